// Test-only host build of the shared host/device arithmetic (lie.cuh, model.cuh and the
// sensor-sized preparation of kernels.cuh), so that the CPU test-suite can check it against the
// oracle without a GPU.  Not part of the product library.
#include "../../eqvio_b200/csrc/kernels.cuh"

using namespace eqvio;

extern "C" {

// ctx_out: RiccatiCtx as doubles; steps_out: nsteps * sizeof(ObsStep)/8 doubles; Xs updated in place
int hook_sensor_prep(const double* xi0s, double* Xs, const double* imu, int nsteps, const double* meanImu, double dtTotal,
                     int discreteLift, const double* qdiag, const double* pdiag, double* ctx_out, double* steps_out) {
    PrepArgs a;
    RiccatiCtx ctx;
    ObsStep* steps = new ObsStep[nsteps > 0 ? nsteps : 1];
    a.xi0s = xi0s;
    a.Xs = Xs;
    a.ctx = &ctx;
    a.steps = steps;
    a.imu = imu;
    FrameHeader fr;
    fr.fs.nsteps = nsteps;
    for (int k = 0; k < 12; ++k) fr.fs.meanImu[k] = meanImu[k];
    fr.fs.dtTotal = dtTotal;
    a.fr = &fr;
    a.discreteLift = discreteLift;
    for (int k = 0; k < 4; ++k) a.qdiag[k] = qdiag[k];
    for (int k = 0; k < 8; ++k) a.pdiag[k] = pdiag[k];
    double XsOut[23];
    a.XsOut = XsOut;
    double As[441] = {0.0}, Bs[252] = {0.0};  // zero on entry, as the kernel's shared arrays are
    for (int part = 0; part < 3; ++part) riccati_small_part(a, ctx, As, Bs, part);
    for (int t = 0; t < 441; ++t) {
        double fs, ns;
        riccati_entry(a, As, Bs, t, fs, ns);
    }
    observer_sensor_body(a);
    for (int k = 0; k < 23; ++k) Xs[k] = XsOut[k];
    memcpy(ctx_out, &ctx, sizeof(ctx));
    memcpy(steps_out, steps, nsteps * sizeof(ObsStep));
    delete[] steps;
    return (int)(sizeof(RiccatiCtx) / 8);
}
int hook_sizeof_ctx() { return (int)sizeof(RiccatiCtx); }
int hook_sizeof_step() { return (int)sizeof(ObsStep); }

void hook_output_block(const double* camv, int coord, const double* q0, const double* Q, double Qa, int useStar, double yu,
                       double yv, double* C) {
    Camera cam;
    cam.model = (int)camv[0];
    cam.width = (int)camv[1];
    cam.height = (int)camv[2];
    cam.ndist = (int)camv[3];
    cam.fx = camv[4];
    cam.fy = camv[5];
    cam.cx = camv[6];
    cam.cy = camv[7];
    for (int i = 0; i < 5; ++i) {
        cam.dist[i] = camv[8 + i];
        cam.inv_dist[i] = camv[13 + i];
    }
    output_block(cam, coord, V3{q0[0], q0[1], q0[2]}, Quat{Q[0], Q[1], Q[2], Q[3]}, Qa, useStar != 0, yu, yv, C);
}
}
