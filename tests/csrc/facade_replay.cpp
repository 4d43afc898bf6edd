// C++ host driving the CUDA path through the facade exactly as main_sim.cpp:128-184 drives the
// reference VIOFilter: processIMUData..., augmentLandmarkStates, processVisionData, stateEstimate.
// Reads a recorded stream (flat doubles, written by tests/test_facade_cpp.py), writes the per-update
// state estimate and Sigma.  Usage: facade_replay <in.bin> <out.bin>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "eqvio_b200_facade.hpp"

using namespace eqvio_b200;

static std::vector<double> slurp(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<double> v(n / 8);
    if (fread(v.data(), 8, v.size(), f) != v.size()) exit(2);
    fclose(f);
    return v;
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    std::vector<double> in = slurp(argv[1]);
    size_t c = 0;
    auto next = [&]() { return in[c++]; };
    VIOFilter::Settings st;
    st.fastRiccati = 1;
    st.coordinateChoice = (int)next();
    st.measurementNoise = next();
    st.outlierThresholdAbs = next();
    st.outlierThresholdProb = next();
    st.featureRetention = next();
    const int frames = (int)next(), N0 = (int)next();
    VIOState xi0;
    double sensor[23];
    for (double& s : sensor) s = next();
    for (int i = 0; i < 6; ++i) xi0.sensor.inputBias[i] = sensor[i];
    for (int i = 0; i < 4; ++i) xi0.sensor.pose.q[i] = sensor[6 + i];
    for (int i = 0; i < 3; ++i) xi0.sensor.pose.x[i] = sensor[10 + i];
    for (int i = 0; i < 3; ++i) xi0.sensor.velocity[i] = sensor[13 + i];
    for (int i = 0; i < 4; ++i) xi0.sensor.cameraOffset.q[i] = sensor[16 + i];
    for (int i = 0; i < 3; ++i) xi0.sensor.cameraOffset.x[i] = sensor[20 + i];
    xi0.cameraLandmarks.resize(N0);
    for (int i = 0; i < N0; ++i) xi0.cameraLandmarks[i].id = (int)next();
    for (int i = 0; i < N0; ++i)
        for (int a = 0; a < 3; ++a) xi0.cameraLandmarks[i].p[a] = next();
    eqvio_camera cam{};
    cam.model = EQVIO_CAMERA_PINHOLE;
    cam.width = (int)next();
    cam.height = (int)next();
    cam.fx = next();
    cam.fy = next();
    cam.cx = next();
    cam.cy = next();

    FILE* out = fopen(argv[2], "wb");
    if (!out) return 2;
    try {
        VIOFilter filter(xi0, st, 0.0, N0 + 16);
        for (int k = 0; k < frames; ++k) {
            VisionMeasurement y;
            y.stamp = next();
            y.camera = cam;
            const int n = (int)next(), ni = (int)next();
            std::vector<int> ids(n);
            for (int j = 0; j < n; ++j) ids[j] = (int)next();
            for (int j = 0; j < n; ++j) {
                double u = next(), v = next();
                y.camCoordinates[ids[j]] = {u, v};
            }
            VIOState provided;
            provided.cameraLandmarks.resize(n);
            for (int j = 0; j < n; ++j) {
                provided.cameraLandmarks[j].id = ids[j];
                for (int a = 0; a < 3; ++a) provided.cameraLandmarks[j].p[a] = next();
            }
            for (int s = 0; s < ni; ++s) {
                IMUVelocity u;
                u.stamp = next();
                for (int a = 0; a < 3; ++a) u.gyr[a] = next();
                for (int a = 0; a < 3; ++a) u.acc[a] = next();
                for (int a = 0; a < 3; ++a) u.gyrBiasVel[a] = next();
                for (int a = 0; a < 3; ++a) u.accBiasVel[a] = next();
                filter.processIMUData(u);
            }
            filter.augmentLandmarkStates(y.getIds(), provided);
            filter.processVisionData(y);
            VIOState est = filter.stateEstimate();
            EqFStateView view = filter.viewEqFState();
            std::vector<double> rec;
            rec.push_back((double)est.cameraLandmarks.size());
            rec.push_back(filter.getTime());
            rec.insert(rec.end(), est.sensor.inputBias.begin(), est.sensor.inputBias.end());
            rec.insert(rec.end(), est.sensor.pose.q.begin(), est.sensor.pose.q.end());
            rec.insert(rec.end(), est.sensor.pose.x.begin(), est.sensor.pose.x.end());
            rec.insert(rec.end(), est.sensor.velocity.begin(), est.sensor.velocity.end());
            rec.insert(rec.end(), est.sensor.cameraOffset.q.begin(), est.sensor.cameraOffset.q.end());
            rec.insert(rec.end(), est.sensor.cameraOffset.x.begin(), est.sensor.cameraOffset.x.end());
            for (const auto& lm : est.cameraLandmarks) rec.push_back((double)lm.id);
            for (const auto& lm : est.cameraLandmarks) rec.insert(rec.end(), lm.p.begin(), lm.p.end());
            rec.insert(rec.end(), view.Sigma.begin(), view.Sigma.end());
            fwrite(rec.data(), 8, rec.size(), out);
        }
    } catch (const Error& e) {
        fprintf(stderr, "eqvio error %d: %s\n", e.code, e.what());
        fclose(out);
        return 1;
    }
    fclose(out);
    return 0;
}
