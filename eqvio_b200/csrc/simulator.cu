// VIOSimulator on the device: IMU samples and vision measurements of the reference's simulator
// (src/VIOSimulator.cpp:128-265, trajectory "wave" of src/dataserver/SimulationDataServer.cpp:46-65) for MANY Monte-Carlo
// instances at once.  The trajectory is the same for every instance, the world points (src/VIOSimulator.cpp:63-126,
// generated on the host from the instance's seed and uploaded once) differ; what is heavy per frame -- transforming and
// testing the visibility of all ~20 N world points, picking the first maxFeatures visible ones in shuffled order, sorting
// them by id -- is one CTA per (frame, instance).  Noise-free streams only (the reference's default, VIOSimulator.h:38-40).
//
// Everything is a pure function of the time stamp: pose samples are recomputed from the closed form where the reference
// reads its pose table, with the same expressions (3.14, not pi; t0 = i / 2000 + 0.0025; times t0 - 0.0025).
#include <cuda_runtime.h>

#include <string>

#include "../../include/eqvio_b200_sim.h"
#include "lie.cuh"

namespace {
using namespace eqvio;

constexpr double TRAJ_FREQUENCY = 10.0 * 200.0;  // 10 * max(imuFreq 200, imageFreq 20): the defaults are in force when the
constexpr double TRAJ_T0 = 0.5 / 200.0;          // trajectory is generated (SimulationDataServer.cpp:223-232, :145)
constexpr int SIM_THREADS = 256;
constexpr int SIM_MAX_FEATURES = 2048;

std::string g_simError;

struct SimParams {
    int numPoints, maxFeatures, numPoses;
    double fx, fy, cx, cy;
    int width, height;
    SE3 camOffset;
};

// Measurement / input noise of the reference's simulator (VIOSimulator.cpp:163-167: IMU, zero-mean Gaussian with covariance
// constructInputGainMatrix() * samplingFrequency; :258-262: pixels, constructOutputGainMatrix(n) = measurementNoise^2 I).
// The reference draws from a std::normal_distribution stream that no other platform reproduces; here every draw is a pure
// function of (instance seed, stream, event index, component) through the counter-based Philox4x32-10 generator and a
// Box-Muller transform, so that any number of Monte-Carlo instances is generated in one launch, in any order, reproducibly
// (simdata/philox.py is the host statement of the same function).
struct NoiseParams {
    int input, output;          // switches (SimulationDataServer.cpp:229-230)
    double imuSigma[12];        // per component: sqrt(variance * samplingFrequency)
    double pixelSigma;
    double imuFreq, imageFreq;  // event index = round(stamp * frequency)
};
enum { STREAM_IMU = 1, STREAM_VISION = 2 };

__host__ __device__ inline void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned (&out)[4]) {
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
        c0 = n0, c1 = n1, c2 = n2, c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}
// two independent N(0, 1) draws for (seed, stream, event, component)
__host__ __device__ inline void normal_pair(unsigned long long seed, int stream, long long event, int comp, double& z0, double& z1) {
    unsigned x[4];
    philox4x32_10((unsigned)event, (unsigned)((unsigned long long)event >> 32), (unsigned)comp, (unsigned)stream, (unsigned)seed,
                  (unsigned)(seed >> 32), x);
    const unsigned long long a = (((unsigned long long)x[0] << 32) | x[1]) >> 11, b = (((unsigned long long)x[2] << 32) | x[3]) >> 11;
    const double u1 = ((double)a + 1.0) * (1.0 / 9007199254740992.0);  // (0, 1]
    const double u2 = (double)b * (1.0 / 9007199254740992.0);          // [0, 1)
    const double r = sqrt(-2.0 * log(u1)), th = 6.283185307179586476925286766559 * u2;
    z0 = r * cos(th);
    z1 = r * sin(th);
}

__host__ __device__ inline double traj_time(int i) { return ((double)i / TRAJ_FREQUENCY + TRAJ_T0) - TRAJ_T0; }
__host__ __device__ inline void traj_pose(int i, Quat& q, V3& x) {
    const double t0 = (double)i / TRAJ_FREQUENCY + TRAJ_T0;
    const double angle = 3.14 * 2 * t0 / 20.0;
    q = so3_exp(V3{0.0, 0.0, angle});
    x = V3{cos(angle), sin(angle), 0.2 * sin(10 * angle)};
}
// number of pose times strictly below t (numpy searchsorted, side = left)
__host__ __device__ inline int time_index(double t, int M) {
    int i = (int)floor(t * TRAJ_FREQUENCY) - 2;
    if (i < 0) i = 0;
    if (i > M) i = M;
    while (i > 0 && !(traj_time(i - 1) < t)) --i;
    while (i < M && traj_time(i) < t) ++i;
    return i;
}
__host__ __device__ inline int clamp_index(int it, int M) {  // VIOSimulator.cpp:131-138
    while (it + 1 >= M) --it;
    while (it - 2 <= 0) ++it;
    return it;
}
// 4x4 inverse by Gauss-Jordan with partial pivoting (the Gram matrix of the cubic fit is badly scaled, not ill-posed)
__host__ __device__ inline void inv4(const double* A, double* Ai) {
    double a[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            a[r][c] = A[4 * r + c];
            a[r][4 + c] = r == c ? 1.0 : 0.0;
        }
    for (int c = 0; c < 4; ++c) {
        int p = c;
        for (int r = c + 1; r < 4; ++r)
            if (fabs(a[r][c]) > fabs(a[p][c])) p = r;
        if (p != c)
            for (int k = 0; k < 8; ++k) {
                const double t = a[c][k];
                a[c][k] = a[p][k];
                a[p][k] = t;
            }
        const double d = a[c][c];
        for (int k = 0; k < 8; ++k) a[c][k] /= d;
        for (int r = 0; r < 4; ++r)
            if (r != c) {
                const double m = a[r][c];
                for (int k = 0; k < 8; ++k) a[r][k] -= m * a[c][k];
            }
    }
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) Ai[4 * r + c] = a[r][4 + c];
}
// getInertialStates (VIOSimulator.cpp:140-161): cubic through four pose samples around `it`, columns position / velocity /
// acceleration at time ct
__host__ __device__ inline void inertial_states(int it, double ct, V3& pos, V3& vel, V3& acc) {
    double tau[4], P[3][4];
    for (int k = 0; k < 4; ++k) {
        Quat q;
        V3 x;
        traj_pose(it - 2 + k, q, x);
        tau[k] = traj_time(it - 2 + k) - ct;
        P[0][k] = x.x;
        P[1][k] = x.y;
        P[2][k] = x.z;
    }
    double T[4][4], G[16], Gi[16];
    for (int k = 0; k < 4; ++k) {
        T[0][k] = 1.0;
        T[1][k] = tau[k];
        T[2][k] = tau[k] * tau[k] / 2.0;
        T[3][k] = tau[k] * tau[k] * tau[k] / 6.0;
    }
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += T[a][k] * T[b][k];
            G[4 * a + b] = s;
        }
    inv4(G, Gi);
    double out[3][3];
    for (int r = 0; r < 3; ++r) {
        double pt[4];
        for (int a = 0; a < 4; ++a) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += P[r][k] * T[a][k];
            pt[a] = s;
        }
        for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int a = 0; a < 4; ++a) s += pt[a] * Gi[4 * a + c];
            out[r][c] = s;
        }
    }
    pos = V3{out[0][0], out[1][0], out[2][0]};
    vel = V3{out[0][1], out[1][1], out[2][1]};
    acc = V3{out[0][2], out[1][2], out[2][2]};
}

// IMU sample at time t (VIOSimulator.cpp:163-214): row = stamp, gyr3, acc3, gyrBiasVel3 (0), accBiasVel3 (0)
// blockIdx.y = instance (one row block per instance when noise is on; a single block otherwise)
__global__ void sim_imu_kernel(const double* __restrict__ stamps, int n, int M, double* __restrict__ rows, NoiseParams np,
                               const unsigned long long* __restrict__ seeds) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double t = stamps[k];
    double* o = rows + 13 * ((size_t)blockIdx.y * n + k);
    auto add_noise = [&]() {  // VIOSimulator.cpp:163-167
        if (!np.input || !seeds) return;
        const long long ev = llround(t * np.imuFreq);
        for (int c = 0; c < 6; ++c) {
            double z0, z1;
            normal_pair(seeds[blockIdx.y], STREAM_IMU, ev, c, z0, z1);
            o[1 + 2 * c] += np.imuSigma[2 * c] * z0;
            o[2 + 2 * c] += np.imuSigma[2 * c + 1] * z1;
        }
    };
    for (int j = 0; j < 13; ++j) o[j] = 0.0;
    o[0] = t;
    int it = time_index(t, M);
    if (it == M) {  // past the end of the trajectory: at rest, gravity only
        Quat q;
        V3 x;
        traj_pose(M - 1, q, x);
        const V3 a = qrot(qinv(q), V3{0, 0, GRAVITY_CONSTANT});
        o[4] = a.x; o[5] = a.y; o[6] = a.z;
        add_noise();
        return;
    }
    it = clamp_index(it, M);
    Quat q1, q2;
    V3 x1, x2;
    traj_pose(it - 1, q1, x1);
    traj_pose(it, q2, x2);
    const double t1 = traj_time(it - 1), t2 = traj_time(it);
    const V3 gyr = so3_log(qmul(qinv(q1), q2)) / (t2 - t1);
    const Quat att = qmul(q1, so3_exp((t - t1) * gyr));
    V3 p, v, a;
    inertial_states(it, t, p, v, a);
    const V3 acc = qrot(qinv(att), a - V3{0, 0, -GRAVITY_CONSTANT});
    o[1] = gyr.x; o[2] = gyr.y; o[3] = gyr.z;
    o[4] = acc.x; o[5] = acc.y; o[6] = acc.z;
    add_noise();
}

__device__ __forceinline__ bool in_domain(const SimParams& sp, V3 pc, double& u, double& v) {
    u = sp.fx * pc.x / pc.z + sp.cx;
    v = sp.fy * pc.y / pc.z + sp.cy;
    return u >= 0 && v >= 0 && u < sp.width && v < sp.height && pc.z > 0;
}

// One CTA per (frame, instance): getMeasurements (VIOSimulator.cpp:216-265) + the true camera-frame positions of the
// measured points and the true sensor state (getFullState, :269-310) at the same stamp.
__global__ void __launch_bounds__(SIM_THREADS)
    sim_vision_kernel(SimParams sp, const double* __restrict__ stamps, const double* __restrict__ points, const int* __restrict__ pointIds,
                      int* __restrict__ nOut, int* __restrict__ idsOut, double* __restrict__ yOut, double* __restrict__ pOut, NoiseParams np,
                      const unsigned long long* __restrict__ seeds,
                      double* __restrict__ sensorOut) {
    __shared__ SE3 sCi, sCi2;
    __shared__ int sEmpty;
    __shared__ int sKey[SIM_MAX_FEATURES], sIdx[SIM_MAX_FEATURES];
    __shared__ int sWarp[SIM_THREADS / 32], sBase;
    const int frame = blockIdx.x, inst = blockIdx.y, nFrames = gridDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double t = stamps[frame];
    const double* P = points + (size_t)inst * sp.numPoints * 3;
    const int* ids = pointIds + (size_t)inst * sp.numPoints;
    const size_t slot = (size_t)inst * nFrames + frame;
    if (tid == 0) {
        const int M = sp.numPoses;
        int it = time_index(t, M);
        sEmpty = it == M;
        if (!sEmpty) {
            // pose at the image stamp: SE(3) geodesic between the two neighbouring samples (:224-233)
            int iv = it;
            while (iv - 1 < 0) ++iv;
            SE3 p0, p1;
            traj_pose(iv - 1, p0.q, p0.x);
            traj_pose(iv, p1.q, p1.x);
            double vel[6];
            se3_log(se3_mul(se3_inv(p0), p1), vel);
            const double dt = traj_time(iv) - traj_time(iv - 1), s = t - traj_time(iv - 1);
            for (int k = 0; k < 6; ++k) vel[k] = vel[k] / dt * s;
            const SE3 cur = se3_mul(p0, se3_exp(V3{vel[0], vel[1], vel[2]}, V3{vel[3], vel[4], vel[5]}));
            sCi = se3_inv(se3_mul(cur, sp.camOffset));
            // true state (:269-310): attitude by so(3) interpolation, position / velocity from the cubic fit
            const int ic = clamp_index(it, M);
            Quat q0, q1;
            V3 x0, x1;
            traj_pose(ic - 1, q0, x0);
            traj_pose(ic, q1, x1);
            const double ta = traj_time(ic - 1), tb = traj_time(ic);
            const V3 w = so3_log(qmul(qinv(q0), q1)) / (tb - ta);
            const Quat pq = qmul(q0, so3_exp((t - ta) * w));
            V3 px, pv, pa;
            inertial_states(ic, t, px, pv, pa);
            const V3 vel_b = qrot(qinv(pq), pv);
            sCi2 = se3_inv(se3_mul(SE3{pq, px}, sp.camOffset));
            if (inst == 0 && sensorOut) {
                double* o = sensorOut + 23 * (size_t)frame;
                for (int k = 0; k < 6; ++k) o[k] = 0.0;
                o[6] = pq.w; o[7] = pq.x; o[8] = pq.y; o[9] = pq.z;
                o[10] = px.x; o[11] = px.y; o[12] = px.z;
                o[13] = vel_b.x; o[14] = vel_b.y; o[15] = vel_b.z;
                o[16] = sp.camOffset.q.w; o[17] = sp.camOffset.q.x; o[18] = sp.camOffset.q.y; o[19] = sp.camOffset.q.z;
                o[20] = sp.camOffset.x.x; o[21] = sp.camOffset.x.y; o[22] = sp.camOffset.x.z;
            }
        }
        sBase = 0;
    }
    __syncthreads();
    if (sEmpty) {
        if (tid == 0) nOut[slot] = 0;
        return;
    }
    const SE3 ci = sCi;
    // first maxFeatures visible points in stored (shuffled) order: ordered compaction, 256 points per round
    for (int p0 = 0; p0 < sp.numPoints; p0 += SIM_THREADS) {
        const int base = sBase;
        if (base >= sp.maxFeatures) break;
        const int p = p0 + tid;
        bool vis = false;
        if (p < sp.numPoints) {
            double u, v;
            vis = in_domain(sp, se3_apply(ci, V3{P[3 * p], P[3 * p + 1], P[3 * p + 2]}), u, v);
        }
        const unsigned m = __ballot_sync(0xffffffffu, vis);
        if (lane == 0) sWarp[warp] = __popc(m);
        __syncthreads();
        int off = 0, total = 0;
        for (int w = 0; w < SIM_THREADS / 32; ++w) {
            if (w < warp) off += sWarp[w];
            total += sWarp[w];
        }
        const int pos = base + off + __popc(m & ((1u << lane) - 1));
        if (vis && pos < sp.maxFeatures) {
            sKey[pos] = ids[p];
            sIdx[pos] = p;
        }
        __syncthreads();
        if (tid == 0) sBase = base + total;
        __syncthreads();
    }
    const int count = min(sBase, sp.maxFeatures);
    // sort by id (ascending, ids are unique): bitonic network over the next power of two
    int n2 = 1;
    while (n2 < count) n2 <<= 1;
    for (int i = count + tid; i < n2; i += SIM_THREADS) {
        sKey[i] = 0x7fffffff;
        sIdx[i] = -1;
    }
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n2; i += SIM_THREADS) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const int a = sKey[i], b = sKey[l];
                    if ((a > b) == up) {
                        sKey[i] = b;
                        sKey[l] = a;
                        const int ti = sIdx[i];
                        sIdx[i] = sIdx[l];
                        sIdx[l] = ti;
                    }
                }
            }
            __syncthreads();
        }
    if (tid == 0) nOut[slot] = count;
    const SE3 ci2 = sCi2;
    for (int i = tid; i < count; i += SIM_THREADS) {
        const int p = sIdx[i];
        const V3 pw = V3{P[3 * p], P[3 * p + 1], P[3 * p + 2]};
        double u, v;
        in_domain(sp, se3_apply(ci, pw), u, v);
        const V3 pt = se3_apply(ci2, pw);
        const size_t o = slot * sp.maxFeatures + i;
        idsOut[o] = sKey[i];
        if (np.output && seeds) {  // VIOSimulator.cpp:258-262: after the visibility test, one pair of draws per measured point (ascending id)
            double z0, z1;
            normal_pair(seeds[blockIdx.y], STREAM_VISION, llround(stamps[blockIdx.x] * np.imageFreq), i, z0, z1);
            u += np.pixelSigma * z0;
            v += np.pixelSigma * z1;
        }
        yOut[2 * o] = u;
        yOut[2 * o + 1] = v;
        pOut[3 * o] = pt.x;
        pOut[3 * o + 1] = pt.y;
        pOut[3 * o + 2] = pt.z;
    }
}
}  // namespace

struct eqvio_sim {
    int device = 0, nInst = 0;
    SimParams sp;
    double* d_points = nullptr;
    int* d_ids = nullptr;
    unsigned long long* d_seeds = nullptr;
    NoiseParams np = {};
    cudaStream_t stream = nullptr;
};

extern "C" {

const char* eqvio_sim_last_error(void) { return g_simError.c_str(); }

eqvio_sim* eqvio_sim_create(int device, int n_instances, int num_points, int max_features, double duration, const double intrinsics[4],
                            int width, int height, const double* points, const int* point_ids) {
    if (n_instances <= 0 || num_points <= 0 || max_features <= 0 || max_features > SIM_MAX_FEATURES || !intrinsics || !points || !point_ids) {
        g_simError = "eqvio_sim_create: invalid arguments (max_features <= 2048)";
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        g_simError = "eqvio_sim_create: no such CUDA device";
        return nullptr;
    }
    eqvio_sim* s = new eqvio_sim();
    s->device = device;
    s->nInst = n_instances;
    s->sp.numPoints = num_points;
    s->sp.maxFeatures = max_features;
    s->sp.numPoses = (int)floor(duration * TRAJ_FREQUENCY);
    s->sp.fx = intrinsics[0];
    s->sp.fy = intrinsics[1];
    s->sp.cx = intrinsics[2];
    s->sp.cy = intrinsics[3];
    s->sp.width = width;
    s->sp.height = height;
    // camera offset R = [[0,0,1],[-1,0,0],[0,-1,0]], x = 0 (SimulationDataServer.cpp:234-236)
    M3 R = M3{{0, 0, 1, -1, 0, 0, 0, -1, 0}};
    s->sp.camOffset = SE3{mat2quat(R), V3{0, 0, 0}};
    const size_t np = (size_t)n_instances * num_points;
    bool ok = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMalloc(&s->d_points, np * 3 * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&s->d_ids, np * sizeof(int)) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(s->d_points, points, np * 3 * sizeof(double), cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(s->d_ids, point_ids, np * sizeof(int), cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(s->stream) == cudaSuccess;
    if (!ok) {
        g_simError = std::string("eqvio_sim_create: ") + cudaGetErrorString(cudaGetLastError());
        eqvio_sim_destroy(s);
        return nullptr;
    }
    return s;
}

void eqvio_sim_destroy(eqvio_sim* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    cudaFree(s->d_points);
    cudaFree(s->d_ids);
    cudaFree(s->d_seeds);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int eqvio_sim_set_noise(eqvio_sim* s, int input_noise, int output_noise, const double imu_sigma[12], double pixel_sigma, double imu_freq,
                        double image_freq, const unsigned long long* seeds) {
    if (!s || ((input_noise || output_noise) && (!seeds || !imu_sigma))) return -1;
    cudaSetDevice(s->device);
    s->np.input = input_noise ? 1 : 0;
    s->np.output = output_noise ? 1 : 0;
    for (int i = 0; i < 12; ++i) s->np.imuSigma[i] = imu_sigma ? imu_sigma[i] : 0.0;
    s->np.pixelSigma = pixel_sigma;
    s->np.imuFreq = imu_freq;
    s->np.imageFreq = image_freq;
    if (seeds) {
        if (!s->d_seeds && cudaMalloc(&s->d_seeds, s->nInst * sizeof(unsigned long long)) != cudaSuccess) return -2;
        if (cudaMemcpyAsync(s->d_seeds, seeds, s->nInst * sizeof(unsigned long long), cudaMemcpyHostToDevice, s->stream) != cudaSuccess ||
            cudaStreamSynchronize(s->stream) != cudaSuccess)
            return -2;
    }
    return 0;
}

static int sim_imu_impl(eqvio_sim* s, int n, const double* stamps, double* rows, int per_instance) {
    if (!s || n < 0 || (n > 0 && (!stamps || !rows))) return -1;
    if (n == 0) return 0;
    cudaSetDevice(s->device);
    const int blocksY = per_instance ? s->nInst : 1;
    const size_t total = (size_t)blocksY * n * 13;
    double *d_t = nullptr, *d_r = nullptr;
    bool ok = cudaMalloc(&d_t, n * sizeof(double)) == cudaSuccess && cudaMalloc(&d_r, total * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(d_t, stamps, n * sizeof(double), cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    if (ok)
        sim_imu_kernel<<<dim3((n + 127) / 128, blocksY), 128, 0, s->stream>>>(d_t, n, s->sp.numPoses, d_r, s->np,
                                                                             per_instance ? s->d_seeds : (const unsigned long long*)nullptr);
    ok = ok && cudaGetLastError() == cudaSuccess;
    ok = ok && cudaMemcpyAsync(rows, d_r, total * sizeof(double), cudaMemcpyDeviceToHost, s->stream) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(s->stream) == cudaSuccess;
    if (!ok) g_simError = std::string("eqvio_sim_imu: ") + cudaGetErrorString(cudaGetLastError());
    cudaFree(d_t);
    cudaFree(d_r);
    return ok ? 0 : -2;
}
int eqvio_sim_imu(eqvio_sim* s, int n, const double* stamps, double* rows) { return sim_imu_impl(s, n, stamps, rows, 0); }
int eqvio_sim_imu_instances(eqvio_sim* s, int n, const double* stamps, double* rows) { return sim_imu_impl(s, n, stamps, rows, 1); }

int eqvio_sim_vision(eqvio_sim* s, int n, const double* stamps, int* n_out, int* ids, double* y, double* provided_p, double* true_sensor,
                     float* device_ms) {
    if (!s || n < 0 || (n > 0 && (!stamps || !n_out || !ids || !y || !provided_p))) return -1;
    if (n == 0) return 0;
    cudaSetDevice(s->device);
    const size_t slots = (size_t)s->nInst * n, feat = slots * s->sp.maxFeatures;
    double *d_t = nullptr, *d_y = nullptr, *d_p = nullptr, *d_s = nullptr;
    int *d_n = nullptr, *d_i = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool ok = cudaMalloc(&d_t, n * sizeof(double)) == cudaSuccess && cudaMalloc(&d_n, slots * sizeof(int)) == cudaSuccess &&
              cudaMalloc(&d_i, feat * sizeof(int)) == cudaSuccess && cudaMalloc(&d_y, feat * 2 * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&d_p, feat * 3 * sizeof(double)) == cudaSuccess && cudaMalloc(&d_s, (size_t)n * 23 * sizeof(double)) == cudaSuccess;
    ok = ok && cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(d_t, stamps, n * sizeof(double), cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    ok = ok && cudaMemsetAsync(d_i, 0, feat * sizeof(int), s->stream) == cudaSuccess;
    ok = ok && cudaMemsetAsync(d_y, 0, feat * 2 * sizeof(double), s->stream) == cudaSuccess;
    ok = ok && cudaMemsetAsync(d_p, 0, feat * 3 * sizeof(double), s->stream) == cudaSuccess;
    ok = ok && cudaMemsetAsync(d_s, 0, (size_t)n * 23 * sizeof(double), s->stream) == cudaSuccess;
    if (ok) {
        cudaEventRecord(e0, s->stream);
        sim_vision_kernel<<<dim3(n, s->nInst), SIM_THREADS, 0, s->stream>>>(s->sp, d_t, s->d_points, s->d_ids, d_n, d_i, d_y, d_p, s->np,
                                                                            (const unsigned long long*)s->d_seeds, d_s);
        cudaEventRecord(e1, s->stream);
    }
    ok = ok && cudaGetLastError() == cudaSuccess;
    ok = ok && cudaMemcpyAsync(n_out, d_n, slots * sizeof(int), cudaMemcpyDeviceToHost, s->stream) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(ids, d_i, feat * sizeof(int), cudaMemcpyDeviceToHost, s->stream) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(y, d_y, feat * 2 * sizeof(double), cudaMemcpyDeviceToHost, s->stream) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(provided_p, d_p, feat * 3 * sizeof(double), cudaMemcpyDeviceToHost, s->stream) == cudaSuccess;
    if (true_sensor) ok = ok && cudaMemcpyAsync(true_sensor, d_s, (size_t)n * 23 * sizeof(double), cudaMemcpyDeviceToHost, s->stream) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(s->stream) == cudaSuccess;
    if (ok && device_ms) cudaEventElapsedTime(device_ms, e0, e1);
    if (!ok) g_simError = std::string("eqvio_sim_vision: ") + cudaGetErrorString(cudaGetLastError());
    cudaFree(d_t);
    cudaFree(d_n);
    cudaFree(d_i);
    cudaFree(d_y);
    cudaFree(d_p);
    cudaFree(d_s);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return ok ? 0 : -2;
}

}  // extern "C"
