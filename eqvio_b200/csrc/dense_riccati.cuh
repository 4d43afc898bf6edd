// integrateRiccatiStateAccurate (src/mathematical/VIO_eqf.cpp:74-91) -- the reference's DEFAULT propagation
// (fastRiccati = false, useDiscreteStateMatrix = false), executed once per buffered IMU sample:
//
//     [A0tExp BtExp; 0 I] = exp( dt [A0t Bt; 0 0] ),
//     Sigma <- A0tExp Sigma A0tExp^T + BtExp (Q / dt) BtExp^T + dt P.
//
// This variant is a dense path on purpose: the matrix exponential is taken with a Pade-13 scaling-and-squaring
// scheme (Higham 2005, the algorithm behind Eigen's unsupported MatrixFunctions `.exp()`), whose products and LU
// solve are plain library calls (cuBLAS DGEMM / cuSOLVER getrf+getrs, resolved with dlopen at first use so that the
// library has no link-time dependency on them).  The hand-written kernels here only assemble dt [A B; 0 0] from
// the same per-landmark blocks the fast path uses, combine matrices, and move Sigma between its padded device
// layout and the dense one.  A structured version (one 36x36 exponential per landmark) is the natural next step.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include "kernels.cuh"

namespace eqvio {
namespace dense {

// ---- minimal cuBLAS / cuSOLVER surface, resolved at run time ---------------------------------------------------
typedef void* blasHandle;
typedef void* solverHandle;
struct Libs {
    void *hBlas = nullptr, *hSolver = nullptr;
    int (*blasCreate)(blasHandle*) = nullptr;
    int (*blasDestroy)(blasHandle) = nullptr;
    int (*blasSetStream)(blasHandle, cudaStream_t) = nullptr;
    int (*dgemm)(blasHandle, int, int, int, int, int, const double*, const double*, int, const double*, int, const double*, double*,
                 int) = nullptr;
    int (*solverCreate)(solverHandle*) = nullptr;
    int (*solverDestroy)(solverHandle) = nullptr;
    int (*solverSetStream)(solverHandle, cudaStream_t) = nullptr;
    int (*getrfBuf)(solverHandle, int, int, double*, int, int*) = nullptr;
    int (*getrf)(solverHandle, int, int, double*, int, double*, int*, int*) = nullptr;
    int (*getrs)(solverHandle, int, int, int, const double*, int, const int*, double*, int, int*) = nullptr;
    bool ok = false;
    const char* why = "";
};
inline void* open_any(const char* a, const char* b) {
    void* h = dlopen(a, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen(b, RTLD_NOW | RTLD_GLOBAL);
    return h;
}
inline Libs& libs() {
    static Libs L;
    static bool tried = false;
    if (tried) return L;
    tried = true;
    L.hBlas = open_any("libcublas.so.12", "/usr/local/cuda/lib64/libcublas.so.12");
    L.hSolver = open_any("libcusolver.so.11", "/usr/local/cuda/lib64/libcusolver.so.11");
    if (!L.hBlas || !L.hSolver) {
        L.why = "libcublas.so.12 / libcusolver.so.11 not found";
        return L;
    }
#define EQ_SYM(field, handle, name)                                  \
    *(void**)(&L.field) = dlsym(handle, name);                       \
    if (!L.field) {                                                  \
        L.why = "missing symbol " name;                              \
        return L;                                                    \
    }
    EQ_SYM(blasCreate, L.hBlas, "cublasCreate_v2")
    EQ_SYM(blasDestroy, L.hBlas, "cublasDestroy_v2")
    EQ_SYM(blasSetStream, L.hBlas, "cublasSetStream_v2")
    EQ_SYM(dgemm, L.hBlas, "cublasDgemm_v2")
    EQ_SYM(solverCreate, L.hSolver, "cusolverDnCreate")
    EQ_SYM(solverDestroy, L.hSolver, "cusolverDnDestroy")
    EQ_SYM(solverSetStream, L.hSolver, "cusolverDnSetStream")
    EQ_SYM(getrfBuf, L.hSolver, "cusolverDnDgetrf_bufferSize")
    EQ_SYM(getrf, L.hSolver, "cusolverDnDgetrf")
    EQ_SYM(getrs, L.hSolver, "cusolverDnDgetrs")
#undef EQ_SYM
    L.ok = true;
    return L;
}
constexpr int OP_N = 0, OP_T = 1;  // cublasOperation_t

// ---- kernels -------------------------------------------------------------------------------------------------------
// M = dt [A B; 0 0] (n x n column-major, n = dim + 12).  Sensor rows come from the Riccati context of the current
// sample (F_s = I + dt A_s, and dt q_gyr B_s[:, 0:3] is NOT enough for B, so the raw dt B_s is passed in dtBs).
__global__ void dense_fill_sensor_kernel(double* __restrict__ M, int n, int dim, const RiccatiCtx* __restrict__ ctx,
                                         const double* __restrict__ dtBs /*21x12 row-major*/) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 441) {
        const int r = t / 21, c = t % 21;
        M[(size_t)c * n + r] = ctx->Fs[t] - (r == c ? 1.0 : 0.0);
    }
    if (t < 252) {
        const int r = t / 12, c = t % 12;
        M[(size_t)(dim + c) * n + r] = dtBs[t];
    }
}
// landmark rows from rows[i] = D(9) | G(36) | Bl(9): D = I + dt A_q, G = dt [-B_l | A_vel | A_cam] on columns c_sidx
__global__ void dense_fill_landmark_kernel(double* __restrict__ M, int n, int dim, int N, const double* __restrict__ rows, double dt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double* ro = rows + (size_t)i * ROWS_STRIDE;
    const int r0 = SENSOR_DIM + 3 * i;
    for (int r = 0; r < 3; ++r) {
        for (int k = 0; k < 12; ++k) M[(size_t)c_sidx[k] * n + r0 + r] = ro[9 + 12 * r + k];
        for (int c = 0; c < 3; ++c) M[(size_t)(r0 + c) * n + r0 + r] = ro[3 * r + c] - (r == c ? 1.0 : 0.0);
        for (int c = 0; c < 3; ++c) M[(size_t)(dim + c) * n + r0 + r] = dt * ro[45 + 3 * r + c];
    }
}
// ---- useDiscreteStateMatrix: stateMatrixADiscrete (EqFMatrices.cpp:24-41) -------------------------------------------
// A0tD = numericalDifferential(a0Discrete, 0) with central differences of step h = cbrt(eps) (Geometry.cpp:25-36).  The
// sensor coordinates of a0Discrete depend on the sensor coordinates only, landmark i's on the sensor coordinates and its
// own three, so A0tD = [[F_s, 0], [G, blockdiag D_i]] needs 43 sensor evaluations (0, +-h e_j) and 49 per landmark
// instead of 2 dim evaluations of the whole state; every entry outside those blocks is an exact zero in the reference
// too (identical function values cancel).  Written into the top-left dim x dim block of R (n x n column-major).
constexpr int DA_SENSOR_EVALS = 1 + 2 * SENSOR_DIM;  // 43
__device__ __forceinline__ double num_diff_step() { return cbrt(2.220446049250313e-16); }
__global__ void discrete_a_sensor_kernel(int coord, const double* __restrict__ xi0s, const double* __restrict__ Xs,
                                         const double* __restrict__ imuRow, double* __restrict__ R, int n, SE3* __restrict__ ccOut) {
    __shared__ double e1[DA_SENSOR_EVALS][SENSOR_DIM];
    const int k = threadIdx.x;
    const double h = num_diff_step();
    if (k < DA_SENSOR_EVALS) {
        const SensorState xi0 = unpack_sensor(xi0s);
        const GroupSensor X = unpack_group(Xs);
        double eps[SENSOR_DIM];
        for (int j = 0; j < SENSOR_DIM; ++j) eps[j] = 0.0;
        if (k > 0) eps[(k - 1) / 2] = ((k - 1) & 1) ? -h : h;
        SE3 cc;
        double out[SENSOR_DIM];
        a0_discrete_sensor(coord, X, xi0, imuRow + 1, imuRow[0], eps, out, cc);
        for (int j = 0; j < SENSOR_DIM; ++j) e1[k][j] = out[j];
        ccOut[k] = cc;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < SENSOR_DIM * SENSOR_DIM; t += blockDim.x) {
        const int r = t / SENSOR_DIM, j = t % SENSOR_DIM;
        R[(size_t)j * n + r] = (e1[1 + 2 * j][r] - e1[2 + 2 * j][r]) / (2 * h);
    }
}
// one CTA per landmark: 43 evaluations under the sensor perturbations + 6 under its own
__global__ void discrete_a_landmark_kernel(const double* __restrict__ lm, int cap, int N, int coord, const double* __restrict__ imuRow,
                                           const SE3* __restrict__ cc, double* __restrict__ R, int n) {
    __shared__ double o[DA_SENSOR_EVALS + 6][3];
    const int i = blockIdx.x, e = threadIdx.x;
    const double h = num_diff_step();
    if (e < DA_SENSOR_EVALS + 6) {
        const V3 p0 = V3{lm[F_Q0X * cap + i], lm[F_Q0Y * cap + i], lm[F_Q0Z * cap + i]};
        const Quat Q = Quat{lm[F_QW * cap + i], lm[F_QX * cap + i], lm[F_QY * cap + i], lm[F_QZ * cap + i]};
        const double a = lm[F_QA * cap + i];
        V3 eps = V3{0, 0, 0};
        int ks = e;  // sensor evaluation whose camera-frame change applies
        if (e >= DA_SENSOR_EVALS) {
            const int c = (e - DA_SENSOR_EVALS) / 2;
            const double d = ((e - DA_SENSOR_EVALS) & 1) ? -h : h;
            eps = V3{c == 0 ? d : 0.0, c == 1 ? d : 0.0, c == 2 ? d : 0.0};
            ks = 0;
        }
        const V3 r = a0_discrete_landmark(coord, p0, Q, a, eps, cc[ks], cc[0]);
        o[e][0] = r.x;
        o[e][1] = r.y;
        o[e][2] = r.z;
    }
    __syncthreads();
    const int r0 = SENSOR_DIM + 3 * i;
    for (int t = threadIdx.x; t < 3 * SENSOR_DIM; t += blockDim.x) {
        const int r = t / SENSOR_DIM, j = t % SENSOR_DIM;
        R[(size_t)j * n + r0 + r] = (o[1 + 2 * j][r] - o[2 + 2 * j][r]) / (2 * h);
    }
    if (threadIdx.x < 9) {
        const int r = threadIdx.x / 3, c = threadIdx.x % 3;
        R[(size_t)(r0 + c) * n + r0 + r] = (o[DA_SENSOR_EVALS + 2 * c][r] - o[DA_SENSOR_EVALS + 2 * c + 1][r]) / (2 * h);
    }
}
// ---- Normal coordinates: A_normal = M A_euclid M^-1, B_normal = M B_euclid (normal.cpp:37-45) ---------------------------
// M = coordinateDifferential_normal_euclid(xi0) (VIOState.cpp:391-401) is the central-difference derivative of the chart
// change; it is block diagonal (sensor 21 x 21, one 3 x 3 per landmark; every other entry is an exact zero in the reference).
// Sensor block and its inverse (Gauss-Jordan with partial pivoting), row-major 21 x 21 each.
__global__ void normal_m_sensor_kernel(const double* __restrict__ xi0s, double* __restrict__ Ms, double* __restrict__ MsInv) {
    __shared__ double e[2 * SENSOR_DIM][SENSOR_DIM];
    __shared__ double aug[SENSOR_DIM][2 * SENSOR_DIM];
    __shared__ double mult[SENSOR_DIM];
    __shared__ int piv;
    const int k = threadIdx.x;
    const double h = normal_diff_step();
    if (k < 2 * SENSOR_DIM) {
        const SensorState xi0 = unpack_sensor(xi0s);
        double eps[SENSOR_DIM], out[SENSOR_DIM];
        for (int j = 0; j < SENSOR_DIM; ++j) eps[j] = 0.0;
        eps[k / 2] = (k & 1) ? -h : h;
        sensor_chart_normal(sensor_chart_std_inv(eps, xi0), xi0, out);
        for (int j = 0; j < SENSOR_DIM; ++j) e[k][j] = out[j];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < SENSOR_DIM * SENSOR_DIM; t += blockDim.x) {
        const int r = t / SENSOR_DIM, j = t % SENSOR_DIM;
        const double v = (e[2 * j][r] - e[2 * j + 1][r]) / (2 * h);
        Ms[t] = v;
        aug[r][j] = v;
        aug[r][SENSOR_DIM + j] = r == j ? 1.0 : 0.0;
    }
    __syncthreads();
    for (int c = 0; c < SENSOR_DIM; ++c) {
        if (threadIdx.x == 0) {
            int p = c;
            for (int r = c + 1; r < SENSOR_DIM; ++r)
                if (fabs(aug[r][c]) > fabs(aug[p][c])) p = r;
            piv = p;
        }
        __syncthreads();
        if (threadIdx.x < 2 * SENSOR_DIM && piv != c) {
            const double t = aug[c][threadIdx.x];
            aug[c][threadIdx.x] = aug[piv][threadIdx.x];
            aug[piv][threadIdx.x] = t;
        }
        __syncthreads();
        const double d = aug[c][c];
        __syncthreads();
        if (threadIdx.x < 2 * SENSOR_DIM) aug[c][threadIdx.x] /= d;
        if (threadIdx.x < SENSOR_DIM) mult[threadIdx.x] = aug[threadIdx.x][c];  // multipliers, read before column c is eliminated
        __syncthreads();
        if (threadIdx.x < 2 * SENSOR_DIM) {
            const double pc = aug[c][threadIdx.x];
            for (int r = 0; r < SENSOR_DIM; ++r)
                if (r != c) aug[r][threadIdx.x] -= mult[r] * pc;
        }
        __syncthreads();
    }
    for (int t = threadIdx.x; t < SENSOR_DIM * SENSOR_DIM; t += blockDim.x) MsInv[t] = aug[t / SENSOR_DIM][SENSOR_DIM + t % SENSOR_DIM];
}
// In place on T = dt [A B; 0 0] (n x n column-major): blocks (s,s), (l_i,s), (l_i,l_i) of A and the rows of B.
// Block 0 handles the sensor rows, block 1 + i landmark i.
__global__ void normal_transform_kernel(double* __restrict__ T, int n, int dim, int N, const double* __restrict__ lm, int cap,
                                        const double* __restrict__ Ms, const double* __restrict__ MsInv) {
    __shared__ double sM[SENSOR_DIM * SENSOR_DIM], sMi[SENSOR_DIM * SENSOR_DIM], sX[SENSOR_DIM * (SENSOR_DIM + 12)], sY[SENSOR_DIM * SENSOR_DIM];
    const int tid = threadIdx.x;
    for (int t = tid; t < SENSOR_DIM * SENSOR_DIM; t += blockDim.x) {
        sM[t] = Ms[t];
        sMi[t] = MsInv[t];
    }
    if (blockIdx.x == 0) {
        // X = [A_ss | B_s] (21 x 33), row-major in shared memory
        for (int t = tid; t < SENSOR_DIM * (SENSOR_DIM + 12); t += blockDim.x) {
            const int r = t / (SENSOR_DIM + 12), c = t % (SENSOR_DIM + 12);
            sX[t] = T[(size_t)(c < SENSOR_DIM ? c : dim + c - SENSOR_DIM) * n + r];
        }
        __syncthreads();
        // Y = A_ss Ms^-1
        for (int t = tid; t < SENSOR_DIM * SENSOR_DIM; t += blockDim.x) {
            const int r = t / SENSOR_DIM, c = t % SENSOR_DIM;
            double acc = 0.0;
            for (int k = 0; k < SENSOR_DIM; ++k) acc += sX[r * (SENSOR_DIM + 12) + k] * sMi[k * SENSOR_DIM + c];
            sY[t] = acc;
        }
        __syncthreads();
        for (int t = tid; t < SENSOR_DIM * (SENSOR_DIM + 12); t += blockDim.x) {
            const int r = t / (SENSOR_DIM + 12), c = t % (SENSOR_DIM + 12);
            double acc = 0.0;
            if (c < SENSOR_DIM) {
                for (int k = 0; k < SENSOR_DIM; ++k) acc += sM[r * SENSOR_DIM + k] * sY[k * SENSOR_DIM + c];
                T[(size_t)c * n + r] = acc;
            } else {
                for (int k = 0; k < SENSOR_DIM; ++k) acc += sM[r * SENSOR_DIM + k] * sX[k * (SENSOR_DIM + 12) + c];
                T[(size_t)(dim + c - SENSOR_DIM) * n + r] = acc;
            }
        }
        return;
    }
    const int i = blockIdx.x - 1;
    if (i >= N) return;
    const int r0 = SENSOR_DIM + 3 * i;
    __shared__ double sMl[9], sMli[9], sR[3 * (SENSOR_DIM + 3 + 12)], sZ[3 * SENSOR_DIM];
    constexpr int W = SENSOR_DIM + 3 + 12;  // [A_is (21) | A_ii (3) | B_i (12)]
    if (tid == 0) {
        const V3 p0 = V3{lm[F_Q0X * cap + i], lm[F_Q0Y * cap + i], lm[F_Q0Z * cap + i]};
        const M3 M = normal_M_landmark(p0), Mi = inverse(M);
        for (int k = 0; k < 9; ++k) {
            sMl[k] = M.m[k];
            sMli[k] = Mi.m[k];
        }
    }
    for (int t = tid; t < 3 * W; t += blockDim.x) {
        const int r = t / W, c = t % W;
        const int col = c < SENSOR_DIM ? c : (c < SENSOR_DIM + 3 ? r0 + c - SENSOR_DIM : dim + c - SENSOR_DIM - 3);
        sR[t] = T[(size_t)col * n + r0 + r];
    }
    __syncthreads();
    // Z = A_is Ms^-1 (3 x 21)
    for (int t = tid; t < 3 * SENSOR_DIM; t += blockDim.x) {
        const int r = t / SENSOR_DIM, c = t % SENSOR_DIM;
        double acc = 0.0;
        for (int k = 0; k < SENSOR_DIM; ++k) acc += sR[r * W + k] * sMi[k * SENSOR_DIM + c];
        sZ[t] = acc;
    }
    __syncthreads();
    for (int t = tid; t < 3 * W; t += blockDim.x) {
        const int r = t / W, c = t % W;
        double acc = 0.0;
        if (c < SENSOR_DIM) {
            for (int k = 0; k < 3; ++k) acc += sMl[3 * r + k] * sZ[k * SENSOR_DIM + c];
            T[(size_t)c * n + r0 + r] = acc;
        } else if (c < SENSOR_DIM + 3) {
            // M_i A_ii M_i^-1
            const int cc = c - SENSOR_DIM;
            for (int k = 0; k < 3; ++k) {
                double inner = 0.0;
                for (int q = 0; q < 3; ++q) inner += sR[k * W + SENSOR_DIM + q] * sMli[3 * q + cc];
                acc += sMl[3 * r + k] * inner;
            }
            T[(size_t)(r0 + cc) * n + r0 + r] = acc;
        } else {
            for (int k = 0; k < 3; ++k) acc += sMl[3 * r + k] * sR[k * W + c];
            T[(size_t)(dim + c - SENSOR_DIM - 3) * n + r0 + r] = acc;
        }
    }
}
// out = a A + b B + c C + d I  (any of A, B, C may be null)
__global__ void dense_lincomb_kernel(double* __restrict__ out, int n, double a, const double* __restrict__ A, double b,
                                     const double* __restrict__ B, double c, const double* __restrict__ C, double d) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    double v = 0.0;
    if (A) v += a * A[idx];
    if (B) v += b * B[idx];
    if (C) v += c * C[idx];
    if (d != 0.0 && idx / n == idx % n) v += d;
    out[idx] = v;
}
// 1-norm (max column abs sum) -> out[0]; one block per column then a host-side max over n values would need a second
// pass: columns are few thousand at most, so one thread per column and an atomic max on the bit pattern suffices.
__global__ void dense_norm1_kernel(const double* __restrict__ M, int n, unsigned long long* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double s = 0.0;
    for (int r = 0; r < n; ++r) s += fabs(M[(size_t)c * n + r]);
    atomicMax(out, (unsigned long long)__double_as_longlong(s));  // non-negative doubles order like their bit patterns
}
// Sigma (padded internal layout) <-> dense dim x dim
__global__ void dense_unpack_sigma_kernel(const double* __restrict__ D, int ldd, int dim, double* __restrict__ S, int ld) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (r >= dim || c >= dim) return;
    const int ri = r < SENSOR_DIM ? r : r + (SOFF - SENSOR_DIM);
    const int ci = c < SENSOR_DIM ? c : c + (SOFF - SENSOR_DIM);
    S[(size_t)ci * ld + ri] = D[(size_t)c * ldd + r];
}
// out (dim x dim, ld = dim) += E_B diag(q / dt) E_B^T + dt P ; E_B = R[0:dim, dim:dim+12] with leading dimension n.
// Symmetrised on the fly: out is also averaged with its transpose? No -- the reference does not symmetrise either.
__global__ void dense_noise_kernel(double* __restrict__ out, int dim, const double* __restrict__ R, int n, double q0, double q1,
                                   double q2, double q3, double invdt, double dt, const double* __restrict__ pdiag /*8*/) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (r >= dim || c >= dim) return;
    const double q[4] = {q0, q1, q2, q3};
    double s = 0.0;
    for (int k = 0; k < 12; ++k) s += R[(size_t)(dim + k) * n + r] * (q[k / 3] * invdt) * R[(size_t)(dim + k) * n + c];
    if (r == c) s += dt * (r < SENSOR_DIM ? pdiag[r / 3] : pdiag[7]);
    out[(size_t)c * dim + r] += s;
}

// ---- workspace + expm ------------------------------------------------------------------------------------------------
struct Workspace {
    int n = 0;
    double* buf[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double* lwork = nullptr;
    int lworkSize = 0;
    int* ipiv = nullptr;
    int* info = nullptr;
    unsigned long long* norm = nullptr;
    double* dtBs = nullptr;
    double* pdiag = nullptr;
    SE3* cc = nullptr;  // camera-frame changes of the 43 sensor evaluations of stateMatrixADiscrete
    blasHandle blas = nullptr;
    solverHandle solver = nullptr;
    void release() {
        for (auto& b : buf) {
            cudaFree(b);
            b = nullptr;
        }
        cudaFree(lwork);
        cudaFree(ipiv);
        cudaFree(info);
        cudaFree(norm);
        cudaFree(dtBs);
        cudaFree(pdiag);
        cudaFree(cc);
        cc = nullptr;
        lwork = nullptr;
        ipiv = nullptr;
        info = nullptr;
        norm = nullptr;
        dtBs = nullptr;
        pdiag = nullptr;
        if (blas) libs().blasDestroy(blas);
        if (solver) libs().solverDestroy(solver);
        blas = nullptr;
        solver = nullptr;
        n = 0;
    }
};

inline const char* ensure(Workspace& w, int n, cudaStream_t stream) {
    Libs& L = libs();
    if (!L.ok) return L.why;
    if (!w.blas) {
        if (L.blasCreate(&w.blas) != 0) return "cublasCreate failed";
        if (L.solverCreate(&w.solver) != 0) return "cusolverDnCreate failed";
    }
    L.blasSetStream(w.blas, stream);
    L.solverSetStream(w.solver, stream);
    if (n > w.n) {
        cudaStreamSynchronize(stream);
        for (auto& b : w.buf) {
            cudaFree(b);
            b = nullptr;
        }
        cudaFree(w.lwork);
        cudaFree(w.ipiv);
        w.lwork = nullptr;
        w.ipiv = nullptr;
        const int cap = n + n / 8 + 16;
        for (auto& b : w.buf)
            if (cudaMalloc(&b, (size_t)cap * cap * sizeof(double)) != cudaSuccess) return "dense Riccati workspace allocation failed";
        if (cudaMalloc(&w.ipiv, cap * sizeof(int)) != cudaSuccess) return "allocation failed";
        int ls = 0;
        if (L.getrfBuf(w.solver, cap, cap, w.buf[0], cap, &ls) != 0) return "cusolverDnDgetrf_bufferSize failed";
        w.lworkSize = ls;
        if (cudaMalloc(&w.lwork, (size_t)ls * sizeof(double)) != cudaSuccess) return "allocation failed";
        if (!w.info) {
            cudaMalloc(&w.info, sizeof(int));
            cudaMalloc(&w.norm, sizeof(unsigned long long));
            cudaMalloc(&w.dtBs, 252 * sizeof(double));
            cudaMalloc(&w.pdiag, 8 * sizeof(double));
            cudaMalloc(&w.cc, DA_SENSOR_EVALS * sizeof(SE3));
        }
        w.n = cap;
    }
    return nullptr;
}

inline int gemm(Workspace& w, int ta, int tb, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
                double beta, double* C, int ldc) {
    return libs().dgemm(w.blas, ta, tb, m, n, k, &alpha, A, lda, B, ldb, &beta, C, ldc);
}

// R = exp(M) for the n x n matrix in w.buf[0]; result in w.buf[8].  Pade-13 with scaling and squaring.
inline const char* expm(Workspace& w, int n, cudaStream_t stream) {
    static const double b[14] = {64764752532480000.0, 32382376266240000.0, 7771770303897600.0, 1187353796428800.0, 129060195264000.0,
                                 10559470521600.0,    670442572800.0,      33522128640.0,      1323241920.0,       40840800.0,
                                 960960.0,            16380.0,             182.0,              1.0};
    double *M = w.buf[0], *A2 = w.buf[1], *A4 = w.buf[2], *A6 = w.buf[3], *U = w.buf[4], *V = w.buf[5], *T1 = w.buf[6], *T2 = w.buf[7],
           *R = w.buf[8];
    const int blocks = (int)(((size_t)n * n + 255) / 256);
    // scaling: s = max(0, ceil(log2(|M|_1 / theta13)))
    cudaMemsetAsync(w.norm, 0, sizeof(unsigned long long), stream);
    dense_norm1_kernel<<<(n + 127) / 128, 128, 0, stream>>>(M, n, w.norm);
    unsigned long long bits = 0;
    cudaMemcpyAsync(&bits, w.norm, sizeof(bits), cudaMemcpyDeviceToHost, stream);
    if (cudaStreamSynchronize(stream) != cudaSuccess) return "norm kernel failed";
    double norm1;
    memcpy(&norm1, &bits, sizeof(double));
    if (!(norm1 == norm1) || norm1 > 1e300) return "non-finite state matrix";
    int s = 0;
    const double theta13 = 5.371920351148152;
    if (norm1 > theta13) s = (int)ceil(log2(norm1 / theta13));
    if (s > 0) dense_lincomb_kernel<<<blocks, 256, 0, stream>>>(M, n, ldexp(1.0, -s), M, 0, nullptr, 0, nullptr, 0.0);
    if (gemm(w, OP_N, OP_N, n, n, n, 1.0, M, n, M, n, 0.0, A2, n)) return "dgemm failed";
    if (gemm(w, OP_N, OP_N, n, n, n, 1.0, A2, n, A2, n, 0.0, A4, n)) return "dgemm failed";
    if (gemm(w, OP_N, OP_N, n, n, n, 1.0, A4, n, A2, n, 0.0, A6, n)) return "dgemm failed";
    // U = M (A6 (b13 A6 + b11 A4 + b9 A2) + b7 A6 + b5 A4 + b3 A2 + b1 I)
    dense_lincomb_kernel<<<blocks, 256, 0, stream>>>(T1, n, b[13], A6, b[11], A4, b[9], A2, 0.0);
    dense_lincomb_kernel<<<blocks, 256, 0, stream>>>(T2, n, b[7], A6, b[5], A4, b[3], A2, b[1]);
    if (gemm(w, OP_N, OP_N, n, n, n, 1.0, A6, n, T1, n, 1.0, T2, n)) return "dgemm failed";
    if (gemm(w, OP_N, OP_N, n, n, n, 1.0, M, n, T2, n, 0.0, U, n)) return "dgemm failed";
    // V = A6 (b12 A6 + b10 A4 + b8 A2) + b6 A6 + b4 A4 + b2 A2 + b0 I
    dense_lincomb_kernel<<<blocks, 256, 0, stream>>>(T1, n, b[12], A6, b[10], A4, b[8], A2, 0.0);
    dense_lincomb_kernel<<<blocks, 256, 0, stream>>>(V, n, b[6], A6, b[4], A4, b[2], A2, b[0]);
    if (gemm(w, OP_N, OP_N, n, n, n, 1.0, A6, n, T1, n, 1.0, V, n)) return "dgemm failed";
    // (V - U) R = (V + U)
    dense_lincomb_kernel<<<blocks, 256, 0, stream>>>(T1, n, 1.0, V, -1.0, U, 0, nullptr, 0.0);
    dense_lincomb_kernel<<<blocks, 256, 0, stream>>>(R, n, 1.0, V, 1.0, U, 0, nullptr, 0.0);
    if (libs().getrf(w.solver, n, n, T1, n, w.lwork, w.ipiv, w.info)) return "getrf failed";
    if (libs().getrs(w.solver, OP_N, n, n, T1, n, w.ipiv, R, n, w.info)) return "getrs failed";
    for (int k = 0; k < s; ++k) {  // undo the scaling
        if (gemm(w, OP_N, OP_N, n, n, n, 1.0, R, n, R, n, 0.0, T2, n)) return "dgemm failed";
        cudaMemcpyAsync(R, T2, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToDevice, stream);
    }
    return nullptr;
}

}  // namespace dense
}  // namespace eqvio
