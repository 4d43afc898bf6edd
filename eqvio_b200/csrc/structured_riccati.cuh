// The Riccati variants beside the fast Euclid / InvDepth step, hand-written and block-structured (no dense dim x dim matrix, no
// library call):
//   * integrateRiccatiStateAccurate (src/mathematical/VIO_eqf.cpp:74-91) -- the reference's DEFAULT (fastRiccati = false), once per
//     buffered IMU sample:   [A0tExp BtExp; 0 I] = exp(dt [A0t Bt; 0 0]),  Sigma <- A0tExp Sigma A0tExp^T + BtExp (Q/dt) BtExp^T + dt P;
//   * integrateRiccatiStateDiscrete (VIO_eqf.cpp:93-103, useDiscreteStateMatrix) with the numerically differentiated
//     stateMatrixADiscrete (EqFMatrices.cpp:24-41);
//   * the Normal coordinate suite (coordinateSuite/normal.cpp:37-45): A = M A_euclid M^-1, B = M B_euclid, fast or per sample.
//
// All of them share one structure.  With the state ordered (sensor 21 | landmarks 3 N) and the 12 input columns appended,
//     dt [A B; 0 0]  has the rows   sensor:      [ dt A_ss (21) | 0 ... 0 | dt B_s (12) ]
//                                   landmark i:  [ dt A_is (21) | dt A_ii (3, own columns only) | dt B_i (12) ]
// (EqFMatrices / euclid.cpp:99-160,186-233: a landmark's velocity depends on the sensor state and on itself), and every function of
// that matrix used here -- I + M, exp(M), the discrete state matrix, the chart change M_normal (block diagonal: VIOState.cpp:391-401) --
// keeps it:  E = [[F_s, 0], [Gamma, blockdiag D_i]],  Bx = [Bx_s; Bx_i].  So the matrices live as COMPACT BLOCKS
//     Ts / Es : 21 x 33 row-major  [ A-part (21) | B-part (12) ]                      (SW = 33)
//     Tl / El : per landmark 3 x 36 row-major  [ sensor (21) | own (3) | B-part (12) ]  (LW = 36)
// and the propagation is the structured one of kernels.cuh with a dense 3 x 21 coupling:
//     Sigma'_ss = F_s Sigma_ss F_s^T + Bx_s Qn Bx_s^T + dt P_s
//     Sigma'_is = H_i F_s^T + Bx_i Qn Bx_s^T,              H_i = Gamma_i Sigma_ss + D_i Sigma_is
//     Sigma'_ij = D_i Sigma_ij D_j^T + U_i V_j^T (+ dt p_l I),   U_i = [Gamma_i | H_i | Bx_i Qn],  V_i = [D_i Sigma_is | Gamma_i | Bx_i]
// (rank 54 instead of the fast path's 27), HBM-bound like prop_ll_kernel: one read and one write of Sigma per Riccati step.
//
// The exponential.  Ordered (sensor, inputs | landmark i), dt [A B; 0 0] restricted to what landmark i's rows can reach is the block
// lower-triangular 36 x 36 matrix  M_i = [[S, 0], [g_i, d_i]]  with the SHARED 33 x 33 block S = dt [[A_ss, B_s], [0, 0]],
// g_i = dt [A_is | B_i] (3 x 33) and d_i = dt A_ii (3 x 3), and  exp(M_i) = [[exp S, 0], [X_i, exp d_i]]:  the landmark rows of the dim + 12
// exponential are exactly (X_i, exp d_i), its sensor rows the top of exp S.  Scaling and squaring with a degree-12 Taylor polynomial
// (|M / 2^s|_inf <= 1/4: truncation 0.25^13 / 13! = 2e-18) needs no inverse and only the landmark ROWS of the powers:
//     rows of M^(k+1):  R_(k+1) = R_k S + d^k g,   squaring:  X <- X E + F X,  F <- F F   with E = exp(S / 2^(s-t)) from the sensor kernel,
// i.e. ~13 products of a 3 x 33 block with the shared 33 x 33 matrix per landmark (45 k FMAs) instead of a dense Pade-13 of order
// dim + 12.  The scaling s comes from a device-side norm (max over the rows of every M_i), so the per-sample loop has no host sync.
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace eqvio {
namespace sric {

constexpr int SW = 33;             // sensor block row: 21 state columns | 12 input columns
constexpr int LW = 36;             // landmark block row: 21 sensor columns | 3 own columns | 12 input columns
constexpr int LSTRIDE = 3 * LW;    // doubles per landmark block
constexpr int SSIZE = SENSOR_DIM * SW;
constexpr int GRANK = 54;          // columns of U_i, V_i
constexpr int GUV_STRIDE = 2 * 3 * GRANK;
constexpr int EXP_DEG = 12;
constexpr int EXP_MAX_SQ = 30;     // |M|_inf up to 2^28: anything beyond is not a state matrix
__host__ __device__ __forceinline__ double exp_theta() { return 0.25; }

struct NoiseArgs {   // settings constants; the step dt comes from the device-side Riccati context (graph replays: kernel arguments are frozen)
    double q[4];    // diagonal of the input gain per 3-block (VIOFilterSettings.h:192-201); applied to Bx as q / dt -- Bx carries dt B
                    // in the one-step and discrete variants, BtExp in the accurate one: the same expression in all three
    double p[8];    // process variance per sensor 3-block, [7] = point process variance (VIOFilterSettings.h:176-190)
};

// number of squarings for a given infinity norm
__device__ __forceinline__ int exp_squarings(double norm) {
    if (!(norm > exp_theta())) return 0;  // also NaN: the polynomial then carries it into Sigma, where the caller's checks see it
    int s = (int)ceil(log2(norm / exp_theta()));
    return s < 0 ? 0 : (s > EXP_MAX_SQ ? EXP_MAX_SQ : s);
}

// ------------------------------------------------------------------------------------------------
// Compact dt [A B; 0 0] from the Riccati context of the sample (ctx->Fs = I + dt A_ss, dtBs = dt B_s) and the landmark rows of
// the landmark CTAs of riccati_prep_kernel (rows[i] = D(9) | G(36) | Bl(9): D = I + dt A_ii, G = dt A_is on the columns c_sidx, Bl = B_i[:, 0:3]).
// Block 0: sensor rows and the norm reset; block 1 + b: 128 landmarks.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    fill_kernel(const RiccatiCtx* __restrict__ ctx, const double* __restrict__ dtBs, const double* __restrict__ rows, int N,
                double* __restrict__ Ts, double* __restrict__ Tl, unsigned long long* __restrict__ norm) {
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) *norm = 0ull;
        for (int t = threadIdx.x; t < SSIZE; t += blockDim.x) {
            const int r = t / SW, c = t % SW;
            Ts[t] = c < SENSOR_DIM ? ctx->Fs[r * SENSOR_DIM + c] - (r == c ? 1.0 : 0.0) : dtBs[r * 12 + c - SENSOR_DIM];
        }
        return;
    }
    const int i = (blockIdx.x - 1) * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double* ro = rows + (size_t)i * ROWS_STRIDE;
    double* o = Tl + (size_t)i * LSTRIDE;
    const double dt = ctx->dt;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < LW; ++c) o[r * LW + c] = 0.0;
        for (int k = 0; k < 12; ++k) o[r * LW + c_sidx[k]] = ro[9 + 12 * r + k];
        for (int c = 0; c < 3; ++c) o[r * LW + SENSOR_DIM + c] = ro[3 * r + c] - (r == c ? 1.0 : 0.0);
        for (int c = 0; c < 3; ++c) o[r * LW + SENSOR_DIM + 3 + c] = dt * ro[45 + 3 * r + c];
    }
}

// E = I + T (in place): the one-step variants (Sigma <- (I + dt A) Sigma (I + dt A)^T + ..., VIO_eqf.cpp:62-72).
__global__ void __launch_bounds__(128) add_identity_kernel(double* __restrict__ Es, double* __restrict__ El, int N) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < SENSOR_DIM) Es[t * SW + t] += 1.0;
    if (t < 3 * N) {
        const int i = t / 3, r = t % 3;
        El[(size_t)i * LSTRIDE + r * LW + SENSOR_DIM + r] += 1.0;
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
// In place on the compact blocks of dt [A B; 0 0].  Block 0: sensor rows  [M_s A_ss M_s^-1 | M_s B_s];  block 1 + i: landmark i,
// [M_i A_is M_s^-1 | M_i A_ii M_i^-1 | M_i B_i].
__global__ void __launch_bounds__(128)
    normal_transform_kernel(double* __restrict__ Ts, double* __restrict__ Tl, int N, const double* __restrict__ lm, int cap,
                            const double* __restrict__ Ms, const double* __restrict__ MsInv) {
    __shared__ double sM[SENSOR_DIM * SENSOR_DIM], sMi[SENSOR_DIM * SENSOR_DIM], sX[SSIZE], sY[SENSOR_DIM * SENSOR_DIM];
    const int tid = threadIdx.x;
    for (int t = tid; t < SENSOR_DIM * SENSOR_DIM; t += blockDim.x) {
        sM[t] = Ms[t];
        sMi[t] = MsInv[t];
    }
    if (blockIdx.x == 0) {
        for (int t = tid; t < SSIZE; t += blockDim.x) sX[t] = Ts[t];
        __syncthreads();
        for (int t = tid; t < SENSOR_DIM * SENSOR_DIM; t += blockDim.x) {  // Y = A_ss M_s^-1
            const int r = t / SENSOR_DIM, c = t % SENSOR_DIM;
            double acc = 0.0;
            for (int k = 0; k < SENSOR_DIM; ++k) acc += sX[r * SW + k] * sMi[k * SENSOR_DIM + c];
            sY[t] = acc;
        }
        __syncthreads();
        for (int t = tid; t < SSIZE; t += blockDim.x) {
            const int r = t / SW, c = t % SW;
            double acc = 0.0;
            if (c < SENSOR_DIM) {
                for (int k = 0; k < SENSOR_DIM; ++k) acc += sM[r * SENSOR_DIM + k] * sY[k * SENSOR_DIM + c];
            } else {
                for (int k = 0; k < SENSOR_DIM; ++k) acc += sM[r * SENSOR_DIM + k] * sX[k * SW + c];
            }
            Ts[t] = acc;
        }
        return;
    }
    const int i = blockIdx.x - 1;
    if (i >= N) return;
    __shared__ double sMl[9], sMli[9], sR[LSTRIDE], sZ[3 * SENSOR_DIM];
    double* blk = Tl + (size_t)i * LSTRIDE;
    if (tid == 0) {
        const V3 p0 = V3{lm[F_Q0X * cap + i], lm[F_Q0Y * cap + i], lm[F_Q0Z * cap + i]};
        const M3 M = normal_M_landmark(p0), Mi = inverse(M);
        for (int k = 0; k < 9; ++k) {
            sMl[k] = M.m[k];
            sMli[k] = Mi.m[k];
        }
    }
    for (int t = tid; t < LSTRIDE; t += blockDim.x) sR[t] = blk[t];
    __syncthreads();
    for (int t = tid; t < 3 * SENSOR_DIM; t += blockDim.x) {  // Z = A_is M_s^-1
        const int r = t / SENSOR_DIM, c = t % SENSOR_DIM;
        double acc = 0.0;
        for (int k = 0; k < SENSOR_DIM; ++k) acc += sR[r * LW + k] * sMi[k * SENSOR_DIM + c];
        sZ[t] = acc;
    }
    __syncthreads();
    for (int t = tid; t < LSTRIDE; t += blockDim.x) {
        const int r = t / LW, c = t % LW;
        double acc = 0.0;
        if (c < SENSOR_DIM) {
            for (int k = 0; k < 3; ++k) acc += sMl[3 * r + k] * sZ[k * SENSOR_DIM + c];
        } else if (c < SENSOR_DIM + 3) {  // M_i A_ii M_i^-1
            const int cc = c - SENSOR_DIM;
            for (int k = 0; k < 3; ++k) {
                double inner = 0.0;
                for (int q = 0; q < 3; ++q) inner += sR[k * LW + SENSOR_DIM + q] * sMli[3 * q + cc];
                acc += sMl[3 * r + k] * inner;
            }
        } else {
            for (int k = 0; k < 3; ++k) acc += sMl[3 * r + k] * sR[k * LW + c];
        }
        blk[t] = acc;
    }
}

// ---- useDiscreteStateMatrix: stateMatrixADiscrete (EqFMatrices.cpp:24-41) -------------------------------------------
// A0tD = numericalDifferential(a0Discrete, 0) with central differences of step h = cbrt(eps) (Geometry.cpp:25-36).  The
// sensor coordinates of a0Discrete depend on the sensor coordinates only, landmark i's on the sensor coordinates and its
// own three, so A0tD = [[F_s, 0], [Gamma, blockdiag D_i]] needs 43 sensor evaluations (0, +-h e_j) and 49 per landmark
// instead of 2 dim evaluations of the whole state; every entry outside those blocks is an exact zero in the reference
// too (identical function values cancel).  Written over the A-parts of the compact blocks (the B-parts keep dt B).
constexpr int DA_SENSOR_EVALS = 1 + 2 * SENSOR_DIM;  // 43
__device__ __forceinline__ double num_diff_step() { return cbrt(2.220446049250313e-16); }
__global__ void discrete_a_sensor_kernel(int coord, const double* __restrict__ xi0s, const double* __restrict__ Xs,
                                         const double* __restrict__ imuRow, double* __restrict__ Es, SE3* __restrict__ ccOut) {
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
        Es[r * SW + j] = (e1[1 + 2 * j][r] - e1[2 + 2 * j][r]) / (2 * h);
    }
}
// one CTA per landmark: 43 evaluations under the sensor perturbations + 6 under its own
__global__ void discrete_a_landmark_kernel(const double* __restrict__ lm, int cap, int N, int coord, const double* __restrict__ imuRow,
                                           const SE3* __restrict__ cc, double* __restrict__ El) {
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
    double* blk = El + (size_t)i * LSTRIDE;
    for (int t = threadIdx.x; t < 3 * SENSOR_DIM; t += blockDim.x) {
        const int r = t / SENSOR_DIM, j = t % SENSOR_DIM;
        blk[r * LW + j] = (o[1 + 2 * j][r] - o[2 + 2 * j][r]) / (2 * h);
    }
    if (threadIdx.x < 9) {
        const int r = threadIdx.x / 3, c = threadIdx.x % 3;
        blk[r * LW + SENSOR_DIM + c] = (o[DA_SENSOR_EVALS + 2 * c][r] - o[DA_SENSOR_EVALS + 2 * c + 1][r]) / (2 * h);
    }
}

// ------------------------------------------------------------------------------------------------
// Exponential, step 1: |M|_inf = max over the sensor rows and every landmark's rows (the input rows are zero).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    exp_norm_kernel(const double* __restrict__ Ts, const double* __restrict__ Tl, int N, unsigned long long* __restrict__ norm) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0.0;
    if (t < SENSOR_DIM) {
        for (int c = 0; c < SW; ++c) s += fabs(Ts[t * SW + c]);
    } else if (t - SENSOR_DIM < 3 * N) {
        const double* row = Tl + (size_t)(t - SENSOR_DIM) * LW;  // rows of consecutive landmarks are contiguous
        for (int c = 0; c < LW; ++c) s += fabs(row[c]);
    } else {
        return;
    }
    if (!(s == s)) s = __longlong_as_double(0x7FF0000000000000ll);  // NaN -> +inf so that it wins the max
    atomicMax(norm, (unsigned long long)__double_as_longlong(s));   // non-negative doubles order like their bit patterns
}

// ------------------------------------------------------------------------------------------------
// Exponential, step 2 (one CTA): top 21 rows of exp(S / 2^s) by Horner on the degree-12 Taylor polynomial, then the
// squaring ladder  E_(t+1) = E_t^2  (t = 0..s-1), every rung kept for the landmark kernel.  A polynomial of S with constant
// term I has the rows [P_top; 0 I], so a product only needs the 21 x 33 top block:
//     (S P)_top[r][c] = sum_{j<21} S[r][j] P[j][c] + (c >= 21 ? S[r][c] : 0).
// Es <- E_s = [F_s | Bx_s].
// ------------------------------------------------------------------------------------------------
constexpr int EXPS_THREADS = 704;  // 693 entries
__global__ void __launch_bounds__(EXPS_THREADS)
    exp_sensor_kernel(const double* __restrict__ Ts, const unsigned long long* __restrict__ norm, double* __restrict__ ladder,
                      double* __restrict__ Es) {
    __shared__ double sS[SSIZE], sP[2][SSIZE];
    const int t = threadIdx.x;
    const int r = t / SW, c = t % SW;
    const bool live = t < SSIZE;
    const int s = exp_squarings(__longlong_as_double((long long)*norm));
    const double scale = ldexp(1.0, -s);
    if (live) {
        sS[t] = scale * Ts[t];
        sP[0][t] = r == c ? 1.0 : 0.0;
    }
    __syncthreads();
    int cur = 0;
    for (int k = EXP_DEG; k >= 1; --k) {  // P <- I + (S / k) P
        if (live) {
            double acc = c >= SENSOR_DIM ? sS[t] : 0.0;
#pragma unroll
            for (int j = 0; j < SENSOR_DIM; ++j) acc += sS[r * SW + j] * sP[cur][j * SW + c];
            sP[1 - cur][t] = (r == c ? 1.0 : 0.0) + acc / (double)k;
        }
        __syncthreads();
        cur = 1 - cur;
    }
    if (live) ladder[t] = sP[cur][t];
    for (int q = 0; q < s; ++q) {  // E <- E E
        if (live) {
            double acc = c >= SENSOR_DIM ? sP[cur][t] : 0.0;
#pragma unroll
            for (int j = 0; j < SENSOR_DIM; ++j) acc += sP[cur][r * SW + j] * sP[cur][j * SW + c];
            sP[1 - cur][t] = acc;
            ladder[(size_t)(q + 1) * SSIZE + t] = acc;
        }
        __syncthreads();
        cur = 1 - cur;
    }
    if (live) Es[t] = sP[cur][t];
}

// ------------------------------------------------------------------------------------------------
// Exponential, step 3 (one CTA per landmark, 128 threads: 99 own an entry of the 3 x 33 row block, 9 an entry of the 3 x 3 one).
// Terms of the Taylor series carried with their 1 / k!:
//     R_1 = g, d_1 = d;   R_(k+1) = (R_k S + d_k g) / (k + 1),  d_(k+1) = d_k d / (k + 1);   X = sum R_k,  F = I + sum d_k,
// then s squarings against the ladder of the sensor kernel.  El[i] <- [X[:, 0:21] | F | X[:, 21:33]].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    exp_landmark_kernel(const double* __restrict__ Ts, const double* __restrict__ Tl, const unsigned long long* __restrict__ norm,
                        const double* __restrict__ ladder, double* __restrict__ El, int N) {
    __shared__ double sS[SSIZE];          // S / 2^s during the series, the current rung during the squarings
    __shared__ double sG[3 * SW];         // g / 2^s
    __shared__ double sD[9];              // d / 2^s
    __shared__ double sR[2][3 * SW];      // R_k, later X
    __shared__ double sK[2][9];           // d_k, later F
    const int i = blockIdx.x;
    if (i >= N) return;
    const int t = threadIdx.x;
    const int s = exp_squarings(__longlong_as_double((long long)*norm));
    const double scale = ldexp(1.0, -s);
    const double* blk = Tl + (size_t)i * LSTRIDE;
    for (int k = t; k < SSIZE; k += blockDim.x) sS[k] = scale * Ts[k];
    const bool isX = t < 3 * SW, isF = t >= 3 * SW && t < 3 * SW + 9;
    const int r = isX ? t / SW : (t - 3 * SW) / 3, c = isX ? t % SW : (t - 3 * SW) % 3;
    if (isX) {
        const double v = scale * blk[r * LW + (c < SENSOR_DIM ? c : c + 3)];
        sG[t] = v;
        sR[0][t] = v;
    } else if (isF) {
        const double v = scale * blk[r * LW + SENSOR_DIM + c];
        sD[t - 3 * SW] = v;
        sK[0][t - 3 * SW] = v;
    }
    __syncthreads();
    double acc = isX ? sR[0][t] : (isF ? (r == c ? 1.0 : 0.0) + sK[0][t - 3 * SW] : 0.0);  // X or F
    int cur = 0;
    for (int k = 1; k < EXP_DEG; ++k) {
        const double inv = 1.0 / (double)(k + 1);
        if (isX) {
            double v = sK[cur][3 * r] * sG[c] + sK[cur][3 * r + 1] * sG[SW + c] + sK[cur][3 * r + 2] * sG[2 * SW + c];
#pragma unroll
            for (int j = 0; j < SENSOR_DIM; ++j) v += sR[cur][r * SW + j] * sS[j * SW + c];
            v *= inv;
            sR[1 - cur][t] = v;
            acc += v;
        } else if (isF) {
            const double v = (sK[cur][3 * r] * sD[c] + sK[cur][3 * r + 1] * sD[3 + c] + sK[cur][3 * r + 2] * sD[6 + c]) * inv;
            sK[1 - cur][t - 3 * SW] = v;
            acc += v;
        }
        __syncthreads();
        cur = 1 - cur;
    }
    // squarings: X <- X E + F X,  F <- F F
    if (isX) sR[cur][t] = acc;
    if (isF) sK[cur][t - 3 * SW] = acc;
    for (int q = 0; q < s; ++q) {
        __syncthreads();  // X, F of this level published; everybody is done with the previous rung
        for (int k = t; k < SSIZE; k += blockDim.x) sS[k] = ladder[(size_t)q * SSIZE + k];
        __syncthreads();
        if (isX) {
            double v = c >= SENSOR_DIM ? sR[cur][t] : 0.0;
            v += sK[cur][3 * r] * sR[cur][c] + sK[cur][3 * r + 1] * sR[cur][SW + c] + sK[cur][3 * r + 2] * sR[cur][2 * SW + c];
#pragma unroll
            for (int j = 0; j < SENSOR_DIM; ++j) v += sR[cur][r * SW + j] * sS[j * SW + c];
            sR[1 - cur][t] = v;
            acc = v;
        } else if (isF) {
            const double v = sK[cur][3 * r] * sK[cur][c] + sK[cur][3 * r + 1] * sK[cur][3 + c] + sK[cur][3 * r + 2] * sK[cur][6 + c];
            sK[1 - cur][t - 3 * SW] = v;
            acc = v;
        }
        cur = 1 - cur;
    }
    double* o = El + (size_t)i * LSTRIDE;
    if (isX) o[r * LW + (c < SENSOR_DIM ? c : c + 3)] = acc;
    if (isF) o[r * LW + SENSOR_DIM + c] = acc;
}

// ------------------------------------------------------------------------------------------------
// Propagation with compact blocks, sensor-sensor part (one CTA):  Sigma'_ss = F_s Sigma_ss F_s^T + Bx_s Qn Bx_s^T + dt P_s,
// and the zero pad rows / columns 21..23 of the sensor block.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(448)
    gprop_sensor_kernel(const double* __restrict__ Es, NoiseArgs nz, const RiccatiCtx* __restrict__ ctx, const double* __restrict__ Sin,
                        double* __restrict__ Sout, int ld) {
    __shared__ double sE[SSIZE], sS[441], sT[441];
    const int t = threadIdx.x;
    const double dt = ctx->dt;
    for (int k = t; k < SSIZE; k += blockDim.x) sE[k] = Es[k];
    if (t < 441) sS[t] = Sin[(size_t)(t % 21) * ld + t / 21];
    __syncthreads();
    const int r = t / 21, c = t % 21;
    if (t < 441) {
        double s = 0;
        for (int k = 0; k < 21; ++k) s += sE[r * SW + k] * sS[k * 21 + c];
        sT[t] = s;
    }
    __syncthreads();
    if (t < 441) {
        double s = r == c ? dt * nz.p[r / 3] : 0.0;
        for (int k = 0; k < 12; ++k) s += sE[r * SW + 21 + k] * (nz.q[k / 3] / dt) * sE[c * SW + 21 + k];
        for (int k = 0; k < 21; ++k) s += sT[r * 21 + k] * sE[c * SW + k];
        Sout[(size_t)c * ld + r] = s;
    }
    if (t < 3 * SOFF) {
        const int p = SENSOR_DIM + t / SOFF, q = t % SOFF;
        Sout[(size_t)q * ld + p] = 0.0;
        Sout[(size_t)p * ld + q] = 0.0;
    }
}

// ------------------------------------------------------------------------------------------------
// Per landmark (32 threads each, 4 per CTA):  H_i, the factors U_i | V_i (3 x 54 each, row-major) and the sensor-landmark block
// Sigma'_is = H_i F_s^T + Bx_i Qn Bx_s^T in both triangles; pad rows / columns 21..23 stay zero.
// ------------------------------------------------------------------------------------------------
constexpr int GS_LM = 4;
__global__ void __launch_bounds__(GS_LM * 32)
    gprop_strip_kernel(const double* __restrict__ Es, const double* __restrict__ El, NoiseArgs nz, const RiccatiCtx* __restrict__ ctx,
                       const double* __restrict__ Sin, double* __restrict__ Sout, int ld, int N, double* __restrict__ uv) {
    const double dt = ctx->dt;
    __shared__ double sE[SSIZE], sS[441];
    __shared__ double sB[GS_LM][LSTRIDE];   // Gamma | D | Bx of the landmark
    __shared__ double sL[GS_LM][63];        // sL[a * 21 + c] = Sigma[row a of the landmark, sensor column c]
    __shared__ double sH[GS_LM][63];
    const int tid = threadIdx.x, li = tid / 32, t = tid % 32;
    const int i = blockIdx.x * GS_LM + li;
    const bool live = i < N;
    const int r0 = SOFF + 3 * i;
    for (int k = tid; k < SSIZE; k += blockDim.x) sE[k] = Es[k];
    for (int k = tid; k < 441; k += blockDim.x) sS[k] = Sin[(size_t)(k % 21) * ld + k / 21];
    if (live) {
        for (int k = t; k < LSTRIDE; k += 32) sB[li][k] = El[(size_t)i * LSTRIDE + k];
        if (t < 21) {
#pragma unroll
            for (int a = 0; a < 3; ++a) sL[li][a * 21 + t] = Sin[(size_t)t * ld + r0 + a];
        }
    }
    __syncthreads();
    const double* B = sB[li];
    if (live) {
        double* U = uv + (size_t)i * GUV_STRIDE;
        double* V = U + 3 * GRANK;
        if (t < 21) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double* D = B + a * LW + SENSOR_DIM;
                const double e = D[0] * sL[li][t] + D[1] * sL[li][21 + t] + D[2] * sL[li][42 + t];
                double h = e;
                for (int k = 0; k < 21; ++k) h += B[a * LW + k] * sS[k * 21 + t];
                sH[li][a * 21 + t] = h;
                U[GRANK * a + t] = B[a * LW + t];
                U[GRANK * a + 21 + t] = h;
                V[GRANK * a + t] = e;
                V[GRANK * a + 21 + t] = B[a * LW + t];
            }
        }
        if (t < 12) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double b = B[a * LW + SENSOR_DIM + 3 + t];
                U[GRANK * a + 42 + t] = b * (nz.q[t / 3] / dt);
                V[GRANK * a + 42 + t] = b;
            }
        }
    }
    __syncthreads();
    if (!live) return;
    if (t < 21) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double v = 0.0;
            for (int k = 0; k < 12; ++k) v += B[a * LW + SENSOR_DIM + 3 + k] * (nz.q[k / 3] / dt) * sE[t * SW + 21 + k];
            for (int k = 0; k < 21; ++k) v += sH[li][a * 21 + k] * sE[t * SW + k];
            Sout[(size_t)(r0 + a) * ld + t] = v;
            Sout[(size_t)t * ld + r0 + a] = v;
        }
    } else if (t < SOFF) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            Sout[(size_t)(r0 + a) * ld + t] = 0.0;
            Sout[(size_t)t * ld + r0 + a] = 0.0;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Landmark-landmark block, HBM-bound (one read of Sigma_in, one write of Sigma_out), same tiling as prop_ll_kernel:
//   Sigma'_ij = D_i Sigma_ij D_j^T + U_i V_j^T (+ dt p_l I),  lower tiles computed, mirror tiles written transposed.
// ------------------------------------------------------------------------------------------------
constexpr int GTP = 16;
constexpr int GLL_LD = 3 * GRANK + 1;  // 163: odd stride, the 16 landmarks of a tile row hit different banks
__global__ void __launch_bounds__(GTP* GTP)
    gprop_ll_kernel(const double* __restrict__ El, const double* __restrict__ uv, const RiccatiCtx* __restrict__ ctx, const double* __restrict__ Sin,
                    double* __restrict__ Sout, int ld, int N) {
    const int ti = blockIdx.y, tj = blockIdx.x;
    if (tj > ti) return;
    const double plDiag = ctx->plDiag;  // dt * point process variance
    extern __shared__ double gsm[];
    double* sU = gsm;                      // [GTP][GLL_LD]
    double* sV = sU + GTP * GLL_LD;        // [GTP][GLL_LD]
    double* sDi = sV + GTP * GLL_LD;       // [GTP][9]
    double* sDj = sDi + GTP * 9;
    const int tid = threadIdx.y * GTP + threadIdx.x;
    const int i0 = ti * GTP, j0 = tj * GTP;
    for (int t = tid; t < GTP * 3 * GRANK; t += GTP * GTP) {
        const int l = t / (3 * GRANK), k = t % (3 * GRANK);
        sU[l * GLL_LD + k] = (i0 + l < N) ? uv[(size_t)(i0 + l) * GUV_STRIDE + k] : 0.0;
        sV[l * GLL_LD + k] = (j0 + l < N) ? uv[(size_t)(j0 + l) * GUV_STRIDE + 3 * GRANK + k] : 0.0;
    }
    for (int t = tid; t < GTP * 9; t += GTP * GTP) {
        const int l = t / 9, k = t % 9;
        sDi[t] = (i0 + l < N) ? El[(size_t)(i0 + l) * LSTRIDE + (k / 3) * LW + SENSOR_DIM + k % 3] : 0.0;
        sDj[t] = (j0 + l < N) ? El[(size_t)(j0 + l) * LSTRIDE + (k / 3) * LW + SENSOR_DIM + k % 3] : 0.0;
    }
    __syncthreads();
    const int li = threadIdx.x, lj = threadIdx.y;  // threadIdx.x walks rows (contiguous in memory)
    const int i = i0 + li, j = j0 + lj;
    if (i >= N || j >= N) return;
    const int r0 = SOFF + 3 * i, c0 = SOFF + 3 * j;
    double S[9];  // S[a*3+b] = Sigma[r0+a, c0+b]
    for (int b = 0; b < 3; ++b)
        for (int a = 0; a < 3; ++a) S[a * 3 + b] = Sin[(size_t)(c0 + b) * ld + r0 + a];
    const double* Di = sDi + 9 * li;
    const double* Dj = sDj + 9 * lj;
    double T[9];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) T[a * 3 + b] = Di[3 * a] * S[b] + Di[3 * a + 1] * S[3 + b] + Di[3 * a + 2] * S[6 + b];
    double O[9];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            double s = T[3 * a] * Dj[3 * b] + T[3 * a + 1] * Dj[3 * b + 1] + T[3 * a + 2] * Dj[3 * b + 2];
            const double* u = sU + li * GLL_LD + GRANK * a;
            const double* v = sV + lj * GLL_LD + GRANK * b;
#pragma unroll 6
            for (int k = 0; k < GRANK; ++k) s += u[k] * v[k];
            O[a * 3 + b] = s;
        }
    if (i == j) {
        O[0] += plDiag;
        O[4] += plDiag;
        O[8] += plDiag;
    }
    for (int b = 0; b < 3; ++b)
        for (int a = 0; a < 3; ++a) Sout[(size_t)(c0 + b) * ld + r0 + a] = O[a * 3 + b];
    if (ti != tj) {
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) Sout[(size_t)(r0 + a) * ld + c0 + b] = O[a * 3 + b];
    }
}
constexpr int GLL_SMEM = (2 * GTP * GLL_LD + 2 * GTP * 9) * (int)sizeof(double);  // 44 KB

}  // namespace sric
}  // namespace eqvio
