// chunk_factor_mma.cuh -- included from kernels.cuh (needs its ChunkStage / TMA / mbarrier helpers and dmma884).
// ------------------------------------------------------------------------------------------------
// chunk_factor_mma_kernel: the chunk factor step of the sequential-chunk correction (performVisionUpdate,
// VIO_eqf.cpp:105-135; derivation above chunk_factor_kernel): S_c = C_c Sigma[L_c, L_c] C_c^T + sigma^2 I, its elimination,
// Y_c = L_c^-1 C_c Sigma for this CTA's state columns, z_c, Gamma += Y_c^T z_c.
//
// The augmented matrix [S_c; W_c^T (COLS state columns); r^T] (64 + COLS + 1 rows, 64 columns) lives in the ACCUMULATOR
// FRAGMENTS of mma.sync.m8n8k4.f64: 8 x 8 tiles, lane l of the owning warp holds entries (l / 4, 2 (l % 4) + {0, 1}).  The
// elimination is right-looking over block columns of four pivots, unscaled columns (after block column J its entries hold
// v_ij = L_ij L_jj), two CTA barriers per block column:
//   1. diagonal: the warp holding the diagonal tile hands its 4 x 4 block to shared memory and eliminates it fraction-free
//      (products only, the four pivot reciprocals side by side) -- the one serial fp64 latency chain of the step; it runs
//      as a look-ahead, right after that warp's trailing update of the previous block column and beside everybody else's;
//   2. panel: the two lanes that hold a row's four entries of the block column finish them (three shuffles per row) and
//      publish the block column as two 4-wide strips: Pa = -(P diag(1 / v_jj)) for all rows, Pb = P for the S rows, rows of
//      the block column itself zeroed;
//   3. trailing update: every tile to the right is ONE DMMA, C += Pa[rows] Pb[cols]^T, with both operand fragments read as
//      32 consecutive doubles of a strip (lane l <-> row l / 4, k = l % 4: conflict-free 8-byte loads).
// The round-1 kernel (4 x 4 register tiles per thread) spends 830 of its ~1560 cycles per block column in the trailing
// update -- 96 fp64 instructions and 32 shared-memory loads per thread and step; here it is 2 loads + 1 DMMA per tile and a
// warp holds 6-8 tiles.  COLS = state columns per CTA (16: 52 CTAs at N = 256; 32 when 16 would exceed one wave of CTAs).
// ------------------------------------------------------------------------------------------------
#pragma once

constexpr int MM_WARPS = 20;
constexpr int MM_THREADS = MM_WARPS * 32;

template <int COLS>
struct MmSmem {
    static constexpr int RHSB = (COLS + 8) / 8;   // 8-row blocks of right-hand sides: COLS state columns, the residual row, padding
    static constexpr int ROWS = CH_R + 8 * RHSB;
    static constexpr int NTILES = 36 + 8 * RHSB;  // lower 8x8 tiles of S_c + the right-hand-side tiles
    static constexpr int TPW = (NTILES + MM_WARPS - 1) / MM_WARPS;  // tiles per warp
    static constexpr int YT_LD = COLS + 2;
    alignas(16) double Pa[ROWS][CH_T];   // -(v_ik / v_kk): row-side operand of the trailing update
    alignas(16) double Pb[CH_R][CH_T];   // v_jk: column-side operand (S rows only)
    alignas(16) double Dt[CH_T][CH_T];   // the diagonal 4x4 block on its way to the eliminating warp
    alignas(16) double Dd[CH_T][CH_T];   // its unscaled columns
    alignas(16) double Dc[CH_T];         // 1 / v_jj
    double Piv[CH_R];                    // v_kk, then 1 / L_kk
    double C[CH_R / 2][6];
    double Yt[CH_R][YT_LD];              // scaled rows of Y for this CTA's columns, the residual z in column COLS
    double Gp[8][COLS];
    int Idx[CH_R / 2];
};
template <int COLS>
__host__ __device__ constexpr int mm_stage_off() { return ((int)sizeof(MmSmem<COLS>) + 127) & ~127; }
template <int COLS>
__host__ __device__ constexpr int mm_smem_bytes() { return mm_stage_off<COLS>() + (int)sizeof(ChunkStage); }

// tile index (column-major: for every column block nK the S tiles mI = nK..7, then the right-hand-side blocks 8..) -> (mI, nK)
template <int RHSB>
__device__ __forceinline__ void mm_decode(int idx, int& mI, int& nK) {
    int n = 0;
    for (;; ++n) {
        const int cnt = (8 - n) + RHSB;
        if (idx < cnt) break;
        idx -= cnt;
    }
    nK = n;
    mI = idx < 8 - n ? n + idx : 8 + (idx - (8 - n));
}

template <int COLS>
__global__ void __launch_bounds__(MM_THREADS)
    chunk_factor_mma_kernel(const double* __restrict__ Sig, int ld, int dimp, const int* __restrict__ lmOf,
                            const double* __restrict__ Cblk, const double* __restrict__ ytilde, int j0, int bc, double r2,
                            const double* __restrict__ GammaIn, double* __restrict__ GammaOut, double* __restrict__ Y,
                            int* __restrict__ status, const int* __restrict__ guard, int tl, int stage,
                            const __grid_constant__ CUtensorMap sigMap) {
    using Smem = MmSmem<COLS>;
    constexpr int RHSB = Smem::RHSB, TPW = Smem::TPW;
    extern __shared__ __align__(128) unsigned char chunk_smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(chunk_smem_raw);
    ChunkStage& stg = *reinterpret_cast<ChunkStage*>(chunk_smem_raw + mm_stage_off<COLS>());
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lr = lane >> 2, q = lane & 3;  // fragment coordinates: row lr, columns 2q, 2q + 1 of an 8x8 tile
    const int rc = 2 * bc;
    CH_STAMP(0);
    // Cblk / lmOf come from meas_kernel and the frame upload, several launches back: staged ahead of the dependency wait
    for (int t = tid; t < bc * 6; t += MM_THREADS) sm.C[t / 6][t % 6] = Cblk[6 * (size_t)j0 + t];
    for (int t = tid; t < bc; t += MM_THREADS) sm.Idx[t] = SOFF + 3 * lmOf[j0 + t];
    for (int t = tid; t < CH_R; t += MM_THREADS) sm.Piv[t] = 1.0;
    if (stage && tid == 0) {
        mbar_init(&stg.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();
    if (*guard) return;
    TL_MARK(tl, 0);
    const int lm0 = lmOf[j0];
    const int consecutive = __syncthreads_and(tid >= bc || lmOf[j0 + tid] == lm0 + tid);
    CH_STAMP(1);
    const bool staged = stage != 0 && consecutive && (lm0 & 1) == 0;  // the TMA box must start 16-byte aligned
    if (staged && tid == 0) {
        mbar_expect_tx(&stg.bar, (uint32_t)sizeof(stg.S));
        tma_load_2d(&stg.S[0][0], &sigMap, SOFF + 3 * lm0, SOFF + 3 * lm0, &stg.bar);
    }
    const int sbase = blockIdx.x * COLS;
    const int nJ = (rc + CH_T - 1) / CH_T;

    // ---- this warp's tiles ----
    int mI[TPW], nK[TPW];
    bool valid[TPW];
    double acc[TPW][2];
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
        const int idx = warp + t * MM_WARPS;
        valid[t] = idx < Smem::NTILES;
        mI[t] = nK[t] = 0;
        if (valid[t]) mm_decode<RHSB>(idx, mI[t], nK[t]);
        acc[t][0] = acc[t][1] = 0.0;
    }
    // right-hand-side entries first (global loads in flight while the S block arrives)
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
        if (!valid[t] || mI[t] < 8) continue;
        const int rr = 8 * (mI[t] - 8) + lr;  // row inside the right-hand-side part
        const int j = (8 * nK[t] + 2 * q) >> 1;  // landmark of the two columns
        if (j >= bc) continue;
        if (rr < COLS) {
            const int s = sbase + rr;  // state column of W_c
            if (s < dimp) {
                const double* sp = Sig + (size_t)sm.Idx[j] * ld + s;  // Sigma[s, cols of j] (symmetric storage)
                const double w0 = sp[0], w1 = sp[ld], w2 = sp[2 * (size_t)ld];
                acc[t][0] = sm.C[j][0] * w0 + sm.C[j][1] * w1 + sm.C[j][2] * w2;
                acc[t][1] = sm.C[j][3] * w0 + sm.C[j][4] * w1 + sm.C[j][5] * w2;
            }
        } else if (rr == COLS) {  // residual row: ytilde_c - C_c Gamma
            const int g = sm.Idx[j];
            const double g0 = GammaIn[g], g1 = GammaIn[g + 1], g2 = GammaIn[g + 2];
            acc[t][0] = ytilde[2 * (j0 + j)] - (sm.C[j][0] * g0 + sm.C[j][1] * g1 + sm.C[j][2] * g2);
            acc[t][1] = ytilde[2 * (j0 + j) + 1] - (sm.C[j][3] * g0 + sm.C[j][4] * g1 + sm.C[j][5] * g2);
        }
    }
    if (staged) mbar_wait(&stg.bar, 0);
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
        if (!valid[t] || mI[t] >= 8) continue;
        const int row = 8 * mI[t] + lr, col = 8 * nK[t] + 2 * q;
        const int i = row >> 1, e = row & 1, j = col >> 1;  // row landmark / component, column landmark
        if (i < bc && j < bc) {
            double P[9];  // P[aa * 3 + b] = Sigma[rows of i (aa), cols of j (b)]
            if (staged) {
#pragma unroll
                for (int aa = 0; aa < 3; ++aa)
#pragma unroll
                    for (int b = 0; b < 3; ++b) P[aa * 3 + b] = stg.S[3 * j + b][3 * i + aa];
            } else {
                const double* sp = Sig + (size_t)sm.Idx[j] * ld + sm.Idx[i];
#pragma unroll
                for (int b = 0; b < 3; ++b)
#pragma unroll
                    for (int aa = 0; aa < 3; ++aa) P[aa * 3 + b] = sp[(size_t)b * ld + aa];
            }
            double T[3];
#pragma unroll
            for (int b = 0; b < 3; ++b) T[b] = sm.C[i][3 * e] * P[b] + sm.C[i][3 * e + 1] * P[3 + b] + sm.C[i][3 * e + 2] * P[6 + b];
            acc[t][0] = T[0] * sm.C[j][0] + T[1] * sm.C[j][1] + T[2] * sm.C[j][2];
            acc[t][1] = T[0] * sm.C[j][3] + T[1] * sm.C[j][4] + T[2] * sm.C[j][5];
        }
        if (row == col) acc[t][0] = row < rc ? acc[t][0] + r2 : 1.0;          // identity padding of a short last chunk
        if (row == col + 1) acc[t][1] = row < rc ? acc[t][1] + r2 : 1.0;
    }
    CH_STAMP(2);

    // diagonal block of block column J: to shared memory, fraction-free elimination by the warp that holds it -- the one serial
    // fp64 latency chain of a step.  Called for J + 1 by that warp right after its own DMMAs of step J (look-ahead): the other
    // warps are still in their trailing updates, so only the panel phase waits for it.
    auto diag_phase = [&](int J) {
        const int nKJ = J >> 1, half = J & 1;
#pragma unroll
        for (int t = 0; t < TPW; ++t) {
            if (!(valid[t] && mI[t] == nKJ && nK[t] == nKJ)) continue;  // warp-uniform
            if ((lr >> 2) == half && (q >> 1) == half) {
                const int i = lr & 3, k0 = 2 * (q & 1);
                *reinterpret_cast<double2*>(&sm.Dt[i][k0]) = make_double2(acc[t][0], acc[t][1]);
            }
            __syncwarp();
            const double a00 = sm.Dt[0][0], a10 = sm.Dt[1][0], a20 = sm.Dt[2][0], a30 = sm.Dt[3][0];
            const double b11 = sm.Dt[1][1], b21 = sm.Dt[2][1], b22 = sm.Dt[2][2], b31 = sm.Dt[3][1], b32 = sm.Dt[3][2], b33 = sm.Dt[3][3];
            const double m11 = b11 * a00 - a10 * a10, m21 = b21 * a00 - a20 * a10, m22 = b22 * a00 - a20 * a20;
            const double m31 = b31 * a00 - a30 * a10, m32 = b32 * a00 - a30 * a20, m33 = b33 * a00 - a30 * a30;
            const double n22 = m22 * m11 - m21 * m21, n32 = m32 * m11 - m31 * m21, n33 = m33 * m11 - m31 * m31;
            const double p33 = n33 * n22 - n32 * n32;
            const double r0 = fast_rcp(a00), r1 = fast_rcp(m11), r2_ = fast_rcp(n22), r3 = fast_rcp(p33);
            const double s2 = r0 * r1, s3 = s2 * r2_, e1 = a00 * m11;
            if (lane == 0) {
                // unscaled columns v_ij = L_ij L_jj (the Schur-complement values) and 1 / v_jj
                sm.Dd[1][0] = a10;
                sm.Dd[2][0] = a20;
                sm.Dd[3][0] = a30;
                sm.Dd[2][1] = m21 * r0;
                sm.Dd[3][1] = m31 * r0;
                sm.Dd[3][2] = n32 * s2;
                sm.Dc[0] = r0;
                sm.Dc[1] = a00 * r1;
                sm.Dc[2] = e1 * r2_;
                sm.Dc[3] = (e1 * n22) * r3;
                sm.Piv[CH_T * J + 0] = a00;
                sm.Piv[CH_T * J + 1] = m11 * r0;
                sm.Piv[CH_T * J + 2] = n22 * s2;
                sm.Piv[CH_T * J + 3] = p33 * s3;
            }
        }
    };
    diag_phase(0);
    __syncthreads();
    for (int J = 0; J < nJ; ++J) {
        const int nKJ = J >> 1, half = J & 1;
        CH_FINE(4 * J);
        CH_FINE(4 * J + 1);
        CH_FINE(4 * J + 2);
        // ---- 2. panel: finish the four columns of block column J, publish the strips ----
        {
            const double c0 = sm.Dc[0], c1 = sm.Dc[1], c2 = sm.Dc[2], c3 = sm.Dc[3];
            const double d10 = sm.Dd[1][0], d20 = sm.Dd[2][0], d30 = sm.Dd[3][0], d21 = sm.Dd[2][1], d31 = sm.Dd[3][1], d32 = sm.Dd[3][2];
            const bool first = (q & 1) == 0;  // holds columns 4J, 4J+1 of its row; the next lane holds 4J+2, 4J+3
#pragma unroll
            for (int t = 0; t < TPW; ++t) {
                if (!(valid[t] && nK[t] == nKJ)) continue;  // warp-uniform
                const int row = 8 * mI[t] + lr;
                double x0 = acc[t][0], x1 = acc[t][1];
                // first lane: a0, a1 -> t0 = a0 c0, a1 -= t0 d10, t1 = a1 c1
                const double t0 = x0 * c0;
                const double y1 = x1 - t0 * d10;
                const double t1 = y1 * c1;
                const double s0 = __shfl_up_sync(0xffffffffu, t0, 1), s1 = __shfl_up_sync(0xffffffffu, t1, 1);
                // second lane: a2 -= t0 d20 + t1 d21, t2 = a2 c2, a3 -= t0 d30 + t1 d31 + t2 d32, t3 = a3 c3
                const double y2 = (x0 - s0 * d20) - s1 * d21;
                const double t2 = y2 * c2;
                const double y3 = ((x1 - s0 * d30) - s1 * d31) - t2 * d32;
                const double t3 = y3 * c3;
                if ((q >> 1) == half) {
                    const int k0 = 2 * (q & 1);
                    if (row >= CH_T * J + CH_T) {
                        if (first) {
                            acc[t][1] = y1;
                            *reinterpret_cast<double2*>(&sm.Pa[row][k0]) = make_double2(-t0, -t1);
                            if (mI[t] < 8) *reinterpret_cast<double2*>(&sm.Pb[row][k0]) = make_double2(x0, y1);
                        } else {
                            acc[t][0] = y2;
                            acc[t][1] = y3;
                            *reinterpret_cast<double2*>(&sm.Pa[row][k0]) = make_double2(-t2, -t3);
                            if (mI[t] < 8) *reinterpret_cast<double2*>(&sm.Pb[row][k0]) = make_double2(y2, y3);
                        }
                    } else if (row >= CH_T * J) {  // rows of the block column itself: no further updates
                        *reinterpret_cast<double2*>(&sm.Pa[row][k0]) = make_double2(0.0, 0.0);
                        *reinterpret_cast<double2*>(&sm.Pb[row][k0]) = make_double2(0.0, 0.0);
                    }
                }
            }
        }
        __syncthreads();
        CH_FINE(4 * J + 3);
        // ---- 3. trailing update: one DMMA per tile to the right ----
        {
            const double* pa = &sm.Pa[0][0];
            const double* pb = &sm.Pb[0][0];
            double af[TPW], bf[TPW];  // all operand fragments first (no branches between the loads), then the DMMAs
#pragma unroll
            for (int t = 0; t < TPW; ++t) {
                af[t] = pa[32 * mI[t] + lane];
                bf[t] = pb[32 * nK[t] + lane];
            }
#pragma unroll
            for (int t = 0; t < TPW; ++t)
                if (valid[t] && (nK[t] > nKJ || (nK[t] == nKJ && half == 0))) dmma884(acc[t][0], acc[t][1], af[t], bf[t]);  // warp-uniform
        }
        // look-ahead: the warp that holds the next diagonal block eliminates it now (Dd / Dc were last read in this step's panel
        // phase, before the barrier above); the barrier below publishes it and orders the strip reads before the next panel phase
        if (J + 1 < nJ) diag_phase(J + 1);
        __syncthreads();
    }
    CH_STAMP(3);
    __syncthreads();
    // 1 / L_kk
    if (tid < CH_R) {
        const double piv = sm.Piv[tid];
        if (!(piv > 0.0)) {
            if (blockIdx.x == 0) atomicOr(status, 1);
            sm.Piv[tid] = 1.0;
        } else {
            sm.Piv[tid] = 1.0 / sqrt(piv);
        }
    }
    __syncthreads();
    CH_STAMP(4);
    // Y[k][s] = v_sk / L_kk, staged so that the global store and the Gamma dot products run in a fixed order
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
        if (!valid[t] || mI[t] < 8) continue;
        const int rr = 8 * (mI[t] - 8) + lr, col = 8 * nK[t] + 2 * q;
        if (rr <= COLS) {
            sm.Yt[col][rr] = acc[t][0] * sm.Piv[col];
            sm.Yt[col + 1][rr] = acc[t][1] * sm.Piv[col + 1];
        }
    }
    __syncthreads();
    CH_STAMP(5);
    // all CH_R rows are written (zero beyond rc and for the pad columns s >= dimp): the downdate reads whole tiles
    for (int t = tid; t < CH_R * COLS; t += MM_THREADS) {
        const int k = t / COLS, sl = t % COLS;
        Y[yb_index(k, sbase + sl)] = sm.Yt[k][sl];
    }
    // Gamma += Y_c^T z_c: eight partial sums per state column (fixed order), combined by the first COLS threads
    {
        const int col = tid % COLS, part = tid / COLS;
        if (part < 8) {
            double g = 0.0;
#pragma unroll
            for (int k = 0; k < CH_R / 8; ++k) g += sm.Yt[8 * part + k][col] * sm.Yt[8 * part + k][COLS];
            sm.Gp[part][col] = g;
        }
        __syncthreads();
        if (tid < COLS && sbase + tid < dimp) {
            const double t = GammaIn[sbase + tid];
            const double acc8 = ((sm.Gp[0][tid] + sm.Gp[1][tid]) + (sm.Gp[2][tid] + sm.Gp[3][tid])) +
                                ((sm.Gp[4][tid] + sm.Gp[5][tid]) + (sm.Gp[6][tid] + sm.Gp[7][tid]));
            GammaOut[sbase + tid] = t + acc8;  // ping-pong: other CTAs may still be reading GammaIn
        }
    }
    CH_STAMP(6);
    TL_MARK(tl, 1);
}
