// chunk_factor_df.cuh -- included from kernels.cuh (needs its ChunkStage / TMA / flag helpers).
// ------------------------------------------------------------------------------------------------
// chunk_factor_df_kernel: the chunk factor step of the sequential-chunk correction (performVisionUpdate,
// VIO_eqf.cpp:105-135; see the derivation above chunk_factor_kernel) as a DATAFLOW elimination.
//
// Same arithmetic as chunk_factor_kernel -- the augmented matrix [S_c; W_c^T (this CTA's state columns); r^T] in 4x4
// register tiles, right-looking, unscaled columns, fraction-free diagonal tiles -- but no CTA-wide barrier inside the
// elimination.  The 64 pivots of S_c form a chain of 16 diagonal tiles; what bounds the old kernel (~1150 cycles per
// block column) is the diag -> barrier -> panel -> barrier -> update hand-off through five warps.  Here:
//   * chain warp (warp 0): walks the diagonal.  At step J it takes tile (J, J-1) and tile (J, J) from their owners
//     (updated through block column J-2), finishes the panel tile (J, J-1) against diagonal J-1, applies its rank-4
//     update to (J, J), eliminates inside (J, J) and publishes.  One thread's dependent fp64 chain, no hand-off to
//     another thread anywhere on it.  Warps 4 / 8 / 12 stay idle: they would share the chain warp's scheduler and its
//     FP64 issue slots (a warp-wide DFMA occupies the sub-partition's pipe for two cycles).
//   * S warps (1,2,3,5,6): one lower tile of S_c per lane, column-major.  Per block column j: panel owners wait for
//     diagonal j and publish their tile; everyone to the right waits for the two panel tiles it needs (per-tile flags,
//     release / acquire in shared memory) and applies the rank-4 update; the two tiles next to the diagonal hand
//     themselves to the chain warp once they are current.
//   * RHS warps (7,9,10,11,13): one tile row (4 state columns, or the residual) per half-warp as before, following the
//     flags instead of a progress counter.
// Every published tile is written once and never overwritten, so the flags are monotonic and there is no WAR hazard.
// All waits are warp-uniform (__all_sync) and bounded: a wait that exceeds CF_SPIN_LIMIT raises status bit 8 and
// the CTA runs to its end (no hang).  Dependencies inside one warp always point backwards in program order
// (phase A of step j publishes what phase B of step j consumes), so the uniform waits cannot deadlock.
// COLS = state columns per CTA (16: more, lighter CTAs -- 52 at N = 256; 32 when 16 would exceed one wave).
// ------------------------------------------------------------------------------------------------
#pragma once

constexpr int CF_WARPS = 14;
constexpr int CF_THREADS = CF_WARPS * 32;
constexpr int CF_SPIN_LIMIT = 1 << 17;  // ~50 cycles per poll: a few ms

template <int COLS>
struct CfSmem {
    static constexpr int RHS_ROWS = COLS / CH_T + 1;      // tile rows of right-hand sides: COLS state columns + the residual row
    static constexpr int YT_LD = RHS_ROWS * CH_T + 1;
    union {
        double Lp[CH_NT][CH_T][CH_T][CH_NT + 1];  // Lp[J][r][j][TI] = v(row 4TI+r, col 4J+j): every finished panel tile is kept
        double Yt[CH_R][YT_LD];                   // afterwards: scaled rows of Y for this CTA's columns (+ the residual z)
    };
    double Dc[CH_NT][CH_T];          // reciprocal pivots of block column J
    double Hand[CH_NT][2][16];       // tiles handed to the chain warp: [J][0] = (J, J-1), [J][1] = (J, J), row-major 4x4
    double C[CH_R / 2][6];
    double Inv[CH_R];
    int Idx[CH_R / 2];
    int flagP[CH_NT][CH_NT + 1];     // flagP[J][TI]: panel tile (TI, J) published (TI > J); flagP[J][J]: diagonal J published
    int hand[CH_NT][2];              // Hand[J][h] written
    int abortFlag;
    int pad_;
};
template <int COLS>
__host__ __device__ constexpr int cf_stage_off() { return ((int)sizeof(CfSmem<COLS>) + 127) & ~127; }
template <int COLS>
__host__ __device__ constexpr int cf_smem_bytes() { return cf_stage_off<COLS>() + (int)sizeof(ChunkStage); }

// warp -> role: 0 chain, 1 S tile warp (idx 0..4), 2 RHS warp (idx 0..4), 3 idle
__device__ __forceinline__ int cf_role(int warp, int& idx) {
    idx = 0;
    switch (warp) {
        case 0: return 0;
        case 1: idx = 0; return 1;
        case 2: idx = 1; return 1;
        case 3: idx = 2; return 1;
        case 5: idx = 3; return 1;
        case 6: idx = 4; return 1;
        case 7: idx = 0; return 2;
        case 9: idx = 1; return 2;
        case 10: idx = 2; return 2;
        case 11: idx = 3; return 2;
        case 13: idx = 4; return 2;
        default: return 3;
    }
}

// warp-uniform bounded wait: every lane with need != 0 waits for both of its flags.  Returns false after a time-out / abort.
__device__ __forceinline__ bool cf_wait(bool need, const int* fa, const int* fb, int* abortFlag) {
    int spins = 0;
    for (;;) {
        const bool ok = !need || (flag_acquire(fa) != 0 && flag_acquire(fb) != 0);
        if (__all_sync(0xffffffffu, ok)) return true;
        if (++spins > CF_SPIN_LIMIT || ((spins & 63) == 0 && flag_acquire(abortFlag) != 0)) {
            flag_release(abortFlag, 1);
            return false;
        }
    }
}

// panel tile against a diagonal tile: a[r][k] -= (a[r][j] c[j]) d[k][j] for j < k (unscaled columns)
__device__ __forceinline__ void cf_panel_op(double (&a)[CH_T][CH_T], const double (&d)[CH_T][CH_T], const double (&c)[CH_T]) {
#pragma unroll
    for (int j = 0; j < CH_T; ++j)
#pragma unroll
        for (int k = j + 1; k < CH_T; ++k)
#pragma unroll
            for (int r = 0; r < CH_T; ++r) a[r][k] -= (a[r][j] * c[j]) * d[k][j];
}

template <int COLS>
__global__ void __launch_bounds__(CF_THREADS)
    chunk_factor_df_kernel(const double* __restrict__ Sig, int ld, int dimp, const int* __restrict__ lmOf,
                           const double* __restrict__ Cblk, const double* __restrict__ ytilde, int j0, int bc, double r2,
                           const double* __restrict__ GammaIn, double* __restrict__ GammaOut, double* __restrict__ Y,
                           int* __restrict__ status, const int* __restrict__ guard, int tl, int stage,
                           const __grid_constant__ CUtensorMap sigMap) {
    using Smem = CfSmem<COLS>;
    constexpr int RHS_ROWS = Smem::RHS_ROWS;
    extern __shared__ __align__(128) unsigned char chunk_smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(chunk_smem_raw);
    ChunkStage& stg = *reinterpret_cast<ChunkStage*>(chunk_smem_raw + cf_stage_off<COLS>());
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rc = 2 * bc;
    CH_STAMP(0);
    // Cblk / lmOf come from meas_kernel and the frame upload, several launches back: staged ahead of the dependency wait
    for (int t = tid; t < bc * 6; t += CF_THREADS) sm.C[t / 6][t % 6] = Cblk[6 * (size_t)j0 + t];
    for (int t = tid; t < bc; t += CF_THREADS) sm.Idx[t] = SOFF + 3 * lmOf[j0 + t];
    for (int t = tid; t < CH_NT * (CH_NT + 1); t += CF_THREADS) (&sm.flagP[0][0])[t] = 0;
    for (int t = tid; t < CH_NT * 2; t += CF_THREADS) (&sm.hand[0][0])[t] = 0;
    if (tid == 0) {
        sm.abortFlag = 0;
        if (stage) {
            mbar_init(&stg.bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
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
    int ridx;
    int role = cf_role(warp, ridx);
    if (role == 2 && ridx * 2 >= RHS_ROWS) role = 3;  // COLS = 16: five half-warps of right-hand sides, warps 11 / 13 have none
    const int sbase = blockIdx.x * COLS;
    const int nJ = (rc + CH_T - 1) / CH_T;
    double a[CH_T][CH_T];
#pragma unroll
    for (int r = 0; r < CH_T; ++r)
#pragma unroll
        for (int c = 0; c < CH_T; ++c) a[r][c] = 0.0;

    if (role == 1) {
        // ================================ S tile warps ================================
        const int s = ridx * 32 + lane;
        const bool owner = s < CH_TILES;
        int TI = 0, TK = 0;
        if (owner) tri_decode_cm(s, CH_NT, TI, TK);
        const bool active = owner && TI < nJ;
        if (active) {
            // S tile = 2x2 landmark pairs: rows from landmarks 2TI, 2TI+1; columns from 2TK, 2TK+1.  All loads first.
            double P[2][2][9];
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    const int j = 2 * TI + u, k = 2 * TK + v;
                    if (j < bc && k < bc && !staged) {
                        const double* sp = Sig + (size_t)sm.Idx[k] * ld + sm.Idx[j];
#pragma unroll
                        for (int b = 0; b < 3; ++b)
#pragma unroll
                            for (int aa = 0; aa < 3; ++aa) P[u][v][aa * 3 + b] = sp[(size_t)b * ld + aa];
                    }
                }
            if (staged) {
                mbar_wait(&stg.bar, 0);
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        const int j = 2 * TI + u, k = 2 * TK + v;
                        if (j < bc && k < bc) {
#pragma unroll
                            for (int aa = 0; aa < 3; ++aa)
#pragma unroll
                                for (int b = 0; b < 3; ++b) P[u][v][aa * 3 + b] = stg.S[3 * k + b][3 * j + aa];  // same entries as the gather
                        }
                    }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    const int j = 2 * TI + u, k = 2 * TK + v;
                    if (j < bc && k < bc) {
                        double T[6];
#pragma unroll
                        for (int e = 0; e < 2; ++e)
#pragma unroll
                            for (int b = 0; b < 3; ++b)
                                T[e * 3 + b] = sm.C[j][3 * e] * P[u][v][b] + sm.C[j][3 * e + 1] * P[u][v][3 + b] + sm.C[j][3 * e + 2] * P[u][v][6 + b];
#pragma unroll
                        for (int e = 0; e < 2; ++e)
#pragma unroll
                            for (int f = 0; f < 2; ++f)
                                a[2 * u + e][2 * v + f] = T[e * 3] * sm.C[k][3 * f] + T[e * 3 + 1] * sm.C[k][3 * f + 1] + T[e * 3 + 2] * sm.C[k][3 * f + 2];
                    }
                }
            if (TI == TK) {
#pragma unroll
                for (int c = 0; c < CH_T; ++c) {
                    if (CH_T * TI + c < rc)
                        a[c][c] += r2;
                    else
                        a[c][c] = 1.0;  // identity padding of a short last chunk
                }
            }
        }
        CH_STAMP(108);  // thread 160 = first lane of S warp 3: its tile is projected
        const bool band = TI == TK || TI == TK + 1;     // handed to the chain warp instead of finishing as a panel tile
        const int jEnd = (TI == TK) ? TK - 1 : TK;      // this lane applies the updates of block columns j < jEnd
        auto hand_over = [&]() {
            const int h = (TI == TK) ? 1 : 0;
#pragma unroll
            for (int r = 0; r < CH_T; ++r)
#pragma unroll
                for (int c = 0; c < CH_T; ++c) sm.Hand[TI][h][r * CH_T + c] = a[r][c];
            flag_release(&sm.hand[TI][h], 1);
        };
        if (active && band && jEnd <= 0) hand_over();   // tiles (0,0), (1,0), (1,1): nothing to apply first
        bool alive = true;
        for (int j = 0; j < nJ && alive; ++j) {
            // ---- phase A: the panel tiles of block column j (not the one next to the diagonal: the chain warp finishes it)
            const bool doPanel = active && !band && TK == j;
            if (__any_sync(0xffffffffu, doPanel)) {
                if (!cf_wait(doPanel, &sm.flagP[j][j], &sm.flagP[j][j], &sm.abortFlag)) { alive = false; break; }
                if (doPanel) {
                    double c[CH_T], d[CH_T][CH_T];
#pragma unroll
                    for (int q = 0; q < CH_T; ++q) c[q] = sm.Dc[j][q];
#pragma unroll
                    for (int i = 0; i < CH_T; ++i)
#pragma unroll
                        for (int q = 0; q < CH_T; ++q) d[i][q] = sm.Lp[j][i][q][j];
                    cf_panel_op(a, d, c);
#pragma unroll
                    for (int r = 0; r < CH_T; ++r)
#pragma unroll
                        for (int q = 0; q < CH_T; ++q) sm.Lp[j][r][q][TI] = a[r][q];
                    flag_release(&sm.flagP[j][TI], 1);
                }
            }
            // ---- phase B: rank-4 update of the tiles to the right of block column j
            const bool doUpd = active && j < jEnd;
            if (__any_sync(0xffffffffu, doUpd)) {
                if (!cf_wait(doUpd, &sm.flagP[j][TI], &sm.flagP[j][TK], &sm.abortFlag)) { alive = false; break; }
                if (doUpd) {
                    double li[CH_T][CH_T], pk[CH_T][CH_T];
#pragma unroll
                    for (int q = 0; q < CH_T; ++q) {
                        const double c = sm.Dc[j][q];
#pragma unroll
                        for (int r = 0; r < CH_T; ++r) li[r][q] = sm.Lp[j][r][q][TI] * c;
                    }
#pragma unroll
                    for (int cc = 0; cc < CH_T; ++cc)
#pragma unroll
                        for (int q = 0; q < CH_T; ++q) pk[cc][q] = sm.Lp[j][cc][q][TK];
#pragma unroll
                    for (int r = 0; r < CH_T; ++r)
#pragma unroll
                        for (int cc = 0; cc < CH_T; ++cc) {
                            double acc = a[r][cc];
#pragma unroll
                            for (int q = 0; q < CH_T; ++q) acc -= li[r][q] * pk[cc][q];
                            a[r][cc] = acc;
                        }
                }
            }
            // ---- phase C: tiles next to the diagonal are current now -> to the chain warp
            if (active && band && jEnd == j + 1) hand_over();
        }
    } else if (role == 0) {
        // ================================ chain warp ================================
        // every lane runs the same dependent chain (no divergence, no exchange); lane 0 publishes
        CH_STAMP(2);
        double d[CH_T][CH_T], c[CH_T];
#pragma unroll
        for (int i = 0; i < CH_T; ++i) {
            c[i] = 0.0;
#pragma unroll
            for (int q = 0; q < CH_T; ++q) d[i][q] = 0.0;
        }
        for (int J = 0; J < nJ; ++J) {
            CH_FINE(4 * J);
            if (!cf_wait(true, &sm.hand[J][1], J > 0 ? &sm.hand[J][0] : &sm.hand[J][1], &sm.abortFlag)) break;
            CH_FINE(4 * J + 1);
            double B[CH_T][CH_T];
#pragma unroll
            for (int r = 0; r < CH_T; ++r)
#pragma unroll
                for (int q = 0; q < CH_T; ++q) B[r][q] = sm.Hand[J][1][r * CH_T + q];
            if (J > 0) {
                double A[CH_T][CH_T];
#pragma unroll
                for (int r = 0; r < CH_T; ++r)
#pragma unroll
                    for (int q = 0; q < CH_T; ++q) A[r][q] = sm.Hand[J][0][r * CH_T + q];
                cf_panel_op(A, d, c);  // against diagonal J-1
                if (lane == 0) {
#pragma unroll
                    for (int r = 0; r < CH_T; ++r)
#pragma unroll
                        for (int q = 0; q < CH_T; ++q) sm.Lp[J - 1][r][q][J] = A[r][q];
                    flag_release(&sm.flagP[J - 1][J], 1);
                }
                // rank-4 update of the diagonal tile by its own panel tile (lower part)
#pragma unroll
                for (int r = 0; r < CH_T; ++r)
#pragma unroll
                    for (int cc = 0; cc <= r; ++cc) {
                        double acc = B[r][cc];
#pragma unroll
                        for (int q = 0; q < CH_T; ++q) acc -= (A[r][q] * c[q]) * A[cc][q];
                        B[r][cc] = acc;
                    }
            }
            CH_FINE(4 * J + 2);
            {
                // 4x4 diagonal tile: fraction-free elimination (products only) and the four pivot reciprocals side by side
                const double a00 = B[0][0], a10 = B[1][0], a20 = B[2][0], a30 = B[3][0];
                const double m11 = B[1][1] * a00 - a10 * a10, m21 = B[2][1] * a00 - a20 * a10, m22 = B[2][2] * a00 - a20 * a20;
                const double m31 = B[3][1] * a00 - a30 * a10, m32 = B[3][2] * a00 - a30 * a20, m33 = B[3][3] * a00 - a30 * a30;
                const double n22 = m22 * m11 - m21 * m21, n32 = m32 * m11 - m31 * m21, n33 = m33 * m11 - m31 * m31;
                const double p33 = n33 * n22 - n32 * n32;
                const double r0 = fast_rcp(a00), r1 = fast_rcp(m11), r2_ = fast_rcp(n22), r3 = fast_rcp(p33);
                const double s2 = r0 * r1, s3 = s2 * r2_, e1 = a00 * m11;
                // unscaled columns v_ij = L_ij L_jj (the Schur-complement values) and 1 / v_jj
                d[0][0] = a00;
                d[1][0] = a10;
                d[2][0] = a20;
                d[3][0] = a30;
                d[1][1] = m11 * r0;
                d[2][1] = m21 * r0;
                d[3][1] = m31 * r0;
                d[2][2] = n22 * s2;
                d[3][2] = n32 * s2;
                d[3][3] = p33 * s3;
                c[0] = r0;
                c[1] = a00 * r1;
                c[2] = e1 * r2_;
                c[3] = (e1 * n22) * r3;
            }
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < CH_T; ++q) sm.Dc[J][q] = c[q];
#pragma unroll
                for (int i = 0; i < CH_T; ++i)
#pragma unroll
                    for (int q = 0; q <= i; ++q) sm.Lp[J][i][q][J] = d[i][q];
                flag_release(&sm.flagP[J][J], 1);
            }
            CH_FINE(4 * J + 3);
            // 1 / L_kk (off the chain: nobody reads it before the closing barrier)
            if (lane < CH_T) {
                const double piv = lane == 0 ? d[0][0] : lane == 1 ? d[1][1] : lane == 2 ? d[2][2] : d[3][3];
                const int k = CH_T * J + lane;
                if (!(piv > 0.0)) {
                    if (blockIdx.x == 0) atomicOr(status, 1);
                    sm.Inv[k] = 1.0;
                } else {
                    sm.Inv[k] = 1.0 / sqrt(piv);
                }
            }
        }
        for (int k = CH_T * nJ + lane; k < CH_R; k += 32) sm.Inv[k] = 1.0;
        CH_STAMP(3);
    } else if (role == 2) {
        // ================================ RHS warps ================================
        const int q = ridx * 32 + lane;
        const int trow = q / CH_NT, TK = q % CH_NT;
        const bool isRhs = trow < RHS_ROWS;
        if (isRhs && trow < COLS / CH_T) {
            // rows s = sbase + 4 trow + r (state columns of W_c), columns k = 4TK + c from landmarks 2TK, 2TK+1
            const int s0 = sbase + CH_T * trow;
            double w[2][3][CH_T];
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const int j = 2 * TK + v;
                if (j < bc && s0 < dimp) {
                    const double* sp = Sig + (size_t)sm.Idx[j] * ld + s0;  // Sigma[s0.., cols of j] (symmetric storage)
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        const double2 p01 = *reinterpret_cast<const double2*>(sp + (size_t)b * ld);
                        const double2 p23 = *reinterpret_cast<const double2*>(sp + (size_t)b * ld + 2);
                        w[v][b][0] = p01.x;
                        w[v][b][1] = p01.y;
                        w[v][b][2] = p23.x;
                        w[v][b][3] = p23.y;
                    }
                }
            }
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const int j = 2 * TK + v;
                if (j < bc && s0 < dimp) {
#pragma unroll
                    for (int r = 0; r < CH_T; ++r)
#pragma unroll
                        for (int e = 0; e < 2; ++e)
                            a[r][2 * v + e] = (s0 + r < dimp) ? sm.C[j][3 * e] * w[v][0][r] + sm.C[j][3 * e + 1] * w[v][1][r] + sm.C[j][3 * e + 2] * w[v][2][r] : 0.0;
                }
            }
        } else if (isRhs) {
            // residual row (r = 0 of the last tile row): ytilde_c - C_c Gamma
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const int j = 2 * TK + v;
                if (j < bc) {
                    const int g = sm.Idx[j];
                    const double g0 = GammaIn[g], g1 = GammaIn[g + 1], g2 = GammaIn[g + 2];
                    a[0][2 * v] = ytilde[2 * (j0 + j)] - (sm.C[j][0] * g0 + sm.C[j][1] * g1 + sm.C[j][2] * g2);
                    a[0][2 * v + 1] = ytilde[2 * (j0 + j) + 1] - (sm.C[j][3] * g0 + sm.C[j][4] * g1 + sm.C[j][5] * g2);
                }
            }
        }
        for (int J = 0; J < nJ; ++J) {
            if (!cf_wait(true, &sm.flagP[J][J], &sm.flagP[J][J], &sm.abortFlag)) break;
            // the lane holding tile column J finishes its four columns ...
            double c[CH_T];
#pragma unroll
            for (int j = 0; j < CH_T; ++j) c[j] = sm.Dc[J][j];
            if (TK == J) {
                double d[CH_T][CH_T];
#pragma unroll
                for (int i = 0; i < CH_T; ++i)
#pragma unroll
                    for (int j = 0; j < CH_T; ++j) d[i][j] = sm.Lp[J][i][j][J];
                cf_panel_op(a, d, c);
            }
            // ... and hands them to the rest of its half-warp (same tile row)
            double li[CH_T][CH_T];
#pragma unroll
            for (int r = 0; r < CH_T; ++r)
#pragma unroll
                for (int j = 0; j < CH_T; ++j) li[r][j] = __shfl_sync(0xffffffffu, a[r][j], J, CH_NT) * c[j];
            const bool upd = isRhs && TK > J && TK < nJ;
            if (!cf_wait(upd, &sm.flagP[J][TK], &sm.flagP[J][TK], &sm.abortFlag)) break;
            if (upd) {
                double pk[CH_T][CH_T];
#pragma unroll
                for (int cc = 0; cc < CH_T; ++cc)
#pragma unroll
                    for (int j = 0; j < CH_T; ++j) pk[cc][j] = sm.Lp[J][cc][j][TK];
#pragma unroll
                for (int r = 0; r < CH_T; ++r)
#pragma unroll
                    for (int cc = 0; cc < CH_T; ++cc) {
                        double acc = a[r][cc];
#pragma unroll
                        for (int j = 0; j < CH_T; ++j) acc -= li[r][j] * pk[cc][j];
                        a[r][cc] = acc;
                    }
            }
        }
    }
    CH_STAMP(109);
    __syncthreads();  // Inv published, every tile final
    CH_STAMP(4);
    if (tid == 0 && sm.abortFlag != 0 && blockIdx.x == 0) atomicOr(status, 8);  // a bounded wait ran out: results are invalid
    // Y[k][s] = v_sk / L_kk, staged so that the global store and the Gamma dot products run in a fixed order
    if (role == 2) {
        const int q = ridx * 32 + lane;
        const int trow = q / CH_NT, TK = q % CH_NT;
        if (trow < RHS_ROWS) {
#pragma unroll
            for (int c = 0; c < CH_T; ++c) {
                const double sc = sm.Inv[CH_T * TK + c];
#pragma unroll
                for (int r = 0; r < CH_T; ++r) sm.Yt[CH_T * TK + c][CH_T * trow + r] = a[r][c] * sc;
            }
        }
    }
    __syncthreads();
    CH_STAMP(5);
    // all CH_R rows are written (zero beyond rc and for the pad columns s >= dimp): the downdate reads whole tiles
    for (int t = tid; t < CH_R * COLS; t += CF_THREADS) {
        const int k = t / COLS, sl = t % COLS;
        Y[yb_index(k, sbase + sl)] = sm.Yt[k][sl];
    }
    // Gamma += Y_c^T z_c: eight partial sums per state column (fixed order), combined by the first COLS threads
    {
        const int col = tid % COLS, part = tid / COLS;
        double* gpart = &sm.Yt[0][0] + CH_R * Smem::YT_LD;  // behind Yt inside the union
        if (part < 8) {
            double g = 0.0;
#pragma unroll
            for (int k = 0; k < CH_R / 8; ++k) g += sm.Yt[8 * part + k][col] * sm.Yt[8 * part + k][COLS];
            gpart[part * COLS + col] = g;
        }
        __syncthreads();
        if (tid < COLS && sbase + tid < dimp) {
            const double t = GammaIn[sbase + tid];
            const double acc = ((gpart[tid] + gpart[COLS + tid]) + (gpart[2 * COLS + tid] + gpart[3 * COLS + tid])) +
                               ((gpart[4 * COLS + tid] + gpart[5 * COLS + tid]) + (gpart[6 * COLS + tid] + gpart[7 * COLS + tid]));
            GammaOut[sbase + tid] = t + acc;  // ping-pong: other CTAs may still be reading GammaIn
        }
    }
    CH_STAMP(6);
    TL_MARK(tl, 1);
}
