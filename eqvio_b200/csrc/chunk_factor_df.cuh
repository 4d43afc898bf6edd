// chunk_factor_df.cuh -- included from kernels.cuh (needs its ChunkStage / TMA / mbarrier helpers).
// ------------------------------------------------------------------------------------------------
// chunk_factor_df_kernel: the chunk factor step of the sequential-chunk correction (performVisionUpdate,
// VIO_eqf.cpp:105-135; see the derivation above chunk_factor_kernel) as a DATAFLOW elimination.
//
// Same arithmetic as chunk_factor_kernel -- the augmented matrix [S_c; W_c^T (this CTA's state columns); r^T] in 4x4
// register tiles, right-looking, unscaled columns, fraction-free diagonal tiles, every operation in the same order (the two
// kernels agree bit for bit) -- but no CTA-wide barrier inside the elimination.  The 64 pivots of S_c are a chain of 16
// diagonal tiles; the round-1 kernel spends ~1150 cycles per block column on diag -> barrier -> panel -> barrier -> update
// hand-offs through five warps.  Here the tile rows are I = 0..15 (S_c) and 16.. (right-hand sides: 4 state columns per
// tile row, the residual in the last one), tile (I, K) lives in one lane, and the warps specialise:
//   * chain warp: walks the diagonal.  At step J it takes tile (J, J-1) and tile (J, J) from the band warp (current
//     through block column J-2), finishes the panel tile (J, J-1) against diagonal J-1 and publishes it, applies its rank-4
//     update to (J, J), eliminates inside (J, J) and publishes the diagonal.  Every lane runs the same dependent fp64
//     chain (no divergence, no exchange: a row-parallel version with shuffles measured slower); it shares its scheduler with
//     nobody who has fp64 work (a warp-wide DFMA occupies the sub-partition's pipe for two cycles).
//   * band warp: the 31 tiles next to the diagonal, (J, J) and (J, J-1).  They only take rank-4 updates and are handed to
//     the chain warp when current; being alone in their warp they never queue behind bulk work.
//   * bulk warps: every other tile, column-major (S rows I >= K + 2, then the right-hand-side rows, column by column), so
//     whole warps retire as the elimination moves right.  Per block column j: the owners of column j wait for diagonal j,
//     finish their panel tile and publish it; everyone to the right waits for the two panel tiles it needs and applies the
//     rank-4 update.  Right-hand-side tiles keep their final value in registers for Y.
//   * aux warp: 1 / L_kk of every finished diagonal tile (a division and a square root: ~570 cycles, off the chain).
// Hand-offs go through single-use mbarriers (one per published tile, arrival count 1, phase 0): the producer stores its tile
// and arrives (SASS: STS ... SYNCS.ARRIVE, no MEMBAR -- a st.release.cta flag costs a MEMBAR.ALL.CTA, hundreds of cycles
// with stores in flight), the consumer blocks in mbarrier.try_wait (SYNCS.PHASECHK.TRYWAIT: the warp sleeps in hardware and
// takes no issue slots from warps that have work) and then reads the tile with plain 16-byte loads.  Every published
// tile is written once and never overwritten: no WAR hazard.  All waits are warp-uniform (__all_sync) and bounded -- a wait
// longer than CF_WAIT_CYCLES raises status bit 8 and the CTA runs to its end (no hang).  Dependencies inside one warp always
// point backwards in program order (phase A of step j publishes what phase B of step j consumes), so the uniform waits
// cannot deadlock.
// COLS = state columns per CTA (16: more, lighter CTAs -- 52 at N = 256; 32 when 16 would exceed one wave of CTAs).
// ------------------------------------------------------------------------------------------------
#pragma once

// warp ids: 0 chain | 1 band | 14 aux | bulk: 2,3,6,7,10,11 (+ 5, 9 when 32 state columns need them) | 4, 8, 12, 13 idle.  The chain
// and the band warp keep their sub-partitions (warp id % 4 = 0 / 1) to themselves: bulk warps beside them cost the chain ~150
// cycles per step (measured), the fp64-heavy aux warp sits with the bulk warps.
constexpr int CF_WARPS = 15;
constexpr int CF_THREADS = CF_WARPS * 32;
constexpr int CF_MAX_BULK = 8;
constexpr long long CF_WAIT_CYCLES = 40000000;  // ~20 ms
constexpr int CF_TILE_LD = 18;  // doubles per published 4x4 tile (16 + 2 pad): 16-byte loads of consecutive tiles spread over all banks

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

template <int COLS>
struct CfSmem {
    static constexpr int RHS_ROWS = COLS / CH_T + 1;      // tile rows of right-hand sides: COLS state columns + the residual row
    static constexpr int ROWS = CH_NT + RHS_ROWS;         // tile rows of the augmented matrix
    static constexpr int YT_LD = RHS_ROWS * CH_T + 1;
    // bulk tiles: column K holds S rows K+2..15 and the RHS_ROWS right-hand-side rows
    static constexpr int BULK_TILES = (CH_NT - 2) * (CH_NT - 1) / 2 + CH_NT * RHS_ROWS;
    static constexpr int BULK_WARPS = (BULK_TILES + 31) / 32;
    static_assert(BULK_WARPS <= CF_MAX_BULK, "not enough bulk warps");
    union {
        double Lp[CH_NT][ROWS][CF_TILE_LD];       // Lp[J][I][4 r + j] = v(row 4I+r, col 4J+j): every finished panel tile is kept
        double Yt[CH_R][YT_LD];                   // afterwards: scaled rows of Y for this CTA's columns (+ the residual z)
    };
    double Ls[CH_NT][ROWS][CF_TILE_LD];           // the same tiles with column j scaled by 1 / v_jj (the row-side operand of the updates)
    alignas(16) double Dc[CH_NT][CH_T];           // reciprocal pivots of block column J
    alignas(16) double Hand[CH_NT][2][16];        // tiles handed to the chain warp: [J][0] = (J, J-1), [J][1] = (J, J), row-major 4x4
    double C[CH_R / 2][6];
    double Inv[CH_R];
    alignas(8) uint64_t barP[CH_NT][ROWS];        // barP[J][I]: panel tile (I, J) published (I > J); barP[J][J]: diagonal J (+ Dc[J])
    alignas(8) uint64_t barH[CH_NT][2];           // Hand[J][h] written
    int Idx[CH_R / 2];
    int abortFlag;
    int pad_;
};
template <int COLS>
__host__ __device__ constexpr int cf_stage_off() { return ((int)sizeof(CfSmem<COLS>) + 127) & ~127; }
template <int COLS>
__host__ __device__ constexpr int cf_smem_bytes() { return cf_stage_off<COLS>() + (int)sizeof(ChunkStage); }

// warp -> role: 0 chain, 1 band, 2 bulk (idx), 3 idle, 4 aux
__device__ __forceinline__ int cf_role(int warp, int& idx) {
    idx = 0;
    switch (warp) {
        case 0: return 0;
        case 1: return 1;
        case 14: return 4;
        case 2: idx = 0; return 2;
        case 3: idx = 1; return 2;
        case 6: idx = 2; return 2;
        case 7: idx = 3; return 2;
        case 10: idx = 4; return 2;
        case 11: idx = 5; return 2;
        case 5: idx = 6; return 2;
        case 9: idx = 7; return 2;
        default: return 3;
    }
}

// warp-uniform bounded wait: every lane with need waits for its two tiles (barriers a and b, possibly the same).  Returns false
// after a time-out / abort.
__device__ __forceinline__ bool cf_wait(bool need, uint64_t* a, uint64_t* b, int* abortFlag) {
    bool ok = !need;
    long long t0 = 0;
    for (int spins = 0;; ++spins) {
        if (!ok) ok = mbar_try_wait(a, 0) && mbar_try_wait(b, 0);
        if (__all_sync(0xffffffffu, ok)) return true;
        if (spins == 0) t0 = clock64();
        if ((spins & 15) == 15 && (clock64() - t0 > CF_WAIT_CYCLES || *(volatile int*)abortFlag != 0)) {
            *(volatile int*)abortFlag = 1;
            return false;
        }
    }
}
// 4x4 tile <-> 16 contiguous doubles (row-major), 16-byte accesses
__device__ __forceinline__ void cf_load_tile(const double* p, double (&t)[CH_T][CH_T]) {
#pragma unroll
    for (int r = 0; r < CH_T; ++r) {
        const double2 v01 = *reinterpret_cast<const double2*>(p + r * CH_T);
        const double2 v23 = *reinterpret_cast<const double2*>(p + r * CH_T + 2);
        t[r][0] = v01.x, t[r][1] = v01.y, t[r][2] = v23.x, t[r][3] = v23.y;
    }
}
__device__ __forceinline__ void cf_store_tile(double* p, const double (&t)[CH_T][CH_T]) {
#pragma unroll
    for (int r = 0; r < CH_T; ++r) {
        *reinterpret_cast<double2*>(p + r * CH_T) = make_double2(t[r][0], t[r][1]);
        *reinterpret_cast<double2*>(p + r * CH_T + 2) = make_double2(t[r][2], t[r][3]);
    }
}
__device__ __forceinline__ void cf_load4(const double* p, double (&c)[CH_T]) {
    const double2 c01 = *reinterpret_cast<const double2*>(p);
    const double2 c23 = *reinterpret_cast<const double2*>(p + 2);
    c[0] = c01.x, c[1] = c01.y, c[2] = c23.x, c[3] = c23.y;
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
// rank-4 update of a tile by block column j: a[r][cc] -= sum_q (p_i[r][q] c[q]) p_k[cc][q]; pi holds the scaled row-side tile
__device__ __forceinline__ void cf_update(double (&a)[CH_T][CH_T], const double (&pi)[CH_T][CH_T], const double (&pk)[CH_T][CH_T]) {
#pragma unroll
    for (int r = 0; r < CH_T; ++r)
#pragma unroll
        for (int cc = 0; cc < CH_T; ++cc) {
            double acc = a[r][cc];
#pragma unroll
            for (int q = 0; q < CH_T; ++q) acc -= pi[r][q] * pk[cc][q];
            a[r][cc] = acc;
        }
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
    for (int t = tid; t < CH_NT * Smem::ROWS; t += CF_THREADS) mbar_init(&sm.barP[0][0] + t, 1);
    for (int t = tid; t < CH_NT * 2; t += CF_THREADS) mbar_init(&sm.barH[0][0] + t, 1);
    if (tid == 0) {
        sm.abortFlag = 0;
        if (stage) mbar_init(&stg.bar, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // made visible to the CTA by the barrier below
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
    if (role == 2 && ridx >= Smem::BULK_WARPS) role = 3;
    const int sbase = blockIdx.x * COLS;
    const int nJ = (rc + CH_T - 1) / CH_T;

    // ---- tile of this lane: (TI, TK), TI < 16: tile of S_c, TI >= 16: right-hand-side tile row TI - 16 ----
    int TI = 0, TK = 0;
    bool owner = false;
    if (role == 1) {  // band warp: lanes 0..15 the diagonal tiles, lanes 16..30 the tiles below them
        owner = lane < 2 * CH_NT - 1;
        TI = lane < CH_NT ? lane : lane - (CH_NT - 1);
        TK = lane < CH_NT ? lane : lane - CH_NT;
    } else if (role == 2) {  // bulk: column K = S rows K+2..15, then the right-hand-side rows
        int t = ridx * 32 + lane;
        owner = t < Smem::BULK_TILES;
        if (owner) {
            int K = 0;
            for (;; ++K) {
                const int nS = K + 2 < CH_NT ? CH_NT - 2 - K : 0;
                if (t < nS + RHS_ROWS) {
                    TI = t < nS ? K + 2 + t : CH_NT + (t - nS);
                    break;
                }
                t -= nS + RHS_ROWS;
            }
            TK = K;
        }
    }
    const bool isRhs = TI >= CH_NT;
    const int trow = TI - CH_NT;
    const bool active = owner && TK < nJ && (isRhs || TI < nJ);
    double a[CH_T][CH_T];
#pragma unroll
    for (int r = 0; r < CH_T; ++r)
#pragma unroll
        for (int c = 0; c < CH_T; ++c) a[r][c] = 0.0;

    // ---- initial tile values ----
    if (active && !isRhs) {
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
    } else if (active && trow < COLS / CH_T) {
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
    } else if (active) {
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
    CH_STAMP(2);

    if (role == 1) {
        // ================================ band warp ================================
        const int h = (TI == TK) ? 1 : 0;
        const int jEnd = (TI == TK) ? TK - 1 : TK;  // this lane applies the updates of block columns j < jEnd, then hands over
        if (active && jEnd <= 0) {                  // tiles (0,0), (1,0), (1,1): nothing to apply first
            cf_store_tile(&sm.Hand[TI][h][0], a);
            mbar_arrive(&sm.barH[TI][h]);
        }
        for (int j = 0; j + 2 < nJ; ++j) {
            const bool doUpd = active && j < jEnd;
            if (!cf_wait(doUpd, &sm.barP[j][TI], &sm.barP[j][TK], &sm.abortFlag)) break;
            if (doUpd) {
                double pi[CH_T][CH_T], pk[CH_T][CH_T];
                cf_load_tile(&sm.Ls[j][TI][0], pi);
                cf_load_tile(&sm.Lp[j][TK][0], pk);
                cf_update(a, pi, pk);
                if (jEnd == j + 1) {
                    cf_store_tile(&sm.Hand[TI][h][0], a);
                    mbar_arrive(&sm.barH[TI][h]);
                }
            }
        }
    } else if (role == 2) {
        // ================================ bulk warps ================================
        for (int j = 0; j < nJ; ++j) {
            // ---- phase A: the panel tiles of block column j
            const bool doPanel = active && TK == j;
            if (__any_sync(0xffffffffu, doPanel)) {
                if (!cf_wait(doPanel, &sm.barP[j][j], &sm.barP[j][j], &sm.abortFlag)) break;
                if (doPanel) {
                    double c[CH_T], d[CH_T][CH_T];
                    cf_load_tile(&sm.Lp[j][j][0], d);
                    cf_load4(&sm.Dc[j][0], c);
                    cf_panel_op(a, d, c);
                    // published for the tiles to the right in the same tile row (a right-hand-side tile is final now and also
                    // stays in registers for Y)
                    cf_store_tile(&sm.Lp[j][TI][0], a);
                    double as[CH_T][CH_T];
#pragma unroll
                    for (int r = 0; r < CH_T; ++r)
#pragma unroll
                        for (int q = 0; q < CH_T; ++q) as[r][q] = a[r][q] * c[q];
                    cf_store_tile(&sm.Ls[j][TI][0], as);
                    mbar_arrive(&sm.barP[j][TI]);
                }
            }
            // ---- phase B: rank-4 update of the tiles to the right of block column j
            const bool doUpd = active && TK > j;
            if (!__any_sync(0xffffffffu, doUpd)) break;  // every tile of this warp is final
            {
                if (!cf_wait(doUpd, &sm.barP[j][TI], &sm.barP[j][TK], &sm.abortFlag)) break;
                if (doUpd) {
                    double pi[CH_T][CH_T], pk[CH_T][CH_T];
                    cf_load_tile(&sm.Ls[j][TI][0], pi);
                    cf_load_tile(&sm.Lp[j][TK][0], pk);
                    cf_update(a, pi, pk);
                }
            }
        }
    } else if (role == 0) {
        // ================================ chain warp ================================
        double d[CH_T][CH_T], c[CH_T];
#pragma unroll
        for (int i = 0; i < CH_T; ++i) {
            c[i] = 0.0;
#pragma unroll
            for (int q = 0; q < CH_T; ++q) d[i][q] = 0.0;
        }
        for (int J = 0; J < nJ; ++J) {
            CH_FINE(4 * J);
            if (!cf_wait(true, &sm.barH[J][1], &sm.barH[J][J > 0 ? 0 : 1], &sm.abortFlag)) break;
            CH_FINE(4 * J + 1);
            double B[CH_T][CH_T];
            cf_load_tile(&sm.Hand[J][1][0], B);
            if (J > 0) {
                double A[CH_T][CH_T];
                cf_load_tile(&sm.Hand[J][0][0], A);
                cf_panel_op(A, d, c);  // against diagonal J-1
                double As[CH_T][CH_T];
#pragma unroll
                for (int r = 0; r < CH_T; ++r)
#pragma unroll
                    for (int q = 0; q < CH_T; ++q) As[r][q] = A[r][q] * c[q];
                if (lane == 0) {
                    cf_store_tile(&sm.Lp[J - 1][J][0], A);
                    cf_store_tile(&sm.Ls[J - 1][J][0], As);
                    mbar_arrive(&sm.barP[J - 1][J]);
                }
                // rank-4 update of the diagonal tile by its own panel tile (lower part)
#pragma unroll
                for (int r = 0; r < CH_T; ++r)
#pragma unroll
                    for (int cc = 0; cc <= r; ++cc) {
                        double acc = B[r][cc];
#pragma unroll
                        for (int q = 0; q < CH_T; ++q) acc -= As[r][q] * A[cc][q];
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
                *reinterpret_cast<double2*>(&sm.Dc[J][0]) = make_double2(c[0], c[1]);
                *reinterpret_cast<double2*>(&sm.Dc[J][2]) = make_double2(c[2], c[3]);
                cf_store_tile(&sm.Lp[J][J][0], d);  // (the strictly upper entries are never read)
                mbar_arrive(&sm.barP[J][J]);
            }
            CH_FINE(4 * J + 3);
        }
        CH_STAMP(3);
    } else if (role == 4) {
        // ================================ aux warp ================================
        for (int J = 0; J < nJ; ++J) {
            if (!cf_wait(true, &sm.barP[J][J], &sm.barP[J][J], &sm.abortFlag)) break;
            if (lane < CH_T) {
                const double piv = sm.Lp[J][J][lane * CH_T + lane];
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
    }
    CH_STAMP(109);
    __syncthreads();  // Inv published, every tile final (Lp is dead from here on: Yt overlays it)
    CH_STAMP(4);
    if (tid == 0 && *(volatile int*)&sm.abortFlag != 0 && blockIdx.x == 0) atomicOr(status, 8);  // a bounded wait ran out: results are invalid
    // Y[k][s] = v_sk / L_kk, staged so that the global store and the Gamma dot products run in a fixed order
    if (role == 2 && owner && isRhs) {
#pragma unroll
        for (int c = 0; c < CH_T; ++c) {
            const double sc = sm.Inv[CH_T * TK + c];
#pragma unroll
            for (int r = 0; r < CH_T; ++r) sm.Yt[CH_T * TK + c][CH_T * trow + r] = a[r][c] * sc;
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
