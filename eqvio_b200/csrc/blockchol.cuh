// Correction, block form with look-ahead (EQVIO_TUNE_CORRECTION = 2): performVisionUpdate (VIO_eqf.cpp:105-135) as ONE
// blocked right-looking Cholesky sweep over the augmented matrix
//        Z = [ S ; W^T ],   S = C Sigma C^T + sigma^2 I  (m x m),   W^T = Sigma C^T  (dimp x m, ytilde^T riding in pad row 21)
// with 64-wide block columns.  After block column k:  P = Z[below, blk k] L_kk^-T  (the W rows of P are Y_k^T, row 21 is z_k^T),
// the trailing blocks take  Z_ij -= P_i P_j^T,  and  Sigma -= Y_k^T Y_k,  Gamma += Y_k^T z_k  follow from the same product.
//
// What is on the critical path is only the chain of 64 x 64 diagonal factorizations:
//   bc_diag_kernel   (1 CTA, stream A)   step k:  D = T_kk - P_{k,k-1} P_{k,k-1}^T  with  P_{k,k-1} = T_{k,k-1} L_{k-1,k-1}^-T  (DMMA, the
//                    look-ahead: T_* are current through step k-2), factor D in 4x4 register tiles; hands over L_kk (transposed) and the
//                    inverses of its four diagonal 16 x 16 blocks.
//   bc_panel_kernel  (stream B)          P = Z[rows below, blk k] L_kk^-T  for every row tile (DMMA block substitution) -> Zp (tile-blocked panels)
//   bc_trail_kernel  (stream B)          S / W trailing tiles  Z_ij -= P_i P_j^T,  Sigma tiles -= Y_a Y_b^T,  Gamma   (DMMA)
// Dependencies:  diag(k) -> panel(k) -> trail(k) -> diag(k+2);  diag(k) -> diag(k+1) is a programmatic (PDL) edge.
// The sequential-chunk form (kernels.cuh) needs the downdated Sigma before the next chunk can even be gathered; here the trailing
// work runs beside the next factorization and only ~2 us of DMMA products per block sit between two factorizations.
#pragma once
#include "kernels.cuh"

namespace eqvio {

constexpr int BC_T = 64;                 // block width of the sweep
constexpr int BC_YROW = 21;              // pad row of the state that carries ytilde / z
constexpr int BC_MAX_ROWS = 768;         // largest m served by this form (beyond that the trailing work dominates: sequential chunks)

// What a diagonal step hands to its consumers (the next diagonal step, the panel / next-column kernels), per block k:
//   LT = L_kk transposed, the six 16 x 16 blocks below the diagonal blocks;  XT = X_b transposed, X_b = (b-th diagonal 16 x 16 block
//   of L_kk)^-1, b = 0..3.   P = T L_kk^-T is then a 4-stage block substitution:  P_b = (T_b - sum_{j<b} P_j L_bj^T) X_b^T.
// Stored compactly: ten 16 x 16 blocks of 16 x BC_XLD doubles -- LT block (b, j), j < b, at index b (b - 1) / 2 + j as [k][n] =
// L[16 b + n][16 j + k]; XT block b at index 6 + b as [k][n] = X_b[n][k]  (25.6 KB: one TMA bulk copy for every consumer).
constexpr int BC_XLD = 20;                          // row stride of a block: 20 mod 16 = 4 -> conflict-free fragment reads
constexpr int BC_XBLK = 16 * BC_XLD;
constexpr size_t BC_LX = 10 * BC_XBLK;
// Hand-over of LT | XT from diag(k) to diag(k+1) while diag(k) is still running: the build kernel fills every block with a sentinel
// (a NaN payload no computation produces), the producer overwrites it with plain 8-byte stores as block rows become final, and the
// consumer simply loads the fragments it needs (L2, bypassing L1) until none of them is the sentinel -- ONE round trip when the data
// is there; a flag would cost two dependent ones (flag, then data), ~0.4 us each, in every stage.
constexpr unsigned long long BC_SENTINEL = 0x7FF8DEADBEEF0001ull;
__device__ __forceinline__ double bc_ld_l2(const double* p) {
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool bc_is_sentinel(double v) { return (unsigned long long)__double_as_longlong(v) == BC_SENTINEL; }

// ------------------------------------------------------------------------------------------------
// Z build.  Tile list: lower tiles of S (i >= j), then the W tiles (w, j).  256 threads per 64 x 64 tile.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    bc_build_kernel(const double* __restrict__ Sig, int ld, int dimp, const int* __restrict__ lmOf, const double* __restrict__ Cblk,
                    const double* __restrict__ ytilde, int nm, double r2, double* __restrict__ Z, int ldz, int nT, int TW,
                    int* __restrict__ trailCnt, double* __restrict__ LxAll, const int* __restrict__ guard, int tl) {
    pdl_wait();
    // completion counters of the urgent trailing launches; sentinels in the LT | XT blocks of every diagonal step (see BC_SENTINEL)
    if (blockIdx.x == 0 && threadIdx.x < 32) trailCnt[threadIdx.x] = 0;
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < (size_t)nT * BC_LX; w += (size_t)gridDim.x * blockDim.x)
        LxAll[w] = __longlong_as_double((long long)BC_SENTINEL);
    if (*guard) return;
    TL_MARK(tl, 0);
    __shared__ double sC[32][6];
    __shared__ int sIdx[32];
    __shared__ double sCr[32][6];
    __shared__ int sIdxR[32];
    const int tid = threadIdx.x;
    const int nS = nT * (nT + 1) / 2;
    const int m = 2 * nm;
    int bid = blockIdx.x;
    if (bid < nS) {
        int ti, tj;
        tri_decode(bid, ti, tj);
        // landmark pairs of the tile rows (32 of them) and tile columns
        if (tid < 32) {
            const int j = 32 * tj + tid;
            sIdx[tid] = j < nm ? SOFF + 3 * lmOf[j] : -1;
            for (int q = 0; q < 6; ++q) sC[tid][q] = j < nm ? Cblk[6 * (size_t)j + q] : 0.0;
        } else if (tid < 64) {
            const int t = tid - 32;
            const int j = 32 * ti + t;
            sIdxR[t] = j < nm ? SOFF + 3 * lmOf[j] : -1;
            for (int q = 0; q < 6; ++q) sCr[t][q] = j < nm ? Cblk[6 * (size_t)j + q] : 0.0;
        }
        __syncthreads();
        // thread: row landmark u = tid & 31, column landmarks v = (tid >> 5) + 8 q
        const int u = tid & 31;
        const int ir = sIdxR[u];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int v = (tid >> 5) + 8 * q;
            const int ic = sIdx[v];
            double o[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            if (ir >= 0 && ic >= 0) {
                double P[3][3];
#pragma unroll
                for (int b = 0; b < 3; ++b)
#pragma unroll
                    for (int a = 0; a < 3; ++a) P[a][b] = Sig[(size_t)(ic + b) * ld + ir + a];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double T[3];
#pragma unroll
                    for (int b = 0; b < 3; ++b) T[b] = sCr[u][3 * e] * P[0][b] + sCr[u][3 * e + 1] * P[1][b] + sCr[u][3 * e + 2] * P[2][b];
#pragma unroll
                    for (int f = 0; f < 2; ++f) o[e][f] = T[0] * sC[v][3 * f] + T[1] * sC[v][3 * f + 1] + T[2] * sC[v][3 * f + 2];
                }
            }
            const int R0 = 64 * ti + 2 * u, C0 = 64 * tj + 2 * v;
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int f = 0; f < 2; ++f) {
                    const int R = R0 + e, Cc = C0 + f;
                    double val = o[e][f];
                    if (R == Cc) val = (R < m) ? val + r2 : 1.0;  // identity padding of a short last block
                    Z[(size_t)Cc * ldz + R] = val;
                }
        }
    } else {
        bid -= nS;
        const int tw = bid / nT, tj = bid % nT;
        if (tid < 32) {
            const int j = 32 * tj + tid;
            sIdx[tid] = j < nm ? SOFF + 3 * lmOf[j] : -1;
            for (int q = 0; q < 6; ++q) sC[tid][q] = j < nm ? Cblk[6 * (size_t)j + q] : 0.0;
        }
        __syncthreads();
        const int sl = tid & 63;
        const int s = 64 * tw + sl;
        double* zrow = Z + (size_t)nT * BC_T + s;  // W rows start behind the nT S tiles
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int v = (tid >> 6) + 4 * q;
            const int ic = sIdx[v];
            double o0 = 0.0, o1 = 0.0;
            if (ic >= 0) {
                if (s == BC_YROW) {
                    o0 = ytilde[2 * (32 * tj + v)];
                    o1 = ytilde[2 * (32 * tj + v) + 1];
                } else if (s < dimp) {
                    const double s0 = Sig[(size_t)ic * ld + s], s1 = Sig[(size_t)(ic + 1) * ld + s], s2 = Sig[(size_t)(ic + 2) * ld + s];
                    o0 = sC[v][0] * s0 + sC[v][1] * s1 + sC[v][2] * s2;
                    o1 = sC[v][3] * s0 + sC[v][4] * s1 + sC[v][5] * s2;
                }
            }
            const int C0 = 64 * tj + 2 * v;
            zrow[(size_t)C0 * ldz] = o0;
            zrow[(size_t)(C0 + 1) * ldz] = o1;
        }
    }
    TL_MARK(tl, 1);
}

// ------------------------------------------------------------------------------------------------
// Diagonal step of the sweep (the critical path).
//
// The factorization follows chunk_factor_kernel (right-looking over 4x4 tiles of the lower triangle, unscaled columns, fraction-free
// elimination of the diagonal tile with its four reciprocals side by side), with two changes measured on this kernel (per-warp clock
// stamps, scripts/bc_timing.py): a lone warp issues one fp64 instruction every ~4 clocks, so what a block column costs is the LENGTH of
// each warp's instruction stream, not the pipe's throughput --
//   * every tile is split between two threads (rows 2h, 2h+1): 40 instead of 80 fp64 instructions per rank-4 update and warp, twice the
//     warps to interleave; published tiles are 16 contiguous doubles (+2 pad) so panels move as 16-byte shared-memory accesses;
//   * the next diagonal tile is eliminated right after its own update (look-ahead), not behind a barrier of its own.
// M_k = L_kk^-1 comes from 256 threads that carry the identity through the same elimination, one block column behind (mbarrier per
// block column), and do not take part in the S group's barriers.
// ------------------------------------------------------------------------------------------------
constexpr int BC_S_WARPS = 9;                      // 136 tiles x 2 half-tile owners = 272 threads
constexpr int BC_S_THREADS = BC_S_WARPS * 32;      // 288
constexpr int BC_X_THREADS = 128;                  // 4 diagonal 16 x 16 blocks x (4 x 4 tiles): their inverses ride along, one warp (16 lanes used) per
                                                   // block -- two blocks in one warp wait on different block columns, and the spinning half starves the other
constexpr int BC_DIAG_THREADS = 512;               // 16 warps for the DMMA prologue; warps 0-8 = S group, 9-12 = X group, 13 = publisher
constexpr int BC_PLD = 18;                         // doubles per published tile: [r][j] and 2 of padding (conflict-free 16-byte reads)
// What a diagonal step hands to its consumers (the next diagonal step, the panel kernel), per block k, two 64 x YB_LD arrays:
//   LT = L_kk transposed, the six 16 x 16 blocks below the diagonal blocks;
//   XT = X_b transposed, X_b = (b-th diagonal 16 x 16 block of L_kk)^-1, b = 0..3.
// P = T L_kk^-T is then a 4-stage block substitution on the FP64 tensor pipe:  P_b = (T_b - sum_{j<b} P_j L_bj^T) X_b^T.
// (An explicit 64 x 64 inverse riding along the whole factorization was measured first: its 256 threads made the loop fp64-pipe
// bound, ~1200 clocks per block column instead of ~650.)
__host__ __device__ __forceinline__ int bc_lt_block(int b, int j) { return b * (b - 1) / 2 + j; }

struct BcDiagSmem {
    double A[BC_T][YB_LD];            // T_{k,k-1} as A[j][r]; stage by stage replaced by P_{k,k-1} as A[c][r]
    double B[BC_T][YB_LD];            // the updated diagonal block as B[c][r]
    double Q[16][YB_LD];              // one stage's 64 x 16 block between its two products
    double Lp[CH_NT][CH_NT][BC_PLD];  // Lp[J][TI][4 r + j]: tile (TI, J) once block column J is finished (unscaled columns); TI = J: the diagonal tile
    double Dc[CH_NT][CH_T];           // reciprocal pivots of block column J
    double Inv[BC_T];
    uint64_t colBar[CH_NT];           // one mbarrier per block column: "its panels are published" (S group -> X group)
};
constexpr int BC_DIAG_SMEM = (int)sizeof(BcDiagSmem);

#ifdef EQVIO_CHUNK_TIMING
__device__ long long g_bc_t[16];
__device__ int g_bc_warp[BC_S_WARPS * 64];
// per-warp stamps go to shared memory (no global stores inside the loop) and are dumped after it
#define BC_FINE(i) do { if ((threadIdx.x & 31) == 0 && threadIdx.x < BC_S_THREADS) bc_stamps[threadIdx.x >> 5][(i)] = (int)clock(); } while (0)
#define BC_STAMP(i) do { if (threadIdx.x == 0 && kblk == 1) g_bc_t[(i)] = clock64(); } while (0)
__device__ unsigned long long g_bc_gt[32];  // globaltimer stamps across the first two diagonal steps (hand-over latency)
#define BC_GT(slot) do { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_bc_gt[(slot)] = t_; } while (0)
#else
#define BC_GT(slot) do { } while (0)
#define BC_FINE(i) do { } while (0)
#define BC_STAMP(i) do { } while (0)
#endif
__device__ __forceinline__ void bc_s_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(BC_S_THREADS) : "memory"); }
__device__ __forceinline__ int bc_diag_tile(int J) { return CH_NT * J - J * (J - 1) / 2; }  // column-major index of tile (J, J)

// Z: augmented matrix (ldz), kblk: block column.  LxAll: per block the LT | XT pair described above.
__global__ void __launch_bounds__(BC_DIAG_THREADS, 1)
    bc_diag_kernel(const double* __restrict__ Z, int ldz, int kblk, double* __restrict__ LxAll, int* __restrict__ status,
                   const int* trailCnt, int waitCnt, const int* __restrict__ guard, int tl) {
    extern __shared__ __align__(128) unsigned char bc_smem_raw[];
    BcDiagSmem& sm = *reinterpret_cast<BcDiagSmem*>(bc_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int k0 = kblk * BC_T;
    BC_STAMP(0);
    if (tid == 0) {
        for (int J = 0; J < CH_NT; ++J) mbar_init(&sm.colBar[J], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // Two dependencies, taken one at a time:
    //  (1) T_{k,k-1}, T_kk are current once trail(k-2) has finished.  That kernel runs on the other stream; an event edge into this
    //      chain costs ~3 us per step inside a replayed graph, so the trailing CTAs count themselves off in trailCnt[k-2] instead
    //      (release) and thread 0 polls it (acquire; bounded).  trail(k-2) never waits for this kernel, and everything it waits for
    //      completed before this grid could start (its stream predecessor, diag(k-1), was past its own dependency wait).
    //  (2) LT | XT of block k-1 come from diag(k-1), which is still RUNNING when this grid starts (programmatic launch: diag(k-1)
    //      releases its dependents once it has all of its own inputs).  It publishes them block row by block row -- block row b of
    //      L_{k-1,k-1} is final after block column 4b+3 of its factorization -- over sentinel-filled blocks (BC_SENTINEL), so three of
    //      the four substitution stages below, and three quarters of D = T_kk - P P^T, run beside the previous factorization.
    if (waitCnt > 0) {
        if (tid == 0) {
            int spins = 0, seen;
            do {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(trailCnt + (kblk - 2)) : "memory");
            } while (seen < waitCnt && ++spins < (1 << 24));
            if (seen < waitCnt) atomicOr(status, 8);
        }
        __syncthreads();
    }
    double2 tpre[4];
    double acc2[3][2];
    if (kblk > 0) {
        const double* Tp = Z + (size_t)(k0 - BC_T) * ldz + k0;  // T_{k,k-1}: 2048 16-byte words, four per thread
        const double* Td = Z + (size_t)k0 * ldz + k0;           // T_{k,k}: the 36 lower 8x8 fragments go round-robin over the 16 warps
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int q = tid + BC_DIAG_THREADS * u;
            tpre[u] = *reinterpret_cast<const double2*>(Tp + (size_t)(q >> 5) * ldz + (q & 31) * 2);
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int f = warp + 16 * u;
            if (f < 36) {
                int fi, fj;
                tri_decode(f, fi, fj);
                acc2[u][0] = -Td[(size_t)(8 * fj + 2 * t4) * ldz + 8 * fi + g];
                acc2[u][1] = -Td[(size_t)(8 * fj + 2 * t4 + 1) * ldz + 8 * fi + g];
            }
        }
    }
    if (kblk == 0) {
        asm volatile("griddepcontrol.wait;" ::: "memory");  // the build kernel (and with it everything earlier on the stream)
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    if (*guard) {
        // a guarded update: nothing will be published, the chain must still unwind
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        return;
    }
    TL_MARK(tl, 0);
    BC_STAMP(1);
    if (tid == 0 && kblk < 2) BC_GT(8 * kblk);
    const bool sGroup = tid < BC_S_THREADS;
    const int t = tid >> 1, h = tid & 1;  // S group: tile t (column-major over the lower triangle), rows 2h, 2h+1 of it
    const bool owner = sGroup && t < CH_TILES;
    int TI = 0, TK = 0;
    if (owner) tri_decode_cm(t, CH_NT, TI, TK);
    double a[2][CH_T];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < CH_T; ++c) a[r][c] = 0.0;

    if (kblk > 0) {
        // ---- look-ahead products on the FP64 tensor pipe:  P = T_{k,k-1} L^-T (block substitution, one stage per published block
        //      row),  D = T_kk - P P^T accumulated stage by stage  ----
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int q = tid + BC_DIAG_THREADS * u;
            *reinterpret_cast<double2*>(&sm.A[q >> 5][(q & 31) * 2]) = tpre[u];
        }
        int fi2[3], fj2[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            fi2[u] = fj2[u] = 0;
            if (warp + 16 * u < 36) tri_decode(warp + 16 * u, fi2[u], fj2[u]);
        }
        const int nf = warp < 4 ? 3 : 2;
        const double* Lxg = LxAll + (size_t)(kblk - 1) * BC_LX;
        __syncthreads();  // T in shared memory
        BC_STAMP(2);
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
            // warp fi < 8 owns rows 8 fi .. 8 fi + 7 of P through all stages (both 8-column fragments of a stage).  Its operands from
            // diag(k-1) -- fragments of LT block row b, then of XT block b -- come straight from L2 into registers: no staging, no
            // CTA-wide barrier between their arrival and the products.
            if (warp < 8) {
                const int fi = warp;
                double q[2][2];
#pragma unroll
                for (int fn = 0; fn < 2; ++fn) {
                    q[fn][0] = -sm.A[16 * b + 8 * fn + 2 * t4][8 * fi + g];
                    q[fn][1] = -sm.A[16 * b + 8 * fn + 2 * t4 + 1][8 * fi + g];
                }
                const double* Xb = Lxg + (size_t)(6 + b) * BC_XBLK;
                double xf[4][2];
                auto load_x = [&]() {
                    bool bad = false;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                        for (int fn = 0; fn < 2; ++fn) {
                            xf[kk][fn] = bc_ld_l2(Xb + (4 * kk + t4) * BC_XLD + 8 * fn + g);
                            bad |= bc_is_sentinel(xf[kk][fn]);
                        }
                    return bad;
                };
                bool xbad = true;
                if (b > 0) {
                    double lf[3][4][2];  // every fragment of the block row (and, optimistically, of XT block b) in flight at once
                    int spins = 0;
                    bool bad;
                    do {
                        bad = false;
#pragma unroll
                        for (int jb = 0; jb < 3; ++jb)
                            if (jb < b) {
                                const double* Lb = Lxg + (size_t)bc_lt_block(b, jb) * BC_XBLK;
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                                    for (int fn = 0; fn < 2; ++fn) {
                                        lf[jb][kk][fn] = bc_ld_l2(Lb + (4 * kk + t4) * BC_XLD + 8 * fn + g);
                                        bad |= bc_is_sentinel(lf[jb][kk][fn]);
                                    }
                            }
                        if (spins == 0) xbad = load_x();
                        bad = __any_sync(0xffffffffu, bad);
                    } while (bad && ++spins < (1 << 22));
                    if (bad) atomicOr(status, 8);
#pragma unroll
                    for (int jb = 0; jb < 3; ++jb)
                        if (jb < b) {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const double af = sm.A[16 * jb + 4 * kk + t4][8 * fi + g];
#pragma unroll
                                for (int fn = 0; fn < 2; ++fn) dmma884(q[fn][0], q[fn][1], af, lf[jb][kk][fn]);
                            }
                        }
                }
#pragma unroll
                for (int fn = 0; fn < 2; ++fn) {
                    sm.Q[8 * fn + 2 * t4][8 * fi + g] = -q[fn][0];
                    sm.Q[8 * fn + 2 * t4 + 1][8 * fi + g] = -q[fn][1];
                }
                {
                    int spins = 0;
                    xbad = __any_sync(0xffffffffu, xbad);
                    while (xbad && ++spins < (1 << 22)) xbad = __any_sync(0xffffffffu, load_x());
                    if (xbad) atomicOr(status, 8);
                }
                if (tid == 0 && kblk == 1) BC_GT(16 + b);
                __syncwarp();
                double p[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const double af = sm.Q[4 * kk + t4][8 * fi + g];
                    if (kk < 2) dmma884(p[0][0], p[0][1], af, xf[kk][0]);
                    dmma884(p[1][0], p[1][1], af, xf[kk][1]);
                }
#pragma unroll
                for (int fn = 0; fn < 2; ++fn) {  // P_b takes the place of T's block column b
                    sm.A[16 * b + 8 * fn + 2 * t4][8 * fi + g] = p[fn][0];
                    sm.A[16 * b + 8 * fn + 2 * t4 + 1][8 * fi + g] = p[fn][1];
                }
            }
            __syncthreads();
            if (b == 3) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // diag(k-1) is done: let diag(k+1) in
            // D -= P_b P_b^T on the 36 lower 8 x 8 fragments (round-robin over the 16 warps: 9 per SM sub-partition)
#pragma unroll
            for (int c4 = 0; c4 < 16; c4 += 4) {
#pragma unroll
                for (int u = 0; u < 3; ++u)
                    if (u < nf) dmma884(acc2[u][0], acc2[u][1], sm.A[16 * b + c4 + t4][8 * fi2[u] + g], sm.A[16 * b + c4 + t4][8 * fj2[u] + g]);
            }
        }
        BC_STAMP(3);
#pragma unroll
        for (int u = 0; u < 3; ++u)
            if (u < nf) {
                sm.B[8 * fj2[u] + 2 * t4][8 * fi2[u] + g] = -acc2[u][0];
                sm.B[8 * fj2[u] + 2 * t4 + 1][8 * fi2[u] + g] = -acc2[u][1];
            }
        __syncthreads();
        BC_STAMP(4);
        if (owner) {
#pragma unroll
            for (int c = 0; c < CH_T; ++c)
#pragma unroll
                for (int r = 0; r < 2; ++r) a[r][c] = sm.B[CH_T * TK + c][CH_T * TI + 2 * h + r];
        }
    } else {
        BC_STAMP(2);
        BC_STAMP(3);
        if (owner) {
            const double* Td = Z + (size_t)k0 * ldz + k0;
#pragma unroll
            for (int c = 0; c < CH_T; ++c) {
                const double2 v = *reinterpret_cast<const double2*>(Td + (size_t)(CH_T * TK + c) * ldz + CH_T * TI + 2 * h);
                a[0][c] = v.x;
                a[1][c] = v.y;
            }
        }
        __syncthreads();  // mbarriers initialised
        BC_STAMP(4);
    }

#ifdef EQVIO_CHUNK_TIMING
    __shared__ int bc_stamps[BC_S_WARPS][64];
#endif
    if (sGroup) {
        // ================================ S group: the factorization ================================
        double pv[CH_T] = {1.0, 1.0, 1.0, 1.0};  // pivots of the diagonal tile eliminated by this thread
        // Diagonal tile (J, J): called by the whole warp that holds it; the h = 0 thread fetches rows 2, 3 from its neighbour and
        // eliminates (fraction-free: products only, then the four reciprocals side by side), publishes the tile and 1 / pivots.
        auto eliminate = [&](int J) {
            const double a20 = __shfl_down_sync(0xffffffffu, a[0][0], 1), g21 = __shfl_down_sync(0xffffffffu, a[0][1], 1),
                         g22 = __shfl_down_sync(0xffffffffu, a[0][2], 1);
            const double a30 = __shfl_down_sync(0xffffffffu, a[1][0], 1), g31 = __shfl_down_sync(0xffffffffu, a[1][1], 1),
                         g32 = __shfl_down_sync(0xffffffffu, a[1][2], 1), g33 = __shfl_down_sync(0xffffffffu, a[1][3], 1);
            if (tid == 2 * bc_diag_tile(J)) {
                const double a00 = a[0][0], a10 = a[1][0];
                const double m11 = a[1][1] * a00 - a10 * a10, m21 = g21 * a00 - a20 * a10, m22 = g22 * a00 - a20 * a20;
                const double m31 = g31 * a00 - a30 * a10, m32 = g32 * a00 - a30 * a20, m33 = g33 * a00 - a30 * a30;
                const double n22 = m22 * m11 - m21 * m21, n32 = m32 * m11 - m31 * m21, n33 = m33 * m11 - m31 * m31;
                const double p33 = n33 * n22 - n32 * n32;
                const double r0 = fast_rcp(a00), r1 = fast_rcp(m11), r2_ = fast_rcp(n22), r3 = fast_rcp(p33);
                const double s2 = r0 * r1, s3 = s2 * r2_, e1 = a00 * m11;
                pv[0] = a00;
                pv[1] = m11 * r0;
                pv[2] = n22 * s2;
                pv[3] = p33 * s3;
                double2* dc = reinterpret_cast<double2*>(&sm.Dc[J][0]);
                dc[0] = make_double2(r0, a00 * r1);
                dc[1] = make_double2(e1 * r2_, (e1 * n22) * r3);
                // unscaled columns below the diagonal (what the panel owners need): d[k][j], j < k
                double2* dt = reinterpret_cast<double2*>(&sm.Lp[J][J][0]);
                dt[2] = make_double2(a10, 0.0);
                dt[4] = make_double2(a20, m21 * r0);
                dt[6] = make_double2(a30, m31 * r0);
                dt[7] = make_double2(n32 * s2, 0.0);
            }
        };
        const int myWarpFirstTile = (warp * 32) >> 1, myWarpLastTile = myWarpFirstTile + 15;
        if (tid == 0 && kblk < 2) BC_GT(8 * kblk + 1);
        if (warp == 0) eliminate(0);
        for (int J = 0; J < CH_NT; ++J) {
            BC_FINE(4 * J);
            BC_FINE(4 * J + 1);
            bc_s_barrier();  // diagonal tile J published
            BC_FINE(4 * J + 2);
            if (owner && TK == J && TI > J) {
                const double2 c01 = *reinterpret_cast<const double2*>(&sm.Dc[J][0]);
                const double c2 = sm.Dc[J][2];
                const double2* dt = reinterpret_cast<const double2*>(&sm.Lp[J][J][0]);
                const double d10 = dt[2].x;
                const double2 d2 = dt[4], d3 = dt[6];
                const double d32 = dt[7].x;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const double t0 = a[r][0] * c01.x;
                    a[r][1] -= t0 * d10;
                    a[r][2] -= t0 * d2.x;
                    a[r][3] -= t0 * d3.x;
                    const double t1 = a[r][1] * c01.y;
                    a[r][2] -= t1 * d2.y;
                    a[r][3] -= t1 * d3.y;
                    const double t2 = a[r][2] * c2;
                    a[r][3] -= t2 * d32;
                }
                double2* out = reinterpret_cast<double2*>(&sm.Lp[J][TI][8 * h]);
                out[0] = make_double2(a[0][0], a[0][1]);
                out[1] = make_double2(a[0][2], a[0][3]);
                out[2] = make_double2(a[1][0], a[1][1]);
                out[3] = make_double2(a[1][2], a[1][3]);
            }
            bc_s_barrier();  // panels of block column J published
            BC_FINE(4 * J + 3);
            // the S group's barrier above ordered every panel store before this arrive (release, cumulative); a release STORE to a
            // flag word would cost a MEMBAR.ALL.CTA in warp 0 on every block column, the mbarrier arrive does not
            if (tid == 0) mbar_arrive(&sm.colBar[J]);
            if (owner && TK > J) {
                const double2* cp = reinterpret_cast<const double2*>(&sm.Dc[J][0]);
                const double2 c01 = cp[0], c23 = cp[1];
                const double2* lp = reinterpret_cast<const double2*>(&sm.Lp[J][TI][8 * h]);
                const double2* pp = reinterpret_cast<const double2*>(&sm.Lp[J][TK][0]);
                double li[2][CH_T];
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const double2 x = lp[2 * r], y = lp[2 * r + 1];
                    li[r][0] = x.x * c01.x;
                    li[r][1] = x.y * c01.y;
                    li[r][2] = y.x * c23.x;
                    li[r][3] = y.y * c23.y;
                }
                // j outermost: eight independent accumulators per level, the four levels pipeline (cc outermost leaves two
                // dependent chains of four per load pair and ~2x the latency)
                double pk[CH_T][CH_T];
#pragma unroll
                for (int cc = 0; cc < CH_T; ++cc) {
                    const double2 x = pp[2 * cc], y = pp[2 * cc + 1];
                    pk[cc][0] = x.x;
                    pk[cc][1] = x.y;
                    pk[cc][2] = y.x;
                    pk[cc][3] = y.y;
                }
#pragma unroll
                for (int j = 0; j < CH_T; ++j)
#pragma unroll
                    for (int cc = 0; cc < CH_T; ++cc)
#pragma unroll
                        for (int r = 0; r < 2; ++r) a[r][cc] = fma(-li[r][j], pk[cc][j], a[r][cc]);
            }
            // look-ahead: the next diagonal tile is eliminated as soon as its own update is in, while the other warps are still in
            // theirs -- the next iteration goes straight to its first barrier
            if (J + 1 < CH_NT) {
                const int dn = bc_diag_tile(J + 1);
                if (dn >= myWarpFirstTile && dn <= myWarpLastTile) eliminate(J + 1);
            }
        }
        if (owner && TI == TK && h == 0) {
#pragma unroll
            for (int c = 0; c < CH_T; ++c)
                if (!(pv[c] > 0.0)) atomicOr(status, 1);  // S is not positive definite
        }
        BC_STAMP(5);
        if (tid == 0 && kblk < 2) BC_GT(8 * kblk + 2);
    } else if (tid < BC_S_THREADS + BC_X_THREADS) {
        // ======== X group: the identity rides along inside each diagonal 16 x 16 block (4 steps per block): X_b = L_bb^-1 ========
        const int q = lane & 15;
        const int blk = warp - BC_S_WARPS, i = (q >> 2) & 3, tk = q & 3;
        if (lane >= 16) goto done;  // (the upper half of the warp only takes part in the final barrier)
        TI = 4 * blk + i;   // tile row: rows 4 TI .. 4 TI + 3 of the identity
        TK = 4 * blk + tk;  // tile column
        const unsigned grp = 0xffffu;
        double b[CH_T][CH_T];
#pragma unroll
        for (int r = 0; r < CH_T; ++r)
#pragma unroll
            for (int c = 0; c < CH_T; ++c) b[r][c] = (i == tk && r == c) ? 1.0 : 0.0;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const int J = 4 * blk + j;
            if (!mbar_wait_bounded(&sm.colBar[J], 0)) atomicOr(status, 8);
            const double2* cp = reinterpret_cast<const double2*>(&sm.Dc[J][0]);
            const double2 c01 = cp[0], c23 = cp[1];
            const double c[CH_T] = {c01.x, c01.y, c23.x, c23.y};
            if (tk == j) {
                const double2* dt = reinterpret_cast<const double2*>(&sm.Lp[J][J][0]);
                const double d10 = dt[2].x;
                const double2 d2 = dt[4], d3 = dt[6];
                const double d32 = dt[7].x;
#pragma unroll
                for (int r = 0; r < CH_T; ++r) {
                    const double t0 = b[r][0] * c[0];
                    b[r][1] -= t0 * d10;
                    b[r][2] -= t0 * d2.x;
                    b[r][3] -= t0 * d3.x;
                    const double t1 = b[r][1] * c[1];
                    b[r][2] -= t1 * d2.y;
                    b[r][3] -= t1 * d3.y;
                    const double t2 = b[r][2] * c[2];
                    b[r][3] -= t2 * d32;
                }
            }
            double li[CH_T][CH_T];
#pragma unroll
            for (int r = 0; r < CH_T; ++r)
#pragma unroll
                for (int jj = 0; jj < CH_T; ++jj) li[r][jj] = __shfl_sync(grp, b[r][jj], j, 4) * c[jj];
            if (tk > j) {
                const double2* pp = reinterpret_cast<const double2*>(&sm.Lp[J][TK][0]);
#pragma unroll
                for (int cc = 0; cc < CH_T; ++cc) {
                    const double2 x = pp[2 * cc], y = pp[2 * cc + 1];
#pragma unroll
                    for (int r = 0; r < CH_T; ++r) {
                        double acc = b[r][cc];
                        acc -= li[r][0] * x.x;
                        acc -= li[r][1] * x.y;
                        acc -= li[r][2] * y.x;
                        acc -= li[r][3] * y.y;
                        b[r][cc] = acc;
                    }
                }
            }
        }
        // 1 / L_jj = sqrt of the reciprocal pivot the diagonal thread published
        double* Xt = LxAll + (size_t)kblk * BC_LX + (size_t)(6 + blk) * BC_XBLK;
        const double2* cq = reinterpret_cast<const double2*>(&sm.Dc[TK][0]);
        const double2 e01 = cq[0], e23 = cq[1];
        const double s0 = sqrt(e01.x), s1 = sqrt(e01.y), s2 = sqrt(e23.x), s3 = sqrt(e23.y);
#pragma unroll
        for (int r = 0; r < CH_T; ++r) {
            double* p = Xt + (CH_T * i + r) * BC_XLD + CH_T * tk;
            *reinterpret_cast<double2*>(p) = make_double2(b[r][0] * s0, b[r][1] * s1);
            *reinterpret_cast<double2*>(p + 2) = make_double2(b[r][2] * s2, b[r][3] * s3);
        }
        if (q == 0 && kblk == 0) BC_GT(3 + blk);
    } else if (warp == (BC_S_THREADS + BC_X_THREADS) / 32) {
        // ======== publisher (one of the warps that are idle during the loop): block row b of L_kk only has entries in block columns
        // < b, final once block column 4b - 1 is finished: scaled, transposed, to global memory, then its flag ========
        double* Ltg = LxAll + (size_t)kblk * BC_LX;
#pragma unroll 1
        for (int b = 1; b < 4; ++b) {
            if (!mbar_wait_bounded(&sm.colBar[4 * b - 1], 0)) atomicOr(status, 8);
            if (lane < 16) sm.Inv[16 * (b - 1) + lane] = sqrt(sm.Dc[4 * (b - 1) + (lane >> 2)][lane & 3]);  // 1 / L_jj of block b-1's columns
            __syncwarp();
            for (int idx = lane; idx < 256 * b; idx += 32) {
                const int n = idx & 15, k = idx >> 4;  // row 16 b + n, column k of L
                const double v = sm.Lp[k >> 2][4 * b + (n >> 2)][4 * (n & 3) + (k & 3)] * sm.Inv[k];
                Ltg[(size_t)bc_lt_block(b, k >> 4) * BC_XBLK + (k & 15) * BC_XLD + n] = v;
            }
        }
    }
done:
    __syncthreads();

#ifdef EQVIO_CHUNK_TIMING
    if (kblk == 1 && tid < BC_S_WARPS * 32) {
        g_bc_warp[2 * tid] = bc_stamps[(2 * tid) / 64][(2 * tid) % 64];
        g_bc_warp[2 * tid + 1] = bc_stamps[(2 * tid + 1) / 64][(2 * tid + 1) % 64];
    }
#endif
    BC_STAMP(6);
    TL_MARK(tl, 1);
}

// ------------------------------------------------------------------------------------------------
// Panel step: P = Z[rows of tile t, blk k] L_kk^-T for every row tile t below the diagonal block (S tiles k+1 .. nT-1, then the
// W tiles), two CTAs per tile (32 rows each), by the 4-stage block substitution described at BcDiagSmem (LT / XT of block k).
// Output in the tile-blocked panel layout of the downdate kernel: Zp[t][c][r].
// ------------------------------------------------------------------------------------------------
constexpr int BC_PA_LD = 36;  // 32 rows of a half tile (+4: conflict-free fragment reads)
struct BcPanelSmem {
    double LX[10][16][BC_XLD];  // LT | XT of block k
    double A[BC_T][BC_PA_LD];   // T as A[j][r]; stage by stage replaced by P as A[c][r]
    double Q[16][BC_PA_LD];
    uint64_t bar;
};
constexpr int BC_PANEL_SMEM = (int)sizeof(BcPanelSmem);

__global__ void __launch_bounds__(128)
    bc_panel_kernel(const double* __restrict__ Z, int ldz, int kblk, const double* __restrict__ LxAll, double* __restrict__ Zp,
                    const int* __restrict__ guard, int tl) {
    pdl_wait();
    if (*guard) return;
    TL_MARK(tl, 0);
    extern __shared__ __align__(128) unsigned char bcp_smem_raw[];
    BcPanelSmem& sm = *reinterpret_cast<BcPanelSmem*>(bcp_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int t = kblk + 1 + (int)(blockIdx.x >> 1), half = blockIdx.x & 1;
    if (tid == 0) {
        mbar_init(&sm.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&sm.bar, (uint32_t)(BC_LX * 8));
        bulk_g2s(&sm.LX[0][0][0], LxAll + (size_t)kblk * BC_LX, (uint32_t)(BC_LX * 8), &sm.bar);
    }
    const double* Tp = Z + (size_t)(kblk * BC_T) * ldz + (size_t)t * BC_T + 32 * half;
    {
        double2 v[8];  // 1024 16-byte words, eight per thread, all in flight before the first store
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int q = tid + 128 * u;
            v[u] = *reinterpret_cast<const double2*>(Tp + (size_t)(q >> 4) * ldz + (q & 15) * 2);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int q = tid + 128 * u;
            *reinterpret_cast<double2*>(&sm.A[q >> 4][(q & 15) * 2]) = v[u];
        }
    }
    __syncthreads();
    mbar_wait(&sm.bar, 0);
    // 32 rows x 16 columns per stage: row fragment = warp, both column fragments
    const int fi = warp;
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
        double q[2][2];
#pragma unroll
        for (int fn = 0; fn < 2; ++fn) {
            q[fn][0] = -sm.A[16 * b + 8 * fn + 2 * t4][8 * fi + g];
            q[fn][1] = -sm.A[16 * b + 8 * fn + 2 * t4 + 1][8 * fi + g];
        }
        for (int k4 = 0; k4 < 16 * b; k4 += 4) {
            const double af = sm.A[k4 + t4][8 * fi + g];
#pragma unroll
            for (int fn = 0; fn < 2; ++fn) dmma884(q[fn][0], q[fn][1], af, sm.LX[bc_lt_block(b, k4 >> 4)][(k4 & 15) + t4][8 * fn + g]);
        }
#pragma unroll
        for (int fn = 0; fn < 2; ++fn) {
            sm.Q[8 * fn + 2 * t4][8 * fi + g] = -q[fn][0];
            sm.Q[8 * fn + 2 * t4 + 1][8 * fi + g] = -q[fn][1];
        }
        __syncwarp();  // a warp only reads back its own 8 rows of Q
        double p[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int k4 = 0; k4 < 16; k4 += 4) {
            const double af = sm.Q[k4 + t4][8 * fi + g];
            if (k4 < 8) dmma884(p[0][0], p[0][1], af, sm.LX[6 + b][k4 + t4][g]);
            dmma884(p[1][0], p[1][1], af, sm.LX[6 + b][k4 + t4][8 + g]);
        }
#pragma unroll
        for (int fn = 0; fn < 2; ++fn) {
            sm.A[16 * b + 8 * fn + 2 * t4][8 * fi + g] = p[fn][0];
            sm.A[16 * b + 8 * fn + 2 * t4 + 1][8 * fi + g] = p[fn][1];
        }
        __syncwarp();  // ... and its own 8 rows of A
    }
    __syncthreads();
    double* out = Zp + (size_t)t * YB_TILE + 32 * half;
    for (int q = tid; q < BC_T * 16; q += 128) {
        const int c = q >> 4, r = (q & 15) * 2;
        *reinterpret_cast<double2*>(out + (size_t)c * YB_LD + r) = make_double2(sm.A[c][r], sm.A[c][r + 1]);
    }
    TL_MARK(tl, 1);
}

// ------------------------------------------------------------------------------------------------
// Trailing step: one 64 x 64 tile per CTA pair (32 rows each), C -= P_a P_b^T with the panels from Zp.
//   S tiles  (i >= j > k, except (k+1, k+1): the next diagonal step applies that one itself)  in Z
//   W tiles  (w, j > k)                                                                         in Z
//   Sigma tiles (a >= b), mirrored on the last step; column BC_YROW of the product is -dGamma, row / column BC_YROW of Sigma stay zero
// ------------------------------------------------------------------------------------------------
// Which trailing tiles of step k a launch covers (local tile coordinates (li, lj) relative to block k+1, q = nT - k - 1):
//   URGENT: (1,0), (1,1) = T(k+2,k+1), T(k+2,k+2), what diag(k+2) reads            -> bc_next_kernel<4> (16-row CTAs, its own stream)
//   NEXT  : the rest of block column k+1 (S tiles (li >= 2, 0), W tiles (w, 0)) and (2,1), (2,2), the urgent pair of the NEXT step
//           (so that urgent(k+1) only waits for this launch, not for the long one below)            -> bc_next_kernel<2>
//   REST  : everything else, and the Sigma tiles                                                       -> bc_trail_kernel, after panel(k)
//   ALL   : one launch (per-kernel profile runs)
enum { BC_PART_ALL = 0, BC_PART_NEXT = 1, BC_PART_REST = 2, BC_PART_URGENT = 3 };
constexpr int BC_TRAIL_URGENT = 8;  // CTAs of the urgent launch (2 tiles x 4 row quarters)
__host__ __device__ __forceinline__ int bc_trail_tiles(int part, int q, int TW) {
    const int all = (q > 0 ? q * (q + 1) / 2 - 1 : 0) + TW * q + TW * (TW + 1) / 2;
    const int urgent = q >= 2 ? 2 : 0;
    const int next = (q >= 3 ? (q - 2) + 2 : 0) + (q >= 1 ? TW : 0);
    return part == BC_PART_ALL ? all : part == BC_PART_URGENT ? urgent : part == BC_PART_NEXT ? next : all - urgent - next;
}
__global__ void __launch_bounds__(DD_THREADS, 3)
    bc_trail_kernel(double* Z, int ldz, double* Sig, int ld, const double* __restrict__ Zp, double* Gamma, const int* __restrict__ guard,
                    int kblk, int nT, int TW, int dimp, int mirrorAll, int part, int* __restrict__ trailCnt, int tl) {
    pdl_wait();
    const int bid = (int)(blockIdx.x >> 1), half = (int)(blockIdx.x & 1);
    const int q = nT - kblk - 1;  // block columns to the right of this one
    // the two tiles diag(k+2) reads, T(k+2, k+1) and T(k+2, k+2): their four CTAs count themselves off (BC_TRAIL_URGENT of them)
    const bool urgent = part == BC_PART_ALL && q >= 2 && bid < 2;
    if (*guard) {
        if (urgent && threadIdx.x == 0) atomicAdd(trailCnt + kblk, 1);  // diag(k+2) counts them whatever they did
        return;
    }
    TL_MARK(tl, 0);
    int ta, tb;  // panel tiles in Zp
    double* C;
    int ldc;
    bool isSig = false;
    // local tile (li, lj) of the trailing S block, li >= lj, (0, 0) excluded; W tile (w, lj); Sigma tile (sa, sb)
    int kind, li = 0, lj = 0;  // 0: S, 1: W, 2: Sigma
    if (part == BC_PART_ALL) {
        const int nS = q > 0 ? q * (q + 1) / 2 - 1 : 0;
        if (bid < nS) {
            kind = 0;
            tri_decode(bid + 1, li, lj);
        } else if (bid < nS + TW * q) {
            kind = 1;
            li = (bid - nS) / q;
            lj = (bid - nS) % q;
        } else {
            kind = 2;
            tri_decode(bid - nS - TW * q, li, lj);
        }
    } else {
        const int skipB = q >= 3 ? 3 : (q >= 2 ? 1 : 0);  // (1,1), (2,1), (2,2) belong to the urgent / next launches
        const int nSB = q >= 2 ? (q - 1) * q / 2 - skipB : 0;
        const int nWB = q >= 2 ? TW * (q - 1) : 0;
        if (bid < nSB) {
            kind = 0;
            tri_decode(bid + skipB, li, lj);
            ++li;
            ++lj;
        } else if (bid < nSB + nWB) {
            kind = 1;
            li = (bid - nSB) / (q - 1);
            lj = 1 + (bid - nSB) % (q - 1);
        } else {
            kind = 2;
            tri_decode(bid - nSB - nWB, li, lj);
        }
    }
    if (kind == 0) {
        ta = kblk + 1 + li;
        tb = kblk + 1 + lj;
        C = Z + (size_t)tb * BC_T * ldz + (size_t)ta * BC_T;
        ldc = ldz;
    } else if (kind == 1) {
        ta = nT + li;
        tb = kblk + 1 + lj;
        C = Z + (size_t)tb * BC_T * ldz + (size_t)ta * BC_T;
        ldc = ldz;
    } else {
        ta = nT + li;
        tb = nT + lj;
        C = Sig + (size_t)lj * BC_T * ld + (size_t)li * BC_T;
        ldc = ld;
        isSig = true;
    }
    extern __shared__ __align__(16) unsigned char bct_smem_raw[];
    double(*sA)[DD_LD] = reinterpret_cast<double(*)[DD_LD]>(bct_smem_raw);
    double(*sB)[DD_LD] = sA + DD_T;
    uint64_t* bar = reinterpret_cast<uint64_t*>(bct_smem_raw + 2 * DD_T * DD_LD * 8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool diag = ta == tb;
    constexpr uint32_t PANEL_BYTES = DD_T * DD_LD * 8;
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar, diag ? PANEL_BYTES : 2 * PANEL_BYTES);
        bulk_g2s(&sA[0][0], Zp + (size_t)ta * YB_TILE, PANEL_BYTES, bar);
        if (!diag) bulk_g2s(&sB[0][0], Zp + (size_t)tb * YB_TILE, PANEL_BYTES, bar);
    }
    const int wm = half * 32 + (warp >> 1) * 16, wn = (warp & 1) * 32;
    const int fr = wm + (lane >> 2), fc = wn + (lane & 3) * 2;
    double acc[2][4][2];
    double* cbase = C + (size_t)fc * ldc + fr;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            acc[a][b][0] = -cbase[(size_t)(b * 8) * ldc + a * 8];
            acc[a][b][1] = -cbase[(size_t)(b * 8 + 1) * ldc + a * 8];
        }
    __syncthreads();
    mbar_wait(bar, 0);
    double(*sBB)[DD_LD] = diag ? sA : sB;
#pragma unroll 4
    for (int k4 = 0; k4 < DD_T; k4 += 4) {
        double af[2], bf[4];
#pragma unroll
        for (int a = 0; a < 2; ++a) af[a] = sA[k4 + (lane & 3)][wm + a * 8 + (lane >> 2)];
#pragma unroll
        for (int b = 0; b < 4; ++b) bf[b] = sBB[k4 + (lane & 3)][wn + b * 8 + (lane >> 2)];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
    if (!isSig) {
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                cbase[(size_t)(b * 8) * ldc + a * 8] = -acc[a][b][0];
                cbase[(size_t)(b * 8 + 1) * ldc + a * 8] = -acc[a][b][1];
            }
    } else {
        const int i0 = (ta - nT) * BC_T, j0 = (tb - nT) * BC_T;
        const bool mirror = !diag && mirrorAll;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int R = i0 + fr + a * 8, Cc = j0 + fc + b * 8;
                const double v0 = -acc[a][b][0], v1 = -acc[a][b][1];
                if (R == BC_YROW) continue;  // the ytilde row of W: not a state row
                if (Cc == BC_YROW - 1) {     // columns 20 | 21: the second one is -dGamma
                    cbase[(size_t)(b * 8) * ldc + a * 8] = v0;
                    if (R < dimp) Gamma[R] -= v1;
                    if (mirror) Sig[(size_t)R * ld + Cc] = v0;
                    continue;
                }
                cbase[(size_t)(b * 8) * ldc + a * 8] = v0;
                cbase[(size_t)(b * 8 + 1) * ldc + a * 8] = v1;
                if (mirror) {
                    double2 tt;
                    tt.x = v0;
                    tt.y = v1;
                    *reinterpret_cast<double2*>(Sig + (size_t)R * ld + Cc) = tt;
                }
            }
    }
    if (urgent) {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(trailCnt + kblk, 1);
        }
    }
    TL_MARK(tl, 1);
}

// ------------------------------------------------------------------------------------------------
// What the NEXT block column needs from step k, in one launch: for every tile of block column k+1 (S tiles (i, k+1), i >= k+2, the W
// tiles (w, k+1)) and the diagonal tile (k+2, k+2), a CTA pair substitutes the two row panels it needs itself (P_a: its 32 rows, P_b:
// the 64 rows of tile k+1 or k+2) and applies  T -= P_a P_b^T  -- panel(k) -> trail<next>(k) as two launches put ~11 us between the end
// of diag(k) and the two tiles diag(k+2) reads; this puts ~5 us there, and panel(k) / trail<rest>(k) run beside it.
// ------------------------------------------------------------------------------------------------
constexpr int BC_NEXT_THREADS = 384;  // 12 warps: 8 row fragments of P_b, 4 of P_a
constexpr int BC_QLD = 12;
struct BcNextSmem {
    double LX[10][16][BC_XLD];
    double Pb[BC_T][YB_LD];
    double Pa[BC_T][BC_PA_LD];
    double Q[12][16][BC_QLD];
    uint64_t bar;
};
constexpr int BC_NEXT_SMEM = (int)sizeof(BcNextSmem);

// four substitution stages on the 8 rows r0 .. r0 + 7 of a panel held as P[c][r] (T on entry, P = T L^-T on exit); warp-local
template <int LD>
__device__ __forceinline__ void bc_substitute_rows(double (*P)[LD], int r0, const double (*LX)[16][BC_XLD], double (*Q)[BC_QLD], int g, int t4) {
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
        double q[2][2];
#pragma unroll
        for (int fn = 0; fn < 2; ++fn) {
            q[fn][0] = -P[16 * b + 8 * fn + 2 * t4][r0 + g];
            q[fn][1] = -P[16 * b + 8 * fn + 2 * t4 + 1][r0 + g];
        }
        for (int k4 = 0; k4 < 16 * b; k4 += 4) {
            const double af = P[k4 + t4][r0 + g];
#pragma unroll
            for (int fn = 0; fn < 2; ++fn) dmma884(q[fn][0], q[fn][1], af, LX[bc_lt_block(b, k4 >> 4)][(k4 & 15) + t4][8 * fn + g]);
        }
#pragma unroll
        for (int fn = 0; fn < 2; ++fn) {
            Q[8 * fn + 2 * t4][g] = -q[fn][0];
            Q[8 * fn + 2 * t4 + 1][g] = -q[fn][1];
        }
        __syncwarp();
        double p[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int k4 = 0; k4 < 16; k4 += 4) {
            const double af = Q[k4 + t4][g];
            if (k4 < 8) dmma884(p[0][0], p[0][1], af, LX[6 + b][k4 + t4][g]);
            dmma884(p[1][0], p[1][1], af, LX[6 + b][k4 + t4][8 + g]);
        }
#pragma unroll
        for (int fn = 0; fn < 2; ++fn) {
            P[16 * b + 8 * fn + 2 * t4][r0 + g] = p[fn][0];
            P[16 * b + 8 * fn + 2 * t4 + 1][r0 + g] = p[fn][1];
        }
        __syncwarp();
    }
}

// The same four stages with the operands taken straight from global memory as the producer publishes them (BC_SENTINEL): for the
// followers of a diagonal step that is still running.  Returns false when a bounded wait ran out.
template <int LD>
__device__ __forceinline__ bool bc_substitute_rows_staged(double (*P)[LD], int r0, const double* __restrict__ Lxg, double (*Q)[BC_QLD], int g, int t4) {
    bool ok = true;
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
        double q[2][2];
#pragma unroll
        for (int fn = 0; fn < 2; ++fn) {
            q[fn][0] = -P[16 * b + 8 * fn + 2 * t4][r0 + g];
            q[fn][1] = -P[16 * b + 8 * fn + 2 * t4 + 1][r0 + g];
        }
        const double* Xb = Lxg + (size_t)(6 + b) * BC_XBLK;
        double xf[4][2];
        auto load_x = [&]() {
            bool bad = false;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int fn = 0; fn < 2; ++fn) {
                    xf[kk][fn] = bc_ld_l2(Xb + (4 * kk + t4) * BC_XLD + 8 * fn + g);
                    bad |= bc_is_sentinel(xf[kk][fn]);
                }
            return bad;
        };
        bool xbad = true;
        if (b > 0) {
            double lf[3][4][2];
            int spins = 0;
            bool bad;
            do {
                bad = false;
#pragma unroll
                for (int jb = 0; jb < 3; ++jb)
                    if (jb < b) {
                        const double* Lb = Lxg + (size_t)bc_lt_block(b, jb) * BC_XBLK;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                            for (int fn = 0; fn < 2; ++fn) {
                                lf[jb][kk][fn] = bc_ld_l2(Lb + (4 * kk + t4) * BC_XLD + 8 * fn + g);
                                bad |= bc_is_sentinel(lf[jb][kk][fn]);
                            }
                    }
                if (spins == 0) xbad = load_x();
                bad = __any_sync(0xffffffffu, bad);
            } while (bad && ++spins < (1 << 22));
            ok &= !bad;
#pragma unroll
            for (int jb = 0; jb < 3; ++jb)
                if (jb < b) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const double af = P[16 * jb + 4 * kk + t4][r0 + g];
#pragma unroll
                        for (int fn = 0; fn < 2; ++fn) dmma884(q[fn][0], q[fn][1], af, lf[jb][kk][fn]);
                    }
                }
        }
#pragma unroll
        for (int fn = 0; fn < 2; ++fn) {
            Q[8 * fn + 2 * t4][g] = -q[fn][0];
            Q[8 * fn + 2 * t4 + 1][g] = -q[fn][1];
        }
        {
            int spins = 0;
            xbad = __any_sync(0xffffffffu, xbad);
            while (xbad && ++spins < (1 << 22)) xbad = __any_sync(0xffffffffu, load_x());
            ok &= !xbad;
        }
        __syncwarp();
        double p[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const double af = Q[4 * kk + t4][g];
            if (kk < 2) dmma884(p[0][0], p[0][1], af, xf[kk][0]);
            dmma884(p[1][0], p[1][1], af, xf[kk][1]);
        }
#pragma unroll
        for (int fn = 0; fn < 2; ++fn) {
            P[16 * b + 8 * fn + 2 * t4][r0 + g] = p[fn][0];
            P[16 * b + 8 * fn + 2 * t4 + 1][r0 + g] = p[fn][1];
        }
        __syncwarp();
    }
    return ok;
}

template <int SPLIT>  // CTAs per tile: 2 (32 rows each) or 4 (16 rows: the urgent pair, where the launch is as long as one CTA)
__global__ void __launch_bounds__(BC_NEXT_THREADS)
    bc_next_kernel(double* Z, int ldz, const double* __restrict__ LxAll, const int* __restrict__ guard, int kblk, int nT, int TW, int part,
                   int* trailCnt, int waitNext, int* __restrict__ status, int tl) {
    // The urgent launch (SPLIT = 4) sits on the chain stream behind diag(k) as a programmatic dependent: it starts while diag(k) is
    // still factoring, lets diag(k+1) in at once, never waits for diag(k) to COMPLETE -- its T tiles are final before it is launched
    // (next(k-1) counts its CTAs off in trailCnt[16 + k-1]; polled here), and L_kk arrives block row by block row over the
    // sentinels, as for diag(k+1).
    constexpr bool STAGED = SPLIT == 4;
    if (STAGED)
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    else
        pdl_wait();
    constexpr int ROWS = BC_T / SPLIT, RF = ROWS / 8, CW = 8 / RF, NB = 8 / CW;
    const int bid = (int)(blockIdx.x / SPLIT), half = (int)(blockIdx.x % SPLIT);
    const int q = nT - kblk - 1;
    const bool urgent = part == BC_PART_URGENT;  // T(k+2, k+1), T(k+2, k+2): what diag(k+2) waits for
    if (*guard) {
        if (threadIdx.x == 0) atomicAdd(trailCnt + (urgent ? 0 : 16) + kblk, 1);
        return;
    }
    if (STAGED && waitNext > 0) {
        if (threadIdx.x == 0) {
            int spins = 0, seen;
            do {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(trailCnt + 16 + (kblk - 1)) : "memory");
            } while (seen < waitNext && ++spins < (1 << 24));
            if (seen < waitNext) atomicOr(status, 8);
        }
        __syncthreads();
    }
    TL_MARK(tl, 0);
    int ta, tb;
    if (urgent) {
        ta = kblk + 2;
        tb = bid == 0 ? kblk + 1 : kblk + 2;
    } else {
        const int nS1 = q >= 3 ? q - 2 : 0;
        if (bid < nS1) {
            ta = kblk + 3 + bid;
            tb = kblk + 1;
        } else if (q >= 3 && bid < nS1 + 2) {
            ta = kblk + 3;
            tb = bid == nS1 ? kblk + 2 : kblk + 3;
        } else {
            ta = nT + (bid - nS1 - (q >= 3 ? 2 : 0));
            tb = kblk + 1;
        }
    }
    extern __shared__ __align__(128) unsigned char bcn_smem_raw[];
    BcNextSmem& sm = *reinterpret_cast<BcNextSmem*>(bcn_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    if (!STAGED && tid == 0) {
        mbar_init(&sm.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&sm.bar, (uint32_t)(BC_LX * 8));
        bulk_g2s(&sm.LX[0][0][0], LxAll + (size_t)kblk * BC_LX, (uint32_t)(BC_LX * 8), &sm.bar);
    }
    const double* Tb = Z + (size_t)(kblk * BC_T) * ldz + (size_t)tb * BC_T;
    const double* Ta = Z + (size_t)(kblk * BC_T) * ldz + (size_t)ta * BC_T + ROWS * half;
    {
        // 2048 + 32 ROWS 16-byte words: every load of a thread in flight before its first store (a load -> store loop pays one L2 round
        // trip per iteration: 3 us of the launch)
        double2 vb[6], va[3];
#pragma unroll
        for (int u = 0; u < 6; ++u) {
            const int w = tid + BC_NEXT_THREADS * u;
            if (w < BC_T * 32) vb[u] = *reinterpret_cast<const double2*>(Tb + (size_t)(w >> 5) * ldz + (w & 31) * 2);
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int w = tid + BC_NEXT_THREADS * u;
            if (w < BC_T * (ROWS / 2)) va[u] = *reinterpret_cast<const double2*>(Ta + (size_t)(w / (ROWS / 2)) * ldz + (w % (ROWS / 2)) * 2);
        }
#pragma unroll
        for (int u = 0; u < 6; ++u) {
            const int w = tid + BC_NEXT_THREADS * u;
            if (w < BC_T * 32) *reinterpret_cast<double2*>(&sm.Pb[w >> 5][(w & 31) * 2]) = vb[u];
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int w = tid + BC_NEXT_THREADS * u;
            if (w < BC_T * (ROWS / 2)) *reinterpret_cast<double2*>(&sm.Pa[w / (ROWS / 2)][(w % (ROWS / 2)) * 2]) = va[u];
        }
    }
    // the tile itself, straight into the accumulator fragments of warps 0-7: rows wm .. wm+7 of this half, columns wn .. wn+31
    const int wm = (warp / CW) * 8, wn = (warp % CW) * (BC_T / CW);
    double* cbase = Z + (size_t)(tb * BC_T + wn + 2 * t4) * ldz + (size_t)ta * BC_T + ROWS * half + wm + g;
    double acc[NB][2];
    if (warp < 8) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            acc[b][0] = -cbase[(size_t)(b * 8) * ldz];
            acc[b][1] = -cbase[(size_t)(b * 8 + 1) * ldz];
        }
    }
    __syncthreads();
    if (STAGED) {
        const double* Lxg = LxAll + (size_t)kblk * BC_LX;
        bool ok = true;
        if (warp < 8)
            ok = bc_substitute_rows_staged<YB_LD>(sm.Pb, 8 * warp, Lxg, sm.Q[warp], g, t4);
        else if (warp < 8 + RF)
            ok = bc_substitute_rows_staged<BC_PA_LD>(sm.Pa, 8 * (warp - 8), Lxg, sm.Q[warp], g, t4);
        if (!ok) atomicOr(status, 8);
    } else {
        mbar_wait(&sm.bar, 0);
        if (warp < 8)
            bc_substitute_rows<YB_LD>(sm.Pb, 8 * warp, sm.LX, sm.Q[warp], g, t4);
        else if (warp < 8 + RF)
            bc_substitute_rows<BC_PA_LD>(sm.Pa, 8 * (warp - 8), sm.LX, sm.Q[warp], g, t4);
    }
    __syncthreads();
    if (warp < 8) {
#pragma unroll 4
        for (int k4 = 0; k4 < BC_T; k4 += 4) {
            const double af = sm.Pa[k4 + t4][wm + g];
#pragma unroll
            for (int b = 0; b < NB; ++b) dmma884(acc[b][0], acc[b][1], af, sm.Pb[k4 + t4][wn + 8 * b + g]);
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            cbase[(size_t)(b * 8) * ldz] = -acc[b][0];
            cbase[(size_t)(b * 8 + 1) * ldz] = -acc[b][1];
        }
    }
    __syncthreads();
    if (tid == 0) {  // urgent: diag(k+2) counts these; next: urgent(k+1) does
        __threadfence();
        atomicAdd(trailCnt + (urgent ? 0 : 16) + kblk, 1);
    }
    TL_MARK(tl, 1);
}

}  // namespace eqvio
