// Stand-alone probe for the phases of bc_diag_kernel's block-column loop (sm_100a): what does a warp pay for
//   (1) a named barrier among W warps, (2) the rank-4 update of a half tile (14 LDS.128 + 8 DMUL + 32 DFMA),
//   (3) the panel operation of a half tile, (4) the fraction-free 4x4 elimination with its four reciprocals,
//   (5) clock() itself -- each alone in a loop, W warps resident.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bc_probe bc_probe.cu && ./bc_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

__global__ void probe(int mode, int iters, double* out, long long* cyc) {
    __shared__ __align__(16) double tiles[64][18];
    __shared__ __align__(16) double dc[4];
    const int tid = threadIdx.x;
    for (int i = tid; i < 64 * 18; i += blockDim.x) (&tiles[0][0])[i] = 1e-3 * (i % 37);
    if (tid < 4) dc[tid] = 0.5 + 0.1 * tid;
    __syncthreads();
    double a[2][4];
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 4; ++c) a[r][c] = 1.0 + 0.01 * (tid + r + c);
    const int TI = (tid >> 1) & 63, TK = (tid >> 6) & 63, h = tid & 1;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (mode == 0) {
            asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x) : "memory");
        } else if (mode == 1) {
            const double2* cp = reinterpret_cast<const double2*>(dc);
            const double2 c01 = cp[0], c23 = cp[1];
            const double2* lp = reinterpret_cast<const double2*>(&tiles[TI][8 * h]);
            const double2* pp = reinterpret_cast<const double2*>(&tiles[(TK + it) & 63][0]);
            double li[2][4];
            for (int r = 0; r < 2; ++r) {
                const double2 x = lp[2 * r], y = lp[2 * r + 1];
                li[r][0] = x.x * c01.x;
                li[r][1] = x.y * c01.y;
                li[r][2] = y.x * c23.x;
                li[r][3] = y.y * c23.y;
            }
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const double2 x = pp[2 * cc], y = pp[2 * cc + 1];
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    double acc = a[r][cc];
                    acc -= li[r][0] * x.x;
                    acc -= li[r][1] * x.y;
                    acc -= li[r][2] * y.x;
                    acc -= li[r][3] * y.y;
                    a[r][cc] = acc;
                }
            }
        } else if (mode == 2) {
            const double2 c01 = *reinterpret_cast<const double2*>(dc);
            const double c2 = dc[2];
            const double2* dt = reinterpret_cast<const double2*>(&tiles[it & 63][0]);
            const double d10 = dt[2].x;
            const double2 d2 = dt[4], d3 = dt[6];
            const double d32 = dt[7].x;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const double t0_ = a[r][0] * c01.x;
                a[r][1] -= t0_ * d10;
                a[r][2] -= t0_ * d2.x;
                a[r][3] -= t0_ * d3.x;
                const double t1 = a[r][1] * c01.y;
                a[r][2] -= t1 * d2.y;
                a[r][3] -= t1 * d3.y;
                const double t2 = a[r][2] * c2;
                a[r][3] -= t2 * d32;
            }
            double2* o = reinterpret_cast<double2*>(&tiles[TI][8 * h]);
            o[0] = make_double2(a[0][0], a[0][1]);
            o[1] = make_double2(a[0][2], a[0][3]);
            o[2] = make_double2(a[1][0], a[1][1]);
            o[3] = make_double2(a[1][2], a[1][3]);
        } else if (mode == 3) {
            if ((tid & 31) == 0) {
                const double a00 = a[0][0], a10 = a[1][0], a20 = a[0][1], a30 = a[1][1];
                const double m11 = a[1][1] * a00 - a10 * a10, m21 = a[0][2] * a00 - a20 * a10, m22 = a[0][3] * a00 - a20 * a20;
                const double m31 = a[1][2] * a00 - a30 * a10, m32 = a[1][3] * a00 - a30 * a20, m33 = (a[1][3] + 3.0) * a00 - a30 * a30;
                const double n22 = m22 * m11 - m21 * m21, n32 = m32 * m11 - m31 * m21, n33 = m33 * m11 - m31 * m31;
                const double p33 = n33 * n22 - n32 * n32;
                const double r0 = fast_rcp(a00), r1 = fast_rcp(m11), r2_ = fast_rcp(n22), r3 = fast_rcp(p33);
                const double s2 = r0 * r1, s3 = s2 * r2_, e1 = a00 * m11;
                double2* d = reinterpret_cast<double2*>(dc);
                d[0] = make_double2(r0, a00 * r1);
                d[1] = make_double2(e1 * r2_, (e1 * n22) * r3);
                double2* dt = reinterpret_cast<double2*>(&tiles[it & 63][0]);
                dt[2] = make_double2(a10, 0.0);
                dt[4] = make_double2(a20, m21 * r0);
                dt[6] = make_double2(a30, m31 * r0);
                dt[7] = make_double2(n32 * s2, 0.0);
                a[0][0] = 1.0 + p33 * s3 * 1e-9;  // carried dependency: the next elimination waits for this one
            }
        } else if (mode == 4) {
            a[0][0] += (double)(int)clock() * 1e-30;
        } else if (mode == 5) {
            // 32 dependent-free DFMAs from one warp: issue rate
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    a[r][c] = fma(a[r][c], 1.0000001, 1e-9);
                    a[r][c] = fma(a[r][c], 1.0000001, 1e-9);
                    a[r][c] = fma(a[r][c], 1.0000001, 1e-9);
                    a[r][c] = fma(a[r][c], 1.0000001, 1e-9);
                }
        } else if (mode == 6) {
            // shared-memory round trip: store a half tile, barrier-free read back of a neighbour's (LDS latency chain)
            double2* o = reinterpret_cast<double2*>(&tiles[TI][8 * h]);
            o[0] = make_double2(a[0][0], a[0][1]);
            __syncwarp();
            const double2 v = *reinterpret_cast<const double2*>(&tiles[(TI + 1) & 63][8 * h]);
            a[0][0] = v.x + 1e-9;
            a[0][1] = v.y;
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 4; ++c) s += a[r][c];
    out[blockIdx.x * blockDim.x + tid] = s;
    if (tid == 0) *cyc = t1 - t0;
}

int main() {
    double* out;
    long long *cyc, h;
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&cyc, 8);
    const char* names[] = {"named barrier", "half-tile rank-4 update (14 LDS.128 + 8 DMUL + 32 DFMA)", "half-tile panel op (+ 4 STS.128)",
                           "4x4 fraction-free elimination (lane 0 of every warp)", "clock()", "32 DFMA, 8 chains of 4", "STS.128 -> LDS.128 round trip"};
    const int iters = 2048;
    for (int mode = 0; mode < 7; ++mode)
        for (int warps : {1, 2, 4, 5, 8, 9, 16}) {
            probe<<<1, 32 * warps>>>(mode, iters, out, cyc);
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-60s warps=%2d: %7.1f cycles per iteration\n", names[mode], warps, (double)h / iters);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
