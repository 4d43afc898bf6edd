// Stand-alone probe: fp64 FMA latency / throughput per SM and DMMA (mma.sync.m8n8k4.f64) throughput on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu && ./fp64_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, long long* cyc) {
    double a[ILP];
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b, c);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void dmma_kernel(double* out, int iters, long long* cyc) {
    double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; k += 2) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[k]), "+d"(c0[k + 1]) : "d"(a), "d"(b));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[k]), "+d"(c1[k + 1]) : "d"(a), "d"(b));
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c0[2] + c0[3] + c1[0] + c1[1] + c1[2] + c1[3];
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    double* out;
    long long *cyc, h;
    cudaMalloc(&out, 1 << 24);
    cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int warps : {1, 4, 8, 16, 32}) {
        dfma_kernel<1><<<1, 32 * warps>>>(out, iters, cyc);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA ILP1 warps=%2d: %.2f cycles per dependent FMA; %.2f FMA/clk/SM\n", warps, (double)h / iters, 32.0 * warps * iters / h);
        dfma_kernel<8><<<1, 32 * warps>>>(out, iters, cyc);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA ILP8 warps=%2d: %.2f cycles per 8 FMAs; %.2f FMA/clk/SM\n", warps, (double)h / iters, 8 * 32.0 * warps * iters / h);
    }
    for (int warps : {1, 4, 8, 16}) {
        dmma_kernel<<<1, 32 * warps>>>(out, iters, cyc);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DMMA m8n8k4 warps=%2d: %.2f cycles per 4 MMAs per warp; %.1f fp64 FMA-equivalents/clk/SM\n", warps, (double)h / iters,
               4.0 * 256 * warps * iters / h);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
