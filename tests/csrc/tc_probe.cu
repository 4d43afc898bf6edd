// Stand-alone probe of the tcgen05 path used by the split-bf16 downdate: D(128x128,f32) = A(128xK) B(128xK)^T with bf16
// operands placed by ordinary stores in the canonical K-major no-swizzle core-matrix layout, accumulator in TMEM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu ; run on a B200.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int R = 128, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
    return d;                // layout_type = 0 (no swizzle), base_offset = 0
}

__global__ void __launch_bounds__(128) tc_probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D) {
    __shared__ __align__(128) __nv_bfloat16 sA[R * K];
    __shared__ __align__(128) __nv_bfloat16 sB[R * K];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t SBO = 128, LBO = (R / 8) * 128;
    for (int t = tid; t < R * K; t += 128) {
        const int r = t / K, k = t % K;
        const uint32_t off = (k / 8) * LBO + (r / 8) * SBO + (r % 8) * 16 + (k % 8) * 2;
        *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(sA) + off) = A[t];
        *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(sB) + off) = B[t];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        // instruction descriptor: c = F32 (1 << 4), a = b = BF16 (1 << 7, 1 << 10), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int kk = 0; kk < K / 16; ++kk) {
            const uint64_t da = make_desc(smem_u32(sA) + (2 * kk) * LBO, LBO, SBO);
            const uint64_t db = make_desc(smem_u32(sB) + (2 * kk) * LBO, LBO, SBO);
            const uint32_t acc = kk > 0 ? 1u : 0u;
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "setp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
                "}\n" ::"r"(tm),
                "l"(da), "l"(db), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    // wait for the MMAs
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra DONE;\n"
        "bra WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(&mbar))
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = 32 * warp + lane;
    for (int c0 = 0; c0 < 128; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tm + ((uint32_t)(32 * warp) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) D[row * 128 + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tm) : "memory");
}

int main() {
    std::vector<__nv_bfloat16> hA(R * K), hB(R * K);
    std::vector<float> fA(R * K), fB(R * K);
    srand(1);
    for (int i = 0; i < R * K; ++i) {
        hA[i] = __float2bfloat16((rand() % 200 - 100) / 64.0f);
        hB[i] = __float2bfloat16((rand() % 200 - 100) / 64.0f);
        fA[i] = __bfloat162float(hA[i]);
        fB[i] = __bfloat162float(hB[i]);
    }
    __nv_bfloat16 *dA, *dB;
    float* dD;
    cudaMalloc(&dA, R * K * 2);
    cudaMalloc(&dB, R * K * 2);
    cudaMalloc(&dD, R * R * 4);
    cudaMemcpy(dA, hA.data(), R * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), R * K * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, R * R * 4);
    tc_probe<<<1, 128>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<float> hD(R * R);
    cudaMemcpy(hD.data(), dD, R * R * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < R; ++j) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)fA[i * K + k] * fB[j * K + k];
            maxerr = fmax(maxerr, fabs(s - hD[i * R + j]));
            maxref = fmax(maxref, fabs(s));
        }
    printf("max |D - ref| = %g (max |ref| = %g)  D[0][0..3] = %g %g %g %g\n", maxerr, maxref, hD[0], hD[1], hD[2], hD[3]);
    return maxerr < 1e-3 * maxref ? 0 : 1;
}
