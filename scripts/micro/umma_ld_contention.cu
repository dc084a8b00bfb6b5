// Microbenchmark: do tcgen05.ld (epilogue) and tcgen05.mma (N=128, SS) slow each other down?
// warp 0 lane 0 issues MMAs into TMEM columns [0,256); warps 4..4+NLD-1 run tcgen05.ld.32x32b.x32 on columns [256,512).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t ph) {
    uint32_t ok; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(ph) : "memory"); return ok; }
constexpr uint64_t DESC_BASE = (uint64_t(128 >> 4) << 16) | (uint64_t(2048 >> 4) << 32) | (uint64_t(1) << 46);
__device__ __forceinline__ uint64_t make_desc(uint32_t a) { return DESC_BASE | uint64_t((a >> 4) & 0x3FFF); }

__global__ void __launch_bounds__(640, 1) k(int mma_iters, int ld_iters, int nld, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar; __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(128 >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    if (warp == 0) {
        if (lane == 0 && mma_iters > 0) {
            const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32768);
            long long t0 = clock64();
            for (int i = 0; i < mma_iters; ++i) {
                const int kk = i & 7;
                const uint32_t d = tm + uint32_t(((i >> 3) & 1) * 128);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(d), "l"(make_desc(a_addr + kk * 256)), "l"(make_desc(b_addr + kk * 256)), "r"(IDESC), "r"(uint32_t(kk > 0)) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            while (!mbar_try(&bar, 0)) {}
            out[blockIdx.x * 2] = clock64() - t0;
        }
    } else if (warp >= 4 && warp < 4 + nld) {
        const uint32_t base = tm + (uint32_t((warp & 3) * 32) << 16) + 256u;
        uint32_t acc = 0;
        long long t0 = clock64();
        for (int i = 0; i < ld_iters; ++i) {
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(base + uint32_t((i * 32) & 255)) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= r[j];
        }
        if (lane == 0 && warp == 4) out[blockIdx.x * 2 + 1] = clock64() - t0;
        if (acc == 0x12345678u) out[1000] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}
int main() {
    long long* out; cudaMalloc(&out, 8192 * 8);
    const int smem = 65536 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    struct Cfg { int mma, ld, nld; const char* name; } cfgs[] = {
        {16000, 0, 0, "MMA alone"}, {0, 8000, 16, "16 ld-warps alone"}, {16000, 8000, 16, "MMA + 16 ld-warps"},
        {16000, 8000, 4, "MMA + 4 ld-warps"}, {0, 8000, 4, "4 ld-warps alone"}};
    for (auto& c : cfgs) {
        for (int rep = 0; rep < 2; ++rep) { k<<<148, 640, smem>>>(c.mma, c.ld, c.nld, out); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; } }
        long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("%-22s: ", c.name);
        if (c.mma) printf("%.1f cycles/MMA(N=128)  ", double(h[0]) / c.mma);
        if (c.ld) printf("%.1f cycles per ld.x32 per warp", double(h[1]) / c.ld);
        printf("\n");
    }
    return 0;
}
