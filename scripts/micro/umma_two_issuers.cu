// Microbenchmark: does issuing tcgen05.mma (M128 N128 K16, SS) from TWO warps raise the issue rate above 1 per ~70 cycles?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t ph) {
    uint32_t ok; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(ph) : "memory"); return ok; }
constexpr uint64_t DESC_BASE = (uint64_t(128 >> 4) << 16) | (uint64_t(2048 >> 4) << 32) | (uint64_t(1) << 46);
__device__ __forceinline__ uint64_t make_desc(uint32_t a) { return DESC_BASE | uint64_t((a >> 4) & 0x3FFF); }
template <int N>
__global__ void __launch_bounds__(128, 1) k(int iters, int nissuers, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[4]; __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    long long t0 = clock64();
    if (warp < nissuers && lane == 0) {
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32768);
        for (int i = 0; i < iters; ++i) {
            const int kk = i & 7;
            const uint32_t d = tm + uint32_t(warp * N);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(d), "l"(make_desc(a_addr + kk * 256)), "l"(make_desc(b_addr + kk * 256)), "r"(IDESC), "r"(uint32_t(kk > 0)) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
        while (!mbar_try(&bar[warp], 0)) {}
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}
template <int N> void run(long long* out) {
    const int smem = 98304 + 1024, iters = 8000;
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int ni : {1, 2, 4}) {
        if (ni * N > 512) continue;
        for (int rep = 0; rep < 2; ++rep) { k<N><<<148, 128, smem>>>(iters, ni, out); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); } }
        long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("N=%3d issuers %d: %.1f cycles per MMA overall (ideal %d)\n", N, ni, double(h) / (double(iters) * ni), N / 2);
    }
}
int main() { long long* out; cudaMalloc(&out, 148 * 8); run<64>(out); run<128>(out); run<256>(out); return 0; }
