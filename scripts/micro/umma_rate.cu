// Microbenchmark: cycles per tcgen05.mma.kind::f16 (M=128, K=16) for N in {64,128,256}, A from smem (SS) or TMEM (TS).
// One thread issues `iters` MMAs back to back (accumulating into the same TMEM tile), commits, waits.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t ph) {
    uint32_t ok; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(ph) : "memory"); return ok; }
constexpr uint64_t DESC_BASE = (uint64_t(128 >> 4) << 16) | (uint64_t(2048 >> 4) << 32) | (uint64_t(1) << 46);
__device__ __forceinline__ uint64_t make_desc(uint32_t a) { return DESC_BASE | uint64_t((a >> 4) & 0x3FFF); }

template <int N, bool TS, int NACC>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar; __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    if (threadIdx.x == 0) {
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32768);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int k = i & 7;
            const uint32_t d = tm + (TS ? 64u : 0u) + uint32_t(((i >> 3) % NACC) * N);
            const uint64_t bd = make_desc(b_addr + k * 256);
            if (TS) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                             ::"r"(d), "r"(tm + uint32_t(k * 8)), "l"(bd), "r"(IDESC), "r"(uint32_t(k > 0)) : "memory");
            } else {
                const uint64_t ad = make_desc(a_addr + k * 256);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(d), "l"(ad), "l"(bd), "r"(IDESC), "r"(uint32_t(k > 0)) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        while (!mbar_try(&bar, 0)) {}
        long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}
template <int N, bool TS, int NACC>
void run(const char* name, long long* cyc) {
    const int iters = 8000, smem = 32768 + 65536 + 1024;
    cudaFuncSetAttribute(k<N, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int grid : {1, 148}) {
        for (int rep = 0; rep < 2; ++rep) { k<N, TS, NACC><<<grid, 128, smem>>>(iters, cyc); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("%s error %s\n", name, cudaGetErrorString(e)); exit(1); } }
        long long h[148]; cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%-26s grid %3d: %.1f cycles/MMA (ideal %d) -> %.0f%% of peak\n", name, grid, double(mx) / iters, N / 2, 100.0 * (N / 2) * iters / mx);
    }
}
int main() {
    long long* cyc; cudaMalloc(&cyc, 148 * 8);
    run<64, false, 1>("SS N=64  1 acc", cyc);  run<64, false, 4>("SS N=64  4 acc", cyc);
    run<128, false, 2>("SS N=128 2 acc", cyc); run<256, false, 2>("SS N=256 2 acc", cyc);
    run<64, true, 4>("TS N=64  4 acc", cyc);   run<128, true, 2>("TS N=128 2 acc", cyc);
    run<192, true, 2>("TS N=192 2 acc", cyc);  run<256, true, 1>("TS N=256 1 acc", cyc);
    return 0;
}
