// Probe: how many tcgen05.mma instructions can one thread have in flight before the ISSUE blocks, and what one MMA of the
// shapes used by the U family costs (cycles per instruction, back to back into one accumulator).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_queue_probe umma_queue_probe.cu && ./umma_queue_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
constexpr int MAXN = 96;
__global__ void probe(long long* out, int n, int M, int N, int a_mn, int b_mn) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_s;
    if (tid == 0) {
        const uint32_t base = smem_u32(smem);
        const uint64_t ad = a_mn ? make_desc(base, 16384, 1024) : make_desc(base, 16, 1024);
        const uint64_t bd = b_mn ? make_desc(base + 32768, 8192, 1024) : make_desc(base + 32768, 16, 1024);
        const uint32_t id = make_idesc(M, N, a_mn, b_mn);
        long long t[MAXN + 2];
        t[0] = clock64();
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            mma(tmem, ad, bd, id, i ? 1u : 0u);
            t[i + 1] = clock64();
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
        t[n + 1] = clock64();
        for (int i = 0; i <= n + 1; ++i) out[i] = t[i] - t[0];
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}
int main() {
    long long* d;
    CK(cudaMalloc(&d, sizeof(long long) * (MAXN + 2)));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560));
    struct Shape { int M, N, a_mn, b_mn; const char* what; } shapes[] = {
        {128, 64, 0, 1, "fwd  M128 N64 A K-major, B MN-major"}, {128, 64, 0, 0, "bwd  M128 N64 A K-major, B K-major"},
        {128, 32, 0, 1, "head M128 N32"}, {64, 64, 1, 1, "dW   M64 N64 A MN, B MN"}, {64, 32, 1, 1, "dW   M64 N32 A MN, B MN"},
        {64, 8, 1, 0, "sum  M64 N8 A MN, B K"}};
    for (auto& s : shapes) {
        for (int n : {1, 2, 4, 8, 16, 32, 64, 96}) {
            probe<<<1, 128, 66560>>>(d, n, s.M, s.N, s.a_mn, s.b_mn);
            CK(cudaDeviceSynchronize());
            long long h[MAXN + 2];
            CK(cudaMemcpy(h, d, sizeof(long long) * (n + 2), cudaMemcpyDeviceToHost));
            printf("%-40s n=%2d  issue done at %6lld  complete at %6lld  (%.1f cyc/mma)", s.what, n, h[n], h[n + 1], (double)h[n + 1] / n);
            if (n == 96) {
                printf("\n    issue stamps:");
                for (int i = 1; i <= n; ++i) printf(" %lld", h[i] - h[i - 1]);
            }
            printf("\n");
        }
    }
    return 0;
}
