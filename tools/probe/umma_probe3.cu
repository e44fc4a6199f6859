// Probe 2: tcgen05.mma kind::f16 with SWIZZLE_128B canonical layouts, K-major and MN-major views of the SAME bytes.
// Matrix [R rows x C cols] fp32 stored as blocks [C/32][R/8][8 rows x 128 B], 16-byte chunks XOR-swizzled with row%8.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <stdint.h>
#include <cuda_bf16.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ inline int sw_off(int r, int c, int R) {  // bf16 element index; blocks of 64 columns (128 B rows)
    return (((c >> 6) * (R >> 3) + (r >> 3)) * 1024 + (r & 7) * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4) + (c & 7) * 2) >> 1;
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// K-major view of a stored [R x C] matrix (mn = row, k = col): descriptor for k-step ks (8 columns)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t base, int R, int ks) {
    return make_desc_sw128(base + (ks >> 2) * (R >> 3) * 1024 + (ks & 3) * 32, 16, 1024);
}
// MN-major view (k = row, mn = col): descriptor for k-step ks (16 rows = two 8-row atoms SBO apart)
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t base, int R, int ks) {
    return make_desc_sw128(base + ks * 2048, (R >> 3) * 1024, 1024);
}

template <int M, int N, int K>
__global__ void probe_kernel(const __nv_bfloat16* gA, const __nv_bfloat16* gB, float* gD, int form) {
    extern __shared__ __align__(1024) float smem[];
    __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem);
    __nv_bfloat16* sB = sA + 128 * 128;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * 128; i += blockDim.x) sA[i] = gA[i];
    for (int i = tid; i < 128 * 128; i += blockDim.x) sB[i] = gB[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        for (int ks = 0; ks < K / 16; ++ks) {
            uint64_t ad, bd;
            uint32_t idesc;
            if (form == 0) {         // act[M x K] K-major ; W stored [in=K rows][out=N cols], B(n=out,k=in) MN-major view
                ad = desc_kmajor(smem_u32(sA), M, ks);
                bd = desc_mnmajor(smem_u32(sB), K, ks);
                idesc = make_idesc_tf32(M, N, 0, 1);
            } else if (form == 1) {  // dY[M x K] K-major ; W stored [in=N rows][out=K cols], B(n=in,k=out) K-major view
                ad = desc_kmajor(smem_u32(sA), M, ks);
                bd = desc_kmajor(smem_u32(sB), N, ks);
                idesc = make_idesc_tf32(M, N, 0, 0);
            } else {                 // act stored [samples=K rows][feat=M cols] -> A(m=feat,k=sample) MN view; dY stored [K rows][N cols] -> B MN view
                ad = desc_mnmajor(smem_u32(sA), K, ks);
                bd = desc_mnmajor(smem_u32(sB), K, ks);
                idesc = make_idesc_tf32(M, N, 1, 1);
            }
            const uint32_t acc = ks > 0 ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                         "l"(ad), "l"(bd), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)));
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u));
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        const int row = warp * 32 + lane;
        for (int j = 0; j < 16; ++j) gD[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}
static float tf32_trunc(float x) { return __bfloat162float(__float2bfloat16(x)); }
template <int M, int N, int K>
int run(int form) {
    std::vector<float> A((size_t)M * K), B((size_t)K * N), Dref((size_t)M * N, 0.f);
    srand(7 + form);
    for (auto& x : A) x = tf32_trunc((rand() % 2001 - 1000) / 1000.f);
    for (auto& x : B) x = tf32_trunc((rand() % 2001 - 1000) / 1000.f);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[k * N + n]; Dref[m * N + n] = (float)s; }
    std::vector<__nv_bfloat16> pa(128 * 128), pb(128 * 128);
    if (form == 0) {
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) pa[sw_off(m, k, M)] = __float2bfloat16(A[m * K + k]);          // act [M rows][K cols]
        for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) pb[sw_off(k, n, K)] = __float2bfloat16(B[k * N + n]);          // W [in=K rows][out=N cols]
    } else if (form == 1) {
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) pa[sw_off(m, k, M)] = __float2bfloat16(A[m * K + k]);          // dY [M rows][K cols]
        for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) pb[sw_off(n, k, N)] = __float2bfloat16(B[k * N + n]);          // W [in=N rows][out=K cols]
    } else {
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) pa[sw_off(k, m, K)] = __float2bfloat16(A[m * K + k]);          // act [samples=K rows][feat=M cols]
        for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) pb[sw_off(k, n, K)] = __float2bfloat16(B[k * N + n]);          // dY [samples=K rows][N cols]
    }
    __nv_bfloat16 *dA, *dB; float* dD;
    CK(cudaMalloc(&dA, pa.size() * 2)); CK(cudaMalloc(&dB, pb.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dA, pa.data(), pa.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, pb.data(), pb.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, 128 * N * 4));
    const int smem = (128 * 128 + 128 * 128) * 2 + 1024;
    CK(cudaFuncSetAttribute(probe_kernel<M, N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<M, N, K><<<1, 128, smem>>>(dA, dB, dD, form);
    CK(cudaDeviceSynchronize());
    std::vector<float> D(128 * N);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    if (M < 128) {
        printf("form %d M=%d lane map: ", form, M);
        for (int lane = 0; lane < 128; ++lane) {
            int found = -1;
            for (int m = 0; m < M; ++m) { bool ok = true; for (int n = 0; n < N && ok; ++n) ok = fabsf(D[lane * N + n] - Dref[m * N + n]) < 1e-3f * K; if (ok) { found = m; break; } }
            printf("%d ", found);
        }
        printf("\n");
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
        return 0;
    }
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < M * N; ++i) { double e = fabs(D[i] - Dref[i]); if (e > maxerr) maxerr = e; if (e > 1e-2) ++bad; }
    printf("BF16 SW128 form %d M=%d N=%d K=%d: max abs err %.3e, bad %d / %d   D[0..3] = %f %f %f %f  ref %f %f %f %f\n", form, M, N, K, maxerr, bad, M * N,
           D[0], D[1], D[2], D[3], Dref[0], Dref[1], Dref[2], Dref[3]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return bad;
}
int main() {
    int bad = 0;
    bad += run<128, 64, 64>(1);
    bad += run<128, 64, 64>(0);
    bad += run<128, 64, 128>(0);
    bad += run<128, 32, 64>(0);
    bad += run<64, 64, 128>(2);
    bad += run<128, 64, 128>(2);
    bad += run<128, 128, 128>(2);
    printf(bad ? "PROBE3: mismatches\n" : "PROBE3 OK\n");
    return 0;
}
