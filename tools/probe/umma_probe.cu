// Probe: validate tcgen05.mma kind::tf32 descriptors / canonical no-swizzle layouts / TMEM load mapping on sm_100a.
//   D[M x N] = A[M x K] * B, fp32 accumulate in TMEM, for the three operand forms the MLP kernels need:
//   form 0 (forward):   A K-major  (act [sample][feat]),  B MN-major (W[in][out] blocked 8in x 4out), D = act * W
//   form 1 (backward):  A K-major  (dY  [sample][out]),   B K-major  (same W bytes),                   D = dY * W^T
//   form 2 (dW):        A MN-major (act bytes as [feat][sample]), B MN-major (dY bytes as [sample][out]), D = act^T * dY
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // version 1 (Blackwell)
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                    // c_format F32
    d |= 2u << 7;                    // a_format TF32
    d |= 2u << 10;                   // b_format TF32
    d |= (uint32_t)a_mn_major << 15;
    d |= (uint32_t)b_mn_major << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

// act blocked: element (row r, col c) of an [R x C] matrix, cores of 8 rows x 4 cols (128 B), cores ordered [r/8][c/4]
__host__ __device__ inline int blk_rc(int r, int c, int C) { return ((r >> 3) * (C >> 2) + (c >> 2)) * 32 + (r & 7) * 4 + (c & 3); }
// W blocked: element (in, out) of W[K_in x N_out]: cores of 8 in x 4 out, ordered [in/8][out/4]
__host__ __device__ inline int blk_w(int in, int out, int N) { return ((in >> 3) * (N >> 2) + (out >> 2)) * 32 + (in & 7) * 4 + (out & 3); }

template <int M, int N, int K>
__global__ void probe_kernel(const float* gA, const float* gB, float* gD, int form, int dump_lanes, int variant) {
    extern __shared__ __align__(128) float smem[];
    float* sA = smem;                 // up to 128*64
    float* sB = smem + 128 * 64;      // up to 64*128
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nA = (form == 2) ? K * M : M * K;   // form 2: A bytes are act [sample=K][feat=M] blocked
    const int nB = (form == 1) ? N * K : K * N;
    for (int i = tid; i < nA; i += blockDim.x) sA[i] = gA[i];
    for (int i = tid; i < nB; i += blockDim.x) sB[i] = gB[i];
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
        uint32_t idesc;
        for (int ks = 0; ks < K / 8; ++ks) {
            uint64_t ad, bd;
            if (form == 0) {
                // A K-major [M x K]: LBO = 128 (next 4-col chunk), SBO = (K/4)*128 (next 8-row group); k-step advances 2 chunks
                ad = make_desc(smem_u32(sA) + ks * 256, 128, (K / 4) * 128);
                // B = W[in=K][out=N] blocked, MN-major: SBO = 128 (next out/4 group), LBO = (N/4)*128 (next in/8 group)
                if (variant == 0) bd = make_desc(smem_u32(sB) + ks * (N / 4) * 128, (N / 4) * 128, 128);
                else if (variant == 1) bd = make_desc(smem_u32(sB) + ks * (N / 4) * 128, 128, (N / 4) * 128);      // LBO/SBO swapped
                else if (variant == 2) bd = make_desc(smem_u32(sB) + ks * (N / 4) * 128, (N / 4) * 128, 128) | (1ull << 52);  // lbo_mode 1
                else bd = make_desc(smem_u32(sB) + ks * (N / 4) * 128, 128, (N / 4) * 128) | (1ull << 52);
                idesc = make_idesc_tf32(M, N, 0, 1);
            } else if (form == 1) {
                // D[M x N] = dY[M x K] * W^T, W[in=N][out=K] blocked: as B (n = in, k = out) K-major:
                // 8 n-rows x 16 B cores; LBO (next out/4 chunk) = 128, SBO (next in/8 group) = (K/4)*128
                ad = make_desc(smem_u32(sA) + ks * 256, 128, (K / 4) * 128);
                bd = make_desc(smem_u32(sB) + ks * 256, 128, (K / 4) * 128);
                idesc = make_idesc_tf32(M, N, 0, 0);
            } else {
                // D[M=feat x N=out] = act^T * dY ; act bytes: [sample=K][feat=M] blocked (cores 8 samples x 4 feats)
                // as A (m = feat, k = sample) MN-major: SBO (next feat/4 group) = 128, LBO (next sample/8 group) = (M/4)*128
                ad = make_desc(smem_u32(sA) + ks * (M / 4) * 128, (M / 4) * 128, 128);
                // dY bytes: [sample=K][out=N] blocked; as B (n = out, k = sample) MN-major
                bd = make_desc(smem_u32(sB) + ks * (N / 4) * 128, (N / 4) * 128, 128);
                idesc = make_idesc_tf32(M, N, 1, 1);
            }
            const uint32_t acc = ks > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)));
    }
    // wait for the MMAs
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u));
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    // every warp reads its 32 lanes, N columns (chunks of 16)
    const int rows_to_dump = dump_lanes ? 128 : M;
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        const int row = warp * 32 + lane;
        if (row < rows_to_dump)
            for (int j = 0; j < 16; ++j) gD[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

template <int M, int N, int K>
int run(int form, int dump_lanes, int variant = 0) {
    // logical operands
    std::vector<float> A((size_t)M * K), B((size_t)K * N), Dref((size_t)M * N, 0.f);
    srand(1 + form);
    for (auto& x : A) x = tf32_trunc((rand() % 2001 - 1000) / 1000.f);
    for (auto& x : B) x = tf32_trunc((rand() % 2001 - 1000) / 1000.f);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[k * N + n]; Dref[m * N + n] = (float)s; }
    std::vector<float> pa((size_t)M * K), pb((size_t)K * N);
    if (form == 0) {        // A = act[M samples][K feats] blocked rows; B = W[K in][N out] blocked
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) pa[blk_rc(m, k, K)] = A[m * K + k];
        for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) pb[blk_w(k, n, N)] = B[k * N + n];
    } else if (form == 1) { // A = dY[M][K outs]; B logical [K][N] = W^T with W[in=N][out=K]: W(in=n,out=k) = B[k][n]
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) pa[blk_rc(m, k, K)] = A[m * K + k];
        for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) pb[blk_w(n, k, K)] = B[k * N + n];
    } else {                // A logical [M feats][K samples] = act^T, act[sample=k][feat=m]; B logical [K samples][N outs] = dY
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) pa[blk_rc(k, m, M)] = A[m * K + k];
        for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) pb[blk_rc(k, n, N)] = B[k * N + n];
    }
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, pa.size() * 4)); CK(cudaMalloc(&dB, pb.size() * 4)); CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dA, pa.data(), pa.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, pb.data(), pb.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, 128 * N * 4));
    const int smem = (128 * 64 + 64 * 128) * 4;
    CK(cudaFuncSetAttribute(probe_kernel<M, N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<M, N, K><<<1, 128, smem>>>(dA, dB, dD, form, dump_lanes, variant);
    CK(cudaDeviceSynchronize());
    std::vector<float> D(128 * N);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    if (dump_lanes) {   // where do the M rows land in TMEM lanes?
        printf("form %d M=%d lane map: ", form, M);
        for (int lane = 0; lane < 128; ++lane) {
            int found = -1;
            for (int m = 0; m < M; ++m) { bool ok = true; for (int n = 0; n < N && ok; ++n) ok = fabsf(D[lane * N + n] - Dref[m * N + n]) < 2e-3f * K; if (ok) { found = m; break; } }
            printf("%d ", found);
        }
        printf("\n");
        return 0;
    }
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < M * N; ++i) { double e = fabs(D[i] - Dref[i]); if (e > maxerr) maxerr = e; if (e > 1e-2) ++bad; }
    printf("variant %d form %d M=%d N=%d K=%d: max abs err %.3e, bad %d / %d   D[0..3] = %f %f %f %f  ref %f %f %f %f\n", variant, form, M, N, K, maxerr, bad, M * N,
           D[0], D[1], D[2], D[3], Dref[0], Dref[1], Dref[2], Dref[3]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return bad;
}

int main() {
    int bad = 0;
    for (int v = 0; v < 4; ++v) bad += run<128, 64, 64>(0, 0, v);
    bad += run<128, 64, 64>(1, 0);
    // does the MMA write at all? K = 8 single step, form 0
    for (int v = 0; v < 2; ++v) bad += run<128, 64, 8>(0, 0, v);
    printf(bad ? "PROBE: mismatches\n" : "PROBE OK\n");
    return 0;
}
