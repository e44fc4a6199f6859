// Probe: cost of the LL (data + sequence number) hand-overs of the U-family epoch kernel between 128 co-resident CTAs.
// Emulates one minibatch round: flush (each CTA writes W LL words to its slab) -> C (owner polls its word in the G slabs of its
// tower) -> partial all-to-all (16 B to every CTA) -> D (poll 2G partials, store the new word to G copies or to one shared copy)
// -> E (poll the W words of the own / shared copy).  Variants: scope of the LL accesses (volatile = relaxed.sys vs relaxed.gpu),
// private vs shared weight copy, nanosleep back-off.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int W = 7496, NTH = 256;

template <int SCOPE>
__device__ __forceinline__ void lst(uint2* p, unsigned d, unsigned s) {
    if (SCOPE == 0) asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(d), "r"(s) : "memory");
    else asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(d), "r"(s) : "memory");
}
template <int SCOPE>
__device__ __forceinline__ uint2 lld(const uint2* p) {
    uint2 v;
    if (SCOPE == 0) asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
template <int SCOPE>
__device__ __forceinline__ uint4 lld4(const void* p) {
    uint4 v;
    if (SCOPE == 0) asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
template <int SCOPE>
__device__ __forceinline__ void lst4(void* p, unsigned a, unsigned b, unsigned c, unsigned d) {
    if (SCOPE == 0) asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
    else asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// SHAREDW: 0 = one weight copy per destination CTA, 1 = one shared copy per tower polled by all G CTAs
template <int SCOPE, int SHAREDW, int SLEEP>
__global__ void __launch_bounds__(NTH, 1) round_kernel(uint2* ll, int rounds, long long* out, float* sink) {
    extern __shared__ unsigned char sm[];
    const int G = gridDim.x, tower = blockIdx.y, tid = threadIdx.x, me = tower * G + blockIdx.x;
    const int wpc = (((W + G - 1) / G) + 3) & ~3;
    const int myw = blockIdx.x * wpc + tid;
    const bool owns = tid < wpc && myw < W;
    const size_t slab_words = (size_t)4 * G * W, w_words = (size_t)4 * G * W;
    long long acc[6] = {0, 0, 0, 0, 0, 0};
    float keep = 0.f;
    __shared__ double red[8];
    for (int r = 1; r <= rounds; ++r) {
        const unsigned seq = r, par = r & 1;
        uint2* slab = ll + ((size_t)(par * 2 + tower) * G) * W;
        uint2* wl = ll + slab_words + ((size_t)(par * 2 + tower) * G) * W;
        uint2* sq = ll + slab_words + w_words + (size_t)par * (2 * G) * (2 * G) * 2;
        long long t0 = clock64();
        // flush
        for (int i = tid; i < W; i += NTH) lst<SCOPE>(slab + (size_t)blockIdx.x * W + i, __float_as_uint(1.0f + i), seq);
        long long t1 = clock64();
        // C
        float g = 0.f;
        if (owns) {
            for (int c0 = 0; c0 < G; c0 += 16) {
                uint2 v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = (c0 + i < G) ? lld<SCOPE>(slab + (size_t)(c0 + i) * W + myw) : make_uint2(0u, seq);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    while (v[i].y != seq) {
                        if (SLEEP) __nanosleep(SLEEP);
                        v[i] = lld<SCOPE>(slab + (size_t)(c0 + i) * W + myw);
                    }
                    g += __uint_as_float(v[i].x);
                }
            }
        }
        __syncthreads();
        if (tid < 2 * G) lst4<SCOPE>(sq + ((size_t)tid * (2 * G) + me) * 2, __float_as_uint(g), seq, 0u, seq);
        long long t2 = clock64();
        // D
        if (tid < 2 * G) {
            uint4 v = lld4<SCOPE>(sq + ((size_t)me * (2 * G) + tid) * 2);
            while (v.y != seq || v.w != seq) {
                if (SLEEP) __nanosleep(SLEEP);
                v = lld4<SCOPE>(sq + ((size_t)me * (2 * G) + tid) * 2);
            }
            keep += __uint_as_float(v.x);
        }
        __syncthreads();
        long long t3 = clock64();
        if (owns) {
            if (SHAREDW) lst<SCOPE>(wl + myw, __float_as_uint(g), seq);
            else
                for (int c = 0; c < G; ++c) lst<SCOPE>(wl + (size_t)c * W + myw, __float_as_uint(g), seq);
        }
        long long t4 = clock64();
        // E
        const uint2* src = SHAREDW ? wl : wl + (size_t)blockIdx.x * W;
        for (int j = tid; 4 * j < W; j += NTH) {
            uint4 lo = lld4<SCOPE>(src + 4 * j), hi = lld4<SCOPE>(src + 4 * j + 2);
            while (lo.y != seq || lo.w != seq || hi.y != seq || hi.w != seq) {
                if (SLEEP) __nanosleep(SLEEP);
                lo = lld4<SCOPE>(src + 4 * j);
                hi = lld4<SCOPE>(src + 4 * j + 2);
            }
            keep += __uint_as_float(lo.x) + __uint_as_float(hi.z);
        }
        __syncthreads();
        long long t5 = clock64();
        acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2; acc[3] += t4 - t3; acc[4] += t5 - t4; acc[5] += t5 - t0;
    }
    if (tid == 0) {
        for (int k = 0; k < 6; ++k) out[me * 6 + k] = acc[k] / rounds;
        sink[me] = keep;
    }
}

template <int SCOPE, int SHAREDW, int SLEEP>
void run(const char* name, uint2* ll, size_t bytes, long long* out, float* sink, int G) {
    cudaMemset(ll, 0, bytes);
    auto k = round_kernel<SCOPE, SHAREDW, SLEEP>;
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int rounds = 200;
    void* args[] = {&ll, &rounds, &out, &sink};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k, dim3(G, 2), dim3(NTH), args, smem, 0);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[6 * 148];
    cudaMemcpy(h, out, sizeof(long long) * 6 * 2 * G, cudaMemcpyDeviceToHost);
    printf("%-44s %s/%s  %.2f us per round | cta0 cycles: flush %lld  C %lld  D-poll %lld  D-store %lld  E %lld  total %lld\n", name,
           cudaGetErrorString(e), cudaGetErrorString(e2), 1e3 * ms / rounds, h[0], h[1], h[2], h[3], h[4], h[5]);
}

int main() {
    const int G = 64;
    const size_t words = (size_t)8 * G * W + (size_t)2 * (2 * G) * (2 * G) * 2;
    uint2* ll;
    cudaMalloc(&ll, words * sizeof(uint2));
    long long* out; float* sink;
    cudaMalloc(&out, sizeof(long long) * 6 * 148);
    cudaMalloc(&sink, sizeof(float) * 148);
    run<0, 0, 0>("volatile(sys), private copies", ll, words * sizeof(uint2), out, sink, G);
    run<1, 0, 0>("relaxed.gpu, private copies", ll, words * sizeof(uint2), out, sink, G);
    run<1, 1, 0>("relaxed.gpu, shared copy", ll, words * sizeof(uint2), out, sink, G);
    run<1, 0, 100>("relaxed.gpu, private copies, nanosleep 100", ll, words * sizeof(uint2), out, sink, G);
    run<1, 1, 100>("relaxed.gpu, shared copy, nanosleep 100", ll, words * sizeof(uint2), out, sink, G);
    run<0, 1, 100>("volatile, shared copy, nanosleep 100", ll, words * sizeof(uint2), out, sink, G);
    return 0;
}
