// Device-side helpers shared by all kernels: network layout, Philox4x32-10, Box-Muller.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ppo {

// Tensor indices in the flat parameter vector (graph gradient order, GRAPH:23738-24074, then q head).
enum TensorId {
    T_PI_FC0_W, T_PI_FC0_B, T_VF_FC0_W, T_VF_FC0_B, T_PI_FC1_W, T_PI_FC1_B, T_VF_FC1_W, T_VF_FC1_B,
    T_VF_W, T_VF_B, T_PI_W, T_PI_B, T_LOGSTD, T_Q_W, T_Q_B, T_COUNT
};

struct NetDims {
    int O, A, H1, H2;
    int off[T_COUNT + 1];  // off[T_Q_W] = P (trainable), off[T_COUNT] = total
    int P, Pq;
    __host__ void init(int o, int a, int h1, int h2) {
        O = o; A = a; H1 = h1; H2 = h2;
        const int sz[T_COUNT] = {O * H1, H1, O * H1, H1, H1 * H2, H2, H1 * H2, H2, H2, 1, H2 * A, A, A, H2 * A, A};
        off[0] = 0;
        for (int i = 0; i < T_COUNT; ++i) off[i + 1] = off[i] + sz[i];
        P = off[T_Q_W];
        Pq = off[T_COUNT];
    }
};

// columns appended to every per-CTA partial gradient slab: loss sums
enum { L_PG = 0, L_VF = 1, L_ENT = 2, L_KL = 3, L_CLIP = 4, L_PAD = 8 };

// 0.5*log(2*pi), 0.5*log(2*pi*e) as the fp32 constants baked in the graph (GRAPH:6103-6672, 10021-10180)
#define PPO_HALF_LOG_2PI 0.9189385175704956f
#define PPO_HALF_LOG_2PIE 1.4189385175704956f

#define PPO_TAG_ACTION 0x50504F32u    // "PPO2": action noise stream
#define PPO_TAG_ENVNOISE 0x454E5631u  // "ENV1": synthetic env process noise
#define PPO_TAG_ENVRESET 0x52535431u  // "RST1": synthetic env reset state

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// Philox4x32-10 (Salmon et al., SC'11) — the generator TF's RandomStandardNormal uses.
__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// TF's Uint32ToFloat: 23 mantissa bits -> [0,1)
__device__ __forceinline__ float u32_to_unit(uint32_t x) { return __uint_as_float((x & 0x7fffffu) | 0x3f800000u) - 1.0f; }

// Box-Muller on two 32-bit words, TF-style (u1 clamped to 1e-7); full-precision logf/sinf/cosf.
__device__ __forceinline__ void box_muller(uint32_t w0, uint32_t w1, float& n0, float& n1) {
    float u1 = fmaxf(u32_to_unit(w0), 1.0e-7f);
    const float u2 = u32_to_unit(w1);
    const float r = sqrtf(-2.0f * logf(u1));
    const float th = 6.2831853071795864769f * u2;
    float s, c;
    sincosf(th, &s, &c);
    n0 = r * s;
    n1 = r * c;
}

// N(0,1) block `blk` (4 values) of stream (seed, a, b, tag)
__device__ __forceinline__ void normal4(uint64_t seed, uint32_t a, uint32_t b, uint32_t blk, uint32_t tag, float out[4]) {
    const uint4 w = philox4x32_10(make_uint4(a, b, blk, tag), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    box_muller(w.x, w.y, out[0], out[1]);
    box_muller(w.z, w.w, out[2], out[3]);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace ppo
