// std::srand / std::rand / std::random_shuffle ON THE DEVICE, bit-exact with glibc + libstdc++ (ppo2/ppo2.hpp:274-288).
//
// Why: the reference shuffles the n_batch indices once per epoch with rand() % (i+1) swaps — a strictly sequential
// chain (~3.5 ns per element on a host core).  At C3 that is 0.9 ms per epoch = 9 ms per update, as long as the whole
// GPU update; with the batch sharded over 8 GPUs the (global) permutation is 8x longer and every rank would wait 70 ms
// per update for its host.  Both parts of the chain parallelise:
//   (1) the rand() stream.  glibc TYPE_3 is the linear recurrence x[k] = x[k-3] + x[k-31] (mod 2^32), output x[k] >> 1
//       (stdlib/random_r.c).  x^k mod (x^31 - x^28 - 1) over Z/2^32 jumps the 31-word window k steps ahead, so thread g
//       computes the window at position g*L from precomputed powers (binary exponentiation, 31x31 polynomial products)
//       and then runs the recurrence for L outputs out of registers.
//   (2) the swap chain  for i in 1..n-1: swap(a[i], a[j_i]), j_i = rand() % (i+1)  applied to the identity.  Position i
//       is untouched before step i, so step i moves the fresh value i to position j_i and the old value of position j_i
//       to position i.  Hence the final value of position p is
//           max S_p                      if S_p = { s > p : j_s = p } is not empty (the last fresh value dropped there),
//           val(j_p, p)                  otherwise — what step p fetched from position j_p, where
//           val(q, t) = max { s in S_q : s < t }  if that set is not empty, else val(j_q, q)   (val(q, .) = q if j_q = q).
//       Every hop moves to a uniformly smaller position, so chains are O(log n) long; S_q are the buckets of a
//       counting sort of the steps by target (expected size ln(n/q)).
//   (3) epochs compound (the reference shuffles the already shuffled array): perm_e = perm_{e-1} o sigma_e, a gather.
// All epochs of an update are produced in one go at the start of the update; the host does nothing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ppo {
namespace shuf {

constexpr int DEG = 31;
constexpr int L = 256;          // outputs per thread of the stream kernel
constexpr int TAB_BITS = 40;    // jump tables cover 2^40 steps

// c = a * b mod (x^31 - x^28 - 1), coefficients mod 2^32
__host__ __device__ inline void polymul(const uint32_t* a, const uint32_t* b, uint32_t* c) {
    uint32_t e[2 * DEG - 1];
#pragma unroll
    for (int i = 0; i < 2 * DEG - 1; ++i) e[i] = 0u;
#pragma unroll
    for (int i = 0; i < DEG; ++i)
#pragma unroll
        for (int j = 0; j < DEG; ++j) e[i + j] += a[i] * b[j];
#pragma unroll
    for (int d = 2 * DEG - 2; d >= DEG; --d) {  // x^d = x^(d-3) + x^(d-31)
        e[d - 3] += e[d];
        e[d - DEG] += e[d];
    }
#pragma unroll
    for (int i = 0; i < DEG; ++i) c[i] = e[i];
}

struct Tables {
    uint32_t pow1[TAB_BITS][DEG];  // x^(2^b)
    uint32_t powL[TAB_BITS][DEG];  // x^(L * 2^b)
};

inline void build_tables(Tables& t) {
    uint32_t x[DEG] = {0};
    x[1] = 1u;
    for (int i = 0; i < DEG; ++i) t.pow1[0][i] = x[i];
    for (int b = 1; b < TAB_BITS; ++b) polymul(t.pow1[b - 1], t.pow1[b - 1], t.pow1[b]);
    int lb = 0;
    while ((1 << lb) < L) ++lb;
    for (int b = 0; b < TAB_BITS; ++b) {
        if (b + lb < TAB_BITS) {
            for (int i = 0; i < DEG; ++i) t.powL[b][i] = t.pow1[b + lb][i];
        } else {
            polymul(t.powL[b - 1], t.powL[b - 1], t.powL[b]);
        }
    }
}

// x^k from a table of x^(unit * 2^b): product over the set bits of k
__device__ __forceinline__ void poly_pow(const uint32_t (*tab)[DEG], unsigned long long k, uint32_t* c) {
#pragma unroll
    for (int i = 0; i < DEG; ++i) c[i] = (i == 0) ? 1u : 0u;
    for (int b = 0; k != 0ull && b < TAB_BITS; ++b, k >>= 1) {
        if (k & 1ull) {
            uint32_t t[DEG], r[DEG];
#pragma unroll
            for (int i = 0; i < DEG; ++i) t[i] = __ldg(&tab[b][i]);
            polymul(c, t, r);
#pragma unroll
            for (int i = 0; i < DEG; ++i) c[i] = r[i];
        }
    }
}

// S[0..60]: the window (oldest first: x[k-31] .. x[k-1]) followed by the next 30 terms
__device__ __forceinline__ void extend_window(const uint32_t* win, uint32_t* S) {
    for (int i = 0; i < DEG; ++i) S[i] = win[i];
    for (int i = DEG; i < 2 * DEG - 1; ++i) S[i] = S[i - 3] + S[i - DEG];
}

// j[e][i] = rand() % (i + 1) for i = 1..n-1 of every epoch e (stream position k = e*(n-1) + i-1); j[e][0] = 0.
// Epoch subsets (multi-GPU: rank r builds the permutations of epochs r, r + R, ... and the ranks exchange them): blockIdx.y
// selects epoch e = e0 + blockIdx.y * estride and the threads cover the L-blocks of the stream that overlap that epoch
// (draws of a neighbouring epoch inside the first / last block are written too: they are correct, just not needed).
__global__ void __launch_bounds__(128) shuffle_draw_kernel(const uint32_t* __restrict__ win, const Tables* __restrict__ tab, int n,
                                                           int epochs, int* __restrict__ jbuf, int e0, int estride) {
    __shared__ uint32_t S[2 * DEG - 1];
    if (threadIdx.x == 0) extend_window(win, S);
    __syncthreads();
    const long long total = (long long)epochs * (n - 1);
    const int my_e = e0 + (int)blockIdx.y * estride;
    const long long g_lo = ((long long)my_e * (n - 1)) / L;
    const long long g = g_lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long k0 = g * L;
    if (blockIdx.x == 0 && threadIdx.x == 0) jbuf[(size_t)my_e * n] = 0;
    if (k0 >= total || k0 >= (long long)(my_e + 1) * (n - 1)) return;
    uint32_t c[DEG], w[DEG];
    poly_pow(tab->powL, (unsigned long long)g, c);
#pragma unroll
    for (int m = 0; m < DEG; ++m) {
        uint32_t acc = 0u;
#pragma unroll
        for (int t = 0; t < DEG; ++t) acc += c[t] * S[t + m];
        w[m] = acc;
    }
    int e = (int)(k0 / (n - 1));
    int i = 1 + (int)(k0 - (long long)e * (n - 1));
    const int count = (int)((total - k0 < (long long)L) ? (total - k0) : (long long)L);
    int done = 0;
    while (done < count) {
#pragma unroll
        for (int t = 0; t < DEG; ++t) {
            w[t] += w[(t + 28) % DEG];  // x[k] = x[k-31] + x[k-3]; slot t held x[k-31]
            if (done < count) {
                const uint32_t r = w[t] >> 1;
                jbuf[(size_t)e * n + i] = (int)(r % (uint32_t)(i + 1));
                ++done;
                if (++i == n) {
                    i = 1;
                    ++e;
                }
            }
        }
    }
}

// the window after `total` draws (stored oldest first), for the next update
__global__ void shuffle_advance_kernel(uint32_t* __restrict__ win, const Tables* __restrict__ tab, unsigned long long total) {
    __shared__ uint32_t S[2 * DEG - 1];
    __shared__ uint32_t cc[DEG];
    if (threadIdx.x == 0) {
        extend_window(win, S);
        uint32_t c[DEG];
        poly_pow(tab->pow1, total, c);
        for (int i = 0; i < DEG; ++i) cc[i] = c[i];
    }
    __syncthreads();
    if (threadIdx.x < DEG) {  // S was copied out of win before the barrier: no hazard with the stores below
        uint32_t acc = 0u;
        for (int t = 0; t < DEG; ++t) acc += cc[t] * S[t + threadIdx.x];
        win[threadIdx.x] = acc;
    }
}

// cnt[e][p] = |S_p| = #{ s > p : j_s = p }
__global__ void shuffle_count_kernel(const int* __restrict__ jbuf, int n, int* __restrict__ cnt, int e0, int estride) {
    const int e = e0 + blockIdx.y * estride;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < 1 || s >= n) return;
    const int j = jbuf[(size_t)e * n + s];
    if (j < s) atomicAdd(cnt + (size_t)e * (n + 1) + j, 1);
}

// exclusive scan of cnt[e][0..n] -> off[e][0..n] (n + 1 entries) and cur = off, in three coalesced passes:
// per-block totals (1024 elements per block), scan of the totals (one block per epoch), local rescan + block offset
constexpr int SCAN_TILE = 1024;

__device__ __forceinline__ int block_exclusive_scan_1024(int v, int* total) {
    __shared__ int warp_tot[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += u;
    }
    __syncthreads();
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int t = warp_tot[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, t, d);
            if (lane >= d) t += u;
        }
        warp_tot[lane] = t;
    }
    __syncthreads();
    if (total) *total = warp_tot[31];
    return incl - v + (wid ? warp_tot[wid - 1] : 0);
}

__global__ void __launch_bounds__(SCAN_TILE) shuffle_scan_totals_kernel(const int* __restrict__ cnt, int n, int nb, int* __restrict__ btot, int e0,
                                                                       int estride) {
    const int e = e0 + blockIdx.y * estride, i = blockIdx.x * SCAN_TILE + threadIdx.x;
    const int v = (i <= n) ? cnt[(size_t)e * (n + 1) + i] : 0;
    int total;
    block_exclusive_scan_1024(v, &total);
    if (threadIdx.x == 0) btot[(size_t)e * nb + blockIdx.x] = total;
}
__global__ void __launch_bounds__(SCAN_TILE) shuffle_scan_blocks_kernel(int nb, int* __restrict__ btot, int e0, int estride) {  // nb <= 1024 * 1024
    __shared__ int carry_s;
    const int e = e0 + blockIdx.x * estride;
    int* b = btot + (size_t)e * nb;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += SCAN_TILE) {
        const int i = base + threadIdx.x;
        const int v = (i < nb) ? b[i] : 0;
        int total;
        const int ex = block_exclusive_scan_1024(v, &total);
        const int carry = carry_s;
        if (i < nb) b[i] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(SCAN_TILE) shuffle_scan_final_kernel(const int* __restrict__ cnt, int n, int nb, const int* __restrict__ btot,
                                                                      int* __restrict__ off, int* __restrict__ cur, int e0, int estride) {
    const int e = e0 + blockIdx.y * estride, i = blockIdx.x * SCAN_TILE + threadIdx.x;
    const int v = (i <= n) ? cnt[(size_t)e * (n + 1) + i] : 0;
    const int ex = block_exclusive_scan_1024(v, nullptr) + btot[(size_t)e * nb + blockIdx.x];
    if (i <= n) {
        off[(size_t)e * (n + 1) + i] = ex;
        cur[(size_t)e * (n + 1) + i] = ex;
    }
}

// list[e][off[j_s] ...] <- s   (order inside a bucket is arbitrary: the resolver takes maxima)
__global__ void shuffle_scatter_kernel(const int* __restrict__ jbuf, int n, int* __restrict__ cur, int* __restrict__ list, int e0, int estride) {
    const int e = e0 + blockIdx.y * estride;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < 1 || s >= n) return;
    const int j = jbuf[(size_t)e * n + s];
    if (j < s) {
        const int pos = atomicAdd(cur + (size_t)e * (n + 1) + j, 1);
        list[(size_t)e * n + pos] = s;
    }
}

// sigma[e][p] = value at position p after the swap chain applied to the identity
__global__ void shuffle_resolve_kernel(const int* __restrict__ jbuf, const int* __restrict__ off, const int* __restrict__ list, int n,
                                       int* __restrict__ sigma, int e0, int estride) {
    const int e = e0 + blockIdx.y * estride;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int* j = jbuf + (size_t)e * n;
    const int* o = off + (size_t)e * (n + 1);
    const int* l = list + (size_t)e * n;
    int q = p, t = n, val;
    while (true) {
        int best = -1;
        const int a = o[q], b = o[q + 1];
        for (int x = a; x < b; ++x) {
            const int s = l[x];
            if (s < t && s > best) best = s;
        }
        if (best >= 0) {
            val = best;
            break;
        }
        const int jq = j[q];  // j[0] = 0
        if (jq == q) {
            val = q;
            break;
        }
        t = q;
        q = jq;
    }
    sigma[(size_t)e * n + p] = val;
}

// perm_e = perm_{e-1} o sigma_e (perm_{-1} = identity); gather_e[perm_e[i]] = physical row of semantic row i
// (Eigen `perm * buf` writes out[perm[i]] = in[i], ppo2.hpp:291-296; physical layout: see build_gather_kernel)
__global__ void shuffle_compose_kernel(const int* __restrict__ prev, const int* __restrict__ sigma, int n, int T, int Nl, int* __restrict__ perm,
                                       int* __restrict__ gather) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int sg = sigma[i];
    const int pv = prev ? prev[sg] : sg;
    perm[i] = pv;
    const int env_g = i / T, t = i - env_g * T;
    const int r = env_g / Nl, el = env_g - r * Nl;
    gather[pv] = r * T * Nl + t * Nl + el;
}

// Multi-GPU: sigma of the epochs this rank resolved -> the same place in every peer's sigma array (NVLink P2P stores;
// peer[r] = this rank's mapping of rank r's array, peer[my rank] is skipped).  16-byte accesses where n allows.
struct SigmaPeers {
    int* p[8];
    int rank, world;
};
__global__ void shuffle_publish_kernel(SigmaPeers peers, const int* __restrict__ sigma, int n, int e0, int estride) {
    const int e = e0 + blockIdx.y * estride;
    const size_t base = (size_t)e * n;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int v = sigma[base + i];
        for (int r = 0; r < peers.world; ++r)
            if (r != peers.rank) peers.p[r][base + i] = v;
    }
}

}  // namespace shuf
}  // namespace ppo
