// Synchronisation primitives of the persistent kernels: a monotonic-counter grid barrier (all CTAs of a cooperative
// launch) and the peer-memory mailbox used for the cross-GPU exchanges (VecNormalize moments, gradients) over
// NVLink P2P stores — no NCCL call on the per-step / per-minibatch path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ppo {

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Grid barrier on a zero-initialised counter that only grows: generation g (1, 2, ...) completes when the counter
// reaches g * nblocks.  All CTAs must be co-resident (cooperative launch).  One thread per CTA adds 1 after a
// __threadfence and polls with plain volatile loads (an acquire load per poll iteration carries a fence each time:
// measured 2x slower; per-CTA flag words polled by a warp: 4x slower; a two-level version — group counters of 16 CTAs,
// the last arriver of a group arriving on a top counter — measured 1.4x slower: the extra dependent hop costs more than
// the same-address serialisation it avoids), then fences once.  arrive/wait are split so that
// independent work can sit between them; every thread of the CTA must call both (they contain __syncthreads).
// `gen` = generations completed so far; it lives in device memory between launches so that a launch can be replayed
// from a CUDA graph (wrap-around safe: only differences are compared).
struct GridBarrier {
    unsigned* ctr;
    unsigned nblocks;
    unsigned gen;
    __device__ __forceinline__ void arrive() {
        __syncthreads();  // all writes of this CTA issued
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(ctr, 1u);
        }
    }
    __device__ __forceinline__ void wait() {
        ++gen;
        if (threadIdx.x == 0) {
            const unsigned target = gen * nblocks;
            while ((int)(*reinterpret_cast<volatile unsigned*>(ctr) - target) < 0) {
            }
            __threadfence();
        }
        __syncthreads();
    }
    __device__ __forceinline__ void sync() {
        if (nblocks == 1) {  // a single CTA (n_envs <= one tile, the reference's own C1 / C2 shapes): no L2 round trips
            __syncthreads();
            ++gen;
            if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned*>(ctr) = gen;  // keeps ctr == gen * nblocks for later launches
            return;
        }
        arrive();
        wait();
    }
};

// Cross-GPU mailbox.  Every rank owns one device allocation, IPC-mapped into every peer of the node:
//   flags   [channels][8] uint32          monotonic sequence numbers, one word per (channel, source rank)
//   moments [2][world][512 B]             payload slots of the VecNormalize exchange, double-buffered by sequence parity
//   grads   [2][world][PS floats]         payload slots of the gradient exchange
// (a PeerMailbox value describes one payload region: data_off / slot_bytes)
// A sender writes its payload into slot [seq & 1][my_rank] of EVERY rank (its own included) with plain stores
// (NVLink P2P for the peers), fences system-wide and release-stores seq into flags[channel][my_rank] of that rank.
// A receiver spins on its own flag words until all sources reached seq, then reads the slots in rank order — every
// rank adds the same numbers in the same order, so replicated state stays bit-identical.  Two slots suffice: a sender
// can be at most one sequence number ahead of the slowest receiver (it needs that receiver's next flag to go on).
constexpr int PPO_MAX_WORLD = 8;
constexpr int PPO_MBOX_CHANNELS = 1024;  // channel = cooperating CTA index (gradient slices); the last one carries the moments
constexpr int PPO_MBOX_MOMENT_CHANNEL = PPO_MBOX_CHANNELS - 1;
constexpr int PPO_MBOX_DONE_CHANNEL = PPO_MBOX_CHANNELS - 2;  // end-of-rollout "my rows are in your buffers"
constexpr int PPO_MBOX_SHUF_CHANNEL = PPO_MBOX_CHANNELS - 3;  // "the permutations of my epochs are in your array"
constexpr size_t PPO_MBOX_MOMENT_SLOT = 2048;  // bytes: 2*(D+1) doubles as 2 LL words each, D <= 32
constexpr size_t PPO_MBOX_FLAG_BYTES = (size_t)PPO_MBOX_CHANNELS * PPO_MAX_WORLD * sizeof(unsigned);

// Low-latency ("LL") payload words: 4 bytes of data + the 4-byte sequence number in ONE 8-byte store.  An aligned
// 8-byte store is a single transaction on NVLink, so data and flag become visible together: no fence and no separate
// flag round trip (a __threadfence_system + flag exchange measured ~10 us per exchange between two B200s; this is one
// NVLink one-way latency).  The receiver polls the payload words themselves.
__device__ __forceinline__ void ll_store(uint2* p, unsigned data, unsigned seq) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(data), "r"(seq) : "memory");
}
__device__ __forceinline__ uint2 ll_load(const uint2* p) {
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}

struct PeerMailbox {
    unsigned char* base[PPO_MAX_WORLD];  // this rank's mapping of rank r's arena
    int rank, world;
    size_t data_off, slot_bytes;
    unsigned* err;  // set to 1 when a wait timed out (a peer died): the caller reports PPO_ERR_COMM
    __device__ __forceinline__ unsigned* flag(int r, int channel, int src) const {
        return reinterpret_cast<unsigned*>(base[r]) + (size_t)channel * PPO_MAX_WORLD + src;
    }
    __device__ __forceinline__ unsigned char* slot(int r, unsigned seq, int src) const {
        return base[r] + data_off + ((size_t)(seq & 1u) * world + src) * slot_bytes;
    }
    __device__ __forceinline__ uint2* ll_slot(int r, unsigned seq, int src) const { return reinterpret_cast<uint2*>(slot(r, seq, src)); }
    // spin until word i of source `src` carries `seq` (bounded: ~4 s, then the error flag is raised and 0 returned)
    __device__ __forceinline__ unsigned ll_wait(unsigned seq, int src, size_t i) const {
        const uint2* p = ll_slot(rank, seq, src) + i;
        uint2 v = ll_load(p);
        if (v.y == seq) return v.x;
        if (*reinterpret_cast<volatile unsigned*>(err)) return 0u;
        const unsigned long long t0 = globaltimer_ns();
        unsigned spins = 0;
        while (true) {
            v = ll_load(p);
            if (v.y == seq) return v.x;
            if (((++spins) & 0x3ffu) == 0u && globaltimer_ns() - t0 > 4000000000ull) {
                *err = 1u;
                return 0u;
            }
        }
    }
    // word i of source ranks [src0, src0 + 4) in one pass: the loads are in flight together (polling the sources one after
    // the other costs one local L2 round trip per rank: measured +6.5 us per exchange at 8 ranks); stale words re-poll.
    // out[k] = payload of rank src0 + k; bounded like ll_wait.
    __device__ __forceinline__ void ll_wait4(unsigned seq, size_t i, int src0, unsigned* out) const {
        uint2 v[4];
        unsigned long long t0 = 0;
        unsigned spins = 0;
        while (true) {
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = (src0 + k < world) ? ll_load(ll_slot(rank, seq, src0 + k) + i) : make_uint2(0u, seq);
#pragma unroll
            for (int k = 0; k < 4; ++k) ok = ok && v[k].y == seq;
            if (ok) break;
            if (((++spins) & 0x3ffu) == 0u) {
                if (t0 == 0) t0 = globaltimer_ns();
                if (*reinterpret_cast<volatile unsigned*>(err) || globaltimer_ns() - t0 > 4000000000ull) {
                    *err = 1u;
                    break;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) out[k] = v[k].x;
    }
    // the same for a pair of adjacent words (i even: one 16-byte load per source), e.g. the two halves of a double
    __device__ __forceinline__ void ll_wait4_pair(unsigned seq, size_t i, int src0, unsigned long long* out) const {
        uint4 v[4];
        unsigned long long t0 = 0;
        unsigned spins = 0;
        while (true) {
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[k] = make_uint4(0u, seq, 0u, seq);
                if (src0 + k < world) {
                    const uint2* p = ll_slot(rank, seq, src0 + k) + i;
                    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[k].x), "=r"(v[k].y), "=r"(v[k].z), "=r"(v[k].w) : "l"(p) : "memory");
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) ok = ok && v[k].y == seq && v[k].w == seq;
            if (ok) break;
            if (((++spins) & 0x3ffu) == 0u) {
                if (t0 == 0) t0 = globaltimer_ns();
                if (*reinterpret_cast<volatile unsigned*>(err) || globaltimer_ns() - t0 > 4000000000ull) {
                    *err = 1u;
                    break;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) out[k] = ((unsigned long long)v[k].z << 32) | (unsigned long long)v[k].x;
    }
    // fenced flag protocol (used once per rollout for "all my rows are in your buffers")
    __device__ __forceinline__ void signal_all(int channel, unsigned seq) const {
        __threadfence_system();
        for (int r = 0; r < world; ++r) st_release_sys(flag(r, channel, rank), seq);
    }
    __device__ __forceinline__ void wait_all(int channel, unsigned seq) const {
        if (*reinterpret_cast<volatile unsigned*>(err)) return;  // a previous wait already timed out: do not stall again
        const unsigned long long t0 = globaltimer_ns();
        for (int src = 0; src < world; ++src) {
            const unsigned* f = flag(rank, channel, src);
            unsigned spins = 0;
            while ((int)(*reinterpret_cast<const volatile unsigned*>(f) - seq) < 0) {
                if (((++spins) & 0x3ffu) == 0u && globaltimer_ns() - t0 > 4000000000ull) {
                    *err = 1u;
                    return;
                }
            }
        }
        __threadfence_system();
    }
};

}  // namespace ppo
