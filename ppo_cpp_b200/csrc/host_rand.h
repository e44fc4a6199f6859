// std::srand / std::rand / std::random_shuffle as the reference's PPO2::learn uses them
// (ppo2/ppo2.hpp:288; seed from ppo2.cpp:159-162), restated so the permutation stream is (a) bit-exact
// with glibc 2.x + libstdc++ and (b) owned by the core instead of process-global state.
//   glibc rand(): TYPE_3 additive feedback generator x[i] = x[i-3] + x[i-31] (mod 2^32), output x[i] >> 1,
//   seeded by the minimal-standard LCG (16807 mod 2^31-1) and 310 discarded outputs (stdlib/random_r.c).
//   libstdc++ random_shuffle(first,last): for i in 1..n-1: swap(a[i], a[rand() % (i+1)]) (bits/stl_algo.h:4581).
#pragma once
#include <cstdint>

namespace ppo {

class GlibcRand {
public:
    explicit GlibcRand(unsigned seed = 1) { srand(seed); }
    void srand(unsigned seed) {
        int32_t word = seed ? static_cast<int32_t>(seed) : 1;
        ring_[0] = static_cast<uint32_t>(word);
        for (int i = 1; i < 31; ++i) {
            const long hi = word / 127773, lo = word % 127773;
            long w = 16807 * lo - 2836 * hi;
            if (w < 0) w += 2147483647;
            word = static_cast<int32_t>(w);
            ring_[i] = static_cast<uint32_t>(word);
        }
        f_ = 3;
        r_ = 0;
        for (int i = 0; i < 310; ++i) (void)rand();
    }
    inline int rand() {
        const uint32_t v = (ring_[f_] += ring_[r_]);
        if (++f_ >= 31) f_ = 0;
        if (++r_ >= 31) r_ = 0;
        return static_cast<int>(v >> 1);
    }
    void random_shuffle(int* a, int n) {
        for (int i = 1; i < n; ++i) {
            const int j = rand() % (i + 1);
            if (i != j) {
                const int t = a[i];
                a[i] = a[j];
                a[j] = t;
            }
        }
    }

    // the last 31 raw words, oldest first (x[k-31] .. x[k-1]): the whole generator state
    void get_window(uint32_t out[31]) const {
        for (int t = 0; t < 31; ++t) out[t] = ring_[(f_ + t) % 31];
    }
    void set_window(const uint32_t in[31]) {
        for (int t = 0; t < 31; ++t) ring_[t] = in[t];
        f_ = 0;   // slot f holds x[k-31], slot r = f - 3 holds x[k-3]
        r_ = 28;
    }

private:
    uint32_t ring_[31];
    int f_ = 3, r_ = 0;
};

}  // namespace ppo
