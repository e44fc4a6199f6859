// Host-side synthetic 18-dim env with the hexapod's interface shape: used where the reference uses
// the DART hexapod (not installable): s' = 0.9 s + 0.1 clamp(a,-1,1) + 0.01 xi, reward = s'[0]-s[0] (the hexapod's
// reward is the x displacement per step, env/hexapod_env.hpp:159-166), 334-step episodes
// (hexapod_env.hpp:41,170), reset to reset_noise_scale*U(-1,1) (hexapod_closed_loop_env.hpp:83).
// Same Philox streams as the device-resident env in csrc/kernels_misc.cuh, so both produce the same trajectory.
#ifndef PPO_B200_ENV_SYNTHETIC_HPP
#define PPO_B200_ENV_SYNTHETIC_HPP

#include <cmath>
#include <cstdint>
#include <cstring>

#include "env.hpp"

class SyntheticEnv : public Env {
public:
    explicit SyntheticEnv(uint64_t seed = 0x1234, uint32_t env_id = 0, int dim = 18) : seed{seed}, env_id{env_id}, dim{dim}, state(1, dim) {
        hard_reset();
    }
    std::string get_action_space() override { return Env::SPACE_CONTINOUS(); }
    std::string get_observation_space() override { return Env::SPACE_CONTINOUS(); }
    int get_action_space_size() override { return dim; }
    int get_observation_space_size() override { return dim; }
    Mat reset() override {
        hard_reset();
        return state;
    }
    std::vector<Mat> step(const Mat& actions) override {
        float xi[32];
        normals(t_env, 0x454E5631u, xi);
        const float s0 = state(0, 0);
        for (int k = 0; k < dim; ++k) {
            float a = actions(0, k);
            a = a < -1.f ? -1.f : (a > 1.f ? 1.f : a);
            state(0, k) = (0.9f * state(0, k) + 0.1f * a) + 0.01f * xi[k];
        }
        last_rew = state(0, 0) - s0;
        ++t_env;
        Mat done = Mat::Zero(1, 1);
        if (t_env % 334u == 0u) {
            done(0, 0) = 1.f;
            soft_reset();
        }
        return {state, Mat::Constant(1, 1, last_rew), done};
    }
    void render() override {}
    float get_time() override { return 0.015f * static_cast<float>(t_env % 334u); }
    Mat get_original_obs() override { return state; }
    Mat get_original_rew() override { return Mat::Constant(1, 1, last_rew); }
    void serialize(nlohmann::json& json) override { json["reset_noise_scale"] = 0.1; }
    void deserialize(nlohmann::json&) override {}

private:
    static uint32_t mulhi(uint32_t a, uint32_t b) { return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32); }
    void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) const {
        uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = mulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0, hi1 = mulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
    static float unit(uint32_t x) {
        const uint32_t bits = (x & 0x7fffffu) | 0x3f800000u;
        float f;
        std::memcpy(&f, &bits, 4);
        return f - 1.0f;
    }
    void normals(uint32_t counter, uint32_t tag, float* out) const {
        for (int blk = 0; blk * 4 < dim; ++blk) {
            uint32_t w[4];
            philox(env_id, counter, static_cast<uint32_t>(blk), tag, w);
            for (int h = 0; h < 2; ++h) {
                float u1 = unit(w[2 * h]);
                if (u1 < 1.0e-7f) u1 = 1.0e-7f;
                const float r = std::sqrt(-2.0f * std::log(u1)), th = 6.2831853071795864769f * unit(w[2 * h + 1]);
                if (blk * 4 + 2 * h < dim) out[blk * 4 + 2 * h] = r * std::sin(th);
                if (blk * 4 + 2 * h + 1 < dim) out[blk * 4 + 2 * h + 1] = r * std::cos(th);
            }
        }
    }
    void soft_reset() {
        for (int blk = 0; blk * 4 < dim; ++blk) {
            uint32_t w[4];
            philox(env_id, resets, static_cast<uint32_t>(blk), 0x52535431u, w);
            for (int k = 0; k < 4 && blk * 4 + k < dim; ++k) state(0, blk * 4 + k) = 0.1f * (2.0f * unit(w[k]) - 1.0f);
        }
        ++resets;
    }
    void hard_reset() {
        resets = 0;
        t_env = env_id % 334u;
        soft_reset();
    }

    uint64_t seed;
    uint32_t env_id;
    int dim;
    Mat state;
    uint32_t t_env = 0, resets = 0;
    float last_rew = 0.f;
};

#endif
