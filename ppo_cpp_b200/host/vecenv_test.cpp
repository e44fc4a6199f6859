// The reference's one unit test, restated without Catch2 (test/vecenv_test.cpp:13-60): N EnvMock(i+1) in a VecEnv,
// 5 steps with zero actions; every obs column and the reward column must equal [1..N]^T.  CPU only.
#include <chrono>
#include <cstdio>
#include <memory>
#include <thread>

#include "env_mock.hpp"
#include "vec_env.hpp"

static int failures = 0;
#define REQUIRE(cond)                                                        \
    do {                                                                     \
        if (!(cond)) {                                                       \
            std::printf("REQUIRE failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
            ++failures;                                                      \
        }                                                                    \
    } while (0)

static void simulate_steps(const int steps, const int num_threads) {
    std::vector<std::shared_ptr<Env>> envs;
    Mat test_column = Mat::Zero(num_threads, 1);
    for (int i = 0; i < num_threads; ++i) {
        envs.push_back(std::make_shared<EnvMock>(i + 1));
        test_column(i, 0) = static_cast<float>(i + 1);
    }
    VecEnv ve{envs};
    std::this_thread::sleep_for(std::chrono::milliseconds(20));
    for (int s = 0; s < steps; ++s) {
        Mat actions = Mat::Zero(ve.get_num_envs(), ve.get_action_space_size());
        const auto result = ve.step(actions);
        const Mat& obs = result[0];
        const Mat& rew = result[1];
        REQUIRE(obs.rows() == ve.get_num_envs());
        REQUIRE(rew.rows() == ve.get_num_envs());
        REQUIRE(obs.cols() == ve.get_observation_space_size());
        REQUIRE(rew.cols() == 1);
        REQUIRE((rew - test_column).squaredNorm() < 1e-2f);
        for (int i = 0; i < ve.get_observation_space_size(); ++i) REQUIRE((obs.col(i) - test_column).squaredNorm() < 1e-2f);
        REQUIRE(result[2].sum() == 0.f);
    }
    // reset() returns the sub-envs' cached original observations without resetting them (vec_env.hpp:94-106)
    const Mat r = ve.reset();
    for (int i = 0; i < num_threads; ++i) REQUIRE(r(i, 0) == static_cast<float>(i + 1));
    REQUIRE((ve.get_original_rew() - test_column).squaredNorm() < 1e-2f);
}

int main() {
    simulate_steps(5, 1);
    simulate_steps(5, 2);
    simulate_steps(5, 16);
    // done flag of the mock fires at every 300th call of each sub-env
    {
        std::vector<std::shared_ptr<Env>> envs{std::make_shared<EnvMock>(1.0), std::make_shared<EnvMock>(2.0)};
        VecEnv ve{envs};
        int dones = 0;
        for (int s = 0; s < 600; ++s) dones += static_cast<int>(ve.step(Mat::Zero(2, 18))[2].sum());
        REQUIRE(dones == 4);
    }
    // json subset round trip used by the checkpoint sidecar
    {
        nlohmann::json j;
        j["obs_rms"]["count"] = 72001473.000001;
        j["obs_rms"]["mean"] = std::vector<float>{0.5f, -0.25f};
        j["n_steps"] = 65536;
        j["model_filename"] = std::string("a/b.meta.txt");
        nlohmann::json k = nlohmann::json::parse(j.dump());
        REQUIRE(k["obs_rms"]["count"].get<double>() == 72001473.000001);
        REQUIRE(k["obs_rms"]["mean"].get<std::vector<float>>()[1] == -0.25f);
        REQUIRE(k["n_steps"].get<int>() == 65536);
        REQUIRE(k["model_filename"].get<std::string>() == "a/b.meta.txt");
    }
    std::printf(failures ? "FAILED (%d)\n" : "vecenv_test OK\n", failures);
    return failures ? 1 : 0;
}
