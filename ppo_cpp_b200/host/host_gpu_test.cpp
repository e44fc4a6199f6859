// GPU test of the reference-named C++ classes: EnvNormalize + VecEnv + PPO2::learn / save / load / eval on the
// B200 core.  argv[1] = graph file (.meta.txt), argv[2] = scratch directory.
#include <cmath>
#include <cstdio>
#include <fstream>
#include <memory>

#include "env_mock.hpp"
#include "env_normalize.hpp"
#include "env_synthetic.hpp"
#include "ppo2.hpp"
#include "vec_env.hpp"

static int failures = 0;
#define REQUIRE(cond)                                                        \
    do {                                                                     \
        if (!(cond)) {                                                       \
            std::printf("REQUIRE failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
            ++failures;                                                      \
        }                                                                    \
    } while (0)

int main(int argc, char** argv) {
    if (argc < 3) {
        std::printf("usage: host_gpu_test <graph.meta.txt> <scratch dir>\n");
        return 2;
    }
    const std::string graph = argv[1], dir = argv[2];
    // --- config C1 shape: EnvMock x4 in a VecEnv, the reference's graph, 2 updates
    {
        std::vector<std::shared_ptr<Env>> envs;
        for (int i = 0; i < 4; ++i) envs.push_back(std::make_shared<SyntheticEnv>(0x1234, i));
        EnvNormalize env{std::make_unique<VecEnv>(envs), true};
        PPO2 algo{graph, env, .99f, 256, 0.f, 3.9e-4f, .5f, .5f, .95f, 32, 4, 0.161f, -1.f, ""};
        algo.set_seed(42, 7);
        float p0[442], p1[442];
        REQUIRE(ppo_core_get_tensor(algo.core().get(), "params", p0, 442) == PPO_OK);
        algo.learn(2 * 4 * 256, 1, dir + "/ckpt.pkl");
        REQUIRE(algo.last_fps > 0);
        for (int i = 0; i < 5; ++i) REQUIRE(std::isfinite(algo.last_losses[i]));
        REQUIRE(std::fabs(algo.last_losses[2] - 25.54f) < 0.2f);  // entropy of 18 unit Gaussians = 25.54
        REQUIRE(ppo_core_get_tensor(algo.core().get(), "params", p1, 442) == PPO_OK);
        double moved = 0;
        for (int i = 0; i < 334; ++i) moved += std::fabs(p1[i] - p0[i]);
        REQUIRE(moved > 1e-4);
        for (int i = 334; i < 442; ++i) REQUIRE(p1[i] == p0[i]);  // q head untouched
        // checkpoint sidecar carries the reference's keys
        std::ifstream in(dir + "/ckpt.pkl.0.json");
        REQUIRE(in.is_open());
        std::stringstream ss;
        ss << in.rdbuf();
        nlohmann::json j = nlohmann::json::parse(ss.str());
        REQUIRE(j["n_steps"].get<int>() == 256 && j["nminibatches"].get<int>() == 32 && j["n_envs"].get<int>() == 4);
        REQUIRE(j["obs_rms"]["mean"].get<std::vector<float>>().size() == 18);
        REQUIRE(j["obs_rms"]["count"].get<double>() > 4 * 256 * 2);
        REQUIRE(j["action_space"].get<std::string>() == "continous");
        // load into a fresh model (playback configuration: training = false) and compare the deterministic action
        Mat probe = Mat::Constant(1, 18, 0.25f);
        const Mat want = algo.eval(probe);
        EnvNormalize env2{std::make_unique<SyntheticEnv>(0x1234, 0), false};
        PPO2 algo2{graph, env2, .99f, 256, 0.f, 3.9e-4f, .5f, .5f, .95f, 32, 4, 0.161f, -1.f, ""};
        algo2.load(dir + "/ckpt.pkl.0");
        const Mat got = algo2.eval(probe);
        REQUIRE((got - want).squaredNorm() == 0.f);
        nlohmann::json j2;
        env2.serialize(j2);
        REQUIRE(j2["obs_rms"]["count"].get<double>() == j["obs_rms"]["count"].get<double>());
        // playback step through EnvNormalize::step with frozen statistics
        Mat obs = env2.reset();
        const auto out = env2.step(algo2.eval(obs));
        REQUIRE(out[0].rows() == 1 && out[0].cols() == 18 && std::fabs(out[0](0, 0)) <= 10.f);
        nlohmann::json j3;
        env2.serialize(j3);
        REQUIRE(j3["obs_rms"]["count"].get<double>() == j["obs_rms"]["count"].get<double>());
    }
    // --- single mock env, graph-less [64,64]
    {
        EnvNormalize env{std::make_unique<EnvMock>(1.0), true};
        PPO2 algo{"", env, .99f, 128, 0.f, 1e-3f, .5f, .5f, .95f, 4, 2, 0.2f, -1.f, ""};
        algo.reset_without_graph(64, 64, 3);
        algo.learn(256);
        REQUIRE(algo.last_fps > 0 && std::isfinite(algo.last_losses[0]));
        // graph-less checkpoint round trip from a FRESH model that knows neither a graph file nor the hidden sizes
        // (what `ppo_cpp -p ckpt` does): the .json sidecar carries the shape
        algo.save(dir + "/graphless.pkl", 0);
        Mat probe = Mat::Constant(1, 18, -0.5f);
        const Mat want = algo.eval(probe);
        EnvNormalize env2{std::make_unique<EnvMock>(1.0), false};
        PPO2 algo2{"", env2, .99f, 128, 0.f, 1e-3f, .5f, .5f, .95f, 4, 2, 0.2f, -1.f, ""};
        algo2.load(dir + "/graphless.pkl.0");
        const Mat got = algo2.eval(probe);
        REQUIRE((got - want).squaredNorm() == 0.f);
        REQUIRE(ppo_core_tensor_size(algo2.core().get(), "model/pi_fc1/w") == 64 * 64);
    }
    std::printf(failures ? "FAILED (%d)\n" : "host_gpu_test OK\n", failures);
    return failures ? 1 : 0;
}
