// EnvMock — constant-valued 18/18 env, done at every 300th call (reference env/env_mock.hpp:20-91).
#ifndef PPO_B200_ENV_MOCK_HPP
#define PPO_B200_ENV_MOCK_HPP

#include "env.hpp"

class EnvMock : public Env {
public:
    explicit EnvMock(double scaling_coeff = 0.) : total_step{0}, scaling_coeff{scaling_coeff} {}
    std::string get_action_space() override { return Env::SPACE_CONTINOUS(); }
    std::string get_observation_space() override { return Env::SPACE_CONTINOUS(); }
    int get_action_space_size() override { return 18; }
    int get_observation_space_size() override { return 18; }
    Mat reset() override { return Mat::Constant(get_num_envs(), get_observation_space_size(), static_cast<float>(scaling_coeff)); }
    std::vector<Mat> step(const Mat& /*actions*/) override {
        ++total_step;
        Mat obs = Mat::Constant(get_num_envs(), get_observation_space_size(), static_cast<float>(scaling_coeff));
        Mat rewards = Mat::Constant(get_num_envs(), 1, static_cast<float>(scaling_coeff));
        Mat dones = (total_step % 300 == 0) ? Mat::Ones(get_num_envs(), 1) : Mat::Zero(get_num_envs(), 1);
        return {obs, rewards, dones};
    }
    Mat get_original_obs() override { return Mat::Constant(get_num_envs(), get_observation_space_size(), static_cast<float>(scaling_coeff)); }
    Mat get_original_rew() override { return Mat::Constant(get_num_envs(), 1, static_cast<float>(scaling_coeff)); }
    void serialize(nlohmann::json&) override {}
    void deserialize(nlohmann::json&) override {}
    void render() override {}
    float get_time() override { return 0; }

private:
    long total_step;
    double scaling_coeff;
};

#endif
