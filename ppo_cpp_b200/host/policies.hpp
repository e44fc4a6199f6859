// MlpPolicy with the reference's three calls (ppo2/policies.hpp:25-82); tensorflow::Tensor is replaced by Mat
// ("that type must be replaced").  Each call is one C-ABI call instead of one Session::Run.
#ifndef PPO_B200_POLICIES_HPP
#define PPO_B200_POLICIES_HPP

#include <vector>

#include "core_handle.hpp"
#include "mat.hpp"

class MlpPolicy {
public:
    explicit MlpPolicy(CorePtr core, int act_dim) : core_{std::move(core)}, act_dim_{act_dim} {}

    // {action [n,A], value [n,1], neglogp [n,1]}  (fetch order of policies.hpp:37)
    std::vector<Mat> step(const Mat& obs) {
        Mat action(obs.rows(), act_dim_), value(obs.rows(), 1), neglogp(obs.rows(), 1);
        ppo_check(ppo_policy_step(core_.get(), obs.data(), obs.rows(), nullptr, action.data(), value.data(), neglogp.data(), PPO_HOST), "step");
        return {std::move(action), std::move(value), std::move(neglogp)};
    }
    Mat get_deterministic_action(const Mat& obs) {
        Mat action(obs.rows(), act_dim_);
        ppo_check(ppo_policy_mean(core_.get(), obs.data(), obs.rows(), action.data(), PPO_HOST), "get_action(): evaluation step");
        return action;
    }
    Mat value(const Mat& obs) {
        Mat v(obs.rows(), 1);
        ppo_check(ppo_policy_value(core_.get(), obs.data(), obs.rows(), v.data(), PPO_HOST), "value()");
        return v;
    }
    const CorePtr& core() const { return core_; }

private:
    CorePtr core_;
    int act_dim_;
};

#endif
