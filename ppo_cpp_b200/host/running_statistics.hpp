// RunningStatistics / MatrixClamp with the reference's interface (common/running_statistics.hpp:12-114,
// common/matrix_clamp.hpp:10-40); the arithmetic runs on the device through the C ABI.
#ifndef PPO_B200_RUNNING_STATISTICS_HPP
#define PPO_B200_RUNNING_STATISTICS_HPP

#include "core_handle.hpp"
#include "env.hpp"

class RunningStatistics : public virtual ISerializable {
public:
    explicit RunningStatistics(int space_size = 1, float epsilon = 1e-6)
        : mean{Mat::Zero(1, space_size)}, var{Mat::Ones(1, space_size)}, count{epsilon}, space_size{space_size} {}

    // any core works: the call is stateless on the device side
    void update(ppo_core* core, const Mat& batch) {
        assert(batch.cols() == space_size);
        ppo_check(ppo_running_stats_update(core, mean.data(), var.data(), &count, space_size, batch.data(), batch.rows(), PPO_HOST),
                  "RunningStatistics::update");
    }
    void serialize(nlohmann::json& json) override {
        json["var"] = std::vector<float>(var.data(), var.data() + var.size());
        json["mean"] = std::vector<float>(mean.data(), mean.data() + mean.size());
        json["count"] = count;
    }
    void deserialize(nlohmann::json& json) override {
        count = json["count"].get<double>();
        auto var_v = json["var"].get<std::vector<float>>();
        auto mean_v = json["mean"].get<std::vector<float>>();
        assert(static_cast<int>(var_v.size()) == space_size && static_cast<int>(mean_v.size()) == space_size);
        for (int i = 0; i < space_size; ++i) {
            var(0, i) = var_v[i];
            mean(0, i) = mean_v[i];
        }
    }

    Mat mean;
    Mat var;
    double count;

private:
    int space_size;
};

class MatrixClamp {
public:
    MatrixClamp(const Mat& like, float clamp) : MatrixClamp(like.rows(), like.cols(), -clamp, clamp) {}
    MatrixClamp(int rows, int cols, float clamp) : MatrixClamp(rows, cols, -clamp, clamp) {}
    MatrixClamp(int rows, int cols, float low, float high) : rows{rows}, cols{cols}, lo{low}, hi{high} {}
    Mat clamp(ppo_core* core, const Mat& mat) const {
        Mat result(mat.rows(), mat.cols());
        ppo_check(ppo_matrix_clamp(core, mat.data(), mat.size(), lo, hi, result.data(), PPO_HOST), "MatrixClamp::clamp");
        return result;
    }

private:
    int rows, cols;
    float lo, hi;
};

#endif
