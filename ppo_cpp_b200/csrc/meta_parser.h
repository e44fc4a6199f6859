// Minimal reader for TensorFlow-1.14 text-proto MetaGraphDef files (.meta.txt).
// Replaces SessionCreator::load_graph (reference ppo2/session_creator.hpp:23-57): the graph is read,
// never executed.  Extracts variable shapes, initial values and the constants the graph bakes in.
#pragma once
#include <map>
#include <string>
#include <vector>

namespace ppo {

struct MetaTensor {
    std::vector<int> shape;
    std::vector<float> data;
};

struct MetaGraph {
    int obs_dim = 0, act_dim = 0, hidden1 = 0, hidden2 = 0;
    float ent_coef = 0, vf_coef = 0, clip_norm = 0, beta1 = 0, beta2 = 0, adam_eps = 0;
    std::map<std::string, MetaTensor> tensors;  // the 15 model/* variables with their initial values
};

// Returns empty string on success, otherwise the error message.
std::string parse_meta_txt(const std::string& path, MetaGraph& out);

// Tensor order of the core's flat parameter vector: the graph's gradient order
// (GRAPH:23738-24074) followed by the untrained q head.
extern const char* const kTensorNames[15];
constexpr int kNumTrainableTensors = 13;
constexpr int kNumTensors = 15;

}  // namespace ppo
