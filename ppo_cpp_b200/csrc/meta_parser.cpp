// See meta_parser.h.  A ~150-line recursive-descent reader for protobuf text format, enough for
// MetaGraphDef: nested messages, string literals with C escapes, numbers and enum identifiers.
#include "meta_parser.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>

namespace ppo {

const char* const kTensorNames[15] = {
    "model/pi_fc0/w", "model/pi_fc0/b", "model/vf_fc0/w", "model/vf_fc0/b", "model/pi_fc1/w",
    "model/pi_fc1/b", "model/vf_fc1/w", "model/vf_fc1/b", "model/vf/w",     "model/vf/b",
    "model/pi/w",     "model/pi/b",     "model/pi/logstd", "model/q/w",     "model/q/b"};

namespace {

struct Msg;
struct Field {
    std::string key;
    std::string scalar;         // string bytes / number / identifier
    std::unique_ptr<Msg> msg;   // non-null for nested messages
};
struct Msg {
    std::vector<Field> fields;
    const Field* find(const char* key) const {
        for (const auto& f : fields)
            if (f.key == key) return &f;
        return nullptr;
    }
    const Msg* sub(const char* key) const {
        const Field* f = find(key);
        return f && f->msg ? f->msg.get() : nullptr;
    }
    std::string str(const char* key) const {
        const Field* f = find(key);
        return f ? f->scalar : std::string();
    }
};

struct Parser {
    const char* p;
    const char* end;
    std::string err;

    void skip_ws() {
        while (p < end) {
            if (*p == '#') {
                while (p < end && *p != '\n') ++p;
            } else if (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r') {
                ++p;
            } else {
                break;
            }
        }
    }
    bool parse_string(std::string& out) {
        ++p;  // opening quote
        while (p < end && *p != '"') {
            if (*p != '\\') {
                out.push_back(*p++);
                continue;
            }
            ++p;
            if (p >= end) return false;
            char c = *p;
            if (c >= '0' && c <= '7') {
                int v = 0, n = 0;
                while (p < end && n < 3 && *p >= '0' && *p <= '7') {
                    v = v * 8 + (*p - '0');
                    ++p;
                    ++n;
                }
                out.push_back(static_cast<char>(v & 0xFF));
                continue;
            }
            ++p;
            switch (c) {
                case 'n': out.push_back('\n'); break;
                case 't': out.push_back('\t'); break;
                case 'r': out.push_back('\r'); break;
                case 'a': out.push_back('\a'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case 'v': out.push_back('\v'); break;
                case 'x': {
                    int v = 0, n = 0;
                    while (p < end && n < 2 && isxdigit(static_cast<unsigned char>(*p))) {
                        char h = *p++;
                        v = v * 16 + (h <= '9' ? h - '0' : (h | 32) - 'a' + 10);
                        ++n;
                    }
                    out.push_back(static_cast<char>(v));
                    break;
                }
                default: out.push_back(c); break;  // \" \' \\ ...
            }
        }
        if (p >= end) return false;
        ++p;  // closing quote
        return true;
    }
    bool parse_msg(Msg& m, bool top) {
        for (;;) {
            skip_ws();
            if (p >= end) return top;
            if (*p == '}') {
                ++p;
                return !top;
            }
            const char* k0 = p;
            while (p < end && (isalnum(static_cast<unsigned char>(*p)) || *p == '_')) ++p;
            if (p == k0) {
                err = "text-proto: field name expected";
                return false;
            }
            Field f;
            f.key.assign(k0, p);
            skip_ws();
            if (p < end && *p == ':') {
                ++p;
                skip_ws();
            }
            if (p >= end) return false;
            if (*p == '{') {
                ++p;
                f.msg.reset(new Msg());
                if (!parse_msg(*f.msg, false)) return false;
            } else if (*p == '"') {
                if (!parse_string(f.scalar)) {
                    err = "text-proto: unterminated string";
                    return false;
                }
            } else {
                const char* v0 = p;
                while (p < end && !isspace(static_cast<unsigned char>(*p)) && *p != '}') ++p;
                f.scalar.assign(v0, p);
            }
            m.fields.push_back(std::move(f));
        }
    }
};

const Msg* attr_value(const Msg& node, const char* key) {
    for (const auto& f : node.fields) {
        if (f.key != "attr" || !f.msg) continue;
        if (f.msg->str("key") == key) return f.msg->sub("value");
    }
    return nullptr;
}

std::vector<int> shape_dims(const Msg* shape) {
    std::vector<int> d;
    if (!shape) return d;
    for (const auto& f : shape->fields)
        if (f.key == "dim" && f.msg) d.push_back(std::atoi(f.msg->str("size").c_str()));
    return d;
}

bool const_tensor(const Msg& node, MetaTensor& t) {
    const Msg* v = attr_value(node, "value");
    const Msg* ten = v ? v->sub("tensor") : nullptr;
    if (!ten) return false;
    t.shape = shape_dims(ten->sub("tensor_shape"));
    size_t count = 1;
    for (int d : t.shape) count *= static_cast<size_t>(d);
    const Field* content = ten->find("tensor_content");
    if (content) {
        if (content->scalar.size() != count * 4) return false;
        t.data.resize(count);
        std::memcpy(t.data.data(), content->scalar.data(), count * 4);  // little-endian fp32
        return true;
    }
    std::vector<float> vals;
    for (const auto& f : ten->fields)
        if (f.key == "float_val") vals.push_back(std::strtof(f.scalar.c_str(), nullptr));
    if (vals.empty()) {
        t.data.assign(count, 0.f);  // all-zero tensors carry no value field
    } else if (vals.size() == 1) {
        t.data.assign(count, vals[0]);
    } else {
        if (vals.size() != count) return false;
        t.data = vals;
    }
    return true;
}

}  // namespace

std::string parse_meta_txt(const std::string& path, MetaGraph& out) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return "cannot open graph file: " + path;
    std::stringstream ss;
    ss << in.rdbuf();
    const std::string text = ss.str();
    Parser ps{text.data(), text.data() + text.size(), {}};
    Msg root;
    if (!ps.parse_msg(root, true)) return "cannot parse " + path + ": " + (ps.err.empty() ? "unbalanced braces" : ps.err);
    const Msg* graph = root.sub("graph_def");
    if (!graph) return "no graph_def in " + path;
    std::map<std::string, const Msg*> nodes;
    for (const auto& f : graph->fields)
        if (f.key == "node" && f.msg) nodes[f.msg->str("name")] = f.msg.get();

    for (int i = 0; i < kNumTensors; ++i) {
        const std::string name = kTensorNames[i];
        auto it = nodes.find(name);
        if (it == nodes.end() || it->second->str("op") != "VariableV2") return "graph has no variable " + name;
        const Msg* shp = attr_value(*it->second, "shape");
        MetaTensor t;
        std::vector<int> vshape = shape_dims(shp ? shp->sub("shape") : nullptr);
        bool found = false;
        for (const char* suffix : {"initial_value", "Const", "zeros"}) {
            auto init = nodes.find(name + "/Initializer/" + suffix);
            if (init != nodes.end() && init->second->str("op") == "Const") {
                if (!const_tensor(*init->second, t)) return "bad initializer for " + name;
                found = true;
                break;
            }
        }
        if (!found) return "no Const initializer for " + name;
        size_t count = 1;
        for (int d : vshape) count *= static_cast<size_t>(d);
        if (t.data.size() != count) return "initializer/variable size mismatch for " + name;
        t.shape = vshape;
        out.tensors[name] = std::move(t);
    }
    auto scalar = [&](const char* name, float& dst) -> bool {
        auto it = nodes.find(name);
        MetaTensor t;
        if (it == nodes.end() || !const_tensor(*it->second, t) || t.data.empty()) return false;
        dst = t.data[0];
        return true;
    };
    if (!scalar("loss/mul_4/y", out.ent_coef)) return "graph has no loss/mul_4/y (entropy coefficient)";
    if (!scalar("loss/mul_5/y", out.vf_coef)) return "graph has no loss/mul_5/y (value coefficient)";
    if (!scalar("loss/clip_by_global_norm/mul/x", out.clip_norm)) return "graph has no clip norm constant";
    if (!scalar("ppo2/_train/beta1", out.beta1) || !scalar("ppo2/_train/beta2", out.beta2) ||
        !scalar("ppo2/_train/epsilon", out.adam_eps))
        return "graph has no Adam constants";
    const auto& w0 = out.tensors["model/pi_fc0/w"].shape;
    const auto& w1 = out.tensors["model/pi_fc1/w"].shape;
    const auto& wp = out.tensors["model/pi/w"].shape;
    if (w0.size() != 2 || w1.size() != 2 || wp.size() != 2 || w0[1] != w1[0] || w1[1] != wp[0])
        return "unexpected MLP variable shapes in " + path;
    out.obs_dim = w0[0];
    out.hidden1 = w0[1];
    out.hidden2 = w1[1];
    out.act_dim = wp[1];
    return std::string();
}

}  // namespace ppo
