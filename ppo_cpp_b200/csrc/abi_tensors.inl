// abi_tensors.inl — part of libppo_core.so's single translation unit (included by ppo_core.cu, in this order): tensors by name, graph file and checkpoint I/O.
// ------------------------------------------------------------------------------------------------ tensors
extern "C" int ppo_core_num_tensors(void) { return kNumTensors; }
extern "C" const char* ppo_core_tensor_name(int i) { return (i >= 0 && i < kNumTensors) ? kTensorNames[i] : nullptr; }

static int resolve_tensor(ppo_core* c, const char* name, float** ptr, int* count) {
    const NetDims& d = c->d;
    const std::string s(name ? name : "");
    if (s == "params") { *ptr = c->params; *count = d.Pq; return PPO_OK; }
    if (s == "params_trainable") { *ptr = c->params; *count = d.P; return PPO_OK; }
    if (s == "adam_m") { *ptr = c->adam_m; *count = d.P; return PPO_OK; }
    if (s == "adam_v") { *ptr = c->adam_v; *count = d.P; return PPO_OK; }
    if (s == "grad") { *ptr = c->grad; *count = d.P; return PPO_OK; }
    if (s == "beta1_power") { *ptr = c->bpow + c->bpow_slot * 2; *count = 1; return PPO_OK; }
    if (s == "beta2_power") { *ptr = c->bpow + c->bpow_slot * 2 + 1; *count = 1; return PPO_OK; }
    for (int t = 0; t < kNumTensors; ++t) {
        const std::string base(kTensorNames[t]);
        const int n = d.off[t + 1] - d.off[t];
        if (s == base) { *ptr = c->params + d.off[t]; *count = n; return PPO_OK; }
        if (t < kNumTrainableTensors) {
            if (s == base + "/Adam") { *ptr = c->adam_m + d.off[t]; *count = n; return PPO_OK; }
            if (s == base + "/Adam_1") { *ptr = c->adam_v + d.off[t]; *count = n; return PPO_OK; }
        }
    }
    return fail(PPO_ERR_INVALID, "unknown tensor name '%s'", s.c_str());
}

extern "C" int ppo_core_tensor_size(ppo_core* c, const char* name) {
    if (!c) return fail(PPO_ERR_INVALID, "core is NULL");
    float* p; int n;
    const int st = resolve_tensor(c, name, &p, &n);
    return st == PPO_OK ? n : st;
}
extern "C" int ppo_core_get_tensor(ppo_core* c, const char* name, float* out, size_t cap) {
    if (!c || !out) return fail(PPO_ERR_INVALID, "NULL argument");
    float* p; int n;
    TRY(resolve_tensor(c, name, &p, &n));
    if (cap < (size_t)n) return fail(PPO_ERR_INVALID, "buffer for '%s' holds %zu floats, need %d", name, cap, n);
    CU(cudaSetDevice(c->desc.device));
    CU(cudaMemcpyAsync(out, p, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return PPO_OK;
}
extern "C" int ppo_core_set_tensor(ppo_core* c, const char* name, const float* in, size_t count) {
    if (!c || !in) return fail(PPO_ERR_INVALID, "NULL argument");
    float* p; int n;
    TRY(resolve_tensor(c, name, &p, &n));
    if (count != (size_t)n) return fail(PPO_ERR_INVALID, "tensor '%s' has %d floats, got %zu", name, n, count);
    CU(cudaSetDevice(c->desc.device));
    CU(cudaMemcpyAsync(p, in, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->wide_images_valid = false;
    return PPO_OK;
}

extern "C" int ppo_core_load_meta_txt(ppo_core* c, const char* path) {
    if (!c || !path) return fail(PPO_ERR_INVALID, "NULL argument");
    ppo_meta_info info;
    std::vector<float> p(c->d.Pq);
    ppo_meta_info probe;
    TRY(ppo_meta_parse(path, &probe, nullptr, 0));
    if (probe.obs_dim != c->d.O || probe.act_dim != c->d.A || probe.hidden1 != c->d.H1 || probe.hidden2 != c->d.H2)
        return fail(PPO_ERR_INVALID, "graph %s is obs %d act %d MLP [%d,%d]; core was created for obs %d act %d MLP [%d,%d]", path,
                    probe.obs_dim, probe.act_dim, probe.hidden1, probe.hidden2, c->d.O, c->d.A, c->d.H1, c->d.H2);
    TRY(ppo_meta_parse(path, &info, p.data(), p.size()));
    // the graph's baked constants win over constructor arguments, as in the reference (SURVEY §3.5 "Consequence")
    c->desc.ent_coef = info.ent_coef; c->desc.vf_coef = info.vf_coef; c->desc.max_grad_norm = info.max_grad_norm;
    c->desc.adam_beta1 = info.adam_beta1; c->desc.adam_beta2 = info.adam_beta2; c->desc.adam_epsilon = info.adam_epsilon;
    TRY(ppo_core_set_tensor(c, "params", p.data(), p.size()));
    // reset() re-creates the session: Adam state starts from zero, beta powers at beta (ppo2.hpp:90-105)
    CU(cudaMemsetAsync(c->adam_m, 0, c->d.P * sizeof(float), c->stream));
    CU(cudaMemsetAsync(c->adam_v, 0, c->d.P * sizeof(float), c->stream));
    const float bp[4] = {info.adam_beta1, info.adam_beta2, info.adam_beta1, info.adam_beta2};
    CU(cudaMemcpyAsync(c->bpow, bp, sizeof(bp), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->bpow_slot = 0;
    return PPO_OK;
}

// Stable-Baselines ortho_init(scale): QR-free variant via modified Gram-Schmidt on a Gaussian matrix
// (scale sqrt(2) hidden, 1.0 value head, 0.01 policy/q heads; biases and logstd zero) — SURVEY §3.4.
extern "C" int ppo_core_init_orthogonal(ppo_core* c, uint64_t seed) {
    if (!c) return fail(PPO_ERR_INVALID, "core is NULL");
    const NetDims& d = c->d;
    std::vector<float> p(d.Pq, 0.f);
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    auto next_u = [&]() -> double {  // splitmix64 -> (0,1)
        s += 0x9E3779B97F4A7C15ull;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        return ((double)(z >> 11) + 0.5) / 9007199254740992.0;
    };
    auto gauss = [&]() -> double { return std::sqrt(-2.0 * std::log(next_u())) * std::cos(6.283185307179586 * next_u()); };
    auto ortho = [&](int t, int rows, int cols, double scale) {
        // orthonormalise the shorter dimension's vectors
        const bool tall = rows >= cols;
        const int nv = tall ? cols : rows, len = tall ? rows : cols;
        std::vector<std::vector<double>> v(nv, std::vector<double>(len));
        for (auto& vec : v) for (auto& x : vec) x = gauss();
        for (int i = 0; i < nv; ++i) {
            for (int j = 0; j < i; ++j) {
                double dot = 0;
                for (int k = 0; k < len; ++k) dot += v[i][k] * v[j][k];
                for (int k = 0; k < len; ++k) v[i][k] -= dot * v[j][k];
            }
            double nrm = 0;
            for (int k = 0; k < len; ++k) nrm += v[i][k] * v[i][k];
            nrm = std::sqrt(nrm);
            for (int k = 0; k < len; ++k) v[i][k] /= nrm;
        }
        float* w = p.data() + d.off[t];
        for (int r = 0; r < rows; ++r)
            for (int cc = 0; cc < cols; ++cc) w[(size_t)r * cols + cc] = (float)(scale * (tall ? v[cc][r] : v[r][cc]));
    };
    const double s2 = std::sqrt(2.0);
    ortho(T_PI_FC0_W, d.O, d.H1, s2); ortho(T_VF_FC0_W, d.O, d.H1, s2);
    ortho(T_PI_FC1_W, d.H1, d.H2, s2); ortho(T_VF_FC1_W, d.H1, d.H2, s2);
    ortho(T_VF_W, d.H2, 1, 1.0); ortho(T_PI_W, d.H2, d.A, 0.01); ortho(T_Q_W, d.H2, d.A, 0.01);
    TRY(ppo_core_set_tensor(c, "params", p.data(), p.size()));
    CU(cudaMemsetAsync(c->adam_m, 0, d.P * sizeof(float), c->stream));
    CU(cudaMemsetAsync(c->adam_v, 0, d.P * sizeof(float), c->stream));
    const float bp[4] = {c->desc.adam_beta1, c->desc.adam_beta2, c->desc.adam_beta1, c->desc.adam_beta2};
    CU(cudaMemcpyAsync(c->bpow, bp, sizeof(bp), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->bpow_slot = 0;
    return PPO_OK;
}

// TF Saver V2 data file: tensors in sorted-name order, raw little-endian fp32
static const int kCkptOrder[15] = {T_PI_B, T_LOGSTD, T_PI_W, T_PI_FC0_B, T_PI_FC0_W, T_PI_FC1_B, T_PI_FC1_W, T_Q_B,
                                   T_Q_W, T_VF_B, T_VF_W, T_VF_FC0_B, T_VF_FC0_W, T_VF_FC1_B, T_VF_FC1_W};

extern "C" int ppo_core_load_checkpoint_data(ppo_core* c, const char* prefix) {
    if (!c || !prefix) return fail(PPO_ERR_INVALID, "NULL argument");
    const std::string path = std::string(prefix) + ".data-00000-of-00001";
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return fail(PPO_ERR_IO, "cannot open %s", path.c_str());
    std::vector<float> raw(c->d.Pq), p(c->d.Pq);
    const size_t got = fread(raw.data(), sizeof(float), raw.size(), f);
    const bool extra = fgetc(f) != EOF;
    fclose(f);
    if (got != raw.size() || extra) return fail(PPO_ERR_IO, "%s does not hold exactly %d floats (MLP [%d,%d])", path.c_str(), c->d.Pq, c->d.H1, c->d.H2);
    size_t off = 0;
    for (int i = 0; i < 15; ++i) {
        const int t = kCkptOrder[i], n = c->d.off[t + 1] - c->d.off[t];
        memcpy(p.data() + c->d.off[t], raw.data() + off, n * sizeof(float));
        off += n;
    }
    return ppo_core_set_tensor(c, "params", p.data(), p.size());
}

extern "C" int ppo_core_save_checkpoint_data(ppo_core* c, const char* prefix) {
    if (!c || !prefix) return fail(PPO_ERR_INVALID, "NULL argument");
    std::vector<float> p(c->d.Pq), raw(c->d.Pq);
    TRY(ppo_core_get_tensor(c, "params", p.data(), p.size()));
    size_t off = 0;
    for (int i = 0; i < 15; ++i) {
        const int t = kCkptOrder[i], n = c->d.off[t + 1] - c->d.off[t];
        memcpy(raw.data() + off, p.data() + c->d.off[t], n * sizeof(float));
        off += n;
    }
    const std::string path = std::string(prefix) + ".data-00000-of-00001";
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return fail(PPO_ERR_IO, "cannot open %s for writing", path.c_str());
    const size_t put = fwrite(raw.data(), sizeof(float), raw.size(), f);
    fclose(f);
    if (put != raw.size()) return fail(PPO_ERR_IO, "short write to %s", path.c_str());
    const int st = ppo_checkpoint_write_index(prefix, c->d.O, c->d.A, c->d.H1, c->d.H2, raw.data(), raw.size());
    return st == PPO_OK ? PPO_OK : fail(st, "cannot write %s.index", prefix);
}
