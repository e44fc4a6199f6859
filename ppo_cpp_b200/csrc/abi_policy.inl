// abi_policy.inl — part of libppo_core.so's single translation unit (included by ppo_core.cu, in this order): policy step / value / mean (MlpPolicy::step, ::value, deterministic action).
// ------------------------------------------------------------------------------------------------ policy
static int launch_wide_policy(ppo_core* c, const PolicyArgs& a);
constexpr int WIDE_POLICY_MIN = 1024;  // below this the single launch of the tile kernel wins over five launches
static int launch_policy(ppo_core* c, PolicyArgs& a) {
    a.d = c->d;
    a.params = c->params;
    a.seed = c->desc.seed;
    a.env_id0 = (uint32_t)c->desc.env_offset;
    a.step_ctr = c->step_ctr;
    if (c->wide && a.n >= WIDE_POLICY_MIN) return launch_wide_policy(c, a);
    if (c->small) {  // thread per env
        const int ntiles = (a.n + small::NTH - 1) / small::NTH;
        const int grid = std::max(1, std::min(ntiles, c->sm_count * 8));
        LAUNCH(c, (small::policy_small_kernel<18, 18, 4, 5>), grid, small::NTH, 0, a);
        CU(cudaGetLastError());
        return PPO_OK;
    }
    if (c->fused) {
        const int ntiles = (a.n + F_TM_POLICY - 1) / F_TM_POLICY;
        const int grid = std::max(1, std::min(ntiles, c->sm_count * 2));
        LAUNCH(c, (policy_fused_kernel<F_TM_POLICY, F_NT_POLICY>), grid, F_NT_POLICY, c->fused_policy_smem, a);
        CU(cudaGetLastError());
        return PPO_OK;
    }
    const int tm = c->tm;
    const int ntiles = (a.n + tm - 1) / tm;
    const int grid = std::max(1, std::min(ntiles, c->sm_count * 4));
    if (tm == 64) LAUNCH(c, policy_tile_kernel<64>, grid, NT, policy_smem_floats<64>(c->d) * sizeof(float), a);
    else LAUNCH(c, policy_tile_kernel<32>, grid, NT, policy_smem_floats<32>(c->d) * sizeof(float), a);
    CU(cudaGetLastError());
    return PPO_OK;
}

static int policy_call(ppo_core* c, int mode, const float* obs, int n, const float* eps, float* action, float* value,
                       float* neglogp, ppo_mem mem) {
    if (!c || !obs || n < 1) return fail(PPO_ERR_INVALID, "policy call: bad arguments");
    CU(cudaSetDevice(c->desc.device));
    const int O = c->d.O, A = c->d.A;
    PolicyArgs a{};
    a.n = n;
    a.mode = mode;
    if (mem == PPO_DEVICE) {
        a.obs = obs; a.eps = eps; a.action = action; a.value = value; a.neglogp = neglogp;
        TRY(launch_policy(c, a));
        if (mode == 0 && !eps) LAUNCH(c, bump_counter_kernel, 1, 1, 0, c->step_ctr);
        return PPO_OK;
    }
    const size_t need = (size_t)n * (O + 2 * A + 2);
    TRY(ensure_scratch(c, need));
    float* d_obs = c->scratch;
    float* d_eps = d_obs + (size_t)n * O;
    float* d_act = d_eps + (size_t)n * A;
    float* d_val = d_act + (size_t)n * A;
    float* d_nlp = d_val + n;
    TRY(h2d(c, d_obs, obs, (size_t)n * O));
    if (eps) TRY(h2d(c, d_eps, eps, (size_t)n * A));
    a.obs = d_obs; a.eps = eps ? d_eps : nullptr;
    a.action = action ? d_act : nullptr; a.value = value ? d_val : nullptr; a.neglogp = neglogp ? d_nlp : nullptr;
    TRY(launch_policy(c, a));
    if (mode == 0 && !eps) LAUNCH(c, bump_counter_kernel, 1, 1, 0, c->step_ctr);
    if (action) TRY(d2h(c, action, d_act, (size_t)n * A));
    if (value) TRY(d2h(c, value, d_val, n));
    if (neglogp) TRY(d2h(c, neglogp, d_nlp, n));
    CU(cudaStreamSynchronize(c->stream));
    return PPO_OK;
}

extern "C" int ppo_policy_step(ppo_core* c, const float* obs, int n, const float* eps, float* action, float* value,
                               float* neglogp, ppo_mem mem) {
    return policy_call(c, 0, obs, n, eps, action, value, neglogp, mem);
}
extern "C" int ppo_policy_value(ppo_core* c, const float* obs, int n, float* value, ppo_mem mem) {
    return policy_call(c, 1, obs, n, nullptr, nullptr, value, nullptr, mem);
}
extern "C" int ppo_policy_mean(ppo_core* c, const float* obs, int n, float* action, ppo_mem mem) {
    return policy_call(c, 2, obs, n, nullptr, action, nullptr, nullptr, mem);
}
