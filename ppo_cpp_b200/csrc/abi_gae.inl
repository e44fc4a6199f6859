// abi_gae.inl — part of libppo_core.so's single translation unit (included by ppo_core.cu, in this order): GAE(lambda).
// ------------------------------------------------------------------------------------------------ GAE
static int launch_gae(ppo_core* c, const float* rew, const float* val, const float* dones, const float* last_val,
                      const float* last_done, int T, int N, float gamma, float lam, float* adv, float* ret) {
    // enough (env, chunk) threads to fill the machine; chunks only when there are few envs
    const int want_threads = c->sm_count * 512;
    if (!(N < want_threads && T > 1024)) {
        LAUNCH(c, gae_kernel, dim3((N + 127) / 128, 1), 128, 0, rew, val, dones, last_val, last_done, T, N, gamma, lam, T, 0, adv, ret);
        CU(cudaGetLastError());
        return PPO_OK;
    }
    // warm-up length after which a wrong starting value has decayed far below fp32 resolution: (gamma*lam)^warm <= 2^-46
    // (2^-22 of an ulp: the chance that the residue flips one rounding is ~2e-7 per chunk boundary)
    const double gl = (double)gamma * (double)lam;
    const double need = (gl > 0.0 && gl < 1.0) ? std::ceil(std::log(std::ldexp(1.0, -46)) / std::log(gl)) : (gl <= 0.0 ? 1.0 : 1e30);
    if (need <= 4096.0) {
        const int warm = std::max(512, (int)need);
        const int nchunks = std::min((T + warm - 1) / warm, std::max(1, want_threads / std::max(N, 1)));
        int chunk = std::max((T + nchunks - 1) / nchunks, std::max(256, warm / 2));
        const dim3 grid((N + 127) / 128, (T + chunk - 1) / chunk);
        LAUNCH(c, gae_kernel, grid, 128, 0, rew, val, dones, last_val, last_done, T, N, gamma, lam, chunk, warm, adv, ret);
        CU(cudaGetLastError());
        return PPO_OK;
    }
    // gamma*lam near (or at) 1: no warm-up contracts -> exact chunk carries (affine maps in fp64, then the reference's fp32 steps)
    const int nchunks0 = std::min((T + 255) / 256, std::max(1, want_threads / std::max(N, 1)));
    const int chunk = (T + nchunks0 - 1) / nchunks0;
    const int nchunks = (T + chunk - 1) / chunk;
    const size_t need_bytes = (size_t)nchunks * N * sizeof(double2);
    if (c->gae_ab_bytes < need_bytes) {
        if (c->gae_ab) cudaFree(c->gae_ab);
        c->gae_ab = nullptr;
        c->gae_ab_bytes = 0;
        CU(cudaMalloc(&c->gae_ab, need_bytes));
        c->gae_ab_bytes = need_bytes;
    }
    const dim3 grid((N + 127) / 128, nchunks);
    LAUNCH(c, gae_affine_kernel, grid, 128, 0, rew, val, dones, last_val, last_done, T, N, gamma, lam, chunk, (double2*)c->gae_ab);
    LAUNCH(c, gae_kernel_carry, grid, 128, 0, rew, val, dones, last_val, last_done, T, N, gamma, lam, chunk, nchunks,
           (const double2*)c->gae_ab, adv, ret);
    CU(cudaGetLastError());
    return PPO_OK;
}

extern "C" int ppo_gae(ppo_core* c, const float* rewards, const float* values, const float* dones, const float* last_values,
                       const float* last_dones, int n_steps, int n_envs, float gamma, float lam, float* advs, float* returns,
                       ppo_mem mem) {
    if (!c || !rewards || !values || !dones || !last_values || !last_dones || !returns || n_steps < 1 || n_envs < 1)
        return fail(PPO_ERR_INVALID, "ppo_gae: bad arguments");
    CU(cudaSetDevice(c->desc.device));
    if (mem == PPO_DEVICE) return launch_gae(c, rewards, values, dones, last_values, last_dones, n_steps, n_envs, gamma, lam, advs, returns);
    const size_t tn = (size_t)n_steps * n_envs;
    TRY(ensure_scratch(c, 5 * tn + 2 * (size_t)n_envs));
    float* d_rew = c->scratch; float* d_val = d_rew + tn; float* d_done = d_val + tn; float* d_adv = d_done + tn; float* d_ret = d_adv + tn;
    float* d_lv = d_ret + tn; float* d_ld = d_lv + n_envs;
    TRY(h2d(c, d_rew, rewards, tn)); TRY(h2d(c, d_val, values, tn)); TRY(h2d(c, d_done, dones, tn));
    TRY(h2d(c, d_lv, last_values, n_envs)); TRY(h2d(c, d_ld, last_dones, n_envs));
    TRY(launch_gae(c, d_rew, d_val, d_done, d_lv, d_ld, n_steps, n_envs, gamma, lam, d_adv, d_ret));
    if (advs) TRY(d2h(c, advs, d_adv, tn));
    TRY(d2h(c, returns, d_ret, tn));
    CU(cudaStreamSynchronize(c->stream));
    return PPO_OK;
}
