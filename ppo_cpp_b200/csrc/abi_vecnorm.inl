// abi_vecnorm.inl — part of libppo_core.so's single translation unit (included by ppo_core.cu, in this order): VecNormalize / RunningStatistics / MatrixClamp entry points.
// ------------------------------------------------------------------------------------------------ VecNormalize
// d_raw_obs/d_raw_rew/d_done are device pointers; outputs device pointers (may alias core state).
static int vecnorm_device(ppo_core* c, const float* d_raw_obs, const float* d_raw_rew, const float* d_done, float* d_obs_out,
                          float* d_rew_out, float* d_dones_out, float* rew_store, float* urew_store, bool bump) {
    const ppo_core_desc& D = c->desc;
    const int N = D.n_envs, O = c->d.O;
    const bool upd_obs = D.training && D.norm_obs;
    const bool upd_ret = D.training && D.norm_reward && d_raw_rew;
    if (upd_obs || d_raw_rew) {
        MomentsArgs m{};
        m.raw_obs = d_raw_obs; m.raw_rew = d_raw_rew; m.ret = c->ret; m.n = N; m.D = O; m.gamma = D.norm_gamma;
        m.partial = c->mom_partial; m.moments = c->moments; m.ticket = c->ticket; m.st = c->st;
        m.update_obs = upd_obs; m.update_ret = upd_ret;
        m.fuse_merge = (D.world_size == 1) && (upd_obs || upd_ret);
        const size_t smem = sizeof(double) * (2 * (size_t)c->mom_threads + 2 * (O + 1) + 64);
        LAUNCH(c, norm_moments_kernel, c->mom_grid, c->mom_threads, smem, m);
        if (D.world_size > 1 && (upd_obs || upd_ret)) {
            TRY(need_comm(c));
            TRY(nccl_check(g_nccl.AllReduce(c->moments, c->moments, 2 * (O + 1) + 1, ncclFloat64C, ncclSumC, c->comm, c->stream), "ncclAllReduce(moments)"));
            LAUNCH(c, norm_merge_kernel, 1, 64, 0, m);
        }
    }
    ApplyArgs a{};
    a.raw_obs = d_raw_obs; a.raw_rew = d_raw_rew; a.done = d_done; a.ret = c->ret; a.n = N; a.D = O; a.st = c->st;
    a.norm_obs = D.norm_obs; a.norm_reward = D.norm_reward; a.clip_obs = D.clip_obs; a.clip_rew = D.clip_reward; a.eps = D.norm_epsilon;
    a.obs_out = d_obs_out; a.rew_out = d_rew_out; a.dones_out = d_dones_out; a.rew_store = rew_store; a.urew_store = urew_store;
    a.step_ctr = bump ? c->step_ctr : nullptr;
    const int grid = std::max(1, std::min(c->sm_count * 8, (int)(((size_t)N * O + 255) / 256)));
    LAUNCH(c, norm_apply_kernel, grid, 256, 0, a);
    CU(cudaGetLastError());
    return PPO_OK;
}

extern "C" int ppo_vecnorm_reset(ppo_core* c, const float* raw_obs, float* obs_out, ppo_mem mem) {
    if (!c || !raw_obs) return fail(PPO_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->desc.device));
    const size_t no = (size_t)c->desc.n_envs * c->d.O;
    CU(cudaMemsetAsync(c->ret, 0, c->desc.n_envs * sizeof(float), c->stream));  // ret = Zero (env_normalize.hpp:114)
    if (mem == PPO_DEVICE) return vecnorm_device(c, raw_obs, nullptr, nullptr, obs_out ? obs_out : c->cur_obs, nullptr, nullptr, nullptr, nullptr, false);
    TRY(h2d(c, c->raw_obs, raw_obs, no));
    TRY(vecnorm_device(c, c->raw_obs, nullptr, nullptr, c->cur_obs, nullptr, nullptr, nullptr, nullptr, false));
    if (obs_out) TRY(d2h(c, obs_out, c->cur_obs, no));
    CU(cudaStreamSynchronize(c->stream));
    return PPO_OK;
}

extern "C" int ppo_vecnorm_step(ppo_core* c, const float* raw_obs, const float* raw_rew, const float* done, float* obs_out,
                                float* rew_out, ppo_mem mem) {
    if (!c || !raw_obs || !raw_rew || !done) return fail(PPO_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->desc.device));
    const int N = c->desc.n_envs;
    const size_t no = (size_t)N * c->d.O;
    if (mem == PPO_DEVICE)
        return vecnorm_device(c, raw_obs, raw_rew, done, obs_out ? obs_out : c->cur_obs, rew_out ? rew_out : c->nrew, c->cur_dones, nullptr, nullptr, false);
    TRY(h2d(c, c->raw_obs, raw_obs, no));
    TRY(h2d(c, c->raw_rew, raw_rew, N));
    TRY(h2d(c, c->raw_done, done, N));
    TRY(vecnorm_device(c, c->raw_obs, c->raw_rew, c->raw_done, c->cur_obs, c->nrew, c->cur_dones, nullptr, nullptr, false));
    if (obs_out) TRY(d2h(c, obs_out, c->cur_obs, no));
    if (rew_out) TRY(d2h(c, rew_out, c->nrew, N));
    CU(cudaStreamSynchronize(c->stream));
    return PPO_OK;
}

// T consecutive ppo_vecnorm_step calls on a recorded trajectory in four launches (kernels_misc.cuh, Replay*)
extern "C" int ppo_vecnorm_replay(ppo_core* c, const float* raw_obs, const float* raw_rew, const float* done, int n_steps,
                                  float* obs_out, float* rew_out, ppo_mem mem) {
    if (!c || !raw_obs || !raw_rew || !done || !obs_out || n_steps < 1) return fail(PPO_ERR_INVALID, "ppo_vecnorm_replay: bad arguments");
    if (c->desc.world_size > 1) return fail(PPO_ERR_UNSUPPORTED, "ppo_vecnorm_replay: single rank only (per-step moments are not exchanged)");
    CU(cudaSetDevice(c->desc.device));
    const ppo_core_desc& D = c->desc;
    const int N = D.n_envs, O = c->d.O, T = n_steps;
    const size_t tn = (size_t)T * N, tno = tn * O;
    const int threads = O * std::max(1, 256 / O);
    const int NB = std::max(1, std::min(64, (int)(((size_t)N * O + (size_t)threads * 16 - 1) / ((size_t)threads * 16))));
    const size_t n_partial = (size_t)T * NB * 2 * (O + 1);  // doubles
    const size_t n_stats = (size_t)T * (2 * O + 1);
    const size_t io = mem == PPO_HOST ? 2 * tno + 3 * tn : 0;
    const size_t n_mom = (size_t)T * 2 * (O + 1);  // floats
    TRY(ensure_scratch(c, io + tn + 2 * n_partial + n_mom + n_stats + 16));
    float* p = c->scratch;
    ReplayArgs a{};
    if (mem == PPO_HOST) {
        float* d_obs = p; p += tno;
        float* d_out = p; p += tno;
        float* d_rew = p; p += tn;
        float* d_done = p; p += tn;
        float* d_rout = p; p += tn;
        TRY(h2d(c, d_obs, raw_obs, tno)); TRY(h2d(c, d_rew, raw_rew, tn)); TRY(h2d(c, d_done, done, tn));
        a.raw_obs = d_obs; a.raw_rew = d_rew; a.done = d_done; a.obs_out = d_out; a.rew_out = rew_out ? d_rout : nullptr;
    } else {
        a.raw_obs = raw_obs; a.raw_rew = raw_rew; a.done = done; a.obs_out = obs_out; a.rew_out = rew_out;
    }
    a.rt = p; p += tn;
    a.partial = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(p) + 7) & ~(uintptr_t)7);
    a.bmom = reinterpret_cast<float*>(a.partial + n_partial);
    a.stats = a.bmom + n_mom;
    a.ret = c->ret; a.T = T; a.n = N; a.D = O; a.NB = NB; a.st = c->st;
    a.update_obs = D.training && D.norm_obs; a.update_ret = D.training && D.norm_reward;
    a.norm_obs = D.norm_obs; a.norm_reward = D.norm_reward;
    a.gamma = D.norm_gamma; a.clip_obs = D.clip_obs; a.clip_rew = D.clip_reward; a.eps = D.norm_epsilon;
    LAUNCH(c, replay_ret_kernel, (N + 127) / 128, 128, 0, a);
    if (a.update_obs || a.update_ret)
        LAUNCH(c, replay_moments_kernel, dim3(NB, T), threads, sizeof(double) * (4 * (size_t)threads + 64), a);
    if (a.update_obs || a.update_ret) LAUNCH(c, replay_reduce_kernel, T, 64, 0, a);
    LAUNCH(c, replay_merge_kernel, 1, 256, REPLAY_CH * sizeof(float) * 4 * (O + 1), a);
    static const int ab_div = getenv("PPO_REPLAY_F4") ? atoi(getenv("PPO_REPLAY_F4")) : 8;  // float4 per thread (4: 0.890 ms, 8: 0.871, 16: 0.870, 32: 0.876 at 16.8 M transitions)
    const int ab = (int)std::max<size_t>(1, std::min<size_t>(1024, ((size_t)N * O / 4 + 256 * ab_div - 1) / (256 * (size_t)ab_div)));
    LAUNCH(c, replay_apply_kernel, dim3(ab, T), 256, sizeof(float) * (2 * O + 1), a);
    CU(cudaGetLastError());
    if (mem == PPO_HOST) {
        TRY(d2h(c, obs_out, a.obs_out, tno));
        if (rew_out) TRY(d2h(c, rew_out, a.rew_out, tn));
        CU(cudaStreamSynchronize(c->stream));
    }
    return PPO_OK;
}

extern "C" int ppo_vecnorm_get_stats(ppo_core* c, float* obs_mean, float* obs_var, double* obs_count, float* ret_mean,
                                     float* ret_var, double* ret_count) {
    if (!c) return fail(PPO_ERR_INVALID, "core is NULL");
    CU(cudaSetDevice(c->desc.device));
    const int O = c->d.O;
    if (obs_mean) CU(cudaMemcpyAsync(obs_mean, c->st.obs_mean, O * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (obs_var) CU(cudaMemcpyAsync(obs_var, c->st.obs_var, O * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (obs_count) CU(cudaMemcpyAsync(obs_count, c->st.obs_count, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (ret_mean) CU(cudaMemcpyAsync(ret_mean, c->st.ret_mean, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (ret_var) CU(cudaMemcpyAsync(ret_var, c->st.ret_var, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (ret_count) CU(cudaMemcpyAsync(ret_count, c->st.ret_count, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return PPO_OK;
}

extern "C" int ppo_vecnorm_set_stats(ppo_core* c, const float* obs_mean, const float* obs_var, double obs_count,
                                     const float* ret_mean, const float* ret_var, double ret_count) {
    if (!c || !obs_mean || !obs_var || !ret_mean || !ret_var) return fail(PPO_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->desc.device));
    const int O = c->d.O;
    CU(cudaMemcpyAsync(c->st.obs_mean, obs_mean, O * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->st.obs_var, obs_var, O * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->st.obs_count, &obs_count, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->st.ret_mean, ret_mean, sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->st.ret_var, ret_var, sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->st.ret_count, &ret_count, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return PPO_OK;
}

extern "C" int ppo_vecnorm_set_training(ppo_core* c, int training) {
    if (!c) return fail(PPO_ERR_INVALID, "core is NULL");
    c->desc.training = training ? 1 : 0;
    return PPO_OK;
}

extern "C" int ppo_running_stats_update(ppo_core* c, float* mean, float* var, double* count, int dim, const float* batch,
                                        int rows, ppo_mem batch_mem) {
    if (!c || !mean || !var || !count || !batch || dim < 1 || dim > 256 || rows < 1) return fail(PPO_ERR_INVALID, "bad arguments");
    CU(cudaSetDevice(c->desc.device));
    const int threads = dim * std::max(1, 256 / dim);
    const int grid = std::max(1, std::min(c->sm_count * 2, (int)(((size_t)rows * dim + threads * 8 - 1) / (threads * 8))));
    // scratch: [batch rows*dim] [mean dim] [var dim] then doubles
    const size_t nd = (size_t)grid * 2 * (dim + 1) + 2 * (dim + 1) + 1 + 2 + 2;  // partial, moments, counts(2), pad
    const size_t floats = (batch_mem == PPO_HOST ? (size_t)rows * dim : 0) + 2 * (size_t)dim + 4 + 2 * nd + 8;
    TRY(ensure_scratch(c, floats));
    float* p = c->scratch;
    const float* d_batch = batch;
    if (batch_mem == PPO_HOST) {
        TRY(h2d(c, p, batch, (size_t)rows * dim));
        d_batch = p;
        p += (size_t)rows * dim;
    }
    float* d_mean = p; p += dim;
    float* d_var = p; p += dim;
    float* d_dummy = p; p += 2;
    double* dd = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(p) + 7) & ~(uintptr_t)7);
    double* d_partial = dd; dd += (size_t)grid * 2 * (dim + 1);
    double* d_moments = dd; dd += 2 * (dim + 1) + 1;
    double* d_count = dd; dd += 1;
    double* d_count2 = dd; dd += 1;
    unsigned int* d_ticket = reinterpret_cast<unsigned int*>(dd);
    CU(cudaMemcpyAsync(d_mean, mean, dim * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_var, var, dim * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_count, count, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(d_ticket, 0, sizeof(unsigned int), c->stream));
    MomentsArgs m{};
    m.raw_obs = d_batch; m.raw_rew = nullptr; m.ret = nullptr; m.n = rows; m.D = dim; m.gamma = 0.f;
    m.partial = d_partial; m.moments = d_moments; m.ticket = d_ticket;
    m.st.obs_mean = d_mean; m.st.obs_var = d_var; m.st.obs_count = d_count;
    m.st.ret_mean = d_dummy; m.st.ret_var = d_dummy + 1; m.st.ret_count = d_count2;
    m.fuse_merge = 1; m.update_obs = 1; m.update_ret = 0;
    const size_t smem = sizeof(double) * (2 * (size_t)threads + 2 * (dim + 1) + 64);
    LAUNCH(c, norm_moments_kernel, grid, threads, smem, m);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(mean, d_mean, dim * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(var, d_var, dim * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(count, d_count, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return PPO_OK;
}

extern "C" int ppo_matrix_clamp(ppo_core* c, const float* x, size_t n, float lo, float hi, float* out, ppo_mem mem) {
    if (!c || !x || !out) return fail(PPO_ERR_INVALID, "NULL argument");
    if (n == 0) return PPO_OK;
    CU(cudaSetDevice(c->desc.device));
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)c->sm_count * 8, (n + 255) / 256));
    if (mem == PPO_DEVICE) {
        LAUNCH(c, clamp_kernel, grid, 256, 0, x, n, lo, hi, out);
        CU(cudaGetLastError());
        return PPO_OK;
    }
    TRY(ensure_scratch(c, n));
    TRY(h2d(c, c->scratch, x, n));
    LAUNCH(c, clamp_kernel, grid, 256, 0, c->scratch, n, lo, hi, c->scratch);
    CU(cudaGetLastError());
    TRY(d2h(c, out, c->scratch, n));
    CU(cudaStreamSynchronize(c->stream));
    return PPO_OK;
}
