// abi_introspect.inl — part of libppo_core.so's single translation unit (included by ppo_core.cu, in this order): counters, kernel families, per-kernel timing (ppo_profile_kernel).
extern "C" int ppo_core_counters(ppo_core* c, ppo_counters* out, int reset) {
    if (!c) return fail(PPO_ERR_INVALID, "core is NULL");
    if (out) *out = c->ctr;
    if (reset) c->ctr = ppo_counters{};
    return PPO_OK;
}

extern "C" const char* ppo_core_kernel_family(ppo_core* c, const char* which) {
    if (!c || !which) return nullptr;
    const std::string w(which);
    if (w == "train") {
        if (c->wide) return "wgemm_kernel (tcgen05.mma kind::f16, fp16x2 split operand images, layer-wise GEMMs with bulk-copy pipeline)";
        if (c->small && c->small_epoch) return "train_small_kernel (thread per sample, fp32 FFMA in registers, warp-transpose gradient sums; persistent: train_small_epoch_kernel, one single-CTA launch per epoch with combine + clip + Adam in shared memory)";
        if (c->small) return "train_small_kernel (thread per sample, fp32 FFMA in registers, warp-transpose gradient sums)";
        if (c->umma && c->persistent_epoch && fast_path(c))
            return "train_umma_kernel (tcgen05.mma kind::f16, fp16x2 split operands, fp32 TMEM accumulators; persistent: one cooperative launch per epoch, reduce + Adam inside)";
        if (c->umma) return "train_umma_kernel (tcgen05.mma kind::f16, fp16x2 split operands, fp32 TMEM accumulators)";
        if (c->fused) return "train_fused_kernel (fp32 FFMA, weights staged in shared memory)";
        return "train_tile_kernel (fp32 FFMA, generic hidden sizes)";
    }
    if (w == "rollout") return (c->persistent_rollout && fast_path(c)) ? "rollout_persistent_kernel (one cooperative launch per rollout)" : "per-step kernels";
    if (w == "policy") {
        if (c->wide && c->desc.n_envs >= WIDE_POLICY_MIN) return "wgemm_kernel forward (tcgen05, split-bf16 operand images) + wide_policy_head_kernel";
        if (c->small) return "policy_small_kernel (thread per env, fp32 FFMA in registers, parameters in shared memory)";
        if (c->fused) return "policy_fused_kernel (fp32 FFMA, weights staged in shared memory)";
        return "policy_tile_kernel (fp32 FFMA, generic hidden sizes)";
    }
    return nullptr;
}

extern "C" int ppo_profile_kernel(ppo_core* c, const char* which, int iters, float* avg_ms, int* launches) {
    if (!c || !which || iters < 1 || !avg_ms) return fail(PPO_ERR_INVALID, "ppo_profile_kernel: bad arguments");
    CU(cudaSetDevice(c->desc.device));
    const std::string w(which);
    const int N = c->desc.n_envs, T = c->desc.n_steps, W = c->desc.world_size;
    if ((w == "train_fwdbwd" || w == "grad_reduce" || w == "adam") && !c->perm_set) {
        // identity permutation is as good as any for timing
        for (int i = 0; i < c->n_batch_global; ++i) c->perm_pinned[i] = i;
        TRY(prepare_epoch(c, c->perm_pinned));
        c->perm_set = true;
    }
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const uint64_t before = c->ctr.kernel_launches;
    int st = PPO_OK;
    const int bpow_slot = c->bpow_slot;
    for (int pass = 0; pass < 2 && st == PPO_OK; ++pass) {  // pass 0 = warm-up
        const int n = pass == 0 ? std::min(iters, 3) : iters;
        if (pass == 1) cudaEventRecord(e0, c->stream);
        for (int i = 0; i < n && st == PPO_OK; ++i) {
            if (w == "train_fwdbwd" || w == "grad_reduce") {
                const int k = i % c->desc.nminibatches, per_rank = c->B_global / W;
                TrainArgs a{};
                a.obs = c->buf[B_OBS]; a.act = c->buf[B_ACTIONS]; a.ret = c->buf[B_RETURNS]; a.val = c->buf[B_VALUES]; a.nlp = c->buf[B_NEGLOGP];
                a.gather = c->cur_gather; a.mbstats = c->cur_mbstats + k; a.slot0 = k * c->B_global + c->desc.rank * per_rank; a.count = per_rank;
                a.invB = 1.0f / (float)c->B_global; a.cliprange = 0.2f;
                a.d = c->d; a.params = c->params; a.ent_coef = c->desc.ent_coef / (float)W; a.vf_coef = c->desc.vf_coef; a.partial = c->partial; a.PS = c->PS;
                if (w == "train_fwdbwd") {
                    st = launch_train_kernel(c, a, false, &c->prof_train_grid);
                } else {
                    if (c->prof_train_grid == 0) st = launch_train_kernel(c, a, false, &c->prof_train_grid);
                    LAUNCH(c, grad_reduce_kernel, c->n_sq_blocks, 256, 0, c->partial, c->prof_train_grid, c->PS, c->d.P, c->grad, c->sq_partial);
                }
            } else if (w == "adam") {
                AdamArgs ad{};
                ad.params = c->params; ad.m = c->adam_m; ad.v = c->adam_v; ad.grad = c->grad; ad.sq_partial = c->sq_partial;
                ad.nblk = c->n_sq_blocks; ad.P = c->d.P; ad.lr = 0.f; ad.beta1 = 1.f; ad.beta2 = 1.f;  // state unchanged
                ad.eps = c->desc.adam_epsilon; ad.clip_norm = c->desc.max_grad_norm;
                ad.bpow_in = c->bpow + bpow_slot * 2; ad.bpow_out = c->bpow + 4 - 4 + (bpow_slot ^ 1) * 2;
                ad.invB = 1.f; ad.inv_world = 1.f; ad.loss_row = c->loss_rows + (size_t)c->desc.noptepochs * c->desc.nminibatches * 5;
                ad.gnorm_out = c->gnorm;
                // keep the beta powers: write the same values to the other slot
                ad.beta1 = 1.f; ad.beta2 = 1.f;
                LAUNCH(c, adam_kernel, (c->d.P + 255) / 256, 256, 0, ad);
            } else if (w == "policy_step") {
                PolicyArgs a{};
                a.obs = c->cur_obs; a.n = N; a.mode = 0; a.action = c->cur_actions; a.value = c->last_values; a.neglogp = c->nrew;
                st = launch_policy(c, a);
            } else if (w == "norm_moments") {
                MomentsArgs m{};
                m.raw_obs = c->raw_obs; m.raw_rew = nullptr; m.ret = c->ret; m.n = N; m.D = c->d.O; m.gamma = c->desc.norm_gamma;
                m.partial = c->mom_partial; m.moments = c->moments; m.ticket = c->ticket; m.st = c->st;
                m.update_obs = 0; m.update_ret = 0; m.fuse_merge = 0;
                LAUNCH(c, norm_moments_kernel, c->mom_grid, c->mom_threads, sizeof(double) * (2 * (size_t)c->mom_threads + 2 * (c->d.O + 1) + 64), m);
            } else if (w == "norm_apply") {
                ApplyArgs a{};
                a.raw_obs = c->raw_obs; a.raw_rew = nullptr; a.done = nullptr; a.ret = c->ret; a.n = N; a.D = c->d.O; a.st = c->st;
                a.norm_obs = 1; a.norm_reward = 1; a.clip_obs = c->desc.clip_obs; a.clip_rew = c->desc.clip_reward; a.eps = c->desc.norm_epsilon;
                a.obs_out = c->cur_obs;
                LAUNCH(c, norm_apply_kernel, std::max(1, std::min(c->sm_count * 8, (int)(((size_t)N * c->d.O + 255) / 256))), 256, 0, a);
            } else if (w == "vecnorm_replay") {  // in place over the rollout buffers (timing only: the statistics keep moving)
                st = ppo_vecnorm_replay(c, slab(c, B_OBS, 0), slab(c, B_TRUE_REW, 0), slab(c, B_DONES, 0), T, slab(c, B_OBS, 0),
                                        slab(c, B_UNNORM_REW, 0), PPO_DEVICE);
            } else if (w == "gae") {
                st = launch_gae(c, slab(c, B_TRUE_REW, 0), slab(c, B_VALUES, 0), slab(c, B_DONES, 0), c->last_values, c->cur_dones, T, N,
                                c->desc.gamma, c->desc.lam, nullptr, slab(c, B_RETURNS, 0));
            } else {
                st = fail(PPO_ERR_INVALID, "unknown kernel '%s'", which);
            }
        }
        if (pass == 1) cudaEventRecord(e1, c->stream);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) st = fail(PPO_ERR_CUDA, "profile: %s", cudaGetErrorString(cudaGetLastError()));
    }
    float ms = 0.f;
    if (st == PPO_OK && cudaEventElapsedTime(&ms, e0, e1) != cudaSuccess) st = fail(PPO_ERR_CUDA, "cudaEventElapsedTime failed");
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (st != PPO_OK) return st;
    if (c->umma_prof && w == "train_fwdbwd") {
        long long h[176];
        cudaMemcpy(h, c->umma_prof, sizeof(h), cudaMemcpyDeviceToHost);
        for (int t = 0; t < 2; ++t) {
            fprintf(stderr, "umma phases, stand-alone kernel, tower %d (cycles):", t);
            for (int i = 1; i < 32 && h[t * 32 + i]; ++i) fprintf(stderr, " %lld", h[t * 32 + i] - h[t * 32 + i - 1]);
            fprintf(stderr, "\numma phases, epoch kernel minibatch 2, tower %d (cycles):", t);
            for (int i = 1; i < 48 && h[64 + t * 48 + i]; ++i) fprintf(stderr, " %lld", h[64 + t * 48 + i] - h[64 + t * 48 + i - 1]);
            fprintf(stderr, "\n   reduce phases (loads | combine + exchange + prefetch | partials / barrier | norm | Adam):");
            for (int i = 1; i < 8 && h[160 + t * 8 + i]; ++i) fprintf(stderr, " %lld", h[160 + t * 8 + i] - h[160 + t * 8 + i - 1]);
            fprintf(stderr, "\n");
        }
        if (c->persistent_epoch && getenv("PPO_UMMA_TIMELINE")) {  // per-CTA timeline of minibatch 2, ns since the earliest start
            std::vector<long long> tl(2 * (size_t)c->epoch_grid * 8);
            cudaMemcpy(tl.data(), c->umma_prof + 256, tl.size() * sizeof(long long), cudaMemcpyDeviceToHost);
            long long t0 = tl[0];
            for (size_t i = 0; i < tl.size(); i += 8) t0 = std::min(t0, tl[i]);
            fprintf(stderr, "timeline: start | weights staged | tiles | flushed | barrier 1 | reduce + Adam | barrier 3\n");
            for (int b = 0; b < 2 * c->epoch_grid; ++b) {
                fprintf(stderr, "cta %3d:", b);
                for (int i = 0; i < 7; ++i) fprintf(stderr, " %6lld", tl[(size_t)b * 8 + i] - t0);
                fprintf(stderr, "\n");
            }
        }
    }
    *avg_ms = ms / (float)iters;
    if (launches) *launches = (int)(c->ctr.kernel_launches - before);
    return PPO_OK;
}
