// abi_rollout.inl — part of libppo_core.so's single translation unit (included by ppo_core.cu, in this order): rollout: Runner::run on the device, the host-env protocol and its one-kernel form.
// ------------------------------------------------------------------------------------------------ rollout
static inline float* slab(ppo_core* c, int b, int t) {  // this rank's time-major slab, row t
    const bool global = (b == B_OBS || b == B_RETURNS || b == B_ACTIONS || b == B_VALUES || b == B_NEGLOGP);
    const size_t base = global ? (size_t)c->desc.rank * c->n_batch_local : 0;
    return c->buf[b] + (base + (size_t)t * c->desc.n_envs) * c->buf_w[b];
}

extern "C" int ppo_runner_reset(ppo_core* c, const float* raw_obs, ppo_mem mem) {
    if (!c || !raw_obs) return fail(PPO_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->desc.device));
    CU(cudaMemsetAsync(c->cur_dones, 0, c->desc.n_envs * sizeof(float), c->stream));  // dones{Zero} (runner.hpp:50)
    CU(cudaMemsetAsync(c->ret, 0, c->desc.n_envs * sizeof(float), c->stream));
    const float* d_raw = raw_obs;
    if (mem == PPO_HOST) {
        TRY(h2d(c, c->raw_obs, raw_obs, (size_t)c->desc.n_envs * c->d.O));
        d_raw = c->raw_obs;
    }
    TRY(vecnorm_device(c, d_raw, nullptr, nullptr, c->cur_obs, nullptr, nullptr, nullptr, nullptr, false));
    if (mem == PPO_HOST) CU(cudaStreamSynchronize(c->stream));
    return PPO_OK;
}

static int runner_act_device(ppo_core* c, int t) {
    c->gathered = false;  // this rank's slab changes: the other ranks' copies are stale until the next allgather
    PolicyArgs a{};
    a.obs = c->cur_obs; a.n = c->desc.n_envs; a.eps = nullptr; a.mode = 0;
    a.action = c->cur_actions;
    a.obs_store = slab(c, B_OBS, t); a.act_store = slab(c, B_ACTIONS, t); a.val_store = slab(c, B_VALUES, t);
    a.nlp_store = slab(c, B_NEGLOGP, t); a.dones_in = c->cur_dones; a.dones_store = slab(c, B_DONES, t);
    return launch_policy(c, a);
}

extern "C" int ppo_runner_act(ppo_core* c, int t, float* actions_out, ppo_mem mem) {
    if (!c || t < 0 || t >= c->desc.n_steps) return fail(PPO_ERR_INVALID, "ppo_runner_act: step %d out of range", t);
    CU(cudaSetDevice(c->desc.device));
    if (t == 0) TRY(prefetch_shuffle(c));
    TRY(runner_act_device(c, t));
    if (actions_out) {
        const size_t na = (size_t)c->desc.n_envs * c->d.A;
        if (mem == PPO_HOST) {
            TRY(d2h_staged_sync(c, actions_out, c->cur_actions, na));
        } else {
            CU(cudaMemcpyAsync(actions_out, c->cur_actions, na * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
        }
    }
    return PPO_OK;
}

extern "C" int ppo_runner_observe(ppo_core* c, int t, const float* raw_obs, const float* raw_rew, const float* done, ppo_mem mem) {
    if (!c || !raw_obs || !raw_rew || !done || t < 0 || t >= c->desc.n_steps) return fail(PPO_ERR_INVALID, "ppo_runner_observe: bad arguments");
    CU(cudaSetDevice(c->desc.device));
    const int N = c->desc.n_envs;
    const float *d_o = raw_obs, *d_r = raw_rew, *d_d = done;
    if (mem == PPO_HOST) {
        const StageCopy cp[3] = {{c->raw_obs, raw_obs, (size_t)N * c->d.O}, {c->raw_rew, raw_rew, (size_t)N}, {c->raw_done, done, (size_t)N}};
        TRY(h2d_staged(c, cp, 3));
        d_o = c->raw_obs; d_r = c->raw_rew; d_d = c->raw_done;
    }
    return vecnorm_device(c, d_o, d_r, d_d, c->cur_obs, c->nrew, c->cur_dones, slab(c, B_TRUE_REW, t), slab(c, B_UNNORM_REW, t), true);
}

extern "C" int ppo_runner_finish(ppo_core* c) {
    if (!c) return fail(PPO_ERR_INVALID, "core is NULL");
    CU(cudaSetDevice(c->desc.device));
    PolicyArgs a{};
    a.obs = c->cur_obs; a.n = c->desc.n_envs; a.mode = 1; a.value = c->last_values;  // model.value(obs) (runner.hpp:161-166)
    TRY(launch_policy(c, a));
    return launch_gae(c, slab(c, B_TRUE_REW, 0), slab(c, B_VALUES, 0), slab(c, B_DONES, 0), c->last_values, c->cur_dones,
                      c->desc.n_steps, c->desc.n_envs, c->desc.gamma, c->desc.lam, nullptr, slab(c, B_RETURNS, 0));
}

// launch arguments of rollout_persistent_kernel for the core's current state (synthetic env; the host-env mode adds its buffers)
static RolloutArgs make_rollout_args(ppo_core* c) {
    const ppo_core_desc& D = c->desc;
    RolloutArgs r{};
    r.d = c->d; r.params = c->params; r.n = D.n_envs; r.T = D.n_steps; r.tpc = c->roll_tpc;
    r.seed = D.seed; r.env_id0 = (uint32_t)D.env_offset; r.step_ctr = c->step_ctr; r.env = c->env; r.st = c->st; r.ret = c->ret;
    r.norm_gamma = D.norm_gamma; r.clip_obs = D.clip_obs; r.clip_rew = D.clip_reward; r.eps = D.norm_epsilon;
    r.norm_obs = D.norm_obs; r.norm_reward = D.norm_reward;
    r.upd_obs = D.training && D.norm_obs; r.upd_ret = D.training && D.norm_reward;
    r.partial = c->roll_partial; r.cur_obs = c->cur_obs; r.cur_dones = c->cur_dones; r.last_values = c->last_values;
    r.obs_store = slab(c, B_OBS, 0); r.act_store = slab(c, B_ACTIONS, 0); r.val_store = slab(c, B_VALUES, 0);
    r.nlp_store = slab(c, B_NEGLOGP, 0); r.dones_store = slab(c, B_DONES, 0); r.rew_store = slab(c, B_TRUE_REW, 0);
    r.urew_store = slab(c, B_UNNORM_REW, 0); r.ret_store = slab(c, B_RETURNS, 0);
    r.gamma = D.gamma; r.lam = D.lam;
    r.bar_ctr = c->sync_vars + SV_ROLL_FLAGS; r.bar_gen = c->sync_vars + SV_ROLL_GEN;
    r.n_global = D.n_envs * D.world_size;
    r.mbox = make_mailbox(c, false); r.mbox_seq = c->sync_vars + SV_MOM_SEQ; r.done_seq = c->sync_vars + SV_DONE_SEQ;
    r.off_obs = c->arena_off[B_OBS]; r.off_act = c->arena_off[B_ACTIONS]; r.off_val = c->arena_off[B_VALUES];
    r.off_nlp = c->arena_off[B_NEGLOGP]; r.off_ret = c->arena_off[B_RETURNS];
    r.row_off = (size_t)D.rank * c->n_batch_local;
    r.host_err = c->sync_vars + SV_ERR;
    return r;
}

static int prefetch_shuffle(ppo_core* c);

// The one-kernel host-env rollout pays one PCIe round trip per env step and reads the env's answer with SM loads from
// mapped host memory: a win while a step is latency-bound (C1: 1 env, 49 -> 17 us per env step), a loss once the
// observations are hundreds of KB per step (C3, 4096 envs: measured 326 us per step against 64 us with the copy engine).
static bool host_persistent_ok(const ppo_core* c) {
    return c->persistent_rollout && c->desc.world_size == 1 && getenv("PPO_DISABLE_HOST_PERSISTENT") == nullptr &&
           (c->desc.n_envs <= 512 || getenv("PPO_DISABLE_HOST_STAGED") == nullptr);
}
// up to 512 envs the kernel reads the env's answer with SM loads from mapped host memory; beyond, the copy engine moves it into a
// device staging buffer and a 4-byte copy behind it raises the flag the kernel polls (SM loads over PCIe: 326 us per step at 4096 envs)
static bool host_persistent_staged(const ppo_core* c) { return c->desc.n_envs > 512 && getenv("PPO_FORCE_HOST_MAPPED") == nullptr; }

// Host-env rollout as ONE persistent kernel (kernels_rollout.cuh, host-env mode): the kernel and this loop hand the actions
// and the env's answers back and forth through mapped pinned memory and two flags.
// direct_actions: the callback may read the actions where the kernel put them (no copy into `actions`); actions_all: the
// caller's pinned [n_steps][n_envs][A] array, written in place by the kernel when it is device-accessible (or NULL)
static int rollout_host_persistent(ppo_core* c, ppo_env_step_fn step, void* user, float* actions, bool direct_actions = false,
                                   float* actions_all = nullptr) {
    const ppo_core_desc& D = c->desc;
    const size_t N = (size_t)D.n_envs, O = (size_t)c->d.O, A = (size_t)c->d.A;
    const bool staged = host_persistent_staged(c);
    if (!c->hx_mem) {
        // [obs flag | act flags (grid) | actions N*A | obs N*O | rew N | done N | flag values 1 .. n_steps, abort], mapped + pinned
        const size_t words = 64 + (size_t)((c->roll_grid + 63) & ~63) + N * A + N * O + 2 * N + (size_t)D.n_steps + 2 + 4 + 2 * (A + O + 2);
        CU(cudaHostAlloc(reinterpret_cast<void**>(&c->hx_mem), words * sizeof(float), cudaHostAllocMapped | cudaHostAllocPortable));
        memset(c->hx_mem, 0, words * sizeof(float));
        CU(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->hx_dev), c->hx_mem, 0));
    }
    if (staged && !c->hx_stage) {
        CU(cudaMalloc(&c->hx_stage, (N * O + 2 * N + 64) * sizeof(float)));
        CU(cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking));
    }
    const size_t off_actf = 64, off_act = off_actf + (size_t)((c->roll_grid + 63) & ~63), off_obs = off_act + N * A, off_rew = off_obs + N * O,
                 off_done = off_rew + N, off_fval = off_done + N;
    unsigned* flag_vals = reinterpret_cast<unsigned*>(c->hx_mem) + off_fval;  // sources of the 4-byte flag copies (staged mode)
    for (int t = 0; t < D.n_steps; ++t) flag_vals[t] = (unsigned)t + 1u;
    flag_vals[D.n_steps] = PPO_HOST_ENV_ABORT;
    flag_vals[D.n_steps + 1] = 0u;
    // a single env: both directions as LL words (value, t + 1) in mapped memory, see RolloutArgs::h_act_ll
    const bool solo_ll = !staged && N == 1 && O + 2 <= 32 && getenv("PPO_DISABLE_HOST_LL") == nullptr;
    const size_t off_ll = (off_fval + (size_t)D.n_steps + 2 + 3) & ~(size_t)3;  // 16-byte aligned
    volatile uint64_t* act_ll = reinterpret_cast<volatile uint64_t*>(c->hx_mem + off_ll);
    volatile uint64_t* ans_ll = act_ll + A;
    if (solo_ll)
        for (size_t k = 0; k < A + O + 2; ++k) act_ll[k] = 0ull;  // sequence numbers restart at 1 with every rollout
    float* s_obs = c->hx_stage;
    unsigned* s_flag = staged ? reinterpret_cast<unsigned*>(c->hx_stage + N * O + 2 * N) : nullptr;
    volatile unsigned* obs_flag = reinterpret_cast<volatile unsigned*>(c->hx_mem);
    volatile unsigned* act_flag = reinterpret_cast<volatile unsigned*>(c->hx_mem) + off_actf;
    float* h_act = c->hx_mem + off_act;
    CU(cudaStreamSynchronize(c->stream));  // nothing of an earlier kernel may still look at the flags
    *obs_flag = 0u;
    for (int b = 0; b < c->roll_grid; ++b) act_flag[b] = 0u;
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
    if (staged) {
        CU(cudaMemcpyAsync(s_flag, flag_vals + D.n_steps + 1, sizeof(unsigned), cudaMemcpyHostToDevice, c->stream3));
        CU(cudaStreamSynchronize(c->stream3));
    }
    TRY(prefetch_shuffle(c));
    RolloutArgs r = make_rollout_args(c);
    r.h_actions = c->hx_dev + off_act;
    r.h_obs = c->hx_dev + off_obs; r.h_rew = c->hx_dev + off_rew; r.h_done = c->hx_dev + off_done;
    r.h_act_flag = reinterpret_cast<unsigned*>(c->hx_dev) + off_actf;
    r.h_obs_flag = reinterpret_cast<const unsigned*>(c->hx_dev);
    if (solo_ll) {
        r.h_act_ll = reinterpret_cast<uint2*>(c->hx_dev + off_ll);
        r.h_ans_ll = reinterpret_cast<const uint2*>(c->hx_dev + off_ll) + A;
    }
    if (staged) {  // observations through the copy engine (ONE API call per step: the kernel recognises the landed sectors, see h_sentinel);
                   // rewards / dones (8 bytes per env) stay in mapped memory
        r.h_obs = s_obs;
        r.h_obs_flag = s_flag;  // abort only
        r.h_sentinel = 1;
    }
    float* act_base = nullptr;  // host address of the kernel's action stores when they go straight into the caller's array
    if (direct_actions && actions_all) {
        void* dp = nullptr;
        if (cudaHostGetDevicePointer(&dp, actions_all, 0) == cudaSuccess && dp) {
            r.h_actions = static_cast<float*>(dp);
            r.h_act_stride = N * A;
            act_base = actions_all;
        } else {
            cudaGetLastError();
        }
    }
    void* kargs[] = {&r};
    CU(cudaLaunchCooperativeKernel((void*)rollout_persistent_kernel, dim3(c->roll_grid), dim3(R_NTH), kargs, c->roll_smem, c->stream));
    c->ctr.kernel_launches++;
    int st = PPO_OK;
    static const bool hx_prof = getenv("PPO_HOST_ROLLOUT_PROF") != nullptr;  // host-side split of an env step: wait | env callback | hand-over
    double prof_wait = 0.0, prof_env = 0.0, prof_push = 0.0;
    for (int t = 0; t < D.n_steps && st == PPO_OK; ++t) {
        // the CTAs' actions of step t
        const auto t0 = std::chrono::steady_clock::now();
        unsigned spins = 0;
        for (int b = 0; b < (solo_ll ? (int)A : c->roll_grid); ++b) {
            while (solo_ll ? (unsigned)(act_ll[b] >> 32) != (unsigned)t + 1u : act_flag[b] != (unsigned)t + 1u) {
                if (((++spins) & 0xfffffu) == 0u) {
                    if (cudaStreamQuery(c->stream) != cudaErrorNotReady) { st = fail(PPO_ERR_CUDA, "host-env rollout: the kernel ended at step %d: %s", t, cudaGetErrorString(cudaGetLastError())); break; }
                    if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(60)) { st = fail(PPO_ERR_CUDA, "host-env rollout: no actions from the device at step %d", t); break; }
                }
            }
            if (st != PPO_OK) break;
        }
        if (st != PPO_OK) break;
        __atomic_thread_fence(__ATOMIC_ACQUIRE);
        if (solo_ll)
            for (size_t j = 0; j < A; ++j) {
                const unsigned bits = (unsigned)act_ll[j];
                memcpy(h_act + j, &bits, sizeof(float));
            }
        const float* acts = (act_base && !solo_ll) ? act_base + (size_t)t * N * A : h_act;  // the kernel's stores into mapped host memory (posted PCIe writes)
        if (!direct_actions) {
            memcpy(actions, acts, N * A * sizeof(float));
            acts = actions;
        }
        c->ctr.d2h_bytes += N * A * sizeof(float);
        const float *o = nullptr, *rw = nullptr, *dn = nullptr;
        const auto t1 = std::chrono::steady_clock::now();
        if (step(user, t, acts, &o, &rw, &dn) != 0 || !o || !rw || !dn) {
            st = fail(PPO_ERR_INVALID, "ppo_runner_rollout_host: the env aborted at step %d", t);
            break;
        }
        const auto t2 = std::chrono::steady_clock::now();
        if (staged) {  // copy engine, then the flag behind the data on the same stream
            memcpy(c->hx_mem + off_rew, rw, N * sizeof(float));
            memcpy(c->hx_mem + off_done, dn, N * sizeof(float));
            __atomic_thread_fence(__ATOMIC_RELEASE);
            if (cudaMemcpyAsync(s_obs, o, N * O * sizeof(float), cudaMemcpyHostToDevice, c->stream3) != cudaSuccess) {
                st = fail(PPO_ERR_CUDA, "host-env rollout: H2D copy of step %d failed: %s", t, cudaGetErrorString(cudaGetLastError()));
                break;
            }
        } else if (solo_ll) {
            const uint64_t seq = (uint64_t)((unsigned)t + 1u) << 32;
            unsigned bits;
            for (size_t k = 0; k < O; ++k) {
                memcpy(&bits, o + k, sizeof(bits));
                ans_ll[k] = seq | bits;  // one aligned 8-byte store: value and sequence number become visible together
            }
            memcpy(&bits, rw, sizeof(bits));
            ans_ll[O] = seq | bits;
            memcpy(&bits, dn, sizeof(bits));
            ans_ll[O + 1] = seq | bits;
        } else {
            memcpy(c->hx_mem + off_obs, o, N * O * sizeof(float));
            memcpy(c->hx_mem + off_rew, rw, N * sizeof(float));
            memcpy(c->hx_mem + off_done, dn, N * sizeof(float));
        }
        c->ctr.h2d_bytes += N * (O + 2) * sizeof(float);
        __atomic_thread_fence(__ATOMIC_RELEASE);
        if (!staged) *obs_flag = (unsigned)t + 1u;
        if (hx_prof) {
            const auto t3 = std::chrono::steady_clock::now();
            prof_wait += std::chrono::duration<double, std::micro>(t1 - t0).count();
            prof_env += std::chrono::duration<double, std::micro>(t2 - t1).count();
            prof_push += std::chrono::duration<double, std::micro>(t3 - t2).count();
        }
    }
    if (hx_prof)
        fprintf(stderr, "[host-env rollout] per env step: wait for the actions %.1f us | env callback %.1f us | hand the answer over %.1f us\n",
                prof_wait / D.n_steps, prof_env / D.n_steps, prof_push / D.n_steps);
    if (staged && st == PPO_OK) {  // the env's arrays of the last step may be released when this call returns
        if (cudaStreamSynchronize(c->stream3) != cudaSuccess) st = fail(PPO_ERR_CUDA, "host-env rollout: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (st != PPO_OK) {
        char keep[1024];
        strncpy(keep, g_err, sizeof(keep));
        *obs_flag = PPO_HOST_ENV_ABORT;  // releases the kernel
        if (solo_ll)
            for (size_t k = 0; k < O + 2; ++k) ans_ll[k] = (uint64_t)PPO_HOST_ENV_ABORT << 32;
        __atomic_thread_fence(__ATOMIC_SEQ_CST);
        if (staged) {
            cudaMemcpyAsync(s_flag, flag_vals + D.n_steps, sizeof(unsigned), cudaMemcpyHostToDevice, c->stream3);
            cudaStreamSynchronize(c->stream3);
        }
        cudaStreamSynchronize(c->stream);
        strncpy(g_err, keep, sizeof(g_err));
        return st;
    }
    return PPO_OK;  // bootstrap value + GAE run at the kernel's end (asynchronous, like ppo_runner_finish)
}

extern "C" int ppo_runner_rollout_host(ppo_core* c, ppo_env_step_fn step, void* user, float* actions) {
    if (!c || !step || !actions) return fail(PPO_ERR_INVALID, "ppo_runner_rollout_host: NULL argument");
    CU(cudaSetDevice(c->desc.device));
    if (host_persistent_ok(c)) return rollout_host_persistent(c, step, user, actions);
    for (int t = 0; t < c->desc.n_steps; ++t) {
        TRY(ppo_runner_act(c, t, actions, PPO_HOST));
        const float *o = nullptr, *r = nullptr, *d = nullptr;
        if (step(user, t, actions, &o, &r, &d) != 0) return fail(PPO_ERR_INVALID, "ppo_runner_rollout_host: the env aborted at step %d", t);
        TRY(ppo_runner_observe(c, t, o, r, d, PPO_HOST));
    }
    return ppo_runner_finish(c);
}

namespace {
struct ReplayEnv {
    const float *obs, *rew, *done;
    float* actions_out;
    size_t no, n, na;
};
int replay_env_step(void* user, int t, const float* actions, const float** raw_obs, const float** raw_rew, const float** done) {
    ReplayEnv* e = static_cast<ReplayEnv*>(user);
    if (e->actions_out && actions != e->actions_out + (size_t)t * e->na) memcpy(e->actions_out + (size_t)t * e->na, actions, e->na * sizeof(float));
    *raw_obs = e->obs + (size_t)t * e->no;
    *raw_rew = e->rew + (size_t)t * e->n;
    *done = e->done + (size_t)t * e->n;
    return 0;
}
}  // namespace

extern "C" int ppo_runner_rollout_replay(ppo_core* c, const float* raw_obs, const float* raw_rew, const float* done, float* actions_out) {
    if (!c || !raw_obs || !raw_rew || !done) return fail(PPO_ERR_INVALID, "ppo_runner_rollout_replay: NULL argument");
    const size_t N = (size_t)c->desc.n_envs;
    ReplayEnv env{raw_obs, raw_rew, done, actions_out, N * c->d.O, N, N * c->d.A};
    const bool persistent = host_persistent_ok(c);
    if (actions_out && !persistent) {  // every step's actions land directly in their row of actions_out
        for (int t = 0; t < c->desc.n_steps; ++t) {
            float* a = actions_out + (size_t)t * env.na;
            TRY(ppo_runner_act(c, t, a, PPO_HOST));
            TRY(ppo_runner_observe(c, t, raw_obs + (size_t)t * env.no, raw_rew + (size_t)t * N, done + (size_t)t * N, PPO_HOST));
        }
        return ppo_runner_finish(c);
    }
    if (persistent) {  // the recorded env reads the actions where the kernel put them (one copy into actions_out, none without it)
        CU(cudaSetDevice(c->desc.device));
        return rollout_host_persistent(c, replay_env_step, &env, nullptr, true, actions_out);
    }
    std::vector<float> scratch(env.na);
    return ppo_runner_rollout_host(c, replay_env_step, &env, scratch.data());
}

// the synthetic env feeds action component k into state component k: it needs obs_dim == act_dim <= 32
static int synth_env_check(const ppo_core* c) {
    if (c->d.O != c->d.A || c->d.O > 32)
        return fail(PPO_ERR_UNSUPPORTED, "the synthetic env needs obs_dim == act_dim <= 32 (have %d/%d); use the host-env protocol", c->d.O, c->d.A);
    return PPO_OK;
}

extern "C" int ppo_synth_env_reset(ppo_core* c) {
    if (!c) return fail(PPO_ERR_INVALID, "core is NULL");
    TRY(synth_env_check(c));
    CU(cudaSetDevice(c->desc.device));
    LAUNCH(c, synth_env_reset_kernel, (c->desc.n_envs + 127) / 128, 128, 0, c->env, c->raw_obs);
    CU(cudaGetLastError());
    return ppo_runner_reset(c, c->raw_obs, PPO_DEVICE);
}

static int rollout_synthetic_enqueue(ppo_core* c) {
    const int N = c->desc.n_envs;
    for (int t = 0; t < c->desc.n_steps; ++t) {
        TRY(runner_act_device(c, t));
        LAUNCH(c, synth_env_step_kernel, (N + 127) / 128, 128, 0, c->env, c->cur_actions, c->raw_obs, c->raw_rew, c->raw_done);
        TRY(vecnorm_device(c, c->raw_obs, c->raw_rew, c->raw_done, c->cur_obs, c->nrew, c->cur_dones, slab(c, B_TRUE_REW, t),
                           slab(c, B_UNNORM_REW, t), true));
    }
    return ppo_runner_finish(c);
}

extern "C" int ppo_rollout_synthetic(ppo_core* c) {
    if (!c) return fail(PPO_ERR_INVALID, "core is NULL");
    TRY(synth_env_check(c));
    CU(cudaSetDevice(c->desc.device));
    TRY(prefetch_shuffle(c));  // the next update's permutations, on stream2, while this rollout runs
    if (c->persistent_rollout && fast_path(c)) {
        const ppo_core_desc& D = c->desc;
        RolloutArgs r = make_rollout_args(c);
        static long long* s_prof = nullptr;
        if (getenv("PPO_ROLLOUT_PROF") && !s_prof) {
            cudaMalloc(&s_prof, sizeof(long long) * 32);
            cudaMemset(s_prof, 0, sizeof(long long) * 32);
        }
        r.prof = s_prof;
        void* kargs[] = {&r};
        CU(cudaLaunchCooperativeKernel((void*)rollout_persistent_kernel, dim3(c->roll_grid), dim3(R_NTH), kargs, c->roll_smem, c->stream));
        c->ctr.kernel_launches++;
        c->gathered = D.world_size > 1;  // the kernel stored this rank's rows into every rank's buffers
        if (s_prof) {
            long long h[32];
            cudaStreamSynchronize(c->stream);
            cudaMemcpy(h, s_prof, sizeof(h), cudaMemcpyDeviceToHost);
            fprintf(stderr, "rollout phases (cycles):");
            for (int i = 1; i < 32 && h[i]; ++i) fprintf(stderr, " %lld", h[i] - h[i - 1]);
            fprintf(stderr, "\n");
        }
        return PPO_OK;
    }
    if (!rollout_graph_ok(c)) return rollout_synthetic_enqueue(c);
    // every launch argument of the rollout is a fixed device address (the Philox step counter lives on the device),
    // so the whole rollout is captured once and replayed; the training flag is baked into the captured launches
    ppo_core::EpochGraph& g = c->rollout_graph;
    if (!g.exec || g.bpow_slot != c->desc.training) {
        if (g.exec) {
            cudaGraphExecDestroy(g.exec);
            g.exec = nullptr;
        }
        const uint64_t k0 = c->ctr.kernel_launches;
        cudaGraph_t graph = nullptr;
        CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        c->wide_images_valid = false;  // the captured rollout must rebuild the W family's weight images itself (it is replayed after updates)
        const int st = rollout_synthetic_enqueue(c);
        const cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
        g.kernels = c->ctr.kernel_launches - k0;
        c->ctr.kernel_launches = k0;
        if (st != PPO_OK) {
            if (graph) cudaGraphDestroy(graph);
            return st;
        }
        if (ce != cudaSuccess) return fail(PPO_ERR_CUDA, "cudaStreamEndCapture(rollout) failed: %s", cudaGetErrorString(ce));
        const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) return fail(PPO_ERR_CUDA, "cudaGraphInstantiate(rollout) failed: %s", cudaGetErrorString(ie));
        g.bpow_slot = c->desc.training;
    }
    CU(cudaGraphLaunch(g.exec, c->stream));
    c->ctr.graph_launches++;
    c->ctr.kernel_launches += g.kernels;
    return PPO_OK;
}

static int buf_index(const char* name) {
    for (int i = 0; i < B_COUNT; ++i)
        if (name && strcmp(name, kBufNames[i]) == 0) return i;
    return -1;
}

extern "C" int ppo_rollout_get(ppo_core* c, const char* name, float* out, size_t cap) {
    const int b = buf_index(name);
    if (!c || !out || b < 0) return fail(PPO_ERR_INVALID, "ppo_rollout_get: unknown buffer '%s'", name ? name : "(null)");
    CU(cudaSetDevice(c->desc.device));
    const size_t n = (size_t)c->n_batch_local * c->buf_w[b];
    if (cap < n) return fail(PPO_ERR_INVALID, "buffer '%s' needs %zu floats, got %zu", name, n, cap);
    TRY(ensure_scratch(c, n));
    LAUNCH(c, export_flat_kernel, std::max(1, std::min(c->sm_count * 8, (int)((n + 255) / 256))), 256, 0, slab(c, b, 0),
           c->desc.n_steps, c->desc.n_envs, c->buf_w[b], c->scratch);
    CU(cudaGetLastError());
    TRY(d2h(c, out, c->scratch, n));
    CU(cudaStreamSynchronize(c->stream));
    return PPO_OK;
}

extern "C" int ppo_rollout_set(ppo_core* c, const char* name, const float* in, size_t count) {
    const int b = buf_index(name);
    if (!c || !in || b < 0) return fail(PPO_ERR_INVALID, "ppo_rollout_set: unknown buffer '%s'", name ? name : "(null)");
    CU(cudaSetDevice(c->desc.device));
    const size_t n = (size_t)c->n_batch_local * c->buf_w[b];
    if (count != n) return fail(PPO_ERR_INVALID, "buffer '%s' has %zu floats, got %zu", name, n, count);
    c->gathered = false;
    TRY(ensure_scratch(c, n));
    TRY(h2d(c, c->scratch, in, n));
    LAUNCH(c, import_flat_kernel, std::max(1, std::min(c->sm_count * 8, (int)((n + 255) / 256))), 256, 0, c->scratch,
           c->desc.n_steps, c->desc.n_envs, c->buf_w[b], slab(c, b, 0));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return PPO_OK;
}
