// abi_core.inl — part of libppo_core.so's single translation unit (included by ppo_core.cu, in this order): the core: state of a ppo_core, creation / destruction, allocation of the rollout buffer and the peer arena.
// ------------------------------------------------------------------------------------------------ the core
enum { B_OBS, B_RETURNS, B_DONES, B_ACTIONS, B_VALUES, B_NEGLOGP, B_TRUE_REW, B_UNNORM_REW, B_COUNT };
static inline bool is_global_buf(int b) { return b == B_OBS || b == B_RETURNS || b == B_ACTIONS || b == B_VALUES || b == B_NEGLOGP; }
static const char* const kBufNames[B_COUNT] = {"obs", "returns", "dones", "actions", "values", "neglogpacs",
                                               "true_rewards", "unnormalized_rewards"};

struct ppo_core {
    ppo_core_desc desc{};
    NetDims d{};
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    int tm = 64;          // tile size of the generic (T family) MLP kernels
    bool fused = false;   // F family usable: weights + one tile fit in shared memory, H1 % 4 == H2 % 4 == 0
    size_t fused_train_smem = 0, fused_policy_smem = 0;
    bool small = false;   // S family (thread per sample, registers): the reference's own [4,5] net with 18/18 obs/act
    bool umma = false;    // U family (tcgen05) train kernel usable: H1 == H2 == 64, obs/act 18/18
    bool wide = false;    // W family (tcgen05, layer-wise GEMMs over operand images): H1 == H2 in {128, 256, 512, 1024}
    wide::WideBufs wb{};
    void* wide_mem = nullptr;
    bool wide_images_valid = false;  // the weight images (and their scale table) were built from the current parameters
    int wide_cap = 0;     // capacity of the W-family buffers in tiles of 128 samples
    int max_train_grid = 0;
    int prof_train_grid = 0;
    long long* umma_prof = nullptr;  // PPO_UMMA_PROF=1: phase timestamps of the U-family train kernel
    int PS = 0;           // partial slab width = P + L_PAD, rounded up to whole float4

    float *params = nullptr, *adam_m = nullptr, *adam_v = nullptr, *bpow = nullptr;  // bpow: 2 slots x 2
    int bpow_slot = 0;

    NormStats st{};
    float* ret = nullptr;
    double *mom_partial = nullptr, *moments = nullptr;
    unsigned int* ticket = nullptr;
    int mom_grid = 0, mom_threads = 0;

    float *cur_obs = nullptr, *cur_dones = nullptr, *cur_actions = nullptr, *last_values = nullptr;
    float *raw_obs = nullptr, *raw_rew = nullptr, *raw_done = nullptr, *nrew = nullptr;
    uint32_t* step_ctr = nullptr;
    SynthEnv env{};

    int n_batch_local = 0, n_batch_global = 0, B_global = 0;
    float* buf[B_COUNT] = {};  // [world][T][Nl][w] slabs
    int buf_w[B_COUNT] = {};

    int *perm_dev = nullptr, *gather = nullptr;
    float2* mbstats = nullptr;
    float *partial = nullptr, *grad = nullptr, *loss_rows = nullptr, *loss_mean = nullptr, *gnorm = nullptr;
    double* sq_partial = nullptr;
    int n_sq_blocks = 0;
    bool perm_set = false;
    bool coop = false;        // fused cooperative reduce+Adam kernel usable (single GPU, grid co-resident)
    bool coop_big = false;    // ... in its many-chunks-per-block form
    int coop_grid = 0;
    bool use_graph = false;   // replay each epoch's launches as a CUDA graph
    struct EpochGraph {
        cudaGraphExec_t exec = nullptr;
        float lr = 0.f, cliprange = 0.f;
        int bpow_slot = -1;
        uint64_t kernels = 0;
        int flip = 0;  // beta-power slot parity change of one replay
    };
    std::vector<EpochGraph> graphs;
    EpochGraph rollout_graph;  // the whole synthetic-env rollout (n_steps x 4 kernels + bootstrap + GAE)
    bool small_epoch = false;         // S family, minibatches of one CTA (C1): all minibatches of an epoch in one single-CTA launch
    bool persistent_epoch = false;    // U family: all minibatches of an epoch in one cooperative launch
    int epoch_grid = 0;
    uint4* sq_ll = nullptr;           // sum-of-squares partials of the gradient step as LL words, [parity][block][block] (or NULL: grid barrier)
    int sq_ll_blocks = 0;
    bool persistent_rollout = false;  // R family: the whole rollout as one cooperative kernel
    int roll_grid = 0, roll_tpc = 0;
    size_t roll_smem = 0;
    double* roll_partial = nullptr;

    GlibcRand rng{1};
    // device-side std::random_shuffle (kernels_shuffle.cuh): generator window + work arrays for all epochs of an update
    bool gpu_shuffle = false, rng_on_device = false;
    uint32_t* rng_win = nullptr;
    shuf::Tables* shuf_tab = nullptr;
    int *sh_j = nullptr, *sh_cnt = nullptr, *sh_off = nullptr, *sh_cur = nullptr, *sh_list = nullptr, *sh_sigma = nullptr,
        *sh_perm = nullptr, *sh_gather = nullptr, *sh_btot = nullptr;
    float2* sh_mbstats = nullptr;
    uint32_t* win_pinned = nullptr;
    const int* cur_gather = nullptr;       // gather list / advantage statistics of the epoch being trained
    const float2* cur_mbstats = nullptr;
    EpochGraph update_graph;               // GPU-shuffle path: advantage statistics + all epochs of an update as one graph
    EpochGraph shuffle_graph;              // ... and the permutations of all its epochs as another: they do not depend on the
                                           // rollout, so the next update's are built on stream2 while the rollout runs
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_main = nullptr, ev_shuf = nullptr;
    bool shuffle_prefetched = false;       // sh_perm / sh_gather already hold the NEXT update's permutations (ev_shuf)
    uint32_t* rng_win_saved = nullptr;     // generator window before the prefetched draws (to undo an unused prefetch)
    std::vector<int> perm_host;
    int* perm_pinned = nullptr;  // [noptepochs][n_batch_global]
    float* stage = nullptr;      // pinned staging for pageable host buffers of the host-env protocol: two slots (step parity)
    size_t stage_floats = 0;
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};  // the H2D copies out of a slot have finished
    unsigned stage_ctr = 0;
    float* scratch = nullptr;    // device scratch for host-pointer calls
    size_t scratch_floats = 0;
    float* hx_mem = nullptr;     // host-env exchange of the persistent rollout kernel: flags, actions, obs / rew / done (mapped pinned)
    float* hx_dev = nullptr;     // ... its device address
    float* hx_stage = nullptr;   // ... with more than 512 envs the env's answer goes through the copy engine: device staging [obs | rew | done | flag]
    cudaStream_t stream3 = nullptr;  // ... on its own stream (the rollout kernel occupies c->stream while it polls)
    void* gae_ab = nullptr;      // per-(chunk, env) affine maps of the exact chunked GAE (gamma*lam near 1)
    size_t gae_ab_bytes = 0;

    ncclComm_t comm = nullptr;
    // peer-memory mailbox (multi-GPU): this rank's allocation, the IPC mappings of the peers', device-resident
    // barrier / sequence variables (sync_vars: see SV_*)
    unsigned char* mbox_mem = nullptr;
    unsigned char* mbox_peer[PPO_MAX_WORLD] = {};
    size_t mbox_bytes = 0, mbox_grad_off = 0, mbox_grad_slot = 0;
    size_t arena_off[8] = {};      // byte offsets of the five train-input buffers inside the arena
    size_t arena_sigma_off = 0;    // ... and of the per-epoch swap-chain results (sh_sigma) when the ranks share their construction
    bool gathered = false;         // the train inputs of every rank are already in place (persistent rollout, P2P stores)
    bool mbox_ready = false;
    unsigned* sync_vars = nullptr;
    ppo_counters ctr{};
};
// sync_vars layout: scalars first, then three barrier flag arrays of SV_MAXBLK words each
enum { SV_COOP_GEN, SV_ROLL_GEN, SV_EPOCH_GEN, SV_GRAD_SEQ, SV_MOM_SEQ, SV_ERR, SV_DONE_SEQ, SV_SHUF_SEQ, SV_SQ_SEQ, SV_SCALARS = 16, SV_MAXBLK = 2048,
       SV_COOP_FLAGS = SV_SCALARS, SV_ROLL_FLAGS = SV_COOP_FLAGS + SV_MAXBLK, SV_EPOCH_FLAGS = SV_ROLL_FLAGS + SV_MAXBLK,
       SV_COUNT = SV_EPOCH_FLAGS + SV_MAXBLK };

static PeerMailbox make_mailbox(const ppo_core* c, bool grads) {
    PeerMailbox m{};
    for (int r = 0; r < PPO_MAX_WORLD; ++r) m.base[r] = c->mbox_peer[r];
    m.rank = c->desc.rank;
    m.world = c->mbox_ready ? c->desc.world_size : 1;
    m.data_off = grads ? c->mbox_grad_off : PPO_MBOX_FLAG_BYTES;
    m.slot_bytes = grads ? c->mbox_grad_slot : PPO_MBOX_MOMENT_SLOT;
    m.err = c->sync_vars + SV_ERR;
    return m;
}
// single GPU, or multi-GPU with the peer mailboxes mapped: the persistent / cooperative kernels carry the exchanges
static inline bool fast_path(const ppo_core* c) { return c->desc.world_size == 1 || c->mbox_ready; }
// CUDA graphs hold only our own kernels.  With more than one rank that requires every exchange of the captured work
// to run through the peer mailboxes inside those kernels; the per-step kernels of the other shapes call NCCL.
static inline bool rollout_graph_ok(const ppo_core* c) { return c->use_graph && c->desc.world_size == 1; }
static inline bool update_graph_ok(const ppo_core* c) {
    return c->use_graph && fast_path(c) && (c->desc.world_size == 1 || c->coop || c->persistent_epoch);
}

#define LAUNCH(core, kernel, grid, block, smem, ...)                               \
    do {                                                                           \
        kernel<<<(grid), (block), (smem), (core)->stream>>>(__VA_ARGS__);          \
        (core)->ctr.kernel_launches++;                                             \
    } while (0)

static int ensure_scratch(ppo_core* c, size_t floats) {
    if (floats <= c->scratch_floats) return PPO_OK;
    if (c->scratch) {
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaFree(c->scratch));
        c->scratch = nullptr;
    }
    CU(cudaMalloc(&c->scratch, floats * sizeof(float)));
    c->scratch_floats = floats;
    return PPO_OK;
}
static int ensure_stage(ppo_core* c, size_t floats) {
    if (floats <= c->stage_floats) return PPO_OK;
    if (c->stage) {
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaFreeHost(c->stage));
        c->stage = nullptr;
    }
    CU(cudaMallocHost(&c->stage, floats * sizeof(float)));
    c->stage_floats = floats;
    return PPO_OK;
}

// copy helpers honouring ppo_mem: returns a device pointer for an input / stages an output
static int h2d(ppo_core* c, float* dst, const float* src, size_t n) {
    CU(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    c->ctr.h2d_bytes += n * sizeof(float);
    return PPO_OK;
}
static int d2h(ppo_core* c, float* dst, const float* src, size_t n) {
    CU(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    c->ctr.d2h_bytes += n * sizeof(float);
    return PPO_OK;
}

// Host buffers of the per-step host-env protocol (Runner::run with host physics, runner.hpp:56-157).  Pinned memory is
// DMA'd in place.  Pageable memory is staged through the core's own pinned double buffer: the caller's memcpy into slot
// (step & 1) overlaps the DMA still reading slot (step - 1) & 1, and the call returns without waiting for the copy
// (cudaMemcpyAsync from pageable memory would block until the driver has staged it).
static bool host_ptr_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}
struct StageCopy { float* dst; const float* src; size_t n; };
static int h2d_staged(ppo_core* c, const StageCopy* cp, int ncp) {
    bool all_pinned = true;
    size_t total = 0;
    for (int i = 0; i < ncp; ++i) {
        all_pinned = all_pinned && host_ptr_pinned(cp[i].src);
        total += cp[i].n;
    }
    if (all_pinned) {
        for (int i = 0; i < ncp; ++i) TRY(h2d(c, cp[i].dst, cp[i].src, cp[i].n));
        return PPO_OK;
    }
    if (2 * total > c->stage_floats) {
        TRY(ensure_stage(c, 2 * total));
        for (int k = 0; k < 2; ++k)
            if (!c->stage_ev[k]) CU(cudaEventCreateWithFlags(&c->stage_ev[k], cudaEventDisableTiming));
    }
    const unsigned slot = c->stage_ctr++ & 1u;
    CU(cudaEventSynchronize(c->stage_ev[slot]));  // copies issued from this slot two steps ago (a fresh event is complete)
    float* p = c->stage + (size_t)slot * (c->stage_floats / 2);
    for (int i = 0; i < ncp; ++i) {
        memcpy(p, cp[i].src, cp[i].n * sizeof(float));
        TRY(h2d(c, cp[i].dst, p, cp[i].n));
        p += cp[i].n;
    }
    CU(cudaEventRecord(c->stage_ev[slot], c->stream));
    return PPO_OK;
}
// device -> host buffer, complete on return
static int d2h_staged_sync(ppo_core* c, float* dst, const float* src, size_t n) {
    if (host_ptr_pinned(dst)) {
        TRY(d2h(c, dst, src, n));
        CU(cudaStreamSynchronize(c->stream));
        return PPO_OK;
    }
    if (2 * n > c->stage_floats) {
        TRY(ensure_stage(c, 2 * n));
        for (int k = 0; k < 2; ++k)
            if (!c->stage_ev[k]) CU(cudaEventCreateWithFlags(&c->stage_ev[k], cudaEventDisableTiming));
    }
    // the stream is synchronised below, so every earlier copy out of the staging slots has finished when we reuse one
    CU(cudaStreamSynchronize(c->stream));
    TRY(d2h(c, c->stage, src, n));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(dst, c->stage, n * sizeof(float));
    return PPO_OK;
}

extern "C" int ppo_core_desc_default(ppo_core_desc* d) {
    if (!d) return fail(PPO_ERR_INVALID, "desc is NULL");
    memset(d, 0, sizeof(*d));
    d->abi_version = PPO_CORE_ABI_VERSION;
    d->obs_dim = 18; d->act_dim = 18; d->hidden1 = 4; d->hidden2 = 5;
    d->n_envs = 1; d->n_steps = 2048; d->nminibatches = 32; d->noptepochs = 10;
    d->gamma = 0.99f; d->lam = 0.95f;
    d->ent_coef = 0.0007160293171182275f; d->vf_coef = 0.5f; d->max_grad_norm = 0.5f;
    d->adam_beta1 = 0.9f; d->adam_beta2 = 0.999f; d->adam_epsilon = 1e-5f;
    d->norm_obs = 1; d->norm_reward = 1; d->training = 1;
    d->clip_obs = 10.f; d->clip_reward = 10.f; d->norm_gamma = 0.99f; d->norm_epsilon = 1e-8f;
    d->seed = 0; d->rank = 0; d->world_size = 1; d->env_offset = 0; d->n_envs_global = 0;
    return PPO_OK;
}

extern "C" int ppo_meta_parse(const char* path, ppo_meta_info* info, float* params_out, size_t cap) {
    if (!path || !info) return fail(PPO_ERR_INVALID, "ppo_meta_parse: NULL argument");
    MetaGraph g;
    const std::string err = parse_meta_txt(path, g);
    if (!err.empty()) return fail(PPO_ERR_IO, "%s", err.c_str());
    NetDims d;
    d.init(g.obs_dim, g.act_dim, g.hidden1, g.hidden2);
    info->obs_dim = g.obs_dim; info->act_dim = g.act_dim; info->hidden1 = g.hidden1; info->hidden2 = g.hidden2;
    info->ent_coef = g.ent_coef; info->vf_coef = g.vf_coef; info->max_grad_norm = g.clip_norm;
    info->adam_beta1 = g.beta1; info->adam_beta2 = g.beta2; info->adam_epsilon = g.adam_eps;
    info->n_params_trainable = d.P; info->n_params_total = d.Pq;
    if (params_out) {
        if (cap < (size_t)d.Pq) return fail(PPO_ERR_INVALID, "params_out holds %zu floats, graph has %d", cap, d.Pq);
        for (int t = 0; t < kNumTensors; ++t) {
            const MetaTensor& mt = g.tensors[kTensorNames[t]];
            if ((int)mt.data.size() != d.off[t + 1] - d.off[t]) return fail(PPO_ERR_IO, "tensor %s has unexpected size", kTensorNames[t]);
            memcpy(params_out + d.off[t], mt.data.data(), mt.data.size() * sizeof(float));
        }
    }
    return PPO_OK;
}

// the largest dynamic shared memory a kernel may ask for: the device's opt-in maximum minus the kernel's static shared memory
template <class K>
static int max_dynamic_smem(K kernel, size_t max_smem) {
    cudaFuncAttributes fa{};
    if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) {
        cudaGetLastError();
        return (int)max_smem;
    }
    return (int)(max_smem - std::min(max_smem, (size_t)fa.sharedSizeBytes));
}

template <int TM>
static int set_smem_attrs(size_t max_smem) {
    // The attribute is per function and per device, i.e. shared by every core of the process: always raise it to the
    // device's opt-in maximum, so that a core created later with smaller hidden sizes (EnvNormalize's private [4,5] core
    // beside a [64,64] PPO2 core) cannot lower the limit under a live core.  What a launch uses is its own smem argument.
    CU(cudaFuncSetAttribute(train_tile_kernel<TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dynamic_smem(train_tile_kernel<TM>, max_smem)));
    CU(cudaFuncSetAttribute(policy_tile_kernel<TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dynamic_smem(policy_tile_kernel<TM>, max_smem)));
    return PPO_OK;
}

extern "C" void ppo_core_destroy(ppo_core* c) {
    if (!c) return;
    cudaSetDevice(c->desc.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (int r = 0; r < PPO_MAX_WORLD; ++r)
        if (c->mbox_peer[r] && r != c->desc.rank) cudaIpcCloseMemHandle(c->mbox_peer[r]);
    if (c->mbox_mem) cudaFree(c->mbox_mem);
    if (c->gae_ab) cudaFree(c->gae_ab);
    if (c->hx_mem) cudaFreeHost(c->hx_mem);
    if (c->hx_stage) cudaFree(c->hx_stage);
    if (c->stream3) cudaStreamDestroy(c->stream3);
    if (c->wide_mem) cudaFree(c->wide_mem);
    if (c->sync_vars) cudaFree(c->sync_vars);
    if (c->sq_ll) cudaFree(c->sq_ll);
    if (c->umma_prof) cudaFree(c->umma_prof);
    for (auto& g : c->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    if (c->rollout_graph.exec) cudaGraphExecDestroy(c->rollout_graph.exec);
    if (c->update_graph.exec) cudaGraphExecDestroy(c->update_graph.exec);
    if (c->stream2) cudaStreamSynchronize(c->stream2);
    if (c->shuffle_graph.exec) cudaGraphExecDestroy(c->shuffle_graph.exec);
    if (c->ev_main) cudaEventDestroy(c->ev_main);
    if (c->ev_shuf) cudaEventDestroy(c->ev_shuf);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->rng_win_saved) cudaFree(c->rng_win_saved);
    if (c->win_pinned) cudaFreeHost(c->win_pinned);
    void* dev_ptrs[] = {c->params, c->adam_m, c->adam_v, c->bpow, c->st.obs_mean, c->st.obs_var, c->st.obs_count,
                        c->st.ret_mean, c->st.ret_var, c->st.ret_count, c->ret, c->mom_partial, c->moments, c->ticket,
                        c->cur_obs, c->cur_dones, c->cur_actions, c->last_values, c->raw_obs, c->raw_rew, c->raw_done,
                        c->nrew, c->step_ctr, c->env.state, c->env.t_env, c->env.resets, c->perm_dev, c->gather,
                        c->mbstats, c->partial, c->grad, c->loss_rows, c->loss_mean, c->gnorm, c->sq_partial, c->scratch,
                        c->roll_partial, c->rng_win, c->shuf_tab, c->sh_j, c->sh_cnt, c->sh_off, c->sh_cur, c->sh_list,
                        c->arena_sigma_off ? nullptr : c->sh_sigma, c->sh_perm, c->sh_gather, c->sh_btot, c->sh_mbstats};
    for (void* p : dev_ptrs)
        if (p) cudaFree(p);
    for (int i = 0; i < B_COUNT; ++i)
        if (c->buf[i] && !(c->mbox_mem && is_global_buf(i))) cudaFree(c->buf[i]);
    if (c->perm_pinned) cudaFreeHost(c->perm_pinned);
    if (c->stage) cudaFreeHost(c->stage);
    for (int k = 0; k < 2; ++k)
        if (c->stage_ev[k]) cudaEventDestroy(c->stage_ev[k]);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

static int ensure_wide(ppo_core* c, int tiles);
static int prefetch_shuffle(ppo_core* c);
static int core_alloc(ppo_core* c) {
    const ppo_core_desc& D = c->desc;
    const NetDims& d = c->d;
    const int N = D.n_envs, O = d.O, A = d.A, T = D.n_steps, W = D.world_size;
    auto zalloc = [&](void** p, size_t bytes) -> int {
        CU(cudaMalloc(p, bytes));
        CU(cudaMemsetAsync(*p, 0, bytes, c->stream));
        return PPO_OK;
    };
#define ZA(ptr, count) TRY(zalloc(reinterpret_cast<void**>(&(ptr)), sizeof(*(ptr)) * (size_t)(count)))
    ZA(c->params, d.Pq); ZA(c->adam_m, d.P); ZA(c->adam_v, d.P); ZA(c->bpow, 4);
    ZA(c->st.obs_mean, O); ZA(c->st.obs_var, O); ZA(c->st.obs_count, 1);
    ZA(c->st.ret_mean, 1); ZA(c->st.ret_var, 1); ZA(c->st.ret_count, 1);
    ZA(c->ret, N);
    c->mom_threads = O * std::max(1, 256 / O);
    c->mom_grid = std::max(1, std::min(c->sm_count * 2, (int)(((size_t)N * O + c->mom_threads * 8 - 1) / (c->mom_threads * 8))));
    ZA(c->mom_partial, (size_t)c->mom_grid * 2 * (O + 1)); ZA(c->moments, 2 * (O + 1) + 1); ZA(c->ticket, 1);
    ZA(c->cur_obs, (size_t)N * O); ZA(c->cur_dones, N); ZA(c->cur_actions, (size_t)N * A); ZA(c->last_values, N);
    ZA(c->raw_obs, (size_t)N * O); ZA(c->raw_rew, N); ZA(c->raw_done, N); ZA(c->nrew, N);
    ZA(c->step_ctr, 1);
    ZA(c->env.state, (size_t)N * O); ZA(c->env.t_env, N); ZA(c->env.resets, N);
    c->env.seed = D.seed ^ 0x1234ull; c->env.env_id0 = (uint32_t)D.env_offset; c->env.n = N; c->env.D = O;
    c->n_batch_local = N * T;
    c->n_batch_global = c->n_batch_local * W;
    c->B_global = c->n_batch_global / D.nminibatches;
    const int widths[B_COUNT] = {O, 1, 1, A, 1, 1, 1, 1};
    c->PS = (d.P + L_PAD + 3) & ~3;  // rows of the slab buffer stay 16-byte aligned (float4 loads of the column reduce)
    if (W > 1) {
        // one arena per rank, IPC-mapped by every peer: [mailbox flags | moment slots | gradient slots | the five train inputs].
        // The persistent rollout kernel stores its rows straight into every rank's copy (NVLink P2P), so the buffers are
        // already "allgathered" when the rollout ends.
        if (W > PPO_MAX_WORLD) return fail(PPO_ERR_UNSUPPORTED, "world_size %d > %d", W, PPO_MAX_WORLD);
        c->mbox_grad_off = PPO_MBOX_FLAG_BYTES + 2 * (size_t)PPO_MAX_WORLD * PPO_MBOX_MOMENT_SLOT;
        c->mbox_grad_slot = (((size_t)c->PS * sizeof(uint2)) + 255) & ~(size_t)255;  // LL words: (value, seq)
        size_t off = c->mbox_grad_off + 2 * (size_t)W * c->mbox_grad_slot;
        for (int i = 0; i < B_COUNT; ++i) {
            if (!is_global_buf(i)) continue;
            c->arena_off[i] = off;
            off += (((size_t)c->n_batch_global * widths[i] * sizeof(float)) + 255) & ~(size_t)255;
        }
        // permutations: rank r resolves the swap chains of epochs r, r + W, ... and stores them into every rank's sigma array
        {
            const long long E = D.noptepochs, nbg = c->n_batch_global;
            if (E >= 1 && nbg >= 2 && E * (nbg - 1) < 0x7fffffffLL) {
                c->arena_sigma_off = off;
                off += (((size_t)E * nbg * sizeof(int)) + 255) & ~(size_t)255;
            }
        }
        c->mbox_bytes = off;
        CU(cudaMalloc(&c->mbox_mem, c->mbox_bytes));
        CU(cudaMemsetAsync(c->mbox_mem, 0, c->mbox_bytes, c->stream));
        c->mbox_peer[D.rank] = c->mbox_mem;
    }
    for (int i = 0; i < B_COUNT; ++i) {
        c->buf_w[i] = widths[i];
        // only the five train inputs are global ([rank][t][env_local][w] slabs); the others stay local-sized
        if (W > 1 && is_global_buf(i)) c->buf[i] = reinterpret_cast<float*>(c->mbox_mem + c->arena_off[i]);
        else ZA(c->buf[i], (size_t)(is_global_buf(i) ? c->n_batch_global : c->n_batch_local) * widths[i]);
    }
    ZA(c->perm_dev, c->n_batch_global); ZA(c->gather, c->n_batch_global);
    {
        const long long E = D.noptepochs, nbg = c->n_batch_global;
        c->gpu_shuffle = E >= 1 && nbg >= 2 && E * (nbg - 1) < 0x7fffffffLL && getenv("PPO_DISABLE_GPU_SHUFFLE") == nullptr;
        if (c->gpu_shuffle) {
            const size_t en = (size_t)E * nbg, en1 = (size_t)E * (nbg + 1);
            const int nb = (int)((nbg + 1 + shuf::SCAN_TILE - 1) / shuf::SCAN_TILE);
            ZA(c->rng_win, 31); ZA(c->shuf_tab, 1);
            ZA(c->sh_j, en); ZA(c->sh_cnt, en1); ZA(c->sh_off, en1); ZA(c->sh_cur, en1); ZA(c->sh_list, en);
            if (c->arena_sigma_off) c->sh_sigma = reinterpret_cast<int*>(c->mbox_mem + c->arena_sigma_off);
            else ZA(c->sh_sigma, en);
            ZA(c->sh_perm, en); ZA(c->sh_gather, en); ZA(c->sh_btot, (size_t)E * nb); ZA(c->sh_mbstats, (size_t)E * D.nminibatches);
            CU(cudaMallocHost(&c->win_pinned, 31 * sizeof(uint32_t)));
            static shuf::Tables host_tab;
            static bool host_tab_ready = false;
            if (!host_tab_ready) {
                shuf::build_tables(host_tab);
                host_tab_ready = true;
            }
            CU(cudaMemcpyAsync(c->shuf_tab, &host_tab, sizeof(host_tab), cudaMemcpyHostToDevice, c->stream));
        }
    }
    ZA(c->mbstats, D.nminibatches);
    c->max_train_grid = c->sm_count * 2;
    ZA(c->partial, (size_t)c->max_train_grid * c->PS); ZA(c->grad, c->PS);
    c->n_sq_blocks = (c->PS + 255) / 256;
    ZA(c->sq_partial, c->n_sq_blocks);
    ZA(c->loss_rows, (size_t)D.noptepochs * D.nminibatches * 5 + 5); ZA(c->loss_mean, 5); ZA(c->gnorm, 1);
#undef ZA
    CU(cudaMallocHost(&c->perm_pinned, sizeof(int) * (size_t)c->n_batch_global * std::max(1, D.noptepochs)));
    c->perm_host.resize(c->n_batch_global);
    // RunningStatistics(): mean 0, var 1, count = (double)1e-6f  (running_statistics.hpp:17-20)
    std::vector<float> ones(O, 1.f);
    const double cnt = (double)1e-6f;
    const float one = 1.f;
    CU(cudaMemcpyAsync(c->st.obs_var, ones.data(), O * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->st.ret_var, &one, sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->st.obs_count, &cnt, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->st.ret_count, &cnt, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    // beta powers start at beta (GRAPH:25426,25579)
    const float bp[4] = {D.adam_beta1, D.adam_beta2, D.adam_beta1, D.adam_beta2};
    CU(cudaMemcpyAsync(c->bpow, bp, sizeof(bp), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return PPO_OK;
}

extern "C" int ppo_core_create(const ppo_core_desc* desc, ppo_core** out) {
    if (!desc || !out) return fail(PPO_ERR_INVALID, "ppo_core_create: NULL argument");
    if (desc->abi_version != PPO_CORE_ABI_VERSION) return fail(PPO_ERR_INVALID, "ABI version mismatch: header %d, library %d", desc->abi_version, PPO_CORE_ABI_VERSION);
    // the reference's envs: closed loop 18/18 (36/18 with velocities, hexapod_closed_loop_env.hpp:20,61-72), open loop 1/18
    // (hexapod_env.hpp:226-238).  18/18 selects the specialised S / U / W families; other widths run on the generic F / T families.
    if (desc->obs_dim < 1 || desc->obs_dim > 64 || desc->act_dim < 1 || desc->act_dim > 64)
        return fail(PPO_ERR_UNSUPPORTED, "obs_dim/act_dim %d/%d unsupported (need 1..64 each)", desc->obs_dim, desc->act_dim);
    if (desc->hidden1 < 1 || desc->hidden2 < 1 || desc->hidden1 > 1024 || desc->hidden2 > 1024)
        return fail(PPO_ERR_UNSUPPORTED, "hidden sizes [%d,%d] out of range 1..1024", desc->hidden1, desc->hidden2);
    if (desc->n_envs < 1 || desc->n_steps < 1 || desc->nminibatches < 1 || desc->noptepochs < 0)
        return fail(PPO_ERR_INVALID, "n_envs, n_steps, nminibatches must be >= 1");
    if (desc->world_size < 1 || desc->rank < 0 || desc->rank >= desc->world_size)
        return fail(PPO_ERR_INVALID, "bad rank/world_size %d/%d", desc->rank, desc->world_size);
    const long nbg = (long)desc->n_envs * desc->n_steps * desc->world_size;
    if (nbg % desc->nminibatches != 0)  // assert((n_batch % nminibatches) == 0), ppo2.hpp:265
        return fail(PPO_ERR_INVALID, "n_batch %ld not divisible by nminibatches %d", nbg, desc->nminibatches);
    if ((nbg / desc->nminibatches) % desc->world_size != 0)
        return fail(PPO_ERR_INVALID, "minibatch size %ld not divisible by world_size %d", nbg / desc->nminibatches, desc->world_size);
    if (nbg > 0x7fffffffL) return fail(PPO_ERR_UNSUPPORTED, "n_batch %ld exceeds int32 (the reference uses int indices)", nbg);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PPO_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    }
    if (desc->device < 0 || desc->device >= ndev) return fail(PPO_ERR_INVALID, "device %d out of range (have %d)", desc->device, ndev);
    CU(cudaSetDevice(desc->device));
    ppo_core* c = new ppo_core();
    c->desc = *desc;
    if (c->desc.n_envs_global <= 0) c->desc.n_envs_global = desc->n_envs * desc->world_size;
    if (c->desc.world_size > 1 && c->desc.env_offset == 0) c->desc.env_offset = desc->rank * desc->n_envs;
    c->d.init(desc->obs_dim, desc->act_dim, desc->hidden1, desc->hidden2);
    int st = PPO_OK;
    do {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, desc->device) != cudaSuccess) { st = fail(PPO_ERR_CUDA, "cudaGetDeviceProperties failed"); break; }
        c->sm_count = prop.multiProcessorCount;
        const size_t max_smem = prop.sharedMemPerBlockOptin;
        if (train_smem_floats<64>(c->d) * sizeof(float) <= max_smem) c->tm = 64;
        else if (train_smem_floats<32>(c->d) * sizeof(float) <= max_smem) c->tm = 32;
        else { st = fail(PPO_ERR_UNSUPPORTED, "hidden sizes [%d,%d] need more shared memory than the device has", desc->hidden1, desc->hidden2); break; }
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { st = fail(PPO_ERR_CUDA, "cudaStreamCreate failed"); break; }
        st = set_smem_attrs<64>(max_smem);
        if (st == PPO_OK) st = set_smem_attrs<32>(max_smem);
        if (st != PPO_OK) break;
        {
            FLayout lt, lp;
            lt.init(c->d, F_TM_TRAIN, true);
            lp.init(c->d, F_TM_POLICY, false);
            c->fused_train_smem = (size_t)lt.total * sizeof(float);
            c->fused_policy_smem = (size_t)lp.total * sizeof(float);
            c->fused = (c->d.H1 % 4 == 0) && (c->d.H2 % 4 == 0) && c->fused_train_smem <= max_smem && c->fused_policy_smem <= max_smem &&
                       getenv("PPO_DISABLE_FUSED") == nullptr;
            if (c->fused) {
                if (cudaFuncSetAttribute(train_fused_kernel<F_TM_TRAIN, F_NT_TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dynamic_smem(train_fused_kernel<F_TM_TRAIN, F_NT_TRAIN>, max_smem)) != cudaSuccess ||
                    cudaFuncSetAttribute(policy_fused_kernel<F_TM_POLICY, F_NT_POLICY>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dynamic_smem(policy_fused_kernel<F_TM_POLICY, F_NT_POLICY>, max_smem)) != cudaSuccess) {
                    st = fail(PPO_ERR_CUDA, "cudaFuncSetAttribute(fused kernels) failed: %s", cudaGetErrorString(cudaGetLastError()));
                    break;
                }
            }
        }
        c->umma = c->d.H1 == umma::HID && c->d.H2 == umma::HID && c->d.O == 18 && c->d.A == 18 && umma::SMEM_BYTES <= max_smem &&
                  prop.major == 10 && getenv("PPO_DISABLE_UMMA") == nullptr && getenv("PPO_DISABLE_FUSED") == nullptr;
        if (c->umma && cudaFuncSetAttribute(umma::train_umma_kernel<18, 18, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)umma::SMEM_BYTES) != cudaSuccess) {
            st = fail(PPO_ERR_CUDA, "cudaFuncSetAttribute(train_umma_kernel) failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        c->small = c->d.O == 18 && c->d.A == 18 && c->d.H1 == 4 && c->d.H2 == 5 && getenv("PPO_DISABLE_SMALL") == nullptr;
        {
            const int H = c->d.H1;
            c->wide = c->d.H1 == c->d.H2 && (H == 128 || H == 256 || H == 512 || H == 1024) && c->d.O == 18 && c->d.A == 18 &&
                      wide::GEMM_SMEM <= max_smem && prop.major == 10 && getenv("PPO_DISABLE_WIDE") == nullptr;
            if (c->wide && cudaFuncSetAttribute(wide::wgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide::GEMM_SMEM) != cudaSuccess) {
                st = fail(PPO_ERR_CUDA, "cudaFuncSetAttribute(wgemm_kernel) failed: %s", cudaGetErrorString(cudaGetLastError()));
                break;
            }
        }
        st = core_alloc(c);
        if (st != PPO_OK) break;
        if (c->wide) {
            const long per_rank_mb = nbg / desc->nminibatches / desc->world_size;
            st = ensure_wide(c, (int)((std::max<long>(per_rank_mb, desc->n_envs) + wide::TM - 1) / wide::TM));
            if (st != PPO_OK) break;
        }
        if (c->umma && getenv("PPO_UMMA_PROF")) {
            if (cudaMalloc(&c->umma_prof, sizeof(long long) * 4096) != cudaSuccess) { st = fail(PPO_ERR_CUDA, "cudaMalloc(umma_prof) failed"); break; }
            cudaMemset(c->umma_prof, 0, sizeof(long long) * 4096);
        }
        {
            int per_sm = 0, coop_ok = 0;
            cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, desc->device);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, grad_reduce_adam_coop_kernel, 256, 0);
            {
                int per_big = 0;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_big, grad_reduce_adam_big_kernel, 256, 0);
                per_sm = std::min(per_sm, per_big);
            }
            // cooperative reduce(+allreduce)+Adam: blocks own 64-column chunks, up to RA_MAXJ chunks each; multi-GPU runs
            // use it once the peer mailboxes are mapped (fast_path), with one mailbox channel per block
            const int nchunks = (c->PS + 63) / 64;
            c->coop_grid = std::min(nchunks, std::min(per_sm * c->sm_count, PPO_MBOX_CHANNELS - 1));
            c->coop = coop_ok && c->coop_grid > 0 && getenv("PPO_DISABLE_COOP") == nullptr;
            c->coop_big = nchunks > c->coop_grid * RA_MAXJ;  // long parameter vectors (W family): grad_reduce_adam_big_kernel
            if (c->coop && c->coop_grid > c->n_sq_blocks) {  // sq_partial is sized for 256-column blocks
                cudaFree(c->sq_partial);
                c->sq_partial = nullptr;
                if (cudaMalloc(&c->sq_partial, sizeof(double) * c->coop_grid) != cudaSuccess) { st = fail(PPO_ERR_CUDA, "cudaMalloc(sq_partial) failed"); break; }
            }
            if (cudaMalloc(&c->sync_vars, sizeof(unsigned) * SV_COUNT) != cudaSuccess) { st = fail(PPO_ERR_CUDA, "cudaMalloc(sync_vars) failed"); break; }
            cudaMemset(c->sync_vars, 0, sizeof(unsigned) * SV_COUNT);
            if (c->umma && coop_ok && getenv("PPO_DISABLE_PERSISTENT") == nullptr) {
                const int per_rank = (int)(nbg / desc->nminibatches / desc->world_size);
                const int ntiles = (per_rank + umma::TM - 1) / umma::TM;
                const int grid = std::max(1, std::min(ntiles, c->sm_count / 2));
                int per = 0;
                if (cudaFuncSetAttribute(umma::train_umma_kernel<18, 18, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)umma::SMEM_BYTES) == cudaSuccess)
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, umma::train_umma_kernel<18, 18, 1>, umma::NTH, umma::SMEM_BYTES);
                else
                    cudaGetLastError();
                if (per > 0 && 2 * grid <= per * c->sm_count && nchunks <= 2 * grid * RA_MAXJ && 2 * grid <= PPO_MBOX_CHANNELS - 1) {
                    c->persistent_epoch = true;
                    c->epoch_grid = grid;
                    if (2 * grid > std::max(c->coop_grid, c->n_sq_blocks)) {
                        cudaFree(c->sq_partial);
                        c->sq_partial = nullptr;
                        if (cudaMalloc(&c->sq_partial, sizeof(double) * 2 * grid) != cudaSuccess) { st = fail(PPO_ERR_CUDA, "cudaMalloc(sq_partial) failed"); break; }
                    }
                }
            }
            // sum-of-squares partials of the cooperative gradient step as LL words (replaces its grid barrier) when one row fits the
            // polling threads and the loss columns P .. P+4 sit in one 64-column chunk
            {
                const int nb = std::max(c->coop_grid, c->persistent_epoch ? 2 * c->epoch_grid : 0);
                if (c->coop && nb <= 256 && (c->d.P & 63) + 5 <= 64 && getenv("PPO_DISABLE_SQ_LL") == nullptr) {
                    if (cudaMalloc(&c->sq_ll, sizeof(uint4) * 2 * (size_t)nb * nb) != cudaSuccess) { st = fail(PPO_ERR_CUDA, "cudaMalloc(sq_ll) failed"); break; }
                    cudaMemset(c->sq_ll, 0, sizeof(uint4) * 2 * (size_t)nb * nb);  // sequence numbers start at 1
                    c->sq_ll_blocks = nb;
                }
            }
            // S family with minibatches of at most 512 samples on one GPU: one single-CTA launch per epoch
            c->small_epoch = c->small && desc->world_size == 1 && nbg / desc->nminibatches <= 512 && getenv("PPO_DISABLE_PERSISTENT") == nullptr;
            c->use_graph = getenv("PPO_DISABLE_GRAPH") == nullptr;  // multi-GPU: only on the fast path (no NCCL inside a graph)
            c->graphs.resize(std::max(1, desc->noptepochs));
            // R family: one CTA per tile of R_TM envs (x tpc tiles) for the whole rollout; needs the parameter vector in
            // shared memory and all CTAs co-resident (grid barrier per env step)
            if (coop_ok && c->d.O == c->d.A && c->d.O <= 32 && getenv("PPO_DISABLE_PERSISTENT") == nullptr) {
                const int ntiles = (desc->n_envs + R_TM - 1) / R_TM;
                for (int tpc = 1; tpc <= 8 && !c->persistent_rollout; ++tpc) {
                    RLayout L;
                    L.init(c->d, tpc);
                    if ((size_t)L.total_bytes > max_smem) break;
                    const int grid = (ntiles + tpc - 1) / tpc;
                    // when one CTA per SM is enough, ask for more than half of the shared memory so that the block scheduler
                    // cannot put two CTAs on one SM (they would run at half speed and everybody waits at the step barrier)
                    size_t smem = (size_t)L.total_bytes;
                    if (grid <= c->sm_count) smem = std::max(smem, std::min(max_smem, (size_t)120 * 1024));
                    if (desc->n_envs == 1) smem = std::max(smem, (size_t)L.total_bytes + rollout_solo_noise_bytes(c->d.O));  // single-env path: noise drawn ahead
                    if (cudaFuncSetAttribute(rollout_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dynamic_smem(rollout_persistent_kernel, max_smem)) != cudaSuccess) {
                        cudaGetLastError();
                        break;
                    }
                    int per = 0;
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, rollout_persistent_kernel, R_NTH, smem);
                    if (per > 0 && grid <= per * c->sm_count && grid <= SV_MAXBLK) {
                        c->persistent_rollout = true;
                        c->roll_grid = grid;
                        c->roll_tpc = tpc;
                        c->roll_smem = smem;
                    }
                }
                if (c->persistent_rollout &&
                    cudaMalloc(&c->roll_partial, sizeof(double) * 2 * (size_t)c->roll_grid * 2 * (c->d.O + 1)) != cudaSuccess) {
                    st = fail(PPO_ERR_CUDA, "cudaMalloc(roll_partial) failed");
                    break;
                }
            }
        }
    } while (0);
    if (st != PPO_OK) {
        char keep[1024];
        strncpy(keep, g_err, sizeof(keep));
        ppo_core_destroy(c);
        strncpy(g_err, keep, sizeof(g_err));
        return st;
    }
    *out = c;
    return PPO_OK;
}

extern "C" int ppo_core_sync(ppo_core* c) {
    if (!c) return fail(PPO_ERR_INVALID, "core is NULL");
    CU(cudaStreamSynchronize(c->stream));
    return PPO_OK;
}
extern "C" void* ppo_core_stream(ppo_core* c) { return c ? (void*)c->stream : nullptr; }
