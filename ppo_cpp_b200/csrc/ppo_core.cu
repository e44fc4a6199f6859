// libppo_core.so — C ABI implementation (see include/ppo_core.h).  sm_100a only; no CPU fallback.
#include "../../include/ppo_core.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <chrono>
#include <vector>

#include "device_common.cuh"
#include "host_rand.h"
#include "kernels_misc.cuh"
#include "kernels_mlp.cuh"
#include "kernels_mlp2.cuh"
#include "kernels_rollout.cuh"
#include "kernels_shuffle.cuh"
#include "kernels_small.cuh"
#include "kernels_umma.cuh"
#include "kernels_wide.cuh"
#include "meta_parser.h"

using namespace ppo;

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CU(expr)                                                                                        \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) return fail(PPO_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
#define TRY(expr)                 \
    do {                          \
        int _s = (expr);          \
        if (_s != PPO_OK) return _s; \
    } while (0)

extern "C" const char* ppo_last_error(void) { return g_err; }
extern "C" int ppo_abi_version(void) { return PPO_CORE_ABI_VERSION; }

// ------------------------------------------------------------------------------------------------ NCCL (dlopen)
namespace {
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueIdC { char internal[128]; };
enum { ncclSuccessC = 0 };
enum { ncclFloat32C = 7, ncclFloat64C = 8, ncclInt32C = 2 };
enum { ncclSumC = 0 };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueIdC*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueIdC, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load() {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return false;
#define LD(sym, field) field = reinterpret_cast<decltype(field)>(dlsym(lib, sym))
        LD("ncclGetUniqueId", GetUniqueId);
        LD("ncclCommInitRank", CommInitRank);
        LD("ncclCommDestroy", CommDestroy);
        LD("ncclAllReduce", AllReduce);
        LD("ncclAllGather", AllGather);
        LD("ncclGetErrorString", GetErrorString);
#undef LD
        return GetUniqueId && CommInitRank && CommDestroy && AllReduce && AllGather && GetErrorString;
    }
};
NcclApi g_nccl;
}  // namespace

constexpr int F_TM_TRAIN = 64, F_NT_TRAIN = 512, F_TM_POLICY = 32, F_NT_POLICY = 256;

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

// ------------------------------------------------------------------------------------------------ comm
static int nccl_check(int r, const char* what) {
    if (r == ncclSuccessC) return PPO_OK;
    return fail(PPO_ERR_COMM, "%s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
}
extern "C" int ppo_comm_get_unique_id(char id[PPO_COMM_ID_BYTES]) {
    if (!g_nccl.load()) return fail(PPO_ERR_COMM, "cannot load libnccl.so.2: %s", dlerror());
    ncclUniqueIdC u;
    TRY(nccl_check(g_nccl.GetUniqueId(&u), "ncclGetUniqueId"));
    memcpy(id, u.internal, PPO_COMM_ID_BYTES);
    return PPO_OK;
}
extern "C" int ppo_comm_init(ppo_core* c, const char id[PPO_COMM_ID_BYTES], int rank, int world_size) {
    if (!c || !id) return fail(PPO_ERR_INVALID, "NULL argument");
    if (rank != c->desc.rank || world_size != c->desc.world_size) return fail(PPO_ERR_INVALID, "rank/world_size differ from the core's desc");
    if (!g_nccl.load()) return fail(PPO_ERR_COMM, "cannot load libnccl.so.2: %s", dlerror());
    CU(cudaSetDevice(c->desc.device));
    ncclUniqueIdC u;
    memcpy(u.internal, id, PPO_COMM_ID_BYTES);
    TRY(nccl_check(g_nccl.CommInitRank(&c->comm, world_size, u, rank), "ncclCommInitRank"));
    return PPO_OK;
}
// peer mailboxes: cudaIpc handle of this rank's allocation / mapping of every peer's (one process per GPU, one node)
extern "C" int ppo_comm_ipc_handle(ppo_core* c, char out[PPO_IPC_HANDLE_BYTES]) {
    if (!c || !out) return fail(PPO_ERR_INVALID, "NULL argument");
    if (!c->mbox_mem) return fail(PPO_ERR_INVALID, "world_size is 1: no mailbox");
    static_assert(sizeof(cudaIpcMemHandle_t) == PPO_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    CU(cudaSetDevice(c->desc.device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, c->mbox_mem));
    memcpy(out, &h, sizeof(h));
    return PPO_OK;
}
extern "C" int ppo_comm_ipc_open(ppo_core* c, const char* handles, int world_size) {
    if (!c || !handles) return fail(PPO_ERR_INVALID, "NULL argument");
    if (world_size != c->desc.world_size || !c->mbox_mem) return fail(PPO_ERR_INVALID, "world_size differs from the core's desc");
    CU(cudaSetDevice(c->desc.device));
    for (int r = 0; r < world_size; ++r) {
        if (r == c->desc.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * PPO_IPC_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(PPO_ERR_COMM, "cudaIpcOpenMemHandle(rank %d) failed: %s (the NCCL path stays in use)", r, cudaGetErrorString(e));
        }
        c->mbox_peer[r] = static_cast<unsigned char*>(p);
    }
    c->mbox_ready = getenv("PPO_DISABLE_P2P") == nullptr;
    return PPO_OK;
}
extern "C" int ppo_comm_set_p2p(ppo_core* c, int enable) {
    if (!c) return fail(PPO_ERR_INVALID, "core is NULL");
    if (enable) {
        for (int r = 0; r < c->desc.world_size; ++r)
            if (!c->mbox_peer[r]) return fail(PPO_ERR_INVALID, "mailbox of rank %d is not mapped (ppo_comm_ipc_open)", r);
    }
    c->mbox_ready = enable != 0 && c->desc.world_size > 1;
    return PPO_OK;
}
// 1 when a peer-mailbox wait timed out since the last call (a peer died or was never launched)
extern "C" int ppo_comm_error(ppo_core* c) {
    if (!c) return fail(PPO_ERR_INVALID, "core is NULL");
    if (!c->sync_vars) return 0;
    unsigned e = 0;
    CU(cudaMemcpy(&e, c->sync_vars + SV_ERR, sizeof(e), cudaMemcpyDeviceToHost));
    return e ? 1 : 0;
}
static int need_comm(ppo_core* c) {
    if (c->desc.world_size > 1 && !c->comm) return fail(PPO_ERR_COMM, "world_size %d but ppo_comm_init was not called", c->desc.world_size);
    return PPO_OK;
}

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

// ------------------------------------------------------------------------------------------------ update
extern "C" int ppo_shuffle_seed(ppo_core* c, unsigned seed) {
    if (!c) return fail(PPO_ERR_INVALID, "core is NULL");
    if (c->shuffle_prefetched) {  // permutations drawn from the old stream: discard (the new window is uploaded by the next update)
        cudaSetDevice(c->desc.device);
        cudaStreamSynchronize(c->stream2);
        c->shuffle_prefetched = false;
    }
    c->rng.srand(seed);
    c->rng_on_device = false;  // the host object is authoritative again; the next device shuffle uploads its window
    return PPO_OK;
}
extern "C" int ppo_host_srand_rand(unsigned seed, int count, int* out) {
    if (!out || count < 0) return fail(PPO_ERR_INVALID, "bad arguments");
    GlibcRand r(seed);
    for (int i = 0; i < count; ++i) out[i] = r.rand();
    return PPO_OK;
}
extern "C" int ppo_host_random_shuffle(unsigned seed, int n, int epochs, int* perms_out) {
    if (!perms_out || n < 0 || epochs < 0) return fail(PPO_ERR_INVALID, "bad arguments");
    GlibcRand r(seed);
    std::vector<int> p(n);
    for (int i = 0; i < n; ++i) p[i] = i;
    for (int e = 0; e < epochs; ++e) {
        r.random_shuffle(p.data(), n);
        memcpy(perms_out + (size_t)e * n, p.data(), sizeof(int) * n);
    }
    return PPO_OK;
}

static int allgather_train_inputs(ppo_core* c) {
    if (c->desc.world_size == 1 || c->gathered) return PPO_OK;
    TRY(need_comm(c));
    const int ids[5] = {B_OBS, B_RETURNS, B_ACTIONS, B_VALUES, B_NEGLOGP};
    for (int b : ids) {
        const size_t n = (size_t)c->n_batch_local * c->buf_w[b];
        TRY(nccl_check(g_nccl.AllGather(slab(c, b, 0), c->buf[b], n, ncclFloat32C, c->comm, c->stream), "ncclAllGather(rollout)"));
    }
    return PPO_OK;
}

// upload one epoch's permutation and derive the gather list + per-minibatch advantage statistics
static int prepare_epoch(ppo_core* c, const int* perm_pinned_or_host) {
    const int nb = c->n_batch_global;
    CU(cudaMemcpyAsync(c->perm_dev, perm_pinned_or_host, sizeof(int) * (size_t)nb, cudaMemcpyHostToDevice, c->stream));
    c->ctr.h2d_bytes += sizeof(int) * (size_t)nb;
    LAUNCH(c, build_gather_kernel, (nb + 255) / 256, 256, 0, c->perm_dev, nb, c->desc.n_steps, c->desc.n_envs, c->gather);
    LAUNCH(c, advnorm_stats_kernel, c->desc.nminibatches, 512, 0, c->buf[B_RETURNS], c->buf[B_VALUES], c->gather, c->B_global, c->mbstats, (size_t)0, 0);
    CU(cudaGetLastError());
    c->cur_gather = c->gather;
    c->cur_mbstats = c->mbstats;
    return PPO_OK;
}

// ------------------------------------------------------------------------------------------------ W family (kernels_wide.cuh)
// One allocation holds every operand image and fp32 result of a minibatch of up to `tiles` tiles.  Zero-filled: the
// chunks of the X' and dY images that no kernel writes must read as zeros.
static int ensure_wide(ppo_core* c, int tiles) {
    if (tiles <= c->wide_cap) return PPO_OK;
    if (c->wide_mem) {
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaFree(c->wide_mem));
        c->wide_mem = nullptr;
        c->wide_cap = 0;
        c->wide_images_valid = false;
        // captured graphs hold the old pointers: drop them, they are re-captured on their next use
        for (auto& g : c->graphs)
            if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
        if (c->update_graph.exec) { cudaGraphExecDestroy(c->update_graph.exec); c->update_graph.exec = nullptr; }
        if (c->rollout_graph.exec) { cudaGraphExecDestroy(c->rollout_graph.exec); c->rollout_graph.exec = nullptr; }
    }
    wide::Geom G;
    G.init(c->d.H1, tiles, tiles);
    const size_t R = (size_t)tiles * wide::TM, H = (size_t)G.H;
    const size_t sizes[] = {(size_t)tiles * G.x_tile, 2 * G.act_tower, 2 * G.act_tower, 2 * G.act_tower, 2 * G.act_tower, 2 * G.dy_tower,
                            2 * G.w0_tower, 2 * G.w1_tower, 2 * G.wh_tower, 2 * R * H * sizeof(float), 2 * R * 64 * sizeof(float),
                            4 * (size_t)tiles * wide::COLPART * sizeof(float), 2 * 4 * (size_t)tiles * H * sizeof(float), 2 * R * H * sizeof(float),
                            (size_t)wide::WMAX_BLOCKS * 8 * sizeof(float), 2 * (size_t)wide::SC_STRIDE * sizeof(float),
                            2 * 4 * (size_t)tiles * H * sizeof(float)};
    size_t off[17], total = 0;
    for (int i = 0; i < 17; ++i) {
        off[i] = total;
        total += (sizes[i] + 1023) & ~(size_t)1023;
    }
    CU(cudaMalloc(&c->wide_mem, total));
    CU(cudaMemsetAsync(c->wide_mem, 0, total, c->stream));
    uint8_t* base = static_cast<uint8_t*>(c->wide_mem);
    wide::WideBufs& w = c->wb;
    w.X = base + off[0]; w.H1 = base + off[1]; w.H2 = base + off[2]; w.dP2 = base + off[3]; w.dP1 = base + off[4]; w.dY = base + off[5];
    w.W0 = base + off[6]; w.W1 = base + off[7]; w.WH = base + off[8];
    w.G1 = reinterpret_cast<float*>(base + off[9]); w.MU = reinterpret_cast<float*>(base + off[10]);
    w.G2 = reinterpret_cast<float*>(base + off[13]);
    w.colloss = reinterpret_cast<float*>(base + off[11]); w.colb1 = reinterpret_cast<float*>(base + off[12]);
    w.pmax = reinterpret_cast<float*>(base + off[14]); w.sc = reinterpret_cast<float*>(base + off[15]);
    w.colb0 = reinterpret_cast<float*>(base + off[16]);
    c->wide_cap = tiles;
    return PPO_OK;
}

static void launch_wgemm(ppo_core* c, const wide::GemmArgs& g) {
    const int grid = std::max(1, std::min(g.ntasks, c->sm_count));
    LAUNCH(c, wide::wgemm_kernel, grid, wide::GEMM_NTH, wide::GEMM_SMEM, g);
}

// policy step / value / mean for n envs on the W family: weight images, X' image, three forward GEMMs, per-env tail
static int launch_wide_policy(ppo_core* c, const PolicyArgs& a) {
    using namespace wide;
    const int NT = (a.n + TM - 1) / TM;
    TRY(ensure_wide(c, NT));
    WideBufs w = c->wb;
    Geom& G = w.G;
    G.init(c->d.H1, NT, c->wide_cap);
    const int H = G.H, nb = G.nb;
    const NetDims& d = c->d;
    const int chunks = 2 * nb * 32 * 8 + 2 * nb * nb * 64 * 8 + 2 * nb * 64 * 8;
    if (!c->wide_images_valid) {  // the weight images are those of the current parameters for the whole rollout
        LAUNCH(c, wide_absmax_kernel, WMAX_BLOCKS, 256, 0, a.params, d, w.pmax);
        LAUNCH(c, wide_prep_weights_kernel, (chunks + 255) / 256, 256, 0, a.params, d, w, 1.0f / (float)c->B_global);
        c->wide_images_valid = true;
    }
    LAUNCH(c, wide_policy_gather_kernel, (G.Bpad * (d.O / 8 + 1) + 255) / 256, 256, 0, a.obs, a.n, a.obs_store, d.O, w);
    GemmArgs g{};
    g.P = a.params; g.img_tower = G.act_tower; g.img_tile = G.act_tile; g.img_piece = G.act_piece; g.cap = G.cap; g.H = H;
    g.sc = w.sc; g.sc_fwd = SC_U_W0;
    g.mode = MODE_FWD;
    g.A = w.X; g.a_tower = 0; g.a_tile = G.x_tile; g.a_piece = BLK16; g.kblocks = 1; g.ksteps = 2;
    g.B = w.W0; g.b_tower = G.w0_tower; g.b_piece = G.w0_piece; g.b_kb = 0; g.b_g = 4096; g.b_bytes = 4096;
    g.n_tile = 128; g.n_blks = H / 128; g.m_tiles = NT; g.ntasks = 2 * NT * g.n_blks;
    g.epi = EPI_ACT; g.bias_off[0] = g.bias_off[1] = -1; g.img_out = w.H1; g.gbuf = nullptr;
    launch_wgemm(c, g);
    g.A = w.H1; g.a_tower = G.act_tower; g.a_tile = G.act_tile; g.a_piece = G.act_piece; g.kblocks = nb; g.ksteps = 4;
    g.B = w.W1; g.b_tower = G.w1_tower; g.b_piece = G.w1_piece; g.b_kb = BLK8; g.b_g = (size_t)nb * BLK8; g.b_bytes = BLK8;
    g.bias_off[0] = d.off[T_PI_FC1_B]; g.bias_off[1] = d.off[T_VF_FC1_B]; g.img_out = w.H2; g.sc_fwd = SC_U_W1;
    launch_wgemm(c, g);
    g.A = w.H2;
    g.B = w.WH; g.b_tower = G.wh_tower; g.b_piece = G.wh_piece; g.b_kb = BLK8; g.b_g = 0; g.b_bytes = BLK8;
    g.n_tile = 64; g.n_blks = 1; g.ntasks = 2 * NT; g.sc_fwd = SC_U_HD;
    g.epi = EPI_STORE; g.C = w.MU; g.c_tower = G.mu_tower; g.ldc = 64;
    launch_wgemm(c, g);
    LAUNCH(c, wide_policy_head_kernel, (a.n + 127) / 128, 128, 0, a, w);
    CU(cudaGetLastError());
    return PPO_OK;
}

// loss forward + backward of one minibatch shard -> KG gradient slabs (split-K groups of the weight-gradient GEMMs)
static int launch_wide_train(ppo_core* c, const TrainArgs& a, int* slabs_out) {
    using namespace wide;
    const int NT = (a.count + TM - 1) / TM;
    TRY(ensure_wide(c, NT));
    WideBufs w = c->wb;
    Geom& G = w.G;
    G.init(c->d.H1, NT, c->wide_cap);
    const int H = G.H, nb = G.nb;
    const NetDims& d = c->d;
    {
        const int chunks = 2 * nb * 32 * 8 + 2 * nb * nb * 64 * 8 + 2 * nb * 64 * 8;
        LAUNCH(c, wide_absmax_kernel, WMAX_BLOCKS, 256, 0, a.params, d, w.pmax);
        LAUNCH(c, wide_prep_weights_kernel, (chunks + 255) / 256, 256, 0, a.params, d, w, a.invB);
        c->wide_images_valid = false;  // an Adam step follows
        LAUNCH(c, wide_gather_kernel, (G.Bpad * (d.O / 8 + 1) + 255) / 256, 256, 0, a, w);
    }
    GemmArgs g{};
    g.P = a.params; g.img_tower = G.act_tower; g.img_tile = G.act_tile; g.img_piece = G.act_piece; g.cap = G.cap; g.H = H;
    g.sc = w.sc; g.sc_fwd = SC_U_W0;
    // ---- layer 0: H1 = tanh(X' W0')  (bias through the ones column of X')
    g.mode = MODE_FWD;
    g.A = w.X; g.a_tower = 0; g.a_tile = G.x_tile; g.a_piece = BLK16; g.kblocks = 1; g.ksteps = 2;
    g.B = w.W0; g.b_tower = G.w0_tower; g.b_piece = G.w0_piece; g.b_kb = 0; g.b_g = 4096; g.b_bytes = 4096;
    g.n_tile = 128; g.n_blks = H / 128; g.m_tiles = NT; g.ntasks = 2 * NT * g.n_blks;
    g.epi = EPI_ACT; g.bias_off[0] = g.bias_off[1] = -1; g.img_out = w.H1; g.gbuf = w.G1;
    launch_wgemm(c, g);
    // ---- layer 1: H2 = tanh(H1 W1 + b1)
    g.A = w.H1; g.a_tower = G.act_tower; g.a_tile = G.act_tile; g.a_piece = G.act_piece; g.kblocks = nb; g.ksteps = 4;
    g.B = w.W1; g.b_tower = G.w1_tower; g.b_piece = G.w1_piece; g.b_kb = BLK8; g.b_g = (size_t)nb * BLK8; g.b_bytes = BLK8;
    g.bias_off[0] = d.off[T_PI_FC1_B]; g.bias_off[1] = d.off[T_VF_FC1_B]; g.img_out = w.H2; g.gbuf = w.G2; g.sc_fwd = SC_U_W1;
    launch_wgemm(c, g);
    // ---- heads: [mu | v] = H2 WH, losses and head gradients (dY image) in the epilogue
    g.A = w.H2;
    g.B = w.WH; g.b_tower = G.wh_tower; g.b_piece = G.wh_piece; g.b_kb = BLK8; g.b_g = 0; g.b_bytes = BLK8;
    g.n_tile = 64; g.n_blks = 1; g.ntasks = 2 * NT;
    g.epi = EPI_LOSS; g.ta = a; g.dY = w.dY; g.dy_tower = G.dy_tower; g.dy_tile = G.dy_tile; g.colloss = w.colloss;
    launch_wgemm(c, g);
    // ---- dP2 = (dY WH^T) (1 - H2^2), column sums -> db1
    g.mode = MODE_BWD;
    g.A = w.dY; g.a_tower = G.dy_tower; g.a_tile = G.dy_tile; g.a_piece = BLK16; g.kblocks = 1; g.ksteps = 2;
    g.B = w.WH; g.b_kb = 0; g.b_g = BLK16; g.b_bytes = BLK16;
    g.n_tile = 128; g.n_blks = H / 128; g.ntasks = 2 * NT * g.n_blks;
    g.epi = EPI_DACT; g.gbuf = w.G2; g.img_out = w.dP2; g.colsum = w.colb1;
    launch_wgemm(c, g);
    // ---- dP1 = (dP2 W1^T) (1 - H1^2)
    g.A = w.dP2; g.a_tower = G.act_tower; g.a_tile = G.act_tile; g.a_piece = G.act_piece; g.kblocks = nb; g.ksteps = 4;
    g.B = w.W1; g.b_tower = G.w1_tower; g.b_piece = G.w1_piece; g.b_kb = (size_t)nb * BLK8; g.b_g = BLK16; g.b_bytes = BLK16;
    g.gbuf = w.G1; g.img_out = w.dP1; g.colsum = w.colb0;  // layer-0 bias gradient from the fp32 values
    launch_wgemm(c, g);
    // ---- weight gradients, split over KG groups of samples: dW1 = H1^T dP2, dWhead = H2^T dY, dW0'^T = dP1^T X'
    GemmArgs q{};
    q.mode = MODE_DW;
    q.HT = 2 * NT;
    static const int kg_env = getenv("PPO_WIDE_KG") ? atoi(getenv("PPO_WIDE_KG")) : 0;  // split-K groups (measurements)
    q.KG = std::min(kg_env > 0 ? kg_env : 16, std::min(q.HT, c->max_train_grid));
    q.partial = a.partial; q.PS = a.PS; q.H = H; q.O = d.O; q.A_dim = d.A; q.sc = w.sc;
    q.off_w1[0] = d.off[T_PI_FC1_W]; q.off_w1[1] = d.off[T_VF_FC1_W];
    q.off_w0[0] = d.off[T_PI_FC0_W]; q.off_w0[1] = d.off[T_VF_FC0_W];
    q.off_b0[0] = d.off[T_PI_FC0_B]; q.off_b0[1] = d.off[T_VF_FC0_B];
    q.off_piw = d.off[T_PI_W]; q.off_vfw = d.off[T_VF_W];
    q.n_dw = 3;
    const int mb = H / 128;
    q.dw[0] = DwProb{w.H1, G.act_tower, G.act_tile, G.act_piece, w.dP2, G.act_tower, G.act_tile, G.act_piece, mb, H / 128, 128, DW_W1, 0, 2 * mb * (H / 128) * q.KG};
    q.dw[1] = DwProb{w.H2, G.act_tower, G.act_tile, G.act_piece, w.dY, G.dy_tower, G.dy_tile, BLK16, mb, 1, 64, DW_HEAD, 0, 2 * mb * q.KG};
    q.dw[2] = DwProb{w.dP1, G.act_tower, G.act_tile, G.act_piece, w.X, 0, G.x_tile, BLK16, mb, 1, 64, DW_W0, 0, 2 * mb * q.KG};
    q.dw[1].task0 = q.dw[0].ntasks;
    q.dw[2].task0 = q.dw[1].task0 + q.dw[1].ntasks;
    q.ntasks = q.dw[2].task0 + q.dw[2].ntasks;
    launch_wgemm(c, q);
    LAUNCH(c, wide_fold_kernel, (4 * H + 2 * d.A + 1 + L_PAD + 7) / 8, 256, 0, a, w, q.KG);
    *slabs_out = q.KG;
    return PPO_OK;
}

static int launch_train_kernel(ppo_core* c, TrainArgs& a, bool with_reduce = true, int* grid_out = nullptr) {
    a.d = c->d;
    a.params = c->params;
    a.ent_coef = c->desc.ent_coef / (float)c->desc.world_size;
    a.vf_coef = c->desc.vf_coef;
    a.partial = c->partial;
    a.PS = c->PS;
    int grid;
    a.prof = c->umma_prof;
    if (c->wide) {  // layer-wise tcgen05 GEMMs; the slabs are the split-K groups of the weight-gradient GEMMs
        TRY(launch_wide_train(c, a, &grid));
    } else if (c->small) {  // thread per sample, gradient sums by transposing warp butterflies
        const int nblocks = (a.count + small::NTH - 1) / small::NTH;
        grid = std::max(1, std::min(nblocks, c->max_train_grid));
        LAUNCH(c, (small::train_small_kernel<18, 18, 4, 5>), grid, small::NTH, 0, a);
    } else if (c->umma) {  // tcgen05 path: one CTA per (tile of 128 samples, tower)
        const int ntiles = (a.count + umma::TM - 1) / umma::TM;
        grid = std::max(1, std::min(ntiles, c->sm_count / 2));
        LAUNCH(c, (umma::train_umma_kernel<18, 18, 0>), dim3(grid, 2), umma::NTH, umma::SMEM_BYTES, a, umma::EpochArgs{});
    } else {
        const int tm = c->fused ? F_TM_TRAIN : c->tm;
        const int ntiles = (a.count + tm - 1) / tm;
        grid = std::max(1, std::min(ntiles, c->fused ? c->sm_count : c->max_train_grid));
        if (c->fused) LAUNCH(c, (train_fused_kernel<F_TM_TRAIN, F_NT_TRAIN>), grid, F_NT_TRAIN, c->fused_train_smem, a);
        else if (tm == 64) LAUNCH(c, train_tile_kernel<64>, grid, NT, train_smem_floats<64>(c->d) * sizeof(float), a);
        else LAUNCH(c, train_tile_kernel<32>, grid, NT, train_smem_floats<32>(c->d) * sizeof(float), a);
    }
    if (with_reduce) LAUNCH(c, grad_reduce_kernel, c->n_sq_blocks, 256, 0, c->partial, grid, c->PS, c->d.P, c->grad, c->sq_partial);
    if (grid_out) *grid_out = grid;
    CU(cudaGetLastError());
    return PPO_OK;
}

// one minibatch train step on the device: loss fwd/bwd -> reduce -> (allreduce) -> clip + Adam
static int train_step_device(ppo_core* c, int k, float lr, float cliprange, int loss_row) {
    const int W = c->desc.world_size;
    const int per_rank = c->B_global / W;
    TrainArgs a{};
    a.obs = c->buf[B_OBS]; a.act = c->buf[B_ACTIONS]; a.ret = c->buf[B_RETURNS]; a.val = c->buf[B_VALUES]; a.nlp = c->buf[B_NEGLOGP];
    a.gather = c->cur_gather;
    a.mbstats = c->cur_mbstats + k;
    a.adv_direct = nullptr;
    a.slot0 = k * c->B_global + c->desc.rank * per_rank;
    a.count = per_rank;
    a.invB = 1.0f / (float)c->B_global;
    a.cliprange = cliprange;
    int train_grid = 0;
    const bool coop = c->coop && fast_path(c);
    TRY(launch_train_kernel(c, a, !coop, &train_grid));
    if (coop) {
        ReduceAdamArgs r{};
        r.partial = c->partial; r.G = train_grid; r.PS = c->PS; r.grad = c->grad; r.sq_partial = c->sq_partial;
        r.bar_ctr = c->sync_vars + SV_COOP_FLAGS; r.bar_gen = c->sync_vars + SV_COOP_GEN;
        r.mbox = make_mailbox(c, true); r.mbox_seq = c->sync_vars + SV_GRAD_SEQ;
        r.sq_ll = (c->sq_ll && !c->coop_big) ? c->sq_ll : nullptr; r.sq_seq = c->sync_vars + SV_SQ_SEQ;
        AdamArgs& ad = r.adam;
        ad.params = c->params; ad.m = c->adam_m; ad.v = c->adam_v; ad.grad = c->grad; ad.sq_partial = c->sq_partial;
        ad.nblk = c->coop_grid; ad.P = c->d.P; ad.lr = lr; ad.beta1 = c->desc.adam_beta1; ad.beta2 = c->desc.adam_beta2;
        ad.eps = c->desc.adam_epsilon; ad.clip_norm = c->desc.max_grad_norm;
        ad.bpow_in = c->bpow + c->bpow_slot * 2; ad.bpow_out = c->bpow + (c->bpow_slot ^ 1) * 2;
        ad.invB = a.invB; ad.inv_world = 1.0f / (float)W;
        ad.loss_row = c->loss_rows + (size_t)loss_row * 5; ad.gnorm_out = c->gnorm;
        void* kargs[] = {&r};
        CU(cudaLaunchCooperativeKernel(c->coop_big ? (void*)grad_reduce_adam_big_kernel : (void*)grad_reduce_adam_coop_kernel, dim3(c->coop_grid),
                                       dim3(256), kargs, 0, c->stream));
        c->ctr.kernel_launches++;
        c->bpow_slot ^= 1;
        return PPO_OK;
    }
    if (W > 1) {
        TRY(need_comm(c));
        TRY(nccl_check(g_nccl.AllReduce(c->grad, c->grad, c->PS, ncclFloat32C, ncclSumC, c->comm, c->stream), "ncclAllReduce(grad)"));
        LAUNCH(c, sqnorm_kernel, c->n_sq_blocks, 256, 0, c->grad, c->d.P, c->sq_partial);
    }
    AdamArgs ad{};
    ad.params = c->params; ad.m = c->adam_m; ad.v = c->adam_v; ad.grad = c->grad; ad.sq_partial = c->sq_partial;
    ad.nblk = c->n_sq_blocks; ad.P = c->d.P; ad.lr = lr; ad.beta1 = c->desc.adam_beta1; ad.beta2 = c->desc.adam_beta2;
    ad.eps = c->desc.adam_epsilon; ad.clip_norm = c->desc.max_grad_norm;
    ad.bpow_in = c->bpow + c->bpow_slot * 2; ad.bpow_out = c->bpow + (c->bpow_slot ^ 1) * 2;
    ad.invB = a.invB; ad.inv_world = 1.0f / (float)W;
    ad.loss_row = c->loss_rows + (size_t)loss_row * 5; ad.gnorm_out = c->gnorm;
    LAUNCH(c, adam_kernel, (c->d.P + 255) / 256, 256, 0, ad);
    CU(cudaGetLastError());
    c->bpow_slot ^= 1;
    return PPO_OK;
}

// GPU-shuffle path: every epoch's permutation, gather list and advantage statistics from device kernels
// (kernels_shuffle.cuh), then the epochs back to back.  Nothing here waits for the host.
static int train_step_device(ppo_core* c, int k, float lr, float cliprange, int loss_row);
static int train_epoch_device(ppo_core* c, float lr, float cliprange, int e);
static int train_epoch_small(ppo_core* c, float lr, float cliprange, int e);
// end of the sigma exchange: "my epochs are in your array" to every rank, then wait for every rank's (fenced flag protocol)
__global__ void shuffle_exchange_kernel(PeerMailbox mbox, unsigned* seq_var) {
    if (threadIdx.x == 0) {
        const unsigned seq = *seq_var + 1u;
        mbox.signal_all(PPO_MBOX_SHUF_CHANNEL, seq);
        mbox.wait_all(PPO_MBOX_SHUF_CHANNEL, seq);
        *seq_var = seq;
    }
}

static int enqueue_shuffle(ppo_core* c) {
    const int n = c->n_batch_global, E = c->desc.noptepochs;
    const long long total = (long long)E * (n - 1);
    // multi-GPU with mapped peer memory: rank r builds sigma of epochs r, r + W, ... (the swap chains of different epochs are
    // independent, only the composition is sequential) and stores them into every rank's array over NVLink; every rank used
    // to build all E x n_global of it (2.8 ms at 8 x 262 144 transitions, the longest thing beside the rollout)
    const bool sharded = c->desc.world_size > 1 && c->mbox_ready && c->arena_sigma_off != 0 && getenv("PPO_DISABLE_SHUFFLE_SHARDING") == nullptr;
    const int e0 = sharded ? c->desc.rank : 0, es = sharded ? c->desc.world_size : 1;
    const int Emy = e0 < E ? (E - e0 + es - 1) / es : 0;
    if (Emy > 0) {
        for (int y = 0; y < Emy; ++y)
            CU(cudaMemsetAsync(c->sh_cnt + (size_t)(e0 + y * es) * (n + 1), 0, sizeof(int) * (size_t)(n + 1), c->stream));
        const int draw_blocks = (int)(((long long)(n - 1) + shuf::L - 1) / shuf::L) + 1;  // L-blocks of the stream overlapping one epoch
        LAUNCH(c, shuf::shuffle_draw_kernel, dim3((draw_blocks + 127) / 128, Emy), 128, 0, c->rng_win, c->shuf_tab, n, E, c->sh_j, e0, es);
    }
    LAUNCH(c, shuf::shuffle_advance_kernel, 1, 32, 0, c->rng_win, c->shuf_tab, (unsigned long long)total);
    if (Emy > 0) {
        const dim3 gn((n + 255) / 256, Emy);
        LAUNCH(c, shuf::shuffle_count_kernel, gn, 256, 0, c->sh_j, n, c->sh_cnt, e0, es);
        const int nb = (n + 1 + shuf::SCAN_TILE - 1) / shuf::SCAN_TILE;
        LAUNCH(c, shuf::shuffle_scan_totals_kernel, dim3(nb, Emy), shuf::SCAN_TILE, 0, c->sh_cnt, n, nb, c->sh_btot, e0, es);
        LAUNCH(c, shuf::shuffle_scan_blocks_kernel, Emy, shuf::SCAN_TILE, 0, nb, c->sh_btot, e0, es);
        LAUNCH(c, shuf::shuffle_scan_final_kernel, dim3(nb, Emy), shuf::SCAN_TILE, 0, c->sh_cnt, n, nb, c->sh_btot, c->sh_off, c->sh_cur, e0, es);
        LAUNCH(c, shuf::shuffle_scatter_kernel, gn, 256, 0, c->sh_j, n, c->sh_cur, c->sh_list, e0, es);
        LAUNCH(c, shuf::shuffle_resolve_kernel, gn, 256, 0, c->sh_j, c->sh_off, c->sh_list, n, c->sh_sigma, e0, es);
        // (resolving the epochs one after the other on an L2-resident working set was measured at n = 2 M: no gain)
    }
    if (sharded) {
        if (Emy > 0) {
            shuf::SigmaPeers sp{};
            sp.rank = c->desc.rank; sp.world = c->desc.world_size;
            for (int r = 0; r < sp.world; ++r) sp.p[r] = reinterpret_cast<int*>(c->mbox_peer[r] + c->arena_sigma_off);
            LAUNCH(c, shuf::shuffle_publish_kernel, dim3(std::min((n + 255) / 256, 4 * c->sm_count), Emy), 256, 0, sp, c->sh_sigma, n, e0, es);
        }
        LAUNCH(c, shuffle_exchange_kernel, 1, 32, 0, make_mailbox(c, false), c->sync_vars + SV_SHUF_SEQ);
    }
    for (int e = 0; e < E; ++e)
        LAUNCH(c, shuf::shuffle_compose_kernel, (n + 255) / 256, 256, 0, e ? c->sh_perm + (size_t)(e - 1) * n : (const int*)nullptr,
               c->sh_sigma + (size_t)e * n, n, c->desc.n_steps, c->desc.n_envs, c->sh_perm + (size_t)e * n, c->sh_gather + (size_t)e * n);
    CU(cudaGetLastError());
    return PPO_OK;
}
static int enqueue_epochs(ppo_core* c, float lr, float cliprange) {
    const int n = c->n_batch_global, E = c->desc.noptepochs, M = c->desc.nminibatches;
    LAUNCH(c, advnorm_stats_kernel, dim3(M, E), 512, 0, c->buf[B_RETURNS], c->buf[B_VALUES], c->sh_gather, c->B_global, c->sh_mbstats, (size_t)n, M);
    CU(cudaGetLastError());
    for (int e = 0; e < E; ++e) {
        c->cur_gather = c->sh_gather + (size_t)e * n;
        c->cur_mbstats = c->sh_mbstats + (size_t)e * M;
        if (c->persistent_epoch && fast_path(c)) TRY(train_epoch_device(c, lr, cliprange, e));
        else if (c->small_epoch) TRY(train_epoch_small(c, lr, cliprange, e));
        else
            for (int k = 0; k < M; ++k) TRY(train_step_device(c, k, lr, cliprange, e * M + k));
    }
    c->perm_set = true;  // cur_gather / cur_mbstats describe the last epoch
    return PPO_OK;
}

// capture `enqueue` (launches on c->stream) once and replay it on `on`; lr / cliprange / beta-power slot are baked in
template <class F>
static int replay_graph(ppo_core* c, ppo_core::EpochGraph& g, cudaStream_t on, float lr, float cliprange, F enqueue, bool uses_adam = true) {
    if (!g.exec || g.lr != lr || g.cliprange != cliprange || (uses_adam && g.bpow_slot != c->bpow_slot)) {
        if (g.exec) {
            cudaGraphExecDestroy(g.exec);
            g.exec = nullptr;
        }
        const int slot0 = c->bpow_slot;
        const uint64_t k0 = c->ctr.kernel_launches;
        cudaGraph_t graph = nullptr;
        CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        const int st = enqueue();
        const cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
        g.kernels = c->ctr.kernel_launches - k0;
        c->ctr.kernel_launches = k0;  // nothing ran yet; the replay accounts for them
        g.flip = c->bpow_slot ^ slot0;
        c->bpow_slot = slot0;
        if (st != PPO_OK) {
            if (graph) cudaGraphDestroy(graph);
            return st;
        }
        if (ce != cudaSuccess) return fail(PPO_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
        const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) return fail(PPO_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
        g.lr = lr; g.cliprange = cliprange; g.bpow_slot = slot0;
    }
    CU(cudaGraphLaunch(g.exec, on));
    c->ctr.graph_launches++;
    c->ctr.kernel_launches += g.kernels;
    c->bpow_slot ^= g.flip;
    return PPO_OK;
}

// The permutations of an update depend only on the rand() stream, not on the rollout: build the next update's on a
// second stream while the rollout runs (called when a rollout starts).  Undone by drop_shuffle_prefetch.
static int prefetch_shuffle(ppo_core* c) {
    if (getenv("PPO_DISABLE_SHUFFLE_PREFETCH") != nullptr || c->shuffle_prefetched || !c->rng_on_device || !(c->gpu_shuffle && fast_path(c)) || !update_graph_ok(c) || c->desc.noptepochs < 1)
        return PPO_OK;
    if (!c->stream2) {
        CU(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_shuf, cudaEventDisableTiming));
        CU(cudaMalloc(&c->rng_win_saved, 31 * sizeof(uint32_t)));
    }
    CU(cudaEventRecord(c->ev_main, c->stream));        // the previous update (it reads sh_gather) has been enqueued before this point
    CU(cudaStreamWaitEvent(c->stream2, c->ev_main, 0));
    CU(cudaMemcpyAsync(c->rng_win_saved, c->rng_win, 31 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream2));
    TRY(replay_graph(c, c->shuffle_graph, c->stream2, 0.f, 0.f, [&]() { return enqueue_shuffle(c); }, false));
    CU(cudaEventRecord(c->ev_shuf, c->stream2));
    c->shuffle_prefetched = true;
    return PPO_OK;
}
// the prefetched permutations will not be used (re-seed, switch to the host shuffle): put the generator back
static int drop_shuffle_prefetch(ppo_core* c, bool restore_window) {
    if (!c->shuffle_prefetched) return PPO_OK;
    CU(cudaStreamSynchronize(c->stream2));
    if (restore_window) {
        CU(cudaMemcpyAsync(c->rng_win, c->rng_win_saved, 31 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    c->shuffle_prefetched = false;
    return PPO_OK;
}

// all minibatches of epoch e in one cooperative launch (U family, persistent): see kernels_umma.cuh
// the same for the S family when one CTA handles a minibatch: see kernels_small.cuh
static int train_epoch_small(ppo_core* c, float lr, float cliprange, int e) {
    const int M = c->desc.nminibatches;
    TrainArgs a{};
    a.obs = c->buf[B_OBS]; a.act = c->buf[B_ACTIONS]; a.ret = c->buf[B_RETURNS]; a.val = c->buf[B_VALUES]; a.nlp = c->buf[B_NEGLOGP];
    a.gather = c->cur_gather; a.mbstats = c->cur_mbstats; a.adv_direct = nullptr; a.slot0 = 0; a.count = c->B_global;
    a.invB = 1.0f / (float)c->B_global; a.cliprange = cliprange;
    a.d = c->d; a.params = c->params; a.ent_coef = c->desc.ent_coef; a.vf_coef = c->desc.vf_coef;
    a.partial = c->partial; a.PS = c->PS; a.prof = nullptr;
    small::SmallEpochArgs ep{};
    ep.M = M; ep.B = c->B_global; ep.mbstats = c->cur_mbstats; ep.loss_rows = c->loss_rows + (size_t)e * M * 5;
    AdamArgs& ad = ep.adam;
    ad.params = c->params; ad.m = c->adam_m; ad.v = c->adam_v; ad.grad = c->grad; ad.sq_partial = c->sq_partial;
    ad.nblk = 1; ad.P = c->d.P; ad.lr = lr; ad.beta1 = c->desc.adam_beta1; ad.beta2 = c->desc.adam_beta2;
    ad.eps = c->desc.adam_epsilon; ad.clip_norm = c->desc.max_grad_norm;
    ad.bpow_in = c->bpow + c->bpow_slot * 2; ad.bpow_out = c->bpow + (c->bpow_slot ^ 1) * 2;
    ad.invB = a.invB; ad.inv_world = 1.0f;
    ad.loss_row = nullptr; ad.gnorm_out = c->gnorm;
    LAUNCH(c, (small::train_small_epoch_kernel<18, 18, 4, 5>), 1, small::NTH, 0, a, ep);
    CU(cudaGetLastError());
    c->bpow_slot ^= 1;
    return PPO_OK;
}

static int train_epoch_device(ppo_core* c, float lr, float cliprange, int e) {
    const int W = c->desc.world_size, M = c->desc.nminibatches;
    const int per_rank = c->B_global / W;
    TrainArgs a{};
    a.obs = c->buf[B_OBS]; a.act = c->buf[B_ACTIONS]; a.ret = c->buf[B_RETURNS]; a.val = c->buf[B_VALUES]; a.nlp = c->buf[B_NEGLOGP];
    a.gather = c->cur_gather; a.mbstats = c->cur_mbstats; a.adv_direct = nullptr; a.slot0 = 0; a.count = per_rank;
    a.invB = 1.0f / (float)c->B_global; a.cliprange = cliprange;
    a.d = c->d; a.params = c->params; a.ent_coef = c->desc.ent_coef / (float)W; a.vf_coef = c->desc.vf_coef;
    a.partial = c->partial; a.PS = c->PS; a.prof = c->umma_prof;
    umma::EpochArgs ep{};
    ep.M = M; ep.B = c->B_global; ep.rank_off = c->desc.rank * per_rank; ep.mbstats = c->cur_mbstats;
    ep.loss_rows = c->loss_rows + (size_t)e * M * 5;
    ReduceAdamArgs& r = ep.ra;
    r.partial = c->partial; r.G = c->epoch_grid; r.PS = c->PS; r.grad = c->grad; r.sq_partial = c->sq_partial;
    r.bar_ctr = c->sync_vars + SV_EPOCH_FLAGS; r.bar_gen = c->sync_vars + SV_EPOCH_GEN;
    r.mbox = make_mailbox(c, true); r.mbox_seq = c->sync_vars + SV_GRAD_SEQ;
    r.sq_ll = c->sq_ll; r.sq_seq = c->sync_vars + SV_SQ_SEQ;
    AdamArgs& ad = r.adam;
    ad.params = c->params; ad.m = c->adam_m; ad.v = c->adam_v; ad.grad = c->grad; ad.sq_partial = c->sq_partial;
    ad.nblk = 2 * c->epoch_grid; ad.P = c->d.P; ad.lr = lr; ad.beta1 = c->desc.adam_beta1; ad.beta2 = c->desc.adam_beta2;
    ad.eps = c->desc.adam_epsilon; ad.clip_norm = c->desc.max_grad_norm;
    ad.bpow_in = c->bpow + c->bpow_slot * 2; ad.bpow_out = c->bpow + (c->bpow_slot ^ 1) * 2;
    ad.invB = a.invB; ad.inv_world = 1.0f / (float)W;
    ad.loss_row = nullptr; ad.gnorm_out = c->gnorm;
    void* kargs[] = {&a, &ep};
    CU(cudaLaunchCooperativeKernel((void*)umma::train_umma_kernel<18, 18, 1>, dim3(c->epoch_grid, 2), dim3(umma::NTH), kargs,
                                   umma::SMEM_BYTES, c->stream));
    c->ctr.kernel_launches++;
    c->bpow_slot ^= 1;
    return PPO_OK;
}

// mean losses of the update to the host; on a multi-GPU run the peer-mailbox error flag travels with them: a wait that
// timed out (a peer died or never arrived) fails the call instead of returning numbers computed from stale slots
static int read_losses_checked(ppo_core* c, float* mean_losses) {
    TRY(d2h(c, mean_losses, c->loss_mean, 5));
    unsigned err = 0;
    const bool check = c->desc.world_size > 1 && c->mbox_ready && c->sync_vars;
    if (check) CU(cudaMemcpyAsync(&err, c->sync_vars + SV_ERR, sizeof(err), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (check && err) return fail(PPO_ERR_COMM, "a peer-mailbox wait timed out during the update (rank %d): a peer is gone or never arrived", c->desc.rank);
    return PPO_OK;
}

extern "C" int ppo_train_update(ppo_core* c, float lr, float cliprange, float* mean_losses) {
    if (!c) return fail(PPO_ERR_INVALID, "core is NULL");
    CU(cudaSetDevice(c->desc.device));
    const int nb = c->n_batch_global, E = c->desc.noptepochs, M = c->desc.nminibatches;
    static const bool timing = getenv("PPO_TIMING") != nullptr;
    cudaEvent_t tg0 = nullptr, tg1 = nullptr;
    if (timing) {
        cudaEventCreate(&tg0); cudaEventCreate(&tg1);
        cudaEventRecord(tg0, c->stream);
    }
    TRY(allgather_train_inputs(c));
    if (timing) cudaEventRecord(tg1, c->stream);
    // host-shuffle path: the previous update's H2D copies out of the pinned permutation buffers must have finished
    if (!(c->gpu_shuffle && fast_path(c)) || timing) CU(cudaStreamSynchronize(c->stream));
    if (timing) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, tg0, tg1);
        fprintf(stderr, "[ppo timing] rank %d allgather of the rollout buffers: %.3f ms\n", c->desc.rank, ms);
        cudaEventDestroy(tg0); cudaEventDestroy(tg1);
    }
    if (c->gpu_shuffle && fast_path(c) && E > 0) {
        // the generator state moves to the device (once; ppo_shuffle_seed moves it back to the host object)
        if (!c->rng_on_device) {
            c->rng.get_window(c->win_pinned);
            CU(cudaMemcpyAsync(c->rng_win, c->win_pinned, 31 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            c->rng_on_device = true;
        }
        if (!update_graph_ok(c)) {
            TRY(enqueue_shuffle(c));
            TRY(enqueue_epochs(c, lr, cliprange));
        } else {
            if (c->shuffle_prefetched) {  // built on stream2 while the rollout ran
                CU(cudaStreamWaitEvent(c->stream, c->ev_shuf, 0));
                c->shuffle_prefetched = false;
            } else {
                TRY(replay_graph(c, c->shuffle_graph, c->stream, 0.f, 0.f, [&]() { return enqueue_shuffle(c); }, false));
            }
            TRY(replay_graph(c, c->update_graph, c->stream, lr, cliprange, [&]() { return enqueue_epochs(c, lr, cliprange); }));
        }
        if (E * M > 0) LAUNCH(c, loss_mean_kernel, 1, 32, 0, c->loss_rows, E * M, c->loss_mean);
        CU(cudaGetLastError());
        if (mean_losses) TRY(read_losses_checked(c, mean_losses));
        return PPO_OK;
    }
    TRY(drop_shuffle_prefetch(c, true));
    if (c->rng_on_device) {  // host path after a device shuffle: bring the generator state back
        CU(cudaMemcpyAsync(c->win_pinned, c->rng_win, 31 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->rng.set_window(c->win_pinned);
        c->rng_on_device = false;
    }
    for (int i = 0; i < nb; ++i) c->perm_host[i] = i;  // perm.setIdentity() once per update (ppo2.hpp:274-275)
    for (int e = 0; e < E; ++e) {
        c->rng.random_shuffle(c->perm_host.data(), nb);  // compounded across epochs (ppo2.hpp:288)
        int* pinned = c->perm_pinned + (size_t)e * nb;
        memcpy(pinned, c->perm_host.data(), sizeof(int) * (size_t)nb);
        if (!update_graph_ok(c)) {
            TRY(prepare_epoch(c, pinned));
            if (c->persistent_epoch && fast_path(c)) TRY(train_epoch_device(c, lr, cliprange, e));
            else if (c->small_epoch) TRY(train_epoch_small(c, lr, cliprange, e));
            else
                for (int k = 0; k < M; ++k) TRY(train_step_device(c, k, lr, cliprange, e * M + k));
            continue;
        }
        ppo_core::EpochGraph& eg = c->graphs[e];
        if (!eg.exec || eg.lr != lr || eg.cliprange != cliprange || eg.bpow_slot != c->bpow_slot) {
            if (eg.exec) {
                cudaGraphExecDestroy(eg.exec);
                eg.exec = nullptr;
            }
            const int slot0 = c->bpow_slot;
            const uint64_t k0 = c->ctr.kernel_launches, h0 = c->ctr.h2d_bytes;
            cudaGraph_t graph = nullptr;
            CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            int st = prepare_epoch(c, pinned);
            if (st == PPO_OK && c->persistent_epoch && fast_path(c)) st = train_epoch_device(c, lr, cliprange, e);
            else if (st == PPO_OK && c->small_epoch) st = train_epoch_small(c, lr, cliprange, e);
            else
                for (int k = 0; k < M && st == PPO_OK; ++k) st = train_step_device(c, k, lr, cliprange, e * M + k);
            const cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
            eg.kernels = c->ctr.kernel_launches - k0;
            c->ctr.kernel_launches = k0;  // nothing ran yet; replay accounts for them
            c->ctr.h2d_bytes = h0;
            eg.flip = c->bpow_slot ^ slot0;
            c->bpow_slot = slot0;
            if (st != PPO_OK) {
                if (graph) cudaGraphDestroy(graph);
                return st;
            }
            if (ce != cudaSuccess) return fail(PPO_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
            const cudaError_t ie = cudaGraphInstantiate(&eg.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ie != cudaSuccess) return fail(PPO_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
            eg.lr = lr; eg.cliprange = cliprange; eg.bpow_slot = slot0;
        }
        CU(cudaGraphLaunch(eg.exec, c->stream));
        c->ctr.graph_launches++;
        c->ctr.kernel_launches += eg.kernels;
        c->ctr.h2d_bytes += sizeof(int) * (size_t)nb;
        c->bpow_slot ^= eg.flip;
    }
    if (E * M > 0) LAUNCH(c, loss_mean_kernel, 1, 32, 0, c->loss_rows, E * M, c->loss_mean);
    CU(cudaGetLastError());
    if (mean_losses) TRY(read_losses_checked(c, mean_losses));
    return PPO_OK;
}

extern "C" int ppo_train_get_permutation(ppo_core* c, int epoch, int* out, int n) {
    if (!c || !out) return fail(PPO_ERR_INVALID, "NULL argument");
    if (n != c->n_batch_global) return fail(PPO_ERR_INVALID, "permutation has %d entries, n_batch is %d", c->n_batch_global, n);
    if (epoch < 0 || epoch >= c->desc.noptepochs) return fail(PPO_ERR_INVALID, "epoch %d out of range", epoch);
    CU(cudaSetDevice(c->desc.device));
    if (c->gpu_shuffle && fast_path(c)) {
        if (c->shuffle_prefetched)
            return fail(PPO_ERR_INVALID, "the permutations of the last update are gone: the next rollout has started (they are rebuilt then)");
        CU(cudaMemcpyAsync(out, c->sh_perm + (size_t)epoch * n, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    } else {
        CU(cudaStreamSynchronize(c->stream));
        memcpy(out, c->perm_pinned + (size_t)epoch * n, sizeof(int) * (size_t)n);
    }
    return PPO_OK;
}

extern "C" int ppo_train_set_permutation(ppo_core* c, const int* perm, int n) {
    if (!c || !perm) return fail(PPO_ERR_INVALID, "NULL argument");
    if (n != c->n_batch_global) return fail(PPO_ERR_INVALID, "permutation has %d entries, n_batch is %d", n, c->n_batch_global);
    CU(cudaSetDevice(c->desc.device));
    std::vector<char> seen(n, 0);
    for (int i = 0; i < n; ++i) {
        if (perm[i] < 0 || perm[i] >= n || seen[perm[i]]) return fail(PPO_ERR_INVALID, "not a permutation (entry %d = %d)", i, perm[i]);
        seen[perm[i]] = 1;
    }
    TRY(allgather_train_inputs(c));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(c->perm_pinned, perm, sizeof(int) * (size_t)n);
    TRY(prepare_epoch(c, c->perm_pinned));
    CU(cudaStreamSynchronize(c->stream));
    c->perm_set = true;
    return PPO_OK;
}

extern "C" int ppo_train_minibatch(ppo_core* c, int k, float lr, float cliprange, float* losses, float* grads) {
    if (!c) return fail(PPO_ERR_INVALID, "core is NULL");
    if (!c->perm_set) return fail(PPO_ERR_INVALID, "call ppo_train_set_permutation first");
    if (k < 0 || k >= c->desc.nminibatches) return fail(PPO_ERR_INVALID, "minibatch %d out of range", k);
    CU(cudaSetDevice(c->desc.device));
    const int row = c->desc.noptepochs * c->desc.nminibatches;  // spare row
    TRY(train_step_device(c, k, lr, cliprange, row));
    if (losses) TRY(d2h(c, losses, c->loss_rows + (size_t)row * 5, 5));
    if (grads) TRY(d2h(c, grads, c->grad, c->d.P));
    CU(cudaStreamSynchronize(c->stream));
    return PPO_OK;
}

extern "C" int ppo_advnorm(ppo_core* c, const float* returns, const float* values, int n, float* advs) {
    if (!c || !returns || !values || !advs || n < 2) return fail(PPO_ERR_INVALID, "ppo_advnorm: bad arguments (the reference asserts rows > 1)");
    CU(cudaSetDevice(c->desc.device));
    TRY(ensure_scratch(c, 3 * (size_t)n + 4));
    float* d_ret = c->scratch; float* d_val = d_ret + n; float* d_out = d_val + n;
    float2* d_st = reinterpret_cast<float2*>((reinterpret_cast<uintptr_t>(d_out + n) + 7) & ~(uintptr_t)7);
    TRY(h2d(c, d_ret, returns, n)); TRY(h2d(c, d_val, values, n));
    LAUNCH(c, advnorm_stats_kernel, 1, 512, 0, d_ret, d_val, (const int*)nullptr, n, d_st);
    LAUNCH(c, advnorm_apply_kernel, (n + 255) / 256, 256, 0, d_ret, d_val, n, d_st, d_out);
    CU(cudaGetLastError());
    TRY(d2h(c, advs, d_out, n));
    CU(cudaStreamSynchronize(c->stream));
    return PPO_OK;
}

extern "C" int ppo_loss_grad(ppo_core* c, const float* obs, const float* actions, const float* advs, const float* returns,
                             const float* old_neglogp, const float* old_values, int B, float cliprange, float* grads, float* losses) {
    if (!c || !obs || !actions || !advs || !returns || !old_neglogp || !old_values || B < 1) return fail(PPO_ERR_INVALID, "ppo_loss_grad: bad arguments");
    CU(cudaSetDevice(c->desc.device));
    const int O = c->d.O, A = c->d.A;
    TRY(ensure_scratch(c, (size_t)B * (O + A + 4)));
    float* d_obs = c->scratch; float* d_act = d_obs + (size_t)B * O; float* d_adv = d_act + (size_t)B * A;
    float* d_ret = d_adv + B; float* d_nlp = d_ret + B; float* d_val = d_nlp + B;
    TRY(h2d(c, d_obs, obs, (size_t)B * O)); TRY(h2d(c, d_act, actions, (size_t)B * A)); TRY(h2d(c, d_adv, advs, B));
    TRY(h2d(c, d_ret, returns, B)); TRY(h2d(c, d_nlp, old_neglogp, B)); TRY(h2d(c, d_val, old_values, B));
    TrainArgs a{};
    a.obs = d_obs; a.act = d_act; a.ret = d_ret; a.val = d_val; a.nlp = d_nlp; a.gather = nullptr; a.mbstats = nullptr;
    a.adv_direct = d_adv; a.slot0 = 0; a.count = B; a.invB = 1.0f / (float)B; a.cliprange = cliprange;
    TRY(launch_train_kernel(c, a));
    std::vector<float> g(c->PS);
    TRY(d2h(c, g.data(), c->grad, c->PS));
    CU(cudaStreamSynchronize(c->stream));
    if (grads) memcpy(grads, g.data(), sizeof(float) * c->d.P);
    if (losses) {
        const float* L = g.data() + c->d.P;
        losses[0] = L[L_PG] * a.invB; losses[1] = 0.5f * (L[L_VF] * a.invB); losses[2] = L[L_ENT];
        losses[3] = 0.5f * (L[L_KL] * a.invB); losses[4] = L[L_CLIP] * a.invB;
    }
    return PPO_OK;
}

extern "C" int ppo_learn_update_synthetic(ppo_core* c, float lr, float cliprange, float* mean_losses) {
    TRY(ppo_rollout_synthetic(c));
    return ppo_train_update(c, lr, cliprange, mean_losses);
}

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
