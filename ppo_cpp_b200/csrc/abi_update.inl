// abi_update.inl — part of libppo_core.so's single translation unit (included by ppo_core.cu, in this order): update: permutations, minibatch launches of every kernel family, epochs as CUDA graphs, ppo_train_update.
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
