// abi_comm.inl — part of libppo_core.so's single translation unit (included by ppo_core.cu, in this order): multi-GPU: communicator set-up, peer mailbox, NCCL fallback.
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
