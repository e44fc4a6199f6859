"""Multi-GPU plumbing: one process per GPU, torch.distributed only for rendezvous (NCCL unique-id broadcast,
barriers, max-over-ranks timing).  The data path collectives (moment allreduce, rollout allgather, gradient
allreduce) are issued by libppo_core.so on its own NCCL communicator and stream."""
from __future__ import annotations

import os
from dataclasses import dataclass


@dataclass
class ShardPlan:
    rank: int
    world: int
    n_envs_local: int
    n_envs_global: int
    env_offset: int
    n_batch_local: int
    n_batch_global: int
    minibatch_global: int
    minibatch_local: int

    def minibatch_slots(self, k: int):
        """[start, stop) of the slots of minibatch k this rank processes (contiguous, equal on every rank)."""
        s = k * self.minibatch_global + self.rank * self.minibatch_local
        return s, s + self.minibatch_local

    def owner_of_row(self, semantic_row: int, n_steps: int) -> int:
        """rank whose rollout produced flat row = env_global*n_steps + t"""
        return (semantic_row // n_steps) // self.n_envs_local

    def physical_row(self, semantic_row: int, n_steps: int) -> int:
        """position inside the allgathered [rank][t][env_local] buffers (build_gather_kernel's mapping)"""
        env_g, t = divmod(semantic_row, n_steps)
        r, el = divmod(env_g, self.n_envs_local)
        return r * n_steps * self.n_envs_local + t * self.n_envs_local + el


def shard_plan(rank: int, world: int, n_envs_local: int, n_steps: int, nminibatches: int) -> ShardPlan:
    nbl = n_envs_local * n_steps
    nbg = nbl * world
    if nbg % nminibatches:
        raise ValueError(f"n_batch {nbg} not divisible by nminibatches {nminibatches}")
    mbg = nbg // nminibatches
    if mbg % world:
        raise ValueError(f"minibatch {mbg} not divisible by world size {world}")
    return ShardPlan(rank, world, n_envs_local, n_envs_local * world, rank * n_envs_local, nbl, nbg, mbg, mbg // world)


def env_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def broadcast_bytes(payload: bytes | None, src: int = 0) -> bytes:
    """Broadcast a byte string (the NCCL unique id) over the default process group (gloo or nccl)."""
    import torch.distributed as dist
    box = [payload]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def setup_comm(core_module, core, rank: int, world: int):
    """Create the core's NCCL communicator: rank 0 makes the id, everyone joins."""
    if world == 1:
        return
    import torch.distributed as dist
    uid = broadcast_bytes(core_module.comm_unique_id() if rank == 0 else None)
    core.comm_init(uid, rank, world)
    # peer mailboxes: gather every rank's cudaIpc handle and map them (NVLink P2P); on failure the NCCL path stays
    if os.environ.get("PPO_DISABLE_P2P"):
        return
    handles = [None] * world
    dist.all_gather_object(handles, core.comm_ipc_handle())
    ok = True
    try:
        core.comm_ipc_open(handles, world)
    except Exception as ex:  # noqa: BLE001
        ok = False
        if rank == 0:
            print(f"[ppo_cpp_b200] peer mailboxes unavailable, using NCCL for every exchange: {ex}", flush=True)
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if not all(flags):  # all or nothing: every rank must take the same path
        core.comm_set_p2p(False)
