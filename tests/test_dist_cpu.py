"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard plan, identical permutations on every rank,
every minibatch slot processed exactly once, unique-id broadcast plumbing."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from ppo_cpp_b200 import core
from ppo_cpp_b200.dist import broadcast_bytes, shard_plan


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_envs_local, n_steps, nmb, epochs = 6, 16, 4, 3
        plan = shard_plan(rank, world, n_envs_local, n_steps, nmb)
        uid = broadcast_bytes(bytes(range(128)) if rank == 0 else None)
        perms = core.host_random_shuffle(42, plan.n_batch_global, epochs)  # same seed -> same stream on every rank
        gathered = [None] * world
        dist.all_gather_object(gathered, perms.tobytes())
        mine = []
        for e in range(epochs):
            gather = np.zeros(plan.n_batch_global, np.int64)
            for i, s in enumerate(perms[e]):
                gather[s] = plan.physical_row(i, n_steps)  # out[perm[i]] = in[i]
            for k in range(nmb):
                a, b = plan.minibatch_slots(k)
                mine.append(gather[a:b])
        allrows = [None] * world
        dist.all_gather_object(allrows, np.concatenate(mine).tolist())
        q.put((rank, uid == bytes(range(128)), len(set(gathered)) == 1, plan, allrows))
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, uid_ok, same_perm, plan, allrows in results:
        assert uid_ok and same_perm
        assert plan.env_offset == rank * 6 and plan.n_batch_global == 192 and plan.minibatch_local == 24
        # per epoch, the union over ranks and minibatches covers every physical row exactly once
        per_rank = [np.array(r).reshape(3, -1) for r in allrows]
        for e in range(3):
            rows = np.concatenate([pr[e] for pr in per_rank])
            assert sorted(rows.tolist()) == list(range(192))


def test_shard_plan_validation_and_mapping():
    with pytest.raises(ValueError):
        shard_plan(0, 2, 3, 5, 4)
    plan = shard_plan(1, 4, 8, 16, 8)
    assert plan.minibatch_slots(3) == (3 * 64 + 16, 3 * 64 + 32)
    # semantic row = env_global*T + t ; env 9 belongs to rank 1 (local env 1)
    row = 9 * 16 + 5
    assert plan.owner_of_row(row, 16) == 1
    assert plan.physical_row(row, 16) == 1 * 16 * 8 + 5 * 8 + 1
