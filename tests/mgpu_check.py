"""Multi-GPU parity check, run under torchrun on N GPUs of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
A run sharded over N ranks (env-sharded rollout, allreduced VecNormalize moments, allgathered buffers, allreduced
gradients) must reproduce the single-GPU run of the same GLOBAL configuration: same Philox streams by global env
id, same global permutation; only fp32 summation order differs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppo_cpp_b200 import core  # noqa: E402
from ppo_cpp_b200.dist import env_world, setup_comm  # noqa: E402


def run(world, rank, device, n_envs_local, hidden, n_steps=32, nmb=4, epochs=2, updates=2, p2p=True):
    c = core.PPOCore(device=device, hidden1=hidden[0], hidden2=hidden[1], n_envs=n_envs_local, n_steps=n_steps, nminibatches=nmb,
                     noptepochs=epochs, seed=99, rank=rank, world_size=world, env_offset=rank * n_envs_local,
                     n_envs_global=n_envs_local * world)
    c.init_orthogonal(5)
    setup_comm(core, c, rank, world)
    if world > 1:
        c.comm_set_p2p(p2p)  # True: peer mailboxes inside the persistent / cooperative kernels; False: NCCL for every exchange
        assert ("persistent" in c.kernel_family("rollout")) == p2p or hidden[0] > 128
    c.shuffle_seed(42)
    c.synth_env_reset()
    losses = None
    for _ in range(updates):
        losses = c.learn_update_synthetic(3e-4, 0.2)
    out = dict(params=c.get_tensor("params"), losses=losses, stats=c.vecnorm_stats(), obs=c.rollout_get("obs"), returns=c.rollout_get("returns"))
    assert not c.comm_error(), "a peer-mailbox wait timed out"
    c.close()
    return out


def main():
    rank, world, local = env_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    cases = (((4, 5), True), ((64, 64), True), ((4, 5), False), ((64, 64), False), ((256, 256), True))
    if os.environ.get("MGPU_CASES") == "wide":
        cases = (((256, 256), True),)
    for hidden, p2p in cases:
        n_local = 48  # 1.5 tiles of the persistent rollout kernel per rank
        sharded = run(world, rank, local, n_local, hidden, p2p=p2p)
        tag = f"hidden {hidden} {'p2p mailbox' if p2p else 'nccl'}"
        if rank == 0:
            single = run(1, 0, local, n_local * world, hidden)
            moved = np.abs(single["params"]).max()
            dp = np.abs(sharded["params"] - single["params"]).max() / moved
            dl = np.abs(sharded["losses"] - single["losses"]).max()
            # rank 0 owns global envs [0, n_local): its local rollout must equal those rows of the single-GPU rollout
            T = 32
            obs1 = single["obs"].reshape(n_local * world, T, 18)[:n_local].reshape(-1, 18)
            dobs = np.abs(sharded["obs"] - obs1).max()
            same_count = sharded["stats"]["obs_count"] == single["stats"]["obs_count"]
            print(f"{tag}: world {world}  param rel diff {dp:.2e}  loss diff {dl:.2e}  obs diff {dobs:.2e}  counts equal {same_count}")
            ok &= dp < 1e-5 and dl < 1e-5 and dobs < 1e-4 and same_count
        # every rank must hold bit-identical parameters after the replicated Adam
        t = torch.tensor(sharded["params"], device="cuda")
        ref = t.clone()
        dist.broadcast(ref, 0)
        same = bool(torch.equal(t, ref))
        flags = [None] * world
        dist.all_gather_object(flags, same)
        if rank == 0:
            print(f"{tag}: replicas bit-identical {all(flags)}")
            ok &= all(flags)
    if rank == 0:
        print("MGPU_CHECK", "OK" if ok else "FAILED")
    dist.destroy_process_group()
    sys.exit(0 if ok or rank != 0 else 1)


if __name__ == "__main__":
    main()
