#!/usr/bin/env python
"""C5 microbench (BASELINE.json configs[4]): GAE scan, VecNormalize and PPO loss fwd/bwd over a large rollout
buffer, each kernel timed alone with CUDA events on the core's stream (ppo_profile_kernel) and reported as
algorithmic GB/s (SURVEY §8d per-transition bytes) against the measured HBM copy bandwidth.

  python tools/microbench_c5.py [--envs 65536] [--steps 256] [--hidden 4 5] [--sweep]

One JSON line per kernel.  --sweep repeats the measurement for n_envs = 4096 ... --envs (x4 steps) so the
latency-bound and the bandwidth-bound regimes are both visible.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# algorithmic bytes per transition / env step
# the profiled VecNormalize launches carry observations only: moments read 72 B, apply reads 72 B + writes 72 B per env step
# vecnorm_replay: the whole VecNormalize of a recorded [T, n_envs] trajectory, 164 B per transition
BYTES = {"gae": 16, "vecnorm_replay": 164, "norm_moments": 72, "norm_apply": 144, "policy_step": 152, "train_fwdbwd": 160}


def measure(n_envs, n_steps, h1, h2, peak, iters):
    from ppo_cpp_b200 import core
    c = core.PPOCore(hidden1=h1, hidden2=h2, n_envs=n_envs, n_steps=n_steps, nminibatches=32, noptepochs=1, seed=3)
    c.init_orthogonal(7)
    c.shuffle_seed(42)
    c.synth_env_reset()
    c.rollout_synthetic()  # fills the buffer with real rollout data
    c.sync()
    n_batch = n_envs * n_steps
    units = {"gae": n_batch, "vecnorm_replay": n_batch, "norm_moments": n_envs, "norm_apply": n_envs, "policy_step": n_envs, "train_fwdbwd": n_batch // 32}
    out = []
    for k in ("gae", "vecnorm_replay", "norm_moments", "norm_apply", "policy_step", "train_fwdbwd"):
        ms = c.profile_kernel(k, iters)
        gbs = units[k] * BYTES[k] / (ms * 1e-3) / 1e9
        out.append({"kernel": k, "family": c.kernel_family("train") if k == "train_fwdbwd" else (c.kernel_family("policy") if k == "policy_step" else None),
                    "n_envs": n_envs, "n_steps": n_steps, "hidden": [h1, h2], "units_per_launch": units[k],
                    "bytes_per_unit": BYTES[k], "ms": ms, "achieved_gbs": gbs, "peak_gbs": peak, "frac": gbs / peak})
    c.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--hidden", type=int, nargs=2, default=[4, 5])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--sweep", action="store_true")
    args = ap.parse_args()
    peak = 6650.0
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = json.load(f)["hbm_gbs"]
    except Exception:
        pass
    sizes = [args.envs]
    if args.sweep:
        sizes = []
        n = 4096
        while n < args.envs:
            sizes.append(n)
            n *= 4
        sizes.append(args.envs)
    for n in sizes:
        for line in measure(n, args.steps, args.hidden[0], args.hidden[1], peak, args.iters):
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
