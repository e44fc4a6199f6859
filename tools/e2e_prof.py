"""GPU box: time of the host-env rollout of the C3 shape (ppo_runner_rollout_replay, 64 env steps x 4096 envs, pinned host
arrays) for the variants of the host-env path.  Usage: python tools/e2e_prof.py"""
import os, subprocess, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))

CHILD = r"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from ppo_cpp_b200 import core
n_envs, n_steps = 4096, 64
c = core.PPOCore(hidden1=64, hidden2=64, n_envs=n_envs, n_steps=n_steps, nminibatches=32, noptepochs=10, seed=1234)
c.init_orthogonal(7)
pin = lambda shape: torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()
rng = np.random.default_rng(0)
raw_obs, raw_rew, raw_done, act = pin((n_steps, n_envs, 18)), pin((n_steps, n_envs)), pin((n_steps, n_envs)), pin((n_steps, n_envs, 18))
raw_obs[:] = rng.standard_normal(raw_obs.shape); raw_rew[:] = rng.standard_normal(raw_rew.shape); raw_done[:] = rng.random(raw_done.shape) < 1 / 334
c.runner_reset(raw_obs[0])
for with_act in (True, False):
    for _ in range(2):
        c.runner_rollout_replay(raw_obs, raw_rew, raw_done, act if with_act else None); c.sync()
    t0 = time.perf_counter()
    for _ in range(5):
        c.runner_rollout_replay(raw_obs, raw_rew, raw_done, act if with_act else None); c.sync()
    dt = (time.perf_counter() - t0) / 5
    print(f"{os.environ.get('VARIANT', 'default'):28s} actions_out={with_act!s:5s}  {1e3 * dt:7.3f} ms per rollout  {1e6 * dt / n_steps:6.1f} us per env step", flush=True)
c.close()
"""
for name, env in (("one kernel, copy engine", {}), ("one kernel, mapped loads", {"PPO_FORCE_HOST_MAPPED": "1"}),
                  ("per-step launches", {"PPO_DISABLE_HOST_PERSISTENT": "1"})):
    e = dict(os.environ, VARIANT=name, **env)
    subprocess.run([sys.executable, "-c", CHILD], env=e, check=False)
