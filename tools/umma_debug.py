"""Debug helper (GPU box): per-tensor gradient error of the kernel families vs the fp64 oracle + U-family phase timing."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import oracle_lib as ol
from ppo_cpp_b200 import core

def rel(a, b): return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
h1 = h2 = 64; B = 8192
rng = np.random.default_rng(B + h1)
o = ol.Oracle(h1=h1, h2=h2)
p = (rng.standard_normal(o.Pq) * min(0.3, 1.5 / np.sqrt(64))).astype(np.float32)
p[o.offset(12):o.offset(13)] = (-0.5 + 0.2 * rng.standard_normal(18)).astype(np.float32)
obs = rng.standard_normal((B, 18)).astype(np.float32)
_, v64, _, m64 = o.policy_step(p, obs, None, "f64")
std = np.exp(p[o.offset(12):o.offset(13)].astype(np.float64))
act = (m64 + std * rng.standard_normal((B, 18))).astype(np.float32)
z = (act - m64) / std
old_nlp = (0.5 * (z * z).sum(1) + 18 * 0.9189385175704956 + np.log(std).sum() + 0.1 * rng.standard_normal(B)).astype(np.float32)
old_v = (v64 + 0.3 * rng.standard_normal(B)).astype(np.float32)
ret = (v64 + 0.5 * rng.standard_normal(B)).astype(np.float32)
adv = rng.standard_normal(B).astype(np.float32)
g64, l64 = o.loss_grad(p, obs, act, adv, ret, old_nlp, old_v, 0.2, "f64")
g32, l32 = o.loss_grad(p, obs, act, adv, ret, old_nlp, old_v, 0.2, "f32")
print("oracle f32 vs f64 per tensor:", " ".join(f"{rel(g32[o.offset(t):o.offset(t+1)], g64[o.offset(t):o.offset(t+1)]):.1e}" for t in range(13)))
for env in (None, "PPO_DISABLE_UMMA", "PPO_DISABLE_FUSED"):
    if env: os.environ[env] = "1"
    c = core.PPOCore(hidden1=h1, hidden2=h2, n_envs=4, n_steps=8, nminibatches=4)
    c.set_tensor("params", p)
    g, l = c.loss_grad(obs, act, adv, ret, old_nlp, old_v, 0.2)
    print(env or "UMMA", "per tensor:", " ".join(f"{rel(g[o.offset(t):o.offset(t+1)], g64[o.offset(t):o.offset(t+1)]):.1e}" for t in range(13)), "losses", l - l64)
    c.close()
    if env: del os.environ[env]
os.environ["PPO_UMMA_PROF"] = "1"
c = core.PPOCore(hidden1=64, hidden2=64, n_envs=4096, n_steps=64, nminibatches=32, noptepochs=10)
c.init_orthogonal(1)
c.synth_env_reset(); c.rollout_synthetic()
print("train_fwdbwd ms:", c.profile_kernel("train_fwdbwd", 50))
c.close()
