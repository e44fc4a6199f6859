"""W family vs T family on ppo_loss_grad: per-tensor relative errors (debug aid)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ppo_cpp_b200 import core
import oracle_lib as ol

def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))

def run(H, B, wide):
    if wide: os.environ.pop("PPO_DISABLE_WIDE", None)
    else: os.environ["PPO_DISABLE_WIDE"] = "1"
    rng = np.random.default_rng(5)
    o = ol.Oracle(h1=H, h2=H)
    p = (rng.standard_normal(o.Pq) * 0.1).astype(np.float32)
    p[o.offset(12):o.offset(13)] = (0.1 * rng.standard_normal(18)).astype(np.float32)
    obs = rng.standard_normal((B, 18)).astype(np.float32)
    eps = rng.standard_normal((B, 18)).astype(np.float32)
    adv = rng.standard_normal(B).astype(np.float32)
    c = core.PPOCore(hidden1=H, hidden2=H, n_envs=4, n_steps=8, nminibatches=4)
    c.set_tensor("params", p)
    fam = c.kernel_family("train")
    act, val, nlp = c.policy_step(obs, eps)
    g, l = c.loss_grad(obs, act, adv, val + 0.1, nlp + 0.01, val - 0.05, 0.2)
    c.close()
    return fam, g, l, o

for H, B in ((256, 128), (256, 300), (128, 1000), (256, 5000)):
    fw, gw, lw, o = run(H, B, True)
    ft, gt, lt, _ = run(H, B, False)
    print(f"H {H} B {B}: {fw[:30]} vs {ft[:30]}  grads {rel(gw, gt):.2e}  losses {lw} {lt}")
    names = ["pi_fc0_w", "pi_fc0_b", "vf_fc0_w", "vf_fc0_b", "pi_fc1_w", "pi_fc1_b", "vf_fc1_w", "vf_fc1_b", "vf_w", "vf_b", "pi_w", "pi_b", "logstd"]
    for t in range(13):
        sl = slice(o.offset(t), o.offset(t + 1))
        print(f"   {names[t]:9s} {rel(gw[sl], gt[sl]):.2e}  |max| {np.abs(gt[sl]).max():.3e}")
