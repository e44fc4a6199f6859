import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ppo_cpp_b200 import core
from ppo_cpp_b200.meta_graph import TENSOR_ORDER, param_layout
from conftest import load_npz_tree, rel_err
lk = load_npz_tree("loss_grad_kat.npz")
for case in ["init_4_5", "rand_8_8", "rand_64_64"]:
    k = lk[case]
    h1, h2 = (int(x) for x in k["hidden"])
    c = core.PPOCore(hidden1=h1, hidden2=h2, n_envs=4, n_steps=8, nminibatches=4, ent_coef=0.0007160293171182275, vf_coef=0.5)
    c.set_tensor("params", k["params"])
    g, l = c.loss_grad(k["obs"], k["act"], k["adv"], k["ret"], k["old_nlp"], k["old_v"], float(k["cliprange"]))
    print(case, "losses", l, k["losses"])
    lay = param_layout(18, 18, h1, h2)
    for name in TENSOR_ORDER[:13]:
        off, shp = lay[name]; n = int(np.prod(shp))
        print("  %-16s rel_err %.3e  |want| %.3e" % (name, rel_err(g[off:off+n], k["grads"][off:off+n]), np.abs(k["grads"][off:off+n]).max()))
    c.close()
