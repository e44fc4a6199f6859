import sys, os, ctypes as C, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
from ppo_cpp_b200 import core
from conftest import load_weights
_, flat = load_weights("ckpt_71_weights.npz")
n_envs, n_steps, seed = 32, 400, 2024
lib = ol.load()
d = ol.LearnerDesc(ol.Dims(18,18,4,5), ol.HParams(0.0007160293171182275,0.5,0.5,0.9,0.999,1e-5), n_envs,n_steps,4,1,0.99,0.95,3.9e-4,0.2,seed,42,0,1)
L = lib.oracle_learner_create(C.byref(d), flat); lib.oracle_learner_rollout(L)
nb = n_envs*n_steps
names = ["obs","returns","dones","actions","values","neglogpacs","true_rewards","unnormalized_rewards"]; widths=[18,1,1,18,1,1,1,1]
want = {n: np.ctypeslib.as_array(lib.oracle_learner_buffer(L,i),(nb,w)).copy() for i,(n,w) in enumerate(zip(names,widths))}
c = core.PPOCore(n_envs=n_envs, n_steps=n_steps, nminibatches=4, noptepochs=1, seed=seed); c.set_tensor("params", flat)
c.synth_env_reset(); c.rollout_synthetic()
for n in names:
    g = c.rollout_get(n).reshape(n_envs, n_steps, -1); w = want[n].reshape(n_envs, n_steps, -1)
    err_t = np.abs(g-w).max(axis=(0,2))
    print("%-22s max|w| %.3f  err at t=0,1,2,10,50,100,200,333,334,335,399: %s" % (n, np.abs(w).max(), " ".join("%.1e"%err_t[t] for t in (0,1,2,10,50,100,200,333,334,335,399))))
    if n == "obs":
        e = np.abs(g-w).max(axis=2); env, t = np.unravel_index(np.argmax(e), e.shape); print("   worst env %d t %d" % (env,t), g[env,t,:4], w[env,t,:4])
