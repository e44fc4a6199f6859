"""GPU box: phase clocks of the persistent U-family epoch kernel (PPO_UMMA_PROF=1) on the C3 shape + ms per update."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
os.environ["PPO_UMMA_PROF"] = "1"
from ppo_cpp_b200 import core
c = core.PPOCore(hidden1=64, hidden2=64, n_envs=4096, n_steps=64, nminibatches=32, noptepochs=10, seed=1234)
c.init_orthogonal(7)
c.shuffle_seed(42)
c.synth_env_reset()
for i in range(6):
    c.rollout_synthetic()
    c.sync()
    t0 = time.perf_counter()
    c.train_update(3.9e-4, 0.161, want_losses=False)
    c.sync()
    t1 = time.perf_counter()
    print(f"update {i}: train {1e3 * (t1 - t0):.3f} ms", flush=True)
print("train_fwdbwd alone ms:", c.profile_kernel("train_fwdbwd", 50))
c.close()
