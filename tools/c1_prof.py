import os, sys, time
sys.path.insert(0, os.getcwd())
os.environ["PPO_ROLLOUT_PROF"] = "1"
from ppo_cpp_b200 import core
for (h1, h2) in ((4, 5), (64, 64)):
    c = core.PPOCore(hidden1=h1, hidden2=h2, n_envs=1, n_steps=2048, nminibatches=32, noptepochs=10, seed=1)
    c.init_orthogonal(7); c.shuffle_seed(42); c.synth_env_reset()
    for i in range(3):
        c.sync(); t0 = time.perf_counter(); c.rollout_synthetic(); c.sync(); t1 = time.perf_counter()
        print(f"[{h1},{h2}] rollout {1e3*(t1-t0):.3f} ms = {1e6*(t1-t0)/2048:.2f} us per env step", flush=True)
    c.close()
