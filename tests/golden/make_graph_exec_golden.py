#!/usr/bin/env python
"""Golden fixtures computed by EXECUTING the reference's own TensorFlow graph file, node by node.

    python tests/golden/make_graph_exec_golden.py       (build container only: needs /root/reference)

writes tests/golden/graph_exec_act.npz and tests/golden/graph_exec_train.npz.  The interpreter is
oracle/tf_graph_exec.py; the graph is the one the reference ships and runs through tensorflow::Session
(resources/ppo_cl/graphs/ppo_cpp_[4_5]_lr_0.0004_cr_0.1610_ent_0.0007.meta.txt, "GRAPH"):

  act model     feed input/Ob:0 (+ the RandomStandardNormal output, whose graph seeds are 0/0) and fetch
                output/_action, _value_flat, _neglogp, _deterministic_action      (ppo2/policies.hpp:33-77)
  train model   feed the eight placeholders of ppo2/ppo2.hpp:430-440, fetch the five loss scalars, the 13
                gradient tensors before clip_by_global_norm (inputs of loss/global_norm/L2Loss*), after it
                (inputs 9 of the ApplyAdam ops), the global norm, and — after running the target ppo2/_train
                (GRAPH:31383) — every variable: weights, Adam m / v, beta1_power, beta2_power.
                Several consecutive train steps per case so the Adam state and the beta powers are pinned too.

Cases.  `init_4_5` / `ckpt_4_5`: the graph exactly as shipped (initial weights; shipped checkpoint weights).
`ties_*`: value head forced to v == 1 exactly (vf/w = 0, vf/b = 1) with old values / returns chosen so that
v - old_v == +-cliprange and (v-R)^2 == (vclip-R)^2 hold EXACTLY, plus zero advantages and ratio == 1: the
tie rules of the graph's Maximum/Minimum gradients (GRAPH:12609-12776, 14975-15142, 15357, 16113).
In every other case the advantages fed are the per-minibatch normalised `returns - old values` of ppo2/ppo2.hpp:401-406.
`rand_H_H`, `obs36_*`, `obs1_*`: the same graph executed with substituted variables of another width (the
graph's op structure is shape-generic; see dynamic_grad_shapes in the interpreter) — hidden [8,8], [64,64],
[128,128] and the observation widths of the reference's other envs (env/hexapod_env.hpp:226-238 -> 1,
env/hexapod_closed_loop_env.hpp:20,61-72 -> 36).

Each quantity is stored twice: `f32` (TF's own arithmetic type) and `f64` (every float tensor widened; the
truth that the 1e-5 tolerances are measured against).
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
import tf_graph_exec as tg  # noqa: E402
from make_golden import CKPT, GRAPH, minibatch, orthogonal_tensors  # noqa: E402
from ppo_cpp_b200.meta_graph import TENSOR_ORDER, parse_meta_txt, read_checkpoint_data  # noqa: E402

LR = np.float32(3.9e-4)


def executor(mg, dtype, tensors):
    ex = tg.GraphExecutor(mg, dtype, dynamic_grad_shapes=True)
    ex.init()
    for n in TENSOR_ORDER:
        ex.set_variable(n, tensors[n])
        if n + "/Adam" in ex.vars:      # fresh optimizer slots of the substituted shape (zeros, as `init` makes them)
            ex.set_variable(n + "/Adam", np.zeros_like(tensors[n]))
            ex.set_variable(n + "/Adam_1", np.zeros_like(tensors[n]))
    return ex


def flat(ex, suffix=""):
    return np.concatenate([np.asarray(ex.vars[n + suffix]).ravel() for n in TENSOR_ORDER[:13]])


def run_act(mg, tensors, obs, eps):
    out = {}
    for tag, dt in (("f32", np.float32), ("f64", np.float64)):
        ex = executor(mg, dt, tensors)
        a, v, nlp, det = ex.run(tg.ACT_FETCH, {tg.ACT_FEED: obs.astype(dt), tg.ACT_NOISE: eps.astype(dt)})
        out.update({f"action_{tag}": a, f"value_{tag}": v, f"neglogp_{tag}": nlp, f"mean_{tag}": det})
        (v2,) = ex.run(["output/_value_flat:0"], {tg.ACT_FEED: obs.astype(dt)})     # MlpPolicy::value
        assert np.array_equal(v, v2)
    return out


def run_train(mg, tensors, batches, cr, store=("f32", "f64"), light=False):
    """batches: list of (obs, act, adv, ret, old_nlp, old_v).  Returns per-step arrays stacked on axis 0."""
    pre = [n.input[0] for n in mg.graph_def.node if n.op == "L2Loss"]
    post = [n.input[9] for n in mg.graph_def.node if n.op == "ApplyAdam"]
    assert len(pre) == 13 and len(post) == 13
    out = {}
    for tag, dt in (("f32", np.float32), ("f64", np.float64)):
        if tag not in store:
            continue
        ex = executor(mg, dt, tensors)
        rec = {k: [] for k in ("losses", "grads", "grads_clipped", "gnorm", "theta", "m", "v", "bpow")}
        for obs, act, adv, ret, old_nlp, old_v in batches:
            F = tg.TRAIN_FEEDS
            feeds = {F["obs"]: obs, F["actions"]: act, F["advs"]: adv, F["returns"]: ret, F["lr"]: LR,
                     F["cliprange"]: np.float32(cr), F["old_neglogp"]: old_nlp, F["old_vpred"]: old_v}
            feeds = {k: np.asarray(v).astype(dt) for k, v in feeds.items()}
            # one Session::Run: fetches and the train target together, as ppo2.hpp:450 does
            res = ex.run(list(tg.LOSS_FETCH) + pre + post + ["loss/global_norm/global_norm:0"], feeds, [tg.TRAIN_TARGET])
            rec["losses"].append(np.array([float(x) for x in res[:5]]))
            rec["grads"].append(np.concatenate([r.ravel() for r in res[5:18]]))
            rec["grads_clipped"].append(np.concatenate([r.ravel() for r in res[18:31]]))
            rec["gnorm"].append(float(res[31]))
            rec["theta"].append(flat(ex))
            rec["m"].append(flat(ex, "/Adam"))
            rec["v"].append(flat(ex, "/Adam_1"))
            rec["bpow"].append(np.array([float(ex.vars["beta1_power"]), float(ex.vars["beta2_power"])]))
        keep = ("losses", "grads", "gnorm", "theta", "bpow") if light else tuple(rec)
        for k in keep:
            out[f"{k}_{tag}"] = np.stack([np.asarray(x, dt if k not in ("losses", "gnorm", "bpow") else np.float64) for x in rec[k]])
    return out


def with_normalised_advantages(batch):
    """advs exactly as PPO2::_train_step makes them from returns and old values (ppo2/ppo2.hpp:401-406, fp32 Eigen)."""
    obs, act, _, ret, old_nlp, old_v = batch
    advs = ret - old_v
    mean = np.float32(advs.mean(dtype=np.float32))
    var = np.float32(((advs - mean) ** 2).sum(dtype=np.float32) / np.float32(advs.size))
    advs = (advs - mean) / np.float32(np.sqrt(var) + 1e-8)       # float(sqrt(var) + 1e-8): the sum is formed in double
    return obs, act, advs.astype(np.float32), ret, old_nlp, old_v


def tie_batch(rng, t, mg, B, cr):
    """Minibatch whose value-loss terms tie exactly (v == 1 for every sample, see module docstring)."""
    obs, act, adv, ret, old_nlp, old_v = minibatch(rng, t, B)
    one = np.float32(1.0)
    old_v[0:8], ret[0:8] = one - np.float32(cr), rng.standard_normal(8).astype(np.float32)      # v - old_v == +cr
    old_v[8:16], ret[8:16] = one + np.float32(cr), rng.standard_normal(8).astype(np.float32)    # v - old_v == -cr
    old_v[16:28], ret[16:28] = 0.0, 0.625            # vclip = 0.25: v-R = .375, vclip-R = -.375  -> l1 == l2
    old_v[28:32], ret[28:32] = 2.0, 1.375            # vclip = 1.75: v-R = -.375, vclip-R = .375 -> l1 == l2
    adv[32:40] = 0.0                                 # -A*ratio == -A*clip(ratio) == -0
    ex = executor(mg, np.float32, t)                 # ratio == 1 in fp32: old_neglogp = the graph's own neglogp
    (nlp,) = ex.run(["loss/Add_1:0"], {tg.TRAIN_FEEDS["obs"]: obs, tg.TRAIN_FEEDS["actions"]: act})   # GRAPH: neglogp of the train model
    old_nlp[40:48] = nlp[40:48]
    assert cr == 0.25
    return obs, act, adv, ret, old_nlp, old_v


def main():
    rng = np.random.default_rng(20261017)
    mg = tg.load_meta_graph(GRAPH)
    meta = parse_meta_txt(GRAPH)
    ck = read_checkpoint_data(CKPT + ".data-00000-of-00001", meta.shapes)
    with open(GRAPH, "rb") as f:
        sha = hashlib.sha256(f.read()).hexdigest()

    nets = {
        "init_4_5": dict(meta.tensors), "ckpt_4_5": dict(ck),
        "rand_8_8": orthogonal_tensors(rng, 18, 18, 8, 8), "rand_64_64": orthogonal_tensors(rng, 18, 18, 64, 64),
        "rand_128_128": orthogonal_tensors(rng, 18, 18, 128, 128),
        "obs36_4_5": orthogonal_tensors(rng, 36, 18, 4, 5), "obs1_4_5": orthogonal_tensors(rng, 1, 18, 4, 5),
        "obs36_64_64": orthogonal_tensors(rng, 36, 18, 64, 64),
    }
    for base in ("ckpt_4_5", "rand_64_64"):
        t = {k: v.copy() for k, v in nets[base].items()}
        t["model/vf/w"] = np.zeros_like(t["model/vf/w"])
        t["model/vf/b"] = np.ones_like(t["model/vf/b"])
        nets["ties_" + base.split("_", 1)[1]] = t

    act, train = {"graph_sha256": np.array(sha)}, {"graph_sha256": np.array(sha)}
    for name, t in nets.items():
        O = t["model/pi_fc0/w"].shape[0]
        n = 16
        obs = rng.standard_normal((n, O)).astype(np.float32)
        obs[0] = 0.0
        eps = rng.standard_normal((n, 18)).astype(np.float32)
        d = dict(params=np.concatenate([t[k].ravel() for k in TENSOR_ORDER]).astype(np.float32),
                 dims=np.array([O, 18, t["model/pi_fc0/w"].shape[1], t["model/pi_fc1/w"].shape[1]]), obs=obs, eps=eps)
        d.update(run_act(mg, t, obs, eps))
        act.update({f"{name}__{k}": v for k, v in d.items()})

    plan = {  # name: (B, cliprange, steps, light)
        "init_4_5": (96, 0.2, 3, False), "ckpt_4_5": (257, 0.16102319955825806, 3, False), "ties_4_5": (64, 0.25, 1, False),
        "rand_8_8": (33, 0.1, 3, False), "rand_64_64": (128, 0.2, 2, False), "ties_64_64": (128, 0.25, 1, True),
        "obs36_4_5": (64, 0.2, 2, False), "obs1_4_5": (64, 0.2, 2, False), "obs36_64_64": (128, 0.2, 1, True),
        "rand_128_128": (256, 0.2, 1, True),
    }
    for name, (B, cr, steps, light) in plan.items():
        t = nets[name]
        if name.startswith("ties_"):
            batches = [tie_batch(rng, t, mg, B, cr)]
        else:
            batches = [with_normalised_advantages(minibatch(rng, t, B)) for _ in range(steps)]
        d = dict(params=np.concatenate([t[k].ravel() for k in TENSOR_ORDER]).astype(np.float32),
                 dims=np.array([t["model/pi_fc0/w"].shape[0], 18, t["model/pi_fc0/w"].shape[1], t["model/pi_fc1/w"].shape[1]]),
                 cliprange=np.float64(np.float32(cr)), lr=np.float64(LR))
        for i, f in enumerate(("obs", "act", "adv", "ret", "old_nlp", "old_v")):
            d[f] = np.stack([b[i] for b in batches])
        d.update(run_train(mg, t, batches, cr, store=("f64",) if light else ("f32", "f64"), light=light))
        train.update({f"{name}__{k}": v for k, v in d.items()})
        print(name, "losses", d["losses_f64"][0], "gnorm", d["gnorm_f64"][0])

    np.savez_compressed(HERE + "/graph_exec_act.npz", **act)
    np.savez_compressed(HERE + "/graph_exec_train.npz", **train)
    for fn in ("graph_exec_act.npz", "graph_exec_train.npz"):
        print(fn, os.path.getsize(HERE + "/" + fn) // 1024, "KiB")


if __name__ == "__main__":
    main()
