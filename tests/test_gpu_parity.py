"""GPU parity tests: every entry point of the C ABI against the CPU oracle / the golden fixtures.

Bar (BASELINE.json north_star): bit-exact for permutation indices and done masks; <= 1e-5 relative
(tensor-wise: max|a-b| / max|b|) for log-probs, GAE, losses, gradients and one Adam step.
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as ol
from conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope="module")
def core_mod():
    from ppo_cpp_b200 import core
    return core


def make_core(core_mod, flat=None, **kw):
    c = core_mod.PPOCore(**kw)
    if flat is not None:
        c.set_tensor("params", flat)
    return c


def rand_params(rng, h1, h2, O=18, A=18):
    o = ol.Oracle(h1=h1, h2=h2)
    # keep pre-activations O(1) for wide layers (a saturated tanh net has ill-conditioned gradients in any fp32 code)
    p = (rng.standard_normal(o.Pq) * min(0.3, 1.5 / np.sqrt(max(h1, h2)))).astype(np.float32)
    p[o.offset(12):o.offset(13)] = (-0.5 + 0.2 * rng.standard_normal(A)).astype(np.float32)  # logstd
    return p


# ------------------------------------------------------------------ policy step (a1-a3)
@pytest.mark.parametrize("wname", ["init", "ckpt"])
def test_policy_step_golden(core_mod, wname, forward_kat, init_weights, ckpt_weights):
    _, flat = init_weights if wname == "init" else ckpt_weights
    k = forward_kat[wname]
    c = make_core(core_mod, flat, n_envs=8, n_steps=4, nminibatches=4)
    act, val, nlp = c.policy_step(k["obs"], k["eps"])
    assert rel_err(act, k["action"]) < TOL and rel_err(val, k["value"]) < TOL and rel_err(nlp, k["neglogp"]) < TOL
    assert rel_err(c.policy_mean(k["obs"]), k["mean"]) < TOL
    assert rel_err(c.policy_value(k["obs"]), k["value"]) < TOL
    c.close()


@pytest.mark.parametrize("h1,h2,n", [(4, 5, 1), (4, 5, 1000), (8, 8, 129), (64, 64, 4096), (64, 32, 77), (256, 256, 300), (5, 3, 65),
                                     (256, 256, 3000), (128, 128, 1024)])  # the last two: W-family forward (tcgen05 GEMMs)
def test_policy_step_vs_oracle(core_mod, h1, h2, n):
    rng = np.random.default_rng(h1 * 1000 + h2 + n)
    p = rand_params(rng, h1, h2)
    obs = rng.standard_normal((n, 18)).astype(np.float32)
    eps = rng.standard_normal((n, 18)).astype(np.float32)
    c = make_core(core_mod, p, hidden1=h1, hidden2=h2, n_envs=4, n_steps=8, nminibatches=4)
    act, val, nlp = c.policy_step(obs, eps)
    o = ol.Oracle(h1=h1, h2=h2)
    a64, v64, n64, m64 = o.policy_step(p, obs, eps, "f64")
    assert rel_err(act, a64) < TOL and rel_err(val, v64) < TOL and rel_err(nlp, n64) < TOL
    assert rel_err(c.policy_mean(obs), m64) < TOL
    c.close()


def test_policy_step_philox_noise(core_mod, ckpt_weights):
    """eps == NULL: noise is Philox4x32-10 on (seed, env id, step counter); same stream as the oracle's."""
    _, flat = ckpt_weights
    n, seed = 500, 0xABCDEF0123
    rng = np.random.default_rng(5)
    obs = rng.standard_normal((n, 18)).astype(np.float32)
    c = make_core(core_mod, flat, n_envs=4, n_steps=8, nminibatches=4, seed=seed)
    lib = ol.load()
    o = ol.Oracle()
    std = np.exp(flat[o.offset(12):o.offset(13)].astype(np.float64))
    mean = o.policy_step(flat, obs, None, "f64")[3]
    for step in range(3):  # the step counter advances by one per call
        act, _, nlp = c.policy_step(obs, None)
        eps = np.zeros((n, 18), np.float32)
        for e in range(n):
            lib.oracle_normal_eps(seed, e, step, 18, eps[e])
        assert np.max(np.abs((act - mean) / std - eps)) < 2e-5
        a64, _, n64, _ = o.policy_step(flat, obs, eps, "f64")
        assert rel_err(nlp, n64) < 5e-5  # eps itself carries ~1e-7 logf/sinf differences, amplified by 1/std^2
    c.close()


# ------------------------------------------------------------------ VecNormalize (a6-a8)
def test_vecnorm_sequence_vs_oracle(core_mod):
    rng = np.random.default_rng(11)
    N, D, K = 96, 18, 12
    c = make_core(core_mod, None, n_envs=N, n_steps=4, nminibatches=4)
    lib = ol.load()
    vn = lib.oracle_vecnorm_create(N, D, 1)
    want_obs, want_rew = np.zeros((N, D), np.float32), np.zeros(N, np.float32)
    raw0 = (rng.standard_normal((N, D)) * 3 + 1).astype(np.float32)
    lib.oracle_vecnorm_reset_f32(vn, raw0, want_obs)
    got = c.vecnorm_reset(raw0)
    assert rel_err(got, want_obs) < TOL
    for k in range(K):
        raw = (rng.standard_normal((N, D)) * (1 + k) + 0.5 * k).astype(np.float32)
        rew = (rng.standard_normal(N) * 5).astype(np.float32)
        done = (rng.random(N) < 0.2).astype(np.float32)
        lib.oracle_vecnorm_step_f32(vn, raw, rew, done, want_obs, want_rew)
        obs, r = c.vecnorm_step(raw, rew, done)
        assert rel_err(obs, want_obs) < TOL and rel_err(r, want_rew) < TOL
    st = c.vecnorm_stats()
    v = vn.contents
    assert st["obs_count"] == v.obs_rms.count and st["ret_count"] == v.ret_rms.count  # exact doubles
    assert st["obs_count"] == pytest.approx(1e-6 + N * (K + 1))
    assert rel_err(st["obs_mean"], np.ctypeslib.as_array(v.obs_rms.mean, (D,))) < TOL
    assert rel_err(st["obs_var"], np.ctypeslib.as_array(v.obs_rms.var, (D,))) < TOL
    assert rel_err(st["ret_var"], np.ctypeslib.as_array(v.ret_rms.var, (1,))) < TOL
    lib.oracle_vecnorm_destroy(vn)
    # frozen statistics (training = false, playback): outputs keep following the oracle, counts stay
    c.vecnorm_set_training(0)
    c.vecnorm_step(raw, rew, done)
    assert c.vecnorm_stats()["obs_count"] == st["obs_count"]
    c.close()


def test_vecnorm_clip_and_constant_input(core_mod):
    """EnvMock-style constant observation (C1): mean -> obs, var -> ~0, output stays finite and inside the clip."""
    N, D = 4, 18
    c = make_core(core_mod, None, n_envs=N, n_steps=4, nminibatches=4)
    out = c.vecnorm_reset(np.ones((N, D), np.float32))
    for _ in range(5):
        out, r = c.vecnorm_step(np.ones((N, D), np.float32), np.ones(N, np.float32), np.zeros(N, np.float32))
        assert np.all(np.isfinite(out)) and np.abs(out).max() <= 10 and np.abs(r).max() <= 10
    st = c.vecnorm_stats()
    assert np.allclose(st["obs_mean"], 1.0, atol=1e-5)
    c.vecnorm_set_training(0)  # frozen statistics: an outlier must hit the clip (matrix_clamp.hpp:32-35)
    big = c.vecnorm_step(np.full((N, D), 1e6, np.float32), np.full(N, 1e9, np.float32), np.zeros(N, np.float32))
    assert big[0].max() == 10.0 and big[1].max() == 10.0
    c.close()


@pytest.mark.parametrize("T,N", [(1, 1), (40, 7), (16, 4096), (300, 33)])
def test_vecnorm_replay_equals_stepwise(core_mod, T, N):
    """ppo_vecnorm_replay (a recorded trajectory in four HBM-bound launches: per-env return scan, per-step moments of all
    steps at once, the T Chan merges replayed by one warp, normalise + clip) against T calls of ppo_vecnorm_step, and the
    stepwise path is itself checked against the oracle above.  Episodes end inside the trajectory (done resets the
    discounted return); a second replay continues from the carried state."""
    rng = np.random.default_rng(T * 1000 + N)
    raw = (3.0 + 2.0 * rng.standard_normal((2 * T, N, 18))).astype(np.float32)
    raw[:, :, 5] *= 40.0  # exercises the clip
    rew = rng.standard_normal((2 * T, N)).astype(np.float32)
    done = (rng.random((2 * T, N)) < 0.1).astype(np.float32)
    a = make_core(core_mod, None, n_envs=N, n_steps=8, nminibatches=1)
    b = make_core(core_mod, None, n_envs=N, n_steps=8, nminibatches=1)
    a.vecnorm_reset(raw[0])
    b.vecnorm_reset(raw[0])
    for half in range(2):
        sl = slice(half * T, (half + 1) * T)
        want = [a.vecnorm_step(raw[t], rew[t], done[t]) for t in range(half * T, (half + 1) * T)]
        obs, r = b.vecnorm_replay(raw[sl], rew[sl], done[sl])
        assert rel_err(obs, np.stack([w[0] for w in want])) < TOL
        assert rel_err(r, np.stack([w[1] for w in want])) < TOL
        sa, sb = a.vecnorm_stats(), b.vecnorm_stats()
        for k in ("obs_mean", "obs_var", "ret_mean", "ret_var"):
            assert rel_err(np.atleast_1d(sb[k]), np.atleast_1d(sa[k])) < TOL, k
        assert sa["obs_count"] == sb["obs_count"] and sa["ret_count"] == sb["ret_count"]
    a.close()
    b.close()


def test_running_stats_and_clamp_standalone(core_mod):
    rng = np.random.default_rng(12)
    c = make_core(core_mod, None, n_envs=4, n_steps=4, nminibatches=4)
    lib = ol.load()
    for D, rows in ((18, 1), (18, 5000), (1, 333), (7, 64)):
        batch = (rng.standard_normal((rows, D)) * 2 + 3).astype(np.float32)
        mean, var = rng.standard_normal(D).astype(np.float32), (rng.random(D) + 0.5).astype(np.float32)
        m, v, cnt = c.running_stats_update(mean, var, 123.5, batch)
        wm, wv, wc = mean.astype(np.float64), var.astype(np.float64), C.c_double(123.5)
        lib.oracle_rstats_update_f64(wm, wv, C.byref(wc), D, batch.astype(np.float64), rows)
        assert cnt == wc.value and rel_err(m, wm) < TOL and rel_err(v, wv) < TOL
    x = (rng.standard_normal(10001) * 20).astype(np.float32)
    want = np.zeros_like(x)
    lib.oracle_matrix_clamp_f32(x, x.size, -10.0, 10.0, want)
    assert np.array_equal(c.matrix_clamp(x, -10.0, 10.0), want)  # bit-exact
    c.close()


# ------------------------------------------------------------------ GAE (a5)
@pytest.mark.parametrize("T,N", [(1, 1), (64, 4096), (2048, 1), (5000, 3), (333, 257)])
def test_gae_vs_oracle(core_mod, T, N):
    rng = np.random.default_rng(T * 7 + N)
    rew, val = rng.standard_normal((T, N)).astype(np.float32), rng.standard_normal((T, N)).astype(np.float32)
    done = (rng.random((T, N)) < 1 / 50).astype(np.float32)
    lv, ld = rng.standard_normal(N).astype(np.float32), (rng.random(N) < 0.3).astype(np.float32)
    c = make_core(core_mod, None, n_envs=4, n_steps=4, nminibatches=4)
    g, lam = np.float32(0.99), np.float32(0.95)
    adv, ret = c.gae(rew, val, done, lv, ld, g, lam)
    o = ol.Oracle()
    a64, r64 = o.gae(rew, val, done, lv, ld, float(g), float(lam), "f64")
    assert rel_err(adv, a64) < TOL and rel_err(ret, r64) < TOL
    a32, r32 = o.gae(rew, val, done, lv, ld, g, lam, "f32")
    assert np.array_equal(adv, a32) and np.array_equal(ret, r32)  # same fp32 operation order: bit-exact
    c.close()


@pytest.mark.parametrize("gamma,lam", [(0.999, 1.0), (1.0, 1.0), (0.99, 0.95), (0.9999, 0.97), (0.5, 0.5)])
@pytest.mark.parametrize("T,N,pdone", [(65536, 1, 1 / 334), (65536, 3, 1 / 334), (65536, 3, 0.0), (20000, 2, 1 / 50)])
def test_gae_chunked_any_gamma_lambda(core_mod, gamma, lam, T, N, pdone):
    """Few envs x long rollouts are scanned in chunks (the C2 shape, 1 x 65 536).  The carried lastgaelam must be right for
    every gamma*lam (runner.hpp:159-191 takes them from the constructor): warm-up chunks when gamma*lam contracts, exact
    affine carries when it is close to or equal to 1."""
    rng = np.random.default_rng(int(gamma * 1e4) + T + N)
    rew = (0.1 * rng.standard_normal((T, N))).astype(np.float32)
    val = rng.standard_normal((T, N)).astype(np.float32)
    done = (rng.random((T, N)) < pdone).astype(np.float32)
    lv, ld = rng.standard_normal(N).astype(np.float32), np.zeros(N, np.float32)
    c = make_core(core_mod, None, n_envs=4, n_steps=4, nminibatches=4)
    g, l = np.float32(gamma), np.float32(lam)
    adv, ret = c.gae(rew, val, done, lv, ld, g, l)
    o = ol.Oracle()
    a32, r32 = o.gae(rew, val, done, lv, ld, g, l, "f32")
    a64, r64 = o.gae(rew, val, done, lv, ld, float(g), float(l), "f64")
    # the fp32 sequential scan itself drifts from the exact recurrence when nothing contracts (gamma = lam = 1, no dones:
    # a 65 536-term running sum); the kernel must be at least as close to the truth as the reference's own fp32 order is
    ref_drift = rel_err(a32, a64)
    assert rel_err(adv, a64) < max(TOL, 2 * ref_drift) and rel_err(ret, r64) < max(TOL, 2 * ref_drift)
    assert rel_err(adv, a32) < max(TOL, 2 * ref_drift)
    if float(g) * float(l) < 0.99:   # contracting: warm-up chunks reproduce the sequential fp32 scan bit for bit
        assert np.array_equal(adv, a32) and np.array_equal(ret, r32)
    c.close()


def test_gae_full_size_properties(core_mod):
    """C5 size (16 M transitions): size-independent properties instead of an oracle run."""
    import torch
    T, N = 256, 65536
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(0)
    rew = torch.randn(T, N, device=dev, generator=gen)
    val = torch.randn(T, N, device=dev, generator=gen)
    lv = torch.randn(N, device=dev, generator=gen)
    c = make_core(core_mod, None, n_envs=4, n_steps=4, nminibatches=4)
    adv, ret = torch.empty_like(rew), torch.empty_like(rew)
    ones, zeros = torch.ones(T, N, device=dev), torch.zeros(N, device=dev)
    torch.cuda.synchronize()
    # every step terminal: adv = rew - val exactly (except the last row, which bootstraps with last_dones=0)
    c.gae_device(rew, val, ones, lv, torch.ones(N, device=dev), 0.99, 0.95, adv, ret)
    c.sync()
    assert torch.equal(adv, rew - val) and torch.equal(ret, (rew - val) + val)
    # gamma = 0: adv = rew - val regardless of dones
    c.gae_device(rew, val, torch.zeros(T, N, device=dev), lv, zeros, 0.0, 0.95, adv, ret)
    c.sync()
    assert torch.equal(adv, rew - val)
    # lam = 1, no dones, gamma = 1: adv_t = sum_{s>=t} rew_s + last_v - val_t  (telescoping) — checked in fp64
    c.gae_device(rew, val, torch.zeros(T, N, device=dev), lv, zeros, 1.0, 1.0, adv, ret)
    c.sync()
    want = torch.flip(torch.cumsum(torch.flip(rew.double(), [0]), 0), [0]) + lv.double() - val.double()
    assert float((adv.double() - want).abs().max() / want.abs().max()) < TOL
    c.close()


# ------------------------------------------------------------------ advantage normalisation (a10)
@pytest.mark.parametrize("n", [2, 64, 2048, 131072])
def test_advnorm_vs_oracle(core_mod, n):
    rng = np.random.default_rng(n)
    r, v = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    c = make_core(core_mod, None, n_envs=4, n_steps=4, nminibatches=4)
    assert rel_err(c.advnorm(r, v), ol.Oracle().advnorm(r, v, "f64")) < TOL
    c.close()


# ------------------------------------------------------------------ loss + gradient (a11)
@pytest.mark.parametrize("case", ["init_4_5", "ckpt_4_5", "rand_64_64", "rand_8_8"])
def test_loss_grad_golden_autograd(core_mod, case, loss_kat, kat):
    k = loss_kat[case]
    h1, h2 = (int(x) for x in k["hidden"])
    cst = kat["consts"]
    c = make_core(core_mod, k["params"], hidden1=h1, hidden2=h2, n_envs=4, n_steps=8, nminibatches=4,
                  ent_coef=cst["ent_coef"], vf_coef=cst["vf_coef"])
    g, l = c.loss_grad(k["obs"], k["act"], k["adv"], k["ret"], k["old_nlp"], k["old_v"], float(k["cliprange"]))
    assert rel_err(g, k["grads"]) < TOL
    assert np.allclose(l, k["losses"], rtol=TOL, atol=1e-7)
    from ppo_cpp_b200.meta_graph import TENSOR_ORDER, param_layout
    lay = param_layout(18, 18, h1, h2)
    for name in TENSOR_ORDER[:13]:  # per tensor, so small tensors cannot hide behind large ones
        off, shp = lay[name]
        n = int(np.prod(shp))
        assert rel_err(g[off:off + n], k["grads"][off:off + n]) < 5 * TOL, name
    c.close()


# ---- the same entry points against fixtures made by EXECUTING the reference's TensorFlow graph file node by node
# (oracle/tf_graph_exec.py; forward GRAPH:1859-6866, loss 9210-11752, autodiff 11773-23699, clip 23738-25392,
# ApplyAdam 25426-31383).  This is the parity pin: nothing below depends on a restatement of the math.
GX_ALL = ["init_4_5", "ckpt_4_5", "ties_4_5", "rand_8_8", "rand_64_64", "ties_64_64", "rand_128_128", "obs36_4_5", "obs1_4_5",
          "obs36_64_64"]


def _gx_core(core_mod, k, kat, **kw):
    O, A, h1, h2 = (int(x) for x in k["dims"])
    cst = kat["consts"]
    return make_core(core_mod, k["params"], obs_dim=O, act_dim=A, hidden1=h1, hidden2=h2, ent_coef=cst["ent_coef"],
                     vf_coef=cst["vf_coef"], max_grad_norm=cst["clip_norm"], adam_beta1=cst["beta1"], adam_beta2=cst["beta2"],
                     adam_epsilon=cst["adam_eps"], **kw)


@pytest.mark.parametrize("case", GX_ALL)
def test_policy_step_vs_executed_graph(core_mod, case, gx_act, kat):
    k = gx_act[case]
    c = _gx_core(core_mod, k, kat, n_envs=8, n_steps=4, nminibatches=4)
    act, val, nlp = c.policy_step(k["obs"], k["eps"])
    assert rel_err(act, k["action_f64"]) < TOL and rel_err(val, k["value_f64"]) < TOL and rel_err(nlp, k["neglogp_f64"]) < TOL
    assert rel_err(c.policy_mean(k["obs"]), k["mean_f64"]) < TOL
    assert rel_err(c.policy_value(k["obs"]), k["value_f64"]) < TOL
    c.close()


@pytest.mark.parametrize("case", GX_ALL)
def test_loss_grad_vs_executed_graph(core_mod, case, gx_train, kat):
    """Losses and all 13 gradient tensors (before the clip) of the graph's first train step, ties included."""
    from ppo_cpp_b200.meta_graph import TENSOR_ORDER, param_layout
    k = gx_train[case]
    c = _gx_core(core_mod, k, kat, n_envs=4, n_steps=8, nminibatches=4)
    g, l = c.loss_grad(k["obs"][0], k["act"][0], k["adv"][0], k["ret"][0], k["old_nlp"][0], k["old_v"][0], float(k["cliprange"]))
    want = k["grads_f64"][0]
    assert rel_err(g, want) < TOL
    O, A, h1, h2 = (int(x) for x in k["dims"])
    lay = param_layout(O, A, h1, h2)
    for name in TENSOR_ORDER[:13]:
        off, shp = lay[name]
        n = int(np.prod(shp))
        assert rel_err(g[off:off + n], want[off:off + n]) < TOL, name
    assert np.allclose(l, k["losses_f64"][0], rtol=TOL, atol=2e-7)
    c.close()


@pytest.mark.parametrize("case", ["init_4_5", "ckpt_4_5", "rand_8_8", "rand_64_64", "obs36_4_5", "obs1_4_5"])
def test_train_steps_vs_executed_graph(core_mod, case, gx_train, kat):
    """Consecutive `ppo2/_train` runs of the graph (GRAPH:31383): the rollout buffers hold the fixture's minibatches back
    to back, the permutation is the identity, minibatch k is train step k.  Checks the per-minibatch advantage
    normalisation (ppo2.hpp:401-406), losses, unclipped gradient, and after every step the weights, Adam m / v and
    the beta powers against the graph's variables."""
    k = gx_train[case]
    steps, B = k["obs"].shape[:2]
    c = _gx_core(core_mod, k, kat, n_envs=1, n_steps=steps * B, nminibatches=steps, noptepochs=1)
    P = c.P
    flat = lambda a: np.ascontiguousarray(a.reshape(steps * B, -1))
    for name, f in (("obs", "obs"), ("actions", "act"), ("returns", "ret"), ("values", "old_v"), ("neglogpacs", "old_nlp")):
        c.rollout_set(name, flat(k[f]))
    c.train_set_permutation(np.arange(steps * B, dtype=np.int32))
    th0 = k["params"][:P].astype(np.float64)
    for s in range(steps):
        losses, grads = c.train_minibatch(s, float(k["lr"]), float(k["cliprange"]))
        assert rel_err(grads, k["grads_f64"][s]) < (TOL if s == 0 else 3 * TOL), s   # later steps start from fp32-rounded weights
        assert np.allclose(losses, k["losses_f64"][s], rtol=3 * TOL, atol=2e-6), (s, losses, k["losses_f64"][s])
        got = c.get_tensor("params")
        assert rel_err(got[:P], k["theta_f64"][s]) < 1e-6, s
        assert rel_err(got[:P] - th0, k["theta_f64"][s] - th0) < 1e-3, s       # the accumulated movement itself (fp32 ulp-limited)
        assert np.array_equal(got[P:], k["params"][P:])
        assert rel_err(c.get_tensor("adam_m"), k["m_f64"][s]) < 3 * TOL and rel_err(c.get_tensor("adam_v"), k["v_f64"][s]) < 3 * TOL
        assert c.get_tensor("beta1_power")[0] == pytest.approx(k["bpow_f64"][s][0], rel=1e-6)
        assert c.get_tensor("beta2_power")[0] == pytest.approx(k["bpow_f64"][s][1], rel=1e-6)
    c.close()


@pytest.mark.parametrize("h1,h2,B", [(4, 5, 64), (4, 5, 2048), (64, 64, 8192), (256, 256, 512), (12, 20, 100), (5, 3, 37)])
def test_loss_grad_vs_oracle(core_mod, h1, h2, B):
    rng = np.random.default_rng(B + h1)
    p = rand_params(rng, h1, h2)
    o = ol.Oracle(h1=h1, h2=h2)
    obs = rng.standard_normal((B, 18)).astype(np.float32)
    _, v64, _, m64 = o.policy_step(p, obs, None, "f64")
    std = np.exp(p[o.offset(12):o.offset(13)].astype(np.float64))
    act = (m64 + std * rng.standard_normal((B, 18))).astype(np.float32)
    z = (act - m64) / std
    old_nlp = (0.5 * (z * z).sum(1) + 18 * 0.9189385175704956 + np.log(std).sum() + 0.1 * rng.standard_normal(B)).astype(np.float32)
    old_v = (v64 + 0.3 * rng.standard_normal(B)).astype(np.float32)
    ret = (v64 + 0.5 * rng.standard_normal(B)).astype(np.float32)
    adv = rng.standard_normal(B).astype(np.float32)
    c = make_core(core_mod, p, hidden1=h1, hidden2=h2, n_envs=4, n_steps=8, nminibatches=4)
    g, l = c.loss_grad(obs, act, adv, ret, old_nlp, old_v, 0.2)
    g64, l64 = o.loss_grad(p, obs, act, adv, ret, old_nlp, old_v, 0.2, "f64")
    assert rel_err(g, g64) < TOL
    for t in range(13):  # every gradient tensor separately
        sl = slice(o.offset(t), o.offset(t + 1))
        assert rel_err(g[sl], g64[sl]) < TOL, f"gradient tensor {t}"
    assert np.allclose(l, l64, rtol=TOL, atol=1e-7)
    assert l[4] == pytest.approx(l64[4], abs=1.5 / B)  # clipfrac is a count
    c.close()


@pytest.mark.parametrize("B", [1000, 128, 77, 20000])
def test_kernel_families_agree(core_mod, monkeypatch, B):
    """[64,64] trains on the tcgen05 (U) family by default; the fused FFMA (F) family and the generic tile (T) family
    must give the same answers (all are also checked against the oracle above).  B = 1000 / 77 leave a ragged last
    tile, B = 20000 makes every CTA of the U family loop over several tiles (accumulation in TMEM)."""
    rng = np.random.default_rng(99)
    h1 = h2 = 64
    p = rand_params(rng, h1, h2)
    obs = rng.standard_normal((B, 18)).astype(np.float32)
    eps = rng.standard_normal((B, 18)).astype(np.float32)
    adv = rng.standard_normal(B).astype(np.float32)
    res = []
    for env in (None, "PPO_DISABLE_UMMA", "PPO_DISABLE_FUSED"):
        if env:
            monkeypatch.setenv(env, "1")
        c = make_core(core_mod, p, hidden1=h1, hidden2=h2, n_envs=4, n_steps=8, nminibatches=4)
        act, val, nlp = c.policy_step(obs, eps)
        g, l = c.loss_grad(obs, act, adv, val + 0.1, nlp + 0.01, val - 0.05, 0.2)
        res.append((act, val, nlp, g, l))
        c.close()
        if env:
            monkeypatch.delenv(env)
    names = ["actions", "values", "neglogp", "grads", "losses"]
    for other in res[1:]:
        for nm, a, b in zip(names, res[0], other):
            assert rel_err(a, b) < 5e-6, nm
    # every gradient tensor separately (a small tensor must not hide behind a big one)
    o = ol.Oracle(h1=h1, h2=h2)
    for t in range(13):
        sl = slice(o.offset(t), o.offset(t + 1))
        assert rel_err(res[0][3][sl], res[2][3][sl]) < 1e-5, f"gradient tensor {t}"


@pytest.mark.parametrize("B", [1, 31, 64, 1000, 50000])
def test_small_family_equals_tile_family(core_mod, monkeypatch, init_weights, B):
    """The reference's [4,5] net trains on the thread-per-sample S family by default; the generic tile (T) family must
    give the same loss and gradient (both are also checked against the oracle).  B = 1 / 31 / 1000 leave ragged warps,
    B = 50000 makes every warp loop over several 32-sample tiles."""
    rng = np.random.default_rng(17 + B)
    _, flat = init_weights
    p = flat + (0.1 * rng.standard_normal(flat.size)).astype(np.float32)
    obs = rng.standard_normal((B, 18)).astype(np.float32)
    eps = rng.standard_normal((B, 18)).astype(np.float32)
    adv = rng.standard_normal(B).astype(np.float32)
    res = []
    for env in (None, "PPO_DISABLE_SMALL"):
        if env:
            monkeypatch.setenv(env, "1")
        c = make_core(core_mod, p, n_envs=4, n_steps=8, nminibatches=4)
        assert ("train_small_kernel" in c.kernel_family("train")) == (env is None)
        act, val, nlp = c.policy_step(obs, eps)
        g, l = c.loss_grad(obs, act, adv, val + 0.1, nlp + 0.01, val - 0.05, 0.2)
        res.append((g, l))
        c.close()
        if env:
            monkeypatch.delenv(env)
    assert rel_err(res[0][1], res[1][1]) < 5e-6
    o = ol.Oracle(h1=4, h2=5)
    for t in range(13):
        sl = slice(o.offset(t), o.offset(t + 1))
        # two fp32 evaluation orders of a sum with cancellation (the 1e-5 bar is against the fp64 oracle, tests above)
        assert rel_err(res[0][0][sl], res[1][0][sl]) < 3e-5, f"gradient tensor {t}"


@pytest.mark.parametrize("H,B", [(256, 128), (256, 300), (128, 1000), (256, 5000), (128, 77)])
def test_wide_family_vs_oracle_and_tile_family(core_mod, monkeypatch, H, B):
    """[256,256] (BASELINE.json C4) and the other wide nets train on the W family: every layer of forward and backward is
    a tcgen05 GEMM over the whole minibatch, activations travel as split-bf16 operand images (kernels_wide.cuh).
    Checked against the fp64 oracle (the 1e-5 bar, every gradient tensor separately) and against the generic fp32 tile
    (T) family.  B = 300 / 1000 / 5000 / 77 leave a ragged last tile; 5000 gives each split-K group several tiles."""
    rng = np.random.default_rng(1000 * H + B)
    p = rand_params(rng, H, H)
    o = ol.Oracle(h1=H, h2=H)
    obs = rng.standard_normal((B, 18)).astype(np.float32)
    _, v64, _, m64 = o.policy_step(p, obs, None, "f64")
    std = np.exp(p[o.offset(12):o.offset(13)].astype(np.float64))
    act = (m64 + std * rng.standard_normal((B, 18))).astype(np.float32)
    z = (act - m64) / std
    old_nlp = (0.5 * (z * z).sum(1) + 18 * 0.9189385175704956 + np.log(std).sum() + 0.1 * rng.standard_normal(B)).astype(np.float32)
    old_v = (v64 + 0.3 * rng.standard_normal(B)).astype(np.float32)
    ret = (v64 + 0.5 * rng.standard_normal(B)).astype(np.float32)
    adv = rng.standard_normal(B).astype(np.float32)
    res = []
    for env in (None, "PPO_DISABLE_WIDE"):
        if env:
            monkeypatch.setenv(env, "1")
        c = make_core(core_mod, p, hidden1=H, hidden2=H, n_envs=4, n_steps=8, nminibatches=4)
        assert ("wgemm_kernel" in c.kernel_family("train")) == (env is None)
        res.append(c.loss_grad(obs, act, adv, ret, old_nlp, old_v, 0.2))
        c.close()
        if env:
            monkeypatch.delenv(env)
    g64, l64 = o.loss_grad(p, obs, act, adv, ret, old_nlp, old_v, 0.2, "f64")
    (gw, lw), (gt, lt) = res
    assert np.allclose(lw, l64, rtol=TOL, atol=1e-7)
    assert rel_err(gw, g64) < TOL
    for t in range(13):
        sl = slice(o.offset(t), o.offset(t + 1))
        if t == 9:
            # vf/b is ONE number, (vf_coef / B) * sum_i c_i (v_i - R_i): a sum with cancellation whose fp32 evaluation is only
            # good to eps * sum|terms| whatever the kernel; bound the error by that instead of by the cancelled result
            atol = 2.0 ** -23 * 0.5 * float(np.mean(np.abs(v64 - ret)))
            assert abs(float(gw[sl][0]) - float(g64[sl][0])) < TOL * abs(float(g64[sl][0])) + atol
            continue
        assert rel_err(gw[sl], g64[sl]) < TOL, f"gradient tensor {t} vs the fp64 oracle"
        assert rel_err(gw[sl], gt[sl]) < 3e-5, f"gradient tensor {t} vs the T family"


def test_wide_family_training_equals_tile_family(core_mod, monkeypatch):
    """One whole update of a [256,256] net (device shuffle, advantage statistics, W-family minibatch steps with a ragged
    last tile, slabs -> reduce + clip + Adam, 8 dependent steps) against the same update on the T family.  The two
    evaluate the gradients in different fp32 orders (each within 1e-5 of the fp64 oracle, test above) and Adam's
    g / (sqrt(v) + eps) magnifies that for the parameters whose gradient is near eps, hence the looser bound here:
    this test is about the plumbing (slab layout, split-K groups, folded bias / loss columns)."""
    rng = np.random.default_rng(31)
    p = rand_params(rng, 256, 256)
    res = []
    for env in (None, "PPO_DISABLE_WIDE"):
        if env:
            monkeypatch.setenv(env, "1")
        c = make_core(core_mod, p, hidden1=256, hidden2=256, n_envs=100, n_steps=32, nminibatches=4, noptepochs=2, seed=21)
        c.shuffle_seed(7)
        c.synth_env_reset()
        losses = c.learn_update_synthetic(3e-4, 0.2)
        res.append(dict(losses=losses, params=c.get_tensor("params"), m=c.get_tensor("adam_m"), v=c.get_tensor("adam_v")))
        c.close()
        if env:
            monkeypatch.delenv(env)
    a, b = res
    moved = np.abs(a["params"] - p).max()
    assert moved > 1e-4
    assert np.abs(a["params"] - b["params"]).max() < 5e-3 * moved
    assert rel_err(a["losses"], b["losses"]) < 1e-4
    assert rel_err(a["m"], b["m"]) < 1e-4 and rel_err(a["v"], b["v"]) < 1e-4


# ------------------------------------------------------------------ minibatch step / whole update (a9-a11)
def _oracle_learner(kat, flat, n_envs, n_steps, nmb, epochs, env_kind, seed=77, shuffle_seed=42, h1=4, h2=5, lr=3.9e-4, cr=0.2):
    c = kat["consts"]
    lib = ol.load()
    d = ol.LearnerDesc(ol.Dims(18, 18, h1, h2), ol.HParams(c["ent_coef"], c["vf_coef"], c["clip_norm"], c["beta1"], c["beta2"], c["adam_eps"]),
                       n_envs, n_steps, nmb, epochs, 0.99, 0.95, lr, cr, seed, shuffle_seed, env_kind, 1)
    return lib, lib.oracle_learner_create(C.byref(d), np.ascontiguousarray(flat, np.float32))


def _buffers(lib, L, nb):
    names = ["obs", "returns", "dones", "actions", "values", "neglogpacs", "true_rewards", "unnormalized_rewards"]
    widths = [18, 1, 1, 18, 1, 1, 1, 1]
    return {n: np.ctypeslib.as_array(lib.oracle_learner_buffer(L, i), (nb, w)).copy() for i, (n, w) in enumerate(zip(names, widths))}


def test_train_minibatch_one_adam_step(core_mod, init_weights, kat):
    """Import the oracle's rollout, set the oracle's permutation, run ONE minibatch: gradient, losses and the
    Adam-updated parameters must match the fp64 oracle chain advnorm -> loss_grad -> clip_adam."""
    _, flat = init_weights
    n_envs, n_steps, nmb = 4, 64, 4
    nb = n_envs * n_steps
    lib, L = _oracle_learner(kat, flat, n_envs, n_steps, nmb, 1, 0)
    lib.oracle_learner_rollout(L)
    bufs = _buffers(lib, L, nb)
    cst = kat["consts"]
    c = make_core(core_mod, flat, n_envs=n_envs, n_steps=n_steps, nminibatches=nmb, noptepochs=1, ent_coef=cst["ent_coef"])
    for name, v in bufs.items():
        c.rollout_set(name, v)
        assert np.array_equal(c.rollout_get(name), v)  # layout round trip (flat row = env*n_steps + t)
    perm = ol.glibc_shuffle(42, nb, 1)[0]
    c.train_set_permutation(perm)
    src = np.zeros(nb, np.int32)
    lib.oracle_perm_to_gather(perm, nb, src)
    o = ol.Oracle()
    B = nb // nmb
    k = 2
    idx = src[k * B:(k + 1) * B]
    adv = o.advnorm(bufs["returns"][idx, 0], bufs["values"][idx, 0], "f64")
    g64, l64 = o.loss_grad(flat, bufs["obs"][idx], bufs["actions"][idx], adv, bufs["returns"][idx, 0], bufs["neglogpacs"][idx, 0],
                           bufs["values"][idx, 0], 0.2, "f64")
    losses, grads = c.train_minibatch(k, 3.9e-4, 0.2)
    # on-policy first minibatch: ratio == 1, so pg_loss = -mean(normalised adv) ~ 0 and approxkl ~ 0 are pure
    # cancellation residue of O(1) terms -> absolute tolerance at the fp32 resolution of those terms
    assert rel_err(grads, g64) < TOL and np.allclose(losses, l64, rtol=TOL, atol=2e-6)
    th0 = flat[:o.P].astype(np.float64)
    th1, m1, v1, _, b1p, b2p, gn = o.clip_adam(3.9e-4, th0, np.zeros(o.P), np.zeros(o.P), g64, float(np.float32(0.9)), float(np.float32(0.999)), "f64")
    got = c.get_tensor("params")
    assert rel_err(got[:o.P], th1) < 1e-6  # one Adam step: parameters to 1e-6 relative
    # the step itself is a difference of nearly equal fp32 numbers: |theta| ~ 0.8 stored in fp32 (ulp 6e-8) against
    # a step of ~lr = 3.9e-4 bounds its relative accuracy at ~2 ulp / step = 3e-4 for ANY fp32 implementation
    assert rel_err(got[:o.P] - flat[:o.P], th1 - th0) < 5e-4
    assert np.array_equal(got[o.P:], flat[o.P:])  # q head untouched
    assert rel_err(c.get_tensor("adam_m"), m1) < TOL and rel_err(c.get_tensor("adam_v"), v1) < TOL
    assert c.get_tensor("beta1_power")[0] == pytest.approx(b1p, rel=1e-6) and c.get_tensor("beta2_power")[0] == pytest.approx(b2p, rel=1e-6)
    lib.oracle_learner_destroy(L)
    c.close()


@pytest.mark.parametrize("h1,h2,n_envs,n_steps,nmb,epochs", [(4, 5, 2, 64, 4, 3), (4, 5, 1, 2048, 32, 2), (64, 64, 16, 32, 8, 2),
                                                              (64, 64, 1024, 16, 2, 2), (256, 256, 128, 32, 1, 1), (64, 64, 4096, 64, 32, 1)])
def test_train_update_vs_oracle_learner(core_mod, kat, init_weights, h1, h2, n_envs, n_steps, nmb, epochs):
    """Whole update (epochs x minibatches, compounded glibc shuffles, scatter semantics) vs the oracle's
    reference-structured learner on the SAME rollout.  The last two cases run the shapes of the benchmark: minibatches of 8192
    samples = 64 tiles x 2 towers on the persistent tcgen05 epoch kernel (C3's minibatch), and a 4096-sample minibatch step of
    the [256,256] layer-wise tcgen05 path (C4's net); the last one is C3 itself (262 144 transitions, 32 minibatches), one epoch."""
    rng = np.random.default_rng(3)
    flat = init_weights[1] if (h1, h2) == (4, 5) else rand_params(rng, h1, h2)
    nb = n_envs * n_steps
    lib, L = _oracle_learner(kat, flat, n_envs, n_steps, nmb, epochs, 0, h1=h1, h2=h2)
    lib.oracle_learner_rollout(L)
    bufs = _buffers(lib, L, nb)
    cst = kat["consts"]
    c = make_core(core_mod, flat, hidden1=h1, hidden2=h2, n_envs=n_envs, n_steps=n_steps, nminibatches=nmb, noptepochs=epochs,
                  ent_coef=cst["ent_coef"])
    for name, v in bufs.items():
        c.rollout_set(name, v)
    c.shuffle_seed(42)
    losses = c.train_update(3.9e-4, 0.2)
    want = np.zeros(5, np.float32)
    lib.oracle_learner_train(L, want)
    o = ol.Oracle(h1=h1, h2=h2)
    p1 = np.ctypeslib.as_array(lib.oracle_learner_params(L), (o.Pq,)).copy()
    got = c.get_tensor("params")
    # after epochs*nmb Adam steps of size ~lr the trajectories must still agree to a small fraction of the total movement
    moved = np.abs(p1[:o.P] - flat[:o.P]).max()
    print(f"whole update [{h1},{h2}] {n_envs}x{n_steps}: param diff / moved = {np.abs(got[:o.P] - p1[:o.P]).max() / moved:.3e}, "
          f"loss rel diff = {np.abs(losses - want).max() / np.abs(want).max():.3e}")
    # (measured: <= 1.6e-4 of the movement over up to 16 Adam steps — about one ulp of a weight per step — and 9.6e-4 after the 32
    # steps of a C3 epoch: early Adam steps are ~lr * sign(g), which amplifies rounding differences of small gradient entries)
    assert np.abs(got[:o.P] - p1[:o.P]).max() < max(5e-4, 4e-5 * epochs * nmb) * moved
    assert np.allclose(losses, want, rtol=2e-6, atol=1e-6), (losses, want)
    lib.oracle_learner_destroy(L)
    c.close()


# ------------------------------------------------------------------ rollout (a4) — device env and host env
@pytest.mark.parametrize("weights,n_steps,tol", [("init", 400, 2e-5), ("ckpt", 12, 2e-5), ("ckpt", 400, 5e-2)])
def test_rollout_synthetic_vs_oracle(core_mod, kat, init_weights, ckpt_weights, weights, n_steps, tol):
    """Whole device-resident rollout vs the oracle's Runner::run restatement.

    Every single step agrees to ~1e-7 (stepwise tests above); over a rollout the closed loop
    obs -> policy -> env -> VecNormalize re-amplifies rounding differences while the running variance is
    still small (gain ~ 0.1 * 1/sqrt(var) * policy gain > 1 during the first steps with the TRAINED policy).
    Hence: the untrained policy (output gain 0.01, practically open loop) must agree to 2e-5 over 400 steps
    that cross every env's episode boundary; the trained policy to 2e-5 over a short horizon and only
    loosely over 400 steps.  Done masks are bit-exact in every case."""
    _, flat = init_weights if weights == "init" else ckpt_weights
    n_envs = 32
    nb = n_envs * n_steps
    lib, L = _oracle_learner(kat, flat, n_envs, n_steps, 4, 1, 0, seed=2024)
    lib.oracle_learner_rollout(L)
    want = _buffers(lib, L, nb)
    c = make_core(core_mod, flat, n_envs=n_envs, n_steps=n_steps, nminibatches=4, noptepochs=1, seed=2024)
    c.synth_env_reset()
    c.rollout_synthetic()
    got = {n: c.rollout_get(n) for n in want}
    assert np.array_equal(got["dones"], want["dones"])  # done masks: bit-exact
    if n_steps >= 334:
        assert want["dones"].sum() >= n_envs  # every env ended at least one episode
    for name in ("obs", "actions", "values", "neglogpacs", "true_rewards", "unnormalized_rewards", "returns"):
        assert rel_err(got[name], want[name]) < tol, name
    om, ov = np.zeros(18, np.float32), np.zeros(18, np.float32)
    oc, rm, rv, rc = C.c_double(), C.c_float(), C.c_float(), C.c_double()
    lib.oracle_learner_get_norm(L, om, ov, C.byref(oc), C.byref(rm), C.byref(rv), C.byref(rc))
    st = c.vecnorm_stats()
    assert st["obs_count"] == oc.value and st["ret_count"] == rc.value
    assert rel_err(st["obs_mean"], om) < max(tol, 1e-4) and rel_err(st["obs_var"], ov) < max(tol, 1e-4)
    assert st["ret_var"][0] == pytest.approx(rv.value, rel=max(tol, 1e-4))
    lib.oracle_learner_destroy(L)
    c.close()


@pytest.mark.parametrize("h1,h2,n_envs,n_steps", [(4, 5, 77, 40), (64, 64, 4096, 16), (64, 64, 33, 700), (4, 5, 40000, 6)])
def test_rollout_persistent_equals_stepwise(core_mod, monkeypatch, h1, h2, n_envs, n_steps):
    """The one-launch cooperative rollout kernel (R family) against the per-step kernels (policy step, env step,
    VecNormalize moments/apply, GAE) over two consecutive rollouts: identical Philox streams and fp32 operation order;
    only the fp64 summation order of the VecNormalize batch moments differs (sums are rounded to fp32 afterwards).
    n_envs = 77 / 33 leave a ragged last tile, 40000 envs puts several tiles on one CTA, 700 steps cross every env's
    episode boundary twice."""
    rng = np.random.default_rng(5)
    p = rand_params(rng, h1, h2)
    res = []
    for env in (None, "PPO_DISABLE_PERSISTENT"):
        if env:
            monkeypatch.setenv(env, "1")
        c = make_core(core_mod, p, hidden1=h1, hidden2=h2, n_envs=n_envs, n_steps=n_steps, nminibatches=1, noptepochs=1, seed=11)
        assert ("persistent" in c.kernel_family("rollout")) == (env is None)
        c.synth_env_reset()
        c.rollout_synthetic()
        first = {n: c.rollout_get(n) for n in ("dones", "obs")}
        c.rollout_synthetic()
        out = {n: c.rollout_get(n) for n in ("obs", "actions", "values", "neglogpacs", "dones", "true_rewards", "unnormalized_rewards", "returns")}
        out["first_dones"], out["first_obs"] = first["dones"], first["obs"]
        out["stats"] = c.vecnorm_stats()
        res.append(out)
        c.close()
        if env:
            monkeypatch.delenv(env)
    a, b = res
    assert np.array_equal(a["dones"], b["dones"]) and np.array_equal(a["first_dones"], b["first_dones"])
    assert np.array_equal(a["unnormalized_rewards"], b["unnormalized_rewards"]) or rel_err(a["unnormalized_rewards"], b["unnormalized_rewards"]) < 1e-5
    for name in ("first_obs", "obs", "actions", "values", "neglogpacs", "true_rewards", "returns"):
        assert rel_err(a[name], b[name]) < 2e-5, name
    assert a["stats"]["obs_count"] == b["stats"]["obs_count"] and a["stats"]["ret_count"] == b["stats"]["ret_count"]
    assert rel_err(a["stats"]["obs_mean"], b["stats"]["obs_mean"]) < 1e-5 and rel_err(a["stats"]["obs_var"], b["stats"]["obs_var"]) < 1e-5
    assert rel_err(a["stats"]["ret_var"], b["stats"]["ret_var"]) < 1e-5


@pytest.mark.parametrize("h,n_envs,n_steps,nmb", [(64, 256, 32, 2), (64, 4096, 16, 8), (64, 1000, 8, 2), (4, 1, 2048, 32), (4, 8, 64, 4), (4, 3, 50, 1)])
def test_epoch_persistent_equals_stepwise(core_mod, monkeypatch, h, n_envs, n_steps, nmb):
    """[64,64]: the persistent epoch kernel (all minibatches of an epoch in one cooperative launch: tcgen05 tiles, column
    reduce over distributed shared memory + LL hand-overs, global norm, Adam inside thread-block clusters) against the
    launch-per-minibatch path over two updates.  Same arithmetic; the order of the cross-CTA gradient sums and of the
    sum-of-squares partials differs, so single parameters may differ by an ulp of their own magnitude.
    h = 4: the reference's [4,5] net with minibatches that one CTA handles (C1): the single-CTA epoch kernel of the S family
    (combine, clip and Adam in shared memory) against train_small_kernel + the cooperative reduce / Adam kernel per minibatch."""
    rng = np.random.default_rng(8)
    h1, h2 = (64, 64) if h == 64 else (4, 5)
    p = rand_params(rng, h1, h2)
    res = []
    for env in (None, "PPO_DISABLE_PERSISTENT"):
        if env:
            monkeypatch.setenv(env, "1")
        c = make_core(core_mod, p, hidden1=h1, hidden2=h2, n_envs=n_envs, n_steps=n_steps, nminibatches=nmb, noptepochs=3, seed=21)
        assert ("persistent" in c.kernel_family("train")) == (env is None)
        c.shuffle_seed(7)
        c.synth_env_reset()
        losses = [c.learn_update_synthetic(3e-4, 0.2) for _ in range(2)]
        res.append(dict(losses=np.stack(losses), params=c.get_tensor("params"), m=c.get_tensor("adam_m"), v=c.get_tensor("adam_v"),
                        b1=c.get_tensor("beta1_power"), b2=c.get_tensor("beta2_power")))
        c.close()
        if env:
            monkeypatch.delenv(env)
    a, b = res
    moved = np.abs(a["params"] - p).max()
    ulp = np.finfo(np.float32).eps * np.abs(p).max()
    assert np.abs(a["params"] - b["params"]).max() < 1e-5 * moved + 2 * ulp
    assert rel_err(a["losses"], b["losses"]) < 1e-5
    assert rel_err(a["m"], b["m"]) < 1e-5 and rel_err(a["v"], b["v"]) < 1e-5
    assert np.array_equal(a["b1"], b["b1"]) and np.array_equal(a["b2"], b["b2"])


@pytest.mark.parametrize("n_envs,n_steps,epochs,seed", [(1, 2, 1, 1), (3, 7, 4, 42), (16, 64, 3, 7), (257, 33, 2, 123456), (4096, 64, 10, 42),
                                                         (16384, 64, 4, 9)])  # the last one: 4 x 1 M entries (the global batch of a multi-GPU run)
def test_device_random_shuffle_bit_exact(core_mod, n_envs, n_steps, epochs, seed):
    """The permutations built on the device (parallel glibc rand() stream by polynomial jump-ahead + parallel resolution
    of the swap chain, kernels_shuffle.cuh) must equal std::srand(seed) + std::random_shuffle bit for bit: every epoch of
    an update (compounded, ppo2.hpp:274-288) and the continuation of the rand() stream into the next update."""
    n = n_envs * n_steps
    nmb = next(m for m in (4, 3, 1) if n % m == 0 and n // m >= 2) if n >= 64 else 1
    c = make_core(core_mod, None, n_envs=n_envs, n_steps=n_steps, nminibatches=nmb, noptepochs=epochs, seed=3)
    c.init_orthogonal(1)
    c.shuffle_seed(seed)
    c.synth_env_reset()
    c.rollout_synthetic()
    c.train_update(1e-4, 0.2)
    want = core_mod.host_random_shuffle(seed, n, epochs)
    for e in range(epochs):
        assert np.array_equal(c.train_get_permutation(e), want[e]), f"update 0 epoch {e}"
    # second update: identity again, the rand() stream goes on where the first update stopped
    c.train_update(1e-4, 0.2)
    if n <= 20000:
        stream = core_mod.host_rand(seed, 2 * epochs * (n - 1))[epochs * (n - 1):]
        a = np.arange(n, dtype=np.int32)
        k = 0
        for e in range(epochs):
            for i in range(1, n):
                j = int(stream[k]) % (i + 1)
                k += 1
                a[i], a[j] = a[j], a[i]
            assert np.array_equal(c.train_get_permutation(e), a), f"update 1 epoch {e}"
    c.close()


def test_shuffle_prefetch_equals_inline(core_mod, monkeypatch):
    """From the second update on, the permutations are built on a second stream while the rollout runs (they depend on the
    rand() stream only).  Same kernels, same generator state: training and permutations are bit-identical to the inline
    order, and a re-seed after the prefetch has started discards it."""
    rng = np.random.default_rng(4)
    p = rand_params(rng, 64, 64)
    n_envs, n_steps, E = 256, 16, 3
    res = []
    for env in (None, "PPO_DISABLE_SHUFFLE_PREFETCH"):
        if env:
            monkeypatch.setenv(env, "1")
        c = make_core(core_mod, p, hidden1=64, hidden2=64, n_envs=n_envs, n_steps=n_steps, nminibatches=4, noptepochs=E, seed=21)
        c.shuffle_seed(7)
        c.synth_env_reset()
        losses = [c.learn_update_synthetic(3e-4, 0.2) for _ in range(3)]
        perms = [c.train_get_permutation(e) for e in range(E)]
        # a rollout starts (prefetch of the 4th update's permutations), then the generator is re-seeded
        c.rollout_synthetic()
        c.shuffle_seed(77)
        c.train_update(3e-4, 0.2)
        reseeded = [c.train_get_permutation(e) for e in range(E)]
        res.append(dict(losses=np.stack(losses), params=c.get_tensor("params"), perms=perms, reseeded=reseeded))
        c.close()
        if env:
            monkeypatch.delenv(env)
    a, b = res
    assert np.array_equal(a["params"], b["params"]) and np.array_equal(a["losses"], b["losses"])
    n = n_envs * n_steps
    want = core_mod.host_random_shuffle(7, n, 3 * E)  # three updates: identity again at each, the stream goes on
    stream3 = core_mod.host_random_shuffle(77, n, E)
    for e in range(E):
        assert np.array_equal(a["perms"][e], b["perms"][e])
        assert np.array_equal(a["reseeded"][e], stream3[e]) and np.array_equal(b["reseeded"][e], stream3[e])
    del want


def test_device_shuffle_equals_host_shuffle_training(core_mod, monkeypatch):
    """Same training run with the permutations from the device kernels and from the host restatement."""
    rng = np.random.default_rng(4)
    p = rand_params(rng, 64, 64)
    res = []
    for env in (None, "PPO_DISABLE_GPU_SHUFFLE"):
        if env:
            monkeypatch.setenv(env, "1")
        c = make_core(core_mod, p, hidden1=64, hidden2=64, n_envs=512, n_steps=16, nminibatches=4, noptepochs=3, seed=21)
        c.shuffle_seed(99)
        c.synth_env_reset()
        losses = [c.learn_update_synthetic(3e-4, 0.2) for _ in range(3)]
        res.append((np.stack(losses), c.get_tensor("params")))
        c.close()
        if env:
            monkeypatch.delenv(env)
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])


def test_runner_host_env_protocol_equals_device_env(core_mod, ckpt_weights):
    """Runner::run through host buffers (act -> env.step on the host -> observe) with the oracle's synthetic env
    as the host env must reproduce the all-device rollout of the same seed."""
    _, flat = ckpt_weights
    n_envs, n_steps, seed = 16, 48, 99
    lib = ol.load()
    env = lib.oracle_synth_env_create(n_envs, 18, seed ^ 0x1234, 0)
    obs, rew, done = np.zeros((n_envs, 18), np.float32), np.zeros(n_envs, np.float32), np.zeros(n_envs, np.float32)
    c = make_core(core_mod, flat, n_envs=n_envs, n_steps=n_steps, nminibatches=4, noptepochs=1, seed=seed)
    lib.oracle_synth_env_reset(env, obs)
    c.runner_reset(obs)
    for t in range(n_steps):
        act = c.runner_act(t)
        lib.oracle_synth_env_step(env, act, obs, rew, done)
        c.runner_observe(t, obs, rew, done)
    c.runner_finish()
    host = {n: c.rollout_get(n) for n in ("obs", "actions", "values", "returns", "dones", "true_rewards")}
    lib.oracle_synth_env_destroy(env)
    c2 = make_core(core_mod, flat, n_envs=n_envs, n_steps=n_steps, nminibatches=4, noptepochs=1, seed=seed)
    c2.synth_env_reset()
    c2.rollout_synthetic()
    for n, v in host.items():
        assert rel_err(v, c2.rollout_get(n)) < 1e-5, n
    c.close()
    c2.close()


@pytest.mark.parametrize("pinned", [False, True])
def test_runner_rollout_replay_equals_stepwise_protocol(core_mod, ckpt_weights, pinned):
    """ppo_runner_rollout_replay (the act/observe loop of Runner::run as one C call, host env = a recorded trajectory)
    against the same trajectory fed step by step; pageable buffers go through the core's pinned double buffer, pinned
    buffers are DMA'd in place.  Same kernels in the same order: bit-identical rollout buffers and actions."""
    _, flat = ckpt_weights
    n_envs, n_steps, seed = 33, 20, 5
    rng = np.random.default_rng(3)
    shape = [(n_steps, n_envs, 18), (n_steps, n_envs), (n_steps, n_envs), (n_steps, n_envs, 18)]
    if pinned:
        import torch
        raw, rew, done, acts = (torch.empty(s_, dtype=torch.float32, pin_memory=True).numpy() for s_ in shape)
    else:
        raw, rew, done, acts = (np.empty(s_, np.float32) for s_ in shape)
    raw[:] = 2.0 + rng.standard_normal(shape[0])
    rew[:] = rng.standard_normal(shape[1])
    done[:] = rng.random(shape[2]) < 0.1
    a = make_core(core_mod, flat, n_envs=n_envs, n_steps=n_steps, nminibatches=4, noptepochs=1, seed=seed)
    b = make_core(core_mod, flat, n_envs=n_envs, n_steps=n_steps, nminibatches=4, noptepochs=1, seed=seed)
    a.runner_reset(raw[0])
    b.runner_reset(raw[0])
    want_acts = []
    for t in range(n_steps):
        want_acts.append(a.runner_act(t).copy())
        a.runner_observe(t, raw[t], rew[t], done[t])
    a.runner_finish()
    b.runner_rollout_replay(raw, rew, done, acts)
    assert np.array_equal(acts, np.stack(want_acts))
    for n in ("obs", "actions", "values", "neglogpacs", "returns", "dones", "true_rewards", "unnormalized_rewards"):
        assert np.array_equal(a.rollout_get(n), b.rollout_get(n)), n
    a.close()
    b.close()


@pytest.mark.parametrize("n_envs", [1, 40, 700])
def test_host_env_one_kernel_rollout_equals_stepwise_and_aborts_cleanly(core_mod, monkeypatch, ckpt_weights, n_envs):
    """ppo_runner_rollout_host with a live host env (Python callback): the whole rollout is ONE persistent kernel that trades
    actions / observations with the host — through mapped pinned memory up to 512 envs, beyond that the env's answer goes
    through the copy engine into a device staging buffer with a 4-byte flag copy behind it (700 envs here).  It must equal
    the per-step protocol over consecutive rollouts (bit for bit up to 512 envs; with more CTAs the fp64 moment partials are
    cut differently: 1e-5), and an env that aborts must release the kernel and leave the core usable."""
    _, flat = ckpt_weights
    n_steps, seed = 24, 8
    lib = ol.load()

    def run(disable):
        if disable:
            monkeypatch.setenv("PPO_DISABLE_HOST_PERSISTENT", "1")
        else:
            monkeypatch.delenv("PPO_DISABLE_HOST_PERSISTENT", raising=False)
        env = lib.oracle_synth_env_create(n_envs, 18, seed ^ 0x1234, 0)
        obs, rew, done = np.zeros((n_envs, 18), np.float32), np.zeros(n_envs, np.float32), np.zeros(n_envs, np.float32)
        c = make_core(core_mod, flat, n_envs=n_envs, n_steps=n_steps, nminibatches=4, noptepochs=1, seed=seed)
        lib.oracle_synth_env_reset(env, obs)
        c.runner_reset(obs)

        def step(t, act):
            lib.oracle_synth_env_step(env, np.ascontiguousarray(act), obs, rew, done)
            return obs, rew, done

        out = []
        for _ in range(2):
            c.runner_rollout_host(step)
            out.append({n: c.rollout_get(n) for n in ("obs", "actions", "values", "neglogpacs", "returns", "dones", "true_rewards", "unnormalized_rewards")})
        st = c.vecnorm_stats()
        lib.oracle_synth_env_destroy(env)
        return c, out, st

    c1, one_kernel, st1 = run(False)
    c2, stepwise, st2 = run(True)
    for a, b in zip(one_kernel, stepwise):
        for n in a:
            if n_envs <= 512 or n == "dones":
                assert np.array_equal(a[n], b[n]), n
            else:
                assert rel_err(a[n], b[n]) < 1e-5, n
    assert st1["obs_count"] == st2["obs_count"]
    if n_envs <= 512:
        assert np.array_equal(st1["obs_mean"], st2["obs_mean"])
    else:
        assert rel_err(st1["obs_mean"], st2["obs_mean"]) < 1e-5
    # an env that gives up at step 5: the call fails, nothing hangs, the next rollout works
    monkeypatch.delenv("PPO_DISABLE_HOST_PERSISTENT", raising=False)
    o, r, d = np.zeros((n_envs, 18), np.float32), np.zeros(n_envs, np.float32), np.zeros(n_envs, np.float32)
    with pytest.raises(core_mod.PPOError):
        c1.runner_rollout_host(lambda t, act: None if t == 5 else (o, r, d))
    c1.runner_reset(o)
    c1.runner_rollout_host(lambda t, act: (o, r, d))
    assert np.all(np.isfinite(c1.rollout_get("returns")))
    c1.close()
    c2.close()


def test_rollout_mock_env_c1(core_mod, init_weights, kat):
    """Config C1: EnvMock (constant obs/reward 1, done every 300th step) driven through the host protocol.
    Rewards, dones and the done-lag of Runner::run are exact; the normalised observation is cancellation noise
    (SURVEY §7 'C1 is numerically degenerate'), so only its clip bound is checked."""
    _, flat = init_weights
    n_steps = 640
    c = make_core(core_mod, flat, n_envs=1, n_steps=n_steps, nminibatches=32, noptepochs=1)
    lib, L = _oracle_learner(kat, flat, 1, n_steps, 32, 1, 1)
    lib.oracle_learner_rollout(L)
    want = _buffers(lib, L, n_steps)
    c.runner_reset(np.ones((1, 18), np.float32))
    for t in range(n_steps):
        c.runner_act(t)
        c.runner_observe(t, np.ones((1, 18), np.float32), np.ones(1, np.float32), np.full(1, 1.0 if (t + 1) % 300 == 0 else 0.0, np.float32))
    c.runner_finish()
    assert np.array_equal(c.rollout_get("dones"), want["dones"])  # dones[t] = done of step t-1
    assert np.flatnonzero(want["dones"][:, 0]).tolist() == [300, 600]
    assert np.array_equal(c.rollout_get("unnormalized_rewards"), want["unnormalized_rewards"])
    assert rel_err(c.rollout_get("true_rewards"), want["true_rewards"]) < 1e-4
    assert np.abs(c.rollout_get("obs")).max() <= 10.0
    lib.oracle_learner_destroy(L)
    c.close()


# ------------------------------------------------------------------ weights I/O (graph file, checkpoint)
def test_graph_and_checkpoint_io(core_mod, init_weights, ckpt_weights, tmp_path):
    from ppo_cpp_b200.meta_graph import CKPT_ORDER, write_meta_txt
    tensors, flat = init_weights
    path = str(tmp_path / "g.meta.txt")
    write_meta_txt(path, tensors, ent_coef=0.0007160293171182275)
    c = make_core(core_mod, None, n_envs=4, n_steps=8, nminibatches=4)
    c.load_meta_txt(path)
    assert np.array_equal(c.get_tensor("params"), flat)
    assert np.array_equal(c.get_tensor("model/pi_fc1/w"), tensors["model/pi_fc1/w"].ravel())
    ck, ckflat = ckpt_weights
    prefix = str(tmp_path / "run.pkl.71")
    np.concatenate([ck[n].ravel() for n in CKPT_ORDER]).astype("<f4").tofile(prefix + ".data-00000-of-00001")
    c.load_checkpoint_data(prefix)
    assert np.array_equal(c.get_tensor("params"), ckflat)
    c.save_checkpoint_data(str(tmp_path / "again"))
    assert open(prefix + ".data-00000-of-00001", "rb").read() == open(str(tmp_path / "again") + ".data-00000-of-00001", "rb").read()
    # ... and the bundle's .index is the file TensorFlow wrote for this checkpoint (the reference's shipped *.pkl.71.index)
    assert open(str(tmp_path / "again") + ".index", "rb").read() == open(os.path.join(os.path.dirname(__file__), "golden", "ckpt_71.index"), "rb").read()
    with pytest.raises(core_mod.PPOError):
        c.load_meta_txt(str(tmp_path / "missing.meta.txt"))
    c.close()
    wide = core_mod.PPOCore(hidden1=64, hidden2=64, n_envs=4, n_steps=8, nminibatches=4)
    with pytest.raises(core_mod.PPOError, match="MLP"):
        wide.load_meta_txt(path)  # shape mismatch is an error, not a silent reshape
    wide.init_orthogonal(3)
    w = wide.get_tensor("model/pi_fc1/w").reshape(64, 64)
    assert np.allclose(w.T @ w, 2 * np.eye(64), atol=1e-4)  # orthogonal, gain sqrt(2)
    wide.close()


# ------------------------------------------------------------------ multi-GPU
def test_two_gpu_sharded_run_equals_single_gpu():
    """Runs tests/mgpu_check.py under torchrun when the box has >= 2 GPUs (gpurun --gpus 2)."""
    import subprocess
    import sys

    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(root, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
