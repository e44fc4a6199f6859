"""CPU tests: the oracle against the golden fixtures (tests/golden/, made by make_golden.py from the
reference's graph / checkpoint, real glibc + libstdc++, Random123 KATs and PyTorch fp64 autograd)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from conftest import rel_err


def test_philox_known_answers(kat):
    lib = ol.load()
    for v in kat["philox4x32_10"]:
        out = np.zeros(4, np.uint32)
        lib.oracle_philox4x32_10(np.array(v["ctr"], np.uint32), np.array(v["key"], np.uint32), out)
        assert out.tolist() == v["out"]


def test_unit_float_and_normals():
    lib = ol.load()
    assert lib.oracle_u32_to_unit_float(0) == 0.0
    assert lib.oracle_u32_to_unit_float(0xFFFFFFFF) == pytest.approx(1.0 - 2.0 ** -23)
    eps = np.zeros((4096, 18), np.float32)
    for e in range(4096):
        lib.oracle_normal_eps(99, e, 7, 18, eps[e])
    assert abs(eps.mean()) < 0.02 and abs(eps.std() - 1.0) < 0.02
    e2 = np.zeros(18, np.float32)
    lib.oracle_normal_eps(99, 5, 7, 18, e2)
    assert np.array_equal(e2, eps[5])  # counter-based: reproducible per (seed, env, step)


def test_glibc_rand_bit_exact(kat):
    lib = ol.load()
    for seed, vals in kat["glibc"]["rand"].items():
        st = ol.GlibcRand()
        lib.oracle_srand(C.byref(st), int(seed))
        assert [lib.oracle_rand(C.byref(st)) for _ in range(len(vals))] == vals


def test_random_shuffle_bit_exact(kat):
    for key, perms in kat["glibc"]["shuffle"].items():
        seed, n, epochs = (int(x) for x in key.split("_"))
        got = ol.glibc_shuffle(seed, n, epochs)
        for g, want in zip(got, perms):
            assert g.tolist() == want
    assert kat["glibc"]["shuffle"]["42_8_1"][0] == [2, 6, 1, 0, 5, 7, 4, 3]


def test_live_libc_matches_restatement():
    libc = C.CDLL("libc.so.6")
    libc.rand.restype = C.c_int
    lib = ol.load()
    st = ol.GlibcRand()
    for seed in (3, 2024):
        libc.srand(seed)
        lib.oracle_srand(C.byref(st), seed)
        assert [libc.rand() for _ in range(1000)] == [lib.oracle_rand(C.byref(st)) for _ in range(1000)]


def test_perm_to_gather_scatter_semantics():
    lib = ol.load()
    perm = ol.glibc_shuffle(5, 37)[0]
    src = np.zeros(37, np.int32)
    lib.oracle_perm_to_gather(perm, 37, src)
    data = np.arange(37) * 10
    out = np.zeros(37, int)
    out[perm] = data  # Eigen perm * v: out[perm[i]] = in[i]
    assert np.array_equal(out, data[src])


@pytest.mark.parametrize("wname", ["init", "ckpt"])
def test_forward_known_answers(wname, forward_kat, init_weights, ckpt_weights, kat):
    _, flat = init_weights if wname == "init" else ckpt_weights
    k = forward_kat[wname]
    o = ol.Oracle()
    for prec, tol in (("f64", 1e-12), ("f32", 2e-6)):
        action, value, nlp, mean = o.policy_step(flat, k["obs"], k["eps"], prec)
        assert rel_err(mean, k["mean"]) < tol
        assert rel_err(value, k["value"]) < tol
        assert rel_err(action, k["action"]) < tol
        assert rel_err(nlp, k["neglogp"]) < tol
    if wname == "init":  # analytic: zero biases => obs 0 gives mean 0, value 0, neglogp = 0.5|a|^2 + 16.540893
        a, v, n, m = o.policy_step(flat, np.zeros((1, 18), np.float32), k["eps"][:1], "f64")
        assert np.all(m == 0) and v[0] == 0
        assert n[0] == pytest.approx(0.5 * float((a ** 2).sum()) + 18 * 0.9189385175704956, rel=1e-12)
    else:
        s = kat["survey_8c"]
        _, v, n, m = o.policy_step(flat, k["obs"][:2], None, "f64")
        assert np.allclose(m[0, :4], [-0.259430, -0.636848, -0.479933, 0.574643], atol=1e-6)
        assert v[0] == pytest.approx(2.587764, abs=1e-6) and v[1] == pytest.approx(2.458381, abs=1e-6)
        assert n[0] == pytest.approx(0.611873, abs=1e-5) == pytest.approx(s["ckpt_neglogp_at_mean"], abs=1e-5)


@pytest.mark.parametrize("case", ["init_4_5", "ckpt_4_5", "rand_64_64", "rand_8_8"])
def test_loss_grad_vs_torch_autograd(case, loss_kat, kat):
    k = loss_kat[case]
    h1, h2 = (int(x) for x in k["hidden"])
    c = kat["consts"]
    o = ol.Oracle(h1=h1, h2=h2, ent_coef=c["ent_coef"], vf_coef=c["vf_coef"])
    g64, l64 = o.loss_grad(k["params"], k["obs"], k["act"], k["adv"], k["ret"], k["old_nlp"], k["old_v"], float(k["cliprange"]), "f64")
    assert rel_err(g64, k["grads"]) < 1e-12
    assert np.allclose(l64, k["losses"], rtol=1e-12, atol=1e-14)
    g32, l32 = o.loss_grad(k["params"], k["obs"], k["act"], k["adv"], k["ret"], k["old_nlp"], k["old_v"], float(k["cliprange"]), "f32")
    assert rel_err(g32, k["grads"]) < 1e-5
    assert np.allclose(l32, k["losses"], rtol=1e-5, atol=1e-7)
    # per-tensor check too (small tensors must not hide behind big ones)
    from ppo_cpp_b200.meta_graph import param_layout, TENSOR_ORDER
    lay = param_layout(18, 18, h1, h2)
    for name in TENSOR_ORDER[:13]:
        off, shp = lay[name]
        n = int(np.prod(shp))
        assert rel_err(g64[off:off + n], k["grads"][off:off + n]) < 1e-10, name


@pytest.mark.parametrize("case", ["init_4_5", "rand_64_64"])
def test_clip_adam_vs_numpy(case, loss_kat, kat):
    k = loss_kat[case]
    h1, h2 = (int(x) for x in k["hidden"])
    o = ol.Oracle(h1=h1, h2=h2)
    theta = k["params"][:o.P]
    p, m, v, g, b1p, b2p, gn = o.clip_adam(3.9e-4, theta, k["adam_m0"], k["adam_v0"], k["grads"], 0.9 ** 3, 0.999 ** 3, "f64")
    assert gn == pytest.approx(float(k["gnorm"]), rel=1e-12)
    assert rel_err(p, k["adam_theta1"]) < 1e-12 and rel_err(m, k["adam_m1"]) < 1e-12 and rel_err(v, k["adam_v1"]) < 1e-12
    assert b1p == pytest.approx(0.9 ** 3 * float(np.float32(0.9)), rel=1e-12)
    p32, m32, v32, *_ = o.clip_adam(3.9e-4, theta, k["adam_m0"], k["adam_v0"], k["grads"], 0.9 ** 3, 0.999 ** 3, "f32")
    assert rel_err(p32 - theta, k["adam_theta1"] - theta.astype(np.float64)) < 1e-4  # the step itself
    assert rel_err(m32, k["adam_m1"]) < 1e-6 and rel_err(v32, k["adam_v1"]) < 1e-6


def np_gae(rew, val, done, last_v, last_d, gamma, lam):
    T, N = rew.shape
    adv = np.zeros((T, N))
    last = np.zeros(N)
    for t in reversed(range(T)):
        nt = 1.0 - (last_d if t == T - 1 else done[t + 1])
        nv = last_v if t == T - 1 else val[t + 1]
        delta = rew[t] + gamma * nv * nt - val[t]
        last = delta + gamma * lam * nt * last
        adv[t] = last
    return adv, adv + val


def test_gae_vs_numpy_and_closed_form():
    rng = np.random.default_rng(0)
    T, N = 97, 13
    rew, val = rng.standard_normal((T, N)).astype(np.float32), rng.standard_normal((T, N)).astype(np.float32)
    done = (rng.random((T, N)) < 0.05).astype(np.float32)
    lv, ld = rng.standard_normal(N).astype(np.float32), (rng.random(N) < 0.5).astype(np.float32)
    o = ol.Oracle()
    g, lam = float(np.float32(0.99)), float(np.float32(0.95))
    a64, r64 = o.gae(rew, val, done, lv, ld, g, lam, "f64")
    wa, wr = np_gae(rew.astype(np.float64), val.astype(np.float64), done.astype(np.float64), lv.astype(np.float64), ld.astype(np.float64), g, lam)
    assert np.allclose(a64, wa, rtol=1e-13, atol=1e-13) and np.allclose(r64, wr, rtol=1e-13, atol=1e-13)
    a32, r32 = o.gae(rew, val, done, lv, ld, g, lam, "f32")
    assert rel_err(a32, wa) < 1e-5 and rel_err(r32, wr) < 1e-5
    # EnvMock-like closed form: constant reward 1, value 0, no dones: adv_t = sum_k (g*lam)^k
    T = 50
    a, _ = o.gae(np.ones((T, 1), np.float32), np.zeros((T, 1), np.float32), np.zeros((T, 1), np.float32),
                 np.zeros(1, np.float32), np.zeros(1, np.float32), g, lam, "f64")
    q = g * lam
    assert np.allclose(a[:, 0], [(1 - q ** (T - t)) / (1 - q) for t in range(T)], rtol=1e-12)


def test_advnorm():
    rng = np.random.default_rng(1)
    r, v = rng.standard_normal(512).astype(np.float32), rng.standard_normal(512).astype(np.float32)
    adv = r.astype(np.float64) - v.astype(np.float64)
    want = (adv - adv.mean()) / (adv.std() + 1e-8)
    o = ol.Oracle()
    assert np.allclose(o.advnorm(r, v, "f64"), want, rtol=1e-12, atol=1e-12)
    assert rel_err(o.advnorm(r, v, "f32"), want) < 1e-5


def test_running_statistics_chan_merge_equals_batch_stats():
    rng = np.random.default_rng(2)
    lib = ol.load()
    D = 18
    data = (rng.standard_normal((40, 16, D)) * 3 + 1).astype(np.float32)
    mean, var, cnt = np.zeros(D), np.ones(D), C.c_double(1e-6)
    vn = lib.oracle_vecnorm_create(16, D, 1)
    for b in data:
        lib.oracle_rstats_update_f64(mean, var, C.byref(cnt), D, np.ascontiguousarray(b, np.float64), 16)
        lib.oracle_rstats_update_f32(C.byref(vn.contents.obs_rms), np.ascontiguousarray(b), 16)
    allrows = data.reshape(-1, D).astype(np.float64)
    assert np.allclose(mean, allrows.mean(0), atol=1e-6) and np.allclose(var, allrows.var(0), rtol=1e-5)
    m32 = np.ctypeslib.as_array(vn.contents.obs_rms.mean, (D,))
    v32 = np.ctypeslib.as_array(vn.contents.obs_rms.var, (D,))
    assert rel_err(m32, mean) < 1e-5 and rel_err(v32, var) < 1e-5
    assert vn.contents.obs_rms.count == pytest.approx(640 + 1e-6)
    lib.oracle_vecnorm_destroy(vn)


def test_vecnorm_step_semantics():
    """env_normalize.hpp:64-116 — count bookkeeping, clip, ret reset on done."""
    lib = ol.load()
    N, D = 8, 18
    rng = np.random.default_rng(3)
    vn = lib.oracle_vecnorm_create(N, D, 1)
    obs_out, rew_out = np.zeros((N, D), np.float32), np.zeros(N, np.float32)
    lib.oracle_vecnorm_reset_f32(vn, (rng.standard_normal((N, D)) * 50).astype(np.float32), obs_out)
    assert np.abs(obs_out).max() <= 10.0
    K = 5
    for k in range(K):
        done = np.zeros(N, np.float32)
        done[k % N] = 1.0
        lib.oracle_vecnorm_step_f32(vn, rng.standard_normal((N, D)).astype(np.float32),
                                    (rng.standard_normal(N) * 100).astype(np.float32), done, obs_out, rew_out)
        ret = np.ctypeslib.as_array(vn.contents.ret, (N,))
        assert ret[k % N] == 0.0 and np.abs(rew_out).max() <= 10.0
    assert vn.contents.obs_rms.count == pytest.approx(1e-6 + N * (K + 1))
    assert vn.contents.ret_rms.count == pytest.approx(1e-6 + N * K)
    lib.oracle_vecnorm_destroy(vn)


def test_synth_env_shapes_and_episode_boundaries():
    lib = ol.load()
    N, D = 700, 18
    env = lib.oracle_synth_env_create(N, D, 1234, 0)
    obs, rew, done = np.zeros((N, D), np.float32), np.zeros(N, np.float32), np.zeros(N, np.float32)
    lib.oracle_synth_env_reset(env, obs)
    assert np.abs(obs).max() <= 0.1
    ndone = np.zeros(N)
    for t in range(334):
        prev0 = obs[:, 0].copy()
        lib.oracle_synth_env_step(env, np.zeros((N, D), np.float32), obs, rew, done)
        ndone += done
        live = done == 0
        assert np.allclose(rew[live], obs[live, 0] - prev0[live], atol=1e-7)
    assert np.all(ndone == 1)  # every env finishes exactly one 334-step episode per 334 steps (phase = id % 334)
    lib.oracle_synth_env_destroy(env)


def test_learner_mock_env_matches_stepwise_pieces(init_weights, kat):
    """The whole-learner port (reference structure) == composing the individual oracle functions."""
    _, flat = init_weights
    c = kat["consts"]
    lib = ol.load()
    d = ol.LearnerDesc(ol.Dims(18, 18, 4, 5), ol.HParams(c["ent_coef"], c["vf_coef"], c["clip_norm"], c["beta1"], c["beta2"], c["adam_eps"]),
                       2, 64, 4, 2, 0.99, 0.95, 3.9e-4, 0.2, 77, 42, 0, 1)
    L = lib.oracle_learner_create(C.byref(d), flat)
    lib.oracle_learner_rollout(L)
    nb = 128
    buf = lambda i, w: np.ctypeslib.as_array(lib.oracle_learner_buffer(L, i), (nb, w)).copy()
    obs, rets, dones, acts, vals, nlps = buf(0, 18), buf(1, 1), buf(2, 1), buf(3, 18), buf(4, 1), buf(5, 1)
    o = ol.Oracle()
    # stored values / neglogp are those of the stored (obs, action) pairs under the rollout policy
    _, v, _, mean = o.policy_step(flat, obs, None, "f64")
    assert rel_err(vals[:, 0], v) < 1e-5
    z = (acts - mean)
    assert rel_err(nlps[:, 0], 0.5 * (z * z).sum(1) + 18 * 0.9189385175704956) < 1e-5
    p0 = np.ctypeslib.as_array(lib.oracle_learner_params(L), (o.Pq,)).copy()
    assert np.array_equal(p0, flat)
    losses = np.zeros(5, np.float32)
    lib.oracle_learner_train(L, losses)
    p1 = np.ctypeslib.as_array(lib.oracle_learner_params(L), (o.Pq,)).copy()
    assert np.all(np.isfinite(losses)) and not np.array_equal(p0[:o.P], p1[:o.P]) and np.array_equal(p0[o.P:], p1[o.P:])
    # replay the 8 minibatch steps with the piecewise functions
    perms = ol.glibc_shuffle(42, nb, 2)
    th, m, vv = flat[:o.P].astype(np.float32), np.zeros(o.P, np.float32), np.zeros(o.P, np.float32)
    b1p, b2p = np.float32(c["beta1"]), np.float32(c["beta2"])
    for perm in perms:
        src = np.zeros(nb, np.int32)
        lib.oracle_perm_to_gather(perm, nb, src)
        for s in range(0, nb, 32):
            idx = src[s:s + 32]
            adv = o.advnorm(rets[idx, 0], vals[idx, 0], "f32")
            full = np.concatenate([th, flat[o.P:]])
            g, _ = o.loss_grad(full, obs[idx], acts[idx], adv, rets[idx, 0], nlps[idx, 0], vals[idx, 0], np.float32(0.2), "f32")
            th, m, vv, _, b1p, b2p, _ = o.clip_adam(np.float32(3.9e-4), th, m, vv, g, b1p, b2p, "f32")
    assert np.array_equal(th, p1[:o.P])
    lib.oracle_learner_destroy(L)


# ----------------------------------------------------------------------------------------------------------
# Parity pin: fixtures produced by EXECUTING the reference's own TensorFlow graph file node by node
# (oracle/tf_graph_exec.py + tests/golden/make_graph_exec_golden.py).  Forward GRAPH:1859-6866, loss
# 9210-11752, autodiff 11773-23699 (tie rules are graph nodes), clip 23738-25392, ApplyAdam 25426-31383.
# ----------------------------------------------------------------------------------------------------------
GX_ACT = ["init_4_5", "ckpt_4_5", "rand_8_8", "rand_64_64", "rand_128_128", "obs36_4_5", "obs1_4_5", "obs36_64_64",
          "ties_4_5", "ties_64_64"]
GX_TRAIN_FULL = ["init_4_5", "ckpt_4_5", "ties_4_5", "rand_8_8", "rand_64_64", "obs36_4_5", "obs1_4_5"]
GX_TRAIN_LIGHT = ["ties_64_64", "obs36_64_64", "rand_128_128"]


def _gx_oracle(k, kat):
    O, A, h1, h2 = (int(x) for x in k["dims"])
    c = kat["consts"]
    return ol.Oracle(obs_dim=O, act_dim=A, h1=h1, h2=h2, ent_coef=c["ent_coef"], vf_coef=c["vf_coef"], clip_norm=c["clip_norm"],
                     beta1=c["beta1"], beta2=c["beta2"], adam_eps=c["adam_eps"])


@pytest.mark.parametrize("case", GX_ACT)
def test_policy_step_vs_executed_graph(case, gx_act, kat):
    k = gx_act[case]
    o = _gx_oracle(k, kat)
    for prec, tol in (("f64", 1e-12), ("f32", 3e-6)):
        action, value, nlp, mean = o.policy_step(k["params"], k["obs"], k["eps"], prec)
        for got, name in ((action, "action"), (value, "value"), (nlp, "neglogp"), (mean, "mean")):
            assert rel_err(got, k[f"{name}_f64"]) < tol, (name, prec)
    # the fp32 oracle against TF's own arithmetic type executed in numpy
    action, value, nlp, mean = o.policy_step(k["params"], k["obs"], k["eps"], "f32")
    for got, name in ((action, "action"), (value, "value"), (nlp, "neglogp"), (mean, "mean")):
        assert rel_err(got, k[f"{name}_f32"]) < 3e-6, name


def _gx_replay(o, k, prec, steps):
    """Consecutive train steps with the piecewise oracle functions; yields per-step (losses, grads, theta, m, v, bpow, gnorm)."""
    dt = np.float32 if prec == "f32" else np.float64
    P = o.P
    th, m, v = k["params"][:P].astype(dt), np.zeros(P, dt), np.zeros(P, dt)
    b1p, b2p = dt(np.float32(o.hp.beta1)), dt(np.float32(o.hp.beta2))
    for s in range(steps):
        full = np.concatenate([th.astype(np.float32), k["params"][P:]])
        if prec == "f64":
            assert s == 0 or True
        g, losses = o.loss_grad(full, k["obs"][s], k["act"][s], k["adv"][s], k["ret"][s], k["old_nlp"][s], k["old_v"][s],
                                dt(k["cliprange"]), prec)
        th, m, v, _, b1p, b2p, gn = o.clip_adam(dt(k["lr"]), th, m, v, g, b1p, b2p, prec)
        yield losses, g, th, m, v, (b1p, b2p), gn


@pytest.mark.parametrize("case", GX_TRAIN_FULL + GX_TRAIN_LIGHT)
def test_train_step_vs_executed_graph(case, gx_train, kat):
    from ppo_cpp_b200.meta_graph import TENSOR_ORDER, param_layout
    k = gx_train[case]
    o = _gx_oracle(k, kat)
    O, A, h1, h2 = (int(x) for x in k["dims"])
    lay = param_layout(O, A, h1, h2)
    # fp64 oracle, first step: gradients (every tensor separately), losses, norm, Adam state — tight
    losses, g, th, m, v, bp, gn = next(_gx_replay(o, k, "f64", 1))
    assert np.allclose(losses, k["losses_f64"][0], rtol=1e-9, atol=1e-12)      # the graph's constants are fp32-baked
    for name in TENSOR_ORDER[:13]:
        off, shp = lay[name]
        n = int(np.prod(shp))
        assert rel_err(g[off:off + n], k["grads_f64"][0][off:off + n]) < 1e-10, name
    assert gn == pytest.approx(float(k["gnorm_f64"][0]), rel=1e-12)
    assert rel_err(th - k["params"][:o.P], k["theta_f64"][0] - k["params"][:o.P]) < 1e-9
    assert bp[0] == pytest.approx(k["bpow_f64"][0][0], rel=1e-12) and bp[1] == pytest.approx(k["bpow_f64"][0][1], rel=1e-12)
    if "m_f64" in k:
        assert rel_err(m, k["m_f64"][0]) < 1e-10 and rel_err(v, k["v_f64"][0]) < 1e-10
    # fp32 oracle, every consecutive step (weights fed back through fp32, as the reference's variables are):
    if "theta_f32" in k:
        steps = k["obs"].shape[0]
        for s, (losses, g, th, m, v, bp, gn) in enumerate(_gx_replay(o, k, "f32", steps)):
            assert np.allclose(losses, k["losses_f32"][s], rtol=2e-5, atol=1e-6), (s, losses, k["losses_f32"][s])
            for name in TENSOR_ORDER[:13]:
                off, shp = lay[name]
                n = int(np.prod(shp))
                assert rel_err(g[off:off + n], k["grads_f32"][s][off:off + n]) < 2e-5, (s, name)
            assert rel_err(th, k["theta_f32"][s]) < 1e-6 and rel_err(m, k["m_f32"][s]) < 2e-5 and rel_err(v, k["v_f32"][s]) < 2e-5
            assert bp[0] == pytest.approx(k["bpow_f32"][s][0], rel=1e-6) and bp[1] == pytest.approx(k["bpow_f32"][s][1], rel=1e-6)
            # and the fp32 run tracks the fp64 truth of the same step count
            assert rel_err(th, k["theta_f64"][s]) < 1e-6


def test_executed_graph_tie_rules_are_discriminating(gx_train):
    """The tie fixture really separates TF's rule (ties -> first argument, GRAPH:14975-15142) from a 0.5/0.5 split:
    samples 16..31 have (v-R)^2 == (vclip-R)^2 with v outside the clip range, so dL/dv = vf_coef*(v-R)/B under TF's
    rule, half of that under a split, and 0 if the clipped branch won.  vf/b's gradient is the sum of dL/dv."""
    k = gx_train["ties_4_5"]
    O, A, h1, h2 = (int(x) for x in k["dims"])
    from ppo_cpp_b200.meta_graph import param_layout
    off, _ = param_layout(O, A, h1, h2)["model/vf/b"]
    B = k["obs"].shape[1]
    v, R, ov, cr = 1.0, k["ret"][0].astype(np.float64), k["old_v"][0].astype(np.float64), float(k["cliprange"])
    vc = ov + np.clip(v - ov, -cr, cr)
    l1, l2 = (v - R) ** 2, (vc - R) ** 2
    inr = (np.abs(v - ov) <= cr)
    assert np.all(l1[16:32] == l2[16:32]) and not inr[16:32].any() and inr[:16].all()
    dv_tf = 0.5 * 0.5 / B * np.where(l1 >= l2, 2 * (v - R), 2 * (vc - R) * inr)
    assert k["grads_f64"][0][off] == pytest.approx(dv_tf.sum(), rel=1e-12)
    dv_split = 0.5 * 0.5 / B * np.where(l1 > l2, 2 * (v - R), np.where(l1 == l2, (v - R) + (vc - R) * inr, 2 * (vc - R) * inr))
    assert abs(dv_split.sum() - dv_tf.sum()) > 1e-3
