#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ from the reference tree.

Run in the build container (needs /root/reference, g++, torch):   python tests/golden/make_golden.py
The GPU box has no /root/reference; tests there read only the files this script wrote.

Sources of truth used here (none of them is the oracle under oracle/):
  * the reference's graph file (initial weights, baked constants) and shipped checkpoint (trained
    weights, VecNormalize statistics) — parsed, not executed (TensorFlow is not installable);
  * an independent numpy fp64 forward pass and a PyTorch fp64 *autograd* evaluation of the loss
    written from the graph description in SURVEY.md §3.4/§3.5 (gradients are autograd's, not hand-derived);
  * the real glibc srand()/rand() (via ctypes) and the real libstdc++ std::random_shuffle (a 10-line
    C++ program compiled here) for the permutation known answers;
  * Random123's published Philox4x32-10 known-answer vectors.
"""
import ctypes
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from ppo_cpp_b200.meta_graph import TENSOR_ORDER, parse_meta_txt, read_checkpoint_data  # noqa: E402

REF = "/root/reference"
GRAPH = REF + "/resources/ppo_cl/graphs/ppo_cpp_[4_5]_lr_0.0004_cr_0.1610_ent_0.0007.meta.txt"
CKPT = REF + "/resources/ppo_cl/2019-08-20_21_13_01_2859_0.pkl.71"

C_NLP = float(np.float32(0.9189385175704956))   # 0.5*log(2*pi) as baked (fp32)
C_ENT = float(np.float32(1.4189385175704956))   # 0.5*log(2*pi*e)


def np_forward(t, x):
    """numpy fp64 forward of the act model."""
    x = x.astype(np.float64)
    W = {k: v.astype(np.float64) for k, v in t.items()}
    h1 = np.tanh(x @ W["model/pi_fc0/w"] + W["model/pi_fc0/b"])
    h2 = np.tanh(h1 @ W["model/pi_fc1/w"] + W["model/pi_fc1/b"])
    mean = h2 @ W["model/pi/w"] + W["model/pi/b"]
    g1 = np.tanh(x @ W["model/vf_fc0/w"] + W["model/vf_fc0/b"])
    g2 = np.tanh(g1 @ W["model/vf_fc1/w"] + W["model/vf_fc1/b"])
    v = (g2 @ W["model/vf/w"] + W["model/vf/b"])[:, 0]
    return mean, v


def np_neglogp(t, a, mean):
    ls = t["model/pi/logstd"].astype(np.float64).reshape(1, -1)
    z = (a - mean) / np.exp(ls)
    return 0.5 * (z * z).sum(1) + C_NLP * a.shape[1] + ls.sum()


def torch_loss(t, obs, act, adv, ret, old_nlp, old_v, cr, ent_coef, vf_coef):
    """PyTorch fp64 restatement of the train graph; returns losses and autograd grads."""
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in t.items() if "/q/" not in k}
    x = torch.tensor(obs, dtype=torch.float64)
    a = torch.tensor(act, dtype=torch.float64)
    A = torch.tensor(adv, dtype=torch.float64)
    R = torch.tensor(ret, dtype=torch.float64)
    on = torch.tensor(old_nlp, dtype=torch.float64)
    ov = torch.tensor(old_v, dtype=torch.float64)
    h1 = torch.tanh(x @ P["model/pi_fc0/w"] + P["model/pi_fc0/b"])
    h2 = torch.tanh(h1 @ P["model/pi_fc1/w"] + P["model/pi_fc1/b"])
    mean = h2 @ P["model/pi/w"] + P["model/pi/b"]
    g1 = torch.tanh(x @ P["model/vf_fc0/w"] + P["model/vf_fc0/b"])
    g2 = torch.tanh(g1 @ P["model/vf_fc1/w"] + P["model/vf_fc1/b"])
    v = (g2 @ P["model/vf/w"] + P["model/vf/b"])[:, 0]
    ls = mean * 0.0 + P["model/pi/logstd"]
    std = torch.exp(ls)
    nlp = 0.5 * (((a - mean) / std) ** 2).sum(-1) + C_NLP * a.shape[1] + ls.sum(-1)
    entropy = (ls + C_ENT).sum(-1).mean()
    vclip = ov + torch.clamp(v - ov, -cr, cr)
    vf_loss = 0.5 * torch.maximum((v - R) ** 2, (vclip - R) ** 2).mean()
    ratio = torch.exp(on - nlp)
    pg_loss = torch.maximum(-A * ratio, -A * torch.clamp(ratio, 1 - cr, 1 + cr)).mean()
    approxkl = 0.5 * ((nlp - on) ** 2).mean()
    clipfrac = ((ratio - 1).abs() > cr).double().mean()
    loss = pg_loss - entropy * ent_coef + vf_loss * vf_coef
    loss.backward()
    grads = np.concatenate([P[n].grad.numpy().ravel() for n in TENSOR_ORDER[:13]])
    losses = np.array([pg_loss.item(), vf_loss.item(), entropy.item(), approxkl.item(), clipfrac.item()])
    return losses, grads


def np_clip_adam(theta, m, v, g, lr, b1, b2, eps, b1p, b2p, clip):
    gn = np.sqrt((g * g).sum())
    scale = clip * min(1.0 / gn, 1.0 / clip)
    g = g * scale
    alpha = lr * np.sqrt(1 - b2p) / (1 - b1p)
    m = m + (g - m) * (1 - b1)
    v = v + (g * g - v) * (1 - b2)
    theta = theta - (m * alpha) / (np.sqrt(v) + eps)
    return theta, m, v, gn


def minibatch(rng, t, B, scale_obs=1.0):
    O, A = t["model/pi_fc0/w"].shape[0], t["model/pi/w"].shape[1]
    obs = (rng.standard_normal((B, O)) * scale_obs).astype(np.float32)
    mean, v = np_forward(t, obs)
    std = np.exp(t["model/pi/logstd"].astype(np.float64).reshape(1, -1))
    act = (mean + std * rng.standard_normal((B, A))).astype(np.float32)
    old_nlp = (np_neglogp(t, act.astype(np.float64), mean) + 0.05 * rng.standard_normal(B)).astype(np.float32)
    old_v = (v + 0.3 * rng.standard_normal(B)).astype(np.float32)
    ret = (v + 0.5 * rng.standard_normal(B)).astype(np.float32)
    adv = rng.standard_normal(B).astype(np.float32)
    return obs, act, adv, ret, old_nlp, old_v


def orthogonal_tensors(rng, O, A, h1, h2):
    def ortho(shape, scale):
        a = rng.standard_normal(shape)
        u, _, vt = np.linalg.svd(a, full_matrices=False)
        q = u if u.shape == shape else vt
        return (scale * q).astype(np.float32)
    t = {
        "model/pi_fc0/w": ortho((O, h1), np.sqrt(2)), "model/pi_fc0/b": (0.1 * rng.standard_normal(h1)).astype(np.float32),
        "model/vf_fc0/w": ortho((O, h1), np.sqrt(2)), "model/vf_fc0/b": (0.1 * rng.standard_normal(h1)).astype(np.float32),
        "model/pi_fc1/w": ortho((h1, h2), np.sqrt(2)), "model/pi_fc1/b": (0.1 * rng.standard_normal(h2)).astype(np.float32),
        "model/vf_fc1/w": ortho((h1, h2), np.sqrt(2)), "model/vf_fc1/b": (0.1 * rng.standard_normal(h2)).astype(np.float32),
        "model/vf/w": ortho((h2, 1), 1.0), "model/vf/b": np.zeros(1, np.float32),
        "model/pi/w": ortho((h2, A), 0.3), "model/pi/b": (0.05 * rng.standard_normal(A)).astype(np.float32),
        "model/pi/logstd": (-0.5 + 0.2 * rng.standard_normal((1, A))).astype(np.float32),
        "model/q/w": ortho((h2, A), 0.01), "model/q/b": np.zeros(A, np.float32),
    }
    return t


def glibc_goldens():
    libc = ctypes.CDLL("libc.so.6")
    libc.rand.restype = ctypes.c_int
    out = {"rand": {}, "shuffle": {}}
    for seed in (0, 1, 42, 12345, 2147483647):
        libc.srand(seed)
        out["rand"][str(seed)] = [libc.rand() for _ in range(400)]
    src = r"""
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
int main(int argc, char** argv) {
  unsigned seed = std::strtoul(argv[1], 0, 10); int n = std::atoi(argv[2]); int epochs = std::atoi(argv[3]);
  std::srand(seed); std::vector<int> p(n); for (int i = 0; i < n; ++i) p[i] = i;
  for (int e = 0; e < epochs; ++e) { std::random_shuffle(p.data(), p.data() + n);
    for (int i = 0; i < n; ++i) std::printf("%d ", p[i]); std::printf("\n"); }
}
"""
    with tempfile.TemporaryDirectory() as td:
        with open(td + "/s.cpp", "w") as f:
            f.write(src)
        subprocess.run(["g++", "-std=c++14", "-O1", "-w", "-o", td + "/s", td + "/s.cpp"], check=True)
        for seed, n, epochs in ((42, 8, 1), (1, 64, 3), (12345, 2048, 3), (7, 1, 2), (7, 2, 2)):
            r = subprocess.run([td + "/s", str(seed), str(n), str(epochs)], check=True, capture_output=True, text=True)
            out["shuffle"][f"{seed}_{n}_{epochs}"] = [[int(x) for x in line.split()] for line in r.stdout.strip().split("\n")]
    return out


def main():
    rng = np.random.default_rng(20191029)
    meta = parse_meta_txt(GRAPH)
    consts = dict(ent_coef=meta.ent_coef, vf_coef=meta.vf_coef, clip_norm=meta.clip_norm, beta1=meta.beta1,
                  beta2=meta.beta2, adam_eps=meta.adam_eps, hidden=meta.hidden, tf_version=meta.tf_version)
    np.savez(HERE + "/graph_4_5_init.npz", **{k.replace("/", "__"): v for k, v in meta.tensors.items()})
    ck = read_checkpoint_data(CKPT + ".data-00000-of-00001", meta.shapes)
    np.savez(HERE + "/ckpt_71_weights.npz", **{k.replace("/", "__"): v for k, v in ck.items()})
    with open(CKPT + ".json") as f:
        ckj = json.load(f)

    kat = {"consts": consts, "ckpt_json": ckj}
    # ---- forward known answers (numpy fp64) ----
    fwd = {}
    for wname, t in (("init", meta.tensors), ("ckpt", ck)):
        obs = np.stack([np.zeros(18), np.linspace(-1, 1, 18)] + [rng.standard_normal(18) for _ in range(6)]).astype(np.float32)
        eps = rng.standard_normal((obs.shape[0], 18)).astype(np.float32)
        mean, v = np_forward(t, obs)
        act = mean + np.exp(t["model/pi/logstd"].astype(np.float64).reshape(1, -1)) * eps
        fwd[wname] = dict(obs=obs, eps=eps, mean=mean, value=v, action=act, neglogp=np_neglogp(t, act, mean),
                          neglogp_at_mean=np_neglogp(t, mean, mean))
    np.savez(HERE + "/forward_kat.npz", **{f"{w}__{k}": v for w, d in fwd.items() for k, v in d.items()})
    kat["survey_8c"] = {  # numbers recorded in SURVEY.md §8(c)2 at survey time, re-derived here
        "ckpt_obs0_mean0_4": fwd["ckpt"]["mean"][0, :4].tolist(), "ckpt_obs0_value": float(fwd["ckpt"]["value"][0]),
        "ckpt_lin_mean0_4": fwd["ckpt"]["mean"][1, :4].tolist(), "ckpt_lin_value": float(fwd["ckpt"]["value"][1]),
        "ckpt_sum_logstd": float(ck["model/pi/logstd"].astype(np.float64).sum()),
        "ckpt_neglogp_at_mean": float(fwd["ckpt"]["neglogp_at_mean"][0]),
    }

    # ---- loss / gradient / Adam known answers (PyTorch fp64 autograd) ----
    cases = {}
    wide = orthogonal_tensors(rng, 18, 18, 64, 64)
    odd = orthogonal_tensors(rng, 18, 18, 8, 8)
    for cname, t, B, cr in (("init_4_5", meta.tensors, 96, 0.2), ("ckpt_4_5", ck, 257, 0.16102319955825806),
                            ("rand_64_64", wide, 128, 0.2), ("rand_8_8", odd, 33, 0.1)):
        obs, act, adv, ret, old_nlp, old_v = minibatch(rng, t, B)
        losses, grads = torch_loss(t, obs, act, adv, ret, old_nlp, old_v, cr, meta.ent_coef, meta.vf_coef)
        theta = np.concatenate([t[n].ravel() for n in TENSOR_ORDER[:13]]).astype(np.float64)
        m0 = 0.01 * rng.standard_normal(theta.size)
        v0 = 1e-4 * rng.random(theta.size)
        th1, m1, v1, gn = np_clip_adam(theta, m0, v0, grads, 3.9e-4, float(np.float32(0.9)), float(np.float32(0.999)),
                                       float(np.float32(1e-5)), 0.9 ** 3, 0.999 ** 3, 0.5)
        cases[cname] = dict(params=np.concatenate([t[n].ravel() for n in TENSOR_ORDER]).astype(np.float32),
                            hidden=np.array([t["model/pi_fc0/w"].shape[1], t["model/pi_fc1/w"].shape[1]]),
                            obs=obs, act=act, adv=adv, ret=ret, old_nlp=old_nlp, old_v=old_v, cliprange=np.float64(cr),
                            losses=losses, grads=grads, adam_m0=m0, adam_v0=v0, adam_theta1=th1, adam_m1=m1, adam_v1=v1,
                            gnorm=np.float64(gn))
    np.savez(HERE + "/loss_grad_kat.npz", **{f"{c}__{k}": v for c, d in cases.items() for k, v in d.items()})

    kat["glibc"] = glibc_goldens()
    kat["philox4x32_10"] = [  # Random123 kat_vectors
        {"ctr": [0, 0, 0, 0], "key": [0, 0], "out": [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]},
        {"ctr": [0xffffffff] * 4, "key": [0xffffffff] * 2, "out": [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]},
        {"ctr": [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], "key": [0xa4093822, 0x299f31d0],
         "out": [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]},
    ]
    with open(HERE + "/kat.json", "w") as f:
        json.dump(kat, f, indent=1)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
