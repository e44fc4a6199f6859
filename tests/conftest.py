import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_npz_tree(name):
    """npz with keys 'case__field' -> {case: {field: array}}"""
    z = np.load(os.path.join(GOLDEN, name))
    out = {}
    for k in z.files:
        c, f = k.split("__", 1)
        out.setdefault(c, {})[f] = z[k]
    return out


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def forward_kat():
    return load_npz_tree("forward_kat.npz")


@pytest.fixture(scope="session")
def loss_kat():
    return load_npz_tree("loss_grad_kat.npz")


@pytest.fixture(scope="session")
def gx_act():
    """act-model outputs of the reference's graph, executed node by node (make_graph_exec_golden.py)."""
    z = np.load(os.path.join(GOLDEN, "graph_exec_act.npz"))
    out = {}
    for k in z.files:
        if "__" in k:
            c, f = k.split("__", 1)
            out.setdefault(c, {})[f] = z[k]
    return out


@pytest.fixture(scope="session")
def gx_train():
    """train-model losses / gradients / clip / ApplyAdam state of the executed reference graph."""
    z = np.load(os.path.join(GOLDEN, "graph_exec_train.npz"))
    out = {}
    for k in z.files:
        if "__" in k:
            c, f = k.split("__", 1)
            out.setdefault(c, {})[f] = z[k]
    return out


def load_weights(name):
    from ppo_cpp_b200.meta_graph import TENSOR_ORDER
    z = np.load(os.path.join(GOLDEN, name))
    t = {k.replace("__", "/"): z[k] for k in z.files}
    flat = np.concatenate([t[n].ravel() for n in TENSOR_ORDER]).astype(np.float32)
    return t, flat


@pytest.fixture(scope="session")
def init_weights():
    return load_weights("graph_4_5_init.npz")


@pytest.fixture(scope="session")
def ckpt_weights():
    return load_weights("ckpt_71_weights.npz")


def rel_err(a, b):
    """max |a-b| / max(|b|) — tensor-wise relative error used for the 1e-5 parity bar."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    denom = max(float(np.max(np.abs(b))), 1e-30)
    return float(np.max(np.abs(a - b))) / denom
