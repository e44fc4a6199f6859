"""The reference-named C++ host classes (ppo_cpp_b200/host/): the reference's own VecEnv unit test restated
(CPU), and PPO2::learn/save/load/eval plus the CLI on the GPU."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_weights

HOST = os.path.join(ROOT, "ppo_cpp_b200", "host")
BIN = os.path.join(HOST, "bin")


def _build():
    subprocess.run(["make", "-s", "-C", HOST], check=True)


def test_vecenv_reference_unit_test():
    """test/vecenv_test.cpp:13-60 — VecEnv row routing with EnvMock for 1, 2 and 16 threads."""
    _build()
    r = subprocess.run([os.path.join(BIN, "vecenv_test")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "vecenv_test OK" in r.stdout


def test_tensorboard_event_file(tmp_path):
    """host/tensorboard.hpp (the reference's TensorboardWriter interface, ppo2/tensorboard.hpp:13-52, without TensorFlow):
    CRC-32C known answers, then an event file read back by TensorBoard's own reader, which checks the masked CRC of every
    record header and payload."""
    _build()
    r = subprocess.run([os.path.join(BIN, "tensorboard_test"), str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "tensorboard_test OK" in r.stdout, r.stdout + r.stderr
    files = [f for f in os.listdir(tmp_path) if f.startswith("PPO2.out.tfevents.")]
    assert len(files) == 1
    ea = pytest.importorskip("tensorboard.backend.event_processing.event_accumulator")
    acc = ea.EventAccumulator(str(tmp_path / files[0]))
    acc.Reload()
    assert sorted(acc.Tags()["scalars"]) == ["episode_reward", "other/scalar"]
    ev = acc.Scalars("episode_reward")
    assert len(ev) == 150
    assert [e.step for e in ev] == [20 * i for i in range(150)]
    assert np.allclose([e.value for e in ev], [np.float32(150.0) / np.float32(i + 1) for i in range(150)], rtol=1e-6)
    assert [e.wall_time for e in ev] == [1000.0 + i for i in range(150)]
    o = acc.Scalars("other/scalar")
    assert len(o) == 1 and o[0].step == 7 and o[0].value == -2.5


def _write_graph(tmp_path):
    from ppo_cpp_b200.meta_graph import write_meta_txt
    tensors, _ = load_weights("graph_4_5_init.npz")
    path = str(tmp_path / "ppo_cpp_[4_5].meta.txt")
    write_meta_txt(path, tensors, ent_coef=0.0007160293171182275)
    return path


@pytest.mark.gpu
def test_ppo2_learn_save_load_eval(tmp_path):
    _build()
    graph = _write_graph(tmp_path)
    r = subprocess.run([os.path.join(BIN, "host_gpu_test"), graph, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "host_gpu_test OK" in r.stdout


@pytest.mark.gpu
def test_cli_flags_and_csv_line(tmp_path):
    """Reference flags (ppo2.cpp:93-128) and the per-update stdout line fps,pg_loss,vf_loss,entropy,approxkl,clipfrac,"""
    _build()
    graph = _write_graph(tmp_path)
    cmd = [os.path.join(BIN, "ppo_cpp"), "-d", str(tmp_path / "exp"), "--graph_path", graph, "--id", "t0", "--steps", "8192", "--lr", "0.00039",
           "--ent", "0.0007", "--cr", "0.161", "--num_epochs", "3", "--batch_steps", "1024", "--threads", "2", "--seed", "5", "--saves", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    rows = [l for l in r.stdout.splitlines() if re.fullmatch(r"(-?[0-9.e+-]+,){6}", l)]
    assert len(rows) == 4  # 8192 / (2 envs * 1024 steps)
    vals = np.array([[float(x) for x in l.rstrip(",").split(",")] for l in rows])
    assert np.all(vals[:, 0] > 0) and np.all(np.isfinite(vals))
    ck = tmp_path / "exp" / "checkpoints" / "t0"
    names = sorted(os.listdir(ck))
    assert "t0.pkl.0.json" in names and "t0.pkl.1.json" in names and "t0.pkl.0.data-00000-of-00001" in names
    assert os.path.getsize(ck / "t0.pkl.0.data-00000-of-00001") == 1768  # 442 fp32, as the reference's checkpoint
    assert os.path.getsize(ck / "t0.pkl.0.index") == 498  # the bundle's table, the size of the reference's own *.pkl.71.index
    tb = tmp_path / "exp" / "tensorboard" / "t0"
    assert any(f.startswith("PPO2.out.tfevents.") for f in os.listdir(tb)), os.listdir(tb)  # TensorboardWriter (ppo2.hpp:248)
    # playback of the saved checkpoint (-p): restores weights + normaliser statistics, no training
    r2 = subprocess.run([os.path.join(BIN, "ppo_cpp"), "-g", graph, "-p", str(ck / "t0.pkl.1"), "--duration", "1.5", "--seed", "5"],
                        capture_output=True, text=True, timeout=120)
    assert r2.returncode == 0 and "episode_reward:" in r2.stdout, r2.stdout[-2000:] + r2.stderr[-2000:]
