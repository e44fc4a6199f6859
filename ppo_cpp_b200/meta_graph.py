"""Reader for the reference's TensorFlow-1.14 ``.meta.txt`` graph files (text-proto MetaGraphDef).

The reference executes the graph through ``tensorflow::Session`` (``ppo2/session_creator.hpp:23-57``);
this framework has no TensorFlow, so the graph file is only *read*: hidden sizes come from the
``VariableV2`` shapes of ``model/{pi,vf}_fc{k}/w``, the initial weights from the ``Const``
``tensor_content`` of ``model/*/Initializer/*`` and the constants that the graph bakes in
(entropy coefficient ``loss/mul_4/y``, value coefficient ``loss/mul_5/y``, clip norm
``loss/clip_by_global_norm/mul/x``, Adam ``ppo2/_train/{beta1,beta2,epsilon}``).

This is the Python twin of ``csrc/meta_parser.cpp`` (the one the product path uses); tests check that
both agree on the reference fixture.  No protobuf dependency: a 60-line recursive text-proto reader.
"""
from __future__ import annotations

import re
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

# Trainable tensors in the order of the graph's gradient list (GRAPH:23738-24074), then the
# untrained q head that the Saver still stores (GRAPH:32396).
TENSOR_ORDER = [
    "model/pi_fc0/w", "model/pi_fc0/b", "model/vf_fc0/w", "model/vf_fc0/b",
    "model/pi_fc1/w", "model/pi_fc1/b", "model/vf_fc1/w", "model/vf_fc1/b",
    "model/vf/w", "model/vf/b", "model/pi/w", "model/pi/b", "model/pi/logstd",
    "model/q/w", "model/q/b",
]
N_TRAINABLE_TENSORS = 13

_TOKEN = re.compile(r'\s*(?:([A-Za-z_][A-Za-z0-9_]*)|([{}:])|("(?:[^"\\]|\\.)*")|([-+0-9.eE]+[A-Za-z]*|-?inf|nan))')


def _unescape(s: str) -> bytes:
    """C-style unescape of a text-proto string literal body (octal \\ooo, \\n, \\t, \\", \\\\ ...)."""
    out = bytearray()
    i, n = 0, len(s)
    simple = {"n": 10, "t": 9, "r": 13, '"': 34, "'": 39, "\\": 92, "a": 7, "b": 8, "f": 12, "v": 11}
    while i < n:
        c = s[i]
        if c != "\\":
            out += c.encode("latin-1") if ord(c) < 256 else c.encode("utf-8")
            i += 1
            continue
        i += 1
        c = s[i]
        if c in "01234567":
            j = i
            while j < n and j < i + 3 and s[j] in "01234567":
                j += 1
            out.append(int(s[i:j], 8) & 0xFF)
            i = j
        elif c == "x":
            j = i + 1
            while j < n and j < i + 3 and s[j] in "0123456789abcdefABCDEF":
                j += 1
            out.append(int(s[i + 1:j], 16))
            i = j
        else:
            out.append(simple[c])
            i += 1
    return bytes(out)


def _parse_block(text: str, pos: int) -> Tuple[dict, int]:
    """Parse fields until the matching '}' (or EOF). Returns ({name: [values...]}, new_pos)."""
    msg: Dict[str, list] = {}
    n = len(text)
    while True:
        m = _TOKEN.match(text, pos)
        if not m:
            return msg, n
        if m.group(2) == "}":
            return msg, m.end()
        key = m.group(1)
        if key is None:
            raise ValueError(f"text-proto: expected field name at offset {pos}: {text[pos:pos+40]!r}")
        pos = m.end()
        m = _TOKEN.match(text, pos)
        if m.group(2) == ":":
            pos = m.end()
            m = _TOKEN.match(text, pos)
        if m.group(2) == "{":
            val, pos = _parse_block(text, m.end())
        elif m.group(3) is not None:
            val, pos = _unescape(m.group(3)[1:-1]), m.end()
        elif m.group(4) is not None:
            val, pos = m.group(4), m.end()
        else:  # enum / bool identifier
            val, pos = m.group(1), m.end()
        msg.setdefault(key, []).append(val)


@dataclass
class MetaInfo:
    hidden: List[int]
    obs_dim: int
    act_dim: int
    tensors: Dict[str, np.ndarray]
    ent_coef: float
    vf_coef: float
    clip_norm: float
    beta1: float
    beta2: float
    adam_eps: float
    tf_version: str = ""
    shapes: Dict[str, Tuple[int, ...]] = field(default_factory=dict)

    def flat_params(self, include_q: bool = True) -> np.ndarray:
        names = TENSOR_ORDER if include_q else TENSOR_ORDER[:N_TRAINABLE_TENSORS]
        return np.concatenate([self.tensors[n].ravel() for n in names]).astype(np.float32)


def _attr(node: dict, key: str) -> dict | None:
    for a in node.get("attr", []):
        if a["key"][0] == key.encode():
            return a["value"][0]
    return None


def _tensor_value(node: dict) -> np.ndarray:
    t = _attr(node, "value")["tensor"][0]
    dims = [int(d["size"][0]) for d in t.get("tensor_shape", [{}])[0].get("dim", [])]
    count = int(np.prod(dims)) if dims else 1
    if "tensor_content" in t:
        arr = np.frombuffer(t["tensor_content"][0], dtype="<f4").copy()
    elif "float_val" in t:
        vals = [float(v) for v in t["float_val"]]
        arr = np.full(count, vals[0], dtype=np.float32) if len(vals) == 1 else np.asarray(vals, np.float32)
    else:  # all-zero tensors carry neither field
        arr = np.zeros(count, dtype=np.float32)
    return arr.reshape(dims) if dims else arr.reshape(())


def parse_meta_txt(path: str) -> MetaInfo:
    with open(path, "r", encoding="latin-1") as f:
        text = f.read()
    root, _ = _parse_block(text, 0)
    graph = root["graph_def"][0]
    nodes = {n["name"][0].decode(): n for n in graph["node"]}

    shapes: Dict[str, Tuple[int, ...]] = {}
    tensors: Dict[str, np.ndarray] = {}
    for name in TENSOR_ORDER:
        var = nodes[name]
        assert var["op"][0] == b"VariableV2", name
        shp = _attr(var, "shape")["shape"][0]
        shapes[name] = tuple(int(d["size"][0]) for d in shp.get("dim", []))
        init = None
        for suffix in ("initial_value", "Const", "zeros"):
            cand = nodes.get(f"{name}/Initializer/{suffix}")
            if cand is not None and cand["op"][0] == b"Const":
                init = _tensor_value(cand)
                break
        if init is None:
            raise ValueError(f"no Const initializer found for {name}")
        tensors[name] = np.ascontiguousarray(init.reshape(shapes[name]), dtype=np.float32)

    def scalar(name: str) -> float:
        return float(_tensor_value(nodes[name]).reshape(-1)[0])

    hidden = [shapes["model/pi_fc0/w"][1], shapes["model/pi_fc1/w"][1]]
    ver = ""
    try:
        ver = root["meta_info_def"][0]["tensorflow_version"][0].decode()
    except Exception:
        pass
    return MetaInfo(
        hidden=hidden,
        obs_dim=shapes["model/pi_fc0/w"][0],
        act_dim=shapes["model/pi/w"][1],
        tensors=tensors,
        ent_coef=scalar("loss/mul_4/y"),
        vf_coef=scalar("loss/mul_5/y"),
        clip_norm=scalar("loss/clip_by_global_norm/mul/x"),
        beta1=scalar("ppo2/_train/beta1"),
        beta2=scalar("ppo2/_train/beta2"),
        adam_eps=scalar("ppo2/_train/epsilon"),
        tf_version=ver,
        shapes=shapes,
    )


# TF Saver V2 bundle: the .data file is the tensors' raw bytes concatenated in the order of the
# (sorted) .index keys (ppo2/ppo2.hpp:107-131 runs the graph's Saver; GRAPH:32396 tensor_names).
CKPT_ORDER = sorted(TENSOR_ORDER)


def read_checkpoint_data(data_path: str, shapes: Dict[str, Tuple[int, ...]]) -> Dict[str, np.ndarray]:
    raw = np.fromfile(data_path, dtype="<f4")
    out, off = {}, 0
    for name in CKPT_ORDER:
        n = int(np.prod(shapes[name]))
        out[name] = raw[off:off + n].reshape(shapes[name]).copy()
        off += n
    if off != raw.size:
        raise ValueError(f"checkpoint data has {raw.size} floats, shapes need {off}")
    return out


def param_layout(obs_dim: int, act_dim: int, h1: int, h2: int) -> Dict[str, Tuple[int, Tuple[int, ...]]]:
    """Offsets (in floats) of every tensor inside the flat parameter vector used by the core."""
    shp = {
        "model/pi_fc0/w": (obs_dim, h1), "model/pi_fc0/b": (h1,),
        "model/vf_fc0/w": (obs_dim, h1), "model/vf_fc0/b": (h1,),
        "model/pi_fc1/w": (h1, h2), "model/pi_fc1/b": (h2,),
        "model/vf_fc1/w": (h1, h2), "model/vf_fc1/b": (h2,),
        "model/vf/w": (h2, 1), "model/vf/b": (1,),
        "model/pi/w": (h2, act_dim), "model/pi/b": (act_dim,),
        "model/pi/logstd": (1, act_dim),
        "model/q/w": (h2, act_dim), "model/q/b": (act_dim,),
    }
    out, off = {}, 0
    for name in TENSOR_ORDER:
        out[name] = (off, shp[name])
        off += int(np.prod(shp[name]))
    out["__total__"] = (off, ())
    return out


def _escape(b: bytes) -> str:
    out = []
    for c in b:
        if c == 0x5C:
            out.append("\\\\")
        elif c == 0x22:
            out.append('\\"')
        elif c == 0x27:
            out.append("\\'")
        elif 32 <= c < 127:
            out.append(chr(c))
        else:
            out.append("\\%03o" % c)
    return "".join(out)


def write_meta_txt(path: str, tensors: Dict[str, np.ndarray], ent_coef: float = 0.0, vf_coef: float = 0.5,
                   clip_norm: float = 0.5, beta1: float = 0.9, beta2: float = 0.999, adam_eps: float = 1e-5) -> None:
    """Write a minimal .meta.txt (variables + initialisers + baked constants) that both readers accept.

    The reference's graphs come from a Stable-Baselines/TensorFlow export script that is not part of the
    repo; this writer lets a user produce a graph file for other hidden sizes without TensorFlow."""
    def dims(shape, ind):
        return "".join(f"{ind}dim {{\n{ind}  size: {d}\n{ind}}}\n" for d in shape)

    def const_node(name, arr):
        arr = np.asarray(arr, np.float32)
        if arr.ndim == 0:
            val = f"          float_val: {float(arr)!r}\n"
        elif not arr.any():
            val = ""
        else:
            val = f'          tensor_content: "{_escape(arr.astype("<f4").tobytes())}"\n'
        return (f'  node {{\n    name: "{name}"\n    op: "Const"\n    attr {{\n      key: "dtype"\n      value {{\n        type: DT_FLOAT\n      }}\n    }}\n'
                f'    attr {{\n      key: "value"\n      value {{\n        tensor {{\n          dtype: DT_FLOAT\n          tensor_shape {{\n'
                f'{dims(arr.shape, "            ")}          }}\n{val}        }}\n      }}\n    }}\n  }}\n')

    def var_node(name, shape):
        return (f'  node {{\n    name: "{name}"\n    op: "VariableV2"\n    attr {{\n      key: "dtype"\n      value {{\n        type: DT_FLOAT\n      }}\n    }}\n'
                f'    attr {{\n      key: "shape"\n      value {{\n        shape {{\n{dims(shape, "          ")}        }}\n      }}\n    }}\n  }}\n')

    parts = ['meta_info_def {\n  tensorflow_version: "1.14.0"\n}\ngraph_def {\n']
    for name in TENSOR_ORDER:
        arr = np.asarray(tensors[name], np.float32)
        parts.append(const_node(f"{name}/Initializer/initial_value", arr))
        parts.append(var_node(name, arr.shape))
    for name, val in (("loss/mul_4/y", ent_coef), ("loss/mul_5/y", vf_coef), ("loss/clip_by_global_norm/mul/x", clip_norm),
                      ("ppo2/_train/beta1", beta1), ("ppo2/_train/beta2", beta2), ("ppo2/_train/epsilon", adam_eps)):
        parts.append(const_node(name, np.float32(val)))
    parts.append("}\n")
    with open(path, "w", encoding="latin-1") as f:
        f.write("".join(parts))
