"""numpy interpreter for the reference's TensorFlow-1.14 graph files — TEST INFRASTRUCTURE ONLY.

The reference does all of its network arithmetic by handing a text-proto ``MetaGraphDef``
(``resources/ppo_cl/graphs/*.meta.txt``) to ``tensorflow::Session::Run`` (``ppo2/policies.hpp:37,53,68``,
``ppo2/ppo2.hpp:450``, ``ppo2/session_creator.hpp:38-54``).  TensorFlow is neither vendored nor installable
here, but the graph itself IS the normative statement of the arithmetic: 928 nodes built from the 55 primitive
ops of its ``stripped_op_list`` (GRAPH:2-1853).  This module executes those nodes one by one in numpy — forward
(GRAPH:1859-6866), loss (9210-11752), the autodiff sub-graph exactly as ``tf.gradients`` emitted it
(11773-23699, tie rules included: they are ``GreaterEqual``/``LessEqual`` + ``Select`` nodes), the global-norm clip
(23738-25392) and the 13 ``ApplyAdam`` ops plus the beta-power updates (25426-31383).  Nothing here restates
SURVEY.md prose: which op feeds which comes from the file, only the per-op kernels (Add, MatMul, Select, ...)
are restated, each a one-liner below.

Only ``tests/golden/make_graph_exec_golden.py`` (fixture generator, run in the build container where
``/root/reference`` exists) and ``tests/`` import this file.  The product path never does.

Arithmetic dtype: ``float32`` (TF's own) or ``float64`` (every DT_FLOAT tensor widened — the truth the 1e-5
tolerances are measured against; constants keep their baked fp32 values).

Shape generality: ``tf.gradients`` freezes the static shapes of forward operands into ``<fwd>_grad/Shape[_1]``
Const nodes.  With ``dynamic_grad_shapes=True`` those Consts are replaced by the run-time shape of the forward
node's operand (what TF emits when the static shape is unknown), so the SAME graph can be executed with
substituted variables of another hidden width or observation width.
"""
from __future__ import annotations

import re
from collections import deque
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

DT_FLOAT, DT_DOUBLE, DT_INT32, DT_STRING, DT_INT64, DT_BOOL = 1, 2, 3, 7, 9, 10
_MUTATING = ("Assign", "ApplyAdam")


def load_meta_graph(path: str):
    from google.protobuf import text_format
    from tensorboard.compat.proto import meta_graph_pb2
    mg = meta_graph_pb2.MetaGraphDef()
    with open(path) as f:
        text_format.Parse(f.read(), mg)
    return mg


def _split_ref(ref: str) -> Tuple[str, int, bool]:
    ctrl = ref.startswith("^")
    if ctrl:
        ref = ref[1:]
    if ":" in ref:
        n, k = ref.rsplit(":", 1)
        return n, int(k), ctrl
    return ref, 0, ctrl


class GraphExecutor:
    def __init__(self, meta_graph, dtype=np.float32, dynamic_grad_shapes: bool = False):
        self.gd = meta_graph.graph_def
        self.nodes = {n.name: n for n in self.gd.node}
        self.f = np.dtype(dtype).type
        self.vars: Dict[str, np.ndarray] = {}
        self.dynamic_grad_shapes = dynamic_grad_shapes
        self.executed: List[str] = []          # op trace of the last run (node names)

    # ---------------------------------------------------------------- helpers
    def _np_dtype(self, dt: int):
        return {DT_FLOAT: self.f, DT_DOUBLE: np.float64, DT_INT32: np.int32, DT_INT64: np.int64, DT_BOOL: np.bool_}[dt]

    def _const(self, node) -> np.ndarray:
        t = node.attr["value"].tensor
        shape = tuple(d.size for d in t.tensor_shape.dim)
        if t.dtype == DT_STRING:
            return np.array(list(t.string_val), dtype=object).reshape(shape) if shape else np.array(t.string_val[0], dtype=object)
        src = {DT_FLOAT: np.float32, DT_DOUBLE: np.float64, DT_INT32: np.int32, DT_INT64: np.int64, DT_BOOL: np.bool_}[t.dtype]
        if t.tensor_content:
            a = np.frombuffer(t.tensor_content, dtype=src).copy()
        else:
            vals = {DT_FLOAT: t.float_val, DT_DOUBLE: t.double_val, DT_INT32: t.int_val, DT_INT64: t.int64_val,
                    DT_BOOL: t.bool_val}[t.dtype]
            a = np.array(list(vals), dtype=src)
            n = int(np.prod(shape)) if shape else 1
            if a.size == 0:
                a = np.zeros(n, src)
            elif a.size < n:                                   # TF: the last value is repeated
                a = np.concatenate([a, np.full(n - a.size, a[-1], src)])
        return a.reshape(shape).astype(self._np_dtype(t.dtype))

    def variable_names(self) -> List[str]:
        return [n.name for n in self.gd.node if n.op == "VariableV2"]

    def init(self):
        """session->Run({}, {}, {"init"}) — ppo2/session_creator.hpp:54."""
        self.run([], {}, ["init"])

    def set_variable(self, name: str, value):
        self.vars[name] = np.asarray(value).astype(self.f if np.asarray(value).dtype.kind == "f" else np.asarray(value).dtype)

    # ---------------------------------------------------------------- scheduling
    def run(self, fetches: Sequence[str], feeds: Optional[Dict[str, np.ndarray]] = None, targets: Iterable[str] = ()):
        feeds = {(_split_ref(k)[0], _split_ref(k)[1]): v for k, v in (feeds or {}).items()}
        fed_nodes = {k[0] for k in feeds}
        want = [_split_ref(f)[:2] for f in fetches]
        roots = [w[0] for w in want] + list(targets)
        # transitive closure (a fed node cuts the traversal: its inputs are not needed)
        need, stack = set(), list(roots)
        while stack:
            n = stack.pop()
            if n in need:
                continue
            need.add(n)
            if n in fed_nodes:
                continue
            for ref in self._inputs(self.nodes[n]):
                stack.append(_split_ref(ref)[0])
        indeg, users = {}, {}
        for n in need:
            ins = [] if n in fed_nodes else {_split_ref(r)[0] for r in self._inputs(self.nodes[n])}
            indeg[n] = len(ins)
            for i in ins:
                users.setdefault(i, []).append(n)
        pure, mut = deque(), deque()
        for n in sorted(need):
            if indeg[n] == 0:
                (mut if self.nodes[n].op in _MUTATING else pure).append(n)
        vals: Dict[Tuple[str, int], np.ndarray] = dict(feeds)
        self.executed = []
        done = 0
        while pure or mut:
            # every read that is not control-dependent on a mutation happens before the mutations:
            # the gradients (clip_by_global_norm needs all 13) are computed from the pre-update variables.
            n = pure.popleft() if pure else mut.popleft()
            if n not in fed_nodes:
                outs = self._exec(self.nodes[n], vals)
                for k, v in enumerate(outs):
                    vals[(n, k)] = v
                self.executed.append(n)
            done += 1
            for u in users.get(n, []):
                indeg[u] -= 1
                if indeg[u] == 0:
                    (mut if self.nodes[u].op in _MUTATING else pure).append(u)
        assert done == len(need), "cycle in graph?"
        return [vals[w] for w in want]

    def _inputs(self, node) -> List[str]:
        ins = list(node.input)
        if self.dynamic_grad_shapes and node.op == "Const":
            fwd = self._frozen_shape_source(node.name)
            if fwd is not None:
                ins = [fwd]
        return ins

    _GRAD_SHAPE = re.compile(r"^loss/gradients/(.+)_grad/Shape(_1)?$")

    def _frozen_shape_source(self, name: str) -> Optional[str]:
        """`loss/gradients/<fwd>_grad/Shape[_1]` Const -> the forward operand whose static shape it froze."""
        m = self._GRAD_SHAPE.match(name)
        if not m or m.group(1) not in self.nodes:
            return None
        fwd = self.nodes[m.group(1)]
        if fwd.op not in ("Add", "Sub", "Mul", "RealDiv", "Maximum", "Minimum"):
            return None
        data = [r for r in fwd.input if not r.startswith("^")]
        return data[1 if m.group(2) else 0]

    # ---------------------------------------------------------------- op kernels
    def _exec(self, node, vals) -> tuple:
        op = node.op
        ins = []
        for ref in self._inputs(node):
            n, k, ctrl = _split_ref(ref)
            if not ctrl:
                ins.append(vals[(n, k)])
        a = node.attr
        f = self.f
        if op == "Const":
            if ins:                                             # dynamic_grad_shapes replacement
                return (np.array(np.shape(ins[0]), np.int32),)
            return (self._const(node),)
        if op == "Placeholder":
            raise KeyError(f"placeholder {node.name} needs a feed")
        if op == "PlaceholderWithDefault" or op == "Identity":
            return (ins[0],)
        if op == "NoOp":
            return ()
        if op == "VariableV2":
            if node.name not in self.vars:
                return (None,)                                  # uninitialised ref (only Assign may consume it)
            return (self.vars[node.name],)
        if op == "Assign":
            ref = _split_ref(node.input[0])[0]
            self.vars[ref] = np.array(ins[1])
            return (self.vars[ref],)
        if op == "Add":
            return (ins[0] + ins[1],)
        if op == "Sub":
            return (ins[0] - ins[1],)
        if op == "Mul":
            return (ins[0] * ins[1],)
        if op == "RealDiv":
            return (ins[0] / ins[1],)
        if op == "Neg":
            return (-ins[0],)
        if op == "Maximum":
            return (np.maximum(ins[0], ins[1]),)
        if op == "Minimum":
            return (np.minimum(ins[0], ins[1]),)
        if op == "Exp":
            return (np.exp(ins[0]),)
        if op == "Tanh":
            return (np.tanh(ins[0]),)
        if op == "TanhGrad":                                    # (y, dy) -> dy * (1 - y*y)
            return (ins[1] * (f(1) - ins[0] * ins[0]),)
        if op == "Square":
            return (ins[0] * ins[0],)
        if op == "Sqrt":
            return (np.sqrt(ins[0]),)
        if op == "Abs":
            return (np.abs(ins[0]),)
        if op == "Greater":
            return (ins[0] > ins[1],)
        if op == "GreaterEqual":
            return (ins[0] >= ins[1],)
        if op == "LessEqual":
            return (ins[0] <= ins[1],)
        if op == "IsFinite":
            return (np.isfinite(ins[0]),)
        if op == "Select":
            return (np.where(ins[0], ins[1], ins[2]),)
        if op == "Cast":
            return (np.asarray(ins[0]).astype(self._np_dtype(a["DstT"].type)),)
        if op == "AddN":
            out = ins[0]
            for x in ins[1:]:
                out = out + x
            return (out,)
        if op == "L2Loss":
            return (np.sum(ins[0] * ins[0]) / f(2),)
        if op in ("Sum", "Mean", "Prod"):
            axes = tuple(int(i) for i in np.atleast_1d(ins[1]))
            fn = {"Sum": np.sum, "Mean": np.mean, "Prod": np.prod}[op]
            x = np.asarray(ins[0])
            return (np.asarray(fn(x, axis=axes, keepdims=bool(a["keep_dims"].b)), dtype=x.dtype),)
        if op == "MatMul":
            x = ins[0].T if a["transpose_a"].b else ins[0]
            y = ins[1].T if a["transpose_b"].b else ins[1]
            return (x @ y,)
        if op == "Shape":
            return (np.array(np.shape(ins[0]), np.int32),)
        if op == "ShapeN":
            return tuple(np.array(np.shape(x), np.int32) for x in ins)
        if op == "Reshape":
            return (np.reshape(ins[0], tuple(int(i) for i in ins[1])),)
        if op == "Pack":
            return (np.stack(ins, axis=a["axis"].i),)
        if op == "ConcatV2":
            return (np.concatenate(ins[:-1], axis=int(ins[-1])),)
        if op == "Split":
            return tuple(np.split(ins[1], a["num_split"].i, axis=int(ins[0])))
        if op == "Slice":
            begin, size = [int(i) for i in ins[1]], [int(i) for i in ins[2]]
            sl = tuple(slice(b, None if s < 0 else b + s) for b, s in zip(begin, size))
            return (ins[0][sl],)
        if op == "StridedSlice":
            return (ins[0][self._strided(a, ins[1], ins[2], ins[3], np.ndim(ins[0]))],)
        if op == "StridedSliceGrad":
            out = np.zeros(tuple(int(i) for i in ins[0]), dtype=ins[4].dtype)
            sl = self._strided(a, ins[1], ins[2], ins[3], out.ndim)
            out[sl] = np.reshape(ins[4], out[sl].shape)
            return (out,)
        if op == "Tile":
            return (np.tile(ins[0], tuple(int(i) for i in ins[1])),)
        if op == "Fill":
            return (np.full(tuple(int(i) for i in np.atleast_1d(ins[0])), ins[1], dtype=np.asarray(ins[1]).dtype),)
        if op == "Range":
            return (np.arange(int(ins[0]), int(ins[1]), int(ins[2]), dtype=np.int32),)
        if op == "FloorDiv":
            return (np.floor_divide(ins[0], ins[1]),)
        if op == "FloorMod":
            return (np.mod(ins[0], ins[1]),)
        if op == "DynamicStitch":
            n = a["N"].i
            idx = [np.atleast_1d(np.asarray(i)) for i in ins[:n]]
            dat = [np.atleast_1d(np.asarray(d)) for d in ins[n:]]
            size = max(int(i.max()) for i in idx if i.size) + 1
            out = np.zeros((size,) + dat[0].shape[1:], dat[0].dtype)
            for i, d in zip(idx, dat):
                out[i] = d
            return (out,)
        if op == "BroadcastGradientArgs":
            return self._bcast_grad_args([int(i) for i in ins[0]], [int(i) for i in ins[1]])
        if op == "ConcatOffset":
            dim, off, outs = int(ins[0]), 0, []
            for s in ins[1:]:
                o = np.zeros(len(s), np.int32)
                o[dim] = off
                off += int(s[dim])
                outs.append(o)
            return tuple(outs)
        if op == "RandomStandardNormal":
            raise KeyError(f"{node.name}: graph seeds are 0/0 (non-reproducible in TF) — feed the noise explicitly")
        if op == "ApplyAdam":
            # TF 1.14 core/kernels/training_ops.cc ApplyAdam<CPUDevice,T> (use_nesterov=false):
            #   alpha = lr * sqrt(1 - beta2_power) / (1 - beta1_power)
            #   m += (g - m) * (1 - beta1);  v += (g*g - v) * (1 - beta2);  var -= (m * alpha) / (sqrt(v) + epsilon)
            assert not a["use_nesterov"].b
            var, m, v = (_split_ref(node.input[i])[0] for i in range(3))
            b1p, b2p, lr, b1, b2, eps, g = (ins[i] for i in range(3, 10))
            b1p, b2p, lr, b1, b2, eps = (f(x) for x in (b1p, b2p, lr, b1, b2, eps))
            alpha = lr * np.sqrt(f(1) - b2p) / (f(1) - b1p)
            mm = self.vars[m] + (g - self.vars[m]) * (f(1) - b1)
            vv = self.vars[v] + (g * g - self.vars[v]) * (f(1) - b2)
            self.vars[m], self.vars[v] = mm, vv
            self.vars[var] = self.vars[var] - (mm * alpha) / (np.sqrt(vv) + eps)
            return (self.vars[var],)
        raise NotImplementedError(f"op {op} ({node.name})")

    @staticmethod
    def _strided(a, begin, end, strides, ndim):
        assert a["ellipsis_mask"].i == 0 and a["new_axis_mask"].i == 0
        bm, em, sm = a["begin_mask"].i, a["end_mask"].i, a["shrink_axis_mask"].i
        sl = []
        for i in range(len(begin)):
            if sm >> i & 1:
                sl.append(int(begin[i]))
            else:
                sl.append(slice(None if bm >> i & 1 else int(begin[i]), None if em >> i & 1 else int(end[i]), int(strides[i])))
        return tuple(sl)

    @staticmethod
    def _bcast_grad_args(s0: List[int], s1: List[int]):
        """tensorflow/core/util/bcast.cc: reduction indices that undo the broadcast of s0 op s1."""
        n = max(len(s0), len(s1))
        x = [1] * (n - len(s0)) + s0
        y = [1] * (n - len(s1)) + s1
        r0, r1 = [], []
        for i in range(n):
            if x[i] == y[i]:
                if x[i] == 1:
                    r0.append(i)
                    r1.append(i)
            elif x[i] == 1:
                r0.append(i)
            elif y[i] == 1:
                r1.append(i)
            else:
                raise ValueError(f"incompatible shapes {s0} {s1}")
        return np.array(r0, np.int32), np.array(r1, np.int32)


# Tensor names the reference feeds and fetches (ppo2/ppo2.hpp:521-544).
ACT_FEED = "input/Ob:0"
ACT_NOISE = "output/random_normal/RandomStandardNormal:0"
ACT_FETCH = ("output/_action:0", "output/_value_flat:0", "output/_neglogp:0", "output/_deterministic_action:0")
TRAIN_FEEDS = dict(obs="train_model/input/Ob:0", actions="loss/action_ph:0", advs="loss/advs_ph:0",
                   returns="loss/rewards_ph:0", lr="loss/learning_rate_ph:0", cliprange="loss/clip_range_ph:0",
                   old_neglogp="loss/old_neglog_pac_ph:0", old_vpred="loss/old_vpred_ph:0")
LOSS_FETCH = ("loss/pg_loss:0", "loss/vf_loss:0", "loss/ppo2/entropy:0", "loss/approxkl:0", "loss/clipfrac:0")
TRAIN_TARGET = "ppo2/_train"
