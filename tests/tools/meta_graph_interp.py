#!/usr/bin/env python3
"""Executor for the reference's OWN TensorFlow graph (test infrastructure, CPU, numpy).

The reference ships the exact op graph of its inference path as
``/root/reference/catfish/ResNetRNN/checkpoints/ckpnt-30000.meta`` (a TF-1.10
``MetaGraphDef``; the graph ``catfish/models/rnn_class.py:38-54`` builds and
``rnn_class.py:213-219`` runs).  TensorFlow is not installed here, so this module
executes that GraphDef node by node with numpy: the 1 049-node subgraph feeding
``accuracy/Sigmoid`` (44 op types, six ``while_loop`` frames with TensorArrays).
Nothing here restates the model: layer order, padding, BN arithmetic, gate order,
time reversal and concat order all come from the graph file; the weights come from
the shipped ``.index/.data`` bundle.  It is what pins ``oracle/tf_graph.py``.

Semantics implemented (TF 1.x executor, restricted to what this graph uses):

* dataflow ops: their documented kernels (``Conv2D`` NHWC/SAME/stride 1, ``MatMul``
  with transpose attrs, ``StridedSlice`` with begin/end/shrink masks, ...); every
  attr a kernel depends on is asserted, an unknown op raises.
* control flow: ``Enter/Merge/Switch/LoopCond/NextIteration/Exit`` evaluated per
  iteration of their (non-nested) frame: ``Merge@0 = Enter``, ``Merge@i =
  NextIteration@(i-1)``, ``Switch`` forwards its data to output 1 while the
  predicate holds and to output 0 (-> ``Exit``) at the first iteration it fails.
* TensorArrays: the ``flow`` value carries the array contents (a persistent
  index -> tensor map), which is equivalent to TF's resource + flow ordering
  because every read in this graph is ordered after its writes by the flow edge.
* ``RandomUniform`` (dropout mask, ``rnn_class.py:151-154``): seeded U[0,1); with
  ``dropout`` fed 1.0 as ``RNN.infer`` does the mask is floor(1+u) = 1 for every u.

``float_dtype=np.float64`` widens every DT_FLOAT const / variable / placeholder so
the graph can be compared with the fp64 restatement at 1e-12; ``np.float32`` runs
it in the reference's own arithmetic type.

CLI:  python tests/tools/meta_graph_interp.py   (prints a self-check)
"""
import os
import sys

import numpy as np

REF_CKPT_DIR = "/root/reference/catfish/ResNetRNN/checkpoints"
META = os.path.join(REF_CKPT_DIR, "ckpnt-30000.meta")
PREDICTIONS = "accuracy/Sigmoid"          # self.predictions, rnn_class.py:84
ACCURACY = "accuracy/Mean"                # self.accuracy, rnn_class.py:85-86
LOSS = "loss/Mean"                        # self.loss, rnn_class.py:75-76
LOGITS = "Reshape_1"                      # self.logits, rnn_class.py:181
Y_PLACEHOLDER = "data/Placeholder_1"      # self.y, rnn_class.py:160
X_PLACEHOLDER = "data/Placeholder"        # self.x, rnn_class.py:159
DROPOUT_PLACEHOLDER = "dropout"           # self.p_dropout, rnn_class.py:161

DT_FLOAT, DT_INT32, DT_BOOL, DT_INT64 = 1, 3, 10, 9


def available():
    return os.path.exists(META)


def _split(ref):
    """'name:1' -> ('name', 1); control inputs ('^name') are ordering-only."""
    if ref.startswith("^"):
        return None
    if ":" in ref:
        n, i = ref.rsplit(":", 1)
        return n, int(i)
    return ref, 0


class _Flow(object):
    """TensorArray contents travelling on the flow edge."""

    def __init__(self, size, items=None):
        self.size = size
        self.items = items or {}

    def write(self, idx, val):
        items = dict(self.items)
        items[int(idx)] = val
        return _Flow(self.size, items)


class MetaGraph(object):
    def __init__(self, meta_path=META, variables=None, float_dtype=np.float32, seed=0, check_shapes=True):
        self.check_shapes = check_shapes
        from tensorboard.compat.proto import meta_graph_pb2
        from tensorboard.util import tensor_util
        self._make_ndarray = tensor_util.make_ndarray
        mg = meta_graph_pb2.MetaGraphDef()
        with open(meta_path, "rb") as f:
            mg.ParseFromString(f.read())
        self.tf_version = mg.meta_info_def.tensorflow_version
        self.nodes = {n.name: n for n in mg.graph_def.node}
        self.fd = np.dtype(float_dtype)
        self.variables = variables or {}
        self.rng = np.random.default_rng(seed)
        self._frames()
        self.executed = {}

    # -------------------------------------------------------------- structure
    def _data_inputs(self, node):
        return [r for r in (_split(i) for i in node.input) if r is not None]

    def _frames(self):
        """frame[name] for nodes inside a while frame (Exit counts as outside)."""
        consumers = {}
        for n in self.nodes.values():
            for src, _ in self._data_inputs(n):
                consumers.setdefault(src, []).append(n.name)
        self.frame = {}
        self.frame_nodes = {}
        for n in self.nodes.values():
            if n.op != "Enter":
                continue
            f = n.attr["frame_name"].s.decode()
            stack = [n.name]
            while stack:
                cur = stack.pop()
                if cur in self.frame:
                    assert self.frame[cur] == f, "nested/crossing frames are not supported"
                    continue
                if self.nodes[cur].op == "Exit":
                    continue
                self.frame[cur] = f
                self.frame_nodes.setdefault(f, []).append(cur)
                stack.extend(consumers.get(cur, []))

    def subgraph(self, target):
        seen, stack = set(), [target]
        while stack:
            n = stack.pop()
            if n in seen:
                continue
            seen.add(n)
            for i in self.nodes[n].input:
                stack.append(i.lstrip("^").split(":")[0])
        return seen

    # -------------------------------------------------------------- execution
    def run(self, fetch, feed):
        self.feed = {k: v for k, v in feed.items()}
        self.memo = {}
        old = sys.getrecursionlimit()
        sys.setrecursionlimit(max(old, 20000))
        try:
            name, idx = _split(fetch)
            return self._eval(name, idx, None)
        finally:
            sys.setrecursionlimit(old)

    def _eval(self, name, idx, it):
        node = self.nodes[name]
        if name in self.feed and node.op != "Placeholder":
            # fed intermediate tensor (TF allows feeding any tensor): used to route the graph's own
            # sub-graphs the way the reference's model variants wire them (neural_network.py:17-18,
            # resnet_class.py:23)
            v = np.asarray(self.feed[name])
            return v.astype(self.fd) if v.dtype.kind == "f" else v
        ctx = it if name in self.frame else None
        if name in self.frame:
            assert it is not None, "frame node %s evaluated outside its frame" % name
        key = (name, ctx)
        if key not in self.memo:
            self.memo[key] = self._exec(node, ctx)
            self.executed[node.op] = self.executed.get(node.op, 0) + 1
        out = self.memo[key]
        val = out[idx]
        assert val is not _DEAD, "dead tensor %s:%d consumed at iteration %s" % (name, idx, ctx)
        return val

    def _in(self, node, k, it):
        src, idx = self._data_inputs(node)[k]
        return self._eval(src, idx, it)

    def _ins(self, node, it):
        return [self._eval(s, i, it) for s, i in self._data_inputs(node)]

    def _np_dtype(self, dt):
        return {DT_FLOAT: self.fd, DT_INT32: np.dtype(np.int32), DT_INT64: np.dtype(np.int64),
                DT_BOOL: np.dtype(bool)}[dt]

    def _exec(self, node, it):
        op = node.op
        fn = getattr(self, "_op_" + op, None)
        if fn is None:
            raise NotImplementedError("op %s (%s)" % (op, node.name))
        out = fn(node, it)
        return out if isinstance(out, tuple) else (out,)

    # ---- sources
    def _op_Const(self, n, it):
        v = self._make_ndarray(n.attr["value"].tensor)
        return v.astype(self.fd) if v.dtype == np.float32 else v

    def _op_Placeholder(self, n, it):
        v = np.asarray(self.feed[n.name])
        dt = self._np_dtype(n.attr["dtype"].type)
        if dt == self.fd:
            v = v.astype(np.float32)         # the feed is cast to the placeholder's DT_FLOAT first
        return v.astype(dt)

    def _op_VariableV2(self, n, it):
        v = np.asarray(self.variables[n.name])
        want = [d.size for d in n.attr["shape"].shape.dim]
        assert list(v.shape) == want or not self.check_shapes, (n.name, v.shape, want)
        assert v.dtype == np.float32 and n.attr["dtype"].type == DT_FLOAT
        return v.astype(self.fd)

    def _op_RandomUniform(self, n, it):
        shape = self._in(n, 0, it)
        return self.rng.random(tuple(int(s) for s in shape)).astype(self._np_dtype(n.attr["dtype"].type))

    # ---- elementwise / shape
    def _op_Identity(self, n, it):
        return self._in(n, 0, it)

    def _bin(self, n, it, f):
        a, b = self._ins(n, it)
        r = f(np.asarray(a), np.asarray(b))
        return r

    def _op_Add(self, n, it):
        return self._bin(n, it, np.add)

    def _op_Sub(self, n, it):
        return self._bin(n, it, np.subtract)

    def _op_Mul(self, n, it):
        return self._bin(n, it, np.multiply)

    def _op_RealDiv(self, n, it):
        return self._bin(n, it, np.true_divide)

    def _op_Maximum(self, n, it):
        return self._bin(n, it, np.maximum)

    def _op_Minimum(self, n, it):
        return self._bin(n, it, np.minimum)

    def _op_Less(self, n, it):
        return self._bin(n, it, np.less)

    def _op_LogicalAnd(self, n, it):
        return self._bin(n, it, np.logical_and)

    def _op_Floor(self, n, it):
        return np.floor(self._in(n, 0, it))

    def _op_Relu(self, n, it):
        x = self._in(n, 0, it)
        return np.maximum(x, x.dtype.type(0))

    def _op_Rsqrt(self, n, it):
        x = self._in(n, 0, it)
        return (x.dtype.type(1) / np.sqrt(x)).astype(x.dtype)

    def _op_Sigmoid(self, n, it):
        x = self._in(n, 0, it)
        with np.errstate(over="ignore"):
            return (x.dtype.type(1) / (x.dtype.type(1) + np.exp(-x))).astype(x.dtype)

    def _op_Tanh(self, n, it):
        return np.tanh(self._in(n, 0, it))

    def _op_BiasAdd(self, n, it):
        x, b = self._ins(n, it)
        fmt = n.attr["data_format"].s.decode() or "NHWC"
        assert fmt == "NHWC" and b.ndim == 1 and b.shape[0] == x.shape[-1]
        return x + b

    def _op_Shape(self, n, it):
        return np.array(np.shape(self._in(n, 0, it)), self._np_dtype(n.attr["out_type"].type or DT_INT32))

    def _op_Reshape(self, n, it):
        x, s = self._ins(n, it)
        return np.reshape(x, tuple(int(v) for v in s))

    def _op_ExpandDims(self, n, it):
        x, ax = self._ins(n, it)
        return np.expand_dims(x, int(ax))

    def _op_Squeeze(self, n, it):
        x = self._in(n, 0, it)
        dims = tuple(int(d) for d in n.attr["squeeze_dims"].list.i)
        return np.squeeze(x, axis=dims if dims else None)

    def _op_Transpose(self, n, it):
        x, perm = self._ins(n, it)
        return np.transpose(x, tuple(int(p) for p in perm))

    def _op_ReverseV2(self, n, it):
        x, axes = self._ins(n, it)
        return np.flip(x, tuple(int(a) for a in np.atleast_1d(axes)))

    def _op_ConcatV2(self, n, it):
        vals = self._ins(n, it)
        assert len(vals) == n.attr["N"].i + 1
        return np.concatenate([np.asarray(v) for v in vals[:-1]], axis=int(vals[-1]))

    def _op_Split(self, n, it):
        ax, x = self._ins(n, it)            # Split(split_dim, value)
        return tuple(np.split(x, n.attr["num_split"].i, axis=int(ax)))

    def _op_Fill(self, n, it):
        dims, v = self._ins(n, it)
        return np.full(tuple(int(d) for d in np.atleast_1d(dims)), v, dtype=np.asarray(v).dtype)

    def _op_Range(self, n, it):
        a, b, d = self._ins(n, it)
        return np.arange(a, b, d, dtype=np.asarray(a).dtype)

    def _op_StridedSlice(self, n, it):
        x, begin, end, strides = self._ins(n, it)
        a = n.attr
        assert a["ellipsis_mask"].i == 0 and a["new_axis_mask"].i == 0
        bm, em, sm = a["begin_mask"].i, a["end_mask"].i, a["shrink_axis_mask"].i
        x = np.asarray(x)
        idx = []
        for d in range(len(begin)):
            st = int(strides[d])
            if (sm >> d) & 1:
                idx.append(int(begin[d]))
                continue
            b = None if (bm >> d) & 1 else int(begin[d])
            e = None if (em >> d) & 1 else int(end[d])
            idx.append(slice(b, e, st))
        return x[tuple(idx)]

    # ---- loss / accuracy heads (rnn_class.py:73-88; run by test_network, rnn_class.py:241)
    def _op_Equal(self, n, it):
        return self._bin(n, it, np.equal)

    def _op_Greater(self, n, it):
        return self._bin(n, it, np.greater)

    def _op_GreaterEqual(self, n, it):
        return self._bin(n, it, np.greater_equal)

    def _op_Select(self, n, it):
        c, a, b = self._ins(n, it)
        return np.where(c, a, b)

    def _op_Neg(self, n, it):
        return -self._in(n, 0, it)

    def _op_Exp(self, n, it):
        return np.exp(self._in(n, 0, it))

    def _op_Log1p(self, n, it):
        return np.log1p(self._in(n, 0, it))

    def _op_ZerosLike(self, n, it):
        return np.zeros_like(self._in(n, 0, it))

    def _op_Round(self, n, it):
        return np.rint(self._in(n, 0, it))          # TF rounds half to even

    def _op_Cast(self, n, it):
        return np.asarray(self._in(n, 0, it)).astype(self._np_dtype(n.attr["DstT"].type))

    def _reduce(self, n, it, f):
        x, axes = self._ins(n, it)
        x = np.asarray(x)
        axes = tuple(int(a) for a in np.atleast_1d(axes))
        return f(x, axis=axes, keepdims=bool(n.attr["keep_dims"].b), dtype=x.dtype)

    def _op_Sum(self, n, it):
        return self._reduce(n, it, np.sum)

    def _op_Mean(self, n, it):
        return self._reduce(n, it, np.mean)

    # ---- contractions
    def _op_MatMul(self, n, it):
        a, b = self._ins(n, it)
        if n.attr["transpose_a"].b:
            a = a.T
        if n.attr["transpose_b"].b:
            b = b.T
        return a @ b

    def _op_Conv2D(self, n, it):
        x, w = self._ins(n, it)
        a = n.attr
        assert (a["data_format"].s.decode() or "NHWC") == "NHWC"
        assert list(a["strides"].list.i) == [1, 1, 1, 1]
        assert list(a["dilations"].list.i) in ([], [1, 1, 1, 1])
        assert a["padding"].s == b"SAME"
        kh, kw, cin, cout = w.shape
        b, h, wd, c = x.shape
        assert c == cin
        # SAME, stride 1: total pad k-1, the smaller half in front (TF: pad_before = (k-1)//2)
        ph, pw = (kh - 1) // 2, (kw - 1) // 2
        xp = np.zeros((b, h + kh - 1, wd + kw - 1, c), x.dtype)
        xp[:, ph:ph + h, pw:pw + wd] = x
        y = np.zeros((b, h, wd, cout), x.dtype)
        for i in range(kh):
            for j in range(kw):
                y += xp[:, i:i + h, j:j + wd].reshape(-1, cin).dot(w[i, j]).reshape(b, h, wd, cout)
        return y

    # ---- control flow (one non-nested frame per while_loop)
    def _op_Enter(self, n, it):
        src, idx = self._data_inputs(n)[0]
        return self._eval(src, idx, None)       # loop-invariant or iteration-0 value, from outside

    def _op_Merge(self, n, it):
        (e, ei), (x, xi) = self._data_inputs(n)
        if self.nodes[e].op != "Enter":
            (e, ei), (x, xi) = (x, xi), (e, ei)
        assert self.nodes[e].op == "Enter" and self.nodes[x].op == "NextIteration"
        if it == 0:
            return self._eval(e, ei, 0), np.int32(0)
        return self._eval(x, xi, it - 1), np.int32(1)

    def _op_NextIteration(self, n, it):
        return self._in(n, 0, it)

    def _op_LoopCond(self, n, it):
        return self._in(n, 0, it)

    def _op_Switch(self, n, it):
        data, pred = self._ins(n, it)
        return (_DEAD, data) if bool(pred) else (data, _DEAD)

    def _op_Exit(self, n, it):
        src, idx = self._data_inputs(n)[0]
        frame = self.frame[src]
        conds = [m for m in self.frame_nodes[frame] if self.nodes[m].op == "LoopCond"]
        nexts = [m for m in self.frame_nodes[frame] if self.nodes[m].op == "NextIteration"]
        assert len(conds) == 1
        i = 0
        while bool(self._eval(conds[0], 0, i)):
            for m in nexts:                      # completes iteration i (bounds the recursion depth)
                self._eval(m, 0, i)
            i += 1
            assert i < 100000
        self.iterations = getattr(self, "iterations", {})
        self.iterations[frame] = i
        return self._eval(src, idx, i)

    # ---- TensorArray (contents carried by the flow value)
    def _op_TensorArrayV3(self, n, it):
        size = int(self._in(n, 0, it))
        return n.name, _Flow(size)

    def _op_TensorArrayScatterV3(self, n, it):
        handle, indices, value, flow = self._ins(n, it)
        for k, i in enumerate(indices):
            flow = flow.write(i, value[k])
        return flow

    def _op_TensorArrayWriteV3(self, n, it):
        handle, index, value, flow = self._ins(n, it)
        assert 0 <= int(index) < flow.size
        return flow.write(index, value)

    def _op_TensorArrayReadV3(self, n, it):
        handle, index, flow = self._ins(n, it)
        return flow.items[int(index)]

    def _op_TensorArrayGatherV3(self, n, it):
        handle, indices, flow = self._ins(n, it)
        return np.stack([flow.items[int(i)] for i in indices])

    def _op_TensorArraySizeV3(self, n, it):
        handle, flow = self._ins(n, it)
        return np.int32(flow.size)


class _Dead(object):
    def __repr__(self):
        return "<dead>"


_DEAD = _Dead()


def load_reference_variables():
    """The 74 inference variables of the shipped TF-V2 bundle, keyed by graph node name."""
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
    if root not in sys.path:
        sys.path.insert(0, root)
    from catfish_b200 import tf_checkpoint
    return tf_checkpoint.load_checkpoint(os.path.join(REF_CKPT_DIR, "ckpnt-30000"))


def predictions(x, float_dtype=np.float32, variables=None, seed=0, graph=None):
    """``sess.run(self.predictions, {x: input_x, p_dropout: 1.0})`` + flatten as
    ``RNN.infer`` does (rnn_class.py:213-219).  x: [B, 35, 1]."""
    g = graph or MetaGraph(META, variables or load_reference_variables(), float_dtype, seed)
    p = g.run(PREDICTIONS, {X_PLACEHOLDER: np.asarray(x).reshape(-1, 35, 1), DROPOUT_PLACEHOLDER: 1.0})
    return np.reshape(p, -1).astype(float), g


def accuracy_loss(x, y, float_dtype=np.float32, variables=None, graph=None):
    """``sess.run([self.accuracy, self.loss], {x, y, p_dropout: 1.0})`` of
    ``RNN.test_network`` (rnn_class.py:240-241).  x, y: [B, 35, 1]."""
    g = graph or MetaGraph(META, variables or load_reference_variables(), float_dtype)
    feed = {X_PLACEHOLDER: np.asarray(x).reshape(-1, 35, 1), Y_PLACEHOLDER: np.asarray(y).reshape(-1, 35, 1),
            DROPOUT_PLACEHOLDER: 1.0}
    acc = g.run(ACCURACY, feed)
    loss = g.run(LOSS, feed)
    return float(acc), float(loss), g


RESNET_OUT = "ResNet_layer_1/residual_block/Relu_3"                       # input of the recurrent stack
RNN_OUT = "recurrent_layer/stack_bidirectional_rnn/cell_2/concat"         # input of the dense head


def predictions_variant(kind, x, variables, float_dtype=np.float64):
    """The reference's model VARIANTS on the shipped graph's own sub-graphs.

    * "RNN" (``neural_network.py:17-18``: ``RNN`` used directly): x is fed where the graph feeds the
      residual stack's output into the recurrent stack; layer-0 kernels are [1+H, .] (H = 64, 3 layers,
      the shipped recurrent hyper-parameters, because the graph's zero-state consts carry H).
    * "ResNet" (``resnet_class.py:23`` commented out): the residual stack's output is fed where the
      graph feeds the recurrent stack's output into ``Reshape``/dense; dense kernel [32, 1].
    Only the wiring between the sub-graphs is ours; every op inside them is the graph's."""
    x = np.asarray(x).reshape(-1, 35, 1)
    g = MetaGraph(META, variables, float_dtype, check_shapes=False)
    if kind == "RNN":
        feed = {RESNET_OUT: x.astype(np.float32), DROPOUT_PLACEHOLDER: 1.0}
    elif kind == "ResNet":
        y = g.run(RESNET_OUT, {X_PLACEHOLDER: x, DROPOUT_PLACEHOLDER: 1.0})
        feed = {RNN_OUT: y, "Reshape/shape": np.array([-1, y.shape[2]], np.int32), DROPOUT_PLACEHOLDER: 1.0}
    else:
        raise ValueError(kind)
    return np.reshape(g.run(PREDICTIONS, feed), -1).astype(float)


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    x = rng.normal(size=(3, 35, 1))
    p, g = predictions(x, np.float64)
    print("TF", g.tf_version, "nodes", len(g.nodes), "subgraph", len(g.subgraph(PREDICTIONS)))
    print("executed", sum(g.executed.values()), "node-iterations;", len(g.executed), "op types; loop trips", g.iterations)
    print(p[:5])
