"""numpy stand-in for the TensorFlow-1.x / TFP ops the reference samplers call -- TEST INFRASTRUCTURE.

TensorFlow, TFP and TFC are not installed in this image, so the reference's
code/coded_importance_sampler.py and code/coded_greedy_sampler.py cannot be imported as they are.
`install()` registers three fake modules (`tensorflow`, `tensorflow.compat.v1`,
`tensorflow_probability`) that implement, lazily and in float32 numpy, exactly the graph-mode surface
those two files and code/misc.py touch: placeholders + `sess.run(feed_dict)`, `tf.while_loop`,
`tf.numpy_function`, `tf.map_fn`, slicing, the elementwise / reduction ops, `tfd.Normal.log_prob`,
`tfd.kl_divergence`, `tf.quantization`.  With it the UNMODIFIED reference sources run in this
container (tests/golden/gen_reference_goldens.py), which pins their control flow, seeds, inequalities,
code concatenation and return types.

What it does NOT pin: the arithmetic inside TensorFlow's own kernels.  Stated assumptions:
  * `tf.random.stateless_normal(shape=[K, G], seed=[s, 42])` is the injection seam: it returns rows
    0..K-1 of the spec stream `s` (oracle.default_normal_source, or whatever `set_normal_source` installed);
  * `tfd.Normal.log_prob` / `tfd.kl_divergence(Normal, Normal)` follow the TFP 0.7-era formulas
    (TF 1.14/1.15 contemporaries; no version is pinned by the reference);
  * reductions are numpy float32 pairwise sums; `tf.argmax` returns the first maximal index;
  * `Normal.sample()` (unseeded in the reference, CIS:150) returns the spec's seeded outlier draw;
  * quantize/dequantize are MIN_COMBINED quint16 as restated in oracle/cwq_oracle.c.
"""
import sys
import types

import numpy as np

_state = {"normal_source": None, "outlier_seed": 0}
_ctx = []   # stack of (feed, memo) of the Session.run calls in flight: loop bodies close over placeholders


def set_normal_source(fn):
    _state["normal_source"] = fn


def set_outlier_seed(seed):
    _state["outlier_seed"] = int(seed)


# ----------------------------------------------------------------------------------------------
# lazy graph
# ----------------------------------------------------------------------------------------------
class DType:
    def __init__(self, name, np_dtype):
        self.name, self.np = name, np_dtype

    def __repr__(self):
        return "<dtype: '{}'>".format(self.name)


float32 = DType("float32", np.float32)
float64 = DType("float64", np.float64)
int32 = DType("int32", np.int32)
int64 = DType("int64", np.int64)
string = DType("string", object)
quint16 = DType("quint16", np.uint16)
_NP2DT = {np.dtype(np.float32): float32, np.dtype(np.float64): float64, np.dtype(np.int32): int32,
          np.dtype(np.int64): int64, np.dtype(np.uint16): quint16}


class TensorShape:
    def __init__(self, dims, tensor=None):
        self.dims, self.tensor = dims, tensor

    def as_list(self):
        if self.tensor is not None:
            # static shape inference is not modelled: evaluate (works whenever no placeholder is involved)
            return list(np.shape(self.tensor.eval({}, {})))
        return list(self.dims)


class Tensor:
    """A node of the lazy graph; evaluated by Session.run with a feed dict."""

    def __init__(self, fn, inputs=(), dtype=None, static_shape=None):
        self.fn, self.inputs, self._dtype, self._static_shape = fn, tuple(inputs), dtype, static_shape

    # -- evaluation --
    def eval(self, feed, memo):
        key = id(self)
        if key in memo:
            return memo[key][1]
        if key in feed:
            v = feed[key]
        else:
            v = self.fn(*[_ev(i, feed, memo) for i in self.inputs])
        memo[key] = (self, v)   # the node is kept alive so a recycled id() can never alias a stale entry
        return v

    @property
    def dtype(self):
        return self._dtype

    @property
    def shape(self):
        if self._static_shape is not None:
            return TensorShape(self._static_shape)
        return TensorShape([None], tensor=self)

    def get_shape(self):
        return self.shape

    # -- operators --
    def __add__(self, o): return _binary(np.add, self, o)
    def __radd__(self, o): return _binary(np.add, o, self)
    def __sub__(self, o): return _binary(np.subtract, self, o)
    def __rsub__(self, o): return _binary(np.subtract, o, self)
    def __mul__(self, o): return _binary(np.multiply, self, o)
    def __rmul__(self, o): return _binary(np.multiply, o, self)
    def __truediv__(self, o): return _binary(np.divide, self, o)
    def __rtruediv__(self, o): return _binary(np.divide, o, self)
    def __neg__(self): return Tensor(lambda a: -a, [self], self._dtype)
    def __le__(self, o): return _binary(np.less_equal, self, o, keep=False)
    def __lt__(self, o): return _binary(np.less, self, o, keep=False)
    def __ge__(self, o): return _binary(np.greater_equal, self, o, keep=False)
    def __gt__(self, o): return _binary(np.greater, self, o, keep=False)
    __hash__ = object.__hash__
    __array_ufunc__ = None   # numpy arrays defer to the reflected operators above (ndarray * Tensor -> Tensor)

    def __getitem__(self, item):
        items = item if isinstance(item, tuple) else (item,)
        flat, spec = [], []
        for it in items:
            if isinstance(it, slice):
                parts = []
                for p in (it.start, it.stop, it.step):
                    parts.append(len(flat) if isinstance(p, Tensor) else None)
                    if isinstance(p, Tensor):
                        flat.append(p)
                spec.append(("slice", it, parts))
            elif isinstance(it, Tensor):
                spec.append(("tensor", len(flat)))
                flat.append(it)
            else:
                spec.append(("const", it))

        def fn(a, *vals):
            idx = []
            for s in spec:
                if s[0] == "slice":
                    sl, parts = s[1], s[2]
                    vs = [int(vals[p]) if p is not None else q for p, q in zip(parts, (sl.start, sl.stop, sl.step))]
                    idx.append(slice(*vs))
                elif s[0] == "tensor":
                    idx.append(int(vals[s[1]]))
                else:
                    idx.append(s[1])
            return np.asarray(a)[tuple(idx)]
        return Tensor(fn, [self] + flat, self._dtype)


def _ev(x, feed, memo):
    if isinstance(x, Tensor):
        return x.eval(feed, memo)
    if isinstance(x, (list, tuple)):
        return type(x)(_ev(i, feed, memo) for i in x)
    return x


def _const_value(x, like=None):
    """Python scalars adopt the dtype of the tensor operand (TF's convert_to_tensor with dtype hint)."""
    if isinstance(x, (bool, int, float, np.floating, np.integer)) and like is not None and like.dtype.kind in "fiu":
        if like.dtype.kind == "f" or isinstance(x, (int, np.integer)):
            return np.asarray(x, dtype=like.dtype)
    return np.asarray(x)


def _binary(op, a, b, keep=True):
    def fn(x, y):
        xa = x if isinstance(x, np.ndarray) else None
        ya = y if isinstance(y, np.ndarray) else None
        xv = np.asarray(x) if xa is not None or not np.isscalar(x) else None
        yv = np.asarray(y) if ya is not None or not np.isscalar(y) else None
        if xv is None and yv is not None:
            xv = _const_value(x, yv)
        elif yv is None and xv is not None:
            yv = _const_value(y, xv)
        elif xv is None and yv is None:
            xv, yv = np.asarray(x), np.asarray(y)
        return op(xv, yv)
    dt = a._dtype if isinstance(a, Tensor) else (b._dtype if isinstance(b, Tensor) else None)
    return Tensor(fn, [a, b], dt if keep else None)


def convert(x, dtype=None):
    if isinstance(x, Tensor):
        return x
    a = np.asarray(x, dtype=dtype.np if dtype is not None and dtype is not string else None)
    return Tensor(lambda: a, [], dtype or _NP2DT.get(a.dtype))


def _unary(f):
    return lambda x, name=None: Tensor(lambda a: f(np.asarray(a)), [x], getattr(x, "_dtype", None))


# ----------------------------------------------------------------------------------------------
# tf.* surface
# ----------------------------------------------------------------------------------------------
def placeholder(dtype, shape=None):
    def missing():
        raise RuntimeError("placeholder was not fed")
    return Tensor(missing, [], dtype, list(shape) if shape is not None else None)


def constant(v, dtype=None):
    return convert(v, dtype or (int32 if isinstance(v, int) else None))


def cast(x, dtype):
    if dtype is quint16:
        return Tensor(lambda a: np.asarray(a).astype(np.uint16), [x], quint16)
    return Tensor(lambda a: np.asarray(a).astype(dtype.np), [x], dtype)


def reduce_sum(x, axis=None):
    return Tensor(lambda a: np.sum(np.asarray(a), axis=axis, dtype=np.asarray(a).dtype), [x], getattr(x, "_dtype", None))


def reduce_prod(x, axis=None):
    return Tensor(lambda a: np.prod(np.asarray(a), axis=axis), [x], int32)


def shape(x):
    return Tensor(lambda a: np.asarray(np.shape(a), dtype=np.int32), [x], int32)


def rank(x):
    return Tensor(lambda a: np.int32(np.ndim(a)), [x], int32)


def zeros_like(x):
    return Tensor(lambda a: np.zeros_like(np.asarray(a)), [x], getattr(x, "_dtype", None))


def ones_like(x):
    return Tensor(lambda a: np.ones_like(np.asarray(a)), [x], getattr(x, "_dtype", None))


def zeros(shp, dtype=float32):
    return Tensor(lambda s: np.zeros(tuple(int(v) for v in np.atleast_1d(s)), dtype=dtype.np), [shp], dtype)


def where(cond, x=None, y=None):
    if x is None:
        return Tensor(lambda c: np.argwhere(np.asarray(c)).astype(np.int64), [cond], int64)
    return Tensor(lambda c, a, b: np.where(np.asarray(c), a, b), [cond, x, y], getattr(x, "_dtype", None))


def gather_nd(params, indices):
    return Tensor(lambda p, i: np.asarray(p)[tuple(np.asarray(i).T)], [params, indices], getattr(params, "_dtype", None))


def reshape(x, shp):
    return Tensor(lambda a, s: np.reshape(np.asarray(a), tuple(int(v) for v in np.atleast_1d(s))), [x, shp],
                  getattr(x, "_dtype", None))


def concat(values, axis):
    vals = list(values)
    return Tensor(lambda *a: np.concatenate([np.atleast_1d(np.asarray(v)) for v in a], axis=axis), vals,
                  next((v._dtype for v in vals if isinstance(v, Tensor)), None))


def squeeze(x):
    return Tensor(lambda a: np.squeeze(np.asarray(a)), [x], getattr(x, "_dtype", None), getattr(x, "_static_shape", None))


def tile(x, multiples):
    return Tensor(lambda a, m: np.tile(np.asarray(a), tuple(int(v) for v in np.atleast_1d(m))), [x, multiples],
                  getattr(x, "_dtype", None))


def expand_dims(x, axis):
    return Tensor(lambda a: np.expand_dims(np.asarray(a), axis), [x], getattr(x, "_dtype", None))


def argmax(x, axis=0):
    return Tensor(lambda a: np.int64(np.argmax(np.asarray(a), axis=axis)), [x], int64)


def equal(a, b):
    return _binary(np.equal, a, b, keep=False)


def scatter_nd(indices, updates, shp):
    def fn(i, u, s):
        out = np.zeros(tuple(int(v) for v in s), dtype=np.asarray(u).dtype)
        np.add.at(out, tuple(np.asarray(i).T), np.asarray(u))
        return out
    return Tensor(fn, [indices, updates, shp], getattr(updates, "_dtype", None))


def _to_tf_string(v):
    if isinstance(v, str):
        return v.encode("utf-8")
    return v


def numpy_function(func, inp, Tout):
    def fn(*args):
        conv = []
        for a in args:
            if isinstance(a, (list, tuple)):
                a = np.array(list(a), dtype=object)
            elif isinstance(a, np.ndarray):
                a = a.copy()   # py functions get their own buffers (binary_io.to_bit_string divides its argument in place)
            conv.append(a)
        r = func(*conv)
        if isinstance(Tout, (list, tuple)):
            return tuple(np.asarray(x, dtype=t.np) if t is not string else _to_tf_string(x) for x, t in zip(r, Tout))
        if Tout is string:
            return _to_tf_string(r)
        return np.asarray(r, dtype=Tout.np)
    node = Tensor(fn, list(inp), Tout if not isinstance(Tout, (list, tuple)) else None)
    if isinstance(Tout, (list, tuple)):
        return tuple(Tensor(lambda r, j=j: r[j], [node], t) for j, t in enumerate(Tout))
    return node


def _sub_eval(node):
    feed, memo = _ctx[-1] if _ctx else ({}, {})
    return _ev(node, feed, memo)


def map_fn(fn, elems, dtype=None):
    def run(e):
        out = []
        for x in np.asarray(e):
            out.append(_sub_eval(fn(convert(x))))
        return np.array(out, dtype=object)
    return Tensor(run, [elems], dtype)


def while_loop(cond, body, loop_vars, shape_invariants=None):
    n = len(loop_vars)

    def run(*vals):
        cur = list(vals)
        while True:
            nodes = [convert(v) for v in cur]
            if not bool(_sub_eval(cond(*nodes))):
                break
            res = body(*nodes)
            cur = [_sub_eval(r) for r in res]
        return tuple(cur)
    node = Tensor(run, list(loop_vars))
    return [Tensor(lambda r, j=j: r[j], [node], getattr(loop_vars[j], "_dtype", None)) for j in range(n)]


class _Substr:
    @staticmethod
    def substr(s, pos, length):
        return Tensor(lambda a, p, l: (a if isinstance(a, bytes) else str(a).encode("utf-8"))[int(p):int(p) + int(l)],
                      [s, pos, length], string)


def stateless_normal(shape, seed):
    def fn(shp, sd):
        shp = [int(v) for v in np.atleast_1d(shp)]
        sd = [int(v) for v in np.atleast_1d(sd)]
        assert len(sd) == 2 and sd[1] == 42, "the reference always seeds with [seed, 42] (misc.py:11)"
        K = shp[0]
        G = int(np.prod(shp[1:])) if len(shp) > 1 else 1
        z = _state["normal_source"](sd[0] & 0xFFFFFFFF, 0, K, G)
        return np.asarray(z, dtype=np.float32).reshape(shp)
    return Tensor(fn, [shape, seed], float32)


class Session:
    def run(self, fetches, feed_dict=None):
        feed = {id(k): (np.asarray(v) if not isinstance(v, (bytes, str)) else v) for k, v in (feed_dict or {}).items()}
        memo = {}
        _ctx.append((feed, memo))
        try:
            return _ev(fetches, feed, memo)
        finally:
            _ctx.pop()


# ----------------------------------------------------------------------------------------------
# tfp.distributions surface (TFP 0.7-era formulas, float32)
# ----------------------------------------------------------------------------------------------
class Normal:
    def __init__(self, loc, scale):
        self.loc, self.scale = loc, scale
        ld = loc._dtype if isinstance(loc, Tensor) else _NP2DT.get(np.asarray(loc).dtype)
        sd = scale._dtype if isinstance(scale, Tensor) else _NP2DT.get(np.asarray(scale).dtype)
        self.dtype = ld if ld is sd else (ld if sd is None else (sd if ld is None else None))

    def log_prob(self, x):
        def fn(v, loc, scale):
            f = np.float32
            z = (np.asarray(v, f) - np.asarray(loc, f)) / np.asarray(scale, f)
            return (f(-0.5) * np.square(z) - (f(0.5 * np.log(2.0 * np.pi)) + np.log(np.asarray(scale, f)))).astype(f)
        return Tensor(fn, [x, self.loc, self.scale], float32)

    def sample(self, seed=None):
        def fn(loc, scale):
            loc, scale = np.asarray(loc, np.float32), np.asarray(scale, np.float32)
            n = loc.size
            z = np.array([_state["normal_source"]((_state["outlier_seed"] + 0x9E3779B9) & 0xFFFFFFFF, i, 1, 1)[0, 0]
                          for i in range(n)], dtype=np.float32).reshape(loc.shape)
            return (scale * z + loc).astype(np.float32)
        return Tensor(fn, [self.loc, self.scale], float32)


def kl_divergence(a, b):
    def fn(al, asc, bl, bsc):
        f = np.float32
        al, asc, bl, bsc = (np.asarray(v, f) for v in (al, asc, bl, bsc))
        s_a2, s_b2 = np.square(asc), np.square(bsc)
        ratio = s_a2 / s_b2
        return (np.square(al - bl) / (f(2) * s_b2) + f(0.5) * (ratio - f(1) - np.log(ratio))).astype(f)
    return Tensor(fn, [a.loc, a.scale, b.loc, b.scale], float32)


class _QuantizeOut:
    def __init__(self, t):
        self.output = t


def _quantize(x, lo, hi, T):
    from oracle import oracle as O
    assert (lo, hi) == (-30, 30) and T is quint16
    return _QuantizeOut(Tensor(lambda a: O.quantize_u16(np.asarray(a, np.float32)).reshape(np.shape(a)), [x], quint16))


def _dequantize(x, lo, hi):
    from oracle import oracle as O
    assert (lo, hi) == (-30, 30)
    return Tensor(lambda a: O.dequantize_u16(np.asarray(a)).reshape(np.shape(a)), [x], float32)


def install():
    """Register the fake modules; returns the `tensorflow.compat.v1` stand-in."""
    tf1 = types.ModuleType("tensorflow.compat.v1")
    for name in ("float32", "float64", "int32", "int64", "string", "quint16", "placeholder", "constant", "cast", "reduce_sum",
                 "reduce_prod", "shape", "rank", "zeros_like", "ones_like", "zeros", "where", "gather_nd", "reshape", "concat",
                 "squeeze", "tile", "expand_dims", "argmax", "equal", "scatter_nd", "numpy_function", "map_fn", "while_loop",
                 "TensorShape", "Session"):
        setattr(tf1, name, globals()[name])
    tf1.exp = _unary(np.exp)
    tf1.math = types.SimpleNamespace(ceil=_unary(np.ceil), log=_unary(np.log))
    tf1.strings = _Substr
    tf1.random = types.SimpleNamespace(stateless_normal=stateless_normal)
    tf1.quantization = types.SimpleNamespace(quantize=_quantize, dequantize=_dequantize)
    tf = types.ModuleType("tensorflow")
    compat = types.ModuleType("tensorflow.compat")
    compat.v1 = tf1
    tf.compat = compat
    tfp = types.ModuleType("tensorflow_probability")
    tfp.distributions = types.SimpleNamespace(Normal=Normal, kl_divergence=kl_divergence)
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.compat"] = compat
    sys.modules["tensorflow.compat.v1"] = tf1
    sys.modules["tensorflow_probability"] = tfp
    return tf1
