"""A numpy stand-in for the slice of jax / optax / flax that the reference's hot-path SOURCE FILES touch, so that
those files can be imported from /root/reference and EXECUTED in this container (jax is not installed and there is no
network).  TEST INFRASTRUCTURE: only tests/golden/make_ref_vectors.py uses it, to write tests/golden/ref_vectors.npz;
nothing under fortuna_b200/ may import it.

What a run through this stand-in pins, and what it does not:

* pinned: the reference's own control flow, formulas, argument plumbing, key schedule (which key goes where, how many
  splits), reduction structure, dtype promotions (re-derived here by RULE from jax's promotion lattice with x64 off,
  not by hand per expression as in oracle/), broadcasting and indexing - everything that is written in the
  reference's .py files.
* not pinned: the numerics of jax PRIMITIVES.  jnp.sum / mean use numpy's pairwise order instead of XLA's; exp / log /
  cos are libm's; jnp.cumsum is bound to the XLA-order scan restated in oracle/scoring.py (flagged there as a
  restatement); jnp.linspace and jnp.quantile are restated below from jax's library source (their formulas, not
  numpy's: numpy's linspace differs from jax's in the last bit of interior points); jax.random is bound to the threefry restatement in oracle/prng.py, which IS pinned by the Random123
  and jax known answers in tests/golden/kat.json.  jax.nn.softmax / one_hot and jax.scipy.special.logsumexp are
  restated from jax's library source (jax/_src/nn/functions.py, jax/_src/ops/special.py).

Arrays are ``JArray`` (an ndarray subclass): every ufunc call first casts its operands to the dtype jax would compute in
(python scalars are weak; float64 -> float32, int64 -> int32), then calls numpy, so arithmetic happens IN float32 and
not in float64-then-rounded.
"""

from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types
from typing import Any, NamedTuple

import numpy as np

REFERENCE_ROOT = "/root/reference"

_CANON = {
    np.dtype(np.float64): np.dtype(np.float32),
    np.dtype(np.int64): np.dtype(np.int32),
    np.dtype(np.uint64): np.dtype(np.uint32),
    np.dtype(np.complex128): np.dtype(np.complex64),
}


def _canon(dt):
    dt = np.dtype(dt)
    return _CANON.get(dt, dt)


_CAT = {"b": 0, "i": 1, "u": 1, "f": 2, "c": 3}
_DEFAULT = {0: np.dtype(np.bool_), 1: np.dtype(np.int32), 2: np.dtype(np.float32), 3: np.dtype(np.complex64)}
# ufuncs whose result is floating even for integer operands (jnp promotes ints to the default float first)
_FLOAT_ONLY = {
    np.sqrt, np.exp, np.exp2, np.log, np.log2, np.log10, np.log1p, np.expm1, np.sin, np.cos, np.tan, np.tanh,
    np.arctan, np.true_divide, np.ceil, np.floor, np.reciprocal, np.cbrt,
}  # fmt: skip
_NO_UNIFY = {np.left_shift, np.right_shift, np.ldexp, np.isnan, np.isinf, np.isfinite, np.logical_not, np.signbit}


def _is_weak(x):
    return isinstance(x, (bool, int, float, complex)) and not isinstance(x, np.generic)


def _weak_cat(x):
    return 0 if isinstance(x, bool) else 1 if isinstance(x, int) else 2 if isinstance(x, float) else 3


def _result_dtype(inputs, float_only=False):
    """jax's promotion with x64 disabled, for the cases the reference produces: the category is the highest one present;
    the width comes from the non-weak operands of that category (after canonicalisation), else the default of it."""
    cat, strong = -1, []
    for x in inputs:
        if _is_weak(x):
            cat = max(cat, _weak_cat(x))
        else:
            dt = _canon(np.asarray(x).dtype)
            c = _CAT[dt.kind]
            cat = max(cat, c)
            strong.append((c, dt))
    if float_only and cat < 2:
        return _DEFAULT[2]
    top = [dt for c, dt in strong if c == cat]
    if not top:
        return _DEFAULT[cat]
    return _canon(np.result_type(*top))


class JArray(np.ndarray):
    __array_priority__ = 1000

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kw):
        raw = [np.asarray(x) if isinstance(x, JArray) else x for x in inputs]
        if ufunc in _NO_UNIFY or method not in ("__call__", "reduce", "accumulate", "outer"):
            args = [np.asarray(x) if not _is_weak(x) else x for x in raw]
        else:
            dt = _result_dtype(raw, float_only=ufunc in _FLOAT_ONLY)
            args = [np.asarray(x, dtype=dt) for x in raw]
            if method == "reduce" and kw.get("dtype") is not None:
                kw["dtype"] = _canon(kw["dtype"])
        if out is not None:
            kw["out"] = tuple(np.asarray(o) if isinstance(o, JArray) else o for o in out)
        res = getattr(ufunc, method)(*args, **kw)
        if out is not None:
            return out[0] if len(out) == 1 else out
        if isinstance(res, tuple):
            return tuple(_j(r) for r in res)
        return _j(res)

    def __array_function__(self, func, types_, args, kwargs):
        # numpy's own implementation runs on the JArrays: the ufuncs inside it (sum -> add.reduce, mean -> add.reduce and
        # true_divide, ...) come back through __array_ufunc__ and follow the promotion rule.
        res = super().__array_function__(func, types_, args, kwargs)
        return _wrap_out(res)

    # jax arrays are immutable; .at[idx].set / add return copies
    @property
    def at(self):
        return _At(self)

    def astype(self, dtype, *a, **k):
        return _j(np.asarray(self).astype(_canon(np.dtype(dtype)), *a, **k))

    def tolist(self):
        return np.asarray(self).tolist()

    def block_until_ready(self):
        return self

    def __hash__(self):
        return id(self)

    def __iter__(self):
        if self.ndim == 0:
            raise TypeError("iteration over a 0-d array")
        return (self[i] for i in range(self.shape[0]))

    def __getitem__(self, idx):
        return _j(np.asarray(self)[_plain(idx)])

    def __bool__(self):
        return bool(np.asarray(self))

    def __index__(self):
        return int(np.asarray(self))

    def __int__(self):
        return int(np.asarray(self))

    def __float__(self):
        return float(np.asarray(self))

    def __format__(self, spec):
        return format(np.asarray(self)[()] if self.ndim == 0 else np.asarray(self), spec)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, _plain(idx)

    def set(self, v):
        a = np.array(self.arr)
        a[self.idx] = np.asarray(v, dtype=a.dtype)
        return _j(a)

    def add(self, v):
        a = np.array(self.arr)
        np.add.at(a, self.idx, np.asarray(v, dtype=a.dtype))
        return _j(a)


def _plain(idx):
    if isinstance(idx, JArray):
        return np.asarray(idx)
    if isinstance(idx, tuple):
        return tuple(_plain(i) for i in idx)
    return idx


def _j(x) -> JArray:
    """Anything array-like -> canonical-dtype JArray (0-d stays an array so that later arithmetic still dispatches)."""
    if isinstance(x, JArray) and x.dtype == _canon(x.dtype):
        return x
    a = np.asarray(x)
    if a.dtype == object:
        return x
    dt = _canon(a.dtype)
    if a.dtype != dt:
        a = a.astype(dt)
    return a.view(JArray)


def _wrap_out(res):
    if isinstance(res, np.ndarray) or isinstance(res, np.generic):
        return _j(res)
    if isinstance(res, tuple):
        return tuple(_wrap_out(r) for r in res)
    if isinstance(res, list):
        return [_wrap_out(r) for r in res]
    return res


def _wrap_in(a):
    if isinstance(a, JArray) or a is None or isinstance(a, (str, np.dtype, type, slice)) or callable(a):
        return a
    if _is_weak(a):
        return a
    if isinstance(a, (np.ndarray, np.generic)):
        return _j(a)
    if isinstance(a, (list, tuple)):
        if len(a) and all(_is_weak(b) for b in a):
            return a  # shapes, axes, python lists of numbers: leave for numpy / jnp.array to interpret
        return type(a)(_wrap_in(b) for b in a) if not hasattr(a, "_fields") else type(a)(*[_wrap_in(b) for b in a])
    return a


def _np_fn(fn, float_out=False):
    """jnp.<name> from np.<name>: operands become JArrays (so the ufuncs inside follow the promotion rule above) and
    results are canonicalised."""

    def wrapped(*args, **kwargs):
        args = [_wrap_in(a) for a in args]
        if args and _is_weak(args[0]):
            args[0] = _j(np.asarray(args[0]))
        if "dtype" in kwargs and kwargs["dtype"] is not None:
            kwargs["dtype"] = _canon(np.dtype(kwargs["dtype"]))
        res = _wrap_out(fn(*args, **kwargs))
        return res

    wrapped.__name__ = getattr(fn, "__name__", "fn")
    return wrapped


# ---------------------------------------------------------------------------------------------------------------------
# pytrees (jax.tree_util): dict keys sorted, None is an empty node, NamedTuples by field order
# ---------------------------------------------------------------------------------------------------------------------
class _Leaf:
    pass


_LEAF = _Leaf()


class TreeDef:
    def __init__(self, skeleton, num_leaves):
        self.skeleton, self.num_leaves = skeleton, num_leaves

    def __eq__(self, other):
        return isinstance(other, TreeDef) and repr(self.skeleton) == repr(other.skeleton)

    def __repr__(self):
        return f"TreeDef({self.skeleton!r})"


def _flatten(tree, leaves, is_leaf=None):
    if is_leaf is not None and is_leaf(tree):
        leaves.append(tree)
        return _LEAF
    if tree is None:
        return None
    if isinstance(tree, dict):
        return ("dict", type(tree), [(k, _flatten(tree[k], leaves, is_leaf)) for k in sorted(tree)])
    if isinstance(tree, tuple) and hasattr(tree, "_fields"):
        return ("namedtuple", type(tree), [_flatten(v, leaves, is_leaf) for v in tree])
    if isinstance(tree, (list, tuple)):
        return ("seq", type(tree), [_flatten(v, leaves, is_leaf) for v in tree])
    leaves.append(tree)
    return _LEAF


def _unflatten(skel, it):
    if skel is _LEAF:
        return next(it)
    if skel is None:
        return None
    kind, typ, children = skel
    if kind == "dict":
        return typ({k: _unflatten(c, it) for k, c in children})
    if kind == "namedtuple":
        return typ(*[_unflatten(c, it) for c in children])
    return typ(_unflatten(c, it) for c in children)


def tree_flatten(tree, is_leaf=None):
    leaves = []
    skel = _flatten(tree, leaves, is_leaf)
    return leaves, TreeDef(skel, len(leaves))


def tree_unflatten(treedef, leaves):
    return _unflatten(treedef.skeleton, iter(leaves))


def tree_leaves(tree, is_leaf=None):
    return tree_flatten(tree, is_leaf)[0]


def tree_structure(tree):
    return tree_flatten(tree)[1]


def tree_map(f, tree, *rest, is_leaf=None):
    leaves, treedef = tree_flatten(tree, is_leaf)
    others = [treedef_flatten_up_to(treedef, r) for r in rest]
    return tree_unflatten(treedef, [f(*xs) for xs in zip(leaves, *others)])


def treedef_flatten_up_to(treedef, tree):
    out = []

    def walk(skel, t):
        if skel is _LEAF:
            out.append(t)
        elif skel is None:
            return
        else:
            kind, _, children = skel
            if kind == "dict":
                for k, c in children:
                    walk(c, t[k])
            else:
                for c, v in zip(children, t):
                    walk(c, v)

    walk(treedef.skeleton, tree)
    return out


def tree_reduce(f, tree, initializer=None):
    leaves = tree_leaves(tree)
    acc = initializer
    start = 0
    if acc is None:
        acc, start = leaves[0], 1
    for x in leaves[start:]:
        acc = f(acc, x)
    return acc


def ravel_pytree(tree):
    leaves, treedef = tree_flatten(tree)
    shapes = [np.shape(x) for x in leaves]
    sizes = [int(np.prod(s)) for s in shapes]
    flat = _j(np.concatenate([np.ravel(np.asarray(x)) for x in leaves])) if leaves else _j(np.zeros(0, np.float32))

    def unravel(v):
        v = np.asarray(v)
        out, o = [], 0
        for s, n in zip(shapes, sizes):
            out.append(_j(v[o : o + n].reshape(s)))
            o += n
        return tree_unflatten(treedef, out)

    return flat, unravel


# ---------------------------------------------------------------------------------------------------------------------
# transformations: batching by explicit loops (the fixtures are small)
# ---------------------------------------------------------------------------------------------------------------------
def vmap(fun=None, in_axes=0, out_axes=0):
    if fun is None:
        return lambda f: vmap(f, in_axes, out_axes)

    def batched(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = tree_leaves(a)[0].shape[ax]
                break
        outs = []
        for i in range(n):
            sl = [a if ax is None else tree_map(lambda v, ax=ax: _j(np.take(np.asarray(v), i, axis=ax)), a)
                  for a, ax in zip(args, axes)]  # fmt: skip
            outs.append(fun(*sl))
        return tree_map(lambda *xs: _j(np.stack([np.asarray(x) for x in xs], axis=out_axes)), outs[0], *outs[1:])

    return batched


def _identity_transform(fun=None, **_):
    if fun is None:
        return lambda f: f
    return fun


def _pmap(*_a, **_k):
    raise NotImplementedError("the stand-in runs the reference with distribute=False")


class _Lax(types.ModuleType):
    @staticmethod
    def cond(pred, true_fun, false_fun, *operands):
        return true_fun(*operands) if bool(pred) else false_fun(*operands)

    @staticmethod
    def map(f, xs):
        n = tree_leaves(xs)[0].shape[0]
        outs = [f(tree_map(lambda v: v[i], xs)) for i in range(n)]
        return tree_map(lambda *ys: _j(np.stack([np.asarray(y) for y in ys], axis=0)), outs[0], *outs[1:])

    @staticmethod
    def scan(f, init, xs, length=None):
        n = tree_leaves(xs)[0].shape[0] if xs is not None else length
        carry, ys = init, []
        for i in range(n):
            carry, y = f(carry, tree_map(lambda v: v[i], xs) if xs is not None else None)
            ys.append(y)
        if ys and ys[0] is not None:
            ys = tree_map(lambda *zs: _j(np.stack([np.asarray(z) for z in zs], axis=0)), ys[0], *ys[1:])
        else:
            ys = None
        return carry, ys

    @staticmethod
    def fori_loop(lower, upper, body, init):
        val = init
        for i in range(int(lower), int(upper)):
            val = body(i, val)
        return val

    @staticmethod
    def stop_gradient(x):
        return x

    @staticmethod
    def select(pred, on_true, on_false):
        return _j(np.where(np.asarray(pred), np.asarray(on_true), np.asarray(on_false)))


# ---------------------------------------------------------------------------------------------------------------------
# module construction
# ---------------------------------------------------------------------------------------------------------------------
def _make_jnp():
    m = types.ModuleType("jax.numpy")
    for name in (
        "sum mean max min argmax argmin argsort sort where concatenate stack take_along_axis linspace arange zeros ones "
        "zeros_like ones_like full sqrt exp log log2 cos sin abs square minimum maximum clip dot transpose roll pad real "
        "conj ceil floor unique nan_to_num split swapaxes expand_dims squeeze prod all any cumprod matmul einsum "
        "isnan var std tile repeat reshape ravel broadcast_to searchsorted outer diff sign power logical_and logical_or "
        "logical_not allclose array_equal moveaxis take median quantile percentile count_nonzero eye diag tril triu "
        "log1p expm1 tanh atleast_1d atleast_2d hstack vstack dstack vsplit hsplit flip isfinite argwhere nonzero round floor_divide mod "
        "equal not_equal less greater less_equal greater_equal add subtract multiply divide negative"
    ).split():
        setattr(m, name, _np_fn(getattr(np, name)))

    def array(x, dtype=None, copy=True):
        if isinstance(x, (list, tuple)):
            x = [np.asarray(v) if isinstance(v, JArray) else v for v in x]
        a = np.array(x, dtype=None if dtype is None else _canon(np.dtype(dtype)))
        return _j(a)

    m.array = array
    m.shape, m.ndim, m.size = np.shape, np.ndim, np.size
    m.copy = lambda x: _j(np.array(x))
    m.unravel_index = lambda i, shape: tuple(_j(v) for v in np.unravel_index(int(i), shape))
    m.asarray = lambda x, dtype=None: array(x, dtype)

    def cumsum(x, axis=None, dtype=None):
        from oracle import scoring as _sc

        a = np.asarray(x)
        if a.dtype.kind != "f":
            return _j(np.cumsum(a, axis=axis))
        if axis is None:
            a, axis = a.ravel(), 0
        moved = np.moveaxis(a.astype(np.float32), axis, -1)
        out = _sc.cumsum_f32(moved)  # XLA-order scan, restated in the oracle (see the header)
        return _j(np.moveaxis(out, -1, axis))

    m.cumsum = cumsum

    def quantile(a, q, axis=None, **kw):
        # jnp.quantile, method="linear": low + (high - low) * frac evaluated in float32 on the sorted values
        a = np.sort(np.asarray(a, dtype=np.float32), axis=0 if axis is None else axis)
        if axis is None:
            a, axis = a.ravel(), 0
            a = np.sort(a)
        qa = np.asarray(q, dtype=np.float32)
        n = a.shape[axis]
        pos = qa * np.float32(n - 1)
        low = np.floor(pos)
        high = np.ceil(pos)
        hw = (pos - low).astype(np.float32)
        lw = (np.float32(1) - hw).astype(np.float32)
        lo = np.clip(low, 0, n - 1).astype(np.int32)
        hi = np.clip(high, 0, n - 1).astype(np.int32)
        am = np.moveaxis(a, axis, 0)
        if qa.ndim == 0:
            return _j(am[lo] * lw + am[hi] * hw)
        sh = (-1,) + (1,) * (am.ndim - 1)
        return _j(am[lo] * lw.reshape(sh) + am[hi] * hw.reshape(sh))

    m.quantile = quantile

    def linspace(start, stop, num=50, endpoint=True, dtype=None):
        # jnp.linspace is a LIBRARY function (jax/_src/numpy/lax_numpy.py), not numpy's formula: in the computation dtype
        # (float32), step = iota(div) / div, out = start * (1 - step) + stop * step, and the endpoint is appended as is.
        assert endpoint and num > 1
        f = np.float32
        start, stop, div = f(start), f(stop), num - 1
        step = (np.arange(div, dtype=f) / f(div)).astype(f)
        out = (start * (f(1) - step) + stop * step).astype(f)
        return _j(np.concatenate([out, np.array([stop], dtype=f)]))

    m.linspace = linspace

    def clip(a, a_min=None, a_max=None, **kw):
        a = _wrap_in(a) if not _is_weak(a) else _j(np.asarray(a))
        if a_min is not None:
            a = np.maximum(a, a_min)
        if a_max is not None:
            a = np.minimum(a, a_max)
        return _j(a)

    m.clip = clip
    m.pi, m.inf, m.nan, m.e, m.newaxis = np.pi, np.inf, np.nan, np.e, None
    m.ndarray = JArray
    for t in ("float32", "float64", "int32", "int64", "uint32", "uint8", "int8", "bool_", "complex64", "float16"):
        setattr(m, t, getattr(np, t))
    fft = types.ModuleType("jax.numpy.fft")
    fft.fft = lambda x, *a, **k: _j(np.fft.fft(np.asarray(x), *a, **k).astype(np.complex64))
    fft.ifft = lambda x, *a, **k: _j(np.fft.ifft(np.asarray(x), *a, **k).astype(np.complex64))
    m.fft = fft
    linalg = types.ModuleType("jax.numpy.linalg")
    linalg.norm = _np_fn(np.linalg.norm)
    m.linalg = linalg
    return m


def _make_random():
    from oracle import prng as P

    m = types.ModuleType("jax.random")
    m.PRNGKey = lambda seed: _j(P.PRNGKey(int(seed)))
    m.key = m.PRNGKey
    m.split = lambda key, num=2: _j(P.split(np.asarray(key), int(num)))
    m.fold_in = lambda key, data: _j(P.fold_in(np.asarray(key), int(data)))

    def normal(key, shape=(), dtype=np.float32):
        n = int(np.prod(shape)) if len(tuple(shape)) else 1
        return _j(P.normal(np.asarray(key), n).reshape(tuple(shape)).astype(np.float32))

    def uniform(key, shape=(), dtype=np.float32, minval=0.0, maxval=1.0):
        n = int(np.prod(shape)) if len(tuple(shape)) else 1
        return _j(P.uniform(np.asarray(key), n, minval, maxval).reshape(tuple(shape)))

    def choice(key, a, shape=(), replace=True, p=None, axis=0):
        if p is None and not replace and tuple(shape) == (int(a),):
            # choice(key, n, (n,), replace=False) == permutation(key, n) (jax/_src/random.py): the sort-based shuffle
            # restated in oracle/multivalid.py (a jax LIBRARY algorithm, flagged there as a restatement)
            from oracle import multivalid as _mv

            return _j(_mv.permutation_from_key(np.asarray(key), int(a)))
        if p is not None and replace:
            # choice with probabilities and replacement (jax/_src/random.py): p_cuml = cumsum(p);
            # r = p_cuml[-1] * (1 - uniform(key, shape)); searchsorted(p_cuml, r) - a jax LIBRARY algorithm, restated
            from oracle.scoring import cumsum_f32

            n_draw = int(np.prod(shape)) if len(tuple(shape)) else 1
            p_cuml = cumsum_f32(np.asarray(p, dtype=np.float32))
            u = P.uniform(np.asarray(key), n_draw).astype(np.float32)
            r = (p_cuml[-1] * (np.float32(1) - u)).astype(np.float32)
            return _j(np.searchsorted(p_cuml, r, side="left").astype(np.int32).reshape(tuple(shape)))
        if p is not None or tuple(shape) != ():
            raise NotImplementedError("stand-in: the scalar uniform random.choice(key, n) of SGMCMCPosterior.sample, the "
                                      "full permutation of the multivalid split and choice(p=) with replacement only")
        return _j(np.int32(P.choice(np.asarray(key), int(a))))

    m.normal, m.uniform, m.choice = normal, uniform, choice
    return m


def _make_nn_and_special():
    nn = types.ModuleType("jax.nn")

    def softmax(x, axis=-1):
        x = np.asarray(x, dtype=np.float32)
        un = np.exp(x - x.max(axis, keepdims=True))
        return _j(un / un.sum(axis, keepdims=True))

    def log_softmax(x, axis=-1):
        x = np.asarray(x, dtype=np.float32)
        sh = x - x.max(axis, keepdims=True)
        return _j(sh - np.log(np.exp(sh).sum(axis, keepdims=True)))

    def one_hot(x, num_classes, dtype=np.float32, axis=-1):
        x = np.asarray(x)
        return _j((x[..., None] == np.arange(num_classes, dtype=_canon(x.dtype))).astype(_canon(np.dtype(dtype))))

    nn.softmax, nn.log_softmax, nn.one_hot = softmax, log_softmax, one_hot
    nn.relu = lambda x: _j(np.maximum(np.asarray(x), 0))
    jsp = types.ModuleType("jax.scipy")
    special = types.ModuleType("jax.scipy.special")

    def logsumexp(a, axis=None, b=None, keepdims=False):
        a = np.asarray(a, dtype=np.float32)
        amax = a.max(axis=axis, keepdims=True)
        amax = np.where(np.isfinite(amax), amax, np.float32(0))
        out = np.log(np.exp(a - amax).sum(axis=axis, keepdims=keepdims))
        return _j(out + (amax if keepdims else np.squeeze(amax, axis=axis)))

    special.logsumexp = logsumexp
    jsp.special = special
    return nn, jsp, special


class _AutoStub(types.ModuleType):
    """A module whose every attribute is an empty class of that name (annotations, base classes, isinstance checks);
    instances keep their keyword arguments as attributes (JointState(params=..., mutable=...))."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (), {"__init__": lambda self, *a, **k: self.__dict__.update(k)})
        setattr(self, name, cls)
        return cls


def _make_optax():
    optax = types.ModuleType("optax")

    class GradientTransformation(NamedTuple):
        init: Any
        update: Any

    class EmptyState(NamedTuple):
        pass

    def sgd(learning_rate, momentum=None, nesterov=False):
        # optax.sgd without momentum == scale_by_learning_rate: updates = -learning_rate * g
        assert momentum is None

        def init(params):
            return (EmptyState(), EmptyState())

        def update(updates, state, params=None):
            return tree_map(lambda g: -learning_rate * g, updates), state

        return GradientTransformation(init, update)

    def apply_updates(params, updates):
        return tree_map(lambda p, u: _j(np.asarray(p + u).astype(np.asarray(p).dtype)), params, updates)

    optax.GradientTransformation, optax.EmptyState, optax.sgd, optax.apply_updates = (
        GradientTransformation, EmptyState, sgd, apply_updates)  # fmt: skip
    src = types.ModuleType("optax._src")
    base = types.ModuleType("optax._src.base")
    base.PyTree, base.OptState, base.Params = Any, Any, Any
    base.GradientTransformation = GradientTransformation
    src.base = base
    optax._src = src
    return {"optax": optax, "optax._src": src, "optax._src.base": base}


def install():
    """Put the stand-in modules into sys.modules (idempotent).  Returns the jax module."""
    if isinstance(sys.modules.get("jax"), types.ModuleType) and getattr(sys.modules["jax"], "_fb200_standin", False):
        return sys.modules["jax"]
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if root not in sys.path:
        sys.path.insert(0, root)
    jax = types.ModuleType("jax")
    jax._fb200_standin = True
    jnp = _make_jnp()
    rnd = _make_random()
    nn, jsp, special = _make_nn_and_special()
    lax = _Lax("jax.lax")
    tu = types.ModuleType("jax.tree_util")
    for f in (tree_map, tree_flatten, tree_unflatten, tree_leaves, tree_structure, tree_reduce):
        setattr(tu, f.__name__, f)
    fu = types.ModuleType("jax.flatten_util")
    fu.ravel_pytree = ravel_pytree
    jax.numpy, jax.random, jax.nn, jax.scipy, jax.lax, jax.tree_util, jax.flatten_util = jnp, rnd, nn, jsp, lax, tu, fu
    jax.vmap, jax.jit, jax.pmap, jax.tree_map = vmap, _identity_transform, _pmap, tree_map
    jax.pure_callback = lambda fn, _shape, *args, **kw: fn(*args)
    jax.Array = JArray
    jax.local_device_count = lambda: 1
    jax.device_get = lambda x: x
    mods = {
        "jax": jax, "jax.numpy": jnp, "jax.numpy.fft": jnp.fft, "jax.random": rnd, "jax.nn": nn, "jax.scipy": jsp,
        "jax.scipy.special": special, "jax.lax": lax, "jax.tree_util": tu, "jax.flatten_util": fu,
    }  # fmt: skip
    mods.update(_make_optax())
    flax = _AutoStub("flax")
    flax.__path__ = []
    core = _AutoStub("flax.core")
    core.FrozenDict = dict
    core.__path__ = []
    flax.core = core
    mods.update({"flax": flax, "flax.core": core})
    for name in ("flax.training", "flax.training.common_utils", "flax.training.checkpoints", "flax.jax_utils",
                 "flax.linen", "flax.core.frozen_dict", "flax.traverse_util", "flax.struct"):  # fmt: skip
        s = _AutoStub(name)
        s.__path__ = []
        mods[name] = s
    mods["flax.core.frozen_dict"].FrozenDict = dict
    mods["flax.training.common_utils"].shard_prng_key = lambda k: k
    mods["flax.training.checkpoints"].PyTree = Any
    sys.modules.update(mods)
    return jax


# ---------------------------------------------------------------------------------------------------------------------
# importing the reference's source files
# ---------------------------------------------------------------------------------------------------------------------
# The files executed as they are.  Everything else under fortuna.* resolves to an auto-stub (empty classes): loaders,
# plotting, model managers, flax modules, trainers - none of which is on the path.
REAL_MODULES = (
    "fortuna.utils.random",
    "fortuna.utils.training",
    "fortuna.conformal.classification.base",
    "fortuna.conformal.classification.adaptive_prediction",
    "fortuna.conformal.classification.simple_prediction",
    "fortuna.conformal.regression.base",
    "fortuna.conformal.regression.quantile",
    "fortuna.conformal.regression.onedim_uncertainty",
    "fortuna.metric.classification",
    "fortuna.metric.regression",
    "fortuna.prob_model.posterior.sgmcmc.sgmcmc_preconditioner",
    "fortuna.prob_model.posterior.sgmcmc.sgmcmc_step_schedule",
    "fortuna.prob_model.posterior.sgmcmc.sgmcmc_diagnostic",
    "fortuna.prob_model.posterior.sgmcmc.sgmcmc_posterior",
    "fortuna.prob_model.posterior.sgmcmc.sghmc.sghmc_integrator",
    "fortuna.prob_model.posterior.sgmcmc.cyclical_sgld.cyclical_sgld_integrator",
    "fortuna.prob_output_layer.base",
    "fortuna.prob_output_layer.classification",
    "fortuna.prob_output_layer.regression",
    "fortuna.likelihood.base",
    "fortuna.likelihood.classification",
    "fortuna.likelihood.regression",
    "fortuna.prob_model.predictive.base",
    "fortuna.prob_model.predictive.classification",
    "fortuna.prob_model.predictive.regression",
    "fortuna.conformal.multivalid.base",
    "fortuna.conformal.multivalid.iterative.base",
    "fortuna.conformal.multivalid.iterative.batch_mvp",
    "fortuna.conformal.multivalid.iterative.multicalibrator",
    "fortuna.conformal.multivalid.iterative.regression.batch_mvp",
    "fortuna.conformal.multivalid.iterative.classification.binary_multicalibrator",
    "fortuna.conformal.multivalid.iterative.classification.top_label_multicalibrator",
    "fortuna.conformal.multivalid.mixins.batchmvp",
    "fortuna.conformal.multivalid.mixins.multicalibrator",
    "fortuna.conformal.multivalid.mixins.classification.binary_multicalibrator",
    "fortuna.conformal.multivalid.mixins.classification.top_label_multicalibrator",
    "fortuna.conformal.multivalid.one_shot.base",
    "fortuna.conformal.multivalid.one_shot.multicalibrator",
    "fortuna.conformal.multivalid.one_shot.classification.binary_multicalibrator",
    "fortuna.conformal.multivalid.one_shot.classification.top_label_multicalibrator",
    "fortuna.conformal.classification.maxcovfixprec_binary_classfication",
    "fortuna.prob_model.posterior.sgmcmc.sgmcmc_sampling_callback",
    "fortuna.prob_model.posterior.sgmcmc.sghmc.sghmc_callback",
    "fortuna.prob_model.posterior.sgmcmc.cyclical_sgld.cyclical_sgld_callback",
    "fortuna.output_calib_model.predictive.base",
    "fortuna.output_calib_model.predictive.classification",
    "fortuna.output_calib_model.predictive.regression",
    "fortuna.loss.classification.focal_loss",
    "fortuna.loss.regression.scaled_mse",
)


class _RefFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname != "fortuna" and not fullname.startswith("fortuna."):
            return None
        return importlib.machinery.ModuleSpec(fullname, self, is_package=fullname not in REAL_MODULES)

    def create_module(self, spec):
        if spec.name in REAL_MODULES:
            return None
        m = _AutoStub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        name = module.__name__
        if name not in REAL_MODULES:
            if name == "fortuna.typing":
                for t in ("Array", "Params", "Mutable", "CalibParams", "CalibMutable", "Batch", "Path", "Status",
                          "OptaxOptimizer", "InputData", "Targets", "Uncertainties", "Outputs", "Predictions",
                          "AnyKey", "Shape"):  # fmt: skip
                    setattr(module, t, Any)
            return
        path = os.path.join(REFERENCE_ROOT, *name.split(".")) + ".py"
        with open(path) as f:
            src = f.read()
        module.__file__ = path
        exec(compile(src, path, "exec"), module.__dict__)


_FINDER = _RefFinder()


def load_reference():
    """install() + make ``import fortuna.<module>`` execute the reference's file for every name in REAL_MODULES."""
    install()
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"{REFERENCE_ROOT} is not present: fixtures can only be regenerated in the build container")
    if _FINDER not in sys.meta_path:
        sys.meta_path.insert(0, _FINDER)
    return importlib.import_module
