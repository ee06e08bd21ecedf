"""TEST INFRASTRUCTURE ONLY -- minimal functional stand-in for `autoray`.

autoray 0.10.1 (pinned in the reference's pixi.lock) is a third-party
dependency of quimb that is not installable offline.  This shim implements
just the dispatch surface quimb imports so that the
*unmodified reference* under /root/reference can be executed on its numpy
backend to generate golden vectors (oracle/make_golden.py).  It is never
imported by the product package `quimb_b200`; when the real autoray is
installed it takes precedence and this directory is simply not put on
sys.path.

Semantics restated from autoray's documented behaviour:
  * infer_backend(x): 'builtins' for python scalars, otherwise the top-level
    module name of type(x) ('numpy' for ndarray);
  * do(name, *args, like=None): look the function up by (backend, name) in
    the registry, else getattr-walk the backend module ('linalg.svd');
  * compose(fn): generic implementation + per-backend overrides through
    `.register(backend)`.
"""

import functools
import importlib
import inspect
import numbers

import numpy as np

from . import lazy  # noqa: F401

__version__ = "0.0.shim"

_FUNCS = {}
_MODULE_ALIASES = {"builtins": "numpy", "autoray.lazy": "numpy"}


def infer_backend(array):
    if isinstance(array, (numbers.Number, bool, np.generic)) and not isinstance(
        array, np.ndarray
    ):
        if isinstance(array, np.generic):
            return "numpy"
        return "builtins"
    if isinstance(array, np.ndarray):
        return "numpy"
    if array is None:
        return "builtins"
    mod = type(array).__module__
    if mod.startswith("autoray.lazy"):
        return "autoray.lazy"
    return mod.split(".")[0]


def infer_backend_multi(*arrays):
    backends = [infer_backend(a) for a in arrays]
    for b in backends:
        if b not in ("builtins", "numpy"):
            return b
    if "numpy" in backends:
        return "numpy"
    return backends[0] if backends else "numpy"


_multi_first = {"stack", "concatenate", "block", "vstack", "hstack"}


def _choose(fn_name, args, kwargs, like):
    if like is not None:
        return like if isinstance(like, str) else infer_backend(like)
    if not args:
        return "numpy"
    first = args[0]
    if fn_name in _multi_first or fn_name == "einsum":
        if fn_name == "einsum":
            return infer_backend_multi(*args[1:])
        return infer_backend_multi(*first)
    if fn_name == "tensordot":
        return infer_backend_multi(*args[:2])
    if isinstance(first, (tuple, list)) and first and not isinstance(
        first[0], numbers.Number
    ):
        return infer_backend_multi(*first)
    return infer_backend(first)


def register_function(backend, name, fn=None, wrap=False):
    if fn is None:
        return functools.partial(register_function, backend, name, wrap=wrap)
    if wrap:
        fn = fn(get_lib_fn(backend, name))
    _FUNCS[backend, name] = fn
    return fn


def _numpy_to_numpy(x):
    return np.asarray(x)


def _complex(re, im):
    return re + 1j * im


_NUMPY_EXTRA = {
    "to_numpy": _numpy_to_numpy,
    "complex": _complex,
    "linalg.expm": None,
}

_COMPOSED = {}


def get_lib_fn(backend, fn):
    try:
        return _FUNCS[backend, fn]
    except KeyError:
        pass
    if fn in _COMPOSED:
        out = _COMPOSED[fn]._default_fn
        _FUNCS[backend, fn] = out
        return out
    modname = _MODULE_ALIASES.get(backend, backend)
    if modname == "numpy" and fn in _NUMPY_EXTRA and _NUMPY_EXTRA[fn]:
        out = _NUMPY_EXTRA[fn]
    else:
        if modname == "numpy" and fn.startswith("scipy."):
            mod = importlib.import_module("scipy")
            path = fn.split(".")[1:]
        else:
            mod = importlib.import_module(modname)
            path = fn.split(".")
        out = mod
        for part in path:
            out = getattr(out, part)
    _FUNCS[backend, fn] = out
    return out


def do(fn, *args, like=None, **kwargs):
    backend = _choose(fn, args, kwargs, like)
    return get_lib_fn(backend, fn)(*args, **kwargs)


class DoFunc:
    def __init__(self, fn):
        self.fn = fn

    def __call__(self, *args, **kwargs):
        return do(self.fn, *args, **kwargs)


class _Namespace:
    def __init__(self, backend, prefix=""):
        self._backend = backend
        self._prefix = prefix

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        full = f"{self._prefix}{name}"
        if name in ("linalg", "random", "fft", "scipy"):
            return _Namespace(self._backend, prefix=f"{full}.")
        fn = get_lib_fn(self._backend, full)
        return fn


def get_namespace(like=None, **kwargs):
    if isinstance(like, str):
        backend = like
    else:
        backend = infer_backend(like)
    return _Namespace(backend)


class Composed:
    def __init__(self, fn, name=None):
        self._default_fn = fn
        self.__name__ = self._name = name or fn.__name__
        self.__doc__ = fn.__doc__
        self.__wrapped__ = fn
        _COMPOSED[self._name] = self
        try:
            self.__signature__ = inspect.signature(fn)
        except (TypeError, ValueError):
            pass

    def register(self, backend, fn=None):
        if fn is None:
            return functools.partial(self.register, backend)
        _FUNCS[backend, self._name] = fn
        return fn

    def __call__(self, *args, like=None, **kwargs):
        backend = _choose(self._name, args, kwargs, like)
        fn = _FUNCS.get((backend, self._name), self._default_fn)
        return fn(*args, **kwargs)


def compose(fn=None, *, name=None):
    if fn is None:
        return functools.partial(compose, name=name)
    return Composed(fn, name=name)


# ---- thin functional wrappers ---------------------------------------------
def conj(x):
    return do("conj", x)


def transpose(x, *args):
    return do("transpose", x, *args)


def dag(x):
    if ndim(x) < 2:
        return conj(x)
    return conj(do("swapaxes", x, -1, -2)) if ndim(x) > 2 else conj(do("transpose", x))


def real(x):
    return do("real", x)


def imag(x):
    return do("imag", x)


def reshape(x, shape):
    return do("reshape", x, shape)


def shape(x):
    return tuple(int(d) for d in x.shape) if hasattr(x, "shape") else np.shape(x)


def size(x):
    try:
        return int(x.size)
    except (AttributeError, TypeError):
        return int(np.size(x))


def ndim(x):
    try:
        return x.ndim
    except AttributeError:
        return np.ndim(x)


def get_dtype_name(x):
    try:
        return x.dtype.name
    except AttributeError:
        return str(np.asarray(x).dtype)


def to_backend_dtype(dtype_name, like):
    return np.dtype(dtype_name)


def get_common_dtype(*arrays):
    return np.result_type(*(get_dtype_name(a) for a in arrays)).name


def astype(x, dtype_name, **kwargs):
    dtype = to_backend_dtype(dtype_name, like=x)
    return x.astype(dtype, **kwargs)


def to_numpy(x):
    return do("to_numpy", x)


def to(x, like=None, backend=None, dtype=None, device=None):
    if backend in (None, "numpy") and (like is None or infer_backend(like) == "numpy"):
        out = np.asarray(x)
        if dtype is not None:
            out = out.astype(dtype)
        return out
    raise NotImplementedError("autoray shim: only numpy conversions supported")


def backend_like(like, set_globally="auto"):
    import contextlib

    return contextlib.nullcontext()


def autojit(fn=None, **kwargs):
    if fn is None:
        return functools.partial(autojit, **kwargs)
    return fn


def tree_map(f, tree, is_leaf=None):
    if isinstance(tree, (list, tuple)):
        return type(tree)(tree_map(f, t) for t in tree)
    if isinstance(tree, dict):
        return {k: tree_map(f, v) for k, v in tree.items()}
    return f(tree)


def tree_flatten(tree, get_ref=False):
    leaves = []

    def rec(t):
        if isinstance(t, (list, tuple)):
            return type(t)(rec(x) for x in t)
        if isinstance(t, dict):
            return {k: rec(v) for k, v in t.items()}
        leaves.append(t)
        return None

    ref = rec(tree)
    return (leaves, ref) if get_ref else leaves


def tree_unflatten(leaves, ref):
    it = iter(leaves)

    def rec(t):
        if isinstance(t, (list, tuple)):
            return type(t)(rec(x) for x in t)
        if isinstance(t, dict):
            return {k: rec(v) for k, v in t.items()}
        return next(it)

    return rec(ref)


# numpy-specific registrations mirroring autoray's translations
register_function("numpy", "linalg.svd",
                  lambda x, **kw: np.linalg.svd(x, full_matrices=False, **kw))
register_function("builtins", "to_numpy", _numpy_to_numpy)
register_function("numpy", "to_numpy", _numpy_to_numpy)
