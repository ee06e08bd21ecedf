"""Device array type of the ``quimb_b200`` backend.

quimb keeps whatever object it is given in ``Tensor._data`` as long as it has
a ``.shape`` (quimb/tensor/array_ops.py:31-33) and dispatches every numeric
call through autoray on ``type(x).__module__.split('.')[0]`` -- which for this
class is ``"quimb_b200"``, so ``do("tensordot", a, b, axes)`` lands on
:func:`quimb_b200.tensordot`.

The array wraps a ``torch.Tensor`` living on a CUDA device.  torch is the
*container* (allocator, strides, streams); the arithmetic on the hot path
(contraction, permute-copy, QR/SVD, Lanczos algebra) is done by the CUDA
kernels behind the C ABI.  Views (``transpose``, slicing, ``conj`` of real
data) never copy: the contraction kernel consumes arbitrary strides.
"""

import numbers

import numpy as np
import torch

from . import _lib

_NP2TORCH = {
    np.dtype("float32"): torch.float32,
    np.dtype("float64"): torch.float64,
    np.dtype("complex64"): torch.complex64,
    np.dtype("complex128"): torch.complex128,
    np.dtype("int64"): torch.int64,
    np.dtype("int32"): torch.int32,
    np.dtype("bool"): torch.bool,
}
_TORCH2NP = {v: k for k, v in _NP2TORCH.items()}


def torch_dtype(dt):
    if isinstance(dt, torch.dtype):
        return dt
    if isinstance(dt, str) and dt.startswith("torch."):
        dt = dt[6:]
    return _NP2TORCH[np.dtype(dt)]


def default_device():
    if not torch.cuda.is_available():
        raise _lib.QuimbB200Error(
            "quimb_b200 needs a CUDA device (no CPU fallback exists)")
    return torch.device("cuda", torch.cuda.current_device())


class Array:
    """A strided view of device memory (``torch.Tensor`` on cuda) plus a lazy
    conjugation flag that the contraction kernel folds into its loads."""

    __slots__ = ("t", "cj")
    __array_priority__ = 1000

    def __init__(self, t, cj=False):
        if isinstance(t, Array):
            t, cj = t.t, (t.cj != cj)
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"expected torch.Tensor, got {type(t)}")
        self.t = t
        # conjugating real data is the identity
        self.cj = bool(cj) and t.dtype.is_complex

    # ---- protocol quimb relies on -------------------------------------
    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def ndim(self):
        return self.t.dim()

    @property
    def size(self):
        return self.t.numel()

    @property
    def dtype(self):
        return _TORCH2NP[self.t.dtype]

    @property
    def device(self):
        return self.t.device

    def __len__(self):
        if self.t.dim() == 0:
            raise TypeError("len() of unsized object")
        return self.t.shape[0]

    def __repr__(self):
        return (f"quimb_b200.Array(shape={self.shape}, dtype={self.dtype.name},"
                f" device={self.t.device}{', conj' if self.cj else ''})")

    # ---- materialisation ----------------------------------------------
    def resolve(self):
        """torch view with any pending conjugation applied (may copy)."""
        if self.cj:
            from .ops import materialize
            return materialize(self).t
        return self.t

    def item(self, *args):
        """numpy's ``item()`` / ``item(flat_index)`` / ``item(i, j, ...)``"""
        t = self.t
        if len(args) == 1 and not isinstance(args[0], tuple):
            t = t.reshape(-1)[int(args[0])]
        elif args:
            t = t[tuple(args[0]) if len(args) == 1 else tuple(int(a) for a in args)]
        v = t.item()
        return v.conjugate() if self.cj else v

    def __float__(self):
        return float(self.item().real if isinstance(self.item(), complex)
                     else self.item())

    def __complex__(self):
        return complex(self.item())

    def __int__(self):
        return int(self.item())

    def __bool__(self):
        return bool(self.item())

    def __array__(self, dtype=None, copy=None):
        out = self.resolve().detach().cpu().numpy()
        return out if dtype is None else out.astype(dtype)

    def to_numpy(self):
        return self.__array__()

    # ---- views ----------------------------------------------------------
    def conj(self):
        return Array(self.t, not self.cj)

    conjugate = conj

    @property
    def real(self):
        return Array(self.resolve().real)

    @property
    def imag(self):
        t = self.resolve()
        return Array(t.imag if t.dtype.is_complex else torch.zeros_like(t))

    @property
    def T(self):
        return Array(self.t.permute(*reversed(range(self.t.dim()))), self.cj)

    @property
    def H(self):
        return self.T.conj()

    def transpose(self, *axes):
        if len(axes) == 1 and not isinstance(axes[0], numbers.Integral):
            axes = tuple(axes[0])
        if not axes:
            axes = tuple(reversed(range(self.t.dim())))
        return Array(self.t.permute(*axes), self.cj)

    def swapaxes(self, a, b):
        return Array(self.t.transpose(a, b), self.cj)

    def reshape(self, *shape):
        from .ops import reshape
        if len(shape) == 1 and not isinstance(shape[0], numbers.Integral):
            shape = tuple(shape[0])
        return reshape(self, shape)

    def ravel(self):
        return self.reshape(-1)

    flatten = ravel

    def squeeze(self, axis=None):
        t = self.t.squeeze() if axis is None else self.t.squeeze(axis)
        return Array(t, self.cj)

    def astype(self, dtype, copy=True):
        td = torch_dtype(dtype)
        t = self.resolve()
        if td == t.dtype:
            return Array(t.clone() if copy else t)
        if t.dtype.is_complex and not td.is_complex:
            t = t.real
        return Array(t.to(td))

    def diagonal(self, offset=0, axis1=0, axis2=1):
        return Array(torch.diagonal(self.resolve(), offset, axis1, axis2))

    # scipy's ``aslinearoperator`` accepts any object with ``shape`` and
    # ``matvec``: with these a dense device matrix can be handed to
    # ``scipy.sparse.linalg.eigsh`` / ``svds`` directly (what quimb's DMRG does
    # with its dense effective Hamiltonian when no eigensolver backend is
    # selected, dmrg.py:690-703 -> scipy_linalg.py:113-128): the Krylov vectors
    # live on the host, the products run on the device.
    def matvec(self, v):
        from .ops import asarray, matmul
        return matmul(self, asarray(np.asarray(v).astype(self.dtype, copy=False))).to_numpy()

    def matmat(self, m):
        return self.matvec(m)

    def rmatvec(self, v):
        from .ops import asarray, matmul
        return matmul(self.conj().swapaxes(0, 1),
                      asarray(np.asarray(v).astype(self.dtype, copy=False))).to_numpy()

    def toarray(self):
        """quimb's ``qarray`` protocol (DMRG reads eigenvectors through
        ``loc_gs.toarray()``, dmrg.py:841): already a plain array."""
        return self

    def copy(self):
        from .ops import materialize
        return materialize(self, force=True)

    def clone(self):
        return self.copy()

    @staticmethod
    def _index(idx):
        """unwrap device index arrays; numpy index arrays go as tensors"""
        def one(i):
            if isinstance(i, Array):
                return i.t
            if isinstance(i, np.ndarray):
                return torch.as_tensor(np.ascontiguousarray(i))
            return i
        if isinstance(idx, tuple):
            return tuple(one(i) for i in idx)
        if isinstance(idx, list) and any(isinstance(i, (Array, np.ndarray)) for i in idx):
            return tuple(one(i) for i in idx)
        return one(idx)

    def __getitem__(self, idx):
        return Array(self.t[self._index(idx)], self.cj)

    def __setitem__(self, idx, val):
        if self.cj:
            raise ValueError("cannot assign into a lazily conjugated view")
        if isinstance(val, Array):
            val = val.resolve()
        elif isinstance(val, np.ndarray):
            val = torch.as_tensor(np.ascontiguousarray(val), device=self.t.device)
        if isinstance(val, torch.Tensor) and val.dtype != self.t.dtype:
            if val.is_complex() and not self.t.is_complex():
                val = val.real                      # numpy: discards the imaginary part
            val = val.to(self.t.dtype)
        self.t[self._index(idx)] = val

    def __iter__(self):
        for i in range(self.t.shape[0]):
            yield Array(self.t[i], self.cj)

    # ---- arithmetic (element-wise: torch as the container library) -------
    def _bin(self, other, fn, reflect=False):
        a = self.resolve()
        if isinstance(other, Array):
            b = other.resolve()
        elif isinstance(other, np.ndarray):
            b = torch.as_tensor(np.ascontiguousarray(other), device=a.device)
        elif isinstance(other, (numbers.Number, np.generic)):
            b = other.item() if isinstance(other, np.generic) else other
        elif isinstance(other, torch.Tensor):
            b = other
        else:
            return NotImplemented
        return Array(fn(b, a) if reflect else fn(a, b))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, lambda x, y: x - y, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __rtruediv__(self, o): return self._bin(o, lambda x, y: x / y, True)
    def __pow__(self, o): return self._bin(o, torch.pow)
    def __rpow__(self, o): return self._bin(o, lambda x, y: x ** y, True)
    def __neg__(self): return Array(-self.resolve())
    def __pos__(self): return self
    def __abs__(self): return Array(self.t.abs())
    def _order(self, other, fn):
        """ordering comparisons; complex operands compare lexicographically
        (real part, then imaginary part) as numpy does -- torch has no complex
        ordering, and quimb's drivers compare 0-d complex overlaps with floats
        (``0.0 < psi.H @ psi < 1.0``)."""
        def cmp(a, b):
            ca = isinstance(a, complex) or (isinstance(a, torch.Tensor) and a.is_complex())
            cb = isinstance(b, complex) or (isinstance(b, torch.Tensor) and b.is_complex())
            if not (ca or cb):
                return fn(torch.as_tensor(a) if not isinstance(a, torch.Tensor) else a, b)
            dev = a.device if isinstance(a, torch.Tensor) else b.device
            a = torch.as_tensor(a, device=dev).to(torch.complex128)
            b = torch.as_tensor(b, device=dev).to(torch.complex128)
            strict = torch.lt if fn in (torch.lt, torch.le) else torch.gt
            return torch.where(a.real != b.real, strict(a.real, b.real), fn(a.imag, b.imag))
        return self._bin(other, cmp)

    def __lt__(self, o): return self._order(o, torch.lt)
    def __le__(self, o): return self._order(o, torch.le)
    def __gt__(self, o): return self._order(o, torch.gt)
    def __ge__(self, o): return self._order(o, torch.ge)
    def _logic(self, o, fn):
        def f(a, b):
            dev = a.device if isinstance(a, torch.Tensor) else b.device
            return fn(torch.as_tensor(a, device=dev), torch.as_tensor(b, device=dev))
        return self._bin(o, f)

    def __or__(self, o): return self._logic(o, torch.logical_or)
    __ror__ = __or__
    def __and__(self, o): return self._logic(o, torch.logical_and)
    __rand__ = __and__
    def __xor__(self, o): return self._logic(o, torch.logical_xor)
    __rxor__ = __xor__
    def __invert__(self): return Array(torch.logical_not(self.resolve()))

    def __format__(self, spec):
        if self.t.dim() == 0 and spec:
            return format(self.item(), spec)
        return repr(self) if spec else str(self)

    def __eq__(self, o): return self._bin(o, torch.eq)
    def __ne__(self, o): return self._bin(o, torch.ne)
    __hash__ = None

    def _inplace(self, r):
        """Write the result of a binary op back into this array's own storage
        (numpy semantics: views and aliases of the buffer see the update; a
        result that does not cast safely into the stored dtype raises)."""
        rt = r.resolve() if isinstance(r, Array) else torch.as_tensor(r, device=self.t.device)
        if rt.dtype != self.t.dtype and not torch.can_cast(rt.dtype, self.t.dtype):
            raise TypeError(f"Cannot cast in-place result from {rt.dtype} to "
                            f"{self.t.dtype} with casting rule 'same_kind'")
        if rt.shape != self.t.shape:
            raise ValueError(f"non-broadcastable output operand with shape "
                             f"{tuple(self.t.shape)} doesn't match the broadcast "
                             f"shape {tuple(rt.shape)}")
        self.t.copy_(rt.conj() if self.cj and rt.is_complex() else rt)
        return self

    def __iadd__(self, o): return self._inplace(self + o)
    def __isub__(self, o): return self._inplace(self - o)
    def __imul__(self, o): return self._inplace(self * o)
    def __itruediv__(self, o): return self._inplace(self / o)

    def __matmul__(self, other):
        from .ops import matmul
        return matmul(self, other)

    def dot(self, other):
        return self.__matmul__(other)

    def __rmatmul__(self, other):
        from .ops import matmul, asarray
        return matmul(asarray(other), self)

    # reductions used by quimb on arrays directly
    def sum(self, axis=None, keepdims=False, **kw):
        from .ops import sum as _sum
        return _sum(self, axis=axis, keepdims=keepdims)

    def max(self, axis=None, keepdims=False):
        from .ops import max as _max
        return _max(self, axis=axis, keepdims=keepdims)

    def min(self, axis=None, keepdims=False):
        from .ops import min as _min
        return _min(self, axis=axis, keepdims=keepdims)

    def all(self):
        return bool(self.t.all().item())

    def any(self):
        return bool(self.t.any().item())

    def tolist(self):
        return self.__array__().tolist()
