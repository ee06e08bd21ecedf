"""``quimb_b200.linalg`` -- what ``do("linalg.svd" | "linalg.qr" | ...,
like="quimb_b200")`` resolves to.  SVD and QR run on the dedicated CUDA
kernels (``csrc/linalg.cu``: cluster-resident Householder panels, one-sided
block Jacobi); nothing here falls back to a CPU or library factorization.
"""

import ctypes

import numpy as np
import torch

from . import _lib, ops
from .array import Array

_WS = {}


def _workspace(nbytes, device):
    key = ("linalg", device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


def _narrow(fn):
    """float32 input: widen exactly, factor in fp64, round the factors once
    (dtype is preserved end to end, as the reference does)."""
    import functools

    @functools.wraps(fn)
    def wrapped(x, *args, **kwargs):
        x = ops.asarray(x)
        if x.t.dtype != torch.float32:
            return fn(x, *args, **kwargs)
        from .contract import convert
        wide = Array(convert(ops.materialize(x).t, torch.float64))
        outs = fn(wide, *args, **kwargs)
        return tuple(Array(convert(o.t, torch.float32)) if isinstance(o, Array) else o
                     for o in outs)
    return wrapped


def _as_matrix(x):
    x = ops.materialize(ops.asarray(x))
    if x.ndim != 2:
        raise ValueError("quimb_b200.linalg: only 2-d arrays are supported "
                         f"(got shape {x.shape})")
    if x.t.dtype != torch.float64:
        raise TypeError(
            f"quimb_b200.linalg: dtype {x.dtype} is not implemented yet "
            "(float64 only); no fallback exists")
    _lib.require_cuda(x.t)
    return x


@_narrow
def qr(x, stabilized=False, want_q=True, want_r=True):
    """Thin QR of a 2-d device array: Q (m, k), R (k, n), k = min(m, n).
    ``stabilized`` makes diag(R) >= 0 (quimb's qr_stabilized convention)."""
    x = _as_matrix(x)
    m, n = x.shape
    k = min(m, n)
    lib = _lib.load()
    dev = x.t.device
    if m < n:
        # QR of the leading m x m block, R2 = Q^T X[:, m:]
        q, r1 = qr(Array(x.t[:, :m]), stabilized=stabilized)
        r2 = ops.tensordot(q, Array(x.t[:, m:]), axes=((0,), (0,)))
        r = torch.cat([r1.t, r2.t], dim=1)
        return (q if want_q else None), (Array(r) if want_r else None)
    need = lib.qb_qr_workspace(_lib.QB_F64, m, n)
    if need < 0:
        raise ValueError(f"quimb_b200.linalg.qr: unsupported shape {x.shape}")
    ws = _workspace(need, dev)
    Q = torch.empty((m, k), dtype=x.t.dtype, device=dev) if want_q else None
    R = torch.empty((k, n), dtype=x.t.dtype, device=dev) if want_r else None
    rc = lib.qb_qr_stab(_lib.QB_F64, m, n, x.t.data_ptr(),
                        Q.data_ptr() if want_q else None,
                        R.data_ptr() if want_r else None,
                        int(bool(stabilized)), ws.data_ptr(), ws.numel(),
                        _lib.stream_ptr())
    _lib.check(rc, "qb_qr_stab")
    return (Array(Q) if want_q else None), (Array(R) if want_r else None)


@_narrow
def svd(x, full_matrices=False, return_sweeps=False):
    """Thin SVD: U (m, k), s (k,) descending, VH (k, n)."""
    if full_matrices:
        raise NotImplementedError("quimb_b200.linalg.svd: thin SVD only")
    x = _as_matrix(x)
    m, n = x.shape
    if m < n:
        xt = ops.materialize(Array(x.t.t()))
        out = svd(xt, return_sweeps=return_sweeps)
        u, s, vh = out[:3]
        res = (Array(vh.t.t()), s, Array(u.t.t()))
        return res + (out[3],) if return_sweeps else res
    lib = _lib.load()
    dev = x.t.device
    need = lib.qb_svd_workspace(_lib.QB_F64, m, n)
    if need < 0:
        raise ValueError(f"quimb_b200.linalg.svd: unsupported shape {x.shape}")
    ws = _workspace(need, dev)
    U = torch.empty((m, n), dtype=x.t.dtype, device=dev)
    S = torch.empty((n,), dtype=x.t.dtype, device=dev)
    VH = torch.empty((n, n), dtype=x.t.dtype, device=dev)
    sweeps = ctypes.c_int(0)
    rc = lib.qb_svd(_lib.QB_F64, m, n, x.t.data_ptr(), U.data_ptr(),
                    S.data_ptr(), VH.data_ptr(), ws.data_ptr(), ws.numel(),
                    ctypes.byref(sweeps), _lib.stream_ptr())
    _lib.check(rc, "qb_svd")
    out = (Array(U), Array(S), Array(VH))
    return out + (sweeps.value,) if return_sweeps else out


def norm(x, ord=None):
    """Frobenius / 2-norm through the deterministic dot kernel."""
    x = ops.asarray(x)
    if ord not in (None, "fro", 2):
        raise NotImplementedError("only the Frobenius / vector 2-norm")
    flat = ops.materialize(x).reshape(-1)
    return ops.sqrt(ops.real(ops.vdot(flat, flat)))


def eigh(x):
    """Small dense symmetric eigenproblems (Lanczos tridiagonals, DMRG's
    dense-Heff branch for prod(dims) < 800) are host-side control logic in
    the reference too (dmrg.py:690); they are solved on the host."""
    if x.shape[0] > 64:
        raise NotImplementedError(
            "quimb_b200.linalg.eigh: only the tiny projected (Krylov) problems "
            "are solved here; for operators use eigh_lanczos (no dense device "
            "eigh yet, and no CPU fallback)")
    a = ops.to_numpy(x)
    w, v = np.linalg.eigh(a)
    return ops.asarray(w), ops.asarray(v)
