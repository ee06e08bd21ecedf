"""Pairwise contraction front-end over the C ABI.

``contract_pair`` is the single primitive that ``tensordot`` / ``einsum`` /
the tree executor lower to: it never transposes or copies an operand, the
permutation implied by the labels is folded into the CUDA kernel's tile
loads (see ``csrc/contract_dmma.cu``).
"""

import ctypes

import torch

from . import _lib

_WS = {}


def _workspace(nbytes, device):
    """Grow-only per-(device, stream) scratch buffer."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8,
                          device=device)
        _WS[key] = buf
    return buf


def contract_pair(a, la, b, lb, lc, conj_a=False, conj_b=False, out=None,
                  engine=0, alpha=1.0, beta=0.0):
    """``out[lc] = sum op(a)[la] * op(b)[lb]`` with integer mode labels.

    Parameters
    ----------
    a, b : torch.Tensor (cuda)
        Arbitrarily strided views; never copied.
    la, lb, lc : sequences of int
        Mode labels of ``a``, ``b`` and the output (its axis order).
    out : torch.Tensor, optional
        Pre-allocated (possibly strided) output view.
    alpha, beta : float
        ``out = alpha * contraction + beta * out`` (beta != 0 needs ``out``).
    """
    _lib.require_cuda(a, "a")
    _lib.require_cuda(b, "b")
    if a.dtype != b.dtype:
        raise TypeError(f"dtype mismatch: {a.dtype} vs {b.dtype}")
    if len(la) != a.dim() or len(lb) != b.dim():
        raise ValueError("label count does not match operand rank")
    lib = _lib.load()
    ext = {}
    for t, ls in ((a, la), (b, lb)):
        for l, s in zip(ls, t.shape):
            if ext.setdefault(l, s) != s:
                raise ValueError(
                    f"extent mismatch for label {l}: {ext[l]} vs {s}")
    if out is None:
        try:
            shape = [ext[l] for l in lc]
        except KeyError as e:
            raise ValueError(f"output label {e} appears in neither input")
        out = torch.empty(shape, dtype=a.dtype, device=a.device)
    else:
        _lib.require_cuda(out, "out")
    da, db, dc = _lib.desc(a), _lib.desc(b), _lib.desc(out)
    pla, plb, plc = _lib.labels(la), _lib.labels(lb), _lib.labels(lc)
    need = lib.qb_contract_pair_workspace(da, pla, db, plb, dc, plc, engine)
    if need < 0:
        _lib.check(int(need), "qb_contract_pair_workspace")
    ws_ptr, ws_n = None, 0
    if need > 0:
        ws = _workspace(need, a.device)
        ws_ptr, ws_n = ctypes.c_void_p(ws.data_ptr()), ws.numel()
    if alpha == 1.0 and beta == 0.0:
        rc = lib.qb_contract_pair(da, pla, db, plb, dc, plc, int(bool(conj_a)),
                                  int(bool(conj_b)), engine, ws_ptr, ws_n,
                                  _lib.stream_ptr())
    else:
        rc = lib.qb_contract_pair_ab(da, pla, db, plb, dc, plc,
                                     int(bool(conj_a)), int(bool(conj_b)),
                                     float(alpha), float(beta), ws_ptr, ws_n,
                                     _lib.stream_ptr())
    _lib.check(rc, "qb_contract_pair")
    return out


def plan_pair(a_shape, a_strides, la, b_shape, b_strides, lb, c_shape,
              c_strides, lc, dtype_code=_lib.QB_F64):
    """Host-only: the GEMM view the C planner derives (no GPU needed)."""
    lib = _lib.load()
    da = _lib.np_desc(a_shape, a_strides, dtype_code)
    db = _lib.np_desc(b_shape, b_strides, dtype_code)
    dc = _lib.np_desc(c_shape, c_strides, dtype_code)
    out = (ctypes.c_int64 * 16)()
    rc = lib.qb_contract_pair_plan(da, _lib.labels(la), db, _lib.labels(lb),
                                   dc, _lib.labels(lc), out)
    _lib.check(rc, "qb_contract_pair_plan")
    keys = ("M", "N", "K", "batch", "n_m", "n_n", "n_k", "n_b", "cfg",
            "splitk", "vecA", "vecB", "vecC", "thrA", "thrB", "degenerate")
    return dict(zip(keys, list(out)))
