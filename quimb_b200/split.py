"""Split drivers: the host-side mirror of quimb's ``array_split`` /
``svd_truncated`` / ``qr_stabilized`` / ``tensor_split`` for device arrays
(quimb/tensor/decomp.py:35-174, 369-424, 829-1118, 2055-2216;
quimb/tensor/tensor_core.py:392-668).  Same names, same option codes, same
error behaviour; the factorizations run on the CUDA kernels of
``quimb_b200.linalg``, the absorb step is one diagonal-scaling kernel, and
the truncation rule is the reference's numba rule evaluated by the C library
on the singular values (the one unavoidable device->host read per split: the
kept rank decides the output shapes).
"""

import ctypes

import numpy as np
import torch

from . import _lib, linalg, ops
from .array import Array

# absorb codes (decomp.py:201-211); None = 'U,s,VH'
get_s = 2
get_Usq = -12
get_VH = -11
get_Us = -10
get_Us_VH = -1
get_Usq_sqVH = 0
get_U_sVH = 1
get_U = 10
get_sVH = 11
get_sqVH = 12

_ABSORB_MAP = {}
for _mode, _aliases in [
    (None, ["U,s,VH"]), (get_s, ["s"]), (get_Usq, ["lsqrt"]),
    (get_VH, ["VH", "rorthog"]), (get_Us, ["Us", "lfactor"]),
    (get_Us_VH, ["Us,VH", "left"]), (get_Usq_sqVH, ["Usq,sqVH", "both"]),
    (get_U_sVH, ["U,sVH", "right"]), (get_U, ["U", "lorthog"]),
    (get_sVH, ["sVH", "rfactor"]), (get_sqVH, ["sqVH", "rsqrt"]),
]:
    _ABSORB_MAP[_mode] = _mode
    for _a in _aliases:
        _ABSORB_MAP[_a] = _mode

_CUTOFF_MODE_MAP = {1: 1, "abs": 1, 2: 2, "rel": 2, 3: 3, "sum2": 3,
                    4: 4, "rsum2": 4, 5: 5, "sum1": 5, 6: 6, "rsum1": 6}
_RENORM_LOOKUP = {3: 2, 4: 2, 5: 1, 6: 1}
_SVD_ABSORBS = set(_ABSORB_MAP.values())
_QR_ABSORBS = {get_U_sVH, get_U, get_sVH, get_Us_VH, get_Us, get_VH}


_DEFAULT_ABSORB = {"svd": get_Usq_sqVH, "qr": get_U_sVH}


def parse_method_absorb(method="auto", absorb="auto", truncation=True):
    """decomp.py:307-365: resolve 'auto' settings, map aliases to codes."""
    if method == "auto":
        if truncation or absorb == "auto":
            method = "svd"
        else:
            absorb = _ABSORB_MAP[absorb]
            method = "qr" if absorb in _QR_ABSORBS else "svd"
    if method.startswith("lq"):
        # lq methods are simply qr with a different default absorb
        method = "qr" + method[2:]
        if absorb == "auto":
            absorb = "left"
    if method not in _DEFAULT_ABSORB:
        raise ValueError(f"quimb_b200: split method {method!r} is not "
                         "implemented (available: 'svd', 'qr', 'lq')")
    if absorb == "auto":
        absorb = _DEFAULT_ABSORB[method]
    else:
        absorb = _ABSORB_MAP[absorb]
    return method, absorb


def parse_split_opts(method="auto", absorb="auto", max_bond=None,
                     cutoff=1e-10, cutoff_mode="rsum2", renorm=None):
    """Resolve options to numeric codes exactly as decomp.py:369-424 does for
    the methods this backend implements ('svd', 'qr' / 'lq')."""
    max_bond = -1 if max_bond is None else max_bond
    cutoff = -1.0 if cutoff is None else cutoff
    truncation = (max_bond > 0) or (cutoff > 0.0)
    method, absorb = parse_method_absorb(method, absorb, truncation)
    opts = {"absorb": absorb}
    if method == "svd":
        cutoff_mode = _CUTOFF_MODE_MAP[cutoff_mode]
        if renorm is True:
            renorm = _RENORM_LOOKUP.get(cutoff_mode, 0)
        else:
            renorm = 0 if renorm is None else renorm
        opts.update(max_bond=max_bond, cutoff=cutoff, cutoff_mode=cutoff_mode,
                    renorm=renorm)
    elif absorb is None:
        raise ValueError("You can't return the singular values separately when "
                         f"`method='{method}'`.")
    return method, opts


def _scale_diag(x, d, side, sqrt_d):
    """x (rows, cols) *= d[col]^p (side=1) or d[row]^p (side=0), in place."""
    lib = _lib.load()
    rows, cols = x.shape
    rc = lib.qb_scale_diag(_lib.qb_dtype(x.dtype), rows, cols, x.data_ptr(),
                           d.data_ptr(), side, int(sqrt_d), _lib.stream_ptr())
    _lib.check(rc, "qb_scale_diag")
    return x


def svals_to_keep(s_host, cutoff, cutoff_mode, max_bond, renorm):
    """(n_keep, renorm factor, truncation error) by the reference's rule
    (decomp.py:901-965, 968-1029), evaluated by the C library."""
    lib = _lib.load()
    s_host = np.ascontiguousarray(s_host, dtype=np.float64)
    n_keep = ctypes.c_int64(0)
    f = ctypes.c_double(1.0)
    err = ctypes.c_double(0.0)
    rc = lib.qb_svals_to_keep(
        s_host.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), s_host.size,
        float(cutoff), int(cutoff_mode), int(max_bond), int(renorm),
        ctypes.byref(n_keep), ctypes.byref(f), ctypes.byref(err))
    _lib.check(rc, "qb_svals_to_keep")
    return int(n_keep.value), float(f.value), float(err.value)


def svd_truncated(x, cutoff=-1.0, cutoff_mode=4, max_bond=-1,
                  absorb=get_Usq_sqVH, renorm=0, info=None):
    """Truncated SVD of a 2-d device array; returns (left, s, right) with
    ``None`` for parts the absorb mode does not request."""
    absorb = _ABSORB_MAP[absorb]
    cutoff_mode = _CUTOFF_MODE_MAP[cutoff_mode]
    U, s, VH = linalg.svd(x)
    s_host = s.t.cpu().numpy()  # the one host read of the split
    n_keep, f, err = svals_to_keep(s_host, cutoff, cutoff_mode, max_bond, renorm)
    if info is not None and "error" in info:
        info["error"] = err
    if info is not None:
        info["n_keep"] = n_keep
    Ut, st, Vt = U.t, s.t, VH.t
    if n_keep < st.shape[0]:
        Ut = Ut[:, :n_keep].contiguous()
        Vt = Vt[:n_keep, :]
        st = st[:n_keep]
        if f != 1.0:
            st = st * f
        st = st.contiguous()
    want_left = absorb in (None, get_Usq, get_Us, get_Us_VH, get_Usq_sqVH,
                           get_U_sVH, get_U)
    want_right = absorb in (None, get_VH, get_Us_VH, get_Usq_sqVH, get_U_sVH,
                            get_sVH, get_sqVH)
    left = right = sv = None
    if absorb is None:
        return Array(Ut), Array(st), Array(Vt)
    if absorb == get_s:
        return None, Array(st), None
    if want_left:
        if absorb in (get_Us_VH, get_Us):
            left = _scale_diag(Ut if Ut.is_contiguous() else Ut.contiguous(), st, 1, False)
        elif absorb in (get_Usq_sqVH, get_Usq):
            left = _scale_diag(Ut if Ut.is_contiguous() else Ut.contiguous(), st, 1, True)
        else:
            left = Ut
        left = Array(left)
    if want_right:
        Vc = Vt if Vt.is_contiguous() else Vt.contiguous()
        if absorb in (get_U_sVH, get_sVH):
            right = _scale_diag(Vc, st, 0, False)
        elif absorb in (get_Usq_sqVH, get_sqVH):
            right = _scale_diag(Vc, st, 0, True)
        else:
            right = Vc
        right = Array(right)
    return left, sv, right


def qr_stabilized(x, absorb=get_U_sVH, stabilized=True):
    """QR (or LQ for the 'left' family of absorbs) with diag(R) >= 0;
    returns (left, None, right) like decomp.py:2055-2144."""
    absorb = _ABSORB_MAP[absorb]
    if absorb in (get_U_sVH, get_U, get_sVH):
        Q, R = linalg.qr(x, stabilized=stabilized, want_q=absorb != get_sVH,
                         want_r=absorb != get_U)
        return Q, None, R
    if absorb in (get_Us_VH, get_Us, get_VH):
        xt = Array(ops.asarray(x).t.t(), ops.asarray(x).cj)
        Q, R = linalg.qr(xt, stabilized=stabilized, want_q=absorb != get_Us,
                         want_r=absorb != get_VH)
        left = None if R is None else Array(R.t.t())
        right = None if Q is None else Array(Q.t.t())
        return left, None, right
    raise ValueError(f"Invalid absorb mode for qr_stabilized: {absorb}")


def array_split(x, method="auto", absorb="auto", max_bond=None, cutoff=1e-10,
                cutoff_mode="rsum2", renorm=None, info=None):
    """decomp.py:35-174 for the implemented methods."""
    method, opts = parse_split_opts(method, absorb, max_bond, cutoff,
                                    cutoff_mode, renorm)
    if method == "svd":
        return svd_truncated(x, info=info, **opts)
    return qr_stabilized(x, absorb=opts["absorb"])


def tensor_split(x, inds, left_inds, right_inds=None, method="auto",
                 absorb="auto", max_bond=None, cutoff=1e-10,
                 cutoff_mode="rel", renorm=None, info=None):
    """Array-level ``tensor_split(..., get='arrays')`` (tensor_core.py:392-668):
    transpose to (left..., right...), fuse to a matrix (one permute-copy
    kernel), split, unfuse.  Returns the non-None parts in order
    (left, [s], right); the new bond is last on left, first on right."""
    x = ops.asarray(x)
    inds = tuple(inds)
    left_inds = tuple(left_inds)
    if right_inds is None:
        right_inds = tuple(ix for ix in inds if ix not in left_inds)
    else:
        right_inds = tuple(right_inds)
    if set(left_inds + right_inds) != set(inds) or len(left_inds + right_inds) != len(inds):
        raise ValueError("'output_inds' must be permutation of the current "
                         "tensor indices")
    perm = [inds.index(ix) for ix in left_inds + right_inds]
    xt = x.transpose(*perm)
    ldims = xt.shape[:len(left_inds)]
    rdims = xt.shape[len(left_inds):]
    mat = xt.reshape(int(np.prod(ldims, dtype=np.int64)),
                     int(np.prod(rdims, dtype=np.int64)))
    left, s, right = array_split(mat, method=method, absorb=absorb,
                                 max_bond=max_bond, cutoff=cutoff,
                                 cutoff_mode=cutoff_mode, renorm=renorm,
                                 info=info)
    out = []
    if left is not None:
        out.append(left.reshape(*ldims, -1))
    if s is not None:
        out.append(s)
    if right is not None:
        out.append(right.reshape(-1, *rdims))
    return tuple(out)


def _out_perm(have, want):
    return [have.index(ix) for ix in want]


def tensor_canonize_bond(a, a_inds, b, b_inds, absorb="right"):
    """Array-level ``tensor_canonize_bond`` (tensor_core.py:671-824): QR ``a``
    over the bond shared with ``b``, absorb R into ``b`` (or the LQ mirror for
    ``absorb='left'``).  Outputs keep the index order of the inputs; the
    absorption is one launch of the contraction kernel writing straight into
    that order."""
    from .contract import contract_pair
    a, b = ops.asarray(a), ops.asarray(b)
    a_inds, b_inds = tuple(a_inds), tuple(b_inds)
    shared = [ix for ix in a_inds if ix in b_inds]
    if len(shared) != 1:
        raise ValueError("The tensors specified don't share an bond.")
    if absorb == "left":
        nb, na = tensor_canonize_bond(b, b_inds, a, a_inds, "right")
        return na, nb
    bond = shared[0]
    lix = tuple(ix for ix in a_inds if ix != bond)
    q, r = tensor_split(a, a_inds, lix, (bond,), method="qr")
    new_a = q.transpose(*_out_perm(lix + (bond,), a_inds))
    lab = {ix: i for i, ix in enumerate(dict.fromkeys(a_inds + b_inds))}
    K = len(lab)  # label of the new bond
    lb = [lab[ix] for ix in b_inds]
    out = [K if ix == bond else lab[ix] for ix in b_inds]
    new_b = Array(contract_pair(r.t, [K, lab[bond]], b.t, lb, out, conj_a=r.cj, conj_b=b.cj))
    return new_a, new_b


def tensor_compress_bond(a, a_inds, b, b_inds, max_bond=None, cutoff=1e-10,
                         cutoff_mode="rel", absorb="both", renorm=None, info=None):
    """Array-level ``tensor_compress_bond`` (tensor_core.py:864-1094, the
    default ``reduced=True`` pipeline): QR(a), LQ(b), truncated SVD of the
    reduced core, factors folded back with two contractions.  Outputs keep the
    index order of the inputs."""
    from .contract import contract_pair
    a, b = ops.asarray(a), ops.asarray(b)
    a_inds, b_inds = tuple(a_inds), tuple(b_inds)
    shared = [ix for ix in a_inds if ix in b_inds]
    if len(shared) != 1:
        raise ValueError("The tensors specified don't share an bond. "
                         "To create one automatically, set `create_bond=True`.")
    bond = shared[0]
    lix = tuple(ix for ix in a_inds if ix != bond)
    rix = tuple(ix for ix in b_inds if ix != bond)
    qa, ra = tensor_split(a, a_inds, lix, (bond,), method="qr")      # (*lix,k1), (k1,bond)
    lb_, qb = tensor_split(b, b_inds, (bond,), rix, method="lq")     # (bond,k2), (k2,*rix)
    core = ops.tensordot(ra, lb_, axes=((1,), (0,)))                 # (k1, k2)
    _, opts = parse_split_opts("svd", absorb, max_bond, cutoff, cutoff_mode, renorm)
    cl, _, cr = svd_truncated(core, info=info, **opts)               # (k1,k), (k,k2)
    lab = {ix: i for i, ix in enumerate(dict.fromkeys(a_inds + b_inds))}
    K1, K2, KB = len(lab), len(lab) + 1, lab[bond]
    la = [lab[ix] for ix in lix] + [K1]
    new_a = Array(contract_pair(qa.t, la, cl.t, [K1, KB], [lab[ix] for ix in a_inds],
                                conj_a=qa.cj, conj_b=cl.cj))
    lq = [K2] + [lab[ix] for ix in rix]
    new_b = Array(contract_pair(cr.t, [KB, K2], qb.t, lq, [lab[ix] for ix in b_inds],
                                conj_a=cr.cj, conj_b=qb.cj))
    return new_a, new_b
