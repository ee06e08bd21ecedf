"""On-device Lanczos for the DMRG local eigensolve.

The reference solves ``eigh(Heff, k=1, which='SA', v0=..., ncv=4, tol=1e-3)``
with scipy's ARPACK (quimb/tensor/tn1d/dmrg.py:626-645 ->
quimb/linalg/scipy_linalg.py:113-128): the Krylov vectors live on the host
and every iteration calls back into ``TNLinearOperator._matvec``.  With a
device matvec that would cross PCIe twice per iteration, so the whole Krylov
process lives on the device here:

  * restarted Lanczos (thick restart with the Ritz vector, whose image H x is
    known, so a restart costs no matvec), full re-orthogonalisation by
    classical Gram-Schmidt applied twice, each pass = two fused HBM-bound
    kernels (``qb_multi_dot`` : h = V w,  ``qb_multi_axpy`` : w -= V^T h);
  * the images ``W_j = H v_j`` are kept, so the projected matrix
    ``V H V^T`` is one skinny contraction;
  * two small device->host reads per restart cycle (the ncv x ncv projected
    matrix, then the true residual norm |H x - theta x|); the tiny dense
    eigenproblem is host control logic, as in ARPACK.

Convergence test as in ARPACK's dsaupd: ``resid <= tol * max(eps^(2/3),
|theta|)``.  The basis size ``ncv`` is a free parameter of the device solver:
ARPACK's 4 is memory-frugal for host vectors; on a 180 GB device a basis of
8-16 vectors (32 MiB each at chi = 1024) needs far fewer matvecs.
``eigh_arpack_host_driver`` is the parity mode (scipy ARPACK on the host
driving the device matvec).
"""

import ctypes

import numpy as np
import torch

from . import _lib, ops
from .array import Array
from .contract import contract_pair
from .linalg import norm as _norm

_J, _N = 0, 1
_MD_WS = {}


def _md_ws(dev):
    ws = _MD_WS.get(dev.index)
    if ws is None:
        ws = torch.empty(_lib.load().qb_multi_dot_workspace(), dtype=torch.uint8,
                         device=dev)
        _MD_WS[dev.index] = ws
    return ws


def _norm_c(x, comm):
    """2-norm of a (possibly row-sharded) vector as a 0-d device Array."""
    if comm is None:
        return _norm(x)
    flat = ops.materialize(x).reshape(-1)
    n2 = ops.real(ops.vdot(flat, flat))
    comm.all_reduce_(n2.t)
    return ops.sqrt(n2)


def _orthogonalise(V, j, w, h, comm=None):
    """Two classical Gram-Schmidt passes of w against V[0..j] (in place).
    With ``comm`` the vectors are row slabs: the m dot products are summed
    over the ranks (one tiny all-reduce per pass)."""
    lib = _lib.load()
    m, n = j + 1, w.numel()
    st = _lib.stream_ptr()
    ws = _md_ws(w.device).data_ptr()
    for _ in range(2):
        rc = lib.qb_multi_dot(_lib.QB_F64, m, n, V.data_ptr(), V.stride(0),
                              w.data_ptr(), h.data_ptr(), ws, st)
        _lib.check(rc, "qb_multi_dot")
        if comm is not None:
            comm.all_reduce_(h)
        rc = lib.qb_multi_axpy(_lib.QB_F64, m, n, V.data_ptr(), V.stride(0),
                               h.data_ptr(), -1.0, w.data_ptr(), st)
        _lib.check(rc, "qb_multi_axpy")


def eigh_lanczos(matvec, v0, which="SA", ncv=4, tol=1e-3, maxiter=None,
                 return_info=False, comm=None):
    """Lowest ('SA') or highest ('LA') eigenpair of a Hermitian operator.

    Parameters
    ----------
    matvec : callable(Array[n]) -> Array[n]
        Device matvec on flat contiguous vectors.
    v0 : Array
        Start vector (any shape, flattened).
    ncv : int
        Krylov basis size between restarts (2 <= ncv <= 16).
    comm : object with ``all_reduce_(tensor)``, optional
        Row-sharded mode (quimb_b200.dist.BondShard): ``v0`` and every vector
        handed to / returned by ``matvec`` is this rank's slab; inner products
        are summed over the ranks, all ranks take identical decisions.
    Returns ``(theta: float, x: Array[n])`` (+ info dict).
    """
    v0 = ops.materialize(ops.asarray(v0)).reshape(-1)
    if v0.t.dtype == torch.complex128:
        # A Hermitian operator on C^n is a real symmetric operator on R^2n
        # (complex storage viewed as interleaved reals); each eigenvalue is
        # doubled and every real eigenvector is the image of a complex one,
        # so the real Krylov machinery applies unchanged.
        def mv_real(vr):
            vc = Array(torch.view_as_complex(vr.t.reshape(-1, 2)))
            out = ops.materialize(matvec(vc)).t.reshape(-1)
            return Array(torch.view_as_real(out).reshape(-1))

        vr0 = Array(torch.view_as_real(v0.t).reshape(-1))
        theta, xr, info = eigh_lanczos(mv_real, vr0, which=which, ncv=ncv, tol=tol,
                                       maxiter=maxiter, return_info=True, comm=comm)
        x = Array(torch.view_as_complex(xr.t.reshape(-1, 2)))
        return (theta, x, info) if return_info else (theta, x)
    n = v0.size
    dt = v0.t.dtype
    if dt != torch.float64:
        raise NotImplementedError(f"eigh_lanczos: dtype {dt} is not implemented "
                                  "yet (float64 only; no fallback)")
    dev = v0.t.device
    m = max(2, min(int(ncv), n, 16))
    if maxiter is None:
        maxiter = 1000
    V = torch.zeros((m, n), dtype=dt, device=dev)
    W = torch.empty((m, n), dtype=dt, device=dev)
    w = torch.empty((n,), dtype=dt, device=dev)
    h = torch.zeros((16,), dtype=dt, device=dev)
    eps23 = np.finfo(np.float64).eps ** (2.0 / 3.0)
    sign = 1.0 if which in ("SA", "SR") else -1.0

    nrm = _norm_c(v0, comm)
    V[0].copy_(v0.t)
    ops.scale_(Array(V[0]), 1.0, div_by=nrm)
    have_w0 = False
    nmv = 0
    theta, resid = None, None
    info = {"restarts": 0, "converged": False}
    x = Array(V[0])
    mmax = m
    for cycle in range(maxiter):
        # adaptive basis: a short first cycle (an already good v0 converges
        # in ~3 matvecs, like ARPACK's ncv=4), longer ones for hard problems
        m = min(mmax, 4 << min(cycle, 4))
        for j in range(m):
            if not (j == 0 and have_w0):
                Wj = matvec(Array(V[j]))
                nmv += 1
                W[j].copy_(ops.materialize(Wj).t.reshape(-1))
            if j + 1 < m:
                w.copy_(W[j])
                _orthogonalise(V, j, w, h, comm)
                bnorm = _norm_c(Array(w), comm)
                V[j + 1].copy_(w)
                ops.scale_(Array(V[j + 1]), 1.0, div_by=bnorm)
        # projected matrix (m x m): host read #1 of the cycle
        Hm = contract_pair(V[:m], [_J, _N], W[:m], [2, _N], [_J, 2])
        if comm is not None:
            comm.all_reduce_(Hm)
        Hh = Hm.cpu().numpy()
        Hh = 0.5 * (Hh + Hh.T)
        evals, evecs = np.linalg.eigh(sign * Hh)
        theta = sign * evals[0]
        y = evecs[:, 0]
        yd = torch.zeros(16, dtype=dt)
        yd[:m] = torch.as_tensor(y, dtype=dt)
        yd = yd.to(dev)
        lib = _lib.load()
        xnew = torch.zeros((n,), dtype=dt, device=dev)
        hx = torch.zeros((n,), dtype=dt, device=dev)
        st = _lib.stream_ptr()
        _lib.check(lib.qb_multi_axpy(_lib.QB_F64, m, n, V.data_ptr(), V.stride(0),
                                     yd.data_ptr(), 1.0, xnew.data_ptr(), st),
                   "qb_multi_axpy")
        _lib.check(lib.qb_multi_axpy(_lib.QB_F64, m, n, W.data_ptr(), W.stride(0),
                                     yd.data_ptr(), 1.0, hx.data_ptr(), st),
                   "qb_multi_axpy")
        # true residual |H x - theta x| (robust to Krylov breakdown):
        # host read #2
        w.copy_(hx)
        ops.axpby(-theta, Array(xnew), 1.0, Array(w))
        resid = float(_norm_c(Array(w), comm).item())
        info["restarts"] = cycle
        if resid <= tol * max(eps23, abs(theta)) or not np.isfinite(resid):
            info["converged"] = bool(np.isfinite(resid))
            x = Array(xnew)
            break
        # restart from the Ritz vector; its image is known
        xn = _norm_c(Array(xnew), comm)
        V[0].copy_(xnew)
        ops.scale_(Array(V[0]), 1.0, div_by=xn)
        W[0].copy_(hx)
        ops.scale_(Array(W[0]), 1.0, div_by=xn)
        have_w0 = True
        x = Array(V[0].clone())
    # normalise the Ritz vector
    xn = _norm_c(x, comm)
    x = ops.scale_(ops.materialize(x, force=True), 1.0, div_by=xn)
    info.update(nmatvec=nmv, resid=resid, theta=theta, ncv=mmax)
    if return_info:
        return float(theta), x, info
    return float(theta), x


def eigh_arpack_host_driver(matvec, v0, which="SA", ncv=4, tol=1e-3):
    """Parity mode: the reference's own eigensolver (scipy ARPACK on the
    host) driving the device matvec; every iteration copies the vector
    host->device->host.  For validation only."""
    import scipy.sparse.linalg as spla

    v0h = ops.to_numpy(ops.asarray(v0)).reshape(-1)
    n = v0h.size
    count = [0]

    def mv(vec):
        count[0] += 1
        out = matvec(ops.asarray(np.ascontiguousarray(vec)))
        return ops.to_numpy(out).reshape(-1)

    A = spla.LinearOperator((n, n), matvec=mv, dtype=v0h.dtype)
    lk, vk = spla.eigsh(A, k=1, which=which, v0=v0h, ncv=ncv, tol=tol)
    return float(lk[0]), ops.asarray(vk[:, 0]), {"nmatvec": count[0]}
