"""On-device Lanczos for the DMRG local eigensolve.

The reference solves ``eigh(Heff, k=1, which='SA', v0=..., ncv=4, tol=1e-3)``
with scipy's ARPACK (quimb/tensor/tn1d/dmrg.py:626-645 ->
quimb/linalg/scipy_linalg.py:113-128): the Krylov vectors live on the host
and every iteration calls back into ``TNLinearOperator._matvec``.  With a
device matvec that would cross PCIe twice per iteration, so the whole Krylov
process lives on the device here:

  * restarted Lanczos with a small basis (``ncv``, like ARPACK's), full
    re-orthogonalisation by classical Gram-Schmidt applied twice, written as
    two skinny contractions ``h = V w`` / ``w -= V^T h`` on the pairwise
    kernel (alpha/beta form);
  * the images ``W_j = H v_j`` are kept, so the projected matrix
    ``V H V^T`` is one contraction, the Ritz vector's image ``H x`` comes for
    free and every restart saves one matvec;
  * two small device->host reads per restart cycle (the ncv x ncv projected
    matrix, then the true residual norm |H x - theta x|); the tiny dense
    eigenproblem is host control logic, as in ARPACK.

Convergence test as in ARPACK's dsaupd: ``resid <= tol * max(eps^(2/3),
|theta|)``.  ``parity mode`` (scipy ARPACK driving the device matvec through
host copies) is available as ``eigh_arpack_host_driver`` for validation.
"""

import numpy as np
import torch

from . import ops
from .array import Array
from .contract import contract_pair
from .linalg import norm as _norm

_J, _N = 0, 1


def eigh_lanczos(matvec, v0, which="SA", ncv=4, tol=1e-3, maxiter=None,
                 return_info=False):
    """Lowest ('SA') or highest ('LA') eigenpair of a Hermitian operator.

    Parameters
    ----------
    matvec : callable(Array[n]) -> Array[n]
        Device matvec on flat contiguous vectors.
    v0 : Array
        Start vector (any shape, flattened).
    Returns ``(theta: float, x: Array[n])`` (+ info dict).
    """
    v0 = ops.materialize(ops.asarray(v0)).reshape(-1)
    n = v0.size
    dt = v0.t.dtype
    if dt.is_complex:
        raise NotImplementedError("eigh_lanczos: complex operators are not "
                                  "implemented yet (no fallback)")
    dev = v0.t.device
    m = max(2, min(int(ncv), n))
    if maxiter is None:
        maxiter = max(10 * n, 300) if n < 30 else 300
    V = torch.zeros((m, n), dtype=dt, device=dev)
    W = torch.empty((m, n), dtype=dt, device=dev)
    w = torch.empty((n,), dtype=dt, device=dev)
    eps23 = np.finfo(np.float64).eps ** (2.0 / 3.0)
    sign = 1.0 if which in ("SA", "SR") else -1.0

    nrm = _norm(v0)
    V[0].copy_(v0.t)
    ops.scale_(Array(V[0]), 1.0, div_by=nrm)
    have_w0 = False
    nmv = 0
    theta, resid = None, None
    info = {"restarts": 0, "converged": False}
    x = Array(V[0])
    for cycle in range(maxiter):
        bnorm = None
        for j in range(m):
            if not (j == 0 and have_w0):
                Wj = matvec(Array(V[j]))
                nmv += 1
                W[j].copy_(ops.materialize(Wj).t.reshape(-1))
            w.copy_(W[j])
            Vj = V[: j + 1]
            for _ in range(2):  # CGS2
                h = contract_pair(Vj, [_J, _N], w, [_N], [_J])
                contract_pair(Vj, [_J, _N], h, [_J], [_N], out=w, alpha=-1.0,
                              beta=1.0)
            bnorm = _norm(Array(w))
            if j + 1 < m:
                V[j + 1].copy_(w)
                ops.scale_(Array(V[j + 1]), 1.0, div_by=bnorm)
        # projected matrix (m x m): host read #1 of the cycle
        Hm = contract_pair(V, [_J, _N], W, [2, _N], [_J, 2])
        Hh = Hm.cpu().numpy()
        Hh = 0.5 * (Hh + Hh.T)
        evals, evecs = np.linalg.eigh(sign * Hh)
        theta = sign * evals[0]
        y = evecs[:, 0]
        yd = torch.as_tensor(y, dtype=dt).to(dev)
        xnew = contract_pair(V, [_J, _N], yd, [_J], [_N])
        hx = contract_pair(W, [_J, _N], yd, [_J], [_N])
        # true residual |H x - theta x| (robust to Krylov breakdown):
        # host read #2
        w.copy_(hx)
        ops.axpby(-theta, Array(xnew), 1.0, Array(w))
        resid = float(_norm(Array(w)).item())
        info["restarts"] = cycle
        if resid <= tol * max(eps23, abs(theta)) or not np.isfinite(resid):
            info["converged"] = bool(np.isfinite(resid))
            x = Array(xnew)
            break
        # restart from the Ritz vector; its image is known
        xn = _norm(Array(xnew))
        V[0].copy_(xnew)
        ops.scale_(Array(V[0]), 1.0, div_by=xn)
        W[0].copy_(hx)
        ops.scale_(Array(W[0]), 1.0, div_by=xn)
        have_w0 = True
        x = Array(V[0].clone())
    # normalise the Ritz vector
    xn = _norm(x)
    x = ops.scale_(ops.materialize(x, force=True), 1.0, div_by=xn)
    info.update(nmatvec=nmv, resid=resid, theta=theta)
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
