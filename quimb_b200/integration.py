"""Plug the backend into an installed quimb / autoray.

quimb never needs to be patched: its arrays-with-``.shape`` rule
(quimb/tensor/array_ops.py:31-33) keeps ``quimb_b200.Array`` objects as
``Tensor._data`` and autoray resolves every ``do(name, x)`` to
``getattr(quimb_b200, name)`` from the array's module.  What quimb *adds* on
top of plain array functions are composed drivers with per-backend overrides
(``@svd_truncated.register("numpy")`` at decomp.py:1058, ``qr_stabilized``
:2198, ``fuse`` / ``norm_fro`` in array_ops.py); this function registers the
fused device versions for backend ``"quimb_b200"`` the same way.  See
INTEGRATION.md for the maintainer-side view.
"""

from . import ops, split
from .array import Array


def to_device(x, dtype=None):
    """``Tensor.apply_to_arrays`` / ``tn.apply_to_arrays`` helper: numpy ->
    device Array."""
    return ops.asarray(x, dtype=dtype)


def register_with_quimb():
    """Register the fused split drivers with quimb's composed functions.
    Returns the list of names registered; raises ImportError if quimb (and
    its autoray / cotengra dependencies) are not importable."""
    import autoray as ar
    from quimb.tensor import array_ops, decomp

    name = "quimb_b200"
    done = []

    def _svd_truncated(x, cutoff=-1.0, cutoff_mode=4, max_bond=-1, absorb=0,
                       renorm=0, info=None, **kwargs):
        return split.svd_truncated(x, cutoff=cutoff, cutoff_mode=cutoff_mode,
                                   max_bond=max_bond, absorb=absorb,
                                   renorm=renorm, info=info)

    def _qr_stabilized(x, absorb=1, stabilized=True, **kwargs):
        return split.qr_stabilized(x, absorb=absorb, stabilized=stabilized)

    decomp.svd_truncated.register(name)(_svd_truncated)
    done.append("svd_truncated")
    decomp.qr_stabilized.register(name)(_qr_stabilized)
    done.append("qr_stabilized")

    def _svd_via_eig_truncated(x, cutoff=-1.0, cutoff_mode=4, max_bond=-1, absorb=0,
                               renorm=0, info=None, **kwargs):
        return split.svd_via_eig_truncated(x, cutoff=cutoff, cutoff_mode=cutoff_mode,
                                           max_bond=max_bond, absorb=absorb,
                                           renorm=renorm, info=info)

    for nm, fn in (("svd_via_eig_truncated", _svd_via_eig_truncated),
                   ("svd_rand_truncated", split.svd_rand_truncated),
                   ("eigh_truncated", split.eigh_truncated),
                   ("cholesky_regularized", split.cholesky_regularized),
                   ("qr_via_cholesky", split.qr_via_cholesky),
                   ("polar_right", split.polar_right),
                   ("polar_left", split.polar_left),
                   ("lu_truncated", split.lu_truncated),
                   ("rddiv", split.rddiv), ("lddiv", split.lddiv),
                   ("rdmul", split.rdmul), ("ldmul", split.ldmul),
                   ("sgn", split.sgn)):
        if hasattr(decomp, nm) and hasattr(getattr(decomp, nm), "register"):
            getattr(decomp, nm).register(name)(fn)
            done.append(nm)

    for nm in ("fuse", "unfuse", "multiply_diagonal", "align_axes"):
        fn = getattr(array_ops, nm, None)
        if fn is not None and hasattr(fn, "register"):
            fn.register(name)(getattr(ops, nm))
            done.append(nm)

    def _norm_fro(x):
        from .linalg import norm
        return norm(x)

    if hasattr(array_ops, "norm_fro") and hasattr(array_ops.norm_fro, "register"):
        array_ops.norm_fro.register(name)(_norm_fro)
        done.append("norm_fro")
    ar.register_function(name, "to_numpy", ops.to_numpy)
    done.append("to_numpy")

    # singular-value-only drivers are plain (numba) functions in quimb, looked
    # up by method name (decomp.py:474-493): wrap them so device arrays take
    # the device routes and everything else still reaches the original
    for method, fn in (("svd", split.svdvals), ("svd:eig", split.svdvals_eig)):
        orig = decomp._SPLIT_VALUES_FNS.get(method)
        if orig is not None and not getattr(orig, "_quimb_b200", False):
            def _svals(x, *a, _orig=orig, _fn=fn, **kw):
                return _fn(x) if isinstance(x, Array) else _orig(x, *a, **kw)
            _svals._quimb_b200 = True
            decomp._SPLIT_VALUES_FNS[method] = _svals
            done.append(f"svals:{method}")

    # partial eigensolver backend: ``eigh(A, k=1, backend="quimb_b200")`` /
    # ``dmrg.opts["local_eig_backend"] = "quimb_b200"`` (one dictionary entry
    # next to "NUMPY" / "SCIPY" / "LOBPCG", quimb/linalg/base_linalg.py:70-77)
    from quimb.linalg import base_linalg
    base_linalg._EIGS_METHODS[name.upper()] = eigs_quimb_b200
    done.append("eigs:QUIMB_B200")
    return done


# ``tol=None`` / 0 means machine precision in the reference's backend contract
# (scipy_linalg.py:109 passes 0 to ARPACK); DMRG passes its own 1e-3
# (dmrg.py:87-88).  The Lanczos residual estimate cannot go below round-off of
# the matvec, so "machine precision" is a relative residual of 1e-12 here.
_EIGS_TIGHT_TOL = 1e-12


def eigs_quimb_b200(A, k=1, *, B=None, which="SA", return_vecs=True, sigma=None,
                    isherm=True, ncv=None, sort=True, tol=None, v0=None,
                    maxiter=None, **backend_opts):
    """quimb partial-eigensolver backend (same contract as ``eigs_scipy``,
    quimb/linalg/scipy_linalg.py:23-133: returns ``lk`` as a host array of
    ``k`` values and ``vk`` with eigenvectors as columns) that keeps the
    whole Krylov process on the device.  ``A`` is a quimb
    ``TNLinearOperator`` over device arrays (its ``_matvec`` accepts and
    returns device arrays, tensor_core.py:12393-12417) or a dense Hermitian
    device array (DMRG's dense branch, dmrg.py:690-703)."""
    import numpy as np
    from . import linalg
    from .lanczos import eigh_lanczos
    if B is not None or sigma is not None or not isherm:
        raise NotImplementedError("quimb_b200 eigensolver: standard Hermitian problems only")
    if which not in ("SA", "LA"):
        raise NotImplementedError(f"quimb_b200 eigensolver: which={which!r}")
    if hasattr(A, "_matvec") and not isinstance(A, Array):
        if k != 1:
            raise NotImplementedError("quimb_b200 eigensolver: k=1 for linear operators")
        n = A.shape[1]
        if v0 is None:
            v0 = np.random.default_rng(0).standard_normal(n).astype(
                np.dtype(A.dtype), copy=False)
        v0 = ops.asarray(v0)
        theta, x = eigh_lanczos(lambda v: ops.asarray(A._matvec(v)), v0, which=which,
                                ncv=max(2, min(64, ncv or 4)),
                                tol=_EIGS_TIGHT_TOL if not tol else tol, maxiter=maxiter)
        lk = np.array([theta])
        if not return_vecs:
            return lk
        return lk, Array(x.resolve().reshape(-1, 1))
    w, v = linalg.eigh(ops.asarray(A))
    wt, vt = w.resolve(), v.resolve()
    if which == "LA":
        wt, vt = wt.flip(0), vt.flip(1)
    lk = wt[:k].cpu().numpy()
    if not return_vecs:
        return lk
    return lk, Array(vt[:, :k].contiguous())


__all__ = ["register_with_quimb", "to_device", "Array"]
