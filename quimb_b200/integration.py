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
                   ("eigh_truncated", split.eigh_truncated)):
        if hasattr(decomp, nm) and hasattr(getattr(decomp, nm), "register"):
            getattr(decomp, nm).register(name)(fn)
            done.append(nm)

    for nm in ("fuse", "unfuse"):
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
    return done


__all__ = ["register_with_quimb", "to_device", "Array"]
