"""``TNLinearOperator`` on the device: the array-level mirror of
quimb/tensor/tensor_core.py:12297-12549.

A linear operator defined by an uncontracted network of constant tensors with
open ``left_inds`` (output) and ``right_inds`` (input): ``matvec`` is a cached
contraction *expression* over (constants..., input vector) -- the tree is
found once per operator (and once per ``matmat`` width), constants stay
resident on the device, every step is one launch of the pairwise kernel.
Conjugation of the operator is a flag folded into the kernel loads of the
vector and of the result (:12396, :12414), never a pass over the constants.
``EffHam2`` (quimb_b200/dmrg.py) is the hand-scheduled special case for the
two-site DMRG problem; this class serves any network (one-site effective
Hamiltonians, environments, transfer operators, projected norms ...).
"""

import math

from . import ops
from .array import Array
from .tree import ContractExpression, tensor_contract


class TNLinearOperator:
    """Parameters
    ----------
    arrays : sequence of device arrays (or anything ``asarray`` accepts)
    inds : sequence of index tuples, one per array
    left_inds, right_inds : sequences of index names
        Output and input indices of the operator.
    ldims, rdims : tuples of int, optional
        Inferred from the arrays when omitted (tensor_core.py:12356-12366).
    optimize : str / path / tree, optional
    is_conj : bool
        The operator is the complex conjugate of the network.
    """

    def __init__(self, arrays, inds, left_inds, right_inds, ldims=None, rdims=None,
                 optimize="auto", is_conj=False):
        self._arrays = [ops.asarray(a) for a in arrays]
        self._inds = [tuple(t) for t in inds]
        if len(self._arrays) != len(self._inds):
            raise ValueError("one index tuple per array is required")
        self.left_inds, self.right_inds = tuple(left_inds), tuple(right_inds)
        if ldims is None or rdims is None:
            sz = {}
            for t, x in zip(self._inds, self._arrays):
                sz.update(zip(t, x.shape))
            ldims = tuple(sz[i] for i in self.left_inds)
            rdims = tuple(sz[i] for i in self.right_inds)
        self.ldims, self.rdims = tuple(ldims), tuple(rdims)
        self.shape = (math.prod(self.ldims), math.prod(self.rdims))
        self.dtype = self._arrays[0].dtype
        self.optimize = optimize
        self.is_conj = bool(is_conj)
        self._contractors = {}
        self._conj = self._adj = self._tr = None
        self.nmatvec = 0

    # ---- application ---------------------------------------------------------
    def _expr(self, key, in_inds, in_shape, out_inds):
        fn = self._contractors.get(key)
        if fn is None:
            n = len(self._arrays)
            fn = ContractExpression(
                self._inds + [tuple(in_inds)], tuple(out_inds),
                [a.shape for a in self._arrays] + [tuple(in_shape)],
                optimize=self.optimize,
                constants=dict(enumerate(self._arrays)))
            assert fn.var_pos == [n]
            self._contractors[key] = fn
        return fn

    def matvec(self, vec):
        """tensor_core.py:12393-12417."""
        self.nmatvec += 1
        x = ops.asarray(vec).reshape(*self.rdims)
        if x.dtype != self.dtype:
            x = x.astype(self.dtype, copy=False)
        if self.is_conj:
            x = x.conj()
        fn = self._expr("matvec", self.right_inds, self.rdims, self.left_inds)
        out = fn(x)
        if self.is_conj:
            out = out.conj()
        return ops.materialize(out).reshape(-1)

    __call__ = matvec

    def matmat(self, mat):
        """tensor_core.py:12419-12448."""
        mat = ops.asarray(mat)
        d = mat.shape[-1]
        x = mat.reshape(*self.rdims, d)
        if x.dtype != self.dtype:
            x = x.astype(self.dtype, copy=False)
        if self.is_conj:
            x = x.conj()
        fn = self._expr(f"matmat_{d}", (*self.right_inds, "_mat_ix"), (*self.rdims, d),
                        (*self.left_inds, "_mat_ix"))
        out = fn(x)
        if self.is_conj:
            out = out.conj()
        return ops.materialize(out).reshape(-1, d)

    def __matmul__(self, other):
        other = ops.asarray(other)
        return self.matvec(other) if other.ndim == 1 else self.matmat(other)

    def dot(self, other):
        return self.__matmul__(other)

    # ---- derived operators -----------------------------------------------------
    def copy(self, conj=False, transpose=False):
        if transpose:
            inds, dims = (self.right_inds, self.left_inds), (self.rdims, self.ldims)
        else:
            inds, dims = (self.left_inds, self.right_inds), (self.ldims, self.rdims)
        return TNLinearOperator(self._arrays, self._inds, *inds, *dims,
                                optimize=self.optimize,
                                is_conj=(not self.is_conj) if conj else self.is_conj)

    def conj(self):
        if self._conj is None:
            self._conj = self.copy(conj=True)
        return self._conj

    @property
    def T(self):
        if self._tr is None:
            self._tr = self.copy(transpose=True)
        return self._tr

    @property
    def H(self):
        if self._adj is None:
            self._adj = self.copy(conj=True, transpose=True)
        return self._adj

    def rmatvec(self, vec):
        return self.H.matvec(vec)

    def trace(self):
        """tensor_core.py:12450-12456: identify left and right indices pairwise
        and contract everything (no dense matrix is formed)."""
        if len(self.left_inds) != len(self.right_inds):
            raise ValueError("trace needs matching left and right indices")
        ren = dict(zip(self.right_inds, self.left_inds))
        inds = [tuple(ren.get(ix, ix) for ix in t) for t in self._inds]
        arrays = [a.conj() for a in self._arrays] if self.is_conj else self._arrays
        out, _ = tensor_contract(arrays, inds, output_inds=(), optimize=self.optimize)
        return out.item() if isinstance(out, Array) else out

    def to_dense(self):
        """Dense (prod(ldims), prod(rdims)) device matrix (tensor_core.py:
        12497-12514)."""
        arrays = [a.conj() for a in self._arrays] if self.is_conj else self._arrays
        out, _ = tensor_contract(arrays, self._inds,
                                 output_inds=self.left_inds + self.right_inds,
                                 optimize=self.optimize)
        return ops.materialize(out).reshape(*self.shape)

    toarray = to_dense

    @property
    def A(self):
        return self.to_dense()

    def astype(self, dtype):
        return TNLinearOperator([a.astype(dtype) for a in self._arrays], self._inds,
                                self.left_inds, self.right_inds, self.ldims, self.rdims,
                                optimize=self.optimize, is_conj=self.is_conj)

    def split(self, **split_opts):
        """tensor_split of the operator seen as a matrix (tensor_core.py:
        12519-12526): dense path through ``to_dense``."""
        from .split import array_split
        return array_split(self.to_dense(), **split_opts)
