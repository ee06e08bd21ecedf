"""GPU tier: the device TNLinearOperator (matvec / matmat / conj / adjoint /
transpose / trace / to_dense) against dense numpy contractions -- the way the
reference tests it (tests/test_tensor/test_tensor_core.py:2181-2202) -- and as
the operator of the device Lanczos solver."""

import numpy as np
import pytest

import quimb_b200 as qb
from quimb_b200.linop import TNLinearOperator
from oracle import dmrg_np as dm

pytestmark = pytest.mark.gpu


def _net(rng, dtype):
    shapes = {"A": (4, 6, 3), "B": (6, 5, 7), "C": (7, 2)}
    inds = {"A": ("a", "x", "b"), "B": ("x", "c", "y"), "C": ("y", "d")}
    arrs = {}
    for k, s in shapes.items():
        v = rng.standard_normal(s)
        if dtype == "complex128":
            v = v + 1j * rng.standard_normal(s)
        arrs[k] = v.astype(dtype)
    return [arrs[k] for k in "ABC"], [inds[k] for k in "ABC"]


@pytest.mark.parametrize("dtype", ["float64", "complex128"])
def test_linop_against_dense(dtype):
    rng = np.random.default_rng(0)
    arrays, inds = _net(rng, dtype)
    dense = np.einsum("axb,xcy,yd->acbd", *arrays).reshape(20, 6)
    op = TNLinearOperator(arrays, inds, ("a", "c"), ("b", "d"))
    assert op.shape == (20, 6) and op.ldims == (4, 5) and op.rdims == (3, 2)
    v = rng.standard_normal(6) + (1j * rng.standard_normal(6) if dtype == "complex128" else 0)
    w = rng.standard_normal(20) + (1j * rng.standard_normal(20) if dtype == "complex128" else 0)
    M = rng.standard_normal((6, 3)).astype(dtype)
    np.testing.assert_allclose(op.matvec(qb.asarray(v)).to_numpy(), dense @ v, atol=1e-12)
    np.testing.assert_allclose((op @ qb.asarray(M)).to_numpy(), dense @ M, atol=1e-12)
    np.testing.assert_allclose(op.conj().matvec(qb.asarray(v)).to_numpy(),
                               dense.conj() @ v, atol=1e-12)
    np.testing.assert_allclose(op.H.matvec(qb.asarray(w)).to_numpy(),
                               dense.conj().T @ w, atol=1e-12)
    np.testing.assert_allclose(op.T.matvec(qb.asarray(w)).to_numpy(), dense.T @ w, atol=1e-12)
    np.testing.assert_allclose(op.rmatvec(qb.asarray(w)).to_numpy(),
                               dense.conj().T @ w, atol=1e-12)
    np.testing.assert_allclose(op.to_dense().to_numpy(), dense, atol=1e-12)
    np.testing.assert_allclose(op.conj().A.to_numpy(), dense.conj(), atol=1e-12)
    assert op.nmatvec == 1
    # the expression is cached: the second call reuses the tree
    assert set(op._contractors) == {"matvec", "matmat_3"}


def test_linop_trace_and_split():
    rng = np.random.default_rng(1)
    a = rng.standard_normal((3, 4, 5, 3))
    b = rng.standard_normal((5, 4))
    op = TNLinearOperator([a, b], [("i", "j", "k", "l"), ("k", "m")], ("i", "j"), ("l", "m"))
    dense = np.einsum("ijkl,km->ijlm", a, b).reshape(12, 12)
    assert abs(op.trace() - np.trace(dense)) < 1e-11
    left, _, right = op.split(method="svd", cutoff=0.0, absorb="both")
    np.testing.assert_allclose(left.to_numpy() @ right.to_numpy(), dense, atol=1e-11)


def test_one_site_effective_hamiltonian_lanczos():
    """L - W - R one-site effective Hamiltonian as a TNLinearOperator driving
    the device Lanczos: lowest eigenvalue vs dense eigh."""
    L = 6
    mpo = dm.mpo_heis(L)
    sites = dm.mps_rand(L, 5, seed=2)
    s = [np.asarray(x) for x in sites]
    dm.right_canonize(s)
    i = 2
    E = np.ones((1, 1, 1))
    for k in range(i):
        E = dm.env_step_left(E, s[k], mpo[k])
    R = np.ones((1, 1, 1))
    for k in range(L - 1, i, -1):
        R = dm.env_step_right(R, s[k], mpo[k])
    W = mpo[i]                                    # (l, r, u, d)
    op = TNLinearOperator([E, W, R], [("ap", "w", "a"), ("w", "v", "p", "q"), ("bp", "v", "b")],
                          ("ap", "q", "bp"), ("a", "p", "b"))
    dense = np.einsum("xwa,wvpq,yvb->xqyapb", E, W, R).reshape(op.shape)
    np.testing.assert_allclose(op.to_dense().to_numpy(), dense, atol=1e-12)
    v0 = np.random.default_rng(3).standard_normal(op.shape[1])
    theta, x = qb.eigh_lanczos(op, qb.asarray(v0), ncv=8, tol=1e-10)
    ref = np.linalg.eigvalsh(0.5 * (dense + dense.T))[0]
    assert abs(theta - ref) < 1e-8


def test_fuse_unfuse_values():
    from oracle import decomp_np as dn
    rng = np.random.default_rng(4)
    x = rng.standard_normal((3, 4, 5, 2, 6)) + 1j * rng.standard_normal((3, 4, 5, 2, 6))
    xa = qb.asarray(x)
    for groups in [((3, 1), (4, 0)), ((0, 1),), ((2,), (4, 3)), ((1, 0, 2, 3, 4),)]:
        f = qb.fuse(xa, *groups)
        np.testing.assert_array_equal(f.to_numpy(), dn.fuse(x, *groups))
    f = qb.fuse(xa.conj(), (0, 1), (2, 3))
    np.testing.assert_array_equal(f.to_numpy(), dn.fuse(x.conj(), (0, 1), (2, 3)))
    u = qb.unfuse(qb.fuse(xa, (1, 2)), 1, (4, 5))
    np.testing.assert_array_equal(u.to_numpy(), x)
    assert qb.fuse(xa) is xa and qb.unfuse(xa, 0, (3,)) is xa
