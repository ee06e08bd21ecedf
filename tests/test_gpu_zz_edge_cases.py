"""GPU tier, last file on purpose: edge cases and callers added after the
round's GPU budget was spent (validated against the oracle / reference golden
values on the CPU tier through tests/abi_emulator.py, not yet on a device).
Rank-deficient complex QR, Krylov breakdown, redundant MPS bonds, MPS
compression / circuit driver, einsum / matmul fuzz, degenerate split inputs."""

import itertools

import numpy as np
import pytest

import quimb_b200 as qb
from quimb_b200 import mps, tebd as tb
from oracle import dmrg_np as dm

pytestmark = pytest.mark.gpu


def _np(x):
    return None if x is None else x.to_numpy()


def _dense(sites):
    return dm.mps_to_dense([np.asarray(s.to_numpy()) for s in sites]).reshape(-1)


def _fresh(raw):
    n = len(raw)
    return [qb.materialize(mps.site_lpr(a, "lrp", i, n), force=True)
            for i, a in enumerate(raw)]


@pytest.mark.parametrize("m,n,r", [(60, 30, 8), (32, 30, 18), (20, 40, 5), (96, 96, 40)])
def test_complex_qr_rank_deficient(m, n, r):
    """Complex QR of numerically rank-deficient input (redundant MPS bonds,
    products of thin factors): the real-embedding route is only valid through
    uniqueness of the full-rank QR, so such input takes the SVD route: x = Q R
    with a complete isometry Q (R is then not triangular)."""
    rng = np.random.default_rng(m + n + r)
    x = ((rng.standard_normal((m, r)) + 1j * rng.standard_normal((m, r)))
         @ (rng.standard_normal((r, n)) + 1j * rng.standard_normal((r, n))))
    Q, R = qb.linalg.qr(qb.asarray(x), stabilized=True)
    q, rr = _np(Q), _np(R)
    k = min(m, n)
    assert q.shape == (m, k) and rr.shape == (k, n)
    scale = np.linalg.norm(x, 2)
    np.testing.assert_allclose(q.conj().T @ q, np.eye(k), atol=1e-11)
    np.testing.assert_allclose(q @ rr, x, atol=1e-11 * scale)
    # the LQ family and the split driver inherit it
    left, _, right = qb.qr_stabilized(qb.asarray(x), absorb="left")
    l, rt = _np(left), _np(right)
    np.testing.assert_allclose(l @ rt, x, atol=1e-11 * scale)
    np.testing.assert_allclose(rt @ rt.conj().T, np.eye(k), atol=1e-11)
    # ill-conditioned but full rank (cond 1e8): still an isometry (to the
    # accuracy of the singular subspaces of the tiniest values: 1e-8 covers a
    # LAPACK-grade SVD; the Jacobi kernel is relatively accurate and far better)
    u, _ = np.linalg.qr(rng.standard_normal((m, k)) + 1j * rng.standard_normal((m, k)))
    v, _ = np.linalg.qr(rng.standard_normal((n, k)) + 1j * rng.standard_normal((n, k)))
    y = (u * np.logspace(0, -8, k)) @ v.conj().T
    Q, R = qb.linalg.qr(qb.asarray(y), stabilized=True)
    q, rr = _np(Q), _np(R)
    np.testing.assert_allclose(q.conj().T @ q, np.eye(k), atol=1e-8)
    np.testing.assert_allclose(q @ rr, y, atol=1e-12)


def test_mps_circuit_simulation_matches_statevector():
    """CircuitMPS-style run (quimb/tensor/circuit/mps.py): one-qubit gates
    contracted in, two-qubit gates (incl. long-range ones) by swap + split;
    amplitudes against a dense state-vector simulation."""
    from tests.circuit_util import _rand_u2, _rand_u4
    rng = np.random.default_rng(12)
    n, depth = 7, 5
    gates = []
    psi = np.zeros([2] * n, dtype=np.complex128)
    psi[(0,) * n] = 1.0
    for layer in range(depth):
        for q in range(n):
            u = _rand_u2(rng)
            gates.append((u, (q,)))
            psi = np.moveaxis(np.tensordot(u, psi, axes=(1, q)), 0, q)
        pairs = [(0, 1), (2, 5), (6, 3)] if layer % 2 == 0 else [(1, 2), (4, 0), (5, 6)]
        for a, b in pairs:
            g = _rand_u4(rng)
            gates.append((g.reshape(4, 4), (a, b)))
            psi = np.moveaxis(np.tensordot(g, psi, axes=((2, 3), (a, b))), (0, 1), (a, b))
    sites = tb.mps_zero_state(n)
    tb.apply_circuit(sites, gates, cutoff=1e-14)
    np.testing.assert_allclose(_dense(sites), psi.reshape(-1), atol=1e-10)
    for bits in ([0] * n, [1, 0, 1, 1, 0, 0, 1], rng.integers(0, 2, n).tolist()):
        assert abs(tb.mps_amplitude(sites, bits) - psi[tuple(bits)]) < 1e-10
    # truncated run stays normalised to the discarded weight
    s2 = tb.mps_zero_state(n)
    tb.apply_circuit(s2, gates, max_bond=4, cutoff=0.0)
    assert max(a.shape[2] for a in s2) <= 4
    ov = abs(np.vdot(psi.reshape(-1), _dense(s2)))
    assert 0.3 < ov <= 1.0 + 1e-12


def test_mps_compress_add_apply_overlap_match_reference(golden_mps_ops):
    """MatrixProductState.compress (all forms) / add_MPS / MPO.apply / overlap
    of the reference (tests/golden/mps_ops.*): dense states and bond dims."""
    data, meta = golden_mps_ops
    n = 7
    p = _fresh([data[f"p__{i}"] for i in range(n)])
    q = _fresh([data[f"q__{i}"] for i in range(n)])
    H = [data[f"H__{i}"] for i in range(n)]
    np.testing.assert_allclose(_dense(p), data["p_dense"], atol=1e-13)
    ov = tb.mps_overlap(p, q)
    assert abs(ov - complex(*meta["overlap_pq"])) < 1e-12
    np.testing.assert_allclose(_dense(tb.mps_add(p, q)), data["add_dense"], atol=1e-12)
    Hp = tb.mpo_apply(H, p, mpo_shape="lrud")
    assert [a.shape[2] for a in Hp[:-1]] == meta["Hp_bonds"]
    np.testing.assert_allclose(_dense(Hp), data["Hp_dense"], atol=1e-12)
    for c in meta["compress"]:
        s = [a.copy() for a in Hp]
        tb.mps_compress(s, **c["kw"])
        assert [a.shape[2] for a in s[:-1]] == c["bonds"], c
        tol = 1e-7 if c["kw"].get("method") == "svd:eig" else 1e-10
        np.testing.assert_allclose(_dense(s), data[c["key"] + "__dense"], atol=tol,
                                   err_msg=str(c))
    # canonical forms: 'right' leaves every site but the first right-isometric
    s = [a.copy() for a in Hp]
    tb.mps_compress(s, form="right", max_bond=5, cutoff=0.0)
    for a in s[1:]:
        m = a.to_numpy().reshape(a.shape[0], -1)
        np.testing.assert_allclose(m @ m.conj().T, np.eye(m.shape[0]), atol=1e-11)
    with pytest.raises(ValueError):
        tb.mps_compress(s, form="up")


def test_compress_and_canonize_complex_mps_with_redundant_bonds():
    """Zero-padded (rank-deficient) complex bonds: canonisation keeps the
    state and compression finds the true bond dimensions (the complex QR takes
    its SVD route on such input)."""
    a = dm.mps_rand(6, 3, seed=5, dtype="complex128")
    pad = []
    for i, x in enumerate(a):
        l, d, r = x.shape
        y = np.zeros((l if i == 0 else l + 4, d, r if i == 5 else r + 4), dtype=complex)
        y[:l, :, :r] = x
        pad.append(y)
    ref = dm.mps_to_dense(a).reshape(-1)
    true_bonds = [x.shape[2] for x in a]
    for form in ("right", "left", 3, "flat"):
        s = [qb.asarray(x) for x in pad]
        tb.mps_compress(s, form=form, cutoff=1e-12)
        np.testing.assert_allclose(_dense(s), ref, atol=1e-12)
        assert [x.shape[2] for x in s] == true_bonds
    s = [qb.asarray(x) for x in pad]
    tb.canonicalize(s, 2)
    np.testing.assert_allclose(_dense(s), ref, atol=1e-12)


@pytest.mark.parametrize("dtype", ["float64", "complex128"])
def test_dmrg2_from_product_state_krylov_breakdown(dtype):
    """A product state start makes the local Krylov spaces tiny: the Lanczos
    basis breaks down exactly (beta = 0) -- also inside the steps queued ahead
    of the host reads -- and must end cleanly at the invariant subspace."""
    L = 8
    mpo = dm.mpo_heis(L)
    p0 = [np.zeros((1, 2, 1), dtype=dtype) for _ in range(L)]
    for i in range(L):
        p0[i][0, i % 2, 0] = 1.0
    d = qb.DMRG2(mpo, [4, 8, 16], cutoffs=1e-10, mpo_shape="lrdu", p0=p0, mps_shape="lpr")
    d.solve(tol=1e-8, max_sweeps=8)
    assert d.state[0].dtype == np.dtype(dtype)
    assert abs(d.energy - np.linalg.eigvalsh(dm.mpo_to_dense(mpo))[0]) < 1e-7
    # exact eigenvector as the start vector: immediate breakdown, no NaNs
    rng = np.random.default_rng(0)
    q, _ = np.linalg.qr(rng.standard_normal((50, 50)))
    lam = np.linspace(-3, 3, 50)
    H = qb.asarray((q * lam) @ q.T)
    theta, x, info = qb.eigh_lanczos(lambda v: qb.tensordot(H, v, axes=((1,), (0,))),
                                     qb.asarray(q[:, 0].copy()), ncv=16, tol=1e-12,
                                     return_info=True)
    assert abs(theta - lam[0]) < 1e-12 and np.all(np.isfinite(x.to_numpy()))
    # start vector inside a 3-dimensional invariant subspace
    v0 = q[:, [0, 7, 20]] @ np.array([0.2, 1.0, -0.5])
    theta, x, info = qb.eigh_lanczos(lambda v: qb.tensordot(H, v, axes=((1,), (0,))),
                                     qb.asarray(v0), ncv=16, tol=1e-12, return_info=True)
    assert abs(theta - lam[0]) < 1e-10 and info["nmatvec"] <= 8


def _rnd(rng, shape, dtype):
    x = rng.standard_normal(shape)
    if "complex" in dtype:
        x = x + 1j * rng.standard_normal(shape)
    return x.astype(dtype)


def test_einsum_fuzz_against_numpy():
    """150 random two-operand einsum signatures: batch / contracted / summed /
    repeated (diagonal) labels, size-1 modes, transposed and conjugated views,
    all four dtypes."""
    rng = np.random.default_rng(1)
    letters = "abcdefgh"
    for trial in range(150):
        dtype = str(rng.choice(["float64", "complex128", "float32", "complex64"]))
        sz = {c: int(rng.choice([1, 2, 3, 4])) for c in letters}
        ra, rb = int(rng.integers(0, 5)), int(rng.integers(0, 5))
        ta = "".join(rng.choice(list(letters), size=ra, replace=bool(rng.random() < 0.2)))
        tb_ = "".join(rng.choice(list(letters), size=rb, replace=bool(rng.random() < 0.2)))
        cand = list(dict.fromkeys(ta + tb_))
        keep = [c for c in cand if rng.random() < 0.6]
        tc = "".join(rng.permutation(keep)) if keep else ""
        a = _rnd(rng, [sz[c] for c in ta], dtype)
        b = _rnd(rng, [sz[c] for c in tb_], dtype)
        A, B = qb.asarray(a), qb.asarray(b)
        if ra and rng.random() < 0.5:
            pa = list(rng.permutation(ra))
            A = qb.asarray(np.ascontiguousarray(np.transpose(a, pa))).transpose(*np.argsort(pa))
        if rng.random() < 0.3:
            A, a = A.conj(), a.conj()
        eq = f"{ta},{tb_}->{tc}"
        out = qb.einsum(eq, A, B).to_numpy()
        ref = np.einsum(eq, a, b)
        tol = 1e-4 if dtype in ("float32", "complex64") else 1e-11
        assert out.shape == ref.shape and out.dtype == ref.dtype, eq
        assert np.abs(out - ref).max(initial=0) <= tol * max(1, np.abs(ref).max(initial=0)), eq


def test_tensordot_and_matmul_numpy_semantics():
    rng = np.random.default_rng(2)
    for trial in range(60):
        dtype = str(rng.choice(["float64", "complex128"]))
        na, nb = int(rng.integers(1, 5)), int(rng.integers(1, 5))
        k = int(rng.integers(0, min(na, nb) + 1))
        sa = [int(rng.integers(1, 4)) for _ in range(na)]
        sb = [int(rng.integers(1, 4)) for _ in range(nb)]
        axa = [int(v) for v in rng.choice(na, size=k, replace=False)]
        axb = [int(v) for v in rng.choice(nb, size=k, replace=False)]
        for i, j in zip(axa, axb):
            sb[j] = sa[i]
        a, b = _rnd(rng, sa, dtype), _rnd(rng, sb, dtype)
        ref = np.tensordot(a, b, axes=(axa, axb))
        out = qb.tensordot(qb.asarray(a), qb.asarray(b), axes=(axa, axb)).to_numpy()
        assert out.shape == ref.shape and np.abs(out - ref).max(initial=0) < 1e-11
        out = qb.tensordot(qb.asarray(a), qb.asarray(b),
                           axes=([x - na for x in axa], [x - nb for x in axb])).to_numpy()
        assert np.abs(out - ref).max(initial=0) < 1e-11
    # matmul: vector promotion, batch dims, size-1 broadcasting
    for sa, sb in [((3, 4), (4, 5)), ((4,), (4, 5)), ((3, 4), (4,)), ((4,), (4,)),
                   ((2, 3, 4), (2, 4, 5)), ((2, 3, 4), (4, 5)), ((3, 4), (2, 4, 5)),
                   ((1, 3, 4), (2, 4, 5)), ((2, 3, 4), (4,)), ((4,), (2, 4, 5)),
                   ((5, 1, 3, 4), (2, 4, 6)), ((2, 1, 3, 4), (1, 5, 4, 2))]:
        a, b = _rnd(rng, sa, "float64"), _rnd(rng, sb, "float64")
        out = qb.asarray(a) @ qb.asarray(b)
        ref = a @ b
        assert out.shape == np.shape(ref), (sa, sb)
        np.testing.assert_allclose(out.to_numpy(), ref, atol=1e-12)
    with pytest.raises(ValueError):
        qb.asarray(np.zeros((2, 3, 4))) @ qb.asarray(np.zeros((3, 4, 5)))
    x = _rnd(rng, (3, 4, 3, 5), "complex128")
    for ax in [(0, 2), (2, 0)]:
        np.testing.assert_allclose(qb.trace(qb.asarray(x), axis1=ax[0], axis2=ax[1]).to_numpy(),
                                   np.trace(x, axis1=ax[0], axis2=ax[1]), atol=1e-12)


@pytest.mark.parametrize("kind", ["full", "lowrank", "rank1", "zero"])
@pytest.mark.parametrize("dtype", ["float64", "complex128", "float32", "complex64"])
def test_split_drivers_on_degenerate_matrices(kind, dtype):
    """Every split method x absorb mode on 1x1, single-row / -column, square,
    tall and wide matrices that are full rank, low rank, rank one or zero:
    finite factors, the input dtype, and left @ right == x."""
    rng = np.random.default_rng(3)
    tol = 1e-4 if dtype in ("float32", "complex64") else 1e-9
    cplx = "complex" in dtype

    def g(*shape):
        v = rng.standard_normal(shape)
        return v + 1j * rng.standard_normal(shape) if cplx else v

    for m, n in [(1, 1), (1, 5), (5, 1), (2, 2), (3, 7), (7, 3), (8, 8), (17, 5)]:
        if kind == "zero":
            x = np.zeros((m, n))
        elif kind == "rank1":
            x = np.outer(g(m), g(n))
        elif kind == "lowrank":
            r = max(1, min(m, n) // 3)
            x = g(m, r) @ g(r, n)
        else:
            x = g(m, n)
        x = x.astype(dtype)
        for method, absorbs in [("svd", [None, "both", "left", "right", "s", "lorthog",
                                         "rorthog", "lfactor", "rfactor"]),
                                ("svd:eig", ["both", "left", "right", None]),
                                ("qr", ["right", "left", "lorthog", "rorthog", "lfactor",
                                        "rfactor"]),
                                ("svd:rand", ["both", "left", "right"])]:
            for absorb in absorbs:
                kw = {}
                if method == "svd:rand":
                    kw = dict(max_bond=min(m, n), seed=1)
                if method in ("svd", "svd:eig"):
                    kw = dict(cutoff=0.0)
                L, s, R = (_np(t) for t in qb.array_split(qb.asarray(x), method=method,
                                                          absorb=absorb, **kw))
                for t in (L, s, R):
                    assert t is None or np.all(np.isfinite(t)), (m, n, method, absorb)
                if L is not None:
                    assert L.dtype == np.dtype(dtype)
                if L is not None and R is not None:
                    rec = L @ (np.diag(s) @ R if s is not None else R)
                    etol = tol * max(1.0, np.abs(x).max()) * (1e3 if method == "svd:eig" else 1)
                    assert np.abs(rec - x).max() <= etol, (m, n, method, absorb)


@pytest.mark.parametrize("dtype,tol", [("float32", 2e-5), ("float64", 1e-8),
                                       ("complex64", 2e-5), ("complex128", 1e-8)])
def test_dmrg2_all_four_dtypes_preserved(dtype, tol):
    """The reference's test_dtypes (tests/test_tensor/test_tn1d/test_dmrg.py:
    290-300): the state keeps the Hamiltonian's dtype for f32 / f64 / c64 /
    c128; single precision runs its Krylov process in double."""
    mpo = dm.mpo_heis(8)
    e0 = np.linalg.eigvalsh(dm.mpo_to_dense(mpo))[0]
    d = qb.DMRG2([w.astype(dtype) for w in mpo], [8, 16], cutoffs=1e-8, mpo_shape="lrdu",
                 seed=1, dtype=dtype)
    d.solve(max_sweeps=3)
    assert {str(a.dtype) for a in d.state} == {dtype}
    assert abs(d.energy - e0) < tol * abs(e0)


def test_contraction_with_more_than_12_interleaved_modes_regroups():
    """A pair whose free / contracted axes interleave so that a group keeps
    more than 12 non-mergeable modes (rank-26 operands of compression drivers,
    found by the reference's test_tn1d/test_compress.py on device tensors) is
    regrouped once with the permute kernel instead of being refused."""
    from quimb_b200.contract import contract_pair, permute_contiguous
    rng = np.random.default_rng(5)
    r = 26
    x = rng.standard_normal((2,) * r)
    y = rng.standard_normal((2,) * 14)
    la = list(range(r))
    k = list(range(0, r, 2))[:13]                  # every other axis of x is contracted
    lb = k + [40]
    lc = [l for l in la if l not in k] + [40]
    perm = rng.permutation(len(lc))
    lc = [lc[p] for p in perm]
    ref = np.einsum(x, la, y, lb, lc, optimize=True)
    out = contract_pair(qb.asarray(x).t, la, qb.asarray(y).t, lb, lc)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-11, atol=1e-11)
    # into a caller-provided output, accumulating
    prev = rng.standard_normal(ref.shape)
    o = qb.asarray(prev.copy()).t
    contract_pair(qb.asarray(x).t, la, qb.asarray(y).t, lb, lc, out=o, alpha=2.0, beta=1.0)
    np.testing.assert_allclose(o.cpu().numpy(), 2.0 * ref + prev, rtol=1e-11, atol=1e-11)
    # the multi-pass permutation on its own
    z = rng.standard_normal((2,) * 22)
    order = [int(p) for p in rng.permutation(22)]
    got = permute_contiguous(qb.asarray(z).t, order)
    assert got.is_contiguous()
    np.testing.assert_array_equal(got.cpu().numpy(), np.transpose(z, order))


def test_materialize_of_a_high_rank_general_permutation():
    """fuse / to_dense of a 22-index tensor under a general axis permutation
    has more modes than one launch of the permute kernel takes (17-18): the copy
    is done in passes; lazy conjugation is applied on the way."""
    from quimb_b200 import ops
    from quimb_b200.contract import PERMUTE_MAX_MODES, permute_modes
    rng = np.random.default_rng(6)
    z = rng.standard_normal((2,) * 22) + 1j * rng.standard_normal((2,) * 22)
    perm = [int(p) for p in rng.permutation(22)]
    view = qb.asarray(z).transpose(*perm)
    assert permute_modes(view.t) > PERMUTE_MAX_MODES
    got = ops.materialize(view)
    assert got.t.is_contiguous() and not got.cj
    np.testing.assert_array_equal(got.to_numpy(), np.transpose(z, perm))
    got_c = ops.materialize(view.conj())
    np.testing.assert_array_equal(got_c.to_numpy(), np.transpose(z, perm).conj())
    # reshape (fuse) of the permuted view goes through the same route
    fused = ops.reshape(view, (2 ** 11, 2 ** 11))
    np.testing.assert_array_equal(fused.to_numpy(), np.transpose(z, perm).reshape(2 ** 11, 2 ** 11))
    # a contiguous tensor and a plain transpose stay single-launch
    assert permute_modes(qb.asarray(z).t) == 1
    assert permute_modes(qb.asarray(z.reshape(2 ** 11, 2 ** 11)).t.t()) == 2


@pytest.mark.parametrize("shape", [(300, 120), (128, 128), (64, 200), (257, 33)])
@pytest.mark.parametrize("kind", ["full", "lowrank"])
def test_fused_truncated_split_all_absorb_modes(shape, kind):
    """qb_svd_trunc through split.svd_truncated for every absorb mode, tall /
    square / wide, full and rank-deficient: the accumulation-free paths (left
    family directly, right family natively for tall input and through the
    transpose for square input) against the oracle's truncation of LAPACK's SVD
    (decomp.py:829-1118): kept rank, singular values, products and the
    isometry of whichever factor the mode leaves orthonormal."""
    from oracle import decomp_np as dn
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    m, n = shape
    if kind == "full":
        x = rng.standard_normal((m, n))
    else:
        r = min(m, n) // 3
        x = rng.standard_normal((m, r)) @ rng.standard_normal((r, n))
    for absorb in (None, 2, -12, -11, -10, -1, 0, 1, 10, 11, 12):
        for kw in (dict(cutoff=1e-3, cutoff_mode=4, max_bond=-1, renorm=0),
                   dict(cutoff=0.0, cutoff_mode=3, max_bond=min(m, n) // 2, renorm=0),
                   dict(cutoff=1e-2, cutoff_mode=3, max_bond=-1, renorm=2)):
            info, oinfo = {"error": None}, {"error": None}
            L, s, R = qb.svd_truncated(qb.asarray(x), absorb=absorb, info=info, **kw)
            Lo, so, Ro = dn.svd_truncated(x, absorb=absorb, info=oinfo, **kw)
            for got, ref in ((L, Lo), (s, so), (R, Ro)):
                assert (got is None) == (ref is None), (absorb, kw)
            if s is not None:
                np.testing.assert_allclose(s.to_numpy(), so, rtol=0, atol=1e-11 * max(so.max(), 1e-300))
            assert abs(info["error"] - oinfo["error"]) <= 1e-10 * max(1.0, oinfo["error"])
            scale = np.linalg.norm(x, 2)
            if L is not None and R is not None:
                mid = np.diag(s.to_numpy()) if s is not None else None
                rec = L.to_numpy() @ (mid @ R.to_numpy() if mid is not None else R.to_numpy())
                reco = Lo @ (np.diag(so) @ Ro if so is not None else Ro)
                assert np.abs(rec - reco).max() <= 1e-10 * scale, (absorb, kw)
            # gauge-free checks of single factors: Gram matrices
            if L is not None:
                assert L.shape == Lo.shape
                assert np.abs(L.to_numpy().T @ L.to_numpy() - Lo.T @ Lo).max() <= 1e-9 * max(1.0, scale ** 2), (absorb, kw)
            if R is not None:
                assert R.shape == Ro.shape
                assert np.abs(R.to_numpy() @ R.to_numpy().T - Ro @ Ro.T).max() <= 1e-9 * max(1.0, scale ** 2), (absorb, kw)


def test_zero_extent_contraction_into_a_strided_output():
    """K = 0 with beta = 0 zero-fills exactly the (strided, possibly 8-byte
    aligned) output view and nothing around it (the vector fill
    ignored strides)."""
    import torch
    buf = torch.full((6, 10), 7.0, dtype=torch.float64, device="cuda")
    out = buf[1:5, 1::2]                                   # strided, 8-byte aligned view
    a = torch.zeros((4, 0), dtype=torch.float64, device="cuda")
    b = torch.zeros((0, 5), dtype=torch.float64, device="cuda")
    qb.contract_pair(a, [0, 1], b, [1, 2], [0, 2], out=out)
    h = buf.cpu().numpy()
    assert np.all(h[1:5, 1::2] == 0.0)
    mask = np.ones_like(h, dtype=bool); mask[1:5, 1::2] = False
    assert np.all(h[mask] == 7.0)
    # complex, contiguous but offset by one element (16-byte aligned for c128)
    cb = torch.full((9,), 1 + 2j, dtype=torch.complex128, device="cuda")
    qb.contract_pair(torch.zeros((3, 0), dtype=torch.complex128, device="cuda"), [0, 1],
                     torch.zeros((0, 2), dtype=torch.complex128, device="cuda"), [1, 2], [0, 2],
                     out=cb[1:7].view(3, 2))
    hc = cb.cpu().numpy()
    assert np.all(hc[1:7] == 0) and hc[0] == 1 + 2j and np.all(hc[7:] == 1 + 2j)
