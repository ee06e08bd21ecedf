"""GPU tier: QR / SVD kernels and the split drivers against the oracle and
the reference-generated golden vectors.  Parity on gauge-invariant outputs:
singular values, kept rank, truncation error, reconstruction, isometry; QR
factors element-wise (the stabilised QR is unique for full column rank)."""

import numpy as np
import pytest

import quimb_b200 as qb
from oracle import decomp_np as dn

pytestmark = pytest.mark.gpu


def _np(x):
    return None if x is None else x.to_numpy()


def test_qr_matches_reference_golden(golden_decomp):
    data, meta = golden_decomp
    for c in meta["qr_cases"]:
        x = data[f"mat__{c['mat']}"]
        left, _, right = qb.qr_stabilized(qb.asarray(x), absorb=c["absorb"])
        for part, nm in ((left, "__left"), (right, "__right")):
            key = c["key"] + nm
            assert (part is not None) == (key in data), c
            if part is not None:
                np.testing.assert_allclose(_np(part), data[key], atol=2e-12)


@pytest.mark.parametrize("m,n", [(8, 8), (33, 17), (200, 64), (2048, 96), (300, 300),
                                 (4100, 40), (64, 100)])
def test_qr_properties(m, n):
    rng = np.random.default_rng(m * 1000 + n)
    x = rng.standard_normal((m, n))
    Q, R = qb.linalg.qr(qb.asarray(x), stabilized=True)
    q, r = _np(Q), _np(R)
    k = min(m, n)
    assert q.shape == (m, k) and r.shape == (k, n)
    np.testing.assert_allclose(q.T @ q, np.eye(k), atol=5e-13)
    np.testing.assert_allclose(q @ r, x, atol=1e-11)
    assert np.all(np.diag(r) >= 0)
    assert np.allclose(np.tril(r[:, :k], -1), 0)
    # against the oracle (unique factorisation)
    lo, _, ro = dn.qr_stabilized(x.copy())
    np.testing.assert_allclose(q, lo, atol=1e-9)
    np.testing.assert_allclose(r, ro, atol=1e-9)


def test_qr_rank_deficient():
    rng = np.random.default_rng(9)
    x = rng.standard_normal((60, 5)) @ rng.standard_normal((5, 20))
    Q, R = qb.linalg.qr(qb.asarray(x), stabilized=True)
    q, r = _np(Q), _np(R)
    np.testing.assert_allclose(q @ r, x, atol=1e-11)
    np.testing.assert_allclose(q.T @ q, np.eye(20), atol=1e-12)


@pytest.mark.parametrize("m,n", [(16, 16), (24, 10), (9, 20), (100, 100), (256, 64),
                                 (130, 130), (512, 512)])
def test_svd_singular_values_and_factors(m, n):
    rng = np.random.default_rng(m * 7 + n)
    x = rng.standard_normal((m, n))
    U, s, VH = qb.linalg.svd(qb.asarray(x))
    u, sv, vh = _np(U), _np(s), _np(VH)
    k = min(m, n)
    assert u.shape == (m, k) and sv.shape == (k,) and vh.shape == (k, n)
    ref = np.linalg.svd(x, compute_uv=False)
    np.testing.assert_allclose(sv, ref, rtol=1e-11, atol=1e-12 * ref[0])
    assert np.all(np.diff(sv) <= 0)
    np.testing.assert_allclose((u * sv) @ vh, x, atol=1e-10)
    np.testing.assert_allclose(u.T @ u, np.eye(k), atol=1e-11)
    np.testing.assert_allclose(vh @ vh.T, np.eye(k), atol=1e-11)


def test_svd_graded_spectrum_relative_accuracy():
    rng = np.random.default_rng(11)
    n = 96
    q1, _ = np.linalg.qr(rng.standard_normal((n, n)))
    q2, _ = np.linalg.qr(rng.standard_normal((n, n)))
    s = 0.7 ** np.arange(n)
    x = (q1 * s) @ q2
    _, sv, _ = qb.linalg.svd(qb.asarray(x))
    np.testing.assert_allclose(_np(sv), s, rtol=1e-9, atol=1e-16)


def test_svd_truncated_matches_reference_golden(golden_decomp):
    data, meta = golden_decomp
    for c in meta["svd_cases"]:
        x = data[f"mat__{c['mat']}"]
        info = {"error": None}
        left, s, right = qb.svd_truncated(
            qb.asarray(x), cutoff=c["cutoff"], cutoff_mode=c["cutoff_mode"],
            max_bond=c["max_bond"], absorb=c["absorb"], renorm=c["renorm"], info=info)
        k = left.shape[1] if left is not None else right.shape[0]
        assert k == c["n_keep"], c                      # integer: exact
        assert info["error"] == pytest.approx(c["error"], rel=1e-8, abs=1e-11)
        if c["key"] + "__s" in data:
            np.testing.assert_allclose(_np(s), data[c["key"] + "__s"], rtol=1e-10, atol=1e-13)
        if c["key"] + "__rec" in data:
            l, r = _np(left), _np(right)
            rec = l @ (np.diag(_np(s)) @ r if s is not None else r)
            np.testing.assert_allclose(rec, data[c["key"] + "__rec"], atol=1e-10)


def test_tensor_split_matches_reference_golden(golden_decomp):
    data, meta = golden_decomp
    x = data["split__x"]
    for c in meta["split_cases"]:
        kw = dict(c["kw"])
        left_inds = kw.pop("left_inds")
        right_inds = kw.pop("right_inds", None)
        out = qb.tensor_split(qb.asarray(x), "abcd", left_inds, right_inds, **kw)
        assert len(out) == c["n_out"]
        refs = [data[f"{c['key']}__{j}"] for j in range(c["n_out"])]
        got = [_np(o) for o in out]
        for g, r in zip(got, refs):
            assert g.shape == r.shape        # bond position / shapes: exact

        def rebuild(parts):
            if len(parts) == 3:
                l, s, r = parts
                return np.tensordot(l * s, r, axes=1)
            return np.tensordot(parts[0], parts[1], axes=1)
        np.testing.assert_allclose(rebuild(got), rebuild(refs), atol=1e-10)


def test_split_bad_options_raise_like_reference():
    x = qb.asarray(np.eye(4))
    with pytest.raises(ValueError):
        qb.array_split(x, method="qr", absorb=None)
    with pytest.raises(ValueError):
        qb.qr_stabilized(x, absorb="both")
    with pytest.raises(KeyError):
        qb.svd_truncated(x, cutoff_mode="nope")


def test_dmrg_size_split_properties():
    """Two-site tensor at a DMRG-like size: truncated SVD keeps the best rank-k
    approximation (Eckart-Young) and the kept factor is an isometry."""
    rng = np.random.default_rng(21)
    chi, d = 128, 2
    x = rng.standard_normal((chi * d, d * chi))
    info = {"error": None}
    left, _, right = qb.svd_truncated(qb.asarray(x), cutoff=0.0, cutoff_mode=3,
                                      max_bond=chi, absorb=1, info=info)
    l, r = _np(left), _np(right)
    assert l.shape == (chi * d, chi) and r.shape == (chi, d * chi)
    np.testing.assert_allclose(l.T @ l, np.eye(chi), atol=1e-11)
    sref = np.linalg.svd(x, compute_uv=False)
    assert info["error"] == pytest.approx(np.sqrt(np.sum(sref[chi:] ** 2)), rel=1e-10)
    assert np.linalg.norm(l @ r - x) == pytest.approx(info["error"], rel=1e-9)


def test_float32_split_preserves_dtype():
    rng = np.random.default_rng(4)
    x = rng.standard_normal((40, 24)).astype(np.float32)
    left, _, right = qb.svd_truncated(qb.asarray(x), cutoff=0.0, max_bond=10, absorb=0)
    assert left.dtype == np.float32 and right.dtype == np.float32
    s = np.linalg.svd(x.astype(np.float64), compute_uv=False)
    err = np.linalg.norm(left.to_numpy().astype(np.float64) @ right.to_numpy().astype(np.float64) - x)
    assert err == pytest.approx(np.sqrt(np.sum(s[10:] ** 2)), rel=1e-4)
    Q, _, R = qb.qr_stabilized(qb.asarray(x))
    assert Q.dtype == np.float32
    np.testing.assert_allclose(Q.to_numpy() @ R.to_numpy(), x, atol=1e-5)


def test_bond_canonize_compress_match_reference(golden_decomp):
    data, meta = golden_decomp
    a, b = data["bond__a"], data["bond__b"]
    na, nb = qb.tensor_canonize_bond(qb.asarray(a), "axb", qb.asarray(b), "cxd")
    np.testing.assert_allclose(_np(na), data["bond__canon_a"], atol=1e-11)   # unique (QR)
    np.testing.assert_allclose(_np(nb), data["bond__canon_b"], atol=1e-11)
    for c in meta["bond_cases"]:
        xa, xb = qb.tensor_compress_bond(qb.asarray(a), "axb", qb.asarray(b), "cxd", **c["kw"])
        ra, rb = data[c["key"] + "_a"], data[c["key"] + "_b"]
        assert xa.shape == ra.shape and xb.shape == rb.shape     # bond position and size: exact
        np.testing.assert_allclose(np.einsum("axb,cxd->abcd", _np(xa), _np(xb)),
                                   np.einsum("axb,cxd->abcd", ra, rb), atol=1e-10)
    with pytest.raises(ValueError):
        qb.tensor_compress_bond(qb.asarray(a), "axb", qb.asarray(b), "cyd")


@pytest.mark.parametrize("m,n", [(12, 14), (40, 24), (64, 64)])
def test_complex_svd_qr_via_real_embedding(m, n):
    rng = np.random.default_rng(m + n)
    x = rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))
    U, s, VH = qb.linalg.svd(qb.asarray(x))
    u, sv, vh = _np(U), _np(s), _np(VH)
    k = min(m, n)
    assert u.shape == (m, k) and vh.shape == (k, n) and sv.dtype == np.float64
    np.testing.assert_allclose(sv, np.linalg.svd(x, compute_uv=False), rtol=1e-11)
    np.testing.assert_allclose((u * sv) @ vh, x, atol=1e-11)
    np.testing.assert_allclose(u.conj().T @ u, np.eye(k), atol=1e-11)
    np.testing.assert_allclose(vh @ vh.conj().T, np.eye(k), atol=1e-11)
    Q, R = qb.linalg.qr(qb.asarray(x), stabilized=True)
    q, r = _np(Q), _np(R)
    np.testing.assert_allclose(q @ r, x, atol=1e-11)
    np.testing.assert_allclose(q.conj().T @ q, np.eye(k), atol=1e-12)
    assert np.abs(np.diag(r).imag).max() < 1e-13 and (np.diag(r).real >= 0).all()
    lo, _, ro = dn.qr_stabilized(x.copy())
    np.testing.assert_allclose(q, lo, atol=1e-10)


def test_complex_svd_degenerate_singular_values():
    rng = np.random.default_rng(5)
    q1, _ = np.linalg.qr(rng.standard_normal((16, 16)) + 1j * rng.standard_normal((16, 16)))
    q2, _ = np.linalg.qr(rng.standard_normal((16, 16)) + 1j * rng.standard_normal((16, 16)))
    s = np.array([3.0] * 4 + [1.0] * 6 + [0.5] * 6)
    x = (q1 * s) @ q2
    U, sv, VH = qb.linalg.svd(qb.asarray(x))
    u, svn, vh = _np(U), _np(sv), _np(VH)
    np.testing.assert_allclose(svn, s, atol=1e-12)
    np.testing.assert_allclose((u * svn) @ vh, x, atol=1e-11)
    np.testing.assert_allclose(u.conj().T @ u, np.eye(16), atol=1e-11)


def test_complex64_split_preserves_dtype():
    rng = np.random.default_rng(6)
    x = (rng.standard_normal((30, 20)) + 1j * rng.standard_normal((30, 20))).astype(np.complex64)
    left, _, right = qb.svd_truncated(qb.asarray(x), cutoff=0.0, max_bond=8, absorb=0)
    assert left.dtype == np.complex64 and right.dtype == np.complex64
    sref = np.linalg.svd(x.astype(np.complex128), compute_uv=False)
    err = np.linalg.norm(left.to_numpy().astype(np.complex128) @ right.to_numpy().astype(np.complex128) - x)
    assert err == pytest.approx(np.sqrt(np.sum(sref[8:] ** 2)), rel=1e-4)
