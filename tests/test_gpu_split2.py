"""GPU tier: the Gram-matrix SVD ('svd:eig'), Hermitian 'eigh' split,
randomized SVD ('svd:rand') and the device ``linalg.eigh`` against the
reference-generated golden vectors (tests/golden/decomp2.*) and the oracle.
Parity on gauge-invariant images: values, kept rank, truncation error,
reconstruction / Gram matrices of single factors."""

import numpy as np
import pytest

import quimb_b200 as qb
from quimb_b200 import split
from oracle import decomp_np as dn

pytestmark = pytest.mark.gpu


def _np(x):
    return None if x is None else x.to_numpy()


def _images(left, sv, right):
    out = {}
    if sv is not None:
        out["s"] = sv
    if left is not None and right is not None:
        out["rec"] = left @ (np.diag(sv) @ right if sv is not None else right)
    elif left is not None:
        out["lgram"] = left @ left.conj().T
    elif right is not None:
        out["rgram"] = right.conj().T @ right
    return out


@pytest.mark.parametrize("n,dtype", [(96, "float64"), (130, "float64"), (200, "float64"),
                                     (80, "complex128"), (72, "float32")])
def test_device_eigh_vs_lapack(n, dtype):
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n))
    if dtype == "complex128":
        a = a + 1j * rng.standard_normal((n, n))
    a = (a + a.conj().T).astype(dtype)
    w, v = qb.linalg.eigh(qb.asarray(a))
    w, v = _np(w), _np(v)
    assert w.dtype == np.dtype(dtype).type(0).real.dtype and v.dtype == np.dtype(dtype)
    ref = np.linalg.eigvalsh(a.astype("complex128" if dtype == "complex128" else "float64"))
    scale = np.abs(ref).max()
    tol = 2e-5 if dtype == "float32" else 1e-12
    assert np.all(np.diff(w) >= -tol * scale)
    np.testing.assert_allclose(w, ref, atol=tol * scale * n ** 0.5)
    np.testing.assert_allclose(v.conj().T @ v, np.eye(n), atol=50 * tol)
    np.testing.assert_allclose((v * w) @ v.conj().T, a, atol=50 * tol * scale)


def test_device_eigh_degenerate_and_tiny():
    rng = np.random.default_rng(3)
    q, _ = np.linalg.qr(rng.standard_normal((100, 100)))
    lam = np.concatenate([np.full(30, -2.0), np.full(30, 2.0), np.linspace(-1, 1, 40)])
    a = (q * lam) @ q.T
    w, v = (_np(t) for t in qb.linalg.eigh(qb.asarray(a)))
    np.testing.assert_allclose(w, np.sort(lam), atol=1e-12)
    np.testing.assert_allclose((v * w) @ v.T, a, atol=1e-11)
    # host-side control-logic branch (tiny projected problems)
    b = rng.standard_normal((6, 6)); b = b + b.T
    w, v = (_np(t) for t in qb.linalg.eigh(qb.asarray(b)))
    np.testing.assert_allclose(w, np.linalg.eigvalsh(b), atol=1e-13)
    with pytest.raises(ValueError):
        qb.linalg.eigh(qb.asarray(rng.standard_normal((4, 5))))


def test_svd_via_eig_matches_reference_golden(golden_decomp2):
    data, meta = golden_decomp2
    for c in meta["eig_cases"]:
        x = data[f"mat__{c['mat']}"]
        info = {"error": None} if c["error"] is not None else None
        left, sv, right = split.svd_via_eig_truncated(
            qb.asarray(x), cutoff=c["cutoff"], cutoff_mode=c["cutoff_mode"],
            max_bond=c["max_bond"], absorb=c["absorb"], renorm=c["renorm"], info=info)
        left, sv, right = _np(left), _np(sv), _np(right)
        assert [left is not None, sv is not None, right is not None] == c["has"], c
        k = (left.shape[1] if left is not None else
             right.shape[0] if right is not None else sv.shape[0])
        assert k == c["n_keep"], c
        smax = np.linalg.norm(x, 2)
        if info is not None:
            # Gram-matrix values carry sqrt(eps) * smax absolute noise
            assert abs(info["error"] - c["error"]) <= 1e-7 * smax, c
        for nm, val in _images(left, sv, right).items():
            ref = data[f"{c['key']}__{nm}"]
            scale = smax ** (2 if nm.endswith("gram") else 1)
            if c["mat"] == "lowrank" and nm.endswith("gram") and c["absorb"] in (10, -11):
                continue  # isometries of the numerical null space are not unique
            np.testing.assert_allclose(val, ref, atol=2e-7 * scale, err_msg=str(c))


def test_svd_via_eig_isometries_and_oracle():
    rng = np.random.default_rng(5)
    for m, n in [(300, 40), (40, 300), (128, 128), (513, 70)]:
        x = rng.standard_normal((m, n))
        U, s, VH = (_np(t) for t in split.svd_via_eig(qb.asarray(x)))
        k = min(m, n)
        np.testing.assert_allclose(s, np.linalg.svd(x, compute_uv=False), rtol=1e-9)
        np.testing.assert_allclose((U * s) @ VH, x, atol=1e-10)
        np.testing.assert_allclose(U.T @ U, np.eye(k), atol=1e-8)
        np.testing.assert_allclose(VH @ VH.T, np.eye(k), atol=1e-8)
        lo, so, ro = dn.svd_via_eig(x)
        np.testing.assert_allclose(s, so, rtol=1e-9)
    xs = np.asarray(_np(split.array_svals(qb.asarray(x), method="svd:eig")))
    np.testing.assert_allclose(xs, np.linalg.svd(x, compute_uv=False), rtol=1e-9)


def test_eigh_truncated_matches_reference_golden(golden_decomp2):
    data, meta = golden_decomp2
    for c in meta["eigh_cases"]:
        x = data[f"mat__{c['mat']}"]
        left, sv, right = (_np(t) for t in split.eigh_truncated(qb.asarray(x), **c["kw"]))
        assert left.shape[1] == c["n_keep"], c
        scale = np.linalg.norm(x, 2)
        for nm, val in _images(left, sv, right).items():
            ref = data[f"{c['key']}__{nm}"]
            np.testing.assert_allclose(val, ref, atol=1e-11 * scale, err_msg=str(c))


def test_svd_rand_matches_reference_accuracy(golden_decomp2):
    data, meta = golden_decomp2
    for c in meta["rand_cases"]:
        x = data[f"mat__{c['mat']}"]
        left, sv, right = (_np(t) for t in split.svd_rand_truncated(
            qb.asarray(x), max_bond=c["max_bond"], absorb=c["absorb"], seed=5))
        assert [left is not None, sv is not None, right is not None] == c["has"], c
        k = left.shape[1] if left is not None else right.shape[0]
        assert k == c["n_keep"], c
        if c["rec_err"] is not None:
            rec = left @ (np.diag(sv) @ right if sv is not None else right)
            err = np.linalg.norm(x - rec)
            # a different Gaussian sketch: same accuracy class as the reference
            assert err <= 1.5 * c["rec_err"] + 1e-10 * np.linalg.norm(x), (c, err)
            assert err >= c["optimal_err"] * (1 - 1e-9) - 1e-9
        if left is not None and c["absorb"] in (1, 10):
            np.testing.assert_allclose(left.conj().T @ left, np.eye(k), atol=1e-10)
        if right is not None and c["absorb"] in (-1, -11):
            np.testing.assert_allclose(right @ right.conj().T, np.eye(k), atol=1e-10)


def test_svd_rand_large_lowrank_and_seed_reproducible():
    rng = np.random.default_rng(8)
    x = rng.standard_normal((600, 24)) @ rng.standard_normal((24, 500))
    xa = qb.asarray(x)
    l1, _, r1 = split.array_split(xa, method="svd:rand", max_bond=24, absorb="right", seed=1)
    l2, _, r2 = split.array_split(xa, method="svd:rand", max_bond=24, absorb="right", seed=1)
    np.testing.assert_allclose(_np(l1) @ _np(r1), x, atol=1e-9 * np.linalg.norm(x, 2))
    np.testing.assert_array_equal(_np(l1), _np(l2))
    with pytest.warns(UserWarning):
        split.svd_rand_truncated(qb.asarray(x[:40, :30]), max_bond=None)


def test_parse_split_opts_new_methods_match_reference(golden_decomp2):
    _, meta = golden_decomp2
    for c in meta["parse_split_opts"]:
        method, opts = split.parse_split_opts(**c["kw"])
        assert method == c["method"], c
        assert opts == c["opts"], c
    with pytest.warns(FutureWarning):
        assert split.parse_split_opts(method="eig")[0] == "svd:eig"
    with pytest.raises(ValueError):
        split.parse_split_opts(method="svds")


def test_array_split_dispatch_and_tensor_split_methods():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((6, 5, 4, 7))
    for method in ("svd", "svd:eig", "svd:rand"):
        kw = dict(seed=0) if method == "svd:rand" else {}
        if method == "svd:rand":
            l, r = (_np(t) for t in qb.tensor_split(
                qb.asarray(x), "abcd", "ac", method=method, max_bond=24, absorb="left", **kw))
        else:
            l, r = (_np(t) for t in qb.tensor_split(
                qb.asarray(x), "abcd", "ac", method=method, cutoff=0.0, absorb="left"))
        rec = np.einsum("ack,kbd->abcd", l, r)
        np.testing.assert_allclose(rec, x, atol=1e-9)


# ---- the reference's own property tests for these drivers, restated --------
# (tests/test_tensor/test_decomp.py:302-375, 378-437, 464-486, 748-790)
@pytest.fixture(params=["host_eigh", "device_eigh"])
def eigh_mode(request, monkeypatch):
    """Small Gram matrices take the host branch of linalg.eigh (n <= 64, control
    logic); 'device_eigh' forces the Jacobi path so it sees the same cases."""
    if request.param == "device_eigh":
        monkeypatch.setattr(qb.linalg, "_EIGH_HOST_BELOW", 0)
    return request.param


@pytest.mark.parametrize("da,db", [(5, 5), (5, 7), (7, 5)])
@pytest.mark.parametrize("k", [-1, 6, 8])
@pytest.mark.parametrize("descending", [True, False])
def test_svd_via_eig_properties(da, db, k, descending, eigh_mode):
    rng = np.random.default_rng(da * 100 + db * 10 + k + 2 + int(descending))
    x = rng.uniform(size=(da, db))
    x /= np.linalg.norm(x)
    Ux, sx, VHx = np.linalg.svd(x, full_matrices=False)
    kk = min(da, db, k) if k > 0 else min(da, db)
    sx, Ux, VHx = sx[:kk], Ux[:, :kk], VHx[:kk]
    if not descending:
        sx, Ux, VHx = sx[::-1], Ux[:, ::-1], VHx[::-1]
    for absorb in ["U", "s", "VH", "Us", "sVH", "U,s,VH", "U,sVH", "Us,VH"]:
        U, s, VH = (_np(t) for t in split.svd_via_eig(
            qb.asarray(x), max_bond=k, absorb=absorb, descending=descending))
        if absorb in ("U", "U,s,VH", "U,sVH"):
            assert U.shape == (da, kk)
            np.testing.assert_allclose(U.T @ U, np.eye(kk), atol=1e-9)
            np.testing.assert_allclose(np.abs(U.T @ Ux), np.eye(kk), atol=1e-8)
        if absorb in ("s", "U,s,VH"):
            assert s.shape == (kk,)
            np.testing.assert_allclose(s, sx, atol=1e-9)
            assert np.all(np.diff(s) <= 0) if descending else np.all(np.diff(s) >= 0)
        if absorb in ("VH", "Us,VH", "U,s,VH"):
            assert VH.shape == (kk, db)
            np.testing.assert_allclose(VH @ VH.T, np.eye(kk), atol=1e-9)
            np.testing.assert_allclose(np.abs(VHx @ VH.T), np.eye(kk), atol=1e-8)
        if absorb in ("Us", "Us,VH"):
            np.testing.assert_allclose(U.T @ U, np.diag(sx ** 2), atol=1e-9)
        if absorb in ("sVH", "U,sVH"):
            np.testing.assert_allclose(VH @ VH.T, np.diag(sx ** 2), atol=1e-9)
        if absorb in ("Us,VH", "U,sVH", "U,s,VH"):
            rec = (U * s) @ VH if absorb == "U,s,VH" else U @ VH
            if k > min(da, db) or k < 0:
                np.testing.assert_allclose(rec, x, atol=1e-9)
            else:
                assert np.linalg.norm(x - rec) < 0.2


@pytest.mark.parametrize("da,db", [(4, 8), (8, 4), (6, 6)])
@pytest.mark.parametrize("right", [True, False, None])
def test_svd_rand_right_param(right, da, db):
    rng = np.random.default_rng(da * 10 + db + (3 if right is None else int(right)))
    x = rng.uniform(size=(da, db))
    x /= np.linalg.norm(x)
    U, s, VH = (_np(t) for t in split.svd_rand_truncated(
        qb.asarray(x), absorb=None, max_bond=3, right=right, seed=17))
    assert U.shape == (da, 3) and s.shape == (3,) and VH.shape == (3, db)
    assert np.all(s >= 0)
    np.testing.assert_allclose(U.T @ U, np.eye(3), atol=1e-9)
    np.testing.assert_allclose(VH @ VH.T, np.eye(3), atol=1e-9)
    assert np.linalg.norm(x - (U * s) @ VH) < 0.5


@pytest.mark.parametrize("absorb", ["U", "s", "VH", "Us", "sVH", "U,s,VH", "U,sVH", "Us,VH"])
@pytest.mark.parametrize("k", [-1, 4])
def test_svd_rand_absorb_modes(absorb, k):
    rng = np.random.default_rng(5 + k)
    x = rng.uniform(size=(7, 5))
    x /= np.linalg.norm(x)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        U, s, VH = (_np(t) for t in split.svd_rand_truncated(
            qb.asarray(x), max_bond=k, absorb=absorb, seed=3))
    kk = 5 if k < 0 else k
    sx = np.linalg.svd(x, compute_uv=False)[:kk]
    want = {"U": (1, 0, 0), "s": (0, 1, 0), "VH": (0, 0, 1), "Us": (1, 0, 0), "sVH": (0, 0, 1),
            "U,s,VH": (1, 1, 1), "U,sVH": (1, 0, 1), "Us,VH": (1, 0, 1)}[absorb]
    assert tuple(int(t is not None) for t in (U, s, VH)) == want
    if absorb in ("s", "U,s,VH"):
        np.testing.assert_allclose(s, sx, atol=1e-6 if k > 0 else 1e-9)
    if absorb in ("U", "U,sVH", "U,s,VH"):
        np.testing.assert_allclose(U.T @ U, np.eye(kk), atol=1e-9)
    if absorb in ("VH", "Us,VH", "U,s,VH"):
        np.testing.assert_allclose(VH @ VH.T, np.eye(kk), atol=1e-9)
    if absorb in ("Us,VH", "U,sVH"):
        assert np.linalg.norm(x - U @ VH) < (1e-9 if k < 0 else 0.3)


def test_eigh_shift_semantics(eigh_mode):
    """array_split(method='eigh', shift=...) (test_decomp.py:748-790): False /
    default add nothing, a float adds shift * trace, True adds eps * trace."""
    x = np.diag([4.0, 1.0, 0.0])
    tr = 5.0

    def vals(**kw):
        return _np(qb.array_split(qb.asarray(x), method="eigh", absorb="s", cutoff=0.0,
                                  positive=1, **kw)[1])
    np.testing.assert_allclose(vals(), [4.0, 1.0, 0.0], atol=1e-12)
    np.testing.assert_allclose(vals(shift=False), [4.0, 1.0, 0.0], atol=1e-12)
    np.testing.assert_allclose(vals(shift=0.1), np.array([4.0, 1.0, 0.0]) + 0.1 * tr,
                               atol=1e-12)
    np.testing.assert_allclose(vals(shift=True),
                               np.array([4.0, 1.0, 0.0]) + np.finfo(float).eps * tr, atol=1e-12)


@pytest.mark.parametrize("method", ["svd", "svd:eig", "svd:rand", "qr", "lq"])
@pytest.mark.parametrize("absorb", ["left", "right"])
@pytest.mark.parametrize("m,n", [(8, 5), (5, 5), (5, 8)])
def test_isometric_factor_across_methods(method, absorb, m, n, eigh_mode):
    """Every method returns x = L @ R with the non-absorbing factor isometric
    (test_decomp.py:517-553)."""
    if eigh_mode == "device_eigh" and method != "svd:eig":
        pytest.skip("eigh mode only matters for svd:eig")
    rng = np.random.default_rng(m * 10 + n)
    x = rng.standard_normal((m, n))
    kw = dict(seed=1, max_bond=min(m, n)) if method == "svd:rand" else {}
    if method in ("svd", "svd:eig"):
        kw["cutoff"] = 0.0
    L, _, R = (_np(t) for t in qb.array_split(qb.asarray(x), method=method, absorb=absorb, **kw))
    np.testing.assert_allclose(L @ R, x, atol=1e-9)
    k = min(m, n)
    if absorb == "right":
        np.testing.assert_allclose(L.T @ L, np.eye(k), atol=1e-8)
    else:
        np.testing.assert_allclose(R @ R.T, np.eye(k), atol=1e-8)
