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
        split.parse_split_opts(method="cholesky")


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
