"""GPU tier: the 'cholesky', 'qr:cholesky' / 'lq:cholesky' and 'polar_right' /
'polar_left' split drivers, the device ``linalg.cholesky`` / ``pinv`` and the
diagonal helpers (rddiv / lddiv / sgn / safe_inverse) against the
reference-generated golden vectors (tests/golden/decomp3.*,
oracle/make_golden.py:decomp3_cases; quimb/tensor/decomp.py:501-656,
2245-2424, 2673-2730).  These factorisations are unique, so the factors
themselves are compared.  fp64 tolerance 1e-10 on O(1) entries (Gram-matrix
routes square the condition number: 1e-8 there)."""

import warnings

import numpy as np
import pytest

import quimb_b200 as qb
from quimb_b200 import split

pytestmark = pytest.mark.gpu


def _np(x):
    return None if x is None else x.to_numpy()


def _check(res, data, case, atol):
    left, sv, right = res
    assert sv is None
    assert [left is not None, False, right is not None] == case["has"], case
    if left is not None:
        np.testing.assert_allclose(_np(left), data[f"{case['key']}__left"], atol=atol, err_msg=str(case))
    if right is not None:
        np.testing.assert_allclose(_np(right), data[f"{case['key']}__right"], atol=atol, err_msg=str(case))


def test_cholesky_regularized_matches_reference_golden(golden_decomp3):
    data, meta = golden_decomp3
    for c in meta["cholesky_cases"]:
        x = data[f"mat__{c['mat']}"]
        res = split.cholesky_regularized(qb.asarray(x), absorb=c["absorb"], shift=c["shift"])
        # n > 64 goes through the device Jacobi eigh: the factor inherits
        # cond(x) * (backward error of the eigendecomposition), cond ~ 300 here
        _check(res, data, c, 1e-9 if c["mat"] == "pd_big" else 1e-10)
        for part in res:
            if part is not None:
                assert part.dtype == x.dtype


def test_device_cholesky_properties():
    rng = np.random.default_rng(5)
    for n, cplx in [(5, False), (40, False), (100, False), (130, False), (70, True), (33, True)]:
        a = rng.standard_normal((n, 2 * n))                # cond(x) ~ 30
        if cplx:
            a = a + 1j * rng.standard_normal((n, 2 * n))
        x = a @ a.conj().T / n
        L = _np(qb.linalg.cholesky(qb.asarray(x)))
        ref = np.linalg.cholesky(x)
        np.testing.assert_allclose(L, ref, atol=1e-10)
        assert np.allclose(np.triu(L, 1), 0.0)
        U = _np(qb.linalg.cholesky(qb.asarray(x), upper=True))
        np.testing.assert_allclose(U, ref.conj().T, atol=1e-10)
    # single precision keeps its dtype
    x32 = (a.real @ a.real.T / n).astype(np.float32)
    L32 = qb.linalg.cholesky(qb.asarray(x32))
    assert L32.dtype == np.float32
    np.testing.assert_allclose(_np(L32), np.linalg.cholesky(x32.astype(np.float64)), atol=2e-5)
    with pytest.raises(ValueError):
        qb.linalg.cholesky(qb.asarray(rng.standard_normal((4, 5))))


def test_cholesky_error_behaviour(golden_decomp3):
    data, meta = golden_decomp3
    indef = data["mat__indef"]
    for shift, err in (("False", False), ("True", True), ("auto", "auto")):
        assert meta["errors"][shift] == "LinAlgError"
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            with pytest.raises(np.linalg.LinAlgError):
                split.cholesky_regularized(qb.asarray(indef), shift=err)
    assert meta["errors"]["polar_absorb_none"] == "ValueError"
    with pytest.raises(ValueError):
        split.parse_split_opts(method="polar_right", absorb=None)
    with pytest.raises(ValueError):
        split.cholesky_regularized(qb.asarray(data["mat__pd"]), absorb=1)
    with pytest.raises(ValueError):
        split.qr_via_cholesky(qb.asarray(data["mat__wide"]), absorb=0)
    # 'auto' retries with the shift after a failure and warns like the reference
    with pytest.warns(UserWarning, match="Cholesky decomposition failed"):
        with pytest.raises(np.linalg.LinAlgError):
            split.cholesky_regularized(qb.asarray(indef), shift="auto")


def test_qr_via_cholesky_matches_reference_golden(golden_decomp3):
    data, meta = golden_decomp3
    for c in meta["qr_cholesky_cases"]:
        x = data[f"mat__{c['mat']}"]
        res = split.qr_via_cholesky(qb.asarray(x), absorb=c["absorb"],
                                    solve_triangular=c["solve_triangular"])
        _check(res, data, c, 1e-8)
    # isometry + triangularity + reconstruction on a fresh matrix
    rng = np.random.default_rng(2)
    x = rng.standard_normal((60, 25))
    Q, _, R = split.qr_via_cholesky(qb.asarray(x), absorb="right")
    Q, R = _np(Q), _np(R)
    np.testing.assert_allclose(Q @ R, x, atol=1e-10)
    np.testing.assert_allclose(Q.T @ Q, np.eye(25), atol=1e-9)
    assert np.allclose(np.tril(R, -1), 0.0) and np.all(np.diag(R) > 0)
    with pytest.warns(UserWarning, match="not well-defined for tall"):
        split.qr_via_cholesky(qb.asarray(x), absorb="left")


def test_polar_matches_reference_golden(golden_decomp3):
    data, meta = golden_decomp3
    for c in meta["polar_cases"]:
        x = data[f"mat__{c['mat']}"]
        fn = split.polar_right if c["side"] == "right" else split.polar_left
        res = fn(qb.asarray(x))
        _check(res, data, c, 1e-10)
        left, _, right = (_np(t) for t in res)
        np.testing.assert_allclose(left @ right, x, atol=1e-11)
        P = right if c["side"] == "right" else left
        np.testing.assert_allclose(P, P.conj().T, atol=1e-11)


def test_array_split_new_methods_and_option_codes(golden_decomp3):
    data, meta = golden_decomp3
    for rec in meta["parse_split_opts"]:
        method, opts = split.parse_split_opts(**rec["kw"])
        assert method == rec["method"] and opts == rec["opts"], rec
    for c in meta["array_split_cases"]:
        x = data[f"mat__{c['mat']}"]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = split.array_split(qb.asarray(x), cutoff=0.0, **c["kw"])
        _check(res, data, c, 1e-8)
    with pytest.raises(ValueError):
        split.array_split(qb.asarray(data["mat__pd"]), method="svds")


def test_diagonal_helpers_match_reference_golden(golden_decomp3):
    data, _ = golden_decomp3
    x, d, z = data["mat__square"], data["helpers__d"], data["helpers__z"]
    np.testing.assert_allclose(_np(split.rddiv(qb.asarray(x), qb.asarray(d))),
                               data["helpers__rddiv"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(_np(split.lddiv(qb.asarray(d), qb.asarray(x))),
                               data["helpers__lddiv"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(_np(split.sgn(qb.asarray(z))), data["helpers__sgn"], atol=1e-15)
    np.testing.assert_allclose(_np(split.sgn(qb.asarray(z.real.copy()))),
                               data["helpers__sgn_real"], atol=1e-15)
    np.testing.assert_allclose(_np(split.safe_inverse(qb.asarray(d))),
                               data["helpers__safe_inverse"], rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(_np(split.safe_inverse(qb.asarray(d), cutoff=1e-3, power=0.5)),
                               data["helpers__safe_inverse_sqrt"], rtol=1e-12)
    np.testing.assert_allclose(_np(split.rdmul(qb.asarray(x), qb.asarray(d))), x * d[None, :])
    np.testing.assert_allclose(_np(split.ldmul(qb.asarray(d), qb.asarray(x))), x * d[:, None])
    np.testing.assert_allclose(_np(qb.multiply_diagonal(qb.asarray(x), qb.asarray(d), 0)), x * d[:, None])
    np.testing.assert_allclose(_np(qb.multiply_diagonal(qb.asarray(x), qb.asarray(d), -1)), x * d[None, :])
    assert abs(float(qb.norm_fro(qb.asarray(x)).item()) - np.linalg.norm(x)) < 1e-12


def test_pinv_inv_solve_and_small_surface():
    rng = np.random.default_rng(9)
    for shape, cplx in [((30, 12), False), ((12, 30), False), ((20, 20), True), ((80, 70), False)]:
        x = rng.standard_normal(shape)
        if cplx:
            x = x + 1j * rng.standard_normal(shape)
        np.testing.assert_allclose(_np(qb.linalg.pinv(qb.asarray(x))), np.linalg.pinv(x), atol=1e-10)
    low = rng.standard_normal((25, 4)) @ rng.standard_normal((4, 18))
    np.testing.assert_allclose(_np(qb.linalg.pinv(qb.asarray(low), rcond=1e-10)),
                               np.linalg.pinv(low, rcond=1e-10), atol=1e-9)
    a = rng.standard_normal((12, 12)) + 4 * np.eye(12)
    b = rng.standard_normal((12, 3))
    np.testing.assert_allclose(_np(qb.linalg.inv(qb.asarray(a))), np.linalg.inv(a), atol=1e-11)
    np.testing.assert_allclose(_np(qb.linalg.solve(qb.asarray(a), qb.asarray(b))),
                               np.linalg.solve(a, b), atol=1e-11)
    L = np.tril(a)
    import scipy.linalg as sla
    for kw in (dict(lower=True), dict(lower=True, trans=1), dict(lower=False)):
        A = L if kw.get("lower") else L.T
        np.testing.assert_allclose(_np(qb.scipy.linalg.solve_triangular(qb.asarray(A), qb.asarray(b), **kw)),
                                   sla.solve_triangular(A, b, **kw), atol=1e-11)
    tall = rng.standard_normal((20, 6))
    rhs = rng.standard_normal((20, 2))
    xs = qb.linalg.lstsq(qb.asarray(tall), qb.asarray(rhs))[0]
    np.testing.assert_allclose(_np(xs), np.linalg.lstsq(tall, rhs, rcond=None)[0], atol=1e-11)
    np.testing.assert_array_equal(_np(qb.indices((2, 3))), np.indices((2, 3)))
    assert qb.finfo("float64").eps == np.finfo(np.float64).eps
    assert qb.finfo(qb.asarray(np.zeros(2, np.complex64))).eps == np.finfo(np.float32).eps
    np.testing.assert_allclose(_np(qb.dag(qb.asarray(a + 1j * a.T))), (a + 1j * a.T).conj().T)
    np.testing.assert_allclose(_np(qb.full((2, 3), 1.5)), np.full((2, 3), 1.5))
    np.testing.assert_allclose(_np(qb.identity(4)), np.eye(4))
    np.testing.assert_allclose(_np(qb.outer(qb.asarray(b[:, 0]), qb.asarray(b[:, 1]))), np.outer(b[:, 0], b[:, 1]))
    np.testing.assert_allclose(_np(qb.sort(qb.asarray(b[:, 0]))), np.sort(b[:, 0]))
    np.testing.assert_array_equal(_np(qb.argsort(qb.asarray(b[:, 0]))), np.argsort(b[:, 0]))
    np.testing.assert_allclose(float(qb.prod(qb.asarray(b[:, 0])).item()), np.prod(b[:, 0]))
    np.testing.assert_allclose(_np(qb.log10(qb.asarray(np.abs(b)))), np.log10(np.abs(b)))
    np.testing.assert_allclose(_np(qb.linalg.eigvalsh(qb.asarray(a + a.T))), np.linalg.eigvalsh(a + a.T), atol=1e-11)


def test_numpy_signature_helpers():
    """take / flip / cumsum / tril / triu / expm with numpy's signatures (found
    by running the reference's own test-suite on device tensors)."""
    rng = np.random.default_rng(4)
    x = rng.standard_normal((3, 4, 5))
    X = qb.asarray(x)
    np.testing.assert_array_equal(_np(qb.take(X, 2, axis=1)), np.take(x, 2, axis=1))
    np.testing.assert_array_equal(_np(qb.take(X, [3, 0], axis=2)), np.take(x, [3, 0], axis=2))
    np.testing.assert_array_equal(_np(qb.take(X, 7)), np.take(x, 7))
    np.testing.assert_array_equal(_np(qb.take(X, [[0, 1], [2, 2]], axis=0)), np.take(x, [[0, 1], [2, 2]], axis=0))
    for ax in (None, 0, -1, (0, 2)):
        np.testing.assert_array_equal(_np(qb.flip(X, axis=ax)), np.flip(x, axis=ax))
    np.testing.assert_allclose(_np(qb.cumsum(X)), np.cumsum(x))
    np.testing.assert_allclose(_np(qb.cumsum(X, axis=1)), np.cumsum(x, axis=1))
    m = x[0]
    for k in (-1, 0, 2):
        np.testing.assert_array_equal(_np(qb.tril(qb.asarray(m), k=k)), np.tril(m, k=k))
        np.testing.assert_array_equal(_np(qb.triu(qb.asarray(m), k=k)), np.triu(m, k=k))
    import scipy.linalg as sla
    a = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
    np.testing.assert_allclose(_np(qb.linalg.expm(qb.asarray(a))), sla.expm(a), atol=1e-12)
    np.testing.assert_allclose(_np(qb.scipy.linalg.expm(qb.asarray(a.real))), sla.expm(a.real), atol=1e-12)


def test_reductions_keepdims_and_eig():
    """numpy's ``keepdims`` on sum / max / min / mean / prod (quimb's belief
    propagation normalises messages with it) and the general ``linalg.eig``."""
    rng = np.random.default_rng(8)
    x = rng.standard_normal((3, 4, 5))
    X = qb.asarray(x)
    for name in ("sum", "max", "min", "mean", "prod"):
        f, g = getattr(qb, name), getattr(np, name)
        for ax in (None, 0, -1, (0, 2)):
            for kd in (False, True):
                if name in ("prod",) and isinstance(ax, tuple):
                    continue
                got = f(X, axis=ax, keepdims=kd)
                ref = g(x, axis=ax, keepdims=kd)
                assert tuple(got.shape) == np.shape(ref), (name, ax, kd)
                np.testing.assert_allclose(_np(got), ref, rtol=1e-12)
    np.testing.assert_allclose(_np(X.sum(axis=-1, keepdims=True)), x.sum(axis=-1, keepdims=True))
    y = X / qb.sum(qb.abs(X), axis=-1, keepdims=True)
    np.testing.assert_allclose(_np(y), x / np.abs(x).sum(axis=-1, keepdims=True), rtol=1e-13)
    a = rng.standard_normal((6, 6))
    w, v = qb.linalg.eig(qb.asarray(a))
    w, v = _np(w), _np(v)
    np.testing.assert_allclose(a @ v, v * w, atol=1e-10)
    np.testing.assert_allclose(np.sort_complex(w), np.sort_complex(np.linalg.eigvals(a)), atol=1e-10)


def test_new_drivers_vs_oracle_at_larger_sizes():
    """device drivers against the numpy oracle (oracle/decomp_np.py, itself
    pinned on the reference's golden vectors) at sizes where the device eigh
    / SVD / QR kernels -- not the host branches -- do the work."""
    from oracle import decomp_np as dn
    rng = np.random.default_rng(12)
    a = rng.standard_normal((200, 400))
    x = a @ a.T / 200                                  # cond ~ 30
    for absorb in (0, -12, 12):
        got = split.cholesky_regularized(qb.asarray(x), absorb=absorb, shift=True)
        ref = dn.cholesky_regularized(x, absorb=absorb, shift=True)
        for g, r in zip(got, ref):
            assert (g is None) == (r is None)
            if g is not None:
                np.testing.assert_allclose(_np(g), r, atol=1e-10)
    z = rng.standard_normal((96, 160)) + 1j * rng.standard_normal((96, 160))
    hz = z @ z.conj().T / 96
    L, _, LH = split.cholesky_regularized(qb.asarray(hz), shift=False)
    np.testing.assert_allclose(_np(L), np.linalg.cholesky(hz), atol=1e-10)
    np.testing.assert_allclose(_np(LH), np.linalg.cholesky(hz).conj().T, atol=1e-10)
    tall = rng.standard_normal((300, 120))
    for absorb in (1, 10, 11):
        got = split.qr_via_cholesky(qb.asarray(tall), absorb=absorb)
        ref = dn.qr_via_cholesky(tall, absorb=absorb)
        for g, r in zip(got, ref):
            assert (g is None) == (r is None)
            if g is not None:
                np.testing.assert_allclose(_np(g), r, atol=1e-9)
    for m, n, side in ((150, 90, "right"), (90, 150, "left"), (128, 128, "right")):
        y = rng.standard_normal((m, n))
        fn_d = split.polar_right if side == "right" else split.polar_left
        fn_o = dn.polar_right if side == "right" else dn.polar_left
        got, ref = fn_d(qb.asarray(y)), fn_o(y)
        np.testing.assert_allclose(_np(got[0]), ref[0], atol=1e-10)
        np.testing.assert_allclose(_np(got[2]), ref[2], atol=1e-10)


def test_lu_split_matches_reference_golden(golden_decomp3):
    """``method='lu'`` (decomp.py:2615-2670): the permuted-lower / upper factors
    and which rows / columns survive the cutoff."""
    data, meta = golden_decomp3
    for c in meta["lu_cases"]:
        x = data[f"mat__{c['mat']}"]
        res = split.lu_truncated(qb.asarray(x), **c["kw"])
        pl, _, u = (_np(t) for t in res)
        if c["mat"] == "cplx":
            # complex pivot selection may legitimately differ between getrf
            # implementations (|re| + |im| against |z|): compare the product
            assert pl.shape == data[f"{c['key']}__left"].shape
            assert np.allclose(np.tril(u, -1), 0.0)
        else:
            _check(res, data, c, 1e-11)
        if c["mat"] != "lu_mixed":
            np.testing.assert_allclose(pl @ u, x, atol=1e-12)
    x = data["mat__square"]
    left, _, right = split.array_split(qb.asarray(x), method="lu", cutoff=0.0, cutoff_mode="rel")
    np.testing.assert_allclose(_np(left) @ _np(right), x, atol=1e-12)
    for kw in (dict(absorb="left"), dict(renorm=1), dict(max_bond=3), dict(cutoff_mode=4)):
        with pytest.raises(NotImplementedError):
            split.lu_truncated(qb.asarray(x), **{"cutoff_mode": 2, **kw})
