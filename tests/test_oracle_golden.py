"""CPU tier: the numpy oracle (oracle/) is pinned against golden vectors
produced by the unmodified reference (oracle/make_golden.py) and against the
known answers the reference's own tests hold for this path."""

import numpy as np
import pytest

from oracle import contract_np as cn
from oracle import decomp_np as dn
from oracle import dmrg_np as dm


def test_contract_matches_reference(golden_contract):
    data, meta = golden_contract
    for name, m in meta.items():
        if name.startswith("_"):
            continue
        arrays = [data[f"{name}__in{k}"] for k in range(len(m["inds"]))]
        inds = [tuple(i) for i in m["inds"]]
        out, inds_out = cn.tensor_contract(arrays, inds, m["output_inds"])
        # index bookkeeping: bit exact
        assert list(inds_out) == m["result_inds"], name
        ref = data[f"{name}__out"]
        assert out.shape == ref.shape, name
        np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-12, err_msg=name)


def test_strip_exponent_matches_reference(golden_contract):
    """tensor_contract(strip_exponent=True, exponent=...) (tensor_core.py:330-341):
    result = mantissa * 10**exponent, the mantissa has unit max-magnitude and
    the inputs (scaled up to 1e160) would overflow without stripping."""
    data, meta = golden_contract
    n = 0
    for name, m in meta.items():
        if name.startswith("_") or "strip" not in m:
            continue
        sc = m["strip"]["input_scales"]
        arrays = [data[f"{name}__in{k}"] * 10.0 ** sc[k] for k in range(len(m["inds"]))]
        inds = [tuple(i) for i in m["inds"]]
        (mant, e), inds_out = cn.tensor_contract(arrays, inds, m["output_inds"],
                                                 strip_exponent=True,
                                                 exponent=m["strip"]["base_exponent"])
        assert list(inds_out) == m["result_inds"]
        assert e == pytest.approx(m["strip"]["exponent"], abs=1e-9), name
        np.testing.assert_allclose(mant, data[f"{name}__strip_mantissa"], rtol=1e-12, atol=1e-13)
        assert np.max(np.abs(mant)) == pytest.approx(1.0, abs=1e-14)
        # consistent with the plain contraction of the unscaled inputs
        np.testing.assert_allclose(mant * 10.0 ** (e - sum(sc) - m["strip"]["base_exponent"]),
                                   data[f"{name}__out"], rtol=1e-11, atol=1e-12)
        n += 1
    assert n >= 4


def test_triple_index_error(golden_contract):
    _, meta = golden_contract
    with pytest.raises(ValueError) as e:
        cn.gen_output_inds(list("ab") + list("bc") + list("bd"))
    assert str(e.value) == meta["_triple_index_error"]


def test_output_inds_order_rule():
    # tests/test_tensor/test_tensor_core.py:434-512 (TestTensorContract)
    assert cn.gen_output_inds("abc" "cd" "ea") == ("b", "d", "e")
    assert cn.gen_output_inds([0, 1, 2, 2, 3]) == (0, 1, 3)


def test_svals_to_keep_known_answers(golden_decomp):
    _, meta = golden_decomp
    s = np.array(meta["svals_to_keep"]["s"])
    for cutoff, mode, expect in meta["svals_to_keep"]["cases"]:
        assert dn.number_svals_to_keep(s, cutoff, mode) == expect
    # the reference's own test (test_decomp.py:52-57)
    s = np.array([3.0, 2.0, 1.0, 0.1])
    assert dn.number_svals_to_keep(s, 1.1, 1) == 2


def test_renorm_known_answers():
    # test_tensor_core.py:677-716: sum of squares 385 / sum 55 preserved
    s = np.arange(10, 0, -1.0)
    f2 = dn.svals_renorm_factor(s, 5, 2)
    assert np.isclose(np.sum((s[:5] * f2) ** 2), 385.0)
    f1 = dn.svals_renorm_factor(s, 5, 1)
    assert np.isclose(np.sum(s[:5] * f1), 55.0)


def test_svd_truncated_matches_reference(golden_decomp):
    data, meta = golden_decomp
    for c in meta["svd_cases"]:
        x = data[f"mat__{c['mat']}"]
        info = {}
        left, s, right = dn.svd_truncated(
            x, cutoff=c["cutoff"], cutoff_mode=c["cutoff_mode"],
            max_bond=c["max_bond"], absorb=c["absorb"], renorm=c["renorm"],
            info=info)
        k = left.shape[1] if left is not None else right.shape[0]
        assert k == c["n_keep"], c
        assert info["error"] == pytest.approx(c["error"], rel=1e-9, abs=1e-12)
        if c["key"] + "__s" in data:
            np.testing.assert_allclose(s, data[c["key"] + "__s"], rtol=1e-10)
        if c["key"] + "__rec" in data:
            rec = left @ (np.diag(s) @ right if s is not None else right)
            np.testing.assert_allclose(rec, data[c["key"] + "__rec"], atol=1e-10)


def test_qr_stabilized_matches_reference(golden_decomp):
    data, meta = golden_decomp
    for c in meta["qr_cases"]:
        x = data[f"mat__{c['mat']}"]
        left, _, right = dn.qr_stabilized(x.copy(), absorb=c["absorb"])
        for part, nm in ((left, "__left"), (right, "__right")):
            key = c["key"] + nm
            assert (part is not None) == (key in data), c
            if part is not None:
                np.testing.assert_allclose(part, data[key], atol=1e-11)


def test_tensor_split_matches_reference(golden_decomp):
    data, meta = golden_decomp
    x = data["split__x"]
    for c in meta["split_cases"]:
        kw = dict(c["kw"])
        left_inds = kw.pop("left_inds")
        right_inds = kw.pop("right_inds", None)
        out = dn.tensor_split(x, "abcd", left_inds, right_inds, **kw)
        got = [o for o in out if o is not None]
        assert len(got) == c["n_out"]
        refs = [data[f"{c['key']}__{j}"] for j in range(c["n_out"])]
        for g, r in zip(got, refs):
            assert g.shape == r.shape
        # factors are gauge dependent (signs); their product is not
        def rebuild(parts):
            if len(parts) == 3:
                l, s, r = parts
                return np.tensordot(l * s, r, axes=1)
            return np.tensordot(parts[0], parts[1], axes=1)
        np.testing.assert_allclose(rebuild(got), rebuild(refs), atol=1e-10)


def test_heisenberg_mpo_matches_reference(golden_mps):
    data, _ = golden_mps
    H = dm.mpo_to_dense(dm.mpo_heis(6))
    np.testing.assert_allclose(H, data["heis6__dense"], atol=1e-13)


def _ref_mps_to_lpr(data, prefix, L):
    """reference layout (l, r, p) [ends (r,p)/(l,p)] -> oracle layout (l,p,r)"""
    sites = []
    for i in range(L):
        x = data[f"{prefix}__{i}"]
        if i == 0:
            x = x[None, :, :]            # (1, r, p)
        elif i == L - 1:
            x = x[:, None, :]            # (l, 1, p)
        sites.append(np.transpose(x, (0, 2, 1)))
    return sites


def test_mps_norm_expec_match_reference(golden_mps):
    data, meta = golden_mps
    sites = _ref_mps_to_lpr(data, "mps12", 12)
    n2 = dm.mps_norm2(sites)
    assert n2 == pytest.approx(meta["mps12_norm2"], rel=1e-12)
    e = dm.mps_expec(sites, dm.mpo_heis(12))
    assert e == pytest.approx(meta["mps12_expec_heis"], rel=1e-11)
    csites = _ref_mps_to_lpr(data, "cmps8", 8)
    assert np.real(dm.mps_norm2(csites)) == pytest.approx(meta["cmps8_norm2"], rel=1e-12)
    assert np.real(dm.mps_expec(csites, dm.mpo_heis(8))) == pytest.approx(
        meta["cmps8_expec_heis"], rel=1e-11)


def test_dmrg2_energies_match_reference(golden_mps):
    _, meta = golden_mps
    for run in meta["dmrg2_runs"][:2]:
        d = dm.DMRG2(dm.mpo_heis(run["L"]), run["bond_dims"], cutoffs=run["cutoffs"], seed=7)
        d.solve(tol=run["tol"], max_sweeps=8)
        # converged energies agree with the reference run (different random
        # start, loose inner tolerance -> compare at the solve tolerance)
        assert d.energy == pytest.approx(run["energies"][-1], abs=20 * run["tol"])
        if run["exact"] is not None:
            assert d.energy == pytest.approx(run["exact"], abs=1e-6)


def test_bond_canonize_compress_match_reference(golden_decomp):
    data, meta = golden_decomp
    a, b = data["bond__a"], data["bond__b"]
    na, nb = dn.tensor_canonize_bond(a, "axb", b, "cxd")
    assert meta["bond_canon_inds"] == [list("axb"), list("cxd")]
    np.testing.assert_allclose(na, data["bond__canon_a"], atol=1e-12)
    np.testing.assert_allclose(nb, data["bond__canon_b"], atol=1e-12)
    for c in meta["bond_cases"]:
        xa, xb = dn.tensor_compress_bond(a, "axb", b, "cxd", **c["kw"])
        ra, rb = data[c["key"] + "_a"], data[c["key"] + "_b"]
        assert xa.shape == ra.shape and xb.shape == rb.shape and xa.shape[1] == c["bond"]
        np.testing.assert_allclose(np.einsum("axb,cxd->abcd", xa, xb),
                                   np.einsum("axb,cxd->abcd", ra, rb), atol=1e-11)


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


def test_svd_via_eig_oracle_matches_reference(golden_decomp2):
    """oracle 'svd:eig' (decomp.py:1168-1444) vs the reference's numba path:
    both lose relative accuracy below sqrt(eps) * smax, by the method."""
    data, meta = golden_decomp2
    for c in meta["eig_cases"]:
        x = data[f"mat__{c['mat']}"]
        info = {"error": None} if c["error"] is not None else None
        left, sv, right = dn.svd_via_eig_truncated(
            x, cutoff=c["cutoff"], cutoff_mode=c["cutoff_mode"], max_bond=c["max_bond"],
            absorb=c["absorb"], renorm=c["renorm"], info=info)
        assert [left is not None, sv is not None, right is not None] == c["has"], c
        smax = np.linalg.norm(x, 2)
        if info is not None:
            assert info["n_keep"] == c["n_keep"], c
            assert abs(info["error"] - c["error"]) <= 1e-7 * smax
        for nm, val in _images(left, sv, right).items():
            if c["mat"] == "lowrank" and nm.endswith("gram") and c["absorb"] in (10, -11):
                continue
            scale = smax ** (2 if nm.endswith("gram") else 1)
            np.testing.assert_allclose(val, data[f"{c['key']}__{nm}"], atol=2e-7 * scale,
                                       err_msg=str(c))


def test_eigh_truncated_oracle_matches_reference(golden_decomp2):
    data, meta = golden_decomp2
    for c in meta["eigh_cases"]:
        x = data[f"mat__{c['mat']}"]
        left, sv, right = dn.eigh_truncated(x, **c["kw"])
        assert left.shape[1] == c["n_keep"], c
        for nm, val in _images(left, sv, right).items():
            np.testing.assert_allclose(val, data[f"{c['key']}__{nm}"],
                                       atol=1e-11 * np.linalg.norm(x, 2), err_msg=str(c))


def test_svd_rand_oracle_matches_reference(golden_decomp2):
    """Same numpy Generator and seed as the reference -> same sketch: the
    reconstruction error agrees to rounding."""
    data, meta = golden_decomp2
    for c in meta["rand_cases"]:
        x = data[f"mat__{c['mat']}"]
        left, sv, right = dn.svd_rand_truncated(x, c["max_bond"], absorb=c["absorb"], seed=5)
        assert [left is not None, sv is not None, right is not None] == c["has"], c
        k = left.shape[1] if left is not None else right.shape[0]
        assert k == c["n_keep"], c
        if c["rec_err"] is not None:
            rec = left @ (np.diag(sv) @ right if sv is not None else right)
            err = np.linalg.norm(x - rec)
            assert abs(err - c["rec_err"]) <= 1e-8 * np.linalg.norm(x) + 1e-6 * c["rec_err"], c


def test_svals_drivers_match_reference(golden_decomp2):
    data, _ = golden_decomp2
    for mname in ("tall", "wide", "cplx"):
        x = data[f"mat__{mname}"]
        np.testing.assert_allclose(np.linalg.svd(x, compute_uv=False),
                                   data[f"svals__{mname}__svd"], rtol=1e-12)
        np.testing.assert_allclose(dn.svd_via_eig(x, absorb=dn.get_s)[1],
                                   data[f"svals__{mname}__eig"], rtol=1e-9)


def _golden_tebd():
    import json
    import os
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return np.load(os.path.join(root, "tebd.npz")), json.load(open(os.path.join(root, "tebd.json")))


def _lpr(a, i, n):
    """quimb 'lrp' site array -> (l, p, r)."""
    if a.ndim == 2:
        a = a[None] if i == 0 else a[:, None]
    return np.transpose(a, (0, 2, 1))


def test_gate_split_and_swaps_oracle_match_reference():
    from oracle import tebd_np as tn
    data, meta = _golden_tebd()
    raw = [data[f"gs_mps__{i}"] for i in range(6)]
    G = data["gs_gate"]
    for c in meta["gate_split"]:
        kw = dict(c["kw"])
        where = tuple(kw.pop("where"))
        s = [_lpr(a, i, 6) for i, a in enumerate(raw)]
        tn.canonicalize(s, where)
        tn.gate_split(s, G, where, **kw)
        assert s[min(where)].shape[2] == c["bond"]
        np.testing.assert_allclose(dm.mps_to_dense(s).reshape(-1), data[c["key"] + "__dense"],
                                   atol=1e-12)
    for c in meta["auto_swap"]:
        s = [_lpr(a, i, 6) for i, a in enumerate(raw)]
        tn.gate_with_auto_swap(s, G, tuple(c["where"]), cutoff=1e-12)
        np.testing.assert_allclose(dm.mps_to_dense(s).reshape(-1), data[c["key"] + "__dense"],
                                   atol=1e-11)


def test_tebd_oracle_matches_reference():
    from oracle import tebd_np as tn
    data, meta = _golden_tebd()
    for o, sched in meta["trotter"].items():
        assert [[k, f] for k, f in tn.trotter_schedule(2, int(o))] == sched
    for r in meta["tebd"]:
        L = r["L"]
        terms = {tuple(map(int, k.split(","))): data[f"{r['key']}__term__{k}"]
                 for k in r["terms"]}
        p0 = [np.zeros((1, 2, 1)) for _ in range(L)]
        for i in range(L):
            p0[i][0, i % 2, 0] = 1.0
        kw = dict(dt=r["dt"]) if r["dt"] is not None else dict(tol=r["tol"])
        t = tn.TEBD(p0, terms, imag=r["imag"], split_opts=dict(cutoff=1e-12), **kw)
        t.update_to(r["T"], order=r["order"])
        assert abs(t.err - r["err"]) <= 1e-12 * max(1.0, r["err"])
        assert max(a.shape[2] for a in t.sites) == r["max_bond"]
        np.testing.assert_allclose(dm.mps_to_dense(t.sites).reshape(-1),
                                   data[r["key"] + "__dense"], atol=1e-8)


def test_dmrg1_oracle_matches_reference_energies(golden_mps):
    """One-site DMRG: the reference pads bonds with unseeded noise, so runs are
    compared at the convergence tolerance (and against exact diagonalisation)."""
    _, meta = golden_mps
    for r in meta["dmrg1_runs"]:
        d = dm.DMRG1(dm.mpo_heis(r["L"]), r["bond_dims"], cutoffs=1e-10, seed=1)
        assert d.solve(tol=r["tol"], max_sweeps=12)
        assert abs(d.energy - r["energies"][-1]) < 50 * r["tol"]
        assert abs(d.energy - r["exact"]) < 1e-6
        assert abs(r["energies"][-1] - r["exact"]) < 1e-6     # the reference itself


def test_cholesky_qr_cholesky_polar_oracle_match_reference(golden_decomp3):
    """oracle restatements of 'cholesky', 'qr:cholesky' and the polar splits
    pinned on the reference's outputs (tests/golden/decomp3.*); these factors
    are unique, so they are compared directly."""
    import warnings
    data, meta = golden_decomp3

    def check(res, c, atol):
        left, sv, right = res
        assert sv is None and [left is not None, False, right is not None] == c["has"], c
        if left is not None:
            np.testing.assert_allclose(left, data[f"{c['key']}__left"], atol=atol)
        if right is not None:
            np.testing.assert_allclose(right, data[f"{c['key']}__right"], atol=atol)
    for c in meta["cholesky_cases"]:
        check(dn.cholesky_regularized(data[f"mat__{c['mat']}"], absorb=c["absorb"], shift=c["shift"]),
              c, 1e-12)
    for c in meta["qr_cholesky_cases"]:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            check(dn.qr_via_cholesky(data[f"mat__{c['mat']}"], absorb=c["absorb"]), c, 1e-9)
    for c in meta["polar_cases"]:
        fn = dn.polar_right if c["side"] == "right" else dn.polar_left
        check(fn(data[f"mat__{c['mat']}"]), c, 1e-12)
    with pytest.raises(np.linalg.LinAlgError):
        dn.cholesky_regularized(data["mat__indef"], shift="auto")
