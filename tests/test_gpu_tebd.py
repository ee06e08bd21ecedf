"""GPU tier: MPS gate application (gate_split, swaps) and TEBD against the
states the unmodified reference produced (tests/golden/tebd.*) and against the
numpy oracle on other inputs.  States are compared as dense vectors (gauge
independent); bond dimensions exactly."""

import numpy as np
import pytest

import quimb_b200 as qb
from quimb_b200 import mps, tebd as tb
from oracle import dmrg_np as dm
from oracle import tebd_np as tn

pytestmark = pytest.mark.gpu


def _dense(sites):
    return dm.mps_to_dense([np.asarray(s.to_numpy()) for s in sites]).reshape(-1)


def _fresh(raw):
    n = len(raw)
    return [qb.materialize(mps.site_lpr(a, "lrp", i, n), force=True)
            for i, a in enumerate(raw)]


def test_gate_split_matches_reference(golden_tebd):
    data, meta = golden_tebd
    raw = [data[f"gs_mps__{i}"] for i in range(6)]
    G = data["gs_gate"]
    for c in meta["gate_split"]:
        kw = dict(c["kw"])
        where = tuple(kw.pop("where"))
        s = _fresh(raw)
        tb.canonicalize(s, where)
        tb.gate_split(s, G, where, **kw)
        assert s[min(where)].shape[2] == c["bond"], c
        np.testing.assert_allclose(_dense(s), data[c["key"] + "__dense"], atol=1e-11)
    with pytest.raises(ValueError):
        tb.gate_split(_fresh(raw), G, (0, 2))


def test_gate_with_auto_swap_matches_reference(golden_tebd):
    data, meta = golden_tebd
    raw = [data[f"gs_mps__{i}"] for i in range(6)]
    for c in meta["auto_swap"]:
        s = _fresh(raw)
        tb.gate_with_auto_swap(s, data["gs_gate"], tuple(c["where"]), cutoff=1e-12)
        np.testing.assert_allclose(_dense(s), data[c["key"] + "__dense"], atol=1e-10)


def test_canonize_sites_keep_state_and_make_isometries():
    sites = [qb.asarray(a) for a in dm.mps_rand(7, 6, seed=3)]
    ref = _dense(sites)
    tb.canonicalize(sites, 3)
    np.testing.assert_allclose(_dense(sites), ref, atol=1e-12)
    for i in range(3):
        a = sites[i].to_numpy()
        m = a.reshape(-1, a.shape[2])
        np.testing.assert_allclose(m.T @ m, np.eye(m.shape[1]), atol=1e-12)
    for i in range(4, 7):
        a = sites[i].to_numpy()
        m = a.reshape(a.shape[0], -1)
        np.testing.assert_allclose(m @ m.T, np.eye(m.shape[0]), atol=1e-12)


def test_tebd_matches_reference(golden_tebd):
    data, meta = golden_tebd
    for o, sched in meta["trotter"].items():
        assert [[k, f] for k, f in tb.trotter_schedule(2, int(o))] == sched
    for r in meta["tebd"]:
        L = r["L"]
        terms = {tuple(map(int, k.split(","))): data[f"{r['key']}__term__{k}"]
                 for k in r["terms"]}
        H = tb.LocalHam1D(L, H2=terms)
        p0 = [np.zeros((1, 2, 1)) for _ in range(L)]
        for i in range(L):
            p0[i][0, i % 2, 0] = 1.0
        kw = dict(dt=r["dt"]) if r["dt"] is not None else dict(tol=r["tol"])
        t = tb.TEBD(p0, H, imag=r["imag"], split_opts=dict(cutoff=1e-12), **kw)
        t.update_to(r["T"], order=r["order"])
        assert abs(t.t - r["t"]) < 1e-12 and abs(t.err - r["err"]) <= 1e-12 * max(1, r["err"])
        assert max(a.shape[2] for a in t.pt) == r["max_bond"], r
        np.testing.assert_allclose(_dense(t.pt), data[r["key"] + "__dense"], atol=1e-8)


def test_local_ham_single_site_terms_and_tebd_vs_oracle(golden_tebd):
    data, _ = golden_tebd
    h2 = data["heis_h2"]
    sz = np.diag([0.5, -0.5])
    L = 6
    H = tb.LocalHam1D(L, H2=h2, H1={None: 0.3 * sz, 2: -0.7 * sz})
    # the terms sum to the full Hamiltonian
    full = np.zeros((2 ** L, 2 ** L))
    for (a, b), h in H.terms.items():
        full += np.kron(np.kron(np.eye(2 ** a), h), np.eye(2 ** (L - b - 1)))
    ref = np.zeros_like(full)
    for i in range(L - 1):
        ref += np.kron(np.kron(np.eye(2 ** i), h2), np.eye(2 ** (L - i - 2)))
    for i in range(L):
        c = -0.7 if i == 2 else 0.3
        ref += np.kron(np.kron(np.eye(2 ** i), c * sz), np.eye(2 ** (L - i - 1)))
    np.testing.assert_allclose(full, ref, atol=1e-13)
    p0 = dm.mps_rand(L, 3, seed=5)
    t = tb.TEBD(p0, H, dt=0.05, split_opts=dict(cutoff=1e-12))
    o = tn.TEBD(p0, H.terms, dt=0.05, split_opts=dict(cutoff=1e-12))
    t.update_to(0.2, order=2)
    o.update_to(0.2, order=2)
    np.testing.assert_allclose(_dense(t.pt), dm.mps_to_dense(o.sites).reshape(-1), atol=1e-8)
    # exact evolution of the dense state
    import scipy.linalg as sla
    psi0 = dm.mps_to_dense(p0).reshape(-1)
    exact = sla.expm(-0.2j * ref) @ psi0
    v = _dense(t.pt)
    assert abs(np.vdot(exact, v)) / (np.linalg.norm(exact) * np.linalg.norm(v)) > 1 - 1e-4
