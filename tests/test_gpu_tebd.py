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
