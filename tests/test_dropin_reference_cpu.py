"""CPU tier, build container only: the UNMODIFIED reference (quimb from
/root/reference, its third-party autoray / cotengra / cytoolz layer served by
oracle/shims) driven with ``quimb_b200.Array`` objects as ``Tensor._data`` --
the drop-in boundary of SURVEY 8(b) exercised from the reference's side:
autoray dispatch on the array type, quimb's composed-driver registration and
its partial-eigensolver backend table (INTEGRATION.md sections 1-4).

Every numeric call quimb makes lands on the product's host layer; the kernel
launching ABI calls underneath are served by tests/abi_emulator.py (no device
here), so this checks names, signatures, the Array protocol and option
plumbing against the reference's own numpy run -- not the CUDA kernels, which
the ``-m gpu`` tier covers.  Skipped where /root/reference does not exist
(the GPU box)."""

import os
import sys
import warnings

import numpy as np
import pytest

REF = os.environ.get("QUIMB_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
SHIMS = os.path.join(os.path.dirname(HERE), "oracle", "shims")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "quimb")),
                                reason="reference source tree not present")


@pytest.fixture(scope="module")
def env():
    for p in (REF, SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import quimb.tensor as qtn
    import quimb_b200 as qb
    from tests.abi_emulator import emulated_abi
    with emulated_abi():
        names = qb.register_with_quimb()
        yield qtn, qb, names


def _dev(tn, qb):
    tn = tn.copy()
    tn.apply_to_arrays(qb.asarray)
    return tn


def test_registration_covers_the_composed_drivers(env):
    qtn, qb, names = env
    for nm in ("svd_truncated", "qr_stabilized", "svd_via_eig_truncated", "eigh_truncated",
               "cholesky_regularized", "polar_right", "polar_left", "fuse", "unfuse",
               "norm_fro", "eigs:QUIMB_B200"):
        assert nm in names


def test_tensor_contract_and_matmul_stay_on_the_backend(env):
    qtn, qb, _ = env
    rng = np.random.default_rng(0)
    a = qtn.Tensor(rng.standard_normal((4, 5, 6)), inds="abc", tags="A")
    b = qtn.Tensor(rng.standard_normal((6, 5, 7)), inds="cbd", tags="B")
    c = qtn.Tensor(rng.standard_normal((7, 3)), inds="de", tags="C")
    ref2 = (a @ b).data
    ref3 = qtn.tensor_contract(a, b, c, output_inds="ea").data
    ad, bd, cd = (_dev(t, qb) for t in (a, b, c))
    out = ad @ bd
    assert isinstance(out.data, qb.Array) and out.inds == ("a", "d") and out.tags == a.tags | b.tags
    np.testing.assert_allclose(out.data.to_numpy(), ref2, atol=1e-13)
    out3 = qtn.tensor_contract(ad, bd, cd, output_inds="ea")
    assert isinstance(out3.data, qb.Array)
    np.testing.assert_allclose(out3.data.to_numpy(), ref3, atol=1e-13)
    # hyper index + explicit outputs goes through do("einsum")
    h1 = qtn.Tensor(rng.standard_normal((3, 4)), inds="xh")
    h2 = qtn.Tensor(rng.standard_normal((4, 5)), inds="hy")
    h3 = qtn.Tensor(rng.standard_normal((4, 2)), inds="hz")
    refh = qtn.tensor_contract(h1, h2, h3, output_inds="xyz").data
    outh = qtn.tensor_contract(*(_dev(t, qb) for t in (h1, h2, h3)), output_inds="xyz")
    np.testing.assert_allclose(outh.data.to_numpy(), refh, atol=1e-13)
    # full contraction to a scalar, complex, with the mantissa / exponent split
    z1 = qtn.Tensor(rng.standard_normal((3, 4)) + 1j * rng.standard_normal((3, 4)), inds="ab")
    z2 = qtn.Tensor(rng.standard_normal((4, 3)) + 1j * rng.standard_normal((4, 3)), inds="ba")
    refz = qtn.tensor_contract(z1, z2)
    outz = qtn.tensor_contract(_dev(z1, qb), _dev(z2, qb))
    assert abs(complex(outz) - complex(refz)) < 1e-13
    m, e = qtn.tensor_contract(_dev(z1, qb), _dev(z2, qb), strip_exponent=True)
    assert abs(complex(m) * 10 ** e - complex(refz)) < 1e-12


@pytest.mark.parametrize("method,kw", [
    ("svd", dict(cutoff=1e-3, cutoff_mode="rel")), ("svd", dict(max_bond=3, absorb="left")),
    ("svd:eig", dict(max_bond=4)), ("qr", {}), ("lq", {}), ("eigh", dict(max_bond=4)),
    ("polar_right", {}), ("polar_left", {}),
])
def test_tensor_split_methods_through_quimb(env, method, kw):
    qtn, qb, _ = env
    rng = np.random.default_rng(1)
    x = rng.standard_normal((6, 4, 5))
    if method == "eigh":
        y = rng.standard_normal((6, 4, 6, 4))
        x = y + y.transpose(2, 3, 0, 1)
        t = qtn.Tensor(x, inds="abcd")
        left = ["a", "b"]
    else:
        t = qtn.Tensor(x, inds="abc")
        left = ["a", "b"] if method in ("qr", "polar_right") else ["a"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = t.split(left_inds=left, method=method, get="arrays", **kw)
        out = _dev(t, qb).split(left_inds=left, method=method, get="arrays", **kw)
    assert all(isinstance(o, qb.Array) for o in out)
    assert [o.shape for o in out] == [r.shape for r in ref]
    # gauge-free comparison: the product of the factors
    prod_ref = np.tensordot(ref[0], ref[-1], 1)
    prod_out = np.tensordot(out[0].to_numpy(), out[-1].to_numpy(), 1)
    np.testing.assert_allclose(prod_out, prod_ref, atol=1e-10)


def test_canonize_compress_and_linear_operator(env):
    qtn, qb, _ = env
    p = qtn.MPS_rand_state(6, 7, seed=4)
    pd = _dev(p, qb)
    p.left_canonize()
    pd.left_canonize()
    for i in range(6):
        assert isinstance(pd[i].data, qb.Array)
    assert abs(float(pd.H @ pd) - float(p.H @ p)) < 1e-12
    p.compress(max_bond=3)
    pd.compress(max_bond=3)
    assert pd.max_bond() == 3
    assert abs(float(pd.H @ pd) - float(p.H @ p)) < 1e-10
    H = qtn.MPO_ham_heis(6)
    e_ref = qtn.expec_TN_1D(p.H, H, p)
    e_dev = qtn.expec_TN_1D(pd.H, _dev(H, qb), pd)
    assert abs(float(e_dev) - float(e_ref)) < 1e-10
    # TNLinearOperator: device matvec fed from / read back by scipy (host vectors)
    rng = np.random.default_rng(2)
    ts = [qtn.Tensor(rng.standard_normal((5, 3, 5)), inds=("a", "w", "b"), tags="L"),
          qtn.Tensor(rng.standard_normal((3, 2, 2)), inds=("w", "p", "q"), tags="W")]
    from quimb.tensor.tensor_core import TNLinearOperator
    A = TNLinearOperator(ts, left_inds=("a", "p"), right_inds=("b", "q"))
    Ad = TNLinearOperator([_dev(t, qb) for t in ts], left_inds=("a", "p"), right_inds=("b", "q"))
    v = rng.standard_normal(10)
    np.testing.assert_allclose(np.asarray(Ad.matvec(v)), A.matvec(v), atol=1e-12)
    np.testing.assert_allclose(np.asarray(Ad.to_dense()), A.to_dense(), atol=1e-12)
    out = Ad._matvec(qb.asarray(v))            # device in, device out
    assert isinstance(out, qb.Array)


def test_reference_dmrg2_default_eigensolver_path(env):
    """No backend selected: quimb hands the dense effective Hamiltonian or the
    TNLinearOperator to scipy ARPACK on the host (dmrg.py:626-645); the
    products run on the device, the Krylov vectors cross the boundary."""
    qtn, qb, _ = env
    L = 8
    H = qtn.MPO_ham_heis(L)
    p0 = qtn.MPS_rand_state(L, 4, seed=3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = qtn.DMRG2(H.copy(), bond_dims=[8, 16, 32], cutoffs=1e-10, p0=p0.copy())
        ref.solve(tol=1e-8, max_sweeps=5, verbosity=0)
        for dense in (None, False):
            dm = qtn.DMRG2(_dev(H, qb), bond_dims=[8, 16, 32], cutoffs=1e-10, p0=_dev(p0, qb))
            dm.opts["local_eig_ham_dense"] = dense
            dm.solve(tol=1e-8, max_sweeps=5, verbosity=0)
            assert abs(float(dm.energy) - float(ref.energy)) < 1e-6


@pytest.mark.parametrize("dense", [True, False])
def test_reference_dmrg2_runs_on_device_arrays(env, dense):
    qtn, qb, _ = env
    L = 8
    H = qtn.MPO_ham_heis(L)
    p0 = qtn.MPS_rand_state(L, 8, seed=3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = qtn.DMRG2(H.copy(), bond_dims=[8, 16], cutoffs=1e-10, p0=p0.copy())
        ref.solve(tol=1e-9, max_sweeps=5, verbosity=0)
        dm = qtn.DMRG2(_dev(H, qb), bond_dims=[8, 16], cutoffs=1e-10, p0=_dev(p0, qb))
        dm.opts["local_eig_backend"] = "quimb_b200"
        dm.opts["local_eig_ham_dense"] = dense
        dm.solve(tol=1e-9, max_sweeps=5, verbosity=0)
    assert all(isinstance(dm.state[i].data, qb.Array) for i in range(L))
    assert abs(float(dm.energy) - float(ref.energy)) < 1e-6
    # exact ground state of the open 8-site Heisenberg chain
    import quimb as qu
    exact = qu.groundenergy(qu.ham_heis(L, cyclic=False, sparse=True))
    assert abs(float(dm.energy) - exact) < 1e-6
    assert [dm.state[i].shape for i in range(L)] == [ref.state[i].shape for i in range(L)]


def test_reference_callers_either_side_of_the_path(env):
    """The reference's own drivers around the hot path (boundary
    contraction, circuits, TEBD, DMRG1, MPS gates / arithmetic, rank
    simplification) on device arrays, against their numpy runs."""
    import quimb as qu
    qtn, qb, _ = env
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # PEPS norm by boundary-MPS contraction (tn2d/core.py:2528-2543)
        peps = qtn.PEPS.rand(4, 4, bond_dim=2, seed=1, dtype="complex128")
        norm = peps.make_norm()
        ref = norm.contract_boundary(max_bond=8, cutoff=0.0, layer_tags=("KET", "BRA"))
        out = _dev(norm, qb).contract_boundary(max_bond=8, cutoff=0.0, layer_tags=("KET", "BRA"))
        assert abs(complex(out) - complex(ref)) < 1e-10 * abs(ref)
        ising = qtn.TN2D_classical_ising_partition_function(4, 4, beta=0.3)
        assert abs(float(_dev(ising, qb).contract_boundary(max_bond=8))
                   - float(ising.contract_boundary(max_bond=8))) < 1e-6

        # circuit amplitude with to_backend (circuit/exact.py:90-98)
        def build(**kw):
            rng = np.random.default_rng(0)
            circ = qtn.Circuit(5, **kw)
            for d in range(4):
                for q in range(5):
                    circ.apply_gate("U3", *rng.uniform(0, 6, 3), q)
                for q in range(d % 2, 4, 2):
                    circ.apply_gate("CZ", q, q + 1)
            return circ
        assert abs(complex(build(to_backend=qb.asarray).amplitude("01001"))
                   - complex(build().amplitude("01001"))) < 1e-12
        tn = build().amplitude_tn("00000")
        assert abs(complex(_dev(tn, qb).full_simplify() ^ all) - complex(tn.full_simplify() ^ all)) < 1e-12

        # TEBD (tn1d/tebd.py)
        ham = qtn.ham_1d_heis(6)
        psi0 = qtn.MPS_neel_state(6)
        t0 = qtn.TEBD(psi0.copy(), ham, progbar=False)
        t0.update_to(0.2, dt=0.05, order=2)
        t1 = qtn.TEBD(_dev(psi0, qb), ham, progbar=False)
        t1.update_to(0.2, dt=0.05, order=2)
        assert isinstance(t1.pt[2].data, qb.Array)
        host = t1.pt.copy()
        host.apply_to_arrays(lambda x: x.to_numpy())
        assert abs(abs(complex(t0.pt.H @ host)) - abs(complex(t0.pt.H @ t0.pt))) < 1e-10

        # DMRG1 with the device eigensolver backend
        H = qtn.MPO_ham_heis(8)
        p0 = qtn.MPS_rand_state(8, 8, seed=3)
        r = qtn.DMRG1(H.copy(), bond_dims=[8, 16], p0=p0.copy())
        r.solve(tol=1e-8, max_sweeps=4, verbosity=0)
        d = qtn.DMRG1(_dev(H, qb), bond_dims=[8, 16], p0=_dev(p0, qb))
        d.opts["local_eig_backend"] = "quimb_b200"
        d.solve(tol=1e-8, max_sweeps=4, verbosity=0)
        assert abs(float(d.energy) - float(r.energy)) < 1e-7

        # MPS gates, MPO application, addition, entropy, dense vector
        p = qtn.MPS_rand_state(6, 4, seed=2)
        G = qu.rand_uni(4, seed=1).reshape(2, 2, 2, 2)
        for where, fn in (((2, 3), "gate_split"), ((0, 4), "gate_with_auto_swap")):
            rr = getattr(p, fn)(G, where, cutoff=1e-12)
            dd = getattr(_dev(p, qb), fn)(qb.asarray(G), where, cutoff=1e-12)
            np.testing.assert_allclose(np.asarray(dd.to_dense()), rr.to_dense(), atol=1e-10)
        Hh = qtn.MPO_ham_heis(6)
        np.testing.assert_allclose(np.asarray(_dev(Hh, qb).apply(_dev(p, qb)).to_dense()),
                                   Hh.apply(p).to_dense(), atol=1e-10)
        q = qtn.MPS_rand_state(6, 3, seed=6)
        np.testing.assert_allclose(np.asarray((_dev(p, qb) + _dev(q, qb)).to_dense()),
                                   (p + q).to_dense(), atol=1e-12)
        assert abs(float(_dev(p, qb).entropy(3)) - float(p.entropy(3))) < 1e-10


def test_reference_compressed_contraction_runs_on_the_backend_and_matches_the_mirror(env):
    """The reference's own ``_contract_compressed_tid_sequence`` (tensor_core.py:
    8560-8780; 'basic' mode, and the default tree-gauged 'virtual-tree' mode the
    array-level mirror does not duplicate) driven with device arrays: same
    values as its numpy run, and -- for the mirrored mode -- as
    ``quimb_b200.contract_compressed`` on the same sequence."""
    qtn, qb, _ = env
    tn = qtn.TN2D_rand(4, 4, D=3, seed=5)
    tids = list(tn.tensor_map)
    seq_pos = []
    # a simple inward sequence: absorb the tensors row by row into the last one
    order = list(range(len(tids)))
    for i in range(len(order) - 1):
        seq_pos.append((order[i], order[i + 1]))
    seq = [(tids[a], tids[b]) for a, b in seq_pos]
    for kw in (dict(max_bond=4, cutoff=0.0, tree_gauge_distance=0, compress_mode="basic"),
               dict(max_bond=6, cutoff=1e-8, tree_gauge_distance=0, compress_mode="basic",
                    compress_late=False),
               dict(max_bond=5, cutoff=0.0)):                 # default: virtual-tree, gauge 1
        ref = tn.copy()._contract_compressed_tid_sequence(seq, output_inds=(), **kw)
        dev = _dev(tn, qb)._contract_compressed_tid_sequence(seq, output_inds=(), **kw)
        dev = dev.item() if hasattr(dev, "item") else complex(dev)
        assert abs(dev - ref) <= 1e-9 * abs(ref), kw
        if kw.get("compress_mode") == "basic":
            arrays = [qb.asarray(np.asarray(tn.tensor_map[t].data)) for t in tids]
            inputs = [tuple(map(str, tn.tensor_map[t].inds)) for t in tids]
            kw2 = {k: v for k, v in kw.items() if k not in ("tree_gauge_distance", "compress_mode")}
            mir = qb.contract_compressed(arrays, inputs, (), seq_pos, **kw2)
            assert abs(mir.item() - ref) <= 1e-9 * abs(ref), kw
