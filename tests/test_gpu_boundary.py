"""GPU tier: boundary-MPS contraction of 2D networks (BASELINE config 5, PEPS
norm with a two-layer boundary) against the values the unmodified reference
produced for the same tensors and options (tests/golden/boundary.*).  The
truncated result is an approximation of the exact contraction, but with the
same sequence of gauge moves it is the SAME approximation: parity is asserted
at rounding level, far below the truncation error itself."""

import numpy as np
import pytest

import quimb_b200 as qb
from quimb_b200 import boundary as bd

pytestmark = pytest.mark.gpu


def _net(data, m, name):
    return [(data[f"{name}__t{k}"], r["inds"], tuple(r["site"]), r["layer"])
            for k, r in enumerate(m["tensors"])]


@pytest.mark.parametrize("name", ["peps44", "peps35", "peps53_c64", "flat55", "flat64"])
def test_contract_boundary_matches_reference(golden_boundary, name):
    data, meta = golden_boundary
    m = meta[name]
    ts = _net(data, m, name)
    single = "64" in m["dtype"] and "complex64" in m["dtype"]
    tol = 2e-5 if single else 1e-10
    for run in m["runs"]:
        v = bd.contract_boundary(ts, m["Lx"], m["Ly"], layer_tags=m["layers"], **run["kw"])
        ref = complex(*run["value"])
        assert abs(v - ref) <= tol * abs(ref), (name, run["kw"], v, ref)


def test_boundary_converges_to_exact(golden_boundary):
    data, meta = golden_boundary
    m = meta["peps44"]
    ts = _net(data, m, "peps44")
    exact = complex(*m["exact"])
    errs = []
    for chi in (4, 16, 81):
        v = bd.contract_boundary(ts, 4, 4, max_bond=chi, cutoff=0.0, layer_tags=m["layers"])
        errs.append(abs(v - exact) / abs(exact))
    assert errs[0] > errs[1] > errs[2]
    assert errs[2] < 1e-10                      # chi = D^4: no truncation at all


def test_peps_norm_tensors_convention(golden_boundary):
    """quimb's PEPS site arrays (index order up, right, down, left, phys) fed
    through peps_norm_tensors give the same value as the reference's own
    norm network."""
    data, meta = golden_boundary
    sites = meta["peps44_site_inds"]
    arrays = [[data[f"peps44_site__{i}_{j}"] for j in range(4)] for i in range(4)]
    bonds = meta["peps44_bonds"]
    # check the assumed order against the reference's index names
    for i in range(4):
        for j in range(4):
            want = []
            if i < 3:
                want.append(bonds[f"{i},{j}"]["up"])
            if j < 3:
                want.append(bonds[f"{i},{j}"]["right"])
            if i > 0:
                want.append(bonds[f"{i - 1},{j}"]["up"])
            if j > 0:
                want.append(bonds[f"{i},{j - 1}"]["right"])
            assert sites[f"{i},{j}"][:-1] == want
    ts, Lx, Ly = bd.peps_norm_tensors(arrays)
    m = meta["peps44"]
    for run in m["runs"][:3]:
        v = bd.contract_boundary(ts, Lx, Ly, layer_tags=("KET", "BRA"), **run["kw"])
        ref = complex(*run["value"])
        assert abs(v - ref) <= 1e-10 * abs(ref)


def test_compress_bond_reduced_modes_agree():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((5, 12, 4))
    b = rng.standard_normal((12, 6, 3))
    full = np.einsum("axb,xcd->abcd", a, b)
    for reduced in (True, False, "left", "right"):
        na, nb = qb.tensor_compress_bond(qb.asarray(a), "axb", qb.asarray(b), "xcd",
                                         max_bond=12, cutoff=0.0, reduced=reduced)
        rec = np.einsum("axb,xcd->abcd", na.to_numpy(), nb.to_numpy())
        if reduced in (True, False):
            np.testing.assert_allclose(rec, full, atol=1e-10)
        assert na.shape[0] == 5 and nb.shape[1:] == (6, 3)
    # truncating: reduced=True and reduced=False give the optimal truncation
    outs = []
    for reduced in (True, False):
        na, nb = qb.tensor_compress_bond(qb.asarray(a), "axb", qb.asarray(b), "xcd",
                                         max_bond=5, cutoff=0.0, reduced=reduced)
        outs.append(np.einsum("axb,xcd->abcd", na.to_numpy(), nb.to_numpy()))
    np.testing.assert_allclose(outs[0], outs[1], atol=1e-10)
    with pytest.raises(ValueError):
        qb.tensor_compress_bond(qb.asarray(a), "axb", qb.asarray(b), "xcd", reduced="up")
