"""GPU tier: whole-tree contraction of circuit-amplitude networks (many small
complex128 tensors, greedy tree) against an independent dense state-vector
simulation, and the slice-parallel decomposition used for multi-GPU runs."""

import numpy as np
import pytest

import quimb_b200 as qb
from oracle import contract_np as cn
from tests.circuit_util import random_circuit_amplitude, random_grid_circuit_amplitude

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nq,depth,seed", [(6, 4, 0), (10, 6, 1), (12, 8, 2)])
def test_circuit_amplitude_matches_statevector(nq, depth, seed):
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2, nq).tolist()
    arrays, inputs, output, amp = random_circuit_amplitude(nq, depth, seed, bits)
    out = qb.array_contract(arrays, inputs, output, optimize="greedy")
    val = complex(out.item())
    assert abs(val - amp) <= 1e-10 * max(1.0, abs(amp))      # BASELINE: atol 1e-10
    # same tree through the numpy oracle
    ref = cn.array_contract(arrays, inputs, output, "greedy")
    assert abs(val - complex(ref)) <= 1e-12


def test_sliced_contraction_sums_to_the_same_amplitude():
    arrays, inputs, output, amp = random_circuit_amplitude(8, 6, 5)
    # slice two internal indices: 4 independent contractions, summed
    counts = {}
    for t in inputs:
        for ix in t:
            counts[ix] = counts.get(ix, 0) + 1
    sliced = [ix for ix, c in counts.items() if c == 2][10:12]
    dev = [qb.asarray(a) for a in arrays]
    total, mine = qb.dist.contract_sliced(dev, inputs, output, sliced, optimize="greedy",
                                          rank=0, world_size=1)
    assert len(mine) == 4
    assert abs(complex(total.item()) - amp) <= 1e-10


def test_cuda_graph_replay_of_a_tree():
    arrays, inputs, output, amp = random_circuit_amplitude(8, 6, 3)
    g = qb.GraphedContraction(inputs, output, arrays, optimize="greedy")
    v1 = complex(g(*arrays).item())
    assert abs(v1 - amp) <= 1e-10
    # new input values (different output bitstring), same structure: one replay
    arrays2, _, _, amp2 = random_circuit_amplitude(8, 6, 3, bits=[1, 0, 1, 1, 0, 0, 1, 0])
    n0 = qb.launch_count()
    v2 = complex(g(*arrays2).item())
    assert qb.launch_count() == n0          # no host-side launches: graph replay
    assert abs(v2 - amp2) <= 1e-10


@pytest.mark.parametrize("Lx,Ly,depth,gate", [(3, 3, 8, "fsim"), (2, 4, 12, "cz"),
                                              (4, 4, 8, "fsim")])
def test_grid_circuit_amplitude_matches_statevector(Lx, Ly, depth, gate):
    """BASELINE configs[3] geometry at CPU-checkable size: qubits on a grid,
    random U3 layers, fSim / CZ on the bond patterns A, B, C, D in turn."""
    from tests.circuit_util import grid_bond_patterns, random_grid_circuit_amplitude
    pats = grid_bond_patterns(Lx, Ly)
    allb = [b for p in pats for b in p]
    assert len(allb) == len(set(allb)) == Lx * (Ly - 1) + Ly * (Lx - 1)
    rng = np.random.default_rng(Lx * 10 + Ly)
    for bits in ([0] * (Lx * Ly), rng.integers(0, 2, Lx * Ly).tolist()):
        arrays, inputs, output, amp = random_grid_circuit_amplitude(
            Lx, Ly, depth, seed=3, bits=bits, gate=gate)
        out = qb.array_contract(arrays, inputs, output, optimize="greedy")
        assert abs(complex(out.item()) - amp) <= 1e-10 * max(1.0, abs(amp))


def test_contract_sliced_with_tree_and_slices_found_together():
    """dist.contract_sliced(optimize='auto-hq', target_width=...): tree and
    sliced indices come from find_sliced_tree (slicing interleaved with
    annealing / reconfiguration), every rank executes its share of the
    slices with the tree of the sliced network; the parts sum to the
    state-vector amplitude."""
    from quimb_b200 import dist
    arrays, inputs, output, amp = random_grid_circuit_amplitude(3, 3, 10, seed=7)
    dev = [qb.asarray(a) for a in arrays]
    total, seen = 0.0, []
    for r in range(2):
        part, mine = dist.contract_sliced(dev, inputs, output, optimize="auto-hq", target_width=6,
                                          rank=r, world_size=2, reduce=False)
        seen += list(mine)
        total += complex(part.item())
    assert sorted(seen) == list(range(len(seen))) and len(seen) >= 2
    assert abs(total - amp) < 1e-12
