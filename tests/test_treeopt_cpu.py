"""Host-only tests of the tree refinement layer (quimb_b200/treeopt.py): the
stand-in for cotengra's subtree / slicing reconfiguration on the path
``array_contract_tree`` -> ``tensor_contract`` (quimb/tensor/contraction.py:
302-313, tensor_core.py:224-358).  Trees are index bookkeeping: validity is
checked by executing them with the numpy oracle against einsum / the dense
state-vector amplitude."""

import itertools as it
import math

import numpy as np
import pytest

from oracle import contract_np as cn
from quimb_b200 import tree, treeopt
from tests.circuit_util import random_grid_circuit_amplitude


def _run_ssa(inputs, output, sizes, ssa, arrays):
    tr = tree.Tree(inputs, output, sizes, ssa)
    nodes = dict(enumerate(arrays))
    inds = dict(enumerate(tr.inputs))
    for i, j, k, res in tr.steps:
        nodes[k] = cn.contract_pair(nodes.pop(i), inds[i], nodes.pop(j), inds[j], res)
        inds[k] = res
    (out,) = nodes.values()
    return out, tr


def _rand_network(rng, n, n_inds, hyper=False, n_out=2):
    """random connected-ish network over single-letter indices, sizes 2..4;
    ``hyper`` lets some indices live on three tensors (then they must be
    summed explicitly or kept as outputs)."""
    letters = [chr(ord("a") + i) for i in range(n_inds)]
    sizes = {ix: int(rng.integers(2, 5)) for ix in letters}
    inputs = [[] for _ in range(n)]
    for ix in letters:
        k = 3 if hyper and rng.random() < 0.25 else 2
        for t in rng.choice(n, size=min(k, n), replace=False):
            inputs[int(t)].append(ix)
    for t in inputs:
        if not t:
            t.append(letters[int(rng.integers(n_inds))])
    inputs = [tuple(t) for t in inputs]
    cnt = {}
    for t in inputs:
        for ix in t:
            cnt[ix] = cnt.get(ix, 0) + 1
    # outputs: a few indices, including (when present) a hyper one
    cands = [ix for ix in letters if ix in cnt]
    output = tuple(rng.choice(cands, size=min(n_out, len(cands)), replace=False))
    arrays = [rng.standard_normal([sizes[ix] for ix in t]) for t in inputs]
    return arrays, inputs, output, sizes


def _einsum_ref(arrays, inputs, output):
    sym = {}
    for t in inputs:
        for ix in t:
            sym.setdefault(ix, chr(ord("a") + len(sym)))
    eq = ",".join("".join(sym[ix] for ix in t) for t in inputs) + "->" + "".join(sym[ix] for ix in output)
    return np.einsum(eq, *arrays)


@pytest.mark.parametrize("hyper", [False, True])
def test_reconfigure_keeps_value_and_never_costs_more(hyper):
    rng = np.random.default_rng(11 + hyper)
    for trial in range(25):
        n = int(rng.integers(4, 11))
        arrays, inputs, output, sizes = _rand_network(rng, n, int(rng.integers(n, 2 * n)), hyper)
        ref = _einsum_ref(arrays, inputs, output)
        ssa0 = tree._greedy_ssa(inputs, output, sizes)
        c0, _ = treeopt.tree_stats(inputs, output, sizes, ssa0)
        t0 = tree.Tree(inputs, output, sizes, ssa0)
        # the bit-set bookkeeping agrees with Tree's own
        assert c0 == pytest.approx(math.log2(t0.contraction_cost()), abs=1e-9)
        for kw in (dict(subtree_size=4), dict(subtree_size=8), dict(subtree_size=6, seed=3),
                   dict(subtree_size=6, minimize="combo")):
            ssa = treeopt.reconfigure(inputs, output, sizes, ssa0, **kw)
            assert len(ssa) == n - 1
            out, tr = _run_ssa(inputs, output, sizes, ssa, arrays)
            np.testing.assert_allclose(out, ref, rtol=1e-10, atol=1e-10)
            if kw.get("minimize") != "combo":
                assert tr.contraction_cost() <= t0.contraction_cost() * (1 + 1e-9)
        # with all tensors in one subtree the result is the DP optimum
        if n <= 8:
            ssa = treeopt.reconfigure(inputs, output, sizes, ssa0, subtree_size=n)
            opt = tree.find_tree(inputs, output, sizes, "optimal")
            assert tree.Tree(inputs, output, sizes, ssa).contraction_cost() == opt.contraction_cost()


def test_simplify_spectral_growth_trees_are_valid():
    arrays, inputs, output, amp = random_grid_circuit_amplitude(3, 3, 8, seed=5)
    sizes = {ix: 2 for t in inputs for ix in t}
    prefix, red, ids = treeopt.simplify_inputs(inputs, output, sizes)
    # vectors and one-qubit gates are all absorbed: only two-qubit gates remain
    assert len(red) < len(inputs) // 3 and max(len(t) for t in red) <= 4
    for finder in (treeopt.spectral_ssa, treeopt.growth_ssa):
        sub = finder(red, output, sizes)
        full = treeopt.compose_ssa(prefix, len(inputs), ids, sub)
        assert len(full) == len(inputs) - 1
        out, tr = _run_ssa(inputs, output, sizes, full, arrays)
        assert abs(complex(out) - amp) < 1e-10
        # the prefix is free: the composed tree costs what the reduced one does
        # plus the (tiny) absorption steps
        c_red, w_red = treeopt.tree_stats(red, output, sizes, sub)
        assert tr.contraction_width() == w_red
    # open outputs / disconnected parts / a hyper index survive simplification
    ins = [("a", "b"), ("b", "c"), ("x", "y"), ("y",), ("c", "h"), ("h", "q"), ("h",)]
    out = ("a", "x", "q")
    sz = dict(a=2, b=3, c=4, x=5, y=6, h=3, q=2)
    rng = np.random.default_rng(0)
    arrs = [rng.standard_normal([sz[i] for i in t]) for t in ins]
    ref = np.einsum("ab,bc,xy,y,ch,hq,h->axq", *arrs)
    prefix, red, ids = treeopt.simplify_inputs(ins, out, sz)
    for finder in (treeopt.spectral_ssa, treeopt.growth_ssa):
        full = treeopt.compose_ssa(prefix, len(ins), ids, finder(red, out, sz))
        val, _ = _run_ssa(ins, out, sz, full, arrs)
        np.testing.assert_allclose(val, ref, rtol=1e-12)


def test_auto_hq_beats_greedy_on_a_deep_grid_circuit():
    # deep relative to its width: the shape BASELINE configs[3] has (6x6, depth 24)
    arrays, inputs, output, amp = random_grid_circuit_amplitude(3, 3, 16, seed=3)
    sizes = {ix: 2 for t in inputs for ix in t}
    g = tree.find_tree(inputs, output, sizes, "greedy")
    hq = tree.find_tree(inputs, output, sizes, "auto-hq")
    assert hq.contraction_cost() < g.contraction_cost()
    assert hq.contraction_width() <= 9 + 1e-9          # never wider than the state vector
    out, _ = _run_ssa(inputs, output, sizes, [(i, j) for i, j, _, _ in hq.steps], arrays)
    assert abs(complex(out) - amp) < 1e-10
    sp = tree.find_tree(inputs, output, sizes, "spectral")
    assert len(sp.steps) == len(inputs) - 1


def test_find_sliced_tree_meets_the_width_and_sums_to_the_amplitude():
    arrays, inputs, output, amp = random_grid_circuit_amplitude(3, 3, 10, seed=7)
    sizes = {ix: 2 for t in inputs for ix in t}
    full = tree.find_tree(inputs, output, sizes, "auto-hq")
    target = int(full.contraction_width()) - 3
    tr, sl = tree.find_sliced_tree(inputs, output, sizes, target, min_slices=16)
    assert tr.contraction_width() <= target + 1e-9
    assert len(set(sl)) == len(sl) and not set(sl) & set(output)
    assert 2 ** len(sl) >= 16
    total = 0.0
    for vals in it.product(*[range(2) for _ in sl]):
        fix = dict(zip(sl, vals))
        sub = [a[tuple(fix[ix] if ix in fix else slice(None) for ix in t)]
               for a, t in zip(arrays, inputs)]
        nodes = dict(enumerate(sub))
        inds = dict(enumerate(tr.inputs))
        for i, j, k, res in tr.steps:
            nodes[k] = cn.contract_pair(nodes.pop(i), inds[i], nodes.pop(j), inds[j], res)
            inds[k] = res
        (part,) = nodes.values()
        total += complex(part)
    assert abs(total - amp) < 1e-10
    # interleaved slicing + reconfiguration is never worse than slicing the
    # finished tree
    sl0, n0, w0, cost0 = tree.find_slices(full, target_width=target)
    tr2, sl2 = tree.find_sliced_tree(inputs, output, sizes, target)
    assert tr2.contraction_cost() * 2 ** len(sl2) <= cost0 * n0 * 1.0001


def test_anneal_keeps_value_and_repairs_a_bad_tree():
    rng = np.random.default_rng(21)
    for hyper in (False, True):
        for trial in range(10):
            n = int(rng.integers(5, 11))
            arrays, inputs, output, sizes = _rand_network(rng, n, int(rng.integers(n, 2 * n)), hyper)
            ref = _einsum_ref(arrays, inputs, output)
            ssa0 = tree._greedy_ssa(inputs, output, sizes)
            ssa = treeopt.anneal(inputs, output, sizes, ssa0, sweeps=40, seed=trial)
            out, tr = _run_ssa(inputs, output, sizes, ssa, arrays)
            np.testing.assert_allclose(out, ref, rtol=1e-10, atol=1e-10)
            # the best tree seen is kept: never worse than the start
            assert tr.contraction_cost() <= tree.Tree(inputs, output, sizes, ssa0).contraction_cost() * (1 + 1e-9)
            ssa_w = treeopt.anneal(inputs, output, sizes, ssa0, sweeps=20, seed=1, target_width=3.0)
            out, _ = _run_ssa(inputs, output, sizes, ssa_w, arrays)
            np.testing.assert_allclose(out, ref, rtol=1e-10, atol=1e-10)
    # a deep grid circuit from a deliberately poor start (plain greedy):
    # annealing alone recovers most of the gap to the refined tree
    arrays, inputs, output, amp = random_grid_circuit_amplitude(3, 3, 16, seed=3)
    sizes = {ix: 2 for t in inputs for ix in t}
    prefix, red, ids = treeopt.simplify_inputs(inputs, output, sizes)
    bad = tree._greedy_ssa(red, output, sizes)
    c_bad, _ = treeopt.tree_stats(red, output, sizes, bad)
    good = treeopt.anneal(red, output, sizes, bad, sweeps=200, seed=0)
    c_good, w_good = treeopt.tree_stats(red, output, sizes, good)
    assert c_good <= c_bad and w_good <= 9 + 1e-9
    full = treeopt.compose_ssa(prefix, len(inputs), ids, good)
    out, _ = _run_ssa(inputs, output, sizes, full, arrays)
    assert abs(complex(out) - amp) < 1e-10


def test_tree_search_does_not_depend_on_the_hash_seed():
    """Ranks of a multi-GPU run each search for the tree and the slices; with
    string index names kept in sets the result would follow PYTHONHASHSEED."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = "\n".join([
        "import sys; sys.path.insert(0, %r)" % root,
        "from tests.circuit_util import random_grid_circuit_amplitude",
        "from quimb_b200 import tree",
        "a, inputs, output, _ = random_grid_circuit_amplitude(3, 3, 10, seed=3, dense=False)",
        "sd = {ix: 2 for t in inputs for ix in t}",
        "tr, sl = tree.find_sliced_tree(inputs, output, sd, 6)",
        "r = tree.find_tree(inputs, output, sd, 'random-greedy')",
        "print(sl, repr(tr.steps), repr(r.steps))",
    ])
    outs = set()
    for seed in ("0", "1", "12345"):
        env = dict(os.environ, PYTHONHASHSEED=seed)
        res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True,
                             timeout=600)
        assert res.returncode == 0, res.stderr[-2000:]
        outs.add(res.stdout)
    assert len(outs) == 1


def test_auto_hq_and_sliced_search_on_random_hyper_networks():
    """the full pipelines (simplify -> greedy / spectral -> anneal -> reconfigure,
    and slicing interleaved with re-optimisation) on random networks with
    hyper-indices, open outputs and mixed index sizes."""
    rng = np.random.default_rng(99)
    for trial in range(16):
        n = int(rng.integers(10, 15))
        arrays, inputs, output, sizes = _rand_network(rng, n, int(rng.integers(n, 2 * n)),
                                                      bool(trial % 2), n_out=int(rng.integers(0, 4)))
        ref = cn.array_contract(arrays, inputs, output, "greedy")
        for opt in ("auto-hq", "spectral"):
            tr = tree.find_tree(inputs, output, sizes, opt)
            out, _ = _run_ssa(inputs, output, sizes, [(i, j) for i, j, _, _ in tr.steps], arrays)
            np.testing.assert_allclose(out, ref, rtol=1e-9, atol=1e-9)
        w = tree.find_tree(inputs, output, sizes, "auto-hq").contraction_width()
        tr, sl = tree.find_sliced_tree(inputs, output, sizes, max(1.0, w - 2))
        assert not set(sl) & set(output)
        total = 0
        for vals in it.product(*[range(sizes[ix]) for ix in sl]):
            fix = dict(zip(sl, vals))
            sub = [a[tuple(fix[ix] if ix in fix else slice(None) for ix in t)]
                   for a, t in zip(arrays, inputs)]
            nodes = dict(enumerate(sub))
            inds = dict(enumerate(tr.inputs))
            for i, j, k, res in tr.steps:
                nodes[k] = cn.contract_pair(nodes.pop(i), inds[i], nodes.pop(j), inds[j], res)
                inds[k] = res
            (part,) = nodes.values()
            total = total + part
        np.testing.assert_allclose(total, ref, rtol=1e-9, atol=1e-9)


def test_tree_traffic_and_peak_bookkeeping():
    # chain A(ab) B(bc) C(cd) contracted left to right, sizes 2, 3, 4, 5
    inputs, output = [("a", "b"), ("b", "c"), ("c", "d")], ("a", "d")
    sizes = dict(a=2, b=3, c=4, d=5)
    ssa = [(0, 1), (3, 2)]
    # node 3 = (a, c): reads 6 + 12, writes 8; node 4 = (a, d): reads 8 + 20, writes 10
    assert treeopt.tree_traffic(inputs, output, sizes, ssa) == pytest.approx(math.log2(6 + 12 + 8 + 8 + 20 + 10))
    # alive: inputs 6 + 12 + 20 = 38; + node 3 (8) = 46 is the peak; then 38 - 18 + 8 = 28, + 10 = 38
    assert treeopt.tree_peak(inputs, output, sizes, ssa) == pytest.approx(math.log2(46))
    c, w = treeopt.tree_stats(inputs, output, sizes, ssa)
    assert c == pytest.approx(math.log2(2 * 3 * 4 + 2 * 4 * 5)) and w == pytest.approx(math.log2(20))
