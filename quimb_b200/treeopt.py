"""Contraction-tree refinement (host-only; the role cotengra's
``subtree_reconfigure`` / ``slicing_reconfigure`` play for
``quimb.tensor.contraction.array_contract_tree``, contraction.py:302-313, whose
trees drive ``tensor_contract``, tensor_core.py:224-358).

The engine takes external trees unchanged; this module exists because the
reference's tree finder (cotengra) is a third-party package that is not
installed here, and BASELINE configs[3] (6 x 6 qubits, depth 24) needs a tree
far better than greedy.  Everything is index bookkeeping on Python integers
used as bit sets: one bit per index, a node's legs = one int.

    reconfigure(...)   repeatedly cut a subtree of <= ``subtree_size`` frontier
                       nodes out of the tree and replace its inside by the
                       exact dynamic-programming optimum for those frontier
                       tensors (an index leaving the subtree counts as output)
    slice_and_reconfigure(...)  alternate greedy slicing with reconfiguration of the
                       sliced network until the width target is met

Trees are exchanged as SSA step lists like everywhere in ``tree.py``.
"""

import math
import random

__all__ = ["reconfigure", "anneal", "slice_and_reconfigure", "tree_stats", "tree_traffic", "tree_peak",
           "spectral_ssa", "growth_ssa", "simplify_inputs", "compose_ssa"]


class _Bits:
    """index <-> bit table with log2 sizes."""

    def __init__(self, inputs, output, size_dict):
        self.bit = {}
        for t in inputs:
            for ix in t:
                if ix not in self.bit:
                    self.bit[ix] = len(self.bit)
        for ix in output:
            if ix not in self.bit:
                self.bit[ix] = len(self.bit)
        self.names = list(self.bit)
        self.w = [math.log2(size_dict[ix]) for ix in self.names]
        self.uniform = len(set(self.w)) <= 1
        self.lw = self.w[0] if self.w else 0.0

    def mask(self, inds):
        m = 0
        for ix in inds:
            m |= 1 << self.bit[ix]
        return m

    def lsize(self, m):
        if self.uniform:
            return m.bit_count() * self.lw
        s, w = 0.0, self.w
        while m:
            lb = m & -m
            s += w[lb.bit_length() - 1]
            m ^= lb
        return s


class _BinTree:
    """Mutable binary contraction tree over leaf tensors 0..n-1."""

    def __init__(self, inputs, output, size_dict, ssa):
        self.bits = B = _Bits(inputs, output, size_dict)
        n = self.n = len(inputs)
        self.out_mask = B.mask(output)
        leaf_masks = [B.mask(t) for t in inputs]
        # appearance counts: an index is contracted at the node under which
        # all of its appearances have been gathered (and it is not an output)
        total = {}
        for t in inputs:
            for ix in set(t):
                total[ix] = total.get(ix, 0) + 1
        self.hyper = 0
        for ix, c in total.items():
            if c > 2:
                self.hyper |= 1 << B.bit[ix]
        self.total = {B.bit[ix]: c for ix, c in total.items() if c > 2}
        self.children = {}
        self.legs = {k: m for k, m in enumerate(leaf_masks)}
        self.hcnt = {}
        for k, m in enumerate(leaf_masks):
            self.hcnt[k] = self._leaf_hcnt(m)
        self.next_id = n
        for i, j in ssa:
            self._merge(i, j)
        self.root = self.next_id - 1 if ssa else 0

    def _leaf_hcnt(self, m):
        h = m & self.hyper
        d = {}
        while h:
            lb = h & -h
            d[lb.bit_length() - 1] = 1
            h ^= lb
        return d

    def _merged_legs(self, li, lj, hi, hj):
        """legs of the contraction of two nodes + its hyper-index counts."""
        shared = li & lj
        gone = shared & ~self.out_mask & ~self.hyper
        hc = None
        if hi or hj:
            hc = dict(hi)
            for b, c in hj.items():
                hc[b] = hc.get(b, 0) + c
            for b, c in list(hc.items()):
                if c == self.total[b] and not (self.out_mask >> b & 1):
                    gone |= 1 << b
                    del hc[b]
        return (li | lj) & ~gone, (hc or {})

    def _merge(self, i, j):
        k = self.next_id
        self.next_id += 1
        self.children[k] = (i, j)
        self.legs[k], self.hcnt[k] = self._merged_legs(
            self.legs[i], self.legs[j], self.hcnt[i], self.hcnt[j])
        return k

    def node_cost(self, k):
        i, j = self.children[k]
        return 2.0 ** self.bits.lsize(self.legs[i] | self.legs[j])

    def total_cost(self):
        return sum(self.node_cost(k) for k in self.children)

    def width(self):
        return max((self.bits.lsize(m) for m in self.legs.values()), default=0.0)

    def to_ssa(self):
        """post-order SSA steps (iterative: trees can be deep chains)."""
        ssa, ids = [], {}
        nxt = self.n
        stack = [(self.root, False)]
        while stack:
            k, seen = stack.pop()
            if k not in self.children:
                ids[k] = k
                continue
            i, j = self.children[k]
            if not seen:
                stack.append((k, True))
                stack.append((j, False))
                stack.append((i, False))
            else:
                ssa.append((ids[i], ids[j]))
                ids[k] = nxt
                nxt += 1
        return ssa


def _dp_optimal(legs, out_legs, lsize, size_weight=0.0):
    """Exact optimum over all binary trees on ``legs`` (<= ~12 nodes).  An
    index is kept by a subset when it also lives outside it or in
    ``out_legs``.  Returns (cost, nested pair structure over 0..k-1)."""
    k = len(legs)
    full = (1 << k) - 1
    inside = [0] * (full + 1)
    for m in range(1, full + 1):
        lb = m & -m
        inside[m] = inside[m ^ lb] | legs[lb.bit_length() - 1]
    sub_legs = [0] * (full + 1)
    for m in range(1, full + 1):
        sub_legs[m] = inside[m] & (inside[full ^ m] | out_legs)
    best = [0.0] * (full + 1)
    how = [None] * (full + 1)
    pow2 = {}
    for m in range(1, full + 1):
        if m & (m - 1) == 0:
            continue
        low = m & -m
        bc, bh = None, None
        a = (m - 1) & m
        while a:
            if a & low:
                b = m ^ a
                la, lb_ = sub_legs[a], sub_legs[b]
                u = la | lb_
                c = pow2.get(u)
                if c is None:
                    c = pow2[u] = 2.0 ** lsize(u)
                c += best[a] + best[b]
                if bc is None or c < bc:
                    bc, bh = c, (a, b)
            a = (a - 1) & m
        if size_weight:
            bc += size_weight * 2.0 ** lsize(sub_legs[m])
        best[m], how[m] = bc, bh
    return best[full], how


def _subtree_cost(bt, root, frontier, size_weight):
    """cost of the internal nodes between ``root`` and ``frontier``."""
    fr = set(frontier)
    tot, internal = 0.0, []
    stack = [root]
    while stack:
        k = stack.pop()
        if k in fr:
            continue
        internal.append(k)
        tot += bt.node_cost(k)
        if size_weight:
            tot += size_weight * 2.0 ** bt.bits.lsize(bt.legs[k])
        stack.extend(bt.children[k])
    return tot, internal


def _grow(bt, root, size, rng=None):
    """frontier of <= ``size`` nodes below ``root``: keep opening the largest
    (or a random) non-leaf frontier node."""
    frontier = list(bt.children[root])
    while len(frontier) < size:
        cands = [k for k in frontier if k in bt.children]
        if not cands:
            break
        if rng is None:
            k = max(cands, key=lambda c: bt.bits.lsize(bt.legs[c]))
        else:
            k = rng.choice(cands)
        frontier.remove(k)
        frontier.extend(bt.children[k])
    return frontier


def _replace(bt, root, frontier, internal, how):
    """rebuild the inside of a subtree from the DP solution; ``root`` keeps
    its id (its legs do not change), so nothing above it is touched."""
    for k in internal:
        if k != root:
            del bt.children[k], bt.legs[k], bt.hcnt[k]
    del bt.children[root]
    full = (1 << len(frontier)) - 1
    new_nodes = []

    def build(m):
        if m & (m - 1) == 0:
            return frontier[m.bit_length() - 1]
        a, b = how[m]
        ia, ib = build(a), build(b)
        if m == full:
            bt.children[root] = (ia, ib)
            return root
        k = bt._merge(ia, ib)
        new_nodes.append(k)
        return k

    build(full)
    return new_nodes


def _reconfigure_bt(bt, subtree_size=8, max_rounds=8, size_weight=0.0, seed=None,
                    max_time=None, rtol=1e-9):
    import time
    t0 = time.time()
    rng = random.Random(seed) if seed is not None else None
    lsize = bt.bits.lsize
    for rnd in range(max_rounds):
        improved = 0.0
        order = sorted(bt.children, key=bt.node_cost, reverse=True)
        for root in order:
            if root not in bt.children:
                continue
            frontier = _grow(bt, root, subtree_size, rng)
            if len(frontier) < 3:
                continue
            old, internal = _subtree_cost(bt, root, frontier, size_weight)
            new, how = _dp_optimal([bt.legs[f] for f in frontier], bt.legs[root],
                                   lsize, size_weight)
            if size_weight:
                # the DP also charged the root's own size, the walk did too
                pass
            if new < old * (1.0 - rtol):
                _replace(bt, root, frontier, internal, how)
                improved += old - new
            if max_time is not None and time.time() - t0 > max_time:
                return bt
        if improved == 0.0:
            if rng is None:
                break
    return bt


def reconfigure(inputs, output, size_dict, ssa, subtree_size=8, max_rounds=8,
                minimize="flops", seed=None, max_time=None):
    """Refine an SSA tree by exact re-optimisation of its subtrees.
    ``minimize`` = 'flops' or 'combo' (flops + 64 x intermediate sizes, the
    memory-traffic-aware objective).  ``seed`` switches the subtree growth
    from largest-first to random (use after the deterministic pass has
    converged).  Returns the new SSA steps."""
    if len(inputs) < 3:
        return list(ssa)
    bt = _BinTree(inputs, output, size_dict, ssa)
    _reconfigure_bt(bt, subtree_size, max_rounds, _size_weight(minimize), seed, max_time)
    return bt.to_ssa()


def tree_stats(inputs, output, size_dict, ssa):
    """(log2 total multiply-adds, log2 largest intermediate)."""
    bt = _BinTree(inputs, output, size_dict, ssa)
    c = bt.total_cost()
    return (math.log2(c) if c > 0 else 0.0), bt.width()


def tree_traffic(inputs, output, size_dict, ssa):
    """log2 of the elements moved by the pairwise executor: every node reads
    its two operands and writes its result (the algorithmic bytes of a tree
    are this times the item size)."""
    bt = _BinTree(inputs, output, size_dict, ssa)
    lsize = bt.bits.lsize
    tot = 0.0
    for k, (i, j) in bt.children.items():
        tot += 2.0 ** lsize(bt.legs[i]) + 2.0 ** lsize(bt.legs[j]) + 2.0 ** lsize(bt.legs[k])
    return math.log2(tot) if tot > 0 else 0.0


def tree_peak(inputs, output, size_dict, ssa):
    """log2 of the largest number of elements alive at once when the tree is
    executed step by step (operands are released after the step that consumes
    them): the memory the executor needs, in elements."""
    bt = _BinTree(inputs, output, size_dict, ssa)
    lsize = bt.bits.lsize
    size = {k: 2.0 ** lsize(m) for k, m in bt.legs.items()}
    live = sum(size[k] for k in range(bt.n))
    peak = live
    # replay in SSA order (ids grow with the steps)
    for k in sorted(bt.children):
        i, j = bt.children[k]
        peak = max(peak, live + size[k])
        live += size[k] - size[i] - size[j]
    return math.log2(peak) if peak > 0 else 0.0


def _greedy_slices(bt, target_width, fixed=(), max_new=None):
    """indices to slice, chosen one at a time: the index on the widest
    intermediates whose removal lowers (width, total cost x slices) most."""
    lsize = bt.bits.lsize
    removed = 0
    for b in fixed:
        removed |= 1 << b
    chosen = list(fixed)
    nodes = list(bt.children)

    def measure(rem):
        keep = ~rem
        width, cost = 0.0, 0.0
        for k in nodes:
            i, j = bt.children[k]
            width = max(width, lsize(bt.legs[k] & keep))
            cost += 2.0 ** lsize((bt.legs[i] | bt.legs[j]) & keep)
        return width, cost

    width, cost = measure(removed)
    while width > target_width + 1e-9:
        if max_new is not None and len(chosen) - len(fixed) >= max_new:
            break
        cands = 0
        for k in nodes:
            m = bt.legs[k] & ~removed
            if lsize(m) >= width - 1e-9:
                cands |= m
        cands &= ~bt.out_mask
        if not cands:
            break
        best = None
        c = cands
        while c:
            lb = c & -c
            c ^= lb
            b = lb.bit_length() - 1
            if bt.bits.w[b] <= 0.0:
                continue
            w2, c2 = measure(removed | lb)
            key = (w2, c2 * 2.0 ** bt.bits.w[b])
            if best is None or key < best[0]:
                best = (key, b, w2, c2)
        if best is None:
            break
        _, b, width, cost = best
        removed |= 1 << b
        chosen.append(b)
    return chosen, width, cost


def slice_and_reconfigure(inputs, output, size_dict, ssa, target_width,
                          subtree_size=8, step=1, minimize="flops", max_rounds=2,
                          anneal_sweeps=300, max_time=None):
    """Interleave slicing with re-optimisation (the scheme of cotengra's
    ``slicing_reconfigure``): slice the ``step`` best indices for the current
    tree, then anneal + reconfigure the tree of the *sliced* network -- the one
    that is executed once per slice -- with intermediates wider than the
    target charged extra, and repeat until the largest intermediate has at
    most ``2**target_width`` elements.  The width target of each round is
    lowered one bit at a time (never below ``target_width``): annealing
    against the final target from the start distorts the early rounds.
    Returns ``(ssa, sliced_inds)``; the tree is valid for the unsliced network
    as well (same leaves)."""
    import time
    t0 = time.time()
    sliced = []
    wprev = None
    while True:
        sl = set(sliced)
        red_inputs = [tuple(ix for ix in t if ix not in sl) for t in inputs]
        red_output = tuple(ix for ix in output if ix not in sl)
        bt = _BinTree(red_inputs, red_output, size_dict, ssa)
        if sliced:
            left = None if max_time is None else max(1.0, max_time - (time.time() - t0))
            if anneal_sweeps and len(inputs) >= 4:
                _anneal_bt(bt, anneal_sweeps, 0.5, 0.02, max(target_width, wprev - 1.0),
                           seed=len(sliced), max_time=left, size_weight=_size_weight(minimize))
            _reconfigure_bt(bt, subtree_size, max_rounds, _size_weight(minimize), None, left)
            ssa = bt.to_ssa()
            bt = _BinTree(red_inputs, red_output, size_dict, ssa)
        width = wprev = bt.width()
        if width <= target_width + 1e-9:
            break
        bits, _, _ = _greedy_slices(bt, target_width, max_new=step)
        if not bits:
            break
        sliced.extend(bt.bits.names[b] for b in bits)
    return ssa, tuple(sliced)


# ----------------------------------------------------------- region growth --
def growth_ssa(inputs, output, size_dict, start=None, temperature=0.0, rng=None):
    """'Boundary sweep' tree: grow ONE contracted region a tensor at a time,
    always absorbing the neighbour that leaves the smallest region boundary
    (ties: cheapest step).  This is the shape deep circuits and long strips
    want -- a greedy pair-merger builds many blobs whose final joins are far
    too wide.  The caterpillar it returns is a starting point for
    ``reconfigure`` (which makes it bushy where that pays)."""
    n = len(inputs)
    if n < 2:
        return []
    bt = _BinTree(inputs, output, size_dict, [])
    lsize = bt.bits.lsize
    where = {}
    for k in range(n):
        m = bt.legs[k]
        while m:
            lb = m & -m
            m ^= lb
            where.setdefault(lb, []).append(k)

    def neighbours(k):
        out, m = set(), bt.legs[k]
        while m:
            lb = m & -m
            m ^= lb
            out.update(where[lb])
        out.discard(k)
        return out

    left = set(range(n))
    if start is None:
        start = min(left, key=lambda k: (lsize(bt.legs[k]), k))
    region = start
    left.discard(start)
    cands = neighbours(start) & left
    ssa = []
    while left:
        if not cands:
            # disconnected: outer product with the smallest remaining tensor
            t = min(left, key=lambda k: (lsize(bt.legs[k]), k))
        else:
            best = None
            lr, hr = bt.legs[region], bt.hcnt[region]
            for t in cands:
                lt = bt.legs[t]
                new, _ = bt._merged_legs(lr, lt, hr, bt.hcnt[t])
                sc = lsize(new) + 1e-3 * lsize(lr | lt)
                if temperature > 0.0:
                    sc += temperature * rng.gauss(0.0, 1.0)
                if best is None or (sc, t) < best:
                    best = (sc, t)
            t = best[1]
        ssa.append((region, t))
        region = bt._merge(region, t)
        left.discard(t)
        cands.discard(t)
        cands |= neighbours(t) & left
    return ssa


def simplify_inputs(inputs, output, size_dict):
    """Rank simplification as a tree prefix (what quimb's ``rank_simplify``
    does to the network before contraction, tensor_core.py:10086-10250):
    absorb a tensor into a neighbour whenever the product is no larger than
    the larger of the two (vectors, one-qubit gates, ...), so this is free.  Returns
    ``(ssa_prefix, reduced_inputs, ids)`` with ``ids[k]`` the SSA id of
    reduced tensor k."""
    n = len(inputs)
    out_set = set(output)
    sets = {k: set(t) for k, t in enumerate(inputs)}
    order = {k: tuple(t) for k, t in enumerate(inputs)}
    where = {}
    for k, t in sets.items():
        for ix in t:
            where.setdefault(ix, set()).add(k)
    ssa, nxt = [], n
    lsz = lambda t: sum(math.log2(size_dict[ix]) for ix in t)  # noqa: E731
    queue = sorted(sets, key=lambda k: (len(sets[k]), k))
    alive = set(sets)
    changed = True
    while changed:
        changed = False
        for k in list(queue):
            if k not in alive:
                continue
            sk = sets[k]
            nb = set()
            for ix in sk:
                nb |= where[ix]
            nb.discard(k)
            tgt = None
            szk = lsz(sk)
            for j in sorted(nb):
                res = [ix for ix in sk | sets[j]
                       if ix in out_set or (where[ix] - {k, j})]
                if lsz(res) <= max(szk, lsz(sets[j])) + 1e-12:
                    tgt = j
                    break
            if tgt is None:
                continue
            a, b = (k, tgt)
            keep = [ix for ix in dict.fromkeys(order[b] + order[a])
                    if ix in out_set or (where[ix] - {a, b})]
            for z in (a, b):
                for ix in sets[z]:
                    where[ix].discard(z)
                alive.discard(z)
                del sets[z], order[z]
            ssa.append((a, b))
            sets[nxt], order[nxt] = set(keep), tuple(keep)
            for ix in keep:
                where[ix].add(nxt)
            alive.add(nxt)
            queue.append(nxt)
            nxt += 1
            changed = True
        queue = sorted(alive, key=lambda k: (len(sets[k]), k))
    ids = sorted(alive)
    return ssa, [order[k] for k in ids], ids


def compose_ssa(prefix, n, ids, sub_ssa):
    """SSA steps of (prefix, then a tree over the reduced tensors): sub-tree
    leaf k is SSA id ``ids[k]``, its new nodes follow the prefix's."""
    base = n + len(prefix)
    m = len(ids)
    out = list(prefix)
    for i, j in sub_ssa:
        ii = ids[i] if i < m else base + (i - m)
        jj = ids[j] if j < m else base + (j - m)
        out.append((ii, jj))
    return out


def spectral_ssa(inputs, output, size_dict):
    """Caterpillar tree along the Fiedler vector of the tensor graph (edge
    weight = log2 of the shared index sizes): a linear arrangement with a
    small cut everywhere, i.e. a sweep along the network's longest direction
    (time for a deep circuit, the long axis of a strip).  Host-only numpy."""
    import numpy as np
    n = len(inputs)
    if n < 3:
        return [(0, 1)] if n == 2 else []
    where = {}
    for k, t in enumerate(inputs):
        for ix in set(t):
            where.setdefault(ix, []).append(k)
    W = np.zeros((n, n))
    for ix, ts in where.items():
        w = math.log2(size_dict[ix])
        for a in ts:
            for b in ts:
                if a != b:
                    W[a, b] += w
    lap = np.diag(W.sum(1)) - W
    if n <= 3000:
        _, vec = np.linalg.eigh(lap)
        f = vec[:, 1]
    else:
        from scipy.sparse.linalg import eigsh
        _, vec = eigsh(lap, k=2, sigma=-1e-3, which="LM")
        f = vec[:, 1]
    order = [int(k) for k in np.argsort(f, kind="stable")]
    ssa, cur, nxt = [], order[0], n
    for k in order[1:]:
        ssa.append((cur, k))
        cur = nxt
        nxt += 1
    return ssa


# ------------------------------------------------------ simulated annealing --
def _anneal_bt(bt, sweeps=200, t_start=1.0, t_end=0.02, target_width=None, seed=0,
               max_time=None, size_weight=0.0):
    """Metropolis over local tree rotations (the move set of cotengra's
    ``simulated_anneal_tree``): at a node p = (l, r) with l = (a, b) the
    sibling r is exchanged with a or b, which changes only the contraction
    that forms l.  Scores are log2 of the local cost; an intermediate wider
    than ``target_width`` is charged 2^(excess) extra.  Keeps the best tree
    seen.  In place; returns the best total cost."""
    import time
    rng = random.Random(seed)
    lsize = bt.bits.lsize
    t0 = time.time()

    def step_cost(li, lj, legs_out):
        """linear cost of one contraction: multiply-adds + size_weight x the
        elements of its result (written once, read once by its consumer)."""
        c = 2.0 ** lsize(li | lj)
        so = lsize(legs_out)
        if size_weight:
            c += size_weight * 2.0 ** so
        if target_width is not None and so > target_width:
            c *= 4.0 ** (so - target_width)
        return c

    def tree_score():
        tot = 0.0
        for k, (i, j) in bt.children.items():
            tot += step_cost(bt.legs[i], bt.legs[j], bt.legs[k])
        return tot

    cur = tree_score()
    best = cur
    best_children = dict(bt.children)
    nodes = [k for k in bt.children]
    for s in range(sweeps):
        T = t_start * (t_end / t_start) ** (s / max(1, sweeps - 1))
        rng.shuffle(nodes)
        for p in nodes:
            l, r = bt.children[p]
            if rng.random() < 0.5:
                l, r = r, l
            if l not in bt.children:
                l, r = r, l
                if l not in bt.children:
                    continue
            a, b = bt.children[l]
            if rng.random() < 0.5:
                a, b = b, a
            # p = ((a, b), r)  ->  ((a, r), b)
            la, lb, lr = bt.legs[a], bt.legs[b], bt.legs[r]
            old_l = bt.legs[l]
            new_l, new_h = bt._merged_legs(la, lr, bt.hcnt[a], bt.hcnt[r])
            lp = bt.legs[p]
            old = step_cost(la, lb, old_l) + step_cost(old_l, lr, lp)
            new = step_cost(la, lr, new_l) + step_cost(new_l, lb, lp)
            if new > old:
                d = math.log2(new) - math.log2(old)
                if rng.random() >= math.exp(-d / T):
                    continue
            bt.children[l] = (a, r)
            bt.legs[l], bt.hcnt[l] = new_l, new_h
            bt.children[p] = (l, b)
            cur += new - old
            if cur < best * (1.0 - 1e-12):
                best = cur
                best_children = dict(bt.children)
        if max_time is not None and time.time() - t0 > max_time:
            break
    # restore the best tree and recompute its legs bottom-up
    bt.children = best_children
    order, stack = [], [bt.root]
    while stack:
        k = stack.pop()
        if k in bt.children:
            order.append(k)
            stack.extend(bt.children[k])
    for k in reversed(order):
        i, j = bt.children[k]
        bt.legs[k], bt.hcnt[k] = bt._merged_legs(bt.legs[i], bt.legs[j], bt.hcnt[i], bt.hcnt[j])
    return best


def _size_weight(minimize):
    """'flops' -> 0; 'combo' -> 64 (cotengra's convention: multiply-adds +
    64 x intermediate elements); ('combo', w) -> w."""
    if minimize in (None, "flops"):
        return 0.0
    if minimize == "combo":
        return 64.0
    if isinstance(minimize, (tuple, list)) and minimize[0] == "combo":
        return float(minimize[1])
    raise ValueError(f"unknown objective {minimize!r}")


def anneal(inputs, output, size_dict, ssa, sweeps=200, t_start=1.0, t_end=0.02,
           target_width=None, seed=0, max_time=None, minimize="flops"):
    """Simulated annealing of an SSA tree (see :func:`_anneal_bt`), followed by
    nothing else -- combine with :func:`reconfigure`.  Returns new SSA steps."""
    if len(inputs) < 4:
        return list(ssa)
    bt = _BinTree(inputs, output, size_dict, ssa)
    _anneal_bt(bt, sweeps, t_start, t_end, target_width, seed, max_time, _size_weight(minimize))
    return bt.to_ssa()
