"""ORACLE (test infrastructure, never imported by the product package).

Plain-numpy restatement of the contraction half of quimb's hot path:

  * output-index rule and error behaviour of ``tensor_contract``
    (quimb/tensor/tensor_core.py:158-170, 300-306);
  * the pairwise executor that cotengra 0.8.2 (third-party, pinned in the
    reference's pixi.lock:1022, absent offline) runs for
    ``array_contract`` (quimb/tensor/contraction.py:272-292): for every node
    of a contraction tree one ``tensordot`` (+ ``transpose``) when the pair
    is a pure tensordot, otherwise one ``einsum``;
  * a contraction-path finder: optimal (dynamic programming over subsets,
    the classic opt_einsum "optimal"/"dp" objective = total flops) for small
    networks, greedy (opt_einsum's published heuristic: prefer the pair
    minimising size(out) - size(a) - size(b)) otherwise.

Parity note: no reference test pins a contraction *path* (SURVEY.md 8c), so
results are compared numerically, index bookkeeping exactly.
"""

import itertools
import math

import numpy as np


def gen_output_inds(all_inds):
    """Indices appearing exactly once, in first-appearance order; raise if any
    index appears more than twice (tensor_core.py:158-170)."""
    freq = {}
    for ix in all_inds:
        freq[ix] = freq.get(ix, 0) + 1
    out = []
    for ix, f in freq.items():
        if f > 2:
            raise ValueError(
                f"The index {ix} appears more than twice! If this is "
                "intentionally a 'hyper' tensor network you will need to "
                "explicitly supply `output_inds` when contracting for example."
            )
        if f == 1:
            out.append(ix)
    return tuple(out)


# --------------------------------------------------------------- paths -----
def _pair_result(ia, ib, keep):
    """Indices of the pairwise result: those of a|b still needed elsewhere,
    ordered a-first then b (cotengra's convention)."""
    out = [ix for ix in ia if ix in keep]
    out += [ix for ix in ib if ix in keep and ix not in ia]
    return tuple(out)


def _needed_elsewhere(terms, skip, output):
    need = set(output)
    for j, t in enumerate(terms):
        if j not in skip and t is not None:
            need.update(t)
    return need


def _flops(ia, ib, size_dict):
    allix = set(ia) | set(ib)
    return math.prod(size_dict[ix] for ix in allix)


def _size(ix, size_dict):
    return math.prod(size_dict[i] for i in ix)


def find_path(inputs, output, size_dict, optimize="auto"):
    """Return an ssa-free linear path [(i, j), ...] in opt_einsum convention
    (positions refer to the *current* list, contracted terms are removed and
    the result appended)."""
    n = len(inputs)
    if n <= 1:
        return []
    if n == 2:
        return [(0, 1)]
    if isinstance(optimize, (list, tuple)):
        return [tuple(p) for p in optimize]
    if optimize in ("optimal", "dp") or (
        optimize in ("auto", "auto-hq") and n <= 8
    ):
        return _path_optimal(inputs, output, size_dict)
    return _path_greedy(inputs, output, size_dict)


def _path_greedy(inputs, output, size_dict):
    """opt_einsum-style greedy: among pairs sharing an index take the one
    minimising size(out) - size(a) - size(b); outer products only when no
    connected pair is left.  Candidate pairs come from an index -> terms map."""
    terms = {i: tuple(t) for i, t in enumerate(inputs)}
    order = list(range(len(inputs)))          # current positions (linear path)
    where = {}
    for i, t in terms.items():
        for ix in set(t):
            where.setdefault(ix, set()).add(i)
    out_set = set(output)
    nxt = len(inputs)
    path = []
    while len(terms) > 1:
        cands = set()
        for ix, ts in where.items():
            if len(ts) >= 2:
                tl = sorted(ts)
                for x in range(len(tl)):
                    for y in range(x + 1, len(tl)):
                        cands.add((tl[x], tl[y]))
        if not cands:
            ks = sorted(terms, key=lambda k: (_size(terms[k], size_dict), k))[:2]
            cands = {(min(ks), max(ks))}
        best = None
        for i, j in cands:
            res = tuple(ix for ix in dict.fromkeys(terms[i] + terms[j])
                        if ix in out_set or (where[ix] - {i, j}))
            key = (_size(res, size_dict) - _size(terms[i], size_dict) - _size(terms[j], size_dict),
                   _flops(terms[i], terms[j], size_dict), i, j)
            if best is None or key < best[0]:
                best = (key, i, j, res)
        _, i, j, res = best
        pi, pj = order.index(i), order.index(j)
        path.append((min(pi, pj), max(pi, pj)))
        order = [k for k in order if k not in (i, j)] + [nxt]
        for k in (i, j):
            for ix in set(terms[k]):
                where[ix].discard(k)
            del terms[k]
        terms[nxt] = res
        for ix in set(res):
            where.setdefault(ix, set()).add(nxt)
        nxt += 1
    return path


def _path_optimal(inputs, output, size_dict):
    n = len(inputs)
    inputs = [tuple(t) for t in inputs]
    full = (1 << n) - 1
    # indices of a subset's intermediate = indices that also appear outside
    def sub_inds(mask):
        inside, outside = [], set(output)
        for k in range(n):
            if mask >> k & 1:
                for ix in inputs[k]:
                    if ix not in inside:
                        inside.append(ix)
            else:
                outside.update(inputs[k])
        return tuple(ix for ix in inside if ix in outside)

    inds = {1 << k: inputs[k] for k in range(n)}
    best = {1 << k: (0, None) for k in range(n)}
    for size in range(2, n + 1):
        for combo in itertools.combinations(range(n), size):
            mask = sum(1 << k for k in combo)
            inds[mask] = sub_inds(mask)
            bestc = None
            # enumerate proper sub-splits (each once)
            sub = (mask - 1) & mask
            while sub:
                other = mask ^ sub
                if sub < other:
                    sub = (sub - 1) & mask
                    continue
                c = (
                    best[sub][0]
                    + best[other][0]
                    + _flops(inds[sub], inds[other], size_dict)
                )
                if bestc is None or c < bestc[0]:
                    bestc = (c, (sub, other))
                sub = (sub - 1) & mask
            best[mask] = bestc
    # unroll into a linear path
    order = []

    def rec(mask):
        split = best[mask][1]
        if split is None:
            return
        rec(split[0])
        rec(split[1])
        order.append(split)

    rec(full)
    current = [1 << k for k in range(n)]
    path = []
    for a, b in order:
        i, j = current.index(a), current.index(b)
        path.append((min(i, j), max(i, j)))
        current = [m for k, m in enumerate(current) if k not in (i, j)] + [a | b]
    return path


def path_cost(inputs, output, size_dict, path):
    """(total flops as sum of 2*M*N*K per step, largest intermediate size)."""
    terms = [tuple(t) for t in inputs]
    flops, width = 0, 0
    for i, j in path:
        need = _needed_elsewhere(terms, (i, j), output)
        res = _pair_result(terms[i], terms[j], need)
        flops += 2 * _flops(terms[i], terms[j], size_dict)
        width = max(width, _size(res, size_dict))
        terms = [t for k, t in enumerate(terms) if k not in (i, j)] + [res]
    return flops, width


# ------------------------------------------------------------ executor -----
def contract_pair(a, ia, b, ib, iout):
    """One tree node: tensordot (+ transpose) if the pair is a pure
    tensordot, einsum otherwise -- what cotengra emits through
    autoray.do(..., like='numpy')."""
    ia, ib, iout = tuple(ia), tuple(ib), tuple(iout)
    sa, sb, so = set(ia), set(ib), set(iout)
    pure = (
        len(sa) == len(ia)
        and len(sb) == len(ib)
        and not (sa & sb & so)          # no batch index
        and (sa - sb) <= so             # nothing summed out of a alone
        and (sb - sa) <= so
    )
    if pure:
        shared = [ix for ix in ia if ix in sb]
        axa = [ia.index(ix) for ix in shared]
        axb = [ib.index(ix) for ix in shared]
        res = np.tensordot(a, b, axes=(axa, axb))
        ires = tuple(ix for ix in ia if ix not in sb) + tuple(
            ix for ix in ib if ix not in sa
        )
        if ires != iout:
            res = np.transpose(res, [ires.index(ix) for ix in iout])
        return res
    symbols = {}
    for ix in ia + ib + iout:
        symbols.setdefault(ix, chr(ord("a") + len(symbols)) if len(symbols) < 26
                           else chr(ord("A") + len(symbols) - 26))
    eq = "{},{}->{}".format(
        "".join(symbols[i] for i in ia),
        "".join(symbols[i] for i in ib),
        "".join(symbols[i] for i in iout),
    )
    return np.einsum(eq, a, b)


def array_contract(arrays, inputs, output, optimize="auto", size_dict=None,
                   strip_exponent=False):
    """Contract ``arrays`` labelled by ``inputs`` into ``output`` order.

    ``strip_exponent`` (tensor_core.py:330-336; executed by cotengra, absent
    offline -- its published behaviour restated): every intermediate is
    divided by its largest magnitude, the log10 of the factors accumulated;
    returns ``(mantissa, exponent)`` with result = mantissa * 10**exponent."""
    if strip_exponent:
        expo = [0.0]

        def strip(x):
            x = np.asarray(x)
            f = np.max(np.abs(x)) if x.size else 1.0
            expo[0] = expo[0] + np.log10(f)
            return x / f
    else:
        def strip(x):
            return x
    res = _array_contract(arrays, inputs, output, optimize, size_dict, strip)
    if strip_exponent:
        return res, float(expo[0])
    return res


def _array_contract(arrays, inputs, output, optimize, size_dict, strip):
    arrays = list(arrays)
    terms = [tuple(t) for t in inputs]
    output = tuple(output)
    if size_dict is None:
        size_dict = {}
        for t, x in zip(terms, arrays):
            for ix, d in zip(t, np.shape(x)):
                size_dict[ix] = int(d)
    if len(arrays) == 1:
        (x,), (t,) = arrays, terms
        if t == output:
            return strip(x)
        if len(set(t)) == len(t) and set(t) == set(output):
            return strip(np.transpose(x, [t.index(ix) for ix in output]))
        eq_in = "".join(chr(97 + list(dict.fromkeys(t)).index(i)) for i in t)
        eq_out = "".join(chr(97 + list(dict.fromkeys(t)).index(i)) for i in output)
        return strip(np.einsum(f"{eq_in}->{eq_out}", x))
    path = find_path(terms, output, size_dict, optimize)
    for i, j in path:
        need = _needed_elsewhere(terms, (i, j), output)
        last = len(terms) == 2
        res_inds = output if last else _pair_result(terms[i], terms[j], need)
        res = strip(contract_pair(arrays[i], terms[i], arrays[j], terms[j], res_inds))
        arrays = [x for k, x in enumerate(arrays) if k not in (i, j)] + [res]
        terms = [t for k, t in enumerate(terms) if k not in (i, j)] + [res_inds]
    return arrays[0]


def tensor_contract(arrays, inds, output_inds=None, optimize="auto",
                    strip_exponent=False, exponent=None):
    """(data, inds_out) of quimb.tensor.tensor_contract for raw arrays; with
    ``strip_exponent``: ((mantissa, exponent), inds_out); a supplied base
    ``exponent`` is added to the stripped one, or scales the plain result by
    10**exponent (tensor_core.py:330-341)."""
    if output_inds is None:
        inds_out = gen_output_inds(itertools.chain.from_iterable(inds))
    else:
        inds_out = tuple(output_inds)
    out = array_contract(arrays, inds, inds_out, optimize=optimize,
                         strip_exponent=strip_exponent)
    if strip_exponent:
        data, e = out
        if exponent is not None:
            e = e + exponent
        return (data, e), inds_out
    if exponent is not None:
        out = out * 10**exponent
    return out, inds_out
