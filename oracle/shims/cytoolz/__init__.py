"""TEST INFRASTRUCTURE ONLY -- the 15 cytoolz functions quimb/utils.py:9-26
imports, restated from the toolz documentation (pure python)."""
import collections
import functools
import itertools

__version__ = "0.0.shim"


def last(seq):
    return collections.deque(seq, maxlen=1)[0] if not hasattr(seq, "__getitem__") else seq[-1]


def concat(seqs):
    return itertools.chain.from_iterable(seqs)


def concatv(*seqs):
    return itertools.chain.from_iterable(seqs)


def frequencies(seq):
    d = {}
    for x in seq:
        d[x] = d.get(x, 0) + 1
    return d


def partition_all(n, seq):
    it = iter(seq)
    while True:
        chunk = tuple(itertools.islice(it, n))
        if not chunk:
            return
        yield chunk


def partition(n, seq):
    it = iter(seq)
    while True:
        chunk = tuple(itertools.islice(it, n))
        if len(chunk) < n:
            return
        yield chunk


def partitionby(func, seq):
    return (tuple(v) for _, v in itertools.groupby(seq, key=func))


def merge_with(func, *dicts):
    if len(dicts) == 1 and not isinstance(dicts[0], dict):
        dicts = dicts[0]
    out = collections.defaultdict(list)
    for d in dicts:
        for k, v in d.items():
            out[k].append(v)
    return {k: func(v) for k, v in out.items()}


def valmap(func, d):
    return {k: func(v) for k, v in d.items()}


def keymap(func, d):
    return {func(k): v for k, v in d.items()}


def identity(x):
    return x


def compose(*funcs):
    if not funcs:
        return identity

    def composed(*args, **kwargs):
        out = funcs[-1](*args, **kwargs)
        for f in reversed(funcs[:-1]):
            out = f(out)
        return out

    return composed


def isiterable(x):
    try:
        iter(x)
        return True
    except TypeError:
        return False


def unique(seq, key=None):
    seen = set()
    for x in seq:
        k = x if key is None else key(x)
        if k not in seen:
            seen.add(k)
            yield x
