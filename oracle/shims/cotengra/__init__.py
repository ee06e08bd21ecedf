"""TEST INFRASTRUCTURE ONLY -- minimal functional stand-in for `cotengra`.

cotengra 0.8.2 (pixi.lock:1022 of the reference) owns the contraction-tree
executor of quimb's hot path but cannot be installed offline.  This shim
provides the entry points quimb calls (quimb/tensor/contraction.py:285-313)
on top of the numpy oracle in oracle/contract_np.py, so that the unmodified
reference can be run here to generate golden vectors.  Never imported by the
product package.
"""
import math
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from oracle import contract_np as _cn  # noqa: E402

from . import utils  # noqa: E402,F401
from . import cotengra  # noqa: E402,F401

__version__ = "0.0.shim"


def get_symbol(i):
    if i < 26:
        return "abcdefghijklmnopqrstuvwxyz"[i]
    if i < 52:
        return "ABCDEFGHIJKLMNOPQRSTUVWXYZ"[i - 26]
    return chr(i + 140)


def get_symbol_map(inputs):
    symbol_map = {}
    c = 0
    for term in inputs:
        for ind in term:
            if ind not in symbol_map:
                symbol_map[ind] = get_symbol(c)
                c += 1
    return symbol_map


class PathOptimizer:
    pass


class ContractionTree:
    """Just enough of a tree: a linear path plus cost queries."""

    def __init__(self, inputs, output, size_dict, path):
        self.inputs = tuple(tuple(t) for t in inputs)
        self.output = tuple(output)
        self.size_dict = dict(size_dict)
        self._path = list(path)
        self.sliced_inds = {}
        self.multiplicity = 1

    def get_path(self):
        return tuple(self._path)

    def contraction_cost(self, log=None):
        fl, _ = _cn.path_cost(self.inputs, self.output, self.size_dict, self._path)
        c = fl // 2
        return math.log(c, log) if log else c

    def contraction_width(self, log=2):
        _, w = _cn.path_cost(self.inputs, self.output, self.size_dict, self._path)
        return math.log(max(w, 1), log) if log else w

    def max_size(self, log=None):
        _, w = _cn.path_cost(self.inputs, self.output, self.size_dict, self._path)
        return w

    def contract(self, arrays, backend=None, **kw):
        return _cn.array_contract(arrays, self.inputs, self.output,
                                  optimize=self._path, size_dict=self.size_dict)


class ContractionTreeCompressed(ContractionTree):
    pass


def _sizes(inputs, shapes):
    size_dict = {}
    for t, s in zip(inputs, shapes):
        for ix, d in zip(t, s):
            size_dict[ix] = int(d)
    return size_dict


def _resolve(optimize, inputs, output, size_dict):
    if isinstance(optimize, ContractionTree):
        return optimize.get_path()
    if optimize is None:
        optimize = "auto"
    if isinstance(optimize, str):
        return _cn.find_path(inputs, output, size_dict, optimize)
    return [tuple(p) for p in optimize]


def array_contract_tree(inputs, output=None, size_dict=None, shapes=None,
                        optimize="auto", **kwargs):
    inputs = tuple(tuple(t) for t in inputs)
    if size_dict is None:
        size_dict = _sizes(inputs, shapes)
    if output is None:
        output = _cn.gen_output_inds(ix for t in inputs for ix in t)
    path = _resolve(optimize, inputs, output, size_dict)
    return ContractionTree(inputs, output, size_dict, path)


def array_contract_path(inputs, output=None, size_dict=None, shapes=None,
                        optimize="auto", **kwargs):
    return array_contract_tree(inputs, output, size_dict, shapes, optimize).get_path()


def _strip(x):
    xmax = np.max(np.abs(x))
    if xmax == 0.0:
        return x, 0.0
    return x / xmax, float(np.log10(xmax))


def array_contract(arrays, inputs, output=None, optimize="auto", backend=None,
                   strip_exponent=False, cache_expression=True, **kwargs):
    inputs = tuple(tuple(t) for t in inputs)
    shapes = [np.shape(a) for a in arrays]
    size_dict = _sizes(inputs, shapes)
    if output is None:
        output = _cn.gen_output_inds(ix for t in inputs for ix in t)
    path = _resolve(optimize, inputs, tuple(output), size_dict)
    return _cn.array_contract(arrays, inputs, tuple(output), optimize=path,
                              size_dict=size_dict, strip_exponent=strip_exponent)


def array_contract_expression(inputs, output=None, size_dict=None, shapes=None,
                              optimize="auto", constants=None, **kwargs):
    inputs = tuple(tuple(t) for t in inputs)
    if size_dict is None:
        size_dict = _sizes(inputs, shapes)
    if output is None:
        output = _cn.gen_output_inds(ix for t in inputs for ix in t)
    output = tuple(output)
    path = _resolve(optimize, inputs, output, size_dict)
    constants = dict(constants or {})
    var_pos = [i for i in range(len(inputs)) if i not in constants]

    def expr(*arrays, backend=None, **kw):
        full = [None] * len(inputs)
        for i, c in constants.items():
            full[i] = c
        for i, a in zip(var_pos, arrays):
            full[i] = a
        return _cn.array_contract(full, inputs, output, optimize=path,
                                  size_dict=size_dict)

    expr.path = path
    return expr


def get_hypergraph(*a, **k):
    raise NotImplementedError("cotengra shim: hypergraph tools unavailable")


HyperGraph = None
