"""CPU tier: the index bookkeeping of the streaming contraction engine
(csrc/contract_stream.cu) checked WITHOUT a device.  The engine's row
arithmetic, operator-table fill and batch addressing are __host__ __device__
functions; ``qb_debug_contract_stream_host`` (a TEST entry of the C ABI, never
called by the product) runs exactly those functions in a host loop over host
buffers.  Random labelled contractions with small N and K -- gate application
on arbitrary axes of strided / permuted / conjugated operands, batch and
summed labels, alpha / beta accumulation -- are compared with numpy einsum
(the operation cotengra's pairwise loop would issue through
``do("tensordot")`` / ``do("einsum")``, quimb/tensor/tensor_core.py:3786-3808).
What this cannot check is the launch itself; the ``-m gpu`` test in
tests/test_gpu_zzz_stream.py does that against the same oracle."""

import ctypes

import numpy as np
import pytest

from quimb_b200 import _lib


def _desc(x):
    code = {np.dtype("float64"): _lib.QB_F64, np.dtype("complex128"): _lib.QB_C128}[x.dtype]
    return _lib.np_desc(x.shape, [s // x.itemsize for s in x.strides], code, ptr=x.ctypes.data)


def _stream_host(a, la, b, lb, out, lc, conj_a=False, conj_b=False, alpha=1.0, beta=0.0):
    lib = _lib.load()
    rc = lib.qb_debug_contract_stream_host(_desc(a), _lib.labels(la), _desc(b), _lib.labels(lb),
                                           _desc(out), _lib.labels(lc), int(conj_a), int(conj_b),
                                           float(alpha), float(beta))
    return rc


def _einsum(a, la, b, lb, lc, conj_a, conj_b):
    sym = {}
    for l in list(la) + list(lb) + list(lc):
        sym.setdefault(l, chr(ord("a") + len(sym)))
    eq = "{},{}->{}".format("".join(sym[l] for l in la), "".join(sym[l] for l in lb),
                            "".join(sym[l] for l in lc))
    return np.einsum(eq, a.conj() if conj_a else a, b.conj() if conj_b else b)


def _rand(rng, shape, cplx):
    x = rng.standard_normal(shape)
    if cplx:
        x = x + 1j * rng.standard_normal(shape)
    return x


def _strided(rng, x):
    """same values behind a random axis permutation / padding of the storage"""
    perm = rng.permutation(x.ndim)
    big_shape = [x.shape[p] + int(rng.integers(0, 2)) for p in perm]
    store = np.zeros(big_shape, dtype=x.dtype)
    view = store[tuple(slice(0, x.shape[p]) for p in perm)]
    view[...] = np.transpose(x, perm)
    return np.transpose(view, np.argsort(perm))


@pytest.mark.parametrize("cplx", [False, True])
def test_gate_application_on_random_axes(cplx):
    rng = np.random.default_rng(7 + cplx)
    for trial in range(40):
        nq = int(rng.integers(3, 9))
        dims = [int(rng.choice([2, 2, 2, 3])) for _ in range(nq)]
        state = _strided(rng, _rand(rng, dims, cplx))
        ng = int(rng.integers(1, 3))                       # one- or two-qubit gate
        where = [int(q) for q in rng.choice(nq, size=ng, replace=False)]
        gdims = [dims[q] for q in where]
        gate = _strided(rng, _rand(rng, gdims + gdims, cplx))     # (out..., in...)
        la = list(range(nq))
        new = [100 + i for i in range(ng)]
        lb = new + where
        lc = [new[where.index(q)] if q in where else q for q in range(nq)]
        if rng.random() < 0.5:                              # consumer wants another axis order
            lc = [lc[p] for p in rng.permutation(nq)]
        conj_a, conj_b = bool(rng.integers(2)), bool(rng.integers(2))
        ref = _einsum(state, la, gate, lb, lc, conj_a and cplx, conj_b and cplx)
        out = _strided(rng, np.zeros(ref.shape, dtype=state.dtype))
        rc = _stream_host(state, la, gate, lb, out, lc, conj_a, conj_b)
        assert rc == 0, _lib.last_error()
        np.testing.assert_allclose(out, ref, rtol=1e-13, atol=1e-13)
        # accumulate form: out = alpha * contraction + beta * out
        prev = _rand(rng, ref.shape, cplx)
        out2 = prev.copy()
        rc = _stream_host(state, la, gate, lb, out2, lc, conj_a, conj_b, alpha=-0.5, beta=2.0)
        assert rc == 0
        np.testing.assert_allclose(out2, -0.5 * ref + 2.0 * prev, rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("cplx", [False, True])
def test_batch_summed_and_degenerate_labels(cplx):
    rng = np.random.default_rng(17 + cplx)
    cases = [
        # (shape a, labels a, shape b, labels b, labels c)
        ((5, 3, 4), [0, 1, 2], (3, 4, 2), [1, 2, 9], [0, 9]),          # K = 12, N = 2
        ((6, 2, 7), [0, 1, 2], (2, 3), [1, 9], [9, 2, 0]),              # output permuted
        ((4, 5, 3), [0, 1, 2], (4, 3, 2), [0, 2, 9], [0, 1, 9]),        # batch label 0
        ((4, 5, 3), [0, 1, 2], (3, 2), [2, 9], [0, 9]),                 # label 1 summed out of a
        ((6, 5), [0, 1], (4,), [9], [0, 1, 9]),                         # outer product (K = 1)
        ((6, 5), [0, 1], (5,), [1], [0]),                               # matrix-vector (N = 1)
        ((3, 4, 2, 2), [0, 1, 2, 3], (2, 2, 2, 2), [8, 9, 2, 3], [8, 0, 9, 1]),
        ((7,), [0], (7,), [0], []),                                     # full contraction, M = 1
        ((2, 2, 2, 2, 2, 2, 2, 2, 2, 2), list(range(10)), (2, 2, 2, 2), [20, 21, 3, 7],
         [0, 1, 2, 20, 4, 5, 6, 21, 8, 9]),
    ]
    for sa, la, sb, lb, lc in cases:
        a, b = _rand(rng, sa, cplx), _rand(rng, sb, cplx)
        ref = np.asarray(_einsum(a, la, b, lb, lc, False, False))
        out = np.zeros(ref.shape, dtype=a.dtype)
        rc = _stream_host(a, la, b, lb, out, lc)
        assert rc == 0, (_lib.last_error(), sa, sb)
        np.testing.assert_allclose(out, ref, rtol=1e-13, atol=1e-13)


def test_shapes_outside_the_engine_are_refused():
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((8, 40)), rng.standard_normal((40, 3))      # K = 40
    out = np.zeros((8, 3))
    assert _stream_host(a, [0, 1], b, [1, 2], out, [0, 2]) == -9
    a, b = rng.standard_normal((8, 3)), rng.standard_normal((3, 40))       # N = 40
    out = np.zeros((8, 40))
    assert _stream_host(a, [0, 1], b, [1, 2], out, [0, 2]) == -9
    # planner-level argument errors come back as they do from qb_contract_pair
    assert _stream_host(a, [0, 1], b, [1, 2], out, [0, 5]) < 0
