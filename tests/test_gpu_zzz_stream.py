"""GPU tier: the streaming contraction engine (csrc/contract_stream.cu,
``engine=QB_ENGINE_STREAM``) against the numpy oracle -- gate application on
arbitrary axes of large strided / conjugated tensors, batch labels,
accumulation, pointer-array batches.  Same oracle and tolerance (1e-11
relative, fp64) as the DMMA engine's parity tests; the index bookkeeping is
additionally checked without a device in tests/test_stream_engine_cpu.py.
Sorted last: the engine is opt-in and was written before it could be run."""

import numpy as np
import pytest
import torch

import quimb_b200 as qb
from quimb_b200 import _lib
from quimb_b200.contract import contract_pair

pytestmark = pytest.mark.gpu

STREAM = _lib.QB_ENGINE_STREAM


def _rand(rng, shape, cplx):
    x = rng.standard_normal(shape)
    if cplx:
        x = x + 1j * rng.standard_normal(shape)
    return x


def _einsum(a, la, b, lb, lc):
    sym = {}
    for l in list(la) + list(lb) + list(lc):
        sym.setdefault(l, chr(ord("a") + len(sym)))
    return np.einsum("{},{}->{}".format("".join(sym[l] for l in la), "".join(sym[l] for l in lb),
                                        "".join(sym[l] for l in lc)), a, b)


@pytest.mark.parametrize("cplx", [False, True])
def test_stream_engine_gate_application(cplx):
    rng = np.random.default_rng(3 + cplx)
    for nq, where in [(10, (3,)), (14, (0, 13)), (18, (5, 6)), (20, (19, 2)), (12, (11,))]:
        state = _rand(rng, (2,) * nq, cplx)
        ng = len(where)
        gate = _rand(rng, (2,) * (2 * ng), cplx)
        la = list(range(nq))
        new = [100 + i for i in range(ng)]
        lb = new + list(where)
        lc = [new[where.index(q)] if q in where else q for q in range(nq)]
        ref = _einsum(state, la, gate, lb, lc)
        A, B = qb.asarray(state), qb.asarray(gate)
        n0 = qb.launch_count()
        out = contract_pair(A.t, la, B.t, lb, lc, engine=STREAM)
        assert qb.launch_count() - n0 == 1                 # one kernel, no staging copies
        np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-11, atol=1e-11)
        # the default engine agrees (bit-for-bit is not required: different summation order)
        out_d = contract_pair(A.t, la, B.t, lb, lc)
        np.testing.assert_allclose(out.cpu().numpy(), out_d.cpu().numpy(), rtol=1e-12, atol=1e-12)
        # conjugated operands are load flags
        if cplx:
            out_c = contract_pair(A.t, la, B.t, lb, lc, conj_a=True, conj_b=True, engine=STREAM)
            np.testing.assert_allclose(out_c.cpu().numpy(), ref.conj(), rtol=1e-11, atol=1e-11)


def test_stream_engine_strided_batch_accumulate_and_fallback():
    rng = np.random.default_rng(11)
    # permuted (non-contiguous) state view, output written into a strided view
    x = _rand(rng, (8, 6, 4, 16, 3), True)
    A = qb.asarray(x).t.permute(3, 0, 4, 1, 2)              # labels (3, 0, 4, 1, 2)
    la = [3, 0, 4, 1, 2]
    g = _rand(rng, (5, 6), True)
    B = qb.asarray(g).t
    lb, lc = [9, 1], [0, 9, 2, 3, 4]
    ref = np.einsum("abcde,fb->afcde", x, g)
    out = contract_pair(A, la, B, lb, lc, engine=STREAM)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-11, atol=1e-11)
    store = torch.zeros((8, 5, 4, 16, 6), dtype=torch.complex128, device=A.device)
    view = store[..., :3]
    contract_pair(A, la, B, lb, lc, out=view, engine=STREAM)
    np.testing.assert_allclose(view.cpu().numpy(), ref, rtol=1e-11, atol=1e-11)
    assert float(store[..., 3:].abs().max()) == 0.0
    # batch label + accumulate
    a, b = _rand(rng, (7, 300, 4), False), _rand(rng, (7, 4, 3), False)
    prev = _rand(rng, (7, 300, 3), False)
    o = qb.asarray(prev.copy()).t
    contract_pair(qb.asarray(a).t, [0, 1, 2], qb.asarray(b).t, [0, 2, 3], [0, 1, 3], out=o,
                  engine=STREAM, alpha=0.25, beta=-1.0)
    np.testing.assert_allclose(o.cpu().numpy(), 0.25 * np.einsum("bmk,bkn->bmn", a, b) - prev,
                               rtol=1e-11, atol=1e-11)
    # a shape the engine does not take goes to the DMMA path under the same flag
    a, b = _rand(rng, (64, 80), False), _rand(rng, (80, 48), False)
    o = contract_pair(qb.asarray(a).t, [0, 1], qb.asarray(b).t, [1, 2], [0, 2], engine=STREAM)
    np.testing.assert_allclose(o.cpu().numpy(), a @ b, rtol=1e-11, atol=1e-11)
    # large M: every row of a 2^22-row operand
    s = _rand(rng, (2,) * 22, False)
    gate = _rand(rng, (2, 2, 2, 2), False)
    la = list(range(22))
    out = contract_pair(qb.asarray(s).t, la, qb.asarray(gate).t, [50, 51, 4, 17],
                        [50 if q == 4 else 51 if q == 17 else q for q in la], engine=STREAM)
    ref = np.einsum(gate, [50, 51, 4, 17], s, la, [50 if q == 4 else 51 if q == 17 else q for q in la])
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-11, atol=1e-11)
