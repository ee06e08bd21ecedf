"""CPU tier: the index bookkeeping of the Jacobi SVD (csrc/svd_jacobi.cu) that
needs no device -- the sweep schedule (every column-block pair exactly once per
sweep, disjoint blocks inside a round, independent groups inside a phase) and
a numpy emulation of the position-layout eigen-solve (slot movement, shuffles,
ping-pong writes) as the kernel's 256 threads perform it."""

import ctypes
import itertools

import numpy as np
import pytest

from quimb_b200 import _lib


def _schedule(nblk, groups):
    lib = _lib.load()
    cap = nblk * nblk
    buf = (ctypes.c_int32 * (5 * cap))()
    used = ctypes.c_int(0)
    n = lib.qb_debug_jacobi_schedule(nblk, groups, buf, cap, ctypes.byref(used))
    assert n >= 0
    recs = np.frombuffer(buf, dtype=np.int32, count=5 * n).reshape(n, 5)
    return recs, used.value


@pytest.mark.parametrize("nblk,groups", [(128, 4), (128, 2), (128, 1), (64, 4), (32, 2), (16, 4),
                                         (8, 2), (6, 4), (2, 1), (24, 4), (40, 4)])
def test_sweep_schedule_visits_every_pair_once_and_streams_are_independent(nblk, groups):
    recs, used = _schedule(nblk, groups)
    pairs = [(int(p), int(q)) for _, _, _, p, q in recs]
    assert all(0 <= p < q < nblk for p, q in pairs)
    assert sorted(pairs) == sorted(itertools.combinations(range(nblk), 2))   # exactly once
    assert used in (1, 2, 4) and used <= max(groups, 1)
    by_phase = {}
    for ph, g, r, p, q in recs:
        by_phase.setdefault(int(ph), {}).setdefault(int(g), {}).setdefault(int(r), []).append((int(p), int(q)))
    for ph, groups_ in by_phase.items():
        touched = {}
        for g, rounds in groups_.items():
            blocks = set()
            for r, prs in rounds.items():
                flat = [b for pr in prs for b in pr]
                assert len(flat) == len(set(flat)), "a block twice in one round"
                blocks.update(flat)
            touched[g] = blocks
        for g1, g2 in itertools.combinations(touched, 2):
            assert not (touched[g1] & touched[g2]), "concurrent streams share a column block"


def test_position_layout_eigensolve_emulation():
    """The Brent-Luk position layout of the in-kernel 32 x 32 eigen-solve,
    thread for thread: after 31 steps the players are back in their starting
    slots, J is orthogonal, J^T G J has the diagonal the threads hold, and a few
    sweeps diagonalise a Gram matrix."""
    JP, JB = 32, 16

    def slot_player(isB, k):
        if not isB:
            return JP - 1 if k == 0 else k
        return 0 if k == 0 else JP - 1 - k

    def slot_next(isB, k):
        if not isB:
            return (0, 0) if k == 0 else ((1, 0) if k == 1 else (0, k - 1))
        return (0, JB - 1) if k == JB - 1 else (1, k + 1)

    rng = np.random.default_rng(0)
    X = rng.standard_normal((200, JP))
    G0 = X.T @ X
    k1, k2 = np.meshgrid(np.arange(16), np.arange(16), indexing="ij")
    sp = np.vectorize(slot_player)
    rA, rB, cA, cB = sp(0, k1), sp(1, k1), sp(0, k2), sp(1, k2)
    gAA, gAB, gBA, gBB = G0[rA, cA], G0[rA, cB], G0[rB, cA], G0[rB, cB]
    jA, jB = np.zeros((2, 16, 16)), np.zeros((2, 16, 16))
    for it in range(2):
        jA[it], jB[it] = (k1 + 16 * it == cA), (k1 + 16 * it == cB)
    GA, GB = np.zeros((32, 16)), np.zeros((32, 16))
    GA[k1, k2], GB[k1, k2], GA[16 + k1, k2], GB[16 + k1, k2] = gAA, gAB, gBA, gBB
    nxt = [[slot_next(0, k) for k in range(16)], [slot_next(1, k) for k in range(16)]]
    offs = []
    for sweep in range(7):
        for step in range(JP - 1):
            c, s = np.ones(16), np.zeros(16)
            for k in range(16):
                app, aqq, apq = GA[k, k], GB[16 + k, k], GB[k, k]
                if abs(apq) > 1e-300:
                    a, b = aqq - app, 2 * apq
                    r = 1 / np.sqrt(a * a + b * b)
                    c2t = abs(a) * r
                    s2t = (-abs(b) if (a < 0) != (b < 0) else abs(b)) * r
                    u = 0.5 * c2t + 0.5
                    ru = 1 / np.sqrt(u)
                    c[k], s[k] = u * ru, 0.5 * s2t * ru
            c2, s2, c1, s1 = c[k2], s[k2], c[k1], s[k1]
            a0, a1 = c2 * gAA - s2 * gAB, s2 * gAA + c2 * gAB
            b0, b1 = c2 * gBA - s2 * gBB, s2 * gBA + c2 * gBB
            n00, n01 = c1 * a0 - s1 * b0, c1 * a1 - s1 * b1
            n10, n11 = s1 * a0 + c1 * b0, s1 * a1 + c1 * b1
            n01, n10 = np.where(k1 == k2, 0, n01), np.where(k1 == k2, 0, n10)
            GAn, GBn = np.full((32, 16), np.nan), np.full((32, 16), np.nan)
            for i in range(16):
                (rAB, rAk), (rBB, rBk) = nxt[0][i], nxt[1][i]
                rowA, rowB = rAk + 16 * rAB, rBk + 16 * rBB
                for j in range(16):
                    (cAB, cAk), (cBB, cBk) = nxt[0][j], nxt[1][j]
                    (GBn if cAB else GAn)[rowA, cAk] = n00[i, j]
                    (GBn if cBB else GAn)[rowA, cBk] = n01[i, j]
                    (GBn if cAB else GAn)[rowB, cAk] = n10[i, j]
                    (GBn if cBB else GAn)[rowB, cBk] = n11[i, j]
            assert not np.isnan(GAn).any() and not np.isnan(GBn).any()   # every slot written once
            GA, GB = GAn, GBn
            for it in range(2):
                ja, jb = c2 * jA[it] - s2 * jB[it], s2 * jA[it] + c2 * jB[it]
                dn = np.roll(ja, -1, axis=1); dn[:, 15] = ja[:, 15]      # shfl_down, width 16
                up = np.roll(jb, 1, axis=1); up[:, 0] = jb[:, 0]         # shfl_up, width 16
                jA[it] = np.where(k2 == 0, ja, np.where(k2 == 15, jb, dn))
                jB[it] = np.where(k2 == 0, dn, up)
            gAA, gAB, gBA, gBB = GA[k1, k2], GB[k1, k2], GA[16 + k1, k2], GB[16 + k1, k2]
        J = np.zeros((JP, JP))
        for it in range(2):
            J[k1 + 16 * it, cA], J[k1 + 16 * it, cB] = jA[it], jB[it]
        D = J.T @ G0 @ J
        assert np.abs(J.T @ J - np.eye(JP)).max() < 1e-13
        dg = np.zeros(JP)
        for k in range(16):
            dg[slot_player(0, k)], dg[slot_player(1, k)] = gAA[k, k], gBB[k, k]
        assert np.abs(dg - np.diag(D)).max() < 1e-11 * np.abs(D).max()
        offs.append(np.abs(D - np.diag(np.diag(D))).max() / np.abs(np.diag(D)).max())
    assert offs[-1] < 1e-12 and all(b <= a * 1.0001 for a, b in zip(offs, offs[1:]))
