"""ORACLE (test infrastructure, never imported by the product package).

Plain-numpy restatement of the MPS gate / TEBD half of the reference:

  gate_split (eager contract + split)  quimb/tensor/tn1d/core.py:2219-2247,
                                       quimb/tensor/gating.py:86-123
  left/right_canonize_site             tn1d/core.py:824-905
  swap_sites_with_compress/swap_site_to tn1d/core.py:1628-1735
  gate_with_auto_swap                  tn1d/core.py:2251-2322
  LocalHam1D / get_gate_expm           tn1d/tebd.py:12-96, tnag/tebd.py:244-392
  trotter_schedule                     tnag/tebd.py:78-126
  TEBD.sweep / step / update_to        tn1d/tebd.py:323-506

Sites are (l, p, r) numpy arrays.  Pinned against the reference itself by
oracle/make_golden.py:tebd_cases (tests/golden/tebd.*).
"""

import numpy as np
import scipy.linalg as sla

from . import decomp_np as dn


def left_canonize_site(sites, i):
    A = sites[i]
    l, d, r = A.shape
    Q, _, R = dn.qr_stabilized(A.reshape(l * d, r).copy(), absorb=dn.get_U_sVH)
    sites[i] = Q.reshape(l, d, -1)
    sites[i + 1] = np.tensordot(R, sites[i + 1], axes=(1, 0))


def right_canonize_site(sites, i):
    A = sites[i]
    l, d, r = A.shape
    Lf, _, Q = dn.qr_stabilized(A.reshape(l, d * r).copy(), absorb=dn.get_Us_VH)
    sites[i] = Q.reshape(-1, d, r)
    sites[i - 1] = np.tensordot(sites[i - 1], Lf, axes=(2, 0))


def canonicalize(sites, where):
    imin, imax = (where, where) if isinstance(where, int) else (min(where), max(where))
    for i in range(0, imin):
        left_canonize_site(sites, i)
    for i in range(len(sites) - 1, imax, -1):
        right_canonize_site(sites, i)


def gate_split(sites, G, where, **opts):
    opts.setdefault("cutoff_mode", "rsum2")
    i, j = where
    a, b = min(i, j), max(i, j)
    A, B = sites[a], sites[b]
    d1, d2 = A.shape[1], B.shape[1]
    G = np.asarray(G)
    if G.ndim == 2:
        G = G.reshape((d1, d2, d1, d2) if i < j else (d2, d1, d2, d1))
    if i > j:
        G = G.transpose(1, 0, 3, 2)
    T = np.einsum("lpm,mqr,PQpq->lPQr", A, B, G)
    left, s, right = dn.tensor_split(T, "lpqr", "lp", "qr", **opts)
    sites[a], sites[b] = left, right
    return s


def swap_sites_with_compress(sites, i, j, **opts):
    opts.setdefault("cutoff_mode", "rsum2")
    T = np.einsum("lpm,mqr->lqpr", sites[i], sites[j])
    left, _, right = dn.tensor_split(T, "lqpr", "lq", "pr", **opts)
    sites[i], sites[j] = left, right


def gate_with_auto_swap(sites, G, where, swap_back=True, **opts):
    i, j = where
    if i > j:
        i, j = j, i
        final_where, absorb = (i + 1, i), "left"
    else:
        final_where, absorb = (i, i + 1), "right"
    need = i + 1 != j
    if need:
        for k in range(j, i + 1, -1):
            canonicalize(sites, (k - 1, k))
            swap_sites_with_compress(sites, k - 1, k, absorb="left", **opts)
    canonicalize(sites, (i, i + 1))
    gate_split(sites, G, final_where, absorb=absorb, **opts)
    if need and swap_back:
        for k in range(i + 1, j):
            canonicalize(sites, (k, k + 1))
            swap_sites_with_compress(sites, k, k + 1, absorb="right", **opts)


def trotter_schedule(nlayers, order=2):
    if order == 1:
        return [(k, 1.0) for k in range(nlayers)]
    if order == 2:
        return ([(k, 0.5) for k in range(nlayers - 1)] + [(nlayers - 1, 1.0)]
                + [(k, 0.5) for k in reversed(range(nlayers - 1))])
    s = 1 / (4 - 4 ** (1 / 3))
    return [(k, frac * f) for f in (s, s, 1 - 4 * s, s, s)
            for k, frac in trotter_schedule(nlayers, 2)]


class TEBD:
    """tn1d/tebd.py:221-506 for an open chain with two-site ``terms``
    {(i, i + 1): d^2 x d^2 array}."""

    def __init__(self, p0, terms, dt=None, tol=None, split_opts=None, imag=False):
        self.sites = [np.asarray(a, dtype=np.result_type(a.dtype, np.float64 if imag
                                                         else np.complex128)) for a in p0]
        self.L = len(self.sites)
        self.terms = dict(terms)
        canonicalize(self.sites, 0)
        self.ham_norm = sum(np.linalg.norm(h) for h in terms.values()) / len(terms)
        self.t, self.err = 0.0, 0.0
        self.dt = self._dt = dt
        self.tol = tol
        self.imag = imag
        self.split_opts = dict(split_opts or {})
        self._queued = None
        self._cache = {}

    def gate(self, dt_frac, where):
        x = -(1.0 if self.imag else 1.0j) * self._dt * dt_frac
        key = (where, complex(x))
        if key not in self._cache:
            self._cache[key] = sla.expm(self.terms[where] * x)
        return self._cache[key]

    def sweep(self, direction, dt_frac, dt=None, queue=False):
        if dt is not None:
            dt_frac *= dt / self._dt
        if queue:
            if self._queued:
                if direction == self._queued[0]:
                    self._queued[1] += dt_frac
                    return
                new = [direction, dt_frac]
                direction, dt_frac = self._queued
                self._queued = new
            else:
                self._queued = [direction, dt_frac]
                return
        elif self._queued:
            qd, qf = self._queued
            self._queued = None
            self.sweep(qd, qf, queue=False)
        s = self.sites
        if direction == "right":
            final = self.L - 1
            for i in range(0, final, 2):
                for k in range(max(0, i - 1), i):
                    left_canonize_site(s, k)
                gate_split(s, self.gate(dt_frac, (i, i + 1)), (i, i + 1), absorb="right",
                           **self.split_opts)
            if self.L % 2 == 1:
                left_canonize_site(s, self.L - 2)
        else:
            final = 1
            for i in reversed(range(final, self.L - 1, 2)):
                for k in range(min(self.L - 1, i + 2), i + 1, -1):
                    right_canonize_site(s, k)
                gate_split(s, self.gate(dt_frac, (i, i + 1)), (i, i + 1), absorb="left",
                           **self.split_opts)
            right_canonize_site(s, 1)
        if self.imag:
            s[final] = s[final] / np.linalg.norm(s[final])

    def step(self, order=2, dt=None, **kw):
        for k, frac in trotter_schedule(2, order):
            self.sweep(("right", "left")[k], frac, dt=dt, **kw)
        dt = self._dt if dt is None else dt
        self.t += dt
        self.err += self.ham_norm * dt ** (order + 1)

    def update_to(self, T, dt=None, tol=None, order=4):
        dt = self.dt if dt is None else dt
        tol = self.tol if tol is None else tol
        self._dt = (tol / ((T - self.t) * self.ham_norm)) ** (1 / order) if dt is None else dt
        while self.t < T - self._dt:
            self.step(order=order, dt=None, queue=True)
        self.step(order=order, dt=T - self.t, queue=False)
