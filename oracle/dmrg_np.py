"""ORACLE (test infrastructure, never imported by the product package).

Plain-numpy/scipy restatement of the 1D part of quimb's hot path on the
numpy backend, open boundary conditions:

  MPS / MPO builders      quimb/tensor/tensor_builder.py:4166-4244 (MPS_rand_state,
                          arrays drawn N(0,1) and `sensibly_scale`d,
                          quimb/tensor/array_ops.py:277-281),
                          tensor_builder.py:4856-4947,5501 (MPO_ham_heis)
  norm / expectation      quimb/tensor/tn1d/core.py:55-95 (expec_TN_1D),
                          tensor_core.py:4879-4918 (norm) -- a left-to-right
                          transfer contraction, 4 d chi^3 flops per site
  canonisation            tn1d/core.py:824-990 via tensor_canonize_bond
                          (tensor_core.py:671-824): QR one site, absorb R
  DMRG2                   tn1d/dmrg.py: MovingEnvironment.init_segment :281-322,
                          move_right/left :383-425, _update_local_state_2site
                          :803-870 (Heff as a linear operator, v0 = old
                          two-site tensor, ARPACK eigsh k=1 ncv=4 tol=1e-3
                          through quimb/linalg/scipy_linalg.py:113-128,
                          SVD split with absorb=direction, cutoff_mode
                          'sum2'), sweep :885-998, solve :1032-1131

Internal array layouts (gauge / layout free choices of this restatement):
MPS site A[l, p, r] (end sites carry a size-1 bond), MPO site W[wl, wr, pu,
pd], environments E[a_bra, w, a_ket].  Parity with the reference is asserted
on layout/gauge-invariant quantities only: norms, energies, singular values,
bond dimensions, truncation errors.
"""

import numpy as np
import scipy.sparse.linalg as spla

from . import decomp_np as dn


# ------------------------------------------------------------------ builders
def sensibly_scale(x):
    """array_ops.py:277-281"""
    return x / np.linalg.norm(x) ** (1.5 / x.ndim)


def mps_rand(L, bond_dim, phys_dim=2, seed=0, dtype="float64"):
    """Random OBC MPS with bonds min(d^i, d^(L-i), bond_dim), unnormalised
    (MPS_rand_state(..., normalize=False) analogue; own RNG stream)."""
    rng = np.random.default_rng(seed)
    d = phys_dim
    bonds = [1]
    for i in range(1, L):
        # guard the power against overflow for long chains
        e = min(i, L - i)
        cap = d ** e if e < 60 else bond_dim
        bonds.append(int(min(cap, bond_dim)))
    bonds.append(1)
    sites = []
    for i in range(L):
        shape = (bonds[i], d, bonds[i + 1])
        x = rng.standard_normal(shape)
        if "complex" in str(dtype):
            x = (x + 1j * rng.standard_normal(shape)) / np.sqrt(2)
        # quimb scales the (l, r, p) array; squeezing size-1 bonds first
        nd = sum(1 for s in shape if s > 1) or 1
        x = x / np.linalg.norm(x) ** (1.5 / nd)
        sites.append(x.astype(dtype))
    return sites


def mpo_heis(L, j=1.0, bz=0.0, S=0.5, dtype="float64"):
    """Heisenberg chain MPO, bond dimension 5, H = sum_i j S_i.S_{i+1}
    (quimb's convention, tensor_builder.py:5501 with spin_ham_mpo_tensor
    :4856-4947: one row/column per two-site term plus identity/finish).
    W[wl, wr, pu, pd]; end sites keep a size-1 outer bond."""
    sx = np.array([[0, 0.5], [0.5, 0]])
    isy = np.array([[0, 0.5], [-0.5, 0]])  # i * S^y (real)
    sz = np.array([[0.5, 0], [0, -0.5]])
    eye = np.eye(2)
    W = np.zeros((5, 5, 2, 2))
    W[0, 0] = eye
    W[4, 4] = eye
    # S.S = Sx Sx - (iSy)(iSy) + Sz Sz
    W[4, 1] = sx
    W[1, 0] = j * sx
    W[4, 2] = isy
    W[2, 0] = -j * isy
    W[4, 3] = sz
    W[3, 0] = j * sz
    W[4, 0] = -bz * sz
    sites = []
    for i in range(L):
        if i == 0:
            sites.append(W[4:5].astype(dtype).copy())
        elif i == L - 1:
            sites.append(W[:, 0:1].astype(dtype).copy())
        else:
            sites.append(W.astype(dtype).copy())
    return sites


def mps_to_dense(sites):
    psi = sites[0]
    for A in sites[1:]:
        psi = np.tensordot(psi, A, axes=(psi.ndim - 1, 0))
    return psi.reshape(-1)


def mpo_to_dense(sites):
    op = sites[0]  # (wl, wr, pu, pd)
    for W in sites[1:]:
        op = np.einsum("abij,bckl->acikjl", op, W).reshape(
            op.shape[0], W.shape[1], op.shape[2] * W.shape[2],
            op.shape[3] * W.shape[3])
    return op[0, 0]


# ---------------------------------------------------------- norm / expec ----
def mps_norm2(sites, return_flops=False):
    """<psi|psi> by the left-to-right transfer contraction."""
    E = np.ones((1, 1), dtype=sites[0].dtype)
    flops = 0
    for A in sites:
        l, d, r = A.shape
        T = np.tensordot(E, A, axes=(1, 0))            # (l', d, r)
        E = np.tensordot(A.conj(), T, axes=((0, 1), (0, 1)))  # (r', r)
        flops += 2 * l * l * d * r + 2 * l * d * r * r
    out = E[0, 0]
    return (out, flops) if return_flops else out


def mps_expec(sites, mpo):
    """<psi|H|psi> (expec_TN_1D(bra, mpo, ket) analogue)."""
    E = np.ones((1, 1, 1), dtype=np.result_type(sites[0].dtype, mpo[0].dtype))
    for A, W in zip(sites, mpo):
        E = env_step_left(E, A, W)
    return E[0, 0, 0]


def env_step_left(E, A, W):
    """E'[b', w', b] = sum E[a', w, a] conj(A)[a', p', b'] W[w, w', p', p] A[a, p, b]"""
    T = np.tensordot(E, A, axes=(2, 0))                      # a' w p b
    T = np.tensordot(T, W, axes=((1, 2), (0, 3)))            # a' b w' p'
    out = np.tensordot(A.conj(), T, axes=((0, 1), (0, 3)))   # b' b w'
    return np.transpose(out, (0, 2, 1))                      # b' w' b


def env_step_right(E, A, W):
    """E'[a', w, a] = sum conj(A)[a', p', b'] W[w, w', p', p] A[a, p, b] E[b', w', b]"""
    T = np.tensordot(A, E, axes=(2, 2))                      # a p b' w'
    T = np.tensordot(W, T, axes=((1, 3), (3, 1)))            # w p' a b'
    return np.tensordot(A.conj(), T, axes=((1, 2), (1, 3)))  # a' w a


# ------------------------------------------------------------ canonisation --
def right_canonize(sites):
    """Sweep from the right: site i -> (R absorbed into i-1) Q^T, so that
    every site but the first is a right isometry (tn1d/core.py:937-990)."""
    sites = [s.copy() for s in sites]
    for i in range(len(sites) - 1, 0, -1):
        A = sites[i]
        l, d, r = A.shape
        # LQ of the (l) x (d r) matrix: A = L Q
        Lf, _, Q = dn.qr_stabilized(A.reshape(l, d * r), absorb=dn.get_Us_VH)
        k = Q.shape[0]
        sites[i] = Q.reshape(k, d, r)
        sites[i - 1] = np.tensordot(sites[i - 1], Lf, axes=(2, 0))
    return sites


def left_canonize(sites):
    sites = [s.copy() for s in sites]
    for i in range(len(sites) - 1):
        A = sites[i]
        l, d, r = A.shape
        Q, _, R = dn.qr_stabilized(A.reshape(l * d, r), absorb=dn.get_U_sVH)
        k = Q.shape[1]
        sites[i] = Q.reshape(l, d, k)
        sites[i + 1] = np.tensordot(R, sites[i + 1], axes=(1, 0))
    return sites


# ------------------------------------------------------------------ DMRG2 ---
class EffHam2(spla.LinearOperator):
    """Two-site effective Hamiltonian as a scipy LinearOperator (the role of
    TNLinearOperator, tensor_core.py:12297-12417); matvec contraction order
    L.x -> .W_i -> .W_{i+1} -> .R (the chi^3 order)."""

    def __init__(self, Lenv, W1, W2, Renv, dims):
        self.L, self.W1, self.W2, self.R = Lenv, W1, W2, Renv
        self.dims = dims  # (a, s, t, b)
        n = int(np.prod(dims))
        self.nmatvec = 0
        super().__init__(dtype=Lenv.dtype, shape=(n, n))

    def _matvec(self, v):
        self.nmatvec += 1
        x = v.reshape(self.dims)                                 # a s t b
        T = np.tensordot(self.L, x, axes=(2, 0))                 # a' w s t b
        T = np.tensordot(T, self.W1, axes=((1, 2), (0, 3)))      # a' t b w1 s'
        T = np.tensordot(T, self.W2, axes=((3, 1), (0, 3)))      # a' b s' w2 t'
        T = np.tensordot(T, self.R, axes=((1, 3), (2, 1)))       # a' s' t' b'
        return T.reshape(-1)


class DMRG2:
    """Restatement of quimb's DMRG2 driver for an OBC MPO (numpy backend)."""

    def __init__(self, mpo, bond_dims, cutoffs=1e-8, p0=None, seed=0):
        self.L = len(mpo)
        self.mpo = [np.asarray(w) for w in mpo]
        self.bond_dims = (bond_dims,) if isinstance(bond_dims, int) else tuple(bond_dims)
        self.cutoffs = (cutoffs,) if isinstance(cutoffs, float) else tuple(cutoffs)
        if p0 is None:
            p0 = mps_rand(self.L, self.bond_dims[0], seed=seed,
                          dtype=self.mpo[0].dtype)
        self.k = [np.asarray(a).copy() for a in p0]
        nrm = np.sqrt(abs(mps_norm2(self.k)))
        self.k[0] = self.k[0] / nrm
        self.energies, self.local_energies, self.total_energies = [], [], []
        self.nmatvecs = []
        self.opts = dict(local_eig_tol=1e-3, local_eig_ncv=4,
                         cutoff_mode="sum2", method="svd",
                         default_sweep_sequence="R")
        self._sweep_idx = 0

    # -- environments -------------------------------------------------------
    def _init_right_envs(self):
        dt = self.k[0].dtype
        self.renv = {self.L - 1: np.ones((1, 1, 1), dtype=dt)}
        for i in range(self.L - 1, 1, -1):
            self.renv[i - 1] = env_step_right(self.renv[i], self.k[i], self.mpo[i])

    def _init_left_envs(self):
        dt = self.k[0].dtype
        self.lenv = {0: np.ones((1, 1, 1), dtype=dt)}
        for i in range(0, self.L - 2):
            self.lenv[i + 1] = env_step_left(self.lenv[i], self.k[i], self.mpo[i])

    # -- local update ---------------------------------------------------------
    def _update_2site(self, i, direction, max_bond, cutoff):
        A, B = self.k[i], self.k[i + 1]
        Lenv, Renv = self.lenv[i], self.renv[i + 1]
        a, s, _ = A.shape
        _, t, b = B.shape
        dims = (a, s, t, b)
        Heff = EffHam2(Lenv, self.mpo[i], self.mpo[i + 1], Renv, dims)
        v0 = np.tensordot(A, B, axes=(2, 0)).reshape(-1)
        n = v0.size
        if n < 800:   # dense below this size, dmrg.py:690
            Hd = Heff @ np.eye(n)
            Hd = 0.5 * (Hd + Hd.conj().T)
            evals, evecs = np.linalg.eigh(Hd)
            loc_en, loc_gs = evals[0], evecs[:, 0]
        else:
            lk, vk = spla.eigsh(Heff, k=1, which="SA", v0=v0,
                                ncv=self.opts["local_eig_ncv"],
                                tol=self.opts["local_eig_tol"])
            loc_en, loc_gs = lk[0], vk[:, 0]
        self.nmatvecs.append(Heff.nmatvec)
        mat = loc_gs.reshape(a * s, t * b)
        absorb = dn.get_U_sVH if direction == "right" else dn.get_Us_VH
        opts = dn.parse_truncation_opts(max_bond, cutoff, self.opts["cutoff_mode"])
        info = {}
        left, _, right = dn.svd_truncated(mat, absorb=absorb, info=info, **opts)
        kdim = left.shape[1]
        self.k[i] = left.reshape(a, s, kdim)
        self.k[i + 1] = right.reshape(kdim, t, b)
        # total energy = full contraction of the local network (dmrg.py:868)
        x = np.tensordot(self.k[i], self.k[i + 1], axes=(2, 0))
        Hx = Heff._matvec(x.reshape(-1))
        tot_en = np.vdot(x.reshape(-1), Hx)
        return float(np.real(loc_en)), float(np.real(tot_en)), info

    def sweep(self, direction, canonize=True, max_bond=None, cutoff=0.0):
        if max_bond is None:
            max_bond = self.bond_dims[-1]
        L = self.L
        loc, tot = [], []
        self.last_infos = []
        if direction == "R":
            if canonize:
                self.k = right_canonize(self.k)
            self._init_right_envs()
            self.lenv = {0: np.ones((1, 1, 1), dtype=self.k[0].dtype)}
            for i in range(L - 1):
                if i > 0:
                    self.lenv[i] = env_step_left(self.lenv[i - 1], self.k[i - 1],
                                                 self.mpo[i - 1])
                le, te, info = self._update_2site(i, "right", max_bond, cutoff)
                loc.append(le); tot.append(te); self.last_infos.append(info)
        else:
            if canonize:
                self.k = left_canonize(self.k)
            self._init_left_envs()
            self.renv = {L - 1: np.ones((1, 1, 1), dtype=self.k[0].dtype)}
            for i in range(L - 2, -1, -1):
                if i < L - 2:
                    self.renv[i + 1] = env_step_right(self.renv[i + 2],
                                                      self.k[i + 2], self.mpo[i + 2])
                le, te, info = self._update_2site(i, "left", max_bond, cutoff)
                loc.append(le); tot.append(te); self.last_infos.append(info)
        self.local_energies.append(tuple(loc))
        self.total_energies.append(tuple(tot))
        return tot[-1]

    def solve(self, tol=1e-4, max_sweeps=10, sweep_sequence=None):
        seq = sweep_sequence or self.opts["default_sweep_sequence"]
        prev = "0"
        for n in range(max_sweeps):
            direction = seq[n % len(seq)]
            idx = self._sweep_idx
            max_bond = self.bond_dims[min(idx, len(self.bond_dims) - 1)]
            cutoff = self.cutoffs[min(idx, len(self.cutoffs) - 1)]
            self._sweep_idx += 1
            canonize = (direction + prev) not in ("LR", "RL")
            en = self.sweep(direction, canonize=canonize, max_bond=max_bond,
                            cutoff=cutoff)
            self.energies.append(en)
            if len(self.energies) >= 2 and abs(self.energies[-2] - self.energies[-1]) < tol:
                return True
            prev = direction
        return False

    @property
    def energy(self):
        return self.energies[-1]


# ------------------------------------------------------------------ DMRG1 ---
class EffHam1(spla.LinearOperator):
    """One-site effective Hamiltonian L - W - R (dmrg.py:756-801)."""

    def __init__(self, Lenv, W, Renv, dims):
        self.L, self.W, self.R = Lenv, W, Renv
        self.dims = dims  # (a, s, b)
        n = int(np.prod(dims))
        self.nmatvec = 0
        super().__init__(dtype=Lenv.dtype, shape=(n, n))

    def _matvec(self, v):
        self.nmatvec += 1
        x = v.reshape(self.dims)                                 # a s b
        T = np.tensordot(self.L, x, axes=(2, 0))                 # a' w s b
        T = np.tensordot(T, self.W, axes=((1, 2), (0, 3)))       # a' b w' s'
        T = np.tensordot(T, self.R, axes=((2, 1), (1, 2)))       # a' s' b'
        return T.reshape(-1)


class DMRG1(DMRG2):
    """quimb's DMRG1 = DMRG(bsz=1) (dmrg.py:756-801, 1100-1106, 1137-1156):
    bonds are padded with noise of strength 1e-6 before every sweep
    (expand_bond_dimension), each site is solved and the orthogonality centre
    moved on by a stabilised QR / LQ."""

    def __init__(self, mpo, bond_dims, cutoffs=1e-8, p0=None, seed=0):
        super().__init__(mpo, bond_dims, cutoffs, p0=p0, seed=seed)
        self._rng = np.random.default_rng(seed + 12345)

    def expand(self, new_bond_dim, rand_strength=1e-6):
        for i in range(self.L - 1):
            cur = self.k[i].shape[2]
            if cur >= new_bond_dim:
                continue
            ex = new_bond_dim - cur
            a, b = self.k[i], self.k[i + 1]
            pa = rand_strength * self._rng.standard_normal((a.shape[0], a.shape[1], ex))
            pb = rand_strength * self._rng.standard_normal((ex, b.shape[1], b.shape[2]))
            self.k[i] = np.concatenate([a, pa.astype(a.dtype)], axis=2)
            self.k[i + 1] = np.concatenate([b, pb.astype(b.dtype)], axis=0)

    def _init_right_envs(self):
        dt = self.k[0].dtype
        self.renv = {self.L - 1: np.ones((1, 1, 1), dtype=dt)}
        for i in range(self.L - 1, 0, -1):
            self.renv[i - 1] = env_step_right(self.renv[i], self.k[i], self.mpo[i])

    def _init_left_envs(self):
        dt = self.k[0].dtype
        self.lenv = {0: np.ones((1, 1, 1), dtype=dt)}
        for i in range(0, self.L - 1):
            self.lenv[i + 1] = env_step_left(self.lenv[i], self.k[i], self.mpo[i])

    def _update_1site(self, i, direction):
        A = self.k[i]
        dims = A.shape
        Heff = EffHam1(self.lenv[i], self.mpo[i], self.renv[i], dims)
        v0 = A.reshape(-1)
        n = v0.size
        if n < 800:
            Hd = Heff @ np.eye(n)
            Hd = 0.5 * (Hd + Hd.conj().T)
            evals, evecs = np.linalg.eigh(Hd)
            loc_en, x = evals[0], evecs[:, 0]
        else:
            lk, vk = spla.eigsh(Heff, k=1, which="SA", v0=v0, ncv=self.opts["local_eig_ncv"],
                                tol=self.opts["local_eig_tol"])
            loc_en, x = lk[0], vk[:, 0]
        self.nmatvecs.append(Heff.nmatvec)
        self.k[i] = x.reshape(dims)
        tot_en = np.vdot(x, Heff._matvec(x))
        if direction == "right" and i < self.L - 1:
            a, d, r = dims
            Q, _, R = dn.qr_stabilized(self.k[i].reshape(a * d, r).copy(), absorb=dn.get_U_sVH)
            self.k[i] = Q.reshape(a, d, -1)
            self.k[i + 1] = np.tensordot(R, self.k[i + 1], axes=(1, 0))
        elif direction == "left" and i > 0:
            a, d, r = dims
            Lf, _, Q = dn.qr_stabilized(self.k[i].reshape(a, d * r).copy(), absorb=dn.get_Us_VH)
            self.k[i] = Q.reshape(-1, d, r)
            self.k[i - 1] = np.tensordot(self.k[i - 1], Lf, axes=(2, 0))
        return float(np.real(loc_en)), float(np.real(tot_en)), {}

    def sweep(self, direction, canonize=True, max_bond=None, cutoff=0.0):
        L = self.L
        loc, tot = [], []
        if direction == "R":
            if canonize:
                self.k = right_canonize(self.k)
            self._init_right_envs()
            self.lenv = {0: np.ones((1, 1, 1), dtype=self.k[0].dtype)}
            for i in range(L):
                if i > 0:
                    self.lenv[i] = env_step_left(self.lenv[i - 1], self.k[i - 1],
                                                 self.mpo[i - 1])
                le, te, _ = self._update_1site(i, "right")
                loc.append(le); tot.append(te)
        else:
            if canonize:
                self.k = left_canonize(self.k)
            self._init_left_envs()
            self.renv = {L - 1: np.ones((1, 1, 1), dtype=self.k[0].dtype)}
            for i in range(L - 1, -1, -1):
                if i < L - 1:
                    self.renv[i] = env_step_right(self.renv[i + 1], self.k[i + 1],
                                                  self.mpo[i + 1])
                le, te, _ = self._update_1site(i, "left")
                loc.append(le); tot.append(te)
        self.local_energies.append(tuple(loc))
        self.total_energies.append(tuple(tot))
        return tot[-1]

    def solve(self, tol=1e-4, max_sweeps=10, sweep_sequence=None):
        seq = sweep_sequence or self.opts["default_sweep_sequence"]
        prev = "0"
        for n in range(max_sweeps):
            direction = seq[n % len(seq)]
            idx = self._sweep_idx
            max_bond = self.bond_dims[min(idx, len(self.bond_dims) - 1)]
            self._sweep_idx += 1
            canonize = (direction + prev) not in ("LR", "RL")
            self.expand(max_bond)
            en = self.sweep(direction, canonize=canonize)
            self.energies.append(en)
            if len(self.energies) >= 2 and abs(self.energies[-2] - self.energies[-1]) < tol:
                return True
            prev = direction
        return False
