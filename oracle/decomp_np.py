"""ORACLE (test infrastructure, never imported by the product package).

Plain-numpy restatement of the split half of quimb's hot path, following the
*numpy backend* of the reference (its numba-accelerated overrides, which are
what `method="svd"` / `"qr"` run on numpy arrays):

  svd_truncated          quimb/tensor/decomp.py:1058-1118 -> :1032-1055
  number of values kept  decomp.py:901-937  (_compute_number_svals_to_keep_numba)
  renorm factor          decomp.py:940-965  (_compute_svals_renorm_factor_numba)
  trim / renorm / error  decomp.py:968-1029 (_trim_and_renorm_svd_result_numba)
  absorb                 decomp.py:693-721  (_do_absorb_numba)
  qr_stabilized          decomp.py:2147-2216 (_qr_stabilized_numba / _lq_...)
  sgn                    decomp.py:634-648
  option parsing         decomp.py:201-291, 369-424 (codes only)
  svd:eig                decomp.py:1168-1361 (svd_via_eig), :1364-1444
  svd:rand               decomp.py:1689-1861 (svd_rand_truncated)
  eigh                   decomp.py:1899-1969 (+ _with_diag_shift :1867-1881)
  safe_inverse           decomp.py:501-551
  fuse / tensor_split    quimb/tensor/array_ops.py:95-180,
                         quimb/tensor/tensor_core.py:392-668 (array part)

Pinned against the reference itself by oracle/make_golden.py (the reference
is executed under the shims in oracle/shims) and against the known answers
of the reference's tests (tests/test_tensor/test_decomp.py:52-57,
test_tensor_core.py:677-716).
"""

import numpy as np

# absorb codes (decomp.py:201-211); None = 'U,s,VH'
get_s = 2
get_Usq = -12
get_VH = -11
get_Us = -10
get_Us_VH = -1
get_Usq_sqVH = 0
get_U_sVH = 1
get_U = 10
get_sVH = 11
get_sqVH = 12

ABSORB_MAP = {
    None: None, "U,s,VH": None,
    get_s: get_s, "s": get_s,
    get_Usq: get_Usq, "lsqrt": get_Usq,
    get_VH: get_VH, "VH": get_VH, "rorthog": get_VH,
    get_Us: get_Us, "Us": get_Us, "lfactor": get_Us,
    get_Us_VH: get_Us_VH, "Us,VH": get_Us_VH, "left": get_Us_VH,
    get_Usq_sqVH: get_Usq_sqVH, "Usq,sqVH": get_Usq_sqVH, "both": get_Usq_sqVH,
    get_U_sVH: get_U_sVH, "U,sVH": get_U_sVH, "right": get_U_sVH,
    get_U: get_U, "U": get_U, "lorthog": get_U,
    get_sVH: get_sVH, "sVH": get_sVH, "rfactor": get_sVH,
    get_sqVH: get_sqVH, "sqVH": get_sqVH, "rsqrt": get_sqVH,
}

CUTOFF_MODE_MAP = {
    1: 1, "abs": 1, 2: 2, "rel": 2, 3: 3, "sum2": 3, 4: 4, "rsum2": 4,
    5: 5, "sum1": 5, 6: 6, "rsum1": 6,
}
RENORM_LOOKUP = {3: 2, 4: 2, 5: 1, 6: 1}


def parse_truncation_opts(max_bond=None, cutoff=1e-10, cutoff_mode="rsum2",
                          renorm=None):
    """Numeric codes as produced by parse_split_opts (decomp.py:369-424)."""
    max_bond = -1 if max_bond is None else max_bond
    cutoff = -1.0 if cutoff is None else cutoff
    cutoff_mode = CUTOFF_MODE_MAP[cutoff_mode]
    if renorm is True:
        renorm = RENORM_LOOKUP.get(cutoff_mode, 0)
    else:
        renorm = 0 if renorm is None else renorm
    return dict(max_bond=max_bond, cutoff=cutoff, cutoff_mode=cutoff_mode,
                renorm=int(renorm))


def number_svals_to_keep(s, cutoff, cutoff_mode):
    """decomp.py:901-937."""
    s = np.asarray(s)
    if cutoff_mode == 1:
        n_chi = int(np.sum(s > cutoff))
    elif cutoff_mode == 2:
        n_chi = int(np.sum(s > cutoff * s[0]))
    else:
        pw = 2 if cutoff_mode in (3, 4) else 1
        target = cutoff
        if cutoff_mode in (4, 6):
            target *= np.sum(s ** pw)
        n_chi = s.size
        ssum = 0.0
        for i in range(s.size - 1, -1, -1):
            s2 = s[i] ** pw
            if not np.isnan(s2):
                ssum += s2
            if ssum > target:
                break
            n_chi -= 1
    return max(n_chi, 1)


def svals_renorm_factor(s, n_chi, renorm):
    """decomp.py:940-965."""
    keep = lose = 0.0
    raise_power = renorm >= 2
    for i in range(s.size):
        s2 = s[i]
        if raise_power:
            s2 = s2 ** renorm
        if not np.isnan(s2):
            if i < n_chi:
                keep += s2
            else:
                lose += s2
    f = (keep + lose) / keep
    if raise_power:
        f = f ** (1 / renorm)
    return f


def do_absorb(U, s, VH, absorb):
    """decomp.py:693-721."""
    if absorb is None:
        return U, s, VH
    if absorb == get_Usq_sqVH:
        sq = np.sqrt(s)
        return U * sq[None, :], None, sq[:, None] * VH
    if absorb == get_U_sVH:
        return U, None, s[:, None] * VH
    if absorb == get_Us_VH:
        return U * s[None, :], None, VH
    if absorb == get_sVH:
        return None, None, s[:, None] * VH
    if absorb == get_Us:
        return U * s[None, :], None, None
    if absorb == get_U:
        return U, None, None
    if absorb == get_VH:
        return None, None, VH
    if absorb == get_Usq:
        return U * np.sqrt(s)[None, :], None, None
    if absorb == get_sqVH:
        return None, None, np.sqrt(s)[:, None] * VH
    if absorb == get_s:
        return None, s, None
    raise ValueError(f"Invalid absorb mode: {absorb}")


def trim_and_renorm(U, s, VH, cutoff, cutoff_mode, max_bond, absorb, renorm,
                    use_abs=False):
    """decomp.py:968-1029; returns (left, s, right, error, n_keep)."""
    sabs = np.abs(s) if use_abs else s
    error = 0.0
    n_keep = s.size
    if (cutoff > 0.0) or (renorm > 0):
        n_chi = number_svals_to_keep(sabs, cutoff, cutoff_mode)
        if max_bond > 0:
            n_chi = min(n_chi, max_bond)
        if n_chi < s.size:
            error = float(np.sqrt(np.sum(sabs[n_chi:] ** 2)))
            if renorm > 0:
                f = svals_renorm_factor(sabs, n_chi, renorm)
                s = s[:n_chi] * f
            else:
                s = s[:n_chi]
            U = U[:, :n_chi]
            VH = VH[:n_chi, :]
            n_keep = n_chi
    elif (max_bond != -1) and (max_bond < s.shape[0]):
        error = float(np.sqrt(np.sum(sabs[max_bond:] ** 2)))
        U = U[:, :max_bond]
        s = s[:max_bond]
        VH = VH[:max_bond, :]
        n_keep = max_bond
    s = np.ascontiguousarray(s)
    left, sv, right = do_absorb(U, s, VH, absorb)
    return left, sv, right, error, n_keep


def svd_truncated(x, cutoff=-1.0, cutoff_mode=4, max_bond=-1, absorb=0,
                  renorm=0, info=None):
    """decomp.py:1032-1118 for 2-d numpy input."""
    absorb = ABSORB_MAP[absorb]
    cutoff_mode = CUTOFF_MODE_MAP[cutoff_mode]
    U, s, VH = np.linalg.svd(x, full_matrices=False)
    left, sv, right, error, n_keep = trim_and_renorm(
        U, s, VH, cutoff, cutoff_mode, max_bond, absorb, renorm)
    if info is not None:
        info["error"] = error
        info["n_keep"] = n_keep
        info["svals"] = s
    return left, sv, right


def safe_inverse(x, cutoff=None, power=1.0):
    """decomp.py:501-551."""
    xmax = np.max(x) if x.ndim == 1 else np.expand_dims(np.max(x, axis=-1), -1)
    xmax = np.where(xmax > 0.0, xmax, 1.0)
    c = np.finfo(x.dtype).eps if cutoff is None else cutoff / xmax
    y = x / xmax
    q = power + 1.0
    return y / ((y ** q + c ** q) * xmax ** power)


def _dag(x):
    return np.conj(np.swapaxes(x, -2, -1))


def svd_via_eig(x, absorb=None, max_bond=-1, descending=True, right=None):
    """decomp.py:1168-1361 (the generic, array-API version)."""
    m, n = x.shape
    xdag = _dag(x)
    absorb = ABSORB_MAP[absorb]
    if right is None:
        if m > n:
            right = True
        elif m < n:
            right = False
        else:
            right = absorb in (get_VH, get_sVH, get_sqVH, get_Us_VH)
    if right:
        s2, V = np.linalg.eigh(xdag @ x)
        if 0 < max_bond < min(m, n):
            s2 = s2[-max_bond:]
            V = V[:, -max_bond:]
        if descending:
            s2 = np.flip(s2, axis=-1)
            V = np.flip(V, axis=-1)
        s2 = np.clip(s2, 0.0, None)
        if absorb == get_s:
            return None, np.sqrt(s2), None
        if absorb == get_VH:
            return None, None, _dag(V)
        if absorb == get_sVH:
            return None, None, np.sqrt(s2)[:, None] * _dag(V)
        if absorb == get_sqVH:
            return None, None, np.sqrt(np.sqrt(s2))[:, None] * _dag(V)
        Us = x @ V
        if absorb == get_Us:
            return Us, None, None
        if absorb == get_Us_VH:
            return Us, None, _dag(V)
        s = np.sqrt(s2)
        eps = np.finfo(s.dtype).eps
        smax = s[0:1] if descending else s[-1:]
        sinv = safe_inverse(s, smax * eps * max(m, n))
        U = Us * sinv[None, :]
        if absorb == get_U:
            return U, None, None
        if absorb == get_Usq:
            return U * np.sqrt(s)[None, :], None, None
        VH = _dag(V)
        if absorb is None:
            return U, s, VH
        if absorb == get_U_sVH:
            return U, None, s[:, None] * VH
        if absorb == get_Usq_sqVH:
            sq = np.sqrt(s)
            return U * sq[None, :], None, sq[:, None] * VH
    else:
        s2, U = np.linalg.eigh(x @ xdag)
        if 0 < max_bond < min(m, n):
            s2 = s2[-max_bond:]
            U = U[:, -max_bond:]
        if descending:
            s2 = np.flip(s2, axis=-1)
            U = np.flip(U, axis=-1)
        s2 = np.clip(s2, 0.0, None)
        if absorb == get_s:
            return None, np.sqrt(s2), None
        if absorb == get_U:
            return U, None, None
        if absorb == get_Us:
            return U * np.sqrt(s2)[None, :], None, None
        if absorb == get_Usq:
            return U * np.sqrt(np.sqrt(s2))[None, :], None, None
        sVH = _dag(U) @ x
        if absorb == get_sVH:
            return None, None, sVH
        if absorb == get_U_sVH:
            return U, None, sVH
        s = np.sqrt(s2)
        eps = np.finfo(s.dtype).eps
        smax = s[0:1] if descending else s[-1:]
        sinv = safe_inverse(s, smax * eps * max(m, n))
        VH = sinv[:, None] * sVH
        if absorb == get_VH:
            return None, None, VH
        if absorb is None:
            return U, s, VH
        if absorb == get_Us_VH:
            return U * s[None, :], None, VH
        sq = np.sqrt(s)
        if absorb == get_Usq_sqVH:
            return U * sq[None, :], None, sq[:, None] * VH
        if absorb == get_sqVH:
            return None, None, sq[:, None] * VH
    raise ValueError(f"Invalid absorb mode: {absorb}")


def svd_via_eig_truncated(x, cutoff=-1.0, cutoff_mode=4, max_bond=-1, absorb=0,
                          renorm=0, info=None):
    """decomp.py:1364-1444."""
    absorb = ABSORB_MAP[absorb]
    cutoff_mode = CUTOFF_MODE_MAP[cutoff_mode]
    need_full = (cutoff > 0.0) or (renorm > 0) or (info is not None and "error" in info)
    if need_full:
        U, s, VH = svd_via_eig(x, absorb=None, max_bond=-1, descending=True)
        left, sv, right, error, n_keep = trim_and_renorm(
            U, s, VH, cutoff, cutoff_mode, max_bond, absorb, renorm)
        if info is not None:
            info["error"] = error
            info["n_keep"] = n_keep
        return left, sv, right
    return svd_via_eig(x, absorb=absorb, max_bond=max_bond, descending=False)


def with_diag_shift(x, shift=0.0):
    """decomp.py:1867-1881."""
    if shift < 0.0:
        shift = np.finfo(x.dtype).eps
    if shift > 0.0:
        x = x + shift * np.trace(x) * np.eye(x.shape[-1])
    return x


def eigh_truncated(x, cutoff=-1.0, cutoff_mode=4, max_bond=-1, absorb=0, renorm=0,
                   positive=0, shift=False):
    """decomp.py:1899-1969 / numba :1972-2020."""
    absorb = ABSORB_MAP[absorb]
    cutoff_mode = CUTOFF_MODE_MAP[cutoff_mode]
    shift = {False: 0.0, True: -1.0}.get(shift, shift)
    x = with_diag_shift(x, shift)
    s, U = np.linalg.eigh(x)
    if not positive:
        k = np.argsort(-np.abs(s))
        s, U = s[k], U[:, k]
    else:
        s = s[::-1].copy()
        U = U[:, ::-1]
        if absorb in (get_Usq_sqVH, get_Usq, get_sqVH):
            s[s < 0.0] = 0.0
    VH = _dag(U)
    left, sv, right, _, _ = trim_and_renorm(U, s, VH, cutoff, cutoff_mode, max_bond,
                                            absorb, renorm, use_abs=not positive)
    return left, sv, right


_ABSORB_TRANSPOSE = {get_U_sVH: get_Us_VH, get_U: get_VH, get_sVH: get_Us,
                     get_Us_VH: get_U_sVH, get_VH: get_U, get_Us: get_sVH}


def cholesky_regularized(x, absorb=get_Usq_sqVH, shift=True):
    """decomp.py:2245-2322 (numba :2325-2335): (L, None, L^H) of a positive
    definite matrix; ``shift`` True -> eps * trace on the diagonal, 'auto' ->
    retry with it after a failure, float -> relative shift."""
    absorb = ABSORB_MAP[absorb]

    def run(sh):
        L = np.linalg.cholesky(with_diag_shift(x, sh))
        if absorb == get_Usq:
            return L, None, None
        if absorb == get_sqVH:
            return None, None, _dag(L)
        if absorb == get_Usq_sqVH:
            return L, None, _dag(L)
        raise ValueError(f"Invalid absorb={absorb} in cholesky_regularized.")
    if isinstance(shift, str) and shift == "auto":
        try:
            return run(0.0)
        except np.linalg.LinAlgError:
            return run(-1.0)
    return run({False: 0.0, True: -1.0}.get(shift, shift))


def qr_via_cholesky(x, absorb=get_Us_VH, shift=True):
    """decomp.py:2359-2424: LQ-like split x = L Q from the Cholesky factor of
    x x^H (QR-like through the transpose)."""
    import scipy.linalg as sla
    absorb = ABSORB_MAP[absorb]
    if absorb in (get_U_sVH, get_U, get_sVH):
        transposed = True
    elif absorb in (get_Us_VH, get_Us, get_VH):
        transposed = False
    else:
        raise ValueError(f"Invalid absorb mode for qr_via_cholesky: {absorb}")
    if transposed:
        absorb = _ABSORB_TRANSPOSE[absorb]
        xT = x.T
        xx = xT @ x.conj()
        x = xT
    else:
        xx = x @ x.conj().T
    L, _, _ = cholesky_regularized(xx, absorb=get_Usq, shift=shift)
    right = sla.solve_triangular(L, x, lower=True) if absorb != get_Us else None
    left = L if absorb != get_VH else None
    if transposed:
        left, right = (None if right is None else right.T), (None if left is None else left.T)
    return left, None, right


def polar_right(x):
    """decomp.py:2673-2700: x = U P, U isometric, P positive semi-definite."""
    W, s, VH = np.linalg.svd(x, full_matrices=False)
    return W @ VH, None, _dag(VH) @ (s[:, None] * VH)


def polar_left(x):
    """decomp.py:2703-2730: x = P U."""
    W, s, VH = np.linalg.svd(x, full_matrices=False)
    return (W * s[None, :]) @ _dag(W), None, W @ VH


def svd_rand_truncated(x, max_bond, absorb=0, oversample=10, num_iterations=2,
                       right=None, seed=None):
    """decomp.py:1689-1861 with method_lorthog='qr', method_reduced='svd'."""
    absorb = ABSORB_MAP[absorb]
    if max_bond is None:
        max_bond = -1
    m, n = x.shape
    k = min(m, n) if max_bond < 0 else min(m, n, max_bond)
    k_sketch = min(m, n, k + oversample)
    if right is None:
        if absorb in (get_U_sVH, get_U, get_sVH):
            right = True
        elif absorb in (get_Us_VH, get_Us, get_VH):
            right = False
        else:
            right = m > n
    rng = np.random.default_rng(seed)
    xdag = _dag(x)
    if right:
        y = x @ rng.normal(size=(n, k_sketch))
        for _ in range(num_iterations):
            y = x @ (xdag @ y)
        Q, _, _ = qr_stabilized(y, absorb=get_U)
        if k >= k_sketch:
            if absorb == get_U_sVH:
                return Q, None, _dag(Q) @ x
            if absorb == get_sVH:
                return None, None, _dag(Q) @ x
            if absorb == get_U:
                return Q, None, None
        B = _dag(Q) @ x
    else:
        y = rng.normal(size=(k_sketch, m)) @ x
        for _ in range(num_iterations):
            y = (y @ xdag) @ x
        Q, _, _ = qr_stabilized(_dag(y), absorb=get_U)
        if k >= k_sketch:
            if absorb == get_Us_VH:
                return x @ Q, None, _dag(Q)
            if absorb == get_Us:
                return x @ Q, None, None
            if absorb == get_VH:
                return None, None, _dag(Q)
        B = x @ Q
    U, s, VH = svd_truncated(B, cutoff=0.0, max_bond=k, absorb=absorb)
    if U is not None and right:
        U = Q @ U
    if VH is not None and not right:
        VH = VH @ _dag(Q)
    return U, s, VH


def sgn(x):
    """decomp.py:634-648: x / sgn(x) is real and non-negative; sgn(0) = 1."""
    x0 = x == 0.0
    return (x + x0) / (np.abs(x) + x0)


def qr_stabilized(x, absorb=get_U_sVH, stabilized=True):
    """decomp.py:2147-2216 (2-d numpy path incl. the LQ variants)."""
    absorb = ABSORB_MAP[absorb]
    if absorb in (get_U_sVH, get_U, get_sVH):
        return _qr_stab(x, absorb, stabilized)
    if absorb in (get_Us_VH, get_Us, get_VH):
        absorb_t = {get_Us: get_sVH, get_VH: get_U}.get(absorb, get_U_sVH)
        Q, _, L = _qr_stab(x.T, absorb_t, stabilized)
        if absorb == get_Us:
            return L.T, None, None
        if absorb == get_VH:
            return None, None, Q.T
        return L.T, None, Q.T
    raise ValueError(f"Invalid absorb mode for qr_stabilized: {absorb}")


def _qr_stab(x, absorb, stabilized):
    Q, R = np.linalg.qr(x)
    if stabilized:
        for i in range(R.shape[0]):
            phase = sgn(R[i, i])
            if phase != 1.0:
                if absorb != get_sVH:
                    Q[:, i] *= np.conj(phase)
                if absorb != get_U:
                    R[i, i:] *= phase
    if absorb == get_U:
        return Q, None, None
    if absorb == get_sVH:
        return None, None, R
    return Q, None, R


# ---- fuse / tensor_split (array level) -------------------------------------
def calc_fuse_perm_and_shape(shape, axes_groups):
    """array_ops.py:95-145: groups are fused and placed at the position of
    the first axis of the first group; ungrouped axes keep their order."""
    ndim = len(shape)
    grouped = set(ax for g in axes_groups for ax in g)
    first = min(min(g) for g in axes_groups if g) if any(axes_groups) else 0
    perm, new_shape = [], []
    placed = False
    for ax in range(ndim):
        if ax in grouped:
            if not placed and ax == first:
                for g in axes_groups:
                    perm.extend(g)
                    n = 1
                    for a in g:
                        n *= shape[a]
                    new_shape.append(n)
                placed = True
            continue
        perm.append(ax)
        new_shape.append(shape[ax])
    return tuple(perm), tuple(new_shape)


def fuse(x, *axes_groups):
    perm, new_shape = calc_fuse_perm_and_shape(x.shape, axes_groups)
    return np.transpose(x, perm).reshape(new_shape)


_QR_ABSORBS = (get_U_sVH, get_U, get_sVH, get_Us_VH, get_Us, get_VH)


def parse_method_absorb(method="auto", absorb="auto", truncation=True):
    """decomp.py:307-365."""
    if method == "auto":
        if truncation or absorb == "auto":
            method = "svd"
        else:
            absorb = ABSORB_MAP[absorb]
            method = "qr" if absorb in _QR_ABSORBS else "svd"
    if method.startswith("lq"):
        method = "qr" + method[2:]
        if absorb == "auto":
            absorb = "left"
    if absorb == "auto":
        absorb = {"qr": get_U_sVH}.get(method, get_Usq_sqVH)
    else:
        absorb = ABSORB_MAP[absorb]
    return method, absorb


def tensor_split(x, inds, left_inds, right_inds=None, method="auto",
                 absorb="auto", max_bond=None, cutoff=1e-10,
                 cutoff_mode="rel", renorm=None, info=None):
    """Array-level restatement of tensor_split (tensor_core.py:392-668) with
    get='arrays': returns (left, [s], right) with the new bond last on the
    left factor and first on the right factor."""
    inds = tuple(inds)
    left_inds = tuple(left_inds)
    if right_inds is None:
        right_inds = tuple(ix for ix in inds if ix not in left_inds)
    perm = [inds.index(ix) for ix in left_inds + tuple(right_inds)]
    xt = np.transpose(x, perm)
    ldims = xt.shape[: len(left_inds)]
    rdims = xt.shape[len(left_inds):]
    mat = xt.reshape(int(np.prod(ldims)), int(np.prod(rdims)))
    mb = -1 if max_bond is None else max_bond
    co = -1.0 if cutoff is None else cutoff
    method, absorb = parse_method_absorb(method, absorb, (mb > 0) or (co > 0.0))
    if method == "svd":
        opts = parse_truncation_opts(max_bond, cutoff, cutoff_mode, renorm)
        left, s, right = svd_truncated(mat, absorb=absorb, info=info, **opts)
    elif method == "qr":
        left, s, right = qr_stabilized(mat, absorb=absorb)
    else:
        raise ValueError(f"oracle: unsupported split method {method!r}")
    if left is not None:
        left = left.reshape(*ldims, -1)
    if right is not None:
        right = right.reshape(-1, *rdims)
    return left, s, right


# ---- bond canonisation / compression (array level) -------------------------
def tensor_canonize_bond(a, a_inds, b, b_inds, absorb="right"):
    """tensor_core.py:671-824: QR ``a`` over the bond it shares with ``b`` and
    absorb R into ``b`` (absorb='right'); LQ of ``b`` for absorb='left'.
    Index orders of both outputs are those of the inputs."""
    a_inds, b_inds = tuple(a_inds), tuple(b_inds)
    (bond,) = [ix for ix in a_inds if ix in b_inds]
    if absorb == "left":
        nb, na = tensor_canonize_bond(b, b_inds, a, a_inds, "right")
        return na, nb
    lix = tuple(ix for ix in a_inds if ix != bond)
    q, _, r = tensor_split(a, a_inds, lix, (bond,), method="qr")
    # q: (*lix, k), r: (k, bond)
    new_a = np.transpose(q, [(lix + (bond,)).index(ix) for ix in a_inds])
    nb = np.tensordot(r, b, axes=(1, b_inds.index(bond)))      # (k, rest of b...)
    rest = tuple(ix for ix in b_inds if ix != bond)
    new_b = np.transpose(nb, [((bond,) + rest).index(ix) for ix in b_inds])
    return new_a, new_b


def tensor_compress_bond(a, a_inds, b, b_inds, max_bond=None, cutoff=1e-10,
                         cutoff_mode="rel", absorb="both", renorm=None, info=None):
    """tensor_core.py:864-1094, ``reduced=True`` branch: QR(a) / LQ(b), SVD of
    the reduced core Ra Rb with truncation, factors folded back.  Output index
    orders equal the input ones."""
    a_inds, b_inds = tuple(a_inds), tuple(b_inds)
    (bond,) = [ix for ix in a_inds if ix in b_inds]
    lix = tuple(ix for ix in a_inds if ix != bond)
    rix = tuple(ix for ix in b_inds if ix != bond)
    qa, _, ra = tensor_split(a, a_inds, lix, (bond,), method="qr")            # (*lix,k1),(k1,bond)
    lb, _, qb = tensor_split(b, b_inds, (bond,), rix, method="lq")            # (bond,k2),(k2,*rix)
    core = ra @ lb                                                            # (k1, k2)
    opts = parse_truncation_opts(max_bond, cutoff, cutoff_mode, renorm)
    _, ab = parse_method_absorb("svd", absorb, True)
    cl, s, cr = svd_truncated(core, absorb=ab, info=info, **opts)            # (k1,k),(k,k2)
    na = np.tensordot(qa, cl, axes=(qa.ndim - 1, 0))                          # (*lix, k)
    nb = np.tensordot(cr, qb, axes=(1, 0))                                    # (k, *rix)
    new_a = np.transpose(na, [(lix + (bond,)).index(ix) for ix in a_inds])
    new_b = np.transpose(nb, [((bond,) + rix).index(ix) for ix in b_inds])
    return new_a, new_b
