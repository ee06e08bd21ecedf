"""Synthetic random-circuit amplitude networks (BASELINE configs[3]-like):
qubits on a line/grid, layers of random single-qubit unitaries and CZ-like
two-qubit gates, complex128.  Returns (arrays, inputs, output) of the closed
tensor network <b| C |0...0>, plus the exact amplitude from a dense
state-vector simulation (independent of any contraction code)."""

import numpy as np


def _rand_u2(rng):
    a = rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2))
    q, r = np.linalg.qr(a)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def _rand_u4(rng):
    a = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
    q, r = np.linalg.qr(a)
    return (q * (np.diag(r) / np.abs(np.diag(r)))).reshape(2, 2, 2, 2)


def random_circuit_amplitude(nq=8, depth=6, seed=0, bits=None):
    rng = np.random.default_rng(seed)
    bits = [0] * nq if bits is None else list(bits)
    arrays, inputs = [], []
    cur = []
    counter = [0]

    def new():
        counter[0] += 1
        return f"i{counter[0]}"
    psi = np.zeros([2] * nq, dtype=np.complex128)
    psi[(0,) * nq] = 1.0
    for q in range(nq):
        ix = new()
        arrays.append(np.array([1.0, 0.0], dtype=np.complex128)); inputs.append((ix,))
        cur.append(ix)
    for layer in range(depth):
        for q in range(nq):
            u = _rand_u2(rng)
            ix = new()
            arrays.append(u); inputs.append((ix, cur[q]))       # u[out, in]
            cur[q] = ix
            psi = np.moveaxis(np.tensordot(u, psi, axes=(1, q)), 0, q)
        for q in range(layer % 2, nq - 1, 2):
            g = _rand_u4(rng)                                    # g[o1, o2, i1, i2]
            o1, o2 = new(), new()
            arrays.append(g); inputs.append((o1, o2, cur[q], cur[q + 1]))
            cur[q], cur[q + 1] = o1, o2
            psi = np.moveaxis(np.tensordot(g, psi, axes=((2, 3), (q, q + 1))), (0, 1), (q, q + 1))
    for q in range(nq):
        v = np.zeros(2, dtype=np.complex128); v[bits[q]] = 1.0
        arrays.append(v); inputs.append((cur[q],))
    return arrays, inputs, (), complex(psi[tuple(bits)])


def _fsim(theta, phi):
    g = np.zeros((4, 4), dtype=np.complex128)
    g[0, 0] = 1.0
    g[1, 1] = g[2, 2] = np.cos(theta)
    g[1, 2] = g[2, 1] = -1j * np.sin(theta)
    g[3, 3] = np.exp(-1j * phi)
    return g.reshape(2, 2, 2, 2)


def _u3(theta, phi, lam):
    return np.array([[np.cos(theta / 2), -np.exp(1j * lam) * np.sin(theta / 2)],
                     [np.exp(1j * phi) * np.sin(theta / 2),
                      np.exp(1j * (phi + lam)) * np.cos(theta / 2)]], dtype=np.complex128)


def grid_bond_patterns(Lx, Ly):
    """The four nearest-neighbour bond patterns A, B, C, D of a square grid
    (SURVEY.md 8d, cfg4): horizontal even / odd, vertical even / odd."""
    q = lambda i, j: i * Ly + j  # noqa: E731
    A = [(q(i, j), q(i, j + 1)) for i in range(Lx) for j in range(0, Ly - 1, 2)]
    B = [(q(i, j), q(i, j + 1)) for i in range(Lx) for j in range(1, Ly - 1, 2)]
    C = [(q(i, j), q(i + 1, j)) for j in range(Ly) for i in range(0, Lx - 1, 2)]
    D = [(q(i, j), q(i + 1, j)) for j in range(Ly) for i in range(1, Lx - 1, 2)]
    return [A, B, C, D]


def random_grid_circuit_amplitude(Lx=3, Ly=3, depth=8, seed=3, bits=None, gate="fsim",
                                  dense=True):
    """BASELINE configs[3]: qubits on an Lx x Ly grid; every layer applies a
    random U3(theta, phi, lambda) (angles U(0, 2 pi)) to every qubit, then an
    fSim (or CZ) gate on one of the four bond patterns cycling A, B, C, D.
    Returns (arrays, inputs, output, exact amplitude or None)."""
    rng = np.random.default_rng(seed)
    nq = Lx * Ly
    bits = [0] * nq if bits is None else list(bits)
    pats = grid_bond_patterns(Lx, Ly)
    arrays, inputs, cur = [], [], []
    counter = [0]

    def new():
        counter[0] += 1
        return f"i{counter[0]}"
    psi = None
    if dense:
        psi = np.zeros([2] * nq, dtype=np.complex128)
        psi[(0,) * nq] = 1.0
    for q in range(nq):
        ix = new()
        arrays.append(np.array([1.0, 0.0], dtype=np.complex128)); inputs.append((ix,))
        cur.append(ix)
    for layer in range(depth):
        for q in range(nq):
            u = _u3(*rng.uniform(0, 2 * np.pi, size=3))
            ix = new()
            arrays.append(u); inputs.append((ix, cur[q]))
            cur[q] = ix
            if dense:
                psi = np.moveaxis(np.tensordot(u, psi, axes=(1, q)), 0, q)
        for a, b in pats[layer % 4]:
            if gate == "cz":
                g = np.diag([1, 1, 1, -1]).astype(np.complex128).reshape(2, 2, 2, 2)
            else:
                g = _fsim(*rng.uniform(0, 2 * np.pi, size=2))
            o1, o2 = new(), new()
            arrays.append(g); inputs.append((o1, o2, cur[a], cur[b]))
            cur[a], cur[b] = o1, o2
            if dense:
                psi = np.moveaxis(np.tensordot(g, psi, axes=((2, 3), (a, b))), (0, 1), (a, b))
    for q in range(nq):
        v = np.zeros(2, dtype=np.complex128); v[bits[q]] = 1.0
        arrays.append(v); inputs.append((cur[q],))
    amp = complex(psi[tuple(bits)]) if dense else None
    return arrays, inputs, (), amp
