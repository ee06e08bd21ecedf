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
