"""CPU ORACLE (test infrastructure only): dense state-vector simulation of a gate list, the independent check
for circuit-amplitude networks ("full amplitude on a scaled-down instance vs state-vector").
Nothing here restates the reference (it has no circuit simulator: ext/TenetPythonCallExt/pytket.jl only imports
Pauli strings); it pins the *network generator + contraction* against first principles."""
import numpy as np


def simulate(nq, gates):
    """gates: [(2x2 matrix, (q,)) | (2x2x2x2 tensor [o1,o2,i1,i2], (q1,q2))]; returns psi[q0,q1,...] from |0..0>."""
    psi = np.zeros((2,) * nq, dtype=np.complex128)
    psi[(0,) * nq] = 1.0
    for g, qs in gates:
        g = np.asarray(g, dtype=np.complex128)
        if len(qs) == 1:
            q = qs[0]
            psi = np.moveaxis(np.tensordot(g, psi, axes=([1], [q])), 0, q)
        else:
            a, b = qs
            psi = np.moveaxis(np.tensordot(g, psi, axes=([2, 3], [a, b])), [0, 1], [a, b])
    return psi


def amplitude(nq, gates, bits):
    return complex(simulate(nq, gates)[tuple(int(b) for b in bits)])
