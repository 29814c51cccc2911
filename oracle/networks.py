"""CPU ORACLE (test infrastructure only): numpy restatement of how the reference's front-ends lay out the tensors
that reach the hot path — written independently of tenet.jl_b200/components.py so the two can be cross-checked.

  ising_1d_mpo           /root/reference/src/Models/Ising.jl:12-30   (order (:i,:o,:l,:r), bond 3, boundary slices)
  MPO index naming       src/Components/MPO.jl:85-132
  MPS index naming       src/Components/MPS.jl:50-94;  rand: MPS.jl:113-165
  env sweep              src/Algorithms/DMRG.jl:106-115
"""
import numpy as np


def ising_1d_mpo(L, h, J):
    Id = np.eye(2)
    sx = np.array([[0.0, 1.0], [1.0, 0.0]])
    sz = np.array([[1.0, 0.0], [0.0, -1.0]])
    W = np.zeros((2, 2, 3, 3), dtype=np.complex128)
    W[:, :, 0, 0] = Id
    W[:, :, 1, 0] = sz
    W[:, :, 2, 0] = -h * sx
    W[:, :, 2, 1] = -J * sz
    W[:, :, 2, 2] = Id
    arrays = [W[:, :, 2, :]] + [W] * (L - 2) + [W[:, :, :, 0]]
    inds = []
    for i in range(1, L + 1):
        t = [("in", i), ("out", i)]
        if i > 1:
            t.append(("w", i - 1))
        if i < L:
            t.append(("w", i))
        inds.append(tuple(t))
    return arrays, inds


def mps_inds(n, tag, plug):
    """(:l,:o,:r) order, boundaries drop the missing bond."""
    inds = []
    for i in range(1, n + 1):
        t = []
        if i > 1:
            t.append((tag, i - 1))
        t.append((plug, i))
        if i < n:
            t.append((tag, i))
        inds.append(tuple(t))
    return inds


def expectation_network(mps_arrays, Ws, Winds):
    """<psi|H|psi>: ket plugs feed the operator inputs, operator outputs meet the conjugated bra."""
    n = len(mps_arrays)
    arrays = list(mps_arrays) + list(Ws) + [np.conj(a) for a in mps_arrays]
    inds = mps_inds(n, "ket", "in") + list(Winds) + mps_inds(n, "bra", "out")
    return arrays, inds


def product_expectation_network(thetas, Ws, Winds):
    n = len(thetas)
    arrs = []
    for k, t in enumerate(thetas):
        v = np.array([np.cos(t / 2), np.sin(t / 2)], dtype=np.complex128)
        a = v.reshape(1, 2, 1)
        if k == 0:
            a = a.reshape(2, 1)
        elif k == n - 1:
            a = a.reshape(1, 2)
        arrs.append(a)
    return expectation_network(arrs, Ws, Winds)


def sweep_steps(n):
    steps, cur = [], 3 * n
    steps.append((0, n))
    steps.append((cur, 2 * n))
    cur += 1
    for i in range(1, n):
        steps.append((cur, i)); cur += 1
        steps.append((cur, n + i)); cur += 1
        steps.append((cur, 2 * n + i)); cur += 1
    return steps


def rand_mps(n, chi, seed, dtype=np.complex128, p=2):
    """Right-canonical random MPS: site i = Q factor of the LQ of a Gaussian (chi_l x p*chi_r) matrix."""
    rng = np.random.default_rng(seed)
    arrays = []
    for i in range(1, n + 1):
        ii = (n + 1 - abs(2 * i - n - 1)) // 2
        cl, cr = min(chi, p ** min(ii - 1, 62)), min(chi, p ** min(ii, 62))
        if n % 2 == 1 and i == n // 2 + 1:
            cr = cl
        elif i > n // 2:
            cl, cr = cr, cl
        x = rng.standard_normal((cl, p * cr))
        if np.dtype(dtype).kind == "c":
            x = x + 1j * rng.standard_normal((cl, p * cr))
        q, _ = np.linalg.qr(x.conj().T)
        arrays.append(np.reshape(q.conj().T[:cl].astype(dtype), (cl, p, cr), order="F"))
    arrays[0] = arrays[0].reshape(p, -1, order="F")
    arrays[-1] = arrays[-1].reshape(-1, p, order="F")
    return arrays


def norm_network(mps_arrays):
    n = len(mps_arrays)
    arrays, inds = [], []
    ki, bi = mps_inds(n, "ket", "p"), mps_inds(n, "bra", "p")
    for k in range(n):
        arrays += [mps_arrays[k], np.conj(mps_arrays[k])]
        inds += [ki[k], bi[k]]
    return arrays, inds


def zipper_steps(n):
    steps, nl = [(0, 1)], 2 * n
    cur = nl
    for i in range(1, n):
        steps.append((cur, 2 * i))
        steps.append((cur + 1, 2 * i + 1))
        cur += 2
    return steps
