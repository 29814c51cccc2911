"""Device-resident thin QR / SVD against numpy.linalg on the same matrices: the defining properties
(isometry, triangularity, reconstruction, singular values) — the factors themselves are unique only up to phases."""
import numpy as np
import pytest

from tolerances import C128_BOUND

pytestmark = pytest.mark.gpu


def rand(rng, shape, dt):
    x = rng.standard_normal(shape)
    if np.dtype(dt).kind == "c":
        x = x + 1j * rng.standard_normal(shape)
    return x.astype(dt)


@pytest.mark.parametrize("dt,tol", [(np.complex128, 100 * C128_BOUND), (np.float64, 100 * C128_BOUND), (np.complex64, 2e-5), (np.float32, 2e-5)])
@pytest.mark.parametrize("shape,rows", [((8, 2, 16), ("l", "p")), ((8, 2, 16), ("r",)), ((4, 3, 5, 2), ("b", "d"))])
def test_qr_thin(ctx, dt, tol, shape, rows):
    """canonize.jl:58: Q, R = tensor_qr_thin(A; inds_q, inds_r, ind_virtual), A an MPS site (l, p, r) or any tensor."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(len(shape) * 7 + len(rows))
    inds = ("l", "p", "r") if len(shape) == 3 else ("a", "b", "c", "d")
    a = rand(rng, shape, dt)
    t = tb.Tensor(a, inds)
    q, r = tb.tensor_qr_thin(t, rows, ind_virtual="v")
    cols = [i for i in inds if i not in rows]
    assert q.inds == tuple(rows) + ("v",) and r.inds == ("v",) + tuple(cols)
    m = int(np.prod([t.size(i) for i in rows])); n = int(np.prod([t.size(i) for i in cols])); k = min(m, n)
    Q = np.reshape(q.parent, (m, k), order="F").astype(np.complex128)
    R = np.reshape(r.parent, (k, n), order="F").astype(np.complex128)
    A = np.reshape(np.transpose(a, [inds.index(i) for i in list(rows) + cols]), (m, n), order="F").astype(np.complex128)
    assert np.abs(Q.conj().T @ Q - np.eye(k)).max() < tol * 10
    assert np.abs(np.tril(R, -1)).max() == 0
    assert np.abs(Q @ R - A).max() / np.abs(A).max() < tol * 10
    # the conj flag of the operand is honoured (no materialised copy on the caller's side)
    q2, r2 = tb.tensor_qr_thin(t.conj(), rows, ind_virtual="v")
    Q2 = np.reshape(q2.parent, (m, k), order="F").astype(np.complex128)
    R2 = np.reshape(r2.parent, (k, n), order="F").astype(np.complex128)
    assert np.abs(Q2 @ R2 - A.conj()).max() / np.abs(A).max() < tol * 10


@pytest.mark.parametrize("dt,tol", [(np.complex128, 100 * C128_BOUND), (np.float64, 100 * C128_BOUND), (np.complex64, 2e-5)])
@pytest.mark.parametrize("shape,rows", [((8, 2, 16), ("l", "p")), ((16, 2, 4), ("l",)), ((6, 2, 2, 6), ("a", "b"))])
def test_svd_thin(ctx, dt, tol, shape, rows):
    """canonize.jl:41, evolve.jl:62, DMRG.jl:338: U, s, V = tensor_svd_thin(A; inds_u, inds_v, ind_s), tall and wide."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(sum(shape))
    inds = ("l", "p", "r") if len(shape) == 3 else ("a", "b", "c", "d")
    a = rand(rng, shape, dt)
    t = tb.Tensor(a, inds)
    u, s, v = tb.tensor_svd_thin(t, rows, ind_s="s")
    cols = [i for i in inds if i not in rows]
    m = int(np.prod([t.size(i) for i in rows])); n = int(np.prod([t.size(i) for i in cols])); k = min(m, n)
    U = np.reshape(u.parent, (m, k), order="F").astype(np.complex128)
    S = s.parent.astype(np.complex128)
    V = np.reshape(v.parent, (k, n), order="F").astype(np.complex128)
    A = np.reshape(np.transpose(a, [inds.index(i) for i in list(rows) + cols]), (m, n), order="F").astype(np.complex128)
    sref = np.linalg.svd(A, compute_uv=False)
    assert np.abs(S.imag).max() == 0 and np.all(np.diff(S.real) <= 1e-6 * sref[0])
    assert np.abs(S.real - sref).max() / sref[0] < tol * 10
    assert np.abs(U.conj().T @ U - np.eye(k)).max() < tol * 10
    assert np.abs(V @ V.conj().T - np.eye(k)).max() < tol * 10
    assert np.abs((U * S) @ V - A).max() / np.abs(A).max() < tol * 10
    # canonize.jl:41-47: absorb s into V with binary_einsum(s, V; dims=[]) and rebuild A with one more einsum, all on the device
    sv = tb.binary_einsum(s, v, dims=())
    back = tb.binary_einsum(u, sv)
    B = np.reshape(np.transpose(back.parent, [back.inds.index(i) for i in list(rows) + cols]), (m, n), order="F")
    assert np.abs(B - A).max() / np.abs(A).max() < tol * 10
