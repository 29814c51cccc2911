"""GPU parity: tnb_binary_einsum (through the Python mirror of Muscle.binary_einsum) vs the numpy oracle on the
same seeded inputs.  Covers SURVEY §8a cases (i)-(vii)."""
import numpy as np
import pytest

from oracle import einsum_oracle as orc

pytestmark = pytest.mark.gpu

TOL = {np.dtype(np.complex128): 1e-12, np.dtype(np.float64): 1e-12,
       np.dtype(np.complex64): 2e-5, np.dtype(np.float32): 2e-5}


def rand(rng, shape, dt):
    dt = np.dtype(dt)
    x = rng.standard_normal(shape)
    if dt.kind == "c":
        x = x + 1j * rng.standard_normal(shape)
    return x.astype(dt)


def check(tb, a, ai, b, bi, dims=None, out=None, conj_a=False, conj_b=False):
    ta, tb_ = tb.Tensor(a, ai), tb.Tensor(b, bi)
    if conj_a:
        ta = ta.conj()
    if conj_b:
        tb_ = tb_.conj()
    c = tb.binary_einsum(ta, tb_, dims=dims, out=out)
    a128 = np.conj(a) if conj_a else a
    b128 = np.conj(b) if conj_b else b
    hi = np.complex128 if np.iscomplexobj(a) or np.iscomplexobj(b) else np.float64
    ref, ri = orc.binary_einsum(a128.astype(hi), ai, b128.astype(hi), bi, dims=dims, out_inds=out)
    assert tuple(c.inds) == tuple(ri)
    got = c.parent
    assert got.shape == ref.shape
    dt = np.result_type(a.dtype, b.dtype)
    if dt.kind in "iu":
        dt = np.dtype(np.float64)
    scale = max(np.abs(ref).max(), 1e-30)
    err = np.abs(got - ref).max() / scale
    assert err < TOL[np.dtype(dt)] * max(1, np.sqrt(a.size / max(ref.size, 1))), f"rel err {err:.3e}"
    return c


@pytest.mark.parametrize("dt", [np.complex128, np.complex64, np.float64, np.float32])
def test_matmul_like(ctx, dt):
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(0)
    for (m, k, n) in [(1, 1, 1), (3, 5, 7), (64, 64, 64), (65, 9, 130), (200, 33, 17), (128, 256, 128)]:
        check(tb, rand(rng, (m, k), dt), ("m", "k"), rand(rng, (k, n), dt), ("k", "n"))
        check(tb, rand(rng, (k, m), dt), ("k", "m"), rand(rng, (n, k), dt), ("n", "k"))


@pytest.mark.parametrize("dt", [np.complex128, np.complex64])
def test_mps_zipper_shapes(ctx, dt):
    """overlap.jl:42-47: [c',c]x[c,p,cr] -> [c',p,cr];  [c',p,cr]x[c',p,cr'] -> [cr,cr']."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(1)
    chi, p = 24, 2
    env = rand(rng, (chi, chi), dt)
    A = rand(rng, (chi, p, chi), dt)
    B = rand(rng, (chi, p, chi), dt)
    t = check(tb, env, ("b", "a"), A, ("a", "p", "ar"))
    check(tb, t.parent, t.inds, B, ("b", "p", "br"), conj_b=True)


def test_batch_dims_empty(ctx):
    """dims=Index[]: shared indices are kept (Hadamard / diagonal scaling) — canonize.jl:44, absorb.jl:31."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(2)
    a = rand(rng, (6, 5, 4), np.complex128)
    lam = rand(rng, (4,), np.complex128)
    check(tb, a, ("l", "p", "r"), lam, ("r",), dims=())
    b = rand(rng, (4, 5, 3), np.complex128)
    check(tb, a, ("l", "p", "r"), b, ("r", "p", "x"), dims=("r",))      # p is batch, r contracted


def test_outer_and_scalars(ctx):
    """no shared index -> outer product, incl. rank-0 operands (DMRG.jl:60-61); full contraction -> rank 0."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(3)
    a = rand(rng, (4, 3), np.complex128)
    s = np.array(2.5 - 1j)
    c = check(tb, s, (), a, ("i", "j"))
    assert c.inds == ("i", "j")
    check(tb, a, ("i", "j"), rand(rng, (5,), np.complex128), ("k",))
    check(tb, s, (), np.array(0.5 + 2j), ())
    c = check(tb, a, ("i", "j"), rand(rng, (4, 3), np.complex128), ("i", "j"))
    assert c.inds == () and c.parent.shape == ()


def test_large_k_splitk(ctx):
    """full contraction of two big vectors/tensors (amplitude-closing step): split-K path."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(4)
    for dt in (np.complex64, np.complex128):
        a = rand(rng, (64, 64, 32), dt) / 100
        b = rand(rng, (32, 64, 64), dt) / 100
        check(tb, a, ("x", "y", "z"), b, ("z", "y", "x"))
        check(tb, a, ("x", "y", "z"), b, ("z", "y", "w"))      # M=64 (x), N=64 (w), K=2048: few tiles, long K


def test_mixed_eltypes_and_views(ctx):
    """Int x Float (sample.jl:32-36), Float64 x ComplexF64 (DMRG.jl:60 x Ising.jl:17), SubArray views
    (compress.jl:46-58), extent-1 indices (MPS.jl:173-177)."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(5)
    a = np.arange(1, 65).reshape(4, 4, 4)
    b = rng.standard_normal((4, 4))
    check(tb, a, ("i", "j", "k"), b, ("k", "l"))
    z = rand(rng, (4, 4), np.complex128)
    check(tb, b, ("k", "l"), z, ("l", "m"))
    big = rand(rng, (6, 5, 8), np.complex128)
    t = tb.Tensor(big, ("l", "p", "r"))
    t.device()
    v = t.view(("r", slice(2, 6)), ("p", 1))
    w = rand(rng, (4, 3), np.complex128)
    c = tb.binary_einsum(v, tb.Tensor(w, ("r", "x")))
    ref = np.einsum("lr,rx->lx", big[:, 1, 2:6], w)
    assert np.abs(c.parent - ref).max() < 1e-12
    one = rand(rng, (1, 2, 1), np.complex128)
    check(tb, one, ("a", "p", "b"), rand(rng, (1, 3), np.complex128), ("b", "c"))


def test_high_rank_small_extents(ctx):
    """circuit-like operands: many modes of extent 2, interleaved free/contracted modes, custom out order."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(6)
    ai = tuple("abcdefghij")
    bi = tuple("jxhyfzdw")
    a = rand(rng, (2,) * 10, np.complex64)
    b = rand(rng, (2,) * 8, np.complex64)
    check(tb, a, ai, b, bi)
    out = tuple("zaxcywegbi")
    check(tb, a, ai, b, bi, out=out)
    a3 = rand(rng, (3, 2, 4, 2, 3), np.complex128)
    b3 = rand(rng, (4, 3, 5, 2), np.complex128)
    check(tb, a3, tuple("abcde"), b3, tuple("caxd"), out=tuple("bxe"))


def test_sum_single_operand_mode(ctx):
    """an index carried by one operand only and listed in dims is summed out."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(7)
    a = rand(rng, (4, 5, 3), np.complex128)
    b = rand(rng, (3, 6), np.complex128)
    check(tb, a, ("i", "j", "k"), b, ("k", "l"), dims=("k", "j"))


def test_errors(ctx):
    import tenet_jl_b200 as tb
    a = tb.Tensor(np.ones((2, 3)), ("i", "j"))
    b = tb.Tensor(np.ones((4, 3)), ("i", "k"))
    with pytest.raises(ValueError):
        tb.binary_einsum(a, b)
    with pytest.raises(ValueError):
        tb.binary_einsum(a, tb.Tensor(np.ones((2, 5)), ("i", "k")), out=("j",))
    with pytest.raises(TypeError):
        tb.Tensor(np.ones((2,), dtype=object), ("i",)).device()
