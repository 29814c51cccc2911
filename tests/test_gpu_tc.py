"""GPU parity of the specialised kernels against the oracle AND against the generic table kernel:
  * c64_tf32x3 (tcgen05 + TMEM, 3xTF32 split): stated bound — max-abs error relative to the largest |C| entry
    below 2e-5 for K <= 8192 with O(1) Gaussian entries, and within 20x of the exact-FP32 SIMT kernel's own error.
  * thin reduction kernel (M,N <= 4, huge K) and split-K path."""
import numpy as np
import pytest

from tolerances import C128_BOUND, C64_STEP_BOUND

pytestmark = pytest.mark.gpu


def crand(rng, shape, dt=np.complex64):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dt)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 128, 40), (128, 128, 16), (384, 512, 200), (512, 512, 1024),
                                   (256, 1024, 8192),
                                   (200, 130, 77), (1000, 34, 300), (64, 4096, 16384), (130, 258, 2050),   # ragged / split-K
                                   (2560, 1280, 40), (2500, 1300, 72), (4096, 1024, 264)])  # > 148 tiles: persistent CTAs, odd k-block counts
def test_tc_matches_oracle(ctx, M, N, K):
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(M + N + K)
    a, b = crand(rng, (M, K)), crand(rng, (N, K))
    c = tb.binary_einsum(tb.Tensor(a, ("m", "k")), tb.Tensor(b, ("n", "k")))
    assert ctx.last_kernel == "c64_tf32x3"
    ref = a.astype(np.complex128) @ b.astype(np.complex128).T
    got = c.parent
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < C64_STEP_BOUND, f"rel err {err:.2e}"
    # same step on the exact-FP32 generic kernel: both must sit within FP32 noise of the c128 truth
    ctx.set_option(tb._lib.TNB_OPT_FORCE_KERNEL, 1)
    try:
        c2 = tb.binary_einsum(tb.Tensor(a, ("m", "k")), tb.Tensor(b, ("n", "k")))
        assert ctx.last_kernel in ("generic", "splitk")
    finally:
        ctx.set_option(tb._lib.TNB_OPT_FORCE_KERNEL, 0)
    err2 = np.abs(c2.parent - ref).max() / np.abs(ref).max()
    assert err2 < C64_STEP_BOUND
    assert err < 20 * max(err2, 1e-7), f"3xTF32 error {err:.2e} vs FP32-SIMT error {err2:.2e}"


@pytest.mark.parametrize("M,N,K", [(512, 256, 256), (256, 128, 40), (384, 512, 200), (2500, 1300, 72), (258, 130, 2050),
                                   (4096, 1024, 264), (1024, 256, 8192), (640, 64, 136), (8192, 4096, 128)])
def test_pair_kernel_bit_identical_to_single_cta(ctx, M, N, K):
    """The CTA-pair GEMM kernel (cta_group::2, dedicated drain warps) issues the same MMAs per output element in the
    same order and chunks K the same way as the 1-CTA kernel: results must agree bit for bit (ragged edges, split-K,
    conj, beta = 1 accumulation included)."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    a, b = crand(rng, (M, K)), crand(rng, (N, K))
    ta, tb_ = tb.Tensor(a, ("m", "k")).conj(), tb.Tensor(b, ("n", "k"))
    out = {}
    for pair in (0, 1):
        ctx.set_option(tb._lib.TNB_OPT_GEMM_PAIR, pair)
        try:
            out[pair] = tb.binary_einsum(ta, tb_).parent.copy()
            assert ctx.last_kernel == "c64_tf32x3"
        finally:
            ctx.set_option(tb._lib.TNB_OPT_GEMM_PAIR, 1)
    ref = np.conj(a).astype(np.complex128) @ b.astype(np.complex128).T
    assert np.abs(out[1] - ref).max() / np.abs(ref).max() < C64_STEP_BOUND
    assert np.array_equal(out[0], out[1]), f"max diff {np.abs(out[0] - out[1]).max():.3e}"


def test_tc_conj_swap_and_scatter(ctx):
    """conj flags, operand swap (M not a multiple of 128 but N is), multi-mode free/contracted groups, custom
    output order (scattered epilogue)."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(1)
    a = crand(rng, (4, 8, 4, 16, 8))          # m1 m2 m3 | k1 k2   -> M = 128, K = 128
    b = crand(rng, (16, 4, 4, 16, 8))         # n1 n2 n3 | k1 k2   -> N = 256
    ta = tb.Tensor(a, ("m1", "m2", "m3", "k1", "k2"))
    tb_ = tb.Tensor(b, ("n1", "n2", "n3", "k1", "k2")).conj()
    ref = np.einsum("abcxy,defxy->abcdef", a.astype(np.complex128), np.conj(b).astype(np.complex128))
    for out in [None, ("n2", "m1", "n1", "m3", "m2", "n3")]:
        c = tb.binary_einsum(ta, tb_, out=out)
        assert ctx.last_kernel == "c64_tf32x3"
        r = ref if out is None else np.einsum("abcdef->eadcbf", ref)
        assert np.abs(c.parent - r).max() / np.abs(r).max() < C64_STEP_BOUND
    a2, b2 = crand(rng, (64, 32)), crand(rng, (256, 32))     # M = 64 -> swapped roles
    c = tb.binary_einsum(tb.Tensor(a2, ("m", "k")).conj(), tb.Tensor(b2, ("n", "k")))
    ref = np.conj(a2).astype(np.complex128) @ b2.astype(np.complex128).T
    assert np.abs(c.parent - ref).max() / np.abs(ref).max() < C64_STEP_BOUND


def test_simt_mode_option(ctx):
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(2)
    a, b = crand(rng, (128, 64)), crand(rng, (256, 64))
    ctx.set_option(tb._lib.TNB_OPT_C64_MODE, tb._lib.TNB_C64_SIMT)
    try:
        tb.binary_einsum(tb.Tensor(a, ("m", "k")), tb.Tensor(b, ("n", "k")))
        assert ctx.last_kernel == "generic"
    finally:
        ctx.set_option(tb._lib.TNB_OPT_C64_MODE, tb._lib.TNB_C64_TF32X3)


@pytest.mark.parametrize("dt,tol", [(np.complex64, C64_STEP_BOUND), (np.complex128, C128_BOUND), (np.float64, C128_BOUND)])
@pytest.mark.parametrize("M,N", [(1, 1), (2, 3), (4, 4), (1, 4)])
def test_thin_reduction(ctx, dt, tol, M, N):
    """amplitude-closing dot products: M*N <= 16, K = 2^20 (+ a ragged K)."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(3)
    for K in (1 << 20, 777777):
        if np.dtype(dt).kind == "c":
            a, b = crand(rng, (K, M), dt), crand(rng, (K, N), dt)
        else:
            a, b = rng.standard_normal((K, M)).astype(dt), rng.standard_normal((K, N)).astype(dt)
        c = tb.binary_einsum(tb.Tensor(a, ("k", "m")), tb.Tensor(b, ("k", "n")))
        assert ctx.last_kernel == "stream"
        ref = a.astype(np.complex128).T @ b.astype(np.complex128)
        assert np.abs(c.parent - ref).max() / np.abs(ref).max() < tol * 10


@pytest.mark.parametrize("dt,tol", [(np.complex64, C64_STEP_BOUND), (np.complex128, C128_BOUND)])
@pytest.mark.parametrize("M,N", [(8, 32), (2, 4), (16, 32), (4, 128), (6, 12), (32, 16), (16, 16), (16, 8)])
def test_k_reduction_kernel(ctx, dt, tol, M, N):
    """small M x N, huge K, dense operands with the free index fastest (the environment-closing steps of a sliced
    path; storage is column-major like Julia): HBM-bound k-reduction kernel; ragged K, conj flags, high-rank groups.
    With TNB_KRED_MMA=1 in the environment the complex64 cases with M % 16 == 0, N % 8 == 0 run on the mma.sync
    (3xTF32) variant of the kernel instead of the FP32-FMA one."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(M * 7 + N)
    hi = np.complex128
    for K in (1 << 18, 300007):
        a, b = crand(rng, (M, K), dt), crand(rng, (N, K), dt)
        c = tb.binary_einsum(tb.Tensor(a, ("m", "k")), tb.Tensor(b, ("n", "k")))
        assert ctx.last_kernel == "stream", ctx.last_kernel
        ref = a.astype(hi) @ b.astype(hi).T
        assert np.abs(c.parent - ref).max() / np.abs(ref).max() < tol * 10
    c = tb.binary_einsum(tb.Tensor(a, ("m", "k")).conj(), tb.Tensor(b, ("n", "k")), out=("n", "m"))
    assert ctx.last_kernel == "stream"
    ref = (np.conj(a).astype(hi) @ b.astype(hi).T).T
    assert np.abs(c.parent - ref).max() / np.abs(ref).max() < tol * 10
    if M == 8:
        a4 = a[:, : 1 << 18].reshape(2, 2, 2, 64, 64, 64)
        b4 = b[:, : 1 << 18].reshape(4, 8, 64, 64, 64)
        c = tb.binary_einsum(tb.Tensor(a4, ("m0", "m1", "m2", "k0", "k1", "k2")), tb.Tensor(b4, ("n0", "n1", "k0", "k1", "k2")))
        assert ctx.last_kernel == "stream"
        ref = np.einsum("xyzabc,uvabc->xyzuv", a4.astype(hi), b4.astype(hi))
        got = np.transpose(np.asarray(c.parent), [list(c.inds).index(i) for i in ("m0", "m1", "m2", "n0", "n1")])
        assert np.abs(got - ref).max() / np.abs(ref).max() < tol * 10


@pytest.mark.parametrize("M,N,K", [(128, 64, 8), (33, 47, 5), (200, 130, 77), (64, 512, 256), (1000, 8, 300), (256, 256, 4096)])
def test_c128_dmma_matches_oracle(ctx, M, N, K):
    """FP64 tensor-core kernel: ragged edges, operand swap (N > M), split-K, against c128 numpy (1e-12)."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    a, b = crand(rng, (M, K), np.complex128), crand(rng, (K, N), np.complex128)
    c = tb.binary_einsum(tb.Tensor(a, ("m", "k")), tb.Tensor(b, ("k", "n")))
    if M * N >= 1024 and M * N * K >= 65536:
        assert ctx.last_kernel == "c128_dmma"
    ref = a @ b
    assert np.abs(c.parent - ref).max() / np.abs(ref).max() < 1e-12
    c2 = tb.binary_einsum(tb.Tensor(a, ("m", "k")).conj(), tb.Tensor(b.T.copy(), ("n", "k")).conj(), out=("n", "m"))
    assert np.abs(c2.parent - np.conj(ref).T).max() / np.abs(ref).max() < 1e-12


def test_c128_dmma_batch_and_high_rank(ctx):
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(11)
    a = crand(rng, (3, 4, 5, 16, 7), np.complex128)     # b x y k z
    b = crand(rng, (16, 3, 8, 9), np.complex128)        # k b u v
    c = tb.binary_einsum(tb.Tensor(a, "bxykz"), tb.Tensor(b, "kbuv"), dims=("k",))
    ref = np.einsum("bxykz,kbuv->xyzuvb", a, b)
    assert c.inds == tuple("xyzuvb")
    assert ctx.last_kernel == "c128_dmma"
    assert np.abs(c.parent - ref).max() / np.abs(ref).max() < 1e-12


def test_tc_misaligned_view_falls_back(ctx):
    """a view whose first element is only 8-byte aligned cannot be bulk-copied: the step runs on the exact-FP32
    generic kernel instead (same result, no error)."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(21)
    big = crand(rng, (130, 64))
    b = crand(rng, (256, 64))
    t = tb.Tensor(big, ("m", "k"))
    t.device()
    v = t.view(("m", slice(1, 129)))
    c = tb.binary_einsum(v, tb.Tensor(b, ("n", "k")))
    ref = big[1:129].astype(np.complex128) @ b.astype(np.complex128).T
    assert np.abs(c.parent - ref).max() / np.abs(ref).max() < C64_STEP_BOUND


@pytest.mark.parametrize("dt,tol", [(np.complex64, 2e-6), (np.complex128, 1e-13)])
def test_stem_stream_kernel(ctx, dt, tol):
    """HBM-bound stem steps: huge dense operand x tiny operand, consumer layouts that interleave the tiny
    operand's modes with the big one's (sorted tile pattern), both orientations, conj flags."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(31)
    big_modes = [f"m{i}" for i in range(17)]
    a = crand(rng, (2,) * 17 + (2, 2), dt)                       # m0..m16 | k0 k1   -> M = 2^17, K = 4
    b = crand(rng, (2, 2, 2, 2, 2), dt)                          # n0 n1 n2 | k0 k1  -> N = 8
    ta = tb.Tensor(a, big_modes + ["k0", "k1"])
    tb_ = tb.Tensor(b, ["n0", "n1", "n2", "k0", "k1"])
    hi = np.complex128
    ref = np.einsum(a.astype(hi), list(range(19)), b.astype(hi), [19, 20, 21, 17, 18], list(range(17)) + [19, 20, 21])
    ref_inds = big_modes + ["n0", "n1", "n2"]
    outs = [None,
            ["n0"] + big_modes[:3] + ["n1"] + big_modes[3:9] + ["n2"] + big_modes[9:],       # tiny modes interleaved
            big_modes[5:] + ["n2", "n1", "n0"] + big_modes[:5]]                               # big modes reordered too
    for out in outs:
        c = tb.binary_einsum(ta, tb_, out=out)
        assert ctx.last_kernel == "stem", ctx.last_kernel
        r = ref if out is None else np.transpose(ref, [ref_inds.index(i) for i in out])
        assert np.abs(c.parent - r).max() / np.abs(r).max() < tol
    # swapped orientation + conj on both
    c = tb.binary_einsum(tb_.conj(), ta.conj())
    assert ctx.last_kernel == "stem"
    r = np.conj(np.transpose(ref, [17, 18, 19] + list(range(17))))
    assert np.abs(c.parent - r).max() / np.abs(r).max() < tol
    # K = 1 (outer product with a tiny tensor) and N = 16, K = 16
    a2 = crand(rng, (1 << 16,), dt)
    b2 = crand(rng, (4,), dt)
    c = tb.binary_einsum(tb.Tensor(a2, ["m"]), tb.Tensor(b2, ["n"]))
    assert np.abs(c.parent - np.outer(a2.astype(hi), b2.astype(hi))).max() < tol * 10
    a3, b3 = crand(rng, (1 << 16, 16), dt), crand(rng, (8, 16), dt)
    c = tb.binary_einsum(tb.Tensor(a3, ["m", "k"]), tb.Tensor(b3, ["n", "k"]), out=["n", "m"])
    assert ctx.last_kernel == "stem"
    r = (a3.astype(hi) @ b3.astype(hi).T).T
    assert np.abs(c.parent - r).max() / np.abs(r).max() < tol * 4


@pytest.mark.parametrize("dt,tol", [(np.complex64, 2e-6), (np.complex128, 1e-13)])
@pytest.mark.parametrize("N,K", [(4, 4), (6, 6), (8, 2), (16, 16)])
def test_stem_direct_form_bit_identical(ctx, monkeypatch, dt, tol, N, K):
    """SIMT stem kernel: rows stored straight from registers where the planner finds whole 64-byte pieces per warp store
    (st_direct; 16-byte stores of two complex64 rows when they are adjacent), staged sorted write-out otherwise and under
    TNB_STEM_DIRECT=0.  Same FMAs in the same order: bit-identical; non-power-of-two small operand (the 6 x 6 MPO step of
    configs[4]), conj, beta = 1."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(N * 11 + K)
    a = crand(rng, (2,) * 17 + (K,), dt)
    b = crand(rng, (N, K), dt)
    big = [f"m{i}" for i in range(17)]
    ta, tb_ = tb.Tensor(a, big + ["k"]).conj(), tb.Tensor(b, ["n", "k"])
    ref = np.tensordot(np.conj(a).astype(np.complex128), b.astype(np.complex128), axes=([17], [1]))
    outs = [None, big[:3] + ["n"] + big[3:], big[:5] + ["n"] + big[5:], big[1:] + ["n"] + big[:1],   # direct
            big[:1] + ["n"] + big[1:], ["n"] + big]                                                   # staged
    kern = "stem_tc" if (dt == np.complex64 and N % 16 == 0 and K % 8 == 0) else "stem"              # 16 x 16 complex64: tensor cores
    for out in outs:
        res = {}
        for direct in ("1", "0"):
            monkeypatch.setenv("TNB_STEM_DIRECT", direct)
            res[direct] = tb.binary_einsum(ta, tb_, out=out).parent.copy()
            assert ctx.last_kernel == kern, ctx.last_kernel
        monkeypatch.delenv("TNB_STEM_DIRECT")
        r = ref if out is None else np.transpose(ref, [(big + ["n"]).index(i) for i in out])
        assert np.abs(res["1"] - r).max() / np.abs(r).max() < (tol if kern == "stem" else C64_STEP_BOUND)
        assert np.array_equal(res["0"], res["1"]), f"out={out}: max diff {np.abs(res['0'] - res['1']).max():.3e}"
    tn = tb.TensorNetwork([ta, tb_])
    out = big + ["n"]
    plan = tb.ContractionPlan(tn, tb.einexpr(tn, output=out), output=out, ctx=ctx)
    assert plan.step_info(0)["kernel_name"] == kern
    plan.zero_output()
    plan.execute(accumulate=True)
    once = plan.result().parent.copy()
    plan.execute(accumulate=True)
    twice = plan.result().parent.copy()
    plan.close()
    assert np.abs(once - ref).max() / np.abs(ref).max() < (tol if kern == "stem" else C64_STEP_BOUND)
    assert np.array_equal(twice, once + once)


@pytest.mark.parametrize("N,K", [(16, 16), (32, 128), (64, 64), (48, 8), (64, 16), (128, 64), (256, 16), (64, 128), (128, 128), (256, 96)])
def test_stem_tc_kernel(ctx, N, K):
    """persistent tcgen05 stem kernel: huge x small, several consumer layouts, both orientations."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(N * 131 + K)
    M = 1 << 17
    a = crand(rng, (2,) * 17 + (K,))
    b = crand(rng, (N, K))
    big = [f"m{i}" for i in range(17)]
    ta, tb_ = tb.Tensor(a, big + ["k"]), tb.Tensor(b, ["n", "k"])
    hi = np.complex128
    ref = np.tensordot(a.astype(hi), b.astype(hi), axes=([17], [1]))          # m0..m16, n
    # last layout: the fastest output mode is neither in the tile nor in the small operand -> runs of one element
    for out in [None, big[:2] + ["n"] + big[2:], ["n"] + big[8:] + big[:8], big[8:] + ["n"] + big[:8]]:
        c = tb.binary_einsum(ta, tb_, out=out)
        assert ctx.last_kernel == "stem_tc", ctx.last_kernel
        r = ref if out is None else np.transpose(ref, [(big + ["n"]).index(i) for i in out])
        err = np.abs(c.parent - r).max() / np.abs(r).max()
        assert err < C64_STEP_BOUND, err
    c = tb.binary_einsum(tb_.conj(), ta)
    assert ctx.last_kernel == "stem_tc"
    r = np.transpose(np.tensordot(a.astype(hi), np.conj(b).astype(hi), axes=([17], [1])), [17] + list(range(17)))
    assert np.abs(c.parent - r).max() / np.abs(r).max() < C64_STEP_BOUND


@pytest.mark.parametrize("N,K", [(128, 128), (256, 64), (512, 128), (128, 72)])
def test_wide_stem_on_cta_pairs_bit_identical(ctx, N, K):
    """128-column stem passes with a dense small operand run on the CTA-pair kernel (staged sorted-pattern epilogue);
    with TNB_OPT_GEMM_PAIR = 0 the 1-CTA stem kernel takes them.  Same MMA order per output element: bit-identical,
    for every consumer layout (coalesced write-out through the planner's sorted tile pattern), both orientations,
    conj, beta = 1."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(N * 17 + K)
    a = crand(rng, (2,) * 17 + (K,))
    b = crand(rng, (N, K))
    big = [f"m{i}" for i in range(17)]
    ta, tb_ = tb.Tensor(a, big + ["k"]).conj(), tb.Tensor(b, ["n", "k"])
    ref = np.tensordot(np.conj(a).astype(np.complex128), b.astype(np.complex128), axes=([17], [1]))
    l0 = ctx.launch_count
    for out in [None, big[:2] + ["n"] + big[2:], ["n"] + big[8:] + big[:8], big[8:] + ["n"] + big[:8]]:
        res = {}
        for pair in (0, 1):
            ctx.set_option(tb._lib.TNB_OPT_GEMM_PAIR, pair)
            try:
                res[pair] = tb.binary_einsum(ta, tb_, out=out).parent.copy()
                assert ctx.last_kernel == "stem_tc", ctx.last_kernel
            finally:
                ctx.set_option(tb._lib.TNB_OPT_GEMM_PAIR, 1)
        r = ref if out is None else np.transpose(ref, [(big + ["n"]).index(i) for i in out])
        assert np.abs(res[1] - r).max() / np.abs(r).max() < C64_STEP_BOUND
        assert np.array_equal(res[0], res[1]), f"out={out}: max diff {np.abs(res[0] - res[1]).max():.3e}"
    c0 = tb.binary_einsum(tb_, ta)                     # swapped orientation
    assert ctx.last_kernel == "stem_tc"
    r = np.transpose(ref, [17] + list(range(17)))
    assert np.abs(c0.parent - r).max() / np.abs(r).max() < C64_STEP_BOUND


@pytest.mark.parametrize("N,K", [(16, 32), (32, 32), (32, 128), (64, 16), (64, 64), (48, 24), (128, 64), (256, 128), (128, 256)])
def test_stem_tc_direct_epilogue_bit_identical(ctx, monkeypatch, N, K):
    """Stem passes (1-CTA kernel: <= 64 columns, CTA-pair kernel: 128 columns) store rows straight from registers when a
    warp's 32 rows of one column are whole 64-byte pieces in the output (planner: st_direct); TNB_STEM_DIRECT=0 at plan
    time keeps the staged, sorted write-out.  Same MMAs, same order: bit-identical, for layouts on both sides of the
    criterion, with conj and beta = 1 accumulation."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(N * 7 + K)
    a = crand(rng, (2,) * 17 + (K,))
    b = crand(rng, (N, K))
    big = [f"m{i}" for i in range(17)]
    ta, tb_ = tb.Tensor(a, big + ["k"]), tb.Tensor(b, ["n", "k"]).conj()
    ref = np.tensordot(a.astype(np.complex128), np.conj(b).astype(np.complex128), axes=([17], [1]))
    # m-fastest (direct), 8 rows then n (direct, 4 pieces per store), 4 rows then n (staged), n fastest (staged)
    # ... and the fastest row as the output's slowest index (direct: two 128-byte pieces per store, tile bases far apart)
    for out in [None, big[:3] + ["n"] + big[3:], big[:5] + ["n"] + big[5:], big[:2] + ["n"] + big[2:], ["n"] + big,
                big[1:] + ["n"] + big[:1]]:
        res = {}
        for direct in ("1", "0"):
            monkeypatch.setenv("TNB_STEM_DIRECT", direct)
            res[direct] = tb.binary_einsum(ta, tb_, out=out).parent.copy()
            assert ctx.last_kernel == "stem_tc", ctx.last_kernel
        monkeypatch.delenv("TNB_STEM_DIRECT")
        r = ref if out is None else np.transpose(ref, [(big + ["n"]).index(i) for i in out])
        assert np.abs(res["1"] - r).max() / np.abs(r).max() < C64_STEP_BOUND
        assert np.array_equal(res["0"], res["1"]), f"out={out}: max diff {np.abs(res['0'] - res['1']).max():.3e}"
    # beta = 1 (slice accumulation into the output): executing the one-step plan twice doubles every element exactly
    tn = tb.TensorNetwork([ta, tb_])
    for out in [big + ["n"], big[:3] + ["n"] + big[3:]]:
        path = tb.einexpr(tn, output=out)
        plan = tb.ContractionPlan(tn, path, output=out, ctx=ctx)
        assert plan.step_info(0)["kernel_name"] == "stem_tc"
        plan.zero_output()
        plan.execute(accumulate=True)
        once = plan.result().parent.copy()
        plan.execute(accumulate=True)
        twice = plan.result().parent.copy()
        plan.close()
        r = np.transpose(ref, [(big + ["n"]).index(i) for i in out])
        assert np.abs(once - r).max() / np.abs(r).max() < C64_STEP_BOUND
        assert np.array_equal(twice, once + once)


def test_stem_tc_rank_table_path(ctx, monkeypatch):
    """the general (non-separable rank) epilogue of the stem kernel: forced through TNB_STEM_NO_ADDITIVE."""
    import tenet_jl_b200 as tb
    monkeypatch.setenv("TNB_STEM_NO_ADDITIVE", "1")
    rng = np.random.default_rng(5)
    a = crand(rng, (2,) * 17 + (32,))
    b = crand(rng, (32, 32))
    big = [f"m{i}" for i in range(17)]
    ta, tb_ = tb.Tensor(a, big + ["k"]), tb.Tensor(b, ["n", "k"])
    hi = np.complex128
    ref = np.tensordot(a.astype(hi), b.astype(hi), axes=([17], [1]))
    for out in [big[:3] + ["n"] + big[3:], big[9:] + ["n"] + big[:9]]:
        c = tb.binary_einsum(ta, tb_, out=out)
        assert ctx.last_kernel == "stem_tc", ctx.last_kernel
        r = np.transpose(ref, [(big + ["n"]).index(i) for i in out])
        err = np.abs(c.parent - r).max() / np.abs(r).max()
        assert err < C64_STEP_BOUND, err


@pytest.mark.parametrize("N,K", [(128, 256), (128, 512), (256, 512), (128, 200)])
def test_stem_tc_long_k(ctx, N, K):
    """huge x small with K > 128: a small operand of exactly 128 columns goes through the 128-column stem pass
    (K <= 512, chunks of 128 k summed round-to-nearest); wider ones take the tile GEMM kernel unless TNB_STEM_KMAX=512
    is in the environment.  Both accumulate chunks round-to-nearest, so the bound is the same."""
    import os
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(N + K)
    a = crand(rng, (2,) * 17 + (K,))
    b = crand(rng, (N, K))
    big = [f"m{i}" for i in range(17)]
    ta, tb_ = tb.Tensor(a, big + ["k"]), tb.Tensor(b, ["n", "k"])
    hi = np.complex128
    ref = np.tensordot(a.astype(hi), b.astype(hi), axes=([17], [1]))
    kmax = int(os.environ.get("TNB_STEM_KMAX", "128"))
    if N == 128:
        kmax = max(kmax, 512)
    for out in [None, big[:2] + ["n"] + big[2:], ["n"] + big[8:] + big[:8]]:
        c = tb.binary_einsum(ta, tb_, out=out)
        assert ctx.last_kernel == ("stem_tc" if K <= kmax and K % 8 == 0 else "c64_tf32x3"), ctx.last_kernel
        r = ref if out is None else np.transpose(ref, [(big + ["n"]).index(i) for i in out])
        err = np.abs(c.parent - r).max() / np.abs(r).max()
        assert err < C64_STEP_BOUND, err
