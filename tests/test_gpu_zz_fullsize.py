"""GPU parity at BASELINE.json's FULL sizes, where no CPU oracle finishes: size-independent properties only.

* linearity over slice ranges (the multi-GPU partition), run-to-run determinism;
* homogeneity: a leaf scaled by 2 doubles the result to within 2 ulp (a power-of-two scale commutes with every
  rounding on the path, including the hi/lo TF32 split of the tcgen05 kernels, so it is normally bit-exact — printed);
* the numpy oracle itself on oracle-sized sub-slices of the full-size network (extra sliced indices);
* cross-kernel: the specialised kernels (tcgen05 3xTF32, stem, k-reduction, DMMA) against the generic table kernel
  (`TNB_OPT_FORCE_KERNEL`, plain FP32/FP64 FMA — the BLAS-equivalent arithmetic) on the same slice;
* cross-path: two different contraction trees / slicings of the same network give the same number;
* the reference-derived known answers K1 (<psi|psi> = 1, MPS.jl:103-104,154-157) and K2 via K4
  (<0..0|H|0..0> = -J(n-1), Ising.jl:12-30) at configs[0] / configs[4] sizes.

The file sorts last on purpose: these are the most expensive GPU tests (tens of seconds, tens of GiB).
"""
import ctypes as C

import numpy as np
import pytest

from tolerances import C128_BOUND, C64_PATH_BOUND

pytestmark = pytest.mark.gpu


def _run(plan, b, s, e):
    plan.zero_output()
    plan.execute(b, s, e, accumulate=True)
    return np.asarray(plan.result().parent).reshape(-1).copy()


def _scale_leaf(tb, ctx, t, factor, dtype):
    """overwrite the device copy of leaf `t` with factor * (its host values)."""
    arr = t._dev
    host = np.asfortranarray(arr.to_numpy() * factor).astype(dtype)
    flat = np.ascontiguousarray(host.reshape(-1, order="F"))
    tb._lib.check(ctx.handle, ctx.lib.tnb_upload(ctx.handle, arr.buffer.handle, arr.offset * arr.dtype.itemsize,
                                                 flat.ctypes.data_as(C.c_void_p), flat.nbytes))
    ctx.sync()          # the staging array may go away after this


def _sliced_properties(tb, ctx, name, cross_tol):
    import bench
    tn, path = bench.build_workload(tb, name)
    dt = np.dtype(np.complex64)
    plan = tb.ContractionPlan(tn, path, ctx=ctx)
    n = plan.nslices
    assert n >= 4
    kernels = {plan.step_info(s)["kernel_name"] for s in range(plan.nsteps)}
    a0 = _run(plan, 0, 1, 1)
    a0b = _run(plan, 0, 1, 1)
    assert np.array_equal(a0, a0b), "same slice, same launches: results must be bit-identical"
    a3 = _run(plan, 3, 1, 4)
    pair = _run(plan, 0, 3, 4)                       # slices {0, 3}
    eps = np.finfo(np.float32).eps
    assert np.abs(pair - (a0 + a3)).max() <= 4 * eps * (np.abs(a0).max() + np.abs(a3).max())
    assert np.abs(a0).max() > 0 and np.all(np.isfinite(a0.view(np.float32)))
    # homogeneity, bit exact
    leaf = tn.tensors[0]
    _scale_leaf(tb, ctx, leaf, 2.0, dt)
    d0 = _run(plan, 0, 1, 1)
    _scale_leaf(tb, ctx, leaf, 0.5, dt)
    exact = bool(np.array_equal(d0, 2 * a0))
    assert np.abs(d0 - 2 * a0).max() <= 2 * eps * np.abs(2 * a0).max(), (d0, a0)
    plan.close()
    # cross-kernel on slice 0
    ctx.set_option(tb._lib.TNB_OPT_FORCE_KERNEL, 1)
    try:
        gplan = tb.ContractionPlan(tn, path, ctx=ctx)
        assert {gplan.step_info(s)["kernel_name"] for s in range(gplan.nsteps)} <= {"generic", "splitk"}
        g0 = _run(gplan, 0, 1, 1)
        gplan.close()
    finally:
        ctx.set_option(tb._lib.TNB_OPT_FORCE_KERNEL, 0)
    rel = np.abs(a0 - g0).max() / max(np.abs(g0).max(), np.abs(a3).max())
    print(f"{name}: slice 0 = {a0[0]:.6e}; x2 leaf scaling bit-exact: {exact}; specialised kernels {sorted(kernels)} "
          f"vs generic FP32: {rel:.2e} of max(|slice 0|, |slice 3|)")
    assert rel < cross_tol, rel
    return kernels


def test_sycamore53_m14_full_size_properties(ctx):
    """configs[2] at full size on the committed path (256 slices x 2^39.9 MACs, 8 GiB peak intermediate).
    Cross-kernel tolerance: the stated complex64 path bound (tests/tolerances.py), relative to max(|slice 0|, |slice 3|);
    measured 1.4e-5 (a slice amplitude is a cancelling sum of ~2^40 products and both sides carry FP32 rounding)."""
    import tenet_jl_b200 as tb
    kernels = _sliced_properties(tb, ctx, "sycamore53_m14", C64_PATH_BOUND)
    assert {"c64_tf32x3", "stem_tc"} <= kernels
    ctx.trim()


def test_regular3_n100_full_size_properties(ctx):
    """configs[1] (largest instance whose best found path fits one GPU: 100 tensors, bond 4, 16 slices)."""
    import tenet_jl_b200 as tb
    _sliced_properties(tb, ctx, "regular3_n100_d4", C64_PATH_BOUND)
    ctx.trim()


@pytest.mark.parametrize("name", ["sycamore53_m14", "regular3_n100_d4"])
def test_full_size_subslices_vs_oracle(ctx, name):
    """The full-size network against the numpy complex128 oracle on the pieces a CPU can finish: the committed path
    with extra sliced indices (bench.subslice_path, <= 2^33 MACs per sub-slice; this is also bench.py's cpu_baseline
    sample), sub-slices first / second / middle / last.  Error relative to the largest of the reference values (a
    single sub-slice can be atypically small): the stated complex64 path bound (3xTF32 + FP32 accumulation along ~250
    steps; measured 2.0e-5)."""
    import bench
    import tenet_jl_b200 as tb
    from oracle import einsum_oracle as orc
    tn, path = bench.build_workload(tb, name)
    p = bench.subslice_path(tb, tn, path, 33.0)
    assert tuple(p.sliced)[:len(path.sliced)] == tuple(path.sliced) and p.log2_macs <= 33.0
    arrays = [t.parent.astype(np.complex128) for t in tn.tensors]
    inds = [t.inds for t in tn.tensors]
    plan = tb.ContractionPlan(tn, p, ctx=ctx)
    assert plan.nslices == p.nslices
    ids = sorted({0, 1, p.nslices // 2 + 1, p.nslices - 1})
    got, ref = [], []
    for i in ids:
        got.append(complex(_run(plan, i, 1, i + 1)[0]))
        r, _ = orc.contract_sliced(arrays, inds, p.steps, list(p.sliced), slice_ids=[i])
        ref.append(complex(r))
    plan.close()
    scale = max(abs(r) for r in ref)
    errs = [abs(g - r) / scale for g, r in zip(got, ref)]
    print(f"{name}: {p.nslices} sub-slices of 2^{p.log2_macs:.1f} MACs; ids {ids}; errors / max|ref| = "
          + ", ".join(f"{e:.1e}" for e in errs))
    assert max(errs) < C64_PATH_BOUND, (got, ref)
    ctx.trim()


def test_sycamore53_m14_full_slices_c64_vs_c128(ctx):
    """FULL slices (2^39.9 MACs each) of the committed path: the complex64 engine (tcgen05 3xTF32 + stem + k-reduction
    kernels) against the SAME engine run in complex128 (FP64 DMMA / DFMA kernels, ~1e-12) on the same slices — the c128
    truth no CPU can produce at this size.  Slices first / second / middle / last and their sum, errors relative to the
    largest |c128 partial|: the stated complex64 path bound."""
    import bench
    import tenet_jl_b200 as tb
    tn, path = bench.build_workload(tb, "sycamore53_m14")
    ids = [0, 1, 129, 255]
    p64 = tb.ContractionPlan(tn, path, ctx=ctx)
    got = [complex(_run(p64, i, 1, i + 1)[0]) for i in ids]
    p64.close()
    ctx.trim()
    p128 = tb.ContractionPlan(tn, path, ctx=ctx, dtype=np.complex128)
    assert "c128_dmma" in {p128.step_info(s)["kernel_name"] for s in range(p128.nsteps)}
    ref = [complex(_run(p128, i, 1, i + 1)[0]) for i in ids]
    p128.close()
    del tn
    ctx.trim()
    scale = max(abs(r) for r in ref)
    errs = [abs(g - r) / scale for g, r in zip(got, ref)]
    esum = abs(sum(got) - sum(ref)) / max(abs(sum(ref)), scale)
    print("sycamore53_m14 full slices " + str(ids) + ": |c64 - c128| / max|c128| = " + ", ".join(f"{e:.1e}" for e in errs)
          + f"; sum of the four: {esum:.1e}")
    assert max(errs) < C64_PATH_BOUND and esum < C64_PATH_BOUND, (got, ref)


def test_peps6x6_d4_two_paths_agree(ctx):
    """configs[3] at full size, complex128: the unsliced row-by-row boundary path and the committed sliced
    hyper-optimised path are different trees over the same 72 tensors; the norm is real and positive."""
    import bench
    import tenet_jl_b200 as tb
    tn, pb = bench.build_workload(tb, "peps6x6_d4_boundary")
    vb = tb.contract(tn, path=pb, ctx=ctx).item()
    tn2, ps = bench.build_workload(tb, "peps6x6_d4")
    assert ps.nslices > 1
    vs = tb.contract(tn2, path=ps, ctx=ctx).item()
    assert vb.real > 0 and abs(vb.imag) < 1e-11 * vb.real
    assert abs(vs - vb) / abs(vb) < 100 * C128_BOUND, (vs, vb)
    ctx.trim()


def test_mps_norm_full_size_K1(ctx):
    """configs[0] at full size: 32 sites, chi = 128, complex128; rand(MPS) is right-canonical => 1."""
    import tenet_jl_b200 as tb
    tn, psi = tb.workloads.mps_norm_network(32, 128, np.complex128, seed=1)
    v = tb.contract(tn, path=tb.workloads.zipper_path(32), ctx=ctx).item()
    assert abs(v - 1.0) < 1e-12, v


def test_mps_mpo_full_size_K2_via_K4(ctx):
    """configs[4] at full size: 100 sites, chi = 1024, complex128, env sweep (DMRG.jl:106-115).  The state is the
    product state |0..0> padded to chi = 1024 with random bond gauges (dense 1024 x 2 x 1024 sites), so
    <psi|H|psi> / <psi|psi> = -J (n-1) exactly (Ising.jl:12-30)."""
    import tenet_jl_b200 as tb
    n, J, h = 100, 1.3, 0.7
    H = tb.ising_1d_mpo(n, h, J)
    psi = tb.workloads.product_mps("0" * n, chi=1024, seed=5)
    assert max(max(t.shape) for t in psi.tensors) == 1024
    plan_tn = tb.expect_network(psi, H)
    v = tb.contract(plan_tn, path=tb.workloads.sweep_path(n), ctx=ctx).item()
    nrm = tb.overlap(psi, psi).item()
    assert abs(nrm - 1.0) < 1e-9, nrm
    assert abs(v / nrm - (-J * (n - 1))) < 1e-8 * J * (n - 1), v
    ctx.trim()
