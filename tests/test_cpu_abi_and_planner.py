"""CPU: the C-ABI library loads, exports every symbol include/tnb200.h declares, refuses to compute without a GPU,
and its PLANNER (dry-run plans: GEMM views, offset tables, intermediate layouts, arena, slice offsets) reproduces the
oracle when replayed with numpy gathers (tests/plan_emulator.py).  No CUDA compute happens here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import einsum_oracle as orc
from plan_emulator import replay

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "tnb200.h")).read()
    declared = set(re.findall(r"\b(tnb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    from tenet_jl_b200 import _lib
    assert declared == set(_lib.ABI_SYMBOLS), declared ^ set(_lib.ABI_SYMBOLS)
    for s in declared:
        assert hasattr(built_lib, s), s


def test_no_gpu_means_loud_error_not_fallback(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import tenet_jl_b200 as tb
    with pytest.raises(tb.TnbError) as e:
        tb.Context(0)
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(tb.TnbError):
        tb.binary_einsum(tb.Tensor(np.ones((2, 2)), "ab"), tb.Tensor(np.ones((2, 2)), "bc"))


def test_product_path_never_imports_oracle():
    pkg = os.path.join(ROOT, "tenet.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("test/oracle", ""), f"{f} mentions the oracle"


def test_default_result_order(built_lib):
    from tenet_jl_b200._lib import make_desc
    a, ka = make_desc(None, 0, 0, [2, 3, 4], [1, 2, 6], [10, 11, 12])
    b, kb = make_desc(None, 0, 0, [4, 5, 3], [1, 4, 20], [12, 13, 11])
    rank = C.c_int32()
    modes = (C.c_int32 * 64)()
    ext = (C.c_int64 * 64)()
    sm = (C.c_int32 * 1)(12)
    assert built_lib.tnb_binary_einsum_result(C.byref(a), C.byref(b), sm, 1, C.byref(rank), modes, ext) == 0
    assert list(modes[:rank.value]) == [10, 13, 11] and list(ext[:rank.value]) == [2, 5, 3]


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("dt", [np.complex128, np.float64])
def test_planner_replay_random_networks(built_lib, seed, dt):
    import tenet_jl_b200 as tb
    tn = tb.workloads.random_regular_network(n=12, bond=3, dtype=dt, seed=seed)
    arrays = [t.parent for t in tn.tensors]
    inds = [t.inds for t in tn.tensors]
    p = tb.einexpr(tn, ntrials=3, seed=seed)
    plan = tb.ContractionPlan(tn, p, dry=True)
    ref, _ = orc.contract_path(arrays, inds, p.steps)
    got = replay(plan, arrays)
    assert abs(got - ref) <= 1e-12 * max(abs(ref), 1e-30)
    # sliced: every slice id exactly once, hoisted steps once
    p2 = tb.einexpr(tn, ntrials=3, seed=seed, max_log2_size=max(2.0, p.log2_max_size - 3.2))
    if p2.nslices > 1 and p2.nslices <= 2000:
        plan2 = tb.ContractionPlan(tn, p2, dry=True)
        got2 = replay(plan2, arrays)
        assert abs(got2 - ref) <= 1e-11 * max(abs(ref), 1e-30)
        half = replay(plan2, arrays, slice_ids=range(0, p2.nslices, 2)) + replay(plan2, arrays, slice_ids=range(1, p2.nslices, 2))
        assert abs(half - ref) <= 1e-11 * max(abs(ref), 1e-30)
        assert plan2.info["nslices"] == p2.nslices


def test_planner_replay_open_indices_conj_and_views(built_lib):
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(9)
    mk = lambda s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    ts = [tb.Tensor(mk((3, 4, 5)), "abc"), tb.Tensor(mk((5, 6, 2)), "cde").conj(), tb.Tensor(mk((4, 6, 7)), "bdf"),
          tb.Tensor(mk((7, 3, 1)), "fgu")]
    tn = tb.TensorNetwork(ts)
    arrays = [t.parent for t in tn.tensors]         # parent applies the conj flag
    raw = [np.conj(a) if t._conj else a for a, t in zip(arrays, ts)]
    inds = [t.inds for t in ts]
    for out in [("a", "e", "g", "u"), ("g", "u", "e", "a")]:
        p = tb.einexpr(tn, output=out, ntrials=2)
        plan = tb.ContractionPlan(tn, p, output=out, dry=True)
        ref, _ = orc.contract_path(arrays, inds, p.steps, output=out)
        got = replay(plan, raw)
        assert np.abs(got - ref).max() < 1e-12 * np.abs(ref).max()


def test_planner_layouts_make_intermediates_dense_matrices(built_lib):
    """the layout rule: every intermediate is consumed as a dense, free-index-fastest matrix (TC-kernel eligible)."""
    import tenet_jl_b200 as tb
    tn, _ = tb.workloads.sycamore_amplitude_network(rows=4, cols=4, cycles=8, seed=1, removed=(), dtype=np.complex64)
    p = tb.einexpr(tn, ntrials=4, seed=0, max_log2_size=10)
    plan = tb.ContractionPlan(tn, p, dry=True)
    lib = plan.lib
    nleaves = len(tn.tensors)
    dense = total = 0
    for s in range(plan.nsteps):
        ids = (C.c_int32 * 3)(); base = (C.c_int64 * 3)(); kind = (C.c_int32 * 3)()
        sst = (C.c_int64 * max(3 * len(p.sliced), 1))(); cj = (C.c_int32 * 2)()
        lib.tnb_plan_dump_step(plan.handle, s, ids, base, kind, sst, cj)
        info = plan.step_info(s)
        for which, (opk, nfree, tab_free, tab_k) in enumerate([(kind[0], info["M"], 0, 1), (kind[1], info["N"], 3, 4)]):
            if opk != 1:
                continue                      # only intermediates (arena operands)
            total += 1
            n = lib.tnb_plan_dump_table(plan.handle, s, tab_free, None, 0)
            buf = (C.c_int64 * max(n, 1))(); lib.tnb_plan_dump_table(plan.handle, s, tab_free, buf, n)
            free = np.frombuffer(buf, dtype=np.int64, count=n)
            nk = lib.tnb_plan_dump_table(plan.handle, s, tab_k, None, 0)
            bk = (C.c_int64 * max(nk, 1))(); lib.tnb_plan_dump_table(plan.handle, s, tab_k, bk, nk)
            kk = np.frombuffer(bk, dtype=np.int64, count=nk)
            if np.array_equal(free, np.arange(n)) and np.array_equal(kk, np.arange(nk) * n):
                dense += 1
    assert total > 0 and dense == total


def test_plan_errors(built_lib):
    import tenet_jl_b200 as tb
    tn = tb.TensorNetwork([tb.Tensor(np.ones((2, 3)), "ab"), tb.Tensor(np.ones((3, 2)), "bc"), tb.Tensor(np.ones((2, 2)), "ca")])
    with pytest.raises(tb.TnbError):
        tb.ContractionPlan(tn, tb.ContractionPath([(0, 1)]), dry=True)              # n-1 steps required
    with pytest.raises(tb.TnbError):
        tb.ContractionPlan(tn, tb.ContractionPath([(0, 1), (0, 2)]), dry=True)      # id consumed twice
    with pytest.raises(ValueError):
        tb.ContractionPlan(tn, tb.ContractionPath([(0, 1), (3, 2)], sliced=("zz",)), dry=True)
    with pytest.raises(tb.TnbError):
        tb.ContractionPlan(tn, tb.ContractionPath([(0, 1), (3, 2)], sliced=("a",)), output=("a",), dry=True)
    bad = tb.TensorNetwork([tb.Tensor(np.ones((2, 3)), "ab"), tb.Tensor(np.ones((4, 2)), "bc")])
    with pytest.raises((tb.TnbError, ValueError)):
        tb.ContractionPlan(bad, tb.ContractionPath([(0, 1)]), dry=True)             # extent mismatch on b


def test_pathfinder_costs_and_slicing():
    import tenet_jl_b200 as tb
    tn = tb.workloads.random_regular_network(n=30, bond=2, dtype=np.complex64, seed=3)
    inputs = [t.inds for t in tn.tensors]
    sizes = tn.sizes()
    p = tb.optimize_path(inputs, sizes, (), ntrials=8, seed=1)
    macs, mx, _ = orc.path_flops(inputs, sizes, p.steps)
    assert abs(np.log2(float(macs)) - p.log2_macs) < 1e-6 and abs(np.log2(float(mx)) - p.log2_max_size) < 1e-6
    p2 = tb.find_slices(inputs, sizes, (), p, p.log2_max_size - 2)
    assert p2.log2_max_size <= p.log2_max_size - 2 + 1e-9 and p2.nslices >= 4
    macs2, mx2, _ = orc.path_flops(inputs, sizes, p2.steps, sliced=p2.sliced)
    assert abs(np.log2(float(macs2)) - p2.log2_macs) < 1e-6
    # committed bench path is loadable and consistent with the generator
    from tools.make_paths import network
    t53 = network("sycamore53_m14")
    bp = tb.pathfinder.load_path(t53.inds("all"), os.path.join(ROOT, "bench_paths", "sycamore53_m14.json"))
    assert len(bp.steps) == len(t53.tensors) - 1 and len(set(bp.sliced)) == len(bp.sliced)
    m3, x3, _ = orc.path_flops([t.inds for t in t53.tensors], t53.sizes(), bp.steps, sliced=bp.sliced)
    assert abs(np.log2(float(m3)) - bp.log2_macs) < 1e-6 and abs(np.log2(float(x3)) - bp.log2_max_size) < 1e-6


@pytest.mark.parametrize("slicing", ["greedy", "interleaved"])
def test_hyper_search_sliced_paths_are_valid_and_exact(slicing):
    """tree search with both slicing strategies: the advertised costs match an independent recount, the peak fits the
    target, and the sliced contraction (oracle, complex128) equals the unsliced one."""
    import tenet_jl_b200 as tb
    from tenet_jl_b200 import treeopt
    tn = tb.workloads.random_regular_network(n=60, bond=2, dtype=np.complex128, seed=11)
    inputs = [t.inds for t in tn.tensors]
    sizes = tn.sizes()
    p0 = tb.optimize_path(inputs, sizes, (), ntrials=8, seed=0)
    target = p0.log2_max_size - 3
    assert target >= 5                      # a handful of slices, not thousands (the oracle loops over them)
    p = treeopt.hyper_search(inputs, sizes, (), ntrials=16, seed=2, target_log2_size=target, reconf_size=6,
                             reconf_rounds=1, keep=2, minimize="time", slicing=slicing)
    assert 2 <= p.nslices <= 1024 and p.log2_max_size <= target + 1e-9
    macs, mx, _ = orc.path_flops(inputs, sizes, p.steps, sliced=p.sliced)
    assert abs(np.log2(float(macs)) - p.log2_macs) < 1e-6 and abs(np.log2(float(mx)) - p.log2_max_size) < 1e-6
    arrays = [t.parent for t in tn.tensors]
    full, _ = orc.contract_path(arrays, inputs, p0.steps)
    sl, _ = orc.contract_sliced(arrays, inputs, p.steps, list(p.sliced))
    assert abs(complex(sl) - complex(full)) < 1e-10 * max(abs(complex(full)), 1e-30)


def test_committed_path_kernel_selection(built_lib):
    """dry plan of the committed 53-qubit path: the planner must route the work to the kernels bench.py reports on
    (no silent fall-back of a big step to the generic kernel), with the workspace the streamed-B / split-K steps need."""
    import tenet_jl_b200 as tb
    from tools.make_paths import network
    tn = network("sycamore53_m14")
    path = tb.pathfinder.load_path(tn.inds("all"), os.path.join(ROOT, "bench_paths", "sycamore53_m14.json"))
    plan = tb.ContractionPlan(tn, path, dry=True)
    info = plan.info
    assert info["nslices"] == 256 and info["nsteps_per_slice"] + info["nsteps_hoisted"] == len(tn.tensors) - 1
    flops = {}
    for s in range(plan.nsteps):
        si = plan.step_info(s)
        if not si["hoisted"]:
            flops[si["kernel_name"]] = flops.get(si["kernel_name"], 0.0) + si["flops"]
            # K <= 128 huge x small steps (small side a multiple of 128, <= 16 passes) belong to the stem kernel's
            # 128-column passes, not to the tile GEMM kernel (measured 136 -> 186 TFLOP/s on 262144 x 2048 x 128)
            if si["kernel_name"] == "c64_tf32x3" and si["flops"] > 1e11:
                assert si["K"] > 128, si
    tot = sum(flops.values())
    assert abs(tot - info["flops_per_slice"]) / tot < 1e-9
    tc = flops.get("c64_tf32x3", 0) + flops.get("stem_tc", 0)
    assert tc / tot > 0.98, flops                              # tensor-core kernels carry the flops
    assert flops.get("generic", 0) + flops.get("splitk", 0) < 1e-3 * tot, flops
    assert flops.get("stem_tc", 0) > 0.45 * tot, flops
    assert "stream" in flops                                   # the environment-closing k-reduction step
    assert info["workspace_bytes"] < 64 * 2 ** 30              # arena + workspace fit one B200 with room to spare


def test_stem_plan_flags(built_lib):
    """stem steps: power-of-two extents give a separable rank (bit deposit) and even run bases; an interleaved output
    layout must still be recognised as a stem step with runs of the expected length."""
    import tenet_jl_b200 as tb
    a = np.zeros((2,) * 17 + (32,), np.complex64)
    b = np.zeros((32, 32), np.complex64)
    big = [f"m{i}" for i in range(17)]
    for out, kern in [(None, "stem_tc"), (big[:3] + ["n"] + big[3:], "stem_tc"), (big[9:] + ["n"] + big[:9], "stem_tc")]:
        tn = tb.TensorNetwork([tb.Tensor(a, big + ["k"]), tb.Tensor(b, ["n", "k"])])
        plan = tb.ContractionPlan(tn, tb.ContractionPath([(0, 1)]), output=out, dry=True)
        si = plan.step_info(0)
        assert si["kernel_name"] == kern and si["M"] * si["N"] == (1 << 17) * 32 and si["K"] == 32
    # small operand too large to stay resident: still the stem kernel (streamed-B mode)
    b2 = np.zeros((128, 128), np.complex64)
    a2 = np.zeros((2,) * 17 + (128,), np.complex64)
    tn = tb.TensorNetwork([tb.Tensor(a2, big + ["k"]), tb.Tensor(b2, ["n", "k"])])
    plan = tb.ContractionPlan(tn, tb.ContractionPath([(0, 1)]), dry=True)
    assert plan.step_info(0)["kernel_name"] == "stem_tc"


def test_stem_direct_epilogue_criterion(built_lib, monkeypatch, capfd):
    """planner.cpp st_direct: a stem step may store rows straight from registers exactly when a warp's 32 consecutive rows
    of one column are whole 32-byte sectors (4 complex64) of the output in at most 4 lines.  Read from the planner trace
    (TNB_DEBUG_STEM)."""
    import re
    import tenet_jl_b200 as tb
    monkeypatch.setenv("TNB_DEBUG_STEM", "1")
    big = [f"m{i}" for i in range(17)]

    def direct_flag(N, K, out, env=None):
        a = np.zeros((2,) * 17 + (K,), np.complex64)
        b = np.zeros((N, K), np.complex64)
        tn = tb.TensorNetwork([tb.Tensor(a, big + ["k"]), tb.Tensor(b, ["n", "k"])])
        if env is not None:
            monkeypatch.setenv("TNB_STEM_DIRECT", env)
        capfd.readouterr()
        plan = tb.ContractionPlan(tn, tb.ContractionPath([(0, 1)]), output=out, dry=True)
        assert plan.step_info(0)["kernel_name"] == "stem_tc"
        err = capfd.readouterr().err
        if env is not None:
            monkeypatch.delenv("TNB_STEM_DIRECT")
        m = re.findall(r"\[stem\].* direct=(\d)", err)
        assert m, err
        return int(m[-1])

    for N, K in [(32, 32), (64, 16), (128, 128), (256, 64)]:
        assert direct_flag(N, K, big + ["n"]) == 1                       # rows fastest: 32 rows = 256 contiguous bytes
        assert direct_flag(N, K, big[:5] + ["n"] + big[5:]) == 1         # 32 rows, then the small index
        assert direct_flag(N, K, big[:3] + ["n"] + big[3:]) == 1         # 8 rows (64 bytes), then the small index
        assert direct_flag(N, K, big[:2] + ["n"] + big[2:]) == 0         # 4 rows = 8 sectors in 8 lines per store: staged write-out
        assert direct_flag(N, K, ["n"] + big) == 0                       # small index fastest: staged write-out
        assert direct_flag(N, K, big[1:] + ["n"] + big[:1]) == 1         # lanes = m0..m4: two 128-byte pieces (m1..m4) per store
        assert direct_flag(N, K, big[3:] + ["n"] + big[:3]) == 0         # only m3, m4 of the lane rows are adjacent: 8 lines per store
        assert direct_flag(N, K, big + ["n"], env="0") == 0              # switch (read at plan time)
