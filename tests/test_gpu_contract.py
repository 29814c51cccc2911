"""GPU parity: tnb_plan_* / tnb_contract_path (through the Python mirror of Tangles.contract) vs the numpy oracle,
the reference's own pins and size-independent properties (slice-sum linearity)."""
import numpy as np
import pytest

from oracle import einsum_oracle as orc
from oracle import statevector as sv

from tolerances import C128_BOUND, C64_PATH_BOUND, C64_STEP_BOUND  # noqa: F401

pytestmark = pytest.mark.gpu


def _arrays(tn):
    return [t.parent for t in tn.tensors], [t.inds for t in tn.tensors]


def _rel(got, ref):
    got, ref = np.asarray(got), np.asarray(ref)
    return np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)


@pytest.mark.parametrize("dt,tol", [(np.complex128, C128_BOUND), (np.complex64, C64_PATH_BOUND), (np.float64, C128_BOUND)])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_regular_closed(ctx, dt, tol, seed):
    import tenet_jl_b200 as tb
    tn = tb.workloads.random_regular_network(n=14, bond=3, dtype=dt, seed=seed)
    path = tb.einexpr(tn, ntrials=4, seed=seed)
    got = tb.contract(tn, path=path).item()
    arrays, inds = _arrays(tn)
    hi = np.complex128 if np.dtype(dt).kind == "c" else np.float64
    ref, _ = orc.contract_path([a.astype(hi) for a in arrays], inds, path.steps)
    assert _rel(got, ref) < tol


def test_open_indices_and_output_order(ctx):
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(3)
    mk = lambda shape: rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    ts = [tb.Tensor(mk((3, 4, 5)), ("a", "b", "c")), tb.Tensor(mk((5, 6, 2)), ("c", "d", "e")),
          tb.Tensor(mk((4, 6, 7)), ("b", "d", "f")), tb.Tensor(mk((7, 3)), ("f", "g"))]
    tn = tb.TensorNetwork(ts)
    arrays, inds = _arrays(tn)
    for out in [None, ("g", "e", "a"), ("a", "g", "e")]:
        c = tb.contract(tn, output=out)
        ref, ri = orc.contract_path(arrays, inds, tb.einexpr(tn, output=out).steps, output=out)
        assert tuple(c.inds) == tuple(ri)
        assert _rel(c.parent, ref) < 1e-12


def test_hyperindex_batch(ctx):
    """an index carried by three tensors stays a batch index until its last carrier is absorbed."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(4)
    mk = lambda shape: rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    ts = [tb.Tensor(mk((3, 4)), ("h", "a")), tb.Tensor(mk((3, 5)), ("h", "b")), tb.Tensor(mk((3, 4, 5)), ("h", "a", "b"))]
    tn = tb.TensorNetwork(ts)
    arrays, inds = _arrays(tn)
    path = tb.ContractionPath([(0, 1), (3, 2)])
    got = tb.contract(tn, path=path, output=()).item()
    ref = np.einsum("ha,hb,hab->", *arrays)
    assert _rel(got, ref) < 1e-12


@pytest.mark.parametrize("dt,tol", [(np.complex128, C128_BOUND), (np.complex64, C64_PATH_BOUND)])
def test_sliced_equals_unsliced_and_oracle(ctx, dt, tol):
    import tenet_jl_b200 as tb
    tn = tb.workloads.random_regular_network(n=24, bond=3, dtype=dt, seed=5)
    p0 = tb.einexpr(tn, ntrials=4, seed=0)
    p1 = tb.einexpr(tn, ntrials=4, seed=0, max_log2_size=p0.log2_max_size - 2.2)
    assert 1 < p1.nslices <= 3 ** 8
    full = tb.contract(tn, path=p0).item()
    sliced = tb.contract(tn, path=p1).item()
    arrays, inds = _arrays(tn)
    ref, _ = orc.contract_sliced([a.astype(np.complex128) for a in arrays], inds, p1.steps, p1.sliced)
    assert _rel(sliced, ref) < tol
    assert _rel(sliced, full) < 10 * tol
    # linearity over slice ranges: two interleaved halves add up to the whole (the multi-GPU partition)
    h0 = tb.contract(tn, path=p1, slice_range=(0, 2, p1.nslices)).item()
    h1 = tb.contract(tn, path=p1, slice_range=(1, 2, p1.nslices)).item()
    assert _rel(h0 + h1, sliced) < 10 * tol
    r0, _ = orc.contract_sliced([a.astype(np.complex128) for a in arrays], inds, p1.steps, p1.sliced,
                                slice_ids=range(0, p1.nslices, 2))
    assert _rel(h0, r0) < tol


def test_mps_norm_is_one_K1(ctx):
    """K1: rand(MPS) is right-canonical => <psi|psi> = 1 (MPS.jl:103-104,154-157); zipper path + greedy path."""
    import tenet_jl_b200 as tb
    tn, psi = tb.workloads.mps_norm_network(n=12, chi=16, dtype=np.complex128, seed=1)
    v = tb.contract(tn, path=tb.workloads.zipper_path(12)).item()
    assert abs(v - 1.0) < 1e-12
    v2 = tb.contract(tn).item()
    assert abs(v2 - 1.0) < 1e-12
    v3 = tb.overlap(psi, psi).item()
    assert abs(v3 - 1.0) < 1e-12


def test_x_gate_kat_P1(ctx):
    """P1: X on [1,0] / [0,1] through binary_einsum (test/unit/simple_update.jl:4-14; simple_update.jl:28-37)."""
    import tenet_jl_b200 as tb
    X = tb.Tensor(np.array([[0.0, 1.0], [1.0, 0.0]]), ("o", "i"))
    for vec, exp in (([1.0, 0.0], [0.0, 1.0]), ([0.0, 1.0], [1.0, 0.0])):
        r = tb.binary_einsum(tb.Tensor(np.array(vec), ("i",)), X)
        assert r.inds == ("o",)
        assert np.array_equal(r.parent, np.array(exp))


def test_tfim_energy_kats_K2_K3(ctx):
    """K2: <0..0|H|0..0> = -J(n-1), <+..+|H|+..+> = -h n (Ising.jl:12-30).  P4/K3: best product state of the n=10
    h=J=1 TFIM has E/n = -1.1902477482849715 (test/unit/dmrg.jl:10-13) — evaluated through the MPO sandwich."""
    import tenet_jl_b200 as tb
    n = 10
    H = tb.ising_1d_mpo(n, 1.0, 1.0)
    for bits, exp in (("0" * n, -(n - 1)), ("+" * n, -n)):
        psi = tb.workloads.product_mps(bits, chi=1)
        tn = tb.expect_network(psi, H)
        v = tb.contract(tn, path=tb.workloads.sweep_path(n)).item()
        assert abs(v - exp) < 1e-12, (bits, v)
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "tfim_product_state.npz"))
    arrays = [np.array([np.cos(t / 2), np.sin(t / 2)]).reshape(1, 2, 1) for t in g["thetas"]]
    arrays[0] = arrays[0].reshape(2, 1)
    arrays[-1] = arrays[-1].reshape(1, 2)
    psi = tb.MPS([a.astype(np.complex128) for a in arrays], order=("l", "o", "r"))
    v = tb.contract(tb.expect_network(psi, H), path=tb.workloads.sweep_path(n)).item()
    assert abs(v.real / n - (-1.1902477482849715)) < 1e-9
    assert abs(v - complex(g["energy"])) < 1e-12


def test_gauge_trick_K4(ctx):
    """K4: product state padded to chi=32 with random bond gauges — dense tensors, same energy."""
    import tenet_jl_b200 as tb
    n = 12
    H = tb.ising_1d_mpo(n, 0.7, 1.3)
    psi = tb.workloads.product_mps("0" * n, chi=32, seed=3)
    v = tb.contract(tb.expect_network(psi, H), path=tb.workloads.sweep_path(n)).item()
    nrm = tb.overlap(psi, psi).item()
    assert abs(v / nrm - (-1.3 * (n - 1))) < 1e-9


@pytest.mark.parametrize("dt,tol", [(np.complex64, C64_PATH_BOUND), (np.complex128, 10 * C128_BOUND)])
def test_circuit_amplitude_vs_statevector(ctx, dt, tol):
    import tenet_jl_b200 as tb
    tn, (nq, gates, bits) = tb.workloads.sycamore_amplitude_network(rows=4, cols=3, cycles=8, seed=11, removed=((0, 1),), dtype=dt)
    assert nq == 11
    ref = sv.amplitude(nq, gates, bits)
    p0 = tb.einexpr(tn, ntrials=8, seed=0)
    got = tb.contract(tn, path=p0).item()
    assert abs(got - ref) / abs(ref) < tol
    p1 = tb.einexpr(tn, ntrials=8, seed=0, max_log2_size=max(4.0, p0.log2_max_size - 3))
    assert p1.nslices >= 2
    got = tb.contract(tn, path=p1).item()
    assert abs(got - ref) / abs(ref) < tol


def test_peps_norm_small(ctx):
    import tenet_jl_b200 as tb
    tn, psi = tb.workloads.peps_norm_network(3, 3, D=2, p=2, dtype=np.complex128, seed=4)
    arrays, inds = _arrays(tn)
    path = tb.workloads.peps_boundary_path(3, 3)
    got = tb.contract(tn, path=path).item()
    ref, _ = orc.contract_path(arrays, inds, path.steps)
    assert _rel(got, ref) < 1e-12
    assert abs(got.imag) < 1e-12 * abs(got.real) and got.real > 0
    got2 = tb.contract(tn).item()
    assert _rel(got2, ref) < 1e-11


def test_plan_reuse_and_info(ctx):
    import tenet_jl_b200 as tb
    tn = tb.workloads.random_regular_network(n=12, bond=4, dtype=np.complex64, seed=9)
    p = tb.einexpr(tn, ntrials=4, max_log2_size=10)
    plan = tb.ContractionPlan(tn, p)
    info = plan.info
    assert info["nslices"] == p.nslices and info["flops_per_slice"] > 0
    plan.execute(accumulate=False)
    a = plan.result().item()
    plan.execute(accumulate=False)
    b = plan.result().item()
    assert a == b          # deterministic: same launches, same order
    plan.execute(accumulate=True)
    c = plan.result().item()
    assert abs(c - 2 * a) < 1e-5 * abs(a)
    macs = sum(plan.step_info(s)["flops"] for s in range(plan.nsteps) if not plan.step_info(s)["hoisted"])
    assert abs(macs - info["flops_per_slice"]) < 1e-6 * macs
    plan.close()


def test_error_paths(ctx):
    import tenet_jl_b200 as tb
    tn = tb.TensorNetwork([tb.Tensor(np.ones((2, 3)), ("a", "b")), tb.Tensor(np.ones((3, 2)), ("b", "c"))])
    # an open index left out of `output` is summed out (einsum semantics), not an error
    r = tb.contract(tn, path=tb.ContractionPath([(0, 1)]), output=("a",))
    assert np.allclose(r.parent, np.full(2, 6.0))
    with pytest.raises((tb.TnbError, ValueError)):
        tb.contract(tn, path=tb.ContractionPath([(0, 1)]), output=("a", "zz"))  # output index nobody carries
    with pytest.raises(tb.TnbError):
        tb.contract(tn, path=tb.ContractionPath([(0, 2)]))                      # id does not exist
    with pytest.raises(ValueError):
        tb.contract(tb.TensorNetwork([]))


def _integer_mps_arrays():
    """test/unit/mps.jl:63-69: MPS([reshape(1:16,4,4), reshape(1:64,4,4,4) x 3, reshape(1:16,4,4)]) (Julia is
    column-major: order="F"), default order (:l, :r, :o)."""
    a2 = np.arange(1, 17, dtype=np.int64).reshape(4, 4, order="F")
    a3 = np.arange(1, 65, dtype=np.int64).reshape(4, 4, 4, order="F")
    return [a2, a3, a3, a3, a2]


def test_integer_valued_mps_contract_is_exact(ctx):
    """Reference-held input (test/unit/mps.jl:63-69,89): the 5-site integer-valued MPS the reference contracts in its
    canonize tests.  Int64 promotes to Float64 on the host (a1 (v)); every product and partial sum is an integer below
    2^53, so the 4^5-element result must equal the exact integer contraction BIT FOR BIT, whatever the summation order."""
    import tenet_jl_b200 as tb
    arrays = _integer_mps_arrays()
    psi = tb.MPS(arrays)
    got = tb.contract(psi)
    assert got.parent.dtype == np.float64 and got.parent.size == 4 ** 5
    # exact integer reference: sites are (r, o), (l, r, o) x 3, (l, o)
    ref = np.einsum("ap,abq,bcr,cds,dt->pqrst", arrays[0], arrays[1], arrays[2], arrays[3], arrays[4])
    assert ref.max() < 2 ** 53
    plugs = [tb.components.plug(i) for i in range(1, 6)]
    g = np.transpose(got.parent, [got.inds.index(i) for i in plugs])
    assert np.array_equal(g, ref.astype(np.float64))


def test_allreduce_refuses_partial_sum_without_communicator(ctx, monkeypatch):
    """with WORLD_SIZE > 1 in the environment but no communicator on the context, contract_distributed must
    raise instead of returning this rank's partial slice sum."""
    import tenet_jl_b200 as tb
    tn, _ = tb.workloads.sycamore_amplitude_network(rows=3, cols=3, cycles=6, seed=3, removed=(), dtype=np.complex64)
    path = tb.einexpr(tn, ntrials=4, seed=0, max_log2_size=4)
    plan = tb.ContractionPlan(tn, path, ctx=ctx)
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("WORLD_SIZE", "2")
    assert ctx.lib.tnb_comm_size(ctx.handle) == 1
    with pytest.raises(tb.TnbError):
        tb.distributed.contract_distributed(plan, ctx)
    plan.close()


def test_empty_slice_range_zeroes_the_output(ctx):
    """accumulate = 0 with an empty slice range (a rank that owns no slice) must leave zeros, not stale data."""
    import tenet_jl_b200 as tb
    tn, _ = tb.workloads.sycamore_amplitude_network(rows=3, cols=3, cycles=6, seed=3, removed=(), dtype=np.complex64)
    path = tb.einexpr(tn, ntrials=4, seed=0, max_log2_size=4)
    plan = tb.ContractionPlan(tn, path, ctx=ctx)
    plan.execute(0, 1, plan.nslices, accumulate=False)
    assert abs(plan.result().item()) > 0
    plan.execute(plan.nslices, 1, plan.nslices, accumulate=False)      # empty range
    assert plan.result().item() == 0
    plan.close()


def test_simt_mode_k_reduction_is_exact_fp32(ctx):
    """TNB_C64_SIMT must keep tensor-core rounding out of the k-reduction steps too: in SIMT mode a
    16 x 32 x K step agrees bitwise with ... itself under TNB_KRED_MMA-independent FMA arithmetic, and differs from the
    default (mma.sync 3xTF32) result only within the step bound."""
    import tenet_jl_b200 as tb
    rng = np.random.default_rng(8)
    K = 1 << 15
    a = (rng.standard_normal((16, K)) + 1j * rng.standard_normal((16, K))).astype(np.complex64)
    b = (rng.standard_normal((32, K)) + 1j * rng.standard_normal((32, K))).astype(np.complex64)
    ta, tb_ = tb.Tensor(a, ("m", "k")), tb.Tensor(b, ("n", "k"))
    c_tc = tb.binary_einsum(ta, tb_).parent.copy()
    assert ctx.last_kernel == "stream"
    ctx.set_option(tb._lib.TNB_OPT_C64_MODE, tb._lib.TNB_C64_SIMT)
    try:
        c1 = tb.binary_einsum(ta, tb_).parent.copy()
        assert ctx.last_kernel == "stream"
        c2 = tb.binary_einsum(ta, tb_).parent.copy()
    finally:
        ctx.set_option(tb._lib.TNB_OPT_C64_MODE, tb._lib.TNB_C64_TF32X3)
    ref = a.astype(np.complex128) @ b.astype(np.complex128).T
    assert np.array_equal(c1, c2)
    assert np.abs(c1 - ref).max() / np.abs(ref).max() < C64_STEP_BOUND
    assert np.abs(c_tc - ref).max() / np.abs(ref).max() < C64_STEP_BOUND
    assert not np.array_equal(c1, c_tc), "SIMT mode still runs the tensor-core k-reduction"


def test_unsliced_plan_replays_as_cuda_graph(ctx):
    """Un-sliced plans are captured once and replayed as one CUDA graph (TNB_OPT_CUDA_GRAPH, default on): same result
    bit for bit as direct launches, on the first (capturing) execute and on replays; launch accounting unchanged."""
    import tenet_jl_b200 as tb
    tn, psi = tb.workloads.mps_norm_network(12, 32, np.complex128, seed=1)
    path = tb.workloads.zipper_path(12)
    vals, launches = {}, {}
    for graphs in (0, 1):
        ctx.set_option(tb._lib.TNB_OPT_CUDA_GRAPH, graphs)
        try:
            plan = tb.ContractionPlan(tn, path, ctx=ctx)
            out = []
            l0 = ctx.launch_count
            for _ in range(3):
                plan.zero_output()
                plan.execute(0, 1, plan.nslices, accumulate=True)
                out.append(plan.result().item())
            launches[graphs] = ctx.launch_count - l0
            vals[graphs] = out
            plan.close()
        finally:
            ctx.set_option(tb._lib.TNB_OPT_CUDA_GRAPH, 1)
    assert vals[0][0] == vals[0][1] == vals[0][2] == vals[1][0] == vals[1][1] == vals[1][2]
    assert abs(vals[1][0] - 1.0) < C128_BOUND
    assert launches[0] == launches[1] > 0
