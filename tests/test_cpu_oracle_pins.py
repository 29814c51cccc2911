"""CPU: pin the oracle (oracle/) against every golden the reference's own tests hold for this path and against
first-principles checks, BEFORE it is trusted as the checker of the CUDA path.

  P1  X-gate KAT                        /root/reference/test/unit/simple_update.jl:4-14
  P4  TFIM product-state energy         /root/reference/test/unit/dmrg.jl:10-13   (-1.1902477482849715 per site)
  K1  <psi|psi> = 1 for rand(MPS)       /root/reference/src/Components/MPS.jl:103-104,154-157
  K2  <0..0|H|0..0> = -J(n-1), <+..+|H|+..+> = -h n   /root/reference/src/Models/Ising.jl:12-30
  P6  rand(MPS; n=8) bond sizes         /root/reference/test/unit/mps.jl:524-527
"""
import os

import numpy as np
import pytest

from oracle import einsum_oracle as orc
from oracle import networks as onet
from oracle import statevector as sv

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_P1_x_gate_kat():
    X = np.array([[0.0, 1.0], [1.0, 0.0]])
    for vec, exp in (([1.0, 0.0], [0.0, 1.0]), ([0.0, 1.0], [1.0, 0.0])):
        c, ci = orc.binary_einsum(np.array(vec), ("i",), X, ("o", "i"))
        assert ci == ("o",) and np.array_equal(c, np.array(exp))


def test_binary_einsum_semantics_vs_numpy_einsum():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((3, 4, 5)) + 1j * rng.standard_normal((3, 4, 5))
    b = rng.standard_normal((5, 4, 6)) + 1j * rng.standard_normal((5, 4, 6))
    c, ci = orc.binary_einsum(a, "xyz", b, "zyw")                       # (i) all shared contracted
    assert ci == ("x", "w") and np.allclose(c, np.einsum("xyz,zyw->xw", a, b))
    c, ci = orc.binary_einsum(a, "xyz", b, "zyw", dims=())              # (ii) Hadamard on shared
    assert ci == ("x", "w", "y", "z") and np.allclose(c, np.einsum("xyz,zyw->xwyz", a, b))
    c, ci = orc.binary_einsum(a, "xyz", b, "zyw", dims=("z",))          # y is batch
    assert np.allclose(c, np.einsum("xyz,zyw->xwy", a, b))
    s = np.array(2.0 - 1j)
    c, ci = orc.binary_einsum(s, (), a, "xyz")                          # (iii) rank-0 operand
    assert ci == ("x", "y", "z") and np.allclose(c, s * a)
    c, ci = orc.binary_einsum(a, "xyz", a.conj(), "xyz")                # (iv) rank-0 result
    assert ci == () and np.allclose(c, np.vdot(a, a))
    c, ci = orc.binary_einsum(a, "xyz", b, "zyw", out_inds=("w", "x"))
    assert np.allclose(c, np.einsum("xyz,zyw->wx", a, b))
    c, ci = orc.binary_einsum(np.arange(12).reshape(3, 4), "ab", rng.standard_normal((4, 2)), "bc")   # (v) Int x Float
    assert c.dtype == np.float64


def test_P4_tfim_product_state_energy_golden():
    g = np.load(os.path.join(GOLD, "tfim_product_state.npz"))
    n = 10
    Ws, Winds = onet.ising_1d_mpo(n, 1.0, 1.0)
    arrays, inds = onet.product_expectation_network(g["thetas"], Ws, Winds)
    v, vi = orc.contract_path(arrays, inds, onet.sweep_steps(n))
    assert vi == ()
    assert abs(v.real / n - (-1.1902477482849715)) < 1e-9      # the value test/unit/dmrg.jl:13 states (atol 1e-4 there)
    assert abs(v - complex(g["energy"])) < 1e-12


@pytest.mark.parametrize("n,h,J", [(10, 1.0, 1.0), (7, 0.3, 2.0)])
def test_K2_ising_product_states(n, h, J):
    Ws, Winds = onet.ising_1d_mpo(n, h, J)
    for theta, exp in ((0.0, -J * (n - 1)), (np.pi / 2, -h * n)):
        arrays, inds = onet.product_expectation_network([theta] * n, Ws, Winds)
        v, _ = orc.contract_path(arrays, inds, onet.sweep_steps(n))
        assert abs(v - exp) < 1e-12


def test_K1_rand_mps_norm_and_P6_bond_sizes():
    arrays = onet.rand_mps(8, 128, seed=1)
    bonds = [a.shape[-1] for a in arrays[:-1]]
    assert bonds == [2, 4, 8, 16, 8, 4, 2]                     # test/unit/mps.jl:524-527
    arrays = onet.rand_mps(12, 16, seed=3)
    a, i = onet.norm_network(arrays)
    v, _ = orc.contract_path(a, i, onet.zipper_steps(12))
    assert abs(v - 1.0) < 1e-13


def test_sliced_sum_equals_full_and_partitions_add():
    rng = np.random.default_rng(5)
    arrays = [rng.standard_normal((3, 4)), rng.standard_normal((4, 5, 2)), rng.standard_normal((5, 3, 2))]
    inds = [("a", "b"), ("b", "c", "s"), ("c", "a", "s")]
    steps = [(0, 1), (3, 2)]
    full, _ = orc.contract_path(arrays, inds, steps)
    sl, _ = orc.contract_sliced(arrays, inds, steps, sliced=("s", "b"))
    assert np.allclose(sl, full)
    p0, _ = orc.contract_sliced(arrays, inds, steps, sliced=("s", "b"), slice_ids=range(0, 8, 2))
    p1, _ = orc.contract_sliced(arrays, inds, steps, sliced=("s", "b"), slice_ids=range(1, 8, 2))
    assert np.allclose(p0 + p1, full)
    empty, _ = orc.contract_sliced(arrays, inds, steps, sliced=("s", "b"), slice_ids=[])
    assert np.all(empty == 0)


def test_circuit_network_vs_statevector():
    """generator + contraction vs dense simulation (independent first-principles check of cfg3's construction)."""
    import tenet_jl_b200 as tb
    tn, (nq, gates, bits) = tb.workloads.sycamore_amplitude_network(rows=3, cols=3, cycles=5, seed=3, removed=(),
                                                                     dtype=np.complex128)
    ref = sv.amplitude(nq, gates, bits)
    arrays = [t.parent for t in tn.tensors]
    inds = [t.inds for t in tn.tensors]
    p = tb.einexpr(tn, ntrials=4)
    v, _ = orc.contract_path(arrays, inds, p.steps)
    assert abs(v - ref) < 1e-13
    # unsimplified network gives the same amplitude
    tn2 = tb.workloads.circuit_amplitude_network(nq, gates, bits, np.complex128, simplify=False)
    p2 = tb.einexpr(tn2, ntrials=2)
    v2, _ = orc.contract_path([t.parent for t in tn2.tensors], [t.inds for t in tn2.tensors], p2.steps)
    assert abs(v2 - ref) < 1e-13
    probs = np.abs(sv.simulate(nq, gates)) ** 2
    assert abs(probs.sum() - 1.0) < 1e-12


def test_package_generators_match_oracle_generators():
    """tenet.jl_b200's front-end mirrors (MPS.rand, ising_1d_mpo, expect_network) vs oracle/networks.py."""
    import tenet_jl_b200 as tb
    n = 6
    H = tb.ising_1d_mpo(n, 0.7, 1.3)
    Ws, _ = onet.ising_1d_mpo(n, 0.7, 1.3)
    for t, w in zip(H.tensors, Ws):
        assert np.array_equal(t.parent, w)
    psi = tb.MPS.rand(n, maxdim=4, eltype=np.complex128, rng=7)
    ref = onet.rand_mps(n, 4, seed=7)
    for t, a in zip(psi.tensors, ref):
        assert np.allclose(t.parent, a)
    tn = tb.expect_network(psi, H)
    v, _ = orc.contract_path([t.parent for t in tn.tensors], [t.inds for t in tn.tensors], tb.workloads.sweep_path(n).steps)
    a, i = onet.expectation_network(ref, *onet.ising_1d_mpo(n, 0.7, 1.3))
    v2, _ = orc.contract_path(a, i, onet.sweep_steps(n))
    assert abs(v - v2) < 1e-12 and abs(v.imag) < 1e-12


def test_integer_valued_mps_oracle_exact():
    """test/unit/mps.jl:63-69,89 input through the oracle's contract_path: exact integers (< 2^53) in float64."""
    from oracle import einsum_oracle as orc
    a2 = np.arange(1, 17, dtype=np.int64).reshape(4, 4, order="F")
    a3 = np.arange(1, 65, dtype=np.int64).reshape(4, 4, 4, order="F")
    arrays = [a2, a3, a3, a3, a2]
    inds = [("b12", "p1"), ("b12", "b23", "p2"), ("b23", "b34", "p3"), ("b34", "b45", "p4"), ("b45", "p5")]
    out = ("p1", "p2", "p3", "p4", "p5")
    got, gi = orc.contract_path([a.astype(np.float64) for a in arrays], inds, [(0, 1), (5, 2), (6, 3), (7, 4)], out)
    ref = np.einsum("ap,abq,bcr,cds,dt->pqrst", *arrays)
    assert tuple(gi) == out and np.array_equal(got, ref.astype(np.float64))
