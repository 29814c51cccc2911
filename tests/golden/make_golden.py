"""Generates tests/golden/*.npz.  Run from the repo root: python tests/golden/make_golden.py

The reference cannot run here (no Julia; arithmetic lives in un-vendored Muscle/Tangles), so goldens are
(1) values the reference's own tests state, reproduced with the numpy oracle, and (2) oracle outputs on seeded
inputs, frozen so that GPU results are compared against committed numbers rather than a moving target.
"""
import os
import sys

import numpy as np
from scipy.optimize import minimize

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import einsum_oracle as orc  # noqa: E402
from oracle import networks as onet  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def tfim_product_state():
    """P4/K3: minimise <psi|H|psi> over real product states for the n=10, h=J=1 TFIM MPO
    (test/unit/dmrg.jl:10-13 states E/n = -1.1902477482849715)."""
    n = 10
    Ws, Winds = onet.ising_1d_mpo(n, 1.0, 1.0)

    def energy(thetas):
        arrays, inds = onet.product_expectation_network(thetas, Ws, Winds)
        v, _ = orc.contract_path(arrays, inds, onet.sweep_steps(n))
        return float(np.real(v))

    best = None
    rng = np.random.default_rng(0)
    for _ in range(8):
        r = minimize(energy, rng.uniform(0.5, 1.5, n), method="BFGS", options={"gtol": 1e-12})
        r = minimize(energy, r.x, method="Nelder-Mead", options={"xatol": 1e-13, "fatol": 1e-15, "maxiter": 20000})
        if best is None or r.fun < best.fun:
            best = r
    print("TFIM product-state E/n =", best.fun / n, "(reference: -1.1902477482849715)")
    assert abs(best.fun / n + 1.1902477482849715) < 1e-9
    np.savez(os.path.join(HERE, "tfim_product_state.npz"), thetas=best.x, energy=best.fun)


if __name__ == "__main__":
    tfim_product_state()
