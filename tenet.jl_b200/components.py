"""Front-ends `ProductState`, `MPS`, `MPO`, `PEPS`, `ising_1d_mpo`, `overlap`, `expect` — only as much as the hot
path needs: they build the `Tensor` lists (index naming, array orders, boundary shapes) exactly as the
reference constructors do, so the networks handed to `contract` / `binary_einsum` are the reference's.
Everything canonical-form / SVD related is out of scope.

  MPS(arrays; order)      /root/reference/src/Components/MPS.jl:50-94   (default order (:l,:r,:o), :6)
  rand(MPS; n, maxdim)    MPS.jl:113-165  (LQ -> right-canonical, norm 1)
  MPO(arrays; order)      src/Components/MPO.jl:85-132 (default (:l,:r,:o,:i), :7)
  ProductState            src/Components/ProductState.jl:45-55,72-92
  PEPS(arrays; order)     src/Components/PEPS.jl:14-74 (default (:l,:r,:u,:d,:o))
  ising_1d_mpo            src/Models/Ising.jl:12-30
  overlap                 src/Operations/overlap.jl:5-50
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .network import TensorNetwork, contract
from .tensor import Tensor, binary_einsum


def plug(i, prime: bool = False):
    return ("plug", i, "'") if prime else ("plug", i)


def bond(a, b):
    return ("bond", a, b)


class _Wrapper:
    """Components wrap a TensorNetwork in field `tn` and delegate (DelegateToField{:tn}, MPS.jl:13-29)."""

    def __init__(self, tensors: Sequence[Tensor]):
        self.tn = TensorNetwork(tensors)

    @property
    def tensors(self) -> List[Tensor]:
        return self.tn.tensors

    def nsites(self):
        return len(self.tn.tensors)

    def tensor_at(self, site: int) -> Tensor:
        """`tensor_at(tn, site"i")`, 1-based like the reference."""
        return self.tn.tensors[site - 1]

    def conj(self):
        c = object.__new__(type(self))
        c.tn = self.tn.conj()
        return c

    def replace(self, mapping):
        c = object.__new__(type(self))
        c.tn = self.tn.replace(mapping)
        return c

    def inds(self, set="all"):
        return self.tn.inds(set)


def _order_check(order, default, name):
    order = tuple(order)
    if set(order) != set(default) or len(order) != len(default):
        raise ValueError(f"order must be a permutation of {default} for {name}")
    return order


class ProductState(_Wrapper):
    def __init__(self, arrays):
        if isinstance(arrays, str):
            s2 = 1 / np.sqrt(2)
            table = {"0": [1.0, 0.0], "1": [0.0, 1.0], "+": [s2, s2], "-": [s2, -s2],
                     "u": [s2, -1j * s2], "d": [-s2, -1j * s2]}
            try:
                arrays = [np.array(table[c]) for c in arrays]
            except KeyError as e:
                raise ValueError(f"invalid character: {e.args[0]}")
        super().__init__([Tensor(np.asarray(a), [plug(i + 1)]) for i, a in enumerate(arrays)])


class MPS(_Wrapper):
    DEFAULT_ORDER = ("l", "r", "o")

    def __init__(self, arrays: Sequence[np.ndarray], order=DEFAULT_ORDER):
        order = _order_check(order, self.DEFAULT_ORDER, "MPS")
        n = len(arrays)
        if np.ndim(arrays[0]) != 2 or np.ndim(arrays[-1]) != 2:
            raise ValueError("First and last arrays must have 2 dimensions")
        if any(np.ndim(a) != 3 for a in arrays[1:-1]):
            raise ValueError("All bulk arrays must have 3 dimensions")
        ts = []
        for k, a in enumerate(arrays):
            i = k + 1
            local = [d for d in order if not (i == 1 and d == "l") and not (i == n and d == "r")]
            inds = [plug(i) if d == "o" else bond(i, i + 1) if d == "r" else bond(i - 1, i) for d in local]
            ts.append(Tensor(a, inds))
        super().__init__(ts)

    @staticmethod
    def bond_dims(n, maxdim, physdim=2):
        """chi_l, chi_r per site as in MPS.jl:122-150."""
        dims = []
        for i in range(1, n + 1):
            after_mid = i > n // 2
            ii = (n + 1 - abs(2 * i - n - 1)) // 2
            cl = min(maxdim, physdim ** min(ii - 1, 62))
            cr = min(maxdim, physdim ** min(ii, 62))
            if n % 2 == 1 and i == n // 2 + 1:
                dims.append((cl, cl))
            else:
                dims.append((cr, cl) if after_mid else (cl, cr))
        return dims

    @classmethod
    def rand(cls, n, maxdim=128, eltype=np.float64, physdim=2, rng=None):
        """Right-canonical random MPS, <psi|psi> = 1 (MPS.jl:113-165).  The LQ/QR factorisation is host-side
        setup (LAPACK), not part of the accelerated path."""
        rng = np.random.default_rng(rng)
        eltype = np.dtype(eltype)
        arrays = []
        for (cl, cr) in cls.bond_dims(n, maxdim, physdim):
            x = rng.standard_normal((cl, physdim * cr))
            if eltype.kind == "c":
                x = x + 1j * rng.standard_normal((cl, physdim * cr))
            # LQ of x: rows of Q orthonormal  <=>  QR of x^H
            q, _ = np.linalg.qr(x.conj().T)
            Q = q.conj().T[:cl]
            arrays.append(np.reshape(Q.astype(eltype), (cl, physdim, cr), order="F"))
        arrays[0] = arrays[0].reshape(physdim, -1, order="F")
        arrays[-1] = arrays[-1].reshape(-1, physdim, order="F")
        return cls(arrays, order=("l", "o", "r"))


class MPO(_Wrapper):
    DEFAULT_ORDER = ("l", "r", "o", "i")

    def __init__(self, arrays, order=DEFAULT_ORDER):
        order = _order_check(order, self.DEFAULT_ORDER, "MPO")
        n = len(arrays)
        if np.ndim(arrays[0]) != 3 or np.ndim(arrays[-1]) != 3:
            raise ValueError("First and last arrays must have 3 dimensions")
        if any(np.ndim(a) != 4 for a in arrays[1:-1]):
            raise ValueError("All bulk arrays must have 4 dimensions")
        ts = []
        for k, a in enumerate(arrays):
            i = k + 1
            local = [d for d in order if not (i == 1 and d == "l") and not (i == n and d == "r")]
            inds = [plug(i) if d == "o" else plug(i, True) if d == "i" else bond(i, i + 1) if d == "r" else bond(i - 1, i)
                    for d in local]
            ts.append(Tensor(a, inds))
        super().__init__(ts)


class PEPS(_Wrapper):
    DEFAULT_ORDER = ("l", "r", "u", "d", "o")

    def __init__(self, arrays, order=DEFAULT_ORDER):
        order = _order_check(order, self.DEFAULT_ORDER, "PEPS")
        m, n = len(arrays), len(arrays[0])
        self.grid = (m, n)
        ts = []
        for i in range(1, m + 1):
            for j in range(1, n + 1):
                dirs = [d for d in order if not (i == 1 and d == "u") and not (i == m and d == "d")
                        and not (j == 1 and d == "l") and not (j == n and d == "r")]
                inds = []
                for d in dirs:
                    if d == "l":
                        inds.append(bond((i, j - 1), (i, j)))
                    elif d == "r":
                        inds.append(bond((i, j), (i, j + 1)))
                    elif d == "u":
                        inds.append(bond((i - 1, j), (i, j)))
                    elif d == "d":
                        inds.append(bond((i, j), (i + 1, j)))
                    else:
                        inds.append(plug((i, j)))
                ts.append(Tensor(arrays[i - 1][j - 1], inds))
        super().__init__(ts)


def ising_1d_mpo(L: int, h: float, J: float) -> MPO:
    """Uniform open-boundary TFIM  H = -J sum Z_i Z_{i+1} - h sum X_i  as a bond-3 MPO (Ising.jl:12-30)."""
    Id = np.eye(2)
    sx = np.array([[0.0, 1.0], [1.0, 0.0]])
    sz = np.array([[1.0, 0.0], [0.0, -1.0]])
    W = np.zeros((2, 2, 3, 3), dtype=np.complex128)
    W[:, :, 0, 0] = Id
    W[:, :, 1, 0] = sz
    W[:, :, 2, 0] = -h * sx
    W[:, :, 2, 1] = -J * sz
    W[:, :, 2, 2] = Id
    W1 = W[:, :, 2, :]
    Wn = W[:, :, :, 0]
    return MPO([W1] + [W for _ in range(L - 2)] + [Wn], order=("i", "o", "l", "r"))


def _conj_reset(phi: _Wrapper, tag):
    """`resetinds!(conj(phi))` + `align!(psi, :outputs, phi, :outputs)`: conjugate, give every inner index a
    fresh name, keep the plugs so they pair with psi's (overlap.jl:7-8,39-40)."""
    c = phi.conj()
    plugs = set(i for i in c.inds("all") if isinstance(i, tuple) and i and i[0] == "plug")
    mapping = {i: (tag, i) for i in c.inds("all") if i not in plugs}
    return c.replace(mapping)


def overlap(a, b, **kw):
    """<b|a>.  MPS x MPS uses the reference's left-to-right zipper of `binary_einsum` (overlap.jl:36-50);
    anything else builds the joint network and calls `contract` (overlap.jl:5-13)."""
    if isinstance(a, MPS) and isinstance(b, MPS) and not kw:
        if a.nsites() != b.nsites():
            raise ValueError("both MPS must have the same number of sites")
        bc = _conj_reset(b, "bra")
        env = binary_einsum(a.tensor_at(1), bc.tensor_at(1))
        for i in range(2, a.nsites() + 1):
            env = binary_einsum(binary_einsum(env, a.tensor_at(i)), bc.tensor_at(i))
        return env
    bc = _conj_reset(b, "bra")
    tn = TensorNetwork()
    tn.append(a.tn)
    tn.append(bc.tn)
    return contract(tn, **kw)


def expect_network(psi: MPS, op: MPO) -> TensorNetwork:
    """<psi| op |psi> as one closed network: ket plugs -> operator inputs (plug i'), operator outputs (plug i)
    -> bra plugs."""
    ket = psi.replace({plug(i): plug(i, True) for i in range(1, psi.nsites() + 1)})
    ket = ket.replace({i: ("ket", i) for i in ket.inds("all") if i[0] == "bond"})
    bra = _conj_reset(psi, "bra")
    tn = TensorNetwork()
    tn.append(ket.tn)
    tn.append(op.tn)
    tn.append(bra.tn)
    return tn
