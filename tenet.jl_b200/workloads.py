"""Synthetic networks of the five BASELINE.json configs plus the paths the reference would walk.

All generators are deterministic in their seed (numpy `default_rng`) and return host-side `TensorNetwork`s of
numpy-backed `Tensor`s — the inputs a Tenet.jl user would hand to `contract`.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .components import MPO, MPS, PEPS, _conj_reset, expect_network, ising_1d_mpo
from .network import TensorNetwork
from .pathfinder import ContractionPath
from .tensor import Tensor


# ---------------------------------------------------------------------------------------------------
# cfg1: MPS <psi|psi>
# ---------------------------------------------------------------------------------------------------
def mps_norm_network(n=32, chi=128, dtype=np.complex128, seed=1):
    """psi = rand(MPS; n, maxdim=chi) (right-canonical => <psi|psi> = 1); network = psi U conj(psi), leaves
    interleaved ket_1, bra_1, ket_2, bra_2 ... so that the zipper (overlap.jl:36-50) is a linear path."""
    psi = MPS.rand(n, maxdim=chi, eltype=dtype, rng=seed)
    bra = _conj_reset(psi, "bra")
    ts = []
    for i in range(n):
        ts += [psi.tensors[i], bra.tensors[i]]
    return TensorNetwork(ts), psi


def zipper_path(n: int) -> ContractionPath:
    """left_env = A1*B1; for i: left_env = (left_env * A_i) * B_i   — on leaves interleaved as above."""
    steps, nl = [], 2 * n
    steps.append((0, 1))
    cur = nl
    for i in range(1, n):
        steps.append((cur, 2 * i))
        steps.append((cur + 1, 2 * i + 1))
        cur += 2
    return ContractionPath(steps)


# ---------------------------------------------------------------------------------------------------
# cfg2: random 3-regular network
# ---------------------------------------------------------------------------------------------------
def random_regular_graph(n: int, d: int, rng) -> List[Tuple[int, int]]:
    """Pairing-model random d-regular simple graph (retry until simple)."""
    assert (n * d) % 2 == 0
    while True:
        stubs = np.repeat(np.arange(n), d)
        rng.shuffle(stubs)
        edges = set()
        ok = True
        for a, b in zip(stubs[0::2], stubs[1::2]):
            a, b = int(min(a, b)), int(max(a, b))
            if a == b or (a, b) in edges:
                ok = False
                break
            edges.add((a, b))
        if ok:
            return sorted(edges)


def random_regular_network(n=200, bond=4, dtype=np.complex64, seed=0, degree=3):
    rng = np.random.default_rng(seed)
    edges = random_regular_graph(n, degree, rng)
    inc: Dict[int, list] = {v: [] for v in range(n)}
    for (a, b) in edges:
        inc[a].append(("e", a, b))
        inc[b].append(("e", a, b))
    dt = np.dtype(dtype)
    scale = 1.0 / math.sqrt(2.0 * bond ** (degree / 2.0) / 2.0)
    ts = []
    for v in range(n):
        shape = (bond,) * len(inc[v])
        x = rng.uniform(-1, 1, shape)
        if dt.kind == "c":
            x = x + 1j * rng.uniform(-1, 1, shape)
        ts.append(Tensor((x * scale / math.sqrt(2)).astype(dt), inc[v]))
    return TensorNetwork(ts)


# ---------------------------------------------------------------------------------------------------
# cfg3: Sycamore-like random circuit amplitude
# ---------------------------------------------------------------------------------------------------
def sycamore_layout(rows=9, cols=6, removed=((0, 3),)):
    """Diagonal (staggered) grid: qubit (r,c); even rows couple down-right to (r+1,c) and down-left to (r+1,c-1),
    odd rows down-left to (r+1,c) and down-right to (r+1,c+1).  9x6 minus one = 53 qubits, 86 couplers.
    Couplers are split into four disjoint matchings A,B (one diagonal) and C,D (the other)."""
    removed = set(removed)
    qubits = [(r, c) for r in range(rows) for c in range(cols) if (r, c) not in removed]
    qid = {q: k for k, q in enumerate(qubits)}
    pats = {"A": [], "B": [], "C": [], "D": []}
    for r in range(rows - 1):
        for c in range(cols):
            if r % 2 == 0:
                dr, dl = (r + 1, c), (r + 1, c - 1)
            else:
                dr, dl = (r + 1, c + 1), (r + 1, c)
            for tgt, name in ((dr, "A" if r % 2 == 0 else "B"), (dl, "C" if r % 2 == 0 else "D")):
                if (r, c) in qid and tgt in qid:
                    pats[name].append((qid[(r, c)], qid[tgt]))
    return qubits, pats


def _fsim(theta, phi):
    c, s = math.cos(theta), math.sin(theta)
    u = np.array([[1, 0, 0, 0], [0, c, -1j * s, 0], [0, -1j * s, c, 0], [0, 0, 0, np.exp(-1j * phi)]], dtype=np.complex128)
    return u.reshape(2, 2, 2, 2)      # [out1, out2, in1, in2]


_SQ = {
    "X": np.array([[1, -1j], [-1j, 1]], dtype=np.complex128) / math.sqrt(2),
    "Y": np.array([[1, -1], [1, 1]], dtype=np.complex128) / math.sqrt(2),
    "W": np.array([[1, -np.sqrt(1j)], [np.sqrt(-1j), 1]], dtype=np.complex128) / math.sqrt(2),
}


def sycamore_circuit(rows=9, cols=6, cycles=14, seed=53, removed=((0, 3),), sequence="ABCDCDAB"):
    """Gate list of a Sycamore-like random circuit: per cycle a random sqrt(X|Y|W) on every qubit (never the
    same gate twice in a row on one qubit) then fSim(pi/2, pi/6) on the cycle's coupler pattern; a final layer of
    single-qubit gates.  Returns (nqubits, gates) with gates = [(matrix, (q,)) | (tensor4, (q1,q2))]."""
    rng = np.random.default_rng(seed)
    qubits, pats = sycamore_layout(rows, cols, removed)
    nq = len(qubits)
    last = [None] * nq
    gates = []
    fs = _fsim(math.pi / 2, math.pi / 6)

    def sq_layer():
        for q in range(nq):
            choices = [g for g in "XYW" if g != last[q]]
            g = choices[int(rng.integers(len(choices)))]
            last[q] = g
            gates.append((_SQ[g], (q,)))

    for cyc in range(cycles):
        sq_layer()
        for (a, b) in pats[sequence[cyc % len(sequence)]]:
            gates.append((fs, (a, b)))
    sq_layer()
    return nq, gates


def circuit_amplitude_network(nq, gates, bitstring: Sequence[int], dtype=np.complex64, simplify=True):
    """<b| C |0..0> as a closed tensor network.  With simplify=True rank-1/rank-2 tensors are absorbed into a
    neighbour on the host (the usual rank simplification), leaving one tensor per two-qubit gate."""
    cur = [("q", q, 0) for q in range(nq)]
    arrays, inds = [], []
    for q in range(nq):
        arrays.append(np.array([1.0, 0.0], dtype=np.complex128))
        inds.append((cur[q],))
    for (g, qs) in gates:
        if len(qs) == 1:
            q = qs[0]
            new = ("q", q, cur[q][2] + 1)
            arrays.append(g)
            inds.append((new, cur[q]))
            cur[q] = new
        else:
            a, b = qs
            na, nb = ("q", a, cur[a][2] + 1), ("q", b, cur[b][2] + 1)
            arrays.append(g)
            inds.append((na, nb, cur[a], cur[b]))
            cur[a], cur[b] = na, nb
    for q in range(nq):
        v = np.zeros(2, dtype=np.complex128)
        v[int(bitstring[q])] = 1.0
        arrays.append(v)
        inds.append((cur[q],))
    if simplify:
        arrays, inds = rank_simplify(arrays, inds)
    dt = np.dtype(dtype)
    return TensorNetwork([Tensor(np.ascontiguousarray(a).astype(dt), i) for a, i in zip(arrays, inds)])


def rank_simplify(arrays, inds, max_rank=2):
    """Absorb every tensor of rank <= max_rank into a neighbour (host-side network preparation, numpy)."""
    arrays = [np.asarray(a) for a in arrays]
    inds = [tuple(i) for i in inds]
    alive = [True] * len(arrays)
    owners: Dict[object, set] = {}
    for k, t in enumerate(inds):
        for i in t:
            owners.setdefault(i, set()).add(k)
    work = [k for k in range(len(arrays)) if len(inds[k]) <= max_rank]
    while work:
        k = work.pop()
        if not alive[k] or len(inds[k]) > max_rank:
            continue
        nb = None
        for i in inds[k]:
            for o in owners[i]:
                if o != k and alive[o]:
                    # prefer the neighbour whose rank grows least
                    if nb is None or len(inds[o]) < len(inds[nb]):
                        nb = o
        if nb is None:
            continue
        a, ai, b, bi = arrays[k], inds[k], arrays[nb], inds[nb]
        shared = [i for i in ai if i in bi]
        # sum shared indices only if nobody else carries them
        ssum = [i for i in shared if len(owners[i]) == 2]
        out = tuple(i for i in bi if i not in ssum) + tuple(i for i in ai if i not in bi)
        if len(out) > max(len(bi), max_rank):
            continue
        sym = {}
        for i in ai + bi:
            sym.setdefault(i, len(sym))
        c = np.einsum(a, [sym[i] for i in ai], b, [sym[i] for i in bi], [sym[i] for i in out])
        for i in ai:
            owners[i].discard(k)
        for i in bi:
            owners[i].discard(nb)
        alive[k] = False
        arrays[nb], inds[nb] = c, out
        for i in out:
            owners[i].add(nb)
        if len(out) <= max_rank:
            work.append(nb)
    keep = [k for k in range(len(arrays)) if alive[k]]
    return [arrays[k] for k in keep], [inds[k] for k in keep]


def sycamore_amplitude_network(rows=9, cols=6, cycles=14, seed=53, removed=((0, 3),), dtype=np.complex64):
    nq, gates = sycamore_circuit(rows, cols, cycles, seed, removed)
    rng = np.random.default_rng(seed + 1)
    bits = rng.integers(0, 2, nq)
    return circuit_amplitude_network(nq, gates, bits, dtype), (nq, gates, bits)


# ---------------------------------------------------------------------------------------------------
# cfg4: PEPS norm
# ---------------------------------------------------------------------------------------------------
def peps_norm_network(m=6, n=6, D=4, p=2, dtype=np.complex128, seed=4):
    rng = np.random.default_rng(seed)
    dt = np.dtype(dtype)
    arrays = []
    for i in range(m):
        row = []
        for j in range(n):
            shape = []
            if j > 0: shape.append(D)       # l
            if j < n - 1: shape.append(D)   # r
            if i > 0: shape.append(D)       # u
            if i < m - 1: shape.append(D)   # d
            shape.append(p)                 # o
            x = rng.standard_normal(shape)
            if dt.kind == "c":
                x = x + 1j * rng.standard_normal(shape)
            row.append((x / math.sqrt(2.0 * D * D * p / 2.0)).astype(dt))
        arrays.append(row)
    psi = PEPS(arrays)
    bra = _conj_reset(psi, "bra")
    tn = TensorNetwork()
    # interleave ket/bra per site: (site k) -> leaves 2k, 2k+1
    for a, b in zip(psi.tensors, bra.tensors):
        tn.append(a)
        tn.append(b)
    return tn, psi


def peps_boundary_path(m: int, n: int) -> ContractionPath:
    """Row-by-row boundary contraction of <psi|psi> on leaves interleaved (ket,bra) per site in row-major order:
    first fuse each site's ket*bra pair, then absorb sites column by column into a growing boundary tensor."""
    nl = 2 * m * n
    steps = []
    cur = nl
    site = {}
    for k in range(m * n):
        steps.append((2 * k, 2 * k + 1))
        site[k] = cur
        cur += 1
    env = site[0]
    for k in range(1, m * n):
        steps.append((env, site[k]))
        env = cur
        cur += 1
    return ContractionPath(steps)


# ---------------------------------------------------------------------------------------------------
# cfg5: <psi|H|psi>
# ---------------------------------------------------------------------------------------------------
def mps_mpo_expectation_network(n=100, chi=1024, h=1.0, J=1.0, dtype=np.complex128, seed=5, psi: Optional[MPS] = None):
    """Leaves ordered ket_1..ket_n, W_1..W_n, bra_1..bra_n (expect_network)."""
    psi = psi or MPS.rand(n, maxdim=chi, eltype=dtype, rng=seed)
    H = ising_1d_mpo(n, h, J)
    return expect_network(psi, H), psi, H


def sweep_path(n: int) -> ContractionPath:
    """Left-to-right environment sweep exactly as DMRG.jl:106-115: env <- ((env * psi_i) * W_i) * conj(psi_i)."""
    steps, nl = [], 3 * n
    cur = nl
    steps.append((0, n))               # psi_1 * W_1
    steps.append((cur, 2 * n))         # * conj(psi_1)
    cur += 1
    for i in range(1, n):
        steps.append((cur, i)); cur += 1
        steps.append((cur, n + i)); cur += 1
        steps.append((cur, 2 * n + i)); cur += 1
    return ContractionPath(steps)


def product_mps(bits: str, chi: int = 1, seed: int = 0, dtype=np.complex128) -> MPS:
    """K4 gauge trick: a product state written as an MPS whose bonds are padded to `chi` and
    scrambled by random invertible gauges G, G^-1 — dense tensors, value unchanged."""
    s2 = 1 / math.sqrt(2)
    table = {"0": [1.0, 0.0], "1": [0.0, 1.0], "+": [s2, s2], "-": [s2, -s2]}
    n = len(bits)
    rng = np.random.default_rng(seed)
    dims = [1] + [min(chi, 2 ** min(i, n - i, 30)) for i in range(1, n)] + [1]
    arrays = []
    Gprev_inv = np.eye(1)
    for i, ch in enumerate(bits):
        cl, cr = dims[i], dims[i + 1]
        if cr > 1:
            G = rng.standard_normal((cr, cr)) + 1j * rng.standard_normal((cr, cr))
            # ||G||_2 of a complex Ginibre matrix with E|g|^2 = 2 is ~ 2 sqrt(2 cr): G/that + 1 is well conditioned
            G = G / (2.0 * math.sqrt(2.0 * cr)) + np.eye(cr)
            Ginv = np.linalg.inv(G)
        else:
            G, Ginv = np.eye(1), np.eye(1)
        # the un-gauged site is e_0 (x) |ch> (x) e_0^T, so G_prev^-1 . site . G is an outer product
        a = Gprev_inv[:, 0][:, None, None] * np.asarray(table[ch])[None, :, None] * G[0, :][None, None, :]
        Gprev_inv = Ginv
        arrays.append(a.astype(dtype))
    arrays[0] = arrays[0].reshape(2, -1)
    arrays[-1] = arrays[-1].reshape(-1, 2)
    return MPS(arrays, order=("l", "o", "r"))
