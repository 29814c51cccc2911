"""CPU ORACLE — test infrastructure only.  Nothing under oracle/ is imported by the product package
(`tenet.jl_b200/`); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import it, and only as the checker / the reported CPU baseline.

What it restates
----------------
The hot path of bsc-quantic/Tenet.jl v0.10.3: `contract(tn; path)` executed as a chain of pairwise
`Muscle.binary_einsum` calls, summed over sliced indices.  The arithmetic of that path is NOT in
/root/reference: `binary_einsum` is Muscle.jl (compat 0.3.9), `contract` is Tangles.jl (compat 0.2.7), the path
is EinExprs.jl — un-vendored registry packages (/root/reference/Project.toml:6-18,29-45), and there is no Julia in
the build container.  So this file restates the *published algorithm* of the default backend
(permutedims both operands to (free, batch, K)/(K, batch, free) -> reshape -> BLAS gemm -> reshape) in numpy on
OpenBLAS — the same BLAS family Julia links — and anchors it on the reference's own call sites:

  * binary_einsum semantics (i)-(vii):   src/Operations/overlap.jl:42,46 (contract all shared inds);
    src/Operations/canonize.jl:44, absorb.jl:31, evolve.jl:76,108 (dims=Index[] keeps shared inds as batch);
    src/Algorithms/DMRG.jl:60-61,75-78 (rank-0 operand / outer product); overlap.jl:49 (rank-0 result);
    overlap.jl:7,39 (operands pre-conjugated).
  * contract(tn) tree walk:              src/Operations/overlap.jl:5-13, test/unit/mps.jl:89.
  * slicing = sum over fixed index values: README.md:20; tensor-level views compress.jl:46-58.

Parity pins (see tests/test_oracle_pins.py and tests/golden/): P1 X-gate KAT (test/unit/simple_update.jl:4-14),
P4/K3 TFIM product-state energy -1.1902477482849715 (test/unit/dmrg.jl:10-13), K1 <psi|psi> = 1 for rand(MPS)
(src/Components/MPS.jl:103-104,154-157), K2 <0..0|H|0..0> = -J(n-1), <+..+|H|+..+> = -h n (src/Models/Ising.jl:12-30).
Beyond those the upstream-internal details (default output order, error types) are PARITY UNPINNED.
"""
from __future__ import annotations

import itertools
from typing import Dict, Hashable, Iterable, List, Optional, Sequence, Tuple

import numpy as np

Ind = Hashable


def binary_einsum(a: np.ndarray, a_inds: Sequence[Ind], b: np.ndarray, b_inds: Sequence[Ind],
                  dims: Optional[Iterable[Ind]] = None,
                  out_inds: Optional[Sequence[Ind]] = None) -> Tuple[np.ndarray, Tuple[Ind, ...]]:
    """C[free(A), free(B), batch] = sum_{dims} A * B   (Muscle.binary_einsum restated).

    dims=None contracts every shared index (overlap.jl:42); dims=() keeps shared indices as batch/Hadamard
    indices (canonize.jl:44).  An index carried by only one operand and listed in `dims` is summed out.
    Default output order: free(A) in A's order, free(B) in B's order, batch in A's order [UPSTREAM-RECALL].
    """
    a_inds = tuple(a_inds)
    b_inds = tuple(b_inds)
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.ndim == len(a_inds) and b.ndim == len(b_inds), "rank / index-count mismatch"
    assert len(set(a_inds)) == len(a_inds) and len(set(b_inds)) == len(b_inds), "repeated index in one tensor"
    shared = [i for i in a_inds if i in b_inds]
    for i in shared:
        assert a.shape[a_inds.index(i)] == b.shape[b_inds.index(i)], f"extent mismatch on {i!r}"
    dims = set(shared) if dims is None else set(dims)
    batch = [i for i in shared if i not in dims]
    k_ab = [i for i in shared if i in dims]
    k_a = [i for i in a_inds if i in dims and i not in b_inds]      # summed, carried by A only
    k_b = [i for i in b_inds if i in dims and i not in a_inds]
    free_a = [i for i in a_inds if i not in b_inds and i not in dims]
    free_b = [i for i in b_inds if i not in a_inds and i not in dims]

    if k_a:
        a = a.sum(axis=tuple(a_inds.index(i) for i in k_a))
        a_inds = tuple(i for i in a_inds if i not in k_a)
    if k_b:
        b = b.sum(axis=tuple(b_inds.index(i) for i in k_b))
        b_inds = tuple(i for i in b_inds if i not in k_b)

    ext = {i: a.shape[a_inds.index(i)] for i in a_inds}
    ext.update({i: b.shape[b_inds.index(i)] for i in b_inds})
    prod = lambda inds: int(np.prod([ext[i] for i in inds], dtype=np.int64)) if inds else 1
    L, M, N, K = prod(batch), prod(free_a), prod(free_b), prod(k_ab)
    # permutedims -> reshape -> gemm -> reshape
    at = np.transpose(a, [a_inds.index(i) for i in batch + free_a + k_ab]).reshape(L, M, K)
    bt = np.transpose(b, [b_inds.index(i) for i in batch + k_ab + free_b]).reshape(L, K, N)
    ct = np.matmul(at, bt)                                             # OpenBLAS gemm per batch entry
    c_inds = tuple(batch + free_a + free_b)
    c = ct.reshape([ext[i] for i in c_inds])
    default = tuple(free_a + free_b + batch)
    want = tuple(out_inds) if out_inds is not None else default
    assert set(want) == set(c_inds) and len(want) == len(c_inds), "out_inds must be a permutation of the result indices"
    c = np.transpose(c, [c_inds.index(i) for i in want])
    return c, want


def result_inds(tensors_inds: Sequence[Sequence[Ind]], output: Optional[Sequence[Ind]] = None) -> Tuple[Ind, ...]:
    """Open indices of a network = indices carried by exactly one tensor (Tangles `inds(tn; set=:open)`)."""
    if output is not None:
        return tuple(output)
    count: Dict[Ind, int] = {}
    for inds in tensors_inds:
        for i in inds:
            count[i] = count.get(i, 0) + 1
    seen, out = set(), []
    for inds in tensors_inds:
        for i in inds:
            if count[i] == 1 and i not in seen:
                seen.add(i)
                out.append(i)
    return tuple(out)


def contract_path(arrays: Sequence[np.ndarray], inds: Sequence[Sequence[Ind]], steps: Sequence[Tuple[int, int]],
                  output: Optional[Sequence[Ind]] = None) -> Tuple[np.ndarray, Tuple[Ind, ...]]:
    """Tangles.contract(tn; path) restated: SSA path, step s contracts ids (i, j) into id n+s.

    An index is summed at the step after which no tensor outside the pair's subtree (and not the output)
    carries it — EinExprs' `suminds`.  Indices carried by more than two tensors stay as batch indices until then.
    """
    n = len(arrays)
    output = result_inds(inds, output)
    total: Dict[Ind, int] = {}
    for t in inds:
        for i in t:
            total[i] = total.get(i, 0) + 1
    for i in output:
        total[i] = total.get(i, 0) + 1
    vals: List[Optional[np.ndarray]] = [np.asarray(a) for a in arrays]
    vinds: List[Optional[Tuple[Ind, ...]]] = [tuple(t) for t in inds]
    cnt: List[Optional[Dict[Ind, int]]] = [{i: 1 for i in t} for t in inds]
    if len(steps) == 0:
        assert n == 1
        a, ai = vals[0], vinds[0]
        drop = [i for i in ai if i not in output]
        if drop:
            a = a.sum(axis=tuple(ai.index(i) for i in drop))
            ai = tuple(i for i in ai if i not in drop)
        return np.transpose(a, [ai.index(i) for i in output]), tuple(output)
    for s, (i, j) in enumerate(steps):
        ci = dict(cnt[i])
        for k, v in cnt[j].items():
            ci[k] = ci.get(k, 0) + v
        dims = [k for k, v in ci.items() if v == total[k]]
        last = s == len(steps) - 1
        c, c_inds = binary_einsum(vals[i], vinds[i], vals[j], vinds[j], dims=dims,
                                  out_inds=output if last else None)
        vals.append(c)
        vinds.append(c_inds)
        cnt.append({k: v for k, v in ci.items() if v < total[k]})
        vals[i] = vals[j] = None
        vinds[i] = vinds[j] = None
        cnt[i] = cnt[j] = None
    return vals[-1], vinds[-1]


def slice_network(arrays: Sequence[np.ndarray], inds: Sequence[Sequence[Ind]], fixed: Dict[Ind, int]):
    """`view(tn, ind => value ...)` restated: fix indices, dropping them from every tensor that carries them."""
    out_a, out_i = [], []
    for a, t in zip(arrays, inds):
        t = tuple(t)
        sel = tuple(fixed[i] if i in fixed else slice(None) for i in t)
        out_a.append(np.asarray(a)[sel])
        out_i.append(tuple(i for i in t if i not in fixed))
    return out_a, out_i


def contract_sliced(arrays, inds, steps, sliced: Sequence[Ind], output=None, slice_ids: Optional[Iterable[int]] = None):
    """Sum of per-slice contractions.  Slice id enumerates `sliced` in mixed radix, sliced[0] fastest — the same
    order the engine uses, so partial sums over the same slice ids agree term by term."""
    output = result_inds([tuple(i for i in t if i not in sliced) for t in inds], output)
    ext = {}
    for a, t in zip(arrays, inds):
        for ax, i in enumerate(t):
            ext[i] = np.asarray(a).shape[ax]
    sizes = [ext[i] for i in sliced]
    nslices = int(np.prod(sizes, dtype=np.int64)) if sliced else 1
    ids = range(nslices) if slice_ids is None else slice_ids
    acc = None
    for sid in ids:
        t, fixed = sid, {}
        for i, d in zip(sliced, sizes):
            fixed[i] = t % d
            t //= d
        sa, si = slice_network(arrays, inds, fixed)
        c, ci = contract_path(sa, si, steps, output)
        acc = c.copy() if acc is None else acc + c
    if acc is None:
        shape = [ext[i] for i in output]
        acc = np.zeros(shape, dtype=np.result_type(*[np.asarray(a).dtype for a in arrays]))
    return acc, tuple(output)


def path_flops(inds: Sequence[Sequence[Ind]], ext: Dict[Ind, int], steps, output=None, sliced: Sequence[Ind] = ()):
    """EinExprs-style cost of a path: MACs = prod of extents of all distinct indices of the pair, per step
    (SURVEY §8a N2); returns (total MACs per slice, max intermediate elements, total elements moved)."""
    sl = set(sliced)
    tinds = [tuple(i for i in t if i not in sl) for t in inds]
    output = result_inds(tinds, output)
    total: Dict[Ind, int] = {}
    for t in tinds:
        for i in t:
            total[i] = total.get(i, 0) + 1
    for i in output:
        total[i] = total.get(i, 0) + 1
    cnt = [{i: 1 for i in t} for t in tinds]
    macs, mx, moved = 0, 0, 0
    size = lambda s: int(np.prod([ext[i] for i in s], dtype=object)) if s else 1
    for (i, j) in steps:
        ci = dict(cnt[i])
        for k, v in cnt[j].items():
            ci[k] = ci.get(k, 0) + v
        macs += size(ci.keys())
        co = {k: v for k, v in ci.items() if v < total[k]}
        mx = max(mx, size(co.keys()))
        moved += size(cnt[i].keys()) + size(cnt[j].keys()) + size(co.keys())
        cnt.append(co)
    return macs, mx, moved
