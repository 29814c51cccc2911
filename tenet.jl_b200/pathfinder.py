"""Contraction-path and slice finder — a stand-in for EinExprs.jl (`einexpr(tn; optimizer=Greedy())`, slice
selection), which the reference names as its path optimiser (/root/reference/README.md:19-20,
docs/refs.bib:36-44 cites Gray & Kourtis hyper-optimisation) but does not vendor.

The path is an *input* of the hot path, so this is host-side Python: randomised greedy
(Boltzmann-sampled pair choice, cost = size(out) - alpha*(size(a)+size(b))) with restarts, plus a greedy
slice finder that removes the index giving the largest drop in peak intermediate size per unit of extra
FLOPs.  Index sets are Python-int bitmasks, sizes are log2 floats.
"""
from __future__ import annotations

import heapq
import math
import random
from dataclasses import dataclass, field
from typing import Dict, Hashable, List, Optional, Sequence, Tuple


@dataclass
class ContractionPath:
    """SSA path: step s contracts ids (i, j) into id n_leaves + s.  `sliced` are summed outside."""
    steps: List[Tuple[int, int]]
    sliced: Tuple[Hashable, ...] = ()
    output: Tuple[Hashable, ...] = ()
    log2_macs: float = 0.0          # per slice
    log2_max_size: float = 0.0      # largest intermediate (elements), per slice
    nslices: int = 1
    info: dict = field(default_factory=dict)

    @property
    def total_macs(self) -> float:
        return (2.0 ** self.log2_macs) * self.nslices


class _Net:
    def __init__(self, inputs: Sequence[Sequence[Hashable]], sizes: Dict[Hashable, int], output: Sequence[Hashable]):
        self.labels = []
        bit = {}
        for t in inputs:
            for i in t:
                if i not in bit:
                    bit[i] = len(self.labels)
                    self.labels.append(i)
        for i in output:
            if i not in bit:
                raise ValueError(f"output index {i!r} is not carried by any tensor")
        self.bit = bit
        self.lg = [math.log2(sizes[l]) for l in self.labels]
        self.masks = [sum(1 << bit[i] for i in t) for t in inputs]
        self.out = sum(1 << bit[i] for i in output)
        self.n = len(inputs)

    def lsize(self, mask: int, removed: int = 0) -> float:
        mask &= ~removed
        s = 0.0
        lg = self.lg
        while mask:
            low = mask & -mask
            s += lg[low.bit_length() - 1]
            mask ^= low
        return s


def _logaddexp2(a, b):
    if a < b:
        a, b = b, a
    return a + math.log2(1.0 + 2.0 ** (b - a))


def path_cost(net: _Net, steps, removed: int = 0):
    """(log2 MACs, log2 max intermediate size, list of per-step (log2 macs, log2 out size))."""
    masks = [m & ~removed for m in net.masks]
    # how many live tensors carry each index (+1 for output)
    count = {}
    for m in masks:
        mm = m
        while mm:
            low = mm & -mm
            count[low] = count.get(low, 0) + 1
            mm ^= low
    mm = net.out & ~removed
    while mm:
        low = mm & -mm
        count[low] = count.get(low, 0) + 1
        mm ^= low
    # count of appearances inside each node's subtree
    sub = [dict() for _ in masks]
    for k, m in enumerate(masks):
        mm = m
        while mm:
            low = mm & -mm
            sub[k][low] = 1
            mm ^= low
    tot, mx, per = -1e9, 0.0, []
    for (i, j) in steps:
        si = dict(sub[i])
        for k, v in sub[j].items():
            si[k] = si.get(k, 0) + v
        union = 0
        outm = 0
        so = {}
        for k, v in si.items():
            union |= k
            if v < count[k]:
                outm |= k
                so[k] = v
        lm = net.lsize(union)
        lo = net.lsize(outm)
        tot = _logaddexp2(tot, lm)
        mx = max(mx, lo)
        per.append((lm, lo))
        masks.append(outm)
        sub.append(so)
        sub[i] = sub[j] = None
    return tot, mx, per


def _greedy_once(net: _Net, rng: random.Random, temperature: float, alpha: float, removed: int = 0):
    """One randomised greedy pass.  Returns SSA steps."""
    masks = {k: (m & ~removed) for k, m in enumerate(net.masks)}
    out = net.out & ~removed
    # index -> set of live tensor ids
    owners: Dict[int, set] = {}
    for k, m in masks.items():
        mm = m
        while mm:
            low = mm & -mm
            owners.setdefault(low, set()).add(k)
            mm ^= low
    sizes = {k: net.lsize(m) for k, m in masks.items()}
    next_id = net.n
    steps = []
    heap = []
    tick = 0

    def result_mask(i, j):
        mi, mj = masks[i], masks[j]
        union = mi | mj
        shared = mi & mj
        keep = union & ~shared
        # shared indices survive if in output or carried by a third tensor
        mm = shared
        while mm:
            low = mm & -mm
            if (low & out) or len(owners[low]) > 2:
                keep |= low
            mm ^= low
        return keep

    def push(i, j):
        nonlocal tick
        rm = result_mask(i, j)
        so = net.lsize(rm)
        # cost in linear space, clipped for stability
        c = 2.0 ** min(so, 200.0) - alpha * (2.0 ** min(sizes[i], 200.0) + 2.0 ** min(sizes[j], 200.0))
        if temperature > 0:
            # Boltzmann noise on a log-compressed score (as in cotengra's greedy)
            g = -math.log(-math.log(rng.random() + 1e-300) + 1e-300)
            sc = math.copysign(math.log2(abs(c) + 1.0), c) - temperature * g
        else:
            sc = math.copysign(math.log2(abs(c) + 1.0), c)
        tick += 1
        heapq.heappush(heap, (sc, tick, i, j, rm))

    def neighbours(i):
        nb = set()
        mm = masks[i]
        while mm:
            low = mm & -mm
            nb |= owners[low]
            mm ^= low
        nb.discard(i)
        return nb

    live = set(masks.keys())
    for i in list(live):
        for j in neighbours(i):
            if i < j:
                push(i, j)
    while len(live) > 1:
        pair = None
        while heap:
            sc, _, i, j, rm = heapq.heappop(heap)
            if i in live and j in live:
                pair = (i, j, rm)
                break
        if pair is None:
            # disconnected components: outer products, smallest first
            rest = sorted(live, key=lambda k: sizes[k])
            i, j = rest[0], rest[1]
            pair = (i, j, result_mask(i, j))
        i, j, rm = pair
        rm = result_mask(i, j)
        k = next_id
        next_id += 1
        steps.append((i, j))
        for t in (i, j):
            mm = masks[t]
            while mm:
                low = mm & -mm
                owners[low].discard(t)
                mm ^= low
            live.discard(t)
        masks[k] = rm
        sizes[k] = net.lsize(rm)
        mm = rm
        while mm:
            low = mm & -mm
            owners.setdefault(low, set()).add(k)
            mm ^= low
        live.add(k)
        for t in neighbours(k):
            push(min(k, t), max(k, t))
        del masks[i], masks[j]
    return steps


def _score(lm, ls, max_log2_size, minimize):
    pen = max(0.0, ls - max_log2_size) if max_log2_size is not None else 0.0
    if minimize == "size":
        return ls + 1e-3 * lm
    if minimize == "combo":
        return _logaddexp2(lm, ls + 6.0) + 4.0 * pen   # flops + 64 * write traffic
    return lm + 4.0 * pen


def optimize_path(inputs, sizes, output=(), ntrials: int = 32, seed: int = 0, minimize: str = "flops",
                  max_log2_size: Optional[float] = None, removed_inds: Sequence[Hashable] = ()) -> ContractionPath:
    """Randomised greedy with restarts (`einexpr(tn; optimizer=Greedy())` + a hyper-search over temperature/alpha)."""
    net = _Net(inputs, sizes, output)
    removed = sum(1 << net.bit[i] for i in removed_inds if i in net.bit)
    rng = random.Random(seed)
    best = None
    if net.n == 1:
        return ContractionPath([], tuple(removed_inds), tuple(output), 0.0, net.lsize(net.masks[0]), 1)
    for trial in range(max(1, ntrials)):
        if trial == 0:
            temp, alpha = 0.0, 1.0
        else:
            temp = rng.choice([0.0, 0.05, 0.1, 0.3, 0.6, 1.0])
            alpha = rng.choice([0.0, 0.5, 1.0, 1.0, 1.5, 2.0])
        steps = _greedy_once(net, rng, temp, alpha, removed)
        lm, ls, _ = path_cost(net, steps, removed)
        sc = _score(lm, ls, max_log2_size, minimize)
        if best is None or sc < best[0]:
            best = (sc, steps, lm, ls, {"temperature": temp, "alpha": alpha, "trial": trial})
    _, steps, lm, ls, info = best
    ns = 1
    for i in removed_inds:
        ns *= sizes[i]
    return ContractionPath(steps, tuple(removed_inds), tuple(output), lm, ls, ns, info)


def find_slices(inputs, sizes, output, path: ContractionPath, target_log2_size: float,
                max_slices_log2: float = 40.0, reoptimize_trials: int = 0, seed: int = 0) -> ContractionPath:
    """Greedy slice finder: repeatedly slice the index (never an output index) that most reduces the peak
    intermediate size, ties broken by the smallest total-FLOP overhead, until the peak <= target."""
    net = _Net(inputs, sizes, output)
    removed = sum(1 << net.bit[i] for i in path.sliced if i in net.bit)
    chosen = list(path.sliced)
    steps = path.steps
    lm, ls, per = path_cost(net, steps, removed)
    lg_slices = sum(math.log2(sizes[i]) for i in chosen)
    while ls > target_log2_size and lg_slices < max_slices_log2:
        # candidate indices: those present in the largest intermediates
        masks = [m & ~removed for m in net.masks]
        cand = 0
        # recompute intermediates' masks cheaply through path_cost's logic
        cand_masks = _intermediate_masks(net, steps, removed)
        thr = ls - 1e-9
        for m in cand_masks:
            if net.lsize(m) >= thr:
                cand |= m
        cand &= ~(net.out)
        if not cand:
            break
        best = None
        mm = cand
        while mm:
            low = mm & -mm
            mm ^= low
            r2 = removed | low
            lm2, ls2, _ = path_cost(net, steps, r2)
            over = lm2 + net.lg[low.bit_length() - 1]   # total flops over all slices
            key = (ls2, over)
            if best is None or key < best[0]:
                best = (key, low, lm2, ls2)
        _, low, lm, ls = best
        removed |= low
        lab = net.labels[low.bit_length() - 1]
        chosen.append(lab)
        lg_slices += math.log2(sizes[lab])
        if reoptimize_trials > 0:
            p2 = optimize_path(inputs, sizes, output, ntrials=reoptimize_trials, seed=seed + len(chosen),
                               removed_inds=chosen)
            if p2.log2_macs < lm or p2.log2_max_size < ls:
                steps, lm, ls = p2.steps, p2.log2_macs, p2.log2_max_size
    ns = 1
    for i in chosen:
        ns *= sizes[i]
    return ContractionPath(list(steps), tuple(chosen), tuple(output), lm, ls, ns, dict(path.info))


def _intermediate_masks(net: _Net, steps, removed: int):
    masks = [m & ~removed for m in net.masks]
    count = {}
    for m in masks + [net.out & ~removed]:
        mm = m
        while mm:
            low = mm & -mm
            count[low] = count.get(low, 0) + 1
            mm ^= low
    sub = []
    for m in masks:
        d = {}
        mm = m
        while mm:
            low = mm & -mm
            d[low] = 1
            mm ^= low
        sub.append(d)
    outs = []
    for (i, j) in steps:
        si = dict(sub[i])
        for k, v in sub[j].items():
            si[k] = si.get(k, 0) + v
        so = {k: v for k, v in si.items() if v < count[k]}
        m = 0
        for k in so:
            m |= k
        outs.append(m)
        sub.append(so)
        sub[i] = sub[j] = None
    return outs


def linear_path(n: int) -> List[Tuple[int, int]]:
    """((0,1),2),3)... — the zipper used by overlap(::MPS, ::MPS) (overlap.jl:36-50) when leaves are interleaved."""
    steps, cur = [], 0
    for k in range(1, n):
        steps.append((cur, k))
        cur = n + k - 1
    return steps


def search(inputs, sizes, output=(), ntrials: int = 64, seed: int = 0, target_log2_size: Optional[float] = None,
           minimize: str = "flops", slice_reopt_trials: int = 0) -> ContractionPath:
    """Path search + slicing in one call: randomised greedy restarts, then (if the peak intermediate exceeds
    `target_log2_size`) the greedy slice finder."""
    p = optimize_path(inputs, sizes, output, ntrials=ntrials, seed=seed, minimize=minimize)
    if target_log2_size is not None and p.log2_max_size > target_log2_size:
        p = find_slices(inputs, sizes, output, p, target_log2_size, reoptimize_trials=slice_reopt_trials, seed=seed)
    return p


def load_path(tn_inds_all, filename) -> ContractionPath:
    """Read a bench_paths/*.json file; `tn_inds_all` = tn.inds("all") maps the stored integer labels back."""
    import json
    with open(filename) as f:
        d = json.load(f)
    labels = list(tn_inds_all)
    sliced = tuple(labels[k] for k in d["sliced"])
    return ContractionPath([tuple(s) for s in d["steps"]], sliced, (), d["log2_macs_per_slice"], d["log2_max_size"],
                           int(round(2 ** d["nslices_log2"])), d.get("search", {}))
