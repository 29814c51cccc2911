"""Contraction-tree refinement: subtree reconfiguration (exhaustive DP on small subtrees) and slicing-aware search.

Stand-in for what EinExprs' / hyper-optimisers' tree-level passes do for the reference (README.md:19-20,
docs/refs.bib:36-44); host-side, an *input* of the hot path.  Works on the bitmask network of pathfinder._Net.

A tree is stored as children[node] = (left, right) for internal nodes (ids >= n_leaves) in SSA numbering.
Reconfiguration: take an internal node, expand it downwards into at most `size` sub-trees (frontier), find the
optimal pairwise order of those frontier tensors by dynamic programming over subsets (cost = sum of 2^|union legs|,
tie-broken by peak size), splice the better order back.  Legs of any subset are those indices that also live
outside the subset (rest of the network, output) — computed from per-index carrier counts.
"""
from __future__ import annotations

import math
import random
from typing import Dict, List, Optional, Sequence, Tuple

from .pathfinder import ContractionPath, _Net, _greedy_once, _logaddexp2, path_cost


# "time" objective: a pairwise step costs max(MACs, BYTE_WEIGHT * elements moved) — on the B200 engine a complex64
# element moved through HBM (8 B at ~3.3 TB/s achieved) costs as much time as ~73 complex MACs (8 flop at ~240 TFLOP/s).
BYTE_WEIGHT_LOG2 = math.log2(float(__import__('os').environ.get('TNB_BYTE_WEIGHT', '73')))


def _step_cost(net, la_mask, lb_mask, lc_mask, minimize):
    lm = net.lsize(la_mask | lb_mask)
    if minimize != "time":
        return lm
    sa, sb, sc = net.lsize(la_mask), net.lsize(lb_mask), net.lsize(lc_mask)
    moved = _logaddexp2(_logaddexp2(sa, sb), sc) + BYTE_WEIGHT_LOG2
    return max(lm, moved)


def _popbits(mask: int):
    while mask:
        low = mask & -mask
        yield low
        mask ^= low


class _Tree:
    def __init__(self, net: _Net, steps: Sequence[Tuple[int, int]], removed: int):
        self.net = net
        self.removed = removed
        self.n = net.n
        self.children: Dict[int, Tuple[int, int]] = {}
        for s, (i, j) in enumerate(steps):
            self.children[self.n + s] = (i, j)
        self.root = self.n + len(steps) - 1
        self.leafmask = [m & ~removed for m in net.masks]
        self.out = net.out & ~removed
        # total carriers of each index over leaves (+1 for output)
        self.total: Dict[int, int] = {}
        for m in self.leafmask + [self.out]:
            for b in _popbits(m):
                self.total[b] = self.total.get(b, 0) + 1
        self._cache: Dict[int, Tuple[int, Dict[int, int]]] = {}

    def sub_counts(self, node) -> Dict[int, int]:
        """index -> number of leaves inside the subtree of `node` carrying it"""
        if node < self.n:
            return {b: 1 for b in _popbits(self.leafmask[node])}
        c = self._cache.get(node)
        if c is not None:
            return c[1]
        l, r = self.children[node]
        cl = dict(self.sub_counts(l))
        for k, v in self.sub_counts(r).items():
            cl[k] = cl.get(k, 0) + v
        self._cache[node] = (0, cl)
        return cl

    def legs(self, counts: Dict[int, int]) -> int:
        m = 0
        for k, v in counts.items():
            if v < self.total[k]:
                m |= k
        return m

    def invalidate(self):
        self._cache.clear()

    def ssa(self) -> List[Tuple[int, int]]:
        """post-order SSA steps of the current tree"""
        steps: List[Tuple[int, int]] = []
        newid: Dict[int, int] = {}
        nxt = self.n
        stack = [(self.root, False)]
        while stack:
            node, done = stack.pop()
            if node < self.n:
                newid[node] = node
                continue
            if done:
                l, r = self.children[node]
                steps.append((newid[l], newid[r]))
                newid[node] = nxt
                nxt += 1
            else:
                stack.append((node, True))
                l, r = self.children[node]
                stack.append((r, False))
                stack.append((l, False))
        return steps


def _dp_optimal(tree: _Tree, items: List[int], minimize: str):
    """optimal pairwise order of the frontier `items`; returns (cost_log2, size_log2, nested tuple order)"""
    net = tree.net
    k = len(items)
    counts = [tree.sub_counts(i) for i in items]
    full = (1 << k) - 1
    # carrier counts of every subset, built incrementally
    sub: Dict[int, Dict[int, int]] = {}
    for a in range(k):
        sub[1 << a] = counts[a]
    legs: Dict[int, int] = {}
    lsz: Dict[int, float] = {}
    best: Dict[int, Tuple[float, float, object]] = {}
    for a in range(k):
        m = tree.legs(counts[a])
        legs[1 << a] = m
        lsz[1 << a] = net.lsize(m)
        best[1 << a] = (-1e9, 0.0, items[a])
    for S in range(1, full + 1):
        if S & (S - 1) == 0:
            continue
        low = S & -S
        rest = S ^ low
        c = dict(sub[rest])
        for kk, v in sub[low].items():
            c[kk] = c.get(kk, 0) + v
        sub[S] = c
        lm = tree.legs(c)
        legs[S] = lm
        lsz[S] = net.lsize(lm)
        bc = None
        # enumerate splits S = S1 | S2 with low in S1
        T = rest
        S2 = T
        while True:
            S1 = S ^ S2
            if S2 and S1:
                c1, c2 = best[S1], best[S2]
                step = _step_cost(net, legs[S1], legs[S2], legs[S], minimize)
                cost = _logaddexp2(_logaddexp2(c1[0], c2[0]), step)
                size = max(c1[1], c2[1], lsz[S])
                key = (cost, size) if minimize in ("flops", "time") else (max(size, 0.0), cost)
                if bc is None or key < bc[0]:
                    bc = (key, cost, size, (c1[2], c2[2]))
            if S2 == 0:
                break
            S2 = (S2 - 1) & T
        best[S] = (bc[1], bc[2], bc[3])
    return best[full]


def _current_cost(tree: _Tree, node: int, frontier: set, minimize: str = "flops"):
    """cost / peak size of the part of the tree between `node` and the frontier"""
    net = tree.net
    if node in frontier:
        return -1e9, 0.0
    l, r = tree.children[node]
    cl, sl = _current_cost(tree, l, frontier, minimize)
    cr, sr = _current_cost(tree, r, frontier, minimize)
    ll = tree.legs(tree.sub_counts(l))
    lr = tree.legs(tree.sub_counts(r))
    lo = tree.legs(tree.sub_counts(node))
    step = _step_cost(net, ll, lr, lo, minimize)
    out = net.lsize(lo)
    return _logaddexp2(_logaddexp2(cl, cr), step), max(sl, sr, out)


def subtree_reconfigure(net: _Net, steps, removed: int = 0, size: int = 8, rounds: int = 2, minimize: str = "flops",
                        rng: Optional[random.Random] = None):
    """Greedy descent: for every internal node (largest first) re-solve its `size`-frontier optimally."""
    tree = _Tree(net, steps, removed)
    rng = rng or random.Random(0)
    next_id = [max(tree.children) + 1 if tree.children else net.n]
    for _ in range(rounds):
        improved = False
        nodes = sorted(tree.children.keys(), key=lambda x: -net.lsize(tree.legs(tree.sub_counts(x))))
        for node in nodes:
            if node not in tree.children:
                continue
            # expand frontier: repeatedly open the internal frontier node with the largest output
            frontier = [node]
            while True:
                cand = [f for f in frontier if f >= tree.n and f in tree.children]
                if not cand or len(frontier) >= size:
                    break
                f = max(cand, key=lambda x: net.lsize(tree.legs(tree.sub_counts(x))))
                frontier.remove(f)
                frontier.extend(tree.children[f])
            if len(frontier) < 3:
                continue
            fs = set(frontier)
            cur_c, cur_s = _current_cost(tree, node, fs, minimize)
            new_c, new_s, order = _dp_optimal(tree, frontier, minimize)
            better = (new_c < cur_c - 1e-9) if minimize in ("flops", "time") else ((new_s, new_c) < (cur_s - 1e-9, cur_c))
            if minimize in ("flops", "time") and new_c <= cur_c + 1e-9 and new_s < cur_s - 1e-9:
                better = True
            if not better:
                continue
            # splice: delete old internal nodes between node and frontier, build new ones; keep `node` as the root id
            def drop(x):
                if x in fs:
                    return
                l, r = tree.children.pop(x)
                drop(l); drop(r)
            drop(node)

            def build(o, top):
                if not isinstance(o, tuple):
                    return o
                l = build(o[0], False)
                r = build(o[1], False)
                if top:
                    nid = node
                else:
                    nid = next_id[0]
                    next_id[0] += 1
                tree.children[nid] = (l, r)
                return nid
            build(order, True)
            tree.invalidate()
            improved = True
        if not improved:
            break
    return tree.ssa()


def tree_time_cost(net: _Net, steps, removed: int = 0) -> float:
    """log2 of sum over steps of max(MACs, BYTE_WEIGHT * elements moved) — per slice"""
    tree = _Tree(net, steps, removed)
    tot = -1e9
    for node, (l, r) in tree.children.items():
        ll = tree.legs(tree.sub_counts(l)); lr = tree.legs(tree.sub_counts(r)); lo = tree.legs(tree.sub_counts(node))
        tot = _logaddexp2(tot, _step_cost(net, ll, lr, lo, "time"))
    return tot


def slice_with_reconfiguration(net: _Net, inputs, sizes, output, steps, target_log2_size: float, reconf_size: int,
                               minimize: str, rng, max_slices_log2: float = 40.0):
    """Interleaved slicing (the `slicing_reconf` idea of hyper-optimisers): slice ONE index — the one, among the indices
    of the largest intermediates, whose removal gives the lowest total cost over all slices (ties: smaller peak) — then
    re-optimise the tree for the sliced network with one round of subtree reconfiguration, and repeat until the peak
    fits.  Returns (steps, chosen labels, removed mask)."""
    from .pathfinder import _intermediate_masks
    removed, chosen, lg_slices = 0, [], 0.0
    lm, ls, _ = path_cost(net, steps, removed)

    def total_cost(st, rem, lgs):
        return (tree_time_cost(net, st, rem) if minimize == "time" else path_cost(net, st, rem)[0]) + lgs

    while ls > target_log2_size and lg_slices < max_slices_log2:
        cand = 0
        for m in _intermediate_masks(net, steps, removed):
            if net.lsize(m) >= ls - 1.0 - 1e-9:          # indices of the (near-)largest intermediates
                cand |= m
        cand &= ~net.out
        if not cand:
            break
        best = None
        for low in _popbits(cand):
            r2 = removed | low
            _, ls2, _ = path_cost(net, steps, r2)
            key = (total_cost(steps, r2, lg_slices + net.lg[low.bit_length() - 1]), ls2)
            if best is None or key < best[0]:
                best = (key, low)
        low = best[1]
        removed |= low
        lab = net.labels[low.bit_length() - 1]
        chosen.append(lab)
        lg_slices += math.log2(sizes[lab])
        s2 = subtree_reconfigure(net, steps, removed, reconf_size, 1, minimize, rng)
        _, ls2, _ = path_cost(net, s2, removed)
        _, ls1, _ = path_cost(net, steps, removed)
        if ls2 <= max(ls1, target_log2_size) + 1e-9 and total_cost(s2, removed, lg_slices) <= total_cost(steps, removed, lg_slices):
            steps = s2
        lm, ls, _ = path_cost(net, steps, removed)
    return steps, chosen, removed


def hyper_search(inputs, sizes, output=(), ntrials: int = 64, seed: int = 0, target_log2_size: Optional[float] = None,
                 reconf_size: int = 8, reconf_rounds: int = 2, keep: int = 4, verbose: bool = False,
                 minimize: str = "flops", slicing: str = "greedy", prescreen: int = 0) -> ContractionPath:
    """Randomised greedy restarts -> subtree reconfiguration of the best few -> greedy slicing with
    reconfiguration of the sliced tree (slicing="greedy"), or slicing interleaved with reconfiguration
    (slicing="interleaved").  Returns the best (total MACs over all slices) path found."""
    from .pathfinder import find_slices
    net = _Net(inputs, sizes, output)
    rng = random.Random(seed)
    cands = []
    for trial in range(max(1, ntrials)):
        temp = 0.0 if trial == 0 else rng.choice([0.0, 0.05, 0.1, 0.3, 0.6, 1.0])
        alpha = 1.0 if trial == 0 else rng.choice([0.0, 0.5, 1.0, 1.0, 1.5, 2.0])
        steps = _greedy_once(net, rng, temp, alpha, 0)
        lm, ls, _ = path_cost(net, steps, 0)
        cands.append((lm, ls, steps))
    cands.sort(key=lambda c: (c[0], c[1]))
    pre_rounds = 0
    if prescreen > keep:
        # the greedy cost is a poor predictor of where reconfiguration ends up (the committed Sycamore tree started as
        # the 200th-best greedy tree: 2^59 -> 2^47.8): give MANY candidates one cheap round, keep the best few after it
        pre = []
        for lm, ls, steps in cands[:prescreen]:
            s1 = subtree_reconfigure(net, steps, 0, reconf_size, 1, minimize, rng)
            c1 = tree_time_cost(net, s1, 0) if minimize == "time" else path_cost(net, s1, 0)[0]
            l1, z1, _ = path_cost(net, s1, 0)
            pre.append((c1, l1, z1, s1))
        pre.sort(key=lambda c: c[0])
        if verbose:
            print("  prescreen (1 round): " + " ".join(f"{c[0]:.1f}" for c in pre[:12]))
        cands = [(c[1], c[2], c[3]) for c in pre]
        pre_rounds = 1
    best = None
    for lm, ls, steps in cands[:keep]:
        s2 = subtree_reconfigure(net, steps, 0, reconf_size, max(1, reconf_rounds - pre_rounds), minimize, rng)
        lm2, ls2, _ = path_cost(net, s2, 0)
        if verbose:
            print(f"  greedy 2^{lm:.2f}/2^{ls:.0f} -> reconf 2^{lm2:.2f}/2^{ls2:.0f}")
        p = ContractionPath(s2, (), tuple(output), lm2, ls2, 1, {})
        if target_log2_size is not None and ls2 > target_log2_size:
            if slicing == "interleaved":
                st, chosen, removed = slice_with_reconfiguration(net, inputs, sizes, output, s2, target_log2_size,
                                                                 reconf_size, minimize, rng)
                lmi, lsi, _ = path_cost(net, st, removed)
                ns = 1
                for i in chosen:
                    ns *= sizes[i]
                p = ContractionPath(list(st), tuple(chosen), tuple(output), lmi, lsi, ns, {})
            else:
                p = find_slices(inputs, sizes, output, p, target_log2_size)
            removed = sum(1 << net.bit[i] for i in p.sliced)
            s3 = subtree_reconfigure(net, p.steps, removed, reconf_size, reconf_rounds, minimize, rng)
            lm3, ls3, _ = path_cost(net, s3, removed)
            ok3 = lm3 < p.log2_macs if minimize != "time" else tree_time_cost(net, s3, removed) < tree_time_cost(net, p.steps, removed)
            if ls3 <= target_log2_size + 1e-9 and ok3:
                p = ContractionPath(s3, p.sliced, tuple(output), lm3, ls3, p.nslices, {})
            if verbose:
                print(f"    sliced x2^{math.log2(p.nslices):.0f}: per-slice 2^{p.log2_macs:.2f}, total 2^{p.log2_macs + math.log2(p.nslices):.2f}")
        if minimize == "time":
            removed = sum(1 << net.bit[i] for i in p.sliced)
            total = tree_time_cost(net, p.steps, removed) + math.log2(p.nslices)
            if verbose:
                print(f"    time-model cost 2^{total:.2f} MAC-equivalents")
        else:
            total = p.log2_macs + math.log2(p.nslices)
        if best is None or total < best[0]:
            best = (total, p)
    p = best[1]
    p.info = {"method": "greedy+subtree-reconf", "trials": ntrials, "reconf_size": reconf_size, "seed": seed,
              "minimize": minimize, "score_log2": best[0]}
    return p
