"""`TensorNetwork`, `einexpr`, `contract` — host-side mirror of the Tangles.jl / EinExprs.jl interface the
reference drives its hot path through (`contract(tn; path, optimizer)`, /root/reference/src/Operations/overlap.jl:5-13;
17+17 assertions in test/unit/mps.jl / mpo.jl; test/integration/itensormps.jl:45).

`contract` hands the whole path (plus sliced indices) to libtnb200's `tnb_plan_*` in ONE call: the library
plans every pairwise step, keeps all intermediates device-resident and sums the slices into the output.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Hashable, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import check, make_desc, tnb_plan_info, tnb_step_info, tnb_tensor
from .context import B200Array, Context, default_context
from .pathfinder import ContractionPath, find_slices, optimize_path
from .tensor import Tensor, _promote_dtype


class TensorNetwork:
    """`GenericTensorNetwork`: a bag of tensors whose shared labels are the contracted (inner) indices."""

    def __init__(self, tensors: Iterable[Tensor] = ()):
        self.tensors: List[Tensor] = list(tensors)

    # -- Tangles-style accessors ------------------------------------------------------------------
    def append(self, t):
        if isinstance(t, TensorNetwork):
            self.tensors.extend(t.tensors)
        else:
            self.tensors.append(t)
        return self

    addtensor = append

    def ntensors(self):
        return len(self.tensors)

    def copy(self):
        return TensorNetwork(list(self.tensors))

    def size(self, ind) -> int:
        for t in self.tensors:
            if ind in t.inds:
                return t.size(ind)
        raise KeyError(ind)

    def sizes(self) -> Dict[Hashable, int]:
        s = {}
        for t in self.tensors:
            for i, d in zip(t.inds, t.shape):
                if s.setdefault(i, d) != d:
                    raise ValueError(f"inconsistent extents for index {i!r}: {s[i]} vs {d}")
        return s

    def inds(self, set: str = "all") -> Tuple[Hashable, ...]:
        """`inds(tn; set=:all | :open | :inner)` in first-appearance order."""
        count, order = {}, []
        for t in self.tensors:
            for i in t.inds:
                if i not in count:
                    order.append(i)
                count[i] = count.get(i, 0) + 1
        if set == "all":
            return tuple(order)
        if set == "open":
            return tuple(i for i in order if count[i] == 1)
        if set == "inner":
            return tuple(i for i in order if count[i] > 1)
        raise ValueError(set)

    def conj(self):
        return TensorNetwork([t.conj() for t in self.tensors])

    def replace(self, mapping: dict):
        return TensorNetwork([t.replace(mapping) for t in self.tensors])

    def view(self, *pairs):
        """`view(tn, ind => i, ...)`: fix indices in every tensor that carries them (network-level slicing)."""
        out = []
        for t in self.tensors:
            ps = [(i, s) for (i, s) in pairs if i in t.inds]
            out.append(t.view(*ps) if ps else t)
        return TensorNetwork(out)

    def __iter__(self):
        return iter(self.tensors)

    def __len__(self):
        return len(self.tensors)


def einexpr(tn: TensorNetwork, optimizer: str = "greedy", output: Optional[Sequence[Hashable]] = None,
            ntrials: int = 16, seed: int = 0, max_log2_size: Optional[float] = None,
            minimize: str = "flops", slice_trials: int = 0) -> ContractionPath:
    """`einexpr(tn; optimizer=Greedy())`: find a pairwise contraction path; if `max_log2_size` is given and the
    best path exceeds it, also pick indices to slice (EinExprs' slicing)."""
    if optimizer not in ("greedy", "Greedy"):
        raise ValueError(f"unknown optimizer {optimizer!r}")
    inputs = [t.inds for t in tn.tensors]
    sizes = tn.sizes()
    output = tuple(tn.inds("open") if output is None else output)
    p = optimize_path(inputs, sizes, output, ntrials=ntrials, seed=seed, minimize=minimize, max_log2_size=max_log2_size)
    if max_log2_size is not None and p.log2_max_size > max_log2_size:
        p = find_slices(inputs, sizes, output, p, max_log2_size, reoptimize_trials=slice_trials, seed=seed)
    return p


class ContractionPlan:
    """A planned, device-resident contraction (tnb_plan): build once, execute for any slice range."""

    def __init__(self, tn: TensorNetwork, path: ContractionPath, output: Optional[Sequence[Hashable]] = None,
                 ctx: Optional[Context] = None, dtype=None, dry: bool = False):
        self.ctx = None if dry else (ctx or default_context())
        self.lib = _lib.load_library()
        self.tn = tn
        self.path = path
        sliced = tuple(path.sliced)
        if output is None:
            output = path.output if path.output else tuple(i for i in tn.inds("open") if i not in sliced)
        self.output = tuple(output)
        dt = np.dtype(dtype) if dtype is not None else _promote_dtype(*[t.dtype for t in tn.tensors])
        self.dtype = dt
        code = _lib.DTYPE_CODE[dt]
        modes = {}
        for t in tn.tensors:
            for i in t.inds:
                modes.setdefault(i, len(modes))
        self.modes = modes
        for i in tuple(sliced) + self.output:
            if i not in modes:
                raise ValueError(f"index {i!r} is not carried by any tensor of the network")
        sizes = tn.sizes()
        n = len(tn.tensors)
        self._keep = []
        descs = (tnb_tensor * n)()
        for k, t in enumerate(tn.tensors):
            if dry:
                from .context import fortran_strides
                d, keep = make_desc(None, 0, code, t.shape, fortran_strides(t.shape), [modes[i] for i in t.inds], t._conj)
            else:
                arr = t.device(self.ctx, dt)
                d, keep = make_desc(arr.buffer.handle, arr.offset, code, arr.shape, arr.strides,
                                    [modes[i] for i in t.inds], t._conj)
                self._keep.append(arr)
            descs[k] = d
            self._keep.append(keep)
        oshape = [sizes[i] for i in self.output]
        if dry:
            from .context import fortran_strides
            self.out_array = None
            dout, keep = make_desc(None, 0, code, oshape, fortran_strides(oshape), [modes[i] for i in self.output])
        else:
            self.out_array = B200Array.zeros(oshape, dt, self.ctx)
            dout, keep = make_desc(self.out_array.buffer.handle, 0, code, oshape, self.out_array.strides,
                                   [modes[i] for i in self.output])
        self._keep.append(keep)
        steps = (C.c_int32 * max(2 * len(path.steps), 1))(*[x for s in path.steps for x in s])
        sm = (C.c_int32 * max(len(sliced), 1))(*[modes[i] for i in sliced])
        h = C.c_void_p()
        if dry:
            rc = self.lib.tnb_plan_create_dry(descs, n, steps, len(path.steps), sm, len(sliced), C.byref(dout), C.byref(h))
            if rc:
                raise _lib.TnbError(rc, (self.lib.tnb_last_error(None) or b"").decode())
        else:
            check(self.ctx.handle, self.lib.tnb_plan_create(self.ctx.handle, descs, n, steps, len(path.steps), sm,
                                                            len(sliced), C.byref(dout), C.byref(h)))
        self.handle = h
        self.nsteps = len(path.steps)
        info = tnb_plan_info()
        self.lib.tnb_plan_get_info(self.handle, C.byref(info))
        self.info = {f: getattr(info, f) for f, _ in tnb_plan_info._fields_}
        self.nslices = int(info.nslices)

    def step_info(self, s: int) -> dict:
        si = tnb_step_info()
        self.lib.tnb_plan_get_step(self.handle, s, C.byref(si))
        d = {f: getattr(si, f) for f, _ in tnb_step_info._fields_}
        d["kernel_name"] = _lib.KERNEL_NAMES.get(d["kernel"], "?")
        return d

    def execute(self, slice_begin: int = 0, slice_step: int = 1, slice_end: Optional[int] = None,
                accumulate: bool = False):
        """out (+)= sum of slices slice_begin, +slice_step, ... < slice_end.  Asynchronous (stream-ordered)."""
        end = self.nslices if slice_end is None else int(slice_end)
        check(self.ctx.handle, self.lib.tnb_plan_execute(self.ctx.handle, self.handle, int(slice_begin),
                                                         int(slice_step), end, 1 if accumulate else 0))

    def profile(self, enable: bool = True):
        """Bracket every step launch with CUDA events (one host sync per slice while enabled)."""
        check(self.ctx.handle, self.lib.tnb_plan_profile(self.ctx.handle, self.handle, 1 if enable else 0))

    def step_times(self):
        """[(ms_total, runs)] per step since profiling was (re-)enabled."""
        out = []
        for s in range(self.nsteps):
            ms, n = C.c_double(), C.c_int64()
            self.lib.tnb_plan_get_step_time(self.handle, s, C.byref(ms), C.byref(n))
            out.append((ms.value, n.value))
        return out

    def zero_output(self):
        a = self.out_array
        check(self.ctx.handle, self.lib.tnb_memset_zero(self.ctx.handle, a.buffer.handle, 0, a.size * a.dtype.itemsize))

    def result(self) -> Tensor:
        return Tensor(self.out_array, self.output)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.tnb_plan_destroy(self.ctx.handle if self.ctx else None, self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def contract(tn, path: Optional[ContractionPath] = None, optimizer: str = "greedy",
             output: Optional[Sequence[Hashable]] = None, ctx: Optional[Context] = None,
             slice_range: Optional[Tuple[int, int, int]] = None, **einexpr_kw) -> Tensor:
    """`contract(tn; path=einexpr(tn; optimizer))` on the B200: every inner index is summed along the path's
    binary tree, open indices survive; sliced indices of the path are summed slice by slice.

    The index order of the result is `output` if given, else the network's open indices in first-appearance
    order (consumers in the reference never rely on it: DMRG.jl:138, test/integration/itensormps.jl:46-48).
    """
    if not isinstance(tn, TensorNetwork):
        tn = TensorNetwork(getattr(tn, "tensors"))   # any front-end exposing .tensors (MPS, MPO, PEPS ...)
    if len(tn.tensors) == 0:
        raise ValueError("cannot contract an empty tensor network")
    if len(tn.tensors) == 1:
        t = tn.tensors[0]
        keep = tuple(t.inds if output is None else output)
        if set(keep) == set(t.inds):
            return t.permutedims(keep)
        raise NotImplementedError("contract of a single tensor with a trace/sum is not on the accelerated path")
    if path is None:
        path = einexpr(tn, optimizer=optimizer, output=output, **einexpr_kw)
    plan = ContractionPlan(tn, path, output=output, ctx=ctx)
    try:
        if slice_range is None:
            plan.execute(0, 1, plan.nslices, accumulate=True)
        else:
            b, s, e = slice_range
            plan.execute(b, s, e, accumulate=True)
        res = plan.result()
    finally:
        plan.close()
    return res


def multi_contract(tn, path: ContractionPath, ngpus: int, output: Optional[Sequence[Hashable]] = None,
                   ctx: Optional[Context] = None) -> Tensor:
    """`contract(tn; path)` over `ngpus` GPUs of this box from ONE process (tnb_multi_contract_path): the leaves live on
    `ctx`'s device, the library broadcasts them, deals slice s to GPU s mod ngpus (one host thread + one context per
    device), sums the per-GPU accumulators with a single NCCL all-reduce and returns the result on `ctx`'s device.
    This is what a single Julia task calling `contract` binds; the one-process-per-GPU form is
    `distributed.contract_distributed`."""
    if not isinstance(tn, TensorNetwork):
        tn = TensorNetwork(getattr(tn, "tensors"))
    ctx = ctx or default_context()
    lib = _lib.load_library()
    sliced = tuple(path.sliced)
    if output is None:
        output = path.output if path.output else tuple(i for i in tn.inds("open") if i not in sliced)
    output = tuple(output)
    dt = _promote_dtype(*[t.dtype for t in tn.tensors])
    code = _lib.DTYPE_CODE[dt]
    modes = {}
    for t in tn.tensors:
        for i in t.inds:
            modes.setdefault(i, len(modes))
    sizes = tn.sizes()
    n = len(tn.tensors)
    keep_alive = []
    descs = (tnb_tensor * n)()
    for k, t in enumerate(tn.tensors):
        arr = t.device(ctx, dt)
        d, keep = make_desc(arr.buffer.handle, arr.offset, code, arr.shape, arr.strides, [modes[i] for i in t.inds], t._conj)
        descs[k] = d
        keep_alive += [arr, keep]
    oshape = [sizes[i] for i in output]
    out_array = B200Array.zeros(oshape, dt, ctx)
    dout, keep = make_desc(out_array.buffer.handle, 0, code, oshape, out_array.strides, [modes[i] for i in output])
    keep_alive.append(keep)
    steps = (C.c_int32 * max(2 * len(path.steps), 1))(*[x for s in path.steps for x in s])
    sm = (C.c_int32 * max(len(sliced), 1))(*[modes[i] for i in sliced])
    check(ctx.handle, lib.tnb_multi_contract_path(ctx.handle, descs, n, steps, len(path.steps), sm, len(sliced),
                                                  C.byref(dout), int(ngpus)))
    return Tensor(out_array, output)
