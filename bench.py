#!/usr/bin/env python
"""bench.py — contraction throughput of the hot path on N B200s (one process per GPU).

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

A "step" contracts one batch of slices of the workload's network along its committed path
(bench_paths/<workload>.json): `slices_per_step` slices on every GPU, dealt round-robin (slice id mod N), summed
into the per-GPU accumulator, followed by the single all-reduce when N > 1.  Work per GPU is fixed as N grows
("weak") in the timed steps; after them the FULL amplitude (all 2^k slices over the N GPUs, hoisted steps on every
GPU, one all-reduce) is contracted once and reported as `full_amplitude_s` — the strong-scaling figure — together
with the amplitude and its distance from the committed single-GPU golden (tests/golden/sycamore53_m14_amplitude.json).

Printed JSON (one line, rank 0): the driver contract keys plus
  `roofline`      dominant kernel: algorithmic flops or bytes per launch / CUDA-event duration, measured in a separate
                  profiled region (per-step events cost one host sync per slice, so `value` is timed WITHOUT them);
  `cpu_baseline`  the numpy/OpenBLAS oracle timed on the host cores on a bounded sample (cores and BLAS threads stated);
  `e2e`           same metric through the public API, leaves from pinned host memory, result read back, every step;
  `configs`       (N = 1 only) one short measurement of each other BASELINE config with its own roofline / cpu_baseline.
"""
import os
import sys


def _early_reference_setup():
    """--impl reference: the BLAS thread pools are sized when numpy loads, and torch.distributed.run exports
    OMP_NUM_THREADS=1 — so fix the environment BEFORE numpy is imported; ranks other than 0 have nothing to do."""
    argv = sys.argv[1:]
    ref = any(a == "--impl=reference" for a in argv) or any(
        a == "--impl" and i + 1 < len(argv) and argv[i + 1] == "reference" for i, a in enumerate(argv))
    if not ref:
        return
    if int(os.environ.get("RANK", "0")) != 0:
        sys.exit(0)
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = str(n)


_early_reference_setup()

import argparse  # noqa: E402
import json  # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (description, dtype)
    "sycamore53_m14": "Sycamore-like 53-qubit depth-14 random-circuit amplitude, complex64, sliced (BASELINE configs[2])",
    "sycamore53_m14_v1": "same network, earlier committed path (flops objective: 4096 slices x 2^37.95 MACs, HBM-bound stem steps dominate)",
    "sycamore53_m14_greedy": "same network, round-1 first-light path (plain greedy, 2^63 MACs total): two 16384x8192x8192 GEMMs dominate",
    "sycamore53_m10": "Sycamore-like 53-qubit depth-10 random-circuit amplitude, complex64, sliced",
    "regular3_n60_d4": "random 3-regular network, 60 tensors, bond 4, complex64 (BASELINE configs[1] scaled to fit)",
    "regular3_n100_d4": "random 3-regular network, 100 tensors, bond 4, complex64, 64 slices (BASELINE configs[1]; 200 tensors needs 2^96 MACs)",
    "peps6x6_d4_boundary": "6x6 PEPS norm, D=4, complex128, row-by-row boundary path, unsliced (BASELINE configs[3])",
    "peps6x6_d4": "6x6 PEPS norm, D=4, complex128 (BASELINE configs[3])",
    "mps_norm": "MPS <psi|psi>, 32 sites chi=128, complex128, zipper path (BASELINE configs[0])",
    "mps_mpo": "MPS-MPO <psi|H|psi>, 100 sites chi=1024, complex128, env sweep (BASELINE configs[4])",
}
EXTRA_CONFIGS = ["mps_norm", "regular3_n100_d4", "peps6x6_d4_boundary", "mps_mpo"]
# measured on this pool's B200s with tools/yardstick.py (cuBLAS ZGEMM 8192^3, never linked into the product;
# profiles/r1_summary.md): the FP64 "complex tensor-core peak" SURVEY §8d asks the c128 fraction to be quoted against
ZGEMM_MEASURED_TFLOPS = 37.0
GOLDEN_AMPLITUDE = os.path.join(ROOT, "tests", "golden", "sycamore53_m14_amplitude.json")
C64_AMPLITUDE_BOUND = 5e-5        # the one stated c64 bound: relative to |amplitude|


def build_workload(tb, name, path_file=""):
    """-> (TensorNetwork, ContractionPath)"""
    from tools.make_paths import network  # noqa
    if name == "peps6x6_d4_boundary":
        tn, _ = tb.workloads.peps_norm_network(6, 6, D=4, p=2, dtype=np.complex128, seed=4)
        return tn, tb.workloads.peps_boundary_path(6, 6)
    if name in ("sycamore53_m14", "sycamore53_m14_v1", "sycamore53_m14_greedy", "sycamore53_m10", "regular3_n60_d4", "regular3_n100_d4", "peps6x6_d4"):
        tn = network(name)
        fn = path_file or os.path.join(ROOT, "bench_paths", name + ".json")
        if not os.path.exists(fn):
            raise SystemExit(f"{fn} missing: run `python tools/make_paths.py {name}`")
        path = tb.pathfinder.load_path(tn.inds("all"), fn)
        return tn, path
    if name == "mps_norm":
        tn, _ = tb.workloads.mps_norm_network(32, 128, np.complex128, seed=1)
        return tn, tb.workloads.zipper_path(32)
    if name == "mps_mpo":
        tn, _, _ = tb.workloads.mps_mpo_expectation_network(100, 1024, dtype=np.complex128, seed=5)
        return tn, tb.workloads.sweep_path(100)
    raise SystemExit(f"unknown workload {name}; choose from {sorted(WORKLOADS)}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(gpu_index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples in the upper half of the power range
        thr = (max(pw) + min(pw)) / 2
        load = [s for s, p in zip(sm, pw) if p >= thr] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def host_threads():
    """(cores this process may run on, BLAS threads numpy's BLAS will use)"""
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    blas = None
    try:
        from threadpoolctl import threadpool_info
        th = [p.get("num_threads") for p in threadpool_info() if p.get("user_api") == "blas"]
        blas = max(th) if th else None
    except Exception:
        pass
    return cores, blas


def subslice_path(tb, tn, path, budget_macs_log2=38.5):
    """The committed path with extra sliced indices, added until one sub-slice costs <= 2^budget MACs: the piece of
    the full-size workload a CPU can finish (also used by tests/test_gpu_zz_fullsize.py as the oracle-sized case)."""
    inputs = [t.inds for t in tn.tensors]
    sizes = tn.sizes()
    p = path
    target = p.log2_max_size
    while p.log2_macs > budget_macs_log2 and target > 8:
        target -= 1.0
        p = tb.find_slices(inputs, sizes, (), tb.ContractionPath(list(path.steps), tuple(p.sliced)), target)
    return p


def cpu_sample(tb, tn, path, budget_macs_log2=38.5, deadline_s=40.0):
    """A bounded CPU sample of the same workload.  Sliced paths: slice further until one sub-slice is ~2^budget
    MACs, then time the oracle on sub-slice 0.  Un-sliced paths that are too long for the budget: the oracle walks
    the same path step by step and stops at the first step boundary after `deadline_s` (rate = MACs done / time).
    Returns (flops, seconds, description)."""
    from oracle import einsum_oracle as orc
    inputs = [t.inds for t in tn.tensors]
    arrays = [t.parent for t in tn.tensors]
    cplx = np.iscomplexobj(arrays[0])
    unit = 8.0 if cplx else 2.0
    if len(path.sliced) == 0:
        # hand-written paths (zipper, env sweep, PEPS boundary) carry no cost estimate: count the MACs here
        macs, _, _ = orc.path_flops(inputs, tn.sizes(), path.steps)
        log2_macs = float(np.log2(max(float(macs), 1.0)))
        if log2_macs > budget_macs_log2 + 0.5:
            return cpu_sample_prefix(orc, arrays, inputs, path, unit, deadline_s, log2_macs)
        t0 = time.perf_counter()
        orc.contract_path(arrays, inputs, path.steps)
        dt = time.perf_counter() - t0
        return unit * float(macs), dt, f"the whole path (2^{log2_macs:.1f} MACs); numpy/OpenBLAS permute->reshape->gemm restatement"
    p = subslice_path(tb, tn, path, budget_macs_log2)
    sl = list(p.sliced)
    t0 = time.perf_counter()
    orc.contract_sliced(arrays, inputs, p.steps, sl, slice_ids=[0])
    dt = time.perf_counter() - t0
    flops = unit * 2.0 ** p.log2_macs
    extra = len(p.sliced) - len(path.sliced)
    what = (f"1 sub-slice (slice 0 with {extra} extra sliced indices, 2^{p.log2_macs:.1f} MACs, peak 2^{p.log2_max_size:.0f} "
            f"elements) of the same path") if extra or len(path.sliced) else f"the whole path (2^{p.log2_macs:.1f} MACs)"
    return flops, dt, what + "; numpy/OpenBLAS permute->reshape->gemm restatement"


def cpu_sample_prefix(orc, arrays, inds, path, unit, deadline_s, log2_macs_total):
    """oracle.contract_path step by step with a wall-clock deadline (un-sliced, long paths: configs[4])."""
    n = len(arrays)
    total = {}
    for t in inds:
        for i in t:
            total[i] = total.get(i, 0) + 1
    vals = [np.asarray(a) for a in arrays]
    vinds = [tuple(t) for t in inds]
    cnt = [{i: 1 for i in t} for t in inds]
    ext = {}
    for a, t in zip(arrays, inds):
        for ax, i in enumerate(t):
            ext[i] = a.shape[ax]
    macs, done = 0.0, 0
    t0 = time.perf_counter()
    for s, (i, j) in enumerate(path.steps):
        ci = dict(cnt[i])
        for k, v in cnt[j].items():
            ci[k] = ci.get(k, 0) + v
        dims = [k for k, v in ci.items() if v == total[k]]
        c, c_inds = orc.binary_einsum(vals[i], vinds[i], vals[j], vinds[j], dims=dims)
        macs += float(np.prod([float(ext[k]) for k in ci]))
        vals.append(c); vinds.append(c_inds)
        cnt.append({k: v for k, v in ci.items() if v < total[k]})
        vals[i] = vals[j] = None
        done = s + 1
        if time.perf_counter() - t0 > deadline_s:
            break
    dt = time.perf_counter() - t0
    return unit * macs, dt, (f"the first {done} of {len(path.steps)} pairwise steps of the same path (2^{np.log2(max(macs, 1)):.1f} of "
                             f"2^{log2_macs_total:.1f} MACs, stopped at the first step boundary after {deadline_s:.0f} s); "
                             f"numpy/OpenBLAS permute->reshape->gemm restatement")


def load_peaks():
    peaks = {}
    pk_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_file):
        peaks = json.load(open(pk_file))
    return peaks


def roofline_of(plan, step_times, dtype, workload):
    """roofline object of the kernel with the largest share of the profiled region + the per-step rows"""
    peaks = load_peaks()
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    bf16_peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    by_kernel, rows = {}, []
    for s in range(plan.nsteps):
        si = plan.step_info(s)
        t_ms, runs = step_times[s]
        if runs == 0:
            continue
        k = by_kernel.setdefault(si["kernel_name"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        k["ms"] += t_ms; k["flops"] += si["flops"] * runs; k["bytes"] += si["bytes"] * runs; k["launches"] += runs
        rows.append({"step": s, **{x: si[x] for x in ("M", "N", "K", "L", "kernel_name", "flops", "bytes")},
                     "ms_avg": t_ms / runs, "runs": runs})
    if not by_kernel:
        return None, rows
    name, k = max(by_kernel.items(), key=lambda kv: kv[1]["ms"])
    ai = k["flops"] / max(k["bytes"], 1.0)
    # c64: a complex MAC is 4 real MACs, each needing 3 TF32 products for fp32-grade accuracy -> the tensor peak in
    # "8 flops per complex MAC" units is tf32_dense / 3, tf32_dense = bf16_dense / 2.  c128: tcgen05 has no FP64 kind;
    # the FP64 peak is the cuBLAS ZGEMM measured on this pool ("% of measured cublasZgemm").
    if dtype == np.complex64 or dtype == np.float32:
        tpeak, basis = bf16_peak / 2.0 / 3.0, "bf16_tflops_sustained/2 (TF32) /3 (3xTF32 split); " + peak_src
    else:
        tpeak, basis = ZGEMM_MEASURED_TFLOPS, "cuBLAS ZGEMM 8192^3 measured on this pool (tools/yardstick.py, profiles/r1_summary.md)"
    ridge = tpeak * 1e12 / (hbm_peak * 1e9)
    if ai >= ridge:
        ach = k["flops"] / (k["ms"] * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak,
                    "traffic": None, "peak_basis": basis}
    else:
        ach = k["bytes"] / (k["ms"] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "traffic": None, "peak_basis": peak_src}
    # DRAM traffic of the dominant kernel: from the ncu --set full capture of THIS build (the file records the digest of
    # the kernel sources it was taken from; a capture of other sources is reported as stale, not silently reused)
    tr_file = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(tr_file):
        trj = json.load(open(tr_file))
        tr = trj.get("workloads", {}).get(workload, {}).get(name)
        if tr:
            cur = None
            try:
                import importlib.util
                spec = importlib.util.spec_from_file_location("_tnb_build", os.path.join(ROOT, "tenet.jl_b200", "build.py"))
                mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
                cur = mod.source_digest()
            except Exception:
                pass
            if cur is not None and cur == trj.get("source_digest"):
                roofline["traffic"] = tr["dram_bytes"]
                roofline["traffic_note"] = (f"ncu dram read+write of the largest launch of this kernel, {tr['launch']}: "
                                            f"{tr['ratio']:.3f} x its algorithmic bytes (profiles/r2_traffic.json, same sources)")
            else:
                roofline["traffic_note"] = "profiles/r2_traffic.json was captured from other kernel sources: stale, not reported"
    roofline.update({"kernel": name, "launches": k["launches"], "avg_launch_ms": k["ms"] / k["launches"],
                     "share_of_step": k["ms"] / max(sum(v["ms"] for v in by_kernel.values()), 1e-9),
                     "arithmetic_intensity": ai,
                     "kernels": {n: {"ms": v["ms"], "tflops": v["flops"] / max(v["ms"], 1e-9) / 1e9,
                                     "gbs": v["bytes"] / max(v["ms"], 1e-9) / 1e6, "launches": v["launches"]}
                                 for n, v in by_kernel.items()}})
    return roofline, rows


class Runner:
    """One workload on this rank's GPU: plan, pinned leaves, step function, timed regions."""

    def __init__(self, tb, torch, dist, ctx, stream, name, rank, world, path_file="", slices_per_step=0):
        self.tb, self.torch, self.dist, self.ctx, self.stream = tb, torch, dist, ctx, stream
        self.name, self.rank, self.world = name, rank, world
        self.tn, self.path = build_workload(tb, name, path_file)
        tn = self.tn
        self.dtype = tb.tensor._promote_dtype(*[t.dtype for t in tn.tensors])
        self.cplx = self.dtype.kind == "c"
        dtype, cplx = self.dtype, self.cplx
        # pinned host copies of the leaves (e2e: uploaded every step) + device residency (value: resident)
        self.pinned = []
        for t in tn.tensors:
            h = np.asfortranarray(t.parent.astype(dtype))
            flat = np.ascontiguousarray(h.reshape(-1, order="F"))
            real = flat.view(np.float32 if dtype.itemsize // (2 if cplx else 1) == 4 else np.float64)
            self.pinned.append(torch.from_numpy(real.copy()).pin_memory())
        self.plan = tb.ContractionPlan(tn, self.path, ctx=ctx)
        self.info = self.plan.info
        self.nslices = self.plan.nslices
        self.flops_slice = self.info["flops_per_slice"]
        S = slices_per_step
        if S <= 0:
            S = max(1, int(15e12 / max(self.flops_slice, 1.0)))      # first guess, refined by calibrate()
        self.S = max(1, min(S, max(1, self.nslices // max(world, 1))))
        self.auto_S = slices_per_step <= 0
        self.replicas = self.nslices == 1   # un-sliced network: nothing to shard -> N independent replicas, no collective

    def step_range(self, i):
        if self.replicas:
            return 0, 1, 1
        base = (i * self.world * self.S) % max(self.nslices - self.world * self.S + 1, 1)
        return base + self.rank, self.world, base + self.world * self.S

    def run_step(self, i, e2e):
        tb, ctx, plan = self.tb, self.ctx, self.plan
        if e2e:
            for t, buf in zip(self.tn.tensors, self.pinned):
                arr = t._dev
                tb._lib.check(ctx.handle, ctx.lib.tnb_upload(ctx.handle, arr.buffer.handle, 0, buf.data_ptr(),
                                                             arr.size * self.dtype.itemsize))
        plan.zero_output()
        b, s, e = self.step_range(i)
        plan.execute(b, s, e, accumulate=True)
        if self.world > 1 and not self.replicas:
            tb.distributed.allreduce_sum(ctx, plan.out_array, self.world)
        if e2e:
            return plan.out_array.to_numpy()
        return None

    def barrier(self):
        self.ctx.sync()
        if self.world > 1:
            self.dist.barrier()

    def calibrate(self):
        """the very first step pays one-time costs (function attributes, NCCL communicator set-up inside the first
        all-reduce): run it once untimed, then size S for ~1 s steps from a second one"""
        torch = self.torch
        self.run_step(0, False)
        self.barrier()
        t0 = time.perf_counter()
        self.run_step(0, False)
        self.ctx.sync()
        first = time.perf_counter() - t0
        if self.auto_S and self.nslices > 1:
            per_slice = first / self.S
            S2 = max(1, min(int(round(1.0 / max(per_slice, 1e-6))), 4096, max(1, self.nslices // max(self.world, 1))))
            if self.world > 1:
                t = torch.tensor([S2], device="cuda")
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
                S2 = int(t.item())
            self.S = S2
        return first

    def timed(self, first_step, nsteps, e2e):
        """K steps bracketed by barrier + sync on both sides, CUDA events on the context stream -> (ms, last result)"""
        torch = self.torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        with torch.cuda.stream(self.stream):
            ev0.record()
        res = None
        for i in range(nsteps):
            res = self.run_step(first_step + i, e2e)
        with torch.cuda.stream(self.stream):
            ev1.record()
        self.barrier()
        return ev0.elapsed_time(ev1), res

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return vals
        t = self.torch.tensor(list(vals), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return tuple(float(x) for x in t)

    def work(self, nsteps):
        total_slices = nsteps * self.world * self.S if not self.replicas else nsteps * self.world
        hoisted = self.info["flops_hoisted"] * nsteps * self.world
        return total_slices, total_slices * self.flops_slice + hoisted

    def full_amplitude(self):
        """ALL slices over the N GPUs (slice s on rank s mod N, hoisted steps on every rank, one all-reduce):
        the strong-scaling measurement.  -> (seconds, result array)"""
        torch, plan = self.torch, self.plan
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        with torch.cuda.stream(self.stream):
            ev0.record()
        plan.zero_output()
        plan.execute(self.rank, self.world, self.nslices, accumulate=True)
        if self.world > 1:
            self.tb.distributed.allreduce_sum(self.ctx, plan.out_array, self.world)
        with torch.cuda.stream(self.stream):
            ev1.record()
        self.barrier()
        (ms,) = self.max_over_ranks(ev0.elapsed_time(ev1))
        return ms * 1e-3, plan.out_array.to_numpy()

    def close(self):
        self.plan.close()
        self.plan = self.tn = None      # leaves are freed with their Tensor objects (stream-ordered allocator)
        self.pinned = []


def dtype_name(dtype):
    return {"complex64": "c64", "complex128": "c128"}.get(str(dtype), str(dtype))


def measure_extra(tb, torch, ctx, stream, name, no_cpu):
    """A short single-GPU measurement of another BASELINE config: 2 warm-up + 2 timed steps, a profiled step for the
    roofline, a bounded CPU sample."""
    t_start = time.perf_counter()
    r = Runner(tb, torch, None, ctx, stream, name, 0, 1)
    r.calibrate()
    r.run_step(1, False)
    ms, _ = r.timed(2, 2, False)
    ms_e2e, res = r.timed(2, 2, True)
    r.plan.profile(True)
    r.run_step(2, False)
    ctx.sync()
    st = r.plan.step_times()
    r.plan.profile(False)
    roofline, _ = roofline_of(r.plan, st, r.dtype, name)
    slices, flops = r.work(2)
    out = {"description": WORKLOADS[name], "dtype": dtype_name(r.dtype), "value": flops / (ms * 1e-3) / 1e12, "unit": "TFLOP/s",
           "ms_per_step": ms / 2, "e2e": flops / (ms_e2e * 1e-3) / 1e12, "slices_per_step": r.S if not r.replicas else None,
           "nslices_total": r.nslices, "flops_per_slice": r.flops_slice, "steps_per_contraction": r.info["nsteps_per_slice"] + r.info["nsteps_hoisted"],
           "result": [float(np.real(res).sum()), float(np.imag(res).sum())], "roofline": roofline}
    if not no_cpu:
        cores, blas = host_threads()
        fl, dt, desc = cpu_sample(tb, r.tn, r.path, budget_macs_log2=36.5, deadline_s=12.0)
        out["cpu_baseline"] = {"value": fl / dt / 1e12, "unit": "TFLOP/s", "cores": cores, "blas_threads": blas, "kind": "port",
                               "sample": desc, "seconds": dt}
    r.close()
    ctx.trim()
    out["wall_s"] = time.perf_counter() - t_start
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="tnb200", choices=["tnb200", "reference"])
    ap.add_argument("--workload", default="sycamore53_m14")
    ap.add_argument("--slices-per-step", type=int, default=0, help="per GPU; 0 = sized for ~1 s steps")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the short measurements of the other BASELINE configs (N = 1)")
    ap.add_argument("--no-full", action="store_true", help="skip the full-amplitude (strong scaling) run")
    ap.add_argument("--c64-mode", default="auto", choices=["auto", "simt", "tf32x3", "tf32x3_fast"])
    ap.add_argument("--dump-steps", default="")
    ap.add_argument("--path-file", default="", help="candidate path JSON to use instead of bench_paths/<workload>.json")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")

    # ------------------------------------------------------------------------------------------------
    if a.impl == "reference":
        # The reference's CPU algorithm (numpy/OpenBLAS restatement: the reference itself cannot run here, BASELINE.md
        # §2) on the host cores, rank 0 only (the other ranks left in _early_reference_setup).  libtnb200.so is never
        # loaded in this process: only the pure-Python workload / path modules of the package are imported.
        import tenet_jl_b200 as tb
        tn, path = build_workload(tb, a.workload, a.path_file)
        dtype = tb.tensor._promote_dtype(*[t.dtype for t in tn.tensors])
        cores, blas = host_threads()
        # size the per-step sample so that the whole --steps/--warmup run stays within ~3 minutes
        fl, dt, desc = cpu_sample(tb, tn, path, budget_macs_log2=36.0, deadline_s=6.0)
        rate = fl / dt
        per_step_s = max(4.0, min(20.0, 150.0 / max(a.steps + a.warmup, 1)))
        unit = 8.0 if dtype.kind == "c" else 2.0
        budget = float(np.log2(max(rate * per_step_s / unit, 2.0 ** 30)))
        vals = []
        for i in range(a.warmup + a.steps):
            fl, dt, desc = cpu_sample(tb, tn, path, budget_macs_log2=budget, deadline_s=per_step_s)
            if i >= a.warmup:
                vals.append((fl, dt))
        fl = sum(v[0] for v in vals)
        dt = sum(v[1] for v in vals)
        val = fl / dt / 1e12
        print(json.dumps({
            "impl": "reference", "metric": "contraction_tflops", "value": val, "unit": "TFLOP/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / max(a.steps, 1) * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype_name(dtype),
            "data": "synthetic", "config": {"workload": a.workload, "description": WORKLOADS[a.workload]},
            "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": cores, "blas_threads": blas, "kind": "port", "sample": desc},
            "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    # ------------------------------------------------------------------------------------------------
    import __graft_entry__ as g
    g.build()
    import tenet_jl_b200 as tb
    import torch
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = tb.default_context(local_rank)
    if a.c64_mode != "auto":
        ctx.set_option(tb._lib.TNB_OPT_C64_MODE, {"simt": 0, "tf32x3": 1, "tf32x3_fast": 2}[a.c64_mode])
    if world > 1:
        tb.distributed.init_comm(ctx, rank, world)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)

    r = Runner(tb, torch, dist, ctx, stream, a.workload, rank, world, a.path_file, a.slices_per_step)
    dtype = r.dtype
    r.calibrate()
    for i in range(1, a.warmup):
        r.run_step(i, False)
    r.barrier()

    # timed region 1: inputs resident in HBM, no per-step events
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = ctx.launch_count
    ms, _ = r.timed(a.warmup, a.steps, False)
    launches = ctx.launch_count - l0
    # timed region 2: end to end (pinned host leaves -> H2D every step, result D2H every step)
    ms_e2e, res = r.timed(a.warmup, a.steps, True)
    clocks = sampler.stop() if sampler else None
    ms, ms_e2e = r.max_over_ranks(ms, ms_e2e)
    # region 3 (not part of any reported throughput): one step with per-step CUDA events -> kernel shares / roofline
    r.plan.profile(True)
    r.run_step(a.warmup, False)
    ctx.sync()
    step_times = r.plan.step_times()
    r.plan.profile(False)
    r.barrier()

    # strong scaling: the whole amplitude over the N GPUs
    full = None
    if not a.no_full and not r.replicas:
        secs, amp = r.full_amplitude()
        z = complex(np.asarray(amp).reshape(-1)[0]) if amp.size == 1 else None
        full = {"seconds": secs, "nslices": r.nslices, "slices_per_gpu": -(-r.nslices // world),
                "tflops": (r.nslices * r.flops_slice + world * r.info["flops_hoisted"]) / secs / 1e12}
        if z is not None:
            full["amplitude"] = [z.real, z.imag]
            if a.workload == "sycamore53_m14" and not a.path_file and os.path.exists(GOLDEN_AMPLITUDE):
                gold = json.load(open(GOLDEN_AMPLITUDE))
                g0 = complex(*gold["amplitude_c64_n1"])
                rel = abs(z - g0) / abs(g0)
                full["rel_diff_vs_golden_n1"] = rel
                full["bound"] = C64_AMPLITUDE_BOUND
                full["within_bound"] = bool(rel <= C64_AMPLITUDE_BOUND)
                if "amplitude_c128" in gold:     # the complex128 truth (same engine, FP64 kernels, tools/full_amplitude.py --dtype c128)
                    t128 = complex(*gold["amplitude_c128"])
                    full["rel_err_vs_c128_truth"] = abs(z - t128) / abs(t128)
                    full["within_bound"] = bool(full["within_bound"] and full["rel_err_vs_c128_truth"] <= C64_AMPLITUDE_BOUND)

    total_slices, total_flops = r.work(a.steps)
    value = total_flops / (ms * 1e-3) / 1e12
    e2e_value = total_flops / (ms_e2e * 1e-3) / 1e12
    h2d = int(sum(t._dev.size for t in r.tn.tensors) * dtype.itemsize)
    d2h = int(r.plan.out_array.size * dtype.itemsize)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    roofline, rows = roofline_of(r.plan, step_times, dtype, a.workload)
    if a.dump_steps:
        os.makedirs(os.path.dirname(os.path.abspath(a.dump_steps)), exist_ok=True)
        json.dump(rows, open(a.dump_steps, "w"))

    cpu = None
    if world == 1 and not a.no_cpu:
        cores, blas = host_threads()
        fl, dt, desc = cpu_sample(tb, r.tn, r.path)
        cpu = {"value": fl / dt / 1e12, "unit": "TFLOP/s", "cores": cores, "blas_threads": blas, "kind": "port", "sample": desc,
               "seconds": dt}

    info = r.info
    out = {
        "metric": "contraction_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": dtype_name(dtype),
        "data": "synthetic",
        "config": {"workload": a.workload, "description": WORKLOADS[a.workload], "slices_per_step_per_gpu": r.S,
                   "nslices_total_log2": float(np.log2(max(r.nslices, 1))), "flops_per_slice": r.flops_slice,
                   "steps_per_slice": info["nsteps_per_slice"], "peak_intermediate_bytes": info["max_intermediate_elems"] * dtype.itemsize,
                   "workspace_bytes": info["workspace_bytes"],
                   "l2": "intermediates (>= hundreds of MB per slice) exceed the 126 MB L2; no explicit flush",
                   "parallelism": ("single GPU" if world == 1 else f"{world} independent replicas (un-sliced network, no collective)"
                                   if r.replicas else f"slices round-robin over {world} GPU(s), one all-reduce per step")},
        "slices_per_s": total_slices / (ms * 1e-3), "slices_per_s_per_gpu": total_slices / (ms * 1e-3) / world,
        "e2e": {"value": e2e_value, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / a.steps, "result_checksum": [float(np.real(res).sum()), float(np.imag(res).sum())] if res is not None else None},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "full_amplitude": full, "full_amplitude_s": full["seconds"] if full else None,
    }
    if world == 1 and not a.no_extras and a.workload == "sycamore53_m14":
        r.close()
        ctx.trim()
        cfgs = {}
        for name in EXTRA_CONFIGS:
            try:
                cfgs[name] = measure_extra(tb, torch, ctx, stream, name, a.no_cpu)
            except Exception as e:  # noqa: BLE001 — an extra must never cost the headline line
                cfgs[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        out["configs"] = cfgs
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
