#!/usr/bin/env python
"""bench.py — contraction throughput of the hot path on N B200s (one process per GPU).

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

A "step" contracts one batch of slices of the workload's network along its committed path
(bench_paths/<workload>.json): `slices_per_step` slices on every GPU, dealt round-robin (slice id mod N), summed
into the per-GPU accumulator, followed by the single all-reduce when N > 1.  Work per GPU is fixed as N grows
("weak"): the full 2^k-slice amplitude is far beyond a benchmark's time budget, so each step samples a different
window of the slice space.

Printed JSON (one line, rank 0): the driver contract keys plus `roofline` (dominant kernel, algorithmic flops or
bytes per launch / CUDA-event duration measured on the context stream inside the timed region), `cpu_baseline`
(the numpy/OpenBLAS oracle timed on the host cores on a bounded sub-slice) and `e2e` (same metric through the
public API with the leaves coming from pinned host memory and the result read back, every step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

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


def build_workload(tb, name, path_file=""):
    """-> (TensorNetwork, ContractionPath, dtype)"""
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


def cpu_sample(tb, tn, path, budget_macs_log2=38.5):
    """A bounded CPU sample of the same workload: slice further until one sub-slice is ~2^budget MACs, then time
    the oracle on sub-slice 0.  Returns (flops, seconds, description)."""
    from oracle import einsum_oracle as orc
    inputs = [t.inds for t in tn.tensors]
    p = subslice_path(tb, tn, path, budget_macs_log2)
    arrays = [t.parent for t in tn.tensors]
    sl = list(p.sliced)
    t0 = time.perf_counter()
    orc.contract_sliced(arrays, inputs, p.steps, sl, slice_ids=[0])
    dt = time.perf_counter() - t0
    cplx = np.iscomplexobj(arrays[0])
    flops = (8.0 if cplx else 2.0) * 2.0 ** p.log2_macs
    extra = len(p.sliced) - len(path.sliced)
    desc = (f"1 sub-slice (slice 0 with {extra} extra sliced indices, 2^{p.log2_macs:.1f} MACs, peak 2^{p.log2_max_size:.0f} "
            f"elements) of the same path; numpy/OpenBLAS permute->reshape->gemm restatement")
    return flops, dt, desc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="tnb200", choices=["tnb200", "reference"])
    ap.add_argument("--workload", default="sycamore53_m14")
    ap.add_argument("--slices-per-step", type=int, default=0, help="per GPU; 0 = sized for ~1 s steps")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--c64-mode", default="auto", choices=["auto", "simt", "tf32x3", "tf32x3_fast"])
    ap.add_argument("--dump-steps", default="")
    ap.add_argument("--path-file", default="", help="candidate path JSON to use instead of bench_paths/<workload>.json")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")

    import __graft_entry__ as g
    g.build()
    import tenet_jl_b200 as tb

    tn, path = build_workload(tb, a.workload, a.path_file)
    dtype = tb.tensor._promote_dtype(*[t.dtype for t in tn.tensors])
    cplx = dtype.kind == "c"
    flops_unit = 8.0 if cplx else 2.0
    ncores = os.cpu_count() or 1

    # ------------------------------------------------------------------------------------------------
    if a.impl == "reference":
        # the reference's CPU algorithm (numpy/OpenBLAS restatement: the reference itself cannot run here, see
        # BASELINE.md §2) on the host cores, rank 0 only.
        if rank != 0:
            return
        vals = []
        for i in range(a.warmup + a.steps):
            fl, dt, desc = cpu_sample(tb, tn, path)
            if i >= a.warmup:
                vals.append((fl, dt))
        fl = sum(v[0] for v in vals)
        dt = sum(v[1] for v in vals)
        val = fl / dt / 1e12
        print(json.dumps({
            "impl": "reference", "metric": "contraction_tflops", "value": val, "unit": "TFLOP/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / max(a.steps, 1) * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "c64" if dtype == np.complex64 else str(dtype),
            "data": "synthetic", "config": {"workload": a.workload, "description": WORKLOADS[a.workload]},
            "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": ncores, "kind": "port", "sample": desc},
            "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    # ------------------------------------------------------------------------------------------------
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

    # pinned host copies of the leaves (e2e: uploaded every step) + device residency (value: resident)
    pinned = []
    for t in tn.tensors:
        h = np.asfortranarray(t.parent.astype(dtype))
        flat = np.ascontiguousarray(h.reshape(-1, order="F"))
        real = flat.view(np.float32 if dtype.itemsize // (2 if cplx else 1) == 4 else np.float64)
        pinned.append(torch.from_numpy(real.copy()).pin_memory())
    plan = tb.ContractionPlan(tn, path, ctx=ctx)
    info = plan.info
    nslices = plan.nslices
    flops_slice = info["flops_per_slice"]
    S = a.slices_per_step
    if S <= 0:
        # ~1 s per step assuming ~15 TFLOP/s for a first guess; refined from the warm-up below
        S = max(1, int(15e12 / max(flops_slice, 1.0)))
    S = max(1, min(S, max(1, nslices // max(world, 1))))

    replicas = nslices == 1       # un-sliced network: nothing to shard -> N independent replicas, no collective

    def step_range(i):
        if replicas:
            return 0, 1, 1
        base = (i * world * S) % max(nslices - world * S + 1, 1)
        return base + rank, world, base + world * S

    def run_step(i, e2e):
        if e2e:
            for t, buf in zip(tn.tensors, pinned):
                arr = t._dev
                tb._lib.check(ctx.handle, ctx.lib.tnb_upload(ctx.handle, arr.buffer.handle, 0, buf.data_ptr(),
                                                             arr.size * dtype.itemsize))
        plan.zero_output()
        b, s, e = step_range(i)
        plan.execute(b, s, e, accumulate=True)
        if world > 1 and not replicas:
            tb.distributed.allreduce_sum(ctx, plan.out_array)
        if e2e:
            return plan.out_array.to_numpy()
        return None

    def barrier():
        ctx.sync()
        if world > 1:
            dist.barrier()

    # warm-up (also calibrates S when auto): the very first step pays one-time costs (function attributes, NCCL
    # communicator set-up inside the first all-reduce), so it is run once untimed before the calibration step
    run_step(0, False)
    barrier()
    t0 = time.perf_counter()
    run_step(0, False)
    ctx.sync()
    first = time.perf_counter() - t0
    if a.slices_per_step <= 0 and nslices > 1:
        per_slice = first / S
        S2 = max(1, min(int(round(1.0 / max(per_slice, 1e-6))), 4096, max(1, nslices // max(world, 1))))
        if world > 1:
            t = torch.tensor([S2], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            S2 = int(t.item())
        S = S2
    for i in range(1, a.warmup):
        run_step(i, False)
    barrier()

    # timed region 1: inputs resident in HBM
    sampler = ClockSampler(local_rank) if rank == 0 else None
    plan.profile(True)
    l0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        ev0.record()
    for i in range(a.steps):
        run_step(a.warmup + i, False)
    with torch.cuda.stream(stream):
        ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - l0
    step_times = plan.step_times()
    plan.profile(False)

    # timed region 2: end to end (pinned host leaves -> H2D every step, result D2H every step)
    barrier()
    with torch.cuda.stream(stream):
        ev0.record()
    res = None
    for i in range(a.steps):
        res = run_step(a.warmup + i, True)
    with torch.cuda.stream(stream):
        ev1.record()
    barrier()
    ms_e2e = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    total_slices = a.steps * world * S
    hoisted = info["flops_hoisted"] * a.steps * world
    total_flops = total_slices * flops_slice + hoisted
    value = total_flops / (ms * 1e-3) / 1e12
    e2e_value = total_flops / (ms_e2e * 1e-3) / 1e12
    h2d = int(sum(t._dev.size for t in tn.tensors) * dtype.itemsize)
    d2h = int(plan.out_array.size * dtype.itemsize)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel (by measured time) ------------------------------------------------
    peaks = {}
    pk_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_file):
        peaks = json.load(open(pk_file))
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    bf16_peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    by_kernel = {}
    rows = []
    for s in range(plan.nsteps):
        si = plan.step_info(s)
        t_ms, runs = step_times[s]
        if runs == 0:
            continue
        k = by_kernel.setdefault(si["kernel_name"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        k["ms"] += t_ms; k["flops"] += si["flops"] * runs; k["bytes"] += si["bytes"] * runs; k["launches"] += runs
        rows.append({"step": s, **{x: si[x] for x in ("M", "N", "K", "L", "kernel_name", "flops", "bytes")},
                     "ms_avg": t_ms / runs, "runs": runs})
    roofline = None
    if by_kernel:
        name, k = max(by_kernel.items(), key=lambda kv: kv[1]["ms"])
        ai = k["flops"] / max(k["bytes"], 1.0)
        # c64: a complex MAC is 4 real MACs, each needing 3 TF32 products for fp32-grade accuracy -> the tensor
        # peak in "8 flops per complex MAC" units is tf32_dense / 3, tf32_dense = bf16_dense / 2.  c128: no FP64
        # tcgen05 kind exists; DMMA/DFMA nominal 37 TFLOP/s (not in MEASURED_PEAKS).
        if dtype == np.complex64 or dtype == np.float32:
            tpeak, basis = bf16_peak / 2.0 / 3.0, "bf16_tflops_sustained/2 (TF32) /3 (3xTF32 split)"
        else:
            tpeak, basis = 37.0, "nominal FP64 (148 SM x 64 FMA/clk x 1.965 GHz); no measured FP64 peak available"
        ridge = tpeak * 1e12 / (hbm_peak * 1e9)
        if ai >= ridge:
            ach = k["flops"] / (k["ms"] * 1e-3) / 1e12
            roofline = {"bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak,
                        "traffic": None, "peak_basis": basis + "; " + peak_src}
        else:
            ach = k["bytes"] / (k["ms"] * 1e-3) / 1e9
            roofline = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                        "traffic": None, "peak_basis": peak_src}
        tr_file = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if a.workload == "sycamore53_m14" and os.path.exists(tr_file):
            tr = json.load(open(tr_file)).get(name)
            if tr:
                roofline["traffic"] = tr["dram_bytes"]
                roofline["traffic_note"] = (f"ncu dram read+write of the largest launch of this kernel, {tr['launch']}: "
                                            f"{tr['ratio']:.3f} x its algorithmic bytes (profiles/r1_traffic.json)")
        roofline.update({"kernel": name, "launches": k["launches"], "avg_launch_ms": k["ms"] / k["launches"],
                         "share_of_step": k["ms"] / max(sum(v["ms"] for v in by_kernel.values()), 1e-9),
                         "arithmetic_intensity": ai,
                         "kernels": {n: {"ms": v["ms"], "tflops": v["flops"] / max(v["ms"], 1e-9) / 1e9,
                                         "gbs": v["bytes"] / max(v["ms"], 1e-9) / 1e6, "launches": v["launches"]}
                                     for n, v in by_kernel.items()}})
    if a.dump_steps:
        os.makedirs(os.path.dirname(os.path.abspath(a.dump_steps)), exist_ok=True)
        json.dump(rows, open(a.dump_steps, "w"))

    cpu = None
    if world == 1 and not a.no_cpu:
        fl, dt, desc = cpu_sample(tb, tn, path)
        cpu = {"value": fl / dt / 1e12, "unit": "TFLOP/s", "cores": ncores, "kind": "port", "sample": desc,
               "seconds": dt}

    out = {
        "metric": "contraction_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"complex64": "c64", "complex128": "c128"}.get(str(dtype), str(dtype)),
        "data": "synthetic",
        "config": {"workload": a.workload, "description": WORKLOADS[a.workload], "slices_per_step_per_gpu": S,
                   "nslices_total_log2": float(np.log2(max(nslices, 1))), "flops_per_slice": flops_slice,
                   "steps_per_slice": info["nsteps_per_slice"], "peak_intermediate_bytes": info["max_intermediate_elems"] * dtype.itemsize,
                   "workspace_bytes": info["workspace_bytes"],
                   "l2": "intermediates (>= hundreds of MB per slice) exceed the 126 MB L2; no explicit flush",
                   "parallelism": ("single GPU" if world == 1 else f"{world} independent replicas (un-sliced network, no collective)"
                                   if replicas else f"slices round-robin over {world} GPU(s), one all-reduce per step")},
        "slices_per_s": total_slices / (ms * 1e-3), "slices_per_s_per_gpu": total_slices / (ms * 1e-3) / world,
        "e2e": {"value": e2e_value, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / a.steps, "result_checksum": [float(np.real(res).sum()), float(np.imag(res).sum())] if res is not None else None},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
