"""Worker of tests/test_gpu_multi.py (launched under torch.distributed.run, one process per GPU): every rank contracts its
round-robin share of a sliced network through the C ABI, the accumulators are summed with tnb_comm_allreduce_sum (NCCL
inside libtnb200), and every rank compares the all-reduced result with the oracle and with a single-GPU run of all slices."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    import tenet_jl_b200 as tb
    from oracle import einsum_oracle as orc
    from tolerances import C64_PATH_BOUND
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = tb.default_context(local)
    tb.distributed.init_comm(ctx, rank, world)
    assert ctx.lib.tnb_comm_size(ctx.handle) == world
    tn, _ = tb.workloads.sycamore_amplitude_network(rows=4, cols=3, cycles=8, seed=7, removed=(), dtype=np.complex64)
    path = tb.einexpr(tn, ntrials=8, seed=0, max_log2_size=4)
    assert path.nslices >= 2 * world
    plan = tb.ContractionPlan(tn, path, ctx=ctx)
    got = complex(tb.distributed.contract_distributed(plan, ctx).item())          # my slices + all-reduce
    plan.zero_output()
    plan.execute(0, 1, plan.nslices, accumulate=True)                               # all slices on this GPU alone
    single = complex(plan.result().item())
    arrays = [t.parent.astype(np.complex128) for t in tn.tensors]
    ref, _ = orc.contract_sliced(arrays, [t.inds for t in tn.tensors], path.steps, path.sliced)
    ref = complex(ref)
    ok = abs(got - ref) <= C64_PATH_BOUND * abs(ref) and abs(got - single) <= C64_PATH_BOUND * abs(ref)
    # a rank whose share is empty (more ranks than slices) must contribute zeros: shrink the plan to `world - 1` slices' worth
    plan.zero_output()
    plan.execute(rank, world, min(plan.nslices, world - 1), accumulate=True)
    tb.distributed.allreduce_sum(ctx, plan.out_array, world)
    part = complex(plan.result().item())
    pref, _ = orc.contract_sliced(arrays, [t.inds for t in tn.tensors], path.steps, path.sliced, slice_ids=range(min(plan.nslices, world - 1)))
    ok = ok and abs(part - complex(pref)) <= C64_PATH_BOUND * max(abs(complex(pref)), abs(ref))
    print(json.dumps({"rank": rank, "world": world, "ok": bool(ok), "got": [got.real, got.imag], "single": [single.real, single.imag],
                      "ref": [ref.real, ref.imag]}), flush=True)
    dist.barrier()
    plan.close()
    ctx.comm_destroy()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
