"""Multi-GPU parity ON HARDWARE (skipped on boxes with one GPU; run with `gpurun --gpus 2`):
  * tnb_multi_contract_path (one process, one host thread + context per device, NCCL inside the library) against the
    single-GPU result and the oracle;
  * the one-process-per-GPU path: tnb_comm_* driven through the C ABI by two processes (torch.distributed.run), the
    all-reduced amplitude against the single-GPU amplitude and the oracle."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from tolerances import C128_BOUND, C64_PATH_BOUND

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("dt,tol", [(np.complex64, C64_PATH_BOUND), (np.complex128, 100 * C128_BOUND)])
def test_multi_contract_path_matches_single_gpu_and_oracle(ctx, dt, tol):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    import tenet_jl_b200 as tb
    from oracle import einsum_oracle as orc
    tn, _ = tb.workloads.sycamore_amplitude_network(rows=4, cols=3, cycles=8, seed=7, removed=(), dtype=dt)
    path = tb.einexpr(tn, ntrials=8, seed=0, max_log2_size=4)
    assert path.nslices >= 4
    single = complex(tb.contract(tn, path=path, ctx=ctx).item())
    multi = complex(tb.multi_contract(tn, path, 2, ctx=ctx).item())
    again = complex(tb.multi_contract(tn, path, 2, ctx=ctx).item())       # cached worker contexts / communicator
    arrays = [t.parent.astype(np.complex128) for t in tn.tensors]
    ref, _ = orc.contract_sliced(arrays, [t.inds for t in tn.tensors], path.steps, path.sliced)
    ref = complex(ref)
    assert abs(multi - ref) <= tol * abs(ref), (multi, ref)
    assert abs(multi - single) <= tol * abs(ref), (multi, single)
    assert multi == again
    # open indices: a 2^3-element result, and an un-sliced path (rank 0 contracts, the others contribute zeros)
    psi = tb.MPS.rand(6, maxdim=8, eltype=np.complex128, rng=2)
    tn2 = tb.TensorNetwork(psi.tensors)
    p2 = tb.einexpr(tn2, ntrials=2, seed=0)
    a = tb.contract(tn2, path=p2, ctx=ctx)
    b = tb.multi_contract(tn2, p2, 2, ctx=ctx)
    bb = np.transpose(b.parent, [b.inds.index(i) for i in a.inds])
    assert np.abs(bb - a.parent).max() <= 100 * C128_BOUND * np.abs(a.parent).max()


def test_two_processes_allreduce_through_the_c_abi():
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ)
    env.pop("RANK", None); env.pop("WORLD_SIZE", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tests", "_mp_worker.py")],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert len(lines) == 2 and all(l["ok"] for l in lines), lines
    assert lines[0]["got"] == lines[1]["got"]
