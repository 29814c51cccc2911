"""CPU, world_size 2, gloo: the N>1 host logic — round-robin slice partition, every slice exactly once, partial sums
all-reduced to the full result.  The arithmetic on each rank is the numpy oracle (there is no GPU here); what is under
test is tenet.jl_b200.distributed's partition + reduction plumbing, the same code bench.py uses on NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import tenet_jl_b200 as tb
    from oracle import einsum_oracle as orc
    tn = tb.workloads.random_regular_network(n=14, bond=3, dtype=np.complex128, seed=2)
    p = tb.einexpr(tn, ntrials=3, seed=0, max_log2_size=3.2)
    arrays = [t.parent for t in tn.tensors]
    inds = [t.inds for t in tn.tensors]
    assert tb.distributed.rank_world() == (rank, world)
    mine = tb.distributed.slices_of_rank(p.nslices, rank, world)
    part, _ = orc.contract_sliced(arrays, inds, p.steps, p.sliced, slice_ids=mine)
    total = tb.distributed.host_allreduce_sum(np.asarray(part))
    full, _ = orc.contract_path(arrays, inds, p.steps)
    counts = np.zeros(p.nslices)
    counts[mine] = 1
    counts = tb.distributed.host_allreduce_sum(counts)
    q.put((rank, complex(total), complex(full), bool(np.all(counts == 1)), len(mine), p.nslices))
    dist.destroy_process_group()


def test_round_robin_partition_and_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, total, full, once, nmine, ns in res:
        assert once and ns > 1
        assert abs(total - full) <= 1e-11 * abs(full)
    assert sum(r[4] for r in res) == res[0][5]


def test_slice_range_helpers():
    import tenet_jl_b200 as tb
    d = tb.distributed
    assert d.slice_range_for_rank(16, 1, 4) == (1, 4, 16)
    assert d.slices_of_rank(10, 3, 4) == [3, 7]
    assert d.slices_of_rank(100, 2, 8, limit=3) == [2, 10, 18]
    allids = sorted(sum((d.slices_of_rank(37, r, 8) for r in range(8)), []))
    assert allids == list(range(37))
