"""Test-only: replay a DRY-RUN plan of libtnb200 with numpy gathers, so that the planner's decisions (GEMM views,
offset tables, intermediate layouts, arena placement, slice offsets) are checked on CPU against the oracle without
a GPU.  This never runs in the product path — it reads the planner's tables through the introspection entry points
(tnb_plan_dump_table / tnb_plan_dump_step) and does the arithmetic in numpy."""
import ctypes as C

import numpy as np


def _table(lib, plan, step, which):
    n = lib.tnb_plan_dump_table(plan.handle, step, which, None, 0)
    buf = (C.c_int64 * max(n, 1))()
    lib.tnb_plan_dump_table(plan.handle, step, which, buf, n)
    return np.frombuffer(buf, dtype=np.int64, count=n).copy()


def replay(plan, host_arrays, slice_ids=None):
    """host_arrays: numpy arrays of the leaves (any layout; flattened in Fortran order like the upload does)."""
    lib = plan.lib
    dt = plan.dtype
    leaves = [np.asarray(a).astype(dt).ravel(order="F") for a in host_arrays]
    conj_leaf = [t._conj for t in plan.tn.tensors]
    esz = dt.itemsize
    arena = np.zeros(max(1, plan.info["workspace_bytes"] // esz), dtype=dt)
    sizes = plan.tn.sizes()
    oshape = [sizes[i] for i in plan.output]
    out = np.zeros(int(np.prod(oshape, dtype=np.int64)) if oshape else 1, dtype=dt)
    ns = len(plan.path.sliced)
    sl_ext = [sizes[i] for i in plan.path.sliced]
    steps = []
    for s in range(plan.nsteps):
        ids = (C.c_int32 * 3)()
        base = (C.c_int64 * 3)()
        kind = (C.c_int32 * 3)()
        sst = (C.c_int64 * max(3 * ns, 1))()
        cj = (C.c_int32 * 2)()
        assert lib.tnb_plan_dump_step(plan.handle, s, ids, base, kind, sst, cj) == 0
        tabs = [_table(lib, plan, s, w) for w in range(9)]
        info = plan.step_info(s)
        steps.append(dict(ids=list(ids), base=list(base), kind=list(kind),
                          sst=np.array(list(sst)[:3 * ns], dtype=np.int64).reshape(3, ns) if ns else np.zeros((3, 0), np.int64),
                          conj=list(cj), tabs=tabs, info=info))

    def storage(kind, idx):
        return leaves[idx] if kind == 0 else arena if kind == 1 else out

    def run(st, digits, acc):
        am, ak, al, bn, bk, bl, cm, cn, cl = st["tabs"]
        offs = [st["base"][i] + int(np.dot(st["sst"][i], digits)) if ns else st["base"][i] for i in range(3)]
        A = storage(st["kind"][0], st["ids"][0])
        B = storage(st["kind"][1], st["ids"][1])
        Cst = storage(st["kind"][2], st["ids"][2])
        ia = offs[0] + al[:, None, None] + am[None, :, None] + ak[None, None, :]
        ib = offs[1] + bl[:, None, None] + bk[None, :, None] + bn[None, None, :]
        a = A[ia]
        b = B[ib]
        if st["conj"][0]:
            a = np.conj(a)
        if st["conj"][1]:
            b = np.conj(b)
        c = np.matmul(a, b)
        ic = offs[2] + cl[:, None, None] + cm[None, :, None] + cn[None, None, :]
        assert len(np.unique(ic)) == ic.size, "output offsets collide"
        if acc and st["kind"][2] == 2:
            Cst[ic] += c
        else:
            Cst[ic] = c

    hoisted = [st for st in steps if st["info"]["hoisted"]]
    dep = [st for st in steps if not st["info"]["hoisted"]]
    zero = np.zeros(ns, dtype=np.int64)
    if not dep:
        for st in hoisted:
            run(st, zero, False)
    else:
        for st in hoisted:
            run(st, zero, False)
        ids = range(plan.nslices) if slice_ids is None else slice_ids
        first = True
        for sid in ids:
            t, dig = sid, []
            for d in sl_ext:
                dig.append(t % d)
                t //= d
            for st in dep:
                run(st, np.array(dig, dtype=np.int64), not first)
            first = False
    return out.reshape(oshape, order="F")
