"""`Tensor`, `Index`, `binary_einsum` — host-side mirror of the Muscle.jl interface the reference re-exports
(/root/reference/src/Tenet.jl:9-12).  Same names and argument meaning; the arithmetic runs in libtnb200.so.

Semantics fixed by the reference's call sites:
  (i)   default: contract every shared index          overlap.jl:42,46; DMRG.jl:17
  (ii)  dims=[] keeps shared indices as batch indices   canonize.jl:44; absorb.jl:31; evolve.jl:76,108
  (iii) no shared index: outer product, rank-0 operands DMRG.jl:60-61
  (iv)  full contraction: rank-0 result                 overlap.jl:49
  (v)   mixed element types are promoted                sample.jl:32-36; DMRG.jl:60 x Ising.jl:17
  (vi)  conj(t) is a flag, not a copy                   overlap.jl:7,39 (the reference materialises it)
  (vii) extent-1 indices                                MPS.jl:173-177
"""
from __future__ import annotations

import ctypes as C
from typing import Hashable, Iterable, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import check, make_desc
from .context import B200Array, Context, default_context


class Index:
    """Muscle `Index(tag)`: a label.  Any hashable works as an index; this wrapper exists for API fidelity."""

    __slots__ = ("tag",)

    def __init__(self, tag: Hashable):
        self.tag = tag.tag if isinstance(tag, Index) else tag

    def __hash__(self):
        return hash(("Index", self.tag))

    def __eq__(self, other):
        return isinstance(other, Index) and self.tag == other.tag

    def __repr__(self):
        return f"Index({self.tag!r})"


def _promote_dtype(*dts):
    dt = np.result_type(*dts)
    if dt.kind in "biu":
        dt = np.dtype(np.float64)
    if dt == np.float16:
        dt = np.dtype(np.float32)
    if dt not in _lib.DTYPE_CODE:
        raise TypeError(f"unsupported element type {dt}")
    return dt


class _Storage:
    """Shared by a tensor and its conj()/replace() aliases, so one upload serves e.g. ket and bra."""
    __slots__ = ("host", "dev")

    def __init__(self, host=None, dev=None):
        self.host, self.dev = host, dev


class Tensor:
    """`Tensor(array, inds)`: an N-d array plus one unique label per dimension.

    `data` is a numpy array (uploaded on first use) or a `B200Array`.  `conj()` flips a flag that the kernels
    honour at load time; `permutedims`, `replace` and `view` are metadata-only.
    """

    def __init__(self, data, inds: Sequence[Hashable] = (), *, _conj: bool = False):
        inds = tuple(inds)
        if len(set(inds)) != len(inds):
            raise ValueError(f"repeated index in {inds!r} (hyperindices within one tensor are not supported)")
        if isinstance(data, B200Array):
            self._st = _Storage(None, data)
            shape = data.shape
        else:
            arr = np.asarray(data)
            self._st = _Storage(arr, None)
            shape = arr.shape
        if len(shape) != len(inds):
            raise ValueError(f"ndims(array)={len(shape)} but {len(inds)} indices were given")
        self.inds = inds
        self._conj = bool(_conj)

    @property
    def _dev(self):
        return self._st.dev

    @property
    def _host(self):
        return self._st.host

    # -- Muscle accessors ---------------------------------------------------------------------------
    @property
    def shape(self):
        return self._dev.shape if self._dev is not None else self._host.shape

    @property
    def ndim(self):
        return len(self.inds)

    @property
    def dtype(self):
        return self._dev.dtype if self._dev is not None else self._host.dtype

    def size(self, ind=None):
        if ind is None:
            return self.shape
        return self.shape[self.inds.index(ind)]

    @property
    def parent(self) -> np.ndarray:
        """`parent(tensor)`: the array, on the host (downloads if the tensor lives on the device)."""
        a = self._dev.to_numpy() if self._dev is not None else np.asarray(self._host)
        return np.conj(a) if self._conj else a

    def __array__(self, dtype=None, copy=None):
        a = self.parent
        return a.astype(dtype) if dtype is not None else a

    def item(self):
        a = self.parent
        if a.size != 1:
            raise ValueError("item() needs a single-element tensor")
        return a.reshape(-1)[0]

    def device(self, ctx: Optional[Context] = None, dtype=None) -> B200Array:
        """`adapt(B200Array, tensor)`: make sure the data is device-resident (with promotion if asked)."""
        want = np.dtype(dtype) if dtype is not None else _promote_dtype(self.dtype)
        if self._dev is not None and self._dev.dtype == want:
            return self._dev
        host = self._dev.to_numpy() if self._dev is not None else np.asarray(self._host)
        self._st.dev = B200Array.from_numpy(host.astype(want, copy=False), ctx or default_context())
        self._st.host = None
        return self._st.dev

    # -- metadata-only operations -----------------------------------------------------
    def conj(self):
        t = Tensor.__new__(Tensor)
        t._st, t.inds, t._conj = self._st, self.inds, not self._conj
        if t.dtype.kind != "c":
            t._conj = False
        return t

    def replace(self, *pairs, **kw):
        """`replace(t, old => new, ...)`: relabel indices."""
        m = dict(pairs[0]) if len(pairs) == 1 and isinstance(pairs[0], dict) else dict(pairs)
        t = Tensor.__new__(Tensor)
        t._st, t._conj = self._st, self._conj
        t.inds = tuple(m.get(i, i) for i in self.inds)
        if len(set(t.inds)) != len(t.inds):
            raise ValueError("replace would create a repeated index")
        return t

    def permutedims(self, inds: Sequence[Hashable]):
        inds = tuple(inds)
        if set(inds) != set(self.inds) or len(inds) != len(self.inds):
            raise ValueError("permutedims needs a permutation of the tensor's indices")
        perm = [self.inds.index(i) for i in inds]
        t = Tensor.__new__(Tensor)
        t._conj, t.inds = self._conj, inds
        if self._dev is not None:
            t._st = _Storage(None, self._dev.transpose(perm))
        else:
            t._st = _Storage(np.transpose(self._host, perm), None)
        return t

    def view(self, *pairs):
        """`view(t, ind => i)` drops the index, `view(t, ind => a:b)` (a Python slice) keeps it
        (compress.jl:46-58, evolve.jl:64-72, sample.jl:48)."""
        t = self
        for ind, sel in pairs:
            ax = t.inds.index(ind)
            n = Tensor.__new__(Tensor)
            n._conj = t._conj
            if t._dev is not None:
                n._st = _Storage(None, t._dev.view_index(ax, sel))
            else:
                idx = [slice(None)] * t.ndim
                idx[ax] = sel
                n._st = _Storage(t._host[tuple(idx)], None)
            n.inds = t.inds if isinstance(sel, slice) else t.inds[:ax] + t.inds[ax + 1:]
            t = n
        return t

    def __repr__(self):
        where = "B200" if self._dev is not None else "host"
        return f"Tensor({where}, {self.dtype}, shape={tuple(self.shape)}, inds={self.inds!r}{', conj' if self._conj else ''})"


def _mode_map(*ind_lists):
    m = {}
    for inds in ind_lists:
        for i in inds:
            if i not in m:
                m[i] = len(m)
    return m


def _desc(t: Tensor, arr: B200Array, modes):
    return make_desc(arr.buffer.handle, arr.offset, arr.dtype_code, arr.shape, arr.strides,
                     [modes[i] for i in t.inds], t._conj)


def binary_einsum(a: Tensor, b: Tensor, dims: Optional[Iterable[Hashable]] = None,
                  out: Optional[Sequence[Hashable]] = None, ctx: Optional[Context] = None) -> Tensor:
    """Muscle.binary_einsum(a, b; dims, out) on the B200.

    dims=None contracts all shared indices; dims=() keeps them as batch indices; `out` fixes the order of the
    result's indices (default: free(a), free(b), batch).  Returns a device-resident Tensor.
    """
    ctx = ctx or default_context()
    lib = ctx.lib
    shared = [i for i in a.inds if i in b.inds]
    for i in shared:
        if a.size(i) != b.size(i):
            raise ValueError(f"extent mismatch on index {i!r}: {a.size(i)} vs {b.size(i)}")
    dims = list(shared) if dims is None else list(dims)
    for i in dims:
        if i not in a.inds and i not in b.inds:
            raise ValueError(f"dims contains {i!r}, which neither operand carries")
    dt = _promote_dtype(a.dtype, b.dtype)
    A, B = a.device(ctx, dt), b.device(ctx, dt)
    free_a = [i for i in a.inds if i not in b.inds and i not in dims]
    free_b = [i for i in b.inds if i not in a.inds and i not in dims]
    batch = [i for i in shared if i not in dims]
    c_inds = tuple(free_a + free_b + batch)
    if out is not None:
        out = tuple(out)
        if set(out) != set(c_inds) or len(out) != len(c_inds):
            raise ValueError(f"out={out!r} is not a permutation of the result indices {c_inds!r}")
        c_inds = out
    ext = {i: a.size(i) for i in a.inds}
    ext.update({i: b.size(i) for i in b.inds})
    Cdev = B200Array.empty([ext[i] for i in c_inds], dt, ctx)
    c = Tensor(Cdev, c_inds)
    modes = _mode_map(a.inds, b.inds)
    da, ka = _desc(a, A, modes)
    db, kb = _desc(b, B, modes)
    dc, kc = _desc(c, Cdev, modes)
    sm = (C.c_int32 * max(len(dims), 1))(*[modes[i] for i in dims])
    check(ctx.handle, lib.tnb_binary_einsum(ctx.handle, C.byref(da), C.byref(db), C.byref(dc), sm, len(dims), None, None))
    return c


def tensor_qr_thin(a: Tensor, inds_q: Sequence[Hashable], inds_r: Optional[Sequence[Hashable]] = None,
                   ind_virtual: Hashable = "qr_virtual", ctx: Optional[Context] = None):
    """Muscle.tensor_qr_thin(a; inds_q, inds_r, ind_virtual) on the device (canonize.jl:58): a = Q * R with
    Q[inds_q..., ind_virtual] an isometry and R[ind_virtual, inds_r...]; the operand never leaves the GPU
    (tnb_qr_thin: gather -> cuSOLVER geqrf/orgqr -> scatter)."""
    ctx = ctx or default_context()
    inds_q = list(inds_q)
    inds_r = [i for i in a.inds if i not in inds_q] if inds_r is None else list(inds_r)
    if set(inds_q) | set(inds_r) != set(a.inds) or set(inds_q) & set(inds_r) or ind_virtual in a.inds:
        raise ValueError("inds_q / inds_r must partition the tensor's indices and ind_virtual must be new")
    dt = _promote_dtype(a.dtype)
    A = a.device(ctx, dt)
    m = int(np.prod([a.size(i) for i in inds_q], dtype=np.int64)) if inds_q else 1
    n = int(np.prod([a.size(i) for i in inds_r], dtype=np.int64)) if inds_r else 1
    k = min(m, n)
    q = Tensor(B200Array.empty([a.size(i) for i in inds_q] + [k], dt, ctx), tuple(inds_q) + (ind_virtual,))
    r = Tensor(B200Array.empty([k] + [a.size(i) for i in inds_r], dt, ctx), (ind_virtual,) + tuple(inds_r))
    modes = _mode_map(a.inds, [ind_virtual])
    da, ka = _desc(a, A, modes)
    dq, kq = _desc(q, q._dev, modes)
    dr, kr = _desc(r, r._dev, modes)
    # column order of the matrix view = the other modes in a's order; R is scattered into inds_r order by the descriptor
    rows = (C.c_int32 * max(len(inds_q), 1))(*[modes[i] for i in inds_q])
    check(ctx.handle, ctx.lib.tnb_qr_thin(ctx.handle, C.byref(da), rows, len(inds_q), modes[ind_virtual], C.byref(dq), C.byref(dr)))
    return q, r


def tensor_svd_thin(a: Tensor, inds_u: Sequence[Hashable], inds_v: Optional[Sequence[Hashable]] = None,
                    ind_s: Hashable = "svd_virtual", ctx: Optional[Context] = None):
    """Muscle.tensor_svd_thin(a; inds_u, inds_v, ind_s) on the device (canonize.jl:41,97; evolve.jl:62,92;
    DMRG.jl:338,437): a = U * diag(s) * V with U[inds_u..., ind_s], s[ind_s] (real values in a's dtype, descending),
    V[ind_s, inds_v...]."""
    ctx = ctx or default_context()
    inds_u = list(inds_u)
    inds_v = [i for i in a.inds if i not in inds_u] if inds_v is None else list(inds_v)
    if set(inds_u) | set(inds_v) != set(a.inds) or set(inds_u) & set(inds_v) or ind_s in a.inds:
        raise ValueError("inds_u / inds_v must partition the tensor's indices and ind_s must be new")
    dt = _promote_dtype(a.dtype)
    A = a.device(ctx, dt)
    m = int(np.prod([a.size(i) for i in inds_u], dtype=np.int64)) if inds_u else 1
    n = int(np.prod([a.size(i) for i in inds_v], dtype=np.int64)) if inds_v else 1
    k = min(m, n)
    u = Tensor(B200Array.empty([a.size(i) for i in inds_u] + [k], dt, ctx), tuple(inds_u) + (ind_s,))
    sv = Tensor(B200Array.empty([k], dt, ctx), (ind_s,))
    v = Tensor(B200Array.empty([k] + [a.size(i) for i in inds_v], dt, ctx), (ind_s,) + tuple(inds_v))
    modes = _mode_map(a.inds, [ind_s])
    da, ka = _desc(a, A, modes)
    du, ku = _desc(u, u._dev, modes)
    ds, ks = _desc(sv, sv._dev, modes)
    dv, kv = _desc(v, v._dev, modes)
    rows = (C.c_int32 * max(len(inds_u), 1))(*[modes[i] for i in inds_u])
    check(ctx.handle, ctx.lib.tnb_svd_thin(ctx.handle, C.byref(da), rows, len(inds_u), modes[ind_s], C.byref(du), C.byref(ds), C.byref(dv)))
    return u, sv, v
