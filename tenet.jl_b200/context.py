"""Device context and device array (`B200Array`) — the storage type that selects this backend.

In the reference a backend is chosen by the array storage type behind a `Tensor` (the Reactant manual shows the
pattern: `adapt(ConcreteRArray, psi)` swaps storage and everything else is unchanged,
/root/reference/docs/src/manual/reactant.md:23-36).  `B200Array` plays that role here: a strided view of a device
buffer owned by libtnb200's stream-ordered caching allocator.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import DTYPE_CODE, check, load_library


class Context:
    """One per (process, device): stream + allocator + error slot (tnb_ctx)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.tnb_ctx_create(int(device), C.byref(h))
        if rc != _lib.TNB_OK:
            raise _lib.TnbError(rc, (self.lib.tnb_last_error(None) or b"").decode())
        self.handle = h
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.tnb_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover - interpreter shutdown order
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.handle, self.lib.tnb_sync(self.handle))

    def set_option(self, option: int, value: int):
        check(self.handle, self.lib.tnb_ctx_set_option(self.handle, option, value))

    @property
    def stream(self) -> int:
        return int(self.lib.tnb_ctx_stream(self.handle) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.lib.tnb_ctx_launch_count(self.handle))

    @property
    def last_kernel(self) -> str:
        return _lib.KERNEL_NAMES.get(int(self.lib.tnb_ctx_last_kernel(self.handle)), "none")

    def mem_stats(self):
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        check(self.handle, self.lib.tnb_mem_stats(self.handle, C.byref(a), C.byref(b), C.byref(c)))
        return {"in_use": a.value, "cached": b.value, "peak": c.value}

    def trim(self):
        check(self.handle, self.lib.tnb_mem_trim(self.handle))

    # -- multi-GPU (one process per GPU) ---------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        check(self.handle, self.lib.tnb_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, uid: bytes, rank: int, nranks: int):
        buf = C.create_string_buffer(uid, 128)
        check(self.handle, self.lib.tnb_comm_init(self.handle, buf, rank, nranks))

    def comm_destroy(self):
        check(self.handle, self.lib.tnb_comm_destroy(self.handle))


_default_ctx = {}


def default_context(device=None) -> Context:
    """Process-wide context for `device` (default: LOCAL_RANK under torchrun, else 0)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    ctx = _default_ctx.get(device)
    if ctx is None or ctx.handle is None:
        ctx = _default_ctx[device] = Context(device)
    return ctx


class _Buffer:
    """Owns one tnb_buf; released to the caching allocator when the last view dies."""

    def __init__(self, ctx: Context, nbytes: int):
        self.ctx = ctx
        h = C.c_void_p()
        check(ctx.handle, ctx.lib.tnb_alloc(ctx.handle, max(int(nbytes), 1), C.byref(h)))
        self.handle = h
        self.nbytes = int(nbytes)

    def __del__(self):  # safe from any thread: tnb_free only takes the ctx mutex
        try:
            if self.handle and self.ctx.handle:
                self.ctx.lib.tnb_free(self.ctx.handle, self.handle)
        except Exception:
            pass
        self.handle = None

    @property
    def ptr(self) -> int:
        return int(self.ctx.lib.tnb_buf_ptr(self.handle) or 0)


def fortran_strides(shape):
    st, s = [], 1
    for d in shape:
        st.append(s)
        s *= int(d)
    return tuple(st)


class B200Array:
    """Strided view (element strides, element offset) of a device buffer.  Column-major by default, like Julia."""

    def __init__(self, ctx, buffer, shape, strides, offset, dtype):
        self.ctx, self.buffer = ctx, buffer
        self.shape = tuple(int(x) for x in shape)
        self.strides = tuple(int(x) for x in strides)
        self.offset = int(offset)
        self.dtype = np.dtype(dtype)
        if self.dtype not in DTYPE_CODE:
            raise TypeError(f"unsupported element type {self.dtype}; supported: complex128, complex64, float64, float32")

    # -- construction -----------------------------------------------------------------------------
    @classmethod
    def empty(cls, shape, dtype, ctx=None):
        ctx = ctx or default_context()
        dtype = np.dtype(dtype)
        n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
        buf = _Buffer(ctx, n * dtype.itemsize)
        return cls(ctx, buf, shape, fortran_strides(shape), 0, dtype)

    @classmethod
    def zeros(cls, shape, dtype, ctx=None):
        a = cls.empty(shape, dtype, ctx)
        check(a.ctx.handle, a.ctx.lib.tnb_memset_zero(a.ctx.handle, a.buffer.handle, 0, a.size * a.dtype.itemsize))
        return a

    @classmethod
    def from_numpy(cls, arr, ctx=None):
        arr = np.asarray(arr)
        if arr.dtype not in DTYPE_CODE:
            raise TypeError(f"unsupported element type {arr.dtype}")
        a = cls.empty(arr.shape, arr.dtype, ctx)
        host = np.asfortranarray(arr)
        if host.ndim == 0:
            host = host.reshape(1)
        check(a.ctx.handle, a.ctx.lib.tnb_upload(a.ctx.handle, a.buffer.handle, 0,
                                                   host.ctypes.data_as(C.c_void_p), host.nbytes))
        a._keep = host  # the copy is stream-ordered; keep the staging array alive until it is consumed
        return a

    # -- properties ---------------------------------------------------------------------------------
    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1

    @property
    def dtype_code(self):
        return DTYPE_CODE[self.dtype]

    @property
    def __cuda_array_interface__(self):
        isz = self.dtype.itemsize
        return {"shape": self.shape, "typestr": self.dtype.str, "version": 3,
                "strides": tuple(s * isz for s in self.strides) if self.shape else None,
                "data": (self.buffer.ptr + self.offset * isz, False), "stream": self.ctx.stream or None}

    # -- views / host copies ------------------------------------------------------------------------
    def view_index(self, axis: int, sel):
        """`view(array, :, i, :)` / `view(array, :, a:b, :)`: int drops the axis, slice keeps it."""
        shape, strides = list(self.shape), list(self.strides)
        if isinstance(sel, slice):
            start, stop, step = sel.indices(shape[axis])
            n = max(0, (stop - start + (step - 1 if step > 0 else step + 1)) // step)
            off = self.offset + start * strides[axis]
            shape[axis], strides[axis] = n, strides[axis] * step
        else:
            i = int(sel)
            if not -shape[axis] <= i < shape[axis]:
                raise IndexError("index out of range")
            i %= shape[axis]
            off = self.offset + i * strides[axis]
            del shape[axis], strides[axis]
        return B200Array(self.ctx, self.buffer, shape, strides, off, self.dtype)

    def transpose(self, perm):
        return B200Array(self.ctx, self.buffer, [self.shape[p] for p in perm], [self.strides[p] for p in perm],
                         self.offset, self.dtype)

    def to_numpy(self) -> np.ndarray:
        isz = self.dtype.itemsize
        lo = hi = self.offset
        for d, s in zip(self.shape, self.strides):
            span = (d - 1) * s
            if span > 0:
                hi += span
            else:
                lo += span
        raw = np.empty(hi - lo + 1, dtype=self.dtype)
        check(self.ctx.handle, self.ctx.lib.tnb_download(self.ctx.handle, self.buffer.handle, lo * isz,
                                                          raw.ctypes.data_as(C.c_void_p), raw.nbytes))
        v = np.lib.stride_tricks.as_strided(raw[self.offset - lo:], shape=self.shape,
                                            strides=tuple(s * isz for s in self.strides))
        return np.array(v, order="F")
