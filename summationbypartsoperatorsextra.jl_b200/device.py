"""Device context and device-resident batch arrays.

A *batch* is the data-parallel axis the reference does not have: B independent nodal vectors that share
one operator.  Layout: Julia `Matrix{Float64}(N, B)` column-major == numpy (B, N) C-order; instance b is
contiguous at offset b*N.  `DeviceBatch` only wraps a device pointer; `mul_(du, D, u)`/`semi(du, u, p, t)`
dispatch on it with no copies (host<->device transfer happens only in `upload`/`to_host`).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ArgumentError, DimensionMismatch, check


class Context:
    """One per device (wraps `sbp_ctx`): a CUDA stream, scratch for reductions, optional NCCL communicator."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        check(self._lib.sbp_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        sm, l2, smem, glob = C.c_int32(), C.c_int64(), C.c_int64(), C.c_int64()
        check(self._lib.sbp_ctx_device_info(h, C.byref(sm), C.byref(l2), C.byref(smem), C.byref(glob)))
        self.sm_count, self.l2_bytes, self.smem_optin, self.global_bytes = sm.value, l2.value, smem.value, glob.value
        self._comm = False

    # -- lifetime ------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.sbp_ctx_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if not self._h:
            raise ArgumentError("context is closed")
        return self._h

    def synchronize(self):
        check(self._lib.sbp_ctx_sync(self.handle))

    def set_stream(self, cuda_stream: int | None):
        """Launch on a caller-owned cudaStream_t (e.g. `torch.cuda.current_stream().cuda_stream`)."""
        check(self._lib.sbp_ctx_set_stream(self.handle, C.c_void_p(cuda_stream or 0)))

    @property
    def stream(self) -> int:
        s = C.c_void_p()
        check(self._lib.sbp_ctx_get_stream(self.handle, C.byref(s)))
        return s.value or 0

    def set_option(self, name: str, value: int):
        """Tuning / debugging option of this context (`sbp_ctx_set_option`; names: csrc/common.cuh enum Tune)."""
        check(self._lib.sbp_ctx_set_option(self.handle, name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        v = C.c_int32()
        check(self._lib.sbp_ctx_get_option(self.handle, name.encode(), C.byref(v)))
        return v.value

    def options(self, **kv):
        """Context manager: set options for the duration of a `with` block, then restore the previous values."""
        import contextlib

        @contextlib.contextmanager
        def scope():
            old = {k: self.get_option(k) for k in kv}
            try:
                for k, v in kv.items():
                    self.set_option(k, v)
                yield self
            finally:
                for k, v in old.items():
                    self.set_option(k, v)

        return scope()

    @property
    def launch_count(self) -> int:
        return int(self._lib.sbp_ctx_launch_count(self.handle))

    # -- memory --------------------------------------------------------------------------------------
    def malloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        check(self._lib.sbp_malloc(self.handle, int(nbytes), C.byref(p)))
        return p.value

    def free(self, dptr: int):
        check(self._lib.sbp_free(self.handle, C.c_void_p(dptr)))

    def upload(self, host: np.ndarray) -> "DeviceBatch":
        """Copy a host array of shape (N,) or (B, N) to a new device batch."""
        a = np.ascontiguousarray(host, dtype=np.float64)
        if a.ndim == 1:
            a = a[None, :]
        if a.ndim != 2:
            raise DimensionMismatch("batch arrays are (B, N) (or a single vector (N,))")
        out = DeviceBatch.empty(self, a.shape[0], a.shape[1])
        if a.size:
            check(self._lib.sbp_memcpy_h2d(self.handle, C.c_void_p(out.ptr), a.ctypes.data_as(C.c_void_p), a.nbytes))
            self.synchronize()
        return out

    def upload_array(self, host: np.ndarray):
        """Raw device copy of any contiguous host array; returns (ptr, owner) with owner.free()."""
        a = np.ascontiguousarray(host)
        raw = RawBuffer(self, max(a.nbytes, 16))
        if a.nbytes:
            check(self._lib.sbp_memcpy_h2d(self.handle, C.c_void_p(raw.ptr), a.ctypes.data_as(C.c_void_p), a.nbytes))
            self.synchronize()
        return raw

    def pci_bus_id(self) -> str:
        buf = C.create_string_buffer(32)
        check(self._lib.sbp_ctx_pci_bus_id(self.handle, buf, 32))
        return buf.value.decode()

    def bind_host_to_gpu_numa(self) -> list[int]:
        """Pin the calling process to the CPUs of the GPU's NUMA node (sysfs `local_cpulist` of the device, intersected
        with the current affinity mask) so that pinned buffers allocated afterwards and the copy-issuing thread are local
        to the GPU's PCIe root.  Returns the CPU list in effect (unchanged when sysfs has no answer)."""
        import os
        cur = sorted(os.sched_getaffinity(0))
        try:
            with open(f"/sys/bus/pci/devices/{self.pci_bus_id()}/local_cpulist") as f:
                txt = f.read().strip()
            cpus = set()
            for part in txt.split(","):
                if not part:
                    continue
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            want = sorted(cpus & set(cur))
            if want and len(want) < len(cur):
                os.sched_setaffinity(0, want)
                return want
        except (OSError, ValueError):
            pass
        return cur

    def pcie_probe(self, nbytes: int = 256 << 20) -> dict:
        """Pinned-memory copy rates of this box in GB/s: {"h2d", "d2h", "bidir"} (`sbp_pcie_probe`)."""
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        check(self._lib.sbp_pcie_probe(self.handle, int(nbytes), C.byref(a), C.byref(b), C.byref(c)))
        return {"h2d": a.value, "d2h": b.value, "bidir": c.value}

    def pinned_empty(self, shape) -> np.ndarray:
        """Page-locked host array (for the end-to-end host-buffer path); release with `free_pinned`."""
        n = int(np.prod(shape))
        p = C.c_void_p()
        check(self._lib.sbp_host_alloc(max(n, 1) * 8, C.byref(p)))
        buf = (C.c_double * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape)
        _PINNED[p.value] = buf
        return arr

    def free_pinned(self, arr: np.ndarray):
        ptr = arr.ctypes.data
        if ptr in _PINNED:
            _PINNED.pop(ptr)
            check(self._lib.sbp_host_free(C.c_void_p(ptr)))

    # -- multi-GPU: one process per GPU -----------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(_lib.NCCL_ID_BYTES)
        check(self._lib.sbp_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        if len(unique_id) != _lib.NCCL_ID_BYTES:
            raise ArgumentError("NCCL unique id must be 128 bytes")
        buf = C.create_string_buffer(unique_id, _lib.NCCL_ID_BYTES)
        check(self._lib.sbp_comm_init(self.handle, int(nranks), int(rank), buf))
        self._comm = True

    def allreduce_sum(self, dptr: int, count: int):
        check(self._lib.sbp_allreduce_sum(self.handle, C.c_void_p(dptr), int(count)))


_PINNED: dict = {}


class RawBuffer:
    def __init__(self, ctx: Context, nbytes: int):
        self.ctx, self.nbytes = ctx, int(nbytes)
        self.ptr = ctx.malloc(nbytes)

    def free(self):
        if self.ptr:
            self.ctx.free(self.ptr)
            self.ptr = 0

    def to_host(self, dtype, count) -> np.ndarray:
        out = np.empty(count, dtype=dtype)
        check(self.ctx._lib.sbp_memcpy_d2h(self.ctx.handle, out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr),
                                           out.nbytes))
        return out

    def __del__(self):  # pragma: no cover
        try:
            self.free()
        except Exception:
            pass


class DeviceBatch:
    """B nodal vectors of N nodes on the device (Float64).  `shape == (B, N)`."""

    def __init__(self, ctx: Context, ptr: int, B: int, N: int, owner=None):
        self.ctx, self.ptr, self.B, self.N = ctx, int(ptr), int(B), int(N)
        self._owner = owner  # RawBuffer, torch tensor, ... (keeps the memory alive)

    @classmethod
    def empty(cls, ctx: Context, B: int, N: int) -> "DeviceBatch":
        raw = RawBuffer(ctx, max(int(B) * int(N) * 8, 16))
        return cls(ctx, raw.ptr, B, N, raw)

    @classmethod
    def zeros(cls, ctx: Context, B: int, N: int) -> "DeviceBatch":
        out = cls.empty(ctx, B, N)
        check(ctx._lib.sbp_memset(ctx.handle, C.c_void_p(out.ptr), 0, max(out.nbytes, 0)))
        return out

    @classmethod
    def from_torch(cls, ctx: Context, t) -> "DeviceBatch":
        """Zero-copy view of a contiguous float64 CUDA tensor of shape (B, N) or (N,)."""
        if t.dtype.__str__() != "torch.float64" or not t.is_cuda or not t.is_contiguous():
            raise ArgumentError("need a contiguous float64 CUDA tensor")
        if t.dim() == 1:
            B, N = 1, t.shape[0]
        elif t.dim() == 2:
            B, N = t.shape
        else:
            raise DimensionMismatch("need a (B, N) or (N,) tensor")
        return cls(ctx, t.data_ptr(), B, N, t)

    @property
    def shape(self):
        return (self.B, self.N)

    @property
    def nbytes(self):
        return self.B * self.N * 8

    def __len__(self):
        return self.N if self.B == 1 else self.B

    def to_host(self) -> np.ndarray:
        out = np.empty((self.B, self.N), dtype=np.float64)
        if out.size:
            check(self.ctx._lib.sbp_memcpy_d2h(self.ctx.handle, out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr),
                                               out.nbytes))
        return out

    def copy_from_host(self, host: np.ndarray):
        a = np.ascontiguousarray(host, dtype=np.float64).reshape(-1)
        if a.size != self.B * self.N:
            raise DimensionMismatch(f"host array has {a.size} entries, batch holds {self.B * self.N}")
        if a.size:
            check(self.ctx._lib.sbp_memcpy_h2d(self.ctx.handle, C.c_void_p(self.ptr), a.ctypes.data_as(C.c_void_p),
                                               a.nbytes))
            self.ctx.synchronize()
        return self

    def slice(self, b0: int, b1: int) -> "DeviceBatch":
        """Instances [b0, b1) as a view (contiguous in this layout)."""
        if not (0 <= b0 <= b1 <= self.B):
            raise ArgumentError("slice out of range")
        return DeviceBatch(self.ctx, self.ptr + b0 * self.N * 8, b1 - b0, self.N, self)

    def similar(self) -> "DeviceBatch":
        return DeviceBatch.empty(self.ctx, self.B, self.N)

    def copy(self) -> "DeviceBatch":
        out = self.similar()
        check(self.ctx._lib.sbp_memset(self.ctx.handle, C.c_void_p(out.ptr), 0, max(out.nbytes, 0)))
        check(self.ctx._lib.sbp_axpbypcz(self.ctx.handle, C.c_void_p(out.ptr), C.c_void_p(self.ptr), None,
                                         self.B * self.N, 1.0, 0.0, 0.0))
        return out

    def axpby_(self, a: float, x: "DeviceBatch", b: float = 1.0, c: float = 0.0, z: "DeviceBatch | None" = None):
        """self <- a*x + b*self (+ c*z): Runge-Kutta stage updates without leaving the device."""
        if x.shape != self.shape or (z is not None and z.shape != self.shape):
            raise DimensionMismatch("axpby_: shapes differ")
        check(self.ctx._lib.sbp_axpbypcz(self.ctx.handle, C.c_void_p(self.ptr), C.c_void_p(x.ptr),
                                         C.c_void_p(z.ptr) if z is not None else None, self.B * self.N, float(a),
                                         float(b), float(c)))
        return self

    def lincomb_(self, a: float, x: "DeviceBatch", c: float = 0.0, z: "DeviceBatch | None" = None, d: float = 0.0,
                 w: "DeviceBatch | None" = None, b: float = 0.0):
        """self <- a*x + c*z + d*w (+ b*self) in one pass over memory (`sbp_lincomb`)."""
        for v in (x, z, w):
            if v is not None and v.shape != self.shape:
                raise DimensionMismatch("lincomb_: shapes differ")
        check(self.ctx._lib.sbp_lincomb(self.ctx.handle, C.c_void_p(self.ptr), C.c_void_p(x.ptr),
                                        C.c_void_p(z.ptr) if z is not None else None,
                                        C.c_void_p(w.ptr) if w is not None else None, self.B * self.N, float(a),
                                        float(b), float(c), float(d)))
        return self

    def free(self):
        if isinstance(self._owner, RawBuffer):
            self._owner.free()
        self.ptr = 0
