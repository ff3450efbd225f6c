"""Input coercion for the wrappers: NumPy arrays (any dtype/shape/strides, flattened in
the same K-order the reference's NpyIter uses, tdigest_stubs.c:652-666) and CUDA arrays
via ``__cuda_array_interface__`` or DLPack.  No PyTorch / CuPy import anywhere."""
import ctypes

import numpy as np

from ._lib import check, lib


class DeviceArray:
    """A device buffer owned by libcrick_cuda, exposing ``__cuda_array_interface__`` (v3) so
    torch / cupy / numba can wrap it zero-copy; ``to_numpy()`` copies it to the host."""

    def __init__(self, shape, dtype="f8"):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.size = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        self.nbytes = self.size * self.dtype.itemsize
        self.ptr = lib.crick_device_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise MemoryError("crick_device_alloc(%d) failed" % self.nbytes)

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr, False),
                "version": 3, "strides": None}

    def to_numpy(self):
        out = np.empty(self.shape, dtype=self.dtype)
        if self.nbytes:
            # whichever stream wrote this array (handles own non-blocking streams) has finished
            check(lib.crick_device_synchronize(), "synchronize")
            check(lib.crick_memcpy_d2h(out.ctypes.data, self.ptr, self.nbytes, None), "memcpy d2h")
        return out

    @classmethod
    def from_numpy(cls, a):
        a = np.ascontiguousarray(a)
        d = cls(a.shape, a.dtype)
        if d.nbytes:
            check(lib.crick_memcpy_h2d(d.ptr, a.ctypes.data, d.nbytes, None), "memcpy h2d")
        return d

    def __del__(self):
        p, self.ptr = getattr(self, "ptr", None), None
        if p:
            lib.crick_device_free(p)


class PinnedArray:
    """Page-locked host buffer viewed as a NumPy array (``.array``): H2D copies from it are
    real asynchronous DMA."""

    def __init__(self, n, dtype="f8"):
        self.dtype = np.dtype(dtype)
        self.nbytes = int(n) * self.dtype.itemsize
        self.ptr = lib.crick_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise MemoryError("crick_host_alloc(%d) failed" % self.nbytes)
        buf = (ctypes.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(n))

    def __del__(self):
        p, self.ptr = getattr(self, "ptr", None), None
        self.array = None
        if p:
            lib.crick_host_free(p)


# ---- DLPack (no framework import: parse the capsule with ctypes) --------------------
class _DLDevice(ctypes.Structure):
    _fields_ = [("device_type", ctypes.c_int), ("device_id", ctypes.c_int)]


class _DLDataType(ctypes.Structure):
    _fields_ = [("code", ctypes.c_uint8), ("bits", ctypes.c_uint8), ("lanes", ctypes.c_uint16)]


class _DLTensor(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("device", _DLDevice), ("ndim", ctypes.c_int),
                ("dtype", _DLDataType), ("shape", ctypes.POINTER(ctypes.c_int64)),
                ("strides", ctypes.POINTER(ctypes.c_int64)), ("byte_offset", ctypes.c_uint64)]


_KDLCUDA, _KDLCUDA_MANAGED = 2, 13
ctypes.pythonapi.PyCapsule_GetPointer.restype = ctypes.c_void_p
ctypes.pythonapi.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]


class DeviceView:
    """(pointer, size, dtype) of a C-contiguous CUDA array; keeps its owner alive."""

    def __init__(self, ptr, shape, dtype, owner):
        self.ptr = int(ptr) if ptr else 0
        self.shape = tuple(int(s) for s in shape)
        self.size = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        self.dtype = np.dtype(dtype)
        self.owner = owner


def _contiguous(shape, strides, itemsize):
    if strides is None:
        return True
    expect = itemsize
    for n, s in zip(reversed(shape), reversed(strides)):
        if n != 1 and s != expect:
            return False
        expect *= max(n, 1)
    return True


def device_view(obj):
    """Return a DeviceView if `obj` is a CUDA array, else None."""
    cai = getattr(obj, "__cuda_array_interface__", None)
    if cai is not None:
        shape = tuple(cai["shape"])
        dtype = np.dtype(cai["typestr"])
        if not _contiguous(shape, cai.get("strides"), dtype.itemsize):
            raise ValueError("CUDA arrays must be C-contiguous")
        return DeviceView(cai["data"][0], shape, dtype, obj)
    if hasattr(obj, "__dlpack__") and hasattr(obj, "__dlpack_device__") and not isinstance(obj, np.ndarray):
        dev_type = int(obj.__dlpack_device__()[0])
        if dev_type not in (_KDLCUDA, _KDLCUDA_MANAGED):
            return None
        cap = obj.__dlpack__()
        p = ctypes.pythonapi.PyCapsule_GetPointer(cap, b"dltensor")
        t = ctypes.cast(p, ctypes.POINTER(_DLTensor)).contents
        code, bits = t.dtype.code, t.dtype.bits
        kind = {0: "i", 1: "u", 2: "f"}.get(code)
        if kind is None or t.dtype.lanes != 1:
            raise TypeError("unsupported DLPack dtype")
        dtype = np.dtype("%s%d" % (kind, bits // 8))
        shape = tuple(t.shape[i] for i in range(t.ndim))
        strides = None
        if t.strides:
            strides = tuple(t.strides[i] * dtype.itemsize for i in range(t.ndim))
        if not _contiguous(shape, strides, dtype.itemsize):
            raise ValueError("CUDA arrays must be C-contiguous")
        # the un-consumed capsule owns the tensor; keep it (and obj) alive with the view
        return DeviceView((t.data or 0) + t.byte_offset, shape, dtype, (obj, cap))
    return None


def korder_pair(x, w, w_dtype="f8", x_dtype="f8"):
    """Flatten broadcast(x, w) to two contiguous 1-D arrays in the order the reference's
    buffered NpyIter visits them (NPY_KEEPORDER: memory order, negative strides flipped only
    when every operand allows it).  `w` may be 0-d: then only x is materialised and the
    scalar is returned.  Raises TypeError on unsafe casts like the iterator would."""
    x = np.asarray(x)
    w = np.asarray(w)
    xd, wd = np.dtype(x_dtype), np.dtype(w_dtype)
    if x.size == 0 or w.size == 0:
        return np.empty(0, xd), (np.empty(0, wd) if w.ndim else w.astype(wd, casting="safe")[()])
    fast_x = x.flags.c_contiguous and x.dtype == xd
    if w.ndim == 0 and fast_x:
        return x.reshape(-1), w.astype(wd, casting="safe")[()]
    if fast_x and w.shape == x.shape and w.flags.c_contiguous and w.dtype == wd:
        return x.reshape(-1), w.reshape(-1)
    it = np.nditer([x, w], flags=["external_loop", "buffered"],
                   op_flags=[["readonly", "aligned"], ["readonly", "aligned"]], order="K",
                   casting="safe", op_dtypes=[xd, wd])
    xs, ws = [], []
    for a, b in it:
        xs.append(a.copy())
        if w.ndim:
            ws.append(b.copy())
    xf = np.concatenate(xs) if len(xs) > 1 else xs[0]
    if w.ndim == 0:
        return xf, w.astype(wd, casting="safe")[()]
    wf = np.concatenate(ws) if len(ws) > 1 else ws[0]
    return xf, wf
