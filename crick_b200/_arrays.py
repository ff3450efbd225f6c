"""Input coercion for the wrappers: NumPy arrays (any dtype/shape/strides, flattened in
the same K-order the reference's NpyIter uses, tdigest_stubs.c:652-666) and CUDA arrays
via ``__cuda_array_interface__`` or DLPack.  No PyTorch / CuPy import anywhere."""
import ctypes

import numpy as np

from ._lib import check, lib


class DeviceArray:
    """A device buffer owned by libcrick_cuda, exposing ``__cuda_array_interface__`` (v3) and
    DLPack (``__dlpack__`` / ``__dlpack_device__``) so torch / cupy / numba / jax can wrap it
    zero-copy; ``to_numpy()`` copies it to the host."""

    def __init__(self, shape, dtype="f8"):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.size = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        self.nbytes = self.size * self.dtype.itemsize
        self.device = lib.crick_get_device()
        self.ptr = lib.crick_device_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise MemoryError("crick_device_alloc(%d) failed" % self.nbytes)

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr, False),
                "version": 3, "strides": None}

    def __dlpack_device__(self):
        return (_KDLCUDA, self.device)

    def __dlpack__(self, stream=None, **_):
        # whichever of the library's own streams wrote this array has finished before the
        # consumer (on any stream of its own) may read it
        check(lib.crick_device_synchronize(), "synchronize")
        return dlpack_capsule(self.ptr, self.shape, self.dtype, self.device, self)

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


# ---- DLPack producer: a DLManagedTensor built with ctypes, kept alive until its deleter runs ----
_DL_DELETER = ctypes.CFUNCTYPE(None, ctypes.c_void_p)


class _DLManagedTensor(ctypes.Structure):
    _fields_ = [("dl_tensor", _DLTensor), ("manager_ctx", ctypes.c_void_p), ("deleter", _DL_DELETER)]


_dl_live = {}          # address of the DLManagedTensor -> (struct, shape array, owner)


@_DL_DELETER
def _dl_deleter(addr):
    _dl_live.pop(addr, None)


_CAPSULE_DESTRUCTOR = ctypes.CFUNCTYPE(None, ctypes.c_void_p)
_capsule_api = ctypes.pythonapi
_PyCapsule_New = _capsule_api.PyCapsule_New
_PyCapsule_New.restype = ctypes.py_object
_PyCapsule_New.argtypes = [ctypes.c_void_p, ctypes.c_char_p, _CAPSULE_DESTRUCTOR]
_PyCapsule_IsValid = ctypes.PYFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_char_p)(
    ("PyCapsule_IsValid", ctypes.pythonapi))
_PyCapsule_GetPointerRaw = ctypes.PYFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p)(
    ("PyCapsule_GetPointer", ctypes.pythonapi))
_DLTENSOR_NAME = b"dltensor"    # (module-level: the capsule keeps a pointer to these bytes)


@_CAPSULE_DESTRUCTOR
def _capsule_destructor(capsule):
    # a consumer renames the capsule to "used_dltensor" and calls the deleter itself; a capsule
    # nobody consumed still carries the original name and is released here
    if _PyCapsule_IsValid(capsule, _DLTENSOR_NAME):
        _dl_deleter(_PyCapsule_GetPointerRaw(capsule, _DLTENSOR_NAME))


def dlpack_capsule(ptr, shape, dtype, device, owner):
    """A DLPack capsule ("dltensor") for a C-contiguous CUDA buffer; `owner` stays alive until
    the consumer calls the tensor's deleter."""
    dtype = np.dtype(dtype)
    code = {"i": 0, "u": 1, "f": 2, "b": 6}.get(dtype.kind)
    if code is None:
        raise TypeError("dtype %s has no DLPack type code" % dtype)
    shape = tuple(int(v) for v in shape)
    arr = (ctypes.c_int64 * max(len(shape), 1))(*shape)
    m = _DLManagedTensor()
    m.dl_tensor.data = ctypes.c_void_p(int(ptr))
    m.dl_tensor.device = _DLDevice(_KDLCUDA, int(device))
    m.dl_tensor.ndim = len(shape)
    m.dl_tensor.dtype = _DLDataType(code, dtype.itemsize * 8, 1)
    m.dl_tensor.shape = ctypes.cast(arr, ctypes.POINTER(ctypes.c_int64))
    m.dl_tensor.strides = ctypes.POINTER(ctypes.c_int64)()      # NULL: compact row-major
    m.dl_tensor.byte_offset = 0
    m.manager_ctx = None
    m.deleter = _dl_deleter
    addr = ctypes.addressof(m)
    _dl_live[addr] = (m, arr, owner)
    return _PyCapsule_New(addr, _DLTENSOR_NAME, _capsule_destructor)


class DeviceView:
    """(pointer, size, dtype) of a C-contiguous CUDA array; keeps its owner alive.
    ``foreign``: the memory belongs to another library (torch, cupy, numba ...), whose allocator
    may hand it to someone else as soon as the owner is dropped -- see :func:`release_views`."""

    def __init__(self, ptr, shape, dtype, owner, foreign=True):
        self.ptr = int(ptr) if ptr else 0
        self.shape = tuple(int(s) for s in shape)
        self.size = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        self.dtype = np.dtype(dtype)
        self.owner = owner
        self.foreign = foreign


def _contiguous(shape, strides, itemsize):
    if strides is None:
        return True
    expect = itemsize
    for n, s in zip(reversed(shape), reversed(strides)):
        if n != 1 and s != expect:
            return False
        expect *= max(n, 1)
    return True


def _producer_sync(stream_field):
    """``stream`` entry of a ``__cuda_array_interface__`` v3 dict: the consumer has to order its
    work behind that stream (1 = legacy default stream, 2 = per-thread default stream, another
    integer = a cudaStream_t; None = nothing to wait for)."""
    if stream_field is None:
        return
    if stream_field in (1, 2):
        check(lib.crick_device_synchronize() if stream_field == 2 else lib.crick_stream_synchronize(None),
              "synchronize")
    else:
        check(lib.crick_stream_synchronize(int(stream_field)), "synchronize")


def device_view(obj, stream=None):
    """Return a DeviceView if `obj` is a CUDA array, else None.

    Stream ordering with the array's producer: `stream` is the cudaStream_t (int) the caller is
    going to enqueue the library's work on.  DLPack objects (torch, cupy, jax) are asked for a
    capsule FOR that stream (``__dlpack__(stream=...)``: the producer makes that stream wait for
    its pending writes); without one the capsule is requested for the legacy default stream and
    that stream is synchronised here, because the library's own streams are non-blocking.
    ``__cuda_array_interface__`` objects are waited for through their ``stream`` entry (v3)."""
    if isinstance(obj, DeviceArray):
        return DeviceView(obj.ptr, obj.shape, obj.dtype, obj, foreign=False)
    has_dl = hasattr(obj, "__dlpack__") and hasattr(obj, "__dlpack_device__") and not isinstance(obj, np.ndarray)
    cai = None if has_dl else getattr(obj, "__cuda_array_interface__", None)
    if cai is not None:
        shape = tuple(cai["shape"])
        dtype = np.dtype(cai["typestr"])
        if not _contiguous(shape, cai.get("strides"), dtype.itemsize):
            raise ValueError("CUDA arrays must be C-contiguous")
        _producer_sync(cai.get("stream"))
        return DeviceView(cai["data"][0], shape, dtype, obj)
    if has_dl:
        dev_type = int(obj.__dlpack_device__()[0])
        if dev_type not in (_KDLCUDA, _KDLCUDA_MANAGED):
            return None
        if stream:
            cap = obj.__dlpack__(stream=int(stream))
        else:
            try:
                cap = obj.__dlpack__(stream=1)           # "the consumer works on the legacy default stream"
            except TypeError:
                cap = obj.__dlpack__()
            check(lib.crick_stream_synchronize(None), "synchronize")
        p = ctypes.pythonapi.PyCapsule_GetPointer(cap, b"dltensor")
        t = ctypes.cast(p, ctypes.POINTER(_DLTensor)).contents
        code, bits = t.dtype.code, t.dtype.bits
        kind = {0: "i", 1: "u", 2: "f", 6: "b"}.get(code)     # (6 = kDLBool: one byte)
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


def release_views(views, stream):
    """Called after the library's work on CUDA arrays of ANOTHER library has been enqueued.  With
    a caller-provided `stream` the usual CUDA rule applies (the caller keeps the arrays alive, or
    tells its allocator, until that stream has passed the call).  Without one the work ran on a
    stream only this library knows, so a caching allocator (torch, cupy) could recycle the
    arrays' memory while the kernels still read it: wait for them here."""
    if stream:
        return
    if any(v is not None and v.foreign for v in views):
        check(lib.crick_device_synchronize(), "synchronize")


def as_f64_device(view, stream=None):
    """A float64 DeviceView of `view`: itself when it already is float64, else a safe cast on the
    device into a library-owned buffer (the reference accepts every dtype that
    ``np.can_cast(dtype, 'f8', 'safe')`` allows, tdigest.pyx:298-302; complex, strings, datetimes
    raise TypeError like there)."""
    if view.dtype == np.float64:
        return view
    if view.dtype.kind not in "fiub" or not np.can_cast(view.dtype, "f8", casting="safe"):
        raise TypeError("CUDA array of dtype %s cannot be safely cast to float64" % view.dtype)
    out = DeviceArray(view.shape, "f8")
    if view.size:
        check(lib.crick_cast_to_f64(view.ptr, ord(view.dtype.kind), view.dtype.itemsize, view.size, out.ptr,
                                    stream), "cast")
        if not stream:
            # the cast ran on the legacy stream; the update runs on the handle's own
            check(lib.crick_stream_synchronize(None), "synchronize")
    v = DeviceView(out.ptr, view.shape, np.float64, (out, view.owner), foreign=view.foreign)
    return v


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
