"""SpaceSaving -- drop-in for ``crick.SpaceSaving`` (reference: crick/space_saving.pyx).  The
int64 and float64 item types run on the GPU.  The object dtype needs ``PyObject_Hash`` /
``RichCompare`` per item (space_saving_stubs.c.in:20-29), which only the interpreter can
evaluate: that one dtype is a host-side Stream-Summary (``_spsv_object.py``)."""
import numpy as np

from ._arrays import device_view, korder_pair, release_views
from ._lib import check, check_handle, handle_array, lib
from ._spsv_object import ObjectSummary

COUNTER_OBJECT_DTYPE = np.dtype([("item", object), ("count", np.int64), ("error", np.int64)])
COUNTER_INT64_DTYPE = np.dtype([("item", np.int64), ("count", np.int64), ("error", np.int64)])
COUNTER_FLOAT64_DTYPE = np.dtype([("item", np.float64), ("count", np.int64), ("error", np.int64)])


class TopKResult(tuple):
    """TopKResult(item, count, error)

    A result from ``SpaceSaving.topk``: the item, its estimated frequency, and an upper
    bound on the error of that estimate.
    """

    __slots__ = ()

    def __new__(cls, item, count, error):
        return tuple.__new__(cls, (item, count, error))

    def __repr__(self):
        return "TopKResult(item=%r, count=%r, error=%r)" % self

    def __getnewargs__(self):
        return tuple(self)

    def __dict__(self):
        return dict(zip(("item", "count", "error"), self))

    def __getstate__(self):
        pass

    @property
    def item(self):
        return self[0]

    @property
    def count(self):
        return self[1]

    @property
    def error(self):
        return self[2]


def _as_int(v, name):
    """Cython's conversion of an ``int`` argument (space_saving.pyx:179,247,304): numbers go
    through ``__int__`` (NaN -> ValueError, +-inf -> OverflowError, both from float.__int__),
    anything without one is a TypeError."""
    if isinstance(v, (str, bytes, bytearray)) or v is None:
        raise TypeError("%s must be an integer, not %s" % (name, type(v).__name__))
    try:
        return int(v)
    except TypeError:
        raise TypeError("%s must be an integer, not %s" % (name, type(v).__name__))


class SpaceSaving:
    """SpaceSaving(capacity=20, dtype='f8')

    Approximate TopK: keeps ``capacity`` counters for the most frequent elements of a
    stream (Metwally et al. Space-Saving / Stream-Summary; merge by Cafaro et al.).
    Integer dtypes are tracked as int64, float dtypes as float64 (by bit pattern).
    """

    __slots__ = ("_h", "_obj", "dtype", "__weakref__")

    def __init__(self, capacity=20, dtype="f8"):
        self._h = None
        self._obj = None
        capacity = _as_int(capacity, "capacity")
        if capacity <= 0:  # space_saving.pyx:180-181
            raise ValueError("capacity must be > 0")
        dtype = np.dtype(dtype)
        if dtype.kind == "O":  # :185-186: host-side Stream-Summary (Python hashing / equality)
            self.dtype = dtype
            self._obj = ObjectSummary(capacity)
            return
        elif dtype.kind in "iu":
            dtype = np.dtype(np.int64)
        elif dtype.kind == "f":
            dtype = np.dtype(np.float64)
        else:
            raise ValueError("dtype %s not supported" % dtype)
        self.dtype = dtype
        self._h = check_handle(lib.crick_spsv_new(capacity), "SpaceSaving()")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.crick_spsv_free(h)

    def __repr__(self):  # space_saving.pyx:207-209
        return "SpaceSaving<capacity={0}, dtype={1}, size={2}>".format(self.capacity, self.dtype,
                                                                       self.size())

    @property
    def capacity(self):
        if self._obj is not None:
            return self._obj.capacity
        return int(lib.crick_spsv_capacity(self._h))

    def size(self):
        """The number of active counters."""
        if self._obj is not None:
            return self._obj.size
        return int(check(lib.crick_spsv_size(self._h), "size"))

    def counters(self):
        """A copy of all the counters in the summary, most frequent first."""
        return self.topk(self.size())

    def __reduce__(self):  # space_saving.pyx:229-230
        return (SpaceSaving, (self.capacity, self.dtype), self.__getstate__())

    def __getstate__(self):
        return (self.counters(),)

    def __setstate__(self, state):  # space_saving.pyx:235-245
        if self._obj is not None:
            c = state[0]
            self._obj.set_state(c["item"].tolist(), c["count"].tolist(), c["error"].tolist())
            return
        counters = np.ascontiguousarray(state[0])
        if self.dtype.kind == "f":
            counters = counters.view(COUNTER_INT64_DTYPE)
        counters = np.ascontiguousarray(counters, dtype=COUNTER_INT64_DTYPE)
        rc = lib.crick_spsv_set_state(self._h, counters.ctypes.data, len(counters))
        if rc < 0:
            from ._lib import last_error
            raise ValueError(last_error())

    def _item_bits(self, item):
        """Cython's conversion of `item` to the C argument of spsv_int64_add
        (space_saving.pyx:265-268): ``asint64(<double>item)`` for float summaries, else
        ``<npy_int64>item`` -- through ``__int__``, so NaN is a ValueError and +-inf an
        OverflowError, and anything that is not a number a TypeError."""
        if isinstance(item, (str, bytes, bytearray)) or item is None:
            raise TypeError("item must be a number, not %s" % type(item).__name__)
        if self.dtype.kind == "f":
            try:
                v = float(item)
            except (TypeError, ValueError):
                raise TypeError("item must be a number, not %s" % type(item).__name__)
            return int(np.float64(v).view(np.int64))
        try:
            v = int(item)
        except TypeError:
            raise TypeError("item must be a number, not %s" % type(item).__name__)
        if not -(1 << 63) <= v < (1 << 63):
            raise OverflowError("Python int too large to convert to C long")
        return v

    def add(self, item, count=1):
        """add(self, item, count=1)

        Add an element to the summary.
        """  # space_saving.pyx:247-268
        count = _as_int(count, "count")
        if count <= 0:
            raise ValueError("count must be > 0")
        if self._obj is not None:
            self._obj.add(item, count)
            return
        check(lib.crick_spsv_add(self._h, self._item_bits(item), count), "add")

    def update(self, item, count=1, stream=None, mode=0, stats=None):
        """update(self, item, count=1)

        Add many elements to the summary.  `item` may be NumPy-compatible or an
        int64 / float64 CUDA array matching the summary's dtype.  ``mode`` (an addition) picks
        the kernel: 0 auto, 1 in-order replay, 2 shard summaries, 3 candidate counting
        (see ``crick_spsv_update`` in include/crick_cuda.h).  ``stats`` (an addition): a
        :class:`SummaryStats` that receives ``stats.update(item)`` in the same pass over the
        items (count 1 only; ``crick_spsv_update_stats``).
        """  # space_saving.pyx:270-302
        mode = int(mode)
        if self._obj is not None:
            if stats is not None:
                raise TypeError("stats= needs a numeric summary dtype")
            return self._update_object(item, count)
        if stats is not None:
            return self._update_with_stats(item, count, stream, stats)
        idv = device_view(item, stream)
        if idv is not None:
            if idv.dtype != self.dtype:
                raise TypeError("CUDA item array must have dtype %s" % self.dtype)
            if idv.size == 0:
                return
            cd = device_view(count, stream)
            if cd is not None:
                if cd.dtype != np.int64 or cd.size != idv.size:
                    raise TypeError("count must be an int64 CUDA array shaped like item")
                check(lib.crick_spsv_update(self._h, idv.ptr, cd.ptr, 0, idv.size, 1, 1, mode, stream),
                      "update")
            else:
                c = _as_int(count, "count")
                if c <= 0:
                    raise ValueError("count must be > 0")
                check(lib.crick_spsv_update(self._h, idv.ptr, None, c, idv.size, 1, 0, mode, stream),
                      "update")
            release_views((idv, cd), stream)
            return
        item = np.asarray(item)
        check_count = not (np.isscalar(count) and count == 1)
        count = np.asarray(count)
        item = item.astype(self.dtype, casting="safe", copy=False)
        count = count.astype(np.int64, casting="safe", copy=False)
        if check_count and (count <= 0).any():
            raise ValueError("count must be > 0")
        if item.size == 0 or count.size == 0:
            return
        if self.dtype.kind == "f":
            item = item.view("i8")
        xf, cf = korder_pair(item, count, w_dtype="i8", x_dtype="i8")
        if isinstance(cf, np.ndarray):
            check(lib.crick_spsv_update(self._h, xf.ctypes.data, cf.ctypes.data, 0, xf.size, 0, 0, mode,
                                        stream), "update")
        else:
            check(lib.crick_spsv_update(self._h, xf.ctypes.data, None, int(cf), xf.size, 0, 0, mode,
                                        stream), "update")

    def _update_with_stats(self, item, count, stream, stats):
        from .stats import SummaryStats
        if not isinstance(stats, SummaryStats):
            raise TypeError("stats must be a SummaryStats")
        if not (np.isscalar(count) and count == 1):
            raise ValueError("stats= needs count == 1")
        f64 = int(self.dtype.kind == "f")
        idv = device_view(item, stream)
        if idv is not None:
            if idv.dtype != self.dtype:
                raise TypeError("CUDA item array must have dtype %s" % self.dtype)
            if idv.size:
                check(lib.crick_spsv_update_stats(self._h, stats._h, idv.ptr, idv.size, 1, f64, stream),
                      "update")
                release_views((idv,), stream)
            return
        item = np.asarray(item).astype(self.dtype, casting="safe", copy=False)
        if item.size == 0:
            return
        if f64:
            item = item.view("i8")
        xf, _ = korder_pair(item, np.int64(1), w_dtype="i8", x_dtype="i8")
        check(lib.crick_spsv_update_stats(self._h, stats._h, xf.ctypes.data, xf.size, 0, f64, stream),
              "update")

    def _update_object(self, item, count):
        """spsv_object_update_ndarray, space_saving_stubs.c.in:367-451: the same coercion,
        broadcasting and K-order walk as the numeric dtypes, one ``add`` per element."""
        item = np.asarray(item)
        check_count = not (np.isscalar(count) and count == 1)
        count = np.asarray(count)
        item = item.astype(self.dtype, casting="safe", copy=False)
        count = count.astype(np.int64, casting="safe", copy=False)
        if check_count and (count <= 0).any():
            raise ValueError("count must be > 0")
        if item.size == 0 or count.size == 0:
            return
        add = self._obj.add
        with np.nditer([item, count], flags=["refs_ok", "external_loop", "buffered"],
                       op_flags=[["readonly", "aligned"], ["readonly", "aligned"]], order="K",
                       casting="safe", op_dtypes=[object, np.int64]) as it:
            for xs, cs in it:
                for x, c in zip(xs.tolist(), cs.tolist()):
                    add(x, c)
                del xs, cs

    def topk(self, k, astuples=False):
        """topk(self, k, astuples=False)

        Estimate the top k elements: a record array with fields item / count / error
        (``count - error <= actual <= count``), or a list of ``TopKResult`` tuples.
        """  # space_saving.pyx:304-341
        k = _as_int(k, "k")
        if k < 0:
            raise ValueError("k must be >= 0")
        k = min(self.size(), k)
        if self._obj is not None:
            out = np.empty(k, dtype=COUNTER_OBJECT_DTYPE)
            for i, rec in enumerate(self._obj.counters(k)):
                out[i] = rec
        else:
            out = np.empty(k, dtype=COUNTER_INT64_DTYPE)
            if k:
                check(lib.crick_spsv_counters(self._h, out.ctypes.data, k), "topk")
            if self.dtype.kind == "f":
                out = out.view(COUNTER_FLOAT64_DTYPE)
        if astuples:
            return [TopKResult(*i) for i in out]  # NumPy scalars, like space_saving.pyx:339-340
        return out

    def merge(self, *args):
        """merge(self, *args)

        Update this summary in-place with data from other summaries.
        """  # space_saving.pyx:343-365
        if not all(isinstance(i, SpaceSaving) for i in args):
            raise TypeError("All arguments to merge must be SpaceSavings")
        if not all(i.dtype == self.dtype for i in args):
            raise ValueError("All arguments to merge must have same dtype")
        if not args:
            return
        if self._obj is not None:
            for o in args:
                self._obj.merge(o._obj)
            return
        arr = handle_array([s._h for s in args])
        check(lib.crick_spsv_merge(self._h, arr, len(args)), "merge")

    # ---- multi-GPU tail (additions; crick_b200/distributed.py) ----
    def record_int64s(self):
        """Length of the fixed-size int64 record ``[size, 0, (item, count, error) x capacity]``."""
        return int(lib.crick_spsv_record_int64s(self._h))

    def export_record(self, device_ptr, stream=None):
        check(lib.crick_spsv_export_record(self._h, device_ptr, stream), "export_record")

    def merge_records(self, device_ptr, n_records, stream=None, bulk=False):
        """Merge `n_records` records (same capacity, device memory) one after the other, each
        exactly like ``self.merge(other)``.  ``bulk=True``: the k-way form of the same merge in
        one kernel -- same guarantees, not the list order nested merges would leave."""
        fn = lib.crick_spsv_merge_records_bulk if bulk else lib.crick_spsv_merge_records
        check(fn(self._h, device_ptr, int(n_records), stream), "merge_records")

    def clear(self, stream=None):
        """Empty the summary in place (the state ``SpaceSaving(capacity, dtype)`` starts with)."""
        if self._obj is not None:
            self._obj = ObjectSummary(self._obj.capacity)
            return
        check(lib.crick_spsv_reset(self._h, stream), "clear")


SpaceSaving.__module__ = "crick.space_saving"
TopKResult.__module__ = "crick.space_saving"
