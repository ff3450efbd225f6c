"""SummaryStats -- drop-in for ``crick.SummaryStats`` (reference: crick/stats.pyx) on the GPU."""
import ctypes

import numpy as np

from ._arrays import device_view, korder_pair, release_views
from ._lib import check, check_handle, handle_array, lib


def _as_double(v, name):
    if isinstance(v, (str, bytes)) or v is None:
        raise TypeError("%s must be a real number, not %s" % (name, type(v).__name__))
    try:
        return float(v)
    except (TypeError, ValueError):
        raise TypeError("%s must be a real number, not %s" % (name, type(v).__name__))


def _as_int(v, name):
    if isinstance(v, (str, bytes, float)) or v is None:
        if isinstance(v, float) and v == int(v):
            return int(v)
        raise TypeError("%s must be an integer" % name)
    try:
        return int(v)
    except (TypeError, ValueError):
        raise TypeError("%s must be an integer" % name)


class SummaryStats:
    """SummaryStats()

    Computes exact summary statistics on a data stream: count, sum, minimum, maximum,
    mean, variance, standard deviation, skewness, kurtosis (Pebay's one-pass update
    formulas, as the reference; evaluated on the GPU).
    """

    __slots__ = ("_h", "__weakref__")

    def __init__(self):
        self._h = None
        self._h = check_handle(lib.crick_stats_new(), "SummaryStats()")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.crick_stats_free(h)

    def __repr__(self):  # stats.pyx:71-72
        return "SummaryStats<count=%d>" % self.count()

    def __reduce__(self):  # stats.pyx:74-75
        return (SummaryStats, (), self.__getstate__())

    def __getstate__(self):  # stats.pyx:77-80
        cnt = ctypes.c_int64(0)
        hom = ctypes.c_int(0)
        f7 = np.empty(7, dtype=np.float64)
        check(lib.crick_stats_getstate(self._h, ctypes.addressof(cnt), f7.ctypes.data,
                                       ctypes.addressof(hom)), "getstate")
        return (cnt.value, float(f7[0]), float(f7[1]), float(f7[2]), float(f7[3]), float(f7[4]),
                float(f7[5]), bool(hom.value), float(f7[6]))

    def __setstate__(self, state):  # stats.pyx:82-91
        f7 = np.array([state[1], state[2], state[3], state[4], state[5], state[6], state[8]],
                      dtype=np.float64)
        check(lib.crick_stats_setstate(self._h, int(state[0]), f7.ctypes.data, int(bool(state[7]))),
              "setstate")

    def add(self, x, count=1):
        """add(self, x, count=1)

        Add an element to the summary.
        """  # stats.pyx:93-106
        x = _as_double(x, "x")
        count = _as_int(count, "count")
        if count <= 0:
            raise ValueError("count must be > 0")
        check(lib.crick_stats_add(self._h, x, count), "add")

    def update(self, x, count=1, stream=None):
        """update(self, x, count=1)

        Add many elements to the summary.  `x` may be NumPy-compatible or a float64 / int64
        CUDA array; `count` a scalar or an array.
        """  # stats.pyx:108-126
        xd = device_view(x, stream)
        if xd is not None:
            if xd.dtype not in (np.dtype("f8"), np.dtype("i8")):
                from ._arrays import as_f64_device
                xd = as_f64_device(xd, stream)          # other numeric dtypes: safe cast on the device
            if xd.size == 0:
                return
            cd = device_view(count, stream)
            if cd is not None:
                if cd.dtype != np.int64 or cd.size != xd.size:
                    raise TypeError("count must be an int64 CUDA array shaped like x")
                check(lib.crick_stats_update(self._h, xd.ptr, int(xd.dtype == np.int64), cd.ptr, 0,
                                             xd.size, 1, 1, stream), "update")
                release_views((xd, cd), stream)
            else:
                c = _as_int(count, "count")
                if c <= 0:
                    raise ValueError("count must be > 0")
                check(lib.crick_stats_update(self._h, xd.ptr, int(xd.dtype == np.int64), None, c,
                                             xd.size, 1, 0, stream), "update")
                release_views((xd,), stream)
            return
        x = np.asarray(x)
        check_count = not (np.isscalar(count) and count == 1)
        count = np.asarray(count, dtype="i8")
        if check_count and (count <= 0).any():
            raise ValueError("count must be > 0")
        if x.size == 0 or count.size == 0:  # stats_stubs.c:155-157
            return
        # the iterator's safe cast to float64 (stats_stubs.c:165-177): TypeError otherwise
        xf, cf = korder_pair(x, count, w_dtype="i8")
        if isinstance(cf, np.ndarray):
            check(lib.crick_stats_update(self._h, xf.ctypes.data, 0, cf.ctypes.data, 0, xf.size, 0, 0,
                                         stream), "update")
        else:
            check(lib.crick_stats_update(self._h, xf.ctypes.data, 0, None, int(cf), xf.size, 0, 0,
                                         stream), "update")

    def merge(self, *args):
        """merge(self, *args)

        Update this summary inplace with data from other summaries.
        """  # stats.pyx:128-142
        if not all(isinstance(i, SummaryStats) for i in args):
            raise TypeError("All arguments to merge must be SummaryStats")
        if not args:
            return
        arr = handle_array([s._h for s in args])
        check(lib.crick_stats_merge(self._h, arr, len(args)), "merge")

    def count(self):
        """The number of elements in the summary."""
        return self.__getstate__()[0]

    def sum(self):
        """The sum of all elements in the summary."""
        return self.__getstate__()[1]

    def min(self):
        """The minimum element added to the summary. Returns ``NaN`` if empty."""
        s = self.__getstate__()
        return s[2] if s[0] > 0 else np.nan

    def max(self):
        """The maximum element added to the summary. Returns ``NaN`` if empty."""
        s = self.__getstate__()
        return s[3] if s[0] > 0 else np.nan

    def mean(self):
        """The mean of all elements in the summary. Returns ``NaN`` if empty."""
        return lib.crick_stats_mean(self._h)

    def var(self, ddof=0):
        """The variance (divisor ``N - ddof``). Returns ``NaN`` if empty."""
        return lib.crick_stats_var(self._h, _as_int(ddof, "ddof"))

    def std(self, ddof=0):
        """The standard deviation (divisor ``N - ddof``). Returns ``NaN`` if empty."""
        return lib.crick_stats_std(self._h, _as_int(ddof, "ddof"))

    def skew(self, bias=True):
        """The skewness; ``bias=False`` applies the statistical bias correction."""
        return lib.crick_stats_skew(self._h, int(bool(bias)))

    def kurt(self, fisher=True, bias=True):
        """The kurtosis (Fisher by default, Pearson with ``fisher=False``)."""
        return lib.crick_stats_kurt(self._h, int(bool(fisher)), int(bool(bias)))

    def clear(self, stream=None):
        """Empty the summary in place (the state ``SummaryStats()`` starts with)."""
        check(lib.crick_stats_reset(self._h, stream), "clear")

    def export_record_device(self, device_ptr, stream=None):
        """The 9-double record written to device memory, asynchronously on `stream`."""
        check(lib.crick_stats_export_record_dev(self._h, device_ptr, stream), "export_record")

    def merge_records_device(self, device_ptr, n_records, stride=9, stream=None):
        """``self.merge(*records)`` in order for 9-double records in device memory, `stride`
        doubles apart; asynchronous on `stream`."""
        check(lib.crick_stats_merge_records_dev(self._h, device_ptr, int(n_records), int(stride), stream),
              "merge_records")

    # ---- multi-GPU wire format: 9 doubles ----
    def export_record(self):
        r = np.empty(9, dtype=np.float64)
        check(lib.crick_stats_export_record(self._h, r.ctypes.data), "export_record")
        return r

    def merge_record(self, record):
        r = np.ascontiguousarray(record, dtype=np.float64)
        check(lib.crick_stats_merge_record(self._h, r.ctypes.data), "merge_record")


SummaryStats.__module__ = "crick.stats"
