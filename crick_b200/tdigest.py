"""TDigest -- drop-in for ``crick.TDigest`` (reference: crick/tdigest.pyx) running on the GPU.

Every method keeps the reference's name, argument meaning, validation order and error
types (file:line cited per method); the numeric work happens in libcrick_cuda.so.
Additions: ``mode=`` (constructor / ``update``), CUDA-array inputs, and the grouped entry
point :class:`TDigestGroup` / :func:`grouped_update` for many independent digests.
"""
import ctypes
import math
import numbers

import numpy as np

from . import _lib
from ._arrays import (DeviceArray, as_f64_device, device_view, dlpack_capsule, korder_pair,
                      release_views)
from ._lib import check, check_handle, lib

CENTROID_DTYPE = np.dtype([("mean", np.float64), ("weight", np.float64)])  # tdigest.pyx:43


def _as_double(v, name):
    """Cython's `double` argument conversion: numbers only, else TypeError."""
    if isinstance(v, (str, bytes)) or v is None:
        raise TypeError("%s must be a real number, not %s" % (name, type(v).__name__))
    try:
        return float(v)
    except (TypeError, ValueError):
        raise TypeError("%s must be a real number, not %s" % (name, type(v).__name__))


class TDigest:
    """TDigest(compression=100.0, mode="auto")

    An approximate histogram (merging t-digest), computed on the GPU.

    Parameters
    ----------
    compression : float, optional
        The compression factor to use (clipped to [20, 1000] like the reference,
        tdigest_stubs.c:57-62).
    mode : {"auto", "reference", "parallel"}, optional
        ``reference``: the reference's sequential buffer/flush algorithm, bit-exact
        with crick's C path.  ``parallel``: k-bucket pre-aggregation for huge arrays
        (needs >= 2048 values per call).  ``auto`` (default): reference below
        2048 values per ``update`` call (where one warp replaying the reference's
        sequential algorithm is still affordable), parallel at or above.
    """

    __slots__ = ("_h", "_mode", "_keepalive", "_deferred", "__weakref__")

    def __init__(self, compression=100.0, mode="auto"):
        self._h = None
        self._deferred = False
        c = _as_double(compression, "compression")
        if not math.isfinite(c):  # tdigest.pyx:81-82
            raise ValueError("Compression must be finite")
        self._mode = _lib.mode_code(mode)
        self._h = check_handle(lib.crick_tdigest_new(c), "TDigest()")  # :83-85

    def __del__(self):  # tdigest.pyx:87-89
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.crick_tdigest_free(h)

    def __repr__(self):  # tdigest.pyx:91-93
        return "TDigest<compression={0}, size={1}>".format(self.compression, self.size())

    @property
    def compression(self):
        """The compression factor for this digest"""
        return lib.crick_tdigest_compression(self._h)

    @property
    def mode(self):
        return {v: k for k, v in _lib._MODES.items()}[self._mode]

    def _flush(self):
        check(lib.crick_tdigest_flush(self._h), "flush")
        if self._deferred:
            self.check_weights()

    def check_weights(self):
        """Verdict of the deferred weight checks (``update(..., check_w="deferred")``) since the
        last call: raises the ValueError the synchronous check would have raised at the call
        (an update whose weights failed left the digest as it was).  Synchronises."""
        self._deferred = False
        bits = check(lib.crick_tdigest_take_bad_weights(self._h), "check_weights")
        if bits & 2:
            raise RuntimeError("a sharded update timed out waiting for another rank's partial digest")
        if bits & 1:
            raise ValueError("w must be > 0 and finite")

    def min(self):
        """min(self)

        The minimum value in the digest."""  # tdigest.pyx:100-107
        self._flush()
        if lib.crick_tdigest_total_weight(self._h) > 0:
            return lib.crick_tdigest_min(self._h)
        return np.nan

    def max(self):
        """max(self)

        The maximum value in the digest."""  # tdigest.pyx:109-116
        self._flush()
        if lib.crick_tdigest_total_weight(self._h) > 0:
            return lib.crick_tdigest_max(self._h)
        return np.nan

    def size(self):
        """size(self)

        The sum of the weights on all centroids."""  # tdigest.pyx:118-122 (no flush)
        return (lib.crick_tdigest_total_weight(self._h)
                + lib.crick_tdigest_buffer_total_weight(self._h))

    def _query(self, x, fn, name, out=None):
        dv = device_view(x)
        if dv is not None:
            dv = as_f64_device(dv)
            if out is None:
                out = DeviceArray(dv.shape, "f8")
                optr = out.ptr
            else:  # caller-provided CUDA output (no allocation on the query path)
                ov = device_view(out)
                if ov is None or ov.dtype != np.float64 or ov.size != dv.size:
                    raise TypeError("out must be a float64 CUDA array with as many elements as %s" % name)
                optr = ov.ptr
            check(fn(self._h, dv.ptr, optr, dv.size, 1, None), name)
            # the kernel ran on this digest's own (non-blocking) stream: make the result
            # visible to whatever stream the caller reads it from
            check(lib.crick_device_synchronize(), name)
            return out
        if out is not None:
            raise TypeError("out= is only supported with CUDA array inputs")
        x = np.asarray(x)
        if not np.can_cast(x.dtype, "f8", casting="safe"):  # tdigest.pyx:135,156
            raise TypeError("%s must be numeric" % name)
        if not np.isfinite(x).all():  # :137,158
            raise ValueError("%s must be finite" % name)
        xin = np.asarray(x, dtype=np.float64)
        if not xin.flags.c_contiguous:  # (ascontiguousarray would turn a 0-d input into 1-d)
            xin = np.ascontiguousarray(xin)
        out = np.empty(xin.shape, dtype=np.float64)
        check(fn(self._h, xin.ctypes.data, out.ctypes.data, xin.size, 0, None), name)
        if self._deferred:
            self.check_weights()
        if out.ndim == 0:  # :140-142
            return np.float64(out)
        return out

    def cdf(self, x, out=None):
        """cdf(self, x)

        Compute an estimate of the fraction of all points added to this digest
        which are <= x.

        Parameters
        ----------
        x : array_like or float (NumPy-compatible or CUDA array)
        """  # tdigest.pyx:124-142
        return self._query(x, lib.crick_tdigest_cdf, "x", out)

    def quantile(self, q, out=None):
        """quantile(self, q)

        Compute an estimate of the qth percentile of the data added to
        this digest.

        Parameters
        ----------
        q : array_like or float
            A number between 0 and 1 inclusive.
        """  # tdigest.pyx:144-163
        return self._query(q, lib.crick_tdigest_quantile, "q", out)

    def histogram(self, bins=10, range=None):
        """histogram(self, bins=10, range=None)

        Compute a histogram from the digest.

        Parameters
        ----------
        bins : int or array_like, optional
            Number of equal-width bins in the given range, or the bin edges
            (rightmost edge inclusive).  Default 10.
        range : (float, float), optional
            Lower and upper bounds used when generating bins; defaults to
            ``(t.min(), t.max())``.  Ignored when edges are given.

        Returns
        -------
        hist : array
        bin_edges : array, ``len(bin_edges) == len(hist) + 1``
        """  # tdigest.pyx:165-229
        left = right = 0.0
        size = self.size()
        if range is None:
            if size != 0:
                left = self.min()
                right = self.max()
        else:
            left = _as_double(range[0], "range[0]")
            right = _as_double(range[1], "range[1]")
            if not (math.isfinite(left) and math.isfinite(right)):
                raise ValueError("range parameters must be finite")
            elif right < left:
                raise ValueError("max must be larger than min for range parameter")
        if right == left:
            left -= 0.5
            right += 0.5

        if isinstance(bins, int):  # (like the reference: a NumPy integer is an array_like here)
            if bins < 1:
                raise ValueError("bins must be >= 1")
            bin_edges = np.linspace(left, right, bins + 1, endpoint=True)
        else:
            bin_edges = np.asarray(bins).astype("f8", copy=False)
            if bin_edges.ndim != 1:
                raise ValueError("bins must be a 1-dimensional array")
            elif not np.isfinite(bin_edges).all():
                raise ValueError("bins must be finite")
            elif (np.diff(bin_edges) < 0).any():
                raise ValueError("bins must increase monotonically")

        hist_size = bin_edges.size - 1
        hist = np.zeros(hist_size, dtype=np.float64)  # no edges at all: ValueError, as in :224-226
        if size != 0 and hist_size > 0:
            cdf = self.cdf(bin_edges)
            hist[:] = (cdf[1:] - cdf[:-1]) * size  # _cdf_to_hist, :46-50
        return hist, bin_edges

    def centroids(self):
        """centroids(self)

        Returns a numpy array of all the centroids in the digest. Note that
        this array is a *copy* of the internal data.
        """  # tdigest.pyx:231-244
        self._flush()
        n = check(lib.crick_tdigest_ncentroids(self._h), "ncentroids")
        result = np.empty(n, dtype=CENTROID_DTYPE)
        if n > 0:
            check(lib.crick_tdigest_get_centroids(self._h, result.ctypes.data, 0), "centroids")
        return result

    def centroids_device(self):
        """Zero-copy view of the live centroids as a CUDA array ``(n, 2)`` of
        (mean, weight); valid until the digest is next modified."""
        self._flush()
        n = check(lib.crick_tdigest_ncentroids(self._h), "ncentroids")
        ptr = lib.crick_tdigest_centroids_device(self._h)
        owner = self

        device = lib.crick_get_device()

        class _View:
            # (exported writable: PyTorch rejects read-only CUDA arrays; treat it as read-only)
            __cuda_array_interface__ = {"shape": (n, 2), "typestr": "<f8", "data": (ptr, False),
                                        "version": 3, "strides": None}
            _keep = owner

            def __dlpack_device__(self):
                return (2, device)       # kDLCUDA

            def __dlpack__(self, stream=None, **_):
                check(lib.crick_device_synchronize(), "synchronize")   # the flush above has run
                return dlpack_capsule(ptr, (n, 2), np.float64, device, owner)
        return _View()

    def __reduce__(self):  # tdigest.pyx:246-247
        return (TDigest, (self.compression,), self.__getstate__())

    def __getstate__(self):  # :249-251
        c = self.centroids()
        return (c, lib.crick_tdigest_total_weight(self._h), lib.crick_tdigest_min(self._h),
                lib.crick_tdigest_max(self._h))

    def __setstate__(self, state):  # :253-263
        c = np.ascontiguousarray(state[0], dtype=CENTROID_DTYPE)
        check(lib.crick_tdigest_set_state(self._h, c.ctypes.data, len(c), float(state[1]),
                                          float(state[2]), float(state[3])), "__setstate__")

    def add(self, x, w=1):
        """add(self, x, w=1)

        Add a sample to this digest.

        Parameters
        ----------
        x : float
            The value to add.
        w : float, optional
            The weight of the value to add. Default is 1.
        """  # tdigest.pyx:265-280
        x = _as_double(x, "x")
        w = _as_double(w, "w")
        if w != 1 and (not math.isfinite(w) or w <= 0):
            raise ValueError("w must be > 0 and finite")
        check(lib.crick_tdigest_add(self._h, x, w), "add")

    def update(self, x, w=1, mode=None, stream=None, check_w=True, stats=None, tail_stream=None):
        """update(self, x, w=1)

        Add many samples to this digest.

        Parameters
        ----------
        x : array_like
            The values to add (NumPy-compatible, or a float64 CUDA array exposing
            ``__cuda_array_interface__`` / DLPack).
        w : array_like, optional
            The weight (or weights) of the values to add. Default is 1.
        mode : optional override of the digest's update mode for this call.
        stream : optional ``cudaStream_t`` (int) to enqueue the work on.  Without it the work
            runs on the digest's own non-blocking stream: CUDA array inputs must then be
            complete (synchronise their producer, or pass the producing stream here).
        stats : optional :class:`SummaryStats`; its moments of ``x`` (``stats.update(x)``) are
            accumulated in the same pass over the data as the digest (parallel path).
        check_w : CUDA-array weights only.  True (default): raise ValueError like the
            reference when a weight is not finite or not > 0 -- checked on the device, which
            synchronises the stream.  False: fully asynchronous, ``w <= eps`` is dropped by
            the kernel exactly like ``tdigest_add`` does.  ``"deferred"`` (parallel path): the
            same check, but nothing waits for it -- a failed update leaves the digest as it
            was and the ValueError is raised by the next call that synchronises anyway
            (``centroids``, ``min``, ``max``, ``quantile`` / ``cdf`` of host arrays, pickling)
            or by ``check_weights()``: update -> export -> all-gather -> merge become one
            enqueue chain.
        tail_stream : a second ``cudaStream_t`` (CUDA-array inputs of at least PARALLEL_MIN values,
            with ``stream``): the pipelined form for back-to-back updates.  The kernels that only
            read ``x`` / ``w`` (sample, edges, the streaming pass) run on ``stream``, the short
            single-CTA tail that folds the result into the digest runs on ``tail_stream`` -- so
            does whatever the caller enqueues there afterwards (export, all-gather, merge) --
            and the next update's sample phase overlaps it.  Weights are checked like
            ``check_w="deferred"``.  Same results as the plain call.
        """  # tdigest.pyx:282-308
        m = self._mode if mode is None else _lib.mode_code(mode)
        if stats is not None:
            return self._update_with_stats(x, w, stream, check_w, stats)
        xd = device_view(x, stream)
        if tail_stream is not None:
            if xd is None or not stream:
                raise TypeError("tail_stream needs CUDA-array inputs and an explicit stream")
            return self._update_pipelined(xd, w, stream, tail_stream)
        if xd is not None:
            return self._update_device(xd, w, m, stream, check_w)
        x = np.asarray(x)
        check_w = not (np.isscalar(w) and w == 1)
        wd = device_view(w)
        if wd is not None:
            raise TypeError("w is a CUDA array but x is not")
        w = np.asarray(w)
        if not np.can_cast(x.dtype, "f8", casting="safe"):
            raise TypeError("x must be numeric")
        if not np.can_cast(w.dtype, "f8", casting="safe"):
            raise TypeError("w must be numeric")
        # Large weight arrays headed for the parallel path are validated on the device in the
        # same pass over the data (crick_tdigest_update_checked); the host check alone costs
        # more than the whole update there.
        try:
            n_b = int(np.prod(np.broadcast_shapes(x.shape, w.shape)))
        except ValueError:  # not broadcastable: korder_pair raises the reference's error below
            n_b = 0
        eff = m
        if m == _lib.MODE_PARALLEL and n_b < _lib.PARALLEL_MIN:
            eff = _lib.MODE_REFERENCE  # too small to sample: same answer, exact algorithm
        goes_parallel = eff == _lib.MODE_PARALLEL or (eff == _lib.MODE_AUTO
                                                      and n_b >= _lib.PARALLEL_MIN)
        device_check = check_w and w.ndim > 0 and goes_parallel
        if check_w and not device_check and (not np.isfinite(w).all() or (w <= 0).any()):
            raise ValueError("w must be > 0 and finite")
        if x.size == 0 or w.size == 0:  # tdigest_stubs.c:648-650
            return
        xf, wf = korder_pair(x, w)
        n = xf.size
        if isinstance(wf, np.ndarray):
            if device_check:
                bad = ctypes.c_int(0)
                check(lib.crick_tdigest_update_checked(self._h, xf.ctypes.data, wf.ctypes.data, 0.0,
                                                       n, 0, 0, eff, stream, ctypes.addressof(bad)),
                      "update")
                if bad.value:
                    raise ValueError("w must be > 0 and finite")
                return
            check(lib.crick_tdigest_update(self._h, xf.ctypes.data, wf.ctypes.data, 0.0, n, 0, 0,
                                           eff, stream), "update")
        else:
            check(lib.crick_tdigest_update(self._h, xf.ctypes.data, None, float(wf), n, 0, 0,
                                           eff, stream), "update")

    def _update_with_stats(self, x, w, stream, check_w, stats):
        """One pass over x for both sketches (crick_tdigest_update_stats); inputs too small for
        the parallel path take the two ordinary calls."""
        from .stats import SummaryStats
        if not isinstance(stats, SummaryStats):
            raise TypeError("stats must be a SummaryStats")
        xd = device_view(x, stream)
        wd = device_view(w, stream)
        bad = ctypes.c_int(0)
        if xd is not None:
            xd = as_f64_device(xd, stream)
            if xd.size < _lib.PARALLEL_MIN:
                self._update_device(xd, w, _lib.MODE_REFERENCE, stream, check_w)
                return stats.update(x)
            if wd is not None:
                wd = as_f64_device(wd, stream)
                if wd.size != xd.size:
                    raise ValueError("x and w CUDA arrays must have the same number of elements")
                wp, wsc = wd.ptr, 0.0
            else:
                wp, wsc = None, _as_double(w, "w")
                if wsc != 1 and (not math.isfinite(wsc) or wsc <= 0):
                    raise ValueError("w must be > 0 and finite")
            check(lib.crick_tdigest_update_stats(self._h, stats._h, xd.ptr, wp, wsc, xd.size, 1,
                                                 1 if wp else 0, stream,
                                                 ctypes.addressof(bad) if check_w else None),
                  "update")
            release_views((xd, wd), stream)
        else:
            x = np.asarray(x)
            w = np.asarray(w)
            if not np.can_cast(x.dtype, "f8", casting="safe"):
                raise TypeError("x must be numeric")
            if not np.can_cast(w.dtype, "f8", casting="safe"):
                raise TypeError("w must be numeric")
            if x.size == 0 or w.size == 0:
                return
            xf, wf = korder_pair(x, w)
            if xf.size < _lib.PARALLEL_MIN:
                self.update(x, w, mode="reference")
                return stats.update(x)
            if isinstance(wf, np.ndarray):
                wp, wsc = wf.ctypes.data, 0.0
            else:
                wp, wsc = None, float(wf)
                if wsc != 1 and (not math.isfinite(wsc) or wsc <= 0):
                    raise ValueError("w must be > 0 and finite")
            check(lib.crick_tdigest_update_stats(self._h, stats._h, xf.ctypes.data, wp, wsc, xf.size,
                                                 0, 0, stream, ctypes.addressof(bad)), "update")
        if bad.value:
            raise ValueError("w must be > 0 and finite")

    def _update_pipelined(self, xd, w, stream, tail_stream):
        xd = as_f64_device(xd, stream)
        n = xd.size
        if n < _lib.PARALLEL_MIN:
            raise ValueError("tail_stream needs at least %d values per call" % _lib.PARALLEL_MIN)
        wd = device_view(w, stream)
        if wd is not None:
            wd = as_f64_device(wd, stream)
            if wd.size != n:
                raise ValueError("x and w CUDA arrays must have the same number of elements")
            check(lib.crick_tdigest_update_pipelined(self._h, xd.ptr, wd.ptr, 0.0, n, stream, tail_stream),
                  "update")
            self._deferred = True
            release_views((xd, wd), stream)
            return
        if not np.isscalar(w) and np.ndim(w) != 0:
            raise TypeError("with a CUDA array x, w must be a scalar or a CUDA array")
        wv = _as_double(w, "w")
        if wv != 1 and (not math.isfinite(wv) or wv <= 0):
            raise ValueError("w must be > 0 and finite")
        check(lib.crick_tdigest_update_pipelined(self._h, xd.ptr, None, wv, n, stream, tail_stream), "update")
        release_views((xd,), stream)

    def _update_device(self, xd, w, m, stream, check_w=True):
        xd = as_f64_device(xd, stream)     # int / float16 / float32 ... CUDA arrays: safe cast on the device
        n = xd.size
        if n == 0:
            return
        eff = m
        if m == _lib.MODE_PARALLEL and n < _lib.PARALLEL_MIN:
            eff = _lib.MODE_REFERENCE
        wd = device_view(w, stream)
        if wd is not None:
            wd = as_f64_device(wd, stream)
            if wd.size != n:
                raise ValueError("x and w CUDA arrays must have the same number of elements")
            parallel = eff == _lib.MODE_PARALLEL or (eff == _lib.MODE_AUTO
                                                     and n >= _lib.PARALLEL_MIN)
            if isinstance(check_w, str):
                if check_w != "deferred":
                    raise ValueError("check_w must be True, False or 'deferred'")
                if parallel:
                    check(lib.crick_tdigest_update_deferred(self._h, xd.ptr, wd.ptr, 0.0, n, eff,
                                                            stream), "update")
                    self._deferred = True
                    release_views((xd, wd), stream)
                    return
            if check_w and parallel:
                bad = ctypes.c_int(0)
                check(lib.crick_tdigest_update_checked(self._h, xd.ptr, wd.ptr, 0.0, n, 1, 1, eff,
                                                       stream, ctypes.addressof(bad)), "update")
                if bad.value:
                    raise ValueError("w must be > 0 and finite")
                return
            # reference-mode sizes: the wrapper's check (tdigest.pyx:304-305) as a small reduction
            # over the device weights before anything is added
            if check_w and check(lib.crick_check_weights(wd.ptr, n, stream), "check weights"):
                raise ValueError("w must be > 0 and finite")
            check(lib.crick_tdigest_update(self._h, xd.ptr, wd.ptr, 0.0, n, 1, 1, eff, stream),
                  "update")
            release_views((xd, wd), stream)
            return
        if not np.isscalar(w) and np.ndim(w) != 0:
            raise TypeError("with a CUDA array x, w must be a scalar or a CUDA array")
        wv = _as_double(w, "w")
        if wv != 1 and (not math.isfinite(wv) or wv <= 0):
            raise ValueError("w must be > 0 and finite")
        check(lib.crick_tdigest_update(self._h, xd.ptr, None, wv, n, 1, 0, eff, stream), "update")
        release_views((xd,), stream)

    def merge(self, *args):
        """merge(self, *args)

        Update this digest inplace with data from other digests.

        Parameters
        ----------
        args : TDigests
            TDigests to merge into this one.
        """  # tdigest.pyx:310-324
        if not all(isinstance(i, TDigest) for i in args):
            raise TypeError("All arguments to merge must be TDigests")
        if not args:
            return
        arr = _lib.handle_array([t._h for t in args])
        check(lib.crick_tdigest_merge(self._h, arr, len(args)), "merge")

    def scale(self, factor):
        """scale(self, factor)

        Return a new TDigest, with the weights all scaled by factor.

        Parameters
        ----------
        factor : number
            The factor to scale the weights by. Must be > 0.
        """  # tdigest.pyx:326-340
        factor = _as_double(factor, "factor")
        if factor <= 0 or not math.isfinite(factor):
            raise ValueError("factor must be > 0 and finite")
        out = TDigest(self.compression, mode=self.mode)  # copy(self) via __reduce__
        out.__setstate__(self.__getstate__())
        check(lib.crick_tdigest_scale(out._h, factor), "scale")
        return out

    def clear(self, stream=None):
        """Empty the digest in place (the state ``TDigest(compression)`` starts with)."""
        check(lib.crick_tdigest_reset(self._h, stream), "clear")

    # ---- multi-GPU wire format ------------------------------------
    def record_doubles(self):
        return lib.crick_tdigest_record_doubles(self._h)

    def export_record(self, device_ptr, stream=None):
        """Flush and write the fixed-size record (4 + 2*size doubles) at `device_ptr`."""
        check(lib.crick_tdigest_export_record(self._h, device_ptr, stream), "export_record")

    def merge_records(self, device_ptr, n_records, skip=-1, stream=None, bulk=False):
        """``self.merge(*records)`` in record order for records gathered on the device.
        ``bulk=True``: one greedy pass over the sorted union of all centroids instead of the
        reference's incremental merge (faster, not bit-identical with ``merge``)."""
        if bulk:
            check(lib.crick_tdigest_merge_records_bulk(self._h, device_ptr, n_records, stream),
                  "merge_records")
        else:
            check(lib.crick_tdigest_merge_records(self._h, device_ptr, n_records, skip, stream),
                  "merge_records")

    def _debug_edges(self):
        buf = np.empty(2048, dtype=np.float64)
        n = check(lib.crick_tdigest_debug_edges(self._h, buf.ctypes.data, buf.size), "debug_edges")
        return buf[:n].copy()

    def _debug_buckets(self):
        w = np.empty(2048, dtype=np.float64)
        s = np.empty(2048, dtype=np.float64)
        mm = np.empty(2, dtype=np.float64)
        n = check(lib.crick_tdigest_debug_buckets(self._h, w.ctypes.data, s.ctypes.data,
                                                  mm.ctypes.data, 2048), "debug_buckets")
        return w[:n].copy(), s[:n].copy(), mm


# Pickles name `crick.tdigest.TDigest` (tdigest.pyx:246-247) so digests move between the
# reference build and this one; the `crick/` shim package re-exports these classes.
TDigest.__module__ = "crick.tdigest"


class TDigestGroup:
    """Many independent digests with one compression, one warp per digest.

    Per group the semantics are exactly ``t = TDigest(c); t.update(values_g, weights_g)``
    in reference-compatible mode (bit-exact with crick's C path): the dask
    groupby-quantile shape in one call.
    """

    def __init__(self, n_groups, compression=100.0):
        self._h = None
        c = _as_double(compression, "compression")
        if not math.isfinite(c):
            raise ValueError("Compression must be finite")
        if int(n_groups) <= 0:
            raise ValueError("n_groups must be > 0")
        self._h = check_handle(lib.crick_tdigest_group_new(c, int(n_groups)), "TDigestGroup()")
        self._compression = min(max(c, 20.0), 1000.0)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.crick_tdigest_group_free(h)

    def __len__(self):
        return int(lib.crick_tdigest_group_count(self._h))

    @property
    def compression(self):
        return self._compression

    def update(self, values, weights=1, offsets=None, stream=None):
        """Add values to every group.

        values : 2-D ``(n_groups, row_len)`` float64 array (host or CUDA), or a 1-D array
            with ``offsets`` (``n_groups + 1`` int64, CSR style) for ragged groups.
        weights : scalar, or an array shaped like ``values``.
        """
        ng = len(self)
        vd = device_view(values, stream)
        if vd is not None:
            vd = as_f64_device(vd, stream)
            wptr, wsc = None, 1.0
            wd = device_view(weights, stream)
            if wd is not None:
                wd = as_f64_device(wd, stream)
                if wd.size != vd.size:
                    raise TypeError("weights must be a CUDA array shaped like values")
                if check(lib.crick_check_weights(wd.ptr, wd.size, stream), "check weights"):
                    raise ValueError("w must be > 0 and finite")
                wptr = wd.ptr
            else:
                wsc = _as_double(weights, "weights")
                if wsc != 1 and (not math.isfinite(wsc) or wsc <= 0):
                    raise ValueError("w must be > 0 and finite")
            if offsets is None:
                if len(vd.shape) != 2 or vd.shape[0] != ng:
                    raise ValueError("values must have shape (n_groups, row_len)")
                check(lib.crick_tdigest_group_update(self._h, vd.ptr, wptr, wsc, None, vd.shape[1],
                                                     vd.shape[1], 1, stream), "group update")
                release_views((vd, wd), stream)
            else:
                od = device_view(offsets, stream)
                if od is None or od.dtype != np.int64 or od.size != ng + 1:
                    raise TypeError("offsets must be an int64 CUDA array of n_groups + 1 entries")
                check(lib.crick_tdigest_group_update(self._h, vd.ptr, wptr, wsc, od.ptr, 0, 0, 1,
                                                     stream), "group update")
                release_views((vd, wd, od), stream)
            return
        v = np.asarray(values)
        if not np.can_cast(v.dtype, "f8", casting="safe"):
            raise TypeError("x must be numeric")
        v = np.ascontiguousarray(v, dtype=np.float64)
        w = np.asarray(weights)
        if not np.can_cast(w.dtype, "f8", casting="safe"):
            raise TypeError("w must be numeric")
        if not (np.isscalar(weights) and weights == 1):
            if not np.isfinite(w).all() or (w <= 0).any():
                raise ValueError("w must be > 0 and finite")
        wptr, wsc, wkeep = None, 1.0, None
        if w.ndim == 0:
            wsc = float(w)
        else:
            wkeep = np.ascontiguousarray(np.broadcast_to(w, v.shape), dtype=np.float64)
            wptr = wkeep.ctypes.data
        if offsets is None:
            if v.ndim != 2 or v.shape[0] != ng:
                raise ValueError("values must have shape (n_groups, row_len)")
            if v.shape[1] == 0:
                return
            check(lib.crick_tdigest_group_update(self._h, v.ctypes.data, wptr, wsc, None, v.shape[1],
                                                 v.shape[1], 0, stream), "group update")
        else:
            off = np.ascontiguousarray(offsets, dtype=np.int64)
            if off.ndim != 1 or off.size != ng + 1 or off[0] != 0 or off[-1] != v.size \
                    or (np.diff(off) < 0).any():
                raise ValueError("offsets must be n_groups + 1 non-decreasing int64 values "
                                 "from 0 to values.size")
            if v.size == 0:
                return
            check(lib.crick_tdigest_group_update(self._h, v.ctypes.data, wptr, wsc, off.ctypes.data,
                                                 0, 0, 0, stream), "group update")

    def flush(self):
        check(lib.crick_tdigest_group_flush(self._h), "group flush")

    def _query(self, q, fn, name):
        dv = device_view(q)
        ng = len(self)
        if dv is not None:
            if dv.dtype != np.float64:
                raise TypeError("%s must be float64 when given as a CUDA array" % name)
            out = DeviceArray((ng, dv.size), "f8")
            check(fn(self._h, dv.ptr, dv.size, out.ptr, 1, None), name)
            check(lib.crick_device_synchronize(), name)   # see TDigest._query
            return out
        q = np.asarray(q)
        if not np.can_cast(q.dtype, "f8", casting="safe"):
            raise TypeError("%s must be numeric" % name)
        if not np.isfinite(q).all():
            raise ValueError("%s must be finite" % name)
        qin = np.ascontiguousarray(q, dtype=np.float64).reshape(-1)
        out = np.empty((ng, qin.size), dtype=np.float64)
        check(fn(self._h, qin.ctypes.data, qin.size, out.ctypes.data, 0, None), name)
        return out

    def quantile(self, q):
        """Per-group quantiles: array ``(n_groups, len(q))``."""
        return self._query(q, lib.crick_tdigest_group_quantile, "q")

    def cdf(self, x):
        """Per-group cdf: array ``(n_groups, len(x))``."""
        return self._query(x, lib.crick_tdigest_group_cdf, "x")

    def meta(self):
        """Structured array (ncentroids, total_weight, min, max) per group (no flush)."""
        out = np.empty((len(self), 4), dtype=np.float64)
        check(lib.crick_tdigest_group_meta(self._h, out.ctypes.data), "group meta")
        return out

    def centroids(self, i):
        """Centroids of group `i` (flushes that group), like ``TDigest.centroids``."""
        cap = 2 * int(math.ceil(self._compression))
        buf = np.empty(cap, dtype=CENTROID_DTYPE)
        n = check(lib.crick_tdigest_group_get_centroids(self._h, int(i), buf.ctypes.data),
                  "group centroids")
        return buf[:n].copy()

    def digest(self, i):
        """Copy group `i` out as a standalone :class:`TDigest`."""
        out = TDigest(self._compression)
        check(lib.crick_tdigest_group_extract(self._h, int(i), out._h), "group extract")
        return out

    def merge_all(self):
        """Merge all groups into one new TDigest in a single pass: every centroid of every group
        is a (mean, weight) point -- what ``TDigest.merge`` feeds through ``add`` -- and the union
        goes through the parallel k-bucket update at once.  Not bit-identical with nested
        ``merge`` calls (:meth:`merge_tree` is), same accuracy class, orders of magnitude
        shorter for many digests.  The group's digests are flushed but otherwise left alone."""
        return self._merge_all()

    def merge_tree(self):
        """Merge all groups pairwise, level by level (left operand = self, right = other,
        exactly a nest of ``TDigest.merge`` calls) and return the resulting TDigest.
        The group's digests are consumed (modified in place)."""
        out = TDigest(self._compression)
        check(lib.crick_tdigest_group_merge_tree(self._h, out._h), "merge_tree")
        return out


    def _merge_all(self):
        out = TDigest(self._compression)
        check(lib.crick_tdigest_group_merge_all(self._h, out._h), "merge_all")
        return out


def grouped_update(values, weights=1, offsets=None, compression=100.0):
    """Build one digest per row of `values` (or per ``offsets`` segment) in one call."""
    if offsets is not None:
        dv = device_view(offsets)
        ng = (dv.size if dv is not None else len(offsets)) - 1
    else:
        dv = device_view(values)
        ng = dv.shape[0] if dv is not None else np.shape(values)[0]
    g = TDigestGroup(ng, compression)
    g.update(values, weights, offsets)
    return g
