"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU checkers for the CUDA path.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package; the
product (crick_b200/) never does.

Two implementations with one Python face:

* ``oracle.restated``  -- our C restatement (oracle/crick_oracle.c -> liboracle.so)
* ``oracle.reference`` -- the unmodified reference C sources behind
  oracle/ref_harness.c (oracle/_ref/libcrick_ref.so); ``None`` when that
  library has not been built (it needs /root/reference at build time).

Each namespace offers TDigest / SummaryStats / SpaceSaving objects with the
reference's method names (tdigest.pyx, stats.pyx, space_saving.pyx) on
contiguous float64 / int64 NumPy inputs -- no validation, no broadcasting:
that logic belongs to the product wrappers and is tested there.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

_c_dp = ctypes.c_void_p
_dbl = ctypes.c_double
_lng = ctypes.c_long
_ll = ctypes.c_longlong
_int = ctypes.c_int

CENTROID_DTYPE = np.dtype([("mean", np.float64), ("weight", np.float64)])
COUNTER_DTYPE = np.dtype([("item", np.int64), ("count", np.int64), ("error", np.int64)])


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _bind(lib, prefix):
    sig = {
        "tdigest_new": (_c_dp, [_dbl]),
        "tdigest_free": (None, [_c_dp]),
        "tdigest_add": (None, [_c_dp, _dbl, _dbl]),
        "tdigest_flush": (None, [_c_dp]),
        "tdigest_merge": (None, [_c_dp, _c_dp]),
        "tdigest_scale": (None, [_c_dp, _dbl]),
        "tdigest_update": (None, [_c_dp, _c_dp, _c_dp, _dbl, _lng]),
        "tdigest_quantile": (None, [_c_dp, _c_dp, _c_dp, _lng]),
        "tdigest_cdf": (None, [_c_dp, _c_dp, _c_dp, _lng]),
        "tdigest_compression": (_dbl, [_c_dp]),
        "tdigest_size": (_int, [_c_dp]),
        "tdigest_buffer_size": (_int, [_c_dp]),
        "tdigest_min": (_dbl, [_c_dp]),
        "tdigest_max": (_dbl, [_c_dp]),
        "tdigest_ncentroids": (_int, [_c_dp]),
        "tdigest_buffer_ncentroids": (_int, [_c_dp]),
        "tdigest_total_weight": (_dbl, [_c_dp]),
        "tdigest_buffer_total_weight": (_dbl, [_c_dp]),
        "tdigest_get_centroids": (None, [_c_dp, _c_dp]),
        "tdigest_set_state": (None, [_c_dp, _c_dp, _int, _dbl, _dbl, _dbl]),
        "tdigest_update_sharded": (_c_dp, [_dbl, _c_dp, _c_dp, _dbl, _lng, _int]),
        "stats_new": (_c_dp, []),
        "stats_free": (None, [_c_dp]),
        "stats_add": (None, [_c_dp, _dbl, _ll]),
        "stats_merge": (None, [_c_dp, _c_dp]),
        "stats_update": (None, [_c_dp, _c_dp, _c_dp, _ll, _lng]),
        "stats_getstate": (None, [_c_dp, _c_dp, _c_dp, _c_dp]),
        "stats_setstate": (None, [_c_dp, _ll, _c_dp, _int]),
        "stats_mean": (_dbl, [_c_dp]),
        "stats_var": (_dbl, [_c_dp, _lng]),
        "stats_std": (_dbl, [_c_dp, _lng]),
        "stats_skew": (_dbl, [_c_dp, _int]),
        "stats_kurt": (_dbl, [_c_dp, _int, _int]),
        "spsv_new": (_c_dp, [_int]),
        "spsv_free": (None, [_c_dp]),
        "spsv_add": (_int, [_c_dp, _ll, _ll]),
        "spsv_update": (None, [_c_dp, _c_dp, _c_dp, _ll, _lng]),
        "spsv_merge": (_int, [_c_dp, _c_dp]),
        "spsv_set_state": (_int, [_c_dp, _c_dp, _lng]),
        "spsv_size": (_lng, [_c_dp]),
        "spsv_capacity": (_lng, [_c_dp]),
        "spsv_counters": (_lng, [_c_dp, _c_dp, _lng]),
    }
    out = {}
    for name, (res, args) in sig.items():
        fn = getattr(lib, prefix + name)
        fn.restype = res
        fn.argtypes = args
        out[name] = fn
    return out


class _Namespace:
    """One implementation (restated or reference) behind the reference's API names."""

    def __init__(self, path, prefix, kind):
        self.path = path
        self.kind = kind
        self.lib = ctypes.CDLL(path)
        f = self.f = _bind(self.lib, prefix)
        ns = self

        class TDigest:
            def __init__(self, compression=100.0, _handle=None):
                self._h = _handle if _handle is not None else f["tdigest_new"](float(compression))

            def __del__(self):
                if getattr(self, "_h", None):
                    f["tdigest_free"](self._h)
                    self._h = None

            @property
            def compression(self):
                return f["tdigest_compression"](self._h)

            def add(self, x, w=1.0):
                f["tdigest_add"](self._h, float(x), float(w))

            def update(self, x, w=1.0):
                x = _f64(x).ravel()
                if np.ndim(w) == 0:
                    f["tdigest_update"](self._h, x.ctypes.data, None, float(w), x.size)
                else:
                    w = _f64(w).ravel()
                    assert w.size == x.size
                    f["tdigest_update"](self._h, x.ctypes.data, w.ctypes.data, 0.0, x.size)

            def flush(self):
                f["tdigest_flush"](self._h)

            def merge(self, *others):
                for o in others:
                    f["tdigest_merge"](self._h, o._h)

            def scale(self, factor):
                out = TDigest(self.compression)
                c, tw, mn, mx = self.__getstate__()
                out.__setstate__((c, tw, mn, mx))
                f["tdigest_scale"](out._h, float(factor))
                return out

            def centroids(self):
                self.flush()
                n = f["tdigest_ncentroids"](self._h)
                out = np.empty(n, dtype=CENTROID_DTYPE)
                if n:
                    f["tdigest_get_centroids"](self._h, out.ctypes.data)
                return out

            def min(self):
                self.flush()
                return f["tdigest_min"](self._h) if f["tdigest_total_weight"](self._h) > 0 else np.nan

            def max(self):
                self.flush()
                return f["tdigest_max"](self._h) if f["tdigest_total_weight"](self._h) > 0 else np.nan

            def size(self):
                return f["tdigest_total_weight"](self._h) + f["tdigest_buffer_total_weight"](self._h)

            def raw(self):
                """(ncentroids, total_weight, min, max, buffer_n, buffer_total) without flushing."""
                return (f["tdigest_ncentroids"](self._h), f["tdigest_total_weight"](self._h),
                        f["tdigest_min"](self._h), f["tdigest_max"](self._h),
                        f["tdigest_buffer_ncentroids"](self._h),
                        f["tdigest_buffer_total_weight"](self._h))

            def quantile(self, q):
                q = _f64(q)
                out = np.empty_like(q)
                f["tdigest_quantile"](self._h, q.ctypes.data, out.ctypes.data, q.size)
                return out

            def cdf(self, x):
                x = _f64(x)
                out = np.empty_like(x)
                f["tdigest_cdf"](self._h, x.ctypes.data, out.ctypes.data, x.size)
                return out

            def __getstate__(self):
                c = self.centroids()
                return (c, f["tdigest_total_weight"](self._h), f["tdigest_min"](self._h),
                        f["tdigest_max"](self._h))

            def __setstate__(self, state):
                c = np.ascontiguousarray(state[0], dtype=CENTROID_DTYPE)
                f["tdigest_set_state"](self._h, c.ctypes.data, len(c), float(state[1]),
                                       float(state[2]), float(state[3]))

            def absorb_sorted(self, items, vmin, vmax):
                """Parallel-mode finalize restated (restated namespace only)."""
                it = np.ascontiguousarray(items, dtype=np.float64).reshape(-1, 2)
                fn = ns.lib.orc_tdigest_absorb_sorted
                fn.restype = None
                fn.argtypes = [_c_dp, _c_dp, _int, _dbl, _dbl]
                fn(self._h, it.ctypes.data, it.shape[0], float(vmin), float(vmax))

            @staticmethod
            def update_sharded(compression, x, w=1.0, nthreads=1):
                """One digest per contiguous shard (one thread each), merged in shard order."""
                x = _f64(x).ravel()
                if np.ndim(w) == 0:
                    h = f["tdigest_update_sharded"](float(compression), x.ctypes.data, None,
                                                    float(w), x.size, int(nthreads))
                else:
                    w = _f64(w).ravel()
                    h = f["tdigest_update_sharded"](float(compression), x.ctypes.data,
                                                    w.ctypes.data, 0.0, x.size, int(nthreads))
                return TDigest(_handle=h)

        class SummaryStats:
            def __init__(self):
                self._h = f["stats_new"]()

            def __del__(self):
                if getattr(self, "_h", None):
                    f["stats_free"](self._h)
                    self._h = None

            def add(self, x, count=1):
                f["stats_add"](self._h, float(x), int(count))

            def update(self, x, count=1):
                x = _f64(x).ravel()
                if np.ndim(count) == 0:
                    f["stats_update"](self._h, x.ctypes.data, None, int(count), x.size)
                else:
                    c = _i64(count).ravel()
                    assert c.size == x.size
                    f["stats_update"](self._h, x.ctypes.data, c.ctypes.data, 0, x.size)

            def merge(self, *others):
                for o in others:
                    f["stats_merge"](self._h, o._h)

            def __getstate__(self):
                cnt = _ll(0)
                f7 = np.empty(7)
                hom = _int(0)
                f["stats_getstate"](self._h, ctypes.addressof(cnt), f7.ctypes.data,
                                    ctypes.addressof(hom))
                return (cnt.value, f7[0], f7[1], f7[2], f7[3], f7[4], f7[5], bool(hom.value), f7[6])

            def __setstate__(self, s):
                f7 = np.array([s[1], s[2], s[3], s[4], s[5], s[6], s[8]], dtype=np.float64)
                f["stats_setstate"](self._h, int(s[0]), f7.ctypes.data, int(bool(s[7])))

            def count(self):
                return self.__getstate__()[0]

            def sum(self):
                return self.__getstate__()[1]

            def min(self):
                s = self.__getstate__()
                return s[2] if s[0] > 0 else np.nan

            def max(self):
                s = self.__getstate__()
                return s[3] if s[0] > 0 else np.nan

            def mean(self):
                return f["stats_mean"](self._h)

            def var(self, ddof=0):
                return f["stats_var"](self._h, ddof)

            def std(self, ddof=0):
                return f["stats_std"](self._h, ddof)

            def skew(self, bias=True):
                return f["stats_skew"](self._h, int(bias))

            def kurt(self, fisher=True, bias=True):
                return f["stats_kurt"](self._h, int(fisher), int(bias))

        class SpaceSaving:
            """int64 items only (float64 items are the same code on the bit pattern,
            space_saving_stubs.c.in:459-466)."""

            def __init__(self, capacity=20):
                self._h = f["spsv_new"](int(capacity))

            def __del__(self):
                if getattr(self, "_h", None):
                    f["spsv_free"](self._h)
                    self._h = None

            def add(self, item, count=1):
                f["spsv_add"](self._h, int(item), int(count))

            def update(self, item, count=1):
                item = _i64(item).ravel()
                if np.ndim(count) == 0:
                    f["spsv_update"](self._h, item.ctypes.data, None, int(count), item.size)
                else:
                    c = _i64(count).ravel()
                    assert c.size == item.size
                    f["spsv_update"](self._h, item.ctypes.data, c.ctypes.data, 0, item.size)

            def merge(self, *others):
                for o in others:
                    f["spsv_merge"](self._h, o._h)

            def size(self):
                return f["spsv_size"](self._h)

            @property
            def capacity(self):
                return f["spsv_capacity"](self._h)

            def topk(self, k):
                k = min(int(k), self.size())
                out = np.empty(k, dtype=COUNTER_DTYPE)
                if k:
                    f["spsv_counters"](self._h, out.ctypes.data, k)
                return out

            def counters(self):
                return self.topk(self.size())

            def set_state(self, counters):
                c = np.ascontiguousarray(counters, dtype=COUNTER_DTYPE)
                return f["spsv_set_state"](self._h, c.ctypes.data, len(c))

        self.TDigest = TDigest
        self.SummaryStats = SummaryStats
        self.SpaceSaving = SpaceSaving


def _load():
    from . import build_oracle
    rest = _Namespace(build_oracle.build_restatement(), "orc_", "port")
    ref_path = build_oracle.build_reference()
    ref = _Namespace(ref_path, "ref_", "reference") if ref_path else None
    return rest, ref


restated, reference = _load()

# fp64-product switch of the restated SummaryStats (the CUDA path's variant)
def stats_fp_products(stats_obj, on=True):
    restated.lib.orc_stats_set_fp_products.argtypes = [_c_dp, _int]
    restated.lib.orc_stats_set_fp_products(stats_obj._h, int(on))
