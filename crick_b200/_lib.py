"""ctypes binding of libcrick_cuda.so (include/crick_cuda.h).

There is deliberately no fallback: if the shared library is missing the import
fails, and if there is no CUDA device every constructor raises.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# CRICK_CUDA_LIB: another build of the same library (A/B kernel experiments on one box)
LIB_PATH = os.environ.get("CRICK_CUDA_LIB") or os.path.join(_HERE, "libcrick_cuda.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "crick_b200: %s is missing. Build it with `make -C crick_b200/csrc` "
        "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
        "There is no CPU fallback." % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

_P = c_void_p
_SIG = {
    # library
    "crick_last_error": (c_char_p, []),
    "crick_version": (c_char_p, []),
    "crick_device_count": (c_int, []),
    "crick_set_device": (c_int, [c_int]),
    "crick_sm_count": (c_int, []),
    "crick_get_device": (c_int, []),
    "crick_cast_to_f64": (c_int, [_P, c_int, c_int, c_int64, _P, _P]),
    "crick_check_weights": (c_int, [_P, c_int64, _P]),
    "crick_launch_count": (c_int64, []),
    "crick_profile_enable": (c_int, [c_int]),
    "crick_group_kernel": (c_int, [c_int]),
    "crick_group_margin": (c_double, [c_double]),
    "crick_stream_create": (_P, []),
    "crick_stream_destroy": (None, [_P]),
    "crick_tdigest_update_pipelined": (c_int, [_P, _P, _P, c_double, c_int64, _P, _P]),
    "crick_pipeline_reserve_sms": (c_int, [c_int]),
    "crick_pipeline_overlap": (c_int, [c_int]),
    "crick_comm_unique_id": (c_int, [_P]),
    "crick_comm_init": (_P, [_P, c_int, c_int]),
    "crick_comm_free": (None, [_P]),
    "crick_comm_rank": (c_int, [_P]),
    "crick_comm_world": (c_int, [_P]),
    "crick_comm_allgather": (c_int, [_P, _P, _P, c_int64, _P]),
    "crick_tdigest_allgather_merge": (c_int, [_P, _P, _P, c_int, _P]),
    "crick_spsv_allgather_merge": (c_int, [_P, _P, _P, c_int, _P]),
    "crick_stats_allgather_merge": (c_int, [_P, _P, _P, _P]),
    "crick_group_exact_fallbacks": (c_int, [_P, c_int]),
    "crick_stats_kernel": (c_int, [c_int]),
    "crick_profile_collect": (c_int, [_P, _P]),
    "crick_host_alloc": (_P, [c_int64]),
    "crick_host_free": (None, [_P]),
    "crick_device_alloc": (_P, [c_int64]),
    "crick_device_free": (None, [_P]),
    "crick_memcpy_h2d": (c_int, [_P, _P, c_int64, _P]),
    "crick_memcpy_d2h": (c_int, [_P, _P, c_int64, _P]),
    "crick_stream_synchronize": (c_int, [_P]),
    "crick_device_synchronize": (c_int, []),
    "crick_synth_fill": (c_int, [_P, _P, c_int64, c_uint64, c_int64, c_int, _P]),
    # tdigest
    "crick_tdigest_new": (_P, [c_double]),
    "crick_tdigest_free": (None, [_P]),
    "crick_tdigest_add": (c_int, [_P, c_double, c_double]),
    "crick_tdigest_update": (c_int, [_P, _P, _P, c_double, c_int64, c_int, c_int, c_int, _P]),
    "crick_tdigest_update_checked": (c_int, [_P, _P, _P, c_double, c_int64, c_int, c_int, c_int, _P, _P]),
    "crick_tdigest_update_deferred": (c_int, [_P, _P, _P, c_double, c_int64, c_int, _P]),
    "crick_tdigest_take_bad_weights": (c_int, [_P]),
    "crick_p2p_buffer_bytes": (c_int64, [c_double, c_int]),
    "crick_p2p_alloc": (_P, [c_int64]),
    "crick_p2p_free": (None, [_P]),
    "crick_ipc_get_handle": (c_int, [_P, _P]),
    "crick_ipc_open_handle": (_P, [_P]),
    "crick_ipc_close_handle": (c_int, [_P]),
    "crick_p2p_new": (_P, [c_int, c_int, _P]),
    "crick_p2p_destroy": (None, [_P]),
    "crick_tdigest_update_exchange": (c_int, [_P, _P, _P, _P, c_double, c_int64, _P]),
    "crick_tdigest_update_stats": (c_int, [_P, _P, _P, _P, c_double, c_int64, c_int, c_int, _P, _P]),
    "crick_tdigest_flush": (c_int, [_P]),
    "crick_tdigest_merge": (c_int, [_P, _P, c_int]),
    "crick_tdigest_scale": (c_int, [_P, c_double]),
    "crick_tdigest_quantile": (c_int, [_P, _P, _P, c_int64, c_int, _P]),
    "crick_tdigest_cdf": (c_int, [_P, _P, _P, c_int64, c_int, _P]),
    "crick_tdigest_compression": (c_double, [_P]),
    "crick_tdigest_size": (c_int, [_P]),
    "crick_tdigest_buffer_size": (c_int, [_P]),
    "crick_tdigest_min": (c_double, [_P]),
    "crick_tdigest_max": (c_double, [_P]),
    "crick_tdigest_ncentroids": (c_int, [_P]),
    "crick_tdigest_buffer_ncentroids": (c_int, [_P]),
    "crick_tdigest_total_weight": (c_double, [_P]),
    "crick_tdigest_buffer_total_weight": (c_double, [_P]),
    "crick_tdigest_get_centroids": (c_int, [_P, _P, c_int]),
    "crick_tdigest_centroids_device": (_P, [_P]),
    "crick_tdigest_set_state": (c_int, [_P, _P, c_int, c_double, c_double, c_double]),
    "crick_tdigest_reset": (c_int, [_P, _P]),
    "crick_tdigest_debug_header": (c_int, [_P, _P]),
    "crick_tdigest_debug_edges": (c_int, [_P, _P, c_int]),
    "crick_tdigest_debug_buckets": (c_int, [_P, _P, _P, _P, c_int]),
    "crick_tdigest_record_doubles": (c_int, [_P]),
    "crick_tdigest_export_record": (c_int, [_P, _P, _P]),
    "crick_tdigest_merge_records": (c_int, [_P, _P, c_int, c_int, _P]),
    "crick_tdigest_merge_records_bulk": (c_int, [_P, _P, c_int, _P]),
    # grouped
    "crick_tdigest_group_new": (_P, [c_double, c_int64]),
    "crick_tdigest_group_free": (None, [_P]),
    "crick_tdigest_group_count": (c_int64, [_P]),
    "crick_tdigest_group_update": (c_int, [_P, _P, _P, c_double, _P, c_int64, c_int64, c_int, _P]),
    "crick_tdigest_group_flush": (c_int, [_P]),
    "crick_tdigest_group_merge_tree": (c_int, [_P, _P]),
    "crick_tdigest_group_merge_all": (c_int, [_P, _P]),
    "crick_tdigest_group_quantile": (c_int, [_P, _P, c_int, _P, c_int, _P]),
    "crick_tdigest_group_cdf": (c_int, [_P, _P, c_int, _P, c_int, _P]),
    "crick_tdigest_group_meta": (c_int, [_P, _P]),
    "crick_tdigest_group_get_centroids": (c_int, [_P, c_int64, _P]),
    "crick_tdigest_group_extract": (c_int, [_P, c_int64, _P]),
    # stats
    "crick_stats_new": (_P, []),
    "crick_stats_free": (None, [_P]),
    "crick_stats_add": (c_int, [_P, c_double, c_int64]),
    "crick_stats_update": (c_int, [_P, _P, c_int, _P, c_int64, c_int64, c_int, c_int, _P]),
    "crick_stats_merge": (c_int, [_P, _P, c_int]),
    "crick_stats_getstate": (c_int, [_P, _P, _P, _P]),
    "crick_stats_setstate": (c_int, [_P, c_int64, _P, c_int]),
    "crick_stats_mean": (c_double, [_P]),
    "crick_stats_var": (c_double, [_P, c_int64]),
    "crick_stats_std": (c_double, [_P, c_int64]),
    "crick_stats_skew": (c_double, [_P, c_int]),
    "crick_stats_kurt": (c_double, [_P, c_int, c_int]),
    "crick_stats_export_record": (c_int, [_P, _P]),
    "crick_stats_merge_record": (c_int, [_P, _P]),
    "crick_stats_export_record_dev": (c_int, [_P, _P, _P]),
    "crick_stats_merge_records_dev": (c_int, [_P, _P, c_int, c_int64, _P]),
    # space saving
    "crick_spsv_new": (_P, [c_int64]),
    "crick_spsv_free": (None, [_P]),
    "crick_spsv_add": (c_int, [_P, c_int64, c_int64]),
    "crick_spsv_update": (c_int, [_P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, _P]),
    "crick_spsv_update_stats": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P]),
    "crick_spsv_merge": (c_int, [_P, _P, c_int]),
    "crick_spsv_set_state": (c_int, [_P, _P, c_int64]),
    "crick_spsv_reset": (c_int, [_P, _P]),
    "crick_stats_reset": (c_int, [_P, _P]),
    "crick_spsv_record_int64s": (c_int64, [_P]),
    "crick_spsv_export_record": (c_int, [_P, _P, _P]),
    "crick_spsv_merge_records": (c_int, [_P, _P, c_int, _P]),
    "crick_spsv_merge_records_bulk": (c_int, [_P, _P, c_int, _P]),
    "crick_spsv_size": (c_int64, [_P]),
    "crick_spsv_capacity": (c_int64, [_P]),
    "crick_spsv_counters": (c_int64, [_P, _P, c_int64]),
}

for _name, (_res, _args) in _SIG.items():
    _fn = getattr(lib, _name)  # AttributeError here = header/library mismatch: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args

MODE_AUTO, MODE_REFERENCE, MODE_PARALLEL = 0, 1, 2
PARALLEL_MIN = 2048
_MODES = {"auto": MODE_AUTO, "reference": MODE_REFERENCE, "parallel": MODE_PARALLEL}


def mode_code(mode):
    try:
        return _MODES[mode]
    except KeyError:
        raise ValueError("mode must be one of 'auto', 'reference', 'parallel'")


def last_error():
    msg = lib.crick_last_error()
    return msg.decode() if msg else ""


def check(rc, what="libcrick_cuda"):
    if rc is not None and rc < 0:
        raise RuntimeError("%s failed: %s" % (what, last_error()))
    return rc


def check_handle(h, what):
    if not h:
        msg = last_error()
        if "cudaMalloc" in msg or "out of memory" in msg:
            raise MemoryError(msg)
        raise RuntimeError("%s failed: %s" % (what, msg or "unknown error"))
    return h


def handle_array(handles):
    arr = (c_void_p * len(handles))(*handles)
    return arr
