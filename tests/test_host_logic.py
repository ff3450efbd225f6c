"""CPU: host-side logic of the wrappers that needs no device -- K-order flattening (the
reference's NpyIter order, tdigest_stubs.c:652-666), safe-cast rules, CUDA-array parsing."""
import numpy as np
import pytest

import oracle
from crick_b200._arrays import _contiguous, device_view, korder_pair
from helpers import sha


def test_korder_matches_reference_iteration_order(golden):
    """Flatten with korder_pair, feed the oracle, and require the reference's own sha for
    every stride/shape case recorded from the built reference (make_golden.py)."""
    g, a = golden
    base, wv = a["korder_base"], a["korder_w"]
    cases = {"rev": (base.ravel()[::-1], 1), "fortran": (np.asfortranarray(base), 1),
             "transposed": (base.T, 1), "strided": (base.ravel()[::2], 1),
             "rev_fw": (base.ravel()[::-1], wv.ravel()), "bcast_w": (base, wv[0]),
             "t_w": (base.T, wv.T), "scalar_x": (1.5, wv.ravel()[:100])}
    for name, (x, w) in cases.items():
        xf, wf = korder_pair(x, w)
        t = oracle.restated.TDigest(100)
        t.update(xf, wf)
        assert sha(t) == g["korder"][name], name


def test_korder_fast_paths_do_not_copy():
    x = np.arange(12.0).reshape(3, 4)
    xf, wf = korder_pair(x, 2.5)
    assert xf.base is not None and np.shares_memory(xf, x) and wf == 2.5
    w = np.ones((3, 4))
    xf, wf = korder_pair(x, w)
    assert np.shares_memory(xf, x) and np.shares_memory(wf, w)


def test_korder_casting_rules():
    xf, wf = korder_pair(np.arange(5, dtype="i4"), np.float32(2))
    assert xf.dtype == np.float64 and xf.tolist() == [0, 1, 2, 3, 4] and wf == 2.0
    with pytest.raises(TypeError):
        korder_pair(np.array([1 + 2j]), 1)
    xf, cf = korder_pair(np.arange(3.0), np.array([1, 2, 3], dtype="i2"), w_dtype="i8")
    assert cf.dtype == np.int64
    with pytest.raises(TypeError):
        korder_pair(np.arange(3.0), np.array([1.5, 2, 3]), w_dtype="i8")
    with pytest.raises(ValueError):
        korder_pair(np.arange(6.0).reshape(2, 3), np.arange(4.0))  # not broadcastable
    xf, wf = korder_pair(np.empty((0,)), 1)
    assert xf.size == 0


def test_contiguity_check():
    assert _contiguous((3, 4), None, 8)
    assert _contiguous((3, 4), (32, 8), 8)
    assert not _contiguous((3, 4), (8, 24), 8)
    assert _contiguous((1, 4), (999, 8), 8)


class _FakeCuda:
    def __init__(self, ptr, shape, typestr="<f8", strides=None):
        self.__cuda_array_interface__ = {"data": (ptr, False), "shape": shape, "typestr": typestr,
                                         "strides": strides, "version": 3}


def test_device_view_parses_cuda_array_interface():
    v = device_view(_FakeCuda(0x1000, (10, 3)))
    assert v.ptr == 0x1000 and v.size == 30 and v.dtype == np.float64 and v.shape == (10, 3)
    v = device_view(_FakeCuda(0x2000, (7,), "<i8"))
    assert v.dtype == np.int64
    with pytest.raises(ValueError):
        device_view(_FakeCuda(0x1000, (10, 3), strides=(8, 80)))
    assert device_view(np.arange(3.0)) is None
    assert device_view([1, 2, 3]) is None


def test_device_view_ignores_cpu_dlpack():
    torch = pytest.importorskip("torch")
    assert device_view(torch.arange(4, dtype=torch.float64)) is None


def test_nccl_comm_rejects_a_malformed_unique_id():
    """NcclComm checks the 128-byte id before anything touches NCCL or a device."""
    from crick_b200.distributed import NcclComm
    with pytest.raises(ValueError):
        NcclComm(b"short", 0, 1)


def test_shard_bounds_cover_the_stream_exactly():
    """Strong scaling shards ONE stream by contiguous ranges (bench.py): the ranges of
    all ranks tile [0, n) without gaps or overlap, for world sizes that do not divide n as well."""
    from crick_b200.distributed import shard_bounds
    for n in (10**9, 10**9 + 7, 5, 0):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_bounds(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            assert max(hi - lo for lo, hi in edges) - min(hi - lo for lo, hi in edges) <= 1
