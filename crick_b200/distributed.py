"""Multi-GPU reduction of sketches: one process per GPU, each builds its partial
sketch over its shard, the small fixed-size records are all-gathered, and every rank merges them
in rank order with the on-device merge kernel -- enqueued on the same stream directly behind the
collective, no host synchronisation in between.

torch.distributed is the plumbing here (NCCL over NVLink on GPUs, gloo on CPU for the tests);
libcrick_cuda itself stays free of any PyTorch dependency.
"""
import numpy as np

from .stats import SummaryStats
from .tdigest import TDigest


def shard_bounds(n, rank, world):
    """Contiguous equal element ranges, rank r owns [r*n/G, (r+1)*n/G)."""
    return (rank * n) // world, ((rank + 1) * n) // world


def record_from_state(centroids, total_weight, vmin, vmax, size):
    """Host mirror of k_pack_record: [n, total, min, max, mean0, w0, ...] padded to 4 + 2*size."""
    c = np.ascontiguousarray(centroids).view(np.float64).reshape(-1, 2)
    if len(c) > size:
        raise ValueError("more centroids than the digest has slots")
    rec = np.zeros(4 + 2 * size, dtype=np.float64)
    rec[0], rec[1], rec[2], rec[3] = len(c), total_weight, vmin, vmax
    rec[4:4 + 2 * len(c)] = c.ravel()
    return rec


def state_from_record(rec):
    """Inverse of record_from_state: (centroids (n, 2), total_weight, min, max)."""
    rec = np.asarray(rec, dtype=np.float64)
    n = int(rec[0])
    return rec[4:4 + 2 * n].reshape(n, 2).copy(), float(rec[1]), float(rec[2]), float(rec[3])


def gather_records(rec, group=None):
    """All-gather one fixed-size record per rank into a (world, len) tensor in RANK ORDER.
    Works on CUDA tensors (NCCL) and CPU tensors (gloo, used by the CPU tests)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    out = torch.empty((world, rec.numel()), dtype=rec.dtype, device=rec.device)
    if rec.is_cuda:
        dist.all_gather_into_tensor(out.view(-1), rec.contiguous().view(-1), group=group)
    else:
        parts = [torch.empty_like(rec) for _ in range(world)]
        dist.all_gather(parts, rec.contiguous(), group=group)
        for r, p in enumerate(parts):
            out[r] = p
    return out


class TDigestAllReduce:
    """Reusable state of the multi-GPU reduction of one digest per rank: the record buffers
    and the output digest are allocated once, so a step is export -> all-gather -> merge, three
    enqueues on the caller's stream and no allocation, free or host synchronisation."""

    def __init__(self, compression, group=None, exact=True):
        """exact=True: the merge is ``TDigest(c).merge(*digests)`` bit for bit (the reference's
        incremental tdigest_merge).  exact=False: one greedy pass over the sorted union of all
        ranks' centroids -- same accuracy, a fraction of the latency."""
        import torch
        import torch.distributed as dist

        self.group = group
        self.exact = exact
        self.world = dist.get_world_size(group)
        self.out = TDigest(compression)
        nd = self.out.record_doubles()
        self.rec = torch.empty(nd, dtype=torch.float64, device="cuda")
        self.gathered = torch.empty((self.world, nd), dtype=torch.float64, device="cuda")

    def __call__(self, digest):
        """``TDigest(c).merge(*[digest of rank 0, 1, ...])`` on every rank.  Enqueued on torch's
        current stream, which must not be the legacy default stream; the returned digest is
        reused by the next call."""
        import torch
        import torch.distributed as dist

        st = torch.cuda.current_stream().cuda_stream
        if st == 0:
            raise RuntimeError("run the reduction under `with torch.cuda.stream(s)`: the legacy "
                               "default stream cannot be named through the C ABI")
        digest.export_record(self.rec.data_ptr(), stream=st)
        dist.all_gather_into_tensor(self.gathered.view(-1), self.rec, group=self.group)
        self.out.clear(stream=st)
        self.out.merge_records(self.gathered.data_ptr(), self.world, stream=st,
                               bulk=not self.exact)
        return self.out


class ShardedTDigest:
    """One TDigest over a stream that is sharded across the GPUs of a node (BASELINE config 2).

    Every rank calls ``update(x_shard, w_shard)`` -- collectively, the same number of times --
    with CUDA arrays; afterwards ``digest`` holds, identically on every rank, the digest of ALL
    ranks' data seen so far.  The exchange does not go through a collective library: one kernel
    per rank turns the rank's bucket sums into <= 2 ov c weighted points, stores them into every
    rank's exchange buffer over NVLink (peer pointers obtained through CUDA IPC), waits for the
    others' points and folds the union into the digest in one greedy pass (``k_exchange_merge``).
    Setup (IPC handle exchange) uses ``torch.distributed``'s object all-gather once.

    Weights are validated on the device like ``update(..., check_w="deferred")``: a rank whose
    weights fail contributes nothing to that step and ``digest.check_weights()`` raises there.
    """

    def __init__(self, compression=100.0, group=None):
        import torch.distributed as dist
        from ._lib import check, check_handle, lib

        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.digest = TDigest(compression, mode="parallel")
        nbytes = lib.crick_p2p_buffer_bytes(float(self.digest.compression), self.world)
        self._own = check_handle(lib.crick_p2p_alloc(nbytes), "crick_p2p_alloc")
        import ctypes
        handle = ctypes.create_string_buffer(64)
        check(lib.crick_ipc_get_handle(self._own, handle), "ipc handle")
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
        else:
            handles[0] = bytes(handle.raw)
        self._opened = []
        ptrs = (ctypes.c_void_p * self.world)()
        for r in range(self.world):
            if r == self.rank:
                ptrs[r] = self._own
            else:
                p = check_handle(lib.crick_ipc_open_handle(ctypes.c_char_p(handles[r])), "ipc open")
                self._opened.append(p)
                ptrs[r] = p
        self._comm = check_handle(lib.crick_p2p_new(self.rank, self.world, ptrs), "crick_p2p_new")
        if self.world > 1:
            dist.barrier(group=group)       # every rank has mapped every buffer before the first step

    def update(self, x, w=1, stream=None):
        """Collective: this rank's shard (CUDA arrays; `w` a CUDA array or a scalar).  Returns the
        digest (valid once the stream has executed the call)."""
        import math
        from ._arrays import device_view
        from ._lib import check, lib

        xd = device_view(x)
        if xd is None or xd.dtype != np.float64:
            raise TypeError("x must be a float64 CUDA array")
        wd = device_view(w)
        if wd is not None:
            if wd.dtype != np.float64 or wd.size != xd.size:
                raise TypeError("w must be a float64 CUDA array shaped like x")
            wp, wsc = wd.ptr, 0.0
        else:
            wp, wsc = None, float(w)
            if wsc != 1 and (not math.isfinite(wsc) or wsc <= 0):
                raise ValueError("w must be > 0 and finite")
        check(lib.crick_tdigest_update_exchange(self.digest._h, self._comm, xd.ptr if xd.size else None, wp,
                                                wsc, xd.size, stream), "sharded update")
        self.digest._deferred = True
        return self.digest

    def close(self):
        from ._lib import lib
        if getattr(self, "_comm", None):
            lib.crick_device_synchronize()
            lib.crick_p2p_destroy(self._comm)
            self._comm = None
            for p in self._opened:
                lib.crick_ipc_close_handle(p)
            self._opened = []
            lib.crick_p2p_free(self._own)
            self._own = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def allgather_merge_tdigest(digest, group=None):
    """One-shot form: a new TDigest equal to ``TDigest(c).merge(*[digest of rank 0, 1, ...])``."""
    import torch

    s = torch.cuda.current_stream()
    if s.cuda_stream == 0:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
    red = TDigestAllReduce(digest.compression, group)
    with torch.cuda.stream(s):
        out = red(digest)
    s.synchronize()
    out._keepalive = red  # owns the gathered records the digest was merged from
    return out


def merge_tdigest_records_host(compression, records):
    """CPU-side plumbing test helper: merge records (list of 1-D float64 arrays laid out as
    [n, total, min, max, mean0, w0, ...]) in order into a fresh TDigest via the device kernel."""
    from ._arrays import DeviceArray
    recs = np.ascontiguousarray(np.stack(records), dtype=np.float64)
    d = DeviceArray.from_numpy(recs)
    out = TDigest(compression)
    out.merge_records(d.ptr, recs.shape[0])
    out.centroids()  # synchronise before `d` goes away
    return out


def spsv_record_from_counters(counters, capacity):
    """Host-side form of the record ``crick_spsv_export_record`` writes: int64
    ``[size, 0, (item, count, error) x capacity]`` in list order, zero padded."""
    rec = np.zeros(2 + 3 * int(capacity), dtype=np.int64)
    n = len(counters)
    rec[0] = n
    if n:
        body = np.stack([np.asarray(counters[f]).view(np.int64) for f in ("item", "count", "error")], 1)
        rec[2:2 + 3 * n] = body.ravel()
    return rec


def spsv_counters_from_record(rec):
    """Inverse of :func:`spsv_record_from_counters`: an (n, 3) int64 array."""
    n = int(rec[0])
    return np.asarray(rec[2:2 + 3 * n], dtype=np.int64).reshape(n, 3).copy()


def allgather_merge_spsv(sp, group=None):
    """``SpaceSaving(capacity).merge(*[summary of rank 0, 1, ...])`` on every rank: the fixed
    24 KB (capacity 1000) int64 records are all-gathered and merged on the device in rank order
    with the reference's merge (space_saving_stubs.c.in:289-364)."""
    import torch
    import torch.distributed as dist
    from .space_saving import SpaceSaving

    world = dist.get_world_size(group)
    s = torch.cuda.current_stream()
    if s.cuda_stream == 0:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
    nd = sp.record_int64s()
    with torch.cuda.stream(s):
        rec = torch.empty(nd, dtype=torch.int64, device="cuda")
        gathered = torch.empty((world, nd), dtype=torch.int64, device="cuda")
        sp.export_record(rec.data_ptr(), stream=s.cuda_stream)
        dist.all_gather_into_tensor(gathered.view(-1), rec, group=group)
        out = SpaceSaving(sp.capacity, sp.dtype)
        out.merge_records(gathered.data_ptr(), world, stream=s.cuda_stream)
    s.synchronize()
    return out


class SketchAllReduce:
    """SpaceSaving + SummaryStats of every rank reduced in one exchange (BASELINE config 5): the
    summary record ``[size, 0, (item, count, error) x capacity]`` and the 9-double moments record
    travel as ONE int64 buffer per rank, one all-gather, then both merges on the device on the same
    stream -- no allocation, no host synchronisation.  ``exact=True``: SpaceSaving records are
    merged one after the other like ``SpaceSaving.merge(rank0, rank1, ...)`` (bit for bit);
    ``exact=False``: the k-way form of the same merge in one kernel."""

    def __init__(self, capacity, dtype="i8", group=None, exact=False):
        import torch
        import torch.distributed as dist
        from .space_saving import SpaceSaving

        self.group, self.exact = group, exact
        self.world = dist.get_world_size(group)
        self.summary = SpaceSaving(capacity, dtype)
        self.stats = SummaryStats()
        self.nd = self.summary.record_int64s()
        self.rec = torch.zeros(self.nd + 9, dtype=torch.int64, device="cuda")
        self.gathered = torch.empty((self.world, self.nd + 9), dtype=torch.int64, device="cuda")

    def __call__(self, summary, stats=None):
        """Returns (merged SpaceSaving, merged SummaryStats or None), both reused by the next call.
        Enqueued on torch's current stream (not the legacy default stream)."""
        import torch
        import torch.distributed as dist

        st = torch.cuda.current_stream().cuda_stream
        if st == 0:
            raise RuntimeError("run the reduction under `with torch.cuda.stream(s)`")
        summary.export_record(self.rec.data_ptr(), stream=st)
        if stats is not None:
            stats.export_record_device(self.rec.data_ptr() + 8 * self.nd, stream=st)
        dist.all_gather_into_tensor(self.gathered.view(-1), self.rec, group=self.group)
        self.summary.clear(stream=st)
        row = self.nd + 9
        # the records are `row` int64 apart and the merges take a contiguous block of summary
        # records: compact them first (one strided device copy)
        if not hasattr(self, "_dense"):
            self._dense = torch.empty((self.world, self.nd), dtype=torch.int64, device="cuda")
        self._dense.copy_(self.gathered[:, :self.nd])
        self.summary.merge_records(self._dense.data_ptr(), self.world, stream=st, bulk=not self.exact)
        if stats is not None:
            self.stats.clear(stream=st)
            self.stats.merge_records_device(self.gathered.data_ptr() + 8 * self.nd, self.world, stride=row,
                                            stream=st)
            return self.summary, self.stats
        return self.summary, None


class NcclComm:
    """The reduction without PyTorch: a NCCL communicator owned by libcrick_cuda (``crick_comm_*``,
    NCCL opened with dlopen), one process per GPU.  Rank 0 calls ``NcclComm.unique_id()`` and ships
    the 128 bytes to the other ranks by any means; every rank then constructs
    ``NcclComm(id, rank, world)`` (collective) on its current device (``crick_set_device``).

    ``reduce_tdigest`` / ``reduce_space_saving`` / ``reduce_stats`` leave, identically on every
    rank, ``out.merge(sketch of rank 0, ..., sketch of rank world - 1)`` in ``out``: export ->
    ncclAllGather -> merge kernel, three enqueues on ``stream`` and no host synchronisation."""

    def __init__(self, unique_id, rank, world):
        from ._lib import check_handle, lib
        if len(unique_id) != 128:
            raise ValueError("unique_id must be the 128 bytes NcclComm.unique_id() returned on rank 0")
        self.rank, self.world = int(rank), int(world)
        self._h = None
        self._h = check_handle(lib.crick_comm_init(bytes(unique_id), self.rank, self.world), "crick_comm_init")
        self.stream = check_handle(lib.crick_stream_create(), "crick_stream_create")

    @staticmethod
    def unique_id():
        import ctypes
        from ._lib import check, lib
        buf = ctypes.create_string_buffer(128)
        check(lib.crick_comm_unique_id(buf), "crick_comm_unique_id")
        return bytes(buf.raw)

    def close(self):
        from ._lib import lib
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.crick_stream_synchronize(self.stream)
            lib.crick_comm_free(h)
            lib.crick_stream_destroy(self.stream)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        from ._lib import check, lib
        check(lib.crick_stream_synchronize(self.stream), "stream synchronize")

    def reduce_tdigest(self, digest, out=None, exact=False, stream=None):
        from ._lib import check, lib
        out = TDigest(digest.compression) if out is None else out
        check(lib.crick_tdigest_allgather_merge(digest._h, self._h, out._h, int(bool(exact)),
                                                stream or self.stream), "crick_tdigest_allgather_merge")
        return out

    def reduce_space_saving(self, summary, out=None, exact=False, stream=None):
        from ._lib import check, lib
        from .space_saving import SpaceSaving
        out = SpaceSaving(summary.capacity, summary.dtype) if out is None else out
        check(lib.crick_spsv_allgather_merge(summary._h, self._h, out._h, int(bool(exact)),
                                             stream or self.stream), "crick_spsv_allgather_merge")
        return out

    def reduce_stats(self, stats, out=None, stream=None):
        from ._lib import check, lib
        out = SummaryStats() if out is None else out
        check(lib.crick_stats_allgather_merge(stats._h, self._h, out._h, stream or self.stream),
              "crick_stats_allgather_merge")
        return out


def allgather_merge_stats(stats, group=None, backend_device="cuda"):
    """All-gather the 9-double SummaryStats records and merge them in rank order."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rec = torch.from_numpy(stats.export_record()).to(backend_device)
    gathered = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(gathered, rec, group=group)
    out = SummaryStats()
    for r in gathered:
        out.merge_record(r.cpu().numpy())
    return out
