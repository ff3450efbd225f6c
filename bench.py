#!/usr/bin/env python
"""bench.py -- headline benchmark of the crick hot path on B200 (contract: see DESIGN.md 5).

Metric (BASELINE.json): TDigest.update Gvalues/s over ONE stream of 1e9 float64 values with
float64 weights (lognormal x, U(0.5,1.5) w, compression 100) -- BASELINE.json configs[1] --
sharded over the N GPUs of the node: rank r owns the contiguous range [r n / N, (r + 1) n / N)
(strong scaling: the total stays 1e9), builds its partial digest, and with N > 1 every step ends
with the all-gather of the 3.2 KB centroid records and the on-device merge in rank order.  The
weak-scaling figure (1e9 values per GPU) of the same step is reported next to it under "weak".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n VALUES]

One JSON line on rank 0.  `value` = device-resident throughput (inputs in HBM before the timed
region), `e2e` = the same through TDigest.update() on pinned HOST arrays (H2D inside the timed
region, centroids read back every step), `e2e_pageable` = the same on ordinary NumPy arrays,
`roofline` = the hot kernel (k_bin_accumulate2) against the measured HBM peak (weighted,
16 B/value) and `roofline_unweighted` the w = 1 call dask makes (8 B/value), `cpu_baseline` = the
reference's own C code on this box's host cores, `other_configs` = BASELINE configs 3, 4 and 5
at full size (value, roofline fraction, parity mismatches against the oracle, CPU baseline).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "TDigest.update throughput (1e9 f64 values + f64 weights, compression=100)"
UNIT = "Gvalues/s"
N_DEFAULT = 10**9
COMPRESSION = 100.0
SEED = 1234
ALGO_BYTES_PER_VALUE = 16  # 8 B x + 8 B w, read once


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_DEFAULT, help="values of the whole stream per step")
    ap.add_argument("--cpu-sample", type=int, default=2 * 10**8)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip BASELINE configs 3, 4, 5")
    ap.add_argument("--no-weak", action="store_true", help="skip the weak-scaling leg (N > 1)")
    ap.add_argument("--overlap", type=int, default=-1, help="overlapped sample phase: 1 / 0 (default 0: see DESIGN.md 5.1)")
    ap.add_argument("--reserve", type=int, default=-1, help="SMs the pass leaves free (default 0 / 2 / 6)")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="plain update calls instead of TDigest.update(..., tail_stream=)")
    return ap.parse_args()


def peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def shard_bounds(n, rank, world):
    return (rank * n) // world, ((rank + 1) * n) // world


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [c.strip() for c in r.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_run(n, threads):
    """The reference's own C code (oracle/_ref, else the oracle port) on host cores: shard per
    thread + merge in shard order (the dask pattern, tdigest.pyx:282-324).  Returns
    (Gvalues/s, kind, seconds)."""
    import oracle
    ns = oracle.reference if oracle.reference is not None else oracle.restated
    rng = np.random.default_rng(SEED)
    x = rng.lognormal(0.0, 1.0, n)
    w = rng.uniform(0.5, 1.5, n)
    t0 = time.perf_counter()
    d = ns.TDigest.update_sharded(COMPRESSION, x, w, nthreads=threads)
    nc = len(d.centroids())
    dt = time.perf_counter() - t0
    return n / dt / 1e9, ns.kind, dt, nc


def run_reference(args, rank, world):
    """--impl reference: the reference CPU implementation, all host threads, bounded sample."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.cpu_sample
    for _ in range(min(args.warmup, 1)):
        cpu_reference_run(max(n // 20, 10**6), threads)
    vals, secs = [], 0.0
    steps = max(1, min(args.steps, 3))
    kind = "port"
    for _ in range(steps):
        v, kind, dt, _nc = cpu_reference_run(n, threads)
        vals.append(v); secs += dt
    value = float(np.mean(vals))
    sample = "%d lognormal f64 values + f64 weights per step (bounded sample of the 1e9 workload)" % n
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * secs / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "config2: single TDigest compression=100 over lognormal f64 + f64 weights",
                   "values_per_step": n, "host_threads": threads,
                   "pattern": "one digest per thread-shard, merged in shard order"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


_STDOUT_FD = None


def emit(obj):
    """The one JSON line of the contract, on the process's original stdout."""
    line = (json.dumps(obj) + "\n").encode()
    if _STDOUT_FD is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_STDOUT_FD, line)


class Bench:
    """State shared by the legs of one run (one process per GPU)."""

    def __init__(self, args):
        import torch
        import crick_b200 as cb
        from crick_b200._lib import lib
        self.args, self.torch, self.cb, self.lib = args, torch, cb, lib
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(self.local_rank)
        lib.crick_set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        self.stream = torch.cuda.Stream()
        self.sptr = self.stream.cuda_stream
        # pipelined updates (TDigest.update(..., tail_stream=)): the digest-side tail of a step and the
        # reduction run here and overlap the next step's sample phase on self.stream
        self.tail = torch.cuda.Stream()
        self.tptr = self.tail.cuda_stream
        self.pipelined = True
        self.peak, self.peak_src = peak_hbm()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.dist is None:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        if self.dist is None:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def profile_collect(self):
        kms, kn = ctypes.c_double(0), ctypes.c_int64(0)
        self.lib.crick_profile_collect(ctypes.addressof(kms), ctypes.addressof(kn))
        return kms.value, kn.value

    # ---- config 2: one stream, `m` values on this rank -----------------------------------
    def timed_updates(self, dx, dw, K, W, reducer, sampler=None):
        """W warm-up + K timed steps of update(+ all-gather + merge); device time on the launching
        stream, max over ranks.  Returns (ms, kernel ms total, kernel launches, merged digest,
        library launches)."""
        cb, lib, torch = self.cb, self.lib, self.torch
        digest = cb.TDigest(COMPRESSION, mode="parallel")

        def step():
            # weights validated on the device ("deferred": the verdict is collected after the
            # timed region, nothing waits for it inside a step)
            if self.pipelined:
                digest.update(dx, 1 if dw is None else dw, stream=self.sptr, tail_stream=self.tptr)
            elif dw is not None:
                digest.update(dx, dw, mode="parallel", stream=self.sptr, check_w="deferred")
            else:
                digest.update(dx, mode="parallel", stream=self.sptr)
            if reducer is not None:
                with torch.cuda.stream(self.tail if self.pipelined else self.stream):
                    return reducer(digest)
            return digest

        merged = None
        for _ in range(W):
            merged = step()
        self.barrier()
        launches0 = cb.launch_count()
        lib.crick_profile_enable(1)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        ev0.record(self.stream)
        for _ in range(K):
            merged = step()
        if self.pipelined:
            self.stream.wait_stream(self.tail)     # the last step's tail and reduction are inside
        ev1.record(self.stream)
        self.barrier()
        ms = self.max_over_ranks(ev0.elapsed_time(ev1))
        lib.crick_profile_enable(0)
        kms, kn = self.profile_collect()
        launches = cb.launch_count() - launches0
        if sampler is not None:
            # nvidia-smi needs ~0.1-0.3 s for its first line and the timed region may be shorter:
            # keep the identical load running (untimed) until a few samples exist
            n_extra = int(min(400, max(0, 600.0 / max(ms / K, 1e-3))))
            for _ in range(n_extra):
                step()
            torch.cuda.synchronize()
        digest.check_weights()      # raises if any deferred weight check failed
        return ms, kms, kn, merged, launches


def kernel_roofline(bench, name, k_ms_total, k_launches, bytes_per_launch, step_ms_total, traffic_key):
    k_ms = k_ms_total / max(k_launches, 1)
    achieved = bytes_per_launch / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    r = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": bench.peak, "unit": "GB/s",
         "frac": achieved / bench.peak, "traffic": None, "peak_source": bench.peak_src,
         "kernel_ms_per_launch": k_ms,
         "kernel_share_of_step": (k_ms_total / step_ms_total) if step_ms_total else None,
         "algorithmic_bytes_per_launch": bytes_per_launch,
         # SURVEY.md 8d asks for both denominators: the box-measured copy bandwidth above and
         # the nominal 8 TB/s of HBM3e
         "peak_nominal": 8000.0, "frac_nominal": achieved / 8000.0}
    try:  # DRAM bytes per launch from the committed ncu --set full capture, scaled to this launch
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        r["traffic"] = tj[traffic_key] * (bytes_per_launch / tj[traffic_key + "_algorithmic"])
        r["traffic_source"] = tj["source"]
    except Exception:
        pass
    return r


def main():
    # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version
    # banner under torchrun, for one) goes to stderr instead
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        return

    B = Bench(args)
    B.pipelined = not args.no_pipeline
    # How the pipelined call is set up (library switches).  One GPU: the tail is
    # the 20 us of k_finalize -- no SM reserve.  More: 2 SMs for the reduction on the tail stream (an
    # NCCL kernel waiting for the other ranks would otherwise hold SMs the pass wants).  The overlapped
    # sample phase (--overlap 1) is NOT used: it is faster (0.36 vs 0.40 ms per step at 8 GPUs) but came
    # out non-deterministic on updates of 4 Mi values (tests/tools/pipe_race.py), cause not found.
    overlap = bool(args.overlap) if args.overlap >= 0 else False
    reserve = args.reserve if args.reserve >= 0 else (6 if overlap else (2 if B.world > 1 else 0))
    B.lib.crick_pipeline_overlap(1 if overlap else 0)
    B.lib.crick_pipeline_reserve_sms(reserve)
    B.pipe_cfg = {"overlap_sample_phase": bool(overlap), "reserved_sms": int(reserve)}
    cb, lib, torch = B.cb, B.lib, B.torch
    rank, world = B.rank, B.world
    from crick_b200.distributed import TDigestAllReduce, allgather_merge_tdigest

    n = args.n
    K, W = args.steps, args.warmup
    lo, hi = shard_bounds(n, rank, world)
    m = hi - lo

    # ---- synthetic, device-resident shard: element i of the stream depends on (seed, i) only
    dx, dw = cb.DeviceArray((m,)), cb.DeviceArray((m,))
    lib.crick_synth_fill(dx.ptr, dw.ptr, m, SEED, lo, 0, None)
    lib.crick_device_synchronize()
    reducer = TDigestAllReduce(COMPRESSION, exact=False) if world > 1 else None

    # clocks are sampled every 100 ms from before the warm-up (the same load) to the end of
    # the timed region, which by itself may be shorter than one sample period
    sampler = ClockSampler(B.local_rank)
    sampler.start()
    ms, kms, kn, merged, launches = B.timed_updates(dx, dw, K, W, reducer, sampler)
    clocks = sampler.stop()
    clocks["note"] = "sampled every 100 ms from the warm-up through the timed region and an identical untimed tail"
    value = n * K / (ms * 1e-3) / 1e9
    nc = len(merged.centroids())
    total_size = merged.size()
    roofline = kernel_roofline(B, "k_bin_accumulate2<weighted>", kms, kn, ALGO_BYTES_PER_VALUE * m, ms,
                               "k_bin_accumulate2_weighted_dram_bytes")

    # ---- the unweighted call (w = 1: what dask's percentile makes), same shard, 8 B/value
    ums, ukms, ukn, _, _ = B.timed_updates(dx, None, K, W, reducer)
    roofline_unweighted = kernel_roofline(B, "k_bin_accumulate2<unweighted>", ukms, ukn, 8 * m, ums,
                                          "k_bin_accumulate2_unweighted_dram_bytes")
    roofline_unweighted["value_gvalues_s"] = n * K / (ums * 1e-3) / 1e9
    roofline_unweighted["ms_per_step"] = ums / K

    # ---- the same step with the exchange fused into one kernel over NVLink peer memory
    # (ShardedTDigest / k_exchange_merge: no collective library call); N > 1 only
    p2p = None
    if world > 1:
        try:
            from crick_b200.distributed import ShardedTDigest
            sh = ShardedTDigest(COMPRESSION)
            for _ in range(W):
                sh.update(dx, dw, stream=B.sptr)
            B.barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(B.stream)
            for _ in range(K):
                sh.update(dx, dw, stream=B.sptr)
            ev1.record(B.stream)
            B.barrier()
            pms = B.max_over_ranks(ev0.elapsed_time(ev1))
            sh.digest.check_weights()
            p2p = {"value": n * K / (pms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": pms / K,
                   "centroids": len(sh.digest.centroids()),
                   "what": "ShardedTDigest.update: bucket points stored into every rank's buffer over NVLink "
                           "(CUDA IPC peer pointers), flags, one merge -- k_exchange_merge instead of "
                           "finalize + pack + NCCL all-gather + merge"}
            sh.close()
        except Exception as exc:
            p2p = {"error": "%s: %s" % (type(exc).__name__, exc)}

    # ---- weak scaling next to it (N > 1): 1e9 values on EVERY GPU, same step
    weak = None
    if world > 1 and not args.no_weak:
        del dx, dw
        dx, dw = cb.DeviceArray((n,)), cb.DeviceArray((n,))
        lib.crick_synth_fill(dx.ptr, dw.ptr, n, SEED, rank * n, 0, None)
        lib.crick_device_synchronize()
        wms, wkms, wkn, _, _ = B.timed_updates(dx, dw, K, W, reducer)
        weak = {"scaling": "weak", "value": world * n * K / (wms * 1e-3) / 1e9, "unit": UNIT,
                "ms_per_step": wms / K, "values_per_gpu_per_step": n,
                "kernel_frac_of_peak": (ALGO_BYTES_PER_VALUE * n) / (wkms / max(wkn, 1) * 1e-3) / 1e9 / B.peak}
        del dx, dw
        dx, dw = cb.DeviceArray((m,)), cb.DeviceArray((m,))
        lib.crick_synth_fill(dx.ptr, dw.ptr, m, SEED, lo, 0, None)
        lib.crick_device_synchronize()

    # ---- e2e: public API on HOST arrays; H2D every step, centroids read back every step
    e2e = e2e_pageable = None
    if not args.no_e2e:
        import psutil
        me = m
        while 16 * me * world * 2 > psutil.virtual_memory().available and me > 10**7:
            me //= 2

        def e2e_leg(hx, hw, steps, api):
            def e2e_step():
                t = cb.TDigest(COMPRESSION, mode="parallel")
                t.update(hx, hw)
                if world > 1:
                    t = allgather_merge_tdigest(t)
                return t.centroids()
            e2e_step()
            B.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                cen = e2e_step()
            B.barrier()
            dt = B.max_over_ranks(time.perf_counter() - t0)
            tot = B.sum_over_ranks(me)
            return {"value": tot * steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": 16 * me,
                    "d2h_bytes_per_step": int(cen.nbytes), "values_per_step_per_gpu": me,
                    "steps": steps, "h2d_gbs_per_rank": 16 * me * steps / dt / 1e9, "api": api}
        try:
            hx, hw = cb.PinnedArray(me), cb.PinnedArray(me)
            lib.crick_memcpy_d2h(hx.ptr, dx.ptr, 8 * me, None)
            lib.crick_memcpy_d2h(hw.ptr, dw.ptr, 8 * me, None)
            e2e = e2e_leg(hx.array, hw.array, max(1, min(K, 5)),
                          "crick_b200.TDigest.update(numpy over pinned host memory) + centroids()")
            # ordinary (pageable) NumPy arrays: what a caller who never heard of pinned memory has
            px, pw = np.array(hx.array), np.array(hw.array)
            del hx, hw
            e2e_pageable = e2e_leg(px, pw, max(1, min(K, 3)),
                                   "crick_b200.TDigest.update(ordinary numpy arrays) + centroids()")
            del px, pw
        except MemoryError as exc:
            e2e = e2e or {"value": None, "unit": UNIT, "error": str(exc)}
    del dx, dw

    # ---- the other configs of BASELINE.json at full size
    others = None
    if not args.no_others:
        from bench_configs import run_other_configs
        others = run_other_configs(B, cpu=(world == 1 and not args.no_cpu))

    # ---- CPU baseline beside it (rank 0, N == 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        v, kind, dt, _ = cpu_reference_run(args.cpu_sample, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": "%d values of the same lognormal+weights stream, one digest per thread-shard "
                         "then merge (%.1f s)" % (args.cpu_sample, dt)}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config2: single TDigest compression=100 over ONE stream of %d lognormal "
                                   "f64 + f64 weights, parallel k-bucket mode, sharded over %d GPU(s) by "
                                   "contiguous ranges" % (n, world)
                                   + ("; centroid all-gather + one-pass merge of the gathered records every step"
                                      if world > 1 else ""),
                       "values_per_step": n, "values_per_gpu_per_step": m, "compression": COMPRESSION,
                       "l2": "inputs (16 B x %d = %.1f GB per GPU per step) far exceed the 126 MB L2" % (m, 16 * m / 1e9),
                       "update_call": ("TDigest.update(x, w, stream=, tail_stream=): the step's single-CTA tail and the "
                                       "reduction run on a second stream and overlap the next step's sample phase"
                                       if B.pipelined else "TDigest.update(x, w, stream=, check_w='deferred')"),
                       "pipeline": B.pipe_cfg if B.pipelined else None,
                       "weights": "validated on the device in the same pass (deferred verdict, no host sync in a step)",
                       "centroids": nc, "digest_size": total_size},
            "roofline": roofline, "roofline_unweighted": roofline_unweighted, "cpu_baseline": cpu,
            "e2e": e2e, "e2e_pageable": e2e_pageable, "weak": weak, "p2p_exchange": p2p,
            "other_configs": others,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        emit(out)
    if world > 1:
        B.dist.destroy_process_group()


if __name__ == "__main__":
    main()
