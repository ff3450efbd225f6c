#!/usr/bin/env python
"""bench.py -- headline benchmark of the crick hot path on B200 (contract: see DESIGN.md 6).

Metric (BASELINE.json): TDigest.update Gvalues/s over 1e9 float64 values with float64 weights
(lognormal x, U(0.5,1.5) w, compression 100) -- BASELINE.json configs[1] -- per GPU, weak
scaling: every rank owns a 1e9-value shard of one logical stream; with N > 1 each step ends with
the all-gather of the (3.2 KB) centroid records and the on-device merge in rank order.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n VALUES]

One JSON line on rank 0.  `value` = device-resident throughput (inputs in HBM before the timed
region), `e2e` = the same through TDigest.update() on pinned HOST arrays (H2D inside the timed
region, centroids read back every step), `roofline` = the hot kernel (k_bin_accumulate) against
the measured HBM peak, `cpu_baseline` = the reference's own C code on this box's host cores.
"""
import argparse
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
    ap.add_argument("--n", type=int, default=N_DEFAULT, help="values per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=2 * 10**8)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
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


def main():
    # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version
    # banner under torchrun, for one) goes to stderr instead
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import crick_b200 as cb
    from crick_b200._lib import lib
    from crick_b200.distributed import allgather_merge_tdigest

    torch.cuda.set_device(local_rank)
    lib.crick_set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.n
    K, W = args.steps, args.warmup
    stream = torch.cuda.Stream()
    sptr = stream.cuda_stream

    # ---- synthetic, device-resident shard: element i of the stream depends on (seed, i) only
    dx, dw = cb.DeviceArray((n,)), cb.DeviceArray((n,))
    lib.crick_synth_fill(dx.ptr, dw.ptr, n, SEED, rank * n, 0, None)
    lib.crick_device_synchronize()

    reducer = None
    if world > 1:
        from crick_b200.distributed import TDigestAllReduce
        reducer = TDigestAllReduce(COMPRESSION, exact=False)

    def step(t):
        t.update(dx, dw, mode="parallel", stream=sptr)
        if world > 1:
            with torch.cuda.stream(stream):
                return reducer(t)
        return t

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    digest = cb.TDigest(COMPRESSION, mode="parallel")
    merged = None
    # clocks are sampled every 100 ms from before the warm-up (the same load) to the end of
    # the timed region, which by itself may be shorter than one sample period
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(W):
        merged = step(digest)
    barrier()
    launches0 = cb.launch_count()
    lib.crick_profile_enable(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(K):
        merged = step(digest)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    lib.crick_profile_enable(0)
    import ctypes
    kms, kn = ctypes.c_double(0), ctypes.c_int64(0)
    lib.crick_profile_collect(ctypes.addressof(kms), ctypes.addressof(kn))
    launches = cb.launch_count() - launches0
    # nvidia-smi needs ~0.1-0.3 s to print its first line and the timed region may be shorter:
    # keep the identical load running (untimed) until a few samples exist
    if world > 1:   # the same count on every rank (a step holds a collective)
        tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    n_extra = int(min(400, max(0, 600.0 / max(ms / K, 1e-3))))
    for _ in range(n_extra):
        step(digest)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    clocks["note"] = "sampled every 100 ms from the warm-up through the timed region and an identical untimed tail"
    value = world * n * K / (ms * 1e-3) / 1e9
    nc = len(merged.centroids())
    total_size = merged.size()

    # ---- roofline of the dominant kernel (k_bin_accumulate), CUDA events around each launch
    peak, peak_src = peak_hbm()
    k_ms = kms.value / max(kn.value, 1)
    achieved = (ALGO_BYTES_PER_VALUE * n) / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_bin_accumulate", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "kernel_ms_per_launch": k_ms, "kernel_share_of_step": (kms.value / ms) if ms else None,
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_VALUE * n,
                # SURVEY.md 8d asks for both denominators: the box-measured copy bandwidth above
                # and the nominal 8 TB/s of HBM3e
                "peak_nominal": 8000.0, "frac_nominal": achieved / 8000.0}
    try:  # DRAM bytes per launch from the committed ncu --set full capture, scaled to n
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        roofline["traffic"] = tj["k_bin_accumulate_dram_bytes_per_value"] * n
        roofline["traffic_source"] = tj["source"]
    except Exception:
        pass

    # ---- e2e: public API on pinned HOST arrays; H2D every step, centroids read back every step
    e2e = None
    if not args.no_e2e:
        import psutil
        need = 16 * n
        avail = psutil.virtual_memory().available
        ne = n
        while 16 * ne * (world if world > 1 else 1) * 2 > avail and ne > 10**7:
            ne //= 2
        try:
            hx, hw = cb.PinnedArray(ne), cb.PinnedArray(ne)
            lib.crick_memcpy_d2h(hx.ptr, dx.ptr, 8 * ne, None)
            lib.crick_memcpy_d2h(hw.ptr, dw.ptr, 8 * ne, None)
            ke = max(1, min(K, 5))
            def e2e_step():
                t = cb.TDigest(COMPRESSION, mode="parallel")
                t.update(hx.array, hw.array)
                if world > 1:
                    t = allgather_merge_tdigest(t)
                return t.centroids()
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(ke):
                cen = e2e_step()
            barrier()
            dt = time.perf_counter() - t0
            if world > 1:
                tdt = torch.tensor([dt], dtype=torch.float64, device="cuda")
                dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
                dt = float(tdt.item())
            e2e = {"value": world * ne * ke / dt / 1e9, "unit": UNIT,
                   "h2d_bytes_per_step": 16 * ne, "d2h_bytes_per_step": int(cen.nbytes),
                   "values_per_step_per_gpu": ne, "steps": ke,
                   "api": "crick_b200.TDigest.update(numpy over pinned host memory) + centroids()"}
            del hx, hw
        except MemoryError as exc:
            e2e = {"value": None, "unit": UNIT, "error": str(exc)}

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
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config2: single TDigest compression=100 over 1e9 lognormal f64 + "
                                   "f64 weights per GPU, parallel k-bucket mode"
                                   + ("; centroid all-gather + one-pass merge of the gathered records every step" if world > 1 else ""),
                       "values_per_gpu_per_step": n, "compression": COMPRESSION,
                       "l2": "inputs (16 B x %d = %.1f GB per step) far exceed the 126 MB L2" % (n, 16 * n / 1e9),
                       "centroids": nc, "digest_size": total_size},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks,
        }
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
