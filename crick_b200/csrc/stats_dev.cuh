// crick_b200/csrc/stats_dev.cuh -- device pieces of the parallel SummaryStats path shared by
// stats.cu (stand-alone pass) and tdigest_parallel.cu (moments fused into the t-digest pass).
#pragma once
#include "common.cuh"

namespace crick {

struct StatsPart {          // mergeable summary of a contiguous index range
    long long count;
    double sum, vmin, vmax, m2, m3, m4;
    long first_idx;          // index of first non-NaN element (LONG_MAX if none)
    long last_nan_idx;       // index of last NaN element (-1 if none)
    double first_value;      // value at first_idx
    double last_value;       // value of the last element seen (for the all-NaN case)
    long last_idx;
};

__device__ __forceinline__ void part_fresh(StatsPart& p) {
    p.count = 0; p.sum = 0; p.vmin = INFINITY; p.vmax = -INFINITY; p.m2 = p.m3 = p.m4 = 0;
    p.first_idx = 0x7fffffffffffffffl; p.last_nan_idx = -1; p.first_value = 0.0;
    p.last_value = 0.0; p.last_idx = -1;
}

__device__ __forceinline__ void part_merge(StatsPart& a, const StatsPart& b) {
    if (b.first_idx < a.first_idx) { a.first_idx = b.first_idx; a.first_value = b.first_value; }
    if (b.last_nan_idx > a.last_nan_idx) a.last_nan_idx = b.last_nan_idx;
    if (b.last_idx > a.last_idx) { a.last_idx = b.last_idx; a.last_value = b.last_value; }
    if (b.count == 0) return;
    if (a.count == 0) {
        a.count = b.count; a.sum = b.sum; a.vmin = b.vmin; a.vmax = b.vmax;
        a.m2 = b.m2; a.m3 = b.m3; a.m4 = b.m4;
        return;
    }
    // Pebay pairwise merge, fp64 products
    double n1 = (double)a.count, n2 = (double)b.count, n = n1 + n2;
    double delta = b.sum / n2 - a.sum / n1;
    double dn = delta / n, dn2 = dn * dn, dn3 = dn2 * dn;
    double n1n2 = n1 * n2;
    double m4 = a.m4 + b.m4 + n1n2 * (n1 * n1 - n1n2 + n2 * n2) * delta * dn3 +
                6.0 * (n1 * n1 * b.m2 + n2 * n2 * a.m2) * dn2 + 4.0 * (n1 * b.m3 - n2 * a.m3) * dn;
    double m3 = a.m3 + b.m3 + n1n2 * (n1 - n2) * delta * dn2 + 3.0 * (n1 * b.m2 - n2 * a.m2) * dn;
    double m2 = a.m2 + b.m2 + n1n2 * delta * dn;
    a.m4 = m4; a.m3 = m3; a.m2 = m2;
    a.count += b.count; a.sum += b.sum;
    a.vmin = fmin(a.vmin, b.vmin); a.vmax = fmax(a.vmax, b.vmax);
}

__device__ __forceinline__ StatsPart part_shfl_xor(const StatsPart& p, int m) {
    StatsPart o;
    o.count = __shfl_xor_sync(CRICK_FULL, p.count, m);
    o.sum = shfl_xor_d(p.sum, m); o.vmin = shfl_xor_d(p.vmin, m); o.vmax = shfl_xor_d(p.vmax, m);
    o.m2 = shfl_xor_d(p.m2, m); o.m3 = shfl_xor_d(p.m3, m); o.m4 = shfl_xor_d(p.m4, m);
    o.first_idx = __shfl_xor_sync(CRICK_FULL, p.first_idx, m);
    o.last_nan_idx = __shfl_xor_sync(CRICK_FULL, p.last_nan_idx, m);
    o.first_value = shfl_xor_d(p.first_value, m);
    o.last_value = shfl_xor_d(p.last_value, m);
    o.last_idx = __shfl_xor_sync(CRICK_FULL, p.last_idx, m);
    return o;
}

// One element into a thread's shifted power sums.  UNIT: every count is 1 (no division).
// Bookkeeping for the homogeneous / first_value quirks (:200-205) only runs on the rare
// events (first value of the thread, a NaN).
struct StatsAcc {
    double K, cN, s1, s2, s3, s4, sx, vmin, vmax;
    long long cnt;
    bool haveK;
};

template <bool UNIT>
__device__ __forceinline__ void stats_one(StatsAcc& a, StatsPart& p, double v, long long c, long i) {
    if (v != v) {               // NaN: skipped by stats_add (:93), but remembered
        if (i > p.last_nan_idx) p.last_nan_idx = i;
        if (i > p.last_idx) { p.last_idx = i; p.last_value = v; }
        return;
    }
    const double cd = UNIT ? 1.0 : (double)c;
    const double u = UNIT ? v : v / cd;   // stats_do_update: u2 = sum2 / n2 (:51)
    if (!a.haveK) {
        a.K = u; a.haveK = true;
        if (i < p.first_idx) { p.first_idx = i; p.first_value = v; }
    }
    const double d = u - a.K, d2 = d * d;
    if (UNIT) { a.s1 += d; a.s2 += d2; a.s3 = __fma_rn(d2, d, a.s3); a.s4 = __fma_rn(d2, d2, a.s4); }
    else {
        a.s1 = __fma_rn(cd, d, a.s1); a.s2 = __fma_rn(cd, d2, a.s2);
        a.s3 = __fma_rn(cd * d2, d, a.s3); a.s4 = __fma_rn(cd * d2, d2, a.s4);
    }
    a.cN += cd; a.cnt += UNIT ? 1 : c; a.sx += v;
    a.vmin = v < a.vmin ? v : a.vmin;
    a.vmax = v > a.vmax ? v : a.vmax;
}


// N elements with count 1 at once.  When the thread already has its shift K and none of the N
// values is a NaN (the common case) nothing but the power sums, the sum and min / max is touched:
// 13 instructions per value instead of ~40 with the per-value bookkeeping of stats_one.
// idx(k) gives the global index of v[k] (only evaluated on the slow path).
template <int N, typename IdxFn>
__device__ __forceinline__ void stats_block_unit(StatsAcc& a, StatsPart& p, const double* v, IdxFn idx) {
    bool slow = !a.haveK;
#pragma unroll
    for (int k = 0; k < N; ++k) slow |= (v[k] != v[k]);
    if (slow) {
#pragma unroll
        for (int k = 0; k < N; ++k) stats_one<true>(a, p, v[k], 1, idx(k));
        return;
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double d = v[k] - a.K, d2 = d * d;
        a.s1 += d; a.s2 += d2; a.s3 = __fma_rn(d2, d, a.s3); a.s4 = __fma_rn(d2, d2, a.s4);
        a.sx += v[k];
        a.vmin = v[k] < a.vmin ? v[k] : a.vmin;
        a.vmax = v[k] > a.vmax ? v[k] : a.vmax;
    }
    a.cN += (double)N;
    a.cnt += N;
}

// The same with the fewest instructions per value (the moments fused into SpaceSaving's counting
// pass, which is bound by its issue rate): min / max are only CHECKED per value (two chained
// compares) and updated on the slow path, NaN is only looked for when the values can be NaN
// (float64 items; an int64 -> double conversion never is), and the sum is not kept per value:
// sum = cN * K + s1 (stats_thread_finish2).  Fast path: 6 fp64 instructions + 2 (+1) compares.
template <int N, bool CHECK_NAN, typename IdxFn>
__device__ __forceinline__ void stats_block_unit2(StatsAcc& a, StatsPart& p, const double* v, IdxFn idx) {
    bool slow = !a.haveK;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        if (CHECK_NAN) slow = slow || (v[k] != v[k]);
        slow = slow || (v[k] < a.vmin) || (v[k] > a.vmax);
    }
    if (slow) {
#pragma unroll
        for (int k = 0; k < N; ++k) stats_one<true>(a, p, v[k], 1, idx(k));
        return;
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double d = v[k] - a.K, d2 = d * d;
        a.s1 += d; a.s2 += d2; a.s3 = __fma_rn(d2, d, a.s3); a.s4 = __fma_rn(d2, d2, a.s4);
    }
    a.cN += (double)N;
    a.cnt += N;
}
// (stats_one keeps a.sx as well; the fast path above does not: rebuild the sum from the shift)
__device__ __forceinline__ void stats_thread_finish2(StatsAcc& a, StatsPart& p);

// per-thread shifted power sums -> central moments of the thread's elements
__device__ __forceinline__ void stats_thread_finish(const StatsAcc& a, StatsPart& p) {
    if (a.haveK) {
        const double md = a.s1 / a.cN;
        p.count = a.cnt; p.sum = a.sx; p.vmin = a.vmin; p.vmax = a.vmax;
        p.m2 = a.s2 - a.cN * md * md;
        p.m3 = a.s3 - 3.0 * md * a.s2 + 2.0 * a.cN * md * md * md;
        p.m4 = a.s4 - 4.0 * md * a.s3 + 6.0 * md * md * a.s2 - 3.0 * a.cN * md * md * md * md;
        if (p.m2 < 0.0) p.m2 = 0.0;
        if (p.m4 < 0.0) p.m4 = 0.0;
        if (p.last_idx < 0) p.last_idx = 0;   // "the batch was not empty"
    }
}

__device__ __forceinline__ void stats_thread_finish2(StatsAcc& a, StatsPart& p) {
    if (a.haveK) a.sx = __fma_rn(a.cN, a.K, a.s1);      // sum of v = sum of (K + d)
    stats_thread_finish(a, p);
}

// xor-shuffle tree with a fixed association (lower lane on the left) => deterministic
__device__ __forceinline__ StatsPart stats_warp_reduce(StatsPart p) {
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
        StatsPart o = part_shfl_xor(p, m);
        if ((threadIdx.x & m) == 0) part_merge(p, o);
        else { StatsPart t = o; part_merge(t, p); p = t; }
    }
    return p;
}

__device__ __forceinline__ void stats_acc_fresh(StatsAcc& a) {
    a.K = 0.0; a.cN = 0.0; a.s1 = a.s2 = a.s3 = a.s4 = a.sx = 0.0;
    a.vmin = INFINITY; a.vmax = -INFINITY; a.cnt = 0; a.haveK = false;
}

}  // namespace crick
