// crick_b200/csrc/tdigest_ref.cuh
//
// Reference-compatible ("bit-exact") t-digest primitives: ONE WARP PER DIGEST.
// Replaces tdigest_add / centroid_sort / integrate / centroid_merge /
// tdigest_flush of the reference (tdigest_stubs.c:111-298).
//
// The flush is split into phases that are parallel across the warp wherever
// the reference's result does not depend on evaluation order, and sequential
// (one lane) where it does:
//   1 stable rank sort of the buffer ('<' on means; ties keep insertion order,
//     -0.0 ties +0.0)                                   [parallel, :111-175]
//   2 two-way merge order, buffer item first only if strictly smaller
//     [parallel: binary searches against the prefix-max of centroid means,
//     which reproduces the two-pointer loop :240-266 even if rounding ever
//     produced a non-monotone centroid array]
//   3 W_{i+1} = fl(W_i + w_i) in merged order           [sequential :251;
//     parallel scan only when every weight is an integer and total < 2^53,
//     where any summation order is exact]
//     kend_i = kscale(c, W_{i+1}/total): one asin per item  [parallel, :194]
//   4 boundary chain: k1 = 0; item i>=1 folds iff kend_i - k1 <= 1, else opens
//     a centroid and k1 = kend_{i-1} (the same fl expression as :213) [sequential]
//   5 per-centroid left-to-right fold, weight += w; mean += (u-mean)*w/weight
//     [parallel across centroids, :205-207]
// All arithmetic uses explicit round-to-nearest intrinsics: no FMA contraction,
// matching the reference build (gcc -O2, no -march).
#pragma once
#include "common.cuh"

namespace crick {

struct WarpWS {
    double2* cen;    // [cap]            live centroids
    double2* buf;    // [bufcap]         insertion-order buffer
    double2* sb;     // [bufcap]         sorted buffer
    double2* items;  // [cap + bufcap]   merged order
    double* kend;    // [cap + bufcap]   prefix-max scratch, then cumulative W, then k
    int* start;      // [cap + bufcap + 1]
};

__host__ __device__ inline size_t warp_ws_bytes(int cap, int bufcap) {
    size_t t = (size_t)cap + (size_t)bufcap;
    size_t b = (size_t)cap * 16 + (size_t)bufcap * 32 + t * 16 + t * 8 + (t + 1) * 4;
    return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ WarpWS carve_ws(unsigned char* base, int cap, int bufcap) {
    WarpWS ws;
    size_t t = (size_t)cap + (size_t)bufcap;
    ws.cen = reinterpret_cast<double2*>(base);
    ws.buf = ws.cen + cap;
    ws.sb = ws.buf + bufcap;
    ws.items = ws.sb + bufcap;
    ws.kend = reinterpret_cast<double*>(ws.items + t);
    ws.start = reinterpret_cast<int*>(ws.kend + t);
    return ws;
}

// Header fields held warp-uniformly in registers while a digest is resident.
struct Hdr {
    double vmin, vmax, total, btotal;
    int n, bn;
};

// k-scale, tdigest_stubs.c:178-189
__device__ __forceinline__ double kscale(double c, double q) {
    q = (q > 1.0) ? 1.0 : q;
    double a = asin(__dadd_rn(__dmul_rn(2.0, q), -1.0));
    return __ddiv_rn(__dmul_rn(c, __dadd_rn(a, CRICK_PI_2)), CRICK_PI);
}

__device__ __forceinline__ bool is_small_int(double w) {
    return w == trunc(w) && w <= 4294967296.0;
}
__device__ __forceinline__ bool is_int(double w) { return w == trunc(w); }

// The two strictly sequential pieces of a flush, run by ONE thread.  The loads do not depend
// on the chain, so they are issued eight at a time and only the arithmetic is serial (the
// one-load-per-step loops cost ~45 cycles per item, 3/4 of it shared-memory latency).
//
// W_{i+1} = fl(W_i + w_i) in merged order (tdigest_stubs.c:251) -> kend[i]
__device__ __forceinline__ void serial_cumsum(const double2* items, double* kend, int T) {
    double W = 0.0;
    int i = 0;
    for (; i + 8 <= T; i += 8) {
        double y[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = items[i + k].y;
#pragma unroll
        for (int k = 0; k < 8; ++k) { W = __dadd_rn(W, y[k]); kend[i + k] = W; }
    }
    for (; i < T; ++i) { W = __dadd_rn(W, items[i].y); kend[i] = W; }
}

// Boundary chain (:192-216 restated): k1 = 0; item 0 opens a centroid;
// item i >= 1 opens one iff !(fl(kend_i - k1) <= 1) (and there is room), and then
// k1 = kend_{i-1}.  Writes start[0..nc] and returns nc.
__device__ __forceinline__ int serial_chain(const double* kend, int* start, int T, int cap) {
    double k1 = 0.0;
    start[0] = 0;
    int nc = 1;
    double kprev = kend[0];
    int i = 1;
    for (; i + 8 <= T; i += 8) {
        double kk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) kk[k] = kend[i + k];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const bool nb = !(__dadd_rn(kk[k], -k1) <= 1.0) && nc < cap;
            if (nb) start[nc] = i + k;
            nc += nb ? 1 : 0;
            k1 = nb ? kprev : k1;
            kprev = kk[k];
        }
    }
    for (; i < T; ++i) {
        const double ki = kend[i];
        const bool nb = !(__dadd_rn(ki, -k1) <= 1.0) && nc < cap;
        if (nb) start[nc] = i;
        nc += nb ? 1 : 0;
        k1 = nb ? kprev : k1;
        kprev = ki;
    }
    start[nc] = T;
    return nc;
}

// Phases 2-5 on `m` sorted new items in ws.sb[0..m) and h.n centroids in ws.cen.
// Requires: ws.sb sorted by '<' (stable), h.total already includes their weight.
__device__ __forceinline__ void warp_merge_sorted(const WarpWS& ws, double comp, Hdr& h,
                                                  int m, int cap, int lane) {
    const int n = h.n;
    const int T = n + m;
    // ---- phase 2a: prefix max of centroid means -> kend[0..n)
    {
        double carry = -INFINITY;
        for (int base = 0; base < n; base += 32) {
            int j = base + lane;
            double v = (j < n) ? ws.cen[j].x : -INFINITY;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                double o = shfl_up_d(v, d);
                if (lane >= d) v = fmax(v, o);
            }
            v = fmax(v, carry);
            if (j < n) ws.kend[j] = v;
            carry = shfl_d(v, 31);
        }
    }
    __syncwarp();
    // ---- phase 2b: merged positions
    for (int i = lane; i < m; i += 32) {
        double2 it = ws.sb[i];
        int lo = 0, hi = n;  // first j with PM[j] > b
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (ws.kend[mid] > it.x) hi = mid; else lo = mid + 1;
        }
        ws.items[i + lo] = it;
    }
    for (int j = lane; j < n; j += 32) {
        double p = ws.kend[j];
        int lo = 0, hi = m;  // number of sorted buffer items strictly below PM[j]
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (ws.sb[mid].x < p) lo = mid + 1; else hi = mid;
        }
        ws.items[j + lo] = ws.cen[j];
    }
    __syncwarp();
    // ---- phase 3: cumulative weights (into kend), then k-scale per item
    bool ints = true;
    for (int i = lane; i < T; i += 32) ints = ints && is_int(ws.items[i].y);
    ints = __all_sync(CRICK_FULL, ints);
    if (ints) {
        double carry = 0.0;
        for (int base = 0; base < T; base += 32) {
            int i = base + lane;
            double v = (i < T) ? ws.items[i].y : 0.0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                double o = shfl_up_d(v, d);
                if (lane >= d) v = __dadd_rn(v, o);
            }
            v = __dadd_rn(v, carry);
            if (i < T) ws.kend[i] = v;
            carry = shfl_d(v, 31);
        }
        // weights are positive, so partial sums are monotone: a final sum below
        // 2^53 proves every partial sum was exact (== the sequential fl chain)
        ints = carry < 9007199254740992.0;
        __syncwarp();
    }
    if (!ints) {
        if (lane == 0) serial_cumsum(ws.items, ws.kend, T);
    }
    __syncwarp();
    for (int i = lane; i < T; i += 32)
        ws.kend[i] = kscale(comp, __ddiv_rn(ws.kend[i], h.total));
    __syncwarp();
    // ---- phase 4: boundary chain
    int nc = 0;
    if (lane == 0) nc = serial_chain(ws.kend, ws.start, T, cap);
    nc = __shfl_sync(CRICK_FULL, nc, 0);
    __syncwarp();
    // ---- phase 5: fold runs
    for (int c = lane; c < nc; c += 32) {
        int s = ws.start[c], e = ws.start[c + 1];
        double2 f = ws.items[s];
        double mean = f.x, wt = f.y;
        for (int i = s + 1; i < e; ++i) {
            double2 it = ws.items[i];
            wt = __dadd_rn(wt, it.y);
            mean = __dadd_rn(mean, __ddiv_rn(__dmul_rn(__dadd_rn(it.x, -mean), it.y), wt));
        }
        ws.cen[c] = make_double2(mean, wt);
    }
    __syncwarp();
    h.n = nc;
}

// tdigest_flush, tdigest_stubs.c:219-273
__device__ __forceinline__ void warp_flush(const WarpWS& ws, double comp, Hdr& h, int cap,
                                           int lane) {
    const int m = h.bn;
    if (m == 0) return;  // :224
    // ---- phase 1: stable rank sort (buffer never holds NaN: tdigest_add drops them)
    for (int i = lane; i < m; i += 32) {
        double xi = ws.buf[i].x;
        int r = 0;
        for (int j = 0; j < m; ++j) {
            double xj = ws.buf[j].x;
            r += ((xj < xi) || (j < i && !(xi < xj))) ? 1 : 0;
        }
        ws.sb[r] = ws.buf[i];
    }
    __syncwarp();
    double s0 = ws.sb[0].x, s1 = ws.sb[m - 1].x;  // :229-232
    if (h.vmin > s0) h.vmin = s0;
    if (h.vmax < s1) h.vmax = s1;
    h.total = __dadd_rn(h.total, h.btotal);  // :237
    h.btotal = 0.0;
    warp_merge_sorted(ws, comp, h, m, cap, lane);
    h.bn = 0;
}

// Feed `len` (x, w) pairs through the buffer exactly like a loop of
// tdigest_add (tdigest_stubs.c:276-298, driver loop :681-695):
//   - drop !isfinite(x) or w <= DBL_EPSILON
//   - flush when the buffer is full BEFORE appending the next accepted value
//   - buffer_total_weight += w in insertion order
// X/W are strided (stride in doubles); W == nullptr means scalar weight wsc.
__device__ __forceinline__ void warp_ingest(const WarpWS& ws, double comp, Hdr& h, int cap,
                                            int bufcap, const double* __restrict__ X, long sx,
                                            const double* __restrict__ W, long sw, double wsc,
                                            long len, int lane) {
    const unsigned lt = (1u << lane) - 1u;
    for (long pos = 0; pos < len; pos += 32) {
        long i = pos + lane;
        bool valid = i < len;
        double x = valid ? X[i * sx] : 0.0;
        double w = (W != nullptr) ? (valid ? W[i * sw] : 0.0) : wsc;
        bool accept = valid && is_finite_d(x) && !(w <= DBL_EPSILON);
        unsigned mask = __ballot_sync(CRICK_FULL, accept);
        int k = __popc(mask);
        if (k == 0) continue;
        int rank = __popc(mask & lt);
        bool ints = __all_sync(CRICK_FULL, !accept || is_small_int(w));
        int done = 0;
        while (done < k) {
            if (h.bn == bufcap) warp_flush(ws, comp, h, cap, lane);
            int room = bufcap - h.bn;
            int take = min(room, k - done);
            bool mine = accept && rank >= done && rank < done + take;
            if (mine) ws.buf[h.bn + rank - done] = make_double2(x, w);
            // buffer_total_weight, insertion order (:297)
            if (ints && h.btotal == trunc(h.btotal) && h.btotal < 4503599627370496.0) {
                double v = mine ? w : 0.0;  // exact in any order: integers below 2^53
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v = __dadd_rn(v, shfl_xor_d(v, d));
                h.btotal = __dadd_rn(h.btotal, v);
            } else if (W == nullptr) {
                double b = h.btotal;
                for (int t = 0; t < take; ++t) b = __dadd_rn(b, wsc);
                h.btotal = b;
            } else {
                __syncwarp();
                double b = h.btotal;
                if (lane == 0)
                    for (int t = 0; t < take; ++t) b = __dadd_rn(b, ws.buf[h.bn + t].y);
                h.btotal = shfl_d(b, 0);
            }
            h.bn += take;
            done += take;
            __syncwarp();
        }
    }
}

__device__ __forceinline__ void load_digest(const BatchView& v, long g, const WarpWS& ws, Hdr& h,
                                            int lane) {
    const DevHdr d = v.hdr[g];
    h.vmin = d.vmin; h.vmax = d.vmax; h.total = d.total; h.btotal = d.btotal;
    h.n = d.n; h.bn = d.bn;
    const double2* c = v.cen + g * (long)v.cap;
    const double2* b = v.buf + g * (long)v.bufcap;
    for (int i = lane; i < h.n; i += 32) ws.cen[i] = c[i];
    for (int i = lane; i < h.bn; i += 32) ws.buf[i] = b[i];
    __syncwarp();
}

__device__ __forceinline__ void store_digest(const BatchView& v, long g, const WarpWS& ws,
                                             const Hdr& h, int lane) {
    __syncwarp();
    double2* c = v.cen + g * (long)v.cap;
    double2* b = v.buf + g * (long)v.bufcap;
    for (int i = lane; i < h.n; i += 32) c[i] = ws.cen[i];
    for (int i = lane; i < h.bn; i += 32) b[i] = ws.buf[i];
    if (lane == 0) {
        DevHdr d;
        d.vmin = h.vmin; d.vmax = h.vmax; d.total = h.total; d.btotal = h.btotal;
        d.n = h.n; d.bn = h.bn; d.pad0 = 0; d.pad1 = 0;
        v.hdr[g] = d;
    }
}

__device__ __forceinline__ void fresh_digest(Hdr& h) {
    h.vmin = DBL_MAX; h.vmax = -DBL_MAX; h.total = 0.0; h.btotal = 0.0; h.n = 0; h.bn = 0;
}

}  // namespace crick
