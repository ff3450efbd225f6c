// crick_b200/csrc/tdigest_kernels.cu
//
// Reference-compatible kernels (one warp per digest) and the batched query
// kernels.  Replaces, on device:
//   tdigest_update_ndarray loop (tdigest_stubs.c:681-695)      -> k_ingest
//   tdigest_flush (:219-273)                                   -> k_flush
//   tdigest_merge (:592-606)                                   -> k_merge_pairs
//   tdigest_scale (:609-629)                                   -> k_scale
//   tdigest_query_prep (:301-316)                              -> k_query_prep
//   tdigest_quantile / tdigest_quantile_ndarray (:483-589)     -> k_quantile
//   tdigest_cdf / tdigest_cdf_ndarray (:355-480)               -> k_cdf
#include "kernels.h"
#include "tdigest_ref.cuh"

namespace crick {

extern __shared__ __align__(16) unsigned char smem_raw[];

// ---------------------------------------------------------------------------
// k_ingest: group g (one warp) streams its values through digest dmap[g] (or g).
//   offsets != nullptr : group g owns X[offsets[g] .. offsets[g+1])
//   else               : group g owns X[g*row_stride .. +fixed_len)
// flags: bit0 = start from a fresh digest (ignore stored state)
//        bit1 = flush at the end (what centroids() would do, tdigest.pyx:238)
__global__ void __launch_bounds__(256) k_ingest(BatchView v, long ngroups, const double* __restrict__ X,
                                                const double* __restrict__ W, double wsc, long sx,
                                                long sw, const long* __restrict__ offsets,
                                                long fixed_len, long row_stride,
                                                const long* __restrict__ dmap, int flags) {
    const int lane = lane_id();
    const int wpc = blockDim.x >> 5;
    const long g = (long)blockIdx.x * wpc + (threadIdx.x >> 5);
    if (g >= ngroups) return;
    const size_t wsb = warp_ws_bytes(v.cap, v.bufcap);
    WarpWS ws = carve_ws(smem_raw + (size_t)(threadIdx.x >> 5) * wsb, v.cap, v.bufcap);
    const long d = dmap ? dmap[g] : g;
    Hdr h;
    if (flags & 1) fresh_digest(h); else load_digest(v, d, ws, h, lane);
    long beg, len;
    if (offsets) { beg = offsets[g]; len = offsets[g + 1] - beg; }
    else { beg = g * row_stride; len = fixed_len; }
    warp_ingest(ws, v.comp, h, v.cap, v.bufcap, X + beg * sx, sx, W ? W + beg * sw : nullptr, sw,
                wsc, len, lane);
    if (flags & 2) warp_flush(ws, v.comp, h, v.cap, lane);
    store_digest(v, d, ws, h, lane);
}

// k_flush: flush digest first + g*step for g in [0, count)
__global__ void __launch_bounds__(256) k_flush(BatchView v, long first, long step, long count) {
    const int lane = lane_id();
    const int wpc = blockDim.x >> 5;
    const long g = (long)blockIdx.x * wpc + (threadIdx.x >> 5);
    if (g >= count) return;
    const long d = first + g * step;
    if (v.hdr[d].bn == 0) return;
    const size_t wsb = warp_ws_bytes(v.cap, v.bufcap);
    WarpWS ws = carve_ws(smem_raw + (size_t)(threadIdx.x >> 5) * wsb, v.cap, v.bufcap);
    Hdr h;
    load_digest(v, d, ws, h, lane);
    warp_flush(ws, v.comp, h, v.cap, lane);
    store_digest(v, d, ws, h, lane);
}

// k_merge_pairs: for g in [0,count): dst = dfirst + g*dstep (in `dv`),
// src = sfirst + g*sstep (in `sv`, ALREADY FLUSHED by k_flush -- the
// reference flushes `other` first, :595).  tdigest_merge :596-605: skip if
// src.total == 0; feed src centroids in order through tdigest_add; widen
// min/max; dst is NOT flushed afterwards.
__global__ void __launch_bounds__(256) k_merge_pairs(BatchView dv, BatchView sv, long dfirst, long dstep,
                                                     long sfirst, long sstep, long count) {
    const int lane = lane_id();
    const int wpc = blockDim.x >> 5;
    const long g = (long)blockIdx.x * wpc + (threadIdx.x >> 5);
    if (g >= count) return;
    const long d = dfirst + g * dstep, s = sfirst + g * sstep;
    const DevHdr sh = sv.hdr[s];
    if (sh.total == 0.0) return;
    const size_t wsb = warp_ws_bytes(dv.cap, dv.bufcap);
    WarpWS ws = carve_ws(smem_raw + (size_t)(threadIdx.x >> 5) * wsb, dv.cap, dv.bufcap);
    Hdr h;
    load_digest(dv, d, ws, h, lane);
    const double* sc = reinterpret_cast<const double*>(sv.cen + s * (long)sv.cap);
    warp_ingest(ws, dv.comp, h, dv.cap, dv.bufcap, sc, 2, sc + 1, 2, 0.0, (long)sh.n, lane);
    if (h.vmin > sh.vmin) h.vmin = sh.vmin;
    if (h.vmax < sh.vmax) h.vmax = sh.vmax;
    store_digest(dv, d, ws, h, lane);
}

// k_scale: tdigest_scale (:609-629) on an already flushed digest, including the
// reference's compaction quirk (weight moves to slot j, the mean does not).
// Sequential compaction + sequential total (fl order), one thread.
__global__ void k_scale(BatchView v, long d, double factor) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    DevHdr h = v.hdr[d];
    if (h.total == 0.0) return;
    double2* c = v.cen + d * (long)v.cap;
    double total = 0.0;
    int j = 0;
    for (int i = 0; i < h.n; ++i) {
        double w = __dmul_rn(c[i].y, factor);
        if (w > DBL_EPSILON) {
            c[j].y = w;
            total = __dadd_rn(total, w);
            ++j;
        }
    }
    h.total = total;
    h.n = j;
    v.hdr[d] = h;
}

// k_query_prep: flush + table[i] = (mean_i, cumW_before_i + w_i/2) with the
// cumulative weight summed sequentially (:309-315).  One warp per digest.
__global__ void __launch_bounds__(256) k_query_prep(BatchView v, long first, long count) {
    const int lane = lane_id();
    const int wpc = blockDim.x >> 5;
    const long g = (long)blockIdx.x * wpc + (threadIdx.x >> 5);
    if (g >= count) return;
    const long d = first + g;
    const size_t wsb = warp_ws_bytes(v.cap, v.bufcap);
    WarpWS ws = carve_ws(smem_raw + (size_t)(threadIdx.x >> 5) * wsb, v.cap, v.bufcap);
    Hdr h;
    load_digest(v, d, ws, h, lane);
    if (h.bn) {
        warp_flush(ws, v.comp, h, v.cap, lane);
        store_digest(v, d, ws, h, lane);
    }
    double2* t = v.table + d * (long)v.cap;
    // integer weights: exact in any order -> warp scan; otherwise one lane, in order
    bool ints = true;
    for (int i = lane; i < h.n; i += 32) ints = ints && is_int(ws.cen[i].y);
    ints = __all_sync(CRICK_FULL, ints);
    if (ints) {
        double carry = 0.0;
        for (int base = 0; base < h.n; base += 32) {
            int i = base + lane;
            double w = (i < h.n) ? ws.cen[i].y : 0.0;
            double s = w;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
                double o = shfl_up_d(s, dd);
                if (lane >= dd) s = __dadd_rn(s, o);
            }
            s = __dadd_rn(s, carry);
            if (i < h.n) ws.kend[i] = s;
            carry = shfl_d(s, 31);
        }
        ints = carry < 9007199254740992.0;
        __syncwarp();
        if (ints)
            for (int i = lane; i < h.n; i += 32) {
                double w = ws.cen[i].y;
                double before = __dadd_rn(ws.kend[i], -w);  // exact: integers
                t[i] = make_double2(ws.cen[i].x, __dadd_rn(before, __ddiv_rn(w, 2.0)));
            }
    }
    if (!ints && lane == 0) {
        double cum = 0.0;
        for (int i = 0; i < h.n; ++i) {
            double2 c = ws.cen[i];
            t[i] = make_double2(c.x, __dadd_rn(cum, __ddiv_rn(c.y, 2.0)));
            cum = __dadd_rn(cum, c.y);
        }
    }
}

// ---------------------------------------------------------------------------
// Query evaluation on a prepared digest staged in shared memory:
// sm[i] = {mean_i, weight_i, cummid_i}
struct QTab {
    const double* mean;
    const double* weight;
    const double* mid;
    int n;
    double total, vmin, vmax;
};

// tdigest_quantile, tdigest_stubs.c:483-516 (+ bisect_weight :318-327)
__device__ __forceinline__ double quantile1(const QTab& t, double q) {
    if (t.total == 0.0) return __longlong_as_double(0x7ff8000000000000ll);
    if (q <= 0.0) return t.vmin;
    if (q >= 1.0) return t.vmax;
    if (t.n == 1) return t.mean[0];
    const double idx = __dmul_rn(q, t.total);
    int lo = 0, hi = t.n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (t.mid[mid] < idx) lo = mid + 1; else hi = mid;
    }
    double x0, y0, x1, y1;
    if (lo == 0) { x0 = 0.0; y0 = t.vmin; } else { x0 = t.mid[lo - 1]; y0 = t.mean[lo - 1]; }
    if (lo == t.n) { x1 = t.total; y1 = t.vmax; } else { x1 = t.mid[lo]; y1 = t.mean[lo]; }
    // y0 + (index - x0) * (y1 - y0) / (x1 - x0)
    return __dadd_rn(y0, __ddiv_rn(__dmul_rn(__dadd_rn(idx, -x0), __dadd_rn(y1, -y0)),
                                   __dadd_rn(x1, -x0)));
}

// tdigest_cdf, tdigest_stubs.c:355-407 (+ bisect_left_mean :330-339, bisect_right_mean :341-352)
__device__ __forceinline__ double cdf1(const QTab& t, double x) {
    if (t.n == 0) return __longlong_as_double(0x7ff8000000000000ll);
    if (t.n == 1) {
        if (x < t.vmin) return 0.0;
        if (x > t.vmax) return 1.0;
        double r = __dadd_rn(t.vmax, -t.vmin);
        if (r < DBL_EPSILON) return 0.5;
        return __ddiv_rn(__dadd_rn(x, -t.vmin), r);
    }
    if (x >= t.vmax) return 1.0;
    if (x <= t.vmin) return 0.0;
    int lo = 0, hi = t.n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (t.mean[mid] < x) lo = mid + 1; else hi = mid;
    }
    const int il = lo;
    if (x < t.mean[0]) {
        // dw = table[0].weight / 2 where table[0].weight is already w0/2 (:386)
        double x0 = t.vmin, x1 = t.mean[0];
        double dw = __ddiv_rn(t.mid[0], 2.0);
        return __ddiv_rn(__ddiv_rn(__dmul_rn(dw, __dadd_rn(x, -x0)), __dadd_rn(x1, -x0)), t.total);
    } else if (il == t.n) {
        double x0 = t.mean[il - 1], x1 = t.vmax;
        double dw = __ddiv_rn(t.weight[il - 1], 2.0);
        return __dadd_rn(1.0, -__ddiv_rn(__ddiv_rn(__dmul_rn(dw, __dadd_rn(x1, -x)),
                                                   __dadd_rn(x1, -x0)), t.total));
    } else if (t.mean[il] == x) {
        lo = il; hi = t.n;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (x < t.mean[mid]) hi = mid; else lo = mid + 1;
        }
        if (lo && t.mean[lo - 1] == x) lo--;
        return __ddiv_rn(t.mid[lo], t.total);
    } else {
        double x0 = t.mean[il - 1], x1 = t.mean[il];
        double dw = __ddiv_rn(__dadd_rn(t.weight[il - 1], t.weight[il]), 2.0);
        double frac = __ddiv_rn(__dmul_rn(dw, __dadd_rn(x, -x0)), __dadd_rn(x1, -x0));
        return __ddiv_rn(__dadd_rn(t.mid[il - 1], frac), t.total);
    }
}

__device__ __forceinline__ QTab stage_table(const BatchView& v, long d, double* sm, int tid,
                                            int nthreads) {
    const DevHdr h = v.hdr[d];
    const double2* c = v.cen + d * (long)v.cap;
    const double2* t = v.table + d * (long)v.cap;
    double* mean = sm;
    double* weight = sm + v.cap;
    double* mid = sm + 2 * (size_t)v.cap;
    for (int i = tid; i < h.n; i += nthreads) {
        double2 ci = c[i];
        mean[i] = ci.x;
        weight[i] = ci.y;
        mid[i] = t[i].y;
    }
    QTab q;
    q.mean = mean; q.weight = weight; q.mid = mid; q.n = h.n;
    q.total = h.total; q.vmin = h.vmin; q.vmax = h.vmax;
    return q;
}

// One digest, many queries: table staged once per CTA in shared memory, two
// queries per thread per step with 128-bit loads/stores when aligned.
template <bool CDF>
__global__ void __launch_bounds__(256) k_query(BatchView v, long d, const double* __restrict__ in,
                                               double* __restrict__ out, long nq) {
    double* sm = reinterpret_cast<double*>(smem_raw);
    QTab t = stage_table(v, d, sm, threadIdx.x, blockDim.x);
    __syncthreads();
    const long stride = (long)gridDim.x * blockDim.x;
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (vec) {
        const long npair = nq >> 1;
        const double2* in2 = reinterpret_cast<const double2*>(in);
        double2* out2 = reinterpret_cast<double2*>(out);
        for (long i = tid; i < npair; i += stride) {
            double2 a = ldg_stream2(in2 + i);
            double2 r;
            r.x = CDF ? cdf1(t, a.x) : quantile1(t, a.x);
            r.y = CDF ? cdf1(t, a.y) : quantile1(t, a.y);
            out2[i] = r;
        }
        if ((nq & 1) && tid == 0) {
            double a = in[nq - 1];
            out[nq - 1] = CDF ? cdf1(t, a) : quantile1(t, a);
        }
    } else {
        for (long i = tid; i < nq; i += stride) {
            double a = in[i];
            out[i] = CDF ? cdf1(t, a) : quantile1(t, a);
        }
    }
}

// Many digests x the same nq queries (the groupby-quantile shape): one warp per
// digest, out[g*nq + j].  Digests must have been prepared by k_query_prep.
template <bool CDF>
__global__ void __launch_bounds__(256) k_query_grouped(BatchView v, long ngroups,
                                                       const double* __restrict__ in,
                                                       double* __restrict__ out, int nq) {
    const int lane = lane_id();
    const int wpc = blockDim.x >> 5;
    const long g = (long)blockIdx.x * wpc + (threadIdx.x >> 5);
    if (g >= ngroups) return;
    double* sm = reinterpret_cast<double*>(smem_raw) + (size_t)(threadIdx.x >> 5) * 3 * v.cap;
    QTab t = stage_table(v, g, sm, lane, 32);
    __syncwarp();
    for (int j = lane; j < nq; j += 32) {
        double a = in[j];
        out[g * (long)nq + j] = CDF ? cdf1(t, a) : quantile1(t, a);
    }
}


// ---------------------------------------------------------------------------
// small state kernels
__global__ void k_init_hdr(DevHdr* hdr, long n) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    DevHdr d;
    d.vmin = DBL_MAX; d.vmax = -DBL_MAX;  // tdigest_stubs.c:68-69
    d.total = 0.0; d.btotal = 0.0; d.n = 0; d.bn = 0; d.pad0 = 0; d.pad1 = 0;
    hdr[i] = d;
}

// wire record: [n, total, min, max, (mean, weight) x cap], zero padded
__global__ void k_pack_record(BatchView v, long d, double* rec) {
    const DevHdr h = v.hdr[d];
    const double2* c = v.cen + d * (long)v.cap;
    if (threadIdx.x == 0) {
        rec[0] = (double)h.n; rec[1] = h.total; rec[2] = h.vmin; rec[3] = h.vmax;
    }
    for (int i = threadIdx.x; i < v.cap; i += blockDim.x) {
        double2 x = i < h.n ? c[i] : make_double2(0.0, 0.0);
        rec[4 + 2 * i] = x.x;
        rec[5 + 2 * i] = x.y;
    }
}

// T.merge(*records) in record order (tdigest.pyx:310-324 over tdigest_merge :592-606);
// the records are already flushed digests.  One warp.
__global__ void __launch_bounds__(32) k_merge_records(BatchView v, long d, const double* __restrict__ recs,
                                                      int nrec, int skip, int rec_doubles) {
    const int lane = lane_id();
    WarpWS ws = carve_ws(smem_raw, v.cap, v.bufcap);
    Hdr h;
    load_digest(v, d, ws, h, lane);
    for (int r = 0; r < nrec; ++r) {
        if (r == skip) continue;
        const double* rec = recs + (size_t)r * rec_doubles;
        const long n = (long)rec[0];
        if (rec[1] == 0.0) continue;  // :596
        warp_ingest(ws, v.comp, h, v.cap, v.bufcap, rec + 4, 2, rec + 5, 2, 0.0, n, lane);
        if (h.vmin > rec[2]) h.vmin = rec[2];
        if (h.vmax < rec[3]) h.vmax = rec[3];
    }
    store_digest(v, d, ws, h, lane);
}

// digest-to-digest copy (same cap/bufcap), one CTA
__global__ void k_copy_digest(BatchView dv, long dd, BatchView sv, long sd) {
    const DevHdr h = sv.hdr[sd];
    const double2* sc = sv.cen + sd * (long)sv.cap;
    const double2* sb = sv.buf + sd * (long)sv.bufcap;
    double2* dc = dv.cen + dd * (long)dv.cap;
    double2* db = dv.buf + dd * (long)dv.bufcap;
    for (int i = threadIdx.x; i < h.n; i += blockDim.x) dc[i] = sc[i];
    for (int i = threadIdx.x; i < h.bn; i += blockDim.x) db[i] = sb[i];
    if (threadIdx.x == 0) dv.hdr[dd] = h;
}

// out_meta[g*4 + {0,1,2,3}] = n, total, min, max
__global__ void k_group_meta(BatchView v, long ng, double* out) {
    long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ng) return;
    const DevHdr h = v.hdr[g];
    out[4 * g] = (double)h.n; out[4 * g + 1] = h.total; out[4 * g + 2] = h.vmin; out[4 * g + 3] = h.vmax;
}

// ---- all centroids of a group as two dense arrays (crick_tdigest_group_merge_all) -------------
// offsets[g] = number of centroids in the digests before g (offsets[ng] = total); one CTA: every
// thread sums a contiguous slice of the counts, one scan of the 1024 slice sums, slice written.
__global__ void __launch_bounds__(1024) k_group_offsets(BatchView v, long ng, long* offsets,
                                                        double* minmax_total) {
    __shared__ long s_w[32];
    __shared__ double s_mn[32], s_mx[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long per = (ng + 1023) / 1024;
    const long g0 = tid * per, g1 = g0 + per < ng ? g0 + per : ng;
    long run = 0;
    double mn = DBL_MAX, mx = -DBL_MAX;
    for (long g = g0; g < g1; ++g) {
        const DevHdr h = v.hdr[g];
        run += h.n;
        if (h.total != 0.0) { mn = fmin(mn, h.vmin); mx = fmax(mx, h.vmax); }   // empty ones are skipped (:596)
    }
    long inc = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        long o = __shfl_up_sync(CRICK_FULL, inc, d);
        if (lane >= d) inc += o;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { mn = fmin(mn, shfl_xor_d(mn, d)); mx = fmax(mx, shfl_xor_d(mx, d)); }
    if (lane == 31) s_w[wid] = inc;
    if (lane == 0) { s_mn[wid] = mn; s_mx[wid] = mx; }
    __syncthreads();
    long t = s_w[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        long o = __shfl_up_sync(CRICK_FULL, t, d);
        if (lane >= d) t += o;
    }
    long before = (wid > 0 ? __shfl_sync(CRICK_FULL, t, wid - 1) : 0) + inc - run;
    for (long g = g0; g < g1; ++g) { offsets[g] = before; before += v.hdr[g].n; }
    if (tid == 1023) offsets[ng] = __shfl_sync(CRICK_FULL, t, 31);
    if (tid == 0) {
        double a = DBL_MAX, b = -DBL_MAX;
        for (int k = 0; k < 32; ++k) { a = fmin(a, s_mn[k]); b = fmax(b, s_mx[k]); }
        minmax_total[0] = a; minmax_total[1] = b;
    }
}

// one warp per digest: its centroids -> X[off ..), W[off ..)
__global__ void __launch_bounds__(256) k_group_gather(BatchView v, long ng, const long* __restrict__ offsets,
                                                      double* __restrict__ X, double* __restrict__ W) {
    const int lane = lane_id();
    const long g = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= ng) return;
    const long off = offsets[g];
    const int n = (int)(offsets[g + 1] - off);
    const double2* c = v.cen + g * (long)v.cap;
    for (int i = lane; i < n; i += 32) {
        const double2 a = c[i];
        X[off + i] = a.x;
        W[off + i] = a.y;
    }
}

// the merged digest's min / max are those of the inputs' (tdigest_merge :603-604), not of their
// centroid means
__global__ void k_widen_minmax(BatchView v, long d, const double* __restrict__ minmax) {
    if (threadIdx.x || blockIdx.x) return;
    DevHdr h = v.hdr[d];
    if (minmax[0] <= minmax[1]) {
        if (h.vmin > minmax[0]) h.vmin = minmax[0];
        if (h.vmax < minmax[1]) h.vmax = minmax[1];
        v.hdr[d] = h;
    }
}

// ---------------------------------------------------------------------------
// host-side launchers
cudaError_t launch_group_offsets(const BatchView& v, long ng, long* offsets, double* minmax, cudaStream_t st) {
    count_launch();
    k_group_offsets<<<1, 1024, 0, st>>>(v, ng, offsets, minmax);
    return cudaGetLastError();
}
cudaError_t launch_group_gather(const BatchView& v, long ng, const long* offsets, double* X, double* W,
                                cudaStream_t st) {
    count_launch();
    k_group_gather<<<(unsigned)((ng + 7) / 8), 256, 0, st>>>(v, ng, offsets, X, W);
    return cudaGetLastError();
}
cudaError_t launch_widen_minmax(const BatchView& v, long d, const double* minmax, cudaStream_t st) {
    count_launch();
    k_widen_minmax<<<1, 32, 0, st>>>(v, d, minmax);
    return cudaGetLastError();
}

static int pick_wpc(int cap, int bufcap, int maxw, size_t* smem_out) {
    size_t wsb = warp_ws_bytes(cap, bufcap);
    size_t budget = 200 * 1024;
    int wpc = (int)(budget / wsb);
    if (wpc > maxw) wpc = maxw;
    if (wpc < 1) wpc = 1;
    *smem_out = wsb * wpc;
    return wpc;
}

template <typename K>
static cudaError_t allow_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t launch_ingest(const BatchView& v, long ngroups, const double* X, const double* W,
                          double wsc, long sx, long sw, const long* offsets, long fixed_len,
                          long row_stride, const long* dmap, int flags, cudaStream_t st,
                          bool pending_uniform) {
    if (ngroups <= 0) return cudaSuccess;
    if (ingest_lanes_wanted(ngroups))
        return launch_ingest_lanes(v, ngroups, X, W, wsc, sx, sw, offsets, fixed_len, row_stride,
                                   dmap, flags, st, pending_uniform);
    size_t smem;
    // many groups: 8 warps/CTA keeps >=2 CTAs per SM resident at c=100; a single
    // digest gets a 1-warp CTA
    int wpc = pick_wpc(v.cap, v.bufcap, ngroups >= 8 ? 8 : 1, &smem);
    cudaError_t e = allow_smem(k_ingest, smem);
    if (e != cudaSuccess) return e;
    long grid = (ngroups + wpc - 1) / wpc;
    count_launch();
    k_ingest<<<(unsigned)grid, wpc * 32, smem, st>>>(v, ngroups, X, W, wsc, sx, sw, offsets,
                                                    fixed_len, row_stride, dmap, flags);
    return cudaGetLastError();
}

cudaError_t launch_flush(const BatchView& v, long first, long step, long count, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    size_t smem;
    int wpc = pick_wpc(v.cap, v.bufcap, count >= 8 ? 8 : 1, &smem);
    cudaError_t e = allow_smem(k_flush, smem);
    if (e != cudaSuccess) return e;
    long grid = (count + wpc - 1) / wpc;
    count_launch();
    k_flush<<<(unsigned)grid, wpc * 32, smem, st>>>(v, first, step, count);
    return cudaGetLastError();
}

cudaError_t launch_merge_pairs(const BatchView& dv, const BatchView& sv, long dfirst, long dstep,
                               long sfirst, long sstep, long count, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    size_t smem;
    int wpc = pick_wpc(dv.cap, dv.bufcap, count >= 8 ? 8 : 1, &smem);
    cudaError_t e = allow_smem(k_merge_pairs, smem);
    if (e != cudaSuccess) return e;
    long grid = (count + wpc - 1) / wpc;
    count_launch();
    k_merge_pairs<<<(unsigned)grid, wpc * 32, smem, st>>>(dv, sv, dfirst, dstep, sfirst, sstep,
                                                         count);
    return cudaGetLastError();
}

cudaError_t launch_scale(const BatchView& v, long d, double factor, cudaStream_t st) {
    count_launch();
    k_scale<<<1, 32, 0, st>>>(v, d, factor);
    return cudaGetLastError();
}

cudaError_t launch_query_prep(const BatchView& v, long first, long count, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    size_t smem;
    int wpc = pick_wpc(v.cap, v.bufcap, count >= 8 ? 8 : 1, &smem);
    cudaError_t e = allow_smem(k_query_prep, smem);
    if (e != cudaSuccess) return e;
    long grid = (count + wpc - 1) / wpc;
    count_launch();
    k_query_prep<<<(unsigned)grid, wpc * 32, smem, st>>>(v, first, count);
    return cudaGetLastError();
}

cudaError_t launch_query(const BatchView& v, long d, bool cdf, const double* in, double* out,
                         long nq, int sm_count, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    size_t smem = (size_t)3 * v.cap * sizeof(double);
    long want = (nq / 2 + 255) / 256;
    long grid = want < 1 ? 1 : want;
    long maxg = (long)sm_count * 8;
    if (grid > maxg) grid = maxg;
    cudaError_t e;
    if (cdf) {
        e = allow_smem(k_query<true>, smem);
        if (e != cudaSuccess) return e;
        count_launch();
        k_query<true><<<(unsigned)grid, 256, smem, st>>>(v, d, in, out, nq);
    } else {
        e = allow_smem(k_query<false>, smem);
        if (e != cudaSuccess) return e;
        count_launch();
        k_query<false><<<(unsigned)grid, 256, smem, st>>>(v, d, in, out, nq);
    }
    return cudaGetLastError();
}

cudaError_t launch_query_grouped(const BatchView& v, long ngroups, bool cdf, const double* in,
                                 double* out, int nq, cudaStream_t st) {
    if (ngroups <= 0 || nq <= 0) return cudaSuccess;
    size_t per = (size_t)3 * v.cap * sizeof(double);
    int wpc = (int)((200 * 1024) / per);
    if (wpc > 8) wpc = 8;
    if (wpc < 1) wpc = 1;
    size_t smem = per * wpc;
    long grid = (ngroups + wpc - 1) / wpc;
    cudaError_t e;
    if (cdf) {
        e = allow_smem(k_query_grouped<true>, smem);
        if (e != cudaSuccess) return e;
        count_launch();
        k_query_grouped<true><<<(unsigned)grid, wpc * 32, smem, st>>>(v, ngroups, in, out, nq);
    } else {
        e = allow_smem(k_query_grouped<false>, smem);
        if (e != cudaSuccess) return e;
        count_launch();
        k_query_grouped<false><<<(unsigned)grid, wpc * 32, smem, st>>>(v, ngroups, in, out, nq);
    }
    return cudaGetLastError();
}

cudaError_t launch_init_hdr(DevHdr* hdr, long n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    count_launch();
    k_init_hdr<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(hdr, n);
    return cudaGetLastError();
}

cudaError_t launch_pack_record(const BatchView& v, long d, double* rec, cudaStream_t st) {
    count_launch();
    k_pack_record<<<1, 256, 0, st>>>(v, d, rec);
    return cudaGetLastError();
}

cudaError_t launch_merge_records(const BatchView& v, long d, const double* recs, int nrec, int skip,
                                 cudaStream_t st) {
    if (nrec <= 0) return cudaSuccess;
    size_t smem = warp_ws_bytes(v.cap, v.bufcap);
    cudaError_t e = allow_smem(k_merge_records, smem);
    if (e != cudaSuccess) return e;
    count_launch();
    k_merge_records<<<1, 32, smem, st>>>(v, d, recs, nrec, skip, 4 + 2 * v.cap);
    return cudaGetLastError();
}

cudaError_t launch_copy_digest(const BatchView& dv, long dd, const BatchView& sv, long sd,
                               cudaStream_t st) {
    count_launch();
    k_copy_digest<<<1, 256, 0, st>>>(dv, dd, sv, sd);
    return cudaGetLastError();
}

cudaError_t launch_group_meta(const BatchView& v, long ng, double* out, cudaStream_t st) {
    if (ng <= 0) return cudaSuccess;
    count_launch();
    k_group_meta<<<(unsigned)((ng + 255) / 256), 256, 0, st>>>(v, ng, out);
    return cudaGetLastError();
}

}  // namespace crick
