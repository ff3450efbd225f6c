// crick_b200/csrc/tdigest_lanes.cu
//
// Reference-compatible grouped ingest with ONE THREAD PER DIGEST, for many independent digests
// (the dask groupby-quantile shape, BASELINE config 3: 1e6 digests x 1000 values).
//
// k_ingest (tdigest_kernels.cu) gives a digest to a warp and runs the strictly sequential parts
// of a flush (cumulative weights, boundary chain) on one lane: ~240 warp instructions per
// value, issue bound.  With >= tens of thousands of digests there is enough parallelism ACROSS
// digests, so here every thread runs the reference's own sequential algorithm on its digest
// (tdigest_add :276-298, tdigest_flush :219-273 with centroid_merge :192-216 in the restated
// one-asin-per-item form of SURVEY.md 8a) and a warp instruction advances 32 digests at once:
//   * the buffer of a digest lives in shared memory, interleaved over the lanes (slot i of lane
//     l at [i * 32 + l]: conflict-free), and is kept sorted by insertion as values arrive --
//     stable, '<' on means, the order centroid_sort (:116-175) produces;
//   * the centroid arrays of the 32 digests of a warp live in a global scratch, interleaved the
//     same way (a warp access is 512 contiguous bytes); a flush is ONE streaming two-pointer
//     merge that reads the old array once and writes the new one once;
//   * state is converted from / to the AoS layout of BatchView once per kernel.
// Same arithmetic (explicit round-to-nearest intrinsics, no FMA contraction) and the same
// decisions in the same order as warp_flush, so results are bit-identical to k_ingest and to
// the reference (tests/test_gpu_tdigest.py grouped cases run both).
#include "kernels.h"
#include "tdigest_ref.cuh"

#include <cstdlib>

namespace crick {

extern __shared__ __align__(16) unsigned char smem_raw[];

// BUFW: the buffer keeps a weight per item (weights given, or items of unknown weight may be
// waiting in the buffers); otherwise every buffered item weighs wsc.
template <bool BUFW>
__global__ void __launch_bounds__(512, 1) k_ingest_lanes(BatchView v, long ngroups,
                                                        const double* __restrict__ X,
                                                        const double* __restrict__ W, double wsc,
                                                        long sx, long sw,
                                                        const long* __restrict__ offsets,
                                                        long fixed_len, long row_stride,
                                                        const long* __restrict__ dmap, int flags,
                                                        double2* __restrict__ scratch) {
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int wpc = blockDim.x >> 5;
    const long gwarp = (long)blockIdx.x * wpc + wib;
    const long nwarps = (long)gridDim.x * wpc;
    const int cap = v.cap, bufcap = v.bufcap;
    const double comp = v.comp;
    double2* A = scratch + (size_t)gwarp * 2 * cap * 32 + lane;   // live centroids, [i * 32]
    double2* Bq = A + (size_t)cap * 32;                            // output of the next flush
    double* bx = reinterpret_cast<double*>(smem_raw) +
                 (size_t)wib * bufcap * 32 * (BUFW ? 2 : 1) + lane;   // sorted buffer means
    double* bwt = bx + (size_t)bufcap * 32;                             // and weights (BUFW)

    for (long g0 = gwarp * 32; g0 < ngroups; g0 += nwarps * 32) {
        const long g = g0 + lane;
        const bool live = g < ngroups;
        const long d = live ? (dmap ? dmap[g] : g) : 0;
        double vmin = DBL_MAX, vmax = -DBL_MAX, total = 0.0, btotal = 0.0;
        int n = 0, bn = 0;
        // stable insertion into the sorted buffer: after every element <= x
        auto insert = [&](double x, double w) {
            int i = bn;
            while (i > 0) {
                const double p = bx[(i - 1) * 32];
                if (!(p > x)) break;
                bx[i * 32] = p;
                if (BUFW) bwt[i * 32] = bwt[(i - 1) * 32];
                --i;
            }
            bx[i * 32] = x;
            if (BUFW) bwt[i * 32] = w;
            ++bn;
        };
        if (live && !(flags & 1)) {
            const DevHdr hd = v.hdr[d];
            vmin = hd.vmin; vmax = hd.vmax; total = hd.total; btotal = hd.btotal; n = hd.n;
            const double2* c = v.cen + d * (long)cap;
            for (int i = 0; i < n; ++i) A[i * 32] = c[i];
            const double2* b = v.buf + d * (long)bufcap;
            for (int i = 0; i < hd.bn; ++i) { const double2 it = b[i]; insert(it.x, it.y); }
        }
        long beg = 0, len = 0;
        if (live) {
            if (offsets) { beg = offsets[g]; len = offsets[g + 1] - beg; }
            else { beg = g * row_stride; len = fixed_len; }
        }
        const double* xp = X + beg * sx;
        const double* wp = W ? W + beg * sw : nullptr;
        long j = 0;
        for (;;) {
            // tdigest_add (:276-298) until an accepted value finds the buffer full
            bool pending = false;
            double px = 0.0, pw = 0.0;
            while (j < len) {
                const double x = xp[j * sx];
                const double w = wp ? wp[j * sw] : wsc;
                ++j;
                if (!(is_finite_d(x) && !(w <= DBL_EPSILON))) continue;   // :285
                if (bn == bufcap) { pending = true; px = x; pw = w; break; }   // :288-290
                insert(x, w);
                btotal = __dadd_rn(btotal, w);                           // :297
            }
            if (!pending && !((flags & 2) && bn > 0)) break;
            // ---- tdigest_flush (:219-273), bn >= 1
            {
                const int m = bn;
                const double s0 = bx[0], s1 = bx[(m - 1) * 32];          // :229-232
                if (vmin > s0) vmin = s0;
                if (vmax < s1) vmax = s1;
                total = __dadd_rn(total, btotal);                         // :237
                btotal = 0.0;
                int ib = 0, ic = 0, nc = 0;
                double Wc = 0.0, k1 = 0.0, kprev = 0.0, cm = 0.0, cw = 0.0;
                double bxv = bx[0], bwv = BUFW ? bwt[0] : wsc;
                double2 cv = n > 0 ? A[0] : make_double2(0.0, 0.0);
                const int T = n + m;
                for (int t = 0; t < T; ++t) {
                    // two-way merge, buffer item first only if strictly smaller (:240-266)
                    double u, w;
                    if (ib < m && (ic >= n || bxv < cv.x)) {
                        u = bxv; w = bwv;
                        if (++ib < m) { bxv = bx[ib * 32]; if (BUFW) bwv = bwt[ib * 32]; }
                    } else {
                        u = cv.x; w = cv.y;
                        if (++ic < n) cv = A[ic * 32];
                    }
                    Wc = __dadd_rn(Wc, w);                                 // :251
                    const double kend = kscale(comp, __ddiv_rn(Wc, total)); // :194
                    if (t == 0) {
                        cm = u; cw = w; nc = 1;                            // :197-202
                    } else if (!(__dadd_rn(kend, -k1) <= 1.0) && nc < cap) {
                        Bq[(nc - 1) * 32] = make_double2(cm, cw);          // :208-214
                        cm = u; cw = w; ++nc;
                        k1 = kprev;
                    } else {
                        cw = __dadd_rn(cw, w);                             // :203-207
                        cm = __dadd_rn(cm, __ddiv_rn(__dmul_rn(__dadd_rn(u, -cm), w), cw));
                    }
                    kprev = kend;
                }
                Bq[(nc - 1) * 32] = make_double2(cm, cw);
                n = nc;
                bn = 0;
                double2* tmp = A; A = Bq; Bq = tmp;                        // :270-272
            }
            if (!pending) break;                                           // that was the final flush
            insert(px, pw);
            btotal = __dadd_rn(btotal, pw);
        }
        if (live) {
            double2* c = v.cen + d * (long)cap;
            for (int i = 0; i < n; ++i) c[i] = A[i * 32];
            double2* b = v.buf + d * (long)bufcap;
            for (int i = 0; i < bn; ++i) b[i] = make_double2(bx[i * 32], BUFW ? bwt[i * 32] : wsc);
            DevHdr hd;
            hd.vmin = vmin; hd.vmax = vmax; hd.total = total; hd.btotal = btotal;
            hd.n = n; hd.bn = bn; hd.pad0 = 0; hd.pad1 = 0;
            v.hdr[d] = hd;
        }
    }
}

// Groups from which the thread-per-digest kernel is used (below it the GPU is not filled:
// 32 digests per warp).  crick_group_kernel() / CRICK_GROUP_LANES=0|1 force the choice.
static int lanes_sm_count() {
    static int sms = [] {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
        return n;
    }();
    return sms;
}

static int g_lanes_mode = [] {
    const char* env = getenv("CRICK_GROUP_LANES");
    return (env && (env[0] == '0' || env[0] == '1')) ? env[0] - '0' : -1;
}();
int ingest_lanes_mode(int mode) {
    int prev = g_lanes_mode;
    if (mode >= -1 && mode <= 1) g_lanes_mode = mode;
    return prev;
}
bool ingest_lanes_wanted(long ngroups) {
    if (g_lanes_mode >= 0) return g_lanes_mode == 1;
    return ngroups >= 32L * 4 * lanes_sm_count();
}

cudaError_t launch_ingest_lanes(const BatchView& v, long ngroups, const double* X, const double* W,
                                double wsc, long sx, long sw, const long* offsets, long fixed_len,
                                long row_stride, const long* dmap, int flags, cudaStream_t st,
                                bool pending_uniform) {
    if (ngroups <= 0) return cudaSuccess;
    const int sm_count = lanes_sm_count();
    const bool bufw = W != nullptr || !(pending_uniform || (flags & 1));
    const size_t per_warp = (size_t)v.bufcap * 32 * 8 * (bufw ? 2 : 1);
    int wpc_max = (int)((200 * 1024) / per_warp);
    if (wpc_max > 16) wpc_max = 16;
    if (wpc_max < 1) return cudaErrorInvalidValue;
    // spread the warps over all SMs before stacking them: a digest is one long dependent chain
    const long want_warps = (ngroups + 31) / 32;
    long wpc_l = (want_warps + sm_count - 1) / sm_count;
    int wpc = (int)(wpc_l > wpc_max ? wpc_max : wpc_l);
    long grid = (want_warps + wpc - 1) / wpc;
    if (grid > sm_count) grid = sm_count;
    const size_t smem = per_warp * wpc;
    const size_t scratch_bytes = (size_t)grid * wpc * 2 * v.cap * 32 * sizeof(double2);
    double2* scratch = nullptr;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&scratch), scratch_bytes, st);
    if (e != cudaSuccess) return e;
    if (bufw) {
        e = cudaFuncSetAttribute(k_ingest_lanes<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) {
            count_launch();
            k_ingest_lanes<true><<<(unsigned)grid, wpc * 32, smem, st>>>(
                v, ngroups, X, W, wsc, sx, sw, offsets, fixed_len, row_stride, dmap, flags, scratch);
            e = cudaGetLastError();
        }
    } else {
        e = cudaFuncSetAttribute(k_ingest_lanes<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) {
            count_launch();
            k_ingest_lanes<false><<<(unsigned)grid, wpc * 32, smem, st>>>(
                v, ngroups, X, W, wsc, sx, sw, offsets, fixed_len, row_stride, dmap, flags, scratch);
            e = cudaGetLastError();
        }
    }
    cudaError_t e2 = cudaFreeAsync(scratch, st);
    return e != cudaSuccess ? e : e2;
}

}  // namespace crick
