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
//
// The boundary decision without an asin per item (CHEAP).  centroid_merge (:192-216) opens a new
// centroid when fl(kend - k1) > 1 with kend = c (asin(2 q - 1) + pi/2) / pi, q = Wc / total and
// k1 the same expression at q1 = W1 / total (W1: cumulative weight before the open centroid's
// first item; k1 = 0 = K(q1 = 0) for the first one).  With K real-valued and increasing,
//     K(q) - K(q1) > 1  <=>  q > qlim(q1) = q1 cos(pi/c) + sin^2(pi/2c) + sqrt(q1 (1 - q1)) sin(pi/c)
// (the sine addition theorem on asin(2 q1 - 1) + pi/c; all terms >= 0, no cancellation), valid while
// q1 < (1 + cos(pi/c)) / 2 -- beyond that K(q1) + 1 > c and no centroid can open.  The rounded chain
// the reference evaluates differs from K by: q (1 ulp), 2q-1 (2^-53 absolute), asin (<= 2 ulp of
// pi/2), three roundings of ~c, the subtraction: together < 2e-15 in units of the asin argument, and
// qlim itself is computed to ~1e-15 -- plus, for a boundary deep in a tail, the quantisation of
// y1 = fl(2 q1 - 1) next to -+1 as seen through asin: <= 1.4e-17 / sqrt(q1 (1 - q1)), which the window
// adds (x3.6) once that root is below 3.2e-4 (tests/test_window_decision.py attacks the window on the
// CPU with the reference's arithmetic).  So ONCE PER CENTROID qlim -> [Wlo, Whi] = total (qlim -+ margin)
// with margin = 6e-13 (>= 100x the bound), and per item only  Wc > Whi -> open,  Wc < Wlo -> fold;
// a Wc inside the window evaluates the reference's own two kscale() and decides from them (counted
// in g_exact_fallbacks).  Means and weights never depend on k, so the result is the same bits.
#include "api_common.h"
#include "kernels.h"
#include "tdigest_ref.cuh"

#include <cstdlib>

namespace crick {

extern __shared__ __align__(16) unsigned char smem_raw[];

// BUFW: the buffer keeps a weight per item (weights given, or items of unknown weight may be
// waiting in the buffers); otherwise every buffered item weighs wsc.
__device__ unsigned long long g_exact_fallbacks = 0;

template <bool BUFW, bool CHEAP>
__global__ void __launch_bounds__(512, 1) k_ingest_lanes(BatchView v, long ngroups,
                                                        const double* __restrict__ X,
                                                        const double* __restrict__ W, double wsc,
                                                        long sx, long sw,
                                                        const long* __restrict__ offsets,
                                                        long fixed_len, long row_stride,
                                                        const long* __restrict__ dmap, int flags,
                                                        double2* __restrict__ scratch, double margin) {
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int wpc = blockDim.x >> 5;
    const long gwarp = (long)blockIdx.x * wpc + wib;
    const long nwarps = (long)gridDim.x * wpc;
    const int cap = v.cap, bufcap = v.bufcap;
    const double comp = v.comp;
    // qlim(q1) coefficients; comp < 2 would put pi/c beyond the quarter period: exact path only
    const double kC = cos(CRICK_PI / comp), kS = sin(CRICK_PI / comp);
    const double kH = sin(CRICK_PI_2 / comp) * sin(CRICK_PI_2 / comp);
    const double kQtop = 0.5 * (1.0 + kC);
    const bool cheap_ok = CHEAP && comp >= 2.0;
    unsigned fallbacks = 0;
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
        // (four slots at a time while the fourth-nearest is still > x -- the buffer is sorted, so
        // then all four move: independent loads, one compare per four shifts; singles after that)
        auto insert = [&](double x, double w) {
            int i = bn;
            while (i >= 4) {
                const double p0 = bx[(i - 1) * 32], p1 = bx[(i - 2) * 32];
                const double p2 = bx[(i - 3) * 32], p3 = bx[(i - 4) * 32];
                if (!(p3 > x)) break;
                bx[i * 32] = p0; bx[(i - 1) * 32] = p1; bx[(i - 2) * 32] = p2; bx[(i - 3) * 32] = p3;
                if (BUFW) {
                    const double q0 = bwt[(i - 1) * 32], q1 = bwt[(i - 2) * 32];
                    const double q2 = bwt[(i - 3) * 32], q3 = bwt[(i - 4) * 32];
                    bwt[i * 32] = q0; bwt[(i - 1) * 32] = q1; bwt[(i - 2) * 32] = q2; bwt[(i - 3) * 32] = q3;
                }
                i -= 4;
            }
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
        // the next value is already in a register while this one is inserted, and the line eight
        // values ahead is on its way to L1 (a lane walks its own row: 32 sectors per warp load)
        double xn = 0.0, wn = wsc;
        if (len > 0) { xn = xp[0]; if (wp) wn = wp[0]; }
        for (;;) {
            // tdigest_add (:276-298) until an accepted value finds the buffer full
            bool pending = false;
            double px = 0.0, pw = 0.0;
            while (j < len) {
                const double x = xn;
                const double w = wn;
                ++j;
                if (j < len) { xn = xp[j * sx]; if (wp) wn = wp[j * sw]; }
                if ((j & 3) == 0 && j + 8 < len) {
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(xp + (j + 8) * sx));
                    if (wp) asm volatile("prefetch.global.L1 [%0];" ::"l"(wp + (j + 8) * sw));
                }
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
                // CHEAP: decision window of the open centroid, and what the exact path needs
                double Wlo = -INFINITY, Whi = INFINITY, W1 = 0.0;
                bool first_c = true;
                const bool windowed = cheap_ok && total > 1e-290 && total < 1e290;
                const double rtotal = windowed ? 1.0 / total : 0.0;
                auto set_window = [&](double q1) {
                    if (!windowed) return;                 // every item takes the exact path
                    q1 = fmin(q1, 1.0);
                    const double r = sqrt(q1 * (1.0 - q1));
                    const double qlim = q1 * kC + kH + r * kS;
                    // k1 is evaluated by the reference at y1 = fl(2 q1 - 1), quantised to 2^-53 next
                    // to -1: in the asin argument that is an error of 2^-54 / cos(theta1) with
                    // cos(theta1) = 2 r, i.e. <= 1.4e-17 / r in qlim -- beyond the margin once
                    // q1 < ~5e-10 (a boundary behind 1e-9 of the total weight).  Widen the window there.
                    double m_ = margin;
                    if (r < 3.2e-4 && r > 0.0) m_ += 5e-17 / r;
                    Wlo = total * (qlim - m_);
                    Whi = q1 > kQtop - margin ? INFINITY : total * (qlim + m_);
                    if (q1 > kQtop - margin) Wlo = total * (1.0 - 16.0 * margin - 1e-11);   // no centroid can open beyond
                };
                if (CHEAP) set_window(0.0);
                double bxv = bx[0], bwv = BUFW ? bwt[0] : wsc;
                // live centroids ic and ic + 1 in registers, ic + 4 on its way to L1
                double2 cv = n > 0 ? A[0] : make_double2(0.0, 0.0);
                double2 cn = n > 1 ? A[32] : make_double2(0.0, 0.0);
                if (n > 4) asm volatile("prefetch.global.L1 [%0];" ::"l"(A + 4 * 32));
                const int T = n + m;
                for (int t = 0; t < T; ++t) {
                    // two-way merge, buffer item first only if strictly smaller (:240-266)
                    double u, w;
                    if (ib < m && (ic >= n || bxv < cv.x)) {
                        u = bxv; w = bwv;
                        if (++ib < m) { bxv = bx[ib * 32]; if (BUFW) bwv = bwt[ib * 32]; }
                    } else {
                        u = cv.x; w = cv.y;
                        cv = cn;
                        if (ic + 2 < n) cn = A[(ic + 2) * 32];
                        if (ic + 5 < n) asm volatile("prefetch.global.L1 [%0];" ::"l"(A + (ic + 5) * 32));
                        ++ic;
                    }
                    const double Wprev = Wc;
                    Wc = __dadd_rn(Wc, w);                                 // :251
                    bool open;
                    double kend = 0.0;
                    if (CHEAP) {
                        if (Wc > Whi) open = true;
                        else if (Wc < Wlo) open = false;
                        else {
                            kend = kscale(comp, __ddiv_rn(Wc, total));     // :194, as the reference
                            k1 = first_c ? 0.0 : kscale(comp, __ddiv_rn(W1, total));
                            open = !(__dadd_rn(kend, -k1) <= 1.0);
                            if (t > 0) ++fallbacks;
                        }
                    } else {
                        kend = kscale(comp, __ddiv_rn(Wc, total));         // :194
                        open = !(__dadd_rn(kend, -k1) <= 1.0);
                    }
                    if (t == 0) {
                        cm = u; cw = w; nc = 1;                            // :197-202
                    } else if (open && nc < cap) {
                        Bq[(nc - 1) * 32] = make_double2(cm, cw);          // :208-214
                        cm = u; cw = w; ++nc;
                        if (CHEAP) { W1 = Wprev; first_c = false; set_window(Wprev * rtotal); }
                        else k1 = kprev;
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
    if (CHEAP && fallbacks) atomicAdd(&g_exact_fallbacks, (unsigned long long)fallbacks);
}

// Groups from which the thread-per-digest kernel is used (below it the GPU is not filled:
// 32 digests per warp).  crick_group_kernel() / CRICK_GROUP_LANES=0|1 force the choice.
static int lanes_sm_count() {
    const int n = sm_count();        // current device (api.cu caches it per device)
    return n > 0 ? n : 148;
}

static int g_lanes_mode = [] {
    const char* env = getenv("CRICK_GROUP_LANES");
    return (env && (env[0] == '0' || env[0] == '1')) ? env[0] - '0' : -1;
}();

// Decision window half-width in q (see the file comment); 0 < margin; a huge value (>= 1) sends
// every item through the exact kscale pair, negative selects the kernel without the window.
static double g_lanes_margin = [] {
    const char* env = getenv("CRICK_GROUP_MARGIN");
    return env ? atof(env) : 6e-13;
}();
double ingest_lanes_margin(double m) {
    double prev = g_lanes_margin;
    if (m == m) g_lanes_margin = m;
    return prev;
}
cudaError_t ingest_lanes_fallbacks(unsigned long long* out, bool reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_exact_fallbacks, sizeof(*out));
    if (e == cudaSuccess && reset) {
        const unsigned long long z = 0;
        e = cudaMemcpyToSymbol(g_exact_fallbacks, &z, sizeof(z));
    }
    return e;
}
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
    const double margin = g_lanes_margin;
    auto go = [&](auto kern) {
        cudaError_t e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e1 != cudaSuccess) return e1;
        count_launch();
        kern<<<(unsigned)grid, wpc * 32, smem, st>>>(v, ngroups, X, W, wsc, sx, sw, offsets, fixed_len,
                                                     row_stride, dmap, flags, scratch, margin);
        return cudaGetLastError();
    };
    if (margin < 0.0) e = bufw ? go(k_ingest_lanes<true, false>) : go(k_ingest_lanes<false, false>);
    else e = bufw ? go(k_ingest_lanes<true, true>) : go(k_ingest_lanes<false, true>);
    cudaError_t e2 = cudaFreeAsync(scratch, st);
    return e != cudaSuccess ? e : e2;
}

}  // namespace crick
