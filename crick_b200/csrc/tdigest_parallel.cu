// crick_b200/csrc/tdigest_parallel.cu
//
// Parallel "k-bucket" update for one huge stream (SURVEY.md 2.3 K1/K3/K5, 7 option (c)).
//
// The reference folds values into centroids 42 at a time through a strictly
// sequential chain (tdigest_stubs.c:219-298); 1e9 values would be 23.8 M
// dependent flushes.  A centroid, however, is only (sum w, sum w*x) over a
// contiguous quantile range, and both sums are order independent.  So:
//
//   1. k_sample        strided+hashed sample of S (x, w) pairs            (reads S*16 B)
//   2. k_bitonic_*     sort the sample by x (S = 16 Ki or 64 Ki)
//   3. k_make_edges    weighted sample quantiles at the k-scale grid
//                      q_j = sin^2(pi*j/(2*ov*c)), j = 1..ov*c-1  (ov sub-buckets
//                      per unit of k, so NO asin per element), heavy duplicates get a
//                      bucket of their own [v, nextafter(v)); plus a 4096-cell lookup
//                      table over the order-preserving bit pattern of x
//   4. k_bin_accumulate  THE HOT KERNEL: every value is read exactly once from HBM
//                      (128-bit streaming loads), filtered like tdigest_add (:285),
//                      mapped to its bucket (one table probe + 0-2 edge compares) and
//                      accumulated into warp-private (sum w, sum w*x) arrays in shared
//                      memory -- no atomics: lanes of a warp that hit the same bucket
//                      are serialised with __match_any_sync.
//   5. k_finalize      bucket sums -> <= ov*c weighted points in sorted order, which are
//                      merged into the digest with the reference's own greedy k-scale
//                      merge (warp_merge_sorted == tdigest_flush phases 2-5).
//
// Result: the digest the reference would produce had it been handed the
// pre-aggregated bucket points in one flush.  Bit parity with the sequential
// reference is impossible by construction (different clustering); parity here
// is (a) bucket sums vs a CPU restatement of the same partition at 1e-12 and
// (b) rank error <= the reference's on the same data (tests/test_parallel.py).
#include "kernels.h"
#include "tdigest_ref.cuh"

#include <cstdlib>
#include <mutex>
#include <utility>
#include <vector>

namespace crick {

extern __shared__ __align__(16) unsigned char smem_raw[];

constexpr int kNC = 4096;        // lookup-table cells
constexpr int kEMax = 1024;      // max edges (buckets - 1)
constexpr int kBPad = kEMax + 8; // row stride of the per-CTA partial sums
constexpr long kSMax = 65536;    // max sample size
constexpr int kSortChunk = 4096; // elements sorted per CTA in shared memory

struct ParHeader {
    int E;       // number of edges; buckets = E + 1, bucket b = [edge[b-1], edge[b])
    int shift;   // cell = ((key - ubase) >> shift) + 1
    unsigned long long ubase;
    long S;      // sample length (power of two)
    int ncta;    // CTAs that wrote partials
    int grid;    // lookup grid: 0 = linear in the ordered bit pattern, 1 = linear in x
    double sample_total;
    double x0;   // grid 1: cell = (int)((x - x0) * inv) + 1
    double inv;
};

struct ParLayout {
    ParHeader* hdr;
    double* edges;       // [kEMax]
    unsigned* tab;       // [kNC]   lo | hi << 16 (the grid k_make_edges chose)
    unsigned* tab2;      // [kNC]   scratch: candidate table of the other grid
    double* sx;          // [kSMax]
    double* sw;          // [kSMax]
    double* cw;          // [kSMax]
    double2* partial;    // [ncta_max * kBPad]
    double2* minmax;     // [ncta_max]
    unsigned char* fin;  // finalize workspace (global fallback)
    size_t fin_bytes;
};

// Sub-buckets per unit of k.  2 is enough: on 1e7 lognormal / normal / gamma values the rank
// error with ov = 2, 3, 4 is the same within noise and at or below the reference's
// (DESIGN.md 4), and the per-warp tables of the hot kernel shrink with it.
static int ov_for(double comp) {
    int c = (int)ceil(comp);
    int ov = kEMax / c;
    if (ov > 2) ov = 2;
    if (ov < 1) ov = 1;
    return ov;
}
static int bmax_for(double comp) { return ov_for(comp) * (int)ceil(comp); }
static int mmax_for(double comp) { return 2 * bmax_for(comp) + 2 * (int)ceil(comp) + 8; }

static size_t fin_ws_bytes(int cap, int mmax) {
    // cen[cap] + sb[mmax] + items[cap+mmax] + kend + start
    size_t t = (size_t)cap + mmax;
    size_t b = (size_t)cap * 16 + (size_t)mmax * 16 + t * 16 + t * 8 + (t + 1) * 4;
    return (b + 15) & ~(size_t)15;
}

size_t parallel_scratch_bytes(double comp, int sm_count) {
    int cap = 2 * (int)ceil(comp);
    size_t b = 256;
    b += sizeof(double) * kEMax + sizeof(unsigned) * kNC;
    b += 3 * sizeof(double) * kSMax;
    b += sizeof(unsigned) * kNC;
    b += sizeof(double2) * (size_t)(2 * sm_count) * kBPad;
    b += sizeof(double2) * (size_t)(2 * sm_count);
    b += fin_ws_bytes(cap, mmax_for(comp)) + 64;
    return b;
}

static ParLayout carve(void* scratch, double comp, int sm_count) {
    ParLayout L;
    unsigned char* p = reinterpret_cast<unsigned char*>(scratch);
    L.hdr = reinterpret_cast<ParHeader*>(p); p += 256;
    L.edges = reinterpret_cast<double*>(p); p += sizeof(double) * kEMax;
    L.tab = reinterpret_cast<unsigned*>(p); p += sizeof(unsigned) * kNC;
    L.sx = reinterpret_cast<double*>(p); p += sizeof(double) * kSMax;
    L.sw = reinterpret_cast<double*>(p); p += sizeof(double) * kSMax;
    L.cw = reinterpret_cast<double*>(p); p += sizeof(double) * kSMax;
    L.tab2 = reinterpret_cast<unsigned*>(p); p += sizeof(unsigned) * kNC;
    L.partial = reinterpret_cast<double2*>(p); p += sizeof(double2) * (size_t)(2 * sm_count) * kBPad;
    L.minmax = reinterpret_cast<double2*>(p); p += sizeof(double2) * (size_t)(2 * sm_count);
    L.fin = p;
    L.fin_bytes = fin_ws_bytes(2 * (int)ceil(comp), mmax_for(comp));
    return L;
}

double* parallel_sample_x(void* scratch) {
    return reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(scratch) + 256 +
                                     sizeof(double) * kEMax + sizeof(unsigned) * kNC);
}
double* parallel_sample_w(void* scratch) { return parallel_sample_x(scratch) + kSMax; }
long parallel_sample_capacity(void*) { return kSMax; }

// next representable double above / below a finite x (zeros are canonical +0.0)
__device__ __forceinline__ double next_up(double x) {
    if (x == 0.0) return __longlong_as_double(1ll);
    long long b = __double_as_longlong(x);
    return __longlong_as_double(x > 0.0 ? b + 1 : b - 1);
}
__device__ __forceinline__ double next_down(double x) {
    if (x == 0.0) return __longlong_as_double((long long)0x8000000000000001ull);
    long long b = __double_as_longlong(x);
    return __longlong_as_double(x > 0.0 ? b - 1 : b + 1);
}

// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// 1. sample: element i comes from segment i of n/S, at a hashed offset (so a
// periodic input cannot alias with the stride).  Rejected values (tdigest_add's
// filter, :285) become (+inf, 0) and sort to the end.
__global__ void k_sample(const double* __restrict__ X, const double* __restrict__ W, double wsc,
                         long n, long S, double* __restrict__ sx, double* __restrict__ sw,
                         ParHeader* hdr) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) hdr->S = S;
    if (i >= S) return;
    long seg = n / S;
    double x = INFINITY, w = 0.0;
    if (seg > 0) {
        long pos = i * seg + (long)(mix64((unsigned long long)i) % (unsigned long long)seg);
        double xv = X[pos];
        double wv = W ? W[pos] : wsc;
        if (is_finite_d(xv) && !(wv <= DBL_EPSILON)) { x = xv + 0.0; w = wv; }
    } else if (i < n) {
        double xv = X[i];
        double wv = W ? W[i] : wsc;
        if (is_finite_d(xv) && !(wv <= DBL_EPSILON)) { x = xv + 0.0; w = wv; }
    }
    sx[i] = x;
    sw[i] = w;
}

// 2. bitonic sort of (key, value) pairs, ascending by key.
__device__ __forceinline__ void ce(double& ka, double& va, double& kb, double& vb, bool asc) {
    bool sw = asc ? (kb < ka) : (ka < kb);
    if (sw) { double t = ka; ka = kb; kb = t; t = va; va = vb; vb = t; }
}

// all stages with stride < kSortChunk for the given k range, inside shared memory
// mode 0: full local sort (k = 2 .. kSortChunk); mode 1: merge tail of level K (j = kSortChunk/2 .. 1)
__global__ void __launch_bounds__(512) k_bitonic_local(double* __restrict__ keys,
                                                       double* __restrict__ vals, long K, int mode) {
    double* sk = reinterpret_cast<double*>(smem_raw);
    double* sv = sk + kSortChunk;
    const long base = (long)blockIdx.x * kSortChunk;
    for (int i = threadIdx.x; i < kSortChunk; i += blockDim.x) {
        sk[i] = keys[base + i];
        sv[i] = vals[base + i];
    }
    __syncthreads();
    long k0 = mode == 0 ? 2 : K;
    long k1 = mode == 0 ? kSortChunk : K;
    for (long k = k0; k <= k1; k <<= 1) {
        for (int j = (int)((k >> 1) < kSortChunk ? (k >> 1) : (kSortChunk >> 1)); j > 0; j >>= 1) {
            for (int p = threadIdx.x; p < kSortChunk / 2; p += blockDim.x) {
                int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                int q = i | j;
                bool asc = (((base + i) & k) == 0);
                double ka = sk[i], va = sv[i], kb = sk[q], vb = sv[q];
                ce(ka, va, kb, vb, asc);
                sk[i] = ka; sv[i] = va; sk[q] = kb; sv[q] = vb;
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < kSortChunk; i += blockDim.x) {
        keys[base + i] = sk[i];
        vals[base + i] = sv[i];
    }
}

__global__ void k_bitonic_global(double* __restrict__ keys, double* __restrict__ vals, long L,
                                 long j, long k) {
    long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= L / 2) return;
    long i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
    long q = i | j;
    bool asc = ((i & k) == 0);
    double ka = keys[i], va = vals[i], kb = keys[q], vb = vals[q];
    bool sw = asc ? (kb < ka) : (ka < kb);
    if (sw) { keys[i] = kb; vals[i] = vb; keys[q] = ka; vals[q] = va; }
}

static cudaError_t sort_sample(double* keys, double* vals, long L, cudaStream_t st) {
    size_t smem = (size_t)kSortChunk * 16;
    cudaError_t e = cudaFuncSetAttribute(k_bitonic_local, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    unsigned nchunk = (unsigned)(L / kSortChunk);
    count_launch();
    k_bitonic_local<<<nchunk, 512, smem, st>>>(keys, vals, 0, 0);
    for (long k = 2L * kSortChunk; k <= L; k <<= 1) {
        for (long j = k >> 1; j >= kSortChunk; j >>= 1) {
            count_launch();
            k_bitonic_global<<<(unsigned)((L / 2 + 255) / 256), 256, 0, st>>>(keys, vals, L, j, k);
        }
        count_launch();
        k_bitonic_local<<<nchunk, 512, smem, st>>>(keys, vals, k, 1);
    }
    return cudaGetLastError();
}

// 3. edges + lookup table.  Single CTA, 1024 threads.
__device__ __forceinline__ int cell_of(unsigned long long key, unsigned long long ubase, int shift) {
    if (key < ubase) return 0;
    unsigned long long c = ((key - ubase) >> shift) + 1ull;
    return c > (unsigned long long)(kNC - 1) ? (kNC - 1) : (int)c;
}

__device__ __forceinline__ int cell_lin(double x, double x0, double inv) {
    if (x < x0) return 0;
    // cvt.rzi.s32.f64 saturates (NaN -> 0): any finite or infinite x gives a valid cell
    int c = __double2int_rz(__dmul_rn(__dadd_rn(x, -x0), inv));
    return (c > kNC - 2 ? kNC - 2 : c) + 1;
}

__global__ void __launch_bounds__(1024) k_make_edges(ParHeader* hdr, const double* __restrict__ sx,
                                                     const double* __restrict__ sw,
                                                     double* __restrict__ cw, double* __restrict__ edges,
                                                     unsigned* __restrict__ tab,
                                                     unsigned* __restrict__ tab2, double comp, int ov) {
    __shared__ double s_part[1024];
    __shared__ double s_raw[kEMax];
    __shared__ int s_cnt[kEMax];
    __shared__ int s_scan[1024];
    __shared__ int s_E;
    const int tid = threadIdx.x;
    const long S = hdr->S;
    const long per = S / 1024 > 0 ? S / 1024 : 1;
    // a. inclusive scan of sample weights (fixed order => deterministic)
    {
        long lo = (long)tid * per, hi = lo + per;
        if (hi > S) hi = S;
        double s = 0.0;
        for (long i = lo; i < hi; ++i) s += sw[i];
        s_part[tid] = (lo < S) ? s : 0.0;
        __syncthreads();
        if (tid == 0) {
            double run = 0.0;
            for (int t = 0; t < 1024; ++t) { double v = s_part[t]; s_part[t] = run; run += v; }
        }
        __syncthreads();
        double run = s_part[tid];
        for (long i = lo; i < hi; ++i) { run += sw[i]; cw[i] = run; }
    }
    __syncthreads();
    const double Wt = cw[S - 1];
    const int c = (int)ceil(comp);
    const int E0 = ov * c - 1;
    // b. raw edges at the k-scale grid
    for (int j = tid; j < E0; j += blockDim.x) {
        double kq = (double)(j + 1) / (double)ov;                 // k in (0, c)
        double sn = sin(CRICK_PI_2 * kq / comp);
        double q = sn * sn;                                        // k^-1(kq)
        if (q > 1.0) q = 1.0;
        double target = q * Wt;
        long lo = 0, hi = S - 1;                                   // first i with cw[i] >= target
        while (lo < hi) {
            long mid = (lo + hi) >> 1;
            if (cw[mid] < target) lo = mid + 1; else hi = mid;
        }
        double e = sx[lo];
        s_raw[j] = e;
    }
    __syncthreads();
    // c. dedupe; a value chosen more than once is "heavy" and gets [v, nextafter(v))
    for (int j = tid; j < E0; j += blockDim.x) {
        double v = s_raw[j];
        int cnt = 0;
        bool first = (j == 0) || (s_raw[j - 1] != v);
        if (first && is_finite_d(v) && Wt > 0.0) {
            cnt = 1;
            if (j + 1 < E0 && s_raw[j + 1] == v) {
                int jn = j + 1;
                while (jn < E0 && s_raw[jn] == v) ++jn;
                double nx = next_up(v);
                if (is_finite_d(nx) && !(jn < E0 && s_raw[jn] == nx)) cnt = 2;
            }
        }
        s_cnt[j] = cnt;
    }
    __syncthreads();
    // exclusive scan of s_cnt (E0 <= 1023): thread t owns element t
    {
        int v = (tid < E0) ? s_cnt[tid] : 0;
        s_scan[tid] = v;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            int o = (tid >= d) ? s_scan[tid - d] : 0;
            __syncthreads();
            s_scan[tid] += o;
            __syncthreads();
        }
        int excl = s_scan[tid] - v;
        if (tid < E0 && v > 0) {
            double e = s_raw[tid];
            edges[excl] = e;
            if (v == 2) edges[excl + 1] = next_up(e);
        }
        if (tid == 1023) s_E = s_scan[1023];
    }
    __syncthreads();
    const int E = s_E;
    // d. lookup tables.  Two candidate grids of kNC cells: linear in the order-preserving bit
    // pattern (~ log x: good for heavy-tailed positive data) and linear in x (good for data
    // that crosses zero, where the bit pattern spends its cells on the denormal range).
    // cost = sum over cells of m * ceil(log2(m + 1)), m = edges in the cell, ~ expected
    // compares per value; the cheaper grid goes to `tab`.
    __shared__ unsigned long long s_cost[2];
    unsigned long long ubase = 0;
    int shift = 0;
    double x0 = 0.0, inv = 0.0;
    bool lin_ok = false;
    if (E > 0) {
        ubase = ordered_key(edges[0]);
        unsigned long long span = ordered_key(edges[E - 1]) - ubase;
        while (shift < 63 && ((span >> shift) + 1ull) > (unsigned long long)(kNC - 2)) ++shift;
        x0 = edges[0];
        double width = edges[E - 1] - x0;
        inv = (double)(kNC - 2) / width;
        lin_ok = E >= 2 && width > 0.0 && is_finite_d(width) && is_finite_d(inv);
    }
    if (tid < 2) s_cost[tid] = 0ull;
    __syncthreads();
    for (int g = 0; g < 2; ++g) {
        unsigned* out = g == 0 ? tab : tab2;
        if (g == 1 && !lin_ok) break;
        unsigned long long local = 0;
        for (int cell = tid; cell < kNC; cell += blockDim.x) {
            // lo = #edges with cell(e) < cell ; hi = #edges with cell(e) <= cell
            int a = 0, b = E;
            while (a < b) {
                int mid = (a + b) >> 1;
                int ce_ = g == 0 ? cell_of(ordered_key(edges[mid]), ubase, shift)
                                 : cell_lin(edges[mid], x0, inv);
                if (ce_ < cell) a = mid + 1; else b = mid;
            }
            int lo = a;
            b = E;
            while (a < b) {
                int mid = (a + b) >> 1;
                int ce_ = g == 0 ? cell_of(ordered_key(edges[mid]), ubase, shift)
                                 : cell_lin(edges[mid], x0, inv);
                if (ce_ <= cell) a = mid + 1; else b = mid;
            }
            out[cell] = (unsigned)lo | ((unsigned)a << 16);
            int m = a - lo, lg = 0;
            while ((1 << lg) < m + 1) ++lg;
            local += (unsigned long long)m * (unsigned long long)lg;
        }
        atomicAdd(&s_cost[g], local);
    }
    __syncthreads();
    const int grid = (lin_ok && s_cost[1] < s_cost[0]) ? 1 : 0;
    if (grid == 1)
        for (int cell = tid; cell < kNC; cell += blockDim.x) tab[cell] = tab2[cell];
    if (tid == 0) {
        hdr->E = E;
        hdr->shift = shift;
        hdr->ubase = ubase;
        hdr->sample_total = Wt;
        hdr->ncta = 0;
        hdr->grid = grid;
        hdr->x0 = x0;
        hdr->inv = inv;
    }
}

// 4. THE HOT KERNEL -----------------------------------------------------------
//
// Weighted histogram of (sum w, sum w*x) over B <= bmax buckets, one pass over HBM.
// round-1 profile (profiles/r1_a_*): resolving intra-warp bucket collisions with
// __match_any_sync saturates the ADU pipe (92 %) at ~0.5 value/clk/SM.  This version needs
// neither MATCH nor atomics:
//   * every warp owns a private table acc[bucket][C] of double2 in shared memory,
//     C = 8 (or 4/2/1 for large compressions) COLUMNS per bucket;
//   * lane l only ever touches column l % C, and the 32/C lanes that share a column take
//     turns in P = 32/C passes separated by __syncwarp().  Inside a pass the C active lanes
//     form (part of) one quarter-warp and hit C distinct 16-byte slots: every LDS.128/STS.128
//     is a single conflict-free wavefront;
//   * a lane's U = 4 values of one group are de-duplicated in registers first (equal buckets
//     are folded, the later one is redirected to a dummy row), so the 4 read-modify-writes of
//     a pass are independent and overlap.
// Bucket lookup: one probe of a 4096-cell table over a grid that is linear either in x or in
// the order-preserving bit pattern of x (k_make_edges picks the cheaper one), then 0-2
// compares against the edges that fall into the cell.
struct Lookup {
    const unsigned* tab;   // shared
    const double* edges;   // shared
    unsigned long long ubase;
    double x0, inv;
    int shift;
    int E;
};

template <int GRID>
__device__ __forceinline__ int bucket_of(const Lookup& L, double x) {
    int cell;
    if (GRID == 1) cell = cell_lin(x, L.x0, L.inv);
    else cell = cell_of(ordered_key(x + 0.0), L.ubase, L.shift);  // -0.0 -> +0.0: zeros tie
    unsigned t = L.tab[cell];
    int lo = (int)(t & 0xffffu), hi = (int)(t >> 16);
    while (lo < hi) {  // #edges <= x, restricted to the cell's candidates
        int mid = (lo + hi) >> 1;
        if (L.edges[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

constexpr int kU = 4;  // values per lane per group

// One group: lane holds kU (x, w) pairs.  acc_col = this warp's table + (lane % C).
template <bool HAS_W, int C, int GRID>
__device__ __forceinline__ void accumulate_group(const Lookup& L, double2* acc_col, int dummy,
                                                 const double* xs, const double* ws, double wsc,
                                                 unsigned valid_mask, int pass, double& vmin,
                                                 double& vmax) {
    constexpr int P = 32 / C;
    int idx[kU];
    double ww[kU], ss[kU];
    const int top = L.E - 1;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
        const double x = xs[u];
        const double w = HAS_W ? ws[u] : wsc;
        bool ok = is_finite_d(x) && !(w <= DBL_EPSILON) && ((valid_mask >> u) & 1u);  // :285
        int b = bucket_of<GRID>(L, x);
        // the minimum lies in bucket 0 or 1 and the maximum in bucket E-1 or E (edges are data
        // values, or nextafter of one): only those values need the fp64 min/max
        if (ok && (b <= 1 || b >= top)) {
            vmin = x < vmin ? x : vmin;
            vmax = x > vmax ? x : vmax;
        }
        // rejected values go to the dummy row (never read back), whatever their w / w*x
        idx[u] = (ok ? b : dummy) * C;
        ww[u] = w;
        ss[u] = __dmul_rn(w, x);
    }
    // fold equal buckets of this lane (fixed order => deterministic sums)
#pragma unroll
    for (int i = 1; i < kU; ++i) {
#pragma unroll
        for (int j = 0; j < i; ++j) {
            if (idx[i] == idx[j]) {
                ww[j] = __dadd_rn(ww[j], ww[i]);
                ss[j] = __dadd_rn(ss[j], ss[i]);
                idx[i] = dummy * C;
            }
        }
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
        if (pass == p) {
            double2 a[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) a[u] = acc_col[idx[u]];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                a[u].x = __dadd_rn(a[u].x, ww[u]);
                a[u].y = __dadd_rn(a[u].y, ss[u]);
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) acc_col[idx[u]] = a[u];
        }
        __syncwarp();
    }
}

template <bool HAS_W, bool VEC>
__device__ __forceinline__ void load_tile(const double* __restrict__ X, const double* __restrict__ W,
                                          long base, int lane, double* xs, double* ws) {
    // tile = 2 groups x 4 values x 32 lanes; lane's values: VEC -> double2 k*32+lane, k = 0..3
    if (VEC) {
        const double2* x2 = reinterpret_cast<const double2*>(X + base);
#pragma unroll
        for (int k = 0; k < kU; ++k) {
            double2 a = ldg_stream2(x2 + k * 32 + lane);
            xs[2 * k] = a.x; xs[2 * k + 1] = a.y;
        }
        if (HAS_W) {
            const double2* w2 = reinterpret_cast<const double2*>(W + base);
#pragma unroll
            for (int k = 0; k < kU; ++k) {
                double2 a = ldg_stream2(w2 + k * 32 + lane);
                ws[2 * k] = a.x; ws[2 * k + 1] = a.y;
            }
        }
    } else {
        // 8-byte aligned input: same element -> lane mapping as the vector variant, so the
        // result does not depend on the alignment of the caller's pointer
#pragma unroll
        for (int k = 0; k < 2 * kU; ++k)
            xs[k] = ldg_stream1(X + base + (k >> 1) * 64 + 2 * lane + (k & 1));
        if (HAS_W) {
#pragma unroll
            for (int k = 0; k < 2 * kU; ++k)
                ws[k] = ldg_stream1(W + base + (k >> 1) * 64 + 2 * lane + (k & 1));
        }
    }
}

template <bool HAS_W, bool VEC, int C, int GRID>
__device__ __forceinline__ void accumulate_stream(const Lookup& L, double2* acc_col, int dummy,
                                                  const double* __restrict__ X,
                                                  const double* __restrict__ W, double wsc, long n,
                                                  long gw, long tw, int lane, double& vmin,
                                                  double& vmax) {
    constexpr long TILE = 32L * 2 * kU;  // values per warp iteration
    const int pass = lane / C;
    const long ntiles = n / TILE;
    double xa[2 * kU], wa[2 * kU], xb[2 * kU], wb[2 * kU];
    double* xs = xa;
    double* ws = wa;
    // register ping-pong: the loads of tile t + tw are in flight while tile t is processed
    long t = gw;
    if (t < ntiles) load_tile<HAS_W, VEC>(X, W, t * TILE, lane, xa, wa);
    while (t < ntiles) {
        long tn = t + tw;
        if (tn < ntiles) load_tile<HAS_W, VEC>(X, W, tn * TILE, lane, xb, wb);
        accumulate_group<HAS_W, C, GRID>(L, acc_col, dummy, xa, wa, wsc, 0xfu, pass, vmin, vmax);
        accumulate_group<HAS_W, C, GRID>(L, acc_col, dummy, xa + kU, wa + kU, wsc, 0xfu, pass,
                                         vmin, vmax);
        t = tn;
        if (t >= ntiles) break;
        tn = t + tw;
        if (tn < ntiles) load_tile<HAS_W, VEC>(X, W, tn * TILE, lane, xa, wa);
        accumulate_group<HAS_W, C, GRID>(L, acc_col, dummy, xb, wb, wsc, 0xfu, pass, vmin, vmax);
        accumulate_group<HAS_W, C, GRID>(L, acc_col, dummy, xb + kU, wb + kU, wsc, 0xfu, pass,
                                         vmin, vmax);
        t = tn;
    }
    // ragged tail (< TILE values): the warp whose turn it would be
    if (gw == ntiles % tw) {
        for (long i0 = ntiles * TILE; i0 < n; i0 += 32 * kU) {
            unsigned m = 0;
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                long i = i0 + u * 32 + lane;
                bool in = i < n;
                xs[u] = in ? X[i] : 0.0;
                if (HAS_W) ws[u] = in ? W[i] : 0.0;
                m |= in ? (1u << u) : 0u;
            }
            accumulate_group<HAS_W, C, GRID>(L, acc_col, dummy, xs, ws, wsc, m, pass, vmin, vmax);
        }
    }
}

template <bool HAS_W, bool VEC, int C>
__global__ void __launch_bounds__(512, 1) k_bin_accumulate(const double* __restrict__ X,
                                                           const double* __restrict__ W, double wsc,
                                                           long n, ParHeader* hdr,
                                                           const double* __restrict__ g_edges,
                                                           const unsigned* __restrict__ g_tab,
                                                           double2* __restrict__ partial,
                                                           double2* __restrict__ minmax, int bmax,
                                                           int accumulate_into) {
    // shared: tab[kNC] | edges[bmax even] | acc[nwarps][bmax + 1][C]
    unsigned* s_tab = reinterpret_cast<unsigned*>(smem_raw);
    double* s_edges = reinterpret_cast<double*>(smem_raw + sizeof(unsigned) * kNC);
    double2* s_acc = reinterpret_cast<double2*>(s_edges + ((bmax + 1) & ~1));
    __shared__ double s_min[32], s_max[32];
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const int E = hdr->E;
    const int B = E + 1;
    const int rows = bmax + 1;  // last row = dummy
    for (int i = threadIdx.x; i < kNC; i += blockDim.x) s_tab[i] = g_tab[i];
    for (int i = threadIdx.x; i < E; i += blockDim.x) s_edges[i] = g_edges[i];
    for (int i = threadIdx.x; i < nwarps * rows * C; i += blockDim.x) s_acc[i] = make_double2(0.0, 0.0);
    __syncthreads();
    Lookup L;
    L.tab = s_tab; L.edges = s_edges; L.ubase = hdr->ubase; L.shift = hdr->shift; L.E = E;
    L.x0 = hdr->x0; L.inv = hdr->inv;
    double2* acc_col = s_acc + (size_t)warp * rows * C + (lane % C);
    double vmin = INFINITY, vmax = -INFINITY;
    const long gw = (long)blockIdx.x * nwarps + warp;
    const long tw = (long)gridDim.x * nwarps;
    if (hdr->grid == 1)
        accumulate_stream<HAS_W, VEC, C, 1>(L, acc_col, bmax, X, W, wsc, n, gw, tw, lane, vmin, vmax);
    else
        accumulate_stream<HAS_W, VEC, C, 0>(L, acc_col, bmax, X, W, wsc, n, gw, tw, lane, vmin, vmax);
    // CTA reduction in fixed (warp, column) order -> deterministic partials
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        vmin = fmin(vmin, shfl_xor_d(vmin, d));
        vmax = fmax(vmax, shfl_xor_d(vmax, d));
    }
    if (lane == 0) { s_min[warp] = vmin; s_max[warp] = vmax; }
    __syncthreads();
    double2* row = partial + (size_t)blockIdx.x * kBPad;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        double sw_ = 0.0, sx_ = 0.0;
        for (int wv = 0; wv < nwarps; ++wv) {
            const double2* r = s_acc + ((size_t)wv * rows + b) * C;
#pragma unroll
            for (int c = 0; c < C; ++c) { sw_ += r[c].x; sx_ += r[c].y; }
        }
        if (accumulate_into) { double2 o = row[b]; sw_ += o.x; sx_ += o.y; }
        row[b] = make_double2(sw_, sx_);
    }
    if (threadIdx.x == 0) {
        double mn = INFINITY, mx = -INFINITY;
        for (int wv = 0; wv < nwarps; ++wv) { mn = fmin(mn, s_min[wv]); mx = fmax(mx, s_max[wv]); }
        if (accumulate_into) { double2 o = minmax[blockIdx.x]; mn = fmin(mn, o.x); mx = fmax(mx, o.y); }
        minmax[blockIdx.x] = make_double2(mn, mx);
        if (blockIdx.x == 0) hdr->ncta = gridDim.x;
    }
}

// 5. finalize: bucket sums -> sorted weighted points -> greedy k-scale merge.
__global__ void __launch_bounds__(1024) k_finalize(BatchView v, long d, ParHeader* hdr,
                                                   const double* __restrict__ edges,
                                                   const double2* __restrict__ partial,
                                                   const double2* __restrict__ minmax,
                                                   unsigned char* gws, int mmax, int use_smem) {
    __shared__ int s_scan[1024 + 1];
    __shared__ double s_mm[2];
    __shared__ int s_m;
    __shared__ double s_W;
    const int tid = threadIdx.x;
    const int E = hdr->E;
    const int B = E + 1;
    const int ncta = hdr->ncta;
    const int cap = v.cap;
    unsigned char* wsbase = use_smem ? smem_raw : gws;
    // custom carve: cen[cap] | sb[mmax] | items[cap+mmax] | kend | start  (no unsorted buf)
    WarpWS ws;
    {
        size_t t = (size_t)cap + mmax;
        ws.cen = reinterpret_cast<double2*>(wsbase);
        ws.sb = ws.cen + cap;
        ws.buf = ws.sb;
        ws.items = ws.sb + mmax;
        ws.kend = reinterpret_cast<double*>(ws.items + t);
        ws.start = reinterpret_cast<int*>(ws.kend + t);
    }
    // a. reduce partials over CTAs (fixed order) and the batch total weight
    double bw = 0.0, bs = 0.0;
    const int b = tid;  // B <= 1024 (ov*c <= kEMax): one thread per bucket
    auto reduce_bucket = [&](int bb, double& w_, double& s_) {
        w_ = 0.0; s_ = 0.0;
        for (int cta = 0; cta < ncta; ++cta) {
            double2 a = partial[(size_t)cta * kBPad + bb];
            w_ += a.x; s_ += a.y;
        }
    };
    if (b < B) reduce_bucket(b, bw, bs);
    if (tid == 0) {
        double mn = INFINITY, mx = -INFINITY;
        for (int cta = 0; cta < ncta; ++cta) {
            mn = fmin(mn, minmax[cta].x); mx = fmax(mx, minmax[cta].y);
        }
        s_mm[0] = mn; s_mm[1] = mx;
    }
    // batch weight: sum of bucket weights in bucket order (deterministic)
    s_scan[tid] = 0;
    __syncthreads();
    __shared__ double s_bw[1024];
    if (b < B) s_bw[b] = bw;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int i = 0; i < B; ++i) t += s_bw[i];
        s_W = t;
    }
    __syncthreads();
    const double Wb = s_W;          // weight of this batch
    const double mn = s_mm[0], mx = s_mm[1];
    // b. pieces per bucket.  A single-valued bucket [v, nextafter(v)) heavier than
    // half a k-unit is split into equal pieces with the same mean, so that the
    // digest keeps several centroids on a heavy duplicate like the reference does.
    auto pieces_of = [&](int bb, double w_, double before) -> int {
        if (!(w_ > 0.0)) return 0;
        bool single = bb >= 1 && bb <= E - 1 && edges[bb] == next_up(edges[bb - 1]);
        if (!single) return 1;
        double q0 = before / Wb, q1 = (before + w_) / Wb;
        double span = kscale(v.comp, q1) - kscale(v.comp, q0);
        int p = (int)ceil(span * 2.0);
        return p < 1 ? 1 : p;
    };
    // cumulative weight before each bucket (exclusive, bucket order)
    __shared__ double s_before[1024];
    if (tid == 0) {
        double run = 0.0;
        for (int i = 0; i < B; ++i) { s_before[i] = run; run += s_bw[i]; }
    }
    __syncthreads();
    int np = (b < B) ? pieces_of(b, bw, s_before[b]) : 0;
    // exclusive scan over 1024 buckets
    s_scan[tid] = np;
    __syncthreads();
    for (int dd = 1; dd < 1024; dd <<= 1) {
        int o = (tid >= dd) ? s_scan[tid - dd] : 0;
        __syncthreads();
        s_scan[tid] += o;
        __syncthreads();
    }
    int pos = s_scan[tid] - np;
    int total1024 = s_scan[1023];
    auto emit = [&](int bb, double w_, double s_, int p, int at) {
        if (p <= 0) return;
        double lo = (bb == 0) ? mn : edges[bb - 1];
        double hi = (bb == E) ? mx : next_down(edges[bb]);
        double mean = s_ / w_;
        if (p > 1 || (bb >= 1 && bb < E && edges[bb] == next_up(edges[bb - 1])))
            mean = edges[bb - 1];                    // every member equals the left edge
        mean = fmin(fmax(mean, lo), hi);
        if (at + p > mmax) p = mmax - at > 0 ? mmax - at : 0;
        double each = w_ / (double)p;
        for (int k = 0; k < p; ++k) {
            double wk = (k == p - 1) ? (w_ - each * (double)(p - 1)) : each;
            ws.sb[at + k] = make_double2(mean, wk);
        }
    };
    if (b < B) emit(b, bw, bs, np, pos);
    if (tid == 0) {
        int m = total1024;
        s_m = m > mmax ? mmax : m;
    }
    __syncthreads();
    // c. merge into the digest: warp 0 only (the greedy chain is sequential)
    if (tid < 32) {
        const int lane = tid;
        const int m = s_m;
        Hdr h;
        const DevHdr dh = v.hdr[d];
        h.vmin = dh.vmin; h.vmax = dh.vmax; h.total = dh.total; h.btotal = 0.0; h.n = dh.n; h.bn = 0;
        const double2* cg = v.cen + d * (long)cap;
        for (int i = lane; i < h.n; i += 32) ws.cen[i] = cg[i];
        __syncwarp();
        if (m > 0) {
            // batch weight in item order, sequential (pieces re-sum to the bucket weight
            // only up to rounding, so sum what is actually merged)
            double bt = 0.0;
            if (lane == 0) for (int i = 0; i < m; ++i) bt = __dadd_rn(bt, ws.sb[i].y);
            bt = shfl_d(bt, 0);
            if (h.vmin > mn) h.vmin = mn;
            if (h.vmax < mx) h.vmax = mx;
            h.total = __dadd_rn(h.total, bt);
            warp_merge_sorted(ws, v.comp, h, m, cap, lane);
            double2* cgo = v.cen + d * (long)cap;
            for (int i = lane; i < h.n; i += 32) cgo[i] = ws.cen[i];
            if (lane == 0) {
                DevHdr o = dh;
                o.vmin = h.vmin; o.vmax = h.vmax; o.total = h.total; o.btotal = 0.0;
                o.n = h.n; o.bn = 0;
                v.hdr[d] = o;
            }
        }
    }
}

// ---------------------------------------------------------------------------
long parallel_sample_pos(long i, long n, long S) {
    long seg = n / S;
    if (seg <= 0) return i;
    return i * seg + (long)(mix64((unsigned long long)i) % (unsigned long long)seg);
}
long parallel_sample_len(long n);
static long sample_len(long n) { return parallel_sample_len(n); }
long parallel_sample_len(long n) {
    long S = (n >= (1L << 22)) ? kSMax : 16384;
    while (S > kSortChunk && S * 4 > n) S >>= 1;
    return S;
}

static cudaError_t edges_from_sorted_sample(const BatchView& v, ParLayout& L, long S,
                                            cudaStream_t st) {
    cudaError_t e = sort_sample(L.sx, L.sw, S, st);
    if (e != cudaSuccess) return e;
    count_launch();
    k_make_edges<<<1, 1024, 0, st>>>(L.hdr, L.sx, L.sw, L.cw, L.edges, L.tab, L.tab2, v.comp,
                                     ov_for(v.comp));
    return cudaGetLastError();
}

// host supplies ns (<= kSMax) sample pairs already copied into parallel_sample_x/w
__global__ void k_pad_sample(double* sx, double* sw, long ns, long S, ParHeader* hdr) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) hdr->S = S;
    if (i >= S) return;
    if (i >= ns) { sx[i] = INFINITY; sw[i] = 0.0; return; }
    double x = sx[i], w = sw[i];
    if (!(is_finite_d(x) && !(w <= DBL_EPSILON))) { sx[i] = INFINITY; sw[i] = 0.0; }
    else sx[i] = x + 0.0;
}

cudaError_t parallel_begin_from_sample(const BatchView& v, long ns, void* scratch,
                                       cudaStream_t st) {
    ParLayout L = carve(scratch, v.comp, 1);
    long S = kSortChunk;
    while (S < ns && S < kSMax) S <<= 1;
    count_launch();
    k_pad_sample<<<(unsigned)((S + 255) / 256), 256, 0, st>>>(L.sx, L.sw, ns, S, L.hdr);
    return edges_from_sorted_sample(v, L, S, st);
}

// Shape of the hot kernel for a bucket capacity: C columns per warp table, nw warps.
struct AccCfg { int C, nw; size_t smem; };
static AccCfg acc_cfg(int bmax) {
    const size_t fixed = sizeof(unsigned) * kNC + sizeof(double) * (size_t)((bmax + 1) & ~1);
    const size_t budget = 230000;                       // of 232448 B opt-in, 512 B are static
    const size_t col = (size_t)(bmax + 1) * 16;         // one column incl. the dummy row
    int cols = (int)((budget - fixed) / col);
    AccCfg c;
    c.C = 8;
    while (c.C > 1 && cols / c.C < 8) c.C >>= 1;        // prefer >= 8 warps
    c.nw = cols / c.C;
    if (c.nw > 16) c.nw = 16;
    if (c.nw < 1) c.nw = 1;
    if (const char* e = getenv("CRICK_ACC_C")) {        // tuning hook (bench sweeps)
        int v = atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8) { c.C = v; c.nw = cols / v > 16 ? 16 : cols / v; }
    }
    if (const char* e = getenv("CRICK_ACC_W")) {
        int v = atoi(e);
        if (v >= 1 && v <= 16 && v * c.C <= cols) c.nw = v;
    }
    c.smem = (fixed + (size_t)c.nw * c.C * col + 15) & ~(size_t)15;
    return c;
}

// ---- optional per-launch timing of the hot kernel (bench.py's roofline leg) ------------
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_ev;

void profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = on != 0;
}
// synchronises the recorded events, returns summed milliseconds and the launch count
int profile_collect(double* total_ms, long long* launches) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    double tot = 0.0;
    long long n = 0;
    for (auto& pr : g_prof_ev) {
        float ms = 0.f;
        if (cudaEventSynchronize(pr.second) == cudaSuccess &&
            cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) { tot += ms; ++n; }
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    g_prof_ev.clear();
    *total_ms = tot;
    *launches = n;
    return 0;
}

static cudaError_t accumulate_impl(const BatchView& v, const double* X, const double* W, double wsc,
                                   long n, void* scratch, int sm_count, cudaStream_t st,
                                   int accumulate_into) {
    ParLayout L = carve(scratch, v.comp, sm_count);
    const int bmax = bmax_for(v.comp);
    const AccCfg cfg = acc_cfg(bmax);
    const size_t smem = cfg.smem;
    const bool vec = ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W)) & 15) == 0;
    const unsigned grid = (unsigned)sm_count;
    const unsigned block = (unsigned)cfg.nw * 32;
    cudaError_t e;
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        if (g_prof_on && cudaEventCreate(&pe0) == cudaSuccess && cudaEventCreate(&pe1) == cudaSuccess)
            cudaEventRecord(pe0, st);
    }
#define CRICK_LAUNCH_ACC(HW, VC, CC)                                                              \
    do {                                                                                          \
        e = cudaFuncSetAttribute(k_bin_accumulate<HW, VC, CC>,                                    \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
        if (e != cudaSuccess) return e;                                                           \
        count_launch();                                                                           \
        k_bin_accumulate<HW, VC, CC><<<grid, block, smem, st>>>(                                  \
            X, W, wsc, n, L.hdr, L.edges, L.tab, L.partial, L.minmax, bmax, accumulate_into);     \
    } while (0)
#define CRICK_LAUNCH_ACC_C(HW, VC)                                                                \
    switch (cfg.C) {                                                                              \
        case 8: CRICK_LAUNCH_ACC(HW, VC, 8); break;                                               \
        case 4: CRICK_LAUNCH_ACC(HW, VC, 4); break;                                               \
        case 2: CRICK_LAUNCH_ACC(HW, VC, 2); break;                                               \
        default: CRICK_LAUNCH_ACC(HW, VC, 1); break;                                              \
    }
    if (W) { if (vec) { CRICK_LAUNCH_ACC_C(true, true) } else { CRICK_LAUNCH_ACC_C(true, false) } }
    else { if (vec) { CRICK_LAUNCH_ACC_C(false, true) } else { CRICK_LAUNCH_ACC_C(false, false) } }
#undef CRICK_LAUNCH_ACC_C
#undef CRICK_LAUNCH_ACC
    if (pe0 && pe1) {
        cudaEventRecord(pe1, st);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof_ev.emplace_back(pe0, pe1);
    }
    return cudaGetLastError();
}

cudaError_t parallel_accumulate(const BatchView& v, const double* X, const double* W, double wsc,
                                long n, void* scratch, int sm_count, cudaStream_t st, int first) {
    return accumulate_impl(v, X, W, wsc, n, scratch, sm_count, st, first ? 0 : 1);
}

cudaError_t parallel_finalize(const BatchView& v, long d, void* scratch, int sm_count,
                              cudaStream_t st) {
    ParLayout L = carve(scratch, v.comp, sm_count);
    const int mmax = mmax_for(v.comp);
    size_t wsb = fin_ws_bytes(v.cap, mmax);
    int use_smem = wsb <= 180 * 1024;
    size_t smem = use_smem ? wsb : 0;
    cudaError_t e = cudaFuncSetAttribute(k_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(use_smem ? wsb : 0));
    if (e != cudaSuccess) return e;
    count_launch();
    k_finalize<<<1, 1024, smem, st>>>(v, d, L.hdr, L.edges, L.partial, L.minmax, L.fin, mmax,
                                      use_smem);
    return cudaGetLastError();
}

cudaError_t launch_parallel_update(const BatchView& v, long d, const double* X, const double* W,
                                   double wsc, long n, void* scratch, int sm_count,
                                   cudaStream_t st) {
    ParLayout L = carve(scratch, v.comp, sm_count);
    long S = sample_len(n);
    count_launch();
    k_sample<<<(unsigned)((S + 255) / 256), 256, 0, st>>>(X, W, wsc, n, S, L.sx, L.sw, L.hdr);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = edges_from_sorted_sample(v, L, S, st);
    if (e != cudaSuccess) return e;
    e = accumulate_impl(v, X, W, wsc, n, scratch, sm_count, st, 0);
    if (e != cudaSuccess) return e;
    return parallel_finalize(v, d, scratch, sm_count, st);
}

// test hook: per-bucket (sum w, sum w*x) of the last update, reduced over CTAs in the same
// order k_finalize uses; also min/max.  Returns the bucket count.
int parallel_debug_buckets(void* scratch, double comp, int sm_count, double* out_w, double* out_s,
                           double* out_minmax, int max_out, cudaStream_t st) {
    ParLayout L = carve(scratch, comp, sm_count);
    ParHeader h;
    if (cudaMemcpyAsync(&h, L.hdr, sizeof(h), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    const int B = h.E + 1;
    if (h.ncta <= 0 || B > max_out) return -1;
    std::vector<double2> rows((size_t)h.ncta * kBPad), mm((size_t)h.ncta);
    if (cudaMemcpyAsync(rows.data(), L.partial, rows.size() * sizeof(double2), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaMemcpyAsync(mm.data(), L.minmax, mm.size() * sizeof(double2), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    for (int b = 0; b < B; ++b) {
        double w = 0.0, s = 0.0;
        for (int c = 0; c < h.ncta; ++c) { w += rows[(size_t)c * kBPad + b].x; s += rows[(size_t)c * kBPad + b].y; }
        out_w[b] = w; out_s[b] = s;
    }
    double mn = INFINITY, mx = -INFINITY;
    for (int c = 0; c < h.ncta; ++c) { mn = fmin(mn, mm[c].x); mx = fmax(mx, mm[c].y); }
    out_minmax[0] = mn; out_minmax[1] = mx;
    return B;
}

int parallel_debug_edges(void* scratch, double* out, int max_out, cudaStream_t st) {
    ParLayout L = carve(scratch, 100.0, 1);
    ParHeader h;
    if (cudaMemcpyAsync(&h, L.hdr, sizeof(h), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    int E = h.E < max_out ? h.E : max_out;
    if (E > 0) {
        if (cudaMemcpyAsync(out, L.edges, sizeof(double) * E, cudaMemcpyDeviceToHost, st) != cudaSuccess)
            return -1;
        if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    }
    return h.E;
}

}  // namespace crick
