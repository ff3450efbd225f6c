// crick_b200/csrc/tdigest_parallel.cu
//
// Parallel "k-bucket" update for one huge stream (SURVEY.md 2.3 K1/K3/K5, 7 option (c)).
//
// The reference folds values into centroids 42 at a time through a strictly
// sequential chain (tdigest_stubs.c:219-298); 1e9 values would be 23.8 M
// dependent flushes.  A centroid, however, is only (sum w, sum w*x) over a
// contiguous quantile range, and both sums are order independent.  So:
//
//   1. k_sort_runs     strided+hashed sample of S = 16 Ki (x, w) pairs (reads S*16 B), sorted
//                      by x in 1024-element runs on 16 SMs (bitonic in shared memory)
//   2. k_merge_runs    the runs merged by ranking every element against the other runs
//   3. k_make_edges    weighted sample quantiles at the k-scale grid
//                      q_j = sin^2(pi*j/(2*ov*c)), j = 1..ov*c-1  (ov = 2 sub-buckets
//                      per unit of k, so NO asin per element), heavy duplicates get a
//                      bucket of their own [v, nextafter(v)); plus an 8192-cell uint16 lookup
//                      table over x (linear) or over the order-preserving bit pattern of x
//   4. k_bin_accumulate  THE HOT KERNEL: every value is read exactly once from HBM
//                      (128-bit streaming loads, next tile prefetched into L2 or staged by
//                      TMA), filtered like tdigest_add (:285), mapped to its bucket (one
//                      table probe + 1-2 predicated edge compares) and accumulated into
//                      column-private (sum w, sum w*x) tables in shared memory -- no atomics,
//                      no MATCH: lane = column, and the warps sharing a table take turns
//                      through a token ring over named barriers.  Optionally the
//                      SummaryStats moments of x and the check of w ride in the same pass.
//   5. k_finalize      bucket sums -> <= ov*c weighted points in sorted order, which are
//                      merged into the digest with the reference's own greedy k-scale
//                      merge (warp_merge_sorted == tdigest_flush phases 2-5).
//   6. k_merge_records_bulk  multi-GPU tail: gathered records merged in one greedy pass.
//
// Result: the digest the reference would produce had it been handed the
// pre-aggregated bucket points in one flush.  Bit parity with the sequential
// reference is impossible by construction (different clustering); parity here
// is (a) bucket sums vs a CPU restatement of the same partition at 1e-12 and
// (b) rank error <= the reference's on the same data (tests/test_gpu_parallel.py).
#include "kernels.h"
#include "stats_dev.cuh"
#include "tdigest_ref.cuh"

#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

namespace crick {

extern __shared__ __align__(16) unsigned char smem_raw[];

// Programmatic dependent launch (opt-in, CRICK_PDL=1; see pdl_enabled()): the setup kernels of one
// update are a chain of short dependent grids, so each can be launched with
// cudaLaunchAttributeProgrammaticStreamSerialization -- it may be scheduled (and run its own
// prologue) while its predecessor in the stream is still running, and executes grid_dep_wait()
// before it touches anything the predecessor wrote.  With plain launches (the default) the two
// instructions below are no-ops.
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// cudaFuncSetAttribute costs a few microseconds of host time per call and an update makes four:
// remember the largest dynamic shared-memory size granted per (device, kernel).
template <typename K>
static cudaError_t need_smem(K kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> granted;
    int dev = 0;
    cudaGetDevice(&dev);
    const std::pair<int, const void*> key(dev, reinterpret_cast<const void*>(kernel));
    std::lock_guard<std::mutex> lk(mu);
    auto it = granted.find(key);
    if (it != granted.end() && it->second >= bytes) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) granted[key] = bytes;
    return e;
}

// OFF by default since round 2 (CRICK_PDL=1 turns it on).  With programmatic launches a digest
// updated on one stream while ANOTHER stream's kernels held SMs came out wrong (median 0.53 instead
// of 1.0; tests/test_gpu_parallel.py::test_pipelined_updates_equal_plain_updates, 5 of 5 runs on one
// box class, 0 of 3 with plain launches): k_finalize -- the secondary of a 148-CTA primary whose last
// CTAs were still waiting for an SM -- read bucket rows of the previous update.  Alone on the GPU
// (every primary CTA resident at once) the chain has never misbehaved, but a library whose handles
// are meant to be driven from several streams cannot assume that.  Cost: ~10 us per update.
static bool pdl_enabled() {
    static const bool on = [] {
        const char* env = getenv("CRICK_PDL");
        return env && env[0] == '1';
    }();
    return on;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_dep(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

constexpr int kNC = 8192;        // lookup-table cells (uint16 entries, + 2 sentinels)
constexpr int kTabLen = kNC + 2;
constexpr size_t kTabBytes = ((size_t)kTabLen * 2 + 15) & ~(size_t)15;
constexpr int kEMax = 1024;      // max edges (buckets - 1)
constexpr int kBPad = kEMax + 8; // row stride of the per-CTA partial sums
constexpr long kSMax = 16384;    // max sample size (rank error with 16 Ki = with 64 Ki)
constexpr int kSortChunk = 1024; // elements per sorted run = threads per CTA of k_sort_runs

struct ParHeader {
    int E;       // number of edges; buckets = E + 1, bucket b = [edge[b-1], edge[b])
    int shift;   // grid 0: cell = ((skey32(x) >> 1) - kbase) >> shift, kbase includes the "+ 1 cell"
    int kbase;
    int badw_sticky;  // set by k_finalize when it refused an update because of badw; cleared by the host
    int p2p_timeout;  // k_exchange_merge gave up waiting for a rank
    int pad1;
    long S;      // sample length (power of two)
    int ncta;    // CTAs that wrote partials
    int grid;    // lookup grid: 0 = linear in the ordered bit pattern, 1 = linear in x
    int kmax;    // most edges in any one cell of that grid
    int badw;    // set by k_bin_accumulate: some weight is not finite or not > 0 (tdigest.pyx:304)
    double sample_total;
    double x0;   // grid 1: cell = floor(fma(x, inv, x0)), x0 = 1 - (first edge) * inv
    double inv;
};

// Timeline stamps (globaltimer, ns) in the 16 u64 slots behind the header (bytes 128..255 of the
// 256-byte header block; crick_tdigest_debug_header reads them; tests/tools/step_probe.py prints
// them): 6 / 7 k_sort_runs CTA 0 start / end, 8 / 9 first tail group after its loads / end,
// 10 / 11 / 12 k_make_edges after the dependency wait / after the tail lists / end, 13 / 14 hot
// kernel CTA 0 after the wait / end, 0 / 15 / 2 / 5 k_finalize after the wait / after the bucket
// reduction / after the points / end (k_exchange_merge uses 0..5 for its own phases).
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void stamp(ParHeader* hdr, int slot) {
    reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(hdr) + 128)[slot] = global_ns();
}

struct ParLayout {
    ParHeader* hdr;
    double* edges;       // [kEMax]
    unsigned short* tab;   // [kTabLen] tab[c] = #edges in cells < c (the grid k_make_edges chose)
    double* sx;          // [kSMax]
    double* sw;          // [kSMax]
    double* cw;          // [kSMax]
    double* sx2;         // [kSMax] merged (sorted) sample when it spans several runs
    double* sw2;         // [kSMax]
    double2* tmin;       // [kSMax] per-thread minimum / maximum of the tail oversampling draws
    double2* tmax;       // [kSMax]
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
    b += sizeof(double) * kEMax + kTabBytes;
    b += 5 * sizeof(double) * kSMax;
    b += 2 * sizeof(double2) * kSMax;
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
    L.tab = reinterpret_cast<unsigned short*>(p); p += kTabBytes;
    L.sx = reinterpret_cast<double*>(p); p += sizeof(double) * kSMax;
    L.sw = reinterpret_cast<double*>(p); p += sizeof(double) * kSMax;
    L.cw = reinterpret_cast<double*>(p); p += sizeof(double) * kSMax;
    L.sx2 = reinterpret_cast<double*>(p); p += sizeof(double) * kSMax;
    L.sw2 = reinterpret_cast<double*>(p); p += sizeof(double) * kSMax;
    L.tmin = reinterpret_cast<double2*>(p); p += sizeof(double2) * kSMax;
    L.tmax = reinterpret_cast<double2*>(p); p += sizeof(double2) * kSMax;
    L.partial = reinterpret_cast<double2*>(p); p += sizeof(double2) * (size_t)(2 * sm_count) * kBPad;
    L.minmax = reinterpret_cast<double2*>(p); p += sizeof(double2) * (size_t)(2 * sm_count);
    L.fin = p;
    L.fin_bytes = fin_ws_bytes(2 * (int)ceil(comp), mmax_for(comp));
    return L;
}

double* parallel_sample_x(void* scratch) {
    return reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(scratch) + 256 +
                                     sizeof(double) * kEMax + kTabBytes);
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
__device__ __forceinline__ void sample_one(const double* __restrict__ X, const double* __restrict__ W,
                                           double wsc, long n, long S, long i, double& x, double& w) {
    const long seg = n / S;
    x = INFINITY; w = 0.0;
    long pos = -1;
    if (seg > 0) pos = i * seg + (long)(mix64((unsigned long long)i) % (unsigned long long)seg);
    else if (i < n) pos = i;
    if (pos >= 0) {
        double xv = X[pos];
        double wv = W ? W[pos] : wsc;
        if (is_finite_d(xv) && !(wv <= DBL_EPSILON)) { x = xv + 0.0; w = wv; }
    }
}

// 2. sort of the (key, value) sample, ascending by key, in two launches:
//   k_sort_runs   S / 1024 CTAs: each draws (SAMPLE) or loads its 1024 pairs and sorts them
//                 (bitonic network on registers: shuffles below distance 32, shared memory above);
//   k_merge_runs  S / 256 CTAs: every CTA stages all S keys (<= 128 KB) in shared memory and
//                 each thread ranks its own element against the other runs with branch-free
//                 binary searches (ties: lower run first, so the order is total and fixed), then
//                 scatters the pair to its final place.
// (r1_q launch list: the previous 15-launch global bitonic network took 73 us of a 195 us setup.)
// Tail oversampling.  The edges at the ends of the k-scale grid sit at q_j = sin^2(pi j / (2 ov c)):
// 6e-5 for c = 100, 2.5e-6 for c = 500 -- below the resolution 1 / S = 6e-5 of the sample, so the
// first / last buckets would all collapse onto the smallest / largest sampled value and the far
// tails (what a t-digest is bought for) would be one centroid.  Next to the S / 1024 sorting CTAs
// the sampling launch therefore runs kTailGroups more CTAs (on other SMs, 128 threads each): a
// group draws kTailGroupDraws = 2048 more values -- a second sample of 256 Ki values in all -- and
// keeps its kTailKeep smallest and largest.  Below t* = the smallest of the groups' kTailKeep-th
// smallest values EVERY draw of the big sample is among the kept ones (a group holding more than
// kTailKeep of them is a 1e-6 event), so the kept values below t* are the lower end of a sample 16
// times denser than the main one; k_make_edges reads the grid points out there from it.  Same on
// the upper side.
constexpr int kTailDraws = 16;          // the big sample has kTailDraws * S values
constexpr int kTailGroups = 128;        // CTAs (= groups) of the big sample
constexpr int kTailThreads = 256;       // active threads per group
constexpr int kTailPerThread = 8;       // draws per thread (all loads in flight, no spills at 64 registers)
constexpr int kTailKeep = 4;            // smallest / largest values kept per group
constexpr int kTailCap = kTailGroups * kTailKeep;   // 512 candidates per side
constexpr long kTailMinN = 1L << 22;
static_assert((long)kTailGroups * kTailThreads * kTailPerThread == kTailDraws * kSMax, "the big sample is kTailDraws x the main one");

template <bool SAMPLE>
__global__ void __launch_bounds__(kSortChunk) k_sort_runs(double* __restrict__ keys,
                                                          double* __restrict__ vals,
                                                          const double* __restrict__ X,
                                                          const double* __restrict__ W, double wsc,
                                                          long n, long S, ParHeader* hdr,
                                                          double2* __restrict__ tmin,
                                                          double2* __restrict__ tmax) {
    // one (key, source index) pair per thread in registers; compare-exchange partners at
    // distance < 32 come by shuffle, the others through a double-buffered shared array (one
    // barrier per stage).  Order: (key, index), so ties are deterministic.
    __shared__ double s_key[2][kSortChunk];
    __shared__ short s_idx[2][kSortChunk];
    __shared__ double s_val[kSortChunk];
    const int t = threadIdx.x;
    const long base = (long)blockIdx.x * kSortChunk;
    if (SAMPLE && blockIdx.x == 0 && t == 0) { hdr->S = S; stamp(hdr, 6); }
    // (no griddepcontrol.launch_dependents here or in k_merge_runs: with both, update -> finalize on
    // one stream became non-deterministic on driver 580 -- k_finalize, the fifth kernel of the
    // programmatic chain, intermittently ran ahead of the streaming kernel's results; ~4 us gained
    // is not worth that.  Measured round 2, tests/tools/step_probe.py.)
    double key, val;
    if (SAMPLE && (long)blockIdx.x >= S / kSortChunk) {
        // a tail oversampling group: 128 threads x 16 draws, all loads in flight; every thread keeps
        // the minimum and maximum of its draws, the group its kTailKeep smallest minima and largest
        // maxima (ranked by counting: ties by thread index), written in order
        __shared__ double2 s_mn[(kTailThreads / 32) * kTailKeep], s_mx[(kTailThreads / 32) * kTailKeep];
        if (t >= kTailThreads) return;
        const long g = (long)blockIdx.x - S / kSortChunk;                         // group 0 .. kTailGroups - 1
        // draw i of the big sample comes from segment i of n / S2 values at a hashed offset
        // (multiply-shift instead of sample_one's 64-bit modulo: no host mirror to agree with)
        const long seg2 = n / (S * kTailDraws);                                   // >= 1 (n >= kTailMinN)
        const unsigned seg32 = seg2 > 0xffffffffl ? 0xffffffffu : (unsigned)seg2;
        // Only the values are loaded for every draw; the weight is fetched for the thread's minimum
        // and maximum alone (a weight tdigest_add would drop, :285, then drops that candidate: such
        // weights are degenerate input, and the oversample only refines edges).  The phase is bound
        // by the loads a SM can keep in flight, so this halves it.
        double xs_[kTailPerThread];
        long ps_[kTailPerThread];
#pragma unroll
        for (int e = 0; e < kTailPerThread; ++e) {
            const long i = (g * kTailThreads + t) * kTailPerThread + e;
            ps_[e] = i * seg2 + (long)__umulhi((unsigned)(mix64((unsigned long long)i) >> 32), seg32);
            xs_[e] = X[ps_[e]];
        }
        double2 mn = make_double2(INFINITY, 0.0), mx = make_double2(-INFINITY, 0.0);
        long pmn = -1, pmx = -1;
#pragma unroll
        for (int e = 0; e < kTailPerThread; ++e) {
            if (is_finite_d(xs_[e])) {
                const double xv = xs_[e] + 0.0;
                if (xv < mn.x) { mn.x = xv; pmn = ps_[e]; }
                if (xv > mx.x) { mx.x = xv; pmx = ps_[e]; }
            }
        }
        {
            const double wlo = pmn < 0 ? 0.0 : (W ? W[pmn] : wsc), whi = pmx < 0 ? 0.0 : (W ? W[pmx] : wsc);
            mn = (wlo <= DBL_EPSILON || !is_finite_d(wlo)) ? make_double2(INFINITY, 0.0) : make_double2(mn.x, wlo);
            mx = (whi <= DBL_EPSILON || !is_finite_d(whi)) ? make_double2(-INFINITY, 0.0) : make_double2(mx.x, whi);
        }
        if (g == 0 && t == 0) stamp(hdr, 8);
        // the kTailKeep smallest minima / largest maxima: per warp by kTailKeep rounds of a shuffle
        // reduction on (value, lane) -- the winner drops out --, then the same over the warps'
        // winners.  Ties fall to the lower lane / warp: deterministic.
        const int lane = t & 31, wid = t >> 5;
        auto select = [&](double2 mine, bool low, double2* s_out) {
            double2 cand = mine;
            for (int r = 0; r < kTailKeep; ++r) {
                double bx = low ? cand.x : -cand.x;                    // minimise bx
                int bl = lane;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    const double ox = shfl_xor_d(bx, d);
                    const int ol = __shfl_xor_sync(CRICK_FULL, bl, d);
                    if (ox < bx || (ox == bx && ol < bl)) { bx = ox; bl = ol; }
                }
                const double wx = shfl_d(cand.x, bl), ww_ = shfl_d(cand.y, bl);
                if (lane == 0) s_out[wid * kTailKeep + r] = make_double2(wx, ww_);
                if (lane == bl) cand = make_double2(low ? INFINITY : -INFINITY, 0.0);
            }
        };
        select(mn, true, s_mn);
        select(mx, false, s_mx);
        asm volatile("bar.sync 1, 256;" ::: "memory");      // (the other 768 threads have left)
        if (wid == 0) {
            // 8 warps x kTailKeep winners = 32 candidates: one per lane, same selection once more
            double2 a = s_mn[lane], b2 = s_mx[lane];
            double2 cand = a;
            for (int r = 0; r < kTailKeep; ++r) {
                double bx = cand.x; int bl = lane;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    const double ox = shfl_xor_d(bx, d);
                    const int ol = __shfl_xor_sync(CRICK_FULL, bl, d);
                    if (ox < bx || (ox == bx && ol < bl)) { bx = ox; bl = ol; }
                }
                const double wx = shfl_d(cand.x, bl), ww_ = shfl_d(cand.y, bl);
                if (lane == 0) tmin[g * kTailKeep + r] = make_double2(wx, ww_);   // (+inf, 0): fewer valid draws
                if (lane == bl) cand = make_double2(INFINITY, 0.0);
            }
            cand = b2;
            for (int r = 0; r < kTailKeep; ++r) {
                double bx = -cand.x; int bl = lane;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    const double ox = shfl_xor_d(bx, d);
                    const int ol = __shfl_xor_sync(CRICK_FULL, bl, d);
                    if (ox < bx || (ox == bx && ol < bl)) { bx = ox; bl = ol; }
                }
                const double wx = shfl_d(cand.x, bl), ww_ = shfl_d(cand.y, bl);
                if (lane == 0) tmax[g * kTailKeep + r] = make_double2(wx, ww_);
                if (lane == bl) cand = make_double2(-INFINITY, 0.0);
            }
            if (g == 0 && lane == 0) stamp(hdr, 9);
        }
        return;
    }
    if (SAMPLE) {
        sample_one(X, W, wsc, n, S, base + t, key, val);
    } else {
        key = keys[base + t];
        val = vals[base + t];
    }
    s_val[t] = val;
    int idx = t;
    int buf = 0;
#pragma unroll
    for (int k = 2; k <= kSortChunk; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            double pk;
            int pi;
            if (j < 32) {
                pk = shfl_xor_d(key, j);
                pi = __shfl_xor_sync(CRICK_FULL, idx, j);
            } else {
                s_key[buf][t] = key;
                s_idx[buf][t] = (short)idx;
                __syncthreads();
                pk = s_key[buf][t ^ j];
                pi = s_idx[buf][t ^ j];
                buf ^= 1;
            }
            const bool partner_less = pk < key || (pk == key && pi < idx);
            const bool take_min = ((t & j) == 0) == ((t & k) == 0);
            if (partner_less == take_min) { key = pk; idx = pi; }
        }
    }
    __syncthreads();                      // s_val complete (k = 1024 had barriers, but be explicit)
    keys[base + t] = key;
    vals[base + t] = s_val[idx];
    if (SAMPLE && blockIdx.x == 0 && t == 0) stamp(hdr, 7);
}

constexpr int kMergeThreads = 256;   // S / 256 CTAs: the searches are a latency chain, so spread them
__device__ __forceinline__ int count_below(const double* run, double key, bool or_equal) {
    // #elements < key (or <= key) of a sorted run of kSortChunk, branch-free
    int pos = 0;
    if (or_equal) {
#pragma unroll
        for (int step = kSortChunk / 2; step > 0; step >>= 1) pos += (run[pos + step - 1] <= key) ? step : 0;
        pos += (run[pos] <= key) ? 1 : 0;
    } else {
#pragma unroll
        for (int step = kSortChunk / 2; step > 0; step >>= 1) pos += (run[pos + step - 1] < key) ? step : 0;
        pos += (run[pos] < key) ? 1 : 0;
    }
    return pos;
}
__global__ void __launch_bounds__(kMergeThreads) k_merge_runs(const double* __restrict__ keys,
                                                              const double* __restrict__ vals, long S,
                                                              double* __restrict__ okeys,
                                                              double* __restrict__ ovals) {
    double* sk = reinterpret_cast<double*>(smem_raw);
    const int tid = threadIdx.x;
    const long e = (long)blockIdx.x * kMergeThreads + tid;      // this thread's element
    grid_dep_wait();                                            // the sorted runs
    const double val = vals[e];
    {
        const double2* src = reinterpret_cast<const double2*>(keys);
        double2* dst = reinterpret_cast<double2*>(sk);
        // S / 2 is a multiple of 16 * kMergeThreads for S >= 8192: 16 loads in flight
        long i = tid;
        for (; i + 15 * kMergeThreads < S / 2; i += 16 * kMergeThreads) {
            double2 a[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) a[k] = src[i + k * kMergeThreads];
#pragma unroll
            for (int k = 0; k < 16; ++k) dst[i + k * kMergeThreads] = a[k];
        }
        for (; i < S / 2; i += kMergeThreads) dst[i] = src[i];
    }
    __syncthreads();
    const int R = (int)(S / kSortChunk);
    const int r = (int)(e / kSortChunk);
    const double key = sk[e];
    int rank = (int)(e - (long)r * kSortChunk);
    // ties: the lower run first -- runs before r count their elements <= key, runs after r < key
#pragma unroll 4
    for (int q = 0; q < r; ++q) rank += count_below(sk + q * kSortChunk, key, true);
#pragma unroll 4
    for (int q = r + 1; q < R; ++q) rank += count_below(sk + q * kSortChunk, key, false);
    okeys[rank] = key;
    ovals[rank] = val;
}

// Sorts S pairs (S a power of two >= kSortChunk).  The sorted pairs end up in (*skeys, *svals):
// the input arrays when S is one run, else (keys2, vals2).
struct SampleSrc { const double* X; const double* W; double wsc; long n; ParHeader* hdr; double2* tmin; double2* tmax; };
static cudaError_t sort_sample(double* keys, double* vals, double* keys2, double* vals2, long S,
                               const SampleSrc* src, cudaStream_t st, const double** skeys,
                               const double** svals) {
    const unsigned nrun = (unsigned)(S / kSortChunk);
    cudaError_t e;
    count_launch();
    if (src)
        k_sort_runs<true><<<src->tmin ? nrun + kTailGroups : nrun, kSortChunk, 0, st>>>(
            keys, vals, src->X, src->W, src->wsc, src->n, S, src->hdr, src->tmin, src->tmax);
    else
        k_sort_runs<false><<<nrun, kSortChunk, 0, st>>>(keys, vals, nullptr, nullptr, 0.0, 0, S, nullptr,
                                                        nullptr, nullptr);
    *skeys = keys; *svals = vals;
    if (nrun > 1) {
        const size_t msm = (size_t)S * 8;
        e = need_smem(k_merge_runs, msm);
        if (e != cudaSuccess) return e;
        count_launch();
        e = launch_dep(k_merge_runs, dim3((unsigned)(S / kMergeThreads)), dim3(kMergeThreads), msm, st,
                       keys, vals, S, keys2, vals2);
        if (e != cudaSuccess) return e;
        *skeys = keys2; *svals = vals2;
    }
    return cudaGetLastError();
}

// 3. edges + lookup table.  Single CTA, 1024 threads.
//
// Two candidate grids of kNC cells over [edges[0], edges[E-1]]; cell 0 = below the first edge.
// Both maps are monotone non-decreasing in x, which is all the lookup needs.
//  grid 0: linear in the (signed-orderable) HIGH WORD of the bit pattern, i.e. ~ log|x| with
//          2^-20 relative resolution -- good for heavy-tailed data of one sign;
//  grid 1: linear in x -- good for data that crosses zero or sits in a narrow band.
// signed 32-bit key of the high word: non-decreasing in x (all the lookup needs), -0.0 ties +0.0
// (negative values get +1, so the largest negative key, 0, is the key of +0.0) -- integer ALU
// only: SHF, LOP3, IADD
__device__ __forceinline__ int skey32(double x) {
    const int hi = __double2hiint(x);
    const int m = hi >> 31;
    return (hi ^ (m & 0x7fffffff)) - m;
}
__device__ __forceinline__ int skey_half(double x) { return skey32(x) >> 1; }   // >> 1: no overflow below
__device__ __forceinline__ int clamp_cell0(int c) {     // -> cell 0 .. kNC - 1
    c = c < 0 ? 0 : c;
    return c > kNC - 1 ? kNC - 1 : c;
}
// grid 0: kb1 = kbase - (1 << shift), i.e. ((skey_half - kbase) >> shift) + 1: cell 0 = below the first edge
__device__ __forceinline__ int cell_key(double x, int kb1, int shift) {
    return clamp_cell0((skey_half(x) - kb1) >> shift);
}
// grid 1: t = (x - x0) * inv + 1 as ONE fused multiply-add with c0 = 1 - x0 * inv (monotone in x:
// a single rounding of an increasing affine map), then cvt.rmi.u32.f64, which saturates (negative
// and NaN -> 0), and one unsigned min: 3 instructions.  The same function places the edges
// (k_make_edges) and the values (k_bin_accumulate), so its exact rounding does not matter.
__device__ __forceinline__ int cell_lin(double x, double c0, double inv) {
    unsigned c;
    const double t = __fma_rn(x, inv, c0);
    asm("cvt.rmi.u32.f64 %0, %1;" : "=r"(c) : "d"(t));
    return (int)(c > (unsigned)(kNC - 1) ? (unsigned)(kNC - 1) : c);
}

// inclusive scan over the 1024 threads of the CTA (warp shuffles + one pass of warp 0 over the
// 32 warp totals); s_w32 is a 32-entry scratch.  Two barriers.
template <typename T>
__device__ __forceinline__ T block_scan_incl(T v, T* s_w32) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(CRICK_FULL, v, d);
        if (lane >= d) v += o;
    }
    __syncthreads();                      // s_w32 may still be read from the previous call
    if (lane == 31) s_w32[wid] = v;
    __syncthreads();
    T t = s_w32[lane];                    // every warp scans the 32 totals itself
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(CRICK_FULL, t, d);
        if (lane >= d) t += o;
    }
    T before = __shfl_sync(CRICK_FULL, t, wid > 0 ? wid - 1 : 0);
    return wid > 0 ? v + before : v;
}

__global__ void __launch_bounds__(1024) k_make_edges(ParHeader* hdr, const double* __restrict__ sx,
                                                     const double* __restrict__ sw,
                                                     double* __restrict__ edges,
                                                     unsigned short* __restrict__ tab, double comp,
                                                     int ov, long S, const double2* __restrict__ tmin,
                                                     const double2* __restrict__ tmax) {
    // dynamic shared memory: cw[S] cumulative sample weights | the candidate table of each grid
    // static: s_buf = raw edges + dedupe counts (phases b-c), then the per-cell counts (phase d)
    __shared__ __align__(16) unsigned char s_buf[(kTabLen + 6) * 4];
    static_assert(sizeof(s_buf) >= kEMax * 8 + kEMax * 4, "phase b-c arrays must fit");
    __shared__ double s_edge[kEMax];
    __shared__ double s_wd[32];
    __shared__ int s_wi[32];
    __shared__ int s_cost[2], s_kmax[2];
    double* s_raw = reinterpret_cast<double*>(s_buf);
    int* s_cnt = reinterpret_cast<int*>(s_buf + kEMax * 8);
    const int tid = threadIdx.x;
    grid_dep_launch();        // k_bin_accumulate may take the other SMs and clear its tables
    grid_dep_wait();          // the sorted sample
    if (tid == 0) stamp(hdr, 10);
    double* cw = reinterpret_cast<double*>(smem_raw);
    unsigned short* s_tab = reinterpret_cast<unsigned short*>(smem_raw + (S + 1024) * 8);   // [2][kTabLen]
    // a. inclusive scan of the sample weights (fixed order => deterministic).  Warp w owns the
    // S / 32 consecutive elements [w * S/32, ...) and every lane `per` = S / 1024 consecutive
    // ones of them: coalesced loads go through the warp's slice of shared memory (one pad word
    // per `per` elements, so both the row-wise writes and the lane-wise reads are conflict
    // free), a serial prefix per lane, ONE shuffle scan of the lane totals (16 scans of 32
    // elements each cost 10 k cycles of SHFL throughput), the offsets of the 32 warp totals.
    // cw stays in the padded layout: element i lives at CW(i).
    const int seg = (int)(S >> 5);                 // 32 .. 512 (S is a power of two >= 1024)
    const int per = seg >> 5;                      // 1 .. 16
    const int lseg = 31 - __clz(seg), lper = 31 - __clz(per);
    auto CW = [&](long i) -> double& {
        const int p_ = (int)(i & (seg - 1));
        return cw[(i >> lseg) * (seg + 32) + p_ + (p_ >> lper)];
    };
    {
        const int lane = tid & 31, wid = tid >> 5;
        const double* src = sw + (long)wid * seg;
        double* wbuf = cw + (long)wid * (seg + 32);
        double r[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) r[k] = k * 32 < seg ? src[k * 32 + lane] : 0.0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int p_ = k * 32 + lane;
            if (k * 32 < seg) wbuf[p_ + (p_ >> lper)] = r[k];
        }
        __syncwarp();
        const int b0 = lane * (per + 1);
#pragma unroll
        for (int k = 0; k < 16; ++k) r[k] = k < per ? wbuf[b0 + k] : 0.0;
#pragma unroll
        for (int k = 1; k < 16; ++k) r[k] += r[k - 1];          // r[15] = the lane's total
        double inc = r[15];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            double o = shfl_up_d(inc, d);
            if (lane >= d) inc += o;
        }
        const double lane_before = inc - r[15];
        if (lane == 31) s_wd[wid] = inc;
        __syncthreads();
        double t = s_wd[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            double o = shfl_up_d(t, d);
            if (lane >= d) t += o;
        }
        const double before = (wid > 0 ? shfl_d(t, wid - 1) : 0.0) + lane_before;
#pragma unroll
        for (int k = 0; k < 16; ++k) if (k < per) wbuf[b0 + k] = before + r[k];
    }
    __syncthreads();
    const double Wt = CW(S - 1);
    const int c = (int)ceil(comp);
    const int E0 = ov * c - 1;
    // a'. tail oversampling (see k_sort_runs): the groups' kept values beyond t*, sorted, with
    // their cumulative weights.  Lists live where the lookup tables of phase d will be (s_tab
    // region, 33 KB): x[512] | cw[512] per side.  Ranking is an O(n^2) count over the 512
    // candidates of a side (broadcast reads), the cumulative weights one block scan in rank order.
    __shared__ int s_tn[2];                 // candidates per side (0 = no refinement there)
    double* t_x = reinterpret_cast<double*>(s_tab);          // [2][kTailCap] sorted values (hi side: descending)
    double* t_cw = t_x + 2 * kTailCap;                       // [2][kTailCap] inclusive cumulative weights
    const bool tails = tmin != nullptr && S == kSMax;
    if (tails) {
        double* c_x = t_cw + 2 * kTailCap;                   // unsorted candidates, same region (4 x 8 KB <= 2 tables)
        double* c_w = c_x + 2 * kTailCap;
        static_assert(8 * kTailCap * sizeof(double) <= 2 * kTabBytes, "tail lists must fit the table region");
        // thread i < 512: low-side candidate i, else high-side candidate i - 512; t* = the least
        // extreme of the groups' kTailKeep-th values (a group with fewer valid draws holds +-inf
        // there: then nothing is refined on that side)
        __shared__ double s_tstar[2];
        {
            const int side = tid >> 9, i = tid & (kTailCap - 1);
            const double2 a = side ? tmax[i] : tmin[i];
            c_x[tid] = a.x; c_w[tid] = a.y;
            double lim = INFINITY;                           // low: min of the kept-th smallest; high: max of the kept-th largest, negated
            if ((i & (kTailKeep - 1)) == kTailKeep - 1) lim = a.y > 0.0 ? (side ? -a.x : a.x) : -INFINITY;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) lim = fmin(lim, shfl_xor_d(lim, d));
            if ((tid & 31) == 0) s_wd[tid >> 5] = lim;
        }
        __syncthreads();
        if (tid < 2) {
            double lim = INFINITY;
            for (int k = 0; k < 16; ++k) lim = fmin(lim, s_wd[tid * 16 + k]);
            s_tstar[tid] = lim;                              // [1] is the NEGATED upper threshold
        }
        __syncthreads();
        {
            // keep the candidates beyond t*, compacted in index order (ballot + warp offsets), then
            // rank them by counting: only the kept ones (~100 per side) take part
            __shared__ int s_wcnt[32];
            const int side = tid >> 9, lane = tid & 31, wid = tid >> 5;
            const double xv = c_x[tid], wv = c_w[tid];
            const bool in = wv > 0.0 && (side ? -xv < s_tstar[1] : xv < s_tstar[0]);
            const unsigned bal = __ballot_sync(CRICK_FULL, in);
            if (lane == 0) s_wcnt[wid] = __popc(bal);
            __syncthreads();
            int before = 0, total = 0;
            for (int k = side * 16; k < side * 16 + 16; ++k) { const int cnt_k = s_wcnt[k]; before += k < wid ? cnt_k : 0; total += cnt_k; }
            if (tid == 0 || tid == kTailCap) s_tn[side] = total;
            __syncthreads();                                   // every thread has read its candidate
            if (in) {
                const int at = side * kTailCap + before + __popc(bal & ((1u << lane) - 1u));
                t_x[at] = xv; t_cw[at] = wv;                   // (compacted, unsorted: t_x / t_cw as staging)
            }
        }
        __syncthreads();
        const int nl = s_tn[0], nh = s_tn[1];
        {
            const int side = tid >> 9, i = tid & (kTailCap - 1);          // threads 0-511: low side, 512-1023: high
            const int nn = side ? nh : nl;
            if (i < nn) {
                const double xi = t_x[side * kTailCap + i], wi = t_cw[side * kTailCap + i];
                int rank = 0;
                for (int o = 0; o < nn; ++o) {
                    const double xo = t_x[side * kTailCap + o], wo = t_cw[side * kTailCap + o];
                    // low side ascending, high side descending; ties by weight, then list position
                    // (entries that tie in both are interchangeable)
                    const bool first = (side ? xo > xi : xo < xi) ||
                                       (xo == xi && (wo < wi || (wo == wi && o < i)));
                    rank += first ? 1 : 0;
                }
                c_x[side * kTailCap + rank] = xi;
                c_w[side * kTailCap + rank] = wi;
            }
        }
        __syncthreads();
        {
            const int side = tid >> 9, i = tid & (kTailCap - 1);
            const bool live = i < (side ? nh : nl);
            t_x[tid] = live ? c_x[tid] : 0.0;                  // sorted values; weights follow through the scan
            t_cw[tid] = live ? c_w[tid] : 0.0;
        }
        __syncthreads();
        {
            // inclusive cumulative weights in rank order (one scan over [low | high], fixed association)
            const int side = tid >> 9, i = tid & (kTailCap - 1);
            const double wv = i < (side ? nh : nl) ? t_cw[tid] : 0.0;
            const double incl = block_scan_incl<double>(wv, s_wd);
            __shared__ double s_lo_total;
            if (tid == kTailCap - 1) s_lo_total = incl;
            __syncthreads();
            t_cw[tid] = side ? incl - s_lo_total : incl;
        }
        __syncthreads();
    } else if (tid < 2) {
        s_tn[tid] = 0;
    }
    __syncthreads();
    if (tid == 0) stamp(hdr, 11);
    // b. raw edges at the k-scale grid
    {
        const int nl = s_tn[0], nh = s_tn[1];
        // the dense tail sample stands for kTailDraws * S draws: its expected total weight is
        // kTailDraws * Wt; it is trusted up to 70 % of what it covers
        const double lim_lo = nl ? 0.7 * t_cw[nl - 1] : 0.0, lim_hi = nh ? 0.7 * t_cw[kTailCap + nh - 1] : 0.0;
        int r_lo = 0, r_hi = 0;                  // the innermost list entries a grid point can be given
        if (nl) { int lo = 0, hi = nl - 1; while (lo < hi) { int mid = (lo + hi) >> 1; if (t_cw[mid] < lim_lo) lo = mid + 1; else hi = mid; } r_lo = lo; }
        if (nh) { int lo = 0, hi = nh - 1; while (lo < hi) { int mid = (lo + hi) >> 1; if (t_cw[kTailCap + mid] < lim_hi) lo = mid + 1; else hi = mid; } r_hi = lo; }
        for (int j = tid; j < E0; j += blockDim.x) {
            double kq = (double)(j + 1) / (double)ov;                 // k in (0, c)
            double sn = sin(CRICK_PI_2 * kq / comp);
            double q = sn * sn;                                        // k^-1(kq)
            if (q > 1.0) q = 1.0;
            const double tl = q * Wt * (double)kTailDraws, th = (1.0 - q) * Wt * (double)kTailDraws;
            if (nl && tl <= lim_lo) {
                int lo = 0, hi = nl - 1;                               // first r with t_cw[r] >= tl
                while (lo < hi) { int mid = (lo + hi) >> 1; if (t_cw[mid] < tl) lo = mid + 1; else hi = mid; }
                s_raw[j] = t_x[lo];
            } else if (nh && th <= lim_hi) {
                int lo = 0, hi = nh - 1;                               // from the top: first r with cw >= th
                while (lo < hi) { int mid = (lo + hi) >> 1; if (t_cw[kTailCap + mid] < th) lo = mid + 1; else hi = mid; }
                s_raw[j] = t_x[kTailCap + lo];
            } else {
                double target = q * Wt;
                long lo = 0, hi = S - 1;                               // first i with cw[i] >= target
                while (lo < hi) {
                    long mid = (lo + hi) >> 1;
                    if (CW(mid) < target) lo = mid + 1; else hi = mid;
                }
                // keep the raw edges monotone next to the refined tails: nothing from the main
                // sample may lie beyond the innermost value the tail lists can hand out
                double e = sx[lo];
                if (nl) e = fmax(e, t_x[r_lo]);
                if (nh) e = fmin(e, t_x[kTailCap + r_hi]);
                s_raw[j] = e;
            }
        }
    }
    __syncthreads();
    // c. dedupe; a value chosen more than once is "heavy" and gets [v, nextafter(v))
    int cnt = 0;                                                   // E0 <= 1023: thread t owns raw edge t
    if (tid < E0) {
        const int j = tid;
        double v = s_raw[j];
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
    }
    const int incl = block_scan_incl<int>(cnt, s_wi);
    if (cnt > 0) {
        const double e = s_raw[tid];
        const int at = incl - cnt;
        s_edge[at] = e; edges[at] = e;
        if (cnt == 2) { s_edge[at + 1] = next_up(e); edges[at + 1] = next_up(e); }
    }
    __shared__ int s_E;
    if (tid == 1023) s_E = incl;
    if (tid < 2) { s_cost[tid] = 0; s_kmax[tid] = 0; }
    __syncthreads();
    const int En = s_E;
    // d. lookup tables: tab[c] = #edges whose cell is < c (so the edges of cell c are
    // tab[c] .. tab[c+1]-1), for both grids; keep the one whose fullest cell holds fewer edges
    // (ties: fewer expected compares).
    int kbase = 0, shift = 0, kb1 = 0;
    double x0 = 0.0, inv = 0.0, c0 = 0.0;
    bool lin_ok = false;
    if (En > 0) {
        kbase = skey_half(s_edge[0]);
        long long span = (long long)skey_half(s_edge[En - 1]) - (long long)kbase;
        while (shift < 31 && ((span >> shift) + 1ll) > (long long)(kNC - 2)) ++shift;
        kb1 = kbase - (1 << shift);        // shift <= 19 (the span is below 2^32)
        x0 = s_edge[0];
        double width = s_edge[En - 1] - x0;
        inv = (double)(kNC - 2) / width;
        c0 = 1.0 - x0 * inv;
        lin_ok = En >= 2 && width > 0.0 && is_finite_d(width) && is_finite_d(inv) && is_finite_d(c0);
    }
    // count the edges of every cell (edges are sorted and cell() is monotone) for both grids in
    // one pass -- grid 0 in the low, grid 1 in the high half-word of the same counter -- then an
    // inclusive scan over the cells: tab[c] = sum of counts of cells < c
    int* s_cnt2 = reinterpret_cast<int*>(s_buf);
    constexpr int kPer = (kTabLen + 1023) / 1024;   // cells per thread
    {
        __syncthreads();                             // s_buf is free (phase c)
        for (int cc = tid; cc < kTabLen + 6; cc += blockDim.x) s_cnt2[cc] = 0;
        __syncthreads();
        for (int j = tid; j < En; j += blockDim.x) {
            const double e = s_edge[j];
            atomicAdd(&s_cnt2[cell_key(e, kb1, shift) + 1], 1);          // counted at index cell + 1
            if (lin_ok) atomicAdd(&s_cnt2[cell_lin(e, c0, inv) + 1], 1 << 16);
        }
        __syncthreads();
        int local[2] = {0, 0}, lmax[2] = {0, 0}, run[2] = {0, 0};
        int vals[2][kPer];
        const int c0 = tid * kPer;
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int cc = c0 + k;
            const int mm = cc < kTabLen ? s_cnt2[cc] : 0;    // = edges in cell cc - 1, both grids
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const int m = g == 0 ? (mm & 0xffff) : (mm >> 16);
                const int lg = m > 0 ? 32 - __clz(m) : 0;    // compares of a binary search over m + 1 outcomes
                local[g] += m * lg;
                lmax[g] = m > lmax[g] ? m : lmax[g];
                run[g] += m;
                vals[g][k] = run[g];
            }
        }
        const int pk = block_scan_incl<int>(run[0] | (run[1] << 16), s_wi);   // <= 1023 per field
        const int off0 = (pk & 0xffff) - run[0], off1 = (pk >> 16) - run[1];
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int cc = c0 + k;
            if (cc < kTabLen) {
                // entry = (#edges in the cells before cc) << 2 | min(#edges in cell cc, 3): the
                // low bits let the hot kernel skip the edge loads of the ~95 % of values whose
                // cell holds no edge (shared-memory wavefronts are what bounds it)
                const int nx = s_cnt2[cc + 1];
                const int n0 = nx & 0xffff, n1 = nx >> 16;
                s_tab[cc] = (unsigned short)(((off0 + vals[0][k]) << 2) | (n0 < 3 ? n0 : 3));
                s_tab[kTabLen + cc] = (unsigned short)(((off1 + vals[1][k]) << 2) | (n1 < 3 ? n1 : 3));
            }
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int lsum = __reduce_add_sync(CRICK_FULL, local[g]);
            const int lmx = __reduce_max_sync(CRICK_FULL, lmax[g]);
            if ((tid & 31) == 0) { atomicAdd(&s_cost[g], lsum); atomicMax(&s_kmax[g], lmx); }
        }
    }
    __syncthreads();
    int grid = 0;
    if (lin_ok) {
        if (s_kmax[0] != s_kmax[1]) grid = s_kmax[1] < s_kmax[0] ? 1 : 0;
        else grid = s_cost[1] <= s_cost[0] ? 1 : 0;
    }
    {
        const unsigned* src = reinterpret_cast<const unsigned*>(s_tab + grid * kTabLen);
        unsigned* dst = reinterpret_cast<unsigned*>(tab);
        for (int i = tid; i < kTabLen / 2; i += blockDim.x) dst[i] = src[i];
    }
    if (tid == 0) {
        hdr->E = En;
        hdr->shift = shift;
        hdr->kbase = kb1;
        hdr->sample_total = Wt;
        hdr->ncta = 0;
        hdr->grid = grid;
        hdr->badw = 0;
        hdr->kmax = s_kmax[grid];
        hdr->x0 = c0;
        hdr->inv = inv;
        stamp(hdr, 12);
    }
}

// 4. THE HOT KERNEL -----------------------------------------------------------
//
// Weighted histogram of (sum w, sum w*x) over B <= bmax buckets, one pass over HBM.
// History (profiles/): (r1_a) resolving intra-warp bucket collisions with __match_any_sync
// saturates the ADU pipe (92 %) at ~0.5 value/clk/SM; (r1_b) per-warp tables with 8 columns and
// 4 lane passes: 103 instructions per 32 values on 8 warps, latency bound.  This version:
//   * TWO tables acc[bucket][C] of double2 in shared memory (acc_cfg below), C = 32 columns
//     when they fit (c <= 100), else 16 / 8 / 4: lane l owns column l % C, so LDS.128 / STS.128
//     of a warp are conflict-free and need neither MATCH nor atomics (C < 32: the 32/C lanes
//     that share a column take turns in passes);
//   * the 7 warps of a table take turns on it: a token travels round-robin over named barriers
//     (bar.sync / bar.arrive on 64 threads) and only the holder does its read-modify-writes,
//     8 LDS + 16 DADD + 8 STS per 256-value tile; loads, bucket lookup and de-duplication run
//     outside the critical section on all 14 warps.  The fixed order makes sums deterministic;
//   * a lane's 4 values of one group are de-duplicated in registers first (equal buckets are
//     folded, the later one is redirected to a dummy row), so the 4 read-modify-writes are
//     independent and overlap.
// Bucket lookup: one uint16 probe of the 8192-cell table (k_make_edges picks the grid), then a
// branch-free count of the <= kmax edges of that cell that are <= x.  Shared memory is
// addressed with 32-bit shared-window addresses (inline PTX) to keep the integer work short.
constexpr int kU = 4;         // values per lane per group
constexpr int kKFast = 8;     // branch-free lookup when no cell holds more edges than this
constexpr int kEdgePad = 8;   // +inf edges after the last one (the branch-free scan overruns)

__device__ __forceinline__ unsigned lds_u16(unsigned addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 lds_f64x2(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64x2(unsigned addr, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void bar_sync64(int id) {
    asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}
__device__ __forceinline__ void bar_arrive64(int id) {
    asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory");
}

// if (ai == aj) { wj += wi; sj += si; ai = daddr; } without branches: f = 1.0 or 0.0 from one
// 32-bit select, then two fused multiply-adds (exact: the product is wi or 0).  ptxas turns a
// predicated add.f64 -- also one written as "@p add.f64" in PTX -- into DADD + 2 FSEL per double
// (r1_g profile: 14 FSEL per value; checked again in round 2).
// Caveat: a non-finite wi / si (w*x overflow) times 0 is NaN.
__device__ __forceinline__ void fold_if_equal(unsigned& ai, unsigned aj, double& wj, double& sj,
                                              double wi, double si, unsigned daddr) {
    const bool eq = ai == aj;
    const double f = __hiloint2double(eq ? 0x3ff00000 : 0, 0);
    wj = __fma_rn(wi, f, wj);
    sj = __fma_rn(si, f, sj);
    ai = eq ? daddr : ai;
}
// the same for one sum (unweighted input)
__device__ __forceinline__ void fold_if_equal1(unsigned& ai, unsigned aj, double& sj, double si,
                                               unsigned daddr) {
    const bool eq = ai == aj;
    const double f = __hiloint2double(eq ? 0x3ff00000 : 0, 0);
    sj = __fma_rn(si, f, sj);
    ai = eq ? daddr : ai;
}

struct Lookup {
    unsigned tab;     // shared-window byte addresses
    unsigned edges;
    unsigned acc;     // this lane's column: acc base + (lane % C) * 16
    int row_shift;    // log2(C * 16): byte stride of a bucket row
    int dummy;        // dummy bucket index (= bmax)
    int kbase, shift; // grid 0
    double x0, inv;   // grid 1
    int E, kmax;
    unsigned ext_lo, ext_span;  // bucket b may hold the min/max iff (b - ext_lo) >= ext_span (unsigned)
};

// One group: kU (x, w) pairs of this lane -> shared addresses of the accumulators + the
// (w, w*x) to add, equal buckets folded.  Returns true if some value may be the min / max.
template <bool HAS_W, int GRID, bool FAST>
__device__ __forceinline__ bool classify_group(const Lookup& L, const double* xs, const double* ws,
                                               double wsc, unsigned* addr, double* ww, double* ss,
                                               bool& badw) {
    int b[kU];
    unsigned ea[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
        int cell = GRID == 1 ? cell_lin(xs[u], L.x0, L.inv) : cell_key(xs[u], L.kbase, L.shift);
        unsigned ta = L.tab + 2u * (unsigned)cell;
        b[u] = (int)(lds_u16(ta) >> 2);                    // (low 2 bits: #edges in the cell, unused here)
        if (FAST) ea[u] = L.edges + 8u * (unsigned)b[u];
        else ea[u] = lds_u16(ta + 2u) >> 2;                // hi = first edge of the next cell
    }
    if (FAST) {
        // edges of later cells are > x and the pad is +inf: counting kmax entries is exact
#pragma unroll 1
        for (int j = 0; j < L.kmax; ++j) {
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                // b += (edge <= x): a predicated add (the compiler's select costs 3 moves)
                asm volatile(
                    "{\n\t.reg .pred p;\n\t.reg .f64 e;\n\t"
                    "ld.shared.f64 e, [%1];\n\t"
                    "setp.le.f64 p, e, %2;\n\t"
                    "@p add.s32 %0, %0, 1;\n\t}"
                    : "+r"(b[u]) : "r"(ea[u]), "d"(xs[u]));
                ea[u] += 8u;
            }
        }
    } else {
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            int lo = b[u], hi = (int)ea[u];
            while (lo < hi) {  // #edges <= x among the cell's candidates
                int mid = (lo + hi) >> 1;
                if (lds_f64(L.edges + 8u * (unsigned)mid) <= xs[u]) lo = mid + 1; else hi = mid;
            }
            b[u] = lo;
        }
    }
    bool ext = false;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
        const double w = HAS_W ? ws[u] : wsc;
        // tdigest_add's filter (:285): finite x, and not (w <= eps)
        // tdigest_add's filter (:285) is "finite x and not (w <= eps)".  Here a weight passes
        // when eps < w <= DBL_MAX; anything else is dropped and, unless it is one of the legal
        // tiny weights 0 < w <= eps, reported (the wrapper's check, tdigest.pyx:304-305) -- the
        // exact test runs only in the rare warp-steps that saw such a weight.  Padding lanes
        // of the last tile carry x = NaN, w = 1.
        const bool finite = fabs(xs[u]) <= DBL_MAX;
        const bool wok = !HAS_W || (w > DBL_EPSILON && w <= DBL_MAX);
        const bool ok = finite && wok;
        if (HAS_W) badw |= !wok;
        ext |= ok && ((unsigned)b[u] - L.ext_lo >= L.ext_span);
        // rejected values go to the dummy row (never read back), whatever their w / w*x --
        // but w*x must stay finite: fold_if_equal multiplies it by 0 (a NaN or infinite x
        // becomes a denormal: high word cleared)
        addr[u] = L.acc + ((unsigned)(ok ? b[u] : L.dummy) << L.row_shift);
        ww[u] = w;
        const double xf = __hiloint2double(finite ? __double2hiint(xs[u]) : 0, __double2loint(xs[u]));
        ss[u] = __dmul_rn(w, xf);
    }
    // fold equal buckets of this lane (fixed order => deterministic sums)
    const unsigned daddr = L.acc + ((unsigned)L.dummy << L.row_shift);
#pragma unroll
    for (int i = 1; i < kU; ++i) {
#pragma unroll
        for (int j = 0; j < i; ++j) {
            fold_if_equal(addr[i], addr[j], ww[j], ss[j], ww[i], ss[i], daddr);
        }
    }
    return ext;
}

// exact min / max: only values of buckets 0, 1, E-1, E can be it (edges are data values, or
// nextafter of one), so this runs for ~1 warp-step in 100
template <bool HAS_W>
__device__ __forceinline__ void minmax_group(const Lookup& L, const double* xs, const double* ws,
                                             double wsc, double& vmin, double& vmax) {
#pragma unroll
    for (int u = 0; u < kU; ++u) {
        const double w = HAS_W ? ws[u] : wsc;
        const bool ok = (fabs(xs[u]) <= DBL_MAX) && (!HAS_W || (w > DBL_EPSILON && w <= DBL_MAX));
        if (ok) {
            vmin = xs[u] < vmin ? xs[u] : vmin;
            vmax = xs[u] > vmax ? xs[u] : vmax;
        }
    }
}

// rare path: some weight of the tile is not in (eps, DBL_MAX]; is any of them illegal?
__device__ __forceinline__ bool any_illegal_weight(const double* ws) {
    bool bad = false;
#pragma unroll
    for (int k = 0; k < 2 * kU; ++k) bad |= !(ws[k] > 0.0 && ws[k] <= DBL_MAX);
    return bad;
}

// read-modify-write of N accumulators with pairwise distinct addresses (or the dummy row):
// all loads first, so the token holder pays one LDS latency, not N
template <int N>
__device__ __forceinline__ void rmw(const unsigned* addr, const double* ww, const double* ss) {
    double2 a[N];
#pragma unroll
    for (int u = 0; u < N; ++u) a[u] = lds_f64x2(addr[u]);
#pragma unroll
    for (int u = 0; u < N; ++u) {
        a[u].x = __dadd_rn(a[u].x, ww[u]);
        a[u].y = __dadd_rn(a[u].y, ss[u]);
    }
#pragma unroll
    for (int u = 0; u < N; ++u) sts_f64x2(addr[u], a[u]);
}

// tile = 2 groups x 4 values x 32 lanes; lane's values are the double2s k*32 + lane, k = 0..3
constexpr long kTile = 32L * 2 * kU;

template <bool HAS_W, bool VEC>
__device__ __forceinline__ void load_tile(const double* __restrict__ X, const double* __restrict__ W,
                                          long base, int lane, double* xs, double* ws) {
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

// the last, partial tile: guarded loads, same element -> lane mapping; returns the valid mask
template <bool HAS_W>
__device__ __forceinline__ unsigned load_partial(const double* __restrict__ X,
                                              const double* __restrict__ W, long base, long n,
                                              int lane, double* xs, double* ws) {
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 2 * kU; ++k) {
        long i = base + (k >> 1) * 64 + 2 * lane + (k & 1);
        bool in = i < n;
        xs[k] = in ? X[i] : __longlong_as_double(0x7ff8000000000000ll);   // NaN: dropped
        if (HAS_W) ws[k] = in ? W[i] : 1.0;
        m |= in ? (1u << k) : 0u;
    }
    return m;
}

struct Ring {
    int id_in, id_out;  // named barriers: token into this warp / into the next one
    bool first;         // ring position 0 starts with the token
    int passes;         // 32 / C
    int pass;           // lane / C
};

// ---- TMA staging: one 4 KB stage (2 KB of x + 2 KB of w) and one mbarrier per warp ---------
// The next tile is fetched by cp.async.bulk (SASS: UBLKCP) while the current one is classified;
// completion is signalled on the mbarrier, so the prefetch shares no register scoreboard with
// the tile in flight (the r1_d/r1_e profiles: with LDG prefetch into registers 32 % of the warp
// time went to long-scoreboard waits on the *newer* load).
constexpr int kStageBytes = 2 * (int)kTile * 8;

__device__ __forceinline__ void mbar_init(unsigned mbar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "MBAR_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MBAR_DONE_%=;\n\t"
        "bra MBAR_WAIT_%=;\n"
        "MBAR_DONE_%=:\n\t}"
        ::"r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(unsigned dst, const void* src, unsigned bytes,
                                            unsigned mbar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

template <bool HAS_W>
__device__ __forceinline__ void stage_issue(const double* __restrict__ X, const double* __restrict__ W,
                                            long base, unsigned stage, unsigned mbar) {
    mbar_expect_tx(mbar, HAS_W ? kStageBytes : kStageBytes / 2);
    tma_load_1d(stage, X + base, kStageBytes / 2, mbar);
    if (HAS_W) tma_load_1d(stage + kStageBytes / 2, W + base, kStageBytes / 2, mbar);
}

// stage-free alternative: pull the next tile into L2; the LDGs of the next round then wait for
// an L2 hit, not for DRAM, and no shared memory is spent on staging.  One 128-byte line per
// lane: lanes 0-15 cover the 2 KB of x, lanes 16-31 the 2 KB of w.
template <bool HAS_W>
__device__ __forceinline__ void l2_prefetch(const double* __restrict__ X, const double* __restrict__ W,
                                            long base, int lane) {
    const double* p = (HAS_W && lane >= 16) ? W + base + (lane - 16) * 16 : X + base + lane * 16;
    if (HAS_W || lane < 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

template <bool HAS_W>
__device__ __forceinline__ void stage_read(unsigned stage, int lane, double* xs, double* ws) {
#pragma unroll
    for (int k = 0; k < kU; ++k) {
        double2 a = lds_f64x2(stage + 16u * (unsigned)(k * 32 + lane));
        xs[2 * k] = a.x; xs[2 * k + 1] = a.y;
    }
    if (HAS_W) {
#pragma unroll
        for (int k = 0; k < kU; ++k) {
            double2 a = lds_f64x2(stage + kStageBytes / 2 + 16u * (unsigned)(k * 32 + lane));
            ws[2 * k] = a.x; ws[2 * k + 1] = a.y;
        }
    }
}

// VEC: 16-byte aligned input.  STAGE: tiles arrive through the per-warp TMA stage; otherwise
// they are prefetched into L2 and loaded with LDG.128 (VEC) or plain loads (!VEC).
template <bool HAS_W, bool VEC, bool STAGE, bool STATS, int GRID, bool FAST>
__device__ __forceinline__ void accumulate_stream(const Lookup& L, const Ring ring, unsigned stage,
                                                  unsigned mbar, const double* __restrict__ X,
                                                  const double* __restrict__ W, double wsc, long n,
                                                  long gw, long tw, int lane, double& vmin,
                                                  double& vmax, int* badw_flag, StatsAcc& sa,
                                                  StatsPart& sp, long idx_base) {
    const long nfull = n / kTile;
    const long ntiles = (n + kTile - 1) / kTile;
    const long rounds = (ntiles + tw - 1) / tw;  // the same for every warp: the token needs all
    double xs[2 * kU], ws[2 * kU];
    unsigned addr[2 * kU];
    double ww[2 * kU], ss[2 * kU];
    unsigned parity = 0;
    bool badw = false;
    long t = gw;
    if (STAGE) {
        if (lane == 0 && t < nfull) stage_issue<HAS_W>(X, W, t * kTile, stage, mbar);
    } else if (t < nfull) {
        // !STAGE pipeline: the tile of round r + 1 is loaded into the (by then dead) registers
        // at the end of round r, its latency hidden behind the wait for the token, and the tile
        // of round r + 2 is pulled into L2 at the same time
        load_tile<HAS_W, VEC>(X, W, t * kTile, lane, xs, ws);
        if (VEC && t + tw < nfull) l2_prefetch<HAS_W>(X, W, (t + tw) * kTile, lane);
    }
    for (long r = 0; r < rounds; ++r, t += tw) {
        const long tn = t + tw;
        const bool have = t < ntiles;   // warp-uniform; m is per lane
        unsigned m = 0;
        if (t < nfull) {
            m = 0xffu;
            if (STAGE) {
                mbar_wait(mbar, parity);
                parity ^= 1u;
                stage_read<HAS_W>(stage, lane, xs, ws);
                __syncwarp();  // every lane has read the stage before it is refilled
                if (lane == 0 && tn < nfull) stage_issue<HAS_W>(X, W, tn * kTile, stage, mbar);
            }
        } else if (have) {
            m = load_partial<HAS_W>(X, W, t * kTile, n, lane, xs, ws);
        }
        if (STATS && have) {
            // SummaryStats moments of x in the same pass (stats_update_ndarray :199-210 with
            // count 1): NaN skipped, everything else -- infinities included -- counted
            const long i0 = idx_base + t * kTile + 2 * lane;
            if (m == 0xffu) {
                stats_block_unit<2 * kU>(sa, sp, xs, [&](int k) { return i0 + (k >> 1) * 64 + (k & 1); });
            } else {
#pragma unroll
                for (int k = 0; k < 2 * kU; ++k)
                    if ((m >> k) & 1u) stats_one<true>(sa, sp, xs[k], 1, i0 + (k >> 1) * 64 + (k & 1));
            }
        }
        if (have) {
            bool susp = false;
            bool e0 = classify_group<HAS_W, GRID, FAST>(L, xs, ws, wsc, addr, ww, ss, susp);
            bool e1 = classify_group<HAS_W, GRID, FAST>(L, xs + kU, ws + kU, wsc, addr + kU,
                                                        ww + kU, ss + kU, susp);
            if (__any_sync(CRICK_FULL, e0 || e1)) {
                minmax_group<HAS_W>(L, xs, ws, wsc, vmin, vmax);
                minmax_group<HAS_W>(L, xs + kU, ws + kU, wsc, vmin, vmax);
            }
            if (HAS_W && __any_sync(CRICK_FULL, susp)) badw |= susp && any_illegal_weight(ws);
        }
        if (!STAGE && tn < nfull) {
            load_tile<HAS_W, VEC>(X, W, tn * kTile, lane, xs, ws);
            if (VEC && tn + tw < nfull) l2_prefetch<HAS_W>(X, W, (tn + tw) * kTile, lane);
        }
        if (!(ring.first && r == 0)) bar_sync64(ring.id_in);
        if (have) {
            // two groups of 4 independent read-modify-writes; a bucket shared by the groups is
            // ordered by the thread's own program order
            if (ring.passes == 1) {
                rmw<kU>(addr, ww, ss);
                rmw<kU>(addr + kU, ww + kU, ss + kU);
            } else {
                for (int p_ = 0; p_ < ring.passes; ++p_) {
                    if (ring.pass == p_) {
                        rmw<kU>(addr, ww, ss);
                        rmw<kU>(addr + kU, ww + kU, ss + kU);
                    }
                    __syncwarp();
                }
            }
        }
        bar_arrive64(ring.id_out);
    }
    if (ring.first && rounds > 0) bar_sync64(ring.id_in);  // absorb the last hand-over
    if (HAS_W && __any_sync(CRICK_FULL, badw) && lane == 0) atomicOr(badw_flag, 1);
}

template <bool HAS_W, bool VEC, bool STAGE, bool STATS>
__global__ void __launch_bounds__(STATS ? 384 : 448, 1) k_bin_accumulate(const double* __restrict__ X,
                                                           const double* __restrict__ W, double wsc,
                                                           long n, ParHeader* hdr,
                                                           const double* __restrict__ g_edges,
                                                           const unsigned short* __restrict__ g_tab,
                                                           double2* __restrict__ partial,
                                                           double2* __restrict__ minmax, int bmax,
                                                           int cols, int ntab, int accumulate_into,
                                                           StatsPart* __restrict__ stats_parts,
                                                           long idx_base) {
    // shared: stage[nwarps][4 KB] (STAGE only) | acc[ntab][bmax + 1][cols] double2 |
    //         edges[bmax + kEdgePad] | tab[kTabLen] uint16 | mbar[nwarps]
    const int nwarps = blockDim.x >> 5;   // a multiple of ntab, <= 14 (one named barrier each)
    unsigned char* s_stage = smem_raw;
    double2* s_acc = reinterpret_cast<double2*>(smem_raw + (STAGE ? (size_t)nwarps * kStageBytes : 0));
    const int rows = bmax + 1;  // last row = dummy
    double* s_edges = reinterpret_cast<double*>(s_acc + (size_t)ntab * rows * cols);
    unsigned short* s_tab = reinterpret_cast<unsigned short*>(s_edges + bmax + kEdgePad);
    unsigned long long* s_mbar = reinterpret_cast<unsigned long long*>(
        reinterpret_cast<unsigned char*>(s_tab) + kTabBytes);
    __shared__ double s_min[32], s_max[32];
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    // clearing the tables needs nothing from k_make_edges: it runs before the dependency wait
    for (int i = threadIdx.x; i < ntab * rows * cols; i += blockDim.x)
        s_acc[i] = make_double2(0.0, 0.0);
    grid_dep_wait();
    const int E = hdr->E;
    const int B = E + 1;
    for (int i = threadIdx.x; i < kTabLen / 2; i += blockDim.x)
        reinterpret_cast<unsigned*>(s_tab)[i] = reinterpret_cast<const unsigned*>(g_tab)[i];
    for (int i = threadIdx.x; i < bmax + kEdgePad; i += blockDim.x)
        s_edges[i] = i < E ? g_edges[i] : INFINITY;
    const unsigned stage = (unsigned)__cvta_generic_to_shared(s_stage) + (unsigned)warp * kStageBytes;
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(s_mbar) + 8u * (unsigned)warp;
    if (STAGE) {
        if (lane == 0) mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    Lookup L;
    L.tab = (unsigned)__cvta_generic_to_shared(s_tab);
    L.edges = (unsigned)__cvta_generic_to_shared(s_edges);
    // table and ring of this warp: warps w, w + ntab, w + 2 ntab, ... share table w % ntab
    const int tab_id = warp % ntab, pos = warp / ntab, rsize = nwarps / ntab;
    L.acc = (unsigned)__cvta_generic_to_shared(s_acc + (size_t)tab_id * rows * cols) +
            16u * (unsigned)(lane & (cols - 1));
    L.row_shift = 4 + __ffs(cols) - 1;
    L.dummy = bmax;
    L.kbase = hdr->kbase; L.shift = hdr->shift; L.x0 = hdr->x0; L.inv = hdr->inv;
    L.E = E; L.kmax = hdr->kmax;
    // buckets that may hold the extremes: b <= 1 or b >= E - 1; fewer than 4 edges: all
    L.ext_lo = E >= 4 ? 2u : 0u;
    L.ext_span = E >= 4 ? (unsigned)(E - 3) : 0u;
    Ring ring;
    ring.id_in = 1 + tab_id * rsize + pos;
    ring.id_out = 1 + tab_id * rsize + (pos + 1 == rsize ? 0 : pos + 1);
    ring.first = pos == 0;
    ring.passes = 32 / cols;
    ring.pass = lane / cols;
    double vmin = INFINITY, vmax = -INFINITY;
    StatsAcc sa;
    StatsPart sp;
    stats_acc_fresh(sa);
    part_fresh(sp);
    const long gw = (long)blockIdx.x * nwarps + warp;
    const long tw = (long)gridDim.x * nwarps;
    const bool fast = L.kmax <= kKFast;
    if (hdr->grid == 1) {
        if (fast) accumulate_stream<HAS_W, VEC, STAGE, STATS, 1, true>(L, ring, stage, mbar, X, W, wsc, n, gw, tw, lane, vmin, vmax, &hdr->badw, sa, sp, idx_base);
        else accumulate_stream<HAS_W, VEC, STAGE, STATS, 1, false>(L, ring, stage, mbar, X, W, wsc, n, gw, tw, lane, vmin, vmax, &hdr->badw, sa, sp, idx_base);
    } else {
        if (fast) accumulate_stream<HAS_W, VEC, STAGE, STATS, 0, true>(L, ring, stage, mbar, X, W, wsc, n, gw, tw, lane, vmin, vmax, &hdr->badw, sa, sp, idx_base);
        else accumulate_stream<HAS_W, VEC, STAGE, STATS, 0, false>(L, ring, stage, mbar, X, W, wsc, n, gw, tw, lane, vmin, vmax, &hdr->badw, sa, sp, idx_base);
    }
    // CTA reduction in fixed column order -> deterministic partials
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        vmin = fmin(vmin, shfl_xor_d(vmin, d));
        vmax = fmax(vmax, shfl_xor_d(vmax, d));
    }
    if (lane == 0) { s_min[warp] = vmin; s_max[warp] = vmax; }
    __shared__ StatsPart s_sp[STATS ? 12 : 1];
    if (STATS) {
        stats_thread_finish(sa, sp);
        sp = stats_warp_reduce(sp);
        if (lane == 0) s_sp[warp] = sp;
    }
    __syncthreads();
    if (STATS && threadIdx.x == 0) {
        StatsPart t = s_sp[0];
        for (int wv = 1; wv < nwarps; ++wv) part_merge(t, s_sp[wv]);
        if (accumulate_into) { StatsPart o = stats_parts[blockIdx.x]; part_merge(o, t); t = o; }
        stats_parts[blockIdx.x] = t;
    }
    // column sums in place: the tables are added into table 0, then a halving tree over the
    // columns (c += c + d, d = cols/2 .. 1): the same order as a xor-shuffle tree, without the
    // 20 shuffles per bucket
    double2* row = partial + (size_t)blockIdx.x * kBPad;
    {
        const int nrc = B * cols;                      // rows 0 .. B-1 of table 0
        for (int tb = 1; tb < ntab; ++tb) {
            for (int i = threadIdx.x; i < nrc; i += blockDim.x) {
                double2 a = s_acc[i], o = s_acc[(size_t)tb * rows * cols + i];
                s_acc[i] = make_double2(a.x + o.x, a.y + o.y);
            }
        }
        __syncthreads();
        for (int dcol = cols >> 1; dcol > 0; dcol >>= 1) {
            const int sh = __ffs(dcol) - 1;
            for (int i = threadIdx.x; i < B * dcol; i += blockDim.x) {
                const int bb = i >> sh, c = i & (dcol - 1);
                double2 a = s_acc[bb * cols + c], o = s_acc[bb * cols + c + dcol];
                s_acc[bb * cols + c] = make_double2(a.x + o.x, a.y + o.y);
            }
            __syncthreads();
        }
        for (int bb = threadIdx.x; bb < B; bb += blockDim.x) {
            double2 a = s_acc[bb * cols];
            if (accumulate_into) { double2 o = row[bb]; a.x += o.x; a.y += o.y; }
            row[bb] = a;
        }
    }
    if (threadIdx.x == 0) {
        double mn = INFINITY, mx = -INFINITY;
        for (int wv = 0; wv < nwarps; ++wv) { mn = fmin(mn, s_min[wv]); mx = fmax(mx, s_max[wv]); }
        if (accumulate_into) { double2 o = minmax[blockIdx.x]; mn = fmin(mn, o.x); mx = fmax(mx, o.y); }
        minmax[blockIdx.x] = make_double2(mn, mx);
        if (blockIdx.x == 0) hdr->ncta = gridDim.x;
    }
}

// ---------------------------------------------------------------------------------------------
// 4b. k_bin_accumulate2 -- the round-2 shape of the hot kernel (the one that runs unless fused
// moments or the TMA-staged single-table variant are asked for).  Same partition, same token ring,
// same deterministic order as k_bin_accumulate; what changed (ncu r1_z: 58.6 warp instructions per
// 32 values at 57 % issue, 28 % of the warp time waiting for the token):
//   * a TILE-LEVEL screen replaces the per-value filter: one chained compare per value decides
//     whether the whole 256-value tile is clean (every x finite, every w in (eps, DBL_MAX]); clean
//     tiles -- all of them in practice -- take classify_lean (no dummy-row selects, no per-value
//     weight logic, edge compares at immediate offsets), the others the exact round-1 path;
//   * cells cost 3 instructions (grid 1: one FMA + one saturating convert + one min) or 8 integer
//     ones (grid 0, no fp64);
//   * UNWEIGHTED input (w is a scalar: the call dask makes, tdigest.pyx:294-295): sum w is a
//     COUNT.  It lives in one u32 counter per bucket updated with red.shared.add.u32 (SASS
//     ATOMS.POPC.INC: the hardware folds equal addresses of a warp, 2.7-3.3 cycles per warp
//     instruction, tests/tools/mb_atoms.cu) outside the token; only sum x (8 bytes) needs the
//     ring.  The rows halve, FOUR tables of 32 columns fit where two did, and the rings shorten
//     from 7 warps to 4 (16 warps per CTA).
template <int GRID>
__device__ __forceinline__ int cell_of(const Lookup& L, double x) {
    return GRID == 1 ? cell_lin(x, L.x0, L.inv) : cell_key(x, L.kbase, L.shift);
}

// b += (edge at [ea + OFF] <= x): LDS.64 at an immediate offset, DSETP, predicated add
template <int OFF>
__device__ __forceinline__ void edge_step(int& b, unsigned ea, double x) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .f64 e;\n\t"
        "ld.shared.f64 e, [%1+%3];\n\t"
        "setp.le.f64 p, e, %2;\n\t"
        "@p add.s32 %0, %0, 1;\n\t}"
        : "+r"(b) : "r"(ea), "d"(x), "n"(OFF));
}

// buckets of kU values: one uint16 probe + KM predicated edge compares (KM = 0: L.kmax <= kKFast
// of them; KM < 0: a binary search over the cell's edges -- data whose mass sits in a few narrow
// far-apart clusters puts many edges into one cell)
template <int GRID, int KM>
__device__ __forceinline__ void buckets_of(const Lookup& L, const double* xs, int* b) {
    unsigned v[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
        const int cell = cell_of<GRID>(L, xs[u]);
        const unsigned ta = L.tab + 2u * (unsigned)cell;
        v[u] = lds_u16(ta);
        if (KM < 0) v[u] |= lds_u16(ta + 2u) << 16;        // + the entry of the next cell
    }
    if (KM < 0) {
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            int lo = (int)((v[u] & 0xffffu) >> 2), hi = (int)(v[u] >> 18);
            while (lo < hi) {                              // #edges <= x among the cell's candidates
                const int mid = (lo + hi) >> 1;
                if (lds_f64(L.edges + 8u * (unsigned)mid) <= xs[u]) lo = mid + 1; else hi = mid;
            }
            b[u] = lo;
        }
    } else if (KM > 0) {
        // the cell holds n = v & 3 <= KM edges: load and compare only those (a lane whose cell has
        // none issues no shared-memory request at all)
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            if (KM == 1) {
                asm volatile("{\n\t.reg .pred p;\n\t.reg .f64 e;\n\t.reg .u32 c, a;\n\t"
                    "and.b32 c, %1, 3;\n\t"
                    "shr.u32 %0, %1, 2;\n\t"
                    "setp.ne.u32 p, c, 0;\n\t"
                    "mad.lo.u32 a, %0, 8, %2;\n\t"
                    "@p ld.shared.f64 e, [a];\n\t"
                    "@p setp.le.f64 p, e, %3;\n\t"
                    "@p add.s32 %0, %0, 1;\n\t}"
                    : "=r"(b[u]) : "r"(v[u]), "r"(L.edges), "d"(xs[u]));
            } else {
                asm volatile("{\n\t.reg .pred p;\n\t.reg .f64 e;\n\t.reg .u32 c, a;\n\t"
                    "and.b32 c, %1, 3;\n\t"
                    "shr.u32 %0, %1, 2;\n\t"
                    "setp.ne.u32 p, c, 0;\n\t"
                    "mad.lo.u32 a, %0, 8, %2;\n\t"
                    "@p ld.shared.f64 e, [a];\n\t"
                    "@p setp.le.f64 p, e, %3;\n\t"
                    "@p add.s32 %0, %0, 1;\n\t"
                    "setp.gt.u32 p, c, 1;\n\t"
                    "@p ld.shared.f64 e, [a+8];\n\t"
                    "@p setp.le.f64 p, e, %3;\n\t"
                    "@p add.s32 %0, %0, 1;\n\t}"
                    : "=r"(b[u]) : "r"(v[u]), "r"(L.edges), "d"(xs[u]));
            }
        }
    } else {
        unsigned ea[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) { b[u] = (int)(v[u] >> 2); ea[u] = L.edges + 8u * (unsigned)b[u]; }
#pragma unroll 1
        for (int j = 0; j < L.kmax; ++j) {
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t.reg .f64 e;\n\t"
                    "ld.shared.f64 e, [%1];\n\t"
                    "setp.le.f64 p, e, %2;\n\t"
                    "@p add.s32 %0, %0, 1;\n\t}"
                    : "+r"(b[u]) : "r"(ea[u]), "d"(xs[u]));
                ea[u] += 8u;
            }
        }
    }
}

struct Lookup2 {
    Lookup L;
    unsigned row_bytes;   // byte stride of a bucket row of this warp's table
    unsigned cnt;         // unweighted: shared-window address of the u32 counters
};

// every x of the tile finite, every w in (eps, DBL_MAX]?  (w: on the high word -- a weight within
// 2^-20 relative of eps counts as suspicious and sends the tile down the exact path)
template <bool HAS_W>
__device__ __forceinline__ bool tile_clean(const double* xs, const double* ws) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 2 * kU; ++k) ok = ok && (fabs(xs[k]) <= DBL_MAX);
    if (HAS_W) {
        unsigned worst = 0;
#pragma unroll
        for (int k = 0; k < 2 * kU; ++k) {
            const unsigned d = (unsigned)__double2hiint(ws[k]) - 0x3CB00001u;   // hi(eps) = 0x3CB00000
            worst = d > worst ? d : worst;
        }
        ok = ok && worst <= 0x7FEFFFFFu - 0x3CB00001u;
    }
    return __all_sync(CRICK_FULL, ok);
}

// clean tile, weighted: addresses + (w, w*x), equal buckets of the lane folded
template <int GRID, int KM>
__device__ __forceinline__ bool classify_lean_w(const Lookup2& Q, const double* xs, const double* ws,
                                                unsigned* addr, double* ww, double* ss) {
    int b[kU];
    buckets_of<GRID, KM>(Q.L, xs, b);
    bool ext = false;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
        ext = ext || ((unsigned)b[u] - Q.L.ext_lo >= Q.L.ext_span);
        addr[u] = Q.L.acc + (unsigned)b[u] * Q.row_bytes;
        ww[u] = ws[u];
        ss[u] = __dmul_rn(ws[u], xs[u]);
    }
    const unsigned daddr = Q.L.acc + (unsigned)Q.L.dummy * Q.row_bytes;
#pragma unroll
    for (int i = 1; i < kU; ++i) {
#pragma unroll
        for (int j = 0; j < i; ++j) fold_if_equal(addr[i], addr[j], ww[j], ss[j], ww[i], ss[i], daddr);
    }
    return ext;
}

// unweighted: counts go to the u32 counters at once (one red per value, no fold needed), the sums
// x through the ring.  CLEAN = false: non-finite x (tdigest_add's filter, :285) is routed to the
// dummy row / dummy counter.
template <int GRID, int KM, bool CLEAN>
__device__ __forceinline__ bool classify_lean_u(const Lookup2& Q, const double* xs, unsigned* addr,
                                                double* ss) {
    int b[kU];
    buckets_of<GRID, KM>(Q.L, xs, b);
    bool ext = false;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
        double x = xs[u];
        if (!CLEAN) {
            const bool ok = fabs(x) <= DBL_MAX;
            b[u] = ok ? b[u] : Q.L.dummy;
            x = ok ? x : 0.0;
        }
        ext = ext || ((unsigned)b[u] - Q.L.ext_lo >= Q.L.ext_span);   // (the dummy row counts as "high")
        addr[u] = Q.L.acc + (unsigned)b[u] * Q.row_bytes;
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(Q.cnt + 4u * (unsigned)b[u]) : "memory");
        ss[u] = x;
    }
    const unsigned daddr = Q.L.acc + (unsigned)Q.L.dummy * Q.row_bytes;
#pragma unroll
    for (int i = 1; i < kU; ++i) {
#pragma unroll
        for (int j = 0; j < i; ++j) fold_if_equal1(addr[i], addr[j], ss[j], ss[i], daddr);
    }
    return ext;
}

__device__ __forceinline__ double lds_f64v(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
template <int N>
__device__ __forceinline__ void rmw1(const unsigned* addr, const double* ss) {
    double a[N];
#pragma unroll
    for (int u = 0; u < N; ++u) a[u] = lds_f64v(addr[u]);
#pragma unroll
    for (int u = 0; u < N; ++u) a[u] = __dadd_rn(a[u], ss[u]);
#pragma unroll
    for (int u = 0; u < N; ++u) sts_f64(addr[u], a[u]);
}

template <bool HAS_W, bool VEC, int GRID, int KM>
__device__ __forceinline__ void accumulate_stream2(const Lookup2& Q, const Ring ring,
                                                   const double* __restrict__ X,
                                                   const double* __restrict__ W, double wsc, long n,
                                                   long gw, long tw, int lane, double& vmin,
                                                   double& vmax, int* badw_flag) {
    const Lookup& L = Q.L;
    const long nfull = n / kTile;
    const long ntiles = (n + kTile - 1) / kTile;
    const long rounds = (ntiles + tw - 1) / tw;  // the same for every warp: the token needs all
    double xs[2 * kU], ws[2 * kU];
    unsigned addr[2 * kU];
    double ww[2 * kU], ss[2 * kU];
    bool badw = false;
    long t = gw;
    if (t < nfull) {
        load_tile<HAS_W, VEC>(X, W, t * kTile, lane, xs, ws);
        if (VEC && t + tw < nfull) l2_prefetch<HAS_W>(X, W, (t + tw) * kTile, lane);
    }
    for (long r = 0; r < rounds; ++r, t += tw) {
        const long tn = t + tw;
        const bool have = t < ntiles;   // warp-uniform
        const bool full = t < nfull;
        if (have && !full) load_partial<HAS_W>(X, W, t * kTile, n, lane, xs, ws);   // pads: x = NaN, w = 1
        if (have) {
            bool ext;
            if (full && tile_clean<HAS_W>(xs, ws)) {
                if (HAS_W) {
                    const bool e0 = classify_lean_w<GRID, KM>(Q, xs, ws, addr, ww, ss);
                    const bool e1 = classify_lean_w<GRID, KM>(Q, xs + kU, ws + kU, addr + kU, ww + kU, ss + kU);
                    ext = e0 || e1;
                } else {
                    const bool e0 = classify_lean_u<GRID, KM, true>(Q, xs, addr, ss);
                    const bool e1 = classify_lean_u<GRID, KM, true>(Q, xs + kU, addr + kU, ss + kU);
                    ext = e0 || e1;
                }
            } else if (HAS_W) {
                // the exact per-value filter of round 1 (dummy-row routing, weight verdict)
                bool susp = false;
                const bool e0 = classify_group<true, GRID, (KM >= 0)>(L, xs, ws, wsc, addr, ww, ss, susp);
                const bool e1 = classify_group<true, GRID, (KM >= 0)>(L, xs + kU, ws + kU, wsc, addr + kU,
                                                                      ww + kU, ss + kU, susp);
                ext = e0 || e1;
                if (__any_sync(CRICK_FULL, susp)) badw |= susp && any_illegal_weight(ws);
            } else {
                const bool e0 = classify_lean_u<GRID, KM, false>(Q, xs, addr, ss);
                const bool e1 = classify_lean_u<GRID, KM, false>(Q, xs + kU, addr + kU, ss + kU);
                ext = e0 || e1;
            }
            if (__any_sync(CRICK_FULL, ext)) {
                minmax_group<HAS_W>(L, xs, ws, wsc, vmin, vmax);
                minmax_group<HAS_W>(L, xs + kU, ws + kU, wsc, vmin, vmax);
            }
        }
        if (tn < nfull) {
            load_tile<HAS_W, VEC>(X, W, tn * kTile, lane, xs, ws);
            if (VEC && tn + tw < nfull) l2_prefetch<HAS_W>(X, W, (tn + tw) * kTile, lane);
        }
        if (!(ring.first && r == 0)) bar_sync64(ring.id_in);
        if (have) {
            if (ring.passes == 1) {
                if (HAS_W) { rmw<kU>(addr, ww, ss); rmw<kU>(addr + kU, ww + kU, ss + kU); }
                else { rmw1<kU>(addr, ss); rmw1<kU>(addr + kU, ss + kU); }
            } else {
                for (int p_ = 0; p_ < ring.passes; ++p_) {
                    if (ring.pass == p_) {
                        if (HAS_W) { rmw<kU>(addr, ww, ss); rmw<kU>(addr + kU, ww + kU, ss + kU); }
                        else { rmw1<kU>(addr, ss); rmw1<kU>(addr + kU, ss + kU); }
                    }
                    __syncwarp();
                }
            }
        }
        bar_arrive64(ring.id_out);
    }
    if (ring.first && rounds > 0) bar_sync64(ring.id_in);  // absorb the last hand-over
    if (HAS_W && __any_sync(CRICK_FULL, badw) && lane == 0) atomicOr(badw_flag, 1);
}

template <bool HAS_W, bool VEC>
__global__ void __launch_bounds__(HAS_W ? 448 : 512, 1)
k_bin_accumulate2(const double* __restrict__ X, const double* __restrict__ W, double wsc, long n,
                  ParHeader* hdr, const double* __restrict__ g_edges,
                  const unsigned short* __restrict__ g_tab, double2* __restrict__ partial,
                  double2* __restrict__ minmax, int bmax, int cols, int ntab, int accumulate_into) {
    // shared: acc[ntab][bmax + 1][cols] (double2 weighted, double unweighted) | cnt[bmax + 1 (+pad)] u32
    //         (unweighted) | edges[bmax + kEdgePad] | tab[kTabLen] uint16
    constexpr int ES = HAS_W ? 16 : 8;            // bytes per accumulator
    const int nwarps = blockDim.x >> 5;
    const int rows = bmax + 1;                    // last row = dummy
    unsigned char* s_acc = smem_raw;
    const size_t acc_bytes = (size_t)ntab * rows * cols * ES;
    const int cnt_len = HAS_W ? 0 : ((rows + 3) & ~3);
    unsigned* s_cnt = reinterpret_cast<unsigned*>(s_acc + acc_bytes);
    double* s_edges = reinterpret_cast<double*>(s_acc + acc_bytes + (size_t)cnt_len * 4);
    unsigned short* s_tab = reinterpret_cast<unsigned short*>(s_edges + bmax + kEdgePad);
    __shared__ double s_min[32], s_max[32];
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    // clearing the tables needs nothing from k_make_edges: it runs before the dependency wait
    {
        double2* z = reinterpret_cast<double2*>(s_acc);
        const int nz = (int)((acc_bytes + (size_t)cnt_len * 4) >> 4);
        for (int i = threadIdx.x; i < nz; i += blockDim.x) z[i] = make_double2(0.0, 0.0);
    }
    grid_dep_wait();
    if (blockIdx.x == 0 && threadIdx.x == 0) stamp(hdr, 13);
    const int E = hdr->E;
    const int B = E + 1;
    for (int i = threadIdx.x; i < kTabLen / 2; i += blockDim.x)
        reinterpret_cast<unsigned*>(s_tab)[i] = reinterpret_cast<const unsigned*>(g_tab)[i];
    for (int i = threadIdx.x; i < bmax + kEdgePad; i += blockDim.x)
        s_edges[i] = i < E ? g_edges[i] : INFINITY;
    __syncthreads();
    Lookup2 Q;
    Lookup& L = Q.L;
    L.tab = (unsigned)__cvta_generic_to_shared(s_tab);
    L.edges = (unsigned)__cvta_generic_to_shared(s_edges);
    const int tab_id = warp % ntab, pos = warp / ntab, rsize = nwarps / ntab;
    L.acc = (unsigned)__cvta_generic_to_shared(s_acc + (size_t)tab_id * rows * cols * ES) +
            (unsigned)ES * (unsigned)(lane & (cols - 1));
    L.row_shift = (HAS_W ? 4 : 3) + __ffs(cols) - 1;
    Q.row_bytes = (unsigned)(cols * ES);
    Q.cnt = (unsigned)__cvta_generic_to_shared(s_cnt);
    L.dummy = bmax;
    L.kbase = hdr->kbase; L.shift = hdr->shift; L.x0 = hdr->x0; L.inv = hdr->inv;
    L.E = E; L.kmax = hdr->kmax;
    L.ext_lo = E >= 4 ? 2u : 0u;
    L.ext_span = E >= 4 ? (unsigned)(E - 3) : 0u;
    Ring ring;
    ring.id_in = 1 + tab_id * rsize + pos;
    ring.id_out = 1 + tab_id * rsize + (pos + 1 == rsize ? 0 : pos + 1);
    ring.first = pos == 0;
    ring.passes = 32 / cols;
    ring.pass = lane / cols;
    double vmin = INFINITY, vmax = -INFINITY;
    const long gw = (long)blockIdx.x * nwarps + warp;
    const long tw = (long)gridDim.x * nwarps;
    const int km = L.kmax;
#define CRICK_ACC2(G, K) accumulate_stream2<HAS_W, VEC, G, K>(Q, ring, X, W, wsc, n, gw, tw, lane, vmin, vmax, &hdr->badw)
    if (hdr->grid == 1) {
        if (km <= 1) CRICK_ACC2(1, 1); else if (km == 2) CRICK_ACC2(1, 2);
        else if (km <= kKFast) CRICK_ACC2(1, 0); else CRICK_ACC2(1, -1);
    } else {
        if (km <= 1) CRICK_ACC2(0, 1); else if (km == 2) CRICK_ACC2(0, 2);
        else if (km <= kKFast) CRICK_ACC2(0, 0); else CRICK_ACC2(0, -1);
    }
#undef CRICK_ACC2
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        vmin = fmin(vmin, shfl_xor_d(vmin, d));
        vmax = fmax(vmax, shfl_xor_d(vmax, d));
    }
    if (lane == 0) { s_min[warp] = vmin; s_max[warp] = vmax; }
    __syncthreads();
    // column sums in place, fixed order (tables into table 0, then a halving tree over the columns)
    double2* row = partial + (size_t)blockIdx.x * kBPad;
    const int nrc = B * cols;
    if (HAS_W) {
        double2* a2 = reinterpret_cast<double2*>(s_acc);
        for (int tb = 1; tb < ntab; ++tb) {
            for (int i = threadIdx.x; i < nrc; i += blockDim.x) {
                double2 a = a2[i], o = a2[(size_t)tb * rows * cols + i];
                a2[i] = make_double2(a.x + o.x, a.y + o.y);
            }
        }
        __syncthreads();
        for (int dcol = cols >> 1; dcol > 0; dcol >>= 1) {
            const int sh = __ffs(dcol) - 1;
            for (int i = threadIdx.x; i < B * dcol; i += blockDim.x) {
                const int bb = i >> sh, c = i & (dcol - 1);
                double2 a = a2[bb * cols + c], o = a2[bb * cols + c + dcol];
                a2[bb * cols + c] = make_double2(a.x + o.x, a.y + o.y);
            }
            __syncthreads();
        }
        for (int bb = threadIdx.x; bb < B; bb += blockDim.x) {
            double2 a = a2[bb * cols];
            if (accumulate_into) { double2 o = row[bb]; a.x += o.x; a.y += o.y; }
            row[bb] = a;
        }
    } else {
        double* a1 = reinterpret_cast<double*>(s_acc);
        for (int i = threadIdx.x; i < nrc; i += blockDim.x) {
            double a = a1[i];
            for (int tb = 1; tb < ntab; ++tb) a += a1[(size_t)tb * rows * cols + i];   // fixed order
            a1[i] = a;
        }
        __syncthreads();
        for (int dcol = cols >> 1; dcol > 0; dcol >>= 1) {
            const int sh = __ffs(dcol) - 1;
            for (int i = threadIdx.x; i < B * dcol; i += blockDim.x) {
                const int bb = i >> sh, c = i & (dcol - 1);
                a1[bb * cols + c] += a1[bb * cols + c + dcol];
            }
            __syncthreads();
        }
        for (int bb = threadIdx.x; bb < B; bb += blockDim.x) {
            // sum w = count * w (exact for w = 1), sum w x = w * sum x
            double2 a = make_double2(__dmul_rn((double)s_cnt[bb], wsc), __dmul_rn(wsc, a1[bb * cols]));
            if (accumulate_into) { double2 o = row[bb]; a.x += o.x; a.y += o.y; }
            row[bb] = a;
        }
    }
    if (threadIdx.x == 0) {
        double mn = INFINITY, mx = -INFINITY;
        for (int wv = 0; wv < nwarps; ++wv) { mn = fmin(mn, s_min[wv]); mx = fmax(mx, s_max[wv]); }
        if (accumulate_into) { double2 o = minmax[blockIdx.x]; mn = fmin(mn, o.x); mx = fmax(mx, o.y); }
        minmax[blockIdx.x] = make_double2(mn, mx);
        if (blockIdx.x == 0) { hdr->ncta = gridDim.x; stamp(hdr, 14); }
    }
}

// warp_merge_sorted (tdigest_ref.cuh) for a whole CTA: the same phases and the same arithmetic,
// with the order-independent ones spread over all threads and the two sequential ones on
// thread 0.  (Integer weights need no special case here: the warp version's parallel scan is
// exact for them, i.e. equal to the sequential sum used below.)
//
// s_bt != nullptr: h.total does not yet include the new items; their weights are summed
// sequentially in ws.sb order (by another thread, next to the cumulative sum) and added first.
// SEQ_SUM = false (the one-pass merge of gathered records, which promises no reference bits): the
// cumulative weights come from a block scan (fixed association: deterministic) instead of the
// single-thread fl chain -- 1000 dependent DADDs are 9 us of a 30 us kernel.
template <bool SEQ_SUM = true>
__device__ __forceinline__ void block_merge_sorted(const WarpWS& ws, double comp, Hdr& h, int m,
                                                   int cap, int* s_nc, double* s_bt) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int n = h.n;
    const int T = n + m;
    // ---- phase 2a: prefix max of centroid means -> kend[0..n)
    if (tid < 32) {
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
    __syncthreads();
    // ---- phase 2b: merged positions
    for (int i = tid; i < m; i += blockDim.x) {
        double2 it = ws.sb[i];
        int lo = 0, hi = n;  // first j with PM[j] > b
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (ws.kend[mid] > it.x) hi = mid; else lo = mid + 1;
        }
        ws.items[i + lo] = it;
    }
    for (int j = tid; j < n; j += blockDim.x) {
        double p = ws.kend[j];
        int lo = 0, hi = m;  // number of sorted buffer items strictly below PM[j]
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (ws.sb[mid].x < p) lo = mid + 1; else hi = mid;
        }
        ws.items[j + lo] = ws.cen[j];
    }
    __syncthreads();
    // ---- phase 3: cumulative weights (into kend), then k-scale per item
    if (SEQ_SUM) {
        if (tid == 0) serial_cumsum(ws.items, ws.kend, T);
    } else {
        __shared__ double s_scan_d[32];
        __shared__ double s_carry;
        if (tid == 0) s_carry = 0.0;
        __syncthreads();
        for (int base = 0; base < T; base += blockDim.x) {
            const int i = base + tid;
            const double incl = block_scan_incl<double>(i < T ? ws.items[i].y : 0.0, s_scan_d);
            const double carry = s_carry;
            if (i < T) ws.kend[i] = carry + incl;
            __syncthreads();
            if (tid == (int)blockDim.x - 1) s_carry = carry + incl;
            __syncthreads();
        }
    }
    if (s_bt != nullptr && tid == 32) {
        double bt = 0.0;
        int i = 0;
        for (; i + 8 <= m; i += 8) {
            double y[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) y[k] = ws.sb[i + k].y;
#pragma unroll
            for (int k = 0; k < 8; ++k) bt = __dadd_rn(bt, y[k]);
        }
        for (; i < m; ++i) bt = __dadd_rn(bt, ws.sb[i].y);
        *s_bt = bt;
    }
    __syncthreads();
    if (s_bt != nullptr) h.total = __dadd_rn(h.total, *s_bt);
    for (int i = tid; i < T; i += blockDim.x)
        ws.kend[i] = kscale(comp, __ddiv_rn(ws.kend[i], h.total));
    __syncthreads();
    // ---- phase 4: boundary chain (serial_chain's recurrence, tdigest_ref.cuh).  While
    // k1 = kend[p] the next centroid opens at nb(p) = first i >= p + 2 with
    // !(fl(kend_i - k1) <= 1), and then k1 = kend[nb(p) - 1]: every nb(p) is found in parallel
    // (a short linear scan, the same test in the same order, so no monotonicity is assumed) and
    // one thread follows the links.  p = -1 stands for the initial k1 = 0.  nb lives in ws.sb,
    // which is dead after phase 2b (4 ints per slot >= T + 1).
    int* nb = reinterpret_cast<int*>(ws.sb);
    for (int p = tid - 1; p < T - 1; p += blockDim.x) {
        const double k1 = p < 0 ? 0.0 : ws.kend[p];
        int i = p + 2;
        while (i < T && __dadd_rn(ws.kend[i], -k1) <= 1.0) ++i;
        nb[p + 1] = i;
    }
    // Node q (0 <= q < T) stands for "a centroid starts at item q" (q = 0: the first one, with
    // k1 = 0) and points to the next start nb[q]; T is the sink.  The starts are the nodes on
    // the path from node 0: marked by pointer jumping (after r rounds: every node within 2^r - 1
    // steps is marked and cur = nb^(2^r)) instead of one thread walking ~1.3 c dependent links.
    // Mark = bit 31 of the node's entry; two arrays (cur / nxt) so that no entry is read and
    // written in the same phase.  (ws.sb holds 4 ints per slot >= 2 (T + 1).)
    constexpr unsigned kMark = 0x80000000u;
    unsigned* cur = reinterpret_cast<unsigned*>(nb);
    unsigned* nxt = cur + (T + 1);
    if (tid == 0) { cur[T] = (unsigned)T; nxt[T] = (unsigned)T; }
    __syncthreads();
    if (tid == 0) cur[0] |= kMark;
    __syncthreads();
    for (int span = 1; span < T; span <<= 1) {
        for (int q = tid; q < T; q += blockDim.x) {
            const unsigned v = cur[q];
            const unsigned tg = v & ~kMark;
            if ((v & kMark) && tg < (unsigned)T) atomicOr(&cur[tg], kMark);
        }
        __syncthreads();
        for (int q = tid; q < T; q += blockDim.x) {
            const unsigned v = cur[q];
            nxt[q] = (v & kMark) | (cur[v & ~kMark] & ~kMark);
        }
        __syncthreads();
        unsigned* t_ = cur; cur = nxt; nxt = t_;
    }
    // the marked nodes in increasing order are start[0], start[1], ...; at most `cap` of them
    // (the sequential chain stops opening centroids once it has `cap`)
    {
        __shared__ int s_scan_w[32];
        __shared__ int s_tot;
        int run = 0;
        for (int base = 0; base < T; base += blockDim.x) {
            const int q = base + tid;
            const int f = (q < T && (cur[q] & kMark)) ? 1 : 0;
            const int incl = block_scan_incl<int>(f, s_scan_w);
            const int at = run + incl - f;
            if (f && at < cap) ws.start[at] = q;
            if (tid == (int)blockDim.x - 1) s_tot = incl;
            __syncthreads();
            run += s_tot;
            __syncthreads();
        }
        if (tid == 0) {
            const int nc_ = run < cap ? run : cap;
            ws.start[nc_] = T;
            *s_nc = nc_;
        }
    }
    __syncthreads();
    const int nc = *s_nc;
    // ---- phase 5: fold runs
    for (int c = tid; c < nc; c += blockDim.x) {
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
    __syncthreads();
    h.n = nc;
}

// 5. finalize: bucket sums -> sorted weighted points -> greedy k-scale merge.
//
// finalize_points: phases a-b -- the per-CTA partials reduced in fixed order and turned into
// <= mmax weighted points in sorted order (sb[0..m)), plus the exact (min, max) of the batch.
// Returns m (uniform over the CTA).  1024 threads.
__device__ __forceinline__ int finalize_points(const ParHeader* hdr, const double* __restrict__ edges,
                                               const double2* __restrict__ partial,
                                               const double2* __restrict__ minmax, double2* sb,
                                               int mmax, double comp, double& mn_out, double& mx_out) {
    __shared__ int s_scan[1024 + 1];
    __shared__ double s_mm[2];
    __shared__ int s_m;
    __shared__ double s_W;
    const int tid = threadIdx.x;
    const int E = hdr->E;
    const int B = E + 1;
    const int ncta = hdr->ncta;
    // a. reduce partials over CTAs and the batch total weight.  Bucket b is summed by NS
    // threads, thread (s, b) taking rows s, s + NS, ... (loads independent, 16 in flight), then
    // the NS slice sums are added in slice order: a fixed order, so the result is deterministic.
    __shared__ double2 s_part[1024];
    __shared__ double s_bw[1024];
    __shared__ double s_before[1024];
    __shared__ double s_wmn[32], s_wmx[32];
    const int b = tid;  // B <= 1024 (ov*c <= kEMax): one thread per bucket
    const int NS = 1024 / B > 0 ? 1024 / B : 1;
    {
        const int sl = tid / B, bb = tid - sl * B;
        double w_ = 0.0, s_ = 0.0;
        if (sl < NS) {
            int cta = sl;
            for (; cta + 15 * NS < ncta; cta += 16 * NS) {
                double2 a[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) a[k] = partial[(size_t)(cta + k * NS) * kBPad + bb];
#pragma unroll
                for (int k = 0; k < 16; ++k) { w_ += a[k].x; s_ += a[k].y; }
            }
            for (; cta < ncta; cta += NS) {
                double2 a = partial[(size_t)cta * kBPad + bb];
                w_ += a.x; s_ += a.y;
            }
        }
        s_part[tid] = make_double2(w_, s_);
        double mn_ = INFINITY, mx_ = -INFINITY;
        for (int cta = tid; cta < ncta; cta += 1024) {
            mn_ = fmin(mn_, minmax[cta].x); mx_ = fmax(mx_, minmax[cta].y);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn_ = fmin(mn_, shfl_xor_d(mn_, o)); mx_ = fmax(mx_, shfl_xor_d(mx_, o));
        }
        if ((tid & 31) == 0) { s_wmn[tid >> 5] = mn_; s_wmx[tid >> 5] = mx_; }
    }
    s_scan[tid] = 0;
    __syncthreads();
    if (tid == 0) stamp(const_cast<ParHeader*>(hdr), 15);
    double bw = 0.0, bs = 0.0;
    if (b < B) {
        for (int sl = 0; sl < NS; ++sl) { bw += s_part[sl * B + b].x; bs += s_part[sl * B + b].y; }
        s_bw[b] = bw;
    }
    if (tid == 32) {
        double mn_ = INFINITY, mx_ = -INFINITY;
        for (int k = 0; k < 32; ++k) { mn_ = fmin(mn_, s_wmn[k]); mx_ = fmax(mx_, s_wmx[k]); }
        s_mm[0] = mn_; s_mm[1] = mx_;
    }
    __syncthreads();
    const double mn = s_mm[0], mx = s_mm[1];
    // b. pieces per bucket.  A single-valued bucket [v, nextafter(v)) heavier than
    // half a k-unit is split into equal pieces with the same mean, so that the
    // digest keeps several centroids on a heavy duplicate like the reference does.
    const bool single = b < B && bw > 0.0 && b >= 1 && b <= E - 1 && edges[b] == next_up(edges[b - 1]);
    int np = (b < B && bw > 0.0) ? 1 : 0;
    if (__syncthreads_or(single)) {
        // batch weight and the cumulative weight before each bucket, in bucket order
        // (sequential: deterministic); only needed when there is such a bucket
        if (tid == 0) {
            double run = 0.0;
            int i = 0;
            for (; i + 8 <= B; i += 8) {        // loads batched, only the adds are serial
                double y[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) y[k] = s_bw[i + k];
#pragma unroll
                for (int k = 0; k < 8; ++k) { s_before[i + k] = run; run += y[k]; }
            }
            for (; i < B; ++i) { s_before[i] = run; run += s_bw[i]; }
            s_W = run;
        }
        __syncthreads();
        if (single) {
            const double Wb = s_W;          // weight of this batch
            const double before = s_before[b];
            double q0 = before / Wb, q1 = (before + bw) / Wb;
            double span = kscale(comp, q1) - kscale(comp, q0);
            int p = (int)ceil(span * 2.0);
            np = p < 1 ? 1 : p;
        }
    }
    // exclusive scan over 1024 buckets
    __shared__ int s_wi[32];
    const int incl = block_scan_incl<int>(np, s_wi);
    int pos = incl - np;
    if (tid == 1023) s_scan[1023] = incl;
    __syncthreads();
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
            sb[at + k] = make_double2(mean, wk);
        }
    };
    if (b < B) emit(b, bw, bs, np, pos);
    if (tid == 0) {
        int m = total1024;
        s_m = m > mmax ? mmax : m;
    }
    __syncthreads();
    mn_out = mn; mx_out = mx;
    return s_m;
}

__global__ void __launch_bounds__(1024) k_finalize(BatchView v, long d, ParHeader* hdr,
                                                   const double* __restrict__ edges,
                                                   const double2* __restrict__ partial,
                                                   const double2* __restrict__ minmax,
                                                   unsigned char* gws, int mmax, int use_smem,
                                                   int honor_badw) {
    const int tid = threadIdx.x;
    grid_dep_wait();          // the partials of k_bin_accumulate
    if (tid == 0) stamp(hdr, 0);
    // validated update (tdigest.pyx:304-305): the digest is only touched when every weight
    // passed; the host reads the same flag after this kernel has been enqueued
    if (honor_badw && hdr->badw) {
        if (tid == 0) hdr->badw_sticky = 1;   // read (and cleared) by parallel_take_bad_weights
        return;
    }
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
    double mn, mx;
    const int m = finalize_points(hdr, edges, partial, minmax, ws.sb, mmax, v.comp, mn, mx);
    if (tid == 0) stamp(hdr, 2);
    // c. merge into the digest (tdigest_flush phases 2-5 with the bucket points as the buffer)
    {
        __shared__ double s_bt;
        __shared__ int s_nc;
        Hdr h;
        const DevHdr dh = v.hdr[d];
        h.vmin = dh.vmin; h.vmax = dh.vmax; h.total = dh.total; h.btotal = 0.0; h.n = dh.n; h.bn = 0;
        const double2* cg = v.cen + d * (long)cap;
        for (int i = tid; i < h.n; i += blockDim.x) ws.cen[i] = cg[i];
        if (m > 0) {
            if (h.vmin > mn) h.vmin = mn;
            if (h.vmax < mx) h.vmax = mx;
            // the batch weight (sum of the pieces actually merged, in item order) is summed
            // inside the merge, next to the cumulative weights; it is added to h.total there
            block_merge_sorted(ws, v.comp, h, m, cap, &s_nc, &s_bt);
            double2* cgo = v.cen + d * (long)cap;
            for (int i = tid; i < h.n; i += blockDim.x) cgo[i] = ws.cen[i];
            if (tid == 0) {
                DevHdr o = dh;
                o.vmin = h.vmin; o.vmax = h.vmax; o.total = h.total; o.btotal = 0.0;
                o.n = h.n; o.bn = 0;
                v.hdr[d] = o;
            }
        }
        if (tid == 0) stamp(hdr, 5);
    }
}

// 5b. k_exchange_merge -- the tail of a multi-GPU step as ONE kernel over NVLink peer memory.
//
// A stream sharded over G GPUs (BASELINE config 2) ends every update with "all ranks' partial
// results -> one digest on every rank".  Through NCCL that is finalize -> pack -> all-gather ->
// reset -> merge: five launches and a collective whose latency (not bandwidth: 3 KB per rank) sits
// on the critical path.  Here the kernel that reduces the rank's bucket partials to <= 2 ov c
// weighted points
//   * stores them, with (count, weight, min, max), straight into slot [step & 1][rank] of EVERY
//     rank's exchange buffer (peer pointers obtained through CUDA IPC: plain st.global over
//     NVLink / NVSwitch), then publishes the step number in flags[rank] of every rank with a
//     system-scope release store;
//   * waits (system-scope acquire loads on its OWN flags) until every rank's points of this step
//     have arrived;
//   * ranks the G sorted point runs against each other in shared memory and folds the union into
//     the digest in one greedy k-scale pass (block_merge_sorted) -- every rank computes the same
//     digest from the same records in the same order.
// Two slots per source rank: a rank can only be one step ahead of the slowest reader (its next
// merge needs that reader's next flag).  The wait gives up after 4 s (flag in ParHeader).
constexpr int kP2PMaxWorld = 16;
constexpr int kP2PFlagDoubles = 32;          // flags[16] u64, padded to 256 B
struct P2PView {
    int rank, world;
    int rec_doubles;                          // 4 + 2 * mmax
    int empty;                                // this rank has no data in this step
    unsigned long long step;
    double* buf[kP2PMaxWorld];                // exchange buffer of every rank (buf[rank] = own)
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(1024) k_exchange_merge(BatchView v, long d, ParHeader* hdr,
                                                         const double* __restrict__ edges,
                                                         const double2* __restrict__ partial,
                                                         const double2* __restrict__ minmax, int mmax,
                                                         int pmax, P2PView pv, int honor_badw) {
    __shared__ int s_off[kP2PMaxWorld + 1];
    __shared__ double s_tot[kP2PMaxWorld], s_rmn[kP2PMaxWorld], s_rmx[kP2PMaxWorld];
    __shared__ double s_scan_d[32];
    __shared__ int s_nc, s_ng, s_timeout;
    const int tid = threadIdx.x;
    const int cap = v.cap;
    // phase timestamps (globaltimer ns) for tests/tools/sharded_check.py: [128 .. 176) of the header block
    unsigned long long* dbg = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(hdr) + 128);
    if (tid == 0) dbg[0] = global_ns();
    grid_dep_wait();          // the partials of k_bin_accumulate
    if (tid == 0) dbg[1] = global_ns();
    // carve: cen[cap] | sb[pmax] | items[cap + pmax] | kend | start; the staged keys of the runs
    // alias items (dead until the merge)
    WarpWS ws;
    {
        size_t t = (size_t)cap + pmax;
        ws.cen = reinterpret_cast<double2*>(smem_raw);
        ws.sb = ws.cen + cap;
        ws.buf = ws.sb;
        ws.items = ws.sb + pmax;
        ws.kend = reinterpret_cast<double*>(ws.items + t);
        ws.start = reinterpret_cast<int*>(ws.kend + t);
    }
    double* keys = reinterpret_cast<double*>(ws.items);
    // ---- 1. this rank's points
    int m = 0;
    double mn = INFINITY, mx = -INFINITY;
    const bool refused = honor_badw && !pv.empty && hdr->badw;     // (uniform)
    if (refused && tid == 0) hdr->badw_sticky = 1;
    if (!pv.empty && !refused) m = finalize_points(hdr, edges, partial, minmax, ws.sb, mmax, v.comp, mn, mx);
    // total weight of the points (block scan: fixed association)
    double wsum = 0.0;
    {
        const double incl = block_scan_incl<double>(tid < m ? ws.sb[tid].y : 0.0, s_scan_d);
        __shared__ double s_last;
        if (tid == 1023) s_last = incl;
        __syncthreads();
        wsum = s_last;           // m <= mmax <= 1024 for c <= 100 ... larger m: second pass below
        for (int base = 1024; base < m; base += 1024) {
            const double inc2 = block_scan_incl<double>(base + tid < m ? ws.sb[base + tid].y : 0.0, s_scan_d);
            if (tid == 1023) s_last = inc2;
            __syncthreads();
            wsum += s_last;
            __syncthreads();
        }
    }
    // ---- 2. publish: points into slot [step & 1][rank] of every rank, then the step number
    const size_t slot = (size_t)kP2PFlagDoubles + ((size_t)(pv.step & 1ull) * pv.world + pv.rank) * pv.rec_doubles;
    for (int r = 0; r < pv.world; ++r) {
        double* dst = pv.buf[r] + slot;
        if (tid == 0) { dst[0] = (double)m; dst[1] = wsum; dst[2] = mn; dst[3] = mx; }
        double2* dp = reinterpret_cast<double2*>(dst + 4);
        for (int i = tid; i < m; i += blockDim.x) dp[i] = ws.sb[i];
    }
    if (tid == 0) dbg[2] = global_ns();
    __threadfence_system();
    __syncthreads();
    if (tid < pv.world)
        st_release_sys(reinterpret_cast<unsigned long long*>(pv.buf[tid]) + pv.rank, pv.step);
    if (tid == 0) dbg[3] = global_ns();
    // ---- 3. wait for every rank's points of this step
    if (tid == 0) s_timeout = 0;
    __syncthreads();
    if (tid < pv.world) {
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(pv.buf[pv.rank]) + tid;
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(f) < pv.step) {
            __nanosleep(64);
            if (global_ns() - t0 > 4000000000ull) { s_timeout = 1; break; }
        }
    }
    __syncthreads();
    if (tid == 0) dbg[4] = global_ns();
    if (s_timeout) {                            // a rank never arrived: leave the digest alone
        if (tid == 0) hdr->p2p_timeout = 1;
        return;
    }
    // ---- 4. merge the G sorted runs into the digest
    const double* own = pv.buf[pv.rank] + (size_t)kP2PFlagDoubles + (size_t)(pv.step & 1ull) * pv.world * pv.rec_doubles;
    Hdr h;
    const DevHdr dh = v.hdr[d];
    h.vmin = dh.vmin; h.vmax = dh.vmax; h.total = dh.total; h.btotal = 0.0; h.n = dh.n; h.bn = 0;
    {
        const double2* cg = v.cen + d * (long)cap;
        for (int i = tid; i < h.n; i += blockDim.x) ws.cen[i] = cg[i];
    }
    bool any = false;
    int r0 = 0;
    while (r0 < pv.world) {                     // groups of records whose points fit the workspace
        if (tid == 0) {
            int run = 0, r = r0;
            for (; r < pv.world; ++r) {
                const double* rec = own + (size_t)r * pv.rec_doubles;
                int n = (int)__ldcg(rec);
                n = n < 0 ? 0 : (n > mmax ? mmax : n);
                if (run + n > pmax && r > r0) break;
                s_off[r - r0] = run;
                run += n;
                s_tot[r - r0] = __ldcg(rec + 1); s_rmn[r - r0] = __ldcg(rec + 2); s_rmx[r - r0] = __ldcg(rec + 3);
            }
            s_off[r - r0] = run;
            s_ng = r - r0;                      // records in this group
        }
        __syncthreads();
        const int ng = s_ng;
        const int T = s_off[ng];
        for (int q = 0; q < ng; ++q) {          // stage the keys (means) of every run
            const double2* pts = reinterpret_cast<const double2*>(own + (size_t)(r0 + q) * pv.rec_doubles + 4);
            const int n = s_off[q + 1] - s_off[q];
            for (int i = tid; i < n; i += blockDim.x) keys[s_off[q] + i] = __ldcg(&pts[i].x);
        }
        __syncthreads();
        for (int e = tid; e < T; e += blockDim.x) {
            int q = 0;
            while (s_off[q + 1] <= e) ++q;
            const double k = keys[e];
            int rank = e - s_off[q];
            for (int o = 0; o < ng; ++o) {
                if (o == q) continue;
                int lo = s_off[o], hi = s_off[o + 1];
                if (o < q) { while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] <= k) lo = mid + 1; else hi = mid; } }
                else { while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] < k) lo = mid + 1; else hi = mid; } }
                rank += lo - s_off[o];
            }
            const double2* pts = reinterpret_cast<const double2*>(own + (size_t)(r0 + q) * pv.rec_doubles + 4);
            ws.sb[rank] = make_double2(k, __ldcg(&pts[e - s_off[q]].y));
        }
        __syncthreads();
        if (T > 0) {                            // uniform
            for (int q = 0; q < ng; ++q) {      // every thread the same fl chain
                if (s_off[q + 1] > s_off[q]) {
                    h.total = __dadd_rn(h.total, s_tot[q]);
                    if (h.vmin > s_rmn[q]) h.vmin = s_rmn[q];
                    if (h.vmax < s_rmx[q]) h.vmax = s_rmx[q];
                }
            }
            block_merge_sorted<false>(ws, v.comp, h, T, cap, &s_nc, nullptr);
            any = true;
        }
        __syncthreads();
        r0 += ng;
    }
    if (any) {
        double2* cgo = v.cen + d * (long)cap;
        for (int i = tid; i < h.n; i += blockDim.x) cgo[i] = ws.cen[i];
        if (tid == 0) {
            DevHdr o = dh;
            o.vmin = h.vmin; o.vmax = h.vmax; o.total = h.total; o.btotal = 0.0;
            o.n = h.n; o.bn = 0;
            v.hdr[d] = o;
        }
    }
    if (tid == 0) dbg[5] = global_ns();
}

// 6. bulk merge of gathered records (multi-GPU tail) -----------------------------------------
// All centroids of the records are sorted by mean (ties: record order) and go through ONE
// greedy k-scale pass -- the digest the reference would build had it been handed the union in a
// single flush.  tdigest_merge's own semantics (through the 42-slot buffer, 23 flushes for 8
// records) are k_merge_records in tdigest_kernels.cu; this one is 5x shorter on the critical
// path of every multi-GPU step.  One CTA; P = power of two >= total centroid count.
__global__ void __launch_bounds__(1024) k_merge_records_bulk(BatchView v, long d,
                                                             const double* __restrict__ recs,
                                                             int nrec, int rec_doubles, int P, int fresh) {
    __shared__ int s_off[64 + 1];
    __shared__ double s_tot, s_mn, s_mx;
    const int tid = threadIdx.x;
    const int cap = v.cap;
    // carve: key[P] | wgt[P] | ord[P] (sort arrays) then cen[cap] | sb[P] | items[cap+P] | kend | start
    double* key = reinterpret_cast<double*>(smem_raw);
    double* wgt = key + P;
    int* ord = reinterpret_cast<int*>(wgt + P);
    unsigned char* wsbase = reinterpret_cast<unsigned char*>(ord + P);
    wsbase = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(wsbase) + 15) & ~(uintptr_t)15);
    WarpWS ws;
    {
        size_t t = (size_t)cap + P;
        ws.cen = reinterpret_cast<double2*>(wsbase);
        ws.sb = ws.cen + cap;
        ws.buf = ws.sb;
        ws.items = ws.sb + P;
        ws.kend = reinterpret_cast<double*>(ws.items + t);
        ws.start = reinterpret_cast<int*>(ws.kend + t);
    }
    if (tid == 0) {
        int run = 0;
        double tot = 0.0, mn = INFINITY, mx = -INFINITY;
        for (int r = 0; r < nrec; ++r) {
            const double* rec = recs + (size_t)r * rec_doubles;
            s_off[r] = run;
            if (rec[1] != 0.0) {                 // empty digests are skipped (:596)
                run += (int)rec[0];
                tot = __dadd_rn(tot, rec[1]);
                mn = fmin(mn, rec[2]); mx = fmax(mx, rec[3]);
            }
        }
        s_off[nrec] = run;
        s_tot = tot; s_mn = mn; s_mx = mx;
    }
    __syncthreads();
    const int m = s_off[nrec];
    const bool by_rank = nrec <= 16;             // the records are sorted runs: rank them against each other
    if (!by_rank) {
        for (int i = tid; i < P; i += blockDim.x) { key[i] = INFINITY; wgt[i] = 0.0; ord[i] = 0x7fffffff; }
        __syncthreads();
    }
    for (int r = 0; r < nrec; ++r) {
        const double* rec = recs + (size_t)r * rec_doubles;
        const int n = s_off[r + 1] - s_off[r];
        for (int i = tid; i < n; i += blockDim.x) {
            key[s_off[r] + i] = rec[4 + 2 * i];
            wgt[s_off[r] + i] = rec[5 + 2 * i];
            ord[s_off[r] + i] = s_off[r] + i;
        }
    }
    __syncthreads();
    if (by_rank) {
        // position of element (r, i) in the sorted union = i + #(elements of earlier records <= key)
        // + #(elements of later records < key): ties keep record order, like the bitonic path's
        // (key, ord) comparison.  A centroid array is sorted by mean (tdigest_flush), so each count
        // is one binary search in shared memory.
        for (int e = tid; e < m; e += blockDim.x) {
            int r = 0;
            while (s_off[r + 1] <= e) ++r;
            const double k = key[e];
            int rank = e - s_off[r];
            for (int q = 0; q < nrec; ++q) {
                if (q == r) continue;
                int lo = s_off[q], hi = s_off[q + 1];
                if (q < r) { while (lo < hi) { const int mid = (lo + hi) >> 1; if (key[mid] <= k) lo = mid + 1; else hi = mid; } }
                else { while (lo < hi) { const int mid = (lo + hi) >> 1; if (key[mid] < k) lo = mid + 1; else hi = mid; } }
                rank += lo - s_off[q];
            }
            ws.sb[rank] = make_double2(k, wgt[e]);
        }
    } else {
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < P / 2; t += blockDim.x) {
                    int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    int q = i | j;
                    bool up = (i & k) == 0;
                    bool q_first = key[q] < key[i] || (key[q] == key[i] && ord[q] < ord[i]);
                    if (q_first == up) {
                        double t1 = key[i]; key[i] = key[q]; key[q] = t1;
                        t1 = wgt[i]; wgt[i] = wgt[q]; wgt[q] = t1;
                        int t2 = ord[i]; ord[i] = ord[q]; ord[q] = t2;
                    }
                }
                __syncthreads();
            }
        }
        for (int i = tid; i < m; i += blockDim.x) ws.sb[i] = make_double2(key[i], wgt[i]);
    }
    __syncthreads();
    if (m > 0) {                                  // uniform
        __shared__ int s_nc;
        Hdr h;
        DevHdr dh;
        if (fresh) {                              // the state tdigest_new leaves (:63-92): no reset launch
            dh.vmin = DBL_MAX; dh.vmax = -DBL_MAX; dh.total = 0.0; dh.btotal = 0.0;
            dh.n = 0; dh.bn = 0; dh.pad0 = 0; dh.pad1 = 0;
        } else {
            dh = v.hdr[d];
        }
        h.vmin = dh.vmin; h.vmax = dh.vmax; h.total = dh.total; h.btotal = 0.0; h.n = dh.n; h.bn = 0;
        const double2* cg = v.cen + d * (long)cap;
        for (int i = tid; i < h.n; i += blockDim.x) ws.cen[i] = cg[i];
        if (h.vmin > s_mn) h.vmin = s_mn;
        if (h.vmax < s_mx) h.vmax = s_mx;
        h.total = __dadd_rn(h.total, s_tot);
        block_merge_sorted<false>(ws, v.comp, h, m, cap, &s_nc, nullptr);   // starts with a barrier
        double2* cgo = v.cen + d * (long)cap;
        for (int i = tid; i < h.n; i += blockDim.x) cgo[i] = ws.cen[i];
        if (tid == 0) {
            DevHdr o = dh;
            o.vmin = h.vmin; o.vmax = h.vmax; o.total = h.total; o.btotal = 0.0;
            o.n = h.n; o.bn = 0;
            v.hdr[d] = o;
        }
    } else if (fresh && tid == 0) {
        DevHdr o;
        o.vmin = DBL_MAX; o.vmax = -DBL_MAX; o.total = 0.0; o.btotal = 0.0; o.n = 0; o.bn = 0; o.pad0 = 0; o.pad1 = 0;
        v.hdr[d] = o;
    }
}

// returns cudaErrorInvalidValue when the union does not fit one CTA's shared memory (the
// caller then uses the sequential k_merge_records)
cudaError_t launch_merge_records_bulk(const BatchView& v, long d, const double* recs, int nrec,
                                      cudaStream_t st, bool fresh) {
    if (nrec <= 0) return fresh ? launch_init_hdr(v.hdr + d, 1, st) : cudaSuccess;
    if (nrec > 64) return cudaErrorInvalidValue;
    int P = 1;
    while (P < nrec * v.cap + 2) P <<= 1;   // (+2: block_merge_sorted keeps 2 (T + 1) ints in ws.sb)
    size_t t = (size_t)v.cap + P;
    size_t smem = (size_t)P * 20 + 16 + (size_t)v.cap * 16 + (size_t)P * 16 + t * 16 + t * 8 + (t + 1) * 4 + 16;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = need_smem(k_merge_records_bulk, smem);
    if (e != cudaSuccess) return e;
    count_launch();
    k_merge_records_bulk<<<1, 1024, smem, st>>>(v, d, recs, nrec, 4 + 2 * v.cap, P, fresh ? 1 : 0);
    return cudaGetLastError();
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
    long S = kSMax;
    while (S > kSortChunk && S * 4 > n) S >>= 1;
    return S;
}

static bool tail_sampling_enabled() {
    static const bool on = [] { const char* e = getenv("CRICK_TAIL_SAMPLE"); return !(e && e[0] == '0'); }();
    return on;
}

static cudaError_t edges_from_sample(const BatchView& v, ParLayout& L, long S, const SampleSrc* src,
                                     cudaStream_t st) {
    const double *skeys, *svals;
    cudaError_t e = sort_sample(L.sx, L.sw, L.sx2, L.sw2, S, src, st, &skeys, &svals);
    if (e != cudaSuccess) return e;
    const size_t smem = (size_t)(S + 1024) * 8 + 2 * kTabBytes;   // cumulative sample weights (padded) + 2 candidate tables
    e = need_smem(k_make_edges, smem);
    if (e != cudaSuccess) return e;
    count_launch();
    const bool tails = src != nullptr && src->tmin != nullptr;
    return launch_dep(k_make_edges, dim3(1), dim3(1024), smem, st, L.hdr, skeys, svals, L.edges, L.tab,
                      v.comp, ov_for(v.comp), S, tails ? (const double2*)src->tmin : (const double2*)nullptr,
                      tails ? (const double2*)src->tmax : (const double2*)nullptr);
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
    return edges_from_sample(v, L, S, nullptr, st);
}

// Shape of the hot kernel for a bucket capacity.  Two tables, each with its own token ring of
// 7 warps (6 with fused moments), tiles prefetched into L2; C = 32 columns per table when they
// fit (c <= 100), else 16 / 8 / 4 with 32 / C lane passes per read-modify-write.  Two tables of
// C columns beat one table of 2C columns + TMA stages (same shared memory, r1_f vs r1_h
// profiles); the staged single-table shape stays selectable with CRICK_ACC_T=1.
struct AccCfg { int cols, ntab, nw, stage; size_t smem; };
static AccCfg acc_cfg(int bmax, bool stats) {
    const size_t budget = 230000;                       // of 232448 B opt-in, ~2 KB are static
    const size_t fixed = kTabBytes + sizeof(double) * (size_t)(bmax + kEdgePad) + 16 * 8;
    const size_t col = (size_t)(bmax + 1) * 16;         // one column incl. the dummy row
    AccCfg c;
    int want_tab = 2;
    if (const char* e = getenv("CRICK_ACC_T")) want_tab = atoi(e) == 1 ? 1 : 2;   // tuning hooks
    c.ntab = want_tab;
    c.stage = want_tab == 1;
    c.nw = want_tab == 1 ? 12 : (stats ? 12 : 14);
    if (const char* e = getenv("CRICK_ACC_W")) {
        int v = atoi(e);
        if (v >= c.ntab && v <= (stats ? 12 : 14)) c.nw = v - v % c.ntab;
    }
    const size_t stage_bytes = c.stage ? (size_t)c.nw * kStageBytes : 0;
    c.cols = 32;
    while (c.cols > 1 && fixed + stage_bytes + (size_t)c.ntab * c.cols * col > budget) c.cols >>= 1;
    c.smem = (fixed + stage_bytes + (size_t)c.ntab * c.cols * col + 15) & ~(size_t)15;
    return c;
}

// Shape of k_bin_accumulate2.  Weighted: two double2 tables, rings of 7 (14 warps), as in round 1.
// Unweighted: 8-byte rows -> three tables with rings of 5 (15 warps: a ring hop needs a named
// barrier of its own and there are 15 besides __syncthreads'), or CRICK_ACC2_T=4: four tables with
// rings of 3.  C = 32 columns when they fit, else 16 / 8 / 4.
static AccCfg acc_cfg2(int bmax, bool has_w) {
    const size_t budget = 230000;
    const size_t es = has_w ? 16 : 8;
    const size_t fixed = kTabBytes + sizeof(double) * (size_t)(bmax + kEdgePad) +
                         (has_w ? 0 : (size_t)((bmax + 1 + 3) & ~3) * 4);
    const size_t col = (size_t)(bmax + 1) * es;
    AccCfg c;
    c.stage = 0;
    c.ntab = has_w ? 2 : 3;
    if (const char* e = getenv("CRICK_ACC2_T")) { int v = atoi(e); if (v >= 1 && v <= 5) c.ntab = v; }
    int per = 15 / c.ntab;                                 // warps per ring
    if (has_w && per > 7) per = 7;                         // launch bounds: 448 threads
    if (const char* e = getenv("CRICK_ACC2_R")) { int v = atoi(e); if (v >= 1 && v <= per) per = v; }
    c.nw = per * c.ntab;
    c.cols = 32;
    while (c.cols > 2 && fixed + (size_t)c.ntab * c.cols * col > budget) c.cols >>= 1;
    c.smem = (fixed + (size_t)c.ntab * c.cols * col + 15) & ~(size_t)15;
    return c;
}
static bool acc2_enabled() {
    static const bool on = [] { const char* e = getenv("CRICK_ACC_V"); return !(e && e[0] == '1'); }();
    return on;
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
                                   int accumulate_into, StatsPart* stats_parts, long idx_base) {
    ParLayout L = carve(scratch, v.comp, sm_count);
    const int bmax = bmax_for(v.comp);
    // the round-2 kernel unless fused moments or the TMA-staged shape are asked for
    const bool v2 = stats_parts == nullptr && acc2_enabled() && !acc_cfg(bmax, false).stage;
    const AccCfg cfg = v2 ? acc_cfg2(bmax, W != nullptr) : acc_cfg(bmax, stats_parts != nullptr);
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
#define CRICK_LAUNCH_ACC_S(HW, VC, SG, SS)                                                        \
    do {                                                                                          \
        e = need_smem(k_bin_accumulate<HW, VC, SG, SS>, smem);                                    \
        if (e != cudaSuccess) return e;                                                           \
        count_launch();                                                                           \
        e = launch_dep(k_bin_accumulate<HW, VC, SG, SS>, dim3(grid), dim3(block), smem, st, X, W, \
                       wsc, n, L.hdr, L.edges, L.tab, L.partial, L.minmax, bmax, cfg.cols,        \
                       cfg.ntab, accumulate_into, stats_parts, idx_base);                         \
        if (e != cudaSuccess) return e;                                                           \
    } while (0)
#define CRICK_LAUNCH_ACC(HW, VC, SG)                                                              \
    do {                                                                                          \
        if (stats_parts) CRICK_LAUNCH_ACC_S(HW, VC, SG, true);                                    \
        else CRICK_LAUNCH_ACC_S(HW, VC, SG, false);                                               \
    } while (0)
#define CRICK_LAUNCH_ACC2(HW, VC)                                                                  \
    do {                                                                                          \
        e = need_smem(k_bin_accumulate2<HW, VC>, smem);                                           \
        if (e != cudaSuccess) return e;                                                           \
        count_launch();                                                                           \
        e = launch_dep(k_bin_accumulate2<HW, VC>, dim3(grid), dim3(block), smem, st, X, W, wsc, n, \
                       L.hdr, L.edges, L.tab, L.partial, L.minmax, bmax, cfg.cols, cfg.ntab,      \
                       accumulate_into);                                                          \
        if (e != cudaSuccess) return e;                                                           \
    } while (0)
    const bool stage = vec && cfg.stage;   // unaligned input: plain loads, no staging
    if (v2) {
        if (W) { if (vec) CRICK_LAUNCH_ACC2(true, true); else CRICK_LAUNCH_ACC2(true, false); }
        else { if (vec) CRICK_LAUNCH_ACC2(false, true); else CRICK_LAUNCH_ACC2(false, false); }
    } else if (W) {
        if (!vec) CRICK_LAUNCH_ACC(true, false, false);
        else if (stage) CRICK_LAUNCH_ACC(true, true, true);
        else CRICK_LAUNCH_ACC(true, true, false);
    } else {
        if (!vec) CRICK_LAUNCH_ACC(false, false, false);
        else if (stage) CRICK_LAUNCH_ACC(false, true, true);
        else CRICK_LAUNCH_ACC(false, true, false);
    }
#undef CRICK_LAUNCH_ACC
#undef CRICK_LAUNCH_ACC_S
#undef CRICK_LAUNCH_ACC2
    if (pe0 && pe1) {
        cudaEventRecord(pe1, st);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof_ev.emplace_back(pe0, pe1);
    }
    return cudaGetLastError();
}

cudaError_t parallel_accumulate(const BatchView& v, const double* X, const double* W, double wsc,
                                long n, void* scratch, int sm_count, cudaStream_t st, int first,
                                StatsPart* stats_parts, long idx_base) {
    return accumulate_impl(v, X, W, wsc, n, scratch, sm_count, st, first ? 0 : 1, stats_parts,
                           idx_base);
}

cudaError_t parallel_finalize(const BatchView& v, long d, void* scratch, int sm_count,
                              cudaStream_t st, bool honor_badw) {
    ParLayout L = carve(scratch, v.comp, sm_count);
    const int mmax = mmax_for(v.comp);
    size_t wsb = fin_ws_bytes(v.cap, mmax);
    int use_smem = wsb <= 180 * 1024;
    size_t smem = use_smem ? wsb : 0;
    cudaError_t e = need_smem(k_finalize, use_smem ? wsb : 0);
    if (e != cudaSuccess) return e;
    count_launch();
    return launch_dep(k_finalize, dim3(1), dim3(1024), smem, st, v, d, L.hdr, L.edges, L.partial,
                      L.minmax, L.fin, mmax, use_smem, honor_badw ? 1 : 0);
}

// reads the "some weight is not > 0 and finite" flag of the accumulate calls since the last
// edges kernel (synchronises `st`)
cudaError_t parallel_bad_weights(void* scratch, cudaStream_t st, int* bad) {
    ParLayout L = carve(scratch, 100.0, 1);
    ParHeader h;
    cudaError_t e = cudaMemcpyAsync(&h, L.hdr, sizeof(h), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    *bad = (e == cudaSuccess) ? h.badw : 0;
    return e;
}

// sticky verdict of the deferred weight checks since the last call (synchronises `st`)
cudaError_t parallel_take_bad_weights(void* scratch, cudaStream_t st, int* bad) {
    ParLayout L = carve(scratch, 100.0, 1);
    ParHeader h;
    cudaError_t e = cudaMemcpyAsync(&h, L.hdr, sizeof(h), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    // bit 0: a deferred weight check failed; bit 1: k_exchange_merge gave up waiting for a rank
    *bad = (e == cudaSuccess) ? ((h.badw_sticky ? 1 : 0) | (h.p2p_timeout ? 2 : 0)) : 0;
    if (e == cudaSuccess && (h.badw_sticky || h.p2p_timeout))
        e = cudaMemsetAsync(&L.hdr->badw_sticky, 0, 2 * sizeof(int), st);
    return e;
}

cudaError_t launch_parallel_update(const BatchView& v, long d, const double* X, const double* W,
                                   double wsc, long n, void* scratch, int sm_count,
                                   cudaStream_t st, int* bad_weights, StatsPart* stats_parts,
                                   bool defer_check) {
    ParLayout L = carve(scratch, v.comp, sm_count);
    long S = sample_len(n);
    const bool tails = S == kSMax && n >= kTailMinN && tail_sampling_enabled();
    SampleSrc src{X, W, wsc, n, L.hdr, tails ? L.tmin : nullptr, tails ? L.tmax : nullptr};
    cudaError_t e = edges_from_sample(v, L, S, &src, st);
    if (e != cudaSuccess) return e;
    e = accumulate_impl(v, X, W, wsc, n, scratch, sm_count, st, 0, stats_parts, 0);
    if (e != cudaSuccess) return e;
    // validated update: k_finalize leaves the digest alone when the pass flagged a weight; the
    // flag is read back once everything is enqueued (one synchronisation, at the end)
    // (defer_check: the same, but the verdict stays on the device -- no synchronisation here)
    const bool checked = (bad_weights || defer_check) && W;
    e = parallel_finalize(v, d, scratch, sm_count, st, checked);
    if (e == cudaSuccess && checked && !defer_check) e = parallel_bad_weights(scratch, st, bad_weights);
    return e;
}

// ---- sharded update: sample -> edges -> accumulate on this rank, then k_exchange_merge ----------
size_t p2p_rec_doubles(double comp) { return 4 + 2 * (size_t)mmax_for(comp); }
size_t p2p_buffer_bytes(double comp, int world) {
    return ((size_t)kP2PFlagDoubles + 2 * (size_t)world * p2p_rec_doubles(comp)) * sizeof(double);
}
// The pieces of launch_parallel_update for pipelined callers: sample -> edges (reads X, W, writes only
// `scratch`), the streaming pass (the same), and the tail that folds the bucket sums into the digest
// (deferred weight check when `checked`).
cudaError_t launch_parallel_prep(const BatchView& v, const double* X, const double* W, double wsc, long n,
                                 void* scratch, int sm_count, cudaStream_t st) {
    ParLayout L = carve(scratch, v.comp, sm_count);
    long S = sample_len(n);
    const bool tails = S == kSMax && n >= kTailMinN && tail_sampling_enabled();
    SampleSrc src{X, W, wsc, n, L.hdr, tails ? L.tmin : nullptr, tails ? L.tmax : nullptr};
    return edges_from_sample(v, L, S, &src, st);
}
cudaError_t launch_parallel_accumulate(const BatchView& v, const double* X, const double* W, double wsc, long n,
                                       void* scratch, int sm_count, cudaStream_t st) {
    return accumulate_impl(v, X, W, wsc, n, scratch, sm_count, st, 0, nullptr, 0);
}
cudaError_t launch_parallel_pass(const BatchView& v, const double* X, const double* W, double wsc, long n,
                                 void* scratch, int sm_count, cudaStream_t st) {
    cudaError_t e = launch_parallel_prep(v, X, W, wsc, n, scratch, sm_count, st);
    if (e != cudaSuccess) return e;
    return launch_parallel_accumulate(v, X, W, wsc, n, scratch, sm_count, st);
}
cudaError_t launch_parallel_tail(const BatchView& v, long d, void* scratch, int sm_count, cudaStream_t st,
                                 bool checked) {
    return parallel_finalize(v, d, scratch, sm_count, st, checked);
}

cudaError_t launch_parallel_update_exchange(const BatchView& v, long d, const double* X, const double* W,
                                            double wsc, long n, void* scratch, int sm_count,
                                            cudaStream_t st, bool check_weights, int rank, int world,
                                            void* const* bufs, unsigned long long step) {
    if (world < 1 || world > kP2PMaxWorld) return cudaErrorInvalidValue;
    ParLayout L = carve(scratch, v.comp, sm_count);
    const int mmax = mmax_for(v.comp);
    // shared memory of the merge: cen[cap] | sb[pmax] | items[cap + pmax] | kend | start, next to
    // ~39 KB of static arrays; at least one rank's points must fit
    const size_t budget = 190000;
    long pmax = ((long)budget - (long)v.cap * 44 - 16) / 44;
    if (pmax > (long)world * mmax) pmax = (long)world * mmax;
    if (pmax < mmax) return cudaErrorInvalidValue;      // (huge compressions: use the NCCL reducer)
    const size_t smem = (size_t)v.cap * 16 + (size_t)pmax * 16 + ((size_t)v.cap + pmax) * 28 + 16;
    P2PView pv;
    pv.rank = rank; pv.world = world; pv.rec_doubles = (int)p2p_rec_doubles(v.comp);
    pv.empty = n <= 0 ? 1 : 0; pv.step = step;
    for (int r = 0; r < kP2PMaxWorld; ++r) pv.buf[r] = r < world ? reinterpret_cast<double*>(bufs[r]) : nullptr;
    cudaError_t e = cudaSuccess;
    if (n > 0) {
        long S = sample_len(n);
        const bool tails = S == kSMax && n >= kTailMinN && tail_sampling_enabled();
        SampleSrc src{X, W, wsc, n, L.hdr, tails ? L.tmin : nullptr, tails ? L.tmax : nullptr};
        e = edges_from_sample(v, L, S, &src, st);
        if (e != cudaSuccess) return e;
        e = accumulate_impl(v, X, W, wsc, n, scratch, sm_count, st, 0, nullptr, 0);
        if (e != cudaSuccess) return e;
    }
    e = need_smem(k_exchange_merge, smem);
    if (e != cudaSuccess) return e;
    count_launch();
    return launch_dep(k_exchange_merge, dim3(1), dim3(1024), smem, st, v, d, L.hdr, L.edges, L.partial,
                      L.minmax, mmax, (int)pmax, pv, (check_weights && W) ? 1 : 0);
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
    const int NS = 1024 / B > 0 ? 1024 / B : 1;   // k_finalize's order: NS strided slices, then slice order
    for (int b = 0; b < B; ++b) {
        double w = 0.0, s = 0.0;
        for (int sl = 0; sl < NS; ++sl) {
            double pw = 0.0, ps = 0.0;
            for (int c = sl; c < h.ncta; c += NS) {
                pw += rows[(size_t)c * kBPad + b].x; ps += rows[(size_t)c * kBPad + b].y;
            }
            w += pw; s += ps;
        }
        out_w[b] = w; out_s[b] = s;
    }
    double mn = INFINITY, mx = -INFINITY;
    for (int c = 0; c < h.ncta; ++c) { mn = fmin(mn, mm[c].x); mx = fmax(mx, mm[c].y); }
    out_minmax[0] = mn; out_minmax[1] = mx;
    return B;
}

// test hook: the 256-byte header block (ParHeader + phase timestamps of k_exchange_merge)
int parallel_debug_header(void* scratch, void* out256, cudaStream_t st) {
    if (cudaMemcpyAsync(out256, scratch, 256, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    return cudaStreamSynchronize(st) == cudaSuccess ? 0 : -1;
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
