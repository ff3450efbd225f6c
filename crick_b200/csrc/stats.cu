// crick_b200/csrc/stats.cu -- SummaryStats on the GPU (replaces crick/stats_stubs.c).
//
// State (stats_t, stats_stubs.c:12-22) lives on the device; the host keeps a mirror
// for the getters and the scalar finalisers (mean/var/std/skew/kurt, :98-136).
//
// Two update paths:
//  * sequential  (n < kStatsParallelMin, and every scalar add): one thread replays
//    stats_update_ndarray's loop (:199-210) with stats_do_update (:47-75) per element --
//    bit-identical to the reference whenever its int64 products do not overflow.
//  * parallel    (large arrays): per-thread shifted power sums -> central moments ->
//    Pebay pairwise merges up warp / CTA / grid, then one merge into the state.
//    Same formulas, different association: compared at tolerance (tests/test_stats.py).
//
// Integer products (row S1 of SURVEY.md 8a): the reference computes n1*n2, n1^2, n2^2 and
// n1n2*(n1^2 - n1n2 + n2^2) in int64 and silently wraps.  Here the int64 value is used when
// it is exact (so results match the reference bit for bit), and fp64 products otherwise
// (where the reference is wrong).  This is a deliberate, documented divergence.
#include "../../include/crick_cuda.h"
#include "api_common.h"
#include "stats_dev.cuh"

#include <cmath>
#include <type_traits>

namespace crick {

struct StatsState {
    long long count;
    double sum, vmin, vmax, m2, m3, m4;
    double first_value;
    int homogeneous;
    int pad;
};

constexpr long kStatsParallelMin = 1 << 8;    // the replay is one dependent chain: 1.5 us per element
static int g_stats_mode = -1;                  // -1 auto, 0 always the sequential replay, 1 always parallel

__host__ __device__ inline void stats_fresh(StatsState& s) {
    s.count = 0; s.sum = 0; s.vmin = DBL_MAX; s.vmax = -DBL_MIN;  // :29-37 (max quirk S0)
    s.m2 = 0; s.m3 = 0; s.m4 = 0; s.homogeneous = 1; s.first_value = 0.0; s.pad = 0;
}

// does a*b fit in int64 (both non-negative)?
__host__ __device__ inline bool mul_fits(long long a, long long b) {
    if (a == 0 || b == 0) return true;
    return a <= 9223372036854775807ll / b;
}

// stats_do_update, stats_stubs.c:47-75
__host__ __device__ inline void stats_combine(StatsState& T, long long n2, double sum2, double min2,
                                              double max2, double m4_2, double m3_2, double m2_2) {
#ifdef __CUDA_ARCH__
#define D_ADD(a, b) __dadd_rn(a, b)
#define D_MUL(a, b) __dmul_rn(a, b)
#define D_DIV(a, b) __ddiv_rn(a, b)
#else
#define D_ADD(a, b) ((a) + (b))
#define D_MUL(a, b) ((a) * (b))
#define D_DIV(a, b) ((a) / (b))
#endif
    const double u2 = D_DIV(sum2, (double)n2);
    const double delta = T.count ? D_ADD(u2, -D_DIV(T.sum, (double)T.count)) : u2;
    const long long n1 = T.count;
    const long long n = n1 + n2;
    const double dn = D_DIV(delta, (double)n);
    const double dn2 = D_MUL(dn, dn);
    const double dn3 = D_MUL(dn2, dn);
    if (min2 < T.vmin) T.vmin = min2;
    if (max2 > T.vmax) T.vmax = max2;
    // integer products: exact int64 when they fit, fp64 otherwise
    double n1n2, n1_2, n2_2, t4, t3;
    bool fits = mul_fits(n1, n2) && mul_fits(n1, n1) && mul_fits(n2, n2);
    if (fits) {
        long long i12 = n1 * n2, i11 = n1 * n1, i22 = n2 * n2;
        // n1^2 - n1n2 + n2^2 >= 0 and <= n1^2 + n2^2
        bool ok = i11 <= 9223372036854775807ll - i22;
        long long inner = ok ? (i11 - i12 + i22) : 0;
        ok = ok && mul_fits(i12, inner);
        n1n2 = (double)i12; n1_2 = (double)i11; n2_2 = (double)i22;
        t4 = ok ? (double)(i12 * inner) : D_MUL((double)i12, D_ADD(D_ADD((double)i11, -(double)i12), (double)i22));
        long long diff = n1 - n2;
        long long ad = diff < 0 ? -diff : diff;
        t3 = mul_fits(i12, ad) ? (double)(i12 * diff) : D_MUL((double)i12, (double)diff);
    } else {
        double f1 = (double)n1, f2 = (double)n2;
        n1n2 = D_MUL(f1, f2); n1_2 = D_MUL(f1, f1); n2_2 = D_MUL(f2, f2);
        t4 = D_MUL(n1n2, D_ADD(D_ADD(n1_2, -n1n2), n2_2));
        t3 = D_MUL(n1n2, D_ADD(f1, -f2));
    }
    const double f1 = (double)n1, f2 = (double)n2;
    // T->m4 += (m4_2 + t4*delta*dn3 + 6*(n1_2*m2_2 + n2_2*T->m2)*dn2 + 4*(n1*m3_2 - n2*T->m3)*dn)
    {
        double a = D_MUL(D_MUL(t4, delta), dn3);
        double b = D_MUL(D_MUL(6.0, D_ADD(D_MUL(n1_2, m2_2), D_MUL(n2_2, T.m2))), dn2);
        double c = D_MUL(D_MUL(4.0, D_ADD(D_MUL(f1, m3_2), -D_MUL(f2, T.m3))), dn);
        T.m4 = D_ADD(T.m4, D_ADD(D_ADD(D_ADD(m4_2, a), b), c));
    }
    // T->m3 += (m3_2 + t3*delta*dn2 + 3*(n1*m2_2 - n2*T->m2)*dn)
    {
        double a = D_MUL(D_MUL(t3, delta), dn2);
        double b = D_MUL(D_MUL(3.0, D_ADD(D_MUL(f1, m2_2), -D_MUL(f2, T.m2))), dn);
        T.m3 = D_ADD(T.m3, D_ADD(D_ADD(m3_2, a), b));
    }
    // T->m2 += m2_2 + n1n2*delta*dn
    T.m2 = D_ADD(T.m2, D_ADD(m2_2, D_MUL(D_MUL(n1n2, delta), dn)));
    T.count += n2;
    T.sum = D_ADD(T.sum, sum2);
#undef D_ADD
#undef D_MUL
#undef D_DIV
}

// stats_merge, stats_stubs.c:78-90
__host__ __device__ inline void stats_merge_state(StatsState& T1, const StatsState& T2) {
    if (T2.count == 0) return;
    if (T1.homogeneous && !T2.homogeneous) T1.homogeneous = 0;
    else if (T1.homogeneous && T2.homogeneous) T1.homogeneous = (T1.first_value == T2.first_value);
    stats_combine(T1, T2.count, T2.sum, T2.vmin, T2.vmax, T2.m4, T2.m3, T2.m2);
}

template <typename XT>
__device__ __forceinline__ double load_x(const XT* x, long i) { return (double)x[i]; }

// sequential replay of stats_update_ndarray's loop (:199-210)
template <typename XT>
__global__ void k_stats_seq(StatsState* st, const XT* __restrict__ x,
                            const long long* __restrict__ cnt, long long cnt_scalar, long n,
                            int track /* 0: replay of scalar stats_add calls, no :200-205 bookkeeping */) {
    if (threadIdx.x || blockIdx.x) return;
    StatsState T = *st;
    for (long i = 0; i < n; ++i) {
        double v = load_x(x, i);
        if (track) {
            if (T.count == 0) T.first_value = v;
            else if (T.homogeneous && T.first_value != v) T.homogeneous = 0;
        }
        if (v != v) continue;  // stats_add :93
        long long c = cnt ? cnt[i] : cnt_scalar;
        stats_combine(T, c, v, v, v, 0.0, 0.0, 0.0);
    }
    *st = T;
}

// ---------------------------------------------------------------------------
// parallel path
// 128-bit loads (2 elements), 4 in flight per thread; HBM bound (8 B / value)
template <typename XT, bool UNIT>
__global__ void __launch_bounds__(512) k_stats_partial(const XT* __restrict__ x,
                                                       const long long* __restrict__ cnt,
                                                       long long cnt_scalar, long n,
                                                       StatsPart* __restrict__ parts) {
    __shared__ StatsPart s_part[16];
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long stride = (long)gridDim.x * blockDim.x;
    StatsAcc a;
    a.K = 0.0; a.cN = 0.0; a.s1 = a.s2 = a.s3 = a.s4 = a.sx = 0.0;
    a.vmin = INFINITY; a.vmax = -INFINITY; a.cnt = 0; a.haveK = false;
    StatsPart p;
    part_fresh(p);
    const bool vec = (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                     (!cnt || (reinterpret_cast<uintptr_t>(cnt) & 15) == 0);
    long done = 0;
    if (vec) {
        const long npair = n >> 1;
        constexpr int U = 4;
        long j0 = tid;
        if (UNIT) {
            // full groups of UF 128-bit loads: no predicates, block update
            constexpr int UF = 8;
            for (; j0 + (UF - 1) * stride < npair; j0 += UF * stride) {
                double v[2 * UF];
#pragma unroll
                for (int k = 0; k < UF; ++k) {
                    const long j = j0 + k * stride;
                    if (sizeof(XT) == 8 && !std::is_same<XT, double>::value) {
                        longlong2 t = __ldg(reinterpret_cast<const longlong2*>(x) + j);
                        v[2 * k] = (double)t.x; v[2 * k + 1] = (double)t.y;
                    } else {
                        double2 t = __ldg(reinterpret_cast<const double2*>(x) + j);
                        v[2 * k] = t.x; v[2 * k + 1] = t.y;
                    }
                }
                stats_block_unit<2 * UF>(a, p, v, [&](int k) { return 2 * (j0 + (k >> 1) * stride) + (k & 1); });
            }
        }
        for (; j0 < npair; j0 += U * stride) {
            double v[2 * U];
            long long c[2 * U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const long j = j0 + k * stride;
                if (j < npair) {
                    if (sizeof(XT) == 8 && !std::is_same<XT, double>::value) {
                        longlong2 t = __ldg(reinterpret_cast<const longlong2*>(x) + j);
                        v[2 * k] = (double)t.x; v[2 * k + 1] = (double)t.y;
                    } else {
                        double2 t = __ldg(reinterpret_cast<const double2*>(x) + j);
                        v[2 * k] = t.x; v[2 * k + 1] = t.y;
                    }
                    if (!UNIT && cnt) {
                        longlong2 t = __ldg(reinterpret_cast<const longlong2*>(cnt) + j);
                        c[2 * k] = t.x; c[2 * k + 1] = t.y;
                    } else { c[2 * k] = c[2 * k + 1] = cnt_scalar; }
                }
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const long j = j0 + k * stride;
                if (j < npair) {
                    stats_one<UNIT>(a, p, v[2 * k], c[2 * k], 2 * j);
                    stats_one<UNIT>(a, p, v[2 * k + 1], c[2 * k + 1], 2 * j + 1);
                }
            }
        }
        done = npair << 1;
    }
    for (long i = done + tid; i < n; i += stride)
        stats_one<UNIT>(a, p, load_x(x, i), cnt ? cnt[i] : cnt_scalar, i);
    stats_thread_finish(a, p);
    p = stats_warp_reduce(p);
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) s_part[warp] = p;
    __syncthreads();
    if (threadIdx.x == 0) {
        StatsPart t = s_part[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) part_merge(t, s_part[w]);
        parts[blockIdx.x] = t;
    }
}

// one CTA of 1024 threads: shuffle tree per warp (fixed association, left = lower index), then
// thread 0 folds the <= 32 warp results -- ~40 dependent Pebay merges instead of nparts
__global__ void __launch_bounds__(1024) k_stats_final(StatsState* st,
                                                      const StatsPart* __restrict__ parts,
                                                      int nparts) {
    __shared__ StatsPart s_w[32];
    StatsPart p;
    part_fresh(p);
    // thread t folds parts t, t + 1024, ... (nparts <= 1024 in practice: one each)
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
        if (i == (int)threadIdx.x) p = parts[i]; else part_merge(p, parts[i]);
    }
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
        StatsPart o = part_shfl_xor(p, m);
        if ((threadIdx.x & m) == 0) part_merge(p, o);
        else { StatsPart t2 = o; part_merge(t2, p); p = t2; }
    }
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = p;
    __syncthreads();
    if (threadIdx.x) return;
    StatsPart t = s_w[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) part_merge(t, s_w[w]);
    StatsState T = *st;
    // homogeneous / first_value bookkeeping of the update loop (:200-205)
    const bool any_nonnan = t.first_idx != 0x7fffffffffffffffl;
    if (T.count == 0) {
        if (any_nonnan) {
            T.first_value = t.first_value;
            bool all_eq = (t.vmin == t.vmax) && (t.vmin == t.first_value);
            bool nan_after = t.last_nan_idx > t.first_idx;
            if (T.homogeneous && (!all_eq || nan_after)) T.homogeneous = 0;
        } else if (t.last_idx >= 0) {
            T.first_value = t.last_value;  // rewritten by every element while count == 0
        }
    } else if (t.last_idx >= 0) {
        bool all_eq = any_nonnan && t.last_nan_idx < 0 && t.vmin == t.vmax && t.vmin == T.first_value;
        if (T.homogeneous && !all_eq) T.homogeneous = 0;
    }
    if (t.count > 0) stats_combine(T, t.count, t.sum, t.vmin, t.vmax, t.m4, t.m3, t.m2);
    *st = T;
}

__global__ void k_stats_merge(StatsState* dst, const StatsState* src) {
    if (threadIdx.x || blockIdx.x) return;
    StatsState T = *dst;
    stats_merge_state(T, *src);
    *dst = T;
}

}  // namespace crick

using namespace crick;

namespace crick {
// 9-double wire record on the device: export, and merge of n records in order
__global__ void k_stats_reset(StatsState* st) {
    if (threadIdx.x) return;
    StatsState f;
    stats_fresh(f);
    *st = f;
}
__global__ void k_stats_export(const StatsState* st, double* r) {
    if (threadIdx.x) return;
    const StatsState m = *st;
    r[0] = (double)m.count; r[1] = m.sum; r[2] = m.vmin; r[3] = m.vmax; r[4] = m.m2; r[5] = m.m3;
    r[6] = m.m4; r[7] = (double)m.homogeneous; r[8] = m.first_value;
}
__global__ void k_stats_merge_records(StatsState* st, const double* __restrict__ recs, int n, long stride) {
    if (threadIdx.x) return;
    StatsState T = *st;
    for (int k = 0; k < n; ++k) {
        const double* r = recs + (size_t)k * stride;
        StatsState o;
        o.count = (long long)r[0]; o.sum = r[1]; o.vmin = r[2]; o.vmax = r[3]; o.m2 = r[4]; o.m3 = r[5];
        o.m4 = r[6]; o.homogeneous = r[7] != 0.0; o.first_value = r[8]; o.pad = 0;
        stats_merge_state(T, o);
    }
    *st = T;
}
}  // namespace crick

struct crick_stats {
    StatsState* d = nullptr;
    StatsPart* parts = nullptr;
    int nparts = 0;
    cudaStream_t own = nullptr, last = nullptr;
    cudaEvent_t ev = nullptr;
    std::vector<double> px;
    std::vector<long long> pc;
    StatsState mirror;
    bool mirror_ok = false;
};

static cudaStream_t ss_stream(crick_stats* s, void* stream) {
    cudaStream_t st = stream ? reinterpret_cast<cudaStream_t>(stream) : s->own;
    if (s->last && s->last != st) {
        cudaEventRecord(s->ev, s->last);
        cudaStreamWaitEvent(st, s->ev, 0);
    }
    s->last = st;
    return st;
}

static int ss_drain(crick_stats* s, cudaStream_t st) {
    size_t n = s->px.size();
    if (!n) return 0;
    unsigned char* d = nullptr;
    cudaError_t e = cudaMallocAsync(&d, n * 16, st);
    if (e != cudaSuccess) return fail("cudaMallocAsync(stats pending)", e);
    double* dx = reinterpret_cast<double*>(d);
    long long* dc = reinterpret_cast<long long*>(d + n * 8);
    e = cudaMemcpyAsync(dx, s->px.data(), n * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dc, s->pc.data(), n * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        count_launch();
        k_stats_seq<double><<<1, 32, 0, st>>>(s->d, dx, dc, 1, (long)n, 0);
        e = cudaGetLastError();
    }
    cudaFreeAsync(d, st);
    s->px.clear(); s->pc.clear();
    s->mirror_ok = false;
    return e == cudaSuccess ? 0 : fail("stats drain", e);
}

static int ss_sync(crick_stats* s) {
    cudaStream_t st = ss_stream(s, nullptr);
    if (ss_drain(s, st)) return -1;
    if (s->mirror_ok) return 0;
    cudaError_t e = cudaMemcpyAsync(&s->mirror, s->d, sizeof(StatsState), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail("stats read state", e);
    s->mirror_ok = true;
    return 0;
}

template <typename XT>
static cudaError_t stats_launch(crick_stats* s, const XT* x, const long long* c, long long cs, long n,
                                cudaStream_t st) {
    if (g_stats_mode == 0 || (g_stats_mode < 0 && n < kStatsParallelMin)) {
        count_launch();
        k_stats_seq<XT><<<1, 32, 0, st>>>(s->d, x, c, cs, n, 1);
        return cudaGetLastError();
    }
    long blocks = (n + 512 * 16 - 1) / (512 * 16);
    if (blocks > s->nparts) blocks = s->nparts;
    count_launch();
    if (!c && cs == 1) k_stats_partial<XT, true><<<(unsigned)blocks, 512, 0, st>>>(x, c, cs, n, s->parts);
    else k_stats_partial<XT, false><<<(unsigned)blocks, 512, 0, st>>>(x, c, cs, n, s->parts);
    count_launch();
    k_stats_final<<<1, 1024, 0, st>>>(s->d, s->parts, (int)blocks);
    return cudaGetLastError();
}

// ---- hooks for the fused t-digest + moments pass (tdigest_parallel.cu) -------------------
// begin: pending scalar adds are applied first; returns the per-CTA parts buffer.
namespace crick {
int stats_fused_begin(crick_stats* s, cudaStream_t st, StatsPart** parts, int* nparts) {
    cudaStream_t own = ss_stream(s, st);
    if (ss_drain(s, own)) return -1;
    s->mirror_ok = false;
    *parts = s->parts;
    *nparts = s->nparts;
    return 0;
}
// finish: fold `nused` parts (one per CTA of the accumulate kernel) into the state
cudaError_t stats_fused_finish(crick_stats* s, int nused, cudaStream_t st) {
    count_launch();
    k_stats_final<<<1, 1024, 0, st>>>(s->d, s->parts, nused);
    cudaError_t e = cudaGetLastError();
    // `st` belongs to ANOTHER handle (the digest's / the summary's own stream) and may be
    // destroyed before this one: hand the ordering back to the stats handle's own stream now
    // instead of remembering a stream that is not ours
    if (e == cudaSuccess && st != s->own) {
        e = cudaEventRecord(s->ev, st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s->own, s->ev, 0);
        s->last = s->own;
    }
    return e;
}
}  // namespace crick

// a call enqueued this handle's work on a stream that belongs to somebody else: order the
// handle's own stream behind it and forget the foreign stream
int crick_stats_release_stream(crick_stats_t* s, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (!st || st == s->own) return 0;
    cudaError_t e = cudaEventRecord(s->ev, st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s->own, s->ev, 0);
    s->last = s->own;
    return e == cudaSuccess ? 0 : fail("stats release stream", e);
}

extern "C" {

int crick_stats_kernel(int mode) {
    int prev = g_stats_mode;
    if (mode >= -1 && mode <= 1) g_stats_mode = mode;
    return prev;
}

crick_stats_t* crick_stats_new(void) {
    if (!sm_count()) { fail_msg("no CUDA device: libcrick_cuda has no CPU fallback"); return nullptr; }
    crick_stats* s = new (std::nothrow) crick_stats();
    if (!s) return nullptr;
    s->nparts = sm_count() * 4;   // 4 CTAs of 512 threads per SM
    cudaError_t e = cudaMalloc(&s->d, sizeof(StatsState));
    if (e == cudaSuccess) e = cudaMalloc(&s->parts, sizeof(StatsPart) * s->nparts);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->own, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev, cudaEventDisableTiming);
    stats_fresh(s->mirror);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s->d, &s->mirror, sizeof(StatsState), cudaMemcpyHostToDevice, s->own);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->own);
    if (e != cudaSuccess) { fail("stats_new", e); crick_stats_free(s); return nullptr; }
    s->mirror_ok = true;
    s->last = s->own;
    return s;
}

void crick_stats_free(crick_stats_t* s) {
    if (!s) return;
    // (only streams this handle owns are touched: `last` may be a caller's stream that no longer
    // exists; cudaFree below waits for the whole device before the memory goes away)
    if (s->own) { cudaStreamSynchronize(s->own); cudaStreamDestroy(s->own); }
    if (s->ev) cudaEventDestroy(s->ev);
    if (s->d) cudaFree(s->d);
    if (s->parts) cudaFree(s->parts);
    delete s;
}

int crick_stats_add(crick_stats_t* s, double x, int64_t count) {
    // stats_add (:92-95) is what SummaryStats.add calls: no homogeneous bookkeeping there
    // (that lives in the ndarray driver), so NaN is simply dropped and the rest is queued.
    if (x != x) return 0;
    s->px.push_back(x);
    s->pc.push_back(count);
    s->mirror_ok = false;
    if (s->px.size() >= (1u << 16)) return ss_drain(s, ss_stream(s, nullptr));
    return 0;
}

int crick_stats_update(crick_stats_t* s, const void* x, int x_is_int64, const int64_t* count,
                       int64_t count_scalar, int64_t n, int x_on_device, int count_on_device,
                       void* stream) {
    if (n <= 0) return 0;
    cudaStream_t st = ss_stream(s, stream);
    if (ss_drain(s, st)) return -1;
    s->mirror_ok = false;
    const void* dx = x;
    const long long* dc = reinterpret_cast<const long long*>(count);
    unsigned char* stage = nullptr;
    cudaError_t e = cudaSuccess;
    size_t need = (x_on_device ? 0 : (size_t)n * 8) + ((count && !count_on_device) ? (size_t)n * 8 : 0);
    if (need) {
        e = cudaMallocAsync(&stage, need, st);
        if (e != cudaSuccess) return fail("cudaMallocAsync(stats stage)", e);
        unsigned char* p = stage;
        if (!x_on_device) {
            e = staged_copy_h2d(p, x, n, st);
            dx = p; p += (size_t)n * 8;
        }
        if (e == cudaSuccess && count && !count_on_device) {
            e = staged_copy_h2d(p, count, n, st);
            dc = reinterpret_cast<const long long*>(p);
        }
    }
    if (e == cudaSuccess) {
        if (x_is_int64) e = stats_launch<long long>(s, reinterpret_cast<const long long*>(dx), dc, count_scalar, n, st);
        else e = stats_launch<double>(s, reinterpret_cast<const double*>(dx), dc, count_scalar, n, st);
    }
    if (stage) {
        cudaFreeAsync(stage, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    return e == cudaSuccess ? 0 : fail("stats_update", e);
}

int crick_stats_merge(crick_stats_t* s, crick_stats_t* const* others, int n_others) {
    cudaStream_t st = ss_stream(s, nullptr);
    if (ss_drain(s, st)) return -1;
    s->mirror_ok = false;
    for (int i = 0; i < n_others; ++i) {
        crick_stats* o = others[i];
        cudaStream_t ost = ss_stream(o, nullptr);
        if (ss_drain(o, ost)) return -1;
        cudaError_t e = cudaSuccess;
        if (o != s) {
            e = cudaEventRecord(o->ev, ost);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(st, o->ev, 0);
        }
        if (e == cudaSuccess) {
            count_launch();
            k_stats_merge<<<1, 32, 0, st>>>(s->d, o->d);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess && o != s) {
            e = cudaEventRecord(s->ev, st);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(ost, s->ev, 0);
        }
        if (e != cudaSuccess) return fail("stats_merge", e);
    }
    return 0;
}

int crick_stats_getstate(crick_stats_t* s, int64_t* count, double* f7, int* homogeneous) {
    if (ss_sync(s)) return -1;
    const StatsState& m = s->mirror;
    *count = m.count;
    f7[0] = m.sum; f7[1] = m.vmin; f7[2] = m.vmax; f7[3] = m.m2; f7[4] = m.m3; f7[5] = m.m4;
    f7[6] = m.first_value;
    *homogeneous = m.homogeneous;
    return 0;
}

int crick_stats_setstate(crick_stats_t* s, int64_t count, const double* f7, int homogeneous) {
    cudaStream_t st = ss_stream(s, nullptr);
    if (ss_drain(s, st)) return -1;
    StatsState m;
    m.count = count; m.sum = f7[0]; m.vmin = f7[1]; m.vmax = f7[2]; m.m2 = f7[3]; m.m3 = f7[4];
    m.m4 = f7[5]; m.first_value = f7[6]; m.homogeneous = homogeneous ? 1 : 0; m.pad = 0;
    s->mirror = m;
    cudaError_t e = cudaMemcpyAsync(s->d, &s->mirror, sizeof(StatsState), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail("stats_setstate", e);
    s->mirror_ok = true;
    return 0;
}

// finalisers: scalar host code on the mirrored state (stats_stubs.c:98-136)
double crick_stats_mean(crick_stats_t* s) {
    if (ss_sync(s)) return NAN;
    const StatsState& m = s->mirror;
    return m.count ? m.sum / m.count : NAN;
}
double crick_stats_var(crick_stats_t* s, int64_t ddof) {
    if (ss_sync(s)) return NAN;
    const StatsState& m = s->mirror;
    return m.count ? m.m2 / (m.count - ddof) : NAN;
}
double crick_stats_std(crick_stats_t* s, int64_t ddof) { return std::sqrt(crick_stats_var(s, ddof)); }
double crick_stats_skew(crick_stats_t* s, int bias) {
    if (ss_sync(s)) return NAN;
    const StatsState& m = s->mirror;
    if (m.homogeneous) return NAN;
    double n = (double)m.count, m2 = m.m2 / m.count, m3 = m.m3 / m.count;
    double skew = m2 ? m3 / (std::sqrt(m2) * m2) : 0;
    if (!bias && n > 2 && m2 > 0) return std::sqrt((n - 1) * n) / (n - 2) * skew;
    return skew;
}
double crick_stats_kurt(crick_stats_t* s, int fisher, int bias) {
    if (ss_sync(s)) return NAN;
    const StatsState& m = s->mirror;
    if (m.homogeneous) return NAN;
    double n = (double)m.count, m2 = m.m2 / m.count, m4 = m.m4 / m.count;
    double kurt = m2 ? m4 / (m2 * m2) : 0;
    if (!bias && n > 3 && m2 > 0) kurt = ((n * n - 1) * kurt - 9 * n + 15) / ((n - 2) * (n - 3));
    return fisher ? kurt - 3 : kurt;
}

int crick_stats_export_record(crick_stats_t* s, double* r) {
    if (ss_sync(s)) return -1;
    const StatsState& m = s->mirror;
    r[0] = (double)m.count; r[1] = m.sum; r[2] = m.vmin; r[3] = m.vmax; r[4] = m.m2; r[5] = m.m3;
    r[6] = m.m4; r[7] = (double)m.homogeneous; r[8] = m.first_value;
    return 0;
}

// back to the state stats_new leaves (:25-39), asynchronous on `stream`
int crick_stats_reset(crick_stats_t* s, void* stream) {
    cudaStream_t st = ss_stream(s, stream);
    s->px.clear(); s->pc.clear();
    count_launch();
    k_stats_reset<<<1, 32, 0, st>>>(s->d);
    cudaError_t e = cudaGetLastError();
    stats_fresh(s->mirror);
    s->mirror_ok = e == cudaSuccess;
    return e == cudaSuccess ? 0 : fail("stats_reset", e);
}

int crick_stats_export_record_dev(crick_stats_t* s, double* dev, void* stream) {
    cudaStream_t st = ss_stream(s, stream);
    if (ss_drain(s, st)) return -1;
    count_launch();
    k_stats_export<<<1, 32, 0, st>>>(s->d, dev);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail("stats_export_record_dev", e);
}

int crick_stats_merge_records_dev(crick_stats_t* s, const double* recs, int n, int64_t stride, void* stream) {
    if (n <= 0) return 0;
    cudaStream_t st = ss_stream(s, stream);
    if (ss_drain(s, st)) return -1;
    s->mirror_ok = false;
    count_launch();
    k_stats_merge_records<<<1, 32, 0, st>>>(s->d, recs, n, (long)stride);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail("stats_merge_records_dev", e);
}

int crick_stats_merge_record(crick_stats_t* s, const double* r) {
    cudaStream_t st = ss_stream(s, nullptr);
    if (ss_drain(s, st)) return -1;
    StatsState o;
    o.count = (long long)r[0]; o.sum = r[1]; o.vmin = r[2]; o.vmax = r[3]; o.m2 = r[4]; o.m3 = r[5];
    o.m4 = r[6]; o.homogeneous = r[7] != 0.0; o.first_value = r[8]; o.pad = 0;
    StatsState* d = nullptr;
    cudaError_t e = cudaMallocAsync(&d, sizeof(StatsState), st);
    if (e != cudaSuccess) return fail("cudaMallocAsync(stats record)", e);
    e = cudaMemcpyAsync(d, &o, sizeof(o), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        count_launch();
        k_stats_merge<<<1, 32, 0, st>>>(s->d, d);
        e = cudaGetLastError();
    }
    cudaFreeAsync(d, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    s->mirror_ok = false;
    return e == cudaSuccess ? 0 : fail("stats_merge_record", e);
}

}  // extern "C"
