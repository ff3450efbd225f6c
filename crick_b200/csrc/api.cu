// crick_b200/csrc/api.cu -- the C ABI of libcrick_cuda.so (include/crick_cuda.h):
// library plumbing, TDigest handles, grouped digests.  Host code only; every
// numeric result comes from the kernels in tdigest_kernels.cu / tdigest_parallel.cu.
#include "../../include/crick_cuda.h"
#include "api_common.h"
#include "tdigest_ref.cuh"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <chrono>
#include <map>
#include <cstdio>
#include <thread>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#endif

namespace crick {

static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int fail(const char* what, cudaError_t e) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    g_err = buf;
    return -1;
}
int fail_msg(const char* what) {
    g_err = what;
    return -1;
}

// SM count of the CURRENT device, cached per device; the first query of a device also sets its
// default memory pool's release threshold (the one place that does): freed stream-ordered staging
// buffers stay cached up to 2 GiB instead of going back to the driver at every synchronisation.
static std::atomic<int> g_sm_counts[64];
int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (dev >= 0 && dev < 64 && (n = g_sm_counts[dev].load(std::memory_order_relaxed)) > 0) return n;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = 2ull << 30;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaGetLastError();
    if (dev >= 0 && dev < 64) g_sm_counts[dev].store(n, std::memory_order_relaxed);
    return n;
}

// ---------------------------------------------------------------------------
// device storage for `nd` digests of one compression
int store_alloc(DigestStore& s, double compression, long nd) {
    if (compression < 20) compression = 20;          // tdigest_stubs.c:57-62
    else if (compression > 1000) compression = 1000;
    s.v.comp = compression;
    s.v.cap = 2 * (int)std::ceil(compression);       // :65
    s.v.bufcap = (int)(7.5 + 0.37 * compression - 2e-4 * compression * compression);  // :66
    s.nd = nd;
    size_t hb = ((size_t)nd * sizeof(DevHdr) + 255) & ~(size_t)255;
    size_t cb = (size_t)nd * s.v.cap * sizeof(double2);
    size_t bb = (((size_t)nd * s.v.bufcap * sizeof(double2)) + 255) & ~(size_t)255;
    s.bytes = hb + 2 * cb + bb;
    // (crick_device_alloc caches freed blocks by size: creating and freeing digests in a loop does
    // not go through cudaMalloc / cudaFree each time)
    s.base = crick_device_alloc((int64_t)s.bytes);
    if (!s.base) return -1;
    unsigned char* p = reinterpret_cast<unsigned char*>(s.base);
    s.v.hdr = reinterpret_cast<DevHdr*>(p); p += hb;
    s.v.cen = reinterpret_cast<double2*>(p); p += cb;
    s.v.table = reinterpret_cast<double2*>(p); p += cb;
    s.v.buf = reinterpret_cast<double2*>(p);
    return 0;
}
void store_free(DigestStore& s) {
    if (s.base) crick_device_free(s.base);
    s.base = nullptr;
}

}  // namespace crick

using namespace crick;

// ---------------------------------------------------------------------------
struct crick_tdigest {
    DigestStore s;
    cudaStream_t own = nullptr;      // the handle's stream
    cudaStream_t last = nullptr;     // stream of the most recent enqueued op
    cudaEvent_t ev = nullptr;
    std::vector<double> px, pw;      // pending scalar adds (tdigest_add), in order
    DevHdr mirror;                   // host copy of the header
    bool mirror_ok = false;
    bool buf_clean = false;          // the insertion buffer is known to be empty (no flush needed)
    bool lazy_reset = false;         // crick_tdigest_reset was asked for but not launched yet: the
                                     // one-pass record merge takes the digest as empty by itself
    void* scratch = nullptr;         // parallel-mode scratch (lazy)
    size_t scratch_bytes = 0;
    // host-input path: copy stream and per-slot events, created once per handle (a stream and six
    // events cost ~130 us per call to create and destroy)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[3] = {nullptr, nullptr, nullptr}, ev_done[3] = {nullptr, nullptr, nullptr};
    // pipelined updates (crick_tdigest_update_pipelined): two complete scratch sets, used in turn, so
    // that the tail of update k (k_finalize, on the tail stream) overlaps the sample / edges phase
    // and the pass of update k + 1 (on the main stream)
    void* pipe_scratch[2] = {nullptr, nullptr};
    cudaEvent_t ev_pass = nullptr, ev_fin[2] = {nullptr, nullptr};
    cudaStream_t pass_stream = nullptr;      // overlapped sample phase: the passes run here, one after the other
    cudaEvent_t ev_prep = nullptr;
    bool fin_recorded[2] = {false, false};
    int pipe_next = 0;
};

static cudaStream_t td_stream(crick_tdigest* t, void* stream, bool keep_lazy_reset = false) {
    // Chain work across streams: if this op runs on a different stream than the previous
    // one, make it wait for everything enqueued so far.
    cudaStream_t st = stream ? reinterpret_cast<cudaStream_t>(stream) : t->own;
    if (t->last && t->last != st) {
        cudaEventRecord(t->ev, t->last);
        cudaStreamWaitEvent(st, t->ev, 0);
    }
    t->last = st;
    if (t->lazy_reset && !keep_lazy_reset) {      // whoever comes next needs the reset header in memory
        t->lazy_reset = false;
        launch_init_hdr(t->s.v.hdr, 1, st);
    }
    return st;
}

static int td_drain(crick_tdigest* t, cudaStream_t st) {
    size_t n = t->px.size();
    if (!n) return 0;
    double* d = nullptr;
    cudaError_t e = cudaMallocAsync(&d, 2 * n * sizeof(double), st);
    if (e != cudaSuccess) return fail("cudaMallocAsync(pending)", e);
    e = cudaMemcpyAsync(d, t->px.data(), n * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d + n, t->pw.data(), n * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess)
        e = launch_ingest(t->s.v, 1, d, d + n, 0.0, 1, 1, nullptr, (long)n, 0, nullptr, 0, st);
    // pageable source: the copies above have consumed the host vectors when they return
    cudaFreeAsync(d, st);
    t->px.clear();
    t->pw.clear();
    t->mirror_ok = false; t->buf_clean = false;
    if (e != cudaSuccess) return fail("drain pending adds", e);
    return 0;
}

static int td_sync_mirror(crick_tdigest* t) {
    cudaStream_t st = td_stream(t, nullptr);
    if (td_drain(t, st)) return -1;
    if (t->mirror_ok) return 0;
    cudaError_t e = cudaMemcpyAsync(&t->mirror, t->s.v.hdr, sizeof(DevHdr), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail("read digest header", e);
    t->mirror_ok = true;
    return 0;
}

static int td_scratch(crick_tdigest* t) {
    if (t->scratch) return 0;
    size_t b = parallel_scratch_bytes(t->s.v.comp, sm_count());
    t->scratch = crick_device_alloc((int64_t)b);
    if (!t->scratch) return -1;
    t->scratch_bytes = b;
    // the header (sticky flags) starts zeroed; ordered before every use: all of them go through
    // td_stream(), which chains the handle's streams
    cudaError_t e = cudaMemsetAsync(t->scratch, 0, 256, t->last ? t->last : t->own);
    if (e != cudaSuccess) return fail("cudaMemsetAsync(scratch)", e);
    return 0;
}

extern "C" {

// ---- library ---------------------------------------------------------------
const char* crick_last_error(void) { return g_err.c_str(); }
const char* crick_version(void) { return "crick_b200 0.1.0 (sm_100a)"; }
int crick_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int crick_set_device(int device) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail("cudaSetDevice", e);
    return 0;
}
int crick_sm_count(void) { return sm_count(); }
void* crick_stream_create(void) {
    cudaStream_t st = nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (e != cudaSuccess) { fail("cudaStreamCreate", e); return nullptr; }
    return st;
}
void crick_stream_destroy(void* stream) {
    if (stream) cudaStreamDestroy(static_cast<cudaStream_t>(stream));
}
int crick_get_device(void) {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
    return dev;
}
int64_t crick_launch_count(void) { return (int64_t)g_launches.load(); }
void* crick_host_alloc(int64_t bytes) {
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) { fail("cudaHostAlloc", e); return nullptr; }
    return p;
}
void crick_host_free(void* p) { if (p) cudaFreeHost(p); }
// Device buffers handed out by the library (DeviceArray, digest stores, parallel scratch): plain
// cudaMalloc, plus a small cache of freed blocks by size.  A cudaMalloc / cudaFree pair costs
// 0.1-25 ms once the process has tens of GB mapped (bench.py's e2e leg creates and frees a digest
// per step; every large device query allocates its output), and returning blocks to the driver's
// stream-ordered pool instead makes it trim at some later synchronisation.  Blocks up to 128 MB
// are kept, 1 GB in total; a block is only cached after a device synchronisation, so whoever
// gets it next cannot race with its previous user.
// Small blocks (a digest's store is 16 KB) are carved from 8 MB slabs -- a cold cudaMalloc per
// digest costs ~0.7 ms -- and go back to the cache, never to the driver.
static std::mutex g_mem_mu;
static std::map<void*, size_t> g_mem_size;                 // live blocks (bit 63: carved from a slab)
static std::map<void*, int> g_mem_dev;                     // ... and the device they live on
static unsigned char* g_slab_of[64] = {};
static size_t g_slab_left_of[64] = {};
constexpr size_t kSlabBytes = (size_t)8 << 20, kSlabMaxBlock = (size_t)256 << 10;
constexpr size_t kCarved = (size_t)1 << 63;
static std::multimap<std::pair<int, size_t>, void*> g_mem_cache;   // freed blocks by (device, size)
static size_t g_mem_cached = 0;
void* crick_device_alloc(int64_t bytes) {
    size_t nb = (size_t)(bytes > 0 ? bytes : 1);
    const bool small = nb <= kSlabMaxBlock;
    if (small) nb = (nb + 255) & ~(size_t)255;
    void* p = nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { fail_msg("no CUDA device"); return nullptr; }
    {
        std::lock_guard<std::mutex> lk(g_mem_mu);
        unsigned char*& g_slab = g_slab_of[dev];
        size_t& g_slab_left = g_slab_left_of[dev];
        auto it = g_mem_cache.find(std::make_pair(dev, nb));
        if (it != g_mem_cache.end()) {
            p = it->second;
            g_mem_cache.erase(it);
            if (!small) g_mem_cached -= nb;
            g_mem_size[p] = nb | (small ? kCarved : 0);
            g_mem_dev[p] = dev;
            return p;
        }
        if (small) {
            if (g_slab_left < nb) {
                void* slab = nullptr;
                if (cudaMalloc(&slab, kSlabBytes) != cudaSuccess) { cudaGetLastError(); slab = nullptr; }
                g_slab = reinterpret_cast<unsigned char*>(slab);
                g_slab_left = slab ? kSlabBytes : 0;
            }
            if (g_slab_left >= nb) {
                p = g_slab;
                g_slab += nb;
                g_slab_left -= nb;
                g_mem_size[p] = nb | kCarved;
                g_mem_dev[p] = dev;
                return p;
            }
        }
    }
    cudaError_t e = cudaMalloc(&p, nb);
    if (e != cudaSuccess) {
        // give the cached blocks back and try once more
        std::vector<void*> drop;
        {
            std::lock_guard<std::mutex> lk(g_mem_mu);
            for (auto it = g_mem_cache.begin(); it != g_mem_cache.end();) {
                if (it->first.second > kSlabMaxBlock) { drop.push_back(it->second); it = g_mem_cache.erase(it); }
                else ++it;                        // slab pieces cannot be returned one by one
            }
            g_mem_cached = 0;
        }
        cudaGetLastError();
        for (void* q : drop) cudaFree(q);
        e = cudaMalloc(&p, nb);
    }
    if (e != cudaSuccess) { fail("cudaMalloc", e); return nullptr; }
    std::lock_guard<std::mutex> lk(g_mem_mu);
    g_mem_size[p] = nb;
    g_mem_dev[p] = dev;
    return p;
}
void crick_device_free(void* p) {
    if (!p) return;
    size_t nb = 0;
    int dev = 0;
    {
        std::lock_guard<std::mutex> lk(g_mem_mu);
        auto it = g_mem_size.find(p);
        if (it != g_mem_size.end()) { nb = it->second; g_mem_size.erase(it); }
        auto jt = g_mem_dev.find(p);
        if (jt != g_mem_dev.end()) { dev = jt->second; g_mem_dev.erase(jt); }
    }
    // a block goes back to the cache only when nothing enqueued on ITS device can still touch it
    auto sync_owner = [&]() {
        int cur = 0;
        if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); cudaDeviceSynchronize(); return; }
        if (cur != dev) cudaSetDevice(dev);
        cudaDeviceSynchronize();
        if (cur != dev) cudaSetDevice(cur);
    };
    if (nb & kCarved) {                       // part of a slab: always back to the cache
        nb &= ~kCarved;
        sync_owner();
        std::lock_guard<std::mutex> lk(g_mem_mu);
        g_mem_cache.emplace(std::make_pair(dev, nb), p);
        return;
    }
    if (nb && nb <= ((size_t)128 << 20)) {
        sync_owner();
        std::lock_guard<std::mutex> lk(g_mem_mu);
        if (g_mem_cached + nb <= ((size_t)1 << 30)) {
            g_mem_cache.emplace(std::make_pair(dev, nb), p);
            g_mem_cached += nb;
            return;
        }
    }
    cudaFree(p);
}
int crick_memcpy_h2d(void* dst, const void* src, int64_t bytes, void* stream) {
    cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice,
                                    reinterpret_cast<cudaStream_t>(stream));
    if (e == cudaSuccess && !stream) e = cudaStreamSynchronize(nullptr);
    return e == cudaSuccess ? 0 : fail("memcpy h2d", e);
}
int crick_memcpy_d2h(void* dst, const void* src, int64_t bytes, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    return e == cudaSuccess ? 0 : fail("memcpy d2h", e);
}
int crick_stream_synchronize(void* stream) {
    cudaError_t e = cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail("cudaStreamSynchronize", e);
}
int crick_device_synchronize(void) {
    cudaError_t e = cudaDeviceSynchronize();
    return e == cudaSuccess ? 0 : fail("cudaDeviceSynchronize", e);
}
int crick_profile_enable(int on) { profile_enable(on); return 0; }
int crick_group_kernel(int mode) { return ingest_lanes_mode(mode); }
double crick_group_margin(double margin) { return ingest_lanes_margin(margin); }
int crick_group_exact_fallbacks(int64_t* out, int reset) {
    if (!out) return fail("crick_group_exact_fallbacks: null out", cudaErrorInvalidValue);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long n = 0;
    if (e == cudaSuccess) e = ingest_lanes_fallbacks(&n, reset != 0);
    *out = (int64_t)n;
    return e == cudaSuccess ? 0 : fail("crick_group_exact_fallbacks", e);
}
int crick_profile_collect(double* total_ms, int64_t* launches) {
    long long n = 0;
    int rc = profile_collect(total_ms, &n);
    *launches = n;
    return rc;
}
int crick_cast_to_f64(const void* src, int kind, int itemsize, int64_t n, double* dst, void* stream) {
    cudaError_t e = launch_cast_f64(src, kind, itemsize, n, dst, reinterpret_cast<cudaStream_t>(stream));
    if (e == cudaErrorInvalidValue) return fail_msg("crick_cast_to_f64: unsupported dtype");
    return e == cudaSuccess ? 0 : fail("cast_to_f64", e);
}
int crick_check_weights(const double* w_dev, int64_t n, void* stream) {
    if (n <= 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int* flag = nullptr;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&flag), sizeof(int), st);
    if (e != cudaSuccess) return fail("cudaMallocAsync(check_weights)", e);
    int bad = 0;
    e = cudaMemsetAsync(flag, 0, sizeof(int), st);
    if (e == cudaSuccess) e = launch_check_weights(w_dev, n, flag, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaFreeAsync(flag, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail("check_weights", e);
    return bad ? 1 : 0;
}
int crick_synth_fill(double* x, double* w, int64_t n, uint64_t seed, int64_t offset, int kind,
                     void* stream) {
    cudaError_t e = launch_synth(x, w, n, seed, offset, kind, reinterpret_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail("synth", e);
}

// ---- TDigest ---------------------------------------------------------------
crick_tdigest_t* crick_tdigest_new(double compression) {
    if (!sm_count()) { fail_msg("no CUDA device: libcrick_cuda has no CPU fallback"); return nullptr; }
    crick_tdigest* t = new (std::nothrow) crick_tdigest();
    if (!t) return nullptr;
    if (store_alloc(t->s, compression, 1)) { delete t; return nullptr; }
    cudaError_t e = cudaStreamCreateWithFlags(&t->own, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = launch_init_hdr(t->s.v.hdr, 1, t->own);
    if (e != cudaSuccess) { fail("tdigest_new", e); crick_tdigest_free(t); return nullptr; }
    t->last = t->own;
    return t;
}

void crick_tdigest_free(crick_tdigest_t* t) {
    if (!t) return;
    // (only streams this handle owns are touched: `last` may be a caller's stream that no longer
    // exists; crick_device_free waits for the device before a block is reused or released)
    if (t->own) { cudaStreamSynchronize(t->own); cudaStreamDestroy(t->own); }
    if (t->ev) cudaEventDestroy(t->ev);
    if (t->copy_stream) { cudaStreamSynchronize(t->copy_stream); cudaStreamDestroy(t->copy_stream); }
    for (int b = 0; b < 3; ++b) {
        if (t->ev_copied[b]) cudaEventDestroy(t->ev_copied[b]);
        if (t->ev_done[b]) cudaEventDestroy(t->ev_done[b]);
    }
    if (t->scratch) crick_device_free(t->scratch);
    for (int b = 0; b < 2; ++b) {
        if (t->pipe_scratch[b]) crick_device_free(t->pipe_scratch[b]);
        if (t->ev_fin[b]) cudaEventDestroy(t->ev_fin[b]);
    }
    if (t->ev_pass) cudaEventDestroy(t->ev_pass);
    if (t->ev_prep) cudaEventDestroy(t->ev_prep);
    if (t->pass_stream) { cudaStreamSynchronize(t->pass_stream); cudaStreamDestroy(t->pass_stream); }
    store_free(t->s);
    delete t;
}

int crick_tdigest_add(crick_tdigest_t* t, double x, double w) {
    t->px.push_back(x);
    t->pw.push_back(w);
    if (t->px.size() >= (1u << 16)) return td_drain(t, td_stream(t, nullptr));
    return 0;
}

// host-resident input, parallel mode: same sample positions as the device sampler, then
// double-buffered chunks (copy of chunk k+1 overlaps the accumulate of chunk k)
// ---- pageable host input ------------------------------------------------------------------
// cudaMemcpyAsync from ordinary (pageable) memory is staged by the driver on one thread at
// ~11 GB/s (tests/tools/pageable_probe.py) -- a fifth of what the link moves from pinned memory.
// NumPy arrays are pageable, so the library stages them itself: a few host threads copy each
// chunk into one of three cached pinned slots while the previous slots are in flight to the GPU.
constexpr int64_t kStageChunk = 1 << 22;   // values per slot and array (32 MiB)
constexpr int kStageSlots = 3;
struct StagePool {
    std::mutex mu;
    double* slot[kStageSlots] = {nullptr, nullptr, nullptr};   // each: x chunk | w chunk
    bool busy = false;
};
static StagePool g_stage;

static bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

// Copy into a pinned staging slot with non-temporal stores: the DMA engine reads the slot right
// afterwards, and lines left dirty in the caches of several cores slow that read down several
// times (8 MB: H2D 160 us after streaming stores or one big memcpy, 260-1200 us after eight
// threads of ordinary stores).  dst must be 16-byte aligned.
static void stream_copy(double* dst, const double* src, int64_t n) {
#if defined(__x86_64__) || defined(_M_X64)
    int64_t i = 0;
    for (; i + 8 <= n; i += 8) {
        const __m128d a = _mm_loadu_pd(src + i), b = _mm_loadu_pd(src + i + 2);
        const __m128d c = _mm_loadu_pd(src + i + 4), d = _mm_loadu_pd(src + i + 6);
        _mm_stream_pd(dst + i, a); _mm_stream_pd(dst + i + 2, b);
        _mm_stream_pd(dst + i + 4, c); _mm_stream_pd(dst + i + 6, d);
    }
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
#else
    memcpy(dst, src, (size_t)n * sizeof(double));
#endif
}

static void parallel_memcpy(double* dst, const double* src, int64_t n, int nthreads) {
    if (nthreads <= 1 || n < (1 << 16)) { stream_copy(dst, src, n); return; }
    std::vector<std::thread> th;
    const int64_t per = (((n + nthreads - 1) / nthreads) + 1) & ~(int64_t)1;   // even: 16-byte aligned parts
    for (int k = 0; k < nthreads; ++k) {
        const int64_t lo = k * per, hi = lo + per < n ? lo + per : n;
        if (lo >= hi) break;
        th.emplace_back([=] { stream_copy(dst + lo, src + lo, hi - lo); });
    }
    for (auto& t : th) t.join();
}

// takes the process-wide staging slots (allocated on first use); false = in use by another
// thread or out of pinned memory: the caller falls back to the driver's pageable copy
static bool stage_acquire() {
    std::lock_guard<std::mutex> lk(g_stage.mu);
    if (g_stage.busy) return false;
    for (int i = 0; i < kStageSlots; ++i) {
        if (!g_stage.slot[i] &&
            cudaHostAlloc(reinterpret_cast<void**>(&g_stage.slot[i]),
                          (size_t)kStageChunk * 2 * sizeof(double), cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            g_stage.slot[i] = nullptr;
            return false;
        }
    }
    g_stage.busy = true;
    return true;
}
static void stage_release() {
    std::lock_guard<std::mutex> lk(g_stage.mu);
    g_stage.busy = false;
}

// Host -> device copy of `n8` 8-byte elements for the other sketches (stats.cu, spsv.cu): pageable
// sources go through the pinned slots in 4 Mi-element pieces (streaming-store copies by several
// threads, overlapped with the DMA of the previous piece); pinned sources, or slots in use by
// another thread, take the plain cudaMemcpyAsync.  Returns when `src` has been consumed.
extern "C++" {
namespace crick {
cudaError_t staged_copy_h2d(void* dst, const void* src, int64_t n8, cudaStream_t st) {
    if (n8 < (1 << 17) || !is_pageable(src) || !stage_acquire())
        return cudaMemcpyAsync(dst, src, (size_t)n8 * 8, cudaMemcpyHostToDevice, st);
    unsigned hw = std::thread::hardware_concurrency();
    const int nthreads = hw >= 16 ? 8 : (hw >= 4 ? (int)hw / 2 : 1);
    cudaEvent_t ev[kStageSlots] = {nullptr, nullptr, nullptr};
    cudaError_t e = cudaSuccess;
    for (int b = 0; b < kStageSlots && e == cudaSuccess; ++b)
        e = cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming);
    const int64_t piece = 2 * kStageChunk;                    // a whole slot (x part | w part)
    int k = 0;
    for (int64_t off = 0; off < n8 && e == cudaSuccess; off += piece, ++k) {
        const int b = k % kStageSlots;
        const int64_t m = (n8 - off) < piece ? (n8 - off) : piece;
        if (k >= kStageSlots) e = cudaEventSynchronize(ev[b]);      // the slot's previous DMA is done
        if (e != cudaSuccess) break;
        parallel_memcpy(g_stage.slot[b], reinterpret_cast<const double*>(src) + off, m,
                        m >= ((int64_t)1 << 19) ? nthreads : 1);
        e = cudaMemcpyAsync(reinterpret_cast<double*>(dst) + off, g_stage.slot[b], (size_t)m * 8,
                            cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaEventRecord(ev[b], st);
    }
    // the slots go back to the pool: their last DMAs must have left them
    for (int b = 0; b < kStageSlots; ++b)
        if (ev[b]) { if (e == cudaSuccess) e = cudaEventSynchronize(ev[b]); cudaEventDestroy(ev[b]); }
    stage_release();
    return e;
}
}  // namespace crick
}  // extern "C++"

static int td_update_host_parallel(crick_tdigest* t, const double* x, const double* w, double wsc,
                                   int64_t n, cudaStream_t st, int* bad_weights,
                                   crick_stats* stats, bool clean) {
    if (td_scratch(t)) return -1;
    // CRICK_TRACE_HOST=1: one line per call with the host-side time of every phase
    static const bool trace = getenv("CRICK_TRACE_HOST") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::micro>(b - a).count();
    };
    const auto t0 = now();
    double t_stage = 0.0;
    cudaError_t e = clean ? cudaSuccess : launch_flush(t->s.v, 0, 1, 1, st);
    if (e != cudaSuccess) return fail("flush", e);
    if (n <= kStageChunk) {
        // one chunk: copy it (pageable sources through the staging slots) and run the
        // device-resident update on the copy -- the sample is then drawn on the GPU at the very
        // positions the host loop below would read, so the result is the same
        const int64_t cw = (n + 1) & ~(int64_t)1;
        double* d = nullptr;
        e = cudaMallocAsync(reinterpret_cast<void**>(&d), (size_t)cw * 8 * (w ? 2 : 1), st);
        if (e != cudaSuccess) return fail("cudaMallocAsync(stage)", e);
        e = staged_copy_h2d(d, x, n, st);
        if (e == cudaSuccess && w) e = staged_copy_h2d(d + cw, w, n, st);
        StatsPart* sp1 = nullptr;
        int sn1 = 0;
        if (e == cudaSuccess && stats && stats_fused_begin(stats, st, &sp1, &sn1)) { cudaFreeAsync(d, st); return -1; }
        if (e == cudaSuccess)
            e = launch_parallel_update(t->s.v, 0, d, w ? d + cw : nullptr, wsc, n, t->scratch, sm_count(), st,
                                       bad_weights, sp1);
        if (e == cudaSuccess && stats && !(bad_weights && *bad_weights)) e = stats_fused_finish(stats, sm_count(), st);
        if (stats) crick_stats_release_stream(stats, st);
        cudaFreeAsync(d, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);    // the caller may reuse x / w at once
        if (trace) fprintf(stderr, "crick host update n=%lld: one chunk, %.0f us\n", (long long)n, us(t0, now()));
        if (e != cudaSuccess) return fail("host parallel update", e);
        t->buf_clean = true;
        return 0;
    }
    const long S = parallel_sample_len(n);
    std::vector<double> sx((size_t)S), sw((size_t)S);
    for (long i = 0; i < S; ++i) {
        long pos = parallel_sample_pos(i, n, S);
        sx[i] = x[pos];
        sw[i] = w ? w[pos] : wsc;
    }
    e = cudaMemcpyAsync(parallel_sample_x(t->scratch), sx.data(), S * sizeof(double),
                        cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(parallel_sample_w(t->scratch), sw.data(), S * sizeof(double),
                            cudaMemcpyHostToDevice, st);
    const auto t1 = now();
    if (e == cudaSuccess) e = parallel_begin_from_sample(t->s.v, S, t->scratch, st);
    if (e != cudaSuccess) return fail("parallel edges", e);
    const auto t2 = now();

    StatsPart* sparts = nullptr;
    int snparts = 0;
    if (stats && stats_fused_begin(stats, st, &sparts, &snparts)) return -1;
    // pinned (or registered) input goes to the GPU straight from the caller's memory in 16 Mi
    // chunks; pageable input through the staging slots in 4 Mi chunks
    const bool staged = (is_pageable(x) || (w && is_pageable(w))) && stage_acquire();
    const int64_t chunk = staged ? kStageChunk : (int64_t)1 << 24;
    const int nbuf = staged ? kStageSlots : 2;
    unsigned hw = std::thread::hardware_concurrency();
    int nthreads = hw >= 16 ? 8 : (hw >= 4 ? (int)hw / 2 : 1);
    if (const char* env = getenv("CRICK_STAGE_THREADS")) { int v = atoi(env); if (v >= 1 && v <= 64) nthreads = v; }
    double* dbuf[kStageSlots] = {nullptr, nullptr, nullptr};
    cudaStream_t cst = nullptr;
    cudaEvent_t copied[kStageSlots] = {nullptr, nullptr, nullptr}, done[kStageSlots] = {nullptr, nullptr, nullptr};
    const int64_t cw = ((n < chunk ? n : chunk) + 1) & ~(int64_t)1;  // even: w half stays 16-byte aligned
    const size_t per = (size_t)cw * sizeof(double) * (w ? 2 : 1);
    int rc = 0;
    static_assert(kStageSlots == 3, "the handle keeps three event pairs");
    if (!t->copy_stream) e = cudaStreamCreateWithFlags(&t->copy_stream, cudaStreamNonBlocking);
    cst = t->copy_stream;
    for (int b = 0; b < nbuf && e == cudaSuccess; ++b) {
        e = cudaMallocAsync(reinterpret_cast<void**>(&dbuf[b]), per, st);   // stream-ordered pool: no device sync
        if (e == cudaSuccess && !t->ev_copied[b]) e = cudaEventCreateWithFlags(&t->ev_copied[b], cudaEventDisableTiming);
        if (e == cudaSuccess && !t->ev_done[b]) e = cudaEventCreateWithFlags(&t->ev_done[b], cudaEventDisableTiming);
        copied[b] = t->ev_copied[b];
        done[b] = t->ev_done[b];
    }
    if (e == cudaSuccess) {
        // the staging buffers come from st's memory pool: the copy stream may touch them only
        // after the allocation point
        cudaEvent_t alloc_ev = copied[0];
        e = cudaEventRecord(alloc_ev, st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(cst, alloc_ev, 0);
        int k = 0;
        const auto t3 = now();
        if (trace) fprintf(stderr, "crick host update n=%lld: sample %.0f us, edges enqueue %.0f us, setup %.0f us",
                           (long long)n, us(t0, t1), us(t1, t2), us(t2, t3));
        for (int64_t off = 0; off < n && e == cudaSuccess; off += chunk, ++k) {
            const int b = k % nbuf;
            const int64_t m = (n - off) < chunk ? (n - off) : chunk;
            if (k >= nbuf) e = cudaStreamWaitEvent(cst, done[b], 0);
            const double* hx = x + off;
            const double* hw_ = w ? w + off : nullptr;
            if (staged) {
                // the slot's previous H2D copy must have left the pinned buffer
                if (k >= nbuf && e == cudaSuccess) e = cudaEventSynchronize(copied[b]);
                const auto ts = now();
                // several writers only pay from ~4 MB (thread start-up); the copies use streaming
                // stores (stream_copy) so that the DMA which follows finds the data in DRAM
                const int nt = m >= ((int64_t)1 << 19) ? nthreads : 1;
                parallel_memcpy(g_stage.slot[b], x + off, m, nt);
                if (w) parallel_memcpy(g_stage.slot[b] + kStageChunk, w + off, m, nt);
                t_stage += us(ts, now());
                hx = g_stage.slot[b];
                hw_ = w ? g_stage.slot[b] + kStageChunk : nullptr;
            }
            cudaEvent_t tr0 = nullptr, tr1 = nullptr;
            if (trace && k == 0) { cudaEventCreate(&tr0); cudaEventCreate(&tr1); cudaEventRecord(tr0, cst); }
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(dbuf[b], hx, m * sizeof(double), cudaMemcpyHostToDevice, cst);
            if (e == cudaSuccess && w)
                e = cudaMemcpyAsync(dbuf[b] + cw, hw_, m * sizeof(double), cudaMemcpyHostToDevice, cst);
            if (tr0) {
                cudaEventRecord(tr1, cst);
                cudaEventSynchronize(tr1);
                float ms_ = 0; cudaEventElapsedTime(&ms_, tr0, tr1);
                fprintf(stderr, " [first H2D %.0f us on the copy stream]", ms_ * 1e3);
                cudaEventDestroy(tr0); cudaEventDestroy(tr1);
            }
            if (e == cudaSuccess) e = cudaEventRecord(copied[b], cst);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(st, copied[b], 0);
            if (e == cudaSuccess)
                e = parallel_accumulate(t->s.v, dbuf[b], w ? dbuf[b] + cw : nullptr, wsc, m,
                                        t->scratch, sm_count(), st, k == 0, sparts, (long)off);
            if (e == cudaSuccess) e = cudaEventRecord(done[b], st);
        }
        if (e == cudaSuccess && bad_weights && w)
            e = parallel_bad_weights(t->scratch, st, bad_weights);
        if (e == cudaSuccess && !(bad_weights && *bad_weights)) {
            e = parallel_finalize(t->s.v, 0, t->scratch, sm_count(), st);
            if (e == cudaSuccess && stats) e = stats_fused_finish(stats, sm_count(), st);
        }
    }
    if (stats) crick_stats_release_stream(stats, st);
    if (e != cudaSuccess) rc = fail("host-chunked parallel update", e);
    const auto t4 = now();
    // buffers are freed only after the work that reads them has finished
    cudaStreamSynchronize(st);
    const auto t5 = now();
    if (cst) cudaStreamSynchronize(cst);
    for (int b = 0; b < nbuf; ++b)
        if (dbuf[b]) cudaFreeAsync(dbuf[b], st);
    if (staged) stage_release();
    if (trace) fprintf(stderr, ", staging memcpy %.0f us, loop total %.0f us, final sync %.0f us, teardown %.0f us\n",
                       t_stage, us(t2, t4), us(t4, t5), us(t5, now()));
    if (rc == 0) t->buf_clean = true;   // flushed above, and k_finalize leaves no buffer
    return rc;
}

static int td_update(crick_tdigest_t* t, const double* x, const double* w, double w_scalar,
                     int64_t n, int x_on_device, int w_on_device, int mode, void* stream,
                     int* bad_weights, crick_stats* stats, bool defer_check = false);

int crick_tdigest_update(crick_tdigest_t* t, const double* x, const double* w, double w_scalar,
                         int64_t n, int x_on_device, int w_on_device, int mode, void* stream) {
    return td_update(t, x, w, w_scalar, n, x_on_device, w_on_device, mode, stream, nullptr, nullptr);
}

int crick_tdigest_update_checked(crick_tdigest_t* t, const double* x, const double* w,
                                 double w_scalar, int64_t n, int x_on_device, int w_on_device,
                                 int mode, void* stream, int* bad_weights) {
    if (!bad_weights) return fail_msg("update_checked: bad_weights must not be NULL");
    *bad_weights = 0;
    return td_update(t, x, w, w_scalar, n, x_on_device, w_on_device, mode, stream, bad_weights,
                     nullptr);
}

int crick_tdigest_update_stats(crick_tdigest_t* t, crick_stats_t* stats, const double* x,
                               const double* w, double w_scalar, int64_t n, int x_on_device,
                               int w_on_device, void* stream, int* bad_weights) {
    if (!stats) return fail_msg("update_stats: stats must not be NULL");
    if (n < CRICK_PARALLEL_MIN) return fail_msg("update_stats needs n >= CRICK_PARALLEL_MIN");
    if (bad_weights) *bad_weights = 0;
    return td_update(t, x, w, w_scalar, n, x_on_device, w_on_device, CRICK_MODE_PARALLEL, stream,
                     bad_weights, stats);
}

int crick_tdigest_update_deferred(crick_tdigest_t* t, const double* x, const double* w,
                                  double w_scalar, int64_t n, int mode, void* stream) {
    const bool parallel = mode == CRICK_MODE_PARALLEL ||
                          (mode == CRICK_MODE_AUTO && n >= (int64_t)CRICK_PARALLEL_MIN);
    if (!parallel) return fail_msg("update_deferred: the deferred weight check exists in parallel mode only");
    return td_update(t, x, w, w_scalar, n, 1, 1, mode, stream, nullptr, nullptr, true);
}

int crick_tdigest_take_bad_weights(crick_tdigest_t* t) {
    if (!t->scratch && !t->pipe_scratch[0]) return 0;
    int bad = 0;
    cudaStream_t st = td_stream(t, nullptr);      // (behind the tail stream of pipelined updates)
    void* sets[3] = {t->scratch, t->pipe_scratch[0], t->pipe_scratch[1]};
    for (void* sc : sets) {
        if (!sc) continue;
        int b = 0;
        cudaError_t e = parallel_take_bad_weights(sc, st, &b);
        if (e != cudaSuccess) return fail("take_bad_weights", e);
        bad |= b;
    }
    return bad;      // bit 0: bad weights; bit 1: a sharded update timed out waiting for a rank
}

// Pipelined parallel-mode update from device arrays.  The sample / edges kernels and the streaming
// pass -- which read only x, w and a scratch set -- are enqueued on `stream`; the tail that folds the
// bucket sums into the digest (k_finalize: one CTA, ~20 us) goes to `tail_stream` behind an event.
// A following pipelined update therefore starts its ~50 us of sample work while the previous tail
// (and whatever the caller enqueues behind it on `tail_stream`: export, all-gather, merge) is still
// running; with back-to-back updates of 1.25e8 values the step shrinks from pass + ~95 us to
// pass + ~55 us.  Every other operation on the handle orders itself behind the tail stream as usual
// (td_stream).  The streaming kernel runs on sm_count - crick_pipeline_reserve_sms() CTAs (default 2
// fewer): the per-CTA partial sums are cut differently, so the digest agrees with
// crick_tdigest_update_deferred's in the last bits only when the reserve is 0 -- deterministic either
// way.  Weights are validated like there (crick_tdigest_take_bad_weights).
// SMs the streaming kernel of a pipelined update leaves free (it runs one CTA per SM and holds that
// SM's shared memory for the whole pass): the tail stream's kernels -- k_finalize, and the caller's
// export / NCCL all-gather / merge, which can sit waiting for other ranks -- would otherwise take
// SMs first and make the pass wait for them (measured at 8 GPUs: pass 0.32 -> 0.40 ms).
static int g_pipe_reserve = [] {
    const char* env = getenv("CRICK_PIPE_RESERVE");
    int v = env ? atoi(env) : 2;
    return v < 0 ? 0 : (v > 32 ? 32 : v);
}();
int crick_pipeline_reserve_sms(int n_sms) {
    int prev = g_pipe_reserve;
    if (n_sms >= 0 && n_sms <= 32) g_pipe_reserve = n_sms;
    return prev;
}

// Overlapped sample phase.  With it the streaming passes of consecutive pipelined updates run back to
// back on a stream of the handle's own, and the sample -> edges kernels of update k + 1 -- which need
// only its x and w -- run on `stream` WHILE the pass of update k streams, on the SMs the reserve
// leaves free (raise the reserve with it: 144 sort CTAs on 2 SMs would take longer than the pass).
// A step then costs the pass plus a kernel boundary instead of pass + ~50 us.  x and w of an update
// must stay valid until `tail_stream` has passed it (the plain pipelined form frees them with `stream`).
static int g_pipe_overlap = [] {
    const char* env = getenv("CRICK_PIPE_OVERLAP");
    return env ? (atoi(env) != 0) : 0;
}();
int crick_pipeline_overlap(int on) {
    int prev = g_pipe_overlap;
    if (on == 0 || on == 1) g_pipe_overlap = on;
    return prev;
}

int crick_tdigest_update_pipelined(crick_tdigest_t* t, const double* x, const double* w, double w_scalar,
                                   int64_t n, void* stream, void* tail_stream) {
    if (!stream || !tail_stream || stream == tail_stream)
        return fail_msg("update_pipelined: needs two different explicit streams");
    if (n <= 0) return 0;
    if (n < CRICK_PARALLEL_MIN) return fail_msg("update_pipelined needs n >= CRICK_PARALLEL_MIN");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // everything that touches the digest itself happens on the tail stream, in the handle's order
    cudaStream_t ts = td_stream(t, tail_stream);
    if (td_drain(t, ts)) return -1;
    const bool clean = t->buf_clean;
    t->mirror_ok = false; t->buf_clean = false;
    if (!w && w_scalar <= DBL_EPSILON) { t->buf_clean = clean; return 0; }      // :285 drops every sample
    cudaError_t e = clean ? cudaSuccess : launch_flush(t->s.v, 0, 1, 1, ts);
    if (e != cudaSuccess) return fail("update_pipelined: flush", e);
    const int p = t->pipe_next;
    t->pipe_next ^= 1;
    if (!t->ev_pass) {
        e = cudaEventCreateWithFlags(&t->ev_pass, cudaEventDisableTiming);
        for (int b = 0; b < 2 && e == cudaSuccess; ++b) e = cudaEventCreateWithFlags(&t->ev_fin[b], cudaEventDisableTiming);
        if (e != cudaSuccess) return fail("update_pipelined: events", e);
    }
    if (!t->pipe_scratch[p]) {
        const size_t b = parallel_scratch_bytes(t->s.v.comp, sm_count());
        t->pipe_scratch[p] = crick_device_alloc((int64_t)b);
        if (!t->pipe_scratch[p]) return -1;
        e = cudaMemsetAsync(t->pipe_scratch[p], 0, 256, st);       // header: sticky flags start cleared
        if (e != cudaSuccess) return fail("update_pipelined: scratch", e);
    }
    // set p was last read by the tail of the update before the previous one
    if (t->fin_recorded[p]) {
        e = cudaStreamWaitEvent(st, t->ev_fin[p], 0);
        if (e != cudaSuccess) return fail("update_pipelined: wait", e);
    }
    const int sms = sm_count() > 2 * g_pipe_reserve ? sm_count() - g_pipe_reserve : sm_count();
    if (g_pipe_overlap) {
        if (!t->pass_stream) {
            e = cudaStreamCreateWithFlags(&t->pass_stream, cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_prep, cudaEventDisableTiming);
            if (e != cudaSuccess) return fail("update_pipelined: pass stream", e);
        }
        e = launch_parallel_prep(t->s.v, x, w, w_scalar, n, t->pipe_scratch[p], sms, st);
        if (e == cudaSuccess) e = cudaEventRecord(t->ev_prep, st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(t->pass_stream, t->ev_prep, 0);
        if (e == cudaSuccess)
            e = launch_parallel_accumulate(t->s.v, x, w, w_scalar, n, t->pipe_scratch[p], sms, t->pass_stream);
        if (e == cudaSuccess) e = cudaEventRecord(t->ev_pass, t->pass_stream);
    } else {
        e = launch_parallel_pass(t->s.v, x, w, w_scalar, n, t->pipe_scratch[p], sms, st);
        if (e == cudaSuccess) e = cudaEventRecord(t->ev_pass, st);
    }
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ts, t->ev_pass, 0);
    if (e == cudaSuccess) e = launch_parallel_tail(t->s.v, 0, t->pipe_scratch[p], sms, ts, w != nullptr);
    if (e == cudaSuccess) e = cudaEventRecord(t->ev_fin[p], ts);
    if (e != cudaSuccess) return fail("update_pipelined", e);
    t->fin_recorded[p] = true;
    t->buf_clean = true;
    return 0;
}

// ---- sharded update over NVLink peer memory -----------------------------------------------------
struct crick_p2p {
    int rank = 0, world = 1;
    void* buf[16] = {};
    unsigned long long step = 0;
};

int64_t crick_p2p_buffer_bytes(double compression, int world) {
    if (compression < 20) compression = 20; else if (compression > 1000) compression = 1000;
    return (int64_t)p2p_buffer_bytes(compression, world);
}
void* crick_p2p_alloc(int64_t bytes) {
    void* p = nullptr;                       // plain cudaMalloc: the block is exported through CUDA IPC
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fail("crick_p2p_alloc", e); if (p) cudaFree(p); return nullptr; }
    return p;
}
void crick_p2p_free(void* p) { if (p) cudaFree(p); }
int crick_ipc_get_handle(void* dev_ptr, void* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaError_t e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), dev_ptr);
    return e == cudaSuccess ? 0 : fail("cudaIpcGetMemHandle", e);
}
void* crick_ipc_open_handle(const void* handle64) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { fail("cudaIpcOpenMemHandle", e); return nullptr; }
    return p;
}
int crick_ipc_close_handle(void* p) {
    cudaError_t e = cudaIpcCloseMemHandle(p);
    return e == cudaSuccess ? 0 : fail("cudaIpcCloseMemHandle", e);
}
crick_p2p_t* crick_p2p_new(int rank, int world, void* const* buffers) {
    if (world < 1 || world > 16 || rank < 0 || rank >= world) { fail_msg("crick_p2p_new: bad rank / world"); return nullptr; }
    crick_p2p* c = new (std::nothrow) crick_p2p();
    if (!c) return nullptr;
    c->rank = rank; c->world = world;
    for (int r = 0; r < world; ++r) {
        if (!buffers[r]) { fail_msg("crick_p2p_new: null buffer"); delete c; return nullptr; }
        c->buf[r] = buffers[r];
    }
    return c;
}
void crick_p2p_destroy(crick_p2p_t* c) { delete c; }

int crick_tdigest_update_exchange(crick_tdigest_t* t, crick_p2p_t* c, const double* x, const double* w,
                                  double w_scalar, int64_t n, void* stream) {
    if (!c) return fail_msg("update_exchange: no communicator");
    if (n > 0 && n < CRICK_PARALLEL_MIN) return fail_msg("update_exchange needs 0 or >= CRICK_PARALLEL_MIN values per rank");
    cudaStream_t st = td_stream(t, stream);
    if (td_drain(t, st)) return -1;
    const bool clean = t->buf_clean;
    t->mirror_ok = false; t->buf_clean = false;
    if (td_scratch(t)) return -1;
    cudaError_t e = clean ? cudaSuccess : launch_flush(t->s.v, 0, 1, 1, st);
    // a scalar weight tdigest_add drops (:285) drops this rank's shard; the exchange still happens
    const int64_t n_eff = (!w && w_scalar <= DBL_EPSILON) ? 0 : n;
    if (e == cudaSuccess)
        e = launch_parallel_update_exchange(t->s.v, 0, x, w, w_scalar, n_eff, t->scratch, sm_count(), st, true,
                                            c->rank, c->world, c->buf, ++c->step);
    if (e == cudaSuccess) t->buf_clean = true;
    return e == cudaSuccess ? 0 : fail("update_exchange", e);
}

static int td_update(crick_tdigest_t* t, const double* x, const double* w, double w_scalar,
                     int64_t n, int x_on_device, int w_on_device, int mode, void* stream,
                     int* bad_weights, crick_stats* stats, bool defer_check) {
    if (n <= 0) return 0;  // tdigest_stubs.c:648-650
    cudaStream_t st = td_stream(t, stream);
    if (td_drain(t, st)) return -1;
    const bool clean = t->buf_clean;   // after the drain: pending adds go through the buffer
    t->mirror_ok = false; t->buf_clean = false;
    bool parallel = mode == CRICK_MODE_PARALLEL ||
                    (mode == CRICK_MODE_AUTO && n >= (int64_t)CRICK_PARALLEL_MIN);
    if (mode == CRICK_MODE_PARALLEL && n < CRICK_PARALLEL_MIN)
        return fail_msg("parallel mode needs n >= CRICK_PARALLEL_MIN");
    const bool w_arr = w != nullptr;
    if (w_arr && (x_on_device != w_on_device))
        return fail_msg("x and w must both be host or both be device arrays");
    // a scalar weight tdigest_add drops (w <= DBL_EPSILON, tdigest_stubs.c:285) drops every
    // sample: the digest stays as it is (the parallel pass only tests weights it loads)
    if (!w_arr && w_scalar <= DBL_EPSILON) {
        t->buf_clean = clean;
        // (the fused moments do not depend on the weight: stats.update(x))
        return stats ? crick_stats_update(stats, x, 0, nullptr, 1, n, x_on_device, 0, stream) : 0;
    }
    if (!x_on_device) {
        if (parallel) return td_update_host_parallel(t, x, w, w_scalar, n, st, bad_weights, stats, clean);
        // reference mode from host memory: stage the whole (small) input
        double* d = nullptr;
        size_t bytes = (size_t)n * sizeof(double) * (w_arr ? 2 : 1);
        cudaError_t e = cudaMallocAsync(&d, bytes, st);
        if (e != cudaSuccess) return fail("cudaMallocAsync(stage)", e);
        e = cudaMemcpyAsync(d, x, n * sizeof(double), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess && w_arr)
            e = cudaMemcpyAsync(d + n, w, n * sizeof(double), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess)
            e = launch_ingest(t->s.v, 1, d, w_arr ? d + n : nullptr, w_scalar, 1, 1, nullptr, n, 0,
                              nullptr, 0, st);
        cudaFreeAsync(d, st);
        if (e != cudaSuccess) return fail("reference-mode update", e);
        // the source may be pageable or reused by the caller right away
        e = cudaStreamSynchronize(st);
        return e == cudaSuccess ? 0 : fail("update sync", e);
    }
    cudaError_t e;
    if (parallel) {
        if (td_scratch(t)) return -1;
        e = clean ? cudaSuccess : launch_flush(t->s.v, 0, 1, 1, st);
        StatsPart* sparts = nullptr;
        int snparts = 0;
        if (stats && stats_fused_begin(stats, st, &sparts, &snparts)) return -1;
        if (e == cudaSuccess)
            e = launch_parallel_update(t->s.v, 0, x, w, w_scalar, n, t->scratch, sm_count(), st,
                                       bad_weights, sparts, defer_check);
        if (e == cudaSuccess && stats && !(bad_weights && *bad_weights))
            e = stats_fused_finish(stats, sm_count(), st);
        if (stats) crick_stats_release_stream(stats, st);
        if (e == cudaSuccess) t->buf_clean = true;   // flushed above, and k_finalize leaves no buffer
    } else {
        e = launch_ingest(t->s.v, 1, x, w, w_scalar, 1, 1, nullptr, n, 0, nullptr, 0, st);
    }
    return e == cudaSuccess ? 0 : fail("update", e);
}

int crick_tdigest_flush(crick_tdigest_t* t) {
    cudaStream_t st = td_stream(t, nullptr);
    if (td_drain(t, st)) return -1;
    t->mirror_ok = false; t->buf_clean = false;
    cudaError_t e = launch_flush(t->s.v, 0, 1, 1, st);
    if (e == cudaSuccess) t->buf_clean = true;
    return e == cudaSuccess ? 0 : fail("flush", e);
}

int crick_tdigest_merge(crick_tdigest_t* t, crick_tdigest_t* const* others, int n_others) {
    cudaStream_t st = td_stream(t, nullptr);
    if (td_drain(t, st)) return -1;
    t->mirror_ok = false; t->buf_clean = false;
    for (int i = 0; i < n_others; ++i) {
        crick_tdigest* o = others[i];
        cudaStream_t ost = td_stream(o, nullptr);
        if (td_drain(o, ost)) return -1;
        o->mirror_ok = false; o->buf_clean = false;
        cudaError_t e = launch_flush(o->s.v, 0, 1, 1, ost);  // tdigest_merge flushes `other` (:595)
        if (e == cudaSuccess && o != t) {
            e = cudaEventRecord(o->ev, ost);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(st, o->ev, 0);
        }
        if (e == cudaSuccess) e = launch_merge_pairs(t->s.v, o->s.v, 0, 0, 0, 0, 1, st);
        if (e == cudaSuccess && o != t) {
            // `o` must not be mutated/freed before the merge has read it
            e = cudaEventRecord(t->ev, st);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(ost, t->ev, 0);
        }
        if (e != cudaSuccess) return fail("merge", e);
    }
    return 0;
}

int crick_tdigest_scale(crick_tdigest_t* t, double factor) {
    cudaStream_t st = td_stream(t, nullptr);
    if (td_drain(t, st)) return -1;
    t->mirror_ok = false; t->buf_clean = false;
    cudaError_t e = launch_flush(t->s.v, 0, 1, 1, st);
    if (e == cudaSuccess) e = launch_scale(t->s.v, 0, factor, st);
    return e == cudaSuccess ? 0 : fail("scale", e);
}

static int td_query(crick_tdigest* t, bool cdf, const double* in, double* out, int64_t n,
                    int on_device, void* stream) {
    cudaStream_t st = td_stream(t, stream);
    if (td_drain(t, st)) return -1;
    t->mirror_ok = false; t->buf_clean = false;  // query_prep flushes (tdigest_stubs.c:307)
    cudaError_t e = launch_query_prep(t->s.v, 0, 1, st);
    if (e != cudaSuccess) return fail("query prep", e);
    if (n <= 0) return 0;
    if (on_device) {
        e = launch_query(t->s.v, 0, cdf, in, out, n, sm_count(), st);
        return e == cudaSuccess ? 0 : fail("query", e);
    }
    double* d = nullptr;
    e = cudaMallocAsync(&d, 2 * (size_t)n * sizeof(double), st);
    if (e != cudaSuccess) return fail("cudaMallocAsync(query)", e);
    e = cudaMemcpyAsync(d, in, n * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = launch_query(t->s.v, 0, cdf, d, d + n, n, sm_count(), st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d + n, n * sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaFreeAsync(d, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    return e == cudaSuccess ? 0 : fail("query", e);
}

int crick_tdigest_quantile(crick_tdigest_t* t, const double* q, double* out, int64_t n,
                           int on_device, void* stream) {
    return td_query(t, false, q, out, n, on_device, stream);
}
int crick_tdigest_cdf(crick_tdigest_t* t, const double* x, double* out, int64_t n, int on_device,
                      void* stream) {
    return td_query(t, true, x, out, n, on_device, stream);
}

double crick_tdigest_compression(crick_tdigest_t* t) { return t->s.v.comp; }
int crick_tdigest_size(crick_tdigest_t* t) { return t->s.v.cap; }
int crick_tdigest_buffer_size(crick_tdigest_t* t) { return t->s.v.bufcap; }
double crick_tdigest_min(crick_tdigest_t* t) { return td_sync_mirror(t) ? NAN : t->mirror.vmin; }
double crick_tdigest_max(crick_tdigest_t* t) { return td_sync_mirror(t) ? NAN : t->mirror.vmax; }
int crick_tdigest_ncentroids(crick_tdigest_t* t) { return td_sync_mirror(t) ? -1 : t->mirror.n; }
int crick_tdigest_buffer_ncentroids(crick_tdigest_t* t) { return td_sync_mirror(t) ? -1 : t->mirror.bn; }
double crick_tdigest_total_weight(crick_tdigest_t* t) { return td_sync_mirror(t) ? NAN : t->mirror.total; }
double crick_tdigest_buffer_total_weight(crick_tdigest_t* t) {
    return td_sync_mirror(t) ? NAN : t->mirror.btotal;
}

int crick_tdigest_get_centroids(crick_tdigest_t* t, double* out, int on_device) {
    if (td_sync_mirror(t)) return -1;
    int n = t->mirror.n;
    if (n <= 0) return 0;
    cudaStream_t st = td_stream(t, nullptr);
    cudaError_t e = cudaMemcpyAsync(out, t->s.v.cen, (size_t)n * sizeof(double2),
                                    on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    return e == cudaSuccess ? n : fail("get_centroids", e);
}

const double* crick_tdigest_centroids_device(crick_tdigest_t* t) {
    return reinterpret_cast<const double*>(t->s.v.cen);
}

int crick_tdigest_set_state(crick_tdigest_t* t, const double* centroids, int n, double total_weight,
                            double vmin, double vmax) {
    if (n < 0 || n > t->s.v.cap) return fail_msg("set_state: more centroids than 2*ceil(compression)");
    if (td_sync_mirror(t)) return -1;
    cudaStream_t st = td_stream(t, nullptr);
    DevHdr h = t->mirror;
    h.total = total_weight; h.vmin = vmin; h.vmax = vmax;
    if (n > 0) h.n = n;  // tdigest.pyx:260-263
    cudaError_t e = cudaSuccess;
    if (n > 0)
        e = cudaMemcpyAsync(t->s.v.cen, centroids, (size_t)n * sizeof(double2), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(t->s.v.hdr, &h, sizeof(h), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail("set_state", e);
    t->mirror = h;
    return 0;
}

int crick_tdigest_reset(crick_tdigest_t* t, void* stream) {
    td_stream(t, stream, true);
    t->px.clear();
    t->pw.clear();
    t->mirror_ok = false;
    // the header is rewritten by whatever touches the digest next (td_stream), or not at all when
    // that is the one-pass record merge, which starts from an empty digest by itself: one launch
    // less on the critical path of a multi-GPU step
    t->lazy_reset = true;
    t->buf_clean = true;                  // an empty digest has an empty buffer
    return 0;
}

int crick_tdigest_debug_edges(crick_tdigest_t* t, double* out, int max_out) {
    if (!t->scratch) return 0;
    return parallel_debug_edges(t->scratch, out, max_out, td_stream(t, nullptr));
}

int crick_tdigest_debug_header(crick_tdigest_t* t, void* out256) {
    if (!t->scratch) return -1;
    return parallel_debug_header(t->scratch, out256, td_stream(t, nullptr));
}

int crick_tdigest_debug_buckets(crick_tdigest_t* t, double* out_w, double* out_s, double* out_minmax,
                                int max_out) {
    if (!t->scratch) return 0;
    return parallel_debug_buckets(t->scratch, t->s.v.comp, sm_count(), out_w, out_s, out_minmax,
                                  max_out, td_stream(t, nullptr));
}

int crick_tdigest_record_doubles(crick_tdigest_t* t) { return 4 + 2 * t->s.v.cap; }

int crick_tdigest_export_record(crick_tdigest_t* t, double* record_dev, void* stream) {
    cudaStream_t st = td_stream(t, stream);
    if (td_drain(t, st)) return -1;
    // (no flush launch when the host knows the insertion buffer is empty, e.g. right after a
    // parallel-mode update: one launch less on the critical path of a multi-GPU step)
    const bool clean = t->buf_clean;
    cudaError_t e = clean ? cudaSuccess : launch_flush(t->s.v, 0, 1, 1, st);
    if (!clean) t->mirror_ok = false;
    t->buf_clean = e == cudaSuccess;
    if (e == cudaSuccess) e = launch_pack_record(t->s.v, 0, record_dev, st);
    return e == cudaSuccess ? 0 : fail("export_record", e);
}

int crick_tdigest_merge_records(crick_tdigest_t* t, const double* records_dev, int n_records, int skip,
                                void* stream) {
    cudaStream_t st = td_stream(t, stream);
    if (td_drain(t, st)) return -1;
    t->mirror_ok = false; t->buf_clean = false;
    cudaError_t e = launch_merge_records(t->s.v, 0, records_dev, n_records, skip, st);
    return e == cudaSuccess ? 0 : fail("merge_records", e);
}

int crick_tdigest_merge_records_bulk(crick_tdigest_t* t, const double* records_dev, int n_records,
                                     void* stream) {
    const bool fresh = t->lazy_reset && t->px.empty();
    cudaStream_t st = td_stream(t, stream, fresh);
    if (td_drain(t, st)) return -1;
    const bool clean = t->buf_clean;
    t->mirror_ok = false; t->buf_clean = false;
    cudaError_t e = clean ? cudaSuccess : launch_flush(t->s.v, 0, 1, 1, st);
    if (e == cudaSuccess) {
        e = launch_merge_records_bulk(t->s.v, 0, records_dev, n_records, st, fresh);
        if (e == cudaErrorInvalidValue) {   // union too large for one CTA: sequential merge
            cudaGetLastError();
            e = fresh ? launch_init_hdr(t->s.v.hdr, 1, st) : cudaSuccess;
            if (e == cudaSuccess) e = launch_merge_records(t->s.v, 0, records_dev, n_records, -1, st);
        }
        t->lazy_reset = false;
    }
    return e == cudaSuccess ? 0 : fail("merge_records_bulk", e);
}

}  // extern "C"

// ---------------------------------------------------------------------------
// grouped digests
struct crick_tdigest_group {
    DigestStore s;
    cudaStream_t own = nullptr;
    cudaStream_t last = nullptr;
    cudaEvent_t ev = nullptr;
    // what may be waiting in the buffers: 0 nothing, 1 items that all weigh pend_w, 2 anything
    int pend = 0;
    double pend_w = 0.0;
};

static cudaStream_t grp_stream(crick_tdigest_group* g, void* stream) {
    cudaStream_t st = stream ? reinterpret_cast<cudaStream_t>(stream) : g->own;
    if (g->last && g->last != st) {
        cudaEventRecord(g->ev, g->last);
        cudaStreamWaitEvent(st, g->ev, 0);
    }
    g->last = st;
    return st;
}

extern "C" {

crick_tdigest_group_t* crick_tdigest_group_new(double compression, int64_t n_groups) {
    if (!sm_count()) { fail_msg("no CUDA device: libcrick_cuda has no CPU fallback"); return nullptr; }
    if (n_groups <= 0) { fail_msg("n_groups must be > 0"); return nullptr; }
    crick_tdigest_group* g = new (std::nothrow) crick_tdigest_group();
    if (!g) return nullptr;
    if (store_alloc(g->s, compression, n_groups)) { delete g; return nullptr; }
    cudaError_t e = cudaStreamCreateWithFlags(&g->own, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = launch_init_hdr(g->s.v.hdr, n_groups, g->own);
    if (e != cudaSuccess) { fail("tdigest_group_new", e); crick_tdigest_group_free(g); return nullptr; }
    g->last = g->own;
    return g;
}

void crick_tdigest_group_free(crick_tdigest_group_t* g) {
    if (!g) return;
    if (g->own) { cudaStreamSynchronize(g->own); cudaStreamDestroy(g->own); }
    if (g->ev) cudaEventDestroy(g->ev);
    store_free(g->s);
    delete g;
}

int64_t crick_tdigest_group_count(crick_tdigest_group_t* g) { return g->s.nd; }

int crick_tdigest_group_update(crick_tdigest_group_t* g, const double* values, const double* weights,
                               double w_scalar, const int64_t* offsets, int64_t row_len,
                               int64_t row_stride, int on_device, void* stream) {
    cudaStream_t st = grp_stream(g, stream);
    const long ng = g->s.nd;
    static_assert(sizeof(long) == sizeof(int64_t), "LP64 expected");
    const bool uniform = weights == nullptr && (g->pend == 0 || (g->pend == 1 && g->pend_w == w_scalar));
    g->pend = uniform ? 1 : 2;
    g->pend_w = w_scalar;
    if (on_device) {
        cudaError_t e = launch_ingest(g->s.v, ng, values, weights, w_scalar, 1, 1,
                                      reinterpret_cast<const long*>(offsets), row_len, row_stride,
                                      nullptr, 0, st, uniform);
        return e == cudaSuccess ? 0 : fail("group_update", e);
    }
    // host input: stage values (+weights, +offsets) once
    int64_t total;
    if (offsets) total = offsets[ng];
    else total = (ng - 1) * row_stride + row_len;
    if (total <= 0) return 0;
    double* d = nullptr;
    long* doff = nullptr;
    size_t bytes = (size_t)total * sizeof(double) * (weights ? 2 : 1);
    cudaError_t e = cudaMallocAsync(&d, bytes, st);
    if (e != cudaSuccess) return fail("cudaMallocAsync(group stage)", e);
    e = staged_copy_h2d(d, values, total, st);
    if (e == cudaSuccess && weights) e = staged_copy_h2d(d + total, weights, total, st);
    if (e == cudaSuccess && offsets) {
        e = cudaMallocAsync(&doff, (size_t)(ng + 1) * sizeof(long), st);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(doff, offsets, (size_t)(ng + 1) * sizeof(long), cudaMemcpyHostToDevice, st);
    }
    if (e == cudaSuccess)
        e = launch_ingest(g->s.v, ng, d, weights ? d + total : nullptr, w_scalar, 1, 1, doff, row_len,
                          row_stride, nullptr, 0, st, uniform);
    cudaFreeAsync(d, st);
    if (doff) cudaFreeAsync(doff, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    return e == cudaSuccess ? 0 : fail("group_update", e);
}

int crick_tdigest_group_flush(crick_tdigest_group_t* g) {
    cudaStream_t st = grp_stream(g, nullptr);
    cudaError_t e = launch_flush(g->s.v, 0, 1, g->s.nd, st);
    if (e == cudaSuccess) g->pend = 0;
    return e == cudaSuccess ? 0 : fail("group_flush", e);
}

int crick_tdigest_group_merge_tree(crick_tdigest_group_t* g, crick_tdigest_t* out) {
    if (out->s.v.cap != g->s.v.cap || out->s.v.bufcap != g->s.v.bufcap)
        return fail_msg("merge_tree: output digest has a different compression");
    g->pend = 2;   // merges feed centroids through the buffers (tdigest_merge :592-606)
    cudaStream_t st = grp_stream(g, nullptr);
    const long n = g->s.nd;
    cudaError_t e = cudaSuccess;
    for (long s = 1; s < n && e == cudaSuccess; s <<= 1) {
        long count = (n - s + 2 * s - 1) / (2 * s);  // pairs (i, i+s), i = 0, 2s, 4s, ... with i+s < n
        e = launch_flush(g->s.v, s, 2 * s, count, st);
        if (e == cudaSuccess) e = launch_merge_pairs(g->s.v, g->s.v, 0, 2 * s, s, 2 * s, count, st);
    }
    if (e != cudaSuccess) return fail("merge_tree", e);
    cudaStream_t ost = td_stream(out, nullptr);
    if (td_drain(out, ost)) return -1;
    out->mirror_ok = false; out->buf_clean = false;
    e = cudaEventRecord(g->ev, st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ost, g->ev, 0);
    if (e == cudaSuccess) e = launch_copy_digest(out->s.v, 0, g->s.v, 0, ost);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ost);
    return e == cudaSuccess ? 0 : fail("merge_tree copy", e);
}

// All digests of the group merged into `out` in ONE pass: their centroids are (mean, weight)
// points -- exactly what tdigest_merge feeds to tdigest_add (:600-602) -- so the union goes through
// the parallel k-bucket update (sample -> k-scale edges -> one streaming pass -> one greedy
// compression) instead of 16 levels of dependent pairwise merges.  Not bit-identical with nested
// merge() calls (that is crick_tdigest_group_merge_tree); same accuracy class.
int crick_tdigest_group_merge_all(crick_tdigest_group_t* g, crick_tdigest_t* out) {
    if (out->s.v.cap != g->s.v.cap || out->s.v.bufcap != g->s.v.bufcap)
        return fail_msg("merge_all: output digest has a different compression");
    cudaStream_t st = grp_stream(g, nullptr);
    const long ng = g->s.nd;
    cudaError_t e = launch_flush(g->s.v, 0, 1, ng, st);      // tdigest_merge flushes `other` (:595)
    if (e == cudaSuccess) g->pend = 0;
    long* off = nullptr;
    double *mm = nullptr, *xw = nullptr;
    if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&off), (size_t)(ng + 1) * 8 + 64, st);
    if (e != cudaSuccess) return fail("merge_all", e);
    mm = reinterpret_cast<double*>(off + ng + 2);
    e = launch_group_offsets(g->s.v, ng, off, mm, st);
    long total = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&total, off + ng, sizeof(long), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    int rc = 0;
    if (e == cudaSuccess && total > 0) {
        const long tw = (total + 1) & ~1L;                   // W stays 16-byte aligned
        e = cudaMallocAsync(reinterpret_cast<void**>(&xw), (size_t)tw * 16, st);
        if (e == cudaSuccess) e = launch_group_gather(g->s.v, ng, off, xw, xw + tw, st);
        if (e == cudaSuccess) {
            // `out` works on its own stream: order it behind the gather, and the frees behind it
            cudaStream_t ost = td_stream(out, nullptr);
            e = cudaEventRecord(g->ev, st);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(ost, g->ev, 0);
            if (e == cudaSuccess) {
                // fewer points than the sampler wants: the reference's own sequential algorithm
                rc = td_update(out, xw, xw + tw, 0.0, total, 1, 1,
                               total >= CRICK_PARALLEL_MIN ? CRICK_MODE_PARALLEL : CRICK_MODE_REFERENCE,
                               nullptr, nullptr, nullptr);
                if (rc == 0) e = launch_widen_minmax(out->s.v, 0, mm, ost);
                out->mirror_ok = false;
                if (e == cudaSuccess) e = cudaEventRecord(out->ev, ost);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(st, out->ev, 0);
            }
        }
    }
    if (xw) cudaFreeAsync(xw, st);
    cudaFreeAsync(off, st);
    if (rc) return rc;
    return e == cudaSuccess ? 0 : fail("merge_all", e);
}

static int grp_query(crick_tdigest_group* g, bool cdf, const double* in, int nq, double* out,
                     int on_device, void* stream) {
    cudaStream_t st = grp_stream(g, stream);
    const long ng = g->s.nd;
    cudaError_t e = launch_query_prep(g->s.v, 0, ng, st);
    if (e != cudaSuccess) return fail("group query prep", e);
    if (nq <= 0) return 0;
    if (on_device) {
        e = launch_query_grouped(g->s.v, ng, cdf, in, out, nq, st);
        return e == cudaSuccess ? 0 : fail("group query", e);
    }
    double* d = nullptr;
    size_t outn = (size_t)ng * nq;
    e = cudaMallocAsync(&d, (nq + outn) * sizeof(double), st);
    if (e != cudaSuccess) return fail("cudaMallocAsync(group query)", e);
    e = cudaMemcpyAsync(d, in, nq * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = launch_query_grouped(g->s.v, ng, cdf, d, d + nq, nq, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d + nq, outn * sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaFreeAsync(d, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    return e == cudaSuccess ? 0 : fail("group query", e);
}

int crick_tdigest_group_quantile(crick_tdigest_group_t* g, const double* q, int nq, double* out,
                                 int on_device, void* stream) {
    return grp_query(g, false, q, nq, out, on_device, stream);
}
int crick_tdigest_group_cdf(crick_tdigest_group_t* g, const double* x, int nq, double* out,
                            int on_device, void* stream) {
    return grp_query(g, true, x, nq, out, on_device, stream);
}

int crick_tdigest_group_meta(crick_tdigest_group_t* g, double* out_meta) {
    cudaStream_t st = grp_stream(g, nullptr);
    const long ng = g->s.nd;
    double* d = nullptr;
    cudaError_t e = cudaMallocAsync(&d, (size_t)ng * 4 * sizeof(double), st);
    if (e != cudaSuccess) return fail("cudaMallocAsync(meta)", e);
    e = launch_group_meta(g->s.v, ng, d, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_meta, d, (size_t)ng * 4 * sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaFreeAsync(d, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    return e == cudaSuccess ? 0 : fail("group_meta", e);
}

int crick_tdigest_group_get_centroids(crick_tdigest_group_t* g, int64_t i, double* out) {
    if (i < 0 || i >= g->s.nd) return fail_msg("group index out of range");
    cudaStream_t st = grp_stream(g, nullptr);
    cudaError_t e = launch_flush(g->s.v, i, 1, 1, st);
    DevHdr h;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h, g->s.v.hdr + i, sizeof(h), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail("group_get_centroids", e);
    if (h.n > 0) {
        e = cudaMemcpyAsync(out, g->s.v.cen + i * (long)g->s.v.cap, (size_t)h.n * sizeof(double2),
                            cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return fail("group_get_centroids", e);
    }
    return h.n;
}

int crick_tdigest_group_extract(crick_tdigest_group_t* g, int64_t i, crick_tdigest_t* out) {
    if (i < 0 || i >= g->s.nd) return fail_msg("group index out of range");
    if (out->s.v.cap != g->s.v.cap || out->s.v.bufcap != g->s.v.bufcap)
        return fail_msg("extract: output digest has a different compression");
    cudaStream_t st = grp_stream(g, nullptr);
    cudaStream_t ost = td_stream(out, nullptr);
    if (td_drain(out, ost)) return -1;
    out->mirror_ok = false; out->buf_clean = false;
    cudaError_t e = cudaEventRecord(g->ev, st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ost, g->ev, 0);
    if (e == cudaSuccess) e = launch_copy_digest(out->s.v, 0, g->s.v, i, ost);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ost);
    return e == cudaSuccess ? 0 : fail("group_extract", e);
}

}  // extern "C"
