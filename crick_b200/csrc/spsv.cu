// crick_b200/csrc/spsv.cu -- SpaceSaving (Stream-Summary) for int64 / float64-bit-pattern items
// on the GPU.  Replaces crick/space_saving_stubs.c.in (int64 instantiation).
//
// The reference keeps a circular doubly-linked list sorted by (count desc, error asc) over a
// slot array, plus a khash map item -> slot.  The list order depends on the exact sequence of
// "walk backwards until a node >= me" steps (:121-138), so the reference-compatible path here
// replays the stream IN ORDER with one warp:
//   * the list is an explicit `order[]` array (position -> slot); the backwards walk is done
//     32 predecessors at a time (ballot), the re-insertion is a warp-wide shift;
//   * the map is an open-addressing table probed 32 slots at a time.
// State lives in HBM between calls and is staged in shared memory during a call when it
// fits (capacity <~ 2400), otherwise the kernel works on the HBM copy directly.
//
// Quirks kept: eviction ignores the weight (:229-231); float64 items are
// compared by bit pattern.  Quirk NOT kept: the dangling hash key the reference leaves behind
// on the merge's early `break` (:334/:345, segfault) -- nothing is inserted here.
#include "../../include/crick_cuda.h"
#include "api_common.h"
#include "stats_dev.cuh"

namespace crick {

struct SpsvHdr {
    long capacity;
    long size;
    int hcap;   // power of two >= 4*capacity
    int tomb;   // deleted markers currently in the table
    long seen;  // elements consumed so far = stream position of the next one
    int mixed;  // 1 after a user-level merge: `last` positions of two streams are mixed
    int pad;
};

struct SpsvView {
    SpsvHdr* hdr;
    long long* item;   // [capacity] slot arrays, creation order
    long long* count;
    long long* error;
    long long* last;   // [capacity] stream position of the slot's latest update (parallel path:
                       //            exact summaries are listed by (count desc, last asc) like
                       //            the reference's list)
    int* order;        // [capacity] list position -> slot (0 = head)
    int* pos;          // [capacity] slot -> list position
    long long* hkey;   // [hcap]
    int* hval;         // [hcap]  slot, -1 empty, -2 deleted
};

static size_t spsv_bytes(long cap, int hcap) {
    return (64 + (size_t)cap * (32 + 8) + (size_t)hcap * 12 + 64 + 15) & ~(size_t)15;
}

__host__ __device__ inline SpsvView spsv_carve(unsigned char* base, long cap, int hcap) {
    SpsvView v;
    v.hdr = reinterpret_cast<SpsvHdr*>(base);
    unsigned char* p = base + 64;
    v.item = reinterpret_cast<long long*>(p); p += cap * 8;
    v.count = reinterpret_cast<long long*>(p); p += cap * 8;
    v.error = reinterpret_cast<long long*>(p); p += cap * 8;
    v.last = reinterpret_cast<long long*>(p); p += cap * 8;
    v.hkey = reinterpret_cast<long long*>(p); p += (size_t)hcap * 8;
    v.order = reinterpret_cast<int*>(p); p += cap * 4;
    v.pos = reinterpret_cast<int*>(p); p += cap * 4;
    v.hval = reinterpret_cast<int*>(p);
    return v;
}

__device__ __forceinline__ unsigned hash64(long long k) {
    unsigned long long x = (unsigned long long)k;
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return (unsigned)x;
}

// working copy: header in registers (warp-uniform), arrays via pointers (shared or global)
struct SpsvW {
    long long *item, *count, *error, *last, *hkey;
    int *order, *pos, *hval;
    long capacity, size, seen;
    int hcap, tomb, mixed;
};

// warp-parallel probe; returns table index of `k` or -1
__device__ __forceinline__ int map_find(const SpsvW& w, long long k, int lane) {
    unsigned base = hash64(k) & (unsigned)(w.hcap - 1);
    for (int step = 0; step < w.hcap; step += 32) {
        int idx = (int)((base + step + lane) & (unsigned)(w.hcap - 1));
        int v = w.hval[idx];
        bool hit = v >= 0 && w.hkey[idx] == k;
        bool empty = v == -1;
        unsigned mh = __ballot_sync(CRICK_FULL, hit), me = __ballot_sync(CRICK_FULL, empty);
        // first event along the probe sequence decides
        int fh = mh ? __ffs(mh) - 1 : 64, fe = me ? __ffs(me) - 1 : 64;
        if (fh < fe) return (int)((base + step + fh) & (unsigned)(w.hcap - 1));
        if (fe < 64) return -1;
    }
    return -1;
}

__device__ __forceinline__ void map_put(SpsvW& w, long long k, int slot, int lane) {
    unsigned base = hash64(k) & (unsigned)(w.hcap - 1);
    for (int step = 0; step < w.hcap; step += 32) {
        int idx = (int)((base + step + lane) & (unsigned)(w.hcap - 1));
        int v = w.hval[idx];
        unsigned mf = __ballot_sync(CRICK_FULL, v < 0);
        if (mf) {
            int f = __ffs(mf) - 1;
            int tgt = (int)((base + step + f) & (unsigned)(w.hcap - 1));
            int old = w.hval[tgt];
            __syncwarp();
            if (lane == 0) { w.hkey[tgt] = k; w.hval[tgt] = slot; }
            if (old == -2) w.tomb--;
            __syncwarp();
            return;
        }
    }
}

__device__ __forceinline__ void map_rebuild(SpsvW& w, int lane) {
    for (int i = lane; i < w.hcap; i += 32) w.hval[i] = -1;
    __syncwarp();
    w.tomb = 0;
    for (long s = 0; s < w.size; ++s) map_put(w, w.item[s], (int)s, lane);
}

__device__ __forceinline__ void map_del(SpsvW& w, long long k, int lane) {
    int idx = map_find(w, k, lane);
    if (idx >= 0) {
        if (lane == 0) w.hval[idx] = -2;
        w.tomb++;
        __syncwarp();
    }
}

// counter_ge (:112-118) with offset
__device__ __forceinline__ bool ctr_ge(long long ac, long long ae, long long bc, long long be,
                                       long long off) {
    long long c = bc + off, e = be + off;
    return ac > c || (ac == c && ae <= e);
}

// Move the slot at list position p towards the head until its predecessor is >= it
// (counter_insert :121-138 / rebalance :167-183): exact linear-walk semantics, 32 at a time.
__device__ __forceinline__ void sift(SpsvW& w, int p, int lane) {
    const int s = w.order[p];
    const long long sc = w.count[s], se = w.error[s];
    int t = p;  // final position
    while (t > 0) {
        int j = t - 1 - lane;
        bool ge = true;  // positions before the head stop the walk
        if (j >= 0) { int o = w.order[j]; ge = ctr_ge(w.count[o], w.error[o], sc, se, 0); }
        unsigned m = __ballot_sync(CRICK_FULL, ge);
        int adv = m ? __ffs(m) - 1 : 32;  // predecessors that are NOT >= me
        t -= adv;
        if (adv < 32) break;
    }
    if (t == p) return;
    // shift order[t .. p) one step towards the tail, highest chunk first
    for (int hi = p; hi > t; hi -= 32) {
        int j = hi - 1 - lane;  // source position
        int o = (j >= t) ? w.order[j] : -1;
        __syncwarp();
        if (j >= t) { w.order[j + 1] = o; w.pos[o] = j + 1; }
        __syncwarp();
    }
    if (lane == 0) { w.order[t] = s; w.pos[s] = t; }
    __syncwarp();
}

// counter_new (:141-164)
__device__ __forceinline__ int ctr_new(SpsvW& w, long long item, long long count, long long error,
                                       long long when, int lane) {
    int s = (int)w.size;
    if (lane == 0) {
        w.item[s] = item; w.count[s] = count; w.error[s] = error; w.last[s] = when;
        w.order[s] = s; w.pos[s] = s;
    }
    w.size++;
    __syncwarp();
    sift(w, s, lane);
    map_put(w, item, s, lane);
    return s;
}

// swap (:186-210): relabel slot s, rebalance
__device__ __forceinline__ void ctr_replace(SpsvW& w, int s, long long item, long long count,
                                            long long error, long long when, int lane) {
    map_del(w, w.item[s], lane);
    if (lane == 0) { w.item[s] = item; w.count[s] = count; w.error[s] = error; w.last[s] = when; }
    __syncwarp();
    sift(w, w.pos[s], lane);
    map_put(w, item, s, lane);
    if (w.tomb > w.hcap / 4) map_rebuild(w, lane);
}

// spsv_add (:213-250)
__device__ __forceinline__ void spsv_add1(SpsvW& w, long long item, long long cnt, long long when,
                                          int lane) {
    int idx = map_find(w, item, lane);
    if (idx < 0) {
        if (w.size == w.capacity) {
            int s = w.order[w.size - 1];
            long long c = w.count[s];
            ctr_replace(w, s, item, c + 1, c, when, lane);  // ignores cnt (quirk P2)
        } else {
            ctr_new(w, item, cnt, 0, when, lane);
        }
    } else {
        int s = w.hval[idx];
        if (lane == 0) { w.count[s] += cnt; w.last[s] = when; }
        __syncwarp();
        sift(w, w.pos[s], lane);
    }
}

extern __shared__ __align__(16) unsigned char smem_raw[];

__device__ __forceinline__ SpsvW open_state(unsigned char* gbase, long cap, int hcap, int use_smem,
                                            int lane) {
    SpsvView g = spsv_carve(gbase, cap, hcap);
    SpsvHdr h = *g.hdr;
    SpsvView v = g;
    if (use_smem) {
        v = spsv_carve(smem_raw, cap, hcap);
        for (long i = lane; i < h.size; i += 32) {
            v.item[i] = g.item[i]; v.count[i] = g.count[i]; v.error[i] = g.error[i];
            v.last[i] = g.last[i];
            v.order[i] = g.order[i]; v.pos[i] = g.pos[i];
        }
        for (int i = lane; i < hcap; i += 32) { v.hkey[i] = g.hkey[i]; v.hval[i] = g.hval[i]; }
        __syncwarp();
    }
    SpsvW w;
    w.item = v.item; w.count = v.count; w.error = v.error; w.last = v.last; w.hkey = v.hkey;
    w.order = v.order; w.pos = v.pos; w.hval = v.hval;
    w.capacity = h.capacity; w.size = h.size; w.hcap = h.hcap; w.tomb = h.tomb;
    w.seen = h.seen; w.mixed = h.mixed;
    return w;
}

__device__ __forceinline__ void close_state(const SpsvW& w, unsigned char* gbase, long cap, int hcap,
                                            int use_smem, int lane) {
    SpsvView g = spsv_carve(gbase, cap, hcap);
    __syncwarp();
    if (use_smem) {
        for (long i = lane; i < w.size; i += 32) {
            g.item[i] = w.item[i]; g.count[i] = w.count[i]; g.error[i] = w.error[i];
            g.last[i] = w.last[i];
            g.order[i] = w.order[i]; g.pos[i] = w.pos[i];
        }
        for (int i = lane; i < hcap; i += 32) { g.hkey[i] = w.hkey[i]; g.hval[i] = w.hval[i]; }
    }
    if (lane == 0) {
        SpsvHdr h;
        h.capacity = w.capacity; h.size = w.size; h.hcap = w.hcap; h.tomb = w.tomb;
        h.seen = w.seen; h.mixed = w.mixed; h.pad = 0;
        *g.hdr = h;
    }
}

// spsv_int64_update_ndarray loop (:422-435), in order
__global__ void __launch_bounds__(32) k_spsv_update(unsigned char* state, long cap, int hcap,
                                                    int use_smem, const long long* __restrict__ item,
                                                    const long long* __restrict__ cnt,
                                                    long long cnt_scalar, long n,
                                                    const long long* __restrict__ when) {
    // when != nullptr: the elements are a sub-sequence of a stream that has already been
    // accounted for in `seen`; when[i] is the stream position of element i (heavy-hitter path)
    const int lane = threadIdx.x;
    SpsvW w = open_state(state, cap, hcap, use_smem, lane);
    for (long base = 0; base < n; base += 32) {
        long i = base + lane;
        long long it = i < n ? item[i] : 0;
        long long c = cnt ? (i < n ? cnt[i] : 0) : cnt_scalar;
        long long wh = when ? (i < n ? when[i] : 0) : w.seen + i;
        int m = (int)((n - base) < 32 ? (n - base) : 32);
        for (int k = 0; k < m; ++k) {
            long long itk = __shfl_sync(CRICK_FULL, it, k);
            long long ck = __shfl_sync(CRICK_FULL, c, k);
            long long whk = __shfl_sync(CRICK_FULL, wh, k);
            spsv_add1(w, itk, ck, whk, lane);
        }
    }
    if (!when) w.seen += n;
    close_state(w, state, cap, hcap, use_smem, lane);
}

// Parallel path, step 1 (Cafaro et al.: one summary per contiguous shard, then merges): warp g
// replays shard g = [g*n/G, (g+1)*n/G) in order into its own summary in HBM/L2.  Positions
// are global (main summary's `seen` + index), so that `last` survives the merges.
__global__ void __launch_bounds__(128) k_spsv_shards(unsigned char* shards, size_t stride, long cap,
                                                     int hcap, int nshards,
                                                     const unsigned char* main_state,
                                                     const long long* __restrict__ item,
                                                     const long long* __restrict__ cnt,
                                                     long long cnt_scalar, long n,
                                                     const long long* __restrict__ when) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= nshards) return;
    const long base_seen = reinterpret_cast<const SpsvHdr*>(main_state)->seen;
    unsigned char* state = shards + (size_t)g * stride;
    SpsvW w = open_state(state, cap, hcap, 0, lane);
    const long lo = (long)(((__int128)g * n) / nshards), hi = (long)(((__int128)(g + 1) * n) / nshards);
    for (long base = lo; base < hi; base += 32) {
        long i = base + lane;
        long long it = i < hi ? item[i] : 0;
        long long c = cnt ? (i < hi ? cnt[i] : 0) : cnt_scalar;
        long long wh = when ? (i < hi ? when[i] : 0) : base_seen + i;
        int m = (int)((hi - base) < 32 ? (hi - base) : 32);
        for (int k = 0; k < m; ++k) {
            long long itk = __shfl_sync(CRICK_FULL, it, k);
            long long ck = __shfl_sync(CRICK_FULL, c, k);
            long long whk = __shfl_sync(CRICK_FULL, wh, k);
            spsv_add1(w, itk, ck, whk, lane);
        }
    }
    w.seen = when ? base_seen : base_seen + hi;
    close_state(w, state, cap, hcap, 0, lane);
}

// set_state (:253-286); returns status in *status: 1 ok, -2 duplicate
__global__ void __launch_bounds__(32) k_spsv_set_state(unsigned char* state, long cap, int hcap,
                                                       int use_smem, const long long* __restrict__ c3,
                                                       long n, int* status) {
    const int lane = threadIdx.x;
    SpsvW w = open_state(state, cap, hcap, use_smem, lane);
    int rc = 1;
    for (long i = 0; i < n; ++i) {
        long long it = c3[3 * i], cn = c3[3 * i + 1], er = c3[3 * i + 2];
        if (map_find(w, it, lane) >= 0) { rc = -2; break; }
        ctr_new(w, it, cn, er, w.seen + i, lane);
    }
    w.seen += n;
    close_state(w, state, cap, hcap, use_smem, lane);
    if (lane == 0) *status = rc;
}

// spsv_merge (:289-364), T2 read-only from HBM.  Block b merges s2 + b*pair_stride into
// s1 + b*pair_stride (one pair for the user-level merge, a whole tree level for the parallel
// update).  mark_mixed: the two summaries count positions of different streams.
__global__ void __launch_bounds__(32) k_spsv_merge(unsigned char* s1, long cap1, int hcap1,
                                                   int use_smem, unsigned char* s2, long cap2,
                                                   int hcap2, size_t pair_stride, int mark_mixed) {
    const int lane = threadIdx.x;
    s1 += (size_t)blockIdx.x * pair_stride;
    s2 += (size_t)blockIdx.x * pair_stride;
    SpsvView g2 = spsv_carve(s2, cap2, hcap2);
    SpsvHdr h2 = *g2.hdr;
    if (h2.size == 0) {
        if (lane == 0 && h2.seen > reinterpret_cast<SpsvHdr*>(s1)->seen && !mark_mixed)
            reinterpret_cast<SpsvHdr*>(s1)->seen = h2.seen;
        return;
    }
    SpsvW w2;
    w2.item = g2.item; w2.count = g2.count; w2.error = g2.error; w2.last = g2.last; w2.hkey = g2.hkey;
    w2.order = g2.order; w2.pos = g2.pos; w2.hval = g2.hval;
    w2.capacity = h2.capacity; w2.size = h2.size; w2.hcap = h2.hcap; w2.tomb = h2.tomb;
    w2.seen = h2.seen; w2.mixed = h2.mixed;
    SpsvW w = open_state(s1, cap1, hcap1, use_smem, lane);
    long long m1 = w.size < w.capacity ? 0 : w.count[w.order[w.size - 1]];
    long long m2 = w2.size < w2.capacity ? 0 : w2.count[w2.order[w2.size - 1]];
    for (long s = 0; s < w.size; ++s) {  // slot order (:308)
        int idx = map_find(w2, w.item[s], lane);
        if (lane == 0) {
            if (idx >= 0) {
                int s2i = w2.hval[idx];
                w.count[s] += w2.count[s2i];
                w.error[s] += w2.error[s2i];
                if (w2.last[s2i] > w.last[s]) w.last[s] = w2.last[s2i];
            } else {
                w.count[s] += m2;
                w.error[s] += m2;
            }
        }
        __syncwarp();
        sift(w, w.pos[s], lane);
    }
    for (long p = 0; p < w2.size; ++p) {  // T2's list order (:330)
        int s2i = w2.order[p];
        long long it = w2.item[s2i], cn = w2.count[s2i], er = w2.error[s2i], la = w2.last[s2i];
        if (map_find(w, it, lane) >= 0) continue;
        if (w.size == w.capacity) {
            int tail = w.order[w.size - 1];
            if (ctr_ge(w.count[tail], w.error[tail], cn, er, m1)) break;
            ctr_replace(w, tail, it, cn + m1, er + m1, la, lane);
        } else {
            ctr_new(w, it, cn + m1, er + m1, la, lane);
        }
    }
    if (mark_mixed || w2.mixed) w.mixed = 1;
    else if (w2.seen > w.seen) w.seen = w2.seen;
    close_state(w, s1, cap1, hcap1, use_smem, lane);
}

// Parallel path, last step: a summary without errors is exact (no eviction happened anywhere),
// and the reference lists an exact summary by (count desc, position of the latest update asc)
// -- SURVEY.md 8c, verified on 50 random streams.  One CTA, bitonic sort in shared memory.
__global__ void __launch_bounds__(1024) k_spsv_reorder_exact(unsigned char* state, long cap, int hcap,
                                                             int P /* pow2 >= size */) {
    extern __shared__ __align__(16) unsigned char sm[];
    long long* kc = reinterpret_cast<long long*>(sm);
    long long* kl = kc + P;
    int* ks = reinterpret_cast<int*>(kl + P);
    __shared__ int s_bad;
    SpsvView g = spsv_carve(state, cap, hcap);
    const int n = (int)g.hdr->size;
    if (threadIdx.x == 0) s_bad = g.hdr->mixed;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (g.error[i] != 0) s_bad = 1;
    __syncthreads();
    if (s_bad || n > P) return;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        if (i < n) { kc[i] = g.count[i]; kl[i] = g.last[i]; ks[i] = i; }
        else { kc[i] = -0x7fffffffffffffffll - 1; kl[i] = 0x7fffffffffffffffll; ks[i] = -1; }
    }
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < P / 2; t += blockDim.x) {
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int q = i | j;
                bool up = (i & k) == 0;
                // "a before b": larger count first, then earlier last update
                bool q_first = kc[q] > kc[i] || (kc[q] == kc[i] && kl[q] < kl[i]);
                if (q_first == up) {
                    long long t1 = kc[i]; kc[i] = kc[q]; kc[q] = t1;
                    t1 = kl[i]; kl[i] = kl[q]; kl[q] = t1;
                    int t2 = ks[i]; ks[i] = ks[q]; ks[q] = t2;
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) { g.order[i] = ks[i]; g.pos[ks[i]] = i; }
}

// int64_counters walk (space_saving.pyx:368-386)
__global__ void k_spsv_counters(unsigned char* state, long cap, int hcap, long long* out, long k) {
    SpsvView g = spsv_carve(state, cap, hcap);
    long n = g.hdr->size < k ? g.hdr->size : k;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        int s = g.order[i];
        out[3 * i] = g.item[s]; out[3 * i + 1] = g.count[s]; out[3 * i + 2] = g.error[s];
    }
}

__global__ void k_spsv_init(unsigned char* state, long cap, int hcap, size_t stride) {
    SpsvView g = spsv_carve(state + (size_t)blockIdx.x * stride, cap, hcap);
    for (int i = threadIdx.x; i < hcap; i += blockDim.x) g.hval[i] = -1;
    if (threadIdx.x == 0) {
        SpsvHdr h; h.capacity = cap; h.size = 0; h.hcap = hcap; h.tomb = 0;
        h.seen = 0; h.mixed = 0; h.pad = 0;
        *g.hdr = h;
    }
}

// Multi-GPU tail: a summary travels as the fixed-size record
// [size, 0, (item, count, error) x capacity] (int64, list order, zero padded) ...
__global__ void k_spsv_export(unsigned char* state, long cap, int hcap, long long* out) {
    SpsvView g = spsv_carve(state, cap, hcap);
    const long n = g.hdr->size;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < 2 + 3 * cap;
         i += (long)gridDim.x * blockDim.x) {
        long long v = 0;
        if (i == 0) v = n;
        else if (i >= 2) {
            const long j = (i - 2) / 3;
            const int f = (int)((i - 2) % 3);
            if (j < n) { const int sl = g.order[j]; v = f == 0 ? g.item[sl] : f == 1 ? g.count[sl] : g.error[sl]; }
        }
        out[i] = v;
    }
}

// ... and is rebuilt on the receiving side as the state set_state (:253-286) would produce from
// those counters (slot i = list position i), hash map included, by one CTA in parallel.
__global__ void __launch_bounds__(1024) k_spsv_from_record(unsigned char* tmp, long cap, int hcap,
                                                           const long long* __restrict__ rec) {
    SpsvView g = spsv_carve(tmp, cap, hcap);
    const int tid = threadIdx.x;
    long n = rec[0];
    n = n < 0 ? 0 : (n > cap ? cap : n);
    for (int i = tid; i < hcap; i += 1024) g.hval[i] = -1;
    __syncthreads();
    for (long i = tid; i < n; i += 1024) {
        const long long key = rec[2 + 3 * i];
        g.item[i] = key; g.count[i] = rec[3 + 3 * i]; g.error[i] = rec[4 + 3 * i]; g.last[i] = i;
        g.order[i] = (int)i; g.pos[i] = (int)i;
        unsigned idx = hash64(key) & (unsigned)(hcap - 1);
        for (;;) {
            if (atomicCAS(&g.hval[idx], -1, (int)i) == -1) { g.hkey[idx] = key; break; }
            idx = (idx + 1) & (unsigned)(hcap - 1);
        }
    }
    if (tid == 0) {
        SpsvHdr h; h.capacity = cap; h.size = n; h.hcap = hcap; h.tomb = 0;
        h.seen = n; h.mixed = 1; h.pad = 0;
        *g.hdr = h;
    }
}


// inclusive scan of one int per thread over a 1024-thread CTA
__device__ __forceinline__ int hh_block_scan_fwd(int v, int* s_w32) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(CRICK_FULL, v, d);
        if (lane >= d) v += o;
    }
    __syncthreads();
    if (lane == 31) s_w32[wid] = v;
    __syncthreads();
    int t = s_w32[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(CRICK_FULL, t, d);
        if (lane >= d) t += o;
    }
    const int before = __shfl_sync(CRICK_FULL, t, wid > 0 ? wid - 1 : 0);
    return wid > 0 ? v + before : v;
}

// ---- k-way merge of gathered records in ONE pass (multi-GPU tail, crick_spsv_merge_records_bulk) --
// Cafaro et al.'s merge (spsv_merge, :289-364) for K summaries at once: an item monitored by
// summary r contributes (count_r, error_r), one that is not contributes r's bound (m_r, m_r) with
// m_r = r's smallest count if r is full, else 0; the `capacity` largest by (count desc, error asc)
// are kept.  Same guarantees as K - 1 pairwise merges (count - error <= true <= count; everything
// dropped is bounded by the smallest kept count) without their K - 1 dependent single-warp list
// walks: sort the <= K * capacity entries by item, reduce equal items, sort by count.  One CTA.
// Not the list order nested merge() calls would produce: crick_spsv_merge_records keeps that.
__device__ __forceinline__ void spsv_bitonic(long long* k1, long long* k2, int* id, int P) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < P / 2; t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int q = i | j;
                const bool up = (i & k) == 0;
                const bool q_first = k1[q] < k1[i] || (k1[q] == k1[i] && (k2[q] < k2[i] || (k2[q] == k2[i] && id[q] < id[i])));
                if (q_first == up) {
                    long long a = k1[i]; k1[i] = k1[q]; k1[q] = a;
                    a = k2[i]; k2[i] = k2[q]; k2[q] = a;
                    int b = id[i]; id[i] = id[q]; id[q] = b;
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(1024) k_spsv_merge_bulk(unsigned char* state, long cap, int hcap,
                                                          const long long* __restrict__ recs, int nrec,
                                                          long rec_len, int P, long long* scratch) {
    extern __shared__ __align__(16) unsigned char sm[];
    long long* k1 = reinterpret_cast<long long*>(sm);      // [P]
    long long* k2 = k1 + P;                                 // [P]
    int* id = reinterpret_cast<int*>(k2 + P);               // [P]
    __shared__ int s_off[65];
    __shared__ long long s_m[64];
    __shared__ long long s_M, s_seen;
    __shared__ int s_w32[32];
    __shared__ int s_D;
    const int tid = threadIdx.x;
    // scratch: agg_item[P] | agg_count[P] | agg_error[P]
    long long* a_item = scratch;
    long long* a_cnt = scratch + P;
    long long* a_err = scratch + 2 * (size_t)P;
    if (tid == 0) {
        int run = 0;
        long long M = 0, seen = 0;
        for (int r = 0; r < nrec; ++r) {
            const long long* rec = recs + (size_t)r * rec_len;
            long n = rec[0];
            n = n < 0 ? 0 : (n > cap ? cap : n);
            s_off[r] = run;
            run += (int)n;
            s_m[r] = n == cap ? rec[2 + 3 * (n - 1) + 1] : 0;     // count of the last (smallest) counter
            M += s_m[r];
            seen += n;
        }
        s_off[nrec] = run;
        s_M = M; s_seen = seen;
    }
    __syncthreads();
    const int T = s_off[nrec];
    for (int i = tid; i < P; i += blockDim.x) { k1[i] = 0x7fffffffffffffffll; k2[i] = 0x7fffffffffffffffll; id[i] = 0x7fffffff; }
    __syncthreads();
    for (int r = 0; r < nrec; ++r) {
        const long long* rec = recs + (size_t)r * rec_len;
        const int n = s_off[r + 1] - s_off[r];
        for (int i = tid; i < n; i += blockDim.x) {
            k1[s_off[r] + i] = rec[2 + 3 * i];               // item
            k2[s_off[r] + i] = r;                            // record (ties keep rank order)
            id[s_off[r] + i] = s_off[r] + i;                 // entry (its record and slot follow from s_off)
        }
    }
    __syncthreads();
    spsv_bitonic(k1, k2, id, P);
    // heads of equal-item runs -> aggregated (count, error); padding sorts last (id = INT_MAX)
    int run_total = 0;
    for (int base = 0; base < P; base += blockDim.x) {
        const int i = base + tid;
        const bool valid = i < P && id[i] != 0x7fffffff;
        const bool head = valid && (i == 0 || k1[i - 1] != k1[i] || id[i - 1] == 0x7fffffff);
        const int incl = hh_block_scan_fwd(head ? 1 : 0, s_w32);
        if (head) {
            long long cnt = 0, err = 0, bound = 0;
            for (int j = i; j < P && id[j] != 0x7fffffff && k1[j] == k1[i]; ++j) {
                const int r = (int)k2[j];
                const long long* rec = recs + (size_t)r * rec_len;
                const int slot = id[j] - s_off[r];
                cnt += rec[2 + 3 * slot + 1];
                err += rec[2 + 3 * slot + 2];
                bound += s_m[r];
            }
            const long long rest = s_M - bound;              // the summaries that do not monitor it
            const int at = run_total + incl - 1;
            a_item[at] = k1[i]; a_cnt[at] = cnt + rest; a_err[at] = err + rest;
        }
        if (tid == (int)blockDim.x - 1) s_D = run_total + incl;
        __syncthreads();
        run_total = s_D;
        __syncthreads();
    }
    const int D = run_total;
    for (int i = tid; i < P; i += blockDim.x) {
        k1[i] = i < D ? -a_cnt[i] : 0x7fffffffffffffffll;    // count descending
        k2[i] = i < D ? a_err[i] : 0x7fffffffffffffffll;     // error ascending
        id[i] = i < D ? i : 0x7fffffff;                      // item ascending (a_item is sorted)
    }
    __syncthreads();
    spsv_bitonic(k1, k2, id, P);
    const int K = D < (int)cap ? D : (int)cap;
    SpsvView g = spsv_carve(state, cap, hcap);
    for (int i = tid; i < hcap; i += blockDim.x) g.hval[i] = -1;
    __syncthreads();
    for (int i = tid; i < K; i += blockDim.x) {
        const int a = id[i];
        const long long key = a_item[a];
        g.item[i] = key; g.count[i] = a_cnt[a]; g.error[i] = a_err[a]; g.last[i] = i;
        g.order[i] = i; g.pos[i] = i;
        unsigned idx = hash64(key) & (unsigned)(hcap - 1);
        for (;;) {
            if (atomicCAS(&g.hval[idx], -1, i) == -1) { g.hkey[idx] = key; break; }
            idx = (idx + 1) & (unsigned)(hcap - 1);
        }
    }
    if (tid == 0) {
        SpsvHdr h; h.capacity = cap; h.size = K; h.hcap = hcap; h.tomb = 0;
        h.seen = s_seen; h.mixed = 1; h.pad = 0;
        *g.hdr = h;
    }
}
}  // namespace crick

#include "spsv_hh.cuh"

using namespace crick;

struct crick_spsv {
    unsigned char* d = nullptr;
    long cap = 0;
    int hcap = 0;
    int use_smem = 0;
    size_t smem = 0;
    cudaStream_t own = nullptr, last = nullptr;
    cudaEvent_t ev = nullptr;
    std::vector<long long> pi, pc;  // pending scalar adds
    long size_mirror = 0;
    bool mirror_ok = false;
    bool known_empty = false;        // nothing has been added since new(): seen == 0, size == 0
};

static cudaStream_t sp_stream(crick_spsv* s, void* stream) {
    cudaStream_t st = stream ? reinterpret_cast<cudaStream_t>(stream) : s->own;
    if (s->last && s->last != st) {
        cudaEventRecord(s->ev, s->last);
        cudaStreamWaitEvent(st, s->ev, 0);
    }
    s->last = st;
    return st;
}

constexpr long kSpsvParallelMin = 1 << 18;   // below: the bit-exact sequential replay
constexpr long kSpsvShardMin = 4096;         // elements per shard, at least
constexpr long kSpsvHHMin = 1 << 16;         // from here: count candidates instead of replaying
constexpr long kSpsvHHMaxCap = 1024;         // (4 x capacity candidates must fit the lookup image)

// Parallel update: shard summaries -> pairwise merge tree -> merge into the main summary ->
// list order of an exact summary.  Valid Space-Saving summary for any input (the merge is the
// reference's own, Cafaro et al.); counts are exact whenever #distinct <= capacity.
static cudaError_t sp_launch_update_parallel(crick_spsv* s, const long long* item,
                                             const long long* cnt, long long cs, long n,
                                             cudaStream_t st, const long long* when = nullptr) {
    const size_t stride = spsv_bytes(s->cap, s->hcap);
    long G = n / kSpsvShardMin;
    const long gmax = 16L * sm_count();
    if (G > gmax) G = gmax;
    while (G > 1 && (size_t)G * stride > ((size_t)8 << 30)) G >>= 1;   // scratch <= 8 GiB
    if (G < 1) G = 1;
    unsigned char* shards = nullptr;
    cudaError_t e = cudaMallocAsync(&shards, (size_t)G * stride, st);
    if (e != cudaSuccess) return e;
    count_launch();
    k_spsv_init<<<(unsigned)G, 256, 0, st>>>(shards, s->cap, s->hcap, stride);
    count_launch();
    k_spsv_shards<<<(unsigned)((G + 3) / 4), 128, 0, st>>>(shards, stride, s->cap, s->hcap, (int)G,
                                                          s->d, item, cnt, cs, n, when);
    for (long step = 1; step < G; step <<= 1) {
        // shard k <- merge(shard k, shard k + step) for k = 0, 2*step, 4*step, ...
        long pairs = (G - step + 2 * step - 1) / (2 * step);
        count_launch();
        k_spsv_merge<<<(unsigned)pairs, 32, 0, st>>>(shards, s->cap, s->hcap, 0, shards + step * stride,
                                                     s->cap, s->hcap, 2 * step * stride, 0);
    }
    e = cudaGetLastError();
    if (e == cudaSuccess && s->smem > 48 * 1024)
        e = cudaFuncSetAttribute(k_spsv_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem);
    if (e == cudaSuccess) {
        count_launch();
        k_spsv_merge<<<1, 32, s->smem, st>>>(s->d, s->cap, s->hcap, s->use_smem, shards, s->cap, s->hcap,
                                             0, 0);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && s->cap <= 8192) {
        int P = 1;
        while (P < s->cap) P <<= 1;
        size_t sm = (size_t)P * 20;
        if (sm > 48 * 1024)
            e = cudaFuncSetAttribute(k_spsv_reorder_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e == cudaSuccess) {
            count_launch();
            k_spsv_reorder_exact<<<1, 1024, sm, st>>>(s->d, s->cap, s->hcap, P);
            e = cudaGetLastError();
        }
    }
    cudaFreeAsync(shards, st);
    return e;
}

// fused moments (crick_spsv_update_stats): per-CTA StatsPart slots of the counting pass
struct HHFuse { StatsPart* parts; int item_is_f64; int used; };
static cudaError_t sp_launch_update_hh(crick_spsv* s, const long long* item, const long long* cnt,
                                       long n, cudaStream_t st, bool* done, bool exact_only,
                                       HHFuse* fuse = nullptr);
static bool hh2_enabled() {
    static const bool on = [] { const char* e = getenv("CRICK_HH_V"); return !(e && e[0] == '1'); }();
    return on;
}
constexpr int kHH2Buckets = 24000;   // per-CTA miss buckets: with L = 8192 the CTA uses 64 + 32 + 32 + 94 KB

static cudaError_t sp_launch_update(crick_spsv* s, const long long* item, const long long* cnt,
                                    long long cs, long n, cudaStream_t st, int mode = 0,
                                    const long long* when = nullptr) {
    // mode: 0 auto, 1 sequential (reference replay), 2 parallel (shards), 3 counting (heavy hitters)
    // below kSpsvHHMin the reference's bits are kept for every input: a fresh summary still tries
    // the counting path, but takes its result only when no eviction can have happened
    const bool small_exact = mode == 0 && n < kSpsvHHMin && s->known_empty;
    if (!when && (cnt || cs == 1) && s->cap <= kSpsvHHMaxCap &&
        (mode == 3 || (mode == 0 && n >= kSpsvHHMin) || small_exact) && n >= 4 * 1024) {
        bool done = false;
        cudaError_t e = sp_launch_update_hh(s, item, cnt, n, st, &done, small_exact);
        if (e != cudaSuccess || done) return e;
        // not applicable to this batch (see k_hh_finish): replay it
    }
    if (mode == 2 || mode == 3 || (mode == 0 && n >= kSpsvParallelMin))
        return sp_launch_update_parallel(s, item, cnt, cs, n, st, when);
    cudaError_t e = cudaSuccess;
    if (s->smem > 48 * 1024)
        e = cudaFuncSetAttribute(k_spsv_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem);
    if (e != cudaSuccess) return e;
    count_launch();
    k_spsv_update<<<1, 32, s->smem, st>>>(s->d, s->cap, s->hcap, s->use_smem, item, cnt, cs, n, when);
    return cudaGetLastError();
}

// Counting path (spsv_hh.cuh).  *done = false: the batch does not suit it, nothing was changed.
static cudaError_t sp_launch_update_hh(crick_spsv* s, const long long* item, const long long* cnt,
                                       long n, cudaStream_t st, bool* done, bool exact_only,
                                       HHFuse* fuse) {
    *done = false;
    // unit counts: the round-2 counting pass (k_hh_count2: shared-memory buckets, one CTA of 32
    // warps per SM); per-element counts keep the round-1 kernel with its 64-bit counters
    const bool v2 = cnt == nullptr && hh2_enabled();
    const int NB = kHH2Buckets;
    const int nw2 = fuse ? 20 : 24;            // warps per CTA of k_hh_count2 (registers)
    const size_t csz = cnt ? 8 : 4;            // counts of the counting pass: u64 with weights
    int C = 256;
    while (C < 4 * s->cap) C <<= 1;
    if (C > kHHCandMax) C = kHHCandMax;
    const int L = 2 * C;
    const long S = n / 4 < kHHSampleMax ? n / 4 : kHHSampleMax;
    // misses are replayed (exactness when #distinct <= capacity) only while they are few: items
    // the sample did not see are rare, a stream with more distinct items than counters misses ~half
    long miss_cap = n / 128 > 4096 ? n / 128 : 4096;
    if (miss_cap > kHHMissCap) miss_cap = kHHMissCap;
    const size_t smem_cnt = v2 ? (size_t)L * 16 + (size_t)NB * 4
                               : (size_t)L * 10 + (size_t)C * (4 + csz);
    const size_t smem_col = (size_t)L * 8;                           // v2, collecting pass: keys only
    int G = v2 ? sm_count() : sm_count() * (2 * smem_cnt + 2048 <= 227 * 1024 ? 2 : 1);
    const long W = (long)G * (v2 ? nw2 : kHHWarps);
    const size_t stride = spsv_bytes(s->cap, s->hcap);
    // one allocation: sample table | buckets | candidates | lookup image | rows | miss counts | status | tmp
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_tabkey = take((size_t)kHHTabSlots * 8), o_tabcnt = take((size_t)kHHTabSlots * 4);
    const size_t o_buckets = take(v2 ? (size_t)G * NB * 4 : (size_t)kHHBuckets * csz);
    const size_t o_cand = take((size_t)C * 8), o_lutk = take((size_t)L * 8), o_lutv = take((size_t)L * 2);
    const size_t o_rc = take((size_t)G * (v2 ? L : C) * csz), o_rl = take((size_t)G * (v2 ? L : C) * 4);
    const size_t o_cslot = take((size_t)C * 4);
    const size_t o_wm = take((size_t)W * 8), o_wo = take((size_t)W * 8);
    const size_t o_cc = take((size_t)C * 8), o_cl = take((size_t)C * 8);
    const size_t o_hist = take(1024 * 4), o_sel = take(sizeof(HHSel));
    const size_t o_cg = take((size_t)kHHNChunk * 4), o_ce = take((size_t)kHHNChunk * 4);
    const size_t o_status = take(sizeof(HHStatus)), o_tmp = take(stride);
    unsigned char* ws = nullptr;
    cudaError_t e = cudaMallocAsync(&ws, off, st);
    if (e != cudaSuccess) return e;
    long long* tabkey = reinterpret_cast<long long*>(ws + o_tabkey);
    unsigned* tabcnt = reinterpret_cast<unsigned*>(ws + o_tabcnt);
    void* buckets = ws + o_buckets;
    long long* cand = reinterpret_cast<long long*>(ws + o_cand);
    long long* lutk = reinterpret_cast<long long*>(ws + o_lutk);
    unsigned short* lutv = reinterpret_cast<unsigned short*>(ws + o_lutv);
    void* rc = ws + o_rc;
    unsigned* rl = reinterpret_cast<unsigned*>(ws + o_rl);
    int* cslot = reinterpret_cast<int*>(ws + o_cslot);
    unsigned long long* wm = reinterpret_cast<unsigned long long*>(ws + o_wm);
    unsigned long long* wo = reinterpret_cast<unsigned long long*>(ws + o_wo);
    long long* ccnt = reinterpret_cast<long long*>(ws + o_cc);
    long long* clast = reinterpret_cast<long long*>(ws + o_cl);
    unsigned* hist = reinterpret_cast<unsigned*>(ws + o_hist);
    HHSel* sel = reinterpret_cast<HHSel*>(ws + o_sel);
    int* chunk_gt = reinterpret_cast<int*>(ws + o_cg);
    int* chunk_eq = reinterpret_cast<int*>(ws + o_ce);
    HHStatus* dstatus = reinterpret_cast<HHStatus*>(ws + o_status);
    unsigned char* tmp = ws + o_tmp;
    long long* miss_item = nullptr;
    HHStatus hs = {};
    long base_seen = 0;
    do {
        // the stream position of this batch's first element (0 for a summary nothing was added to)
        const bool fresh = s->known_empty;
        if (!fresh) {
            e = cudaMemcpyAsync(&base_seen, s->d + offsetof(SpsvHdr, seen), sizeof(long), cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) break;
        }
        e = cudaMemsetAsync(tabcnt, 0, (size_t)kHHTabSlots * 4, st);
        if (e == cudaSuccess && !v2) e = cudaMemsetAsync(buckets, 0, (size_t)kHHBuckets * csz, st);
        if (e != cudaSuccess) break;
        e = cudaMemsetAsync(dstatus, 0, sizeof(HHStatus), st);
        if (e != cudaSuccess) break;
        count_launch();
        k_hh_fill<<<1024, 256, 0, st>>>(tabkey, kHHTabSlots, kHHEmpty);
        count_launch();
        k_hh_sample<<<(unsigned)((S + kHHSampleThreads - 1) / kHHSampleThreads), kHHSampleThreads, 0, st>>>(item, n, S, tabkey, tabcnt);
        e = cudaMemsetAsync(hist, 0, 1024 * 4, st);
        if (e != cudaSuccess) break;
        count_launch();
        k_hh_hist<<<128, 1024, 0, st>>>(tabcnt, hist);
        count_launch();
        k_hh_thresh<<<1, 1024, 0, st>>>(hist, C, sel);
        count_launch();
        k_hh_chunks<<<kHHNChunk, 1024, 0, st>>>(tabcnt, sel, chunk_gt, chunk_eq);
        count_launch();
        k_hh_place<<<kHHNChunk, 1024, 0, st>>>(tabkey, tabcnt, sel, chunk_gt, chunk_eq, cand);
        count_launch();
        if (v2) k_hh_lut2<<<1, 1024, 0, st>>>(cand, sel, chunk_eq, lutk, cslot, L, dstatus);
        else k_hh_lut<<<1, 1024, 0, st>>>(cand, sel, chunk_eq, lutk, lutv, L, dstatus);
        if (v2) {
            e = fuse ? cudaFuncSetAttribute(k_hh_count2<0, true, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cnt)
                     : cudaFuncSetAttribute(k_hh_count2<0, false, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cnt);
        } else {
            e = cnt ? cudaFuncSetAttribute(k_hh_count<unsigned long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cnt)
                    : cudaFuncSetAttribute(k_hh_count<unsigned>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cnt);
        }
        if (e != cudaSuccess) break;
        if (!fresh) e = cudaStreamSynchronize(st);           // base_seen has arrived
        if (e != cudaSuccess) break;
        count_launch();
        if (v2 && fuse) {
            k_hh_count2<0, true, 20><<<G, 20 * 32, smem_cnt, st>>>(
                item, n, lutk, L, NB, sel, static_cast<unsigned*>(rc), rl,
                static_cast<unsigned*>(buckets), wm, nullptr, nullptr, nullptr, base_seen, fuse->parts,
                fuse->item_is_f64);
            fuse->used = G;
        } else if (v2)
            k_hh_count2<0, false, 24><<<G, 24 * 32, smem_cnt, st>>>(
                item, n, lutk, L, NB, sel, static_cast<unsigned*>(rc), rl,
                static_cast<unsigned*>(buckets), wm, nullptr, nullptr, nullptr, base_seen, nullptr, 0);
        else if (cnt)
            k_hh_count<unsigned long long><<<G, kHHWarps * 32, smem_cnt, st>>>(
                item, cnt, n, lutk, lutv, L, C, static_cast<unsigned long long*>(rc), rl,
                static_cast<unsigned long long*>(buckets), wm, 0, nullptr, nullptr, nullptr, nullptr, base_seen);
        else
            k_hh_count<unsigned><<<G, kHHWarps * 32, smem_cnt, st>>>(
                item, nullptr, n, lutk, lutv, L, C, static_cast<unsigned*>(rc), rl,
                static_cast<unsigned*>(buckets), wm, 0, nullptr, nullptr, nullptr, nullptr, base_seen);
        const size_t smem_fin = (size_t)C * 20;
        e = cudaFuncSetAttribute(k_hh_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fin);
        if (e != cudaSuccess) break;
        count_launch();
        if (v2)
            k_hh_reduce2<<<148, 256, (size_t)G * sizeof(long), st>>>(
                static_cast<const unsigned*>(rc), rl, static_cast<const unsigned*>(buckets), G, nw2, C, L, NB, cslot,
                wm, n, base_seen, sel, ccnt, clast, dstatus);
        else if (cnt)
            k_hh_reduce<unsigned long long><<<64, 256, (size_t)G * sizeof(long), st>>>(
                static_cast<const unsigned long long*>(rc), rl, G, C,
                static_cast<const unsigned long long*>(buckets), wm, n, base_seen, ccnt, clast, dstatus);
        else
            k_hh_reduce<unsigned><<<64, 256, (size_t)G * sizeof(long), st>>>(
                static_cast<const unsigned*>(rc), rl, G, C, static_cast<const unsigned*>(buckets), wm, n,
                base_seen, ccnt, clast, dstatus);
        count_launch();
        k_hh_finish<<<1, 1024, smem_fin, st>>>(ccnt, clast, C, cand, n, base_seen, tmp, s->cap, s->hcap,
                                               miss_cap, dstatus);
        e = cudaMemcpyAsync(&hs, dstatus, sizeof(HHStatus), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) break;
        if (hs.fallback || hs.lutfail) break;    // *done stays false: nothing was changed
        // exact_only (fresh summary, small batch): every distinct item must get its own counter
        if (exact_only && (long)hs.npos + hs.others > s->cap) break;
        // fold the batch's summary into the object (spsv_merge).  Merging into an empty summary
        // of the same capacity reproduces the other one slot for slot (:330-350 with m1 = 0), so
        // a fresh object simply takes the image
        if (fresh) {
            e = cudaMemcpyAsync(s->d, tmp, stride, cudaMemcpyDeviceToDevice, st);
        } else {
            if (s->smem > 48 * 1024)
                e = cudaFuncSetAttribute(k_spsv_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem);
            if (e != cudaSuccess) break;
            count_launch();
            k_spsv_merge<<<1, 32, s->smem, st>>>(s->d, s->cap, s->hcap, s->use_smem, tmp, s->cap, s->hcap, 0, 0);
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) break;
        *done = true;
        if (hs.others > 0 && hs.others <= miss_cap) {
            // few elements outside the candidates: collect them in stream order and replay them
            e = cudaMallocAsync(&miss_item, (size_t)hs.others * 24, st);
            if (e != cudaSuccess) break;
            long long* miss_pos = miss_item + hs.others;
            long long* miss_cnt = cnt ? miss_pos + hs.others : nullptr;
            count_launch();
            k_hh_scan<<<1, 1024, 0, st>>>(wm, (int)W, wo);
            count_launch();
            if (v2) {
                e = nw2 == 20 ? cudaFuncSetAttribute(k_hh_count2<1, false, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_col)
                              : cudaFuncSetAttribute(k_hh_count2<1, false, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_col);
                if (e != cudaSuccess) break;
                if (nw2 == 20)
                    k_hh_count2<1, false, 20><<<G, 20 * 32, smem_col, st>>>(
                        item, n, lutk, L, NB, sel, nullptr, nullptr, nullptr, nullptr, wo, miss_item,
                        miss_pos, base_seen, nullptr, 0);
                else
                    k_hh_count2<1, false, 24><<<G, 24 * 32, smem_col, st>>>(
                        item, n, lutk, L, NB, sel, nullptr, nullptr, nullptr, nullptr, wo, miss_item,
                        miss_pos, base_seen, nullptr, 0);
            } else if (cnt)
                k_hh_count<unsigned long long><<<G, kHHWarps * 32, smem_cnt, st>>>(
                    item, cnt, n, lutk, lutv, L, C, static_cast<unsigned long long*>(rc), rl,
                    static_cast<unsigned long long*>(buckets), wm, 1, wo, miss_item, miss_pos, miss_cnt, base_seen);
            else
                k_hh_count<unsigned><<<G, kHHWarps * 32, smem_cnt, st>>>(
                    item, nullptr, n, lutk, lutv, L, C, static_cast<unsigned*>(rc), rl,
                    static_cast<unsigned*>(buckets), wm, 1, wo, miss_item, miss_pos, nullptr, base_seen);
            e = cudaGetLastError();
            if (e != cudaSuccess) break;
            e = sp_launch_update(s, miss_item, miss_cnt, 1, (long)hs.others, st,
                                 hs.others >= kSpsvParallelMin ? 2 : 1, miss_pos);
            if (e != cudaSuccess) break;
        }
        if (s->cap <= 8192) {
            int P = 1;
            while (P < s->cap) P <<= 1;
            size_t sm = (size_t)P * 20;
            if (sm > 48 * 1024)
                e = cudaFuncSetAttribute(k_spsv_reorder_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            if (e != cudaSuccess) break;
            count_launch();
            k_spsv_reorder_exact<<<1, 1024, sm, st>>>(s->d, s->cap, s->hcap, P);
            e = cudaGetLastError();
        }
    } while (0);
    if (miss_item) cudaFreeAsync(miss_item, st);
    cudaFreeAsync(ws, st);
    return e;
}

static int sp_drain(crick_spsv* s, cudaStream_t st) {
    size_t n = s->pi.size();
    if (!n) return 0;
    long long* d = nullptr;
    cudaError_t e = cudaMallocAsync(&d, n * 16, st);
    if (e != cudaSuccess) return fail("cudaMallocAsync(spsv pending)", e);
    e = cudaMemcpyAsync(d, s->pi.data(), n * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + n, s->pc.data(), n * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = sp_launch_update(s, d, d + n, 1, (long)n, st);
    cudaFreeAsync(d, st);
    s->pi.clear(); s->pc.clear();
    s->mirror_ok = false;
    s->known_empty = false;
    return e == cudaSuccess ? 0 : fail("spsv drain", e);
}

extern "C" {

crick_spsv_t* crick_spsv_new(int64_t capacity) {
    if (!sm_count()) { fail_msg("no CUDA device: libcrick_cuda has no CPU fallback"); return nullptr; }
    if (capacity <= 0 || capacity > (1 << 26)) { fail_msg("capacity out of range"); return nullptr; }
    crick_spsv* s = new (std::nothrow) crick_spsv();
    if (!s) return nullptr;
    s->cap = capacity;
    int h = 64;
    while (h < 4 * capacity) h <<= 1;
    s->hcap = h;
    size_t bytes = spsv_bytes(s->cap, s->hcap);
    s->use_smem = bytes <= 200 * 1024;
    s->smem = s->use_smem ? bytes : 0;
    cudaError_t e = cudaMalloc(&s->d, bytes);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->own, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev, cudaEventDisableTiming);
    if (e == cudaSuccess) {
        count_launch();
        k_spsv_init<<<1, 256, 0, s->own>>>(s->d, s->cap, s->hcap, 0);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) { fail("spsv_new", e); crick_spsv_free(s); return nullptr; }
    s->last = s->own;
    s->size_mirror = 0;
    s->mirror_ok = true;
    s->known_empty = true;
    return s;
}

// back to the state spsv_new leaves, without releasing anything (asynchronous on `stream`): lets
// a multi-GPU reduction reuse one output summary per step
int crick_spsv_reset(crick_spsv_t* s, void* stream) {
    cudaStream_t st = sp_stream(s, stream);
    s->pi.clear(); s->pc.clear();
    count_launch();
    k_spsv_init<<<1, 256, 0, st>>>(s->d, s->cap, s->hcap, 0);
    cudaError_t e = cudaGetLastError();
    s->size_mirror = 0;
    s->mirror_ok = e == cudaSuccess;
    s->known_empty = e == cudaSuccess;
    return e == cudaSuccess ? 0 : fail("spsv_reset", e);
}

void crick_spsv_free(crick_spsv_t* s) {
    if (!s) return;
    // (only streams this handle owns are touched: `last` may be a caller's stream that no longer
    // exists; cudaFree below waits for the whole device before the memory goes away)
    if (s->own) { cudaStreamSynchronize(s->own); cudaStreamDestroy(s->own); }
    if (s->ev) cudaEventDestroy(s->ev);
    if (s->d) cudaFree(s->d);
    delete s;
}

int crick_spsv_add(crick_spsv_t* s, int64_t item, int64_t count) {
    s->pi.push_back(item);
    s->pc.push_back(count);
    s->mirror_ok = false;
    if (s->pi.size() >= (1u << 16)) return sp_drain(s, sp_stream(s, nullptr));
    return 0;
}

int crick_spsv_update(crick_spsv_t* s, const int64_t* item, const int64_t* count, int64_t count_scalar,
                      int64_t n, int item_on_device, int count_on_device, int mode, void* stream) {
    if (n <= 0) return 0;
    cudaStream_t st = sp_stream(s, stream);
    if (sp_drain(s, st)) return -1;
    s->mirror_ok = false;
    const long long* di = reinterpret_cast<const long long*>(item);
    const long long* dc = reinterpret_cast<const long long*>(count);
    unsigned char* stage = nullptr;
    cudaError_t e = cudaSuccess;
    size_t need = (item_on_device ? 0 : (size_t)n * 8) + ((count && !count_on_device) ? (size_t)n * 8 : 0);
    if (need) {
        e = cudaMallocAsync(&stage, need, st);
        if (e != cudaSuccess) return fail("cudaMallocAsync(spsv stage)", e);
        unsigned char* p = stage;
        if (!item_on_device) {
            e = staged_copy_h2d(p, item, n, st);
            di = reinterpret_cast<const long long*>(p); p += (size_t)n * 8;
        }
        if (e == cudaSuccess && count && !count_on_device) {
            e = staged_copy_h2d(p, count, n, st);
            dc = reinterpret_cast<const long long*>(p);
        }
    }
    if (e == cudaSuccess) e = sp_launch_update(s, di, dc, count_scalar, n, st, mode);
    s->known_empty = false;
    if (stage) {
        cudaFreeAsync(stage, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    return e == cudaSuccess ? 0 : fail("spsv_update", e);
}

// SpaceSaving.update(item) and SummaryStats.update(item) in ONE pass over the items (BASELINE
// config 5).  Unit counts.  The moments are fused into the counting pass of the heavy-hitter path;
// where that path does not apply (short batches, capacity > 1024, a batch it hands back) the two
// sketches are updated one after the other -- same results, two passes.
int crick_spsv_update_stats(crick_spsv_t* s, crick_stats_t* stats, const int64_t* item, int64_t n,
                            int item_on_device, int item_is_f64, void* stream) {
    if (!stats) return fail_msg("spsv_update_stats: stats must not be NULL");
    if (n <= 0) return 0;
    cudaStream_t st = sp_stream(s, stream);
    if (sp_drain(s, st)) return -1;
    s->mirror_ok = false;
    const long long* di = reinterpret_cast<const long long*>(item);
    unsigned char* stage = nullptr;
    cudaError_t e = cudaSuccess;
    if (!item_on_device) {
        e = cudaMallocAsync(&stage, (size_t)n * 8, st);
        if (e != cudaSuccess) return fail("cudaMallocAsync(spsv stage)", e);
        e = staged_copy_h2d(stage, item, n, st);
        di = reinterpret_cast<const long long*>(stage);
    }
    bool done = false;
    if (e == cudaSuccess && s->cap <= kSpsvHHMaxCap && n >= kSpsvHHMin && hh2_enabled()) {
        HHFuse fuse{nullptr, item_is_f64 ? 1 : 0, 0};
        int nparts = 0;
        if (stats_fused_begin(stats, st, &fuse.parts, &nparts)) { if (stage) cudaFreeAsync(stage, st); return -1; }
        if (sm_count() <= nparts) {
            e = sp_launch_update_hh(s, di, nullptr, n, st, &done, false, &fuse);
            if (e == cudaSuccess && done) e = stats_fused_finish(stats, fuse.used, st);
        }
        crick_stats_release_stream(stats, st);
    }
    int rc = 0;
    if (e == cudaSuccess && !done) {
        // two passes (the device copy is reused, so host input crosses the bus once)
        rc = crick_stats_update(stats, di, item_is_f64 ? 0 : 1, nullptr, 1, n, 1, 0, st);
        if (rc == 0) rc = crick_stats_release_stream(stats, st);     // `st` is this summary's stream, not theirs
        if (rc == 0) e = sp_launch_update(s, di, nullptr, 1, n, st, 0);
    }
    s->known_empty = false;
    if (stage) {
        cudaFreeAsync(stage, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    if (rc) return rc;
    return e == cudaSuccess ? 0 : fail("spsv_update_stats", e);
}

int crick_spsv_merge(crick_spsv_t* s, crick_spsv_t* const* others, int n_others) {
    cudaStream_t st = sp_stream(s, nullptr);
    if (sp_drain(s, st)) return -1;
    s->mirror_ok = false; s->known_empty = false;
    for (int i = 0; i < n_others; ++i) {
        crick_spsv* o = others[i];
        if (o == s) return fail_msg("spsv_merge: cannot merge a summary into itself");
        cudaStream_t ost = sp_stream(o, nullptr);
        if (sp_drain(o, ost)) return -1;
        cudaError_t e = cudaEventRecord(o->ev, ost);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, o->ev, 0);
        if (e == cudaSuccess && s->smem > 48 * 1024)
            e = cudaFuncSetAttribute(k_spsv_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem);
        if (e == cudaSuccess) {
            count_launch();
            k_spsv_merge<<<1, 32, s->smem, st>>>(s->d, s->cap, s->hcap, s->use_smem, o->d, o->cap, o->hcap,
                                                 0, 1);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) {
            e = cudaEventRecord(s->ev, st);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(ost, s->ev, 0);
        }
        if (e != cudaSuccess) return fail("spsv_merge", e);
    }
    return 0;
}

int64_t crick_spsv_record_int64s(crick_spsv_t* s) { return 2 + 3 * (int64_t)s->cap; }

int crick_spsv_export_record(crick_spsv_t* s, int64_t* dev_out, void* stream) {
    cudaStream_t st = sp_stream(s, stream);
    if (sp_drain(s, st)) return -1;
    count_launch();
    k_spsv_export<<<8, 256, 0, st>>>(s->d, s->cap, s->hcap, reinterpret_cast<long long*>(dev_out));
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail("spsv_export_record", e);
}

int crick_spsv_merge_records(crick_spsv_t* s, const int64_t* dev_recs, int n_records, void* stream) {
    if (n_records <= 0) return 0;
    cudaStream_t st = sp_stream(s, stream);
    if (sp_drain(s, st)) return -1;
    s->mirror_ok = false; s->known_empty = false;
    const size_t stride = spsv_bytes(s->cap, s->hcap);
    unsigned char* tmp = nullptr;
    cudaError_t e = cudaMallocAsync(&tmp, stride, st);
    if (e != cudaSuccess) return fail("cudaMallocAsync(spsv record)", e);
    if (s->smem > 48 * 1024)
        e = cudaFuncSetAttribute(k_spsv_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem);
    const int64_t nd = 2 + 3 * (int64_t)s->cap;
    for (int r = 0; r < n_records && e == cudaSuccess; ++r) {
        count_launch();
        k_spsv_from_record<<<1, 1024, 0, st>>>(tmp, s->cap, s->hcap,
                                               reinterpret_cast<const long long*>(dev_recs + (size_t)r * nd));
        count_launch();
        k_spsv_merge<<<1, 32, s->smem, st>>>(s->d, s->cap, s->hcap, s->use_smem, tmp, s->cap, s->hcap, 0, 1);
        e = cudaGetLastError();
    }
    cudaFreeAsync(tmp, st);
    return e == cudaSuccess ? 0 : fail("spsv_merge_records", e);
}

int crick_spsv_merge_records_bulk(crick_spsv_t* s, const int64_t* dev_recs, int n_records, void* stream) {
    if (n_records <= 0) return 0;
    const int64_t nd = 2 + 3 * (int64_t)s->cap;
    const bool with_self = !s->known_empty || !s->pi.empty();
    const int K = n_records + (with_self ? 1 : 0);
    int P = 1;
    while (P < K * s->cap) P <<= 1;
    const size_t smem = (size_t)P * 20;
    if (K > 64 || smem > 200 * 1024)                       // too many entries for one CTA's sort
        return crick_spsv_merge_records(s, dev_recs, n_records, stream);
    cudaStream_t st = sp_stream(s, stream);
    if (sp_drain(s, st)) return -1;
    s->mirror_ok = false; s->known_empty = false;
    long long* ws = nullptr;                               // [own record + copies] | scratch 3 P
    const size_t recs_bytes = with_self ? (size_t)K * nd * 8 : 0;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&ws), recs_bytes + (size_t)3 * P * 8, st);
    if (e != cudaSuccess) return fail("cudaMallocAsync(spsv bulk merge)", e);
    const long long* recs = reinterpret_cast<const long long*>(dev_recs);
    if (with_self) {                                       // the handle's own summary goes first
        count_launch();
        k_spsv_export<<<8, 256, 0, st>>>(s->d, s->cap, s->hcap, ws);
        e = cudaMemcpyAsync(ws + nd, dev_recs, (size_t)n_records * nd * 8, cudaMemcpyDeviceToDevice, st);
        recs = ws;
    }
    if (e == cudaSuccess && smem > 48 * 1024)
        e = cudaFuncSetAttribute(k_spsv_merge_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        count_launch();
        k_spsv_merge_bulk<<<1, 1024, smem, st>>>(s->d, s->cap, s->hcap, recs, K, (long)nd, P,
                                                 ws + recs_bytes / 8);
        e = cudaGetLastError();
    }
    cudaFreeAsync(ws, st);
    return e == cudaSuccess ? 0 : fail("spsv_merge_records_bulk", e);
}

int crick_spsv_set_state(crick_spsv_t* s, const int64_t* counters3, int64_t n) {
    if (n > s->cap) return fail_msg("deserialization failed, size > capacity");
    if (n <= 0) return 1;
    cudaStream_t st = sp_stream(s, nullptr);
    if (sp_drain(s, st)) return -1;
    s->mirror_ok = false; s->known_empty = false;
    long long* d = nullptr;
    cudaError_t e = cudaMallocAsync(&d, (size_t)n * 24 + 8, st);
    if (e != cudaSuccess) return fail("cudaMallocAsync(spsv set_state)", e);
    int* dstatus = reinterpret_cast<int*>(d + 3 * n);
    int status = 0;
    e = cudaMemcpyAsync(d, counters3, (size_t)n * 24, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && s->smem > 48 * 1024)
        e = cudaFuncSetAttribute(k_spsv_set_state, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem);
    if (e == cudaSuccess) {
        count_launch();
        k_spsv_set_state<<<1, 32, s->smem, st>>>(s->d, s->cap, s->hcap, s->use_smem, d, n, dstatus);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&status, dstatus, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaFreeAsync(d, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail("spsv_set_state", e);
    if (status == -2) return fail_msg("deserialization failed, duplicate items found");
    return 1;
}

int64_t crick_spsv_size(crick_spsv_t* s) {
    cudaStream_t st = sp_stream(s, nullptr);
    if (sp_drain(s, st)) return -1;
    if (s->mirror_ok) return s->size_mirror;
    SpsvHdr h;
    cudaError_t e = cudaMemcpyAsync(&h, s->d, sizeof(h), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail("spsv_size", e);
    s->size_mirror = h.size;
    s->mirror_ok = true;
    return h.size;
}

int64_t crick_spsv_capacity(crick_spsv_t* s) { return s->cap; }

int64_t crick_spsv_counters(crick_spsv_t* s, int64_t* out3, int64_t k) {
    int64_t size = crick_spsv_size(s);
    if (size < 0) return -1;
    int64_t n = size < k ? size : k;
    if (n <= 0) return 0;
    cudaStream_t st = sp_stream(s, nullptr);
    long long* d = nullptr;
    cudaError_t e = cudaMallocAsync(&d, (size_t)n * 24, st);
    if (e != cudaSuccess) return fail("cudaMallocAsync(spsv counters)", e);
    count_launch();
    k_spsv_counters<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s->d, s->cap, s->hcap, d, n);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out3, d, (size_t)n * 24, cudaMemcpyDeviceToHost, st);
    cudaFreeAsync(d, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    return e == cudaSuccess ? n : fail("spsv_counters", e);
}

}  // extern "C"
