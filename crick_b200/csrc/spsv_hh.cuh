// crick_b200/csrc/spsv_hh.cuh -- included by spsv.cu (namespace crick).
//
// SpaceSaving update for LONG unit-count streams (>= 4 Mi elements): counting instead of
// replaying.  The replay (k_spsv_update / k_spsv_shards) costs microseconds per element because
// every element is a dependent walk through a hash map and a sorted list; a Stream-Summary,
// however, only has to satisfy  count - error <= true count <= count  for the items it keeps and
// true count <= smallest counter  for the ones it does not (Metwally et al.), and Cafaro et
// al.'s merge (spsv_merge, space_saving_stubs.c.in:289-364) combines any two such summaries.
// So the batch is summarised by EXACT counting of candidates:
//   1. k_hh_sample    1 Mi elements at strided+hashed positions into a global hash table;
//   2. k_hh_select    the C = 4 x capacity most frequent sampled items become candidates (one
//                     CTA: histogram of the sample counts, threshold, compaction in slot order
//                     => deterministic) + a linear-probing lookup image for them;
//   3. k_hh_count     ONE streaming pass over the batch (8 B / element, HBM bound): lookup in
//                     shared memory, hits -> per-CTA count / latest-position arrays (shared
//                     atomics), misses -> 1 Mi bucket counters in L2 + a per-warp miss count;
//   4. k_hh_finish    exact counts of the candidates, sorted by (count desc, latest position
//                     asc); the top `capacity` become a summary with error 0.  Every item
//                     that is not kept is bounded by max(next candidate, fullest miss bucket)
//                     =: E; if E exceeds the smallest kept count all counts and errors are
//                     raised by the difference (still a valid summary, tighter than a replay's);
//   5. k_spsv_merge   the reference's merge folds it into the object's summary.
// When only a few elements missed the candidates (<= 1 Mi) a second streaming pass collects them
// in stream order (per-warp contiguous ranges, offsets from the miss counts => deterministic) and
// they are replayed with their true positions: the result is then exact whenever the replay is,
// in particular bit-identical to the reference when #distinct <= capacity.
#pragma once

namespace crick {

constexpr long long kHHEmpty = (long long)0x8000000000000000ull;   // never a candidate
constexpr int kHHSampleMax = 1 << 20;
constexpr int kHHTabSlots = 1 << 21;
constexpr int kHHBuckets = 1 << 20;
constexpr long kHHMissCap = 1 << 20;
constexpr int kHHWarps = 16;            // warps per CTA of the counting pass
constexpr int kHHCandMax = 4096;

struct HHStatus {
    long long others;   // elements that are not candidates
    long long emax;     // fullest miss bucket
    int kept, ncand, fallback, pad;
};

__device__ __forceinline__ unsigned long long hh_mix(unsigned long long z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

__global__ void k_hh_fill(long long* p, long n, long long v) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        p[i] = v;
}

__global__ void k_hh_sample(const long long* __restrict__ item, long n, long S,
                            long long* tabkey, unsigned* tabcnt) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const long seg = n / S;
    const long pos = i * seg + (long)(hh_mix((unsigned long long)i) % (unsigned long long)seg);
    const long long key = item[pos];
    if (key == kHHEmpty) return;
    unsigned idx = hash64(key) & (unsigned)(kHHTabSlots - 1);
    for (;;) {
        const long long old = (long long)atomicCAS(reinterpret_cast<unsigned long long*>(&tabkey[idx]),
                                                   (unsigned long long)kHHEmpty, (unsigned long long)key);
        if (old == kHHEmpty || old == key) { atomicAdd(&tabcnt[idx], 1u); return; }
        idx = (idx + 1) & (unsigned)(kHHTabSlots - 1);
    }
}

// inclusive scan of one int per thread over a 1024-thread CTA
__device__ __forceinline__ int hh_block_scan(int v, int* s_w32) {
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

// One CTA.  cand_key[0 .. ncand): the C most frequent sampled items (ties by table slot);
// lut_key / lut_val: linear-probing image (L slots, L = 2 C) mapping a candidate to its index.
__global__ void __launch_bounds__(1024) k_hh_select(const long long* __restrict__ tabkey,
                                                    const unsigned* __restrict__ tabcnt, int C,
                                                    long long* cand_key, long long* lut_key,
                                                    unsigned short* lut_val, int L, HHStatus* status) {
    __shared__ int s_hist[1024];
    __shared__ int s_w32[32];
    __shared__ int s_tstar, s_above, s_quota, s_run_gt, s_run_eq;
    const int tid = threadIdx.x;
    s_hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < kHHTabSlots; i += 1024) {
        const unsigned c = tabcnt[i];
        if (c) atomicAdd(&s_hist[c < 1023u ? c : 1023u], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int cum = 0, tstar = 0, above = 0;
        for (int t = 1023; t >= 1; --t) {
            if (cum + s_hist[t] >= C) { tstar = t; above = cum; break; }
            cum += s_hist[t];
            above = cum;
        }
        s_tstar = tstar; s_above = above;
        s_quota = tstar > 0 ? C - above : 0;   // how many of the entries == tstar are taken
        s_run_gt = 0; s_run_eq = 0;
    }
    __syncthreads();
    const int tstar = s_tstar, above = s_above, quota = s_quota;
    // counts above 1023 land in the last histogram bin: they compare as 1023 here as well
    for (int base = 0; base < kHHTabSlots; base += 4096) {
        const int i0 = base + tid * 4;
        const uint4 c4 = *reinterpret_cast<const uint4*>(tabcnt + i0);
        const unsigned cc[4] = {c4.x, c4.y, c4.z, c4.w};
        int gt = 0, eq = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = (int)(cc[k] < 1023u ? cc[k] : 1023u);
            gt += c > tstar ? 1 : 0;
            eq += (tstar > 0 && c == tstar) ? 1 : 0;
        }
        const int pk = hh_block_scan((gt << 16) | eq, s_w32);   // <= 4096 per field
        int at_gt = s_run_gt + (pk >> 16) - gt, at_eq = s_run_eq + (pk & 0xffff) - eq;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = (int)(cc[k] < 1023u ? cc[k] : 1023u);
            if (c > tstar) cand_key[at_gt++] = tabkey[i0 + k];
            else if (tstar > 0 && c == tstar) { if (at_eq < quota) cand_key[above + at_eq] = tabkey[i0 + k]; ++at_eq; }
        }
        __syncthreads();
        if (tid == 1023) { s_run_gt += pk >> 16; s_run_eq += pk & 0xffff; }
        __syncthreads();
    }
    const int total_eq = s_run_eq;
    const int ncand = above + (total_eq < quota ? total_eq : quota);
    for (int i = tid; i < L; i += 1024) lut_key[i] = kHHEmpty;
    __syncthreads();
    for (int j = tid; j < ncand; j += 1024) {
        const long long key = cand_key[j];
        unsigned idx = hash64(key) & (unsigned)(L - 1);
        for (;;) {
            const long long old = (long long)atomicCAS(reinterpret_cast<unsigned long long*>(&lut_key[idx]),
                                                       (unsigned long long)kHHEmpty, (unsigned long long)key);
            if (old == kHHEmpty) { lut_val[idx] = (unsigned short)j; break; }
            idx = (idx + 1) & (unsigned)(L - 1);
        }
    }
    if (tid == 0) status->ncand = ncand;
}

// The streaming pass.  Warp w of the grid owns the contiguous elements [w n / W, (w+1) n / W).
// collect == 0: count.  collect == 1: write the misses, in stream order, at miss_off[w] ...
__global__ void __launch_bounds__(kHHWarps * 32) k_hh_count(const long long* __restrict__ item, long n,
                                                            const long long* __restrict__ lut_key,
                                                            const unsigned short* __restrict__ lut_val,
                                                            int L, int C, unsigned* rows_cnt,
                                                            unsigned* rows_last, unsigned* buckets,
                                                            unsigned long long* warp_miss, int collect,
                                                            const unsigned long long* __restrict__ miss_off,
                                                            long long* miss_item, long long* miss_pos,
                                                            long base_seen) {
    extern __shared__ __align__(16) unsigned char sm[];
    long long* s_key = reinterpret_cast<long long*>(sm);                    // [L]
    unsigned* s_cnt = reinterpret_cast<unsigned*>(s_key + L);               // [C]
    unsigned* s_last = s_cnt + C;                                           // [C]
    unsigned short* s_val = reinterpret_cast<unsigned short*>(s_last + C);  // [L]
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    for (int i = tid; i < L; i += blockDim.x) { s_key[i] = lut_key[i]; s_val[i] = lut_val[i]; }
    for (int i = tid; i < C; i += blockDim.x) { s_cnt[i] = 0; s_last[i] = 0; }
    __syncthreads();
    const long W = (long)gridDim.x * kHHWarps;
    const long w = (long)blockIdx.x * kHHWarps + wib;
    const long cta_lo = (long)(((__int128)((long)blockIdx.x * kHHWarps) * n) / W);
    const long lo = (long)(((__int128)w * n) / W), hi = (long)(((__int128)(w + 1) * n) / W);
    const unsigned lmask = (unsigned)L - 1u;
    unsigned long long nmiss = 0;
    unsigned long long out = collect ? miss_off[w] : 0ull;
    constexpr int U = 4;
    for (long base = lo; base < hi; base += 32 * U) {
        long long key[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long i = base + u * 32 + lane;
            key[u] = i < hi ? item[i] : kHHEmpty;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long i = base + u * 32 + lane;
            const bool valid = i < hi;
            int v = -1;
            if (valid && key[u] != kHHEmpty) {
                unsigned idx = hash64(key[u]) & lmask;
                for (;;) {
                    const long long k = s_key[idx];
                    if (k == key[u]) { v = s_val[idx]; break; }
                    if (k == kHHEmpty) break;
                    idx = (idx + 1) & lmask;
                }
            }
            const bool miss = valid && v < 0;
            if (!collect) {
                if (v >= 0) {
                    atomicAdd(&s_cnt[v], 1u);
                    atomicMax(&s_last[v], (unsigned)(i - cta_lo) + 1u);
                } else if (miss) {
                    atomicAdd(&buckets[(unsigned)(hh_mix((unsigned long long)key[u]) >> 40) & (kHHBuckets - 1)], 1u);
                    ++nmiss;
                }
            } else {
                const unsigned mm = __ballot_sync(CRICK_FULL, miss);
                if (miss) {
                    const unsigned long long at = out + __popc(mm & ((1u << lane) - 1u));
                    miss_item[at] = key[u];
                    miss_pos[at] = base_seen + i;
                }
                out += __popc(mm);
            }
        }
    }
    if (!collect) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) nmiss += __shfl_xor_sync(CRICK_FULL, nmiss, d);
        if (lane == 0) warp_miss[w] = nmiss;
        __syncthreads();
        for (int i = tid; i < C; i += blockDim.x) {
            rows_cnt[(size_t)blockIdx.x * C + i] = s_cnt[i];
            rows_last[(size_t)blockIdx.x * C + i] = s_last[i];
        }
    }
}

// exclusive scan of the per-warp miss counts (W <= 4096)
__global__ void __launch_bounds__(1024) k_hh_scan(const unsigned long long* warp_miss, int W,
                                                  unsigned long long* miss_off) {
    __shared__ unsigned long long s_tot[1024];
    const int tid = threadIdx.x;
    const int per = (W + 1023) / 1024;
    unsigned long long loc = 0;
    for (int k = 0; k < per; ++k) { int i = tid * per + k; if (i < W) loc += warp_miss[i]; }
    s_tot[tid] = loc;
    __syncthreads();
    if (tid == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 1024; ++i) { unsigned long long t = s_tot[i]; s_tot[i] = run; run += t; }
    }
    __syncthreads();
    unsigned long long run = s_tot[tid];
    for (int k = 0; k < per; ++k) { int i = tid * per + k; if (i < W) { miss_off[i] = run; run += warp_miss[i]; } }
}

// One CTA: exact candidate counts -> sorted -> the batch's summary written into `tmp` (a full
// state image: slot i = list position i, hash map included).
__global__ void __launch_bounds__(1024) k_hh_finish(const unsigned* __restrict__ rows_cnt,
                                                    const unsigned* __restrict__ rows_last, int G, int C,
                                                    const long long* __restrict__ cand_key,
                                                    const unsigned* __restrict__ buckets,
                                                    const unsigned long long* __restrict__ warp_miss,
                                                    long n, long base_seen, unsigned char* tmp, long cap,
                                                    int hcap, HHStatus* status) {
    extern __shared__ __align__(16) unsigned char sm[];
    long long* kc = reinterpret_cast<long long*>(sm);   // [C] counts
    long long* kl = kc + C;                             // [C] latest positions
    int* ks = reinterpret_cast<int*>(kl + C);           // [C] candidate index
    __shared__ unsigned long long s_red[1024];
    __shared__ long long s_others, s_emax;
    __shared__ int s_kept;
    const int tid = threadIdx.x;
    const int ncand = status->ncand;
    const long W = (long)G * kHHWarps;
    for (int c = tid; c < C; c += 1024) {
        long long cnt = 0, last = -1;
        if (c < ncand) {
            for (int g = 0; g < G; ++g) {
                const unsigned rc = rows_cnt[(size_t)g * C + c];
                if (rc) {
                    cnt += rc;
                    const long cta_lo = (long)(((__int128)((long)g * kHHWarps) * n) / W);
                    const long long p = base_seen + cta_lo + (long long)rows_last[(size_t)g * C + c] - 1;
                    last = p > last ? p : last;
                }
            }
        }
        kc[c] = cnt > 0 ? cnt : -1;          // unused slots sort to the end
        kl[c] = cnt > 0 ? last : 0x7fffffffffffffffll;
        ks[c] = c;
    }
    unsigned long long oth = 0, emax = 0;
    for (long i = tid; i < W; i += 1024) oth += warp_miss[i];
    for (int i = tid; i < kHHBuckets; i += 1024) { unsigned b = buckets[i]; emax = b > emax ? b : emax; }
    s_red[tid] = oth;
    __syncthreads();
    for (int d = 512; d > 0; d >>= 1) { if (tid < d) s_red[tid] += s_red[tid + d]; __syncthreads(); }
    if (tid == 0) s_others = (long long)s_red[0];
    __syncthreads();
    s_red[tid] = emax;
    __syncthreads();
    for (int d = 512; d > 0; d >>= 1) {
        if (tid < d) s_red[tid] = s_red[tid] > s_red[tid + d] ? s_red[tid] : s_red[tid + d];
        __syncthreads();
    }
    if (tid == 0) s_emax = (long long)s_red[0];
    __syncthreads();
    // bitonic sort by (count desc, latest position asc, index asc)
    for (int k = 2; k <= C; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < C / 2; t += 1024) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int q = i | j;
                const bool up = (i & k) == 0;
                const bool q_first = kc[q] > kc[i] ||
                                     (kc[q] == kc[i] && (kl[q] < kl[i] || (kl[q] == kl[i] && ks[q] < ks[i])));
                if (q_first == up) {
                    long long t1 = kc[i]; kc[i] = kc[q]; kc[q] = t1;
                    t1 = kl[i]; kl[i] = kl[q]; kl[q] = t1;
                    int t2 = ks[i]; ks[i] = ks[q]; ks[q] = t2;
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        int lo = 0, hi = C;                 // number of entries with a positive count
        while (lo < hi) { int mid = (lo + hi) >> 1; if (kc[mid] > 0) lo = mid + 1; else hi = mid; }
        s_kept = lo < cap ? lo : (int)cap;
    }
    __syncthreads();
    const int K = s_kept;
    const long long others = s_others;
    // bound for every item that is not kept: the next candidate's exact count, or (for items
    // that were never candidates) the fullest miss bucket -- unless the misses are replayed
    const bool replay = others <= kHHMissCap;
    long long bound = (K < C && kc[K] > 0) ? kc[K] : 0;
    if (!replay && s_emax > bound) bound = s_emax;
    const long long m = K == cap ? kc[K - 1] : 0;   // what the merge will use as the bound (:297-300)
    const bool fallback = K < cap && bound > 0;      // not full, yet something is left out
    const long long delta = (K == cap && bound > m) ? bound - m : 0;
    SpsvView g = spsv_carve(tmp, cap, hcap);
    for (int i = tid; i < hcap; i += 1024) g.hval[i] = -1;
    __syncthreads();
    for (int i = tid; i < K; i += 1024) {
        const long long key = cand_key[ks[i]];
        g.item[i] = key; g.count[i] = kc[i] + delta; g.error[i] = delta; g.last[i] = kl[i];
        g.order[i] = i; g.pos[i] = i;
        unsigned idx = hash64(key) & (unsigned)(hcap - 1);
        for (;;) {
            if (atomicCAS(&g.hval[idx], -1, i) == -1) { g.hkey[idx] = key; break; }
            idx = (idx + 1) & (unsigned)(hcap - 1);
        }
    }
    if (tid == 0) {
        SpsvHdr h; h.capacity = cap; h.size = K; h.hcap = hcap; h.tomb = 0;
        h.seen = base_seen + n; h.mixed = 0; h.pad = 0;
        *g.hdr = h;
        status->others = others; status->emax = s_emax; status->kept = K; status->fallback = fallback ? 1 : 0;
    }
}

}  // namespace crick
