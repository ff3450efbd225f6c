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
//   2. k_hh_hist ... k_hh_lut  the C = 4 x capacity most frequent sampled items become
//                     candidates (histogram of the sample counts, threshold, compaction in slot
//                     order => deterministic) + a linear-probing lookup image for them;
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
    long long emax;     // fullest miss bucket (sum of the counts of its elements)
    int kept, ncand, fallback;
    int npos;           // candidates that occur in the batch at all
    int lutfail, pad;   // k_hh_lut2 could not place every candidate (cuckoo insertion gave up)
};

__device__ __forceinline__ unsigned long long hh_mix(unsigned long long z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// lookup image / miss buckets: one 64-bit multiply (Fibonacci hashing); different bit ranges of
// the product index the image (top bits) and the buckets
__device__ __forceinline__ unsigned long long hh_hash(long long k) {
    return (unsigned long long)k * 0x9e3779b97f4a7c15ull;
}
__device__ __forceinline__ unsigned hh_lut_slot(unsigned long long h, int lbits) { return (unsigned)(h >> (64 - lbits)); }
__device__ __forceinline__ unsigned hh_bucket(unsigned long long h) { return (unsigned)(h >> 21) & (kHHBuckets - 1); }

__global__ void k_hh_fill(long long* p, long n, long long v) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        p[i] = v;
}

// The sampled items are first counted in a CTA-local table in shared memory (1024 draws into
// 2048 slots), then one global atomic per DISTINCT item of the CTA folds them into the global
// table: on a Zipf stream the top items take ~10 % of the draws, and 1 Mi threads hammering the
// same few global words with atomicCAS + atomicAdd cost 190 us (ncu, 1.25e8 items); the counts and
// the slot an item ends up in do not depend on the order of the atomics beyond what they did before.
constexpr int kHHSampleThreads = 1024;
constexpr int kHHLocalSlots = 2048;
__global__ void __launch_bounds__(kHHSampleThreads) k_hh_sample(const long long* __restrict__ item, long n,
                                                                long S, long long* tabkey, unsigned* tabcnt) {
    __shared__ long long s_key[kHHLocalSlots];
    __shared__ unsigned s_cnt[kHHLocalSlots];
    for (int j = threadIdx.x; j < kHHLocalSlots; j += kHHSampleThreads) { s_key[j] = kHHEmpty; s_cnt[j] = 0; }
    __syncthreads();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S) {
        const long seg = n / S;
        const long pos = i * seg + (long)(hh_mix((unsigned long long)i) % (unsigned long long)seg);
        const long long key = item[pos];
        if (key != kHHEmpty) {
            unsigned idx = hash64(key) & (unsigned)(kHHLocalSlots - 1);
            for (;;) {
                const long long old = (long long)atomicCAS(reinterpret_cast<unsigned long long*>(&s_key[idx]),
                                                           (unsigned long long)kHHEmpty, (unsigned long long)key);
                if (old == kHHEmpty || old == key) { atomicAdd(&s_cnt[idx], 1u); break; }
                idx = (idx + 1) & (unsigned)(kHHLocalSlots - 1);
            }
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < kHHLocalSlots; j += kHHSampleThreads) {
        const long long key = s_key[j];
        if (key == kHHEmpty) continue;
        const unsigned cnt = s_cnt[j];
        unsigned idx = hash64(key) & (unsigned)(kHHTabSlots - 1);
        for (;;) {
            const long long old = (long long)atomicCAS(reinterpret_cast<unsigned long long*>(&tabkey[idx]),
                                                       (unsigned long long)kHHEmpty, (unsigned long long)key);
            if (old == kHHEmpty || old == key) { atomicAdd(&tabcnt[idx], cnt); break; }
            idx = (idx + 1) & (unsigned)(kHHTabSlots - 1);
        }
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

// Candidate selection: the C most frequent sampled items, ties by table slot (deterministic).
//   k_hh_hist    histogram of the sample counts (saturated at 1023)
//   k_hh_thresh  t* = the count at which the cumulative histogram (from the top) reaches C
//   k_hh_chunks  per 4096-slot chunk: entries above t* / equal to t*
//   k_hh_place   candidates written in slot order: all entries above t*, then the first
//                C - #above entries equal to t*
//   k_hh_lut     linear-probing lookup image (L = 2 C slots) candidate -> index
struct HHSel { int tstar, above, quota, pad; };
constexpr int kHHChunk = 4096;
constexpr int kHHNChunk = kHHTabSlots / kHHChunk;

__global__ void __launch_bounds__(1024) k_hh_hist(const unsigned* __restrict__ tabcnt, unsigned* hist) {
    __shared__ unsigned s_hist[1024];
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    for (long i = (long)blockIdx.x * 1024 + threadIdx.x; i < kHHTabSlots / 4; i += (long)gridDim.x * 1024) {
        const uint4 c4 = reinterpret_cast<const uint4*>(tabcnt)[i];
        const unsigned cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) if (cc[k]) atomicAdd(&s_hist[cc[k] < 1023u ? cc[k] : 1023u], 1u);
    }
    __syncthreads();
    if (s_hist[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s_hist[threadIdx.x]);
}

// (1024 threads: thread i owns count t = 1023 - i; an inclusive scan gives the number of entries with
// a count >= t; t* is the largest t >= 1 at which that reaches C.  The serial form of this loop --
// one thread, 1023 dependent global loads -- took 63 us.)
__global__ void __launch_bounds__(1024) k_hh_thresh(const unsigned* __restrict__ hist, int C, HHSel* sel) {
    __shared__ int s_w32[32];
    __shared__ int s_first;
    const int i = threadIdx.x, t = 1023 - i;
    if (i == 0) s_first = 1024;
    const int h = t >= 1 ? (int)hist[t] : 0;
    const int incl = hh_block_scan(h, s_w32);        // (two barriers inside: s_first is visible)
    if (t >= 1 && incl >= C) atomicMin(&s_first, i);
    __syncthreads();
    const int f = s_first;
    if (f < 1024) {
        if (i == f) { sel->tstar = t; sel->above = incl - h; sel->quota = C - (incl - h); }
    } else if (i == 1022) {                            // t = 1: fewer than C sampled entries in all
        sel->tstar = 0; sel->above = incl; sel->quota = 0;
    }
}

// counts above 1023 sit in the last histogram bin: they compare as 1023 here as well
__device__ __forceinline__ void hh_classify4(const unsigned* tabcnt, int i0, int tstar, unsigned* cc,
                                             int& gt, int& eq) {
    const uint4 c4 = *reinterpret_cast<const uint4*>(tabcnt + i0);
    cc[0] = c4.x; cc[1] = c4.y; cc[2] = c4.z; cc[3] = c4.w;
    gt = 0; eq = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        cc[k] = cc[k] < 1023u ? cc[k] : 1023u;
        gt += (int)cc[k] > tstar ? 1 : 0;
        eq += (tstar > 0 && (int)cc[k] == tstar) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(1024) k_hh_chunks(const unsigned* __restrict__ tabcnt,
                                                    const HHSel* __restrict__ sel, int* chunk_gt,
                                                    int* chunk_eq) {
    __shared__ int s_w32[32];
    unsigned cc[4];
    int gt, eq;
    hh_classify4(tabcnt, blockIdx.x * kHHChunk + threadIdx.x * 4, sel->tstar, cc, gt, eq);
    const int pk = hh_block_scan((gt << 16) | eq, s_w32);   // <= 4096 per field
    if (threadIdx.x == 1023) { chunk_gt[blockIdx.x] = pk >> 16; chunk_eq[blockIdx.x] = pk & 0xffff; }
}

__global__ void __launch_bounds__(1024) k_hh_place(const long long* __restrict__ tabkey,
                                                   const unsigned* __restrict__ tabcnt,
                                                   const HHSel* __restrict__ sel,
                                                   const int* __restrict__ chunk_gt,
                                                   const int* __restrict__ chunk_eq, long long* cand_key) {
    __shared__ int s_w32[32];
    __shared__ int s_pre[2];
    const int tid = threadIdx.x, ch = blockIdx.x;
    const int tstar = sel->tstar, above = sel->above, quota = sel->quota;
    // entries above / equal in the chunks before this one
    int pg = 0, pe = 0;
    for (int j = tid; j < ch; j += 1024) { pg += chunk_gt[j]; pe += chunk_eq[j]; }
    pg = hh_block_scan(pg, s_w32);
    if (tid == 1023) s_pre[0] = pg;
    pe = hh_block_scan(pe, s_w32);
    if (tid == 1023) s_pre[1] = pe;
    unsigned cc[4];
    int gt, eq;
    const int i0 = ch * kHHChunk + tid * 4;
    hh_classify4(tabcnt, i0, tstar, cc, gt, eq);
    const int pk = hh_block_scan((gt << 16) | eq, s_w32);   // (ends with the barrier s_pre needs)
    __syncthreads();
    int at_gt = s_pre[0] + (pk >> 16) - gt, at_eq = s_pre[1] + (pk & 0xffff) - eq;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if ((int)cc[k] > tstar) cand_key[at_gt++] = tabkey[i0 + k];
        else if (tstar > 0 && (int)cc[k] == tstar) {
            if (at_eq < quota) cand_key[above + at_eq] = tabkey[i0 + k];
            ++at_eq;
        }
    }
}

__global__ void __launch_bounds__(1024) k_hh_lut(const long long* __restrict__ cand_key,
                                                 const HHSel* __restrict__ sel,
                                                 const int* __restrict__ chunk_eq, long long* lut_key,
                                                 unsigned short* lut_val, int L, HHStatus* status) {
    __shared__ int s_w32[32];
    __shared__ int s_ncand;
    const int tid = threadIdx.x;
    int te = 0;
    for (int j = tid; j < kHHNChunk; j += 1024) te += chunk_eq[j];
    te = hh_block_scan(te, s_w32);
    if (tid == 1023) s_ncand = sel->above + (te < sel->quota ? te : sel->quota);
    for (int i = tid; i < L; i += 1024) lut_key[i] = kHHEmpty;
    __syncthreads();
    const int ncand = s_ncand;
    for (int j = tid; j < ncand; j += 1024) {
        const long long key = cand_key[j];
        unsigned idx = hh_lut_slot(hh_hash(key), 31 - __clz(L));
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
// CT: unsigned for unit counts (cnt == nullptr), unsigned long long when every element carries
// its own count (the candidates' counts and the miss buckets are then sums of counts).
template <typename CT>
__global__ void __launch_bounds__(kHHWarps * 32) k_hh_count(const long long* __restrict__ item,
                                                            const long long* __restrict__ cnt, long n,
                                                            const long long* __restrict__ lut_key,
                                                            const unsigned short* __restrict__ lut_val,
                                                            int L, int C, CT* rows_cnt,
                                                            unsigned* rows_last, CT* buckets,
                                                            unsigned long long* warp_miss, int collect,
                                                            const unsigned long long* __restrict__ miss_off,
                                                            long long* miss_item, long long* miss_pos,
                                                            long long* miss_cnt, long base_seen) {
    extern __shared__ __align__(16) unsigned char sm[];
    long long* s_key = reinterpret_cast<long long*>(sm);                    // [L]
    CT* s_cnt = reinterpret_cast<CT*>(s_key + L);                           // [C]
    unsigned* s_last = reinterpret_cast<unsigned*>(s_cnt + C);              // [C]
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
    const int lbits = 31 - __clz(L);
    unsigned long long nmiss = 0;
    unsigned long long out = collect ? miss_off[w] : 0ull;
    constexpr int U = 4;
    for (long base = lo; base < hi; base += 32 * U) {
        long long key[U], cv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long i = base + u * 32 + lane;
            key[u] = i < hi ? item[i] : kHHEmpty;
            cv[u] = (cnt && i < hi) ? cnt[i] : 1;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long i = base + u * 32 + lane;
            const bool valid = i < hi;
            int v = -1;
            const unsigned long long h = hh_hash(key[u]);
            if (valid && key[u] != kHHEmpty) {
                unsigned idx = hh_lut_slot(h, lbits);
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
                    atomicAdd(&s_cnt[v], (CT)cv[u]);
                    atomicMax(&s_last[v], (unsigned)(i - cta_lo) + 1u);
                } else if (miss) {
                    atomicAdd(&buckets[hh_bucket(h)], (CT)cv[u]);
                    ++nmiss;
                }
            } else {
                const unsigned mm = __ballot_sync(CRICK_FULL, miss);
                if (miss) {
                    const unsigned long long at = out + __popc(mm & ((1u << lane) - 1u));
                    miss_item[at] = key[u];
                    miss_pos[at] = base_seen + i;
                    if (cnt) miss_cnt[at] = cv[u];
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

// Exact count and latest stream position of every candidate (sum / max over the CTAs of the
// counting pass), the number of misses and the fullest miss bucket.  Grid-stride, any grid.
template <typename CT>
__global__ void __launch_bounds__(256) k_hh_reduce(const CT* __restrict__ rows_cnt,
                                                   const unsigned* __restrict__ rows_last, int G, int C,
                                                   const CT* __restrict__ buckets,
                                                   const unsigned long long* __restrict__ warp_miss,
                                                   long n, long base_seen, long long* cand_cnt,
                                                   long long* cand_last, HHStatus* status) {
    extern __shared__ __align__(16) unsigned char sm[];
    long* s_lo = reinterpret_cast<long*>(sm);   // [G] first element of every counting CTA
    const long W = (long)G * kHHWarps;
    for (int g = threadIdx.x; g < G; g += blockDim.x)
        s_lo[g] = (long)(((__int128)((long)g * kHHWarps) * n) / W);
    __syncthreads();
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long)gridDim.x * blockDim.x;
    for (long c = tid; c < C; c += nthr) {
        long long cnt = 0, last = -1;
        for (int g = 0; g < G; ++g) {
            const CT rc = rows_cnt[(size_t)g * C + c];
            const unsigned rl = rows_last[(size_t)g * C + c];
            cnt += (long long)rc;
            const long long p = rc ? base_seen + s_lo[g] + (long long)rl - 1 : -1;
            last = p > last ? p : last;
        }
        cand_cnt[c] = cnt;
        cand_last[c] = last;
    }
    unsigned long long oth = 0, emax = 0;
    for (long i = tid; i < W; i += nthr) oth += warp_miss[i];
    for (long i = tid; i < kHHBuckets; i += nthr) { const unsigned long long b = buckets[i]; emax = b > emax ? b : emax; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        oth += __shfl_xor_sync(CRICK_FULL, oth, d);
        const unsigned long long o = __shfl_xor_sync(CRICK_FULL, emax, d);
        emax = o > emax ? o : emax;
    }
    if ((threadIdx.x & 31) == 0) {
        if (oth) atomicAdd(reinterpret_cast<unsigned long long*>(&status->others), oth);
        if (emax) atomicMax(reinterpret_cast<unsigned long long*>(&status->emax), emax);
    }
}

// One CTA: the candidates sorted by exact count -> the batch's summary written into `tmp` (a
// full state image: slot i = list position i, hash map included).
__global__ void __launch_bounds__(1024) k_hh_finish(const long long* __restrict__ cand_cnt,
                                                    const long long* __restrict__ cand_last, int C,
                                                    const long long* __restrict__ cand_key,
                                                    long n, long base_seen, unsigned char* tmp, long cap,
                                                    int hcap, long miss_cap, HHStatus* status) {
    extern __shared__ __align__(16) unsigned char sm[];
    long long* kc = reinterpret_cast<long long*>(sm);   // [C] counts
    long long* kl = kc + C;                             // [C] latest positions
    int* ks = reinterpret_cast<int*>(kl + C);           // [C] candidate index
    __shared__ int s_kept;
    const int tid = threadIdx.x;
    const int ncand = status->ncand;
    for (int c = tid; c < C; c += 1024) {
        const long long cnt = c < ncand ? cand_cnt[c] : 0;
        kc[c] = cnt > 0 ? cnt : -1;          // unused slots sort to the end
        kl[c] = cnt > 0 ? cand_last[c] : 0x7fffffffffffffffll;
        ks[c] = c;
    }
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
        status->npos = lo;
    }
    __syncthreads();
    const int K = s_kept;
    const long long others = status->others, s_emax = status->emax;
    // bound for every item that is not kept: the next candidate's exact count, or (for items
    // that were never candidates) the fullest miss bucket -- unless the misses are replayed
    const bool replay = others <= miss_cap;
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

// ---- round 2: the counting pass for unit counts -------------------------------------------------
// ncu of k_hh_count on Zipf(1.1) items (profiles/r1_hh_count_summary.md, r2_b): 0.32 of the HBM
// peak, 88 warp instructions per item-row (64-bit hash multiply, 64-bit index arithmetic, a
// branchy probe loop; 18 of 32 lanes active), `wait` the top stall, and every second item (the
// ones that are no candidates) a REDG into the 1 Mi L2 bucket counters.  This version:
//   * 32-bit arithmetic throughout: a 2-multiply hash of the two key halves, int offsets inside the
//     warp's range, 32-bit miss counts;
//   * the lookup image is keys only: the counters are indexed by the SLOT the key sits in (the
//     reduction maps slots back to candidates), so a hit needs no second load;
//   * the first probe of all items of a round is batched, the second is predicated on the lanes
//     that need one, only the rare third and later probes loop;
//   * ONE atomic per item, whatever it is: hits go to the slot's u32 counter, misses to one of NB
//     per-CTA miss buckets, both in shared memory and both by `red.shared.add.u32 ..., 1` (SASS
//     ATOMS.POPC.INC: the hardware folds the lanes of a warp that hit the same address -- a Zipf
//     stream's top items arrive several per warp), selected by address, not by a branch; no
//     global atomics.  The per-CTA bucket rows are summed by k_hh_reduce2;
//   * the latest position of a candidate (ATOMS.MAX) is only tracked while the summary may come
//     out exact -- every distinct sampled item is a candidate (sel->tstar == 0); list order by
//     latest update is only defined for exact summaries;
//   * STATS: the SummaryStats moments of the items (stats_update_ndarray :199-210, count 1) ride
//     in the same pass -- BASELINE config 5 reads its 8 B per item once for both sketches.
// MODE 0: count.  MODE 1: collect the misses in stream order at miss_off[w] (same warp ranges).
constexpr int kHH2U = 8;   // warps per CTA: a template parameter (24 plain, 16 with fused moments: registers)

__device__ __forceinline__ unsigned hh2_hash(long long k) {
    const unsigned lo = (unsigned)k, hi = (unsigned)((unsigned long long)k >> 32);
    return lo * 0x9E3779B1u + hi * 0x85EBCA77u;
}
// three slot choices per key (3-way cuckoo hashing: at load 0.5 the insertion succeeds with
// overwhelming probability and a lookup is three INDEPENDENT probes -- no probe loop whose trip
// count is the maximum over the 32 lanes of a warp)
__device__ __forceinline__ unsigned hh2_slot(unsigned h, int lbits, int which) {
    if (which == 0) return h >> (32 - lbits);
    if (which == 1) return (h * 0xC2B2AE3Du) >> (32 - lbits);
    return ((h ^ (h >> 15)) * 0x27D4EB2Fu) >> (32 - lbits);
}
__device__ __forceinline__ unsigned hh2_bucket(unsigned h, unsigned NB) {
    return __umulhi(__funnelshift_l(h, h, 13) * 0xC2B2AE3Du, NB);     // NB need not be a power of two
}

// lookup image for k_hh_count2: keys only, 3-way cuckoo (parallel insertion with atomicExch: a
// thread that displaces a key carries it to that key's next choice).  Which slot a key ends up in
// depends on the interleaving, but nothing downstream does: counters are per slot, and
// cand_slot[j] = the slot candidate j landed in.
__global__ void __launch_bounds__(1024) k_hh_lut2(const long long* __restrict__ cand_key,
                                                  const HHSel* __restrict__ sel,
                                                  const int* __restrict__ chunk_eq, long long* lut_key,
                                                  int* cand_slot, int L, HHStatus* status) {
    __shared__ int s_w32[32];
    __shared__ int s_ncand;
    const int tid = threadIdx.x;
    int te = 0;
    for (int j = tid; j < kHHNChunk; j += 1024) te += chunk_eq[j];
    te = hh_block_scan(te, s_w32);
    if (tid == 1023) s_ncand = sel->above + (te < sel->quota ? te : sel->quota);
    for (int i = tid; i < L; i += 1024) lut_key[i] = kHHEmpty;
    __syncthreads();
    const int ncand = s_ncand;
    const int lbits = 31 - __clz(L);
    bool failed = false;
    for (int j = tid; j < ncand; j += 1024) {
        long long key = cand_key[j];
        int which = 0;
        for (int iter = 0; iter < 2000 && key != kHHEmpty; ++iter) {
            const unsigned pos = hh2_slot(hh2_hash(key), lbits, which);
            const long long old = (long long)atomicExch(reinterpret_cast<unsigned long long*>(&lut_key[pos]),
                                                        (unsigned long long)key);
            if (old != kHHEmpty) {        // displaced: it moves on to the choice after the one it had here
                const unsigned ho = hh2_hash(old);
                const int w = hh2_slot(ho, lbits, 0) == pos ? 0 : (hh2_slot(ho, lbits, 1) == pos ? 1 : 2);
                which = w == 2 ? 0 : w + 1;
            }
            key = old;
        }
        failed |= key != kHHEmpty;
    }
    if (__syncthreads_or(failed) && tid == 0) status->lutfail = 1;
    __threadfence_block();
    for (int j = tid; j < ncand; j += 1024) {
        const long long key = cand_key[j];
        const unsigned h = hh2_hash(key);
        int slot = -1;
        for (int wch = 0; wch < 3; ++wch) {
            const unsigned pos = hh2_slot(h, lbits, wch);
            if (lut_key[pos] == key) { slot = (int)pos; break; }
        }
        cand_slot[j] = slot < 0 ? 0 : slot;
    }
    if (tid == 0) status->ncand = ncand;
}

// predicated 8-byte shared load: only the lanes with `p` issue a request (no branch)
__device__ __forceinline__ void hh2_lds_if(bool p, unsigned addr, unsigned& x, unsigned& y) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q ld.shared.v2.u32 {%0, %1}, [%2];\n\t}"
                 : "+r"(x), "+r"(y) : "r"(addr), "r"((unsigned)p));
}

struct HH2Ctx {
    unsigned a_key, a_cnt, a_last, a_bkt, NB, off_cta;
    int lbits;
};

// One round: U items per lane.  FULL: all 32 * U items exist (no validity logic at all).
// Branch-free: the second and third cuckoo choices are predicated loads, hit / miss pick the
// address of the single atomic.
template <int MODE, bool FULL, bool TRACK, int U>
__device__ __forceinline__ void hh2_round(const HH2Ctx& c, const long long* key, int base, int len,
                                          int lane, unsigned& nmiss, unsigned long long& out,
                                          long long* __restrict__ miss_item,
                                          long long* __restrict__ miss_pos, long pos0) {
    unsigned h[U], ix[U], kx[U], ky[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        h[u] = hh2_hash(key[u]);
        ix[u] = hh2_slot(h[u], c.lbits, 0);
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(kx[u]), "=r"(ky[u]) : "r"(c.a_key + 8u * ix[u]));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const unsigned klo = (unsigned)key[u], khi = (unsigned)((unsigned long long)key[u] >> 32);
        bool hit = ((kx[u] ^ klo) | (ky[u] ^ khi)) == 0u;
        const unsigned p1 = hh2_slot(h[u], c.lbits, 1), p2 = hh2_slot(h[u], c.lbits, 2);
        unsigned x1 = ~klo, y1 = 0u, x2 = ~klo, y2 = 0u;            // (never equal to the key)
        hh2_lds_if(!hit, c.a_key + 8u * p1, x1, y1);
        const bool hit1 = ((x1 ^ klo) | (y1 ^ khi)) == 0u;
        hh2_lds_if(!hit && !hit1, c.a_key + 8u * p2, x2, y2);
        const bool hit2 = ((x2 ^ klo) | (y2 ^ khi)) == 0u;
        ix[u] = hit1 ? p1 : (hit2 ? p2 : ix[u]);
        // kHHEmpty itself (INT64_MIN, also what padding lanes carry) is never a candidate
        hit = (hit || hit1 || hit2) && ((klo | (khi ^ 0x80000000u)) != 0u);
        const int i = base + u * 32 + lane;
        const bool valid = FULL || i < len;
        const bool miss = valid && !hit;
        if (MODE == 0) {
            const unsigned a = hit ? c.a_cnt + 4u * ix[u] : c.a_bkt + 4u * hh2_bucket(h[u], c.NB);
            if (FULL) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory");
            else if (valid) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory");
            if (TRACK && hit && valid)
                asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(c.a_last + 4u * ix[u]), "r"(c.off_cta + (unsigned)i + 1u) : "memory");
            nmiss += miss ? 1u : 0u;
        } else {
            const unsigned mm = __ballot_sync(CRICK_FULL, miss);
            if (miss) {
                const unsigned long long at = out + __popc(mm & ((1u << lane) - 1u));
                miss_item[at] = key[u];
                miss_pos[at] = pos0 + i;
            }
            out += __popc(mm);
        }
    }
}

template <int MODE, bool STATS, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
k_hh_count2(const long long* __restrict__ item, long n, const long long* __restrict__ lut_key, int L,
            int NB, const HHSel* __restrict__ sel, unsigned* __restrict__ rows_cnt,
            unsigned* __restrict__ rows_last, unsigned* __restrict__ rows_bkt,
            unsigned long long* __restrict__ warp_miss, const unsigned long long* __restrict__ miss_off,
            long long* __restrict__ miss_item, long long* __restrict__ miss_pos, long base_seen,
            StatsPart* __restrict__ stats_parts, int item_is_f64) {
    extern __shared__ __align__(16) unsigned char sm[];
    uint2* s_key = reinterpret_cast<uint2*>(sm);                            // [L] (lo, hi)
    unsigned* s_cnt = reinterpret_cast<unsigned*>(s_key + L);               // [L]   (MODE 0)
    unsigned* s_last = s_cnt + L;                                           // [L]
    unsigned* s_bkt = s_last + L;                                           // [NB]
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    for (int i = tid; i < L; i += blockDim.x) s_key[i] = reinterpret_cast<const uint2*>(lut_key)[i];
    if (MODE == 0) {
        for (int i = tid; i < 2 * L + NB; i += blockDim.x) s_cnt[i] = 0;    // cnt | last | buckets
    }
    __syncthreads();
    const bool track_last = MODE == 0 && sel->tstar == 0;
    const long W = (long)gridDim.x * NW;
    const long w = (long)blockIdx.x * NW + wib;
    const long cta_lo = (long)(((__int128)((long)blockIdx.x * NW) * n) / W);
    const long lo = (long)(((__int128)w * n) / W), hi = (long)(((__int128)(w + 1) * n) / W);
    const int len = (int)(hi - lo);                     // a warp's range: < 2^31 items
    const long long* __restrict__ src = item + lo;
    HH2Ctx c;
    c.a_key = (unsigned)__cvta_generic_to_shared(s_key);
    c.a_cnt = (unsigned)__cvta_generic_to_shared(s_cnt);
    c.a_last = (unsigned)__cvta_generic_to_shared(s_last);
    c.a_bkt = (unsigned)__cvta_generic_to_shared(s_bkt);
    c.NB = (unsigned)NB;
    c.off_cta = (unsigned)(lo - cta_lo);                // position of the range inside the CTA's
    c.lbits = 31 - __clz(L);
    unsigned nmiss = 0;
    unsigned long long out = MODE == 1 ? miss_off[w] : 0ull;
    const long pos0 = base_seen + lo;
    StatsAcc sa;
    StatsPart sp;
    stats_acc_fresh(sa);
    part_fresh(sp);
    constexpr int U = kHH2U;
    int base = 0;
    for (; base + 32 * U <= len; base += 32 * U) {
        long long key[U];
#pragma unroll
        for (int u = 0; u < U; ++u) key[u] = __ldg(src + base + u * 32 + lane);
        if (STATS) {
            double xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) xv[u] = item_is_f64 ? __longlong_as_double(key[u]) : (double)key[u];
            if (item_is_f64) stats_block_unit2<U, true>(sa, sp, xv, [&](int k) { return pos0 + base + k * 32 + lane; });
            else stats_block_unit2<U, false>(sa, sp, xv, [&](int k) { return pos0 + base + k * 32 + lane; });
        }
        if (track_last) hh2_round<MODE, true, true, U>(c, key, base, len, lane, nmiss, out, miss_item, miss_pos, pos0);
        else hh2_round<MODE, true, false, U>(c, key, base, len, lane, nmiss, out, miss_item, miss_pos, pos0);
    }
    if (base < len) {
        long long key[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * 32 + lane;
            key[u] = i < len ? __ldg(src + i) : kHHEmpty;
            if (STATS && i < len)
                stats_one<true>(sa, sp, item_is_f64 ? __longlong_as_double(key[u]) : (double)key[u], 1, pos0 + i);
        }
        hh2_round<MODE, false, true, U>(c, key, base, len, lane, nmiss, out, miss_item, miss_pos, pos0);
    }
    if (MODE == 0) {
        nmiss = __reduce_add_sync(CRICK_FULL, nmiss);
        if (lane == 0) warp_miss[w] = nmiss;
        __shared__ StatsPart s_sp[STATS ? NW : 1];
        if (STATS) {
            stats_thread_finish2(sa, sp);
            sp = stats_warp_reduce(sp);
            if (lane == 0) s_sp[wib] = sp;
        }
        __syncthreads();
        for (int i = tid; i < L; i += blockDim.x) {
            rows_cnt[(size_t)blockIdx.x * L + i] = s_cnt[i];
            rows_last[(size_t)blockIdx.x * L + i] = s_last[i];
        }
        for (int i = tid; i < NB; i += blockDim.x) rows_bkt[(size_t)blockIdx.x * NB + i] = s_bkt[i];
        if (STATS && tid == 0) {
            StatsPart t = s_sp[0];
            for (int wv = 1; wv < NW; ++wv) part_merge(t, s_sp[wv]);
            stats_parts[blockIdx.x] = t;
        }
    }
}

// Sums over the counting CTAs: exact count and latest position of every candidate, the number of
// misses and the fullest miss bucket.  Candidates first (C threads' worth of work), then the
// bucket columns; any grid.
__global__ void __launch_bounds__(256) k_hh_reduce2(const unsigned* __restrict__ rows_cnt,
                                                    const unsigned* __restrict__ rows_last,
                                                    const unsigned* __restrict__ rows_bkt, int G, int nw,
                                                    int C, int L, int NB, const int* __restrict__ cand_slot,
                                                    const unsigned long long* __restrict__ warp_miss,
                                                    long n, long base_seen, const HHSel* __restrict__ sel,
                                                    long long* cand_cnt, long long* cand_last,
                                                    HHStatus* status) {
    extern __shared__ __align__(16) unsigned char sm[];
    long* s_lo = reinterpret_cast<long*>(sm);   // [G] first element of every counting CTA
    const long W = (long)G * nw;
    for (int g = threadIdx.x; g < G; g += blockDim.x)
        s_lo[g] = (long)(((__int128)((long)g * nw) * n) / W);
    __syncthreads();
    const bool tracked = sel->tstar == 0;
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long)gridDim.x * blockDim.x;
    const int ncand = status->ncand;
    for (long c = tid; c < C; c += nthr) {
        long long cnt = 0, last = -1;
        if (c < ncand) {
            const int slot = cand_slot[c];              // the counters are indexed by lookup slot
            for (int g = 0; g < G; ++g) {
                const unsigned rc = rows_cnt[(size_t)g * L + slot];
                const unsigned rl = rows_last[(size_t)g * L + slot];
                cnt += (long long)rc;
                const long long p = rc ? base_seen + s_lo[g] + (long long)rl - 1 : -1;
                last = p > last ? p : last;
            }
        }
        cand_cnt[c] = cnt;
        // untracked (the summary cannot come out exact): ties fall by candidate index
        cand_last[c] = tracked ? last : (cnt ? base_seen + n - 1 : -1);
    }
    unsigned long long oth = 0, emax = 0;
    for (long i = tid; i < W; i += nthr) oth += warp_miss[i];
    for (long b = tid; b < NB; b += nthr) {
        unsigned long long t = 0;
        for (int g = 0; g < G; ++g) t += rows_bkt[(size_t)g * NB + b];
        emax = t > emax ? t : emax;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        oth += __shfl_xor_sync(CRICK_FULL, oth, d);
        const unsigned long long o = __shfl_xor_sync(CRICK_FULL, emax, d);
        emax = o > emax ? o : emax;
    }
    if ((threadIdx.x & 31) == 0) {
        if (oth) atomicAdd(reinterpret_cast<unsigned long long*>(&status->others), oth);
        if (emax) atomicMax(reinterpret_cast<unsigned long long*>(&status->emax), emax);
    }
}

}  // namespace crick
