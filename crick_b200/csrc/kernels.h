// crick_b200/csrc/kernels.h -- host-callable launchers (internal to libcrick_cuda.so)
#pragma once
#include "common.cuh"

struct crick_stats;
namespace crick {

struct StatsPart;   // stats_dev.cuh

// ---- reference-compatible path (tdigest_kernels.cu)
cudaError_t launch_ingest(const BatchView& v, long ngroups, const double* X, const double* W,
                          double wsc, long sx, long sw, const long* offsets, long fixed_len,
                          long row_stride, const long* dmap, int flags, cudaStream_t st,
                          bool pending_uniform = false);
// pending_uniform: the caller knows that W == nullptr and every item still waiting in a buffer
// of these digests has weight wsc (or the buffers are empty), so no per-item weights are kept
// thread-per-digest variant for many groups (tdigest_lanes.cu); same arguments and results
bool ingest_lanes_wanted(long ngroups);
int ingest_lanes_mode(int mode);   // -1 auto, 0 warp per digest, 1 thread per digest; returns previous
double ingest_lanes_margin(double m);   // decision window of the thread-per-digest kernel; returns previous
cudaError_t ingest_lanes_fallbacks(unsigned long long* out, bool reset);
cudaError_t launch_ingest_lanes(const BatchView& v, long ngroups, const double* X, const double* W,
                                double wsc, long sx, long sw, const long* offsets, long fixed_len,
                                long row_stride, const long* dmap, int flags, cudaStream_t st,
                                bool pending_uniform);
cudaError_t launch_flush(const BatchView& v, long first, long step, long count, cudaStream_t st);
cudaError_t launch_merge_pairs(const BatchView& dv, const BatchView& sv, long dfirst, long dstep,
                               long sfirst, long sstep, long count, cudaStream_t st);
cudaError_t launch_scale(const BatchView& v, long d, double factor, cudaStream_t st);
cudaError_t launch_query_prep(const BatchView& v, long first, long count, cudaStream_t st);
cudaError_t launch_query(const BatchView& v, long d, bool cdf, const double* in, double* out,
                         long nq, int sm_count, cudaStream_t st);
cudaError_t launch_query_grouped(const BatchView& v, long ngroups, bool cdf, const double* in,
                                 double* out, int nq, cudaStream_t st);

cudaError_t launch_init_hdr(DevHdr* hdr, long n, cudaStream_t st);
cudaError_t launch_pack_record(const BatchView& v, long d, double* rec, cudaStream_t st);
cudaError_t launch_merge_records(const BatchView& v, long d, const double* recs, int nrec, int skip,
                                 cudaStream_t st);
// all records in one greedy pass over the sorted union (tdigest_parallel.cu);
// cudaErrorInvalidValue = does not fit, use launch_merge_records
// fresh: the digest is to be taken as empty whatever its memory holds (a reset that was not launched)
cudaError_t launch_merge_records_bulk(const BatchView& v, long d, const double* recs, int nrec,
                                      cudaStream_t st, bool fresh = false);
cudaError_t launch_copy_digest(const BatchView& dv, long dd, const BatchView& sv, long sd,
                               cudaStream_t st);
cudaError_t launch_group_meta(const BatchView& v, long ng, double* out, cudaStream_t st);
// all centroids of a (flushed) group as dense arrays: offsets[ng + 1] (exclusive scan of the
// centroid counts), minmax[2] (over the non-empty digests), then X / W gathered
cudaError_t launch_group_offsets(const BatchView& v, long ng, long* offsets, double* minmax, cudaStream_t st);
cudaError_t launch_group_gather(const BatchView& v, long ng, const long* offsets, double* X, double* W,
                                cudaStream_t st);
cudaError_t launch_widen_minmax(const BatchView& v, long d, const double* minmax, cudaStream_t st);

// ---- parallel k-bucket path (tdigest_parallel.cu)
size_t parallel_scratch_bytes(double compression, int sm_count);
// X/W device pointers (W may be null -> scalar wsc). Accumulates the whole array
// into digest `d` of `v` (which must have an empty buffer): sample -> edges ->
// accumulate -> finalize, all on `st`, no host synchronisation.
// bad_weights != nullptr: validate w on the device (finite and > 0, tdigest.pyx:304-305);
// synchronises, and leaves the digest untouched when *bad_weights comes back non-zero.
cudaError_t launch_parallel_update(const BatchView& v, long d, const double* X, const double* W,
                                   double wsc, long n, void* scratch, int sm_count,
                                   cudaStream_t st, int* bad_weights, StatsPart* stats_parts,
                                   bool defer_check = false);
cudaError_t parallel_bad_weights(void* scratch, cudaStream_t st, int* bad);
// defer_check: validate like bad_weights != nullptr does (a failed check leaves the digest
// untouched) but keep the verdict on the device, sticky, until parallel_take_bad_weights reads
// and clears it -- the update then needs no host synchronisation at all
cudaError_t parallel_take_bad_weights(void* scratch, cudaStream_t st, int* bad);
// Sharded update over peer memory (k_exchange_merge): this rank's shard is binned, its bucket points
// are stored into every rank's exchange buffer (bufs[r], peer pointers, bufs[rank] = own) and the
// digest absorbs all ranks' points of step `step` (1, 2, ... : the same on every rank).
// cudaErrorInvalidValue = does not fit (compression too large): use the all-gather reducer.
size_t p2p_rec_doubles(double comp);
size_t p2p_buffer_bytes(double comp, int world);
cudaError_t launch_parallel_prep(const BatchView& v, const double* X, const double* W, double wsc, long n,
                                 void* scratch, int sm_count, cudaStream_t st);
cudaError_t launch_parallel_accumulate(const BatchView& v, const double* X, const double* W, double wsc, long n,
                                       void* scratch, int sm_count, cudaStream_t st);
cudaError_t launch_parallel_pass(const BatchView& v, const double* X, const double* W, double wsc, long n,
                                 void* scratch, int sm_count, cudaStream_t st);
cudaError_t launch_parallel_tail(const BatchView& v, long d, void* scratch, int sm_count, cudaStream_t st,
                                 bool checked);
cudaError_t launch_parallel_update_exchange(const BatchView& v, long d, const double* X, const double* W,
                                            double wsc, long n, void* scratch, int sm_count,
                                            cudaStream_t st, bool check_weights, int rank, int world,
                                            void* const* bufs, unsigned long long step);
// The same pipeline in pieces, for host-resident input streamed in chunks: edges
// from a caller-provided sample (copied to parallel_sample_x/w), any number of
// accumulate calls (first = 1 on the first), then one finalize.
cudaError_t parallel_begin_from_sample(const BatchView& v, long ns, void* scratch, cudaStream_t st);
// stats_parts != nullptr: also accumulate SummaryStats moments of x (count 1 each) into one
// StatsPart per CTA; idx_base = index of X[0] in the caller's whole array
cudaError_t parallel_accumulate(const BatchView& v, const double* X, const double* W, double wsc,
                                long n, void* scratch, int sm_count, cudaStream_t st, int first,
                                StatsPart* stats_parts, long idx_base);
cudaError_t parallel_finalize(const BatchView& v, long d, void* scratch, int sm_count,
                              cudaStream_t st, bool honor_badw = false);
// debug/test accessor: copy the bucket edges of the last parallel update to the host
int parallel_debug_edges(void* scratch, double* out, int max_out, cudaStream_t st);
int parallel_debug_header(void* scratch, void* out256, cudaStream_t st);
int parallel_debug_buckets(void* scratch, double comp, int sm_count, double* out_w, double* out_s,
                           double* out_minmax, int max_out, cudaStream_t st);
double* parallel_sample_x(void* scratch);
double* parallel_sample_w(void* scratch);
long parallel_sample_capacity(void* scratch);
// host mirrors of the device sampler, so that host-resident input picks the same sample
long parallel_sample_len(long n);
long parallel_sample_pos(long i, long n, long S);

// fused moments: per-CTA StatsPart slots written by k_bin_accumulate, folded by stats.cu
int stats_fused_begin(crick_stats* s, cudaStream_t st, StatsPart** parts, int* nparts);
cudaError_t stats_fused_finish(crick_stats* s, int nused, cudaStream_t st);

void profile_enable(int on);
int profile_collect(double* total_ms, long long* launches);

// ---- input coercion (synth.cu): safe casts of CUDA arrays to float64 (kind 'f' 2/4/8 bytes,
// 'i' / 'u' 1/2/4/8 bytes, 'b' bool), cudaErrorInvalidValue for anything else; and the wrapper's
// weight check (w finite and > 0, tdigest.pyx:304-305) as a tiny reduction: *flag |= 1 if violated
cudaError_t launch_cast_f64(const void* src, int kind, int itemsize, long n, double* dst, cudaStream_t st);
cudaError_t launch_check_weights(const double* w, long n, int* flag, cudaStream_t st);

// ---- synthetic data (synth.cu)
cudaError_t launch_synth(double* x, double* w, long n, unsigned long long seed, long offset,
                         int kind, cudaStream_t st);

}  // namespace crick
