/* include/crick_cuda.h -- C ABI of libcrick_cuda.so (sm_100a, CUDA 12.9).
 *
 * The drop-in boundary for crick's hot path.  In the reference the boundary is
 * the `cdef extern from "*_stubs.c"` blocks of the Cython classes
 * (crick/tdigest.pyx:13-40, crick/stats.pyx:8-31, crick/space_saving.pyx:10-68):
 * C structs malloc'ed per sketch and `static inline` functions #included into
 * the extension.  Here every one of those functions has a plain-C, opaque-handle
 * equivalent that runs on the GPU.  No Python.h, no NumPy C-API, no PyTorch types:
 * plain pointers and sizes only.  Pointers flagged `on_device` are CUDA device
 * pointers (from __cuda_array_interface__ / DLPack / cudaMalloc); otherwise they
 * are host pointers (pageable or pinned) and the library stages the copy.
 *
 * Conventions
 *   - int-returning functions: 0 = ok, <0 = error; crick_last_error() gives the
 *     text (thread-local).  NULL from a constructor = allocation/CUDA failure.
 *   - `stream`: a cudaStream_t passed as void*; NULL = the handle's own stream.
 *     Work is enqueued asynchronously; getters and host-destination copies
 *     synchronise that stream.
 *   - one handle = one thread at a time (as in the reference); different
 *     handles may be used concurrently from different threads.
 *   - there is NO CPU fallback: without a CUDA device every call fails.
 */
#ifndef CRICK_CUDA_H
#define CRICK_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct crick_tdigest crick_tdigest_t;
typedef struct crick_tdigest_group crick_tdigest_group_t;
typedef struct crick_stats crick_stats_t;
typedef struct crick_spsv crick_spsv_t;
typedef struct crick_p2p crick_p2p_t;
typedef struct crick_comm crick_comm_t;

/* ---- library ---------------------------------------------------------- */
const char *crick_last_error(void);
const char *crick_version(void);
/* number of CUDA devices visible (0 => nothing will work), current device, SM count */
int crick_device_count(void);
int crick_set_device(int device);
int crick_sm_count(void);
int crick_get_device(void);
/* kernels launched by this library since load (all threads) -- bench.py's gpu_launches */
int64_t crick_launch_count(void);
/* per-launch CUDA-event timing of the hot kernel (k_bin_accumulate), on the stream it is
 * launched on: enable, run, then collect (synchronises) -> summed ms and launch count */
int crick_profile_enable(int on);
int crick_profile_collect(double *total_ms, int64_t *launches);
/* grouped reference-compatible ingest: which kernel runs.  -1 (default) one thread per digest
 * from 4 warps of digests per SM (18944 groups on a B200), one warp per digest below; 0 always
 * one warp per digest; 1 always one thread per digest.  Results are bit-identical either way.
 * Returns the previous setting. */
int crick_group_kernel(int mode);
/* The thread-per-digest kernel decides "new centroid or fold" (tdigest_stubs.c:192-216) from a
 * per-centroid weight window instead of an asin per item, and evaluates the reference's own
 * k-scale pair only for cumulative weights inside the window (tdigest_lanes.cu, file comment).
 * margin: half-width of the window in quantile units (default 6e-13; >= 1: every item exact;
 * < 0: the kernel without the window).  NaN only queries.  Returns the previous value.
 * crick_group_exact_fallbacks: items decided by the exact pair on the current device since the
 * last reset (synchronises the device). */
double crick_group_margin(double margin);
int crick_group_exact_fallbacks(int64_t *out, int reset);
/* pinned host memory for callers that want true async H2D (bench e2e leg) */
void *crick_host_alloc(int64_t bytes);
void crick_host_free(void *p);
/* plain device memory + copies, so that callers need no other CUDA binding */
void *crick_device_alloc(int64_t bytes);
void crick_device_free(void *p);
int crick_memcpy_h2d(void *dst_dev, const void *src_host, int64_t bytes, void *stream);
int crick_memcpy_d2h(void *dst_host, const void *src_dev, int64_t bytes, void *stream);
int crick_stream_synchronize(void *stream);
/* a non-blocking stream on the current device, for callers with no other CUDA binding */
void *crick_stream_create(void);
void crick_stream_destroy(void *stream);
int crick_device_synchronize(void);
/* Safe cast of a contiguous CUDA array to float64 (what the reference's NpyIter does on the host,
 * tdigest_stubs.c:652-666, for every dtype np.can_cast(dtype, 'f8', 'safe') allows): kind 'f'
 * (2 / 4 / 8 bytes), 'i' or 'u' (1 / 2 / 4 / 8 bytes), 'b' (bool).  Asynchronous on `stream`. */
int crick_cast_to_f64(const void *src_dev, int kind, int itemsize, int64_t n, double *dst_dev, void *stream);
/* the wrapper's weight check (tdigest.pyx:304-305) for weights that live on the device: 1 if some
 * weight is not finite or not > 0, else 0 (synchronises `stream`) */
int crick_check_weights(const double *w_dev, int64_t n, void *stream);
/* synthetic counter-based streams for benchmarks/tests: element i depends only on
 * (seed, offset + i).  kind 0 lognormal(0,1) + U(.5,1.5) weights, 1 uniform, 2 normal,
 * 3 zipf(1.1)-like integers.  w_dev may be NULL. */
int crick_synth_fill(double *x_dev, double *w_dev, int64_t n, uint64_t seed, int64_t offset,
                     int kind, void *stream);

/* ---- TDigest ------------------------------------------------------------
 * update modes */
#define CRICK_MODE_AUTO 0      /* reference-compatible below CRICK_PARALLEL_MIN values, else parallel */
#define CRICK_MODE_REFERENCE 1 /* bit-exact with the reference's sequential algorithm */
#define CRICK_MODE_PARALLEL 2  /* k-bucket pre-aggregation (needs n >= CRICK_PARALLEL_MIN) */
#define CRICK_PARALLEL_MIN 2048

/* tdigest_new, tdigest_stubs.c:50-99 (compression clipped to [20, 1000]) */
crick_tdigest_t *crick_tdigest_new(double compression);
/* tdigest_free, tdigest_stubs.c:102-108 */
void crick_tdigest_free(crick_tdigest_t *t);
/* tdigest_add, tdigest_stubs.c:276-298 (queued on the host, applied in order at the next
 * device operation) */
int crick_tdigest_add(crick_tdigest_t *t, double x, double w);
/* tdigest_update_ndarray, tdigest_stubs.c:632-710, for contiguous float64 input that the
 * wrapper has already flattened in K-order.  w == NULL => scalar weight w_scalar. */
int crick_tdigest_update(crick_tdigest_t *t, const double *x, const double *w, double w_scalar,
                         int64_t n, int x_on_device, int w_on_device, int mode, void *stream);
/* The same with the wrapper's weight check (tdigest.pyx:304-305, "w must be > 0 and finite")
 * done on the device in the same pass instead of on the host.  Parallel mode only validates
 * (other modes leave *bad_weights = 0: validate on the host).  When some weight fails,
 * *bad_weights is set non-zero and the digest is left as it was (its buffer flushed).
 * Synchronises the stream. */
int crick_tdigest_update_checked(crick_tdigest_t *t, const double *x, const double *w,
                                 double w_scalar, int64_t n, int x_on_device, int w_on_device,
                                 int mode, void *stream, int *bad_weights);
/* The checked update for device-resident x and w WITHOUT any host synchronisation (parallel mode
 * only): the verdict of the weight check stays on the device; an update whose weights failed leaves
 * the digest as it was, and crick_tdigest_take_bad_weights() later returns 1 (and clears the
 * flag; it synchronises) if any deferred update since the last call was refused.  Lets a
 * multi-GPU step enqueue update -> export -> all-gather -> merge as one chain. */
int crick_tdigest_update_deferred(crick_tdigest_t *t, const double *x, const double *w,
                                  double w_scalar, int64_t n, int mode, void *stream);
int crick_tdigest_take_bad_weights(crick_tdigest_t *t);   /* bit 0 bad weights, bit 1 exchange timeout */
/* Pipelined form of crick_tdigest_update_deferred for back-to-back updates from device arrays
 * (tdigest.pyx:282-308 per call): the sample / edges kernels and the streaming pass are enqueued on
 * `stream`, the single-CTA tail that folds the bucket sums into the digest on `tail_stream` behind an
 * event, on alternating scratch sets -- so the next update's sample work (and the caller's own work
 * on `tail_stream`: export, all-gather, merge) overlaps.  Two different explicit streams; every
 * other call on the handle orders itself behind the tail as usual.  n >= CRICK_PARALLEL_MIN. */
int crick_tdigest_update_pipelined(crick_tdigest_t *t, const double *x, const double *w, double w_scalar,
                                   int64_t n, void *stream, void *tail_stream);
/* SMs a pipelined update's streaming kernel leaves to the tail stream (default 2, env
 * CRICK_PIPE_RESERVE; 0: same grid, hence bit-identical results, as the plain update).  Returns the
 * previous value; a negative argument only queries. */
int crick_pipeline_reserve_sms(int n_sms);
/* Overlapped sample phase for pipelined updates (default off, env CRICK_PIPE_OVERLAP=1): the streaming
 * passes run back to back on a stream of the handle's own and the sample / edges kernels of the NEXT
 * update run on `stream` while the current pass streams, on the reserved SMs (use a reserve of ~8
 * with it).  x and w must then stay valid until `tail_stream` has passed the update.  Returns the
 * previous setting; any argument other than 0 / 1 only queries. */
int crick_pipeline_overlap(int on);
/* ---- one stream sharded over the GPUs of a node: update + exchange + merge over peer memory -----
 * Every rank (one process per GPU) owns an exchange buffer of crick_p2p_buffer_bytes() bytes
 * (crick_p2p_alloc: plain cudaMalloc, zeroed), exports it with crick_ipc_get_handle (64 bytes, sent
 * to the other ranks by whatever transport the caller has), opens the others' handles with
 * crick_ipc_open_handle and builds a communicator from the `world` pointers (buffers[rank] = own).
 * crick_tdigest_update_exchange is COLLECTIVE: every rank calls it the same number of times with
 * its shard (device-resident; n = 0 or >= CRICK_PARALLEL_MIN); afterwards `t` on every rank has
 * absorbed ALL ranks' shards (identical digests).  One kernel per rank stores the rank's bucket
 * points into every rank's buffer over NVLink, waits for the others' and merges: no collective
 * library call, no host synchronisation.  Weights are validated like update_deferred
 * (crick_tdigest_take_bad_weights: bit 0 = bad weights, bit 1 = a rank did not arrive in 4 s). */
int64_t crick_p2p_buffer_bytes(double compression, int world);
void *crick_p2p_alloc(int64_t bytes);
void crick_p2p_free(void *p);
int crick_ipc_get_handle(void *dev_ptr, void *handle64);
void *crick_ipc_open_handle(const void *handle64);
int crick_ipc_close_handle(void *p);
crick_p2p_t *crick_p2p_new(int rank, int world, void *const *buffers);
void crick_p2p_destroy(crick_p2p_t *c);
int crick_tdigest_update_exchange(crick_tdigest_t *t, crick_p2p_t *c, const double *x, const double *w,
                                  double w_scalar, int64_t n, void *stream);
/* Parallel-mode update with the SummaryStats moments of x (count 1 per element, NaN skipped:
 * stats_update_ndarray, stats_stubs.c:199-210) accumulated into `stats` IN THE SAME PASS over
 * the data.  n >= CRICK_PARALLEL_MIN.  bad_weights may be NULL (no weight validation). */
int crick_tdigest_update_stats(crick_tdigest_t *t, crick_stats_t *stats, const double *x,
                               const double *w, double w_scalar, int64_t n, int x_on_device,
                               int w_on_device, void *stream, int *bad_weights);
/* tdigest_flush, tdigest_stubs.c:219-273 */
int crick_tdigest_flush(crick_tdigest_t *t);
/* tdigest_merge, tdigest_stubs.c:592-606, for each of `others` in order (flushes them) */
int crick_tdigest_merge(crick_tdigest_t *t, crick_tdigest_t *const *others, int n_others);
/* tdigest_scale, tdigest_stubs.c:609-629 (in place; the wrapper copies first, tdigest.pyx:338) */
int crick_tdigest_scale(crick_tdigest_t *t, double factor);
/* tdigest_quantile_ndarray / tdigest_cdf_ndarray, tdigest_stubs.c:519-589 / :410-480 */
int crick_tdigest_quantile(crick_tdigest_t *t, const double *q, double *out, int64_t n,
                           int on_device, void *stream);
int crick_tdigest_cdf(crick_tdigest_t *t, const double *x, double *out, int64_t n,
                      int on_device, void *stream);
/* struct fields read by tdigest.pyx:98,105-106,122,239-243.  min/max/ncentroids/total flush
 * nothing by themselves: call crick_tdigest_flush first where the reference does. */
double crick_tdigest_compression(crick_tdigest_t *t);
int crick_tdigest_size(crick_tdigest_t *t);        /* centroid slots, 2*ceil(c) */
int crick_tdigest_buffer_size(crick_tdigest_t *t); /* buffer slots */
double crick_tdigest_min(crick_tdigest_t *t);
double crick_tdigest_max(crick_tdigest_t *t);
int crick_tdigest_ncentroids(crick_tdigest_t *t);
int crick_tdigest_buffer_ncentroids(crick_tdigest_t *t);
double crick_tdigest_total_weight(crick_tdigest_t *t);
double crick_tdigest_buffer_total_weight(crick_tdigest_t *t);
/* centroids() copy-out, tdigest.pyx:231-244: out holds 2*ncentroids doubles (mean, weight) */
int crick_tdigest_get_centroids(crick_tdigest_t *t, double *out, int on_device);
/* device pointer of the live centroid array (zero-copy view, valid until the next mutation) */
const double *crick_tdigest_centroids_device(crick_tdigest_t *t);
/* __setstate__, tdigest.pyx:253-263 (bounds-checked here: n <= size) */
int crick_tdigest_set_state(crick_tdigest_t *t, const double *centroids, int n,
                            double total_weight, double min, double max);
/* back to the state tdigest_new leaves (tdigest_stubs.c:63-92) without releasing anything:
 * lets a reduction reuse one output digest per step (asynchronous on `stream`) */
int crick_tdigest_reset(crick_tdigest_t *t, void *stream);
/* test hook: bucket edges chosen by the last parallel-mode update (returns their count) */
int crick_tdigest_debug_edges(crick_tdigest_t *t, double *out, int max_out);
/* test hook: the 256-byte header block of the parallel scratch (lookup parameters, sticky flags, and
 * from byte 128 the globaltimer timestamps of the last k_exchange_merge: start, partials ready,
 * points stored, flags published, all ranks arrived, done) */
int crick_tdigest_debug_header(crick_tdigest_t *t, void *out256);
/* test hook: per-bucket (sum w, sum w*x) and (min, max) of the last parallel-mode update,
 * reduced in the kernel's own order (returns the bucket count = edges + 1) */
int crick_tdigest_debug_buckets(crick_tdigest_t *t, double *out_w, double *out_s,
                                double *out_minmax, int max_out);

/* fixed-size wire record for the multi-GPU all-gather:
 * doubles [ncentroids, total_weight, min, max, mean0, weight0, mean1, ...], padded to
 * crick_tdigest_record_doubles(t) = 4 + 2*size.  export flushes first; the record lives on
 * the device.  merge_records feeds `n_records` records (device, contiguous, in order)
 * through tdigest_merge semantics, skipping index `skip` (pass -1 to keep all). */
int crick_tdigest_record_doubles(crick_tdigest_t *t);
int crick_tdigest_export_record(crick_tdigest_t *t, double *record_dev, void *stream);
int crick_tdigest_merge_records(crick_tdigest_t *t, const double *records_dev, int n_records,
                                int skip, void *stream);

/* The same records merged in ONE greedy k-scale pass over the sorted union of their centroids
 * (what a single tdigest_flush of the union would build) instead of tdigest_merge's incremental
 * walk through the buffer: not bit-identical with merge(*others), same accuracy class, ~5x
 * shorter on the critical path of a multi-GPU step. */
int crick_tdigest_merge_records_bulk(crick_tdigest_t *t, const double *records_dev, int n_records,
                                     void *stream);

/* ---- grouped digests: many independent digests, one warp each ------------
 * (the added entry point; semantics per group are exactly
 *  `t = TDigest(c); t.update(values_g, weights_g)` in reference-compatible mode) */
crick_tdigest_group_t *crick_tdigest_group_new(double compression, int64_t n_groups);
void crick_tdigest_group_free(crick_tdigest_group_t *g);
int64_t crick_tdigest_group_count(crick_tdigest_group_t *g);
/* values: n_groups rows of row_len doubles, row r at values + r*row_stride (fixed-length
 * form, offsets == NULL) or group r = values[offsets[r] .. offsets[r+1]) (ragged form;
 * offsets has n_groups+1 int64 entries, same residency as values).  weights NULL => scalar. */
int crick_tdigest_group_update(crick_tdigest_group_t *g, const double *values,
                               const double *weights, double w_scalar, const int64_t *offsets,
                               int64_t row_len, int64_t row_stride, int on_device, void *stream);
int crick_tdigest_group_flush(crick_tdigest_group_t *g);
/* pairwise, level-by-level tree merge (left operand = self, right = other, exactly a nest of
 * tdigest_merge calls); the result lands in group 0 and is copied into `out` */
int crick_tdigest_group_merge_tree(crick_tdigest_group_t *g, crick_tdigest_t *out);
/* all digests of the group merged into `out` in ONE pass over their centroids: the centroids are
 * the (mean, weight) points tdigest_merge feeds to tdigest_add (:600-602), taken through the
 * parallel k-bucket update in one go.  Not bit-identical with nested merge() calls (use
 * crick_tdigest_group_merge_tree for that), same accuracy class; 65536 digests in well under a
 * millisecond instead of 16 dependent levels.  The group's digests are flushed, not consumed. */
int crick_tdigest_group_merge_all(crick_tdigest_group_t *g, crick_tdigest_t *out);
/* per-group queries: out[g*nq + j]; q on host or device per on_device */
int crick_tdigest_group_quantile(crick_tdigest_group_t *g, const double *q, int nq, double *out,
                                 int on_device, void *stream);
int crick_tdigest_group_cdf(crick_tdigest_group_t *g, const double *x, int nq, double *out,
                            int on_device, void *stream);
/* headers: out_meta[g*4 + {0,1,2,3}] = ncentroids, total_weight, min, max (host) */
int crick_tdigest_group_meta(crick_tdigest_group_t *g, double *out_meta);
/* centroids of group i (flushes it): up to 2*size doubles to host; returns ncentroids */
int crick_tdigest_group_get_centroids(crick_tdigest_group_t *g, int64_t i, double *out);
/* copy group i into a standalone digest handle (same compression) */
int crick_tdigest_group_extract(crick_tdigest_group_t *g, int64_t i, crick_tdigest_t *out);

/* ---- SummaryStats (stats_stubs.c) ----------------------------------------- */
crick_stats_t *crick_stats_new(void);                 /* stats_new :25-39  */
void crick_stats_free(crick_stats_t *s);              /* stats_free :42-44 */
int crick_stats_add(crick_stats_t *s, double x, int64_t count);      /* stats_add :92-95 */
/* stats_update_ndarray :139-226; count == NULL => scalar.  x is float64, or int64 when
 * x_is_int64 (converted on the fly as the NumPy safe cast does). */
int crick_stats_update(crick_stats_t *s, const void *x, int x_is_int64, const int64_t *count,
                       int64_t count_scalar, int64_t n, int x_on_device, int count_on_device,
                       void *stream);
/* which kernel array updates use: -1 (default) the sequential replay of stats_do_update :47-75
 * (bit-exact with the reference, one dependent chain: ~1.5 us per element) below 256 elements
 * and the parallel pass (Pebay pairwise merges; moments within ~1e-12 relative) from there;
 * 0 always the replay, 1 always the parallel pass.  Returns the previous setting. */
int crick_stats_kernel(int mode);
int crick_stats_merge(crick_stats_t *s, crick_stats_t *const *others, int n_others); /* :78-90 */
/* state = count + (sum, min, max, m2, m3, m4, first_value) + homogeneous, stats.pyx:77-91 */
int crick_stats_getstate(crick_stats_t *s, int64_t *count, double *f7, int *homogeneous);
int crick_stats_setstate(crick_stats_t *s, int64_t count, const double *f7, int homogeneous);
double crick_stats_mean(crick_stats_t *s);                       /* :98-100  */
double crick_stats_var(crick_stats_t *s, int64_t ddof);          /* :103-105 */
double crick_stats_std(crick_stats_t *s, int64_t ddof);          /* :108-110 */
double crick_stats_skew(crick_stats_t *s, int bias);             /* :113-123 */
double crick_stats_kurt(crick_stats_t *s, int fisher, int bias); /* :126-136 */
/* 9-double wire record for all-gather: count, sum, min, max, m2, m3, m4, homogeneous, first */
int crick_stats_export_record(crick_stats_t *s, double *record9_host);
int crick_stats_merge_record(crick_stats_t *s, const double *record9_host);
/* the same on the device, asynchronous on `stream`: export into a device buffer, merge
 * `n_records` records (`stride` doubles apart, in order) with stats_merge :78-90 each */
int crick_stats_export_record_dev(crick_stats_t *s, double *record9_dev, void *stream);
int crick_stats_merge_records_dev(crick_stats_t *s, const double *records_dev, int n_records,
                                  int64_t stride, void *stream);

/* ---- SpaceSaving, int64 / float64-bit-pattern items (space_saving_stubs.c.in) ---- */
crick_spsv_t *crick_spsv_new(int64_t capacity);                  /* spsv_int64_new :75-95 */
void crick_spsv_free(crick_spsv_t *s);                           /* :98-109 */
int crick_spsv_add(crick_spsv_t *s, int64_t item, int64_t count);            /* :213-250 */
/* spsv_int64_update_ndarray :367-451; count == NULL => scalar.
 * mode 0 auto: below 65536 elements the stream is replayed in order
 * (bit-exact with the reference, list order included); above, shard summaries are built in parallel and merged with
 * the reference's own merge (:289-364, Cafaro et al.) -- a valid Space-Saving summary for any
 * input, with exact counts and the reference's list order whenever #distinct <= capacity.
 * From 65536 elements (count 1 or a count array; capacity <= 1024) the batch is COUNTED instead of replayed:
 * the 4 x capacity most frequent items of a 1 Mi sample are counted exactly in one streaming
 * pass, the top `capacity` form a summary with error 0 (every other item is bounded by the next
 * candidate / the fullest of 1 Mi miss buckets; counts and errors are raised by the excess if
 * that bound exceeds the smallest kept count), merged with :289-364; if only a few elements
 * (<= max(4096, n/128), at most 1 Mi) missed the candidates they are replayed with their stream
 * positions, so the result is still bit-identical to the reference whenever #distinct <= capacity.
 * Below 65536 elements a summary nothing was added to yet also takes the counting path, but keeps
 * its result only when no eviction can have happened (else the in-order replay runs).
 * mode 1 forces the sequential replay, mode 2 the shard path, mode 3 the counting path (falls
 * back to 2 when it does not apply). */
int crick_spsv_update(crick_spsv_t *s, const int64_t *item, const int64_t *count,
                      int64_t count_scalar, int64_t n, int item_on_device, int count_on_device,
                      int mode, void *stream);
/* spsv_int64_update_ndarray :367-451 (count 1) and stats_update_ndarray (stats_stubs.c:139-226,
 * count 1) of the SAME items in ONE pass over them (BASELINE config 5: 8 B per item read once for
 * both sketches): the moments are accumulated inside the counting pass.  item_is_f64: the items
 * are float64 bit patterns (SpaceSaving dtype f8), else int64 (cast to double for the moments
 * like NumPy's safe cast).  Batches the counting path does not take (fewer than 65536 items,
 * capacity > 1024) update the two sketches one after the other: same results, two passes. */
int crick_spsv_update_stats(crick_spsv_t *s, crick_stats_t *stats, const int64_t *item, int64_t n,
                            int item_on_device, int item_is_f64, void *stream);
int crick_spsv_merge(crick_spsv_t *s, crick_spsv_t *const *others, int n_others); /* :289-364 */
int crick_spsv_set_state(crick_spsv_t *s, const int64_t *counters3, int64_t n);   /* :253-286 */
/* Multi-GPU tail: a summary as the fixed-size int64 record
 * [size, 0, (item, count, error) x capacity] in list order (device buffer of
 * crick_spsv_record_int64s() values); merge_records folds `n_records` such records -- all of this
 * handle's capacity -- into the handle one after the other, each exactly like
 * spsv_merge(self, other) :289-364 would.  Asynchronous on `stream`. */
/* back to the state spsv_new / stats_new leave, without releasing anything (asynchronous on `stream`) */
int crick_spsv_reset(crick_spsv_t *s, void *stream);
int crick_stats_reset(crick_stats_t *s, void *stream);
int64_t crick_spsv_record_int64s(crick_spsv_t *s);
int crick_spsv_export_record(crick_spsv_t *s, int64_t *dev_out, void *stream);
int crick_spsv_merge_records(crick_spsv_t *s, const int64_t *dev_records, int n_records, void *stream);
/* The same records merged in ONE pass (the k-way form of :289-364: an item a summary does not monitor
 * contributes that summary's smallest count to count and error; the `capacity` largest are kept):
 * the guarantees of nested merges without their list order, one kernel instead of n_records
 * dependent single-warp merges (8 x 1000 counters: ~0.1 ms instead of ~9 ms). */
int crick_spsv_merge_records_bulk(crick_spsv_t *s, const int64_t *dev_records, int n_records, void *stream);
int64_t crick_spsv_size(crick_spsv_t *s);
int64_t crick_spsv_capacity(crick_spsv_t *s);
/* int64_counters walk, space_saving.pyx:368-386: out = k x (item, count, error); returns rows */
int64_t crick_spsv_counters(crick_spsv_t *s, int64_t *out3, int64_t k);

/* ---- multi-GPU reduction inside the library: one process per GPU, NCCL loaded with dlopen ----
 * The reference has no multi-device path; its unit of reduction is merge(*others) of sketches
 * (tdigest.pyx:310-324, space_saving.pyx:343-366, stats.pyx:128-142).  Here every rank exports its
 * sketch as a fixed-size device record, one ncclAllGather moves the records over NVLink, and every
 * rank merges them in rank order with the kernels merge() uses: three enqueues on `stream`, no
 * host synchronisation, no PyTorch.  libnccl.so.2 (or $CRICK_NCCL_LIB) is opened on first use.
 *
 * crick_comm_unique_id: rank 0 obtains CRICK_COMM_ID_BYTES bytes and ships them to the other ranks
 * by any means (file, socket, MPI); crick_comm_init is collective over all ranks, on each
 * rank's current device.  *_allgather_merge leave in `out` (!= the input), identically on every
 * rank, what out.merge(sketch of rank 0, ..., sketch of rank world-1) gives: exact != 0 through the
 * reference's incremental merge, 0 through the one-pass k-way merge (*_merge_records_bulk).
 * crick_comm_allgather is the raw collective on device buffers (recv = world x bytes_per_rank). */
#define CRICK_COMM_ID_BYTES 128
int crick_comm_unique_id(void *id_out);
crick_comm_t *crick_comm_init(const void *id_bytes, int rank, int world);
void crick_comm_free(crick_comm_t *c);
int crick_comm_rank(crick_comm_t *c);
int crick_comm_world(crick_comm_t *c);
int crick_comm_allgather(crick_comm_t *c, const void *send_dev, void *recv_dev, int64_t bytes_per_rank,
                         void *stream);
int crick_tdigest_allgather_merge(crick_tdigest_t *t, crick_comm_t *c, crick_tdigest_t *out, int exact,
                                  void *stream);
int crick_spsv_allgather_merge(crick_spsv_t *s, crick_comm_t *c, crick_spsv_t *out, int exact,
                               void *stream);
int crick_stats_allgather_merge(crick_stats_t *s, crick_comm_t *c, crick_stats_t *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CRICK_CUDA_H */
