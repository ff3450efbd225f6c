/* oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin C-ABI shim around the UNMODIFIED reference C sources, compiled where they
 * lie under /root/reference (see oracle/build_oracle.py). Nothing from the
 * reference is copied into this repository: the three `#include`s below pull
 * the reference translation units in at build time (they are `static inline`
 * single-file "stubs", crick/common.h:1-11), and the functions here only
 * forward to them so that ctypes can call the reference's real code.
 *
 * Output: oracle/_ref/libcrick_ref.so (git-ignored; shipped to the GPU box).
 *
 * The only logic restated here (because it lives inside the NumPy-iterator
 * drivers that cannot be called without a live ndarray) is
 *   - the per-element loop of tdigest_update_ndarray  (tdigest_stubs.c:681-695)
 *   - the per-element loop + homogeneous tracking of stats_update_ndarray
 *     (stats_stubs.c:199-210)
 *   - the per-element loop of spsv_int64_update_ndarray
 *     (space_saving_stubs.c.in:422-435)
 *   - the query loops (tdigest_stubs.c:451-465, 560-574)
 * each a plain `for` over a contiguous array.
 */
#include <pthread.h>
#include <string.h>

#include "tdigest_stubs.c"        /* -I/root/reference/crick */
#include "stats_stubs.c"          /* -I/root/reference/crick */
#include "space_saving_stubs.c"   /* Tempita-expanded into oracle/_ref/ at build time */

#define API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ tdigest */
API void *ref_tdigest_new(double compression) { return tdigest_new(compression); }
API void ref_tdigest_free(void *T) { tdigest_free((tdigest_t *)T); }
API void ref_tdigest_add(void *T, double x, double w) { tdigest_add((tdigest_t *)T, x, w); }
API void ref_tdigest_flush(void *T) { tdigest_flush((tdigest_t *)T); }
API void ref_tdigest_merge(void *T, void *other) { tdigest_merge((tdigest_t *)T, (tdigest_t *)other); }
API void ref_tdigest_scale(void *T, double f) { tdigest_scale((tdigest_t *)T, f); }

/* loop of tdigest_update_ndarray, tdigest_stubs.c:681-695; w == NULL -> scalar */
API void ref_tdigest_update(void *T, const double *x, const double *w,
                            double w_scalar, long n) {
    long i;
    for (i = 0; i < n; i++)
        tdigest_add((tdigest_t *)T, x[i], w ? w[i] : w_scalar);
}

/* tdigest_quantile_ndarray loop, tdigest_stubs.c:560-574 */
API void ref_tdigest_quantile(void *T, const double *q, double *out, long n) {
    long i;
    tdigest_query_prep((tdigest_t *)T);
    for (i = 0; i < n; i++) out[i] = tdigest_quantile((tdigest_t *)T, q[i]);
}

/* tdigest_cdf_ndarray loop, tdigest_stubs.c:451-465 */
API void ref_tdigest_cdf(void *T, const double *x, double *out, long n) {
    long i;
    tdigest_query_prep((tdigest_t *)T);
    for (i = 0; i < n; i++) out[i] = tdigest_cdf((tdigest_t *)T, x[i]);
}

/* field access as done by tdigest.pyx:98,105-106,122,239-243,250-263 */
API double ref_tdigest_compression(void *T) { return ((tdigest_t *)T)->compression; }
API int ref_tdigest_size(void *T) { return ((tdigest_t *)T)->size; }
API int ref_tdigest_buffer_size(void *T) { return ((tdigest_t *)T)->buffer_size; }
API double ref_tdigest_min(void *T) { return ((tdigest_t *)T)->min; }
API double ref_tdigest_max(void *T) { return ((tdigest_t *)T)->max; }
API int ref_tdigest_ncentroids(void *T) { return ((tdigest_t *)T)->ncentroids; }
API int ref_tdigest_buffer_ncentroids(void *T) { return ((tdigest_t *)T)->buffer_ncentroids; }
API double ref_tdigest_total_weight(void *T) { return ((tdigest_t *)T)->total_weight; }
API double ref_tdigest_buffer_total_weight(void *T) { return ((tdigest_t *)T)->buffer_total_weight; }
API void ref_tdigest_get_centroids(void *T, double *out /* 2*ncentroids */) {
    tdigest_t *t = (tdigest_t *)T;
    memcpy(out, t->centroids, (size_t)t->ncentroids * sizeof(centroid_t));
}
/* tdigest.pyx:253-263 (__setstate__) */
API void ref_tdigest_set_state(void *T, const double *centroids, int n,
                               double total_weight, double min, double max) {
    tdigest_t *t = (tdigest_t *)T;
    t->total_weight = total_weight;
    t->min = min;
    t->max = max;
    if (n > 0) {
        memcpy(t->centroids, centroids, (size_t)n * sizeof(centroid_t));
        t->ncentroids = n;
    }
}

/* Multi-threaded CPU baseline, the dask map-reduce pattern of
 * tdigest.pyx:282-324: one digest per shard (thread), then merge in shard
 * order into a fresh digest. Returns the merged digest. */
typedef struct {
    const double *x, *w; double ws; long n; double c; tdigest_t *T;
} shard_job_t;
static void *shard_main(void *p) {
    shard_job_t *j = (shard_job_t *)p;
    j->T = tdigest_new(j->c);
    ref_tdigest_update(j->T, j->x, j->w, j->ws, j->n);
    return NULL;
}
API void *ref_tdigest_update_sharded(double compression, const double *x,
                                     const double *w, double w_scalar, long n,
                                     int nthreads) {
    pthread_t th[256]; shard_job_t jobs[256];
    int t; tdigest_t *out;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    for (t = 0; t < nthreads; t++) {
        long lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
        jobs[t].x = x + lo; jobs[t].w = w ? w + lo : NULL; jobs[t].ws = w_scalar;
        jobs[t].n = hi - lo; jobs[t].c = compression; jobs[t].T = NULL;
        pthread_create(&th[t], NULL, shard_main, &jobs[t]);
    }
    for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    out = tdigest_new(compression);
    for (t = 0; t < nthreads; t++) {
        tdigest_merge(out, jobs[t].T);
        tdigest_free(jobs[t].T);
    }
    return out;
}

/* -------------------------------------------------------------------- stats */
API void *ref_stats_new(void) { return stats_new(); }
API void ref_stats_free(void *T) { stats_free((stats_t *)T); }
API void ref_stats_add(void *T, double x, long long count) { stats_add((stats_t *)T, x, count); }
API void ref_stats_merge(void *T1, void *T2) { stats_merge((stats_t *)T1, (stats_t *)T2); }
/* stats_update_ndarray loop incl. homogeneous/first_value, stats_stubs.c:199-210 */
API void ref_stats_update(void *Tv, const double *x, const long long *count,
                          long long count_scalar, long n) {
    stats_t *T = (stats_t *)Tv; long i;
    for (i = 0; i < n; i++) {
        double value = x[i];
        if (T->count == 0) T->first_value = value;
        else if (T->homogeneous && T->first_value != value) T->homogeneous = false;
        stats_add(T, value, count ? count[i] : count_scalar);
    }
}
/* stats.pyx:77-91 */
API void ref_stats_getstate(void *Tv, long long *count, double *f7 /* sum,min,max,m2,m3,m4,first */,
                            int *homogeneous) {
    stats_t *T = (stats_t *)Tv;
    *count = T->count; f7[0] = T->sum; f7[1] = T->min; f7[2] = T->max;
    f7[3] = T->m2; f7[4] = T->m3; f7[5] = T->m4; f7[6] = T->first_value;
    *homogeneous = T->homogeneous;
}
API void ref_stats_setstate(void *Tv, long long count, const double *f7, int homogeneous) {
    stats_t *T = (stats_t *)Tv;
    T->count = count; T->sum = f7[0]; T->min = f7[1]; T->max = f7[2];
    T->m2 = f7[3]; T->m3 = f7[4]; T->m4 = f7[5]; T->first_value = f7[6];
    T->homogeneous = homogeneous;
}
API double ref_stats_mean(void *T) { return stats_mean((stats_t *)T); }
API double ref_stats_var(void *T, long ddof) { return stats_var((stats_t *)T, ddof); }
API double ref_stats_std(void *T, long ddof) { return stats_std((stats_t *)T, ddof); }
API double ref_stats_skew(void *T, int bias) { return stats_skew((stats_t *)T, bias); }
API double ref_stats_kurt(void *T, int fisher, int bias) { return stats_kurt((stats_t *)T, fisher, bias); }

/* ------------------------------------------------------------- space saving */
API void *ref_spsv_new(int capacity) { return spsv_int64_new(capacity); }
API void ref_spsv_free(void *T) { spsv_int64_free((spsv_int64_t *)T); }
API int ref_spsv_add(void *T, long long item, long long count) {
    return spsv_int64_add((spsv_int64_t *)T, item, count);
}
/* spsv_int64_update_ndarray loop, space_saving_stubs.c.in:422-435 */
API void ref_spsv_update(void *T, const long long *item, const long long *count,
                         long long count_scalar, long n) {
    long i;
    for (i = 0; i < n; i++)
        spsv_int64_add((spsv_int64_t *)T, item[i], count ? count[i] : count_scalar);
}
API int ref_spsv_merge(void *T1, void *T2) {
    return spsv_int64_merge((spsv_int64_t *)T1, (spsv_int64_t *)T2);
}
API int ref_spsv_set_state(void *T, const long long *counters /* n x 3 */, long n) {
    return spsv_int64_set_state((spsv_int64_t *)T, (counter_int64_t *)counters, n);
}
API long ref_spsv_size(void *T) { return ((spsv_int64_t *)T)->size; }
API long ref_spsv_capacity(void *T) { return ((spsv_int64_t *)T)->capacity; }
/* int64_counters walk, space_saving.pyx:368-386 */
API long ref_spsv_counters(void *Tv, long long *out /* k x 3 */, long k) {
    spsv_int64_t *T = (spsv_int64_t *)Tv;
    long n = T->size < k ? T->size : k, c; npy_intp i = T->head;
    for (c = 0; c < n; c++) {
        out[3 * c + 0] = T->list[i].counter.item;
        out[3 * c + 1] = T->list[i].counter.count;
        out[3 * c + 2] = T->list[i].counter.error;
        i = T->list[i].next;
    }
    return n;
}
