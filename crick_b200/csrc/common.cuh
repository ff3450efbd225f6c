// crick_b200/csrc/common.cuh -- shared device/host helpers for libcrick_cuda.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

#define CRICK_FULL 0xffffffffu
#define CRICK_PI 3.141592653589793238462643383279502884   /* NPY_PI   */
#define CRICK_PI_2 1.570796326794896619231321691639751442 /* NPY_PI_2 */

namespace crick {

// kernels launched by this library since load (crick_launch_count)
void count_launch();

// ---------------------------------------------------------------------------
// Per-digest header as it lives in HBM (one per digest, 48 B).  Mirrors the
// scalar fields of the reference's tdigest_t (tdigest_stubs.c:22-47).
struct DevHdr {
    double vmin, vmax;  // min / max
    double total;       // total_weight
    double btotal;      // buffer_total_weight
    int n;              // ncentroids
    int bn;             // buffer_ncentroids
    int pad0, pad1;
};

// A batch of digests sharing one compression: digest g owns hdr[g],
// cen[g*cap .. +cap), buf[g*bufcap .. +bufcap), table[g*cap .. +cap).
// Centroids are AoS double2 {mean, weight} exactly like centroid_t
// (tdigest_stubs.c:16-19) so pickles/centroids() are a plain memcpy.
struct BatchView {
    DevHdr* hdr;
    double2* cen;
    double2* buf;
    double2* table;  // query scratch (mean, cumulative midpoint), tdigest_stubs.c:301-316
    int cap;
    int bufcap;
    double comp;
};

// ---------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ double shfl_d(double v, int src) {
    return __shfl_sync(CRICK_FULL, v, src);
}
__device__ __forceinline__ double shfl_up_d(double v, int d) {
    return __shfl_up_sync(CRICK_FULL, v, d);
}
__device__ __forceinline__ double shfl_xor_d(double v, int m) {
    return __shfl_xor_sync(CRICK_FULL, v, m);
}

// Streaming 128-bit load that does not pollute L1 (input is read exactly once).
__device__ __forceinline__ double2 ldg_stream2(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
                 : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double ldg_stream1(const double* p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}

// Order-preserving map double -> uint64 (total order; -0.0 < +0.0, callers
// canonicalise zeros first when ties must tie).
__host__ __device__ __forceinline__ uint64_t ordered_key(double x) {
#ifdef __CUDA_ARCH__
    uint64_t b = (uint64_t)__double_as_longlong(x);
#else
    uint64_t b; memcpy(&b, &x, 8);
#endif
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ bool is_finite_d(double x) {
    // exponent field != all ones
    return ((__double2hiint(x) >> 20) & 0x7ff) != 0x7ff;
}

}  // namespace crick
