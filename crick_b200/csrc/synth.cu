// crick_b200/csrc/synth.cu -- synthetic, counter-based input generators for
// bench.py and the GPU tests (NOT part of the hot path).  Element i of the
// stream depends only on (seed, i), so any sharding of a stream across GPUs
// reproduces the same values.
#include "kernels.h"

#include <cuda_fp16.h>

namespace crick {

// ---- safe casts to float64 for CUDA-array inputs that are not float64 (the reference accepts
// anything np.can_cast(dtype, 'f8', 'safe') allows, tdigest.pyx:298-302) -------------------------
template <typename T>
__device__ __forceinline__ double to_f64(T v) { return (double)v; }
template <>
__device__ __forceinline__ double to_f64<__half>(__half v) { return (double)__half2float(v); }

template <typename T>
__global__ void k_cast_f64(const T* __restrict__ src, long n, double* __restrict__ dst) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = to_f64<T>(src[i]);
}

template <typename T>
static cudaError_t cast_launch(const void* src, long n, double* dst, cudaStream_t st) {
    long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    count_launch();
    k_cast_f64<T><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const T*>(src), n, dst);
    return cudaGetLastError();
}

cudaError_t launch_cast_f64(const void* src, int kind, int itemsize, long n, double* dst, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    if (kind == 'f') {
        if (itemsize == 8) return cast_launch<double>(src, n, dst, st);
        if (itemsize == 4) return cast_launch<float>(src, n, dst, st);
        if (itemsize == 2) return cast_launch<__half>(src, n, dst, st);
    } else if (kind == 'i') {
        if (itemsize == 8) return cast_launch<long long>(src, n, dst, st);
        if (itemsize == 4) return cast_launch<int>(src, n, dst, st);
        if (itemsize == 2) return cast_launch<short>(src, n, dst, st);
        if (itemsize == 1) return cast_launch<signed char>(src, n, dst, st);
    } else if (kind == 'u') {
        if (itemsize == 8) return cast_launch<unsigned long long>(src, n, dst, st);
        if (itemsize == 4) return cast_launch<unsigned>(src, n, dst, st);
        if (itemsize == 2) return cast_launch<unsigned short>(src, n, dst, st);
        if (itemsize == 1) return cast_launch<unsigned char>(src, n, dst, st);
    } else if (kind == 'b' && itemsize == 1) {
        return cast_launch<unsigned char>(src, n, dst, st);
    }
    return cudaErrorInvalidValue;
}

// w finite and > 0 for every element?  (tdigest.pyx:304-305)
__global__ void k_check_weights(const double* __restrict__ w, long n, int* flag) {
    const long stride = (long)gridDim.x * blockDim.x;
    bool bad = false;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v = w[i];
        bad |= !(v > 0.0 && v <= DBL_MAX);
    }
    if (__any_sync(CRICK_FULL, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}
cudaError_t launch_check_weights(const double* w, long n, int* flag, cudaStream_t st) {
    long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    count_launch();
    k_check_weights<<<(unsigned)blocks, 256, 0, st>>>(w, n, flag);
    return cudaGetLastError();
}

__device__ __forceinline__ unsigned long long sm64(unsigned long long z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(unsigned long long r) {  // (0, 1)
    return ((double)(r >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

// kind 0: x = exp(N(0,1)) (lognormal), w = U(0.5, 1.5)
// kind 1: x = U(0, 1),                 w = U(0.5, 1.5)
// kind 2: x = N(0, 1),                 w = U(0.5, 1.5)
// kind 3: x = Zipf(1.1)-like heavy-tailed positive integers stored as f64 (inverse-CDF of the
//         continuous Pareto with the same tail, floored), w = 1
// kind 4: the same integers stored as int64 bit patterns (BASELINE config 5's int64 stream)
__global__ void k_synth(double* __restrict__ x, double* __restrict__ w, long n,
                        unsigned long long seed, long offset, int kind) {
    long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        unsigned long long g = (unsigned long long)(i + offset);
        unsigned long long r0 = sm64(seed ^ sm64(3 * g));
        unsigned long long r1 = sm64(seed ^ sm64(3 * g + 1));
        unsigned long long r2 = sm64(seed ^ sm64(3 * g + 2));
        double u0 = u01(r0), u1 = u01(r1), u2 = u01(r2);
        double v;
        if (kind == 1) v = u0;
        else if (kind == 3 || kind == 4) {
            double p = pow(u0, -10.0);           // Pareto(alpha = 0.1): tail ~ k^-1.1
            v = p > 9.0e18 ? 9.0e18 : floor(p);
        } else {
            double nrm = sqrt(-2.0 * log(u0)) * cospi(2.0 * u1);
            v = (kind == 0) ? exp(nrm) : nrm;
        }
        x[i] = kind == 4 ? __longlong_as_double((long long)v) : v;
        if (w) w[i] = (kind == 3 || kind == 4) ? 1.0 : 0.5 + u2;
    }
}

cudaError_t launch_synth(double* x, double* w, long n, unsigned long long seed, long offset,
                         int kind, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    count_launch();
    k_synth<<<(unsigned)blocks, 256, 0, st>>>(x, w, n, seed, offset, kind);
    return cudaGetLastError();
}

}  // namespace crick
