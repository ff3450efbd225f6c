// Micro-benchmark (not product code): how many bytes must one SM keep in flight to stream two
// f64 arrays at HBM speed?  Same tile -> warp assignment as k_bin_accumulate.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ double2 ldg2(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
// DEPTH tiles of 256 values (x and w) in flight per warp
template <int DEPTH>
__global__ void k_stream(const double* __restrict__ X, const double* __restrict__ W, long n, double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long gw = (long)blockIdx.x * nw + warp, tw = (long)gridDim.x * nw;
    const long ntiles = n / 256;
    double2 bx[DEPTH][4], bw[DEPTH][4];
    double acc = 0;
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
        long t = gw + d * tw;
        if (t < ntiles) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                bx[d][k] = ldg2(reinterpret_cast<const double2*>(X + t * 256) + k * 32 + lane);
                bw[d][k] = ldg2(reinterpret_cast<const double2*>(W + t * 256) + k * 32 + lane);
            }
        }
    }
    for (long t0 = gw; t0 < ntiles; t0 += (long)DEPTH * tw) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            long t = t0 + d * tw;
            if (t < ntiles) {
#pragma unroll
                for (int k = 0; k < 4; ++k) acc += bx[d][k].x * bw[d][k].x + bx[d][k].y * bw[d][k].y;
                long tn = t + (long)DEPTH * tw;
                if (tn < ntiles) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        bx[d][k] = ldg2(reinterpret_cast<const double2*>(X + tn * 256) + k * 32 + lane);
                        bw[d][k] = ldg2(reinterpret_cast<const double2*>(W + tn * 256) + k * 32 + lane);
                    }
                }
            }
        }
    }
    if (acc == 123.456) out[0] = acc;
}
template <int DEPTH>
void run(const double* X, const double* W, long n, double* out, int warps, int ctas_per_sm) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int grid = 148 * ctas_per_sm;
    for (int i = 0; i < 2; ++i) k_stream<DEPTH><<<grid, warps * 32>>>(X, W, n, out);
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) k_stream<DEPTH><<<grid, warps * 32>>>(X, W, n, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    printf("depth %d warps/cta %2d ctas/sm %d : %.3f ms  %.0f GB/s  (in flight/SM %d KB) %s\n", DEPTH, warps, ctas_per_sm, ms,
           16.0 * n / ms / 1e6, DEPTH * warps * ctas_per_sm * 4, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    long n = 500000000L;
    double *X, *W, *out;
    cudaMalloc(&X, n * 8); cudaMalloc(&W, n * 8); cudaMalloc(&out, 8);
    cudaMemset(X, 0, n * 8); cudaMemset(W, 0, n * 8);
    for (int w : {8, 12, 16, 32}) {
        run<1>(X, W, n, out, w, 1); run<2>(X, W, n, out, w, 1); run<4>(X, W, n, out, w, 1);
    }
    run<1>(X, W, n, out, 32, 2); run<2>(X, W, n, out, 16, 4);
    return 0;
}
