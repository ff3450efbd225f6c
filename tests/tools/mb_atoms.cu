// Micro-benchmark (sm_100a): throughput of shared-memory integer atomics without a return value
// (red.shared.add.u32 -> ATOMS.ADD) against a plain LDS / IADD / STS read-modify-write, with the
// access patterns the unweighted k_bin_accumulate variant needs.  Cycles per warp instruction per
// SM, 16 warps per CTA, one CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/tools/mb_atoms.bin tests/tools/mb_atoms.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ROWS = 201, ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(unsigned* out, long long* cyc) {
    extern __shared__ unsigned tab[];   // [ROWS][32] (u32) or [ROWS][32] u64
    for (int i = threadIdx.x; i < ROWS * 64; i += blockDim.x) tab[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned s = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    const unsigned base = (unsigned)__cvta_generic_to_shared(tab);
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
        s = s * 1664525u + 1013904223u;
        unsigned row = ((s >> 8) * (unsigned)ROWS) >> 24;          // 0 .. ROWS-1, pseudo-random per lane
        unsigned a;
        if (MODE == 0) a = base + (row * 32 + lane) * 4;           // lane = column: conflict-free
        else if (MODE == 1) a = base + row * 4;                    // one counter per bucket: collisions
        else a = base + (row * 32 + lane) * 4;
        if (MODE <= 1) {
            asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory");
        } else if (MODE == 2) {                                    // plain RMW (racy across warps: only timing)
            unsigned v;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
            v += 1;
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
        } else if (MODE == 3) {                                    // atom with return value
            unsigned v;
            asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(v) : "r"(a) : "memory");
            s ^= v & 1;
        } else if (MODE == 4) {                                    // 64-bit add of a variable, lane = column
            unsigned long long v = ((unsigned long long)s << 13) | 0x4330000000000000ull;
            a = base + (row * 32 + lane) * 8;
            asm volatile("red.shared.add.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
        } else if (MODE == 5) {                                    // 64-bit add, one accumulator per bucket
            unsigned long long v = ((unsigned long long)s << 13) | 0x4330000000000000ull;
            a = base + row * 8;
            asm volatile("red.shared.add.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
        } else if (MODE == 6) {                                    // 32-bit add of a variable, lane = column
            asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(s >> 20) : "memory");
        } else if (MODE == 7) {                                    // 64-bit add, 4 replicas per bucket
            unsigned long long v = ((unsigned long long)s << 13) | 0x4330000000000000ull;
            a = base + (row * 4 + (lane & 3)) * 8;
            asm volatile("red.shared.add.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
        } else {                                                   // LDS.64 + DADD + STS.64 lane = column (racy: timing only)
            a = base + (row * 32 + lane) * 8;
            double v;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
            v += __longlong_as_double(((long long)s << 13) | 0x3ff0000000000000ll);
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    unsigned acc = 0;
    for (int i = threadIdx.x; i < ROWS * 64; i += blockDim.x) acc += tab[i];
    atomicAdd(out, acc + s);
}

template <int MODE>
void run(const char* name, int warps) {
    unsigned* out; long long* cyc;
    cudaMalloc(&out, 4); cudaMalloc(&cyc, 148 * 8);
    cudaMemset(out, 0, 4);
    size_t smem = ROWS * 64 * 4;
    k<MODE><<<148, warps * 32, smem>>>(out, cyc);
    k<MODE><<<148, warps * 32, smem>>>(out, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("%-44s warps=%2d  %.2f cycles per warp-instruction per SM (%.2f per warp)\n", name, warps,
           avg / ((double)ITERS * warps), avg / ITERS);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 16}) {
        run<0>("red.shared.add.u32 lane=column (no conflict)", w);
        run<1>("red.shared.add.u32 one counter per bucket", w);
        run<2>("LDS.32 + IADD + STS.32 lane=column", w);
        run<3>("atom.shared.add.u32 (return) lane=column", w);
        run<6>("red.shared.add.u32 variable lane=column", w);
        run<4>("red.shared.add.u64 variable lane=column", w);
        run<5>("red.shared.add.u64 one per bucket", w);
        run<7>("red.shared.add.u64 4 replicas per bucket", w);
        run<8>("LDS.64 + DADD + STS.64 lane=column", w);
    }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
