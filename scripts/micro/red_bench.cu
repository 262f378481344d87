// Microbenchmark: cost of the 5 reductions a same-event accepted pair makes, by accumulator layout.
//   SoA: five separate arrays (count u64, cos, qo, qs, ql f64) indexed by the bin       (current layout)
//   AoS8: one 64-byte record per bin {count, cos, qo, qs, ql, pad x3}: the five REDs of a pair hit one line
//   AoS5: 40-byte records
// Each lane draws pseudo-random bins (spread addresses, like the drain's 32 distinct bins per round).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o red_bench red_bench.cu && ./red_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void red_f64(double *p, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void red_u64(unsigned long long *p) {
    asm volatile("red.global.add.u64 [%0], 1;" ::"l"(p) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(32, 18) k(unsigned long long *cnt, double *f, unsigned nbins, int iters) {
    unsigned x = (blockIdx.x * 32u + threadIdx.x) * 2654435761u + 12345u;
    double v = 1.0 + threadIdx.x * 1e-3;
    for (int i = 0; i < iters; i++) {
        x = x * 1664525u + 1013904223u;
        const unsigned bin = (x >> 8) % nbins;
        if (MODE == 0) {
            red_u64(cnt + bin);
            red_f64(f + bin, v); red_f64(f + nbins + bin, v); red_f64(f + 2ull * nbins + bin, v); red_f64(f + 3ull * nbins + bin, v);
        } else if (MODE == 1) {
            double *r = f + 8ull * bin;
            red_u64(reinterpret_cast<unsigned long long *>(r));
            red_f64(r + 1, v); red_f64(r + 2, v); red_f64(r + 3, v); red_f64(r + 4, v);
        } else if (MODE == 2) {
            double *r = f + 5ull * bin;
            red_u64(reinterpret_cast<unsigned long long *>(r));
            red_f64(r + 1, v); red_f64(r + 2, v); red_f64(r + 3, v); red_f64(r + 4, v);
        } else if (MODE == 3) {  // one RED per pair (what a mixed-event pair does)
            red_u64(cnt + bin);
        } else if (MODE == 4) {  // AoS4 f64 (32-byte record) + separate count
            double *r = f + 4ull * bin;
            red_u64(cnt + bin);
            red_f64(r, v); red_f64(r + 1, v); red_f64(r + 2, v); red_f64(r + 3, v);
        } else if (MODE == 5) {
            // TRANSPOSED, 64-byte records of five f64 (the count kept as a double): the 160 reductions of a round of
            // 32 pairs are issued as 5 instructions in which consecutive lanes take consecutive components of the
            // same record: 6.4 records (lines) per instruction instead of 32
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const int s = k * 32 + threadIdx.x, pr = s / 5, c = s - 5 * pr;
                const unsigned b = __shfl_sync(0xffffffffu, bin, pr);
                red_f64(f + 8ull * b + c, v);
            }
        } else if (MODE == 6) {
            // TRANSPOSED, 32-byte records of four f64 (one sector) + the count as a separate spread u64 RED
            red_u64(cnt + bin);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int s = k * 32 + threadIdx.x, pr = s >> 2, c = s & 3;
                const unsigned b = __shfl_sync(0xffffffffu, bin, pr);
                red_f64(f + 4ull * b + c, v);
            }
        } else if (MODE == 7) {
            // TRANSPOSED, 64-byte records, 8 lanes per record (5 live + 3 idle lanes): 4 records per instruction, 8 instructions
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int pr = k * 4 + (threadIdx.x >> 3), c = threadIdx.x & 7;
                const unsigned b = __shfl_sync(0xffffffffu, bin, pr);
                if (c < 5) red_f64(f + 8ull * b + c, v);
            }
        }
    }
}

template <int MODE>
void run(const char *name, unsigned long long *cnt, double *f, unsigned nbins, int sms, double ghz) {
    const int iters = 20000, grid = sms * 18;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<grid, 32>>>(cnt, f, nbins, 100);
    cudaEventRecord(a);
    k<MODE><<<grid, 32>>>(cnt, f, nbins, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double lanes = double(grid) * 32 * iters * (MODE == 3 ? 1 : 5);
    printf("%-28s %8.3f ms  %6.3f RED lanes/clk/SM  (%.3f cycles per lane)\n", name, ms, lanes / (ms * 1e-3) / (ghz * 1e9) / sms,
           (ms * 1e-3) * (ghz * 1e9) * sms / lanes);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const double ghz = p.clockRate * 1e-6;
    const unsigned nbins = 4 * 41 * 41 * 41;
    unsigned long long *cnt;
    double *f;
    cudaMalloc(&cnt, 8ull * nbins);
    cudaMalloc(&f, 8ull * 8 * nbins);
    cudaMemset(cnt, 0, 8ull * nbins);
    cudaMemset(f, 0, 8ull * 8 * nbins);
    printf("%s, %d SMs, %.3f GHz (nominal), %u bins\n", p.name, p.multiProcessorCount, ghz, nbins);
    for (int rep = 0; rep < 2; rep++) {
        run<0>("SoA 5 arrays (current)", cnt, f, nbins, p.multiProcessorCount, ghz);
        run<1>("AoS 64-byte records", cnt, f, nbins, p.multiProcessorCount, ghz);
        run<2>("AoS 40-byte records", cnt, f, nbins, p.multiProcessorCount, ghz);
        run<4>("AoS 32-byte f64 + count", cnt, f, nbins, p.multiProcessorCount, ghz);
        run<3>("1 RED per pair", cnt, f, nbins, p.multiProcessorCount, ghz);
        run<5>("transposed 64-byte records", cnt, f, nbins, p.multiProcessorCount, ghz);
        run<6>("transposed 32-byte + count", cnt, f, nbins, p.multiProcessorCount, ghz);
        run<7>("transposed 64-byte, 8 lanes", cnt, f, nbins, p.multiProcessorCount, ghz);
    }
    return 0;
}
