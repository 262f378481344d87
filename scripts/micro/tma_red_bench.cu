// Microbenchmark: the reductions of a same-event accepted pair through the TMA unit instead of the LSU.
//   cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [bin record], [record staged in shared memory], bytes
// One lane = one pair per round: it writes its record {cos, q_out, q_side, q_long (, count as a double ...)} to its own
// shared-memory slot and issues ONE bulk reduction (16-byte multiples) to a pseudo-random bin record, instead of 4-5
// red.global.add.f64 with 32 spread addresses each.  DEPTH rounds are kept in flight (rotating staging buffers,
// cp.async.bulk.wait_group.read before a buffer is overwritten).
// Also: the plain-RED numbers of red_bench.cu with only part of the SMs busy (is the ceiling on the SM or the L2 side?).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_red_bench tma_red_bench.cu && ./tma_red_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void red_f64(double *p, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void red_u64(unsigned long long *p) {
    asm volatile("red.global.add.u64 [%0], 1;" ::"l"(p) : "memory");
}

template <int BYTES>
__device__ __forceinline__ void bulk_red_f64(double *gdst, unsigned ssrc) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "n"(BYTES)
                 : "memory");
}

// BYTES per record (32: four sums, the count goes as a plain RED; 48 / 64: count as a double inside the record)
template <int BYTES, int DEPTH, bool COUNT_RED>
__global__ void __launch_bounds__(32, 18) k_tma(unsigned long long *cnt, double *f, unsigned nbins, int iters) {
    __shared__ __align__(128) unsigned char stage[DEPTH][32][BYTES];
    unsigned x = (blockIdx.x * 32u + threadIdx.x) * 2654435761u + 12345u;
    double v = 1.0 + threadIdx.x * 1e-3;
    constexpr int STRIDE = BYTES / 8;  // doubles per global record
    for (int i = 0; i < iters; i++) {
        x = x * 1664525u + 1013904223u;
        const unsigned bin = (x >> 8) % nbins;
        const int slot = i % DEPTH;
        // the bulk operations that read this slot DEPTH rounds ago must have finished reading it
        asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");
        double *s = reinterpret_cast<double *>(stage[slot][threadIdx.x]);
#pragma unroll
        for (int q = 0; q < BYTES / 8; q++) s[q] = v + q;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bulk_red_f64<BYTES>(f + static_cast<unsigned long long>(STRIDE) * bin, static_cast<unsigned>(__cvta_generic_to_shared(s)));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (COUNT_RED) red_u64(cnt + bin);
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// one elected lane issues the 32 records of the round as ONE... no: records go to 32 different bins, so 32 operations;
// this variant lets lane 0 issue all 32 (is the cost per issuing thread or per operation?)
template <int BYTES, int DEPTH>
__global__ void __launch_bounds__(32, 18) k_tma_lane0(double *f, unsigned nbins, int iters) {
    __shared__ __align__(128) unsigned char stage[DEPTH][32][BYTES];
    __shared__ unsigned bins[DEPTH][32];
    unsigned x = (blockIdx.x * 32u + threadIdx.x) * 2654435761u + 12345u;
    double v = 1.0 + threadIdx.x * 1e-3;
    constexpr int STRIDE = BYTES / 8;
    for (int i = 0; i < iters; i++) {
        x = x * 1664525u + 1013904223u;
        const unsigned bin = (x >> 8) % nbins;
        const int slot = i % DEPTH;
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");
        __syncwarp();
        double *s = reinterpret_cast<double *>(stage[slot][threadIdx.x]);
#pragma unroll
        for (int q = 0; q < BYTES / 8; q++) s[q] = v + q;
        bins[slot][threadIdx.x] = bin;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (threadIdx.x == 0) {
            for (int l = 0; l < 32; l++)
                bulk_red_f64<BYTES>(f + static_cast<unsigned long long>(STRIDE) * bins[slot][l],
                                    static_cast<unsigned>(__cvta_generic_to_shared(stage[slot][l])));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// 18 warps per CTA and (through a large dynamic shared-memory request) one CTA per SM: a grid of k CTAs keeps exactly k SMs busy
template <int MODE>
__global__ void __launch_bounds__(576, 1) k_red(unsigned long long *cnt, double *f, unsigned nbins, int iters) {
    unsigned x = (blockIdx.x * 576u + threadIdx.x) * 2654435761u + 12345u;
    double v = 1.0 + threadIdx.x * 1e-3;
    for (int i = 0; i < iters; i++) {
        x = x * 1664525u + 1013904223u;
        const unsigned bin = (x >> 8) % nbins;
        if (MODE == 0) {
            red_u64(cnt + bin);
            red_f64(f + bin, v); red_f64(f + nbins + bin, v); red_f64(f + 2ull * nbins + bin, v); red_f64(f + 3ull * nbins + bin, v);
        } else {
            red_u64(cnt + bin);
        }
    }
}

struct Timer {
    cudaEvent_t a, b;
    Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    void start() { cudaEventRecord(a); }
    float stop() { cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms; }
};

int g_sms;
double g_ghz;

void report(const char *name, float ms, double records, int sms_used, int values_per_record) {
    const double clk = ms * 1e-3 * g_ghz * 1e9;
    cudaError_t e = cudaGetLastError();
    printf("%-44s %8.3f ms  %6.3f records/clk/SM  %6.3f values/clk/SM  (%d SMs busy)%s%s\n", name, ms, records / clk / sms_used,
           records * values_per_record / clk / sms_used, sms_used, e == cudaSuccess ? "" : "  ERROR: ", e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int BYTES, int DEPTH, bool COUNT_RED>
void run_tma(const char *name, unsigned long long *cnt, double *f, unsigned nbins, int warps_per_sm = 18) {
    const int iters = 20000, grid = g_sms * warps_per_sm;
    Timer t;
    k_tma<BYTES, DEPTH, COUNT_RED><<<grid, 32>>>(cnt, f, nbins, 100);
    t.start();
    k_tma<BYTES, DEPTH, COUNT_RED><<<grid, 32>>>(cnt, f, nbins, iters);
    const float ms = t.stop();
    report(name, ms, double(grid) * 32 * iters, g_sms, BYTES / 8 + (COUNT_RED ? 1 : 0));
}

template <int BYTES, int DEPTH>
void run_tma_lane0(const char *name, double *f, unsigned nbins) {
    const int iters = 5000, grid = g_sms * 18;
    Timer t;
    k_tma_lane0<BYTES, DEPTH><<<grid, 32>>>(f, nbins, 100);
    t.start();
    k_tma_lane0<BYTES, DEPTH><<<grid, 32>>>(f, nbins, iters);
    const float ms = t.stop();
    report(name, ms, double(grid) * 32 * iters, g_sms, BYTES / 8);
}

template <int MODE>
void run_red(const char *name, unsigned long long *cnt, double *f, unsigned nbins, int sms_used) {
    const int iters = 20000, grid = sms_used;
    const size_t smem = 150 * 1024;
    cudaFuncSetAttribute(k_red<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    Timer t;
    k_red<MODE><<<grid, 576, smem>>>(cnt, f, nbins, 100);
    t.start();
    k_red<MODE><<<grid, 576, smem>>>(cnt, f, nbins, iters);
    const float ms = t.stop();
    report(name, ms, double(grid) * 576 * iters, sms_used, MODE == 0 ? 5 : 1);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    g_ghz = p.clockRate * 1e-6;
    g_sms = p.multiProcessorCount;
    const unsigned nbins = 4 * 41 * 41 * 41;
    unsigned long long *cnt;
    double *f;
    cudaMalloc(&cnt, 8ull * nbins);
    cudaMalloc(&f, 8ull * 8 * nbins);
    cudaMemset(cnt, 0, 8ull * nbins);
    cudaMemset(f, 0, 8ull * 8 * nbins);
    printf("%s, %d SMs, %.3f GHz (nominal), %u bins; per-SM rates are per BUSY SM\n", p.name, g_sms, g_ghz, nbins);
    for (int rep = 0; rep < 2; rep++) {
        run_red<0>("RED x5 SoA, 18 warps on every SM", cnt, f, nbins, g_sms);
        run_red<0>("RED x5 SoA, 18 warps on half of the SMs", cnt, f, nbins, g_sms / 2);
        run_red<0>("RED x5 SoA, 18 warps on a quarter of the SMs", cnt, f, nbins, g_sms / 4);
        run_red<1>("RED x1, 18 warps on every SM", cnt, f, nbins, g_sms);
        run_tma<32, 1, false>("bulk add.f64 32 B, depth 1", cnt, f, nbins);
        run_tma<32, 2, false>("bulk add.f64 32 B, depth 2", cnt, f, nbins);
        run_tma<32, 4, false>("bulk add.f64 32 B, depth 4", cnt, f, nbins);
        run_tma<32, 2, true>("bulk add.f64 32 B + count RED, depth 2", cnt, f, nbins);
        run_tma<48, 2, false>("bulk add.f64 48 B, depth 2", cnt, f, nbins);
        run_tma<64, 2, false>("bulk add.f64 64 B, depth 2", cnt, f, nbins);
        run_tma<64, 4, false>("bulk add.f64 64 B, depth 4", cnt, f, nbins);
        run_tma<32, 2, false>("bulk add.f64 32 B, depth 2, 9 warps/SM", cnt, f, nbins, 9);
        run_tma<32, 2, false>("bulk add.f64 32 B, depth 2, 4 warps/SM", cnt, f, nbins, 4);
        run_tma_lane0<32, 2>("bulk add.f64 32 B, lane 0 issues, depth 2", f, nbins);
    }
    return 0;
}
