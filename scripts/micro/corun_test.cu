// Do two persistent single-warp-CTA kernels with different static shared memory sizes share the SMs when launched on
// two streams?  A: 11.4 KB / CTA, nA CTAs per SM; B: 8.6 KB / CTA, 24 CTAs per SM.  Each CTA spins for a fixed number
// of clock cycles; the elapsed time of both tells whether they ran next to each other or one after the other.
#include <cstdio>
#include <cuda_runtime.h>
template <int SMEM>
__global__ void __launch_bounds__(32, 1) spin(long long cycles, unsigned *sink, unsigned long long *first_last) {
    __shared__ unsigned char sm[SMEM];
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const long long c0 = clock64();
    unsigned acc = 0;
    while (clock64() - c0 < cycles) { sm[(threadIdx.x * 4 + acc) % SMEM] = static_cast<unsigned char>(acc); acc += sm[(acc * 7) % SMEM] + 1; }
    if (acc == 0xdeadbeef) sink[0] = acc;
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (threadIdx.x == 0) { atomicMin(&first_last[0], t0); atomicMax(&first_last[1], t1); }
}
int main(int argc, char **argv) {
    const int nA = argc > 1 ? atoi(argv[1]) : 9;
    const int carve = argc > 2 ? atoi(argv[2]) : -2;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int nsm = p.multiProcessorCount;
    cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    unsigned *sink; cudaMalloc(&sink, 4);
    unsigned long long *fl, h[4]; cudaMalloc(&fl, 32);
    if (carve > -2) {
        cudaFuncSetAttribute(spin<11392>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(spin<8576>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    }
    int oa, ob;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oa, spin<11392>, 32, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ob, spin<8576>, 32, 0);
    const long long cyc = 10000000;  // ~5 ms
    for (int rep = 0; rep < 2; rep++) {
        h[0] = ~0ull; h[1] = 0; h[2] = ~0ull; h[3] = 0;
        cudaMemcpy(fl, h, 32, cudaMemcpyHostToDevice);
        cudaEvent_t e0, e1, f; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreateWithFlags(&f, cudaEventDisableTiming);
        cudaDeviceSynchronize();
        cudaEventRecord(e0, s1);
        cudaEventRecord(f, s1);
        cudaStreamWaitEvent(s2, f, 0);
        spin<11392><<<nsm * nA, 32, 0, s1>>>(cyc, sink, fl);
        spin<8576><<<nsm * 10, 32, 0, s2>>>(cyc, sink, fl + 2);
        cudaEvent_t j; cudaEventCreateWithFlags(&j, cudaEventDisableTiming);
        cudaEventRecord(j, s2); cudaStreamWaitEvent(s1, j, 0);
        cudaEventRecord(e1, s1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h, fl, 32, cudaMemcpyDeviceToHost);
        printf("nA=%d carve=%d occA=%d occB=%d total %.2f ms; A [%.2f, %.2f] B [%.2f, %.2f] ms (err %s)\n", nA, carve, oa, ob, ms, 0.0,
               (h[1] - h[0]) * 1e-6, (double)(long long)(h[2] - h[0]) * 1e-6, (double)(long long)(h[3] - h[0]) * 1e-6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
