// Shared-memory / L2 atomic throughput on one SM (B200).  Build: nvcc -arch=sm_100a -O3 -o atomics_bench atomics_bench.cu
// Each warp issues ITER warp-wide atomicAdd(+1) operations at pseudo-random word addresses inside a
// 128 KB table (shared memory, or a per-CTA slice of global memory that stays L2 resident).
// Prints lane-atomics per SM clock.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 10; }

// mode bit0: use return value; mode bit1: warps with (warp & 1) use the global table instead of shared;
// mode bit2: ALL warps use the global table; spread: address mask (0x7fff = 32K words, 0 = single address)
__global__ void __launch_bounds__(1024, 1) bench(uint32_t* gtab, int iters, int mode, uint32_t mask, unsigned long long* cycles, uint32_t* sink) {
    extern __shared__ uint32_t tab[];
    for (int i = threadIdx.x; i < 32768; i += blockDim.x) tab[i] = 0;
    uint32_t* mytab = gtab + (size_t)blockIdx.x * 32768;
    __syncthreads();
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x * 97u + 12345u, acc = 0;
    const bool use_ret = mode & 1;
    const bool global = (mode & 4) || ((mode & 2) && ((threadIdx.x >> 5) & 1));
    __syncthreads();
    const unsigned long long t0 = clock64();
    if (!global) {
        for (int i = 0; i < iters; ++i) {
            const uint32_t a = lcg(s) & mask;
            if (use_ret) acc += atomicAdd(&tab[a], 1u); else atomicAdd(&tab[a], 1u);
        }
    } else {
        for (int i = 0; i < iters; ++i) {
            const uint32_t a = lcg(s) & mask;
            if (use_ret) acc += atomicAdd(&mytab[a], 1u); else atomicAdd(&mytab[a], 1u);
        }
    }
    __syncthreads();
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0xdeadbeef) sink[0] = acc;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *gtab, *sink;
    unsigned long long* cyc;
    cudaMalloc(&gtab, (size_t)sms * 32768 * 4);
    cudaMemset(gtab, 0, (size_t)sms * 32768 * 4);
    cudaMalloc(&sink, 4);
    cudaMalloc(&cyc, sms * 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    const int iters = 2000;
    printf("sms=%d iters=%d (lane-atomics per SM clock; full chip running)\n", sms, iters);
    const char* names[] = {"shared  no-ret", "shared  ret   ", "mixed   no-ret", "mixed   ret   ", "global  no-ret", "global  ret   "};
    const int modes[] = {0, 1, 2, 3, 4, 5};
    for (uint32_t mask : {0x7fffu, 0x3ffu, 0x0u}) {
        for (int m = 0; m < 6; ++m) {
            for (int threads : {128, 256, 512, 1024}) {
                bench<<<sms, threads, 131072>>>(gtab, iters, modes[m], mask, cyc, sink);   // warm
                bench<<<sms, threads, 131072>>>(gtab, iters, modes[m], mask, cyc, sink);
                cudaDeviceSynchronize();
                unsigned long long h[256];
                cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
                double mean = 0;
                for (int i = 0; i < sms; ++i) mean += (double)h[i];
                mean /= sms;
                printf("mask=%05x %s threads=%4d  cycles=%9.0f  lane-atomics/clk/SM=%6.3f\n", mask, names[m], threads, mean,
                       (double)threads * iters / mean);
            }
        }
    }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
