// FP64 / integer pipe throughput per SM on B200.  Build: nvcc -arch=sm_100a -O3 -o fp64_bench fp64_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024, 1) bench(double* out, int iters, unsigned long long* cycles) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 0.9999999;
    unsigned long long i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
    unsigned int m = threadIdx.x | 1;
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {           // 8 independent DFMA chains
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        } else if (MODE == 1) {    // DADD
            a0 += b; a1 += b; a2 += b; a3 += b; a4 += b; a5 += b; a6 += b; a7 += b;
        } else if (MODE == 2) {    // DMUL
            a0 *= b; a1 *= b; a2 *= b; a3 *= b; a4 *= b; a5 *= b; a6 *= b; a7 *= b;
        } else if (MODE == 3) {    // IMAD.WIDE.U32 with 64-bit accumulate
            asm volatile("mad.wide.u32 %0, %1, %1, %0;" : "+l"(i0) : "r"(m));
            asm volatile("mad.wide.u32 %0, %1, %1, %0;" : "+l"(i1) : "r"(m));
            asm volatile("mad.wide.u32 %0, %1, %1, %0;" : "+l"(i2) : "r"(m));
            asm volatile("mad.wide.u32 %0, %1, %1, %0;" : "+l"(i3) : "r"(m));
            asm volatile("mad.wide.u32 %0, %1, %1, %0;" : "+l"(i0) : "r"(m));
            asm volatile("mad.wide.u32 %0, %1, %1, %0;" : "+l"(i1) : "r"(m));
            asm volatile("mad.wide.u32 %0, %1, %1, %0;" : "+l"(i2) : "r"(m));
            asm volatile("mad.wide.u32 %0, %1, %1, %0;" : "+l"(i3) : "r"(m));
        }
    }
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + (double)(i0 + i1 + i2 + i3);
}

template <int MODE>
void run(const char* name, int sms, double* out, unsigned long long* cyc) {
    for (int threads : {128, 256, 512, 1024}) {
        const int iters = 4000;
        bench<MODE><<<sms, threads>>>(out, iters, cyc);
        bench<MODE><<<sms, threads>>>(out, iters, cyc);
        cudaDeviceSynchronize();
        unsigned long long h[256];
        cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
        double mean = 0;
        for (int i = 0; i < sms; ++i) mean += (double)h[i];
        mean /= sms;
        printf("%-22s threads=%4d  lane-ops/clk/SM = %7.2f\n", name, threads, (double)threads * iters * 8 / mean);
    }
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; unsigned long long* cyc;
    cudaMalloc(&out, sms * 1024 * 8); cudaMalloc(&cyc, sms * 8);
    run<0>("DFMA", sms, out, cyc);
    run<1>("DADD", sms, out, cyc);
    run<2>("DMUL", sms, out, cyc);
    run<3>("IMAD.WIDE.U32 (acc64)", sms, out, cyc);
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
