// fp64_micro.cu -- DFMA latency/throughput on B200 as a function of warps per SM sub-partition
// and independent chains per thread (what the filter kernels' occupancy/ILP trade-off rests on).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a profiles/fp64_micro.cu -o /tmp/fp64_micro
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void chain_kernel(double *out, int iters, double a, double b) {
    double v[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) v[k] = threadIdx.x * 1e-3 + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) v[k] = fma(v[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += v[k];
    if (s == 12345.678) out[0] = s;
}

// mixed: DFMA chain interleaved with integer work (IMAD) per FMA, like the loop overhead
template <int CHAINS>
__global__ void mixed_kernel(double *out, int iters, double a, double b, int c) {
    double v[CHAINS];
    int w[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) { v[k] = threadIdx.x * 1e-3 + k; w[k] = threadIdx.x + k; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) { v[k] = fma(v[k], a, b); w[k] = w[k] * c + i; }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += v[k] + w[k];
    if (s == 12345.678) out[0] = s;
}

template <int CHAINS>
static void run(int warps_per_sm, bool mixed) {
    int dev = 0, sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double *out;
    cudaMalloc(&out, 8);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = warps_per_sm * 32;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mixed) mixed_kernel<CHAINS><<<sms, threads>>>(out, iters, 1.0000001, 1e-9, 3);
        else chain_kernel<CHAINS><<<sms, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    const double fma_total = (double)sms * threads * CHAINS * iters;
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%s warps/SM %2d chains %d: %7.3f ms  %6.2f TFLOP/s  %5.1f DFMA/clk/SM  %5.1f cycles per chain step\n",
           mixed ? "mixed" : "dfma ", warps_per_sm, CHAINS, ms, 2 * fma_total / (ms * 1e-3) / 1e12,
           fma_total / sms / cycles, cycles / iters);
    cudaFree(out);
}

int main() {
    for (int w : {4, 8, 16, 20, 24, 32}) {
        run<1>(w, false); run<2>(w, false); run<4>(w, false); run<8>(w, false);
    }
    for (int w : {16, 24}) { run<1>(w, true); run<2>(w, true); run<4>(w, true); }
    return 0;
}
