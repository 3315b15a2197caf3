// DFMA throughput of the speculative gammatone cascade (13 DFMA per sample, 4 skewed chains) as a function of where the
// coefficient operands live: per-thread registers (lane = channel) or uniform/constant operands (lane = utterance).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o fp64_cascade fp64_cascade.cu && ./fp64_cascade
#include <cstdio>
#include <cuda_runtime.h>
template <bool REGCOEF>
__global__ void __launch_bounds__(128) casc(double *out, int iters, double c1, double c2, double c3, double c4, double na1, double na2, double x0)
{
    double xp = 0, p1 = 0, q1 = 0, p2 = 0, q2 = 0, p3 = 0, q3 = 0, p4 = 0, q4 = 0, acc = 0;
    if (REGCOEF) {
        const double t = threadIdx.x * 1e-9;
        c1 += t; c2 += t; c3 += t; c4 += t; na1 += t; na2 -= t;
    }
    double x = x0 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 8
        for (int u = 0; u < 8; ++u) {
            const double x_ = x; x = -x;
            const double n1 = fma(na1, p1, fma(na2, q1, fma(c1, xp, x_)));
            const double n2 = fma(na1, p2, fma(na2, q2, fma(c2, q1, p1)));
            const double n3 = fma(na1, p3, fma(na2, q3, fma(c3, q2, p2)));
            const double n4 = fma(na1, p4, fma(na2, q4, fma(c4, q3, p3)));
            xp = x_;
            q1 = p1; p1 = n1; q2 = p2; p2 = n2; q3 = p3; p3 = n3; q4 = p4; p4 = n4;
            acc = fma(n4, n4, acc);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <bool REGCOEF>
void run(const char *name)
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *out; cudaMalloc(&out, 8 * 148 * 16 * 128);
    for (int w = 1; w <= 8; ++w) {
        const int blocks = p.multiProcessorCount * w, iters = 4000;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        casc<REGCOEF><<<blocks, 128>>>(out, 10, -0.9, -0.8, -0.7, -0.6, 1.9, -0.95, 1e-3);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        casc<REGCOEF><<<blocks, 128>>>(out, iters, -0.9, -0.8, -0.7, -0.6, 1.9, -0.95, 1e-3);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%s, %d warps/SMSP: %.1f G lane-ops/s (%.3f ms)\n", name, w, (double)blocks * 128 * iters * 8 * 13 / ms / 1e6, ms);
    }
    cudaFree(out);
}
int main()
{
    run<false>("cascade, uniform coefficients   ");
    run<true>("cascade, per-thread coefficients");
    return 0;
}
