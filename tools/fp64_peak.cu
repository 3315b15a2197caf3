// Measured fp64 pipe ceiling on this GPU: independent DADD/DMUL (and DFMA) chains in registers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int ILP>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b)
{
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = a + threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) { v[i] = __dmul_rn(v[i], b); v[i] = __dadd_rn(v[i], a); }      // DMUL + DADD (what K1 issues)
            else { v[i] = __fma_rn(v[i], b, a); v[i] = __fma_rn(v[i], b, a); }            // DFMA + DFMA
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int ILP>
void run(const char *name, int warps_per_smsp)
{
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int threads = 128;                       // 4 warps = 1 per SMSP
    const int blocks = p.multiProcessorCount * warps_per_smsp;
    const int iters = 20000;
    double *out; cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, ILP><<<blocks, threads>>>(out, 100, 1.0, 0.999999);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE, ILP><<<blocks, threads>>>(out, iters, 1.0, 0.999999);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads * iters * ILP * 2;
    printf("%-12s ILP %d warps/SMSP %d: %.1f G fp64 instr-lanes/s (%.3f ms)\n", name, ILP, warps_per_smsp, ops / ms / 1e6, ms);
    cudaFree(out);
}

int main()
{
    run<0, 1>("DMUL+DADD", 1); run<0, 2>("DMUL+DADD", 1); run<0, 4>("DMUL+DADD", 1); run<0, 8>("DMUL+DADD", 1);
    run<0, 4>("DMUL+DADD", 2); run<0, 4>("DMUL+DADD", 4); run<0, 8>("DMUL+DADD", 4); run<0, 8>("DMUL+DADD", 8);
    run<1, 8>("DFMA", 4); run<1, 8>("DFMA", 8); run<1, 1>("DFMA", 1);
    return 0;
}
