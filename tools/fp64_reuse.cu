// Does the order of the 13 DFMAs (operand reuse across consecutive instructions) lift the 3-register-operand DFMA rate?
#include <cstdio>
#include <cuda_runtime.h>
#define FMA(d, a, b, c) asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c))
template <int VAR>
__global__ void __launch_bounds__(128) casc(double *out, int iters, double c1, double c2, double c3, double c4, double na1, double na2, double x0)
{
    double xp = 0, p1 = 0, q1 = 0, p2 = 0, q2 = 0, p3 = 0, q3 = 0, p4 = 0, q4 = 0, acc = 0;
    const double t = threadIdx.x * 1e-9;
    c1 += t; c2 += t; c3 += t; c4 += t; na1 += t; na2 -= t;
    double x = x0 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 8
        for (int u = 0; u < 8; ++u) {
            const double x_ = x; x = -x;
            double t1, t2, t3, t4, n1, n2, n3, n4;
            if (VAR == 0) {          // coefficient-major order: na2 x4, na1 x4 back to back (shared operand in the same slot)
                FMA(t1, c1, xp, x_); FMA(t2, c2, q1, p1); FMA(t3, c3, q2, p2); FMA(t4, c4, q3, p3);
                FMA(t1, na2, q1, t1); FMA(t2, na2, q2, t2); FMA(t3, na2, q3, t3); FMA(t4, na2, q4, t4);
                FMA(n1, na1, p1, t1); FMA(n2, na1, p2, t2); FMA(n3, na1, p3, t3); FMA(n4, na1, p4, t4);
            } else if (VAR == 1) {   // stage-major order
                FMA(t1, c1, xp, x_); FMA(t1, na2, q1, t1); FMA(n1, na1, p1, t1);
                FMA(t2, c2, q1, p1); FMA(t2, na2, q2, t2); FMA(n2, na1, p2, t2);
                FMA(t3, c3, q2, p2); FMA(t3, na2, q3, t3); FMA(n3, na1, p3, t3);
                FMA(t4, c4, q3, p3); FMA(t4, na2, q4, t4); FMA(n4, na1, p4, t4);
            } else {                 // state operand shared: (c_k, q_{k-1}) then (na2, q_{k-1})... pairs sharing q / p
                FMA(t1, c1, xp, x_); FMA(t2, c2, q1, p1); FMA(t1, na2, q1, t1);
                FMA(t3, c3, q2, p2); FMA(t2, na2, q2, t2); FMA(t4, c4, q3, p3); FMA(t3, na2, q3, t3); FMA(t4, na2, q4, t4);
                FMA(n1, na1, p1, t1); FMA(n2, na1, p2, t2); FMA(n3, na1, p3, t3); FMA(n4, na1, p4, t4);
            }
            xp = x_;
            q1 = p1; p1 = n1; q2 = p2; p2 = n2; q3 = p3; p3 = n3; q4 = p4; p4 = n4;
            FMA(acc, n4, n4, acc);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int VAR>
void run(const char *name)
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *out; cudaMalloc(&out, 8 * 148 * 16 * 128);
    for (int w = 1; w <= 6; w += (w < 2 ? 1 : 2)) {
        const int blocks = p.multiProcessorCount * w, iters = 4000;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        casc<VAR><<<blocks, 128>>>(out, 10, -0.9, -0.8, -0.7, -0.6, 1.9, -0.95, 1e-3);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        casc<VAR><<<blocks, 128>>>(out, iters, -0.9, -0.8, -0.7, -0.6, 1.9, -0.95, 1e-3);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%s, %d warps/SMSP: %.1f G lane-ops/s\n", name, w, (double)blocks * 128 * iters * 8 * 13 / ms / 1e6);
    }
    cudaFree(out);
}
int main()
{
    run<0>("coefficient-major");
    run<1>("stage-major      ");
    run<2>("mixed            ");
    return 0;
}
