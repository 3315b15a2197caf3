// Checks sqrt_rn_inline and div_rn_inline (csrc/sqrt_rn.cuh) against __dsqrt_rn / __ddiv_rn on the GPU, bit for bit.  Square root: the arguments the mel kernel
// produces: x = r^2 + i^2 in fp64 of two float32 values (all exponents, denormals, zeros, infinities, NaNs), 2^33 pairs.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/sqrt_check tools/sqrt_check.cu && /tmp/sqrt_check
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "../lsm_speech_classifier_b200/csrc/sqrt_rn.cuh"

__device__ __forceinline__ uint32_t mix(uint64_t &s)
{
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    uint64_t z = s;
    z ^= z >> 33; z *= 0xff51afd7ed558ccdull; z ^= z >> 33;
    return (uint32_t)z;
}

__global__ void check(unsigned long long *bad, unsigned long long *first, int per_thread)
{
    uint64_t s = 0x9e3779b97f4a7c15ull * (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x + 1);
    unsigned long long nbad = 0;
    for (int k = 0; k < per_thread; ++k) {
        const uint32_t a = mix(s), b = mix(s);
        float r = __uint_as_float(a), i = __uint_as_float(b);           // every bit pattern: all exponents, denormals, inf, NaN
        if ((k & 15) == 1) i = 0.0f;                                    // pure real
        if ((k & 15) == 2) { r = 0.0f; i = (k & 16) ? 0.0f : i; }       // zero and pure imaginary
        if ((k & 15) == 3) i = r;                                       // equal parts
        if ((k & 15) == 4) i = __uint_as_float((a & 0xff800000u) | (b & 0x007fffffu));   // same exponent
        const double r64 = (double)r, i64 = (double)i;
        const double x = __dadd_rn(__dmul_rn(r64, r64), __dmul_rn(i64, i64));
        const double want = __dsqrt_rn(x), got = sqrt_rn_inline(x);
        const bool same = (__double_as_longlong(want) == __double_as_longlong(got)) || (want != want && got != got);
        if (!same) { if (!nbad) atomicMin(first, (unsigned long long)__double_as_longlong(x)); ++nbad; }
    }
    if (nbad) atomicAdd(bad, nbad);
}

// the division inside lsm_log10: x = any positive normal double, reduced to [sqrt(2)/2, sqrt(2)) as fdlibm does, f = x - 1, s = f / (2 + f)
__global__ void check_div(unsigned long long *bad, int per_thread)
{
    uint64_t s = 0xd1342543de82ef95ull * (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x + 1);
    unsigned long long nbad = 0;
    for (int k = 0; k < per_thread; ++k) {
        uint64_t bits = ((uint64_t)mix(s) << 32) | mix(s);
        if ((k & 7) == 1) bits &= 0xffffffff00000000ull;                 // short mantissas
        if ((k & 7) == 2) bits = (bits & 0xfff0000000000000ull) | ((bits & 0xff) << ((bits >> 8) % 45));   // a few bits set
        if ((k & 7) == 3) bits |= 0x000fffffff000000ull;                 // just below a power of two
        if ((k & 31) == 4) bits &= 0xfff0000000000000ull;                // exact powers of two: f = 0
        int hx = (int)(bits >> 32) & 0x000fffff;
        const int i = (hx + 0x95f64) & 0x100000;
        const double x = __hiloint2double(hx | (i ^ 0x3ff00000), (int)(uint32_t)bits);
        const double f = __dsub_rn(x, 1.0), b = __dadd_rn(2.0, f);
        const double want = __ddiv_rn(f, b), got = div_rn_inline(f, b);
        if (__double_as_longlong(want) != __double_as_longlong(got)) ++nbad;
    }
    if (nbad) atomicAdd(bad, nbad);
}

int main()
{
    unsigned long long *d, h[2] = {0, ~0ull};
    cudaMalloc(&d, 16);
    cudaMemcpy(d, h, 16, cudaMemcpyHostToDevice);
    const int blocks = 148 * 16, threads = 256, per_thread = 1 << 14;
    for (int rep = 0; rep < 14; ++rep) check<<<blocks, threads>>>(d, d + 1, per_thread + rep);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 2; }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("sqrt_rn_inline vs __dsqrt_rn: %.3g arguments, %llu differ", 14.0 * blocks * threads * per_thread, h[0]);
    if (h[0]) printf(" (smallest differing argument bits 0x%016llx)", h[1]);
    printf("\n");
    unsigned long long zero = 0, dbad = 0;
    cudaMemcpy(d, &zero, 8, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 14; ++rep) check_div<<<blocks, threads>>>(d, per_thread + rep);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 2; }
    cudaMemcpy(&dbad, d, 8, cudaMemcpyDeviceToHost);
    printf("div_rn_inline vs __ddiv_rn on lsm_log10's f / (2 + f): %.3g arguments, %llu differ\n", 14.0 * blocks * threads * per_thread, dbad);
    return (h[0] || dbad) ? 1 : 0;
}
