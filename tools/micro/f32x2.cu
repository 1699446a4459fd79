// Throughput of scalar FADD/FMUL versus packed FADD2/FMUL2 (add/mul.rn.f32x2) on sm_100a, no FMA contraction.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o f32x2 f32x2.cu && ./f32x2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b)
{ unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b)
{ unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int PACKED>
__global__ void k(float* out, int iters, float s, float t)
{
    float a[16];
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 0.001f + i;
    if (PACKED) {
        unsigned long long v[8], ss, tt;
        for (int i = 0; i < 8; i++) asm("mov.b64 %0, {%1, %2};" : "=l"(v[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
        asm("mov.b64 %0, {%1, %1};" : "=l"(ss) : "f"(s));
        asm("mov.b64 %0, {%1, %1};" : "=l"(tt) : "f"(t));
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = add2(mul2(v[i], ss), tt);
        }
        for (int i = 0; i < 8; i++) asm("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(v[i]));
    } else {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = a[i] * s + t;
        }
    }
    float r = 0;
    for (int i = 0; i < 16; i++) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main()
{
    float* d;
    cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int packed = 0; packed < 2; packed++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (packed) k<1><<<148 * 8, 256>>>(d, iters, 0.999f, 0.001f); else k<0><<<148 * 8, 256>>>(d, iters, 0.999f, 0.001f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double ops = 148.0 * 8 * 256 * (double)iters * 16 * 2;   // one mul + one add per element
            if (rep) printf("%s: %.3f ms  %.3e lane-op/s\n", packed ? "packed f32x2" : "scalar      ", ms, ops / (ms * 1e-3));
        }
    }
    return 0;
}
