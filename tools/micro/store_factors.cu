// Microbenchmark: which property of a store kernel decides between 6.0 TB/s (K1-like persistent CTAs writing two arrays) and the
// 7.5 TB/s of torch's fill_ / cudaMemset on B200?  Variants of "write 16 GiB":
//   arrays: 1 (one 16 GiB buffer) or 2 (4 GiB + 12 GiB written in step, like dist + rgb)
//   grid:   persistent grid-stride (P CTAs) or one CTA per chunk (non-persistent, memory order)
//   policy: st.global.cs or default
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_factors store_factors.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <bool CS>
__device__ __forceinline__ void st4(float4* p, float4 v) { if (CS) __stcs(p, v); else *p = v; }

// chunk c = 16 KiB: ONE: bytes [16K*c, 16K*(c+1)) of a;  TWO: 4 KiB of a + 12 KiB of b
template <bool TWO, bool CS>
__global__ void __launch_bounds__(256) k(float4* a, float4* b, size_t nchunks)
{
    const float4 q = make_float4(1.f, 2.f, 3.f, 4.f);
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        if (TWO) {
            st4<CS>(a + c * 256 + threadIdx.x, q);
            float4* g = b + c * 768;
            st4<CS>(g + threadIdx.x, q); st4<CS>(g + 256 + threadIdx.x, q); st4<CS>(g + 512 + threadIdx.x, q);
        } else {
            float4* g = a + c * 1024;
            st4<CS>(g + threadIdx.x, q); st4<CS>(g + 256 + threadIdx.x, q); st4<CS>(g + 512 + threadIdx.x, q); st4<CS>(g + 768 + threadIdx.x, q);
        }
    }
}

template <class F>
static void timeit(const char* name, F launch, double bytes)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 2; i++) launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int i = 0; i < 4; i++) {
        cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-64s %7.3f ms = %6.0f GB/s %s\n", name, best, bytes / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
    fflush(stdout);
}

int main()
{
    const size_t total = (size_t)16 << 30, nchunks = total / 16384;
    float4 *a, *b;
    cudaMalloc(&a, total);
    b = a + (total / 4) / 16;      // TWO: a = first 4 GiB, b = the following 12 GiB
    char name[160];
    for (int grid : {148, 296, 592, 1184, 0}) {
        const unsigned g = grid ? (unsigned)grid : (unsigned)nchunks;
        snprintf(name, sizeof name, "one array,  cs,      %s", grid ? (snprintf(name + 100, 50, "%d persistent CTAs", grid), name + 100) : "one CTA per 16 KiB");
        timeit(name, [&] { k<false, true><<<g, 256>>>(a, b, nchunks); }, (double)total);
        snprintf(name, sizeof name, "one array,  default, %s", grid ? (snprintf(name + 100, 50, "%d persistent CTAs", grid), name + 100) : "one CTA per 16 KiB");
        timeit(name, [&] { k<false, false><<<g, 256>>>(a, b, nchunks); }, (double)total);
        snprintf(name, sizeof name, "two arrays, cs,      %s", grid ? (snprintf(name + 100, 50, "%d persistent CTAs", grid), name + 100) : "one CTA per 16 KiB");
        timeit(name, [&] { k<true, true><<<g, 256>>>(a, b, nchunks); }, (double)total);
        snprintf(name, sizeof name, "two arrays, default, %s", grid ? (snprintf(name + 100, 50, "%d persistent CTAs", grid), name + 100) : "one CTA per 16 KiB");
        timeit(name, [&] { k<true, false><<<g, 256>>>(a, b, nchunks); }, (double)total);
    }
    timeit("cudaMemsetAsync 16 GiB", [&] { cudaMemsetAsync(a, 0, total); }, (double)total);
    return 0;
}
