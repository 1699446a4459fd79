// Microbenchmark: how fast can the classify access pattern (warp = 128-float x-chunk, R+1 rows, marching K planes)
// stream a 1024^3 float field, with increasing amounts of per-voxel work?  Build: nvcc -arch=sm_100a -O3 -o sr stream_read.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define FULL 0xFFFFFFFFu
template <int R, int MODE, int MINB>
__global__ void __launch_bounds__(256, MINB) k(const float* __restrict__ d, unsigned* __restrict__ out, int nx, int ny, int nz, int K, unsigned ntiles, float iso)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned gw = blockIdx.x * 8 + (threadIdx.x >> 5), nw = gridDim.x * 8;
    const unsigned cpr = nx / 128, njb = (ny - 1 + R - 1) / R;
    for (unsigned tile = gw; tile < ntiles; tile += nw) {
        const unsigned xc = tile % cpr, t2 = tile / cpr, jb = t2 % njb, kb = t2 / njb;
        const float* lp = d + ((size_t)(kb * K) * ny + jb * R) * nx + xc * 128 + lane * 4;
        unsigned acc = 0, prev[R + 1];
#pragma unroll
        for (int r = 0; r <= R; r++) prev[r] = 0;
        const int nl = min(K, nz - 1 - (int)(kb * K));
#pragma unroll 1
        for (int kk = 0; kk <= nl; kk++) {
            float4 q[R + 1];
            float e[R + 1];
#pragma unroll
            for (int r = 0; r <= R; r++) {
                q[r] = __ldg((const float4*)(lp + (size_t)r * nx));
                if (MODE >= 1) e[r] = __ldg(lp + (size_t)r * nx + (lane == 31 ? 4 : 3));
            }
            unsigned cur[R + 1];
#pragma unroll
            for (int r = 0; r <= R; r++) {
                if (MODE == 0) { acc += __float_as_uint(q[r].x) ^ __float_as_uint(q[r].y) ^ __float_as_uint(q[r].z) ^ __float_as_uint(q[r].w); cur[r] = 0; }
                else {
                    unsigned s = (q[r].x > iso ? 1u : 0u) | (q[r].y > iso ? 2u : 0u) | (q[r].z > iso ? 4u : 0u) | (q[r].w > iso ? 8u : 0u);
                    unsigned nb = __shfl_down_sync(FULL, s, 1);
                    if (lane == 31) nb = e[r] > iso;
                    cur[r] = s | ((nb & 1u) << 4);
                }
            }
            if (MODE >= 2) {
                unsigned o = 0, a = 31;
#pragma unroll
                for (int r = 0; r <= R; r++) { o |= cur[r] | prev[r]; a &= cur[r] & prev[r]; }
                const bool quiet = (o == 0) || (a == 31);
                if (__all_sync(FULL, quiet)) { if (lane < R) out[((size_t)(kb * K + kk) * (ny - 1) + jb * R + lane) * cpr + xc] = 0; }
                else acc += o + a;
            } else {
#pragma unroll
                for (int r = 0; r <= R; r++) acc += cur[r];
            }
#pragma unroll
            for (int r = 0; r <= R; r++) prev[r] = cur[r];
            lp += (size_t)nx * ny;
        }
        if (acc == 0x12345678u) out[tile] = acc;
    }
}
template <int R, int MODE, int MINB>
float run(const float* d, unsigned* out, int n, int K, const char* name)
{
    const unsigned cpr = n / 128, njb = (n - 1 + R - 1) / R, nkb = (n - 1 + K - 1) / K;
    const unsigned ntiles = cpr * njb * nkb;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int it = 0; it < 5; it++) {
        cudaEventRecord(a);
        k<R, MODE, MINB><<<148 * 8, 256>>>(d, out, n, n, n, K, ntiles, 0.0f);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    printf("%-28s R=%d K=%d minb=%d : %.3f ms  %.0f GB/s (algorithmic 4 B/voxel)  err=%s\n", name, R, K, MINB, best, 4.0 * n * n * (double)n / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    return best;
}
__global__ void fill(float* d, size_t n) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = 1.0f + (float)(i & 1023) * 1e-3f; }
int main()
{
    const int n = 1024;
    const size_t N = (size_t)n * n * n;
    float* d; unsigned* out;
    cudaMalloc(&d, N * 4 + (16 * n + 64) * 4); cudaMalloc(&out, (N / 128 + 1024) * 4);
    fill<<<148 * 8, 256>>>(d, N + 16 * n + 64); cudaDeviceSynchronize();
    run<8, 0, 2>(d, out, n, 16, "loads + xor");
    run<8, 0, 4>(d, out, n, 16, "loads + xor");
    run<8, 1, 2>(d, out, n, 16, "loads + e + sign bits");
    run<8, 1, 3>(d, out, n, 16, "loads + e + sign bits");
    run<8, 2, 2>(d, out, n, 16, "+ quiet test + count store");
    run<8, 2, 3>(d, out, n, 16, "+ quiet test + count store");
    run<4, 2, 4>(d, out, n, 16, "+ quiet test + count store");
    run<8, 2, 2>(d, out, n, 32, "+ quiet test + count store");
    run<8, 2, 2>(d, out, n, 64, "+ quiet test + count store");
    run<16, 2, 2>(d, out, n, 16, "+ quiet test + count store");
    return 0;
}
