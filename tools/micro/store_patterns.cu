// Microbenchmark: write-only HBM bandwidth of the ORDER in which a sampling kernel could write an n^3 voxel field
// (dist: 4 B/voxel, rgb: 12 B/voxel, x fastest; a warp stores 512 B of dist + 1536 B of rgb per row tile of 128 voxels).
// cudaMemset reaches 7.5 TB/s on this box, K1's order (a CTA owns a row and walks z: consecutive stores 4 MB / 12 MB apart)
// 6.5 TB/s.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_patterns store_patterns.cu
//   pattern 0: CTA (8 warps) = row y, walks z over [z0, z1) (K1 today; ZS = z segments per column)
//   pattern 1: CTA = (z-group of G slices, y range): for y { for z in group { store row } }   (keeps per-lane sign words: G = 8 / 32)
//   pattern 2: grid-stride over rows in memory order (the memset order)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void store_row(float* dist, float* rgb, size_t row, int nx, unsigned warp, unsigned lane, float v)
{
    // row = z*ny + y ; this warp's tile = 128 voxels
    const size_t vb = row * (size_t)nx + warp * 128u;
    const float4 q = make_float4(v, v, v, v);
    __stcs(reinterpret_cast<float4*>(dist + vb) + lane, q);
    float4* g = reinterpret_cast<float4*>(rgb + vb * 3);
    __stcs(g + lane, q);
    __stcs(g + lane + 32, q);
    __stcs(g + lane + 64, q);
}

// nx = 1024 -> 8 tiles per row = the 8 warps of a CTA
__global__ void __launch_bounds__(256) p0(float* dist, float* rgb, int nx, int ny, int nz, int zs, unsigned nwork)
{
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    for (unsigned w = blockIdx.x; w < nwork; w += gridDim.x) {
        const unsigned seg = w / (unsigned)ny, y = w % (unsigned)ny;
        const int z0 = (int)((long long)seg * nz / zs), z1 = (int)((long long)(seg + 1) * nz / zs);
        for (int z = z0; z < z1; z++) store_row(dist, rgb, (size_t)z * ny + y, nx, warp, lane, (float)z);
    }
}

__global__ void __launch_bounds__(256) p1(float* dist, float* rgb, int nx, int ny, int nz, int G, int ysegs, unsigned nwork)
{
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    for (unsigned w = blockIdx.x; w < nwork; w += gridDim.x) {
        const unsigned zg = w / (unsigned)ysegs, ys = w % (unsigned)ysegs;
        const int y0 = (int)((long long)ys * ny / ysegs), y1 = (int)((long long)(ys + 1) * ny / ysegs);
        const int z0 = (int)zg * G, z1 = min(nz, z0 + G);
        for (int y = y0; y < y1; y++)
            for (int z = z0; z < z1; z++) store_row(dist, rgb, (size_t)z * ny + y, nx, warp, lane, (float)z);
    }
}

__global__ void __launch_bounds__(256) p2(float* dist, float* rgb, int nx, size_t nrows)
{
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    for (size_t r = blockIdx.x; r < nrows; r += gridDim.x) store_row(dist, rgb, r, nx, warp, lane, 1.0f);
}

// pattern 3: items handed out IN ORDER (non-persistent CTAs, or persistent CTAs drawing from an atomic counter):
//   item i = (z-group i / ny, row i % ny): the CTA stores the row's G slices, so the rows in flight form G compact windows
__global__ void __launch_bounds__(256) p3(float* dist, float* rgb, int nx, int ny, int nz, int G, unsigned nitems, unsigned* counter)
{
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    __shared__ unsigned s_item;
    for (;;) {
        unsigned w;
        if (counter) {
            if (threadIdx.x == 0) s_item = atomicAdd(counter, 1u);
            __syncthreads();
            w = s_item;
            __syncthreads();
        } else w = blockIdx.x;
        if (w >= nitems) return;
        const unsigned zg = w / (unsigned)ny, y = w % (unsigned)ny;
        const int z0 = (int)zg * G, z1 = min(nz, z0 + G);
        for (int z = z0; z < z1; z++) store_row(dist, rgb, (size_t)z * ny + y, nx, warp, lane, (float)z);
        if (!counter) return;
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
        cudaEventRecord(a);
        launch();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-58s %7.3f ms = %6.0f GB/s %s\n", name, best, bytes / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
    fflush(stdout);
}

int main(int argc, char** argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 1024;
    const int nx = n, ny = n, nz = n;
    if (nx != 1024) { printf("nx must be 1024 (8 tiles = 8 warps)\n"); return 1; }
    const size_t nv = (size_t)nx * ny * nz;
    float *dist, *rgb;
    cudaMalloc(&dist, nv * 4);
    cudaMalloc(&rgb, nv * 12);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double bytes = (double)nv * 16;
    char name[128];
    timeit("cudaMemset dist+rgb", [&] { cudaMemsetAsync(dist, 0, nv * 4); cudaMemsetAsync(rgb, 0, nv * 12); }, bytes);
    if (argc > 2 && argv[2][0] == 'd') {   // in-order hand-out
        unsigned* counter;
        cudaMalloc(&counter, 4);
        for (int G : {1, 4, 8, 32, 128, 1024}) {
            const unsigned nitems = (unsigned)(((nz + G - 1) / G) * ny);
            snprintf(name, sizeof name, "p3 in order, G=%d, one CTA per item (%u)", G, nitems);
            timeit(name, [&] { p3<<<nitems, 256>>>(dist, rgb, nx, ny, nz, G, nitems, nullptr); }, bytes);
            for (int bps : {4, 8}) {
                snprintf(name, sizeof name, "p3 in order, G=%d, %d persistent CTAs + atomic counter", G, sms * bps);
                timeit(name, [&] { cudaMemsetAsync(counter, 0, 4); p3<<<sms * bps, 256>>>(dist, rgb, nx, ny, nz, G, nitems, counter); }, bytes);
            }
        }
        return 0;
    }
    if (argc > 2) {   // sweep of the number of CTAs (concurrent writers)
        for (int grid : {37, 74, 111, 148, 185, 222, 296, 444, 592, 1184}) {
            snprintf(name, sizeof name, "p2 memory order, %d CTAs", grid);
            timeit(name, [&] { p2<<<grid, 256>>>(dist, rgb, nx, (size_t)ny * nz); }, bytes);
            snprintf(name, sizeof name, "p0 CTA=row walks z, zsplit 1, %d CTAs", grid);
            timeit(name, [&] { p0<<<grid, 256>>>(dist, rgb, nx, ny, nz, 1, (unsigned)ny); }, bytes);
            snprintf(name, sizeof name, "p0 CTA=row walks z, zsplit 4, %d CTAs", grid);
            timeit(name, [&] { p0<<<grid, 256>>>(dist, rgb, nx, ny, nz, 4, (unsigned)ny * 4u); }, bytes);
        }
        return 0;
    }
    for (int bps : {4, 8}) {
        snprintf(name, sizeof name, "p2 memory order, grid %d/SM", bps);
        timeit(name, [&] { p2<<<sms * bps, 256>>>(dist, rgb, nx, (size_t)ny * nz); }, bytes);
    }
    for (int zs : {1, 2, 4, 8}) {
        for (int bps : {4, 8}) {
            const unsigned nwork = (unsigned)(ny * zs);
            const unsigned grid = nwork < (unsigned)(sms * bps) ? nwork : (unsigned)(sms * bps);
            snprintf(name, sizeof name, "p0 CTA=row walks z, zsplit %d, grid<=%d/SM (%u CTAs)", zs, bps, grid);
            timeit(name, [&] { p0<<<grid, 256>>>(dist, rgb, nx, ny, nz, zs, nwork); }, bytes);
        }
    }
    for (int G : {1, 8, 32}) {
        for (int ysegs : {1, 4, 19, 74}) {
            const unsigned nwork = (unsigned)(((nz + G - 1) / G) * ysegs);
            for (int bps : {4}) {
                const unsigned grid = nwork < (unsigned)(sms * bps) ? nwork : (unsigned)(sms * bps);
                snprintf(name, sizeof name, "p1 CTA=(%d slices, 1/%d of y) walks y, %u items, %u CTAs", G, ysegs, nwork, grid);
                timeit(name, [&] { p1<<<grid, 256>>>(dist, rgb, nx, ny, nz, G, ysegs, nwork); }, bytes);
            }
        }
    }
    return 0;
}
