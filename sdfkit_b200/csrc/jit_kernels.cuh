// jit_kernels.cuh -- kernels instantiated per SdfExpr by NVRTC (sm_100a).
//
// This text is appended, at sdfk_sdf_compile() time, after csrc/sdfk_prelude.h and the lowered
//     SK_FN sk_float4 sdf_eval(sk_float3 p) { <body> }
// and compiled with --fmad=false --prec-div=true --prec-sqrt=true --ftz=false so that every float
// operation is the IEEE binary32 operation the reference's CPU path performs.
// It must not include any header (NVRTC gets no include path).
//
//   sdfk_k_sample        K1  Voxels.SampleSdf + ClipToBounds      SdfKit/Voxels.cs:72-125,133-167
//   sdfk_k_eval          K6  the Sdf delegate on a point batch    SdfKit/SdfExpr.cs:240-271
//   sdfk_k_render        K5  RayMarcher.Render                    SdfKit/RayMarcher.cs:95-204
//   sdfk_k_render_depth      RayMarcher.RenderDepth               SdfKit/RayMarcher.cs:69-93

struct sdfk_fastdiv { unsigned mul, sh1, sh2; };   // q = n / d for 32-bit n (host-computed magic)

static __device__ __forceinline__ unsigned sdfk_div(unsigned n, sdfk_fastdiv d)
{
    unsigned t = __umulhi(d.mul, n);
    return (t + ((n - t) >> d.sh1)) >> d.sh2;
}

struct sdfk_sample_params {
    float m0, m1, m2;          // min + 0.5*delta                               (Voxels.cs:81)
    float dx, dy, dz;          // delta = (max-min)/n                           (Voxels.cs:32-34)
    float clip_value;          // Size.X / NX                                   (Voxels.cs:139)
    int clip;                  // ClipToBounds fused as a predicate
    int nx, ny, nz;            // dimensions of the WHOLE grid
    int z_begin;               // first z slice held by this slab
    int nzl;                   // slices in this slab
    unsigned tiles_per_row;    // ceil(nx / 128)
    unsigned ncol;             // columns = tiles_per_row * ny
    unsigned zsplit;           // z-segments per column (> 1 only when there are fewer columns than warps)
    unsigned nwork;            // ncol * zsplit
    sdfk_fastdiv div_tpr, div_ncol;
    float sign_iso;            // iso value of the sign planes (see below)
};

// the work decomposition of sdfk_k_sample_dist8, whose warps own PAIRS of tiles (256 x of one y).  (A separate kernel
// parameter: appended to sdfk_sample_params it changed the code generated for the other two sampling kernels, +13 % on K1d.)
struct sdfk_sample8_params {
    unsigned ncol8, zsplit8, nwork8;
    sdfk_fastdiv div_tpr8, div_ncol8;
};

#define SDFK_SAMPLE_WARPS 8

// cache policy of the voxel / sign-block stores (the field is write-once): 0 = st.global.cs (evict-first), 1 = default
// (write-back), 2 = st.global.cg, 3 = st.global.wt
#ifndef SDFK_STORE_POLICY
#define SDFK_STORE_POLICY 0
#endif
#if SDFK_STORE_POLICY == 0
#define SDFK_ST(p, v) __stcs(p, v)
#elif SDFK_STORE_POLICY == 1
#define SDFK_ST(p, v) (*(p) = (v))
#elif SDFK_STORE_POLICY == 2
#define SDFK_ST(p, v) __stcg(p, v)
#else
#define SDFK_ST(p, v) __stwt(p, v)
#endif

// Sign blocks: a by-product of sampling that lets marching cubes find the active cells without re-reading the
// distance field (1 bit per voxel instead of 4 bytes).  A warp walks its column (128 x of one y) in z, so every lane
// keeps the signs of its 4 voxels over 8 consecutive z slices in ONE register, four such registers make a uint4, and
// the warp stores 32 uint4 = 512 contiguous bytes per 32 slices:
//     signs[((y*tiles_per_row + xc)*nzg + zl/32)*32 + L]   (uint4; nzg = ceil(nzl/32))
//     word (zl/8)%4, bit 4*(zl%8) + k   <->   voxel (xc*128 + 4L + k, y, zl) has value > iso
// (strict, Cell.cs:221-228; evaluated on the value actually stored, i.e. after ClipToBounds).  Bits of voxels beyond
// the row / slab are unspecified / zero; the consumer (mc_classify_signs, mc_kernels.cu) masks the cells that do not
// exist.  z segments of a split column are 32-aligned, so a uint4 has exactly one writer.
static __device__ __forceinline__ unsigned sdfk_gt_mask(float a, float b)   // 0xFFFFFFFF if a > b (false for NaN) else 0: one FSET
{
    unsigned m;
    asm("set.gt.u32.f32 %0, %1, %2;" : "=r"(m) : "f"(a), "f"(b));
    return m;
}

static __device__ __forceinline__ unsigned sdfk_sign_nibble(const float* d, float iso)
{
    const unsigned t1 = sdfk_gt_mask(d[1], iso) & 2u;
    const unsigned t2 = (sdfk_gt_mask(d[0], iso) & 1u) | t1;
    const unsigned t3 = (sdfk_gt_mask(d[2], iso) & 4u) | t2;
    return (sdfk_gt_mask(d[3], iso) & 8u) | t3;
}

static __device__ __forceinline__ void sdfk_zsegment(const sdfk_sample_params& P, unsigned seg, int& zl0, int& zl1)
{
    const long long a = ((long long)seg * P.nzl) / P.zsplit, b = ((long long)(seg + 1) * P.nzl) / P.zsplit;
    zl0 = seg == 0u ? 0 : min(P.nzl, (int)((a + 31) & ~31ll));
    zl1 = seg + 1u == P.zsplit ? P.nzl : min(P.nzl, (int)((b + 31) & ~31ll));
}

// Device layout (DESIGN.md "data layout"): x fastest.  dist[(zl*ny + y)*nx + x], rgb[((zl*ny + y)*nx + x)*3 + c],
// zl = z - z_begin.  A warp owns a COLUMN -- 128 consecutive x of one y -- and walks it in z: everything
// sdf_eval derives from p.x and p.y alone is loop-invariant and hoisted out of the z loop by the compiler (for
// axis-separable trees such as RepeatXY that is nearly all of the arithmetic, including every IEEE division).
// Per z step every lane evaluates 4 consecutive voxels; the 4 distances leave as one float4 (512 B per warp
// store) and the 12 colour floats are transposed through a warp-private shared-memory stage so that the
// 1536 B of colours also leave as three fully coalesced 512 B float4 stores.  Neighbouring warps hold
// neighbouring columns, so at any moment the grid is writing a few contiguous planes.  Streaming
// (evict-first) stores: the field is write-once.
extern "C" __global__ void __launch_bounds__(SDFK_SAMPLE_WARPS * 32)
sdfk_k_sample(const sdfk_sample_params P, float* __restrict__ dist, float* __restrict__ rgb, uint4* __restrict__ signs)
{
    __shared__ float4 stage[SDFK_SAMPLE_WARPS][96];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned warp = threadIdx.x >> 5;
    float4* const st = stage[warp];
    const bool vec = (P.nx & 3) == 0;
    const size_t plane = (size_t)P.nx * (size_t)P.ny;

    for (unsigned work = blockIdx.x * SDFK_SAMPLE_WARPS + warp; work < P.nwork; work += gridDim.x * SDFK_SAMPLE_WARPS) {
        const unsigned seg = sdfk_div(work, P.div_ncol);
        const unsigned col = work - seg * P.ncol;
        const unsigned uy = sdfk_div(col, P.div_tpr);
        const unsigned xc = col - uy * P.tiles_per_row;
        const int iy = (int)uy;
        int zl0, zl1;
        sdfk_zsegment(P, seg, zl0, zl1);
        const int x0 = (int)(xc * 128u + lane * 4u);
        const float fx0 = (float)x0;                               // exact; fx0 + k is exact below 2^24
        const float py = P.m1 + (float)iy * P.dy;                  // p = min' + i*delta, one mul + one add (Voxels.cs:104-106)
        const bool ywall = P.clip && (iy == 0 || iy == P.ny - 1);
        float px[4];
        bool xywall[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            px[k] = P.m0 + (fx0 + (float)k) * P.dx;
            xywall[k] = ywall || (P.clip && (x0 + k == 0 || x0 + k == P.nx - 1));
        }
        const int nvalid = min(128, P.nx - (int)(xc * 128u));
        const int nq = (nvalid * 3) >> 2;                          // float4s of colour per tile on the vector path
        size_t vbase = ((size_t)zl0 * P.ny + uy) * (size_t)P.nx + (size_t)(xc * 128u);   // first voxel of the tile
        uint4* const scol = signs + ((size_t)uy * P.tiles_per_row + xc) * (size_t)((P.nzl + 31) >> 5) * 32u + lane;   // this lane's sign words

        for (int zg0 = zl0; zg0 < zl1; zg0 += 32) {                 // zl0 is a multiple of 32: one uint4 of signs per lane and 32 slices
        uint4 sw = make_uint4(0u, 0u, 0u, 0u);
        for (int zb0 = zg0; zb0 < min(zg0 + 32, zl1); zb0 += 8) {   // one 32-bit sign word per lane and 8 slices
        const int zend = min(zb0 + 8, zl1);
        unsigned sacc = 0u, ssh = 0u;
        for (int zl = zb0; zl < zend; zl++, vbase += plane, ssh += 4u) {
            const int iz = zl + P.z_begin;
            const float pz = P.m2 + (float)iz * P.dz;
            const bool zwall = P.clip && (iz == 0 || iz == P.nz - 1);
            float d[4];
            float c[12];
            sk_float4 r4[4];
            sdf_eval_grid(px, py, pz, r4);                          // the lane's 4 voxels of this row at once: shared y/z work and range guards
#pragma unroll
            for (int k = 0; k < 4; k++) {
                d[k] = (zwall || xywall[k]) ? P.clip_value : r4[k].w;
                c[3 * k + 0] = r4[k].x;
                c[3 * k + 1] = r4[k].y;
                c[3 * k + 2] = r4[k].z;
            }
            sacc |= sdfk_sign_nibble(d, P.sign_iso) << ssh;
            if (vec) {
                if (x0 < P.nx) SDFK_ST(reinterpret_cast<float4*>(dist + vbase) + lane, make_float4(d[0], d[1], d[2], d[3]));
                st[lane * 3 + 0] = make_float4(c[0], c[1], c[2], c[3]);
                st[lane * 3 + 1] = make_float4(c[4], c[5], c[6], c[7]);
                st[lane * 3 + 2] = make_float4(c[8], c[9], c[10], c[11]);
                __syncwarp();
                float4* const g = reinterpret_cast<float4*>(rgb + vbase * 3);
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const int q = (int)lane + 32 * k;
                    if (q < nq) SDFK_ST(g + q, st[q]);
                }
                __syncwarp();
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (x0 + k < P.nx) {
                        const size_t v = vbase + lane * 4u + k;
                        dist[v] = d[k];
                        rgb[v * 3 + 0] = c[3 * k + 0];
                        rgb[v * 3 + 1] = c[3 * k + 1];
                        rgb[v * 3 + 2] = c[3 * k + 2];
                    }
                }
            }
        }
        const int sq = (zb0 >> 3) & 3;
        if (sq == 0) sw.x = sacc; else if (sq == 1) sw.y = sacc; else if (sq == 2) sw.z = sacc; else sw.w = sacc;
        }
        if (signs) SDFK_ST(scol + (size_t)(zg0 >> 5) * 32u, sw);     // evict-first like the voxel stream (a second cache policy in the stream costs 3 %)
        }
    }
}

// One block of up to 8 z slices of a column for sdfk_k_sample_dist: evaluates, stores the distances, returns the lane's sign
// word.  WALLS = false: no slice of the block lies on a z wall and the row is not a wall row.
template <bool WALLS>
static __device__ __forceinline__ unsigned sdfk_sample_dist_block(const sdfk_sample_params& P, float* __restrict__ dist, size_t& vbase,
                                                                  size_t plane, int zb0, int zend, const float* px, float py,
                                                                  const unsigned* keep, const unsigned* setb, bool rowwall, bool vec,
                                                                  int x0, unsigned lane)
{
    unsigned sacc = 0u, ssh = 0u;
    for (int zl = zb0; zl < zend; zl++, vbase += plane, ssh += 4u) {
        const int iz = zl + P.z_begin;
        float d[4];
        if (WALLS && (rowwall || (P.clip && (iz == 0 || iz == P.nz - 1)))) {      // warp-uniform
            d[0] = d[1] = d[2] = d[3] = P.clip_value;
        } else {
            const float pz = P.m2 + (float)iz * P.dz;
            sk_float4 r4[4];
            sdf_eval_grid(px, py, pz, r4);                          // the lane's 4 voxels of this row at once: shared y/z work and range guards
#pragma unroll
            for (int k = 0; k < 4; k++)
                d[k] = __uint_as_float((__float_as_uint(r4[k].w) & keep[k]) | setb[k]);
        }
        sacc |= sdfk_sign_nibble(d, P.sign_iso) << ssh;
        if (vec) {
            if (x0 < P.nx) SDFK_ST(reinterpret_cast<float4*>(dist + vbase) + lane, make_float4(d[0], d[1], d[2], d[3]));
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (x0 + k < P.nx) {
                    const size_t v = vbase + lane * 4u + k;
                    dist[v] = d[k];
                }
            }
        }
    }
    return sacc;
}

// Distance-only variant of K1 for Sdf.ToMesh (SdfKit/Sdf.cs:59-63), where the caller never sees the voxels: 4 B/voxel
// instead of 16; the colours marching cubes needs (two corners per created vertex) are evaluated afterwards by
// sdfk_k_vertex_colors.  Same traversal and arithmetic as sdfk_k_sample.
extern "C" __global__ void __launch_bounds__(SDFK_SAMPLE_WARPS * 32)
sdfk_k_sample_dist(const sdfk_sample_params P, float* __restrict__ dist, uint4* __restrict__ signs)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned warp = threadIdx.x >> 5;
    const bool vec = (P.nx & 3) == 0;
    const size_t plane = (size_t)P.nx * (size_t)P.ny;

    for (unsigned work = blockIdx.x * SDFK_SAMPLE_WARPS + warp; work < P.nwork; work += gridDim.x * SDFK_SAMPLE_WARPS) {
        const unsigned seg = sdfk_div(work, P.div_ncol);
        const unsigned col = work - seg * P.ncol;
        const unsigned uy = sdfk_div(col, P.div_tpr);
        const unsigned xc = col - uy * P.tiles_per_row;
        const int iy = (int)uy;
        int zl0, zl1;
        sdfk_zsegment(P, seg, zl0, zl1);
        const int x0 = (int)(xc * 128u + lane * 4u);
        const float fx0 = (float)x0;                               // exact; fx0 + k is exact below 2^24
        const float py = P.m1 + (float)iy * P.dy;                  // p = min' + i*delta, one mul + one add (Voxels.cs:104-106)
        // ClipToBounds without per-voxel branches: a wall row (y) or wall slice (z) skips the evaluation altogether, the two
        // x walls are bit masks on the result: d = (d & keep) | setb
        const bool rowwall = P.clip && (iy == 0 || iy == P.ny - 1);
        float px[4];
        unsigned keep[4], setb[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            px[k] = P.m0 + (fx0 + (float)k) * P.dx;
            const bool w = P.clip && (x0 + k == 0 || x0 + k == P.nx - 1);
            keep[k] = w ? 0u : 0xFFFFFFFFu;
            setb[k] = w ? __float_as_uint(P.clip_value) : 0u;
        }
        size_t vbase = ((size_t)zl0 * P.ny + uy) * (size_t)P.nx + (size_t)(xc * 128u);   // first voxel of the tile
        uint4* const scol = signs + ((size_t)uy * P.tiles_per_row + xc) * (size_t)((P.nzl + 31) >> 5) * 32u + lane;   // this lane's sign words

        for (int zg0 = zl0; zg0 < zl1; zg0 += 32) {                 // zl0 is a multiple of 32: one uint4 of signs per lane and 32 slices
        uint4 sw = make_uint4(0u, 0u, 0u, 0u);
        for (int zb0 = zg0; zb0 < min(zg0 + 32, zl1); zb0 += 8) {   // one 32-bit sign word per lane and 8 slices
        const int zend = min(zb0 + 8, zl1);
        // only the first / last 8-slice block of the grid (z walls) and the two wall rows need the ClipToBounds logic per slice:
        // every other block runs the lean instance of the loop (13 instructions less per 128 voxels of an issue-bound kernel).
        // (Measured and dropped: a fully unrolled 8-slice block -- slower for every scene, 2x for large bodies; moving the
        // x-wall masks into the wall instance -- no gain.)
        const bool walls = rowwall || (P.clip && (zb0 + P.z_begin == 0 || zend + P.z_begin == P.nz));
        const unsigned sacc = walls ? sdfk_sample_dist_block<true>(P, dist, vbase, plane, zb0, zend, px, py, keep, setb, rowwall, vec, x0, lane)
                                    : sdfk_sample_dist_block<false>(P, dist, vbase, plane, zb0, zend, px, py, keep, setb, rowwall, vec, x0, lane);
        const int sq = (zb0 >> 3) & 3;
        if (sq == 0) sw.x = sacc; else if (sq == 1) sw.y = sacc; else if (sq == 2) sw.z = sacc; else sw.w = sacc;
        }
        if (signs) SDFK_ST(scol + (size_t)(zg0 >> 5) * 32u, sw);     // evict-first like the voxel stream (a second cache policy in the stream costs 3 %)
        }
    }
}


// K1d with 8 voxels per lane: a warp owns a PAIR of neighbouring tiles (256 consecutive x of one y).  The distance-only
// sampler is bound by instruction issue, not by HBM, and a third of its instructions per slice do not depend on how many
// voxels the lane evaluates (slice position, loop control, addresses, constant loads): twice the voxels per slice halve
// that share.  Used when every row is a whole number of tile pairs (nx % 256 == 0); same sign blocks, same values.
// Only for small SDF bodies (the lowering defines SDFK_DIST8): measured at 1024^3, README scene 0.765 -> 0.726 ms, but the
// Perf scene 1.60 -> 1.74 ms and CSG-50 7.6 -> 14.2 ms -- twice the hoisted per-voxel state no longer fits the registers.
#ifdef SDFK_DIST8
extern "C" __global__ void __launch_bounds__(SDFK_SAMPLE_WARPS * 32)
sdfk_k_sample_dist8(const sdfk_sample_params P, const sdfk_sample8_params Q, float* __restrict__ dist, uint4* __restrict__ signs)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned warp = threadIdx.x >> 5;
    const size_t plane = (size_t)P.nx * (size_t)P.ny;
    const size_t colwords = (size_t)((P.nzl + 31) >> 5) * 32u;

    for (unsigned work = blockIdx.x * SDFK_SAMPLE_WARPS + warp; work < Q.nwork8; work += gridDim.x * SDFK_SAMPLE_WARPS) {
        const unsigned seg = sdfk_div(work, Q.div_ncol8);
        const unsigned col = work - seg * Q.ncol8;
        const unsigned uy = sdfk_div(col, Q.div_tpr8);
        const unsigned xp = col - uy * (P.tiles_per_row >> 1);          // tile pair of the row: tiles 2 xp, 2 xp + 1
        const int iy = (int)uy;
        int zl0, zl1;
        {
            const long long a = ((long long)seg * P.nzl) / Q.zsplit8, b = ((long long)(seg + 1) * P.nzl) / Q.zsplit8;
            zl0 = seg == 0u ? 0 : min(P.nzl, (int)((a + 31) & ~31ll));
            zl1 = seg + 1u == Q.zsplit8 ? P.nzl : min(P.nzl, (int)((b + 31) & ~31ll));
        }
        const int x0 = (int)(xp * 256u + lane * 4u);                   // first voxel of the lane in tile A; tile B: + 128
        const float py = P.m1 + (float)iy * P.dy;
        const bool rowwall = P.clip && (iy == 0 || iy == P.ny - 1);
        float px[8];
        unsigned keep[8], setb[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int x = x0 + (k & 3) + ((k >> 2) << 7);
            px[k] = P.m0 + (float)x * P.dx;
            const bool w = P.clip && (x == 0 || x == P.nx - 1);
            keep[k] = w ? 0u : 0xFFFFFFFFu;
            setb[k] = w ? __float_as_uint(P.clip_value) : 0u;
        }
        size_t vbase = ((size_t)zl0 * P.ny + uy) * (size_t)P.nx + (size_t)(xp * 256u);
        uint4* const scol = signs + ((size_t)uy * P.tiles_per_row + 2u * xp) * colwords + lane;     // tile A; tile B: + colwords

        for (int zg0 = zl0; zg0 < zl1; zg0 += 32) {
            uint4 swa = make_uint4(0u, 0u, 0u, 0u), swb = make_uint4(0u, 0u, 0u, 0u);
            for (int zb0 = zg0; zb0 < min(zg0 + 32, zl1); zb0 += 8) {
                const int zend = min(zb0 + 8, zl1);
                const bool walls = rowwall || (P.clip && (zb0 + P.z_begin == 0 || zend + P.z_begin == P.nz));
                unsigned sa = 0u, sb = 0u, ssh = 0u;
                for (int zl = zb0; zl < zend; zl++, vbase += plane, ssh += 4u) {
                    const int iz = zl + P.z_begin;
                    float d[8];
                    if (walls && (rowwall || (P.clip && (iz == 0 || iz == P.nz - 1)))) {      // warp-uniform
#pragma unroll
                        for (int k = 0; k < 8; k++) d[k] = P.clip_value;
                    } else {
                        const float pz = P.m2 + (float)iz * P.dz;
                        sk_float4 r8[8];
                        sdf_eval_grid(px, py, pz, r8);
                        sdf_eval_grid(px + 4, py, pz, r8 + 4);
#pragma unroll
                        for (int k = 0; k < 8; k++) d[k] = __uint_as_float((__float_as_uint(r8[k].w) & keep[k]) | setb[k]);
                    }
                    sa |= sdfk_sign_nibble(d, P.sign_iso) << ssh;
                    sb |= sdfk_sign_nibble(d + 4, P.sign_iso) << ssh;
                    float4* const out = reinterpret_cast<float4*>(dist + vbase) + lane;
                    SDFK_ST(out, make_float4(d[0], d[1], d[2], d[3]));
                    SDFK_ST(out + 32, make_float4(d[4], d[5], d[6], d[7]));
                }
                const int sq = (zb0 >> 3) & 3;
                if (sq == 0) { swa.x = sa; swb.x = sb; } else if (sq == 1) { swa.y = sa; swb.y = sb; }
                else if (sq == 2) { swa.z = sa; swb.z = sb; } else { swa.w = sa; swb.w = sb; }
            }
            if (signs) {
                SDFK_ST(scol + (size_t)(zg0 >> 5) * 32u, swa);
                SDFK_ST(scol + colwords + (size_t)(zg0 >> 5) * 32u, swb);
            }
        }
    }
}
#endif

// Deferred vertex colours for distance-only voxels.  recipes[t] = (cell, edge) that created vertex t; the colour is
// recomputed exactly as Cell.AddFaceFromEdgeIndex / CalculateCenterVertex do (Cell.cs:313-357,501-549): corner colours
// come from sdf_eval at the corner voxels' sample positions (the very values K1 would have stored), the weights
// 1 / (1e-7 + |v - iso|) from the stored distances, in double, mixed exactly like the reference.
struct sdfk_color_params {
    float m0, m1, m2, dx, dy, dz;   // sample positions: p = m + i * d (as in K1)
    float iso;
    int nx, ny;                     // voxel grid (x, y)
    int z0;                         // global z of the first local slice
    int step;
    int ncx, ncy;                   // cells per row / rows per layer
    int k0;                         // global cell layer of local layer 0 of the cell ids
};

// weight and sample position of corner (dxc, dyc, dzc) of cell (i, j, kg)
static __device__ __forceinline__ sk_float3 sdfk_corner(const sdfk_color_params& P, const float* __restrict__ dist, int i, int j, int kg,
                                                        int dxc, int dyc, int dzc, double& w)
{
    const int x = (i + dxc) * P.step, y = (j + dyc) * P.step, z = (kg + dzc) * P.step;
    const float v = dist[((size_t)(z - P.z0) * P.ny + y) * (size_t)P.nx + x];
    w = 1.0 / (0.0000001 + fabs((double)v - (double)P.iso));
    return sk_make3(P.m0 + (float)x * P.dx, P.m1 + (float)y * P.dy, P.m2 + (float)z * P.dz);
}

extern "C" __global__ void __launch_bounds__(128)
sdfk_k_vertex_colors(const sdfk_color_params P, const uint2* __restrict__ recipes, const float* __restrict__ dist,
                     float* __restrict__ cols, long long nverts)
{
    // end corners of edge e as (dx | dy << 1 | dz << 2), EDGETORELATIVEPOS{X,Y,Z} (Luts.cs:26-28)
    const unsigned char end1[12] = {0, 1, 3, 2, 4, 5, 7, 6, 0, 1, 3, 2};
    const unsigned char end2[12] = {1, 3, 2, 0, 5, 7, 6, 4, 4, 5, 7, 6};
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nverts; t += (long long)gridDim.x * blockDim.x) {
        const uint2 rc = recipes[t];
        const int i = (int)(rc.x % (unsigned)P.ncx);
        const unsigned t2 = rc.x / (unsigned)P.ncx;
        const int j = (int)(t2 % (unsigned)P.ncy);
        const int kg = P.k0 + (int)(t2 / (unsigned)P.ncy);
        float cr, cg, cb;
        if (rc.y < 12u) {
            const int a = end1[rc.y], c = end2[rc.y];
            double w1, w2;
            const sk_float3 p1 = sdfk_corner(P, dist, i, j, kg, a & 1, (a >> 1) & 1, a >> 2, w1);
            const sk_float3 p2 = sdfk_corner(P, dist, i, j, kg, c & 1, (c >> 1) & 1, c >> 2, w2);
            sk_float4 c1, c2;
            sdf_eval2(p1, p2, c1, c2);                              // both end corners in one packed evaluation
            double ff = 0.0;
            ff += w1;
            ff += w2;
            const float f1 = (float)w1, f2 = (float)w2;
            cr = (float)((double)(c1.x * f1 + c2.x * f2) / ff);
            cg = (float)((double)(c1.y * f1 + c2.y * f2) / ff);
            cb = (float)((double)(c1.z * f1 + c2.z * f2) / ff);
        } else {   // centre vertex: corners 0..7 in the reference's numbering
            double ff = 0.0;
            float fr = 0.f, fg = 0.f, fb = 0.f;
#pragma unroll
            for (int q = 0; q < 8; q += 2) {
                double wa, wb;
                const sk_float3 pa = sdfk_corner(P, dist, i, j, kg, (0x66 >> q) & 1, (0xCC >> q) & 1, (0xF0 >> q) & 1, wa);
                const sk_float3 pb = sdfk_corner(P, dist, i, j, kg, (0x66 >> (q + 1)) & 1, (0xCC >> (q + 1)) & 1, (0xF0 >> (q + 1)) & 1, wb);
                sk_float4 ca, cb4;
                sdf_eval2(pa, pb, ca, cb4);
                ff += wa;
                const float fa = (float)wa;
                if (q == 0) { fr = ca.x * fa; fg = ca.y * fa; fb = ca.z * fa; }
                else { fr = fr + ca.x * fa; fg = fg + ca.y * fa; fb = fb + ca.z * fa; }
                ff += wb;
                const float fbw = (float)wb;
                fr = fr + cb4.x * fbw; fg = fg + cb4.y * fbw; fb = fb + cb4.z * fbw;
            }
            cr = (float)((double)fr / ff);
            cg = (float)((double)fg / ff);
            cb = (float)((double)fb / ff);
        }
        cols[t * 3 + 0] = cr;
        cols[t * 3 + 1] = cg;
        cols[t * 3 + 2] = cb;
    }
}

// K6: colorsAndDistances[i] = sdf(points[i])  (Sdf.cs:8).  xyz: n*3 floats, rgbd: n*4 floats.
extern "C" __global__ void __launch_bounds__(256)
sdfk_k_eval(const float* __restrict__ xyz, float* __restrict__ rgbd, long long n)
{
    const long long npair = (n + 1) >> 1;                             // two points per thread: packed f32x2 evaluation
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < npair; q += (long long)gridDim.x * blockDim.x) {
        const long long i = 2 * q, i1 = (i + 1 < n) ? i + 1 : i;
        sk_float4 r0, r1;
        sdf_eval2(sk_make3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), sk_make3(xyz[3 * i1], xyz[3 * i1 + 1], xyz[3 * i1 + 2]), r0, r1);
        reinterpret_cast<float4*>(rgbd)[i] = make_float4(r0.x, r0.y, r0.z, r0.w);
        if (i1 != i) reinterpret_cast<float4*>(rgbd)[i1] = make_float4(r1.x, r1.y, r1.z, r1.w);
    }
}

struct sdfk_render_params {
    int w, h;                  // full image size
    int row_begin, row_end;    // rows rendered by this call (row band, RayMarcher.cs:50-61)
    float cam[3];              // camera position = translation of inverse(view)   (RayMarcher.cs:97-99)
    float ivp[16];             // inverse(view * projection), row-major             (RayMarcher.cs:107-108)
    float nearp, farp;
    int iters;
};

// Ray through pixel (i,j): RayMarcher.GetCameraRays per-pixel part (RayMarcher.cs:111-125)
static __device__ __forceinline__ sk_float3 sdfk_ray_dir(const sdfk_render_params& P, int i, int j)
{
    const float y = 1.0f - 2.0f * (float)j / (float)(P.h - 1);
    const float x = -1.0f + 2.0f * (float)i / (float)(P.w - 1);
    // Vector4.Transform((x, y, 0, 1), ivp): ((x*m1 + y*m2) + 0*m3) + 1*m4
    const float tx = x * P.ivp[0] + y * P.ivp[4] + 0.0f * P.ivp[8] + 1.0f * P.ivp[12];
    const float ty = x * P.ivp[1] + y * P.ivp[5] + 0.0f * P.ivp[9] + 1.0f * P.ivp[13];
    const float tz = x * P.ivp[2] + y * P.ivp[6] + 0.0f * P.ivp[10] + 1.0f * P.ivp[14];
    const float tw = x * P.ivp[3] + y * P.ivp[7] + 0.0f * P.ivp[11] + 1.0f * P.ivp[15];
    const float dx = tx / tw - P.cam[0], dy = ty / tw - P.cam[1], dz = tz / tw - P.cam[2];
    const float len = sk_sqrt((dx * dx + dy * dy) + dz * dz);      // Vector3.Normalize = v / v.Length()
    return sk_make3(dx / len, dy / len, dz / len);
}

// Shading of one pixel from its march result (RayMarcher.cs:146-160): gradient taps already evaluated.
static __device__ __forceinline__ void sdfk_shade(const sdfk_render_params& P, float sx, float sy, float sz, float depth, float cr, float cg, float cb,
                                                  float px, float py, float pz, float mx, float my, float mz, float* __restrict__ out)
{
    float nx = px - mx, ny = py - my, nz = pz - mz;
    const float nl = sk_sqrt(nx * nx + ny * ny + nz * nz);     // NormalizeInplace (VectorData.cs:490-510)
    if (nl > 0.0f) { const float r = 1.0f / nl; nx = nx * r; ny = ny * r; nz = nz * r; }
    float lx = 5.0f - sx, ly = 5.0f - sy, lz = 10.0f - sz;     // light (5,5,10) (RayMarcher.cs:149-150)
    const float ll = sk_sqrt(lx * lx + ly * ly + lz * lz);
    if (ll > 0.0f) { const float r = 1.0f / ll; lx = lx * r; ly = ly * r; lz = lz * r; }
    float dv = nx * lx + ny * ly + nz * lz;                    // Dot (VectorData.cs:464-475)
    dv = (dv != dv) ? dv : ((dv > 0.0f) ? dv : 0.0f);          // MaxInplace(0): MathF.Max, NaN propagates
    const float mask = depth > P.farp ? 1.0f : 0.0f;           // bgMask (RayMarcher.cs:156)
    const float notmask = mask == 0.0f ? 1.0f : 0.0f;
    // fg = (dv*colour + 0.1) * notmask + mask*bg ; frag(=0) += fg  (RayMarcher.cs:154-160)
    out[0] = 0.0f + ((dv * cr + 0.1f) * notmask + mask * 0.5f);
    out[1] = 0.0f + ((dv * cg + 0.1f) * notmask + mask * 0.75f);
    out[2] = 0.0f + ((dv * cb + 0.1f) * notmask + mask * 1.0f);
}

// One thread per group of PPT = 2 neighbouring pixels (every SDF evaluation is one two-point sdf_eval2 call, whose range guards
// are shared by both points).  The reference's ~12 full-image temporaries per iteration live in registers.  (PPT = 4 with a
// four-point evaluator was measured: 0.158 -> 0.168 ms on the README scene at 1080p, 100 registers instead of 62.)
template <int PPT>
static __device__ __forceinline__ void sdfk_eval_n(const sk_float3* p, sk_float4* r)
{
#pragma unroll
    for (int k = 0; k < PPT; k += 2) sdf_eval2(p[k], p[k + 1], r[k], r[k + 1]);
}

template <int PPT, bool DEPTH_ONLY>
static __device__ __forceinline__ void sdfk_render_body(const sdfk_render_params& P, float* __restrict__ out)
{
    const long long npix = (long long)(P.row_end - P.row_begin) * P.w;
    const long long ngrp = (npix + PPT - 1) / PPT;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < ngrp; q += (long long)gridDim.x * blockDim.x) {
        long long pix[PPT];
        sk_float3 rd[PPT];
        float depth[PPT];
#pragma unroll
        for (int k = 0; k < PPT; k++) {
            pix[k] = (PPT * q + k < npix) ? PPT * q + k : PPT * q;       // a ragged tail repeats the group's first pixel
            rd[k] = sdfk_ray_dir(P, (int)(pix[k] % P.w), P.row_begin + (int)(pix[k] / P.w));
            depth[k] = P.nearp - 0.1f;                                   // RayMarcher.cs:136
        }
        // fixed count, no early out (RayMarcher.cs:138-145).  Only the LAST sample's colour is kept (RayMarcher.cs:143-144), so
        // the last iteration is peeled: in the loop the colour outputs are dead and the compiler drops everything that only
        // feeds them (for the README scene two of the four divisions per evaluation).
        sk_float3 p[PPT];
        sk_float4 s[PPT];
        for (int it = 0; it + 1 < P.iters; it++) {
#pragma unroll
            for (int k = 0; k < PPT; k++) p[k] = sk_make3(rd[k].x * depth[k] + P.cam[0], rd[k].y * depth[k] + P.cam[1], rd[k].z * depth[k] + P.cam[2]);
            sdfk_eval_n<PPT>(p, s);
#pragma unroll
            for (int k = 0; k < PPT; k++) depth[k] = depth[k] + s[k].w;
        }
        float col[PPT][3];
#pragma unroll
        for (int k = 0; k < PPT; k++) col[k][0] = col[k][1] = col[k][2] = 0.0f;
        if (P.iters > 0) {
#pragma unroll
            for (int k = 0; k < PPT; k++) p[k] = sk_make3(rd[k].x * depth[k] + P.cam[0], rd[k].y * depth[k] + P.cam[1], rd[k].z * depth[k] + P.cam[2]);
            sdfk_eval_n<PPT>(p, s);
#pragma unroll
            for (int k = 0; k < PPT; k++) {
                depth[k] = depth[k] + s[k].w;
                col[k][0] = col[k][0] + s[k].x; col[k][1] = col[k][1] + s[k].y; col[k][2] = col[k][2] + s[k].z;
            }
        }
        if (DEPTH_ONLY) {
#pragma unroll
            for (int k = 0; k < PPT; k++)
                if (k == 0 || pix[k] != pix[0]) out[pix[k]] = depth[k];
            continue;
        }
        sk_float3 hit[PPT];
#pragma unroll
        for (int k = 0; k < PPT; k++) hit[k] = sk_make3(P.cam[0] + rd[k].x * depth[k], P.cam[1] + rd[k].y * depth[k], P.cam[2] + rd[k].z * depth[k]);
        // DistanceGradient: taps +x,+y,+z,-x,-y,-z at GradOffset = 1e-5 (RayMarcher.cs:29,164-204), all PPT pixels per call
        const float go = 1e-5f;
        float g[PPT][6];
#pragma unroll
        for (int t = 0; t < 6; t++) {
            const float sg = t < 3 ? go : -go;
            const float ox = (t % 3 == 0) ? 1.0f : 0.0f, oy = (t % 3 == 1) ? 1.0f : 0.0f, oz = (t % 3 == 2) ? 1.0f : 0.0f;
#pragma unroll
            for (int k = 0; k < PPT; k++) p[k] = sk_make3(hit[k].x + sg * ox, hit[k].y + sg * oy, hit[k].z + sg * oz);
            sdfk_eval_n<PPT>(p, s);
#pragma unroll
            for (int k = 0; k < PPT; k++) g[k][t] = s[k].w;
        }
#pragma unroll
        for (int k = 0; k < PPT; k++)
            if (k == 0 || pix[k] != pix[0])
                sdfk_shade(P, hit[k].x, hit[k].y, hit[k].z, depth[k], col[k][0], col[k][1], col[k][2], g[k][0], g[k][1], g[k][2], g[k][3], g[k][4], g[k][5],
                           out + pix[k] * 3);
    }
}

extern "C" __global__ void __launch_bounds__(128)
sdfk_k_render(const sdfk_render_params P, float* __restrict__ rgb)
{
    sdfk_render_body<2, false>(P, rgb);
}

extern "C" __global__ void __launch_bounds__(128)
sdfk_k_render_depth(const sdfk_render_params P, float* __restrict__ depth_out)
{
    sdfk_render_body<2, true>(P, depth_out);
}

