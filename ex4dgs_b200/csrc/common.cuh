// Shared device helpers, buffer layouts and kernel parameter blocks of the sm_100a rasterizer.
//
// Written from scratch for B200; the *arithmetic* of the integer-determining quantities (depth
// key, pixel centre, radius, tile rectangle, alpha thresholds) is pinned with explicit
// __fmaf_rn/__fmul_rn/__fadd_rn so that it reproduces, bit for bit, what nvcc's FMA contraction
// makes of the reference's expressions (read off the SASS of the reference build, see DESIGN.md
// "pinned arithmetic"; reference sources: cuda_rasterizer/forward.cu:74-124,128-162,165-269,
// auxiliary.h:41-87,267-294).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#define EX_TILE 16            // config.h:16-17 (BLOCK_X/BLOCK_Y) - part of the key contract
#define EX_TILE_PIX 256
#define EX_INVISIBLE_KEY 0xFFFFFFFFu

// ---- pinned float arithmetic -------------------------------------------------------------------
__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float ff(float a, float b, float c) { return __fmaf_rn(a, b, c); }
// a0*b0 + a1*b1 + a2*b2 as the reference build evaluates it:  fma(a2,b2, fma(a0,b0, rn(a1*b1)))
__device__ __forceinline__ float sum3(float a0, float b0, float a1, float b1, float a2, float b2)
{
    return ff(a2, b2, ff(a0, b0, fm(a1, b1)));
}
// m[c]*x + m[c+4]*y + m[c+8]*z + m[c+12]   (auxiliary.h:68-87)
__device__ __forceinline__ float xform_row(const float* __restrict__ m, int c, float x, float y, float z)
{
    return fa(sum3(x, m[c], y, m[c + 4], z, m[c + 8]), m[c + 12]);
}

// Per-Gaussian record consumed by the per-tile compositing loops (forward and backward).  One
// 64-byte line per Gaussian, gathered by id into shared memory with 16-byte async copies.
//   a = (pixel x, pixel y, view depth, skip threshold)       b = (conic A, conic B, conic C, opacity*coef)
//   c = (r, g, b, id as int bits)                            d = (dir3D x, y, z, 0)
struct __align__(16) SplatRec {
    float4 a, b, c, d;
};

// Per-Gaussian gradient accumulator written by the backward compositing loop with 16-byte
// vector reductions, consumed by the fused preprocess backward.
//   g0 = (dL/dmean2D.x, .y, .z, dL/dopacity)   g1 = (dL/dconic.x, .y, .w, 0)
//   g2 = (dL/dcolor r, g, b, 0)                g3 = (dL/ddir x, y, z, 0)
struct __align__(16) GradAcc {
    float4 g0, g1, g2, g3;
};

struct Carver {
    char* base;
    size_t off;
    __host__ explicit Carver(void* b) : base(reinterpret_cast<char*>(b)), off(0) {}
    template <typename T>
    __host__ T* take(size_t n)
    {
        off = (off + 255) & ~size_t(255);
        T* p = reinterpret_cast<T*>(base + off);   // base may be nullptr: size computation only
        off += n * sizeof(T);
        return p;
    }
};

struct GeometryState {
    uint32_t* key_in;         // [P] depth bits, EX_INVISIBLE_KEY when culled
    uint32_t* val_in;         // [P] iota
    uint32_t* key_sorted;     // [P]
    uint32_t* order;          // [P] Gaussian ids by (depth, id)
    uint32_t* tiles_touched;  // [P]
    uint32_t* offsets;        // [P] inclusive scan of tiles_touched in `order`
    SplatRec* rec;            // [P]
    uint8_t* clamped;         // [P] bit c = colour channel c was clamped (forward.cu:67-69)
    GradAcc* gacc;            // [P]
    uint32_t* meta;           // [64] misc device scalars
    char* temp;               // cub temp storage
    size_t temp_bytes;
    size_t total;
};

struct BinningState {
    uint16_t* tile_unsorted;  // [R]
    uint32_t* val_unsorted;   // [R]
    uint16_t* tile_sorted;    // [R]
    uint32_t* point_list;     // [R] Gaussian ids, by (tile, depth, id): == reference point_list
    char* temp;
    size_t temp_bytes;
    size_t total;
};

struct ImageState {
    float* final_T;           // [W*H]
    uint32_t* n_contrib;      // [W*H]
    uint2* ranges;            // [tiles]
    uint32_t* tile_batches;   // [tiles] number of 256-splat batches the forward fetched (stats)
    size_t total;
};

GeometryState carve_geometry(void* base, int P, size_t temp_bytes);
BinningState carve_binning(void* base, int R, size_t temp_bytes);
ImageState carve_image(void* base, int width, int height);

// ---- kernel parameter blocks ---------------------------------------------------------------------
struct PreprocessParams {
    int P, D, M;
    const float* means3D;
    const float* dir3D;
    const float* scales;
    const float* rotations;
    const float* opacities;
    const float* shs;
    const float* cov3D_precomp;
    const float* colors_precomp;
    float scale_modifier;
    const float* view;   // [16] device
    const float* proj;   // [16] device
    const float* cam;    // [3]  device
    int W, H;
    float tan_fovx, tan_fovy, focal_x, focal_y, kernel_size;
    float min_depth, max_depth;
    int grid_x, grid_y;
    int prefiltered;
    unsigned flags;
    const float* pad_ptr;   // max |subpixel offset| (device scalar), read when EX4DGS_FLAG_TILE_CULL
    int* radii;
    uint32_t* key_in;
    uint32_t* val_in;
    uint32_t* tiles_touched;
    SplatRec* rec;
    uint8_t* clamped;
};

struct RenderParams {
    const uint2* ranges;
    const uint32_t* point_list;
    const SplatRec* rec;
    int W, H, grid_x;
    const float2* subpixel_offset;
    const float* bg;     // [3] device
    float min_depth, max_depth;
    // forward outputs / backward inputs
    float* final_T;
    uint32_t* n_contrib;
    uint32_t* tile_batches;
    float* out_color;
    float* out_depth;
    float* out_acc;
    float* out_flow;
    int* out_idx;
    // backward
    const float* dL_dpix;
    const float* dL_ddepth;
    const float* dL_dflow;
    const float* dL_dacc;
    GradAcc* gacc;
};

struct PreprocessBwdParams {
    int P, D, M;
    const float* means3D;
    const float* scales;
    const float* rotations;
    const float* shs;
    const float* cov3D_precomp;
    const float* colors_precomp;
    float scale_modifier;
    const float* view;   // [16] device
    const float* proj;   // [16] device
    const float* cam;    // [3]  device
    float tan_fovx, tan_fovy, focal_x, focal_y, kernel_size;
    const int* radii;
    const uint8_t* clamped;
    const GradAcc* gacc;
    float* dL_dmean2D;
    float* dL_dopacity;
    float* dL_dcolor;
    float* dL_dmean3D;
    float* dL_dcov3D;
    float* dL_dsh;
    float* dL_dscale;
    float* dL_drot;
    float* dL_ddir;
};

// ---- launchers (defined in the .cu files) -----------------------------------------------------------
void launch_preprocess_fwd(const PreprocessParams& p, cudaStream_t s);
void launch_preprocess_bwd(const PreprocessBwdParams& p, cudaStream_t s);
void launch_mark_visible(int P, const float* means3D, const float* view, const float* proj,
                         float min_depth, float max_depth, uint8_t* present, cudaStream_t s);
void launch_render_fwd(const RenderParams& p, int grid_x, int grid_y, cudaStream_t s);
void launch_render_bwd(const RenderParams& p, int grid_x, int grid_y, cudaStream_t s);

cudaError_t launch_subpixel_absmax(const float* subpixel_offset, size_t n, uint32_t* out, cudaStream_t s);
size_t binning_stage1_temp_bytes(int P);
size_t binning_stage2_temp_bytes(int R);
// sort Gaussians by depth bits, scan tiles_touched in that order; returns cudaError
cudaError_t binning_stage1(const GeometryState& g, int P, cudaStream_t s);
// emit (tile, id) pairs in depth order, stable-sort by tile, find per-tile ranges
cudaError_t binning_stage2(const GeometryState& g, const BinningState& b, const ImageState& img,
                           const int* radii, int P, int R, int grid_x, int grid_y, unsigned flags,
                           cudaStream_t s);

// ---- tile rectangle (auxiliary.h:46-56), shared by preprocess and the duplicate kernel -----------------
__device__ __forceinline__ void tile_rect(float px, float py, int radius, int grid_x, int grid_y,
                                          int& x0, int& y0, int& x1, int& y1)
{
    const float r = (float)radius;
    // (int)((p - r) / 16): the reference build multiplies by 0.0625f (exact) and truncates
    int ax0 = (int)fm(fa(px, -r), 0.0625f);
    int ay0 = (int)fm(fa(py, -r), 0.0625f);
    // (int)((p + r + 16 - 1) / 16), evaluated left to right in float
    int ax1 = (int)fm(fa(fa(fa(px, r), 16.0f), -1.0f), 0.0625f);
    int ay1 = (int)fm(fa(fa(fa(py, r), 16.0f), -1.0f), 0.0625f);
    x0 = min(grid_x, max(0, ax0));
    y0 = min(grid_y, max(0, ay0));
    x1 = min(grid_x, max(0, ax1));
    y1 = min(grid_y, max(0, ay1));
}

// Exact-output tile culling (EX4DGS_FLAG_TILE_CULL): true when NO pixel of tile (tx,ty) can pass
// the `alpha >= 1/255` test of the compositing loop for this splat, i.e. when the maximum of
// `power` over the tile's pixel rectangle is below the splat's skip threshold.  Pixel centres
// are integers + subpixel offset; callers pass a rectangle already widened by the maximal
// |subpixel offset| they allow (offsets are in [-0.5,0.5] in the application; the flag is ignored
// by the host wrapper when it cannot guarantee that).
__device__ __forceinline__ bool tile_cannot_contribute(float cx, float cy, float A, float B, float C,
                                                       float thr, int tx, int ty, float pad)
{
    // NOTE: evaluated by two different kernels (count in preprocess, emit in duplicate) that must
    // agree exactly, so every operation is pinned (no compiler-chosen FMA contraction).
    // q(d) = 0.5*(A dx^2 + C dy^2) + B dx dy  (= -power), minimised over the tile's pixel rectangle.
    // The quadratic is convex when A,C > 0 and AC > B^2; otherwise be conservative.
    if (!(pad <= 4096.f)) return false;          // NaN / absurd subpixel offsets: never cull
    const float detc = fa(fm(A, C), -fm(B, B));
    if (!(A > 0.f) || !(C > 0.f) || !(detc > 0.f)) return false;
    const float dx0 = fa(fa((float)(tx * EX_TILE), -pad), -cx), dx1 = fa(fa((float)(tx * EX_TILE + EX_TILE - 1), pad), -cx);
    const float dy0 = fa(fa((float)(ty * EX_TILE), -pad), -cy), dy1 = fa(fa((float)(ty * EX_TILE + EX_TILE - 1), pad), -cy);
    if (dx0 <= 0.f && dx1 >= 0.f && dy0 <= 0.f && dy1 >= 0.f) return false;   // centre inside
    // centre outside => the minimum lies on the boundary: minimise over the 4 edges
    float qmin = 3.4e38f;
#pragma unroll
    for (int e = 0; e < 2; e++) {
        const float dx = e ? dx1 : dx0;
        const float dy = fminf(dy1, fmaxf(dy0, __fdiv_rn(-fm(B, dx), C)));
        const float q = fa(fm(0.5f, fa(fm(fm(A, dx), dx), fm(fm(C, dy), dy))), fm(fm(B, dx), dy));
        qmin = fminf(qmin, q);
    }
#pragma unroll
    for (int e = 0; e < 2; e++) {
        const float dy = e ? dy1 : dy0;
        const float dx = fminf(dx1, fmaxf(dx0, __fdiv_rn(-fm(B, dy), A)));
        const float q = fa(fm(0.5f, fa(fm(fm(A, dx), dx), fm(fm(C, dy), dy))), fm(fm(B, dx), dy));
        qmin = fminf(qmin, q);
    }
    // power_max = -qmin.  The compositing loop evaluates `power` in float with a handful of
    // roundings on terms that may cancel; the evaluated value is <= -q*(1 - eps*kappa) where
    // kappa = (sqrt(AC)+|B|)^2/(AC-B^2) bounds (sum of |terms|)/q.  eps = 2e-5 is ~50x the real
    // worst case (and also covers the rounding of this bound itself); very ill-conditioned conics
    // are simply not culled.
    const float sAC = fa(__fsqrt_rn(fm(A, C)), fabsf(B));
    const float kappa = __fdiv_rn(fm(sAC, sAC), detc);
    const float shrink = fa(1.0f, -fm(2e-5f, kappa));
    if (!(shrink > 0.5f)) return false;
    return fa(fm(-qmin, shrink), 1e-3f) < thr;
}
