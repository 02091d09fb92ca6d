// Shared device helpers, buffer layouts and kernel parameter blocks of the sm_100a rasterizer.
//
// Written from scratch for B200; the *arithmetic* of the integer-determining quantities (depth
// key, pixel centre, radius, tile rectangle, alpha thresholds) is pinned with explicit
// __fmaf_rn/__fmul_rn/__fadd_rn so that it reproduces, bit for bit, what nvcc's FMA contraction
// makes of the reference's expressions (read off the SASS of the reference build, see DESIGN.md
// "pinned arithmetic"; reference sources: cuda_rasterizer/forward.cu:74-124,128-162,165-269,
// auxiliary.h:41-87,267-294).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#define EX_TILE 16            // config.h:16-17 (BLOCK_X/BLOCK_Y) - part of the key contract
#define EX_TILE_PIX 256
#define EX_INVISIBLE_KEY 0xFFFFFFFFu

// tuning knobs of the compositing kernels (defaults = measured best on B200, see profiles/)
#ifndef EX_FWD_MINBLOCKS
#define EX_FWD_MINBLOCKS 7      // CTAs of 128 threads (2 pixels per lane): 72 registers, 28 resident warps; 0.392 ms vs 0.407 (6) / 0.398 (8, groups of 2) at C3
#endif
#ifndef EX_BWD_MINBLOCKS
#define EX_BWD_MINBLOCKS 6      // CTAs of 128 threads (2 pixels per lane, packed): 78 registers, 24 resident warps; 0.712 ms vs 0.744 (5) / 0.758 (4) at C3
#endif
#ifndef EX_FWD_STAGE_LDGSTS
#define EX_FWD_STAGE_LDGSTS 1   // forward staging: 1 = per-thread 16-byte cp.async (LDGSTS), 0 = one TMA bulk copy per splat.
                                // A bulk copy takes uniform registers, so 32 per-lane copies become a 32-trip loop of 9
                                // instructions (7 % of the kernel's issue slots, ncu r1i); measured 0.422 vs 0.439 ms at C3.
#endif
#ifndef EX_BWD_STAGE_GATHER4
#define EX_BWD_STAGE_GATHER4 1  // backward staging: 1 = TMA tile::gather4 (UTMALDG.2D.GATHER4: FOUR 64-byte records per instruction, row
                                // indices = Gaussian ids, over a [P,16] float tensor map of the record array), 0 = one 48-byte bulk copy per record
#endif
#ifndef EX_BWD_FAST_RCP
#define EX_BWD_FAST_RCP 1       // T /= (1 - alpha) with MUFU.RCP: 1.011 vs 1.075 ms at C3, gradients within the 1e-3 budget
#endif
#ifndef EX_BWD_FAST_EXP
#define EX_BWD_FAST_EXP 1       // backward recomputes G = exp(power) with MUFU.EX2 directly (see render_bwd.cu)
#endif
#ifndef EX_BWD_PPT
#define EX_BWD_PPT 2            // pixels per thread in the backward compositing kernel (1: 8 warps x 8x4, 2: 4 warps x 8x8);
                                // PPT = 1 needs EX_BWD_MINBLOCKS <= 4 (256-thread CTAs); measured 0.982 (1) vs 0.926 ms (2)
#endif
#define EX_BLOCK_TEST block_reject   // an exact 4-edge variant was measured slower (more instructions than it saves)

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

// Per-Gaussian gradient accumulator written by the backward compositing loop with warp-level
// reductions, consumed by the fused preprocess backward through gacc_load().
//   g0 = (dL/dmean2D.x, .y, .z, dL/dopacity)   g1 = (dL/dconic.x, .y, .w, 0)
//   g2 = (dL/dcolor r, g, b, 0)                g3 = (dL/ddir x, y, z, 0)
// The compositing loop leaves everything that is constant per Gaussian out of its inner loop: it stores
// g0.x = S1 = sum gG dx and g0.y = S2 = sum gG dy instead of dL/dmean2D = (-W/2 (A S1 + B S2), -H/2 (C S2 + B S1))
// (the 2x2 conic is applied here, once per Gaussian, from the record), and g1.xyz / (-1/2).
struct __align__(16) GradAcc {
    float4 g0, g1, g2, g3;
};

__device__ __forceinline__ GradAcc gacc_load(const GradAcc* gacc, const SplatRec* rec, int idx, float W, float H)
{
    GradAcc g = gacc[idx];
    if (g.g0.x != 0.f || g.g0.y != 0.f) {       // only Gaussians the compositing backward touched have a record
        const float4 b = rec[idx].b;
        const float S1 = g.g0.x, S2 = g.g0.y;
        g.g0.x = (b.x * S1 + b.y * S2) * (-0.5f * W);
        g.g0.y = (b.z * S2 + b.y * S1) * (-0.5f * H);
    }
    g.g1.x *= -0.5f;
    g.g1.y *= -0.5f;
    g.g1.z *= -0.5f;
    return g;
}

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

// device scalars in GeometryState::meta (zeroed at the start of every forward)
#define EX_META_PAD 0        // bits of max |subpixel offset| (exact tile culling)
#define EX_META_FLOW 1       // does any visible Gaussian carry a non-zero dir3D?
#define EX_META_INEXACT 2    // alpha thresholds that fell back to the conservative value (preprocess.cu)
#define EX_META_NVIS 3       // visible Gaussians (depth histogram kernel)
#define EX_META_TOTAL 4      // R: (Gaussian, tile) instances (touched_sums_kernel)
#define EX_META_ERROR 5      // bit 0: a look-back did not complete; bit 1: more instances than the binning buffer holds
#define EX_META_LISTED 6     // entries of the sorted (tile, id) list with exact tile culling: instances the test kept, at most the capacity (tile histogram kernel)
#define EX_META_TICKETS 8    // [0..3] depth passes, [5..8] tile passes

struct GeometryState {
    uint32_t* key_in;         // [P] depth bits, EX_INVISIBLE_KEY when culled
    uint32_t* key_a;          // [P] ping-pong arrays of the depth sort (only the first N_vis entries are used)
    uint32_t* val_a;          // [P]
    uint32_t* key_b;          // [P]
    uint32_t* order;          // [P] ids of the visible Gaussians by (depth, id): N_vis entries
    uint32_t* tiles_touched;  // [P]
    SplatRec* rec;            // [P]
    uint8_t* clamped;         // [P] bit c = colour channel c was clamped (forward.cu:67-69)
    GradAcc* gacc;            // [P]
    uint32_t* meta;           // [64] device scalars (EX_META_*), directly followed by ...
    char* temp;               // ... the sort scratch (binning.cu: histograms, look-back words)
    size_t temp_bytes;
    size_t total;
};

// Two sets of (tile, id) arrays of `cap` entries each.  Set 0 sits at the front of the buffer (point_list at
// offset 0) and holds the sorted lists at the end; the duplicate kernel writes set 0 when the tile sort has two
// passes (0 -> 1 -> 0) and set 1 when it has one.
struct BinningState {
    void* tile[2];            // [cap] uint16 tile ids (uint32 when the image has more than 65535 tiles); tile[0] = sorted tile ids at the end
    int key_bytes;            // 2 or 4
    uint32_t* val[2];         // [cap]; val[0] = point_list: Gaussian ids by (tile, depth, id) == reference point_list
    uint32_t* status;         // look-back words of the tile passes
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
BinningState carve_binning(void* base, int cap, size_t status_bytes, int key_bytes);
ImageState carve_image(void* base, int width, int height);

// ---- kernel parameter blocks ---------------------------------------------------------------------
// SH coefficients kept in the model's four tensors (EX4DGS_FLAG_SH_SEGMENTED): index 0 = static
// Gaussians [0, n_static), 1 = dynamic; dc [n,1,3], rest [n,15,3].  enabled = 0: plain [P,M,3] array.
struct ShSegments {
    int enabled;
    int n_static;
    float* dc[2];
    float* rest[2];
};

struct PreprocessParams {
    int P, D, M;
    const float* means3D;
    const float* dir3D;
    const float* scales;
    const float* rotations;
    const float* opacities;
    const float* shs;
    ShSegments seg;
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
    uint32_t* tiles_touched;
    SplatRec* rec;
    uint8_t* clamped;
    uint32_t* flow_flag;    // device word, set to 1 when a visible Gaussian has a non-zero dir3D component
    uint32_t* inexact_thr;  // device counter: visible Gaussians whose alpha threshold fell back to the conservative value
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
    ShSegments seg;       // inputs when segmented
    ShSegments dseg;      // gradient outputs when segmented (dL_dsh unused then)
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
    const SplatRec* rec; // forward records (conic for gacc_load)
    float W, H;          // image size (gacc_load)
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
void launch_render_fwd(const RenderParams& p, const CUtensorMap* rec_map, int grid_x, int grid_y, bool with_flow, cudaStream_t s);
bool render_fwd_uses_gather();    // compiled with EX_FWD_STAGE=2: needs the record tensor map
void render_fwd_geometry(int* batch, int* warps);     // splats staged per batch, warps per tile (statistics word of tile_batches)
void launch_render_bwd(const RenderParams& p, const CUtensorMap* rec_map, int grid_x, int grid_y, cudaStream_t s);
// [P,16] float32 tensor map over the record array (box {16,1}) for the TMA gathers; false when the driver entry point is missing
bool make_record_tensor_map(CUtensorMap* out, const SplatRec* rec, int P);
// photometric loss (loss.cu): scratch = per-block partial sums + the three SSIM derivative maps
size_t loss_scratch_bytes(int W, int H);
cudaError_t launch_loss_forward(int W, int H, const float* img, const float* gt, float lambda, char* scratch,
                                float* out3, float* l1_err, float* ssim_err, cudaStream_t stream);
cudaError_t launch_loss_backward(int W, int H, const float* img, const float* gt, float lambda, const char* scratch,
                                 const float* dL_dloss, float* dL_dimg, cudaStream_t stream);

// fused RAdam step (optim.cu)
#define EX_OPT_MAX_TENSORS 32
struct RAdamTensorDesc {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    size_t numel;
    float S;            // -sqrt(1 - beta2^t) * lr * rect / (1 - beta1^t)
    float U;            // -lr / (1 - beta1^t) while the variance is not tractable (rho_t <= 5), else 0
    int rectified;      // rho_t > 5
    int aligned;        // all four pointers 16-byte aligned (filled by the launcher)
    int index;          // position in the caller's table (slot of nan_flags)
    int check_nan;      // report NaN parameters of this tensor in nan_flags[index]
    int sanitize_grad;  // nan_to_num the gradient first
};
cudaError_t launch_radam(const RAdamTensorDesc* tensors, int n, double beta1, double beta2, double eps, double grad_scale,
                         int* nan_flags, cudaStream_t s);

// row gathers of densification / pruning (compact.cu)
#define EX_GATHER_MAX_JOBS 64
struct GatherJob {            // dst[r] = r < n_a ? a[index ? index[r] : r] : (b ? b[r - n_a] : 0),  r in [0, n_out), rows of `words` 4-byte words
    const uint32_t* a;
    const uint32_t* b;
    uint32_t* dst;
    const long long* index;
    unsigned words, n_a, n_out;
    int vec4;                 // 128-bit accesses possible (filled by the launcher)
};
cudaError_t launch_gather_rows(const GatherJob* jobs, int n, cudaStream_t s);

size_t l1_scratch_bytes();
cudaError_t launch_l1_forward(size_t n, const float* a, const float* b, char* scratch, float* out, cudaStream_t s);
cudaError_t launch_l1_backward(size_t n, const float* a, const float* b, const float* g, float* da, cudaStream_t s);

// per-iteration statistics and regularisers (stats.cu)
struct StatsArrays {          // one set for the static Gaussians, one for the dynamic ones
    float* max_radii2D;
    float* min_radii2D;
    float* xyz_gradient_accum;
    float* denom;
    float* error_accum;       // xyz_error_accum / motion_xyz_error_mean
    float* error_min;
    float* error_min_timestamp;
    float* ssim_error_accum;
    float* error_denom;
};
struct IterStatsParams {
    int Ns, Nd;
    const int* radii;             // [Ns + Nd]
    const float* grad_means2D;    // [Ns + Nd, 3]
    const float* grad_error;      // [Ns + Nd, 3] or nullptr (l1_accum off)
    float timestamp;
    int densify;
    StatsArrays stat, dyn;
};
struct RegParams {
    long long Ns, Nd;
    int K;
    const float* xyz_disp;        // [Ns, 3]
    const float* xyz_motion;      // [Nd, K, 3]
    float static_coef;            // static_reg / Ns          (0: term off)
    float motion_coef;            // motion_reg / (Nd (K-1))  (0: term off)
    const float* dL_dloss;        // device scalar or nullptr (= 1)
    float* dL_dxyz_disp;          // nullptr: value only
    float* dL_dxyz_motion;
    int accumulate_disp, accumulate_motion;   // 1: += into the existing gradient, 0: overwrite
    double* part;                 // 2 * regularizer_blocks() doubles
};
cudaError_t launch_iteration_stats(const IterStatsParams& p, cudaStream_t s);
int regularizer_blocks();
cudaError_t launch_regularizers(RegParams p, float* out2, cudaStream_t s);

cudaError_t launch_subpixel_absmax(const float* subpixel_offset, size_t n, uint32_t* out, cudaStream_t s);
size_t binning_geometry_scratch_bytes(int P);      // sort scratch behind GeometryState::meta ...
size_t binning_geometry_zero_bytes(int P);         // ... and how much of it has to be zero at the start of a forward
size_t binning_status_bytes(int cap, int passes);  // look-back words of `passes` tile passes for a buffer of `cap` instances
int binning_tile_passes(int grid_x, int grid_y);   // 8-bit passes of the tile sort (1 .. 4)
int binning_tile_key_bytes(int grid_x, int grid_y);    // 2, or 4 for images of more than 65535 tiles
int binning_duplicate_set(int grid_x, int grid_y);
// depth order of the visible Gaussians -> g.order, their number -> meta[EX_META_NVIS], the instance count R ->
// meta[EX_META_TOTAL]; returns cudaError
cudaError_t binning_depth_order(const GeometryState& g, int P, cudaStream_t s);
// emit the (tile, id) pairs in depth order (entries at positions >= cap are dropped: the buffer is sized before the
// instance count is known on the host) ...
cudaError_t binning_duplicate(const GeometryState& g, const BinningState& b, const int* radii, int P, int cap,
                              int grid_x, int grid_y, unsigned flags, cudaStream_t s);
// ... stable-sort the min(R, cap) pairs by tile, find the per-tile ranges
cudaError_t binning_sort_ranges(const GeometryState& g, const BinningState& b, const ImageState& img, int P, int cap,
                                int grid_x, int grid_y, unsigned flags, cudaStream_t s);
// clear what duplicate / tile sort accumulate, for a second run within one frame
cudaError_t binning_reset_instances(const GeometryState& g, int P, cudaStream_t s);

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

// ---- exact-output tile culling (EX4DGS_FLAG_TILE_CULL) ---------------------------------------------
// A (Gaussian, tile) instance can be dropped when NO pixel of the tile can pass the `alpha >= 1/255`
// test of the compositing loop, i.e. when the maximum of `power` over the tile's pixel rectangle
// (pixel centres are integers + subpixel offset, so the rectangle is widened by the frame's maximal
// |offset| = pad) is below the splat's skip threshold.  q(d) = 0.5*(A dx^2 + C dy^2) + B dx dy = -power
// is convex, so its minimum over a rectangle that does not contain the centre lies on one of the
// four edges and has a closed form.  Two kernels evaluate this (count in preprocess, emit in
// duplicate) and must agree exactly: every operation is pinned, per-splat invariants are hoisted.
struct CullCtx {
    float cx, cy, A, B, C;
    float nBoC, nBoA;     // -B/C, -B/A
    float shrink;         // rounding guard, see cull_prepare
    float tq;             // cull iff qmin*shrink > tq
    int x0, y0, w;        // tile rectangle origin and width
    uint32_t id;
    int ok;               // 0: never cull (non-convex / ill-conditioned / bad pad)
    int pad0, pad1;
};

__device__ __forceinline__ CullCtx cull_prepare(float cx, float cy, float A, float B, float C, float thr,
                                                int x0, int y0, int w, uint32_t id, float pad)
{
    CullCtx c;
    c.cx = cx; c.cy = cy; c.A = A; c.B = B; c.C = C; c.x0 = x0; c.y0 = y0; c.w = w; c.id = id;
    c.pad0 = c.pad1 = 0;
    const float detc = fa(fm(A, C), -fm(B, B));
    c.ok = (A > 0.f) && (C > 0.f) && (detc > 0.f) && (pad <= 4096.f);
    c.nBoC = __fdiv_rn(-B, C);
    c.nBoA = __fdiv_rn(-B, A);
    // The compositing loop evaluates `power` in float with a handful of roundings on terms that may
    // cancel; the evaluated value is <= -q*(1 - eps*kappa) where kappa = (sqrt(AC)+|B|)^2/(AC-B^2)
    // bounds (sum of |terms|)/q.  eps = 2e-5 is ~50x the real worst case (and also covers the
    // rounding of this bound itself); very ill-conditioned conics are simply not culled.
    const float sAC = fa(__fsqrt_rn(fm(A, C)), fabsf(B));
    const float kappa = __fdiv_rn(fm(sAC, sAC), detc);
    c.shrink = fa(1.0f, -fm(2e-5f, kappa));
    if (!(c.shrink > 0.5f)) c.ok = 0;
    c.tq = fa(2e-3f, -thr);          // -qmin*shrink + 2e-3 < thr  <=>  qmin*shrink > 2e-3 - thr   (thr = exact alpha crossing)
    return c;
}

// true when tile (tx,ty) cannot contribute
__device__ __forceinline__ bool cull_test(const CullCtx& c, int tx, int ty, float pad)
{
    if (!c.ok) return false;
    const float dx0 = fa(fa((float)(tx * EX_TILE), -pad), -c.cx), dx1 = fa(fa((float)(tx * EX_TILE + EX_TILE - 1), pad), -c.cx);
    const float dy0 = fa(fa((float)(ty * EX_TILE), -pad), -c.cy), dy1 = fa(fa((float)(ty * EX_TILE + EX_TILE - 1), pad), -c.cy);
    if (dx0 <= 0.f && dx1 >= 0.f && dy0 <= 0.f && dy1 >= 0.f) return false;   // centre inside
    float qmin = 3.4e38f;
#pragma unroll
    for (int e = 0; e < 2; e++) {
        const float dx = e ? dx1 : dx0;
        const float dy = fminf(dy1, fmaxf(dy0, fm(c.nBoC, dx)));
        const float q = fa(fm(0.5f, fa(fm(fm(c.A, dx), dx), fm(fm(c.C, dy), dy))), fm(fm(c.B, dx), dy));
        qmin = fminf(qmin, q);
    }
#pragma unroll
    for (int e = 0; e < 2; e++) {
        const float dy = e ? dy1 : dy0;
        const float dx = fminf(dx1, fmaxf(dx0, fm(c.nBoA, dy)));
        const float q = fa(fm(0.5f, fa(fm(fm(c.A, dx), dx), fm(fm(c.C, dy), dy))), fm(fm(c.B, dx), dy));
        qmin = fminf(qmin, q);
    }
    return fm(qmin, c.shrink) > c.tq;
}

// Conservative pre-filter in front of cull_test (EX4DGS_FLAG_TILE_CULL only): the reference's tile rectangle
// is the square of half-width ceil(3 sigma_max) around the centre; the pixels that can pass `alpha >= 1/255`
// lie inside the ellipse q(d) <= L with L = tq / shrink (same tq, shrink as cull_prepare), whose axis-aligned
// bounding box has half-extents sqrt(2 L C / det), sqrt(2 L A / det) - much narrower than the square for
// elongated or faint splats.  The rectangle is cut down to the tiles whose pixel-centre range (widened by
// pad) meets that box.  Evaluated by preprocess (tiles_touched, hence the scan and num_rendered) and by the
// duplicate kernel from the same stored floats with pinned operations, so both agree bit for bit; tiles it
// drops would all be dropped by cull_test as well (guards: +1e-4 relative, +1e-3 + 1e-6 |c| absolute).
__device__ __forceinline__ float approx_sqrt(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void tight_rect(float cx, float cy, float A, float B, float C, float thr, float pad,
                                           int& x0, int& y0, int& x1, int& y1)
{
    // MUFU-based reciprocal / square root (2 ulp, deterministic: both call sites execute the same instructions
    // on the same inputs); the 1e-4 relative guard below is 400x their error
    const float detc = fa(fm(A, C), -fm(B, B));
    if (!((A > 0.f) && (C > 0.f) && (detc > 0.f) && (pad <= 4096.f))) return;
    const float rdet = __fdividef(1.0f, detc);
    const float sAC = fa(approx_sqrt(fm(A, C)), fabsf(B));
    const float kappa = fm(fm(sAC, sAC), rdet);
    const float shrink = fa(1.0f, -fm(2.01e-5f, kappa));
    if (!(shrink > 0.5f)) return;
    const float tq = fa(2e-3f, -thr);
    if (!(tq > 0.f)) return;                       // nothing can pass anyway; cull_test sorts it out
    const float twoL = fm(__fdividef(fm(2.0f, tq), shrink), rdet);
    float ex = approx_sqrt(fm(twoL, C));
    float ey = approx_sqrt(fm(twoL, A));
    if (!(ex < 1e6f) || !(ey < 1e6f)) return;      // NaN / huge: keep the reference rectangle
    ex = fa(fa(fm(ex, 1.0001f), 1e-3f), fm(fabsf(cx), 1e-6f));
    ey = fa(fa(fm(ey, 1.0001f), 1e-3f), fm(fabsf(cy), 1e-6f));
    // tile t holds pixel centres in [16 t - pad, 16 t + 15 + pad]
    const float lx = fm(fa(fa(fa(cx, -ex), -pad), -15.0f), 0.0625f), hx = fm(fa(fa(cx, ex), pad), 0.0625f);
    const float ly = fm(fa(fa(fa(cy, -ey), -pad), -15.0f), 0.0625f), hy = fm(fa(fa(cy, ey), pad), 0.0625f);
    // clamp before the float -> int conversion (far off-screen centres)
    const int tx0 = (int)ceilf(fmaxf(lx, -1.0f)), tx1 = (int)floorf(fminf(hx, 70000.0f)) + 1;
    const int ty0 = (int)ceilf(fmaxf(ly, -1.0f)), ty1 = (int)floorf(fminf(hy, 70000.0f)) + 1;
    x0 = max(x0, tx0); x1 = min(x1, tx1);
    y0 = max(y0, ty0); y1 = min(y1, ty1);
    if (x1 < x0) x1 = x0;
    if (y1 < y0) y1 = y0;
}

// Warp-cooperative enumeration of the (lane, tile) items of 32 rectangles: `area` items per lane,
// flattened in (lane, row-major tile) order and dealt 32 at a time to the lanes, so that one huge
// rectangle does not serialise the warp (the per-thread loops of rasterizer_impl.cu:100-111 do).
// s_prefix: 32 ints of the warp (exclusive prefix of area).  Returns the owner lane of `item`.
__device__ __forceinline__ int expand_owner(const int* s_prefix, int item)
{
    int lo = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1)
        if (s_prefix[lo + step] <= item) lo += step;      // last lane with prefix <= item
    return lo;
}

// ---- warp-level culling inside the compositing kernels ------------------------------------------
// Bounding box of the pixel centres a warp works on (8x4 block incl. subpixel offsets).
struct BlockBox {
    float x0, x1, y0, y1;
};

// warp-wide union of per-lane boxes (an unused lane passes x0 = y0 = 3e38, x1 = y1 = -3e38)
__device__ __forceinline__ BlockBox block_box_merge(float x0, float x1, float y0, float y1)
{
    BlockBox b;
    b.x0 = x0; b.x1 = x1; b.y0 = y0; b.y1 = y1;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        b.x0 = fminf(b.x0, __shfl_xor_sync(0xffffffffu, b.x0, d));
        b.x1 = fmaxf(b.x1, __shfl_xor_sync(0xffffffffu, b.x1, d));
        b.y0 = fminf(b.y0, __shfl_xor_sync(0xffffffffu, b.y0, d));
        b.y1 = fmaxf(b.y1, __shfl_xor_sync(0xffffffffu, b.y1, d));
    }
    return b;
}

__device__ __forceinline__ BlockBox block_box(float pxf, float pyf, bool use)
{
    BlockBox b;
    b.x0 = use ? pxf : 3.0e38f;  b.x1 = use ? pxf : -3.0e38f;
    b.y0 = use ? pyf : 3.0e38f;  b.y1 = use ? pyf : -3.0e38f;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        b.x0 = fminf(b.x0, __shfl_xor_sync(0xffffffffu, b.x0, d));
        b.x1 = fmaxf(b.x1, __shfl_xor_sync(0xffffffffu, b.x1, d));
        b.y0 = fminf(b.y0, __shfl_xor_sync(0xffffffffu, b.y0, d));
        b.y1 = fmaxf(b.y1, __shfl_xor_sync(0xffffffffu, b.y1, d));
    }
    return b;
}

// true when the splat (record words a = (x, y, depth, thr), b = (A, B, C, opacity)) cannot reach
// alpha >= 1/255 at any pixel centre inside `box`.  q = -power >= 0.5*(det/C)*dx^2 and
// >= 0.5*(det/A)*dy^2 (completing the square), so the alpha >= 1/255 region lies inside the
// axis-aligned box |dx| <= sqrt(2 t C/det), |dy| <= sqrt(2 t A/det) with t = -thr.  Conservative by
// construction; the guard `shrink` covers the float rounding of the per-pixel power exactly like
// cull_prepare (kappa is bounded above by 4AC/det), ill-conditioned or non-convex conics are kept.
__device__ __forceinline__ bool block_reject(const float4& a, const float4& b, const BlockBox& box)
{
    const float A = b.x, B = b.y, C = b.z;
    const float det = A * C - B * B;
    if (!(A > 0.f) || !(C > 0.f) || !(det > 0.f)) return false;
    const float inv = 1.0f / det;
    const float shrink = 1.0f - 8e-5f * (A * C * inv);
    if (!(shrink > 0.5f)) return false;
    const float tq = 2e-3f - a.w;                       // NaN thr -> comparisons below are false -> keep
    const float lim = 2.0f * tq * inv / shrink * 1.0001f;
    const float dx = fmaxf(fmaxf(box.x0 - a.x, a.x - box.x1), 0.f);
    const float dy = fmaxf(fmaxf(box.y0 - a.y, a.y - box.y1), 0.f);
    return (dx * dx > lim * C) || (dy * dy > lim * A);
}

// ---- packed FP32 pairs (sm_100: fma/mul/add.rn.f32x2 -> SASS FFMA2 / FMUL2 / FADD2) ------------------------------
// Two IEEE round-to-nearest operations per issued instruction; each half rounds exactly like the scalar
// __fmaf_rn / __fmul_rn / __fadd_rn, so pinned arithmetic stays bit-exact.  A scalar operand is written bc(x): the
// SASS forms take a 32-bit register (or immediate) as a broadcast operand (`R10.F32`), so bc() costs nothing.
// The compositing kernels pair the TWO PIXELS a lane owns.
struct f2 {
    unsigned long long v;
};
__device__ __forceinline__ f2 mk2(float lo, float hi)
{
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f2 bc(float s) { return mk2(s, s); }
__device__ __forceinline__ void split2(f2 a, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
}
__device__ __forceinline__ float lo2(f2 a) { float x, y; split2(a, x, y); return x; }
__device__ __forceinline__ float hi2(f2 a) { float x, y; split2(a, x, y); return y; }
__device__ __forceinline__ f2 fm2(f2 a, f2 b)
{
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 fa2(f2 a, f2 b)
{
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 ff2(f2 a, f2 b, f2 c)
{
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
__device__ __forceinline__ float hsum2(f2 a) { float x, y; split2(a, x, y); return x + y; }

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier transaction counting -------------------
// The compositing kernels stage each splat's 64-byte record with ONE bulk copy issued by the
// thread that owns the slot; completion is tracked by an mbarrier per ring buffer whose expected
// transaction bytes are announced by thread 0 (the tx-count may run ahead of the expectation).
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// per-thread 16-byte asynchronous copy global -> shared (LDGSTS), tracked by cp.async groups
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// 16-byte shared-memory load from a 32-bit shared address (keeps the per-buffer base in one register:
// through a generic pointer the compiler re-derives the shared window base in every loop iteration)
__device__ __forceinline__ float4 lds128(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// TMA gather: rows i0..i3 (all 16 floats = 64 bytes each) of the 2-D tensor `tmap` ([P,16] float32, box {16,1}) land densely at
// smem_dst (128-byte aligned): 256 bytes per instruction, completion counted on `bar` (probed in tools/tma_gather4_probe.cu)
__device__ __forceinline__ void tma_gather4_g2s(void* smem_dst, const CUtensorMap* tmap, int i0, int i1, int i2, int i3,
                                                unsigned long long* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n"
                 ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(0), "r"(i0), "r"(i1), "r"(i2), "r"(i3), "r"(smem_u32(bar)) : "memory");
}

// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

