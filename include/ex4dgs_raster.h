/*
 * ex4dgs_raster.h - C ABI of the B200-native differentiable 4D-Gaussian rasterizer.
 *
 * Drop-in boundary for the hot path of juno181/Ex4DGS: every entry point below replaces one
 * method of the reference's native interface `CudaRasterizer::Rasterizer`
 * (submodules/diff_gaussian_rasterization_df/cuda_rasterizer/rasterizer.h:23-107), which the
 * reference binds to Python through rasterize_points.cu:36-259 + ext.cpp:15-19.
 *
 * Conventions
 *   - plain C: pointers, sizes and scalars only; no torch / C++ types cross this boundary.
 *   - every pointer is a DEVICE pointer (float32 unless stated) valid on the current CUDA device,
 *     except callbacks / `user` cookies / the stream handle.
 *   - "absent" optional inputs are passed as NULL (the reference passes the data_ptr of a CPU
 *     0-element tensor, which is nullptr: rasterize_points.cu:103-130, forward.cu:218,254).
 *   - matrices are 16 floats in the reference's transposed storage (auxiliary.h:68-87).
 *   - all work is enqueued on `stream` (a cudaStream_t cast to void*; NULL = legacy default stream,
 *     which is what the reference uses).  ex4dgs_forward reads num_rendered back like
 *     rasterizer_impl.cu:299, but waits for it only after the whole frame is queued (the instance
 *     count stays on the device for the kernels); with EX4DGS_FLAG_NO_HOST_WAIT it does not wait at all.
 *   - functions return a negative ex4dgs_status on failure; ex4dgs_last_error() has the message.
 *   - the library is re-entrant per device/stream: no global mutable state besides the
 *     thread-local error string.
 */
#ifndef EX4DGS_RASTER_H_
#define EX4DGS_RASTER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EX4DGS_ABI_VERSION 1

typedef enum ex4dgs_status {
    EX4DGS_OK = 0,
    EX4DGS_ERR_INVALID = -1,   /* bad argument combination (mirrors the Python-side Exceptions)   */
    EX4DGS_ERR_CUDA = -2,      /* CUDA runtime error (message carries cudaGetErrorString)          */
    EX4DGS_ERR_ALLOC = -3,     /* an allocator callback returned NULL                              */
    EX4DGS_ERR_UNSUPPORTED = -4
} ex4dgs_status;

/* flags for ex4dgs_forward / ex4dgs_backward (must be identical in both calls of one frame) */
#define EX4DGS_FLAG_NONE 0u
/* Drop (Gaussian,tile) instances that provably cannot reach alpha >= 1/255 in any pixel of the
 * tile.  Outputs (image, depth, acc, flow, idx, radii, all gradients) are unchanged; only the
 * internal tile lists get shorter.  With the flag clear the tile lists (point_list, ranges,
 * n_contrib) are bit-identical to the reference's (rasterizer_impl.cu:72-140). */
#define EX4DGS_FLAG_TILE_CULL 1u
/* Segmented SH input (fused model front-end): the model keeps its SH coefficients in
 * four tensors - features_dc [Ns,1,3] / features_rest [Ns,15,3] for the static Gaussians and the same
 * pair for the dynamic ones - which CGaussianModel.get_features() concatenates into [P,16,3] on every
 * frame (scene/c_gaussian_model.py:337-353: ~1.5 GB of copies at 2 M Gaussians, and as much again in
 * the backward).  With this flag the `shs` argument of ex4dgs_forward / ex4dgs_backward is not a
 * device array but a HOST pointer to an ex4dgs_sh_segments (cast to const float*), the kernels read
 * the four tensors in place, and `dL_dsh` of ex4dgs_backward is likewise a host pointer to an
 * ex4dgs_sh_segments holding the four gradient outputs (every element written).  M must be 16.
 * Gaussians [0, n_static) are the static ones (static first, as the reference concatenates). */
#define EX4DGS_FLAG_SH_SEGMENTED 2u
/* Forward without any host wait (CUDA-graph capturable): the binning buffer is sized from the capacity hint
 * (ex4dgs_set_capacity_hint, or the largest instance count this thread has seen + 25 %; EX4DGS_ERR_INVALID without
 * one), num_rendered is NOT read back: the return value is that capacity - an upper bound of the instance count,
 * which ex4dgs_backward accepts as R.  If the frame has more instances than the capacity, the tile lists are
 * truncated and bit 1 of the device word `meta[5]` of the geometry buffer (ex4dgs_describe_buffers: "meta") is
 * set - check it whenever the host synchronises anyway (e.g. with the loss) and redo the frame without the flag. */
#define EX4DGS_FLAG_NO_HOST_WAIT 4u
typedef struct ex4dgs_sh_segments {
    int n_static;
    float* dc_static;      /* [n_static, 1, 3]        */
    float* rest_static;    /* [n_static, 15, 3]       */
    float* dc_dynamic;     /* [P - n_static, 1, 3]    */
    float* rest_dynamic;   /* [P - n_static, 15, 3]   */
} ex4dgs_sh_segments;

/* Resizable scratch buffers, the C form of `std::function<char*(size_t)>` in
 * rasterizer.h:38-40 / rasterize_points.cu:27-33: called once per buffer per forward (the binning
 * allocator may be called a second time when a size estimate proved too small - the last
 * returned pointer is the one in use);
 * must return a device pointer to at least `nbytes` bytes, 256-byte aligned, that stays valid
 * until the matching ex4dgs_backward (the Python layer keeps the torch byte tensors alive
 * through autograd exactly like ctx.save_for_backward in __init__.py:106). */
typedef void* (*ex4dgs_alloc_fn)(void* user, size_t nbytes);

/* ---- replaces CudaRasterizer::Rasterizer::forward (rasterizer.h:37-66, rasterizer_impl.cu:204-363)
 * Returns num_rendered (R >= 0) = number of (Gaussian,tile) instances in the sorted tile lists,
 * or a negative ex4dgs_status.
 *   P            number of Gaussians;  D active SH degree (0..3);  M SH coefficients per Gaussian
 *   background   [3]           means3D [P,3]      dir3D [P,3] (must be non-NULL when P > 0)
 *   shs          [P,M,3] or NULL (then colors_precomp [P,3] must be given)
 *   opacities    [P]           scales [P,3] + rotations [P,4]  xor  cov3D_precomp [P,6]
 *   subpixel_offset [H,W,2]
 *   outputs (every element is written by the call, no pre-fill needed):
 *     out_color [3,H,W], out_depth [H,W], out_acc [H,W], out_flow [3,H,W], out_idx [H,W] (int32,
 *     -1 = no contributor), radii [P] (int32, 0 = culled)
 */
/* Sizing hint for the calling thread's next forwards: expected number of (Gaussian, tile) instances.  The binning
 * buffer is requested with 25 % on top before the count of the frame is known; 0 forgets the history (the next forward
 * then waits for the count before it sizes the buffer, like the reference). */
void ex4dgs_set_capacity_hint(int instances);

int ex4dgs_forward(
    ex4dgs_alloc_fn geometryBuffer, void* geometry_user,
    ex4dgs_alloc_fn binningBuffer, void* binning_user,
    ex4dgs_alloc_fn imageBuffer, void* image_user,
    int P, int D, int M,
    const float* background,
    int width, int height,
    const float* means3D,
    const float* dir3D,
    const float* shs,
    const float* colors_precomp,
    const float* opacities,
    const float* scales,
    float scale_modifier,
    const float* rotations,
    const float* cov3D_precomp,
    const float* viewmatrix,
    const float* projmatrix,
    const float* cam_pos,
    float tan_fovx, float tan_fovy,
    float kernel_size,
    const float* subpixel_offset,
    int prefiltered,
    float* out_color,
    float min_depth,
    float max_depth,
    float* out_depth,
    float* out_acc,
    float* out_flow,
    int* out_idx,
    int* radii,
    int debug,
    unsigned flags,
    void* stream);

/* ---- replaces CudaRasterizer::Rasterizer::backward (rasterizer.h:68-106, rasterizer_impl.cu:367-486)
 * geom/binning/image buffers are the ones handed out by the allocator callbacks of the matching
 * forward; R is its return value.  Gradient outputs (all fully written, no pre-fill needed):
 *   dL_dmean2D [P,3]  dL_dopacity [P]  dL_dcolor [P,3]  dL_dmean3D [P,3]  dL_dsh [P,M,3] (NULL if M==0)
 *   dL_dscale [P,3]  dL_drot [P,4] (NULL when cov3D_precomp is used)  dL_dcov3D [P,6] (may be NULL)
 *   dL_ddir [P,3]
 * The reference's intermediate dL_dconic [P,2,2] (rasterize_points.cu:181) lives inside the
 * scratch buffers here.  The backward reproduces the reference's deviations from the true
 * derivative.  Returns EX4DGS_OK or a negative status. */
int ex4dgs_backward(
    int P, int D, int M, int R,
    const float* background,
    int width, int height,
    const float* means3D,
    const float* shs,
    const float* colors_precomp,
    const float* scales,
    float scale_modifier,
    const float* rotations,
    const float* acc_depth,      /* forward out_depth */
    const float* acc,            /* forward out_acc   */
    float min_depth,
    float max_depth,
    const float* cov3D_precomp,
    const float* viewmatrix,
    const float* projmatrix,
    const float* campos,
    float tan_fovx, float tan_fovy,
    float kernel_size,
    const float* subpixel_offset,
    const int* radii,
    void* geom_buffer,
    void* binning_buffer,
    void* image_buffer,
    const float* dL_dpix,        /* [3,H,W] */
    const float* dL_ddepth,      /* [H,W]   or NULL = no upstream gradient (read as zeros, terms skipped) */
    const float* dL_dflow,       /* [3,H,W] or NULL */
    const float* dL_dacc,        /* [H,W]   or NULL */
    float* dL_dmean2D,
    float* dL_dopacity,
    float* dL_dcolor,
    float* dL_dmean3D,
    float* dL_dcov3D,
    float* dL_dsh,
    float* dL_dscale,
    float* dL_drot,
    float* dL_ddir,
    int debug,
    unsigned flags,
    void* stream);

/* ---- replaces CudaRasterizer::Rasterizer::markVisible (rasterizer.h:27-35, rasterizer_impl.cu:143-159)
 * present [P] bytes (0/1) = in_frustum (auxiliary.h:267-294). */
int ex4dgs_mark_visible(
    int P,
    const float* means3D,
    const float* viewmatrix,
    const float* projmatrix,
    float min_depth,
    float max_depth,
    uint8_t* present,
    void* stream);

/* ---- fused model front-end ("next" row N1 of SURVEY.md section 8f): the per-frame getters of
 * CGaussianModel (scene/c_gaussian_model.py:170-215,330-375; utils/interpolations.py:33-61,81-93)
 * evaluated in one pass straight from the model's native static / dynamic tensors into the flat
 * [P,.] rasterizer inputs (static Gaussians first), replacing ~40 PyTorch kernels and five
 * torch.cat copies per frame.
 *   static : xyz [Ns,3], xyz_disp [Ns,3], rotation [Ns,4] (raw), scaling [Ns,3] (log), opacity [Ns] (logit)
 *   dynamic: xyz_motion [Nd,K,3], rotation_motion [Nd,K,4], scaling_motion [Nd,3] (log),
 *            opacity_motion [Nd] (logit), opacity_center [Nd,2], opacity_var [Nd,2]
 *   t timestamp; duration, interval, time_shift, var_min (= var_pad/interval) as in the model (doubles:
 *   Python floats).
 *   outputs: means3D [P,3], rotations [P,4], scales [P,3], opacities [P]   (P = Ns + Nd)
 */
int ex4dgs_frontend_forward(
    int Ns, int Nd, int K,
    const float* xyz, const float* xyz_disp, const float* rotation,
    const float* scaling, const float* opacity,
    const float* xyz_motion, const float* rotation_motion, const float* scaling_motion,
    const float* opacity_motion, const float* opacity_center, const float* opacity_var,
    double t, double duration, double interval, double time_shift, double var_min,
    float* means3D, float* rotations, float* scales, float* opacities,
    void* stream);

/* Backward of ex4dgs_frontend_forward.  Gradient outputs for the keyframe tensors
 * (dL_dxyz_motion [Nd,K,3], dL_drotation_motion [Nd,K,4]) are fully written (zeros outside the
 * 4 / 2 keyframes that the frame touches).  The timing scalars are doubles so that the host
 * reproduces the Python arithmetic of c_gaussian_model.py:184-187 exactly. */
int ex4dgs_frontend_backward(
    int Ns, int Nd, int K,
    const float* rotation_motion, const float* scaling, const float* opacity,
    const float* scaling_motion, const float* opacity_motion,
    const float* opacity_center, const float* opacity_var,
    double t, double duration, double interval, double time_shift, double var_min,
    const float* dL_dmeans3D, const float* dL_drotations, const float* dL_dscales, const float* dL_dopacities,
    float* dL_dxyz, float* dL_dxyz_disp, float* dL_drotation, float* dL_dscaling, float* dL_dopacity,
    float* dL_dxyz_motion, float* dL_drotation_motion, float* dL_dscaling_motion,
    float* dL_dopacity_motion, float* dL_dopacity_center, float* dL_dopacity_var,
    void* stream);

/* ---- photometric loss -------------------------------------------
 * Replaces, for one rendered frame, the loss block of the reference's training step
 * (train.py:144-151 with utils/loss_utils.py:22-25 l1_loss and :33-81 ssim/_ssim):
 *     Ll1   = mean |image - gt_image|
 *     ssim  = mean ssim_map(image, gt_image)      11x11 Gaussian window, sigma 1.5, zero padding 5
 *     loss  = (1 - lambda_dssim) * Ll1 + lambda_dssim * (1 - ssim)
 *     l1_errors   [H,W] = mean over channels of |image - gt_image|          (train.py:149)
 *     ssim_errors [H,W] = mean over channels of ssim_map                    (train.py:150)
 * image, gt_image: device float [3,H,W].  out_loss3: device float[3] = {loss, Ll1, ssim}.
 * scratch: device bytes, ex4dgs_loss_scratch_bytes(width, height), kept by the caller between the
 * forward and the backward of the same frame.  Sums are reduced in a fixed order (bit-reproducible). */
size_t ex4dgs_loss_scratch_bytes(int width, int height);
int ex4dgs_loss_forward(int width, int height, const float* image, const float* gt_image, float lambda_dssim,
                        char* scratch, float* out_loss3, float* l1_errors, float* ssim_errors, void* stream);
/* dL_dloss: device scalar (the gradient arriving at `loss`); dL_dimage: device float [3,H,W], overwritten. */
int ex4dgs_loss_backward(int width, int height, const float* image, const float* gt_image, float lambda_dssim,
                         const char* scratch, const float* dL_dloss, float* dL_dimage, void* stream);

/* Plain L1 loss, utils/loss_utils.py:22-25 `torch.abs((network_output - gt)).mean()` over n floats: value
 * (device float) and, in the backward, dL_da = sgn(a - b) * dL_dloss / n.  scratch: ex4dgs_l1_scratch_bytes()
 * device bytes.  Fixed-order double accumulation (bit-reproducible). */
size_t ex4dgs_l1_scratch_bytes(void);
int ex4dgs_l1_forward(size_t n, const float* a, const float* b, char* scratch, float* out_loss, void* stream);
int ex4dgs_l1_backward(size_t n, const float* a, const float* b, const float* dL_dloss, float* dL_da, void* stream);

/* ---- optimizer step ------------------------------------------------
 * Replaces `gaussians.optimizer.step()` of the reference's training iteration (train.py:250): the
 * optimizer is torch.optim.RAdam(l, lr=0.001) over 15 single-tensor parameter groups with per-group
 * learning rates (scene/c_gaussian_model.py:430-449; betas (0.9, 0.999), eps 1e-8, weight_decay 0).
 * One launch updates every tensor of the table: param/exp_avg/exp_avg_sq in place, following the
 * arithmetic of torch's CUDA ("foreach") RAdam operation by operation.
 *   step        the tensor's step count AFTER this update (state['step'] + 1), >= 1
 *   grad_scale  multiplies every gradient first (1/world_size after a SUM all-reduce; 1 otherwise)
 * At most EX4DGS_RADAM_MAX_TENSORS tensors per call.  All pointers: device float, `numel` elements. */
#define EX4DGS_RADAM_MAX_TENSORS 32
typedef struct ex4dgs_radam_tensor {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    size_t numel;
    double lr;
    long long step;
} ex4dgs_radam_tensor;
int ex4dgs_radam_step(const ex4dgs_radam_tensor* tensors, int n, double beta1, double beta2, double eps,
                      double grad_scale, void* stream);
/* Same step with the two per-iteration guards of train.py:244-253 folded in (bit i of a mask = tensors[i]):
 *   sanitize_grad_mask  the gradient is passed through torch.nan_to_num first (NaN -> 0, +-inf -> +-FLT_MAX;
 *                       train.py:246-248 does it for _opacity_duration_var.grad)
 *   check_nan_mask      nan_flags[i] (device int[n], caller-zeroed, only ever set to 1) reports that tensor i
 *                       received a NaN parameter - what CGaussianModel.prune_nan_points
 *                       (scene/c_gaussian_model.py:1229-1241) finds out with isnan().any() reductions and a host
 *                       wait for _xyz and _xyz_motion after every step. */
int ex4dgs_radam_step_ex(const ex4dgs_radam_tensor* tensors, int n, double beta1, double beta2, double eps,
                         double grad_scale, unsigned check_nan_mask, unsigned sanitize_grad_mask, int* nan_flags,
                         void* stream);
/* Host-only helper (no GPU needed): the two per-tensor scalars the kernel receives for step `step`,
 * p += m * (rectified ? 1 / ((sqrt(v) + eps) / S) : U)  - exposed so the CPU tests can pin them against
 * torch/optim/radam.py's expressions. */
int ex4dgs_radam_scalars(double lr, long long step, double beta1, double beta2, float* S, float* U, int* rectified);

/* ---- tensor surgery of densification / pruning ----------------------
 * Replaces, in ONE launch for all tensors, the per-tensor boolean-mask gathers and concatenations the reference
 * performs on its 15 parameter tensors, both optimizer moments of each and the per-Gaussian statistics:
 *   CGaussianModel._prune_optimizer / prune_points       (scene/c_gaussian_model.py:693-763)   x[mask]
 *   CGaussianModel.cat_tensors_to_optimizer              (:765-787)                            torch.cat((x, extension)), zero moments
 *   densify_and_clone / densify_and_split                (:874-1017)                           extension = x[selected](.repeat(N))
 * One job copies rows of `row_bytes` bytes (a multiple of 4):
 *   dst[r] = r < n_a ? a[index ? index[r] : r] : (b ? b[r - n_a] : 0)        for r in [0, n_out)
 * (pruning: index = the kept rows, n_a = n_out; concatenation: index NULL, n_a = rows of a, b = the extension or NULL
 * for zero rows; cloning: index = [0..n) ++ selected, n_a = n_out for parameters, n_a = n for the moments).
 * All pointers are device pointers, `index` is int64 as torch.nonzero produces it and is not range-checked;
 * dst must not overlap a or b.  At most EX4DGS_GATHER_MAX_JOBS jobs per call. */
#define EX4DGS_GATHER_MAX_JOBS 64
typedef struct ex4dgs_gather_job {
    const void* a;
    const void* b;
    void* dst;
    const long long* index;
    size_t row_bytes;
    long long n_a;
    long long n_out;
} ex4dgs_gather_job;
int ex4dgs_gather_rows(const ex4dgs_gather_job* jobs, int n, void* stream);

/* ---- per-iteration statistics (SURVEY.md section 8, row N4 "stats updates") -----------------------
 * Replaces, in ONE launch without host synchronisation, the bookkeeping train.py:196-215 performs after
 * loss.backward() through boolean-mask indexing (~60 PyTorch kernels, a nonzero() + host wait per mask):
 *   CGaussianModel.mark_prune_stats      (scene/c_gaussian_model.py:1105-1117)   when grad_error != NULL
 *   max_radii2D / motion_max_radii2D     (train.py:205-206)                      when densify
 *   CGaussianModel.add_densification_stats (:1095-1103)                          when densify
 *   CGaussianModel.add_l1_ssim_stats       (:1119-1145)                          when densify && grad_error
 * for the Ns static Gaussians (arrays `stat`) followed by the Nd dynamic ones (arrays `dyn`; the model's
 * motion_* tensors, `error_accum` being motion_xyz_error_mean).  All arrays: device float, one element per
 * Gaussian, updated in place.
 *   radii         [Ns+Nd] int32, the rasterizer's output (visibility_filter = radii > 0)
 *   grad_means2D  [Ns+Nd,3] = viewspace_point_tensor.grad
 *   grad_error    [Ns+Nd,3] = viewspace_point_error_tensor.grad (acc, l1, ssim back-projections), or NULL
 *                 when opt.l1_accum is off
 *   timestamp     the camera's timestamp (stored into *_error_min_timestamp)
 *   densify       iteration < opt.densify_until_iter */
typedef struct ex4dgs_stats_arrays {
    float* max_radii2D;
    float* min_radii2D;
    float* xyz_gradient_accum;
    float* denom;
    float* error_accum;
    float* error_min;
    float* error_min_timestamp;
    float* ssim_error_accum;
    float* error_denom;
} ex4dgs_stats_arrays;
int ex4dgs_iteration_stats(int Ns, int Nd, const int* radii, const float* grad_means2D, const float* grad_error,
                           float timestamp, int densify,
                           const ex4dgs_stats_arrays* stat, const ex4dgs_stats_arrays* dyn, void* stream);

/* ---- regularisation terms of the loss (train.py:156-162) ------------------------------------------
 *   out_terms[0] = static_reg * mean_i log(|xyz_disp[i]| + 0.001)                     (Ns > 0, static_reg != 0)
 *   out_terms[1] = motion_reg * mean_{i,k>=1} |xyz_motion[i,0] - xyz_motion[i,k]|     (Nd > 0, K > 1, motion_reg != 0)
 * and their gradients times *dL_dloss (device scalar; NULL = 1), written to (accumulate = 0) or added in
 * place to (accumulate = 1, the gradient the backward pass has already produced) dL_dxyz_disp [Ns,3] /
 * dL_dxyz_motion [Nd,K,3]; a NULL gradient pointer computes the value only.  A term whose weight is 0 is
 * skipped (train.py tests `opt.*_reg > 0`) and reports 0.  scratch: ex4dgs_regularizer_scratch_bytes()
 * device bytes.  Sums are reduced in a fixed order (bit-reproducible). */
size_t ex4dgs_regularizer_scratch_bytes(void);
int ex4dgs_regularizers(int Ns, int Nd, int K, const float* xyz_disp, const float* xyz_motion,
                        float static_reg, float motion_reg, const float* dL_dloss,
                        float* dL_dxyz_disp, int accumulate_disp, float* dL_dxyz_motion, int accumulate_motion,
                        float* out_terms, char* scratch, void* stream);

/* ---- introspection (used by the parity tests to look inside the opaque scratch buffers) ------- */
typedef struct ex4dgs_array_desc {
    const char* name;    /* e.g. "point_list"                                  */
    int buffer;          /* 0 = geometry, 1 = binning, 2 = image               */
    size_t offset;       /* byte offset from the (256-B aligned) buffer base   */
    size_t elem_size;    /* bytes per element                                  */
    size_t count;        /* number of elements                                 */
} ex4dgs_array_desc;

/* Fills `out` (capacity `max`) with the layout the library uses for (P, R, width, height);
 * returns the number of arrays described. */
int ex4dgs_describe_buffers(int P, int R, int width, int height,
                            ex4dgs_array_desc* out, int max);

/* Size in bytes each allocator callback will be asked for. */
size_t ex4dgs_geometry_bytes(int P);
size_t ex4dgs_binning_bytes(int R);
size_t ex4dgs_image_bytes(int width, int height);

/* ---- measurement hooks (bench.py; not part of the reference interface) -------------------------
 * While profiling is on (process-wide: autograd runs the backward on its own thread), every
 * ex4dgs_forward / ex4dgs_backward records
 * CUDA events at its stage boundaries on the launching stream.  ex4dgs_profile_read synchronises
 * those events, ADDS the elapsed milliseconds per stage into ms[0..EX4DGS_NUM_STAGES) and the
 * number of frames into *frames, and clears the recorded set.  Stages:
 *   0 preprocess fwd   1 depth sort + scan   2 duplicate + tile sort + ranges   3 render fwd
 *   4 render bwd (incl. accumulator memset)   5 preprocess bwd */
#define EX4DGS_NUM_STAGES 6
void ex4dgs_profile_enable(int on);
int ex4dgs_profile_read(double* ms, int* frames_fwd, int* frames_bwd);
/* number of kernels of this library (not CUB's) launched by this process so far */
unsigned long long ex4dgs_launch_count(void);

int ex4dgs_abi_version(void);
const char* ex4dgs_last_error(void);
/* Diagnostics of the calling thread's last ex4dgs_forward: the number of visible Gaussians whose exact alpha >= 1/255
 * threshold (the float thr with  min(0.99, opacity * expf(power)) >= 1/255  <=>  power >= thr, csrc/preprocess.cu
 * alpha_threshold) could not be established because expf was not monotone on the nine floats around it, so that the
 * conservative threshold was stored instead.  0 on everything observed so far; when non-zero, the backward may treat
 * pairs of those Gaussians within 1e-3 of the limit as contributing although the forward (forward.cu:384-387) skipped
 * them. */
unsigned ex4dgs_last_inexact_thresholds(void);
/* Work decomposition of the forward compositing kernel, needed to read the per-tile statistics word `tile_batches`
 * (ex4dgs_describe_buffers): low 8 bits = batches of *batch splats the tile fetched before all its pixels were done
 * (R_eff of SURVEY.md 8d = sum over tiles of min(range length, batch * batches)), high 24 bits = (warp, splat) pairs
 * that survived the per-warp block test, out of *warps warps per tile. */
void ex4dgs_forward_geometry(int* batch, int* warps);

#ifdef __cplusplus
}
#endif
#endif /* EX4DGS_RASTER_H_ */
