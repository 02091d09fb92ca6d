// C ABI of the rasterizer (include/ex4dgs_raster.h): host orchestration of the forward and
// backward pipelines.  Mirrors the stage order of the reference's host code
// (cuda_rasterizer/rasterizer_impl.cu:204-363 forward, :367-486 backward; rasterize_points.cu
// for argument handling) on a caller-supplied stream, with caller-supplied scratch allocators.
#include "../../include/ex4dgs_raster.h"
#include "common.cuh"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <vector>

namespace {

thread_local char g_err[512] = "";
thread_local int g_last_R = 0;     // sizing hint only (previous frame of this thread)
thread_local bool g_last_flow = false;   // did the last frame of this thread have flow (non-zero dir3D)? (which compositing instantiation to queue before the flag is read)
thread_local int g_last_cap = 0;   // capacity of the binning buffer of this thread's last forward (ex4dgs_describe_buffers)
thread_local unsigned g_last_inexact = 0;   // Gaussians of the last forward whose alpha threshold fell back (preprocess.cu)

// Pinned landing pad of the forward's one read-back (R and the flow flag), one per host thread and
// device: with pinned memory both 4-byte copies are truly asynchronous and cost a single sync.
struct Readback {
    uint32_t* host = nullptr;
    int device = -1;
    cudaEvent_t ev = nullptr;
    // (re)creates the pinned words and the event when the calling thread has moved to another device
    uint32_t* get()
    {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
        if (host == nullptr || dev != device) {
            if (host) cudaFreeHost(host);
            host = nullptr;
            ev = nullptr;        // an event of the previous device stays with that device (never destroyed: the context may be gone)
            if (cudaHostAlloc(reinterpret_cast<void**>(&host), 64, cudaHostAllocDefault) != cudaSuccess) host = nullptr;
            if (host) memset(host, 0, 64);
            if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { ev = nullptr; cudaGetLastError(); }
            device = dev;
        }
        return host;
    }
    // event recorded right after the read-back copies: the host waits for it instead of the whole stream, so
    // that work queued behind the copies (the speculative duplicate kernel) runs while the host wakes up
    cudaEvent_t event() { return ev; }
    ~Readback() { if (host) cudaFreeHost(host); }
};
thread_local Readback g_readback;
// Measurement state is process-wide: autograd runs the backward on its own thread.
std::atomic<bool> g_prof{false};
std::atomic<unsigned long long> g_launches{0};
struct ProfFrame { cudaEvent_t ev[5]; int n; bool bwd; };
std::mutex g_prof_mu;
std::vector<ProfFrame> g_frames;
std::vector<cudaEvent_t> g_event_pool;

cudaEvent_t pool_get()
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_event_pool.empty()) {
        cudaEvent_t e = g_event_pool.back();
        g_event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

struct Prof {
    ProfFrame f;
    bool on;
    cudaStream_t s;
    Prof(bool bwd, cudaStream_t st) : on(g_prof.load()), s(st) { f.n = 0; f.bwd = bwd; }
    void mark()
    {
        if (!on || f.n >= 5) return;
        f.ev[f.n] = pool_get();
        cudaEventRecord(f.ev[f.n], s);
        f.n++;
    }
    void done()
    {
        if (!on) return;
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_frames.push_back(f);
    }
};

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CK(expr)                                                                                     \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) return fail(EX4DGS_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

// with `debug` the reference synchronises and checks after every stage (auxiliary.h:296-303)
#define STAGE(debug, s, what)                                                                        \
    do {                                                                                             \
        cudaError_t _e = cudaGetLastError();                                                         \
        if (_e == cudaSuccess && (debug)) _e = cudaStreamSynchronize(s);                             \
        if (_e != cudaSuccess) return fail(EX4DGS_ERR_CUDA, "%s: %s", what, cudaGetErrorString(_e));  \
    } while (0)

}  // namespace

GeometryState carve_geometry(void* base, int P, size_t temp_bytes)
{
    GeometryState g;
    Carver c(base);
    const size_t n = (size_t)P;
    g.key_in = c.take<uint32_t>(n);
    g.key_a = c.take<uint32_t>(n);
    g.val_a = c.take<uint32_t>(n);
    g.key_b = c.take<uint32_t>(n);
    g.order = c.take<uint32_t>(n);
    g.tiles_touched = c.take<uint32_t>(n);
    g.rec = c.take<SplatRec>(n);
    g.clamped = c.take<uint8_t>(n);
    g.gacc = c.take<GradAcc>(n);
    g.meta = c.take<uint32_t>(64);
    g.temp = c.take<char>(temp_bytes);
    g.temp_bytes = temp_bytes;
    g.total = c.off + 256;
    return g;
}

// Layout of the binning buffer: two sets of (ids, tiles) of `cap` entries each (BinningState), then the look-back
// words of the tile sort.  The sorted id list (point_list, all the backward needs) is at offset 0 whatever `cap` is,
// so a buffer sized for a guess cap >= R can be filled and sorted before R is known on the host.
BinningState carve_binning(void* base, int cap, size_t status_bytes, int key_bytes)
{
    BinningState b;
    const size_t m = (size_t)(cap > 0 ? cap : 0);
    Carver c(base);
    b.key_bytes = key_bytes;
    b.val[0] = c.take<uint32_t>(m);
    b.tile[0] = c.take<char>(m * (size_t)key_bytes);
    b.val[1] = c.take<uint32_t>(m);
    b.tile[1] = c.take<char>(m * (size_t)key_bytes);
    b.status = reinterpret_cast<uint32_t*>(c.take<char>(status_bytes));
    b.total = c.off + 256;
    return b;
}

ImageState carve_image(void* base, int width, int height)
{
    ImageState im;
    Carver c(base);
    const size_t n = (size_t)width * height;
    const size_t tiles = (size_t)((width + EX_TILE - 1) / EX_TILE) * ((height + EX_TILE - 1) / EX_TILE);
    im.final_T = c.take<float>(n);
    im.n_contrib = c.take<uint32_t>(n);
    im.ranges = c.take<uint2>(tiles);
    im.tile_batches = c.take<uint32_t>(tiles);
    im.total = c.off + 256;
    return im;
}

extern "C" {

int ex4dgs_abi_version(void) { return EX4DGS_ABI_VERSION; }

void ex4dgs_profile_enable(int on) { g_prof.store(on != 0); }
unsigned long long ex4dgs_launch_count(void) { return g_launches.load(); }

int ex4dgs_profile_read(double* ms, int* frames_fwd, int* frames_bwd)
{
    int nf = 0, nb = 0;
    std::vector<ProfFrame> frames;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        frames.swap(g_frames);
    }
    for (ProfFrame& f : frames) {
        if (f.n > 0) cudaEventSynchronize(f.ev[f.n - 1]);
        for (int i = 0; i + 1 < f.n; i++) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, f.ev[i], f.ev[i + 1]) == cudaSuccess && ms) ms[(f.bwd ? 4 : 0) + i] += t;
        }
        {
            std::lock_guard<std::mutex> lk(g_prof_mu);
            for (int i = 0; i < f.n; i++) g_event_pool.push_back(f.ev[i]);
        }
        if (f.bwd) nb++; else nf++;
    }
    if (frames_fwd) *frames_fwd += nf;
    if (frames_bwd) *frames_bwd += nb;
    return EX4DGS_OK;
}
const char* ex4dgs_last_error(void) { return g_err; }
void ex4dgs_set_capacity_hint(int instances) { g_last_R = instances > 0 ? instances : 0; }
unsigned ex4dgs_last_inexact_thresholds(void) { return g_last_inexact; }
void ex4dgs_forward_geometry(int* batch, int* warps)
{
    int b = 0, w = 0;
    render_fwd_geometry(&b, &w);
    if (batch) *batch = b;
    if (warps) *warps = w;
}

size_t ex4dgs_geometry_bytes(int P) { return carve_geometry(nullptr, P, binning_geometry_scratch_bytes(P > 0 ? P : 1)).total; }
// upper bound over all image sizes (32-bit tile keys, four tile passes); the forward asks for what the frame's image needs
size_t ex4dgs_binning_bytes(int R) { return carve_binning(nullptr, R, binning_status_bytes(R, 4), 4).total; }
size_t ex4dgs_image_bytes(int width, int height) { return carve_image(nullptr, width, height).total; }

int ex4dgs_describe_buffers(int P, int R, int width, int height, ex4dgs_array_desc* out, int max)
{
    // the binning buffer is laid out from its capacity: the one of the calling thread's last forward when that covers R
    const int cap = g_last_cap >= R ? g_last_cap : R;
    const GeometryState g = carve_geometry(nullptr, P, binning_geometry_scratch_bytes(P > 0 ? P : 1));
    const int gx = (width + EX_TILE - 1) / EX_TILE, gy = (height + EX_TILE - 1) / EX_TILE;
    const int key_bytes = binning_tile_key_bytes(gx, gy);
    const BinningState b = carve_binning(nullptr, cap, binning_status_bytes(cap, binning_tile_passes(gx, gy)), key_bytes);
    const ImageState im = carve_image(nullptr, width, height);
    const size_t n = (size_t)P, r = (size_t)(R > 0 ? R : 0), px = (size_t)width * height;
    const size_t tiles = (size_t)((width + EX_TILE - 1) / EX_TILE) * ((height + EX_TILE - 1) / EX_TILE);
    const ex4dgs_array_desc all[] = {
        {"depth_key", 0, (size_t)g.key_in, 4, n},
        {"order", 0, (size_t)g.order, 4, n},
        {"tiles_touched", 0, (size_t)g.tiles_touched, 4, n},
        {"rec", 0, (size_t)g.rec, sizeof(SplatRec), n},
        {"clamped", 0, (size_t)g.clamped, 1, n},
        {"gacc", 0, (size_t)g.gacc, sizeof(GradAcc), n},
        {"meta", 0, (size_t)g.meta, 4, 64},
        {"tile_sorted", 1, (size_t)b.tile[0], (size_t)key_bytes, r},
        {"point_list", 1, (size_t)b.val[0], 4, r},
        {"final_T", 2, (size_t)im.final_T, 4, px},
        {"n_contrib", 2, (size_t)im.n_contrib, 4, px},
        {"ranges", 2, (size_t)im.ranges, 8, tiles},
        {"tile_batches", 2, (size_t)im.tile_batches, 4, tiles},
    };
    const int count = (int)(sizeof(all) / sizeof(all[0]));
    for (int i = 0; i < count && i < max; i++) out[i] = all[i];
    return count;
}

// EX4DGS_FLAG_SH_SEGMENTED: `ptr` is a host ex4dgs_sh_segments; returns EX4DGS_OK or the (already recorded) error
static int read_segments(const float* ptr, int P, int M, ShSegments* out, const char* what)
{
    memset(out, 0, sizeof(*out));
    const ex4dgs_sh_segments* h = reinterpret_cast<const ex4dgs_sh_segments*>(ptr);
    if (!h) return fail(EX4DGS_ERR_INVALID, "%s: NULL ex4dgs_sh_segments", what);
    if (M != 16) return fail(EX4DGS_ERR_UNSUPPORTED, "%s: segmented SH needs M = 16 coefficients, got %d", what, M);
    if (h->n_static < 0 || h->n_static > P) return fail(EX4DGS_ERR_INVALID, "%s: n_static=%d outside [0, %d]", what, h->n_static, P);
    if ((h->n_static > 0 && (!h->dc_static || !h->rest_static)) || (h->n_static < P && (!h->dc_dynamic || !h->rest_dynamic)))
        return fail(EX4DGS_ERR_INVALID, "%s: a non-empty SH segment has a NULL pointer", what);
    out->enabled = 1;
    out->n_static = h->n_static;
    out->dc[0] = h->dc_static; out->rest[0] = h->rest_static;
    out->dc[1] = h->dc_dynamic; out->rest[1] = h->rest_dynamic;
    return EX4DGS_OK;
}

static void* align256(void* p) { return (void*)(((uintptr_t)p + 255) & ~(uintptr_t)255); }

int ex4dgs_forward(
    ex4dgs_alloc_fn geometryBuffer, void* geometry_user,
    ex4dgs_alloc_fn binningBuffer, void* binning_user,
    ex4dgs_alloc_fn imageBuffer, void* image_user,
    int P, int D, int M,
    const float* background, int width, int height,
    const float* means3D, const float* dir3D, const float* shs, const float* colors_precomp,
    const float* opacities, const float* scales, float scale_modifier, const float* rotations,
    const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix, const float* cam_pos,
    float tan_fovx, float tan_fovy, float kernel_size, const float* subpixel_offset, int prefiltered,
    float* out_color, float min_depth, float max_depth, float* out_depth, float* out_acc, float* out_flow,
    int* out_idx, int* radii, int debug, unsigned flags, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    g_err[0] = 0;
    Prof prof(false, s);
    if (P < 0 || width <= 0 || height <= 0) return fail(EX4DGS_ERR_INVALID, "bad sizes P=%d W=%d H=%d", P, width, height);
    if (!geometryBuffer || !binningBuffer || !imageBuffer) return fail(EX4DGS_ERR_INVALID, "allocator callbacks are required");
    if (!background || !viewmatrix || !projmatrix || !cam_pos || !subpixel_offset)
        return fail(EX4DGS_ERR_INVALID, "background/viewmatrix/projmatrix/cam_pos/subpixel_offset are required");
    if (!out_color || !out_depth || !out_acc || !out_flow || !out_idx) return fail(EX4DGS_ERR_INVALID, "output pointers are required");
    if (P > 0) {
        if (!means3D || !opacities || !dir3D || !radii) return fail(EX4DGS_ERR_INVALID, "means3D/opacities/dir3D/radii are required");
        if ((shs == nullptr) == (colors_precomp == nullptr))
            return fail(EX4DGS_ERR_INVALID, "Please provide excatly one of either SHs or precomputed colors!");
        if (((scales == nullptr || rotations == nullptr) && cov3D_precomp == nullptr) ||
            ((scales != nullptr || rotations != nullptr) && cov3D_precomp != nullptr))
            return fail(EX4DGS_ERR_INVALID, "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
        if (shs != nullptr && (M <= 0 || M < (D + 1) * (D + 1) || D < 0 || D > 3))
            return fail(EX4DGS_ERR_INVALID, "SH degree %d needs %d coefficients, got M=%d", D, (D + 1) * (D + 1), M);
    }
    const int grid_x = (width + EX_TILE - 1) / EX_TILE, grid_y = (height + EX_TILE - 1) / EX_TILE;
    if ((long long)grid_x * grid_y > 0x3fffffffLL)
        return fail(EX4DGS_ERR_UNSUPPORTED, "more than 2^30-1 tiles (%dx%d) is not supported", grid_x, grid_y);
    const int tile_passes = binning_tile_passes(grid_x, grid_y), key_bytes = binning_tile_key_bytes(grid_x, grid_y);

    // image-sized scratch
    const size_t img_bytes = carve_image(nullptr, width, height).total;
    void* img_base = imageBuffer(image_user, img_bytes);
    if (!img_base) return fail(EX4DGS_ERR_ALLOC, "imageBuffer(%zu) returned NULL", img_bytes);
    const ImageState img = carve_image(align256(img_base), width, height);

    RenderParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.W = width; rp.H = height; rp.grid_x = grid_x;
    rp.subpixel_offset = reinterpret_cast<const float2*>(subpixel_offset);
    rp.min_depth = min_depth; rp.max_depth = max_depth;
    rp.final_T = img.final_T; rp.n_contrib = img.n_contrib; rp.tile_batches = img.tile_batches; rp.ranges = img.ranges;
    rp.out_color = out_color; rp.out_depth = out_depth; rp.out_acc = out_acc; rp.out_flow = out_flow; rp.out_idx = out_idx;

    int R = 0;
    uint32_t flow32 = 0;      // read back with R: does any visible Gaussian carry a non-zero dir3D?
    GeometryState geom;
    memset(&geom, 0, sizeof(geom));
    BinningState bin;
    memset(&bin, 0, sizeof(bin));

    PreprocessParams pp;
    memset(&pp, 0, sizeof(pp));
    // camera constants stay on the device: the kernels stage them through shared memory
    pp.view = viewmatrix; pp.proj = projmatrix; pp.cam = cam_pos;
    rp.bg = background;

    // compositing of the sorted lists in `bin` (everything else of rp is set)
    auto render = [&](bool with_flow) -> int {
        rp.point_list = bin.val[0];
        rp.rec = geom.rec;
        CUtensorMap rec_map;
        const bool have_map = P > 0 && render_fwd_uses_gather() && make_record_tensor_map(&rec_map, geom.rec, P);
        if (P > 0 && render_fwd_uses_gather() && !have_map)
            return fail(EX4DGS_ERR_CUDA, "cuTensorMapEncodeTiled is not available (forward built with TMA gather staging)");
        launch_render_fwd(rp, have_map ? &rec_map : nullptr, grid_x, grid_y, with_flow, s);
        g_launches += 1;
        return EX4DGS_OK;
    };

    if (P <= 0) {
        void* bin_base = binningBuffer(binning_user, carve_binning(nullptr, 0, binning_status_bytes(0, tile_passes), key_bytes).total);
        if (!bin_base) return fail(EX4DGS_ERR_ALLOC, "binningBuffer returned NULL");
        CK(cudaMemsetAsync(img.ranges, 0, sizeof(uint2) * (size_t)grid_x * grid_y, s));
        prof.mark();
        const int rc = render(false);
        if (rc < 0) return rc;
        STAGE(debug, s, "render");
        prof.mark();
        prof.done();
        g_last_cap = 0;
        return 0;
    }

    const size_t scratch = binning_geometry_scratch_bytes(P);
    const size_t geom_bytes = carve_geometry(nullptr, P, scratch).total;
    void* geom_base = geometryBuffer(geometry_user, geom_bytes);
    if (!geom_base) return fail(EX4DGS_ERR_ALLOC, "geometryBuffer(%zu) returned NULL", geom_bytes);
    geom = carve_geometry(align256(geom_base), P, scratch);

    pp.P = P; pp.D = D; pp.M = M;
    pp.means3D = means3D; pp.dir3D = dir3D; pp.scales = scales; pp.rotations = rotations;
    pp.opacities = opacities; pp.shs = shs; pp.cov3D_precomp = cov3D_precomp; pp.colors_precomp = colors_precomp;
    if ((flags & EX4DGS_FLAG_SH_SEGMENTED) && shs != nullptr) {
        const int rc = read_segments(shs, P, M, &pp.seg, "forward");
        if (rc < 0) return rc;
        pp.shs = nullptr;
    }
    pp.scale_modifier = scale_modifier;
    pp.W = width; pp.H = height;
    pp.tan_fovx = tan_fovx; pp.tan_fovy = tan_fovy;
    pp.focal_y = height / (2.0f * tan_fovy);
    pp.focal_x = width / (2.0f * tan_fovx);
    pp.kernel_size = kernel_size;
    pp.min_depth = min_depth; pp.max_depth = max_depth;
    pp.grid_x = grid_x; pp.grid_y = grid_y;
    pp.prefiltered = prefiltered; pp.flags = flags;
    pp.radii = radii; pp.key_in = geom.key_in; pp.tiles_touched = geom.tiles_touched;
    pp.rec = geom.rec; pp.clamped = geom.clamped;
    pp.pad_ptr = reinterpret_cast<const float*>(geom.meta + EX_META_PAD);
    pp.flow_flag = geom.meta + EX_META_FLOW;
    pp.inexact_thr = geom.meta + EX_META_INEXACT;
    // device scalars, digit histograms and look-back words: one clear (the sort scratch directly follows meta)
    CK(cudaMemsetAsync(geom.meta, 0, (size_t)(geom.temp - reinterpret_cast<char*>(geom.meta)) + binning_geometry_zero_bytes(P), s));
    if (flags & EX4DGS_FLAG_TILE_CULL)
        CK(launch_subpixel_absmax(subpixel_offset, (size_t)width * height * 2, geom.meta + EX_META_PAD, s));
    prof.mark();
    launch_preprocess_fwd(pp, s);
    g_launches += 1;
    STAGE(debug, s, "preprocess");
    prof.mark();

    CK(binning_depth_order(geom, P, s));
    g_launches += 6;
    STAGE(debug, s, "depth sort");
    prof.mark();

    const bool no_wait = (flags & EX4DGS_FLAG_NO_HOST_WAIT) != 0;
    if (no_wait && g_last_R <= 0)
        return fail(EX4DGS_ERR_INVALID, "EX4DGS_FLAG_NO_HOST_WAIT needs a capacity hint (ex4dgs_set_capacity_hint, or an earlier forward of this thread)");
    // R is on the device now; the host reads it (rasterizer_impl.cu:299) but waits for it only after everything is queued
    uint32_t* rb = no_wait ? nullptr : g_readback.get();
    if (!rb && !no_wait) return fail(EX4DGS_ERR_ALLOC, "cudaHostAlloc of the read-back words failed");
    cudaEvent_t rb_ev = no_wait ? nullptr : g_readback.event();
    if (!no_wait) {
        CK(cudaMemcpyAsync(rb, geom.meta + EX_META_FLOW, 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        if (rb_ev) CK(cudaEventRecord(rb_ev, s));
    }

    // Nothing below waits for the instance count R before it is launched: the binning buffer is asked for with a
    // capacity guessed from the previous frames of this thread (largest recent R + 25 %), the duplicate kernel drops
    // what does not fit, the tile sort and the ranges kernel read min(R, capacity) from the device.  The host waits
    // for R only after the compositing kernel is queued - the GPU never idles.  If the guess was too small (or there
    // is none: first frame) the instance stages run (once more) into a buffer of exactly R entries.
    bool marked = false;
    int cap = 0;
    if (g_last_R > 0) {
        const long long guess = (long long)g_last_R + g_last_R / 4 + 65536;
        cap = (int)(guess < 0x3fffffffLL ? guess : 0x3fffffffLL);
    }
    for (int attempt = 0; attempt < 2; attempt++) {
        if (attempt == 1) CK(binning_reset_instances(geom, P, s));
        if (cap > 0 || attempt == 1) {
            const size_t status_bytes = binning_status_bytes(cap, tile_passes);
            const size_t bin_bytes = carve_binning(nullptr, cap, status_bytes, key_bytes).total;
            void* bin_base = binningBuffer(binning_user, bin_bytes);
            if (!bin_base) return fail(EX4DGS_ERR_ALLOC, "binningBuffer(%zu) returned NULL", bin_bytes);
            bin = carve_binning(align256(bin_base), cap, status_bytes, key_bytes);
        }
        if (cap > 0) {
            CK(binning_duplicate(geom, bin, radii, P, cap, grid_x, grid_y, flags, s));
            g_launches += 1;
        }
        if (cap > 0 || attempt == 1) {
            CK(binning_sort_ranges(geom, bin, img, P, cap, grid_x, grid_y, flags, s));
            g_launches += cap > 0 ? 4 : 0;
            STAGE(debug, s, "duplicate + tile sort + ranges");
            if (!marked) prof.mark();
            // the flow flag is known on the host only after the wait: the first attempt carries the flow accumulators
            // whenever a previous frame of this thread did
            const int rc = render(no_wait ? true : (attempt == 0 ? g_last_flow : flow32 != 0));
            if (rc < 0) return rc;
            STAGE(debug, s, "render");
            if (!marked) prof.mark();
            marked = true;
        }
        if (attempt == 1) break;
        if (no_wait) {            // nothing is read back: the capacity stands in for R (truncation is flagged on the device)
            R = cap;
            break;
        }
        if (rb_ev) CK(cudaEventSynchronize(rb_ev));
        else CK(cudaStreamSynchronize(s));
        if (rb[8] & 1u) {
            rb[8] = 0;
            return fail(EX4DGS_ERR_CUDA, "binning: a look-back of the previous forward's tile sort did not complete");
        }
        flow32 = rb[EX_META_FLOW - EX_META_FLOW];
        g_last_inexact = rb[EX_META_INEXACT - EX_META_FLOW];
        const uint32_t total = rb[EX_META_TOTAL - EX_META_FLOW];
        if (rb[EX_META_ERROR - EX_META_FLOW] != 0 || total > 0x3fffffffu)
            return fail(EX4DGS_ERR_UNSUPPORTED, "binning: a look-back did not complete or there are more than 2^30-1 (Gaussian, tile) instances (read %u)", total);
        R = (int)total;
        // sizing hint for the next frame: follows R upwards at once, downwards slowly (views alternate in training)
        g_last_R = R > g_last_R ? R : (int)(((long long)g_last_R * 15 + R) / 16);
        const bool flow_ok = (flow32 != 0) == g_last_flow || g_last_flow;     // carrying unused flow accumulators is harmless
        g_last_flow = flow32 != 0;
        if (cap > 0 && R <= cap && flow_ok) break;
        cap = R;
    }
    g_last_cap = cap;
    // look-back failures of the tile passes (queued behind the read-back above) are reported by the next forward
    if (!no_wait) CK(cudaMemcpyAsync(rb + 8, geom.meta + EX_META_ERROR, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    prof.done();
    return R;
}

int ex4dgs_backward(
    int P, int D, int M, int R,
    const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* scales, float scale_modifier, const float* rotations,
    const float* acc_depth, const float* acc, float min_depth, float max_depth,
    const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix, const float* campos,
    float tan_fovx, float tan_fovy, float kernel_size, const float* subpixel_offset, const int* radii,
    void* geom_buffer, void* binning_buffer, void* image_buffer,
    const float* dL_dpix, const float* dL_ddepth, const float* dL_dflow, const float* dL_dacc,
    float* dL_dmean2D, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D,
    float* dL_dsh, float* dL_dscale, float* dL_drot, float* dL_ddir,
    int debug, unsigned flags, void* stream)
{
    (void)flags;
    cudaStream_t s = (cudaStream_t)stream;
    g_err[0] = 0;
    Prof prof(true, s);
    if (P <= 0) return EX4DGS_OK;
    if (!geom_buffer || !binning_buffer || !image_buffer) return fail(EX4DGS_ERR_INVALID, "scratch buffers of the forward are required");
    // dL_ddepth / dL_dflow / dL_dacc may be NULL = "no upstream gradient" (all zero): their terms are skipped
    if (!dL_dpix) return fail(EX4DGS_ERR_INVALID, "the upstream colour gradient is required");
    if (!dL_dmean2D || !dL_dopacity || !dL_dcolor || !dL_dmean3D || !dL_ddir) return fail(EX4DGS_ERR_INVALID, "gradient outputs are required");
    if (shs && !dL_dsh) return fail(EX4DGS_ERR_INVALID, "dL_dsh is required when shs is given");
    if (scales && (!dL_dscale || !dL_drot)) return fail(EX4DGS_ERR_INVALID, "dL_dscale/dL_drot are required when scales are given");

    const int grid_x = (width + EX_TILE - 1) / EX_TILE, grid_y = (height + EX_TILE - 1) / EX_TILE;
    // the CUB temp areas are the last sub-arrays and are not used by the backward
    const GeometryState geom = carve_geometry(align256(geom_buffer), P, 0);
    const BinningState bin = carve_binning(align256(binning_buffer), 0, 0, 2);   // the sorted id list is at offset 0
    const ImageState img = carve_image(align256(image_buffer), width, height);

    PreprocessBwdParams bp;
    memset(&bp, 0, sizeof(bp));
    bp.view = viewmatrix; bp.proj = projmatrix; bp.cam = campos;
    prof.mark();
    CK(cudaMemsetAsync(geom.gacc, 0, sizeof(GradAcc) * (size_t)P, s));

    RenderParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.ranges = img.ranges; rp.point_list = bin.val[0]; rp.rec = geom.rec;
    rp.W = width; rp.H = height; rp.grid_x = grid_x;
    rp.subpixel_offset = reinterpret_cast<const float2*>(subpixel_offset);
    rp.bg = background;
    rp.min_depth = min_depth; rp.max_depth = max_depth;
    rp.final_T = img.final_T; rp.n_contrib = img.n_contrib;
    rp.out_depth = const_cast<float*>(acc_depth); rp.out_acc = const_cast<float*>(acc);
    rp.dL_dpix = dL_dpix; rp.dL_ddepth = dL_ddepth; rp.dL_dflow = dL_dflow; rp.dL_dacc = dL_dacc;
    rp.gacc = geom.gacc;
    if (R > 0) {
        CUtensorMap rec_map;
        if (!make_record_tensor_map(&rec_map, geom.rec, P) && EX_BWD_STAGE_GATHER4)
            return fail(EX4DGS_ERR_CUDA, "cuTensorMapEncodeTiled is not available (the backward stages its records with TMA gathers)");
        launch_render_bwd(rp, &rec_map, grid_x, grid_y, s);
        g_launches += 1;
    }
    STAGE(debug, s, "render backward");
    prof.mark();

    bp.P = P; bp.D = D; bp.M = M;
    bp.means3D = means3D; bp.scales = scales; bp.rotations = rotations; bp.shs = shs;
    if ((flags & EX4DGS_FLAG_SH_SEGMENTED) && shs != nullptr) {
        int rc = read_segments(shs, P, M, &bp.seg, "backward");
        if (rc < 0) return rc;
        rc = read_segments(dL_dsh, P, M, &bp.dseg, "backward (dL_dsh)");
        if (rc < 0) return rc;
        if (bp.dseg.n_static != bp.seg.n_static) return fail(EX4DGS_ERR_INVALID, "backward: dL_dsh segments differ from the inputs'");
        bp.shs = nullptr;
        dL_dsh = nullptr;
    }
    bp.cov3D_precomp = cov3D_precomp; bp.colors_precomp = colors_precomp;
    bp.scale_modifier = scale_modifier;
    bp.tan_fovx = tan_fovx; bp.tan_fovy = tan_fovy;
    bp.focal_y = height / (2.0f * tan_fovy);
    bp.focal_x = width / (2.0f * tan_fovx);
    bp.kernel_size = kernel_size;
    bp.radii = radii; bp.clamped = geom.clamped; bp.gacc = geom.gacc; bp.rec = geom.rec;
    bp.W = (float)width; bp.H = (float)height;
    bp.dL_dmean2D = dL_dmean2D; bp.dL_dopacity = dL_dopacity; bp.dL_dcolor = dL_dcolor; bp.dL_dmean3D = dL_dmean3D;
    bp.dL_dcov3D = dL_dcov3D; bp.dL_dsh = dL_dsh; bp.dL_dscale = dL_dscale; bp.dL_drot = dL_drot; bp.dL_ddir = dL_ddir;
    launch_preprocess_bwd(bp, s);
    g_launches += 1;
    STAGE(debug, s, "preprocess backward");
    prof.mark();
    prof.done();
    return EX4DGS_OK;
}

int ex4dgs_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                        float min_depth, float max_depth, uint8_t* present, void* stream)
{
    g_err[0] = 0;
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !projmatrix || !present)))
        return fail(EX4DGS_ERR_INVALID, "bad arguments");
    launch_mark_visible(P, means3D, viewmatrix, projmatrix, min_depth, max_depth, present, (cudaStream_t)stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(EX4DGS_ERR_CUDA, "mark_visible: %s", cudaGetErrorString(e));
    return EX4DGS_OK;
}

size_t ex4dgs_loss_scratch_bytes(int width, int height)
{
    return (width > 0 && height > 0) ? loss_scratch_bytes(width, height) : 0;
}

int ex4dgs_loss_forward(int width, int height, const float* image, const float* gt_image, float lambda_dssim,
                        char* scratch, float* out_loss3, float* l1_errors, float* ssim_errors, void* stream)
{
    g_err[0] = 0;
    if (width <= 0 || height <= 0 || !image || !gt_image || !scratch || !out_loss3 || !l1_errors || !ssim_errors)
        return fail(EX4DGS_ERR_INVALID, "bad arguments");
    cudaError_t e = launch_loss_forward(width, height, image, gt_image, lambda_dssim, scratch, out_loss3,
                                        l1_errors, ssim_errors, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(EX4DGS_ERR_CUDA, "loss_forward: %s", cudaGetErrorString(e));
    g_launches += 2;
    return EX4DGS_OK;
}

int ex4dgs_loss_backward(int width, int height, const float* image, const float* gt_image, float lambda_dssim,
                         const char* scratch, const float* dL_dloss, float* dL_dimage, void* stream)
{
    g_err[0] = 0;
    if (width <= 0 || height <= 0 || !image || !gt_image || !scratch || !dL_dloss || !dL_dimage)
        return fail(EX4DGS_ERR_INVALID, "bad arguments");
    cudaError_t e = launch_loss_backward(width, height, image, gt_image, lambda_dssim, scratch, dL_dloss,
                                         dL_dimage, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(EX4DGS_ERR_CUDA, "loss_backward: %s", cudaGetErrorString(e));
    g_launches += 1;
    return EX4DGS_OK;
}

// per-tensor scalars in double, expression for expression as torch/optim/radam.py (_multi_tensor_radam)
static void radam_scalars(double lr, double step, double beta1, double beta2, float* S, float* U, int* rectified)
{
    const double rho_inf = 2.0 / (1.0 - beta2) - 1.0;
    const double b2t = pow(beta2, step);
    const double rho_t = rho_inf - 2.0 * step * b2t / (1.0 - b2t);
    const double rect = rho_t > 5.0 ? sqrt((rho_t - 4.0) * (rho_t - 2.0) * rho_inf / ((rho_inf - 4.0) * (rho_inf - 2.0) * rho_t)) : 0.0;
    const double unrectified = rect > 0.0 ? 0.0 : 1.0;
    const double bc1 = 1.0 - pow(beta1, step);
    *U = (float)((lr * unrectified / bc1) * -1.0);
    *S = (float)(sqrt(1.0 - b2t) * (lr * rect / bc1) * -1.0);
    *rectified = rect > 0.0;
}

int ex4dgs_radam_scalars(double lr, long long step, double beta1, double beta2, float* S, float* U, int* rectified)
{
    if (step < 1 || !S || !U || !rectified) return EX4DGS_ERR_INVALID;
    radam_scalars(lr, (double)step, beta1, beta2, S, U, rectified);
    return EX4DGS_OK;
}

int ex4dgs_radam_step_ex(const ex4dgs_radam_tensor* tensors, int n, double beta1, double beta2, double eps,
                         double grad_scale, unsigned check_nan_mask, unsigned sanitize_grad_mask, int* nan_flags,
                         void* stream)
{
    g_err[0] = 0;
    if (n < 0 || n > EX4DGS_RADAM_MAX_TENSORS || (n > 0 && !tensors))
        return fail(EX4DGS_ERR_INVALID, "radam_step: n=%d outside [0, %d]", n, EX4DGS_RADAM_MAX_TENSORS);
    if (!(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0) || !(eps >= 0.0))
        return fail(EX4DGS_ERR_INVALID, "radam_step: bad hyper-parameters beta1=%g beta2=%g eps=%g", beta1, beta2, eps);
    if (check_nan_mask && !nan_flags) return fail(EX4DGS_ERR_INVALID, "radam_step: check_nan_mask needs nan_flags");
    RAdamTensorDesc d[EX_OPT_MAX_TENSORS];
    int m = 0;
    for (int i = 0; i < n; i++) {
        const ex4dgs_radam_tensor& t = tensors[i];
        if (t.numel == 0) continue;
        if (!t.param || !t.grad || !t.exp_avg || !t.exp_avg_sq) return fail(EX4DGS_ERR_INVALID, "radam_step: tensor %d has a NULL pointer", i);
        if (t.step < 1) return fail(EX4DGS_ERR_INVALID, "radam_step: tensor %d step=%lld (must be >= 1)", i, t.step);
        d[m].param = t.param; d[m].grad = t.grad; d[m].exp_avg = t.exp_avg; d[m].exp_avg_sq = t.exp_avg_sq;
        d[m].numel = t.numel;
        radam_scalars(t.lr, (double)t.step, beta1, beta2, &d[m].S, &d[m].U, &d[m].rectified);
        d[m].aligned = 0;
        d[m].index = i;
        d[m].check_nan = (check_nan_mask >> i) & 1u;
        d[m].sanitize_grad = (sanitize_grad_mask >> i) & 1u;
        m++;
    }
    cudaError_t e = launch_radam(d, m, beta1, beta2, eps, grad_scale, nan_flags, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(EX4DGS_ERR_CUDA, "radam_step: %s", cudaGetErrorString(e));
    if (m > 0) g_launches += 1;
    return EX4DGS_OK;
}

int ex4dgs_radam_step(const ex4dgs_radam_tensor* tensors, int n, double beta1, double beta2, double eps,
                      double grad_scale, void* stream)
{
    return ex4dgs_radam_step_ex(tensors, n, beta1, beta2, eps, grad_scale, 0u, 0u, nullptr, stream);
}

int ex4dgs_gather_rows(const ex4dgs_gather_job* jobs, int n, void* stream)
{
    g_err[0] = 0;
    if (n < 0 || n > EX4DGS_GATHER_MAX_JOBS || (n > 0 && !jobs))
        return fail(EX4DGS_ERR_INVALID, "gather_rows: n=%d outside [0, %d]", n, EX4DGS_GATHER_MAX_JOBS);
    GatherJob d[EX_GATHER_MAX_JOBS];
    int m = 0;
    for (int i = 0; i < n; i++) {
        const ex4dgs_gather_job& j = jobs[i];
        if (j.n_out == 0) continue;
        if (j.n_out < 0 || j.n_a < 0 || j.n_a > j.n_out || j.n_out > 0xFFFFFFFFLL)
            return fail(EX4DGS_ERR_INVALID, "gather_rows: job %d has n_a=%lld n_out=%lld", i, j.n_a, j.n_out);
        if (j.row_bytes == 0 || (j.row_bytes & 3) != 0 || j.row_bytes > 0x3FFFFFFCu)
            return fail(EX4DGS_ERR_INVALID, "gather_rows: job %d row_bytes=%zu (must be a positive multiple of 4)", i, j.row_bytes);
        if (!j.dst || (j.n_a > 0 && !j.a)) return fail(EX4DGS_ERR_INVALID, "gather_rows: job %d has a NULL pointer", i);
        if (((uintptr_t)j.a | (uintptr_t)j.b | (uintptr_t)j.dst) & 3) return fail(EX4DGS_ERR_INVALID, "gather_rows: job %d is not 4-byte aligned", i);
        d[m].a = (const uint32_t*)j.a; d[m].b = (const uint32_t*)j.b; d[m].dst = (uint32_t*)j.dst; d[m].index = j.index;
        d[m].words = (unsigned)(j.row_bytes / 4); d[m].n_a = (unsigned)j.n_a; d[m].n_out = (unsigned)j.n_out; d[m].vec4 = 0;
        m++;
    }
    cudaError_t e = launch_gather_rows(d, m, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(EX4DGS_ERR_CUDA, "gather_rows: %s", cudaGetErrorString(e));
    if (m > 0) g_launches += 1;
    return EX4DGS_OK;
}

size_t ex4dgs_l1_scratch_bytes(void) { return l1_scratch_bytes(); }

int ex4dgs_l1_forward(size_t n, const float* a, const float* b, char* scratch, float* out_loss, void* stream)
{
    g_err[0] = 0;
    if (n == 0 || !a || !b || !scratch || !out_loss) return fail(EX4DGS_ERR_INVALID, "l1_forward: bad arguments");
    cudaError_t e = launch_l1_forward(n, a, b, scratch, out_loss, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(EX4DGS_ERR_CUDA, "l1_forward: %s", cudaGetErrorString(e));
    g_launches += 2;
    return EX4DGS_OK;
}

int ex4dgs_l1_backward(size_t n, const float* a, const float* b, const float* dL_dloss, float* dL_da, void* stream)
{
    g_err[0] = 0;
    if (n == 0 || !a || !b || !dL_dloss || !dL_da) return fail(EX4DGS_ERR_INVALID, "l1_backward: bad arguments");
    cudaError_t e = launch_l1_backward(n, a, b, dL_dloss, dL_da, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(EX4DGS_ERR_CUDA, "l1_backward: %s", cudaGetErrorString(e));
    g_launches += 1;
    return EX4DGS_OK;
}

int ex4dgs_iteration_stats(int Ns, int Nd, const int* radii, const float* grad_means2D, const float* grad_error,
                           float timestamp, int densify,
                           const ex4dgs_stats_arrays* stat, const ex4dgs_stats_arrays* dyn, void* stream)
{
    g_err[0] = 0;
    if (Ns < 0 || Nd < 0) return fail(EX4DGS_ERR_INVALID, "iteration_stats: negative counts");
    if (Ns + Nd == 0) return EX4DGS_OK;
    if (!radii || !grad_means2D) return fail(EX4DGS_ERR_INVALID, "iteration_stats: radii and grad_means2D are required");
    IterStatsParams p;
    memset(&p, 0, sizeof(p));
    p.Ns = Ns; p.Nd = Nd; p.radii = radii; p.grad_means2D = grad_means2D; p.grad_error = grad_error;
    p.timestamp = timestamp; p.densify = densify ? 1 : 0;
    const ex4dgs_stats_arrays* src[2] = {stat, dyn};
    StatsArrays* dst[2] = {&p.stat, &p.dyn};
    const int cnt[2] = {Ns, Nd};
    for (int i = 0; i < 2; i++) {
        if (cnt[i] == 0) continue;
        const ex4dgs_stats_arrays* a = src[i];
        if (!a) return fail(EX4DGS_ERR_INVALID, "iteration_stats: %s arrays are required", i ? "dynamic" : "static");
        const bool need_err = grad_error != nullptr;
        if ((need_err && !a->min_radii2D) ||
            (densify && (!a->max_radii2D || !a->xyz_gradient_accum || !a->denom)) ||
            (densify && need_err && (!a->error_accum || !a->error_min || !a->error_min_timestamp || !a->ssim_error_accum || !a->error_denom)))
            return fail(EX4DGS_ERR_INVALID, "iteration_stats: a %s statistics array is NULL", i ? "dynamic" : "static");
        dst[i]->max_radii2D = a->max_radii2D; dst[i]->min_radii2D = a->min_radii2D;
        dst[i]->xyz_gradient_accum = a->xyz_gradient_accum; dst[i]->denom = a->denom;
        dst[i]->error_accum = a->error_accum; dst[i]->error_min = a->error_min;
        dst[i]->error_min_timestamp = a->error_min_timestamp; dst[i]->ssim_error_accum = a->ssim_error_accum;
        dst[i]->error_denom = a->error_denom;
    }
    cudaError_t e = launch_iteration_stats(p, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(EX4DGS_ERR_CUDA, "iteration_stats: %s", cudaGetErrorString(e));
    g_launches += 1;
    return EX4DGS_OK;
}

size_t ex4dgs_regularizer_scratch_bytes(void)
{
    return (size_t)regularizer_blocks() * 2 * sizeof(double) + 256;
}

int ex4dgs_regularizers(int Ns, int Nd, int K, const float* xyz_disp, const float* xyz_motion,
                        float static_reg, float motion_reg, const float* dL_dloss,
                        float* dL_dxyz_disp, int accumulate_disp, float* dL_dxyz_motion, int accumulate_motion,
                        float* out_terms, char* scratch, void* stream)
{
    g_err[0] = 0;
    if (Ns < 0 || Nd < 0 || K < 0) return fail(EX4DGS_ERR_INVALID, "regularizers: negative sizes");
    if (!out_terms || !scratch) return fail(EX4DGS_ERR_INVALID, "regularizers: out_terms and scratch are required");
    RegParams p;
    memset(&p, 0, sizeof(p));
    p.Ns = Ns; p.Nd = Nd; p.K = K;
    p.xyz_disp = xyz_disp; p.xyz_motion = xyz_motion;
    // mean() over an empty tensor is NaN in torch; the reference only evaluates the motion term when
    // _xyz_motion has rows (train.py:159) and always has static Gaussians: empty inputs switch a term off
    p.static_coef = (static_reg != 0.0f && Ns > 0) ? static_reg / (float)Ns : 0.0f;
    p.motion_coef = (motion_reg != 0.0f && Nd > 0 && K > 1) ? motion_reg / ((float)Nd * (float)(K - 1)) : 0.0f;
    if (p.static_coef != 0.0f && !xyz_disp) return fail(EX4DGS_ERR_INVALID, "regularizers: xyz_disp is NULL");
    if (p.motion_coef != 0.0f && !xyz_motion) return fail(EX4DGS_ERR_INVALID, "regularizers: xyz_motion is NULL");
    p.dL_dloss = dL_dloss;
    p.dL_dxyz_disp = dL_dxyz_disp; p.dL_dxyz_motion = dL_dxyz_motion;
    p.accumulate_disp = accumulate_disp ? 1 : 0; p.accumulate_motion = accumulate_motion ? 1 : 0;
    p.part = reinterpret_cast<double*>(align256(scratch));
    cudaError_t e = launch_regularizers(p, out_terms, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(EX4DGS_ERR_CUDA, "regularizers: %s", cudaGetErrorString(e));
    g_launches += 2;
    return EX4DGS_OK;
}

}  // extern "C"
