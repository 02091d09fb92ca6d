// Tile binning for sm_100a: depth ordering of Gaussians, duplication into (tile, id) instances,
// stable tile sort and per-tile ranges.
//
// Behavioural spec: rasterizer_impl.cu:72-140 (duplicateWithKeys, identifyTileRanges) and
// :293-336 (scan, 64-bit key radix sort over bits [0, 32+bit), ranges).  The reference sorts R
// 64-bit keys (tile << 32 | depth bits).  Here the same permutation is produced with two stable
// sorts on much less data (SURVEY.md section 7 "Sort traffic"):
//   1. stable sort of the P Gaussians by their 32 depth bits          -> order by (depth, id)
//   2. emit the (tile, id) instances in that order, stable sort by the `bit`-bit tile id only
//      (uint16 keys) -> (tile, depth, id), which is exactly the order of the reference's stable
//      LSD sort with its emission order (= Gaussian index) as tie-break.
// Both sorts are cub::DeviceRadixSort as the north star prescribes.
#include "common.cuh"
#include <cub/cub.cuh>

namespace {

struct TilesInOrder {
    const uint32_t* tiles_touched;
    __host__ __device__ __forceinline__ uint32_t operator()(const uint32_t& id) const { return tiles_touched[id]; }
};

constexpr int kDupThreads = 256;

// One warp handles 32 consecutive entries of `order`.  The (Gaussian, tile) items of the 32
// rectangles are flattened in emission order (depth-sorted Gaussian, then row-major tile) and dealt
// 32 at a time to the lanes: every store instruction writes 32 consecutive list entries, and one
// huge rectangle no longer serialises its warp (rasterizer_impl.cu:100-111 loops per thread and
// writes 2- and 4-byte items at 32 unrelated addresses per instruction).  With exact-output
// culling the surviving items are compacted with a ballot; their order, and therefore the
// offsets of the scan, are preserved.
__global__ void __launch_bounds__(kDupThreads) duplicate_kernel(int P, const uint32_t* __restrict__ order,
                                                                const uint32_t* __restrict__ key_sorted,
                                                                const uint32_t* __restrict__ offsets,
                                                                const SplatRec* __restrict__ rec, const int* __restrict__ radii,
                                                                int grid_x, int grid_y, unsigned flags, const float* __restrict__ pad_ptr,
                                                                uint16_t* __restrict__ tile_out, uint32_t* __restrict__ val_out, uint32_t cap)
{
    __shared__ CullCtx s_ctx[kDupThreads / 32][32];
    __shared__ int s_prefix[kDupThreads / 32][32];
    const unsigned full = 0xffffffffu;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool cull = (flags & 1u) != 0;
    const float pad = cull ? __ldg(pad_ptr) : 0.f;
    const int jw = (j & ~31);
    if (jw >= P) return;
    // the warp's output window starts at wbase: offsets are an inclusive scan in this order
    const uint32_t wbase = (jw == 0) ? 0u : offsets[jw - 1];
    if (offsets[min(jw + 31, P - 1)] == wbase) return;       // nothing to emit

    CullCtx c;
    c.ok = 0; c.w = 0; c.x0 = c.y0 = 0; c.id = 0;
    int area = 0;
    if (j < P && key_sorted[j] != EX_INVISIBLE_KEY) {
        const uint32_t id = order[j];
        const float4 a = rec[id].a;
        int x0, y0, x1, y1;
        tile_rect(a.x, a.y, radii[id], grid_x, grid_y, x0, y0, x1, y1);
        if (cull) {
            const float4 b = rec[id].b;
            tight_rect(a.x, a.y, b.x, b.y, b.z, a.w, pad, x0, y0, x1, y1);     // the rectangle preprocess counted
            c = cull_prepare(a.x, a.y, b.x, b.y, b.z, a.w, x0, y0, x1 - x0, id, pad);
        } else {
            c.x0 = x0; c.y0 = y0; c.w = x1 - x0; c.id = id;
        }
        area = (x1 - x0) * (y1 - y0);
    }
    int incl = area;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(full, incl, d);
        if (lane >= d) incl += t;
    }
    const int total = __shfl_sync(full, incl, 31);
    s_prefix[warp][lane] = incl - area;
    s_ctx[warp][lane] = c;
    __syncwarp();
    for (int base = 0; base < total; base += 32) {
        const int item = base + lane;
        bool keep = false;
        uint16_t tile = 0;
        uint32_t id = 0;
        if (item < total) {
            const int src = expand_owner(s_prefix[warp], item);
            const int local = item - s_prefix[warp][src];
            const int w = s_ctx[warp][src].w;
            const int ty = s_ctx[warp][src].y0 + local / w, tx = s_ctx[warp][src].x0 + local % w;
            id = s_ctx[warp][src].id;
            tile = (uint16_t)(ty * grid_x + tx);
            keep = !(cull && cull_test(s_ctx[warp][src], tx, ty, pad));
        }
        // culled instances keep their slot (the scan counted the full rectangle) but are keyed to
        // the dump tile 0xFFFF, which the stable tile sort moves behind every real tile
        if (item < total && wbase + item < cap) {      // cap: capacity of a speculatively sized buffer (api.cu)
            tile_out[wbase + item] = keep ? tile : (uint16_t)0xFFFF;
            val_out[wbase + item] = id;
        }
    }
}

// ranges[tile] = [first, last+1) of the tile's entries in the sorted list (rasterizer_impl.cu:118-140);
// eight 16-bit keys per thread from one 128-bit load.
__global__ void __launch_bounds__(256) tile_ranges_kernel(int L, const uint16_t* __restrict__ tiles, uint2* __restrict__ ranges, uint32_t dump)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int base = g * 8;
    if (base >= L) return;
    uint16_t t[8];
    if (base + 8 <= L) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(tiles) + g);
        t[0] = v.x & 0xffff; t[1] = v.x >> 16; t[2] = v.y & 0xffff; t[3] = v.y >> 16;
        t[4] = v.z & 0xffff; t[5] = v.z >> 16; t[6] = v.w & 0xffff; t[7] = v.w >> 16;
    } else {
        for (int k = 0; k < 8; k++) t[k] = (base + k < L) ? tiles[base + k] : 0;
    }
    uint32_t prev = (base == 0) ? 0xffffffffu : (uint32_t)tiles[base - 1];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int idx = base + k;
        if (idx < L) {
            const uint32_t cur = t[k];
            if (idx == 0) {
                if (cur != dump) ranges[cur].x = 0;
            } else if (cur != prev) {
                ranges[prev].y = idx;                 // prev is never the dump tile (it sorts last)
                if (cur != dump) ranges[cur].x = idx;
            }
            if (idx == L - 1 && cur != dump) ranges[cur].y = L;
            prev = cur;
        }
    }
}

// max |subpixel offset| -> *out (float bits, non-negative, so integer max orders correctly).
// Needed by the exact tile culling: pixel centres are integer + offset.
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ v, size_t n, uint32_t* out)
{
    float m = 0.f;
    auto upd = [&](float x) {
        const float a = fabsf(x);
        m = (a > m || a != a) ? a : m;     // NaN propagates
    };
    const size_t n4 = n >> 2;               // subpixel_offset is [H,W,2] floats, 16-byte aligned in practice
    const bool aligned = (reinterpret_cast<uintptr_t>(v) & 15) == 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (aligned) {
        const float4* v4 = reinterpret_cast<const float4*>(v);
        for (size_t i = t0; i < n4; i += stride) {
            const float4 x = __ldg(v4 + i);
            upd(x.x); upd(x.y); upd(x.z); upd(x.w);
        }
        for (size_t i = (n4 << 2) + t0; i < n; i += stride) upd(__ldg(v + i));
    } else {
        for (size_t i = t0; i < n; i += stride) upd(__ldg(v + i));
    }
    uint32_t b = __float_as_uint(m);
    b = __reduce_max_sync(0xffffffffu, b);
    if ((threadIdx.x & 31) == 0 && b) atomicMax(out, b);
}

// smallest b with (n >> b) == 0, computed like rasterizer_impl.cu:35-50
uint32_t higher_msb(uint32_t n)
{
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

}  // namespace

// The CUB size queries touch the driver (device attributes, function attributes): memoised, because
// they sit on the critical path right after the forward's host synchronisation.
size_t binning_stage1_temp_bytes(int P)
{
    static thread_local int last_p = -1, last_dev = -1;
    static thread_local size_t last_bytes = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (P == last_p && dev == last_dev) return last_bytes;
    last_dev = dev;
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, P);
    cub::TransformInputIterator<uint32_t, TilesInOrder, const uint32_t*> it(nullptr, TilesInOrder{nullptr});
    cub::DeviceScan::InclusiveSum(nullptr, b, it, (uint32_t*)nullptr, P);
    last_p = P;
    last_bytes = (a > b ? a : b) + 256;
    return last_bytes;
}

size_t binning_stage2_temp_bytes(int R)
{
    // temp size grows monotonically with the item count: query once per 1M-item bucket
    static thread_local long long last_bucket = -1;
    static thread_local int last_dev = -1;
    static thread_local size_t last_bytes = 0;
    const long long bucket = ((long long)(R > 0 ? R : 1) + 0xFFFFF) >> 20;
    int dev = 0;
    cudaGetDevice(&dev);
    if (bucket == last_bucket && dev == last_dev) return last_bytes;
    last_dev = dev;
    const long long n = bucket << 20;
    size_t a = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint16_t*)nullptr, (uint16_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)(n < 0x7fffffffLL ? n : 0x7fffffffLL));
    last_bucket = bucket;
    last_bytes = a + 256;
    return last_bytes;
}

// `out` must have been zeroed by the caller
cudaError_t launch_subpixel_absmax(const float* subpixel_offset, size_t n, uint32_t* out, cudaStream_t s)
{
    absmax_kernel<<<148 * 8, 256, 0, s>>>(subpixel_offset, n, out);
    return cudaGetLastError();
}

cudaError_t binning_stage1(const GeometryState& g, int P, cudaStream_t s)
{
    size_t tb = g.temp_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(g.temp, tb, g.key_in, g.key_sorted, g.val_in, g.order, P, 0, 32, s);
    if (e != cudaSuccess) return e;
    cub::TransformInputIterator<uint32_t, TilesInOrder, const uint32_t*> it(g.order, TilesInOrder{g.tiles_touched});
    tb = g.temp_bytes;
    return cub::DeviceScan::InclusiveSum(g.temp, tb, it, g.offsets, P, s);
}

cudaError_t binning_duplicate(const GeometryState& g, const BinningState& b, const int* radii, int P, int cap,
                              int grid_x, int grid_y, unsigned flags, cudaStream_t s)
{
    if (P <= 0 || cap <= 0) return cudaSuccess;
    duplicate_kernel<<<(P + kDupThreads - 1) / kDupThreads, kDupThreads, 0, s>>>(
        P, g.order, g.key_sorted, g.offsets, g.rec, radii, grid_x, grid_y, flags,
        reinterpret_cast<const float*>(g.meta), b.tile_unsorted, b.val_unsorted, (uint32_t)cap);
    return cudaGetLastError();
}

cudaError_t binning_sort_ranges(const BinningState& b, const ImageState& img, int R, int grid_x, int grid_y,
                                unsigned flags, cudaStream_t s)
{
    const int tiles = grid_x * grid_y;
    cudaError_t e = cudaMemsetAsync(img.ranges, 0, sizeof(uint2) * (size_t)tiles, s);
    if (e != cudaSuccess || R <= 0) return e;
    size_t tb = b.temp_bytes;
    const int bit = (int)higher_msb((uint32_t)tiles);
    const bool cull = (flags & 1u) != 0;
    e = cub::DeviceRadixSort::SortPairs(b.temp, tb, b.tile_unsorted, b.tile_sorted, b.val_unsorted, b.point_list,
                                        R, 0, (cull || bit > 16) ? 16 : bit, s);
    if (e != cudaSuccess) return e;
    const int groups = (R + 7) / 8;
    tile_ranges_kernel<<<(groups + 255) / 256, 256, 0, s>>>(R, b.tile_sorted, img.ranges, cull ? 0xFFFFu : 0xFFFFFFFFu);
    return cudaGetLastError();
}
