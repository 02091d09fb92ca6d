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

// One warp handles 32 consecutive entries of `order`; Gaussians with few tiles are written by
// their own lane, large rectangles are spread over the whole warp (the per-thread serial loop of
// rasterizer_impl.cu:100-111 is badly imbalanced for big splats).
__global__ void __launch_bounds__(256) duplicate_kernel(int P, const uint32_t* __restrict__ order,
                                                        const uint32_t* __restrict__ key_sorted,
                                                        const uint32_t* __restrict__ offsets,
                                                        const uint32_t* __restrict__ tiles_touched,
                                                        const SplatRec* __restrict__ rec, const int* __restrict__ radii,
                                                        int grid_x, int grid_y, unsigned flags,
                                                        uint16_t* __restrict__ tile_out, uint32_t* __restrict__ val_out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    uint32_t id = 0, cnt = 0, off = 0;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    float px = 0, py = 0, A = 0, B = 0, C = 0, thr = 0;
    if (j < P && key_sorted[j] != EX_INVISIBLE_KEY) {
        id = order[j];
        cnt = tiles_touched[id];
        off = (j == 0) ? 0u : offsets[j - 1];
        const float4 a = rec[id].a;
        px = a.x; py = a.y; thr = a.w;
        tile_rect(px, py, radii[id], grid_x, grid_y, x0, y0, x1, y1);
        if (flags & 1u) {
            const float4 b = rec[id].b;
            A = b.x; B = b.y; C = b.z;
        }
    }
    const bool cull = (flags & 1u) != 0;
    const bool big = cnt > 32;
    if (cnt && !big) {
        for (int ty = y0; ty < y1; ty++)
            for (int tx = x0; tx < x1; tx++) {
                if (cull && tile_cannot_contribute(px, py, A, B, C, thr, tx, ty, 0.5f)) continue;
                tile_out[off] = (uint16_t)(ty * grid_x + tx);
                val_out[off] = id;
                off++;
            }
    }
    // cooperative path for large rectangles
    unsigned todo = __ballot_sync(0xffffffffu, big);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t bid = __shfl_sync(0xffffffffu, id, src);
        const uint32_t boff = __shfl_sync(0xffffffffu, off, src);
        const int bx0 = __shfl_sync(0xffffffffu, x0, src), by0 = __shfl_sync(0xffffffffu, y0, src);
        const int bx1 = __shfl_sync(0xffffffffu, x1, src), by1 = __shfl_sync(0xffffffffu, y1, src);
        const int w = bx1 - bx0, n = w * (by1 - by0);
        if (!cull) {
            for (int t = lane; t < n; t += 32) {
                const int ty = by0 + t / w, tx = bx0 + t % w;
                tile_out[boff + t] = (uint16_t)(ty * grid_x + tx);
                val_out[boff + t] = bid;
            }
        } else {
            const float bpx = __shfl_sync(0xffffffffu, px, src), bpy = __shfl_sync(0xffffffffu, py, src);
            const float bA = __shfl_sync(0xffffffffu, A, src), bB = __shfl_sync(0xffffffffu, B, src);
            const float bC = __shfl_sync(0xffffffffu, C, src), bthr = __shfl_sync(0xffffffffu, thr, src);
            uint32_t base = boff;
            for (int t0 = 0; t0 < n; t0 += 32) {
                const int t = t0 + lane;
                bool keep = false;
                int ty = 0, tx = 0;
                if (t < n) {
                    ty = by0 + t / w; tx = bx0 + t % w;
                    keep = !tile_cannot_contribute(bpx, bpy, bA, bB, bC, bthr, tx, ty, 0.5f);
                }
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const uint32_t o = base + __popc(m & ((1u << lane) - 1u));
                    tile_out[o] = (uint16_t)(ty * grid_x + tx);
                    val_out[o] = bid;
                }
                base += __popc(m);
            }
        }
    }
}

__global__ void __launch_bounds__(256) tile_ranges_kernel(int L, const uint16_t* __restrict__ tiles, uint2* __restrict__ ranges)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L) return;
    const uint32_t cur = tiles[idx];
    if (idx == 0)
        ranges[cur].x = 0;
    else {
        const uint32_t prev = tiles[idx - 1];
        if (cur != prev) {
            ranges[prev].y = idx;
            ranges[cur].x = idx;
        }
    }
    if (idx == L - 1) ranges[cur].y = L;
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

size_t binning_stage1_temp_bytes(int P)
{
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, P);
    cub::TransformInputIterator<uint32_t, TilesInOrder, const uint32_t*> it(nullptr, TilesInOrder{nullptr});
    cub::DeviceScan::InclusiveSum(nullptr, b, it, (uint32_t*)nullptr, P);
    return (a > b ? a : b) + 256;
}

size_t binning_stage2_temp_bytes(int R)
{
    size_t a = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint16_t*)nullptr, (uint16_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, R > 0 ? R : 1);
    return a + 256;
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

cudaError_t binning_stage2(const GeometryState& g, const BinningState& b, const ImageState& img,
                           const int* radii, int P, int R, int grid_x, int grid_y, unsigned flags,
                           cudaStream_t s)
{
    const int tiles = grid_x * grid_y;
    cudaError_t e = cudaMemsetAsync(img.ranges, 0, sizeof(uint2) * (size_t)tiles, s);
    if (e != cudaSuccess || R <= 0) return e;
    duplicate_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, g.order, g.key_sorted, g.offsets, g.tiles_touched, g.rec,
                                                     radii, grid_x, grid_y, flags, b.tile_unsorted, b.val_unsorted);
    size_t tb = b.temp_bytes;
    const int bit = (int)higher_msb((uint32_t)tiles);
    e = cub::DeviceRadixSort::SortPairs(b.temp, tb, b.tile_unsorted, b.tile_sorted, b.val_unsorted, b.point_list,
                                        R, 0, bit < 16 ? bit : 16, s);
    if (e != cudaSuccess) return e;
    tile_ranges_kernel<<<(R + 255) / 256, 256, 0, s>>>(R, b.tile_sorted, img.ranges);
    return cudaGetLastError();
}
