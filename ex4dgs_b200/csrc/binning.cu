// Tile binning for sm_100a: depth ordering of the visible Gaussians, duplication into (tile, id)
// instances, stable tile sort and per-tile ranges - every item count stays on the device.
//
// Behavioural spec: rasterizer_impl.cu:72-140 (duplicateWithKeys, identifyTileRanges) and
// :293-336 (scan, 64-bit key radix sort over bits [0, 32+bit), ranges).  The reference sorts R
// 64-bit keys (tile << 32 | depth bits).  Here the same permutation is produced with two stable
// sorts on much less data (SURVEY.md section 7 "Sort traffic"):
//   1. stable sort of the VISIBLE Gaussians by their 32 depth bits      -> order by (depth, id)
//   2. emit the (tile, id) instances in that order, stable sort by the `bit`-bit tile id only
//      (uint16 keys; uint32 keys and three or four passes for images of more than 65535 tiles)
//      -> (tile, depth, id), which is exactly the order of the reference's stable LSD sort with its
//      emission order (= Gaussian index) as tie-break.
//
// Both sorts are the one-sweep LSD radix sort below (8-bit digits, one kernel per digit: rank the
// tile's keys per warp through shared-memory lane masks, publish the tile's digit counts, chained look-back over the previous
// tiles, scatter through shared memory).  It does what cub::DeviceRadixSort does, with three
// differences that the pipeline needs: the item count is read from DEVICE memory (the number of
// visible Gaussians / of instances is produced by the kernel in front, the host never waits for it
// before launching), the first pass of a sort can drop the all-ones key on the fly (depth sort: the
// culled Gaussians - P keys in, P_vis pairs out, the later passes and the duplicate kernel only see
// visible ones; tile sort: the instances the exact tile test rejected), and the values of the first
// depth pass are the indices themselves (no iota array).  The prefix sum of `tiles_touched` in depth
// order (rasterizer_impl.cu:295 InclusiveSum) is replaced by three levels of partial sums
// (touched_sums_kernel, which also leaves the instance count R on the device) from which every warp
// of the duplicate kernel adds up its own first output position.
#include "common.cuh"
#include <type_traits>

namespace {

constexpr int kBins = 256;                       // 8-bit digits
// shape of a pass: CTAs of THREADS threads take tiles of THREADS * ITEMS pairs; MINBLOCKS resident CTAs per SM
#ifndef EX_SORT_THREADS_DEPTH
#define EX_SORT_THREADS_DEPTH 512
#endif
#ifndef EX_SORT_ITEMS_DEPTH
#define EX_SORT_ITEMS_DEPTH 8
#endif
#ifndef EX_SORT_MINBLOCKS_DEPTH
#define EX_SORT_MINBLOCKS_DEPTH 2
#endif
#ifndef EX_SORT_THREADS_TILE
#define EX_SORT_THREADS_TILE 256
#endif
#ifndef EX_SORT_ITEMS_TILE
#define EX_SORT_ITEMS_TILE 16
#endif
#ifndef EX_SORT_MINBLOCKS_TILE
#define EX_SORT_MINBLOCKS_TILE 3
#endif
#ifndef EX_SORT_WINDOW
#define EX_SORT_WINDOW 8                         // status words a look-back step loads at once
#endif
constexpr uint32_t kFlagAgg = 1u << 30;          // status word = 2 flag bits | 30-bit count
constexpr uint32_t kFlagPrefix = 2u << 30;
constexpr uint32_t kFlagMask = 3u << 30;
constexpr uint32_t kValMask = ~kFlagMask;
constexpr int kSpinLimit = 1 << 22;              // a look-back that never completes raises meta[kMetaError] instead of hanging the GPU

__device__ __forceinline__ uint32_t ld_status(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(uint32_t* p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- digit histograms of all passes in one read of the keys -------------------------------------
// DROP: keys equal to EX_INVISIBLE_KEY are not counted; *n_valid receives the number of counted keys.
template <typename KeyT, int NPASS, bool DROP>
__global__ void __launch_bounds__(1024) radix_hist_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ count_ptr,
                                                         uint32_t cap, uint32_t* __restrict__ hist, uint32_t* __restrict__ n_valid)
{
    constexpr int KPV = 16 / (int)sizeof(KeyT);          // keys per 128-bit load
    __shared__ uint32_t sh[NPASS][kBins];
    for (int i = threadIdx.x; i < NPASS * kBins; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    const uint32_t n = count_ptr ? min(__ldg(count_ptr), cap) : cap;
    const uint32_t nvec = n / KPV;                       // full vectors; the tail is counted by one thread per key
    uint32_t mine = 0;
    auto count = [&](uint32_t k) {
        if (DROP && k == (uint32_t)(KeyT)EX_INVISIBLE_KEY) return;
        mine++;
#pragma unroll
        for (int p = 0; p < NPASS; p++) atomicAdd(&sh[p][(k >> (8 * p)) & 255u], 1u);
    };
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += gridDim.x * blockDim.x) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(keys) + v);
        if (sizeof(KeyT) == 4) {
            count(q.x); count(q.y); count(q.z); count(q.w);
        } else {
            count(q.x & 0xffff); count(q.x >> 16); count(q.y & 0xffff); count(q.y >> 16);
            count(q.z & 0xffff); count(q.z >> 16); count(q.w & 0xffff); count(q.w >> 16);
        }
    }
    if (blockIdx.x == 0 && nvec * KPV + threadIdx.x < n) count((uint32_t)keys[nvec * KPV + threadIdx.x]);
    __syncthreads();
    for (int i = threadIdx.x; i < NPASS * kBins; i += blockDim.x) {
        const uint32_t c = (&sh[0][0])[i];
        if (c) atomicAdd(hist + i, c);
    }
    if (DROP) {
        mine = __reduce_add_sync(0xffffffffu, mine);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_valid, mine);
    }
}

// ---- one digit pass ---------------------------------------------------------------------------------
// Stable: a tile is taken in ticket order (the ticket is the tile index, so every earlier tile is already
// running when a tile starts waiting for it), items of a tile are ranked in (warp, round, lane) order, which
// is their index order (index = tile base + warp * 32 * ITEMS + round * 32 + lane).
// Ranking inside a warp: the lanes OR their lane bit into a per-(warp, digit) mask in shared memory - the result does
// not depend on the order in which the hardware serialises conflicting lanes - and read the mask of their digit back:
// lanes with the same digit, 12 instructions per key.  (MATCH.ANY resolves one group of equal values per iteration:
// ~200 cycles on 32 distinct digits, 30 us of an 86 us tile pass; nine ballots + bit logic cost 45 instructions.)
// DROP: keys equal to the all-ones key (EX_INVISIBLE_KEY / the dump tile 0xFFFF) are dropped (`hist` does not count them:
// `count_ptr` items in, fewer out).  IOTA: the values are the indices themselves.
// !WRITE_KEYS (last pass of a sort whose keys are not needed afterwards): side_out[g] = side_src[value] rides along.
template <typename KeyT, int THREADS, int ITEMS>
struct PassSmem {
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * ITEMS;
    static constexpr int kMaskWords = 2 * WARPS * kBins;
    static constexpr int kRawWords = TILE > kMaskWords ? TILE : kMaskWords;
    // cnt[WARPS][256] | raw[kRawWords] (lane masks while ranking, then the values in sorted order) | delta[256] | part[2][8] | ticket | key[TILE]
    static constexpr size_t kBytes = sizeof(uint32_t) * (WARPS * kBins + kRawWords + kBins + 16 + 4) + sizeof(KeyT) * TILE;
};

template <typename KeyT, int THREADS, int ITEMS, int MINBLOCKS, bool DROP, bool IOTA, bool WRITE_KEYS>
__global__ void __launch_bounds__(THREADS, MINBLOCKS) radix_pass_kernel(
    const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, KeyT* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, const uint32_t* __restrict__ count_ptr, uint32_t cap, int shift,
    const uint32_t* __restrict__ hist, uint32_t* __restrict__ status, uint32_t* __restrict__ ticket_ctr,
    uint32_t* __restrict__ err, const uint32_t* __restrict__ side_src, uint32_t* __restrict__ side_out)
{
    using L = PassSmem<KeyT, THREADS, ITEMS>;
    constexpr int WARPS = L::WARPS, TILE = L::TILE;
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* const s_cnt = smem;                               // [WARPS][256] per-warp digit counts, then the warp's first local position per digit
    uint32_t* const s_raw = s_cnt + WARPS * kBins;
    uint32_t* const s_delta = s_raw + L::kRawWords;             // [256] (global position) - (position in the sorted tile) of a digit's items
    uint32_t* const s_part = s_delta + kBins;                   // [2][8]
    uint32_t* const s_ticket = s_part + 16;
    KeyT* const s_key = reinterpret_cast<KeyT*>(s_ticket + 4);  // [TILE] the tile in sorted order
    uint32_t* const s_val = s_raw;
    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool owner = tid < kBins;                             // thread `tid` owns digit `tid`
    if (tid == 0) *s_ticket = atomicAdd(ticket_ctr, 1u);
    for (int i = tid; i < WARPS * kBins; i += THREADS) s_cnt[i] = 0;
    for (int i = tid; i < L::kMaskWords; i += THREADS) s_raw[i] = 0;
    const uint32_t ghist = owner ? __ldg(hist + tid) : 0u;      // items of digit `tid` in the whole array
    const uint32_t n = count_ptr ? min(__ldg(count_ptr), cap) : cap;
    __syncthreads();
    const uint32_t ticket = *s_ticket;
    if ((uint64_t)ticket * TILE >= n) return;
    const uint32_t tile_base = ticket * TILE;
    const uint32_t warp_base = tile_base + warp * (32 * ITEMS) + lane;
    const bool all_valid = !DROP && (uint64_t)tile_base + TILE <= n;

    uint32_t key[ITEMS], pos[ITEMS];
    uint32_t valid = 0;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const uint32_t idx = warp_base + i * 32;
        bool ok = all_valid || idx < n;
        key[i] = ok ? (uint32_t)__ldg(keys_in + idx) : 0u;
        if (DROP) ok = ok && key[i] != (uint32_t)(KeyT)EX_INVISIBLE_KEY;
        valid |= (ok ? 1u : 0u) << i;
    }
    // rank inside the warp (two mask arrays, alternating: the clear of round i cannot meet the ORs of round i + 1)
    const uint32_t lane_bit = 1u << lane, lt = lane_bit - 1u;
    uint32_t* const my_cnt = s_cnt + warp * kBins;
    auto rank_rounds = [&](auto all_c) {
        constexpr bool ALL = decltype(all_c)::value;
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const bool ok = ALL || ((valid >> i) & 1u);
            const uint32_t d = (key[i] >> shift) & 255u;
            uint32_t* const mask = s_raw + ((i & 1) * WARPS + warp) * kBins + d;
            if (ok) atomicOr(mask, lane_bit);
            __syncwarp();
            const uint32_t peers = ok ? *mask : 0u;
            const uint32_t before = my_cnt[d];
            __syncwarp();
            if (ok && (peers & lt) == 0u) {                 // the lowest lane of the group
                my_cnt[d] = before + __popc(peers);
                *mask = 0u;
            }
            pos[i] = before + __popc(peers & lt);
        }
    };
    if (all_valid) rank_rounds(std::true_type{}); else rank_rounds(std::false_type{});
    __syncthreads();
    // exclusive offsets of the warps, the tile's count of the digit
    uint32_t cnt = 0;
    uint32_t* const my_status = status + (size_t)ticket * kBins + tid;
    if (owner) {
#pragma unroll
        for (int w = 0; w < WARPS; w++) {
            const uint32_t c = s_cnt[w * kBins + tid];
            s_cnt[w * kBins + tid] = cnt;
            cnt += c;
        }
        // publish as early as possible: the tiles behind wait for this
        st_status(my_status, (ticket == 0 ? kFlagPrefix : kFlagAgg) | cnt);
    }
    // exclusive scans over the digits: position of the digit inside the sorted tile / inside the whole output
    uint32_t ia = cnt, ib = ghist;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t ta = __shfl_up_sync(full, ia, d), tb = __shfl_up_sync(full, ib, d);
        if (lane >= d) { ia += ta; ib += tb; }
    }
    if (owner && lane == 31) { s_part[warp] = ia; s_part[8 + warp] = ib; }
    __syncthreads();
    uint32_t wa = 0, wb = 0, n_tile = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const uint32_t a = s_part[w], b = s_part[8 + w];
        if (w < warp) { wa += a; wb += b; }
        n_tile += a;
    }
    if (owner) {
        const uint32_t tile_start = wa + ia - cnt;      // first position of digit `tid` in the sorted tile
        const uint32_t glob_start = wb + ib - ghist;    // first position of digit `tid` in the output
        // chained look-back: items of digit `tid` in all earlier tiles (kWindow independent loads in flight per step)
        uint32_t prefix = 0;
        if (ticket != 0) {
            constexpr int kWindow = EX_SORT_WINDOW;
            int t = (int)ticket - 1, spins = 0;
            bool done = false;
            while (!done) {
                uint32_t v[kWindow];
#pragma unroll
                for (int k = 0; k < kWindow; k++) v[k] = t - k >= 0 ? ld_status(status + (size_t)(t - k) * kBins + tid) : kFlagPrefix;
#pragma unroll
                for (int k = 0; k < kWindow; k++) {
                    if (!done) {
                        const uint32_t f = v[k] & kFlagMask;
                        if (f == 0) {                    // not published yet: read again from here
                            if (++spins > kSpinLimit) { atomicOr(err, 1u); done = true; }
                            break;
                        }
                        prefix += v[k] & kValMask;
                        --t;
                        if (f == kFlagPrefix) done = true;
                    }
                }
            }
            st_status(my_status, kFlagPrefix | ((prefix + cnt) & kValMask));
        }
        s_delta[tid] = glob_start + prefix - tile_start;
#pragma unroll
        for (int w = 0; w < WARPS; w++) s_cnt[w * kBins + tid] += tile_start;
    }
    __syncthreads();
    // scatter into the sorted tile (the values are fetched only now: fewer live registers while ranking) ...
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        if ((valid >> i) & 1u) {
            const uint32_t idx = warp_base + i * 32;
            const uint32_t d = (key[i] >> shift) & 255u;
            const uint32_t lp = my_cnt[d] + pos[i];
            s_key[lp] = (KeyT)key[i];
            s_val[lp] = IOTA ? idx : __ldg(vals_in + idx);
        }
    }
    __syncthreads();
    // ... and from there to the output: consecutive threads write consecutive addresses within a digit
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const uint32_t p = i * THREADS + tid;
        if (p < n_tile) {
            const KeyT k = s_key[p];
            const uint32_t g = s_delta[((uint32_t)k >> shift) & 255u] + p;
            if (WRITE_KEYS) keys_out[g] = k;
            const uint32_t v = s_val[p];
            vals_out[g] = v;
            if (!WRITE_KEYS && side_src) side_out[g] = __ldg(side_src + v);     // last depth pass: tiles_touched in depth order
        }
    }
}

template <typename KeyT, int THREADS, int ITEMS, int MINBLOCKS, bool DROP, bool IOTA, bool WRITE_KEYS>
cudaError_t launch_pass(int grid, cudaStream_t s, const KeyT* keys_in, const uint32_t* vals_in, KeyT* keys_out, uint32_t* vals_out,
                        const uint32_t* count_ptr, uint32_t cap, int shift, const uint32_t* hist, uint32_t* status, uint32_t* ticket,
                        uint32_t* err, const uint32_t* side_src = nullptr, uint32_t* side_out = nullptr)
{
    constexpr size_t bytes = PassSmem<KeyT, THREADS, ITEMS>::kBytes;
    auto kernel = radix_pass_kernel<KeyT, THREADS, ITEMS, MINBLOCKS, DROP, IOTA, WRITE_KEYS>;
    if (bytes > 48 * 1024) {
        // once per device: a new device (or thread) simply sets the attribute again
        static thread_local int configured_dev = -1;
        int dev = 0;
        cudaGetDevice(&dev);
        if (configured_dev != dev) {
            const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            if (e != cudaSuccess) return e;
            configured_dev = dev;
        }
    }
    kernel<<<grid, THREADS, bytes, s>>>(keys_in, vals_in, keys_out, vals_out, count_ptr, cap, shift, hist, status, ticket, err, side_src, side_out);
    return cudaSuccess;
}

// Sums of tiles_touched (in depth order) over 32 / 256 / 16384 consecutive Gaussians and their total R: with them a
// warp of the duplicate kernel finds its first output position from < 100 words (rasterizer_impl.cu:295 runs an
// inclusive scan over all P Gaussians).
__global__ void __launch_bounds__(256) touched_sums_kernel(const uint32_t* __restrict__ n_vis_ptr, const uint32_t* __restrict__ touched_in_order,
                                                           uint32_t* __restrict__ warp_sum, uint32_t* __restrict__ block_sum,
                                                           uint32_t* __restrict__ super_sum, uint32_t* __restrict__ total)
{
    __shared__ uint32_t s_w[8][8];          // [256-block of this CTA][warp]
    const uint32_t n_vis = __ldg(n_vis_ptr);
    const uint32_t base = blockIdx.x * 2048u;
    if (base >= n_vis) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t j = base + i * 256u + threadIdx.x;
        v[i] = j < n_vis ? touched_in_order[j] : 0u;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t j = base + i * 256u + threadIdx.x;
        const uint32_t w = __reduce_add_sync(0xffffffffu, v[i]);
        if (lane == 0) {
            if ((j & ~31u) < n_vis) warp_sum[j >> 5] = w;
            s_w[i][warp] = w;
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t b = 0;
        if (lane < 8) {
#pragma unroll
            for (int k = 0; k < 8; k++) b += s_w[lane][k];
            if (base + lane * 256u < n_vis) block_sum[blockIdx.x * 8 + lane] = b;
        }
        b = __reduce_add_sync(0xffffffffu, b);
        if (lane == 0) {
            atomicAdd(super_sum + (blockIdx.x >> 3), b);
            atomicAdd(total, b);
        }
    }
}

constexpr int kDupThreads = 256;

// One warp handles 32 consecutive entries of `order`.  The (Gaussian, tile) items of the 32
// rectangles are flattened in emission order (depth-sorted Gaussian, then row-major tile) and dealt
// 32 at a time to the lanes: every store instruction writes 32 consecutive list entries, and one
// huge rectangle no longer serialises its warp (rasterizer_impl.cu:100-111 loops per thread and
// writes 2- and 4-byte items at 32 unrelated addresses per instruction).  With exact-output
// culling the surviving items are compacted with a ballot; their order, and therefore the
// offsets of the scan, are preserved.
// The warp's first output position is the number of items of all earlier Gaussians, summed from the three levels of
// touched_sums_kernel; warps are independent (no barrier, no chain).  The rectangle is recomputed from the stored
// record with the same inline functions on the same floats as in the preprocess kernel (its area IS tiles_touched).
template <typename KeyT>
__global__ void __launch_bounds__(kDupThreads) duplicate_kernel(const uint32_t* __restrict__ n_vis_ptr, const uint32_t* __restrict__ order,
                                                                const uint32_t* __restrict__ touched_in_order,
                                                                const uint32_t* __restrict__ warp_sum, const uint32_t* __restrict__ block_sum,
                                                                const uint32_t* __restrict__ super_sum,
                                                                const SplatRec* __restrict__ rec, const int* __restrict__ radii,
                                                                int grid_x, int grid_y, unsigned flags, const float* __restrict__ pad_ptr,
                                                                KeyT* __restrict__ tile_out, uint32_t* __restrict__ val_out, uint32_t cap)
{
    __shared__ CullCtx s_ctx[kDupThreads / 32][32];
    __shared__ int s_prefix[kDupThreads / 32][32];
    const unsigned full = 0xffffffffu;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_vis = __ldg(n_vis_ptr);
    if ((j & ~31u) >= n_vis) return;
    const bool cull = (flags & 1u) != 0;
    const float pad = cull ? __ldg(pad_ptr) : 0.f;
    const bool have = j < n_vis;

    const int area = have ? (int)touched_in_order[j] : 0;
    const uint32_t id = have ? order[j] : 0u;
    // first output position of the warp
    const uint32_t wg = j >> 5, b = wg >> 3, sb = b >> 6;
    uint32_t part = 0;
    for (uint32_t k = lane; k < sb; k += 32) part += super_sum[k];
    for (uint32_t k = sb * 64 + lane; k < b; k += 32) part += block_sum[k];
    if ((uint32_t)lane < (wg & 7u)) part += warp_sum[b * 8 + lane];
    const uint32_t wbase = __reduce_add_sync(full, part);

    int incl = area;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(full, incl, d);
        if (lane >= d) incl += t;
    }
    const int total = __shfl_sync(full, incl, 31);
    if (total == 0) return;

    CullCtx c;
    c.ok = 0; c.w = 0; c.x0 = c.y0 = 0; c.id = 0;
    if (have) {
        const float4 a = rec[id].a;
        int x0, y0, x1, y1;
        tile_rect(a.x, a.y, radii[id], grid_x, grid_y, x0, y0, x1, y1);
        if (cull) {
            const float4 bb = rec[id].b;
            tight_rect(a.x, a.y, bb.x, bb.y, bb.z, a.w, pad, x0, y0, x1, y1);     // the rectangle preprocess counted
            c = cull_prepare(a.x, a.y, bb.x, bb.y, bb.z, a.w, x0, y0, x1 - x0, id, pad);
        } else {
            c.x0 = x0; c.y0 = y0; c.w = x1 - x0; c.id = id;
        }
    }
    s_prefix[warp][lane] = incl - area;
    s_ctx[warp][lane] = c;
    __syncwarp();
    for (int base = 0; base < total; base += 32) {
        const int item = base + lane;
        bool keep = false;
        KeyT tile = 0;
        uint32_t idv = 0;
        if (item < total) {
            const int src = expand_owner(s_prefix[warp], item);
            const int local = item - s_prefix[warp][src];
            const int w = s_ctx[warp][src].w;
            const int ty = s_ctx[warp][src].y0 + local / w, tx = s_ctx[warp][src].x0 + local % w;
            idv = s_ctx[warp][src].id;
            tile = (KeyT)(ty * grid_x + tx);
            keep = !(cull && cull_test(s_ctx[warp][src], tx, ty, pad));
        }
        // culled instances keep their slot (the count is the full rectangle) but are keyed to
        // the dump tile (all ones: 0xFFFF / 0xFFFFFFFF), which the first pass of the tile sort drops
        if (item < total && wbase + item < cap) {      // cap: capacity of the buffer, sized before R is known (api.cu)
            tile_out[wbase + item] = keep ? tile : (KeyT)EX_INVISIBLE_KEY;
            val_out[wbase + item] = idv;
        }
    }
}

// ranges[tile] = [first, last+1) of the tile's entries in the sorted list (rasterizer_impl.cu:118-140);
// eight 16-bit (four 32-bit) keys per thread from one 128-bit load.  `listed` = entries of the sorted list (the instances
// the exact tile test kept, at most the buffer's capacity), `total` = R.
template <typename KeyT>
__global__ void __launch_bounds__(256) tile_ranges_kernel(const uint32_t* __restrict__ listed_ptr, const uint32_t* __restrict__ total_ptr, uint32_t cap,
                                                          const KeyT* __restrict__ tiles, uint2* __restrict__ ranges, uint32_t* __restrict__ err)
{
    constexpr int KPV = 16 / (int)sizeof(KeyT);
    const int L = (int)min(__ldg(listed_ptr), cap);
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g == 0 && __ldg(total_ptr) > cap) atomicOr(err, 2u);      // the lists are truncated (EX4DGS_FLAG_NO_HOST_WAIT: nobody else notices)
    const int base = g * KPV;
    if (base >= L) return;
    uint32_t t[KPV];
    if (base + KPV <= L) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(tiles) + g);
        if (sizeof(KeyT) == 2) {
            t[0] = v.x & 0xffff; t[1] = v.x >> 16; t[2 % KPV] = v.y & 0xffff; t[3 % KPV] = v.y >> 16;
            t[4 % KPV] = v.z & 0xffff; t[5 % KPV] = v.z >> 16; t[6 % KPV] = v.w & 0xffff; t[7 % KPV] = v.w >> 16;
        } else {
            t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
        }
    } else {
        for (int k = 0; k < KPV; k++) t[k] = (base + k < L) ? (uint32_t)tiles[base + k] : 0u;
    }
    uint32_t prev = (base == 0) ? 0xffffffffu : (uint32_t)tiles[base - 1];
#pragma unroll
    for (int k = 0; k < KPV; k++) {
        const int idx = base + k;
        if (idx < L) {
            const uint32_t cur = t[k];
            if (idx == 0) {
                ranges[cur].x = 0;
            } else if (cur != prev) {
                ranges[prev].y = idx;
                ranges[cur].x = idx;
            }
            if (idx == L - 1) ranges[cur].y = L;
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

// ---- host side -------------------------------------------------------------------------------------
// Scratch words behind GeometryState::meta.  Zeroed together with it at the start of every forward:
//   hist[8][256]                     digit histograms: 4 depth passes, up to 4 tile passes
//   super_sum[ceil(P / 16384)]       sums of tiles_touched (depth order) over 16384 Gaussians
//   depth_status[4][tiles][256]      look-back words of the depth passes
// not zeroed (fully written before they are read):
//   block_sum[ceil(P / 256)], warp_sum[ceil(P / 32)]
namespace {
constexpr int kDepthPasses = 4;
constexpr int kDepthTile = EX_SORT_THREADS_DEPTH * EX_SORT_ITEMS_DEPTH;
constexpr int kTileTile = EX_SORT_THREADS_TILE * EX_SORT_ITEMS_TILE;

struct SortScratch {
    uint32_t *hist, *super_sum, *depth_status, *block_sum, *warp_sum;
    size_t depth_tiles, zero_words, words;
};
SortScratch sort_scratch(char* base, int P)
{
    SortScratch sc;
    const size_t n = (size_t)(P > 0 ? P : 0);
    auto up = [](size_t v) { return (v + 63) & ~size_t(63); };
    uint32_t* w = reinterpret_cast<uint32_t*>(base);
    sc.depth_tiles = (n + kDepthTile - 1) / kDepthTile;
    sc.hist = w;
    sc.super_sum = sc.hist + 8 * kBins;
    sc.depth_status = sc.super_sum + up((n + 16383) / 16384 + 1);
    sc.block_sum = sc.depth_status + (size_t)kDepthPasses * sc.depth_tiles * kBins;
    sc.zero_words = (size_t)(sc.block_sum - w);
    sc.warp_sum = sc.block_sum + up((n + 255) / 256);
    sc.words = (size_t)(sc.warp_sum - w) + up((n + 31) / 32);
    return sc;
}
// 8-bit passes over the bits of the largest tile id (the dump tile 0xFF..FF is dropped by the first pass: it needs no bits)
int tile_sort_passes(int grid_x, int grid_y)
{
    const int bit = (int)higher_msb((uint32_t)(grid_x * grid_y));
    return bit <= 8 ? 1 : (bit + 7) / 8;
}
}  // namespace

size_t binning_geometry_scratch_bytes(int P) { return sort_scratch(nullptr, P).words * sizeof(uint32_t) + 256; }
size_t binning_geometry_zero_bytes(int P) { return sort_scratch(nullptr, P).zero_words * sizeof(uint32_t); }
int binning_tile_passes(int grid_x, int grid_y) { return tile_sort_passes(grid_x, grid_y); }
// 16-bit tile keys while every tile id (and the all-ones dump key) fits, 32-bit keys for images of more than 65535 tiles
int binning_tile_key_bytes(int grid_x, int grid_y) { return (long long)grid_x * grid_y > 65535 ? 4 : 2; }
size_t binning_status_bytes(int cap, int passes)
{
    const size_t tiles = ((size_t)(cap > 0 ? cap : 0) + kTileTile - 1) / kTileTile;
    return (size_t)passes * tiles * kBins * sizeof(uint32_t) + 256;
}
int binning_duplicate_set(int grid_x, int grid_y) { return tile_sort_passes(grid_x, grid_y) & 1; }

// `out` must have been zeroed by the caller
cudaError_t launch_subpixel_absmax(const float* subpixel_offset, size_t n, uint32_t* out, cudaStream_t s)
{
    absmax_kernel<<<148 * 8, 256, 0, s>>>(subpixel_offset, n, out);
    return cudaGetLastError();
}

// depth order of the visible Gaussians: g.order[0 .. N_vis), N_vis in meta[EX_META_NVIS]; tiles_touched in that order
// (g.key_b) with its partial sums, the instance count R in meta[EX_META_TOTAL]
cudaError_t binning_depth_order(const GeometryState& g, int P, cudaStream_t s)
{
    if (P <= 0) return cudaSuccess;
    const SortScratch sc = sort_scratch(g.temp, P);
    uint32_t* const n_vis = g.meta + EX_META_NVIS;
    uint32_t* const err = g.meta + EX_META_ERROR;
    radix_hist_kernel<uint32_t, kDepthPasses, true><<<148, 1024, 0, s>>>(g.key_in, nullptr, (uint32_t)P, sc.hist, n_vis);
    const int grid = (int)sc.depth_tiles;
    // ping-pong between (key_a, val_a) and (key_b, order): the last pass leaves the ids in `order`, writes no keys and
    // gathers tiles_touched into key_b (depth order) for the duplicate kernel
    constexpr int T = EX_SORT_THREADS_DEPTH, I = EX_SORT_ITEMS_DEPTH, M = EX_SORT_MINBLOCKS_DEPTH;
    const size_t st = sc.depth_tiles * kBins;
    uint32_t* const tk = g.meta + EX_META_TICKETS;
    cudaError_t e = launch_pass<uint32_t, T, I, M, true, true, true>(grid, s, g.key_in, nullptr, g.key_a, g.val_a, nullptr, (uint32_t)P, 0,
                                                               sc.hist, sc.depth_status, tk + 0, err);
    if (e == cudaSuccess)
        e = launch_pass<uint32_t, T, I, M, false, false, true>(grid, s, g.key_a, g.val_a, g.key_b, g.order, n_vis, (uint32_t)P, 8,
                                                        sc.hist + kBins, sc.depth_status + st, tk + 1, err);
    if (e == cudaSuccess)
        e = launch_pass<uint32_t, T, I, M, false, false, true>(grid, s, g.key_b, g.order, g.key_a, g.val_a, n_vis, (uint32_t)P, 16,
                                                        sc.hist + 2 * kBins, sc.depth_status + 2 * st, tk + 2, err);
    if (e == cudaSuccess)
        e = launch_pass<uint32_t, T, I, M, false, false, false>(grid, s, g.key_a, g.val_a, nullptr, g.order, n_vis, (uint32_t)P, 24,
                                                         sc.hist + 3 * kBins, sc.depth_status + 3 * st, tk + 3, err, g.tiles_touched, g.key_b);
    if (e != cudaSuccess) return e;
    touched_sums_kernel<<<(P + 2047) / 2048, 256, 0, s>>>(n_vis, g.key_b, sc.warp_sum, sc.block_sum, sc.super_sum, g.meta + EX_META_TOTAL);
    return cudaGetLastError();
}

// before a second duplicate / tile sort of the same frame (the first buffer was too small): clear what the tile sort accumulates
cudaError_t binning_reset_instances(const GeometryState& g, int P, cudaStream_t s)
{
    const SortScratch sc = sort_scratch(g.temp, P);
    cudaError_t e = cudaMemsetAsync(g.meta + EX_META_TICKETS + 5, 0, 4 * sizeof(uint32_t), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(g.meta + EX_META_LISTED, 0, sizeof(uint32_t), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(sc.hist + 4 * kBins, 0, 4 * kBins * sizeof(uint32_t), s);
    return e;
}

// emit the (tile, id) pairs in depth order into set `binning_duplicate_set()` of the buffer; entries at positions
// >= cap are dropped (the buffer is sized before the count is known on the host)
cudaError_t binning_duplicate(const GeometryState& g, const BinningState& b, const int* radii, int P, int cap,
                              int grid_x, int grid_y, unsigned flags, cudaStream_t s)
{
    if (P <= 0) return cudaSuccess;
    const SortScratch sc = sort_scratch(g.temp, P);
    const int set = binning_duplicate_set(grid_x, grid_y);
    const int grid = (P + kDupThreads - 1) / kDupThreads;
    const float* const pad = reinterpret_cast<const float*>(g.meta);
    const uint32_t c = (uint32_t)(cap > 0 ? cap : 0);
    if (b.key_bytes == 2)
        duplicate_kernel<uint16_t><<<grid, kDupThreads, 0, s>>>(g.meta + EX_META_NVIS, g.order, g.key_b, sc.warp_sum, sc.block_sum, sc.super_sum,
                                                                g.rec, radii, grid_x, grid_y, flags, pad,
                                                                static_cast<uint16_t*>(b.tile[set]), b.val[set], c);
    else
        duplicate_kernel<uint32_t><<<grid, kDupThreads, 0, s>>>(g.meta + EX_META_NVIS, g.order, g.key_b, sc.warp_sum, sc.block_sum, sc.super_sum,
                                                                g.rec, radii, grid_x, grid_y, flags, pad,
                                                                static_cast<uint32_t*>(b.tile[set]), b.val[set], c);
    return cudaGetLastError();
}

// stable sort of the min(R, cap) pairs by tile into set 0 (point_list, tile_sorted), per-tile ranges
namespace {
template <typename KeyT>
cudaError_t sort_ranges(const GeometryState& g, const BinningState& b, const ImageState& img, int P, int cap, int passes, bool cull, cudaStream_t s)
{
    const SortScratch sc = sort_scratch(g.temp, P);
    const uint32_t* const total = g.meta + EX_META_TOTAL;
    uint32_t* const err = g.meta + EX_META_ERROR;
    // entries of the sorted list: with the exact tile test, what the histogram kernel counted (it skips the dump tile)
    const uint32_t* const listed = cull ? g.meta + EX_META_LISTED : total;
    const int grid = (cap + kTileTile - 1) / kTileTile;
    const size_t st = (size_t)grid * kBins;
    constexpr int T = EX_SORT_THREADS_TILE, I = EX_SORT_ITEMS_TILE, M = EX_SORT_MINBLOCKS_TILE;
    uint32_t* const tk = g.meta + EX_META_TICKETS;
    uint32_t* const hist = sc.hist + 4 * kBins;
    KeyT* const tile[2] = {static_cast<KeyT*>(b.tile[0]), static_cast<KeyT*>(b.tile[1])};
    int cur = passes & 1;                                   // the duplicate kernel's set; the last pass ends in set 0
    uint32_t* const n_listed = cull ? g.meta + EX_META_LISTED : nullptr;
    switch (passes * 2 + (cull ? 1 : 0)) {
        case 2: radix_hist_kernel<KeyT, 1, false><<<148, 1024, 0, s>>>(tile[cur], total, (uint32_t)cap, hist, n_listed); break;
        case 3: radix_hist_kernel<KeyT, 1, true><<<148, 1024, 0, s>>>(tile[cur], total, (uint32_t)cap, hist, n_listed); break;
        case 4: radix_hist_kernel<KeyT, 2, false><<<148, 1024, 0, s>>>(tile[cur], total, (uint32_t)cap, hist, n_listed); break;
        case 5: radix_hist_kernel<KeyT, 2, true><<<148, 1024, 0, s>>>(tile[cur], total, (uint32_t)cap, hist, n_listed); break;
        case 6: radix_hist_kernel<KeyT, 3, false><<<148, 1024, 0, s>>>(tile[cur], total, (uint32_t)cap, hist, n_listed); break;
        case 7: radix_hist_kernel<KeyT, 3, true><<<148, 1024, 0, s>>>(tile[cur], total, (uint32_t)cap, hist, n_listed); break;
        case 8: radix_hist_kernel<KeyT, 4, false><<<148, 1024, 0, s>>>(tile[cur], total, (uint32_t)cap, hist, n_listed); break;
        default: radix_hist_kernel<KeyT, 4, true><<<148, 1024, 0, s>>>(tile[cur], total, (uint32_t)cap, hist, n_listed); break;
    }
    cudaError_t e = cudaSuccess;
    for (int p = 0; p < passes && e == cudaSuccess; p++, cur ^= 1) {
        // the first pass reads the min(R, cap) emitted pairs and (exact tile culling) drops the dump-tile ones on the fly
        if (p == 0 && cull)
            e = launch_pass<KeyT, T, I, M, true, false, true>(grid, s, tile[cur], b.val[cur], tile[cur ^ 1], b.val[cur ^ 1], total, (uint32_t)cap, 0,
                                                              hist, b.status, tk + 5, err);
        else
            e = launch_pass<KeyT, T, I, M, false, false, true>(grid, s, tile[cur], b.val[cur], tile[cur ^ 1], b.val[cur ^ 1], p == 0 ? total : listed,
                                                               (uint32_t)cap, 8 * p, hist + p * kBins, b.status + p * st, tk + 5 + p, err);
    }
    if (e != cudaSuccess) return e;
    constexpr int KPV = 16 / (int)sizeof(KeyT);
    const int groups = (cap + KPV - 1) / KPV;
    tile_ranges_kernel<KeyT><<<(groups + 255) / 256, 256, 0, s>>>(listed, total, (uint32_t)cap, tile[0], img.ranges, err);
    return cudaGetLastError();
}
}  // namespace

cudaError_t binning_sort_ranges(const GeometryState& g, const BinningState& b, const ImageState& img, int P, int cap,
                                int grid_x, int grid_y, unsigned flags, cudaStream_t s)
{
    const int tiles = grid_x * grid_y;
    cudaError_t e = cudaMemsetAsync(img.ranges, 0, sizeof(uint2) * (size_t)tiles, s);
    if (e != cudaSuccess || cap <= 0 || P <= 0) return e;
    const int passes = tile_sort_passes(grid_x, grid_y);
    e = cudaMemsetAsync(b.status, 0, binning_status_bytes(cap, passes) - 256, s);
    if (e != cudaSuccess) return e;
    const bool cull = (flags & 1u) != 0;
    return b.key_bytes == 2 ? sort_ranges<uint16_t>(g, b, img, P, cap, passes, cull, s) : sort_ranges<uint32_t>(g, b, img, P, cap, passes, cull, s);
}
