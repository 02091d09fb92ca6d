// Per-tile backward of the alpha compositing for sm_100a.
//
// Behavioural spec: cuda_rasterizer/backward.cu:426-682 (renderCUDA): back-to-front over the
// tile's list, recomputing alpha and T, producing per-Gaussian gradients w.r.t. 2D mean (x, y and
// the depth slot z), conic, opacity, colour and dir3D - including the reference's deviations from
// the true derivative (Q3 depth term added before the `*= T`, Q4 cumulative
// dL_dacc *= T entering only opacity, Q5 no alpha-gradient from flow).
//
// B200 design (DESIGN.md "render backward"):
//  * the reference issues 14 float atomicAdd per contributing (pixel, splat) pair
//    (backward.cu:613-679).  Here the 13 distinct values of a splat are first reduced across the
//    32 pixels of a warp with a 16-value butterfly (16 SHFL instead of 13x5), the 8 warps of the
//    tile deposit their partial sums in private shared-memory slices (no shared atomics), and one
//    thread per (splat, 4-value group) folds the 8 slices and issues a single 16-byte vector
//    reduction (REDG.E.ADD.F32x4) into a 64-byte per-Gaussian accumulator: at most 4 global
//    reductions per (tile, splat) instead of 14 per (pixel, splat);
//  * traversal starts at the tile's largest `n_contrib` instead of the end of the range
//    (backward.cu:552-577 re-loads the whole range and skips entries one by one);
//  * splat records are gathered with TMA bulk copies (48 bytes per splat, mbarrier-tracked) one
//    sub-batch ahead.
#include "common.cuh"

namespace {

constexpr int kSub = 64;   // splats per sub-batch
constexpr int kRing = 4;   // staging buffers
constexpr int kAhead = 2;  // sub-batches in flight ahead of the one being consumed

__device__ __forceinline__ void red_add_f32(float* addr, float v)
{
    asm volatile("red.global.add.f32 [%0], %1;\n" ::"l"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ void red_add_v4(float* addr, float4 v)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// Reduce 16 per-lane value slots over the warp; afterwards lane L holds the total of slot
// k(L) = 8*bit4 + 4*bit3 + 2*bit2 + bit1 of L (both lanes of a pair hold the same total).
// Slots 7, 11 and 15 are unused by the caller: their exchanges are dropped or left unselected, the
// lanes that would hold their totals end up with don't-care values.
__device__ __forceinline__ float butterfly16(float (&v)[16], int lane)
{
    const unsigned full = 0xffffffffu;
    bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 7; i++) {
        if (i == 3) {                       // slot 11 unused: the upper half-warp's result is don't-care
            v[3] += __shfl_xor_sync(full, v[3], 16);
            continue;
        }
        const float mine = hi ? v[i + 8] : v[i];
        const float other = hi ? v[i] : v[i + 8];
        v[i] = mine + __shfl_xor_sync(full, other, 16);
    }
    hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (i == 3) {                       // slots 7 / 15 unused
            v[3] += __shfl_xor_sync(full, v[3], 8);
            continue;
        }
        const float mine = hi ? v[i + 4] : v[i];
        const float other = hi ? v[i] : v[i + 4];
        v[i] = mine + __shfl_xor_sync(full, other, 8);
    }
    hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const float mine = hi ? v[i + 2] : v[i];
        const float other = hi ? v[i] : v[i + 2];
        v[i] = mine + __shfl_xor_sync(full, other, 4);
    }
    hi = lane & 2;
    {
        const float mine = hi ? v[1] : v[0];
        const float other = hi ? v[0] : v[1];
        v[0] = mine + __shfl_xor_sync(full, other, 2);
    }
    v[0] += __shfl_xor_sync(full, v[0], 1);
    return v[0];
}

__global__ void __launch_bounds__(256, EX_BWD_MINBLOCKS) render_bwd_kernel(const __grid_constant__ RenderParams p)
{
    __shared__ float4 s_rec[kRing][kSub * 3];
    __shared__ uint8_t s_list[8][kSub];
    __shared__ int s_start;
    __shared__ __align__(8) unsigned long long s_full[kRing];     // records of a sub-batch have landed (TMA byte count)
    __shared__ __align__(8) unsigned long long s_empty[kRing];    // all 8 warps are done with the buffer

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.y * p.grid_x + blockIdx.x;
    const int pix_x = blockIdx.x * EX_TILE + (warp & 1) * 8 + (lane & 7);
    const int pix_y = blockIdx.y * EX_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = pix_x < p.W && pix_y < p.H;
    const int pix_id = p.W * pix_y + pix_x;
    const size_t HW = (size_t)p.H * p.W;

    const uint2 range = p.ranges[tile];
    if (tid == 0) {
        s_start = 0;
#pragma unroll
        for (int i = 0; i < kRing; i++) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], 8);
        }
        mbar_fence_init();
    }
    __syncthreads();

    float pxf = (float)pix_x, pyf = (float)pix_y;
    float T_final = 0.f, final_acc = 0.f, final_depth = 0.f;
    int last_contributor = 0;
    float dL_ddepth = 0.f, dL_dacc = 0.f;
    float dflow0 = 0.f, dflow1 = 0.f, dflow2 = 0.f;
    float dpix0 = 0.f, dpix1 = 0.f, dpix2 = 0.f;
    if (inside) {
        const float2 so = __ldg(p.subpixel_offset + pix_id);
        pxf = fa(pxf, so.x);
        pyf = fa(pyf, so.y);
        T_final = p.final_T[pix_id];
        last_contributor = (int)p.n_contrib[pix_id];
        final_acc = __ldg(p.out_acc + pix_id);
        final_depth = __ldg(p.out_depth + pix_id);
        dL_ddepth = __ldg(p.dL_ddepth + pix_id);
        if (final_acc > 0.0f) {
            dL_ddepth = dL_ddepth / final_acc;
            dflow0 = __ldg(p.dL_dflow + pix_id) / final_acc;
            dflow1 = __ldg(p.dL_dflow + HW + pix_id) / final_acc;
            dflow2 = __ldg(p.dL_dflow + 2 * HW + pix_id) / final_acc;
            dL_dacc = __ldg(p.dL_dacc + pix_id);
        }
        dpix0 = __ldg(p.dL_dpix + pix_id);
        dpix1 = __ldg(p.dL_dpix + HW + pix_id);
        dpix2 = __ldg(p.dL_dpix + 2 * HW + pix_id);
    }
    {
        const int wmax = __reduce_max_sync(0xffffffffu, last_contributor);
        if (lane == 0 && wmax > 0) atomicMax(&s_start, wmax);
    }
    __syncthreads();
    const int start = s_start;            // positions >= start contribute to no pixel of the tile
    if (start == 0) return;
    const int rounds = (start + kSub - 1) / kSub;

    // pixels that received nothing in the forward never contribute: keep them out of the warp's box
    const BlockBox box = block_box(pxf, pyf, last_contributor > 0);
    const bool warp_idle = __all_sync(0xffffffffu, last_contributor == 0);
    const int warp_last = __reduce_max_sync(0xffffffffu, last_contributor);   // list positions >= this are dead for the warp
    const float bg_dot_dpixel = __ldg(p.bg + 0) * dpix0 + __ldg(p.bg + 1) * dpix1 + __ldg(p.bg + 2) * dpix2;
    float T = T_final;
    float accum_rec0 = 0.f, accum_rec1 = 0.f, accum_rec2 = 0.f;
    float last_alpha = 0.f, last_c0 = 0.f, last_c1 = 0.f, last_c2 = 0.f;

    // TMA staging: every warp fetches 8 records of each sub-batch (lanes 0-7, one 48-byte bulk copy
    // each; the dir3D word is not needed here), kAhead sub-batches ahead; thread 0 announces the byte
    // count of the whole sub-batch to the buffer's `full` barrier.  No block-wide barrier in the loop:
    // a warp only waits for the records (full) and, before refilling a buffer, for the slowest warp to
    // have left it (empty) - with kRing buffers the warps of a tile may drift kRing - kAhead
    // sub-batches apart, which absorbs the imbalance between the 8x4 pixel blocks.
    const int slot = warp * 8 + lane;      // valid for lane < 8
    auto list_id = [&](int r) -> int {
        const int q = start - 1 - (r * kSub + slot);
        return (lane < 8 && r < rounds && q >= 0) ? (int)__ldg(p.point_list + range.x + q) : -1;
    };
    auto stage = [&](int r, int id) {
        const int buf = r % kRing;
        if (tid == 0) mbar_arrive_expect_tx(&s_full[buf], (unsigned)(min(kSub, start - r * kSub) * 48));
        if (id >= 0) tma_bulk_g2s(&s_rec[buf][slot * 3], p.rec + id, 48, &s_full[buf]);
    };
#pragma unroll
    for (int r = 0; r < kAhead; r++)
        if (r < rounds) stage(r, list_id(r));
    int id_next = list_id(kAhead);

    for (int r = 0; r < rounds; r++) {
        const int buf = r % kRing;
        if (r + kAhead < rounds) {
            const int rn = r + kAhead;                 // refills the buffer sub-batch rn - kRing used
            if (rn >= kRing) mbar_wait(&s_empty[rn % kRing], (unsigned)(((rn - kRing) / kRing) & 1));
            stage(rn, id_next);
            id_next = list_id(rn + 1);
        }
        mbar_wait(&s_full[buf], (unsigned)((r / kRing) & 1));      // sub-batch r has landed
        const float4* __restrict__ s = s_rec[buf];
        const int cnt = min(kSub, start - r * kSub);
        // which splats of the sub-batch can touch this warp's pixel block at all (exact, see block_reject)
        int nw = 0;
        if (!warp_idle && (start - r * kSub - cnt) < warp_last) {      // some entry of the sub-batch is still live for this warp
            for (int g = 0; g < cnt; g += 32) {
                const int jj = g + lane;
                bool keep = false;
                // entry jj sits at list position q = start-1-(r*kSub+jj); only q < warp_last can matter
                if (jj < cnt && (start - 1 - (r * kSub + jj)) < warp_last) keep = !EX_BLOCK_TEST(s[jj * 3], s[jj * 3 + 1], box);
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (keep) s_list[warp][nw + __popc(m & ((1u << lane) - 1u))] = (uint8_t)jj;
                nw += __popc(m);
            }
            __syncwarp();
        }
#pragma unroll 2
        for (int e = 0; e < nw; e++) {
            const int j = s_list[warp][e];
            const int q = start - 1 - (r * kSub + j);
            const float4 a = s[j * 3 + 0];
            const float4 b = s[j * 3 + 1];
            const float dx = fa(a.x, -pxf), dy = fa(a.y, -pyf);
            const float power = ff(ff(dx, fm(dx, b.x), fm(fm(b.z, dy), dy)), -0.5f, -fm(fm(b.y, dx), dy));
            bool contributes = (q < last_contributor) && !(power > 0.0f) && !(power < a.w);
            float G = 0.f, alpha = 0.f;
            if (contributes) {
                G = expf(power);
                alpha = fminf(0.99f, fm(b.w, G));
                contributes = !(alpha < 1.0f / 255.0f);
            }
            if (!__any_sync(0xffffffffu, contributes)) continue;
            const float4 c = s[j * 3 + 2];
            float v[16];
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = 0.f;
            if (contributes) {
                const float inv1ma = 1.f / (1.f - alpha);
                T = T * inv1ma;
                const float w = alpha * T;               // dchannel_dcolor
                float dL_dalpha = 0.0f;
                const float dep = a.z;
                if ((dep > p.min_depth) & (w > 0.0f)) {
                    v[2] = dL_ddepth * w;
                    dL_dalpha += (final_depth - dep) * dL_ddepth * T;
                }
                accum_rec0 = last_alpha * last_c0 + (1.f - last_alpha) * accum_rec0;
                accum_rec1 = last_alpha * last_c1 + (1.f - last_alpha) * accum_rec1;
                accum_rec2 = last_alpha * last_c2 + (1.f - last_alpha) * accum_rec2;
                last_c0 = c.x; last_c1 = c.y; last_c2 = c.z;
                dL_dalpha += (c.x - accum_rec0) * dpix0;
                dL_dalpha += (c.y - accum_rec1) * dpix1;
                dL_dalpha += (c.z - accum_rec2) * dpix2;
                v[8] = w * dpix0; v[9] = w * dpix1; v[10] = w * dpix2;
                v[12] = w * dflow0; v[13] = w * dflow1; v[14] = w * dflow2;
                dL_dalpha *= T;
                dL_dacc *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final * inv1ma) * bg_dot_dpixel;
                // constant factors (-W/2, -H/2, -1/2) are applied once per Gaussian by gacc_load()
                const float gG = G * (b.w * dL_dalpha);
                const float X = gG * dx, Y = gG * dy;
                v[0] = X * b.x + Y * b.y;
                v[1] = Y * b.z + X * b.y;
                v[4] = X * dx;
                v[5] = X * dy;
                v[6] = Y * dy;
                v[3] = G * dL_dalpha + G * dL_dacc;
            }
            const float tot = butterfly16(v, lane);
            // even lanes hold the 16 totals; 13 of them are real.  One warp-level reduction instruction:
            // 13 lanes add into the 64-byte accumulator of the splat (2 sectors).
            const int k = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            if (!(lane & 1) && (k == 3 || (k & 3) != 3))
                red_add_f32(reinterpret_cast<float*>(p.gacc + __float_as_int(c.w)) + k, tot);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[buf]);     // this warp no longer reads buffer `buf`
    }
}


}  // namespace

void launch_render_bwd(const RenderParams& p, int grid_x, int grid_y, cudaStream_t s)
{
    dim3 grid(grid_x, grid_y, 1);
    render_bwd_kernel<<<grid, 256, 0, s>>>(p);
}
