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
//    (backward.cu:613-679).  Here a lane owns PPT = 2 pixels, sums their gradient terms in registers, the 13
//    distinct values of a splat are reduced across the warp with a 16-value butterfly (16 SHFL instead of
//    13x5) and 13 lanes issue ONE warp-level reduction instruction (REDG.E.ADD.F32) into the splat's 64-byte
//    accumulator line: one butterfly + one reduction per (64 pixels, splat) instead of 14 atomics per
//    (pixel, splat); constant factors (-W/2, -H/2, -1/2) are applied once per Gaussian when the line is read;
//  * traversal starts at the tile's largest `n_contrib` instead of the end of the range
//    (backward.cu:552-577 re-loads the whole range and skips entries one by one);
//  * splat records are gathered with TMA bulk copies (48 bytes per splat, mbarrier-tracked) kAhead
//    sub-batches ahead through a ring of kRing buffers guarded by full / empty mbarriers: no block barrier in
//    the loop, the warps of a tile may drift kRing - kAhead sub-batches apart;
//  * per warp and sub-batch the splats that cannot touch the warp's 8x8 pixel block are dropped first
//    (block_reject, exact), as in the forward;
//  * NULL upstream gradients for depth / acc (autograd: "not used by the loss") select an instantiation
//    without those terms.
#include "common.cuh"

namespace {

#ifndef EX_BWD_SUB
#define EX_BWD_SUB 64
#endif
#ifndef EX_BWD_RING
#define EX_BWD_RING 4
#endif
#ifndef EX_BWD_AHEAD
#define EX_BWD_AHEAD 2
#endif
#ifndef EX_BWD_UNROLL
#define EX_BWD_UNROLL 1     // entry loop not unrolled: 0.834 vs 0.856 ms at C3 (the doubled body thrashes the instruction cache)
#endif
constexpr int kSub = EX_BWD_SUB;     // splats per sub-batch
constexpr int kRing = EX_BWD_RING;   // staging buffers
constexpr int kAhead = EX_BWD_AHEAD; // sub-batches in flight ahead of the one being consumed
constexpr int kUnroll = EX_BWD_UNROLL;

__device__ __forceinline__ void red_add_f32(float* addr, float v)
{
    asm volatile("red.global.add.f32 [%0], %1;\n" ::"l"(addr), "f"(v) : "memory");
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

// first pixel of a lane assigns, further pixels add (keeps the PPT = 1 code free of `0 + x` adds)
__device__ __forceinline__ void acc_to(float& dst, float x, int u)
{
    if (u == 0) dst = x; else dst += x;
}

// PPT = pixels per thread.  PPT = 1: 8 warps, each an 8x4 pixel block.  PPT = 2: 4 warps, each an
// 8x8 block (a lane owns pixels (x, y) and (x, y + 4)): the per-lane gradient terms of the two
// pixels are added before the warp butterfly, so one butterfly + one reduction instruction serves
// 64 pixels instead of 32.
// DA = false: the caller has no upstream gradient for the depth and accumulated-alpha images (NULL
// dL_ddepth and dL_dacc - what autograd reports when the loss does not use them, as in train.py): their
// terms are exactly zero and are dropped at compile time.  A NULL dL_dflow is read as zeros.
template <int PPT, bool DA>
__global__ void __launch_bounds__(256 / PPT, EX_BWD_MINBLOCKS) render_bwd_kernel(const __grid_constant__ RenderParams p)
{
    constexpr int NW = 8 / PPT;            // warps per tile
    constexpr int SLOTS = kSub / NW;       // records each warp fetches per sub-batch
    __shared__ float4 s_rec[kRing][kSub * 3];
    __shared__ uint8_t s_list[NW][kSub];
    __shared__ int s_start;
    __shared__ __align__(8) unsigned long long s_full[kRing];     // records of a sub-batch have landed (TMA byte count)
    __shared__ __align__(8) unsigned long long s_empty[kRing];    // all warps are done with the buffer

    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.y * p.grid_x + blockIdx.x;
    const size_t HW = (size_t)p.H * p.W;

    const uint2 range = p.ranges[tile];
    if (tid == 0) {
        s_start = 0;
#pragma unroll
        for (int i = 0; i < kRing; i++) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], NW);
        }
        mbar_fence_init();
    }
    __syncthreads();

    float pxf[PPT], pyf[PPT], T_final[PPT], final_depth[PPT], dL_ddepth[PPT], dL_dacc[PPT];
    float dflow0[PPT], dflow1[PPT], dflow2[PPT], dpix0[PPT], dpix1[PPT], dpix2[PPT];
    int last_contributor[PPT];
    int lane_last = 0;
#pragma unroll
    for (int u = 0; u < PPT; u++) {
        const int pix_x = blockIdx.x * EX_TILE + (warp & 1) * 8 + (lane & 7);
        const int pix_y = blockIdx.y * EX_TILE + (warp >> 1) * (4 * PPT) + (lane >> 3) + 4 * u;
        const bool inside = pix_x < p.W && pix_y < p.H;
        const int pix_id = p.W * pix_y + pix_x;
        pxf[u] = (float)pix_x; pyf[u] = (float)pix_y;
        T_final[u] = 0.f; final_depth[u] = 0.f; dL_ddepth[u] = 0.f; dL_dacc[u] = 0.f;
        dflow0[u] = dflow1[u] = dflow2[u] = 0.f;
        dpix0[u] = dpix1[u] = dpix2[u] = 0.f;
        last_contributor[u] = 0;
        if (inside) {
            const float2 so = __ldg(p.subpixel_offset + pix_id);
            pxf[u] = fa(pxf[u], so.x);
            pyf[u] = fa(pyf[u], so.y);
            T_final[u] = p.final_T[pix_id];
            last_contributor[u] = (int)p.n_contrib[pix_id];
            const float final_acc = __ldg(p.out_acc + pix_id);
            if (DA) {
                final_depth[u] = __ldg(p.out_depth + pix_id);
                if (p.dL_ddepth) dL_ddepth[u] = __ldg(p.dL_ddepth + pix_id);
            }
            if (final_acc > 0.0f) {
                if (DA) dL_ddepth[u] = dL_ddepth[u] / final_acc;
                if (p.dL_dflow) {
                    dflow0[u] = __ldg(p.dL_dflow + pix_id) / final_acc;
                    dflow1[u] = __ldg(p.dL_dflow + HW + pix_id) / final_acc;
                    dflow2[u] = __ldg(p.dL_dflow + 2 * HW + pix_id) / final_acc;
                }
                if (DA && p.dL_dacc) dL_dacc[u] = __ldg(p.dL_dacc + pix_id);
            }
            dpix0[u] = __ldg(p.dL_dpix + pix_id);
            dpix1[u] = __ldg(p.dL_dpix + HW + pix_id);
            dpix2[u] = __ldg(p.dL_dpix + 2 * HW + pix_id);
        }
        lane_last = max(lane_last, last_contributor[u]);
    }
    const int warp_last = __reduce_max_sync(full, lane_last);   // list positions >= this are dead for the warp
    if (lane == 0 && warp_last > 0) atomicMax(&s_start, warp_last);
    __syncthreads();
    const int start = s_start;            // positions >= start contribute to no pixel of the tile
    if (start == 0) return;
    const int rounds = (start + kSub - 1) / kSub;

    // pixels that received nothing in the forward never contribute: keep them out of the warp's box
    BlockBox box;
    {
        float x0 = 3.0e38f, x1 = -3.0e38f, y0 = 3.0e38f, y1 = -3.0e38f;
#pragma unroll
        for (int u = 0; u < PPT; u++)
            if (last_contributor[u] > 0) {
                x0 = fminf(x0, pxf[u]); x1 = fmaxf(x1, pxf[u]);
                y0 = fminf(y0, pyf[u]); y1 = fmaxf(y1, pyf[u]);
            }
        box = block_box_merge(x0, x1, y0, y1);
    }
    const bool warp_idle = warp_last == 0;
    const float bg0 = __ldg(p.bg + 0), bg1 = __ldg(p.bg + 1), bg2 = __ldg(p.bg + 2);
    float bg_dot_dpixel[PPT], T[PPT];
    float accum_rec0[PPT], accum_rec1[PPT], accum_rec2[PPT];
    float last_alpha[PPT], last_c0[PPT], last_c1[PPT], last_c2[PPT];
#pragma unroll
    for (int u = 0; u < PPT; u++) {
        bg_dot_dpixel[u] = bg0 * dpix0[u] + bg1 * dpix1[u] + bg2 * dpix2[u];
        T[u] = T_final[u];
        accum_rec0[u] = accum_rec1[u] = accum_rec2[u] = 0.f;
        last_alpha[u] = last_c0[u] = last_c1[u] = last_c2[u] = 0.f;
    }

    // TMA staging: every warp fetches SLOTS records of each sub-batch (lanes 0..SLOTS-1, one 48-byte
    // bulk copy each; the dir3D word is not needed here), kAhead sub-batches ahead; thread 0 announces
    // the byte count of the whole sub-batch to the buffer's `full` barrier.  No block-wide barrier in
    // the loop: a warp only waits for the records (full) and, before refilling a buffer, for the
    // slowest warp to have left it (empty) - with kRing buffers the warps of a tile may drift
    // kRing - kAhead sub-batches apart, which absorbs the imbalance between the pixel blocks.
    const int slot = warp * SLOTS + lane;      // valid for lane < SLOTS
    auto list_id = [&](int r) -> int {
        const int q = start - 1 - (r * kSub + slot);
        return (lane < SLOTS && r < rounds && q >= 0) ? (int)__ldg(p.point_list + range.x + q) : -1;
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
                const unsigned m = __ballot_sync(full, keep);
                if (keep) s_list[warp][nw + __popc(m & ((1u << lane) - 1u))] = (uint8_t)jj;
                nw += __popc(m);
            }
            __syncwarp();
        }
        // entry j of the sub-batch sits at list position q = start-1-(r*kSub+j):  q < last_contributor  <=>  j >= jthr
        int jthr[PPT];
#pragma unroll
        for (int u = 0; u < PPT; u++) jthr[u] = start - r * kSub - last_contributor[u];
        const unsigned sb = smem_u32(s);
#pragma unroll(kUnroll)
        for (int e = 0; e < nw; e++) {
            const int j = s_list[warp][e];
            const float4 a = lds128(sb + j * 48);
            const float4 b = lds128(sb + j * 48 + 16);
            float dx[PPT], dy[PPT], G[PPT], alpha[PPT];
            bool contributes[PPT];
            bool any_c = false;
#pragma unroll
            for (int u = 0; u < PPT; u++) {
                dx[u] = fa(a.x, -pxf[u]);
                dy[u] = fa(a.y, -pyf[u]);
                const float power = ff(ff(dx[u], fm(dx[u], b.x), fm(fm(b.z, dy[u]), dy[u])), -0.5f, -fm(fm(b.y, dx[u]), dy[u]));
                contributes[u] = (j >= jthr[u]) && !(power > 0.0f) && !(power < a.w);
                G[u] = 0.f; alpha[u] = 0.f;
                if (contributes[u]) {
#if EX_BWD_FAST_EXP
                    // ex2.approx(power * log2 e): 2 instructions instead of libdevice's 10; relative error < 1e-6 for
                    // power in [-6, 0], three orders of magnitude inside the gradient budget (the FORWARD keeps expf:
                    // its images are bit-identical to the reference's)
                    G[u] = __expf(power);
#else
                    G[u] = expf(power);
#endif
                    alpha[u] = fminf(0.99f, fm(b.w, G[u]));
                    contributes[u] = !(alpha[u] < 1.0f / 255.0f);
                }
                any_c |= contributes[u];
            }
            if (!__any_sync(full, any_c)) continue;
            const float4 c = lds128(sb + j * 48 + 32);
            float v[16];
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = 0.f;
#pragma unroll
            for (int u = 0; u < PPT; u++) {
                if (contributes[u]) {
#if EX_BWD_FAST_RCP
                    // 1 - alpha is in [0.01, 1]: MUFU.RCP (1 ulp) instead of the 12-instruction IEEE division
                    float inv1ma;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv1ma) : "f"(1.f - alpha[u]));
#else
                    const float inv1ma = 1.f / (1.f - alpha[u]);
#endif
                    T[u] = T[u] * inv1ma;
                    const float w = alpha[u] * T[u];               // dchannel_dcolor
                    float dL_dalpha = 0.0f;
                    const float dep = a.z;
                    if (DA) {
                        if ((dep > p.min_depth) & (w > 0.0f)) {
                            acc_to(v[2], dL_ddepth[u] * w, u);
                            dL_dalpha += (final_depth[u] - dep) * dL_ddepth[u] * T[u];
                        }
                    }
                    accum_rec0[u] = last_alpha[u] * last_c0[u] + (1.f - last_alpha[u]) * accum_rec0[u];
                    accum_rec1[u] = last_alpha[u] * last_c1[u] + (1.f - last_alpha[u]) * accum_rec1[u];
                    accum_rec2[u] = last_alpha[u] * last_c2[u] + (1.f - last_alpha[u]) * accum_rec2[u];
                    last_c0[u] = c.x; last_c1[u] = c.y; last_c2[u] = c.z;
                    dL_dalpha += (c.x - accum_rec0[u]) * dpix0[u];
                    dL_dalpha += (c.y - accum_rec1[u]) * dpix1[u];
                    dL_dalpha += (c.z - accum_rec2[u]) * dpix2[u];
                    acc_to(v[8], w * dpix0[u], u); acc_to(v[9], w * dpix1[u], u); acc_to(v[10], w * dpix2[u], u);
                    acc_to(v[12], w * dflow0[u], u); acc_to(v[13], w * dflow1[u], u); acc_to(v[14], w * dflow2[u], u);
                    dL_dalpha *= T[u];
                    if (DA) dL_dacc[u] *= T[u];
                    last_alpha[u] = alpha[u];
                    dL_dalpha += (-T_final[u] * inv1ma) * bg_dot_dpixel[u];
                    // constant factors (-W/2, -H/2, -1/2) are applied once per Gaussian by gacc_load()
                    const float gG = G[u] * (b.w * dL_dalpha);
                    const float X = gG * dx[u], Y = gG * dy[u];
                    acc_to(v[0], X * b.x + Y * b.y, u);
                    acc_to(v[1], Y * b.z + X * b.y, u);
                    acc_to(v[4], X * dx[u], u);
                    acc_to(v[5], X * dy[u], u);
                    acc_to(v[6], Y * dy[u], u);
                    if (DA) acc_to(v[3], G[u] * dL_dalpha + G[u] * dL_dacc[u], u);
                    else acc_to(v[3], G[u] * dL_dalpha, u);
                }
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
    if (p.dL_ddepth || p.dL_dacc) render_bwd_kernel<EX_BWD_PPT, true><<<grid, 256 / EX_BWD_PPT, 0, s>>>(p);
    else render_bwd_kernel<EX_BWD_PPT, false><<<grid, 256 / EX_BWD_PPT, 0, s>>>(p);
}
