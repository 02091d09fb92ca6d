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
#include <string.h>

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
// lanes that would hold their totals end up with don't-care values.  The additions of neighbouring slots
// are issued as packed pairs (FADD2).
__device__ __forceinline__ float butterfly16(float (&v)[16], int lane)
{
    const unsigned full = 0xffffffffu;
    float m[8], r[8];
    bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 7; i++) {
        if (i == 3) {                       // slot 11 unused: the upper half-warp's result is don't-care
            m[3] = v[3];
            r[3] = __shfl_xor_sync(full, v[3], 16);
            continue;
        }
        m[i] = hi ? v[i + 8] : v[i];
        r[i] = __shfl_xor_sync(full, hi ? v[i] : v[i + 8], 16);
    }
    split2(fa2(mk2(m[0], m[1]), mk2(r[0], r[1])), v[0], v[1]);
    split2(fa2(mk2(m[2], m[3]), mk2(r[2], r[3])), v[2], v[3]);
    split2(fa2(mk2(m[4], m[5]), mk2(r[4], r[5])), v[4], v[5]);
    v[6] = m[6] + r[6];
    hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (i == 3) {                       // slots 7 / 15 unused
            m[3] = v[3];
            r[3] = __shfl_xor_sync(full, v[3], 8);
            continue;
        }
        m[i] = hi ? v[i + 4] : v[i];
        r[i] = __shfl_xor_sync(full, hi ? v[i] : v[i + 4], 8);
    }
    split2(fa2(mk2(m[0], m[1]), mk2(r[0], r[1])), v[0], v[1]);
    split2(fa2(mk2(m[2], m[3]), mk2(r[2], r[3])), v[2], v[3]);
    hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        m[i] = hi ? v[i + 2] : v[i];
        r[i] = __shfl_xor_sync(full, hi ? v[i] : v[i + 2], 4);
    }
    split2(fa2(mk2(m[0], m[1]), mk2(r[0], r[1])), v[0], v[1]);
    hi = lane & 2;
    {
        const float mine = hi ? v[1] : v[0];
        const float other = hi ? v[0] : v[1];
        v[0] = mine + __shfl_xor_sync(full, other, 2);
    }
    v[0] += __shfl_xor_sync(full, v[0], 1);
    return v[0];
}

__device__ __forceinline__ float fast_rcp(float x)
{
#if EX_BWD_FAST_RCP
    // 1 - alpha is in [0.01, 1]: MUFU.RCP (1 ulp) instead of the 12-instruction IEEE division
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.f / x;
#endif
}

// G = exp(power), alpha = min(0.99, opacity * G) of the lane's two pixels (packed).  Whether a pair CONTRIBUTED is known
// before this is called: the record's threshold a.w is the exact crossing of the forward's alpha >= 1/255 test
// (preprocess.cu alpha_threshold), so `power >= thr` decides membership exactly like the forward and the reference
// (forward.cu:384-387) and exp() only supplies values.
// EX_BWD_FAST_EXP: ex2.approx(power * log2 e) - MUFU.EX2 directly instead of libdevice's 10 instructions, relative
// error < 1e-6 (power lies in [thr, 0], thr > -6: no denormal handling needed).
__device__ __forceinline__ void pair_alpha2(f2 pw, float opac, bool c0, bool c1, f2& G, f2& al)
{
    float G0, G1, a0, a1;
#if EX_BWD_FAST_EXP
    float e0, e1;
    split2(fm2(pw, bc(1.4426950408889634f)), e0, e1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(G0) : "f"(e0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(G1) : "f"(e1));
#else
    G0 = expf(lo2(pw)); G1 = expf(hi2(pw));
#endif
    // a pixel the splat does not contribute to runs along as a phantom with G = alpha = 0
    G = mk2(c0 ? G0 : 0.f, c1 ? G1 : 0.f);
    split2(fm2(bc(opac), G), a0, a1);
    al = mk2(fminf(0.99f, a0), fminf(0.99f, a1));
}

// 4 warps per tile, each an 8x8 pixel block: a lane owns pixels (x, y) and (x, y + 4), and every per-pixel
// quantity lives in a packed pair (f2: low half = upper pixel) processed with FFMA2 / FMUL2 / FADD2 - one issued
// instruction for both pixels.  A pixel the splat does not contribute to runs along as a phantom with
// G = alpha = 0: T * 1, accum_rec * 1 + 0 and every gradient term * 0 leave its state and the sums unchanged bit
// for bit, so there are no per-pixel branches in the gradient math.  The per-lane terms of the two pixels are
// added before the warp butterfly: one butterfly + one reduction instruction serves 64 pixels.
// DA = false: the caller has no upstream gradient for the depth and accumulated-alpha images (NULL
// dL_ddepth and dL_dacc - what autograd reports when the loss does not use them, as in train.py): their
// terms are exactly zero and are dropped at compile time.  A NULL dL_dflow is read as zeros.
#ifdef EX_BWD_HIST
__device__ unsigned long long g_bwd_hist[66];     // [0..32]: entries by number of lanes with a contributing pixel; [33..65]: by number of contributing pixels / 2
#endif

template <bool DA>
__global__ void __launch_bounds__(128, EX_BWD_MINBLOCKS) render_bwd_kernel(const __grid_constant__ RenderParams p,
                                                                            const __grid_constant__ CUtensorMap rec_map)
{
    constexpr int PPT = 2;
    constexpr int NW = 4;                  // warps per tile
    constexpr int SLOTS = kSub / NW;       // records each warp fetches per sub-batch
    constexpr int RS = EX_BWD_STAGE_GATHER4 ? 4 : 3;      // float4 words per staged record (the gather moves whole 64-byte rows)
    constexpr unsigned RB = RS * 16;                       // bytes per staged record
    static_assert(!EX_BWD_STAGE_GATHER4 || (SLOTS % 4 == 0), "a warp's slots are gathered four at a time");
    __shared__ __align__(128) float4 s_rec[kRing][kSub * RS];
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
        const int pix_y = blockIdx.y * EX_TILE + (warp >> 1) * 8 + (lane >> 3) + 4 * u;
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

    // per-pixel state and constants of the lane's two pixels, packed
    const f2 npx = mk2(-pxf[0], -pxf[1]), npy = mk2(-pyf[0], -pyf[1]);
    const f2 dp0 = mk2(dpix0[0], dpix0[1]), dp1 = mk2(dpix1[0], dpix1[1]), dp2 = mk2(dpix2[0], dpix2[1]);
    const f2 df0 = mk2(dflow0[0], dflow0[1]), df1 = mk2(dflow1[0], dflow1[1]), df2 = mk2(dflow2[0], dflow2[1]);
    // -T_final * (bg . dL_dpix): factor of 1 / (1 - alpha) in the background term (backward.cu:656-659)
    const f2 nTbg = mk2(-T_final[0] * (bg0 * dpix0[0] + bg1 * dpix1[0] + bg2 * dpix2[0]),
                        -T_final[1] * (bg0 * dpix0[1] + bg1 * dpix1[1] + bg2 * dpix2[1]));
    const f2 fdep = mk2(final_depth[0], final_depth[1]), ddep = mk2(dL_ddepth[0], dL_ddepth[1]);
    f2 dacc = mk2(dL_dacc[0], dL_dacc[1]);
    f2 T = mk2(T_final[0], T_final[1]);
    f2 ar0 = bc(0.f), ar1 = bc(0.f), ar2 = bc(0.f);       // colour accumulated behind the current splat (accum_rec)

    // TMA staging: every warp fetches SLOTS records of each sub-batch, kAhead sub-batches ahead; thread 0 announces the byte
    // count of the whole sub-batch to the buffer's `full` barrier.  EX_BWD_STAGE_GATHER4: lanes 0..SLOTS/4-1 issue ONE
    // tile::gather4 each (four 64-byte records per instruction, row index = Gaussian id; slots beyond the list re-fetch
    // record 0 and are never read); otherwise lanes 0..SLOTS-1 issue one 48-byte bulk copy each.  No block-wide barrier in
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
        const int cnt_r = min(kSub, start - r * kSub);
#if EX_BWD_STAGE_GATHER4
        if (tid == 0) {
            int groups = 0;
#pragma unroll
            for (int w = 0; w < NW; w++) groups += (max(0, min(SLOTS, cnt_r - w * SLOTS)) + 3) >> 2;
            mbar_arrive_expect_tx(&s_full[buf], (unsigned)(groups * 256));
        }
        const int i0 = __shfl_sync(full, id, (4 * lane) & 31), i1 = __shfl_sync(full, id, (4 * lane + 1) & 31);
        const int i2 = __shfl_sync(full, id, (4 * lane + 2) & 31), i3 = __shfl_sync(full, id, (4 * lane + 3) & 31);
        if (lane < SLOTS / 4 && i0 >= 0)
            tma_gather4_g2s(&s_rec[buf][(warp * SLOTS + 4 * lane) * RS], &rec_map, i0, max(i1, 0), max(i2, 0), max(i3, 0), &s_full[buf]);
#else
        if (tid == 0) mbar_arrive_expect_tx(&s_full[buf], (unsigned)(cnt_r * 48));
        if (id >= 0) tma_bulk_g2s(&s_rec[buf][slot * RS], p.rec + id, 48, &s_full[buf]);
#endif
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
                if (jj < cnt && (start - 1 - (r * kSub + jj)) < warp_last) keep = !EX_BLOCK_TEST(s[jj * RS], s[jj * RS + 1], box);
                const unsigned m = __ballot_sync(full, keep);
                if (keep) s_list[warp][nw + __popc(m & ((1u << lane) - 1u))] = (uint8_t)jj;
                nw += __popc(m);
            }
            __syncwarp();
        }
        // entry j of the sub-batch sits at list position q = start-1-(r*kSub+j):  q < last_contributor  <=>  j >= jthr
        const int jthr0 = start - r * kSub - last_contributor[0], jthr1 = start - r * kSub - last_contributor[1];
        const unsigned sb = smem_u32(s);
#pragma unroll(kUnroll)
        for (int e = 0; e < nw; e++) {
            const int j = s_list[warp][e];
            const float4 a = lds128(sb + j * RB);
            const float4 b = lds128(sb + j * RB + 16);
            // power of both pixels with the forward's FMA placement (render_fwd.cu pair_power), packed
            const f2 dx = fa2(bc(a.x), npx), dy = fa2(bc(a.y), npy);
            const f2 pw = ff2(ff2(dx, fm2(dx, bc(b.x)), fm2(fm2(bc(b.z), dy), dy)), bc(-0.5f), fm2(fm2(bc(-b.y), dx), dy));
            float pw0, pw1;
            split2(pw, pw0, pw1);
            bool c0 = (j >= jthr0) && !(pw0 > 0.0f) && !(pw0 < a.w);
            bool c1 = (j >= jthr1) && !(pw1 > 0.0f) && !(pw1 < a.w);
            if (!__any_sync(full, c0 | c1)) continue;
#ifdef EX_BWD_HIST
            {
                const unsigned m = __ballot_sync(full, c0 | c1);
                const int npix = __popc(__ballot_sync(full, c0)) + __popc(__ballot_sync(full, c1));
                if (lane == 0) { atomicAdd(&g_bwd_hist[__popc(m)], 1ull); atomicAdd(&g_bwd_hist[33 + npix / 2], 1ull); }
            }
#endif
            f2 G, al;
            pair_alpha2(pw, b.w, c0, c1, G, al);
            const float4 c = lds128(sb + j * RB + 32);
            const f2 oma = ff2(al, bc(-1.0f), bc(1.0f));                // 1 - alpha (exact product: == fa(1, -alpha))
            float om0, om1;
            split2(oma, om0, om1);
            const f2 inv = mk2(fast_rcp(om0), fast_rcp(om1));
            T = fm2(T, inv);
            const f2 w = fm2(al, T);                                    // dchannel_dcolor
            f2 dLa;
            float v[16];
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = 0.f;
            // colour: dL_dalpha = sum_ch (c_ch - accum_rec_ch) dL_dpix_ch, then fold this splat into accum_rec
            // (accum_rec <- alpha c + (1 - alpha) accum_rec, written as accum_rec + alpha (c - accum_rec))
            const f2 t0 = ff2(ar0, bc(-1.0f), bc(c.x)), t1 = ff2(ar1, bc(-1.0f), bc(c.y)), t2 = ff2(ar2, bc(-1.0f), bc(c.z));
            dLa = ff2(t2, dp2, ff2(t1, dp1, fm2(t0, dp0)));
            ar0 = ff2(al, t0, ar0);
            ar1 = ff2(al, t1, ar1);
            ar2 = ff2(al, t2, ar2);
            if (DA) {
                // depth (backward.cu:604-622; its alpha term enters BEFORE the `*= T`, deviation Q3)
                float w0, w1, T0, T1;
                split2(w, w0, w1);
                split2(T, T0, T1);
                const float dep = a.z;
                const bool d0 = (dep > p.min_depth) & (w0 > 0.0f), d1 = (dep > p.min_depth) & (w1 > 0.0f);
                const f2 dd = fm2(fm2(fa2(fdep, bc(-dep)), ddep), T);
                v[2] = (d0 ? dL_ddepth[0] * w0 : 0.f) + (d1 ? dL_ddepth[1] * w1 : 0.f);
                dLa = fa2(dLa, mk2(d0 ? lo2(dd) : 0.f, d1 ? hi2(dd) : 0.f));
                // cumulative dL_dacc *= T over the contributors (deviation Q4): phantoms must not touch it
                dacc = fm2(dacc, mk2(c0 ? T0 : 1.0f, c1 ? T1 : 1.0f));
            }
            dLa = fm2(dLa, T);
            dLa = ff2(nTbg, inv, dLa);                                  // background term
            // what is constant per Gaussian (the 2x2 conic of dL/dmean2D, -W/2, -H/2, -1/2) is applied by gacc_load()
            const f2 gG = fm2(G, fm2(bc(b.w), dLa));
            const f2 X = fm2(gG, dx), Y = fm2(gG, dy);
            v[0] = hsum2(X);
            v[1] = hsum2(Y);
            v[4] = hsum2(fm2(X, dx));
            v[5] = hsum2(fm2(X, dy));
            v[6] = hsum2(fm2(Y, dy));
            if (DA) v[3] = hsum2(ff2(G, dacc, fm2(G, dLa)));
            else v[3] = hsum2(fm2(G, dLa));
            v[8] = hsum2(fm2(w, dp0)); v[9] = hsum2(fm2(w, dp1)); v[10] = hsum2(fm2(w, dp2));
            v[12] = hsum2(fm2(w, df0)); v[13] = hsum2(fm2(w, df1)); v[14] = hsum2(fm2(w, df2));
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

#ifdef EX_BWD_HIST
extern "C" void ex4dgs_debug_bwd_hist(unsigned long long* out)
{
    cudaMemcpyFromSymbol(out, g_bwd_hist, sizeof(g_bwd_hist));
    unsigned long long z[66] = {0};
    cudaMemcpyToSymbol(g_bwd_hist, z, sizeof(z));
}
#endif

void launch_render_bwd(const RenderParams& p, const CUtensorMap* rec_map, int grid_x, int grid_y, cudaStream_t s)
{
    dim3 grid(grid_x, grid_y, 1);
    if (p.dL_ddepth || p.dL_dacc) render_bwd_kernel<true><<<grid, 128, 0, s>>>(p, *rec_map);
    else render_bwd_kernel<false><<<grid, 128, 0, s>>>(p, *rec_map);
}

bool make_record_tensor_map(CUtensorMap* out, const SplatRec* rec, int P)
{
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn enc = nullptr;      // resolved once through the runtime: the library does not link libcuda
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<EncodeFn>(fn);
        else
            cudaGetLastError();
    }
    memset(out, 0, sizeof(*out));
    if (!enc || P <= 0) return false;
    const cuuint64_t dims[2] = {16, (cuuint64_t)P};
    const cuuint64_t strides[1] = {sizeof(SplatRec)};
    const cuuint32_t box[2] = {16, 1}, estr[2] = {1, 1};
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<SplatRec*>(rec), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
