// Per-Gaussian stages of the rasterizer for sm_100a: forward preprocess (cull, 3D->2D EWA
// projection with mip filter, SH->RGB, tile-rectangle, depth key), the fused backward of those
// steps, and the frustum test.
//
// Behavioural spec (reference, read for behaviour only):
//   forward : cuda_rasterizer/forward.cu:20-71 (SH), :74-124 (cov2D + mip coef), :128-162 (cov3D),
//             :165-269 (preprocessCUDA); auxiliary.h:41-56,68-87,267-294
//   backward: cuda_rasterizer/backward.cu:20-139 (SH), :144-300 (computeCov2DCUDA), :304-367
//             (cov3D), :372-423 (preprocessCUDA) including the deviations listed in SURVEY.md A.3
//             (Q1: mip-coef gradient dropped, Q2: cov2D->mean term overwritten, Q6: no quaternion
//             normalisation).
//
// Provenance note: the kernel structure here is this repository's own (fused K10 + K11, shared-memory staged coalesced
// I/O, zero-fill folded in, exact alpha threshold), but the closed-form derivative EXPRESSIONS of the backward
// (bwd_conic_to_cov3d, bwd_mean2d_to_mean3d, bwd_cov3d_to_scale_rot, bwd_dir_to_mean) and the real-SH basis / derivative
// tables restate backward.cu:225-251, :403-410, :342-366, auxiliary.h:235-245 and forward.cu:30-59 / backward.cu:47-123 term
// for term with renamed identifiers: the order of operations is part of the 1e-3 gradient-parity contract, the SH basis
// is fixed mathematics, and a derivative has one closed form.  They live in ONE set of __device__ functions used by both
// backward kernels.
#include "common.cuh"
#include <math.h>
#include <stdio.h>

namespace {

__device__ __constant__ float kC0 = 0.28209479177387814f;
__device__ __constant__ float kC1 = 0.4886025119029199f;
__device__ __constant__ float kC2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                        -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float kC3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                        0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                        -0.5900435899266435f};

// Sigma = (S R)^T (S R) from an UN-normalised quaternion (forward.cu:137), upper triangle.
// Which sums are fused is pinned to the reference build.
__device__ __forceinline__ void cov3d_from_scale_rot(float sx, float sy, float sz, float mod,
                                                     float r, float x, float y, float z, float* c)
{
    sx = fm(mod, sx); sy = fm(mod, sy); sz = fm(mod, sz);
    const float xz = fm(x, z), rx = fm(r, x), rz = fm(r, z), yy = fm(y, y), zz = fm(z, z);
    // rotation matrix entries, R[c][r] in column-major (glm) naming
    const float t_yyzz = fa(yy, zz);                 // y*y + z*z   (two rounded products, plain add)
    const float t_xxzz = ff(x, x, zz);               // x*x + z*z
    const float t_xxyy = ff(x, x, yy);               // x*x + y*y
    const float R00 = fa(1.f, -fa(t_yyzz, t_yyzz));
    const float R11 = fa(1.f, -fa(t_xxzz, t_xxzz));
    const float R22 = fa(1.f, -fa(t_xxyy, t_xxyy));
    float v;
    v = ff(x, y, -rz); const float R01 = fa(v, v);   // 2(xy - rz)
    v = ff(r, y, xz);  const float R02 = fa(v, v);   // 2(xz + ry)
    v = ff(x, y, rz);  const float R10 = fa(v, v);   // 2(xy + rz)
    v = ff(y, z, -rx); const float R12 = fa(v, v);   // 2(yz - rx)
    v = ff(-r, y, xz); const float R20 = fa(v, v);   // 2(xz - ry)
    v = ff(y, z, rx);  const float R21 = fa(v, v);   // 2(yz + rx)
    // M = S * R : M[c][r] = s_r * R[c][r]
    const float M00 = fm(sx, R00), M01 = fm(sy, R01), M02 = fm(sz, R02);
    const float M10 = fm(sx, R10), M11 = fm(sy, R11), M12 = fm(sz, R12);
    const float M20 = fm(sx, R20), M21 = fm(sy, R21), M22 = fm(sz, R22);
    // Sigma[c][r] = sum_k M[r][k] * M[c][k]
    c[0] = sum3(M00, M00, M01, M01, M02, M02);
    c[1] = sum3(M00, M10, M01, M11, M02, M12);
    c[2] = sum3(M00, M20, M01, M21, M02, M22);
    c[3] = sum3(M10, M10, M11, M11, M12, M12);
    c[4] = sum3(M10, M20, M11, M21, M12, M22);
    c[5] = sum3(M20, M20, M21, M21, M22, M22);
}

// Upper 2x2 of J W Sigma W^T J^T (EWA), before the mip filter.  Also returns the two rows of T.
struct Cov2D {
    float a, b, c;          // cov[0][0], cov[0][1], cov[1][1]
    float T0[3], T1[3];     // T[0][r], T[1][r]
    float tx, ty, tz;       // clamped view-space mean
    float txtz, tytz;
};

__device__ __forceinline__ Cov2D cov2d_project(float mx, float my, float mz, const float* __restrict__ V,
                                               float focal_x, float focal_y, float tan_fovx, float tan_fovy,
                                               const float* cov3D)
{
    Cov2D o;
    float tx = xform_row(V, 0, mx, my, mz);
    float ty = xform_row(V, 1, mx, my, mz);
    const float tz = xform_row(V, 2, mx, my, mz);
    const float limx = fm(1.3f, tan_fovx), limy = fm(1.3f, tan_fovy);
    o.txtz = __fdiv_rn(tx, tz);
    o.tytz = __fdiv_rn(ty, tz);
    tx = fm(fminf(limx, fmaxf(-limx, o.txtz)), tz);
    ty = fm(fminf(limy, fmaxf(-limy, o.tytz)), tz);
    o.tx = tx; o.ty = ty; o.tz = tz;
    const float tz2 = fm(tz, tz);
    const float J00 = __fdiv_rn(focal_x, tz);
    const float J02 = __fdiv_rn(-fm(focal_x, tx), tz2);
    const float J11 = __fdiv_rn(focal_y, tz);
    const float J12 = __fdiv_rn(-fm(focal_y, ty), tz2);
    // T = W * J with W[k][r] = V[4r + k]:  T[0][r] = V[4r]*J00 + V[4r+2]*J02 ; T[1][r] = V[4r+1]*J11 + V[4r+2]*J12
#pragma unroll
    for (int r = 0; r < 3; r++) {
        o.T0[r] = ff(J02, V[4 * r + 2], fm(V[4 * r + 0], J00));
        o.T1[r] = ff(J12, V[4 * r + 2], fm(V[4 * r + 1], J11));
    }
    const float c0 = cov3D[0], c1 = cov3D[1], c2 = cov3D[2], c3 = cov3D[3], c4 = cov3D[4], c5 = cov3D[5];
    // X[k][r] = sum_j T[r][j] * Vrk[k][j]
    const float X00 = sum3(o.T0[0], c0, o.T0[1], c1, o.T0[2], c2);
    const float X10 = sum3(o.T0[0], c1, o.T0[1], c3, o.T0[2], c4);
    const float X20 = sum3(o.T0[0], c2, o.T0[1], c4, o.T0[2], c5);
    const float X01 = sum3(o.T1[0], c0, o.T1[1], c1, o.T1[2], c2);
    const float X11 = sum3(o.T1[0], c1, o.T1[1], c3, o.T1[2], c4);
    const float X21 = sum3(o.T1[0], c2, o.T1[1], c4, o.T1[2], c5);
    // cov[c][r] = sum_k X[k][r] * T[c][k]
    o.a = sum3(o.T0[0], X00, o.T0[1], X10, o.T0[2], X20);
    o.b = sum3(o.T0[0], X01, o.T0[1], X11, o.T0[2], X21);
    o.c = sum3(o.T1[0], X01, o.T1[1], X11, o.T1[2], X21);
    return o;
}

// Real SH basis of the view direction up to degree D (forward.cu:20-71).  The expressions are NOT pinned: the
// forward image is bit-identical to the reference's because nvcc contracts them like the reference build does,
// and that choice depends on the surrounding code (inlined into the two instantiations of the kernel the
// degree >= 2 terms came out 1 ulp apart).  EX_PRE_FWD_TEMPLATE = 1 (default) compiles the function once
// (__noinline__) so that both instantiations share one contraction - verified bit-identical to the reference
// on every golden / live case; = 0 keeps a single kernel with a run-time test of p.seg.enabled (0.136 vs 0.127 ms).
#ifndef EX_PRE_FWD_TEMPLATE
#define EX_PRE_FWD_TEMPLATE 1
#endif
#if EX_PRE_FWD_TEMPLATE
__device__ __noinline__
#else
__device__ __forceinline__
#endif
int sh_basis_fwd(int D, float dx, float dy, float dz, float* b)
{
    int nb = 1;
    b[0] = kC0;
    if (D > 0) {
        b[1] = -kC1 * dy; b[2] = kC1 * dz; b[3] = -kC1 * dx; nb = 4;
        if (D > 1) {
            const float xx = dx * dx, yy = dy * dy, zz = dz * dz, xy = dx * dy, yz = dy * dz, xz = dx * dz;
            b[4] = kC2[0] * xy; b[5] = kC2[1] * yz; b[6] = kC2[2] * (2.0f * zz - xx - yy);
            b[7] = kC2[3] * xz; b[8] = kC2[4] * (xx - yy); nb = 9;
            if (D > 2) {
                b[9] = kC3[0] * dy * (3.0f * xx - yy);
                b[10] = kC3[1] * xy * dz;
                b[11] = kC3[2] * dy * (4.0f * zz - xx - yy);
                b[12] = kC3[3] * dz * (2.0f * zz - 3.0f * xx - 3.0f * yy);
                b[13] = kC3[4] * dx * (4.0f * zz - xx - yy);
                b[14] = kC3[5] * dz * (xx - yy);
                b[15] = kC3[6] * dx * (xx - 3.0f * yy);
                nb = 16;
            }
        }
    }
    return nb;
}

// The compositing loop blends a (pixel, splat) pair iff alpha = min(0.99, opac * expf(power)) >= 1/255
// (forward.cu:384-387).  alpha is a non-decreasing function of power, so there is ONE float thr with
//     alpha >= 1/255  <=>  power >= thr.
// It is found here, once per visible Gaussian, by bisection over the float bit patterns around p0 = log(1 / (255 opac))
// (window: +-1e-6 absolute and relative - logf and expf are good to a few 1e-7), evaluating the loop's own expression
// (libdevice expf, the same rounding of the product): 3 - 12 evaluations, more only for opacities within 1e-3 of
// 1/255.  So the forward's skip test, the backward's membership test and the reference's alpha test agree on EVERY
// pair, and the backward needs no exp() to know whether a pair contributed.  If the window does not bracket the
// crossing or the two floats next to it contradict monotonicity (never observed), the conservative threshold
// p0 - 1e-3 is stored instead and the frame's counter of such Gaussians is raised (the forward still applies the
// alpha test itself, so images are unaffected).
//   opac < 1/255: nothing passes (+inf);  NaN: every comparison with thr is false, the pair is kept as in the reference.
__device__ __forceinline__ float alpha_threshold(float opac, uint32_t* inexact_counter)
{
    if (opac != opac) return opac;
    if (opac < 1.0f / 255.0f) return __int_as_float(0x7f800000);
    const float p0 = fminf(logf(1.0f / (255.0f * opac)), -0.0f);
    if (!(p0 > -3.0e38f)) return p0;                                   // opac = +inf: alpha is always 0.99
    // work on magnitudes m = bits(-power): m + 1 is one ulp more negative (lower alpha)
    auto passes = [&](uint32_t m) { return !(fminf(0.99f, fm(opac, expf(-__uint_as_float(m)))) < 1.0f / 255.0f); };
    const float mag = -p0, d = fmaf(mag, 1e-6f, 1e-6f);
    uint32_t lo = __float_as_uint(fmaxf(mag - d, 0.0f)), hi = __float_as_uint(mag + d);
    bool ok = passes(lo) && !passes(hi);
    if (ok) {
        while (hi - lo > 1u) {                                         // passes(lo) && !passes(hi)
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (passes(mid)) lo = mid; else hi = mid;
        }
        ok = (lo == 0u || passes(lo - 1u)) && !passes(hi + 1u);
    }
    if (!ok) {
        atomicAdd(inexact_counter, 1u);
        return p0 - 1e-3f;
    }
    return -__uint_as_float(lo);
}

// SEG: the SH coefficients arrive as the model's four tensors (EX4DGS_FLAG_SH_SEGMENTED).  A template
// parameter, not a run-time test of SEG: with the test inside, the contiguous-SH instantiation
// pays for address selects in its inner loops (measured 0.122 -> 0.130 ms forward, 0.237 -> 0.277 ms backward).
template <bool SEG_T>
__global__ void __launch_bounds__(256) preprocess_fwd_kernel(const __grid_constant__ PreprocessParams p)
{
    const bool SEG = EX_PRE_FWD_TEMPLATE ? SEG_T : (p.seg.enabled != 0);
    __shared__ float s_cam[36];   // view[16] | proj[16] | campos[3]
    if (threadIdx.x < 16) s_cam[threadIdx.x] = __ldg(p.view + threadIdx.x);
    else if (threadIdx.x < 32) s_cam[threadIdx.x] = __ldg(p.proj + threadIdx.x - 16);
    else if (threadIdx.x < 35) s_cam[threadIdx.x] = __ldg(p.cam + threadIdx.x - 32);
    __syncthreads();
    const float* view = s_cam;
    const float* proj = s_cam + 16;
    const float* cam = s_cam + 32;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.P) return;

    int radius_out = 0;
    uint32_t tiles = 0, key = EX_INVISIBLE_KEY;
    // Every per-Gaussian input except the SH row is requested up front, before the frustum test decides whether it is
    // needed: the kernel is latency-bound (ncu: long-scoreboard stalls, 50 % issue), and this turns four dependent DRAM
    // round trips (mean -> scale / rotation -> SH -> opacity / dir3D) into two.  The extra bytes are few: culled and
    // visible Gaussians share their 32-byte sectors anyway (measured 360 MB read against 279 MB algorithmic before).
    const float mx = __ldg(p.means3D + 3 * idx), my = __ldg(p.means3D + 3 * idx + 1), mz = __ldg(p.means3D + 3 * idx + 2);
    float4 q_in = make_float4(0.f, 0.f, 0.f, 0.f);
    float sx_in = 0.f, sy_in = 0.f, sz_in = 0.f;
    if (p.cov3D_precomp == nullptr) {
        q_in = __ldg(reinterpret_cast<const float4*>(p.rotations) + idx);
        sx_in = __ldg(p.scales + 3 * idx); sy_in = __ldg(p.scales + 3 * idx + 1); sz_in = __ldg(p.scales + 3 * idx + 2);
    }
    const float opac_in = __ldg(p.opacities + idx);
    const float dir_x = __ldg(p.dir3D + 3 * idx), dir_y = __ldg(p.dir3D + 3 * idx + 1), dir_z = __ldg(p.dir3D + 3 * idx + 2);

    do {
        // ---- frustum test (auxiliary.h:267-294)
        const float hx = xform_row(proj, 0, mx, my, mz);
        const float hy = xform_row(proj, 1, mx, my, mz);
        const float hw = xform_row(proj, 3, mx, my, mz);
        const float pw = __fdiv_rn(1.0f, fa(hw, 0.0000001f));
        const float ndc_x = fm(hx, pw), ndc_y = fm(hy, pw);
        const float depth = xform_row(view, 2, mx, my, mz);
        if ((depth <= p.min_depth) || (depth > p.max_depth) ||
            ((double)ndc_x < -1.3 || (double)ndc_x > 1.3 || (double)ndc_y < -1.3 || (double)ndc_y > 1.3)) {
            if (p.prefiltered) {
                printf("Point is filtered although prefiltered is set. This shouldn't happen!");
                __trap();
            }
            break;
        }
        // ---- 3D covariance
        float cov3D[6];
        if (p.cov3D_precomp != nullptr) {
#pragma unroll
            for (int i = 0; i < 6; i++) cov3D[i] = __ldg(p.cov3D_precomp + 6 * idx + i);
        } else {
            cov3d_from_scale_rot(sx_in, sy_in, sz_in, p.scale_modifier, q_in.x, q_in.y, q_in.z, q_in.w, cov3D);
        }
        // ---- EWA projection + mip filter (forward.cu:74-124)
        const Cov2D cv = cov2d_project(mx, my, mz, view, p.focal_x, p.focal_y, p.tan_fovx, p.tan_fovy, cov3D);
        const float bb = fm(cv.b, cv.b);
        const float det0f = ff(cv.a, cv.c, -bb);
        const float ak = fa(cv.a, p.kernel_size), ck = fa(cv.c, p.kernel_size);
        const float det = ff(ak, ck, -bb);             // == det_1 before clamping == det of filtered cov
        const float det_0 = (float)fmax(1e-6, (double)det0f);
        const float det_1 = (float)fmax(1e-6, (double)det);
        float coef = (float)sqrt((double)det_0 / ((double)det_1 + 1e-6) + 1e-6);
        if ((double)det_0 <= 1e-6 || (double)det_1 <= 1e-6) coef = 0.0f;
        if (det == 0.0f) break;
        const float det_inv = __fdiv_rn(1.f, det);
        const float conA = fm(ck, det_inv), conB = fm(-cv.b, det_inv), conC = fm(ak, det_inv);
        // ---- extent (forward.cu:242-250)
        const float mid = fm(0.5f, fa(ak, ck));
        const float disc = __fsqrt_rn(fmaxf(0.1f, ff(mid, mid, -det)));
        const float lam = fmaxf(fa(mid, disc), fa(mid, -disc));
        const float radf = ceilf(fm(3.f, __fsqrt_rn(lam)));
        const int radius = (int)radf;
        // ndc2Pix in double: ((v + 1.0) * S - 1.0) * 0.5   (auxiliary.h:41-44)
        const float px = (float)((((double)ndc_x + 1.0) * (double)p.W - 1.0) * 0.5);
        const float py = (float)((((double)ndc_y + 1.0) * (double)p.H - 1.0) * 0.5);
        int x0, y0, x1, y1;
        tile_rect(px, py, radius, p.grid_x, p.grid_y, x0, y0, x1, y1);
        if ((uint32_t)(x1 - x0) * (uint32_t)(y1 - y0) == 0) break;

        // ---- colour
        float rgb[3];
        uint8_t clamp_bits = 0;
        if (p.colors_precomp == nullptr) {
            float dx = fa(mx, -cam[0]), dy = fa(my, -cam[1]), dz = fa(mz, -cam[2]);
            const float len = __fsqrt_rn(sum3(dx, dx, dy, dy, dz, dz));
            dx = __fdiv_rn(dx, len); dy = __fdiv_rn(dy, len); dz = __fdiv_rn(dz, len);
            const float* sh = SEG ? nullptr : p.shs + (size_t)idx * p.M * 3;
            float b[16];
            const int nb = sh_basis_fwd(p.D, dx, dy, dz, b);
            float acc[3] = {0.f, 0.f, 0.f};
            if (SEG) {
                // the model's own tensors, read in place: 3 floats of dc + 3*(nb-1) floats of the 180-byte rest row
                const int sgi = idx >= p.seg.n_static;
                const size_t li = (size_t)(idx - (sgi ? p.seg.n_static : 0));
                const float* dc = p.seg.dc[sgi] + li * 3;
                const float* rs = p.seg.rest[sgi] + li * 45;
                acc[0] = fmaf(b[0], __ldg(dc), acc[0]);
                acc[1] = fmaf(b[0], __ldg(dc + 1), acc[1]);
                acc[2] = fmaf(b[0], __ldg(dc + 2), acc[2]);
#pragma unroll
                for (int k = 1; k < 16; k++) {
                    if (k < nb) {
                        acc[0] = fmaf(b[k], __ldg(rs + 3 * (k - 1)), acc[0]);
                        acc[1] = fmaf(b[k], __ldg(rs + 3 * (k - 1) + 1), acc[1]);
                        acc[2] = fmaf(b[k], __ldg(rs + 3 * (k - 1) + 2), acc[2]);
                    }
                }
            } else if (p.M == 16) {
                // 192-byte row, 16-byte aligned: twelve 128-bit loads
                const float4* s4 = reinterpret_cast<const float4*>(sh);
                float v[48];
#pragma unroll
                for (int i = 0; i < 12; i++) {
                    const float4 t = __ldg(s4 + i);
                    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
                }
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    if (k < nb) {
                        acc[0] = fmaf(b[k], v[3 * k], acc[0]);
                        acc[1] = fmaf(b[k], v[3 * k + 1], acc[1]);
                        acc[2] = fmaf(b[k], v[3 * k + 2], acc[2]);
                    }
                }
            } else {
                for (int k = 0; k < nb; k++) {
                    acc[0] = fmaf(b[k], __ldg(sh + 3 * k), acc[0]);
                    acc[1] = fmaf(b[k], __ldg(sh + 3 * k + 1), acc[1]);
                    acc[2] = fmaf(b[k], __ldg(sh + 3 * k + 2), acc[2]);
                }
            }
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                const float v = acc[ch] + 0.5f;
                if (v < 0.f) clamp_bits |= (1u << ch);
                rgb[ch] = fmaxf(v, 0.0f);
            }
        } else {
#pragma unroll
            for (int ch = 0; ch < 3; ch++) rgb[ch] = __ldg(p.colors_precomp + 3 * idx + ch);
        }
        p.clamped[idx] = clamp_bits;

        const float opac = fm(opac_in, coef);
        // skip threshold of the compositing loops: power < thr  <=>  min(0.99, opac*expf(power)) < 1/255, exactly
        const float thr = alpha_threshold(opac, p.inexact_thr);

        // with EX4DGS_FLAG_TILE_CULL the rectangle is first cut down to the bounding box of the alpha >= 1/255
        // ellipse (tight_rect); the duplicate kernel then keys the instances its exact test rejects to the
        // dump tile, so that no second counting pass is needed before the scan.  A Gaussian whose box meets no
        // pixel centre keeps its radius (it is "visible" to the caller, as in the reference) but enters no list.
        if (p.flags & 1u)     // EX4DGS_FLAG_TILE_CULL
            tight_rect(px, py, conA, conB, conC, thr, __ldg(p.pad_ptr), x0, y0, x1, y1);
        const uint32_t count = (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0);

        SplatRec rc;
        rc.a = make_float4(px, py, depth, thr);
        rc.b = make_float4(conA, conB, conC, opac);
        rc.c = make_float4(rgb[0], rgb[1], rgb[2], __int_as_float(idx));
        rc.d = make_float4(dir_x, dir_y, dir_z, 0.f);
        // anything but +-0 (NaN/inf included) makes the compositing kernel carry the flow accumulators
        if (((__float_as_uint(rc.d.x) | __float_as_uint(rc.d.y) | __float_as_uint(rc.d.z)) & 0x7fffffffu) != 0u) *p.flow_flag = 1u;
        p.rec[idx] = rc;

        radius_out = radius;
        tiles = count;
        key = __float_as_uint(depth);
    } while (false);

    p.radii[idx] = radius_out;
    p.tiles_touched[idx] = tiles;
    p.key_in[idx] = tiles ? key : EX_INVISIBLE_KEY;
}

__global__ void __launch_bounds__(256) mark_visible_kernel(int P, const float* __restrict__ means,
                                                           const float* __restrict__ view, const float* __restrict__ proj,
                                                           float min_depth, float max_depth, uint8_t* __restrict__ present)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float mx = __ldg(means + 3 * idx), my = __ldg(means + 3 * idx + 1), mz = __ldg(means + 3 * idx + 2);
    const float hx = xform_row(proj, 0, mx, my, mz);
    const float hy = xform_row(proj, 1, mx, my, mz);
    const float hw = xform_row(proj, 3, mx, my, mz);
    const float pw = __fdiv_rn(1.0f, fa(hw, 0.0000001f));
    const float ndc_x = fm(hx, pw), ndc_y = fm(hy, pw);
    const float depth = xform_row(view, 2, mx, my, mz);
    const bool out = (depth <= min_depth) || (depth > max_depth) ||
                     ((double)ndc_x < -1.3 || (double)ndc_x > 1.3 || (double)ndc_y < -1.3 || (double)ndc_y > 1.3);
    present[idx] = out ? 0 : 1;
}

__device__ __forceinline__ void sh_basis_and_grad(int D, float x, float y, float z, int k,
                                                  float& b, float& bx, float& by, float& bz)
{
    // value and d/d(x,y,z) of real SH basis k (forward.cu:30-59, backward.cu:47-123), k < (D+1)^2
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    bx = by = bz = 0.f;
    switch (k) {
    case 0: b = kC0; break;
    case 1: b = -kC1 * y; by = -kC1; break;
    case 2: b = kC1 * z; bz = kC1; break;
    case 3: b = -kC1 * x; bx = -kC1; break;
    case 4: b = kC2[0] * xy; bx = kC2[0] * y; by = kC2[0] * x; break;
    case 5: b = kC2[1] * yz; by = kC2[1] * z; bz = kC2[1] * y; break;
    case 6: b = kC2[2] * (2.f * zz - xx - yy); bx = kC2[2] * 2.f * -x; by = kC2[2] * 2.f * -y; bz = kC2[2] * 2.f * 2.f * z; break;
    case 7: b = kC2[3] * xz; bx = kC2[3] * z; bz = kC2[3] * x; break;
    case 8: b = kC2[4] * (xx - yy); bx = kC2[4] * 2.f * x; by = kC2[4] * 2.f * -y; break;
    case 9: b = kC3[0] * y * (3.f * xx - yy); bx = kC3[0] * 3.f * 2.f * xy; by = kC3[0] * 3.f * (xx - yy); break;
    case 10: b = kC3[1] * xy * z; bx = kC3[1] * yz; by = kC3[1] * xz; bz = kC3[1] * xy; break;
    case 11: b = kC3[2] * y * (4.f * zz - xx - yy); bx = kC3[2] * -2.f * xy; by = kC3[2] * (-3.f * yy + 4.f * zz - xx); bz = kC3[2] * 4.f * 2.f * yz; break;
    case 12: b = kC3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy); bx = kC3[3] * -3.f * 2.f * xz; by = kC3[3] * -3.f * 2.f * yz; bz = kC3[3] * 3.f * (2.f * zz - xx - yy); break;
    case 13: b = kC3[4] * x * (4.f * zz - xx - yy); bx = kC3[4] * (-3.f * xx + 4.f * zz - yy); by = kC3[4] * -2.f * xy; bz = kC3[4] * 4.f * 2.f * xz; break;
    case 14: b = kC3[5] * z * (xx - yy); bx = kC3[5] * 2.f * xz; by = kC3[5] * -2.f * yz; bz = kC3[5] * (xx - yy); break;
    default: b = kC3[6] * x * (xx - 3.f * yy); bx = kC3[6] * 3.f * (xx - yy); by = kC3[6] * -3.f * 2.f * xy; break;
    }
}

// ---- shared pieces of the two backward kernels (closed-form derivatives; the expressions follow backward.cu:225-251,
// :403-410, :342-366 and auxiliary.h:235-245 term for term - a derivative has one form - including the reference's
// deviations from the true gradient) ---------------------------------------------------------------

// conic gradient -> cov2D -> cov3D (backward.cu:144-257).  The cov2D -> mean term (backward.cu:259-299) is overwritten by
// the assignment at backward.cu:414 in the reference, so it is not computed (A.3-Q2); the mip-coefficient gradient is
// dropped as in the reference (A.3-Q1).
__device__ __forceinline__ void bwd_conic_to_cov3d(const PreprocessBwdParams& p, const float* view, float mx, float my, float mz,
                                                   const float* cov3D, float dcx, float dcy, float dcz, float* dcov)
{
    const Cov2D cv = cov2d_project(mx, my, mz, view, p.focal_x, p.focal_y, p.tan_fovx, p.tan_fovy, cov3D);
    const float a = cv.a + p.kernel_size, b = cv.b, c = cv.c + p.kernel_size;
    const float denom = a * c - b * b;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    if (denom2inv != 0) {
        const float dL_da = denom2inv * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz);
        const float dL_dc = denom2inv * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx);
        const float dL_db = denom2inv * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
        const float* T0 = cv.T0; const float* T1 = cv.T1;
        dcov[0] = (T0[0] * T0[0] * dL_da + T0[0] * T1[0] * dL_db + T1[0] * T1[0] * dL_dc);
        dcov[3] = (T0[1] * T0[1] * dL_da + T0[1] * T1[1] * dL_db + T1[1] * T1[1] * dL_dc);
        dcov[5] = (T0[2] * T0[2] * dL_da + T0[2] * T1[2] * dL_db + T1[2] * T1[2] * dL_dc);
        dcov[1] = 2 * T0[0] * T0[1] * dL_da + (T0[0] * T1[1] + T0[1] * T1[0]) * dL_db + 2 * T1[0] * T1[1] * dL_dc;
        dcov[2] = 2 * T0[0] * T0[2] * dL_da + (T0[0] * T1[2] + T0[2] * T1[0]) * dL_db + 2 * T1[0] * T1[2] * dL_dc;
        dcov[4] = 2 * T0[2] * T0[1] * dL_da + (T0[1] * T1[2] + T0[2] * T1[1]) * dL_db + 2 * T1[1] * T1[2] * dL_dc;
    }
}

// mean2D -> mean3D through the projection (backward.cu:396-414; an assignment, not an accumulation)
__device__ __forceinline__ void bwd_mean2d_to_mean3d(const float* pr, float mx, float my, float mz, float gx, float gy, float gz,
                                                     float* dmean)
{
    const float hw = pr[3] * mx + pr[7] * my + pr[11] * mz + pr[15];
    const float m_w = 1.0f / (hw + 0.0000001f);
    const float mul1 = (pr[0] * mx + pr[4] * my + pr[8] * mz + pr[12]) * m_w * m_w;
    const float mul2 = (pr[1] * mx + pr[5] * my + pr[9] * mz + pr[13]) * m_w * m_w;
    const float mul3 = (pr[2] * mx + pr[6] * my + pr[10] * mz + pr[14]) * m_w * m_w;
    dmean[0] = (pr[0] * m_w - pr[3] * mul1) * gx + (pr[1] * m_w - pr[3] * mul2) * gy + (pr[2] * m_w - pr[3] * mul3) * gz;
    dmean[1] = (pr[4] * m_w - pr[7] * mul1) * gx + (pr[5] * m_w - pr[7] * mul2) * gy + (pr[6] * m_w - pr[7] * mul3) * gz;
    dmean[2] = (pr[8] * m_w - pr[11] * mul1) * gx + (pr[9] * m_w - pr[11] * mul2) * gy + (pr[10] * m_w - pr[11] * mul3) * gz;
}

// gradient of the SH view direction through dir = o / |o| (auxiliary.h:235-245), added to dmean
__device__ __forceinline__ void bwd_dir_to_mean(float ox, float oy, float oz, float ddx, float ddy, float ddz, float* dmean)
{
    const float sum2 = ox * ox + oy * oy + oz * oz;
    const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    dmean[0] += ((+sum2 - ox * ox) * ddx - oy * ox * ddy - oz * ox * ddz) * invsum32;
    dmean[1] += (-ox * oy * ddx + (sum2 - oy * oy) * ddy - oz * oy * ddz) * invsum32;
    dmean[2] += (-ox * oz * ddx - oy * oz * ddy + (sum2 - oz * oz) * ddz) * invsum32;
}

// cov3D -> scale / rotation (backward.cu:304-367), no quaternion-normalisation Jacobian (A.3-Q6)
__device__ __forceinline__ void bwd_cov3d_to_scale_rot(float4 q, float sx, float sy, float sz, float mod, const float* dcov,
                                                       float* dscale, float4& drot)
{
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    // R[c][r] column-major as in the forward
    const float R[3][3] = {
        {1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
        {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
        {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
    const float s[3] = {mod * sx, mod * sy, mod * sz};
    // M[c][r] = s_r R[c][r];  dL_dSigma symmetric with halved off-diagonals
    const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                            {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                            {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
    // dL_dM = 2 M dL_dSigma : dL_dM[c][r] = 2 sum_k M[k][r] dS[c][k]
    float dM[3][3];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int rr = 0; rr < 3; rr++)
            dM[c][rr] = 2.0f * (s[rr] * R[0][rr] * dS[c][0] + s[rr] * R[1][rr] * dS[c][1] + s[rr] * R[2][rr] * dS[c][2]);
    // dL_dscale_a = sum_b R[b][a] dM[b][a]
#pragma unroll
    for (int a = 0; a < 3; a++) dscale[a] = R[0][a] * dM[0][a] + R[1][a] * dM[1][a] + R[2][a] * dM[2][a];
    float Mt[3][3];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b2 = 0; b2 < 3; b2++) Mt[a][b2] = dM[b2][a] * s[a];
    drot.x = 2 * z * (Mt[0][1] - Mt[1][0]) + 2 * y * (Mt[2][0] - Mt[0][2]) + 2 * x * (Mt[1][2] - Mt[2][1]);
    drot.y = 2 * y * (Mt[1][0] + Mt[0][1]) + 2 * z * (Mt[2][0] + Mt[0][2]) + 2 * r * (Mt[1][2] - Mt[2][1]) - 4 * x * (Mt[2][2] + Mt[1][1]);
    drot.z = 2 * x * (Mt[1][0] + Mt[0][1]) + 2 * r * (Mt[2][0] - Mt[0][2]) + 2 * z * (Mt[1][2] + Mt[2][1]) - 4 * y * (Mt[2][2] + Mt[0][0]);
    drot.w = 2 * r * (Mt[0][1] - Mt[1][0]) + 2 * x * (Mt[2][0] + Mt[0][2]) + 2 * y * (Mt[1][2] + Mt[2][1]) - 4 * z * (Mt[1][1] + Mt[0][0]);
}

// ---------------------------------------------------------------------------------------------
// Fused backward of the per-Gaussian steps (backward.cu computeCov2DCUDA + preprocessCUDA in one
// pass, zero-fill of invisible Gaussians folded in).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) preprocess_bwd_kernel(const __grid_constant__ PreprocessBwdParams p)
{
    __shared__ float s_cam[36];   // view[16] | proj[16] | campos[3]
    if (threadIdx.x < 16) s_cam[threadIdx.x] = __ldg(p.view + threadIdx.x);
    else if (threadIdx.x < 32) s_cam[threadIdx.x] = __ldg(p.proj + threadIdx.x - 16);
    else if (threadIdx.x < 35) s_cam[threadIdx.x] = __ldg(p.cam + threadIdx.x - 32);
    __syncthreads();
    const float* view = s_cam;
    const float* cam = s_cam + 32;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.P) return;

    const GradAcc g = gacc_load(p.gacc, p.rec, idx, p.W, p.H);
    // pass-through gradients (zero for Gaussians the compositing loop never touched)
    p.dL_dmean2D[3 * idx + 0] = g.g0.x; p.dL_dmean2D[3 * idx + 1] = g.g0.y; p.dL_dmean2D[3 * idx + 2] = g.g0.z;
    p.dL_dopacity[idx] = g.g0.w;
    p.dL_dcolor[3 * idx + 0] = g.g2.x; p.dL_dcolor[3 * idx + 1] = g.g2.y; p.dL_dcolor[3 * idx + 2] = g.g2.z;
    p.dL_ddir[3 * idx + 0] = g.g3.x; p.dL_ddir[3 * idx + 1] = g.g3.y; p.dL_ddir[3 * idx + 2] = g.g3.z;

    float dmean[3] = {0.f, 0.f, 0.f};
    float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float dscale[3] = {0.f, 0.f, 0.f};
    float drot[4] = {0.f, 0.f, 0.f, 0.f};
    const bool visible = p.radii[idx] > 0;
    float4* dsh4 = (p.dL_dsh != nullptr && p.M == 16) ? reinterpret_cast<float4*>(p.dL_dsh + (size_t)idx * 48) : nullptr;

    if (visible) {
        const float mx = __ldg(p.means3D + 3 * idx), my = __ldg(p.means3D + 3 * idx + 1), mz = __ldg(p.means3D + 3 * idx + 2);
        float cov3D[6];
        float sx = 0, sy = 0, sz = 0;
        float4 q = make_float4(0, 0, 0, 0);
        if (p.cov3D_precomp != nullptr) {
#pragma unroll
            for (int i = 0; i < 6; i++) cov3D[i] = __ldg(p.cov3D_precomp + 6 * idx + i);
        } else {
            q = __ldg(reinterpret_cast<const float4*>(p.rotations) + idx);
            sx = __ldg(p.scales + 3 * idx); sy = __ldg(p.scales + 3 * idx + 1); sz = __ldg(p.scales + 3 * idx + 2);
            cov3d_from_scale_rot(sx, sy, sz, p.scale_modifier, q.x, q.y, q.z, q.w, cov3D);
        }
        bwd_conic_to_cov3d(p, view, mx, my, mz, cov3D, g.g1.x, g.g1.y, g.g1.z, dcov);
        bwd_mean2d_to_mean3d(s_cam + 16, mx, my, mz, g.g0.x, g.g0.y, g.g0.z, dmean);
        // ---- SH backward (backward.cu:20-139)
        if (p.shs != nullptr) {
            const float ox = mx - cam[0], oy = my - cam[1], oz = mz - cam[2];
            const float len = sqrtf(ox * ox + oy * oy + oz * oz);
            const float x = ox / len, y = oy / len, z = oz / len;
            const uint8_t cl = p.clamped[idx];
            float dRGB[3] = {(cl & 1) ? 0.f : g.g2.x, (cl & 2) ? 0.f : g.g2.y, (cl & 4) ? 0.f : g.g2.z};
            const float* sh = p.shs + (size_t)idx * p.M * 3;
            const int nb = (p.D + 1) * (p.D + 1);
            float b[16], dbx[16], dby[16], dbz[16];      // basis_k and d basis_k / d(x,y,z)
#pragma unroll
            for (int k = 0; k < 16; k++) {
                b[k] = dbx[k] = dby[k] = dbz[k] = 0.f;
                if (k < nb) sh_basis_and_grad(p.D, x, y, z, k, b[k], dbx[k], dby[k], dbz[k]);
            }
            float ddir[3] = {0.f, 0.f, 0.f};
            if (p.M == 16) {
                const float4* s4 = reinterpret_cast<const float4*>(sh);
                float v[48], o[48];
#pragma unroll
                for (int i = 0; i < 12; i++) {
                    const float4 t = __ldg(s4 + i);
                    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
                }
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const bool on = k < nb;
#pragma unroll
                    for (int ch = 0; ch < 3; ch++) {
                        o[3 * k + ch] = on ? b[k] * dRGB[ch] : 0.f;
                        if (on) {
                            const float s = v[3 * k + ch] * dRGB[ch];
                            ddir[0] += dbx[k] * s; ddir[1] += dby[k] * s; ddir[2] += dbz[k] * s;
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 12; i++) dsh4[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
            } else {
                float* dsh = p.dL_dsh + (size_t)idx * p.M * 3;
                for (int k = 0; k < p.M; k++) {
                    const bool on = k < nb;
                    for (int ch = 0; ch < 3; ch++) {
                        dsh[3 * k + ch] = on ? b[k] * dRGB[ch] : 0.f;
                        if (on) {
                            const float s = __ldg(sh + 3 * k + ch) * dRGB[ch];
                            ddir[0] += dbx[k] * s; ddir[1] += dby[k] * s; ddir[2] += dbz[k] * s;
                        }
                    }
                }
            }
            bwd_dir_to_mean(ox, oy, oz, ddir[0], ddir[1], ddir[2], dmean);
        }
        if (p.scales != nullptr) {
            float4 dr;
            bwd_cov3d_to_scale_rot(q, sx, sy, sz, p.scale_modifier, dcov, dscale, dr);
            drot[0] = dr.x; drot[1] = dr.y; drot[2] = dr.z; drot[3] = dr.w;
        }
    } else if (p.dL_dsh != nullptr) {
        if (dsh4 != nullptr) {
#pragma unroll
            for (int i = 0; i < 12; i++) dsh4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            float* dsh = p.dL_dsh + (size_t)idx * p.M * 3;
            for (int k = 0; k < p.M * 3; k++) dsh[k] = 0.f;
        }
    }

    p.dL_dmean3D[3 * idx + 0] = dmean[0]; p.dL_dmean3D[3 * idx + 1] = dmean[1]; p.dL_dmean3D[3 * idx + 2] = dmean[2];
    if (p.dL_dcov3D != nullptr) {
#pragma unroll
        for (int i = 0; i < 6; i++) p.dL_dcov3D[6 * idx + i] = dcov[i];
    }
    if (p.dL_dscale != nullptr) {
        p.dL_dscale[3 * idx + 0] = dscale[0]; p.dL_dscale[3 * idx + 1] = dscale[1]; p.dL_dscale[3 * idx + 2] = dscale[2];
    }
    if (p.dL_drot != nullptr)
        reinterpret_cast<float4*>(p.dL_drot)[idx] = make_float4(drot[0], drot[1], drot[2], drot[3]);
}

// ---------------------------------------------------------------------------------------------
// Same math, B200 memory layout: 128 Gaussians per CTA, every [P,3] / [P,16,3] tensor crosses
// HBM with fully coalesced 128-bit accesses through a shared-memory staging area, the 192-byte SH
// row is read and its gradient written in place there (no 48+48 register arrays -> 3x the occupancy).
// Contiguous SH ([P,16,3]): the block's rows are requested with 16-byte cp.async (LDGSTS: no register
// staging, all twelve requests of a thread in flight at once) into rows padded to 52 floats - 16-byte
// aligned, and 13 float4 per row is odd, so the per-thread LDS.128 / STS.128 of the compute phase are
// conflict-free - and everything that does not need the row (accumulator line, conic -> cov3D ->
// scale / rotation, mean2D -> mean3D) runs while they are in flight.
// Handles M == 16 (degree-3 layout) and M == 0 (precomputed colours); other M use the kernel above.
// ---------------------------------------------------------------------------------------------
constexpr int kBT = 128;        // threads = Gaussians per CTA
constexpr int kRow = 52;        // padded SH row (floats)

template <bool SEG>
__global__ void __launch_bounds__(kBT, 4) preprocess_bwd_staged_kernel(const __grid_constant__ PreprocessBwdParams p)
{
    extern __shared__ __align__(16) float smem[];
    float* s_cam = smem;                       // 36 (+4 pad)
    float* s_o3 = smem + 40;                   // 5 x [kBT*3]: mean2D, color, dir, mean3D, scale
    float* s_sh = s_o3 + 5 * kBT * 3;          // [kBT][kRow]   (only when M == 16)
    const int tid = threadIdx.x;
    if (tid < 16) s_cam[tid] = __ldg(p.view + tid);
    else if (tid < 32) s_cam[tid] = __ldg(p.proj + tid - 16);
    else if (tid < 35) s_cam[tid] = __ldg(p.cam + tid - 32);
    const int base = blockIdx.x * kBT;
    const int idx = base + tid;
    const int nvalid = min(kBT, p.P - base);
    const bool has_sh = (p.shs != nullptr) || SEG;    // implies M == 16 here

    // cooperative, coalesced load of the block's SH rows
    // Segmented SH: the block's dc values (3 floats per Gaussian) and rest rows (45 floats) are contiguous
    // inside each of the model's tensors, so they are staged FLAT, exactly as they lie in memory:
    // s_dc[3 * row + c], s_rest[45 * row + 3 * (k - 1) + c] - a pure 128-bit copy in and out, and both
    // per-thread strides (3, 45) are odd: the per-row accesses of the compute phase are conflict-free.
    float* s_dc = s_sh;
    float* s_rest = s_sh + kBT * 3;
    const int seg_ns = p.seg.n_static;
    const bool seg_one = SEG && ((base >= seg_ns) || (base + nvalid <= seg_ns));
    const int seg_i = base >= seg_ns;
    const size_t seg_li0 = (size_t)(base - (seg_i ? seg_ns : 0));
    if (SEG) {
        if (seg_one) {
            const float* dc = p.seg.dc[seg_i] + seg_li0 * 3;
            const float* rs = p.seg.rest[seg_i] + seg_li0 * 45;
            const int nd = nvalid * 3, nr = nvalid * 45;
            if (((reinterpret_cast<uintptr_t>(dc) | reinterpret_cast<uintptr_t>(rs)) & 15) == 0 && (nvalid & 3) == 0) {
                for (int f4 = tid; f4 < nd / 4; f4 += kBT) cp_async16(reinterpret_cast<float4*>(s_dc) + f4, reinterpret_cast<const float4*>(dc) + f4);
                for (int f4 = tid; f4 < nr / 4; f4 += kBT) cp_async16(reinterpret_cast<float4*>(s_rest) + f4, reinterpret_cast<const float4*>(rs) + f4);
            } else {
                for (int f = tid; f < nd; f += kBT) s_dc[f] = __ldg(dc + f);
                for (int f = tid; f < nr; f += kBT) s_rest[f] = __ldg(rs + f);
            }
        } else {      // the one block that straddles the static / dynamic boundary
            for (int f = tid; f < nvalid * 3; f += kBT) {
                const int row = f / 3, g = base + row;
                const int sgi = g >= seg_ns;
                s_dc[f] = __ldg(p.seg.dc[sgi] + (size_t)(g - (sgi ? seg_ns : 0)) * 3 + (f - row * 3));
            }
            for (int f = tid; f < nvalid * 45; f += kBT) {
                const int row = f / 45, g = base + row;
                const int sgi = g >= seg_ns;
                s_rest[f] = __ldg(p.seg.rest[sgi] + (size_t)(g - (sgi ? seg_ns : 0)) * 45 + (f - row * 45));
            }
        }
    } else if (has_sh) {
        const float4* src = reinterpret_cast<const float4*>(p.shs + (size_t)base * 48);
        const int n4 = nvalid * 12;
        for (int f = tid; f < n4; f += kBT) {
            const int row = f / 12, col = f % 12;
            cp_async16(s_sh + row * kRow + col * 4, src + f);
        }
    }
    cp_async_commit();
    __syncthreads();                          // camera constants
    const float* view = s_cam;
    const float* pr = s_cam + 16;
    const float* cam = s_cam + 32;

    // ---- phase A (the SH rows are still in flight): accumulator line, conic -> cov3D -> scale / rotation, mean2D -> mean3D
    bool vis = false;
    float dmean[3] = {0.f, 0.f, 0.f};
    float dscale[3] = {0.f, 0.f, 0.f};
    float ox = 0.f, oy = 0.f, oz = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f;
    if (idx < p.P) {
        const GradAcc g = gacc_load(p.gacc, p.rec, idx, p.W, p.H);
        s_o3[0 * kBT * 3 + 3 * tid + 0] = g.g0.x; s_o3[0 * kBT * 3 + 3 * tid + 1] = g.g0.y; s_o3[0 * kBT * 3 + 3 * tid + 2] = g.g0.z;
        s_o3[1 * kBT * 3 + 3 * tid + 0] = g.g2.x; s_o3[1 * kBT * 3 + 3 * tid + 1] = g.g2.y; s_o3[1 * kBT * 3 + 3 * tid + 2] = g.g2.z;
        s_o3[2 * kBT * 3 + 3 * tid + 0] = g.g3.x; s_o3[2 * kBT * 3 + 3 * tid + 1] = g.g3.y; s_o3[2 * kBT * 3 + 3 * tid + 2] = g.g3.z;
        p.dL_dopacity[idx] = g.g0.w;

        float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float4 drot = make_float4(0.f, 0.f, 0.f, 0.f);
        vis = p.radii[idx] > 0;
        if (vis) {
            const float mx = __ldg(p.means3D + 3 * idx), my = __ldg(p.means3D + 3 * idx + 1), mz = __ldg(p.means3D + 3 * idx + 2);
            float cov3D[6];
            float sx = 0, sy = 0, sz = 0;
            float4 q = make_float4(0, 0, 0, 0);
            if (p.cov3D_precomp != nullptr) {
#pragma unroll
                for (int i = 0; i < 6; i++) cov3D[i] = __ldg(p.cov3D_precomp + 6 * idx + i);
            } else {
                q = __ldg(reinterpret_cast<const float4*>(p.rotations) + idx);
                sx = __ldg(p.scales + 3 * idx); sy = __ldg(p.scales + 3 * idx + 1); sz = __ldg(p.scales + 3 * idx + 2);
                cov3d_from_scale_rot(sx, sy, sz, p.scale_modifier, q.x, q.y, q.z, q.w, cov3D);
            }
            bwd_conic_to_cov3d(p, view, mx, my, mz, cov3D, g.g1.x, g.g1.y, g.g1.z, dcov);
            bwd_mean2d_to_mean3d(pr, mx, my, mz, g.g0.x, g.g0.y, g.g0.z, dmean);
            if (p.scales != nullptr) bwd_cov3d_to_scale_rot(q, sx, sy, sz, p.scale_modifier, dcov, dscale, drot);
            if (has_sh) {
                ox = mx - cam[0]; oy = my - cam[1]; oz = mz - cam[2];
                const uint8_t cl = p.clamped[idx];
                d0 = (cl & 1) ? 0.f : g.g2.x; d1 = (cl & 2) ? 0.f : g.g2.y; d2 = (cl & 4) ? 0.f : g.g2.z;
            }
        }
        if (p.dL_dcov3D != nullptr) {
#pragma unroll
            for (int i = 0; i < 6; i++) p.dL_dcov3D[6 * idx + i] = dcov[i];
        }
        if (p.dL_drot != nullptr) reinterpret_cast<float4*>(p.dL_drot)[idx] = drot;
    }
    cp_async_wait_all();
    __syncthreads();                          // SH rows

    // ---- phase B: SH backward in place on the staged row (backward.cu:20-139)
    if (idx < p.P) {
        if (has_sh) {
            const int nb = (p.D + 1) * (p.D + 1);
            float ddx = 0.f, ddy = 0.f, ddz = 0.f;
            float x = 0.f, y = 0.f, z = 0.f;
            if (vis) {
                const float len = sqrtf(ox * ox + oy * oy + oz * oz);
                x = ox / len; y = oy / len; z = oz / len;
            }
            if (SEG) {
                // coefficient k of this thread's row lives at (k == 0 ? row0 : row)[3 * k + c]
                float* row = s_rest + 45 * tid - 3;
                float* row0 = s_dc + 3 * tid;
                if (vis) {
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        float* rk = (k == 0) ? row0 : row;
                        if (k < nb) {
                            float b, bx, by, bz;
                            sh_basis_and_grad(p.D, x, y, z, k, b, bx, by, bz);
                            const float s = rk[3 * k] * d0 + rk[3 * k + 1] * d1 + rk[3 * k + 2] * d2;
                            ddx += bx * s; ddy += by * s; ddz += bz * s;
                            rk[3 * k] = b * d0; rk[3 * k + 1] = b * d1; rk[3 * k + 2] = b * d2;
                        } else {
                            rk[3 * k] = 0.f; rk[3 * k + 1] = 0.f; rk[3 * k + 2] = 0.f;
                        }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 48; k++) ((k < 3) ? row0 : row)[k] = 0.f;
                }
            } else {
                // four coefficients (three 16-byte words) at a time
                float4* r4 = reinterpret_cast<float4*>(s_sh + tid * kRow);
                if (vis) {
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const float4 q0 = r4[3 * c], q1 = r4[3 * c + 1], q2 = r4[3 * c + 2];
                        const float v[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
                        float o[12];
#pragma unroll
                        for (int kk = 0; kk < 4; kk++) {
                            const int k = 4 * c + kk;
                            if (k < nb) {
                                float b, bx, by, bz;
                                sh_basis_and_grad(p.D, x, y, z, k, b, bx, by, bz);
                                const float s = v[3 * kk] * d0 + v[3 * kk + 1] * d1 + v[3 * kk + 2] * d2;
                                ddx += bx * s; ddy += by * s; ddz += bz * s;
                                o[3 * kk] = b * d0; o[3 * kk + 1] = b * d1; o[3 * kk + 2] = b * d2;
                            } else {
                                o[3 * kk] = 0.f; o[3 * kk + 1] = 0.f; o[3 * kk + 2] = 0.f;
                            }
                        }
                        r4[3 * c] = make_float4(o[0], o[1], o[2], o[3]);
                        r4[3 * c + 1] = make_float4(o[4], o[5], o[6], o[7]);
                        r4[3 * c + 2] = make_float4(o[8], o[9], o[10], o[11]);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 12; c++) r4[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            if (vis) bwd_dir_to_mean(ox, oy, oz, ddx, ddy, ddz, dmean);
        }
        s_o3[3 * kBT * 3 + 3 * tid + 0] = dmean[0]; s_o3[3 * kBT * 3 + 3 * tid + 1] = dmean[1]; s_o3[3 * kBT * 3 + 3 * tid + 2] = dmean[2];
        s_o3[4 * kBT * 3 + 3 * tid + 0] = dscale[0]; s_o3[4 * kBT * 3 + 3 * tid + 1] = dscale[1]; s_o3[4 * kBT * 3 + 3 * tid + 2] = dscale[2];
    }
    __syncthreads();

    // coalesced write-out
    float* outs[5] = {p.dL_dmean2D, p.dL_dcolor, p.dL_ddir, p.dL_dmean3D, p.dL_dscale};
#pragma unroll
    for (int a = 0; a < 5; a++) {
        if (outs[a] == nullptr) continue;
        float* dst = outs[a] + (size_t)base * 3;
        for (int f = tid; f < nvalid * 3; f += kBT) dst[f] = s_o3[a * kBT * 3 + f];
    }
    if (SEG) {
        if (seg_one) {
            float* dc = p.dseg.dc[seg_i] + seg_li0 * 3;
            float* rs = p.dseg.rest[seg_i] + seg_li0 * 45;
            const int nd = nvalid * 3, nr = nvalid * 45;
            if (((reinterpret_cast<uintptr_t>(dc) | reinterpret_cast<uintptr_t>(rs)) & 15) == 0 && (nvalid & 3) == 0) {
                for (int f4 = tid; f4 < nd / 4; f4 += kBT) reinterpret_cast<float4*>(dc)[f4] = reinterpret_cast<const float4*>(s_dc)[f4];
                for (int f4 = tid; f4 < nr / 4; f4 += kBT) reinterpret_cast<float4*>(rs)[f4] = reinterpret_cast<const float4*>(s_rest)[f4];
            } else {
                for (int f = tid; f < nd; f += kBT) dc[f] = s_dc[f];
                for (int f = tid; f < nr; f += kBT) rs[f] = s_rest[f];
            }
        } else {
            for (int f = tid; f < nvalid * 3; f += kBT) {
                const int row = f / 3, g = base + row;
                const int sgi = g >= seg_ns;
                p.dseg.dc[sgi][(size_t)(g - (sgi ? seg_ns : 0)) * 3 + (f - row * 3)] = s_dc[f];
            }
            for (int f = tid; f < nvalid * 45; f += kBT) {
                const int row = f / 45, g = base + row;
                const int sgi = g >= seg_ns;
                p.dseg.rest[sgi][(size_t)(g - (sgi ? seg_ns : 0)) * 45 + (f - row * 45)] = s_rest[f];
            }
        }
    } else if (has_sh) {
        float4* dst = reinterpret_cast<float4*>(p.dL_dsh + (size_t)base * 48);
        const int n4 = nvalid * 12;
        for (int f = tid; f < n4; f += kBT) {
            const int row = f / 12, col = f % 12;
            dst[f] = *reinterpret_cast<const float4*>(s_sh + row * kRow + col * 4);
        }
    }
}

}  // namespace

void launch_preprocess_fwd(const PreprocessParams& p, cudaStream_t s)
{
    if (p.P <= 0) return;
    if (EX_PRE_FWD_TEMPLATE && p.seg.enabled) preprocess_fwd_kernel<true><<<(p.P + 255) / 256, 256, 0, s>>>(p);
    else preprocess_fwd_kernel<false><<<(p.P + 255) / 256, 256, 0, s>>>(p);
}

void launch_preprocess_bwd(const PreprocessBwdParams& p, cudaStream_t s)
{
    if (p.P <= 0) return;
    if (p.shs == nullptr || p.M == 16) {     // segmented SH (shs == nullptr, seg.enabled) always has M == 16
        const size_t smem = sizeof(float) * (40 + 5 * kBT * 3 + ((p.shs != nullptr || p.seg.enabled) ? kBT * kRow : 0));
        if (p.seg.enabled) preprocess_bwd_staged_kernel<true><<<(p.P + kBT - 1) / kBT, kBT, smem, s>>>(p);
        else preprocess_bwd_staged_kernel<false><<<(p.P + kBT - 1) / kBT, kBT, smem, s>>>(p);
    } else {
        preprocess_bwd_kernel<<<(p.P + 255) / 256, 256, 0, s>>>(p);
    }
}

void launch_mark_visible(int P, const float* means3D, const float* view, const float* proj,
                         float min_depth, float max_depth, uint8_t* present, cudaStream_t s)
{
    if (P <= 0) return;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, view, proj, min_depth, max_depth, present);
}
