// Fused model front-end: the per-frame getters of CGaussianModel evaluated in
// one pass from the model's native static / dynamic tensors into the flat [P,.] rasterizer inputs
// (static Gaussians first), and their backward.
//
// Behavioural spec (reference, Python):
//   static  xyz      scene/c_gaussian_model.py:178-180   xyz + xyz_disp * t / duration
//   dynamic xyz      :182-193 + :108-119, utils/interpolations.py:81-93  Catmull-Rom Hermite on keyframes k-1..k+2
//   rotation         :195-215, interpolations.py:33-52   static raw, dynamic slerp WITHOUT shortest-path flip
//   opacity          :363-375, interpolations.py:55-61   sigmoid / bi-Gaussian window * sigmoid
//   scaling          :330-335                            exp
// Replaces ~40 elementwise PyTorch kernels and four of the five torch.cat copies per frame (the
// SH cat stays with the caller).  interp_type "cube" + rot_interp_type "slerp" (the defaults of every
// config, arguments/__init__.py:59) are implemented.
#include "../../include/ex4dgs_raster.h"
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>

namespace {

struct FrontParams {
    int Ns, Nd, K, k;                 // k = keyframe index of the frame
    float tf, durf;                   // t and duration as float32 scalars (torch semantics)
    float h00, h10, h01, h11;         // Hermite basis at delta
    float delta;                      // slerp parameter
    float tau, vmin236;               // (t + shift) / interval, var_min / 2.36
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

struct Slerp {
    float v1[4], v2[4], n1, n2;
    float d_raw, d, omega_raw, omega, s_raw, sn, a0, a1, ps_raw, ps, p0, p1;
    float r[4], ret[4], nret, out[4];
    bool use_r;
};

__device__ __forceinline__ void slerp_fwd(const float* q1, const float* q2, float t, Slerp& s)
{
    s.n1 = sqrtf(q1[0] * q1[0] + q1[1] * q1[1] + q1[2] * q1[2] + q1[3] * q1[3]);
    s.n2 = sqrtf(q2[0] * q2[0] + q2[1] * q2[1] + q2[2] * q2[2] + q2[3] * q2[3]);
#pragma unroll
    for (int i = 0; i < 4; i++) { s.v1[i] = q1[i] / s.n1; s.v2[i] = q2[i] / s.n2; }
    s.d_raw = s.v1[0] * s.v2[0] + s.v1[1] * s.v2[1] + s.v1[2] * s.v2[2] + s.v1[3] * s.v2[3];
    s.d = fminf(fmaxf(s.d_raw, -1.f + 1e-4f), 1.f - 1e-4f);
    s.omega_raw = acosf(s.d);
    s.omega = fmaxf(s.omega_raw, 1e-4f);
    s.s_raw = sinf(s.omega);
    s.sn = fmaxf(s.s_raw, 1e-4f);
    s.a0 = sinf((1.f - t) * s.omega) / s.sn;
    s.a1 = sinf(t * s.omega) / s.sn;
    s.ps_raw = s.a0 + s.a1;
    s.ps = fmaxf(s.ps_raw, 1e-4f);
    s.p0 = s.a0 / s.ps;
    s.p1 = s.a1 / s.ps;
    float asum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) { s.r[i] = s.v1[i] * s.p0 + s.v2[i] * s.p1; asum += fabsf(s.r[i]); }
    s.use_r = asum > 1e-4f;
#pragma unroll
    for (int i = 0; i < 4; i++) s.ret[i] = s.use_r ? s.r[i] : s.v1[i];
    s.nret = sqrtf(s.ret[0] * s.ret[0] + s.ret[1] * s.ret[1] + s.ret[2] * s.ret[2] + s.ret[3] * s.ret[3]);
#pragma unroll
    for (int i = 0; i < 4; i++) s.out[i] = s.ret[i] / s.nret;
}

// bi-Gaussian temporal window (interpolations.py:55-61); returns the window f and fills the pieces
// the backward needs
__device__ __forceinline__ float bigauss(float c0, float c1, float v0, float v1, float tau, float vmin236,
                                         bool& flat, int& jc, int& jv, float& m, float& Dn, float& g)
{
    const float m0 = tau - c0, m1 = tau - c1;
    jc = (m1 < m0) ? 1 : 0;                   // torch.min returns the first minimum
    m = jc ? m1 : m0;
    jv = ((tau > c0) || (tau > c1)) ? 1 : 0;
    const float v = jv ? v1 : v0;
    Dn = expf(v) + vmin236;
    g = expf(-1.f * ((m * m) / (Dn * Dn)));
    flat = ((c0 - tau) * (c1 - tau) < 0.f);
    return flat ? 1.0f : g;
}

__global__ void __launch_bounds__(256) frontend_fwd_kernel(
    FrontParams f,
    const float* __restrict__ xyz, const float* __restrict__ xyz_disp, const float* __restrict__ rotation,
    const float* __restrict__ scaling, const float* __restrict__ opacity,
    const float* __restrict__ xyz_motion, const float* __restrict__ rotation_motion, const float* __restrict__ scaling_motion,
    const float* __restrict__ opacity_motion, const float* __restrict__ opacity_center, const float* __restrict__ opacity_var,
    float* __restrict__ means3D, float* __restrict__ rotations, float* __restrict__ scales, float* __restrict__ opacities)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int P = f.Ns + f.Nd;
    if (i >= P) return;
    if (i < f.Ns) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            means3D[3 * i + c] = __fadd_rn(xyz[3 * i + c], __fdiv_rn(__fmul_rn(xyz_disp[3 * i + c], f.tf), f.durf));
            scales[3 * i + c] = expf(scaling[3 * i + c]);
        }
        reinterpret_cast<float4*>(rotations)[i] = __ldg(reinterpret_cast<const float4*>(rotation) + i);
        opacities[i] = sigmoidf_(opacity[i]);
        return;
    }
    const int j = i - f.Ns;
    const float* y = xyz_motion + ((size_t)j * f.K + (f.k - 1)) * 3;      // keyframes k-1 .. k+2
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float ykm1 = y[c], yk = y[3 + c], yk1 = y[6 + c], yk2 = y[9 + c];
        const float mk = __fdiv_rn(__fadd_rn(yk1, -ykm1), 2.f), mk1 = __fdiv_rn(__fadd_rn(yk2, -yk), 2.f);
        // h00*yk + h10*mk + h01*yk1 + h11*mk1, left to right, every op rounded (torch elementwise ops)
        float v = __fadd_rn(__fmul_rn(f.h00, yk), __fmul_rn(f.h10, mk));
        v = __fadd_rn(v, __fmul_rn(f.h01, yk1));
        v = __fadd_rn(v, __fmul_rn(f.h11, mk1));
        means3D[3 * i + c] = v;
        scales[3 * i + c] = expf(scaling_motion[3 * j + c]);
    }
    {
        const float4 qa = __ldg(reinterpret_cast<const float4*>(rotation_motion) + (size_t)j * f.K + f.k);
        const float4 qb = __ldg(reinterpret_cast<const float4*>(rotation_motion) + (size_t)j * f.K + f.k + 1);
        const float q1[4] = {qa.x, qa.y, qa.z, qa.w}, q2[4] = {qb.x, qb.y, qb.z, qb.w};
        Slerp s;
        slerp_fwd(q1, q2, f.delta, s);
        reinterpret_cast<float4*>(rotations)[i] = make_float4(s.out[0], s.out[1], s.out[2], s.out[3]);
    }
    {
        bool flat; int jc, jv; float m, Dn, g;
        const float w = bigauss(opacity_center[2 * j], opacity_center[2 * j + 1], opacity_var[2 * j], opacity_var[2 * j + 1],
                                f.tau, f.vmin236, flat, jc, jv, m, Dn, g);
        opacities[i] = w * sigmoidf_(opacity_motion[j]);
    }
}

__global__ void __launch_bounds__(256) frontend_bwd_kernel(
    FrontParams f,
    const float* __restrict__ rotation_motion, const float* __restrict__ scaling, const float* __restrict__ opacity,
    const float* __restrict__ scaling_motion, const float* __restrict__ opacity_motion,
    const float* __restrict__ opacity_center, const float* __restrict__ opacity_var,
    const float* __restrict__ g_means, const float* __restrict__ g_rot, const float* __restrict__ g_scales, const float* __restrict__ g_opac,
    float* __restrict__ d_xyz, float* __restrict__ d_disp, float* __restrict__ d_rotation, float* __restrict__ d_scaling, float* __restrict__ d_opacity,
    float* __restrict__ d_xyz_motion, float* __restrict__ d_rot_motion, float* __restrict__ d_scaling_motion,
    float* __restrict__ d_opacity_motion, float* __restrict__ d_center, float* __restrict__ d_var)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int P = f.Ns + f.Nd;
    if (i >= P) return;
    if (i < f.Ns) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float g = g_means[3 * i + c];
            d_xyz[3 * i + c] = g;
            d_disp[3 * i + c] = g * f.tf / f.durf;
            d_scaling[3 * i + c] = g_scales[3 * i + c] * expf(scaling[3 * i + c]);
        }
        reinterpret_cast<float4*>(d_rotation)[i] = __ldg(reinterpret_cast<const float4*>(g_rot) + i);
        const float sg = sigmoidf_(opacity[i]);
        d_opacity[i] = g_opac[i] * sg * (1.f - sg);
        return;
    }
    const int j = i - f.Ns;
    {   // Hermite: keyframes k-1..k+2 (everything else was zero-filled by the launcher)
        float* dy = d_xyz_motion + ((size_t)j * f.K + (f.k - 1)) * 3;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float g = g_means[3 * i + c];
            dy[c] = -0.5f * f.h10 * g;
            dy[3 + c] = (f.h00 - 0.5f * f.h11) * g;
            dy[6 + c] = (f.h01 + 0.5f * f.h10) * g;
            dy[9 + c] = 0.5f * f.h11 * g;
            d_scaling_motion[3 * j + c] = g_scales[3 * i + c] * expf(scaling_motion[3 * j + c]);
        }
    }
    {   // slerp backward
        const float4 qa = __ldg(reinterpret_cast<const float4*>(rotation_motion) + (size_t)j * f.K + f.k);
        const float4 qb = __ldg(reinterpret_cast<const float4*>(rotation_motion) + (size_t)j * f.K + f.k + 1);
        const float q1[4] = {qa.x, qa.y, qa.z, qa.w}, q2[4] = {qb.x, qb.y, qb.z, qb.w};
        Slerp s;
        slerp_fwd(q1, q2, f.delta, s);
        const float4 go4 = __ldg(reinterpret_cast<const float4*>(g_rot) + i);
        const float go[4] = {go4.x, go4.y, go4.z, go4.w};
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < 4; c++) dot += s.out[c] * go[c];
        float g_ret[4], g_v1[4] = {0, 0, 0, 0}, g_v2[4] = {0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < 4; c++) g_ret[c] = (go[c] - s.out[c] * dot) / s.nret;
        if (s.use_r) {
            float g_p0 = 0.f, g_p1 = 0.f;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                g_v1[c] = s.p0 * g_ret[c];
                g_v2[c] = s.p1 * g_ret[c];
                g_p0 += s.v1[c] * g_ret[c];
                g_p1 += s.v2[c] * g_ret[c];
            }
            float g_a0 = g_p0 / s.ps, g_a1 = g_p1 / s.ps;
            const float g_ps = -(g_p0 * s.a0 + g_p1 * s.a1) / (s.ps * s.ps);
            if (s.ps_raw >= 1e-4f) { g_a0 += g_ps; g_a1 += g_ps; }
            const float t = f.delta;
            float g_omega = g_a0 * cosf((1.f - t) * s.omega) * (1.f - t) / s.sn + g_a1 * cosf(t * s.omega) * t / s.sn;
            const float g_s = -(g_a0 * s.a0 + g_a1 * s.a1) / s.sn;
            if (s.s_raw >= 1e-4f) g_omega += g_s * cosf(s.omega);
            float g_d = 0.f;
            if (s.omega_raw >= 1e-4f) g_d = -g_omega / sqrtf(1.f - s.d * s.d);
            if (s.d_raw >= -1.f + 1e-4f && s.d_raw <= 1.f - 1e-4f) {
#pragma unroll
                for (int c = 0; c < 4; c++) { g_v1[c] += g_d * s.v2[c]; g_v2[c] += g_d * s.v1[c]; }
            }
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++) g_v1[c] = g_ret[c];
        }
        float d1 = 0.f, d2 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; c++) { d1 += s.v1[c] * g_v1[c]; d2 += s.v2[c] * g_v2[c]; }
        float4 o1, o2;
        o1.x = (g_v1[0] - s.v1[0] * d1) / s.n1; o1.y = (g_v1[1] - s.v1[1] * d1) / s.n1;
        o1.z = (g_v1[2] - s.v1[2] * d1) / s.n1; o1.w = (g_v1[3] - s.v1[3] * d1) / s.n1;
        o2.x = (g_v2[0] - s.v2[0] * d2) / s.n2; o2.y = (g_v2[1] - s.v2[1] * d2) / s.n2;
        o2.z = (g_v2[2] - s.v2[2] * d2) / s.n2; o2.w = (g_v2[3] - s.v2[3] * d2) / s.n2;
        reinterpret_cast<float4*>(d_rot_motion)[(size_t)j * f.K + f.k] = o1;
        reinterpret_cast<float4*>(d_rot_motion)[(size_t)j * f.K + f.k + 1] = o2;
    }
    {   // opacity window backward
        bool flat; int jc, jv; float m, Dn, g;
        const float w = bigauss(opacity_center[2 * j], opacity_center[2 * j + 1], opacity_var[2 * j], opacity_var[2 * j + 1],
                                f.tau, f.vmin236, flat, jc, jv, m, Dn, g);
        const float sg = sigmoidf_(opacity_motion[j]);
        const float go = g_opac[i];
        d_opacity_motion[j] = go * w * sg * (1.f - sg);
        float dc[2] = {0.f, 0.f}, dv[2] = {0.f, 0.f};
        if (!flat) {
            const float gg = go * sg;                       // dL/dg
            const float dg_dm = g * (-2.f * m / (Dn * Dn));
            const float dg_dD = g * (2.f * m * m / (Dn * Dn * Dn));
            dc[jc] = -gg * dg_dm;                           // m = tau - c_jc
            dv[jv] = gg * dg_dD * (Dn - f.vmin236);         // dD/dv = exp(v)
        }
        d_center[2 * j] = dc[0]; d_center[2 * j + 1] = dc[1];
        d_var[2 * j] = dv[0]; d_var[2 * j + 1] = dv[1];
    }
}

thread_local char g_ferr[256] = "";

bool make_params(int Ns, int Nd, int K, double t, double duration, double interval, double time_shift, double var_min,
                 FrontParams& f)
{
    f.Ns = Ns; f.Nd = Nd; f.K = K;
    // Python semantics of c_gaussian_model.py:184-187: t += shift; k = t // interval; delta = (t % interval) / interval
    const double tt = t + time_shift;
    const double kk = floor(tt / interval);
    double rem = fmod(tt, interval);
    if (rem < 0) rem += interval;
    const double d = rem / interval;
    f.k = (int)kk;
    f.tf = (float)t; f.durf = (float)duration;
    f.h00 = (float)(2 * pow(d, 3) - 3 * pow(d, 2) + 1);
    f.h10 = (float)(pow(d, 3) - 2 * pow(d, 2) + d);
    f.h01 = (float)(-2 * pow(d, 3) + 3 * pow(d, 2));
    f.h11 = (float)(pow(d, 3) - pow(d, 2));
    f.delta = (float)d;
    f.tau = (float)(tt / interval);
    f.vmin236 = (float)(var_min / 2.36);
    if (Nd > 0 && (f.k - 1 < 0 || f.k + 2 >= K)) return false;
    return true;
}

}  // namespace

extern "C" {

int ex4dgs_frontend_forward(int Ns, int Nd, int K,
                            const float* xyz, const float* xyz_disp, const float* rotation, const float* scaling, const float* opacity,
                            const float* xyz_motion, const float* rotation_motion, const float* scaling_motion,
                            const float* opacity_motion, const float* opacity_center, const float* opacity_var,
                            double t, double duration, double interval, double time_shift, double var_min,
                            float* means3D, float* rotations, float* scales, float* opacities, void* stream)
{
    FrontParams f;
    if (Ns < 0 || Nd < 0 || !make_params(Ns, Nd, K, t, duration, interval, time_shift, var_min, f)) return EX4DGS_ERR_INVALID;
    const int P = Ns + Nd;
    if (P == 0) return EX4DGS_OK;
    frontend_fwd_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(f, xyz, xyz_disp, rotation, scaling, opacity,
        xyz_motion, rotation_motion, scaling_motion, opacity_motion, opacity_center, opacity_var,
        means3D, rotations, scales, opacities);
    return cudaGetLastError() == cudaSuccess ? EX4DGS_OK : EX4DGS_ERR_CUDA;
}

int ex4dgs_frontend_backward(int Ns, int Nd, int K,
                             const float* rotation_motion, const float* scaling, const float* opacity,
                             const float* scaling_motion, const float* opacity_motion,
                             const float* opacity_center, const float* opacity_var,
                             double t, double duration, double interval, double time_shift, double var_min,
                             const float* dL_dmeans3D, const float* dL_drotations, const float* dL_dscales, const float* dL_dopacities,
                             float* dL_dxyz, float* dL_dxyz_disp, float* dL_drotation, float* dL_dscaling, float* dL_dopacity,
                             float* dL_dxyz_motion, float* dL_drotation_motion, float* dL_dscaling_motion,
                             float* dL_dopacity_motion, float* dL_dopacity_center, float* dL_dopacity_var, void* stream)
{
    FrontParams f;
    if (Ns < 0 || Nd < 0 || !make_params(Ns, Nd, K, t, duration, interval, time_shift, var_min, f)) return EX4DGS_ERR_INVALID;
    const int P = Ns + Nd;
    if (P == 0) return EX4DGS_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (Nd > 0) {
        if (cudaMemsetAsync(dL_dxyz_motion, 0, sizeof(float) * 3 * (size_t)Nd * K, s) != cudaSuccess) return EX4DGS_ERR_CUDA;
        if (cudaMemsetAsync(dL_drotation_motion, 0, sizeof(float) * 4 * (size_t)Nd * K, s) != cudaSuccess) return EX4DGS_ERR_CUDA;
    }
    frontend_bwd_kernel<<<(P + 255) / 256, 256, 0, s>>>(f, rotation_motion, scaling, opacity, scaling_motion, opacity_motion,
        opacity_center, opacity_var, dL_dmeans3D, dL_drotations, dL_dscales, dL_dopacities,
        dL_dxyz, dL_dxyz_disp, dL_drotation, dL_dscaling, dL_dopacity,
        dL_dxyz_motion, dL_drotation_motion, dL_dscaling_motion, dL_dopacity_motion, dL_dopacity_center, dL_dopacity_var);
    return cudaGetLastError() == cudaSuccess ? EX4DGS_OK : EX4DGS_ERR_CUDA;
}

}  // extern "C"
