// Row N4 of SURVEY.md §8f: the optimizer step that follows the backward in the reference's training
// iteration (train.py:250 `gaussians.optimizer.step()`, the optimizer being torch.optim.RAdam over 15
// parameter groups, scene/c_gaussian_model.py:430-449).
//
// torch's RAdam on CUDA runs its "foreach" implementation (torch/optim/radam.py, _multi_tensor_radam):
// nine element-wise passes per step over parameters, gradients and both moments (~100 bytes of HBM
// traffic per parameter element).  Here the whole step of ALL parameter tensors is ONE launch that
// reads p, g, m, v and writes p, m, v once (28 bytes per element) with 128-bit accesses: a table of up
// to 32 tensors travels in the kernel parameters, the work is cut into 4096-element chunks that a
// persistent grid (a multiple of the SM count) strides over.
//
// Arithmetic follows the foreach code path operation by operation (non-capturable, weight_decay = 0,
// maximize = False):
//     m   = m + (1 - beta1) * (g - m)                       _foreach_lerp_
//     v   = v * beta2;  v = v + (1 - beta2) * g * g         _foreach_mul_, _foreach_addcmul_
//     buf = 1 / ((sqrt(v) + eps) / S) + U                   _foreach_sqrt/add_/div_/reciprocal_/add_
//     p   = p + m * buf                                     _foreach_addcmul_
// with the per-tensor scalars S = -sqrt(1 - beta2^t) * lr * rect / (1 - beta1^t) and U = 0 when the
// variance is tractable (rho_t > 5), else S = 0 (buf = 1/inf = 0) and U = -lr / (1 - beta1^t); the host
// computes rect, S and U in double exactly like the Python code.
#include "common.cuh"

namespace {

constexpr int kChunk = 4096;          // elements per chunk (16 KB per array)
constexpr int kOptThreads = 256;

struct RAdamKernelParams {
    int n;
    float w1;            // 1 - beta1
    float beta2;
    float w2;            // 1 - beta2
    float eps;
    float grad_scale;    // gradients are multiplied by this first (1/world_size after a sum all-reduce)
    int* nan_flags;      // device int[n] or nullptr: nan_flags[i] = 1 when tensor i (check_nan set) received a NaN parameter
    unsigned chunk_end[EX_OPT_MAX_TENSORS];      // exclusive prefix of chunks per tensor
    RAdamTensorDesc t[EX_OPT_MAX_TENSORS];
};

__device__ __forceinline__ void radam_elem(float& p, float g, float& m, float& v, const RAdamKernelParams& k,
                                           const RAdamTensorDesc& d)
{
    // FMA placement = what nvcc makes of the foreach functors (lerp: self + w * (end - self);
    // addcmul: self + value * (t1 * t2))
    if (d.sanitize_grad) {       // torch.nan_to_num (train.py:246-248): NaN -> 0, +-inf -> +-FLT_MAX
        if (g != g) g = 0.0f;
        else g = fminf(fmaxf(g, -3.402823466e+38f), 3.402823466e+38f);
    }
    g = __fmul_rn(g, k.grad_scale);
    m = __fmaf_rn(k.w1, __fsub_rn(g, m), m);
    v = __fmul_rn(v, k.beta2);
    v = __fmaf_rn(k.w2, __fmul_rn(g, g), v);
    float buf;
    if (d.rectified) buf = __frcp_rn(__fdiv_rn(__fadd_rn(__fsqrt_rn(v), k.eps), d.S));
    else buf = d.U;
    p = __fadd_rn(p, __fmul_rn(m, buf));
}

__global__ void __launch_bounds__(kOptThreads) radam_kernel(const __grid_constant__ RAdamKernelParams k, unsigned total_chunks)
{
    for (unsigned c = blockIdx.x; c < total_chunks; c += gridDim.x) {
        int ti = 0;
#pragma unroll 1
        while (ti + 1 < k.n && c >= k.chunk_end[ti]) ti++;
        const RAdamTensorDesc& d = k.t[ti];
        const unsigned first = ti ? k.chunk_end[ti - 1] : 0u;
        const size_t base = (size_t)(c - first) * kChunk;
        const size_t left = d.numel - base;
        const int cnt = left < (size_t)kChunk ? (int)left : kChunk;
        float* __restrict__ P = d.param + base;
        const float* __restrict__ G = d.grad + base;
        float* __restrict__ M = d.exp_avg + base;
        float* __restrict__ V = d.exp_avg_sq + base;
        bool bad = false;
        if (d.aligned && cnt == kChunk) {
#pragma unroll
            for (int i = 0; i < kChunk / (4 * kOptThreads); i++) {
                const int f = i * kOptThreads + threadIdx.x;
                float4 p = reinterpret_cast<float4*>(P)[f];
                const float4 g = __ldg(reinterpret_cast<const float4*>(G) + f);
                float4 m = reinterpret_cast<float4*>(M)[f];
                float4 v = reinterpret_cast<float4*>(V)[f];
                radam_elem(p.x, g.x, m.x, v.x, k, d);
                radam_elem(p.y, g.y, m.y, v.y, k, d);
                radam_elem(p.z, g.z, m.z, v.z, k, d);
                radam_elem(p.w, g.w, m.w, v.w, k, d);
                bad |= (p.x != p.x) | (p.y != p.y) | (p.z != p.z) | (p.w != p.w);
                reinterpret_cast<float4*>(P)[f] = p;
                reinterpret_cast<float4*>(M)[f] = m;
                reinterpret_cast<float4*>(V)[f] = v;
            }
        } else {
            for (int f = threadIdx.x; f < cnt; f += kOptThreads) {
                float p = P[f], m = M[f], v = V[f];
                radam_elem(p, __ldg(G + f), m, v, k, d);
                bad |= (p != p);
                P[f] = p; M[f] = m; V[f] = v;
            }
        }
        // prune_nan_points (c_gaussian_model.py:1229-1241) tests _xyz / _xyz_motion for NaN after every step with
        // two reductions and two host waits; here the written values are tested on the fly
        if (d.check_nan && k.nan_flags && bad) k.nan_flags[d.index] = 1;
    }
}

}  // namespace

cudaError_t launch_radam(const RAdamTensorDesc* tensors, int n, double beta1, double beta2, double eps, double grad_scale,
                         int* nan_flags, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    RAdamKernelParams k;
    k.n = n;
    k.w1 = (float)(1.0 - beta1);      // Python computes 1 - beta in double, the foreach kernels receive it as float
    k.beta2 = (float)beta2;
    k.w2 = (float)(1.0 - beta2);
    k.eps = (float)eps;
    k.grad_scale = (float)grad_scale;
    k.nan_flags = nan_flags;
    unsigned long long chunks = 0;
    for (int i = 0; i < n; i++) {
        k.t[i] = tensors[i];
        const uintptr_t a = (uintptr_t)tensors[i].param | (uintptr_t)tensors[i].grad | (uintptr_t)tensors[i].exp_avg |
                            (uintptr_t)tensors[i].exp_avg_sq;
        k.t[i].aligned = (a & 15) == 0;
        chunks += (tensors[i].numel + kChunk - 1) / kChunk;
        if (chunks > 0xFFFFFFFFull) return cudaErrorInvalidValue;
        k.chunk_end[i] = (unsigned)chunks;
    }
    for (int i = n; i < EX_OPT_MAX_TENSORS; i++) k.chunk_end[i] = (unsigned)chunks;
    if (chunks == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned long long want = (unsigned long long)sms * 8;
    const unsigned grid = (unsigned)(chunks < want ? chunks : want);
    radam_kernel<<<grid, kOptThreads, 0, s>>>(k, (unsigned)chunks);
    return cudaGetLastError();
}
