// Per-iteration bookkeeping around the optimizer step ("stats updates"):
//
//  (1) iteration_stats_kernel - what train.py:196-215 does after loss.backward() with ~60 PyTorch
//      kernels, most of them boolean-mask gathers/scatters (each `x[mask]` runs nonzero() and waits for
//      its count on the host): CGaussianModel.mark_prune_stats (scene/c_gaussian_model.py:1105-1117), the
//      max_radii2D updates (train.py:205-206), add_densification_stats (:1095-1103) and add_l1_ssim_stats
//      (:1119-1145), for the static and the dynamic Gaussians.  One thread per Gaussian, every array read
//      and written once, no host synchronisation.  HBM-bound: 4 + 12 + 12 bytes read, up to 9 x 8 bytes
//      read-modify-written per Gaussian.
//
//  (2) regularizer_kernel - the two default-on regularisation terms of the loss (train.py:156-162) and
//      their gradients, added in place to the gradients the backward pass has produced:
//          static_reg * mean_i log(|xyz_disp_i| + 0.001)
//          motion_reg * mean_{i, k>=1} |xyz_motion[i,0] - xyz_motion[i,k]|
//      (rot_reg is 0.0 in arguments/__init__.py:136 and `if opt.rot_reg > 0` never runs.)  The reference
//      lets autograd run slice, sub, norm, mean and their backward over the [Nd,K,3] keyframe tensor
//      (eight passes); here the tensor is read once and its gradient read-modify-written once, one half-warp per
//      [K,3] row, staged through shared memory so that both cross HBM as whole 128-bit lines.  Sums are
//      reduced in a fixed order (per-block partials in double, then one block): bit-reproducible.
#include "common.cuh"

namespace {

// torch.clamp_min propagates NaN (fmaxf would not)
__device__ __forceinline__ float clamp_min_nan(float x, float lo) { return x < lo ? lo : x; }

__global__ void __launch_bounds__(256) iteration_stats_kernel(const __grid_constant__ IterStatsParams p)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.Ns + p.Nd) return;
    const bool dyn = i >= p.Ns;
    const StatsArrays& a = dyn ? p.dyn : p.stat;
    const int j = dyn ? i - p.Ns : i;
    const int radius = __ldg(p.radii + i);
    const float rf = (float)radius;               // torch.min/max(float tensor, int tensor) promotes to float
    const bool vis = radius > 0;                  // gaussian_renderer/__init__.py:123 visibility_filter
    float e0 = 0.f, e1 = 0.f, e2 = 0.f;
    if (p.grad_error) {
        e0 = __ldg(p.grad_error + 3 * (size_t)i);
        e1 = __ldg(p.grad_error + 3 * (size_t)i + 1);
        e2 = __ldg(p.grad_error + 3 * (size_t)i + 2);
        // mark_prune_stats (c_gaussian_model.py:1105-1117): filter = error-tensor gradient [:, 0] > 0
        if (e0 > 0.0f) {
            const float m = a.min_radii2D[j];
            a.min_radii2D[j] = (rf < m || rf != rf) ? rf : m;     // torch.min propagates NaN; rf never is
        }
    }
    if (!p.densify || !vis) return;
    // train.py:205-206
    {
        const float m = a.max_radii2D[j];
        a.max_radii2D[j] = (rf > m) ? rf : m;
    }
    // add_densification_stats (c_gaussian_model.py:1095-1103)
    {
        const float gx = __ldg(p.grad_means2D + 3 * (size_t)i), gy = __ldg(p.grad_means2D + 3 * (size_t)i + 1);
        a.xyz_gradient_accum[j] += sqrtf(gx * gx + gy * gy);
        a.denom[j] += 1.0f;
    }
    // add_l1_ssim_stats (c_gaussian_model.py:1119-1145)
    if (p.grad_error) {
        const float den = clamp_min_nan(e0, 1e-4f);
        const float l1 = e1 / den;
        const float emin = a.error_min[j];
        const bool better = (emin > l1) && (e0 > 0.01f);
        a.error_accum[j] += l1;
        if (better) {
            a.error_min_timestamp[j] = p.timestamp;
            a.error_min[j] = l1;
        }
        a.ssim_error_accum[j] += e2 / den;
        a.error_denom[j] += (e0 > 0.0f) ? 1.0f : 0.0f;
    }
}

constexpr int kRegThreads = 256;
constexpr int kRegWarps = kRegThreads / 32;
constexpr int kRegMaxK = 64;          // keyframes per row the staged path holds (2 x 768 B per warp); more: direct path

// One half-warp per dynamic Gaussian row ([K,3] floats, contiguous) and one thread per static Gaussian; block
// partial sums (double) go to part[2 * block + {0,1}].
__global__ void __launch_bounds__(kRegThreads, 5) regularizer_kernel(const __grid_constant__ RegParams p)
{
    __shared__ double s_part[2][kRegWarps];
    __shared__ __align__(16) float s_row[kRegWarps][2][2][kRegMaxK * 3];     // [warp][half-warp][values | gradient]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned full = 0xffffffffu;
    const float gscale = p.dL_dloss ? __ldg(p.dL_dloss) : 1.0f;
    double sum_static = 0.0, sum_motion = 0.0;

    if (p.static_coef != 0.0f) {
        const float c = p.static_coef * gscale;        // static_reg / Ns, times the gradient arriving at the loss
        for (long long i = (long long)blockIdx.x * kRegThreads + threadIdx.x; i < p.Ns; i += (long long)gridDim.x * kRegThreads) {
            const float x = __ldg(p.xyz_disp + 3 * i), y = __ldg(p.xyz_disp + 3 * i + 1), z = __ldg(p.xyz_disp + 3 * i + 2);
            const float n = sqrtf(x * x + y * y + z * z);
            sum_static += (double)logf(n + 0.001f);
            if (p.dL_dxyz_disp) {
                // d log(n + 0.001) = 1 / (n + 0.001); norm backward: x * (g / n), 0 where n == 0
                const float s = (n == 0.0f) ? 0.0f : (c / (n + 0.001f)) / n;
                float* g = p.dL_dxyz_disp + 3 * i;
                if (p.accumulate_disp) { g[0] += x * s; g[1] += y * s; g[2] += z * s; }
                else { g[0] = x * s; g[1] = y * s; g[2] = z * s; }
            }
        }
    }
    if (p.motion_coef != 0.0f && p.K > 1) {
        const float c = p.motion_coef * gscale;        // motion_reg / (Nd (K - 1))
        const long long wid = (long long)blockIdx.x * kRegWarps + warp;
        const int row_f = p.K * 3;                     // floats per row
        // rows are staged through shared memory so that the keyframe tensor and its gradient cross HBM as whole
        // 128-bit lines (a lane's own (x, y, z) triple sits at a 12-byte stride)
        const bool staged = (p.K <= kRegMaxK) && (row_f % 4 == 0) &&
                            ((reinterpret_cast<uintptr_t>(p.xyz_motion) | reinterpret_cast<uintptr_t>(p.dL_dxyz_motion)) & 15) == 0;
        if (staged) {
            // one HALF-warp per row: two rows of a warp in flight at once (the kernel waits on its row loads - rows in
            // flight per SM are what hides the latency), K - 1 keyframes over 16 lanes
            const int hw = lane >> 4, hl = lane & 15;
            float* sv = s_row[warp][hw][0];
            float* sg = s_row[warp][hw][1];
            const int n4 = row_f / 4;
            for (long long pair = wid; 2 * pair < p.Nd; pair += (long long)gridDim.x * kRegWarps) {
                const long long i = 2 * pair + hw;
                const bool valid = i < p.Nd;
                const float* row = p.xyz_motion + i * (long long)row_f;
                float* grow = (p.dL_dxyz_motion && valid) ? p.dL_dxyz_motion + i * (long long)row_f : nullptr;
                float s0x = 0.f, s0y = 0.f, s0z = 0.f, sn = 0.f;
                if (valid) {
                    for (int f = hl; f < n4; f += 16) {
                        reinterpret_cast<float4*>(sv)[f] = __ldg(reinterpret_cast<const float4*>(row) + f);
                        if (grow) reinterpret_cast<float4*>(sg)[f] = p.accumulate_motion ? reinterpret_cast<const float4*>(grow)[f]
                                                                                         : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                __syncwarp();
                if (valid) {
                    const float x0 = sv[0], y0 = sv[1], z0 = sv[2];
                    for (int k = 1 + hl; k < p.K; k += 16) {
                        const float dx = x0 - sv[3 * k], dy = y0 - sv[3 * k + 1], dz = z0 - sv[3 * k + 2];
                        const float n = sqrtf(dx * dx + dy * dy + dz * dz);
                        sn += n;
                        if (grow) {
                            const float s = (n == 0.0f) ? 0.0f : c / n;
                            const float gx = dx * s, gy = dy * s, gz = dz * s;     // d/d y_0 ; d/d y_k is the negative
                            s0x += gx; s0y += gy; s0z += gz;
                            sg[3 * k] -= gx; sg[3 * k + 1] -= gy; sg[3 * k + 2] -= gz;
                        }
                    }
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {          // within the half-warp
                    s0x += __shfl_xor_sync(full, s0x, o);
                    s0y += __shfl_xor_sync(full, s0y, o);
                    s0z += __shfl_xor_sync(full, s0z, o);
                    sn += __shfl_xor_sync(full, sn, o);
                }
                if (hl == 0 && valid) sum_motion += (double)sn;
                if (grow) {
                    if (hl == 0) { sg[0] += s0x; sg[1] += s0y; sg[2] += s0z; }
                }
                __syncwarp();
                if (grow) {
                    for (int f = hl; f < n4; f += 16) reinterpret_cast<float4*>(grow)[f] = reinterpret_cast<const float4*>(sg)[f];
                }
                __syncwarp();                      // the buffers are reused by the half-warp's next row
            }
        } else {
            for (long long i = wid; i < p.Nd; i += (long long)gridDim.x * kRegWarps) {
                const float* row = p.xyz_motion + i * (long long)row_f;
                float* grow = p.dL_dxyz_motion ? p.dL_dxyz_motion + i * (long long)row_f : nullptr;
                float s0x = 0.f, s0y = 0.f, s0z = 0.f, sn = 0.f;
                const float x0 = __ldg(row), y0 = __ldg(row + 1), z0 = __ldg(row + 2);
                for (int k = 1 + lane; k < p.K; k += 32) {
                    const float dx = x0 - __ldg(row + 3 * k), dy = y0 - __ldg(row + 3 * k + 1), dz = z0 - __ldg(row + 3 * k + 2);
                    const float n = sqrtf(dx * dx + dy * dy + dz * dz);
                    sn += n;
                    if (grow) {
                        const float s = (n == 0.0f) ? 0.0f : c / n;
                        const float gx = dx * s, gy = dy * s, gz = dz * s;
                        s0x += gx; s0y += gy; s0z += gz;
                        if (p.accumulate_motion) { grow[3 * k] -= gx; grow[3 * k + 1] -= gy; grow[3 * k + 2] -= gz; }
                        else { grow[3 * k] = -gx; grow[3 * k + 1] = -gy; grow[3 * k + 2] = -gz; }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    s0x += __shfl_xor_sync(full, s0x, o);
                    s0y += __shfl_xor_sync(full, s0y, o);
                    s0z += __shfl_xor_sync(full, s0z, o);
                    sn += __shfl_xor_sync(full, sn, o);
                }
                if (lane == 0) sum_motion += (double)sn;
                if (lane == 0 && grow) {
                    if (p.accumulate_motion) { grow[0] += s0x; grow[1] += s0y; grow[2] += s0z; }
                    else { grow[0] = s0x; grow[1] = s0y; grow[2] = s0z; }
                }
            }
        }
    }
    // block partials, fixed order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum_static += __shfl_xor_sync(full, sum_static, o);
        sum_motion += __shfl_xor_sync(full, sum_motion, o);
    }
    if (lane == 0) { s_part[0][warp] = sum_static; s_part[1][warp] = sum_motion; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < kRegWarps; w++) { a += s_part[0][w]; b += s_part[1][w]; }
        p.part[2 * blockIdx.x] = a;
        p.part[2 * blockIdx.x + 1] = b;
    }
}

// fixed-order reduction of the per-block partials: thread t sums blocks t, t + 256, ... ; then a fixed tree
__global__ void __launch_bounds__(256) regularizer_finish_kernel(const double* part, int blocks, float static_coef, float motion_coef, float* out2)
{
    __shared__ double sa[256], sb[256];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < blocks; i += 256) { a += part[2 * i]; b += part[2 * i + 1]; }
    sa[threadIdx.x] = a; sb[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sb[threadIdx.x] += sb[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out2[0] = (float)(sa[0] * (double)static_coef);
        out2[1] = (float)(sb[0] * (double)motion_coef);
    }
}

}  // namespace

cudaError_t launch_iteration_stats(const IterStatsParams& p, cudaStream_t s)
{
    const long long n = (long long)p.Ns + p.Nd;
    if (n <= 0) return cudaSuccess;
    iteration_stats_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p);
    return cudaGetLastError();
}

int regularizer_blocks()
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms * 10;      // two waves of the 5 resident CTAs per SM
}

cudaError_t launch_regularizers(RegParams p, float* out2, cudaStream_t s)
{
    const int blocks = regularizer_blocks();
    regularizer_kernel<<<blocks, kRegThreads, 0, s>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    regularizer_finish_kernel<<<1, 256, 0, s>>>(p.part, blocks, p.static_coef, p.motion_coef, out2);
    return cudaGetLastError();
}
