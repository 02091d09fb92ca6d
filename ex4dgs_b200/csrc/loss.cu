// Row N2 of SURVEY.md §8: the photometric loss that follows the rasterizer in the training step,
//   loss = (1 - lambda) * mean|img - gt| + lambda * (1 - mean(ssim_map(img, gt)))
// (reference train.py:144-146; utils/loss_utils.py:22-25 l1_loss, :33-81 ssim with an 11x11
// Gaussian window, sigma 1.5, zero padding 5, C1 = 0.01^2, C2 = 0.03^2), plus the two per-pixel
// error maps of the l1_accum branch (train.py:149-151): mean over channels of |img - gt| and of the
// SSIM map.  The reference runs 5 depthwise conv2d + ~20 elementwise kernels per ssim() call and
// calls ssim() twice per step; here it is one forward kernel (separable window in shared memory,
// all five moments at once, the partial derivatives of the SSIM map stored for the backward), one
// tiny deterministic reduction, and one backward kernel (the three derivative maps convolved with
// the same symmetric window and combined with the L1 sign term).
//
// HBM traffic per frame (C = 3 channels, N = C*H*W floats = 16.4 MB at 1352x1014):
//   forward  reads img, gt (2N), writes 3 derivative maps (3N) + 2 error maps (2N/3)
//   backward reads 3 derivative maps + img + gt (5N), writes dL_dimg (N)           ~ 11.7 N = 192 MB
#include "common.cuh"

namespace {

constexpr int LT_X = 32;            // tile width  (one warp per tile row)
constexpr int LT_Y = 16;            // tile height
constexpr int LHALO = 5;            // window radius
constexpr int LWIN = 11;
constexpr int LP_X = LT_X + 2 * LHALO;   // 42
constexpr int LP_Y = LT_Y + 2 * LHALO;   // 26
constexpr int LTHREADS = LT_X * LT_Y;    // 512

struct LossParams {
    int H, W;
    const float* img;      // [3,H,W]
    const float* gt;       // [3,H,W]
    float* dmap;           // [3][3,H,W]  d ssim_map / d (conv(img), conv(img^2), conv(img*gt))
    float* l1_err;         // [H,W]
    float* ssim_err;       // [H,W]
    double* partials;      // [2 * blocks]  per-block sums of |img-gt| and ssim_map
    float win[LWIN];       // 1-D window (kernel-parameter constant bank: uniform reads)
};

__global__ void __launch_bounds__(LTHREADS)
loss_fwd_kernel(const LossParams p)
{
    __shared__ float s_x[LP_Y][LP_X];
    __shared__ float s_y[LP_Y][LP_X];
    __shared__ float s_h[5][LP_Y][LT_X];
    __shared__ float s_red[2][LTHREADS / 32];

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * LT_X + tx;
    const int x0 = blockIdx.x * LT_X, y0 = blockIdx.y * LT_Y;
    const int px = x0 + tx, py = y0 + ty;
    const bool inside = px < p.W && py < p.H;
    const size_t HW = (size_t)p.H * p.W;
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;

    float l1_px = 0.f, ssim_px = 0.f;
    for (int ch = 0; ch < 3; ch++) {
        const float* gx = p.img + ch * HW;
        const float* gy = p.gt + ch * HW;
        for (int i = tid; i < LP_Y * LP_X; i += LTHREADS) {
            const int r = i / LP_X, c = i - r * LP_X;
            const int yy = y0 + r - LHALO, xx = x0 + c - LHALO;
            float vx = 0.f, vy = 0.f;
            if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
                vx = __ldg(gx + (size_t)yy * p.W + xx);
                vy = __ldg(gy + (size_t)yy * p.W + xx);
            }
            s_x[r][c] = vx;
            s_y[r][c] = vy;
        }
        __syncthreads();
        // rows of the window: 26 x 32 outputs, 5 moments each
        for (int i = tid; i < LP_Y * LT_X; i += LTHREADS) {
            const int r = i / LT_X, c = i - r * LT_X;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
            for (int k = 0; k < LWIN; k++) {
                const float w = p.win[k], x = s_x[r][c + k], y = s_y[r][c + k];
                const float wx = w * x, wy = w * y;
                a0 += wx; a1 += wy;
                a2 = fmaf(wx, x, a2); a3 = fmaf(wy, y, a3); a4 = fmaf(wx, y, a4);
            }
            s_h[0][r][c] = a0; s_h[1][r][c] = a1; s_h[2][r][c] = a2; s_h[3][r][c] = a3; s_h[4][r][c] = a4;
        }
        __syncthreads();
        float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < LWIN; k++) {
            const float w = p.win[k];
            mu1 = fmaf(w, s_h[0][ty + k][tx], mu1);
            mu2 = fmaf(w, s_h[1][ty + k][tx], mu2);
            e11 = fmaf(w, s_h[2][ty + k][tx], e11);
            e22 = fmaf(w, s_h[3][ty + k][tx], e22);
            e12 = fmaf(w, s_h[4][ty + k][tx], e12);
        }
        if (inside) {
            const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
            const float sig1 = e11 - mu1_sq, sig2 = e22 - mu2_sq, sig12 = e12 - mu12;
            const float n1 = 2.f * mu12 + C1, n2 = 2.f * sig12 + C2;
            const float d1 = mu1_sq + mu2_sq + C1, d2 = sig1 + sig2 + C2;
            const float inv = 1.f / (d1 * d2);
            const float m = n1 * n2 * inv;
            // partials of m with (mu1, sigma1_sq, sigma12) as independent variables ...
            const float f_mu1 = 2.f * mu2 * n2 * inv - m * (2.f * mu1) / d1;
            const float f_s1 = -m / d2;
            const float f_s12 = 2.f * n1 * inv;
            // ... and with (conv(x), conv(x^2), conv(x*y)) as independent variables
            const size_t o = ch * HW + (size_t)py * p.W + px;
            p.dmap[o] = f_mu1 - 2.f * mu1 * f_s1 - mu2 * f_s12;
            p.dmap[3 * HW + o] = f_s1;
            p.dmap[6 * HW + o] = f_s12;
            ssim_px += m;
            l1_px += fabsf(s_x[ty + LHALO][tx + LHALO] - s_y[ty + LHALO][tx + LHALO]);
        }
        __syncthreads();
    }
    if (inside) {
        p.l1_err[(size_t)py * p.W + px] = l1_px / 3.f;
        p.ssim_err[(size_t)py * p.W + px] = ssim_px / 3.f;
    }
    // block sums, fixed order
    float a = l1_px, b = ssim_px;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (tx == 0) { s_red[0][ty] = a; s_red[1][ty] = b; }
    __syncthreads();
    if (tid == 0) {
        double sa = 0.0, sb = 0.0;
        for (int i = 0; i < LTHREADS / 32; i++) { sa += s_red[0][i]; sb += s_red[1][i]; }
        const int blk = blockIdx.y * gridDim.x + blockIdx.x;
        p.partials[2 * blk] = sa;
        p.partials[2 * blk + 1] = sb;
    }
}

// out[0] = loss, out[1] = Ll1 = mean|img-gt|, out[2] = mean ssim_map
__global__ void __launch_bounds__(256)
loss_reduce_kernel(const double* partials, int blocks, double inv_n, float lambda, float* out)
{
    __shared__ double s[2][256];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < blocks; i += 256) { a += partials[2 * i]; b += partials[2 * i + 1]; }
    s[0][threadIdx.x] = a; s[1][threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s[0][threadIdx.x] += s[0][threadIdx.x + o]; s[1][threadIdx.x] += s[1][threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float l1 = (float)(s[0][0] * inv_n), ss = (float)(s[1][0] * inv_n);
        out[0] = (1.0f - lambda) * l1 + lambda * (1.0f - ss);
        out[1] = l1;
        out[2] = ss;
    }
}

struct LossBwdParams {
    int H, W;
    const float* img;
    const float* gt;
    const float* dmap;
    const float* dL_dloss;   // device scalar
    float lambda, inv_n;
    float* dL_dimg;          // [3,H,W]
    float win[LWIN];
};

__global__ void __launch_bounds__(LTHREADS)
loss_bwd_kernel(const LossBwdParams p)
{
    __shared__ float s_d[3][LP_Y][LP_X];
    __shared__ float s_h[3][LP_Y][LT_X];

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * LT_X + tx;
    const int x0 = blockIdx.x * LT_X, y0 = blockIdx.y * LT_Y;
    const int px = x0 + tx, py = y0 + ty;
    const bool inside = px < p.W && py < p.H;
    const size_t HW = (size_t)p.H * p.W;
    const float g = __ldg(p.dL_dloss);
    const float g_l1 = g * (1.f - p.lambda) * p.inv_n, g_ss = -g * p.lambda * p.inv_n;

    for (int ch = 0; ch < 3; ch++) {
        for (int i = tid; i < LP_Y * LP_X; i += LTHREADS) {
            const int r = i / LP_X, c = i - r * LP_X;
            const int yy = y0 + r - LHALO, xx = x0 + c - LHALO;
            float v0 = 0.f, v1 = 0.f, v2 = 0.f;
            if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
                const size_t o = ch * HW + (size_t)yy * p.W + xx;
                v0 = __ldg(p.dmap + o); v1 = __ldg(p.dmap + 3 * HW + o); v2 = __ldg(p.dmap + 6 * HW + o);
            }
            s_d[0][r][c] = v0; s_d[1][r][c] = v1; s_d[2][r][c] = v2;
        }
        __syncthreads();
        for (int i = tid; i < LP_Y * LT_X; i += LTHREADS) {
            const int r = i / LT_X, c = i - r * LT_X;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int k = 0; k < LWIN; k++) {
                const float w = p.win[k];
                a0 = fmaf(w, s_d[0][r][c + k], a0);
                a1 = fmaf(w, s_d[1][r][c + k], a1);
                a2 = fmaf(w, s_d[2][r][c + k], a2);
            }
            s_h[0][r][c] = a0; s_h[1][r][c] = a1; s_h[2][r][c] = a2;
        }
        __syncthreads();
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int k = 0; k < LWIN; k++) {
            const float w = p.win[k];
            c0 = fmaf(w, s_h[0][ty + k][tx], c0);
            c1 = fmaf(w, s_h[1][ty + k][tx], c1);
            c2 = fmaf(w, s_h[2][ty + k][tx], c2);
        }
        if (inside) {
            const size_t o = ch * HW + (size_t)py * p.W + px;
            const float x = __ldg(p.img + o), y = __ldg(p.gt + o);
            const float d = x - y;
            const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);     // torch: d|u|/du = sign(u), 0 at 0
            p.dL_dimg[o] = g_l1 * sgn + g_ss * (c0 + 2.f * x * c1 + y * c2);
        }
        __syncthreads();
    }
}

static void fill_window(float* w)
{
    // utils/loss_utils.py:33-35: exp(-(x - 5)^2 / (2 * 1.5^2)) as float32, normalised by the float32 sum
    float s = 0.f;
    for (int i = 0; i < LWIN; i++) { w[i] = (float)exp(-(double)((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5)); s += w[i]; }
    for (int i = 0; i < LWIN; i++) w[i] = w[i] / s;
}

}  // namespace

size_t loss_scratch_bytes(int W, int H)
{
    const size_t blocks = (size_t)((W + LT_X - 1) / LT_X) * ((H + LT_Y - 1) / LT_Y);
    return 9 * (size_t)W * H * sizeof(float) + 2 * blocks * sizeof(double) + 256;
}

cudaError_t launch_loss_forward(int W, int H, const float* img, const float* gt, float lambda, char* scratch,
                                float* out3, float* l1_err, float* ssim_err, cudaStream_t stream)
{
    const dim3 grid((W + LT_X - 1) / LT_X, (H + LT_Y - 1) / LT_Y), block(LT_X, LT_Y);
    const size_t HW = (size_t)W * H;
    char* base = (char*)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
    LossParams p;
    p.H = H; p.W = W; p.img = img; p.gt = gt;
    p.partials = (double*)base;
    p.dmap = (float*)(base + (((size_t)2 * grid.x * grid.y * sizeof(double) + 255) & ~(size_t)255));
    p.l1_err = l1_err; p.ssim_err = ssim_err;
    fill_window(p.win);
    loss_fwd_kernel<<<grid, block, 0, stream>>>(p);
    loss_reduce_kernel<<<1, 256, 0, stream>>>(p.partials, (int)(grid.x * grid.y), 1.0 / (3.0 * (double)HW), lambda, out3);
    return cudaGetLastError();
}

cudaError_t launch_loss_backward(int W, int H, const float* img, const float* gt, float lambda, const char* scratch,
                                 const float* dL_dloss, float* dL_dimg, cudaStream_t stream)
{
    const dim3 grid((W + LT_X - 1) / LT_X, (H + LT_Y - 1) / LT_Y), block(LT_X, LT_Y);
    const size_t HW = (size_t)W * H;
    const char* base = (const char*)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
    LossBwdParams p;
    p.H = H; p.W = W; p.img = img; p.gt = gt;
    p.dmap = (const float*)(base + (((size_t)2 * grid.x * grid.y * sizeof(double) + 255) & ~(size_t)255));
    p.dL_dloss = dL_dloss; p.lambda = lambda; p.inv_n = (float)(1.0 / (3.0 * (double)HW));
    p.dL_dimg = dL_dimg;
    fill_window(p.win);
    loss_bwd_kernel<<<grid, block, 0, stream>>>(p);
    return cudaGetLastError();
}

// ---- plain L1 loss (utils/loss_utils.py:22-25: torch.abs(network_output - gt).mean()) ---------------------
// What an L1-only training or evaluation step computes around the rasterizer.  torch: sub, abs, mean forward and
// fill, div, sign, mul backward - seven element-wise passes (~7 N floats of traffic each way); here one read of both
// images per direction.  Per-block partial sums in double, reduced in a fixed order (bit-reproducible).
namespace {

constexpr int kL1Threads = 256;

__global__ void __launch_bounds__(kL1Threads) l1_fwd_kernel(size_t n, const float* __restrict__ a, const float* __restrict__ b,
                                                             double* __restrict__ part)
{
    __shared__ double s_w[kL1Threads / 32];
    double acc = 0.0;
    const size_t n4 = n / 4;
    const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
    const size_t stride = (size_t)gridDim.x * kL1Threads;
    if (vec) {
        for (size_t i = (size_t)blockIdx.x * kL1Threads + threadIdx.x; i < n4; i += stride) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(a) + i), y = __ldg(reinterpret_cast<const float4*>(b) + i);
            acc += (double)(fabsf(x.x - y.x) + fabsf(x.y - y.y)) + (double)(fabsf(x.z - y.z) + fabsf(x.w - y.w));
        }
        for (size_t i = n4 * 4 + (size_t)blockIdx.x * kL1Threads + threadIdx.x; i < n; i += stride) acc += (double)fabsf(a[i] - b[i]);
    } else {
        for (size_t i = (size_t)blockIdx.x * kL1Threads + threadIdx.x; i < n; i += stride) acc += (double)fabsf(a[i] - b[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kL1Threads / 32; w++) t += s_w[w];
        part[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) l1_finish_kernel(const double* __restrict__ part, int blocks, double inv_n, float* out)
{
    __shared__ double s[256];
    double t = 0.0;
    for (int i = threadIdx.x; i < blocks; i += 256) t += part[i];
    s[threadIdx.x] = t;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (float)(s[0] * inv_n);
}

// dL/da = sgn(a - b) * (g / n)   (torch: abs backward = grad * sgn(x), sgn(0) = 0; mean backward = grad / n, which
// torch's division-by-a-scalar kernel evaluates as grad * (1.0f / n): reproduced, the gradient is bit-identical)
__global__ void __launch_bounds__(kL1Threads) l1_bwd_kernel(size_t n, const float* __restrict__ a, const float* __restrict__ b,
                                                             const float* __restrict__ g, float n_f, float* __restrict__ da)
{
    const float s = __fmul_rn(__ldg(g), __fdiv_rn(1.0f, n_f));
    const size_t n4 = n / 4;
    const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(da)) & 15) == 0;
    const size_t stride = (size_t)gridDim.x * kL1Threads;
    auto sg = [s](float d) { return d > 0.f ? s : (d < 0.f ? -s : (d == 0.f ? 0.f * s : d * s)); };   // NaN propagates
    if (vec) {
        for (size_t i = (size_t)blockIdx.x * kL1Threads + threadIdx.x; i < n4; i += stride) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(a) + i), y = __ldg(reinterpret_cast<const float4*>(b) + i);
            reinterpret_cast<float4*>(da)[i] = make_float4(sg(x.x - y.x), sg(x.y - y.y), sg(x.z - y.z), sg(x.w - y.w));
        }
        for (size_t i = n4 * 4 + (size_t)blockIdx.x * kL1Threads + threadIdx.x; i < n; i += stride) da[i] = sg(a[i] - b[i]);
    } else {
        for (size_t i = (size_t)blockIdx.x * kL1Threads + threadIdx.x; i < n; i += stride) da[i] = sg(a[i] - b[i]);
    }
}

int l1_blocks()
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms * 8;
}

}  // namespace

size_t l1_scratch_bytes() { return (size_t)l1_blocks() * sizeof(double) + 256; }

cudaError_t launch_l1_forward(size_t n, const float* a, const float* b, char* scratch, float* out, cudaStream_t s)
{
    double* part = reinterpret_cast<double*>(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
    const int blocks = l1_blocks();
    l1_fwd_kernel<<<blocks, kL1Threads, 0, s>>>(n, a, b, part);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    l1_finish_kernel<<<1, 256, 0, s>>>(part, blocks, 1.0 / (double)n, out);
    return cudaGetLastError();
}

cudaError_t launch_l1_backward(size_t n, const float* a, const float* b, const float* g, float* da, cudaStream_t s)
{
    l1_bwd_kernel<<<l1_blocks(), kL1Threads, 0, s>>>(n, a, b, g, (float)n, da);
    return cudaGetLastError();
}
