/*
 * cpu_raster.c - CPU ORACLE for the differentiable 4D-Gaussian rasterizer.   TEST INFRASTRUCTURE.
 *
 * A plain-C restatement of the algorithm of the reference CUDA extension
 * submodules/diff_gaussian_rasterization_df (juno181/Ex4DGS), function by function, so that the
 * CUDA product can be checked against an independent implementation on a machine without a
 * GPU.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this file's shared object; the product (ex4dgs_b200) never does.
 *
 * Parity status: the reference ships no tests or golden vectors; this oracle is
 * pinned instead against outputs of the compiled, unmodified reference itself, generated on a
 * B200 by oracle/make_golden.py and committed under tests/golden/ (see tests/test_oracle.py).
 *
 * Arithmetic: compiled with -ffp-contract=off.  Where the result decides an integer output
 * (depth key, radius, tile rectangle, alpha thresholds) the float operations are written out with
 * explicit fmaf() in the places where nvcc contracts the reference's expressions (read from the
 * SASS of the reference build; see DESIGN.md "pinned arithmetic").  expf/sqrt come from libm
 * (CUDA's expf differs by <= 2 ulp; no integer output depends on it except through the
 * alpha/transmittance thresholds, where a flip needs a value within 1 ulp of the threshold).
 *
 * Reference map (file:line under submodules/diff_gaussian_rasterization_df/cuda_rasterizer/):
 *   or_preprocess      forward.cu:165-269 (+ :20-71 SH, :74-124 cov2D/mip, :128-162 cov3D; auxiliary.h:41-56,68-87,267-294)
 *   or_bin             rasterizer_impl.cu:72-113 (keys), :293-299 (scan), :318-336 (sort over bits [0,32+bit), ranges)
 *   or_render          forward.cu:274-462
 *   or_render_bwd      backward.cu:426-682
 *   or_cov2d_bwd       backward.cu:144-300
 *   or_preprocess_bwd  backward.cu:372-423 (+ :20-139 SH, :304-367 cov3D)
 *   stage order        rasterizer_impl.cu:204-363, :367-486; output fills rasterize_points.cu:73-78,178-187
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TILE 16
#define NCH 3

typedef struct {
    int P, D, M, W, H, gx, gy;
    float tanx, tany, fx, fy, ksize, smod, dmin, dmax;
    const float *bg, *means, *dir, *shs, *colpre, *opac, *scales, *rots, *covpre, *view, *proj, *cam, *subpix;
    /* geometry state (rasterizer_impl.h:31-45) */
    float *depths, *means2D, *cov3D, *conic_op, *rgb;
    uint8_t *clamped;
    int *radii;
    uint32_t *tiles_touched, *point_offsets;
    /* binning state */
    uint32_t R;
    uint64_t *keys;
    uint32_t *plist;
    /* image state */
    uint32_t *ranges;      /* [tiles][2] */
    float *final_T;
    uint32_t *n_contrib;
    float *out_depth, *out_acc; /* kept for the backward */
} Ctx;

static const float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
static const float C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f, 0.5462742152960396f};
static const float C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                            -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

/* a0*b0 + a1*b1 + a2*b2 with the contraction of the reference build */
static inline float sum3(float a0, float b0, float a1, float b1, float a2, float b2)
{
    return fmaf(a2, b2, fmaf(a0, b0, a1 * b1));
}
/* row c of a transposed-storage 4x4 applied to (x,y,z,1): auxiliary.h:68-87 */
static inline float xrow(const float* m, int c, float x, float y, float z)
{
    return sum3(x, m[c], y, m[c + 4], z, m[c + 8]) + m[c + 12];
}

void* or_create(void) { return calloc(1, sizeof(Ctx)); }

static void free_state(Ctx* c)
{
    free(c->depths); free(c->means2D); free(c->cov3D); free(c->conic_op); free(c->rgb); free(c->clamped);
    free(c->radii); free(c->tiles_touched); free(c->point_offsets); free(c->keys); free(c->plist);
    free(c->ranges); free(c->final_T); free(c->n_contrib); free(c->out_depth); free(c->out_acc);
    c->depths = c->means2D = c->cov3D = c->conic_op = c->rgb = NULL; c->clamped = NULL; c->radii = NULL;
    c->tiles_touched = c->point_offsets = NULL; c->keys = NULL; c->plist = NULL; c->ranges = NULL;
    c->final_T = NULL; c->n_contrib = NULL; c->out_depth = c->out_acc = NULL;
}
void or_destroy(void* h) { if (h) { free_state((Ctx*)h); free(h); } }

/* forward.cu:128-162 - Sigma from scale and UN-normalised quaternion */
static void cov3d(const float* s3, float mod, const float* q, float* c)
{
    const float sx = mod * s3[0], sy = mod * s3[1], sz = mod * s3[2];
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    const float xz = x * z, rx = r * x, rz = r * z, yy = y * y, zz = z * z;
    float t, R[3][3];                     /* R[c][r] = glm column c, row r */
    t = yy + zz;           R[0][0] = 1.f - (t + t);
    t = fmaf(x, x, zz);    R[1][1] = 1.f - (t + t);
    t = fmaf(x, x, yy);    R[2][2] = 1.f - (t + t);
    t = fmaf(x, y, -rz);   R[0][1] = t + t;
    t = fmaf(r, y, xz);    R[0][2] = t + t;
    t = fmaf(x, y, rz);    R[1][0] = t + t;
    t = fmaf(y, z, -rx);   R[1][2] = t + t;
    t = fmaf(-r, y, xz);   R[2][0] = t + t;
    t = fmaf(y, z, rx);    R[2][1] = t + t;
    float M[3][3];
    for (int cc = 0; cc < 3; cc++) { M[cc][0] = sx * R[cc][0]; M[cc][1] = sy * R[cc][1]; M[cc][2] = sz * R[cc][2]; }
    int k = 0;
    for (int a = 0; a < 3; a++)
        for (int b = a; b < 3; b++)
            c[k++] = sum3(M[a][0], M[b][0], M[a][1], M[b][1], M[a][2], M[b][2]);
}

typedef struct { float a, b, c, T0[3], T1[3], tz, txtz, tytz; } Cov2;

/* forward.cu:74-106 (projection part, shared with backward.cu:167-199) */
static Cov2 cov2d(const float* m, const Ctx* c, const float* cv)
{
    Cov2 o;
    const float* V = c->view;
    float tx = xrow(V, 0, m[0], m[1], m[2]), ty = xrow(V, 1, m[0], m[1], m[2]);
    const float tz = xrow(V, 2, m[0], m[1], m[2]);
    const float limx = 1.3f * c->tanx, limy = 1.3f * c->tany;
    o.txtz = tx / tz; o.tytz = ty / tz; o.tz = tz;
    tx = fminf(limx, fmaxf(-limx, o.txtz)) * tz;
    ty = fminf(limy, fmaxf(-limy, o.tytz)) * tz;
    const float tz2 = tz * tz;
    const float J00 = c->fx / tz, J02 = -(c->fx * tx) / tz2, J11 = c->fy / tz, J12 = -(c->fy * ty) / tz2;
    for (int r = 0; r < 3; r++) {
        o.T0[r] = fmaf(J02, V[4 * r + 2], V[4 * r] * J00);
        o.T1[r] = fmaf(J12, V[4 * r + 2], V[4 * r + 1] * J11);
    }
    const float X00 = sum3(o.T0[0], cv[0], o.T0[1], cv[1], o.T0[2], cv[2]);
    const float X10 = sum3(o.T0[0], cv[1], o.T0[1], cv[3], o.T0[2], cv[4]);
    const float X20 = sum3(o.T0[0], cv[2], o.T0[1], cv[4], o.T0[2], cv[5]);
    const float X01 = sum3(o.T1[0], cv[0], o.T1[1], cv[1], o.T1[2], cv[2]);
    const float X11 = sum3(o.T1[0], cv[1], o.T1[1], cv[3], o.T1[2], cv[4]);
    const float X21 = sum3(o.T1[0], cv[2], o.T1[1], cv[4], o.T1[2], cv[5]);
    o.a = sum3(o.T0[0], X00, o.T0[1], X10, o.T0[2], X20);
    o.b = sum3(o.T0[0], X01, o.T0[1], X11, o.T0[2], X21);
    o.c = sum3(o.T1[0], X01, o.T1[1], X11, o.T1[2], X21);
    return o;
}

/* auxiliary.h:46-56 */
static void get_rect(float px, float py, int radius, int gx, int gy, int* x0, int* y0, int* x1, int* y1)
{
    const float r = (float)radius;
    int v;
    v = (int)((px - r) * 0.0625f);               *x0 = v < 0 ? 0 : (v > gx ? gx : v);
    v = (int)((py - r) * 0.0625f);               *y0 = v < 0 ? 0 : (v > gy ? gy : v);
    v = (int)((((px + r) + 16.f) - 1.f) * 0.0625f); *x1 = v < 0 ? 0 : (v > gx ? gx : v);
    v = (int)((((py + r) + 16.f) - 1.f) * 0.0625f); *y1 = v < 0 ? 0 : (v > gy ? gy : v);
}

/* SH basis for direction (x,y,z): forward.cu:30-59 */
static int sh_basis(int D, float x, float y, float z, float* b)
{
    b[0] = C0;
    if (D < 1) return 1;
    b[1] = -C1 * y; b[2] = C1 * z; b[3] = -C1 * x;
    if (D < 2) return 4;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = C2[0] * xy; b[5] = C2[1] * yz; b[6] = C2[2] * (2.0f * zz - xx - yy); b[7] = C2[3] * xz; b[8] = C2[4] * (xx - yy);
    if (D < 3) return 9;
    b[9] = C3[0] * y * (3.0f * xx - yy);
    b[10] = C3[1] * xy * z;
    b[11] = C3[2] * y * (4.0f * zz - xx - yy);
    b[12] = C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
    b[13] = C3[4] * x * (4.0f * zz - xx - yy);
    b[14] = C3[5] * z * (xx - yy);
    b[15] = C3[6] * x * (xx - 3.0f * yy);
    return 16;
}

static void or_preprocess(Ctx* c)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < c->P; i++) {
        c->radii[i] = 0; c->tiles_touched[i] = 0;                        /* forward.cu:201-202 */
        const float* m = c->means + 3 * i;
        /* in_frustum, auxiliary.h:267-294 */
        const float hx = xrow(c->proj, 0, m[0], m[1], m[2]), hy = xrow(c->proj, 1, m[0], m[1], m[2]);
        const float hw = xrow(c->proj, 3, m[0], m[1], m[2]);
        const float pw = 1.0f / (hw + 0.0000001f);
        const float nx = hx * pw, ny = hy * pw;
        const float depth = xrow(c->view, 2, m[0], m[1], m[2]);
        if (depth <= c->dmin || depth > c->dmax || (double)nx < -1.3 || (double)nx > 1.3 || (double)ny < -1.3 || (double)ny > 1.3)
            continue;
        float* cv = c->cov3D + 6 * i;
        if (c->covpre) memcpy(cv, c->covpre + 6 * i, 6 * sizeof(float));
        else cov3d(c->scales + 3 * i, c->smod, c->rots + 4 * i, cv);
        const Cov2 o = cov2d(m, c, cv);
        /* mip filter coefficient, forward.cu:112-121 (double sub-expressions) */
        const float bb = o.b * o.b;
        const float det0f = fmaf(o.a, o.c, -bb);
        const float ak = o.a + c->ksize, ck = o.c + c->ksize;
        const float det = fmaf(ak, ck, -bb);
        const float det_0 = (float)fmax(1e-6, (double)det0f);
        const float det_1 = (float)fmax(1e-6, (double)det);
        float coef = (float)sqrt((double)det_0 / ((double)det_1 + 1e-6) + 1e-6);
        if ((double)det_0 <= 1e-6 || (double)det_1 <= 1e-6) coef = 0.0f;
        if (det == 0.0f) continue;                                         /* forward.cu:233 */
        const float dinv = 1.f / det;
        const float conA = ck * dinv, conB = -o.b * dinv, conC = ak * dinv;
        const float mid = 0.5f * (ak + ck);
        const float disc = sqrtf(fmaxf(0.1f, fmaf(mid, mid, -det)));
        const float lam = fmaxf(mid + disc, mid - disc);
        const int radius = (int)ceilf(3.f * sqrtf(lam));
        const float px = (float)((((double)nx + 1.0) * (double)c->W - 1.0) * 0.5);   /* auxiliary.h:41-44 */
        const float py = (float)((((double)ny + 1.0) * (double)c->H - 1.0) * 0.5);
        int x0, y0, x1, y1;
        get_rect(px, py, radius, c->gx, c->gy, &x0, &y0, &x1, &y1);
        if ((x1 - x0) * (y1 - y0) == 0) continue;
        if (!c->colpre) {                                                  /* forward.cu:20-71 */
            float dx = m[0] - c->cam[0], dy = m[1] - c->cam[1], dz = m[2] - c->cam[2];
            const float len = sqrtf(sum3(dx, dx, dy, dy, dz, dz));
            dx /= len; dy /= len; dz /= len;
            float b[16];
            const int nb = sh_basis(c->D, dx, dy, dz, b);
            const float* sh = c->shs + (size_t)i * c->M * 3;
            for (int ch = 0; ch < 3; ch++) {
                float v = 0.f;
                for (int k = 0; k < nb; k++) v = fmaf(b[k], sh[3 * k + ch], v);
                v += 0.5f;
                c->clamped[3 * i + ch] = v < 0.f;
                c->rgb[3 * i + ch] = v < 0.f ? 0.f : v;
            }
        }
        c->depths[i] = depth;
        c->radii[i] = radius;
        c->means2D[2 * i] = px; c->means2D[2 * i + 1] = py;
        c->conic_op[4 * i] = conA; c->conic_op[4 * i + 1] = conB; c->conic_op[4 * i + 2] = conC;
        c->conic_op[4 * i + 3] = c->opac[i] * coef;
        c->tiles_touched[i] = (uint32_t)((y1 - y0) * (x1 - x0));
    }
}

/* rasterizer_impl.cu:35-50 */
static uint32_t higher_msb(uint32_t n)
{
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) { step /= 2; if (n >> msb) msb += step; else msb -= step; }
    if (n >> msb) msb++;
    return msb;
}

static int or_bin(Ctx* c)
{
    uint32_t run = 0;                                                       /* rasterizer_impl.cu:295 */
    for (int i = 0; i < c->P; i++) { run += c->tiles_touched[i]; c->point_offsets[i] = run; }
    c->R = c->P ? c->point_offsets[c->P - 1] : 0;
    const uint32_t R = c->R;
    const int tiles = c->gx * c->gy;
    memset(c->ranges, 0, sizeof(uint32_t) * 2 * (size_t)tiles);
    if (!R) return 0;
    uint64_t* k0 = (uint64_t*)malloc(sizeof(uint64_t) * R), *k1 = (uint64_t*)malloc(sizeof(uint64_t) * R);
    uint32_t* v0 = (uint32_t*)malloc(sizeof(uint32_t) * R), *v1 = (uint32_t*)malloc(sizeof(uint32_t) * R);
    if (!k0 || !k1 || !v0 || !v1) return -1;
    for (int i = 0; i < c->P; i++) {                                        /* rasterizer_impl.cu:72-113 */
        if (c->radii[i] <= 0) continue;
        uint32_t off = i == 0 ? 0 : c->point_offsets[i - 1];
        int x0, y0, x1, y1;
        get_rect(c->means2D[2 * i], c->means2D[2 * i + 1], c->radii[i], c->gx, c->gy, &x0, &y0, &x1, &y1);
        uint32_t dbits; memcpy(&dbits, &c->depths[i], 4);
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++) {
                k0[off] = ((uint64_t)(uint32_t)(y * c->gx + x) << 32) | dbits;
                v0[off] = (uint32_t)i;
                off++;
            }
    }
    /* stable LSD radix sort on key bits [0, 32+bit) (rasterizer_impl.cu:318-326) */
    const int end_bit = 32 + (int)higher_msb((uint32_t)tiles);
    for (int shift = 0; shift < end_bit; shift += 8) {
        const int nb = end_bit - shift < 8 ? end_bit - shift : 8;
        const uint32_t mask = (1u << nb) - 1;
        size_t cnt[257] = {0};
        for (uint32_t i = 0; i < R; i++) cnt[((k0[i] >> shift) & mask) + 1]++;
        for (int b = 0; b < 256; b++) cnt[b + 1] += cnt[b];
        for (uint32_t i = 0; i < R; i++) { const size_t d = cnt[(k0[i] >> shift) & mask]++; k1[d] = k0[i]; v1[d] = v0[i]; }
        uint64_t* tk = k0; k0 = k1; k1 = tk;
        uint32_t* tv = v0; v0 = v1; v1 = tv;
    }
    c->keys = k0; c->plist = v0; free(k1); free(v1);
    for (uint32_t i = 0; i < R; i++) {                                      /* rasterizer_impl.cu:118-140 */
        const uint32_t cur = (uint32_t)(k0[i] >> 32);
        if (i == 0) c->ranges[2 * cur] = 0;
        else { const uint32_t prev = (uint32_t)(k0[i - 1] >> 32); if (cur != prev) { c->ranges[2 * prev + 1] = i; c->ranges[2 * cur] = i; } }
        if (i == R - 1) c->ranges[2 * cur + 1] = R;
    }
    return 0;
}

/* the alpha of one (pixel, splat) pair; returns 0 when the pair is skipped (forward.cu:368-381) */
static inline int pair_alpha(const float* xy, const float* co, float pxf, float pyf, float* dx, float* dy, float* G, float* alpha)
{
    *dx = xy[0] - pxf; *dy = xy[1] - pyf;
    const float power = fmaf(fmaf(*dx, *dx * co[0], (co[2] * *dy) * *dy), -0.5f, -((co[1] * *dx) * *dy));
    if (power > 0.0f) return 0;
    *G = expf(power);
    *alpha = fminf(0.99f, co[3] * *G);
    if (*alpha < 1.0f / 255.0f) return 0;
    return 1;
}

static void or_render(Ctx* c, float* out_color, float* out_depth, float* out_acc, float* out_flow, int* out_idx)
{
    const int W = c->W, H = c->H;
    const size_t HW = (size_t)W * H;
    const float* feat = c->colpre ? c->colpre : c->rgb;                    /* rasterizer_impl.cu:339 */
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < c->gx * c->gy; tile++) {
        const int tx = tile % c->gx, ty = tile / c->gx;
        const uint32_t r0 = c->ranges[2 * tile], r1 = c->ranges[2 * tile + 1];
        for (int ly = 0; ly < TILE; ly++)
            for (int lx = 0; lx < TILE; lx++) {
                const int x = tx * TILE + lx, y = ty * TILE + ly;
                if (x >= W || y >= H) continue;
                const size_t pix = (size_t)W * y + x;
                const float pxf = (float)x + c->subpix[2 * pix], pyf = (float)y + c->subpix[2 * pix + 1];
                float T = 1.f, C[3] = {0, 0, 0}, Dp = 0.f, acc = 0.f, F[3] = {0, 0, 0}, max_vis = 0.f;
                uint32_t contributor = 0, last = 0;
                int best = -1;
                /* Note: the CUDA kernel stops a whole tile when all 256 pixels are done; per pixel
                 * that is the same as stopping at the pixel's own `done` (forward.cu:341-387). */
                for (uint32_t q = r0; q < r1; q++) {
                    const uint32_t id = c->plist[q];
                    contributor++;
                    float dx, dy, G, alpha;
                    if (!pair_alpha(c->means2D + 2 * id, c->conic_op + 4 * id, pxf, pyf, &dx, &dy, &G, &alpha)) continue;
                    const float test_T = T * (1.f - alpha);
                    if (test_T < 0.0001f) break;
                    for (int ch = 0; ch < 3; ch++) C[ch] = fmaf(T, alpha * feat[3 * id + ch], C[ch]);
                    Dp = fmaf(T, alpha * c->depths[id], Dp);
                    const float w = T * alpha;
                    acc += w;
                    for (int ch = 0; ch < 3; ch++) F[ch] = fmaf(T, alpha * c->dir[3 * id + ch], F[ch]);
                    if (w > max_vis) { max_vis = w; best = (int)id; }
                    T = test_T;
                    last = contributor;
                }
                if (acc == 0.0f) Dp = fmaf(1.0f - acc, c->dmax, Dp);          /* forward.cu:428-446 */
                else { Dp /= acc; for (int ch = 0; ch < 3; ch++) F[ch] /= acc; }
                c->final_T[pix] = T; c->n_contrib[pix] = last;
                for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pix] = fmaf(c->bg[ch], T, C[ch]);
                out_depth[pix] = Dp; out_acc[pix] = acc;
                for (int ch = 0; ch < 3; ch++) out_flow[ch * HW + pix] = F[ch];
                out_idx[pix] = best;
                c->out_depth[pix] = Dp; c->out_acc[pix] = acc;
            }
    }
}

/* Returns num_rendered (R) or -1.  Outputs as rasterize_points.cu:73-78. */
int or_forward(void* h, int P, int D, int M, const float* bg, int W, int H,
               const float* means, const float* dir, const float* shs, const float* colpre, const float* opac,
               const float* scales, float smod, const float* rots, const float* covpre,
               const float* view, const float* proj, const float* cam, float tanx, float tany, float ksize,
               const float* subpix, float dmin, float dmax,
               float* out_color, float* out_depth, float* out_acc, float* out_flow, int* out_idx, int* radii)
{
    Ctx* c = (Ctx*)h;
    free_state(c);
    c->P = P; c->D = D; c->M = M; c->W = W; c->H = H; c->gx = (W + TILE - 1) / TILE; c->gy = (H + TILE - 1) / TILE;
    c->tanx = tanx; c->tany = tany; c->fx = W / (2.0f * tanx); c->fy = H / (2.0f * tany);
    c->ksize = ksize; c->smod = smod; c->dmin = dmin; c->dmax = dmax;
    c->bg = bg; c->means = means; c->dir = dir; c->shs = shs; c->colpre = colpre; c->opac = opac; c->scales = scales;
    c->rots = rots; c->covpre = covpre; c->view = view; c->proj = proj; c->cam = cam; c->subpix = subpix;
    const size_t HW = (size_t)W * H, n = P > 0 ? (size_t)P : 1;
    for (size_t i = 0; i < 3 * HW; i++) { out_color[i] = 0.f; out_flow[i] = 0.f; }
    for (size_t i = 0; i < HW; i++) { out_depth[i] = 0.f; out_acc[i] = 0.f; out_idx[i] = -1; }
    if (P == 0) return 0;                                                    /* rasterize_points.cu:89-90 */
    c->depths = (float*)calloc(n, 4); c->means2D = (float*)calloc(n, 8); c->cov3D = (float*)calloc(n, 24);
    c->conic_op = (float*)calloc(n, 16); c->rgb = (float*)calloc(n, 12); c->clamped = (uint8_t*)calloc(n, 3);
    c->radii = (int*)calloc(n, 4); c->tiles_touched = (uint32_t*)calloc(n, 4); c->point_offsets = (uint32_t*)calloc(n, 4);
    c->ranges = (uint32_t*)calloc((size_t)c->gx * c->gy, 8);
    c->final_T = (float*)calloc(HW, 4); c->n_contrib = (uint32_t*)calloc(HW, 4);
    c->out_depth = (float*)calloc(HW, 4); c->out_acc = (float*)calloc(HW, 4);
    or_preprocess(c);
    if (or_bin(c) != 0) return -1;
    or_render(c, out_color, out_depth, out_acc, out_flow, out_idx);
    memcpy(radii, c->radii, sizeof(int) * (size_t)P);
    return (int)c->R;
}

/* copies of the intermediates for the parity tests; any pointer may be NULL */
void or_get_state(void* h, float* depths, float* means2D, float* cov3D, float* conic_op, float* rgb, uint8_t* clamped,
                  uint32_t* tiles_touched, uint32_t* point_offsets, uint64_t* keys, uint32_t* plist,
                  uint32_t* ranges, float* final_T, uint32_t* n_contrib)
{
    Ctx* c = (Ctx*)h;
    const size_t P = (size_t)c->P, HW = (size_t)c->W * c->H, tiles = (size_t)c->gx * c->gy;
    if (depths) memcpy(depths, c->depths, 4 * P);
    if (means2D) memcpy(means2D, c->means2D, 8 * P);
    if (cov3D) memcpy(cov3D, c->cov3D, 24 * P);
    if (conic_op) memcpy(conic_op, c->conic_op, 16 * P);
    if (rgb) memcpy(rgb, c->rgb, 12 * P);
    if (clamped) memcpy(clamped, c->clamped, 3 * P);
    if (tiles_touched) memcpy(tiles_touched, c->tiles_touched, 4 * P);
    if (point_offsets) memcpy(point_offsets, c->point_offsets, 4 * P);
    if (keys && c->R) memcpy(keys, c->keys, 8 * (size_t)c->R);
    if (plist && c->R) memcpy(plist, c->plist, 4 * (size_t)c->R);
    if (ranges) memcpy(ranges, c->ranges, 8 * tiles);
    if (final_T) memcpy(final_T, c->final_T, 4 * HW);
    if (n_contrib) memcpy(n_contrib, c->n_contrib, 4 * HW);
}

/* The reference adds the per-pixel terms with float atomicAdd in whatever order the hardware schedules them
 * (backward.cu:604-672): every summation order is "the reference".  The oracle accumulates them in double and rounds
 * once at the end - the value all those orders scatter around - which also makes it reproducible from run to run
 * although its tiles are processed by OpenMP threads in arbitrary order. */
static inline void atomic_addf(double* p, float v)
{
#pragma omp atomic
    *p += (double)v;
}

/* backward.cu:426-682 */
static void or_render_bwd(Ctx* c, const float* dpix, const float* ddepth, const float* dflow, const float* dacc,
                          double* dmean2D /*[P,3]*/, double* dconic /*[P,4]*/, double* ddir, double* dopac, double* dcol)
{
    const int W = c->W, H = c->H;
    const size_t HW = (size_t)W * H;
    const float* feat = c->colpre ? c->colpre : c->rgb;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < c->gx * c->gy; tile++) {
        const int tx = tile % c->gx, ty = tile / c->gx;
        const uint32_t r0 = c->ranges[2 * tile], r1 = c->ranges[2 * tile + 1];
        for (int ly = 0; ly < TILE; ly++)
            for (int lx = 0; lx < TILE; lx++) {
                const int x = tx * TILE + lx, y = ty * TILE + ly;
                if (x >= W || y >= H) continue;
                const size_t pix = (size_t)W * y + x;
                const float pxf = (float)x + c->subpix[2 * pix], pyf = (float)y + c->subpix[2 * pix + 1];
                const float T_final = c->final_T[pix];
                float T = T_final;
                const uint32_t last = c->n_contrib[pix];
                const float acc = c->out_acc[pix], final_depth = c->out_depth[pix];
                float g_depth = ddepth[pix], g_flow[3] = {0, 0, 0}, g_acc = 0.f;
                if (acc > 0.0f) {
                    g_depth /= acc;
                    for (int ch = 0; ch < 3; ch++) g_flow[ch] = dflow[ch * HW + pix] / acc;
                    g_acc = dacc[pix];
                }
                float g_pix[3], accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0.f;
                for (int ch = 0; ch < 3; ch++) g_pix[ch] = dpix[ch * HW + pix];
                float bg_dot = 0.f;
                for (int ch = 0; ch < 3; ch++) bg_dot += c->bg[ch] * g_pix[ch];
                /* walk the list back to front; entries at list position >= n_contrib are skipped (backward.cu:575-577) */
                for (uint32_t q = r1; q-- > r0;) {
                    const uint32_t pos = q - r0;
                    if (pos >= last) continue;
                    const uint32_t id = c->plist[q];
                    float dx, dy, G, alpha;
                    const float* co = c->conic_op + 4 * id;
                    if (!pair_alpha(c->means2D + 2 * id, co, pxf, pyf, &dx, &dy, &G, &alpha)) continue;
                    T = T / (1.f - alpha);
                    const float w = alpha * T;
                    float dL_dalpha = 0.f;
                    const float dep = c->depths[id];
                    if ((dep > c->dmin) & (w > 0.0f)) {                      /* backward.cu:604-622 (A.3-Q3) */
                        atomic_addf(&dmean2D[3 * id + 2], g_depth * w);
                        dL_dalpha += (final_depth - dep) * g_depth * T;
                    }
                    for (int ch = 0; ch < 3; ch++) {
                        const float col = feat[3 * id + ch];
                        accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                        last_color[ch] = col;
                        dL_dalpha += (col - accum_rec[ch]) * g_pix[ch];
                        atomic_addf(&dcol[3 * id + ch], w * g_pix[ch]);
                    }
                    for (int ch = 0; ch < 3; ch++) atomic_addf(&ddir[3 * id + ch], w * g_flow[ch]);   /* A.3-Q5 */
                    dL_dalpha *= T;
                    g_acc *= T;                                              /* cumulative, A.3-Q4 */
                    last_alpha = alpha;
                    dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                    const float dL_dG = co[3] * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * co[0] - gdy * co[1];
                    const float dG_ddely = -gdy * co[2] - gdx * co[1];
                    atomic_addf(&dmean2D[3 * id], dL_dG * dG_ddelx * ddelx_dx);
                    atomic_addf(&dmean2D[3 * id + 1], dL_dG * dG_ddely * ddely_dy);
                    atomic_addf(&dconic[4 * id], -0.5f * gdx * dx * dL_dG);
                    atomic_addf(&dconic[4 * id + 1], -0.5f * gdx * dy * dL_dG);
                    atomic_addf(&dconic[4 * id + 3], -0.5f * gdy * dy * dL_dG);
                    atomic_addf(&dopac[id], G * dL_dalpha);
                    atomic_addf(&dopac[id], G * g_acc);
                }
            }
    }
}

/* backward.cu:144-257 (cov2D -> cov3D; the mean term of :259-299 is overwritten later, A.3-Q2) and
 * backward.cu:372-423 with :20-139 (SH) and :304-367 (scale / rotation) */
static void or_preprocess_bwd(Ctx* c, const float* dmean2D, const float* dconic, const float* dcol,
                              float* dmean3D, float* dcov3D, float* dsh, float* dscale, float* drot)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < c->P; i++) {
        if (!(c->radii[i] > 0)) continue;
        const float* m = c->means + 3 * i;
        const float* cv = c->covpre ? c->covpre + 6 * i : c->cov3D + 6 * i;
        float* dc = dcov3D + 6 * i;
        {
            const Cov2 o = cov2d(m, c, cv);
            const float a = o.a + c->ksize, b = o.b, cc = o.c + c->ksize;
            const float gx = dconic[4 * i], gy = dconic[4 * i + 1], gz = dconic[4 * i + 3];
            const float denom = a * cc - b * b;
            const float d2 = 1.0f / ((denom * denom) + 0.0000001f);
            if (d2 != 0) {
                const float da = d2 * (-cc * cc * gx + 2 * b * cc * gy + (denom - a * cc) * gz);
                const float dcc = d2 * (-a * a * gz + 2 * a * b * gy + (denom - a * cc) * gx);
                const float db = d2 * 2 * (b * cc * gx - (denom + 2 * b * b) * gy + a * b * gz);
                const float* T0 = o.T0; const float* T1 = o.T1;
                dc[0] = (T0[0] * T0[0] * da + T0[0] * T1[0] * db + T1[0] * T1[0] * dcc);
                dc[3] = (T0[1] * T0[1] * da + T0[1] * T1[1] * db + T1[1] * T1[1] * dcc);
                dc[5] = (T0[2] * T0[2] * da + T0[2] * T1[2] * db + T1[2] * T1[2] * dcc);
                dc[1] = 2 * T0[0] * T0[1] * da + (T0[0] * T1[1] + T0[1] * T1[0]) * db + 2 * T1[0] * T1[1] * dcc;
                dc[2] = 2 * T0[0] * T0[2] * da + (T0[0] * T1[2] + T0[2] * T1[0]) * db + 2 * T1[0] * T1[2] * dcc;
                dc[4] = 2 * T0[2] * T0[1] * da + (T0[1] * T1[2] + T0[2] * T1[1]) * db + 2 * T1[1] * T1[2] * dcc;
            } else {
                for (int k = 0; k < 6; k++) dc[k] = 0;
            }
        }
        float dm[3];
        {                                                                    /* backward.cu:396-414 */
            const float* pr = c->proj;
            const float hw = pr[3] * m[0] + pr[7] * m[1] + pr[11] * m[2] + pr[15];
            const float mw = 1.0f / (hw + 0.0000001f);
            const float mul1 = (pr[0] * m[0] + pr[4] * m[1] + pr[8] * m[2] + pr[12]) * mw * mw;
            const float mul2 = (pr[1] * m[0] + pr[5] * m[1] + pr[9] * m[2] + pr[13]) * mw * mw;
            const float mul3 = (pr[2] * m[0] + pr[6] * m[1] + pr[10] * m[2] + pr[14]) * mw * mw;
            const float gx = dmean2D[3 * i], gy = dmean2D[3 * i + 1], gz = dmean2D[3 * i + 2];
            for (int k = 0; k < 3; k++)
                dm[k] = (pr[4 * k] * mw - pr[4 * k + 3] * mul1) * gx + (pr[4 * k + 1] * mw - pr[4 * k + 3] * mul2) * gy +
                        (pr[4 * k + 2] * mw - pr[4 * k + 3] * mul3) * gz;
        }
        if (c->shs) {                                                        /* backward.cu:20-139 */
            const float ox = m[0] - c->cam[0], oy = m[1] - c->cam[1], oz = m[2] - c->cam[2];
            const float len = sqrtf(ox * ox + oy * oy + oz * oz);
            const float x = ox / len, y = oy / len, z = oz / len;
            float g[3];
            for (int ch = 0; ch < 3; ch++) g[ch] = c->clamped[3 * i + ch] ? 0.f : dcol[3 * i + ch];
            float b[16];
            const int nb = sh_basis(c->D, x, y, z, b);
            const float* sh = c->shs + (size_t)i * c->M * 3;
            float* out = dsh + (size_t)i * c->M * 3;
            for (int k = 0; k < nb; k++)
                for (int ch = 0; ch < 3; ch++) out[3 * k + ch] = b[k] * g[ch];
            /* d(colour)/d(direction), contracted with g */
            float dRx[3] = {0, 0, 0}, dRy[3] = {0, 0, 0}, dRz[3] = {0, 0, 0};
            if (c->D > 0) {
                for (int ch = 0; ch < 3; ch++) {
                    const float* s = sh + ch;
#define SH(k) s[3 * (k)]
                    dRx[ch] = -C1 * SH(3); dRy[ch] = -C1 * SH(1); dRz[ch] = C1 * SH(2);
                    if (c->D > 1) {
                        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                        dRx[ch] += C2[0] * y * SH(4) + C2[2] * 2.f * -x * SH(6) + C2[3] * z * SH(7) + C2[4] * 2.f * x * SH(8);
                        dRy[ch] += C2[0] * x * SH(4) + C2[1] * z * SH(5) + C2[2] * 2.f * -y * SH(6) + C2[4] * 2.f * -y * SH(8);
                        dRz[ch] += C2[1] * y * SH(5) + C2[2] * 2.f * 2.f * z * SH(6) + C2[3] * x * SH(7);
                        if (c->D > 2) {
                            dRx[ch] += (C3[0] * SH(9) * 3.f * 2.f * xy + C3[1] * SH(10) * yz + C3[2] * SH(11) * -2.f * xy +
                                        C3[3] * SH(12) * -3.f * 2.f * xz + C3[4] * SH(13) * (-3.f * xx + 4.f * zz - yy) +
                                        C3[5] * SH(14) * 2.f * xz + C3[6] * SH(15) * 3.f * (xx - yy));
                            dRy[ch] += (C3[0] * SH(9) * 3.f * (xx - yy) + C3[1] * SH(10) * xz + C3[2] * SH(11) * (-3.f * yy + 4.f * zz - xx) +
                                        C3[3] * SH(12) * -3.f * 2.f * yz + C3[4] * SH(13) * -2.f * xy + C3[5] * SH(14) * -2.f * yz +
                                        C3[6] * SH(15) * -3.f * 2.f * xy);
                            dRz[ch] += (C3[1] * SH(10) * xy + C3[2] * SH(11) * 4.f * 2.f * yz + C3[3] * SH(12) * 3.f * (2.f * zz - xx - yy) +
                                        C3[4] * SH(13) * 4.f * 2.f * xz + C3[5] * SH(14) * (xx - yy));
                        }
                    }
#undef SH
                }
            }
            const float dd[3] = {dRx[0] * g[0] + dRx[1] * g[1] + dRx[2] * g[2], dRy[0] * g[0] + dRy[1] * g[1] + dRy[2] * g[2],
                                 dRz[0] * g[0] + dRz[1] * g[1] + dRz[2] * g[2]};
            const float sum2 = ox * ox + oy * oy + oz * oz;                  /* auxiliary.h:235-245 */
            const float inv = 1.0f / sqrtf(sum2 * sum2 * sum2);
            dm[0] += ((+sum2 - ox * ox) * dd[0] - oy * ox * dd[1] - oz * ox * dd[2]) * inv;
            dm[1] += (-ox * oy * dd[0] + (sum2 - oy * oy) * dd[1] - oz * oy * dd[2]) * inv;
            dm[2] += (-ox * oz * dd[0] - oy * oz * dd[1] + (sum2 - oz * oz) * dd[2]) * inv;
        }
        for (int k = 0; k < 3; k++) dmean3D[3 * i + k] = dm[k];
        if (c->scales) {                                                     /* backward.cu:304-367 */
            const float* q = c->rots + 4 * i;
            const float r = q[0], x = q[1], y = q[2], z = q[3];
            const float Rm[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                                    {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                                    {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
            const float s[3] = {c->smod * c->scales[3 * i], c->smod * c->scales[3 * i + 1], c->smod * c->scales[3 * i + 2]};
            const float dS[3][3] = {{dc[0], 0.5f * dc[1], 0.5f * dc[2]}, {0.5f * dc[1], dc[3], 0.5f * dc[4]}, {0.5f * dc[2], 0.5f * dc[4], dc[5]}};
            float dM[3][3], Mt[3][3];
            for (int cc = 0; cc < 3; cc++)
                for (int rr = 0; rr < 3; rr++) {
                    float t = 0.f;
                    for (int k = 0; k < 3; k++) t += (2.0f * s[rr] * Rm[k][rr]) * dS[cc][k];
                    dM[cc][rr] = t;
                }
            for (int a = 0; a < 3; a++) {
                float t = 0.f;
                for (int b2 = 0; b2 < 3; b2++) t += Rm[b2][a] * dM[b2][a];
                dscale[3 * i + a] = t;
                for (int b2 = 0; b2 < 3; b2++) Mt[a][b2] = dM[b2][a] * s[a];
            }
            float* o = drot + 4 * i;                                         /* no normalisation Jacobian, A.3-Q6 */
            o[0] = 2 * z * (Mt[0][1] - Mt[1][0]) + 2 * y * (Mt[2][0] - Mt[0][2]) + 2 * x * (Mt[1][2] - Mt[2][1]);
            o[1] = 2 * y * (Mt[1][0] + Mt[0][1]) + 2 * z * (Mt[2][0] + Mt[0][2]) + 2 * r * (Mt[1][2] - Mt[2][1]) - 4 * x * (Mt[2][2] + Mt[1][1]);
            o[2] = 2 * x * (Mt[1][0] + Mt[0][1]) + 2 * r * (Mt[2][0] - Mt[0][2]) + 2 * z * (Mt[1][2] + Mt[2][1]) - 4 * y * (Mt[2][2] + Mt[0][0]);
            o[3] = 2 * r * (Mt[0][1] - Mt[1][0]) + 2 * x * (Mt[2][0] + Mt[0][2]) + 2 * y * (Mt[1][2] + Mt[2][1]) - 4 * z * (Mt[1][1] + Mt[0][0]);
        }
    }
}

/* Gradient tensors as rasterize_points.cu:178-187 (all zero-filled first), order of the reference's
 * return tuple (:233).  dconic [P,4] is the reference's dL_dconic [P,2,2]. */
int or_backward(void* h, const float* dpix, const float* ddepth, const float* dflow, const float* dacc,
                float* dmean2D, float* dcol, float* dopac, float* dmean3D, float* dcov3D, float* dsh,
                float* dscale, float* drot, float* ddir, float* dconic)
{
    Ctx* c = (Ctx*)h;
    const size_t P = (size_t)c->P;
    memset(dmean2D, 0, 12 * P); memset(dcol, 0, 12 * P); memset(dopac, 0, 4 * P); memset(dmean3D, 0, 12 * P);
    memset(dcov3D, 0, 24 * P); if (c->M) memset(dsh, 0, 12 * P * (size_t)c->M); memset(dscale, 0, 12 * P);
    memset(drot, 0, 16 * P); memset(ddir, 0, 12 * P); memset(dconic, 0, 16 * P);
    if (!c->P) return 0;
    {
        double* a = (double*)calloc(14 * P, sizeof(double));      /* dmean2D[3P] | dconic[4P] | ddir[3P] | dopac[P] | dcol[3P] */
        if (!a) return -1;
        or_render_bwd(c, dpix, ddepth, dflow, dacc, a, a + 3 * P, a + 7 * P, a + 10 * P, a + 11 * P);
        for (size_t i = 0; i < 3 * P; i++) dmean2D[i] = (float)a[i];
        for (size_t i = 0; i < 4 * P; i++) dconic[i] = (float)a[3 * P + i];
        for (size_t i = 0; i < 3 * P; i++) ddir[i] = (float)a[7 * P + i];
        for (size_t i = 0; i < P; i++) dopac[i] = (float)a[10 * P + i];
        for (size_t i = 0; i < 3 * P; i++) dcol[i] = (float)a[11 * P + i];
        free(a);
    }
    or_preprocess_bwd(c, dmean2D, dconic, dcol, dmean3D, dcov3D, dsh, dscale, drot);
    return 0;
}

/* auxiliary.h:267-294 via rasterizer_impl.cu:54-68 */
void or_mark_visible(int P, const float* means, const float* view, const float* proj, float dmin, float dmax, uint8_t* present)
{
    for (int i = 0; i < P; i++) {
        const float* m = means + 3 * i;
        const float hx = xrow(proj, 0, m[0], m[1], m[2]), hy = xrow(proj, 1, m[0], m[1], m[2]), hw = xrow(proj, 3, m[0], m[1], m[2]);
        const float pw = 1.0f / (hw + 0.0000001f);
        const float nx = hx * pw, ny = hy * pw, depth = xrow(view, 2, m[0], m[1], m[2]);
        present[i] = !(depth <= dmin || depth > dmax || (double)nx < -1.3 || (double)nx > 1.3 || (double)ny < -1.3 || (double)ny > 1.3);
    }
}
