/* CPU restatement of the exact-output tile culling of ex4dgs_b200 (TEST INFRASTRUCTURE, see oracle/README.md):
 *   tile_rect      the reference's rectangle, auxiliary.h:46-56 (getRect)
 *   tight_rect     the bounding-box cut of ex4dgs_b200/csrc/common.cuh
 *   cull_prepare / cull_test   the exact per-tile test of ex4dgs_b200/csrc/common.cuh
 * followed by a brute-force verifier: for every (splat, tile) instance of the reference rectangle that the two steps
 * drop, the compositing loop's own arithmetic (forward.cu:365-381: power with the FMA placement of the reference build,
 * alpha = min(0.99, o * exp(power)), the `power > 0` and `alpha < 1/255` rejections) is evaluated at all 256 pixel
 * centres of the tile, each shifted to the four corners and the centre of the +-pad subpixel box (power is a concave
 * quadratic in the pixel position, so its maximum over the pad box is bounded by the value at the pixel nearest to the
 * splat centre - also sampled): a dropped instance with a contributing pixel is a violation.  Float operations are
 * spelled out one by one (compiled with -ffp-contract=off) so that they round like the CUDA code; where the CUDA code
 * uses MUFU approximations (tight_rect) the exact operation is used here - its guard bands are 400x wider than the
 * difference. */
#include <math.h>
#include <stdint.h>

#define TILE 16

static void tile_rect(float px, float py, int radius, int gx, int gy, int* x0, int* y0, int* x1, int* y1)
{
    const float r = (float)radius;
    int ax0 = (int)((px - r) * 0.0625f), ay0 = (int)((py - r) * 0.0625f);
    int ax1 = (int)((((px + r) + 16.0f) - 1.0f) * 0.0625f), ay1 = (int)((((py + r) + 16.0f) - 1.0f) * 0.0625f);
    *x0 = ax0 < 0 ? 0 : (ax0 > gx ? gx : ax0);
    *y0 = ay0 < 0 ? 0 : (ay0 > gy ? gy : ay0);
    *x1 = ax1 < 0 ? 0 : (ax1 > gx ? gx : ax1);
    *y1 = ay1 < 0 ? 0 : (ay1 > gy ? gy : ay1);
}

static void tight_rect(float cx, float cy, float A, float B, float C, float thr, float pad, int* x0, int* y0, int* x1, int* y1)
{
    const float detc = A * C - B * B;
    if (!((A > 0.f) && (C > 0.f) && (detc > 0.f) && (pad <= 4096.f))) return;
    const float rdet = 1.0f / detc;
    const float sAC = sqrtf(A * C) + fabsf(B);
    const float kappa = (sAC * sAC) * rdet;
    const float shrink = 1.0f - 2.01e-5f * kappa;
    if (!(shrink > 0.5f)) return;
    const float tq = 2e-3f - thr;
    if (!(tq > 0.f)) return;
    const float twoL = ((2.0f * tq) / shrink) * rdet;
    float ex = sqrtf(twoL * C), ey = sqrtf(twoL * A);
    if (!(ex < 1e6f) || !(ey < 1e6f)) return;
    ex = (ex * 1.0001f + 1e-3f) + fabsf(cx) * 1e-6f;
    ey = (ey * 1.0001f + 1e-3f) + fabsf(cy) * 1e-6f;
    const float lx = (((cx - ex) - pad) - 15.0f) * 0.0625f, hx = ((cx + ex) + pad) * 0.0625f;
    const float ly = (((cy - ey) - pad) - 15.0f) * 0.0625f, hy = ((cy + ey) + pad) * 0.0625f;
    const int tx0 = (int)ceilf(fmaxf(lx, -1.0f)), tx1 = (int)floorf(fminf(hx, 70000.0f)) + 1;
    const int ty0 = (int)ceilf(fmaxf(ly, -1.0f)), ty1 = (int)floorf(fminf(hy, 70000.0f)) + 1;
    if (tx0 > *x0) *x0 = tx0;
    if (tx1 < *x1) *x1 = tx1;
    if (ty0 > *y0) *y0 = ty0;
    if (ty1 < *y1) *y1 = ty1;
    if (*x1 < *x0) *x1 = *x0;
    if (*y1 < *y0) *y1 = *y0;
}

typedef struct { float cx, cy, A, B, C, nBoC, nBoA, shrink, tq; int ok; } Ctx;

static Ctx cull_prepare(float cx, float cy, float A, float B, float C, float thr, float pad)
{
    Ctx c;
    c.cx = cx; c.cy = cy; c.A = A; c.B = B; c.C = C;
    const float detc = A * C - B * B;
    c.ok = (A > 0.f) && (C > 0.f) && (detc > 0.f) && (pad <= 4096.f);
    c.nBoC = -B / C;
    c.nBoA = -B / A;
    const float sAC = sqrtf(A * C) + fabsf(B);
    const float kappa = (sAC * sAC) / detc;
    c.shrink = 1.0f - 2e-5f * kappa;
    if (!(c.shrink > 0.5f)) c.ok = 0;
    c.tq = 2e-3f - thr;
    return c;
}

static int cull_test(const Ctx* c, int tx, int ty, float pad)
{
    if (!c->ok) return 0;
    const float dx0 = ((float)(tx * TILE) - pad) - c->cx, dx1 = ((float)(tx * TILE + TILE - 1) + pad) - c->cx;
    const float dy0 = ((float)(ty * TILE) - pad) - c->cy, dy1 = ((float)(ty * TILE + TILE - 1) + pad) - c->cy;
    if (dx0 <= 0.f && dx1 >= 0.f && dy0 <= 0.f && dy1 >= 0.f) return 0;
    float qmin = 3.4e38f;
    for (int e = 0; e < 2; e++) {
        const float dx = e ? dx1 : dx0;
        const float dy = fminf(dy1, fmaxf(dy0, c->nBoC * dx));
        const float q = 0.5f * ((c->A * dx) * dx + (c->C * dy) * dy) + (c->B * dx) * dy;
        qmin = fminf(qmin, q);
    }
    for (int e = 0; e < 2; e++) {
        const float dy = e ? dy1 : dy0;
        const float dx = fminf(dx1, fmaxf(dx0, c->nBoA * dy));
        const float q = 0.5f * ((c->A * dx) * dx + (c->C * dy) * dy) + (c->B * dx) * dy;
        qmin = fminf(qmin, q);
    }
    return qmin * c->shrink > c->tq;
}

/* the compositing loop's test for one pixel centre (forward.cu:365-381) */
static int contributes(float cx, float cy, float A, float B, float C, float opac, float pxf, float pyf)
{
    const float dx = cx - pxf, dy = cy - pyf;
    const float power = fmaf(fmaf(dx, A * dx, (C * dy) * dy), -0.5f, -((B * dx) * dy));
    if (power > 0.0f) return 0;
    const float alpha = fminf(0.99f, opac * expf(power));
    return !(alpha < 1.0f / 255.0f);
}

/* csrc/preprocess.cu alpha_threshold(): the float thr with  min(0.99, o*expf(p)) >= 1/255  <=>  p >= thr, found by
 * bisection over the float bit patterns around log(1/(255 o)) with the loop's own expression (here: this file's expf);
 * conservative fallback when the window does not bracket the crossing. */
static int alpha_passes(float opac, unsigned m)
{
    union { float f; unsigned u; } cv;
    cv.u = m;
    return !(fminf(0.99f, opac * expf(-cv.f)) < 1.0f / 255.0f);
}

static float alpha_threshold(float opac)
{
    if (opac != opac) return opac;
    if (opac < 1.0f / 255.0f) return INFINITY;
    const float p0 = fminf(logf(1.0f / (255.0f * opac)), -0.0f);
    if (!(p0 > -3.0e38f)) return p0;
    union { float f; unsigned u; } cv;
    const float mag = -p0, d = fmaf(mag, 1e-6f, 1e-6f);
    cv.f = fmaxf(mag - d, 0.0f);
    unsigned lo = cv.u;
    cv.f = mag + d;
    unsigned hi = cv.u;
    int ok = alpha_passes(opac, lo) && !alpha_passes(opac, hi);
    if (ok) {
        while (hi - lo > 1u) {
            const unsigned mid = lo + ((hi - lo) >> 1);
            if (alpha_passes(opac, mid)) lo = mid; else hi = mid;
        }
        ok = (lo == 0u || alpha_passes(opac, lo - 1u)) && !alpha_passes(opac, hi + 1u);
    }
    if (!ok) return p0 - 1e-3f;
    cv.u = lo;
    return -cv.f;
}

/* For n splats (pixel centre cx, cy; conic A, B, C; opacity o; integer radius): counts[0] += instances of the reference
 * rectangles, [1] += instances of the tight rectangles, [2] += instances surviving the exact test, [3] += instances with
 * a contributing pixel (brute force, pad box sampled), [4] += VIOLATIONS (dropped although a pixel contributes).
 * Returns the number of violations; the index of the first offending splat goes to *first_bad (or -1). */
long cull_check(int n, const float* cx, const float* cy, const float* A, const float* B, const float* C, const float* opac,
                const int* radius, float pad, int grid_x, int grid_y, long long* counts, int* first_bad)
{
    long bad = 0;
    *first_bad = -1;
    for (int i = 0; i < n; i++) {
        int x0, y0, x1, y1;
        tile_rect(cx[i], cy[i], radius[i], grid_x, grid_y, &x0, &y0, &x1, &y1);
        if ((x1 - x0) * (y1 - y0) == 0) continue;
        const float thr = alpha_threshold(opac[i]);
        int tx0 = x0, ty0 = y0, tx1 = x1, ty1 = y1;
        tight_rect(cx[i], cy[i], A[i], B[i], C[i], thr, pad, &tx0, &ty0, &tx1, &ty1);
        const Ctx c = cull_prepare(cx[i], cy[i], A[i], B[i], C[i], thr, pad);
        for (int ty = y0; ty < y1; ty++)
            for (int tx = x0; tx < x1; tx++) {
                counts[0]++;
                const int in_tight = tx >= tx0 && tx < tx1 && ty >= ty0 && ty < ty1;
                const int kept = in_tight && !cull_test(&c, tx, ty, pad);
                counts[1] += in_tight;
                counts[2] += kept;
                int any = 0;
                for (int py = 0; py < TILE && !any; py++)
                    for (int px = 0; px < TILE && !any; px++) {
                        const float bx = (float)(tx * TILE + px), by = (float)(ty * TILE + py);
                        /* centre, four corners of the pad box, and the point of the box nearest to the splat centre */
                        const float nx = fminf(bx + pad, fmaxf(bx - pad, cx[i])), ny = fminf(by + pad, fmaxf(by - pad, cy[i]));
                        const float sx[6] = {bx, bx - pad, bx + pad, bx - pad, bx + pad, nx};
                        const float sy[6] = {by, by - pad, by - pad, by + pad, by + pad, ny};
                        for (int s = 0; s < 6 && !any; s++) any = contributes(cx[i], cy[i], A[i], B[i], C[i], opac[i], sx[s], sy[s]);
                    }
                counts[3] += any;
                if (any && !kept) {
                    counts[4]++;
                    if (bad == 0) *first_bad = i;
                    bad++;
                }
            }
    }
    return bad;
}

/* ---- the two finer culling levels of the compositing kernels --------------------------------------------------
 * block_reject (csrc/common.cuh): a warp drops a splat for its whole 8 x bh pixel block (bh = 4 forward, 8 backward)
 * when the block's bounding box of pixel centres lies outside the alpha >= 1/255 ellipse's axis-aligned box;
 * skip threshold (render_fwd.cu / render_bwd.cu): a (pixel, splat) pair with power < thr (the exact crossing of the alpha
 * test, alpha_threshold() of csrc/preprocess.cu restated above with this file's expf) is skipped without evaluating exp().  Both are verified against the compositing loop's own test at every pixel. */
static int block_reject(float cx, float cy, float thr, float A, float B, float C, float x0, float x1, float y0, float y1)
{
    const float det = A * C - B * B;
    if (!(A > 0.f) || !(C > 0.f) || !(det > 0.f)) return 0;
    const float inv = 1.0f / det;
    const float shrink = 1.0f - 8e-5f * (A * C * inv);
    if (!(shrink > 0.5f)) return 0;
    const float tq = 2e-3f - thr;
    const float lim = 2.0f * tq * inv / shrink * 1.0001f;
    const float dx = fmaxf(fmaxf(x0 - cx, cx - x1), 0.f);
    const float dy = fmaxf(fmaxf(y0 - cy, cy - y1), 0.f);
    return (dx * dx > lim * C) || (dy * dy > lim * A);
}

/* counts[0] += (splat, block) pairs examined, [1] += rejected blocks, [2] += (pixel, splat) pairs below the skip
 * threshold, [3] += contributing pairs, [4] += violations of block_reject, [5] += violations of the skip threshold.
 * (ox, oy): a subpixel offset applied to every pixel centre of the run. */
long block_check(int n, const float* cx, const float* cy, const float* A, const float* B, const float* C, const float* opac,
                 const int* radius, int bh, float ox, float oy, int grid_x, int grid_y, long long* counts)
{
    long bad = 0;
    for (int i = 0; i < n; i++) {
        int x0, y0, x1, y1;
        tile_rect(cx[i], cy[i], radius[i], grid_x, grid_y, &x0, &y0, &x1, &y1);
        const float thr = alpha_threshold(opac[i]);
        for (int by = y0 * TILE; by < y1 * TILE; by += bh)
            for (int bx = x0 * TILE; bx < x1 * TILE; bx += 8) {
                const float fx0 = (float)bx + ox, fx1 = (float)(bx + 7) + ox, fy0 = (float)by + oy, fy1 = (float)(by + bh - 1) + oy;
                const int rej = block_reject(cx[i], cy[i], thr, A[i], B[i], C[i], fx0, fx1, fy0, fy1);
                counts[0]++;
                counts[1] += rej;
                for (int py = 0; py < bh; py++)
                    for (int px = 0; px < 8; px++) {
                        const float pxf = (float)(bx + px) + ox, pyf = (float)(by + py) + oy;
                        const float dx = cx[i] - pxf, dy = cy[i] - pyf;
                        const float power = fmaf(fmaf(dx, A[i] * dx, (C[i] * dy) * dy), -0.5f, -((B[i] * dx) * dy));
                        const int con = contributes(cx[i], cy[i], A[i], B[i], C[i], opac[i], pxf, pyf);
                        counts[3] += con;
                        if (power < thr) {
                            counts[2]++;
                            if (con) { counts[5]++; bad++; }
                        }
                        if (rej && con) { counts[4]++; bad++; }
                    }
            }
    }
    return bad;
}
