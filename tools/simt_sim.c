/* Analysis tool (not product, not a test): replays the forward compositing of one frame on the CPU from the oracle's
 * tile lists and counts, per warp-shaped pixel block, how often the warp-level code sections of render_fwd_kernel
 * would execute under different work mappings.  Build: gcc -O2 -fopenmp -shared -fPIC -I../oracle ...  (tools/simt_sim.py) */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include "cull_check.c"

/* per (pixel, splat): 0 = fails the cheap test, 1 = passes the cheap test (blend body entered), 2 = blended */
static inline int pair_state(float cx, float cy, float A, float B, float C, float o, float thr, float pxf, float pyf, float* alpha)
{
    const float dx = cx - pxf, dy = cy - pyf;
    const float power = fmaf(fmaf(dx, A * dx, (C * dy) * dy), -0.5f, -((B * dx) * dy));
    if (power > 0.0f || power < thr) return 0;
    *alpha = fminf(0.99f, o * expf(power));
    return (*alpha < 1.0f / 255.0f) ? 1 : 2;
}

/* out[0] (warp, splat) pairs after the block test; out[1] warp-level blend-body executions, current scheme (a splat's
 * body runs when any live lane passes the cheap test); out[2] lane-slots active in them; out[3] blend phases of the
 * lane-private-cursor scheme (per batch: max over lanes of the passing splats); out[4] lane-slots active in them;
 * out[5] cursor-advance steps of that scheme (per batch: max over lanes of candidates examined);
 * out[6] pixel-level blends; out[7] fetched instances (R_eff); out[8] listed instances (after tile culling). */
void simulate(int W, int H, int bw, int bh, int batch, int use_cull,
              const float* means2D, const float* conic_opacity, const uint32_t* point_list, const uint32_t* ranges,
              long long* out)
{
    const int gx = (W + 15) / 16, gy = (H + 15) / 16;
    long long acc[9];
    memset(acc, 0, sizeof(acc));
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : acc[:9])
    for (int tile = 0; tile < gx * gy; tile++) {
        const int tx = tile % gx, ty = tile / gx;
        const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        /* exact tile culling first (what the duplicate kernel does) */
        uint32_t* list = (uint32_t*)malloc(sizeof(uint32_t) * (r1 - r0 + 1));
        int n = 0;
        for (uint32_t q = r0; q < r1; q++) {
            const uint32_t id = point_list[q];
            const float cx = means2D[2 * id], cy = means2D[2 * id + 1];
            const float A = conic_opacity[4 * id], B = conic_opacity[4 * id + 1], C = conic_opacity[4 * id + 2], o = conic_opacity[4 * id + 3];
            if (use_cull) {
                const float thr = logf(1.0f / (255.0f * o)) - 1e-3f;
                const Ctx c = cull_prepare(cx, cy, A, B, C, thr, 0.f);
                if (cull_test(&c, tx, ty, 0.f)) continue;
            }
            list[n++] = id;
        }
        acc[8] += n;
        const int nbx = 16 / bw, nby = 16 / bh, nblk = nbx * nby, lanes = bw * bh;
        float T[256];
        int done[256];
        for (int i = 0; i < 256; i++) { T[i] = 1.0f; done[i] = 0; }
        for (int p = 0; p < 256; p++) {
            const int px = tx * 16 + p % 16, py = ty * 16 + p / 16;
            if (px >= W || py >= H) done[p] = 1;
        }
        for (int b0 = 0; b0 < n; b0 += batch) {
            int alive = 0;
            for (int p = 0; p < 256; p++) alive += !done[p];
            if (!alive) break;
            const int cnt = (n - b0 < batch) ? n - b0 : batch;
            acc[7] += cnt;
            for (int blk = 0; blk < nblk; blk++) {
                const int ox = (blk % nbx) * bw, oy = (blk / nbx) * bh;
                int any_alive = 0;
                for (int l = 0; l < lanes; l++) any_alive |= !done[(oy + l / bw) * 16 + ox + l % bw];
                if (!any_alive) continue;
                const float fx0 = (float)(tx * 16 + ox), fx1 = fx0 + (float)(bw - 1), fy0 = (float)(ty * 16 + oy), fy1 = fy0 + (float)(bh - 1);
                int lane_pass[64], lane_seen[64];
                for (int l = 0; l < lanes; l++) lane_pass[l] = lane_seen[l] = 0;
                int cand = 0;
                for (int j = 0; j < cnt; j++) {
                    const uint32_t id = list[b0 + j];
                    const float cx = means2D[2 * id], cy = means2D[2 * id + 1];
                    const float A = conic_opacity[4 * id], B = conic_opacity[4 * id + 1], C = conic_opacity[4 * id + 2], o = conic_opacity[4 * id + 3];
                    const float thr = logf(1.0f / (255.0f * o)) - 1e-3f;
                    if (block_reject(cx, cy, thr, A, B, C, fx0, fx1, fy0, fy1)) continue;
                    acc[0]++;
                    cand++;
                    int any_pass = 0, n_pass = 0;
                    for (int l = 0; l < lanes; l++) {
                        const int p = (oy + l / bw) * 16 + ox + l % bw;
                        if (done[p]) continue;
                        lane_seen[l] = cand;
                        float alpha = 0.f;
                        const int st = pair_state(cx, cy, A, B, C, o, thr, (float)(tx * 16 + ox + l % bw), (float)(ty * 16 + oy + l / bw), &alpha);
                        if (st == 0) continue;
                        any_pass = 1; n_pass++;
                        lane_pass[l]++;
                        if (st == 2) {
                            const float test_T = T[p] * (1.0f - alpha);
                            if (test_T < 0.0001f) { done[p] = 1; continue; }
                            T[p] = test_T;
                            acc[6]++;
                        }
                    }
                    if (any_pass) { acc[1]++; acc[2] += n_pass; }
                }
                int mp = 0, ms = 0;
                for (int l = 0; l < lanes; l++) { if (lane_pass[l] > mp) mp = lane_pass[l]; if (lane_seen[l] > ms) ms = lane_seen[l]; acc[4] += lane_pass[l]; }
                acc[3] += mp;
                acc[5] += ms;
            }
        }
        free(list);
    }
    for (int i = 0; i < 9; i++) out[i] = acc[i];
}

/* Lane-private cursors, faithful round structure.  Per warp and batch: every live lane walks from its cursor to its next
 * candidate that passes the cheap test (lanes that hit early wait for the slowest walker), then all lanes that found one
 * run the blend body together.  out[0] walk steps (warp-level: per round the longest walk), out[1] blend rounds,
 * out[2] lane-slots active in the blend rounds, out[3] lane-slots active in the walk steps, out[4] pixel-level blends. */
void simulate_cursor(int W, int H, int bw, int bh, int batch,
                     const float* means2D, const float* conic_opacity, const uint32_t* point_list, const uint32_t* ranges,
                     long long* out)
{
    const int gx = (W + 15) / 16, gy = (H + 15) / 16;
    long long acc[5];
    memset(acc, 0, sizeof(acc));
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : acc[:5])
    for (int tile = 0; tile < gx * gy; tile++) {
        const int tx = tile % gx, ty = tile / gx;
        const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        uint32_t* list = (uint32_t*)malloc(sizeof(uint32_t) * (r1 - r0 + 1));
        int n = 0;
        for (uint32_t q = r0; q < r1; q++) {
            const uint32_t id = point_list[q];
            const float cx = means2D[2 * id], cy = means2D[2 * id + 1];
            const float A = conic_opacity[4 * id], B = conic_opacity[4 * id + 1], C = conic_opacity[4 * id + 2], o = conic_opacity[4 * id + 3];
            const float thr = logf(1.0f / (255.0f * o)) - 1e-3f;
            const Ctx c = cull_prepare(cx, cy, A, B, C, thr, 0.f);
            if (cull_test(&c, tx, ty, 0.f)) continue;
            list[n++] = id;
        }
        const int nbx = 16 / bw, nby = 16 / bh, nblk = nbx * nby, lanes = bw * bh;
        float T[256];
        int done[256];
        for (int i = 0; i < 256; i++) { T[i] = 1.0f; done[i] = 0; }
        for (int p = 0; p < 256; p++) {
            const int px = tx * 16 + p % 16, py = ty * 16 + p / 16;
            if (px >= W || py >= H) done[p] = 1;
        }
        unsigned char* st = (unsigned char*)malloc((size_t)batch * 64);
        float* al = (float*)malloc(sizeof(float) * (size_t)batch * 64);
        for (int b0 = 0; b0 < n; b0 += batch) {
            int alive = 0;
            for (int p = 0; p < 256; p++) alive += !done[p];
            if (!alive) break;
            const int cnt = (n - b0 < batch) ? n - b0 : batch;
            for (int blk = 0; blk < nblk; blk++) {
                const int ox = (blk % nbx) * bw, oy = (blk / nbx) * bh;
                const float fx0 = (float)(tx * 16 + ox), fx1 = fx0 + (float)(bw - 1), fy0 = (float)(ty * 16 + oy), fy1 = fy0 + (float)(bh - 1);
                int cand = 0;
                for (int j = 0; j < cnt; j++) {
                    const uint32_t id = list[b0 + j];
                    const float cx = means2D[2 * id], cy = means2D[2 * id + 1];
                    const float A = conic_opacity[4 * id], B = conic_opacity[4 * id + 1], C = conic_opacity[4 * id + 2], o = conic_opacity[4 * id + 3];
                    const float thr = logf(1.0f / (255.0f * o)) - 1e-3f;
                    if (block_reject(cx, cy, thr, A, B, C, fx0, fx1, fy0, fy1)) continue;
                    for (int l = 0; l < lanes; l++) {
                        float alpha = 0.f;
                        st[cand * 64 + l] = (unsigned char)pair_state(cx, cy, A, B, C, o, thr, (float)(tx * 16 + ox + l % bw),
                                                                      (float)(ty * 16 + oy + l / bw), &alpha);
                        al[cand * 64 + l] = alpha;
                    }
                    cand++;
                }
                int cur[64];
                for (int l = 0; l < lanes; l++) cur[l] = 0;
                for (;;) {
                    int maxgap = 0, found_any = 0, nfound = 0, nwalk = 0;
                    int hit[64];
                    for (int l = 0; l < lanes; l++) {
                        const int p = (oy + l / bw) * 16 + ox + l % bw;
                        hit[l] = -1;
                        if (done[p] || cur[l] >= cand) continue;
                        int c = cur[l];
                        while (c < cand && st[c * 64 + l] == 0) c++;
                        const int gap = (c < cand) ? c - cur[l] + 1 : cand - cur[l];
                        if (gap > maxgap) maxgap = gap;
                        nwalk += gap;
                        if (c < cand) { hit[l] = c; found_any = 1; nfound++; }
                        cur[l] = (c < cand) ? c + 1 : cand;
                    }
                    acc[0] += maxgap;
                    acc[3] += nwalk;
                    if (!found_any) break;
                    acc[1]++;
                    acc[2] += nfound;
                    for (int l = 0; l < lanes; l++) {
                        if (hit[l] < 0) continue;
                        const int p = (oy + l / bw) * 16 + ox + l % bw;
                        if (st[hit[l] * 64 + l] == 2) {
                            const float test_T = T[p] * (1.0f - al[hit[l] * 64 + l]);
                            if (test_T < 0.0001f) { done[p] = 1; continue; }
                            T[p] = test_T;
                            acc[4]++;
                        }
                    }
                }
            }
        }
        free(st); free(al); free(list);
    }
    for (int i = 0; i < 5; i++) out[i] = acc[i];
}

/* Broadcast test as today + a per-lane FIFO (depth Q) of candidates that passed the cheap test; a blend round (every
 * lane with a non-empty FIFO pops one and runs the blend body) is issued whenever some lane's FIFO is full, and the
 * FIFOs are drained at the end of the batch.  out[0] blend rounds, out[1] lane-slots active in them, out[2] blends. */
void simulate_queue(int W, int H, int bw, int bh, int batch, int Q,
                    const float* means2D, const float* conic_opacity, const uint32_t* point_list, const uint32_t* ranges,
                    long long* out)
{
    const int gx = (W + 15) / 16, gy = (H + 15) / 16;
    long long acc[3];
    memset(acc, 0, sizeof(acc));
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : acc[:3])
    for (int tile = 0; tile < gx * gy; tile++) {
        const int tx = tile % gx, ty = tile / gx;
        const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        uint32_t* list = (uint32_t*)malloc(sizeof(uint32_t) * (r1 - r0 + 1));
        int n = 0;
        for (uint32_t q = r0; q < r1; q++) {
            const uint32_t id = point_list[q];
            const float cx = means2D[2 * id], cy = means2D[2 * id + 1];
            const float A = conic_opacity[4 * id], B = conic_opacity[4 * id + 1], C = conic_opacity[4 * id + 2], o = conic_opacity[4 * id + 3];
            const float thr = logf(1.0f / (255.0f * o)) - 1e-3f;
            const Ctx c = cull_prepare(cx, cy, A, B, C, thr, 0.f);
            if (cull_test(&c, tx, ty, 0.f)) continue;
            list[n++] = id;
        }
        const int nbx = 16 / bw, nby = 16 / bh, nblk = nbx * nby, lanes = bw * bh;
        float T[256];
        int done[256];
        for (int i = 0; i < 256; i++) { T[i] = 1.0f; done[i] = 0; }
        for (int p = 0; p < 256; p++) {
            const int px = tx * 16 + p % 16, py = ty * 16 + p / 16;
            if (px >= W || py >= H) done[p] = 1;
        }
        for (int b0 = 0; b0 < n; b0 += batch) {
            int alive = 0;
            for (int p = 0; p < 256; p++) alive += !done[p];
            if (!alive) break;
            const int cnt = (n - b0 < batch) ? n - b0 : batch;
            for (int blk = 0; blk < nblk; blk++) {
                const int ox = (blk % nbx) * bw, oy = (blk / nbx) * bh;
                const float fx0 = (float)(tx * 16 + ox), fx1 = fx0 + (float)(bw - 1), fy0 = (float)(ty * 16 + oy), fy1 = fy0 + (float)(bh - 1);
                float qa[64][64];        /* FIFO of alphas (-1: passed the cheap test but alpha < 1/255) */
                int qn[64];
                for (int l = 0; l < lanes; l++) qn[l] = 0;
                for (int j = 0; j <= cnt; j++) {
                    int full = 0;
                    if (j < cnt) {
                        const uint32_t id = list[b0 + j];
                        const float cx = means2D[2 * id], cy = means2D[2 * id + 1];
                        const float A = conic_opacity[4 * id], B = conic_opacity[4 * id + 1], C = conic_opacity[4 * id + 2], o = conic_opacity[4 * id + 3];
                        const float thr = logf(1.0f / (255.0f * o)) - 1e-3f;
                        if (block_reject(cx, cy, thr, A, B, C, fx0, fx1, fy0, fy1)) continue;
                        for (int l = 0; l < lanes; l++) {
                            const int p = (oy + l / bw) * 16 + ox + l % bw;
                            if (done[p]) continue;
                            float alpha = 0.f;
                            const int st = pair_state(cx, cy, A, B, C, o, thr, (float)(tx * 16 + ox + l % bw), (float)(ty * 16 + oy + l / bw), &alpha);
                            if (st == 0) continue;
                            qa[l][qn[l]++] = (st == 2) ? alpha : -1.f;
                            if (qn[l] >= Q) full = 1;
                        }
                    }
                    /* one round when a FIFO is full; drain completely at the end of the batch */
                    for (;;) {
                        int any = 0, act = 0;
                        if (!(full || j == cnt)) break;
                        for (int l = 0; l < lanes; l++) {
                            if (qn[l] == 0) continue;
                            const int p = (oy + l / bw) * 16 + ox + l % bw;
                            const float alpha = qa[l][0];
                            for (int k = 1; k < qn[l]; k++) qa[l][k - 1] = qa[l][k];
                            qn[l]--;
                            if (done[p]) { qn[l] = 0; continue; }
                            any = 1; act++;
                            if (alpha >= 0.f) {
                                const float test_T = T[p] * (1.0f - alpha);
                                if (test_T < 0.0001f) { done[p] = 1; qn[l] = 0; continue; }
                                T[p] = test_T;
                                acc[2]++;
                            }
                        }
                        if (any) { acc[0]++; acc[1] += act; }
                        full = 0;
                        if (j < cnt) break;
                        int left = 0;
                        for (int l = 0; l < lanes; l++) left += qn[l];
                        if (!left) break;
                    }
                }
            }
        }
        free(list);
    }
    for (int i = 0; i < 3; i++) out[i] = acc[i];
}
