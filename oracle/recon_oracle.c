/*
 * recon_oracle.c - CPU restatement of the reference's VP8 reconstruction path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path: it is
 * imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; the
 * product (libvpx.opencl_b200/, hostdec/) never links, loads or calls it.
 *
 * Parity status: PINNED.  (a) tests/test_oracle_vs_reference.py checks every function here
 * against the unmodified reference compiled from /root/reference (oracle/_ref/libvpxref.so)
 * on seeded random inputs; (b) tests/test_oracle_golden.py replays committed record dumps
 * of reference-encoded streams and requires the per-frame MD5s printed by the reference's
 * own `vpxdec --md5` (tests/golden/<case>.md5).  The reference tree itself holds no golden
 * vectors, so the pins are generated from the reference, with the
 * generating scripts committed (tools/make_golden.py).
 *
 * It consumes the same per-frame records as the C ABI (include/vp8b200.h) and walks the
 * frame in the reference's own order - raster MB loop, then whole-frame loop filter, then
 * border extension - scribbling the same helper bytes into the frame buffer that the
 * reference does (127/129 intra edges, 4-byte row extension, above-right down copy), so the
 * complete buffer including borders can be compared byte for byte.
 *
 * Each function cites the reference file:line it restates (paths relative to the
 * reference root).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "vp8b200.h"

typedef struct oracle_dec {
    int w, h;                     /* coded size (multiples of 16) */
    int mb_cols, mb_rows;
    int y_stride, uv_stride;
    size_t yplane, uvplane, frame_size;
    int n_fb;
    uint8_t *fb[VP8B200_MAX_FB];
} oracle_dec;

/* frame-buffer geometry: vpx_scale/generic/yv12config.c:55-110 */
static uint8_t *plane_y(const oracle_dec *d, int fb) { return d->fb[fb] + 32 * d->y_stride + 32; }
static uint8_t *plane_u(const oracle_dec *d, int fb) { return d->fb[fb] + d->yplane + 16 * d->uv_stride + 16; }
static uint8_t *plane_v(const oracle_dec *d, int fb) { return d->fb[fb] + d->yplane + d->uvplane + 16 * d->uv_stride + 16; }

oracle_dec *oracle_create(int width, int height, int n_fb)
{
    oracle_dec *d;
    int i;
    if ((width & 15) || (height & 15) || n_fb < 1 || n_fb > VP8B200_MAX_FB) return NULL;
    d = (oracle_dec *)calloc(1, sizeof *d);
    d->w = width; d->h = height;
    d->mb_cols = width >> 4; d->mb_rows = height >> 4;
    d->y_stride = ((width + 64) + 31) & ~31;
    d->uv_stride = d->y_stride >> 1;
    d->yplane = (size_t)(height + 64) * d->y_stride;
    d->uvplane = (size_t)((height >> 1) + 32) * d->uv_stride;
    d->frame_size = d->yplane + 2 * d->uvplane;
    d->n_fb = n_fb;
    for (i = 0; i < n_fb; i++) d->fb[i] = (uint8_t *)calloc(1, d->frame_size);
    return d;
}

void oracle_destroy(oracle_dec *d)
{
    int i;
    if (!d) return;
    for (i = 0; i < d->n_fb; i++) free(d->fb[i]);
    free(d);
}

size_t oracle_frame_size(const oracle_dec *d) { return d->frame_size; }
uint8_t *oracle_fb(oracle_dec *d, int fb) { return d->fb[fb]; }
int oracle_y_stride(const oracle_dec *d) { return d->y_stride; }

static uint8_t clamp255(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

/* ------------------------------------------------------------------------------------------
 * A1-A3: dequantisation, inverse DCT, inverse WHT
 * ---------------------------------------------------------------------------------------- */

/* vp8/common/idctllm.c:28-110 (vp8_short_idct4x4llm_c): vertical pass into 16-bit storage,
 * horizontal pass with (x+4)>>3, add to the predictor already sitting in dst, clamp. */
void oracle_idct_add(const int16_t *in, uint8_t *dst, int stride)
{
    int16_t mid[16];
    int i, r, c;
    for (i = 0; i < 4; i++) {
        int a = in[i] + in[8 + i];
        int b = in[i] - in[8 + i];
        int t1 = (in[4 + i] * 35468) >> 16;
        int t2 = in[12 + i] + ((in[12 + i] * 20091) >> 16);
        int cc = t1 - t2;
        int dd;
        t1 = in[4 + i] + ((in[4 + i] * 20091) >> 16);
        t2 = (in[12 + i] * 35468) >> 16;
        dd = t1 + t2;
        mid[i]      = (int16_t)(a + dd);
        mid[12 + i] = (int16_t)(a - dd);
        mid[4 + i]  = (int16_t)(b + cc);
        mid[8 + i]  = (int16_t)(b - cc);
    }
    for (r = 0; r < 4; r++) {
        const int16_t *m = mid + 4 * r;
        int16_t o[4];
        int a = m[0] + m[2];
        int b = m[0] - m[2];
        int t1 = (m[1] * 35468) >> 16;
        int t2 = m[3] + ((m[3] * 20091) >> 16);
        int cc = t1 - t2;
        int dd;
        t1 = m[1] + ((m[1] * 20091) >> 16);
        t2 = (m[3] * 35468) >> 16;
        dd = t1 + t2;
        o[0] = (int16_t)((a + dd + 4) >> 3);
        o[3] = (int16_t)((a - dd + 4) >> 3);
        o[1] = (int16_t)((b + cc + 4) >> 3);
        o[2] = (int16_t)((b - cc + 4) >> 3);
        for (c = 0; c < 4; c++) dst[r * stride + c] = clamp255(o[c] + dst[r * stride + c]);
    }
}

/* vp8/common/idctllm.c:112-138 (vp8_dc_only_idct_add_c) */
void oracle_dc_add(int16_t dc, uint8_t *dst, int stride)
{
    int a = (dc + 4) >> 3, r, c;
    for (r = 0; r < 4; r++)
        for (c = 0; c < 4; c++) dst[r * stride + c] = clamp255(a + dst[r * stride + c]);
}

/* vp8/common/idctllm.c:140-192 (vp8_short_inv_walsh4x4_c); out[i] is the DC of Y block i */
void oracle_iwalsh(const int16_t *in, int16_t *out)
{
    int16_t mid[16];
    int i;
    for (i = 0; i < 4; i++) {
        int a = in[i] + in[12 + i];
        int b = in[4 + i] + in[8 + i];
        int c = in[4 + i] - in[8 + i];
        int d = in[i] - in[12 + i];
        mid[i] = (int16_t)(a + b);
        mid[4 + i] = (int16_t)(c + d);
        mid[8 + i] = (int16_t)(a - b);
        mid[12 + i] = (int16_t)(d - c);
    }
    for (i = 0; i < 4; i++) {
        const int16_t *m = mid + 4 * i;
        int a = m[0] + m[3];
        int b = m[1] + m[2];
        int c = m[1] - m[2];
        int d = m[0] - m[3];
        out[4 * i + 0] = (int16_t)((a + b + 3) >> 3);
        out[4 * i + 1] = (int16_t)((c + d + 3) >> 3);
        out[4 * i + 2] = (int16_t)((a - b + 3) >> 3);
        out[4 * i + 3] = (int16_t)((d - c + 3) >> 3);
    }
}

/* vp8/common/dequantize.c:29-43 + idct_blk.c:20-44: one block.  `present` = the record
 * carries coefficients for the block.  The reference takes the full transform for eob > 1
 * and the DC shortcut otherwise; the two agree whenever only the DC is non-zero, so
 * "present" selects the full transform and absent blocks keep the DC shortcut on whatever
 * DC is in q[0] (the WHT output for Y2 macroblocks, else 0). */
static void dequant_idct_block(int16_t *q, int present, int dc_factor, int ac_factor,
                               uint8_t *dst, int stride)
{
    if (present) {
        int16_t dq[16];
        int i;
        dq[0] = (int16_t)(q[0] * dc_factor);                       /* dequantize.c:36 */
        for (i = 1; i < 16; i++) dq[i] = (int16_t)(q[i] * ac_factor);
        oracle_idct_add(dq, dst, stride);
    } else {
        oracle_dc_add((int16_t)(q[0] * dc_factor), dst, stride);   /* idct_blk.c:34 */
    }
}

/* ------------------------------------------------------------------------------------------
 * A4-A5: sub-pixel interpolation
 * ---------------------------------------------------------------------------------------- */

static const int16_t k_sixtap[8][6] = {           /* vp8/common/filter.c:28-39 */
    {0, 0, 128, 0, 0, 0},    {0, -6, 123, 12, -1, 0}, {2, -11, 108, 36, -8, 1},
    {0, -9, 93, 50, -6, 0},  {3, -16, 77, 77, -16, 3}, {0, -6, 50, 93, -9, 0},
    {1, -8, 36, 108, -11, 2}, {0, -1, 12, 123, -6, 0}};
static const int16_t k_bilinear[8][2] = {         /* vp8/common/filter.c:16-26 */
    {128, 0}, {112, 16}, {96, 32}, {80, 48}, {64, 64}, {48, 80}, {32, 96}, {16, 112}};

/* filter.c:41-129,152-246: H pass over h+5 rows starting two rows up, rounded, clamped to
 * 0..255; V pass over that, same rounding and clamp.  Always both passes. */
void oracle_sixtap(const uint8_t *src, int sstride, int xoff, int yoff,
                   uint8_t *dst, int dstride, int w, int h)
{
    int tmp[21 * 16];
    const int16_t *hf = k_sixtap[xoff], *vf = k_sixtap[yoff];
    int r, c;
    for (r = 0; r < h + 5; r++) {
        const uint8_t *s = src + (r - 2) * sstride;
        for (c = 0; c < w; c++) {
            int t = s[c - 2] * hf[0] + s[c - 1] * hf[1] + s[c] * hf[2] + s[c + 1] * hf[3] +
                    s[c + 2] * hf[4] + s[c + 3] * hf[5] + 64;
            tmp[r * w + c] = clamp255(t >> 7);
        }
    }
    for (r = 0; r < h; r++)
        for (c = 0; c < w; c++) {
            const int *t = tmp + (r + 2) * w + c;
            int v = t[-2 * w] * vf[0] + t[-w] * vf[1] + t[0] * vf[2] + t[w] * vf[3] +
                    t[2 * w] * vf[4] + t[3 * w] * vf[5] + 64;
            dst[r * dstride + c] = clamp255(v >> 7);
        }
}

/* filter.c:271-397: H pass over h+1 rows into 16-bit, V pass; no clamp anywhere */
void oracle_bilinear(const uint8_t *src, int sstride, int xoff, int yoff,
                     uint8_t *dst, int dstride, int w, int h)
{
    uint16_t tmp[17 * 16];
    const int16_t *hf = k_bilinear[xoff], *vf = k_bilinear[yoff];
    int r, c;
    for (r = 0; r < h + 1; r++)
        for (c = 0; c < w; c++)
            tmp[r * w + c] = (uint16_t)((src[r * sstride + c] * hf[0] +
                                         src[r * sstride + c + 1] * hf[1] + 64) >> 7);
    for (r = 0; r < h; r++)
        for (c = 0; c < w; c++)
            dst[r * dstride + c] =
                (uint8_t)((tmp[r * w + c] * vf[0] + tmp[(r + 1) * w + c] * vf[1] + 64) >> 7);
}

/* one prediction block: reconinter.c:125-221 (copy when both fractions are 0, else filter) */
static void predict_block(int bilinear, const uint8_t *ref, int stride, int mv_row, int mv_col,
                          uint8_t *dst, int w, int h)
{
    const uint8_t *src = ref + (mv_row >> 3) * stride + (mv_col >> 3);
    if ((mv_row & 7) || (mv_col & 7)) {
        if (bilinear) oracle_bilinear(src, stride, mv_col & 7, mv_row & 7, dst, stride, w, h);
        else          oracle_sixtap(src, stride, mv_col & 7, mv_row & 7, dst, stride, w, h);
    } else {
        int r;
        for (r = 0; r < h; r++) memcpy(dst + r * stride, src + r * stride, (size_t)w);
    }
}

typedef struct { int row, col; } mv_t;
typedef struct { int left, right, top, bottom; } edges_t;   /* decodframe.c:351-365, 1/8 pel */

/* reconinter.c:348-368 */
static mv_t clamp_mv(mv_t mv, const edges_t *e)
{
    if (mv.col < e->left - (19 << 3)) mv.col = e->left - (16 << 3);
    else if (mv.col > e->right + (18 << 3)) mv.col = e->right + (16 << 3);
    if (mv.row < e->top - (19 << 3)) mv.row = e->top - (16 << 3);
    else if (mv.row > e->bottom + (18 << 3)) mv.row = e->bottom + (16 << 3);
    return mv;
}

/* reconinter.c:371-382 */
static mv_t clamp_uvmv(mv_t mv, const edges_t *e)
{
    if (2 * mv.col < e->left - (19 << 3)) mv.col = (e->left - (16 << 3)) >> 1;
    if (2 * mv.col > e->right + (18 << 3)) mv.col = (e->right + (16 << 3)) >> 1;
    if (2 * mv.row < e->top - (19 << 3)) mv.row = (e->top - (16 << 3)) >> 1;
    if (2 * mv.row > e->bottom + (18 << 3)) mv.row = (e->bottom + (16 << 3)) >> 1;
    return mv;
}

/* ------------------------------------------------------------------------------------------
 * A6: inter predictor builder (reconinter.c:384-573)
 * ---------------------------------------------------------------------------------------- */
static void inter_predict_mb(const oracle_dec *d, const vp8b200_frame_hdr *hdr,
                             const vp8b200_mb *mb, const vp8b200_aux *aux,
                             int mb_row, int mb_col, uint8_t *dy, uint8_t *du, uint8_t *dv)
{
    int ref_fb = mb->ref_frame == VP8B200_LAST_FRAME ? hdr->fb_last
               : mb->ref_frame == VP8B200_GOLDEN_FRAME ? hdr->fb_golden : hdr->fb_altref;
    size_t yoff = (size_t)mb_row * 16 * d->y_stride + mb_col * 16;
    size_t uvoff = (size_t)mb_row * 8 * d->uv_stride + mb_col * 8;
    const uint8_t *ry = plane_y(d, ref_fb) + yoff;
    const uint8_t *ru = plane_u(d, ref_fb) + uvoff;
    const uint8_t *rv = plane_v(d, ref_fb) + uvoff;
    int bil = hdr->use_bilinear_mc;
    int fpmask = hdr->full_pixel ? ~7 : ~0;
    int need_clamp = (mb->flags & VP8B200_MBF_CLAMP_MVS) != 0;
    edges_t e;
    e.left = -((mb_col * 16) << 3);
    e.right = ((d->mb_cols - 1 - mb_col) * 16) << 3;
    e.top = -((mb_row * 16) << 3);
    e.bottom = ((d->mb_rows - 1 - mb_row) * 16) << 3;

    if (mb->y_mode != VP8B200_SPLITMV) {
        /* reconinter.c:384-441 */
        mv_t mv, uvmv;
        mv.row = mb->u.mv.row; mv.col = mb->u.mv.col;
        if (need_clamp) mv = clamp_mv(mv, &e);
        predict_block(bil, ry, d->y_stride, mv.row, mv.col, dy, 16, 16);
        /* chroma MV from the (clamped) luma MV; the fields are 16-bit, so the shifts by 31
         * act on the sign-extended value: reconinter.c:419-424 */
        uvmv.row = (int16_t)(mv.row + (1 | (mv.row >> 31)));
        uvmv.col = (int16_t)(mv.col + (1 | (mv.col >> 31)));
        uvmv.row = (int16_t)(uvmv.row / 2);
        uvmv.col = (int16_t)(uvmv.col / 2);
        uvmv.row &= fpmask;
        uvmv.col &= fpmask;
        uvmv.row = (int16_t)uvmv.row; uvmv.col = (int16_t)uvmv.col;
        predict_block(bil, ru, d->uv_stride, uvmv.row, uvmv.col, du, 8, 8);
        predict_block(bil, rv, d->uv_stride, uvmv.row, uvmv.col, dv, 8, 8);
    } else {
        /* build_4x4uvmvs (reconinter.c:520-558) then build_inter4x4_predictors_mb (:443-517).
         * Block partition (8x8 / 8x4 / 4x4) does not change any output pixel: each pixel is
         * a function of its own block MV only, so every luma 4x4 is predicted by itself. */
        const vp8b200_aux *a = &aux[mb->u.aux];
        mv_t uv[4];
        int i, j, b;
        for (i = 0; i < 2; i++)
            for (j = 0; j < 2; j++) {
                int y0 = i * 8 + j * 2, t;
                mv_t m;
                t = a->mv[y0].row + a->mv[y0 + 1].row + a->mv[y0 + 4].row + a->mv[y0 + 5].row;
                t += 4 + ((t >> 31) << 3);
                m.row = (int16_t)((t / 8) & fpmask);
                t = a->mv[y0].col + a->mv[y0 + 1].col + a->mv[y0 + 4].col + a->mv[y0 + 5].col;
                t += 4 + ((t >> 31) << 3);
                m.col = (int16_t)((t / 8) & fpmask);
                if (need_clamp) m = clamp_uvmv(m, &e);
                uv[i * 2 + j] = m;
            }
        for (b = 0; b < 16; b++) {
            mv_t m;
            int bx = (b & 3) * 4, by = (b >> 2) * 4;
            m.row = a->mv[b].row; m.col = a->mv[b].col;
            if (need_clamp) m = clamp_mv(m, &e);
            predict_block(bil, ry + by * d->y_stride + bx, d->y_stride, m.row, m.col,
                          dy + by * d->y_stride + bx, 4, 4);
        }
        for (b = 0; b < 4; b++) {
            int bx = (b & 1) * 4, by = (b >> 1) * 4;
            predict_block(bil, ru + by * d->uv_stride + bx, d->uv_stride, uv[b].row, uv[b].col,
                          du + by * d->uv_stride + bx, 4, 4);
            predict_block(bil, rv + by * d->uv_stride + bx, d->uv_stride, uv[b].row, uv[b].col,
                          dv + by * d->uv_stride + bx, 4, 4);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * A6 (intra): reconintra.c:139-263 (luma 16x16), :403-546 (chroma), reconintra4x4.c:16-296
 * ---------------------------------------------------------------------------------------- */
static void intra_predict_plane(uint8_t *dst, int stride, int size, int mode,
                                int up_avail, int left_avail)
{
    const uint8_t *above = dst - stride;
    uint8_t left[16];
    uint8_t tl = above[-1];
    int r, c, lg = size == 16 ? 3 : 2;
    for (r = 0; r < size; r++) left[r] = dst[r * stride - 1];
    switch (mode) {
    case VP8B200_DC_PRED: {
        int dc = 128;
        if (up_avail || left_avail) {
            int sum = 0, shift = lg + up_avail + left_avail;
            if (up_avail) for (c = 0; c < size; c++) sum += above[c];
            if (left_avail) for (r = 0; r < size; r++) sum += left[r];
            dc = (sum + (1 << (shift - 1))) >> shift;
        }
        for (r = 0; r < size; r++) memset(dst + r * stride, dc, (size_t)size);
        break;
    }
    case VP8B200_V_PRED:
        for (r = 0; r < size; r++) memcpy(dst + r * stride, above, (size_t)size);
        break;
    case VP8B200_H_PRED:
        for (r = 0; r < size; r++) memset(dst + r * stride, left[r], (size_t)size);
        break;
    case VP8B200_TM_PRED:
        for (r = 0; r < size; r++)
            for (c = 0; c < size; c++) dst[r * stride + c] = clamp255(left[r] + above[c] - tl);
        break;
    default:
        break;
    }
}

/* reconintra4x4.c:16-296; predicts in place like the reference call (decodframe.c:215) */
void oracle_intra4x4(uint8_t *dst, int stride, int mode)
{
    const uint8_t *A = dst - stride;          /* A[-1] top-left, A[0..7] above + above-right */
    uint8_t L[4], tl = A[-1];
    uint8_t E[9];                              /* L3 L2 L1 L0 tl A0 A1 A2 A3 */
    uint8_t a8[8];
    int r, c;
    for (r = 0; r < 4; r++) L[r] = dst[r * stride - 1];
    for (c = 0; c < 8; c++) a8[c] = A[c];
    E[0] = L[3]; E[1] = L[2]; E[2] = L[1]; E[3] = L[0]; E[4] = tl;
    E[5] = A[0]; E[6] = A[1]; E[7] = A[2]; E[8] = A[3];
#define P(r_, c_) dst[(r_) * stride + (c_)]
#define AVG3(x, y, z) (uint8_t)(((x) + 2 * (y) + (z) + 2) >> 2)
#define AVG2(x, y) (uint8_t)(((x) + (y) + 1) >> 1)
    switch (mode) {
    case VP8B200_B_DC_PRED: {
        int s = 4;
        for (c = 0; c < 4; c++) s += a8[c] + L[c];
        s >>= 3;
        for (r = 0; r < 4; r++) for (c = 0; c < 4; c++) P(r, c) = (uint8_t)s;
        break;
    }
    case VP8B200_B_TM_PRED:
        for (r = 0; r < 4; r++) for (c = 0; c < 4; c++) P(r, c) = clamp255(a8[c] - tl + L[r]);
        break;
    case VP8B200_B_VE_PRED: {
        uint8_t v[4];
        v[0] = AVG3(tl, a8[0], a8[1]); v[1] = AVG3(a8[0], a8[1], a8[2]);
        v[2] = AVG3(a8[1], a8[2], a8[3]); v[3] = AVG3(a8[2], a8[3], a8[4]);
        for (r = 0; r < 4; r++) for (c = 0; c < 4; c++) P(r, c) = v[c];
        break;
    }
    case VP8B200_B_HE_PRED: {
        uint8_t v[4];
        v[0] = AVG3(tl, L[0], L[1]); v[1] = AVG3(L[0], L[1], L[2]);
        v[2] = AVG3(L[1], L[2], L[3]); v[3] = AVG3(L[2], L[3], L[3]);
        for (r = 0; r < 4; r++) for (c = 0; c < 4; c++) P(r, c) = v[r];
        break;
    }
    case VP8B200_B_LD_PRED:
        /* anti-diagonal k = r + c uses above[k..k+2], last one repeats a8[7] */
        for (r = 0; r < 4; r++) for (c = 0; c < 4; c++) {
            int k = r + c;
            P(r, c) = k < 6 ? AVG3(a8[k], a8[k + 1], a8[k + 2]) : AVG3(a8[6], a8[7], a8[7]);
        }
        break;
    case VP8B200_B_RD_PRED:
        /* diagonal k = 3 - r + c walks the edge array E */
        for (r = 0; r < 4; r++) for (c = 0; c < 4; c++) {
            int k = 3 - r + c;
            P(r, c) = AVG3(E[k], E[k + 1], E[k + 2]);
        }
        break;
    case VP8B200_B_VR_PRED:
        P(3, 0) = AVG3(E[1], E[2], E[3]);
        P(2, 0) = AVG3(E[2], E[3], E[4]);
        P(3, 1) = P(1, 0) = AVG3(E[3], E[4], E[5]);
        P(2, 1) = P(0, 0) = AVG2(E[4], E[5]);
        P(3, 2) = P(1, 1) = AVG3(E[4], E[5], E[6]);
        P(2, 2) = P(0, 1) = AVG2(E[5], E[6]);
        P(3, 3) = P(1, 2) = AVG3(E[5], E[6], E[7]);
        P(2, 3) = P(0, 2) = AVG2(E[6], E[7]);
        P(1, 3) = AVG3(E[6], E[7], E[8]);
        P(0, 3) = AVG2(E[7], E[8]);
        break;
    case VP8B200_B_VL_PRED:
        P(0, 0) = AVG2(a8[0], a8[1]);
        P(1, 0) = AVG3(a8[0], a8[1], a8[2]);
        P(2, 0) = P(0, 1) = AVG2(a8[1], a8[2]);
        P(1, 1) = P(3, 0) = AVG3(a8[1], a8[2], a8[3]);
        P(2, 1) = P(0, 2) = AVG2(a8[2], a8[3]);
        P(3, 1) = P(1, 2) = AVG3(a8[2], a8[3], a8[4]);
        P(0, 3) = P(2, 2) = AVG2(a8[3], a8[4]);
        P(1, 3) = P(3, 2) = AVG3(a8[3], a8[4], a8[5]);
        P(2, 3) = AVG3(a8[4], a8[5], a8[6]);
        P(3, 3) = AVG3(a8[5], a8[6], a8[7]);
        break;
    case VP8B200_B_HD_PRED:
        P(3, 0) = AVG2(E[0], E[1]);
        P(3, 1) = AVG3(E[0], E[1], E[2]);
        P(2, 0) = P(3, 2) = AVG2(E[1], E[2]);
        P(2, 1) = P(3, 3) = AVG3(E[1], E[2], E[3]);
        P(2, 2) = P(1, 0) = AVG2(E[2], E[3]);
        P(2, 3) = P(1, 1) = AVG3(E[2], E[3], E[4]);
        P(1, 2) = P(0, 0) = AVG2(E[3], E[4]);
        P(1, 3) = P(0, 1) = AVG3(E[3], E[4], E[5]);
        P(0, 2) = AVG3(E[4], E[5], E[6]);
        P(0, 3) = AVG3(E[5], E[6], E[7]);
        break;
    case VP8B200_B_HU_PRED:
        P(0, 0) = AVG2(L[0], L[1]);
        P(0, 1) = AVG3(L[0], L[1], L[2]);
        P(0, 2) = P(1, 0) = AVG2(L[1], L[2]);
        P(0, 3) = P(1, 1) = AVG3(L[1], L[2], L[3]);
        P(1, 2) = P(2, 0) = AVG2(L[2], L[3]);
        P(1, 3) = P(2, 1) = AVG3(L[2], L[3], L[3]);
        P(2, 2) = P(2, 3) = P(3, 0) = P(3, 1) = P(3, 2) = P(3, 3) = L[3];
        break;
    default:
        break;
    }
#undef P
#undef AVG3
#undef AVG2
}

/* ------------------------------------------------------------------------------------------
 * one macroblock: decodframe.c:190-304 (prediction + residual)
 * ---------------------------------------------------------------------------------------- */
static void recon_mb(const oracle_dec *d, const vp8b200_frame_hdr *hdr, const vp8b200_mb *mb,
                     const vp8b200_aux *aux, const int16_t *coef, int mb_row, int mb_col)
{
    int16_t q[25][16];
    int present[25];
    const int16_t (*dqf)[2] = hdr->dequant[mb->flags & VP8B200_MBF_SEGMENT_MASK];
    int skip = (mb->flags & VP8B200_MBF_SKIP) != 0;
    uint8_t *dy = plane_y(d, hdr->fb_new) + (size_t)mb_row * 16 * d->y_stride + mb_col * 16;
    uint8_t *du = plane_u(d, hdr->fb_new) + (size_t)mb_row * 8 * d->uv_stride + mb_col * 8;
    uint8_t *dv = plane_v(d, hdr->fb_new) + (size_t)mb_row * 8 * d->uv_stride + mb_col * 8;
    int b, n = 0;

    memset(q, 0, sizeof q);
    for (b = 0; b < 25; b++) {
        present[b] = !skip && ((mb->coef_mask >> b) & 1);
        if (present[b]) memcpy(q[b], coef + ((size_t)mb->coef_off + n++) * 16, 32);
    }

    if (mb->ref_frame == VP8B200_INTRA_FRAME) {
        int up = mb_row != 0, left = mb_col != 0;
        intra_predict_plane(du, d->uv_stride, 8, mb->uv_mode, up, left);
        intra_predict_plane(dv, d->uv_stride, 8, mb->uv_mode, up, left);
        if (mb->y_mode != VP8B200_B_PRED) {
            intra_predict_plane(dy, d->y_stride, 16, mb->y_mode, up, left);
        } else {
            /* decodframe.c:200-237, with vp8_intra_prediction_down_copy
             * (reconintra4x4.c:305-317): the 4 pixels above-right of the MB are copied to
             * rows 3, 7, 11 right of the MB so blocks 7, 11, 15 see them as "above-right" */
            const vp8b200_aux *a = &aux[mb->u.aux];
            const uint8_t *ar = dy - d->y_stride + 16;
            memcpy(dy + 3 * d->y_stride + 16, ar, 4);
            memcpy(dy + 7 * d->y_stride + 16, ar, 4);
            memcpy(dy + 11 * d->y_stride + 16, ar, 4);
            for (b = 0; b < 16; b++) {
                uint8_t *blk = dy + (b >> 2) * 4 * d->y_stride + (b & 3) * 4;
                oracle_intra4x4(blk, d->y_stride, a->b_mode[b]);
                if (present[b]) dequant_idct_block(q[b], 1, dqf[0][0], dqf[0][1], blk, d->y_stride);
            }
        }
    } else {
        inter_predict_mb(d, hdr, mb, aux, mb_row, mb_col, dy, du, dv);
    }

    if (skip) return;                                       /* decodframe.c:252 */

    if (mb->y_mode != VP8B200_B_PRED) {
        int y_dc_factor = dqf[0][0];
        if (mb->y_mode != VP8B200_SPLITMV) {
            /* second-order transform, decodframe.c:259-292 */
            int16_t dq2[16], dc[16];
            int i;
            dq2[0] = (int16_t)(q[24][0] * dqf[1][0]);
            for (i = 1; i < 16; i++) dq2[i] = (int16_t)(q[24][i] * dqf[1][1]);
            oracle_iwalsh(dq2, dc);
            for (i = 0; i < 16; i++) q[i][0] = dc[i];
            y_dc_factor = 1;                                /* dequant_y1_dc, decodframe.c:92,291 */
        }
        for (b = 0; b < 16; b++)
            dequant_idct_block(q[b], present[b], y_dc_factor, dqf[0][1],
                               dy + (b >> 2) * 4 * d->y_stride + (b & 3) * 4, d->y_stride);
    }
    for (b = 0; b < 4; b++) {
        int o = (b >> 1) * 4 * d->uv_stride + (b & 1) * 4;
        dequant_idct_block(q[16 + b], present[16 + b], dqf[2][0], dqf[2][1], du + o, d->uv_stride);
        dequant_idct_block(q[20 + b], present[20 + b], dqf[2][0], dqf[2][1], dv + o, d->uv_stride);
    }
}

/* ------------------------------------------------------------------------------------------
 * A7: loop filter (loopfilter.c, loopfilter_filters.c)
 * ---------------------------------------------------------------------------------------- */
static int sc(int v) { return v < -128 ? -128 : (v > 127 ? 127 : v); }   /* :19-24 */
static int iabs(int v) { return v < 0 ? -v : v; }

/* loopfilter_filters.c:27-49: returns 1 when the edge is to be filtered */
static int lf_mask(int lim, int blim, const uint8_t *s, int st)
{
    int p3 = s[-4 * st], p2 = s[-3 * st], p1 = s[-2 * st], p0 = s[-st];
    int q0 = s[0], q1 = s[st], q2 = s[2 * st], q3 = s[3 * st];
    if (iabs(p3 - p2) > lim || iabs(p2 - p1) > lim || iabs(p1 - p0) > lim ||
        iabs(q1 - q0) > lim || iabs(q2 - q1) > lim || iabs(q3 - q2) > lim) return 0;
    return iabs(p0 - q0) * 2 + iabs(p1 - q1) / 2 <= blim;
}
static int lf_hev(int thr, const uint8_t *s, int st)
{
    return iabs(s[-2 * st] - s[-st]) > thr || iabs(s[st] - s[0]) > thr;
}

/* loopfilter_filters.c:51-97 (inner edges) */
static void lf_inner(uint8_t *s, int st, int hev)
{
    int ps1 = (int8_t)(s[-2 * st] ^ 0x80), ps0 = (int8_t)(s[-st] ^ 0x80);
    int qs0 = (int8_t)(s[0] ^ 0x80), qs1 = (int8_t)(s[st] ^ 0x80);
    int f = hev ? sc(ps1 - qs1) : 0;
    int f1, f2, u;
    f = sc(f + 3 * (qs0 - ps0));
    f1 = sc(f + 4) >> 3;
    f2 = sc(f + 3) >> 3;
    s[0] = (uint8_t)(sc(qs0 - f1) ^ 0x80);
    s[-st] = (uint8_t)(sc(ps0 + f2) ^ 0x80);
    u = (f1 + 1) >> 1;
    if (hev) u = 0;
    s[st] = (uint8_t)(sc(qs1 - u) ^ 0x80);
    s[-2 * st] = (uint8_t)(sc(ps1 + u) ^ 0x80);
}

/* loopfilter_filters.c:161-214 (macroblock edges) */
static void lf_mbedge(uint8_t *s, int st, int hev)
{
    int ps2 = (int8_t)(s[-3 * st] ^ 0x80), ps1 = (int8_t)(s[-2 * st] ^ 0x80);
    int ps0 = (int8_t)(s[-st] ^ 0x80), qs0 = (int8_t)(s[0] ^ 0x80);
    int qs1 = (int8_t)(s[st] ^ 0x80), qs2 = (int8_t)(s[2 * st] ^ 0x80);
    int f = sc(sc(ps1 - qs1) + 3 * (qs0 - ps0));
    int f2 = hev ? f : 0, f1, u;
    f1 = sc(f2 + 4) >> 3;
    f2 = sc(f2 + 3) >> 3;
    qs0 = sc(qs0 - f1);
    ps0 = sc(ps0 + f2);
    if (hev) f = 0;
    u = sc((63 + f * 27) >> 7);
    s[0] = (uint8_t)(sc(qs0 - u) ^ 0x80);
    s[-st] = (uint8_t)(sc(ps0 + u) ^ 0x80);
    u = sc((63 + f * 18) >> 7);
    s[st] = (uint8_t)(sc(qs1 - u) ^ 0x80);
    s[-2 * st] = (uint8_t)(sc(ps1 + u) ^ 0x80);
    u = sc((63 + f * 9) >> 7);
    s[2 * st] = (uint8_t)(sc(qs2 - u) ^ 0x80);
    s[-3 * st] = (uint8_t)(sc(ps2 + u) ^ 0x80);
}

/* loopfilter_filters.c:292-315 */
static void lf_simple(uint8_t *s, int st, int blim)
{
    int p1 = (int8_t)(s[-2 * st] ^ 0x80), p0 = (int8_t)(s[-st] ^ 0x80);
    int q0 = (int8_t)(s[0] ^ 0x80), q1 = (int8_t)(s[st] ^ 0x80);
    int f;
    if (iabs(s[-st] - s[0]) * 2 + iabs(s[-2 * st] - s[st]) / 2 > blim) return;
    f = sc(sc(p1 - q1) + 3 * (q0 - p0));
    s[0] = (uint8_t)(sc(q0 - (sc(f + 4) >> 3)) ^ 0x80);
    s[-st] = (uint8_t)(sc(p0 + (sc(f + 3) >> 3)) ^ 0x80);
}

/* one edge of `n` pixels; `along` = step between the n pixels, `across` = step across it */
void oracle_edge_normal(uint8_t *s, int along, int across, int n, int mbedge,
                        int elim, int ilim, int thr)
{
    int i;
    for (i = 0; i < n; i++, s += along)
        if (lf_mask(ilim, elim, s, across)) {
            int hev = lf_hev(thr, s, across);
            if (mbedge) lf_mbedge(s, across, hev); else lf_inner(s, across, hev);
        }
}
void oracle_edge_simple(uint8_t *s, int along, int across, int n, int blim)
{
    int i;
    for (i = 0; i < n; i++, s += along) lf_simple(s, across, blim);
}

/* loopfilter.c:66-96 */
static void lf_limits(int sharp, int lvl, int *ilim, int *blim, int *mblim)
{
    int il = lvl >> (sharp > 0);
    il >>= (sharp > 4);
    if (sharp > 0 && il > 9 - sharp) il = 9 - sharp;
    if (il < 1) il = 1;
    *ilim = il;
    *blim = 2 * lvl + il;
    *mblim = 2 * (lvl + 2) + il;
}

/* loopfilter.c:117-201: level for (segment, ref_frame, mode class) */
static void lf_levels(const vp8b200_frame_hdr *hdr, uint8_t lvl[4][4][4])
{
    int seg, ref, mode;
    for (seg = 0; seg < 4; seg++) {
        int base = hdr->filter_level;
        if (hdr->segmentation_enabled) {
            if (hdr->segment_abs_delta) base = hdr->segment_lf[seg];
            else {
                base += hdr->segment_lf[seg];
                base = base > 0 ? (base > 63 ? 63 : base) : 0;
            }
        }
        if (!hdr->mode_ref_lf_delta_enabled) {
            for (ref = 0; ref < 4; ref++) for (mode = 0; mode < 4; mode++) lvl[seg][ref][mode] = (uint8_t)base;
            continue;
        }
        {
            int r = base + hdr->ref_lf_deltas[0];
            int m = r + hdr->mode_lf_deltas[0];
            memset(lvl[seg][0], 0, 4);
            lvl[seg][0][0] = (uint8_t)(m > 0 ? (m > 63 ? 63 : m) : 0);
            lvl[seg][0][1] = (uint8_t)(r > 0 ? (r > 63 ? 63 : r) : 0);
        }
        for (ref = 1; ref < 4; ref++) {
            int r = base + hdr->ref_lf_deltas[ref];
            lvl[seg][ref][0] = 0;
            for (mode = 1; mode < 4; mode++) {
                int m = r + hdr->mode_lf_deltas[mode];
                lvl[seg][ref][mode] = (uint8_t)(m > 0 ? (m > 63 ? 63 : m) : 0);
            }
        }
    }
}

static const uint8_t k_mode_lf_lut[10] = {1, 1, 1, 1, 0, 2, 2, 1, 2, 3};   /* loopfilter.c:52-63 */

/* loopfilter.c:203-316 */
static void loop_filter_frame(const oracle_dec *d, const vp8b200_frame_hdr *hdr,
                              const vp8b200_mb *mbs)
{
    uint8_t lvl[4][4][4];
    int r, c;
    lf_levels(hdr, lvl);
    for (r = 0; r < d->mb_rows; r++)
        for (c = 0; c < d->mb_cols; c++) {
            const vp8b200_mb *mb = &mbs[r * d->mb_cols + c];
            int skip_lf = mb->y_mode != VP8B200_B_PRED && mb->y_mode != VP8B200_SPLITMV &&
                          (mb->flags & VP8B200_MBF_SKIP);
            int level = lvl[mb->flags & 3][mb->ref_frame][k_mode_lf_lut[mb->y_mode]];
            int ys = d->y_stride, us = d->uv_stride;
            uint8_t *y = plane_y(d, hdr->fb_new) + (size_t)r * 16 * ys + c * 16;
            uint8_t *u = plane_u(d, hdr->fb_new) + (size_t)r * 8 * us + c * 8;
            uint8_t *v = plane_v(d, hdr->fb_new) + (size_t)r * 8 * us + c * 8;
            int ilim, blim, mblim, thr, k;
            if (!level) continue;
            lf_limits(hdr->sharpness_level, level, &ilim, &blim, &mblim);
            /* loopfilter.c:28-50, frame_type 0 = key */
            if (hdr->frame_type == 0) thr = level >= 40 ? 2 : (level >= 15 ? 1 : 0);
            else thr = level >= 40 ? 3 : (level >= 20 ? 2 : (level >= 15 ? 1 : 0));
            if (hdr->filter_type == 0) {
                if (c > 0) {
                    oracle_edge_normal(y, ys, 1, 16, 1, mblim, ilim, thr);
                    oracle_edge_normal(u, us, 1, 8, 1, mblim, ilim, thr);
                    oracle_edge_normal(v, us, 1, 8, 1, mblim, ilim, thr);
                }
                if (!skip_lf) {
                    for (k = 4; k < 16; k += 4) oracle_edge_normal(y + k, ys, 1, 16, 0, blim, ilim, thr);
                    oracle_edge_normal(u + 4, us, 1, 8, 0, blim, ilim, thr);
                    oracle_edge_normal(v + 4, us, 1, 8, 0, blim, ilim, thr);
                }
                if (r > 0) {
                    oracle_edge_normal(y, 1, ys, 16, 1, mblim, ilim, thr);
                    oracle_edge_normal(u, 1, us, 8, 1, mblim, ilim, thr);
                    oracle_edge_normal(v, 1, us, 8, 1, mblim, ilim, thr);
                }
                if (!skip_lf) {
                    for (k = 4; k < 16; k += 4) oracle_edge_normal(y + k * ys, 1, ys, 16, 0, blim, ilim, thr);
                    oracle_edge_normal(u + 4 * us, 1, us, 8, 0, blim, ilim, thr);
                    oracle_edge_normal(v + 4 * us, 1, us, 8, 0, blim, ilim, thr);
                }
            } else {
                if (c > 0) oracle_edge_simple(y, ys, 1, 16, mblim);
                if (!skip_lf) for (k = 4; k < 16; k += 4) oracle_edge_simple(y + k, ys, 1, 16, blim);
                if (r > 0) oracle_edge_simple(y, 1, ys, 16, mblim);
                if (!skip_lf) for (k = 4; k < 16; k += 4) oracle_edge_simple(y + k * ys, 1, ys, 16, blim);
            }
        }
}

/* ------------------------------------------------------------------------------------------
 * A8: border handling
 * ---------------------------------------------------------------------------------------- */

/* vp8/common/setupintrarecon.c:15-32 */
static void setup_intra_recon(const oracle_dec *d, int fb)
{
    uint8_t *p[3];
    int strides[3], widths[3], heights[3], k, i;
    p[0] = plane_y(d, fb); p[1] = plane_u(d, fb); p[2] = plane_v(d, fb);
    strides[0] = d->y_stride; strides[1] = strides[2] = d->uv_stride;
    widths[0] = d->w; widths[1] = widths[2] = d->w >> 1;
    heights[0] = d->h; heights[1] = heights[2] = d->h >> 1;
    for (k = 0; k < 3; k++) {
        memset(p[k] - 1 - strides[k], 127, (size_t)widths[k] + 5);
        for (i = 0; i < heights[k]; i++) p[k][(size_t)strides[k] * i - 1] = 129;
    }
}

/* vp8/common/extend.c:160-185, as called from decodframe.c:430-433 */
static void extend_mb_row(const oracle_dec *d, int fb, int mb_row)
{
    uint8_t *p[3];
    int strides[3], k, i, line;
    p[0] = plane_y(d, fb) + (size_t)(mb_row * 16 + 14) * d->y_stride + d->w;
    p[1] = plane_u(d, fb) + (size_t)(mb_row * 8 + 6) * d->uv_stride + (d->w >> 1);
    p[2] = plane_v(d, fb) + (size_t)(mb_row * 8 + 6) * d->uv_stride + (d->w >> 1);
    strides[0] = d->y_stride; strides[1] = strides[2] = d->uv_stride;
    for (k = 0; k < 3; k++)
        for (line = 0; line < 2; line++)
            for (i = 0; i < 4; i++) p[k][line * strides[k] + i] = p[k][line * strides[k] + i - 1];
}

/* vpx_scale/generic/yv12extend.c:23-145 */
static void extend_plane(uint8_t *base, int stride, int w, int h, int border)
{
    int i;
    for (i = 0; i < h; i++) {
        uint8_t *row = base + (size_t)i * stride;
        memset(row - border, row[0], (size_t)border);
        memset(row + w, row[w - 1], (size_t)border);
    }
    for (i = 0; i < border; i++) {
        memcpy(base - border - (size_t)(i + 1) * stride, base - border, (size_t)stride);
        memcpy(base - border + (size_t)(h + i) * stride, base - border + (size_t)(h - 1) * stride,
               (size_t)stride);
    }
}

void oracle_extend_borders(oracle_dec *d, int fb)
{
    extend_plane(plane_y(d, fb), d->y_stride, d->w, d->h, 32);
    extend_plane(plane_u(d, fb), d->uv_stride, d->w >> 1, d->h >> 1, 16);
    extend_plane(plane_v(d, fb), d->uv_stride, d->w >> 1, d->h >> 1, 16);
}

void oracle_loop_filter(oracle_dec *d, const vp8b200_frame_hdr *hdr, const vp8b200_mb *mb)
{
    if (hdr->filter_level) loop_filter_frame(d, hdr, mb);
}

/* stages: bit0 = prediction+residual, bit1 = loop filter, bit2 = border extension */
void oracle_frame_stages(oracle_dec *d, const vp8b200_frame_hdr *hdr, const vp8b200_mb *mb,
                         const vp8b200_aux *aux, const int16_t *coef, int stages)
{
    int r, c;
    if (stages & 1) {
        setup_intra_recon(d, hdr->fb_new);                       /* decodframe.c:1064 */
        for (r = 0; r < d->mb_rows; r++) {                       /* decodframe.c:334-436 */
            for (c = 0; c < d->mb_cols; c++)
                recon_mb(d, hdr, &mb[r * d->mb_cols + c], aux, coef, r, c);
            extend_mb_row(d, hdr->fb_new, r);
        }
    }
    if (stages & 2) oracle_loop_filter(d, hdr, mb);              /* onyxd_if.c:576-586 */
    if (stages & 4) oracle_extend_borders(d, hdr->fb_new);       /* onyxd_if.c:607 */
}

/* the whole frame, in the reference's order (onyxd_if.c:514-607) */
void oracle_frame(oracle_dec *d, const vp8b200_frame_hdr *hdr, const vp8b200_mb *mb,
                  const vp8b200_aux *aux, const int16_t *coef)
{
    oracle_frame_stages(d, hdr, mb, aux, coef, 7);
}
