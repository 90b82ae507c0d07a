/* katharness.c - frame-level known-answer drivers around the UNMODIFIED reference's own
 * reconstruction functions (oracle/_ref/libkat.so, linked against libvpxref.so).
 * TEST INFRASTRUCTURE ONLY: the reference tree has no golden
 * vectors, and stock vpxenc never emits some of the branches the hot path has (non-zero
 * per-segment loop-filter levels, far clamped / SPLITMV motion vectors next to every frame
 * edge, every intra mode at every frame edge).  These drivers take the SAME per-macroblock
 * records our C ABI takes (include/vp8b200.h), build the reference's own VP8_COMMON /
 * MODE_INFO / MACROBLOCKD / YV12_BUFFER_CONFIG around caller-provided frame allocations
 * (reference layout: border 32, yv12config.c:55-110) and call
 *   kat_loop_filter_frame : vp8_loop_filter_frame            (vp8/common/loopfilter.c:203)
 *   kat_inter_frame       : vp8_build_inter_predictors_mb    (vp8/common/reconinter.c:560)
 *   kat_intra_frame       : vp8_build_intra_predictors_mby_s / mbuv_s, vp8_intra4x4_predict,
 *                           vp8_intra_prediction_down_copy, vp8_setup_intra_recon,
 *                           vp8_extend_mb_row   (the prediction half of decodframe.c:190-238,
 *                           :343-436; no residual, so every macroblock predicts from the
 *                           predictions of its neighbours)
 * so that tests can require oracle == reference and CUDA == reference on seeded random records.
 * Only driver glue lives here; every pixel is produced by reference code.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "vpx_config.h"
#include "vpx_rtcd.h"
#include "vp8/common/onyxc_int.h"
#include "vp8/common/blockd.h"
#include "vp8/common/loopfilter.h"
#include "vp8/common/reconinter.h"
#include "vp8/common/reconintra4x4.h"
#include "vp8/common/setupintrarecon.h"
#include "vp8/common/extend.h"
#include "vp8/common/findnearmv.h"
#include "vp8b200.h"

static void wrap_fb(YV12_BUFFER_CONFIG *f, uint8_t *alloc, int w, int h)
{
    const int border = 32;
    memset(f, 0, sizeof *f);
    f->y_width = w; f->y_height = h; f->y_stride = ((w + 2 * border) + 31) & ~31;
    f->uv_width = w >> 1; f->uv_height = h >> 1; f->uv_stride = f->y_stride >> 1;
    f->border = border;
    f->buffer_alloc = alloc;
    f->frame_size = (h + 2 * border) * f->y_stride + 2 * ((h >> 1) + border) * f->uv_stride;
    f->y_buffer = alloc + border * f->y_stride + border;
    f->u_buffer = alloc + (h + 2 * border) * f->y_stride + (border / 2) * f->uv_stride + border / 2;
    f->v_buffer = f->u_buffer + ((h >> 1) + border) * f->uv_stride;
}

/* MODE_INFO array with the reference's border column (alloccommon.c:111-120) from our records */
static MODE_INFO *build_mi(const vp8b200_mb *mb, const vp8b200_aux *aux, int cols, int rows)
{
    const int stride = cols + 1;
    MODE_INFO *mip = (MODE_INFO *)calloc((size_t)(rows + 1) * stride, sizeof(MODE_INFO));
    MODE_INFO *mi = mip + stride + 1;
    int r, c, i;
    if (!mip) return NULL;
    for (r = 0; r < rows; r++)
        for (c = 0; c < cols; c++) {
            const vp8b200_mb *m = &mb[r * cols + c];
            MODE_INFO *d = &mi[r * stride + c];
            d->mbmi.mode = m->y_mode;
            d->mbmi.uv_mode = m->uv_mode;
            d->mbmi.ref_frame = m->ref_frame;
            d->mbmi.segment_id = m->flags & VP8B200_MBF_SEGMENT_MASK;
            d->mbmi.mb_skip_coeff = (m->flags & VP8B200_MBF_SKIP) != 0;
            d->mbmi.need_to_clamp_mvs = (m->flags & VP8B200_MBF_CLAMP_MVS) != 0;
            if (m->y_mode == SPLITMV) {
                const vp8b200_aux *a = &aux[m->u.aux];
                int quad_equal = 1;
                for (i = 0; i < 16; i++) {
                    d->bmi[i].mv.as_mv.row = a->mv[i].row;
                    d->bmi[i].mv.as_mv.col = a->mv[i].col;
                }
                /* the bitstream's partitioning (decodemv.c): 0-2 = every 8x8 quadrant has one MV
                 * (the reference then predicts 8x8 blocks from blocks 0, 2, 8, 10), 3 = 4x4 */
                for (i = 0; i < 16; i++) {
                    const int qd = (i & 8) | ((i & 2));          /* first block of the quadrant */
                    if (d->bmi[i].mv.as_int != d->bmi[qd].mv.as_int) quad_equal = 0;
                }
                d->mbmi.partitioning = quad_equal ? 2 : 3;
                d->mbmi.mv.as_int = d->bmi[15].mv.as_int;
            } else if (m->y_mode == B_PRED) {
                const vp8b200_aux *a = &aux[m->u.aux];
                for (i = 0; i < 16; i++) d->bmi[i].as_mode = (B_PREDICTION_MODE)a->b_mode[i];
            } else {
                d->mbmi.mv.as_mv.row = m->u.mv.row;
                d->mbmi.mv.as_mv.col = m->u.mv.col;
            }
        }
    return mip;
}

/* vp8_loop_filter_frame on `frame` (whole allocation, coded w x h) with the header's filter
 * parameters and the records' mode / ref_frame / segment / skip */
int kat_loop_filter_frame(int w, int h, const vp8b200_frame_hdr *hdr, const vp8b200_mb *mb,
                          const vp8b200_aux *aux, uint8_t *frame)
{
    VP8_COMMON *cm = (VP8_COMMON *)calloc(1, sizeof *cm);
    MACROBLOCKD *xd = (MACROBLOCKD *)calloc(1, sizeof *xd);
    YV12_BUFFER_CONFIG fb;
    MODE_INFO *mip;
    int i;
    if (!cm || !xd) return -1;
    cm->mb_cols = w >> 4; cm->mb_rows = h >> 4;
    cm->Width = w; cm->Height = h;
    cm->mode_info_stride = cm->mb_cols + 1;
    mip = build_mi(mb, aux, cm->mb_cols, cm->mb_rows);
    if (!mip) return -1;
    cm->mip = mip; cm->mi = mip + cm->mode_info_stride + 1;
    wrap_fb(&fb, frame, w, h);
    cm->frame_to_show = &fb;
    cm->frame_type = (FRAME_TYPE)hdr->frame_type;
    cm->filter_type = (LOOPFILTERTYPE)hdr->filter_type;
    cm->filter_level = hdr->filter_level;
    cm->sharpness_level = hdr->sharpness_level;
    vp8_loop_filter_init(cm);                              /* onyxd_if.c:107 */
    xd->segmentation_enabled = hdr->segmentation_enabled;
    xd->mb_segement_abs_delta = hdr->segment_abs_delta ? SEGMENT_ABSDATA : SEGMENT_DELTADATA;
    xd->mode_ref_lf_delta_enabled = hdr->mode_ref_lf_delta_enabled;
    for (i = 0; i < 4; i++) {
        xd->segment_feature_data[MB_LVL_ALT_LF][i] = hdr->segment_lf[i];
        xd->ref_lf_deltas[i] = hdr->ref_lf_deltas[i];
        xd->mode_lf_deltas[i] = hdr->mode_lf_deltas[i];
    }
    if (cm->filter_level) vp8_loop_filter_frame(cm, xd);   /* onyxd_if.c:576-586 */
    free(mip); free(cm); free(xd);
    return 0;
}

static void set_mb_position(MACROBLOCKD *xd, YV12_BUFFER_CONFIG *dst, int cols, int rows, int mb_row, int mb_col)
{
    /* decodframe.c:343-397 */
    xd->up_available = mb_row != 0;
    xd->left_available = mb_col != 0;
    xd->mb_to_top_edge = -((mb_row * 16)) << 3;
    xd->mb_to_bottom_edge = ((rows - 1 - mb_row) * 16) << 3;
    xd->mb_to_left_edge = -((mb_col * 16) << 3);
    xd->mb_to_right_edge = ((cols - 1 - mb_col) * 16) << 3;
    xd->dst.y_buffer = dst->y_buffer + mb_row * 16 * dst->y_stride + mb_col * 16;
    xd->dst.u_buffer = dst->u_buffer + mb_row * 8 * dst->uv_stride + mb_col * 8;
    xd->dst.v_buffer = dst->v_buffer + mb_row * 8 * dst->uv_stride + mb_col * 8;
}

/* vp8_build_inter_predictors_mb for every inter macroblock of the records; fb[0] = destination,
 * fb[1..3] = last / golden / altref (border-extended) */
int kat_inter_frame(int w, int h, const vp8b200_frame_hdr *hdr, const vp8b200_mb *mb,
                    const vp8b200_aux *aux, uint8_t *const fb[4])
{
    const int cols = w >> 4, rows = h >> 4;
    MACROBLOCKD *xd = (MACROBLOCKD *)calloc(1, sizeof *xd);
    YV12_BUFFER_CONFIG f[4];
    MODE_INFO *mip = build_mi(mb, aux, cols, rows), *mi;
    int r, c, i;
    if (!xd || !mip) return -1;
    for (i = 0; i < 4; i++) wrap_fb(&f[i], fb[i], w, h);
    mi = mip + cols + 2;
    memcpy(&xd->pre, &f[1], sizeof(YV12_BUFFER_CONFIG));   /* decodframe.c:1057-1068 */
    memcpy(&xd->dst, &f[0], sizeof(YV12_BUFFER_CONFIG));
    vp8_setup_block_dptrs(xd);
    vp8_build_block_doffsets(xd);
    if (!hdr->use_bilinear_mc) {                           /* decodframe.c:654-676 */
        xd->subpixel_predict = vp8_sixtap_predict4x4; xd->subpixel_predict8x4 = vp8_sixtap_predict8x4;
        xd->subpixel_predict8x8 = vp8_sixtap_predict8x8; xd->subpixel_predict16x16 = vp8_sixtap_predict16x16;
    } else {
        xd->subpixel_predict = vp8_bilinear_predict4x4; xd->subpixel_predict8x4 = vp8_bilinear_predict8x4;
        xd->subpixel_predict8x8 = vp8_bilinear_predict8x8; xd->subpixel_predict16x16 = vp8_bilinear_predict16x16;
    }
    xd->fullpixel_mask = hdr->full_pixel ? 0xfffffff8 : 0xffffffff;
    xd->mode_info_stride = cols + 1;
    for (r = 0; r < rows; r++)
        for (c = 0; c < cols; c++) {
            const YV12_BUFFER_CONFIG *ref;
            xd->mode_info_context = &mi[r * (cols + 1) + c];
            if (xd->mode_info_context->mbmi.ref_frame == INTRA_FRAME) continue;
            set_mb_position(xd, &f[0], cols, rows, r, c);
            ref = &f[xd->mode_info_context->mbmi.ref_frame];          /* decodframe.c:398-407 */
            xd->pre.y_buffer = ref->y_buffer + r * 16 * ref->y_stride + c * 16;
            xd->pre.u_buffer = ref->u_buffer + r * 8 * ref->uv_stride + c * 8;
            xd->pre.v_buffer = ref->v_buffer + r * 8 * ref->uv_stride + c * 8;
            vp8_build_inter_predictors_mb(xd);
        }
    free(mip); free(xd);
    return 0;
}

/* intra prediction of every macroblock in raster order, no residual: the prediction half of
 * decode_macroblock / decode_mb_row (decodframe.c:190-238, :343-436) with the reference's
 * frame-edge setup (vp8_setup_intra_recon) and per-row extension (vp8_extend_mb_row) */
int kat_intra_frame(int w, int h, const vp8b200_mb *mb, const vp8b200_aux *aux, uint8_t *frame)
{
    const int cols = w >> 4, rows = h >> 4;
    MACROBLOCKD *xd = (MACROBLOCKD *)calloc(1, sizeof *xd);
    YV12_BUFFER_CONFIG f;
    MODE_INFO *mip = build_mi(mb, aux, cols, rows), *mi;
    int r, c, i;
    if (!xd || !mip) return -1;
    wrap_fb(&f, frame, w, h);
    mi = mip + cols + 2;
    memcpy(&xd->dst, &f, sizeof(YV12_BUFFER_CONFIG));
    memcpy(&xd->pre, &f, sizeof(YV12_BUFFER_CONFIG));
    vp8_setup_intra_recon(&f);                             /* decodframe.c:1064 */
    vp8_setup_block_dptrs(xd);
    vp8_build_block_doffsets(xd);
    xd->mode_info_stride = cols + 1;
    for (r = 0; r < rows; r++) {
        for (c = 0; c < cols; c++) {
            xd->mode_info_context = &mi[r * (cols + 1) + c];
            set_mb_position(xd, &f, cols, rows, r, c);
            vp8_build_intra_predictors_mbuv_s(xd);
            if (xd->mode_info_context->mbmi.mode != B_PRED) {
                vp8_build_intra_predictors_mby_s(xd);
            } else {
                vp8_intra_prediction_down_copy(xd);
                for (i = 0; i < 16; i++) {
                    BLOCKD *b = &xd->block[i];
                    vp8_intra4x4_predict(*(b->base_dst) + b->dst, b->dst_stride,
                                         xd->mode_info_context->bmi[i].as_mode,
                                         *(b->base_dst) + b->dst, b->dst_stride);
                }
            }
        }
        /* decodframe.c:430: last pixels of the row replicated to the right (above-right source) */
        vp8_extend_mb_row(&f, xd->dst.y_buffer + 16, xd->dst.u_buffer + 8, xd->dst.v_buffer + 8);
    }
    free(mip); free(xd);
    return 0;
}
