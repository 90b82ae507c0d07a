/*
 * vp8b200_seam.c - glue compiled INTO the reference's host decoder (a scratch copy of the
 * reference sources patched by hostdec/apply_seams.py).  It is the reference-side binding
 * of include/vp8b200.h: it turns the parser's per-macroblock state (MODE_INFO, qcoeff,
 * eobs) into vp8b200_mb records and drives the C ABI at the frame-level seams.
 *
 * Our own code; it includes reference headers at build time only.  Nothing here
 * reconstructs pixels: with the seams applied, the reference's prediction / IDCT / loop
 * filter / border functions are never called on the decode path.
 *
 * Seams (reference file:line -> function here):
 *   vp8/decoder/decodframe.c:1064  vp8_setup_intra_recon      -> vp8b200_seam_frame_begin
 *   vp8/decoder/decodframe.c:127   vp8_decode_mb_tokens       -> vp8b200_seam_decode_tokens
 *   vp8/decoder/decodframe.c:191   "do prediction" .. :304    -> vp8b200_seam_record_mb
 *   vp8/decoder/decodframe.c:430   vp8_extend_mb_row          -> (dropped; device rule)
 *   vp8/decoder/onyxd_if.c:576-607 loop filter + extend       -> vp8b200_seam_frame_submit
 *   vp8/decoder/onyxd_if.c:729     *sd = *frame_to_show       -> vp8b200_seam_fetch
 *   vp8/decoder/onyxd_if.c:390     vp8_yv12_copy_frame_ptr    -> vp8b200_seam_copy_fb
 *   vp8/decoder/onyxd_if.c:155     vp8_remove_common          -> vp8b200_seam_destroy
 *   vpx_scale/generic/yv12config.c:92,27  vpx_memalign/free   -> vp8b200_seam_alloc/free
 *
 * Environment:
 *   VP8B200_DEVICE=<n>     CUDA device ordinal (default 0)
 *   VP8B200_DUMP=<path>    also write every frame's records to a .rec file
 *                          (include/vp8b200_recfile.h); "%p" in the path -> decoder address
 *   VP8B200_TOKENS=ref     keep the reference's vp8_decode_mb_tokens + qcoeff scan instead of
 *                          the fused token reader (vp8b200_tokens.c); for A/B tests
 *   VP8B200_NO_DEVICE=1    record-capture only: no device is touched and NO pixels are
 *                          produced (frames handed back are undefined).  Exists so that
 *                          golden .rec fixtures can be produced on a machine without a GPU;
 *                          it is not a decode path.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vpx_config.h"
#include "vp8/decoder/onyxd_int.h"
#include "vp8/common/onyxc_int.h"
#include "vp8/common/blockd.h"
#include "vpx/internal/vpx_codec_internal.h"

#include "vp8b200.h"
#include "vp8b200_recfile.h"
#include "vp8b200_seam.h"
#include "vp8b200_tokens.h"
#include "vp8/decoder/detokenize.h"

typedef struct seam_state {
    vp8b200_ctx *ctx;
    int width, height;            /* coded size of ctx */
    int no_device;
    FILE *dump;
    int dump_hdr_written;
    /* current frame */
    int open;
    vp8b200_frame_hdr hdr;
    vp8b200_frame_bufs bufs;
    uint32_t n_aux, n_coef;
    int overflow;
    int ref_tokens;               /* VP8B200_TOKENS=ref */
    int tok_valid;                /* tok_off/tok_mask describe the current macroblock */
    uint32_t tok_off, tok_mask;
    /* host-memory record buffers for VP8B200_NO_DEVICE */
    vp8b200_mb *h_mb; vp8b200_aux *h_aux; int16_t *h_coef;
} seam_state;

static void seam_fail(VP8D_COMP *pbi, const char *what, int status)
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    vpx_internal_error(&pbi->common.error, VPX_CODEC_ERROR, "vp8b200: %s: %s (%s)", what,
                       vp8b200_strerror(status),
                       s && s->ctx ? vp8b200_last_error(s->ctx) : "no context");
}

static seam_state *seam_get(VP8D_COMP *pbi)
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    if (!s) {
        const char *e;
        s = (seam_state *)calloc(1, sizeof *s);
        if (!s) vpx_internal_error(&pbi->common.error, VPX_CODEC_MEM_ERROR, "vp8b200: seam state");
        e = getenv("VP8B200_NO_DEVICE");
        s->no_device = e && atoi(e);
        e = getenv("VP8B200_TOKENS");
        s->ref_tokens = e && !strcmp(e, "ref");
        e = getenv("VP8B200_DUMP");
        if (e && *e) {
            char path[1024];
            const char *pp = strstr(e, "%p");
            if (pp) snprintf(path, sizeof path, "%.*s%p%s", (int)(pp - e), e, (void *)pbi, pp + 2);
            else snprintf(path, sizeof path, "%s", e);
            s->dump = fopen(path, "wb");
            if (!s->dump) fprintf(stderr, "vp8b200: cannot open dump file %s\n", path);
        }
        pbi->b200_seam = s;
    }
    return s;
}

void *vp8b200_seam_alloc(size_t bytes)
{
    const char *e = getenv("VP8B200_NO_DEVICE");
    if (e && atoi(e)) {
        void *p = NULL;
        return posix_memalign(&p, 64, bytes) ? NULL : p;
    }
    return vp8b200_host_alloc(bytes);
}

void vp8b200_seam_free(void *p)
{
    const char *e = getenv("VP8B200_NO_DEVICE");
    if (e && atoi(e)) free(p); else vp8b200_host_free(p);
}

void vp8b200_seam_destroy(VP8D_COMP *pbi)
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    if (!s) return;
    if (s->ctx) vp8b200_destroy(s->ctx);
    if (s->dump) fclose(s->dump);
    free(s->h_mb); free(s->h_aux); free(s->h_coef);
    free(s);
    pbi->b200_seam = NULL;
}

/* QIndex of a segment: mb_init_dequantizer, vp8/decoder/decodframe.c:67-87 */
static int segment_qindex(const VP8_COMMON *pc, const MACROBLOCKD *xd, int seg)
{
    int q = pc->base_qindex;
    if (xd->segmentation_enabled) {
        if (xd->mb_segement_abs_delta == SEGMENT_ABSDATA)
            q = xd->segment_feature_data[MB_LVL_ALT_Q][seg];
        else {
            q += xd->segment_feature_data[MB_LVL_ALT_Q][seg];
            q = q >= 0 ? (q <= MAXQ ? q : MAXQ) : 0;
        }
    }
    return q & 127;
}

void vp8b200_seam_frame_begin(VP8D_COMP *pbi)
{
    VP8_COMMON *pc = &pbi->common;
    MACROBLOCKD *xd = &pbi->mb;
    seam_state *s = seam_get(pbi);
    int w = pc->mb_cols * 16, h = pc->mb_rows * 16, seg, i, st;
    uint32_t n_mb = (uint32_t)(pc->mb_rows * pc->mb_cols);
    vp8b200_frame_hdr *hd = &s->hdr;

    if (s->open && s->ctx) vp8b200_frame_abort(s->ctx);   /* frame abandoned by a longjmp */
    s->open = 0;

    if (s->width != w || s->height != h) {                /* first frame or size change */
        if (s->ctx) { vp8b200_destroy(s->ctx); s->ctx = NULL; }
        free(s->h_mb); free(s->h_aux); free(s->h_coef);
        s->h_mb = NULL; s->h_aux = NULL; s->h_coef = NULL;
        if (!s->no_device) {
            const char *e = getenv("VP8B200_DEVICE");
            st = vp8b200_create(&s->ctx, e ? atoi(e) : 0, w, h, NUM_YV12_BUFFERS);
            if (st) seam_fail(pbi, "vp8b200_create", st);
            if ((int)vp8b200_y_stride(s->ctx) != pc->yv12_fb[0].y_stride ||
                vp8b200_frame_size(s->ctx) != (size_t)pc->yv12_fb[0].frame_size)
                vpx_internal_error(&pc->error, VPX_CODEC_ERROR, "vp8b200: frame layout mismatch");
        } else {
            s->h_mb = (vp8b200_mb *)malloc(sizeof(vp8b200_mb) * n_mb);
            s->h_aux = (vp8b200_aux *)malloc(sizeof(vp8b200_aux) * n_mb);
            s->h_coef = (int16_t *)malloc((size_t)32 * 25 * n_mb);
        }
        s->width = w; s->height = h;
        if (s->dump && !s->dump_hdr_written) {
            vp8b200_rec_file_hdr fh;
            memset(&fh, 0, sizeof fh);
            fh.magic = VP8B200_REC_MAGIC; fh.version = VP8B200_ABI_VERSION;
            fh.display_width = (uint32_t)pc->Width; fh.display_height = (uint32_t)pc->Height;
            fh.coded_width = (uint32_t)w; fh.coded_height = (uint32_t)h;
            fh.n_fb = NUM_YV12_BUFFERS;
            fwrite(&fh, sizeof fh, 1, s->dump);
            s->dump_hdr_written = 1;
        }
    }

    memset(hd, 0, sizeof *hd);
    hd->frame_type = (uint8_t)pc->frame_type;
    hd->use_bilinear_mc = (uint8_t)(pc->use_bilinear_mc_filter != 0);
    hd->full_pixel = (uint8_t)(pc->full_pixel != 0);
    hd->filter_type = (uint8_t)pc->filter_type;
    hd->filter_level = (uint8_t)pc->filter_level;
    hd->sharpness_level = (uint8_t)pc->sharpness_level;
    hd->segmentation_enabled = (uint8_t)(xd->segmentation_enabled != 0);
    hd->segment_abs_delta = (uint8_t)(xd->mb_segement_abs_delta == SEGMENT_ABSDATA);
    hd->mode_ref_lf_delta_enabled = (uint8_t)(xd->mode_ref_lf_delta_enabled != 0);
    hd->fb_new = (uint8_t)pc->new_fb_idx;
    hd->fb_last = (uint8_t)pc->lst_fb_idx;
    hd->fb_golden = (uint8_t)pc->gld_fb_idx;
    hd->fb_altref = (uint8_t)pc->alt_fb_idx;
    for (i = 0; i < 4; i++) {
        hd->segment_lf[i] = xd->segment_feature_data[MB_LVL_ALT_LF][i];
        hd->ref_lf_deltas[i] = xd->ref_lf_deltas[i];
        hd->mode_lf_deltas[i] = xd->mode_lf_deltas[i];
    }
    for (seg = 0; seg < 4; seg++) {
        int q = segment_qindex(pc, xd, seg);
        hd->dequant[seg][0][0] = pc->Y1dequant[q][0]; hd->dequant[seg][0][1] = pc->Y1dequant[q][1];
        hd->dequant[seg][1][0] = pc->Y2dequant[q][0]; hd->dequant[seg][1][1] = pc->Y2dequant[q][1];
        hd->dequant[seg][2][0] = pc->UVdequant[q][0]; hd->dequant[seg][2][1] = pc->UVdequant[q][1];
    }

    if (s->ctx) {
        st = vp8b200_frame_begin(s->ctx, hd, &s->bufs);
        if (st) seam_fail(pbi, "vp8b200_frame_begin", st);
    } else {
        s->bufs.mb = s->h_mb; s->bufs.aux = s->h_aux; s->bufs.coef = s->h_coef;
        s->bufs.aux_capacity = n_mb; s->bufs.coef_capacity = 25 * n_mb;
    }
    s->n_aux = 0; s->n_coef = 0; s->overflow = 0; s->tok_valid = 0;
    s->open = 1;
}

/* Called from decode_macroblock (vp8/decoder/decodframe.c:127) in place of
 * vp8_decode_mb_tokens: reads the macroblock's tokens straight into the coefficient arena.
 * Returns eobtotal like the function it replaces. */
int vp8b200_seam_decode_tokens(VP8D_COMP *pbi, MACROBLOCKD *xd)
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    BOOL_DECODER *bc = xd->current_bc;
    const int mode = xd->mode_info_context->mbmi.mode;
    int16_t scratch[25 * 16];
    int16_t *dst;
    vp8b200_booldec bd;
    uint32_t mask = 0;
    int eobtotal;

    if (!s || !s->open || s->ref_tokens) return vp8_decode_mb_tokens(pbi, xd);
    if (s->n_coef + 25 > s->bufs.coef_capacity) { s->overflow = 1; dst = scratch; }
    else dst = s->bufs.coef + (size_t)s->n_coef * 16;
    bd.buf = bc->user_buffer; bd.buf_end = bc->user_buffer_end;
    bd.value = bc->value; bd.count = bc->count; bd.range = bc->range;
    eobtotal = vp8b200_decode_mb_tokens(&bd, &pbi->common.fc.coef_probs[0][0][0][0],
                                        (signed char *)xd->above_context, (signed char *)xd->left_context,
                                        mode != B_PRED && mode != SPLITMV, dst, &mask);
    bc->user_buffer = bd.buf; bc->value = bd.value; bc->count = bd.count; bc->range = bd.range;
    s->tok_off = s->n_coef;
    s->tok_mask = s->overflow ? 0 : mask;
    if (!s->overflow) s->n_coef += (uint32_t)__builtin_popcount(mask);
    s->tok_valid = 1;
    return eobtotal;
}

/* Called from decode_macroblock (vp8/decoder/decodframe.c) once tokens are decoded and
 * mb_skip_coeff has its final value; replaces everything from "do prediction" to the end
 * of the function.  Also clears the coefficients it consumed, as the reference's
 * dequant/IDCT functions do (dequantize.c:41, idct_blk.c:35), because the token decoder
 * relies on an all-zero qcoeff[] at the start of every macroblock. */
void vp8b200_seam_record_mb(VP8D_COMP *pbi, MACROBLOCKD *xd, unsigned int mb_idx)
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    const MODE_INFO *mi = xd->mode_info_context;
    const MB_MODE_INFO *mbmi = &mi->mbmi;
    vp8b200_mb *r = &s->bufs.mb[mb_idx];
    int mode = mbmi->mode, i;
    int skip = mbmi->mb_skip_coeff != 0;
    int has_y2 = mode != B_PRED && mode != SPLITMV;

    r->y_mode = (uint8_t)mode;
    r->uv_mode = (uint8_t)mbmi->uv_mode;
    r->ref_frame = (uint8_t)mbmi->ref_frame;
    r->flags = (uint8_t)((mbmi->segment_id & 3) | (skip ? VP8B200_MBF_SKIP : 0) |
                         (mbmi->need_to_clamp_mvs ? VP8B200_MBF_CLAMP_MVS : 0));
    r->coef_mask = 0;
    r->coef_off = s->n_coef;

    if (mode == B_PRED || mode == SPLITMV) {
        if (s->n_aux >= s->bufs.aux_capacity) { s->overflow = 1; return; }
        r->u.aux = s->n_aux;
        {
            vp8b200_aux *a = &s->bufs.aux[s->n_aux++];
            if (mode == B_PRED) {
                memset(a, 0, sizeof *a);
                for (i = 0; i < 16; i++) a->b_mode[i] = (uint8_t)mi->bmi[i].as_mode;
            } else {
                for (i = 0; i < 16; i++) {
                    a->mv[i].row = mi->bmi[i].mv.as_mv.row;
                    a->mv[i].col = mi->bmi[i].mv.as_mv.col;
                }
            }
        }
    } else {
        r->u.mv.row = mbmi->mv.as_mv.row;
        r->u.mv.col = mbmi->mv.as_mv.col;
    }

    if (!s->ref_tokens) {
        /* the fused token reader already stored this macroblock's blocks */
        if (s->tok_valid) { r->coef_off = s->tok_off; r->coef_mask = s->tok_mask; }
        s->tok_valid = 0;
    } else if (!skip) {
        /* eobs semantics: detokenize.c:183-384.  Y blocks of a Y2 macroblock start at
         * position 1, so they carry coefficients only when eob > 1. */
        uint32_t mask = 0;
        for (i = 0; i < 25; i++) {
            int eob = xd->eobs[i];
            int present;
            if (i == 24 && !has_y2) continue;
            present = (i < 16 && has_y2) ? eob > 1 : eob > 0;
            if (!present) continue;
            if (s->n_coef >= s->bufs.coef_capacity) { s->overflow = 1; break; }
            memcpy(s->bufs.coef + (size_t)s->n_coef * 16, xd->qcoeff + i * 16, 32);
            s->n_coef++;
            mask |= 1u << i;
        }
        r->coef_mask = mask;
        memset(xd->qcoeff, 0, sizeof(xd->qcoeff));
    }
}

static void seam_dump_frame(seam_state *s, VP8D_COMP *pbi, uint32_t n_mb)
{
    VP8_COMMON *cm = &pbi->common;
    vp8b200_rec_frame_hdr fh;
    memset(&fh, 0, sizeof fh);
    fh.magic = VP8B200_REC_FRAME_MAGIC;
    fh.n_mb = n_mb; fh.n_aux = s->n_aux; fh.n_coef = s->n_coef;
    fh.show_frame = (uint8_t)cm->show_frame;
    fh.fb_show = (uint8_t)(cm->frame_to_show - cm->yv12_fb);
    fh.hdr = s->hdr;
    fwrite(&fh, sizeof fh, 1, s->dump);
    fwrite(s->bufs.mb, sizeof(vp8b200_mb), n_mb, s->dump);
    fwrite(s->bufs.aux, sizeof(vp8b200_aux), s->n_aux, s->dump);
    fwrite(s->bufs.coef, 32, s->n_coef, s->dump);
    fflush(s->dump);
}

/* after swap_frame_buffers (onyxd_if.c:560): takes the place of vp8_loop_filter_frame
 * (onyxd_if.c:576-586) and vp8_yv12_extend_frame_borders_ptr (:607) */
void vp8b200_seam_frame_submit(VP8D_COMP *pbi)
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    VP8_COMMON *cm = &pbi->common;
    int st;
    if (!s || !s->open) return;
    if (s->overflow)
        vpx_internal_error(&cm->error, VPX_CODEC_ERROR, "vp8b200: record arena overflow");
    if (s->dump) seam_dump_frame(s, pbi, (uint32_t)(cm->mb_rows * cm->mb_cols));
    s->open = 0;
    if (s->ctx) {
        st = vp8b200_frame_submit(s->ctx, s->n_aux, s->n_coef);
        if (st) seam_fail(pbi, "vp8b200_frame_submit", st);
    }
}

/* vp8dx_get_raw_frame (onyxd_if.c:707-745): make the host mirror of the shown buffer valid */
void vp8b200_seam_fetch(VP8D_COMP *pbi)
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    VP8_COMMON *cm = &pbi->common;
    int st;
    if (!s || !s->ctx || !cm->frame_to_show) return;
    st = vp8b200_frame_fetch(s->ctx, (int)(cm->frame_to_show - cm->yv12_fb),
                             cm->frame_to_show->buffer_alloc, (size_t)cm->frame_to_show->frame_size);
    if (st) seam_fail(pbi, "vp8b200_frame_fetch", st);
}

/* onyxd_if.c:390: the missing-frame path moves `last` to its own buffer */
void vp8b200_seam_copy_fb(VP8D_COMP *pbi, int dst_idx, int src_idx)
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    int st;
    if (!s || !s->ctx) return;
    st = vp8b200_frame_copy(s->ctx, dst_idx, src_idx);
    if (st) seam_fail(pbi, "vp8b200_frame_copy", st);
}
