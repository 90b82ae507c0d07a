/*
 * vp8b200.h - C ABI of the B200-native VP8 reconstruction path.
 *
 * This is the drop-in boundary.  The host
 * bitstream parser (the reference's vp8_dx_iface.c / onyxd_if.c / decodframe.c /
 * decodemv.c / detokenize.c, unchanged) records one macroblock record per MB and
 * hands the frame over; every reconstruction op then runs on the GPU:
 *
 *   dequant + IDCT/WHT add      replaces vp8_dequantize_b / vp8_dequant_idct_add* /
 *                               vp8_short_idct4x4llm / vp8_dc_only_idct_add /
 *                               vp8_short_inv_walsh4x4*    (vp8/common/rtcd_defs.sh:18-35,89-108)
 *   inter prediction            replaces vp8_build_inter_predictors_mb and the
 *                               vp8_sixtap_predict* / vp8_bilinear_predict* /
 *                               vp8_copy_mem* names         (rtcd_defs.sh:111-121,172-205;
 *                                                            vp8/common/reconinter.c:560)
 *   intra prediction            replaces vp8_build_intra_predictors_mby_s / mbuv_s /
 *                               vp8_intra4x4_predict / vp8_intra_prediction_down_copy /
 *                               vp8_setup_intra_recon       (rtcd_defs.sh:123-140)
 *   loop filter                 replaces vp8_loop_filter_frame (normal + simple) and the
 *                               vp8_loop_filter_{mbv,bv,mbh,bh}, vp8_loop_filter_simple_*
 *                               names                       (rtcd_defs.sh:37-87;
 *                                                            vp8/common/loopfilter.c:203)
 *   border extension            replaces vp8_extend_mb_row (vp8/common/extend.c:160) and
 *                               vp8_yv12_extend_frame_borders_ptr
 *                                                           (vpx_scale/generic/yv12extend.c:23)
 *
 * Call sites in the reference that bind to this ABI (INTEGRATION.md shows the stubs):
 *   vp8b200_create / destroy        vp8/common/alloccommon.c:59 (vp8_alloc_frame_buffers),
 *                                   vp8/decoder/onyxd_if.c:141 (vp8dx_remove_decompressor)
 *   vp8b200_host_alloc / host_free  vpx_scale/generic/yv12config.c:92,29 (frame-buffer memory)
 *   vp8b200_frame_begin             vp8/decoder/decodframe.c:1057-1068 (before the MB loop)
 *   (host fills vp8b200_mb records) vp8/decoder/decodframe.c:190-304 (decode_macroblock tail)
 *   vp8b200_frame_submit            vp8/decoder/onyxd_if.c:560-607 (after swap_frame_buffers;
 *                                   takes the place of vp8_loop_filter_frame +
 *                                   vp8_yv12_extend_frame_borders_ptr)
 *   vp8b200_frame_fetch             vp8/decoder/onyxd_if.c:707 (vp8dx_get_raw_frame)
 *
 * Plain C, plain pointers and sizes, no CUDA/torch types.  All functions return 0 on
 * success or a negative vp8b200_status; none throws, none uses process globals, contexts
 * are independent (thread-compatible: one thread per context at a time).  There is NO CPU
 * fallback: if no CUDA device is usable vp8b200_create fails with VP8B200_ERR_NO_DEVICE.
 * Environment: VP8B200_SYNC=block makes frame_fetch sleep on an event instead of spinning
 * (many decoder threads per core).
 */
#ifndef VP8B200_H
#define VP8B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VP8B200_ABI_VERSION 1
#define VP8B200_MAX_FB      8      /* the reference uses NUM_YV12_BUFFERS = 4 (onyxc_int.h:37) */
#define VP8B200_BORDER      32     /* VP8BORDERINPIXELS (vpx_scale/yv12config.h)                */

typedef enum vp8b200_status {
    VP8B200_OK               =  0,
    VP8B200_ERR_INVALID      = -1,   /* bad argument / call order                     */
    VP8B200_ERR_NO_DEVICE    = -2,   /* no usable CUDA device (there is no CPU path)  */
    VP8B200_ERR_NOMEM        = -3,
    VP8B200_ERR_CUDA         = -4,   /* a CUDA call failed; see vp8b200_last_error()  */
    VP8B200_ERR_OVERFLOW     = -5    /* more aux / coefficient entries than capacity  */
} vp8b200_status;

/* ---- macroblock-level record (16 bytes, raster order, mb_rows*mb_cols per frame) ------- */

/* y_mode / uv_mode use the numbering of MB_PREDICTION_MODE (vp8/common/blockd.h:75-90),
 * sub-block modes that of B_PREDICTION_MODE (blockd.h:110-132), ref_frame that of
 * MV_REFERENCE_FRAME (blockd.h:147-154). */
enum { VP8B200_DC_PRED = 0, VP8B200_V_PRED, VP8B200_H_PRED, VP8B200_TM_PRED, VP8B200_B_PRED,
       VP8B200_NEARESTMV, VP8B200_NEARMV, VP8B200_ZEROMV, VP8B200_NEWMV, VP8B200_SPLITMV };
enum { VP8B200_B_DC_PRED = 0, VP8B200_B_TM_PRED, VP8B200_B_VE_PRED, VP8B200_B_HE_PRED,
       VP8B200_B_LD_PRED, VP8B200_B_RD_PRED, VP8B200_B_VR_PRED, VP8B200_B_VL_PRED,
       VP8B200_B_HD_PRED, VP8B200_B_HU_PRED };
enum { VP8B200_INTRA_FRAME = 0, VP8B200_LAST_FRAME, VP8B200_GOLDEN_FRAME, VP8B200_ALTREF_FRAME };

#define VP8B200_MBF_SEGMENT_MASK 0x03u  /* MB_MODE_INFO.segment_id                                */
#define VP8B200_MBF_SKIP         0x04u  /* MB_MODE_INFO.mb_skip_coeff AFTER decodframe.c:129      */
#define VP8B200_MBF_CLAMP_MVS    0x08u  /* MB_MODE_INFO.need_to_clamp_mvs                         */

typedef struct vp8b200_mb {
    uint8_t  y_mode;
    uint8_t  uv_mode;
    uint8_t  ref_frame;
    uint8_t  flags;        /* VP8B200_MBF_* */
    union {
        struct { int16_t row, col; } mv;   /* 16x16 MV in 1/8 pel units as decoded (mv.h:16-20) */
        uint32_t aux;                      /* B_PRED / SPLITMV: index of the vp8b200_aux entry  */
    } u;
    uint32_t coef_mask;    /* bit b (0..24): block b owns one 16-coefficient arena entry.
                              Blocks 0-15 Y (raster), 16-19 U, 20-23 V, 24 Y2.  Entries of one
                              MB are consecutive, in increasing b.  Zero when SKIP is set.      */
    uint32_t coef_off;     /* arena index (units of 16 int16) of the MB's first entry           */
} vp8b200_mb;

/* 64-byte side record for the two MB kinds that carry 16 sub-block values */
typedef union vp8b200_aux {
    struct { int16_t row, col; } mv[16];   /* SPLITMV: MODE_INFO.bmi[i].mv, unclamped           */
    uint8_t b_mode[16];                    /* B_PRED : MODE_INFO.bmi[i].as_mode                 */
    uint8_t raw[64];
} vp8b200_aux;

/* ---- frame header ------------------------------------------------------------------------ */

typedef struct vp8b200_frame_hdr {
    uint8_t  frame_type;            /* 0 = KEY_FRAME, 1 = INTER_FRAME (blockd.h:69-73)          */
    uint8_t  use_bilinear_mc;       /* VP8_COMMON.use_bilinear_mc_filter (alloccommon.c:153)    */
    uint8_t  full_pixel;            /* VP8_COMMON.full_pixel -> chroma MV &= ~7                 */
    uint8_t  filter_type;           /* 0 = NORMAL_LOOPFILTER, 1 = SIMPLE (decodframe.c:878)     */
    uint8_t  filter_level;          /* VP8_COMMON.filter_level; 0 = no loop filter              */
    uint8_t  sharpness_level;
    uint8_t  segmentation_enabled;
    uint8_t  segment_abs_delta;     /* 1 = SEGMENT_ABSDATA                                      */
    uint8_t  mode_ref_lf_delta_enabled;
    uint8_t  fb_new;                /* frame buffer being reconstructed (cm->new_fb_idx)        */
    uint8_t  fb_last, fb_golden, fb_altref;   /* reference buffers for ref_frame 1,2,3          */
    uint8_t  reserved[3];
    int8_t   segment_lf[4];         /* segment_feature_data[MB_LVL_ALT_LF][seg]                 */
    int8_t   ref_lf_deltas[4];
    int8_t   mode_lf_deltas[4];
    int16_t  dequant[4][3][2];      /* [segment][0 Y1, 1 Y2, 2 UV][0 dc, 1 ac] factors chosen by
                                       mb_init_dequantizer (decodframe.c:67-109).  When
                                       segmentation is off all four rows hold the frame's.     */
} vp8b200_frame_hdr;

/* pinned host buffers the parser fills between frame_begin and frame_submit */
typedef struct vp8b200_frame_bufs {
    vp8b200_mb  *mb;        /* mb_rows*mb_cols records                                          */
    vp8b200_aux *aux;       /* capacity aux_capacity entries                                    */
    int16_t     *coef;      /* capacity coef_capacity entries of 16 int16, raster order within
                               the block exactly as MACROBLOCKD.qcoeff (quantised, NOT yet
                               dequantised: the device multiplies by hdr.dequant)               */
    uint32_t     aux_capacity;
    uint32_t     coef_capacity;
} vp8b200_frame_bufs;

typedef struct vp8b200_ctx vp8b200_ctx;       /* one per decoder instance (VP8D_COMP)           */

/* ---- life cycle --------------------------------------------------------------------------- */

int  vp8b200_abi_version(void);
const char *vp8b200_strerror(int status);
const char *vp8b200_last_error(const vp8b200_ctx *ctx);   /* detail of the last CUDA failure   */
int  vp8b200_device_count(void);

/* width/height: coded size rounded up to 16 (VP8_COMMON: (Width+15)&~15).  Frame buffers use
 * the reference layout of vp8_yv12_alloc_frame_buffer (yv12config.c:55-110) with border 32:
 * y_stride = ((w+64)+31)&~31, uv_stride = y_stride/2, one allocation Y|U|V of
 * vp8b200_frame_size() bytes, so a whole-buffer copy is a drop-in for buffer_alloc. */
int  vp8b200_create(vp8b200_ctx **out, int device, int width, int height, int n_fb);
void vp8b200_destroy(vp8b200_ctx *ctx);
size_t vp8b200_frame_size(const vp8b200_ctx *ctx);
int  vp8b200_y_stride(const vp8b200_ctx *ctx);

/* page-locked host memory for the decoder's own YV12 mirrors (so frame_fetch is one DMA) */
void *vp8b200_host_alloc(size_t bytes);
void *vp8b200_host_alloc_on(int device, size_t bytes);   /* same, after selecting `device` */
void  vp8b200_host_free(void *p);

/* ---- per frame ---------------------------------------------------------------------------- */

int vp8b200_frame_begin(vp8b200_ctx *ctx, const vp8b200_frame_hdr *hdr, vp8b200_frame_bufs *bufs);
/* Queue H2D of the records + every reconstruction kernel of the frame on the context's
 * stream and return without waiting.  n_aux / n_coef = entries actually written. */
int vp8b200_frame_submit(vp8b200_ctx *ctx, uint32_t n_aux, uint32_t n_coef);
/* Coalesced submit (SURVEY 8b "shared batch scheduler across ctxs", 8f N2): hand the frame to the
 * per-device engine and return.  One engine thread per device gathers the frames that the
 * decoder threads of all contexts have queued (bounded wait: VP8B200_BATCH_WINDOW_US, default
 * 1000, until about half of the live contexts have a frame; a lone context is issued at once;
 * VP8B200_BATCH_MAX frames at most) and issues ONE launch of each kernel over all of them, their
 * record uploads before and - for show_fb >= 0 - the device->host copy of that frame buffer
 * after (display_w/h as in vp8b200_frame_fetch_begin).  The caller's thread makes no CUDA
 * launch; vp8b200_frame_fetch_wait is where it waits.  Errors of the issue are reported by the
 * next call on the context. */
int vp8b200_frame_submit_show(vp8b200_ctx *ctx, uint32_t n_aux, uint32_t n_coef,
                              int show_fb, uint8_t *dst, int display_w, int display_h);
/* [0] batches, [1] frames the engine of `device` has issued */
void vp8b200_engine_stats(int device, uint64_t out[2]);
/* Abandon a frame opened by frame_begin (the reference's longjmp error path). */
int vp8b200_frame_abort(vp8b200_ctx *ctx);
/* Wait for frame buffer `fb` and copy the whole allocation (borders included) to `dst`. */
int vp8b200_frame_fetch(vp8b200_ctx *ctx, int fb, uint8_t *dst, size_t bytes);
/* Lazy fetch (binds at vp8dx_get_raw_frame, onyxd_if.c:707-745, and
 * vp8_get_frame, vp8_dx_iface.c:485-503).  fetch_begin queues the device->host copy of frame
 * buffer `fb` behind the frame's kernels and returns at once; fetch_wait blocks until the
 * pixels are in host memory.  `dst` is the host image of the WHOLE allocation (the decoder's
 * YV12 mirror, buffer_alloc): with display_w/h > 0 only the visible samples are copied
 * (display_w x display_h luma, ((w+1)/2) x ((h+1)/2) chroma - what vpx_codec_get_frame's
 * caller may read, vpxdec.c:1093-1115), each to the offset it has in the allocation;
 * display_w == display_h == 0 copies the whole allocation, borders included.  Up to two copies
 * may be in flight per context; fetch_wait collects the OLDEST one (frame-delay mode queues
 * picture N before it collects picture N-1), a third fetch_begin first waits for the oldest. */
int vp8b200_frame_fetch_begin(vp8b200_ctx *ctx, int fb, uint8_t *dst, int display_w, int display_h);
int vp8b200_frame_fetch_wait(vp8b200_ctx *ctx);
/* Upload a whole frame buffer (VP8_SET_REFERENCE, onyxd_if.c:161-230). */
int vp8b200_frame_upload(vp8b200_ctx *ctx, int fb, const uint8_t *src, size_t bytes);
/* Device-side copy fb_src -> fb_dst (vp8_yv12_copy_frame_ptr call sites, onyxd_if.c:186,390). */
int vp8b200_frame_copy(vp8b200_ctx *ctx, int fb_dst, int fb_src);
int vp8b200_sync(vp8b200_ctx *ctx);

/* ---- resident frames and batched replay (many independent streams per launch) ------------- */

/* A staged frame = header + records copied into device memory owned by the context, to be
 * reconstructed later (or repeatedly) without any host traffic.  Used by the batched
 * multi-stream driver and by bench.py's HBM-resident measurement. */
typedef struct vp8b200_staged vp8b200_staged;
int  vp8b200_stage_frame(vp8b200_ctx *ctx, const vp8b200_frame_hdr *hdr,
                         const vp8b200_mb *mb, const vp8b200_aux *aux, uint32_t n_aux,
                         const int16_t *coef, uint32_t n_coef, vp8b200_staged **out);
void vp8b200_staged_free(vp8b200_ctx *ctx, vp8b200_staged *s);
/* Reconstruct frame[i] on ctx[i] for i < n with ONE launch of each kernel covering all n
 * streams (all contexts must share device and geometry).  Asynchronous on ctx[0]'s stream
 * (the "leader").  Ordering is handled by the library: the batch waits for work the members
 * queued on their own streams, and a member's later frame_submit / frame_fetch / sync waits
 * for the batch.  A staged frame may be replayed any number of times and on any context of
 * the same geometry. */
int  vp8b200_batch_run(vp8b200_ctx *const *ctx, vp8b200_staged *const *frame, int n);

/* process-wide monotonic counters: [0] bytes copied host->device, [1] device->host,
 * [2] kernels launched, [3] frames reconstructed */
void vp8b200_global_stats(uint64_t out[4]);
/* per-kernel device timing (CUDA events around every launch on the context's stream):
 * kind 0 inter prediction+residual, 1 intra wavefront, 2 loop filter, 3 border extension.
 * profile_read waits for the stream, returns accumulated ms / launch counts and resets. */
int  vp8b200_profile_enable(vp8b200_ctx *ctx, int enable);
int  vp8b200_profile_read(vp8b200_ctx *ctx, double ms[4], uint64_t count[4]);

/* number of kernels this library has launched on the context (for bench accounting) */
uint64_t vp8b200_launch_count(const vp8b200_ctx *ctx);
/* the CUDA stream (cudaStream_t) the context launches on, for event timing by the caller */
void *vp8b200_stream(const vp8b200_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* VP8B200_H */
