/* vp8b200_internal.h - shared between the runtime (runtime.cu) and the kernels. */
#ifndef VP8B200_INTERNAL_H
#define VP8B200_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>
#include "vp8b200.h"

/* Frame-buffer geometry: the reference's YV12 layout (vpx_scale/generic/yv12config.c:55-110),
 * identical for every job of a batched launch. */
struct Geo {
    int mb_cols, mb_rows;
    int width, height;          /* coded luma size */
    int y_stride, uv_stride;
    int y_off, u_off, v_off;    /* byte offset of pixel (0,0) of each plane in the allocation */
    int uv_rows_alloc;          /* height/2 + 32 */
};

/* One frame of one stream, as the kernels see it (lives in device memory). */
struct FrameJob {
    uint8_t *dst;                 /* buffer being reconstructed (hdr.fb_new)                  */
    const uint8_t *ref[4];        /* [1] last, [2] golden, [3] altref; [0] unused             */
    const vp8b200_mb *mb;
    const vp8b200_aux *aux;
    const int16_t *coef;
    unsigned long long *intra_msg; /* per-MB exported borders: 16 tagged 64-bit words          */
    const uint32_t *intra_list;   /* indices of the intra MBs, sorted by wavefront c + 2r     */
    uint8_t *lf_msg;              /* loop-filter hand-off between CTAs: 512 B per macroblock of every LF_ROWS_PER_CTA-th row */
    unsigned int epoch_intra;     /* progress values of this frame are (epoch << 13) + columns; */
    unsigned int epoch_lf;        /* each counter advances only when its kernel really runs     */
    unsigned int n_intra;         /* intra macroblocks in the frame (0 => intra kernel idle)  */
    unsigned int n_split;         /* SPLITMV macroblocks in the frame                         */
    vp8b200_frame_hdr hdr;
};

#define VP8B200_EPOCH_SHIFT 13    /* mb_cols <= 4096 < 2^13 */

void vp8b200_upload_constants();        /* filter taps -> __constant__ on the current device */
void vp8b200_upload_intra_constants();  /* B_PRED predictor table */

/* launch wrappers (kernels_*.cu); `tickets` points at two device counters owned by the
 * launching context, ticket_base = value of the counter before this launch */
void vp8b200_launch_inter(cudaStream_t s, const FrameJob *jobs, int n_jobs, const Geo &g, bool any_split);
void vp8b200_launch_intra(cudaStream_t s, const FrameJob *jobs, int n_jobs, const Geo &g,
                          unsigned int max_intra, unsigned int *ticket, unsigned int ticket_base,
                          int *n_ctas);
void vp8b200_launch_loopfilter(cudaStream_t s, const FrameJob *jobs, int n_jobs, const Geo &g,
                               unsigned int *ticket, unsigned int ticket_base, int *n_ctas);
size_t vp8b200_lf_msg_bytes(const Geo &g);   /* size of FrameJob.lf_msg */
void vp8b200_launch_border(cudaStream_t s, const FrameJob *jobs, int n_jobs, const Geo &g);

#endif
