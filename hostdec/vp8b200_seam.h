/* vp8b200_seam.h - prototypes of the seam functions called from the patched reference
 * host decoder (see vp8b200_seam.c for the reference file:line of every call site). */
#ifndef VP8B200_SEAM_H
#define VP8B200_SEAM_H
#include <stddef.h>
#include <stdint.h>
struct VP8D_COMP;
struct macroblockd;
struct yv12_buffer_config;
void *vp8b200_seam_alloc(size_t bytes);
void  vp8b200_seam_free(void *p);
void  vp8b200_seam_destroy(struct VP8D_COMP *pbi);
void  vp8b200_seam_frame_begin(struct VP8D_COMP *pbi);
int   vp8b200_seam_decode_tokens(struct VP8D_COMP *pbi, struct macroblockd *xd);
void  vp8b200_seam_record_mb(struct VP8D_COMP *pbi, struct macroblockd *xd, unsigned int mb_idx);
int   vp8b200_seam_decode_rows(struct VP8D_COMP *pbi, struct macroblockd *xd, void (*row_fn)(void *, int, void *));
void  vp8b200_seam_mb_wait(int mb_row, int mb_col);
void  vp8b200_seam_mb_done(int mb_row, int mb_col);
void  vp8b200_seam_frame_submit(struct VP8D_COMP *pbi);
int   vp8b200_seam_show(struct VP8D_COMP *pbi, struct yv12_buffer_config *sd);
int   vp8b200_seam_wait(struct VP8D_COMP *pbi);
int   vp8b200_seam_get_raw_frame(struct VP8D_COMP *pbi, struct yv12_buffer_config *sd, int64_t *time_stamp, int64_t *time_end_stamp);
int   vp8b200_seam_flush(struct VP8D_COMP *pbi);
int   vp8b200_seam_sync_fb(struct VP8D_COMP *pbi, int idx);
int   vp8b200_seam_upload_fb(struct VP8D_COMP *pbi, int idx);
void  vp8b200_seam_copy_fb(struct VP8D_COMP *pbi, int dst_idx, int src_idx);
#endif
