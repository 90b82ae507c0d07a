/*
 * vp8b200_recfile.h - on-disk container for per-frame macroblock records.
 *
 * A ".rec" file is exactly what the host parser hands to the C ABI (vp8b200.h), frame by
 * frame, written to disk: it lets tests, the oracle and bench.py replay a real stream's
 * reconstruction work without the bitstream parser.  Little endian, no padding surprises:
 * every struct below is a multiple of 4 bytes and only holds fixed-width fields.
 *
 *   vp8b200_rec_file_hdr
 *   repeat per decoded frame:
 *       vp8b200_rec_frame_hdr
 *       vp8b200_mb   [n_mb]
 *       vp8b200_aux  [n_aux]
 *       int16_t      [n_coef * 16]
 */
#ifndef VP8B200_RECFILE_H
#define VP8B200_RECFILE_H

#include "vp8b200.h"

#define VP8B200_REC_MAGIC       0x52385056u   /* "VP8R" */
#define VP8B200_REC_FRAME_MAGIC 0x314d5246u   /* "FRM1" */

typedef struct vp8b200_rec_file_hdr {
    uint32_t magic;
    uint32_t version;          /* = VP8B200_ABI_VERSION */
    uint32_t display_width;    /* VP8_COMMON.Width / Height: what vpxdec hashes */
    uint32_t display_height;
    uint32_t coded_width;      /* rounded up to 16 */
    uint32_t coded_height;
    uint32_t n_fb;
    uint32_t reserved;
} vp8b200_rec_file_hdr;

typedef struct vp8b200_rec_frame_hdr {
    uint32_t magic;
    uint32_t n_mb;
    uint32_t n_aux;
    uint32_t n_coef;
    uint8_t  show_frame;       /* VP8_COMMON.show_frame */
    uint8_t  fb_show;          /* index of cm->frame_to_show after swap_frame_buffers */
    uint8_t  reserved[2];
    vp8b200_frame_hdr hdr;     /* 76 bytes */
} vp8b200_rec_frame_hdr;

#endif
