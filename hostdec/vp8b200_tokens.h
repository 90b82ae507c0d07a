/* vp8b200_tokens.h - VP8 coefficient-token decoder that writes the sparse coefficient arena
 * of include/vp8b200.h directly ("sparse coefficient packing
 * straight from vp8_decode_mb_tokens").
 *
 * It replaces, on the B200 host path, the pair
 *     vp8_decode_mb_tokens      (reference vp8/decoder/detokenize.c:183-384)
 *   + the qcoeff[400] scan / copy / memset of the per-macroblock seam
 * and produces the same records, the same entropy contexts and the same bool-decoder state.
 * No reference header is needed: the caller hands over plain pointers.
 */
#ifndef VP8B200_TOKENS_H
#define VP8B200_TOKENS_H
#include <stddef.h>
#include <stdint.h>

/* The bool decoder state, field for field what the reference keeps in BOOL_DECODER
 * (vp8/decoder/dboolhuff.h:29-36): `value` is MSB-aligned, `count` = valid bits - 8. */
typedef struct vp8b200_booldec {
    const uint8_t *buf, *buf_end;
    uint64_t value;
    int count;
    unsigned range;
} vp8b200_booldec;

/* Decodes the tokens of one macroblock.
 *   probs      coef_probs[4][8][3][11] of the frame (reference onyxc_int.h:47)
 *   above,left the 9 entropy contexts of the macroblock column / row (blockd.h:53-60:
 *              y1[4] u[2] v[2] y2), updated in place
 *   has_y2     macroblock mode is neither B_PRED nor SPLITMV
 *   coef       where this macroblock's first stored block goes (16 int16 per block, raster
 *              order inside the block, i.e. already de-zigzagged); must have room for 25
 *              blocks and be ZERO where not yet used - the function keeps that invariant
 *              (blocks it does not keep are left all-zero)
 *   mask_out   bit i set = block i stored (0..15 Y, 16..19 U, 20..23 V, 24 Y2); stored
 *              blocks follow each other in ascending bit order
 * Returns the reference's `eobtotal` (sum of end-of-block positions, minus 16 when has_y2).
 */
int vp8b200_decode_mb_tokens(vp8b200_booldec *bd, const uint8_t *probs, signed char *above,
                             signed char *left, int has_y2, int16_t *coef, uint32_t *mask_out);

#endif
