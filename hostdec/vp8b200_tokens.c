/* vp8b200_tokens.c - see vp8b200_tokens.h.
 *
 * Our own implementation of the VP8 DCT-token reader (bitstream semantics: RFC 6386
 * section 13; behaviour pinned against the reference's vp8/decoder/detokenize.c:183-384 by
 * tests/test_hostdec_tokens.py, which compares whole record dumps of both paths).  What it
 * does differently from the reference, and why it is faster on the B200 host path:
 *   - coefficients go straight into the sparse arena that is DMA'd to the GPU (one 32-byte
 *     slot per non-empty block) instead of qcoeff[400] + a scan + a copy + an 800-byte
 *     memset per macroblock;
 *   - a slot is cleared only when its block turns out to be non-empty (most blocks of a P
 *     frame are an immediate end-of-block);
 *   - the 64-bit window is refilled with one big-endian load, normalisation uses clz.
 * The arithmetic (split, compare, shift) and the resulting decoder state are exactly the
 * reference's (dboolhuff.h:46-110), so mode parsing and the next macroblock continue from it.
 */
#include <string.h>
#include "vp8b200_tokens.h"

#define LOTS_OF_BITS 0x40000000         /* dboolhuff.h:27: marks "ran past the end of the data" */

/* dboolhuff.h:46-75 restated: top up `value` with as many whole bytes as fit */
static void refill(vp8b200_booldec *d)
{
    int shift = 48 - d->count;          /* bit position of the next byte */
    const size_t left = (size_t)(d->buf_end - d->buf);
    if (left >= 8) {
        uint64_t x;
        const int n = (shift >> 3) + 1;
        memcpy(&x, d->buf, 8);
        x = __builtin_bswap64(x);
        d->value |= (x >> (56 - shift)) & ~((1ull << (shift & 7)) - 1);
        d->buf += n;
        d->count += 8 * n;
        return;
    }
    {   /* tail of the partition: byte by byte, flag exhaustion like the reference */
        const int bits_left = (int)left * 8;
        const int x = shift + 8 - bits_left;
        int loop_end = 0;
        if (x >= 0) {
            d->count += LOTS_OF_BITS;
            loop_end = x;
            if (!bits_left) return;
        }
        while (shift >= loop_end) {
            d->count += 8;
            d->value |= (uint64_t)*d->buf++ << shift;
            shift -= 8;
        }
    }
}

/* the decoder state lives in locals of the caller (a struct member would be reloaded
 * after every store) */
#define RD_DECL   uint64_t value = bd->value; int count = bd->count; unsigned range = bd->range
#define RD_SYNC() do { bd->value = value; bd->count = count; bd->range = range; } while (0)
#define RD_LOAD() do { value = bd->value; count = bd->count; range = bd->range; } while (0)
#define RD_FILL() do { if (count < 0) { RD_SYNC(); refill(bd); RD_LOAD(); } } while (0)

/* one binary decision with probability p/256 of a zero (dboolhuff.h:78-119) */
#define RD_BOOL(bit, p) do {                                              \
        const unsigned split_ = 1 + (((range - 1) * (unsigned)(p)) >> 8); \
        uint64_t big_;                                                    \
        int sh_;                                                          \
        RD_FILL();                                                        \
        big_ = (uint64_t)split_ << 56;                                    \
        if (value >= big_) { range -= split_; value -= big_; (bit) = 1; } \
        else { range = split_; (bit) = 0; }                               \
        sh_ = __builtin_clz(range) - 24;                                  \
        range <<= sh_; value <<= sh_; count -= sh_;                       \
    } while (0)

static const uint8_t k_zigzag[16] = { 0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15 };
/* offset of the band of coefficient position c inside probs[type]: band * 3 contexts * 11 nodes */
static const uint8_t k_band_off[16] = { 0, 33, 66, 99, 198, 132, 165, 198, 198, 198, 198, 198, 198, 198, 198, 231 };
/* extra-bit probabilities of the value categories (RFC 6386 13.2), most significant bit first */
static const uint8_t k_cat3[] = { 173, 148, 140 };
static const uint8_t k_cat4[] = { 176, 155, 140, 135 };
static const uint8_t k_cat5[] = { 180, 157, 141, 134, 130 };
static const uint8_t k_cat6[] = { 254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129 };

/* Tokens of one 4x4 block starting at position `first` with neighbour context `ctx`.
 * Returns the end-of-block position as the reference counts it (detokenize.c:347-349; a
 * block whose last position is coded ends at 15, not 16).  `dst` is cleared on the first
 * token, so it is untouched when the block is empty. */
static inline int block_tokens(vp8b200_booldec *bd, const uint8_t *type_probs, int ctx, int first, int16_t *dst)
{
    RD_DECL;
    const uint8_t *p = type_probs + k_band_off[first] + ctx * 11;
    int c = first, bit;

    RD_BOOL(bit, p[0]);                           /* end of block right away? */
    if (!bit) { RD_SYNC(); return c; }
    memset(dst, 0, 32);
    for (;;) {
        int v;
        RD_BOOL(bit, p[1]);
        if (!bit) {                               /* DCT_0: no end-of-block test after a zero */
            if (c == 15) break;                   /* malformed input, detokenize.c:136-142 */
            c++;
            p = type_probs + k_band_off[c];
            continue;
        }
        RD_BOOL(bit, p[2]);
        if (!bit) {
            v = 1;
            p = type_probs + 11;                  /* next context: one */
        } else {
            RD_BOOL(bit, p[3]);
            if (!bit) {
                RD_BOOL(bit, p[4]);
                if (!bit) v = 2;
                else { RD_BOOL(bit, p[5]); v = 3 + bit; }
            } else {
                RD_BOOL(bit, p[6]);
                if (!bit) {
                    RD_BOOL(bit, p[7]);
                    if (!bit) { RD_BOOL(bit, 159); v = 5 + bit; }
                    else { int b1; RD_BOOL(b1, 165); RD_BOOL(bit, 145); v = 7 + 2 * b1 + bit; }
                } else {
                    const uint8_t *xp;
                    int nb, k, hi;
                    RD_BOOL(hi, p[8]);
                    RD_BOOL(bit, p[9 + hi]);
                    switch (2 * hi + bit) {
                    case 0:  xp = k_cat3; nb = 3;  v = 11; break;
                    case 1:  xp = k_cat4; nb = 4;  v = 19; break;
                    case 2:  xp = k_cat5; nb = 5;  v = 35; break;
                    default: xp = k_cat6; nb = 11; v = 67; break;
                    }
                    for (k = 0; k < nb; k++) { RD_BOOL(bit, xp[k]); v += bit << (nb - 1 - k); }
                }
            }
            p = type_probs + 22;                  /* next context: more than one */
        }
        RD_BOOL(bit, 128);                        /* sign */
        dst[k_zigzag[c]] = (int16_t)(bit ? -v : v);
        if (c == 15) break;
        c++;
        p += k_band_off[c];
        RD_BOOL(bit, p[0]);
        if (!bit) break;                          /* end of block */
    }
    RD_SYNC();
    return c;
}

int vp8b200_decode_mb_tokens(vp8b200_booldec *bd_io, const uint8_t *probs, signed char *above_,
                             signed char *left, int has_y2, int16_t *coef, uint32_t *mask_out)
{
    enum { TYPE = 8 * 3 * 11 };                   /* bytes per block type */
    /* work on a private copy: stores through the context pointers (char) could alias the
     * caller's struct and would force a reload of the state after every block */
    vp8b200_booldec state = *bd_io, *const bd = &state;
    /* ... and of the column's context entry (9 bytes, ENTROPY_CONTEXT_PLANES): in the
     * partition-parallel parser neighbouring columns of the shared array belong to other
     * threads, and touching it once per macroblock instead of twice per block keeps the
     * cache line from bouncing */
    signed char above[9], *const above_io = above_;
    int16_t y2[16];
    uint32_t mask = 0;
    int eobtotal = 0, i, eob, first = 0, have_y2 = 0;
    const uint8_t *tp = probs + 3 * TYPE;         /* type 3: Y with DC */

    memcpy(above, above_io, 9);

    if (has_y2) {                                 /* type 1: Y2, decoded first, stored last */
        eob = block_tokens(bd, probs + 1 * TYPE, above[8] + left[8], 0, y2);
        above[8] = left[8] = (signed char)(eob > 0);
        have_y2 = eob > 0;
        eobtotal = eob - 16;
        first = 1;
        tp = probs;                               /* type 0: Y after Y2 */
    }
    for (i = 0; i < 16; i++) {
        signed char *a = above + (i & 3), *l = left + (i >> 2);
        eob = block_tokens(bd, tp, *a + *l, first, coef);
        *a = *l = (signed char)(eob > first);
        if (eob > first) { mask |= 1u << i; coef += 16; }
        eobtotal += eob;
    }
    tp = probs + 2 * TYPE;                        /* type 2: chroma */
    for (i = 16; i < 24; i++) {
        signed char *a = above + 4 + ((i - 16) >> 2) * 2 + (i & 1), *l = left + 4 + ((i - 16) >> 2) * 2 + ((i >> 1) & 1);
        eob = block_tokens(bd, tp, *a + *l, 0, coef);
        *a = *l = (signed char)(eob > 0);
        if (eob > 0) { mask |= 1u << i; coef += 16; }
        eobtotal += eob;
    }
    if (have_y2) { memcpy(coef, y2, 32); mask |= 1u << 24; }
    if (bd->count < 0) refill(bd);                /* detokenize.c:377 */
    memcpy(above_io, above, 9);
    *bd_io = state;
    *mask_out = mask;
    return eobtotal;
}
