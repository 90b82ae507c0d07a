/* lf_packed.cuh - the VP8 loop-filter edge arithmetic on TWO pixel lines at once.
 *
 * Restates vp8/common/loopfilter_filters.c (vp8_filter_mask :27-40, vp8_hevmask :43-49,
 * vp8_filter :51-97, vp8_mbfilter :161-214, vp8_simple_filter :292-315) for registers that
 * hold two independent lines, one per 16-bit half ("16x2"): sm_100a executes packed 16-bit
 * add / min / max / 3-input max / add-then-min-then-relu and byte-wise |a-b| in ONE ALU
 * instruction, so an edge costs about half the instructions of a scalar per-line version, and
 * the integer ALU pipe (one warp instruction per two cycles per scheduler) is what bounds
 * k_loopfilter.
 *
 * Number representation.  Pixels are 0..255 in each half.  The reference works on
 * signed-char values (pixel ^ 0x80); here every intermediate is kept NON-NEGATIVE in its half
 * by a bias, so that plain 32-bit IADD / IMAD (which go to the otherwise idle FMA pipe) never
 * borrow or carry across the two lines:
 *   - comparisons x <= lim become the sign bit of (lim | 0x8000) - x, expanded to a lane mask
 *     by one sign-replicating PRMT;
 *   - clamp(v, -128, 127) + 128 is one VIADDMNMX.RELU (add a constant, min 255, max 0);
 *   - the arithmetic shifts of the reference (>> 3, >> 7, (x + 1) >> 1) are byte extractions:
 *     the value is scaled so that the wanted quotient lands in byte 1 of its half, a constant
 *     folds the bias away, and a sign-replicating PRMT pulls out byte 1 of both halves as two
 *     signed 16-bit deltas; the NEGATED deltas (q side) come from the complemented input with
 *     the rounding constant adjusted (-floor(a/n) = floor((n-1-a)/n)), not from a negation;
 *   - the delta is applied with one VIADDMNMX.RELU (pixel + delta, min 255, max 0).
 * tests/test_lf_packed.py runs exactly this source on the CPU (the primitives below have a
 * plain C twin) against the reference's own edge functions over exhaustive / random inputs.
 */
#ifndef VP8B200_LF_PACKED_CUH
#define VP8B200_LF_PACKED_CUH

#include <stdint.h>

typedef uint32_t u32;
#define K2(x) ((u32)((((u32)(x)) & 0xffffu) | ((((u32)(x)) & 0xffffu) << 16)))

#if defined(__CUDA_ARCH__)
#define LFP __device__ __forceinline__
LFP u32 lfp_ad(u32 a, u32 b) { return __vabsdiffu4(a, b); }                          /* VABSDIFF4 */
LFP u32 lfp_max2(u32 a, u32 b) { return __vmaxs2(a, b); }                           /* VIMNMX.S16x2 */
LFP u32 lfp_max3(u32 a, u32 b, u32 c) { return __vimax3_s16x2(a, b, c); }           /* VIMNMX3.S16x2 */
LFP u32 lfp_addmin(u32 a, u32 b, u32 c) { return __viaddmin_s16x2(a, b, c); }       /* VIADDMNMX.S16x2 */
LFP u32 lfp_addmax(u32 a, u32 b, u32 c) { return __viaddmax_s16x2(a, b, c); }
LFP u32 lfp_addmin_relu(u32 a, u32 b, u32 c) { return __viaddmin_s16x2_relu(a, b, c); }
/* __byte_perm masks the selector to 3 bits per nibble; the sign-replicating form needs raw PTX */
LFP u32 lfp_prmt(u32 a, u32 b, u32 sel)
{
    u32 r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
#else
/* plain C twins of the primitives (test builds only; the product is device code) */
#define LFP static inline
LFP int16_t lfp_lo(u32 v) { return (int16_t)(v & 0xffff); }
LFP int16_t lfp_hi(u32 v) { return (int16_t)(v >> 16); }
LFP u32 lfp_mk(int lo, int hi) { return ((u32)lo & 0xffffu) | (((u32)hi & 0xffffu) << 16); }
LFP u32 lfp_ad(u32 a, u32 b)
{
    u32 r = 0;
    for (int i = 0; i < 4; i++) {
        int x = (a >> (8 * i)) & 255, y = (b >> (8 * i)) & 255;
        r |= (u32)(x > y ? x - y : y - x) << (8 * i);
    }
    return r;
}
LFP int lfp_mx(int a, int b) { return a > b ? a : b; }
LFP int lfp_mn(int a, int b) { return a < b ? a : b; }
LFP u32 lfp_max2(u32 a, u32 b) { return lfp_mk(lfp_mx(lfp_lo(a), lfp_lo(b)), lfp_mx(lfp_hi(a), lfp_hi(b))); }
LFP u32 lfp_max3(u32 a, u32 b, u32 c) { return lfp_max2(lfp_max2(a, b), c); }
LFP u32 lfp_addmin(u32 a, u32 b, u32 c)
{
    return lfp_mk(lfp_mn((int16_t)(lfp_lo(a) + lfp_lo(b)), lfp_lo(c)), lfp_mn((int16_t)(lfp_hi(a) + lfp_hi(b)), lfp_hi(c)));
}
LFP u32 lfp_addmax(u32 a, u32 b, u32 c)
{
    return lfp_mk(lfp_mx((int16_t)(lfp_lo(a) + lfp_lo(b)), lfp_lo(c)), lfp_mx((int16_t)(lfp_hi(a) + lfp_hi(b)), lfp_hi(c)));
}
LFP u32 lfp_addmin_relu(u32 a, u32 b, u32 c)
{
    u32 m = lfp_addmin(a, b, c);
    return lfp_mk(lfp_mx(lfp_lo(m), 0), lfp_mx(lfp_hi(m), 0));
}
LFP u32 lfp_prmt(u32 a, u32 b, u32 sel)
{
    const uint64_t src = ((uint64_t)b << 32) | a;
    u32 r = 0;
    for (int i = 0; i < 4; i++) {
        const u32 n = (sel >> (4 * i)) & 15;
        u32 byte = (u32)(src >> (8 * (n & 7))) & 255;
        if (n & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
#endif

/* Limits of one macroblock, replicated into both halves and pre-biased for the sign test:
 *   ilimB = interior limit | 0x8000
 *   mbEB  = (2 * mblim + 1) | 0x8000     macroblock-edge limit (loopfilter.c:78-94)
 *   inEB  = (2 * blim + 1) | 0x8000      inner-edge limit, or LFP_NEVER when the macroblock has
 *                                        no inner edges (loopfilter.c:245-253)
 *   thrB  = hev threshold | 0x8000
 * 2*lim + 1 because |p0-q0|*2 + |p1-q1|/2 <= lim  <=>  4*|p0-q0| + |p1-q1| <= 2*lim + 1. */
struct LfPk { u32 ilimB, mbEB, inEB, thrB; };
#define LFP_NEVER 0x7fff7fffu

/* all-ones in the halves whose biased difference has its sign bit set (i.e. lim >= value) */
LFP u32 lfp_sign_mask(u32 s) { return lfp_prmt(s, 0u, 0xbb99u); }
/* byte 1 of each half as a signed 16-bit value */
LFP u32 lfp_byte1_s(u32 x) { return lfp_prmt(x, 0u, 0xb391u); }
LFP u32 lfp_sel(u32 m, u32 a, u32 b) { return (a & m) | (b & ~m); }
LFP u32 lfp_apply(u32 px, u32 delta) { return lfp_addmin_relu(px, delta, K2(255)); }

/* shared head of the normal filters: edge mask and "not high edge variance" mask */
LFP void lfp_masks(u32 p3, u32 p2, u32 p1, u32 p0, u32 q0, u32 q1, u32 q2, u32 q3,
                   u32 ilimB, u32 EB, u32 thrB, u32 &mask, u32 &nhev)
{
    const u32 a10 = lfp_ad(p1, p0), b10 = lfp_ad(q1, q0);
    u32 m = lfp_max3(lfp_ad(p3, p2), lfp_ad(p2, p1), a10);
    m = lfp_max3(m, b10, lfp_ad(q2, q1));
    m = lfp_max2(m, lfp_ad(q3, q2));
    const u32 E = lfp_ad(p0, q0) * 4u + lfp_ad(p1, q1);
    mask = lfp_sign_mask((ilimB - m) & (EB - E));
    nhev = lfp_sign_mask(thrB - lfp_max2(a10, b10));
}

/* clamp(p1 - q1, -128, 127) + 128 */
LFP u32 lfp_a128(u32 p1, u32 q1) { return lfp_addmin_relu(p1 + K2(384) - q1, K2(-256), K2(255)); }
/* clamp(a + 3 * (q0 - p0), -128, 127) + 128 for a128 = a + 128 */
LFP u32 lfp_f128(u32 a128, u32 p0, u32 q0)
{
    const u32 t = q0 + K2(256) - p0;
    return lfp_addmin_relu(t * 3u + a128, K2(-768), K2(255));
}

/* macroblock edge: loopfilter_filters.c:161-214 */
LFP void lfp_mbedge(u32 p3, u32 &p2, u32 &p1, u32 &p0, u32 &q0, u32 &q1, u32 &q2, u32 q3, const LfPk &P)
{
    u32 mask, nhev;
    lfp_masks(p3, p2, p1, p0, q0, q1, q2, q3, P.ilimB, P.mbEB, P.thrB, mask, nhev);
    const u32 fm = lfp_sel(mask, lfp_f128(lfp_a128(p1, q1), p0, q0), K2(128));
    const u32 g = lfp_sel(nhev, K2(128), fm);                 /* hev ? f : 0   (+128) */
    const u32 W = fm + K2(128) - g;                           /* hev ? 0 : f   (+128) */
    const u32 Wc = g + K2(127) - fm;                          /* 255 - W */
    const u32 gc = K2(255) - g;
    /* Filter2 = min(g + 3, 127) >> 3 and -Filter1 = -(min(g + 4, 127) >> 3), each + 16, times 8 */
    const u32 G2 = lfp_addmin(g, K2(3), K2(255)) & K2(0xfff8);
    const u32 Gn = lfp_addmax(gc, K2(4), K2(8)) & K2(0xfff8);
    /* u = (63 + w * 27) >> 7 (|u| <= 27: the reference's clamp of u is a no-op); exactly one of
     * Filter and u is non-zero, so both are applied with a single clamp */
    p0 = lfp_apply(p0, lfp_byte1_s(G2 * 32u + (W * 54u + 0xD57ED57Eu)));
    q0 = lfp_apply(q0, lfp_byte1_s(Gn * 32u + (Wc * 54u + 0xD5B6D5B6u)));
    p1 = lfp_apply(p1, lfp_byte1_s(W * 36u + 0xEE7EEE7Eu));
    q1 = lfp_apply(q1, lfp_byte1_s(Wc * 36u + 0xEEA4EEA4u));
    p2 = lfp_apply(p2, lfp_byte1_s(W * 18u + 0xF77EF77Eu));
    q2 = lfp_apply(q2, lfp_byte1_s(Wc * 18u + 0xF792F792u));
}

/* inner edge: loopfilter_filters.c:51-97 */
LFP void lfp_inner(u32 p3, u32 p2, u32 &p1, u32 &p0, u32 &q0, u32 &q1, u32 q2, u32 q3, const LfPk &P)
{
    u32 mask, nhev;
    lfp_masks(p3, p2, p1, p0, q0, q1, q2, q3, P.ilimB, P.inEB, P.thrB, mask, nhev);
    const u32 a = lfp_sel(nhev, K2(128), lfp_a128(p1, q1));   /* hev ? clamp(p1 - q1) : 0 */
    const u32 fm = lfp_sel(mask, lfp_f128(a, p0, q0), K2(128));
    const u32 G2 = lfp_addmin(fm, K2(3), K2(255));
    const u32 Gn = lfp_addmax(K2(255) - fm, K2(4), K2(8));
    /* u = hev ? 0 : (Filter1 + 1) >> 1, from G1 = min(f + 4, 127) + 128: u + 8 = (G1 + 8) >> 4 */
    const u32 G1 = lfp_sel(nhev, lfp_addmin(fm, K2(4), K2(255)), K2(132));
    p0 = lfp_apply(p0, lfp_byte1_s(G2 * 32u + 0xF000F000u));
    q0 = lfp_apply(q0, lfp_byte1_s(Gn * 32u + 0xF000F000u));
    p1 = lfp_apply(p1, lfp_byte1_s(G1 * 16u + 0xF880F880u));
    q1 = lfp_apply(q1, lfp_byte1_s((K2(255) - G1) * 16u + 0xF880F880u));
}

/* simple filter: loopfilter_filters.c:281-315; EB = (2 * limit + 1) | 0x8000 */
LFP void lfp_simple(u32 p1, u32 &p0, u32 &q0, u32 q1, u32 EB)
{
    const u32 E = lfp_ad(p0, q0) * 4u + lfp_ad(p1, q1);
    const u32 mask = lfp_sign_mask(EB - E);
    const u32 fm = lfp_sel(mask, lfp_f128(lfp_a128(p1, q1), p0, q0), K2(128));
    const u32 G2 = lfp_addmin(fm, K2(3), K2(255));
    const u32 Gn = lfp_addmax(K2(255) - fm, K2(4), K2(8));
    p0 = lfp_apply(p0, lfp_byte1_s(G2 * 32u + 0xF000F000u));
    q0 = lfp_apply(q0, lfp_byte1_s(Gn * 32u + 0xF000F000u));
}

/* words of two pixel rows (a = first line, b = second line, 4 pixels each) <-> four packed pairs */
LFP void lfp_unpack(u32 a, u32 b, u32 &x0, u32 &x1, u32 &x2, u32 &x3)
{
    const u32 lo = lfp_prmt(a, b, 0x5410u), hi = lfp_prmt(a, b, 0x7632u);   /* a0 a1 b0 b1 | a2 a3 b2 b3 */
    x0 = lo & 0x00ff00ffu; x1 = lfp_prmt(lo, 0u, 0x4341u);
    x2 = hi & 0x00ff00ffu; x3 = lfp_prmt(hi, 0u, 0x4341u);
}
LFP void lfp_pack(u32 x0, u32 x1, u32 x2, u32 x3, u32 &a, u32 &b)
{
    const u32 lo = x1 * 256u + x0, hi = x3 * 256u + x2;
    a = lfp_prmt(lo, hi, 0x5410u);
    b = lfp_prmt(lo, hi, 0x7632u);
}

#endif
