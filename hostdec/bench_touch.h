/* bench_touch.h - the frame "consumer" of both bench harnesses (hostdec/b200bench.c on our arm,
 * oracle/refbuild/refbench.c on the reference arm): reads every visible sample of a decoded
 * vpx_image_t and returns their sum.  One source for both arms so that the work is identical;
 * written so that the compiler vectorises it (a byte-at-a-time loop cost 0.8 ms per 1080p
 * frame, a quarter of our arm's whole host time per frame and 3 % of the reference's). */
#ifndef BENCH_TOUCH_H
#define BENCH_TOUCH_H
#include <stdint.h>
#include <string.h>
#include "vpx/vpx_image.h"

#if defined(__SSE2__)
#include <emmintrin.h>
static inline uint64_t touch_row(const uint8_t *r, unsigned n)
{
    /* psadbw against zero adds eight bytes per 64-bit lane (SSE2 is baseline x86-64) */
    __m128i acc = _mm_setzero_si128();
    const __m128i z = _mm_setzero_si128();
    uint64_t total;
    unsigned i = 0;
    for (; i + 16 <= n; i += 16) acc = _mm_add_epi64(acc, _mm_sad_epu8(_mm_loadu_si128((const __m128i *)(r + i)), z));
    total = (uint64_t)_mm_cvtsi128_si64(acc) + (uint64_t)_mm_cvtsi128_si64(_mm_unpackhi_epi64(acc, acc));
    for (; i < n; i++) total += r[i];
    return total;
}
#else
static inline uint64_t touch_row(const uint8_t *r, unsigned n)
{
    /* sum of n bytes: 16-bit partial sums inside 64-bit words, folded every 128 words */
    uint64_t total = 0;
    unsigned i = 0;
    while (n - i >= 8) {
        uint64_t lo = 0, hi = 0;
        unsigned k, words = (n - i) / 8;
        if (words > 128) words = 128;
        for (k = 0; k < words; k++) {
            uint64_t w;
            memcpy(&w, r + i + 8 * k, 8);
            lo += w & 0x00ff00ff00ff00ffull;
            hi += (w >> 8) & 0x00ff00ff00ff00ffull;
        }
        lo += hi;
        total += (lo & 0xffff) + ((lo >> 16) & 0xffff) + ((lo >> 32) & 0xffff) + (lo >> 48);
        i += 8 * words;
    }
    for (; i < n; i++) total += r[i];
    return total;
}
#endif

static inline uint64_t touch_image(const vpx_image_t *img)
{
    uint64_t s = 0;
    unsigned y;
    for (y = 0; y < img->d_h; y++) s += touch_row(img->planes[0] + (size_t)y * img->stride[0], img->d_w);
    for (y = 0; y < (img->d_h + 1) / 2; y++) {
        s += touch_row(img->planes[1] + (size_t)y * img->stride[1], (img->d_w + 1) / 2);
        s += touch_row(img->planes[2] + (size_t)y * img->stride[2], (img->d_w + 1) / 2);
    }
    return s;
}
#endif
