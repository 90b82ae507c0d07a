/* reftest - exercises VP8_COPY_REFERENCE / VP8_SET_REFERENCE (reference vp8_dx_iface.c:611-650,
 * vp8/decoder/onyxd_if.c:161-230) through the public vpx_codec API and prints a checksum of
 * everything a caller can observe.  The SAME source is built twice: against the unmodified
 * reference (oracle/_ref/reftest_ref, by oracle/refbuild/Makefile) and against the host decoder
 * with the B200 seams (hostdec/_build/reftest_b200); tests/test_gpu_hostdec.py requires the
 * two outputs to be identical.
 *
 * usage: reftest clip.ivf     (coded size must be a multiple of 16: the control compares the
 *                              image size with the decoder's aligned buffer size)
 * Sequence: decode frames 0..3, copying the LAST / GOLDEN / ALTREF references out after each;
 * replace LAST by a synthetic picture; decode frames 4..7 (they now predict from it) and copy
 * the references out again.  One line per observation.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "vpx/vpx_decoder.h"
#include "vpx/vp8dx.h"

static uint64_t fnv(uint64_t h, const uint8_t *p, size_t n)
{
    size_t i;
    for (i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}

static uint64_t img_sum(const vpx_image_t *img)
{
    uint64_t h = 0xcbf29ce484222325ull;
    unsigned y;
    for (y = 0; y < img->d_h; y++) h = fnv(h, img->planes[0] + (size_t)y * img->stride[0], img->d_w);
    for (y = 0; y < (img->d_h + 1) / 2; y++) h = fnv(h, img->planes[1] + (size_t)y * img->stride[1], (img->d_w + 1) / 2);
    for (y = 0; y < (img->d_h + 1) / 2; y++) h = fnv(h, img->planes[2] + (size_t)y * img->stride[2], (img->d_w + 1) / 2);
    return h;
}

static void copy_refs(vpx_codec_ctx_t *dec, int w, int h, int frame)
{
    static const int kinds[3] = { VP8_LAST_FRAME, VP8_GOLD_FRAME, VP8_ALTR_FRAME };
    static const char *names[3] = { "last", "golden", "altref" };
    int k;
    for (k = 0; k < 3; k++) {
        vpx_ref_frame_t ref;
        memset(&ref, 0, sizeof ref);
        ref.frame_type = kinds[k];
        if (!vpx_img_alloc(&ref.img, VPX_IMG_FMT_I420, (unsigned)w, (unsigned)h, 1)) { fprintf(stderr, "img alloc\n"); exit(2); }
        memset(ref.img.img_data, 0x5a, (size_t)w * h * 3 / 2);
        if (vpx_codec_control(dec, VP8_COPY_REFERENCE, &ref)) {
            printf("frame %d copy %s: error %s\n", frame, names[k], vpx_codec_error(dec));
        } else {
            printf("frame %d copy %s: %016llx\n", frame, names[k], (unsigned long long)img_sum(&ref.img));
        }
        vpx_img_free(&ref.img);
    }
}

int main(int argc, char **argv)
{
    FILE *f;
    uint8_t hdr[32], fh[12], *buf = NULL;
    size_t cap = 0;
    vpx_codec_ctx_t dec;
    vpx_codec_dec_cfg_t cfg = {0};
    int frame = 0, w, h;
    if (argc < 2 || !(f = fopen(argv[1], "rb"))) { fprintf(stderr, "usage: reftest clip.ivf\n"); return 2; }
    if (fread(hdr, 1, 32, f) != 32 || memcmp(hdr, "DKIF", 4)) { fprintf(stderr, "not IVF\n"); return 2; }
    w = hdr[12] | (hdr[13] << 8); h = hdr[14] | (hdr[15] << 8);
    if ((w & 15) || (h & 15)) { fprintf(stderr, "size must be a multiple of 16\n"); return 2; }
    if (vpx_codec_dec_init(&dec, vpx_codec_vp8_dx(), &cfg, 0)) { fprintf(stderr, "init failed\n"); return 2; }
    while (frame < 8 && fread(fh, 1, 12, f) == 12) {
        size_t n = fh[0] | (fh[1] << 8) | (fh[2] << 16) | ((size_t)fh[3] << 24);
        vpx_codec_iter_t it = NULL;
        vpx_image_t *img;
        if (n > cap) { buf = (uint8_t *)realloc(buf, n); cap = n; }
        if (fread(buf, 1, n, f) != n) break;
        if (frame == 4) {
            /* replace LAST by a synthetic picture (borders come from the decoder's own extension) */
            vpx_ref_frame_t ref;
            int x, y;
            memset(&ref, 0, sizeof ref);
            ref.frame_type = VP8_LAST_FRAME;
            vpx_img_alloc(&ref.img, VPX_IMG_FMT_I420, (unsigned)w, (unsigned)h, 1);
            for (y = 0; y < h; y++) for (x = 0; x < w; x++) ref.img.planes[0][y * ref.img.stride[0] + x] = (uint8_t)(x * 3 + y * 5 + ((x ^ y) & 7) * 9);
            for (y = 0; y < h / 2; y++) for (x = 0; x < w / 2; x++) {
                ref.img.planes[1][y * ref.img.stride[1] + x] = (uint8_t)(64 + x + 2 * y);
                ref.img.planes[2][y * ref.img.stride[2] + x] = (uint8_t)(200 - x + y);
            }
            if (vpx_codec_control(&dec, VP8_SET_REFERENCE, &ref)) printf("set last: error %s\n", vpx_codec_error(&dec));
            else printf("set last: ok\n");
            vpx_img_free(&ref.img);
            copy_refs(&dec, w, h, -1);
        }
        if (vpx_codec_decode(&dec, buf, (unsigned)n, NULL, 0)) {
            printf("frame %d decode: error %s\n", frame, vpx_codec_error(&dec));
            return 1;
        }
        while ((img = vpx_codec_get_frame(&dec, &it)))
            printf("frame %d shown: %016llx\n", frame, (unsigned long long)img_sum(img));
        copy_refs(&dec, w, h, frame);
        frame++;
    }
    vpx_codec_destroy(&dec);
    fclose(f);
    free(buf);
    return 0;
}
