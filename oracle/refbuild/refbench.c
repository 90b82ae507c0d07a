/* refbench - times the UNMODIFIED reference CPU decoder (oracle/_ref/libvpxref.so)
 * through its public API (vpx_codec_dec_init / vpx_codec_decode /
 * vpx_codec_get_frame; reference vpx/vpx_decoder.h, call order as in the
 * reference's vpxdec.c:985-1067).  TEST/BASELINE INFRASTRUCTURE ONLY.
 *
 * The reference snapshot has no working multithreaded decoder (SURVEY.md fact
 * 3), so host parallelism is process-level: P forked workers, worker w decodes
 * streams w, w+P, ... (each `repeat` times).  Wall time is taken from a common
 * start signal to the last worker's exit, so the figure is whole-job aggregate
 * fps on P cores.
 *
 * usage: refbench [--procs P] [--repeat R] [--frames N] [--touch] a.ivf [b.ivf ...]
 *   --frames N : decode exactly N frames of every stream (whole passes of the clip, then the
 *             first N mod F frames of one more pass) instead of R whole passes - bench.py's
 *             "--steps K" on the reference arm
 *   --touch : also read every visible output pixel (a checksum), the analogue
 *             of vpxdec writing/hashing the frame.
 * prints one JSON line.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>
#include <unistd.h>
#include <sys/wait.h>
#include "vpx/vpx_decoder.h"
#include "vpx/vp8dx.h"
#include "bench_touch.h"            /* hostdec/bench_touch.h: the same consumer as our arm */

typedef struct { uint8_t *data; size_t size; } blob_t;

static blob_t read_file(const char *path)
{
    blob_t b = {0, 0};
    FILE *f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END);
    b.size = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    b.data = (uint8_t *)malloc(b.size);
    if (fread(b.data, 1, b.size, f) != b.size) { perror("fread"); exit(2); }
    fclose(f);
    return b;
}

static uint32_t le32(const uint8_t *p)
{
    return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* decode one IVF blob; returns frames shown */
static long decode_stream(const blob_t *b, int touch, uint64_t *sum, long max_frames, long *decoded)
{
    vpx_codec_ctx_t dec;
    vpx_codec_dec_cfg_t cfg = {0};
    long shown = 0, frames_in = 0;
    size_t pos = 32;
    if (b->size < 32 || memcmp(b->data, "DKIF", 4)) { fprintf(stderr, "not IVF\n"); exit(2); }
    if (vpx_codec_dec_init(&dec, vpx_codec_vp8_dx(), &cfg, 0)) { fprintf(stderr, "init failed\n"); exit(2); }
    while (pos + 12 <= b->size && (max_frames < 0 || frames_in < max_frames)) {
        uint32_t fsz = le32(b->data + pos);
        vpx_codec_iter_t it = NULL;
        vpx_image_t *img;
        pos += 12;
        if (pos + fsz > b->size) break;
        if (vpx_codec_decode(&dec, b->data + pos, fsz, NULL, 0)) {
            fprintf(stderr, "decode error: %s\n", vpx_codec_error(&dec));
            exit(2);
        }
        pos += fsz;
        frames_in++;
        while ((img = vpx_codec_get_frame(&dec, &it))) {
            shown++;
            if (touch) *sum += touch_image(img);
        }
    }
    vpx_codec_destroy(&dec);
    if (decoded) *decoded = frames_in;
    return shown;
}

int main(int argc, char **argv)
{
    int procs = 1, repeat = 1, touch = 0, nfiles = 0, i, w;
    long frames = -1;
    const char *files[4096];
    blob_t *blobs;
    int go[2], done[2];
    double t0, t1;
    long total = 0;
    uint64_t checksum = 0;

    for (i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--procs") && i + 1 < argc) procs = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--repeat") && i + 1 < argc) repeat = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--touch")) touch = 1;
        else if (!strcmp(argv[i], "--frames") && i + 1 < argc) frames = atol(argv[++i]);
        else if (nfiles < 4096) files[nfiles++] = argv[i];
    }
    if (!nfiles) { fprintf(stderr, "usage: refbench [--procs P] [--repeat R] [--touch] a.ivf ...\n"); return 2; }
    if (procs > nfiles) procs = nfiles;
    blobs = (blob_t *)calloc(nfiles, sizeof(blob_t));
    for (i = 0; i < nfiles; i++) blobs[i] = read_file(files[i]);

    if (pipe(go) || pipe(done)) { perror("pipe"); return 2; }
    for (w = 0; w < procs; w++) {
        pid_t pid = fork();
        if (pid < 0) { perror("fork"); return 2; }
        if (pid == 0) {
            char c;
            long n = 0;
            uint64_t sum = 0;
            int r;
            close(go[1]);
            if (read(go[0], &c, 1) < 0) _exit(3);      /* wait for the start signal (EOF) */
            if (frames < 0) {
                for (r = 0; r < repeat; r++)
                    for (i = w; i < nfiles; i += procs) n += decode_stream(&blobs[i], touch, &sum, -1, NULL);
            } else {
                for (i = w; i < nfiles; i += procs) {
                    long left = frames;
                    while (left > 0) {               /* whole passes, then a partial one */
                        long in = 0;
                        n += decode_stream(&blobs[i], touch, &sum, left, &in);
                        if (in <= 0) break;
                        left -= in;
                    }
                }
            }
            {
                uint64_t msg[2];
                msg[0] = (uint64_t)n; msg[1] = sum;
                if (write(done[1], msg, sizeof msg) != sizeof msg) _exit(3);
            }
            _exit(0);
        }
    }
    close(go[0]);
    t0 = now_s();
    close(go[1]);                      /* releases every worker at once */
    for (w = 0; w < procs; w++) {
        uint64_t msg[2] = {0, 0};
        if (read(done[0], msg, sizeof msg) == sizeof msg) { total += (long)msg[0]; checksum += msg[1]; }
    }
    t1 = now_s();
    for (w = 0; w < procs; w++) { int st; wait(&st); if (!WIFEXITED(st) || WEXITSTATUS(st)) { fprintf(stderr, "worker failed\n"); return 1; } }
    printf("{\"frames\": %ld, \"wall_s\": %.6f, \"fps\": %.3f, \"procs\": %d, \"streams\": %d, \"repeat\": %d, "
           "\"touch\": %d, \"checksum\": %llu}\n",
           total, t1 - t0, total / (t1 - t0), procs, nfiles, repeat, touch, (unsigned long long)checksum);
    return 0;
}
