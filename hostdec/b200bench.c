/* b200bench - multi-stream end-to-end driver over the reference's PUBLIC decoder API
 * (vpx_codec_dec_init / vpx_codec_decode / vpx_codec_get_frame, reference
 * vpx/vpx_decoder.h; call order of vpxdec.c:985-1067), linked against the host decoder with
 * the B200 seams (libvpx_b200.so).  Input: IVF files in host memory; output: vpx_image_t in
 * host memory.  So every timed frame pays host parse + H2D of the records + device
 * reconstruction + D2H of the frame - this is bench.py's "e2e" figure.
 *
 * S decoder instances (instance i plays file i % nfiles), T worker threads (thread t owns
 * instances t, t+T, ...; it advances them round-robin one frame at a time).  Each worker
 * first creates its decoders and decodes one warm-up pass (device allocations, clocks),
 * then all workers meet at a barrier and decode `repeat` timed passes of their clips.
 *
 * usage: b200bench [--threads T] [--streams S] [--repeat R] [--touch] [--sum] [--pipeline] [--delay] a.ivf [b.ivf ...]
 *   --touch : read every visible pixel of EVERY frame of every instance (a byte sum, printed as
 *           "checksum") - exactly what oracle/_ref/refbench --touch does on the reference arm,
 *           so both arms of bench.py consume their frames the same way
 *   --pipeline : a worker collects an instance's frame (vpx_codec_get_frame) only after it has
 *           parsed the next frame of its other instances.  vpx_codec_decode only queues the
 *           device work (the wait sits in vpx_codec_get_frame), so a worker with several
 *           instances never waits for the device
 *   --delay : opt-in frame-delay mode of the decoder (VP8B200_FRAME_DELAY=1): get_frame after
 *           decode(N) returns frame N-1, decode(NULL, 0) flushes; one instance then overlaps
 *           its own host parse with its own device reconstruction
 *   --sum : byte-sum only of instance 0's last pass (cheap check that pixels really arrive)
 * prints one JSON line.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "vpx/vpx_decoder.h"
#include "vpx/vp8dx.h"
#include "vp8b200.h"
#include "bench_touch.h"

typedef struct { uint8_t *data; size_t size; int nframes; size_t *off; uint32_t *len; } clip_t;
typedef struct { vpx_codec_ctx_t dec; const clip_t *clip; int ready; } inst_t;

static int g_threads = 1, g_streams = 1, g_repeat = 1, g_sum = 0, g_nclips = 0, g_pipeline = 0, g_touch = 0, g_delay = 0;
static uint64_t g_tsum[1024];                      /* per-worker --touch sums */
static clip_t g_clips[1024];
static inst_t *g_inst;
static pthread_barrier_t g_bar;
static double g_t0, g_t1[1024];
static long g_frames[1024];
static uint64_t g_checksum;
static double g_cpu_dec[1024], g_cpu_get[1024];   /* thread CPU seconds inside the two API calls */
static double g_runq[1024];                        /* seconds a worker sat runnable without a core */

/* /proc/thread-self/schedstat: "<on-cpu ns> <run-queue wait ns> <timeslices>" */
static double runq_s(void)
{
    FILE *f = fopen("/proc/thread-self/schedstat", "r");
    unsigned long long cpu = 0, wait = 0;
    if (!f) return 0;
    if (fscanf(f, "%llu %llu", &cpu, &wait) != 2) wait = 0;
    fclose(f);
    return 1e-9 * (double)wait;
}

static double cpu_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void load_clip(clip_t *c, const char *path)
{
    FILE *f = fopen(path, "rb");
    size_t pos = 32;
    int cap = 1024;
    if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END); c->size = (size_t)ftell(f); fseek(f, 0, SEEK_SET);
    c->data = (uint8_t *)malloc(c->size);
    if (fread(c->data, 1, c->size, f) != c->size) { perror("fread"); exit(2); }
    fclose(f);
    if (c->size < 32 || memcmp(c->data, "DKIF", 4)) { fprintf(stderr, "%s: not IVF\n", path); exit(2); }
    c->off = (size_t *)malloc(sizeof(size_t) * cap);
    c->len = (uint32_t *)malloc(sizeof(uint32_t) * cap);
    c->nframes = 0;
    while (pos + 12 <= c->size) {
        const uint8_t *p = c->data + pos;
        uint32_t n = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
        pos += 12;
        if (pos + n > c->size) break;
        if (c->nframes == cap) {
            cap *= 2;
            c->off = (size_t *)realloc(c->off, sizeof(size_t) * cap);
            c->len = (uint32_t *)realloc(c->len, sizeof(uint32_t) * cap);
        }
        c->off[c->nframes] = pos; c->len[c->nframes] = n; c->nframes++;
        pos += n;
    }
}

/* vpx_codec_decode of frame f: parse on the host, records to the device, kernels queued */
static void submit_one(inst_t *in, int f, double *cpu)
{
    const clip_t *c = in->clip;
    double c0 = cpu ? cpu_s() : 0;
    if (vpx_codec_decode(&in->dec, c->data + c->off[f], c->len[f], NULL, 0)) {
        fprintf(stderr, "decode failed: %s (%s)\n", vpx_codec_error(&in->dec),
                vpx_codec_error_detail(&in->dec) ? vpx_codec_error_detail(&in->dec) : "");
        exit(3);
    }
    if (cpu) cpu[0] += cpu_s() - c0;
}

/* vpx_codec_get_frame until empty: waits for the device, frame arrives in host memory */
static long drain_one(inst_t *in, uint64_t *sum, double *cpu)
{
    vpx_codec_iter_t it = NULL;
    vpx_image_t *img;
    long shown = 0;
    double c0 = cpu ? cpu_s() : 0;
    while ((img = vpx_codec_get_frame(&in->dec, &it))) {
        shown++;
        if (sum) {
            *sum += touch_image(img);
        } else {
            /* touch one byte per plane so the frame is really consumed from host memory */
            volatile uint8_t t = img->planes[0][0] ^ img->planes[1][0] ^ img->planes[2][0];
            (void)t;
        }
    }
    if (cpu) cpu[1] += cpu_s() - c0;
    return shown;
}

static long decode_one(inst_t *in, int f, uint64_t *sum, double *cpu)
{
    submit_one(in, f, cpu);
    return drain_one(in, sum, cpu);
}

static void *worker(void *arg)
{
    int t = (int)(intptr_t)arg, i, f, r, maxf = 0;
    long n = 0;
    double cpu[2] = {0, 0};
    uint64_t tsum = 0;
    for (i = t; i < g_streams; i += g_threads) {
        vpx_codec_dec_cfg_t cfg = {0};
        inst_t *in = &g_inst[i];
        in->clip = &g_clips[i % g_nclips];
        if (vpx_codec_dec_init(&in->dec, vpx_codec_vp8_dx(), &cfg, 0)) { fprintf(stderr, "init failed\n"); exit(3); }
        if (in->clip->nframes > maxf) maxf = in->clip->nframes;
    }
    /* warm-up pass (untimed): creates the device contexts, pins memory */
    for (f = 0; f < maxf; f++)
        for (i = t; i < g_streams; i += g_threads)
            if (f < g_inst[i].clip->nframes) decode_one(&g_inst[i], f, NULL, NULL);
    if (g_delay)
        for (i = t; i < g_streams; i += g_threads) {
            vpx_codec_decode(&g_inst[i].dec, NULL, 0, NULL, 0);
            drain_one(&g_inst[i], NULL, NULL);
        }
    pthread_barrier_wait(&g_bar);
    if (t == 0) g_t0 = now_s();
    pthread_barrier_wait(&g_bar);
    g_runq[t] = -runq_s();
    for (r = 0; r < g_repeat; r++)
        for (f = 0; f < maxf; f++)
            for (i = t; i < g_streams; i += g_threads) {
                inst_t *in = &g_inst[i];
                uint64_t *sum = g_touch ? &tsum : (g_sum && i == 0 && r == g_repeat - 1) ? &g_checksum : NULL;
                if (f >= in->clip->nframes) continue;
                if (!g_pipeline) { n += decode_one(in, f, sum, cpu); continue; }
                /* --pipeline: a worker that owns several instances collects an instance's frame
                 * only when it comes back to that instance, i.e. after it has parsed a frame of
                 * each of its other instances - by then the device has finished and
                 * vpx_codec_get_frame does not block.  Per instance the call order is still
                 * decode, get_frame, decode, ... (the image stays valid until the next decode). */
                if (in->ready) n += drain_one(in, g_touch ? &tsum : (g_sum && i == 0 && in->ready == 2) ? &g_checksum : NULL, cpu);
                submit_one(in, f, cpu);
                in->ready = r == g_repeat - 1 ? 2 : 1;       /* 2: a frame of the last pass is pending */
            }
    for (i = t; i < g_streams; i += g_threads)
        if (g_inst[i].ready) { n += drain_one(&g_inst[i], g_touch ? &tsum : (g_sum && i == 0 && g_inst[i].ready == 2) ? &g_checksum : NULL, cpu); g_inst[i].ready = 0; }
    if (g_delay)                                   /* flush the picture every decoder still holds */
        for (i = t; i < g_streams; i += g_threads) {
            if (vpx_codec_decode(&g_inst[i].dec, NULL, 0, NULL, 0)) { fprintf(stderr, "flush failed\n"); exit(3); }
            n += drain_one(&g_inst[i], g_touch ? &tsum : (g_sum && i == 0) ? &g_checksum : NULL, cpu);
        }
    g_t1[t] = now_s();
    g_tsum[t] = tsum;
    g_cpu_dec[t] = cpu[0]; g_cpu_get[t] = cpu[1];
    g_runq[t] += runq_s();
    g_frames[t] = n;
    pthread_barrier_wait(&g_bar);
    for (i = t; i < g_streams; i += g_threads) vpx_codec_destroy(&g_inst[i].dec);
    return NULL;
}

int main(int argc, char **argv)
{
    pthread_t th[1024];
    int i;
    long total = 0;
    double tend = 0;
    uint64_t st0[4], st1[4], eng[2] = {0, 0};
    double cpu_dec = 0, cpu_get = 0, runq = 0, thread_wall = 0;
    for (i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--threads") && i + 1 < argc) g_threads = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--streams") && i + 1 < argc) g_streams = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--repeat") && i + 1 < argc) g_repeat = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--sum")) g_sum = 1;
        else if (!strcmp(argv[i], "--pipeline")) g_pipeline = 1;
        else if (!strcmp(argv[i], "--touch")) g_touch = 1;
        else if (!strcmp(argv[i], "--delay")) { g_delay = 1; setenv("VP8B200_FRAME_DELAY", "1", 1); }
        else if (g_nclips < 1024) load_clip(&g_clips[g_nclips++], argv[i]);
    }
    if (!g_nclips) { fprintf(stderr, "usage: b200bench [--threads T] [--streams S] [--repeat R] [--sum] a.ivf ...\n"); return 2; }
    if (g_threads > g_streams) g_threads = g_streams;
    if (g_threads > 1024) g_threads = 1024;
    g_inst = (inst_t *)calloc((size_t)g_streams, sizeof(inst_t));
    if (g_threads > 1) {
        /* the reference initialises its static tables on the first decoded frame behind a plain
         * flag (vp8dx_initialize, onyxd_if.c:60-70): do that once before the workers start */
        inst_t prime;
        vpx_codec_dec_cfg_t cfg = {0};
        memset(&prime, 0, sizeof prime);
        prime.clip = &g_clips[0];
        if (vpx_codec_dec_init(&prime.dec, vpx_codec_vp8_dx(), &cfg, 0)) { fprintf(stderr, "init failed\n"); return 3; }
        if (prime.clip->nframes) decode_one(&prime, 0, NULL, NULL);
        vpx_codec_destroy(&prime.dec);
    }
    pthread_barrier_init(&g_bar, NULL, (unsigned)g_threads);
    for (i = 0; i < g_threads; i++) pthread_create(&th[i], NULL, worker, (void *)(intptr_t)i);
    /* stats snapshot is taken by differencing around the whole run minus warm-up: the
     * workers are symmetric, so scale by the timed share */
    vp8b200_global_stats(st0);
    for (i = 0; i < g_threads; i++) pthread_join(th[i], NULL);
    vp8b200_global_stats(st1);
    vp8b200_engine_stats(getenv("VP8B200_DEVICE") ? atoi(getenv("VP8B200_DEVICE")) : 0, eng);
    for (i = 0; i < g_threads; i++) {
        total += g_frames[i]; cpu_dec += g_cpu_dec[i]; cpu_get += g_cpu_get[i];
        if (g_touch) g_checksum += g_tsum[i];
        runq += g_runq[i]; thread_wall += g_t1[i] - g_t0;
        if (g_t1[i] > tend) tend = g_t1[i];
    }
    {
        /* warm-up = 1 pass, timed = g_repeat passes: per-frame byte counts are identical */
        double share = (double)g_repeat / (double)(g_repeat + 1);
        printf("{\"frames\": %ld, \"wall_s\": %.6f, \"fps\": %.3f, \"threads\": %d, \"streams\": %d, "
               "\"repeat\": %d, \"h2d_bytes\": %.0f, \"d2h_bytes\": %.0f, \"kernel_launches\": %.0f, "
               "\"cpu_ms_per_frame_decode\": %.4f, \"cpu_ms_per_frame_get_frame\": %.4f, "
               "\"runq_wait_ms_per_frame\": %.4f, \"blocked_ms_per_frame\": %.4f, \"touch\": %d, \"pipeline\": %d, "
               "\"frame_delay\": %d, \"engine_batches\": %llu, \"engine_frames\": %llu, \"checksum\": %llu}\n",
               total, tend - g_t0, total / (tend - g_t0), g_threads, g_streams, g_repeat,
               (double)(st1[0] - st0[0]) * share, (double)(st1[1] - st0[1]) * share,
               (double)(st1[2] - st0[2]) * share, 1e3 * cpu_dec / (total ? total : 1),
               1e3 * cpu_get / (total ? total : 1), 1e3 * runq / (total ? total : 1),
               1e3 * (thread_wall - cpu_dec - cpu_get - runq) / (total ? total : 1), g_touch, g_pipeline, g_delay,
               (unsigned long long)eng[0], (unsigned long long)eng[1], (unsigned long long)g_checksum);
    }
    return 0;
}
