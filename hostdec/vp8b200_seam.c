/*
 * vp8b200_seam.c - glue compiled INTO the reference's host decoder (a scratch copy of the
 * reference sources patched by hostdec/apply_seams.py).  It is the reference-side binding
 * of include/vp8b200.h: it turns the parser's per-macroblock state (MODE_INFO, qcoeff,
 * eobs) into vp8b200_mb records and drives the C ABI at the frame-level seams.
 *
 * Our own code; it includes reference headers at build time only.  Nothing here
 * reconstructs pixels: with the seams applied, the reference's prediction / IDCT / loop
 * filter / border functions are never called on the decode path.
 *
 * Seams (reference file:line -> function here):
 *   vp8/decoder/decodframe.c:1064  vp8_setup_intra_recon      -> vp8b200_seam_frame_begin
 *   vp8/decoder/decodframe.c:127   vp8_decode_mb_tokens       -> vp8b200_seam_decode_tokens
 *   vp8/decoder/decodframe.c:191   "do prediction" .. :304    -> vp8b200_seam_record_mb
 *   vp8/decoder/decodframe.c:430   vp8_extend_mb_row          -> (dropped; device rule)
 *   vp8/decoder/onyxd_if.c:576-607 loop filter + extend       -> vp8b200_seam_frame_submit
 *   vp8/decoder/onyxd_if.c:729     *sd = *frame_to_show       -> vp8b200_seam_show (queues the D2H, no wait)
 *   vp8/vp8_dx_iface.c:497         img = &ctx->img            -> vp8b200_seam_wait (vpx_codec_get_frame waits)
 *   vp8/decoder/onyxd_if.c:709     vp8dx_get_raw_frame entry  -> vp8b200_seam_get_raw_frame (frame-delay mode)
 *   vp8/decoder/onyxd_if.c:375     missing-frame path         -> vp8b200_seam_flush (decode(NULL,0) = flush)
 *   vp8/decoder/onyxd_if.c:186,226 VP8_COPY/SET_REFERENCE     -> vp8b200_seam_sync_fb / vp8b200_seam_upload_fb
 *   vp8/decoder/onyxd_if.c:390     vp8_yv12_copy_frame_ptr    -> vp8b200_seam_copy_fb
 *   vp8/decoder/onyxd_if.c:155     vp8_remove_common          -> vp8b200_seam_destroy
 *   vpx_scale/generic/yv12config.c:92,27  vpx_memalign/free   -> vp8b200_seam_alloc/free
 *
 * Environment:
 *   VP8B200_DEVICE=<n>     CUDA device ordinal (default 0)
 *   VP8B200_DUMP=<path>    also write every frame's records to a .rec file
 *                          (include/vp8b200_recfile.h); "%p" in the path -> decoder address
 *   VP8B200_TOKENS=ref     keep the reference's vp8_decode_mb_tokens + qcoeff scan instead of
 *                          the fused token reader (vp8b200_tokens.c); for A/B tests
 *   VP8B200_PARSE_THREADS=<n>  threads of the partition-parallel token parser for streams
 *                          with several token partitions (default: one per partition, at
 *                          most 8 and the number of online cores; 1 = serial)
 *   VP8B200_FRAME_DELAY=1  opt-in one-frame-delay mode: vpx_codec_get_frame after
 *                          decode(N) returns frame N-1, so the host parses frame N+1 while the
 *                          device reconstructs frame N; decode(NULL, 0) flushes the last frame
 *   VP8B200_COALESCE=0     submit every frame on the decoder's own CUDA stream instead of handing
 *                          it to the per-device engine that batches the frames of all decoder
 *                          instances into one launch per kernel (default: 1, coalesce)
 *   VP8B200_FETCH=full     copy the whole allocation (borders included) instead of the visible
 *                          samples only
 *   VP8B200_NO_DEVICE=1    record-capture only: no device is touched and NO pixels are
 *                          produced (frames handed back are undefined).  Exists so that
 *                          golden .rec fixtures can be produced on a machine without a GPU;
 *                          it is not a decode path.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vpx_config.h"
#include "vp8/decoder/onyxd_int.h"
#include "vp8/common/onyxc_int.h"
#include "vp8/common/blockd.h"
#include "vpx/internal/vpx_codec_internal.h"

#include "vp8b200.h"
#include "vp8b200_recfile.h"
#include "vp8b200_seam.h"
#include "vp8b200_tokens.h"
#include "vp8/decoder/detokenize.h"

#include <pthread.h>
#include <sched.h>
#include <unistd.h>

/* Where the next aux entry / coefficient block of the macroblocks being parsed goes.  The
 * serial parser has one cursor for the frame; the partition-parallel parser
 * gives every macroblock ROW its own region of the arenas (row r starts at r * mb_cols aux
 * entries and r * mb_cols * 25 blocks - the dense worst case, so regions cannot collide) and
 * packs the rows together before the frame is submitted. */
typedef struct seam_cursor {
    uint32_t n_aux, n_coef;       /* next free aux entry / coefficient block (absolute index) */
    uint32_t aux_end, coef_end;   /* end of the region this cursor may fill */
    int overflow;
    int tok_valid;                /* tok_off/tok_mask describe the current macroblock */
    uint32_t tok_off, tok_mask;
} seam_cursor;

struct seam_mt;
static __thread seam_cursor *tls_cur;          /* NULL: the frame's serial cursor */
static __thread struct seam_mt *tls_mt;        /* set in the workers of a parallel parse */
static __thread int tls_seen;                  /* progress of the row above this thread last saw */

typedef struct seam_state {
    vp8b200_ctx *ctx;
    int width, height;            /* coded size of ctx */
    int no_device;
    FILE *dump;
    int dump_hdr_written;
    /* current frame */
    int open;
    vp8b200_frame_hdr hdr;
    vp8b200_frame_bufs bufs;
    seam_cursor cur;              /* serial cursor; after a parallel parse: the packed totals */
    int ref_tokens;               /* VP8B200_TOKENS=ref */
    int parse_threads;            /* VP8B200_PARSE_THREADS (0 = auto) */
    struct seam_mt *mt;           /* partition-parallel parser, created on first use */
    /* host-memory record buffers for VP8B200_NO_DEVICE */
    vp8b200_mb *h_mb; vp8b200_aux *h_aux; int16_t *h_coef;
    /* lazy fetch: queued in vp8dx_get_raw_frame, waited for in vp8_get_frame */
    int fetch_full;               /* VP8B200_FETCH=full */
    int coalesce;                 /* VP8B200_COALESCE (default 1): submit through the per-device engine */
    int fetch_queued;             /* the engine was asked to copy the shown frame (frame_submit_show) */
    uint8_t *fetch_queued_dst;
    int fetch_pending;            /* a fetch_begin nobody has waited for yet */
    int fetch_fb;                 /* its frame buffer index */
    int device_failed;            /* sticky: a device call failed, every later frame is an error */
    /* frame-delay mode: pictures are copied to two private pinned images in turn */
    int frame_delay;
    uint8_t *out_buf[2];
    size_t out_size;
    int out_cur;
    int delayed_valid;
    YV12_BUFFER_CONFIG delayed_sd;
    int64_t delayed_ts;
} seam_state;

static void seam_fail(VP8D_COMP *pbi, const char *what, int status)
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    vpx_internal_error(&pbi->common.error, VPX_CODEC_ERROR, "vp8b200: %s: %s (%s)", what,
                       vp8b200_strerror(status),
                       s && s->ctx ? vp8b200_last_error(s->ctx) : "no context");
}

static seam_state *seam_get(VP8D_COMP *pbi)
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    if (!s) {
        const char *e;
        s = (seam_state *)calloc(1, sizeof *s);
        if (!s) {
            vpx_internal_error(&pbi->common.error, VPX_CODEC_MEM_ERROR, "vp8b200: seam state");
            return NULL;                         /* only reached when no setjmp is armed */
        }
        e = getenv("VP8B200_NO_DEVICE");
        s->no_device = e && atoi(e);
        e = getenv("VP8B200_TOKENS");
        s->ref_tokens = e && !strcmp(e, "ref");
        e = getenv("VP8B200_PARSE_THREADS");
        s->parse_threads = e ? atoi(e) : 0;
        e = getenv("VP8B200_FRAME_DELAY");
        s->frame_delay = e && atoi(e) && !s->no_device;
        e = getenv("VP8B200_FETCH");
        s->fetch_full = e && !strcmp(e, "full");
        e = getenv("VP8B200_COALESCE");
        s->coalesce = e ? atoi(e) != 0 : 1;
        e = getenv("VP8B200_DUMP");
        if (e && *e) {
            char path[1024];
            const char *pp = strstr(e, "%p");
            if (pp) snprintf(path, sizeof path, "%.*s%p%s", (int)(pp - e), e, (void *)pbi, pp + 2);
            else snprintf(path, sizeof path, "%s", e);
            s->dump = fopen(path, "wb");
            if (!s->dump) fprintf(stderr, "vp8b200: cannot open dump file %s\n", path);
        }
        pbi->b200_seam = s;
    }
    return s;
}


/* ---- partition-parallel token parser ---------------------------------------
 * A frame with 2^k token partitions stores the tokens of macroblock row r in partition
 * r mod 2^k (decodframe.c:1116-1129), each partition with its own bool decoder, so rows of
 * different partitions can be parsed by different threads.  The only coupling is the entropy
 * context of the row above (VP8_COMMON.above_context, one entry per macroblock column): row r
 * may read column c once row r - 1 has finished column c.  The reference's own threaded
 * decoder (vp8/decoder/threading.c) does not build in this snapshot and couples parsing to
 * reconstruction; this one only parses - every thread runs the reference's decode_mb_row on a
 * private MACROBLOCKD and left context, its records go to the row's own arena region, and the
 * rows are packed afterwards so that the frame's records are byte-identical to the serial
 * parser's (tests/test_hostdec_tokens.py).  Mode / motion-vector parsing (first partition)
 * has already happened for the whole frame (vp8_decode_mode_mvs, decodframe.c:1087). */
typedef void (*seam_row_fn)(void *pbi, int mb_row, void *xd);
#define XD_STRIDE ((sizeof(MACROBLOCKD) + 127) & ~(size_t)127)
#define MT_PAD 16                               /* ints per progress counter: one cache line */

typedef struct seam_mt {
    int n_threads;                              /* workers incl. the calling thread */
    int n_use;                                  /* how many of them take rows in this frame */
    pthread_t *th;
    pthread_mutex_t mu;
    pthread_cond_t cv_start, cv_done;
    unsigned generation;
    int pending, quit;
    /* the frame being parsed */
    VP8D_COMP *pbi;
    seam_row_fn row_fn;
    int mb_rows, mb_cols, num_part;
    int rows_cap;
    volatile int *progress;                     /* [mb_rows][MT_PAD]: macroblocks finished in the row */
    seam_cursor *rowcur;                        /* [mb_rows] */
    MACROBLOCKD *xds;                           /* [n_threads] private copies */
    struct seam_thread *thr;                    /* [n_threads] */
} seam_mt;

/* per-thread state that is written for every block or macroblock: one cache line each, or
 * the threads slow each other down through false sharing */
typedef struct seam_thread {
    ENTROPY_CONTEXT_PLANES left;
    int corrupted;
    BOOL_DECODER bc;                            /* private copy of the partition's decoder */
    seam_cursor cur;                            /* cursor of the row being parsed */
} __attribute__((aligned(128))) seam_thread;

typedef struct { seam_mt *mt; int t; } seam_worker_arg;

/* Parse threads at work in this process, over all decoder instances: a frame takes only as
 * many extra threads as there are idle cores, so a server that already runs one decoder per
 * core parses serially (measured: 8 instances on 8 cores, 160 fps serial vs 108 fps when each
 * spawns 8 spinning threads) while a single 4K stream spreads over its partitions. */
static int g_parse_busy;

static void seam_mt_rows(seam_mt *mt, int t)
{
    VP8D_COMP *pbi = mt->pbi;
    VP8_COMMON *pc = &pbi->common;
    MACROBLOCKD *xd = (MACROBLOCKD *)((char *)mt->xds + (size_t)t * XD_STRIDE);
    int r;
    seam_thread *me = &mt->thr[t];
    xd->left_context = &me->left;
    xd->corrupted = 0;
    tls_mt = mt;
    for (r = 0; r < mt->mb_rows; r++) {
        const int part = r % mt->num_part;
        if (part % mt->n_use != t) continue;
        me->bc = pbi->mbc[part];                 /* this thread is the partition's only user */
        xd->current_bc = &me->bc;
        xd->mode_info_context = pc->mi + r * pc->mode_info_stride;
        me->cur = mt->rowcur[r];
        tls_cur = &me->cur;
        tls_seen = 0;
        mt->row_fn(pbi, r, xd);
        pbi->mbc[part] = me->bc;
        mt->rowcur[r] = me->cur;
    }
    me->corrupted = xd->corrupted;
    tls_cur = NULL;
    tls_mt = NULL;
}

static void *seam_mt_worker(void *argp)
{
    seam_worker_arg *arg = (seam_worker_arg *)argp;
    seam_mt *mt = arg->mt;
    const int t = arg->t;
    unsigned seen = 0;
    free(arg);
    for (;;) {
        pthread_mutex_lock(&mt->mu);
        while (mt->generation == seen && !mt->quit) pthread_cond_wait(&mt->cv_start, &mt->mu);
        if (mt->quit) { pthread_mutex_unlock(&mt->mu); return NULL; }
        seen = mt->generation;
        pthread_mutex_unlock(&mt->mu);
        seam_mt_rows(mt, t);
        pthread_mutex_lock(&mt->mu);
        if (--mt->pending == 0) pthread_cond_signal(&mt->cv_done);
        pthread_mutex_unlock(&mt->mu);
    }
}

static void seam_mt_destroy(seam_state *s)
{
    seam_mt *mt = s->mt;
    int i;
    if (!mt) return;
    pthread_mutex_lock(&mt->mu);
    mt->quit = 1;
    pthread_cond_broadcast(&mt->cv_start);
    pthread_mutex_unlock(&mt->mu);
    for (i = 1; i < mt->n_threads; i++) pthread_join(mt->th[i], NULL);
    pthread_mutex_destroy(&mt->mu);
    pthread_cond_destroy(&mt->cv_start);
    pthread_cond_destroy(&mt->cv_done);
    free(mt->th); free((void *)mt->progress); free(mt->rowcur); free(mt->xds); free(mt->thr);
    free(mt);
    s->mt = NULL;
}

static seam_mt *seam_mt_get(seam_state *s, int n_threads, int mb_rows)
{
    seam_mt *mt = s->mt;
    int i;
    if (mt && mt->n_threads != n_threads) { seam_mt_destroy(s); mt = NULL; }
    if (!mt) {
        mt = (seam_mt *)calloc(1, sizeof *mt);
        if (!mt) return NULL;
        mt->n_threads = n_threads;
        mt->th = (pthread_t *)calloc((size_t)n_threads, sizeof *mt->th);
        if (posix_memalign((void **)&mt->xds, 128, XD_STRIDE * (size_t)n_threads)) mt->xds = NULL;
        if (posix_memalign((void **)&mt->thr, 128, sizeof(seam_thread) * (size_t)n_threads)) mt->thr = NULL;
        pthread_mutex_init(&mt->mu, NULL);
        pthread_cond_init(&mt->cv_start, NULL);
        pthread_cond_init(&mt->cv_done, NULL);
        s->mt = mt;
        if (!mt->th || !mt->xds || !mt->thr) { mt->n_threads = 1; seam_mt_destroy(s); return NULL; }
        for (i = 1; i < n_threads; i++) {
            seam_worker_arg *arg = (seam_worker_arg *)malloc(sizeof *arg);
            if (!arg) { mt->n_threads = i; seam_mt_destroy(s); return NULL; }
            arg->mt = mt; arg->t = i;
            if (pthread_create(&mt->th[i], NULL, seam_mt_worker, arg)) {
                free(arg);
                mt->n_threads = i;              /* join only the threads that exist */
                seam_mt_destroy(s);
                return NULL;
            }
        }
    }
    if (mb_rows > mt->rows_cap) {
        free((void *)mt->progress); free(mt->rowcur);
        mt->progress = NULL;
        if (posix_memalign((void **)&mt->progress, 64, sizeof(int) * MT_PAD * (size_t)mb_rows)) mt->progress = NULL;
        mt->rowcur = (seam_cursor *)malloc(sizeof(seam_cursor) * (size_t)mb_rows);
        mt->rows_cap = (mt->progress && mt->rowcur) ? mb_rows : 0;
        if (!mt->rows_cap) { seam_mt_destroy(s); return NULL; }
    }
    return mt;
}

/* Replaces the macroblock-row loop of vp8_decode_frame (decodframe.c:1116-1129) when the
 * frame has several token partitions.  Returns 1 when all rows were parsed here, 0 when the
 * caller should run its own (serial) loop. */
int vp8b200_seam_decode_rows(VP8D_COMP *pbi, MACROBLOCKD *xd, void (*row_fn)(void *, int, void *))
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    VP8_COMMON *pc = &pbi->common;
    const int num_part = 1 << pc->multi_token_partition;
    int n_threads, r, t, i;
    seam_mt *mt;
    uint32_t co = 0, ao = 0;

    if (!s || !s->open || s->ref_tokens || num_part < 2 || s->parse_threads == 1) return 0;
    /* per-row regions need the dense worst case per row */
    if (s->bufs.aux_capacity < (uint32_t)(pc->mb_rows * pc->mb_cols) ||
        s->bufs.coef_capacity < (uint32_t)(pc->mb_rows * pc->mb_cols) * 25u) return 0;
    n_threads = s->parse_threads > 0 ? s->parse_threads : (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (n_threads > num_part) n_threads = num_part;
    if (n_threads > 8) n_threads = 8;
    if (n_threads > pc->mb_rows) n_threads = pc->mb_rows;
    if (n_threads < 2) return 0;
    mt = seam_mt_get(s, n_threads, pc->mb_rows);
    if (!mt) return 0;
    {
        int n_use = n_threads;
        if (s->parse_threads <= 0) {             /* auto: leave the busy cores alone */
            const int idle = (int)sysconf(_SC_NPROCESSORS_ONLN) - __atomic_load_n(&g_parse_busy, __ATOMIC_RELAXED);
            if (n_use > idle) n_use = idle;
        }
        if (n_use < 2) {
            int done;
            __atomic_add_fetch(&g_parse_busy, 1, __ATOMIC_RELAXED);
            for (done = 0; done < pc->mb_rows; done++) {
                xd->current_bc = &pbi->mbc[done % num_part];
                row_fn(pbi, done, xd);
            }
            __atomic_sub_fetch(&g_parse_busy, 1, __ATOMIC_RELAXED);
            return 1;
        }
        mt->n_use = n_use;
        __atomic_add_fetch(&g_parse_busy, n_use, __ATOMIC_RELAXED);
    }

    mt->pbi = pbi; mt->row_fn = row_fn;
    mt->mb_rows = pc->mb_rows; mt->mb_cols = pc->mb_cols; mt->num_part = num_part;
    for (r = 0; r < pc->mb_rows; r++) {
        seam_cursor *cu = &mt->rowcur[r];
        memset(cu, 0, sizeof *cu);
        cu->n_aux = (uint32_t)(r * pc->mb_cols);
        cu->aux_end = cu->n_aux + (uint32_t)pc->mb_cols;
        cu->n_coef = (uint32_t)(r * pc->mb_cols) * 25u;
        cu->coef_end = cu->n_coef + (uint32_t)pc->mb_cols * 25u;
        mt->progress[r * MT_PAD] = 0;
    }
    for (t = 0; t < n_threads; t++) memcpy((char *)mt->xds + (size_t)t * XD_STRIDE, xd, sizeof *xd);

    pthread_mutex_lock(&mt->mu);
    mt->pending = n_threads - 1;
    mt->generation++;
    pthread_cond_broadcast(&mt->cv_start);
    pthread_mutex_unlock(&mt->mu);
    seam_mt_rows(mt, 0);                        /* the calling thread is worker 0 */
    pthread_mutex_lock(&mt->mu);
    while (mt->pending) pthread_cond_wait(&mt->cv_done, &mt->mu);
    pthread_mutex_unlock(&mt->mu);
    __atomic_sub_fetch(&g_parse_busy, mt->n_use, __ATOMIC_RELAXED);

    /* pack the rows: afterwards offsets and arenas are what the serial parser produces */
    for (r = 0; r < pc->mb_rows; r++) {
        seam_cursor *cu = &mt->rowcur[r];
        const uint32_t a0 = (uint32_t)(r * pc->mb_cols), c0 = a0 * 25u;
        const uint32_t na = cu->n_aux - a0, nc = cu->n_coef - c0;
        vp8b200_mb *row = s->bufs.mb + (size_t)r * pc->mb_cols;
        const uint32_t da = a0 - ao, dc = c0 - co;
        if (cu->overflow) s->cur.overflow = 1;
        if (da && na) memmove(s->bufs.aux + ao, s->bufs.aux + a0, (size_t)na * sizeof(vp8b200_aux));
        if (dc && nc) memmove(s->bufs.coef + (size_t)co * 16, s->bufs.coef + (size_t)c0 * 16, (size_t)nc * 32);
        if (da || dc) {
            for (i = 0; i < pc->mb_cols; i++) {
                row[i].coef_off -= dc;
                if (row[i].y_mode == B_PRED || row[i].y_mode == SPLITMV) row[i].u.aux -= da;
            }
        }
        ao += na; co += nc;
    }
    s->cur.n_aux = ao; s->cur.n_coef = co;
    for (t = 0; t < n_threads; t++) xd->corrupted |= mt->thr[t].corrupted;
    /* leave the caller's MACROBLOCKD where the serial loop would: past the last row */
    xd->mode_info_context = pc->mi + pc->mb_rows * pc->mode_info_stride;
    return 1;
}

/* called around decode_macroblock in decode_mb_row (decodframe.c:409): the row above must
 * have finished this column before its entropy context is read, and this column is
 * published once its context is written */
#define MT_LAG 8        /* columns a row stays behind the row above: more than one cache line of
                        * context entries (9 bytes each), so two threads do not share a line */
void vp8b200_seam_mb_wait(int mb_row, int mb_col)
{
    seam_mt *mt = tls_mt;
    if (!mt || mb_row == 0) return;
    {
        volatile int *p = &mt->progress[(mb_row - 1) * MT_PAD];
        const int need = mb_col + MT_LAG < mt->mb_cols ? mb_col + MT_LAG : mt->mb_cols;
        if (tls_seen >= need) return;
        int v, spins = 0;
        while ((v = __atomic_load_n(p, __ATOMIC_ACQUIRE)) < need) {
            if (++spins < 200) __builtin_ia32_pause(); else { sched_yield(); spins = 0; }
        }
        tls_seen = v;
    }
}

void vp8b200_seam_mb_done(int mb_row, int mb_col)
{
    seam_mt *mt = tls_mt;
    if (!mt) return;
    /* published every MT_LAG columns: per-macroblock work can be well under a microsecond, a
     * cache-line hand-off per macroblock would cost as much as the parse itself */
    if (((mb_col + 1) & (MT_LAG - 1)) == 0 || mb_col + 1 == mt->mb_cols)
        __atomic_store_n(&mt->progress[mb_row * MT_PAD], mb_col + 1, __ATOMIC_RELEASE);
}

/* Frame-buffer memory (yv12config.c:92,27).  Page-locked so that the fetch is a DMA; the kind
 * of allocation is remembered in a 64-byte header in front of the block, so freeing does not
 * depend on the environment at that time. */
#define SEAM_HDR 64
#define SEAM_PINNED 0x50494e44u   /* "PIND" */
#define SEAM_MALLOC 0x4d414c43u   /* "MALC" */
static int seam_device(void)
{
    const char *e = getenv("VP8B200_DEVICE");
    return e ? atoi(e) : 0;
}

void *vp8b200_seam_alloc(size_t bytes)
{
    const char *e = getenv("VP8B200_NO_DEVICE");
    uint8_t *p = NULL;
    uint32_t kind;
    if (e && atoi(e)) {
        void *q = NULL;
        if (posix_memalign(&q, 64, bytes + SEAM_HDR)) return NULL;
        p = (uint8_t *)q; kind = SEAM_MALLOC;
    } else {
        p = (uint8_t *)vp8b200_host_alloc_on(seam_device(), bytes + SEAM_HDR);
        if (!p) return NULL;
        kind = SEAM_PINNED;
    }
    memcpy(p, &kind, sizeof kind);
    return p + SEAM_HDR;
}

void vp8b200_seam_free(void *ptr)
{
    uint8_t *p = (uint8_t *)ptr;
    uint32_t kind;
    if (!p) return;
    p -= SEAM_HDR;
    memcpy(&kind, p, sizeof kind);
    if (kind == SEAM_PINNED) vp8b200_host_free(p);
    else if (kind == SEAM_MALLOC) free(p);
    /* anything else was not allocated here: leave it alone */
}

void vp8b200_seam_destroy(VP8D_COMP *pbi)
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    if (!s) return;
    seam_mt_destroy(s);
    if (s->ctx) vp8b200_destroy(s->ctx);
    vp8b200_host_free(s->out_buf[0]); vp8b200_host_free(s->out_buf[1]);
    if (s->dump) fclose(s->dump);
    free(s->h_mb); free(s->h_aux); free(s->h_coef);
    free(s);
    pbi->b200_seam = NULL;
}

/* QIndex of a segment: mb_init_dequantizer, vp8/decoder/decodframe.c:67-87 */
static int segment_qindex(const VP8_COMMON *pc, const MACROBLOCKD *xd, int seg)
{
    int q = pc->base_qindex;
    if (xd->segmentation_enabled) {
        if (xd->mb_segement_abs_delta == SEGMENT_ABSDATA)
            q = xd->segment_feature_data[MB_LVL_ALT_Q][seg];
        else {
            q += xd->segment_feature_data[MB_LVL_ALT_Q][seg];
            q = q >= 0 ? (q <= MAXQ ? q : MAXQ) : 0;
        }
    }
    return q & 127;
}

void vp8b200_seam_frame_begin(VP8D_COMP *pbi)
{
    VP8_COMMON *pc = &pbi->common;
    MACROBLOCKD *xd = &pbi->mb;
    seam_state *s = seam_get(pbi);
    int w = pc->mb_cols * 16, h = pc->mb_rows * 16, seg, i, st;
    uint32_t n_mb = (uint32_t)(pc->mb_rows * pc->mb_cols);
    vp8b200_frame_hdr *hd = &s->hdr;

    if (!s) return;
    if (s->device_failed)
        vpx_internal_error(&pc->error, VPX_CODEC_ERROR, "vp8b200: the device failed on an earlier frame");
    if (s->open && s->ctx) vp8b200_frame_abort(s->ctx);   /* frame abandoned by a longjmp */
    s->open = 0;

    if (s->width != w || s->height != h) {                /* first frame or size change */
        if (s->ctx) { vp8b200_destroy(s->ctx); s->ctx = NULL; }
        s->fetch_pending = 0;                             /* destroy waited for the device */
        vp8b200_host_free(s->out_buf[0]); vp8b200_host_free(s->out_buf[1]);
        s->out_buf[0] = s->out_buf[1] = NULL;
        free(s->h_mb); free(s->h_aux); free(s->h_coef);
        s->h_mb = NULL; s->h_aux = NULL; s->h_coef = NULL;
        if (!s->no_device) {
            st = vp8b200_create(&s->ctx, seam_device(), w, h, NUM_YV12_BUFFERS);
            if (st) seam_fail(pbi, "vp8b200_create", st);
            if ((int)vp8b200_y_stride(s->ctx) != pc->yv12_fb[0].y_stride ||
                vp8b200_frame_size(s->ctx) != (size_t)pc->yv12_fb[0].frame_size)
                vpx_internal_error(&pc->error, VPX_CODEC_ERROR, "vp8b200: frame layout mismatch");
        } else {
            s->h_mb = (vp8b200_mb *)malloc(sizeof(vp8b200_mb) * n_mb);
            s->h_aux = (vp8b200_aux *)malloc(sizeof(vp8b200_aux) * n_mb);
            s->h_coef = (int16_t *)malloc((size_t)32 * 25 * n_mb);
        }
        s->width = w; s->height = h;
        if (s->dump && !s->dump_hdr_written) {
            vp8b200_rec_file_hdr fh;
            memset(&fh, 0, sizeof fh);
            fh.magic = VP8B200_REC_MAGIC; fh.version = VP8B200_ABI_VERSION;
            fh.display_width = (uint32_t)pc->Width; fh.display_height = (uint32_t)pc->Height;
            fh.coded_width = (uint32_t)w; fh.coded_height = (uint32_t)h;
            fh.n_fb = NUM_YV12_BUFFERS;
            fwrite(&fh, sizeof fh, 1, s->dump);
            s->dump_hdr_written = 1;
        }
    }

    memset(hd, 0, sizeof *hd);
    hd->frame_type = (uint8_t)pc->frame_type;
    hd->use_bilinear_mc = (uint8_t)(pc->use_bilinear_mc_filter != 0);
    hd->full_pixel = (uint8_t)(pc->full_pixel != 0);
    hd->filter_type = (uint8_t)pc->filter_type;
    hd->filter_level = (uint8_t)pc->filter_level;
    hd->sharpness_level = (uint8_t)pc->sharpness_level;
    hd->segmentation_enabled = (uint8_t)(xd->segmentation_enabled != 0);
    hd->segment_abs_delta = (uint8_t)(xd->mb_segement_abs_delta == SEGMENT_ABSDATA);
    hd->mode_ref_lf_delta_enabled = (uint8_t)(xd->mode_ref_lf_delta_enabled != 0);
    hd->fb_new = (uint8_t)pc->new_fb_idx;
    hd->fb_last = (uint8_t)pc->lst_fb_idx;
    hd->fb_golden = (uint8_t)pc->gld_fb_idx;
    hd->fb_altref = (uint8_t)pc->alt_fb_idx;
    for (i = 0; i < 4; i++) {
        hd->segment_lf[i] = xd->segment_feature_data[MB_LVL_ALT_LF][i];
        hd->ref_lf_deltas[i] = xd->ref_lf_deltas[i];
        hd->mode_lf_deltas[i] = xd->mode_lf_deltas[i];
    }
    for (seg = 0; seg < 4; seg++) {
        int q = segment_qindex(pc, xd, seg);
        hd->dequant[seg][0][0] = pc->Y1dequant[q][0]; hd->dequant[seg][0][1] = pc->Y1dequant[q][1];
        hd->dequant[seg][1][0] = pc->Y2dequant[q][0]; hd->dequant[seg][1][1] = pc->Y2dequant[q][1];
        hd->dequant[seg][2][0] = pc->UVdequant[q][0]; hd->dequant[seg][2][1] = pc->UVdequant[q][1];
    }

    if (s->ctx) {
        st = vp8b200_frame_begin(s->ctx, hd, &s->bufs);
        if (st) seam_fail(pbi, "vp8b200_frame_begin", st);
    } else {
        s->bufs.mb = s->h_mb; s->bufs.aux = s->h_aux; s->bufs.coef = s->h_coef;
        s->bufs.aux_capacity = n_mb; s->bufs.coef_capacity = 25 * n_mb;
    }
    memset(&s->cur, 0, sizeof s->cur);
    s->cur.aux_end = s->bufs.aux_capacity; s->cur.coef_end = s->bufs.coef_capacity;
    s->open = 1;
}

/* Called from decode_macroblock (vp8/decoder/decodframe.c:127) in place of
 * vp8_decode_mb_tokens: reads the macroblock's tokens straight into the coefficient arena.
 * Returns eobtotal like the function it replaces. */
int vp8b200_seam_decode_tokens(VP8D_COMP *pbi, MACROBLOCKD *xd)
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    BOOL_DECODER *bc = xd->current_bc;
    const int mode = xd->mode_info_context->mbmi.mode;
    int16_t scratch[25 * 16];
    int16_t *dst;
    vp8b200_booldec bd;
    uint32_t mask = 0;
    int eobtotal;

    seam_cursor *cu;
    if (!s || !s->open || s->ref_tokens) return vp8_decode_mb_tokens(pbi, xd);
    cu = tls_cur ? tls_cur : &s->cur;
    if (cu->n_coef + 25 > cu->coef_end) { cu->overflow = 1; dst = scratch; }
    else dst = s->bufs.coef + (size_t)cu->n_coef * 16;
    bd.buf = bc->user_buffer; bd.buf_end = bc->user_buffer_end;
    bd.value = bc->value; bd.count = bc->count; bd.range = bc->range;
    eobtotal = vp8b200_decode_mb_tokens(&bd, &pbi->common.fc.coef_probs[0][0][0][0],
                                        (signed char *)xd->above_context, (signed char *)xd->left_context,
                                        mode != B_PRED && mode != SPLITMV, dst, &mask);
    bc->user_buffer = bd.buf; bc->value = bd.value; bc->count = bd.count; bc->range = bd.range;
    cu->tok_off = cu->n_coef;
    cu->tok_mask = cu->overflow ? 0 : mask;
    if (!cu->overflow) cu->n_coef += (uint32_t)__builtin_popcount(mask);
    cu->tok_valid = 1;
    return eobtotal;
}

/* Called from decode_macroblock (vp8/decoder/decodframe.c) once tokens are decoded and
 * mb_skip_coeff has its final value; replaces everything from "do prediction" to the end
 * of the function.  Also clears the coefficients it consumed, as the reference's
 * dequant/IDCT functions do (dequantize.c:41, idct_blk.c:35), because the token decoder
 * relies on an all-zero qcoeff[] at the start of every macroblock. */
void vp8b200_seam_record_mb(VP8D_COMP *pbi, MACROBLOCKD *xd, unsigned int mb_idx)
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    const MODE_INFO *mi = xd->mode_info_context;
    const MB_MODE_INFO *mbmi = &mi->mbmi;
    seam_cursor *cu = tls_cur ? tls_cur : &s->cur;
    vp8b200_mb *r = &s->bufs.mb[mb_idx];
    int mode = mbmi->mode, i;
    int skip = mbmi->mb_skip_coeff != 0;
    int has_y2 = mode != B_PRED && mode != SPLITMV;

    r->y_mode = (uint8_t)mode;
    r->uv_mode = (uint8_t)mbmi->uv_mode;
    r->ref_frame = (uint8_t)mbmi->ref_frame;
    r->flags = (uint8_t)((mbmi->segment_id & 3) | (skip ? VP8B200_MBF_SKIP : 0) |
                         (mbmi->need_to_clamp_mvs ? VP8B200_MBF_CLAMP_MVS : 0));
    r->coef_mask = 0;
    r->coef_off = cu->n_coef;

    if (mode == B_PRED || mode == SPLITMV) {
        if (cu->n_aux >= cu->aux_end) { cu->overflow = 1; return; }
        r->u.aux = cu->n_aux;
        {
            vp8b200_aux *a = &s->bufs.aux[cu->n_aux++];
            if (mode == B_PRED) {
                memset(a, 0, sizeof *a);
                for (i = 0; i < 16; i++) a->b_mode[i] = (uint8_t)mi->bmi[i].as_mode;
            } else {
                for (i = 0; i < 16; i++) {
                    a->mv[i].row = mi->bmi[i].mv.as_mv.row;
                    a->mv[i].col = mi->bmi[i].mv.as_mv.col;
                }
            }
        }
    } else {
        r->u.mv.row = mbmi->mv.as_mv.row;
        r->u.mv.col = mbmi->mv.as_mv.col;
    }

    if (!s->ref_tokens) {
        /* the fused token reader already stored this macroblock's blocks */
        if (cu->tok_valid) { r->coef_off = cu->tok_off; r->coef_mask = cu->tok_mask; }
        cu->tok_valid = 0;
    } else if (!skip) {
        /* eobs semantics: detokenize.c:183-384.  Y blocks of a Y2 macroblock start at
         * position 1, so they carry coefficients only when eob > 1. */
        uint32_t mask = 0;
        for (i = 0; i < 25; i++) {
            int eob = xd->eobs[i];
            int present;
            if (i == 24 && !has_y2) continue;
            present = (i < 16 && has_y2) ? eob > 1 : eob > 0;
            if (!present) continue;
            if (cu->n_coef >= cu->coef_end) { cu->overflow = 1; break; }
            memcpy(s->bufs.coef + (size_t)cu->n_coef * 16, xd->qcoeff + i * 16, 32);
            cu->n_coef++;
            mask |= 1u << i;
        }
        r->coef_mask = mask;
        memset(xd->qcoeff, 0, sizeof(xd->qcoeff));
    }
}

/* frame-delay mode: the two private pinned images pictures are copied to in turn */
static uint8_t *seam_out_buffer(seam_state *s, size_t frame_size)
{
    uint8_t *out;
    if (!s->out_buf[0] || s->out_size != frame_size) {
        vp8b200_host_free(s->out_buf[0]); vp8b200_host_free(s->out_buf[1]);
        s->out_buf[0] = (uint8_t *)vp8b200_host_alloc_on(seam_device(), frame_size);
        s->out_buf[1] = (uint8_t *)vp8b200_host_alloc_on(seam_device(), frame_size);
        s->out_size = frame_size;
        if (!s->out_buf[0] || !s->out_buf[1]) return NULL;
    }
    out = s->out_buf[s->out_cur];
    s->out_cur ^= 1;
    return out;
}

static void seam_dump_frame(seam_state *s, VP8D_COMP *pbi, uint32_t n_mb)
{
    VP8_COMMON *cm = &pbi->common;
    vp8b200_rec_frame_hdr fh;
    memset(&fh, 0, sizeof fh);
    fh.magic = VP8B200_REC_FRAME_MAGIC;
    fh.n_mb = n_mb; fh.n_aux = s->cur.n_aux; fh.n_coef = s->cur.n_coef;
    fh.show_frame = (uint8_t)cm->show_frame;
    fh.fb_show = (uint8_t)(cm->frame_to_show - cm->yv12_fb);
    fh.hdr = s->hdr;
    fwrite(&fh, sizeof fh, 1, s->dump);
    fwrite(s->bufs.mb, sizeof(vp8b200_mb), n_mb, s->dump);
    fwrite(s->bufs.aux, sizeof(vp8b200_aux), s->cur.n_aux, s->dump);
    fwrite(s->bufs.coef, 32, s->cur.n_coef, s->dump);
    fflush(s->dump);
}

/* after swap_frame_buffers (onyxd_if.c:560): takes the place of vp8_loop_filter_frame
 * (onyxd_if.c:576-586) and vp8_yv12_extend_frame_borders_ptr (:607) */
void vp8b200_seam_frame_submit(VP8D_COMP *pbi)
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    VP8_COMMON *cm = &pbi->common;
    int st;
    if (!s || !s->open) return;
    if (s->cur.overflow)
        vpx_internal_error(&cm->error, VPX_CODEC_ERROR, "vp8b200: record arena overflow");
    if (s->dump) seam_dump_frame(s, pbi, (uint32_t)(cm->mb_rows * cm->mb_cols));
    s->open = 0;
    s->fetch_queued = 0;
    if (s->ctx && s->coalesce) {
        /* hand the frame AND the request for its picture to the device's engine: this thread
         * makes no CUDA launch (swap_frame_buffers has run: frame_to_show is this frame) */
        int show_fb = -1;
        uint8_t *dst = NULL;
        if (cm->show_frame && cm->frame_to_show) {
            show_fb = (int)(cm->frame_to_show - cm->yv12_fb);
            dst = cm->frame_to_show->buffer_alloc;
            if (s->frame_delay) {
                dst = seam_out_buffer(s, (size_t)cm->frame_to_show->frame_size);
                if (!dst) vpx_internal_error(&cm->error, VPX_CODEC_MEM_ERROR, "vp8b200: pinned output image");
            }
        }
        st = vp8b200_frame_submit_show(s->ctx, s->cur.n_aux, s->cur.n_coef, show_fb, dst,
                                       s->fetch_full ? 0 : cm->Width, s->fetch_full ? 0 : cm->Height);
        if (st) seam_fail(pbi, "vp8b200_frame_submit_show", st);
        if (show_fb >= 0) { s->fetch_queued = 1; s->fetch_queued_dst = dst; s->fetch_fb = show_fb; }
    } else if (s->ctx) {
        st = vp8b200_frame_submit(s->ctx, s->cur.n_aux, s->cur.n_coef);
        if (st) seam_fail(pbi, "vp8b200_frame_submit", st);
    }
}

/* A device call failed outside the decoder's setjmp scope (vp8dx_get_raw_frame / vp8_get_frame
 * run after vp8dx_receive_compressed_data has returned): record the error where
 * vp8_decode / the next frame will find it, mark the picture corrupt - never hand out a stale
 * image as if it were this frame. */
static int seam_device_error(VP8D_COMP *pbi, seam_state *s, const char *what, int status)
{
    VP8_COMMON *cm = &pbi->common;
    s->device_failed = 1;
    s->fetch_pending = 0;
    s->delayed_valid = 0;
    if (cm->frame_to_show) cm->frame_to_show->corrupted = 1;
    cm->error.setjmp = 0;
    seam_fail(pbi, what, status);              /* no longjmp: only records code + detail */
    return -1;
}

static int seam_queue_fetch(VP8D_COMP *pbi, seam_state *s, uint8_t *dst)
{
    VP8_COMMON *cm = &pbi->common;
    const int fb = (int)(cm->frame_to_show - cm->yv12_fb);
    int st = vp8b200_frame_fetch_begin(s->ctx, fb, dst, s->fetch_full ? 0 : cm->Width, s->fetch_full ? 0 : cm->Height);
    if (st) return seam_device_error(pbi, s, "vp8b200_frame_fetch_begin", st);
    s->fetch_pending = 1;
    s->fetch_fb = fb;
    return 0;
}

/* vp8dx_get_raw_frame (onyxd_if.c:729), default mode: hand out the shown buffer's host mirror
 * and only QUEUE its device->host copy; vpx_codec_decode returns while the device is still
 * reconstructing, vpx_codec_get_frame (vp8b200_seam_wait) is where the caller waits. */
int vp8b200_seam_show(VP8D_COMP *pbi, YV12_BUFFER_CONFIG *sd)
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    VP8_COMMON *cm = &pbi->common;
    *sd = *cm->frame_to_show;
    if (!s || !s->ctx) return 0;               /* record-capture mode: no pixels */
    if (s->device_failed) return -1;
    if (s->fetch_queued) {                       /* the engine copies it behind the frame's batch */
        s->fetch_queued = 0;
        s->fetch_pending = 1;
        return 0;
    }
    return seam_queue_fetch(pbi, s, cm->frame_to_show->buffer_alloc);
}

/* vp8_get_frame (vp8_dx_iface.c:485-503): the image is about to be handed to the caller */
int vp8b200_seam_wait(VP8D_COMP *pbi)
{
    seam_state *s = pbi ? (seam_state *)pbi->b200_seam : NULL;
    int st;
    if (!s || !s->ctx) return 0;
    if (s->device_failed) return -1;
    if (!s->fetch_pending || s->frame_delay) return 0;   /* delay mode waits one call later */
    s->fetch_pending = 0;
    st = vp8b200_frame_fetch_wait(s->ctx);
    if (st) return seam_device_error(pbi, s, "vp8b200_frame_fetch_wait", st);
    return 0;
}

/* Entry of vp8dx_get_raw_frame (onyxd_if.c:707).  Returns 1 when the reference's own body
 * should run (default mode).  In frame-delay mode it does the whole job: the picture decoded
 * by THIS call is queued for copy into a private pinned image and the picture of the PREVIOUS
 * call - whose copy has had a whole host parse to finish - is handed out (0), or -1 when there
 * is none yet.  decode(NULL, 0) (vp8b200_seam_flush) drains the last one. */
int vp8b200_seam_get_raw_frame(VP8D_COMP *pbi, YV12_BUFFER_CONFIG *sd, int64_t *time_stamp, int64_t *time_end_stamp)
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    VP8_COMMON *cm = &pbi->common;
    int have_prev, ret = -1;
    if (!s || !s->frame_delay || !s->ctx) return 1;
    if (s->device_failed) return -1;
    have_prev = s->delayed_valid;
    if (have_prev) {
        int st = 0;
        if (s->fetch_pending) { s->fetch_pending = 0; st = vp8b200_frame_fetch_wait(s->ctx); }
        if (st) return seam_device_error(pbi, s, "vp8b200_frame_fetch_wait", st);
        *sd = s->delayed_sd;
        *time_stamp = s->delayed_ts;
        *time_end_stamp = 0;
        s->delayed_valid = 0;
        ret = 0;
    }
    if (pbi->ready_for_new_data == 0) {            /* a frame was decoded by this call */
        pbi->ready_for_new_data = 1;
        if (cm->show_frame && cm->frame_to_show) {
            const YV12_BUFFER_CONFIG *f = cm->frame_to_show;
            uint8_t *out;
            if (s->fetch_queued) {                /* the engine copies it behind the frame's batch */
                out = s->fetch_queued_dst;
                s->fetch_queued = 0;
                s->fetch_pending = 1;
            } else {
                out = seam_out_buffer(s, (size_t)f->frame_size);
                if (!out) return seam_device_error(pbi, s, "pinned output image", VP8B200_ERR_NOMEM);
                if (seam_queue_fetch(pbi, s, out)) return -1;
            }
            s->delayed_sd = *f;
            s->delayed_sd.buffer_alloc = out;
            s->delayed_sd.y_buffer = out + (f->y_buffer - f->buffer_alloc);
            s->delayed_sd.u_buffer = out + (f->u_buffer - f->buffer_alloc);
            s->delayed_sd.v_buffer = out + (f->v_buffer - f->buffer_alloc);
            s->delayed_sd.y_width = cm->Width;
            s->delayed_sd.y_height = cm->Height;
            s->delayed_sd.uv_height = cm->Height / 2;
            s->delayed_sd.clrtype = cm->clr_type;
            s->delayed_ts = pbi->last_time_stamp;
            s->delayed_valid = 1;
        }
    }
    return ret;
}

/* vp8dx_receive_compressed_data with no data (onyxd_if.c:336-373): in frame-delay mode that
 * is the flush request, not a lost frame */
int vp8b200_seam_flush(VP8D_COMP *pbi)
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    return s && s->frame_delay && s->delayed_valid;
}

/* VP8_COPY_REFERENCE (vp8dx_get_reference, onyxd_if.c:161-189): the host mirror of a reference
 * buffer is not kept current (only visible samples of shown frames are fetched), so bring the
 * whole buffer over before the reference's host-side copy reads it. */
int vp8b200_seam_sync_fb(VP8D_COMP *pbi, int idx)
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    VP8_COMMON *cm = &pbi->common;
    int st;
    if (!s || s->no_device) return 0;
    if (!s->ctx || s->device_failed) {
        vpx_internal_error(&cm->error, VPX_CODEC_ERROR, "vp8b200: no device frame to copy the reference from");
        return -1;
    }
    if (s->fetch_pending) { s->fetch_pending = 0; vp8b200_frame_fetch_wait(s->ctx); }
    st = vp8b200_frame_fetch(s->ctx, idx, cm->yv12_fb[idx].buffer_alloc, (size_t)cm->yv12_fb[idx].frame_size);
    if (st) return seam_device_error(pbi, s, "vp8b200_frame_fetch", st);
    return 0;
}

/* VP8_SET_REFERENCE (vp8dx_set_reference, onyxd_if.c:192-230): the reference's host-side copy
 * (which also extends the borders) has filled the mirror of buffer idx; the device copy is
 * what later frames predict from. */
int vp8b200_seam_upload_fb(VP8D_COMP *pbi, int idx)
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    VP8_COMMON *cm = &pbi->common;
    int st;
    if (!s || s->no_device) return 0;
    if (!s->ctx || s->device_failed) {
        vpx_internal_error(&cm->error, VPX_CODEC_ERROR, "vp8b200: no device context to set the reference in");
        return -1;
    }
    st = vp8b200_frame_upload(s->ctx, idx, cm->yv12_fb[idx].buffer_alloc, (size_t)cm->yv12_fb[idx].frame_size);
    if (st) return seam_device_error(pbi, s, "vp8b200_frame_upload", st);
    return 0;
}

/* onyxd_if.c:390: the missing-frame path moves `last` to its own buffer */
void vp8b200_seam_copy_fb(VP8D_COMP *pbi, int dst_idx, int src_idx)
{
    seam_state *s = (seam_state *)pbi->b200_seam;
    int st;
    if (!s || !s->ctx) return;
    st = vp8b200_frame_copy(s->ctx, dst_idx, src_idx);
    if (st) seam_fail(pbi, "vp8b200_frame_copy", st);
}
