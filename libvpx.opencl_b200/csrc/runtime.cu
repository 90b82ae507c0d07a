/* runtime.cu - contexts, device frame buffers, pinned staging ring, launch sequence and the
 * C ABI of include/vp8b200.h.  No CPU reconstruction exists in this library: when CUDA is
 * unavailable every entry point that would need the device fails.
 *
 * Per frame, on the context's stream:
 *   H2D(records) -> k_inter -> k_intra -> k_loopfilter -> k_border     (-> D2H on fetch)
 * which replaces, in the reference, decode_macroblock's tail (decodframe.c:190-304),
 * vp8_loop_filter_frame (onyxd_if.c:576-586) and vp8_yv12_extend_frame_borders_ptr
 * (onyxd_if.c:607).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <linux/futex.h>
#include <sys/resource.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <new>
#include <thread>
#include <vector>
#include "vp8b200_internal.h"

/* process-wide statistics (monotonic counters only; no behaviour depends on them) */
static std::atomic<uint64_t> g_h2d_bytes{0}, g_d2h_bytes{0}, g_launches{0}, g_frames{0};

struct ProfSpan { int kind; cudaEvent_t a, b; };

#define NSLOT 3          /* pinned/device record slots: parse N+1 while N uploads / runs */
#define MAX_SPANS 4096   /* profiling spans kept between two profile_read calls */
#define NBJOB 4          /* job-array ring for batched launches */

struct Slot {
    vp8b200_mb *h_mb, *d_mb;
    vp8b200_aux *h_aux, *d_aux;
    int16_t *h_coef, *d_coef;
    FrameJob *h_job, *d_job;
    uint32_t *h_ilist, *d_ilist;   /* intra MB indices in wavefront order (P frames) */
    cudaEvent_t h2d_done;
    bool pending;
};

struct vp8b200_staged {
    vp8b200_frame_hdr hdr;
    uint8_t *d_blob;
    vp8b200_mb *d_mb;
    vp8b200_aux *d_aux;
    int16_t *d_coef;
    uint32_t *d_ilist;             /* NULL on key frames: the context's static order is used */
    unsigned n_intra, n_split;
};

struct vp8b200_ctx {
    int device;
    Geo geo;
    size_t frame_size;
    int n_fb;
    uint32_t n_mb;
    uint8_t *fb[VP8B200_MAX_FB];
    cudaStream_t stream;
    Slot slot[NSLOT];
    int cur;
    bool open;
    vp8b200_frame_hdr cur_hdr;
    unsigned long long *d_imsg;    /* per-MB exported intra borders, 16 tagged words each */
    uint32_t *d_diag;              /* all MB indices sorted by wavefront index c + 2r */
    int *diag_tmp;                 /* host scratch for the counting sort */
    uint8_t *d_lfmsg;              /* loop-filter hand-off messages between CTAs (vp8b200_lf_msg_bytes) */
    unsigned *d_tickets;           /* [0] intra, [1] loop filter */
    unsigned ticket_base[2];
    unsigned epoch_intra, epoch_lf;
    FrameJob *h_bjobs[NBJOB], *d_bjobs[NBJOB];
    cudaEvent_t bjobs_done[NBJOB];
    bool bjobs_pending[NBJOB];
    int bjobs_cap, bjobs_cur;
    uint64_t launches;
    bool blocking_sync;            /* VP8B200_SYNC=block: sleep instead of spinning in fetch */
    /* lazy device->host fetch: the copy is queued on its own stream behind
     * recon_done (recorded on the launch stream); its event fires when the pixels are in host
     * memory.  Up to two copies are in flight per context, oldest first (frame-delay mode
     * collects picture N-1 after picture N has been queued): fetch[i].fb >= 0 while that copy
     * may still be reading the device buffer, .seq != 0 when the ENGINE thread records the
     * event while issuing the submit with that sequence number. */
    cudaStream_t copy_stream;
    cudaEvent_t recon_done;
    struct { cudaEvent_t ev; int fb; uint32_t seq; } fetch[2];
    int fetch_head, fetch_n;
    bool profiling, prof_skip;
    std::vector<ProfSpan> *spans;
    /* cross-stream ordering between a context's own stream and a batch leader's stream */
    cudaEvent_t own_ev;            /* recorded on this context's stream when a batch must wait for it */
    bool own_dirty;                /* work queued on the own stream since the last batch / sync */
    cudaEvent_t batch_ev;          /* event of the batch (on the leader's stream) that last touched this context */
    bool batch_pending;            /* own stream has not yet waited for batch_ev */
    vp8b200_ctx *batch_leader;     /* whose stream that batch ran on */
    cudaEvent_t lead_ev[NBJOB];    /* as a leader: one event per in-flight batch */
    /* per-device submit coalescer (see Engine below): frames handed to it / issued by it */
    struct Engine *eng;
    uint64_t eng_submitted;        /* frames handed to the engine (decoder thread) */
    uint32_t eng_issued;           /* ... issued by it (engine thread; a futex word: the decoder thread sleeps on it) */
    int eng_status;                /* first error of an issue, reported by the next call on this context */
    char err[256];
};

static const char *k_noerr = "";

#define CK(ctx, call)                                                                      \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s: %s", #call, cudaGetErrorString(e_)); \
            return VP8B200_ERR_CUDA;                                                       \
        }                                                                                  \
    } while (0)

/* ---- per-device submit coalescer ("engine"), SURVEY 8b "shared batch scheduler across ctxs" ----
 * vp8b200_frame_submit_show hands a parsed frame to the engine of the context's device and
 * returns; ONE engine thread per device gathers what the decoder threads of all contexts have
 * queued and issues it as one batch: the records' H2D copies, ONE launch of each kernel over
 * all gathered frames (the same batched launch bench.py's resident replay uses), and the D2H
 * copies of the frames that are shown.  Decoder threads make no CUDA launch at all; the only
 * place they wait is vp8b200_frame_fetch_wait (vpx_codec_get_frame). */
struct EngineSubmit {
    vp8b200_ctx *c;
    int slot;
    uint32_t n_aux, n_coef;
    unsigned n_intra, n_split;
    vp8b200_frame_hdr hdr;
    uint64_t seq;
    int show_fb;                   /* < 0: not shown */
    int show_rec;                  /* which of the context's two fetch records tracks the copy */
    uint8_t *show_dst;
    int show_w, show_h;
};
#define ENG_RING 8
struct Engine {
    int device;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<EngineSubmit> q;
    int n_ctx;                     /* contexts that have used this engine and are alive */
    bool started, failed;
    cudaStream_t stream[2], copy_stream;
    FrameJob *h_jobs[ENG_RING], *d_jobs[ENG_RING];
    cudaEvent_t jobs_ev[ENG_RING], batch_ev[ENG_RING];
    bool ring_pending[ENG_RING];
    int cap, cur;
    unsigned *d_tickets[2];
    unsigned ticket_base[2][2];
    int window_us, max_batch;
    std::atomic<uint64_t> batches{0}, frames{0};
};
static std::mutex g_eng_mu;
static Engine *g_engines[64];
static void engine_thread(Engine *e);
static void engine_settle(vp8b200_ctx *c);

extern "C" int vp8b200_abi_version(void) { return VP8B200_ABI_VERSION; }

extern "C" const char *vp8b200_strerror(int st)
{
    switch (st) {
    case VP8B200_OK: return "ok";
    case VP8B200_ERR_INVALID: return "invalid argument or call order";
    case VP8B200_ERR_NO_DEVICE: return "no usable CUDA device (this library has no CPU path)";
    case VP8B200_ERR_NOMEM: return "out of memory";
    case VP8B200_ERR_CUDA: return "CUDA error";
    case VP8B200_ERR_OVERFLOW: return "record arena overflow";
    default: return "unknown status";
    }
}

extern "C" const char *vp8b200_last_error(const vp8b200_ctx *ctx) { return ctx ? ctx->err : k_noerr; }

extern "C" int vp8b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" void *vp8b200_host_alloc(size_t bytes)
{
    void *p = NULL;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return NULL; }
    return p;
}
extern "C" void *vp8b200_host_alloc_on(int device, size_t bytes)
{
    /* the first CUDA call of a thread creates a primary context on the current device: select
     * the decoder's device first, or every rank's frame buffers pin through GPU 0 */
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return NULL; }
    return vp8b200_host_alloc(bytes);
}
extern "C" void vp8b200_host_free(void *p) { if (p) cudaFreeHost(p); }

/* Live contexts.  A batch member borrows an event owned by its leader (batch_ev = the
 * leader's lead_ev[r]), so a leader that goes away first has to retire those references. */
static std::mutex g_live_mu;
static std::vector<vp8b200_ctx *> g_live;

static void free_ctx(vp8b200_ctx *c)
{
    if (!c) return;
    engine_settle(c);
    if (c->eng) {
        std::lock_guard<std::mutex> lk(c->eng->mu);
        c->eng->n_ctx--;
        c->eng = NULL;
    }
    cudaSetDevice(c->device);
    if (c->batch_pending) cudaEventSynchronize(c->batch_ev);   /* a batch on another leader's stream may still use us */
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    for (int i = 0; i < 2; i++)                                 /* copies queued by the engine on its own copy stream */
        if (c->fetch[i].ev && c->fetch[i].fb >= 0) cudaEventSynchronize(c->fetch[i].ev);
    {
        /* every batch this context led has finished now: members need not (and, once the
         * events below are destroyed, must not) wait for them any more */
        std::lock_guard<std::mutex> lk(g_live_mu);
        g_live.erase(std::remove(g_live.begin(), g_live.end(), c), g_live.end());
        for (vp8b200_ctx *m : g_live)
            if (m->batch_leader == c) { m->batch_pending = false; m->batch_leader = NULL; m->batch_ev = NULL; }
    }
    for (int i = 0; i < c->n_fb; i++) cudaFree(c->fb[i]);
    for (int i = 0; i < NSLOT; i++) {
        Slot &s = c->slot[i];
        cudaFreeHost(s.h_mb); cudaFreeHost(s.h_aux); cudaFreeHost(s.h_coef); cudaFreeHost(s.h_job);
        cudaFreeHost(s.h_ilist); cudaFree(s.d_ilist);
        cudaFree(s.d_mb); cudaFree(s.d_aux); cudaFree(s.d_coef); cudaFree(s.d_job);
        if (s.h2d_done) cudaEventDestroy(s.h2d_done);
    }
    for (int i = 0; i < NBJOB; i++) {
        cudaFreeHost(c->h_bjobs[i]); cudaFree(c->d_bjobs[i]);
        if (c->bjobs_done[i]) cudaEventDestroy(c->bjobs_done[i]);
        if (c->lead_ev[i]) cudaEventDestroy(c->lead_ev[i]);
    }
    if (c->own_ev) cudaEventDestroy(c->own_ev);
    cudaFree(c->d_imsg); cudaFree(c->d_diag); cudaFree(c->d_tickets); cudaFree(c->d_lfmsg);
    free(c->diag_tmp);
    for (int i = 0; i < 2; i++) if (c->fetch[i].ev) cudaEventDestroy(c->fetch[i].ev);
    if (c->recon_done) cudaEventDestroy(c->recon_done);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->spans) {
        for (auto &sp : *c->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
        delete c->spans;
    }
    if (c->stream) cudaStreamDestroy(c->stream);
    cudaGetLastError();
    delete c;
}

/* Indices of the intra macroblocks (all macroblocks when mb == NULL) sorted by wavefront
 * index d = col + 2*row, raster order inside one d (counting sort).  Every neighbour an intra
 * MB depends on (left d-1, above-left d-3, above d-2, above-right d-1) sorts before it. */
static unsigned wavefront_order(const Geo &g, const vp8b200_mb *mb, uint32_t n_mb, uint32_t *out, int *cnt)
{
    const int nd = g.mb_cols + 2 * g.mb_rows;
    for (int d = 0; d <= nd; d++) cnt[d] = 0;
    for (uint32_t i = 0; i < n_mb; i++)
        if (!mb || mb[i].ref_frame == VP8B200_INTRA_FRAME) {
            const int r = (int)(i / (uint32_t)g.mb_cols), c = (int)(i % (uint32_t)g.mb_cols);
            cnt[c + 2 * r + 1]++;
        }
    for (int d = 0; d < nd; d++) cnt[d + 1] += cnt[d];
    const unsigned total = (unsigned)cnt[nd];
    for (uint32_t i = 0; i < n_mb; i++)
        if (!mb || mb[i].ref_frame == VP8B200_INTRA_FRAME) {
            const int r = (int)(i / (uint32_t)g.mb_cols), c = (int)(i % (uint32_t)g.mb_cols);
            out[cnt[c + 2 * r]++] = i;
        }
    return total;
}

static int create_impl(vp8b200_ctx *c)
{
    const Geo &g = c->geo;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CK(c, cudaEventCreateWithFlags(&c->recon_done, cudaEventDisableTiming));
    c->fetch[0].fb = c->fetch[1].fb = -1;
    c->fetch_head = c->fetch_n = 0;
    vp8b200_upload_constants();
    vp8b200_upload_intra_constants();
    CK(c, cudaGetLastError());
    for (int i = 0; i < c->n_fb; i++) {
        CK(c, cudaMalloc((void **)&c->fb[i], c->frame_size));
        CK(c, cudaMemsetAsync(c->fb[i], 0, c->frame_size, c->stream));
    }
    const size_t n_mb = c->n_mb;
    for (int i = 0; i < NSLOT; i++) {
        Slot &s = c->slot[i];
        CK(c, cudaHostAlloc((void **)&s.h_mb, n_mb * sizeof(vp8b200_mb), cudaHostAllocPortable));
        CK(c, cudaHostAlloc((void **)&s.h_aux, n_mb * sizeof(vp8b200_aux), cudaHostAllocPortable));
        CK(c, cudaHostAlloc((void **)&s.h_coef, n_mb * 25 * 32, cudaHostAllocPortable));
        CK(c, cudaHostAlloc((void **)&s.h_job, sizeof(FrameJob), cudaHostAllocPortable));
        CK(c, cudaMalloc((void **)&s.d_mb, n_mb * sizeof(vp8b200_mb)));
        CK(c, cudaMalloc((void **)&s.d_aux, n_mb * sizeof(vp8b200_aux)));
        CK(c, cudaMalloc((void **)&s.d_coef, n_mb * 25 * 32));
        CK(c, cudaMalloc((void **)&s.d_job, sizeof(FrameJob)));
        CK(c, cudaHostAlloc((void **)&s.h_ilist, n_mb * sizeof(uint32_t), cudaHostAllocPortable));
        CK(c, cudaMalloc((void **)&s.d_ilist, n_mb * sizeof(uint32_t)));
        CK(c, cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
    }
    for (int i = 0; i < NBJOB; i++) {
        CK(c, cudaEventCreateWithFlags(&c->bjobs_done[i], cudaEventDisableTiming));
        CK(c, cudaEventCreateWithFlags(&c->lead_ev[i], cudaEventDisableTiming));
    }
    CK(c, cudaEventCreateWithFlags(&c->own_ev, cudaEventDisableTiming));
    CK(c, cudaMalloc((void **)&c->d_imsg, n_mb * 128));
    CK(c, cudaMemsetAsync(c->d_imsg, 0, n_mb * 128, c->stream));
    CK(c, cudaMalloc((void **)&c->d_diag, n_mb * sizeof(uint32_t)));
    c->diag_tmp = (int *)malloc(sizeof(int) * (size_t)(g.mb_cols + 2 * g.mb_rows + 2));
    if (!c->diag_tmp) return VP8B200_ERR_NOMEM;
    {
        /* static wavefront order of ALL macroblocks (key frames) */
        uint32_t *tmp = (uint32_t *)malloc(n_mb * sizeof(uint32_t));
        if (!tmp) return VP8B200_ERR_NOMEM;
        wavefront_order(g, NULL, (uint32_t)n_mb, tmp, c->diag_tmp);
        cudaError_t e = cudaMemcpy(c->d_diag, tmp, n_mb * sizeof(uint32_t), cudaMemcpyHostToDevice);
        free(tmp);
        CK(c, e);
    }
    CK(c, cudaMalloc((void **)&c->d_lfmsg, vp8b200_lf_msg_bytes(g)));
    CK(c, cudaMemsetAsync(c->d_lfmsg, 0, vp8b200_lf_msg_bytes(g), c->stream));
    CK(c, cudaMalloc((void **)&c->d_tickets, 2 * sizeof(unsigned)));
    CK(c, cudaMemsetAsync(c->d_tickets, 0, 2 * sizeof(unsigned), c->stream));
    {
        const char *e = getenv("VP8B200_SYNC");
        c->blocking_sync = e && !strcmp(e, "block");
        for (int i = 0; i < 2; i++)
            CK(c, cudaEventCreateWithFlags(&c->fetch[i].ev, cudaEventDisableTiming |
                                           (c->blocking_sync ? cudaEventBlockingSync : 0)));
    }
    CK(c, cudaStreamSynchronize(c->stream));
    return VP8B200_OK;
}

extern "C" int vp8b200_create(vp8b200_ctx **out, int device, int width, int height, int n_fb)
{
    if (!out || width <= 0 || height <= 0 || (width & 15) || (height & 15) || width > 65536 ||
        height > 65536 || n_fb < 1 || n_fb > VP8B200_MAX_FB)
        return VP8B200_ERR_INVALID;
    *out = NULL;
    int ndev = vp8b200_device_count();
    if (ndev <= 0 || device < 0 || device >= ndev) return VP8B200_ERR_NO_DEVICE;
    vp8b200_ctx *c = new (std::nothrow) vp8b200_ctx();
    if (!c) return VP8B200_ERR_NOMEM;
    memset(c, 0, sizeof *c);
    c->device = device;
    c->n_fb = n_fb;
    Geo &g = c->geo;
    g.width = width; g.height = height;
    g.mb_cols = width >> 4; g.mb_rows = height >> 4;
    g.y_stride = ((width + 2 * VP8B200_BORDER) + 31) & ~31;          /* yv12config.c:61 */
    g.uv_stride = g.y_stride >> 1;
    const size_t yplane = (size_t)(height + 2 * VP8B200_BORDER) * g.y_stride;
    g.uv_rows_alloc = (height >> 1) + VP8B200_BORDER;
    const size_t uvplane = (size_t)g.uv_rows_alloc * g.uv_stride;
    c->frame_size = yplane + 2 * uvplane;
    g.y_off = VP8B200_BORDER * g.y_stride + VP8B200_BORDER;          /* yv12config.c:107-109 */
    g.u_off = (int)yplane + (VP8B200_BORDER / 2) * g.uv_stride + VP8B200_BORDER / 2;
    g.v_off = (int)(yplane + uvplane) + (VP8B200_BORDER / 2) * g.uv_stride + VP8B200_BORDER / 2;
    c->n_mb = (uint32_t)(g.mb_cols * g.mb_rows);
    c->epoch_intra = c->epoch_lf = 0;
    int st = create_impl(c);
    if (st != VP8B200_OK) {
        fprintf(stderr, "vp8b200_create: %s\n", c->err);
        free_ctx(c);
        return st;
    }
    {
        std::lock_guard<std::mutex> lk(g_live_mu);
        g_live.push_back(c);
    }
    *out = c;
    return VP8B200_OK;
}

extern "C" void vp8b200_destroy(vp8b200_ctx *ctx) { free_ctx(ctx); }
extern "C" size_t vp8b200_frame_size(const vp8b200_ctx *ctx) { return ctx ? ctx->frame_size : 0; }
extern "C" int vp8b200_y_stride(const vp8b200_ctx *ctx) { return ctx ? ctx->geo.y_stride : 0; }
extern "C" uint64_t vp8b200_launch_count(const vp8b200_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" void *vp8b200_stream(const vp8b200_ctx *ctx) { return ctx ? (void *)ctx->stream : NULL; }

/* A context that took part in vp8b200_batch_run was written by kernels on the LEADER's stream;
 * make its own stream wait for that batch before queueing anything on it. */
static int join_batch(vp8b200_ctx *c)
{
    if (c->batch_pending) {
        CK(c, cudaStreamWaitEvent(c->stream, c->batch_ev, 0));
        c->batch_pending = false;
    }
    return VP8B200_OK;
}

/* A copy queued by fetch_begin may still be reading device buffer `fb`: order work that will
 * WRITE that buffer (on stream `s`) behind it.  Reads need no ordering. */
static int order_after_fetch(vp8b200_ctx *c, cudaStream_t s, int fb)
{
    for (int i = 0; i < 2; i++)
        if (__atomic_load_n(&c->fetch[i].fb, __ATOMIC_RELAXED) == fb && fb >= 0)
            CK(c, cudaStreamWaitEvent(s, c->fetch[i].ev, 0));
    return VP8B200_OK;
}

static bool hdr_ok(const vp8b200_ctx *c, const vp8b200_frame_hdr *h)
{
    if (!(h->fb_new < c->n_fb && h->fb_last < c->n_fb && h->fb_golden < c->n_fb &&
          h->fb_altref < c->n_fb && h->frame_type <= 1 && h->filter_type <= 1 &&
          h->filter_level <= 63 && h->sharpness_level <= 7))
        return false;
    /* an inter frame is never reconstructed into one of its own references (the reference's
     * get_free_fb, onyxd_if.c:242-259, only hands out buffers nobody refers to): k_inter would
     * read the buffer it is writing */
    if (h->frame_type == 1 && (h->fb_new == h->fb_last || h->fb_new == h->fb_golden || h->fb_new == h->fb_altref))
        return false;
    return true;
}

extern "C" int vp8b200_frame_begin(vp8b200_ctx *c, const vp8b200_frame_hdr *hdr, vp8b200_frame_bufs *bufs)
{
    if (!c || !hdr || !bufs || !hdr_ok(c, hdr)) return VP8B200_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    c->open = false;
    if ((uint32_t)c->eng_submitted - __atomic_load_n(&c->eng_issued, __ATOMIC_ACQUIRE) > (uint32_t)(NSLOT - 2)) engine_settle(c);
    Slot &s = c->slot[c->cur];
    if (s.pending) {                         /* the upload that last used this slot */
        CK(c, cudaEventSynchronize(s.h2d_done));
        s.pending = false;
    }
    c->cur_hdr = *hdr;
    bufs->mb = s.h_mb; bufs->aux = s.h_aux; bufs->coef = s.h_coef;
    bufs->aux_capacity = c->n_mb; bufs->coef_capacity = c->n_mb * 25;
    c->open = true;
    return VP8B200_OK;
}

extern "C" int vp8b200_frame_abort(vp8b200_ctx *c)
{
    if (!c) return VP8B200_ERR_INVALID;
    c->open = false;
    return VP8B200_OK;
}

/* One pass over the frame's records at submit time: validates what the kernels will index
 * with (a corrupt record must not become a wild read), counts SPLITMV macroblocks, and - for
 * P frames - builds the wavefront order of the intra macroblocks (counting sort on
 * d = col + 2*row, see wavefront_order).  Returns false on an invalid record. */
static bool scan_records(const Geo &g, const vp8b200_mb *mb, const vp8b200_aux *aux, uint32_t n_aux,
                         uint32_t n_coef, bool key, uint32_t *ilist, int *cnt, unsigned *n_intra_out,
                         unsigned *n_split_out)
{
    const int nd = g.mb_cols + 2 * g.mb_rows;
    unsigned n_intra = 0, n_split = 0;
    if (!key) for (int d = 0; d <= nd; d++) cnt[d] = 0;
    const vp8b200_mb *m = mb;
    for (int row = 0; row < g.mb_rows; row++) {
        /* range clamp_mv_to_umv_border (reconinter.c:348-368) leaves alone: an unclamped MV
         * must already be inside it, or the 32-pixel border would not cover the filter window */
        const int lo_r = -((row * 16) << 3) - (19 << 3), hi_r = (((g.mb_rows - 1 - row) * 16) << 3) + (18 << 3);
        for (int col = 0; col < g.mb_cols; col++, m++) {
            if (m->y_mode > VP8B200_SPLITMV || m->uv_mode > VP8B200_TM_PRED || m->ref_frame > 3) return false;
            const bool intra = m->ref_frame == VP8B200_INTRA_FRAME;
            if (intra != (m->y_mode <= VP8B200_B_PRED)) return false;
            if (key && !intra) return false;
            const bool has_aux = m->y_mode == VP8B200_B_PRED || m->y_mode == VP8B200_SPLITMV;
            if (has_aux && m->u.aux >= n_aux) return false;
            if (m->y_mode == VP8B200_B_PRED) {             /* k_intra indexes its predictor table with these */
                const uint8_t *bm = aux[m->u.aux].b_mode;
                for (int k = 0; k < 16; k++) if (bm[k] > VP8B200_B_HU_PRED) return false;
            }
            if (intra) {
                if (!key) cnt[col + 2 * row + 1]++;
                n_intra++;
            } else {
                const bool split = m->y_mode == VP8B200_SPLITMV;
                n_split += split;
                if (!(m->flags & VP8B200_MBF_CLAMP_MVS)) {
                    const int lo_c = -((col * 16) << 3) - (19 << 3), hi_c = (((g.mb_cols - 1 - col) * 16) << 3) + (18 << 3);
                    const int nmv = split ? 16 : 1;
                    for (int k = 0; k < nmv; k++) {
                        const int r = split ? aux[m->u.aux].mv[k].row : m->u.mv.row;
                        const int c = split ? aux[m->u.aux].mv[k].col : m->u.mv.col;
                        if (r < lo_r || r > hi_r || c < lo_c || c > hi_c) return false;
                    }
                }
            }
            if (m->coef_mask >> 25) return false;
            if (m->coef_mask && (uint64_t)m->coef_off + (unsigned)__builtin_popcount(m->coef_mask) > n_coef) return false;
        }
    }
    if (!key && n_intra) {
        for (int d = 0; d < nd; d++) cnt[d + 1] += cnt[d];
        m = mb;
        uint32_t i = 0;
        for (int row = 0; row < g.mb_rows; row++)
            for (int col = 0; col < g.mb_cols; col++, m++, i++)
                if (m->ref_frame == VP8B200_INTRA_FRAME) ilist[cnt[col + 2 * row]++] = i;
    }
    *n_intra_out = n_intra;
    *n_split_out = n_split;
    return true;
}

static void fill_job(vp8b200_ctx *c, FrameJob *j, const vp8b200_frame_hdr &h, const vp8b200_mb *d_mb,
                     const vp8b200_aux *d_aux, const int16_t *d_coef, const uint32_t *d_ilist,
                     unsigned n_intra, unsigned n_split, bool run_intra, bool run_lf)
{
    memset(j, 0, sizeof *j);
    j->dst = c->fb[h.fb_new];
    j->ref[1] = c->fb[h.fb_last]; j->ref[2] = c->fb[h.fb_golden]; j->ref[3] = c->fb[h.fb_altref];
    j->mb = d_mb; j->aux = d_aux; j->coef = d_coef;
    j->intra_msg = c->d_imsg;
    j->intra_list = d_ilist ? d_ilist : c->d_diag;
    j->lf_msg = c->d_lfmsg;
    if (run_intra) c->epoch_intra++;
    if (run_lf) c->epoch_lf++;
    j->epoch_intra = c->epoch_intra;
    j->epoch_lf = c->epoch_lf;
    j->n_intra = n_intra;
    j->n_split = n_split;
    j->hdr = h;
}

/* the launch sequence shared by frame_submit (n = 1) and batch_run */
static void prof_mark(vp8b200_ctx *c, int kind, bool begin)
{
    if (!c->profiling) return;
    if (begin) {
        if (c->spans->size() >= MAX_SPANS) { c->prof_skip = true; return; }   /* nobody reads: stop collecting */
        c->prof_skip = false;
        ProfSpan sp; sp.kind = kind;
        cudaEventCreate(&sp.a); cudaEventCreate(&sp.b);
        cudaEventRecord(sp.a, c->stream);
        c->spans->push_back(sp);
    } else if (!c->prof_skip) {
        cudaEventRecord(c->spans->back().b, c->stream);
    }
}

/* `st` / `tickets` / `ticket_base`: the launching stream and its two wavefront ticket counters
 * (the context's own for frame_submit and batch_run, the engine's for coalesced submits) */
static int run_jobs_on(vp8b200_ctx *c, cudaStream_t st, unsigned *tickets, unsigned *ticket_base,
                       const FrameJob *d_jobs, int n, bool any_inter, bool any_split,
                       unsigned max_intra, bool any_lf, bool prof)
{
    const bool any_intra = max_intra > 0;
    int nctas = 0;
    unsigned k = 0;
    if (any_inter) {
        if (prof) prof_mark(c, 0, true);
        vp8b200_launch_inter(st, d_jobs, n, c->geo, any_split);
        if (prof) prof_mark(c, 0, false);
        k += any_split ? 2 : 1;
    }
    if (any_intra) {
        if (prof) prof_mark(c, 1, true);
        vp8b200_launch_intra(st, d_jobs, n, c->geo, max_intra, tickets + 0, ticket_base[0], &nctas);
        if (prof) prof_mark(c, 1, false);
        ticket_base[0] += (unsigned)nctas; k++;
    }
    if (any_lf) {
        if (prof) prof_mark(c, 2, true);
        vp8b200_launch_loopfilter(st, d_jobs, n, c->geo, tickets + 1, ticket_base[1], &nctas);
        if (prof) prof_mark(c, 2, false);
        ticket_base[1] += (unsigned)nctas; k++;
    }
    if (prof) prof_mark(c, 3, true);
    vp8b200_launch_border(st, d_jobs, n, c->geo);
    if (prof) prof_mark(c, 3, false);
    k++;
    c->launches += k;
    g_launches += k;
    g_frames += (uint64_t)n;
    CK(c, cudaGetLastError());
    return VP8B200_OK;
}

static int run_jobs(vp8b200_ctx *c, const FrameJob *d_jobs, int n, bool any_inter, bool any_split,
                    unsigned max_intra, bool any_lf)
{
    return run_jobs_on(c, c->stream, c->d_tickets, c->ticket_base, d_jobs, n, any_inter, any_split, max_intra, any_lf, true);
}

extern "C" int vp8b200_frame_submit(vp8b200_ctx *c, uint32_t n_aux, uint32_t n_coef)
{
    if (!c || !c->open) return VP8B200_ERR_INVALID;
    engine_settle(c);
    c->open = false;
    if (n_aux > c->n_mb || n_coef > c->n_mb * 25) return VP8B200_ERR_OVERFLOW;
    CK(c, cudaSetDevice(c->device));
    { int js = join_batch(c); if (js) return js; }
    c->own_dirty = true;
    Slot &s = c->slot[c->cur];
    const vp8b200_frame_hdr &h = c->cur_hdr;
    { int os = order_after_fetch(c, c->stream, h.fb_new); if (os) return os; }
    const bool key = h.frame_type == 0;
    /* key frames use the context's static wavefront order; P frames list their intra MBs */
    unsigned n_intra = 0, n_split = 0;
    if (!scan_records(c->geo, s.h_mb, s.h_aux, n_aux, n_coef, key, s.h_ilist, c->diag_tmp, &n_intra, &n_split)) {
        snprintf(c->err, sizeof c->err, "macroblock records failed validation");
        return VP8B200_ERR_INVALID;
    }
    const bool run_intra = n_intra > 0, run_lf = h.filter_level != 0;
    fill_job(c, s.h_job, h, s.d_mb, s.d_aux, s.d_coef, key ? NULL : s.d_ilist, n_intra, n_split, run_intra, run_lf);
    if (!key && n_intra)
        CK(c, cudaMemcpyAsync(s.d_ilist, s.h_ilist, (size_t)n_intra * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemcpyAsync(s.d_mb, s.h_mb, (size_t)c->n_mb * sizeof(vp8b200_mb), cudaMemcpyHostToDevice, c->stream));
    if (n_aux) CK(c, cudaMemcpyAsync(s.d_aux, s.h_aux, (size_t)n_aux * sizeof(vp8b200_aux), cudaMemcpyHostToDevice, c->stream));
    if (n_coef) CK(c, cudaMemcpyAsync(s.d_coef, s.h_coef, (size_t)n_coef * 32, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemcpyAsync(s.d_job, s.h_job, sizeof(FrameJob), cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaEventRecord(s.h2d_done, c->stream));
    s.pending = true;
    g_h2d_bytes += (uint64_t)c->n_mb * sizeof(vp8b200_mb) + (uint64_t)n_aux * sizeof(vp8b200_aux) +
                   (uint64_t)n_coef * 32 + sizeof(FrameJob) + (key ? 0 : (uint64_t)n_intra * 4);
    int st = run_jobs(c, s.d_job, 1, !key, n_split > 0, n_intra, run_lf);
    c->cur = (c->cur + 1) % NSLOT;
    return st;
}

/* SURVEY 8f N2 / N3.  display_w == 0: the whole allocation (borders included); otherwise only
 * what a caller of vpx_codec_get_frame can see - display_w x display_h luma and
 * ((w+1)/2) x ((h+1)/2) chroma samples (vpxdec.c:1093-1115) - as three pitched copies that land
 * at the same offsets as in the device buffer, so img->planes[] / stride[] of the host mirror
 * stay valid (vp8_dx_iface.c:319-348). */

/* ---- engine: queue, gather, issue ---------------------------------------------------------- */

static Engine *engine_get(vp8b200_ctx *c)
{
    std::lock_guard<std::mutex> lk(g_eng_mu);
    if (c->device < 0 || c->device >= 64) return NULL;
    Engine *e = g_engines[c->device];
    if (!e) {
        e = new (std::nothrow) Engine();
        if (!e) return NULL;
        e->device = c->device;
        e->n_ctx = 0; e->started = false; e->failed = false; e->cap = 0; e->cur = 0;
        for (int i = 0; i < ENG_RING; i++) { e->h_jobs[i] = NULL; e->d_jobs[i] = NULL; e->ring_pending[i] = false; }
        const char *w = getenv("VP8B200_BATCH_WINDOW_US");
        e->window_us = w ? atoi(w) : 1000;
        const char *m = getenv("VP8B200_BATCH_MAX");
        e->max_batch = m ? atoi(m) : 64;
        if (e->max_batch < 1) e->max_batch = 1;
        g_engines[c->device] = e;
    }
    return e;
}

/* wait until the engine thread has issued everything this context handed over */
static void engine_settle(vp8b200_ctx *c)
{
    /* acquire: everything the engine thread wrote into the context while issuing is visible.
     * The wait is a futex on the context's own word: a batch wakes exactly its members. */
    if (!c || !c->eng) return;
    const uint32_t want = (uint32_t)c->eng_submitted;
    for (;;) {
        const uint32_t have = __atomic_load_n(&c->eng_issued, __ATOMIC_ACQUIRE);
        if (have == want) return;
        syscall(SYS_futex, &c->eng_issued, FUTEX_WAIT_PRIVATE, have, NULL, NULL, 0);
    }
}

/* wait until the engine thread has issued this context's submit number `seq` */
static void engine_wait_seq(vp8b200_ctx *c, uint32_t seq)
{
    if (!c->eng) return;
    for (;;) {
        const uint32_t have = __atomic_load_n(&c->eng_issued, __ATOMIC_ACQUIRE);
        if ((int32_t)(have - seq) >= 0) return;
        syscall(SYS_futex, &c->eng_issued, FUTEX_WAIT_PRIVATE, have, NULL, NULL, 0);
    }
}

/* the oldest copy in flight is in host memory */
static int fetch_retire_oldest(vp8b200_ctx *c)
{
    if (c->fetch_n == 0) return VP8B200_OK;
    const int i = c->fetch_head;
    if (c->fetch[i].seq) engine_wait_seq(c, c->fetch[i].seq);   /* its event is recorded by the engine thread */
    c->fetch_head = (i + 1) & 1;
    c->fetch_n--;
    if (c->eng_status) { __atomic_store_n(&c->fetch[i].fb, -1, __ATOMIC_RELAXED); int st = c->eng_status; c->eng_status = 0; return st; }
    CK(c, cudaSetDevice(c->device));
    cudaError_t e = cudaEventSynchronize(c->fetch[i].ev);
    __atomic_store_n(&c->fetch[i].fb, -1, __ATOMIC_RELAXED);
    CK(c, e);
    return VP8B200_OK;
}

/* a record for a new copy of buffer fb (the caller never collected two older ones: the oldest is
 * waited for here) */
static int fetch_reserve(vp8b200_ctx *c, int fb, uint32_t seq, int *rec)
{
    if (c->fetch_n == 2) { int st = fetch_retire_oldest(c); if (st) return st; }
    const int i = (c->fetch_head + c->fetch_n) & 1;
    c->fetch[i].seq = seq;
    __atomic_store_n(&c->fetch[i].fb, fb, __ATOMIC_RELAXED);
    c->fetch_n++;
    *rec = i;
    return VP8B200_OK;
}

#define ECK(call)                                                                          \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            snprintf(err, sizeof err, "%s: %s", #call, cudaGetErrorString(e_));           \
            return VP8B200_ERR_CUDA;                                                       \
        }                                                                                  \
    } while (0)

static int engine_init_device(Engine *e, char (&err)[256])
{
    ECK(cudaSetDevice(e->device));
    for (int i = 0; i < 2; i++) {
        ECK(cudaStreamCreateWithFlags(&e->stream[i], cudaStreamNonBlocking));
        ECK(cudaMalloc((void **)&e->d_tickets[i], 2 * sizeof(unsigned)));
        ECK(cudaMemset(e->d_tickets[i], 0, 2 * sizeof(unsigned)));
        e->ticket_base[i][0] = e->ticket_base[i][1] = 0;
    }
    ECK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < ENG_RING; i++) {
        ECK(cudaEventCreateWithFlags(&e->jobs_ev[i], cudaEventDisableTiming));
        ECK(cudaEventCreateWithFlags(&e->batch_ev[i], cudaEventDisableTiming));
    }
    return VP8B200_OK;
}

/* one batch: frames of DISTINCT contexts with the same geometry */
static int engine_issue(Engine *e, std::vector<EngineSubmit> &b, char (&err)[256])
{
    const int n = (int)b.size();
    vp8b200_ctx *c0 = b[0].c;
    if (n > e->cap) {
        ECK(cudaDeviceSynchronize());
        for (int i = 0; i < ENG_RING; i++) {
            cudaFreeHost(e->h_jobs[i]); cudaFree(e->d_jobs[i]);
            e->h_jobs[i] = NULL; e->d_jobs[i] = NULL; e->ring_pending[i] = false;
            ECK(cudaHostAlloc((void **)&e->h_jobs[i], (size_t)n * sizeof(FrameJob), cudaHostAllocPortable));
            ECK(cudaMalloc((void **)&e->d_jobs[i], (size_t)n * sizeof(FrameJob)));
        }
        e->cap = n;
    }
    const int r = e->cur;
    e->cur = (r + 1) % ENG_RING;
    const int si = r & 1;                                  /* consecutive batches alternate streams: */
    cudaStream_t st = e->stream[si];                       /* the copies of one overlap the kernels of the other */
    if (e->ring_pending[r]) { ECK(cudaEventSynchronize(e->jobs_ev[r])); e->ring_pending[r] = false; }
    bool any_inter = false, any_lf = false, any_split = false;
    unsigned max_intra = 0;
    for (int i = 0; i < n; i++) {
        EngineSubmit &sb = b[i];
        vp8b200_ctx *c = sb.c;
        Slot &s = c->slot[sb.slot];
        const vp8b200_frame_hdr &h = sb.hdr;
        const bool key = h.frame_type == 0;
        /* order behind whatever last touched this context: a batch on the other engine stream or
         * under a batch_run leader, work on its own stream, a copy still reading the new buffer */
        if (c->batch_pending) { ECK(cudaStreamWaitEvent(st, c->batch_ev, 0)); c->batch_pending = false; }
        if (c->own_dirty) {
            ECK(cudaEventRecord(c->own_ev, c->stream));
            ECK(cudaStreamWaitEvent(st, c->own_ev, 0));
            c->own_dirty = false;
        }
        for (int k = 0; k < 2; k++)                            /* an OLDER copy may still be reading the buffer we write */
            if (__atomic_load_n(&c->fetch[k].fb, __ATOMIC_RELAXED) == (int)h.fb_new && !(sb.show_fb >= 0 && k == sb.show_rec))
                ECK(cudaStreamWaitEvent(st, c->fetch[k].ev, 0));
        if (!key && sb.n_intra)
            ECK(cudaMemcpyAsync(s.d_ilist, s.h_ilist, (size_t)sb.n_intra * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        ECK(cudaMemcpyAsync(s.d_mb, s.h_mb, (size_t)c->n_mb * sizeof(vp8b200_mb), cudaMemcpyHostToDevice, st));
        if (sb.n_aux) ECK(cudaMemcpyAsync(s.d_aux, s.h_aux, (size_t)sb.n_aux * sizeof(vp8b200_aux), cudaMemcpyHostToDevice, st));
        if (sb.n_coef) ECK(cudaMemcpyAsync(s.d_coef, s.h_coef, (size_t)sb.n_coef * 32, cudaMemcpyHostToDevice, st));
        ECK(cudaEventRecord(s.h2d_done, st));
        s.pending = true;
        g_h2d_bytes += (uint64_t)c->n_mb * sizeof(vp8b200_mb) + (uint64_t)sb.n_aux * sizeof(vp8b200_aux) +
                       (uint64_t)sb.n_coef * 32 + sizeof(FrameJob) + (key ? 0 : (uint64_t)sb.n_intra * 4);
        any_inter |= !key; any_lf |= h.filter_level != 0; any_split |= sb.n_split > 0;
        if (sb.n_intra > max_intra) max_intra = sb.n_intra;
    }
    for (int i = 0; i < n; i++) {
        EngineSubmit &sb = b[i];
        Slot &s = sb.c->slot[sb.slot];
        fill_job(sb.c, &e->h_jobs[r][i], sb.hdr, s.d_mb, s.d_aux, s.d_coef, sb.hdr.frame_type == 0 ? NULL : s.d_ilist,
                 sb.n_intra, sb.n_split, sb.n_intra > 0, any_lf && sb.hdr.filter_level != 0);
    }
    ECK(cudaMemcpyAsync(e->d_jobs[r], e->h_jobs[r], (size_t)n * sizeof(FrameJob), cudaMemcpyHostToDevice, st));
    ECK(cudaEventRecord(e->jobs_ev[r], st));
    e->ring_pending[r] = true;
    {
        /* launch bookkeeping goes to the first member (vp8b200_launch_count is per context) */
        int rs = run_jobs_on(c0, st, e->d_tickets[si], e->ticket_base[si], e->d_jobs[r], n, any_inter, any_split, max_intra, any_lf, false);
        if (rs) { snprintf(err, sizeof err, "%s", c0->err); return rs; }
    }
    ECK(cudaEventRecord(e->batch_ev[r], st));
    bool any_show = false;
    for (int i = 0; i < n; i++) {
        vp8b200_ctx *c = b[i].c;
        c->batch_ev = e->batch_ev[r]; c->batch_pending = true; c->batch_leader = NULL;
        any_show |= b[i].show_fb >= 0;
    }
    if (any_show) {
        ECK(cudaStreamWaitEvent(e->copy_stream, e->batch_ev[r], 0));
        for (int i = 0; i < n; i++) {
            EngineSubmit &sb = b[i];
            if (sb.show_fb < 0) continue;
            vp8b200_ctx *c = sb.c;
            const Geo &g = c->geo;
            if (sb.show_w == 0) {
                ECK(cudaMemcpyAsync(sb.show_dst, c->fb[sb.show_fb], c->frame_size, cudaMemcpyDeviceToHost, e->copy_stream));
                g_d2h_bytes += c->frame_size;
            } else {
                const int cw = (sb.show_w + 1) >> 1, ch = (sb.show_h + 1) >> 1;
                ECK(cudaMemcpy2DAsync(sb.show_dst + g.y_off, (size_t)g.y_stride, c->fb[sb.show_fb] + g.y_off, (size_t)g.y_stride,
                                      (size_t)sb.show_w, (size_t)sb.show_h, cudaMemcpyDeviceToHost, e->copy_stream));
                ECK(cudaMemcpy2DAsync(sb.show_dst + g.u_off, (size_t)g.uv_stride, c->fb[sb.show_fb] + g.u_off, (size_t)g.uv_stride,
                                      (size_t)cw, (size_t)ch, cudaMemcpyDeviceToHost, e->copy_stream));
                ECK(cudaMemcpy2DAsync(sb.show_dst + g.v_off, (size_t)g.uv_stride, c->fb[sb.show_fb] + g.v_off, (size_t)g.uv_stride,
                                      (size_t)cw, (size_t)ch, cudaMemcpyDeviceToHost, e->copy_stream));
                g_d2h_bytes += (uint64_t)sb.show_w * sb.show_h + 2ull * cw * ch;
            }
            ECK(cudaEventRecord(c->fetch[sb.show_rec].ev, e->copy_stream));
        }
    }
    e->batches++;
    e->frames += (uint64_t)n;
    return VP8B200_OK;
}

static void engine_thread(Engine *e)
{
    char err[256] = "";
    /* the engine sleeps most of the time and must run the moment a frame is queued, also on a
     * host whose cores are all busy parsing: ask for a better scheduling priority (needs
     * CAP_SYS_NICE; silently stays at the default otherwise) */
    setpriority(PRIO_PROCESS, (id_t)syscall(SYS_gettid), -10);
    int init = engine_init_device(e, err);
    std::vector<EngineSubmit> batch, rest;
    for (;;) {
        batch.clear();
        {
            std::unique_lock<std::mutex> lk(e->mu);
            e->cv_work.wait(lk, [&] { return !e->q.empty(); });
            /* gather: with several contexts alive wait (bounded) until about half of them have a
             * frame queued; a lone context is issued at once */
            const int target = std::min(e->max_batch, std::max(1, e->n_ctx / 2));
            if ((int)e->q.size() < target && e->window_us > 0) {
                const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(e->window_us);
                e->cv_work.wait_until(lk, deadline, [&] { return (int)e->q.size() >= target; });
            }
            /* at most one frame per context (a later frame depends on the earlier one), one geometry */
            rest.clear();
            const vp8b200_ctx *g0 = e->q.front().c;
            while (!e->q.empty()) {
                EngineSubmit sb = e->q.front();
                e->q.pop_front();
                bool take = (int)batch.size() < e->max_batch && sb.c->geo.width == g0->geo.width && sb.c->geo.height == g0->geo.height;
                for (size_t k = 0; take && k < batch.size(); k++) if (batch[k].c == sb.c) take = false;
                for (size_t k = 0; take && k < rest.size(); k++) if (rest[k].c == sb.c) take = false;   /* keep per-context order */
                (take ? batch : rest).push_back(sb);
            }
            for (auto &sb : rest) e->q.push_back(sb);
        }
        int st = init ? init : (e->failed ? VP8B200_ERR_CUDA : engine_issue(e, batch, err));
        if (st && !init) e->failed = true;                 /* a CUDA failure is sticky for the device */
        {
            std::lock_guard<std::mutex> lk(e->mu);
            for (auto &sb : batch) {
                if (st) { sb.c->eng_status = st; snprintf(sb.c->err, sizeof sb.c->err, "engine: %s", err); }
                __atomic_store_n(&sb.c->eng_issued, (uint32_t)sb.seq, __ATOMIC_RELEASE);
                syscall(SYS_futex, &sb.c->eng_issued, FUTEX_WAKE_PRIVATE, 1, NULL, NULL, 0);
            }
        }
    }
}

/* Coalesced submit: validate on the caller's thread, hand the frame to the
 * device's engine, return.  show_fb >= 0 also queues the device->host copy of that buffer (as
 * vp8b200_frame_fetch_begin would: display_w/h = 0 -> whole allocation) behind the frame's
 * kernels; vp8b200_frame_fetch_wait is then the caller's only wait. */
extern "C" int vp8b200_frame_submit_show(vp8b200_ctx *c, uint32_t n_aux, uint32_t n_coef,
                                         int show_fb, uint8_t *dst, int display_w, int display_h)
{
    if (!c || !c->open) return VP8B200_ERR_INVALID;
    c->open = false;
    if (n_aux > c->n_mb || n_coef > c->n_mb * 25) return VP8B200_ERR_OVERFLOW;
    if (show_fb >= c->n_fb || (show_fb >= 0 && (!dst || display_w < 0 || display_h < 0 || display_w > c->geo.width ||
                                                 display_h > c->geo.height || (display_w == 0) != (display_h == 0))))
        return VP8B200_ERR_INVALID;
    if (c->eng_status) { int st = c->eng_status; c->eng_status = 0; return st; }
    if (!c->eng) {
        Engine *e = engine_get(c);
        if (!e) return VP8B200_ERR_NOMEM;
        std::lock_guard<std::mutex> lk(e->mu);
        if (!e->started) { e->started = true; std::thread(engine_thread, e).detach(); }
        e->n_ctx++;
        c->eng = e;
    }
    Slot &s = c->slot[c->cur];
    EngineSubmit sb;
    sb.c = c; sb.slot = c->cur; sb.n_aux = n_aux; sb.n_coef = n_coef; sb.hdr = c->cur_hdr;
    sb.show_fb = show_fb; sb.show_dst = dst; sb.show_w = display_w; sb.show_h = display_h;
    if (!scan_records(c->geo, s.h_mb, s.h_aux, n_aux, n_coef, sb.hdr.frame_type == 0, s.h_ilist, c->diag_tmp, &sb.n_intra, &sb.n_split)) {
        snprintf(c->err, sizeof c->err, "macroblock records failed validation");
        return VP8B200_ERR_INVALID;
    }
    c->cur = (c->cur + 1) % NSLOT;
    sb.show_rec = 0;
    if (show_fb >= 0) {
        int st = fetch_reserve(c, show_fb, (uint32_t)(c->eng_submitted + 1), &sb.show_rec);
        if (st) return st;
    }
    {
        std::lock_guard<std::mutex> lk(c->eng->mu);
        sb.seq = ++c->eng_submitted;
        c->eng->q.push_back(sb);
    }
    c->eng->cv_work.notify_one();
    return VP8B200_OK;
}

/* engine statistics of the context's device: [0] batches issued, [1] frames issued */
extern "C" void vp8b200_engine_stats(int device, uint64_t out[2])
{
    out[0] = out[1] = 0;
    std::lock_guard<std::mutex> lk(g_eng_mu);
    if (device >= 0 && device < 64 && g_engines[device]) {
        out[0] = g_engines[device]->batches.load(); out[1] = g_engines[device]->frames.load();
    }
}

extern "C" int vp8b200_frame_fetch_begin(vp8b200_ctx *c, int fb, uint8_t *dst, int display_w, int display_h)
{
    if (!c || fb < 0 || fb >= c->n_fb || !dst || display_w < 0 || display_h < 0 ||
        display_w > c->geo.width || display_h > c->geo.height || (display_w == 0) != (display_h == 0))
        return VP8B200_ERR_INVALID;
    engine_settle(c);
    CK(c, cudaSetDevice(c->device));
    { int js = join_batch(c); if (js) return js; }
    int rec = 0;
    { int rs = fetch_reserve(c, fb, 0, &rec); if (rs) return rs; }
    const Geo &g = c->geo;
    CK(c, cudaEventRecord(c->recon_done, c->stream));
    CK(c, cudaStreamWaitEvent(c->copy_stream, c->recon_done, 0));
    if (display_w == 0) {
        CK(c, cudaMemcpyAsync(dst, c->fb[fb], c->frame_size, cudaMemcpyDeviceToHost, c->copy_stream));
        g_d2h_bytes += c->frame_size;
    } else {
        const int cw = (display_w + 1) >> 1, ch = (display_h + 1) >> 1;
        CK(c, cudaMemcpy2DAsync(dst + g.y_off, (size_t)g.y_stride, c->fb[fb] + g.y_off, (size_t)g.y_stride,
                                (size_t)display_w, (size_t)display_h, cudaMemcpyDeviceToHost, c->copy_stream));
        CK(c, cudaMemcpy2DAsync(dst + g.u_off, (size_t)g.uv_stride, c->fb[fb] + g.u_off, (size_t)g.uv_stride,
                                (size_t)cw, (size_t)ch, cudaMemcpyDeviceToHost, c->copy_stream));
        CK(c, cudaMemcpy2DAsync(dst + g.v_off, (size_t)g.uv_stride, c->fb[fb] + g.v_off, (size_t)g.uv_stride,
                                (size_t)cw, (size_t)ch, cudaMemcpyDeviceToHost, c->copy_stream));
        g_d2h_bytes += (uint64_t)display_w * display_h + 2ull * cw * ch;
    }
    CK(c, cudaEventRecord(c->fetch[rec].ev, c->copy_stream));
    return VP8B200_OK;
}

/* Wait until the OLDEST copy in flight (fetch_begin / frame_submit_show) is in host memory: up
 * to two are tracked, so a caller in frame-delay mode collects picture N-1 while picture N is
 * still being reconstructed.  Device faults of the frame's kernels surface here (or at the
 * next call) as VP8B200_ERR_CUDA. */
extern "C" int vp8b200_frame_fetch_wait(vp8b200_ctx *c)
{
    if (!c) return VP8B200_ERR_INVALID;
    if (c->fetch_n == 0) {
        /* nothing in flight: still the place where an engine failure of an unshown frame surfaces */
        engine_settle(c);
        if (c->eng_status) { int st = c->eng_status; c->eng_status = 0; return st; }
        return VP8B200_OK;
    }
    return fetch_retire_oldest(c);
}

extern "C" int vp8b200_frame_fetch(vp8b200_ctx *c, int fb, uint8_t *dst, size_t bytes)
{
    if (!c || fb < 0 || fb >= c->n_fb || !dst || bytes > c->frame_size) return VP8B200_ERR_INVALID;
    engine_settle(c);
    if (bytes == c->frame_size) {
        /* blocking: every copy in flight, this one last, is in host memory on return */
        int st = vp8b200_frame_fetch_begin(c, fb, dst, 0, 0);
        while (!st && c->fetch_n) st = fetch_retire_oldest(c);
        return st;
    }
    CK(c, cudaSetDevice(c->device));
    { int js = join_batch(c); if (js) return js; }
    CK(c, cudaMemcpyAsync(dst, c->fb[fb], bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    g_d2h_bytes += bytes;
    return VP8B200_OK;
}

extern "C" int vp8b200_frame_upload(vp8b200_ctx *c, int fb, const uint8_t *src, size_t bytes)
{
    if (!c || fb < 0 || fb >= c->n_fb || !src || bytes > c->frame_size) return VP8B200_ERR_INVALID;
    engine_settle(c);
    CK(c, cudaSetDevice(c->device));
    { int js = join_batch(c); if (js) return js; }
    { int os = order_after_fetch(c, c->stream, fb); if (os) return os; }
    c->own_dirty = true;
    CK(c, cudaMemcpyAsync(c->fb[fb], src, bytes, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return VP8B200_OK;
}

extern "C" int vp8b200_frame_copy(vp8b200_ctx *c, int fb_dst, int fb_src)
{
    if (!c || fb_dst < 0 || fb_dst >= c->n_fb || fb_src < 0 || fb_src >= c->n_fb) return VP8B200_ERR_INVALID;
    if (fb_dst == fb_src) return VP8B200_OK;
    engine_settle(c);
    CK(c, cudaSetDevice(c->device));
    { int js = join_batch(c); if (js) return js; }
    { int os = order_after_fetch(c, c->stream, fb_dst); if (os) return os; }
    c->own_dirty = true;
    CK(c, cudaMemcpyAsync(c->fb[fb_dst], c->fb[fb_src], c->frame_size, cudaMemcpyDeviceToDevice, c->stream));
    return VP8B200_OK;
}

extern "C" int vp8b200_sync(vp8b200_ctx *c)
{
    if (!c) return VP8B200_ERR_INVALID;
    engine_settle(c);
    CK(c, cudaSetDevice(c->device));
    { int js = join_batch(c); if (js) return js; }
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaStreamSynchronize(c->copy_stream));
    c->own_dirty = false;
    while (c->fetch_n) { int st = fetch_retire_oldest(c); if (st) return st; }
    return VP8B200_OK;
}

/* ---- resident frames and batched replay ------------------------------------------------- */

extern "C" int vp8b200_stage_frame(vp8b200_ctx *c, const vp8b200_frame_hdr *hdr, const vp8b200_mb *mb,
                                   const vp8b200_aux *aux, uint32_t n_aux, const int16_t *coef,
                                   uint32_t n_coef, vp8b200_staged **out)
{
    if (!c || !hdr || !mb || !out || !hdr_ok(c, hdr) || (n_aux && !aux) || (n_coef && !coef))
        return VP8B200_ERR_INVALID;
    if (n_aux > c->n_mb || n_coef > c->n_mb * 25) return VP8B200_ERR_OVERFLOW;
    const bool key = hdr->frame_type == 0;
    uint32_t *ilist = NULL;
    unsigned n_intra = 0, n_split = 0;
    if (!key) {
        ilist = (uint32_t *)malloc((size_t)c->n_mb * sizeof(uint32_t));
        if (!ilist) return VP8B200_ERR_NOMEM;
    }
    if (!scan_records(c->geo, mb, aux, n_aux, n_coef, key, ilist, c->diag_tmp, &n_intra, &n_split)) {
        snprintf(c->err, sizeof c->err, "macroblock records failed validation");
        free(ilist);
        return VP8B200_ERR_INVALID;
    }
    { cudaError_t e0 = cudaSetDevice(c->device); if (e0 != cudaSuccess) { free(ilist); CK(c, e0); } }
    vp8b200_staged *s = new (std::nothrow) vp8b200_staged();
    if (!s) { free(ilist); return VP8B200_ERR_NOMEM; }
    const size_t mb_b = (size_t)c->n_mb * sizeof(vp8b200_mb);
    const size_t aux_b = ((size_t)n_aux * sizeof(vp8b200_aux) + 255) & ~(size_t)255;
    const size_t coef_b = (size_t)n_coef * 32;
    const size_t mb_pad = (mb_b + 255) & ~(size_t)255;
    const size_t coef_pad = (coef_b + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc((void **)&s->d_blob, mb_pad + aux_b + coef_pad + (size_t)n_intra * 4 + 256);
    if (e != cudaSuccess) {
        snprintf(c->err, sizeof c->err, "cudaMalloc(staged): %s", cudaGetErrorString(e));
        free(ilist);
        delete s;
        return VP8B200_ERR_CUDA;
    }
    s->hdr = *hdr;
    s->d_mb = (vp8b200_mb *)s->d_blob;
    s->d_aux = (vp8b200_aux *)(s->d_blob + mb_pad);
    s->d_coef = (int16_t *)(s->d_blob + mb_pad + aux_b);
    s->n_intra = n_intra;
    s->n_split = n_split;
    s->d_ilist = key ? NULL : (uint32_t *)(s->d_blob + mb_pad + aux_b + coef_pad);
    if (!key && n_intra) {
        cudaError_t e2 = cudaMemcpy(s->d_ilist, ilist, (size_t)n_intra * 4, cudaMemcpyHostToDevice);
        free(ilist);
        CK(c, e2);
    } else {
        free(ilist);
    }
    /* staging is a setup-time operation: plain synchronous copies from pageable memory */
    CK(c, cudaMemcpy(s->d_mb, mb, mb_b, cudaMemcpyHostToDevice));
    if (n_aux) CK(c, cudaMemcpy(s->d_aux, aux, (size_t)n_aux * sizeof(vp8b200_aux), cudaMemcpyHostToDevice));
    if (n_coef) CK(c, cudaMemcpy(s->d_coef, coef, coef_b, cudaMemcpyHostToDevice));
    *out = s;
    return VP8B200_OK;
}

extern "C" void vp8b200_staged_free(vp8b200_ctx *c, vp8b200_staged *s)
{
    if (!s) return;
    /* a staged frame can be in flight on any leader's stream: teardown path, wait for the device */
    if (c) { cudaSetDevice(c->device); cudaDeviceSynchronize(); }
    cudaFree(s->d_blob);
    delete s;
}

extern "C" int vp8b200_batch_run(vp8b200_ctx *const *ctx, vp8b200_staged *const *frame, int n)
{
    if (!ctx || !frame || n < 1 || !ctx[0]) return VP8B200_ERR_INVALID;
    vp8b200_ctx *c = ctx[0];
    for (int i = 0; i < n; i++) {
        if (!ctx[i] || !frame[i] || ctx[i]->device != c->device ||
            ctx[i]->geo.width != c->geo.width || ctx[i]->geo.height != c->geo.height)
            return VP8B200_ERR_INVALID;
        for (int k = 0; k < i; k++) if (ctx[k] == ctx[i]) return VP8B200_ERR_INVALID;
    }
    for (int i = 0; i < n; i++) engine_settle(ctx[i]);
    CK(c, cudaSetDevice(c->device));
    if (n > c->bjobs_cap) {
        CK(c, cudaStreamSynchronize(c->stream));
        for (int i = 0; i < NBJOB; i++) {
            cudaFreeHost(c->h_bjobs[i]); cudaFree(c->d_bjobs[i]);
            c->h_bjobs[i] = NULL; c->d_bjobs[i] = NULL; c->bjobs_pending[i] = false;
            CK(c, cudaHostAlloc((void **)&c->h_bjobs[i], (size_t)n * sizeof(FrameJob), cudaHostAllocPortable));
            CK(c, cudaMalloc((void **)&c->d_bjobs[i], (size_t)n * sizeof(FrameJob)));
        }
        c->bjobs_cap = n;
    }
    const int r = c->bjobs_cur;
    if (c->bjobs_pending[r]) { CK(c, cudaEventSynchronize(c->bjobs_done[r])); c->bjobs_pending[r] = false; }
    /* order the batch after whatever the members queued on their own streams, and after the
     * batch (possibly under another leader) that last touched them */
    { int js = join_batch(c); if (js) return js; }
    { int os = order_after_fetch(c, c->stream, frame[0]->hdr.fb_new); if (os) return os; }
    for (int i = 1; i < n; i++) {
        vp8b200_ctx *m = ctx[i];
        if (m->batch_pending && m->batch_leader != c)       /* same leader = same stream: already ordered */
            CK(c, cudaStreamWaitEvent(c->stream, m->batch_ev, 0));
        { int os = order_after_fetch(m, c->stream, frame[i]->hdr.fb_new); if (os) return os; }
        if (m->own_dirty) {
            CK(c, cudaEventRecord(m->own_ev, m->stream));
            CK(c, cudaStreamWaitEvent(c->stream, m->own_ev, 0));
            m->own_dirty = false;
        }
    }
    bool any_inter = false, any_lf = false, any_split = false;
    unsigned max_intra = 0;
    for (int i = 0; i < n; i++) {
        const vp8b200_staged *s = frame[i];
        any_inter |= s->hdr.frame_type != 0; any_lf |= s->hdr.filter_level != 0; any_split |= s->n_split > 0;
        if (s->n_intra > max_intra) max_intra = s->n_intra;
    }
    for (int i = 0; i < n; i++) {
        const vp8b200_staged *s = frame[i];
        /* a context's intra / loop-filter epoch advances only when that kernel really has
         * work for it (done flags and message tags of idle jobs stay untouched) */
        fill_job(ctx[i], &c->h_bjobs[r][i], s->hdr, s->d_mb, s->d_aux, s->d_coef, s->d_ilist, s->n_intra,
                 s->n_split, s->n_intra > 0, any_lf && s->hdr.filter_level != 0);
    }
    CK(c, cudaMemcpyAsync(c->d_bjobs[r], c->h_bjobs[r], (size_t)n * sizeof(FrameJob), cudaMemcpyHostToDevice, c->stream));
    g_h2d_bytes += (uint64_t)n * sizeof(FrameJob);
    CK(c, cudaEventRecord(c->bjobs_done[r], c->stream));
    c->bjobs_pending[r] = true;
    c->bjobs_cur = (r + 1) % NBJOB;
    int st = run_jobs(c, c->d_bjobs[r], n, any_inter, any_split, max_intra, any_lf);
    if (st) return st;
    CK(c, cudaEventRecord(c->lead_ev[r], c->stream));
    /* the leader's own buffers were written on its own stream: if it later joins a batch under
     * another leader, that batch has to wait for this one */
    c->own_dirty = true;
    for (int i = 1; i < n; i++) { ctx[i]->batch_ev = c->lead_ev[r]; ctx[i]->batch_pending = true; ctx[i]->batch_leader = c; }
    return VP8B200_OK;
}

/* ---- statistics and per-kernel profiling -------------------------------------------------- */

extern "C" void vp8b200_global_stats(uint64_t out[4])
{
    out[0] = g_h2d_bytes.load(); out[1] = g_d2h_bytes.load();
    out[2] = g_launches.load(); out[3] = g_frames.load();
}

extern "C" int vp8b200_profile_enable(vp8b200_ctx *c, int enable)
{
    if (!c) return VP8B200_ERR_INVALID;
    if (!c->spans) c->spans = new (std::nothrow) std::vector<ProfSpan>();
    if (!c->spans) return VP8B200_ERR_NOMEM;
    c->profiling = enable != 0;
    return VP8B200_OK;
}

extern "C" int vp8b200_profile_read(vp8b200_ctx *c, double ms[4], uint64_t count[4])
{
    if (!c || !ms || !count) return VP8B200_ERR_INVALID;
    for (int i = 0; i < 4; i++) { ms[i] = 0; count[i] = 0; }
    if (!c->spans) return VP8B200_OK;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->stream));
    for (auto &sp : *c->spans) {
        float t = 0;
        CK(c, cudaEventElapsedTime(&t, sp.a, sp.b));
        ms[sp.kind] += t; count[sp.kind]++;
        cudaEventDestroy(sp.a); cudaEventDestroy(sp.b);
    }
    c->spans->clear();
    return VP8B200_OK;
}
