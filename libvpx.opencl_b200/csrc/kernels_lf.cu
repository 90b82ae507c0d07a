/* kernels_lf.cu - in-loop deblocking filter, normal and simple variants.
 *
 * Restates vp8_loop_filter_frame (vp8/common/loopfilter.c:203-316) with the edge filters of
 * vp8/common/loopfilter_filters.c, level selection of vp8_loop_filter_frame_init
 * (loopfilter.c:117-201) and the limit tables of vp8_loop_filter_update_sharpness (:66-96).
 *
 * Schedule: the reference filters macroblocks in raster order and each macroblock reads
 * pixels its left, above and above-right neighbours have already modified.  One warp owns
 * one macroblock ROW and walks it left to right; row r can do the horizontal edges of column
 * c once row r-1 has done the left edge of column c+1.  There are no flags and no fences on
 * that path: a row hands the bottom 4 pixel rows of each finished macroblock DOWN as a
 * message and the row below finishes (top-edge filter) and stores the 3 rows it modifies.
 *   - rows in the same CTA: message through a shared-memory ring (LF_RING slots per row);
 *   - across CTAs: tagged 64-bit words in global memory, {32 bits of pixels, 32-bit frame
 *     tag}; an aligned 64-bit access is single-copy atomic, so a word whose tag matches
 *     carries valid pixels and the consumer simply polls the words (NCCL's LL idea).
 *
 * Inside a macroblock the warp first filters the vertical edges with lane = pixel row
 * (lanes 0-15 luma rows, 16-23 U rows, 24-31 V rows; rows live in registers, the 4 pixels
 * left of the MB are carried over from the previous column), transposes through a 512-byte
 * shared-memory tile, filters the horizontal edges with lane = pixel column, and transposes
 * back.  The last 4 columns of a macroblock are stored one iteration later, after the next
 * macroblock's left-edge filter has modified them.  Filters are branch-free (select on the
 * mask) so that a lane's instruction stream is short and free of divergence; macroblocks
 * without inner edges (skip_lf) only move the 8 rows the top edge needs through the tile.
 */
#include "vp8b200_dev.cuh"

#ifndef LF_ROWS_PER_CTA
#define LF_ROWS_PER_CTA 4
#endif
#ifndef LF_RING
#define LF_RING 4                 /* shared-memory message slots per row (power of two) */
#endif
static_assert(1 + (LF_ROWS_PER_CTA - 1) * LF_RING <= 16, "one named barrier per (row pair, ring slot)");
#ifndef LF_PF
#define LF_PF 2                   /* prefetch distance in macroblocks (1: 0.80 ms, 2: 0.72, 3: 0.77, 5: 0.88) */
#endif

__device__ __forceinline__ int sc(int v) { return max(min(v, 127), -128); }
__device__ __forceinline__ int ad(int a, int b) { return __sad(a, b, 0); }      /* |a-b|, one VABSDIFF */
__device__ __forceinline__ int c255(int v) { return __vimin_s32_relu(v, 255); }  /* clamp to 0..255 */

struct LfParams { int ilim, blim, mblim, thr; };

/* loopfilter_filters.c:27-49; pixels as plain 0..255 ints */
__device__ __forceinline__ bool lf_mask(int p3, int p2, int p1, int p0, int q0, int q1, int q2, int q3,
                                        int ilim, int elim)
{
    int m = max(__vimax3_s32(ad(p3, p2), ad(p2, p1), ad(p1, p0)), __vimax3_s32(ad(q1, q0), ad(q2, q1), ad(q3, q2)));
    return m <= ilim && ad(p0, q0) * 2 + (ad(p1, q1) >> 1) <= elim;
}

/* inner edge: loopfilter_filters.c:51-97.  sc(qs0 - F) + 128 == clamp(q0 - F, 0, 255), so the
 * signed-char arithmetic of the reference is done directly on pixel values. */
__device__ __forceinline__ void lf_inner(int p3, int p2, int &p1, int &p0, int &q0, int &q1, int q2, int q3,
                                         const LfParams &P)
{
    const bool mask = lf_mask(p3, p2, p1, p0, q0, q1, q2, q3, P.ilim, P.blim);
    const bool hev = max(ad(p1, p0), ad(q1, q0)) > P.thr;
    int f = hev ? sc(p1 - q1) : 0;
    f = sc(f + 3 * (q0 - p0));
    f = mask ? f : 0;
    const int f1 = min(f + 4, 127) >> 3, f2 = min(f + 3, 127) >> 3;
    const int u = hev ? 0 : (f1 + 1) >> 1;
    q0 = c255(q0 - f1);
    p0 = c255(p0 + f2);
    q1 = c255(q1 - u);
    p1 = c255(p1 + u);
}

/* macroblock edge: loopfilter_filters.c:161-214 */
__device__ __forceinline__ void lf_mbedge(int p3, int &p2, int &p1, int &p0, int &q0, int &q1, int &q2, int q3,
                                          const LfParams &P)
{
    const bool mask = lf_mask(p3, p2, p1, p0, q0, q1, q2, q3, P.ilim, P.mblim);
    const bool hev = max(ad(p1, p0), ad(q1, q0)) > P.thr;
    int f = sc(sc(p1 - q1) + 3 * (q0 - p0));
    f = mask ? f : 0;
    const int g = hev ? f : 0, w = hev ? 0 : f;
    const int f1 = min(g + 4, 127) >> 3, f2 = min(g + 3, 127) >> 3;
    /* |(63 + w*k) >> 7| <= 27: the reference's clamp of u is a no-op */
    const int u27 = (63 + w * 27) >> 7, u18 = (63 + w * 18) >> 7, u9 = (63 + w * 9) >> 7;
    q0 = c255(c255(q0 - f1) - u27);
    p0 = c255(c255(p0 + f2) + u27);
    q1 = c255(q1 - u18);
    p1 = c255(p1 + u18);
    q2 = c255(q2 - u9);
    p2 = c255(p2 + u9);
}

/* simple filter: loopfilter_filters.c:281-315 */
__device__ __forceinline__ void lf_simple(int p1, int &p0, int &q0, int q1, int blim)
{
    const bool mask = ad(p0, q0) * 2 + (ad(p1, q1) >> 1) <= blim;
    int f = sc(sc(p1 - q1) + 3 * (q0 - p0));
    f = mask ? f : 0;
    q0 = c255(q0 - (min(f + 4, 127) >> 3));
    p0 = c255(p0 + (min(f + 3, 127) >> 3));
}

__device__ __forceinline__ void unpack(unsigned w, int &a, int &b, int &c, int &d)
{
    a = __byte_perm(w, 0, 0x4440); b = __byte_perm(w, 0, 0x4441);
    c = __byte_perm(w, 0, 0x4442); d = __byte_perm(w, 0, 0x4443);
}
__device__ __forceinline__ unsigned pack(int a, int b, int c, int d)
{
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

template <bool MB>
__device__ __forceinline__ void edge8(int &p3, int &p2, int &p1, int &p0, int &q0, int &q1, int &q2, int &q3,
                                      bool simple, const LfParams &P)
{
    if (simple) lf_simple(p1, p0, q0, q1, MB ? P.mblim : P.blim);
    else if (MB) lf_mbedge(p3, p2, p1, p0, q0, q1, q2, q3, P);
    else lf_inner(p3, p2, p1, p0, q0, q1, q2, q3, P);
}

/* ---- in-CTA hand-off: named barriers ---------------------------------------------------
 * Producer row w and consumer row w+1 meet on barrier 1 + w*LF_RING + (col % LF_RING): the
 * producer stores the ring slot and ARRIVES (does not wait), the consumer SYNCs.  A waiting
 * consumer is suspended by the hardware and issues nothing, where a shared-memory spin took
 * half of the kernel's issue slots (profiles/r01_summary_v8.md).  A barrier id is reused
 * every LF_RING columns; the ring-full check keeps the producer from arriving at a barrier
 * whose previous phase the consumer has not left. */
#ifndef LF_BAR
#define LF_BAR 1
#endif
__device__ __forceinline__ void bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void bar_wait(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

/* ---- global hand-off (across CTAs) ------------------------------------------------------ */
__device__ __forceinline__ void st_msg2(uint8_t *p, unsigned a, unsigned b, unsigned tag)
{
    unsigned long long x = ((unsigned long long)tag << 32) | a, y = ((unsigned long long)tag << 32) | b;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(x), "l"(y) : "memory");
}
__device__ __forceinline__ void ld_msg2(const uint8_t *p, unsigned long long &x, unsigned long long &y)
{
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "l"(p) : "memory");
}
/* slot: 256 B per MB = luma rows 12..15 (4 x 32 B), U rows 4..7 (4 x 16 B), V rows 4..7 */
__device__ __forceinline__ void g_send(uint8_t *slot, const unsigned (&m)[4], unsigned tag, bool luma)
{
    if (luma) { st_msg2(slot, m[0], m[1], tag); st_msg2(slot + 16, m[2], m[3], tag); }
    else st_msg2(slot, m[0], m[1], tag);
}
__device__ __forceinline__ void g_recv(const uint8_t *slot, unsigned (&m)[4], unsigned tag, bool luma)
{
    unsigned long long a, b, c = 0, d = 0;
    int tries = 0;
    for (;;) {
        ld_msg2(slot, a, b);
        if (luma) ld_msg2(slot + 16, c, d);
        bool ok = (unsigned)(a >> 32) == tag && (unsigned)(b >> 32) == tag;
        if (luma) ok = ok && (unsigned)(c >> 32) == tag && (unsigned)(d >> 32) == tag;
        if (ok) break;
        if (++tries > 8) __nanosleep(100);
    }
    m[0] = (unsigned)a; m[1] = (unsigned)b; m[2] = (unsigned)c; m[3] = (unsigned)d;
}

__global__ void __launch_bounds__(LF_ROWS_PER_CTA * 32, 32 / LF_ROWS_PER_CTA)
k_loopfilter(const FrameJob *__restrict__ jobs, const int n_jobs, const Geo g,
             unsigned *ticket, const unsigned ticket_base)
{
    __shared__ FrameJob job;
    __shared__ unsigned s_ticket;
    __shared__ uint8_t s_lvl[64];                          /* [seg][ref][mode class] */
    __shared__ __align__(16) uint8_t s_tile[LF_ROWS_PER_CTA][512];
    /* message ring of row w -> row w+1: 128 B = luma rows 12..15 (4x16), U 4..7 (4x8), V 4..7 */
    __shared__ __align__(16) uint8_t s_ring[LF_ROWS_PER_CTA][LF_RING][128];
    __shared__ volatile unsigned s_sent[LF_ROWS_PER_CTA], s_rcvd[LF_ROWS_PER_CTA];
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u) - ticket_base;
    if (threadIdx.x < LF_ROWS_PER_CTA) { s_sent[threadIdx.x] = 0; s_rcvd[threadIdx.x] = 0; }
    __syncthreads();
    const unsigned t = s_ticket;
    const int ji = t % n_jobs, group = t / n_jobs;
    {
        const unsigned *s = reinterpret_cast<const unsigned *>(&jobs[ji]);
        unsigned *d = reinterpret_cast<unsigned *>(&job);
        for (int i = threadIdx.x; i < (int)(sizeof(FrameJob) / 4); i += blockDim.x) d[i] = s[i];
    }
    __syncthreads();
    const vp8b200_frame_hdr &h = job.hdr;
    if (h.filter_level == 0) return;                       /* onyxd_if.c:576 */
    if (threadIdx.x < 64) {
        /* vp8_loop_filter_frame_init, loopfilter.c:117-201 */
        const int seg = threadIdx.x >> 4, ref = (threadIdx.x >> 2) & 3, mode = threadIdx.x & 3;
        int lvl = h.filter_level;
        if (h.segmentation_enabled) {
            if (h.segment_abs_delta) lvl = h.segment_lf[seg];
            else lvl = min(max(lvl + h.segment_lf[seg], 0), 63);
        }
        if (h.mode_ref_lf_delta_enabled) {
            lvl += h.ref_lf_deltas[ref];
            if (ref == 0) { if (mode == 0) lvl += h.mode_lf_deltas[0]; }   /* B_PRED only */
            else lvl += h.mode_lf_deltas[mode];
            lvl = min(max(lvl, 0), 63);
        }
        s_lvl[threadIdx.x] = (uint8_t)lvl;
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb_row = group * LF_ROWS_PER_CTA + warp;
    if (mb_row >= g.mb_rows) return;
    const unsigned tag = job.epoch_lf;                     /* marks this frame's global messages */
    const bool simple = h.filter_type != 0;
    const bool key = h.frame_type == 0;
    const int sharp = h.sharpness_level;

    /* lane geometry */
    const bool luma = lane < 16;
    const int pi = luma ? lane : (lane & 7);               /* row (V phase) / column (H phase) */
    const int stride = luma ? g.y_stride : g.uv_stride;
    const int mbw = luma ? 16 : 8;                         /* MB width = height in this plane */
    uint8_t *plane = job.dst + (luma ? g.y_off : (lane < 24 ? g.u_off : g.v_off));
    uint8_t *rowp = plane + (size_t)(mb_row * mbw + pi) * stride;      /* my pixel row, x = 0 */
    const bool lane_on = luma || !simple;                  /* simple filter: luma only */
    uint8_t *tile = s_tile[warp] + (luma ? 0 : (lane < 24 ? 320 : 416));
    const bool top = mb_row > 0;
    const bool last_row = mb_row == g.mb_rows - 1;
    /* rows >= keep of every MB (not in the last MB row) are finished and stored by the row
     * below; this row hands rows keep.. down as a message instead */
    const int keep = luma ? 12 : 4;
    const bool owns_store = lane_on && (last_row || pi <= keep);
    const bool sender = lane_on && !last_row && pi >= keep;
    const bool receiver = lane_on && top && pi < 4;
    const bool send_smem = warp < LF_ROWS_PER_CTA - 1;     /* consumer row lives in this CTA */
    const bool recv_smem = warp > 0;
    /* global slot offsets */
    const int gs_off = luma ? (pi - 12) * 32 : (lane < 24 ? 128 : 192) + (pi - 4) * 16;
    const int gr_off = luma ? pi * 32 : (lane < 24 ? 128 : 192) + pi * 16;
    uint8_t *gmsg_out = job.lf_msg + (size_t)mb_row * g.mb_cols * 256 + gs_off;
    const uint8_t *gmsg_in = job.lf_msg + (size_t)(mb_row - 1) * g.mb_cols * 256 + gr_off;
    /* shared ring offsets */
    const int ss_off = luma ? (pi - 12) * 16 : (lane < 24 ? 64 : 96) + (pi - 4) * 8;
    const int sr_off = luma ? pi * 16 : (lane < 24 ? 64 : 96) + pi * 8;

    const unsigned *mbrec = reinterpret_cast<const unsigned *>(job.mb + (size_t)mb_row * g.mb_cols);
    /* pixel rows are prefetched LF_PF macroblocks ahead: under load the DRAM/L2 latency of a
     * row is several iterations long */
    unsigned cur[4] = {0, 0, 0, 0}, pf[LF_PF][4], prev[3] = {0, 0, 0}, halo = 0;
    auto load_row = [&](int col, unsigned (&d)[4]) {
        if (lane_on && col < g.mb_cols) {
            if (luma) { uint4 v = *reinterpret_cast<const uint4 *>(rowp + col * 16); d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
            else { uint2 v = *reinterpret_cast<const uint2 *>(rowp + col * 8); d[0] = v.x; d[1] = v.y; }
        }
    };
#pragma unroll
    for (int i = 0; i < LF_PF; i++) { pf[i][0] = pf[i][1] = pf[i][2] = pf[i][3] = 0; }
    load_row(0, cur);
#pragma unroll
    for (int i = 0; i < LF_PF - 1; i++) load_row(i + 1, pf[i]);
    unsigned rec = mbrec[0], rec_n = 0;

    /* message for MB `col` of this row: words of rows keep.. after the next MB's left edge */
    auto send = [&](int col) {
        unsigned m[4];
        if (luma) { m[0] = prev[0]; m[1] = prev[1]; m[2] = prev[2]; m[3] = halo; }
        else { m[0] = prev[0]; m[1] = halo; m[2] = 0; m[3] = 0; }
        if (send_smem) {
            while ((int)(col - s_rcvd[warp]) >= LF_RING) { }             /* ring full: wait for the consumer */
            if (sender) {
                uint8_t *slot = s_ring[warp][col & (LF_RING - 1)] + ss_off;
                if (luma) *reinterpret_cast<uint4 *>(slot) = make_uint4(m[0], m[1], m[2], m[3]);
                else *reinterpret_cast<uint2 *>(slot) = make_uint2(m[0], m[1]);
            }
#if LF_BAR
            bar_arrive(1 + warp * LF_RING + (col & (LF_RING - 1)));
#else
            __threadfence_block();
            __syncwarp();
            if (lane == 0) s_sent[warp] = (unsigned)col + 1;
#endif
        } else if (sender) {
            g_send(gmsg_out + (size_t)col * 256, m, tag, luma);
        }
    };

    for (int c = 0; c < g.mb_cols; c++) {
        /* prefetch: the record of the next macroblock, the rows of the one LF_PF ahead */
        if (c + 1 < g.mb_cols) rec_n = mbrec[(c + 1) * 4];
        load_row(c + LF_PF, pf[LF_PF - 1]);
        /* per-MB decisions, loopfilter.c:245-253 */
        const int y_mode = rec & 255, ref = (rec >> 16) & 255, flags = rec >> 24;
        const bool skip_lf = y_mode != VP8B200_B_PRED && y_mode != VP8B200_SPLITMV && (flags & VP8B200_MBF_SKIP);
        /* mode_lf_lut, loopfilter.c:52-63: DC,V,H,TM,ZEROMV -> 1 ; B_PRED -> 0 ; NEAREST,NEAR,NEW -> 2 ; SPLIT -> 3 */
        const int mclass = y_mode == VP8B200_B_PRED ? 0 : y_mode == VP8B200_SPLITMV ? 3
                         : (y_mode <= VP8B200_TM_PRED || y_mode == VP8B200_ZEROMV) ? 1 : 2;
        const int level = s_lvl[((flags & 3) << 4) | (ref << 2) | mclass];
        uint8_t *colp = rowp + c * mbw;                     /* my row at this MB's x = 0 */
        LfParams P;
        {   /* loopfilter.c:66-96 and :28-50 */
            int il = level >> (sharp > 0);
            il >>= (sharp > 4);
            if (sharp > 0) il = min(il, 9 - sharp);
            il = max(il, 1);
            P.ilim = il; P.blim = 2 * level + il; P.mblim = 2 * (level + 2) + il;
            P.thr = key ? (level >= 40 ? 2 : (level >= 15 ? 1 : 0))
                        : (level >= 40 ? 3 : (level >= 20 ? 2 : (level >= 15 ? 1 : 0)));
        }
        /* ---- vertical edges, lane = pixel row, pixels unpacked once ---- */
        if (lane_on && level) {
            int x[8];
            if (c > 0) {
                int h0, h1, h2, h3;
                unpack(halo, h0, h1, h2, h3);
                unpack(cur[0], x[0], x[1], x[2], x[3]);
                edge8<true>(h0, h1, h2, h3, x[0], x[1], x[2], x[3], simple, P);
                halo = pack(h0, h1, h2, h3);
                if (skip_lf) cur[0] = pack(x[0], x[1], x[2], x[3]);
            } else if (!skip_lf) {
                unpack(cur[0], x[0], x[1], x[2], x[3]);
            }
            if (!skip_lf) {
                unpack(cur[1], x[4], x[5], x[6], x[7]);
                edge8<false>(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7], simple, P);
                cur[0] = pack(x[0], x[1], x[2], x[3]);
                if (luma) {
                    int y[8];
                    unpack(cur[2], y[0], y[1], y[2], y[3]);
                    edge8<false>(x[4], x[5], x[6], x[7], y[0], y[1], y[2], y[3], simple, P);
                    cur[1] = pack(x[4], x[5], x[6], x[7]);
                    unpack(cur[3], y[4], y[5], y[6], y[7]);
                    edge8<false>(y[0], y[1], y[2], y[3], y[4], y[5], y[6], y[7], simple, P);
                    cur[2] = pack(y[0], y[1], y[2], y[3]);
                    cur[3] = pack(y[4], y[5], y[6], y[7]);
                } else {
                    cur[1] = pack(x[4], x[5], x[6], x[7]);
                }
            }
        }
        /* the previous MB of this row is now final: store / hand down its last 4 columns */
        if (c > 0) {
            if (owns_store) *reinterpret_cast<unsigned *>(colp - 4) = halo;
            if (!last_row) send(c - 1);
        }
        /* ---- the 4 rows above arrive as a message from the row above ---- */
        if (top) {
            if (recv_smem) {
#if LF_BAR
                bar_wait(1 + (warp - 1) * LF_RING + (c & (LF_RING - 1)));
#else
                for (int tries = 0; s_sent[warp - 1] <= (unsigned)c; tries++) if (tries > 24) __nanosleep(64);   /* spin briefly, then back off */
                __threadfence_block();
#endif
                if (receiver) {
                    const uint8_t *slot = s_ring[warp - 1][c & (LF_RING - 1)] + sr_off;
                    if (luma) *reinterpret_cast<uint4 *>(tile + pi * 16) = *reinterpret_cast<const uint4 *>(slot);
                    else *reinterpret_cast<uint2 *>(tile + pi * 8) = *reinterpret_cast<const uint2 *>(slot);
                }
                __syncwarp();
                if (lane == 0) s_rcvd[warp - 1] = (unsigned)c + 1;
            } else if (receiver) {
                unsigned m[4];
                g_recv(gmsg_in + (size_t)c * 256, m, tag, luma);
                if (luma) *reinterpret_cast<uint4 *>(tile + pi * 16) = make_uint4(m[0], m[1], m[2], m[3]);
                else *reinterpret_cast<uint2 *>(tile + pi * 8) = make_uint2(m[0], m[1]);
            }
        }
        if (level) {
            /* rows the horizontal edges touch: all of them, or only rows 0..3 for the top edge
             * of a macroblock without inner edges (nothing at all if that has no top either) */
            const int nrows = skip_lf ? 4 : mbw;            /* MB rows entering the tile */
            if (lane_on && (!skip_lf || top) && pi < nrows) {
                if (luma) *reinterpret_cast<uint4 *>(tile + (pi + 4) * 16) = make_uint4(cur[0], cur[1], cur[2], cur[3]);
                else *reinterpret_cast<uint2 *>(tile + (pi + 4) * 8) = make_uint2(cur[0], cur[1]);
            }
            __syncwarp();
            /* ---- horizontal edges, lane = pixel column ---- */
            if (lane_on && (!skip_lf || top)) {
                int v[8];
                if (top) {
#pragma unroll
                    for (int r = 0; r < 8; r++) v[r] = tile[r * mbw + pi];
                    edge8<true>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], simple, P);
#pragma unroll
                    for (int r = 1; r < 4; r++) tile[r * mbw + pi] = (uint8_t)v[r];
                    if (skip_lf) {
#pragma unroll
                        for (int r = 4; r < 7; r++) tile[r * mbw + pi] = (uint8_t)v[r];
                    }
                } else {
#pragma unroll
                    for (int r = 4; r < 8; r++) v[r] = tile[r * mbw + pi];
                }
                if (!skip_lf) {
                    int w[8];
#pragma unroll
                    for (int r = 0; r < 4; r++) w[r] = tile[(r + 8) * mbw + pi];
                    edge8<false>(v[4], v[5], v[6], v[7], w[0], w[1], w[2], w[3], simple, P);
#pragma unroll
                    for (int r = 4; r < 8; r++) tile[r * mbw + pi] = (uint8_t)v[r];
                    if (luma) {
#pragma unroll
                        for (int r = 4; r < 8; r++) w[r] = tile[(r + 8) * mbw + pi];
                        edge8<false>(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], simple, P);
#pragma unroll
                        for (int r = 0; r < 4; r++) tile[(r + 8) * mbw + pi] = (uint8_t)w[r];
#pragma unroll
                        for (int r = 0; r < 4; r++) v[r] = tile[(r + 16) * mbw + pi];
                        edge8<false>(w[4], w[5], w[6], w[7], v[0], v[1], v[2], v[3], simple, P);
#pragma unroll
                        for (int r = 4; r < 8; r++) tile[(r + 8) * mbw + pi] = (uint8_t)w[r];
#pragma unroll
                        for (int r = 0; r < 4; r++) tile[(r + 16) * mbw + pi] = (uint8_t)v[r];
                    } else {
#pragma unroll
                        for (int r = 0; r < 4; r++) tile[(r + 8) * mbw + pi] = (uint8_t)w[r];
                    }
                }
            }
            __syncwarp();
            if (lane_on && (!skip_lf || top) && pi < nrows) {
                if (luma) { uint4 v = *reinterpret_cast<const uint4 *>(tile + (pi + 4) * 16); cur[0] = v.x; cur[1] = v.y; cur[2] = v.z; cur[3] = v.w; }
                else { uint2 v = *reinterpret_cast<const uint2 *>(tile + (pi + 4) * 8); cur[0] = v.x; cur[1] = v.y; }
            }
        } else {
            __syncwarp();                                   /* message rows visible in the tile */
        }
        /* ---- rows out; the last word waits for the next MB's left edge ---- */
        if (owns_store) {
            *reinterpret_cast<unsigned *>(colp) = cur[0];
            if (luma) { *reinterpret_cast<unsigned *>(colp + 4) = cur[1]; *reinterpret_cast<unsigned *>(colp + 8) = cur[2]; }
        }
        if (receiver && pi >= 1) {                          /* rows -3..-1: this row finishes them */
            uint8_t *ap = plane + (size_t)(mb_row * mbw - 4 + pi) * stride + c * mbw;
            if (luma) *reinterpret_cast<uint4 *>(ap) = *reinterpret_cast<const uint4 *>(tile + pi * 16);
            else *reinterpret_cast<uint2 *>(ap) = *reinterpret_cast<const uint2 *>(tile + pi * 8);
        }
        if (luma) { prev[0] = cur[0]; prev[1] = cur[1]; prev[2] = cur[2]; halo = cur[3]; }
        else { prev[0] = cur[0]; halo = cur[1]; }
        __syncwarp();                                       /* tile is reused by the next MB */
        rec = rec_n;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            cur[i] = pf[0][i];
#pragma unroll
            for (int k = 0; k + 1 < LF_PF; k++) pf[k][i] = pf[k + 1][i];
        }
    }
    /* last 4 columns of the row */
    if (owns_store) *reinterpret_cast<unsigned *>(rowp + g.mb_cols * mbw - 4) = halo;
    if (!last_row) send(g.mb_cols - 1);
}

void vp8b200_launch_loopfilter(cudaStream_t s, const FrameJob *jobs, int n_jobs, const Geo &g,
                               unsigned int *ticket, unsigned int ticket_base, int *n_ctas)
{
    int groups = (g.mb_rows + LF_ROWS_PER_CTA - 1) / LF_ROWS_PER_CTA;
    *n_ctas = groups * n_jobs;
    k_loopfilter<<<groups * n_jobs, LF_ROWS_PER_CTA * 32, 0, s>>>(jobs, n_jobs, g, ticket, ticket_base);
}
