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
 *   - rows in the same CTA: message through a shared-memory ring (LF_RING slots per row),
 *     producer and consumer meeting on a named barrier per ring slot (bar.arrive / bar.sync:
 *     a waiting row is suspended by the hardware and issues nothing);
 *   - across CTAs: tagged 64-bit words in global memory, {32 bits of pixels, 32-bit frame
 *     tag}; an aligned 64-bit access is single-copy atomic, so a word whose tag matches
 *     carries valid pixels and the consumer simply polls the words (NCCL's LL idea).
 *
 * Inside a macroblock the warp first filters the vertical edges with lane = pixel row
 * (lanes 0-15 luma rows, 16-23 U rows, 24-31 V rows; rows live in registers, the 4 pixels
 * left of the MB are carried over from the previous column), transposes through a 512-byte
 * shared-memory tile, filters the horizontal edges with lane = pixel column, and transposes
 * back.  A macroblock is stored one iteration later, after the next macroblock's left-edge
 * filter has modified its last 3 columns: one 16-byte (luma) / 8-byte (chroma) store per pixel
 * row.  Filters are branch-free (select on the mask) and chroma lanes run the two luma-only
 * inner edges on scratch data, so the normal filter has no divergent branch; macroblocks
 * without inner edges (skip_lf) only move the 8 rows the top edge needs through the tile.
 *
 * Memory: pixel rows arrive through a lane-private cp.async ring LF_PF macroblocks deep (the
 * row latency under a 64-stream load is several iterations), record words are read 32
 * macroblocks at a time one batch ahead and turned into filter limits off the chain
 * (measurements: profiles/r01_summary_v8.md).
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
#define LF_PF 16                  /* cp.async prefetch distance in macroblocks (power of two, >= 2; per 64x1080p launch: 2: 1.1 ms, 4: 0.77, 8: 0.49, 16: 0.465; 32 would cost a resident CTA) */
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

template <bool MB, bool SIMPLE>
__device__ __forceinline__ void edge8(int &p3, int &p2, int &p1, int &p0, int &q0, int &q1, int &q2, int &q3,
                                      const LfParams &P)
{
    if (SIMPLE) lf_simple(p1, p0, q0, q1, MB ? P.mblim : P.blim);
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
#ifndef LF_MIN_CTAS
#define LF_MIN_CTAS 4
#endif
#ifndef LF_UNIFORM_POLL
#define LF_UNIFORM_POLL 1
#endif
#ifndef LF_POLL_SLEEP
#define LF_POLL_SLEEP 100         /* ns between polls of a global message after 8 immediate tries */
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
/* w holds an earlier read of the slot (the loop prefetches the next macroblock's message while
 * it works on the current one, so a row that runs a little behind the row above never waits
 * for L2 here); re-read until every word carries this frame's tag */
__device__ __forceinline__ void g_load(const uint8_t *slot, unsigned long long (&w)[4], bool luma)
{
    ld_msg2(slot, w[0], w[1]);
    if (luma) ld_msg2(slot + 16, w[2], w[3]);
}
__device__ __forceinline__ void g_recv(const uint8_t *slot, unsigned long long (&w)[4], unsigned (&m)[4], unsigned tag, bool luma,
                                       bool receiver)
{
#if LF_UNIFORM_POLL
    /* the WARP leaves the loop together: lanes that break out one by one leave it diverged */
    int tries = 0;
    for (;;) {
        bool ok = !receiver || ((unsigned)(w[0] >> 32) == tag && (unsigned)(w[1] >> 32) == tag);
        if (luma && receiver) ok = ok && (unsigned)(w[2] >> 32) == tag && (unsigned)(w[3] >> 32) == tag;
        if (__all_sync(0xffffffffu, ok)) break;
        if (++tries > 8) __nanosleep(LF_POLL_SLEEP);
        if (!ok) g_load(slot, w, luma);
    }
#else
    if (receiver) {
        int tries = 0;
        for (;;) {
            bool ok = (unsigned)(w[0] >> 32) == tag && (unsigned)(w[1] >> 32) == tag;
            if (luma) ok = ok && (unsigned)(w[2] >> 32) == tag && (unsigned)(w[3] >> 32) == tag;
            if (ok) break;
            if (++tries > 8) __nanosleep(LF_POLL_SLEEP);
            g_load(slot, w, luma);
        }
    }
#endif
    m[0] = (unsigned)w[0]; m[1] = (unsigned)w[1]; m[2] = (unsigned)w[2]; m[3] = (unsigned)w[3];
}

/* mode_lf_lut, loopfilter.c:52-63, two bits per y_mode: DC,V,H,TM,ZEROMV -> 1 ; B_PRED -> 0 ;
 * NEARESTMV,NEARMV,NEWMV -> 2 ; SPLITMV -> 3 */
#define LF_MODE_CLASS_LUT (1u | 1u << 2 | 1u << 4 | 1u << 6 | 0u << 8 | 2u << 10 | 2u << 12 | 1u << 14 | 2u << 16 | 3u << 18)

/* one macroblock row, left to right (see the header comment) */
template <bool SIMPLE>
__device__ __forceinline__ void lf_row(const FrameJob &job, const Geo &g, const int mb_row, const int warp, const int lane,
                                       const unsigned *s_par, uint8_t *s_tile_w, uint8_t *s_scratch_w, uint8_t *s_pf_w,
                                       uint8_t (*s_ring)[LF_RING][192], volatile unsigned *s_rcvd)
{
    const unsigned tag = job.epoch_lf;                     /* marks this frame's global messages */

    /* lane geometry */
    const bool luma = lane < 16;
    const int pi = luma ? lane : (lane & 7);               /* row (V phase) / column (H phase) */
    const int stride = luma ? g.y_stride : g.uv_stride;
    const int mbw = luma ? 16 : 8;                         /* MB width = height in this plane */
    uint8_t *plane = job.dst + (luma ? g.y_off : (lane < 24 ? g.u_off : g.v_off));
    uint8_t *rowp = plane + (size_t)(mb_row * mbw + pi) * stride;      /* my pixel row, x = 0 */
    const bool lane_on = luma || !SIMPLE;                  /* simple filter: luma only */
    /* tile: 16 bytes per pixel row for every plane (a chroma row uses the first 8), so that
     * all lanes move rows with the same 16-byte operations: luma rows -4..15 at 0, U rows
     * -4..7 at 320, V at 512.  Tile rows 12.. exist for luma only; chroma lanes run the same
     * (branch-free) code on a scratch area instead of diverging */
    uint8_t *tile = s_tile_w + (luma ? 0 : (lane < 24 ? 320 : 512));
    uint8_t *tile_hi = luma ? tile : s_scratch_w + (lane < 24 ? 0 : 320);
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
    const int ss_off = luma ? (pi - 12) * 16 : (lane < 24 ? 64 : 128) + (pi - 4) * 16;
    const int sr_off = luma ? pi * 16 : (lane < 24 ? 64 : 128) + pi * 16;
    const bool own16 = owns_store && luma, own8 = owns_store && !luma;

    const unsigned *mbrec = reinterpret_cast<const unsigned *>(job.mb + (size_t)mb_row * g.mb_cols);
    /* Per-MB decisions (loopfilter.c:245-253), 32 macroblocks at a time: lane l loads the first
     * record word of macroblock batch + l one batch ahead and turns it into the parameter word
     * (level -> limits through s_par, bit 31 = no inner edges); the loop broadcasts one word per
     * macroblock by shuffle, so neither the record load nor the level arithmetic sits on the
     * per-macroblock dependency chain. */
    auto par_of = [&](unsigned rec) -> unsigned {
        const int y_mode = rec & 255, ref = (rec >> 16) & 255, flags = rec >> 24;
        const int mclass = (LF_MODE_CLASS_LUT >> (2 * y_mode)) & 3;
        const bool skip = mclass != 0 && mclass != 3 && (flags & VP8B200_MBF_SKIP);
        return s_par[((flags & 3) << 4) | (ref << 2) | mclass] | (skip ? 0x80000000u : 0u);
    };
    auto load_rec = [&](int col) -> unsigned { return col < g.mb_cols ? mbrec[col * 4] : 0u; };
    unsigned par_lane = par_of(load_rec(lane));             /* macroblocks 0..31 */
    unsigned rec_next = load_rec(32 + lane), par_next = 0;  /* macroblocks 32..63 */

    /* Pixel rows are prefetched LF_PF macroblocks ahead with cp.async into a per-lane
     * shared-memory ring (lane-private slots: no barrier, only cp.async.wait_group): under
     * load the L2 / DRAM latency of a row is several iterations long, and a register
     * prefetch is consumed (moved) one iteration after it was issued. */
    const unsigned pf_base = (unsigned)__cvta_generic_to_shared(s_pf_w) + lane * 16;
    auto prefetch = [&](int col) {
        const unsigned dst = pf_base + (col & (LF_PF - 1)) * 512;
        const int in = col < g.mb_cols;
        asm volatile("{ .reg .pred p, q;\n\t"
                     "setp.ne.b32 p, %2, 0;\n\t"
                     "setp.ne.b32 q, %3, 0;\n\t"
                     "@p cp.async.ca.shared.global [%0], [%1], 16;\n\t"
                     "@q cp.async.ca.shared.global [%0], [%1], 8;\n\t"
                     "cp.async.commit_group; }"
                     ::"r"(dst), "l"(rowp + col * mbw), "r"((int)(luma && in)), "r"((int)(lane_on && !luma && in)) : "memory");
    };
    auto fetch = [&](int col, unsigned (&d)[4]) {
        const uint4 v = *reinterpret_cast<const uint4 *>(s_pf_w + (col & (LF_PF - 1)) * 512 + lane * 16);
        d[0] = v.x; d[1] = v.y; d[2] = luma ? v.z : 0u; d[3] = luma ? v.w : 0u;
    };
    unsigned cur[4], nxt[4], prev[3] = {0, 0, 0}, halo = 0;
    const unsigned long long no_msg = (unsigned long long)~tag << 32;     /* a word that is not this frame's */
    unsigned long long gw[4] = {no_msg, no_msg, no_msg, no_msg};          /* last read of the next global message */
#pragma unroll
    for (int i = 0; i < LF_PF; i++) prefetch(i);
    asm volatile("cp.async.wait_group %0;" ::"n"(LF_PF - 1) : "memory");
    fetch(0, cur);

    /* rows of the macroblock LEFT of x = colp, final once the left edge at colp is filtered */
    auto store_prev = [&](uint8_t *colp) {
        if (own16) *reinterpret_cast<uint4 *>(colp - 16) = make_uint4(prev[0], prev[1], prev[2], halo);
        if (own8) *reinterpret_cast<uint2 *>(colp - 8) = make_uint2(prev[0], halo);
    };
    /* message for MB `col` of this row: words of rows keep.. after the next MB's left edge */
    auto send = [&](int col) {
        unsigned m[4];
        if (luma) { m[0] = prev[0]; m[1] = prev[1]; m[2] = prev[2]; m[3] = halo; }
        else { m[0] = prev[0]; m[1] = halo; m[2] = 0; m[3] = 0; }
        if (send_smem) {
            while ((int)(col - s_rcvd[warp]) >= LF_RING) { }             /* ring full: wait for the consumer */
            if (sender) {
                uint8_t *slot = s_ring[warp][col & (LF_RING - 1)] + ss_off;
                *reinterpret_cast<uint4 *>(slot) = make_uint4(m[0], m[1], m[2], m[3]);
            }
            bar_arrive(1 + warp * LF_RING + (col & (LF_RING - 1)));
        } else if (sender) {
            g_send(gmsg_out + (size_t)col * 256, m, tag, luma);
        }
    };

    for (int c = 0; c < g.mb_cols; c++) {
        if ((c & 31) == 16) par_next = par_of(rec_next);
        if ((c & 31) == 0 && c) { par_lane = par_next; rec_next = load_rec(c + 32 + lane); }
        const unsigned par = __shfl_sync(0xffffffffu, par_lane, c & 31);
        const bool skip_lf = (par >> 31) != 0;
        /* A macroblock whose filter level is 0 (loopfilter.c:256 skips it) carries all-zero
         * limits here: the masks then pass only where every difference is zero, where each
         * filter is the identity - so it takes the ordinary path and costs no branch. */
        LfParams P;
        P.ilim = par & 255; P.blim = (par >> 8) & 255; P.mblim = (par >> 16) & 255; P.thr = (par >> 24) & 3;
        /* the next macroblock's rows have landed (LF_PF - 2 younger groups may be in flight);
         * its slot is then free for the macroblock LF_PF ahead */
        asm volatile("cp.async.wait_group %0;" ::"n"(LF_PF - 2) : "memory");
        fetch(c + 1, nxt);
        prefetch(c + LF_PF);
        uint8_t *colp = rowp + c * mbw;                     /* my row at this MB's x = 0 */

        /* ---- vertical edges, lane = pixel row ---- */
        int x[8];
        unpack(cur[0], x[0], x[1], x[2], x[3]);
        if (c > 0) {
            if (lane_on) {
                int h0, h1, h2, h3;
                unpack(halo, h0, h1, h2, h3);
                edge8<true, SIMPLE>(h0, h1, h2, h3, x[0], x[1], x[2], x[3], P);
                halo = pack(h0, h1, h2, h3);
            }
        }
        if (lane_on) {
            if (!skip_lf) {
                /* all lanes run the luma sequence; a chroma lane's third and fourth word are
                 * zeros and its second word is taken before the edge at x = 8 touches it */
                int y[8];
                unpack(cur[1], x[4], x[5], x[6], x[7]);
                edge8<false, SIMPLE>(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7], P);
                const unsigned c1 = pack(x[4], x[5], x[6], x[7]);
                unpack(cur[2], y[0], y[1], y[2], y[3]);
                edge8<false, SIMPLE>(x[4], x[5], x[6], x[7], y[0], y[1], y[2], y[3], P);
                unpack(cur[3], y[4], y[5], y[6], y[7]);
                edge8<false, SIMPLE>(y[0], y[1], y[2], y[3], y[4], y[5], y[6], y[7], P);
                cur[1] = luma ? pack(x[4], x[5], x[6], x[7]) : c1;
                cur[2] = pack(y[0], y[1], y[2], y[3]);
                cur[3] = pack(y[4], y[5], y[6], y[7]);
            }
            cur[0] = pack(x[0], x[1], x[2], x[3]);
        }
        if (c > 0) {
            store_prev(colp);
            if (!last_row) send(c - 1);
        }
        /* ---- the 4 rows above arrive as a message from the row above ---- */
        if (top) {
            if (recv_smem) {
                bar_wait(1 + (warp - 1) * LF_RING + (c & (LF_RING - 1)));
                if (receiver) {
                    const uint8_t *slot = s_ring[warp - 1][c & (LF_RING - 1)] + sr_off;
                    *reinterpret_cast<uint4 *>(tile + pi * 16) = *reinterpret_cast<const uint4 *>(slot);
                }
                __syncwarp();
                if (lane == 0) s_rcvd[warp - 1] = (unsigned)c + 1;
            } else {
                unsigned m[4];
                g_recv(gmsg_in + (size_t)c * 256, gw, m, tag, luma, receiver);
                if (receiver) {
                    if (c + 1 < g.mb_cols) g_load(gmsg_in + (size_t)(c + 1) * 256, gw, luma);
                    *reinterpret_cast<uint4 *>(tile + pi * 16) = make_uint4(m[0], m[1], m[2], m[3]);
                }
            }
        }
        {
            /* rows the horizontal edges touch: all of them, or only rows 0..3 for the top edge
             * of a macroblock without inner edges (nothing at all if that has no top either) */
            const int nrows = skip_lf ? 4 : mbw;            /* MB rows entering the tile */
            if (lane_on && (!skip_lf || top) && pi < nrows)
                *reinterpret_cast<uint4 *>(tile + (pi + 4) * 16) = make_uint4(cur[0], cur[1], cur[2], cur[3]);
            __syncwarp();
            /* ---- horizontal edges, lane = pixel column ---- */
            if (lane_on && (!skip_lf || top)) {
                int v[8];
                if (top) {
#pragma unroll
                    for (int r = 0; r < 8; r++) v[r] = tile[r * 16 + pi];
                    edge8<true, SIMPLE>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], P);
#pragma unroll
                    for (int r = 1; r < 4; r++) tile[r * 16 + pi] = (uint8_t)v[r];
                    if (skip_lf) {
#pragma unroll
                        for (int r = 4; r < 7; r++) tile[r * 16 + pi] = (uint8_t)v[r];
                    }
                } else {
#pragma unroll
                    for (int r = 4; r < 8; r++) v[r] = tile[r * 16 + pi];
                }
                if (!skip_lf) {
                    /* tile rows 8..11 are the last real rows of a chroma MB: a chroma lane keeps
                     * them as they are after the first inner edge and plays the two luma-only
                     * edges on its scratch rows */
                    int w[8], u[4], wb[4];
#pragma unroll
                    for (int r = 0; r < 4; r++) w[r] = tile[(r + 8) * 16 + pi];
#pragma unroll
                    for (int r = 4; r < 8; r++) w[r] = tile_hi[(r + 8) * 16 + pi];
#pragma unroll
                    for (int r = 0; r < 4; r++) u[r] = tile_hi[(r + 16) * 16 + pi];
                    edge8<false, SIMPLE>(v[4], v[5], v[6], v[7], w[0], w[1], w[2], w[3], P);
#pragma unroll
                    for (int r = 4; r < 8; r++) tile[r * 16 + pi] = (uint8_t)v[r];
#pragma unroll
                    for (int r = 0; r < 4; r++) wb[r] = w[r];
                    edge8<false, SIMPLE>(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], P);
#pragma unroll
                    for (int r = 0; r < 4; r++) tile[(r + 8) * 16 + pi] = (uint8_t)(luma ? w[r] : wb[r]);
                    edge8<false, SIMPLE>(w[4], w[5], w[6], w[7], u[0], u[1], u[2], u[3], P);
#pragma unroll
                    for (int r = 4; r < 8; r++) tile_hi[(r + 8) * 16 + pi] = (uint8_t)w[r];
#pragma unroll
                    for (int r = 0; r < 4; r++) tile_hi[(r + 16) * 16 + pi] = (uint8_t)u[r];
                }
            }
            __syncwarp();
            if (lane_on && (!skip_lf || top) && pi < nrows) {
                const uint4 v = *reinterpret_cast<const uint4 *>(tile + (pi + 4) * 16);
                cur[0] = v.x; cur[1] = v.y; cur[2] = v.z; cur[3] = v.w;
            }
        }
        /* (this MB's own rows go out one iteration later, after the next MB's left edge) */
        if (receiver && pi >= 1) {                          /* rows -3..-1: this row finishes them */
            uint8_t *ap = plane + (size_t)(mb_row * mbw - 4 + pi) * stride + c * mbw;
            if (luma) *reinterpret_cast<uint4 *>(ap) = *reinterpret_cast<const uint4 *>(tile + pi * 16);
            else *reinterpret_cast<uint2 *>(ap) = *reinterpret_cast<const uint2 *>(tile + pi * 16);
        }
        if (luma) { prev[0] = cur[0]; prev[1] = cur[1]; prev[2] = cur[2]; halo = cur[3]; }
        else { prev[0] = cur[0]; halo = cur[1]; }
        __syncwarp();                                       /* tile is reused by the next MB */
#pragma unroll
        for (int i = 0; i < 4; i++) cur[i] = nxt[i];
    }
    /* the last macroblock of the row */
    store_prev(rowp + g.mb_cols * mbw);
    if (!last_row) send(g.mb_cols - 1);
}

__global__ void __launch_bounds__(LF_ROWS_PER_CTA * 32, LF_MIN_CTAS)
k_loopfilter(const FrameJob *__restrict__ jobs, const int n_jobs, const Geo g,
             unsigned *ticket, const unsigned ticket_base)
{
    __shared__ FrameJob job;
    __shared__ unsigned s_ticket;
    /* [seg][ref][mode class] -> ilim | blim << 8 | mblim << 16 | hev threshold << 24; 0: level 0 */
    __shared__ unsigned s_par[64];
    __shared__ __align__(16) uint8_t s_tile[LF_ROWS_PER_CTA][704];
    __shared__ __align__(16) uint8_t s_scratch[LF_ROWS_PER_CTA][640];
    __shared__ __align__(16) uint8_t s_pf[LF_ROWS_PER_CTA][LF_PF][512];   /* cp.async ring: [slot][lane][16 B] */
    /* message ring of row w -> row w+1: 192 B = luma rows 12..15, U rows 4..7, V rows 4..7, 16 B each */
    __shared__ __align__(16) uint8_t s_ring[LF_ROWS_PER_CTA][LF_RING][192];
    __shared__ volatile unsigned s_rcvd[LF_ROWS_PER_CTA];
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u) - ticket_base;
    if (threadIdx.x < LF_ROWS_PER_CTA) s_rcvd[threadIdx.x] = 0;
    __syncthreads();
    const unsigned t = s_ticket;
    const int ji = t % n_jobs, group = t / n_jobs;
    {
        const unsigned *s = reinterpret_cast<const unsigned *>(&jobs[ji]);
        unsigned *d = reinterpret_cast<unsigned *>(&job);
        for (int i = threadIdx.x; i < (int)(sizeof(FrameJob) / 4); i += LF_ROWS_PER_CTA * 32) d[i] = s[i];
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
        /* limits of that level: loopfilter.c:66-96 and :28-50 */
        const int sharp = h.sharpness_level;
        const bool key = h.frame_type == 0;
        int il = lvl >> (sharp > 0);
        il >>= (sharp > 4);
        if (sharp > 0) il = min(il, 9 - sharp);
        il = max(il, 1);
        const int thr = key ? (lvl >= 40 ? 2 : (lvl >= 15 ? 1 : 0))
                            : (lvl >= 40 ? 3 : (lvl >= 20 ? 2 : (lvl >= 15 ? 1 : 0)));
        s_par[threadIdx.x] = lvl ? (unsigned)il | (unsigned)(2 * lvl + il) << 8 | (unsigned)(2 * (lvl + 2) + il) << 16 | (unsigned)thr << 24 : 0u;
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb_row = group * LF_ROWS_PER_CTA + warp;
    if (mb_row >= g.mb_rows) return;
    if (h.filter_type != 0) lf_row<true>(job, g, mb_row, warp, lane, s_par, s_tile[warp], s_scratch[warp], &s_pf[warp][0][0], s_ring, s_rcvd);
    else lf_row<false>(job, g, mb_row, warp, lane, s_par, s_tile[warp], s_scratch[warp], &s_pf[warp][0][0], s_ring, s_rcvd);
}

/* size of FrameJob.lf_msg: one 256-byte slot per macroblock */
size_t vp8b200_lf_msg_bytes(const Geo &g) { return (size_t)g.mb_cols * g.mb_rows * 256; }

void vp8b200_launch_loopfilter(cudaStream_t s, const FrameJob *jobs, int n_jobs, const Geo &g,
                               unsigned int *ticket, unsigned int ticket_base, int *n_ctas)
{
    int groups = (g.mb_rows + LF_ROWS_PER_CTA - 1) / LF_ROWS_PER_CTA;
    *n_ctas = groups * n_jobs;
    k_loopfilter<<<groups * n_jobs, LF_ROWS_PER_CTA * 32, 0, s>>>(jobs, n_jobs, g, ticket, ticket_base);
}
