/* kernels_lf.cu - in-loop deblocking filter, normal and simple variants.
 *
 * Restates vp8_loop_filter_frame (vp8/common/loopfilter.c:203-316) with the edge filters of
 * vp8/common/loopfilter_filters.c (lf_packed.cuh), level selection of
 * vp8_loop_filter_frame_init (loopfilter.c:117-201) and the limit tables of
 * vp8_loop_filter_update_sharpness (:66-96).
 *
 * Dependencies.  The reference filters macroblocks in raster order; a macroblock reads pixels
 * its left, above and above-right neighbours have already modified, and inside a macroblock
 * the eight edges (left MB edge, x = 4, 8, 12, top MB edge, y = 4, 8, 12) form one chain.  So a
 * frame offers one chain per macroblock ROW, each row one macroblock behind the row above.
 *
 * Work decomposition (round 2; the round-1 kernel - one warp per row, one pixel line per lane,
 * scalar arithmetic - is kept in profiles/experiments/ and was bound by the integer ALU pipe:
 * 375 warp instructions per macroblock, 41 % of the ALU issue slots at 12 % of HBM):
 *   - every lane filters TWO pixel lines per instruction in packed 16x2 arithmetic
 *     (lf_packed.cuh), and carries a luma line pair AND a chroma line pair (U row k with V row
 *     k: same limits, same edges) as two independent instruction streams: 8 lanes cover a
 *     whole macroblock - luma rows (2k, 2k+1), chroma row k of U and of V - with no lane ever
 *     running an edge on scratch data;
 *   - a warp therefore owns FOUR consecutive macroblock rows ("quarters", lane = 8 q + k),
 *     skewed by one macroblock: in iteration i quarter q works on column i - q, which is
 *     exactly the lag the dependency allows, so three of four row-to-row hand-offs are a
 *     shared-memory store, a __syncwarp and a load inside one warp - no barrier, no polling;
 *   - the fourth hand-off goes to the next warp of the CTA through a ring of slots with one
 *     named barrier per slot (bar.arrive by the producer, bar.sync by the consumer: a waiting
 *     warp issues nothing), and every 16th row to the next CTA through tagged 64-bit words in
 *     global memory ({32 bits of payload, 32-bit frame tag}; an aligned 64-bit access is
 *     single-copy atomic, so a word whose tag matches carries valid data - no flag, no fence).
 *
 * Inside an iteration a quarter filters the vertical edges with lane = line pair (pixels of a
 * row live in 16 + 8 registers, the previous macroblock's last four columns are carried over),
 * transposes through shared memory as 2x2 blocks of packed pairs (8-byte stores, 16-byte
 * loads, two PRMTs per block - pixels are never repacked to bytes between the phases; the
 * chroma (U, V) pairs transpose as plain 32-bit words), filters the horizontal edges with lane
 * = column pair, and transposes back.  A macroblock is stored one iteration later, after the
 * next macroblock's left-edge filter has modified its last three columns, as one 16-byte
 * (luma) / 8-byte (chroma) store per pixel row; its bottom rows go DOWN as the message and the
 * row below stores the three rows its top-edge filter modifies.
 *
 * Filters are branch-free: a macroblock without inner edges, a frame edge or an idle quarter
 * simply carries a limit that never passes (identity); only when no quarter of the warp has
 * inner edges are those six luma + two chroma filters skipped as a whole.  Pixel rows arrive
 * through a lane-private cp.async ring LF_PF macroblocks deep, record words are read eight
 * macroblocks at a time one batch ahead and turned into limits off the chain.
 */
#include "vp8b200_dev.cuh"
#include "lf_packed.cuh"

#ifndef LF_WARPS
#define LF_WARPS 2                /* warps per CTA; a warp owns 4 macroblock rows */
#endif
#define LF_ROWS_PER_CTA (4 * LF_WARPS)
#ifndef LF_RING
#define LF_RING 4                 /* message slots between consecutive warps (power of two) */
#endif
static_assert(1 + (LF_WARPS - 1) * LF_RING <= 16, "one named barrier per (warp pair, ring slot)");
#ifndef LF_PF
#define LF_PF 8                   /* cp.async prefetch distance in macroblocks (power of two, >= 2) */
#endif
#ifndef LF_POLL_SLEEP
#define LF_POLL_SLEEP 100         /* ns between polls of a global message after 8 immediate tries */
#endif

/* shared memory of one warp */
#define LF_T64 704                /* 8 x 80-byte rows of 2x2 blocks + 64: quarters land on alternating bank halves */
#define LF_T32 288                /* 8 x 32-byte chroma rows + 32: quarters land on different banks */
#define LF_MSG 256                /* message: 128 B luma (rows 12-15 as 2x2 blocks) + 128 B chroma (rows 4-7 of U|V) */
struct __align__(16) LfWarpSmem {
    uint8_t pf[LF_PF][32][48];    /* cp.async ring: [slot][lane]{luma row 2k, luma row 2k+1, U row k, V row k} */
    uint8_t t64[2][4][LF_T64];    /* luma transpose tiles, [0] V->H, [1] H->V, per quarter */
    uint8_t t32[2][4][LF_T32];    /* chroma transpose tiles */
    uint8_t intra[4][LF_MSG];     /* [q] message of quarter q to quarter q + 1; [3] = landing slot of a global message */
    uint8_t ring[LF_RING][LF_MSG];/* messages of this warp's quarter 3 to the next warp */
};

__device__ __forceinline__ void bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void bar_wait(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

/* ---- global hand-off (across CTAs): 64 tagged words per macroblock ---------------------- */
__device__ __forceinline__ void st_msg2(uint8_t *p, unsigned a, unsigned b, unsigned tag)
{
    unsigned long long x = ((unsigned long long)tag << 32) | a, y = ((unsigned long long)tag << 32) | b;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(x), "l"(y) : "memory");
}
__device__ __forceinline__ void ld_msg2(const uint8_t *p, unsigned long long &x, unsigned long long &y)
{
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "l"(p) : "memory");
}

/* mode_lf_lut, loopfilter.c:52-63, two bits per y_mode: DC,V,H,TM,ZEROMV -> 1 ; B_PRED -> 0 ;
 * NEARESTMV,NEARMV,NEWMV -> 2 ; SPLITMV -> 3 */
#define LF_MODE_CLASS_LUT (1u | 1u << 2 | 1u << 4 | 1u << 6 | 0u << 8 | 2u << 10 | 2u << 12 | 1u << 14 | 2u << 16 | 3u << 18)

template <bool SIMPLE>
__device__ __forceinline__ void edge_mb(u32 &p3, u32 &p2, u32 &p1, u32 &p0, u32 &q0, u32 &q1, u32 &q2, u32 &q3,
                                        const LfPk &P, u32 EB)
{
    if (SIMPLE) lfp_simple(p1, p0, q0, q1, EB);
    else { LfPk Q = P; Q.mbEB = EB; lfp_mbedge(p3, p2, p1, p0, q0, q1, q2, q3, Q); }
}
template <bool SIMPLE>
__device__ __forceinline__ void edge_in(u32 &p3, u32 &p2, u32 &p1, u32 &p0, u32 &q0, u32 &q1, u32 &q2, u32 &q3,
                                        const LfPk &P)
{
    if (SIMPLE) lfp_simple(p1, p0, q0, q1, P.inEB);
    else lfp_inner(p3, p2, p1, p0, q0, q1, q2, q3, P);
}

/* four macroblock rows (base_row + q), left to right (see the header comment) */
template <bool SIMPLE>
__device__ __forceinline__ void lf_rows(const FrameJob &job, const Geo &g, const int base_row, const int warp,
                                        const int lane, const uint4 *s_par4, LfWarpSmem *sm_all,
                                        volatile unsigned *s_rcvd)
{
    LfWarpSmem &sm = sm_all[warp];
    const unsigned tag = job.epoch_lf;
    const int q = lane >> 3, k = lane & 7;
    const int mb_row = base_row + q;
    const bool row_on = mb_row < g.mb_rows;
    const bool top = mb_row > 0;
    const bool last_row = mb_row == g.mb_rows - 1;
    const int n_cols = g.mb_cols;

    /* my two luma rows and my U / V row, x = 0 */
    uint8_t *const yrow = job.dst + g.y_off + (size_t)(mb_row * 16 + 2 * k) * g.y_stride;
    uint8_t *const urow = job.dst + g.u_off + (size_t)(mb_row * 8 + k) * g.uv_stride;
    uint8_t *const vrow = job.dst + g.v_off + (size_t)(mb_row * 8 + k) * g.uv_stride;
    /* H phase: my column pair / chroma column in the rows above this macroblock row */
    uint8_t *const ytop = job.dst + g.y_off + (size_t)(mb_row * 16 - 4) * g.y_stride + 2 * k;
    uint8_t *const utop = job.dst + g.u_off + (size_t)(mb_row * 8 - 4) * g.uv_stride + k;
    uint8_t *const vtop = job.dst + g.v_off + (size_t)(mb_row * 8 - 4) * g.uv_stride + k;

    /* rows >= 13 (luma) / >= 5 (chroma) of every macroblock are finished and stored by the row
     * below, unless this is the last row */
    const bool own_y0 = row_on && (last_row || 2 * k <= 12), own_y1 = row_on && (last_row || 2 * k + 1 <= 12);
    const bool own_c = !SIMPLE && row_on && (last_row || k <= 4);
    const bool sends = row_on && !last_row;
    const bool send_luma = sends && k >= 6, send_chroma = !SIMPLE && sends && k >= 4;

    /* where quarter q's message goes / comes from */
    const bool q3_smem = warp < LF_WARPS - 1;              /* quarter 3 -> next warp of this CTA */
    const bool q0_smem = warp > 0;                         /* quarter 0 <- previous warp of this CTA */
    const bool q0_global = warp == 0 && base_row > 0;      /* quarter 0 <- previous CTA */
    uint8_t *const gmsg_out = job.lf_msg + (size_t)(mb_row / LF_ROWS_PER_CTA) * n_cols * 512;   /* by sender row group */
    const uint8_t *const gmsg_in = job.lf_msg + (size_t)((mb_row > 0 ? mb_row - 1 : 0) / LF_ROWS_PER_CTA) * n_cols * 512;
    uint8_t *const t64a = sm.t64[0][q], *const t64b = sm.t64[1][q];
    uint8_t *const t32a = sm.t32[0][q], *const t32b = sm.t32[1][q];
    const uint8_t *const msg_in_fixed = q > 0 ? sm.intra[q - 1] : sm.intra[3];

    /* Per-MB decisions (loopfilter.c:245-253), eight macroblocks at a time: lane k of a quarter
     * loads the first record word of macroblock batch + k one batch ahead and turns it into
     * the class index (bit 31 = no inner edges); the loop broadcasts one word per macroblock by
     * shuffle, so neither the record load nor the class arithmetic sits on the chain. */
    const unsigned *mbrec = reinterpret_cast<const unsigned *>(job.mb + (size_t)mb_row * n_cols);
    auto par_of = [&](unsigned rec) -> unsigned {
        const int y_mode = rec & 255, ref = (rec >> 16) & 255, flags = rec >> 24;
        const int mclass = (LF_MODE_CLASS_LUT >> (2 * y_mode)) & 3;
        const bool skip = mclass != 0 && mclass != 3 && (flags & VP8B200_MBF_SKIP);
        return (unsigned)(((flags & 3) << 4) | (ref << 2) | mclass) | (skip ? 0x80000000u : 0u);
    };
    auto load_rec = [&](int col) -> unsigned { return (row_on && col >= 0 && col < n_cols) ? __ldg(mbrec + col * 4) : 0u; };
    unsigned par_lane = par_of(load_rec(k));               /* macroblocks 0..7 */
    unsigned rec_next = load_rec(8 + k), par_next = 0;     /* macroblocks 8..15 */

    /* Pixel rows are prefetched LF_PF macroblocks ahead with cp.async into a lane-private ring
     * (no barrier, only cp.async.wait_group) and read into registers one macroblock ahead. */
    const unsigned pf_base = (unsigned)__cvta_generic_to_shared(&sm.pf[0][lane][0]);
    auto prefetch = [&](int col) {
        const unsigned dst = pf_base + (unsigned)(col & (LF_PF - 1)) * (32 * 48);
        const int on = row_on && col >= 0 && col < n_cols;
        const uint8_t *y0 = yrow + col * 16;
        asm volatile("{ .reg .pred p, c;\n\t"
                     "setp.ne.b32 p, %5, 0;\n\t"
                     "setp.ne.b32 c, %6, 0;\n\t"
                     "@p cp.async.ca.shared.global [%0], [%1], 16;\n\t"
                     "@p cp.async.ca.shared.global [%0+16], [%2], 16;\n\t"
                     "@c cp.async.ca.shared.global [%0+32], [%3], 8;\n\t"
                     "@c cp.async.ca.shared.global [%0+40], [%4], 8;\n\t"
                     "cp.async.commit_group; }"
                     ::"r"(dst), "l"(y0), "l"(y0 + g.y_stride), "l"(urow + col * 8), "l"(vrow + col * 8),
                       "r"(on), "r"((int)(on && !SIMPLE)) : "memory");
    };
    uint4 ra, rb, rc;                                      /* raw rows of the next macroblock: luma a, luma b, {U, V} */
    auto fetch = [&](int col) {
        const uint8_t *p = &sm.pf[col & (LF_PF - 1)][lane][0];
        ra = *reinterpret_cast<const uint4 *>(p);
        rb = *reinterpret_cast<const uint4 *>(p + 16);
        if (!SIMPLE) rc = *reinterpret_cast<const uint4 *>(p + 32);
    };
    /* quarter q runs q macroblocks behind: its columns start at -q */
#pragma unroll
    for (int i = 0; i < LF_PF; i++) prefetch(i - q);
    asm volatile("cp.async.wait_group %0;" ::"n"(LF_PF - 1) : "memory");
    fetch(-q);

    /* the landing slot doubles as the "rows above" of the frame's top row, whose top-edge limit
     * never passes - but only for valid pixel pairs: clear it (uninitialised shared memory is
     * not a pixel pair and can overflow the biased comparison) */
    *reinterpret_cast<uint2 *>(sm.intra[3] + lane * 8) = make_uint2(0u, 0u);
    __syncwarp();

    u32 X[16], Xp[16], C[8], Cp[8];
#pragma unroll
    for (int j = 0; j < 16; j++) Xp[j] = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) Cp[j] = 0;
    const unsigned long long no_msg = (unsigned long long)~tag << 32;
    unsigned long long gw[8];                              /* last read of my 8 words of the next global message */
#pragma unroll
    for (int j = 0; j < 8; j++) gw[j] = no_msg;

    const int n_iter = n_cols + 4;                          /* quarter 3 flushes its last macroblock at n_cols + 3 */
    for (int i = 0; i < n_iter; i++) {
        const int c = i - q;                               /* my column */
        const bool in = row_on && c >= 0 && c < n_cols;
        const bool prev_in = row_on && c >= 1 && c <= n_cols;   /* macroblock c - 1 exists: finish it */
        /* ---- limits of my macroblock ---- */
        if ((c & 7) == 4) par_next = par_of(rec_next);
        if ((c & 7) == 0 && c > 0) { par_lane = par_next; rec_next = load_rec(c + 8 + k); }
        const unsigned par = __shfl_sync(FULL_MASK, par_lane, (lane & 24) | (c & 7));
        const bool skip_lf = (par >> 31) != 0 || !in;
        LfPk P;
        {
            const uint4 v = s_par4[in ? (par & 63) : 64];
            P.ilimB = v.x; P.mbEB = v.y; P.inEB = skip_lf ? LFP_NEVER : v.z; P.thrB = v.w;
        }
        const u32 EB_left = c > 0 ? P.mbEB : LFP_NEVER;    /* no filtering across the frame edge */
        const u32 EB_top = top ? P.mbEB : LFP_NEVER;
        const bool any_inner = __any_sync(FULL_MASK, !skip_lf);

        /* ---- the next macroblock's rows have landed; its slot is free for column + LF_PF ---- */
        {
            u32 w[4], z[4];
            w[0] = ra.x; w[1] = ra.y; w[2] = ra.z; w[3] = ra.w; z[0] = rb.x; z[1] = rb.y; z[2] = rb.z; z[3] = rb.w;
#pragma unroll
            for (int j = 0; j < 4; j++) lfp_unpack(w[j], z[j], X[4 * j], X[4 * j + 1], X[4 * j + 2], X[4 * j + 3]);
            if (!SIMPLE) {
                lfp_unpack(rc.x, rc.z, C[0], C[1], C[2], C[3]);
                lfp_unpack(rc.y, rc.w, C[4], C[5], C[6], C[7]);
            }
        }
        asm volatile("cp.async.wait_group %0;" ::"n"(LF_PF - 2) : "memory");
        fetch(c + 1);
        prefetch(c + LF_PF);

        /* ---- vertical edges, lane = line pair ---- */
        edge_mb<SIMPLE>(Xp[12], Xp[13], Xp[14], Xp[15], X[0], X[1], X[2], X[3], P, EB_left);
        if (!SIMPLE) edge_mb<false>(Cp[4], Cp[5], Cp[6], Cp[7], C[0], C[1], C[2], C[3], P, EB_left);
        if (any_inner) {
            edge_in<SIMPLE>(X[0], X[1], X[2], X[3], X[4], X[5], X[6], X[7], P);
            if (!SIMPLE) edge_in<false>(C[0], C[1], C[2], C[3], C[4], C[5], C[6], C[7], P);
            edge_in<SIMPLE>(X[4], X[5], X[6], X[7], X[8], X[9], X[10], X[11], P);
            edge_in<SIMPLE>(X[8], X[9], X[10], X[11], X[12], X[13], X[14], X[15], P);
        }

        /* ---- macroblock c - 1 is final now: store it, hand its bottom rows down ---- */
        {
            u32 wa[4], wb[4];
#pragma unroll
            for (int j = 0; j < 4; j++) lfp_pack(Xp[4 * j], Xp[4 * j + 1], Xp[4 * j + 2], Xp[4 * j + 3], wa[j], wb[j]);
            uint8_t *dsty = yrow + (c - 1) * 16;
            if (prev_in && own_y0) *reinterpret_cast<uint4 *>(dsty) = make_uint4(wa[0], wa[1], wa[2], wa[3]);
            if (prev_in && own_y1) *reinterpret_cast<uint4 *>(dsty + g.y_stride) = make_uint4(wb[0], wb[1], wb[2], wb[3]);
            if (!SIMPLE) {
                u32 ua[2], va[2];
                lfp_pack(Cp[0], Cp[1], Cp[2], Cp[3], ua[0], va[0]);
                lfp_pack(Cp[4], Cp[5], Cp[6], Cp[7], ua[1], va[1]);
                if (prev_in && own_c) {
                    *reinterpret_cast<uint2 *>(urow + (c - 1) * 8) = make_uint2(ua[0], ua[1]);
                    *reinterpret_cast<uint2 *>(vrow + (c - 1) * 8) = make_uint2(va[0], va[1]);
                }
            }
        }
        {
            /* quarter 3 of the last warp sends through global memory, everybody else through
             * shared memory: quarters 0-2 to the next quarter, quarter 3 to the next warp */
            const bool to_global = q == 3 && !q3_smem;
            const int cp3 = i - 4;                         /* column quarter 3 is sending */
            const bool q3_active = cp3 >= 0 && cp3 < n_cols && (base_row + 3 < g.mb_rows - 1);
            if (q3_smem && q3_active) {
                while ((int)(cp3 - s_rcvd[warp]) >= LF_RING) { }          /* ring full: wait for the consumer */
            }
            uint8_t *slot = q < 3 ? sm.intra[q] : sm.ring[cp3 & (LF_RING - 1)];
            if (prev_in && !to_global) {
                if (send_luma) {
                    uint8_t *d = slot + (k - 6) * 8;       /* [k'][lane 6: X(2k'), X(2k'+1) | lane 7: ...] */
#pragma unroll
                    for (int j = 0; j < 8; j++) *reinterpret_cast<uint2 *>(d + j * 16) = make_uint2(Xp[2 * j], Xp[2 * j + 1]);
                }
                if (send_chroma) {
                    uint8_t *d = slot + 128 + (k - 4) * 32;
                    *reinterpret_cast<uint4 *>(d) = make_uint4(Cp[0], Cp[1], Cp[2], Cp[3]);
                    *reinterpret_cast<uint4 *>(d + 16) = make_uint4(Cp[4], Cp[5], Cp[6], Cp[7]);
                }
            }
            if (prev_in && to_global) {
                uint8_t *gm = gmsg_out + (size_t)(c - 1) * 512;
                if (send_luma) {
#pragma unroll
                    for (int j = 0; j < 8; j++) st_msg2(gm + (j * 4 + (k - 6) * 2) * 8, Xp[2 * j], Xp[2 * j + 1], tag);
                }
                if (send_chroma) {
#pragma unroll
                    for (int j = 0; j < 4; j++) st_msg2(gm + (32 + (k - 4) * 8 + 2 * j) * 8, Cp[2 * j], Cp[2 * j + 1], tag);
                }
            }
            if (q3_smem && q3_active) bar_arrive(1 + warp * LF_RING + (cp3 & (LF_RING - 1)));
        }

        /* ---- a message from the previous CTA lands in shared memory (quarter 0 of warp 0) ---- */
        if (q0_global && q == 0) {
            if (in && (!SIMPLE || k < 4)) {               /* the simple filter sends no chroma words */
                const uint8_t *gm = gmsg_in + (size_t)c * 512 + k * 64;
                int tries = 0;
                for (;;) {
                    bool ok = true;
#pragma unroll
                    for (int j = 0; j < 8; j++) ok = ok && (unsigned)(gw[j] >> 32) == tag;
                    if (ok) break;
                    if (++tries > 8) __nanosleep(LF_POLL_SLEEP);
#pragma unroll
                    for (int j = 0; j < 4; j++) ld_msg2(gm + j * 16, gw[2 * j], gw[2 * j + 1]);
                }
                uint8_t *d = sm.intra[3] + k * 32;
                *reinterpret_cast<uint4 *>(d) = make_uint4((unsigned)gw[0], (unsigned)gw[1], (unsigned)gw[2], (unsigned)gw[3]);
                *reinterpret_cast<uint4 *>(d + 16) = make_uint4((unsigned)gw[4], (unsigned)gw[5], (unsigned)gw[6], (unsigned)gw[7]);
                /* read the next macroblock's message one iteration ahead: a row that runs a little
                 * behind the row above never waits for L2 */
                if (c + 1 < n_cols) {
#pragma unroll
                    for (int j = 0; j < 4; j++) ld_msg2(gm + 512 + j * 16, gw[2 * j], gw[2 * j + 1]);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; j++) gw[j] = no_msg;
                }
            }
        }

        /* ---- V -> H: 2x2 blocks of packed pairs (luma), words (chroma) ---- */
#pragma unroll
        for (int j = 0; j < 8; j++) *reinterpret_cast<uint2 *>(t64a + j * 80 + k * 8) = make_uint2(X[2 * j], X[2 * j + 1]);
        if (!SIMPLE) {
            *reinterpret_cast<uint4 *>(t32a + k * 32) = make_uint4(C[0], C[1], C[2], C[3]);
            *reinterpret_cast<uint4 *>(t32a + k * 32 + 16) = make_uint4(C[4], C[5], C[6], C[7]);
        }
        __syncwarp();
        const uint8_t *msg_in = msg_in_fixed;
        if (q0_smem && i < n_cols) {
            bar_wait(1 + (warp - 1) * LF_RING + (i & (LF_RING - 1)));
            if (q == 0) msg_in = sm_all[warp - 1].ring[i & (LF_RING - 1)];
        }

        /* ---- horizontal edges, lane = column pair: Y[0..3] rows -4..-1, Y[4..19] rows 0..15 ---- */
        u32 Y[20], D[12];
        {
            const uint4 m = *reinterpret_cast<const uint4 *>(msg_in + k * 16);
            Y[0] = lfp_prmt(m.x, m.y, 0x5410u); Y[1] = lfp_prmt(m.x, m.y, 0x7632u);
            Y[2] = lfp_prmt(m.z, m.w, 0x5410u); Y[3] = lfp_prmt(m.z, m.w, 0x7632u);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint4 v = *reinterpret_cast<const uint4 *>(t64a + k * 80 + j * 16);
                Y[4 + 4 * j] = lfp_prmt(v.x, v.y, 0x5410u); Y[5 + 4 * j] = lfp_prmt(v.x, v.y, 0x7632u);
                Y[6 + 4 * j] = lfp_prmt(v.z, v.w, 0x5410u); Y[7 + 4 * j] = lfp_prmt(v.z, v.w, 0x7632u);
            }
            if (!SIMPLE) {
#pragma unroll
                for (int r = 0; r < 4; r++) D[r] = *reinterpret_cast<const u32 *>(msg_in + 128 + r * 32 + k * 4);
#pragma unroll
                for (int r = 0; r < 8; r++) D[4 + r] = *reinterpret_cast<const u32 *>(t32a + r * 32 + k * 4);
            }
        }
        edge_mb<SIMPLE>(Y[0], Y[1], Y[2], Y[3], Y[4], Y[5], Y[6], Y[7], P, EB_top);
        if (!SIMPLE) edge_mb<false>(D[0], D[1], D[2], D[3], D[4], D[5], D[6], D[7], P, EB_top);
        if (any_inner) {
            edge_in<SIMPLE>(Y[4], Y[5], Y[6], Y[7], Y[8], Y[9], Y[10], Y[11], P);
            if (!SIMPLE) edge_in<false>(D[4], D[5], D[6], D[7], D[8], D[9], D[10], D[11], P);
            edge_in<SIMPLE>(Y[8], Y[9], Y[10], Y[11], Y[12], Y[13], Y[14], Y[15], P);
            edge_in<SIMPLE>(Y[12], Y[13], Y[14], Y[15], Y[16], Y[17], Y[18], Y[19], P);
        }
        /* rows -3..-1 are final: this row stores them (two adjacent luma bytes / one byte per plane) */
        if (in && top) {
            uint8_t *py = ytop + c * 16;
#pragma unroll
            for (int r = 1; r < 4; r++)
                *reinterpret_cast<unsigned short *>(py + (size_t)r * g.y_stride) = (unsigned short)lfp_prmt(Y[r], 0u, 0x4420u);
            if (!SIMPLE) {
#pragma unroll
                for (int r = 1; r < 4; r++) {
                    utop[(size_t)r * g.uv_stride + c * 8] = (uint8_t)D[r];
                    vtop[(size_t)r * g.uv_stride + c * 8] = (uint8_t)(D[r] >> 16);
                }
            }
        }

        /* ---- H -> V ---- */
#pragma unroll
        for (int j = 0; j < 8; j++) *reinterpret_cast<uint2 *>(t64b + j * 80 + k * 8) = make_uint2(Y[4 + 2 * j], Y[5 + 2 * j]);
        if (!SIMPLE) {
#pragma unroll
            for (int r = 0; r < 8; r++) *reinterpret_cast<u32 *>(t32b + r * 32 + k * 4) = D[4 + r];
        }
        __syncwarp();
        if (q0_smem && i < n_cols && lane == 0) s_rcvd[warp - 1] = (unsigned)i + 1;   /* the ring slot may be reused */
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint4 v = *reinterpret_cast<const uint4 *>(t64b + k * 80 + j * 16);
            Xp[4 * j] = lfp_prmt(v.x, v.y, 0x5410u); Xp[4 * j + 1] = lfp_prmt(v.x, v.y, 0x7632u);
            Xp[4 * j + 2] = lfp_prmt(v.z, v.w, 0x5410u); Xp[4 * j + 3] = lfp_prmt(v.z, v.w, 0x7632u);
        }
        if (!SIMPLE) {
            const uint4 v0 = *reinterpret_cast<const uint4 *>(t32b + k * 32), v1 = *reinterpret_cast<const uint4 *>(t32b + k * 32 + 16);
            Cp[0] = v0.x; Cp[1] = v0.y; Cp[2] = v0.z; Cp[3] = v0.w; Cp[4] = v1.x; Cp[5] = v1.y; Cp[6] = v1.z; Cp[7] = v1.w;
        }
    }
}

__global__ void __launch_bounds__(LF_WARPS * 32)
k_loopfilter(const FrameJob *__restrict__ jobs, const int n_jobs, const Geo g,
             unsigned *ticket, const unsigned ticket_base)
{
    extern __shared__ __align__(16) uint8_t s_dyn[];
    LfWarpSmem *sm_all = reinterpret_cast<LfWarpSmem *>(s_dyn);
    __shared__ FrameJob job;
    __shared__ unsigned s_ticket;
    /* [seg][ref][mode class] -> limits of that level, packed and biased (lf_packed.cuh); [64] never passes */
    __shared__ uint4 s_par4[65];
    __shared__ volatile unsigned s_rcvd[LF_WARPS];
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u) - ticket_base;
    if (threadIdx.x < LF_WARPS) s_rcvd[threadIdx.x] = 0;
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
    for (int e = threadIdx.x; e < 65; e += blockDim.x) {
        /* vp8_loop_filter_frame_init, loopfilter.c:117-201 */
        const int seg = (e >> 4) & 3, ref = (e >> 2) & 3, mode = e & 3;
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
        uint4 v;
        if (e == 64) {
            v = make_uint4(LFP_NEVER, LFP_NEVER, LFP_NEVER, LFP_NEVER);
        } else if (lvl == 0) {
            /* loopfilter.c:256 skips the macroblock: all-zero limits pass only where every
             * difference is zero, where each filter is the identity */
            v = make_uint4(K2(0x8000), K2(0x8001), K2(0x8001), K2(0x8000));
        } else {
            const int blim = 2 * lvl + il, mblim = 2 * (lvl + 2) + il;
            v = make_uint4(K2(il | 0x8000), K2((2 * mblim + 1) | 0x8000), K2((2 * blim + 1) | 0x8000), K2(thr | 0x8000));
        }
        s_par4[e] = v;
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int base_row = group * LF_ROWS_PER_CTA + warp * 4;
    if (base_row >= g.mb_rows) return;
    if (h.filter_type != 0) lf_rows<true>(job, g, base_row, warp, lane, s_par4, sm_all, s_rcvd);
    else lf_rows<false>(job, g, base_row, warp, lane, s_par4, sm_all, s_rcvd);
}

size_t vp8b200_lf_msg_bytes(const Geo &g)
{
    return (size_t)((g.mb_rows + LF_ROWS_PER_CTA - 1) / LF_ROWS_PER_CTA) * g.mb_cols * 512;
}

void vp8b200_launch_loopfilter(cudaStream_t s, const FrameJob *jobs, int n_jobs, const Geo &g,
                               unsigned int *ticket, unsigned int ticket_base, int *n_ctas)
{
    static bool attr_set[64];
    const size_t smem = sizeof(LfWarpSmem) * LF_WARPS;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaFuncSetAttribute(k_loopfilter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set[dev] = true;
    }
    int groups = (g.mb_rows + LF_ROWS_PER_CTA - 1) / LF_ROWS_PER_CTA;
    *n_ctas = groups * n_jobs;
    k_loopfilter<<<groups * n_jobs, LF_WARPS * 32, smem, s>>>(jobs, n_jobs, g, ticket, ticket_base);
}
