/* kernels_intra.cu - intra prediction + residual for the intra macroblocks of a frame.
 *
 * Restates vp8_build_intra_predictors_mby_s / mbuv_s (vp8/common/reconintra.c:139-263,
 * :403-546), vp8_intra4x4_predict (reconintra4x4.c:16-296) with its above-right rule
 * (vp8_intra_prediction_down_copy, :305-317), the 127/129 frame-edge rule of
 * vp8_setup_intra_recon (setupintrarecon.c:15-32), the right-edge rule of vp8_extend_mb_row
 * (extend.c:160-185) and the residual add of decodframe.c:192-304.
 *
 * Intra prediction reads the UNFILTERED reconstruction of the left, above-left, above and
 * above-right neighbours, so intra macroblocks form a wavefront.  One warp owns one intra
 * macroblock.  The host hands over the list of intra MBs sorted by wavefront index c + 2r
 * (all of them on key frames, a handful on P frames, where inter MBs were already finished by
 * k_inter); warps take list entries through an atomic ticket in that order, so every
 * dependency belongs to an earlier ticket, i.e. to a warp that is already running: no
 * deadlock, and no reliance on block scheduling order.  A warp waits only for those of its
 * four neighbours that are themselves intra, by polling the tagged border words they export
 * (no flags, no fences).  The polling loop is left by the WARP as a whole (__all_sync): lanes
 * that break out one by one leave the warp diverged, and everything after the wait - the
 * dependency chain - then runs its shuffles through the divergent-warp fallback, several times
 * slower (measured: profiles/r02_summary.md).
 *
 * Inside a macroblock: borders are gathered into a shared-memory tile (frame-edge values are
 * synthesised, never read from the frame), residuals of all blocks are computed first
 * (lane = 4x4 block; they do not depend on prediction), whole-block modes predict with
 * lane = block, and B_PRED runs its 16 sub-blocks as a 10-step anti-diagonal wavefront with
 * lane = PIXEL (two sub-blocks x 16 pixels per step) driven by a constant-memory table that
 * encodes every directional predictor as a 3-tap or 2-tap average of the edge array.
 */
#include "vp8b200_dev.cuh"

#define INTRA_WARPS 4
#ifndef INTRA_MIN_CTAS
#define INTRA_MIN_CTAS 8          /* caps the kernel at 64 registers */
#endif
#ifndef INTRA_SPIN
#define INTRA_SPIN 0
#endif
#ifndef INTRA_SLEEP
#define INTRA_SLEEP 32
#endif
/* shared-memory accesses by 32-bit shared-space address (no generic pointer to rebuild inside a
 * register-capped loop) */
__device__ __forceinline__ int lds_u8(unsigned a) { int v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_u8(unsigned a, int v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint2 lds_v2(unsigned a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v; }

#define YS 48             /* luma tile pitch: rows -1..15, cols -16..31 ; index (r+1)*48 + 16 + c */
#define CS 16             /* chroma tile pitch: rows -1..7, cols -4..11 ; index (r+1)*16 + 4 + c  */

/* B_PRED predictor table: [mode][pixel] = i0 | i1<<4 | i2<<8 | kind<<12 over the edge array
 * E[0..3] = L3 L2 L1 L0, E[4] = top-left, E[5..12] = A0..A7.
 * kind 0: (E[i0] + 2 E[i1] + E[i2] + 2) >> 2, which with i2 = i0 is also the 2-tap average
 * (E[i0] + E[i1] + 1) >> 1 ; 2: DC ; 3: TM = clamp(E[i0] - E[i1] + E[i2]) with
 * (i0, i1, i2) = (above, top-left, left).  Kinds 0 and 3 are one formula,
 * clamp((E[i0] + w E[i1] + E[i2] + s) >> s) with (w, s) = (2, 2) or (-1, 0): no divergent code. */
__constant__ unsigned short c_bpred[10][16];
/* B_PRED wavefront geometry, [step][which]: blk | active << 4 | (column 3) << 5 | tile offset of the
 * block's pixel (0,0) << 8 */
__constant__ unsigned c_bstep[10][2];

static unsigned short ent(int kind, int a, int b, int c) { return (unsigned short)(a | (b << 4) | (c << 8) | (kind << 12)); }

void vp8b200_upload_intra_constants()
{
    unsigned short t[10][16];
    auto A3 = [](int a, int b, int c) { return ent(0, a, b, c); };
    auto A2 = [](int a, int b) { return ent(0, a, b, a); };   /* (a + b + 1) >> 1 == (a + 2b + a + 2) >> 2 */
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            const int p = r * 4 + c;
            t[VP8B200_B_DC_PRED][p] = ent(2, 0, 0, 0);
            t[VP8B200_B_TM_PRED][p] = ent(3, 5 + c, 4, 3 - r);
            t[VP8B200_B_VE_PRED][p] = A3(4 + c, 5 + c, 6 + c);
            /* rows: (tl,L0,L1) (L0,L1,L2) (L1,L2,L3) (L2,L3,L3) with L_k = E[3-k] */
            t[VP8B200_B_HE_PRED][p] = r == 0 ? A3(4, 3, 2) : r == 1 ? A3(3, 2, 1) : r == 2 ? A3(2, 1, 0) : A3(1, 0, 0);
            { int k = r + c; t[VP8B200_B_LD_PRED][p] = k < 6 ? A3(5 + k, 6 + k, 7 + k) : A3(11, 12, 12); }
            { int k = 3 - r + c; t[VP8B200_B_RD_PRED][p] = A3(k, k + 1, k + 2); }
        }
#define S(m, r, c, v) t[m][(r) * 4 + (c)] = (v)
    /* reconintra4x4.c:178-210 (VR), E = pp */
    S(VP8B200_B_VR_PRED, 3, 0, A3(1, 2, 3)); S(VP8B200_B_VR_PRED, 2, 0, A3(2, 3, 4));
    S(VP8B200_B_VR_PRED, 3, 1, A3(3, 4, 5)); S(VP8B200_B_VR_PRED, 1, 0, A3(3, 4, 5));
    S(VP8B200_B_VR_PRED, 2, 1, A2(4, 5));    S(VP8B200_B_VR_PRED, 0, 0, A2(4, 5));
    S(VP8B200_B_VR_PRED, 3, 2, A3(4, 5, 6)); S(VP8B200_B_VR_PRED, 1, 1, A3(4, 5, 6));
    S(VP8B200_B_VR_PRED, 2, 2, A2(5, 6));    S(VP8B200_B_VR_PRED, 0, 1, A2(5, 6));
    S(VP8B200_B_VR_PRED, 3, 3, A3(5, 6, 7)); S(VP8B200_B_VR_PRED, 1, 2, A3(5, 6, 7));
    S(VP8B200_B_VR_PRED, 2, 3, A2(6, 7));    S(VP8B200_B_VR_PRED, 0, 2, A2(6, 7));
    S(VP8B200_B_VR_PRED, 1, 3, A3(6, 7, 8)); S(VP8B200_B_VR_PRED, 0, 3, A2(7, 8));
    /* :212-240 (VL), pp = Above -> E[5 + i] */
    S(VP8B200_B_VL_PRED, 0, 0, A2(5, 6));     S(VP8B200_B_VL_PRED, 1, 0, A3(5, 6, 7));
    S(VP8B200_B_VL_PRED, 2, 0, A2(6, 7));     S(VP8B200_B_VL_PRED, 0, 1, A2(6, 7));
    S(VP8B200_B_VL_PRED, 1, 1, A3(6, 7, 8));  S(VP8B200_B_VL_PRED, 3, 0, A3(6, 7, 8));
    S(VP8B200_B_VL_PRED, 2, 1, A2(7, 8));     S(VP8B200_B_VL_PRED, 0, 2, A2(7, 8));
    S(VP8B200_B_VL_PRED, 3, 1, A3(7, 8, 9));  S(VP8B200_B_VL_PRED, 1, 2, A3(7, 8, 9));
    S(VP8B200_B_VL_PRED, 0, 3, A2(8, 9));     S(VP8B200_B_VL_PRED, 2, 2, A2(8, 9));
    S(VP8B200_B_VL_PRED, 1, 3, A3(8, 9, 10)); S(VP8B200_B_VL_PRED, 3, 2, A3(8, 9, 10));
    S(VP8B200_B_VL_PRED, 2, 3, A3(9, 10, 11)); S(VP8B200_B_VL_PRED, 3, 3, A3(10, 11, 12));
    /* :242-276 (HD), E = pp */
    S(VP8B200_B_HD_PRED, 3, 0, A2(0, 1));    S(VP8B200_B_HD_PRED, 3, 1, A3(0, 1, 2));
    S(VP8B200_B_HD_PRED, 2, 0, A2(1, 2));    S(VP8B200_B_HD_PRED, 3, 2, A2(1, 2));
    S(VP8B200_B_HD_PRED, 2, 1, A3(1, 2, 3)); S(VP8B200_B_HD_PRED, 3, 3, A3(1, 2, 3));
    S(VP8B200_B_HD_PRED, 2, 2, A2(2, 3));    S(VP8B200_B_HD_PRED, 1, 0, A2(2, 3));
    S(VP8B200_B_HD_PRED, 2, 3, A3(2, 3, 4)); S(VP8B200_B_HD_PRED, 1, 1, A3(2, 3, 4));
    S(VP8B200_B_HD_PRED, 1, 2, A2(3, 4));    S(VP8B200_B_HD_PRED, 0, 0, A2(3, 4));
    S(VP8B200_B_HD_PRED, 1, 3, A3(3, 4, 5)); S(VP8B200_B_HD_PRED, 0, 1, A3(3, 4, 5));
    S(VP8B200_B_HD_PRED, 0, 2, A3(4, 5, 6)); S(VP8B200_B_HD_PRED, 0, 3, A3(5, 6, 7));
    /* :278-294 (HU), pp = Left: L_k = E[3-k] */
    S(VP8B200_B_HU_PRED, 0, 0, A2(3, 2));    S(VP8B200_B_HU_PRED, 0, 1, A3(3, 2, 1));
    S(VP8B200_B_HU_PRED, 0, 2, A2(2, 1));    S(VP8B200_B_HU_PRED, 1, 0, A2(2, 1));
    S(VP8B200_B_HU_PRED, 0, 3, A3(2, 1, 0)); S(VP8B200_B_HU_PRED, 1, 1, A3(2, 1, 0));
    S(VP8B200_B_HU_PRED, 1, 2, A2(1, 0));    S(VP8B200_B_HU_PRED, 2, 0, A2(1, 0));
    S(VP8B200_B_HU_PRED, 1, 3, A3(1, 0, 0)); S(VP8B200_B_HU_PRED, 2, 1, A3(1, 0, 0));
    S(VP8B200_B_HU_PRED, 2, 2, A3(0, 0, 0)); S(VP8B200_B_HU_PRED, 2, 3, A3(0, 0, 0));
    S(VP8B200_B_HU_PRED, 3, 0, A3(0, 0, 0)); S(VP8B200_B_HU_PRED, 3, 1, A3(0, 0, 0));
    S(VP8B200_B_HU_PRED, 3, 2, A3(0, 0, 0)); S(VP8B200_B_HU_PRED, 3, 3, A3(0, 0, 0));
#undef S
    cudaMemcpyToSymbol(c_bpred, t, sizeof t);
    unsigned st[10][2];
    for (int step = 0; step < 10; step++)
        for (int which = 0; which < 2; which++) {
            const int br = (step > 3 ? (step - 2) >> 1 : 0) + which, bc = step - 2 * br;
            const bool act = br <= 3 && bc >= 0 && bc <= 3;
            st[step][which] = act ? (unsigned)((br * 4 + bc) | 16 | (bc == 3 ? 32 : 0) | ((br * 4 * YS + bc * 4) << 8)) : 0u;
        }
    cudaMemcpyToSymbol(c_bstep, st, sizeof st);
}

/* whole-block modes (reconintra.c:139-263, :403-546) for the lane's 4x4 block at (bx, by);
 * T points at tile pixel (0,0) with pitch ts; dc is the precomputed DC value */
__device__ __forceinline__ void block_mode(int mode, const uint8_t *T, int ts, int bx, int by, int dc, unsigned (&px)[4])
{
    const uint8_t *above = T - ts;
    if (mode == VP8B200_DC_PRED) {
        const unsigned v = (unsigned)dc * 0x01010101u;
#pragma unroll
        for (int r = 0; r < 4; r++) px[r] = v;
    } else if (mode == VP8B200_V_PRED) {
        const unsigned v = *reinterpret_cast<const unsigned *>(above + bx);
#pragma unroll
        for (int r = 0; r < 4; r++) px[r] = v;
    } else if (mode == VP8B200_H_PRED) {
#pragma unroll
        for (int r = 0; r < 4; r++) px[r] = (unsigned)T[(by + r) * ts - 1] * 0x01010101u;
    } else { /* TM_PRED */
        const int tl = above[-1];
        int a[4];
#pragma unroll
        for (int c = 0; c < 4; c++) a[c] = above[bx + c] - tl;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int l = T[(by + r) * ts - 1];
            px[r] = pack4(clamp255(l + a[0]), clamp255(l + a[1]), clamp255(l + a[2]), clamp255(l + a[3]));
        }
    }
}


__global__ void __launch_bounds__(INTRA_WARPS * 32, INTRA_MIN_CTAS)
k_intra(const FrameJob *__restrict__ jobs, const int n_jobs, const Geo g, const unsigned max_intra,
        unsigned *ticket, const unsigned ticket_base)
{
    __shared__ FrameJob s_job[INTRA_WARPS];
    __shared__ unsigned s_ticket;
    __shared__ __align__(16) uint8_t s_yt[INTRA_WARPS][17 * YS];
    __shared__ __align__(16) uint8_t s_ct[INTRA_WARPS][2][9 * CS];
    __shared__ __align__(16) short s_res[INTRA_WARPS][24][16];   /* residuals parked until their phase (16 registers less across the wait) */
    __shared__ uint8_t s_modes[INTRA_WARPS][16];
    /* B_PRED, per step and lane: .x = table entry | residual << 16, .y = tile offset of the
     * lane's edge element | tile offset of its pixel << 16 (0xffff: lane idle in this step) */
    __shared__ uint2 s_pre[INTRA_WARPS][10][32];
    __shared__ unsigned short s_bpred[160];              /* per-lane indexing: shared, not constant */
    __shared__ unsigned s_bstep[20];
    for (int i = threadIdx.x; i < 160; i += INTRA_WARPS * 32) s_bpred[i] = (&c_bpred[0][0])[i];
    if (threadIdx.x < 20) s_bstep[threadIdx.x] = (&c_bstep[0][0])[threadIdx.x];
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u) - ticket_base;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned t = s_ticket * INTRA_WARPS + warp;
    const int ji = t % n_jobs;
    const unsigned k = t / n_jobs;
    if (k >= max_intra) return;
    FrameJob &job = s_job[warp];
    {
        const unsigned *s = reinterpret_cast<const unsigned *>(&jobs[ji]);
        unsigned *d = reinterpret_cast<unsigned *>(&job);
        for (int i = lane; i < (int)(sizeof(FrameJob) / 4); i += 32) d[i] = s[i];
    }
    __syncwarp();
    if (k >= job.n_intra) return;
    const unsigned epoch = job.epoch_intra;
    const int mbi = (int)job.intra_list[k];
    const int mb_row = mbi / g.mb_cols, mb_col = mbi - mb_row * g.mb_cols;
    vp8b200_mb mb;
    *reinterpret_cast<uint4 *>(&mb) = __ldg(reinterpret_cast<const uint4 *>(job.mb + mbi));
    const bool up = mb_row != 0, left = mb_col != 0;
    const bool right = mb_col != g.mb_cols - 1;

    /* ---- residuals first: they need only the coefficient records, so their global loads and
     * the IDCT / WHT overlap the wait for the neighbours (lane = 4x4 block) ---- */
    const bool bpred = mb.y_mode == VP8B200_B_PRED;
    int res[16];
    bool has_res = false;
    if (lane < 24) has_res = block_residual(job, mb, lane, !bpred, res);
    if (lane < 24) {                                     /* parked in shared memory until the block is predicted */
        /* clamp(pixel + r, 0, 255) == clamp(pixel + clamp(r, -255, 255), 0, 255) for a pixel in
         * 0..255: clamping here, off the chain, lets the add after the wait be a packed 16x2 one */
#pragma unroll
        for (int i = 0; i < 16; i++) res[i] = max(min(res[i], 255), -255);
        uint4 *o = reinterpret_cast<uint4 *>(s_res[warp][lane]);
        o[0] = make_uint4((res[0] & 0xffff) | (res[1] << 16), (res[2] & 0xffff) | (res[3] << 16),
                          (res[4] & 0xffff) | (res[5] << 16), (res[6] & 0xffff) | (res[7] << 16));
        o[1] = make_uint4((res[8] & 0xffff) | (res[9] << 16), (res[10] & 0xffff) | (res[11] << 16),
                          (res[12] & 0xffff) | (res[13] << 16), (res[14] & 0xffff) | (res[15] << 16));
    }
    /* four rows of four pixels + the parked residual, two pixels per instruction: expand two
     * bytes to 16x2 (PRMT), add (VIADD.16x2), clamp to 0..255 (VIMNMX.S16x2), pack (PRMT) */
    auto add_parked = [&](unsigned (&px)[4]) {
        const uint4 *q = reinterpret_cast<const uint4 *>(s_res[warp][lane]);
        const uint4 q0 = q[0], q1 = q[1];
        const unsigned qw[8] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w };
#pragma unroll
        for (int r = 0; r < 4; r++) {
            unsigned lo = __vadd2(__byte_perm(px[r], 0, 0x4140), qw[2 * r]);
            unsigned hi = __vadd2(__byte_perm(px[r], 0, 0x4342), qw[2 * r + 1]);
            lo = __vmins2(__vmaxs2(lo, 0u), 0x00ff00ffu);
            hi = __vmins2(__vmaxs2(hi, 0u), 0x00ff00ffu);
            px[r] = __byte_perm(lo, hi, 0x6420);
        }
    };
    if (bpred && lane < 16) s_modes[warp][lane] = reinterpret_cast<const uint8_t *>(job.aux + mb.u.aux)[lane];
    const int pix = lane & 15, pr = pix >> 2, pc = pix & 3, which = lane >> 4;
    /* B_PRED runs its 16 sub-blocks as a 10-step anti-diagonal wavefront (block (br,bc) at step
     * bc + 2*br needs left, above and above-right, decodframe.c:200-237), at most two blocks per
     * step: lanes 0-15 are the pixels of the first, 16-31 of the second.  Everything a step
     * needs that does not depend on pixels - the predictor table entry of (mode, pixel) and the
     * residual - is gathered here, before the wait for the neighbours.  (The loops stay
     * rolled: every warp runs this code once per macroblock, so unrolled straight-line code
     * would be fetched from the instruction cache hierarchy with no reuse.) */
    if (bpred) {
        __syncwarp();
        /* lane p of a block fetches element p of the block's edge array E[0..3] = L3..L0,
         * E[4] = top-left, E[5..12] = above / above-right */
        const int e = min(pix, 12);
        const int e_off = e < 4 ? (3 - e) * YS - 1 : e - 5 - YS;
        const int ar_off = -YS + 16 + e - 9;                 /* above-right of column 3: row -1 of the MB (reconintra4x4.c:305-317) */
        const int px_off = pr * YS + pc;
#pragma unroll 1
        for (int step = 0; step < 10; step++) {
            const unsigned sg = s_bstep[step * 2 + which];
            const int blk = sg & 15;
            const bool act = (sg & 16) != 0;
            const int b_off = sg >> 8;                        /* block pixel (0,0) in the tile */
            const int ld_off = (e >= 9 && (sg & 32)) ? ar_off : b_off + e_off;
            const int st_off = act ? b_off + px_off : 0xffff;
            const unsigned ent = s_bpred[s_modes[warp][blk] * 16 + pix];
            s_pre[warp][step][lane] = make_uint2(
                ent | ((unsigned)(unsigned short)s_res[warp][blk][pix] << 16),
                (unsigned)(ld_off & 0xffff) | ((unsigned)st_off << 16));
        }
    }
    uint8_t *YT = s_yt[warp] + YS + 16;                  /* tile pixel (0,0) */
    uint8_t *UT = s_ct[warp][0] + CS + 4, *VT = s_ct[warp][1] + CS + 4;
    uint8_t *const dy = job.dst + g.y_off + (size_t)mb_row * 16 * g.y_stride + mb_col * 16;
    uint8_t *const du = job.dst + g.u_off + (size_t)mb_row * 8 * g.uv_stride + mb_col * 8;
    uint8_t *const dv = job.dst + g.v_off + (size_t)mb_row * 8 * g.uv_stride + mb_col * 8;

    /* ---- borders.  Every finished intra MB exports its bottom row and right column as 16
     * tagged 64-bit words (W0-3 Y row 15, W4-5 U row 7, W6-7 V row 7, W8-11 Y column 15,
     * W12-13 U column 7, W14-15 V column 7; {32 bits of pixels, 32-bit frame tag}).  A word
     * whose tag matches is valid (aligned 64-bit accesses are single-copy atomic), so the
     * consumer polls the words themselves: no flag, no fence, no second round trip, and the
     * producer's frame stores are off the dependency chain.  Inter neighbours were finished by
     * k_inter and are read from the frame; frame edges are synthesised (setupintrarecon.c:15-32,
     * extend.c:160-185).  Lanes 0-7 left column words, 8-15 above row words, 16-18 the top-left
     * pixels of Y/U/V, 19 the above-right luma word.
     *
     * The macroblock is done in two phases, luma then chroma, each: fetch that plane's
     * borders, predict + add the residual, export that plane's words.  The right-hand
     * neighbour's luma (the long part: the B_PRED wavefront) needs only the luma words, so
     * the chroma work of this macroblock overlaps with the neighbour's luma instead of
     * sitting on the dependency chain.  The code of a phase exists once (rolled loop: every
     * warp runs it once per macroblock, straight-line copies would only miss in the
     * instruction cache). ---- */
    /* where this lane's border word goes in the tiles: 1 = four bytes down a column (left
     * neighbour's words), 2 = one word of a row (above / above-right), 3 = the top-left byte */
    uint8_t *sc_ptr = YT;
    int sc_pitch = 0, sc_kind = 0;
    if (lane < 4) { sc_ptr = YT + 4 * lane * YS - 1; sc_pitch = YS; sc_kind = 1; }
    else if (lane < 8) { sc_ptr = (lane < 6 ? UT : VT) + 4 * (lane & 1) * CS - 1; sc_pitch = CS; sc_kind = 1; }
    else if (lane < 12) { sc_ptr = YT - YS + 4 * (lane - 8); sc_kind = 2; }
    else if (lane < 16) { sc_ptr = (lane < 14 ? UT : VT) - CS + 4 * (lane & 1); sc_kind = 2; }
    else if (lane < 19) { sc_ptr = lane == 16 ? YT - YS - 1 : (lane == 17 ? UT : VT) - CS - 1; sc_kind = 3; }
    else if (lane == 19) { sc_ptr = YT - YS + 16; sc_kind = 2; }
    const unsigned long long *msg = job.intra_msg;
    const bool luma_lane = lane < 4 || (lane >= 8 && lane < 12) || lane == 16 || lane == 19;
    /* Each lane < 20 has a fixed role - one border word of one neighbour - whatever the phase;
     * everything about it is settled here, BEFORE the wait: the word's address when the
     * neighbour is intra (pp), else its value: synthesised at the frame edge or read from the
     * frame (inter neighbours were finished by k_inter, so those loads are off the chain too). */
    unsigned w_fixed = 0;
    const unsigned long long *pp_fixed = nullptr;
    if (lane < 20) {
        /* which neighbour this lane reads, which word of its export, or which frame bytes */
        const int grp = lane < 8 ? 0 : (lane < 16 ? 1 : (lane < 19 ? 2 : 3));   /* left, above, above-left, above-right */
        const bool exists = grp == 0 ? left : (up && (grp == 2 ? left : (grp == 3 ? right : true)));
        const int ni = mbi + (grp == 0 ? -1 : (grp == 1 ? -g.mb_cols : (grp == 2 ? -g.mb_cols - 1 : -g.mb_cols + 1)));
        const int j = grp == 0 ? lane : (grp == 1 ? lane - 8 : lane - 16);     /* index inside the group */
        /* plane of the word: words 0-3 Y, 4-5 U, 6-7 V (both for rows and columns) */
        const int pl = grp >= 2 ? (grp == 3 ? 0 : j) : (j < 4 ? 0 : (j < 6 ? 1 : 2));
        const int stride = pl == 0 ? g.y_stride : g.uv_stride;
        const uint8_t *base = pl == 0 ? dy : (pl == 1 ? du : dv);
        const int size = pl == 0 ? 16 : 8;
        if (!exists) {
            /* row above the frame is 127 (incl. top-left and above-right), column left of it 129;
             * above-right of the last column replicates the above MB's last pixel (filled below) */
            w_fixed = (grp == 0 || (grp == 2 && up)) ? 0x81818181u : 0x7f7f7f7fu;
        } else {
            const bool n_intra = ((__ldg(reinterpret_cast<const unsigned *>(job.mb + ni)) >> 16) & 0xff) == VP8B200_INTRA_FRAME;
            if (n_intra) {
                const int word = grp == 0 ? 8 + j : (grp == 1 ? j : (grp == 2 ? 3 + 2 * j : 0));
                pp_fixed = msg + (size_t)ni * 16 + word;
            } else if (grp == 0) {
                /* 4 pixels of the column left of the MB: rows 4*jj .. 4*jj+3 of plane pl */
                const int jj = pl == 0 ? j : (j & 1);
                const uint8_t *q = base + (size_t)(4 * jj) * stride - 1;
                w_fixed = q[0] | (q[stride] << 8) | (q[2 * stride] << 16) | ((unsigned)q[3 * stride] << 24);
            } else if (grp == 1) {
                const int jj = pl == 0 ? j : (j & 1);
                w_fixed = *reinterpret_cast<const unsigned *>(base - stride + 4 * jj);
            } else if (grp == 2) {
                w_fixed = (unsigned)base[-stride - 1] << 24;                  /* same byte position as in the word */
            } else {
                w_fixed = *reinterpret_cast<const unsigned *>(base - stride + size);
            }
        }
    }
#pragma unroll 1
    for (int phase = 0; phase < 2; phase++) {
        const bool mine = lane < 20 && luma_lane == (phase == 0);
        unsigned w = mine ? w_fixed : 0u;
        const unsigned long long *pp = mine ? pp_fixed : nullptr;
        {
            /* every lane that has a tagged word polls it (hard first: the hand-off is on the
             * chain); the WARP leaves the loop together */
            unsigned long long v = 0;
            bool ok = pp == nullptr;
            for (int tries = 0;; tries++) {
                if (!ok) {
                    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(pp) : "memory");
                    ok = (unsigned)(v >> 32) == epoch;
                }
                if (__all_sync(FULL_MASK, ok)) break;
                if (tries > INTRA_SPIN) __nanosleep(INTRA_SLEEP);
            }
            if (pp) w = (unsigned)v;
        }
        if (phase == 0) {
            /* above-right of the last MB column: replicate the last pixel of the above row */
            const unsigned w11 = __shfl_sync(FULL_MASK, w, 11);
            if (lane == 19 && up && !right) w = (w11 >> 24) * 0x01010101u;
        }
        /* scatter into the tiles: per-lane pointer and shape were set up before the wait, so
         * that what follows the wait is a handful of predicated stores, not a six-way branch */
        {
            if (mine && sc_kind == 2) *reinterpret_cast<unsigned *>(sc_ptr) = w;
            if (mine && sc_kind != 2) sc_ptr[0] = (uint8_t)(sc_kind == 3 ? w >> 24 : w);
            if (mine && sc_kind == 1) {
                sc_ptr[sc_pitch] = (uint8_t)(w >> 8);
                sc_ptr[2 * sc_pitch] = (uint8_t)(w >> 16);
                sc_ptr[3 * sc_pitch] = (uint8_t)(w >> 24);
            }
        }
        __syncwarp();
        /* ---- DC value (reconintra.c:167-195, :434-462) when this phase's mode is DC_PRED ---- */
        int dc = 128;
        if ((phase == 0 ? (!bpred && mb.y_mode == VP8B200_DC_PRED) : mb.uv_mode == VP8B200_DC_PRED) && (up || left)) {
            /* ONE warp reduction: the luma sum of lanes 0-15 (phase 0), or the U sum of lanes
             * 16-23 in the low half and the V sum of lanes 24-31 in the high half (phase 1) */
            const uint8_t *T = lane < 16 ? YT : (lane < 24 ? UT : VT);
            const int ts = lane < 16 ? YS : CS, i = lane < 16 ? lane : (lane & 7);
            const unsigned mine_sum = (unsigned)((up ? T[-ts + i] : 0) + (left ? T[i * ts - 1] : 0));
            const bool in_phase = (lane < 16) == (phase == 0);
            const unsigned sums = __reduce_add_sync(FULL_MASK, in_phase ? mine_sum << (lane >= 24 ? 16 : 0) : 0u);
            const int shift = (phase == 0 ? 3 : 2) + (up ? 1 : 0) + (left ? 1 : 0);
            /* luma blocks are predicted by lanes 0-15, U by 16-19, V by 20-23 */
            const unsigned sum = (phase == 1 && lane >= 20) ? sums >> 16 : sums & 0xffffu;
            dc = (int)(sum + (1u << (shift - 1))) >> shift;
        }
        if (phase == 0) {
            if (!bpred) {
                /* whole-block luma modes, lane = 4x4 block */
                if (lane < 16) {
                    const int bx = (lane & 3) * 4, by = (lane >> 2) * 4;
                    unsigned px[4];
                    block_mode(mb.y_mode, YT, YS, bx, by, dc, px);
                    if (has_res) add_parked(px);
                    store4x4(YT + by * YS + bx, YS, px);              /* for the export below; the frame gets it after the hand-off */
                }
            } else {
                /* Per step one shared-memory load per lane - lane p of a block fetches element p
                 * of the block's edge array - then the three taps of the lane's pixel come from
                 * the other lanes by shuffle.  The step's dependent chain is kept to: edge load,
                 * shuffle, four integer operations, store, warp barrier.  Everything else is
                 * off it: the table entry of the NEXT step is fetched (and its lane numbers,
                 * offsets and residual unpacked) while this step's edge load is in flight, and
                 * the B_DC_PRED mean of both blocks comes from ONE warp reduction of the edge
                 * values packed 16 bits per block, issued next to the shuffles in every step (a
                 * uniform branch around it for steps without such a block made the loop 60 %
                 * slower; it was a vote plus four dependent shuffles before). */
                const bool e_dc = pix < 4 || (pix >= 5 && pix < 9);
                const int half = lane & 16;
                const unsigned yt_s = (unsigned)__cvta_generic_to_shared(YT);
                unsigned pre_s = (unsigned)__cvta_generic_to_shared(&s_pre[warp][0][lane]);
                uint2 t = lds_v2(pre_s);
#pragma unroll 1
                for (int step = 0; step < 10; step++) {
                    const int edge = lds_u8(yt_s + (short)(t.y & 0xffff));
                    const int la = half + (t.x & 15), lb = half + ((t.x >> 4) & 15), lc = half + ((t.x >> 8) & 15);
                    const int kind = (t.x >> 12) & 3;
                    const int wb = kind == 3 ? -1 : 2, rs = kind == 3 ? 0 : 2;
                    const int res_px = (short)(t.x >> 16);
                    const unsigned st = t.y >> 16;
                    if (step < 9) pre_s += 32 * sizeof(uint2);
                    t = lds_v2(pre_s);
                    const int ea = __shfl_sync(FULL_MASK, edge, la);
                    const int eb = __shfl_sync(FULL_MASK, edge, lb);
                    const int ec = __shfl_sync(FULL_MASK, edge, lc);
                    int v = clamp255((ea + wb * eb + ec + rs) >> rs);
                    {
                        const unsigned sums = __reduce_add_sync(FULL_MASK, e_dc ? (unsigned)edge << half : 0u);
                        const int dcv = (int)(((sums >> half) & 0xffffu) + 4) >> 3;
                        v = kind == 2 ? dcv : v;
                    }
                    v = clamp255(v + res_px);
                    if (st != 0xffff) sts_u8(yt_s + st, v);
                    __syncwarp();
                }
            }
        } else if (lane >= 16 && lane < 24) {
            /* chroma, lane = 4x4 block */
            const int j = lane & 3, bx = (j & 1) * 4, by = (j >> 1) * 4;
            unsigned px[4];
            block_mode(mb.uv_mode, lane < 20 ? UT : VT, CS, bx, by, dc, px);
            if (has_res) add_parked(px);
            store4x4((lane < 20 ? UT : VT) + by * CS + bx, CS, px);   /* for the export below; the frame gets it after the hand-off */
        }
        /* ---- export this plane's bottom row + right column for the neighbours still to come ---- */
        __syncwarp();
        if (lane < 16 && ((lane & 4) != 0) == (phase == 1)) {
            unsigned x;
            if (lane < 4) x = *reinterpret_cast<const unsigned *>(YT + 15 * YS + 4 * lane);
            else if (lane < 8) x = *reinterpret_cast<const unsigned *>((lane < 6 ? UT : VT) + 7 * CS + 4 * (lane & 1));
            else if (lane < 12) {
                const uint8_t *q = YT + (4 * (lane - 8)) * YS + 15;
                x = q[0] | (q[YS] << 8) | (q[2 * YS] << 16) | ((unsigned)q[3 * YS] << 24);
            } else {
                const uint8_t *q = (lane < 14 ? UT : VT) + (4 * (lane & 1)) * CS + 7;
                x = q[0] | (q[CS] << 8) | (q[2 * CS] << 16) | ((unsigned)q[3 * CS] << 24);
            }
            const unsigned long long v = ((unsigned long long)epoch << 32) | x;
            unsigned long long *p = job.intra_msg + (size_t)mbi * 16 + lane;
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
        }
        /* the finished plane goes out to the frame row by row, AFTER the hand-off: the export
         * must not queue behind these stores */
        if (phase == 0 && lane < 16) {
            const unsigned *r = reinterpret_cast<const unsigned *>(YT + lane * YS);
            *reinterpret_cast<uint4 *>(dy + lane * g.y_stride) = make_uint4(r[0], r[1], r[2], r[3]);
        }
        if (phase == 1 && lane >= 16) {
            const unsigned *r = reinterpret_cast<const unsigned *>((lane < 24 ? UT : VT) + (lane & 7) * CS);
            *reinterpret_cast<uint2 *>((lane < 24 ? du : dv) + (lane & 7) * g.uv_stride) = make_uint2(r[0], r[1]);
        }
    }
}

void vp8b200_launch_intra(cudaStream_t s, const FrameJob *jobs, int n_jobs, const Geo &g,
                          unsigned int max_intra, unsigned int *ticket, unsigned int ticket_base,
                          int *n_ctas)
{
    const unsigned long long warps = (unsigned long long)max_intra * (unsigned)n_jobs;
    const int ctas = (int)((warps + INTRA_WARPS - 1) / INTRA_WARPS);
    *n_ctas = ctas;
    if (ctas) k_intra<<<ctas, INTRA_WARPS * 32, 0, s>>>(jobs, n_jobs, g, max_intra, ticket, ticket_base);
}
