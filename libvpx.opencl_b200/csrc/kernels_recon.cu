/* kernels_recon.cu - prediction + residual.
 *
 *   k_inter : every inter macroblock of the frame, fully parallel (one warp per MB).
 *             Restates vp8_build_inter_predictors_mb (vp8/common/reconinter.c:560-573) with
 *             the sub-pixel filters of vp8/common/filter.c, fused with the residual
 *             (dequant + WHT + IDCT + add, decodframe.c:252-304).
 *   (k_intra, the intra macroblocks, lives in kernels_intra.cu)
 *
 * Bilinear prediction is evaluated by the same code as six-tap: a two-tap filter {f0,f1} is
 * the six-tap {0,0,f0,f1,0,0}; rounding is identical ((x+64)>>7) and the six-tap clamp is a
 * no-op on bilinear values (filter.c:41-129 vs :271-350).  A zero fraction is the identity
 * tap {0,0,128,0,0,0}, which is also what the reference's full-pel copy produces.
 */
#include "vp8b200_dev.cuh"

#define WARPS_PER_CTA 8

/* taps packed for dp4a: .x = taps 0..3, .y = taps 4,5,0,0 ; [0..7] six-tap, [8..15] bilinear */
__constant__ int2 c_taps[16];
/* plain 6-tap sets for k_inter16's shifted-tap scheme: .x = taps 0..3, .y = taps 4,5,0,0; the
 * identity tap 128 is stored as 0 (a zero fraction selects the source bytes) */
__constant__ int2 c_tapw[16];

static __host__ int pack_s8(int a, int b, int c, int d)
{
    return (a & 255) | ((b & 255) << 8) | ((c & 255) << 16) | ((d & 255) << 24);
}

/* called by vp8b200_create on the context's device (constant memory is per device) */
void vp8b200_upload_constants()
{
    static const int six[8][6] = {{0, 0, 128, 0, 0, 0},   {0, -6, 123, 12, -1, 0},
                                  {2, -11, 108, 36, -8, 1}, {0, -9, 93, 50, -6, 0},
                                  {3, -16, 77, 77, -16, 3}, {0, -6, 50, 93, -9, 0},
                                  {1, -8, 36, 108, -11, 2}, {0, -1, 12, 123, -6, 0}};
    static const int bil[8][2] = {{128, 0}, {112, 16}, {96, 32}, {80, 48},
                                  {64, 64}, {48, 80},  {32, 96}, {16, 112}};
    int2 h[16];
    /* .x multiplies window bytes 0..3, .y multiplies window bytes 2..5 (so .y = {*, 0, t4, t5}).
     * 128 does not fit a signed byte: the full-weight tap is split as 127 in .x plus 1 on the
     * same pixel (byte 2 of the window = byte 0 of the .y operand). */
    for (int i = 0; i < 8; i++) {
        const int s2 = six[i][2] == 128 ? 127 : six[i][2], s2b = six[i][2] == 128 ? 1 : 0;
        h[i].x = pack_s8(six[i][0], six[i][1], s2, six[i][3]);
        h[i].y = pack_s8(s2b, 0, six[i][4], six[i][5]);
        const int b2 = bil[i][0] == 128 ? 127 : bil[i][0], b2b = bil[i][0] == 128 ? 1 : 0;
        h[8 + i].x = pack_s8(0, 0, b2, bil[i][1]);
        h[8 + i].y = pack_s8(b2b, 0, 0, 0);
    }
    cudaMemcpyToSymbol(c_taps, h, sizeof h);
    /* plain taps for k_inter16's shifted-tap scheme; the identity tap 128 is stored as 0 */
    int2 w[16];
    for (int i = 0; i < 8; i++) {
        const int c2 = six[i][2] == 128 ? 0 : six[i][2];
        w[i].x = pack_s8(six[i][0], six[i][1], c2, six[i][3]);
        w[i].y = pack_s8(six[i][4], six[i][5], 0, 0);
        const int b0 = bil[i][0] == 128 ? 0 : bil[i][0];
        w[8 + i].x = pack_s8(0, 0, b0, bil[i][1]);
        w[8 + i].y = 0;
    }
    cudaMemcpyToSymbol(c_tapw, w, sizeof w);
}

/* One six-tap evaluation: lo = window bytes 0..3, mid = window bytes 2..5 (see the tap table).
 * Result: clamp(((sum) + 64) >> 7); the rounding constant rides in the dp4a accumulator. */
__device__ __forceinline__ int filt6(unsigned lo, unsigned mid, int2 t)
{
    int s = dp4a_us(lo, t.x, 64);
    s = dp4a_us(mid, t.y, s);
    return clamp255(s >> 7);
}
/* window bytes 2..5 for output j of a row held as v0 = bytes 0..3, v1 = 4..7, v2 = 8.. */
#define WIN_LO(v0, v1, j) ((j) ? __funnelshift_r((v0), (v1), 8 * (j)) : (v0))
#define WIN_MID(v0, v1, v2, j) ((j) < 2 ? __funnelshift_r((v0), (v1), 8 * ((j) + 2)) : (j) == 2 ? (v1) : __funnelshift_r((v1), (v2), 8))

struct MV { int row, col; };

/* reconinter.c:348-368; edges in 1/8 pel (decodframe.c:351-365) */
__device__ __forceinline__ MV clamp_mv(MV mv, int left, int right, int top, int bottom)
{
    if (mv.col < left - (19 << 3)) mv.col = left - (16 << 3);
    else if (mv.col > right + (18 << 3)) mv.col = right + (16 << 3);
    if (mv.row < top - (19 << 3)) mv.row = top - (16 << 3);
    else if (mv.row > bottom + (18 << 3)) mv.row = bottom + (16 << 3);
    return mv;
}

/* reconinter.c:371-382 */
__device__ __forceinline__ MV clamp_uvmv(MV mv, int left, int right, int top, int bottom)
{
    if (2 * mv.col < left - (19 << 3)) mv.col = (left - (16 << 3)) >> 1;
    if (2 * mv.col > right + (18 << 3)) mv.col = (right + (16 << 3)) >> 1;
    if (2 * mv.row < top - (19 << 3)) mv.row = (top - (16 << 3)) >> 1;
    if (2 * mv.row > bottom + (18 << 3)) mv.row = (bottom + (16 << 3)) >> 1;
    return mv;
}

/* 4x4 block at `src` (points at the integer-pel position of the block's top-left pixel in
 * the reference plane), fraction (xo, yo), filter table base tb (0 six-tap, 8 bilinear).
 * need_v: warp-uniform "some lane has yo != 0" (rows outside the block are skipped when
 * nobody needs them). */
__device__ __forceinline__ void predict4x4(const uint8_t *src, int stride, int xo, int yo, int tb,
                                           bool need_v, unsigned (&px)[4])
{
    const int2 th = c_taps[tb + xo], tv = c_taps[tb + yo];
    /* the 9 source bytes x-2 .. x+6 of a row live in 3 aligned words */
    const uint8_t *p = src - 2;
    unsigned sh = ((unsigned)(uintptr_t)p & 3u) * 8u;
    const unsigned *wp = reinterpret_cast<const unsigned *>(p - (sh >> 3));
    const int wstride = stride >> 2;
    /* first pass: column j of the intermediate, rows -2..6 packed as bytes */
    unsigned c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0};
    const int r_lo = need_v ? 0 : 2, r_hi = need_v ? 9 : 6;
#pragma unroll
    for (int r = 0; r < 9; r++) {
        if (r < r_lo || r >= r_hi) continue;
        const unsigned *row = wp + (r - 2) * wstride;
        unsigned w0 = __ldg(row), w1 = __ldg(row + 1), w2 = __ldg(row + 2);
        unsigned v0 = __funnelshift_r(w0, w1, sh);      /* p0..p3 */
        unsigned v1 = __funnelshift_r(w1, w2, sh);      /* p4..p7 */
        unsigned v2 = w2 >> sh;                         /* p8 ..  */
#pragma unroll
        for (int j = 0; j < 4; j++) {
            unsigned f = (unsigned)filt6(WIN_LO(v0, v1, j), WIN_MID(v0, v1, v2, j), th);
            if (r < 4) c0[j] |= f << (8 * r);
            else if (r < 8) c1[j] |= f << (8 * (r - 4));
            else c2[j] = f;
        }
    }
    /* second pass down each column */
    int o[4][4];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int r = 0; r < 4; r++) {
            o[r][j] = filt6(WIN_LO(c0[j], c1[j], r), WIN_MID(c0[j], c1[j], c2[j], r), tv);
        }
#pragma unroll
    for (int r = 0; r < 4; r++) px[r] = pack4(o[r][0], o[r][1], o[r][2], o[r][3]);
}

/* position of the lane's block: plane offset in the allocation, stride, pixel x / y */
struct BlkPos { int off, stride, x, y; };
__device__ __forceinline__ BlkPos block_pos(const Geo &g, int lane, int mb_row, int mb_col)
{
    BlkPos b;
    if (lane < 16) {
        b.off = g.y_off; b.stride = g.y_stride;
        b.x = mb_col * 16 + (lane & 3) * 4; b.y = mb_row * 16 + (lane >> 2) * 4;
    } else {
        int j = lane & 3;
        b.off = lane < 20 ? g.u_off : g.v_off; b.stride = g.uv_stride;
        b.x = mb_col * 8 + (j & 1) * 4; b.y = mb_row * 8 + (j >> 1) * 4;
    }
    return b;
}

/* SPLITMV macroblocks: one warp per MB, lane = 4x4 block with its own motion vector */
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_inter_split(const FrameJob *__restrict__ jobs, const Geo g)
{
    __shared__ FrameJob job;
    {
        const unsigned *s = reinterpret_cast<const unsigned *>(&jobs[blockIdx.y]);
        unsigned *d = reinterpret_cast<unsigned *>(&job);
        for (int i = threadIdx.x; i < (int)(sizeof(FrameJob) / 4); i += WARPS_PER_CTA * 32) d[i] = s[i];
    }
    __syncthreads();
    if (job.hdr.frame_type == 0 || job.n_split == 0) return;   /* nothing to do in this frame */
    const int lane = threadIdx.x & 31;
    const int mbi = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    if (mbi >= g.mb_cols * g.mb_rows) return;
    vp8b200_mb mb;
    *reinterpret_cast<uint4 *>(&mb) = __ldg(reinterpret_cast<const uint4 *>(job.mb + mbi));
    if (mb.ref_frame == VP8B200_INTRA_FRAME) return;           /* k_intra's */
    if (mb.y_mode != VP8B200_SPLITMV) return;                  /* k_inter16's */
    const int mb_row = mbi / g.mb_cols, mb_col = mbi - mb_row * g.mb_cols;
    const bool active = lane < 24;
    const int blk = active ? lane : 0;
    const BlkPos bp = block_pos(g, blk, mb_row, mb_col);

    /* motion vector of the lane's block */
    const int e_left = -((mb_col * 16) << 3), e_right = ((g.mb_cols - 1 - mb_col) * 16) << 3;
    const int e_top = -((mb_row * 16) << 3), e_bottom = ((g.mb_rows - 1 - mb_row) * 16) << 3;
    const bool clampmv = mb.flags & VP8B200_MBF_CLAMP_MVS;
    const int fpmask = job.hdr.full_pixel ? ~7 : ~0;
    MV mv;
    if (mb.y_mode != VP8B200_SPLITMV) {
        mv.row = mb.u.mv.row; mv.col = mb.u.mv.col;
        if (clampmv) mv = clamp_mv(mv, e_left, e_right, e_top, e_bottom);
        if (blk >= 16) {                                       /* reconinter.c:419-424 */
            mv.row = s16(mv.row + (1 | (mv.row >> 31)));
            mv.col = s16(mv.col + (1 | (mv.col >> 31)));
            mv.row = (mv.row / 2) & fpmask;
            mv.col = (mv.col / 2) & fpmask;
        }
    } else {
        const unsigned *a = reinterpret_cast<const unsigned *>(job.aux + mb.u.aux);
        if (blk < 16) {
            unsigned v = __ldg(a + blk);
            mv.row = (short)(v & 0xffff); mv.col = (short)(v >> 16);
            if (clampmv) mv = clamp_mv(mv, e_left, e_right, e_top, e_bottom);
        } else {                                               /* reconinter.c:520-558 */
            int j = blk & 3, y0 = (j >> 1) * 8 + (j & 1) * 2;
            unsigned v0 = __ldg(a + y0), v1 = __ldg(a + y0 + 1), v2 = __ldg(a + y0 + 4), v3 = __ldg(a + y0 + 5);
            int tr = (short)(v0 & 0xffff) + (short)(v1 & 0xffff) + (short)(v2 & 0xffff) + (short)(v3 & 0xffff);
            int tc = (short)(v0 >> 16) + (short)(v1 >> 16) + (short)(v2 >> 16) + (short)(v3 >> 16);
            tr += 4 + ((tr >> 31) << 3);
            tc += 4 + ((tc >> 31) << 3);
            mv.row = s16((tr / 8) & fpmask);
            mv.col = s16((tc / 8) & fpmask);
            if (clampmv) mv = clamp_uvmv(mv, e_left, e_right, e_top, e_bottom);
        }
    }

    const uint8_t *ref = job.ref[mb.ref_frame] + bp.off;
    const uint8_t *src = ref + (bp.y + (mv.row >> 3)) * bp.stride + bp.x + (mv.col >> 3);
    const int xo = mv.col & 7, yo = mv.row & 7;
    const bool need_v = __any_sync(FULL_MASK, active && yo != 0);
    unsigned px[4];
    predict4x4(src, bp.stride, xo, yo, job.hdr.use_bilinear_mc ? 8 : 0, need_v, px);
    if (!active) return;
    add_residual(job, mb, blk, mb.y_mode != VP8B200_SPLITMV, px);
    store4x4(job.dst + bp.off + bp.y * bp.stride + bp.x, bp.stride, px);
}


/* ---------------------------------------------------------------------------------------
 * k_inter16: macroblocks with ONE motion vector (everything except SPLITMV), the common case.
 *
 * Reference windows are read straight from global memory through L1 (four aligned words per
 * window row and strip, one funnel shift per word for the 0..3-byte offset).  Staging the
 * windows in shared memory first - which the north star asks for - was built twice on top of
 * this very arithmetic, is bit-exact both ways, and is SLOWER on this access pattern (64 x 1080p
 * P frame per launch): TMA boxes 0.80 ms (one 2-D box for luma, one 3-D box for U + V per
 * macroblock; the copy engine's per-row cost dominates for boxes this small, and a box whose x
 * is not a multiple of 16 bytes is an illegal-instruction fault, tools/probes/tma_probe.cu),
 * cp.async chunks 0.36 ms (shared-memory wavefronts at 66-76 % of peak: four macroblocks x
 * (4 luma + 4 chroma) strips at motion-vector-dependent offsets cannot be kept off each other's
 * banks), direct loads as below: see profiles/r02_summary.md.  Both staged kernels are kept in
 * profiles/experiments/ (kernels_recon_tma_k_inter16.cu.txt, kernels_recon_cpasync_staged_*).
 *
 * Warps work in pairs on eight macroblocks: the even warp owns their luma (four lanes per
 * macroblock = the four 4-pixel-wide column strips, 16 rows), the odd warp their chroma (two
 * U strips + two V strips per macroblock, 8 rows) - a chroma strip is half as tall, so mixing
 * both in one warp would leave the chroma lanes idle for half of the row loop.  A strip
 * reads four aligned words per window row and evaluates its four outputs with nine dp4a whose
 * TAP words are pre-shifted (taps slide over the bytes instead of bytes over the taps): no
 * per-output window extraction.  Sums are shifted, clamped and packed four at a time by two
 * saturating pack instructions (I2IP).  First-pass rows are kept column-wise (4x4 byte
 * transposes) so that the second pass is the same nine-dp4a pattern down each column.  A zero
 * fraction (and full-pixel motion) selects the source bytes themselves - the reference's
 * identity tap 128 does not fit a signed byte.  Residual (dequant + WHT + IDCT) is added per
 * 4x4 block right before the block's four row stores.
 * ------------------------------------------------------------------------------------- */
#define I16_WARPS 4
#define I16_MBW 8                  /* macroblocks per warp (pair: one luma warp + one chroma warp) */

/* DCs of the four luma blocks of column `col` from the second-order block (idctllm.c:140-192) */
__device__ __forceinline__ void iwalsh_col(const int16_t *y2, int dc_f, int ac_f, int col, int (&dc)[4])
{
    int q[16];
    load_coefs(y2, q);
    q[0] = s16(q[0] * dc_f);
#pragma unroll
    for (int i = 1; i < 16; i++) q[i] = s16(q[i] * ac_f);
    int mid[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int a = q[i] + q[12 + i];
        int b = q[4 + i] + q[8 + i];
        int c = q[4 + i] - q[8 + i];
        int d = q[i] - q[12 + i];
        mid[i] = s16(a + b);
        mid[4 + i] = s16(c + d);
        mid[8 + i] = s16(a - b);
        mid[12 + i] = s16(d - c);
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        int m0 = mid[4 * r], m1 = mid[4 * r + 1], m2 = mid[4 * r + 2], m3 = mid[4 * r + 3];
        int a = m0 + m3, b = m1 + m2, c = m1 - m2, d = m0 - m3;
        int o = col == 0 ? a + b : col == 1 ? c + d : col == 2 ? a - b : d - c;
        dc[r] = s16((o + 3) >> 3);
    }
}

/* residual of one block given the (optional) WHT DC; same arithmetic as add_residual */
__device__ __forceinline__ void residual_dc(const FrameJob &job, const vp8b200_mb &mb, int blk, bool luma,
                                            int dc, unsigned (&px)[4])
{
    const int16_t(*dq)[2] = job.hdr.dequant[mb.flags & VP8B200_MBF_SEGMENT_MASK];
    const unsigned mask = mb.coef_mask;
    if ((mask >> blk) & 1u) {
        int q[16];
        load_coefs(job.coef + ((size_t)mb.coef_off + __popc(mask & ((1u << blk) - 1u))) * 16, q);
        const int plane = luma ? 0 : 2;
        q[0] = luma ? dc : s16(q[0] * dq[plane][0]);      /* luma DC comes from the WHT, factor 1 */
#pragma unroll
        for (int i = 1; i < 16; i++) q[i] = s16(q[i] * dq[plane][1]);
        idct4x4_add(q, px);
    } else if (luma && dc != 0) {
        dc_add(dc, px);
    } else if (luma) {
        /* (0 + 4) >> 3 == 0: nothing to add */
    }
}

/* four s32 -> clamp to 0..255 -> one word, two saturating packs (I2IP) */
__device__ __forceinline__ unsigned pack_sat4(int a, int b, int c, int d)
{
    unsigned hi, r;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(hi));
    return r;
}

/* the nine pre-shifted tap words of one 6-tap set */
struct Taps9 { int a0, b0, a1, b1, a2, b2, a3, b3, c3; };
__device__ __forceinline__ Taps9 make_taps(int2 t)
{
    Taps9 T;
    const unsigned A = (unsigned)t.x, B = (unsigned)t.y;
    T.a0 = (int)A;          T.b0 = (int)B;
    T.a1 = (int)(A << 8);   T.b1 = (int)__funnelshift_r(A, B, 24);
    T.a2 = (int)(A << 16);  T.b2 = (int)__funnelshift_r(A, B, 16);
    T.a3 = (int)(A << 24);  T.b3 = (int)__funnelshift_r(A, B, 8);
    T.c3 = (int)(B >> 8);
    return T;
}
/* four filtered outputs from 12 consecutive bytes (w0 | w1 | w2), rounded and shifted, NOT yet clamped */
__device__ __forceinline__ void filt4(unsigned w0, unsigned w1, unsigned w2, const Taps9 &T, int (&o)[4])
{
    o[0] = dp4a_us(w1, T.b0, dp4a_us(w0, T.a0, 64)) >> 7;
    o[1] = dp4a_us(w1, T.b1, dp4a_us(w0, T.a1, 64)) >> 7;
    o[2] = dp4a_us(w1, T.b2, dp4a_us(w0, T.a2, 64)) >> 7;
    o[3] = dp4a_us(w2, T.c3, dp4a_us(w1, T.b3, dp4a_us(w0, T.a3, 64))) >> 7;
}
/* 4x4 byte transpose: four row words -> four column words */
__device__ __forceinline__ void transpose4(unsigned r0, unsigned r1, unsigned r2, unsigned r3, unsigned (&c)[4])
{
    const unsigned t0 = __byte_perm(r0, r1, 0x5140), t1 = __byte_perm(r2, r3, 0x5140);
    const unsigned t2 = __byte_perm(r0, r1, 0x7362), t3 = __byte_perm(r2, r3, 0x7362);
    c[0] = __byte_perm(t0, t1, 0x5410); c[1] = __byte_perm(t0, t1, 0x7632);
    c[2] = __byte_perm(t2, t3, 0x5410); c[3] = __byte_perm(t2, t3, 0x7632);
}

#ifndef I16_L1_PREFETCH
#define I16_L1_PREFETCH 2          /* iterations of look-ahead of the L1 prefetch (0: none) */
#endif
#ifndef I16_MINB
#define I16_MINB 8          /* 64 registers: 0.264 ms per 64 x 1080p against 0.302 at 96 (5 CTAs per SM) */
#endif
__global__ void __launch_bounds__(I16_WARPS * 32, I16_MINB)
k_inter16(const FrameJob *__restrict__ jobs, const Geo g)
{
    __shared__ FrameJob job;
    {
        const unsigned *s = reinterpret_cast<const unsigned *>(&jobs[blockIdx.y]);
        unsigned *d = reinterpret_cast<unsigned *>(&job);
        for (int i = threadIdx.x; i < (int)(sizeof(FrameJob) / 4); i += I16_WARPS * 32) d[i] = s[i];
    }
    __syncthreads();
    if (job.hdr.frame_type == 0) return;                       /* key frame: nothing inter */
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    /* warp role: even warps do luma, odd warps the chroma of the same eight macroblocks */
    const bool luma = (warp & 1) == 0;
    const int gi = lane >> 2, li = lane & 3;                   /* macroblock of the warp, strip */
    const int n_mb = g.mb_cols * g.mb_rows;
    const int mbi = (blockIdx.x * (I16_WARPS / 2) + (warp >> 1)) * I16_MBW + gi;
    bool active = mbi < n_mb;
    vp8b200_mb mb;
    *reinterpret_cast<uint4 *>(&mb) = active ? __ldg(reinterpret_cast<const uint4 *>(job.mb + mbi)) : make_uint4(0, 0, 0, 0);
    active = active && mb.ref_frame != VP8B200_INTRA_FRAME && mb.y_mode != VP8B200_SPLITMV;
    if (!__any_sync(FULL_MASK, active)) return;
    const int mb_row = active ? mbi / g.mb_cols : 0, mb_col = active ? mbi - mb_row * g.mb_cols : 0;
    const int plane = luma ? 0 : (li < 2 ? 1 : 2);
    const int strip = luma ? li : (li & 1);

    /* motion vector: reconinter.c:384-441 */
    MV mv;
    mv.row = active ? mb.u.mv.row : 0; mv.col = active ? mb.u.mv.col : 0;
    if (active && (mb.flags & VP8B200_MBF_CLAMP_MVS))
        mv = clamp_mv(mv, -((mb_col * 16) << 3), ((g.mb_cols - 1 - mb_col) * 16) << 3,
                      -((mb_row * 16) << 3), ((g.mb_rows - 1 - mb_row) * 16) << 3);
    if (!luma) {                                               /* reconinter.c:419-424 */
        const int fpmask = job.hdr.full_pixel ? ~7 : ~0;
        mv.row = s16(mv.row + (1 | (mv.row >> 31)));
        mv.col = s16(mv.col + (1 | (mv.col >> 31)));
        mv.row = (mv.row / 2) & fpmask;
        mv.col = (mv.col / 2) & fpmask;
    }
    const int xo = mv.col & 7, yo = mv.row & 7;

    /* ---- taps, second-order transform ---- */
    const int tb = job.hdr.use_bilinear_mc ? 8 : 0;
    const Taps9 TH = make_taps(c_tapw[tb + xo]), TV = make_taps(c_tapw[tb + yo]);
    const int c3v[4] = { TV.c3, TV.c3 << 8, TV.c3 << 16, TV.c3 << 24 };   /* tap 5 on column j of a ROW word */
    const bool skip = (mb.flags & VP8B200_MBF_SKIP) != 0;
    int dc4[4] = {0, 0, 0, 0};
    if (active && luma && !skip && (mb.coef_mask & (1u << 24))) {
        const int16_t(*dq)[2] = job.hdr.dequant[mb.flags & VP8B200_MBF_SEGMENT_MASK];
        iwalsh_col(job.coef + ((size_t)mb.coef_off + __popc(mb.coef_mask & 0xffffffu)) * 16, dq[1][0], dq[1][1], strip, dc4);
    }
    const int stride = luma ? g.y_stride : g.uv_stride;
    const int poff = luma ? g.y_off : (plane == 1 ? g.u_off : g.v_off);
    uint8_t *dstp = job.dst + poff + (size_t)(luma ? mb_row * 16 : mb_row * 8) * stride + (luma ? mb_col * 16 : mb_col * 8) + strip * 4;
    /* my strip's window in the reference plane: first byte = pixel (x - 2, y - 2) of the strip,
     * read as aligned words */
    const uint8_t *refbuf = job.ref[active ? mb.ref_frame : 1];
    const uint8_t *win = refbuf + poff + (ptrdiff_t)((luma ? mb_row * 16 : mb_row * 8) + (mv.row >> 3) - 2) * stride +
                         (luma ? mb_col * 16 : mb_col * 8) + (mv.col >> 3) - 2 + strip * 4;
    const unsigned sh = ((unsigned)(uintptr_t)win & 3u) * 8u;
    const unsigned *wrow = reinterpret_cast<const unsigned *>(win - (sh >> 3));
    const int wstride = stride >> 2;
    const int n_units = luma ? 4 : 2;

    /* first-pass row r of my strip -> one word of four clamped pixels.  (Warp-uniform short cuts
     * for integer motion - skip the pass when no lane has a fraction - were measured: the
     * branches keep the compiler from batching the 16 loads of an iteration and the kernel gets
     * slower, 0.293 against 0.263 ms; profiles/experiments/README.md.) */
    auto hrow = [&](int r) -> unsigned {
        const unsigned *p = wrow + r * wstride;
        const unsigned p0 = __ldg(p), p1 = __ldg(p + 1), p2 = __ldg(p + 2), p3 = __ldg(p + 3);
        const unsigned w0 = __funnelshift_r(p0, p1, sh), w1 = __funnelshift_r(p1, p2, sh), w2 = __funnelshift_r(p2, p3, sh);
        int o[4];
        filt4(w0, w1, w2, TH, o);
        const unsigned f = pack_sat4(o[0], o[1], o[2], o[3]);
        return xo ? f : __byte_perm(w0, w1, 0x5432);           /* zero fraction: pixels 2..5 themselves */
    };

#if I16_L1_PREFETCH
#pragma unroll
    for (int r = 5; r <= 4 + 4 * I16_L1_PREFETCH; r++) asm volatile("prefetch.global.L1 [%0];" ::"l"(wrow + r * wstride));
#endif
    unsigned ga[4], gb[4];                      /* column words of row groups u and u + 1 */
    unsigned r4, keep2, keep3;                                  /* first row of group u + 1 ; rows 2, 3 of group u (identity second pass) */
    {
        const unsigned h0 = hrow(0), h1 = hrow(1), h2 = hrow(2), h3 = hrow(3);
        transpose4(h0, h1, h2, h3, ga);
        keep2 = h2; keep3 = h3;
        r4 = hrow(4);
    }
#pragma unroll 1
    for (int u = 0; u < n_units; u++) {
        /* rows 4u+5 .. 4u+8 of the window */
#if I16_L1_PREFETCH
        /* window rows of a later iteration are requested into L1 while this one computes: the
         * kernel's largest stall is the wait for its loads (7.85 -> 7.55 ms per 30 steps) */
        if (u + I16_L1_PREFETCH < n_units) {
#pragma unroll
            for (int r = 5; r <= 8; r++) asm volatile("prefetch.global.L1 [%0];" ::"l"(wrow + (4 * (u + I16_L1_PREFETCH) + r) * wstride));
        }
#endif
        const unsigned h5 = hrow(4 * u + 5), h6 = hrow(4 * u + 6), h7 = hrow(4 * u + 7), h8 = hrow(4 * u + 8);
        transpose4(r4, h5, h6, h7, gb);
        unsigned px[4];
        {
            int o[4][4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                o[0][j] = dp4a_us(gb[j], TV.b0, dp4a_us(ga[j], TV.a0, 64)) >> 7;
                o[1][j] = dp4a_us(gb[j], TV.b1, dp4a_us(ga[j], TV.a1, 64)) >> 7;
                o[2][j] = dp4a_us(gb[j], TV.b2, dp4a_us(ga[j], TV.a2, 64)) >> 7;
                o[3][j] = dp4a_us(h8, c3v[j], dp4a_us(gb[j], TV.b3, dp4a_us(ga[j], TV.a3, 64))) >> 7;
            }
#pragma unroll
            for (int i = 0; i < 4; i++) px[i] = pack_sat4(o[i][0], o[i][1], o[i][2], o[i][3]);
            if (yo == 0) { px[0] = keep2; px[1] = keep3; px[2] = r4; px[3] = h5; }   /* rows 4u+2 .. 4u+5 */
        }
        if (active) {
            if (!skip) {
                const int blk = luma ? 4 * u + strip : (plane == 1 ? 16 : 20) + 2 * u + strip;
                const int dcu = u == 0 ? dc4[0] : u == 1 ? dc4[1] : u == 2 ? dc4[2] : dc4[3];
                residual_dc(job, mb, blk, luma, luma ? dcu : 0, px);
            }
            store4x4(dstp + (size_t)(4 * u) * stride, stride, px);
        }
        /* slide down four rows */
#pragma unroll
        for (int j = 0; j < 4; j++) ga[j] = gb[j];
        keep2 = h6; keep3 = h7; r4 = h8;
    }
}

void vp8b200_launch_inter(cudaStream_t s, const FrameJob *jobs, int n_jobs, const Geo &g, bool any_split)
{
    const int n_mb = g.mb_cols * g.mb_rows;
    const int per_cta = (I16_WARPS / 2) * I16_MBW;
    dim3 grid16((n_mb + per_cta - 1) / per_cta, n_jobs);
    k_inter16<<<grid16, I16_WARPS * 32, 0, s>>>(jobs, g);
    if (any_split) {
        dim3 grid((n_mb + WARPS_PER_CTA - 1) / WARPS_PER_CTA, n_jobs);
        k_inter_split<<<grid, WARPS_PER_CTA * 32, 0, s>>>(jobs, g);
    }
}
