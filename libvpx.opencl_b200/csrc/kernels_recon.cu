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
        for (int i = threadIdx.x; i < (int)(sizeof(FrameJob) / 4); i += blockDim.x) d[i] = s[i];
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
 * Six lanes per macroblock, five macroblocks per warp: lanes 0-3 of a group own the four
 * 4-pixel-wide luma column strips (16 rows), lanes 4-5 own the two chroma strips and do U
 * then V (8 rows each).  A strip is walked top to bottom in units of four rows with a
 * sliding 9-row window of first-pass results, so the horizontal pass runs over h+5 rows
 * exactly once (21 for luma; 13 + 13 for chroma) instead of 9 rows per 4x4 block, and both
 * passes use dp4a on byte-packed operands.  Residual (dequant + WHT + IDCT) is added per
 * 4x4 block right before the block's four 32-bit row stores.
 * ------------------------------------------------------------------------------------- */
#define I16_WARPS 4
#define I16_MB_PER_WARP 5

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

__global__ void __launch_bounds__(I16_WARPS * 32)
k_inter16(const FrameJob *__restrict__ jobs, const Geo g)
{
    __shared__ FrameJob job;
    {
        const unsigned *s = reinterpret_cast<const unsigned *>(&jobs[blockIdx.y]);
        unsigned *d = reinterpret_cast<unsigned *>(&job);
        for (int i = threadIdx.x; i < (int)(sizeof(FrameJob) / 4); i += blockDim.x) d[i] = s[i];
    }
    __syncthreads();
    if (job.hdr.frame_type == 0) return;                       /* key frame: nothing inter */
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gi = lane / 6, li = lane - gi * 6;               /* group (macroblock) and role */
    const int n_mb = g.mb_cols * g.mb_rows;
    const int mbi = (blockIdx.x * I16_WARPS + warp) * I16_MB_PER_WARP + gi;
    bool active = gi < I16_MB_PER_WARP && mbi < n_mb;
    vp8b200_mb mb;
    *reinterpret_cast<uint4 *>(&mb) = active ? __ldg(reinterpret_cast<const uint4 *>(job.mb + mbi)) : make_uint4(0, 0, 0, 0);
    active = active && mb.ref_frame != VP8B200_INTRA_FRAME && mb.y_mode != VP8B200_SPLITMV;
    if (!__any_sync(FULL_MASK, active)) return;
    const int mb_row = active ? mbi / g.mb_cols : 0, mb_col = active ? mbi - mb_row * g.mb_cols : 0;
    const bool luma = li < 4;
    const int strip = luma ? li : li - 4;

    /* motion vector: reconinter.c:384-441 */
    MV mv;
    mv.row = active ? mb.u.mv.row : 0; mv.col = active ? mb.u.mv.col : 0;   /* idle lanes read (0,0) */
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
    const int tb = job.hdr.use_bilinear_mc ? 8 : 0;
    const int2 th = c_taps[tb + xo], tv = c_taps[tb + yo];
    const bool need_v = __any_sync(FULL_MASK, active && yo != 0);

    const int stride = luma ? g.y_stride : g.uv_stride;
    const int wstride = stride >> 2;
    const int px0 = (luma ? mb_col * 16 : mb_col * 8) + strip * 4;
    const int py0 = luma ? mb_row * 16 : mb_row * 8;
    const uint8_t *refbuf = job.ref[active ? mb.ref_frame : 1];
    /* second-order transform once per luma lane */
    const bool skip = (mb.flags & VP8B200_MBF_SKIP) != 0;
    int dc4[4] = {0, 0, 0, 0};
    if (active && luma && !skip && (mb.coef_mask & (1u << 24))) {
        const int16_t(*dq)[2] = job.hdr.dequant[mb.flags & VP8B200_MBF_SEGMENT_MASK];
        iwalsh_col(job.coef + ((size_t)mb.coef_off + __popc(mb.coef_mask & 0xffffffu)) * 16, dq[1][0], dq[1][1], strip, dc4);
    }

    unsigned c0[4], c1[4], c2[4];                              /* 9-row window, one column per index */
    const unsigned *wp = nullptr;                              /* word holding pixel (x-2) of row 0 of the plane strip */
    unsigned sh = 0;
    uint8_t *dstp = nullptr;

    /* one first-pass row into window slot W (compile-time) */
#define HROW(ROWREL, W)                                                                         \
    {                                                                                           \
        const unsigned *row_ = wp + (ROWREL) * wstride;                                         \
        const unsigned w0_ = __ldg(row_), w1_ = __ldg(row_ + 1), w2_ = __ldg(row_ + 2);         \
        const unsigned v0_ = __funnelshift_r(w0_, w1_, sh), v1_ = __funnelshift_r(w1_, w2_, sh), v2_ = w2_ >> sh; \
        _Pragma("unroll") for (int j_ = 0; j_ < 4; j_++) {                                      \
            const unsigned f_ = (unsigned)filt6(WIN_LO(v0_, v1_, j_), WIN_MID(v0_, v1_, v2_, j_), th); \
            if ((W) < 4) c0[j_] = __byte_perm(c0[j_], f_, (W) == 0 ? 0x3214 : (W) == 1 ? 0x3240 : (W) == 2 ? 0x3410 : 0x4210); \
            else if ((W) < 8) c1[j_] = __byte_perm(c1[j_], f_, (W) == 4 ? 0x3214 : (W) == 5 ? 0x3240 : (W) == 6 ? 0x3410 : 0x4210); \
            else c2[j_] = f_;                                                                   \
        }                                                                                       \
    }

#pragma unroll 1
    for (int u = 0; u < 4; u++) {
        const bool start = u == 0 || (!luma && u == 2);        /* first unit of a plane strip */
        const int ur = luma ? u : (u & 1);                     /* unit index inside the plane */
        if (start) {
            const int poff = luma ? g.y_off : (u < 2 ? g.u_off : g.v_off);
            const uint8_t *p = refbuf + poff + (size_t)(py0 + (mv.row >> 3)) * stride + px0 + (mv.col >> 3) - 2;
            sh = ((unsigned)(uintptr_t)p & 3u) * 8u;
            wp = reinterpret_cast<const unsigned *>(p - (sh >> 3));
            dstp = job.dst + poff + (size_t)py0 * stride + px0;
        }
        unsigned px[4];
        if (need_v) {
            if (start) {
#pragma unroll
                for (int j = 0; j < 4; j++) { c0[j] = 0; c1[j] = 0; c2[j] = 0; }
                HROW(-2, 0) HROW(-1, 1) HROW(0, 2) HROW(1, 3) HROW(2, 4)
            }
            const unsigned *wsave = wp;
            wp += (4 * ur) * wstride;
            HROW(3, 5) HROW(4, 6) HROW(5, 7) HROW(6, 8)
            wp = wsave;
            int o[4][4];
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    o[r][j] = filt6(WIN_LO(c0[j], c1[j], r), WIN_MID(c0[j], c1[j], c2[j], r), tv);
                }
#pragma unroll
            for (int r = 0; r < 4; r++) px[r] = pack4(o[r][0], o[r][1], o[r][2], o[r][3]);
            /* slide the window down four rows */
#pragma unroll
            for (int j = 0; j < 4; j++) { c0[j] = c1[j]; c1[j] = c2[j]; }
        } else {
            /* no lane has a vertical fraction: the second pass is the identity */
            const unsigned *wsave = wp;
            wp += (4 * ur) * wstride;
            HROW(0, 0) HROW(1, 1) HROW(2, 2) HROW(3, 3)
            wp = wsave;
            /* c0[j] holds column j: transpose to rows */
#pragma unroll
            for (int r = 0; r < 4; r++)
                px[r] = pack4((c0[0] >> (8 * r)) & 255, (c0[1] >> (8 * r)) & 255, (c0[2] >> (8 * r)) & 255, (c0[3] >> (8 * r)) & 255);
        }
        if (active) {
            if (!skip) {
                const int blk = luma ? 4 * u + strip : (u < 2 ? 16 : 20) + 2 * ur + strip;
                const int dcu = u == 0 ? dc4[0] : u == 1 ? dc4[1] : u == 2 ? dc4[2] : dc4[3];
                residual_dc(job, mb, blk, luma, luma ? dcu : 0, px);
            }
            store4x4(dstp + (size_t)(4 * ur) * stride, stride, px);
        }
    }
#undef HROW
}

void vp8b200_launch_inter(cudaStream_t s, const FrameJob *jobs, int n_jobs, const Geo &g, bool any_split)
{
    const int n_mb = g.mb_cols * g.mb_rows;
    const int per_cta = I16_WARPS * I16_MB_PER_WARP;
    dim3 grid16((n_mb + per_cta - 1) / per_cta, n_jobs);
    k_inter16<<<grid16, I16_WARPS * 32, 0, s>>>(jobs, g);
    if (any_split) {
        dim3 grid((n_mb + WARPS_PER_CTA - 1) / WARPS_PER_CTA, n_jobs);
        k_inter_split<<<grid, WARPS_PER_CTA * 32, 0, s>>>(jobs, g);
    }
}
