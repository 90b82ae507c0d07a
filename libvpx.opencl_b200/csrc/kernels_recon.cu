/* kernels_recon.cu - prediction + residual.
 *
 *   k_inter : every inter macroblock of the frame, fully parallel (one warp per MB).
 *             Restates vp8_build_inter_predictors_mb (vp8/common/reconinter.c:560-573) with
 *             the sub-pixel filters of vp8/common/filter.c, fused with the residual
 *             (dequant + WHT + IDCT + add, decodframe.c:252-304).
 *   k_intra : the intra macroblocks, in a macroblock wavefront (one warp per MB row, row r
 *             may process column c once row r-1 has finished column c+1), because intra
 *             prediction reads the unfiltered reconstruction of the left / above /
 *             above-right neighbours (reconintra.c, reconintra4x4.c, decodframe.c:192-238).
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
    /* 128 does not fit a signed byte: the identity / full-weight tap is applied as -128 and
     * the sum negated (see filt6), so store tap values t with the 128 case flagged by
     * keeping -128 in the byte. */
    for (int i = 0; i < 8; i++) {
        h[i].x = pack_s8(six[i][0], six[i][1], six[i][2] == 128 ? -128 : six[i][2], six[i][3]);
        h[i].y = pack_s8(six[i][4], six[i][5], 0, 0);
        h[8 + i].x = pack_s8(0, 0, bil[i][0] == 128 ? -128 : bil[i][0], bil[i][1]);
        h[8 + i].y = 0;
    }
    cudaMemcpyToSymbol(c_taps, h, sizeof h);
}

/* One six-tap evaluation on 8 consecutive bytes lo (p0..p3) / hi (p4..p7).
 * Fraction 0 is the only filter with a 128 tap; its packed byte is -128, so the dot product
 * comes out as -128*p2 and is negated.  Result: clamp((sum + 64) >> 7). */
__device__ __forceinline__ int filt6(unsigned lo, unsigned hi, int2 t, bool identity)
{
    int s = dp4a_us(lo, t.x, 0);
    s = dp4a_us(hi, t.y, s);
    s = identity ? -s : s;
    return clamp255((s + 64) >> 7);
}

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
    const bool idh = xo == 0, idv = yo == 0;
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
            unsigned lo = j ? __funnelshift_r(v0, v1, 8 * j) : v0;
            unsigned hi = j ? __funnelshift_r(v1, v2, 8 * j) : v1;
            unsigned f = (unsigned)filt6(lo, hi, th, idh);
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
            unsigned lo = r ? __funnelshift_r(c0[j], c1[j], 8 * r) : c0[j];
            unsigned hi = r ? __funnelshift_r(c1[j], c2[j], 8 * r) : c1[j];
            o[r][j] = filt6(lo, hi, tv, idv);
        }
#pragma unroll
    for (int r = 0; r < 4; r++) px[r] = pack4(o[r][0], o[r][1], o[r][2], o[r][3]);
}

__device__ __forceinline__ void store4x4(uint8_t *dst, int stride, const unsigned (&px)[4])
{
#pragma unroll
    for (int r = 0; r < 4; r++) *reinterpret_cast<unsigned *>(dst + r * stride) = px[r];
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

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_inter(const FrameJob *__restrict__ jobs, const Geo g)
{
    __shared__ FrameJob job;
    {
        const unsigned *s = reinterpret_cast<const unsigned *>(&jobs[blockIdx.y]);
        unsigned *d = reinterpret_cast<unsigned *>(&job);
        for (int i = threadIdx.x; i < (int)(sizeof(FrameJob) / 4); i += blockDim.x) d[i] = s[i];
    }
    __syncthreads();
    if (job.hdr.frame_type == 0) return;                       /* key frame: nothing inter */
    const int lane = threadIdx.x & 31;
    const int mbi = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    if (mbi >= g.mb_cols * g.mb_rows) return;
    vp8b200_mb mb;
    *reinterpret_cast<uint4 *>(&mb) = __ldg(reinterpret_cast<const uint4 *>(job.mb + mbi));
    if (mb.ref_frame == VP8B200_INTRA_FRAME) return;           /* k_intra's */
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
 * intra wavefront
 * ------------------------------------------------------------------------------------- */
#define INTRA_ROWS_PER_CTA 4
#define YT_STRIDE 32      /* luma tile: rows -1..15, cols -4..27 ; index (r+1)*32 + c+4   */
#define CT_STRIDE 16      /* chroma tile: rows -1..7, cols -4..11 ; index (r+1)*16 + c+4  */

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}


/* reconintra4x4.c:16-296.  A[0] = top-left, A[1..8] = above + above-right, L[0..3] = left. */
__device__ __forceinline__ void intra4x4(int mode, const int (&A)[9], const int (&L)[4], unsigned (&px)[4])
{
    int o[4][4];
#define AVG3(x, y, z) (((x) + 2 * (y) + (z) + 2) >> 2)
#define AVG2(x, y) (((x) + (y) + 1) >> 1)
    /* edge array E: L3 L2 L1 L0 tl A0 A1 A2 A3 */
    int E[9] = {L[3], L[2], L[1], L[0], A[0], A[1], A[2], A[3], A[4]};
    const int *a = &A[1];
    switch (mode) {
    case VP8B200_B_DC_PRED: {
        int s = (a[0] + a[1] + a[2] + a[3] + L[0] + L[1] + L[2] + L[3] + 4) >> 3;
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) o[r][c] = s;
        break;
    }
    case VP8B200_B_TM_PRED:
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) o[r][c] = clamp255(a[c] - A[0] + L[r]);
        break;
    case VP8B200_B_VE_PRED:
#pragma unroll
        for (int c = 0; c < 4; c++) {
            int v = AVG3(A[c], A[c + 1], A[c + 2]);
#pragma unroll
            for (int r = 0; r < 4; r++) o[r][c] = v;
        }
        break;
    case VP8B200_B_HE_PRED: {
        int v[4] = {AVG3(A[0], L[0], L[1]), AVG3(L[0], L[1], L[2]), AVG3(L[1], L[2], L[3]), AVG3(L[2], L[3], L[3])};
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) o[r][c] = v[r];
        break;
    }
    case VP8B200_B_LD_PRED: {
        int d[7];
#pragma unroll
        for (int k = 0; k < 6; k++) d[k] = AVG3(a[k], a[k + 1], a[k + 2]);
        d[6] = AVG3(a[6], a[7], a[7]);
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) o[r][c] = d[r + c];
        break;
    }
    case VP8B200_B_RD_PRED: {
        int d[7];
#pragma unroll
        for (int k = 0; k < 7; k++) d[k] = AVG3(E[k], E[k + 1], E[k + 2]);
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) o[r][c] = d[3 - r + c];
        break;
    }
    case VP8B200_B_VR_PRED:
        o[3][0] = AVG3(E[1], E[2], E[3]);
        o[2][0] = AVG3(E[2], E[3], E[4]);
        o[3][1] = o[1][0] = AVG3(E[3], E[4], E[5]);
        o[2][1] = o[0][0] = AVG2(E[4], E[5]);
        o[3][2] = o[1][1] = AVG3(E[4], E[5], E[6]);
        o[2][2] = o[0][1] = AVG2(E[5], E[6]);
        o[3][3] = o[1][2] = AVG3(E[5], E[6], E[7]);
        o[2][3] = o[0][2] = AVG2(E[6], E[7]);
        o[1][3] = AVG3(E[6], E[7], E[8]);
        o[0][3] = AVG2(E[7], E[8]);
        break;
    case VP8B200_B_VL_PRED:
        o[0][0] = AVG2(a[0], a[1]);
        o[1][0] = AVG3(a[0], a[1], a[2]);
        o[2][0] = o[0][1] = AVG2(a[1], a[2]);
        o[1][1] = o[3][0] = AVG3(a[1], a[2], a[3]);
        o[2][1] = o[0][2] = AVG2(a[2], a[3]);
        o[3][1] = o[1][2] = AVG3(a[2], a[3], a[4]);
        o[0][3] = o[2][2] = AVG2(a[3], a[4]);
        o[1][3] = o[3][2] = AVG3(a[3], a[4], a[5]);
        o[2][3] = AVG3(a[4], a[5], a[6]);
        o[3][3] = AVG3(a[5], a[6], a[7]);
        break;
    case VP8B200_B_HD_PRED:
        o[3][0] = AVG2(E[0], E[1]);
        o[3][1] = AVG3(E[0], E[1], E[2]);
        o[2][0] = o[3][2] = AVG2(E[1], E[2]);
        o[2][1] = o[3][3] = AVG3(E[1], E[2], E[3]);
        o[2][2] = o[1][0] = AVG2(E[2], E[3]);
        o[2][3] = o[1][1] = AVG3(E[2], E[3], E[4]);
        o[1][2] = o[0][0] = AVG2(E[3], E[4]);
        o[1][3] = o[0][1] = AVG3(E[3], E[4], E[5]);
        o[0][2] = AVG3(E[4], E[5], E[6]);
        o[0][3] = AVG3(E[5], E[6], E[7]);
        break;
    default: /* VP8B200_B_HU_PRED */
        o[0][0] = AVG2(L[0], L[1]);
        o[0][1] = AVG3(L[0], L[1], L[2]);
        o[0][2] = o[1][0] = AVG2(L[1], L[2]);
        o[0][3] = o[1][1] = AVG3(L[1], L[2], L[3]);
        o[1][2] = o[2][0] = AVG2(L[2], L[3]);
        o[1][3] = o[2][1] = AVG3(L[2], L[3], L[3]);
        o[2][2] = o[2][3] = o[3][0] = o[3][1] = o[3][2] = o[3][3] = L[3];
        break;
    }
#undef AVG3
#undef AVG2
#pragma unroll
    for (int r = 0; r < 4; r++) px[r] = pack4(o[r][0], o[r][1], o[r][2], o[r][3]);
}

/* 16x16 luma / 8x8 chroma whole-block modes (reconintra.c:139-263, :403-546) for the lane's
 * 4x4 sub-block at (bx, by) inside a block of `size`; T = tile with the borders loaded,
 * addressed T[(r+1)*ts + c+4]. */
__device__ __forceinline__ void intra_block_mode(int mode, const uint8_t *T, int ts, int size,
                                                 int bx, int by, bool up, bool left, unsigned (&px)[4])
{
    int o[4][4];
    const uint8_t *above = T + 4;                      /* row -1, col 0 */
    switch (mode) {
    case VP8B200_DC_PRED: {
        int dc = 128;
        if (up || left) {
            int sum = 0, shift = (size == 16 ? 3 : 2) + (up ? 1 : 0) + (left ? 1 : 0);
            if (up) for (int c = 0; c < size; c++) sum += above[c];
            if (left) for (int r = 0; r < size; r++) sum += T[(r + 1) * ts + 3];
            dc = (sum + (1 << (shift - 1))) >> shift;
        }
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) o[r][c] = dc;
        break;
    }
    case VP8B200_V_PRED:
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) o[r][c] = above[bx + c];
        break;
    case VP8B200_H_PRED:
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) o[r][c] = T[(by + r + 1) * ts + 3];
        break;
    default: { /* TM_PRED */
        int tl = T[3];
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) o[r][c] = clamp255(T[(by + r + 1) * ts + 3] + above[bx + c] - tl);
        break;
    }
    }
#pragma unroll
    for (int r = 0; r < 4; r++) px[r] = pack4(o[r][0], o[r][1], o[r][2], o[r][3]);
}

__global__ void __launch_bounds__(INTRA_ROWS_PER_CTA * 32)
k_intra(const FrameJob *__restrict__ jobs, const int n_jobs, const Geo g,
        unsigned *ticket, const unsigned ticket_base)
{
    __shared__ FrameJob job;
    __shared__ unsigned s_ticket;
    __shared__ __align__(16) uint8_t s_yt[INTRA_ROWS_PER_CTA][17 * YT_STRIDE];
    __shared__ __align__(16) uint8_t s_ct[INTRA_ROWS_PER_CTA][2][9 * CT_STRIDE];
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u) - ticket_base;
    __syncthreads();
    const unsigned t = s_ticket;
    const int ji = t % n_jobs, group = t / n_jobs;
    {
        const unsigned *s = reinterpret_cast<const unsigned *>(&jobs[ji]);
        unsigned *d = reinterpret_cast<unsigned *>(&job);
        for (int i = threadIdx.x; i < (int)(sizeof(FrameJob) / 4); i += blockDim.x) d[i] = s[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb_row = group * INTRA_ROWS_PER_CTA + warp;
    if (mb_row >= g.mb_rows) return;
    const unsigned base = job.epoch_intra << VP8B200_EPOCH_SHIFT;
    unsigned *my_prog = job.progress + mb_row;
    if (job.n_intra == 0) {                              /* nothing to do in this frame */
        if (lane == 0) st_release(my_prog, base + g.mb_cols);
        return;
    }
    const unsigned *up_prog = job.progress + mb_row - 1;
    uint8_t *YT = s_yt[warp];
    uint8_t *UT = s_ct[warp][0], *VT = s_ct[warp][1];
    uint8_t *const dy = job.dst + g.y_off, *const du = job.dst + g.u_off, *const dv = job.dst + g.v_off;
    const bool up = mb_row != 0;
    unsigned seen = base;                                /* last value read from up_prog */

    for (int c0 = 0; c0 < g.mb_cols; c0 += 32) {
        /* scan 32 macroblock records at a time for intra ones */
        int mbi = mb_row * g.mb_cols + c0 + lane;
        bool is_intra = false;
        if (c0 + lane < g.mb_cols) is_intra = job.mb[mbi].ref_frame == VP8B200_INTRA_FRAME;
        unsigned todo = __ballot_sync(FULL_MASK, is_intra);
        while (todo) {
            const int k = __ffs(todo) - 1;
            todo &= todo - 1;
            const int mb_col = c0 + k;
            vp8b200_mb mb;
            *reinterpret_cast<uint4 *>(&mb) =
                *reinterpret_cast<const uint4 *>(job.mb + mb_row * g.mb_cols + mb_col);
            const bool left = mb_col != 0;
            /* everything left of this macroblock is final (inter MBs were done by k_inter) */
            if (lane == 0) st_release(my_prog, base + mb_col);
            /* dependency: row above finished column mb_col+1 (above-right), clamped to the row end */
            if (up) {
                const unsigned need = base + (unsigned)min(mb_col + 2, g.mb_cols);
                if ((int)(seen - need) < 0) {
                    if (lane == 0) {
                        unsigned v = ld_acquire(up_prog);
                        while ((int)(v - need) < 0) { __nanosleep(40); v = ld_acquire(up_prog); }
                        seen = v;
                    }
                    seen = __shfl_sync(FULL_MASK, seen, 0);
                    __syncwarp();                  /* order every lane's loads after the acquire */
                }
            }
            __syncwarp();
            /* ---- borders into the tiles (setupintrarecon.c:15-32 rules at frame edges) ---- */
            {
                const uint8_t *row_above = dy + (mb_row * 16 - 1) * g.y_stride + mb_col * 16;
                if (lane < 21) {                              /* cols -1..19 of row -1 */
                    int c = lane - 1, v;
                    if (!up) v = 127;
                    else if (c < 0) v = left ? __ldcg(row_above - 1) : 129;
                    else if (c >= 16 && mb_col == g.mb_cols - 1) v = __ldcg(row_above + 15);  /* extend.c:160-185 */
                    else v = __ldcg(row_above + c);
                    YT[c + 4] = (uint8_t)v;
                }
                if (lane < 16) {                              /* col -1 of rows 0..15 */
                    int v = left ? __ldcg(dy + (mb_row * 16 + lane) * g.y_stride + mb_col * 16 - 1) : 129;
                    YT[(lane + 1) * YT_STRIDE + 3] = (uint8_t)v;
                }
                /* chroma: lanes 0..8 row -1 (cols -1..7), lanes 9..16 col -1, for U; 16.. for V */
                {
                    const int half = lane >> 4;               /* 0 = U, 1 = V */
                    const int q = lane & 15;
                    uint8_t *CT = half ? VT : UT;
                    const uint8_t *pl = half ? dv : du;
                    const uint8_t *ca = pl + (mb_row * 8 - 1) * g.uv_stride + mb_col * 8;
                    if (q < 9) {
                        int c = q - 1, v;
                        if (!up) v = 127;
                        else if (c < 0) v = left ? __ldcg(ca - 1) : 129;
                        else v = __ldcg(ca + c);
                        CT[c + 4] = (uint8_t)v;
                    }
                    if (q < 8) {
                        int v = left ? __ldcg(pl + (mb_row * 8 + q) * g.uv_stride + mb_col * 8 - 1) : 129;
                        CT[(q + 1) * CT_STRIDE + 3] = (uint8_t)v;
                    }
                }
            }
            __syncwarp();
            /* ---- chroma (lanes 16..23) and whole-block luma (lanes 0..15) ---- */
            const bool bpred = mb.y_mode == VP8B200_B_PRED;
            if (lane >= 16 && lane < 24) {
                const int j = lane & 3, bx = (j & 1) * 4, by = (j >> 1) * 4;
                unsigned px[4];
                intra_block_mode(mb.uv_mode, lane < 20 ? UT : VT, CT_STRIDE, 8, bx, by, up, left, px);
                add_residual(job, mb, lane, false, px);
                store4x4((lane < 20 ? du : dv) + (mb_row * 8 + by) * g.uv_stride + mb_col * 8 + bx, g.uv_stride, px);
            } else if (lane < 16 && !bpred) {
                const int bx = (lane & 3) * 4, by = (lane >> 2) * 4;
                unsigned px[4];
                intra_block_mode(mb.y_mode, YT, YT_STRIDE, 16, bx, by, up, left, px);
                add_residual(job, mb, lane, true, px);
                store4x4(dy + (mb_row * 16 + by) * g.y_stride + mb_col * 16 + bx, g.y_stride, px);
            }
            if (bpred) {
                /* 16 sub-blocks, anti-diagonal wavefront: block (br,bc) at step bc + 2*br
                 * (needs left, above, above-right); decodframe.c:200-237 */
                int bmode = 0;
                if (lane < 16) bmode = reinterpret_cast<const uint8_t *>(job.aux + mb.u.aux)[lane];
                const int bc = lane & 3, br = lane >> 2;
                for (int step = 0; step < 10; step++) {
                    if (lane < 16 && bc + 2 * br == step) {
                        int A[9], L[4];
                        const uint8_t *arow = YT + (br * 4) * YT_STRIDE + bc * 4 + 4;   /* row br*4-1 */
#pragma unroll
                        for (int i = 0; i < 5; i++) A[i] = arow[i - 1];
                        /* above-right: column 3 always takes row -1 of the MB (the reference's
                         * down-copy, reconintra4x4.c:305-317) */
                        const uint8_t *ar = bc == 3 ? YT + 16 + 4 : arow + 4;
#pragma unroll
                        for (int i = 0; i < 4; i++) A[5 + i] = ar[i];
#pragma unroll
                        for (int i = 0; i < 4; i++) L[i] = YT[(br * 4 + i + 1) * YT_STRIDE + bc * 4 + 3];
                        unsigned px[4];
                        intra4x4(bmode, A, L, px);
                        add_residual(job, mb, lane, false, px);
#pragma unroll
                        for (int r = 0; r < 4; r++)
                            *reinterpret_cast<unsigned *>(YT + (br * 4 + r + 1) * YT_STRIDE + bc * 4 + 4) = px[r];
                        store4x4(dy + (mb_row * 16 + br * 4) * g.y_stride + mb_col * 16 + bc * 4, g.y_stride, px);
                    }
                    __syncwarp();
                }
            }
            /* publish: everything up to and including this column of the row is final
             * (warp barrier, then one cumulative release by lane 0) */
            __syncwarp();
            if (lane == 0) st_release(my_prog, base + mb_col + 1);
        }
    }
    __syncwarp();
    if (lane == 0) st_release(my_prog, base + g.mb_cols);
}

void vp8b200_launch_inter(cudaStream_t s, const FrameJob *jobs, int n_jobs, const Geo &g)
{
    dim3 grid((g.mb_cols * g.mb_rows + WARPS_PER_CTA - 1) / WARPS_PER_CTA, n_jobs);
    k_inter<<<grid, WARPS_PER_CTA * 32, 0, s>>>(jobs, g);
}

void vp8b200_launch_intra(cudaStream_t s, const FrameJob *jobs, int n_jobs, const Geo &g,
                          unsigned int *ticket, unsigned int ticket_base, int *n_ctas)
{
    int groups = (g.mb_rows + INTRA_ROWS_PER_CTA - 1) / INTRA_ROWS_PER_CTA;
    *n_ctas = groups * n_jobs;
    k_intra<<<groups * n_jobs, INTRA_ROWS_PER_CTA * 32, 0, s>>>(jobs, n_jobs, g, ticket, ticket_base);
}
