/* vp8b200_dev.cuh - device helpers shared by the reconstruction kernels.
 *
 * Work decomposition used by both prediction kernels: one warp owns one macroblock and
 * lane L < 24 owns one 4x4 block - L 0..15 the luma blocks in raster order, 16..19 the U
 * blocks, 20..23 the V blocks (the reference's BLOCKD numbering, vp8/common/blockd.h:232).
 * A lane keeps its 16 pixels packed in four 32-bit registers, one per row.
 */
#ifndef VP8B200_DEV_CUH
#define VP8B200_DEV_CUH

#include "vp8b200_internal.h"

#define FULL_MASK 0xffffffffu

__device__ __forceinline__ int clamp255(int v) { return __vimin_s32_relu(v, 255); }

__device__ __forceinline__ unsigned pack4(int a, int b, int c, int d)
{
    return (unsigned)a | ((unsigned)b << 8) | ((unsigned)c << 16) | ((unsigned)d << 24);
}

/* unsigned bytes of a x signed bytes of b, accumulated into c */
__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

/* ---------------------------------------------------------------------------------------
 * A1-A3 residual: dequant (int16 wrap), IDCT (first pass stored in int16), WHT
 * vp8/common/dequantize.c:17-43, idctllm.c:28-204, idct_blk.c:20-89
 * ------------------------------------------------------------------------------------- */

__device__ __forceinline__ int s16(int v) { return (int)(short)v; }

/* in: 16 dequantised coefficients (already wrapped to int16); pred: 4 packed rows.
 * Returns the reconstructed 4 rows. */
__device__ __forceinline__ void idct4x4_add(const int (&in)[16], unsigned (&px)[4])
{
    int mid[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int a = in[i] + in[8 + i];
        int b = in[i] - in[8 + i];
        int t1 = (in[4 + i] * 35468) >> 16;
        int t2 = in[12 + i] + ((in[12 + i] * 20091) >> 16);
        int c = t1 - t2;
        t1 = in[4 + i] + ((in[4 + i] * 20091) >> 16);
        t2 = (in[12 + i] * 35468) >> 16;
        int d = t1 + t2;
        mid[i]      = s16(a + d);
        mid[12 + i] = s16(a - d);
        mid[4 + i]  = s16(b + c);
        mid[8 + i]  = s16(b - c);
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        int m0 = mid[4 * r], m1 = mid[4 * r + 1], m2 = mid[4 * r + 2], m3 = mid[4 * r + 3];
        int a = m0 + m2;
        int b = m0 - m2;
        int t1 = (m1 * 35468) >> 16;
        int t2 = m3 + ((m3 * 20091) >> 16);
        int c = t1 - t2;
        t1 = m1 + ((m1 * 20091) >> 16);
        t2 = (m3 * 35468) >> 16;
        int d = t1 + t2;
        int o0 = s16((a + d + 4) >> 3);
        int o3 = s16((a - d + 4) >> 3);
        int o1 = s16((b + c + 4) >> 3);
        int o2 = s16((b - c + 4) >> 3);
        unsigned p = px[r];
        px[r] = pack4(clamp255(o0 + (int)(p & 255)), clamp255(o1 + (int)((p >> 8) & 255)),
                      clamp255(o2 + (int)((p >> 16) & 255)), clamp255(o3 + (int)(p >> 24)));
    }
}

__device__ __forceinline__ void dc_add(int dc, unsigned (&px)[4])
{
    int a = (dc + 4) >> 3;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        unsigned p = px[r];
        px[r] = pack4(clamp255(a + (int)(p & 255)), clamp255(a + (int)((p >> 8) & 255)),
                      clamp255(a + (int)((p >> 16) & 255)), clamp255(a + (int)(p >> 24)));
    }
}

__device__ __forceinline__ void load_coefs(const int16_t *p, int (&q)[16])
{
    const uint4 *v = reinterpret_cast<const uint4 *>(p);
    uint4 a = __ldg(v), b = __ldg(v + 1);
    unsigned w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        q[2 * i] = (int)(short)(w[i] & 0xffff);
        q[2 * i + 1] = (int)(short)(w[i] >> 16);
    }
}

/* DC of luma block `blk` from the second-order block (idctllm.c:140-192); every luma lane
 * evaluates the rows it needs itself. */
__device__ __forceinline__ int iwalsh_dc(const int16_t *y2, int dc_f, int ac_f, int blk)
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
    int row = blk >> 2, col = blk & 3;
    int m0 = 0, m1 = 0, m2 = 0, m3 = 0;
#pragma unroll
    for (int r = 0; r < 4; r++)
        if (r == row) { m0 = mid[4 * r]; m1 = mid[4 * r + 1]; m2 = mid[4 * r + 2]; m3 = mid[4 * r + 3]; }
    int a = m0 + m3, b = m1 + m2, c = m1 - m2, d = m0 - m3;
    int o = col == 0 ? a + b : col == 1 ? c + d : col == 2 ? a - b : d - c;
    return s16((o + 3) >> 3);
}

/* Residual of the lane's block `blk` (0..23) of macroblock `mb`, added to px in place.
 * decodframe.c:252-304 for the 16x16 / inter / chroma cases; B_PRED luma calls this with
 * has_y2 = false per sub-block (decodframe.c:217-236). */
__device__ __forceinline__ void add_residual(const FrameJob &job, const vp8b200_mb &mb, int blk,
                                             bool has_y2, unsigned (&px)[4])
{
    if (mb.flags & VP8B200_MBF_SKIP) return;
    const int16_t(*dq)[2] = job.hdr.dequant[mb.flags & VP8B200_MBF_SEGMENT_MASK];
    unsigned mask = mb.coef_mask;
    bool present = (mask >> blk) & 1u;
    int plane = blk < 16 ? 0 : 2;
    int dc_f = dq[plane][0], ac_f = dq[plane][1];
    int dc = 0;
    bool luma_y2 = has_y2 && blk < 16;
    if (luma_y2) {
        if (mask & (1u << 24)) {
            const int16_t *y2 = job.coef + ((size_t)mb.coef_off + __popc(mask & 0xffffffu)) * 16;
            dc = iwalsh_dc(y2, dq[1][0], dq[1][1], blk);
        }
        dc_f = 1;                                   /* dequant_y1_dc, decodframe.c:92,291 */
    }
    if (present) {
        int q[16];
        load_coefs(job.coef + ((size_t)mb.coef_off + __popc(mask & ((1u << blk) - 1u))) * 16, q);
        if (luma_y2) q[0] = dc;                     /* written there by the WHT */
        q[0] = s16(q[0] * dc_f);
#pragma unroll
        for (int i = 1; i < 16; i++) q[i] = s16(q[i] * ac_f);
        idct4x4_add(q, px);
    } else if (luma_y2) {
        dc_add(dc, px);                             /* idct_blk.c:32-36 */
    }
}

/* 4x4 inverse DCT only (idctllm.c:28-92): 16 residual values, before the add */
__device__ __forceinline__ void idct4x4(const int (&in)[16], int (&out)[16])
{
    int mid[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int a = in[i] + in[8 + i];
        int b = in[i] - in[8 + i];
        int t1 = (in[4 + i] * 35468) >> 16;
        int t2 = in[12 + i] + ((in[12 + i] * 20091) >> 16);
        int c = t1 - t2;
        t1 = in[4 + i] + ((in[4 + i] * 20091) >> 16);
        t2 = (in[12 + i] * 35468) >> 16;
        int d = t1 + t2;
        mid[i]      = s16(a + d);
        mid[12 + i] = s16(a - d);
        mid[4 + i]  = s16(b + c);
        mid[8 + i]  = s16(b - c);
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        int m0 = mid[4 * r], m1 = mid[4 * r + 1], m2 = mid[4 * r + 2], m3 = mid[4 * r + 3];
        int a = m0 + m2;
        int b = m0 - m2;
        int t1 = (m1 * 35468) >> 16;
        int t2 = m3 + ((m3 * 20091) >> 16);
        int c = t1 - t2;
        t1 = m1 + ((m1 * 20091) >> 16);
        t2 = (m3 * 35468) >> 16;
        int d = t1 + t2;
        out[4 * r + 0] = s16((a + d + 4) >> 3);
        out[4 * r + 3] = s16((a - d + 4) >> 3);
        out[4 * r + 1] = s16((b + c + 4) >> 3);
        out[4 * r + 2] = s16((b - c + 4) >> 3);
    }
}

/* Residual of block `blk` (0..23) as 16 values (zero when the block adds nothing), computed
 * independently of the prediction: decodframe.c:252-304 / :217-236.  Returns true when any
 * value may be non-zero. */
__device__ __forceinline__ bool block_residual(const FrameJob &job, const vp8b200_mb &mb, int blk,
                                               bool has_y2, int (&res)[16])
{
#pragma unroll
    for (int i = 0; i < 16; i++) res[i] = 0;
    if (mb.flags & VP8B200_MBF_SKIP) return false;
    const int16_t(*dq)[2] = job.hdr.dequant[mb.flags & VP8B200_MBF_SEGMENT_MASK];
    const unsigned mask = mb.coef_mask;
    const bool present = (mask >> blk) & 1u;
    const int plane = blk < 16 ? 0 : 2;
    int dc_f = dq[plane][0];
    const int ac_f = dq[plane][1];
    int dc = 0;
    const bool luma_y2 = has_y2 && blk < 16;
    if (luma_y2) {
        if (mask & (1u << 24)) {
            const int16_t *y2 = job.coef + ((size_t)mb.coef_off + __popc(mask & 0xffffffu)) * 16;
            dc = iwalsh_dc(y2, dq[1][0], dq[1][1], blk);
        }
        dc_f = 1;
    }
    if (present) {
        int q[16];
        load_coefs(job.coef + ((size_t)mb.coef_off + __popc(mask & ((1u << blk) - 1u))) * 16, q);
        if (luma_y2) q[0] = dc;
        q[0] = s16(q[0] * dc_f);
#pragma unroll
        for (int i = 1; i < 16; i++) q[i] = s16(q[i] * ac_f);
        idct4x4(q, res);
        return true;
    }
    if (luma_y2) {
        const int a = (dc + 4) >> 3;                /* idct_blk.c:32-36 */
#pragma unroll
        for (int i = 0; i < 16; i++) res[i] = a;
        return a != 0;
    }
    return false;
}

__device__ __forceinline__ void add_res(unsigned (&px)[4], const int (&res)[16])
{
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const unsigned p = px[r];
        px[r] = pack4(clamp255(res[4 * r] + (int)(p & 255)), clamp255(res[4 * r + 1] + (int)((p >> 8) & 255)),
                      clamp255(res[4 * r + 2] + (int)((p >> 16) & 255)), clamp255(res[4 * r + 3] + (int)(p >> 24)));
    }
}

__device__ __forceinline__ void store4x4(uint8_t *dst, int stride, const unsigned (&px)[4])
{
#pragma unroll
    for (int r = 0; r < 4; r++) *reinterpret_cast<unsigned *>(dst + r * stride) = px[r];
}

#endif
