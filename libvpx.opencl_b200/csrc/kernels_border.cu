/* kernels_border.cu - whole-frame border extension after the loop filter.
 *
 * Restates vp8_yv12_extend_frame_borders (vpx_scale/generic/yv12extend.c:23-145): 32 (luma)
 * / 16 (chroma) pixels replicated left and right of every row, then the first / last
 * extended row copied into the border rows above / below, over the full stride.
 * Every border byte is a pure function of an edge pixel, so one pass suffices:
 *   out(y, x) = plane(clamp(y, 0, H-1), clamp(x, 0, W-1)).
 * Interior rows write only their two side borders: one LANE per row (a warp takes 32 rows,
 * every lane reads its row's two edge pixels and stores 2 x B bytes), so a frame needs 72 warps
 * for them instead of 2,176 one-row warps whose two dependent loads each cost a DRAM round trip
 * (35 -> ~12 us per 64 x 1080p launch).  The 2 x B border rows above and below a plane are
 * whole-stride copies: one warp per row.
 */
#include "vp8b200_dev.cuh"

#define BORDER_WARPS 4

__global__ void __launch_bounds__(BORDER_WARPS * 32)
k_border(const FrameJob *__restrict__ jobs, const Geo g)
{
    uint8_t *fb = jobs[blockIdx.y].dst;
    const int lane = threadIdx.x & 31;
    int w = blockIdx.x * BORDER_WARPS + (threadIdx.x >> 5);
    const int Hy = g.height, Hc = g.height >> 1;
    const int gy = (Hy + 31) >> 5, gc = (Hc + 31) >> 5;          /* 32-row groups of interior rows */
    const int n_groups = gy + 2 * gc;
    if (w < n_groups) {
        /* ---- side borders of 32 interior rows, lane = row ---- */
        int W, H, B, stride;
        uint8_t *base;                                  /* pixel (0,0) of the plane */
        if (w < gy) { W = g.width; H = Hy; B = 32; stride = g.y_stride; base = fb + g.y_off; }
        else {
            w -= gy;
            W = g.width >> 1; H = Hc; B = 16; stride = g.uv_stride;
            if (w < gc) base = fb + g.u_off; else { w -= gc; base = fb + g.v_off; }
        }
        const int row = w * 32 + lane;
        if (row >= H) return;
        uint8_t *dst = base + (size_t)row * stride;
        const unsigned lpix = dst[0] * 0x01010101u, rpix = dst[W - 1] * 0x01010101u;
        uint2 *l = reinterpret_cast<uint2 *>(dst - B), *r = reinterpret_cast<uint2 *>(dst + W);
        const int n2 = B >> 3;                           /* 8-byte stores per side: 4 (luma) or 2 (chroma) */
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < n2) { l[i] = make_uint2(lpix, lpix); r[i] = make_uint2(rpix, rpix); }
        return;
    }
    /* ---- border rows above / below the planes: full row over the stride, one warp per row ---- */
    w -= n_groups;
    int W, H, B, stride, row;
    uint8_t *base;
    if (w < 64) { W = g.width; H = Hy; B = 32; stride = g.y_stride; base = fb + g.y_off; row = w < 32 ? w - 32 : H + (w - 32); }
    else {
        w -= 64;
        if (w >= 64) return;
        W = g.width >> 1; H = Hc; B = 16; stride = g.uv_stride;
        base = fb + (w < 32 ? g.u_off : g.v_off);
        w &= 31;
        row = w < 16 ? w - 16 : H + (w - 16);
    }
    const int sy = row < 0 ? 0 : H - 1;
    const uint8_t *src = base + (size_t)sy * stride;
    uint8_t *dst = base + (size_t)row * stride;
    const unsigned lpix = src[0] * 0x01010101u, rpix = src[W - 1] * 0x01010101u;
    const unsigned *s4 = reinterpret_cast<const unsigned *>(src);
    unsigned *d4 = reinterpret_cast<unsigned *>(dst - B);
    const int total = stride >> 2, lw = B >> 2, iw = W >> 2;
    for (int i = lane; i < total; i += 32) {
        unsigned v;
        if (i < lw) v = lpix;
        else if (i < lw + iw) v = s4[i - lw];
        else v = rpix;
        d4[i] = v;
    }
}

void vp8b200_launch_border(cudaStream_t s, const FrameJob *jobs, int n_jobs, const Geo &g)
{
    const int warps = ((g.height + 31) >> 5) + 2 * (((g.height >> 1) + 31) >> 5) + 64 + 64;
    dim3 grid((warps + BORDER_WARPS - 1) / BORDER_WARPS, n_jobs);
    k_border<<<grid, BORDER_WARPS * 32, 0, s>>>(jobs, g);
}
