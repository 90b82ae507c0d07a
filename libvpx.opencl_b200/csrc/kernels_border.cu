/* kernels_border.cu - whole-frame border extension after the loop filter.
 *
 * Restates vp8_yv12_extend_frame_borders (vpx_scale/generic/yv12extend.c:23-145): 32 (luma)
 * / 16 (chroma) pixels replicated left and right of every row, then the first / last
 * extended row copied into the border rows above / below, over the full stride.
 * Every border byte is a pure function of an edge pixel, so one pass suffices:
 *   out(y, x) = plane(clamp(y, 0, H-1), clamp(x, 0, W-1)).
 * One warp per allocated row; interior rows write only their two side borders.
 */
#include "vp8b200_dev.cuh"

#define BORDER_WARPS 8

__global__ void __launch_bounds__(BORDER_WARPS * 32)
k_border(const FrameJob *__restrict__ jobs, const Geo g)
{
    uint8_t *fb = jobs[blockIdx.y].dst;
    const int lane = threadIdx.x & 31;
    int row = blockIdx.x * BORDER_WARPS + (threadIdx.x >> 5);
    const int y_rows = g.height + 64, c_rows = g.uv_rows_alloc;
    if (row >= y_rows + 2 * c_rows) return;
    int W, H, B, stride;
    uint8_t *base;                                  /* pixel (0,0) of the plane */
    if (row < y_rows) { W = g.width; H = g.height; B = 32; stride = g.y_stride; base = fb + g.y_off; row -= 32; }
    else {
        row -= y_rows;
        W = g.width >> 1; H = g.height >> 1; B = 16; stride = g.uv_stride;
        if (row < c_rows) base = fb + g.u_off; else { row -= c_rows; base = fb + g.v_off; }
        row -= 16;
    }
    const int sy = min(max(row, 0), H - 1);
    const uint8_t *src = base + (size_t)sy * stride;
    uint8_t *dst = base + (size_t)row * stride;
    const unsigned lpix = src[0] * 0x01010101u, rpix = src[W - 1] * 0x01010101u;
    if (row >= 0 && row < H) {
        /* side borders only: B bytes each = B/4 words (8 or 4) */
        const int nwords = B >> 2;
        if (lane < nwords) reinterpret_cast<unsigned *>(dst - B)[lane] = lpix;
        else if (lane < 2 * nwords) reinterpret_cast<unsigned *>(dst + W)[lane - nwords] = rpix;
    } else {
        /* full row over the stride: [-B, stride - B) */
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
}

void vp8b200_launch_border(cudaStream_t s, const FrameJob *jobs, int n_jobs, const Geo &g)
{
    int rows = g.height + 64 + 2 * g.uv_rows_alloc;
    dim3 grid((rows + BORDER_WARPS - 1) / BORDER_WARPS, n_jobs);
    k_border<<<grid, BORDER_WARPS * 32, 0, s>>>(jobs, g);
}
