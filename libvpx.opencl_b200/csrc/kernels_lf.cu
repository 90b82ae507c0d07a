/* kernels_lf.cu - in-loop deblocking filter, normal and simple variants.
 *
 * Restates vp8_loop_filter_frame (vp8/common/loopfilter.c:203-316) with the edge filters of
 * vp8/common/loopfilter_filters.c, level selection of vp8_loop_filter_frame_init
 * (loopfilter.c:117-201) and the limit tables of vp8_loop_filter_update_sharpness (:66-96).
 *
 * Schedule: the reference filters macroblocks in raster order and each macroblock reads
 * pixels its left, above and above-right neighbours have already modified.  One warp owns
 * one macroblock ROW and walks it left to right; row r may filter column c once row r-1 has
 * finished column c+1 (per-row progress counters in global memory, release/acquire).
 *
 * Inside a macroblock the warp first filters the four vertical edges with lane = pixel row
 * (lanes 0-15 luma rows, 16-23 U rows, 24-31 V rows; the row lives in registers, the 4
 * pixels left of the MB are carried over from the previous column), transposes through a
 * 512-byte shared-memory tile, filters the four horizontal edges with lane = pixel column,
 * and transposes back.  The last 4 columns of a macroblock are stored one iteration later,
 * after the next macroblock's left-edge filter has modified them.
 */
#include "vp8b200_dev.cuh"

#define LF_ROWS_PER_CTA 4

__device__ __forceinline__ unsigned lf_ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void lf_st_release(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ int sc(int v) { return max(min(v, 127), -128); }

struct LfParams { int ilim, blim, mblim, thr; };

/* loopfilter_filters.c:27-49 on ints */
__device__ __forceinline__ bool lf_mask(int p3, int p2, int p1, int p0, int q0, int q1, int q2, int q3,
                                        int ilim, int elim)
{
    int m = max(max(abs(p3 - p2), abs(p2 - p1)), max(abs(p1 - p0), abs(q1 - q0)));
    m = max(m, max(abs(q2 - q1), abs(q3 - q2)));
    return m <= ilim && abs(p0 - q0) * 2 + (abs(p1 - q1) >> 1) <= elim;
}

/* inner edge: loopfilter_filters.c:51-97 */
__device__ __forceinline__ void lf_inner(int p3, int p2, int &p1, int &p0, int &q0, int &q1, int q2, int q3,
                                         const LfParams &P)
{
    if (!lf_mask(p3, p2, p1, p0, q0, q1, q2, q3, P.ilim, P.blim)) return;
    bool hev = abs(p1 - p0) > P.thr || abs(q1 - q0) > P.thr;
    int ps1 = p1 - 128, ps0 = p0 - 128, qs0 = q0 - 128, qs1 = q1 - 128;
    int f = hev ? sc(ps1 - qs1) : 0;
    f = sc(f + 3 * (qs0 - ps0));
    int f1 = sc(f + 4) >> 3, f2 = sc(f + 3) >> 3;
    q0 = sc(qs0 - f1) + 128;
    p0 = sc(ps0 + f2) + 128;
    int u = hev ? 0 : (f1 + 1) >> 1;
    q1 = sc(qs1 - u) + 128;
    p1 = sc(ps1 + u) + 128;
}

/* macroblock edge: loopfilter_filters.c:161-214 */
__device__ __forceinline__ void lf_mbedge(int p3, int &p2, int &p1, int &p0, int &q0, int &q1, int &q2, int q3,
                                          const LfParams &P)
{
    if (!lf_mask(p3, p2, p1, p0, q0, q1, q2, q3, P.ilim, P.mblim)) return;
    bool hev = abs(p1 - p0) > P.thr || abs(q1 - q0) > P.thr;
    int ps2 = p2 - 128, ps1 = p1 - 128, ps0 = p0 - 128, qs0 = q0 - 128, qs1 = q1 - 128, qs2 = q2 - 128;
    int f = sc(sc(ps1 - qs1) + 3 * (qs0 - ps0));
    int g = hev ? f : 0;
    int f1 = sc(g + 4) >> 3, f2 = sc(g + 3) >> 3;
    qs0 = sc(qs0 - f1);
    ps0 = sc(ps0 + f2);
    if (hev) f = 0;
    int u = sc((63 + f * 27) >> 7);
    q0 = sc(qs0 - u) + 128;
    p0 = sc(ps0 + u) + 128;
    u = sc((63 + f * 18) >> 7);
    q1 = sc(qs1 - u) + 128;
    p1 = sc(ps1 + u) + 128;
    u = sc((63 + f * 9) >> 7);
    q2 = sc(qs2 - u) + 128;
    p2 = sc(ps2 + u) + 128;
}

/* simple filter: loopfilter_filters.c:281-315 */
__device__ __forceinline__ void lf_simple(int p1, int &p0, int &q0, int q1, int blim)
{
    if (abs(p0 - q0) * 2 + (abs(p1 - q1) >> 1) > blim) return;
    int f = sc(sc((p1 - 128) - (q1 - 128)) + 3 * (q0 - p0));
    int f1 = sc(f + 4) >> 3, f2 = sc(f + 3) >> 3;
    q0 = sc(q0 - 128 - f1) + 128;
    p0 = sc(p0 - 128 + f2) + 128;
}

__device__ __forceinline__ void unpack(unsigned w, int &a, int &b, int &c, int &d)
{
    a = w & 255; b = (w >> 8) & 255; c = (w >> 16) & 255; d = w >> 24;
}

/* filter the vertical edge between packed words l (4 px left) and r (4 px right) */
__device__ __forceinline__ void vedge(unsigned &l, unsigned &r, bool mbedge, bool simple, const LfParams &P)
{
    int p3, p2, p1, p0, q0, q1, q2, q3;
    unpack(l, p3, p2, p1, p0);
    unpack(r, q0, q1, q2, q3);
    if (simple) lf_simple(p1, p0, q0, q1, mbedge ? P.mblim : P.blim);
    else if (mbedge) lf_mbedge(p3, p2, p1, p0, q0, q1, q2, q3, P);
    else lf_inner(p3, p2, p1, p0, q0, q1, q2, q3, P);
    l = pack4(p3, p2, p1, p0);
    r = pack4(q0, q1, q2, q3);
}

/* filter the horizontal edge between v[i-4..i-1] and v[i..i+3] of a pixel column */
template <int N>
__device__ __forceinline__ void hedge(int (&v)[N], int i, bool mbedge, bool simple, const LfParams &P)
{
    if (simple) lf_simple(v[i - 2], v[i - 1], v[i], v[i + 1], mbedge ? P.mblim : P.blim);
    else if (mbedge) lf_mbedge(v[i - 4], v[i - 3], v[i - 2], v[i - 1], v[i], v[i + 1], v[i + 2], v[i + 3], P);
    else lf_inner(v[i - 4], v[i - 3], v[i - 2], v[i - 1], v[i], v[i + 1], v[i + 2], v[i + 3], P);
}

__global__ void __launch_bounds__(LF_ROWS_PER_CTA * 32)
k_loopfilter(const FrameJob *__restrict__ jobs, const int n_jobs, const Geo g,
             unsigned *ticket, const unsigned ticket_base)
{
    __shared__ FrameJob job;
    __shared__ unsigned s_ticket;
    __shared__ uint8_t s_lvl[64];                          /* [seg][ref][mode class] */
    __shared__ __align__(16) uint8_t s_tile[LF_ROWS_PER_CTA][512];
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
    const unsigned base = job.epoch_lf << VP8B200_EPOCH_SHIFT;
    unsigned *my_prog = job.progress + g.mb_rows + mb_row;
    const unsigned *up_prog = my_prog - 1;
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

    const unsigned *mbrec = reinterpret_cast<const unsigned *>(job.mb + (size_t)mb_row * g.mb_cols);
    unsigned cur[4] = {0, 0, 0, 0}, nxt[4] = {0, 0, 0, 0}, halo = 0;
    if (lane_on) {
        if (luma) { uint4 v = *reinterpret_cast<const uint4 *>(rowp); cur[0] = v.x; cur[1] = v.y; cur[2] = v.z; cur[3] = v.w; }
        else { uint2 v = *reinterpret_cast<const uint2 *>(rowp); cur[0] = v.x; cur[1] = v.y; }
    }
    unsigned rec = mbrec[0], rec_n = 0;
    unsigned seen = base;                                  /* last value read from up_prog */

    for (int c = 0; c < g.mb_cols; c++) {
        /* prefetch the next macroblock's rows and record */
        if (c + 1 < g.mb_cols) {
            rec_n = mbrec[(c + 1) * 4];
            if (lane_on) {
                if (luma) { uint4 v = *reinterpret_cast<const uint4 *>(rowp + (c + 1) * 16); nxt[0] = v.x; nxt[1] = v.y; nxt[2] = v.z; nxt[3] = v.w; }
                else { uint2 v = *reinterpret_cast<const uint2 *>(rowp + (c + 1) * 8); nxt[0] = v.x; nxt[1] = v.y; }
            }
        }
        /* per-MB decisions, loopfilter.c:245-253 */
        const int y_mode = rec & 255, ref = (rec >> 16) & 255, flags = rec >> 24;
        const bool skip_lf = y_mode != VP8B200_B_PRED && y_mode != VP8B200_SPLITMV && (flags & VP8B200_MBF_SKIP);
        /* mode_lf_lut, loopfilter.c:52-63: DC,V,H,TM,ZEROMV -> 1 ; B_PRED -> 0 ; NEAREST,NEAR,NEW -> 2 ; SPLIT -> 3 */
        const int mclass = y_mode == VP8B200_B_PRED ? 0 : y_mode == VP8B200_SPLITMV ? 3
                         : (y_mode <= VP8B200_TM_PRED || y_mode == VP8B200_ZEROMV) ? 1 : 2;
        const int level = s_lvl[((flags & 3) << 4) | (ref << 2) | mclass];
        uint8_t *colp = rowp + c * mbw;                     /* my row at this MB's x = 0 */

        if (level == 0) {
            /* untouched macroblock: pass the pixels through */
            if (lane_on) {
                if (c > 0) *reinterpret_cast<unsigned *>(colp - 4) = halo;
                if (luma) { *reinterpret_cast<unsigned *>(colp) = cur[0]; *reinterpret_cast<unsigned *>(colp + 4) = cur[1];
                            *reinterpret_cast<unsigned *>(colp + 8) = cur[2]; halo = cur[3]; }
                else { *reinterpret_cast<unsigned *>(colp) = cur[0]; halo = cur[1]; }
            }
        } else {
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
            /* ---- vertical edges, lane = pixel row ---- */
            if (lane_on) {
                if (c > 0) vedge(halo, cur[0], true, simple, P);
                if (!skip_lf) {
                    vedge(cur[0], cur[1], false, simple, P);
                    if (luma) { vedge(cur[1], cur[2], false, simple, P); vedge(cur[2], cur[3], false, simple, P); }
                }
                if (c > 0) *reinterpret_cast<unsigned *>(colp - 4) = halo;
                /* rows into the tile (rows 4.. of the tile = rows 0.. of the MB) */
                if (luma) *reinterpret_cast<uint4 *>(tile + (pi + 4) * 16) = make_uint4(cur[0], cur[1], cur[2], cur[3]);
                else *reinterpret_cast<uint2 *>(tile + (pi + 4) * 8) = make_uint2(cur[0], cur[1]);
            }
            /* ---- the 4 rows above, once the row above is far enough ---- */
            if (top) {
                /* row above must have finished iteration c+1 (which stores the last 4
                 * columns of its MB c); for the last column that is its end-of-row flush,
                 * published as mb_cols+1.  The counter is re-read only when the value seen
                 * last time is not enough (the row above is usually several MBs ahead);
                 * lane 0 polls, the acquire is extended to the warp by __syncwarp. */
                const unsigned need = base + (unsigned)(c + 2);
                if ((int)(seen - need) < 0) {
                    if (lane == 0) {
                        unsigned v = lf_ld_acquire(up_prog);
                        while ((int)(v - need) < 0) { __nanosleep(40); v = lf_ld_acquire(up_prog); }
                        seen = v;
                    }
                    seen = __shfl_sync(FULL_MASK, seen, 0);
                    __syncwarp();                  /* order every lane's loads after the acquire */
                }
                if (lane_on && pi < 4) {
                    const uint8_t *ap = plane + (size_t)(mb_row * mbw - 4 + pi) * stride + c * mbw;
                    if (luma) *reinterpret_cast<uint4 *>(tile + pi * 16) = __ldcg(reinterpret_cast<const uint4 *>(ap));
                    else *reinterpret_cast<uint2 *>(tile + pi * 8) = __ldcg(reinterpret_cast<const uint2 *>(ap));
                }
            }
            __syncwarp();
            /* ---- horizontal edges, lane = pixel column ---- */
            if (lane_on) {
                if (luma) {
                    int v[20];
#pragma unroll
                    for (int r = 0; r < 20; r++) v[r] = (r >= 4 || top) ? tile[r * 16 + pi] : 0;
                    if (top) hedge(v, 4, true, simple, P);
                    if (!skip_lf) { hedge(v, 8, false, simple, P); hedge(v, 12, false, simple, P); hedge(v, 16, false, simple, P); }
#pragma unroll
                    for (int r = 1; r < 20; r++) if (r >= 4 || top) tile[r * 16 + pi] = (uint8_t)v[r];
                } else {
                    int v[12];
#pragma unroll
                    for (int r = 0; r < 12; r++) v[r] = (r >= 4 || top) ? tile[r * 8 + pi] : 0;
                    if (top) hedge(v, 4, true, simple, P);
                    if (!skip_lf) hedge(v, 8, false, simple, P);
#pragma unroll
                    for (int r = 1; r < 12; r++) if (r >= 4 || top) tile[r * 8 + pi] = (uint8_t)v[r];
                }
            }
            __syncwarp();
            /* ---- rows back out; the last word waits for the next MB's left edge ---- */
            if (lane_on) {
                if (luma) {
                    uint4 v = *reinterpret_cast<const uint4 *>(tile + (pi + 4) * 16);
                    *reinterpret_cast<unsigned *>(colp) = v.x; *reinterpret_cast<unsigned *>(colp + 4) = v.y;
                    *reinterpret_cast<unsigned *>(colp + 8) = v.z; halo = v.w;
                } else {
                    uint2 v = *reinterpret_cast<const uint2 *>(tile + (pi + 4) * 8);
                    *reinterpret_cast<unsigned *>(colp) = v.x; halo = v.y;
                }
                if (top && pi >= 1 && pi < 4) {             /* rows -3..-1 of the MB above */
                    uint8_t *ap = plane + (size_t)(mb_row * mbw - 4 + pi) * stride + c * mbw;
                    if (luma) *reinterpret_cast<uint4 *>(ap) = *reinterpret_cast<const uint4 *>(tile + pi * 16);
                    else *reinterpret_cast<uint2 *>(ap) = *reinterpret_cast<const uint2 *>(tile + pi * 8);
                }
            }
            __syncwarp();
        }
        /* publish: columns < c are final for the row below once c+1 is published.  One
         * release by lane 0 after the warp barrier covers every lane's stores (the barrier
         * orders them before the release; release is cumulative). */
        __syncwarp();
        if (lane == 0) lf_st_release(my_prog, base + c + 1);
        rec = rec_n;
#pragma unroll
        for (int i = 0; i < 4; i++) cur[i] = nxt[i];
    }
    /* last 4 columns of the row */
    if (lane_on) *reinterpret_cast<unsigned *>(rowp + g.mb_cols * mbw - 4) = halo;
    __syncwarp();
    if (lane == 0) lf_st_release(my_prog, base + g.mb_cols + 1);
}

void vp8b200_launch_loopfilter(cudaStream_t s, const FrameJob *jobs, int n_jobs, const Geo &g,
                               unsigned int *ticket, unsigned int ticket_base, int *n_ctas)
{
    int groups = (g.mb_rows + LF_ROWS_PER_CTA - 1) / LF_ROWS_PER_CTA;
    *n_ctas = groups * n_jobs;
    k_loopfilter<<<groups * n_jobs, LF_ROWS_PER_CTA * 32, 0, s>>>(jobs, n_jobs, g, ticket, ticket_base);
}
