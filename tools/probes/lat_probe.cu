// lat_probe.cu - dependent-chain latencies of the warp-level operations the intra wavefront uses
// (one warp, clock64 around N dependent repetitions).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lat_probe lat_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
#define FULL 0xffffffffu
__device__ __forceinline__ int lds_u8(unsigned a) { int v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_u8(unsigned a, int v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint2 lds_v2(unsigned a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ int clamp255(int v) { return min(max(v, 0), 255); }
__global__ void probe(long long *out, int seed)
{
    __shared__ uint2 s_pre[10][32];
    __shared__ unsigned char s_tile[17 * 48];
    __shared__ unsigned sm[64];
    __shared__ unsigned char sb[1024];
    const int lane = threadIdx.x;
    sm[lane] = (lane + 1) & 31; sm[lane + 32] = lane;
    for (int i = lane; i < 1024; i += 32) sb[i] = (unsigned char)((i * 7 + seed) & 31);
    __syncwarp();
    long long t0, t1; unsigned v = lane + seed, a;
    // 0: integer add chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) v = v * 3 + 1;
    t1 = clock64(); if (lane == 0) out[0] = t1 - t0; out[16 + lane] = v;
    // 1: shuffle chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) v = __shfl_sync(FULL, v, (v + lane) & 31);
    t1 = clock64(); if (lane == 0) out[1] = t1 - t0; out[16 + lane] += v;
    // 2: redux chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) v = __reduce_add_sync(FULL, v & 0xff) + lane;
    t1 = clock64(); if (lane == 0) out[2] = t1 - t0; out[16 + lane] += v;
    // 3: LDS.U8 pointer chase
    a = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) a = sb[a * 32 + lane];
    t1 = clock64(); if (lane == 0) out[3] = t1 - t0; out[16 + lane] += a;
    // 4: STS -> __syncwarp -> LDS of another lane's value
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { sm[lane] = v; __syncwarp(); v = sm[(lane + 1) & 31] + 1; __syncwarp(); }
    t1 = clock64(); if (lane == 0) out[4] = t1 - t0; out[16 + lane] += v;
    // 5: STS.U8 -> __syncwarp -> LDS.U8 (one barrier per round, like the B_PRED step)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { sb[(i & 1) * 64 + lane] = (unsigned char)v; __syncwarp(); v = sb[(i & 1) * 64 + ((lane + 1) & 31)] + 1; }
    t1 = clock64(); if (lane == 0) out[5] = t1 - t0; out[16 + lane] += v;
    // 6: __syncwarp alone
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { __syncwarp(); }
    t1 = clock64(); if (lane == 0) out[6] = t1 - t0;
    // 7: vote chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) v = __ballot_sync(FULL, (v + i) & 1) + lane;
    t1 = clock64(); if (lane == 0) out[7] = t1 - t0; out[16 + lane] += v;
    // 8: replica of one B_PRED step: LDS.U8 -> 3 SHFL + REDUX -> 6 ALU -> STS.U8 -> syncwarp
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) {
        const int e = sb[(i & 1) * 64 + lane];
        const int ea = __shfl_sync(FULL, e, (lane + 1) & 31), eb = __shfl_sync(FULL, e, (lane + 2) & 31), ec = __shfl_sync(FULL, e, (lane + 5) & 31);
        const unsigned s = __reduce_add_sync(FULL, (unsigned)e);
        int x = (ea + 2 * eb + ec + 2) >> 2;
        x = min(max(x, 0), 255);
        x = (lane & 7) == 3 ? (int)((s + 4) >> 3) : x;
        x = min(max(x + (int)(lane - 16), 0), 255);
        sb[((i + 1) & 1) * 64 + lane] = (unsigned char)x;
        __syncwarp();
    }
    t1 = clock64(); if (lane == 0) out[8] = t1 - t0;
    // 10: the real B_PRED loop of kernels_intra.cu on a synthetic table (10 steps per round)
    {
        for (int st = 0; st < 10; st++) {
            const int pix = threadIdx.x & 15, which = threadIdx.x >> 4;
            const int br = (st > 3 ? (st - 2) >> 1 : 0) + which, bc = st - 2 * br;
            const bool act = br <= 3 && bc >= 0 && bc <= 3;
            const int b_off = act ? br * 4 * 48 + bc * 4 : 0;
            const int e = min(pix, 12);
            const int e_off = e < 4 ? (3 - e) * 48 - 1 : e - 5 - 48;
            const unsigned ent = ((pix + st) % 13) | (((pix + 3) % 13) << 4) | (((pix + 7) % 13) << 8) | (((st + pix) % 5 == 0 ? 2u : ((st + pix) % 7 == 0 ? 3u : 0u)) << 12);
            s_pre[st][threadIdx.x] = make_uint2(ent | ((unsigned)(unsigned short)(pix - 8) << 16),
                                                (unsigned)((b_off + e_off) & 0xffff) | ((unsigned)(act ? b_off + (pix >> 2) * 48 + (pix & 3) : 0xffff) << 16));
        }
        for (int i = lane; i < 17 * 48; i += 32) s_tile[i] = (unsigned char)(i * 3 + seed);
        __syncwarp();
        const bool e_dc = (lane & 15) < 4 || ((lane & 15) >= 5 && (lane & 15) < 9);
        const int half = lane & 16;
        const unsigned yt_s = (unsigned)__cvta_generic_to_shared(s_tile + 48 + 16);
        t0 = clock64();
        for (int round = 0; round < 32; round++) {
            unsigned pre_s = (unsigned)__cvta_generic_to_shared(&s_pre[0][lane]);
            uint2 t = lds_v2(pre_s);
#pragma unroll 1
            for (int step = 0; step < 10; step++) {
                const int edge = lds_u8(yt_s + (short)(t.y & 0xffff));
                const int la = half + (t.x & 15), lb3 = half + ((t.x >> 4) & 15), lc = half + ((t.x >> 8) & 15);
                const int kind = (t.x >> 12) & 3;
                const int wb = kind == 3 ? -1 : 2, rs = kind == 3 ? 0 : 2;
                const int res_px = (short)(t.x >> 16);
                const unsigned st = t.y >> 16;
                if (step < 9) pre_s += 32 * sizeof(uint2);
                t = lds_v2(pre_s);
                const int ea = __shfl_sync(FULL, edge, la);
                const int eb = __shfl_sync(FULL, edge, lb3);
                const int ec = __shfl_sync(FULL, edge, lc);
                int x = clamp255((ea + wb * eb + ec + rs) >> rs);
                {
                    const unsigned sums = __reduce_add_sync(FULL, e_dc ? (unsigned)edge << half : 0u);
                    const int dcv = (int)(((sums >> half) & 0xffffu) + 4) >> 3;
                    x = kind == 2 ? dcv : x;
                }
                x = clamp255(x + res_px);
                if (st != 0xffff) sts_u8(yt_s + st, x);
                __syncwarp();
            }
        }
        t1 = clock64(); if (lane == 0) out[10] = (t1 - t0) * N / 320;
    }
    // 9: same without the reduction
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) {
        const int e = sb[(i & 1) * 64 + lane];
        const int ea = __shfl_sync(FULL, e, (lane + 1) & 31), eb = __shfl_sync(FULL, e, (lane + 2) & 31), ec = __shfl_sync(FULL, e, (lane + 5) & 31);
        int x = (ea + 2 * eb + ec + 2) >> 2;
        x = min(max(x, 0), 255);
        x = min(max(x + (int)(lane - 16), 0), 255);
        sb[((i + 1) & 1) * 64 + lane] = (unsigned char)x;
        __syncwarp();
    }
    t1 = clock64(); if (lane == 0) out[9] = t1 - t0;
}
int main()
{
    long long *d, h[48];
    cudaMalloc(&d, sizeof h);
    for (int rep = 0; rep < 3; rep++) { probe<<<1, 32>>>(d, rep); cudaDeviceSynchronize(); }
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    const char *names[] = { "imad chain", "shfl chain", "redux chain", "lds.u8 chase", "sts+sync+lds+sync", "sts.u8+sync+lds.u8", "syncwarp", "ballot chain", "bpred step replica", "replica without redux", "real B_PRED loop, per step" };
    for (int i = 0; i < 11; i++) printf("%-24s %.1f cycles per round\n", names[i], (double)h[i] / N);
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
