// Probe (GPU): one 2-D TMA box load of bytes from a pitched plane, with the tensor map passed
// (a) as a __grid_constant__ kernel parameter, (b) from global memory + tensormap proxy fence,
// (c) from global memory without the fence.  usage: tma_probe a|b|c
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <stdint.h>

__device__ __forceinline__ void mbar_init(unsigned bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile("{ .reg .pred p;\n\tW_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=; }" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const void *tmap, int x, int y, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(bar) : "memory");
}

__global__ void k_param(const __grid_constant__ CUtensorMap tm, uint8_t *out, int x, int y)
{
    __shared__ __align__(128) uint8_t tile[24 * 32];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) mbar_init(b, 1);
    __syncwarp();
    if (threadIdx.x == 0) { mbar_expect_tx(b, 768); tma_load_2d((unsigned)__cvta_generic_to_shared(tile), &tm, x, y, b); }
    mbar_wait(b, 0);
    for (int i = threadIdx.x; i < 768; i += 32) out[i] = tile[i];
}
__global__ void k_global(const void *tm, uint8_t *out, int x, int y, int fence)
{
    __shared__ __align__(128) uint8_t tile[24 * 32];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) mbar_init(b, 1);
    __syncwarp();
    if (threadIdx.x == 0) {
        if (fence) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tm) : "memory");
        mbar_expect_tx(b, 768);
        tma_load_2d((unsigned)__cvta_generic_to_shared(tile), tm, x, y, b);
    }
    mbar_wait(b, 0);
    for (int i = threadIdx.x; i < 768; i += 32) out[i] = tile[i];
}

int main(int argc, char **argv)
{
    const char mode = argc > 1 ? argv[1][0] : 'a';
    const int stride = 416, rows = 352;
    uint8_t *h = (uint8_t *)malloc(stride * rows), *d, *dout, hout[768];
    for (int i = 0; i < stride * rows; i++) h[i] = (uint8_t)(i * 7 + (i >> 8));
    cudaMalloc(&d, stride * rows); cudaMemcpy(d, h, stride * rows, cudaMemcpyHostToDevice);
    cudaMalloc(&dout, 768);
    PFN_cuTensorMapEncodeTiled_v12000 encode = NULL;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &q);
    CUtensorMap tm;
    const cuuint64_t dims[2] = { (cuuint64_t)stride, (cuuint64_t)rows };
    const cuuint64_t strides[1] = { (cuuint64_t)stride };
    const cuuint32_t box[2] = { 32, 24 }, es[2] = { 1, 1 };
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    const int x = argc > 2 ? atoi(argv[2]) : 37, y = argc > 3 ? atoi(argv[3]) : 11;
    if (mode == 'a') k_param<<<1, 32>>>(tm, dout, x, y);
    else {
        void *dtm; cudaMalloc(&dtm, 128 * 4); cudaMemcpy((char *)dtm + 128, &tm, 128, cudaMemcpyHostToDevice);
        k_global<<<1, 32>>>((char *)dtm + 128, dout, x, y, mode == 'b');
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %c: %s\n", mode, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        cudaMemcpy(hout, dout, 768, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int rr = 0; rr < 24; rr++) for (int c = 0; c < 32; c++) bad += hout[rr * 32 + c] != h[(y + rr) * stride + x + c];
        printf("mode %c: %d mismatching bytes of 768 (box at unaligned x=%d)\n", mode, bad, x);
    }
    return 0;
}
