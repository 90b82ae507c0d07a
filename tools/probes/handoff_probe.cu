// handoff_probe.cu - latency of a tagged-word hand-off between two CTAs on different SMs:
// the producer posts word i (st.relaxed.gpu) every `gap` ns and records %globaltimer, the consumer
// polls word i (ld.relaxed.gpu) and records when it sees the tag.
#include <cstdio>
#include <cuda_runtime.h>
#define N 64
__device__ unsigned long long now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void k(unsigned long long *msg, unsigned long long *t_post, unsigned long long *t_seen, unsigned tag, int gap_ns, int lines_apart, int pollers)
{
    if (blockIdx.x == 0) {                  // producer, one thread
        if (threadIdx.x == 0) {
            unsigned long long t = now();
            for (int i = 0; i < N; i++) {
                while (now() - t < (unsigned long long)gap_ns) { }
                t = now();
                unsigned long long v = ((unsigned long long)tag << 32) | (unsigned)i;
                asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(msg + (size_t)i * lines_apart), "l"(v) : "memory");
                t_post[i] = now();
            }
        }
    } else if (blockIdx.x == gridDim.x - 1) {   // consumer: `pollers` lanes poll the same word
        if (threadIdx.x < pollers) {
            for (int i = 0; i < N; i++) {
                unsigned long long v;
                do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(msg + (size_t)i * lines_apart) : "memory"); } while ((unsigned)(v >> 32) != tag);
                if (threadIdx.x == 0) t_seen[i] = now();
            }
        }
    } else {
        // filler CTAs so that producer and consumer sit on different SMs
        unsigned long long t = now(); while (now() - t < 2000) { }
    }
}
int main()
{
    unsigned long long *msg, *tp, *ts, hp[N], hs[N];
    cudaMalloc(&msg, N * 16 * 128); cudaMalloc(&tp, sizeof hp); cudaMalloc(&ts, sizeof hs);
    cudaMemset(msg, 0, N * 16 * 128);
    unsigned tag = 1;
    int gaps[] = { 500, 2000, 10000 };
    for (int g = 0; g < 3; g++)
        for (int pollers = 1; pollers <= 6; pollers += 5) {
            k<<<148, 32>>>(msg, tp, ts, tag++, gaps[g], 16, pollers);
            cudaDeviceSynchronize();
            cudaMemcpy(hp, tp, sizeof hp, cudaMemcpyDeviceToHost); cudaMemcpy(hs, ts, sizeof hs, cudaMemcpyDeviceToHost);
            double sum = 0, mx = 0; for (int i = 8; i < N; i++) { double d = (double)hs[i] - (double)hp[i]; sum += d; if (d > mx) mx = d; }
            printf("gap %5d ns, %d polling lanes: hand-off mean %.0f ns, max %.0f ns\n", gaps[g], pollers, sum / (N - 8), mx);
        }
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
