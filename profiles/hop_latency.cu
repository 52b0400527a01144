// hop_latency.cu -- what does one dependency hop of the sweep kernel cost at the very least?  Two CTAs on
// different SMs play ping-pong through global memory with the same instructions the sweep uses to publish and
// poll (st.relaxed.gpu / ld.relaxed.gpu): time per one-way hop = publish -> visible in L2 -> seen by the poller.
// build + run on the GPU box:  nvcc -arch=sm_100a -O3 -o /tmp/hop profiles/hop_latency.cu && /tmp/hop
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ld_poll(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_pub(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__global__ void pingpong(unsigned long long *a, unsigned long long *b, int iters, int partner, long long *cycles, unsigned *smids)
{
    if (threadIdx.x != 0) return;
    unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
    if (blockIdx.x == 0) {
        smids[0] = smid;
        const long long t0 = clock64();
        for (int i = 1; i <= iters; ++i) { st_pub(a, i); while (ld_poll(b) != (unsigned long long)i) { } }
        *cycles = clock64() - t0;
    } else if ((int)blockIdx.x == partner) {
        smids[1] = smid;
        for (int i = 1; i <= iters; ++i) { while (ld_poll(a) != (unsigned long long)i) { } st_pub(b, i); }
    }
}
int main()
{
    unsigned long long *a, *b; long long *cyc; unsigned *sm;
    cudaMalloc(&a, 256); cudaMalloc(&b, 256); cudaMalloc(&cyc, 8); cudaMalloc(&sm, 8);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    for (int partner : {1, 2, 37, 73, 74, 100, 147}) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(a, 0, 256); cudaMemset(b, 0, 256);
            pingpong<<<148, 32>>>(a, b, iters, partner, cyc, sm);
            cudaDeviceSynchronize();
        }
        long long c; unsigned s[2];
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(s, sm, 8, cudaMemcpyDeviceToHost);
        printf("blocks 0 <-> %3d (SM %3u <-> SM %3u): %.1f cycles = %.3f us per one-way hop (SM clock %d MHz)\n", partner, s[0], s[1],
               (double)c / (2.0 * iters), (double)c / (2.0 * iters) / (clk * 1e-3), clk / 1000);
    }
    return 0;
}
