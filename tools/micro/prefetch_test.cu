// prefetch_test.cu -- does prefetch.global.L1 / .L2 hide the latency of the next stage's loads on this GPU?
// One warp per SM walks `stages` blocks of 64 x 256 B (the solver's SoA stage footprint), doing ~W dependent DFMAs per stage.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0 none, 1 prefetch.L1 one stage ahead, 2 prefetch.L2, 3 L1 two ahead
__global__ void walk(const double *base, int stages, int work, double *out, long long *cyc, long stride)
{
    const double *p = base + (long)blockIdx.x * stride + threadIdx.x;
    double acc = threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int k = 0; k < stages; k++) {
        const double *cur = p + (long)k * 64 * 32;
        if (MODE != 0) {
            const double *nx = cur + (MODE == 3 ? 2 : 1) * 64 * 32;
#pragma unroll
            for (int e = 0; e < 64; e++) {
                if (MODE == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + e * 32));
                else asm volatile("prefetch.global.L1 [%0];" ::"l"(nx + e * 32));
            }
        }
        for (int i = 0; i < work; i++) acc = acc * 0.999999 + 1e-6;      // dependent chain: ~9 cycles each
        double s = 0;
#pragma unroll
        for (int e = 0; e < 64; e++) s += cur[e * 32];
        acc += s;
    }
    long long t1 = clock64();
    out[blockIdx.x * 32 + threadIdx.x] = acc;
    if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = t1 - t0;
}

int main()
{
    const int stages = 40, nblk = 148;
    const long stride = (long)(stages + 4) * 64 * 32;
    const size_t n = (size_t)nblk * stride * 64;                 // 64 disjoint regions so that every run starts cold
    double *buf, *out; long long *cyc, h;
    cudaMalloc(&buf, n * 8); cudaMalloc(&out, nblk * 32 * 8); cudaMalloc(&cyc, 8);
    cudaMemset(buf, 0, n * 8);
    int region = 0;
    for (int work : { 100, 400, 1600 }) {
        for (int mode = 0; mode < 4; mode++) {
            const double *b = buf + (size_t)(region++ % 64) * nblk * stride;
            switch (mode) {
            case 0: walk<0><<<nblk, 32>>>(b, stages, work, out, cyc, stride); break;
            case 1: walk<1><<<nblk, 32>>>(b, stages, work, out, cyc, stride); break;
            case 2: walk<2><<<nblk, 32>>>(b, stages, work, out, cyc, stride); break;
            default: walk<3><<<nblk, 32>>>(b, stages, work, out, cyc, stride); break;
            }
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("work %4d dependent DFMAs/stage, mode %d (%s): %.0f cycles per stage (compute alone ~%d)\n", work, mode,
                   mode == 0 ? "no prefetch" : mode == 1 ? "prefetch.L1 +1" : mode == 2 ? "prefetch.L2 +1" : "prefetch.L1 +2",
                   (double)h / stages, work * 9);
        }
    }
    return 0;
}
