// fp64_lat.cu -- dependent-issue latency and per-warp throughput of DFMA / DMUL / FFMA on this GPU
// (one warp per SM sub-partition and 1..8 independent chains per thread).  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>

template <typename T, int CH>
__global__ void chain(T *out, int iters, long long *cyc)
{
    T a[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) a[c] = T(threadIdx.x + c) * T(1e-3);
    const T m = T(0.999999), b = T(1e-6);
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) a[c] = a[c] * m + b;
    }
    long long t1 = clock64();
    T s = T(0);
#pragma unroll
    for (int c = 0; c < CH; c++) s += a[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <typename T, int CH>
void run(const char *name, int threads)
{
    T *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * sizeof(T)); cudaMalloc(&cyc, 8);
    const int iters = 1 << 14;
    chain<T, CH><<<1, threads>>>(out, iters, cyc);
    chain<T, CH><<<1, threads>>>(out, iters, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%s chains=%d threads=%4d: %.2f cycles per FMA per chain-step, %.2f cycles per instruction\n", name, CH, threads,
           (double)h / iters, (double)h / iters / CH);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<double, 1>("f64", 32); run<double, 2>("f64", 32); run<double, 4>("f64", 32); run<double, 8>("f64", 32);
    run<double, 1>("f64", 128); run<double, 2>("f64", 128); run<double, 4>("f64", 128); run<double, 8>("f64", 128);
    run<double, 1>("f64", 256); run<double, 4>("f64", 256); run<double, 8>("f64", 256);
    run<double, 1>("f64", 512); run<double, 4>("f64", 512); run<double, 1>("f64", 1024); run<double, 4>("f64", 1024);
    run<float, 1>("f32", 32); run<float, 4>("f32", 32); run<float, 8>("f32", 32); run<float, 1>("f32", 256); run<float, 8>("f32", 256);
    return 0;
}
