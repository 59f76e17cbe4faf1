// FP64 / FP32 dependent-chain latency and per-SM throughput on the GPU at hand (B200: is the FP64 pipe "half rate"?).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int ILP>
__global__ void chain(T *out, long long *clk, T a, T b, int iters) {
    T x[ILP];
    for (int k = 0; k < ILP; ++k) x[k] = a + (T)k + (T)threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = x[k] * b + a;   // contracts to FMA
    }
    long long t1 = clock64();
    T s = 0;
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
__global__ void divchain(double *out, long long *clk, double a, double b, int iters) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = b / x + a;
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
__global__ void sqrtchain(double *out, long long *clk, double a, double b, int iters) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = sqrt(x) + a;
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <typename T, int ILP>
void run(const char *name, int blocks, int threads) {
    T *out; long long *clk, h;
    cudaMalloc(&out, sizeof(T) * blocks * threads); cudaMalloc(&clk, 8);
    int iters = 4096;
    chain<T, ILP><<<blocks, threads>>>(out, clk, (T)1.0000001, (T)0.9999999, iters);
    chain<T, ILP><<<blocks, threads>>>(out, clk, (T)1.0000001, (T)0.9999999, iters);
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / iters;
    printf("%-6s ILP %d  blocks %4d x %4d thr: %7.2f clk per iteration  (%.2f clk per FMA per warp; %.1f FMA lanes/clk/SM)\n", name, ILP, blocks, threads,
           per, per / ILP, (double)ILP * threads / per);
    cudaFree(out); cudaFree(clk);
}
int main() {
    run<double, 1>("f64", 1, 32);
    run<double, 2>("f64", 1, 32);
    run<double, 4>("f64", 1, 32);
    run<double, 8>("f64", 1, 32);
    run<double, 8>("f64", 148, 128);
    run<double, 8>("f64", 148, 512);
    run<double, 8>("f64", 148, 1024);
    run<float, 1>("f32", 1, 32);
    run<float, 8>("f32", 1, 32);
    run<float, 8>("f32", 148, 512);
    run<float, 8>("f32", 148, 1024);
    double *out; long long *clk, h;
    cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&clk, 8);
    for (int thr : {32, 512}) {
        divchain<<<thr == 32 ? 1 : 148, thr>>>(out, clk, 1.5, 2.5, 1024);
        cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        printf("f64 div+add chain, %d thr/block: %.1f clk per iteration\n", thr, (double)h / 1024);
        sqrtchain<<<thr == 32 ? 1 : 148, thr>>>(out, clk, 1.5, 2.5, 1024);
        cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        printf("f64 sqrt+add chain, %d thr/block: %.1f clk per iteration\n", thr, (double)h / 1024);
    }
    return 0;
}
