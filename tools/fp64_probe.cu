// Developer probe: FP64 latency / throughput, shared-memory atomics, barriers and global atomics on the B200 at hand.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat(double *out, double a, double b, int n, long long *cyc)
{
    double x = a;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = x * b + a;     // dependent DFMA/DMUL+DADD chain
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void k_flat(float *out, float a, float b, int n, long long *cyc)
{
    float x = a;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = x * b + a;
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void k_thr(double *out, double a, double b, int n, long long *cyc)
{
    double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { x0 = x0 * b + a; x1 = x1 * b + a; x2 = x2 * b + a; x3 = x3 * b + a; x4 = x4 * b + a; x5 = x5 * b + a; x6 = x6 * b + a; x7 = x7 * b + a; }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void k_bar(int n, long long *cyc)
{
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_atoms(int n, long long *cyc, int *g)
{
    __shared__ int s[64];
    if (threadIdx.x < 64) s[threadIdx.x] = 0;
    __syncthreads();
    long long t0 = clock64();
    int v = 0;
    for (int i = 0; i < n; ++i) v += atomicExch(&s[(threadIdx.x + v) & 63], i);   // dependent shared atomics
    long long t1 = clock64();
    int w = 0;
    for (int i = 0; i < n; ++i) w += atomicExch(&g[(threadIdx.x * 97 + w) & 65535], i);   // dependent global atomics
    long long t2 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
    g[65536 + threadIdx.x] = v + w;
}
int main()
{
    double *out; float *fo; long long *cyc; int *g;
    cudaMalloc(&out, 8 * 1024 * 1024); cudaMalloc(&fo, 4 * 1024 * 1024); cudaMallocManaged(&cyc, 64); cudaMalloc(&g, 4 * 70000); cudaMemset(g, 0, 4 * 70000);
    const int n = 4096;
    k_lat<<<1, 32>>>(out, 1.0, 0.999, n, cyc); cudaDeviceSynchronize(); printf("FP64 dependent mul+add (no FMA contraction unless compiled in): %.1f cycles per iteration\n", (double)cyc[0] / n);
    k_flat<<<1, 32>>>(fo, 1.0f, 0.999f, n, cyc); cudaDeviceSynchronize(); printf("FP32 dependent: %.1f cycles per iteration\n", (double)cyc[0] / n);
    for (int nt = 32; nt <= 1024; nt *= 2) { k_thr<<<1, nt>>>(out, 1.0, 0.999, n, cyc); cudaDeviceSynchronize(); printf("FP64 throughput, %4d threads on one SM: %.2f cycles per warp-instruction pair (mul+add) [%.1f cycles / iteration of 8]\n", nt, (double)cyc[0] / n / 8 / (nt / 32.0), (double)cyc[0] / n); }
    for (int nt = 32; nt <= 1024; nt *= 4) { k_bar<<<1, nt>>>(n, cyc); cudaDeviceSynchronize(); printf("__syncthreads, %4d threads: %.1f cycles\n", nt, (double)cyc[0] / n); }
    k_atoms<<<1, 32>>>(1024, cyc, g); cudaDeviceSynchronize(); printf("dependent shared atomicExch: %.1f cycles, dependent global atomicExch: %.1f cycles\n", (double)cyc[0] / 1024, (double)cyc[1] / 1024);
    return 0;
}
