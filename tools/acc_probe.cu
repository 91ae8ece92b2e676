// Developer probe: cost per element of a sequential float sum fed from shared memory (the refinement's accumulators in k_rho).
#include <cstdio>
#include <cuda_runtime.h>
#define T2 256
__global__ void __launch_bounds__(512) k_acc(int ntiles, long long *cyc, float *out, int mode)
{
    extern __shared__ float s[];      // 36 x (T2 + 1)
    for (int i = threadIdx.x; i < 36 * (T2 + 1); i += 512) s[i] = 1e-3f * (float)(i % 97);
    __syncthreads();
    float acc = 0.0f;
    long long t0 = clock64();
    if (threadIdx.x < 36) {
        const float *row = s + threadIdx.x * (T2 + 1);
        for (int t = 0; t < ntiles; ++t) {
            if (mode == 0) {
                for (int q = 0; q + 8 <= T2; q += 8) {
                    const float v0 = row[q], v1 = row[q + 1], v2 = row[q + 2], v3 = row[q + 3], v4 = row[q + 4], v5 = row[q + 5], v6 = row[q + 6], v7 = row[q + 7];
                    acc += v0; acc += v1; acc += v2; acc += v3; acc += v4; acc += v5; acc += v6; acc += v7;
                }
            } else if (mode == 1) {
                float v[16], w[16];
#pragma unroll
                for (int k = 0; k < 16; k++) v[k] = row[k];
                for (int q = 0; q + 16 <= T2; q += 16) {
#pragma unroll
                    for (int k = 0; k < 16; k++) w[k] = row[min(q + 16 + k, T2)];
#pragma unroll
                    for (int k = 0; k < 16; k++) acc += v[k];
#pragma unroll
                    for (int k = 0; k < 16; k++) v[k] = w[k];
                }
            } else {
                const float4 *r4 = reinterpret_cast<const float4 *>(s + threadIdx.x * 260);     // 16-byte aligned rows
#pragma unroll 4
                for (int q = 0; q < T2 / 4; ++q) {
                    const float4 v = r4[q];
                    acc += v.x; acc += v.y; acc += v.z; acc += v.w;
                }
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[mode] = t1 - t0;
    if (threadIdx.x < 36) out[threadIdx.x] = acc;
}
int main()
{
    long long *cyc; float *out;
    cudaMallocManaged(&cyc, 64); cudaMalloc(&out, 4096);
    cudaFuncSetAttribute(k_acc, cudaFuncAttributeMaxDynamicSharedMemorySize, 36 * 264 * 4);
    const int nt = 40;
    for (int mode = 0; mode < 3; ++mode) {
        k_acc<<<1, 512, 36 * 264 * 4>>>(nt, cyc, out, mode);
        cudaDeviceSynchronize();
        printf("mode %d: %.2f cycles per element\n", mode, (double)cyc[mode] / (nt * T2));
    }
    return 0;
}
