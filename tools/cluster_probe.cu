// Developer probe: cost of cluster barriers, L2 round trips of data written by another SM, and DSMEM loads on the B200 at hand.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
template <int NT> __global__ void __launch_bounds__(NT) k_csync(int n, long long *cyc, int slot)
{
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) cluster.sync();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[slot] = t1 - t0;
}
// every thread writes a word, cluster barrier, reads the word written by the same thread index of the next CTA (L2, ld.cg), dependent chain of 1
__global__ void __launch_bounds__(1024) k_xread(int n, long long *cyc, int *buf, int slot, int *sink)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int G = gridDim.x * blockDim.x, g = blockIdx.x * blockDim.x + threadIdx.x, peer = ((blockIdx.x + 1) % gridDim.x) * blockDim.x + threadIdx.x;
    int acc = 0;
    long long tw = 0, tr = 0, ts = 0;
    for (int i = 0; i < n; ++i) {
        long long t0 = clock64();
        buf[(i & 1) * G + g] = i + acc;
        long long t1 = clock64();
        cluster.sync();
        long long t2 = clock64();
        acc += __ldcg(&buf[(i & 1) * G + peer]);
        acc += __ldcg(&buf[(i & 1) * G + ((peer + acc) % G)]);      // second, dependent round trip
        long long t3 = clock64();
        tw += t1 - t0; ts += t2 - t1; tr += t3 - t2;
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[slot] = tw; cyc[slot + 1] = ts; cyc[slot + 2] = tr; }
    sink[g] = acc;
}
__global__ void __launch_bounds__(1024) k_dsmem(int n, long long *cyc, int slot, int *sink)
{
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ int s[1024];
    s[threadIdx.x] = threadIdx.x;
    cluster.sync();
    const int *remote = cluster.map_shared_rank(s, (cluster.block_rank() + 1) % cluster.num_blocks());
    int v = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) v = remote[v & 1023];     // dependent DSMEM loads
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[slot] = t1 - t0;
    sink[blockIdx.x * 1024 + threadIdx.x] = v;
    cluster.sync();
}
template <class K, class... A> static void launch(K k, int ctas, int nt, A... a)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(nt);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, a...);
    if (e != cudaSuccess) printf("launch: %s\n", cudaGetErrorString(e));
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("sync: %s\n", cudaGetErrorString(e));
}
int main()
{
    long long *cyc; int *buf, *sink;
    cudaMallocManaged(&cyc, 64 * 8); cudaMalloc(&buf, 4 * 2 * 16 * 1024); cudaMalloc(&sink, 4 * 16 * 1024);
    const int n = 256;
    for (int ctas : {2, 4, 8, 16}) {
        launch(k_csync<1024>, ctas, 1024, n, cyc, 0);
        launch(k_csync<256>, ctas, 256, n, cyc, 1);
        launch(k_xread, ctas, 1024, n, cyc, buf, 2, sink);
        launch(k_dsmem, ctas, 1024, n, cyc, 5, sink);
        printf("cluster of %2d CTAs: cluster.sync %6.0f cycles (1024 thr) %6.0f (256 thr) | store %5.0f, sync after stores %6.0f, two dependent ld.cg of peer data %6.0f | dependent DSMEM load %5.0f\n",
               ctas, (double)cyc[0] / n, (double)cyc[1] / n, (double)cyc[2] / n, (double)cyc[3] / n, (double)cyc[4] / n, (double)cyc[5] / n);
    }
    return 0;
}
