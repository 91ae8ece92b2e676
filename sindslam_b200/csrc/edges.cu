// edges.cu -- depth-gradient edges, valid-depth area and edge end points.
//
// Replaces the first half of DynaDetect::CalOccluded (ORB_SLAM2/src/DynaDetect.cc:434-536):
// convertTo(CV_32F) + medianBlur(5) (:435-436), minMaxLoc (:438), the 5x5 max |depth difference| test with
// the invalid-neighbour skip and imgTotalArea (:443-482), MORPH_OPEN 4x4 (:494), end-point detection on the
// 12-point ring (:498-532, ring offsets DynaDetect.h:113-125) and applyNMS(6 px) (:110-143,536).
// EndPoint::curvature is never assigned in the reference (DynaDetect.cc:526 commented out), so its sort is
// a no-op on equal keys; the deterministic restatement keeps raster order.
#include "edges.cuh"

#include "morph.cuh"

#define ET_W 32
#define ET_H 8

// true median of the 5x5 window with replicated borders (cv::medianBlur CV_32F k=5), + global max
__global__ void __launch_bounds__(ET_W *ET_H) k_median5(const uint16_t *__restrict__ depth, int W, int H, float *__restrict__ out,
                                                         unsigned int *__restrict__ gmax)
{
    __shared__ float tile[ET_H + 4][ET_W + 4];
    __shared__ float smax[ET_W * ET_H / 32];
    const int x0 = blockIdx.x * ET_W - 2, y0 = blockIdx.y * ET_H - 2;
    for (int i = threadIdx.y * ET_W + threadIdx.x; i < (ET_W + 4) * (ET_H + 4); i += ET_W * ET_H) {
        int ty = i / (ET_W + 4), tx = i - ty * (ET_W + 4);
        int gx = min(max(x0 + tx, 0), W - 1), gy = min(max(y0 + ty, 0), H - 1);
        tile[ty][tx] = (float)depth[(size_t)gy * W + gx];
    }
    __syncthreads();
    const int x = blockIdx.x * ET_W + threadIdx.x, y = blockIdx.y * ET_H + threadIdx.y;
    float med = 0.f;
    if (x < W && y < H) {
        float v[25];
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j < 5; ++j) v[i * 5 + j] = tile[threadIdx.y + i][threadIdx.x + j];
        // rank selection: the median is the element with #less <= 12 and #less_or_equal >= 13
#pragma unroll
        for (int a = 0; a < 25; ++a) {
            int lt = 0, le = 0;
#pragma unroll
            for (int b = 0; b < 25; ++b) { lt += v[b] < v[a]; le += v[b] <= v[a]; }
            if (lt <= 12 && le >= 13) med = v[a];
        }
        out[(size_t)y * W + x] = med;
    }
    float m = med;
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    int t = threadIdx.y * ET_W + threadIdx.x;
    if ((t & 31) == 0) smax[t >> 5] = m;
    __syncthreads();
    if (t == 0) {
        for (int i = 1; i < ET_W * ET_H / 32; ++i) m = fmaxf(m, smax[i]);
        atomicMax(gmax, __float_as_uint(m));
    }
}

// DynaDetect.cc:443-482
__global__ void __launch_bounds__(ET_W *ET_H) k_grad_edges(const float *__restrict__ filt, int W, int H, const unsigned int *__restrict__ gmax,
                                                            float depth_scale, uint8_t *__restrict__ total_area, uint8_t *__restrict__ occl)
{
    __shared__ float tile[ET_H + 4][ET_W + 4];
    const int x0 = blockIdx.x * ET_W - 2, y0 = blockIdx.y * ET_H - 2;
    for (int i = threadIdx.y * ET_W + threadIdx.x; i < (ET_W + 4) * (ET_H + 4); i += ET_W * ET_H) {
        int ty = i / (ET_W + 4), tx = i - ty * (ET_W + 4);
        int gx = min(max(x0 + tx, 0), W - 1), gy = min(max(y0 + ty, 0), H - 1);
        tile[ty][tx] = filt[(size_t)gy * W + gx];
    }
    __syncthreads();
    const int x = blockIdx.x * ET_W + threadIdx.x, y = blockIdx.y * ET_H + threadIdx.y;
    if (x >= W || y >= H) return;
    uint8_t ta = 0, oc = 0;
    const int range = 3;
    if (x >= range && x < W - range && y >= range && y < H - range) {
        const float half_max = __uint_as_float(*gmax) * 0.5f;
        float depth1 = tile[threadIdx.y + 2][threadIdx.x + 2];
        if (depth1 > 0.0f && depth1 / depth_scale < 6.0f) ta = 255;
        float val_max = 0.0f;
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                float nb = tile[threadIdx.y + i][threadIdx.x + j];
                float d = depth1 - nb;
                if (d > half_max) continue;
                val_max = fmaxf(fabsf(val_max), fabsf(d));
            }
        if (val_max > depth1 * 0.03f && val_max > 400.0f) oc = 255;
    }
    total_area[(size_t)y * W + x] = ta;
    occl[(size_t)y * W + x] = oc;
}

__constant__ int c_ring_x[12] = {0, 1, 2, 2, 2, 1, 0, -1, -2, -2, -2, -1};
__constant__ int c_ring_y[12] = {-2, -2, -1, 0, 1, 2, 2, 2, 1, 0, -1, -2};

// DynaDetect.cc:498-532: edge pixels with at most 4 of the 12 ring pixels set; appended unordered (raster index)
__global__ void k_endpoints(const uint8_t *__restrict__ occl, int W, int H, int *__restrict__ list, int *__restrict__ count, int cap)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < 3 || x >= W - 3 || y < 3 || y >= H - 3) return;
    if (occl[(size_t)y * W + x] != 255) return;
    int s = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) s += occl[(size_t)(y + c_ring_y[i]) * W + x + c_ring_x[i]] == 255;
    if (s <= 4) {
        int k = atomicAdd(count, 1);
        if (k < cap) list[k] = y * W + x;
    }
}

// single CTA: bitonic sort of the raster indices, then the greedy NMS of applyNMS (DynaDetect.cc:110-143)
__global__ void __launch_bounds__(1024) k_endpoints_sort_nms(int *__restrict__ list, const int *__restrict__ count, int cap, int W,
                                                             float dist_thr, int *__restrict__ out_xy, int *__restrict__ out_n,
                                                             int *__restrict__ overflow)
{
    extern __shared__ int keys[];  // cap entries (power of two)
    int n = *count;
    if (n > cap) { if (threadIdx.x == 0) *overflow = 1; n = cap; }
    for (int i = threadIdx.x; i < cap; i += blockDim.x) keys[i] = i < n ? list[i] : 0x7fffffff;
    __syncthreads();
    for (int k = 2; k <= cap; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < cap; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    int a = keys[i], b = keys[ixj];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    // greedy NMS, warp 0: candidate i is kept iff no kept point lies within dist_thr; kept points are in
    // raster order, so only the tail with y >= y_i - ceil(thr) can suppress.
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    const float thr2 = dist_thr * dist_thr;
    const int reach = (int)ceilf(dist_thr);
    int nk = 0;
    for (int i = 0; i < n; ++i) {
        int key = keys[i];
        int py = key / W, px = key - py * W;
        bool suppressed = false;
        for (int base = nk - 1; base >= 0; base -= 32) {
            int j = base - lane;
            bool hit = false, stop = false;
            if (j >= 0) {
                int qx = out_xy[2 * j], qy = out_xy[2 * j + 1];
                if (qy < py - reach) stop = true;
                else {
                    int dx = px - qx, dy = py - qy;
                    float d2 = (float)dx * dx + dy * dy;
                    if (d2 < thr2) hit = true;
                }
            } else stop = true;
            if (__any_sync(0xffffffffu, hit)) { suppressed = true; break; }
            if (__any_sync(0xffffffffu, stop)) break;
        }
        if (!suppressed) {
            if (lane == 0) { out_xy[2 * nk] = px; out_xy[2 * nk + 1] = py; }
            ++nk;
        }
        __syncwarp();
    }
    if (lane == 0) *out_n = nk;
}

int edges_init(sindyn_base *ctx, EdgeStage *e, int W, int H)
{
    e->W = W; e->H = H;
    SD_CHECK(ctx->dalloc(&e->filtered, (size_t)W * H));
    SD_CHECK(ctx->dalloc(&e->total_area, (size_t)W * H));
    SD_CHECK(ctx->dalloc(&e->occl_raw, (size_t)W * H));
    SD_CHECK(ctx->dalloc(&e->grad_edges, (size_t)W * H));
    SD_CHECK(ctx->dalloc(&e->tmp, (size_t)W * H));
    SD_CHECK(ctx->dalloc(&e->ep_list, EDGE_EP_CAP));
    SD_CHECK(ctx->dalloc(&e->ep_xy, 2 * EDGE_EP_CAP));
    SD_CHECK(ctx->dalloc(&e->scalars, 8));
    CU_CHECK(ctx, cudaFuncSetAttribute(k_endpoints_sort_nms, cudaFuncAttributeMaxDynamicSharedMemorySize, EDGE_EP_CAP * (int)sizeof(int)));
    return SINDYN_OK;
}

int edges_run(sindyn_base *ctx, EdgeStage *e, const uint16_t *depth, float depth_scale)
{
    const int W = e->W, H = e->H;
    dim3 blk(ET_W, ET_H), grd(cdiv(W, ET_W), cdiv(H, ET_H));
    CU_CHECK(ctx, cudaMemsetAsync(e->scalars, 0, sizeof(int) * 8, ctx->stream));
    unsigned int *gmax = (unsigned int *)e->scalars;
    int *ep_count = e->scalars + 1, *ep_n = e->scalars + 2, *overflow = e->scalars + 3;
    LAUNCH(ctx, k_median5, grd, blk, 0, depth, W, H, e->filtered, gmax);
    LAUNCH(ctx, k_grad_edges, grd, blk, 0, e->filtered, W, H, gmax, depth_scale, e->total_area, e->occl_raw);
    SD_CHECK(morph_run(ctx, e->occl_raw, e->grad_edges, e->tmp, W, H, 4, MORPH_OPEN));
    LAUNCH(ctx, k_endpoints, grd, blk, 0, e->grad_edges, W, H, e->ep_list, ep_count, EDGE_EP_CAP);
    LAUNCH(ctx, k_endpoints_sort_nms, 1, 1024, EDGE_EP_CAP * sizeof(int), e->ep_list, ep_count, EDGE_EP_CAP, W, 6.0f, e->ep_xy, ep_n, overflow);
    LAUNCH_CHECK(ctx);
    return SINDYN_OK;
}
