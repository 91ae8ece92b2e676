// morph.cu -- flat morphology with OpenCV's elliptic structuring elements.
// Semantics (pinned against cv2 by the oracle tests):
//   out(y,x) = max/min over {(i,j): SE[i][j] = 1} of src(y + i - k/2, x + j - k/2), out-of-image samples ignored,
//   OPEN = erode -> dilate, CLOSE = dilate -> erode; even k are asymmetric (anchor k/2).
// Every SE row is one contiguous span, so a row costs (j2 - j1) shared-memory reads.
#include "morph.cuh"

#include <math.h>

__constant__ signed char c_j1[MORPH_MAX_K + 1][MORPH_MAX_K];
__constant__ signed char c_j2[MORPH_MAX_K + 1][MORPH_MAX_K];

void ellipse_spans(int k, int *j1, int *j2)
{
    // cv::getStructuringElement(MORPH_ELLIPSE, Size(k,k)): r = k/2, c = k/2, inv_r2 = 1/r^2;
    // row i: dy = i - r; |dy| <= r ? dx = saturate_cast<int>(c * sqrt((r*r - dy*dy) * inv_r2)) : empty
    int r = k / 2, c = k / 2;
    double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
    for (int i = 0; i < k; ++i) {
        int dy = i - r;
        int a = 0, b = 0;
        if (abs(dy) <= r) {
            int dx = (int)lrint(c * sqrt((r * r - dy * dy) * inv_r2));
            a = c - dx > 0 ? c - dx : 0;
            b = c + dx + 1 < k ? c + dx + 1 : k;
        }
        j1[i] = a;
        j2[i] = b;
    }
}

int morph_init(sindyn_base *ctx)
{
    signed char h1[MORPH_MAX_K + 1][MORPH_MAX_K] = {}, h2[MORPH_MAX_K + 1][MORPH_MAX_K] = {};
    for (int k = 1; k <= MORPH_MAX_K; ++k) {
        int a[MORPH_MAX_K], b[MORPH_MAX_K];
        ellipse_spans(k, a, b);
        for (int i = 0; i < k; ++i) { h1[k][i] = (signed char)a[i]; h2[k][i] = (signed char)b[i]; }
    }
    CU_CHECK(ctx, cudaMemcpyToSymbol(c_j1, h1, sizeof h1));
    CU_CHECK(ctx, cudaMemcpyToSymbol(c_j2, h2, sizeof h2));
    return SINDYN_OK;
}

#define MT_W 32
#define MT_H 16
#define MT_PW (MT_W + MORPH_MAX_K)
#define MT_PH (MT_H + MORPH_MAX_K)

template <bool ERODE>
__global__ void __launch_bounds__(MT_W *MT_H) k_morph(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int W, int H, int k)
{
    __shared__ uint8_t tile[MT_PH][MT_PW + 1];
    const int a = k / 2;            // anchor
    const int x0 = blockIdx.x * MT_W - a, y0 = blockIdx.y * MT_H - a;
    const int tw = MT_W + k - 1, th = MT_H + k - 1;
    const uint8_t fill = ERODE ? 255 : 0;
    for (int i = threadIdx.y * MT_W + threadIdx.x; i < tw * th; i += MT_W * MT_H) {
        int ty = i / tw, tx = i - ty * tw;
        int gx = x0 + tx, gy = y0 + ty;
        tile[ty][tx] = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? src[(size_t)gy * W + gx] : fill;
    }
    __syncthreads();
    const int x = blockIdx.x * MT_W + threadIdx.x, y = blockIdx.y * MT_H + threadIdx.y;
    if (x >= W || y >= H) return;
    int v = fill;
    for (int i = 0; i < k; ++i) {
        const int j1 = c_j1[k][i], j2 = c_j2[k][i];
        const uint8_t *row = &tile[threadIdx.y + i][threadIdx.x];
        for (int j = j1; j < j2; ++j) v = ERODE ? min(v, (int)row[j]) : max(v, (int)row[j]);
    }
    dst[(size_t)y * W + x] = (uint8_t)v;
}

static void launch_one(sindyn_base *ctx, const uint8_t *src, uint8_t *dst, int W, int H, int k, bool erode)
{
    dim3 blk(MT_W, MT_H), grd(cdiv(W, MT_W), cdiv(H, MT_H));
    if (erode) LAUNCH(ctx, k_morph<true>, grd, blk, 0, src, dst, W, H, k);
    else LAUNCH(ctx, k_morph<false>, grd, blk, 0, src, dst, W, H, k);
}

int morph_run(sindyn_base *ctx, const uint8_t *src, uint8_t *dst, uint8_t *tmp, int W, int H, int k, int op)
{
    if (k < 1 || k > MORPH_MAX_K) { ctx->err = "morph: k out of range"; return SINDYN_ERR_INVALID; }
    switch (op) {
    case MORPH_DILATE: launch_one(ctx, src, dst, W, H, k, false); break;
    case MORPH_ERODE: launch_one(ctx, src, dst, W, H, k, true); break;
    case MORPH_OPEN: launch_one(ctx, src, tmp, W, H, k, true); launch_one(ctx, tmp, dst, W, H, k, false); break;
    case MORPH_CLOSE: launch_one(ctx, src, tmp, W, H, k, false); launch_one(ctx, tmp, dst, W, H, k, true); break;
    default: return SINDYN_ERR_INVALID;
    }
    LAUNCH_CHECK(ctx);
    return SINDYN_OK;
}

// ------------------------------------------------------------------ bitset morphology
__device__ __forceinline__ uint16_t b_or(uint16_t a, uint16_t b) { return a | b; }
__device__ __forceinline__ uint16_t b_and(uint16_t a, uint16_t b) { return a & b; }
__device__ __forceinline__ ulonglong2 b_or(ulonglong2 a, ulonglong2 b) { return make_ulonglong2(a.x | b.x, a.y | b.y); }
__device__ __forceinline__ ulonglong2 b_and(ulonglong2 a, ulonglong2 b) { return make_ulonglong2(a.x & b.x, a.y & b.y); }
template <class T> __device__ __forceinline__ T b_fill(bool ones);
template <> __device__ __forceinline__ uint16_t b_fill<uint16_t>(bool ones) { return ones ? 0xFFFFu : 0u; }
template <> __device__ __forceinline__ ulonglong2 b_fill<ulonglong2>(bool ones)
{
    return ones ? make_ulonglong2(~0ull, ~0ull) : make_ulonglong2(0ull, 0ull);
}

#define BT_W 32
#define BT_H 8
template <class T, bool ERODE>
__global__ void __launch_bounds__(BT_W *BT_H) k_morph_bits(const T *__restrict__ src, T *__restrict__ dst, int W, int H, int k)
{
    __shared__ T tile[BT_H + MORPH_MAX_K][BT_W + MORPH_MAX_K];
    const int a = k / 2;
    const int x0 = blockIdx.x * BT_W - a, y0 = blockIdx.y * BT_H - a;
    const int tw = BT_W + k - 1, th = BT_H + k - 1;
    const T fill = b_fill<T>(ERODE);
    for (int i = threadIdx.y * BT_W + threadIdx.x; i < tw * th; i += BT_W * BT_H) {
        int ty = i / tw, tx = i - ty * tw;
        int gx = x0 + tx, gy = y0 + ty;
        tile[ty][tx] = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? src[(size_t)gy * W + gx] : fill;
    }
    __syncthreads();
    const int x = blockIdx.x * BT_W + threadIdx.x, y = blockIdx.y * BT_H + threadIdx.y;
    if (x >= W || y >= H) return;
    T v = fill;
    for (int i = 0; i < k; ++i) {
        const int j1 = c_j1[k][i], j2 = c_j2[k][i];
        for (int j = j1; j < j2; ++j) {
            T t = tile[threadIdx.y + i][threadIdx.x + j];
            v = ERODE ? b_and(v, t) : b_or(v, t);
        }
    }
    dst[(size_t)y * W + x] = v;
}

int morph_bits_run(sindyn_base *ctx, const void *src, void *dst, int W, int H, int k, bool erode, int elem_bytes)
{
    if (k < 1 || k > MORPH_MAX_K || (elem_bytes != 2 && elem_bytes != 16)) { ctx->err = "morph_bits: bad arguments"; return SINDYN_ERR_INVALID; }
    dim3 blk(BT_W, BT_H), grd(cdiv(W, BT_W), cdiv(H, BT_H));
    if (elem_bytes == 2) {
        if (erode) LAUNCH(ctx, (k_morph_bits<uint16_t, true>), grd, blk, 0, (const uint16_t *)src, (uint16_t *)dst, W, H, k);
        else LAUNCH(ctx, (k_morph_bits<uint16_t, false>), grd, blk, 0, (const uint16_t *)src, (uint16_t *)dst, W, H, k);
    } else {
        if (erode) LAUNCH(ctx, (k_morph_bits<ulonglong2, true>), grd, blk, 0, (const ulonglong2 *)src, (ulonglong2 *)dst, W, H, k);
        else LAUNCH(ctx, (k_morph_bits<ulonglong2, false>), grd, blk, 0, (const ulonglong2 *)src, (ulonglong2 *)dst, W, H, k);
    }
    LAUNCH_CHECK(ctx);
    return SINDYN_OK;
}
