// preproc.cu -- colour conversion and OpenCV-exact u8 bilinear resize.
//
// Replaces cv::cvtColor(BGR2GRAY) (ORB_SLAM2/src/DynaDetect.cc:1390-1392), cv::resize(gray, 0.6x,
// INTER_LINEAR) (DynaDetect.cc:1037-1039), GpuMat::convertTo(CV_32F, 1/255) (DynaDetect.cc:1046-1048)
// and the resize inside ORBextractor::ComputePyramid (ORB_SLAM2/src/ORBextractor.cc:1179).
// Arithmetic contracts (tests/test_flow_gpu.py pins them bit for bit against cv2.cvtColor / cv2.resize):
//   gray = (B*3735 + G*19235 + R*9798 + 16384) >> 15
//   resize: 11-bit integer weights, horizontal sum kept as int, vertical
//           dst = (((b0*(h0>>4))>>16) + ((b1*(h1>>4))>>16) + 2) >> 2
#include "preproc.cuh"

#include <math.h>

__global__ void k_bgr2gray(const uint8_t *__restrict__ bgr, int W, int H, uint8_t *__restrict__ gray)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    const uint8_t *p = bgr + 3 * (size_t)i;
    gray[i] = (uint8_t)((p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + 16384) >> 15);
}

int launch_bgr2gray(sindyn_base *ctx, const uint8_t *bgr, int W, int H, uint8_t *gray)
{
    LAUNCH(ctx, k_bgr2gray, cdiv(W * H, 256), 256, 0, bgr, W, H, gray);
    return SINDYN_OK;
}

// src/dst may live inside padded buffers (pitches in bytes); also optionally writes dst/255 as float
__global__ void k_resize_u8(const uint8_t *__restrict__ src, int spitch, uint8_t *__restrict__ dst, int dpitch, int dw, int dh,
                            const int *__restrict__ xofs, const short *__restrict__ xw, const int *__restrict__ yofs,
                            const short *__restrict__ yw, float *__restrict__ dst_f32, float fscale)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    int sx = xofs[x], sy = yofs[y];
    int a0 = xw[2 * x], a1 = xw[2 * x + 1], b0 = yw[2 * y], b1 = yw[2 * y + 1];
    // xofs/yofs are pre-clamped so that (s, s+1) are both valid source indices when the second weight != 0
    const uint8_t *r0 = src + (size_t)sy * spitch, *r1 = src + (size_t)(sy + (b1 != 0)) * spitch;
    int sx1 = sx + (a1 != 0);
    int h0 = r0[sx] * a0 + r0[sx1] * a1;
    int h1 = r1[sx] * a0 + r1[sx1] * a1;
    int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
    uint8_t o = (uint8_t)min(max(v, 0), 255);
    dst[(size_t)y * dpitch + x] = o;
    if (dst_f32) dst_f32[y * dw + x] = (float)o * fscale;
}

static inline short sat_short_round(float v)
{
    long r = lrintf(v);  // round-half-even like cvRound
    if (r > 32767) r = 32767;
    if (r < -32768) r = -32768;
    return (short)r;
}

// Host-side table construction follows cv::resize's INTER_LINEAR set-up (double scale, float fraction).
static void build_axis(int s, int d, std::vector<int> &ofs, std::vector<short> &wts)
{
    ofs.resize(d);
    wts.resize(2 * d);
    double inv_scale = (double)d / (double)s;
    double scale = 1.0 / inv_scale;
    for (int i = 0; i < d; ++i) {
        float f = (float)((i + 0.5) * scale - 0.5);
        int si = (int)floorf(f);
        f -= (float)si;
        if (si < 0) { si = 0; f = 0.f; }
        if (si >= s - 1) { si = s - 1; f = 0.f; }
        ofs[i] = si;
        wts[2 * i] = sat_short_round((1.f - f) * 2048.f);
        wts[2 * i + 1] = sat_short_round(f * 2048.f);
    }
}

int resize_plan_init(sindyn_base *ctx, ResizePlanU8 *p, int sw, int sh, int dw, int dh)
{
    p->sw = sw; p->sh = sh; p->dw = dw; p->dh = dh;
    std::vector<int> xo, yo;
    std::vector<short> xw, yw;
    build_axis(sw, dw, xo, xw);
    build_axis(sh, dh, yo, yw);
    SD_CHECK(ctx->dalloc(&p->xofs, dw));
    SD_CHECK(ctx->dalloc(&p->yofs, dh));
    SD_CHECK(ctx->dalloc(&p->xw, 2 * dw));
    SD_CHECK(ctx->dalloc(&p->yw, 2 * dh));
    CU_CHECK(ctx, cudaMemcpyAsync(p->xofs, xo.data(), sizeof(int) * dw, cudaMemcpyHostToDevice, ctx->stream));
    CU_CHECK(ctx, cudaMemcpyAsync(p->yofs, yo.data(), sizeof(int) * dh, cudaMemcpyHostToDevice, ctx->stream));
    CU_CHECK(ctx, cudaMemcpyAsync(p->xw, xw.data(), sizeof(short) * 2 * dw, cudaMemcpyHostToDevice, ctx->stream));
    CU_CHECK(ctx, cudaMemcpyAsync(p->yw, yw.data(), sizeof(short) * 2 * dh, cudaMemcpyHostToDevice, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));  // host vectors die at return
    return SINDYN_OK;
}

int launch_resize_u8(sindyn_base *ctx, const ResizePlanU8 *p, const uint8_t *src, int spitch, uint8_t *dst, int dpitch,
                     float *dst_f32, float fscale)
{
    dim3 blk(32, 8), grd(cdiv(p->dw, 32), cdiv(p->dh, 8));
    LAUNCH(ctx, k_resize_u8, grd, blk, 0, src, spitch, dst, dpitch, p->dw, p->dh, p->xofs, p->xw, p->yofs, p->yw, dst_f32, fscale);
    return SINDYN_OK;
}

// cv::resize of CV_32FC2, INTER_LINEAR (DynaDetect.cc:1144) fused with the 1/scale multiply (:1147)
__global__ void k_resize_flow(const float2 *__restrict__ src, int sw, int sh, float2 *__restrict__ dst, int dw, int dh,
                              float fx, float fy, float mul)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    float sy = ((float)y + 0.5f) * fy - 0.5f, sx = ((float)x + 0.5f) * fx - 0.5f;
    int y0 = (int)floorf(sy), x0 = (int)floorf(sx);
    float ty = sy - (float)y0, tx = sx - (float)x0;
    if (y0 < 0) { y0 = 0; ty = 0.f; }
    if (y0 >= sh - 1) { y0 = sh - 1; ty = 0.f; }
    if (x0 < 0) { x0 = 0; tx = 0.f; }
    if (x0 >= sw - 1) { x0 = sw - 1; tx = 0.f; }
    int x1 = min(x0 + 1, sw - 1), y1 = min(y0 + 1, sh - 1);
    float2 a = src[y0 * sw + x0], b = src[y0 * sw + x1], c = src[y1 * sw + x0], d = src[y1 * sw + x1];
    // OpenCV float path: horizontal pass a*(1-tx) + b*tx, then vertical
    float w0 = 1.f - tx, v0 = 1.f - ty;
    float2 top = make_float2(a.x * w0 + b.x * tx, a.y * w0 + b.y * tx);
    float2 bot = make_float2(c.x * w0 + d.x * tx, c.y * w0 + d.y * tx);
    dst[y * dw + x] = make_float2((top.x * v0 + bot.x * ty) * mul, (top.y * v0 + bot.y * ty) * mul);
}

int launch_resize_flow(sindyn_base *ctx, const float *src, int sw, int sh, float *dst, int dw, int dh, float mul)
{
    dim3 blk(32, 8), grd(cdiv(dw, 32), cdiv(dh, 8));
    float fx = (float)(1.0 / ((double)dw / (double)sw)), fy = (float)(1.0 / ((double)dh / (double)sh));
    LAUNCH(ctx, k_resize_flow, grd, blk, 0, (const float2 *)src, sw, sh, (float2 *)dst, dw, dh, fx, fy, mul);
    return SINDYN_OK;
}
