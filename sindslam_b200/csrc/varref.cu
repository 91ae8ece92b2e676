// varref.cu -- cv::VariationalRefinement::create()->calc(I0, I1, flow) with the default parameters
// (ORB_SLAM2/src/DynaDetect.cc:1133-1143): fixedPointIterations 5, sorIterations 5, alpha 20, delta 5, gamma 10,
// omega 1.6, zeta 0.1, epsilon 0.001.
//
// OpenCV's video module is un-vendored third-party code; the algorithm restated here (and pinned numerically against the
// real cv2.VariationalRefinement by the parity tests, max difference ~1e-5 px):
//   warp I1 by the input flow (cv::remap INTER_LINEAR, 1/32-px quantised coordinates, BORDER_REPLICATE),
//   A = (I0 + I1w)/2, Iz = I1w - I0, central differences [-1 0 1] (Sobel ksize 1, replicated border) for
//   Ix Iy Ixz Iyz and Ixx Ixy Iyy; then per fixed-point iteration: normalised brightness + gradient constancy data terms
//   (robust weights delta/2, gamma/2), smoothness weights alpha/2 / sqrt(|grad(W + dW)|^2 + eps^2) from FORWARD
//   differences (one weight per pixel, shared by its right and lower edge), and red-black SOR sweeps on (dWu, dWv).
// The red-black sweeps are global (one launch per colour): 110 592 pixels x 50 half-sweeps is ~0.2 ms, far below the
// Brox solve it follows.
#include "varref.cuh"

#define VR_FP 5
#define VR_SOR 5
#define VR_ALPHA 20.0f
#define VR_DELTA 5.0f
#define VR_GAMMA 10.0f
#define VR_OMEGA 1.6f
#define VR_ZETA 0.1f
#define VR_EPS 0.001f

enum { P_A = 0, P_IZ, P_IX, P_IY, P_IXZ, P_IYZ, P_IXX, P_IXY, P_IYY, P_WU, P_WV, P_DU, P_DV, P_WGT, P_A11, P_A12, P_A22, P_B1, P_B2, P_COUNT };

// cv::remap(I1 as float, x + u, y + v, INTER_LINEAR, BORDER_REPLICATE) + averaged image + temporal difference
__global__ void k_vr_warp(const uint8_t *__restrict__ I0, const uint8_t *__restrict__ I1, const float2 *__restrict__ flow, int w, int h,
                          float *__restrict__ A, float *__restrict__ Iz, float *__restrict__ Wu, float *__restrict__ Wv, float *__restrict__ du,
                          float *__restrict__ dv)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const int i = y * w + x;
    const float2 f = flow[i];
    // convertMaps: fixed-point coordinates with 5 fractional bits (INTER_BITS), round to nearest even
    const int sx = __float2int_rn(((float)x + f.x) * 32.0f), sy = __float2int_rn(((float)y + f.y) * 32.0f);
    const int ix = sx >> 5, iy = sy >> 5;
    const float fx = (float)(sx & 31) * (1.0f / 32.0f), fy = (float)(sy & 31) * (1.0f / 32.0f);
    const int x0 = min(max(ix, 0), w - 1), x1 = min(max(ix + 1, 0), w - 1), y0 = min(max(iy, 0), h - 1), y1 = min(max(iy + 1, 0), h - 1);
    const float a = I1[y0 * w + x0], b = I1[y0 * w + x1], c = I1[y1 * w + x0], d = I1[y1 * w + x1];
    const float w00 = (1.0f - fy) * (1.0f - fx), w01 = (1.0f - fy) * fx, w10 = fy * (1.0f - fx), w11 = fy * fx;
    const float wi = a * w00 + b * w01 + c * w10 + d * w11;
    const float i0 = I0[i];
    A[i] = 0.5f * (i0 + wi);
    Iz[i] = wi - i0;
    Wu[i] = f.x;
    Wv[i] = f.y;
    du[i] = 0.0f;
    dv[i] = 0.0f;
}

__device__ __forceinline__ float cdx(const float *__restrict__ f, int w, int x, int y) { return f[y * w + min(x + 1, w - 1)] - f[y * w + max(x - 1, 0)]; }
__device__ __forceinline__ float cdy(const float *__restrict__ f, int w, int h, int x, int y) { return f[min(y + 1, h - 1) * w + x] - f[max(y - 1, 0) * w + x]; }

__global__ void k_vr_deriv1(const float *__restrict__ A, const float *__restrict__ Iz, int w, int h, float *__restrict__ Ix, float *__restrict__ Iy,
                            float *__restrict__ Ixz, float *__restrict__ Iyz)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const int i = y * w + x;
    Ix[i] = cdx(A, w, x, y);
    Iy[i] = cdy(A, w, h, x, y);
    Ixz[i] = cdx(Iz, w, x, y);
    Iyz[i] = cdy(Iz, w, h, x, y);
}

__global__ void k_vr_deriv2(const float *__restrict__ Ix, const float *__restrict__ Iy, int w, int h, float *__restrict__ Ixx, float *__restrict__ Ixy,
                            float *__restrict__ Iyy)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const int i = y * w + x;
    Ixx[i] = cdx(Ix, w, x, y);
    Ixy[i] = cdy(Ix, w, h, x, y);
    Iyy[i] = cdy(Iy, w, h, x, y);
}

// smoothness weight of every pixel from forward differences of the current flow W + dW
__global__ void k_vr_weights(const float *__restrict__ Wu, const float *__restrict__ Wv, const float *__restrict__ du, const float *__restrict__ dv,
                             int w, int h, float *__restrict__ wgt)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const int i = y * w + x;
    const float cu = Wu[i] + du[i], cv = Wv[i] + dv[i];
    float ux = 0.f, vx = 0.f, uy = 0.f, vy = 0.f;
    if (x < w - 1) { ux = (Wu[i + 1] + du[i + 1]) - cu; vx = (Wv[i + 1] + dv[i + 1]) - cv; }
    if (y < h - 1) { uy = (Wu[i + w] + du[i + w]) - cu; vy = (Wv[i + w] + dv[i + w]) - cv; }
    wgt[i] = (VR_ALPHA * 0.5f) / sqrtf(ux * ux + vx * vx + uy * uy + vy * vy + VR_EPS * VR_EPS);
}

// data terms + smoothness contributions -> the per-pixel 2x2 system
__global__ void k_vr_system(float *const *__restrict__ P, int w, int h)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const int i = y * w + x;
    const float z2 = VR_ZETA * VR_ZETA, e2 = VR_EPS * VR_EPS;
    const float ix = P[P_IX][i], iy = P[P_IY][i], iz = P[P_IZ][i], ixx = P[P_IXX][i], ixy = P[P_IXY][i], iyy = P[P_IYY][i], ixz = P[P_IXZ][i],
                iyz = P[P_IYZ][i];
    const float du = P[P_DU][i], dv = P[P_DV][i];
    // brightness constancy, normalised (Zimmer et al.)
    float dn = ix * ix + iy * iy + z2;
    const float ik1z = iz + ix * du + iy * dv;
    float wt = ((VR_DELTA * 0.5f) / sqrtf(ik1z * ik1z / dn + e2)) / dn;
    float a11 = wt * (ix * ix) + z2, a12 = wt * (ix * iy), a22 = wt * (iy * iy) + z2, b1 = -wt * (iz * ix), b2 = -wt * (iz * iy);
    // gradient constancy
    const float dn1 = ixx * ixx + ixy * ixy + z2, dn2 = iyy * iyy + ixy * ixy + z2;
    const float ik1zx = ixz + ixx * du + ixy * dv, ik1zy = iyz + ixy * du + iyy * dv;
    wt = (VR_GAMMA * 0.5f) / sqrtf(ik1zx * ik1zx / dn1 + ik1zy * ik1zy / dn2 + e2);
    a11 += wt * (ixx * ixx / dn1 + ixy * ixy / dn2);
    a12 += wt * (ixx * ixy / dn1 + ixy * iyy / dn2);
    a22 += wt * (ixy * ixy / dn1 + iyy * iyy / dn2);
    b1 -= wt * (ixx * ixz / dn1 + ixy * iyz / dn2);
    b2 -= wt * (ixy * ixz / dn1 + iyy * iyz / dn2);
    // smoothness: edge (p, right) and (p, down) carry the weight of p; edge (left, p) / (up, p) the weight of left / up
    const float *wgt = P[P_WGT], *Wu = P[P_WU], *Wv = P[P_WV];
    const float wc = wgt[i];
    const float u0 = Wu[i], v0 = Wv[i];
    if (x < w - 1) { a11 += wc; a22 += wc; b1 += wc * (Wu[i + 1] - u0); b2 += wc * (Wv[i + 1] - v0); }
    if (y < h - 1) { a11 += wc; a22 += wc; b1 += wc * (Wu[i + w] - u0); b2 += wc * (Wv[i + w] - v0); }
    if (x > 0) { const float wl = wgt[i - 1]; a11 += wl; a22 += wl; b1 -= wl * (u0 - Wu[i - 1]); b2 -= wl * (v0 - Wv[i - 1]); }
    if (y > 0) { const float wu = wgt[i - w]; a11 += wu; a22 += wu; b1 -= wu * (u0 - Wu[i - w]); b2 -= wu * (v0 - Wv[i - w]); }
    P[P_A11][i] = a11; P[P_A12][i] = a12; P[P_A22][i] = a22; P[P_B1][i] = b1; P[P_B2][i] = b2;
}

__global__ void k_vr_sor(float *const *__restrict__ P, int w, int h, int colour)
{
    const int xh = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (y >= h) return;
    const int x = 2 * xh + ((y + colour) & 1);
    if (x >= w) return;
    const int i = y * w + x;
    const float *wgt = P[P_WGT];
    float *du = P[P_DU], *dv = P[P_DV];
    const float wc = wgt[i];
    float sU = 0.f, sV = 0.f;
    if (x > 0) { const float wl = wgt[i - 1]; sU += wl * du[i - 1]; sV += wl * dv[i - 1]; }
    if (x < w - 1) { sU += wc * du[i + 1]; sV += wc * dv[i + 1]; }
    if (y > 0) { const float wu = wgt[i - w]; sU += wu * du[i - w]; sV += wu * dv[i - w]; }
    if (y < h - 1) { sU += wc * du[i + w]; sV += wc * dv[i + w]; }
    const float a12 = P[P_A12][i];
    float u = du[i], v = dv[i];
    u += VR_OMEGA * ((sU + P[P_B1][i] - v * a12) / P[P_A11][i] - u);
    v += VR_OMEGA * ((sV + P[P_B2][i] - u * a12) / P[P_A22][i] - v);
    du[i] = u;
    dv[i] = v;
}

__global__ void k_vr_final(float *const *__restrict__ P, int n, float2 *__restrict__ flow)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flow[i] = make_float2(P[P_WU][i] + P[P_DU][i], P[P_WV][i] + P[P_DV][i]);
}

int varref_init(sindyn_base *ctx, VarRefStage *v, int w, int h)
{
    v->w = w; v->h = h;
    const size_t n = (size_t)w * h;
    SD_CHECK(ctx->dalloc(&v->buf, n * P_COUNT));
    for (int k = 0; k < P_COUNT; ++k) v->planes[k] = v->buf + n * k;
    float **tab = nullptr;
    SD_CHECK(ctx->dalloc(&tab, 32));
    CU_CHECK(ctx, cudaMemcpyAsync(tab, v->planes, sizeof(float *) * 32, cudaMemcpyHostToDevice, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    v->planes_dev = tab;
    v->built = true;
    return SINDYN_OK;
}

int varref_run(sindyn_base *ctx, VarRefStage *v, const uint8_t *I0, const uint8_t *I1, float *flow)
{
    const int w = v->w, h = v->h;
    const dim3 blk(32, 8), grd(cdiv(w, 32), cdiv(h, 8)), grdh(cdiv((w + 1) / 2, 32), cdiv(h, 8));
    float **p = v->planes;
    LAUNCH(ctx, k_vr_warp, grd, blk, 0, I0, I1, (const float2 *)flow, w, h, p[P_A], p[P_IZ], p[P_WU], p[P_WV], p[P_DU], p[P_DV]);
    LAUNCH(ctx, k_vr_deriv1, grd, blk, 0, p[P_A], p[P_IZ], w, h, p[P_IX], p[P_IY], p[P_IXZ], p[P_IYZ]);
    LAUNCH(ctx, k_vr_deriv2, grd, blk, 0, p[P_IX], p[P_IY], w, h, p[P_IXX], p[P_IXY], p[P_IYY]);
    for (int it = 0; it < VR_FP; ++it) {
        LAUNCH(ctx, k_vr_weights, grd, blk, 0, p[P_WU], p[P_WV], p[P_DU], p[P_DV], w, h, p[P_WGT]);
        LAUNCH(ctx, k_vr_system, grd, blk, 0, v->planes_dev, w, h);
        for (int s = 0; s < VR_SOR; ++s) {
            LAUNCH(ctx, k_vr_sor, grdh, blk, 0, v->planes_dev, w, h, 0);
            LAUNCH(ctx, k_vr_sor, grdh, blk, 0, v->planes_dev, w, h, 1);
        }
    }
    LAUNCH(ctx, k_vr_final, cdiv(w * h, 256), 256, 0, v->planes_dev, w * h, (float2 *)flow);
    LAUNCH_CHECK(ctx);
    return SINDYN_OK;
}
