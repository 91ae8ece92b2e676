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
// The 10 red-black half-sweeps of one fixed-point iteration run in one temporally blocked launch (k_vr_sor_fused);
// k_vr_sor is the plain one-colour-per-launch sweep kept for reference.
#include "varref.cuh"

#define VR_FP 5
#define VR_SOR 5
#define VR_ALPHA 20.0f
#define VR_DELTA 5.0f
#define VR_GAMMA 10.0f
#define VR_OMEGA 1.6f
#define VR_ZETA 0.1f
#define VR_EPS 0.001f

enum { P_A = 0, P_IZ, P_IX, P_IY, P_IXZ, P_IYZ, P_IXX, P_IXY, P_IYY, P_WU, P_WV, P_DU, P_DV, P_WGT, P_A11, P_A12, P_A22, P_B1, P_B2, P_DU2, P_DV2, P_COUNT };

// cv::remap(I1 as float, x + u, y + v, INTER_LINEAR, BORDER_REPLICATE) + averaged image + temporal difference
// (I1_alt, sel): when sel is non-null and *sel != 0 the reference image is I1_alt -- the large-motion decision taken on the
// device inside the captured flow graph (flow.cu)
__global__ void k_vr_warp(const uint8_t *__restrict__ I0, const uint8_t *__restrict__ I1_def, const uint8_t *__restrict__ I1_alt,
                          const int *__restrict__ sel, const float2 *__restrict__ flow, int w, int h,
                          float *__restrict__ A, float *__restrict__ Iz, float *__restrict__ Wu, float *__restrict__ Wv, float *__restrict__ du,
                          float *__restrict__ dv)
{
    const uint8_t *__restrict__ I1 = (sel && *sel) ? I1_alt : I1_def;
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

// All VR_SOR red-black sweeps of one fixed-point iteration in ONE launch (temporal blocking, like k_brox_inner): a CTA
// stages its 32x24 tile plus a 2*VR_SOR-px halo of (du, dv) and the six per-pixel system planes in shared memory, sweeps
// there with a shrinking valid region (bit-identical to global sweeps) and writes only its interior.
constexpr int VRT_W = 32, VRT_H = 24, VRT_R = 2 * VR_SOR, VRT_PW = VRT_W + 2 * VRT_R, VRT_PH = VRT_H + 2 * VRT_R, VRT_PP = VRT_PW * VRT_PH;
constexpr int VRT_NT = 512;
constexpr size_t VRT_SMEM = sizeof(float) * 8 * VRT_PP;

__global__ void __launch_bounds__(VRT_NT) k_vr_sor_fused(float *const *__restrict__ P, int w, int h)
{
    extern __shared__ float vsm[];
    float *s_du = vsm, *s_dv = vsm + VRT_PP, *s_w = vsm + 2 * VRT_PP, *s_a11 = vsm + 3 * VRT_PP, *s_a12 = vsm + 4 * VRT_PP, *s_a22 = vsm + 5 * VRT_PP,
          *s_b1 = vsm + 6 * VRT_PP, *s_b2 = vsm + 7 * VRT_PP;
    const int gx0 = blockIdx.x * VRT_W, gy0 = blockIdx.y * VRT_H, ox = gx0 - VRT_R, oy = gy0 - VRT_R;
    const int tid = threadIdx.x;
    for (int r = tid; r < VRT_PP; r += VRT_NT) {
        const int ly = r / VRT_PW, lx = r - ly * VRT_PW;
        const int x = ox + lx, y = oy + ly;
        if (x >= 0 && x < w && y >= 0 && y < h) {
            const int g = y * w + x;
            s_du[r] = P[P_DU][g]; s_dv[r] = P[P_DV][g]; s_w[r] = P[P_WGT][g];
            s_a11[r] = P[P_A11][g]; s_a12[r] = P[P_A12][g]; s_a22[r] = P[P_A22][g]; s_b1[r] = P[P_B1][g]; s_b2[r] = P[P_B2][g];
        } else {
            s_du[r] = s_dv[r] = s_w[r] = s_a12[r] = s_b1[r] = s_b2[r] = 0.0f;
            s_a11[r] = s_a22[r] = 1.0f;
        }
    }
    __syncthreads();
    for (int k = 1; k <= 2 * VR_SOR; ++k) {
        // half-sweep k is valid inside radius VRT_R - k around the tile (clipped to the image)
        const int rad = VRT_R - k;
        const int xa = max(0, gx0 - rad), xb = min(w - 1, gx0 + VRT_W - 1 + rad), ya = max(0, gy0 - rad), yb = min(h - 1, gy0 + VRT_H - 1 + rad);
        const int rw = xb - xa + 1, rh = yb - ya + 1, hc = (rw + 1) >> 1;
        const int colour = (k - 1) & 1;
        for (int i = tid; i < hc * rh; i += VRT_NT) {
            const int yy = i / hc, y = ya + yy;
            const int x = xa + 2 * (i - yy * hc) + ((xa + y + colour) & 1);
            if (x > xb) continue;
            const int s = (y - oy) * VRT_PW + (x - ox);
            const float wc = s_w[s];
            float sU = 0.f, sV = 0.f;
            if (x > 0) { const float wl = s_w[s - 1]; sU += wl * s_du[s - 1]; sV += wl * s_dv[s - 1]; }
            if (x < w - 1) { sU += wc * s_du[s + 1]; sV += wc * s_dv[s + 1]; }
            if (y > 0) { const float wu = s_w[s - VRT_PW]; sU += wu * s_du[s - VRT_PW]; sV += wu * s_dv[s - VRT_PW]; }
            if (y < h - 1) { sU += wc * s_du[s + VRT_PW]; sV += wc * s_dv[s + VRT_PW]; }
            const float a12 = s_a12[s];
            float u = s_du[s], v = s_dv[s];
            u += VR_OMEGA * ((sU + s_b1[s] - v * a12) / s_a11[s] - u);
            v += VR_OMEGA * ((sV + s_b2[s] - u * a12) / s_a22[s] - v);
            s_du[s] = u;
            s_dv[s] = v;
        }
        __syncthreads();
    }
    for (int i = tid; i < VRT_W * VRT_H; i += VRT_NT) {
        const int ly = i / VRT_W, lx = i - ly * VRT_W;
        const int x = gx0 + lx, y = gy0 + ly;
        if (x < w && y < h) {
            const int s = (y - oy) * VRT_PW + (x - ox);
            // results go to the second pair of increment planes: neighbouring CTAs still read the old ones as their halo
            P[P_DU2][y * w + x] = s_du[s];
            P[P_DV2][y * w + x] = s_dv[s];
        }
    }
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
    // two plane tables: the fused SOR kernel reads (DU, DV) and writes (DU2, DV2), so the tables swap those pairs
    float **tab = nullptr;
    SD_CHECK(ctx->dalloc(&tab, 64));
    float *host_tab[64] = {};
    for (int k = 0; k < P_COUNT; ++k) host_tab[k] = host_tab[32 + k] = v->planes[k];
    host_tab[32 + P_DU] = v->planes[P_DU2]; host_tab[32 + P_DV] = v->planes[P_DV2];
    host_tab[32 + P_DU2] = v->planes[P_DU]; host_tab[32 + P_DV2] = v->planes[P_DV];
    CU_CHECK(ctx, cudaMemcpyAsync(tab, host_tab, sizeof(host_tab), cudaMemcpyHostToDevice, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    v->planes_dev = tab;
    CU_CHECK(ctx, cudaFuncSetAttribute(k_vr_sor_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VRT_SMEM));
    v->built = true;
    return SINDYN_OK;
}

int varref_run(sindyn_base *ctx, VarRefStage *v, const uint8_t *I0, const uint8_t *I1, float *flow) { return varref_run_sel(ctx, v, I0, I1, nullptr, nullptr, flow); }

int varref_run_sel(sindyn_base *ctx, VarRefStage *v, const uint8_t *I0, const uint8_t *I1, const uint8_t *I1_alt, const int *sel, float *flow)
{
    const int w = v->w, h = v->h;
    const dim3 blk(32, 8), grd(cdiv(w, 32), cdiv(h, 8));
    float **p = v->planes;
    LAUNCH(ctx, k_vr_warp, grd, blk, 0, I0, I1, I1_alt, sel, (const float2 *)flow, w, h, p[P_A], p[P_IZ], p[P_WU], p[P_WV], p[P_DU], p[P_DV]);
    LAUNCH(ctx, k_vr_deriv1, grd, blk, 0, p[P_A], p[P_IZ], w, h, p[P_IX], p[P_IY], p[P_IXZ], p[P_IYZ]);
    LAUNCH(ctx, k_vr_deriv2, grd, blk, 0, p[P_IX], p[P_IY], w, h, p[P_IXX], p[P_IXY], p[P_IYY]);
    const dim3 grdf(cdiv(w, VRT_W), cdiv(h, VRT_H));
    int cur = 0;
    for (int it = 0; it < VR_FP; ++it) {
        float *const *tab = v->planes_dev + 32 * cur;
        LAUNCH(ctx, k_vr_weights, grd, blk, 0, p[P_WU], p[P_WV], p[cur ? P_DU2 : P_DU], p[cur ? P_DV2 : P_DV], w, h, p[P_WGT]);
        LAUNCH(ctx, k_vr_system, grd, blk, 0, tab, w, h);
        LAUNCH(ctx, k_vr_sor_fused, grdf, VRT_NT, VRT_SMEM, tab, w, h);
        cur ^= 1;
    }
    LAUNCH(ctx, k_vr_final, cdiv(w * h, 256), 256, 0, v->planes_dev + 32 * cur, w * h, (float2 *)flow);
    LAUNCH_CHECK(ctx);
    return SINDYN_OK;
}
