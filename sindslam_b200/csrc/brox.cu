// brox.cu -- hand-written sm_100a Brox optical-flow solver.
//
// Replaces cv::cuda::BroxOpticalFlow::create(0.197f, 50.0f, 0.8f, 10, 77, 10)->calc(cur, older, flow)
// (ORB_SLAM2/src/DynaDetect.cc:1029,1072,1124).  Algorithm = oracle/brox_cpu.c (Brox et al. ECCV'04
// with the reference's parameters); the flow-parity gate is a mean end-point-error tolerance.
//
// B200 design: the problem (110 592 px at level 0, 15 levels) is latency- not bandwidth-bound, so
// the solver_iterations red-black SOR sweeps of one lagged-nonlinearity iteration run INSIDE ONE
// launch: each CTA stages its tile plus a (2*sweeps+1)-pixel halo of (du,dv) and the six per-pixel
// system coefficients in shared memory (156 KB of the 227 KB), performs all sweeps there with a
// shrinking valid region (temporal blocking: results are identical to global sweeps), and writes
// only its interior.  32x24 interiors give 144 CTAs at 384x288 = one wave on 148 SMs.  The whole
// pyramid (about 210 launches) is captured in one CUDA graph.
#include "brox.cuh"

#include <math.h>

#define BROX_EPS2 1e-6f

// ------------------------------------------------------------------ small kernels
__global__ void k_resample_f32(const float *__restrict__ src, int sw, int sh, float *__restrict__ dst, int dw, int dh,
                               float fx, float fy, float mul)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    float sy = ((float)y + 0.5f) * fy - 0.5f, sx = ((float)x + 0.5f) * fx - 0.5f;
    int y0 = (int)floorf(sy), x0 = (int)floorf(sx);
    float ty = sy - (float)y0, tx = sx - (float)x0;
    int y0c = min(max(y0, 0), sh - 1), y1c = min(max(y0 + 1, 0), sh - 1);
    int x0c = min(max(x0, 0), sw - 1), x1c = min(max(x0 + 1, 0), sw - 1);
    float a = src[y0c * sw + x0c], b = src[y0c * sw + x1c], c = src[y1c * sw + x0c], d = src[y1c * sw + x1c];
    float top = a + tx * (b - a), bot = c + tx * (d - c);
    dst[y * dw + x] = (top + ty * (bot - top)) * mul;
}

int launch_resample_f32(sindyn_base *ctx, const float *src, int sw, int sh, float *dst, int dw, int dh, float mul)
{
    dim3 blk(32, 8), grd(cdiv(dw, 32), cdiv(dh, 8));
    LAUNCH(ctx, k_resample_f32, grd, blk, 0, src, sw, sh, dst, dw, dh, (float)sw / (float)dw, (float)sh / (float)dh, mul);
    return SINDYN_OK;
}

// two pyramids in one launch (blockIdx.z selects the image)
__global__ void k_brox_pyr_down(const float *__restrict__ s0, const float *__restrict__ s1, int sw, int sh,
                                float *__restrict__ d0, float *__restrict__ d1, int dw, int dh, float fx, float fy)
{
    const float *src = blockIdx.z ? s1 : s0;
    float *dst = blockIdx.z ? d1 : d0;
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    float sy = ((float)y + 0.5f) * fy - 0.5f, sx = ((float)x + 0.5f) * fx - 0.5f;
    int y0 = (int)floorf(sy), x0 = (int)floorf(sx);
    float ty = sy - (float)y0, tx = sx - (float)x0;
    int y0c = min(max(y0, 0), sh - 1), y1c = min(max(y0 + 1, 0), sh - 1);
    int x0c = min(max(x0, 0), sw - 1), x1c = min(max(x0 + 1, 0), sw - 1);
    float a = src[y0c * sw + x0c], b = src[y0c * sw + x1c], c = src[y1c * sw + x0c], d = src[y1c * sw + x1c];
    float top = a + tx * (b - a), bot = c + tx * (d - c);
    dst[y * dw + x] = top + ty * (bot - top);
}

__device__ __forceinline__ float bilinear_clamped(const float *__restrict__ img, int w, int h, float x, float y)
{
    x = fminf(fmaxf(x, 0.0f), (float)(w - 1));
    y = fminf(fmaxf(y, 0.0f), (float)(h - 1));
    int x0 = (int)floorf(x), y0 = (int)floorf(y);
    int x1 = x0 + 1 < w ? x0 + 1 : w - 1, y1 = y0 + 1 < h ? y0 + 1 : h - 1;
    float tx = x - (float)x0, ty = y - (float)y0;
    float a = img[y0 * w + x0], b = img[y0 * w + x1], c = img[y1 * w + x0], d = img[y1 * w + x1];
    float top = a + tx * (b - a), bot = c + tx * (d - c);
    return top + ty * (bot - top);
}

// warp I1 by (u,v); A = (I0 + I1w)/2, Iz = I1w - I0; reset the increment
__global__ void k_brox_warp(const float *__restrict__ I0, const float *__restrict__ I1, const float *__restrict__ u,
                            const float *__restrict__ v, int w, int h, float *__restrict__ A, float *__restrict__ Iz,
                            float *__restrict__ du, float *__restrict__ dv)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    int i = y * w + x;
    float iw = bilinear_clamped(I1, w, h, (float)x + u[i], (float)y + v[i]);
    float i0 = I0[i];
    A[i] = 0.5f * (i0 + iw);
    Iz[i] = iw - i0;
    du[i] = 0.0f;
    dv[i] = 0.0f;
}

__device__ __forceinline__ float d5x(const float *__restrict__ f, int w, int x, int y)
{
    const float *r = f + y * w;
    return (r[max(x - 2, 0)] - 8.0f * r[max(x - 1, 0)] + 8.0f * r[min(x + 1, w - 1)] - r[min(x + 2, w - 1)]) * (1.0f / 12.0f);
}
__device__ __forceinline__ float d5y(const float *__restrict__ f, int w, int h, int x, int y)
{
    return (f[max(y - 2, 0) * w + x] - 8.0f * f[max(y - 1, 0) * w + x] + 8.0f * f[min(y + 1, h - 1) * w + x]
            - f[min(y + 2, h - 1) * w + x]) * (1.0f / 12.0f);
}

__global__ void k_brox_deriv1(const float *__restrict__ A, const float *__restrict__ Iz, int w, int h,
                              float *__restrict__ Ix, float *__restrict__ Iy, float *__restrict__ Ixz, float *__restrict__ Iyz)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    int i = y * w + x;
    Ix[i] = d5x(A, w, x, y);
    Iy[i] = d5y(A, w, h, x, y);
    Ixz[i] = d5x(Iz, w, x, y);
    Iyz[i] = d5y(Iz, w, h, x, y);
}

__global__ void k_brox_deriv2(const float *__restrict__ Ix, const float *__restrict__ Iy, int w, int h,
                              float *__restrict__ Ixx, float *__restrict__ Ixy, float *__restrict__ Iyy)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    int i = y * w + x;
    Ixx[i] = d5x(Ix, w, x, y);
    Ixy[i] = d5y(Ix, w, h, x, y);
    Iyy[i] = d5y(Iy, w, h, x, y);
}

// ------------------------------------------------------------------ the solver kernel
struct BroxInnerP {
    const float *Ix, *Iy, *Iz, *Ixx, *Ixy, *Iyy, *Ixz, *Iyz, *u, *v;
    const float *dub, *dvb;  // increment at the start of this lagged-nonlinearity iteration (coefficients)
    const float *dui, *dvi;  // increment at the start of this launch's sweeps
    float *duo, *dvo;
    int w, h;
    float alpha, gamma, omega;
    int nsweeps;
};

constexpr int BROX_TW = 32, BROX_TH = 24, BROX_SMAX = 10, BROX_NT = 512;
constexpr int BROX_RMAX = 2 * BROX_SMAX + 1;
constexpr int BROX_PW = BROX_TW + 2 * BROX_RMAX, BROX_PH = BROX_TH + 2 * BROX_RMAX;
constexpr int BROX_PP = BROX_PW * BROX_PH;
constexpr size_t BROX_SMEM = (size_t)BROX_PP * 8 * sizeof(float);

__global__ void __launch_bounds__(BROX_NT, 1) k_brox_inner(BroxInnerP p)
{
    extern __shared__ float sm[];
    float *s_du = sm, *s_dv = sm + BROX_PP, *s_ps = sm + 2 * BROX_PP, *s_j12 = sm + 3 * BROX_PP;
    float *s_b1 = sm + 4 * BROX_PP, *s_b2 = sm + 5 * BROX_PP, *s_d1 = sm + 6 * BROX_PP, *s_d2 = sm + 7 * BROX_PP;
    const int gx0 = blockIdx.x * BROX_TW, gy0 = blockIdx.y * BROX_TH;
    const int ox = gx0 - BROX_RMAX, oy = gy0 - BROX_RMAX;
    const int w = p.w, h = p.h;
    const int ns2 = 2 * p.nsweeps;
    const int tid = threadIdx.x;
#define SIDX(gx, gy) (((gy)-oy) * BROX_PW + ((gx)-ox))
#define REGION(r)                                                                             \
    const int xa = max(0, gx0 - (r)), xb = min(w - 1, gx0 + BROX_TW - 1 + (r));                \
    const int ya = max(0, gy0 - (r)), yb = min(h - 1, gy0 + BROX_TH - 1 + (r));                \
    const int rw = xb - xa + 1, rh = yb - ya + 1;

    {   // phase 0: stage (du,dv) and the total flow Uc = u + du_base (scratch in s_b1/s_b2), radius 2ns+1
        REGION(ns2 + 1)
        for (int i = tid; i < rw * rh; i += BROX_NT) {
            int y = ya + i / rw, x = xa + i % rw;
            int g = y * w + x, s = SIDX(x, y);
            s_du[s] = p.dui[g];
            s_dv[s] = p.dvi[g];
            s_b1[s] = p.u[g] + p.dub[g];
            s_b2[s] = p.v[g] + p.dvb[g];
        }
    }
    __syncthreads();
    {   // phase 1: smoothness diffusivity psi'_s from the gradient of the total flow, radius 2ns
        REGION(ns2)
        for (int i = tid; i < rw * rh; i += BROX_NT) {
            int y = ya + i / rw, x = xa + i % rw;
            int xm = max(x - 1, 0), xp = min(x + 1, w - 1), ym = max(y - 1, 0), yp = min(y + 1, h - 1);
            float ux = 0.5f * (s_b1[SIDX(xp, y)] - s_b1[SIDX(xm, y)]), uy = 0.5f * (s_b1[SIDX(x, yp)] - s_b1[SIDX(x, ym)]);
            float vx = 0.5f * (s_b2[SIDX(xp, y)] - s_b2[SIDX(xm, y)]), vy = 0.5f * (s_b2[SIDX(x, yp)] - s_b2[SIDX(x, ym)]);
            s_ps[SIDX(x, y)] = 0.5f / sqrtf(ux * ux + uy * uy + vx * vx + vy * vy + BROX_EPS2);
        }
    }
    __syncthreads();
    {   // phase 2: data-term weights and the per-pixel 2x2 system, radius 2ns-1
        REGION(ns2 - 1)
        const float alpha = p.alpha, gamma = p.gamma;
        for (int i = tid; i < rw * rh; i += BROX_NT) {
            int y = ya + i / rw, x = xa + i % rw;
            int g = y * w + x, s = SIDX(x, y);
            float ix = p.Ix[g], iy = p.Iy[g], iz = p.Iz[g], ixx = p.Ixx[g], ixy = p.Ixy[g], iyy = p.Iyy[g], ixz = p.Ixz[g], iyz = p.Iyz[g];
            float dub = p.dub[g], dvb = p.dvb[g];
            float q0 = iz + ix * dub + iy * dvb;
            float q1 = ixz + ixx * dub + ixy * dvb;
            float q2 = iyz + ixy * dub + iyy * dvb;
            float psid = 0.5f / sqrtf(q0 * q0 + gamma * (q1 * q1 + q2 * q2) + BROX_EPS2);
            float j11 = psid * (ix * ix + gamma * (ixx * ixx + ixy * ixy));
            float j12 = psid * (ix * iy + gamma * (ixx * ixy + ixy * iyy));
            float j22 = psid * (iy * iy + gamma * (ixy * ixy + iyy * iyy));
            float j13 = psid * (ix * iz + gamma * (ixx * ixz + ixy * iyz));
            float j23 = psid * (iy * iz + gamma * (ixy * ixz + iyy * iyz));
            float ps = s_ps[s];
            float uc = p.u[g], vc = p.v[g];
            float su = 0.0f, sv = 0.0f, sw_ = 0.0f;
            if (x > 0) { float wl = alpha * 0.5f * (ps + s_ps[s - 1]); su += wl * (p.u[g - 1] - uc); sv += wl * (p.v[g - 1] - vc); sw_ = wl; }
            float wr = 0.0f, wu = 0.0f, wd = 0.0f;
            if (x < w - 1) { wr = alpha * 0.5f * (ps + s_ps[s + 1]); su += wr * (p.u[g + 1] - uc); sv += wr * (p.v[g + 1] - vc); }
            if (y > 0) { wu = alpha * 0.5f * (ps + s_ps[s - BROX_PW]); su += wu * (p.u[g - w] - uc); sv += wu * (p.v[g - w] - vc); }
            if (y < h - 1) { wd = alpha * 0.5f * (ps + s_ps[s + BROX_PW]); su += wd * (p.u[g + w] - uc); sv += wd * (p.v[g + w] - vc); }
            sw_ = sw_ + wr + wu + wd;
            s_j12[s] = j12;
            s_b1[s] = su - j13;
            s_b2[s] = sv - j23;
            s_d1[s] = 1.0f / (j11 + sw_);
            s_d2[s] = 1.0f / (j22 + sw_);
        }
    }
    __syncthreads();
    // red-black SOR: half-sweep k updates colour (k-1)&1 inside radius 2ns-k
    const float alpha = p.alpha, omega = p.omega, om1 = 1.0f - p.omega;
    for (int k = 1; k <= ns2; ++k) {
        REGION(ns2 - k)
        const int color = (k - 1) & 1;
        const int hc = (rw + 1) >> 1;
        for (int i = tid; i < hc * rh; i += BROX_NT) {
            int y = ya + i / hc;
            int x = xa + 2 * (i % hc) + ((xa + y + color) & 1);
            if (x > xb) continue;
            int s = SIDX(x, y);
            float ps = s_ps[s];
            float su = 0.0f, sv = 0.0f;
            if (x > 0) { float wl = alpha * 0.5f * (ps + s_ps[s - 1]); su += wl * s_du[s - 1]; sv += wl * s_dv[s - 1]; }
            if (x < w - 1) { float wr = alpha * 0.5f * (ps + s_ps[s + 1]); su += wr * s_du[s + 1]; sv += wr * s_dv[s + 1]; }
            if (y > 0) { float wu = alpha * 0.5f * (ps + s_ps[s - BROX_PW]); su += wu * s_du[s - BROX_PW]; sv += wu * s_dv[s - BROX_PW]; }
            if (y < h - 1) { float wd = alpha * 0.5f * (ps + s_ps[s + BROX_PW]); su += wd * s_du[s + BROX_PW]; sv += wd * s_dv[s + BROX_PW]; }
            float j12 = s_j12[s];
            float du_new = om1 * s_du[s] + omega * (s_b1[s] - j12 * s_dv[s] + su) * s_d1[s];
            float dv_new = om1 * s_dv[s] + omega * (s_b2[s] - j12 * du_new + sv) * s_d2[s];
            s_du[s] = du_new;
            s_dv[s] = dv_new;
        }
        __syncthreads();
    }
    {
        REGION(0)
        for (int i = tid; i < rw * rh; i += BROX_NT) {
            int y = ya + i / rw, x = xa + i % rw;
            int g = y * w + x, s = SIDX(x, y);
            p.duo[g] = s_du[s];
            p.dvo[g] = s_dv[s];
        }
    }
#undef SIDX
#undef REGION
}

// level -> finer level: (u+du, v+dv) bilinear, scaled by the size ratios
__global__ void k_brox_prolong(const float *__restrict__ u, const float *__restrict__ v, const float *__restrict__ du,
                               const float *__restrict__ dv, int sw, int sh, float *__restrict__ u2, float *__restrict__ v2,
                               int dw, int dh, float fx, float fy, float mulx, float muly)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    float sy = ((float)y + 0.5f) * fy - 0.5f, sx = ((float)x + 0.5f) * fx - 0.5f;
    int y0 = (int)floorf(sy), x0 = (int)floorf(sx);
    float ty = sy - (float)y0, tx = sx - (float)x0;
    int y0c = min(max(y0, 0), sh - 1), y1c = min(max(y0 + 1, 0), sh - 1);
    int x0c = min(max(x0, 0), sw - 1), x1c = min(max(x0 + 1, 0), sw - 1);
    int ia = y0c * sw + x0c, ib = y0c * sw + x1c, ic = y1c * sw + x0c, id = y1c * sw + x1c;
    {
        float a = u[ia] + du[ia], b = u[ib] + du[ib], c = u[ic] + du[ic], d = u[id] + du[id];
        float top = a + tx * (b - a), bot = c + tx * (d - c);
        u2[y * dw + x] = (top + ty * (bot - top)) * mulx;
    }
    {
        float a = v[ia] + dv[ia], b = v[ib] + dv[ib], c = v[ic] + dv[ic], d = v[id] + dv[id];
        float top = a + tx * (b - a), bot = c + tx * (d - c);
        v2[y * dw + x] = (top + ty * (bot - top)) * muly;
    }
}

__global__ void k_brox_final(const float *__restrict__ u, const float *__restrict__ v, const float *__restrict__ du,
                             const float *__restrict__ dv, int n, float2 *__restrict__ out, float sign)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = make_float2(sign * (u[i] + du[i]), sign * (v[i] + dv[i]));
}

// ------------------------------------------------------------------ host side
int brox_num_levels_host(int w, int h, float scale, int outer, int *ws, int *hs)
{
    int n = 0;
    double s = 1.0;
    while (n < outer && n < BROX_MAX_LEVELS) {
        int lw = (int)ceil((double)w * s - 1e-9), lh = (int)ceil((double)h * s - 1e-9);
        if (n > 0 && (lw < 12 || lh < 12)) break;
        ws[n] = lw;
        hs[n] = lh;
        ++n;
        s *= (double)scale;
    }
    return n;
}

int brox_init(sindyn_base *ctx, BroxSolver *b, int w, int h, float alpha, float gamma, float scale, int inner, int outer,
              int solver, float omega)
{
    b->w = w; b->h = h;
    b->alpha = alpha; b->gamma = gamma; b->scale = scale; b->omega = omega;
    b->inner = inner; b->outer = outer; b->solver = solver;
    b->nl = brox_num_levels_host(w, h, scale, outer, b->ws, b->hs);
    size_t tot = 0;
    for (int k = 1; k < b->nl; ++k) { b->off[k] = tot; tot += (size_t)b->ws[k] * b->hs[k]; }
    b->off[0] = 0;
    SD_CHECK(ctx->dalloc(&b->pyr0, tot));
    SD_CHECK(ctx->dalloc(&b->pyr1, tot));
    size_t n = (size_t)w * h;
    float **planes[] = {&b->A, &b->Iz, &b->Ix, &b->Iy, &b->Ixz, &b->Iyz, &b->Ixx, &b->Ixy, &b->Iyy,
                        &b->u[0], &b->u[1], &b->v[0], &b->v[1], &b->du[0], &b->du[1], &b->du[2], &b->dv[0], &b->dv[1], &b->dv[2]};
    for (float **pp : planes) SD_CHECK(ctx->dalloc(pp, n));
    CU_CHECK(ctx, cudaFuncSetAttribute(k_brox_inner, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BROX_SMEM));
    return SINDYN_OK;
}

static int brox_enqueue(sindyn_base *ctx, BroxSolver *b, const float *I0, const float *I1, float *flow_out, float sign)
{
    const dim3 blk(32, 8);
    // pyramids
    for (int k = 1; k < b->nl; ++k) {
        const float *s0 = k == 1 ? I0 : b->pyr0 + b->off[k - 1], *s1 = k == 1 ? I1 : b->pyr1 + b->off[k - 1];
        dim3 grd(cdiv(b->ws[k], 32), cdiv(b->hs[k], 8), 2);
        LAUNCH(ctx, k_brox_pyr_down, grd, blk, 0, s0, s1, b->ws[k - 1], b->hs[k - 1], b->pyr0 + b->off[k], b->pyr1 + b->off[k],
               b->ws[k], b->hs[k], (float)b->ws[k - 1] / (float)b->ws[k], (float)b->hs[k - 1] / (float)b->hs[k]);
    }
    int cur = 0;  // u[cur], v[cur] hold the flow of the current level
    {
        int kc = b->nl - 1;
        size_t nb = sizeof(float) * (size_t)b->ws[kc] * b->hs[kc];
        CU_CHECK(ctx, cudaMemsetAsync(b->u[cur], 0, nb, ctx->stream));
        CU_CHECK(ctx, cudaMemsetAsync(b->v[cur], 0, nb, ctx->stream));
    }
    int fin = 0;
    for (int k = b->nl - 1; k >= 0; --k) {
        const int w = b->ws[k], h = b->hs[k];
        const float *L0 = k == 0 ? I0 : b->pyr0 + b->off[k], *L1 = k == 0 ? I1 : b->pyr1 + b->off[k];
        dim3 grd(cdiv(w, 32), cdiv(h, 8));
        int base = 0;
        LAUNCH(ctx, k_brox_warp, grd, blk, 0, L0, L1, b->u[cur], b->v[cur], w, h, b->A, b->Iz, b->du[base], b->dv[base]);
        LAUNCH(ctx, k_brox_deriv1, grd, blk, 0, b->A, b->Iz, w, h, b->Ix, b->Iy, b->Ixz, b->Iyz);
        LAUNCH(ctx, k_brox_deriv2, grd, blk, 0, b->Ix, b->Iy, w, h, b->Ixx, b->Ixy, b->Iyy);
        dim3 tgrd(cdiv(w, BROX_TW), cdiv(h, BROX_TH));
        for (int it = 0; it < b->inner; ++it) {
            int in = base, remaining = b->solver;
            while (remaining > 0) {
                int ns = remaining < BROX_SMAX ? remaining : BROX_SMAX;
                int out = 0;
                while (out == base || out == in) ++out;
                BroxInnerP p;
                p.Ix = b->Ix; p.Iy = b->Iy; p.Iz = b->Iz; p.Ixx = b->Ixx; p.Ixy = b->Ixy; p.Iyy = b->Iyy; p.Ixz = b->Ixz; p.Iyz = b->Iyz;
                p.u = b->u[cur]; p.v = b->v[cur];
                p.dub = b->du[base]; p.dvb = b->dv[base];
                p.dui = b->du[in]; p.dvi = b->dv[in];
                p.duo = b->du[out]; p.dvo = b->dv[out];
                p.w = w; p.h = h; p.alpha = b->alpha; p.gamma = b->gamma; p.omega = b->omega; p.nsweeps = ns;
                const bool prof = b->prof_ev && b->prof_n + 2 <= b->prof_cap;
                if (prof) cudaEventRecord(b->prof_ev[b->prof_n++], ctx->stream);
                LAUNCH(ctx, k_brox_inner, tgrd, BROX_NT, BROX_SMEM, p);
                if (prof) { cudaEventRecord(b->prof_ev[b->prof_n++], ctx->stream); b->prof_px += (long long)w * h; }
                in = out;
                remaining -= ns;
            }
            base = in;
        }
        if (k > 0) {
            const int fw = b->ws[k - 1], fh = b->hs[k - 1];
            dim3 fgrd(cdiv(fw, 32), cdiv(fh, 8));
            LAUNCH(ctx, k_brox_prolong, fgrd, blk, 0, b->u[cur], b->v[cur], b->du[base], b->dv[base], w, h, b->u[cur ^ 1],
                   b->v[cur ^ 1], fw, fh, (float)w / (float)fw, (float)h / (float)fh, (float)fw / (float)w, (float)fh / (float)h);
            cur ^= 1;
        } else {
            fin = base;
        }
    }
    int n = b->w * b->h;
    LAUNCH(ctx, k_brox_final, cdiv(n, 256), 256, 0, b->u[cur], b->v[cur], b->du[fin], b->dv[fin], n, (float2 *)flow_out, sign);
    LAUNCH_CHECK(ctx);
    return SINDYN_OK;
}

static void brox_drop_graphs(BroxSolver *b)
{
    for (auto &g : b->slots) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        if (g.graph) cudaGraphDestroy(g.graph);
        g = BroxSolver::GraphSlot();
    }
    b->next_slot = 0;
}

int brox_run(sindyn_base *ctx, BroxSolver *b, const float *I0, const float *I1, float *flow_out, float sign, bool use_graph)
{
    if (!use_graph) return brox_enqueue(ctx, b, I0, I1, flow_out, sign);
    if (!b->graph_ok) { brox_drop_graphs(b); b->graph_ok = true; }
    BroxSolver::GraphSlot *g = nullptr;
    for (auto &s : b->slots)
        if (s.ok && s.I0 == I0 && s.I1 == I1 && s.out == flow_out && s.sign == sign) g = &s;
    if (!g) {
        g = &b->slots[b->next_slot];
        b->next_slot = (b->next_slot + 1) % 4;
        if (g->exec) cudaGraphExecDestroy(g->exec);
        if (g->graph) cudaGraphDestroy(g->graph);
        *g = BroxSolver::GraphSlot();
        unsigned long long before = ctx->launches;
        CU_CHECK(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        int st = brox_enqueue(ctx, b, I0, I1, flow_out, sign);
        cudaError_t e = cudaStreamEndCapture(ctx->stream, &g->graph);
        g->launches = ctx->launches - before;
        ctx->launches = before;
        if (st != SINDYN_OK) return st;
        CU_CHECK(ctx, e);
        CU_CHECK(ctx, cudaGraphInstantiate(&g->exec, g->graph, 0));
        g->I0 = I0; g->I1 = I1; g->out = flow_out; g->sign = sign;
        g->ok = true;
    }
    CU_CHECK(ctx, cudaGraphLaunch(g->exec, ctx->stream));
    ctx->launches += g->launches;
    return SINDYN_OK;
}

void brox_destroy(BroxSolver *b) { brox_drop_graphs(b); b->graph_ok = false; }
