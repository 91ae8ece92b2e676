/*
 * oracle/brox_cpu.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, load or call this file.  The product path is libsindyn_cuda.so and must never link it.
 *
 * What it restates: the dense-flow engine the reference calls as
 *   cv::cuda::BroxOpticalFlow::create(0.197f, 50.0f, 0.8f, 10, 77, 10)   (ORB_SLAM2/src/DynaDetect.cc:1029)
 *   denseFlow3->calc(cur_32F, older_32F, flow)                           (ORB_SLAM2/src/DynaDetect.cc:1072,1124)
 * i.e. (alpha, gamma, scale_factor, inner_iterations, outer_iterations, solver_iterations) on
 * 384x288 float images in [0,1] (DynaDetect.cc:1033-1048).
 *
 * PARITY UNPINNED: the solver itself lives in OpenCV-contrib 4.2.0 (cudaoptflow -> cudalegacy
 * NCVBroxOpticalFlow), an un-vendored dependency absent from /root/reference, and cv2 4.13-headless
 * in this image has no cudaoptflow.  This file therefore restates the PUBLISHED algorithm
 * (Brox, Bruhn, Papenberg, Weickert, "High accuracy optical flow estimation based on a theory
 * for warping", ECCV 2004) with the reference's parameters:
 *   - coarse-to-fine pyramid, factor `scale`, at most `outer` levels, smallest side >= 12 px;
 *   - per level ONE warp of I1 by the current flow, 5-tap derivatives [1 -8 0 8 -1]/12 of the
 *     averaged image, then `inner` lagged-nonlinearity iterations each followed by `solver`
 *     red-black SOR sweeps (relaxation omega) on the increment (du,dv);
 *   - robust function Psi(s^2)=sqrt(s^2+eps^2), eps^2 = 1e-6, gradient-constancy weight gamma,
 *     smoothness weight alpha.
 * Flow parity of the CUDA solver is a mean end-point-error tolerance against THIS solver run
 * with identical parameters (BASELINE.json north_star), plus EPE against the renderer's
 * analytic flow.
 *
 * Build: make -C oracle   (-> oracle/_build/libbrox_cpu.so)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BROX_EPS2 1e-6f
#define MAX_LEVELS 128

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* pixel-centre aligned bilinear resample (same coordinate rule as cv::resize INTER_LINEAR) */
static void resample_bilinear(const float *src, int sw, int sh, float *dst, int dw, int dh, float mul)
{
    const float fx = (float)sw / (float)dw, fy = (float)sh / (float)dh;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dh; ++y) {
        float sy = ((float)y + 0.5f) * fy - 0.5f;
        int y0 = (int)floorf(sy);
        float ty = sy - (float)y0;
        int y0c = clampi(y0, 0, sh - 1), y1c = clampi(y0 + 1, 0, sh - 1);
        for (int x = 0; x < dw; ++x) {
            float sx = ((float)x + 0.5f) * fx - 0.5f;
            int x0 = (int)floorf(sx);
            float tx = sx - (float)x0;
            int x0c = clampi(x0, 0, sw - 1), x1c = clampi(x0 + 1, 0, sw - 1);
            float a = src[y0c * sw + x0c], b = src[y0c * sw + x1c];
            float c = src[y1c * sw + x0c], d = src[y1c * sw + x1c];
            float top = a + tx * (b - a), bot = c + tx * (d - c);
            dst[y * dw + x] = (top + ty * (bot - top)) * mul;
        }
    }
}

static inline float sample_bilinear_clamped(const float *img, int w, int h, float x, float y)
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

static inline float ddx(const float *f, int w, int x, int y)
{
    const float *r = f + y * w;
    return (r[clampi(x - 2, 0, w - 1)] - 8.0f * r[clampi(x - 1, 0, w - 1)]
            + 8.0f * r[clampi(x + 1, 0, w - 1)] - r[clampi(x + 2, 0, w - 1)]) * (1.0f / 12.0f);
}
static inline float ddy(const float *f, int w, int h, int x, int y)
{
    return (f[clampi(y - 2, 0, h - 1) * w + x] - 8.0f * f[clampi(y - 1, 0, h - 1) * w + x]
            + 8.0f * f[clampi(y + 1, 0, h - 1) * w + x] - f[clampi(y + 2, 0, h - 1) * w + x]) * (1.0f / 12.0f);
}

int brox_num_levels(int w, int h, float scale, int outer, int *ws, int *hs)
{
    int n = 0;
    double s = 1.0;
    while (n < outer && n < MAX_LEVELS) {
        int lw = (int)ceil((double)w * s - 1e-9), lh = (int)ceil((double)h * s - 1e-9);
        if (n > 0 && (lw < 12 || lh < 12)) break;
        if (ws) ws[n] = lw;
        if (hs) hs[n] = lh;
        ++n;
        s *= (double)scale;
    }
    return n;
}

/* One level: warp, derivatives, `inner` x (coefficients + `solver` red-black SOR sweeps). */
static void brox_level(const float *I0, const float *I1, int w, int h, float *u, float *v,
                       float alpha, float gamma, int inner, int solver, float omega)
{
    const int n = w * h;
    float *buf = (float *)malloc(sizeof(float) * (size_t)n * 19);
    float *A = buf, *Iz = A + n, *Ix = Iz + n, *Iy = Ix + n, *Ixz = Iy + n, *Iyz = Ixz + n;
    float *Ixx = Iyz + n, *Ixy = Ixx + n, *Iyy = Ixy + n, *du = Iyy + n, *dv = du + n;
    float *psis = dv + n, *J12 = psis + n, *b1 = J12 + n, *b2 = b1 + n, *d1 = b2 + n, *d2 = d1 + n;
    float *Uc = d2 + n, *Vc = Uc + n;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int i = y * w + x;
            float iw = sample_bilinear_clamped(I1, w, h, (float)x + u[i], (float)y + v[i]);
            A[i] = 0.5f * (I0[i] + iw);
            Iz[i] = iw - I0[i];
            du[i] = 0.0f;
            dv[i] = 0.0f;
        }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int i = y * w + x;
            Ix[i] = ddx(A, w, x, y);
            Iy[i] = ddy(A, w, h, x, y);
            Ixz[i] = ddx(Iz, w, x, y);
            Iyz[i] = ddy(Iz, w, h, x, y);
        }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int i = y * w + x;
            Ixx[i] = ddx(Ix, w, x, y);
            Ixy[i] = ddy(Ix, w, h, x, y);
            Iyy[i] = ddy(Iy, w, h, x, y);
        }
    for (int it = 0; it < inner; ++it) {
        /* smoothness diffusivity from the gradient of the CURRENT total flow u+du, v+dv */
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; ++i) { Uc[i] = u[i] + du[i]; Vc[i] = v[i] + dv[i]; }
#pragma omp parallel for schedule(static)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                int i = y * w + x;
                int xm = x > 0 ? x - 1 : 0, xp = x < w - 1 ? x + 1 : w - 1;
                int ym = y > 0 ? y - 1 : 0, yp = y < h - 1 ? y + 1 : h - 1;
                float ux = 0.5f * (Uc[y * w + xp] - Uc[y * w + xm]), uy = 0.5f * (Uc[yp * w + x] - Uc[ym * w + x]);
                float vx = 0.5f * (Vc[y * w + xp] - Vc[y * w + xm]), vy = 0.5f * (Vc[yp * w + x] - Vc[ym * w + x]);
                psis[i] = 0.5f / sqrtf(ux * ux + uy * uy + vx * vx + vy * vy + BROX_EPS2);
            }
#pragma omp parallel for schedule(static)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                int i = y * w + x;
                float ix = Ix[i], iy = Iy[i], iz = Iz[i], ixx = Ixx[i], ixy = Ixy[i], iyy = Iyy[i], ixz = Ixz[i], iyz = Iyz[i];
                float q0 = iz + ix * du[i] + iy * dv[i];
                float q1 = ixz + ixx * du[i] + ixy * dv[i];
                float q2 = iyz + ixy * du[i] + iyy * dv[i];
                float psid = 0.5f / sqrtf(q0 * q0 + gamma * (q1 * q1 + q2 * q2) + BROX_EPS2);
                float j11 = psid * (ix * ix + gamma * (ixx * ixx + ixy * ixy));
                float j12 = psid * (ix * iy + gamma * (ixx * ixy + ixy * iyy));
                float j22 = psid * (iy * iy + gamma * (ixy * ixy + iyy * iyy));
                float j13 = psid * (ix * iz + gamma * (ixx * ixz + ixy * iyz));
                float j23 = psid * (iy * iz + gamma * (ixy * ixz + iyy * iyz));
                float ps = psis[i];
                float wl = x > 0 ? alpha * 0.5f * (ps + psis[i - 1]) : 0.0f;
                float wr = x < w - 1 ? alpha * 0.5f * (ps + psis[i + 1]) : 0.0f;
                float wu = y > 0 ? alpha * 0.5f * (ps + psis[i - w]) : 0.0f;
                float wd = y < h - 1 ? alpha * 0.5f * (ps + psis[i + w]) : 0.0f;
                float uc = u[i], vc = v[i];
                float su = 0.0f, sv = 0.0f;
                if (x > 0) { su += wl * (u[i - 1] - uc); sv += wl * (v[i - 1] - vc); }
                if (x < w - 1) { su += wr * (u[i + 1] - uc); sv += wr * (v[i + 1] - vc); }
                if (y > 0) { su += wu * (u[i - w] - uc); sv += wu * (v[i - w] - vc); }
                if (y < h - 1) { su += wd * (u[i + w] - uc); sv += wd * (v[i + w] - vc); }
                float sw_ = wl + wr + wu + wd;
                J12[i] = j12;
                b1[i] = su - j13;
                b2[i] = sv - j23;
                d1[i] = 1.0f / (j11 + sw_);
                d2[i] = 1.0f / (j22 + sw_);
            }
        for (int s = 0; s < solver; ++s)
            for (int color = 0; color < 2; ++color) {
#pragma omp parallel for schedule(static)
                for (int y = 0; y < h; ++y)
                    for (int x = (y + color) & 1; x < w; x += 2) {
                        int i = y * w + x;
                        float ps = psis[i];
                        float su = 0.0f, sv = 0.0f;
                        if (x > 0) { float wl = alpha * 0.5f * (ps + psis[i - 1]); su += wl * du[i - 1]; sv += wl * dv[i - 1]; }
                        if (x < w - 1) { float wr = alpha * 0.5f * (ps + psis[i + 1]); su += wr * du[i + 1]; sv += wr * dv[i + 1]; }
                        if (y > 0) { float wu = alpha * 0.5f * (ps + psis[i - w]); su += wu * du[i - w]; sv += wu * dv[i - w]; }
                        if (y < h - 1) { float wd = alpha * 0.5f * (ps + psis[i + w]); su += wd * du[i + w]; sv += wd * dv[i + w]; }
                        float du_new = (1.0f - omega) * du[i] + omega * (b1[i] - J12[i] * dv[i] + su) * d1[i];
                        float dv_new = (1.0f - omega) * dv[i] + omega * (b2[i] - J12[i] * du_new + sv) * d2[i];
                        du[i] = du_new;
                        dv[i] = dv_new;
                    }
            }
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) { u[i] += du[i]; v[i] += dv[i]; }
    free(buf);
}

/* I0 = current frame, I1 = older frame (argument order of DynaDetect.cc:1072); flow is w with
 * I0(x) ~ I1(x + w(x)), interleaved (u,v) like CV_32FC2. Returns the number of pyramid levels. */
int brox_flow_cpu(const float *I0, const float *I1, int w, int h, float alpha, float gamma, float scale,
                  int inner, int outer, int solver, float omega, float *flow_uv)
{
    int ws[MAX_LEVELS], hs[MAX_LEVELS];
    int nl = brox_num_levels(w, h, scale, outer, ws, hs);
    float *p0[MAX_LEVELS], *p1[MAX_LEVELS];
    p0[0] = (float *)I0;
    p1[0] = (float *)I1;
    for (int k = 1; k < nl; ++k) {
        p0[k] = (float *)malloc(sizeof(float) * (size_t)ws[k] * hs[k]);
        p1[k] = (float *)malloc(sizeof(float) * (size_t)ws[k] * hs[k]);
        resample_bilinear(p0[k - 1], ws[k - 1], hs[k - 1], p0[k], ws[k], hs[k], 1.0f);
        resample_bilinear(p1[k - 1], ws[k - 1], hs[k - 1], p1[k], ws[k], hs[k], 1.0f);
    }
    float *u = (float *)calloc((size_t)w * h, sizeof(float)), *v = (float *)calloc((size_t)w * h, sizeof(float));
    float *u2 = (float *)malloc(sizeof(float) * (size_t)w * h), *v2 = (float *)malloc(sizeof(float) * (size_t)w * h);
    for (int k = nl - 1; k >= 0; --k) {
        brox_level(p0[k], p1[k], ws[k], hs[k], u, v, alpha, gamma, inner, solver, omega);
        if (k > 0) {
            resample_bilinear(u, ws[k], hs[k], u2, ws[k - 1], hs[k - 1], (float)ws[k - 1] / (float)ws[k]);
            resample_bilinear(v, ws[k], hs[k], v2, ws[k - 1], hs[k - 1], (float)hs[k - 1] / (float)hs[k]);
            float *t = u; u = u2; u2 = t;
            t = v; v = v2; v2 = t;
        }
    }
    for (int i = 0; i < w * h; ++i) { flow_uv[2 * i] = u[i]; flow_uv[2 * i + 1] = v[i]; }
    for (int k = 1; k < nl; ++k) { free(p0[k]); free(p1[k]); }
    free(u); free(v); free(u2); free(v2);
    return nl;
}

int brox_cpu_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
