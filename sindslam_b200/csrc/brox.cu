// brox.cu -- hand-written sm_100a Brox optical-flow solver.
//
// Replaces cv::cuda::BroxOpticalFlow::create(0.197f, 50.0f, 0.8f, 10, 77, 10)->calc(cur, older, flow)
// (ORB_SLAM2/src/DynaDetect.cc:1029,1072,1124).  Algorithm = oracle/brox_cpu.c (Brox et al. ECCV'04
// with the reference's parameters); the flow-parity gate is a mean end-point-error tolerance.
//
// B200 design: the problem (110 592 px at level 0, 15 levels) is latency- not bandwidth-bound, so the
// solver_iterations red-black SOR sweeps of one lagged-nonlinearity iteration run INSIDE ONE launch (temporal
// blocking in shared memory, k_brox_sor) after one k_brox_system launch that prepares the per-pixel systems without halo
// redundancy; levels pick the smallest tile that still fits one wave of 148 SMs, the six coarsest levels (<= 2100 px) run
// all inner iterations in a single one-CTA launch (k_brox_level), and the whole pyramid (about 320 launches) is captured
// in one CUDA graph.
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
    pdl_wait();
    pdl_trigger();
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

// ------------------------------------------------------------------ the solver kernels
// Shared design of the sweep kernels (k_brox_level: small levels, one CTA, all inner iterations in one launch;
// k_brox_sor: tiles of larger levels, one inner iteration per launch, systems precomputed by k_brox_system):
//   * red and black pixels are stored DE-INTERLEAVED (two arrays of half-width rows), so the stride-2 access pattern of a
//     red-black sweep becomes stride-1 and bank-conflict free; (du,dv) is a float2 -> LDS.64, the two edge weights are two
//     float arrays; zero guards around every colour array replace bounds logic in the sweep;
//   * every thread owns a FIXED set of <= 3 red + 3 black pixels for the whole launch: their current (du,dv) stay in
//     registers, the 2x2 system of a pixel (j12, b1, b2, 1/d1 | 1/d2) in the thread's private shared-memory slots
//     (1024 threads, 64 registers each);
//   * tiles: the tile plus a (2 x sweeps + 1)-pixel ring lives in shared memory -- shipped: 5 fused sweeps per launch, ring of
//     11 (two launches per lagged-nonlinearity iteration; 10 sweeps / ring 21 measured slower); the fused sweeps are exact
//     because a pixel at distance d from a halo edge stays valid for d half-sweeps (temporal blocking) and only the
//     interior is written.
// Warp I1 by (u, v): A = (I0 + I1w) / 2, Iz = I1w - I0; 5-tap first derivatives of A and Iz, second derivatives of A; reset the
// increment -- one launch: a 32 x 8 tile warps its pixels plus a 4-pixel ring into shared
// memory (the 5-tap second derivatives reach 4 pixels out; 2.5x redundant bilinear fetches, no intermediate planes through
// L2, two launches less per pyramid level).  Border handling equals the separate kernels: every stencil tap clamps its
// coordinate into the image, and a derivative "at" a clamped coordinate is the derivative of that border pixel.
constexpr int BWD_W = 32, BWD_H = 8, BWD_R = 4, BWD_PW = BWD_W + 2 * BWD_R, BWD_PH = BWD_H + 2 * BWD_R;

__global__ void __launch_bounds__(BWD_W *BWD_H) k_brox_warp_deriv(const float *__restrict__ I0, const float *__restrict__ I1, const float *__restrict__ u,
                                                                  const float *__restrict__ v, int w, int h, float *__restrict__ A,
                                                                  float *__restrict__ Iz, float *__restrict__ Ix, float *__restrict__ Iy,
                                                                  float *__restrict__ Ixz, float *__restrict__ Iyz, float *__restrict__ Ixx,
                                                                  float *__restrict__ Ixy, float *__restrict__ Iyy, float *__restrict__ du,
                                                                  float *__restrict__ dv)
{
    __shared__ float s_A[BWD_PH][BWD_PW], s_Iz[BWD_PH][BWD_PW], s_Ix[BWD_PH][BWD_PW], s_Iy[BWD_PH][BWD_PW];
    pdl_wait();
    pdl_trigger();
    const int tx0 = blockIdx.x * BWD_W - BWD_R, ty0 = blockIdx.y * BWD_H - BWD_R;   // image coordinate of shared (0, 0)
    const int tid = threadIdx.y * BWD_W + threadIdx.x;
    for (int i = tid; i < BWD_PW * BWD_PH; i += BWD_W * BWD_H) {
        const int ly = i / BWD_PW, lx = i - ly * BWD_PW;
        const int x = tx0 + lx, y = ty0 + ly;
        if (x >= 0 && x < w && y >= 0 && y < h) {
            const int g = y * w + x;
            const float iw = bilinear_clamped(I1, w, h, (float)x + u[g], (float)y + v[g]);
            const float i0 = I0[g];
            s_A[ly][lx] = 0.5f * (i0 + iw);
            s_Iz[ly][lx] = iw - i0;
        }
    }
    __syncthreads();
    // taps at clamped image coordinates (a clamped coordinate of a pixel within 2 of this tile's 2-ring lies inside the staged region)
#define BWD_D5X(P, X, Y)                                                                                                       \
    (((P)[(Y) - ty0][max((X) - 2, 0) - tx0] - 8.0f * (P)[(Y) - ty0][max((X) - 1, 0) - tx0] + 8.0f * (P)[(Y) - ty0][min((X) + 1, w - 1) - tx0] - \
      (P)[(Y) - ty0][min((X) + 2, w - 1) - tx0]) * (1.0f / 12.0f))
#define BWD_D5Y(P, X, Y)                                                                                                       \
    (((P)[max((Y) - 2, 0) - ty0][(X) - tx0] - 8.0f * (P)[max((Y) - 1, 0) - ty0][(X) - tx0] + 8.0f * (P)[min((Y) + 1, h - 1) - ty0][(X) - tx0] - \
      (P)[min((Y) + 2, h - 1) - ty0][(X) - tx0]) * (1.0f / 12.0f))
    for (int i = tid; i < (BWD_W + 4) * (BWD_H + 4); i += BWD_W * BWD_H) {
        const int ly = i / (BWD_W + 4) + 2, lx = i - (ly - 2) * (BWD_W + 4) + 2;
        const int x = tx0 + lx, y = ty0 + ly;
        if (x >= 0 && x < w && y >= 0 && y < h) {
            s_Ix[ly][lx] = BWD_D5X(s_A, x, y);
            s_Iy[ly][lx] = BWD_D5Y(s_A, x, y);
        }
    }
    __syncthreads();
    const int x = tx0 + BWD_R + threadIdx.x, y = ty0 + BWD_R + threadIdx.y;
    if (x >= w || y >= h) return;
    const int g = y * w + x;
    A[g] = s_A[y - ty0][x - tx0];
    Iz[g] = s_Iz[y - ty0][x - tx0];
    Ix[g] = s_Ix[y - ty0][x - tx0];
    Iy[g] = s_Iy[y - ty0][x - tx0];
    Ixz[g] = BWD_D5X(s_Iz, x, y);
    Iyz[g] = BWD_D5Y(s_Iz, x, y);
    Ixx[g] = BWD_D5X(s_Ix, x, y);
    Ixy[g] = BWD_D5Y(s_Ix, x, y);
    Iyy[g] = BWD_D5Y(s_Iy, x, y);
    du[g] = 0.0f;
    dv[g] = 0.0f;
#undef BWD_D5X
#undef BWD_D5Y
}

#ifdef SINDYN_BROX_PHASE_CLOCKS
// developer instrumentation: clock64 deltas of the centre CTA of every 32x24-tile k_brox_sor launch, summed per phase
__device__ unsigned long long g_brox_clk[16];
#define BROX_CLK(slot)                                                                                        \
    if (prof_on) { const long long t_ = clock64(); atomicAdd(&g_brox_clk[slot], (unsigned long long)(t_ - t_prev)); t_prev = t_; }
extern "C" int sindyn_dbg_brox_phase_clocks(unsigned long long *out, int reset)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_brox_clk, sizeof(g_brox_clk));
    if (reset) { unsigned long long z[16] = {}; cudaMemcpyToSymbol(g_brox_clk, z, sizeof(z)); }
    return 0;
}
#else
#define BROX_CLK(slot)
#endif

struct BroxInnerP {
    const float *Ix, *Iy, *Iz, *Ixx, *Ixy, *Iyy, *Ixz, *Iyz, *u, *v;
    const float *dub, *dvb;  // increment at the start of the launch
    float *duo, *dvo;
    int w, h;
    float alpha, gamma, omega;
    int nsweeps, n_inner;
};

#ifndef SINDYN_BROX_SMAX
#define SINDYN_BROX_SMAX 5
#endif
#ifndef SINDYN_BROX_SMAX_SMALL
#define SINDYN_BROX_SMAX_SMALL 5
#endif
// sweeps fused into one k_brox_sor launch: levels on the three larger tiles / levels on 8x8 tiles (where the fixed cost of a
// launch weighs more than the halo redundancy)
constexpr int BROX_SMAX = SINDYN_BROX_SMAX, BROX_SMAX_SMALL = SINDYN_BROX_SMAX_SMALL, BROX_NT = 1024;
// Tile menu: a level uses the smallest tile whose grid still fits into one wave of 148 SMs -- the time of a launch is
// the time of ONE CTA, which is proportional to the staged region (tile + 2 x (2 x sweeps + 1) halo), so mid-size levels run on
// many small tiles instead of a few large ones.
template <int TW_, int TH_, int SMAX_ = BROX_SMAX> struct BroxTile {
    static constexpr int TW = TW_, TH = TH_, SMAX = SMAX_, R = 2 * SMAX_ + 1;
    static constexpr int PW = TW + 2 * R, PH = TH + 2 * R;   // staged region
    static constexpr int HW = PW / 2;                                        // pixels of one colour per row
    static constexpr int NPC = PH * HW;                                      // pixels per colour
    static constexpr int M = (NPC + BROX_NT - 1) / BROX_NT;                  // owned pixels per colour and thread
    static constexpr int PP = PW * PH;
    static constexpr int G = HW + 1;                 // zero guard before / after every colour array (wrapped neighbour reads)
    static constexpr int NPCP = NPC + 2 * G;
    // (du,dv) + edge weights | per-pixel 2x2 systems (5 floats; the staging planes ta / tb / ps alias their start) | u, v
    static constexpr size_t SMEM = sizeof(float2) * 4 * NPCP + sizeof(float) * 10 * NPC + sizeof(float) * 2 * PP;
    static constexpr size_t SMEM_SOR = sizeof(float2) * 2 * NPCP;   // k_brox_sor: the (du, dv) pairs only
    static_assert(3 * PP <= 10 * NPC, "ta / tb / ps must fit into the coefficient area they alias");
    static_assert((PW & 1) == 0, "de-interleaving needs an even region width");
    static_assert(NPC < 4096, "pixel index must fit into 12 bits");
};
typedef BroxTile<8, 8, BROX_SMAX_SMALL> BroxTileS;
typedef BroxTile<32, 24> BroxTileL;   // 74 x 66 region: also the single-tile mode of the coarse levels


// Whole level in ONE CTA (levels of up to 1024 * M pixels per colour): all n_inner lagged-nonlinearity iterations x nsweeps
// red-black sweeps in one launch, no halo.  The geometry is a run-time parameter (region = the level itself, width rounded
// up to even plus a dead column) so that the owned pixels are packed into the first warps with full lanes; the block is sized to the level.
template <int M>
__global__ void __launch_bounds__(BROX_NT, 1) k_brox_level(BroxInnerP p)
{
    const int w = p.w, h = p.h;
    // at least one dead column ends every row: with rows packed back to back the element before a row's first pixel is the
    // previous row's last one, and the sweeps rely on its weights being zero
    const int PW = (w + 2) & ~1, HW = PW >> 1, PH = h, NPC = PH * HW, PP = PW * PH, G = HW + 1, NPCP = NPC + 2 * G;
    const int NT = blockDim.x;
    extern __shared__ float4 sm4[];
    float2 *s_uv = (float2 *)sm4 + G;                                  // [2][NPCP] (du, dv) by colour, zero guards
    float *s_wr = (float *)((float2 *)sm4 + 2 * NPCP) + G;             // [2][NPCP] weight to the right neighbour
    float *s_wd = s_wr + 2 * NPCP;                                     // [2][NPCP] weight to the lower neighbour
    float4 *s_c4 = (float4 *)((float2 *)sm4 + 4 * NPCP);               // [2][NPC]  (j12, b1, b2, 1/d1), private to the owner
    float *s_c1 = (float *)(s_c4 + 2 * NPC);                           // [2][NPC]  1/d2
    float *s_ta = (float *)s_c4, *s_tb = s_ta + PP, *s_ps = s_tb + PP; // staging planes alias the systems (3 PP <= 10 NPC)
    float *s_u = (float *)s_c4 + 10 * NPC, *s_v = s_u + PP;            // [PP] flow of this level
    const int tid = threadIdx.x;
    const float alpha = p.alpha, gamma = p.gamma, omega = p.omega, om1 = 1.0f - p.omega;
    const float inv_pw = 1.0f / (float)PW, inv_hw = 1.0f / (float)HW;   // exact row index for the index ranges used here

    // ---- per-thread pixel table: pk = idx | parity << 12 | live << 14
    unsigned pk[2][M];
    float rdu[2][M], rdv[2][M];
    // image derivatives of the owned pixels: constant over the inner iterations, kept in registers when one pixel per colour is
    // owned (with two the kernel would spill: they are re-read from L2 every iteration)
    constexpr bool KEEP = M == 1;
    float dI[2][KEEP ? M : 1][8];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const int q = tid + NT * m;
            unsigned v = 0;
            if (q < NPC) {
                const int ly = (int)(((float)q + 0.5f) * inv_hw), i = q - ly * HW;
                const int par = (ly + c) & 1, lx = 2 * i + par;
                v = (unsigned)q | ((unsigned)par << 12) | ((unsigned)(lx < w) << 14);
            }
            pk[c][m] = v;
            rdu[c][m] = rdv[c][m] = 0.0f;
        }
    for (int r = tid; r < 4 * NPCP; r += NT) ((float2 *)sm4)[r] = make_float2(0.0f, 0.0f);
    pdl_wait();      // everything above is independent of the previous kernel's output
    pdl_trigger();
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int m = 0; m < (KEEP ? M : 0); ++m) {
            const unsigned k = pk[c][m];
            int g = 0;
            if ((k >> 14) & 1u) {
                const int idx = k & 0xfff, par = (k >> 12) & 1;
                const int y = (int)(((float)idx + 0.5f) * inv_hw), x = 2 * (idx - y * HW) + par;
                g = y * w + x;
            }
            dI[c][m][0] = p.Ix[g]; dI[c][m][1] = p.Iy[g]; dI[c][m][2] = p.Iz[g]; dI[c][m][3] = p.Ixx[g];
            dI[c][m][4] = p.Ixy[g]; dI[c][m][5] = p.Iyy[g]; dI[c][m][6] = p.Ixz[g]; dI[c][m][7] = p.Iyz[g];
        }
    __syncthreads();
    for (int it = 0; it < p.n_inner; ++it) {
        // ---- phase 0: (du, dv), the level's flow (u, v) and the total flow
        for (int r = tid; r < PP; r += NT) {
            const int ly = (int)(((float)r + 0.5f) * inv_pw), lx = r - ly * PW;
            const int ci = ((lx + ly) & 1) * NPCP + ly * HW + (lx >> 1);
            float ta = 0.0f, tb = 0.0f;
            if (lx < w) {
                if (it == 0) {
                    const int g = ly * w + lx;
                    const float bu = p.dub[g], bv = p.dvb[g], uu = p.u[g], vv = p.v[g];
                    s_uv[ci] = make_float2(bu, bv);
                    s_u[r] = uu;
                    s_v[r] = vv;
                    ta = uu + bu;
                    tb = vv + bv;
                } else {   // the increment of the previous inner iteration is already in shared memory
                    const float2 d = s_uv[ci];
                    ta = s_u[r] + d.x;
                    tb = s_v[r] + d.y;
                }
            } else if (it == 0) {
                s_u[r] = 0.0f;
                s_v[r] = 0.0f;
            }
            s_ta[r] = ta;
            s_tb[r] = tb;
        }
        __syncthreads();
        // ---- phase 1: smoothness diffusivity psi'_s from the gradient of the total flow
        for (int r = tid; r < PP; r += NT) {
            const int y = (int)(((float)r + 0.5f) * inv_pw), x = r - y * PW;
            float ps = 0.0f;
            if (x < w) {
                const int rm = x > 0 ? r - 1 : r, rp = x < w - 1 ? r + 1 : r, ru = y > 0 ? r - PW : r, rd = y < h - 1 ? r + PW : r;
                const float ux = 0.5f * (s_ta[rp] - s_ta[rm]), uy = 0.5f * (s_ta[rd] - s_ta[ru]);
                const float vx = 0.5f * (s_tb[rp] - s_tb[rm]), vy = 0.5f * (s_tb[rd] - s_tb[ru]);
                ps = 0.5f / sqrtf(ux * ux + uy * uy + vx * vx + vy * vy + BROX_EPS2);
            }
            s_ps[r] = ps;
        }
        __syncthreads();
        // ---- phase 2a: edge weights (right, down) of every pixel; zero across the image border (Neumann)
        for (int r = tid; r < PP; r += NT) {
            const int y = (int)(((float)r + 0.5f) * inv_pw), x = r - y * PW;
            float wr = 0.0f, wd = 0.0f;
            if (x < w) {
                const float ps = s_ps[r];
                if (x < w - 1) wr = alpha * 0.5f * (ps + s_ps[r + 1]);
                if (y < h - 1) wd = alpha * 0.5f * (ps + s_ps[r + PW]);
            }
            const int wi = ((x + y) & 1) * NPCP + y * HW + (x >> 1);
            s_wr[wi] = wr;
            s_wd[wi] = wd;
        }
        __syncthreads();
        // ---- phase 2b: data term and the 2x2 system of the owned pixels
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const unsigned k = pk[c][m];
                if (!((k >> 14) & 1u)) continue;
                const int idx = k & 0xfff, par = (k >> 12) & 1;
                const int y = (int)(((float)idx + 0.5f) * inv_hw), x = 2 * (idx - y * HW) + par;
                const float wl = x > 0 ? s_wr[(c ^ 1) * NPCP + idx - 1 + par] : 0.0f, wr = s_wr[c * NPCP + idx];
                const float wu = y > 0 ? s_wd[(c ^ 1) * NPCP + idx - HW] : 0.0f, wd = s_wd[c * NPCP + idx];
                const float2 own = s_uv[c * NPCP + idx];
                rdu[c][m] = own.x;
                rdv[c][m] = own.y;
                const float dub = own.x, dvb = own.y;   // == the increment at the start of this inner iteration
                float ix, iy, iz, ixx, ixy, iyy, ixz, iyz;
                if (KEEP) {
                    ix = dI[c][KEEP ? m : 0][0]; iy = dI[c][KEEP ? m : 0][1]; iz = dI[c][KEEP ? m : 0][2]; ixx = dI[c][KEEP ? m : 0][3];
                    ixy = dI[c][KEEP ? m : 0][4]; iyy = dI[c][KEEP ? m : 0][5]; ixz = dI[c][KEEP ? m : 0][6]; iyz = dI[c][KEEP ? m : 0][7];
                } else {
                    const int g = y * w + x;
                    ix = p.Ix[g]; iy = p.Iy[g]; iz = p.Iz[g]; ixx = p.Ixx[g]; ixy = p.Ixy[g]; iyy = p.Iyy[g]; ixz = p.Ixz[g]; iyz = p.Iyz[g];
                }
                const float q0 = iz + ix * dub + iy * dvb;
                const float q1 = ixz + ixx * dub + ixy * dvb;
                const float q2 = iyz + ixy * dub + iyy * dvb;
                const float psid = 0.5f / sqrtf(q0 * q0 + gamma * (q1 * q1 + q2 * q2) + BROX_EPS2);
                const float j11 = psid * (ix * ix + gamma * (ixx * ixx + ixy * ixy));
                const float j12 = psid * (ix * iy + gamma * (ixx * ixy + ixy * iyy));
                const float j22 = psid * (iy * iy + gamma * (ixy * ixy + iyy * iyy));
                const float j13 = psid * (ix * iz + gamma * (ixx * ixz + ixy * iyz));
                const float j23 = psid * (iy * iz + gamma * (ixy * ixz + iyy * iyz));
                const int r = y * PW + x;
                const float uc = s_u[r], vc = s_v[r];
                float su = 0.0f, sv = 0.0f;
                if (x > 0) { su += wl * (s_u[r - 1] - uc); sv += wl * (s_v[r - 1] - vc); }
                if (x < w - 1) { su += wr * (s_u[r + 1] - uc); sv += wr * (s_v[r + 1] - vc); }
                if (y > 0) { su += wu * (s_u[r - PW] - uc); sv += wu * (s_v[r - PW] - vc); }
                if (y < h - 1) { su += wd * (s_u[r + PW] - uc); sv += wd * (s_v[r + PW] - vc); }
                const float sw_ = wl + wr + wu + wd;
                s_c4[c * NPC + idx] = make_float4(j12, su - j13, sv - j23, 1.0f / (j11 + sw_));
                s_c1[c * NPC + idx] = 1.0f / (j22 + sw_);
            }
        // (no barrier needed: phase 2b reads none of the planes the systems alias, and the sweeps read s_uv / s_wr / s_wd,
        //  complete since the last barrier, plus the thread's own systems)
        // ---- red-black SOR
#define BROX_HALF(C)                                                                                              \
    {                                                                                                             \
        const float2 *__restrict__ uo = s_uv + ((C) ^ 1) * NPCP;                                                   \
        float2 *__restrict__ uc_ = s_uv + (C) * NPCP;                                                              \
        const float *__restrict__ wro = s_wr + ((C) ^ 1) * NPCP;                                                   \
        const float *__restrict__ wdo = s_wd + ((C) ^ 1) * NPCP;                                                   \
        const float *__restrict__ wrc = s_wr + (C) * NPCP;                                                         \
        const float *__restrict__ wdc = s_wd + (C) * NPCP;                                                         \
        _Pragma("unroll") for (int m = 0; m < M; ++m)                                                             \
        {                                                                                                         \
            const unsigned k_ = pk[C][m];                                                                         \
            if ((k_ >> 14) & 1u) {                                                                                \
                const int idx = k_ & 0xfff, par = (k_ >> 12) & 1;                                                 \
                const float2 l = uo[idx - 1 + par], r = uo[idx + par], u_ = uo[idx - HW], d = uo[idx + HW];       \
                const float wr_ = wrc[idx], wd_ = wdc[idx];                                                       \
                const float wl = wro[idx - 1 + par], wu = wdo[idx - HW];                                          \
                const float su = wl * l.x + wr_ * r.x + wu * u_.x + wd_ * d.x;                                    \
                const float sv = wl * l.y + wr_ * r.y + wu * u_.y + wd_ * d.y;                                    \
                const float4 cf = s_c4[(C) * NPC + idx];                                                          \
                const float cd2_ = s_c1[(C) * NPC + idx];                                                         \
                const float du_new = om1 * rdu[C][m] + omega * (cf.y - cf.x * rdv[C][m] + su) * cf.w;             \
                const float dv_new = om1 * rdv[C][m] + omega * (cf.z - cf.x * du_new + sv) * cd2_;                \
                rdu[C][m] = du_new;                                                                               \
                rdv[C][m] = dv_new;                                                                               \
                uc_[idx] = make_float2(du_new, dv_new);                                                           \
            }                                                                                                     \
        }                                                                                                         \
        __syncthreads();                                                                                          \
    }
        for (int sw = 0; sw < p.nsweeps; ++sw) {
            BROX_HALF(0)
            BROX_HALF(1)
        }
#undef BROX_HALF
    }
    // ---- write the level from registers
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const unsigned k = pk[c][m];
            if (!((k >> 14) & 1u)) continue;
            const int idx = k & 0xfff, par = (k >> 12) & 1;
            const int y = (int)(((float)idx + 0.5f) * inv_hw), x = 2 * (idx - y * HW) + par;
            p.duo[y * w + x] = rdu[c][m];
            p.dvo[y * w + x] = rdv[c][m];
        }
}

// ------------------------------------------------------------------ tiled levels: system kernel + SOR kernel
// For levels larger than one tile the lagged-nonlinearity coefficients are computed ONCE per pixel by k_brox_system
// (no halo redundancy: computed inside the sweep kernel they ran on the whole tile + halo region -- 6.4x the pixels of
// the tile with the 21-px ring of 10 fused sweeps -- and took 44 % of the launch) and the temporally blocked sweeps
// (k_brox_sor) only stage them.
struct BroxSysP {
    const float *Ix, *Iy, *Iz, *Ixx, *Ixy, *Iyy, *Ixz, *Iyz, *u, *v, *dub, *dvb;
    float4 *W;       // (wl, wr, wu, wd): the four edge weights of the pixel (zero across the image border)
    float4 *C4;
    float *C1;
    int w, h;
    float alpha, gamma;
};
constexpr int BSY_W = 32, BSY_H = 8;

__global__ void __launch_bounds__(BSY_W *BSY_H) k_brox_system(BroxSysP p)
{
    pdl_wait();
    pdl_trigger();
    __shared__ float s_ta[BSY_H + 4][BSY_W + 4], s_tb[BSY_H + 4][BSY_W + 4], s_u[BSY_H + 4][BSY_W + 4], s_v[BSY_H + 4][BSY_W + 4];
    __shared__ float s_ps[BSY_H + 2][BSY_W + 2];
    const int w = p.w, h = p.h;
    const int gx0 = blockIdx.x * BSY_W, gy0 = blockIdx.y * BSY_H;
    const int tid = threadIdx.y * BSY_W + threadIdx.x;
    // the thread's own pixel: its image derivatives are loaded first, their latency overlaps the two staging phases below
    const int x = gx0 + threadIdx.x, y = gy0 + threadIdx.y;
    const bool own = x < w && y < h;
    const int g = own ? y * w + x : 0;
    const float dub = p.dub[g], dvb = p.dvb[g];
    const float ix = p.Ix[g], iy = p.Iy[g], iz = p.Iz[g], ixx = p.Ixx[g], ixy = p.Ixy[g], iyy = p.Iyy[g], ixz = p.Ixz[g], iyz = p.Iyz[g];
    for (int i = tid; i < (BSY_H + 4) * (BSY_W + 4); i += BSY_W * BSY_H) {
        const int ly = i / (BSY_W + 4), lx = i - ly * (BSY_W + 4);
        const int x = gx0 - 2 + lx, y = gy0 - 2 + ly;
        float uu = 0.0f, vv = 0.0f, bu = 0.0f, bv = 0.0f;
        if (x >= 0 && x < w && y >= 0 && y < h) {
            const int g = y * w + x;
            uu = p.u[g]; vv = p.v[g]; bu = p.dub[g]; bv = p.dvb[g];
        }
        s_u[ly][lx] = uu; s_v[ly][lx] = vv;
        s_ta[ly][lx] = uu + bu; s_tb[ly][lx] = vv + bv;
    }
    __syncthreads();
    // smoothness diffusivity psi'_s on the tile + 1 ring (central differences, replicated at the image border)
    for (int i = tid; i < (BSY_H + 2) * (BSY_W + 2); i += BSY_W * BSY_H) {
        const int ly = i / (BSY_W + 2), lx = i - ly * (BSY_W + 2);
        const int x = gx0 - 1 + lx, y = gy0 - 1 + ly;
        float ps = 0.0f;
        if (x >= 0 && x < w && y >= 0 && y < h) {
            const int cy = ly + 1, cx = lx + 1;
            const int xm = x > 0 ? cx - 1 : cx, xp = x < w - 1 ? cx + 1 : cx, ym = y > 0 ? cy - 1 : cy, yp = y < h - 1 ? cy + 1 : cy;
            const float ux = 0.5f * (s_ta[cy][xp] - s_ta[cy][xm]), uy = 0.5f * (s_ta[yp][cx] - s_ta[ym][cx]);
            const float vx = 0.5f * (s_tb[cy][xp] - s_tb[cy][xm]), vy = 0.5f * (s_tb[yp][cx] - s_tb[ym][cx]);
            ps = 0.5f / sqrtf(ux * ux + uy * uy + vx * vx + vy * vy + BROX_EPS2);
        }
        s_ps[ly][lx] = ps;
    }
    __syncthreads();
    if (!own) return;
    const int py = threadIdx.y + 1, px = threadIdx.x + 1;   // s_ps coordinates
    const float ps = s_ps[py][px];
    // edge weights; zero across the image border (Neumann).  wl / wu are the right / lower weights of the left / upper pixel
    const float wr = x < w - 1 ? p.alpha * 0.5f * (ps + s_ps[py][px + 1]) : 0.0f;
    const float wd = y < h - 1 ? p.alpha * 0.5f * (ps + s_ps[py + 1][px]) : 0.0f;
    const float wl = x > 0 ? p.alpha * 0.5f * (s_ps[py][px - 1] + ps) : 0.0f;
    const float wu = y > 0 ? p.alpha * 0.5f * (s_ps[py - 1][px] + ps) : 0.0f;
    const float gamma = p.gamma;
    const float q0 = iz + ix * dub + iy * dvb;
    const float q1 = ixz + ixx * dub + ixy * dvb;
    const float q2 = iyz + ixy * dub + iyy * dvb;
    const float psid = 0.5f / sqrtf(q0 * q0 + gamma * (q1 * q1 + q2 * q2) + BROX_EPS2);
    const float j11 = psid * (ix * ix + gamma * (ixx * ixx + ixy * ixy));
    const float j12 = psid * (ix * iy + gamma * (ixx * ixy + ixy * iyy));
    const float j22 = psid * (iy * iy + gamma * (ixy * ixy + iyy * iyy));
    const float j13 = psid * (ix * iz + gamma * (ixx * ixz + ixy * iyz));
    const float j23 = psid * (iy * iz + gamma * (ixy * ixz + iyy * iyz));
    const int cy = threadIdx.y + 2, cx = threadIdx.x + 2;   // s_u coordinates
    const float uc = s_u[cy][cx], vc = s_v[cy][cx];
    float su = 0.0f, sv = 0.0f;
    if (x > 0) { su += wl * (s_u[cy][cx - 1] - uc); sv += wl * (s_v[cy][cx - 1] - vc); }
    if (x < w - 1) { su += wr * (s_u[cy][cx + 1] - uc); sv += wr * (s_v[cy][cx + 1] - vc); }
    if (y > 0) { su += wu * (s_u[cy - 1][cx] - uc); sv += wu * (s_v[cy - 1][cx] - vc); }
    if (y < h - 1) { su += wd * (s_u[cy + 1][cx] - uc); sv += wd * (s_v[cy + 1][cx] - vc); }
    const float sw_ = wl + wr + wu + wd;
    p.W[g] = make_float4(wl, wr, wu, wd);
    p.C4[g] = make_float4(j12, su - j13, sv - j23, 1.0f / (j11 + sw_));
    p.C1[g] = 1.0f / (j22 + sw_);
}

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct BroxSorP {
    const float4 *W;
    const float4 *C4;
    const float *C1;
    const float *dui, *dvi;   // increment at the start of this launch's sweeps
    float *duo, *dvo;
    int w, h;
    float omega;
    int nsweeps;
};

// nsweeps (<= SMAX_, shipped 5) red-black SOR sweeps on one tile + (2 x nsweeps + 1)-px halo (temporal blocking).  The sweep is
// bound by shared-memory bandwidth, so only what the sweeps EXCHANGE lives there: the (du, dv) pairs of the region, de-interleaved
// by colour with zero guards (see above).  Everything that is constant over the sweeps of a launch -- the four edge weights and
// the 2x2 system (j12, b1, b2, 1/d1, 1/d2) of every owned pixel -- is loaded from global memory (L2 hits: k_brox_system has just
// written it) straight into registers: 9 floats x <= 2 x SOR_M owned pixels, BROX_SOR_NT = 512 threads of <= 128 registers.
// A pixel update then costs four LDS.64 + one STS.64 (40 B of shared-memory traffic instead of 76).
constexpr int BROX_SOR_NT = 512;
template <class T>
__global__ void __launch_bounds__(BROX_SOR_NT, 1) k_brox_sor(BroxSorP p)
{
    constexpr int PW = T::PW, PH = T::PH, HW = T::HW, NPC = T::NPC, G = T::G, NPCP = T::NPCP;
    constexpr int M = (NPC + BROX_SOR_NT - 1) / BROX_SOR_NT;
    extern __shared__ float4 sm4[];
    float2 *s_uv = (float2 *)sm4 + G;                               // [2][NPCP] (du, dv) by colour, zero guards
    const int w = p.w, h = p.h;
    const int gx0 = blockIdx.x * T::TW, gy0 = blockIdx.y * T::TH;
    const int ox = gx0 - T::R, oy = gy0 - T::R;                     // ox + oy is even: local colour == global colour
    const int tid = threadIdx.x;
    const int ns2 = 2 * p.nsweeps;
    const float omega = p.omega, om1 = 1.0f - p.omega;
    const unsigned thr0 = (unsigned)(T::R - ns2);
#ifdef SINDYN_BROX_PHASE_CLOCKS
    const bool prof_on = threadIdx.x == 0 && T::TW == 32 && blockIdx.x == gridDim.x / 2 && blockIdx.y == gridDim.y / 2;
    long long t_prev = clock64();
    if (prof_on) atomicAdd(&g_brox_clk[15], 1ull);
#endif
    // guards
    for (int r = tid; r < 4 * G; r += BROX_SOR_NT) {
        const int a = r / (2 * G), o = r - a * 2 * G;
        const int off = o < G ? o - G : NPC + (o - G);
        s_uv[a * NPCP + off] = make_float2(0.0f, 0.0f);
    }
    // ---- owned pixels: table (independent of the previous kernel's output: runs under its tail, see pdl_wait)
    unsigned pk[2][M];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const int q = tid + BROX_SOR_NT * m;
            unsigned v = 0;
            if (q < NPC) {
                const int ly = q / HW, i = q - ly * HW;
                const int par = (ly + c) & 1, lx = 2 * i + par;
                const int x = ox + lx, y = oy + ly;
                const bool inside = x >= 0 && x < w && y >= 0 && y < h;
                int dist = 255;
                if (ox > 0) dist = min(dist, lx);
                if (oy > 0) dist = min(dist, ly);
                if (ox + PW < w) dist = min(dist, PW - 1 - lx);
                if (oy + PH < h) dist = min(dist, PH - 1 - ly);
                const bool interior = inside && x >= gx0 && x < gx0 + T::TW && y >= gy0 && y < gy0 + T::TH;
                const bool live = inside && dist >= (int)thr0 + 1;   // updated by at least the first half-sweep
                v = (unsigned)q | ((unsigned)par << 12) | ((unsigned)interior << 13) | ((unsigned)live << 14) | ((unsigned)dist << 16);
            }
            pk[c][m] = v;
        }
    pdl_wait();
    pdl_trigger();
    BROX_CLK(0)
    // ---- stage (du, dv) of the whole region with asynchronous global -> shared copies (rows over warps, columns over lanes)
    for (int ly = tid >> 5; ly < PH; ly += BROX_SOR_NT / 32) {
        const int y = oy + ly;
        const bool row_in = y >= 0 && y < h;
        for (int lx = tid & 31; lx < PW; lx += 32) {
            const int x = ox + lx;
            const int ci = ((lx + ly) & 1) * NPCP + ly * HW + (lx >> 1);
            if (row_in && x >= 0 && x < w) {
                const int g = y * w + x;
                cp_async4(&s_uv[ci].x, p.dui + g);
                cp_async4(&s_uv[ci].y, p.dvi + g);
            } else {
                s_uv[ci] = make_float2(0.0f, 0.0f);
            }
        }
    }
    // ---- the systems and the edge weights of the owned pixels: global -> registers (a weight towards a pixel outside the image
    // is zero: Neumann boundary; k_brox_system stores all four weights of a pixel)
    float rdu[2][M], rdv[2][M], wl[2][M], wr[2][M], wu[2][M], wd[2][M], cj[2][M], cb1[2][M], cb2[2][M], cd1[2][M], cd2[2][M];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const unsigned k = pk[c][m];
            wl[c][m] = wr[c][m] = wu[c][m] = wd[c][m] = 0.0f;
            cj[c][m] = cb1[c][m] = cb2[c][m] = cd1[c][m] = cd2[c][m] = 0.0f;
            if (!((k >> 14) & 1u)) continue;
            const int idx = k & 0xfff, par = (k >> 12) & 1;
            const int ly = idx / HW, lx = 2 * (idx - ly * HW) + par;
            const int x = ox + lx, y = oy + ly;
            const int g = y * w + x;
            const float4 W0 = p.W[g];
            const float4 c4 = p.C4[g];
            cd2[c][m] = p.C1[g];
            wl[c][m] = W0.x; wr[c][m] = W0.y; wu[c][m] = W0.z; wd[c][m] = W0.w;
            cj[c][m] = c4.x; cb1[c][m] = c4.y; cb2[c][m] = c4.z; cd1[c][m] = c4.w;
        }
    cp_async_wait_all();
    __syncthreads();
    BROX_CLK(1)
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const unsigned k = pk[c][m];
            float2 own = make_float2(0.0f, 0.0f);
            if ((k >> 14) & 1u) own = s_uv[c * NPCP + (k & 0xfff)];
            rdu[c][m] = own.x;
            rdv[c][m] = own.y;
        }
#define BROX_HALF(C, K)                                                                                           \
    {                                                                                                             \
        const float2 *__restrict__ uo = s_uv + ((C) ^ 1) * NPCP;                                                   \
        float2 *__restrict__ uc_ = s_uv + (C) * NPCP;                                                              \
        _Pragma("unroll") for (int m = 0; m < M; ++m)                                                             \
        {                                                                                                         \
            const unsigned k_ = pk[C][m];                                                                         \
            if (((k_ >> 14) & 1u) && (k_ >> 16) >= thr0 + (unsigned)(K)) {                                        \
                const int idx = k_ & 0xfff, par = (k_ >> 12) & 1;                                                 \
                const float2 l = uo[idx - 1 + par], r = uo[idx + par], u_ = uo[idx - HW], d = uo[idx + HW];       \
                const float su = wl[C][m] * l.x + wr[C][m] * r.x + wu[C][m] * u_.x + wd[C][m] * d.x;              \
                const float sv = wl[C][m] * l.y + wr[C][m] * r.y + wu[C][m] * u_.y + wd[C][m] * d.y;              \
                const float du_new = om1 * rdu[C][m] + omega * (cb1[C][m] - cj[C][m] * rdv[C][m] + su) * cd1[C][m]; \
                const float dv_new = om1 * rdv[C][m] + omega * (cb2[C][m] - cj[C][m] * du_new + sv) * cd2[C][m];  \
                rdu[C][m] = du_new;                                                                               \
                rdv[C][m] = dv_new;                                                                               \
                uc_[idx] = make_float2(du_new, dv_new);                                                           \
            }                                                                                                     \
        }                                                                                                         \
        __syncthreads();                                                                                          \
    }
    for (int sw = 0; sw < p.nsweeps; ++sw) {
        BROX_HALF(0, 2 * sw + 1)
        BROX_HALF(1, 2 * sw + 2)
    }
#undef BROX_HALF
    BROX_CLK(2)
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const unsigned k = pk[c][m];
            if (!((k >> 13) & 1u)) continue;
            const int idx = k & 0xfff, par = (k >> 12) & 1;
            const int ly = idx / HW, lx = 2 * (idx - ly * HW) + par;
            const int g = (oy + ly) * w + ox + lx;
            p.duo[g] = rdu[c][m];
            p.dvo[g] = rdv[c][m];
        }
    BROX_CLK(3)
}

// level -> finer level: (u+du, v+dv) bilinear, scaled by the size ratios
__global__ void k_brox_prolong(const float *__restrict__ u, const float *__restrict__ v, const float *__restrict__ du,
                               const float *__restrict__ dv, int sw, int sh, float *__restrict__ u2, float *__restrict__ v2,
                               int dw, int dh, float fx, float fy, float mulx, float muly)
{
    pdl_wait();
    pdl_trigger();
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
    pdl_wait();
    pdl_trigger();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = make_float2(sign * (u[i] + du[i]), sign * (v[i] + dv[i]));
}

// ------------------------------------------------------------------ host side
// shared memory of k_brox_level for a w x h level
static size_t brox_level_smem(int w, int h)
{
    const size_t pw = (size_t)(w + 2) & ~(size_t)1, hw = pw / 2, npc = hw * h, npcp = npc + 2 * (hw + 1);
    return 32 * npcp + 40 * npc + 8 * pw * h;
}
constexpr int BROX_LEVEL_MAX_PX = 2100;

constexpr size_t BROX_SMEM_MAX = 227 * 1024;
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
    SD_CHECK(ctx->dalloc(&b->sysW, n));
    SD_CHECK(ctx->dalloc(&b->sysC4, n));
    SD_CHECK(ctx->dalloc(&b->sysC1, n));
    CU_CHECK(ctx, cudaFuncSetAttribute(k_brox_sor<BroxTileS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BroxTileS::SMEM_SOR));
    CU_CHECK(ctx, cudaFuncSetAttribute(k_brox_sor<BroxTile<16, 12>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BroxTile<16, 12>::SMEM_SOR));
    CU_CHECK(ctx, cudaFuncSetAttribute(k_brox_sor<BroxTile<24, 16>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BroxTile<24, 16>::SMEM_SOR));
    CU_CHECK(ctx, cudaFuncSetAttribute(k_brox_sor<BroxTileL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BroxTileL::SMEM_SOR));
    CU_CHECK(ctx, cudaFuncSetAttribute(k_brox_level<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BROX_SMEM_MAX));
    CU_CHECK(ctx, cudaFuncSetAttribute(k_brox_level<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BROX_SMEM_MAX));
    return SINDYN_OK;
}

template <class T> static bool brox_tile_fits(int w, int h) { return cdiv(w, T::TW) * cdiv(h, T::TH) <= SINDYN_NUM_SMS_B200; }
template <class T> static void brox_launch_sor(sindyn_base *ctx, BroxSorP &p)
{
    LAUNCH_PDL(ctx, k_brox_sor<T>, dim3(cdiv(p.w, T::TW), cdiv(p.h, T::TH)), BROX_SOR_NT, T::SMEM_SOR, p);
}

static int brox_enqueue(sindyn_base *ctx, BroxSolver *b, const float *I0, const float *I1, float *flow_out, float sign)
{
    const dim3 blk(32, 8);
    // pyramids
    for (int k = 1; k < b->nl; ++k) {
        const float *s0 = k == 1 ? I0 : b->pyr0 + b->off[k - 1], *s1 = k == 1 ? I1 : b->pyr1 + b->off[k - 1];
        dim3 grd(cdiv(b->ws[k], 32), cdiv(b->hs[k], 8), 2);
        LAUNCH_PDL(ctx, k_brox_pyr_down, grd, blk, 0, s0, s1, b->ws[k - 1], b->hs[k - 1], b->pyr0 + b->off[k], b->pyr1 + b->off[k],
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
        LAUNCH_PDL(ctx, k_brox_warp_deriv, dim3(cdiv(w, BWD_W), cdiv(h, BWD_H)), dim3(BWD_W, BWD_H), 0, L0, L1, b->u[cur], b->v[cur], w, h, b->A, b->Iz,
                   b->Ix, b->Iy, b->Ixz, b->Iyz, b->Ixx, b->Ixy, b->Iyy, b->du[base], b->dv[base]);
        BroxInnerP p;
        p.Ix = b->Ix; p.Iy = b->Iy; p.Iz = b->Iz; p.Ixx = b->Ixx; p.Ixy = b->Ixy; p.Iyy = b->Iyy; p.Ixz = b->Ixz; p.Iyz = b->Iyz;
        p.u = b->u[cur]; p.v = b->v[cur];
        p.w = w; p.h = h; p.alpha = b->alpha; p.gamma = b->gamma; p.omega = b->omega;
        // One CTA is faster than a grid while the level is small: above BROX_LEVEL_MAX_PX the tiled path (system + SOR
        // launch per inner iteration, spread over up to 148 SMs) wins over 10 inner iterations serialised on one SM.
        const int npc = h * (w / 2 + 1);   // pixels per colour in k_brox_level's packed layout
        if (w * h <= BROX_LEVEL_MAX_PX && npc <= 2 * BROX_NT && brox_level_smem(w, h) <= BROX_SMEM_MAX) {
            const int out = 1;
            p.dub = b->du[base]; p.dvb = b->dv[base];
            p.duo = b->du[out]; p.dvo = b->dv[out];
            p.nsweeps = b->solver; p.n_inner = b->inner;
            const size_t smem = brox_level_smem(w, h);
            if (npc <= BROX_NT) LAUNCH_PDL(ctx, k_brox_level<1>, dim3(1, 1), max(128, (npc + 31) & ~31), smem, p);
            else LAUNCH_PDL(ctx, k_brox_level<2>, dim3(1, 1), ((npc + 1) / 2 + 31) & ~31, smem, p);
            base = out;
        } else {
            const int tile = brox_tile_fits<BroxTileS>(w, h) ? 0 : brox_tile_fits<BroxTile<16, 12>>(w, h) ? 1 : brox_tile_fits<BroxTile<24, 16>>(w, h) ? 2 : 3;
            BroxSysP sp;
            sp.Ix = b->Ix; sp.Iy = b->Iy; sp.Iz = b->Iz; sp.Ixx = b->Ixx; sp.Ixy = b->Ixy; sp.Iyy = b->Iyy; sp.Ixz = b->Ixz; sp.Iyz = b->Iyz;
            sp.u = b->u[cur]; sp.v = b->v[cur]; sp.W = b->sysW; sp.C4 = b->sysC4; sp.C1 = b->sysC1;
            sp.w = w; sp.h = h; sp.alpha = b->alpha; sp.gamma = b->gamma;
            BroxSorP q;
            q.W = b->sysW; q.C4 = b->sysC4; q.C1 = b->sysC1; q.w = w; q.h = h; q.omega = b->omega;
            for (int it = 0; it < b->inner; ++it) {
                int in = base, remaining = b->solver;
                sp.dub = b->du[base]; sp.dvb = b->dv[base];
                LAUNCH_PDL(ctx, k_brox_system, dim3(cdiv(w, BSY_W), cdiv(h, BSY_H)), dim3(BSY_W, BSY_H), 0, sp);
                while (remaining > 0) {
                    const int chunks_left = cdiv(remaining, tile == 0 ? BROX_SMAX_SMALL : BROX_SMAX);
                    const int ns = cdiv(remaining, chunks_left);   // balanced chunks of <= BROX_SMAX sweeps (10 -> 5 + 5)
                    int out = 0;
                    while (out == base || out == in) ++out;
                    q.dui = b->du[in]; q.dvi = b->dv[in];
                    q.duo = b->du[out]; q.dvo = b->dv[out];
                    q.nsweeps = ns;
                    // measurement hook: every k_brox_sor launch is bracketed by events
                    const bool prof = b->prof_ev && b->prof_n + 2 <= b->prof_cap;
                    if (prof) cudaEventRecord(b->prof_ev[b->prof_n++], ctx->stream);
                    switch (tile) {
                    case 0: brox_launch_sor<BroxTileS>(ctx, q); break;
                    case 1: brox_launch_sor<BroxTile<16, 12>>(ctx, q); break;
                    case 2: brox_launch_sor<BroxTile<24, 16>>(ctx, q); break;
                    default: brox_launch_sor<BroxTileL>(ctx, q); break;
                    }
                    if (prof) { cudaEventRecord(b->prof_ev[b->prof_n++], ctx->stream); b->prof_px += (long long)w * h * ns; }
                    in = out;
                    remaining -= ns;
                }
                base = in;
            }
        }
        if (k > 0) {
            const int fw = b->ws[k - 1], fh = b->hs[k - 1];
            dim3 fgrd(cdiv(fw, 32), cdiv(fh, 8));
            LAUNCH_PDL(ctx, k_brox_prolong, fgrd, blk, 0, b->u[cur], b->v[cur], b->du[base], b->dv[base], w, h, b->u[cur ^ 1],
                   b->v[cur ^ 1], fw, fh, (float)w / (float)fw, (float)h / (float)fh, (float)fw / (float)w, (float)fh / (float)h);
            cur ^= 1;
        } else {
            fin = base;
        }
    }
    int n = b->w * b->h;
    LAUNCH_PDL(ctx, k_brox_final, cdiv(n, 256), 256, 0, b->u[cur], b->v[cur], b->du[fin], b->dv[fin], n, (float2 *)flow_out, sign);
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
