// residual.cu -- ego-motion flow residual, 8-bit normalisation, Otsu / Triangle thresholds with
// the reference's clamp logic, and the low/high error masks.
//
// Replaces ORB_SLAM2/src/DynaDetect.cc:1236-1367 (homography residual loop :1252-1267, cartToPolar
// :1271, minMaxLoc + scale + convertTo :1279-1282, threshold OTSU/TRIANGLE :1284-1285, clamp logic
// :1309-1367).  k_residual_pose is the north_star variant (depth back-projection with an SE(3)
// pose instead of the homography).  No host round trip: max -> histogram -> thresholds -> masks
// are chained on the device (4 launches).
#include "residual.cuh"

#include <float.h>

__device__ __forceinline__ void block_max_to_global(float m, unsigned int *gmax)
{
    // magnitudes are >= 0, so the IEEE bit pattern orders like the value
    __shared__ float smax[32];
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) smax[wid] = m;
    __syncthreads();
    if (wid == 0) {
        int nw = (blockDim.x + 31) >> 5;
        m = lane < nw ? smax[lane] : 0.0f;
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) atomicMax(gmax, __float_as_uint(m));
    }
}

struct Homog { double h[9]; };

// Evaluated without FMA contraction so that the result is bit-identical to the plain IEEE double
// arithmetic of the reference loop (DynaDetect.cc:1260-1261) as restated by the numpy oracle.
__device__ __forceinline__ void homog_flow(const double *h, int col, int row, double &fx2, double &fy2)
{
    double c = (double)col, r = (double)row;
    double den = __dadd_rn(__dadd_rn(__dmul_rn(h[6], c), __dmul_rn(h[7], r)), h[8]);
    double nx = __dadd_rn(__dadd_rn(__dmul_rn(h[0], c), __dmul_rn(h[1], r)), h[2]);
    double ny = __dadd_rn(__dadd_rn(__dmul_rn(h[3], c), __dmul_rn(h[4], r)), h[5]);
    fx2 = __dsub_rn(c, __ddiv_rn(nx, den));
    fy2 = __dsub_rn(r, __ddiv_rn(ny, den));
}
__device__ __forceinline__ float mag2(float rx, float ry)
{
    // cv::magnitude (cartToPolar): sqrt(x*x + y*y); MAG_FMA selects the contraction cv2's SIMD path uses
#ifndef SINDYN_MAG_NOFMA  // cv2 4.13 (AVX2 dispatch) computes v_sqrt(v_muladd(x, x, y*y)): verified bit-exact in oracle tests
    return __fsqrt_rn(__fmaf_rn(rx, rx, __fmul_rn(ry, ry)));
#else
    return __fsqrt_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)));
#endif
}


__global__ void k_residual_h(const float2 *__restrict__ flow, int W, int H, Homog hm, float *__restrict__ mag,
                             unsigned int *__restrict__ gmax)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.0f;
    if (i < W * H) {
        int row = i / W, col = i - row * W;
        double fx2, fy2;
        homog_flow(hm.h, col, row, fx2, fy2);
        float2 f = flow[i];
        float rx = __fsub_rn(f.x, (float)fx2), ry = __fsub_rn(f.y, (float)fy2);
        m = mag2(rx, ry);
        mag[i] = m;
    }
    block_max_to_global(m, gmax);
}

struct PoseIntr { double T[12]; float fx, fy, cx, cy, inv_depth_scale; };

__global__ void k_residual_pose(const float2 *__restrict__ flow, const uint16_t *__restrict__ depth, int W, int H, PoseIntr P,
                                float *__restrict__ mag, unsigned int *__restrict__ gmax)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.0f;
    if (i < W * H) {
        int row = i / W, col = i - row * W;
        uint16_t d = depth[i];
        if (d != 0) {
            double z = (double)((float)d * P.inv_depth_scale);
            double X = ((double)col - P.cx) * z / P.fx, Y = ((double)row - P.cy) * z / P.fy;
            double xo = P.T[0] * X + P.T[1] * Y + P.T[2] * z + P.T[3];
            double yo = P.T[4] * X + P.T[5] * Y + P.T[6] * z + P.T[7];
            double zo = P.T[8] * X + P.T[9] * Y + P.T[10] * z + P.T[11];
            if (zo > 1e-6) {
                double fx2 = col - (P.fx * xo / zo + P.cx);
                double fy2 = row - (P.fy * yo / zo + P.cy);
                float2 f = flow[i];
                float rx = f.x - (float)fx2, ry = f.y - (float)fy2;
                m = mag2(rx, ry);
            }
        }
        mag[i] = m;
    }
    block_max_to_global(m, gmax);
}

// mag * (float)(255/max) -> u8 (round half to even, saturate) + 256-bin histogram
__global__ void k_mag_u8_hist(const float *__restrict__ mag, int n, const unsigned int *__restrict__ gmax,
                              uint8_t *__restrict__ m8, unsigned int *__restrict__ hist)
{
    __shared__ unsigned int sh[256];
    for (int j = threadIdx.x; j < 256; j += blockDim.x) sh[j] = 0;
    __syncthreads();
    const float scale = (float)(255.0 / (double)__uint_as_float(*gmax));
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float v = mag[i] * scale;
        int q = __float2int_rn(v);
        q = min(max(q, 0), 255);
        m8[i] = (uint8_t)q;
        atomicAdd(&sh[q], 1u);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 256; j += blockDim.x)
        if (sh[j]) atomicAdd(&hist[j], sh[j]);
}

__device__ int count_above(const unsigned int *h, float t)
{
    int c = 0;
    for (int i = 0; i < 256; i++) if ((double)i > (double)t) c += (int)h[i];
    return c;
}

// thr[0]=otsu thr[1]=triangle thr[2]=t_low thr[3]=t_high (DynaDetect.cc:1284-1367)
//
// One CTA of 256 threads.  Otsu's class statistics (q1, mu1) are a serial floating-point recurrence with a division per bin,
// so bit-exactness pins them to ONE thread; everything around it is order independent and runs in parallel under that
// chain: the between-class variance of every bin + first-maximum search (all threads, after the chain), and the whole
// Triangle threshold (warp 1: bounds, first maximum, distance arg-max are exact integer-valued computations).
__global__ void __launch_bounds__(256) k_thresholds(const unsigned int *__restrict__ hist, const unsigned int *__restrict__ gmax, int W, int H,
                                                    float *__restrict__ thr)
{
    __shared__ unsigned int sh[256];
    __shared__ double s_q1[256], s_mu1[256];
    __shared__ unsigned char s_ok[256];
    __shared__ double s_sig[8];
    __shared__ int s_idx[8];
    __shared__ double s_mu;
    __shared__ float s_tri;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    sh[tid] = hist[tid];
    __syncthreads();
    const int total = W * H;
    const double scale = 1. / (double)total;
    if (tid == 0) {
        // serial part of cv::threshold(THRESH_OTSU) (getThreshVal_Otsu_8u): q1 / mu1 recurrence
        double mu1 = 0, q1 = 0;
        for (int i = 0; i < 256; i++) {
            const double p_i = sh[i] * scale;
            mu1 *= q1;
            q1 += p_i;
            const double q2 = 1. - q1;
            if (fmin(q1, q2) < FLT_EPSILON || fmax(q1, q2) > 1. - FLT_EPSILON) { s_ok[i] = 0; continue; }
            mu1 = (mu1 + i * p_i) / q1;
            s_q1[i] = q1; s_mu1[i] = mu1; s_ok[i] = 1;
        }
    } else if (wid == 1) {
        // cv::threshold(THRESH_TRIANGLE) (getThreshVal_Triangle_8u); lane l owns bins 8l .. 8l+7
        int first = 256, last = 0, mx = 0, mi = 0;
        for (int k = 0; k < 8; ++k) {
            const int i = 8 * lane + k, v = (int)sh[i];
            if (v > 0) { first = min(first, i); if (i > 0) last = max(last, i); }
            if (v > mx) { mx = v; mi = i; }
        }
        for (int off = 16; off > 0; off >>= 1) {
            first = min(first, __shfl_xor_sync(0xffffffffu, first, off));
            last = max(last, __shfl_xor_sync(0xffffffffu, last, off));
            const int om = __shfl_xor_sync(0xffffffffu, mx, off), oi = __shfl_xor_sync(0xffffffffu, mi, off);
            if (om > mx || (om == mx && oi < mi)) { mx = om; mi = oi; }   // first index of the maximum
        }
        int left_bound = first < 256 ? first : 0, right_bound = last, max_ind = mi;
        if (left_bound > 0) left_bound--;
        if (right_bound < 255) right_bound++;
        const bool flipped = max_ind - left_bound < right_bound - max_ind;
        if (flipped) { left_bound = 255 - right_bound; max_ind = 255 - max_ind; }
        // distance arg-max over (left_bound, max_ind]: a*i + b*h[i] is integer valued (exact in double), first maximum wins,
        // and it has to exceed 0 to move the threshold off left_bound
        const double a = mx, b2 = left_bound - max_ind;
        double bd = 0;
        int bi = -1;
        for (int k = 0; k < 8; ++k) {
            const int i = 8 * lane + k;
            if (i > left_bound && i <= max_ind) {
                const double t = a * i + b2 * (double)(int)sh[flipped ? 255 - i : i];
                if (t > bd) { bd = t; bi = i; }
            }
        }
        for (int off = 16; off > 0; off >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (oi >= 0 && (od > bd || (od == bd && (bi < 0 || oi < bi)))) { bd = od; bi = oi; }
        }
        double thresh = bi >= 0 ? bi : left_bound;
        thresh--;
        if (flipped) thresh = 255 - thresh;
        if (lane == 0) s_tri = (float)thresh;
    } else if (wid == 2) {
        // mean of the histogram: integer-valued partial sums, exact in any order
        double m = 0;
        for (int k = 0; k < 8; ++k) m += (8 * lane + k) * (double)sh[8 * lane + k];
        for (int off = 16; off > 0; off >>= 1) m += __shfl_xor_sync(0xffffffffu, m, off);
        if (lane == 0) s_mu = m * scale;
    }
    __syncthreads();
    // between-class variance of every bin, first maximum (it has to exceed 0)
    double sigma = 0;
    int si = -1;
    if (s_ok[tid]) {
        const double q1 = s_q1[tid], mu1 = s_mu1[tid], q2 = 1. - q1;
        const double mu2 = (s_mu - q1 * mu1) / q2;
        const double sg = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
        if (sg > 0) { sigma = sg; si = tid; }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double os = __shfl_xor_sync(0xffffffffu, sigma, off);
        const int oi = __shfl_xor_sync(0xffffffffu, si, off);
        if (oi >= 0 && (os > sigma || (os == sigma && (si < 0 || oi < si)))) { sigma = os; si = oi; }
    }
    if (lane == 0) { s_sig[wid] = sigma; s_idx[wid] = si; }
    __syncthreads();
    if (tid != 0) return;
    for (int k = 1; k < 8; ++k)
        if (s_idx[k] >= 0 && (s_sig[k] > sigma || (s_sig[k] == sigma && (si < 0 || s_idx[k] < si)))) { sigma = s_sig[k]; si = s_idx[k]; }
    const unsigned int *h = sh;
    float thred1 = (float)(si >= 0 ? si : 0);
    float thred2 = s_tri;
    thr[0] = thred1;
    thr[1] = thred2;
    const float maxErrorf = __uint_as_float(*gmax);
    float tl, th;
    if (thred1 < thred2) {
        if (thred1 < 1.7f * 255.0f / maxErrorf) thred1 = 1.7f * 255.0f / maxErrorf;
        else if (thred1 > 3.0f * 255.0f / maxErrorf) thred1 = 3.0f * 255.0f / maxErrorf;
        if ((double)count_above(h, thred1) > 0.5 * W * H) thred1 = thred1 + 0.2f * 255.0f / maxErrorf;
        float m = fmaxf(3.0f * 255.0f / maxErrorf, thred1 * 1.2f);
        if (thred2 < m) thred2 = m;
        else if (thred2 > 10.0f * 255.0f / maxErrorf) thred2 = 10.0f * 255.0f / maxErrorf;
        tl = thred1;
        th = thred2;
    } else {
        if (thred2 < 1.7f * 255.0f / maxErrorf) thred2 = 1.7f * 255.0f / maxErrorf;
        else if (thred2 > 3.0f * 255.0f / maxErrorf) thred2 = 3.0f * 255.0f / maxErrorf;
        // reference quirk: countNonZero(thred2) on the scalar never exceeds 0.5*W*H (DynaDetect.cc:1348)
        float m = fmaxf(3.0f * 255.0f / maxErrorf, thred2 * 1.2f);
        if (thred1 < m) thred1 = m;
        else if (thred1 > 10.0f * 255.0f / maxErrorf) thred1 = 10.0f * 255.0f / maxErrorf;
        tl = thred2;
        th = thred1;
    }
    thr[2] = tl;
    thr[3] = th;
}

__global__ void k_masks(const uint8_t *__restrict__ m8, int n, const float *__restrict__ thr, uint8_t *__restrict__ low,
                        uint8_t *__restrict__ high)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float tl = thr[2], th = thr[3];
    float v = (float)m8[i];
    low[i] = v > tl ? 128 : 0;   // (mask * 0.5) rounds 127.5 to 128 (DynaDetect.cc:1334,1365)
    high[i] = v > th ? 255 : 0;
}

int residual_init(sindyn_base *ctx, ResidualStage *r, int W, int H)
{
    r->W = W; r->H = H;
    SD_CHECK(ctx->dalloc(&r->mag, (size_t)W * H));
    SD_CHECK(ctx->dalloc(&r->m8, (size_t)W * H));
    SD_CHECK(ctx->dalloc(&r->hist, 256 + 1));
    r->gmax = r->hist + 256;
    SD_CHECK(ctx->dalloc(&r->thr, 4));
    return SINDYN_OK;
}

static int residual_tail(sindyn_base *ctx, ResidualStage *r, uint8_t *low, uint8_t *high)
{
    const int n = r->W * r->H;
    LAUNCH(ctx, k_mag_u8_hist, SINDYN_NUM_SMS_B200 * 2, 256, 0, r->mag, n, r->gmax, r->m8, r->hist);
    LAUNCH(ctx, k_thresholds, 1, 256, 0, r->hist, r->gmax, r->W, r->H, r->thr);
    LAUNCH(ctx, k_masks, cdiv(n, 256), 256, 0, r->m8, n, r->thr, low, high);
    LAUNCH_CHECK(ctx);
    return SINDYN_OK;
}

int residual_homography_run(sindyn_base *ctx, ResidualStage *r, const float *flow, const double *Hm, uint8_t *low, uint8_t *high)
{
    const int n = r->W * r->H;
    Homog hm;
    for (int i = 0; i < 9; ++i) hm.h[i] = Hm[i];
    CU_CHECK(ctx, cudaMemsetAsync(r->hist, 0, sizeof(unsigned int) * 257, ctx->stream));
    LAUNCH(ctx, k_residual_h, cdiv(n, 256), 256, 0, (const float2 *)flow, r->W, r->H, hm, r->mag, r->gmax);
    return residual_tail(ctx, r, low, high);
}

__global__ void k_residual_h_dev(const float2 *__restrict__ flow, int W, int H, const double *__restrict__ Hd,
                                 float *__restrict__ mag, unsigned int *__restrict__ gmax)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.0f;
    if (i < W * H) {
        int row = i / W, col = i - row * W;
        double fx2, fy2;
        homog_flow(Hd, col, row, fx2, fy2);
        float2 f = flow[i];
        float rx = __fsub_rn(f.x, (float)fx2), ry = __fsub_rn(f.y, (float)fy2);
        m = mag2(rx, ry);
        mag[i] = m;
    }
    block_max_to_global(m, gmax);
}

// H resident on the device (written by the homography kernel): no host round trip in the fused path
int residual_homography_run_dev(sindyn_base *ctx, ResidualStage *r, const float *flow, const double *H_dev, uint8_t *low,
                                uint8_t *high)
{
    const int n = r->W * r->H;
    CU_CHECK(ctx, cudaMemsetAsync(r->hist, 0, sizeof(unsigned int) * 257, ctx->stream));
    LAUNCH(ctx, k_residual_h_dev, cdiv(n, 256), 256, 0, (const float2 *)flow, r->W, r->H, H_dev, r->mag, r->gmax);
    return residual_tail(ctx, r, low, high);
}

int residual_pose_run(sindyn_base *ctx, ResidualStage *r, const float *flow, const uint16_t *depth, const double *T,
                      float fx, float fy, float cx, float cy, float depth_scale, uint8_t *low, uint8_t *high)
{
    const int n = r->W * r->H;
    PoseIntr P;
    for (int i = 0; i < 12; ++i) P.T[i] = T[i];
    P.fx = fx; P.fy = fy; P.cx = cx; P.cy = cy;
    P.inv_depth_scale = 1.0f / depth_scale;
    CU_CHECK(ctx, cudaMemsetAsync(r->hist, 0, sizeof(unsigned int) * 257, ctx->stream));
    LAUNCH(ctx, k_residual_pose, cdiv(n, 256), 256, 0, (const float2 *)flow, depth, r->W, r->H, P, r->mag, r->gmax);
    return residual_tail(ctx, r, low, high);
}
