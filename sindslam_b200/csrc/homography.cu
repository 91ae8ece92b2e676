// homography.cu -- weighted grid sampling of the dense flow and robust homography estimation.
//
// Replaces ORB_SLAM2/src/DynaDetect.cc:1163-1235:
//   clusterWeight (:1169-1177), the 10-px sample grid with weight = RNG(12345).gaussian(0.5) + class term
//   (:1182-1204), std::sort by weight descending (:1217-1219), the inBorder filter with its int truncation and
//   inclusive upper bound (:1221-1231, inBorder :103-106), and cv::findHomography(pts, ptsLast, noArray(), RHO)
//   (:1235).
// The sample list is bit-exact against the oracle; the homography is bit-identical to the library's (k_rho below).
#include "homography.cuh"

#include <math_constants.h>

__device__ const float c_gauss[HG_GAUSS_N] = {
#include "gauss_table.inc"
};

// counts[0][i] = |label_last == i|, counts[1][i] = |label_last == i && dyna_last == 255|
__global__ void k_cluster_weight_counts(const uint8_t *__restrict__ label_last, const uint8_t *__restrict__ dyna_last, int n,
                                        int *__restrict__ counts)
{
    __shared__ int sc[2 * 16];
    if (threadIdx.x < 32) sc[threadIdx.x] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int l = label_last[i];
        if (l >= 1 && l < 12) {
            atomicAdd(&sc[l], 1);
            if (dyna_last[i] == 255) atomicAdd(&sc[16 + l], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x < 32 && sc[threadIdx.x]) atomicAdd(&counts[threadIdx.x], sc[threadIdx.x]);
}

__device__ __forceinline__ unsigned int f2ord_desc(float f)
{
    unsigned int u = __float_as_uint(f);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending order of f
    return ~u;                                        // descending
}

// single CTA: weights, stable descending sort, inBorder filter, ordered compaction
__global__ void __launch_bounds__(1024) k_sample_pairs(const float2 *__restrict__ flow, const uint8_t *__restrict__ label_last,
                                                       const uint8_t *__restrict__ dyna_last, const int *__restrict__ counts, int W, int H,
                                                       float2 *__restrict__ pts, float2 *__restrict__ pts_last, int *__restrict__ n_out)
{
    __shared__ unsigned long long keys[HG_MAX_SAMPLES];
    __shared__ float cw[12];
    __shared__ int s_scan[1024 / 32];
    __shared__ int s_base;
    const int ncol = (W - 1) / 10, nrow = (H - 1) / 10;  // rows 10,20,.. < H ; cols 10,20,.. < W
    const int ns = ncol * nrow;
    if (threadIdx.x < 12) {
        int i = threadIdx.x;
        cw[i] = i >= 1 ? (float)counts[16 + i] / ((float)counts[i] + 1.0f) : 0.0f;
    }
    __syncthreads();
    // Bitonic sort of HG_MAX_SAMPLES (weight desc, index) keys, 4 consecutive keys per thread: compare-exchange distances 1-2
    // stay in the thread's registers, 4-64 are warp shuffles, only distances >= 128 go through shared memory (15 of the 78
    // stages).  The keys are unique (the index is part of the key), so the order equals std::sort's on distinct weights and
    // the stable order on equal ones.
    static_assert(HG_MAX_SAMPLES == 4096, "4 keys per thread x 1024 threads");
    unsigned long long e[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int k = 4 * threadIdx.x + m;
        unsigned long long key = ~0ull;
        if (k < ns) {
            int r = k / ncol, c = k - r * ncol;
            int row = 10 + 10 * r, col = 10 + 10 * c;
            float randomd = c_gauss[k];
            uint8_t dl = dyna_last[row * W + col];
            float w;
            if (dl < 20) w = randomd + 1.0f;
            else if ((unsigned)(dl - 20) <= 230u - 20u) w = randomd + 1.2f * (1.0f - cw[min((int)label_last[row * W + col], 11)]);
            else w = randomd + 0.4f;
            key = ((unsigned long long)f2ord_desc(w) << 32) | (unsigned)k;
        }
        e[m] = key;
    }
    for (int k = 2; k <= HG_MAX_SAMPLES; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 128) {
#pragma unroll
                for (int m = 0; m < 4; ++m) keys[4 * threadIdx.x + m] = e[m];
                __syncthreads();
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const int i = 4 * threadIdx.x + m;
                    const unsigned long long o = keys[i ^ j];
                    const bool take_min = ((i & j) == 0) == ((i & k) == 0);
                    e[m] = take_min ? (o < e[m] ? o : e[m]) : (o > e[m] ? o : e[m]);
                }
                __syncthreads();
            } else if (j >= 4) {
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const int i = 4 * threadIdx.x + m;
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, e[m], j >> 2);
                    const bool take_min = ((i & j) == 0) == ((i & k) == 0);
                    e[m] = take_min ? (o < e[m] ? o : e[m]) : (o > e[m] ? o : e[m]);
                }
            } else {
                // (distances 1 and 2 spelled out: a run-time register index would push e[] into local memory)
#define HG_CE(A, B)                                                                      \
    {                                                                                    \
        const bool up = ((4 * threadIdx.x + (A)) & k) == 0;                              \
        const unsigned long long a_ = e[A], b_ = e[B];                                   \
        if ((a_ > b_) == up) { e[A] = b_; e[B] = a_; }                                   \
    }
                if (j == 2) { HG_CE(0, 2) HG_CE(1, 3) }
                else { HG_CE(0, 1) HG_CE(2, 3) }
#undef HG_CE
            }
        }
#pragma unroll
    for (int m = 0; m < 4; ++m) keys[4 * threadIdx.x + m] = e[m];
    __syncthreads();
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int start = 0; start < ns; start += blockDim.x) {
        int i = start + threadIdx.x;
        bool keep = false;
        float pc = 0.f, pr = 0.f, lx = 0.f, ly = 0.f;
        if (i < ns) {
            int k = (int)(keys[i] & 0xffffffffull);
            int r = k / ncol, c = k - r * ncol;
            int row = 10 + 10 * r, col = 10 + 10 * c;
            float2 f = flow[row * W + col];
            pc = (float)col; pr = (float)row;
            lx = pc - f.x; ly = pr - f.y;
            int irow = (int)ly, icol = (int)lx;  // float -> int truncation at the inBorder call (DynaDetect.cc:1226)
            keep = ((unsigned)irow <= (unsigned)H) && ((unsigned)icol <= (unsigned)W);
        }
        unsigned bal = __ballot_sync(0xffffffffu, keep);
        int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) s_scan[wid] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w2 = 0; w2 < wid; ++w2) off += s_scan[w2];
        off += __popc(bal & ((1u << lane) - 1u));
        if (keep) { pts[off] = make_float2(pc, pr); pts_last[off] = make_float2(lx, ly); }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) t += s_scan[w2];
            s_base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = s_base;
}

// ---------------------------------------------------------------- cv::findHomography(..., RHO)
// OpenCV's RHO estimator (calib3d rho.cpp, class RHO_HEST_REFC as driven by fundam.cpp: reprojection threshold 3 px,
// 2000 iterations, confidence 0.995, beta 0.35, non-randomness criterion + final refinement), restated operation by
// operation so that H is BIT-IDENTICAL to the library's on the same ordered sample list (oracle/rho_cpu.c is the same
// restatement in C, pinned against cv2.findHomography(RHO) in tests/test_rho_cpu.py; this file is compiled without FMA
// contraction, the float expressions keep the association of the C code).
//
// The estimator is a serial chain: every hypothesis depends on the SPRT / PROSAC state the previous ones left behind.  One
// CTA runs it; thread 0 is the control thread (PROSAC phase + xorshift128+ sampling, degeneracy tests, the unrolled
// 4-point solve, SPRT bookkeeping, non-randomness optimisation, the 8x8 Cholesky of the refinement) and the data-parallel
// parts are spread over the CTA:
//   * model evaluation: all N reprojection tests at once (one ballot word per warp); the sequential probability ratio
//     lambda is replayed exactly point by point only while a decision can depend on it -- once an upper bound on lambda is
//     so small that even a word of 32 rejections cannot lift it over the threshold A, whole words are skipped with the
//     bound alone (if a later word is not safe by the bound, the exact product is replayed from the last exact checkpoint);
//   * the inlier buffers keep the library's stale-tail behaviour (an evaluation overwrites only the points it tested);
//   * refinement: the float accumulations of J^T J, J^T e and the squared error run strictly in inlier order (one thread
//     per accumulator over per-point products staged in shared memory), because float addition is order dependent.
#define RHO_NT 512
#define RHO_TILE2 256            // inliers per product tile of the refinement (two tiles: double buffered)
#define RHO_NE (RHO_NT - 32)     // threads that evaluate a model (warps 1 .. 15); warp 0 holds the control thread
#define RHO_ACC 36               // the 27 entries of JtJ the library updates (lower triangle without the zero block), 8 of Jte, S
#define RHO_WORDS (HG_MAX_SAMPLES / 32)
#define RHO_SMEM (sizeof(float2) * 2 * HG_MAX_SAMPLES + sizeof(unsigned short) * (2 * HG_MAX_SAMPLES + 8) + sizeof(unsigned) * 3 * RHO_WORDS + \
                  sizeof(float) * 2 * RHO_ACC * (RHO_TILE2 + 1))

struct RhoPrng { unsigned long long s0, s1; };
__device__ __forceinline__ double rho_random(RhoPrng &g)
{
    unsigned long long x = g.s0;
    const unsigned long long y = g.s1;
    x ^= x << 23;
    x ^= x >> 17;
    x ^= y ^ (y >> 26);
    g.s0 = y;
    g.s1 = x;
    const unsigned long long s = x + y;
    return (double)s * 5.421010862427522e-20;   // 2^-64
}

__device__ void rho_rnd_smpl(RhoPrng &g, unsigned sampleSize, unsigned *cs, unsigned dataSetSize)
{
    unsigned i, j;
    if (sampleSize * 2 > dataSetSize) {          // selection sampling (Knuth, Algorithm S)
        for (i = 0, j = 0; i < dataSetSize && j < sampleSize; i++) {
            const double U = rho_random(g);
            if ((double)(dataSetSize - i) * U < (double)(sampleSize - j)) cs[j++] = i;
        }
    } else {                                     // draw until distinct
        for (i = 0; i < sampleSize; i++) {
            bool inList;
            do {
                cs[i] = (unsigned)((double)dataSetSize * rho_random(g));
                inList = false;
                for (j = 0; j < i; j++)
                    if (cs[i] == cs[j]) { inList = true; break; }
            } while (inList);
        }
    }
}

__device__ unsigned rho_iter_bound(double confidence, double inlierRate, unsigned maxIterBound)
{
    unsigned retVal;
    const double q = 1. - pow(inlierRate, 4.0);
    if (q >= 1.) retVal = maxIterBound;
    else if (q <= 0.) retVal = 1;
    else retVal = (unsigned)ceil(log(1. - confidence) / log(q));
    return retVal <= maxIterBound ? retVal : maxIterBound;
}

__device__ double rho_design_sprt(double delta, double epsilon)
{
    const double C = (1 - delta) * log((1 - delta) / (1 - epsilon)) + delta * log(delta / epsilon);
    const double K = 25.0 * C / 1.0 + 1;        // t_M = 25, m_S = 1
    double An = K, prevAn;
    unsigned i = 0;
    do {
        prevAn = An;
        An = K + log(An);
    } while ((An - prevAn > 1.5e-8) && (++i < 10));
    return An;
}

// strong geometric constraint + coincident coordinates (RHO_HEST_REFC::isSampleDegenerate); k = 4 src points, 4 dst points
__device__ bool rho_sample_degenerate(const float2 *k)
{
    if (k[0].x == k[1].x || k[1].x == k[2].x || k[2].x == k[3].x || k[0].x == k[2].x || k[0].x == k[3].x || k[1].x == k[3].x ||
        k[0].y == k[1].y || k[1].y == k[2].y || k[2].y == k[3].y || k[0].y == k[2].y || k[0].y == k[3].y || k[1].y == k[3].y ||
        k[4].x == k[5].x || k[5].x == k[6].x || k[6].x == k[7].x || k[4].x == k[6].x || k[4].x == k[7].x || k[5].x == k[7].x ||
        k[4].y == k[5].y || k[5].y == k[6].y || k[6].y == k[7].y || k[4].y == k[6].y || k[4].y == k[7].y || k[5].y == k[7].y)
        return true;
    const float c0s0 = k[0].y - k[1].y, c0s1 = k[1].x - k[0].x, c0s2 = k[0].x * k[1].y - k[0].y * k[1].x;
    const float dots0 = c0s0 * k[2].x + c0s1 * k[2].y + c0s2;
    const float c0d0 = k[4].y - k[5].y, c0d1 = k[5].x - k[4].x, c0d2 = k[4].x * k[5].y - k[4].y * k[5].x;
    const float dotd0 = c0d0 * k[6].x + c0d1 * k[6].y + c0d2;
    if (((int)dots0 ^ (int)dotd0) < 0) return true;
    const float dots1 = c0s0 * k[3].x + c0s1 * k[3].y + c0s2;
    const float dotd1 = c0d0 * k[7].x + c0d1 * k[7].y + c0d2;
    if (((int)dots1 ^ (int)dotd1) < 0) return true;
    const float c2s0 = k[2].y - k[3].y, c2s1 = k[3].x - k[2].x, c2s2 = k[2].x * k[3].y - k[2].y * k[3].x;
    const float dots2 = c2s0 * k[0].x + c2s1 * k[0].y + c2s2;
    const float c2d0 = k[6].y - k[7].y, c2d1 = k[7].x - k[6].x, c2d2 = k[6].x * k[7].y - k[6].y * k[7].x;
    const float dotd2 = c2d0 * k[4].x + c2d1 * k[4].y + c2d2;
    if (((int)dots2 ^ (int)dotd2) < 0) return true;
    const float dots3 = c2s0 * k[1].x + c2s1 * k[1].y + c2s2;
    const float dotd3 = c2d0 * k[5].x + c2d1 * k[5].y + c2d2;
    if (((int)dots3 ^ (int)dotd3) < 0) return true;
    return false;
}

// hFuncRefC: hand-unrolled Gaussian elimination of the 8x9 system of a 4-point homography, all in float
__device__ void rho_h_func(const float2 *k, float *H)
{
    const float x0 = k[0].x, y0 = k[0].y, x1 = k[1].x, y1 = k[1].y, x2 = k[2].x, y2 = k[2].y, x3 = k[3].x, y3 = k[3].y;
    const float X0 = k[4].x, Y0 = k[4].y, X1 = k[5].x, Y1 = k[5].y, X2 = k[6].x, Y2 = k[6].y, X3 = k[7].x, Y3 = k[7].y;
    const float x0X0 = x0 * X0, x1X1 = x1 * X1, x2X2 = x2 * X2, x3X3 = x3 * X3;
    const float x0Y0 = x0 * Y0, x1Y1 = x1 * Y1, x2Y2 = x2 * Y2, x3Y3 = x3 * Y3;
    const float y0X0 = y0 * X0, y1X1 = y1 * X1, y2X2 = y2 * X2, y3X3 = y3 * X3;
    const float y0Y0 = y0 * Y0, y1Y1 = y1 * Y1, y2Y2 = y2 * Y2, y3Y3 = y3 * Y3;
    float n00 = x0 - x2, n01 = x1 - x2, n02 = x2, n03 = x3 - x2;
    float n10 = y0 - y2, n11 = y1 - y2, n12 = y2, n13 = y3 - y2;
    float a0 = x2X2 - x0X0, a1 = x2X2 - x1X1, a2 = -x2X2, a3 = x2X2 - x3X3, a4 = x2Y2 - x0Y0, a5 = x2Y2 - x1Y1, a6 = -x2Y2, a7 = x2Y2 - x3Y3;
    float b0 = y2X2 - y0X0, b1 = y2X2 - y1X1, b2 = -y2X2, b3 = y2X2 - y3X3, b4 = y2Y2 - y0Y0, b5 = y2Y2 - y1Y1, b6 = -y2Y2, b7 = y2Y2 - y3Y3;
    float c0 = (X0 - X2), c1 = (X1 - X2), c2 = (X2), c3 = (X3 - X2), c4 = (Y0 - Y2), c5 = (Y1 - Y2), c6 = (Y2), c7 = (Y3 - Y2);
    float s1 = n00, s2 = n01;
    n11 = n11 * s1 - n10 * s2;
    a1 = a1 * s1 - a0 * s2; b1 = b1 * s1 - b0 * s2; c1 = c1 * s1 - c0 * s2;
    a5 = a5 * s1 - a4 * s2; b5 = b5 * s1 - b4 * s2; c5 = c5 * s1 - c4 * s2;
    s2 = n03;
    n13 = n13 * s1 - n10 * s2;
    a3 = a3 * s1 - a0 * s2; b3 = b3 * s1 - b0 * s2; c3 = c3 * s1 - c0 * s2;
    a7 = a7 * s1 - a4 * s2; b7 = b7 * s1 - b4 * s2; c7 = c7 * s1 - c4 * s2;
    s1 = n11; s2 = n13;
    a3 = a3 * s1 - a1 * s2; b3 = b3 * s1 - b1 * s2; c3 = c3 * s1 - c1 * s2;
    a7 = a7 * s1 - a5 * s2; b7 = b7 * s1 - b5 * s2; c7 = c7 * s1 - c5 * s2;
    s2 = n10;
    n00 = n00 * s1 - n01 * s2;
    a0 = a0 * s1 - a1 * s2; b0 = b0 * s1 - b1 * s2; c0 = c0 * s1 - c1 * s2;
    a4 = a4 * s1 - a5 * s2; b4 = b4 * s1 - b5 * s2; c4 = c4 * s1 - c5 * s2;
    s1 = 1.0f / n00;
    a0 *= s1; b0 *= s1; c0 *= s1; a4 *= s1; b4 *= s1; c4 *= s1;
    s1 = 1.0f / n11;
    a1 *= s1; b1 *= s1; c1 *= s1; a5 *= s1; b5 *= s1; c5 *= s1;
    s1 = n02; s2 = n12;
    a2 -= a0 * s1 + a1 * s2; b2 -= b0 * s1 + b1 * s2; c2 -= c0 * s1 + c1 * s2;
    a6 -= a4 * s1 + a5 * s2; b6 -= b4 * s1 + b5 * s2; c6 -= c4 * s1 + c5 * s2;
    s1 = a7;
    b7 /= s1; c7 /= s1;
    s1 = a0; b0 -= s1 * b7; c0 -= s1 * c7;
    s1 = a1; b1 -= s1 * b7; c1 -= s1 * c7;
    s1 = a2; b2 -= s1 * b7; c2 -= s1 * c7;
    s1 = a3; b3 -= s1 * b7; c3 -= s1 * c7;
    s1 = a4; b4 -= s1 * b7; c4 -= s1 * c7;
    s1 = a5; b5 -= s1 * b7; c5 -= s1 * c7;
    s1 = a6; b6 -= s1 * b7; c6 -= s1 * c7;
    s1 = b3;
    c3 /= s1;
    s1 = b0; c0 -= s1 * c3;
    s1 = b1; c1 -= s1 * c3;
    s1 = b2; c2 -= s1 * c3;
    s1 = b4; c4 -= s1 * c3;
    s1 = b5; c5 -= s1 * c3;
    s1 = b6; c6 -= s1 * c3;
    s1 = b7; c7 -= s1 * c3;
    H[0] = c0; H[1] = c1; H[2] = c2; H[3] = c4; H[4] = c5; H[5] = c6; H[6] = c7; H[7] = c3; H[8] = 1.0f;
}

// sacChol8x8Damped / sacTRInv8x8 / sacTRISolve8x8 (lower triangles only; the association of every product was fixed by
// comparing with the library bit for bit, see oracle/rho_cpu.c)
__device__ bool rho_chol8(const float (*A)[8], float lambda, float (*L)[8])
{
    const float lambdap1 = lambda + 1.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
    #pragma unroll
    for (int j = 0; j < i; j++) {
            float x = A[i][j];
        #pragma unroll
    for (int k = 0; k < j; k++) x -= L[i][k] * L[j][k];
            L[i][j] = x / L[j][j];
        }
        float x = A[i][i] * lambdap1;
    #pragma unroll
    for (int k = 0; k < i; k++) x -= L[i][k] * L[i][k];
        if (x < 0) return false;
        L[i][i] = sqrtf(x);
    }
    return true;
}
__device__ void rho_tr_inv8(float (*M)[8])     // in place: M = L on entry
{
    float s[2][2], t[2][2], u[4][4], v[4][4];
#pragma unroll
    for (int i = 0; i < 8; i++) M[i][i] = 1.0f / M[i][i];
    M[1][0] = -M[1][1] * M[1][0] * M[0][0];
    M[3][2] = -M[3][3] * M[3][2] * M[2][2];
    M[5][4] = -M[5][5] * M[5][4] * M[4][4];
    M[7][6] = -M[7][7] * M[7][6] * M[6][6];
#pragma unroll
    for (int blk = 0; blk < 2; blk++) {
        const int o = 4 * blk;
        s[0][0] = -M[o + 2][o + 2] * M[o + 2][o + 0];
        s[0][1] = -M[o + 2][o + 2] * M[o + 2][o + 1];
        s[1][0] = -M[o + 3][o + 2] * M[o + 2][o + 0] + -M[o + 3][o + 3] * M[o + 3][o + 0];
        s[1][1] = -M[o + 3][o + 2] * M[o + 2][o + 1] + -M[o + 3][o + 3] * M[o + 3][o + 1];
        t[0][0] = s[0][0] * M[o + 0][o + 0] + s[0][1] * M[o + 1][o + 0];
        t[0][1] = s[0][1] * M[o + 1][o + 1];
        t[1][0] = s[1][0] * M[o + 0][o + 0] + s[1][1] * M[o + 1][o + 0];
        t[1][1] = s[1][1] * M[o + 1][o + 1];
        M[o + 2][o + 0] = t[0][0]; M[o + 2][o + 1] = t[0][1]; M[o + 3][o + 0] = t[1][0]; M[o + 3][o + 1] = t[1][1];
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
        u[0][c] = -M[4][4] * M[4][c];
        u[1][c] = -M[5][4] * M[4][c] + -M[5][5] * M[5][c];
        u[2][c] = -M[6][4] * M[4][c] + -M[6][5] * M[5][c] + -M[6][6] * M[6][c];
        u[3][c] = -M[7][4] * M[4][c] + -M[7][5] * M[5][c] + -M[7][6] * M[6][c] + -M[7][7] * M[7][c];
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        v[r][0] = u[r][0] * M[0][0] + u[r][1] * M[1][0] + u[r][2] * M[2][0] + u[r][3] * M[3][0];
        v[r][1] = u[r][1] * M[1][1] + u[r][2] * M[2][1] + u[r][3] * M[3][1];
        v[r][2] = u[r][2] * M[2][2] + u[r][3] * M[3][2];
        v[r][3] = u[r][3] * M[3][3];
    }
#pragma unroll
    for (int r = 0; r < 4; r++)
    #pragma unroll
    for (int c = 0; c < 4; c++) M[4 + r][c] = v[r][c];
}
__device__ void rho_tri_solve8(const float (*L)[8], const float *Jte, float *dH)
{
    float t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float v = L[i][0] * Jte[0];
    #pragma unroll
    for (int k = 1; k <= i; k++) v += L[i][k] * Jte[k];
        t[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float v = L[i][i] * t[i];
    #pragma unroll
    for (int k = i + 1; k < 8; k++) v += L[k][i] * t[k];
        dH[i] = v;
    }
}

// accumulator a of sacCalcJacobianErrors <-> (row, column) of JtJ (lower triangle, in the library's update order), Jte, S
__constant__ signed char c_rho_acc_r[27] = {0, 1, 1, 2, 2, 2, 3, 4, 4, 5, 5, 5, 6, 6, 6, 6, 6, 6, 6, 7, 7, 7, 7, 7, 7, 7, 7};
__constant__ signed char c_rho_acc_c[27] = {0, 0, 1, 0, 1, 2, 3, 3, 4, 3, 4, 5, 0, 1, 2, 3, 4, 5, 6, 0, 1, 2, 3, 4, 5, 6, 7};

__device__ __forceinline__ void rho_ebar() { asm volatile("bar.sync 1, %0;" ::"n"(RHO_NE) : "memory"); }    // barrier of the evaluating warps

__global__ void __launch_bounds__(RHO_NT) k_rho(const float2 *__restrict__ g_src, const float2 *__restrict__ g_dst, const int *__restrict__ n_ptr,
                                               double *__restrict__ H_out, int *__restrict__ info_out, unsigned char *__restrict__ mask_out)
{
    extern __shared__ unsigned char rho_sm[];
    float2 *s_src = (float2 *)rho_sm, *s_dst = s_src + HG_MAX_SAMPLES;
    unsigned short *s_tbl = (unsigned short *)(s_dst + HG_MAX_SAMPLES);     // non-randomness table, N + 1 entries
    unsigned short *s_idx = s_tbl + HG_MAX_SAMPLES + 8;                      // inlier indexes of the best model, ascending
    unsigned *s_new = (unsigned *)(s_idx + HG_MAX_SAMPLES), *s_buf0 = s_new + RHO_WORDS, *s_buf1 = s_buf0 + RHO_WORDS;
    float *s_prod = (float *)(s_buf1 + RHO_WORDS);                           // 2 x RHO_ACC x (RHO_TILE2 + 1)
    __shared__ float s_H[9], s_Hn[9], s_acc[RHO_ACC];
    __shared__ int s_go, s_ninl_list, s_warp_cnt[RHO_WORDS], s_tot_inl, s_nstar, s_ns_which, s_ns_stop, s_cert, s_unc, s_curw, s_cnt;
    __shared__ double s_logAcc, s_logRej, s_logA, s_scan_s[RHO_NT / 32], s_scan_m[RHO_NT / 32], s_pre_s[RHO_NT / 32], s_pre_m[RHO_NT / 32];
    __shared__ unsigned s_ns_n[RHO_NT], s_ns_i[RHO_NT], s_ns_wn[RHO_NT / 32], s_ns_wi[RHO_NT / 32], s_ns_out[2];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int N = min(n_ptr[0], HG_MAX_SAMPLES);
    const int nwords = (N + 31) >> 5;
    if (N < 5) {   // fewer than 4 points: failure; exactly 4: cv::findHomography bypasses RHO -- neither occurs on the 2961-sample grid
        if (tid < 9) H_out[tid] = 0.0;
        if (tid == 0) for (int q = 1; q < 12; q++) info_out[q] = 0;
        return;
    }
    for (int i = tid; i < N; i += RHO_NT) { s_src[i] = g_src[i]; s_dst[i] = g_dst[i]; }
    for (int w = tid; w < RHO_WORDS; w += RHO_NT) { s_buf0[w] = 0u; s_buf1[w] = 0u; }
    {   // sacInitNonRand(beta = 0.35)
        const double bb = sqrt(0.35 * (1.0 - 0.35)) * 1.645;
        for (int n = tid; n <= N; n += RHO_NT) {
            unsigned v = 0;
            if (n >= 5 && n < N) { const double mu = n * 0.35, sigma = sqrt((double)n) * bb; v = (unsigned)ceil(4 + mu + sigma); }
            s_tbl[n] = (unsigned short)min(v, 65535u);
        }
    }
    __syncthreads();
    // ---- control state (thread 0)
    RhoPrng prng;
    unsigned it = 0, phNum = 4, phEndI = 1, phMax = N, phNumInl = 0, maxI = 2000, numInl_b = 0, n_models = 0;
    double phEndFpI = 0, epsilon = 0.1, delta = 0.01, A = 0, lamAcc = 0, lamRej = 0, logAcc = 0, logRej = 0, logA_d = 0;
    int dbg_replay = 0;
    long long clk_loop = 0, clk_nstar = 0, clk_lm = 0, clk_t0 = clock64(), clk_a = 0, clk_b = 0, clk_c = 0, clk_d = 0, clk_x = clock64();
    float Hb[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    unsigned *cur = s_buf0, *best = s_buf1;
    RhoPrng sp_prng = prng;
    unsigned sp_it = 0, sp_phNum = 4, sp_phEndI = 1;
    double sp_phEndFpI = 0;
    bool have_spec = false;
    int spec_go = 0;
    if (tid == 0) {
        prng.s0 = ~0ull; prng.s1 = 0ull;
        for (int i = 0; i < 20; i++) rho_random(prng);
        double numer = 1, denom = 1;
        for (unsigned i = 0; i < 4; i++) { numer *= 4 - i; denom *= N - i; }
        phEndFpI = 2000 * numer / denom;
        A = rho_design_sprt(delta, epsilon);
        lamRej = (1.0 - delta) / (1.0 - epsilon);
        lamAcc = delta / epsilon;
        logAcc = log(lamAcc); logRej = log(lamRej); logA_d = log(A);
    }
    const float distSq = 3.0f * 3.0f;
    if (tid == 0) { s_nstar = 0; s_curw = 0; }
    for (;;) {
        // ---- nStarOptimize of the model that has just become the best one (RHO_HEST_REFC::nStarOptimize), CTA-parallel.
        // The library walks test_n = N .. 21 with a running best ratio (inliers among the first test_n points) / test_n, strict
        // improvement only, stops at the first improving test_n whose inlier count is below the non-randomness table, or when the
        // prefix holds no inlier.  Here: prefix counts from the bit words, per-thread chunk maxima, an exclusive "first maximum"
        // scan over the threads, and the earliest stop position decides which prefix of improvements counts.
        __syncthreads();
        if (s_nstar) {
            const unsigned *bb = s_ns_which ? s_buf1 : s_buf0;
            for (int w = tid; w < RHO_WORDS; w += RHO_NT) s_warp_cnt[w] = w < nwords ? __popc(bb[w]) : 0;
            if (tid == 0) s_ns_stop = 0x7fffffff;
            __syncthreads();
            if (tid == 0) { int acc = 0; for (int w = 0; w < nwords; w++) { const int c = s_warp_cnt[w]; s_warp_cnt[w] = acc; acc += c; } }
            __syncthreads();
            auto I_of = [&](int n) -> unsigned { return n >= 32 * nwords ? (unsigned)(s_warp_cnt[nwords - 1] + __popc(bb[nwords - 1]))
                                                                        : (unsigned)(s_warp_cnt[n >> 5] + __popc(bb[n >> 5] & ((1u << (n & 31)) - 1u))); };
            // position j <-> test_n = N - j, j = 0 .. N - 21 ; chunk of this thread
            const int total = max(N - 20, 1), C = (total + RHO_NT - 1) / RHO_NT;   // (N <= 20: the walk is empty, position 0 = the initial state)
            const int j0 = min(tid * C, total), j1 = min(j0 + C, total);
            auto better = [](unsigned bi, unsigned bn, unsigned ai, unsigned an) { return bi * an > ai * bn; };   // b strictly better than a
            unsigned ln = 0, li = 0;     // chunk maximum (first occurrence); ln == 0: none
            for (int j = j0; j < j1; ++j) {
                const unsigned n = N - j, i = I_of((int)n);
                if (!ln || better(i, n, li, ln)) { ln = n; li = i; }
            }
            s_ns_n[tid] = ln; s_ns_i[tid] = li;
            __syncthreads();
            unsigned an = 0, ai = 0;     // first maximum over the earlier chunks of this warp
            for (int l = 0; l < lane; ++l) {
                const unsigned bn = s_ns_n[wid * 32 + l], bi = s_ns_i[wid * 32 + l];
                if (bn && (!an || better(bi, bn, ai, an))) { an = bn; ai = bi; }
            }
            if (lane == 31) {
                unsigned wn = an, wi = ai;
                if (ln && (!wn || better(li, ln, wi, wn))) { wn = ln; wi = li; }
                s_ns_wn[wid] = wn; s_ns_wi[wid] = wi;
            }
            __syncthreads();
            unsigned cn = 0, ci = 0;     // incoming running best of this thread's chunk
            for (int w = 0; w < wid; ++w) {
                const unsigned bn = s_ns_wn[w], bi = s_ns_wi[w];
                if (bn && (!cn || better(bi, bn, ci, cn))) { cn = bn; ci = bi; }
            }
            if (an && (!cn || better(ai, an, ci, cn))) { cn = an; ci = ai; }
            // walk the chunk with the true state: first stop position (zero prefix, or an improvement below the table)
            int stop = 0x7fffffff;
            {
                unsigned rn = cn, ri = ci;
                for (int j = j0; j < j1; ++j) {
                    const unsigned n = N - j, i = I_of((int)n);
                    if (i == 0) { stop = j; break; }
                    if (!rn) { rn = n; ri = i; continue; }            // j == 0: the initial state (N, numInl)
                    if (better(i, n, ri, rn)) {
                        if (i < s_tbl[n]) { stop = j; break; }
                        rn = n; ri = i;
                    }
                }
            }
            if (stop != 0x7fffffff) atomicMin(&s_ns_stop, stop);
            __syncthreads();
            const int E = min(s_ns_stop, total);     // positions j < E count; E >= 1 (position 0 is the initial state)
            if (j0 < E && E <= j1) {                 // exactly one thread: its chunk holds the last counted position
                unsigned rn = cn, ri = ci;
                for (int j = j0; j < E; ++j) {
                    const unsigned n = N - j, i = I_of((int)n);
                    if (!rn) { rn = n; ri = i; continue; }
                    if (better(i, n, ri, rn)) { rn = n; ri = i; }
                }
                s_ns_out[0] = rn; s_ns_out[1] = ri;
            }
            __syncthreads();
        }
        // ---- hypothesize (thread 0): PROSAC sample until a non-degenerate model or the end of the loop.  While warps 1 .. 15
        // evaluate model i, thread 0 already generates hypothesis i + 1 (2.5 k cycles of serial work: sampling, degeneracy
        // tests, the 8 x 9 elimination) on a COPY of the generator state.  Nothing the generator reads changes unless model i
        // becomes the best one (maxI, phMax); in that case the speculation is dropped and redone after the update.
        auto generate = [&](RhoPrng &g_prng, unsigned &g_it, unsigned &g_phNum, unsigned &g_phEndI, double &g_phEndFpI, float *Hdst) -> int {
            while (g_it < maxI || g_it < 100) {
                if (g_it >= g_phEndI && g_phNum < phMax) {
                    g_phNum++;
                    const double next = (g_phEndFpI * g_phNum) / (g_phNum - 4);
                    g_phEndI += (unsigned)ceil(next - g_phEndFpI);
                    g_phEndFpI = next;
                }
                unsigned smpl[4];
                if (g_it > g_phEndI) rho_rnd_smpl(g_prng, 4, smpl, g_phNum);
                else { rho_rnd_smpl(g_prng, 3, smpl, g_phNum - 1); smpl[3] = g_phNum - 1; }
                float2 k[8];
                for (int q = 0; q < 4; q++) { k[q] = s_src[smpl[q]]; k[4 + q] = s_dst[smpl[q]]; }
                if (rho_sample_degenerate(k)) { ++g_it; continue; }
                float Hc[9];
                rho_h_func(k, Hc);
                const float f = Hc[0] + Hc[1] + Hc[2] + Hc[3] + Hc[4] + Hc[5] + Hc[6] + Hc[7];
                if (f != f) { ++g_it; continue; }
                for (int q = 0; q < 9; q++) Hdst[q] = Hc[q];
                return 1;
            }
            return 0;
        };
        if (tid == 0) {
            if (s_nstar) {
                s_nstar = 0;
                const unsigned best_n = s_ns_out[0], bestNumInl = s_ns_out[1];
                if (bestNumInl * phMax > phNumInl * best_n) {
                    phMax = best_n;
                    phNumInl = bestNumInl;
                    maxI = rho_iter_bound(0.995, (double)phNumInl / phMax, maxI);
                }
            }
            int go;
            if (have_spec) { for (int q = 0; q < 9; q++) s_H[q] = s_Hn[q]; go = spec_go; }     // accepted speculation of the last round
            else go = generate(prng, it, phNum, phEndI, phEndFpI, s_H);
            s_go = go;
            s_tot_inl = 0; s_cert = 0x7fffffff; s_unc = 0x7fffffff; s_cnt = 0;
            s_logAcc = logAcc; s_logRej = logRej; s_logA = logA_d;
            { const long long t_ = clock64(); clk_a += t_ - clk_x; clk_x = t_; }
        }
        __syncthreads();
        if (!s_go) break;
        if (wid == 0) {
            if (tid == 0) {      // hypothesis i + 1, speculatively (the iteration counter advances by one at the end of this round)
                sp_prng = prng; sp_it = it + 1; sp_phNum = phNum; sp_phEndI = phEndI; sp_phEndFpI = phEndFpI;
                spec_go = generate(sp_prng, sp_it, sp_phNum, sp_phEndI, sp_phEndFpI, s_Hn);
            }
        } else {
        // ---- warps 1 .. 15 (RHO_NE threads, barrier 1): all reprojection tests of the model (evaluateModelSPRT's per-point arithmetic)
        const int et = tid - 32, ew = wid - 1;
        {
            const float h0 = s_H[0], h1 = s_H[1], h2 = s_H[2], h3 = s_H[3], h4 = s_H[4], h5 = s_H[5], h6 = s_H[6], h7 = s_H[7];
            for (int base = 0; base < nwords * 32; base += RHO_NE) {
                const int i = base + et;
                bool inl = false;
                if (i < N) {
                    const float x = s_src[i].x, y = s_src[i].y, X = s_dst[i].x, Y = s_dst[i].y;
                    float rx = h0 * x + h1 * y + h2;
                    float ry = h3 * x + h4 * y + h5;
                    const float rz = h6 * x + h7 * y + 1.0f;
                    rx /= rz; ry /= rz;
                    rx -= X; ry -= Y;
                    rx *= rx; ry *= ry;
                    inl = (rx + ry) <= distSq;
                }
                const unsigned m = __ballot_sync(0xffffffffu, inl);
                if (lane == 0 && (i >> 5) < RHO_WORDS) { s_new[i >> 5] = m; if (m) atomicAdd(&s_tot_inl, __popc(m)); }
            }
        }
        rho_ebar();
        // ---- SPRT: the first point at which lambda = prod(accept / reject factors) exceeds A (evaluateModelSPRT stops there).
        // The library multiplies sequentially in FP64; a dependent FP64 chain over up to 3000 points costs ~0.1 ms per model
        // on this part, so the decision is taken in the log domain, in parallel: S_i = FP64 prefix sum of the log factors
        // (a block scan), compared with log A.  |S_i - log(lambda_i)| < 1e-11, so outside a band of 1e-9 around log A the
        // comparison "lambda_i <= A" is decided exactly as the library decides it; the sequential product can also underflow
        // (to a denormal, or to zero where it stays): C_i = S_i - min(0, min_j<=i S_j + 700) bounds it from above after such a
        // dip.  Only if a point inside the band (or a recovery from a dip) comes before the first certain exit -- never
        // observed -- thread 0 falls back to the library's sequential product.
        {
            const int CH = (N + RHO_NE - 1) / RHO_NE;                    // consecutive points per thread
            const int i0 = min(et * CH, N), i1 = min(i0 + CH, N);
            double loc = 0.0, locmin = 0.0;                              // chunk sum, minimum prefix inside the chunk (0 = empty prefix)
            for (int i = i0; i < i1; ++i) {
                loc += ((s_new[i >> 5] >> (i & 31)) & 1u) ? s_logAcc : s_logRej;
                locmin = fmin(locmin, loc);
            }
            // inclusive scan of (sum, min prefix) over the threads: (s1, m1) . (s2, m2) = (s1 + s2, min(m1, s1 + m2))
            double ss = loc, mm = locmin;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const double ps = __shfl_up_sync(0xffffffffu, ss, off), pm = __shfl_up_sync(0xffffffffu, mm, off);
                if (lane >= off) { mm = fmin(pm, ps + mm); ss = ps + ss; }
            }
            if (lane == 31) { s_scan_s[ew] = ss; s_scan_m[ew] = mm; }
            rho_ebar();
            if (ew == 0) {     // exclusive scan of the RHO_NE / 32 warp totals with the same operator (one warp, four shuffle steps)
                double ws = lane < RHO_NE / 32 ? s_scan_s[lane] : 0.0, wm = lane < RHO_NE / 32 ? s_scan_m[lane] : 0.0;
#pragma unroll
                for (int off = 1; off < 16; off <<= 1) {
                    const double ps = __shfl_up_sync(0xffffffffu, ws, off), pm = __shfl_up_sync(0xffffffffu, wm, off);
                    if (lane >= off) { wm = fmin(pm, ps + wm); ws = ps + ws; }
                }
                double xs = __shfl_up_sync(0xffffffffu, ws, 1), xm = __shfl_up_sync(0xffffffffu, wm, 1);
                if (lane == 0) { xs = 0.0; xm = 0.0; }
                if (lane < RHO_NE / 32) { s_pre_s[lane] = xs; s_pre_m[lane] = xm; }
            }
            rho_ebar();
            const double bs = s_pre_s[ew], bm = s_pre_m[ew];             // everything before this warp
            // exclusive prefix of this thread = (before the warp) . (inclusive of the previous lane)
            double es = __shfl_up_sync(0xffffffffu, ss, 1), em = __shfl_up_sync(0xffffffffu, mm, 1);
            if (lane == 0) { es = 0.0; em = 0.0; }
            double S = bs + es, M = fmin(bm, bs + em);
            const double lA = s_logA;
            int cert = 0x7fffffff, unc = 0x7fffffff;
            if (!(lA == CUDART_INF))             // A = inf (epsilon = 1): "lambda <= A" can never fail
            for (int i = i0; i < i1; ++i) {
                S += ((s_new[i >> 5] >> (i & 31)) & 1u) ? s_logAcc : s_logRej;
                M = fmin(M, S);
                if (M >= -700.0) {
                    if (S > lA + 1e-9) { if (cert == 0x7fffffff) cert = i; }
                    else if (!(S < lA - 1e-9)) { if (unc == 0x7fffffff) unc = i; }       // inside the band (or NaN)
                } else if (!(S - (M + 700.0) < lA - 1e-9)) { if (unc == 0x7fffffff) unc = i; }
            }
            if (cert != 0x7fffffff) atomicMin(&s_cert, cert);
            if (unc != 0x7fffffff) atomicMin(&s_unc, unc);
        }
        rho_ebar();
        // ---- the evaluation overwrote the inlier flags of the points it tested only: merge the tested prefix into the current
        // buffer and count its inliers, one word per thread (unless the sequential fallback has to find the exit first)
        if (!(s_unc < s_cert)) {
            const int Nt = s_cert != 0x7fffffff ? s_cert + 1 : N;
            const int full = Nt >> 5, rem = Nt & 31;
            unsigned *cw = s_curw ? s_buf1 : s_buf0;
            int c = 0;
            if (et < full) { const unsigned b = s_new[et]; cw[et] = b; c = __popc(b); }
            else if (et == full && rem) { const unsigned msk = (1u << rem) - 1u, b = s_new[full] & msk; cw[full] = (cw[full] & ~msk) | b; c = __popc(b); }
            if (et < ((nwords + 31) & ~31)) {       // whole warps
                c = __reduce_add_sync(0xffffffffu, c);
                if (lane == 0 && c) atomicAdd(&s_cnt, c);
            }
        }
        }
        __syncthreads();
        const bool fallback = s_unc < s_cert;
        if (tid == 0) {
            { const long long t_ = clock64(); clk_b += t_ - clk_x; clk_x = t_; }
            ++n_models;
            int Ntested = N;
            bool good = true;
            unsigned numInl_c = 0;
            if (fallback) {
                // fallback: the library's sequential FP64 product
                ++dbg_replay;
                double lam = 1.0;
                for (int i = 0; i < N; ++i) {
                    lam *= ((s_new[i >> 5] >> (i & 31)) & 1u) ? lamAcc : lamRej;
                    if (!(lam <= A)) { good = false; Ntested = i + 1; break; }
                }
                const int full = Ntested >> 5, rem = Ntested & 31;
                for (int q = 0; q < full; q++) { const unsigned b = s_new[q]; cur[q] = b; numInl_c += __popc(b); }
                if (rem) {
                    const unsigned msk = (1u << rem) - 1u, b = s_new[full] & msk;
                    cur[full] = (cur[full] & ~msk) | b;
                    numInl_c += __popc(b);
                }
            } else {
                if (s_cert != 0x7fffffff) { good = false; Ntested = s_cert + 1; }
                numInl_c = (unsigned)s_cnt;
            }
            { const long long t_ = clock64(); clk_c += t_ - clk_x; clk_x = t_; }
            // updateSPRT
            if (good) {
                if (numInl_c > numInl_b) {
                    epsilon = (double)numInl_c / N;
                    A = rho_design_sprt(delta, epsilon);
                    lamRej = (1.0 - delta) / (1.0 - epsilon);
                    lamAcc = delta / epsilon;
                    logAcc = log(lamAcc); logRej = log(lamRej); logA_d = log(A);
                }
            } else {
                const double newDelta = (double)numInl_c / Ntested;
                if (newDelta > 0) {
                    const double relChange = fabs(delta - newDelta) / delta;
                    if (relChange > 0.1) {
                        delta = newDelta;
                        A = rho_design_sprt(delta, epsilon);
                        lamRej = (1.0 - delta) / (1.0 - epsilon);
                        lamAcc = delta / epsilon;
                        logAcc = log(lamAcc); logRej = log(lamRej); logA_d = log(A);
                    }
                }
            }
            if (numInl_c > numInl_b) {      // saveBestModel, updateBounds, nStarOptimize
                for (int q = 0; q < 9; q++) Hb[q] = s_H[q];
                unsigned *t = cur; cur = best; best = t;
                s_curw = cur == s_buf1 ? 1 : 0;
                numInl_b = numInl_c;
                maxI = rho_iter_bound(0.995, (double)numInl_b / N, maxI);
                s_nstar = 1; s_ns_which = best == s_buf1 ? 1 : 0;      // nStarOptimize runs CTA-wide at the top of the next iteration
            }
            ++it;
            // the speculated hypothesis i + 1 stands unless this model changed what the generator reads (maxI here, phMax after nStarOptimize)
            have_spec = !s_nstar;
            if (have_spec) { prng = sp_prng; it = sp_it; phNum = sp_phNum; phEndI = sp_phEndI; phEndFpI = sp_phEndFpI; }
            { const long long t_ = clock64(); clk_d += t_ - clk_x; clk_x = t_; }
        }
    }
    // ---- final refinement (RHO_HEST_REFC::refine) over the inliers of the best model, canRefine: more than 4 inliers
    __shared__ int s_nb, s_which, s_lm;
    if (tid == 0) { clk_loop = clock64() - clk_t0; clk_t0 = clock64(); }
    if (tid == 0) { s_nb = (int)numInl_b; s_which = best == s_buf1 ? 1 : 0; for (int q = 0; q < 9; q++) s_H[q] = Hb[q]; }
    __syncthreads();
    const unsigned *bestb = s_which ? s_buf1 : s_buf0;
    const int nb = s_nb;
    if (mask_out)
        for (int i = tid; i < N; i += RHO_NT) mask_out[i] = nb >= 4 ? (unsigned char)((bestb[i >> 5] >> (i & 31)) & 1u) : 0;
    int lm_iters = 0;
    if (nb > 4) {
        // ascending list of inlier indexes
        for (int w = tid; w < RHO_WORDS; w += RHO_NT) s_warp_cnt[w] = w < nwords ? __popc(bestb[w]) : 0;
        __syncthreads();
        if (tid == 0) { int acc = 0; for (int w = 0; w < nwords; w++) { const int c = s_warp_cnt[w]; s_warp_cnt[w] = acc; acc += c; } s_ninl_list = acc; }
        __syncthreads();
        for (int w = tid; w < nwords; w += RHO_NT) {
            unsigned b = bestb[w];
            int o = s_warp_cnt[w];
            while (b) { const int q = __ffs(b) - 1; b &= b - 1; s_idx[o++] = (unsigned short)(32 * w + q); }
        }
        __syncthreads();
        const int ni = s_ninl_list;
        // sacCalcJacobianErrors at the homography in s_H: per-point products staged per tile, accumulators summed in inlier order
        // The sequential sums (one dependent FADD per inlier and accumulator, ~6 cycles each) are the critical path of the
        // refinement; the per-point products hide behind them: while the RHO_ACC accumulator threads (warps 0 and 1, alone on
        // the SM's schedulers 0 and 1) sum tile t, the producer warps (those of schedulers 2 and 3: warps 2, 3, 6, 7, ...) fill
        // tile t + 1 of the double buffer.
        const int wq = tid >> 5;
        const int pt = (wq & 2) ? (((wq >> 2) * 2 + (wq & 1)) * 32 + (tid & 31)) : -1;     // producer index 0 .. RHO_TILE2 - 1
        auto produce = [&](int t0, int tn, float *buf) {
            if (pt < 0 || pt >= tn) return;
            const int i = s_idx[t0 + pt];
            const float x = s_src[i].x, y = s_src[i].y, X = s_dst[i].x, Y = s_dst[i].y;
            const float Wd = s_H[6] * x + s_H[7] * y + 1.0f;
            const float iW = fabsf(Wd) > 1.1920929e-07f ? 1.0f / Wd : 0;
            const float rX = (s_H[0] * x + s_H[1] * y + s_H[2]) * iW;
            const float rY = (s_H[3] * x + s_H[4] * y + s_H[5]) * iW;
            const float eX = rX - X, eY = rY - Y;
            const float e = eX * eX + eY * eY;
            const float dxh11 = x * iW, dxh12 = y * iW, dxh13 = iW, dxh31 = -rX * x * iW, dxh32 = -rX * y * iW;
            const float dyh21 = x * iW, dyh22 = y * iW, dyh23 = iW, dyh31 = -rY * x * iW, dyh32 = -rY * y * iW;
            float *col = buf + pt;
#define RHO_P(a, v) col[(a) * (RHO_TILE2 + 1)] = (v)
            RHO_P(0, dxh11 * dxh11);
            RHO_P(1, dxh11 * dxh12); RHO_P(2, dxh12 * dxh12);
            RHO_P(3, dxh11 * dxh13); RHO_P(4, dxh12 * dxh13); RHO_P(5, dxh13 * dxh13);
            RHO_P(6, dyh21 * dyh21);
            RHO_P(7, dyh21 * dyh22); RHO_P(8, dyh22 * dyh22);
            RHO_P(9, dyh21 * dyh23); RHO_P(10, dyh22 * dyh23); RHO_P(11, dyh23 * dyh23);
            RHO_P(12, dxh11 * dxh31); RHO_P(13, dxh12 * dxh31); RHO_P(14, dxh13 * dxh31);
            RHO_P(15, dyh21 * dyh31); RHO_P(16, dyh22 * dyh31); RHO_P(17, dyh23 * dyh31);
            RHO_P(18, dxh31 * dxh31 + dyh31 * dyh31);
            RHO_P(19, dxh11 * dxh32); RHO_P(20, dxh12 * dxh32); RHO_P(21, dxh13 * dxh32);
            RHO_P(22, dyh21 * dyh32); RHO_P(23, dyh22 * dyh32); RHO_P(24, dyh23 * dyh32);
            RHO_P(25, dxh31 * dxh32 + dyh31 * dyh32);
            RHO_P(26, dxh32 * dxh32 + dyh32 * dyh32);
            RHO_P(27, eX * dxh11); RHO_P(28, eX * dxh12); RHO_P(29, eX * dxh13);
            RHO_P(30, eY * dyh21); RHO_P(31, eY * dyh22); RHO_P(32, eY * dyh23);
            RHO_P(33, eX * dxh31 + eY * dyh31);
            RHO_P(34, eX * dxh32 + eY * dyh32);
            RHO_P(35, e);
#undef RHO_P
        };
        auto eval = [&]() {
            float acc = 0.0f;
            __syncthreads();                                   // s_H of this evaluation is visible; the buffers are free
            produce(0, min(RHO_TILE2, ni), s_prod);
            __syncthreads();
            int par = 0;
            for (int t0 = 0; t0 < ni; t0 += RHO_TILE2, par ^= 1) {
                const int tn = min(RHO_TILE2, ni - t0);
                if (t0 + RHO_TILE2 < ni) produce(t0 + RHO_TILE2, min(RHO_TILE2, ni - t0 - RHO_TILE2), s_prod + (par ^ 1) * RHO_ACC * (RHO_TILE2 + 1));
                if (tid < RHO_ACC) {
                    const float *row = s_prod + par * RHO_ACC * (RHO_TILE2 + 1) + tid * (RHO_TILE2 + 1);
                    int q = 0;
                    for (; q + 8 <= tn; q += 8) {
                        const float v0 = row[q], v1 = row[q + 1], v2 = row[q + 2], v3 = row[q + 3], v4 = row[q + 4], v5 = row[q + 5], v6 = row[q + 6], v7 = row[q + 7];
                        acc += v0; acc += v1; acc += v2; acc += v3; acc += v4; acc += v5; acc += v6; acc += v7;
                    }
                    for (; q < tn; q++) acc += row[q];
                }
                __syncthreads();
            }
            if (tid < RHO_ACC) s_acc[tid] = acc;
            __syncthreads();
        };
        eval();
        // Levenberg-Marquardt loop: thread 0 holds the state; the candidate is evaluated in full (J^T J, J^T e, S) at once,
        // which is what the library computes in two calls when the step is accepted
        float S = 0, L = 100.0f, JtJ[8][8], Jte[8], Hcur[8];
        if (tid == 0) {
            for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) JtJ[r][c] = 0.0f;
            for (int a = 0; a < 27; a++) JtJ[c_rho_acc_r[a]][c_rho_acc_c[a]] = s_acc[a];
            for (int q = 0; q < 8; q++) { Jte[q] = s_acc[27 + q]; Hcur[q] = s_H[q]; }
            S = s_acc[35];
        }
        float dH[8];
        for (int i = 0; i < 100; i++) {
            if (tid == 0) {
                float T[8][8];
                while (!rho_chol8(JtJ, L, T)) L *= 2.0f;
                rho_tr_inv8(T);
                rho_tri_solve8(T, Jte, dH);
                for (int q = 0; q < 8; q++) s_H[q] = Hcur[q] - dH[q];
            }
            eval();
            int stop = 0;
            if (tid == 0) {
                ++lm_iters;
                const float newS = s_acc[35];
                const float dS = S - newS;
                float dL = 0;
                for (int q = 0; q < 8; q++) dL += dH[q] * dH[q];
                dL *= L;
                for (int q = 0; q < 8; q++) dL += dH[q] * Jte[q];
                dL *= 0.5f;
                const float gain = fabsf(dL) < 1.1920929e-07f ? dS : dS / dL;
                if (gain < 0.25f) {
                    L *= 8;
                    if (L > 1000.0f / 1.1920929e-07f) stop = 1;
                } else if (gain > 0.75f) {
                    L *= 0.5f;
                }
                if (!stop && gain > 0) {
                    S = newS;
                    for (int q = 0; q < 8; q++) Hcur[q] = s_H[q];
                    for (int a = 0; a < 27; a++) JtJ[c_rho_acc_r[a]][c_rho_acc_c[a]] = s_acc[a];
                    for (int q = 0; q < 8; q++) Jte[q] = s_acc[27 + q];
                }
                s_lm = stop;
            }
            __syncthreads();
            if (s_lm) break;
        }
        if (tid == 0) for (int q = 0; q < 8; q++) Hb[q] = Hcur[q];
    }
    if (tid == 0) {
        const bool ok = numInl_b >= 4;
        for (int q = 0; q < 9; q++) H_out[q] = ok ? (double)Hb[q] : 0.0;
        clk_lm = clock64() - clk_t0;
        info_out[1] = ok ? (int)numInl_b : 0;
        info_out[2] = (int)n_models;
        info_out[3] = lm_iters;
        info_out[4] = (int)(clk_loop >> 10); info_out[5] = (int)(clk_nstar >> 10); info_out[6] = (int)(clk_lm >> 10); info_out[7] = (int)it;
        info_out[8] = (int)(clk_a >> 10); info_out[9] = (int)(clk_b >> 10); info_out[10] = (int)(clk_d >> 10); info_out[11] = dbg_replay;
        (void)clk_c;
    }
}

int homography_init(sindyn_base *ctx, HomographyStage *g, int W, int H)
{
    g->W = W; g->H = H;
    int ns = ((W - 1) / 10) * ((H - 1) / 10);
    if (ns > HG_MAX_SAMPLES || ns > HG_GAUSS_N) { ctx->err = "homography: sample grid exceeds HG_MAX_SAMPLES"; return SINDYN_ERR_CAPACITY; }
    SD_CHECK(ctx->dalloc(&g->counts, 32));
    SD_CHECK(ctx->dalloc(&g->pts, 2 * HG_MAX_SAMPLES));
    SD_CHECK(ctx->dalloc(&g->pts_last, 2 * HG_MAX_SAMPLES));
    SD_CHECK(ctx->dalloc(&g->n_pairs, 12));
    SD_CHECK(ctx->dalloc(&g->H_dev, 9));
    SD_CHECK(ctx->dalloc(&g->inl_mask, HG_MAX_SAMPLES));
    CU_CHECK(ctx, cudaFuncSetAttribute(k_rho, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RHO_SMEM));
    return SINDYN_OK;
}

int homography_sample(sindyn_base *ctx, HomographyStage *g, const float *flow, const uint8_t *label_last, const uint8_t *dyna_last)
{
    CU_CHECK(ctx, cudaMemsetAsync(g->counts, 0, sizeof(int) * 32, ctx->stream));
    LAUNCH(ctx, k_cluster_weight_counts, SINDYN_NUM_SMS_B200, 256, 0, label_last, dyna_last, g->W * g->H, g->counts);
    LAUNCH(ctx, k_sample_pairs, 1, 1024, 0, (const float2 *)flow, label_last, dyna_last, g->counts, g->W, g->H, (float2 *)g->pts,
           (float2 *)g->pts_last, g->n_pairs);
    LAUNCH_CHECK(ctx);
    return SINDYN_OK;
}

int homography_estimate(sindyn_base *ctx, HomographyStage *g)
{
    LAUNCH(ctx, k_rho, 1, RHO_NT, RHO_SMEM, (const float2 *)g->pts, (const float2 *)g->pts_last, g->n_pairs, g->H_dev, g->n_pairs, g->inl_mask);
    LAUNCH_CHECK(ctx);
    return SINDYN_OK;
}
