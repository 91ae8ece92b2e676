/*
 * oracle/rho_cpu.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or call this
 * file.  The product path is libsindyn_cuda.so and must never link it.
 *
 * What it restates: cv::findHomography(inputPoints, inputPointsLast, cv::noArray(), cv::RHO) as the reference calls it
 * (ORB_SLAM2/src/DynaDetect.cc:1235) -- i.e. OpenCV's calib3d "RHO" estimator (modules/calib3d/src/rho.cpp, class
 * RHO_HEST_REFC, driven by createAndRunRHORegistrator in fundam.cpp with ransacReprojThreshold 3, maxIters 2000,
 * confidence 0.995, beta 0.35, flags NR | FINAL_REFINEMENT): PROSAC sampling with the estimator's own xorshift128+
 * generator (seeded with ~0 at construction, 20 warm-up draws), the degeneracy tests of the 4-point sample (coincident
 * coordinates, orientation of the four point triples), a hand-unrolled 4-point homography solve in float, SPRT
 * evaluation of every model (Matas & Chum, ICCV 2005), the non-randomness criterion (PROSAC, Chum & Matas, CVPR 2005) and
 * the final Levenberg-Marquardt refinement over the inliers of the best model (float, Cholesky with multiplicative
 * damping).
 *
 * OpenCV is an un-vendored dependency of the reference (OpenCV 4.2.0 EXACT, ORB_SLAM2/CMakeLists.txt:44); its source is
 * not in this container.  This restatement is PINNED against the real library: tests/test_rho_cpu.py runs it next to
 * cv2.findHomography(..., cv2.RHO) (cv2 4.13) on the sample lists of synthetic frames and on randomised correspondences
 * and requires the same inlier mask and a BIT-IDENTICAL H (N >= 5; for N == 4 cv::findHomography bypasses the estimator).  The CUDA kernel (sindslam_b200/csrc/homography.cu) follows this
 * file step by step and is compared with cv2 itself in tests/test_homography_gpu.py.
 *
 * Build: make -C oracle   (-> oracle/_build/librho_cpu.so)
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SMPL_SIZE 4
#define SPRT_T_M 25
#define SPRT_M_S 1
#define SPRT_EPSILON 0.1
#define SPRT_DELTA 0.01
#define MIN_DELTA_CHNG 0.1
#define CHI_SQ 1.645
#define MAXLEVMARQITERS 100
#define LM_GAIN_LO 0.25
#define LM_GAIN_HI 0.75

typedef struct {
    /* arguments */
    const float *src, *dst;
    unsigned N;
    float maxD;
    unsigned maxI, rConvg;
    double cfd;
    unsigned minInl;
    double beta;
    /* PROSAC control */
    unsigned i, phNum, phEndI;
    double phEndFpI;
    unsigned phMax, phNumInl, numModels;
    unsigned smpl[4];
    /* models */
    float pkd[16];
    float Hc[9], Hb[9];
    char *inl_c, *inl_b;
    unsigned numInl_c, numInl_b;
    /* NR */
    unsigned *tbl;
    /* SPRT */
    double t_M, m_S, epsilon, delta, A, lambdaAccept, lambdaReject;
    unsigned Ntested, Ntestedtotal;
    int good;
    /* PRNG */
    uint64_t s[2];
    /* diagnostics */
    unsigned n_degenerate_sample, n_degenerate_model, lm_iters;
} Rho;

static double fast_random(Rho *p)
{
    uint64_t x = p->s[0];
    const uint64_t y = p->s[1];
    x ^= x << 23;
    x ^= x >> 17;
    x ^= y ^ (y >> 26);
    p->s[0] = y;
    p->s[1] = x;
    const uint64_t s = x + y;
    return s * 5.421010862427522e-20; /* 2^-64 */
}
static void fast_seed(Rho *p, uint64_t seed)
{
    p->s[0] = seed;
    p->s[1] = ~seed;
    for (int i = 0; i < 20; i++) fast_random(p);
}

static double sac_init_pend_fpi(unsigned ransacConvg, unsigned n, unsigned s)
{
    double numer = 1, denom = 1;
    for (unsigned i = 0; i < s; i++) { numer *= s - i; denom *= n - i; }
    return ransacConvg * numer / denom;
}

static unsigned sac_calc_iter_bound(double confidence, double inlierRate, unsigned sampleSize, unsigned maxIterBound)
{
    unsigned retVal;
    const double atLeastOneOutlierProbability = 1. - pow(inlierRate, (double)sampleSize);
    if (atLeastOneOutlierProbability >= 1.) retVal = maxIterBound;
    else if (atLeastOneOutlierProbability <= 0.) retVal = 1;
    else retVal = (unsigned)ceil(log(1. - confidence) / log(atLeastOneOutlierProbability));
    return retVal <= maxIterBound ? retVal : maxIterBound;
}

static double sac_design_sprt(double delta, double epsilon, double t_M, double m_S)
{
    const double C = (1 - delta) * log((1 - delta) / (1 - epsilon)) + delta * log(delta / epsilon);
    const double K = t_M * C / m_S + 1;
    double An = K, prevAn;
    unsigned i = 0;
    do {
        prevAn = An;
        An = K + log(An);
    } while ((An - prevAn > 1.5e-8) && (++i < 10));
    return An;
}
static void design_sprt(Rho *p)
{
    p->A = sac_design_sprt(p->delta, p->epsilon, p->t_M, p->m_S);
    p->lambdaReject = ((1.0 - p->delta) / (1.0 - p->epsilon));
    p->lambdaAccept = ((p->delta) / (p->epsilon));
}

static void sac_init_nonrand(double beta, unsigned start, unsigned N, unsigned *tbl)
{
    unsigned n = SMPL_SIZE + 1 > start ? SMPL_SIZE + 1 : start;
    const double beta_beta1_sq_chi = sqrt(beta * (1.0 - beta)) * CHI_SQ;
    for (; n < N; n++) {
        const double mu = n * beta;
        const double sigma = sqrt((double)n) * beta_beta1_sq_chi;
        tbl[n] = (unsigned)ceil(SMPL_SIZE + mu + sigma);
    }
}

static void rnd_smpl(Rho *p, unsigned sampleSize, unsigned *currentSample, unsigned dataSetSize)
{
    unsigned i, j;
    if (sampleSize * 2 > dataSetSize) {
        /* selection sampling (Knuth, Algorithm S) */
        for (i = 0, j = 0; i < dataSetSize && j < sampleSize; i++) {
            const double U = fast_random(p);
            if ((dataSetSize - i) * U < (sampleSize - j)) currentSample[j++] = i;
        }
    } else {
        /* draw until sampleSize indexes are distinct */
        for (i = 0; i < sampleSize; i++) {
            int inList;
            do {
                currentSample[i] = (unsigned)(dataSetSize * fast_random(p));
                inList = 0;
                for (j = 0; j < i; j++)
                    if (currentSample[i] == currentSample[j]) { inList = 1; break; }
            } while (inList);
        }
    }
}

static int is_sample_degenerate(Rho *p)
{
    const unsigned i0 = p->smpl[0], i1 = p->smpl[1], i2 = p->smpl[2], i3 = p->smpl[3];
    typedef struct { float x, y; } Pt;
    Pt *k = (Pt *)p->pkd;
    const Pt *src = (const Pt *)p->src, *dst = (const Pt *)p->dst;
    k[0] = src[i0]; k[1] = src[i1]; k[2] = src[i2]; k[3] = src[i3];
    k[4] = dst[i0]; k[5] = dst[i1]; k[6] = dst[i2]; k[7] = dst[i3];
    /* coincident coordinates in either image */
    if (k[0].x == k[1].x || k[1].x == k[2].x || k[2].x == k[3].x || k[0].x == k[2].x || k[0].x == k[3].x || k[1].x == k[3].x ||
        k[0].y == k[1].y || k[1].y == k[2].y || k[2].y == k[3].y || k[0].y == k[2].y || k[0].y == k[3].y || k[1].y == k[3].y ||
        k[4].x == k[5].x || k[5].x == k[6].x || k[6].x == k[7].x || k[4].x == k[6].x || k[4].x == k[7].x || k[5].x == k[7].x ||
        k[4].y == k[5].y || k[5].y == k[6].y || k[6].y == k[7].y || k[4].y == k[6].y || k[4].y == k[7].y || k[5].y == k[7].y)
        return 1;
    /* strong geometric constraint: the orientation of the point triples must be preserved */
    /* (0 x 1) * 2 */
    const float cross0s0 = k[0].y - k[1].y, cross0s1 = k[1].x - k[0].x, cross0s2 = k[0].x * k[1].y - k[0].y * k[1].x;
    const float dots0 = cross0s0 * k[2].x + cross0s1 * k[2].y + cross0s2;
    const float cross0d0 = k[4].y - k[5].y, cross0d1 = k[5].x - k[4].x, cross0d2 = k[4].x * k[5].y - k[4].y * k[5].x;
    const float dotd0 = cross0d0 * k[6].x + cross0d1 * k[6].y + cross0d2;
    if (((int)dots0 ^ (int)dotd0) < 0) return 1;
    /* (0 x 1) * 3 */
    const float dots1 = cross0s0 * k[3].x + cross0s1 * k[3].y + cross0s2;
    const float dotd1 = cross0d0 * k[7].x + cross0d1 * k[7].y + cross0d2;
    if (((int)dots1 ^ (int)dotd1) < 0) return 1;
    /* (2 x 3) * 0 */
    const float cross2s0 = k[2].y - k[3].y, cross2s1 = k[3].x - k[2].x, cross2s2 = k[2].x * k[3].y - k[2].y * k[3].x;
    const float dots2 = cross2s0 * k[0].x + cross2s1 * k[0].y + cross2s2;
    const float cross2d0 = k[6].y - k[7].y, cross2d1 = k[7].x - k[6].x, cross2d2 = k[6].x * k[7].y - k[6].y * k[7].x;
    const float dotd2 = cross2d0 * k[4].x + cross2d1 * k[4].y + cross2d2;
    if (((int)dots2 ^ (int)dotd2) < 0) return 1;
    /* (2 x 3) * 1 */
    const float dots3 = cross2s0 * k[1].x + cross2s1 * k[1].y + cross2s2;
    const float dotd3 = cross2d0 * k[5].x + cross2d1 * k[5].y + cross2d2;
    if (((int)dots3 ^ (int)dotd3) < 0) return 1;
    return 0;
}

/* hand-unrolled Gaussian elimination of the 8x9 system of a 4-point homography, all in float */
static void h_func(const float *pp, float *H)
{
    const float x0 = pp[0], y0 = pp[1], x1 = pp[2], y1 = pp[3], x2 = pp[4], y2 = pp[5], x3 = pp[6], y3 = pp[7];
    const float X0 = pp[8], Y0 = pp[9], X1 = pp[10], Y1 = pp[11], X2 = pp[12], Y2 = pp[13], X3 = pp[14], Y3 = pp[15];
    const float x0X0 = x0 * X0, x1X1 = x1 * X1, x2X2 = x2 * X2, x3X3 = x3 * X3;
    const float x0Y0 = x0 * Y0, x1Y1 = x1 * Y1, x2Y2 = x2 * Y2, x3Y3 = x3 * Y3;
    const float y0X0 = y0 * X0, y1X1 = y1 * X1, y2X2 = y2 * X2, y3X3 = y3 * X3;
    const float y0Y0 = y0 * Y0, y1Y1 = y1 * Y1, y2Y2 = y2 * Y2, y3Y3 = y3 * Y3;
    float minor[2][4] = {{x0 - x2, x1 - x2, x2, x3 - x2}, {y0 - y2, y1 - y2, y2, y3 - y2}};
    float major[3][8] = {{x2X2 - x0X0, x2X2 - x1X1, -x2X2, x2X2 - x3X3, x2Y2 - x0Y0, x2Y2 - x1Y1, -x2Y2, x2Y2 - x3Y3},
                         {y2X2 - y0X0, y2X2 - y1X1, -y2X2, y2X2 - y3X3, y2Y2 - y0Y0, y2Y2 - y1Y1, -y2Y2, y2Y2 - y3Y3},
                         {(X0 - X2), (X1 - X2), (X2), (X3 - X2), (Y0 - Y2), (Y1 - Y2), (Y2), (Y3 - Y2)}};
    /* eliminate column 0 of rows 1 and 3 */
    float scalar1 = minor[0][0], scalar2 = minor[0][1];
    minor[1][1] = minor[1][1] * scalar1 - minor[1][0] * scalar2;
    major[0][1] = major[0][1] * scalar1 - major[0][0] * scalar2;
    major[1][1] = major[1][1] * scalar1 - major[1][0] * scalar2;
    major[2][1] = major[2][1] * scalar1 - major[2][0] * scalar2;
    major[0][5] = major[0][5] * scalar1 - major[0][4] * scalar2;
    major[1][5] = major[1][5] * scalar1 - major[1][4] * scalar2;
    major[2][5] = major[2][5] * scalar1 - major[2][4] * scalar2;
    scalar2 = minor[0][3];
    minor[1][3] = minor[1][3] * scalar1 - minor[1][0] * scalar2;
    major[0][3] = major[0][3] * scalar1 - major[0][0] * scalar2;
    major[1][3] = major[1][3] * scalar1 - major[1][0] * scalar2;
    major[2][3] = major[2][3] * scalar1 - major[2][0] * scalar2;
    major[0][7] = major[0][7] * scalar1 - major[0][4] * scalar2;
    major[1][7] = major[1][7] * scalar1 - major[1][4] * scalar2;
    major[2][7] = major[2][7] * scalar1 - major[2][4] * scalar2;
    /* eliminate column 1 of rows 0 and 3 */
    scalar1 = minor[1][1]; scalar2 = minor[1][3];
    major[0][3] = major[0][3] * scalar1 - major[0][1] * scalar2;
    major[1][3] = major[1][3] * scalar1 - major[1][1] * scalar2;
    major[2][3] = major[2][3] * scalar1 - major[2][1] * scalar2;
    major[0][7] = major[0][7] * scalar1 - major[0][5] * scalar2;
    major[1][7] = major[1][7] * scalar1 - major[1][5] * scalar2;
    major[2][7] = major[2][7] * scalar1 - major[2][5] * scalar2;
    scalar2 = minor[1][0];
    minor[0][0] = minor[0][0] * scalar1 - minor[0][1] * scalar2;
    major[0][0] = major[0][0] * scalar1 - major[0][1] * scalar2;
    major[1][0] = major[1][0] * scalar1 - major[1][1] * scalar2;
    major[2][0] = major[2][0] * scalar1 - major[2][1] * scalar2;
    major[0][4] = major[0][4] * scalar1 - major[0][5] * scalar2;
    major[1][4] = major[1][4] * scalar1 - major[1][5] * scalar2;
    major[2][4] = major[2][4] * scalar1 - major[2][5] * scalar2;
    /* eliminate columns 0 and 1 of row 2 */
    scalar1 = 1.0f / minor[0][0];
    major[0][0] *= scalar1; major[1][0] *= scalar1; major[2][0] *= scalar1;
    major[0][4] *= scalar1; major[1][4] *= scalar1; major[2][4] *= scalar1;
    scalar1 = 1.0f / minor[1][1];
    major[0][1] *= scalar1; major[1][1] *= scalar1; major[2][1] *= scalar1;
    major[0][5] *= scalar1; major[1][5] *= scalar1; major[2][5] *= scalar1;
    scalar1 = minor[0][2]; scalar2 = minor[1][2];
    major[0][2] -= major[0][0] * scalar1 + major[0][1] * scalar2;
    major[1][2] -= major[1][0] * scalar1 + major[1][1] * scalar2;
    major[2][2] -= major[2][0] * scalar1 + major[2][1] * scalar2;
    major[0][6] -= major[0][4] * scalar1 + major[0][5] * scalar2;
    major[1][6] -= major[1][4] * scalar1 + major[1][5] * scalar2;
    major[2][6] -= major[2][4] * scalar1 + major[2][5] * scalar2;
    /* only major matters now: rows 3 and 7 correspond to the hollowed-out rows */
    scalar1 = major[0][7];
    major[1][7] /= scalar1;
    major[2][7] /= scalar1;
    scalar1 = major[0][0]; major[1][0] -= scalar1 * major[1][7]; major[2][0] -= scalar1 * major[2][7];
    scalar1 = major[0][1]; major[1][1] -= scalar1 * major[1][7]; major[2][1] -= scalar1 * major[2][7];
    scalar1 = major[0][2]; major[1][2] -= scalar1 * major[1][7]; major[2][2] -= scalar1 * major[2][7];
    scalar1 = major[0][3]; major[1][3] -= scalar1 * major[1][7]; major[2][3] -= scalar1 * major[2][7];
    scalar1 = major[0][4]; major[1][4] -= scalar1 * major[1][7]; major[2][4] -= scalar1 * major[2][7];
    scalar1 = major[0][5]; major[1][5] -= scalar1 * major[1][7]; major[2][5] -= scalar1 * major[2][7];
    scalar1 = major[0][6]; major[1][6] -= scalar1 * major[1][7]; major[2][6] -= scalar1 * major[2][7];
    /* one column left */
    scalar1 = major[1][3];
    major[2][3] /= scalar1;
    scalar1 = major[1][0]; major[2][0] -= scalar1 * major[2][3];
    scalar1 = major[1][1]; major[2][1] -= scalar1 * major[2][3];
    scalar1 = major[1][2]; major[2][2] -= scalar1 * major[2][3];
    scalar1 = major[1][4]; major[2][4] -= scalar1 * major[2][3];
    scalar1 = major[1][5]; major[2][5] -= scalar1 * major[2][3];
    scalar1 = major[1][6]; major[2][6] -= scalar1 * major[2][3];
    scalar1 = major[1][7]; major[2][7] -= scalar1 * major[2][3];
    H[0] = major[2][0]; H[1] = major[2][1]; H[2] = major[2][2];
    H[3] = major[2][4]; H[4] = major[2][5]; H[5] = major[2][6];
    H[6] = major[2][7]; H[7] = major[2][3]; H[8] = 1.0;
}

static int is_model_degenerate(const float *H)
{
    const float f = H[0] + H[1] + H[2] + H[3] + H[4] + H[5] + H[6] + H[7];
    return isnan(f);
}

static void evaluate_model_sprt(Rho *p)
{
    unsigned i;
    double lambda = 1.0;
    const float distSq = p->maxD * p->maxD;
    const float *src = p->src, *dst = p->dst, *H = p->Hc;
    char *inl = p->inl_c;
    p->numModels++;
    p->numInl_c = 0;
    p->Ntested = 0;
    p->good = 1;
    for (i = 0; i < p->N && p->good; i++) {
        const float x = src[i * 2], y = src[i * 2 + 1];
        const float X = dst[i * 2], Y = dst[i * 2 + 1];
        float reprojX = H[0] * x + H[1] * y + H[2];
        float reprojY = H[3] * x + H[4] * y + H[5];
        const float reprojZ = H[6] * x + H[7] * y + 1.0f;
        reprojX /= reprojZ;
        reprojY /= reprojZ;
        reprojX -= X;
        reprojY -= Y;
        reprojX *= reprojX;
        reprojY *= reprojY;
        const float reprojDist = reprojX + reprojY;
        const unsigned isInlier = reprojDist <= distSq;
        p->numInl_c += isInlier;
        *inl++ = (char)isInlier;
        lambda *= isInlier ? p->lambdaAccept : p->lambdaReject;
        p->good = lambda <= p->A;
    }
    p->Ntested = i;
    p->Ntestedtotal += i;
}

static void update_sprt(Rho *p)
{
    if (p->good) {
        if (p->numInl_c > p->numInl_b) {
            p->epsilon = (double)p->numInl_c / p->N;
            design_sprt(p);
        }
    } else {
        const double newDelta = (double)p->numInl_c / p->Ntested;
        if (newDelta > 0) {
            const double relChange = fabs(p->delta - newDelta) / p->delta;
            if (relChange > MIN_DELTA_CHNG) {
                p->delta = newDelta;
                design_sprt(p);
            }
        }
    }
}

static void n_star_optimize(Rho *p)
{
    const unsigned min_sample_length = 10 * 2;
    unsigned best_n = p->N, test_n = best_n, bestNumInl = p->numInl_b, testNumInl = bestNumInl;
    for (; test_n > min_sample_length && testNumInl; test_n--) {
        if (testNumInl * best_n > bestNumInl * test_n) {
            if (testNumInl < p->tbl[test_n]) break;
            best_n = test_n;
            bestNumInl = testNumInl;
        }
        testNumInl -= !!p->inl_b[test_n - 1];
    }
    if (bestNumInl * p->phMax > p->phNumInl * best_n) {
        p->phMax = best_n;
        p->phNumInl = bestNumInl;
        p->maxI = sac_calc_iter_bound(p->cfd, (double)p->phNumInl / p->phMax, SMPL_SIZE, p->maxI);
    }
}

/* ---- Levenberg-Marquardt refinement over the inliers of the best model (float) */
static void calc_jacobian_errors(const float *H, const float *src, const float *dst, const char *inl, unsigned N, float (*JtJ)[8], float *Jte, float *Sp)
{
    float S = 0.0f;
    if (JtJ) memset(JtJ, 0, 8 * 8 * sizeof(float));
    if (Jte) memset(Jte, 0, 8 * sizeof(float));
    for (unsigned i = 0; i < N; i++) {
        if (!inl[i]) continue;
        const float x = src[2 * i + 0], y = src[2 * i + 1], X = dst[2 * i + 0], Y = dst[2 * i + 1];
        const float W = H[6] * x + H[7] * y + 1.0f;
        const float iW = fabs(W) > FLT_EPSILON ? 1.0f / W : 0;
        const float reprojX = (H[0] * x + H[1] * y + H[2]) * iW;
        const float reprojY = (H[3] * x + H[4] * y + H[5]) * iW;
        const float eX = reprojX - X, eY = reprojY - Y;
        const float e = eX * eX + eY * eY;
        S += e;
        if (JtJ || Jte) {
            const float dxh11 = x * iW, dxh12 = y * iW, dxh13 = iW, dxh31 = -reprojX * x * iW, dxh32 = -reprojX * y * iW;
            const float dyh21 = x * iW, dyh22 = y * iW, dyh23 = iW, dyh31 = -reprojY * x * iW, dyh32 = -reprojY * y * iW;
            if (Jte) {
                Jte[0] += eX * dxh11; Jte[1] += eX * dxh12; Jte[2] += eX * dxh13;
                Jte[3] += eY * dyh21; Jte[4] += eY * dyh22; Jte[5] += eY * dyh23;
                Jte[6] += eX * dxh31 + eY * dyh31;
                Jte[7] += eX * dxh32 + eY * dyh32;
            }
            if (JtJ) {
                JtJ[0][0] += dxh11 * dxh11;
                JtJ[1][0] += dxh11 * dxh12; JtJ[1][1] += dxh12 * dxh12;
                JtJ[2][0] += dxh11 * dxh13; JtJ[2][1] += dxh12 * dxh13; JtJ[2][2] += dxh13 * dxh13;
                JtJ[3][3] += dyh21 * dyh21;
                JtJ[4][3] += dyh21 * dyh22; JtJ[4][4] += dyh22 * dyh22;
                JtJ[5][3] += dyh21 * dyh23; JtJ[5][4] += dyh22 * dyh23; JtJ[5][5] += dyh23 * dyh23;
                JtJ[6][0] += dxh11 * dxh31; JtJ[6][1] += dxh12 * dxh31; JtJ[6][2] += dxh13 * dxh31;
                JtJ[6][3] += dyh21 * dyh31; JtJ[6][4] += dyh22 * dyh31; JtJ[6][5] += dyh23 * dyh31;
                JtJ[6][6] += dxh31 * dxh31 + dyh31 * dyh31;
                JtJ[7][0] += dxh11 * dxh32; JtJ[7][1] += dxh12 * dxh32; JtJ[7][2] += dxh13 * dxh32;
                JtJ[7][3] += dyh21 * dyh32; JtJ[7][4] += dyh22 * dyh32; JtJ[7][5] += dyh23 * dyh32;
                JtJ[7][6] += dxh31 * dxh32 + dyh31 * dyh32;
                JtJ[7][7] += dxh32 * dxh32 + dyh32 * dyh32;
            }
        }
    }
    if (Sp) *Sp = S;
}

static int chol8_damped(const float (*A)[8], float lambda, float (*L)[8])
{
    const float lambdap1 = lambda + 1.0f;
    for (int i = 0; i < 8; i++) {
        for (int j = 0; j < i; j++) {
            float x = A[i][j];
            for (int k = 0; k < j; k++) x -= L[i][k] * L[j][k];
            L[i][j] = x / L[j][j];
        }
        {
            const int j = i;
            float x = A[j][j] * lambdap1;
            for (int k = 0; k < j; k++) x -= L[j][k] * L[j][k];
            if (x < 0) return 0;
            L[j][j] = sqrtf(x);
        }
    }
    return 1;
}

/* inverse of a lower-triangular 8x8, recursive block-wise (1x1 -> 2x2 -> 4x4 -> 8x8); L and M may alias.  The association
 * of every product below was fixed by comparing with cv2.findHomography(RHO) bit for bit (tests/test_rho_cpu.py). */
static void tr_inv8(const float (*L)[8], float (*M)[8])
{
    float s[2][2], t[2][2];
    float u[4][4], v[4][4];
    M[0][0] = 1.0f / L[0][0]; M[1][1] = 1.0f / L[1][1]; M[2][2] = 1.0f / L[2][2]; M[3][3] = 1.0f / L[3][3];
    M[4][4] = 1.0f / L[4][4]; M[5][5] = 1.0f / L[5][5]; M[6][6] = 1.0f / L[6][6]; M[7][7] = 1.0f / L[7][7];
    /* four 2x2 blocks */
    M[1][0] = -M[1][1] * L[1][0] * M[0][0];
    M[3][2] = -M[3][3] * L[3][2] * M[2][2];
    M[5][4] = -M[5][5] * L[5][4] * M[4][4];
    M[7][6] = -M[7][7] * L[7][6] * M[6][6];
    /* two 4x4 blocks: s = -C^-1 B, t = s A^-1 */
    for (int blk = 0; blk < 2; blk++) {
        const int o = 4 * blk;
        s[0][0] = -M[o + 2][o + 2] * L[o + 2][o + 0];
        s[0][1] = -M[o + 2][o + 2] * L[o + 2][o + 1];
        s[1][0] = -M[o + 3][o + 2] * L[o + 2][o + 0] + -M[o + 3][o + 3] * L[o + 3][o + 0];
        s[1][1] = -M[o + 3][o + 2] * L[o + 2][o + 1] + -M[o + 3][o + 3] * L[o + 3][o + 1];
        t[0][0] = s[0][0] * M[o + 0][o + 0] + s[0][1] * M[o + 1][o + 0];
        t[0][1] = s[0][1] * M[o + 1][o + 1];
        t[1][0] = s[1][0] * M[o + 0][o + 0] + s[1][1] * M[o + 1][o + 0];
        t[1][1] = s[1][1] * M[o + 1][o + 1];
        M[o + 2][o + 0] = t[0][0]; M[o + 2][o + 1] = t[0][1]; M[o + 3][o + 0] = t[1][0]; M[o + 3][o + 1] = t[1][1];
    }
    /* the 8x8: u = -C^-1 B, v = u A^-1 */
    for (int c = 0; c < 4; c++) {
        u[0][c] = -M[4][4] * L[4][c];
        u[1][c] = -M[5][4] * L[4][c] + -M[5][5] * L[5][c];
        u[2][c] = -M[6][4] * L[4][c] + -M[6][5] * L[5][c] + -M[6][6] * L[6][c];
        u[3][c] = -M[7][4] * L[4][c] + -M[7][5] * L[5][c] + -M[7][6] * L[6][c] + -M[7][7] * L[7][c];
    }
    for (int r = 0; r < 4; r++) {
        v[r][0] = u[r][0] * M[0][0] + u[r][1] * M[1][0] + u[r][2] * M[2][0] + u[r][3] * M[3][0];
        v[r][1] = u[r][1] * M[1][1] + u[r][2] * M[2][1] + u[r][3] * M[3][1];
        v[r][2] = u[r][2] * M[2][2] + u[r][3] * M[3][2];
        v[r][3] = u[r][3] * M[3][3];
    }
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) M[4 + r][c] = v[r][c];
}

/* dH = Linv^T Linv Jte (Linv = inverse of the Cholesky factor) */
static void tri_solve8(const float (*L)[8], const float *Jte, float *dH)
{
    float t[8];
    for (int i = 0; i < 8; i++) {
        float v = L[i][0] * Jte[0];
        for (int k = 1; k <= i; k++) v += L[i][k] * Jte[k];
        t[i] = v;
    }
    for (int i = 0; i < 8; i++) {
        float v = L[i][i] * t[i];
        for (int k = i + 1; k < 8; k++) v += L[k][i] * t[k];
        dH[i] = v;
    }
}

static float lm_gain(const float *dH, const float *Jte, float S, float newS, float lambda)
{
    const float dS = S - newS;
    float dL = 0;
    for (int i = 0; i < 8; i++) dL += dH[i] * dH[i];
    dL *= lambda;
    for (int i = 0; i < 8; i++) dL += dH[i] * Jte[i];
    dL *= 0.5f;
    return fabs(dL) < FLT_EPSILON ? dS : dS / dL;
}

static void refine(Rho *p)
{
    float S, newS, gain, L = 100.0f, dH[8], newH[8];
    float JtJ[8][8], tmp1[8][8], Jte[8];
    calc_jacobian_errors(p->Hb, p->src, p->dst, p->inl_b, p->N, JtJ, Jte, &S);
    for (int i = 0; i < MAXLEVMARQITERS; i++) {
        p->lm_iters++;
        while (!chol8_damped(JtJ, L, tmp1)) L *= 2.0f;
        tr_inv8(tmp1, tmp1);
        tri_solve8(tmp1, Jte, dH);
        for (int k = 0; k < 8; k++) newH[k] = p->Hb[k] - dH[k];
        calc_jacobian_errors(newH, p->src, p->dst, p->inl_b, p->N, NULL, NULL, &newS);
        gain = lm_gain(dH, Jte, S, newS, L);
        if (gain < LM_GAIN_LO) {
            L *= 8;
            if (L > 1000.0f / FLT_EPSILON) break;
        } else if (gain > LM_GAIN_HI) {
            L *= 0.5f;
        }
        if (gain > 0) {
            S = newS;
            memcpy(p->Hb, newH, sizeof(newH));
            calc_jacobian_errors(p->Hb, p->src, p->dst, p->inl_b, p->N, JtJ, Jte, &S);
        }
    }
}

/*
 * src, dst: N x 2 float (x, y).  H_out: 9 doubles (row-major, H[8] = 1) or all zeros when fewer than 4 inliers were found.
 * mask_out (optional): N bytes, 1 = inlier of the best model.  diag (optional, 8 unsigned): iterations run, models
 * evaluated, degenerate samples, degenerate models, best inlier count, final maxI, LM iterations, total points tested.
 * Returns the number of inliers (0 = failure), like rhoHest.
 */
unsigned rho_find_homography(const float *src, const float *dst, unsigned N, float maxD, unsigned maxI, double cfd, double beta,
                             double *H_out, unsigned char *mask_out, unsigned *diag)
{
    Rho r;
    memset(&r, 0, sizeof r);
    for (int k = 0; k < 9; k++) H_out[k] = 0.0;
    if (mask_out) memset(mask_out, 0, N);
    if (!src || !dst || N < SMPL_SIZE) return 0;
    r.src = src; r.dst = dst; r.N = N; r.maxD = maxD; r.maxI = maxI; r.rConvg = maxI; r.cfd = cfd; r.minInl = 4; r.beta = beta;
    fast_seed(&r, (uint64_t)~0ull);
    r.tbl = (unsigned *)calloc(N + 1, sizeof(unsigned));
    sac_init_nonrand(beta, 0, N, r.tbl);
    r.inl_c = (char *)calloc(N, 1);
    r.inl_b = (char *)calloc(N, 1);
    r.i = 0; r.phNum = SMPL_SIZE; r.phEndI = 1; r.phEndFpI = sac_init_pend_fpi(r.rConvg, N, SMPL_SIZE);
    r.phMax = N; r.phNumInl = 0; r.numModels = 0;
    r.numInl_c = r.numInl_b = 0;
    r.t_M = SPRT_T_M; r.m_S = SPRT_M_S; r.epsilon = SPRT_EPSILON; r.delta = SPRT_DELTA;
    design_sprt(&r);
    for (r.i = 0; r.i < r.maxI || r.i < 100; r.i++) {
        /* hypothesize */
        if (r.i >= r.phEndI && r.phNum < r.phMax) {
            r.phNum++;
            const double next = (r.phEndFpI * r.phNum) / (r.phNum - SMPL_SIZE);
            r.phEndI += (unsigned)ceil(next - r.phEndFpI);
            r.phEndFpI = next;
        }
        if (r.i > r.phEndI) rnd_smpl(&r, 4, r.smpl, r.phNum);
        else { rnd_smpl(&r, 3, r.smpl, r.phNum - 1); r.smpl[3] = r.phNum - 1; }
        if (is_sample_degenerate(&r)) { r.n_degenerate_sample++; continue; }
        h_func(r.pkd, r.Hc);
        if (is_model_degenerate(r.Hc)) { r.n_degenerate_model++; continue; }
        /* verify */
        evaluate_model_sprt(&r);
        update_sprt(&r);
        if (r.numInl_c > r.numInl_b) {
            float t[9];
            memcpy(t, r.Hc, sizeof t); memcpy(r.Hc, r.Hb, sizeof t); memcpy(r.Hb, t, sizeof t);
            char *ti = r.inl_c; r.inl_c = r.inl_b; r.inl_b = ti;
            r.numInl_b = r.numInl_c;
            r.maxI = sac_calc_iter_bound(r.cfd, (double)r.numInl_b / r.N, SMPL_SIZE, r.maxI);
            n_star_optimize(&r);
        }
    }
    const unsigned iters = r.i;
    if (r.numInl_b > (unsigned)SMPL_SIZE) refine(&r);     /* canRefine: best.numInl > SMPL_SIZE */
    const int good = r.numInl_b >= r.minInl;
    if (good) {
        for (int k = 0; k < 9; k++) H_out[k] = (double)r.Hb[k];
        if (mask_out) for (unsigned k = 0; k < N; k++) mask_out[k] = r.inl_b[k] ? 1 : 0;
    }
    if (diag) {
        diag[0] = iters; diag[1] = r.numModels; diag[2] = r.n_degenerate_sample; diag[3] = r.n_degenerate_model;
        diag[4] = r.numInl_b; diag[5] = r.maxI; diag[6] = r.lm_iters; diag[7] = r.Ntestedtotal;
    }
    const unsigned ret = good ? r.numInl_b : 0;
    free(r.tbl); free(r.inl_c); free(r.inl_b);
    return ret;
}
