// homography.cuh -- flow sampling + robust homography (DynaDetect.cc:1163-1235).
#pragma once
#include "common.cuh"

#define HG_MAX_SAMPLES 4096   // power of two >= ((W-1)/10)*((H-1)/10)
#define HG_GAUSS_N 4096
#define HG_M 2048             // hypotheses
#define HG_THR2 9.0           // (3 px)^2: cv::findHomography's default ransacReprojThreshold
#define HG_GN_ITERS 10
#define HG_TOPK 64             // best minimal-sample hypotheses that are refined before the final choice

struct HomographyStage {
    int W = 0, H = 0;
    int *counts = nullptr;       // [0..15] |label==i|, [16..31] |label==i && dyna==255|
    float *pts = nullptr, *pts_last = nullptr;  // ordered sample pairs (x,y)
    int *n_pairs = nullptr;      // [0] n, [1] best inlier count, [2] best hypothesis, [3] refined inlier count
    double *Hs = nullptr;
    int *scores = nullptr;
    double *H_dev = nullptr;     // 3x3 row-major result
    int *top = nullptr;          // HG_TOPK best hypotheses
    double *Hc = nullptr, *cand_cost = nullptr;  // refined candidates: 8 parameters; (consensus size, squared error)
};

int homography_init(sindyn_base *ctx, HomographyStage *g, int W, int H);
int homography_sample(sindyn_base *ctx, HomographyStage *g, const float *flow, const uint8_t *label_last, const uint8_t *dyna_last);
int homography_estimate(sindyn_base *ctx, HomographyStage *g);
