// homography.cuh -- flow sampling + robust homography (DynaDetect.cc:1163-1235).
#pragma once
#include "common.cuh"

#define HG_MAX_SAMPLES 4096   // power of two >= ((W-1)/10)*((H-1)/10)
#define HG_GAUSS_N 4096

struct HomographyStage {
    int W = 0, H = 0;
    int *counts = nullptr;       // [0..15] |label==i|, [16..31] |label==i && dyna==255|
    float *pts = nullptr, *pts_last = nullptr;  // ordered sample pairs (x,y)
    int *n_pairs = nullptr;      // [0] n, [1] inliers of the best model (0 = failure), [2] models evaluated, [3] refinement iterations
    double *H_dev = nullptr;     // 3x3 row-major result
    unsigned char *inl_mask = nullptr;   // inlier flags of the best model, one per sample
};

int homography_init(sindyn_base *ctx, HomographyStage *g, int W, int H);
int homography_sample(sindyn_base *ctx, HomographyStage *g, const float *flow, const uint8_t *label_last, const uint8_t *dyna_last);
int homography_estimate(sindyn_base *ctx, HomographyStage *g);
