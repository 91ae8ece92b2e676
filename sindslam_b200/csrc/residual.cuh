// residual.cuh -- flow residual + thresholds + masks (DynaDetect.cc:1236-1367).
#pragma once
#include "common.cuh"

struct ResidualStage {
    int W = 0, H = 0;
    float *mag = nullptr;        // residual magnitude, W x H
    uint8_t *m8 = nullptr;       // normalised 8-bit magnitude
    unsigned int *hist = nullptr, *gmax = nullptr;  // 256 bins + float-bits max
    float *thr = nullptr;        // {otsu, triangle, t_low, t_high}
};

int residual_init(sindyn_base *ctx, ResidualStage *r, int W, int H);
int residual_homography_run(sindyn_base *ctx, ResidualStage *r, const float *flow, const double *Hm, uint8_t *low, uint8_t *high);
int residual_homography_run_dev(sindyn_base *ctx, ResidualStage *r, const float *flow, const double *H_dev, uint8_t *low, uint8_t *high);
int residual_pose_run(sindyn_base *ctx, ResidualStage *r, const float *flow, const uint16_t *depth, const double *T,
                      float fx, float fy, float cx, float cy, float depth_scale, uint8_t *low, uint8_t *high);
