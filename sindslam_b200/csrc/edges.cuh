// edges.cuh -- gradient depth edges + end points (first half of DynaDetect::CalOccluded, DynaDetect.cc:434-536).
#pragma once
#include "common.cuh"

#define EDGE_EP_CAP 8192  // candidate end points before NMS (power of two)

struct EdgeStage {
    int W = 0, H = 0;
    float *filtered = nullptr;      // medianBlur5(depth as float)
    uint8_t *total_area = nullptr;  // imgTotalArea (0/255)
    uint8_t *occl_raw = nullptr;    // before OPEN 4x4
    uint8_t *grad_edges = nullptr;  // imgOccluded after OPEN 4x4 (= imgOccludedForPlane)
    uint8_t *tmp = nullptr;
    int *ep_list = nullptr;         // unordered raster indices
    int *ep_xy = nullptr;           // after NMS: (x, y) pairs, raster order
    int *scalars = nullptr;         // [0] max depth bits [1] candidate count [2] kept count [3] overflow flag
};

int edges_init(sindyn_base *ctx, EdgeStage *e, int W, int H);
int edges_run(sindyn_base *ctx, EdgeStage *e, const uint16_t *depth, float depth_scale);
