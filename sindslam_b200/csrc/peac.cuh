// peac.cuh -- PEAC plane-contour edges (DynaDetect::CalOccluded, ORB_SLAM2/src/DynaDetect.cc:558-593;
// ORB_SLAM2/include/PEAC/AHCPlaneFitter.hpp, plane_fitter_pcl.hpp:275-317).
#pragma once
#include "common.cuh"

struct PeacStage {
    int W = 0, H = 0;
    bool built = false;
    void *impl = nullptr;
};

int peac_init(sindyn_base *ctx, PeacStage *p, int W, int H);
// depth: W x H u16 device; plane_edges_out: W x H u8 device (imgEdgeByPlane)
int peac_run(sindyn_base *ctx, PeacStage *p, const uint16_t *depth, float fx, float fy, float cx, float cy, float depth_scale,
             uint8_t *plane_edges_out);
