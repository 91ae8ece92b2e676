// peac.cuh -- PEAC plane-contour edges (DynaDetect::CalOccluded, ORB_SLAM2/src/DynaDetect.cc:558-593;
// ORB_SLAM2/include/PEAC/AHCPlaneFitter.hpp, plane_fitter_pcl.hpp:275-317).
#pragma once
#include "common.cuh"

struct ReclusterStage;

struct PeacStage {
    int W = 0, H = 0;
    bool built = false;
    void *impl = nullptr;
};

int peac_init(sindyn_base *ctx, PeacStage *p, int W, int H);
// depth: W x H u16 device; plane_edges_out: W x H u8 device (imgEdgeByPlane)
// rc: scratch planes / CCL buffers of the re-clustering stage (the contour drawing reuses them)
int peac_run(sindyn_base *ctx, PeacStage *p, ReclusterStage *rc, const uint16_t *depth, float fx, float fy, float cx, float cy,
             float depth_scale, uint8_t *plane_edges_out);
// asynchronous copy of the control header {n_planes, n_final, overflow, n_ex} (valid after the stream has been synchronised)
int peac_copy_header(sindyn_base *ctx, PeacStage *p, int *host4);
// asynchronous copy of the instance's sticky overflow flag (an overflow in any frame it has processed so far)
int peac_copy_sticky_overflow(sindyn_base *ctx, PeacStage *p, int *host1);
// test hook: final membership image (final plane id or -1), extracted planes (rid, N, final id), counts
int peac_get_debug(sindyn_base *ctx, PeacStage *p, int *label_out, int *planes_rid_n, int *n_planes, int *n_final);
