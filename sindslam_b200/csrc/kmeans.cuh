// kmeans.cuh -- DynaDetect::SegByKmeans + cluster ordering (DynaDetect.cc:315-420,1425-1491).
#pragma once
#include "common.cuh"

#define KM_K 12

struct KmState {
    unsigned long long sums[2][KM_K * 3];  // 2^-36 fixed-point coordinate sums (ping-pong between centre passes)
    int counts[2][KM_K];
    float centers[KM_K * 3], old[KM_K * 3];
    int final_counts[KM_K];
    int done, iters;
};

struct ClusterOrder {
    int kept[KM_K];        // allLabels order: k-means cluster ids by ascending centre depth, <60 px dropped
    int kept_area[KM_K];
    int rank_of[KM_K];     // inverse of kept (-1 = dropped)
    uint8_t seg_edge_lut[KM_K];
    int n_kept, count0;
};

struct KmeansStage {
    int W = 0, H = 0;
    int lw[4], lh[4];
    uint16_t *depth_pyr[4] = {};
    int *labels[4] = {};
    float *points = nullptr, *points_lvl = nullptr;
    uint8_t *labels_u8 = nullptr, *seg_edge = nullptr;
    KmState *state = nullptr;     // one per level
    ClusterOrder *order = nullptr;
    int *nz_flag = nullptr;       // countNonZero(imgLabelLast)
};

int kmeans_init(sindyn_base *ctx, KmeansStage *k, int W, int H);
// depth: W x H u16 device; label_last: W x H u8 device (imgLabelLast)
int kmeans_run(sindyn_base *ctx, KmeansStage *k, const uint16_t *depth, const uint8_t *label_last, const sindyn_config *cfg);
