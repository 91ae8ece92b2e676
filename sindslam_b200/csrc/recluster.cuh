// recluster.cuh -- geometric re-clustering: split the k-means clusters at depth edges, build the region
// adjacency graph and merge (DynaDetect::SegAndMergeV2, ORB_SLAM2/src/DynaDetect.cc:653-1018), plus the
// plane-edge filter of CalOccluded (DynaDetect.cc:598-641).
//
// The reference keeps one full-frame image per component (cluster, dilated cluster, fake-edge strip) and runs
// morphology / AND / countNonZero per component and per PAIR of components.  Here every pixel carries a 128-bit
// membership set instead (bit c = "pixel belongs to component c's image"), so each of those per-component image
// operations is ONE pass over one bitset plane, and the O(C^2) pair loop becomes one pass that counts bit pairs.
#pragma once
#include "ccl.cuh"
#include "common.cuh"
#include "kmeans.cuh"

#define RC_MAXC 128          // qualified components (bits of the membership set)
#define RC_MAXP 11           // k-means clusters processed (all but the farthest kept one)
#define RC_PF_MAXC 64        // plane-edge contours considered by the filter (DynaDetect.cc:607-636)

struct ReclusterControl {     // small device-resident control block
    int n_planes;             // clusters processed = allLabels.size() - 1 (DynaDetect.cc:664)
    int n_raw;                // qualified contours found (unordered)
    int n_comp;               // C = min(n_raw, RC_MAXC)
    int overflow;             // n_raw > RC_MAXC
    int n_labels;             // merged label count
    unsigned int depth_max;
    unsigned long long raw_key[RC_MAXC];  // plane << 32 | root pixel
    int comp_plane[RC_MAXC], comp_root[RC_MAXC];
    ulonglong2 plane_bits[RC_MAXP + 1];   // components of each k-means plane
    int area[RC_MAXC];
    long long zsum[RC_MAXC];  // 2^-36 fixed-point sum of the weighted depth of the component's points
    int cnt1[RC_MAXC];        // |temp1_c| (DynaDetect.cc:694-697)
    int lj_area[RC_MAXC];
    int order[RC_MAXC];       // sorted rank -> component
    int rank_of[RC_MAXC];     // component -> sorted rank
    float score[RC_MAXC];
    uint8_t lut[RC_MAXC + 2]; // sorted rank -> merged label
    int pf_n, pf_overflow;    // plane-edge filter: contours with >= 25 points
    int pf_root[RC_PF_MAXC];
    int pf_hit[RC_PF_MAXC];
};

struct ReclusterStage {
    int W = 0, H = 0;
    ReclusterControl *ctl = nullptr;
    int8_t *plane_of = nullptr;                       // k-means plane of every pixel (-1 = not processed)
    uint16_t *mcl = nullptr, *mcl_tmp = nullptr;      // per-pixel plane bitmask (cluster - edges, before/after OPEN 4x4)
    uint8_t *cls = nullptr;                           // RC_MAXC planes of 0/1
    int *labels = nullptr, *top = nullptr;            // RC_MAXC x (N+1), RC_MAXC x N
    RegionStats *stats = nullptr;                     // RC_MAXC x N (valid at region roots)
    ulonglong2 *F = nullptr, *CI = nullptr, *CD = nullptr, *T1 = nullptr, *LJ = nullptr, *tmpb = nullptr, *tmpb2 = nullptr;
    uint8_t *occl_dil = nullptr, *seg_dil = nullptr, *tmp8 = nullptr, *tmp8b = nullptr;
    int *hist = nullptr;                              // RC_MAXC x 256
    int *ov = nullptr, *ove = nullptr, *lo = nullptr; // RC_MAXC x RC_MAXC pair counts (component index space)
    float *Tmat = nullptr;                            // (RC_MAXC+1)^2 RAG matrix in rank space
    uint8_t *label_out = nullptr;
    // plane-edge filter
    uint8_t *pf_e = nullptr, *occl1 = nullptr, *occl2 = nullptr;
};

int recluster_init(sindyn_base *ctx, ReclusterStage *r, int W, int H);
// plane_edges / grad_edges: W x H u8 device; endpoints: device (x,y) pairs + device count. Results: r->occl1 / r->occl2.
int plane_edge_filter_run(sindyn_base *ctx, ReclusterStage *r, const uint8_t *plane_edges, const uint8_t *grad_edges, const int *ep_xy,
                          const int *ep_n);
// labels / order / seg_edge (undilated) / points from the k-means stage; occl1/occl2 W x H u8 device. Result: r->label_out.
int recluster_run(sindyn_base *ctx, ReclusterStage *r, const KmeansStage *km, const uint8_t *occl1, const uint8_t *occl2,
                  const uint16_t *depth);
// PEAC output stage (AHCPlaneFitter.hpp:366-399): per-plane membership bitset PB (bit f = final plane f, *n_planes_dev
// planes) -> CLOSE 3x3 -> external contours -> drawContours(thickness 2) of all planes into out (0/255).
int plane_contours_run(sindyn_base *ctx, ReclusterStage *r, const ulonglong2 *PB, const int *n_planes_dev, uint8_t *out);
