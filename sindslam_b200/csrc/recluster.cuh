// recluster.cuh -- geometric re-clustering: split the k-means clusters at depth edges, build the region
// adjacency graph and merge (DynaDetect::SegAndMergeV2, ORB_SLAM2/src/DynaDetect.cc:653-1018), plus the
// plane-edge filter of CalOccluded (DynaDetect.cc:598-641).
#pragma once
#include "ccl.cuh"
#include "common.cuh"
#include "kmeans.cuh"

#define RC_MAXC 128          // qualified components (2 x 64-bit membership words per pixel)
#define RC_MAXP 11           // k-means clusters processed (all but the farthest)
#define RC_PF_MAXC 64        // plane-edge contours considered by the filter

struct Bits128 { unsigned long long w[2]; };

struct ReclusterControl {     // small device-resident control block
    int n_planes;             // clusters processed = n_kept - 1
    int n_comp;               // qualified components C
    int overflow;             // capacity flags
    int n_labels;             // merged label count
    int comp_plane[RC_MAXC], comp_root[RC_MAXC];
    Bits128 plane_bits[RC_MAXP + 1];
    int area[RC_MAXC];
    long long zsum[RC_MAXC];
    int cnt1[RC_MAXC];        // |temp1_c|
    int lj_area[RC_MAXC];
    int n_planes2;            // = n_comp (active planes of the lianjie CCL)
    int order[RC_MAXC];       // sorted rank -> component
    int rank_of[RC_MAXC];     // component -> sorted rank
    float score[RC_MAXC];
    uint8_t lut[RC_MAXC + 2]; // sorted rank -> merged label
    unsigned int depth_max;
    int pf_n;                 // plane-edge filter: contours kept
};

struct ReclusterStage {
    int W = 0, H = 0;
    ReclusterControl *ctl = nullptr;
    uint16_t *mcl = nullptr, *mcl_tmp = nullptr;      // per-pixel cluster-plane bitmask (before/after OPEN 4x4)
    uint8_t *cls = nullptr;                           // RC_MAXC planes of 0/1 (also used with RC_MAXP planes)
    int *labels = nullptr, *top = nullptr;            // RC_MAXC x (N+1), RC_MAXC x N
    RegionStats *stats = nullptr;                     // RC_MAXC x N
    int *comp_of = nullptr;                           // RC_MAXP x N : root pixel -> component index
    Bits128 *F = nullptr, *CI = nullptr, *CD = nullptr, *T1 = nullptr, *LJ = nullptr, *tmpb = nullptr;
    uint8_t *occl_dil = nullptr, *seg_dil = nullptr, *depth_norm = nullptr, *tmp8 = nullptr, *tmp8b = nullptr;
    int *hist = nullptr;                              // RC_MAXC x 256
    int *ov = nullptr, *ove = nullptr, *lo = nullptr; // RC_MAXC x RC_MAXC pair counts
    float *Tmat = nullptr;                            // (RC_MAXC+1)^2 RAG matrix (debug / tests)
    uint8_t *label_out = nullptr;
    // plane-edge filter scratch
    uint8_t *pf_e = nullptr, *occl1 = nullptr, *occl2 = nullptr;
};

int recluster_init(sindyn_base *ctx, ReclusterStage *r, int W, int H);
// plane_edges / grad_edges: W x H u8 device; endpoints: device (x,y) pairs + device count. Results: r->occl1 / r->occl2.
int plane_edge_filter_run(sindyn_base *ctx, ReclusterStage *r, const uint8_t *plane_edges, const uint8_t *grad_edges, const int *ep_xy,
                          const int *ep_n);
// labels_km / order / seg_edge(undilated) / points from the k-means stage; occl1/occl2 W x H u8 device. Result: r->label_out.
int recluster_run(sindyn_base *ctx, ReclusterStage *r, const KmeansStage *km, const uint8_t *occl1, const uint8_t *occl2,
                  const uint16_t *depth);
