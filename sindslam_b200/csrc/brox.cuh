// brox.cuh -- coarse-to-fine Brox variational optical flow (replaces cv::cuda::BroxOpticalFlow,
// ORB_SLAM2/src/DynaDetect.cc:1029,1072,1124).
#pragma once
#include "common.cuh"

#define BROX_MAX_LEVELS 96

struct BroxSolver {
    int w = 0, h = 0, nl = 0;
    int ws[BROX_MAX_LEVELS], hs[BROX_MAX_LEVELS];
    size_t off[BROX_MAX_LEVELS];  // element offset of level k inside pyr0/pyr1 (level 0 is the caller's buffer)
    float alpha, gamma, scale, omega;
    int inner, outer, solver;
    float *pyr0 = nullptr, *pyr1 = nullptr;
    // per-level scratch, sized for level 0
    float *A, *Iz, *Ix, *Iy, *Ixz, *Iyz, *Ixx, *Ixy, *Iyy;
    float *u[2], *v[2], *du[3], *dv[3];
    // per-pixel linear systems of one lagged-nonlinearity iteration (k_brox_system -> k_brox_sor), image layout
    float4 *sysW = nullptr;   // edge weights (left, right, up, down)
    float4 *sysC4 = nullptr;  // (j12, b1, b2, 1/d1)
    float *sysC1 = nullptr;   // 1/d2
    // captured CUDA graphs, keyed by the (I0, I1, out, sign) tuple: the frame ring of the handle rotates
    // through three pointer combinations, each gets its own graph
    struct GraphSlot {
        bool ok = false;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        const float *I0 = nullptr, *I1 = nullptr;
        float *out = nullptr;
        float sign = 0.f;
        unsigned long long launches = 0;
    } slots[4];
    int next_slot = 0;
    bool graph_ok = false;  // set to false to drop every cached graph (stream change)
    // measurement hook (sindyn_brox_profile): when non-null, every k_brox_sor launch is bracketed by events
    cudaEvent_t *prof_ev = nullptr;
    int prof_n = 0, prof_cap = 0;
    long long prof_px = 0;
};

int brox_num_levels_host(int w, int h, float scale, int outer, int *ws, int *hs);
int brox_init(sindyn_base *ctx, BroxSolver *b, int w, int h, float alpha, float gamma, float scale, int inner, int outer,
              int solver, float omega);
// I0/I1: dense w x h float device images. flow_out: w x h x 2 interleaved, multiplied by `sign`
// (the reference negates the solver output right away, DynaDetect.cc:1080).
int brox_run(sindyn_base *ctx, BroxSolver *b, const float *I0, const float *I1, float *flow_out, float sign, bool use_graph);
void brox_destroy(BroxSolver *b);

// generic pixel-centre aligned bilinear resample of `planes` float planes (cv::resize INTER_LINEAR rule, float path)
int launch_resample_f32(sindyn_base *ctx, const float *src, int sw, int sh, float *dst, int dw, int dh, float mul);
