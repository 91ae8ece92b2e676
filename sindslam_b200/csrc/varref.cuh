// varref.cuh -- cv::VariationalRefinement::create()->calc(I0, I1, flow) equivalent (DynaDetect.cc:1133-1143).
#pragma once
#include "common.cuh"

struct VarRefStage {
    int w = 0, h = 0;
    bool built = false;
    float *buf = nullptr;
    float *planes[32] = {};
    float **planes_dev = nullptr;   // device copy of the plane table
};

int varref_init(sindyn_base *ctx, VarRefStage *v, int w, int h);
// I0, I1: u8 w x h device images (I0 = current); flow: w x h x 2 interleaved, refined in place
int varref_run(sindyn_base *ctx, VarRefStage *v, const uint8_t *I0, const uint8_t *I1, float *flow);
// same with the reference image chosen on the device: I1_alt when *sel != 0 (sel may be null)
int varref_run_sel(sindyn_base *ctx, VarRefStage *v, const uint8_t *I0, const uint8_t *I1, const uint8_t *I1_alt, const int *sel, float *flow);
