// decide.cuh -- mask fusion + per-cluster dynamic decision + final mask (DynaDetect.cc:1553-1636).
#pragma once
#include "ccl.cuh"
#include "common.cuh"

#define DD_MAXL 128   // merged cluster labels 1..DD_MAXL (one CCL plane each)

struct DecideStage {
    int W = 0, H = 0;
    void *ctl = nullptr;
    uint8_t *low0 = nullptr, *low = nullptr;   // fused low-error mask before / after the 5x5 dilation
    uint8_t *filled = nullptr, *dyna = nullptr, *tmp = nullptr, *out = nullptr;
    // (the two keyed planes -- cluster id where low == 128 / where low != 128, 255 = inactive -- and their labels are planes
    //  DD_MAXL, DD_MAXL + 1 of the caller's cls / labelsL scratch)
    uint8_t *seedflag = nullptr;               // 2 x N, indexed by key-region root
};

int decide_init(sindyn_base *ctx, DecideStage *d, int W, int H);
// cls / labelsL / stats / top: scratch planes shared with the re-clustering stage (DD_MAXL planes each; cls / labelsL: DD_MAXL + 2).
// low_in (0/128), high (0/255), high_last, total_area, labels: W x H u8 device.  Result: d->out ({0,125,255}).
int decide_run(sindyn_base *ctx, DecideStage *d, uint8_t *cls, int *labelsL, RegionStats *stats, int *top, const uint8_t *low_in,
               const uint8_t *high, const uint8_t *high_last, const uint8_t *total_area, const uint8_t *labels);
