// preproc.cuh -- BGR->gray, OpenCV-exact u8 bilinear resize, float flow resize.
#pragma once
#include "common.cuh"

struct ResizePlanU8 {
    int sw = 0, sh = 0, dw = 0, dh = 0;
    int *xofs = nullptr, *yofs = nullptr;
    short *xw = nullptr, *yw = nullptr;
};

int launch_bgr2gray(sindyn_base *ctx, const uint8_t *bgr, int W, int H, uint8_t *gray);
int resize_plan_init(sindyn_base *ctx, ResizePlanU8 *p, int sw, int sh, int dw, int dh);
int launch_resize_u8(sindyn_base *ctx, const ResizePlanU8 *p, const uint8_t *src, int spitch, uint8_t *dst, int dpitch,
                     float *dst_f32, float fscale);
int launch_resize_flow(sindyn_base *ctx, const float *src, int sw, int sh, float *dst, int dw, int dh, float mul);
