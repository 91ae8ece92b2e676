// morph.cuh -- cv::morphologyEx with getStructuringElement(MORPH_ELLIPSE, k x k), default anchor
// (ORB_SLAM2/src/DynaDetect.cc:51-59 and every morphologyEx call in that file; driver
// rgbd_tum_noros.cc:108,136-139).
#pragma once
#include "common.cuh"

enum { MORPH_DILATE = 0, MORPH_ERODE = 1, MORPH_OPEN = 2, MORPH_CLOSE = 3 };
#define MORPH_MAX_K 15

int morph_init(sindyn_base *ctx);
// src/dst dense W x H u8 device images; tmp is needed for open/close (may alias neither). src may equal dst
// only for open/close (the intermediate goes through tmp).
int morph_run(sindyn_base *ctx, const uint8_t *src, uint8_t *dst, uint8_t *tmp, int W, int H, int k, int op);
// host helper: row spans [j1, j2) of the k x k ellipse (the formula of cv::getStructuringElement)
void ellipse_spans(int k, int *j1, int *j2);

// Morphology on per-pixel membership BITSETS (bit c = "pixel belongs to image c"): dilate = OR, erode = AND over
// the same elliptic structuring element, out-of-image samples ignored -- i.e. up to 16 / 128 of the reference's
// per-cluster morphologyEx calls (DynaDetect.cc:671,683,688) in one pass.  elem_bytes = 2 (uint16_t) or 16 (ulonglong2).
int morph_bits_run(sindyn_base *ctx, const void *src, void *dst, int W, int H, int k, bool erode, int elem_bytes);
