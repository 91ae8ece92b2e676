// ccl.cuh -- connected-region labelling and closed-form contour statistics.
//
// This module is what replaces cv::findContours / contourArea / arcLength / drawContours in
// ORB_SLAM2/src/DynaDetect.cc (:605,:617,:675-690,:700-713,:1579-1602).  OpenCV's border following is a
// raster-sequential algorithm; the quantities the reference consumes are re-expressed as parallel image
// operations (all verified exactly against cv2 on random blob images, see tests/test_contours_cpu.py):
//   * regions: one union-find pass labels 8-connected foreground components AND 4-connected background
//     regions; background touching the image border is the exterior.  A region's root is its raster-first
//     pixel, so parent(region) = region of the pixel left of the root (Suzuki's nesting).
//   * RETR_EXTERNAL contour c  <->  top-level foreground component A;  drawContours(FILLED) = F(A) = A plus
//     everything nested in it.
//   * contours[c].size() = #axis steps + #diagonal steps, arcLength = axis + sqrt(2)*diag, where steps are
//     counted per 2x2 pixel quad of F(A): two adjacent set -> 1 axis step, three set -> 1 diagonal step,
//     two diagonal set -> 2 diagonal steps.  contourArea = |Green sum over the same boundary quads| / 2.
//   * RETR_CCOMP additionally yields one hole contour per non-exterior background region, with the same
//     quad rules applied to the filled hole.
//   * drawContours(thickness = 2) = cross-dilation of the boundary pixels of F(A) plus a 4x4-minus-corners
//     block around every quad that carries a diagonal step.
#pragma once
#include "common.cuh"

enum { CCL_REGION = 0, CCL_KEY8 = 1 };

struct RegionStats {       // per (plane, root pixel)
    unsigned long long steps;  // low 32 bits: axis steps, high 32 bits: diagonal steps
    long long area2;           // signed Green sum = 2 * contourArea
};

// cls: planes x N bytes. CCL_REGION: 0 = background, 1 = foreground. CCL_KEY8: 255 = inactive, else key.
// labels: planes x (N + 1) ints (index N is the exterior node in CCL_REGION mode).
// active_planes (device, may be null): planes >= *active_planes exit immediately.
int ccl_run(sindyn_base *ctx, const uint8_t *cls, int *labels, int W, int H, int planes, int mode, const int *active_planes);
// the same with key_planes further planes (index planes, planes + 1, ...) that are CCL_KEY8 and always active: one launch chain
int ccl_run_mixed(sindyn_base *ctx, const uint8_t *cls, int *labels, int W, int H, int planes, int key_planes, int mode, const int *active_planes);

// Region helpers usable from other translation units (device inline)
__device__ __forceinline__ int rc_exterior(const int *L, int N) { return L[N]; }
// parent region of the region with root pixel r (r != exterior root); returns the exterior root for top-level regions
__device__ __forceinline__ int rc_parent(const int *L, int N, int W, int r)
{
    int x = r % W;
    return x == 0 ? L[N] : L[r - 1];
}
// root of the top-level foreground component enclosing pixel p, or -1 if p is exterior
__device__ __forceinline__ int rc_top(const int *L, int N, int W, int p)
{
    const int ext = L[N];
    int r = L[p];
    if (r == ext) return -1;
    for (int it = 0; it < 64; ++it) {
        int pr = rc_parent(L, N, W, r);
        if (pr == ext) return r;
        r = pr;
    }
    return r;
}

// top[plane][p] = rc_top; zero_stats (may be null): clears stats[plane][root] of every region root (fg and bg)
int ccl_top_image(sindyn_base *ctx, const int *labels, int *top, int W, int H, int planes, const int *active_planes,
                  RegionStats *zero_stats);

// Quad statistics.
//  mode 0 (EXTERNAL): stats of top-level filled components, indexed by their root; `top` from ccl_top_image.
//  mode 1 (CCOMP):    stats of every foreground component (outer border of the component with all nested
//                     regions filled) and every hole (hole border), indexed by the region root; uses labels+cls.
int ccl_quad_stats_external(sindyn_base *ctx, const int *top, RegionStats *stats, int W, int H, int planes, const int *active_planes);
int ccl_quad_stats_ccomp(sindyn_base *ctx, const uint8_t *cls, const int *labels, RegionStats *stats, int W, int H, int planes,
                         const int *active_planes);

// First pixel, in OpenCV's border-following order (icvFetchContour), of the contour that starts at pixel
// `start` for which hit(p) is true; -1 if none.  fg: one plane of the 0/1 image.  Serial; meant to be called by
// one thread per contour with an early exit.
template <class Hit>
__device__ int rc_trace_first_hit(const uint8_t *fg, int W, int H, int start, bool is_hole, Hit hit)
{
    // direction codes of OpenCV: 0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE
    const int dx[8] = {1, 1, 0, -1, -1, -1, 0, 1};
    const int dy[8] = {0, -1, -1, -1, 0, 1, 1, 1};
    auto at = [&](int x, int y) -> int { return (x >= 0 && x < W && y >= 0 && y < H) ? fg[y * W + x] : 0; };
    const int x0 = start % W, y0 = start / W;
    int s_end = is_hole ? 0 : 4, s = s_end;
    int x1, y1;
    do {
        s = (s - 1) & 7;
        x1 = x0 + dx[s];
        y1 = y0 + dy[s];
    } while (at(x1, y1) == 0 && s != s_end);
    if (s == s_end) return hit(start) ? start : -1;  // single-pixel contour
    int x3 = x0, y3 = y0;
    for (int guard = 0; guard < 4 * (W + H) * 8; ++guard) {
        int x4, y4;
        for (;;) {
            ++s;
            x4 = x3 + dx[s & 7];
            y4 = y3 + dy[s & 7];
            if (at(x4, y4) != 0) break;
        }
        s &= 7;
        int p3 = y3 * W + x3;
        if (hit(p3)) return p3;
        if (x4 == x0 && y4 == y0 && x3 == x1 && y3 == y1) break;
        x3 = x4;
        y3 = y4;
        s = (s + 4) & 7;
    }
    return -1;
}
