// ccl.cu -- union-find region labelling (8-connected foreground + 4-connected background + exterior) and
// the per-quad contour statistics described in ccl.cuh.
#include "ccl.cuh"

__device__ __forceinline__ int uf_find(const int *L, int a)
{
    int p = L[a];
    while (p != a) { a = p; p = L[a]; }
    return a;
}

__device__ __forceinline__ void uf_union(int *L, int a, int b)
{
    for (;;) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a > b) { int t = a; a = b; b = t; }
        int old = atomicMin(&L[b], a);  // attach the larger root under the smaller
        if (old == b) return;
        b = old;
    }
}

// one warp = 32 consecutive pixels of a row: label = first pixel of the horizontal run inside the segment
// All CCL kernels are launched with gridDim.z = capacity planes but only *active planes do work: a fixed, small number
// of blocks per plane walks the plane with a grid-stride loop, so that inactive planes cost a handful of empty blocks
// instead of N/256 of them (the decision stage launches 128 planes for ~10 active ones).
#define CCL_BPP 96   // blocks per plane

// key_from: planes key_from, key_from + 1, ... are CCL_KEY8 planes that are always active (the decision stage labels its
// cluster planes and its two keyed planes in one launch chain)
__global__ void k_ccl_init(const uint8_t *__restrict__ cls, int *__restrict__ labels, int W, int H, const int *__restrict__ active, int key_from)
{
    const int plane = blockIdx.z;
    if (plane < key_from && active && plane >= *active) return;
    const int N = W * H;
    const int segs = (W + 31) >> 5;
    const int lane = threadIdx.x & 31;
    const uint8_t *c = cls + (size_t)plane * N;
    int *L = labels + (size_t)plane * (N + 1);
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; warp < segs * H; warp += nwarps) {
        const int y = warp / segs, x = (warp - y * segs) * 32 + lane;
        int v = x < W ? c[y * W + x] : 256 + lane;  // out-of-row lanes never match
        int pv = __shfl_up_sync(0xffffffffu, v, 1);
        bool start = lane == 0 || pv != v;
        unsigned m = __ballot_sync(0xffffffffu, start);
        int sl = 31 - __clz(m & (0xffffffffu >> (31 - lane)));
        if (x < W) L[y * W + x] = y * W + (x - lane + sl);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) L[N] = N;
}

__global__ void k_ccl_merge(const uint8_t *__restrict__ cls, int *__restrict__ labels, int W, int H, int mode, const int *__restrict__ active, int key_from)
{
    const int plane = blockIdx.z;
    if (plane < key_from && active && plane >= *active) return;
    if (plane >= key_from) mode = CCL_KEY8;
    const int N = W * H;
    const uint8_t *c = cls + (size_t)plane * N;
    int *L = labels + (size_t)plane * (N + 1);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
        const int y = p / W, x = p - y * W;
        const int v = c[p];
        if (mode == CCL_KEY8 && v == 255) continue;
        const bool eight = (mode == CCL_KEY8) || v == 1;
        const int vl = x > 0 ? c[p - 1] : -1, vu = y > 0 ? c[p - W] : -1;
        const int vul = (x > 0 && y > 0) ? c[p - W - 1] : -1, vur = (x < W - 1 && y > 0) ? c[p - W + 1] : -1;
        const int vr = x < W - 1 ? c[p + 1] : -1;
        if ((x & 31) == 0 && vl == v) uf_union(L, p, p - 1);                 // runs are cut at 32-px segment starts
        if (vu == v && !(vl == v && vul == v)) uf_union(L, p, p - W);         // a new vertical contact begins here
        if (eight) {
            if (vul == v && vu != v && vl != v) uf_union(L, p, p - W - 1);
            if (vur == v && vu != v && vr != v) uf_union(L, p, p - W + 1);
        }
        if (mode == CCL_REGION && v == 0 && (x == 0 || y == 0 || x == W - 1 || y == H - 1)) {
            // background on the image border belongs to the exterior; one union per border run is enough
            bool first = (y == 0 || y == H - 1) ? (x == 0 || vl != 0) : true;
            if (first) uf_union(L, p, N);
        }
    }
}

__global__ void k_ccl_flatten(int *__restrict__ labels, int N, const int *__restrict__ active, int key_from)
{
    const int plane = blockIdx.z;
    if (plane < key_from && active && plane >= *active) return;
    int *L = labels + (size_t)plane * (N + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= N; i += gridDim.x * blockDim.x) L[i] = uf_find(L, i);
}

static inline int ccl_blocks(int work_items) { int b = cdiv(work_items, 256); return b < CCL_BPP ? b : CCL_BPP; }

int ccl_run(sindyn_base *ctx, const uint8_t *cls, int *labels, int W, int H, int planes, int mode, const int *active_planes)
{
    return ccl_run_mixed(ctx, cls, labels, W, H, planes, 0, mode, active_planes);
}

int ccl_run_mixed(sindyn_base *ctx, const uint8_t *cls, int *labels, int W, int H, int planes, int key_planes, int mode, const int *active_planes)
{
    const int N = W * H;
    const int segs = (W + 31) >> 5;
    const int key_from = key_planes ? planes : 0x7fffffff, nz = planes + key_planes;
    LAUNCH(ctx, k_ccl_init, dim3(ccl_blocks(segs * H * 32), 1, nz), 256, 0, cls, labels, W, H, active_planes, key_from);
    LAUNCH(ctx, k_ccl_merge, dim3(ccl_blocks(N), 1, nz), 256, 0, cls, labels, W, H, mode, active_planes, key_from);
    LAUNCH(ctx, k_ccl_flatten, dim3(ccl_blocks(N + 1), 1, nz), 256, 0, labels, N, active_planes, key_from);
    LAUNCH_CHECK(ctx);
    return SINDYN_OK;
}

__global__ void k_ccl_top(const int *__restrict__ labels, int *__restrict__ top, int W, int H, const int *__restrict__ active,
                          RegionStats *__restrict__ zero_stats)
{
    const int plane = blockIdx.z;
    if (active && plane >= *active) return;
    const int N = W * H;
    const int *L = labels + (size_t)plane * (N + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        int t = rc_top(L, N, W, i);
        top[(size_t)plane * N + i] = t;
        // statistics are accumulated only at region roots: clear those entries here instead of a full memset
        if (zero_stats && L[i] == i) { RegionStats z; z.steps = 0; z.area2 = 0; zero_stats[(size_t)plane * N + i] = z; }
    }
}

int ccl_top_image(sindyn_base *ctx, const int *labels, int *top, int W, int H, int planes, const int *active_planes,
                  RegionStats *zero_stats)
{
    LAUNCH(ctx, k_ccl_top, dim3(ccl_blocks(W * H), 1, planes), 256, 0, labels, top, W, H, active_planes, zero_stats);
    LAUNCH_CHECK(ctx);
    return SINDYN_OK;
}

// ------------------------------------------------------------------ quad rules
// Quad with top-left pixel (x, y): a=(x,y) b=(x+1,y) c=(x,y+1) d=(x+1,y+1); m = membership bits a|b<<1|c<<2|d<<3.
// OUTER rule (polygon through member pixels): adds the step counts and the Green term x_P*y_Q - x_Q*y_P of each
// step P->Q, oriented so that cross(Q - P, outside direction) > 0.
__device__ __forceinline__ void quad_outer(int m, int x, int y, int &axis, int &diag, long long &g)
{
    const int ax = x, ay = y, bx = x + 1, by = y, cx = x, cy = y + 1, dx = x + 1, dy = y + 1;
#define GREEN(px, py, qx, qy) ((long long)(px) * (qy) - (long long)(qx) * (py))
    switch (m) {
    case 0x3: axis = 1; g = GREEN(ax, ay, bx, by); break;   // a,b set, outside below: a -> b
    case 0xC: axis = 1; g = GREEN(dx, dy, cx, cy); break;   // c,d set, outside above: d -> c
    case 0x5: axis = 1; g = GREEN(cx, cy, ax, ay); break;   // a,c set, outside right: c -> a
    case 0xA: axis = 1; g = GREEN(bx, by, dx, dy); break;   // b,d set, outside left: b -> d
    case 0xD: diag = 1; g = GREEN(dx, dy, ax, ay); break;   // b missing: d -> a
    case 0xB: diag = 1; g = GREEN(ax, ay, dx, dy); break;   // c missing: a -> d
    case 0xE: diag = 1; g = GREEN(bx, by, cx, cy); break;   // a missing: b -> c
    case 0x7: diag = 1; g = GREEN(cx, cy, bx, by); break;   // d missing: c -> b
    case 0x9: diag = 2; g = 0; break;                       // a,d only: both directions cancel
    case 0x6: diag = 2; g = 0; break;                       // b,c only
    default: break;
    }
}
// HOLE rule (polygon through the NON-member pixels surrounding the filled hole; m = membership in the hole):
// the same steps as the outer rule applied to the complement pattern, traversed in the opposite sense.
__device__ __forceinline__ void quad_hole(int m, int x, int y, int &axis, int &diag, long long &g)
{
    int cm = (~m) & 0xF;
    if (cm == 0x9 || cm == 0x6) {
        // two diagonal non-members: they are linked only if the two members belong to the same hole, which is what
        // m encodes (both member bits set) -> the border passes twice; nothing if they are different holes (m has one bit)
        diag = 2; g = 0;
        return;
    }
    if (m == 0x1 || m == 0x2 || m == 0x4 || m == 0x8) { quad_outer(cm, x, y, axis, diag, g); g = -g; return; }  // 1 member: diagonal cut
    if (m == 0x3 || m == 0xC || m == 0x5 || m == 0xA) { quad_outer(cm, x, y, axis, diag, g); g = -g; return; }  // 2 adjacent members
    // 3 members: a vertex of the border, no step; 0 or 4 members: nothing
}
#undef GREEN

__global__ void k_quad_external(const int *__restrict__ top, RegionStats *__restrict__ stats, int W, int H, const int *__restrict__ active)
{
    const int plane = blockIdx.z;
    if (active && plane >= *active) return;
    const int N = W * H;
    const int *T = top + (size_t)plane * N;
    // quads have top-left pixel (x, y) with x in [-1, W-1], y in [-1, H-1]
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < (W + 1) * (H + 1); q += gridDim.x * blockDim.x) {
    const int y = q / (W + 1) - 1, x = q - (y + 1) * (W + 1) - 1;
    auto at = [&](int xx, int yy) -> int { return (xx >= 0 && xx < W && yy >= 0 && yy < H) ? T[yy * W + xx] : -1; };
    const int ta = at(x, y), tb = at(x + 1, y), tc = at(x, y + 1), td = at(x + 1, y + 1);
    int X = ta >= 0 ? ta : (tb >= 0 ? tb : (tc >= 0 ? tc : td));
    if (X < 0) continue;
    // different top-level components are never 8-adjacent, so every non-negative id in the quad equals X
    int m = (ta == X) | ((tb == X) << 1) | ((tc == X) << 2) | ((td == X) << 3);
    if (m == 0xF) continue;
    int axis = 0, diag = 0;
    long long g = 0;
    quad_outer(m, x, y, axis, diag, g);
    if (axis | diag) {
        RegionStats *s = stats + (size_t)plane * N + X;
        atomicAdd(&s->steps, (unsigned long long)axis | ((unsigned long long)diag << 32));
        if (g) atomicAdd((unsigned long long *)&s->area2, (unsigned long long)g);
    }
    }
}

int ccl_quad_stats_external(sindyn_base *ctx, const int *top, RegionStats *stats, int W, int H, int planes, const int *active_planes)
{
    LAUNCH(ctx, k_quad_external, dim3(ccl_blocks((W + 1) * (H + 1)), 1, planes), 256, 0, top, stats, W, H, active_planes);
    LAUNCH_CHECK(ctx);
    return SINDYN_OK;
}

__global__ void k_quad_ccomp(const uint8_t *__restrict__ cls, const int *__restrict__ labels, RegionStats *__restrict__ stats, int W, int H,
                             const int *__restrict__ active)
{
    const int plane = blockIdx.z;
    if (active && plane >= *active) return;
    const int N = W * H;
    const int *L = labels + (size_t)plane * (N + 1);
    const uint8_t *c = cls + (size_t)plane * N;
    const int ext = L[N];
    for (int q_ = blockIdx.x * blockDim.x + threadIdx.x; q_ < (W + 1) * (H + 1); q_ += gridDim.x * blockDim.x) {
    const int y = q_ / (W + 1) - 1, x = q_ - (y + 1) * (W + 1) - 1;
    int r[4], par[4], fgv[4];
    const int xs[4] = {x, x + 1, x, x + 1}, ys[4] = {y, y, y + 1, y + 1};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        bool in = xs[k] >= 0 && xs[k] < W && ys[k] >= 0 && ys[k] < H;
        int p = ys[k] * W + xs[k];
        r[k] = in ? L[p] : ext;
        fgv[k] = in ? c[p] : 0;
        par[k] = r[k] == ext ? -1 : rc_parent(L, N, W, r[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int X = r[k];
        if (X == ext) continue;
        bool seen = false;
        for (int q = 0; q < k; ++q) seen |= r[q] == X;
        if (seen) continue;
        // membership in S(X): the region itself or a region nested directly inside it
        int m = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) m |= ((r[q] == X) || (par[q] == X)) << q;
        if (m == 0xF) continue;
        int axis = 0, diag = 0;
        long long g = 0;
        if (fgv[k]) quad_outer(m, x, y, axis, diag, g);
        else quad_hole(m, x, y, axis, diag, g);
        if (axis | diag) {
            RegionStats *s = stats + (size_t)plane * N + X;
            atomicAdd(&s->steps, (unsigned long long)axis | ((unsigned long long)diag << 32));
            if (g) atomicAdd((unsigned long long *)&s->area2, (unsigned long long)g);
        }
    }
    }
}

int ccl_quad_stats_ccomp(sindyn_base *ctx, const uint8_t *cls, const int *labels, RegionStats *stats, int W, int H, int planes,
                         const int *active_planes)
{
    LAUNCH(ctx, k_quad_ccomp, dim3(ccl_blocks((W + 1) * (H + 1)), 1, planes), 256, 0, cls, labels, stats, W, H, active_planes);
    LAUNCH_CHECK(ctx);
    return SINDYN_OK;
}
