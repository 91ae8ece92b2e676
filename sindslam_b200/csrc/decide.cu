// decide.cu -- mask fusion, per-cluster dynamic decision and the final mask
// (DynaDetect::DetectDynaArea, ORB_SLAM2/src/DynaDetect.cc:1553-1636).
//
// The reference loops over the merged clusters; for each one it runs findContours(RETR_CCOMP) on
// (cluster & highError), and for every qualifying blob flood-fills the low-error mask inside the cluster
// from the first contour pixel whose low value is 128.  Here:
//   * one CCL pass with a plane per cluster gives all blobs and holes at once; contourArea / arcLength come from the
//     quad statistics (ccl.cuh); only the seed search follows the border (one thread per qualifying contour);
//   * floodFill(lo = up = 5, 8-connectivity, MASK_ONLY, mask = ~cluster) on an image with values {0, 128} is the
//     8-connected component of (low == low[seed]) & cluster that contains the seed: two keyed CCL
//     planes (value 128 / value 0) label all candidate regions, seeds flag their roots.
#include "decide.cuh"

#include "morph.cuh"

struct DecideControl {
    int nmax;                         // maxNumClusteri (DynaDetect.cc:1522-1524)
    int cnt_label[DD_MAXL + 1];       // |cluster n|
    int cnt_high[DD_MAXL + 1];        // |cluster n & highError|
    int cnt_fill[DD_MAXL + 1];        // |flood-filled pixels of cluster n|
};

__global__ void k_dd_reset(DecideControl *ctl)
{
    int t = threadIdx.x;
    if (t == 0) ctl->nmax = 0;
    if (t <= DD_MAXL) { ctl->cnt_label[t] = 0; ctl->cnt_high[t] = 0; ctl->cnt_fill[t] = 0; }
}

// low = ((highLast | low) != 0 ? 128 : 0) & totalArea (DynaDetect.cc:1549-1552) + per-cluster counts
__global__ void k_dd_low_counts(const uint8_t *__restrict__ low_in, const uint8_t *__restrict__ high, const uint8_t *__restrict__ high_last,
                                const uint8_t *__restrict__ total_area, const uint8_t *__restrict__ labels, int n, uint8_t *__restrict__ low0,
                                DecideControl *ctl)
{
    __shared__ int s_l[DD_MAXL + 1], s_h[DD_MAXL + 1];
    __shared__ int s_max;
    for (int j = threadIdx.x; j <= DD_MAXL; j += blockDim.x) { s_l[j] = 0; s_h[j] = 0; }
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint8_t v = (high_last[i] | low_in[i]) ? 128 : 0;
        low0[i] = v & total_area[i];
        int l = labels[i];
        if (l > 0 && l <= DD_MAXL) {
            atomicAdd(&s_l[l], 1);
            if (high[i]) atomicAdd(&s_h[l], 1);
            if (l > s_max) atomicMax(&s_max, l);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j <= DD_MAXL; j += blockDim.x) {
        if (s_l[j]) atomicAdd(&ctl->cnt_label[j], s_l[j]);
        if (s_h[j]) atomicAdd(&ctl->cnt_high[j], s_h[j]);
    }
    if (threadIdx.x == 0 && s_max) atomicMax(&ctl->nmax, s_max);
}

// plane n-1 = cluster n & highError (only when it has > 100 px, DynaDetect.cc:1563); key planes of the dilated low mask
__global__ void k_dd_planes(const uint8_t *__restrict__ labels, const uint8_t *__restrict__ high, const uint8_t *__restrict__ low, int n,
                            const DecideControl *__restrict__ ctl, uint8_t *__restrict__ cls, uint8_t *__restrict__ key)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int nmax = ctl->nmax;
    const int l = labels[i];
    const bool h = high[i] != 0;
    for (int pl = 0; pl < nmax; ++pl) cls[(size_t)pl * n + i] = (l == pl + 1 && h && ctl->cnt_high[pl + 1] > 100) ? 1 : 0;
    const bool on = l > 0 && l <= DD_MAXL;
    const bool v128 = low[i] == 128;
    key[i] = (on && v128) ? (uint8_t)l : (uint8_t)255;
    key[(size_t)n + i] = (on && !v128) ? (uint8_t)l : (uint8_t)255;
}

// one thread per region root of every cluster plane: contour tests + seed search (DynaDetect.cc:1570-1596)
__global__ void k_dd_seeds(const uint8_t *__restrict__ cls, const int *__restrict__ L_all, const RegionStats *__restrict__ stats,
                           const uint8_t *__restrict__ low, const uint8_t *__restrict__ labels, const int *__restrict__ keyL, int W, int H,
                           const DecideControl *__restrict__ ctl, uint8_t *__restrict__ seedflag)
{
    const int pl = blockIdx.z;
    if (pl >= ctl->nmax) return;
    const int N = W * H;
    const int *L = L_all + (size_t)pl * (N + 1);
    const uint8_t *fg = cls + (size_t)pl * N;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    if (L[i] != i || i == L[N]) continue;               // not a root, or the exterior
    const bool is_hole = fg[i] == 0;
    const RegionStats s = stats[(size_t)pl * N + i];
    const double axis = (double)(s.steps & 0xffffffffull), diag = (double)(s.steps >> 32);
    const double area = (double)(s.area2 < 0 ? -s.area2 : s.area2) * 0.5;
    const double len = axis + diag * (double)1.41421354f;   // arcLength: float segment lengths summed in double
    const double roundness = (4.0 * 3.141592653589793 * area) / (len * len);
    if (!((area > 100.0 && roundness > 0.2) || area > 2000.0)) continue;
    const int start = is_hole ? i - 1 : i;
    int seed = rc_trace_first_hit(fg, W, H, start, is_hole, [&](int p) { return low[p] == 128; });
    int plane2 = 0;
    if (seed < 0) {
        // cv::Point2f seedPoint stays (0,0): floodFill starts there if the cluster contains that pixel
        if (labels[0] != pl + 1) continue;
        seed = 0;
        plane2 = low[0] == 128 ? 0 : 1;
    }
    seedflag[(size_t)plane2 * N + keyL[(size_t)plane2 * (N + 1) + seed]] = 1;
    }
}

__global__ void k_dd_filled(const uint8_t *__restrict__ labels, const uint8_t *__restrict__ low, const int *__restrict__ keyL,
                            const uint8_t *__restrict__ seedflag, int n, DecideControl *ctl, uint8_t *__restrict__ filled)
{
    __shared__ int s_f[DD_MAXL + 1];
    for (int j = threadIdx.x; j <= DD_MAXL; j += blockDim.x) s_f[j] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int l = labels[i];
        uint8_t f = 0;
        if (l > 0 && l <= DD_MAXL) {
            const int pl = low[i] == 128 ? 0 : 1;
            f = seedflag[(size_t)pl * n + keyL[(size_t)pl * (n + 1) + i]];
            if (f) atomicAdd(&s_f[l], 1);
        }
        filled[i] = f;
    }
    __syncthreads();
    for (int j = threadIdx.x; j <= DD_MAXL; j += blockDim.x)
        if (s_f[j]) atomicAdd(&ctl->cnt_fill[j], s_f[j]);
}

// filled > 0.5 * |cluster| -> whole cluster, else the filled pixels (DynaDetect.cc:1612-1619)
__global__ void k_dd_dyna(const uint8_t *__restrict__ labels, const uint8_t *__restrict__ filled, int n, const DecideControl *__restrict__ ctl,
                          uint8_t *__restrict__ dyna)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = labels[i];
    bool on = false;
    if (l > 0 && l <= DD_MAXL) on = ((double)ctl->cnt_fill[l] > 0.5 * (double)ctl->cnt_label[l]) || filled[i];
    dyna[i] = on ? 255 : 0;
}

// imgDyna += (imgTotalArea - imgDyna) * 125/255 (DynaDetect.cc:1633-1634)
__global__ void k_dd_final(const uint8_t *__restrict__ dyna, const uint8_t *__restrict__ total_area, int n, uint8_t *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = dyna[i], t = total_area[i];
    int s = t - d;
    s = s < 0 ? 0 : s;
    const int st = __float2int_rn((float)s * (float)(125.0 / 255.0));
    const int o = d + st;
    out[i] = (uint8_t)(o > 255 ? 255 : o);
}

int decide_init(sindyn_base *ctx, DecideStage *d, int W, int H)
{
    d->W = W; d->H = H;
    const size_t N = (size_t)W * H;
    DecideControl *c = nullptr;
    SD_CHECK(ctx->dalloc(&c, 1));
    d->ctl = c;
    uint8_t **up[] = {&d->low0, &d->low, &d->filled, &d->dyna, &d->tmp, &d->out};
    for (uint8_t **p : up) SD_CHECK(ctx->dalloc(p, N));
    SD_CHECK(ctx->dalloc(&d->seedflag, 2 * N));
    return SINDYN_OK;
}

int decide_run(sindyn_base *ctx, DecideStage *d, uint8_t *cls, int *labelsL, RegionStats *stats, int *top, const uint8_t *low_in,
               const uint8_t *high, const uint8_t *high_last, const uint8_t *total_area, const uint8_t *labels)
{
    const int W = d->W, H = d->H, N = W * H;
    DecideControl *ctl = (DecideControl *)d->ctl;
    LAUNCH(ctx, k_dd_reset, 1, 256, 0, ctl);
    LAUNCH(ctx, k_dd_low_counts, SINDYN_NUM_SMS_B200 * 2, 256, 0, low_in, high, high_last, total_area, labels, N, d->low0, ctl);
    SD_CHECK(morph_run(ctx, d->low0, d->low, d->tmp, W, H, 5, MORPH_DILATE));
    // the two keyed planes of the flood-fill surrogate are planes DD_MAXL, DD_MAXL + 1 of the caller's scratch: all labelling in
    // one launch chain (the keyed pass alone costs as much as the 128-plane one: it is latency, not work)
    uint8_t *key = cls + (size_t)DD_MAXL * N;
    int *keyL = labelsL + (size_t)DD_MAXL * (N + 1);
    LAUNCH(ctx, k_dd_planes, cdiv(N, 256), 256, 0, labels, high, d->low, N, ctl, cls, key);
    SD_CHECK(ccl_run_mixed(ctx, cls, labelsL, W, H, DD_MAXL, 2, CCL_REGION, &ctl->nmax));
    SD_CHECK(ccl_top_image(ctx, labelsL, top, W, H, DD_MAXL, &ctl->nmax, stats));
    SD_CHECK(ccl_quad_stats_ccomp(ctx, cls, labelsL, stats, W, H, DD_MAXL, &ctl->nmax));
    CU_CHECK(ctx, cudaMemsetAsync(d->seedflag, 0, 2 * (size_t)N, ctx->stream));
    LAUNCH(ctx, k_dd_seeds, dim3(96, 1, DD_MAXL), 256, 0, cls, labelsL, stats, d->low, labels, keyL, W, H, ctl, d->seedflag);
    LAUNCH(ctx, k_dd_filled, SINDYN_NUM_SMS_B200 * 2, 256, 0, labels, d->low, keyL, d->seedflag, N, ctl, d->filled);
    LAUNCH(ctx, k_dd_dyna, cdiv(N, 256), 256, 0, labels, d->filled, N, ctl, d->dyna);
    SD_CHECK(morph_run(ctx, d->dyna, d->filled, d->tmp, W, H, 9, MORPH_DILATE));
    LAUNCH(ctx, k_dd_final, cdiv(N, 256), 256, 0, d->filled, total_area, N, d->out);
    LAUNCH_CHECK(ctx);
    return SINDYN_OK;
}
