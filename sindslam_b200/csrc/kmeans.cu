// kmeans.cu -- pyramid K-means on back-projected depth + cluster ordering.
//
// Replaces DynaDetect::SegByKmeans (ORB_SLAM2/src/DynaDetect.cc:315-420: 4-level u16 depth pyramid,
// back-projection to (X, Y, 1.5 Z), label propagation by bilinear resize of the label image,
// cv::kmeans(K=12, TermCriteria(EPS+COUNT, 4, 0.07), 1 attempt, KMEANS_USE_INITIAL_LABELS)) and the
// cluster ordering of DynaDetect.cc:1425-1491.
//
// cv::kmeans semantics follow SURVEY.md Appendix C.1 (verified against cv2 4.13) incl. the
// empty-cluster repair.  ONE documented deviation: OpenCV sums the centre coordinates sequentially
// in float32, which no parallel reduction reproduces; here the sums are accumulated in 2^-36
// fixed point (int64 atomics: exact, order independent, run-to-run deterministic).  The oracle's
// kmeans_fx does the same, the CUDA path is bit-exact against it, and the agreement with
// cv2.kmeans itself is reported by the tests (a few labels out of 307 200).
//
// No host synchronisation: the EPS stop criterion is a device-side flag that later launches test.
#include "kmeans.cuh"

#define KM_FIX_SCALE 68719476736.0 /* 2^36 */

// cv::resize(u16, 0.5x, INTER_LINEAR) takes the INTER_AREA fast path: rint(mean of 2x2), ties to even
__global__ void k_depth_half(const uint16_t *__restrict__ src, int sw, uint16_t *__restrict__ dst, int dw, int dh)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const uint16_t *r0 = src + (size_t)(2 * y) * sw + 2 * x, *r1 = r0 + sw;
    unsigned s = (unsigned)r0[0] + r0[1] + r1[0] + r1[1];
    unsigned q = s >> 2, r = s & 3u;
    if (r == 3u || (r == 2u && (q & 1u))) q++;
    dst[y * dw + x] = (uint16_t)q;
}

struct KmIntr { float fx, fy, cx, cy, depth_scale, depth_weight; };

// DynaDetect.cc:347-369 in float32, evaluated left to right without contraction (file built with -fmad=false)
__device__ __forceinline__ void km_point(const uint16_t *__restrict__ depth, int i, int w, float s, const KmIntr &K, float &X, float &Y, float &Z)
{
    const int row = i / w, col = i - row * w;
    const float dfull = (float)depth[i] * s;
    const uint16_t d = (uint16_t)dfull;  // ushort depth = pyr(row,col) * scales[level]
    X = 0.f; Y = 0.f; Z = 0.f;
    const float df = (float)d;
    if (!(df / K.depth_scale >= 6.0f || d == 0)) {
        const float depth2 = df * (1.0f / K.depth_scale);
        Z = depth2 * K.depth_weight;
        X = (((float)col - K.cx * s) * depth2) * (1.0f / (K.fx * s));
        Y = (((float)row - K.cy * s) * depth2) * (1.0f / (K.fy * s));
    }
}

__global__ void k_km_points(const uint16_t *__restrict__ depth, int w, int h, float s, KmIntr K, float *__restrict__ pts)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w * h) return;
    float X, Y, Z;
    km_point(depth, i, w, s, K, X, Y, Z);
    pts[3 * (size_t)i] = X;
    pts[3 * (size_t)i + 1] = Y;
    pts[3 * (size_t)i + 2] = Z;
}

__global__ void k_count_nonzero_u8(const uint8_t *__restrict__ img, int n, int *__restrict__ out)
{
    int c = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) c += img[i] != 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// coarsest-level initial labels: 3x4 grid when countNonZero(imgLabelLast) == 0 (DynaDetect.cc:374-386),
// else imgLabelLast -> float -> cv::resize INTER_LINEAR -> int (DynaDetect.cc:390-394)
__global__ void k_km_init_labels(const uint8_t *__restrict__ label_last, int sw, int sh, const int *__restrict__ nz,
                                 int *__restrict__ labels, int w, int h, int nrow, int ncol, float fx, float fy)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    if (*nz == 0) {
        float br = (float)h / (float)nrow, bc = (float)w / (float)ncol;
        labels[y * w + x] = (int)floorf((float)y / br) * ncol + (int)floorf((float)x / bc);
        return;
    }
    float sy = ((float)y + 0.5f) * fy - 0.5f, sx = ((float)x + 0.5f) * fx - 0.5f;
    int y0 = (int)floorf(sy), x0 = (int)floorf(sx);
    float ty = sy - (float)y0, tx = sx - (float)x0;
    if (y0 < 0) { y0 = 0; ty = 0.f; }
    if (y0 >= sh - 1) { y0 = sh - 1; ty = 0.f; }
    if (x0 < 0) { x0 = 0; tx = 0.f; }
    if (x0 >= sw - 1) { x0 = sw - 1; tx = 0.f; }
    int x1 = min(x0 + 1, sw - 1), y1 = min(y0 + 1, sh - 1);
    float a = (float)label_last[y0 * sw + x0], b = (float)label_last[y0 * sw + x1];
    float c = (float)label_last[y1 * sw + x0], d = (float)label_last[y1 * sw + x1];
    float top = a * (1.f - tx) + b * tx, bot = c * (1.f - tx) + d * tx;
    labels[y * w + x] = __float2int_rn(top * (1.f - ty) + bot * ty);
}

// label image -> float -> cv::resize INTER_LINEAR -> int (round half even)  (DynaDetect.cc:390-394,402-406)
template <typename T>
__global__ void k_km_labels_resize(const T *__restrict__ src, int sw, int sh, int *__restrict__ dst, int dw, int dh, float fx, float fy)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    float sy = ((float)y + 0.5f) * fy - 0.5f, sx = ((float)x + 0.5f) * fx - 0.5f;
    int y0 = (int)floorf(sy), x0 = (int)floorf(sx);
    float ty = sy - (float)y0, tx = sx - (float)x0;
    if (y0 < 0) { y0 = 0; ty = 0.f; }
    if (y0 >= sh - 1) { y0 = sh - 1; ty = 0.f; }
    if (x0 < 0) { x0 = 0; tx = 0.f; }
    if (x0 >= sw - 1) { x0 = sw - 1; tx = 0.f; }
    int x1 = min(x0 + 1, sw - 1), y1 = min(y0 + 1, sh - 1);
    float a = (float)src[y0 * sw + x0], b = (float)src[y0 * sw + x1], c = (float)src[y1 * sw + x0], d = (float)src[y1 * sw + x1];
    float top = a * (1.f - tx) + b * tx, bot = c * (1.f - tx) + d * tx;
    dst[y * dw + x] = __float2int_rn(top * (1.f - ty) + bot * ty);
}

// Centre sums (2^-36 fixed point, exact and order independent).  A warp walks spans of KM_SPAN consecutive pixels (lane + 32 j):
// coalesced, and the pixels one lane sees in a row lie 32 px apart on the same image rows, i.e. almost always in the same
// cluster -- so every lane keeps a RUN (label, three 64-bit sums, count) in registers and touches shared memory only
// when the label changes (about once per span instead of four atomics per pixel); block totals go to the level's state
// with one global atomic per (block, cluster, component).  The back-projected point is recomputed from the u16 depth
// (2 B per pixel instead of the 12 B of the point plane; same float expressions as k_km_points).
struct KmRun { long long sx, sy, sz; int cnt, lab; };
__device__ __forceinline__ void km_run_flush(KmRun &r, unsigned long long *s_sum, int *s_cnt)
{
    if (r.cnt) {
        atomicAdd(&s_sum[r.lab * 3 + 0], (unsigned long long)r.sx);
        atomicAdd(&s_sum[r.lab * 3 + 1], (unsigned long long)r.sy);
        atomicAdd(&s_sum[r.lab * 3 + 2], (unsigned long long)r.sz);
        atomicAdd(&s_cnt[r.lab], r.cnt);
    }
    r.sx = r.sy = r.sz = 0; r.cnt = 0;
}
__device__ __forceinline__ void km_run_add(KmRun &r, int lab, float x, float y, float z, unsigned long long *s_sum, int *s_cnt)
{
    if (lab != r.lab) { km_run_flush(r, s_sum, s_cnt); r.lab = lab; }
    r.sx += __double2ll_rn((double)x * KM_FIX_SCALE);
    r.sy += __double2ll_rn((double)y * KM_FIX_SCALE);
    r.sz += __double2ll_rn((double)z * KM_FIX_SCALE);
    r.cnt++;
}
#define KM_SPAN 128

// first centre pass of a level: sums of the initial labels
__global__ void k_km_accum(const uint16_t *__restrict__ depth, int w, float s, KmIntr K, const int *__restrict__ labels, int n, KmState *st)
{
    __shared__ unsigned long long s_sum[KM_K * 3];
    __shared__ int s_cnt[KM_K];
    for (int j = threadIdx.x; j < KM_K * 3; j += blockDim.x) s_sum[j] = 0ull;
    for (int j = threadIdx.x; j < KM_K; j += blockDim.x) s_cnt[j] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    KmRun r; r.sx = r.sy = r.sz = 0; r.cnt = 0; r.lab = 0;
    for (int base = gw * KM_SPAN; base < n; base += nw * KM_SPAN)
#pragma unroll 4
        for (int j = 0; j < KM_SPAN / 32; ++j) {
            const int i = base + j * 32 + lane;
            if (i >= n) break;
            const int lab = labels[i];
            if ((unsigned)lab >= (unsigned)KM_K) continue;
            float x, y, z;
            km_point(depth, i, w, s, K, x, y, z);
            km_run_add(r, lab, x, y, z, s_sum, s_cnt);
        }
    km_run_flush(r, s_sum, s_cnt);
    __syncthreads();
    for (int j = threadIdx.x; j < KM_K * 3; j += blockDim.x)
        if (s_sum[j]) atomicAdd(&st->sums[0][j], s_sum[j]);
    for (int j = threadIdx.x; j < KM_K; j += blockDim.x)
        if (s_cnt[j]) atomicAdd(&st->counts[0][j], s_cnt[j]);
}

// assignment (argmin of float32 squared distance, first minimum wins) fused with the next centre pass
__global__ void k_km_assign_accum(const uint16_t *__restrict__ depth, int w, float s, KmIntr K, int *__restrict__ labels, int n, KmState *st, int it)
{
    if (st->done) return;
    __shared__ unsigned long long s_sum[KM_K * 3];
    __shared__ int s_cnt[KM_K];
    __shared__ float s_c[KM_K * 3];
    for (int j = threadIdx.x; j < KM_K * 3; j += blockDim.x) { s_sum[j] = 0ull; s_c[j] = st->centers[j]; }
    for (int j = threadIdx.x; j < KM_K; j += blockDim.x) s_cnt[j] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    KmRun r; r.sx = r.sy = r.sz = 0; r.cnt = 0; r.lab = 0;
    for (int base = gw * KM_SPAN; base < n; base += nw * KM_SPAN)
#pragma unroll 2
        for (int j = 0; j < KM_SPAN / 32; ++j) {
            const int i = base + j * 32 + lane;
            if (i >= n) break;
            float x, y, z;
            km_point(depth, i, w, s, K, x, y, z);
            float best = 0.f;
            int lab = 0;
#pragma unroll
            for (int k = 0; k < KM_K; ++k) {
                float d0 = x - s_c[3 * k], d1 = y - s_c[3 * k + 1], d2 = z - s_c[3 * k + 2];
                float dist = d0 * d0;
                dist = dist + d1 * d1;
                dist = dist + d2 * d2;
                if (k == 0 || dist < best) { best = dist; lab = k; }
            }
            labels[i] = lab;
            km_run_add(r, lab, x, y, z, s_sum, s_cnt);
        }
    km_run_flush(r, s_sum, s_cnt);
    __syncthreads();
    const int b = (it + 1) & 1;
    for (int j = threadIdx.x; j < KM_K * 3; j += blockDim.x)
        if (s_sum[j]) atomicAdd(&st->sums[b][j], s_sum[j]);
    for (int j = threadIdx.x; j < KM_K; j += blockDim.x)
        if (s_cnt[j]) atomicAdd(&st->counts[b][j], s_cnt[j]);
}

// single CTA: empty-cluster repair, centres = sum * (1/count), max centre shift, stop flag
__global__ void __launch_bounds__(1024) k_km_finalize(const float *__restrict__ pts, int *__restrict__ labels, int n, KmState *st,
                                                      int it, int max_iter, double eps2)
{
    if (st->done) return;
    const int b = it & 1;
    __shared__ unsigned long long s_best;
    __shared__ int s_maxk;
    __shared__ float s_base[3];
    long long *sums = (long long *)st->sums[b];
    int *counts = st->counts[b];
    if (threadIdx.x < KM_K * 3) st->old[threadIdx.x] = st->centers[threadIdx.x];
    __syncthreads();
    for (int k = 0; k < KM_K; ++k) {
        if (counts[k] != 0) continue;  // uniform: counts is only written by thread 0 between barriers
        if (threadIdx.x == 0) {
            int max_k = 0;
            for (int k1 = 1; k1 < KM_K; ++k1)
                if (counts[max_k] < counts[k1]) max_k = k1;
            float scale = 1.f / (float)counts[max_k];
            for (int j = 0; j < 3; ++j) {
                float bc = (float)((double)sums[max_k * 3 + j] * (1.0 / KM_FIX_SCALE)) * scale;
                s_base[j] = bc;
                st->old[max_k * 3 + j] = bc;  // cv::kmeans overwrites old_centers[max_k] with the normalised centre
            }
            s_maxk = max_k;
            s_best = 0ull;
        }
        __syncthreads();
        const int max_k = s_maxk;
        unsigned long long best = 0ull;
        bool any = false;
        // one CTA walks all n points: eight label loads of a thread are in flight together (the walk is latency bound)
        for (int i0 = threadIdx.x; i0 < n; i0 += 8 * blockDim.x) {
            int lb[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int i = i0 + u * blockDim.x; lb[u] = i < n ? labels[i] : -1; }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (lb[u] != max_k) continue;
                const int i = i0 + u * blockDim.x;
                float d0 = pts[3 * (size_t)i] - s_base[0], d1 = pts[3 * (size_t)i + 1] - s_base[1], d2 = pts[3 * (size_t)i + 2] - s_base[2];
                float dist = d0 * d0;
                dist = dist + d1 * d1;
                dist = dist + d2 * d2;
                unsigned long long key = ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned)i;  // ties: last index wins
                if (!any || key > best) { best = key; any = true; }
            }
        }
        if (any) atomicMax(&s_best, best);
        __syncthreads();
        if (threadIdx.x == 0) {
            int far = (int)(s_best & 0xffffffffull);
            counts[max_k]--;
            counts[k]++;
            labels[far] = k;
            for (int j = 0; j < 3; ++j) {
                long long f = __double2ll_rn((double)pts[3 * (size_t)far + j] * KM_FIX_SCALE);
                sums[max_k * 3 + j] -= f;
                sums[k * 3 + j] += f;
            }
        }
        __syncthreads();
    }
    // centres, centre shift and the stop decision: one thread per cluster (same operations and order per cluster as the serial loop)
    __shared__ double s_dist[KM_K];
    if (threadIdx.x < KM_K) {
        const int k = threadIdx.x;
        const float scale = 1.f / (float)counts[k];
        double dist = 0.0;
        for (int j = 0; j < 3; ++j) {
            const float c = (float)((double)sums[k * 3 + j] * (1.0 / KM_FIX_SCALE)) * scale;
            st->centers[k * 3 + j] = c;
            const double t = (double)c - (double)st->old[k * 3 + j];
            dist += t * t;
        }
        s_dist[k] = dist;
        st->final_counts[k] = counts[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double shift = 0.0;
        for (int k = 0; k < KM_K; ++k)
            if (it > 0 && s_dist[k] > shift) shift = s_dist[k];
        const bool last = (it + 1 == max_iter) || (it > 0 && shift <= eps2);
        if (last) st->done = 1;
        st->iters = it + 1;
    }
    // zero the accumulators of the next centre pass
    if (threadIdx.x < KM_K * 3) st->sums[b ^ 1][threadIdx.x] = 0ull;
    if (threadIdx.x < KM_K) st->counts[b ^ 1][threadIdx.x] = 0;
}

__global__ void k_i32_to_u8(const int *__restrict__ src, uint8_t *__restrict__ dst, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (uint8_t)min(max(src[i], 0), 255);
}

// DynaDetect.cc:1425-1490: depth ordering of the 12 centres, <60 px clusters dropped, first <=6 clusters
// covering <60 % of the image flagged for imgLabelForSegEdge.
__global__ void k_cluster_order(KmState *st, int n_pixels, ClusterOrder *out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float z[KM_K];
    int idx[KM_K];
    for (int k = 0; k < KM_K; ++k) {
        float v = st->centers[k * 3 + 2];
        if (v < 0.2f) v += 20.0f;
        z[k] = v;
        idx[k] = k;
    }
    for (int i = 1; i < KM_K; ++i) {  // stable insertion sort, ascending
        int id = idx[i];
        int j = i - 1;
        while (j >= 0 && z[idx[j]] > z[id]) { idx[j + 1] = idx[j]; --j; }
        idx[j + 1] = id;
    }
    float ratioArea = 0.f;
    const float total = (float)n_pixels;
    int count0 = 0, nk = 0;
    for (int k = 0; k < KM_K; ++k) { out->seg_edge_lut[k] = 0; out->rank_of[k] = -1; out->kept[k] = -1; }
    for (int i = 0; i < KM_K; ++i) {
        int id = idx[i];
        int cnt = st->final_counts[id];
        if (cnt < 60) continue;
        out->kept[nk] = id;
        out->rank_of[id] = nk;
        out->kept_area[nk] = cnt;
        ++nk;
        float ratio = (float)cnt * (1.0f / total);
        ratioArea += ratio;
        if (count0 <= 5 && ratioArea < 0.6f) { out->seg_edge_lut[id] = 255; ++count0; }
    }
    out->n_kept = nk;
    out->count0 = count0;
}

__global__ void k_seg_edge(const uint8_t *__restrict__ labels, int n, const ClusterOrder *__restrict__ ord, uint8_t *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t l = labels[i];
    out[i] = l < KM_K ? ord->seg_edge_lut[l] : 0;
}

int kmeans_init(sindyn_base *ctx, KmeansStage *k, int W, int H)
{
    k->W = W; k->H = H;
    const float scales[4] = {1.0f, 0.5f, 0.25f, 0.125f};
    for (int l = 0; l < 4; ++l) {
        k->lw[l] = (int)((float)W * scales[l]);
        k->lh[l] = (int)((float)H * scales[l]);
        // the depth pyramid halves the previous level (DynaDetect.cc:333); sizes coincide when W,H are multiples of 8
        if (l > 0) SD_CHECK(ctx->dalloc(&k->depth_pyr[l], (size_t)k->lw[l] * k->lh[l]));
        SD_CHECK(ctx->dalloc(&k->labels[l], (size_t)k->lw[l] * k->lh[l]));
    }
    if (W % 8 || H % 8) { ctx->err = "kmeans: width and height must be multiples of 8"; return SINDYN_ERR_INVALID; }
    SD_CHECK(ctx->dalloc(&k->points, (size_t)W * H * 3));
    SD_CHECK(ctx->dalloc(&k->points_lvl, (size_t)(W / 2) * (H / 2) * 3));
    SD_CHECK(ctx->dalloc(&k->labels_u8, (size_t)W * H));
    SD_CHECK(ctx->dalloc(&k->seg_edge, (size_t)W * H));
    SD_CHECK(ctx->dalloc(&k->state, 4));
    SD_CHECK(ctx->dalloc(&k->order, 1));
    SD_CHECK(ctx->dalloc(&k->nz_flag, 1));
    return SINDYN_OK;
}

int kmeans_run(sindyn_base *ctx, KmeansStage *k, const uint16_t *depth, const uint8_t *label_last, const sindyn_config *cfg)
{
    const dim3 blk(32, 8);
    CU_CHECK(ctx, cudaMemsetAsync(k->nz_flag, 0, sizeof(int), ctx->stream));
    LAUNCH(ctx, k_count_nonzero_u8, SINDYN_NUM_SMS_B200, 256, 0, label_last, k->W * k->H, k->nz_flag);
    k->depth_pyr[0] = (uint16_t *)depth;
    for (int l = 1; l < 4; ++l) {
        dim3 grd(cdiv(k->lw[l], 32), cdiv(k->lh[l], 8));
        LAUNCH(ctx, k_depth_half, grd, blk, 0, k->depth_pyr[l - 1], k->lw[l - 1], k->depth_pyr[l], k->lw[l], k->lh[l]);
    }
    CU_CHECK(ctx, cudaMemsetAsync(k->state, 0, sizeof(KmState) * 4, ctx->stream));
    KmIntr K{cfg->fx, cfg->fy, cfg->cx, cfg->cy, cfg->depth_scale, cfg->depth_weight};
    const float scales[4] = {1.0f, 0.5f, 0.25f, 0.125f};
    for (int l = 3; l >= 0; --l) {
        const int w = k->lw[l], h = k->lh[l], n = w * h;
        float *pts = l == 0 ? k->points : k->points_lvl;
        KmState *st = k->state + l;
        LAUNCH(ctx, k_km_points, cdiv(n, 256), 256, 0, k->depth_pyr[l], w, h, scales[l], K, pts);
        dim3 grd(cdiv(w, 32), cdiv(h, 8));
        if (l == 3) {
            float fx = (float)(1.0 / ((double)w / (double)k->W)), fy = (float)(1.0 / ((double)h / (double)k->H));
            LAUNCH(ctx, k_km_init_labels, grd, blk, 0, label_last, k->W, k->H, k->nz_flag, k->labels[l], w, h, cfg->n_row_cluster,
                   cfg->n_col_cluster, fx, fy);
        } else {
            float fx = (float)(1.0 / ((double)w / (double)k->lw[l + 1])), fy = (float)(1.0 / ((double)h / (double)k->lh[l + 1]));
            LAUNCH(ctx, k_km_labels_resize<int>, grd, blk, 0, k->labels[l + 1], k->lw[l + 1], k->lh[l + 1], k->labels[l], w, h, fx, fy);
        }
        int nblk = min(cdiv(n, KM_SPAN * 8), SINDYN_NUM_SMS_B200 * 8);     // 8 warps per block, one span per warp and trip
        LAUNCH(ctx, k_km_accum, nblk, 256, 0, k->depth_pyr[l], w, scales[l], K, k->labels[l], n, st);
        for (int it = 0; it < 4; ++it) {
            LAUNCH(ctx, k_km_finalize, 1, 1024, 0, pts, k->labels[l], n, st, it, 4, 0.07 * 0.07);
            if (it < 3) LAUNCH(ctx, k_km_assign_accum, nblk, 256, 0, k->depth_pyr[l], w, scales[l], K, k->labels[l], n, st, it);
        }
    }
    const int N = k->W * k->H;
    LAUNCH(ctx, k_i32_to_u8, cdiv(N, 256), 256, 0, k->labels[0], k->labels_u8, N);
    LAUNCH(ctx, k_cluster_order, 1, 32, 0, k->state, N, k->order);
    LAUNCH(ctx, k_seg_edge, cdiv(N, 256), 256, 0, k->labels_u8, N, k->order, k->seg_edge);
    LAUNCH_CHECK(ctx);
    return SINDYN_OK;
}
