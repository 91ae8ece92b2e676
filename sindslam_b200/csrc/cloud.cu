// cloud.cu -- per-keyframe point-cloud generation of the dense-map consumer:
// octomap_pub/src/pubPointCloud.cc generatePointCloud, single-frame overload (:392-470) and the cross-frame consistency
// overload (:471-678).  The consumer rejects pixels of the dynamic mask (>= 240), back-projects the depth image, votes per
// K-means cluster whether the cluster is occluded / moving against the previous key frame, paints rejected clusters into the
// mask and hands the transformed cloud to octomap (the octree itself is third-party and stays on the host).
//
//   k_cloud_single    one thread per 3rd pixel: NaN / back-projection / colour / world transform
//   k_cloud_vote      one thread per 2nd pixel: re-projection into the previous key frame, occlusion vote, local point,
//                     per-block per-cluster sample counts (for the ordered compaction)
//   k_cloud_label_hist  countNonZero(label == i), i < 12, over the full image
//   k_cloud_plan      one CTA: keep / reject decision per cluster, output offsets of every (block, cluster)
//   k_cloud_scatter   ordered compaction: cluster 0, then every kept cluster, raster order inside a cluster (the order of
//                     the reference's `*tmp += *tmpCluster[i]`), world transform
//   k_cloud_mask_new  imgDynaMaskNew: rejected clusters set to 255
// Arithmetic follows the reference's mixed float / double expressions operation by operation (this file is compiled
// without FMA contraction); pcl::transformPointCloud is restated as float(R p + t) in double, NaN points copied.
#include "ctx.cuh"

#define H_CHECK(h)                       \
    if (!(h)) return SINDYN_ERR_INVALID; \
    cudaSetDevice((h)->device)

#define CLOUD_K 12
#define CLOUD_NT 256

struct CloudIntr { double fx, fy, cx, cy, depth_scale, inv_scale; float fxf, fyf, cxf, cyf, inv_fx_f, inv_fy_f; };
struct CloudPose { double m[16]; };   // row-major 4 x 4

struct CloudStage {
    uint8_t *bgr = nullptr, *mask = nullptr, *mask_last = nullptr, *label = nullptr, *mask_new = nullptr;
    uint16_t *depth = nullptr, *depth_last = nullptr, *depth_new = nullptr;
    sindyn_point *tmp = nullptr, *out = nullptr;
    uint8_t *lab_tmp = nullptr;
    int *block_cnt = nullptr, *block_off = nullptr;   // [n_blocks][12]
    int *ctl = nullptr;                               // occlusion[12] | label_count[12] | kept[12] | n_out
    int n_blocks = 0;
};

static __device__ __forceinline__ void cloud_transform(const CloudPose &T, float &x, float &y, float &z)
{
    if (!(isfinite(x) && isfinite(y) && isfinite(z))) return;
    const double px = x, py = y, pz = z;
    const double ox = ((T.m[0] * px + T.m[1] * py) + T.m[2] * pz) + T.m[3];
    const double oy = ((T.m[4] * px + T.m[5] * py) + T.m[6] * pz) + T.m[7];
    const double oz = ((T.m[8] * px + T.m[9] * py) + T.m[10] * pz) + T.m[11];
    x = (float)ox; y = (float)oy; z = (float)oz;
}

__global__ void k_cloud_single(const uint8_t *__restrict__ bgr, const uint16_t *__restrict__ depth, const uint8_t *__restrict__ mask, int W, int H,
                               int nw, int nh, CloudIntr K, CloudPose Twc, sindyn_point *__restrict__ out)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nw * nh) return;
    const int m = 3 * (s / nw), n = 3 * (s - (s / nw) * nw);
    const int i = m * W + n;
    const float d = (float)((double)depth[i] * K.inv_scale);
    sindyn_point p;
    if ((int)mask[i] >= 240 || (double)d < 0.01 || (double)d > 10) {
        p.x = p.y = p.z = __int_as_float(0x7fc00000);
    } else {
        p.z = d;
        p.x = ((float)n - K.cxf) * p.z * K.inv_fx_f;
        p.y = ((float)m - K.cyf) * p.z * K.inv_fy_f;
    }
    p.b = bgr[3 * i]; p.g = bgr[3 * i + 1]; p.r = bgr[3 * i + 2]; p.a = 0;
    cloud_transform(Twc, p.x, p.y, p.z);
    out[s] = p;
}

__global__ void __launch_bounds__(CLOUD_NT) k_cloud_vote(const uint8_t *__restrict__ bgr, const uint16_t *__restrict__ depth,
                                                         const uint16_t *__restrict__ depth_last, const uint8_t *__restrict__ mask,
                                                         const uint8_t *__restrict__ mask_last, const uint8_t *__restrict__ label, int W, int H, int nw,
                                                         int nh, CloudIntr K, CloudPose T, sindyn_point *__restrict__ tmp,
                                                         uint8_t *__restrict__ lab_tmp, uint16_t *__restrict__ depth_new, int *__restrict__ block_cnt,
                                                         int *__restrict__ ctl)
{
    __shared__ int s_occ[CLOUD_K], s_cnt[CLOUD_K];
    if (threadIdx.x < CLOUD_K) { s_occ[threadIdx.x] = 0; s_cnt[threadIdx.x] = 0; }
    __syncthreads();
    const int s = blockIdx.x * CLOUD_NT + threadIdx.x;
    if (s < nw * nh) {
        const int m = 2 * (s / nw), n = 2 * (s - (s / nw) * nw);
        const int i = m * W + n;
        const float dcur = (float)((double)depth[i] * K.inv_scale);
        const int lab = label[i];
        uint8_t lt = 255;
        if (lab < CLOUD_K) {
            lt = (uint8_t)lab;
            const double px = (double)(((float)n - K.cxf) * dcur / K.fxf), py = (double)(((float)m - K.cyf) * dcur / K.fyf), pz = (double)dcur;
            const double q0 = ((T.m[0] * px + T.m[1] * py) + T.m[2] * pz) + T.m[3];
            const double q1 = ((T.m[4] * px + T.m[5] * py) + T.m[6] * pz) + T.m[7];
            const double q2 = ((T.m[8] * px + T.m[9] * py) + T.m[10] * pz) + T.m[11];
            const double u = K.fx * q0 + K.cx * q2, v = K.fy * q1 + K.cy * q2;
            const float xt = (float)(u / q2), yt = (float)(v / q2);
            float dlast = 0.0f;
            bool dyn_last = false;
            if (yt >= 0.0f && yt < (float)H && xt >= 0.0f && xt < (float)W) {
                const int j = (int)yt * W + (int)xt;
                dlast = (float)((double)depth_last[j] * K.inv_scale);
                depth_new[i] = (uint16_t)((double)dlast * K.depth_scale);
                dyn_last = mask_last[j] > 240;
            }
            if (dcur >= 0.0f && dcur < 10.0f && dlast >= 0.0f && dlast < 10.0f) {
                const float diff = dcur - dlast;
                const double lhs = (double)(diff * diff), r = 0.13 * (double)dcur;
                if (lhs > r * r || dyn_last) atomicAdd(&s_occ[lab], 1);
            }
            atomicAdd(&s_cnt[lab], 1);
            sindyn_point p;
            if ((int)mask[i] >= 240 || (double)dcur < 0.01 || (double)dcur > 10) {
                p.x = p.y = p.z = __int_as_float(0x7fc00000);
            } else {
                p.z = dcur;
                p.x = (float)(((double)n - K.cx) * (double)p.z / K.fx);
                p.y = (float)(((double)m - K.cy) * (double)p.z / K.fy);
            }
            p.b = bgr[3 * i]; p.g = bgr[3 * i + 1]; p.r = bgr[3 * i + 2]; p.a = 0;
            tmp[s] = p;
        }
        lab_tmp[s] = lt;
    }
    __syncthreads();
    if (threadIdx.x < CLOUD_K) {
        block_cnt[blockIdx.x * CLOUD_K + threadIdx.x] = s_cnt[threadIdx.x];
        if (s_occ[threadIdx.x]) atomicAdd(&ctl[threadIdx.x], s_occ[threadIdx.x]);
    }
}

__global__ void k_cloud_label_hist(const uint8_t *__restrict__ label, int n, int *__restrict__ ctl)
{
    __shared__ int s_h[CLOUD_K];
    if (threadIdx.x < CLOUD_K) s_h[threadIdx.x] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int l = label[i];
        if (l < CLOUD_K) atomicAdd(&s_h[l], 1);
    }
    __syncthreads();
    if (threadIdx.x < CLOUD_K && s_h[threadIdx.x]) atomicAdd(&ctl[CLOUD_K + threadIdx.x], s_h[threadIdx.x]);
}

// cluster 0 is always kept; cluster i is kept when occlusion_i * 9 <= 0.4 * |label == i| (pubPointCloud.cc:643-657)
__global__ void k_cloud_plan(const int *__restrict__ block_cnt, int n_blocks, int *__restrict__ block_off, int *__restrict__ ctl)
{
    __shared__ int s_tot[CLOUD_K], s_base[CLOUD_K], s_keep[CLOUD_K];
    const int l = threadIdx.x;
    if (l < CLOUD_K) {
        int t = 0;
        for (int b = 0; b < n_blocks; ++b) t += block_cnt[b * CLOUD_K + l];
        s_tot[l] = t;
        s_keep[l] = (l == 0 || (double)ctl[l] * 9 <= 0.4 * (double)ctl[CLOUD_K + l]) ? 1 : 0;
        ctl[2 * CLOUD_K + l] = s_keep[l];
    }
    __syncthreads();
    if (l == 0) {
        int acc = 0;
        for (int k = 0; k < CLOUD_K; ++k) { s_base[k] = acc; if (s_keep[k]) acc += s_tot[k]; }
        ctl[3 * CLOUD_K] = acc;
    }
    __syncthreads();
    if (l < CLOUD_K) {
        int acc = s_base[l];
        for (int b = 0; b < n_blocks; ++b) { block_off[b * CLOUD_K + l] = acc; acc += block_cnt[b * CLOUD_K + l]; }
    }
}

__global__ void __launch_bounds__(CLOUD_NT) k_cloud_scatter(const sindyn_point *__restrict__ tmp, const uint8_t *__restrict__ lab_tmp, int ns,
                                                            const int *__restrict__ block_off, const int *__restrict__ ctl, CloudPose Twc,
                                                            sindyn_point *__restrict__ out)
{
    __shared__ int s_w[CLOUD_NT / 32][CLOUD_K];
    const int s = blockIdx.x * CLOUD_NT + threadIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int lab = s < ns ? lab_tmp[s] : 255;
    int rank = 0;
    for (int l = 0; l < CLOUD_K; ++l) {
        const unsigned b = __ballot_sync(0xffffffffu, lab == l);
        if (lane == 0) s_w[wid][l] = __popc(b);
        if (lab == l) rank = __popc(b & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (lab >= CLOUD_K || !ctl[2 * CLOUD_K + lab]) return;
    for (int w = 0; w < wid; ++w) rank += s_w[w][lab];
    sindyn_point p = tmp[s];
    cloud_transform(Twc, p.x, p.y, p.z);
    out[block_off[blockIdx.x * CLOUD_K + lab] + rank] = p;
}

__global__ void k_cloud_mask_new(const uint8_t *__restrict__ mask, const uint8_t *__restrict__ label, int n, const int *__restrict__ ctl,
                                 uint8_t *__restrict__ mask_new)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = label[i];
    mask_new[i] = (l < CLOUD_K && !ctl[2 * CLOUD_K + l]) ? 255 : mask[i];
}

static int cloud_stage(sindyn_ctx *h, CloudStage **out)
{
    if (!h->cloud) {
        CloudStage *c = new CloudStage();
        const size_t N = h->N, ns = (size_t)((h->W + 1) / 2) * ((h->H + 1) / 2);
        c->n_blocks = (int)((ns + CLOUD_NT - 1) / CLOUD_NT);
        SD_CHECK(h->dalloc(&c->bgr, N * 3));
        SD_CHECK(h->dalloc(&c->mask, N)); SD_CHECK(h->dalloc(&c->mask_last, N)); SD_CHECK(h->dalloc(&c->label, N)); SD_CHECK(h->dalloc(&c->mask_new, N));
        SD_CHECK(h->dalloc(&c->depth, N)); SD_CHECK(h->dalloc(&c->depth_last, N)); SD_CHECK(h->dalloc(&c->depth_new, N));
        SD_CHECK(h->dalloc(&c->tmp, ns)); SD_CHECK(h->dalloc(&c->out, ns)); SD_CHECK(h->dalloc(&c->lab_tmp, ns));
        SD_CHECK(h->dalloc(&c->block_cnt, (size_t)c->n_blocks * CLOUD_K)); SD_CHECK(h->dalloc(&c->block_off, (size_t)c->n_blocks * CLOUD_K));
        SD_CHECK(h->dalloc(&c->ctl, 4 * CLOUD_K));
        h->cloud = c;
    }
    *out = h->cloud;
    return SINDYN_OK;
}

void cloud_stage_destroy(sindyn_ctx *h)
{
    delete h->cloud;   // device buffers belong to the handle's allocation list
    h->cloud = nullptr;
}

static CloudIntr cloud_intr(const sindyn_ctx *h, const double *intr)
{
    CloudIntr K;
    K.fx = intr ? intr[0] : (double)h->cfg.fx; K.fy = intr ? intr[1] : (double)h->cfg.fy;
    K.cx = intr ? intr[2] : (double)h->cfg.cx; K.cy = intr ? intr[3] : (double)h->cfg.cy;
    K.depth_scale = intr ? intr[4] : (double)h->cfg.depth_scale;
    K.inv_scale = 1.0 / K.depth_scale;
    K.fxf = (float)K.fx; K.fyf = (float)K.fy; K.cxf = (float)K.cx; K.cyf = (float)K.cy;
    K.inv_fx_f = (float)(1.0 / K.fx); K.inv_fy_f = (float)(1.0 / K.fy);
    return K;
}

extern "C" int sindyn_cloud_single(sindyn_handle h, const uint8_t *bgr, size_t bgr_step, const uint16_t *depth, size_t depth_step,
                                   const uint8_t *mask, size_t mask_step, const double *Twc16, const double *intr5, sindyn_point *points_out,
                                   int *n_out)
{
    H_CHECK(h);
    if (!bgr || !depth || !mask || !Twc16 || !points_out) return SINDYN_ERR_INVALID;
    CloudStage *c;
    SD_CHECK(cloud_stage(h, &c));
    const int W = h->W, H = h->H, nw = (W + 2) / 3, nh = (H + 2) / 3;
    CU_CHECK(h, copy_in_2d(c->bgr, bgr, bgr_step, (size_t)W * 3, H, h->stream));
    CU_CHECK(h, copy_in_2d(c->depth, depth, depth_step, (size_t)W * 2, H, h->stream));
    CU_CHECK(h, copy_in_2d(c->mask, mask, mask_step, W, H, h->stream));
    CloudPose T;
    for (int i = 0; i < 16; ++i) T.m[i] = Twc16[i];
    LAUNCH(h, k_cloud_single, cdiv(nw * nh, 256), 256, 0, c->bgr, c->depth, c->mask, W, H, nw, nh, cloud_intr(h, intr5), T, c->out);
    LAUNCH_CHECK(h);
    CU_CHECK(h, cudaMemcpyAsync(points_out, c->out, sizeof(sindyn_point) * nw * nh, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    if (n_out) *n_out = nw * nh;
    return SINDYN_OK;
}

extern "C" int sindyn_cloud_consistent(sindyn_handle h, const uint8_t *bgr, size_t bgr_step, const uint16_t *depth, size_t depth_step,
                                       const uint16_t *depth_last, size_t depth_last_step, const uint8_t *mask, size_t mask_step,
                                       const uint8_t *mask_last, size_t mask_last_step, const uint8_t *label, size_t label_step,
                                       const double *T_rel16, const double *Twc16, const double *intr5, sindyn_point *points_out, int *n_out,
                                       uint8_t *mask_new_out, size_t mask_new_step, int *stats36_out, uint16_t *depth_new_out)
{
    H_CHECK(h);
    if (!bgr || !depth || !depth_last || !mask || !mask_last || !label || !T_rel16 || !Twc16 || !points_out || !n_out) return SINDYN_ERR_INVALID;
    CloudStage *c;
    SD_CHECK(cloud_stage(h, &c));
    const int W = h->W, H = h->H, nw = (W + 1) / 2, nh = (H + 1) / 2, ns = nw * nh;
    CU_CHECK(h, copy_in_2d(c->bgr, bgr, bgr_step, (size_t)W * 3, H, h->stream));
    CU_CHECK(h, copy_in_2d(c->depth, depth, depth_step, (size_t)W * 2, H, h->stream));
    CU_CHECK(h, copy_in_2d(c->depth_last, depth_last, depth_last_step, (size_t)W * 2, H, h->stream));
    CU_CHECK(h, copy_in_2d(c->mask, mask, mask_step, W, H, h->stream));
    CU_CHECK(h, copy_in_2d(c->mask_last, mask_last, mask_last_step, W, H, h->stream));
    CU_CHECK(h, copy_in_2d(c->label, label, label_step, W, H, h->stream));
    CU_CHECK(h, cudaMemsetAsync(c->ctl, 0, sizeof(int) * 4 * CLOUD_K, h->stream));
    CU_CHECK(h, cudaMemsetAsync(c->depth_new, 0, sizeof(uint16_t) * h->N, h->stream));
    CloudPose Tr, Tw;
    for (int i = 0; i < 16; ++i) { Tr.m[i] = T_rel16[i]; Tw.m[i] = Twc16[i]; }
    const CloudIntr K = cloud_intr(h, intr5);
    LAUNCH(h, k_cloud_vote, c->n_blocks, CLOUD_NT, 0, c->bgr, c->depth, c->depth_last, c->mask, c->mask_last, c->label, W, H, nw, nh, K, Tr, c->tmp,
           c->lab_tmp, c->depth_new, c->block_cnt, c->ctl);
    LAUNCH(h, k_cloud_label_hist, SINDYN_NUM_SMS_B200, 256, 0, c->label, h->N, c->ctl);
    LAUNCH(h, k_cloud_plan, 1, 32, 0, c->block_cnt, c->n_blocks, c->block_off, c->ctl);
    LAUNCH(h, k_cloud_scatter, c->n_blocks, CLOUD_NT, 0, c->tmp, c->lab_tmp, ns, c->block_off, c->ctl, Tw, c->out);
    LAUNCH(h, k_cloud_mask_new, cdiv(h->N, 256), 256, 0, c->mask, c->label, h->N, c->ctl, c->mask_new);
    LAUNCH_CHECK(h);
    int ctl_host[4 * CLOUD_K];
    CU_CHECK(h, cudaMemcpyAsync(ctl_host, c->ctl, sizeof(ctl_host), cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    const int n = ctl_host[3 * CLOUD_K];
    *n_out = n;
    if (n > 0) CU_CHECK(h, cudaMemcpyAsync(points_out, c->out, sizeof(sindyn_point) * n, cudaMemcpyDeviceToHost, h->stream));
    if (mask_new_out) CU_CHECK(h, copy_out_2d(mask_new_out, mask_new_step, c->mask_new, W, H, h->stream));
    if (depth_new_out) CU_CHECK(h, cudaMemcpyAsync(depth_new_out, c->depth_new, sizeof(uint16_t) * h->N, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    if (stats36_out)
        for (int i = 0; i < 3 * CLOUD_K; ++i) stats36_out[i] = ctl_host[i];
    return SINDYN_OK;
}
