// flow.cu -- the flow branch of DynaDetect (ORB_SLAM2/src/DynaDetect.cc:1023-1147,1163-1235) and its C-ABI
// entry points: Brox(cur, lastlast) -> negate -> large-motion test -> optional Brox(cur, last) -> variational
// refinement -> up-sample x 1/0.6 -> sample weighting -> homography.
#include "ctx.cuh"

#define H_CHECK(h)                       \
    if (!(h)) return SINDYN_ERR_INVALID; \
    cudaSetDevice((h)->device)

// |flow| + global max (cartToPolar magnitude: sqrt(fma(x, x, y*y)), see residual.cu)
__global__ void k_flow_mag(const float2 *__restrict__ flow, int n, float *__restrict__ mag, unsigned int *__restrict__ gmax)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.f;
    if (i < n) {
        float2 f = flow[i];
        m = __fsqrt_rn(__fmaf_rn(f.x, f.x, __fmul_rn(f.y, f.y)));
        mag[i] = m;
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(gmax, __float_as_uint(m));
}

__global__ void k_u8_hist(const float *__restrict__ mag, int n, const unsigned int *__restrict__ gmax, unsigned int *__restrict__ hist)
{
    __shared__ unsigned int sh[256];
    for (int j = threadIdx.x; j < 256; j += blockDim.x) sh[j] = 0;
    __syncthreads();
    const float scale = (float)(255.0 / (double)__uint_as_float(*gmax));
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int q = __float2int_rn(mag[i] * scale);
        atomicAdd(&sh[min(max(q, 0), 255)], 1u);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 256; j += blockDim.x)
        if (sh[j]) atomicAdd(&hist[j], sh[j]);
}

// DynaDetect.cc:1097-1114
__global__ void k_large_motion(const unsigned int *__restrict__ hist, const unsigned int *__restrict__ gmax, int W, int H, float scale_element,
                               int *__restrict__ out)
{
    if (threadIdx.x || blockIdx.x) return;
    double maxFlow = (double)__uint_as_float(*gmax);
    double ef = (double)(10.0f * scale_element * 255.0f) / maxFlow;
    int endFlow = ef >= 2147483647.0 ? 2147483647 : (int)ef;
    int endFlow2 = 0;
    float totalpixel = (float)(W * H) * scale_element * scale_element;
    float ratio = 0.0f;
    for (int i = 0; i < 255; ++i) {
        ratio += (float)hist[i];
        if (ratio > 0.3f * totalpixel) { endFlow2 = i; break; }
    }
    out[0] = endFlow2 > endFlow ? 1 : 0;
    out[1] = endFlow;
    out[2] = endFlow2;
}

int flow_branch_init(sindyn_ctx *c)
{
    SD_CHECK(c->dalloc(&c->fb_mag, (size_t)c->fw * c->fh));
    SD_CHECK(c->dalloc(&c->fb_hist, 260));
    SD_CHECK(c->dalloc(&c->fb_flag, 4));
    SD_CHECK(c->halloc(&c->fb_flag_host, 4));
    SD_CHECK(homography_init(c, &c->homog, c->W, c->H));
    SD_CHECK(varref_init(c, &c->varref, c->fw, c->fh));
    CU_CHECK(c, cudaEventCreateWithFlags(&c->ev_flag, cudaEventDisableTiming));
    return SINDYN_OK;
}

// Runs on the handle's resident frames (gsmall_f / gsmall of cur, last, lastlast). Result: c->flow_full.
// Split in two so that a caller can enqueue other work between the first Brox solve and the host decision:
//   flow_branch_begin: Brox(cur, lastlast), large-motion statistics, asynchronous copy of the flag
//   flow_branch_finish: wait for the flag (the only host decision of the flow branch; the reference does the same D2H + sync,
//                       DynaDetect.cc:1073), optional Brox(cur, last), refinement, up-sampling
int flow_branch_begin(sindyn_ctx *c)
{
    const bool g = c->cfg.use_graphs != 0;
    const int nf = c->fw * c->fh;
    SD_CHECK(brox_run(c, &c->brox, c->gsmall_f[c->i_cur], c->gsmall_f[c->i_lastlast], c->flow_small, -1.0f, g));
    if (c->cfg.stage_timing && c->ev_ok) CU_CHECK(c, cudaEventRecord(c->ev[2], c->stream));
    CU_CHECK(c, cudaMemsetAsync(c->fb_hist, 0, sizeof(unsigned int) * 260, c->stream));
    unsigned int *gmax = c->fb_hist + 256;
    LAUNCH(c, k_flow_mag, cdiv(nf, 256), 256, 0, (const float2 *)c->flow_small, nf, c->fb_mag, gmax);
    LAUNCH(c, k_u8_hist, SINDYN_NUM_SMS_B200, 256, 0, c->fb_mag, nf, gmax, c->fb_hist);
    LAUNCH(c, k_large_motion, 1, 32, 0, c->fb_hist, gmax, c->W, c->H, c->cfg.flow_scale, c->fb_flag);
    LAUNCH_CHECK(c);
    CU_CHECK(c, cudaMemcpyAsync(c->fb_flag_host, c->fb_flag, sizeof(int) * 4, cudaMemcpyDeviceToHost, c->stream));
    CU_CHECK(c, cudaEventRecord(c->ev_flag, c->stream));
    return SINDYN_OK;
}

// refinement + up-sampling (DynaDetect.cc:1133-1147)
static int flow_refine_upsample(sindyn_ctx *c, int i_ref)
{
    if (c->cfg.refine) SD_CHECK(varref_run(c, &c->varref, c->gsmall[c->i_cur], c->gsmall[i_ref], c->flow_small));
    SD_CHECK(launch_resize_flow(c, c->flow_small, c->fw, c->fh, c->flow_full, c->W, c->H, 1.0f / c->cfg.flow_scale));
    LAUNCH_CHECK(c);
    return SINDYN_OK;
}

int flow_branch_finish(sindyn_ctx *c, int *large_motion)
{
    const bool g = c->cfg.use_graphs != 0;
    CU_CHECK(c, cudaEventSynchronize(c->ev_flag));
    const int lm = c->fb_flag_host[0];
    if (large_motion) *large_motion = lm;
    int i_ref = c->i_lastlast;
    if (lm) {
        // second solver object: keeps both CUDA graphs alive
        SD_CHECK(brox_run(c, &c->brox_lm, c->gsmall_f[c->i_cur], c->gsmall_f[c->i_last], c->flow_small, -1.0f, g));
        i_ref = c->i_last;
    }
    return flow_refine_upsample(c, i_ref);
}

// Everything after the large-motion decision up to the two masks (DynaDetect.cc:1121-1367): optional second Brox solve,
// refinement, up-sampling, sample weighting + homography, residual + thresholds.  ~90 small launches whose host enqueue
// time would otherwise starve the GPU right after the host wait; with use_graphs they are captured once per frame-ring
// position (the key is the pair of gray images the refinement reads) and replayed as ONE graph launch.
void flow_tail_drop_graphs(sindyn_ctx *c)
{
    for (auto &t : c->tail) {
        if (t.exec) cudaGraphExecDestroy(t.exec);
        t = sindyn_ctx::TailGraph();
    }
    c->tail_next = 0;
}

int flow_finish_all(sindyn_ctx *c, int *large_motion)
{
    const bool timing = c->cfg.stage_timing && c->ev_ok;
    const bool g = c->cfg.use_graphs != 0 && !c->cfg.stage_timing;
    CU_CHECK(c, cudaEventSynchronize(c->ev_flag));
    const int lm = c->fb_flag_host[0];
    if (large_motion) *large_motion = lm;
    int i_ref = c->i_lastlast;
    if (lm) {
        SD_CHECK(brox_run(c, &c->brox_lm, c->gsmall_f[c->i_cur], c->gsmall_f[c->i_last], c->flow_small, -1.0f, c->cfg.use_graphs != 0));
        i_ref = c->i_last;
    }
    if (g) {
        const uint8_t *k0 = c->gsmall[c->i_cur], *k1 = c->gsmall[i_ref];
        sindyn_ctx::TailGraph *t = nullptr;
        for (auto &q : c->tail)
            if (q.exec && q.k0 == k0 && q.k1 == k1 && q.stream == c->stream) t = &q;
        if (!t) {
            t = &c->tail[c->tail_next];
            c->tail_next = (c->tail_next + 1) % 8;
            if (t->exec) cudaGraphExecDestroy(t->exec);
            *t = sindyn_ctx::TailGraph();
            const unsigned long long before = c->launches;
            cudaGraph_t gr = nullptr;
            CU_CHECK(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            int st = flow_refine_upsample(c, i_ref);
            if (st == SINDYN_OK) st = homography_sample(c, &c->homog, c->flow_full, c->label_last, c->dyna_last);
            if (st == SINDYN_OK) st = homography_estimate(c, &c->homog);
            if (st == SINDYN_OK) st = residual_homography_run_dev(c, &c->resid, c->flow_full, c->homog.H_dev, c->mask_low, c->mask_high);
            cudaError_t e = cudaStreamEndCapture(c->stream, &gr);
            t->launches = c->launches - before;
            c->launches = before;
            SD_CHECK(st);
            CU_CHECK(c, e);
            CU_CHECK(c, cudaGraphInstantiate(&t->exec, gr, 0));
            cudaGraphDestroy(gr);
            t->k0 = k0; t->k1 = k1; t->stream = c->stream;
        }
        CU_CHECK(c, cudaGraphLaunch(t->exec, c->stream));
        c->launches += t->launches;
        return SINDYN_OK;
    }
    SD_CHECK(flow_refine_upsample(c, i_ref));
    if (timing) CU_CHECK(c, cudaEventRecord(c->ev[3], c->stream));
    SD_CHECK(homography_sample(c, &c->homog, c->flow_full, c->label_last, c->dyna_last));
    SD_CHECK(homography_estimate(c, &c->homog));
    if (timing) CU_CHECK(c, cudaEventRecord(c->ev[4], c->stream));
    SD_CHECK(residual_homography_run_dev(c, &c->resid, c->flow_full, c->homog.H_dev, c->mask_low, c->mask_high));
    if (timing) CU_CHECK(c, cudaEventRecord(c->ev[5], c->stream));
    return SINDYN_OK;
}

int flow_branch_run(sindyn_ctx *c, int *large_motion)
{
    SD_CHECK(flow_branch_begin(c));
    return flow_branch_finish(c, large_motion);
}

extern "C" int sindyn_flow_branch(sindyn_handle h, const uint8_t *bgr_cur, size_t step_cur, float *flow_out, int *large_motion_out)
{
    H_CHECK(h);
    if (!bgr_cur || !flow_out) return SINDYN_ERR_INVALID;
    if (!h->have_prev) { h->err = "sindyn_flow_branch: call sindyn_set_prev_frames first"; return SINDYN_ERR_STATE; }
    CU_CHECK(h, copy_in_2d(h->bgr[h->i_cur], bgr_cur, step_cur, (size_t)h->W * 3, h->H, h->stream));
    SD_CHECK(sindyn_prep_frame(h, h->i_cur));
    SD_CHECK(flow_branch_run(h, large_motion_out));
    CU_CHECK(h, cudaMemcpyAsync(flow_out, h->flow_full, sizeof(float) * 2 * h->N, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    return SINDYN_OK;
}

extern "C" int sindyn_flow_refine(sindyn_handle h, const uint8_t *I0, const uint8_t *I1, int w, int hgt, float *flow_uv)
{
    H_CHECK(h);
    if (!I0 || !I1 || !flow_uv || w != h->fw || hgt != h->fh) { h->err = "sindyn_flow_refine: size must equal the flow grid"; return SINDYN_ERR_INVALID; }
    const size_t n = (size_t)w * hgt;
    CU_CHECK(h, cudaMemcpyAsync(h->scratch_u0, I0, n, cudaMemcpyHostToDevice, h->stream));
    CU_CHECK(h, cudaMemcpyAsync(h->scratch_u1, I1, n, cudaMemcpyHostToDevice, h->stream));
    CU_CHECK(h, cudaMemcpyAsync(h->flow_small, flow_uv, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, h->stream));
    SD_CHECK(varref_run(h, &h->varref, h->scratch_u0, h->scratch_u1, h->flow_small));
    CU_CHECK(h, cudaMemcpyAsync(flow_uv, h->flow_small, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    return SINDYN_OK;
}

extern "C" int sindyn_sample_pairs(sindyn_handle h, const float *flow, float *pts, float *pts_last, int capacity, int *n_out)
{
    H_CHECK(h);
    if (!flow || !n_out) return SINDYN_ERR_INVALID;
    CU_CHECK(h, cudaMemcpyAsync(h->flow_full, flow, sizeof(float) * 2 * h->N, cudaMemcpyHostToDevice, h->stream));
    SD_CHECK(homography_sample(h, &h->homog, h->flow_full, h->label_last, h->dyna_last));
    int n = 0;
    CU_CHECK(h, cudaMemcpyAsync(&n, h->homog.n_pairs, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    *n_out = n;
    if (n > capacity) { h->err = "sindyn_sample_pairs: capacity too small"; return SINDYN_ERR_CAPACITY; }
    if (pts) CU_CHECK(h, cudaMemcpy(pts, h->homog.pts, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost));
    if (pts_last) CU_CHECK(h, cudaMemcpy(pts_last, h->homog.pts_last, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost));
    return SINDYN_OK;
}

extern "C" int sindyn_estimate_homography(sindyn_handle h, const float *flow, double *H_out, int *n_pairs_out)
{
    H_CHECK(h);
    if (!flow || !H_out) return SINDYN_ERR_INVALID;
    CU_CHECK(h, cudaMemcpyAsync(h->flow_full, flow, sizeof(float) * 2 * h->N, cudaMemcpyHostToDevice, h->stream));
    SD_CHECK(homography_sample(h, &h->homog, h->flow_full, h->label_last, h->dyna_last));
    SD_CHECK(homography_estimate(h, &h->homog));
    int info[4];
    CU_CHECK(h, cudaMemcpyAsync(info, h->homog.n_pairs, sizeof info, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaMemcpyAsync(H_out, h->homog.H_dev, sizeof(double) * 9, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    if (n_pairs_out) *n_pairs_out = info[0];
    return SINDYN_OK;
}
