// flow.cu -- the flow branch of DynaDetect (ORB_SLAM2/src/DynaDetect.cc:1023-1147,1163-1235) and its C-ABI
// entry points: Brox(cur, lastlast) -> negate -> large-motion test -> optional Brox(cur, last) -> variational
// refinement -> up-sample x 1/0.6 -> sample weighting -> homography.
#include "ctx.cuh"

#include <cstdio>

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
// cond / use_cond: inside the captured flow graph the flag also drives the IF node that holds the second Brox solve
__global__ void k_large_motion(const unsigned int *__restrict__ hist, const unsigned int *__restrict__ gmax, int W, int H, float scale_element,
                               int *__restrict__ out, cudaGraphConditionalHandle cond, int use_cond)
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
    if (use_cond) cudaGraphSetConditional(cond, endFlow2 > endFlow ? 1u : 0u);
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

// ---------------------------------------------------------------- the whole flow branch as ONE graph
// Brox(cur, lastlast) -> large-motion statistics -> IF(large motion) { Brox(cur, last) } -> refinement (reference image chosen by
// the same device flag) -> up-sampling -> sample weighting + homography -> residual + thresholds -> masks.  The large-motion
// decision, which the reference takes on the host after a D2H copy (DynaDetect.cc:1073-1114), is a conditional graph node:
// no host round trip in the middle of the frame, one graph launch per frame.  One graph per frame-ring position (the key is
// the triple of resident gray images).  Graph construction mixes stream capture with one explicitly added IF node.
void flow_graph_drop(sindyn_ctx *c)
{
    for (auto &t : c->flow_graph) {
        if (t.exec) cudaGraphExecDestroy(t.exec);
        if (t.graph) cudaGraphDestroy(t.graph);
        t = sindyn_ctx::FlowGraph();
    }
    c->flow_graph_next = 0;
}

// variant 0: the whole branch; variant 1 (frame pipeline): up to the up-sampled flow only
static int flow_graph_build(sindyn_ctx *c, sindyn_ctx::FlowGraph *t, int variant, const FlowRes &r)
{
    const int nf = c->fw * c->fh;
    cudaGraph_t g = nullptr;
    CU_CHECK(c, cudaGraphCreate(&g, 0));
    t->graph = g;
    cudaGraphConditionalHandle cond;
    CU_CHECK(c, cudaGraphConditionalHandleCreate(&cond, g, 0, cudaGraphCondAssignDefault));
    const unsigned long long before = c->launches;
    CU_CHECK(c, cudaStreamBeginCaptureToGraph(c->stream, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    int st = brox_run(c, r.brox, c->gsmall_f[c->i_cur], c->gsmall_f[c->i_lastlast], r.flow_small, -1.0f, false);
    unsigned int *gmax = r.fb_hist + 256;
    cudaGraphNode_t cnode = nullptr;
    cudaGraph_t body = nullptr;
    cudaError_t e = cudaSuccess;
    if (st == SINDYN_OK) {
        e = cudaMemsetAsync(r.fb_hist, 0, sizeof(unsigned int) * 260, c->stream);
        LAUNCH(c, k_flow_mag, cdiv(nf, 256), 256, 0, (const float2 *)r.flow_small, nf, r.fb_mag, gmax);
        LAUNCH(c, k_u8_hist, SINDYN_NUM_SMS_B200, 256, 0, r.fb_mag, nf, gmax, r.fb_hist);
        LAUNCH(c, k_large_motion, 1, 32, 0, r.fb_hist, gmax, c->W, c->H, c->cfg.flow_scale, r.fb_flag, cond, 1);
        // the IF node hangs off everything captured so far; what is captured next hangs off the IF node
        cudaStreamCaptureStatus cs;
        const cudaGraphNode_t *deps = nullptr;
        size_t ndeps = 0;
        if (e == cudaSuccess) e = cudaStreamGetCaptureInfo_v2(c->stream, &cs, nullptr, nullptr, &deps, &ndeps);
        cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};   // (aggregate initialisation: the union member has no default constructor)
        cp.type = cudaGraphNodeTypeConditional;
        cp.conditional.handle = cond;
        cp.conditional.type = cudaGraphCondTypeIf;
        cp.conditional.size = 1;
        if (e == cudaSuccess) e = cudaGraphAddNode(&cnode, g, deps, ndeps, &cp);
        if (e == cudaSuccess) body = cp.conditional.phGraph_out[0];
        if (e == cudaSuccess) e = cudaStreamUpdateCaptureDependencies(c->stream, &cnode, 1, cudaStreamSetCaptureDependencies);
    }
    if (st == SINDYN_OK && e == cudaSuccess) {
        if (c->cfg.refine)
            st = varref_run_sel(c, r.varref, c->gsmall[c->i_cur], c->gsmall[c->i_lastlast], c->gsmall[c->i_last], r.fb_flag, r.flow_small);
        if (st == SINDYN_OK) st = launch_resize_flow(c, r.flow_small, c->fw, c->fh, r.flow_full, c->W, c->H, 1.0f / c->cfg.flow_scale);
        if (variant == 0) {
            if (st == SINDYN_OK) st = homography_sample(c, &c->homog, r.flow_full, c->label_last, c->dyna_last);
            if (st == SINDYN_OK) st = homography_estimate(c, &c->homog);
            if (st == SINDYN_OK) st = residual_homography_run_dev(c, &c->resid, r.flow_full, c->homog.H_dev, c->mask_low, c->mask_high);
        }
        if (st == SINDYN_OK) e = cudaMemcpyAsync(r.fb_flag_host, r.fb_flag, sizeof(int) * 4, cudaMemcpyDeviceToHost, c->stream);
    }
    cudaGraph_t g_out = nullptr;
    const cudaError_t e_end = cudaStreamEndCapture(c->stream, &g_out);
    t->launches = c->launches - before;
    // body of the IF node: the second Brox solve (its own solver object: both flows' scratch stays alive)
    cudaError_t e_body = cudaSuccess;
    unsigned long long body_launches = 0;
    if (st == SINDYN_OK && e == cudaSuccess && e_end == cudaSuccess && body) {
        const unsigned long long b0 = c->launches;
        e_body = cudaStreamBeginCaptureToGraph(c->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
        if (e_body == cudaSuccess) {
            st = brox_run(c, r.brox_lm, c->gsmall_f[c->i_cur], c->gsmall_f[c->i_last], r.flow_small, -1.0f, false);
            cudaGraph_t b_out = nullptr;
            e_body = cudaStreamEndCapture(c->stream, &b_out);
        }
        body_launches = c->launches - b0;
    }
    c->launches = before;
    SD_CHECK(st);
    CU_CHECK(c, e);
    CU_CHECK(c, e_end);
    CU_CHECK(c, e_body);
    CU_CHECK(c, cudaGraphInstantiate(&t->exec, g, 0));
    t->body_launches = body_launches;
    t->k0 = c->gsmall[c->i_cur]; t->k1 = c->gsmall[c->i_last]; t->k2 = c->gsmall[c->i_lastlast]; t->stream = c->stream;
    t->variant = variant;
    return SINDYN_OK;
}

// launches the flow graph of the current frame-ring position; returns SINDYN_OK with *launched = false when this device /
// driver cannot build it (the caller then takes the two-graph path with the host decision)
static int flow_graph_launch(sindyn_ctx *c, bool *launched, int variant = 0)
{
    *launched = false;
    if (c->flow_graph_broken) return SINDYN_OK;
    const uint8_t *k0 = c->gsmall[c->i_cur], *k1 = c->gsmall[c->i_last], *k2 = c->gsmall[c->i_lastlast];
    sindyn_ctx::FlowGraph *t = nullptr;
    for (auto &q : c->flow_graph)
        if (q.exec && q.k0 == k0 && q.k1 == k1 && q.k2 == k2 && q.stream == c->stream && q.variant == variant) t = &q;
    if (!t) {
        t = &c->flow_graph[c->flow_graph_next];
        c->flow_graph_next = (c->flow_graph_next + 1) % 12;
        if (t->exec) cudaGraphExecDestroy(t->exec);
        if (t->graph) cudaGraphDestroy(t->graph);
        *t = sindyn_ctx::FlowGraph();
        const FlowRes own = {&c->brox, &c->brox_lm, &c->varref, c->fb_mag, c->fb_hist, c->fb_flag, c->fb_flag_host, c->flow_small, c->flow_full};
        if (flow_graph_build(c, t, variant, variant ? pipe_flow_res(c, variant - 1) : own) != SINDYN_OK) {
            // leave capture mode if the failure happened inside it, remember not to try again
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing(c->stream, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) { cudaGraph_t junk = nullptr; cudaStreamEndCapture(c->stream, &junk); }
            cudaGetLastError();
            if (t->exec) cudaGraphExecDestroy(t->exec);
            if (t->graph) cudaGraphDestroy(t->graph);
            *t = sindyn_ctx::FlowGraph();
            c->flow_graph_broken = true;
            fprintf(stderr, "sindyn: one-graph flow branch unavailable (%s); using the two-graph path with the host decision\n", c->err.c_str());
            return SINDYN_OK;
        }
    }
    CU_CHECK(c, cudaGraphLaunch(t->exec, c->stream));
    c->launches += t->launches;   // (+ t->body_launches on large-motion frames: added when the flag is read)
    c->flow_graph_last = t;
    c->flow_flag_pending = true;
    *launched = true;
    return SINDYN_OK;
}

// the large-motion flag of the last flow-graph launch, valid after the stream has been synchronised
void flow_collect_flag(sindyn_ctx *c)
{
    if (!c->flow_flag_pending) return;
    c->flow_flag_pending = false;
    c->large_motion_last = c->fb_flag_host[0];
    if (c->large_motion_last && c->flow_graph_last) c->launches += c->flow_graph_last->body_launches;
}

// ---------------------------------------------------------------- frame pipeline (pipe.cu): the branch in two parts
// Part A depends on the three gray images only, part B also on the previous frame's decision (label_last, dyna_last).  The
// caller has pointed c->flow_full / c->fb_flag / c->fb_flag_host at the buffers of `parity` and c->stream at the stream of the
// part; variant = 1 + parity keys the graph cache (the pointers are baked into the graph).
int flow_part_a(sindyn_ctx *c, int parity)
{
    bool launched = false;
    SD_CHECK(flow_graph_launch(c, &launched, 1 + parity));
    if (!launched) { c->err = "frame pipeline: the flow graph could not be built"; return SINDYN_ERR_CUDA; }
    return SINDYN_OK;
}

int flow_part_b(sindyn_ctx *c, int parity)
{
    sindyn_ctx::TailGraph *t = nullptr;
    const uint8_t *k0 = (const uint8_t *)c->flow_full, *k1 = (const uint8_t *)(uintptr_t)(0x100 + parity);
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
        int st = homography_sample(c, &c->homog, c->flow_full, c->label_last, c->dyna_last);
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

// Runs on the handle's resident frames (gsmall_f / gsmall of cur, last, lastlast). Result: c->flow_full.
// Split in two so that a caller can enqueue other work between the first Brox solve and the host decision:
//   flow_branch_begin: Brox(cur, lastlast), large-motion statistics, asynchronous copy of the flag
//   flow_branch_finish: wait for the flag (the only host decision of the flow branch; the reference does the same D2H + sync,
//                       DynaDetect.cc:1073), optional Brox(cur, last), refinement, up-sampling
// whole_frame: the caller will go on with flow_finish_all (homography, residual, masks), so the entire flow branch may be
// launched as the one captured graph; the stage-level entry point (flow only) passes false
int flow_branch_begin(sindyn_ctx *c, bool whole_frame)
{
    const bool g = c->cfg.use_graphs != 0;
    const int nf = c->fw * c->fh;
    c->flow_graph_active = false;
    if (whole_frame && g && !c->cfg.stage_timing && c->flow_one_graph) {
        bool launched = false;
        SD_CHECK(flow_graph_launch(c, &launched));
        if (launched) { c->flow_graph_active = true; return SINDYN_OK; }
    }
    SD_CHECK(brox_run(c, &c->brox, c->gsmall_f[c->i_cur], c->gsmall_f[c->i_lastlast], c->flow_small, -1.0f, g));
    if (c->cfg.stage_timing && c->ev_ok) CU_CHECK(c, cudaEventRecord(c->ev[2], c->stream));
    CU_CHECK(c, cudaMemsetAsync(c->fb_hist, 0, sizeof(unsigned int) * 260, c->stream));
    unsigned int *gmax = c->fb_hist + 256;
    LAUNCH(c, k_flow_mag, cdiv(nf, 256), 256, 0, (const float2 *)c->flow_small, nf, c->fb_mag, gmax);
    LAUNCH(c, k_u8_hist, SINDYN_NUM_SMS_B200, 256, 0, c->fb_mag, nf, gmax, c->fb_hist);
    LAUNCH(c, k_large_motion, 1, 32, 0, c->fb_hist, gmax, c->W, c->H, c->cfg.flow_scale, c->fb_flag, (cudaGraphConditionalHandle)0, 0);
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
    if (c->flow_graph_active) {   // everything is already enqueued; the flag arrives with the next synchronisation
        if (large_motion) *large_motion = c->large_motion_last;
        return SINDYN_OK;
    }
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
    SD_CHECK(flow_branch_begin(c, false));
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

extern "C" int sindyn_find_homography_rho(sindyn_handle h, const float *src_xy, const float *dst_xy, int n, double *H_out, uint8_t *inlier_mask_out,
                                          int *info_out)
{
    if (!h) return SINDYN_ERR_INVALID;
    cudaSetDevice(h->device);
    if (!src_xy || !dst_xy || !H_out || n < 0 || n > HG_MAX_SAMPLES) return SINDYN_ERR_INVALID;
    int info[12] = {n, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    CU_CHECK(h, cudaMemcpyAsync(h->homog.n_pairs, info, sizeof info, cudaMemcpyHostToDevice, h->stream));
    if (n) {
        CU_CHECK(h, cudaMemcpyAsync(h->homog.pts, src_xy, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, h->stream));
        CU_CHECK(h, cudaMemcpyAsync(h->homog.pts_last, dst_xy, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, h->stream));
    }
    SD_CHECK(homography_estimate(h, &h->homog));
    CU_CHECK(h, cudaMemcpyAsync(info, h->homog.n_pairs, sizeof info, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaMemcpyAsync(H_out, h->homog.H_dev, sizeof(double) * 9, cudaMemcpyDeviceToHost, h->stream));
    if (inlier_mask_out && n) CU_CHECK(h, cudaMemcpyAsync(inlier_mask_out, h->homog.inl_mask, n, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    if (info_out) for (int k = 0; k < 12; ++k) info_out[k] = info[k];
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
