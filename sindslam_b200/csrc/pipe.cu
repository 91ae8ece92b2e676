// pipe.cu -- software pipeline over consecutive frames of one sequence.
//
// One frame of DetectDynaArea (DynaDetect.cc:1377-1666) is a chain of ~4 ms on an otherwise nearly idle GPU, but only part of it
// depends on the previous frame's result:
//   A  gray / resize, Brox, large-motion decision, refinement, up-sampling      <- the three gray images only
//   P  PEAC plane fitter + plane contours                                        <- the depth image only
//   B  sample weighting, RHO homography, residual, masks                          <- A + the previous frame's labels / dynamic mask
//   C  k-means (warm-started by the previous labels), gradient edges, plane-edge filter (needs P), re-clustering
//   D  decision, state roll                                                       <- B + C
// The reference's driver hands the detector one frame after the other (rgbd_tum_noros.cc:113-192); nothing in A or P reads the
// detector's state.  So A and P of frame i + 1 run on their own streams while B, C, D of frame i are still in flight: the host
// enqueues a frame without waiting (CUDA graphs, events), and the streams only wait for what they really read.  Everything a
// later frame's A / P overwrites while an earlier frame still reads it exists twice (parity of the frame number): the up-sampled
// flow, the large-motion flag, the depth image, the plane-edge image, the plane fitter and the scratch of its contour pass.
// Results are bit-identical to the frame-at-a-time path (same kernels, same inputs, same order per buffer).
#include "ctx.cuh"

#define PIPE_NB 4
struct FramePipe {
    cudaStream_t sa[2] = {nullptr, nullptr};        // part A (+ the input copies) of even / odd frames: two frames' flow solves overlap
    BroxSolver brox[2], brox_lm[2];                 // ... each in its own solver / refinement / scratch (index 1; index 0 = the handle's)
    VarRefStage varref1;
    float *fb_mag1 = nullptr, *flow_small1 = nullptr;
    unsigned int *fb_hist1 = nullptr;
    cudaStream_t sc = nullptr;                      // the state chain (part B, decision, state roll): highest priority, its ~30 kernels are
                                                    // what bounds the frame rate and must not queue behind a wave of flow-solve CTAs
    cudaStream_t sp[PIPE_NB] = {};                  // gradient edges + plane fitter of frame k (2.9 ms each, more under load: up to four frames' fitters overlap)
    // buffers a later frame's image-only stages overwrite while an earlier frame still reads them: ring of PIPE_NB, k = frame % PIPE_NB
    float *flow_full[PIPE_NB] = {};
    int *fb_flag[PIPE_NB] = {}, *fb_flag_host[PIPE_NB] = {};
    uint16_t *depth[PIPE_NB] = {};
    uint8_t *plane_edges[PIPE_NB] = {};
    PeacStage peac[PIPE_NB];
    ReclusterStage rc_peac[PIPE_NB];
    // the decision's own CCL scratch and two label images: the clustering of frame i + 1 (warm-started by frame i's labels) no
    // longer has to wait for the decision of frame i, which reads those labels and used to share the re-clustering's scratch
    uint8_t *dd_cls = nullptr; int *dd_labels = nullptr, *dd_top = nullptr; RegionStats *dd_stats = nullptr;
    uint8_t *label_out[2] = {nullptr, nullptr}, *own_label_out = nullptr;
    EdgeStage edges[PIPE_NB], own_edges;            // gradient edges / end points / total area of a frame (depth only: run ahead, read until the decision)
    cudaEvent_t ev_in[PIPE_NB] = {}, ev_a[PIPE_NB] = {}, ev_p[PIPE_NB] = {}, ev_done[PIPE_NB] = {}, ev_e[PIPE_NB] = {}, ev_join = nullptr, ev_sync = nullptr, ev_gray[4] = {};
    bool gray_pending[4] = {false, false, false, false};
    cudaEvent_t ev_ddout = nullptr;                 // the extractor's stream has read the final mask (dd.out) of the last frame
    bool ddout_pending = false;
    cudaGraphExec_t g_c1[PIPE_NB] = {}, g_c2[PIPE_NB] = {}, g_p[PIPE_NB] = {}, g_e[PIPE_NB] = {}, g_d[PIPE_NB] = {};
    unsigned long long n_c1[PIPE_NB] = {}, n_c2[PIPE_NB] = {}, n_p[PIPE_NB] = {}, n_e[PIPE_NB] = {}, n_d[PIPE_NB] = {};
    cudaStream_t built_for = nullptr;               // the handle stream the graphs were captured under
    unsigned long long frame_no = 0;
    bool fresh = true;                              // the pipeline's streams have to wait for the handle's stream first
    int last_slot = 0, last_k = 0;
    int hdr_host[PIPE_NB][4] = {};
    uint8_t *in_bgr[2] = {nullptr, nullptr};        // pinned bounce buffers for pageable host inputs (asynchronous submit)
    uint16_t *in_depth[2] = {nullptr, nullptr};
    cudaEvent_t ev_h2d[2] = {};
    // what the handle's own pointers were before the pipeline redirected them
    float *own_flow_full = nullptr; int *own_fb_flag = nullptr, *own_fb_flag_host = nullptr; uint16_t *own_depth = nullptr; uint8_t *own_plane_edges = nullptr;
};

bool pipe_usable(const sindyn_ctx *c) { return c->cfg.use_graphs && !c->cfg.stage_timing && c->flow_one_graph && !c->flow_graph_broken; }

static void pipe_drop_graphs(FramePipe *P)
{
    for (int p = 0; p < PIPE_NB; ++p) {
        if (P->g_c1[p]) cudaGraphExecDestroy(P->g_c1[p]);
        if (P->g_c2[p]) cudaGraphExecDestroy(P->g_c2[p]);
        if (P->g_p[p]) cudaGraphExecDestroy(P->g_p[p]);
        if (P->g_e[p]) cudaGraphExecDestroy(P->g_e[p]);
        if (P->g_d[p]) cudaGraphExecDestroy(P->g_d[p]);
        P->g_c1[p] = P->g_c2[p] = P->g_p[p] = P->g_e[p] = P->g_d[p] = nullptr;
    }
}

static int pipe_init(sindyn_ctx *c)
{
    if (c->pipe) return SINDYN_OK;
    FramePipe *P = new FramePipe();
    c->pipe = P;
    CU_CHECK(c, cudaStreamCreateWithFlags(&P->sa[0], cudaStreamNonBlocking));
    CU_CHECK(c, cudaStreamCreateWithFlags(&P->sa[1], cudaStreamNonBlocking));
    {
        const sindyn_config &g = c->cfg;
        SD_CHECK(brox_init(c, &P->brox[1], c->fw, c->fh, g.brox_alpha, g.brox_gamma, g.brox_pyr_scale, g.brox_inner, g.brox_outer, g.brox_solver, g.brox_omega));
        SD_CHECK(brox_init(c, &P->brox_lm[1], c->fw, c->fh, g.brox_alpha, g.brox_gamma, g.brox_pyr_scale, g.brox_inner, g.brox_outer, g.brox_solver, g.brox_omega));
        SD_CHECK(varref_init(c, &P->varref1, c->fw, c->fh));
        SD_CHECK(c->dalloc(&P->fb_mag1, (size_t)c->fw * c->fh));
        SD_CHECK(c->dalloc(&P->fb_hist1, 260));
        SD_CHECK(c->dalloc(&P->flow_small1, (size_t)c->fw * c->fh * 2));
    }
    {
        int lo = 0, hi = 0;
        CU_CHECK(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU_CHECK(c, cudaStreamCreateWithPriority(&P->sc, cudaStreamNonBlocking, hi));
    }
    P->sp[0] = c->stream3;
    for (int k = 1; k < PIPE_NB; ++k) CU_CHECK(c, cudaStreamCreateWithFlags(&P->sp[k], cudaStreamNonBlocking));
    P->own_flow_full = c->flow_full; P->own_fb_flag = c->fb_flag; P->own_fb_flag_host = c->fb_flag_host; P->own_depth = c->depth; P->own_plane_edges = c->plane_edges;
    P->flow_full[0] = c->flow_full; P->fb_flag[0] = c->fb_flag; P->fb_flag_host[0] = c->fb_flag_host; P->depth[0] = c->depth; P->plane_edges[0] = c->plane_edges;
    {
        const size_t N = (size_t)c->N;
        SD_CHECK(c->dalloc(&P->dd_cls, N * (DD_MAXL + 2)));
        SD_CHECK(c->dalloc(&P->dd_labels, (N + 1) * (DD_MAXL + 2)));
        SD_CHECK(c->dalloc(&P->dd_top, N * DD_MAXL));
        SD_CHECK(c->dalloc(&P->dd_stats, N * DD_MAXL));
        P->own_label_out = c->rc.label_out;
        P->label_out[0] = c->rc.label_out;
        SD_CHECK(c->dalloc(&P->label_out[1], N));
    }
    P->own_edges = c->edges;
    P->edges[0] = c->edges;
    for (int k = 1; k < PIPE_NB; ++k) {
        SD_CHECK(edges_init(c, &P->edges[k], c->W, c->H));
        SD_CHECK(c->dalloc(&P->flow_full[k], (size_t)c->N * 2));
        SD_CHECK(c->dalloc(&P->fb_flag[k], 4));
        SD_CHECK(c->halloc(&P->fb_flag_host[k], 4));
        SD_CHECK(c->dalloc(&P->depth[k], (size_t)c->N));
        SD_CHECK(c->dalloc(&P->plane_edges[k], (size_t)c->N));
    }
    for (int k = 0; k < PIPE_NB; ++k) {
        CU_CHECK(c, cudaEventCreateWithFlags(&P->ev_in[k], cudaEventDisableTiming));
        CU_CHECK(c, cudaEventCreateWithFlags(&P->ev_a[k], cudaEventDisableTiming));
        CU_CHECK(c, cudaEventCreateWithFlags(&P->ev_p[k], cudaEventDisableTiming));
        CU_CHECK(c, cudaEventCreateWithFlags(&P->ev_done[k], cudaEventDisableTiming));
        CU_CHECK(c, cudaEventCreateWithFlags(&P->ev_e[k], cudaEventDisableTiming));
    }
    for (int k = 0; k < PIPE_NB; ++k)
        if (c->cfg.plane_edges) {
            SD_CHECK(peac_init(c, &P->peac[k], c->W, c->H));
            SD_CHECK(recluster_init(c, &P->rc_peac[k], c->W, c->H));
        }
    for (int p = 0; p < 2; ++p) CU_CHECK(c, cudaEventCreateWithFlags(&P->ev_h2d[p], cudaEventDisableTiming));
    for (int k = 0; k < 4; ++k) CU_CHECK(c, cudaEventCreateWithFlags(&P->ev_gray[k], cudaEventDisableTiming));
    CU_CHECK(c, cudaEventCreateWithFlags(&P->ev_join, cudaEventDisableTiming));
    CU_CHECK(c, cudaEventCreateWithFlags(&P->ev_ddout, cudaEventDisableTiming));
    CU_CHECK(c, cudaEventCreateWithFlags(&P->ev_sync, cudaEventDisableTiming));
    return SINDYN_OK;
}

void pipe_destroy(sindyn_ctx *c)
{
    FramePipe *P = c->pipe;
    if (!P) return;
    pipe_drop_graphs(P);
    c->edges = P->own_edges;
    c->rc.label_out = P->own_label_out;
    c->flow_full = P->own_flow_full; c->fb_flag = P->own_fb_flag; c->fb_flag_host = P->own_fb_flag_host; c->depth = P->own_depth; c->plane_edges = P->own_plane_edges;
    for (int k = 0; k < PIPE_NB; ++k) { cudaEventDestroy(P->ev_in[k]); cudaEventDestroy(P->ev_a[k]); cudaEventDestroy(P->ev_p[k]); cudaEventDestroy(P->ev_done[k]); cudaEventDestroy(P->ev_e[k]); }
    for (int p = 0; p < 2; ++p) cudaEventDestroy(P->ev_h2d[p]);
    for (int k = 0; k < 4; ++k) cudaEventDestroy(P->ev_gray[k]);
    brox_destroy(&P->brox[1]); brox_destroy(&P->brox_lm[1]);
    cudaEventDestroy(P->ev_join); cudaEventDestroy(P->ev_sync); cudaEventDestroy(P->ev_ddout);
    if (P->sa[0]) cudaStreamDestroy(P->sa[0]);
    if (P->sa[1]) cudaStreamDestroy(P->sa[1]);
    for (int k = 1; k < PIPE_NB; ++k) if (P->sp[k]) cudaStreamDestroy(P->sp[k]);
    if (P->sc) cudaStreamDestroy(P->sc);
    delete P;
    c->pipe = nullptr;
}

void pipe_invalidate(sindyn_ctx *c)
{
    if (c->pipe) c->pipe->fresh = true;
}

// the handle's stream waits for everything enqueued on the pipeline's streams
int pipe_join(sindyn_ctx *c)
{
    FramePipe *P = c->pipe;
    if (!P) return SINDYN_OK;
    cudaStream_t ss[4 + PIPE_NB] = {P->sa[0], P->sa[1], c->stream2, P->sc};
    for (int k = 0; k < PIPE_NB; ++k) ss[4 + k] = P->sp[k];
    for (cudaStream_t s : ss) {
        CU_CHECK(c, cudaEventRecord(P->ev_sync, s));
        CU_CHECK(c, cudaStreamWaitEvent(c->stream, P->ev_sync, 0));
    }
    P->fresh = true;     // ... and the pipeline's next frame follows whatever the handle's stream holds by then (a join is two-way)
    return SINDYN_OK;
}

FlowRes pipe_flow_res(sindyn_ctx *c, int k)     // k = frame % PIPE_NB: output buffers k, solver set k & 1
{
    FramePipe *P = c->pipe;
    if ((k & 1) == 0) return FlowRes{&c->brox, &c->brox_lm, &c->varref, c->fb_mag, c->fb_hist, P->fb_flag[k], P->fb_flag_host[k], c->flow_small, P->flow_full[k]};
    return FlowRes{&P->brox[1], &P->brox_lm[1], &P->varref1, P->fb_mag1, P->fb_hist1, P->fb_flag[k], P->fb_flag_host[k], P->flow_small1, P->flow_full[k]};
}

cudaEvent_t pipe_input_event(sindyn_ctx *c) { return c->pipe ? c->pipe->ev_in[c->pipe->last_k] : nullptr; }
cudaStream_t pipe_chain_stream(sindyn_ctx *c) { return c->pipe ? c->pipe->sc : c->stream; }
cudaEvent_t pipe_done_event(sindyn_ctx *c) { return c->pipe ? c->pipe->ev_done[c->pipe->last_k] : nullptr; }

// the caller dilates / copies the last frame's final mask (dd.out) on another stream: the next decision overwrites it only after that
int pipe_note_mask_read(sindyn_ctx *c, cudaStream_t s)
{
    FramePipe *P = c->pipe;
    if (!P) return SINDYN_OK;
    CU_CHECK(c, cudaEventRecord(P->ev_ddout, s));
    P->ddout_pending = true;
    return SINDYN_OK;
}

// the ORB extractor reads the BGR ring slot of the frame on its own stream: the slot may be overwritten only after that
int pipe_note_gray_read(sindyn_ctx *c, cudaStream_t orb_stream)
{
    FramePipe *P = c->pipe;
    if (!P) return SINDYN_OK;
    CU_CHECK(c, cudaEventRecord(P->ev_gray[P->last_slot], orb_stream));
    P->gray_pending[P->last_slot] = true;
    return SINDYN_OK;
}

int pipe_copy_headers(sindyn_ctx *c)
{
    FramePipe *P = c->pipe;
    if (!P || !c->cfg.plane_edges) return SINDYN_OK;
    for (int p = 0; p < PIPE_NB; ++p)
        if (P->peac[p].built) SD_CHECK(peac_copy_sticky_overflow(c, &P->peac[p], &P->hdr_host[p][2]));
    return SINDYN_OK;
}

bool pipe_overflow(const sindyn_ctx *c)
{
    const FramePipe *P = c->pipe;
    if (!P) return false;
    for (int p = 0; p < PIPE_NB; ++p)
        if (P->hdr_host[p][2]) return true;
    return false;
}

// capture fn() on stream s into an executable graph
template <class F> static int pipe_capture(sindyn_ctx *c, cudaStream_t s, cudaGraphExec_t *exec, unsigned long long *n_launches, F fn)
{
    cudaStream_t keep = c->stream;
    const unsigned long long before = c->launches;
    cudaGraph_t gr = nullptr;
    CU_CHECK(c, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    c->stream = s;
    const int st = fn();
    c->stream = keep;
    const cudaError_t e = cudaStreamEndCapture(s, &gr);
    *n_launches = c->launches - before;
    c->launches = before;
    SD_CHECK(st);
    CU_CHECK(c, e);
    CU_CHECK(c, cudaGraphInstantiate(exec, gr, 0));
    cudaGraphDestroy(gr);
    return SINDYN_OK;
}

int pipe_detect_run(sindyn_ctx *c, const uint8_t *bgr_dev, const uint16_t *depth_dev) { return pipe_detect_run_src(c, bgr_dev, 0, depth_dev, 0, false, nullptr); }

// host_src: bgr / depth are (pitched) host images; flags (pinned, optional): capacity flags of this frame, copied before a later
// frame can reset them -- valid once the handle's stream has passed the event the caller records next
int pipe_detect_run_src(sindyn_ctx *c, const uint8_t *bgr, size_t bgr_step, const uint16_t *depth, size_t depth_step, bool host_src, PipeFlags *flags)
{
    if (!c->have_prev) { c->err = "detect: call sindyn_set_prev_frames first"; return SINDYN_ERR_STATE; }
    SD_CHECK(pipe_init(c));
    FramePipe *P = c->pipe;
    cudaStream_t user_s = c->stream, main_s = P->sc, s2 = c->stream2;   // main_s: where the state chain runs
    if (P->built_for != user_s) { pipe_drop_graphs(P); P->built_for = user_s; }
    const int k = (int)(P->frame_no % PIPE_NB), p = k & 1, slot = c->i_cur;    // buffers k, solver / fitter instance and streams p
    if (P->fresh) {   // whatever the handle's stream has done so far (state set by the caller, frames of the other path) comes first
        CU_CHECK(c, cudaMemcpyAsync(P->label_out[p ^ 1], c->label_last, c->N, cudaMemcpyDeviceToDevice, user_s));   // the warm start of this frame's k-means
        CU_CHECK(c, cudaEventRecord(P->ev_sync, user_s));
        CU_CHECK(c, cudaStreamWaitEvent(main_s, P->ev_sync, 0));
        CU_CHECK(c, cudaStreamWaitEvent(P->sa[0], P->ev_sync, 0));
        CU_CHECK(c, cudaStreamWaitEvent(P->sa[1], P->ev_sync, 0));
        CU_CHECK(c, cudaStreamWaitEvent(s2, P->ev_sync, 0));
        for (int q = 0; q < PIPE_NB; ++q) CU_CHECK(c, cudaStreamWaitEvent(P->sp[q], P->ev_sync, 0));
        P->fresh = false;
    }
    c->flow_full = P->flow_full[k]; c->fb_flag = P->fb_flag[k]; c->fb_flag_host = P->fb_flag_host[k]; c->depth = P->depth[k]; c->plane_edges = P->plane_edges[k];
    c->edges = P->edges[k];
    c->rc.label_out = P->label_out[p];
    // ---- stream A of this parity: inputs, gray / resize, Brox .. up-sampling.  The ring slot written here held frame i - 4;
    // its readers (the flow solves of frames i - 4 .. i - 2) are done once frame i - 2 is decided
    cudaStream_t sa = P->sa[p];
    CU_CHECK(c, cudaStreamWaitEvent(sa, P->ev_done[k], 0));                 // frame i - PIPE_NB has finished with buffers k
    // the ring slot written below is frame i - 3's "last" image: its large-motion solve / refinement ran on the OTHER part-A
    // stream, nothing else orders it before this frame's input copy
    CU_CHECK(c, cudaStreamWaitEvent(sa, P->ev_a[(k + 1) % PIPE_NB], 0));
    if (P->gray_pending[slot]) CU_CHECK(c, cudaStreamWaitEvent(sa, P->ev_gray[slot], 0));   // ... and the extractor with this ring slot
    if (!host_src) {
        CU_CHECK(c, cudaMemcpyAsync(c->bgr[slot], bgr, (size_t)c->N * 3, cudaMemcpyDeviceToDevice, sa));
        CU_CHECK(c, cudaMemcpyAsync(c->depth, depth, (size_t)c->N * 2, cudaMemcpyDeviceToDevice, sa));
    } else {
        // pinned caller memory is read by the DMA directly; pageable memory goes through this parity's bounce buffers (free
        // again once the copies of frame i - 2 are done)
        const bool pin_b = host_ptr_is_pinned(bgr), pin_d = host_ptr_is_pinned(depth);
        if (!pin_b || !pin_d) {
            if (!P->in_bgr[p]) { SD_CHECK(c->halloc(&P->in_bgr[p], (size_t)c->N * 3)); SD_CHECK(c->halloc(&P->in_depth[p], (size_t)c->N)); }
            CU_CHECK(c, cudaEventSynchronize(P->ev_h2d[p]));
        }
        CU_CHECK(c, stage_in_2d(c->bgr[slot], bgr, bgr_step, (size_t)c->W * 3, c->H, pin_b ? nullptr : P->in_bgr[p], sa));
        CU_CHECK(c, stage_in_2d(c->depth, depth, depth_step, (size_t)c->W * 2, c->H, pin_d ? nullptr : P->in_depth[p], sa));
        CU_CHECK(c, cudaEventRecord(P->ev_h2d[p], sa));
    }
    if (!c->cfg.plane_edges) CU_CHECK(c, cudaMemsetAsync(c->plane_edges, 0, c->N, sa));
    CU_CHECK(c, cudaEventRecord(P->ev_in[k], sa));
    c->stream = sa;
    int st = sindyn_prep_frame(c, slot);
    if (st == SINDYN_OK) st = flow_part_a(c, k);
    c->stream = user_s;
    SD_CHECK(st);
    CU_CHECK(c, cudaEventRecord(P->ev_a[k], sa));
    // ---- this ring position's depth stream: gradient edges, then the plane fitter (both depth only)
    {
        cudaStream_t s3 = P->sp[k];
        CU_CHECK(c, cudaStreamWaitEvent(s3, P->ev_in[k], 0));
        if (!P->g_e[k]) SD_CHECK(pipe_capture(c, s3, &P->g_e[k], &P->n_e[k], [&]() { return edges_run(c, &c->edges, c->depth, c->cfg.depth_scale); }));
        CU_CHECK(c, cudaGraphLaunch(P->g_e[k], s3));
        c->launches += P->n_e[k];
        CU_CHECK(c, cudaEventRecord(P->ev_e[k], s3));
    }
    if (c->cfg.plane_edges) {
        cudaStream_t s3 = P->sp[k];
        if (!P->g_p[k])
            SD_CHECK(pipe_capture(c, s3, &P->g_p[k], &P->n_p[k], [&]() {
                return peac_run(c, &P->peac[k], &P->rc_peac[k], c->depth, c->cfg.fx, c->cfg.fy, c->cfg.cx, c->cfg.cy, c->cfg.depth_scale, c->plane_edges);
            }));
        CU_CHECK(c, cudaGraphLaunch(P->g_p[k], s3));
        c->launches += P->n_p[k];
        CU_CHECK(c, cudaEventRecord(P->ev_p[k], s3));
    }
    // ---- stream 2: k-means (warm-started by the previous frame's labels: it follows that frame's re-clustering in stream order),
    // then the plane-edge filter and the re-clustering.  This parity's label image is free once frame i - 2 is decided.
    CU_CHECK(c, cudaStreamWaitEvent(s2, P->ev_in[k], 0));
    CU_CHECK(c, cudaStreamWaitEvent(s2, P->ev_done[(k + 2) % PIPE_NB], 0));
    if (!P->g_c1[k]) SD_CHECK(pipe_capture(c, s2, &P->g_c1[k], &P->n_c1[k], [&]() { return kmeans_run(c, &c->km, c->depth, P->label_out[p ^ 1], &c->cfg); }));
    CU_CHECK(c, cudaGraphLaunch(P->g_c1[k], s2));
    c->launches += P->n_c1[k];
    CU_CHECK(c, cudaStreamWaitEvent(s2, P->ev_e[k], 0));
    if (c->cfg.plane_edges) CU_CHECK(c, cudaStreamWaitEvent(s2, P->ev_p[k], 0));
    if (!P->g_c2[k]) SD_CHECK(pipe_capture(c, s2, &P->g_c2[k], &P->n_c2[k], [&]() { return cluster_part2(c); }));
    CU_CHECK(c, cudaGraphLaunch(P->g_c2[k], s2));
    c->launches += P->n_c2[k];
    if (flags) CU_CHECK(c, cudaMemcpyAsync(&flags->rc, c->rc.ctl, sizeof(ReclusterControl), cudaMemcpyDeviceToHost, s2));   // before the next frame's re-clustering resets it
    CU_CHECK(c, cudaEventRecord(P->ev_join, s2));
    // ---- the handle's stream: part B, decision, state roll
    CU_CHECK(c, cudaStreamWaitEvent(main_s, P->ev_a[k], 0));
    c->stream = main_s;
    st = flow_part_b(c, k);
    c->stream = user_s;
    SD_CHECK(st);
    CU_CHECK(c, cudaStreamWaitEvent(main_s, P->ev_join, 0));
    if (P->ddout_pending) { CU_CHECK(c, cudaStreamWaitEvent(main_s, P->ev_ddout, 0)); P->ddout_pending = false; }
    if (!P->g_d[k])     // decision + state roll (DynaDetect.cc:1543-1636,1660-1664): ~25 launches and three copies as one graph
        SD_CHECK(pipe_capture(c, main_s, &P->g_d[k], &P->n_d[k], [&]() -> int {
            SD_CHECK(decide_run(c, &c->dd, P->dd_cls, P->dd_labels, P->dd_stats, P->dd_top, c->mask_low, c->mask_high, c->high_last, c->edges.total_area,
                                c->rc.label_out));
            CU_CHECK(c, cudaMemcpyAsync(c->dyna_last, c->dd.out, c->N, cudaMemcpyDeviceToDevice, c->stream));
            CU_CHECK(c, cudaMemcpyAsync(c->high_last, c->mask_high, c->N, cudaMemcpyDeviceToDevice, c->stream));
            CU_CHECK(c, cudaMemcpyAsync(c->label_last, c->rc.label_out, c->N, cudaMemcpyDeviceToDevice, c->stream));
            return (int)SINDYN_OK;
        }));
    CU_CHECK(c, cudaGraphLaunch(P->g_d[k], main_s));
    c->launches += P->n_d[k];
    if (flags) {
        CU_CHECK(c, cudaMemcpyAsync(flags->edge_scalars, c->edges.scalars, sizeof(int) * 4, cudaMemcpyDeviceToHost, main_s));
        flags->peac_hdr[0] = flags->peac_hdr[1] = flags->peac_hdr[2] = flags->peac_hdr[3] = 0;
        if (c->cfg.plane_edges) {
            c->stream = main_s;
            st = peac_copy_sticky_overflow(c, &P->peac[k], &flags->peac_hdr[2]);
            c->stream = user_s;
            SD_CHECK(st);
        }
    }
    CU_CHECK(c, cudaEventRecord(P->ev_done[k], main_s));
    P->last_slot = slot; P->last_k = k;
    ++P->frame_no;
    c->roll_ring();
    return SINDYN_OK;
}
