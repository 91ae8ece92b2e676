// detect.cu -- DynaDetect::DetectDynaArea (ORB_SLAM2/src/DynaDetect.cc:1377-1666) end to end, and the stage-level
// entry points of the clustering / decision half (plane edges, plane-edge filter, re-clustering, decision).
#include "ctx.cuh"

#define H_CHECK(h)                       \
    if (!(h)) return SINDYN_ERR_INVALID; \
    cudaSetDevice((h)->device)

static int ensure_events(sindyn_ctx *c)
{
    if (c->ev_ok) return SINDYN_OK;
    for (auto &e : c->ev) CU_CHECK(c, cudaEventCreate(&e));
    c->ev_ok = true;
    return SINDYN_OK;
}

#define MARK(c, i)                                                                        \
    do {                                                                                  \
        if ((c)->cfg.stage_timing) CU_CHECK(c, cudaEventRecord((c)->ev[i], (c)->stream)); \
    } while (0)

// clustering branch (DynaDetect.cc:1410-1516) on the handle's current stream; depth already in c->depth
static int cluster_branch(sindyn_ctx *c)
{
    // the plane fitter (a long serial chain on one SM) depends on the depth image only: it runs on its own stream next to
    // k-means and the gradient edges and joins before the plane-edge filter
    cudaStream_t s2 = c->stream;
    if (c->cfg.plane_edges) {
        CU_CHECK(c, cudaEventRecord(c->ev_peac_fork, s2));
        CU_CHECK(c, cudaStreamWaitEvent(c->stream3, c->ev_peac_fork, 0));
        c->stream = c->stream3;
        if (c->cfg.stage_timing) cudaEventRecord(c->ev[15], c->stream3);
        int st = peac_run(c, &c->peac, &c->rc, c->depth, c->cfg.fx, c->cfg.fy, c->cfg.cx, c->cfg.cy, c->cfg.depth_scale, c->plane_edges);
        if (c->cfg.stage_timing) cudaEventRecord(c->ev[16], c->stream3);
        cudaEventRecord(c->ev_peac_join, c->stream3);
        c->stream = s2;
        SD_CHECK(st);
    } else {
        CU_CHECK(c, cudaMemsetAsync(c->plane_edges, 0, c->N, c->stream));
    }
    MARK(c, 10);
    SD_CHECK(kmeans_run(c, &c->km, c->depth, c->label_last, &c->cfg));
    MARK(c, 11);
    SD_CHECK(edges_run(c, &c->edges, c->depth, c->cfg.depth_scale));
    MARK(c, 12);
    if (c->cfg.plane_edges) CU_CHECK(c, cudaStreamWaitEvent(s2, c->ev_peac_join, 0));
    SD_CHECK(plane_edge_filter_run(c, &c->rc, c->plane_edges, c->edges.grad_edges, c->edges.ep_xy, c->edges.scalars + 2));
    MARK(c, 13);
    SD_CHECK(recluster_run(c, &c->rc, &c->km, c->rc.occl1, c->rc.occl2, c->depth));
    MARK(c, 14);
    return SINDYN_OK;
}

// the two halves of the clustering branch without the plane fitter and the gradient edges (frame pipeline: both run ahead on
// their own stream)
int cluster_part1(sindyn_ctx *c)      // (the gradient edges depend on the depth image only: they run ahead, see pipe.cu)
{
    return kmeans_run(c, &c->km, c->depth, c->label_last, &c->cfg);
}
int cluster_part2(sindyn_ctx *c)
{
    SD_CHECK(plane_edge_filter_run(c, &c->rc, c->plane_edges, c->edges.grad_edges, c->edges.ep_xy, c->edges.scalars + 2));
    return recluster_run(c, &c->rc, &c->km, c->rc.occl1, c->rc.occl2, c->depth);
}

static int check_capacity(sindyn_ctx *c)
{
    ReclusterControl ctl;
    int sc[4];
    SD_CHECK(pipe_join(c));      // frames of the frame pipeline finish first (their decision runs on the pipeline's own stream)
    CU_CHECK(c, cudaMemcpyAsync(&ctl, c->rc.ctl, sizeof ctl, cudaMemcpyDeviceToHost, c->stream));
    CU_CHECK(c, cudaMemcpyAsync(sc, c->edges.scalars, sizeof sc, cudaMemcpyDeviceToHost, c->stream));
    int peac_hdr[4] = {0, 0, 0, 0};
    if (c->cfg.plane_edges && c->peac.built) SD_CHECK(peac_copy_header(c, &c->peac, peac_hdr));
    SD_CHECK(pipe_copy_headers(c));      // plane-fitter instances of the frame pipeline (asynchronous, on c->stream)
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    flow_collect_flag(c);
    if (pipe_overflow(c)) peac_hdr[2] = 1;
    if (peac_hdr[2]) { c->err = "plane fitter: a fixed-capacity list overflowed (> 64 planes, > 512 neighbours of one node, or a region-growing level > 131072 entries)"; return SINDYN_ERR_CAPACITY; }
    if (ctl.overflow) { c->err = "recluster: more than RC_MAXC components"; return SINDYN_ERR_CAPACITY; }
    if (ctl.pf_overflow) { c->err = "plane-edge filter: more than RC_PF_MAXC contours"; return SINDYN_ERR_CAPACITY; }
    if (sc[3]) { c->err = "depth_edges: more than EDGE_EP_CAP candidate end points"; return SINDYN_ERR_CAPACITY; }
    return SINDYN_OK;
}

// One frame, inputs already on the device (c->bgr[i_cur], c->depth).  The clustering branch runs on stream2 while the
// flow branch (which contains the one host decision, large motion) runs on the main stream -- the reference's thread split
// (DynaDetect.cc:1396-1398,1553).  Results: c->dd.out (mask), c->rc.label_out (labels); state rolled.
static int detect_run(sindyn_ctx *c)
{
    if (!c->have_prev) { c->err = "detect: call sindyn_set_prev_frames first"; return SINDYN_ERR_STATE; }
    if (c->cfg.stage_timing) SD_CHECK(ensure_events(c));
    SD_CHECK(pipe_join(c));          // frames still in the pipeline finish first; its streams re-synchronise on their next frame
    pipe_invalidate(c);
    cudaStream_t main_s = c->stream;
    MARK(c, 0);
    CU_CHECK(c, cudaEventRecord(c->ev_fork, main_s));
    // ---- flow branch, first part: the Brox solve is one graph launch, so the GPU starts on the critical path at once
    SD_CHECK(sindyn_prep_frame(c, c->i_cur));
    MARK(c, 1);
    SD_CHECK(flow_branch_begin(c, true));      // marks ev[2] after the first Brox solve
    // ---- clustering branch on stream2 (+ stream3 for the plane fitter): ~300 launches without a host decision.  With
    // use_graphs it is captured once (cross-stream fork / join included) and replayed as ONE graph launch per frame.
    CU_CHECK(c, cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
    const bool graph = c->cfg.use_graphs && !c->cfg.stage_timing;
    if (graph && c->cluster_graph && c->cluster_graph_stream == c->stream2) {
        CU_CHECK(c, cudaGraphLaunch(c->cluster_graph, c->stream2));
        c->launches += c->cluster_graph_launches;
    } else {
        c->stream = c->stream2;
        int st = SINDYN_OK;
        cudaGraph_t gr = nullptr;
        const unsigned long long before = c->launches;
        if (graph) {
            if (c->cluster_graph) { cudaGraphExecDestroy(c->cluster_graph); c->cluster_graph = nullptr; }
            if (cudaStreamBeginCapture(c->stream2, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { c->stream = main_s; c->err = "cluster graph capture failed to start"; return SINDYN_ERR_CUDA; }
        }
        st = cluster_branch(c);
        if (graph) {
            cudaError_t e = cudaStreamEndCapture(c->stream2, &gr);
            c->stream = main_s;
            SD_CHECK(st);
            CU_CHECK(c, e);
            CU_CHECK(c, cudaGraphInstantiate(&c->cluster_graph, gr, 0));
            cudaGraphDestroy(gr);
            c->cluster_graph_stream = c->stream2;
            c->cluster_graph_launches = c->launches - before;
            CU_CHECK(c, cudaGraphLaunch(c->cluster_graph, c->stream2));
        } else {
            c->stream = main_s;
            SD_CHECK(st);
        }
    }
    CU_CHECK(c, cudaEventRecord(c->ev_join, c->stream2));
    // ---- flow branch, second part (contains the one host decision, large motion)
    int lm = 0;
    SD_CHECK(flow_finish_all(c, &lm));   // marks ev[3], ev[4], ev[5]
    c->large_motion_last = lm;
    // ---- join + decision (DynaDetect.cc:1543-1636)
    CU_CHECK(c, cudaStreamWaitEvent(main_s, c->ev_join, 0));
    MARK(c, 6);
    SD_CHECK(decide_run(c, &c->dd, c->rc.cls, c->rc.labels, c->rc.stats, c->rc.top, c->mask_low, c->mask_high, c->high_last,
                        c->edges.total_area, c->rc.label_out));
    // ---- state roll (DynaDetect.cc:1660-1664)
    CU_CHECK(c, cudaMemcpyAsync(c->dyna_last, c->dd.out, c->N, cudaMemcpyDeviceToDevice, main_s));
    CU_CHECK(c, cudaMemcpyAsync(c->high_last, c->mask_high, c->N, cudaMemcpyDeviceToDevice, main_s));
    CU_CHECK(c, cudaMemcpyAsync(c->label_last, c->rc.label_out, c->N, cudaMemcpyDeviceToDevice, main_s));
    MARK(c, 7);
    c->roll_ring();
    return SINDYN_OK;
}

int detect_run_public(sindyn_ctx *c) { return detect_run(c); }
int detect_check_capacity(sindyn_ctx *c) { return check_capacity(c); }

static int collect_detect_ms(sindyn_ctx *c)
{
    if (!c->cfg.stage_timing) return SINDYN_OK;
    CU_CHECK(c, cudaEventSynchronize(c->ev[7]));
    auto el = [&](int a, int b) { float ms = 0.f; if (cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]) != cudaSuccess) { cudaGetLastError(); ms = 0.f; } return ms; };
    c->stage_ms[0] = el(0, 1); c->stage_ms[1] = el(1, 2); c->stage_ms[2] = el(2, 3); c->stage_ms[3] = el(3, 4); c->stage_ms[4] = el(4, 5);
    c->stage_ms[5] = el(10, 11); c->stage_ms[6] = el(11, 12); c->stage_ms[7] = el(12, 13); c->stage_ms[8] = el(13, 14);
    c->stage_ms[9] = el(6, 7); c->stage_ms[10] = el(0, 7);
    c->stage_ms[11] = c->cfg.plane_edges ? el(15, 16) : 0.f;   // PEAC plane edges (own stream, concurrent with k-means / gradient edges)
    return SINDYN_OK;
}

extern "C" int sindyn_detect(sindyn_handle h, const uint8_t *bgr, size_t bgr_step, const uint16_t *depth, size_t depth_step,
                             uint8_t *mask_out, size_t mask_step, uint8_t *label_out, size_t label_step, int frame_idx)
{
    (void)frame_idx;   // nImg only divides the reference's running time sums (DynaDetect.cc:1647)
    H_CHECK(h);
    if (!bgr || !depth) return SINDYN_ERR_INVALID;
    CU_CHECK(h, stage_in_2d(h->bgr[h->i_cur], bgr, bgr_step, (size_t)h->W * 3, h->H, h->pin_bgr, h->stream));
    CU_CHECK(h, stage_in_2d(h->depth, depth, depth_step, (size_t)h->W * 2, h->H, h->pin_depth, h->stream));
    SD_CHECK(detect_run(h));
    if (mask_out) CU_CHECK(h, stage_out_begin(h->pin_out0, h->dd.out, h->N, h->stream));
    if (label_out) CU_CHECK(h, stage_out_begin(h->pin_out1, h->rc.label_out, h->N, h->stream));
    SD_CHECK(check_capacity(h));   // synchronises the stream
    if (mask_out) stage_out_finish(mask_out, mask_step, h->pin_out0, h->W, h->H);
    if (label_out) stage_out_finish(label_out, label_step, h->pin_out1, h->W, h->H);
    return collect_detect_ms(h);
}

extern "C" int sindyn_detect_resident(sindyn_handle h, int slot, int frame_idx)
{
    (void)frame_idx;
    H_CHECK(h);
    if (slot < 0 || slot >= SINDYN_MAX_SLOTS || !h->slot_bgr[slot]) { h->err = "detect_resident: empty slot"; return SINDYN_ERR_INVALID; }
    CU_CHECK(h, cudaMemcpyAsync(h->bgr[h->i_cur], h->slot_bgr[slot], (size_t)h->N * 3, cudaMemcpyDeviceToDevice, h->stream));
    CU_CHECK(h, cudaMemcpyAsync(h->depth, h->slot_depth[slot], (size_t)h->N * 2, cudaMemcpyDeviceToDevice, h->stream));
    return detect_run(h);
}

extern "C" int sindyn_get_detect_results(sindyn_handle h, uint8_t *mask, uint8_t *labels)
{
    H_CHECK(h);
    if (mask) CU_CHECK(h, cudaMemcpyAsync(mask, h->dd.out, h->N, cudaMemcpyDeviceToHost, h->stream));
    if (labels) CU_CHECK(h, cudaMemcpyAsync(labels, h->rc.label_out, h->N, cudaMemcpyDeviceToHost, h->stream));
    SD_CHECK(check_capacity(h));
    return collect_detect_ms(h);
}

// ------------------------------------------------------------------ stage-level entry points
extern "C" int sindyn_plane_edges(sindyn_handle h, const uint16_t *depth, size_t depth_step, uint8_t *plane_edges_out)
{
    H_CHECK(h);
    if (!depth || !plane_edges_out) return SINDYN_ERR_INVALID;
    CU_CHECK(h, copy_in_2d(h->depth, depth, depth_step, (size_t)h->W * 2, h->H, h->stream));
    SD_CHECK(peac_run(h, &h->peac, &h->rc, h->depth, h->cfg.fx, h->cfg.fy, h->cfg.cx, h->cfg.cy, h->cfg.depth_scale, h->plane_edges));
    CU_CHECK(h, cudaMemcpyAsync(plane_edges_out, h->plane_edges, h->N, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    return SINDYN_OK;
}

extern "C" int sindyn_filter_plane_edges(sindyn_handle h, const uint8_t *plane_edges, const uint8_t *grad_edges, const int *endpoints,
                                         int n_endpoints, uint8_t *occluded1_out, uint8_t *occluded2_out)
{
    H_CHECK(h);
    if (!plane_edges || !grad_edges || n_endpoints < 0 || n_endpoints > EDGE_EP_CAP || (n_endpoints && !endpoints)) return SINDYN_ERR_INVALID;
    CU_CHECK(h, cudaMemcpyAsync(h->plane_edges, plane_edges, h->N, cudaMemcpyHostToDevice, h->stream));
    CU_CHECK(h, cudaMemcpyAsync(h->edges.grad_edges, grad_edges, h->N, cudaMemcpyHostToDevice, h->stream));
    if (n_endpoints) CU_CHECK(h, cudaMemcpyAsync(h->edges.ep_xy, endpoints, sizeof(int) * 2 * n_endpoints, cudaMemcpyHostToDevice, h->stream));
    CU_CHECK(h, cudaMemcpyAsync(h->edges.scalars + 2, &n_endpoints, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    SD_CHECK(plane_edge_filter_run(h, &h->rc, h->plane_edges, h->edges.grad_edges, h->edges.ep_xy, h->edges.scalars + 2));
    if (occluded1_out) CU_CHECK(h, cudaMemcpyAsync(occluded1_out, h->rc.occl1, h->N, cudaMemcpyDeviceToHost, h->stream));
    if (occluded2_out) CU_CHECK(h, cudaMemcpyAsync(occluded2_out, h->rc.occl2, h->N, cudaMemcpyDeviceToHost, h->stream));
    ReclusterControl ctl;
    CU_CHECK(h, cudaMemcpyAsync(&ctl, h->rc.ctl, sizeof ctl, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    if (ctl.pf_overflow) { h->err = "plane-edge filter: more than RC_PF_MAXC contours"; return SINDYN_ERR_CAPACITY; }
    return SINDYN_OK;
}

extern "C" int sindyn_recluster(sindyn_handle h, const uint8_t *occluded1, const uint8_t *occluded2, const uint16_t *depth, size_t depth_step,
                                uint8_t *label_out, int *n_components_out)
{
    H_CHECK(h);
    if (!occluded1 || !occluded2 || !depth) return SINDYN_ERR_INVALID;
    CU_CHECK(h, cudaMemcpyAsync(h->rc.occl1, occluded1, h->N, cudaMemcpyHostToDevice, h->stream));
    CU_CHECK(h, cudaMemcpyAsync(h->rc.occl2, occluded2, h->N, cudaMemcpyHostToDevice, h->stream));
    CU_CHECK(h, copy_in_2d(h->depth, depth, depth_step, (size_t)h->W * 2, h->H, h->stream));
    SD_CHECK(recluster_run(h, &h->rc, &h->km, h->rc.occl1, h->rc.occl2, h->depth));
    if (label_out) CU_CHECK(h, cudaMemcpyAsync(label_out, h->rc.label_out, h->N, cudaMemcpyDeviceToHost, h->stream));
    ReclusterControl ctl;
    CU_CHECK(h, cudaMemcpyAsync(&ctl, h->rc.ctl, sizeof ctl, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    if (n_components_out) *n_components_out = ctl.n_comp;
    if (ctl.overflow) { h->err = "recluster: more than RC_MAXC components"; return SINDYN_ERR_CAPACITY; }
    return SINDYN_OK;
}

// debug / test hook: the RAG matrix ((n+1)^2 floats, rank space) and per-component scalars of the last recluster run
extern "C" int sindyn_get_recluster_debug(sindyn_handle h, float *T_out, int t_capacity, int *area_out, float *score_out, int *order_out)
{
    H_CHECK(h);
    ReclusterControl ctl;
    CU_CHECK(h, cudaMemcpyAsync(&ctl, h->rc.ctl, sizeof ctl, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    const int n = ctl.n_comp, s = n + 1;
    if (T_out) {
        if (t_capacity < s * s) return SINDYN_ERR_CAPACITY;
        CU_CHECK(h, cudaMemcpy(T_out, h->rc.Tmat, sizeof(float) * s * s, cudaMemcpyDeviceToHost));
    }
    for (int i = 0; i < n; ++i) {
        if (area_out) area_out[i] = ctl.area[i];
        if (score_out) score_out[i] = ctl.score[i];
        if (order_out) order_out[i] = ctl.order[i];
    }
    return n;
}

extern "C" int sindyn_dynamic_decide(sindyn_handle h, const uint8_t *mask_low, const uint8_t *mask_high, const uint8_t *total_area,
                                     const uint8_t *labels, uint8_t *dyna_out)
{
    H_CHECK(h);
    if (!mask_low || !mask_high || !total_area || !labels) return SINDYN_ERR_INVALID;
    CU_CHECK(h, cudaMemcpyAsync(h->mask_low, mask_low, h->N, cudaMemcpyHostToDevice, h->stream));
    CU_CHECK(h, cudaMemcpyAsync(h->mask_high, mask_high, h->N, cudaMemcpyHostToDevice, h->stream));
    CU_CHECK(h, cudaMemcpyAsync(h->edges.total_area, total_area, h->N, cudaMemcpyHostToDevice, h->stream));
    CU_CHECK(h, cudaMemcpyAsync(h->rc.label_out, labels, h->N, cudaMemcpyHostToDevice, h->stream));
    SD_CHECK(decide_run(h, &h->dd, h->rc.cls, h->rc.labels, h->rc.stats, h->rc.top, h->mask_low, h->mask_high, h->high_last,
                        h->edges.total_area, h->rc.label_out));
    // state roll of the decision inputs (DynaDetect.cc:1660,1663-1664)
    CU_CHECK(h, cudaMemcpyAsync(h->dyna_last, h->dd.out, h->N, cudaMemcpyDeviceToDevice, h->stream));
    CU_CHECK(h, cudaMemcpyAsync(h->high_last, h->mask_high, h->N, cudaMemcpyDeviceToDevice, h->stream));
    CU_CHECK(h, cudaMemcpyAsync(h->label_last, h->rc.label_out, h->N, cudaMemcpyDeviceToDevice, h->stream));
    if (dyna_out) CU_CHECK(h, cudaMemcpyAsync(dyna_out, h->dd.out, h->N, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    return SINDYN_OK;
}

// test hook for the PEAC stage: final plane membership image and the extracted planes of the last sindyn_plane_edges / detect
extern "C" int sindyn_get_peac_debug(sindyn_handle h, int *membership_out, int *planes_rid_n_final, int *n_planes_out, int *n_final_out)
{
    H_CHECK(h);
    return peac_get_debug(h, &h->peac, membership_out, planes_rid_n_final, n_planes_out, n_final_out);
}
