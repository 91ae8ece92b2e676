// ctx.cuh -- the DynaDetect handle: configuration, device-resident frames and inter-frame state
// (ORB_SLAM2/include/DynaDetect.h:164-188), and the per-stage solver objects.
#pragma once
#include "brox.cuh"
#include "common.cuh"
#include "decide.cuh"
#include "edges.cuh"
#include "homography.cuh"
#include "kmeans.cuh"
#include "morph.cuh"
#include "peac.cuh"
#include "preproc.cuh"
#include "recluster.cuh"
#include "residual.cuh"
#include "varref.cuh"

struct sindyn_ctx : sindyn_base {
    sindyn_config cfg;
    int W = 0, H = 0, N = 0;
    int fw = 0, fh = 0;  // flow grid: (int)(flow_scale*W) x (int)(flow_scale*H)  (DynaDetect.cc:1037)
    bool have_prev = false;

    // frames: BGR (W*H*3), gray full, gray small (u8 + float/255)
    // ring of four: cur / last / lastlast by index rotation + one spare slot, the next frame's cur (with two frames' flow
    // solves in flight in the frame pipeline, the next frame is written while lastlast is still being read)
    uint8_t *bgr[4] = {nullptr, nullptr, nullptr, nullptr};
    uint8_t *gray[4] = {nullptr, nullptr, nullptr, nullptr};
    uint8_t *gsmall[4] = {nullptr, nullptr, nullptr, nullptr};
    float *gsmall_f[4] = {nullptr, nullptr, nullptr, nullptr};
    int i_cur = 0, i_last = 1, i_lastlast = 2, i_spare = 3;
    void roll_ring() { const int t = i_spare; i_spare = i_lastlast; i_lastlast = i_last; i_last = i_cur; i_cur = t; }   // DynaDetect.cc:1661-1662
    uint16_t *depth = nullptr;
    // pinned bounce buffers of the per-frame entry points (see stage_in_2d)
    uint8_t *pin_bgr = nullptr, *pin_out0 = nullptr, *pin_out1 = nullptr;
    uint16_t *pin_depth = nullptr;
    // device-resident staging slots for kernel-only timing
    uint8_t *slot_bgr[SINDYN_MAX_SLOTS] = {};
    uint16_t *slot_depth[SINDYN_MAX_SLOTS] = {};

    // inter-frame state (DynaDetect.h:172-178)
    uint8_t *dyna_last = nullptr, *high_last = nullptr, *label_last = nullptr;

    // flow branch
    ResizePlanU8 plan_flow;
    BroxSolver brox, brox_lm;  // (cur, lastlast) and the large-motion (cur, last) solver
    float *fb_mag = nullptr;
    unsigned int *fb_hist = nullptr;
    int *fb_flag = nullptr, *fb_flag_host = nullptr;
    HomographyStage homog;
    VarRefStage varref;
    float *flow_small = nullptr;  // fw x fh x 2
    float *flow_full = nullptr;   // W x H x 2
    // generic scratch for stage-level entry points
    float *scratch_f0 = nullptr, *scratch_f1 = nullptr;  // N*2 floats each
    uint8_t *scratch_u0 = nullptr, *scratch_u1 = nullptr, *scratch_u2 = nullptr, *scratch_u3 = nullptr;  // N*3 bytes each

    ResidualStage resid;
    uint8_t *mask_low = nullptr, *mask_high = nullptr;

    KmeansStage km;
    EdgeStage edges;
    PeacStage peac;
    ReclusterStage rc;
    DecideStage dd;
    uint8_t *plane_edges = nullptr;   // imgEdgeByPlane (zeros when cfg.plane_edges == 0)
    cudaStream_t stream2 = nullptr;   // clustering branch (the reference runs the flow branch in its own std::thread)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_flag = nullptr;    // large-motion flag has arrived on the host
    struct TailGraph { cudaGraphExec_t exec = nullptr; const uint8_t *k0 = nullptr, *k1 = nullptr; cudaStream_t stream = nullptr; unsigned long long launches = 0; };
    TailGraph tail[8];                // captured post-decision part of the flow branch, per frame-ring position
    int tail_next = 0;
    struct FlowGraph { cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr; const uint8_t *k0 = nullptr, *k1 = nullptr, *k2 = nullptr; cudaStream_t stream = nullptr; unsigned long long launches = 0, body_launches = 0; int variant = 0; };
    FlowGraph flow_graph[12];         // the whole flow branch incl. the device-side large-motion decision (flow.cu), per ring position (x parity in the frame pipeline)
    FlowGraph *flow_graph_last = nullptr;
    int flow_graph_next = 0;
    bool flow_one_graph = true, flow_graph_broken = false, flow_graph_active = false, flow_flag_pending = false;
    cudaGraphExec_t cluster_graph = nullptr;   // captured clustering branch (three streams, no host decisions)
    cudaStream_t cluster_graph_stream = nullptr;
    unsigned long long cluster_graph_launches = 0;
    cudaStream_t stream3 = nullptr;   // PEAC plane fitter, concurrent with k-means / gradient edges
    cudaEvent_t ev_peac_fork = nullptr, ev_peac_join = nullptr;

    struct FramePipe *pipe = nullptr;     // software pipeline over consecutive frames (pipe.cu), allocated on first use

    struct CloudStage *cloud = nullptr;   // dense-map consumer stage (cloud.cu), allocated on first use

    float stage_ms[16] = {};
    cudaEvent_t ev[24] = {};
    bool ev_ok = false;
    int large_motion_last = 0;
};

// the resources one flow solve (part A of the branch) works in: the handle's own, or one of the frame pipeline's two sets
struct FlowRes {
    BroxSolver *brox, *brox_lm;
    VarRefStage *varref;
    float *fb_mag;
    unsigned int *fb_hist;
    int *fb_flag, *fb_flag_host;
    float *flow_small, *flow_full;
};
FlowRes pipe_flow_res(sindyn_ctx *c, int parity);          // pipe.cu

// internal cross-module helpers (C++ linkage)
int sindyn_prep_frame(sindyn_ctx *c, int idx);          // api.cu: BGR -> gray -> 0.6x gray (u8 + float)
int flow_branch_init(sindyn_ctx *c);                    // flow.cu
int flow_branch_run(sindyn_ctx *c, int *large_motion);  // flow.cu
int flow_branch_begin(sindyn_ctx *c, bool whole_frame);  // flow.cu
int flow_branch_finish(sindyn_ctx *c, int *large_motion);  // flow.cu
int flow_finish_all(sindyn_ctx *c, int *large_motion);     // flow.cu: finish + homography + residual/masks (marks ev[3..5])
void flow_tail_drop_graphs(sindyn_ctx *c);                 // flow.cu
void flow_graph_drop(sindyn_ctx *c);                       // flow.cu
void flow_collect_flag(sindyn_ctx *c);                     // flow.cu: large-motion flag of the last flow-graph launch (after a sync)
// flow.cu, frame pipeline: part A = Brox .. up-sampling into c->flow_full (on c->stream, one graph launch; variant 1 + parity keys
// the graph cache), part B = sample weighting + homography + residual / masks from c->flow_full
int flow_part_a(sindyn_ctx *c, int parity);
int flow_part_b(sindyn_ctx *c, int parity);
int pipe_detect_run(sindyn_ctx *c, const uint8_t *bgr_dev, const uint16_t *depth_dev);   // pipe.cu: one frame through the frame pipeline
struct PipeFlags { ReclusterControl rc; int edge_scalars[4]; int peac_hdr[4]; };   // capacity flags of one pipelined frame (pinned host memory)
int pipe_detect_run_src(sindyn_ctx *c, const uint8_t *bgr, size_t bgr_step, const uint16_t *depth, size_t depth_step, bool host_src, PipeFlags *flags);
int pipe_join(sindyn_ctx *c);                              // pipe.cu: the handle's stream waits for everything the pipeline has enqueued
void pipe_invalidate(sindyn_ctx *c);                       // pipe.cu: state was changed outside the pipeline: re-synchronise its streams
void pipe_destroy(sindyn_ctx *c);                          // pipe.cu
bool pipe_usable(const sindyn_ctx *c);                     // pipe.cu: graphs on, no stage timing
int pipe_copy_headers(sindyn_ctx *c);                      // pipe.cu: asynchronous copy of the plane-fitter headers of both pipeline instances
bool pipe_overflow(const sindyn_ctx *c);                   // pipe.cu: ... and their overflow flags, valid after the stream was synchronised
cudaStream_t pipe_chain_stream(sindyn_ctx *c);             // pipe.cu: the stream the state chain (part B, decision, state roll) runs on
cudaEvent_t pipe_done_event(sindyn_ctx *c);                // pipe.cu: the last frame is decided (final mask, labels, rolled state)
int pipe_note_mask_read(sindyn_ctx *c, cudaStream_t s);    // pipe.cu: stream s has read the last frame's final mask
cudaEvent_t pipe_input_event(sindyn_ctx *c);               // pipe.cu: the last frame's inputs are in place (bgr ring slot, depth)
int cluster_part1(sindyn_ctx *c);                          // detect.cu: k-means
int cluster_part2(sindyn_ctx *c);                          // detect.cu: plane-edge filter + re-clustering
int flow_residual_run(sindyn_ctx *c, const uint8_t *bgr_dev, bool roll);  // pipeline.cu
int sindyn_ctx_init_stages(sindyn_ctx *c);              // stages.cu
void sindyn_ctx_destroy_stages(sindyn_ctx *c);          // stages.cu
void cloud_stage_destroy(sindyn_ctx *c);                // cloud.cu
