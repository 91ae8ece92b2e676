// pipeline.cu -- fused per-frame entry points: the flow + residual branch
// (DynaDetect::DetectDynaByDenseOpticalFLow, ORB_SLAM2/src/DynaDetect.cc:1023-1374), device-resident frame
// slots for kernel-only timing, and the Brox measurement hook.
#include <cstdio>
#include <cstdlib>
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

#define STAGE_MARK(c, i)                                                      \
    do {                                                                      \
        if ((c)->cfg.stage_timing) CU_CHECK(c, cudaEventRecord((c)->ev[i], (c)->stream)); \
    } while (0)

// Flow branch + homography + residual on a BGR frame already on the device (bgr_dev may alias c->bgr[i_cur]).
// Results stay resident: c->flow_full, c->homog.H_dev, c->resid.thr, c->mask_low, c->mask_high.
int flow_residual_run(sindyn_ctx *c, const uint8_t *bgr_dev, bool roll)
{
    if (!c->have_prev) { c->err = "flow_residual: call sindyn_set_prev_frames first"; return SINDYN_ERR_STATE; }
    if (c->cfg.stage_timing) SD_CHECK(ensure_events(c));
    SD_CHECK(pipe_join(c));
    pipe_invalidate(c);
    STAGE_MARK(c, 0);
    if (bgr_dev != c->bgr[c->i_cur])
        CU_CHECK(c, cudaMemcpyAsync(c->bgr[c->i_cur], bgr_dev, (size_t)c->N * 3, cudaMemcpyDeviceToDevice, c->stream));
    SD_CHECK(sindyn_prep_frame(c, c->i_cur));
    STAGE_MARK(c, 1);
    int lm = 0;
    SD_CHECK(flow_branch_begin(c, true));      // marks ev[2] (after Brox) itself
    SD_CHECK(flow_finish_all(c, &lm));   // marks ev[3], ev[4], ev[5]
    c->large_motion_last = lm;
    if (roll) {  // imgRGBLastLast <- imgRGBLast <- cur (DynaDetect.cc:1661-1662): index rotation, no copies
        c->roll_ring();
    }
    return SINDYN_OK;
}

static int collect_stage_ms(sindyn_ctx *c, int last_ev)
{
    if (!c->cfg.stage_timing) return SINDYN_OK;
    CU_CHECK(c, cudaEventSynchronize(c->ev[last_ev]));
    for (int i = 0; i < last_ev; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ev[i], c->ev[i + 1]) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
        c->stage_ms[i] = ms;
    }
    float tot = 0.f;
    cudaEventElapsedTime(&tot, c->ev[0], c->ev[last_ev]);
    c->stage_ms[10] = tot;
    return SINDYN_OK;
}

extern "C" int sindyn_upload_frame(sindyn_handle h, int slot, const uint8_t *bgr, size_t bgr_step, const uint16_t *depth, size_t depth_step)
{
    H_CHECK(h);
    if (slot < 0 || slot >= SINDYN_MAX_SLOTS || !bgr) return SINDYN_ERR_INVALID;
    if (!h->slot_bgr[slot]) {
        SD_CHECK(h->dalloc(&h->slot_bgr[slot], (size_t)h->N * 3));
        SD_CHECK(h->dalloc(&h->slot_depth[slot], (size_t)h->N));
    }
    CU_CHECK(h, copy_in_2d(h->slot_bgr[slot], bgr, bgr_step, (size_t)h->W * 3, h->H, h->stream));
    if (depth) CU_CHECK(h, copy_in_2d(h->slot_depth[slot], depth, depth_step, (size_t)h->W * 2, h->H, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    return SINDYN_OK;
}

extern "C" int sindyn_flow_residual(sindyn_handle h, const uint8_t *bgr, size_t bgr_step, uint8_t *mask_low, uint8_t *mask_high, int roll)
{
    H_CHECK(h);
    if (!bgr) return SINDYN_ERR_INVALID;
    CU_CHECK(h, stage_in_2d(h->bgr[h->i_cur], bgr, bgr_step, (size_t)h->W * 3, h->H, h->pin_bgr, h->stream));
    SD_CHECK(flow_residual_run(h, h->bgr[h->i_cur], roll != 0));
    const bool lo_direct = mask_low && host_ptr_is_pinned(mask_low), hi_direct = mask_high && host_ptr_is_pinned(mask_high);
    if (mask_low) CU_CHECK(h, stage_out_begin(lo_direct ? (void *)mask_low : (void *)h->pin_out0, h->mask_low, h->N, h->stream));
    if (mask_high) CU_CHECK(h, stage_out_begin(hi_direct ? (void *)mask_high : (void *)h->pin_out1, h->mask_high, h->N, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    flow_collect_flag(h);
    if (mask_low && !lo_direct) stage_out_finish(mask_low, 0, h->pin_out0, h->N, 1);
    if (mask_high && !hi_direct) stage_out_finish(mask_high, 0, h->pin_out1, h->N, 1);
    return collect_stage_ms(h, 5);
}

extern "C" int sindyn_flow_residual_resident(sindyn_handle h, int slot, int roll)
{
    H_CHECK(h);
    if (slot < 0 || slot >= SINDYN_MAX_SLOTS || !h->slot_bgr[slot]) { h->err = "flow_residual_resident: empty slot"; return SINDYN_ERR_INVALID; }
    return flow_residual_run(h, h->slot_bgr[slot], roll != 0);
}

extern "C" int sindyn_get_flow_results(sindyn_handle h, float *flow, double *H_out, float *thresholds, uint8_t *mask_low, uint8_t *mask_high,
                                       int *large_motion)
{
    H_CHECK(h);
    if (flow) CU_CHECK(h, cudaMemcpyAsync(flow, h->flow_full, sizeof(float) * 2 * h->N, cudaMemcpyDeviceToHost, h->stream));
    if (H_out) CU_CHECK(h, cudaMemcpyAsync(H_out, h->homog.H_dev, sizeof(double) * 9, cudaMemcpyDeviceToHost, h->stream));
    if (thresholds) CU_CHECK(h, cudaMemcpyAsync(thresholds, h->resid.thr, sizeof(float) * 4, cudaMemcpyDeviceToHost, h->stream));
    if (mask_low) CU_CHECK(h, cudaMemcpyAsync(mask_low, h->mask_low, h->N, cudaMemcpyDeviceToHost, h->stream));
    if (mask_high) CU_CHECK(h, cudaMemcpyAsync(mask_high, h->mask_high, h->N, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    flow_collect_flag(h);
    if (large_motion) *large_motion = h->large_motion_last;
    if (h->cfg.stage_timing) collect_stage_ms(h, 5);
    return SINDYN_OK;
}

extern "C" int sindyn_get_path_info(sindyn_handle h, int info[4])
{
    H_CHECK(h);
    if (!info) return SINDYN_ERR_INVALID;
    info[0] = h->flow_graph_active ? 1 : 0;
    info[1] = h->flow_graph_broken ? 1 : 0;
    info[2] = (h->cluster_graph != nullptr) ? 1 : 0;
    info[3] = 0;
    return SINDYN_OK;
}

extern "C" int sindyn_brox_profile(sindyn_handle h, double *out4)
{
    H_CHECK(h);
    if (!out4) return SINDYN_ERR_INVALID;
    if (!h->have_prev) { h->err = "brox_profile: call sindyn_set_prev_frames first"; return SINDYN_ERR_STATE; }
    BroxSolver *b = &h->brox;
    const int cap = 2 * (b->nl * b->inner * b->solver + 4);   // upper bound: one launch per sweep
    std::vector<cudaEvent_t> ev(cap + 2);
    for (auto &e : ev) CU_CHECK(h, cudaEventCreate(&e));
    b->prof_ev = ev.data(); b->prof_n = 0; b->prof_cap = cap; b->prof_px = 0;
    CU_CHECK(h, cudaEventRecord(ev[cap], h->stream));
    int st = brox_run(h, b, h->gsmall_f[h->i_last], h->gsmall_f[h->i_lastlast], h->flow_small, -1.0f, false);
    cudaEventRecord(ev[cap + 1], h->stream);
    const int n = b->prof_n;
    const long long px = b->prof_px;
    b->prof_ev = nullptr; b->prof_n = 0; b->prof_cap = 0;
    if (st == SINDYN_OK && cudaStreamSynchronize(h->stream) != cudaSuccess) st = SINDYN_ERR_CUDA;
    double sor = 0.0;
    float ms = 0.f;
    if (st == SINDYN_OK) {
        const bool trace = getenv("SINDYN_BROX_TRACE") != nullptr;   // developer aid: per-interval times on stderr
        for (int i = 0; i + 1 < n; i += 2) {
            cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            sor += ms;
            if (trace) fprintf(stderr, "brox interval %d: %.1f us\n", i / 2, 1e3 * ms);
        }
        cudaEventElapsedTime(&ms, ev[cap], ev[cap + 1]);
    }
    for (auto &e : ev) cudaEventDestroy(e);
    SD_CHECK(st);
    out4[0] = sor; out4[1] = n / 2; out4[2] = ms; out4[3] = (double)px;
    return SINDYN_OK;
}
