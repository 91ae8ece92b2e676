// api.cu -- C-ABI entry points of libsindyn_cuda (include/sindyn.h): life cycle and the
// flow-branch stages.  Each function cites the reference interface it replaces in sindyn.h.
#include "ctx.cuh"

#include <new>

#define H_CHECK(h)                  \
    if (!(h)) return SINDYN_ERR_INVALID; \
    cudaSetDevice((h)->device)

extern "C" const char *sindyn_version(void) { return "sindyn-cuda 0.1 (sm_100a)"; }

extern "C" void sindyn_default_config(sindyn_config *c, int width, int height)
{
    memset(c, 0, sizeof *c);
    c->width = width; c->height = height;
    c->fx = 535.4f; c->fy = 539.2f; c->cx = 320.1f; c->cy = 247.6f;  // Examples/RGB-D/TUM3.yaml
    c->depth_scale = 5000.0f;
    c->flow_scale = 0.6f;
    c->brox_alpha = 0.197f; c->brox_gamma = 50.0f; c->brox_pyr_scale = 0.8f;
    c->brox_inner = 10; c->brox_outer = 77; c->brox_solver = 10;
    c->brox_omega = 1.99f;
    c->refine = 1;   // cv::VariationalRefinement pass (DynaDetect.cc:1133-1143)
    c->n_row_cluster = 3; c->n_col_cluster = 4;
    c->depth_weight = 1.5f;
    c->device = 0;
    c->use_graphs = 1;
    c->plane_edges = 1;   // PEAC plane-contour edges (DynaDetect.cc:592-593)
}


extern "C" int sindyn_create(const sindyn_config *cfg, sindyn_handle *out)
{
    if (!cfg || !out || cfg->width < 64 || cfg->height < 64) return SINDYN_ERR_INVALID;
    // The number of k-means clusters is a compile-time constant of the kernels (numCluster = nRowCluster * nColCluster = 12,
    // DynaDetect.cc:46-47); the remaining parameters must leave a usable flow grid and a finite solver.
    if (cfg->n_row_cluster * cfg->n_col_cluster != 12 || cfg->n_row_cluster < 1 || cfg->n_col_cluster < 1) return SINDYN_ERR_INVALID;
    if (!(cfg->flow_scale > 0.0f && cfg->flow_scale <= 1.0f) || (int)(cfg->flow_scale * (float)cfg->width) < 16 ||
        (int)(cfg->flow_scale * (float)cfg->height) < 16)
        return SINDYN_ERR_INVALID;
    if (!(cfg->brox_alpha > 0.0f) || !(cfg->brox_gamma >= 0.0f) || !(cfg->brox_pyr_scale > 0.0f && cfg->brox_pyr_scale < 1.0f) || cfg->brox_inner < 1 ||
        cfg->brox_outer < 1 || cfg->brox_solver < 1 || !(cfg->brox_omega > 0.0f && cfg->brox_omega < 2.0f))
        return SINDYN_ERR_INVALID;
    if (!(cfg->fx > 0.0f) || !(cfg->fy > 0.0f) || !(cfg->depth_scale > 0.0f) || !(cfg->depth_weight > 0.0f)) return SINDYN_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device >= ndev) return SINDYN_ERR_NO_DEVICE;
    sindyn_ctx *c = new (std::nothrow) sindyn_ctx();
    if (!c) return SINDYN_ERR_INVALID;
    c->cfg = *cfg;
    c->device = cfg->device;
    c->W = cfg->width; c->H = cfg->height; c->N = c->W * c->H;
    c->fw = (int)(cfg->flow_scale * (float)c->W);
    c->fh = (int)(cfg->flow_scale * (float)c->H);
    *out = c;
    if (cudaSetDevice(c->device) != cudaSuccess) { c->err = "cudaSetDevice failed"; return SINDYN_ERR_CUDA; }
    CU_CHECK(c, cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    const size_t N = (size_t)c->N, NF = (size_t)c->fw * c->fh;
    for (int i = 0; i < 4; ++i) {
        SD_CHECK(c->dalloc(&c->bgr[i], N * 3));
        SD_CHECK(c->dalloc(&c->gray[i], N));
        SD_CHECK(c->dalloc(&c->gsmall[i], NF));
        SD_CHECK(c->dalloc(&c->gsmall_f[i], NF));
    }
    SD_CHECK(c->dalloc(&c->depth, N));
    SD_CHECK(c->halloc(&c->pin_bgr, N * 3));
    SD_CHECK(c->halloc(&c->pin_depth, N));
    SD_CHECK(c->halloc(&c->pin_out0, N));
    SD_CHECK(c->halloc(&c->pin_out1, N));
    SD_CHECK(c->dalloc(&c->dyna_last, N));
    SD_CHECK(c->dalloc(&c->high_last, N));
    SD_CHECK(c->dalloc(&c->label_last, N));
    SD_CHECK(c->dalloc(&c->flow_small, NF * 2));
    SD_CHECK(c->dalloc(&c->flow_full, N * 2));
    SD_CHECK(c->dalloc(&c->scratch_f0, N * 2));
    SD_CHECK(c->dalloc(&c->scratch_f1, N * 2));
    SD_CHECK(c->dalloc(&c->scratch_u0, N * 3));
    SD_CHECK(c->dalloc(&c->scratch_u1, N * 3));
    SD_CHECK(c->dalloc(&c->scratch_u2, N * 3));
    SD_CHECK(c->dalloc(&c->scratch_u3, N * 3));
    SD_CHECK(c->dalloc(&c->mask_low, N));
    SD_CHECK(c->dalloc(&c->mask_high, N));
    SD_CHECK(resize_plan_init(c, &c->plan_flow, c->W, c->H, c->fw, c->fh));
    SD_CHECK(brox_init(c, &c->brox, c->fw, c->fh, cfg->brox_alpha, cfg->brox_gamma, cfg->brox_pyr_scale, cfg->brox_inner,
                       cfg->brox_outer, cfg->brox_solver, cfg->brox_omega));
    SD_CHECK(brox_init(c, &c->brox_lm, c->fw, c->fh, cfg->brox_alpha, cfg->brox_gamma, cfg->brox_pyr_scale, cfg->brox_inner,
                       cfg->brox_outer, cfg->brox_solver, cfg->brox_omega));
    SD_CHECK(residual_init(c, &c->resid, c->W, c->H));
    SD_CHECK(flow_branch_init(c));
    SD_CHECK(sindyn_ctx_init_stages(c));
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    return SINDYN_OK;
}


extern "C" int sindyn_destroy(sindyn_handle h)
{
    if (!h) return SINDYN_ERR_INVALID;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();     // (the frame pipeline and the clustering branch run on further streams)
    pipe_destroy(h);
    brox_destroy(&h->brox);
    brox_destroy(&h->brox_lm);
    flow_tail_drop_graphs(h);
    flow_graph_drop(h);
    sindyn_ctx_destroy_stages(h);
    h->free_all();
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return SINDYN_OK;
}

extern "C" const char *sindyn_last_error(sindyn_handle h) { return h ? h->err.c_str() : "null handle"; }

extern "C" int sindyn_set_stream(sindyn_handle h, void *s)
{
    H_CHECK(h);
    cudaStream_t ns = s ? (cudaStream_t)s : h->own_stream;
    if (ns != h->stream) {
        CU_CHECK(h, cudaStreamSynchronize(h->stream));
        h->stream = ns;
        h->brox.graph_ok = false;
        h->brox_lm.graph_ok = false;  // graphs are stream-agnostic, but re-capture keeps capture semantics simple
        flow_tail_drop_graphs(h);
        flow_graph_drop(h);
        pipe_invalidate(h);
    }
    return SINDYN_OK;
}

extern "C" int sindyn_synchronize(sindyn_handle h)
{
    H_CHECK(h);
    SD_CHECK(pipe_join(h));      // frames in the frame pipeline run on further streams
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    return SINDYN_OK;
}

extern "C" unsigned long long sindyn_launch_count(sindyn_handle h) { return h ? h->launches : 0ull; }

static int prep_frame(sindyn_ctx *c, int slot_idx)
{
    // BGR -> gray -> 0.6x gray (u8 and float/255)   (DynaDetect.cc:1390-1392,1037-1039,1046-1048)
    SD_CHECK(launch_bgr2gray(c, c->bgr[slot_idx], c->W, c->H, c->gray[slot_idx]));
    SD_CHECK(launch_resize_u8(c, &c->plan_flow, c->gray[slot_idx], c->W, c->gsmall[slot_idx], c->fw, c->gsmall_f[slot_idx],
                              1.0f / 255.0f));
    LAUNCH_CHECK(c);
    return SINDYN_OK;
}
int sindyn_prep_frame(sindyn_ctx *c, int idx) { return prep_frame(c, idx); }

extern "C" int sindyn_set_prev_frames(sindyn_handle h, const uint8_t *bgr_last, size_t step_last, const uint8_t *bgr_lastlast,
                                      size_t step_lastlast)
{
    H_CHECK(h);
    if (!bgr_last || !bgr_lastlast) return SINDYN_ERR_INVALID;
    SD_CHECK(pipe_join(h));
    pipe_invalidate(h);
    const size_t rb = (size_t)h->W * 3;
    CU_CHECK(h, copy_in_2d(h->bgr[h->i_last], bgr_last, step_last, rb, h->H, h->stream));
    CU_CHECK(h, copy_in_2d(h->bgr[h->i_lastlast], bgr_lastlast, step_lastlast, rb, h->H, h->stream));
    SD_CHECK(prep_frame(h, h->i_last));
    SD_CHECK(prep_frame(h, h->i_lastlast));
    CU_CHECK(h, cudaMemsetAsync(h->dyna_last, 0, h->N, h->stream));
    CU_CHECK(h, cudaMemsetAsync(h->high_last, 0, h->N, h->stream));
    CU_CHECK(h, cudaMemsetAsync(h->label_last, 0, h->N, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    h->have_prev = true;
    return SINDYN_OK;
}

extern "C" int sindyn_flow_brox(sindyn_handle h, const float *I0, const float *I1, int w, int hgt, float *flow_uv)
{
    H_CHECK(h);
    if (!I0 || !I1 || !flow_uv || w != h->fw || hgt != h->fh) {
        h->err = "sindyn_flow_brox: size must equal the handle's flow grid";
        return SINDYN_ERR_INVALID;
    }
    const size_t nb = sizeof(float) * (size_t)w * hgt;
    CU_CHECK(h, cudaMemcpyAsync(h->scratch_f0, I0, nb, cudaMemcpyHostToDevice, h->stream));
    CU_CHECK(h, cudaMemcpyAsync(h->scratch_f1, I1, nb, cudaMemcpyHostToDevice, h->stream));
    SD_CHECK(brox_run(h, &h->brox, h->scratch_f0, h->scratch_f1, h->flow_small, 1.0f, h->cfg.use_graphs != 0));
    CU_CHECK(h, cudaMemcpyAsync(flow_uv, h->flow_small, nb * 2, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    return SINDYN_OK;
}

extern "C" int sindyn_gray_resize(sindyn_handle h, const uint8_t *bgr, size_t bgr_step, uint8_t *gray_full, uint8_t *gray_small)
{
    H_CHECK(h);
    if (!bgr) return SINDYN_ERR_INVALID;
    CU_CHECK(h, copy_in_2d(h->scratch_u0, bgr, bgr_step, (size_t)h->W * 3, h->H, h->stream));
    SD_CHECK(launch_bgr2gray(h, h->scratch_u0, h->W, h->H, h->scratch_u1));
    SD_CHECK(launch_resize_u8(h, &h->plan_flow, h->scratch_u1, h->W, h->scratch_u2, h->fw, nullptr, 0.f));
    LAUNCH_CHECK(h);
    if (gray_full) CU_CHECK(h, cudaMemcpyAsync(gray_full, h->scratch_u1, h->N, cudaMemcpyDeviceToHost, h->stream));
    if (gray_small) CU_CHECK(h, cudaMemcpyAsync(gray_small, h->scratch_u2, (size_t)h->fw * h->fh, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    return SINDYN_OK;
}

static int residual_out(sindyn_ctx *h, float *residual_mag_out, uint8_t *mask_low, uint8_t *mask_high, float *thr_out)
{
    if (residual_mag_out) CU_CHECK(h, cudaMemcpyAsync(residual_mag_out, h->resid.mag, sizeof(float) * h->N, cudaMemcpyDeviceToHost, h->stream));
    if (mask_low) CU_CHECK(h, cudaMemcpyAsync(mask_low, h->mask_low, h->N, cudaMemcpyDeviceToHost, h->stream));
    if (mask_high) CU_CHECK(h, cudaMemcpyAsync(mask_high, h->mask_high, h->N, cudaMemcpyDeviceToHost, h->stream));
    if (thr_out) CU_CHECK(h, cudaMemcpyAsync(thr_out, h->resid.thr, sizeof(float) * 4, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    return SINDYN_OK;
}

extern "C" int sindyn_residual_homography(sindyn_handle h, const float *flow, const double *Hm, float *residual_mag_out,
                                          uint8_t *mask_low, uint8_t *mask_high, float *thresholds_out)
{
    H_CHECK(h);
    if (!flow || !Hm) return SINDYN_ERR_INVALID;
    CU_CHECK(h, cudaMemcpyAsync(h->flow_full, flow, sizeof(float) * 2 * h->N, cudaMemcpyHostToDevice, h->stream));
    SD_CHECK(residual_homography_run(h, &h->resid, h->flow_full, Hm, h->mask_low, h->mask_high));
    return residual_out(h, residual_mag_out, mask_low, mask_high, thresholds_out);
}

extern "C" int sindyn_residual_pose(sindyn_handle h, const float *flow, const uint16_t *depth, size_t depth_step,
                                    const double *T_old_cur, float *residual_mag_out, uint8_t *mask_low, uint8_t *mask_high,
                                    float *thresholds_out)
{
    H_CHECK(h);
    if (!flow || !depth || !T_old_cur) return SINDYN_ERR_INVALID;
    CU_CHECK(h, cudaMemcpyAsync(h->flow_full, flow, sizeof(float) * 2 * h->N, cudaMemcpyHostToDevice, h->stream));
    CU_CHECK(h, copy_in_2d(h->depth, depth, depth_step, (size_t)h->W * 2, h->H, h->stream));
    SD_CHECK(residual_pose_run(h, &h->resid, h->flow_full, h->depth, T_old_cur, h->cfg.fx, h->cfg.fy, h->cfg.cx, h->cfg.cy,
                               h->cfg.depth_scale, h->mask_low, h->mask_high));
    return residual_out(h, residual_mag_out, mask_low, mask_high, thresholds_out);
}

extern "C" int sindyn_get_state(sindyn_handle h, int which, uint8_t *out)
{
    H_CHECK(h);
    if (!out) return SINDYN_ERR_INVALID;
    const uint8_t *src = nullptr;
    size_t nb = h->N;
    switch (which) {
    case 0: src = h->dyna_last; break;
    case 1: src = h->high_last; break;
    case 2: src = h->label_last; break;
    case 3: src = h->bgr[h->i_last]; nb *= 3; break;
    case 4: src = h->bgr[h->i_lastlast]; nb *= 3; break;
    default: return SINDYN_ERR_INVALID;
    }
    SD_CHECK(pipe_join(h));
    CU_CHECK(h, cudaMemcpyAsync(out, src, nb, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    return SINDYN_OK;
}

extern "C" int sindyn_set_state(sindyn_handle h, int which, const uint8_t *in)
{
    H_CHECK(h);
    if (!in) return SINDYN_ERR_INVALID;
    uint8_t *dst = nullptr;
    size_t nb = h->N;
    int prep = -1;
    switch (which) {
    case 0: dst = h->dyna_last; break;
    case 1: dst = h->high_last; break;
    case 2: dst = h->label_last; break;
    case 3: dst = h->bgr[h->i_last]; nb *= 3; prep = h->i_last; break;
    case 4: dst = h->bgr[h->i_lastlast]; nb *= 3; prep = h->i_lastlast; break;
    default: return SINDYN_ERR_INVALID;
    }
    SD_CHECK(pipe_join(h));      // frames still in the frame pipeline finish first; it re-reads the state on its next frame
    pipe_invalidate(h);
    CU_CHECK(h, cudaMemcpyAsync(dst, in, nb, cudaMemcpyHostToDevice, h->stream));
    if (prep >= 0) { SD_CHECK(prep_frame(h, prep)); h->have_prev = true; }
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    return SINDYN_OK;
}

extern "C" int sindyn_get_stage_ms(sindyn_handle h, float *ms, int n)
{
    if (!h || !ms || n < 0) return SINDYN_ERR_INVALID;
    for (int i = 0; i < n && i < 16; ++i) ms[i] = h->stage_ms[i];
    return SINDYN_OK;
}
