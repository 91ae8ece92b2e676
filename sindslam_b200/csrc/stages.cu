// stages.cu -- C-ABI entry points of the clustering branch (k-means, cluster order, depth edges,
// morphology).  See include/sindyn.h for the reference lines each one replaces.
#include "ctx.cuh"

#define H_CHECK(h)                       \
    if (!(h)) return SINDYN_ERR_INVALID; \
    cudaSetDevice((h)->device)

int sindyn_ctx_init_stages(sindyn_ctx *c)
{
    SD_CHECK(morph_init(c));
    SD_CHECK(kmeans_init(c, &c->km, c->W, c->H));
    SD_CHECK(edges_init(c, &c->edges, c->W, c->H));
    SD_CHECK(peac_init(c, &c->peac, c->W, c->H));
    SD_CHECK(recluster_init(c, &c->rc, c->W, c->H));
    SD_CHECK(decide_init(c, &c->dd, c->W, c->H));
    SD_CHECK(c->dalloc(&c->plane_edges, (size_t)c->N));
    CU_CHECK(c, cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    CU_CHECK(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CU_CHECK(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    CU_CHECK(c, cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking));
    CU_CHECK(c, cudaEventCreateWithFlags(&c->ev_peac_fork, cudaEventDisableTiming));
    CU_CHECK(c, cudaEventCreateWithFlags(&c->ev_peac_join, cudaEventDisableTiming));
    return SINDYN_OK;
}
void sindyn_ctx_destroy_stages(sindyn_ctx *c)
{
    cloud_stage_destroy(c);
    if (c->cluster_graph) cudaGraphExecDestroy(c->cluster_graph);
    if (c->ev_flag) cudaEventDestroy(c->ev_flag);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->stream3) cudaStreamDestroy(c->stream3);
    if (c->ev_peac_fork) cudaEventDestroy(c->ev_peac_fork);
    if (c->ev_peac_join) cudaEventDestroy(c->ev_peac_join);
}

extern "C" int sindyn_morph_ellipse(sindyn_handle h, const uint8_t *src, size_t src_step, uint8_t *dst, size_t dst_step, int width,
                                    int height, int k, int op)
{
    H_CHECK(h);
    if (!src || !dst || width <= 0 || height <= 0 || (size_t)width * height > (size_t)h->N * 3) return SINDYN_ERR_INVALID;
    CU_CHECK(h, copy_in_2d(h->scratch_u0, src, src_step, width, height, h->stream));
    SD_CHECK(morph_run(h, h->scratch_u0, h->scratch_u1, h->scratch_u2, width, height, k, op));
    CU_CHECK(h, copy_out_2d(dst, dst_step, h->scratch_u1, width, height, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    return SINDYN_OK;
}

extern "C" int sindyn_kmeans(sindyn_handle h, const uint16_t *depth, size_t depth_step, uint8_t *labels_out, float *points_out,
                             float *centers_out)
{
    H_CHECK(h);
    if (!depth) return SINDYN_ERR_INVALID;
    CU_CHECK(h, copy_in_2d(h->depth, depth, depth_step, (size_t)h->W * 2, h->H, h->stream));
    SD_CHECK(kmeans_run(h, &h->km, h->depth, h->label_last, &h->cfg));
    if (labels_out) CU_CHECK(h, cudaMemcpyAsync(labels_out, h->km.labels_u8, h->N, cudaMemcpyDeviceToHost, h->stream));
    if (points_out) CU_CHECK(h, cudaMemcpyAsync(points_out, h->km.points, sizeof(float) * 3 * h->N, cudaMemcpyDeviceToHost, h->stream));
    if (centers_out)
        CU_CHECK(h, cudaMemcpyAsync(centers_out, h->km.state[0].centers, sizeof(float) * 3 * KM_K, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    return SINDYN_OK;
}

extern "C" int sindyn_cluster_order(sindyn_handle h, uint8_t *label_for_seg_edge_out, int *order_out, int *n_kept_out)
{
    H_CHECK(h);
    // k-means already ran the ordering kernels; dilate 7x7 (DynaDetect.cc:1491) happens here for the stage API
    SD_CHECK(morph_run(h, h->km.seg_edge, h->scratch_u0, h->scratch_u1, h->W, h->H, 7, MORPH_DILATE));
    ClusterOrder co;
    CU_CHECK(h, cudaMemcpyAsync(&co, h->km.order, sizeof co, cudaMemcpyDeviceToHost, h->stream));
    if (label_for_seg_edge_out) CU_CHECK(h, cudaMemcpyAsync(label_for_seg_edge_out, h->scratch_u0, h->N, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    if (order_out) for (int i = 0; i < KM_K; ++i) order_out[i] = co.kept[i];
    if (n_kept_out) *n_kept_out = co.n_kept;
    return SINDYN_OK;
}

extern "C" int sindyn_depth_edges(sindyn_handle h, const uint16_t *depth, size_t depth_step, uint8_t *total_area_out,
                                  uint8_t *grad_edges_out, int *endpoints_out, int capacity, int *n_endpoints_out)
{
    H_CHECK(h);
    if (!depth) return SINDYN_ERR_INVALID;
    CU_CHECK(h, copy_in_2d(h->depth, depth, depth_step, (size_t)h->W * 2, h->H, h->stream));
    SD_CHECK(edges_run(h, &h->edges, h->depth, h->cfg.depth_scale));
    int sc[4];
    CU_CHECK(h, cudaMemcpyAsync(sc, h->edges.scalars, sizeof sc, cudaMemcpyDeviceToHost, h->stream));
    if (total_area_out) CU_CHECK(h, cudaMemcpyAsync(total_area_out, h->edges.total_area, h->N, cudaMemcpyDeviceToHost, h->stream));
    if (grad_edges_out) CU_CHECK(h, cudaMemcpyAsync(grad_edges_out, h->edges.grad_edges, h->N, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    if (sc[3]) { h->err = "depth_edges: more than EDGE_EP_CAP candidate end points"; return SINDYN_ERR_CAPACITY; }
    int n = sc[2];
    if (n_endpoints_out) *n_endpoints_out = n;
    if (endpoints_out) {
        if (n > capacity) { h->err = "depth_edges: endpoint buffer too small"; return SINDYN_ERR_CAPACITY; }
        CU_CHECK(h, cudaMemcpy(endpoints_out, h->edges.ep_xy, sizeof(int) * 2 * n, cudaMemcpyDeviceToHost));
    }
    return SINDYN_OK;
}
