// peac.cu -- placeholder until the AHC plane-fitter kernels land (fails loudly; create the handle with plane_edges = 0).
#include "peac.cuh"

int peac_init(sindyn_base *, PeacStage *p, int W, int H) { p->W = W; p->H = H; return SINDYN_OK; }
int peac_run(sindyn_base *ctx, PeacStage *, const uint16_t *, float, float, float, float, float, uint8_t *)
{
    ctx->err = "PEAC plane edges are not built yet; create the handle with plane_edges = 0";
    return SINDYN_ERR_STATE;
}
