// varref.cu -- placeholder until the VariationalRefinement-equivalent kernels land (fails loudly).
#include "varref.cuh"

int varref_init(sindyn_base *, VarRefStage *v, int w, int h) { v->w = w; v->h = h; return SINDYN_OK; }
int varref_run(sindyn_base *ctx, VarRefStage *, const uint8_t *, const uint8_t *, float *)
{
    ctx->err = "flow refinement (cv::VariationalRefinement equivalent) is not built yet; create the handle with refine=0";
    return SINDYN_ERR_STATE;
}
