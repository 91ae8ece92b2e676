// homography.cu -- weighted grid sampling of the dense flow and robust homography estimation.
//
// Replaces ORB_SLAM2/src/DynaDetect.cc:1163-1235:
//   clusterWeight (:1169-1177), the 10-px sample grid with weight = RNG(12345).gaussian(0.5) + class term
//   (:1182-1204), std::sort by weight descending (:1217-1219), the inBorder filter with its int truncation and
//   inclusive upper bound (:1221-1231, inBorder :103-106), and cv::findHomography(pts, ptsLast, noArray(), RHO)
//   (:1235).
// The sample list is bit-exact against the oracle.  cv::findHomography(RHO) is OpenCV's PROSAC+SPRT estimator
// (un-vendored; order sensitive); it is replaced by a deterministic PROSAC-style parallel
// hypothesise-and-verify estimator: HG_M minimal 4-point hypotheses drawn progressively from the top of the
// sorted list (one warp each: DLT in double + inlier count at the RHO default 3 px), best consensus, then
// Gauss-Newton refinement of the reprojection error on the inliers.  Parity with RHO is a stated tolerance
// on the induced flow (tests/test_homography_gpu.py).
#include "homography.cuh"

__device__ const float c_gauss[HG_GAUSS_N] = {
#include "gauss_table.inc"
};

// counts[0][i] = |label_last == i|, counts[1][i] = |label_last == i && dyna_last == 255|
__global__ void k_cluster_weight_counts(const uint8_t *__restrict__ label_last, const uint8_t *__restrict__ dyna_last, int n,
                                        int *__restrict__ counts)
{
    __shared__ int sc[2 * 16];
    if (threadIdx.x < 32) sc[threadIdx.x] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int l = label_last[i];
        if (l >= 1 && l < 12) {
            atomicAdd(&sc[l], 1);
            if (dyna_last[i] == 255) atomicAdd(&sc[16 + l], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x < 32 && sc[threadIdx.x]) atomicAdd(&counts[threadIdx.x], sc[threadIdx.x]);
}

__device__ __forceinline__ unsigned int f2ord_desc(float f)
{
    unsigned int u = __float_as_uint(f);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending order of f
    return ~u;                                        // descending
}

// single CTA: weights, stable descending sort, inBorder filter, ordered compaction
__global__ void __launch_bounds__(1024) k_sample_pairs(const float2 *__restrict__ flow, const uint8_t *__restrict__ label_last,
                                                       const uint8_t *__restrict__ dyna_last, const int *__restrict__ counts, int W, int H,
                                                       float2 *__restrict__ pts, float2 *__restrict__ pts_last, int *__restrict__ n_out)
{
    __shared__ unsigned long long keys[HG_MAX_SAMPLES];
    __shared__ float cw[12];
    __shared__ int s_scan[1024 / 32];
    __shared__ int s_base;
    const int ncol = (W - 1) / 10, nrow = (H - 1) / 10;  // rows 10,20,.. < H ; cols 10,20,.. < W
    const int ns = ncol * nrow;
    if (threadIdx.x < 12) {
        int i = threadIdx.x;
        cw[i] = i >= 1 ? (float)counts[16 + i] / ((float)counts[i] + 1.0f) : 0.0f;
    }
    __syncthreads();
    // Bitonic sort of HG_MAX_SAMPLES (weight desc, index) keys, 4 consecutive keys per thread: compare-exchange distances 1-2
    // stay in the thread's registers, 4-64 are warp shuffles, only distances >= 128 go through shared memory (15 of the 78
    // stages).  The keys are unique (the index is part of the key), so the order equals std::sort's on distinct weights and
    // the stable order on equal ones.
    static_assert(HG_MAX_SAMPLES == 4096, "4 keys per thread x 1024 threads");
    unsigned long long e[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int k = 4 * threadIdx.x + m;
        unsigned long long key = ~0ull;
        if (k < ns) {
            int r = k / ncol, c = k - r * ncol;
            int row = 10 + 10 * r, col = 10 + 10 * c;
            float randomd = c_gauss[k];
            uint8_t dl = dyna_last[row * W + col];
            float w;
            if (dl < 20) w = randomd + 1.0f;
            else if ((unsigned)(dl - 20) <= 230u - 20u) w = randomd + 1.2f * (1.0f - cw[min((int)label_last[row * W + col], 11)]);
            else w = randomd + 0.4f;
            key = ((unsigned long long)f2ord_desc(w) << 32) | (unsigned)k;
        }
        e[m] = key;
    }
    for (int k = 2; k <= HG_MAX_SAMPLES; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 128) {
#pragma unroll
                for (int m = 0; m < 4; ++m) keys[4 * threadIdx.x + m] = e[m];
                __syncthreads();
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const int i = 4 * threadIdx.x + m;
                    const unsigned long long o = keys[i ^ j];
                    const bool take_min = ((i & j) == 0) == ((i & k) == 0);
                    e[m] = take_min ? (o < e[m] ? o : e[m]) : (o > e[m] ? o : e[m]);
                }
                __syncthreads();
            } else if (j >= 4) {
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const int i = 4 * threadIdx.x + m;
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, e[m], j >> 2);
                    const bool take_min = ((i & j) == 0) == ((i & k) == 0);
                    e[m] = take_min ? (o < e[m] ? o : e[m]) : (o > e[m] ? o : e[m]);
                }
            } else {
                // (distances 1 and 2 spelled out: a run-time register index would push e[] into local memory)
#define HG_CE(A, B)                                                                      \
    {                                                                                    \
        const bool up = ((4 * threadIdx.x + (A)) & k) == 0;                              \
        const unsigned long long a_ = e[A], b_ = e[B];                                   \
        if ((a_ > b_) == up) { e[A] = b_; e[B] = a_; }                                   \
    }
                if (j == 2) { HG_CE(0, 2) HG_CE(1, 3) }
                else { HG_CE(0, 1) HG_CE(2, 3) }
#undef HG_CE
            }
        }
#pragma unroll
    for (int m = 0; m < 4; ++m) keys[4 * threadIdx.x + m] = e[m];
    __syncthreads();
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int start = 0; start < ns; start += blockDim.x) {
        int i = start + threadIdx.x;
        bool keep = false;
        float pc = 0.f, pr = 0.f, lx = 0.f, ly = 0.f;
        if (i < ns) {
            int k = (int)(keys[i] & 0xffffffffull);
            int r = k / ncol, c = k - r * ncol;
            int row = 10 + 10 * r, col = 10 + 10 * c;
            float2 f = flow[row * W + col];
            pc = (float)col; pr = (float)row;
            lx = pc - f.x; ly = pr - f.y;
            int irow = (int)ly, icol = (int)lx;  // float -> int truncation at the inBorder call (DynaDetect.cc:1226)
            keep = ((unsigned)irow <= (unsigned)H) && ((unsigned)icol <= (unsigned)W);
        }
        unsigned bal = __ballot_sync(0xffffffffu, keep);
        int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) s_scan[wid] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w2 = 0; w2 < wid; ++w2) off += s_scan[w2];
        off += __popc(bal & ((1u << lane) - 1u));
        if (keep) { pts[off] = make_float2(pc, pr); pts_last[off] = make_float2(lx, ly); }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) t += s_scan[w2];
            s_base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = s_base;
}

// ---------------------------------------------------------------- minimal solver
// 8 x 8 Gaussian elimination with partial pivoting (first largest pivot) + back substitution, spread over a warp: lane r (< 8)
// owns row r of the augmented matrix in registers, pivot search / row swap / pivot-row broadcast go through shuffles.  Every
// element sees exactly the operations of the textbook serial elimination in the same order (results verified bit-identical to
// the serial version this replaced); the system never touches local memory and the row updates of a column run in parallel.
// All 32 lanes must call it (lanes >= 8 only take part in the shuffles); the solution is returned in every lane.
__device__ bool solve8_warp(double row[9], double x[8])
{
    const int lane = threadIdx.x & 31;
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        // first row r >= c with the largest |A[r][c]|
        double best = (lane >= c && lane < 8) ? fabs(row[c]) : -1.0;
        int piv = lane;
        for (int o = 4; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int op = __shfl_xor_sync(0xffffffffu, piv, o);
            if (ob > best || (ob == best && op < piv)) { best = ob; piv = op; }
        }
        best = __shfl_sync(0xffffffffu, best, 0);
        piv = __shfl_sync(0xffffffffu, piv, 0);
        if (!(best > 1e-12)) { ok = false; break; }
        const int src = lane == c ? piv : (lane == piv ? c : lane);
        double inv = 0.0, f = 0.0;
#pragma unroll
        for (int k = 0; k < 9; ++k) row[k] = __shfl_sync(0xffffffffu, row[k], src);   // swap rows c and piv (whole rows: columns < c are dead)
        {
            const double pc = __shfl_sync(0xffffffffu, row[c], c);
            inv = 1.0 / pc;
            f = row[c] * inv;
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const double pk = __shfl_sync(0xffffffffu, row[k], c);
            if (k >= c && lane > c && lane < 8 && f != 0.0) row[k] -= f * pk;
        }
    }
    if (!ok) return false;
#pragma unroll
    for (int r = 7; r >= 0; --r) {
        double s = row[8];
#pragma unroll
        for (int k = r + 1; k < 8; ++k) s -= row[k] * x[k];
        const double xr = s / row[r];
        x[r] = __shfl_sync(0xffffffffu, xr, r);
    }
    return true;
}

__device__ __forceinline__ unsigned int hg_hash(unsigned int x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// one warp per hypothesis
__global__ void __launch_bounds__(256) k_homog_hypotheses(const float2 *__restrict__ pts, const float2 *__restrict__ pts_last,
                                                          const int *__restrict__ n_ptr, double *__restrict__ Hs, int *__restrict__ scores)
{
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (m >= HG_M) return;
    const int n = *n_ptr;
    double h[8];
    bool ok = n >= 4;
    if (ok) {
        // PROSAC-style progressive sampling pool: the top T of the weight-sorted list
        int T = max(8, (int)(((long long)n * (m + 1) + HG_M - 1) / HG_M));
        T = min(T, n);
        int idx[4];
        unsigned int s = hg_hash(0x9e3779b9u * (unsigned)(m + 1));
        for (int k = 0; k < 4; ++k) {
            for (int tries = 0; tries < 16; ++tries) {
                s = hg_hash(s + 0x632be5abu);
                idx[k] = (int)(s % (unsigned)T);
                bool dup = false;
                for (int q = 0; q < k; ++q) dup |= idx[q] == idx[k];
                if (!dup) break;
            }
        }
        // DLT rows: lane 2k holds the u-equation of sample k, lane 2k + 1 its v-equation
        double row[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (lane < 8) {
            const int k = lane >> 1;
            const double x = pts[idx[k]].x, y = pts[idx[k]].y, u = pts_last[idx[k]].x, v = pts_last[idx[k]].y;
            if (lane & 1) { row[3] = x; row[4] = y; row[5] = 1; row[6] = -v * x; row[7] = -v * y; row[8] = v; }
            else { row[0] = x; row[1] = y; row[2] = 1; row[6] = -u * x; row[7] = -u * y; row[8] = u; }
        }
        ok = solve8_warp(row, h);
        for (int c = 0; c < 8; ++c) ok = ok && isfinite(h[c]);
    }
    int cnt = 0;
    if (ok) {
        for (int i = lane; i < n; i += 32) {
            double x = pts[i].x, y = pts[i].y;
            double w = h[6] * x + h[7] * y + 1.0;
            double ex = (h[0] * x + h[1] * y + h[2]) / w - pts_last[i].x;
            double ey = (h[3] * x + h[4] * y + h[5]) / w - pts_last[i].y;
            cnt += (w > 1e-9 && ex * ex + ey * ey <= HG_THR2) ? 1 : 0;
        }
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) {
        scores[m] = ok ? cnt : -1;
        for (int c = 0; c < 8; ++c) Hs[m * 8 + c] = ok ? h[c] : 0.0;
    }
}

// ------------------------------------------------------------------ locally optimised consensus search
// rank of every hypothesis by inlier count (ties: lower index first); the HG_TOPK best go on to refinement
__global__ void __launch_bounds__(256) k_homog_rank(const int *__restrict__ scores, int *__restrict__ top)
{
    __shared__ int sc[HG_M];
    for (int i = threadIdx.x; i < HG_M; i += blockDim.x) sc[i] = scores[i];
    __syncthreads();
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= HG_M) return;
    const int mine = sc[m];
    int rank = 0;
    for (int j = 0; j < HG_M; ++j) rank += (sc[j] > mine) || (sc[j] == mine && j < m);
    if (rank < HG_TOPK) top[rank] = m;
}

// One CTA per candidate: Gauss-Newton on the reprojection error of the inliers (re-selected every iteration), then the
// final consensus size.  A minimal-sample hypothesis that looks second best can refine into the largest consensus set
// (two competing planes), which is what the reference's RHO estimator returns -- hence HG_TOPK candidates, not one.
// 256 threads: the 45 normal-equation accumulators of a thread stay in registers, the cross-warp reduction is done by
// 45 threads in parallel, and the loop stops once the update is below 1e-12.
#define HG_RT 256
__global__ void __launch_bounds__(HG_RT) k_homog_refine(const float2 *__restrict__ pts, const float2 *__restrict__ pts_last,
                                                        const int *__restrict__ n_ptr, const double *__restrict__ Hs,
                                                        const int *__restrict__ scores, const int *__restrict__ top,
                                                        double *__restrict__ Hc, double *__restrict__ cand_cost)
{
    constexpr int NW = HG_RT / 32;
    __shared__ double s_acc[NW][45];
    __shared__ double s_tot[45];
    __shared__ double s_h[8];
    __shared__ int s_nin[NW];
    __shared__ int s_done;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = *n_ptr;
    const int cand = blockIdx.x;
    const int m = top[cand];
    const int best = scores[m];
    if (tid == 0) {
        if (best >= 4) for (int c = 0; c < 8; ++c) s_h[c] = Hs[m * 8 + c];
        else { for (int c = 0; c < 8; ++c) s_h[c] = 0.0; s_h[0] = 1.0; s_h[4] = 1.0; }  // identity when no consensus
        s_done = best >= 4 ? 0 : 1;
    }
    __syncthreads();
    int final_n = 0;
    double final_sse = 0.0;
    for (int iter = 0; iter <= HG_GN_ITERS; ++iter) {
        // the last pass (iter == HG_GN_ITERS, or right after convergence) only evaluates the consensus of the final H
        const bool last = iter == HG_GN_ITERS || s_done;
        double acc[45];
#pragma unroll
        for (int k = 0; k < 45; ++k) acc[k] = 0.0;
        int nin = 0;
        double h[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) h[c] = s_h[c];
        for (int i = tid; i < n; i += HG_RT) {
            const float2 pp = pts[i], pl = pts_last[i];
            const double x = pp.x, y = pp.y;
            const double w = h[6] * x + h[7] * y + 1.0;
            const double iw = 1.0 / w;
            const double uh = (h[0] * x + h[1] * y + h[2]) * iw, vh = (h[3] * x + h[4] * y + h[5]) * iw;
            const double rx = pl.x - uh, ry = pl.y - vh;
            if (!(w > 1e-9) || rx * rx + ry * ry > HG_THR2) continue;
            ++nin;
            acc[44] += rx * rx + ry * ry;
            if (last) continue;
            const double Ju[8] = {x * iw, y * iw, iw, 0, 0, 0, -uh * x * iw, -uh * y * iw};
            const double Jv[8] = {0, 0, 0, x * iw, y * iw, iw, -vh * x * iw, -vh * y * iw};
            int k = 0;
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = a; b < 8; ++b) acc[k++] += Ju[a] * Ju[b] + Jv[a] * Jv[b];
#pragma unroll
            for (int a = 0; a < 8; ++a) acc[36 + a] += Ju[a] * rx + Jv[a] * ry;
        }
#pragma unroll
        for (int k = 0; k < 45; ++k) {
            double v = acc[k];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_acc[wid][k] = v;
        }
        for (int o = 16; o > 0; o >>= 1) nin += __shfl_xor_sync(0xffffffffu, nin, o);
        if (lane == 0) s_nin[wid] = nin;
        __syncthreads();
        if (tid < 45) { double v = 0; for (int w2 = 0; w2 < NW; ++w2) v += s_acc[w2][tid]; s_tot[tid] = v; }
        __syncthreads();
        int tn = 0;
        for (int w2 = 0; w2 < NW; ++w2) tn += s_nin[w2];
        final_n = tn;
        final_sse = s_tot[44];
        if (last) break;
        if (wid == 0) {   // normal equations: warp-cooperative solve (bit-identical to the serial elimination)
            if (tn >= 8) {
                double row[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, dx[8];
                if (lane < 8) {
                    const int a = lane;
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        const int lo = a < b ? a : b, hi = a < b ? b : a;
                        row[b] = s_tot[lo * 8 - (lo * (lo - 1)) / 2 + (hi - lo)];
                    }
                    row[8] = s_tot[36 + a];
                }
#pragma unroll
                for (int b = 0; b < 8; ++b)
                    if (lane == b) row[b] *= 1.0 + 1e-9;
                if (solve8_warp(row, dx)) {
                    bool fin = true;
                    double mx = 0.0;
                    for (int c = 0; c < 8; ++c) { fin = fin && isfinite(dx[c]); mx = fmax(mx, fabs(dx[c])); }
                    if (lane == 0) {
                        if (fin) for (int c = 0; c < 8; ++c) s_h[c] += dx[c];
                        if (!fin || mx < 1e-12) s_done = 1;
                    }
                } else if (lane == 0) s_done = 1;
            } else if (lane == 0) s_done = 1;
        }
        __syncthreads();
    }
    if (tid == 0) {
        for (int c = 0; c < 8; ++c) Hc[cand * 8 + c] = s_h[c];
        cand_cost[cand * 2] = best >= 4 ? (double)final_n : -1.0;
        cand_cost[cand * 2 + 1] = final_sse;
    }
}

// largest refined consensus (ties: smaller squared error, then better original rank)
__global__ void k_homog_pick(const double *__restrict__ Hc, const double *__restrict__ cand_cost, const int *__restrict__ top,
                             const int *__restrict__ scores, double *__restrict__ H_out, int *__restrict__ info)
{
    if (threadIdx.x) return;
    int bi = 0;
    for (int c = 1; c < HG_TOPK; ++c)
        if (cand_cost[2 * c] > cand_cost[2 * bi] || (cand_cost[2 * c] == cand_cost[2 * bi] && cand_cost[2 * c + 1] < cand_cost[2 * bi + 1])) bi = c;
    for (int c = 0; c < 8; ++c) H_out[c] = Hc[bi * 8 + c];
    H_out[8] = 1.0;
    info[0] = scores[top[0]];
    info[1] = top[bi];
    info[2] = (int)cand_cost[2 * bi];
}

int homography_init(sindyn_base *ctx, HomographyStage *g, int W, int H)
{
    g->W = W; g->H = H;
    int ns = ((W - 1) / 10) * ((H - 1) / 10);
    if (ns > HG_MAX_SAMPLES || ns > HG_GAUSS_N) { ctx->err = "homography: sample grid exceeds HG_MAX_SAMPLES"; return SINDYN_ERR_CAPACITY; }
    SD_CHECK(ctx->dalloc(&g->counts, 32));
    SD_CHECK(ctx->dalloc(&g->pts, 2 * HG_MAX_SAMPLES));
    SD_CHECK(ctx->dalloc(&g->pts_last, 2 * HG_MAX_SAMPLES));
    SD_CHECK(ctx->dalloc(&g->n_pairs, 4));
    SD_CHECK(ctx->dalloc(&g->Hs, 8 * HG_M));
    SD_CHECK(ctx->dalloc(&g->scores, HG_M));
    SD_CHECK(ctx->dalloc(&g->H_dev, 9));
    SD_CHECK(ctx->dalloc(&g->top, HG_TOPK));
    SD_CHECK(ctx->dalloc(&g->Hc, 8 * HG_TOPK));
    SD_CHECK(ctx->dalloc(&g->cand_cost, 2 * HG_TOPK));
    return SINDYN_OK;
}

int homography_sample(sindyn_base *ctx, HomographyStage *g, const float *flow, const uint8_t *label_last, const uint8_t *dyna_last)
{
    CU_CHECK(ctx, cudaMemsetAsync(g->counts, 0, sizeof(int) * 32, ctx->stream));
    LAUNCH(ctx, k_cluster_weight_counts, SINDYN_NUM_SMS_B200, 256, 0, label_last, dyna_last, g->W * g->H, g->counts);
    LAUNCH(ctx, k_sample_pairs, 1, 1024, 0, (const float2 *)flow, label_last, dyna_last, g->counts, g->W, g->H, (float2 *)g->pts,
           (float2 *)g->pts_last, g->n_pairs);
    LAUNCH_CHECK(ctx);
    return SINDYN_OK;
}

int homography_estimate(sindyn_base *ctx, HomographyStage *g)
{
    LAUNCH(ctx, k_homog_hypotheses, cdiv(HG_M * 32, 256), 256, 0, (const float2 *)g->pts, (const float2 *)g->pts_last, g->n_pairs, g->Hs,
           g->scores);
    LAUNCH(ctx, k_homog_rank, cdiv(HG_M, 256), 256, 0, g->scores, g->top);
    LAUNCH(ctx, k_homog_refine, HG_TOPK, HG_RT, 0, (const float2 *)g->pts, (const float2 *)g->pts_last, g->n_pairs, g->Hs, g->scores, g->top,
           g->Hc, g->cand_cost);
    LAUNCH(ctx, k_homog_pick, 1, 32, 0, g->Hc, g->cand_cost, g->top, g->scores, g->H_dev, g->n_pairs + 1);
    LAUNCH_CHECK(ctx);
    return SINDYN_OK;
}
