// orb.cu -- the masked ORB extractor of SInDSLAM (ORB_SLAM2::ORBextractor, ORB_SLAM2/src/ORBextractor.cc,
// include/ORBextractor.h) as sm_100a kernels behind the C ABI of include/sindyn.h:
//   pyramid (8 levels, OpenCV-exact u8 bilinear, +19 px REFLECT_101)            ORBextractor.cc:1166-1191
//   FAST-9/16 score map, per-cell two-threshold detection with in-window NMS    :765-829
//   quadtree distribution (one CTA per level, list order emulated in shared memory) :481-763
//   intensity-centroid orientation (one warp per keypoint, fastAtan2 polynomial)    :77-104
//   dynamic-mask erasure with the < 250 fallback                                    :1058-1116
//   7x7 sigma-2 Gaussian blur (fixed point, exact) + 256-bit steered BRIEF          :108-147,1135-1151
// All control (counts, offsets, list surgery) stays on the device; the host reads back one counter.
#include "ctx.cuh"

#include <math.h>

#define ORB_MAX_LEVELS 12
#define ORB_EDGE 19              // EDGE_THRESHOLD
#define ORB_HALF_PATCH 15
#define ORB_KMAX 16384           // FAST candidates per level handled by the quadtree
#define ORB_NODES 4096           // quadtree node pool per level
#define ORB_MAX_CELLS 2048       // cells per level
#define ORB_OUT_MAX 8192         // keypoints over all levels

__constant__ signed char c_orb_pattern[256 * 4] = {
#include "orb_pattern.inc"
};
__constant__ int c_umax[ORB_HALF_PATCH + 1];

struct OrbLevel {
    int w, h, pitch;             // unpadded size, padded row pitch (w + 38)
    size_t pad_off;              // offset of the padded image inside pyr
    size_t img_off;              // offset of the unpadded planes (score / blurred)
    int min_b, max_bx, max_by;   // minBorder = 16, maxBorder = size - 16
    int n_cols, n_rows, w_cell, h_cell, cell_off;
    int quota;                   // mnFeaturesPerLevel
    float scale, mask_scale, size;
};

struct OrbCand { unsigned short x, y; unsigned short resp, pad; };   // relative to minBorder

struct OrbKp { float x, y, angle, resp; int level; int keep; };

struct OrbControl {
    int cand_count[ORB_MAX_LEVELS];
    int kp_count[ORB_MAX_LEVELS];
    int overflow;
    int n_out, n_total;
};

struct sindyn_orb : sindyn_base {
    int nfeatures = 0, nlevels = 0, ini_th = 0, min_th = 0, W = 0, H = 0;
    float scale_factor = 0.f;
    OrbLevel lv[ORB_MAX_LEVELS];
    OrbLevel *lv_dev = nullptr;
    ResizePlanU8 plan[ORB_MAX_LEVELS];
    uint8_t *pyr = nullptr, *score = nullptr, *blur = nullptr, *mask = nullptr;
    int *cell_count = nullptr, *cell_offs = nullptr;    // nlevels x ORB_MAX_CELLS
    OrbCand *cand = nullptr;                            // nlevels x ORB_KMAX
    OrbKp *kps = nullptr, *out = nullptr;               // nlevels x ORB_NODES ; ORB_OUT_MAX
    uint8_t *desc = nullptr;
    sindyn_keypoint *out_host_fmt = nullptr;
    OrbControl *ctl = nullptr, *ctl_host = nullptr;
    uint8_t *pin_gray = nullptr, *pin_mask = nullptr;   // pinned bounce buffers (see stage_in_2d)
    sindyn_keypoint *pin_kp = nullptr;
    uint8_t *pin_desc = nullptr;
    // Frame construction
    uint16_t *depth = nullptr, *pin_depth = nullptr;
    float *fr_un = nullptr, *fr_depth = nullptr, *fr_uright = nullptr, *fr_bounds = nullptr;
    int *fr_offsets = nullptr, *fr_indices = nullptr;
    struct MatchStage *match = nullptr;   // descriptor matching (row f4), allocated on first use
    void *track = nullptr;                // events of the fused per-frame entry (sindyn_track_frame)
    size_t pad_total = 0, img_total = 0;
};

// ------------------------------------------------------------------ pyramid
__global__ void k_orb_pad(uint8_t *__restrict__ pad, int w, int h, int pitch)
{
    // fill the 19-px frame of a padded level from its interior (BORDER_REFLECT_101)
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const int pw = w + 2 * ORB_EDGE, ph = h + 2 * ORB_EDGE;
    if (x >= pw || y >= ph) return;
    int sx = x - ORB_EDGE, sy = y - ORB_EDGE;
    if (sx >= 0 && sx < w && sy >= 0 && sy < h) return;
    sx = sx < 0 ? -sx : (sx >= w ? 2 * w - 2 - sx : sx);
    sy = sy < 0 ? -sy : (sy >= h ? 2 * h - 2 - sy : sy);
    pad[(size_t)y * pitch + x] = pad[(size_t)(sy + ORB_EDGE) * pitch + sx + ORB_EDGE];
}

// ------------------------------------------------------------------ FAST-9/16 score
__global__ void k_orb_fast_score(const uint8_t *__restrict__ pad, int w, int h, int pitch, uint8_t *__restrict__ score)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t *c = pad + (size_t)(y + ORB_EDGE) * pitch + x + ORB_EDGE;
    const int v = c[0];
    const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = v - (int)c[dy[k] * pitch + dx[k]];
    // NOTE: keep the bright and dark arc extrema in SEPARATE accumulators.  The obvious form
    //   best = max(best, max(mn, -mx))
    // is miscompiled by nvcc 12.9 for sm_100a (min/max/negate fusion into VIMNMX3; repro: tools/ptxas_minmax_repro.cu,
    // 97 % wrong scores on a B200) -- found through the bit-exact FAST parity test.
    int bb = -255, bd = 255;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
#pragma unroll
        for (int j = 1; j < 9; ++j) { int t = d[(k + j) & 15]; mn = min(mn, t); mx = max(mx, t); }
        bb = max(bb, mn);
        bd = min(bd, mx);
    }
    const int best = max(0, max(bb, -bd));
    score[(size_t)y * w + x] = (uint8_t)best;
}

// One CTA per cell window (ORBextractor.cc:787-828).  WRITE = false: count; WRITE = true: emit in raster order.
template <bool WRITE>
__global__ void __launch_bounds__(256) k_orb_cells(const uint8_t *__restrict__ score_all, const OrbLevel *__restrict__ lvs, int level, int ini_th,
                                                   int min_th, int *__restrict__ cell_count, const int *__restrict__ cell_offs,
                                                   OrbCand *__restrict__ cand_all, OrbControl *ctl)
{
    const OrbLevel L = lvs[level];
    const int j = blockIdx.x, i = blockIdx.y;
    const int cell = i * L.n_cols + j;
    int *my_count = cell_count + level * ORB_MAX_CELLS + cell;
    const float iniY = (float)(L.min_b + i * L.h_cell), iniX = (float)(L.min_b + j * L.w_cell);
    float maxY = iniY + (float)L.h_cell + 6.0f, maxX = iniX + (float)L.w_cell + 6.0f;
    const bool skip = iniY >= (float)(L.max_by - 3) || iniX >= (float)(L.max_bx - 6);
    if (skip) { if (!WRITE && threadIdx.x == 0) *my_count = 0; return; }
    if (maxY > (float)L.max_by) maxY = (float)L.max_by;
    if (maxX > (float)L.max_bx) maxX = (float)L.max_bx;
    const int x0 = (int)iniX, y0 = (int)iniY, x1 = (int)maxX, y1 = (int)maxY;
    const int ww = x1 - x0, wh = y1 - y0;          // window size (<= 64 x 64)
    __shared__ uint8_t s[66][68];
    __shared__ int s_cnt[2];
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const uint8_t *score = score_all + L.img_off;
    // scores of the window interior (3-px margin), zero elsewhere, with a 1-px zero ring for the NMS
    for (int t = threadIdx.x; t < (wh + 2) * (ww + 2); t += 256) {
        const int ty = t / (ww + 2), tx = t - ty * (ww + 2);
        const int lx = tx - 1, ly = ty - 1;
        uint8_t v = 0;
        if (lx >= 3 && lx < ww - 3 && ly >= 3 && ly < wh - 3) v = score[(size_t)(y0 + ly) * L.w + x0 + lx];
        s[ty][tx] = v;
    }
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    auto is_kp = [&](int lx, int ly, int th) -> bool {
        const int v = s[ly + 1][lx + 1];
        if (v <= th) return false;
        // strict maximum over the 8 neighbours; neighbours that are not corners at this threshold count as 0 < v
        return v > s[ly][lx] && v > s[ly][lx + 1] && v > s[ly][lx + 2] && v > s[ly + 1][lx] && v > s[ly + 1][lx + 2] &&
               v > s[ly + 2][lx] && v > s[ly + 2][lx + 1] && v > s[ly + 2][lx + 2];
    };
    const int npx = ww * wh;
    int c0 = 0, c1 = 0;
    for (int t = threadIdx.x; t < npx; t += 256) {
        const int ly = t / ww, lx = t - ly * ww;
        c0 += is_kp(lx, ly, ini_th);
        c1 += is_kp(lx, ly, min_th);
    }
    if (c0) atomicAdd(&s_cnt[0], c0);
    if (c1) atomicAdd(&s_cnt[1], c1);
    __syncthreads();
    const int th = s_cnt[0] > 0 ? ini_th : min_th;
    const int total = s_cnt[0] > 0 ? s_cnt[0] : s_cnt[1];
    if (!WRITE) { if (threadIdx.x == 0) *my_count = total; return; }
    if (total == 0) return;
    const int base0 = cell_offs[level * ORB_MAX_CELLS + cell];
    OrbCand *out = cand_all + (size_t)level * ORB_KMAX;
    if (threadIdx.x == 0) s_base = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t0 = 0; t0 < npx; t0 += 256) {
        __syncthreads();
        const int t = t0 + threadIdx.x;
        bool f = false;
        int lx = 0, ly = 0;
        if (t < npx) { ly = t / ww; lx = t - ly * ww; f = is_kp(lx, ly, th); }
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = 0, tile_total = 0;
        for (int k = 0; k < 8; ++k) { if (k < warp) before += s_warp[k]; tile_total += s_warp[k]; }
        const int pos = base0 + s_base + before + __popc(m & ((1u << lane) - 1u));
        if (f) {
            if (pos < ORB_KMAX) {
                OrbCand c;
                c.x = (unsigned short)(x0 + lx - L.min_b);
                c.y = (unsigned short)(y0 + ly - L.min_b);
                c.resp = (unsigned short)(s[ly + 1][lx + 1] - 1);
                c.pad = 0;
                out[pos] = c;
            } else ctl->overflow = 1;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += tile_total;
    }
}

// exclusive scan of the cell counts of every level (one CTA per level)
__global__ void k_orb_cell_scan(const OrbLevel *__restrict__ lvs, const int *__restrict__ cell_count, int *__restrict__ cell_offs, OrbControl *ctl)
{
    const int level = blockIdx.x;
    const OrbLevel L = lvs[level];
    const int n = L.n_cols * L.n_rows;
    __shared__ int s_part[256];
    const int per = (n + 255) / 256;
    int sum = 0;
    for (int k = 0; k < per; ++k) { int idx = threadIdx.x * per + k; if (idx < n) sum += cell_count[level * ORB_MAX_CELLS + idx]; }
    s_part[threadIdx.x] = sum;
    __syncthreads();
    int base = 0;
    for (int k = 0; k < threadIdx.x; ++k) base += s_part[k];
    for (int k = 0; k < per; ++k) {
        int idx = threadIdx.x * per + k;
        if (idx < n) { cell_offs[level * ORB_MAX_CELLS + idx] = base; base += cell_count[level * ORB_MAX_CELLS + idx]; }
    }
    if (threadIdx.x == 255) {
        ctl->cand_count[level] = base > ORB_KMAX ? ORB_KMAX : base;
        if (base > ORB_KMAX) ctl->overflow = 1;
    }
}

// ------------------------------------------------------------------ quadtree (DistributeOctTree)
struct QNode { short x0, y0, x1, y1; unsigned short prev, next, cbase; unsigned short flags; int count; };
#define QN_NONE 0xFFFFu
#define QF_NOMORE 1u
#define QF_MARK 2u      // considered for division in the current batch
#define QF_DIVIDED 4u   // actually divided in the current batch

struct QState {
    int head, tail, size, n_nodes, overflow;
    int n_batch;
};

__device__ __forceinline__ int q_quadrant(const QNode &p, int kx, int ky)
{
    // DivideNode (ORBextractor.cc:481-537): halfX = ceil((UR.x - UL.x) / 2), halfY = ceil((BR.y - UL.y) / 2)
    const int mx = p.x0 + (int)ceilf((float)(p.x1 - p.x0) / 2.0f), my = p.y0 + (int)ceilf((float)(p.y1 - p.y0) / 2.0f);
    return kx < mx ? (ky < my ? 0 : 2) : (ky < my ? 1 : 3);
}

__global__ void __launch_bounds__(512) k_orb_quadtree(const OrbLevel *__restrict__ lvs, const OrbCand *__restrict__ cand_all, OrbControl *ctl,
                                                      OrbKp *__restrict__ kps_all)
{
    extern __shared__ unsigned char smem_raw[];
    const int level = blockIdx.x;
    const OrbLevel L = lvs[level];
    const int K = ctl->cand_count[level];
    const int N = L.quota;
    QNode *nodes = (QNode *)smem_raw;                                   // ORB_NODES
    unsigned short *knode = (unsigned short *)(nodes + ORB_NODES);      // ORB_KMAX
    unsigned int *kxy = (unsigned int *)(knode + ORB_KMAX);             // ORB_KMAX  (y << 16 | x)
    unsigned int *best = (unsigned int *)(kxy + ORB_KMAX);              // ORB_NODES
    unsigned short *batch = (unsigned short *)(best + ORB_NODES);       // ORB_NODES: vSizeAndPointerToNode
    unsigned short *batch2 = batch + ORB_NODES;
    __shared__ QState st;
    const OrbCand *cand = cand_all + (size_t)level * ORB_KMAX;
    const int tid = threadIdx.x, nt = blockDim.x;
    OrbKp *kps = kps_all + (size_t)level * ORB_NODES;

    // ---- initial nodes (ORBextractor.cc:543-584)
    const int dx = L.max_bx - L.min_b, dy = L.max_by - L.min_b;
    const int nIni = (int)roundf((float)dx / (float)dy);
    const float hX = (float)dx / (float)nIni;
    if (tid == 0) {
        st.head = QN_NONE; st.tail = QN_NONE; st.size = 0; st.n_nodes = nIni; st.overflow = 0; st.n_batch = 0;
        for (int i = 0; i < nIni; ++i) {
            QNode n;
            n.x0 = (short)(int)(hX * (float)i); n.x1 = (short)(int)(hX * (float)(i + 1)); n.y0 = 0; n.y1 = (short)dy;
            n.prev = n.next = QN_NONE; n.cbase = 0; n.flags = 0; n.count = 0;
            nodes[i] = n;
        }
    }
    __syncthreads();
    for (int k = tid; k < K; k += nt) {
        const OrbCand c = cand[k];
        int ni = (int)((float)c.x / hX);
        ni = ni < nIni ? ni : nIni - 1;
        knode[k] = (unsigned short)ni;
        kxy[k] = ((unsigned int)c.y << 16) | c.x;
        atomicAdd(&nodes[ni].count, 1);
    }
    __syncthreads();
    if (tid == 0) {
        for (int i = 0; i < nIni; ++i) {             // push_back order; empty nodes are erased
            if (nodes[i].count == 0) continue;
            if (nodes[i].count == 1) nodes[i].flags = QF_NOMORE;
            nodes[i].prev = (unsigned short)st.tail; nodes[i].next = QN_NONE;
            if (st.tail != (int)QN_NONE) nodes[st.tail].next = (unsigned short)i; else st.head = i;
            st.tail = i;
            ++st.size;
        }
    }
    __syncthreads();

    // list helpers (thread 0 only)
    auto push_front = [&](int id) {
        nodes[id].prev = QN_NONE; nodes[id].next = (unsigned short)st.head;
        if (st.head != (int)QN_NONE) nodes[st.head].prev = (unsigned short)id; else st.tail = id;
        st.head = id; ++st.size;
    };
    auto erase = [&](int id) {
        const int p = nodes[id].prev, n = nodes[id].next;
        if (p != (int)QN_NONE) nodes[p].next = (unsigned short)n; else st.head = n;
        if (n != (int)QN_NONE) nodes[n].prev = (unsigned short)p; else st.tail = p;
        --st.size;
    };
    // allocate the four children of `id` (geometry only; counts are filled by the parallel pass)
    auto alloc_children = [&](int id) -> bool {
        if (st.n_nodes + 4 > ORB_NODES) { st.overflow = 1; return false; }
        const QNode p = nodes[id];
        const int hx = (int)ceilf((float)(p.x1 - p.x0) / 2.0f), hy = (int)ceilf((float)(p.y1 - p.y0) / 2.0f);
        const int cb = st.n_nodes;
        st.n_nodes += 4;
        for (int q = 0; q < 4; ++q) {
            QNode c;
            c.x0 = (short)((q & 1) ? p.x0 + hx : p.x0); c.x1 = (short)((q & 1) ? p.x1 : p.x0 + hx);
            c.y0 = (short)((q & 2) ? p.y0 + hy : p.y0); c.y1 = (short)((q & 2) ? p.y1 : p.y0 + hy);
            c.prev = c.next = QN_NONE; c.cbase = 0; c.flags = 0; c.count = 0;
            nodes[cb + q] = c;
        }
        nodes[id].cbase = (unsigned short)cb;
        nodes[id].flags |= QF_MARK;
        return true;
    };
    // parallel passes over the keys
    auto count_children = [&]() {
        for (int k = tid; k < K; k += nt) {
            const QNode p = nodes[knode[k]];
            if (p.flags & QF_MARK) atomicAdd(&nodes[p.cbase + q_quadrant(p, kxy[k] & 0xffff, kxy[k] >> 16)].count, 1);
        }
    };
    auto move_keys = [&]() {
        for (int k = tid; k < K; k += nt) {
            const QNode p = nodes[knode[k]];
            if (p.flags & QF_DIVIDED) knode[k] = (unsigned short)(p.cbase + q_quadrant(p, kxy[k] & 0xffff, kxy[k] >> 16));
        }
    };

    __shared__ int s_finish, s_phase2, s_nexp, s_prev;
    if (tid == 0) { s_finish = 0; s_phase2 = 0; }
    __syncthreads();
    while (!s_finish && !st.overflow) {
        // ---- phase 1 pass: divide every node that holds more than one key (ORBextractor.cc:603-663)
        if (tid == 0) {
            s_prev = st.size;
            for (int id = st.head; id != (int)QN_NONE; id = nodes[id].next)
                if (!(nodes[id].flags & QF_NOMORE)) { if (!alloc_children(id)) break; }
        }
        __syncthreads();
        count_children();
        __syncthreads();
        if (tid == 0) {
            int nexp = 0, nb = 0;
            int id = st.head;
            while (id != (int)QN_NONE) {
                const int nx = nodes[id].next;
                if (nodes[id].flags & QF_MARK) {
                    const int cb = nodes[id].cbase;
                    for (int q = 0; q < 4; ++q) {
                        const int c = cb + q;
                        if (nodes[c].count > 0) {
                            push_front(c);
                            if (nodes[c].count > 1) { ++nexp; batch[nb++] = (unsigned short)c; }
                            else nodes[c].flags = QF_NOMORE;
                        }
                    }
                    erase(id);
                    nodes[id].flags = (nodes[id].flags & ~QF_MARK) | QF_DIVIDED;
                }
                id = nx;
            }
            s_nexp = nexp; st.n_batch = nb;
        }
        __syncthreads();
        move_keys();
        __syncthreads();
        for (int i = tid; i < st.n_nodes; i += nt) nodes[i].flags &= ~(QF_MARK | QF_DIVIDED);
        if (tid == 0) {
            if (st.size >= N || st.size == s_prev) s_finish = 1;
            else if (st.size + s_nexp * 3 > N) s_phase2 = 1;
        }
        __syncthreads();
        // ---- phase 2: expand the largest nodes first until N is reached (ORBextractor.cc:671-735)
        while (s_phase2 && !s_finish && !st.overflow) {
            const int nb = st.n_batch;
            // sort (size, creation order) ascending, walk from the end = descending
            for (int a = tid; a < nb; a += nt) {
                const int ia = batch[a];
                const long long ka = ((long long)nodes[ia].count << 16) | ia;
                int rank = 0;
                for (int b = 0; b < nb; ++b) { const int ib = batch[b]; rank += ((((long long)nodes[ib].count << 16) | ib) > ka); }
                batch2[rank] = (unsigned short)ia;
            }
            __syncthreads();
            if (tid == 0) {
                s_prev = st.size;
                for (int a = 0; a < nb; ++a) if (!alloc_children(batch2[a])) break;
            }
            __syncthreads();
            count_children();
            __syncthreads();
            if (tid == 0) {
                int nb2 = 0;
                for (int a = 0; a < nb; ++a) {
                    const int id = batch2[a];
                    if (!(nodes[id].flags & QF_MARK)) break;
                    const int cb = nodes[id].cbase;
                    for (int q = 0; q < 4; ++q) {
                        const int c = cb + q;
                        if (nodes[c].count > 0) {
                            push_front(c);
                            if (nodes[c].count > 1) batch[nb2++] = (unsigned short)c;
                            else nodes[c].flags = QF_NOMORE;
                        }
                    }
                    erase(id);
                    nodes[id].flags |= QF_DIVIDED;
                    if (st.size >= N) break;
                }
                st.n_batch = nb2;
            }
            __syncthreads();
            move_keys();
            __syncthreads();
            for (int i = tid; i < st.n_nodes; i += nt) nodes[i].flags &= ~(QF_MARK | QF_DIVIDED);
            if (tid == 0 && (st.size >= N || st.size == s_prev)) s_finish = 1;
            __syncthreads();
        }
    }
    // ---- best response per node, first maximum in candidate order (ORBextractor.cc:738-760)
    for (int i = tid; i < st.n_nodes; i += nt) best[i] = 0u;
    __syncthreads();
    for (int k = tid; k < K; k += nt) atomicMax(&best[knode[k]], ((unsigned int)cand[k].resp << 14) | (unsigned int)(ORB_KMAX - 1 - k));
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int id = st.head; id != (int)QN_NONE; id = nodes[id].next) batch[n++] = (unsigned short)id;
        ctl->kp_count[level] = n;
        if (st.overflow) ctl->overflow = 1;
        st.n_batch = n;
    }
    __syncthreads();
    for (int i = tid; i < st.n_batch; i += nt) {
        const int k = ORB_KMAX - 1 - (int)(best[batch[i]] & (ORB_KMAX - 1));
        const OrbCand c = cand[k];
        OrbKp o;
        o.x = (float)(c.x + L.min_b); o.y = (float)(c.y + L.min_b);   // keypoints[i].pt += minBorder (ORBextractor.cc:842-843)
        o.angle = 0.f; o.resp = (float)c.resp; o.level = level; o.keep = 1;
        kps[i] = o;
    }
}

// ------------------------------------------------------------------ orientation (IC_Angle) + mask test
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    // cv::fastAtan2
    const float p1 = 0.9997878412794807f * 57.29577951308232f, p3 = -0.3258083974640975f * 57.29577951308232f;
    const float p5 = 0.1555786518463281f * 57.29577951308232f, p7 = -0.04432655554792128f * 57.29577951308232f;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + 2.220446049250313e-16f);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + 2.220446049250313e-16f);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

__global__ void k_orb_orient(const uint8_t *__restrict__ pyr, const OrbLevel *__restrict__ lvs, const OrbControl *__restrict__ ctl,
                             OrbKp *__restrict__ kps_all, const uint8_t *__restrict__ mask, int mask_w)
{
    const int level = blockIdx.y;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= ctl->kp_count[level]) return;
    const OrbLevel L = lvs[level];
    OrbKp *kp = kps_all + (size_t)level * ORB_NODES + warp;
    const int cx = __float2int_rn(kp->x), cy = __float2int_rn(kp->y);
    const uint8_t *center = pyr + L.pad_off + (size_t)(cy + ORB_EDGE) * L.pitch + cx + ORB_EDGE;
    int m01 = 0, m10 = 0;
    const int u = lane - ORB_HALF_PATCH;
    if (lane < 31) {
        m10 = u * (int)center[u];
        for (int v = 1; v <= ORB_HALF_PATCH; ++v) {
            if (u >= -c_umax[v] && u <= c_umax[v]) {
                const int vp = center[u + v * L.pitch], vm = center[u - v * L.pitch];
                m01 += v * (vp - vm);
                m10 += u * (vp + vm);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) { m01 += __shfl_xor_sync(0xffffffffu, m01, o); m10 += __shfl_xor_sync(0xffffffffu, m10, o); }
    if (lane == 0) {
        kp->angle = fast_atan2_deg((float)m01, (float)m10);
        int keep = 1;
        if (mask) {   // DynaMask.at<uchar>(pt.y * scale, pt.x * scale) == 255 (ORBextractor.cc:1073-1075)
            const int my = (int)(kp->y * L.mask_scale), mx = (int)(kp->x * L.mask_scale);
            keep = mask[(size_t)my * mask_w + mx] != 255;
        }
        kp->keep = keep;
    }
}

// erase + "< 250 -> restore" + level-ordered output (ORBextractor.cc:1063-1116,1154-1162); single CTA
__global__ void __launch_bounds__(1024) k_orb_select(const OrbLevel *__restrict__ lvs, int nlevels, OrbControl *ctl, const OrbKp *__restrict__ kps_all,
                                                     OrbKp *__restrict__ out)
{
    __shared__ int s_kept, s_total, s_base;
    __shared__ int s_warp[32];
    if (threadIdx.x == 0) { s_kept = 0; s_total = 0; s_base = 0; }
    __syncthreads();
    int kept = 0, total = 0;
    for (int l = 0; l < nlevels; ++l) {
        const int n = ctl->kp_count[l];
        for (int i = threadIdx.x; i < n; i += blockDim.x) { kept += kps_all[(size_t)l * ORB_NODES + i].keep; ++total; }
    }
    if (kept) atomicAdd(&s_kept, kept);
    if (total) atomicAdd(&s_total, total);
    __syncthreads();
    const bool all = s_kept < 250;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int l = 0; l < nlevels; ++l) {
        const int n = ctl->kp_count[l];
        const float sc = lvs[l].scale;
        for (int t0 = 0; t0 < n; t0 += blockDim.x) {
            __syncthreads();
            const int i = t0 + threadIdx.x;
            OrbKp k;
            bool f = false;
            if (i < n) { k = kps_all[(size_t)l * ORB_NODES + i]; f = all || k.keep; }
            const unsigned m = __ballot_sync(0xffffffffu, f);
            if (lane == 0) s_warp[warp] = __popc(m);
            __syncthreads();
            int before = 0, tile = 0;
            for (int w = 0; w < 32; ++w) { if (w < warp) before += s_warp[w]; tile += s_warp[w]; }
            const int pos = s_base + before + __popc(m & ((1u << lane) - 1u));
            if (f) { if (pos < ORB_OUT_MAX) out[pos] = k; else ctl->overflow = 1; }
            (void)sc;
            __syncthreads();
            if (threadIdx.x == 0) s_base += tile;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) { ctl->n_out = s_base > ORB_OUT_MAX ? ORB_OUT_MAX : s_base; ctl->n_total = s_total; }
}

// ------------------------------------------------------------------ blur + descriptors
// GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on u8: kernel [18 34 48 56 48 34 18]/256,
// horizontal 8.8 fixed point, vertical 16.16, (v + 32768) >> 16.  Reads the padded level (its frame IS the reflection).
__global__ void k_orb_blur(const uint8_t *__restrict__ pad, int w, int h, int pitch, uint8_t *__restrict__ dst)
{
    __shared__ int hs[8 + 6][32];
    const int kx[7] = {18, 34, 48, 56, 48, 34, 18};
    const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 8;
    for (int r = threadIdx.y; r < 14; r += 8) {
        const int y = y0 + r - 3;
        int acc = 0;
        if (x < w && y >= -ORB_EDGE && y < h + ORB_EDGE) {
            const uint8_t *row = pad + (size_t)(y + ORB_EDGE) * pitch + x + ORB_EDGE;
#pragma unroll
            for (int k = 0; k < 7; ++k) acc += kx[k] * (int)row[k - 3];
        }
        hs[r][threadIdx.x] = acc;
    }
    __syncthreads();
    const int y = y0 + threadIdx.y;
    if (x >= w || y >= h) return;
    int v = 0;
#pragma unroll
    for (int k = 0; k < 7; ++k) v += kx[k] * hs[threadIdx.y + k][threadIdx.x];
    dst[(size_t)y * w + x] = (uint8_t)((v + 32768) >> 16);
}

// computeOrbDescriptor (ORBextractor.cc:108-147): one warp per keypoint, lane = descriptor byte
__global__ void k_orb_desc(const uint8_t *__restrict__ blur, const OrbLevel *__restrict__ lvs, const OrbControl *__restrict__ ctl,
                           const OrbKp *__restrict__ out, uint8_t *__restrict__ desc, sindyn_keypoint *__restrict__ kp_out)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= ctl->n_out) return;
    const OrbKp k = out[warp];
    const OrbLevel L = lvs[k.level];
    const float angle = k.angle * (float)(3.14159265358979323846 / 180.f);
    const float a = (float)cos((double)angle), b = (float)sin((double)angle);
    const uint8_t *img = blur + L.img_off;
    const int cx = __float2int_rn(k.x), cy = __float2int_rn(k.y);
    int val = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const signed char *p = c_orb_pattern + (lane * 8 + t) * 4;
        const float x0 = (float)p[0], y0 = (float)p[1], x1 = (float)p[2], y1 = (float)p[3];
        const int r0 = __float2int_rn(x0 * b + y0 * a), c0 = __float2int_rn(x0 * a - y0 * b);
        const int r1 = __float2int_rn(x1 * b + y1 * a), c1 = __float2int_rn(x1 * a - y1 * b);
        const int t0 = img[(size_t)(cy + r0) * L.w + cx + c0], t1 = img[(size_t)(cy + r1) * L.w + cx + c1];
        val |= (t0 < t1) << t;
    }
    desc[(size_t)warp * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        sindyn_keypoint o;
        o.x = k.level ? k.x * L.scale : k.x;     // keypoint->pt *= scale for level != 0 (ORBextractor.cc:1154-1160)
        o.y = k.level ? k.y * L.scale : k.y;
        o.size = L.size; o.angle = k.angle; o.response = k.resp; o.octave = k.level;
        kp_out[warp] = o;
    }
}

// ------------------------------------------------------------------ host side
static inline int cv_round_f(float v) { return (int)lrintf(v); }

extern "C" int sindyn_orb_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast, int width, int height,
                                 int device, sindyn_orb_handle *out)
{
    if (!out || nlevels < 1 || nlevels > ORB_MAX_LEVELS || nfeatures < 1 || width < 64 || height < 64 || scale_factor <= 1.0f) return SINDYN_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device >= ndev) return SINDYN_ERR_NO_DEVICE;
    sindyn_orb *o = new sindyn_orb();
    *out = o;
    o->device = device;
    o->nfeatures = nfeatures; o->nlevels = nlevels; o->ini_th = ini_th_fast; o->min_th = min_th_fast; o->W = width; o->H = height;
    o->scale_factor = scale_factor;
    if (cudaSetDevice(device) != cudaSuccess) { o->err = "cudaSetDevice failed"; return SINDYN_ERR_CUDA; }
    CU_CHECK(o, cudaStreamCreateWithFlags(&o->own_stream, cudaStreamNonBlocking));
    o->stream = o->own_stream;
    // scale tables and feature quotas (ORBextractor.cc:415-447)
    float sf[ORB_MAX_LEVELS];
    sf[0] = 1.0f;
    for (int i = 1; i < nlevels; ++i) sf[i] = sf[i - 1] * scale_factor;
    const float factor = (float)(1.0 / (double)scale_factor);   // the reference member is a double holding the float
    float nd = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    int quota[ORB_MAX_LEVELS];
    for (int l = 0; l < nlevels - 1; ++l) { quota[l] = cv_round_f(nd); sum += quota[l]; nd *= factor; }
    quota[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    size_t pad_off = 0, img_off = 0;
    int cell_off = 0;
    for (int l = 0; l < nlevels; ++l) {
        OrbLevel &L = o->lv[l];
        const float inv = 1.0f / sf[l];
        L.w = cv_round_f((float)width * inv); L.h = cv_round_f((float)height * inv);    // ORBextractor.cc:1171
        L.pitch = L.w + 2 * ORB_EDGE;
        L.pad_off = pad_off; pad_off += (size_t)L.pitch * (L.h + 2 * ORB_EDGE);
        L.img_off = img_off; img_off += (size_t)L.w * L.h;
        L.min_b = ORB_EDGE - 3; L.max_bx = L.w - ORB_EDGE + 3; L.max_by = L.h - ORB_EDGE + 3;
        const float fw = (float)(L.max_bx - L.min_b), fh = (float)(L.max_by - L.min_b);
        L.n_cols = (int)(fw / 30.0f); L.n_rows = (int)(fh / 30.0f);                       // ORBextractor.cc:769-785
        if (L.n_cols < 1 || L.n_rows < 1) { o->err = "orb: pyramid level too small"; return SINDYN_ERR_INVALID; }
        L.w_cell = (int)ceilf(fw / (float)L.n_cols); L.h_cell = (int)ceilf(fh / (float)L.n_rows);
        if (L.n_cols * L.n_rows > ORB_MAX_CELLS || L.w_cell + 6 > 64 || L.h_cell + 6 > 64) { o->err = "orb: cell grid out of range"; return SINDYN_ERR_INVALID; }
        L.cell_off = cell_off; cell_off += L.n_cols * L.n_rows;
        L.quota = quota[l];
        if (L.quota + 8 > ORB_NODES / 4) { o->err = "orb: nfeatures too large for the node pool"; return SINDYN_ERR_INVALID; }
        L.scale = sf[l];
        L.mask_scale = (float)pow((double)scale_factor, (double)l);                       // ORBextractor.cc:1074
        L.size = (float)(int)(31.0f * sf[l]);                                             // scaledPatchSize (:837)
    }
    o->pad_total = pad_off; o->img_total = img_off;
    SD_CHECK(o->dalloc(&o->pyr, pad_off));
    SD_CHECK(o->dalloc(&o->score, img_off));
    SD_CHECK(o->dalloc(&o->blur, img_off));
    SD_CHECK(o->dalloc(&o->mask, (size_t)width * height));
    SD_CHECK(o->dalloc(&o->lv_dev, ORB_MAX_LEVELS));
    SD_CHECK(o->dalloc(&o->cell_count, (size_t)nlevels * ORB_MAX_CELLS));
    SD_CHECK(o->dalloc(&o->cell_offs, (size_t)nlevels * ORB_MAX_CELLS));
    SD_CHECK(o->dalloc(&o->cand, (size_t)nlevels * ORB_KMAX));
    SD_CHECK(o->dalloc(&o->kps, (size_t)nlevels * ORB_NODES));
    SD_CHECK(o->dalloc(&o->out, ORB_OUT_MAX));
    SD_CHECK(o->dalloc(&o->desc, (size_t)ORB_OUT_MAX * 32));
    SD_CHECK(o->dalloc(&o->out_host_fmt, ORB_OUT_MAX));
    SD_CHECK(o->dalloc(&o->ctl, 1));
    SD_CHECK(o->halloc(&o->ctl_host, 1));
    SD_CHECK(o->halloc(&o->pin_gray, (size_t)width * height));
    SD_CHECK(o->halloc(&o->pin_mask, (size_t)width * height));
    SD_CHECK(o->halloc(&o->pin_kp, ORB_OUT_MAX));
    SD_CHECK(o->halloc(&o->pin_desc, (size_t)ORB_OUT_MAX * 32));
    SD_CHECK(o->dalloc(&o->depth, (size_t)width * height));
    SD_CHECK(o->halloc(&o->pin_depth, (size_t)width * height));
    SD_CHECK(o->dalloc(&o->fr_un, 2 * ORB_OUT_MAX));
    SD_CHECK(o->dalloc(&o->fr_depth, ORB_OUT_MAX));
    SD_CHECK(o->dalloc(&o->fr_uright, ORB_OUT_MAX));
    SD_CHECK(o->dalloc(&o->fr_bounds, 4));
    SD_CHECK(o->dalloc(&o->fr_offsets, 64 * 48 + 1));
    SD_CHECK(o->dalloc(&o->fr_indices, ORB_OUT_MAX));
    CU_CHECK(o, cudaMemcpyAsync(o->lv_dev, o->lv, sizeof(OrbLevel) * ORB_MAX_LEVELS, cudaMemcpyHostToDevice, o->stream));
    for (int l = 1; l < nlevels; ++l) SD_CHECK(resize_plan_init(o, &o->plan[l], o->lv[l - 1].w, o->lv[l - 1].h, o->lv[l].w, o->lv[l].h));
    // umax (ORBextractor.cc:450-467)
    int umax[ORB_HALF_PATCH + 1];
    {
        int v, v0;
        const int vmax = (int)floorf(ORB_HALF_PATCH * sqrtf(2.f) / 2 + 1), vmin = (int)ceilf(ORB_HALF_PATCH * sqrtf(2.f) / 2);
        const double hp2 = ORB_HALF_PATCH * ORB_HALF_PATCH;
        for (v = 0; v <= vmax; ++v) umax[v] = (int)lrint(sqrt(hp2 - v * v));
        for (v = ORB_HALF_PATCH, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
    }
    CU_CHECK(o, cudaMemcpyToSymbolAsync(c_umax, umax, sizeof umax, 0, cudaMemcpyHostToDevice, o->stream));
    const size_t qsmem = sizeof(QNode) * ORB_NODES + 2 * ORB_KMAX + 4 * ORB_KMAX + 4 * ORB_NODES + 2 * 2 * ORB_NODES;
    CU_CHECK(o, cudaFuncSetAttribute(k_orb_quadtree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qsmem));
    CU_CHECK(o, cudaStreamSynchronize(o->stream));
    return SINDYN_OK;
}

static void match_stage_free(sindyn_orb *o);

static void track_free(sindyn_orb *o);
extern "C" int sindyn_orb_destroy(sindyn_orb_handle h)
{
    if (!h) return SINDYN_ERR_INVALID;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    match_stage_free(h);
    track_free(h);
    h->free_all();
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return SINDYN_OK;
}

extern "C" int sindyn_orb_set_stream(sindyn_orb_handle h, void *s)
{
    if (!h) return SINDYN_ERR_INVALID;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    h->stream = s ? (cudaStream_t)s : h->own_stream;
    return SINDYN_OK;
}

extern "C" unsigned long long sindyn_orb_launch_count(sindyn_orb_handle h) { return h ? h->launches : 0ull; }
extern "C" const char *sindyn_orb_last_error(sindyn_orb_handle h) { return h ? h->err.c_str() : "null handle"; }

// first half of one extraction (pyramid, FAST, cells, quadtree): needs the gray image only (level-0 interior of pyr)
static int orb_enqueue_unmasked(sindyn_orb *o)
{
    const dim3 blk(32, 8);
    const size_t qsmem = sizeof(QNode) * ORB_NODES + 2 * ORB_KMAX + 4 * ORB_KMAX + 4 * ORB_NODES + 2 * 2 * ORB_NODES;
    CU_CHECK(o, cudaMemsetAsync(o->ctl, 0, sizeof(OrbControl), o->stream));
    for (int l = 0; l < o->nlevels; ++l) {
        const OrbLevel &L = o->lv[l];
        uint8_t *pad = o->pyr + L.pad_off;
        if (l > 0) {
            const OrbLevel &P = o->lv[l - 1];
            SD_CHECK(launch_resize_u8(o, &o->plan[l], o->pyr + P.pad_off + (size_t)ORB_EDGE * P.pitch + ORB_EDGE, P.pitch,
                                      pad + (size_t)ORB_EDGE * L.pitch + ORB_EDGE, L.pitch, nullptr, 0.f));
        }
        LAUNCH(o, k_orb_pad, dim3(cdiv(L.pitch, 32), cdiv(L.h + 2 * ORB_EDGE, 8)), blk, 0, pad, L.w, L.h, L.pitch);
        LAUNCH(o, k_orb_fast_score, dim3(cdiv(L.w, 32), cdiv(L.h, 8)), blk, 0, pad, L.w, L.h, L.pitch, o->score + L.img_off);
        LAUNCH(o, k_orb_blur, dim3(cdiv(L.w, 32), cdiv(L.h, 8)), blk, 0, pad, L.w, L.h, L.pitch, o->blur + L.img_off);
        LAUNCH(o, (k_orb_cells<false>), dim3(L.n_cols, L.n_rows), 256, 0, o->score, o->lv_dev, l, o->ini_th, o->min_th, o->cell_count,
               o->cell_offs, o->cand, o->ctl);
    }
    LAUNCH(o, k_orb_cell_scan, o->nlevels, 256, 0, o->lv_dev, o->cell_count, o->cell_offs, o->ctl);
    for (int l = 0; l < o->nlevels; ++l) {
        const OrbLevel &L = o->lv[l];
        LAUNCH(o, (k_orb_cells<true>), dim3(L.n_cols, L.n_rows), 256, 0, o->score, o->lv_dev, l, o->ini_th, o->min_th, o->cell_count,
               o->cell_offs, o->cand, o->ctl);
    }
    LAUNCH(o, k_orb_quadtree, o->nlevels, 512, qsmem, o->lv_dev, o->cand, o->ctl, o->kps);
    LAUNCH_CHECK(o);
    return SINDYN_OK;
}

// second half: everything that needs the dynamic mask (erasure inside k_orb_orient) and follows it
static int orb_enqueue_masked(sindyn_orb *o, bool have_mask)
{
    LAUNCH(o, k_orb_orient, dim3(cdiv(ORB_NODES * 32, 256), o->nlevels), 256, 0, o->pyr, o->lv_dev, o->ctl, o->kps,
           have_mask ? o->mask : (const uint8_t *)nullptr, o->W);
    LAUNCH(o, k_orb_select, 1, 1024, 0, o->lv_dev, o->nlevels, o->ctl, o->kps, o->out);
    LAUNCH(o, k_orb_desc, cdiv(ORB_OUT_MAX * 32, 256), 256, 0, o->blur, o->lv_dev, o->ctl, o->out, o->desc, o->out_host_fmt);
    LAUNCH_CHECK(o);
    return SINDYN_OK;
}

extern "C" int sindyn_orb_extract(sindyn_orb_handle h, const uint8_t *gray, size_t gray_step, const uint8_t *mask, size_t mask_step,
                                  sindyn_keypoint *kps, uint8_t *desc, int capacity, int *n_out)
{
    if (!h || !n_out) return SINDYN_ERR_INVALID;
    cudaSetDevice(h->device);
    *n_out = 0;
    if (!gray) return SINDYN_OK;   // operator() returns silently on an empty image (ORBextractor.cc:1046-1047)
    const OrbLevel &L0 = h->lv[0];
    {   // gray -> interior of the padded level 0 (pitched destination), through the pinned bounce buffer
        const void *src = gray;
        size_t step = gray_step ? gray_step : (size_t)h->W;
        if (!host_ptr_is_pinned(gray)) {
            for (int r = 0; r < h->H; ++r) memcpy(h->pin_gray + (size_t)r * h->W, gray + r * step, h->W);
            src = h->pin_gray; step = h->W;
        }
        CU_CHECK(h, cudaMemcpy2DAsync(h->pyr + (size_t)ORB_EDGE * L0.pitch + ORB_EDGE, L0.pitch, src, step, h->W, h->H, cudaMemcpyHostToDevice, h->stream));
    }
    if (mask) CU_CHECK(h, stage_in_2d(h->mask, mask, mask_step, h->W, h->H, h->pin_mask, h->stream));
    SD_CHECK(orb_enqueue_unmasked(h));
    SD_CHECK(orb_enqueue_masked(h, mask != nullptr));
    CU_CHECK(h, cudaMemcpyAsync(h->ctl_host, h->ctl, sizeof(OrbControl), cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    if (h->ctl_host->overflow) { h->err = "orb: a fixed-capacity list overflowed (candidates / nodes / output)"; return SINDYN_ERR_CAPACITY; }
    const int n = h->ctl_host->n_out;
    *n_out = n;
    if (n > capacity) { h->err = "orb: output capacity too small"; return SINDYN_ERR_CAPACITY; }
    if (n && kps) CU_CHECK(h, cudaMemcpyAsync(h->pin_kp, h->out_host_fmt, sizeof(sindyn_keypoint) * n, cudaMemcpyDeviceToHost, h->stream));
    if (n && desc) CU_CHECK(h, cudaMemcpyAsync(h->pin_desc, h->desc, (size_t)32 * n, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    if (n && kps) memcpy(kps, h->pin_kp, sizeof(sindyn_keypoint) * n);
    if (n && desc) memcpy(desc, h->pin_desc, (size_t)32 * n);
    return SINDYN_OK;
}

// ------------------------------------------------------------------ one driver iteration: DetectDynaArea + dilation + masked ORB
// rgbd_tum_noros.cc:132-139 (DetectDynaArea, 15x15 dilation of the mask) followed by what System::TrackRGBD does with the frame
// up to the key points: Tracking::GrabImageRGBD's colour conversion (Tracking.cc:246-252: RGB2GRAY when Camera.RGB is set,
// else BGR2GRAY) and Frame::ExtractORB2 -> ORBextractor::operator() on the dilated mask (Frame.cc:300-317).
// The ORB pyramid / FAST / quadtree half needs the frame only, so it runs on the extractor's stream concurrently with the
// detection; the erasure half waits for the dilated mask.
__global__ void k_orb_gray_in(const uint8_t *__restrict__ bgr, int W, int H, int rgb_order, uint8_t *__restrict__ dst, int pitch)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const uint8_t *p = bgr + 3 * ((size_t)y * W + x);
    // cvtColor: gray = (c0 * 3735 + c1 * 19235 + c2 * 9798 + 16384) >> 15 with (c0, c2) = (B, R); RGB2GRAY reads the same bytes
    // as (R, G, B), i.e. the outer weights swap
    const int w0 = rgb_order ? 9798 : 3735, w2 = rgb_order ? 3735 : 9798;
    dst[(size_t)y * pitch + x] = (uint8_t)((p[0] * w0 + p[1] * 19235 + p[2] * w2 + 16384) >> 15);
}

int detect_run_public(sindyn_ctx *c);        // detect.cu
int detect_check_capacity(sindyn_ctx *c);    // detect.cu (synchronises the handle's stream)

#define TRACK_NB 3      // frames in flight through sindyn_track_submit (SINDYN_TRACK_MAX_IN_FLIGHT)
struct TrackSync {
    cudaEvent_t ev_in = nullptr, ev_mask = nullptr, ev_orb_done = nullptr;
    // asynchronous entry (sindyn_track_submit / sindyn_track_collect): results of the frames in flight, pinned, ring of TRACK_NB
    uint8_t *r_mask[TRACK_NB] = {}, *r_label[TRACK_NB] = {}, *r_desc[TRACK_NB] = {};
    sindyn_keypoint *r_kp[TRACK_NB] = {};
    OrbControl *r_ctl[TRACK_NB] = {};
    PipeFlags *r_flags[TRACK_NB] = {};
    cudaEvent_t ev_res_main[TRACK_NB] = {}, ev_res_orb[TRACK_NB] = {};
    int r_cap = 0;
    unsigned long long n_submit = 0, n_collect = 0;
};
static TrackSync *track_sync(sindyn_orb *o)
{
    if (!o->track) {
        TrackSync *t = new TrackSync();
        cudaEventCreateWithFlags(&t->ev_in, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&t->ev_mask, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&t->ev_orb_done, cudaEventDisableTiming);
        cudaEventRecord(t->ev_orb_done, o->stream);
        o->track = t;
    }
    return (TrackSync *)o->track;
}

static void track_free(sindyn_orb *o)
{
    TrackSync *t = (TrackSync *)o->track;
    if (!t) return;
    cudaEventDestroy(t->ev_in); cudaEventDestroy(t->ev_mask); cudaEventDestroy(t->ev_orb_done);
    for (int p = 0; p < TRACK_NB; ++p) {
        if (t->ev_res_main[p]) cudaEventDestroy(t->ev_res_main[p]);
        if (t->ev_res_orb[p]) cudaEventDestroy(t->ev_res_orb[p]);
    }
    delete t;
    o->track = nullptr;
}

// inputs already in c->bgr[c->i_cur] / c->depth (enqueued on c->stream).  Leaves the dilated mask in o->mask, labels in
// c->rc.label_out, key points / descriptors in o->out_host_fmt / o->desc; nothing is synchronised.
static int track_enqueue(sindyn_ctx *c, sindyn_orb *o, int rgb_order, int dilate_k)
{
    if (o->device != c->device || o->W != c->W || o->H != c->H) { c->err = "track_frame: the two handles differ in device or image size"; return SINDYN_ERR_INVALID; }
    if (dilate_k < 0 || dilate_k > MORPH_MAX_K) { c->err = "track_frame: dilate_k out of range"; return SINDYN_ERR_INVALID; }
    TrackSync *t = track_sync(o);
    const OrbLevel &L0 = o->lv[0];
    const uint8_t *bgr = c->bgr[c->i_cur];
    CU_CHECK(c, cudaEventRecord(t->ev_in, c->stream));
    CU_CHECK(o, cudaStreamWaitEvent(o->stream, t->ev_in, 0));
    LAUNCH(o, k_orb_gray_in, dim3(cdiv(o->W, 32), cdiv(o->H, 8)), dim3(32, 8), 0, bgr, o->W, o->H, rgb_order,
           o->pyr + (size_t)ORB_EDGE * L0.pitch + ORB_EDGE, L0.pitch);
    SD_CHECK(orb_enqueue_unmasked(o));
    SD_CHECK(detect_run_public(c));
    // the previous frame's erasure half may still be reading o->mask
    CU_CHECK(c, cudaStreamWaitEvent(c->stream, t->ev_orb_done, 0));
    if (dilate_k > 1) SD_CHECK(morph_run(c, c->dd.out, o->mask, nullptr, c->W, c->H, dilate_k, MORPH_DILATE));
    else CU_CHECK(c, cudaMemcpyAsync(o->mask, c->dd.out, c->N, cudaMemcpyDeviceToDevice, c->stream));
    LAUNCH_CHECK(c);
    CU_CHECK(c, cudaEventRecord(t->ev_mask, c->stream));
    CU_CHECK(o, cudaStreamWaitEvent(o->stream, t->ev_mask, 0));
    SD_CHECK(orb_enqueue_masked(o, true));
    CU_CHECK(o, cudaEventRecord(t->ev_orb_done, o->stream));
    return SINDYN_OK;
}

extern "C" int sindyn_track_frame(sindyn_handle h, sindyn_orb_handle o, const uint8_t *bgr, size_t bgr_step, const uint16_t *depth, size_t depth_step,
                                  int rgb_order, int dilate_k, uint8_t *mask_out, size_t mask_step, uint8_t *label_out, size_t label_step,
                                  sindyn_keypoint *kps, uint8_t *desc, int capacity, int *n_out, int frame_idx)
{
    (void)frame_idx;
    if (!h || !o || !bgr || !depth || !n_out) return SINDYN_ERR_INVALID;
    cudaSetDevice(h->device);
    *n_out = 0;
    CU_CHECK(h, stage_in_2d(h->bgr[h->i_cur], bgr, bgr_step, (size_t)h->W * 3, h->H, h->pin_bgr, h->stream));
    CU_CHECK(h, stage_in_2d(h->depth, depth, depth_step, (size_t)h->W * 2, h->H, h->pin_depth, h->stream));
    SD_CHECK(track_enqueue(h, o, rgb_order, dilate_k));
    // outputs in pinned caller memory are written by the DMA directly, pageable ones go through the handle's bounce buffers
    const bool m_direct = mask_out && (mask_step == 0 || mask_step == (size_t)h->W) && host_ptr_is_pinned(mask_out);
    const bool l_direct = label_out && (label_step == 0 || label_step == (size_t)h->W) && host_ptr_is_pinned(label_out);
    const bool k_direct = kps && host_ptr_is_pinned(kps), d_direct = desc && host_ptr_is_pinned(desc);
    if (mask_out) CU_CHECK(h, stage_out_begin(m_direct ? (void *)mask_out : (void *)h->pin_out0, o->mask, h->N, h->stream));
    if (label_out) CU_CHECK(h, stage_out_begin(l_direct ? (void *)label_out : (void *)h->pin_out1, h->rc.label_out, h->N, h->stream));
    // key points: the full fixed-capacity buffers are small next to the frame (ORB_OUT_MAX x (24 + 32) B would be 458 KB), so
    // read the count first and then exactly n entries
    CU_CHECK(o, cudaMemcpyAsync(o->ctl_host, o->ctl, sizeof(OrbControl), cudaMemcpyDeviceToHost, o->stream));
    CU_CHECK(o, cudaStreamSynchronize(o->stream));
    if (o->ctl_host->overflow) { h->err = o->err = "orb: a fixed-capacity list overflowed (candidates / nodes / output)"; return SINDYN_ERR_CAPACITY; }
    const int n = o->ctl_host->n_out;
    *n_out = n;
    if (n > capacity) { h->err = o->err = "orb: output capacity too small"; return SINDYN_ERR_CAPACITY; }
    if (n && kps) CU_CHECK(o, cudaMemcpyAsync(k_direct ? (void *)kps : (void *)o->pin_kp, o->out_host_fmt, sizeof(sindyn_keypoint) * n, cudaMemcpyDeviceToHost, o->stream));
    if (n && desc) CU_CHECK(o, cudaMemcpyAsync(d_direct ? (void *)desc : (void *)o->pin_desc, o->desc, (size_t)32 * n, cudaMemcpyDeviceToHost, o->stream));
    SD_CHECK(detect_check_capacity(h));   // synchronises h->stream (mask / label copies included)
    CU_CHECK(o, cudaStreamSynchronize(o->stream));
    if (mask_out && !m_direct) stage_out_finish(mask_out, mask_step, h->pin_out0, h->W, h->H);
    if (label_out && !l_direct) stage_out_finish(label_out, label_step, h->pin_out1, h->W, h->H);
    if (n && kps && !k_direct) memcpy(kps, o->pin_kp, sizeof(sindyn_keypoint) * n);
    if (n && desc && !d_direct) memcpy(desc, o->pin_desc, (size_t)32 * n);
    return SINDYN_OK;
}

int pipe_note_gray_read(sindyn_ctx *c, cudaStream_t orb_stream);   // pipe.cu

// the same through the frame pipeline (pipe.cu): the detector's image-only stages of this frame may run while the previous
// frame is still being decided; the inputs are device buffers the pipeline copies on its own stream
static int track_enqueue_pipe(sindyn_ctx *c, sindyn_orb *o, const uint8_t *bgr_src, size_t bgr_step, const uint16_t *depth_src, size_t depth_step, bool host_src,
                              PipeFlags *flags, int rgb_order, int dilate_k)
{
    if (o->device != c->device || o->W != c->W || o->H != c->H) { c->err = "track_frame: the two handles differ in device or image size"; return SINDYN_ERR_INVALID; }
    if (dilate_k < 0 || dilate_k > MORPH_MAX_K) { c->err = "track_frame: dilate_k out of range"; return SINDYN_ERR_INVALID; }
    TrackSync *t = track_sync(o);
    const OrbLevel &L0 = o->lv[0];
    const uint8_t *bgr = c->bgr[c->i_cur];
    SD_CHECK(pipe_detect_run_src(c, bgr_src, bgr_step, depth_src, depth_step, host_src, flags));
    CU_CHECK(o, cudaStreamWaitEvent(o->stream, pipe_input_event(c), 0));
    LAUNCH(o, k_orb_gray_in, dim3(cdiv(o->W, 32), cdiv(o->H, 8)), dim3(32, 8), 0, bgr, o->W, o->H, rgb_order,
           o->pyr + (size_t)ORB_EDGE * L0.pitch + ORB_EDGE, L0.pitch);
    SD_CHECK(pipe_note_gray_read(c, o->stream));
    SD_CHECK(orb_enqueue_unmasked(o));
    // the 15x15 dilation runs on the extractor's stream (after the previous frame's erasure half in stream order, so o->mask is
    // free): the detector's stream goes straight on to the next frame's homography -- that chain bounds the frame rate
    CU_CHECK(o, cudaStreamWaitEvent(o->stream, pipe_done_event(c), 0));
    {
        cudaStream_t keep = c->stream;
        c->stream = o->stream;
        int st = SINDYN_OK;
        if (dilate_k > 1) st = morph_run(c, c->dd.out, o->mask, nullptr, c->W, c->H, dilate_k, MORPH_DILATE);
        else if (cudaMemcpyAsync(o->mask, c->dd.out, c->N, cudaMemcpyDeviceToDevice, o->stream) != cudaSuccess) st = SINDYN_ERR_CUDA;
        c->stream = keep;
        SD_CHECK(st);
    }
    LAUNCH_CHECK(c);
    SD_CHECK(pipe_note_mask_read(c, o->stream));
    SD_CHECK(orb_enqueue_masked(o, true));
    CU_CHECK(o, cudaEventRecord(t->ev_orb_done, o->stream));
    return SINDYN_OK;
}

extern "C" int sindyn_track_frame_resident(sindyn_handle h, sindyn_orb_handle o, int slot, int rgb_order, int dilate_k, int frame_idx)
{
    (void)frame_idx;
    if (!h || !o) return SINDYN_ERR_INVALID;
    cudaSetDevice(h->device);
    if (slot < 0 || slot >= SINDYN_MAX_SLOTS || !h->slot_bgr[slot]) { h->err = "track_frame_resident: empty slot"; return SINDYN_ERR_INVALID; }
    if (pipe_usable(h)) return track_enqueue_pipe(h, o, h->slot_bgr[slot], 0, h->slot_depth[slot], 0, false, nullptr, rgb_order, dilate_k);
    CU_CHECK(h, cudaMemcpyAsync(h->bgr[h->i_cur], h->slot_bgr[slot], (size_t)h->N * 3, cudaMemcpyDeviceToDevice, h->stream));
    CU_CHECK(h, cudaMemcpyAsync(h->depth, h->slot_depth[slot], (size_t)h->N * 2, cudaMemcpyDeviceToDevice, h->stream));
    return track_enqueue(h, o, rgb_order, dilate_k);
}

// The frames enqueued by sindyn_track_frame_resident run on several streams; after this call everything they have enqueued
// precedes whatever is enqueued next on the detector handle's stream (an event record, a synchronisation)
extern "C" int sindyn_track_join(sindyn_handle h, sindyn_orb_handle o)
{
    if (!h || !o) return SINDYN_ERR_INVALID;
    cudaSetDevice(h->device);
    SD_CHECK(pipe_join(h));
    if (o->track) CU_CHECK(h, cudaStreamWaitEvent(h->stream, ((TrackSync *)o->track)->ev_orb_done, 0));
    return SINDYN_OK;
}

// ---------------------------------------------------------------- asynchronous per-frame entry: submit frame i + 1, then collect frame i
// sindyn_track_frame returns a frame's results before it accepts the next frame, so the GPU never sees two frames at once.  A
// caller that has the next image at hand (a dataset on disk, rgbd_tum_noros.cc:113-192; a camera running ahead of the tracker)
// submits it first: the upload and the image-only stages of frames i + 1, i + 2 then overlap the decision, the extractor and the
// download of frame i (pipe.cu).  At most three frames are in flight; results are collected in submission order.
static int track_async_init(sindyn_ctx *c, sindyn_orb *o, TrackSync *t)
{
    if (t->r_mask[0]) return SINDYN_OK;
    t->r_cap = o->nfeatures * 2 + 64 < ORB_OUT_MAX ? o->nfeatures * 2 + 64 : ORB_OUT_MAX;
    for (int p = 0; p < TRACK_NB; ++p) {
        SD_CHECK(o->halloc(&t->r_mask[p], (size_t)c->N));
        SD_CHECK(o->halloc(&t->r_label[p], (size_t)c->N));
        SD_CHECK(o->halloc(&t->r_kp[p], (size_t)t->r_cap));
        SD_CHECK(o->halloc(&t->r_desc[p], (size_t)t->r_cap * 32));
        SD_CHECK(o->halloc(&t->r_ctl[p], 1));
        SD_CHECK(o->halloc(&t->r_flags[p], 1));
        CU_CHECK(o, cudaEventCreateWithFlags(&t->ev_res_main[p], cudaEventDisableTiming));
        CU_CHECK(o, cudaEventCreateWithFlags(&t->ev_res_orb[p], cudaEventDisableTiming));
    }
    return SINDYN_OK;
}

extern "C" int sindyn_track_submit(sindyn_handle h, sindyn_orb_handle o, const uint8_t *bgr, size_t bgr_step, const uint16_t *depth, size_t depth_step,
                                   int rgb_order, int dilate_k, int frame_idx)
{
    (void)frame_idx;
    if (!h || !o || !bgr || !depth) return SINDYN_ERR_INVALID;
    cudaSetDevice(h->device);
    if (!pipe_usable(h)) { h->err = "track_submit: needs the CUDA-graph path (use_graphs = 1, stage_timing = 0)"; return SINDYN_ERR_STATE; }
    TrackSync *t = track_sync(o);
    SD_CHECK(track_async_init(h, o, t));
    if (t->n_submit - t->n_collect >= TRACK_NB) { h->err = "track_submit: three frames are in flight already, collect one first"; return SINDYN_ERR_STATE; }
    const int p = (int)(t->n_submit % TRACK_NB);
    SD_CHECK(track_enqueue_pipe(h, o, bgr, bgr_step, depth, depth_step, true, t->r_flags[p], rgb_order, dilate_k));
    // results into this slot's pinned buffers: the labels on the detector's stream (from the rolled state: a later frame's
    // re-clustering may already overwrite rc.label_out), the dilated mask and the key points on the extractor's
    CU_CHECK(h, cudaMemcpyAsync(t->r_label[p], h->label_last, h->N, cudaMemcpyDeviceToHost, pipe_chain_stream(h)));   // (before the next frame's state roll)
    CU_CHECK(h, cudaEventRecord(t->ev_res_main[p], pipe_chain_stream(h)));
    CU_CHECK(o, cudaMemcpyAsync(t->r_mask[p], o->mask, h->N, cudaMemcpyDeviceToHost, o->stream));
    CU_CHECK(o, cudaMemcpyAsync(t->r_ctl[p], o->ctl, sizeof(OrbControl), cudaMemcpyDeviceToHost, o->stream));
    CU_CHECK(o, cudaMemcpyAsync(t->r_kp[p], o->out_host_fmt, sizeof(sindyn_keypoint) * t->r_cap, cudaMemcpyDeviceToHost, o->stream));
    CU_CHECK(o, cudaMemcpyAsync(t->r_desc[p], o->desc, (size_t)32 * t->r_cap, cudaMemcpyDeviceToHost, o->stream));
    CU_CHECK(o, cudaEventRecord(t->ev_res_orb[p], o->stream));
    ++t->n_submit;
    return SINDYN_OK;
}

extern "C" int sindyn_track_collect(sindyn_handle h, sindyn_orb_handle o, uint8_t *mask_out, size_t mask_step, uint8_t *label_out, size_t label_step,
                                    sindyn_keypoint *kps, uint8_t *desc, int capacity, int *n_out)
{
    if (!h || !o || !n_out) return SINDYN_ERR_INVALID;
    cudaSetDevice(h->device);
    *n_out = 0;
    TrackSync *t = (TrackSync *)o->track;
    if (!t || t->n_collect >= t->n_submit) { h->err = "track_collect: no frame in flight"; return SINDYN_ERR_STATE; }
    const int p = (int)(t->n_collect % TRACK_NB);
    ++t->n_collect;
    CU_CHECK(h, cudaEventSynchronize(t->ev_res_main[p]));
    CU_CHECK(o, cudaEventSynchronize(t->ev_res_orb[p]));
    const PipeFlags *f = t->r_flags[p];
    if (f->peac_hdr[2]) { h->err = "plane fitter: a fixed-capacity list overflowed (> 64 planes, > 512 neighbours of one node, or a region-growing level > 131072 entries)"; return SINDYN_ERR_CAPACITY; }
    if (f->rc.overflow) { h->err = "recluster: more than RC_MAXC components"; return SINDYN_ERR_CAPACITY; }
    if (f->rc.pf_overflow) { h->err = "plane-edge filter: more than RC_PF_MAXC contours"; return SINDYN_ERR_CAPACITY; }
    if (f->edge_scalars[3]) { h->err = "depth_edges: more than EDGE_EP_CAP candidate end points"; return SINDYN_ERR_CAPACITY; }
    if (t->r_ctl[p]->overflow) { h->err = o->err = "orb: a fixed-capacity list overflowed (candidates / nodes / output)"; return SINDYN_ERR_CAPACITY; }
    const int n = t->r_ctl[p]->n_out;
    *n_out = n;
    if (n > capacity || n > t->r_cap) { h->err = o->err = "orb: output capacity too small"; return SINDYN_ERR_CAPACITY; }
    if (mask_out) stage_out_finish(mask_out, mask_step, t->r_mask[p], h->W, h->H);
    if (label_out) stage_out_finish(label_out, label_step, t->r_label[p], h->W, h->H);
    if (n && kps) memcpy(kps, t->r_kp[p], sizeof(sindyn_keypoint) * n);
    if (n && desc) memcpy(desc, t->r_desc[p], (size_t)32 * n);
    return SINDYN_OK;
}

extern "C" int sindyn_track_get_results(sindyn_handle h, sindyn_orb_handle o, uint8_t *mask_dilated, uint8_t *labels, sindyn_keypoint *kps, uint8_t *desc,
                                        int capacity, int *n_out)
{
    if (!h || !o || !n_out) return SINDYN_ERR_INVALID;
    cudaSetDevice(h->device);
    CU_CHECK(o, cudaStreamSynchronize(o->stream));
    if (mask_dilated) CU_CHECK(h, cudaMemcpyAsync(mask_dilated, o->mask, h->N, cudaMemcpyDeviceToHost, h->stream));
    if (labels) CU_CHECK(h, cudaMemcpyAsync(labels, h->rc.label_out, h->N, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaMemcpyAsync(o->ctl_host, o->ctl, sizeof(OrbControl), cudaMemcpyDeviceToHost, h->stream));
    SD_CHECK(detect_check_capacity(h));
    if (o->ctl_host->overflow) { h->err = o->err = "orb: a fixed-capacity list overflowed (candidates / nodes / output)"; return SINDYN_ERR_CAPACITY; }
    const int n = o->ctl_host->n_out;
    *n_out = n;
    if (n > capacity) { h->err = "orb: output capacity too small"; return SINDYN_ERR_CAPACITY; }
    if (n && kps) CU_CHECK(h, cudaMemcpy(kps, o->out_host_fmt, sizeof(sindyn_keypoint) * n, cudaMemcpyDeviceToHost));
    if (n && desc) CU_CHECK(h, cudaMemcpy(desc, o->desc, (size_t)32 * n, cudaMemcpyDeviceToHost));
    return SINDYN_OK;
}

extern "C" int sindyn_orb_get_pyramid_level(sindyn_orb_handle h, int level, uint8_t *out, int *w_out, int *h_out)
{
    if (!h || level < 0 || level >= h->nlevels) return SINDYN_ERR_INVALID;
    cudaSetDevice(h->device);
    const OrbLevel &L = h->lv[level];
    if (w_out) *w_out = L.w;
    if (h_out) *h_out = L.h;
    if (out) {
        CU_CHECK(h, cudaMemcpy2DAsync(out, L.w, h->pyr + L.pad_off + (size_t)ORB_EDGE * L.pitch + ORB_EDGE, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost,
                                      h->stream));
        CU_CHECK(h, cudaStreamSynchronize(h->stream));
    }
    return SINDYN_OK;
}

// test hook: FAST candidates of one level in distribution order (x, y relative to minBorder, response)
extern "C" int sindyn_orb_get_candidates(sindyn_orb_handle h, int level, int *xyr, int capacity, int *n_out)
{
    if (!h || level < 0 || level >= h->nlevels || !n_out) return SINDYN_ERR_INVALID;
    cudaSetDevice(h->device);
    OrbControl c;
    CU_CHECK(h, cudaMemcpy(&c, h->ctl, sizeof c, cudaMemcpyDeviceToHost));
    const int n = c.cand_count[level];
    *n_out = n;
    if (n > capacity) return SINDYN_ERR_CAPACITY;
    std::vector<OrbCand> tmp(n);
    if (n) CU_CHECK(h, cudaMemcpy(tmp.data(), h->cand + (size_t)level * ORB_KMAX, sizeof(OrbCand) * n, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) { xyr[3 * i] = tmp[i].x; xyr[3 * i + 1] = tmp[i].y; xyr[3 * i + 2] = tmp[i].resp; }
    return SINDYN_OK;
}

// test hook: raw planes of one level: which = 0 padded image, 1 FAST score map, 2 blurred image
extern "C" int sindyn_orb_get_plane(sindyn_orb_handle h, int level, int which, uint8_t *out, int *w_out, int *h_out)
{
    if (!h || level < 0 || level >= h->nlevels || !out) return SINDYN_ERR_INVALID;
    cudaSetDevice(h->device);
    const OrbLevel &L = h->lv[level];
    const int w = which == 0 ? L.pitch : L.w, hh = which == 0 ? L.h + 2 * ORB_EDGE : L.h;
    const uint8_t *src = which == 0 ? h->pyr + L.pad_off : (which == 1 ? h->score + L.img_off : h->blur + L.img_off);
    if (w_out) *w_out = w;
    if (h_out) *h_out = hh;
    CU_CHECK(h, cudaMemcpy(out, src, (size_t)w * hh, cudaMemcpyDeviceToHost));
    return SINDYN_OK;
}

// ------------------------------------------------------------------ Frame construction right after ORB (row f2)
// ORB_SLAM2::Frame (Frame.cc:143-170): UndistortKeyPoints (:714-753), ComputeImageBounds (:507-535),
// ComputeStereoFromRGBD (depth lookup + virtual right coordinate), AssignFeaturesToGrid / PosInGrid (:453-463).
// The keypoints stay on the device between sindyn_orb_extract and this call.
#define FR_COLS 64
#define FR_ROWS 48

// cv::undistortPoints(pts, K, dist, noArray(), K): 5 fixed-point iterations in double (verified bit-exact against cv2)
__device__ __forceinline__ void fr_undistort(float u, float v, const sindyn_frame_params &P, float &uo, float &vo)
{
    const double fx = P.fx, fy = P.fy, cx = P.cx, cy = P.cy;
    const double k1 = P.k1, k2 = P.k2, p1 = P.p1, p2 = P.p2, k3 = P.k3;
    const double x0 = ((double)u - cx) / fx, y0 = ((double)v - cy) / fy;
    double x = x0, y = y0;
    for (int it = 0; it < 5; ++it) {
        const double r2 = x * x + y * y;
        double icdist = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2);
        if (icdist < 0) icdist = 1.0;
        const double dx = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x), dy = p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y;
        x = (x0 - dx) * icdist;
        y = (y0 - dy) * icdist;
    }
    uo = (float)(x * fx + cx);
    vo = (float)(y * fy + cy);
}

__global__ void __launch_bounds__(1024) k_orb_frame(const sindyn_keypoint *__restrict__ kps, const OrbControl *__restrict__ ctl,
                                                     const uint16_t *__restrict__ depth, int W, int H, sindyn_frame_params P,
                                                     float *__restrict__ un, float *__restrict__ dep, float *__restrict__ uright,
                                                     float *__restrict__ bounds, int *__restrict__ offsets, int *__restrict__ indices)
{
    __shared__ float s_b[4];
    __shared__ int s_cnt[FR_COLS * FR_ROWS];
    __shared__ short s_cell[ORB_OUT_MAX];
    const int n = ctl->n_out;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int c = tid; c < FR_COLS * FR_ROWS; c += nt) s_cnt[c] = 0;
    if (tid == 0) {
        if (P.k1 != 0.0f) {   // ComputeImageBounds: undistorted image corners
            float x[4], y[4];
            const float cxs[4] = {0.f, (float)W, 0.f, (float)W}, cys[4] = {0.f, 0.f, (float)H, (float)H};
            for (int k = 0; k < 4; ++k) fr_undistort(cxs[k], cys[k], P, x[k], y[k]);
            s_b[0] = fminf(x[0], x[2]); s_b[1] = fmaxf(x[1], x[3]); s_b[2] = fminf(y[0], y[1]); s_b[3] = fmaxf(y[2], y[3]);
        } else {
            s_b[0] = 0.f; s_b[1] = (float)W; s_b[2] = 0.f; s_b[3] = (float)H;
        }
        for (int k = 0; k < 4; ++k) bounds[k] = s_b[k];
    }
    __syncthreads();
    const float inv_w = (float)FR_COLS / (s_b[1] - s_b[0]), inv_h = (float)FR_ROWS / (s_b[3] - s_b[2]);
    for (int i = tid; i < n; i += nt) {
        const sindyn_keypoint k = kps[i];
        float ux = k.x, uy = k.y;
        if (P.k1 != 0.0f) fr_undistort(k.x, k.y, P, ux, uy);
        un[2 * i] = ux; un[2 * i + 1] = uy;
        // imDepth.at<float>(v, u): float -> int truncation of the DISTORTED keypoint; depth = raw * (1 / DepthMapFactor)
        const float d = (float)depth[(int)k.y * W + (int)k.x] * P.depth_map_factor;
        dep[i] = d > 0.f ? d : -1.0f;
        uright[i] = d > 0.f ? ux - P.bf / d : -1.0f;
        // PosInGrid: C round() of float products
        const int px = (int)round((double)((ux - s_b[0]) * inv_w)), py = (int)round((double)((uy - s_b[2]) * inv_h));
        int cell = -1;
        if (px >= 0 && px < FR_COLS && py >= 0 && py < FR_ROWS) { cell = px * FR_ROWS + py; atomicAdd(&s_cnt[cell], 1); }
        s_cell[i] = (short)cell;
    }
    __syncthreads();
    if (tid == 0) {   // exclusive scan over the 3072 cells (mGrid[i][j] order: column-major in i)
        int acc = 0;
        for (int c = 0; c < FR_COLS * FR_ROWS; ++c) { const int v = s_cnt[c]; offsets[c] = acc; s_cnt[c] = acc; acc += v; }
        offsets[FR_COLS * FR_ROWS] = acc;
    }
    __syncthreads();
    // stable fill: keypoint i goes after every j < i of the same cell (push_back order of AssignFeaturesToGrid)
    for (int i = tid; i < n; i += nt) {
        const int cell = s_cell[i];
        if (cell < 0) continue;
        int before = 0;
        for (int j = 0; j < i; ++j) before += s_cell[j] == cell;
        indices[s_cnt[cell] + before] = i;
    }
}

extern "C" int sindyn_orb_frame_features(sindyn_orb_handle h, const uint16_t *depth_raw, size_t depth_step, const sindyn_frame_params *params,
                                         float *keys_un, float *depth_out, float *u_right_out, float *bounds_out, int *grid_offsets,
                                         int *grid_indices, int capacity, int *n_out)
{
    if (!h || !depth_raw || !params || !n_out) return SINDYN_ERR_INVALID;
    cudaSetDevice(h->device);
    const int n = h->ctl_host->n_out;   // result of the last sindyn_orb_extract
    *n_out = n;
    if (n > capacity) { h->err = "orb frame: output capacity too small"; return SINDYN_ERR_CAPACITY; }
    CU_CHECK(h, stage_in_2d(h->depth, depth_raw, depth_step, (size_t)h->W * 2, h->H, h->pin_depth, h->stream));
    LAUNCH(h, k_orb_frame, 1, 1024, 0, h->out_host_fmt, h->ctl, h->depth, h->W, h->H, *params, h->fr_un, h->fr_depth, h->fr_uright, h->fr_bounds,
           h->fr_offsets, h->fr_indices);
    LAUNCH_CHECK(h);
    if (keys_un && n) CU_CHECK(h, cudaMemcpyAsync(keys_un, h->fr_un, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, h->stream));
    if (depth_out && n) CU_CHECK(h, cudaMemcpyAsync(depth_out, h->fr_depth, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
    if (u_right_out && n) CU_CHECK(h, cudaMemcpyAsync(u_right_out, h->fr_uright, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
    if (bounds_out) CU_CHECK(h, cudaMemcpyAsync(bounds_out, h->fr_bounds, sizeof(float) * 4, cudaMemcpyDeviceToHost, h->stream));
    if (grid_offsets) CU_CHECK(h, cudaMemcpyAsync(grid_offsets, h->fr_offsets, sizeof(int) * (FR_COLS * FR_ROWS + 1), cudaMemcpyDeviceToHost, h->stream));
    if (grid_indices && n) CU_CHECK(h, cudaMemcpyAsync(grid_indices, h->fr_indices, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    return SINDYN_OK;
}

// ------------------------------------------------------------------ descriptor matching
// ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono) (src/ORBmatcher.cc:1328-1470) of
// Tracking::TrackWithMotionModel, with Frame::GetFeaturesInArea (src/Frame.cc:398-452), ORBmatcher::DescriptorDistance
// (:1647-1665) and ComputeThreeMaxima (:1601-1643).  The current frame is the one resident in this handle (last
// sindyn_orb_extract + sindyn_orb_frame_features); the last frame's map points arrive as plain arrays.
//   k_match_candidates  one warp per last-frame point: projection, grid window, level / window / right-coordinate tests and the
//                       256-bit Hamming distance of every surviving candidate, in the reference's visiting order
//   k_match_assign      one CTA packs the candidate lists into shared memory, then one warp walks the last-frame points IN ORDER (the reference's result depends on it: a key point whose
//                       map point has observations blocks later matches), lanes take the minimum over the candidate list;
//                       rotation histogram + three-maxima filter
#define MT_MAXC 512                 // candidates kept per last-frame point (th = 30, the retry of TrackWithMotionModel, needs ~300)
#define MT_TH_HIGH 100
#define MT_HISTO 30

struct MatchParams {
    float fx, fy, cx, cy, bf;
    float R[9], t[3];               // CurrentFrame.mTcw
    float th;
    int mode;                       // 0: octave window [o-1, o+1]; 1: forward (>= o); 2: backward (<= o)
    int check_ori, nlevels;
    float scale[ORB_MAX_LEVELS];    // mvScaleFactors
};

struct MatchStage {
    float *xyz = nullptr, *angle = nullptr;
    uint8_t *valid = nullptr, *observed = nullptr, *desc = nullptr, *blocked = nullptr;
    int *octave = nullptr, *match = nullptr, *ctl = nullptr;   // ctl: nmatches, overflow
    unsigned short *cand_i2 = nullptr, *cand_d = nullptr;
    int *cand_n = nullptr;
    unsigned *best0 = nullptr;
};

__device__ __forceinline__ float match_gemv_row(const float *R, const float *x, float t)
{
    // cv::Mat float product + addend: double accumulation, one rounding (GEMMSingleMul<float, double>)
    return (float)(((double)R[0] * (double)x[0] + (double)R[1] * (double)x[1] + (double)R[2] * (double)x[2]) * 1.0 + (double)t * 1.0);
}

// One WARP per last-frame point.  The CSR grid stores cell (ix, iy) at ix * 48 + iy, so the cells iy = r0 .. r1 of one grid column
// are ONE contiguous index run in the reference's visiting order: lanes take consecutive entries of the run, a ballot keeps the
// order when the survivors are appended to the candidate list.  best0 = first minimum over the candidates that are not blocked
// on entry (k_match_assign re-scans a list only when that candidate was taken by an earlier point in the meantime).
__global__ void __launch_bounds__(128) k_match_candidates(int n_last, const float *__restrict__ xyz, const uint8_t *__restrict__ valid,
                                                          const uint8_t *__restrict__ desc_last, const int *__restrict__ octave_last,
                                                          const sindyn_keypoint *__restrict__ kps, const float *__restrict__ un,
                                                          const float *__restrict__ uright, const uint8_t *__restrict__ desc,
                                                          const float *__restrict__ bounds, const int *__restrict__ offsets,
                                                          const int *__restrict__ indices, const uint8_t *__restrict__ blocked_in, MatchParams P,
                                                          unsigned short *__restrict__ cand_i2, unsigned short *__restrict__ cand_d,
                                                          int *__restrict__ cand_n, unsigned *__restrict__ best0, int *__restrict__ ctl)
{
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n_last) return;
    if (lane == 0) { cand_n[i] = 0; best0[i] = 0xffffffffu; }
    if (!valid[i]) return;
    const float X[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    const float xc = match_gemv_row(P.R, X, P.t[0]), yc = match_gemv_row(P.R + 3, X, P.t[1]), zc = match_gemv_row(P.R + 6, X, P.t[2]);
    const float invzc = (float)(1.0 / (double)zc);
    if (invzc < 0) return;
    const float u = P.fx * xc * invzc + P.cx, v = P.fy * yc * invzc + P.cy;
    const float minx = bounds[0], maxx = bounds[1], miny = bounds[2], maxy = bounds[3];
    if (isnan(u) || isnan(v)) return;
    if (u < minx || u > maxx || v < miny || v > maxy) return;
    const int o = octave_last[i];
    const float r = P.th * P.scale[o];
    const int min_level = P.mode == 2 ? 0 : (P.mode == 1 ? o : o - 1), max_level = P.mode == 1 ? -1 : (P.mode == 2 ? o : o + 1);
    // Frame::GetFeaturesInArea
    const float inv_w = (float)FR_COLS / (maxx - minx), inv_h = (float)FR_ROWS / (maxy - miny);
    const int c0 = max(0, (int)floor((double)((u - minx - r) * inv_w)));
    if (c0 >= FR_COLS) return;
    const int c1 = min(FR_COLS - 1, (int)ceil((double)((u - minx + r) * inv_w)));
    if (c1 < 0) return;
    const int r0 = max(0, (int)floor((double)((v - miny - r) * inv_h)));
    if (r0 >= FR_ROWS) return;
    const int r1 = min(FR_ROWS - 1, (int)ceil((double)((v - miny + r) * inv_h)));
    if (r1 < 0) return;
    const bool check = min_level > 0 || max_level >= 0;
    unsigned dl[8];
    for (int k = 0; k < 8; ++k) dl[k] = ((const unsigned *)desc_last)[8 * i + k];
    const float ur = u - P.bf * invzc;
    int n = 0;
    unsigned best = 0xffffffffu;
    for (int ix = c0; ix <= c1; ++ix) {
        const int j0 = offsets[ix * FR_ROWS + r0], j1 = offsets[ix * FR_ROWS + r1 + 1];
        for (int jb = j0; jb < j1; jb += 32) {
            const int j = jb + lane;
            bool pass = false;
            int i2 = 0, d = 0;
            if (j < j1) {
                i2 = indices[j];
                pass = true;
                if (check) {
                    const int oc = kps[i2].octave;
                    if (oc < min_level || (max_level >= 0 && oc > max_level)) pass = false;
                }
                if (pass) {
                    const float dx = un[2 * i2] - u, dy = un[2 * i2 + 1] - v;
                    pass = fabsf(dx) < r && fabsf(dy) < r;
                }
                // (the "already matched to an observed map point" test depends on earlier points: k_match_assign)
                if (pass) {
                    const float urt = uright[i2];
                    if (urt > 0 && fabsf(ur - urt) > r) pass = false;
                }
                if (pass)
                    for (int k = 0; k < 8; ++k) d += __popc(dl[k] ^ ((const unsigned *)desc)[8 * i2 + k]);
            }
            const unsigned m = __ballot_sync(0xffffffffu, pass);
            if (pass) {
                const int pos = n + __popc(m & ((1u << lane) - 1u));
                if (pos < MT_MAXC) {
                    cand_i2[(size_t)i * MT_MAXC + pos] = (unsigned short)i2;
                    cand_d[(size_t)i * MT_MAXC + pos] = (unsigned short)d;
                    if (!(blocked_in && blocked_in[i2])) best = min(best, ((unsigned)d << 16) | (unsigned)pos);
                }
            }
            n += __popc(m);
        }
    }
    for (int off = 16; off > 0; off >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, off));
    if (lane == 0) {
        if (n > MT_MAXC) { ctl[1] = 1; n = MT_MAXC; }
        cand_n[i] = n;
        best0[i] = best;
    }
}

// The ordered pass is a chain of n_last dependent steps, so everything a step touches sits in shared memory: the first-minimum
// candidate of every point (best0 and its key point), the blocked flags, the result.  A point whose first-minimum candidate is
// still free takes it (the minimum over a subset that contains the old minimum is the old minimum); only when an earlier point
// with observations took it in the meantime is the list re-scanned from global memory.  All 32 lanes of warp 0 execute the chain
// redundantly (same-value stores), so no step needs a warp barrier; the rotation histogram, the three-maxima filter and the
// write-back are order independent and run on the whole CTA afterwards.
#define MT_NT 1024

__global__ void __launch_bounds__(MT_NT) k_match_assign(int n_last, int n_cur, const uint8_t *__restrict__ observed, const float *__restrict__ angle_last,
                                                        const sindyn_keypoint *__restrict__ kps, const unsigned short *__restrict__ cand_i2,
                                                        const unsigned short *__restrict__ cand_d, const int *__restrict__ cand_n,
                                                        const unsigned *__restrict__ best0, const uint8_t *__restrict__ blocked_in, int check_ori,
                                                        int *__restrict__ match, int *__restrict__ ctl)
{
    extern __shared__ unsigned mt_sm[];
    unsigned *s_best = mt_sm;                                            // [n_last] distance << 16 | position
    float *s_ang = (float *)(mt_sm + ORB_OUT_MAX);                       // [n_cur]  mvKeysUn[i2].angle
    float *s_angl = s_ang + ORB_OUT_MAX;                                 // [n_last] LastFrame.mvKeysUn[i].angle
    unsigned short *s_bi2 = (unsigned short *)(s_angl + ORB_OUT_MAX);    // [n_last] key point of the first-minimum candidate | observed << 15
    unsigned short *s_n = s_bi2 + ORB_OUT_MAX;                           // [n_last] list length
    short *s_match = (short *)(s_n + ORB_OUT_MAX);                       // [n_cur]  result
    unsigned short *s_rec = (unsigned short *)(s_match + ORB_OUT_MAX);   // [n_last] accepted matches in order: key point
    unsigned short *s_reci = s_rec + ORB_OUT_MAX;                        // [n_last]                            last-frame point
    uint8_t *s_rbin = (uint8_t *)(s_reci + ORB_OUT_MAX);                 // [n_last] rotation bin of every accepted match
    __shared__ uint8_t s_blk[ORB_OUT_MAX];
    __shared__ int s_hist[MT_HISTO];
    __shared__ int s_nrec, s_removed;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int k = tid; k < n_cur; k += MT_NT) { s_blk[k] = blocked_in ? blocked_in[k] : 0; s_match[k] = -1; s_ang[k] = kps[k].angle; }
    for (int i = tid; i < n_last; i += MT_NT) {
        const unsigned b = best0[i];
        s_best[i] = b;
        s_n[i] = (unsigned short)cand_n[i];
        s_angl[i] = angle_last[i];
        s_bi2[i] = (unsigned short)((b != 0xffffffffu ? cand_i2[(size_t)i * MT_MAXC + (b & 0xffff)] : 0) | (observed[i] ? 0x8000u : 0u));
    }
    if (tid < MT_HISTO) s_hist[tid] = 0;
    if (tid == 0) s_removed = 0;
    __syncthreads();
    if (tid < 32) {
        int nrec = 0;
        // (next point's first-minimum candidate is fetched one step ahead: the chain of a step is one s_blk read)
        unsigned nbest = n_last > 0 ? s_best[0] : 0xffffffffu;
        unsigned nbi2 = n_last > 0 ? s_bi2[0] : 0;
        for (int i = 0; i < n_last; ++i) {
            unsigned best = nbest;
            const unsigned pk = nbi2;                     // key point | observed << 15
            if (i + 1 < n_last) { nbest = s_best[i + 1]; nbi2 = s_bi2[i + 1]; }
            if (best == 0xffffffffu) continue;            // no candidate that was free on entry (the blocked set only grows)
            int i2 = pk & 0x7fff;
            if (s_blk[i2]) {                              // taken by an earlier point with observations: re-scan the list
                const int n = s_n[i];
                best = 0xffffffffu;
                for (int k = lane; k < n; k += 32) {
                    const int c2 = cand_i2[(size_t)i * MT_MAXC + k];
                    if (s_blk[c2]) continue;
                    best = min(best, ((unsigned)cand_d[(size_t)i * MT_MAXC + k] << 16) | (unsigned)k);
                }
                for (int off = 16; off > 0; off >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, off));
                if (best == 0xffffffffu) continue;
                i2 = cand_i2[(size_t)i * MT_MAXC + (best & 0xffff)];
            }
            if ((best >> 16) > MT_TH_HIGH) continue;
            s_match[i2] = (short)i;                        // (a later point may overwrite it: the reference counts both)
            if (pk >> 15) s_blk[i2] = 1;
            s_rec[nrec] = (unsigned short)i2;
            s_reci[nrec] = (unsigned short)i;
            ++nrec;
        }
        if (lane == 0) s_nrec = nrec;
    }
    __syncthreads();
    const int nrec = s_nrec;
    if (check_ori) {
        const float factor = 1.0f / MT_HISTO;
        for (int k = tid; k < nrec; k += MT_NT) {
            float rot = s_angl[s_reci[k]] - s_ang[s_rec[k]];
            if (rot < 0.0f) rot += 360.0f;
            int bin = (int)roundf(rot * factor);
            if (bin == MT_HISTO) bin = 0;
            s_rbin[k] = (uint8_t)bin;
            atomicAdd(&s_hist[bin], 1);
        }
        __syncthreads();
        int ind1 = -1, ind2 = -1, ind3 = -1;
        {   // ComputeThreeMaxima (every thread computes the same thing)
            int max1 = 0, max2 = 0, max3 = 0;
            for (int k = 0; k < MT_HISTO; ++k) {
                const int sz = s_hist[k];
                if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = k; }
                else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = k; }
                else if (sz > max3) { max3 = sz; ind3 = k; }
            }
            if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
            else if ((float)max3 < 0.1f * (float)max1) ind3 = -1;
        }
        for (int k = tid; k < nrec; k += MT_NT) {
            const int b = s_rbin[k];
            if (b != ind1 && b != ind2 && b != ind3) { s_match[s_rec[k]] = -1; atomicAdd(&s_removed, 1); }
        }
        __syncthreads();
    }
    for (int k = tid; k < n_cur; k += MT_NT) match[k] = s_match[k];
    if (tid == 0) ctl[0] = nrec - s_removed;
}
#define MT_ASSIGN_SMEM ((sizeof(unsigned) + 2 * sizeof(float) + 5 * sizeof(unsigned short) + 1) * ORB_OUT_MAX)
static_assert(ORB_OUT_MAX <= 0x8000, "key point index and the observed flag share 16 bits");

static void match_stage_free(sindyn_orb *o)
{
    delete o->match;   // the device buffers belong to the handle's allocation list
    o->match = nullptr;
}

static int match_stage(sindyn_orb *o, MatchStage **out)
{
    if (!o->match) {
        MatchStage *m = new MatchStage();
        SD_CHECK(o->dalloc(&m->xyz, 3 * ORB_OUT_MAX)); SD_CHECK(o->dalloc(&m->angle, ORB_OUT_MAX));
        SD_CHECK(o->dalloc(&m->valid, ORB_OUT_MAX)); SD_CHECK(o->dalloc(&m->observed, ORB_OUT_MAX)); SD_CHECK(o->dalloc(&m->blocked, ORB_OUT_MAX));
        SD_CHECK(o->dalloc(&m->desc, 32 * ORB_OUT_MAX));
        SD_CHECK(o->dalloc(&m->octave, ORB_OUT_MAX)); SD_CHECK(o->dalloc(&m->match, ORB_OUT_MAX)); SD_CHECK(o->dalloc(&m->ctl, 2));
        SD_CHECK(o->dalloc(&m->cand_i2, (size_t)ORB_OUT_MAX * MT_MAXC)); SD_CHECK(o->dalloc(&m->cand_d, (size_t)ORB_OUT_MAX * MT_MAXC));
        SD_CHECK(o->dalloc(&m->cand_n, ORB_OUT_MAX)); SD_CHECK(o->dalloc(&m->best0, ORB_OUT_MAX));
        CU_CHECK(o, cudaFuncSetAttribute(k_match_assign, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MT_ASSIGN_SMEM));
        o->match = m;
    }
    *out = o->match;
    return SINDYN_OK;
}

extern "C" int sindyn_orb_search_by_projection(sindyn_orb_handle h, const sindyn_match_params *mp, int n_last, const float *last_xyz_world,
                                               const uint8_t *last_valid, const uint8_t *last_desc, const int *last_octave,
                                               const float *last_angle, const uint8_t *last_observed, const uint8_t *cur_blocked,
                                               int *match_out, int capacity, int *n_cur_out, int *nmatches_out)
{
    if (!h || !mp || !last_xyz_world || !last_valid || !last_desc || !last_octave || !last_angle || !last_observed || !match_out || !nmatches_out)
        return SINDYN_ERR_INVALID;
    cudaSetDevice(h->device);
    const int n_cur = h->ctl_host->n_out;   // current frame = last sindyn_orb_extract + sindyn_orb_frame_features on this handle
    if (n_cur_out) *n_cur_out = n_cur;
    if (n_last < 0 || n_last > ORB_OUT_MAX) { h->err = "search_by_projection: n_last out of range"; return SINDYN_ERR_INVALID; }
    if (n_cur > capacity) { h->err = "search_by_projection: match_out capacity too small"; return SINDYN_ERR_CAPACITY; }
    for (int i = 0; i < n_last; ++i)
        if (last_valid[i] && (last_octave[i] < 0 || last_octave[i] >= h->nlevels)) { h->err = "search_by_projection: octave out of range"; return SINDYN_ERR_INVALID; }
    MatchStage *m;
    SD_CHECK(match_stage(h, &m));
    MatchParams P;
    P.fx = mp->fx; P.fy = mp->fy; P.cx = mp->cx; P.cy = mp->cy; P.bf = mp->bf; P.th = mp->th;
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) P.R[3 * r + c] = mp->Tcw_cur[4 * r + c]; P.t[r] = mp->Tcw_cur[4 * r + 3]; }
    // twc = -Rcw^T tcw ; tlc = Rlw twc + tlw (ORBmatcher.cc:1338-1349): forward / backward motion along the optical axis
    float twc[3], tlc_z;
    {
        float nRt[9];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) nRt[3 * r + c] = -P.R[3 * c + r];
        for (int r = 0; r < 3; ++r) twc[r] = (float)(((double)nRt[3 * r] * P.t[0] + (double)nRt[3 * r + 1] * P.t[1] + (double)nRt[3 * r + 2] * P.t[2]) * 1.0);
        const float *L = mp->Tcw_last;
        tlc_z = (float)(((double)L[8] * twc[0] + (double)L[9] * twc[1] + (double)L[10] * twc[2]) * 1.0 + (double)L[11] * 1.0);
    }
    const bool forward = tlc_z > mp->b && !mp->mono, backward = -tlc_z > mp->b && !mp->mono;
    P.mode = forward ? 1 : (backward ? 2 : 0);
    P.check_ori = mp->check_orientation; P.nlevels = h->nlevels;
    for (int l = 0; l < ORB_MAX_LEVELS; ++l) P.scale[l] = l < h->nlevels ? h->lv[l].scale : 1.0f;
    if (n_last > 0) {
        CU_CHECK(h, cudaMemcpyAsync(m->xyz, last_xyz_world, sizeof(float) * 3 * n_last, cudaMemcpyHostToDevice, h->stream));
        CU_CHECK(h, cudaMemcpyAsync(m->valid, last_valid, n_last, cudaMemcpyHostToDevice, h->stream));
        CU_CHECK(h, cudaMemcpyAsync(m->desc, last_desc, (size_t)32 * n_last, cudaMemcpyHostToDevice, h->stream));
        CU_CHECK(h, cudaMemcpyAsync(m->octave, last_octave, sizeof(int) * n_last, cudaMemcpyHostToDevice, h->stream));
        CU_CHECK(h, cudaMemcpyAsync(m->angle, last_angle, sizeof(float) * n_last, cudaMemcpyHostToDevice, h->stream));
        CU_CHECK(h, cudaMemcpyAsync(m->observed, last_observed, n_last, cudaMemcpyHostToDevice, h->stream));
    }
    if (cur_blocked && n_cur > 0) CU_CHECK(h, cudaMemcpyAsync(m->blocked, cur_blocked, n_cur, cudaMemcpyHostToDevice, h->stream));
    CU_CHECK(h, cudaMemsetAsync(m->ctl, 0, sizeof(int) * 2, h->stream));
    const uint8_t *blk = cur_blocked ? m->blocked : (const uint8_t *)nullptr;
    if (n_last > 0)
        LAUNCH(h, k_match_candidates, (n_last + 3) / 4, 128, 0, n_last, m->xyz, m->valid, m->desc, m->octave, h->out_host_fmt, h->fr_un, h->fr_uright,
               h->desc, h->fr_bounds, h->fr_offsets, h->fr_indices, blk, P, m->cand_i2, m->cand_d, m->cand_n, m->best0, m->ctl);
    LAUNCH(h, k_match_assign, 1, MT_NT, MT_ASSIGN_SMEM, n_last, n_cur, m->observed, m->angle, h->out_host_fmt, m->cand_i2, m->cand_d, m->cand_n, m->best0,
           blk, P.check_ori, m->match, m->ctl);
    LAUNCH_CHECK(h);
    int ctl_host[2];
    CU_CHECK(h, cudaMemcpyAsync(ctl_host, m->ctl, sizeof(ctl_host), cudaMemcpyDeviceToHost, h->stream));
    if (n_cur > 0) CU_CHECK(h, cudaMemcpyAsync(match_out, m->match, sizeof(int) * n_cur, cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    if (ctl_host[1]) { h->err = "search_by_projection: more than 512 candidates in one search window"; return SINDYN_ERR_CAPACITY; }
    *nmatches_out = ctl_host[0];
    return SINDYN_OK;
}
