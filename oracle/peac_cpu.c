/*
 * oracle/peac_cpu.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or call this
 * file.  The product path is libsindyn_cuda.so and must never link it.
 *
 * C restatement of the PEAC plane fitter as DynaDetect::CalOccluded drives it (reference: ORB_SLAM2/src/DynaDetect.cc:558-593;
 * ORB_SLAM2/include/PEAC/AHCPlaneFitter.hpp, AHCPlaneSeg.hpp, AHCParamSet.hpp, DisjointSet.hpp, eig33sym.hpp,
 * plane_fitter_pcl.hpp:160-317): PlaneFitter::run with doRefine --
 *   initGraph            AHCPlaneFitter.hpp:881-1039   16x16 blocks, INIT_STRICT, Stats::push in raster order (double sums)
 *   ahCluster            AHCPlaneFitter.hpp:1050-1256  min-MSE heap, best-neighbour merge, union-find, minSupport 2000
 *   findBlockMembership  AHCPlaneFitter.hpp:603-705    ERODE_ALL_BORDER, seeds of the region growing (rfQueue)
 *   floodFill            AHCPlaneFitter.hpp:546-594    ONE serial FIFO queue over the pixels
 *   final re-merge       AHCPlaneFitter.hpp:291-323    second ahCluster over the planes that met, plidmap
 * It is the same restatement as oracle/peac_oracle.py (which it replaces on the hot loops: that file is 2.5 s per frame in
 * pure Python; this one is a few ms, i.e. of the order of the real PEAC, "> 35 Hz VGA"), kept bit-compatible with it:
 * tests/test_peac_cpu.py asserts identical membership images on synthetic frames.
 *
 * PARITY UNPINNED by the reference (no tests / golden vectors; PCL and Eigen are un-vendored).  Restatement choices where
 * the reference is implementation-defined (all measure-zero on real data): std::set<PlaneSeg*> neighbour order (pointer
 * values, AHCPlaneSeg.hpp:166) -> node creation order; std::priority_queue / std::sort tie order -> creation order;
 * Eigen::SelfAdjointEigenSolver (eig33sym.hpp:45-51) -> cyclic Jacobi in double.
 *
 * Build: make -C oracle   (-> oracle/_build/libpeac_cpu.so)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define WIN 16
#define MIN_SUPPORT 2000
#define DEPTH_SIGMA 3e-6
#define STD_TOL_INIT 10.0
#define STD_TOL_MERGE 17.0
#define Z_NEAR 500.0
#define Z_FAR 6000.0
#define PI 3.14159265358979323846

static double t_mse(int init, double z)
{
    const double t = DEPTH_SIGMA * z * z + (init ? STD_TOL_INIT : STD_TOL_MERGE);
    return t * t;
}
static double t_ang_init(double z)
{
    const double a_near = 10.0 * PI / 180.0, a_far = 20.0 * PI / 180.0;
    const double cz = z < Z_NEAR ? Z_NEAR : (z > Z_FAR ? Z_FAR : z);
    const double factor = (a_far - a_near) / (Z_FAR - Z_NEAR);
    return cos(factor * cz + a_near - factor * Z_NEAR);
}

typedef struct {
    double st[9];   /* sx sy sz sxx syy szz sxy syz sxz */
    double c[3], n[3], mse;
    int N, rid, nouse;
    int *nb, nnb, cap; /* neighbour node ids, ascending (= creation order) */
} Seg;

/* smallest eigenpair of a symmetric 3x3 (cyclic Jacobi, double) */
static void eig_min(double A[3][3], double *lmin, double v[3])
{
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 60; ++sweep) {
        const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
        const double dia = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
        if (off <= 1e-40 * dia || off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < 3; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = cs * akp - sn * akq;
                    A[k][q] = sn * akp + cs * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = cs * apk - sn * aqk;
                    A[q][k] = sn * apk + cs * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = cs * vkp - sn * vkq;
                    V[k][q] = sn * vkp + cs * vkq;
                }
            }
    }
    int m = 0;
    if (A[1][1] < A[m][m]) m = 1;
    if (A[2][2] < A[m][m]) m = 2;
    *lmin = A[m][m];
    v[0] = V[0][m]; v[1] = V[1][m]; v[2] = V[2][m];
}

/* Stats::compute (AHCPlaneSeg.hpp:84-116) */
static void seg_compute(Seg *s)
{
    const double *t = s->st;
    const double sc = 1.0 / (double)s->N;
    s->c[0] = t[0] * sc; s->c[1] = t[1] * sc; s->c[2] = t[2] * sc;
    double K[3][3];
    K[0][0] = t[3] - t[0] * t[0] * sc; K[0][1] = t[6] - t[0] * t[1] * sc; K[0][2] = t[8] - t[0] * t[2] * sc;
    K[1][1] = t[4] - t[1] * t[1] * sc; K[1][2] = t[7] - t[1] * t[2] * sc;
    K[2][2] = t[5] - t[2] * t[2] * sc;
    K[1][0] = K[0][1]; K[2][0] = K[0][2]; K[2][1] = K[1][2];
    double l, v[3];
    eig_min(K, &l, v);
    if (v[0] * s->c[0] + v[1] * s->c[1] + v[2] * s->c[2] > 0) { v[0] = -v[0]; v[1] = -v[1]; v[2] = -v[2]; }
    s->n[0] = v[0]; s->n[1] = v[1]; s->n[2] = v[2];
    s->mse = l * sc;
}

static double seg_sim(const Seg *a, const Seg *b) { return fabs(a->n[0] * b->n[0] + a->n[1] * b->n[1] + a->n[2] * b->n[2]); }

static void nb_add(Seg *s, int id)
{
    int lo = 0, hi = s->nnb;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (s->nb[m] < id) lo = m + 1; else hi = m; }
    if (lo < s->nnb && s->nb[lo] == id) return;
    if (s->nnb == s->cap) { s->cap = s->cap ? 2 * s->cap : 8; s->nb = (int *)realloc(s->nb, sizeof(int) * s->cap); }
    memmove(s->nb + lo + 1, s->nb + lo, sizeof(int) * (s->nnb - lo));
    s->nb[lo] = id;
    s->nnb++;
}
static void nb_del(Seg *s, int id)
{
    int lo = 0, hi = s->nnb;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (s->nb[m] < id) lo = m + 1; else hi = m; }
    if (lo < s->nnb && s->nb[lo] == id) { memmove(s->nb + lo, s->nb + lo + 1, sizeof(int) * (s->nnb - lo - 1)); s->nnb--; }
}

typedef struct { Seg *a; int n, cap; } SegPool;
static int pool_new(SegPool *P)
{
    if (P->n == P->cap) { P->cap = P->cap ? 2 * P->cap : 4096; P->a = (Seg *)realloc(P->a, sizeof(Seg) * P->cap); }
    memset(&P->a[P->n], 0, sizeof(Seg));
    return P->n++;
}
static void connect(SegPool *P, int a, int b) { nb_add(&P->a[a], b); nb_add(&P->a[b], a); }
static void disconnect_all(SegPool *P, int a)
{
    Seg *s = &P->a[a];
    for (int k = 0; k < s->nnb; ++k) nb_del(&P->a[s->nb[k]], a);
    s->nnb = 0;
}

/* DisjointSet.hpp */
static int ds_find(int *parent, int x)
{
    while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; }
    return x;
}
static int ds_union(int *parent, int *size, int x, int y)
{
    const int xr = ds_find(parent, x), yr = ds_find(parent, y);
    if (xr == yr) return xr;
    if (size[xr] < size[yr]) { parent[xr] = yr; size[yr] += size[xr]; return yr; }
    parent[yr] = xr; size[xr] += size[yr];
    return xr;
}

/* binary min-heap on (mse, id) */
typedef struct { int *a; int n, cap; } Heap;
static int heap_less(const SegPool *P, int x, int y)
{
    const double mx = P->a[x].mse, my = P->a[y].mse;
    return mx < my || (mx == my && x < y);
}
static void heap_push(Heap *h, const SegPool *P, int id)
{
    if (h->n == h->cap) { h->cap = h->cap ? 2 * h->cap : 4096; h->a = (int *)realloc(h->a, sizeof(int) * h->cap); }
    int i = h->n++;
    h->a[i] = id;
    while (i > 0) {
        const int p = (i - 1) >> 1;
        if (!heap_less(P, h->a[i], h->a[p])) break;
        const int t = h->a[i]; h->a[i] = h->a[p]; h->a[p] = t;
        i = p;
    }
}
static int heap_pop(Heap *h, const SegPool *P)
{
    const int top = h->a[0];
    h->a[0] = h->a[--h->n];
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < h->n && heap_less(P, h->a[l], h->a[m])) m = l;
        if (r < h->n && heap_less(P, h->a[r], h->a[m])) m = r;
        if (m == i) break;
        const int t = h->a[i]; h->a[i] = h->a[m]; h->a[m] = t;
        i = m;
    }
    return top;
}

/* PlaneFitter::ahCluster (AHCPlaneFitter.hpp:1050-1256).  queue: node ids; extracted: output ids sorted by N (stable). */
static int ah_cluster(SegPool *P, const int *queue, int nq, int *parent, int *size, int *extracted, long *n_pops)
{
    const double SIM_MERGE = cos(15.0 * PI / 180.0);
    Heap h = {0, 0, 0};
    for (int k = 0; k < nq; ++k) heap_push(&h, P, queue[k]);
    int nex = 0;
    while (h.n) {
        const int p = heap_pop(&h, P);
        if (P->a[p].nouse) continue;
        if (n_pops) ++*n_pops;
        int have = 0, cand_nb = -1;
        Seg cand;
        memset(&cand, 0, sizeof cand);
        for (int k = 0; k < P->a[p].nnb; ++k) {
            const int o = P->a[p].nb[k];
            if (seg_sim(&P->a[p], &P->a[o]) < SIM_MERGE) continue;
            Seg m;
            memset(&m, 0, sizeof m);
            for (int d = 0; d < 9; ++d) m.st[d] = P->a[p].st[d] + P->a[o].st[d];
            m.N = P->a[p].N + P->a[o].N;
            m.rid = P->a[p].N >= P->a[o].N ? P->a[p].rid : P->a[o].rid;
            seg_compute(&m);
            if (!have || cand.mse > m.mse) { cand = m; cand_nb = o; have = 1; }
        }
        if (have && cand.mse < t_mse(0, cand.c[2])) {
            const int q = pool_new(P);   /* may move the pool: re-read pointers below */
            P->a[q] = cand;
            P->a[q].nb = NULL; P->a[q].nnb = 0; P->a[q].cap = 0;
            heap_push(&h, P, q);
            ds_union(parent, size, P->a[p].rid, P->a[cand_nb].rid);
            /* mergeNbsFrom (AHCPlaneSeg.hpp:357-386) */
            for (int k = 0; k < P->a[p].nnb; ++k) { const int o = P->a[p].nb[k]; if (o != cand_nb) nb_add(&P->a[q], o); }
            for (int k = 0; k < P->a[cand_nb].nnb; ++k) { const int o = P->a[cand_nb].nb[k]; if (o != p) nb_add(&P->a[q], o); }
            disconnect_all(P, p);
            disconnect_all(P, cand_nb);
            for (int k = 0; k < P->a[q].nnb; ++k) nb_add(&P->a[P->a[q].nb[k]], q);
            P->a[p].nouse = P->a[cand_nb].nouse = 1;
        } else {
            if (P->a[p].N >= MIN_SUPPORT) extracted[nex++] = p;
            disconnect_all(P, p);
        }
    }
    free(h.a);
    /* PlaneSegSizeCmp, stable */
    for (int a = 1; a < nex; ++a) {
        const int v = extracted[a];
        int b = a - 1;
        while (b >= 0 && P->a[extracted[b]].N < P->a[v].N) { extracted[b + 1] = extracted[b]; --b; }
        extracted[b + 1] = v;
    }
    return nex;
}

/*
 * pts: H x W x 3 float (organised cloud, NaN z where invalid; DynaDetect.cc:562-587).
 * member_out: H x W int, final plane ids or -1.  grown_out (optional): membership after floodFill, before the relabel.
 * blk_map_out (optional): Nh x Nw.  coarse_out (optional): (rid, N) pairs of the coarse planes, up to max_coarse pairs.
 * stats_out (optional, 8 longs): [0] pops of the first ahCluster, [1] seeds, [2] queue entries processed, [3] FIFO levels,
 * [4] largest level, [5] valid blocks, [6] planes after the first ahCluster, [7] final planes.
 * Returns the number of final planes, or -1 on allocation failure.
 */
int peac_plane_fit(const float *pts, int W, int H, int *member_out, int *grown_out, int *blk_map_out, int *coarse_out, int max_coarse,
                   int *n_coarse_out, long *stats_out)
{
    const int Nh = H / WIN, Nw = W / WIN, NB = Nh * Nw;
    const double SIM_REFINE = cos(20.0 * PI / 180.0);
    SegPool P = {0, 0, 0};
    int *parent = (int *)malloc(sizeof(int) * NB), *size = (int *)malloc(sizeof(int) * NB);
    int *G = (int *)malloc(sizeof(int) * NB), *queue = (int *)malloc(sizeof(int) * NB);
    int *extracted = (int *)malloc(sizeof(int) * 2 * NB);
    long n_pops = 0;
    int nq = 0;
    for (int b = 0; b < NB; ++b) { parent[b] = b; size[b] = 1; G[b] = -1; }
    /* ---- initGraph */
    for (int i = 0; i < Nh; ++i)
        for (int j = 0; j < Nw; ++j) {
            double st[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            int ok = 1;
            for (int y = 0; y < WIN && ok; ++y)
                for (int x = 0; x < WIN; ++x) {
                    const float *p = pts + 3 * ((size_t)(i * WIN + y) * W + j * WIN + x);
                    if (isnan(p[2])) { ok = 0; break; }
                    const double X = p[0], Y = p[1], Z = p[2];
                    st[0] += X; st[1] += Y; st[2] += Z;
                    st[3] += X * X; st[4] += Y * Y; st[5] += Z * Z;
                    st[6] += X * Y; st[7] += Y * Z; st[8] += X * Z;
                }
            if (!ok) continue;
            Seg s;
            memset(&s, 0, sizeof s);
            memcpy(s.st, st, sizeof st);
            s.N = WIN * WIN; s.rid = i * Nw + j;
            seg_compute(&s);
            if (s.mse < t_mse(1, s.c[2])) {
                const int id = pool_new(&P);
                P.a[id] = s;
                G[i * Nw + j] = id;
                queue[nq++] = id;
            }
        }
    const int n_valid_blocks = nq;
#define SIM(i0, i1) seg_sim(&P.a[i0], &P.a[i1])
    for (int i = 0; i < Nh; ++i) {
        int j = 1;
        while (j < Nw) {
            const int c = i * Nw + j;
            if (G[c - 1] < 0) { j += 1; continue; }
            if (G[c] < 0) { j += 2; continue; }
            if (j < Nw - 1 && G[c + 1] < 0) { j += 3; continue; }
            const double th = t_ang_init(P.a[G[c]].c[2]);
            if ((j < Nw - 1 && SIM(G[c - 1], G[c + 1]) >= th) || (j == Nw - 1 && SIM(G[c], G[c - 1]) >= th)) {
                connect(&P, G[c], G[c - 1]);
                if (j < Nw - 1) connect(&P, G[c], G[c + 1]);
                j += 2;
            } else j += 1;
        }
    }
    for (int j = 0; j < Nw; ++j) {
        int i = 1;
        while (i < Nh) {
            const int c = i * Nw + j;
            if (G[c - Nw] < 0) { i += 1; continue; }
            if (G[c] < 0) { i += 2; continue; }
            if (i < Nh - 1 && G[c + Nw] < 0) { i += 3; continue; }
            const double th = t_ang_init(P.a[G[c]].c[2]);
            if ((i < Nh - 1 && SIM(G[c - Nw], G[c + Nw]) >= th) || (i == Nh - 1 && SIM(G[c], G[c - Nw]) >= th)) {
                connect(&P, G[c], G[c - Nw]);
                if (i < Nh - 1) connect(&P, G[c], G[c + Nw]);
                i += 2;
            } else i += 1;
        }
    }
    const int nex = ah_cluster(&P, queue, nq, parent, size, extracted, &n_pops);
    if (n_coarse_out) *n_coarse_out = nex;
    if (coarse_out)
        for (int k = 0; k < nex && k < max_coarse; ++k) { coarse_out[2 * k] = P.a[extracted[k]].rid; coarse_out[2 * k + 1] = P.a[extracted[k]].N; }
    /* ---- findBlockMembership */
    int *rid2plid = (int *)malloc(sizeof(int) * NB);
    for (int b = 0; b < NB; ++b) rid2plid[b] = -1;
    for (int k = 0; k < nex; ++k) rid2plid[P.a[extracted[k]].rid] = k;
    int *member = grown_out ? grown_out : (int *)malloc(sizeof(int) * (size_t)W * H);
    for (size_t i = 0; i < (size_t)W * H; ++i) member[i] = -1;
    int *blk_map = (int *)malloc(sizeof(int) * NB);
    int *is_valid = (int *)calloc(nex + 1, sizeof(int));
    size_t rf_cap = (size_t)W * H + 1024, rf_n = 0;
    int *rf_idx = (int *)malloc(sizeof(int) * rf_cap), *rf_pl = (int *)malloc(sizeof(int) * rf_cap);
#define RF_PUSH(idx, pl)                                                                   \
    do {                                                                                   \
        if (rf_n == rf_cap) {                                                              \
            rf_cap *= 2;                                                                   \
            rf_idx = (int *)realloc(rf_idx, sizeof(int) * rf_cap);                         \
            rf_pl = (int *)realloc(rf_pl, sizeof(int) * rf_cap);                           \
        }                                                                                  \
        rf_idx[rf_n] = (idx); rf_pl[rf_n] = (pl); ++rf_n;                                  \
    } while (0)
    for (int b = 0; b < NB; ++b) blk_map[b] = -1;
    for (int i = 0; i < Nh; ++i)
        for (int j = 0; j < Nw; ++j) {
            const int b = i * Nw + j;
            const int setid = ds_find(parent, b);
            if (size[setid] * WIN * WIN >= MIN_SUPPORT) {
                int same = 1;
                if (j > 0) same &= ds_find(parent, b - 1) == setid;
                if (j < Nw - 1) same &= ds_find(parent, b + 1) == setid;
                if (i > 0) same &= ds_find(parent, b - Nw) == setid;
                if (i < Nh - 1) same &= ds_find(parent, b + Nw) == setid;
                const int plid = rid2plid[setid];
                if (same && plid >= 0) {
                    blk_map[b] = plid;
                    for (int y = 0; y < WIN; ++y)
                        for (int x = 0; x < WIN; ++x) member[(size_t)(i * WIN + y) * W + j * WIN + x] = plid;
                    is_valid[plid] = 1;
                }
            }
            if (blk_map[b] < 0) {
                if (i > 0 && blk_map[b - Nw] >= 0) {
                    const int sp = (i * WIN - 1) * W + j * WIN;
                    for (int k = 1; k < WIN; ++k) RF_PUSH(sp + k, blk_map[b - Nw]);
                }
                if (j > 0 && blk_map[b - 1] >= 0) {
                    const int sp = (i * WIN) * W + j * WIN - 1;
                    for (int k = 0; k < WIN - 1; ++k) RF_PUSH(sp + k * W, blk_map[b - 1]);
                }
            } else {
                const int plid = blk_map[b];
                if (i > 0 && blk_map[b - Nw] != plid) {
                    const int sp = (i * WIN) * W + j * WIN;
                    for (int k = 0; k < WIN - 1; ++k) RF_PUSH(sp + k, plid);
                }
                if (j > 0 && blk_map[b - 1] != plid) {
                    const int sp = (i * WIN) * W + j * WIN;
                    for (int k = 1; k < WIN; ++k) RF_PUSH(sp + k * W, plid);
                }
            }
        }
    if (blk_map_out) memcpy(blk_map_out, blk_map, sizeof(int) * NB);
    const long n_seeds = (long)rf_n;
    /* ---- floodFill */
    float *dist_map = (float *)malloc(sizeof(float) * (size_t)W * H);
    for (size_t i = 0; i < (size_t)W * H; ++i) dist_map[i] = 3.402823466e+38f;
    long levels = 0, max_level = 0;
    size_t level_end = rf_n, level_begin = 0;
    for (size_t k = 0; k < rf_n; ++k) {
        if (k == level_end) {
            if ((long)(level_end - level_begin) > max_level) max_level = (long)(level_end - level_begin);
            ++levels; level_begin = level_end; level_end = rf_n;
        }
        const int s_idx = rf_idx[k], plid = rf_pl[k];
        const Seg *pl = &P.a[extracted[plid]];
        const double thr = 9.0 * pl->mse + 1e-5;
        const int sy = s_idx / W, sx = s_idx - sy * W;
        int nb[4], nn = 0;
        if (sx > 0) nb[nn++] = s_idx - 1;
        if (sx < W - 1) nb[nn++] = s_idx + 1;
        if (sy > 0) nb[nn++] = s_idx - W;
        if (sy < H - 1) nb[nn++] = s_idx + W;
        for (int t = 0; t < nn; ++t) {
            const int c = nb[t];
            const int trail = member[c];
            if (trail <= -6) continue;
            if (trail >= 0 && trail == plid) continue;
            const int cy = c / W, cx = c - cy * W;
            const int bx = cx / WIN, by = cy / WIN;
            if (by < Nh && bx < Nw && blk_map[by * Nw + bx] >= 0) continue;
            int ok = 0;
            float cdist = -1.0f;
            const float *p = pts + 3 * (size_t)c;
            if (!isnan(p[2])) {
                const double d = pl->n[0] * ((double)p[0] - pl->c[0]) + pl->n[1] * ((double)p[1] - pl->c[1]) + pl->n[2] * ((double)p[2] - pl->c[2]);
                cdist = (float)fabs(d);
                ok = (double)cdist * (double)cdist < thr;
            }
            if (ok) {
                if (trail >= 0) {
                    const int a = extracted[trail], b2 = extracted[plid];
                    if (seg_sim(&P.a[b2], &P.a[a]) >= SIM_REFINE) connect(&P, a, b2);
                }
                if (cdist < dist_map[c]) {
                    member[c] = plid;
                    dist_map[c] = cdist;
                    RF_PUSH(c, plid);
                } else if (trail < 0) member[c] = trail - 1;
            } else if (trail < 0) member[c] = trail - 1;
        }
    }
    if ((long)(rf_n - level_begin) > max_level) max_level = (long)(rf_n - level_begin);
    if (rf_n > 0) ++levels;
    /* ---- final merge of the planes that met during region growing (AHCPlaneFitter.hpp:291-323) */
    int *q2 = (int *)malloc(sizeof(int) * (nex + 1)), nq2 = 0;
    for (int k = 0; k < nex; ++k)
        if (is_valid[k]) q2[nq2++] = extracted[k];
    int *final_ex = (int *)malloc(sizeof(int) * 2 * (nex + 1));
    ah_cluster(&P, q2, nq2, parent, size, final_ex, NULL);
    int *plidmap = (int *)malloc(sizeof(int) * (nex + 1));
    for (int k = 0; k < nex; ++k) plidmap[k] = -1;
    int n_final = 0;
    for (int i = 0; i < nex; ++i) {
        if (!is_valid[i]) continue;
        const int rid = P.a[extracted[i]].rid;
        const int root = ds_find(parent, rid);
        if (root == rid) {
            if (plidmap[i] < 0) plidmap[i] = n_final++;
        } else {
            const int npid = rid2plid[root];
            if (plidmap[npid] < 0) { plidmap[i] = plidmap[npid] = n_final++; }
            else plidmap[i] = plidmap[npid];
        }
    }
    for (size_t i = 0; i < (size_t)W * H; ++i) member_out[i] = member[i] >= 0 ? plidmap[member[i]] : -1;
    if (stats_out) {
        stats_out[0] = n_pops; stats_out[1] = n_seeds; stats_out[2] = (long)rf_n; stats_out[3] = levels; stats_out[4] = max_level;
        stats_out[5] = n_valid_blocks; stats_out[6] = nex; stats_out[7] = n_final;
    }
    for (int k = 0; k < P.n; ++k) free(P.a[k].nb);
    free(P.a); free(parent); free(size); free(G); free(queue); free(extracted); free(rid2plid);
    if (!grown_out) free(member);
    free(blk_map); free(is_valid); free(rf_idx); free(rf_pl); free(dist_map); free(q2); free(final_ex); free(plidmap);
    return n_final;
}
