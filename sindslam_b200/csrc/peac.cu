// peac.cu -- PEAC plane-contour edges of DynaDetect::CalOccluded (ORB_SLAM2/src/DynaDetect.cc:558-593) on the device:
// the agglomerative hierarchical clustering plane fitter of ORB_SLAM2/include/PEAC (AHCPlaneFitter.hpp, AHCPlaneSeg.hpp,
// AHCParamSet.hpp, DisjointSet.hpp) with the parameters actually in force (16x16 blocks, INIT_STRICT, minSupport 2000,
// metre coordinates against mm-tuned thresholds), followed by per-plane CLOSE 3x3 + external contours
// drawn with thickness 2 (AHCPlaneFitter.hpp:366-399).
//
//   k_peac_blocks      one thread per 16x16 block: the nine second-order sums in double, accumulated in the reference's
//                      raster order (bit-identical statistics), PCA plane (Jacobi 3x3) -> node table
//   k_peac_ahc         ONE CTA: initial graph edges, min-MSE binary heap in shared memory, union-find by size.  The serial
//                      control (pop / merge / extract) runs on thread 0; the candidate evaluation of a pop (all graph
//                      neighbours: merged statistics + eigen solve) is spread over the CTA's threads.
//   k_peac_blockmap    block erosion (ERODE_ALL_BORDER) and the initial membership image
//   k_peac_grow        pixel-level region growing as a parallel min-distance label relaxation (tile-local iterations in
//                      shared memory).  The reference grows with ONE serial FIFO queue (AHCPlaneFitter.hpp:546-594), whose
//                      result depends on the visiting order at pixels contested by two planes; the relaxation converges
//                      to the order-independent fixed point instead (documented deviation; plane-edge
//                      agreement with the serial oracle is measured in tests/test_peac_gpu.py).
//   k_peac_merge       final re-merge of planes that met during growing (serial, <= 64 planes) + relabel map
//   plane_contours_run (recluster.cu) CLOSE 3x3, external contours, thickness-2 drawing on membership bitsets
#include "peac.cuh"

#include "recluster.cuh"

#include <cooperative_groups.h>

#define PEAC_WIN 16
#define PEAC_MAXB 1600          // blocks (53 x 30 = 1590 at 848 x 480); bounded by the shared memory of k_peac_ahc
#define PEAC_MAXP 64            // extracted planes
#define PEAC_MIN_SUPPORT 2000
#define PEAC_HEAP (2 * PEAC_MAXB)
#define PEAC_MAXCAND 1024       // distinct graph neighbours of one node
#define PEAC_MAXE (2 * PEAC_MAXB)   // the row pass and the column pass add at most one edge per block each

struct PeacNode {               // ahc::PlaneSeg (AHCPlaneSeg.hpp:29-188), indexed by root block id
    double st[9];               // sx sy sz sxx syy szz sxy syz sxz
    double center[3], normal[3], mse;
    int N, version, alive, valid;
};

struct PeacPlane { double center[3], normal[3], mse, thr; int N, rid, valid, final_id; double st[9]; };

#define PEAC_AHC_CTAS_MAX 64
struct PeacExtract { double st[9], normal[3], mse, pop_key; int N, rid; };   // one extracted plane of one graph component, unsorted

struct PeacControl {
    int n_planes, n_final, overflow, n_ex;      // header (16 ints), zeroed every frame
    int done, n_comp, grow_levels, grow_entries;
    int grow_n, grow_buf, hdr_pad[6];           // frontier handed from the cluster kernel to the single-CTA kernel
    int grow_tot[16];                           // per-CTA counts of the cluster kernel's ordered compaction
    int sticky_overflow;                        // never reset: an overflow of ANY frame this instance has processed (frame pipeline: the
                                                // header above may already belong to a later frame when the host looks)
    long long clk[PEAC_AHC_CTAS_MAX][12];      // PEAC_CLOCKS builds only
    int parent[PEAC_MAXB], size[PEAC_MAXB];
    int blk_map[PEAC_MAXB];
    PeacExtract ex[PEAC_MAXP];
    PeacPlane pl[PEAC_MAXP];
    unsigned long long conn[PEAC_MAXP];      // planes that met during region growing with similar normals
};

#define PG_CAP 131072          // entries of one FIFO level of the region growing (observed: <= 12 k at 848 x 480)

struct PeacImpl {
    PeacNode *nodes = nullptr;
    PeacControl *ctl = nullptr;
    int *label = nullptr;        // membershipImg
    float *dist = nullptr;
    int *head = nullptr;         // per pixel: head of the list of this level's visits
    int *qa = nullptr, *qb = nullptr;                                        // FIFO level n / n + 1: (plane << 20) | pixel
    struct PgRec *rec = nullptr;                                            // visit records of the large levels (cluster kernel)
    int *v_pix = nullptr, *v_info = nullptr, *v_next = nullptr;              // visit records of one level, slot = 4 * entry + direction
    float *v_dist = nullptr;
    unsigned char *v_push = nullptr;
    ulonglong2 *PB = nullptr;    // final plane membership bitset
    int Nw = 0, Nh = 0;
};

// ---------------------------------------------------------------- small dense eigen solver
// Smallest eigenpair of a symmetric positive semi-definite 3x3 (reference: Eigen::SelfAdjointEigenSolver through
// LA::eig33sym, eig33sym.hpp:45-51 -- un-vendored).  The clustering loop is a serial chain of ~1000 candidate evaluations
// that need the smallest eigenVALUE only, so the two halves are separate:
//   eig33_min_val: Newton's method on the characteristic cubic p(l) = det(K - l I) from l = 0.  For a PSD matrix p is
//                  positive, decreasing and convex on (-inf, l_min], so the iteration rises monotonically to l_min and
//                  converges quadratically (3-5 steps, ~10 FP64 operations each; the reciprocal of p' is a float
//                  approximation with one FP64 Newton correction);
//   eig33_vec:     the eigenvector from the best-conditioned cross product of two rows of K - l I (winner only).
// The smallest eigenvalue agrees with numpy.linalg.eigh / cyclic Jacobi to ~1e-11 relative for plane-like covariances.
struct Sym3 { double a00, a11, a22, a01, a02, a12; };

__device__ __forceinline__ double peac_rcp(double x)
{
    const double r = (double)__frcp_rn((float)x);
    return r * (2.0 - x * r);                  // relative error ~1e-14
}

__device__ __forceinline__ double eig33_min_val(const Sym3 &K)
{
    const double m0 = K.a11 * K.a22 - K.a12 * K.a12, m1 = K.a00 * K.a22 - K.a02 * K.a02, m2 = K.a00 * K.a11 - K.a01 * K.a01;
    const double c2 = K.a00 + K.a11 + K.a22;
    const double c1 = m0 + m1 + m2;
    const double c0 = K.a00 * m0 - K.a01 * (K.a01 * K.a22 - K.a12 * K.a02) + K.a02 * (K.a01 * K.a12 - K.a11 * K.a02);
    // three float iterations from 0 (cheap: the FP64 chain below costs ~50 cycles per dependent operation), then FP64 to convergence.
    // A float iterate may land slightly right of the root; p < 0 and p' < 0 there, so the FP64 steps walk back onto it.
    float lf = 0.0f;
    {
        const float f2 = (float)c2, f1 = (float)c1, f0 = (float)c0;
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const float pf = ((f2 - lf) * lf - f1) * lf + f0;
            const float df = (2.0f * f2 - 3.0f * lf) * lf - f1;
            if (df < 0.0f) lf -= pf * __frcp_rn(df);
        }
        if (!(lf >= 0.0f) || !(lf < f2)) lf = 0.0f;     // NaN / runaway guard: restart from the safe side
    }
    double l = (double)lf;
#pragma unroll 1
    for (int it = 0; it < 12; ++it) {
        const double p = ((c2 - l) * l - c1) * l + c0;
        const double dp = (2.0 * c2 - 3.0 * l) * l - c1;
        if (!(dp < 0.0)) break;                  // flat cubic (rank-deficient block): keep the current iterate
        const double dl = p * peac_rcp(dp);
        l -= dl;
        if (fabs(dl) <= 1e-8 * fabs(l)) break;   // quadratic convergence: the iterate just written is already exact to ~1e-16
    }
    return l;
}

__device__ __forceinline__ void eig33_vec(const Sym3 &K, double l, double v[3])
{
    // rows of K - l I; the eigenvector is the best-conditioned cross product of two of them.  Scalars and selects only: a
    // pointer into local arrays would put them into local memory, which is an L2 round trip in the kernels that carve out
    // almost all of the L1 for shared memory.
    const double r00 = K.a00 - l, r01 = K.a01, r02 = K.a02, r11 = K.a11 - l, r12 = K.a12, r22 = K.a22 - l;
    const double ax = r01 * r12 - r02 * r11, ay = r02 * r01 - r00 * r12, az = r00 * r11 - r01 * r01;     // row0 x row1
    const double bx = r01 * r22 - r02 * r12, by = r02 * r02 - r00 * r22, bz = r00 * r12 - r01 * r02;     // row0 x row2
    const double cx = r11 * r22 - r12 * r12, cy = r12 * r02 - r01 * r22, cz = r01 * r12 - r11 * r02;     // row1 x row2
    const double n0 = ax * ax + ay * ay + az * az, n1 = bx * bx + by * by + bz * bz, n2 = cx * cx + cy * cy + cz * cz;
    double x = ax, y = ay, z = az, nn = n0;
    if (n1 > nn) { x = bx; y = by; z = bz; nn = n1; }
    if (n2 > nn) { x = cx; y = cy; z = cz; nn = n2; }
    if (nn <= 0.0) { v[0] = 0; v[1] = 0; v[2] = 1; return; }
    const double in = rsqrt(nn);
    v[0] = x * in; v[1] = y * in; v[2] = z * in;
}

// scatter matrix of Stats::compute (AHCPlaneSeg.hpp:84-116)
__device__ __forceinline__ void peac_cov(const double st[9], double sc, double center[3], Sym3 &K)
{
    center[0] = st[0] * sc; center[1] = st[1] * sc; center[2] = st[2] * sc;
    K.a00 = st[3] - st[0] * st[0] * sc; K.a01 = st[6] - st[0] * st[1] * sc; K.a02 = st[8] - st[0] * st[2] * sc;
    K.a11 = st[4] - st[1] * st[1] * sc; K.a12 = st[7] - st[1] * st[2] * sc;
    K.a22 = st[5] - st[2] * st[2] * sc;
}

// Stats::compute (AHCPlaneSeg.hpp:84-116)
__device__ void peac_compute(const double st[9], int N, double center[3], double normal[3], double &mse)
{
    const double sc = 1.0 / (double)N;
    Sym3 K;
    peac_cov(st, sc, center, K);
    const double l = eig33_min_val(K);
    double v[3];
    eig33_vec(K, l, v);
    const double d = v[0] * center[0] + v[1] * center[1] + v[2] * center[2];
    const double sgn = d <= 0 ? 1.0 : -1.0;     // normal points towards the camera
    normal[0] = sgn * v[0]; normal[1] = sgn * v[1]; normal[2] = sgn * v[2];
    mse = l * sc;
}

__device__ __forceinline__ double peac_t_mse(bool init, double z)
{
    const double t = 3e-6 * z * z + (init ? 10.0 : 17.0);
    return t * t;
}
__device__ __forceinline__ double peac_t_ang_init(double z)
{
    const double z_near = 500.0, z_far = 6000.0, a_near = 10.0 * 3.14159265358979323846 / 180.0, a_far = 20.0 * 3.14159265358979323846 / 180.0;
    const double cz = fmin(fmax(z, z_near), z_far);
    const double factor = (a_far - a_near) / (z_far - z_near);
    return cos(factor * cz + a_near - factor * z_near);
}
#define PEAC_SIM_MERGE 0.96592582628906831   /* cos 15 deg */
#define PEAC_SIM_REFINE 0.93969262078590843  /* cos 20 deg */

// organised cloud point (DynaDetect.cc:562-587): float arithmetic, then widened to double by OrganizedImage3D::get
__device__ __forceinline__ bool peac_point(const uint16_t *__restrict__ depth, int W, int x, int y, float fx, float fy, float cx, float cy,
                                           float inv_scale, double p[3])
{
    const float d = (float)depth[y * W + x];
    if (d < 1e-3f) return false;
    const float z = d * inv_scale;
    const float px = ((float)x - cx) * z / fx, py = ((float)y - cy) * z / fy;
    p[0] = px; p[1] = py; p[2] = z;
    return true;
}

// ---------------------------------------------------------------- block statistics (PlaneSeg constructor)
__global__ void k_peac_blocks(const uint16_t *__restrict__ depth, int W, int H, float fx, float fy, float cx, float cy, float inv_scale, int Nw, int Nh,
                              PeacNode *__restrict__ nodes)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= Nw * Nh) return;
    const int by = b / Nw, bx = b - by * Nw;
    double st[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    bool ok = true;
    for (int i = 0; i < PEAC_WIN && ok; ++i)
        for (int j = 0; j < PEAC_WIN; ++j) {
            double p[3];
            if (!peac_point(depth, W, bx * PEAC_WIN + j, by * PEAC_WIN + i, fx, fy, cx, cy, inv_scale, p)) { ok = false; break; }   // INIT_STRICT
            // depthDisContinuous: |dz| > 0.04 |z| + 20 can never hold for metre coordinates
            st[0] += p[0]; st[1] += p[1]; st[2] += p[2];
            st[3] += p[0] * p[0]; st[4] += p[1] * p[1]; st[5] += p[2] * p[2];
            st[6] += p[0] * p[1]; st[7] += p[1] * p[2]; st[8] += p[0] * p[2];
        }
    PeacNode n;
    for (int k = 0; k < 9; ++k) n.st[k] = ok ? st[k] : 0.0;
    n.N = ok ? PEAC_WIN * PEAC_WIN : 0;
    n.version = 0;
    n.mse = 0; n.center[0] = n.center[1] = n.center[2] = 0; n.normal[0] = n.normal[1] = n.normal[2] = 0;
    bool valid = ok;
    if (ok) {
        peac_compute(n.st, n.N, n.center, n.normal, n.mse);
        valid = n.mse < peac_t_mse(true, n.center[2]);
    }
    n.valid = valid ? 1 : 0;
    n.alive = valid ? 1 : 0;
    nodes[b] = n;
}

// unsigned key with the order of the double (x == y <=> same key, -0 folded into +0): lets the arg-min use integer warp reductions
__device__ __forceinline__ unsigned long long peac_key(double m)
{
    if (m == 0.0) m = 0.0;
    const unsigned long long b = (unsigned long long)__double_as_longlong(m);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// ---------------------------------------------------------------- AHC (Algorithm 2 edges + Algorithm 3 clustering)
struct MergeResult { double st[9], c[3], n[3], mse; };

#define PEAC_AHC_SMEM ((sizeof(double) * (3 + 1 + 9) + sizeof(int) * 2 + 1 + sizeof(unsigned short) * 2) * PEAC_MAXB + sizeof(unsigned short) * 2 * PEAC_MAXE + \
                       sizeof(unsigned short) * 4 * PEAC_MAXB + 64)
#define PF_ALIVE 1
#define PF_VALID 2
#define PF_QUEUED 4

// The pop / merge / extract loop is a chain of ~1000 dependent steps per frame, so everything it touches sits in shared
// memory (structure-of-arrays, ~200 KB) and every step is spread over the CTA:
//   * "pop the min-MSE node" is a parallel arg-min over the live queued nodes (cheaper than a binary heap walked by one
//     thread, no stale entries): every thread caches the minimum of the nodes it owns and re-scans them only after one of
//     them changed, warps reduce with three integer redux operations on an order-preserving key; ties resolve by creation
//     order like the oracle's insertion-ordered queue;
//   * the union-find is kept FLAT (root[] of every block is rewritten on each merge by all threads);
//   * every thread owns a column of the edge array and compacts it lazily (edges that became internal to a node or touch
//     an extracted node are dropped when met), so the neighbour scan shrinks as the clustering proceeds;
//   * the distinct neighbours of the popped node are collected first (stamp + compaction), then one candidate merge
//     (merged sums + eigen solve) runs per thread; the winning thread hands its merged plane over through shared memory;
//   * the bookkeeping of a merge is done by warp 0 (shuffle arg-min over the 8 warp results, lanes copy the sums).
// developer instrumentation (SINDYN_NVCC_EXTRA=-DPEAC_CLOCKS): cycle counts of the clustering loop's phases, thread 0 of every CTA
#ifdef PEAC_CLOCKS
#define PCLK(i) do { if (tid == 0) { const long long t_ = clock64(); clk[i] += t_ - t_last; t_last = t_; } } while (0)
#else
#define PCLK(i) do { } while (0)
#endif
#define PEAC_AHC_NT 512   // threads of k_peac_ahc (a power of two: node b is owned by thread b & (PEAC_AHC_NT - 1))
#define PEAC_BATCH 2      // queue heads processed per round (PEAC_AHC_NT / PEAC_BATCH threads evaluate the candidates of one of them)
#define PEAC_CANDK 512    // distinct graph neighbours of one node
#define PEAC_WTOP 2        // queue heads every warp contributes to the selection of a round (16 warps x 2 = 32 lanes)
// (measured on the B200, bit-exact in every variant: 256 threads / batch 4 / 3 heads per warp 1.77 ms; 256 / 2 / 2 1.64; 128 / 2 / 2
//  1.97; 512 / 4 / 2 1.63; 512 / 2 / 2 1.55 -- a round commits ~2 pops whatever the batch size)
#define PEAC_AHC_CTAS 32  // CTAs of k_peac_ahc: one connected component of the block graph each (more components: round robin)

// Edges exist only between blocks of one connected component of the initial graph and every later edge is inherited from
// them, so the clustering of one component never reads or writes another one: the global pop order of the reference's
// single queue only INTERLEAVES the components.  Every CTA therefore rebuilds the (tiny) initial graph, labels its
// connected components and clusters the components blockIdx.x, blockIdx.x + gridDim.x, ... on its own SM; components of
// fewer than minSupport / 256 blocks can never reach minSupport and are skipped.  The serial chain of a frame shrinks from
// "all valid blocks" (~1000 pops) to "the blocks of the largest component".  The order in which the reference extracts
// planes (it only matters for the stable size sort, AHCPlaneFitter.hpp:1251-1254) is recovered from the MSE of the
// extracted node: the popped MSEs of the single queue are non-decreasing (a merged node is never better than the popped
// minimum it contains), so "extraction order" == "ascending pop MSE".
__global__ void __launch_bounds__(PEAC_AHC_NT) k_peac_ahc(const PeacNode *__restrict__ nodes, PeacControl *ctl, int Nw, int Nh)
{
    extern __shared__ unsigned char smraw[];
    double *nrm = (double *)smraw;                                       // 3 x PEAC_MAXB
    double *mse_a = nrm + 3 * PEAC_MAXB;                                 // PEAC_MAXB
    double *sst = mse_a + PEAC_MAXB;                                     // 9 x PEAC_MAXB second-order sums
    int *N_a = (int *)(sst + 9 * PEAC_MAXB);                             // PEAC_MAXB
    int *seq_a = N_a + PEAC_MAXB;                                        // PEAC_MAXB
    unsigned short *root = (unsigned short *)(seq_a + PEAC_MAXB);        // PEAC_MAXB: flat union-find
    unsigned short *ssize = root + PEAC_MAXB;                            // PEAC_MAXB
    unsigned short *eu = ssize + PEAC_MAXB, *ev = eu + PEAC_MAXE;        // edges, column t = entries t, t + 256, ...
    unsigned char *flags = (unsigned char *)(ev + PEAC_MAXE);            // PEAC_MAXB
    // PEAC_BATCH x PEAC_MAXB round stamps ("o is a neighbour of the batch's j-th node in round r"); the same bytes hold the
    // component labels (int) while the graph is set up
    unsigned short *stampK = (unsigned short *)(flags + PEAC_MAXB);
    int *comp = (int *)stampK;
    __shared__ int s_ne, s_seq, s_nex, s_changed, s_nactive, s_nmerge, s_lose[PEAC_BATCH], s_win[PEAC_BATCH], s_ncand[PEAC_BATCH], s_guard[PEAC_BATCH];
    __shared__ unsigned short cand[PEAC_BATCH][PEAC_CANDK];
    __shared__ unsigned long long w4_key[PEAC_AHC_NT / 32][PEAC_BATCH];
    __shared__ int w4_p[PEAC_AHC_NT / 32][PEAC_BATCH], w4_seq[PEAC_AHC_NT / 32][PEAC_BATCH];
    constexpr int WPS = PEAC_AHC_NT / PEAC_BATCH / 32;      // warps that evaluate the candidates of one batch node
    __shared__ double ws_mse[PEAC_BATCH][WPS], ws_c2[PEAC_BATCH][WPS], pl_M[PEAC_BATCH], pl_mp[PEAC_BATCH];
    __shared__ int pl_ext[PEAC_BATCH], pl_clash[PEAC_BATCH];
    __shared__ int pl_O[PEAC_BATCH], pl_MG[PEAC_BATCH], pl_seq[PEAC_BATCH], pl_ex[PEAC_BATCH], pl_L, pl_nm, pl_ne, s_P[PEAC_BATCH];
    __shared__ int ws_o[PEAC_BATCH][WPS];
    __shared__ unsigned short s_active[PEAC_MAXB / 8 + 1];
    __shared__ int s_ex[PEAC_MAXP];
    __shared__ double s_exkey[PEAC_MAXP];
    const int tid = threadIdx.x, nt = PEAC_AHC_NT, lane = tid & 31, wid = tid >> 5;
    const int NB = Nw * Nh;
#ifdef PEAC_CLOCKS
    long long clk[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, t_last = clock64();
#endif
    auto sim = [&](int a, int b) { return fabs(nrm[3 * a] * nrm[3 * b] + nrm[3 * a + 1] * nrm[3 * b + 1] + nrm[3 * a + 2] * nrm[3 * b + 2]); };
    auto add_edge = [&](int a, int b) {
        int e = atomicAdd(&s_ne, 1);
        if (e < PEAC_MAXE) { eu[e] = (unsigned short)a; ev[e] = (unsigned short)b; }
    };
    for (int round = 0;; ++round) {
        __syncthreads();
        for (int b = tid; b < NB; b += nt) {
            root[b] = (unsigned short)b; ssize[b] = 1; comp[b] = b;
            nrm[3 * b] = nodes[b].normal[0]; nrm[3 * b + 1] = nodes[b].normal[1]; nrm[3 * b + 2] = nodes[b].normal[2];
            mse_a[b] = nodes[b].mse; N_a[b] = nodes[b].N; seq_a[b] = 0;
            flags[b] = nodes[b].valid ? PF_VALID : 0;
            for (int k = 0; k < 9; ++k) sst[b * 9 + k] = nodes[b].st[k];
        }
        if (tid == 0) { s_ne = 0; s_seq = NB; s_nex = 0; s_nmerge = 0; s_nactive = 0; }
        __syncthreads();
        auto valid = [&](int b) { return (flags[b] & PF_VALID) != 0; };
        // initGraph edges (AHCPlaneFitter.hpp:958-1014): rows and columns are independent, one thread each
        if (tid < Nh) {
            const int i = tid;
            for (int j = 1; j < Nw; j += 2) {
                const int c = i * Nw + j;
                if (!valid(c - 1)) { --j; continue; }
                if (!valid(c)) continue;
                if (j < Nw - 1 && !valid(c + 1)) { ++j; continue; }
                const double th = peac_t_ang_init(nodes[c].center[2]);
                if ((j < Nw - 1 && sim(c - 1, c + 1) >= th) || (j == Nw - 1 && sim(c, c - 1) >= th)) {
                    add_edge(c, c - 1);
                    if (j < Nw - 1) add_edge(c, c + 1);
                } else --j;
            }
        }
        if (tid >= 64 && tid - 64 < Nw) {
            const int j = tid - 64;
            for (int i = 1; i < Nh; i += 2) {
                const int c = i * Nw + j;
                if (!valid(c - Nw)) { --i; continue; }
                if (!valid(c)) continue;
                if (i < Nh - 1 && !valid(c + Nw)) { ++i; continue; }
                const double th = peac_t_ang_init(nodes[c].center[2]);
                if ((i < Nh - 1 && sim(c - Nw, c + Nw) >= th) || (i == Nh - 1 && sim(c, c - Nw) >= th)) {
                    add_edge(c, c - Nw);
                    if (i < Nh - 1) add_edge(c, c + Nw);
                } else --i;
            }
        }
        __syncthreads();
        const int NE = min(s_ne, PEAC_MAXE);
        if (tid == 0 && s_ne > PEAC_MAXE) ctl->overflow = ctl->sticky_overflow = 1;
        PCLK(0);   // load + initial edges
        // ---- connected components of the initial graph: minimum-label propagation over the edges + pointer jumping
        for (;;) {
            __syncthreads();
            if (tid == 0) s_changed = 0;
            __syncthreads();
            bool ch = false;
            for (int e = tid; e < NE; e += nt) {
                const int u = eu[e], v = ev[e];
                const int a = comp[u], b = comp[v];
                if (a < b) { atomicMin(&comp[v], a); atomicMin(&comp[b], a); ch = true; }
                else if (b < a) { atomicMin(&comp[u], b); atomicMin(&comp[a], b); ch = true; }
            }
            if (ch) s_changed = 1;
            __syncthreads();
            for (int b = tid; b < NB; b += nt) {
                int c = comp[b];
                while (comp[c] != c) c = comp[c];
                comp[b] = c;     // racing writers only ever move a label further down its own chain
            }
            __syncthreads();
            if (!s_changed) break;
        }
        // component sizes (seq_a as scratch) and the ascending list of components that can reach minSupport
        for (int b = tid; b < NB; b += nt)
            if (valid(b)) atomicAdd(&seq_a[comp[b]], 1);
        __syncthreads();
        if (wid == 0) {
            int n = 0;
            for (int b0 = 0; b0 < NB; b0 += 32) {
                const int b = b0 + lane;
                const bool is = b < NB && valid(b) && comp[b] == b && seq_a[b] * PEAC_WIN * PEAC_WIN >= PEAC_MIN_SUPPORT;
                const unsigned m = __ballot_sync(0xffffffffu, is);
                if (is) s_active[n + __popc(m & ((1u << lane) - 1))] = (unsigned short)b;
                n += __popc(m);
            }
            if (lane == 0) s_nactive = n;
        }
        __syncthreads();
        const int n_active = s_nactive;
        PCLK(1);   // components
        if (round == 0 && blockIdx.x == 0) {
            // blocks outside the clustered components stay singletons (their sets can never reach minSupport)
            for (int b = tid; b < NB; b += nt) {
                const int c = comp[b];
                if (!(valid(b) && seq_a[c] * PEAC_WIN * PEAC_WIN >= PEAC_MIN_SUPPORT)) { ctl->parent[b] = b; ctl->size[b] = 1; }
            }
            if (tid == 0) ctl->n_comp = n_active;
        }
        const int ci = blockIdx.x + round * gridDim.x;
        if (ci >= n_active) break;
        const int my_comp = s_active[ci];
        __syncthreads();
        for (int b = tid; b < NB; b += nt) {
            const bool mine = valid(b) && comp[b] == my_comp;
            flags[b] = mine ? (PF_ALIVE | PF_VALID | PF_QUEUED) : 0;
            seq_a[b] = b;
        }
        __syncthreads();
        // edges of the other components go away in the first compaction; from here on stamp[] is the candidate stamp
        int my_cnt = NE > tid ? (NE - tid + nt - 1) / nt : 0;   // live entries of this thread's edge column
        for (int i = 0; i < my_cnt;) {
            const int e = i * nt + tid;
            if (!(flags[eu[e]] & PF_ALIVE)) {
                const int last = (my_cnt - 1) * nt + tid;
                eu[e] = eu[last]; ev[e] = ev[last];
                --my_cnt;
                continue;
            }
            ++i;
        }
        __syncthreads();
        for (int i = tid; i < PEAC_BATCH * PEAC_MAXB; i += nt) stampK[i] = 0;
        if (tid < PEAC_BATCH) { s_ncand[tid] = 0; s_guard[tid] = 0; }
        int round_id = 0;
        // arg-min state of this thread over the nodes it owns (b = tid, tid + nt, ...): recomputed only after one of them changed
        unsigned long long my_key = ~0ull, my_key2 = ~0ull;   // order-preserving bit pattern of the mse (peac_key): best, second best
        int my_p = -1, my_s = 0x7fffffff, my_p2 = -1, my_s2 = 0x7fffffff;
        bool my_dirty = true;
        // lexicographic (mse bits, seq) minimum of the warp with three hardware reductions; seq is unique
        auto warp_argmin = [&](unsigned long long key, int sq, int node) -> int {
            const unsigned hi = node >= 0 ? (unsigned)(key >> 32) : 0xffffffffu, lo = (unsigned)key;
            const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
            bool c = node >= 0 && hi == mh;
            const unsigned ml = __reduce_min_sync(0xffffffffu, c ? lo : 0xffffffffu);
            c = c && lo == ml;
            const unsigned ms = __reduce_min_sync(0xffffffffu, c ? (unsigned)sq : 0xffffffffu);
            const unsigned who = __ballot_sync(0xffffffffu, c && (unsigned)sq == ms);
            return who ? __ffs(who) - 1 : -1;   // lane that holds the minimum, -1 = no live node in this warp
        };
        __syncthreads();
        PCLK(2);   // component setup
        // ---- rounds.  The reference pops ONE node at a time.  A round takes the PEAC_BATCH smallest queue entries p_1 <= p_2 <=
        // ... in queue order, evaluates the candidate merges of all of them on the state before the round, and commits the
        // longest prefix for which that is exactly what the serial loop would have done: p_j is committed iff (a) it and all
        // its graph neighbours are untouched by the commits of p_1 .. p_(j-1) (their merged pairs are different nodes, so p_j
        // sees the same candidates with the same statistics), and (b) no node created by those commits has a smaller MSE than
        // p_j (it would be popped first).  The first entry always commits.  Measured on the largest component of the synthetic
        // frames: 3.2 pops per round.
        for (;;) {
            ++round_id;
            const unsigned short rid = (unsigned short)round_id;
            if (tid < PEAC_BATCH) s_guard[tid] = 0;     // (next used after three barriers)
            // ---- flatten the union-find after the merges of the previous round; threads whose nodes changed rescan them
            {
                const int nm = s_nmerge;
                for (int q = 0; q < nm; ++q) {
                    const int lose = s_lose[q], win = s_win[q];
                    for (int b2 = tid; b2 < NB; b2 += nt)
                        if (root[b2] == lose) root[b2] = (unsigned short)win;
                }
            }
            if (my_dirty) {     // best and second best of the nodes this thread owns, one pass
                my_key = ~0ull; my_p = -1; my_s = 0x7fffffff;
                my_key2 = ~0ull; my_p2 = -1; my_s2 = 0x7fffffff;
                for (int b2 = tid; b2 < NB; b2 += nt)
                    if ((flags[b2] & (PF_ALIVE | PF_QUEUED)) == (PF_ALIVE | PF_QUEUED)) {
                        const unsigned long long kb = peac_key(mse_a[b2]);
                        const int sq = seq_a[b2];
                        if (kb < my_key || (kb == my_key && sq < my_s)) {
                            my_key2 = my_key; my_p2 = my_p; my_s2 = my_s;
                            my_key = kb; my_p = b2; my_s = sq;
                        } else if (kb < my_key2 || (kb == my_key2 && sq < my_s2)) { my_key2 = kb; my_p2 = b2; my_s2 = sq; }
                    }
                my_dirty = false;
            }
            // ---- the PEAC_WTOP smallest (mse, seq) of every warp: a lane that gave its best continues with its second best
            // (a lane that has to give a third rescans its nodes; rare) ...
            {
                unsigned long long k1 = my_key, k2 = my_key2;
                int p1 = my_p, s1 = my_s, p2 = my_p2, s2 = my_s2, given = 0, g0 = -1, g1 = -1;
#pragma unroll
                for (int r = 0; r < PEAC_WTOP; ++r) {
                    const int wl = warp_argmin(k1, s1, p1);
                    const int src = wl >= 0 ? wl : 0;
                    const unsigned long long k0 = __shfl_sync(0xffffffffu, k1, src);
                    const int p0 = __shfl_sync(0xffffffffu, p1, src), s0 = __shfl_sync(0xffffffffu, s1, src);
                    if (lane == 0) { w4_key[wid][r] = k0; w4_p[wid][r] = wl >= 0 ? p0 : -1; w4_seq[wid][r] = s0; }
                    if (r + 1 < PEAC_WTOP && wl >= 0 && lane == wl) {
                        if (given == 0) { g0 = p1; k1 = k2; p1 = p2; s1 = s2; given = 1; }
                        else {
                            g1 = p1; given = 2;
                            k1 = ~0ull; p1 = -1; s1 = 0x7fffffff;
                            for (int b2 = tid; b2 < NB; b2 += nt)
                                if (b2 != g0 && b2 != g1 && (flags[b2] & (PF_ALIVE | PF_QUEUED)) == (PF_ALIVE | PF_QUEUED)) {
                                    const unsigned long long kb = peac_key(mse_a[b2]);
                                    const int sq = seq_a[b2];
                                    if (kb < k1 || (kb == k1 && sq < s1)) { k1 = kb; p1 = b2; s1 = sq; }
                                }
                        }
                    }
                }
            }
            __syncthreads();
            // ---- ... and of the CTA: every warp merges the 8 x PEAC_WTOP warp entries (lane l holds entry l).  Beyond a warp's
            // last entry its next one is unknown, so the merged order is provably the queue order only up to and including the
            // first "last entry" taken: the batch ends there.
            int P[PEAC_BATCH], npk = 0;
            {
                const bool has = lane < (PEAC_AHC_NT / 32) * PEAC_WTOP;
                const int ew = lane / PEAC_WTOP, er = lane - ew * PEAC_WTOP;
                unsigned long long k1 = has ? w4_key[ew][er] : ~0ull;
                int p1 = has ? w4_p[ew][er] : -1, s1 = has ? w4_seq[ew][er] : 0x7fffffff;
                bool open = true;
#pragma unroll
                for (int r = 0; r < PEAC_BATCH; ++r) {
                    P[r] = -1;
                    const int wl = open ? warp_argmin(k1, s1, p1) : -1;
                    if (wl >= 0) {
                        P[r] = __shfl_sync(0xffffffffu, p1, wl);
                        if (lane == wl) p1 = -1;
                        npk = r + 1;
                        if (wl % PEAC_WTOP == PEAC_WTOP - 1) open = false;     // a warp's last entry: nothing after it is certain
                    } else open = false;
                }
            }
            if (tid < PEAC_BATCH) {       // (dynamic indexing of P[] would put it into local memory)
#pragma unroll
                for (int r = 0; r < PEAC_BATCH; ++r) if (tid == r) s_P[r] = P[r];
            }
            PCLK(3);   // flatten + selection
            if (npk == 0) break;
            // ---- distinct live graph neighbours of the batch nodes (AHCPlaneFitter.hpp:1092-1117); dead edges (internal to a
            // node, or touching an extracted node) are only skipped here and compacted out of the thread's edge column every
            // 16th round.  The stamp is written without an atomic: two threads that race on the same neighbour add it twice,
            // which costs one redundant evaluation and changes nothing.
            if ((round_id & 15) == 0) {
                for (int i = 0; i < my_cnt;) {
                    const int e = i * nt + tid;
                    const int ru = root[eu[e]], rv = root[ev[e]];
                    if (ru == rv || !(flags[ru] & PF_ALIVE) || !(flags[rv] & PF_ALIVE)) {   // edge is gone for good
                        const int last = (my_cnt - 1) * nt + tid;
                        eu[e] = eu[last]; ev[e] = ev[last];
                        --my_cnt;
                        continue;
                    }
                    ++i;
                }
            }
#pragma unroll 2
            for (int i = 0; i < my_cnt; ++i) {
                const int e = i * nt + tid;
                const int ru = root[eu[e]], rv = root[ev[e]];
                if (ru == rv) continue;
#pragma unroll
                for (int j = 0; j < PEAC_BATCH; ++j) {
                    const int o = ru == P[j] ? rv : (rv == P[j] ? ru : -1);
                    if (o < 0 || j >= npk || !(flags[o] & PF_ALIVE)) continue;
                    unsigned short *st = stampK + j * PEAC_MAXB + o;
                    if (*st == rid) continue;
                    *st = rid;
                    const int slot = atomicAdd(&s_ncand[j], 1);
                    if (slot < PEAC_CANDK) cand[j][slot] = (unsigned short)o;
                }
            }
            __syncthreads();
            PCLK(4);   // edge scan
            // ---- candidate merges: PEAC_AHC_NT / PEAC_BATCH threads per batch node; every thread evaluates its candidates
            // completely (smallest eigenvalue, and for its best one the plane normal), so that the winner can commit at once
            const int slot_j = tid / (PEAC_AHC_NT / PEAC_BATCH), slot_t = tid % (PEAC_AHC_NT / PEAC_BATCH);
            const int pj = s_P[slot_j];
            double best = 1e300;
            int best_o = -1;
            double bst[9], bc2 = 0.0, bn[3] = {0.0, 0.0, 1.0};
            if (slot_j < npk) {
                const int ncj = min(s_ncand[slot_j], PEAC_CANDK);
#ifdef PEAC_CLOCKS
                if (slot_t == 0 && slot_j == 0) { clk[8] += 1; clk[9] += ncj; clk[10] += npk; }
#endif
                for (int ci2 = slot_t; ci2 < ncj; ci2 += PEAC_AHC_NT / PEAC_BATCH) {
                    const int o = cand[slot_j][ci2];
                    if (sim(pj, o) < PEAC_SIM_MERGE) continue;
                    double st[9], c[3];
#pragma unroll
                    for (int k = 0; k < 9; ++k) st[k] = sst[pj * 9 + k] + sst[o * 9 + k];
                    const double sc = 1.0 / (double)(N_a[pj] + N_a[o]);
                    Sym3 K;
                    peac_cov(st, sc, c, K);
                    const double l = eig33_min_val(K);
                    const double mse = l * sc;
                    if (mse < best || (mse == best && o < best_o)) {
                        best = mse; best_o = o; bc2 = c[2];
#pragma unroll
                        for (int k = 0; k < 9; ++k) bst[k] = st[k];
                        double v[3];
                        eig33_vec(K, l, v);
                        const double d = v[0] * c[0] + v[1] * c[1] + v[2] * c[2];
                        const double sgn = d <= 0 ? 1.0 : -1.0;     // normal points towards the camera
                        bn[0] = sgn * v[0]; bn[1] = sgn * v[1]; bn[2] = sgn * v[2];
                    }
                }
            }
            {   // best of the warp (a batch node's threads are WPS whole warps): three integer warp reductions on the order-preserving key
                const int wl = warp_argmin(peac_key(best), best_o, best_o);      // ties in the MSE: the smaller node id
                if (wl >= 0 ? lane == wl : lane == 0) { ws_mse[slot_j][wid % WPS] = best; ws_o[slot_j][wid % WPS] = wl >= 0 ? best_o : -1; ws_c2[slot_j][wid % WPS] = bc2; }
            }
            __syncthreads();
            PCLK(5);   // candidate evaluation + warp reduction
            // ---- commit plan (warp 0; lane j = batch slot j, through shared memory): winners, valid prefix, queue sequence numbers,
            // extraction slots
            if (wid == 0) {
                if (lane < PEAC_BATCH) {
                    int ko = -1;
                    double km = 0.0, mp = 0.0;
                    bool mg = false, ext = false;
                    if (lane < npk) {
                        km = ws_mse[lane][0];
                        double kc = ws_c2[lane][0];
                        ko = ws_o[lane][0];
                        for (int q = 1; q < WPS; ++q) {      // best of the slot's warps: smallest MSE, ties to the smaller node id
                            const double om = ws_mse[lane][q];
                            const int oo = ws_o[lane][q];
                            if (oo >= 0 && (ko < 0 || om < km || (om == km && oo < ko))) { km = om; ko = oo; kc = ws_c2[lane][q]; }
                        }
                        mg = ko >= 0 && km < peac_t_mse(false, kc);
                        const int pme = s_P[lane];
                        mp = mse_a[pme];
                        ext = !mg && N_a[pme] >= PEAC_MIN_SUPPORT;
                    }
                    pl_O[lane] = ko; pl_M[lane] = km; pl_MG[lane] = mg ? 1 : 0; pl_mp[lane] = mp; pl_ext[lane] = ext ? 1 : 0;
                }
                __syncwarp();
                if (lane < PEAC_BATCH) {
                    // conflicts of slot `lane` with every earlier slot i (independent of the prefix): consumed / neighbour changed
                    bool clash = false;
                    if (lane < npk)
                        for (int i = 0; i < lane; ++i) {
                            const int pi = s_P[i], oi = pl_O[i];
                            const bool mi = pl_MG[i] != 0;
                            if (mi && oi == s_P[lane]) clash = true;
                            if (stampK[lane * PEAC_MAXB + pi] == rid) clash = true;
                            if (mi && stampK[lane * PEAC_MAXB + oi] == rid) clash = true;
                        }
                    pl_clash[lane] = clash ? 1 : 0;
                }
                __syncwarp();
                if (lane == 0) {
                    int L = 0, n_mrg = 0, n_ext = 0;
                    double minq = 1e300;
                    for (int j = 0; j < npk; ++j) {
                        if (pl_clash[j] || (j > 0 && minq < pl_mp[j])) break;
                        L = j + 1;
                        pl_seq[j] = n_mrg; pl_ex[j] = n_ext;
                        if (pl_MG[j]) { minq = fmin(minq, pl_M[j]); ++n_mrg; } else if (pl_ext[j]) ++n_ext;
                    }
                    pl_L = L; pl_nm = n_mrg; pl_ne = n_ext;
                }
            }
            __syncthreads();
            PCLK(11);  // commit plan
            const int L = pl_L, n_mrg = pl_nm, n_ext = pl_ne;
            const int my_O = pl_O[slot_j], my_seq_new = pl_seq[slot_j], my_ex_slot = pl_ex[slot_j];
            const double my_M = pl_M[slot_j];
            const bool my_MG = pl_MG[slot_j] != 0;
            const int seq_base = s_seq, nex_base = s_nex;
            __syncthreads();      // every thread has read the pre-commit state (mse_a, N_a, stamps, s_seq, s_nex)
            // ---- commit (parallel: the committed pairs are disjoint)
            if (slot_j < L) {
                const int j = slot_j, p = pj;
                if (my_MG) {
                    if (best_o == my_O && best == my_M) {   // one thread per distinct candidate; a duplicated candidate: the first one commits
                        if (atomicExch(&s_guard[j], 1) == 0) {
                            const int o = my_O;
                            // PlaneSeg(pa, pb): rid of the larger parent; DisjointSet::Union by size keeps the same root
                            const int wn = N_a[p] >= N_a[o] ? p : o, ls = wn == p ? o : p;
                            const int Nsum = N_a[p] + N_a[o];
#pragma unroll
                            for (int k = 0; k < 9; ++k) sst[wn * 9 + k] = bst[k];
                            nrm[3 * wn] = bn[0]; nrm[3 * wn + 1] = bn[1]; nrm[3 * wn + 2] = bn[2];
                            mse_a[wn] = best; N_a[wn] = Nsum;
                            seq_a[wn] = seq_base + my_seq_new;            // the merged node is a NEW queue entry
                            ssize[wn] += ssize[ls];
                            flags[wn] |= PF_ALIVE | PF_QUEUED;
                            flags[ls] &= ~(PF_ALIVE | PF_QUEUED);
                            s_lose[my_seq_new] = ls; s_win[my_seq_new] = wn;
                        }
                    }
                } else if (slot_t == 0) {   // extract p (or drop it) and cut it out of the graph
                    if (N_a[p] >= PEAC_MIN_SUPPORT) {
                        if (nex_base + my_ex_slot < PEAC_MAXP) { s_ex[nex_base + my_ex_slot] = p; s_exkey[nex_base + my_ex_slot] = mse_a[p]; }
                        else ctl->overflow = ctl->sticky_overflow = 1;
                    }
                    flags[p] &= ~(PF_ALIVE | PF_QUEUED);
                }
            }
            if (tid == 0) {
                s_seq = seq_base + n_mrg; s_nex = min(nex_base + n_ext, PEAC_MAXP); s_nmerge = n_mrg;
                for (int j = 0; j < PEAC_BATCH; ++j) { if (s_ncand[j] > PEAC_CANDK) ctl->overflow = ctl->sticky_overflow = 1; s_ncand[j] = 0; }
            }
            // nodes this thread owns that changed: the committed p_j and their partners
#pragma unroll
            for (int j = 0; j < PEAC_BATCH; ++j) {
                if (j < L && (P[j] & (PEAC_AHC_NT - 1)) == tid) my_dirty = true;
                if (j < L && pl_MG[j] && (pl_O[j] & (PEAC_AHC_NT - 1)) == tid) my_dirty = true;
            }
            __syncthreads();
            PCLK(6);   // merge / extract
        }
        // ---- results of this component: extracted planes (unsorted, global list) and the union-find of its blocks
        __syncthreads();
        if (tid < s_nex) {
            const int slot = atomicAdd(&ctl->n_ex, 1);
            if (slot < PEAC_MAXP) {
                const int r = s_ex[tid];
                PeacExtract &x = ctl->ex[slot];
                for (int d = 0; d < 9; ++d) x.st[d] = sst[r * 9 + d];
                for (int d = 0; d < 3; ++d) x.normal[d] = nrm[3 * r + d];
                x.mse = mse_a[r]; x.pop_key = s_exkey[tid]; x.N = N_a[r]; x.rid = r;
            } else ctl->overflow = ctl->sticky_overflow = 1;
        }
        for (int b = tid; b < NB; b += nt)
            if ((flags[b] & PF_VALID)) { ctl->parent[b] = root[b]; ctl->size[b] = ssize[root[b]]; }
        PCLK(7);   // outputs
#ifdef PEAC_CLOCKS
        if (tid == 0) for (int k = 0; k < 12; ++k) ctl->clk[blockIdx.x][k] = clk[k];
#endif
        if ((int)blockIdx.x + (round + 1) * (int)gridDim.x >= n_active) break;   // nothing left for this CTA: skip the rebuild
    }
    // ---- the last CTA to finish orders the planes: size descending, stable in the reference's extraction order
    __threadfence();
    __syncthreads();
    __shared__ int s_last;
    if (tid == 0) s_last = atomicAdd(&ctl->done, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid == 0) {
        const int n = min(*(volatile int *)&ctl->n_ex, PEAC_MAXP);
        int *idx = s_ex;      // (the kernel keeps no local-memory arrays: with 219 KB of shared memory carved out the L1 is tiny)
        for (int a = 0; a < n; ++a) idx[a] = a;
        volatile PeacExtract *ex = ctl->ex;
        auto before = [&](int a, int b) {   // a sorts before b
            if (ex[a].N != ex[b].N) return ex[a].N > ex[b].N;
            if (ex[a].pop_key != ex[b].pop_key) return ex[a].pop_key < ex[b].pop_key;
            return ex[a].rid < ex[b].rid;
        };
        for (int a = 1; a < n; ++a) {
            const int v = idx[a];
            int b = a - 1;
            while (b >= 0 && before(v, idx[b])) { idx[b + 1] = idx[b]; --b; }
            idx[b + 1] = v;
        }
        ctl->n_planes = n;
        for (int k = 0; k < n; ++k) {
            volatile PeacExtract &x = ex[idx[k]];
            PeacPlane &pl = ctl->pl[k];
            const double sc = 1.0 / (double)x.N;
            for (int d = 0; d < 3; ++d) { pl.center[d] = x.st[d] * sc; pl.normal[d] = x.normal[d]; }
            for (int d = 0; d < 9; ++d) pl.st[d] = x.st[d];
            pl.mse = x.mse; pl.thr = 9.0 * x.mse + 1e-5; pl.N = x.N; pl.rid = x.rid; pl.valid = 0; pl.final_id = -1;
            ctl->conn[k] = 0ull;
        }
    }
}

// ---------------------------------------------------------------- block erosion (findBlockMembership)
__global__ void k_peac_blockmap(PeacControl *ctl, int Nw, int Nh)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= Nw * Nh) return;
    auto find = [&](int x) { while (ctl->parent[x] != x) x = ctl->parent[x]; return x; };
    const int i = b / Nw, j = b - i * Nw;
    const int setid = find(b);
    int plid = -1;
    if (ctl->size[setid] * PEAC_WIN * PEAC_WIN >= PEAC_MIN_SUPPORT) {
        bool same = true;   // ERODE_ALL_BORDER: every 4-neighbour block must be in the same set
        if (j > 0) same &= find(b - 1) == setid;
        if (j < Nw - 1) same &= find(b + 1) == setid;
        if (i > 0) same &= find(b - Nw) == setid;
        if (i < Nh - 1) same &= find(b + Nw) == setid;
        if (same)
            for (int k = 0; k < ctl->n_planes; ++k)
                if (ctl->pl[k].rid == setid) { plid = k; break; }
    }
    ctl->blk_map[b] = plid;
    if (plid >= 0) ctl->pl[plid].valid = 1;
}

__global__ void k_peac_init_labels(const PeacControl *__restrict__ ctl, int W, int H, int Nw, int Nh, int *__restrict__ label, float *__restrict__ dist,
                                   int *__restrict__ head)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const int bx = x / PEAC_WIN, by = y / PEAC_WIN;
    label[y * W + x] = (bx < Nw && by < Nh) ? ctl->blk_map[by * Nw + bx] : -1;
    dist[y * W + x] = 3.402823466e+38f;
    head[y * W + x] = -1;
}

// ---------------------------------------------------------------- region growing (floodFill), order-faithful
// The reference grows all planes with ONE serial FIFO queue (AHCPlaneFitter.hpp:546-594): pop (pixel s, plane), visit the
// <= 4 neighbours of s in the order left / right / up / down, and per neighbour c update a small state machine
// (membershipImg[c], distMap[c]) that may append (c, plane) to the queue.  The outcome at a pixel depends on the ORDER in
// which the visits reach it (first come first served below 3 sigma, a closer plane may take a pixel over later, five
// failed visits close it for good), so it is reproduced exactly, level by level:
//   * a FIFO is processed in levels (the seeds are level 0, the entries appended while level n is processed form level
//     n + 1), and an entry never reads the state of its OWN pixel -- only visits to a pixel c touch c's state;
//   * phase A (parallel over the entries of the level): visit slot key = 4 * rank + direction; everything that does not
//     depend on the order (interior-block test, point validity, point-plane distance) is evaluated, the slot is linked
//     into the list of visits of pixel c;
//   * phase B (parallel over pixels): the visit that was linked first owns the pixel; it replays the pixel's visits in
//     ascending key order -- exactly the order of the serial queue -- through the reference's state machine and flags the
//     visits that append to the queue;
//   * phase C: the flagged visits, compacted in key order, are the next level -- the same sequence the serial queue holds.
// ~200 levels per frame, one 1024-thread CTA (the levels are a dependent chain; a level is 1-10 k entries).
#define PG_NT 512             // threads per CTA of the cluster kernel (PG_U visits each: 128 registers)
#define PG_U 4
__device__ __forceinline__ int pg_block_scan(int v, int *s_warp, int &total)
{
    // exclusive scan of v over the CTA (PG_NT threads); all threads must call
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    __syncthreads();
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
        int winc = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, winc, off);
            if (lane >= off) winc += t;
        }
        s_warp[lane] = winc - w;
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    total = s_warp[32];
    return inc - v + s_warp[wid];
}

// The first levels after the seeds are large (every boundary of every plane advances one pixel per level through the
// eroded border blocks: ~8 k entries, 30 k visits per level) and throughput bound on one SM, so they run on a cluster of
// PG_CL CTAs (phases separated by cluster barriers, everything in global memory, loads of data written by other CTAs
// bypass L1); as soon as a level fits the shared-memory path the frontier is handed to the single-CTA kernel below.
#define PG_CL 16                // CTAs of the cluster (non-portable size: opted in at init)
#define PG_SCAP 2048
#define PG_HAND 1024           // levels of at most PG_HAND entries are handed to the single-CTA kernel
#define PG_FNT 1024            // threads of the single-CTA kernel (levels of a few hundred entries: more warps only add barrier and scan overhead)
struct __align__(16) PgRec { int pix, info, next; float dist; };   // one visit: target pixel (-1: none), plane | ok << 8 | push << 9, list link, distance

__global__ void __launch_bounds__(PG_NT)
k_peac_grow_cluster(const uint16_t *__restrict__ depth, int W, int H, float fx, float fy, float cx, float cy, float inv_scale, int Nw, int Nh,
                    PeacControl *ctl, int *member, float *dist, int *head, int *qa, int *qb, PgRec *rec, float *curd)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ int s_blk[PEAC_MAXB];
    __shared__ double s_pn[PEAC_MAXP][3], s_pc[PEAC_MAXP][3], s_thr[PEAC_MAXP];
    __shared__ int s_warp[33];
    __shared__ int s_tot, s_base[2];
    __shared__ unsigned long long s_conn[PEAC_MAXP];    // plane adjacency found by this CTA (every boundary pixel would hit the same two global words)
    const int tid = threadIdx.x, NB = Nw * Nh, rank = (int)cluster.block_rank();
    const int g = rank * PG_NT + tid, G = PG_CL * PG_NT;
    const int np = ctl->n_planes;
    for (int b = tid; b < NB; b += PG_NT) s_blk[b] = ctl->blk_map[b];
    for (int k = tid; k < np; k += PG_NT) {
        for (int d = 0; d < 3; ++d) { s_pn[k][d] = ctl->pl[k].normal[d]; s_pc[k][d] = ctl->pl[k].center[d]; }
        s_thr[k] = ctl->pl[k].thr;
        s_conn[k] = 0ull;
    }
    __syncthreads();
    // ---- seeds in the order of findBlockMembership (AHCPlaneFitter.hpp:660-703): blocks in raster order, per block the run
    // along its top edge, then the run along its left edge (CTA 0 writes them; every CTA computes the count)
    int *cur = qa, *nxt = qb;
    int n = 0;
    {
        constexpr int SB = (PEAC_MAXB + PG_NT - 1) / PG_NT;      // consecutive blocks per thread
        int cnt[SB], offs[SB];
#pragma unroll
        for (int r = 0; r < SB; ++r) {
            const int b = tid * SB + r;
            cnt[r] = 0;
            if (b < NB) {
                const int i = b / Nw, j = b - i * Nw, m = s_blk[b];
                const bool top = i > 0 && (m < 0 ? s_blk[b - Nw] >= 0 : s_blk[b - Nw] != m);
                const bool left = j > 0 && (m < 0 ? s_blk[b - 1] >= 0 : s_blk[b - 1] != m);
                cnt[r] = (PEAC_WIN - 1) * ((top ? 1 : 0) + (left ? 1 : 0));
            }
        }
        int total;
        int mine = 0;
#pragma unroll
        for (int r = 0; r < SB; ++r) mine += cnt[r];
        const int base = pg_block_scan(mine, s_warp, total);
        offs[0] = base;
#pragma unroll
        for (int r = 1; r < SB; ++r) offs[r] = offs[r - 1] + cnt[r - 1];
        n = total;
        if (rank == 0 && n <= PG_CAP)
#pragma unroll
            for (int r = 0; r < SB; ++r) {
                const int b = tid * SB + r;
                if (b >= NB || !cnt[r]) continue;
                const int i = b / Nw, j = b - i * Nw, m = s_blk[b];
                int o = offs[r];
                if (m < 0) {
                    if (i > 0 && s_blk[b - Nw] >= 0) {
                        const int sp = (i * PEAC_WIN - 1) * W + j * PEAC_WIN, pl = s_blk[b - Nw];
                        for (int k = 1; k < PEAC_WIN; ++k) cur[o++] = (pl << 20) | (sp + k);
                    }
                    if (j > 0 && s_blk[b - 1] >= 0) {
                        const int sp = (i * PEAC_WIN) * W + j * PEAC_WIN - 1, pl = s_blk[b - 1];
                        for (int k = 0; k < PEAC_WIN - 1; ++k) cur[o++] = (pl << 20) | (sp + k * W);
                    }
                } else {
                    if (i > 0 && s_blk[b - Nw] != m) {
                        const int sp = (i * PEAC_WIN) * W + j * PEAC_WIN;
                        for (int k = 0; k < PEAC_WIN - 1; ++k) cur[o++] = (m << 20) | (sp + k);
                    }
                    if (j > 0 && s_blk[b - 1] != m) {
                        const int sp = (i * PEAC_WIN) * W + j * PEAC_WIN;
                        for (int k = 1; k < PEAC_WIN; ++k) cur[o++] = (m << 20) | (sp + k * W);
                    }
                }
            }
    }
    cluster.sync();
    int levels = 0, entries = 0, buf = 0;
#ifdef PEAC_CLOCKS
    long long gc[4] = {0, 0, 0, 0}, gt = clock64();
#define GCLK(i) do { if (g == 0) { const long long t_ = clock64(); gc[i] += t_ - gt; gt = t_; } } while (0)
#else
#define GCLK(i) do { } while (0)
#endif
    // Every phase is a chain of L2 round trips (~800 cycles each) and cluster barriers (~600 cycles, ~1500 after global stores), so
    // a thread works on PG_U visits at a time with the loads of all of them in flight together; a level of up to PG_U * G
    // visits (practically every level) is ONE batch, and then the thread's own visit records stay in registers between the
    // phases.
    while (n > PG_HAND && n <= PG_CAP) {
        ++levels; entries += n;
        const int nk = 4 * n;
        const bool single = nk <= PG_U * G;
        int oc[PG_U], oi[PG_U], ol[PG_U];      // own visit records of the (last) batch: pixel, info, link
        float od[PG_U], ocur[PG_U];             // point-plane distance, distMap of the target pixel
        // ---- phase A: one thread per visit (key = 4 * queue rank + direction).  The membership and the distance of the target
        // pixel only change in phase B, so they are fetched here, together with the depth and the list link: ONE L2 round trip
        for (int v0 = g; v0 < nk; v0 += PG_U * G) {
            int ent[PG_U];
            unsigned dv[PG_U];
            int tr[PG_U];
#pragma unroll
            for (int j = 0; j < PG_U; ++j) {
                const int v = v0 + j * G;
                ent[j] = v < nk ? __ldcg(&cur[v >> 2]) : -1;
            }
#pragma unroll
            for (int j = 0; j < PG_U; ++j) {
                const int v = v0 + j * G, q = v & 3;
                int c = -1;
                if (ent[j] >= 0) {
                    const int s = ent[j] & 0xfffff;
                    const int sy = s / W, sx = s - sy * W;
                    const int ccx = sx + (q == 0 ? -1 : (q == 1 ? 1 : 0)), ccy = sy + (q == 2 ? -1 : (q == 3 ? 1 : 0));
                    if (ccx >= 0 && ccx < W && ccy >= 0 && ccy < H) {
                        const int bx = ccx / PEAC_WIN, by = ccy / PEAC_WIN;
                        if (!(bx < Nw && by < Nh && s_blk[by * Nw + bx] >= 0)) c = ccy * W + ccx;
                    }
                }
                oc[j] = c; ol[j] = -1; dv[j] = 0; tr[j] = 0; ocur[j] = 0.0f;
                if (c >= 0) {
                    ol[j] = atomicExch(&head[c], v);
                    dv[j] = depth[c];
                    tr[j] = __ldcg(&member[c]);
                    ocur[j] = __ldcg(&dist[c]);
                }
            }
#pragma unroll
            for (int j = 0; j < PG_U; ++j) {
                const int v = v0 + j * G, c = oc[j];
                if (v >= nk) continue;
                if (c < 0) { rec[v].pix = -1; continue; }
                const int plid = ent[j] >> 20;
                const int ccy = c / W, ccx = c - ccy * W;
                const float df = (float)dv[j];
                const float z = df * inv_scale;
                const double P0 = ((float)ccx - cx) * z / fx, P1 = ((float)ccy - cy) * z / fy, P2 = z;
                const float cd = (float)fabs(s_pn[plid][0] * (P0 - s_pc[plid][0]) + s_pn[plid][1] * (P1 - s_pc[plid][1]) + s_pn[plid][2] * (P2 - s_pc[plid][2]));
                const bool ok = !(df < 1e-3f) && (double)cd * (double)cd < s_thr[plid];
                oi[j] = plid | (ok ? 256 : 0) | ((tr[j] + 6) << 16);
                od[j] = cd;
                int4 r;
                r.x = c; r.y = oi[j]; r.z = ol[j]; r.w = __float_as_int(cd);
                reinterpret_cast<int4 *>(rec)[v] = r;
                if (!single) curd[v] = ocur[j];
            }
        }
        cluster.sync();
        GCLK(0);
        // ---- phase B: the visit linked first owns the pixel.  It walks the pixel's list ONCE (a dependent chain of L2 loads, one
        // per further visit; the chains of the thread's PG_U visits advance together), sorts the <= 4 visits by key in
        // registers and replays them in queue order
        for (int v0 = g; v0 < nk; v0 += PG_U * G) {
            if (!single) {
#pragma unroll
                for (int j = 0; j < PG_U; ++j) {
                    const int v = v0 + j * G;
                    oc[j] = -1;
                    if (v < nk) {
                        const int4 t = __ldcg(reinterpret_cast<const int4 *>(rec) + v);
                        oc[j] = t.x; oi[j] = t.y; ol[j] = t.z; od[j] = __int_as_float(t.w);
                        ocur[j] = __ldcg(&curd[v]);
                    }
                }
            }
            int w[PG_U], h0[PG_U];
            // visits of the pixel: key << 10 | (plane | ok << 8), distance; slot 0 is filled from the list head
            int KI[PG_U][4];
            float DD[PG_U][4];
#pragma unroll
            for (int j = 0; j < PG_U; ++j) {
                const bool owner = oc[j] >= 0 && ol[j] == -1;
                h0[j] = owner ? __ldcg(&head[oc[j]]) : -1;
#pragma unroll
                for (int r = 0; r < 4; ++r) { KI[j][r] = 0x7fffffff; DD[j][r] = 0.0f; }
            }
#pragma unroll
            for (int j = 0; j < PG_U; ++j) {
                w[j] = h0[j];
                if (h0[j] >= 0) head[oc[j]] = -1;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                int4 t[PG_U];
#pragma unroll
                for (int j = 0; j < PG_U; ++j) {
                    const int v = v0 + j * G;
                    t[j].x = 0; t[j].y = 0; t[j].z = -1; t[j].w = 0;
                    if (w[j] >= 0) {
                        if (w[j] == v) { t[j].y = oi[j]; t[j].z = -1; t[j].w = __float_as_int(od[j]); }
                        else t[j] = __ldcg(reinterpret_cast<const int4 *>(rec) + w[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < PG_U; ++j)
                    if (w[j] >= 0) { KI[j][r] = (w[j] << 10) | (t[j].y & 0x1ff); DD[j][r] = __int_as_float(t[j].w); w[j] = t[j].z; }
            }
#pragma unroll
            for (int j = 0; j < PG_U; ++j) {
                if (h0[j] < 0) continue;
                const int c = oc[j];
                int trail = (oi[j] >> 16) - 6;
                float d = ocur[j];
                auto visit = [&](int k, int info, float cd) {       // the reference's per-visit state machine (AHCPlaneFitter.hpp:563-590)
                    const int plid = info & 255;
                    if (trail >= 0 && trail == plid) return;
                    if (info & 256) {
                        if (trail >= 0) {
                            const double sm = fabs(s_pn[plid][0] * s_pn[trail][0] + s_pn[plid][1] * s_pn[trail][1] + s_pn[plid][2] * s_pn[trail][2]);
                            if (sm >= PEAC_SIM_REFINE && !((s_conn[trail] >> plid) & 1ull)) { atomicOr(&s_conn[trail], 1ull << plid); atomicOr(&s_conn[plid], 1ull << trail); }
                        }
                        if (cd < d) { trail = plid; d = cd; rec[k].info = info | 512; }   // (phase C reads the plane and the flag only)
                        else if (trail < 0) trail -= 1;
                    } else if (trail < 0) trail -= 1;
                };
                if (w[j] < 0) {
#define PG_CSWAP(A, B)                                                                                          \
                    if (KI[j][B] < KI[j][A]) { const int tk = KI[j][A]; KI[j][A] = KI[j][B]; KI[j][B] = tk; const float td = DD[j][A]; DD[j][A] = DD[j][B]; DD[j][B] = td; }
                    PG_CSWAP(0, 1) PG_CSWAP(2, 3) PG_CSWAP(0, 2) PG_CSWAP(1, 3) PG_CSWAP(1, 2)
#undef PG_CSWAP
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (trail > -6 && KI[j][r] != 0x7fffffff) visit(KI[j][r] >> 10, KI[j][r] & 0x1ff, DD[j][r]);
                } else {      // more than four visits of one pixel in one level (a pixel queued twice): walk the list per visit
                    int last = -1;
                    for (;;) {
                        if (trail <= -6) break;
                        int k = 0x7fffffff;
                        for (int w2 = h0[j]; w2 >= 0; w2 = __ldcg(&rec[w2].next))
                            if (w2 > last && w2 < k) k = w2;
                        if (k == 0x7fffffff) break;
                        last = k;
                        visit(k, __ldcg(&rec[k].info) & 0x1ff, __ldcg(&rec[k].dist));
                    }
                }
                member[c] = trail;
                dist[c] = d;
            }
        }
        cluster.sync();
        GCLK(1);
        // ---- phase C: ordered compaction over the whole cluster (every thread a contiguous range of keys; totals through
        // distributed shared memory)
        {
            const int L = (nk + G - 1) / G;
            const int k0 = min(g * L, nk), k1 = min(k0 + L, nk);
            int e[PG_U];                                 // the appended entries of the first PG_U keys stay in registers
            int cnt = 0;
            if (L <= PG_U) {
                int4 t[PG_U];
#pragma unroll
                for (int j = 0; j < PG_U; ++j) {
                    t[j].x = -1; t[j].y = 0;
                    if (k0 + j < k1) t[j] = __ldcg(reinterpret_cast<const int4 *>(rec) + k0 + j);
                }
#pragma unroll
                for (int j = 0; j < PG_U; ++j) {
                    const bool f = t[j].x >= 0 && (t[j].y & 512);
                    e[j] = f ? (((t[j].y & 255) << 20) | t[j].x) : -1;
                    cnt += f ? 1 : 0;
                }
            } else
                for (int k = k0; k < k1; ++k) {
                    const int4 t = __ldcg(reinterpret_cast<const int4 *>(rec) + k);
                    cnt += (t.x >= 0 && (t.y & 512)) ? 1 : 0;
                }
            int total;
            int o = pg_block_scan(cnt, s_warp, total);
            if (tid == 0) s_tot = total;
            cluster.sync();
            if (tid < 32) {      // warp 0 reads the PG_CL totals
                const int t = tid < PG_CL ? *cluster.map_shared_rank(&s_tot, tid) : 0;
                const int al = __reduce_add_sync(0xffffffffu, t), bs = __reduce_add_sync(0xffffffffu, tid < rank ? t : 0);
                if (tid == 0) { s_base[0] = bs; s_base[1] = al; }
            }
            __syncthreads();
            const int all = s_base[1];
            o += s_base[0];
            if (all <= PG_CAP) {
                if (L <= PG_U) {
#pragma unroll
                    for (int j = 0; j < PG_U; ++j)
                        if (e[j] >= 0) nxt[o++] = e[j];
                } else
                    for (int k = k0; k < k1; ++k) {
                        const int4 t = __ldcg(reinterpret_cast<const int4 *>(rec) + k);
                        if (t.x >= 0 && (t.y & 512)) nxt[o++] = ((t.y & 255) << 20) | t.x;
                    }
            }
            n = all;
        }
        cluster.sync();
        GCLK(2);
        int *t = cur; cur = nxt; nxt = t;
        buf ^= 1;
    }
#ifdef PEAC_CLOCKS
    if (g == 0) { ctl->clk[60][0] = gc[0]; ctl->clk[60][1] = gc[1]; ctl->clk[60][2] = gc[2]; ctl->clk[60][3] = levels; }
#endif
    __syncthreads();
    for (int k = tid; k < np; k += PG_NT)
        if (s_conn[k]) atomicOr(&ctl->conn[k], s_conn[k]);
    if (g == 0) {
        ctl->grow_n = n; ctl->grow_buf = buf; ctl->grow_levels = levels; ctl->grow_entries = entries;
        if (n > PG_CAP) ctl->overflow = ctl->sticky_overflow = 1;
    }
}

// Levels of at most PG_SCAP entries (the common case: a frame has a few levels of 5-10 k entries right after the seeds and
// then ~170 levels of a few hundred) run entirely out of shared memory: the queue, the visit records AND the per-pixel
// visit lists.  A level costs ONE L2 round trip: phase A (one thread per visit) loads the target pixel's depth, membership
// and distance together -- the state of a pixel only changes in phase B, so the values phase B needs can be fetched here --
// and links the visit into the pixel's list through a shared-memory hash table keyed by the pixel (word = pixel << 13 |
// newest visit; compare-and-swap).  Phase B and C touch shared memory only (the new state is written back with plain
// stores).  Larger levels take the generic path over the global buffers (per-pixel list heads in head[]).
#define PG_TAB 8192            // hash slots (visits per level <= 4 * PG_SCAP = 8192 = 2^13, distinct pixels fewer)
#define PG_TAB_EMPTY 0xffffffffu
#define PG_NONE 0xffffu
#define PG_SMEM ((size_t)2 * PG_SCAP * 4 + (size_t)4 * PG_SCAP * (4 + 2 + 2 + 4 + 4 + 1) + (size_t)PG_TAB * 4)
__device__ __forceinline__ unsigned pg_hash(int pix) { return ((unsigned)pix * 2654435761u) >> 19; }   // 13 bits

__global__ void __launch_bounds__(PG_FNT) k_peac_grow_fifo(const uint16_t *__restrict__ depth, int W, int H, float fx, float fy, float cx, float cy,
                                                          float inv_scale, int Nw, int Nh, PeacControl *ctl, int *__restrict__ member,
                                                          float *__restrict__ dist, int *__restrict__ head, int *qa, int *qb,
                                                          int *g_pix, int *g_info, float *g_dist, int *g_next, unsigned char *g_push)
{
    extern __shared__ int pg_dyn[];
    int *s_q0 = pg_dyn, *s_q1 = s_q0 + PG_SCAP;
    int *s_pix = s_q1 + PG_SCAP;                                        // [4 PG_SCAP] target pixel of the visit, -1: none
    float *s_dist = (float *)(s_pix + 4 * PG_SCAP);                     // point-plane distance of the visit
    float *s_cur = s_dist + 4 * PG_SCAP;                                // distMap of the target pixel before this level
    unsigned *s_tab = (unsigned *)(s_cur + 4 * PG_SCAP);                // [PG_TAB] pixel << 13 | newest visit
    unsigned short *s_next = (unsigned short *)(s_tab + PG_TAB);        // list link (older visit), PG_NONE: first
    unsigned short *s_info = s_next + 4 * PG_SCAP;                      // plane | ok << 6 | (membership + 6) << 7
    unsigned char *s_push = (unsigned char *)(s_info + 4 * PG_SCAP);
    __shared__ int s_blk[PEAC_MAXB];
    __shared__ double s_pn[PEAC_MAXP][3], s_pc[PEAC_MAXP][3], s_thr[PEAC_MAXP];
    __shared__ unsigned long long s_conn[PEAC_MAXP];
    __shared__ int s_warp[33];
    const int tid = threadIdx.x, NB = Nw * Nh;
    const int np = ctl->n_planes;
    for (int b = tid; b < NB; b += PG_FNT) s_blk[b] = ctl->blk_map[b];
    for (int k = tid; k < np; k += PG_FNT) {
        for (int d = 0; d < 3; ++d) { s_pn[k][d] = ctl->pl[k].normal[d]; s_pc[k][d] = ctl->pl[k].center[d]; }
        s_thr[k] = ctl->pl[k].thr;
        s_conn[k] = 0ull;
    }
    for (int i = tid; i < PG_TAB; i += PG_FNT) s_tab[i] = PG_TAB_EMPTY;
    __syncthreads();
    // ---- the frontier left by k_peac_grow_cluster (seeds + the large levels)
    int n = ctl->grow_n;
    int *cur = ctl->grow_buf ? qb : qa, *nxt_g = ctl->grow_buf ? qa : qb;
    bool nxt_s_is_q1 = true;      // which shared queue buffer is free for the next level
    if (n > 0 && n <= PG_SCAP) {
        for (int e = tid; e < n; e += PG_FNT) s_q0[e] = cur[e];
        cur = s_q0;
    }
    __syncthreads();
    int levels = ctl->grow_levels, entries = ctl->grow_entries;
#ifdef PEAC_CLOCKS
    long long fc[6] = {0, 0, 0, 0, 0, 0}, ft = clock64();
    const int lv0 = levels;
    __shared__ int s_seg[8];
    long long sg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (tid < 8) s_seg[tid] = 0;
    __syncthreads();
#define SEGT(i, t0_) do { const int d_ = (int)(clock64() - (t0_)); if (d_ > s_seg[i]) atomicMax(&s_seg[i], d_); } while (0)
#define FCLK(i) do { if (tid == 0) { const long long t_ = clock64(); fc[i] += t_ - ft; ft = t_; } } while (0)
#else
#define FCLK(i) do { } while (0)
#define SEGT(i, t0_) do { } while (0)
#endif
    while (n > 0) {
        if (n > PG_CAP) { if (tid == 0) ctl->overflow = ctl->sticky_overflow = 1; break; }
        ++levels; entries += n;
        int *nxt;
#ifdef PEAC_CLOCKS
        const long long tw0 = clock64();
        const int ncls = n <= 64 ? 0 : (n <= 256 ? 1 : (n <= 1024 ? 2 : 3));
#endif
        if (n <= PG_SCAP) {
            // ---- phase A: one thread per visit (key = 4 * queue rank + direction)
            for (int v = tid; v < 4 * n; v += PG_FNT) {
                const int ent = cur[v >> 2], q = v & 3;
                const int s = ent & 0xfffff, plid = ent >> 20;
                const int sy = s / W, sx = s - sy * W;
                const int ccx = sx + (q == 0 ? -1 : (q == 1 ? 1 : 0)), ccy = sy + (q == 2 ? -1 : (q == 3 ? 1 : 0));
                int c = -1;
                if (ccx >= 0 && ccx < W && ccy >= 0 && ccy < H) {
                    const int bx = ccx / PEAC_WIN, by = ccy / PEAC_WIN;
                    if (!(bx < Nw && by < Nh && s_blk[by * Nw + bx] >= 0)) c = ccy * W + ccx;     // only pixels of "black" blocks grow (:567-568)
                }
                s_pix[v] = c;
                if (c < 0) continue;
                const unsigned dv = depth[c];                     // the three loads are in flight together
                const int tr = member[c];
                const float dcur = dist[c];
                unsigned link = PG_NONE;                          // link the visit into the list of pixel c
                for (unsigned h = pg_hash(c);; h = (h + 1) & (PG_TAB - 1)) {
                    unsigned w = s_tab[h];
                    bool done = false;
                    for (;;) {
                        if (w == PG_TAB_EMPTY) {
                            const unsigned old = atomicCAS(&s_tab[h], PG_TAB_EMPTY, ((unsigned)c << 13) | (unsigned)v);
                            if (old == PG_TAB_EMPTY) { done = true; break; }
                            w = old;
                        } else if ((int)(w >> 13) == c) {
                            const unsigned old = atomicCAS(&s_tab[h], w, ((unsigned)c << 13) | (unsigned)v);
                            if (old == w) { link = w & 8191u; done = true; break; }
                            w = old;
                        } else break;                             // another pixel's slot: probe on
                    }
                    if (done) break;
                }
                const float df = (float)dv;
                const float z = df * inv_scale;                                       // organised cloud point (DynaDetect.cc:562-587)
                const double P0 = ((float)ccx - cx) * z / fx, P1 = ((float)ccy - cy) * z / fy, P2 = z;
                const float cd = (float)fabs(s_pn[plid][0] * (P0 - s_pc[plid][0]) + s_pn[plid][1] * (P1 - s_pc[plid][1]) + s_pn[plid][2] * (P2 - s_pc[plid][2]));
                const bool ok = !(df < 1e-3f) && (double)cd * (double)cd < s_thr[plid];   // valid point within 3 sigma of the plane
                s_info[v] = (unsigned short)(plid | (ok ? 64 : 0) | ((tr + 6) << 7));
                s_dist[v] = cd;
                s_cur[v] = dcur;
                s_push[v] = 0;
                s_next[v] = (unsigned short)link;
            }
            __syncthreads();
            FCLK(0);
            // ---- phase B: the visit linked first owns the pixel and replays its visits in key order
            for (int v = tid; v < 4 * n; v += PG_FNT) {
                const int c = s_pix[v];
                if (c < 0 || s_next[v] != PG_NONE) continue;
                unsigned h = pg_hash(c);
                while ((int)(s_tab[h] >> 13) != c) h = (h + 1) & (PG_TAB - 1);
                const int h0 = (int)(s_tab[h] & 8191u);
                int trail = (int)(s_info[v] >> 7) - 6;
                float d = s_cur[v];
                int last = -1;
                for (;;) {
                    if (trail <= -6) break;                      // visited from 4 neighbours already (:563); nothing can change any more
                    int k = 0x7fffffff;                          // next visit in key order
                    for (int w = h0; w != (int)PG_NONE; w = s_next[w])
                        if (w > last && w < k) k = w;
                    if (k == 0x7fffffff) break;
                    last = k;
                    const int info = s_info[k], plid = info & 63;
                    if (trail >= 0 && trail == plid) continue;   // visited by the same plane (:564)
                    if (info & 64) {
                        if (trail >= 0) {                        // two planes meet: potential merge (:575-580)
                            const double sm = fabs(s_pn[plid][0] * s_pn[trail][0] + s_pn[plid][1] * s_pn[trail][1] + s_pn[plid][2] * s_pn[trail][2]);
                            if (sm >= PEAC_SIM_REFINE && !((s_conn[trail] >> plid) & 1ull)) { atomicOr(&s_conn[trail], 1ull << plid); atomicOr(&s_conn[plid], 1ull << trail); }
                        }
                        const float cd = s_dist[k];
                        if (cd < d) { trail = plid; d = cd; s_push[k] = 1; }
                        else if (trail < 0) trail -= 1;
                    } else if (trail < 0) trail -= 1;
                }
                member[c] = trail;
                dist[c] = d;
            }
            __syncthreads();
            FCLK(1);
            // ---- phase C: the appended entries, in key order, are the next level; the hash table is emptied
            if (4 * n <= PG_FNT) {      // one key per thread: ballot + one pass over the 32 warp counts (two barriers instead of four)
                const bool f = tid < 4 * n && s_pix[tid] >= 0 && s_push[tid];
                const unsigned m = __ballot_sync(0xffffffffu, f);
                const int lane = tid & 31, wid = tid >> 5;
                if (lane == 0) s_warp[wid] = __popc(m);
                for (int i = tid; i < PG_TAB; i += PG_FNT) s_tab[i] = PG_TAB_EMPTY;
                __syncthreads();
                const int wc = s_warp[lane];
                const int total = __reduce_add_sync(0xffffffffu, wc), base = __reduce_add_sync(0xffffffffu, lane < wid ? wc : 0);
                nxt = nxt_s_is_q1 ? s_q1 : s_q0;     // total <= 4 n <= PG_SCAP
                if (f) nxt[base + __popc(m & ((1u << lane) - 1))] = ((s_info[tid] & 63) << 20) | s_pix[tid];
                n = total;
                nxt_s_is_q1 = !nxt_s_is_q1;
            } else {
                const int total_keys = 4 * n;
                const int L = (total_keys + PG_FNT - 1) / PG_FNT;
                const int k0 = min(tid * L, total_keys), k1 = min(k0 + L, total_keys);
                int cnt = 0;
                for (int k = k0; k < k1; ++k) cnt += (s_pix[k] >= 0 && s_push[k]) ? 1 : 0;
                for (int i = tid; i < PG_TAB; i += PG_FNT) s_tab[i] = PG_TAB_EMPTY;
                int total;
                int o = pg_block_scan(cnt, s_warp, total);
                nxt = total <= PG_SCAP ? (nxt_s_is_q1 ? s_q1 : s_q0) : nxt_g;
                if (total <= PG_CAP)
                    for (int k = k0; k < k1; ++k)
                        if (s_pix[k] >= 0 && s_push[k]) nxt[o++] = ((s_info[k] & 63) << 20) | s_pix[k];
                n = total;
                if (total <= PG_SCAP) nxt_s_is_q1 = !nxt_s_is_q1;
                else nxt_g = nxt_g == qa ? qb : qa;
            }
            __syncthreads();
            FCLK(2);
#ifdef PEAC_CLOCKS
            if (tid == 0) { sg[ncls] += clock64() - tw0; sg[4 + ncls] += 1; }
#endif
            cur = nxt;
            continue;
        }
        // ---- generic path (a level of more than PG_SCAP entries after the cluster kernel handed over: rare)
        // phase A
        for (int e = tid; e < n; e += PG_FNT) {
            const int ent = cur[e];
            const int s = ent & 0xfffff, plid = ent >> 20;
            const int sy = s / W, sx = s - sy * W;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool ex = q == 0 ? sx > 0 : (q == 1 ? sx < W - 1 : (q == 2 ? sy > 0 : sy < H - 1));
                const int key = e * 4 + q;
                int c = -1;
                const int ccx = sx + (q == 0 ? -1 : (q == 1 ? 1 : 0)), ccy = sy + (q == 2 ? -1 : (q == 3 ? 1 : 0));
                if (ex) {
                    const int bx = ccx / PEAC_WIN, by = ccy / PEAC_WIN;
                    if (!(bx < Nw && by < Nh && s_blk[by * Nw + bx] >= 0)) c = ccy * W + ccx;
                }
                g_pix[key] = c;
                if (c < 0) continue;
                const int link = atomicExch(&head[c], key);
                const float df = (float)depth[c];
                const float z = df * inv_scale;
                const double P0 = ((float)ccx - cx) * z / fx, P1 = ((float)ccy - cy) * z / fy, P2 = z;
                const float cd = (float)fabs(s_pn[plid][0] * (P0 - s_pc[plid][0]) + s_pn[plid][1] * (P1 - s_pc[plid][1]) + s_pn[plid][2] * (P2 - s_pc[plid][2]));
                const bool has = !(df < 1e-3f);
                g_info[key] = plid | ((has && (double)cd * (double)cd < s_thr[plid]) ? 256 : 0);
                g_dist[key] = has ? cd : -1.0f;
                g_push[key] = 0;
                g_next[key] = link;
            }
        }
        __syncthreads();
        // phase B
        for (int key = tid; key < 4 * n; key += PG_FNT) {
            const int c = g_pix[key];
            if (c < 0 || g_next[key] != -1) continue;            // the visit linked first (list tail) owns the pixel
            int trail = member[c];
            float d = dist[c];
            const int h0 = head[c];
            head[c] = -1;
            int last = -1;
            for (;;) {
                if (trail <= -6) break;
                int k = 0x7fffffff;
                for (int w = h0; w >= 0; w = g_next[w])
                    if (w > last && w < k) k = w;
                if (k == 0x7fffffff) break;
                last = k;
                const int info = g_info[k], plid = info & 255;
                if (trail >= 0 && trail == plid) continue;
                if (info & 256) {
                    if (trail >= 0) {
                        const double sm = fabs(s_pn[plid][0] * s_pn[trail][0] + s_pn[plid][1] * s_pn[trail][1] + s_pn[plid][2] * s_pn[trail][2]);
                        if (sm >= PEAC_SIM_REFINE && !((s_conn[trail] >> plid) & 1ull)) { atomicOr(&s_conn[trail], 1ull << plid); atomicOr(&s_conn[plid], 1ull << trail); }
                    }
                    const float cd = g_dist[k];
                    if (cd < d) { trail = plid; d = cd; g_push[k] = 1; }
                    else if (trail < 0) trail -= 1;
                } else if (trail < 0) trail -= 1;
            }
            member[c] = trail;
            dist[c] = d;
        }
        __syncthreads();
        // phase C
        {
            const int total_keys = 4 * n;
            const int L = (total_keys + PG_FNT - 1) / PG_FNT;
            const int k0 = min(tid * L, total_keys), k1 = min(k0 + L, total_keys);
            int cnt = 0;
            for (int k = k0; k < k1; ++k) cnt += (g_pix[k] >= 0 && g_push[k]) ? 1 : 0;
            int total;
            int o = pg_block_scan(cnt, s_warp, total);
            nxt = total <= PG_SCAP ? (nxt_s_is_q1 ? s_q1 : s_q0) : nxt_g;
            if (total <= PG_CAP)
                for (int k = k0; k < k1; ++k)
                    if (g_pix[k] >= 0 && g_push[k]) nxt[o++] = ((g_info[k] & 255) << 20) | g_pix[k];
            n = total;
            if (total <= PG_SCAP) nxt_s_is_q1 = !nxt_s_is_q1;
            else nxt_g = nxt_g == qa ? qb : qa;
        }
        __syncthreads();
        cur = nxt;
    }
#ifdef PEAC_CLOCKS
    if (tid == 0) { ctl->clk[61][0] = fc[0]; ctl->clk[61][1] = fc[1]; ctl->clk[61][2] = fc[2]; ctl->clk[61][3] = levels - lv0; for (int q = 0; q < 8; ++q) ctl->clk[61][4 + q] = sg[q]; }
#endif
    __syncthreads();
    for (int k = tid; k < np; k += PG_FNT) ctl->conn[k] |= s_conn[k];
    if (tid == 0) { ctl->grow_levels = levels; ctl->grow_entries = entries; }
}

// ---------------------------------------------------------------- final re-merge (second ahCluster) + relabel map
__global__ void k_peac_merge(PeacControl *ctl)
{
    if (threadIdx.x || blockIdx.x) return;
    const int n0 = ctl->n_planes;
    // nodes: 0..n0-1 = extracted planes, merged nodes appended
    const int MAXN = 2 * PEAC_MAXP;
    double st[MAXN][9], nrm[MAXN][3], mse[MAXN], cz[MAXN];
    int N[MAXN], rid[MAXN], seq[MAXN];
    bool alive[MAXN], inq[MAXN];
    unsigned long long adj[MAXN][2];
    int parent[PEAC_MAXP];     // union-find over the OLD plane ids (stands for ds->Union on their root block ids)
    int psize[PEAC_MAXP];
    int nn = n0;
    for (int k = 0; k < n0; ++k) {
        const PeacPlane &p = ctl->pl[k];
        for (int d = 0; d < 9; ++d) st[k][d] = p.st[d];
        for (int d = 0; d < 3; ++d) nrm[k][d] = p.normal[d];
        mse[k] = p.mse; cz[k] = p.center[2]; N[k] = p.N; rid[k] = k; seq[k] = k;
        alive[k] = p.valid != 0; inq[k] = alive[k];
        adj[k][0] = p.valid ? ctl->conn[k] : 0ull; adj[k][1] = 0ull;
        parent[k] = k; psize[k] = p.N / (PEAC_WIN * PEAC_WIN);   // DisjointSet sizes count blocks
    }
    for (int k = 0; k < n0; ++k)      // connections only between valid planes
        for (int j = 0; j < n0; ++j)
            if (!(ctl->pl[j].valid) || !(ctl->pl[k].valid)) adj[k][0] &= ~(1ull << j);
    auto has = [&](int a, int b) { return (adj[a][b >> 6] >> (b & 63)) & 1ull; };
    auto setb = [&](int a, int b) { adj[a][b >> 6] |= 1ull << (b & 63); };
    auto clrb = [&](int a, int b) { adj[a][b >> 6] &= ~(1ull << (b & 63)); };
    auto find = [&](int x) { while (parent[x] != x) x = parent[x]; return x; };
    int n_final_nodes = 0;
    for (;;) {
        int p = -1;   // min-MSE live queue entry
        for (int k = 0; k < nn; ++k)
            if (inq[k] && alive[k] && (p < 0 || mse[k] < mse[p] || (mse[k] == mse[p] && seq[k] < seq[p]))) p = k;
        if (p < 0) break;
        inq[p] = false;
        int best = -1;
        double bst[9], bc[3], bn[3], bm = 0;
        for (int o = 0; o < nn; ++o) {
            if (!has(p, o) || !alive[o]) continue;
            const double s = fabs(nrm[p][0] * nrm[o][0] + nrm[p][1] * nrm[o][1] + nrm[p][2] * nrm[o][2]);
            if (s < PEAC_SIM_MERGE) continue;
            double s9[9], c[3], n[3], m;
            for (int d = 0; d < 9; ++d) s9[d] = st[p][d] + st[o][d];
            peac_compute(s9, N[p] + N[o], c, n, m);
            if (best < 0 || bm > m) { best = o; bm = m; for (int d = 0; d < 9; ++d) bst[d] = s9[d]; for (int d = 0; d < 3; ++d) { bc[d] = c[d]; bn[d] = n[d]; } }
        }
        if (best >= 0 && bm < peac_t_mse(false, bc[2]) && nn < MAXN) {
            const int o = best, q = nn++;
            for (int d = 0; d < 9; ++d) st[q][d] = bst[d];
            for (int d = 0; d < 3; ++d) nrm[q][d] = bn[d];
            mse[q] = bm; cz[q] = bc[2]; N[q] = N[p] + N[o]; seq[q] = q;
            rid[q] = N[p] >= N[o] ? rid[p] : rid[o];
            // ds->Union(pa.rid, pb.rid) by set size (block counts are proportional to N for INIT_STRICT blocks)
            {
                const int a = find(rid[p]), b = find(rid[o]);
                if (a != b) { if (psize[a] < psize[b]) { parent[a] = b; psize[b] += psize[a]; } else { parent[b] = a; psize[a] += psize[b]; } }
            }
            adj[q][0] = (adj[p][0] | adj[o][0]); adj[q][1] = (adj[p][1] | adj[o][1]);
            clrb(q, p); clrb(q, o);
            for (int k = 0; k < nn; ++k) {
                if (has(k, p) || has(k, o)) { clrb(k, p); clrb(k, o); if (k != q) setb(k, q); }
            }
            adj[p][0] = adj[p][1] = adj[o][0] = adj[o][1] = 0ull;
            alive[p] = alive[o] = false;
            alive[q] = true; inq[q] = true;
        } else {
            if (N[p] >= PEAC_MIN_SUPPORT) ++n_final_nodes;
            for (int k = 0; k < nn; ++k) clrb(k, p);
            adj[p][0] = adj[p][1] = 0ull;
            alive[p] = false;
        }
    }
    // plidmap (AHCPlaneFitter.hpp:304-323); psize here counts planes, the reference's set sizes count blocks: use N
    // the root of a merged set in the reference is the root block id of the side with more blocks at every union
    int nf = 0;
    int plidmap[PEAC_MAXP];
    for (int k = 0; k < n0; ++k) plidmap[k] = -1;
    for (int i = 0; i < n0; ++i) {
        if (!ctl->pl[i].valid) continue;
        const int root = find(i);
        if (root == i) { if (plidmap[i] < 0) plidmap[i] = nf++; }
        else if (plidmap[root] < 0) { plidmap[i] = plidmap[root] = nf++; }
        else plidmap[i] = plidmap[root];
    }
    for (int k = 0; k < n0; ++k) ctl->pl[k].final_id = plidmap[k];
    ctl->n_final = nf;
    (void)n_final_nodes; (void)cz;
}

__global__ void k_peac_bits(const int *__restrict__ label, int n, const PeacControl *__restrict__ ctl, ulonglong2 *__restrict__ PB)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = label[i];
    ulonglong2 b = make_ulonglong2(0ull, 0ull);
    if (l >= 0) {
        const int f = ctl->pl[l].final_id;
        if (f >= 0 && f < 64) b.x = 1ull << f;
    }
    PB[i] = b;
}

// ---------------------------------------------------------------- host side
int peac_init(sindyn_base *ctx, PeacStage *p, int W, int H)
{
    p->W = W; p->H = H;
    PeacImpl *im = new PeacImpl();
    p->impl = im;
    im->Nw = W / PEAC_WIN; im->Nh = H / PEAC_WIN;
    if (im->Nw * im->Nh > PEAC_MAXB || im->Nw > 128 || im->Nh > 64) { ctx->err = "peac: image too large for the block tables"; return SINDYN_ERR_INVALID; }
    SD_CHECK(ctx->dalloc(&im->nodes, PEAC_MAXB));
    SD_CHECK(ctx->dalloc(&im->ctl, 1));
    CU_CHECK(ctx, cudaMemsetAsync(im->ctl, 0, sizeof(PeacControl), ctx->stream));
    SD_CHECK(ctx->dalloc(&im->label, (size_t)W * H));
    SD_CHECK(ctx->dalloc(&im->dist, (size_t)W * H));
    SD_CHECK(ctx->dalloc(&im->head, (size_t)W * H));
    SD_CHECK(ctx->dalloc(&im->qa, (size_t)PG_CAP));
    SD_CHECK(ctx->dalloc(&im->qb, (size_t)PG_CAP));
    SD_CHECK(ctx->dalloc(&im->v_pix, (size_t)4 * PG_CAP));
    SD_CHECK(ctx->dalloc(&im->v_info, (size_t)4 * PG_CAP));
    SD_CHECK(ctx->dalloc(&im->v_next, (size_t)4 * PG_CAP));
    SD_CHECK(ctx->dalloc(&im->v_dist, (size_t)4 * PG_CAP));
    SD_CHECK(ctx->dalloc(&im->v_push, (size_t)4 * PG_CAP));
    SD_CHECK(ctx->dalloc(&im->PB, (size_t)W * H));
    SD_CHECK(ctx->dalloc(&im->rec, (size_t)4 * PG_CAP));
    CU_CHECK(ctx, cudaFuncSetAttribute(k_peac_grow_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    if ((size_t)W * H >= (1u << 20)) { ctx->err = "peac: image too large for the packed queue entries"; return SINDYN_ERR_INVALID; }
    const size_t smem = PEAC_AHC_SMEM;
    CU_CHECK(ctx, cudaFuncSetAttribute(k_peac_ahc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU_CHECK(ctx, cudaFuncSetAttribute(k_peac_grow_fifo, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PG_SMEM));
    p->built = true;
    return SINDYN_OK;
}

int peac_run(sindyn_base *ctx, PeacStage *p, ReclusterStage *rc, const uint16_t *depth, float fx, float fy, float cx, float cy, float depth_scale,
             uint8_t *plane_edges_out)
{
    PeacImpl *im = (PeacImpl *)p->impl;
    const int W = p->W, H = p->H, Nw = im->Nw, Nh = im->Nh, NB = Nw * Nh;
    const float inv_scale = 1.0f / depth_scale;
    const size_t smem = PEAC_AHC_SMEM;
    CU_CHECK(ctx, cudaMemsetAsync(im->ctl, 0, 16 * sizeof(int), ctx->stream));
    LAUNCH(ctx, k_peac_blocks, cdiv(NB, 64), 64, 0, depth, W, H, fx, fy, cx, cy, inv_scale, Nw, Nh, im->nodes);
    LAUNCH(ctx, k_peac_ahc, PEAC_AHC_CTAS, PEAC_AHC_NT, smem, im->nodes, im->ctl, Nw, Nh);
    LAUNCH(ctx, k_peac_blockmap, cdiv(NB, 128), 128, 0, im->ctl, Nw, Nh);
    const dim3 blk(32, 8), grd(cdiv(W, 32), cdiv(H, 8));
    LAUNCH(ctx, k_peac_init_labels, grd, blk, 0, im->ctl, W, H, Nw, Nh, im->label, im->dist, im->head);
    {   // one cluster of PG_CL CTAs
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(PG_CL); cfg.blockDim = dim3(PG_NT); cfg.dynamicSmemBytes = 0; cfg.stream = ctx->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = PG_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CU_CHECK(ctx, cudaLaunchKernelEx(&cfg, k_peac_grow_cluster, depth, W, H, fx, fy, cx, cy, inv_scale, Nw, Nh, im->ctl, im->label, im->dist, im->head, im->qa,
                                         im->qb, im->rec, im->v_dist));
        ctx->launches++;
    }
    LAUNCH(ctx, k_peac_grow_fifo, 1, PG_FNT, PG_SMEM, depth, W, H, fx, fy, cx, cy, inv_scale, Nw, Nh, im->ctl, im->label, im->dist, im->head, im->qa, im->qb,
           im->v_pix, im->v_info, im->v_dist, im->v_next, im->v_push);
    LAUNCH(ctx, k_peac_merge, 1, 32, 0, im->ctl);
    LAUNCH(ctx, k_peac_bits, cdiv(W * H, 256), 256, 0, im->label, W * H, im->ctl, im->PB);
    LAUNCH_CHECK(ctx);
    return plane_contours_run(ctx, rc, im->PB, &im->ctl->n_final, plane_edges_out);
}

int peac_copy_header(sindyn_base *ctx, PeacStage *p, int *host4)
{
    PeacImpl *im = (PeacImpl *)p->impl;
    CU_CHECK(ctx, cudaMemcpyAsync(host4, im->ctl, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    return SINDYN_OK;
}

int peac_copy_sticky_overflow(sindyn_base *ctx, PeacStage *p, int *host1)
{
    PeacImpl *im = (PeacImpl *)p->impl;
    CU_CHECK(ctx, cudaMemcpyAsync(host1, &im->ctl->sticky_overflow, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    return SINDYN_OK;
}

int peac_get_debug(sindyn_base *ctx, PeacStage *p, int *label_out, int *planes_rid_n, int *n_planes, int *n_final)
{
    PeacImpl *im = (PeacImpl *)p->impl;
    static PeacControl host;   // large struct: keep it off the stack
    CU_CHECK(ctx, cudaMemcpyAsync(&host, im->ctl, sizeof host, cudaMemcpyDeviceToHost, ctx->stream));
    if (label_out) CU_CHECK(ctx, cudaMemcpyAsync(label_out, im->label, sizeof(int) * p->W * p->H, cudaMemcpyDeviceToHost, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
#ifdef PEAC_CLOCKS
    for (int c = 0; c < PEAC_AHC_CTAS; ++c)
        if (host.clk[c][8]) {
            fprintf(stderr, "peac ahc cta %2d: rounds %lld cand(first)/round %.1f batch/round %.2f | kcycles: load %lld comps %lld setup %lld select %lld scan %lld cand %lld plan %lld commit %lld out %lld\n", c,
                    host.clk[c][8], (double)host.clk[c][9] / host.clk[c][8], (double)host.clk[c][10] / host.clk[c][8], host.clk[c][0] / 1000, host.clk[c][1] / 1000,
                    host.clk[c][2] / 1000, host.clk[c][3] / 1000, host.clk[c][4] / 1000, host.clk[c][5] / 1000, host.clk[c][11] / 1000, host.clk[c][6] / 1000, host.clk[c][7] / 1000);
        }
    fprintf(stderr, "peac grow: levels %d entries %d | cluster kernel: %lld levels, kcycles A %lld B %lld C %lld | single-CTA kernel: %lld levels, kcycles A %lld B %lld C %lld (levels of <= 64 / 256 / 1024 / 2048 entries: %lld / %lld / %lld / %lld levels, %lld / %lld / %lld / %lld kcycles)\n",
            host.grow_levels, host.grow_entries, host.clk[60][3], host.clk[60][0] / 1000, host.clk[60][1] / 1000, host.clk[60][2] / 1000, host.clk[61][3],
            host.clk[61][0] / 1000, host.clk[61][1] / 1000, host.clk[61][2] / 1000, host.clk[61][8], host.clk[61][9], host.clk[61][10], host.clk[61][11], host.clk[61][4] / 1000, host.clk[61][5] / 1000, host.clk[61][6] / 1000, host.clk[61][7] / 1000);
#endif
    if (n_planes) *n_planes = host.n_planes;
    if (n_final) *n_final = host.n_final;
    if (planes_rid_n)
        for (int k = 0; k < host.n_planes && k < PEAC_MAXP; ++k) { planes_rid_n[3 * k] = host.pl[k].rid; planes_rid_n[3 * k + 1] = host.pl[k].N; planes_rid_n[3 * k + 2] = host.pl[k].final_id; }
    if (label_out)   // map grown labels to final plane ids like the oracle's membership image
        for (int i = 0; i < p->W * p->H; ++i) label_out[i] = label_out[i] >= 0 ? host.pl[label_out[i]].final_id : -1;
    return host.overflow ? SINDYN_ERR_CAPACITY : SINDYN_OK;
}
