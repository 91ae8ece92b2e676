// recluster.cu -- DynaDetect::SegAndMergeV2 (ORB_SLAM2/src/DynaDetect.cc:653-1018) and the plane-edge filter of
// CalOccluded (DynaDetect.cc:598-641) on per-pixel 128-bit membership sets.  See recluster.cuh / ccl.cuh.
#include "recluster.cuh"

#include "morph.cuh"

#define KM_FIX_SCALE 68719476736.0 /* 2^36, same fixed point as the k-means centre sums */

typedef ulonglong2 B128;
__device__ __forceinline__ B128 b0() { return make_ulonglong2(0ull, 0ull); }
__device__ __forceinline__ B128 bor(B128 a, B128 b) { return make_ulonglong2(a.x | b.x, a.y | b.y); }
__device__ __forceinline__ B128 band(B128 a, B128 b) { return make_ulonglong2(a.x & b.x, a.y & b.y); }
__device__ __forceinline__ B128 bnot(B128 a) { return make_ulonglong2(~a.x, ~a.y); }
__device__ __forceinline__ bool bany(B128 a) { return (a.x | a.y) != 0ull; }
__device__ __forceinline__ B128 bbit(int c) { return c < 64 ? make_ulonglong2(1ull << c, 0ull) : make_ulonglong2(0ull, 1ull << (c - 64)); }
__device__ __forceinline__ bool btest(B128 a, int c) { return c < 64 ? (a.x >> c) & 1ull : (a.y >> (c - 64)) & 1ull; }
// iterate set bits: returns the lowest set bit index and clears it
__device__ __forceinline__ int bpop(B128 &a)
{
    if (a.x) { int c = __ffsll((long long)a.x) - 1; a.x &= a.x - 1; return c; }
    int c = __ffsll((long long)a.y) - 1; a.y &= a.y - 1; return c + 64;
}

// ------------------------------------------------------------------ split (DynaDetect.cc:664-676)
__global__ void k_rc_reset(ReclusterControl *ctl, const ClusterOrder *ord)
{
    int t = threadIdx.x;
    if (t == 0) {
        int np = ord->n_kept - 1;
        ctl->n_planes = np < 0 ? 0 : (np > RC_MAXP ? RC_MAXP : np);
        ctl->n_raw = 0; ctl->n_comp = 0; ctl->overflow = 0; ctl->n_labels = 0; ctl->depth_max = 0;
    }
    if (t < RC_MAXC) { ctl->area[t] = 0; ctl->zsum[t] = 0; ctl->cnt1[t] = 0; ctl->lj_area[t] = 0; }
    if (t <= RC_MAXP) ctl->plane_bits[t] = make_ulonglong2(0ull, 0ull);
}

__global__ void k_rc_prepare(const uint8_t *__restrict__ labels_km, const uint8_t *__restrict__ occl1, const uint16_t *__restrict__ depth,
                             int n, const ClusterOrder *__restrict__ ord, ReclusterControl *ctl, int8_t *__restrict__ plane_of,
                             uint16_t *__restrict__ mcl)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int d = 0;
    if (i < n) {
        int l = labels_km[i];
        int pl = l < KM_K ? ord->rank_of[l] : -1;
        if (pl >= ctl->n_planes) pl = -1;   // the last kept cluster (farthest / invalid depth) is not processed
        plane_of[i] = (int8_t)pl;
        mcl[i] = (pl >= 0 && occl1[i] < 255) ? (uint16_t)(1u << pl) : (uint16_t)0;   // eachLabel - imgOccluded (saturating)
        d = depth[i];
    }
    for (int o = 16; o > 0; o >>= 1) d = max(d, __shfl_xor_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0 && d) atomicMax(&ctl->depth_max, d);
}

__global__ void k_rc_expand16(const uint16_t *__restrict__ m, int n, const int *__restrict__ active, uint8_t *__restrict__ cls)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int np = *active;
    uint16_t v = m[i];
    for (int pl = 0; pl < np; ++pl) cls[(size_t)pl * n + i] = (v >> pl) & 1;
}

// contours[c].size() > 50 && contourArea > 80 (DynaDetect.cc:678)
__global__ void k_rc_qualify(const int *__restrict__ top, RegionStats *__restrict__ stats, int n, ReclusterControl *ctl)
{
    const int pl = blockIdx.z;
    if (pl >= ctl->n_planes) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (top[(size_t)pl * n + i] != i) return;
    RegionStats *s = stats + (size_t)pl * n + i;
    const unsigned long long st = s->steps;
    const long long steps = (long long)(st & 0xffffffffull) + (long long)(st >> 32);
    const long long a2 = s->area2 < 0 ? -s->area2 : s->area2;
    bool ok = steps > 50 && a2 > 160;
    s->area2 = 0;   // becomes (component index + 1) for qualified roots in k_rc_sort
    if (ok) {
        int slot = atomicAdd(&ctl->n_raw, 1);
        if (slot < RC_MAXC) ctl->raw_key[slot] = ((unsigned long long)pl << 32) | (unsigned int)i;
        else ctl->overflow = 1;
    }
}

// deterministic component numbering: by (k-means plane in depth order, raster-first pixel)
__global__ void k_rc_sort(ReclusterControl *ctl, RegionStats *stats, int n)
{
    __shared__ unsigned long long key[RC_MAXC];
    const int t = threadIdx.x;
    const int nc = min(ctl->n_raw, RC_MAXC);
    if (t < nc) key[t] = ctl->raw_key[t];
    __syncthreads();
    if (t < nc) {
        int rank = 0;
        for (int j = 0; j < nc; ++j) rank += key[j] < key[t];
        const int pl = (int)(key[t] >> 32), root = (int)(key[t] & 0xffffffffull);
        ctl->comp_plane[rank] = pl;
        ctl->comp_root[rank] = root;
        stats[(size_t)pl * n + root].area2 = rank + 1;
        if (rank < 64) atomicOr(&ctl->plane_bits[pl].x, 1ull << rank);
        else atomicOr(&ctl->plane_bits[pl].y, 1ull << (rank - 64));
    }
    if (t == 0) ctl->n_comp = nc;
}

// F(c) = drawContours(FILLED) of component c: the top-level region and everything nested in it (DynaDetect.cc:682)
__global__ void k_rc_fill_bits(const int *__restrict__ top, const RegionStats *__restrict__ stats, int n, const ReclusterControl *__restrict__ ctl,
                               B128 *__restrict__ F)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int np = ctl->n_planes;
    B128 b = b0();
    for (int pl = 0; pl < np; ++pl) {
        int t = top[(size_t)pl * n + i];
        if (t < 0) continue;
        int idx = (int)stats[(size_t)pl * n + t].area2;
        if (idx > 0) b = bor(b, bbit(idx - 1));
    }
    F[i] = b;
}

// cluster image = dilate9(F) & original k-means cluster (DynaDetect.cc:683-684); area (:687); weighted-depth sum for
// calCenterPoint (:256-293); masked depth histogram for cal_hist (:1691-1696)
__global__ void k_rc_cluster_bits(const B128 *__restrict__ Fd, const int8_t *__restrict__ plane_of, const float *__restrict__ points,
                                  const uint16_t *__restrict__ depth, int n, ReclusterControl *ctl, B128 *__restrict__ CI, int *__restrict__ hist)
{
    __shared__ int s_area[RC_MAXC];
    __shared__ unsigned long long s_z[RC_MAXC];
    for (int j = threadIdx.x; j < RC_MAXC; j += blockDim.x) { s_area[j] = 0; s_z[j] = 0ull; }
    __syncthreads();
    // imgDepth / depth_max * 255 -> 8U (DynaDetect.cc:767-770): float scale, round half to even
    const double dmax = (double)ctl->depth_max;
    const float alpha = dmax > 0.0 ? (float)((1.0 / dmax) * 255.0) : 0.0f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int pl = plane_of[i];
        B128 b = b0();
        if (pl >= 0) b = band(Fd[i], ctl->plane_bits[pl]);
        CI[i] = b;
        if (bany(b)) {
            const long long zf = __double2ll_rn((double)points[3 * (size_t)i + 2] * KM_FIX_SCALE);
            int v = __float2int_rn((float)depth[i] * alpha);
            v = v > 255 ? 255 : v;
            // calcHist range {0,255} with 256 bins: bin = floor(v * 256/255), the value 255 falls outside
            const int bin = (int)floor((double)v * (256.0 / 255.0));
            while (bany(b)) {
                int c = bpop(b);
                atomicAdd(&s_area[c], 1);
                atomicAdd(&s_z[c], (unsigned long long)zf);
                if (bin < 256) atomicAdd(&hist[c * 256 + bin], 1);
            }
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < RC_MAXC; j += blockDim.x) {
        if (s_area[j]) {
            atomicAdd(&ctl->area[j], s_area[j]);
            atomicAdd((unsigned long long *)&ctl->zsum[j], s_z[j]);
        }
    }
}

// boundary pixels of F(c) and quads with a diagonal contour step (ccl.cuh: drawContours thickness 2)
__global__ void k_rc_bnd(const B128 *__restrict__ F, int W, int H, B128 *__restrict__ BND, B128 *__restrict__ DQ)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const int p = y * W + x;
    const B128 z = b0();
    const B128 a = F[p];
    const B128 l = x > 0 ? F[p - 1] : z, r = x < W - 1 ? F[p + 1] : z, u = y > 0 ? F[p - W] : z, d = y < H - 1 ? F[p + W] : z;
    BND[p] = band(a, bnot(band(band(l, r), band(u, d))));
    // quad with top-left pixel p: a=(x,y) b=(x+1,y) c=(x,y+1) d=(x+1,y+1)
    B128 dq = z;
    if (x < W - 1 && y < H - 1) {
        const B128 b = r, c = d, dd = F[p + W + 1];
        const B128 three = bor(bor(band(band(a, b), band(c, bnot(dd))), band(band(a, b), band(bnot(c), dd))),
                               bor(band(band(a, bnot(b)), band(c, dd)), band(band(bnot(a), b), band(c, dd))));
        const B128 two = bor(band(band(a, dd), band(bnot(b), bnot(c))), band(band(b, c), band(bnot(a), bnot(dd))));
        dq = bor(three, two);
    }
    DQ[p] = dq;
}

// drawContours(thickness 2) per component, then (optionally) - dilate10(edges), & imgLabelForSegEdge (DynaDetect.cc:692-697)
__global__ void k_rc_thick2(const B128 *__restrict__ BND, const B128 *__restrict__ DQ, int W, int H, const uint8_t *__restrict__ occl_dil,
                            const uint8_t *__restrict__ seg_dil, B128 *__restrict__ T1, int *__restrict__ cnt)
{
    __shared__ int s_cnt[RC_MAXC];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (cnt) { for (int j = tid; j < RC_MAXC; j += blockDim.x * blockDim.y) s_cnt[j] = 0; __syncthreads(); }
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < W && y < H) {
        const int p = y * W + x;
        B128 b = BND[p];
        if (x > 0) b = bor(b, BND[p - 1]);
        if (x < W - 1) b = bor(b, BND[p + 1]);
        if (y > 0) b = bor(b, BND[p - W]);
        if (y < H - 1) b = bor(b, BND[p + W]);
        // quads (x', y') whose 4x4-minus-corners block [x'-1, x'+2] x [y'-1, y'+2] covers p
#pragma unroll
        for (int dy = -2; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -2; dx <= 1; ++dx) {
                if ((dx == -2 || dx == 1) && (dy == -2 || dy == 1)) continue;
                const int qx = x + dx, qy = y + dy;
                if (qx >= 0 && qx < W && qy >= 0 && qy < H) b = bor(b, DQ[qy * W + qx]);
            }
        if (occl_dil && !(occl_dil[p] == 0 && seg_dil[p] != 0)) b = b0();
        T1[p] = b;
        if (cnt) while (bany(b)) atomicAdd(&s_cnt[bpop(b)], 1);
    }
    if (cnt) {
        __syncthreads();
        for (int j = tid; j < RC_MAXC; j += blockDim.x * blockDim.y)
            if (s_cnt[j]) atomicAdd(&cnt[j], s_cnt[j]);
    }
}

__global__ void k_rc_expand128(const B128 *__restrict__ T1, int n, const ReclusterControl *__restrict__ ctl, uint8_t *__restrict__ cls)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int nc = ctl->n_comp;
    const B128 b = T1[i];
    for (int c = 0; c < nc; ++c) cls[(size_t)c * n + i] = (btest(b, c) && ctl->cnt1[c] > 20) ? 1 : 0;
}

// lianjie(c) = union of the filled external contours of temp1_c with >= 30 points (DynaDetect.cc:698-716)
__global__ void k_rc_lianjie(const int *__restrict__ top, const RegionStats *__restrict__ stats, int n, ReclusterControl *ctl,
                             B128 *__restrict__ LJ)
{
    __shared__ int s_cnt[RC_MAXC];
    for (int j = threadIdx.x; j < RC_MAXC; j += blockDim.x) s_cnt[j] = 0;
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int nc = ctl->n_comp;
        B128 b = b0();
        for (int c = 0; c < nc; ++c) {
            int t = top[(size_t)c * n + i];
            if (t < 0) continue;
            const unsigned long long st = stats[(size_t)c * n + t].steps;
            long long steps = (long long)(st & 0xffffffffull) + (long long)(st >> 32);
            if (steps == 0) steps = 1;   // a single-pixel contour has one point
            if (steps >= 30) { b = bor(b, bbit(c)); atomicAdd(&s_cnt[c], 1); }
        }
        LJ[i] = b;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < RC_MAXC; j += blockDim.x)
        if (s_cnt[j]) atomicAdd(&ctl->lj_area[j], s_cnt[j]);
}

// pair counts: |dil_i & dil_j|, |dil_i & dil_j & occluded2|, |lianjie_i & lianjie_j| (DynaDetect.cc:829-866)
__global__ void k_rc_pairs(const B128 *__restrict__ CD, const B128 *__restrict__ LJ, const uint8_t *__restrict__ occl2, int n,
                           int *__restrict__ ov, int *__restrict__ ove, int *__restrict__ lo)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    B128 b = CD[i];
    if (__popcll(b.x) + __popcll(b.y) >= 2) {
        const bool e = occl2[i] != 0;
        while (bany(b)) {
            int c = bpop(b);
            B128 r = b;
            while (bany(r)) {
                int d = bpop(r);
                atomicAdd(&ov[c * RC_MAXC + d], 1);
                if (e) atomicAdd(&ove[c * RC_MAXC + d], 1);
            }
        }
    }
    b = LJ[i];
    if (__popcll(b.x) + __popcll(b.y) >= 2) {
        while (bany(b)) {
            int c = bpop(b);
            B128 r = b;
            while (bany(r)) atomicAdd(&lo[c * RC_MAXC + bpop(r)], 1);
        }
    }
}

// score = area * 0.0003 - centre.z, descending (DynaDetect.cc:736-748)
__global__ void k_rc_rank(ReclusterControl *ctl)
{
    __shared__ float sc[RC_MAXC];
    const int t = threadIdx.x, nc = ctl->n_comp;
    if (t < nc) {
        const float cnt = (float)ctl->area[t];
        const float z = (float)((double)ctl->zsum[t] * (1.0 / KM_FIX_SCALE)) / cnt;
        sc[t] = cnt * 0.0003f - z;
        ctl->score[t] = sc[t];
    }
    __syncthreads();
    if (t < nc) {
        int rank = 0;
        for (int j = 0; j < nc; ++j) rank += (sc[j] > sc[t]) || (sc[j] == sc[t] && j < t);
        ctl->order[rank] = t;
        ctl->rank_of[t] = rank;
    }
}

// one warp per pair of ranks (i < j): RAG entry (DynaDetect.cc:784-894)
__global__ void k_rc_rag(const ReclusterControl *__restrict__ ctl, const int *__restrict__ hist, const int *__restrict__ ov,
                         const int *__restrict__ ove, const int *__restrict__ lo, float *__restrict__ T)
{
    const int nc = ctl->n_comp;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nc * nc) return;
    const int i = warp / nc, j = warp - i * nc;
    if (i >= j) return;
    const int stride = nc + 1;
    const int c1 = ctl->order[i], c2 = ctl->order[j];
    const float a1 = (float)ctl->area[c1], a2 = (float)ctl->area[c2];
    float lessArea; int lessLabel;
    if (a1 < a2) { lessArea = a1; lessLabel = i; } else { lessArea = a2; lessLabel = j; }
    const int smallLabel = (int)fminf(0.7f * (float)nc, 15.0f);
    float wt = 1.0f;
    if (lessLabel < 10) wt = 0.7f;
    else if (lessLabel > smallLabel) wt = 2.0f;
    const int lo_c = min(c1, c2), hi_c = max(c1, c2);
    const int overlap = ov[lo_c * RC_MAXC + hi_c];
    float out = 0.0f;
    if ((float)overlap > fminf(200.0f, lessArea * 0.4f)) {
        // cal_hist (DynaDetect.cc:1685-1739)
        const int *h1i = hist + c1 * 256, *h2i = hist + c2 * 256;
        float h1[8], h2[8];
        int m1 = 0, m2 = 0, n1 = 0x7fffffff, n2 = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int u = h1i[lane + 32 * k], v = h2i[lane + 32 * k];
            h1[k] = (float)u; h2[k] = (float)v;
            m1 = max(m1, u); m2 = max(m2, v); n1 = min(n1, u); n2 = min(n2, v);
        }
        for (int o = 16; o > 0; o >>= 1) {
            m1 = max(m1, __shfl_xor_sync(0xffffffffu, m1, o)); m2 = max(m2, __shfl_xor_sync(0xffffffffu, m2, o));
            n1 = min(n1, __shfl_xor_sync(0xffffffffu, n1, o)); n2 = min(n2, __shfl_xor_sync(0xffffffffu, n2, o));
        }
        // normalize(NORM_MINMAX, 0..400) of the histogram with the larger peak; the other one is divided by (peak / 400)
        {
            const bool first = m1 > m2;
            const double smax = first ? (double)m1 : (double)m2, smin = first ? (double)n1 : (double)n2;
            const double scale = 400.0 * (smax - smin > 2.220446049250313e-16 ? 1.0 / (smax - smin) : 0.0);
            const float sa = (float)scale, sb = (float)(0.0 - smin * scale);
            const float div = (float)(1.0 / (smax / 400.0));
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (first) { h1[k] = h1[k] * sa + sb; h2[k] = h2[k] * div; }
                else { h2[k] = h2[k] * sa + sb; h1[k] = h1[k] * div; }
            }
        }
        double s1 = 0, s2 = 0, s11 = 0, s12 = 0, s22 = 0, sb = 0, si = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double a = h1[k], b = h2[k];
            s1 += a; s2 += b; s11 += a * a; s12 += a * b; s22 += b * b;
            sb += sqrt(a * b);
            si += fmin(a, b);
        }
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            s11 += __shfl_xor_sync(0xffffffffu, s11, o); s12 += __shfl_xor_sync(0xffffffffu, s12, o);
            s22 += __shfl_xor_sync(0xffffffffu, s22, o); sb += __shfl_xor_sync(0xffffffffu, sb, o);
            si += __shfl_xor_sync(0xffffffffu, si, o);
        }
        // compareHist: CORREL, BHATTACHARYYA, INTERSECT
        const double sc = 1.0 / 256.0;
        const double num = s12 - s1 * s2 * sc, den2 = (s11 - s1 * s1 * sc) * (s22 - s2 * s2 * sc);
        const double correl = fabs(den2) > 2.220446049250313e-16 ? num / sqrt(den2) : 1.0;
        double ss = s1 * s2;
        ss = fabs(ss) > 1.1920928955078125e-07 ? 1.0 / sqrt(ss) : 1.0;
        const double bhat = sqrt(fmax(1.0 - sb * ss, 0.0));
        const float v3 = (float)(correl + (1.0 - bhat) + si * 0.0005);
        const int ovEdge = ove[lo_c * RC_MAXC + hi_c];
        bool reject = false;
        if (ovEdge > 100 && lessLabel < smallLabel) reject = true;
        else if (v3 < 0.19f && lessLabel < smallLabel) reject = true;
        if (!reject) {
            float v2 = 0.0f;
            const int l1 = ctl->lj_area[c1], l2 = ctl->lj_area[c2];
            if (l1 > 0 && l2 > 0) {
                const int ol = lo[lo_c * RC_MAXC + hi_c];
                if (ol > 0 && ol > min(50, (int)(0.5 * (double)min(l1, l2)))) {
                    v2 = (float)ol;
                    if ((double)ol > 0.62 * (double)l1 || (double)ol > 0.62 * (double)l2) v2 = (float)max(250, ol);
                }
            }
            out = (v2 * 0.01f + v3) * 1.0f * wt;
        }
    }
    if (lane == 0) { T[i * stride + j] = out; T[j * stride + i] = out; }
}

// greedy merge + relabel (DynaDetect.cc:936-1016), strictly serial: the matrix is staged in shared memory and one
// thread walks it
__global__ void k_rc_merge(ReclusterControl *ctl, const float *__restrict__ Tg, int num_cluster)
{
    extern __shared__ float T[];
    __shared__ unsigned char merged[RC_MAXC + 1];
    __shared__ short mlist[RC_MAXC + 1][RC_MAXC];   // merge[i] lists
    __shared__ short mcount[RC_MAXC + 1];
    __shared__ float col[RC_MAXC + 1];
    const int n = ctl->n_comp, s = n + 1;
    for (int k = threadIdx.x; k < s * s; k += blockDim.x) T[k] = Tg[k];
    for (int k = threadIdx.x; k <= RC_MAXC; k += blockDim.x) { merged[k] = 0; mcount[k] = 0; }
    __syncthreads();
    if (threadIdx.x != 0) return;
    int countMerged = 0;
    auto do_merge = [&](int target, int j) {
        merged[j] = 1;
        mlist[target][mcount[target]++] = (short)j;
        for (int k = 0; k < s; ++k) col[k] = T[k * s + j];
        for (int k = 0; k < s; ++k) T[k * s + target] += col[k];
        for (int k = 0; k < s; ++k) T[target * s + k] += col[k];
        for (int k = 0; k < s; ++k) { T[k * s + j] = 0.0f; T[j * s + k] = 0.0f; }
    };
    for (int i = 0; i < min(num_cluster - 1 + countMerged, n); ++i) {
        for (int j = i + 1; j < min(num_cluster - 1 + countMerged, n); ++j) {
            const float score = T[j * s + i];
            if (score > 0.9f) {
                int toMerge = i;
                for (int k = 0; k < j; ++k)
                    if (T[k * s + j] > score) toMerge = k;
                do_merge(toMerge, j);
                ++countMerged;
            }
        }
    }
    for (int i = min(num_cluster - 1 + countMerged, n); i < n; ++i) {
        int mc = n;
        float best = 0.2f;
        for (int j = 0; j < i; ++j) {
            const float sc = T[j * s + i];
            if (sc > best) { best = sc; mc = j; }
        }
        do_merge(mc, i);
    }
    for (int k = 0; k < RC_MAXC + 2; ++k) ctl->lut[k] = 0;
    int idx = 1;
    for (int i = 0; i < n; ++i) {
        if (merged[i]) continue;
        ctl->lut[i] = (uint8_t)idx;
        for (int a = 0; a < mcount[i]; ++a) {
            const int ml = mlist[i][a];
            ctl->lut[ml] = (uint8_t)idx;
            for (int b = 0; b < mcount[ml]; ++b) ctl->lut[mlist[ml][b]] = (uint8_t)idx;   // two levels only (quirk B#12)
        }
        ++idx;
    }
    ctl->n_labels = idx - 1;
}

// imgTotalCluster.setTo(i, cluster_i) in rank order (later ranks overwrite) + relabel (DynaDetect.cc:749-754,996-1013)
__global__ void k_rc_apply(const B128 *__restrict__ CI, int n, const ReclusterControl *__restrict__ ctl, uint8_t *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    B128 b = CI[i];
    int best = -1;
    while (bany(b)) best = max(best, ctl->rank_of[bpop(b)]);
    out[i] = best >= 0 ? ctl->lut[best] : 0;
}

// ------------------------------------------------------------------ host side
int recluster_init(sindyn_base *ctx, ReclusterStage *r, int W, int H)
{
    r->W = W; r->H = H;
    const size_t N = (size_t)W * H;
    SD_CHECK(ctx->dalloc(&r->ctl, 1));
    SD_CHECK(ctx->dalloc(&r->plane_of, N));
    SD_CHECK(ctx->dalloc(&r->mcl, N));
    SD_CHECK(ctx->dalloc(&r->mcl_tmp, N));
    SD_CHECK(ctx->dalloc(&r->cls, N * (RC_MAXC + 2)));            // (+ 2: the decision stage's keyed planes, decide.cu)
    SD_CHECK(ctx->dalloc(&r->labels, (N + 1) * (RC_MAXC + 2)));
    SD_CHECK(ctx->dalloc(&r->top, N * RC_MAXC));
    SD_CHECK(ctx->dalloc(&r->stats, N * RC_MAXC));
    B128 **bp[] = {&r->F, &r->CI, &r->CD, &r->T1, &r->LJ, &r->tmpb, &r->tmpb2};
    for (B128 **p : bp) SD_CHECK(ctx->dalloc(p, N));
    uint8_t **up[] = {&r->occl_dil, &r->seg_dil, &r->tmp8, &r->tmp8b, &r->label_out, &r->pf_e, &r->occl1, &r->occl2};
    for (uint8_t **p : up) SD_CHECK(ctx->dalloc(p, N));
    SD_CHECK(ctx->dalloc(&r->hist, RC_MAXC * 256));
    SD_CHECK(ctx->dalloc(&r->ov, RC_MAXC * RC_MAXC));
    SD_CHECK(ctx->dalloc(&r->ove, RC_MAXC * RC_MAXC));
    SD_CHECK(ctx->dalloc(&r->lo, RC_MAXC * RC_MAXC));
    SD_CHECK(ctx->dalloc(&r->Tmat, (RC_MAXC + 1) * (RC_MAXC + 1)));
    CU_CHECK(ctx, cudaFuncSetAttribute(k_rc_merge, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(sizeof(float) * (RC_MAXC + 1) * (RC_MAXC + 1))));
    return SINDYN_OK;
}

int recluster_run(sindyn_base *ctx, ReclusterStage *r, const KmeansStage *km, const uint8_t *occl1, const uint8_t *occl2,
                  const uint16_t *depth)
{
    const int W = r->W, H = r->H, N = W * H;
    const dim3 blk2(32, 8), grd2(cdiv(W, 32), cdiv(H, 8));
    ReclusterControl *ctl = r->ctl;
    LAUNCH(ctx, k_rc_reset, 1, 128, 0, ctl, km->order);
    LAUNCH(ctx, k_rc_prepare, cdiv(N, 256), 256, 0, km->labels_u8, occl1, depth, N, km->order, ctl, r->plane_of, r->mcl);
    // MORPH_OPEN 4x4 of every (cluster - edges) image at once (DynaDetect.cc:671)
    SD_CHECK(morph_bits_run(ctx, r->mcl, r->mcl_tmp, W, H, 4, true, 2));
    SD_CHECK(morph_bits_run(ctx, r->mcl_tmp, r->mcl, W, H, 4, false, 2));
    LAUNCH(ctx, k_rc_expand16, cdiv(N, 256), 256, 0, r->mcl, N, &ctl->n_planes, r->cls);
    // findContours(RETR_EXTERNAL) + size / area tests (DynaDetect.cc:674-678)
    SD_CHECK(ccl_run(ctx, r->cls, r->labels, W, H, RC_MAXP, CCL_REGION, &ctl->n_planes));
    SD_CHECK(ccl_top_image(ctx, r->labels, r->top, W, H, RC_MAXP, &ctl->n_planes, r->stats));
    SD_CHECK(ccl_quad_stats_external(ctx, r->top, r->stats, W, H, RC_MAXP, &ctl->n_planes));
    LAUNCH(ctx, k_rc_qualify, dim3(cdiv(N, 256), 1, RC_MAXP), 256, 0, r->top, r->stats, N, ctl);
    LAUNCH(ctx, k_rc_sort, 1, RC_MAXC, 0, ctl, r->stats, N);
    LAUNCH(ctx, k_rc_fill_bits, cdiv(N, 256), 256, 0, r->top, r->stats, N, ctl, r->F);
    SD_CHECK(morph_bits_run(ctx, r->F, r->tmpb, W, H, 9, false, 16));
    CU_CHECK(ctx, cudaMemsetAsync(r->hist, 0, sizeof(int) * RC_MAXC * 256, ctx->stream));
    LAUNCH(ctx, k_rc_cluster_bits, SINDYN_NUM_SMS_B200 * 4, 256, 0, r->tmpb, r->plane_of, km->points, depth, N, ctl, r->CI, r->hist);
    SD_CHECK(morph_bits_run(ctx, r->CI, r->CD, W, H, 7, false, 16));
    // fake-edge strips (DynaDetect.cc:668-670,692-716)
    SD_CHECK(morph_run(ctx, occl1, r->occl_dil, r->tmp8, W, H, 10, MORPH_DILATE));
    SD_CHECK(morph_run(ctx, km->seg_edge, r->seg_dil, r->tmp8, W, H, 7, MORPH_DILATE));
    LAUNCH(ctx, k_rc_bnd, grd2, blk2, 0, r->F, W, H, r->tmpb, r->tmpb2);
    LAUNCH(ctx, k_rc_thick2, grd2, blk2, 0, r->tmpb, r->tmpb2, W, H, r->occl_dil, r->seg_dil, r->T1, ctl->cnt1);
    LAUNCH(ctx, k_rc_expand128, cdiv(N, 256), 256, 0, r->T1, N, ctl, r->cls);
    SD_CHECK(ccl_run(ctx, r->cls, r->labels, W, H, RC_MAXC, CCL_REGION, &ctl->n_comp));
    SD_CHECK(ccl_top_image(ctx, r->labels, r->top, W, H, RC_MAXC, &ctl->n_comp, r->stats));
    SD_CHECK(ccl_quad_stats_external(ctx, r->top, r->stats, W, H, RC_MAXC, &ctl->n_comp));
    LAUNCH(ctx, k_rc_lianjie, cdiv(N, 256), 256, 0, r->top, r->stats, N, ctl, r->LJ);
    // region adjacency graph
    CU_CHECK(ctx, cudaMemsetAsync(r->ov, 0, sizeof(int) * RC_MAXC * RC_MAXC, ctx->stream));
    CU_CHECK(ctx, cudaMemsetAsync(r->ove, 0, sizeof(int) * RC_MAXC * RC_MAXC, ctx->stream));
    CU_CHECK(ctx, cudaMemsetAsync(r->lo, 0, sizeof(int) * RC_MAXC * RC_MAXC, ctx->stream));
    LAUNCH(ctx, k_rc_pairs, cdiv(N, 256), 256, 0, r->CD, r->LJ, occl2, N, r->ov, r->ove, r->lo);
    LAUNCH(ctx, k_rc_rank, 1, RC_MAXC, 0, ctl);
    CU_CHECK(ctx, cudaMemsetAsync(r->Tmat, 0, sizeof(float) * (RC_MAXC + 1) * (RC_MAXC + 1), ctx->stream));
    LAUNCH(ctx, k_rc_rag, cdiv(RC_MAXC * RC_MAXC * 32, 256), 256, 0, ctl, r->hist, r->ov, r->ove, r->lo, r->Tmat);
    LAUNCH(ctx, k_rc_merge, 1, 256, sizeof(float) * (RC_MAXC + 1) * (RC_MAXC + 1), ctl, r->Tmat, KM_K);
    LAUNCH(ctx, k_rc_apply, cdiv(N, 256), 256, 0, r->CI, N, ctl, r->label_out);
    LAUNCH_CHECK(ctx);
    return SINDYN_OK;
}

// ------------------------------------------------------------------ plane-edge filter (DynaDetect.cc:598-641)
__global__ void k_pf_prepare(const uint8_t *__restrict__ plane_edges, const uint8_t *__restrict__ grad_edges, int n, uint8_t *__restrict__ cls,
                             ReclusterControl *ctl)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { ctl->pf_n = 0; ctl->pf_overflow = 0; }
    if (i >= n) return;
    cls[i] = plane_edges[i] > grad_edges[i] ? 1 : 0;   // imgEdgeByPlane - imgOccludedForPlane (saturating)
}

__global__ void k_pf_qualify(const int *__restrict__ top, RegionStats *__restrict__ stats, int n, ReclusterControl *ctl)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || top[i] != i) return;
    const unsigned long long st = stats[i].steps;
    long long steps = (long long)(st & 0xffffffffull) + (long long)(st >> 32);
    if (steps == 0) steps = 1;
    stats[i].area2 = 0;
    if (steps >= 25) {   // contours[i].size() < 25 -> skipped (DynaDetect.cc:611)
        int slot = atomicAdd(&ctl->pf_n, 1);
        if (slot < RC_PF_MAXC) { ctl->pf_root[slot] = i; ctl->pf_hit[slot] = 0; stats[i].area2 = slot + 1; }
        else ctl->pf_overflow = 1;
    }
}

__global__ void k_pf_fill_bits(const int *__restrict__ top, const RegionStats *__restrict__ stats, int n, B128 *__restrict__ F)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = top[i];
    B128 b = b0();
    if (t >= 0) {
        int idx = (int)stats[t].area2;
        if (idx > 0) b = bbit(idx - 1);
    }
    F[i] = b;
}

// does the dilated thick contour cover an end point? (DynaDetect.cc:621-631)
__global__ void k_pf_hits(const B128 *__restrict__ D, int W, const int *__restrict__ ep_xy, const int *__restrict__ ep_n, ReclusterControl *ctl)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *ep_n) return;
    B128 b = D[ep_xy[2 * i + 1] * W + ep_xy[2 * i]];
    while (bany(b)) { int c = bpop(b); if (c < RC_PF_MAXC) ctl->pf_hit[c] = 1; }
}

__global__ void k_pf_out(const B128 *__restrict__ E, const uint8_t *__restrict__ grad_edges, int n, const ReclusterControl *__restrict__ ctl,
                         uint8_t *__restrict__ occl2, uint8_t *__restrict__ occl1_pre)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    B128 b = E[i];
    bool on = false;
    while (bany(b)) { int c = bpop(b); if (c < RC_PF_MAXC && ctl->pf_hit[c]) on = true; }
    occl2[i] = on ? 255 : 0;
    occl1_pre[i] = (on || grad_edges[i]) ? 255 : 0;
}

int plane_edge_filter_run(sindyn_base *ctx, ReclusterStage *r, const uint8_t *plane_edges, const uint8_t *grad_edges, const int *ep_xy,
                          const int *ep_n)
{
    const int W = r->W, H = r->H, N = W * H;
    const dim3 blk2(32, 8), grd2(cdiv(W, 32), cdiv(H, 8));
    ReclusterControl *ctl = r->ctl;
    LAUNCH(ctx, k_pf_prepare, cdiv(N, 256), 256, 0, plane_edges, grad_edges, N, r->cls, ctl);
    SD_CHECK(ccl_run(ctx, r->cls, r->labels, W, H, 1, CCL_REGION, nullptr));
    SD_CHECK(ccl_top_image(ctx, r->labels, r->top, W, H, 1, nullptr, r->stats));
    SD_CHECK(ccl_quad_stats_external(ctx, r->top, r->stats, W, H, 1, nullptr));
    LAUNCH(ctx, k_pf_qualify, cdiv(N, 256), 256, 0, r->top, r->stats, N, ctl);
    LAUNCH(ctx, k_pf_fill_bits, cdiv(N, 256), 256, 0, r->top, r->stats, N, r->F);
    LAUNCH(ctx, k_rc_bnd, grd2, blk2, 0, r->F, W, H, r->tmpb, r->tmpb2);
    LAUNCH(ctx, k_rc_thick2, grd2, blk2, 0, r->tmpb, r->tmpb2, W, H, (const uint8_t *)nullptr, (const uint8_t *)nullptr, r->T1, (int *)nullptr);
    SD_CHECK(morph_bits_run(ctx, r->T1, r->tmpb, W, H, 10, false, 16));
    LAUNCH(ctx, k_pf_hits, cdiv(8192, 256), 256, 0, r->tmpb, W, ep_xy, ep_n, ctl);
    SD_CHECK(morph_bits_run(ctx, r->tmpb, r->tmpb2, W, H, 7, true, 16));
    LAUNCH(ctx, k_pf_out, cdiv(N, 256), 256, 0, r->tmpb2, grad_edges, N, ctl, r->occl2, r->tmp8);
    SD_CHECK(morph_run(ctx, r->tmp8, r->occl1, r->tmp8b, W, H, 3, MORPH_CLOSE));
    return SINDYN_OK;
}

// ------------------------------------------------------------------ plane contours (PEAC output stage)
__global__ void k_pl_expand(const B128 *__restrict__ bits, int n, const int *__restrict__ n_planes, uint8_t *__restrict__ cls)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int np = min(*n_planes, RC_MAXC);
    const B128 b = bits[i];
    for (int c = 0; c < np; ++c) cls[(size_t)c * n + i] = btest(b, c) ? 1 : 0;
}

__global__ void k_pl_fill(const int *__restrict__ top, int n, const int *__restrict__ n_planes, B128 *__restrict__ F)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int np = min(*n_planes, RC_MAXC);
    B128 b = b0();
    for (int c = 0; c < np; ++c)
        if (top[(size_t)c * n + i] >= 0) b = bor(b, bbit(c));
    F[i] = b;
}

__global__ void k_pl_any(const B128 *__restrict__ bits, int n, uint8_t *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = bany(bits[i]) ? 255 : 0;
}

int plane_contours_run(sindyn_base *ctx, ReclusterStage *r, const ulonglong2 *PB, const int *n_planes_dev, uint8_t *out)
{
    const int W = r->W, H = r->H, N = W * H;
    const dim3 blk2(32, 8), grd2(cdiv(W, 32), cdiv(H, 8));
    SD_CHECK(morph_bits_run(ctx, PB, r->tmpb, W, H, 3, false, 16));       // MORPH_CLOSE 3x3 of every plane image at once
    SD_CHECK(morph_bits_run(ctx, r->tmpb, r->tmpb2, W, H, 3, true, 16));
    LAUNCH(ctx, k_pl_expand, cdiv(N, 256), 256, 0, r->tmpb2, N, n_planes_dev, r->cls);
    SD_CHECK(ccl_run(ctx, r->cls, r->labels, W, H, RC_MAXC, CCL_REGION, n_planes_dev));
    SD_CHECK(ccl_top_image(ctx, r->labels, r->top, W, H, RC_MAXC, n_planes_dev, nullptr));
    LAUNCH(ctx, k_pl_fill, cdiv(N, 256), 256, 0, r->top, N, n_planes_dev, r->F);
    LAUNCH(ctx, k_rc_bnd, grd2, blk2, 0, r->F, W, H, r->tmpb, r->tmpb2);
    LAUNCH(ctx, k_rc_thick2, grd2, blk2, 0, r->tmpb, r->tmpb2, W, H, (const uint8_t *)nullptr, (const uint8_t *)nullptr, r->T1, (int *)nullptr);
    LAUNCH(ctx, k_pl_any, cdiv(N, 256), 256, 0, r->T1, N, out);
    LAUNCH_CHECK(ctx);
    return SINDYN_OK;
}
