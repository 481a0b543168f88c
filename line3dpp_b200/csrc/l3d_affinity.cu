// l3d_affinity.cu — pairwise affinities between 3D hypotheses and the replicator-dynamics diffusion.
//
//   k_affinity        : Line3D::similarity (line3D.cc:1467-1553) for every kept match whose two segments both have a
//                       3D estimate; the reference does this on the host under three mutexes (computingAffinityMatrix,
//                       line3D.cc:1852-1979).  One thread per match slot, deterministic output order.
//   l3d_rdd           : replicator_dynamics_diffusion_GPU (cudawrapper.cu:708-766) on CSR/CSC SoA arrays instead of two
//                       AoS float4 COO copies: 20 B/nnz/iteration of compulsory traffic instead of >= 64.  The arithmetic
//                       (positional lock-step product, per-row sequential sums, clamps) is kept operation for operation,
//                       so the result is bit-identical to the reference kernels; the linear search for the transposed
//                       slot (cudawrapper.cu:524-542) is replaced by a precomputed permutation.
#include "l3d_ctx.cuh"
#include "l3d_sweep.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <algorithm>
#include <cstdlib>
#include <vector>

#define L3D_EPS_D 1e-12
#define L3D_PI_D 3.14159265358979323846

struct A3 { double x, y, z; };
__device__ __forceinline__ A3 a3(double x, double y, double z) { A3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ A3 asub(A3 a, A3 b) { return a3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ double adot(A3 a, A3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double anorm(A3 a) { return sqrt(adot(a, a)); }
struct ASeg { A3 P1, P2, dir; float length; };
// Segment3D(P1,P2) (segment3D.h:48-66)
__device__ __forceinline__ ASeg make_seg(const double* p)
{
    ASeg s;
    A3 P1 = a3(p[0], p[1], p[2]), P2 = a3(p[3], p[4], p[5]);
    s.length = (float)anorm(asub(P1, P2));
    if (s.length > L3D_EPS_D) {
        s.P1 = P1; s.P2 = P2;
        A3 d = asub(P2, P1);
        double n2 = adot(d, d);
        if (n2 > 0) { double n = sqrt(n2); d = a3(d.x / n, d.y / n, d.z / n); }
        s.dir = d;
    } else { s.P1 = s.P2 = s.dir = a3(0, 0, 0); s.length = 0.0f; }
    return s;
}
// Segment3D::distance_Point2Line (segment3D.h:69-73): P1 + (dir * (P-P1)^T) * dir, evaluated as (outer product) * dir
__device__ __forceinline__ float dist_p2l(const ASeg& s, A3 P)
{
    A3 w = asub(P, s.P1);
    double d[3] = {s.dir.x, s.dir.y, s.dir.z}, ww[3] = {w.x, w.y, w.z}, h[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) h[i] = (d[i] * ww[0]) * d[0] + (d[i] * ww[1]) * d[1] + (d[i] * ww[2]) * d[2];
    A3 hp = a3(s.P1.x + h[0], s.P1.y + h[1], s.P1.z + h[2]);
    return (float)anorm(asub(hp, P));
}

// Line3D::similarity(s1, m1, seg2, truncate = false) (line3D.cc:1467-1553) for two segments that both have a 3D estimate:
// P1/P2 = their estimates (6 doubles each), m1/m2 = depths of their best matches, v1/v2 = their views
__device__ __forceinline__ float aff_similarity(const L3DViewDev* v1, const L3DViewDev* v2, const double* P1, const double* P2, float4 m1, float4 m2,
                                                float two_sigA_sqr, float med_scene_depth_lines)
{
    const ASeg s1 = make_seg(P1), s2 = make_seg(P2);
    if (s1.length < L3D_EPS_D || s2.length < L3D_EPS_D) return 0.0f;
    float dot_p = (float)adot(s1.dir, s2.dir);                               // angleBetweenSeg3D (line3D.cc:1571-1583)
    float angle = (float)((double)acosf(fmaxf(fminf(dot_p, 1.0f), -1.0f)) / L3D_PI_D * (double)180.0f);
    if (angle > 90.0f) angle = 180.0f - angle;
    float sim_a = expf(-angle * angle / two_sigA_sqr);
    float cutoff1 = v1->median_depth, cutoff2 = v2->median_depth;
    if (med_scene_depth_lines > L3D_EPS_D) { cutoff1 = fminf(cutoff1, med_scene_depth_lines); cutoff2 = fminf(cutoff2, med_scene_depth_lines); }
    float d11 = dist_p2l(s2, s1.P1), d12 = dist_p2l(s2, s1.P2), d21 = dist_p2l(s1, s2.P1), d22 = dist_p2l(s1, s2.P2);
    float sig11 = m1.x > cutoff1 ? cutoff1 * v1->k : m1.x * v1->k;
    float sig12 = m1.y > cutoff1 ? cutoff1 * v1->k : m1.y * v1->k;
    float reg11 = 2.0f * sig11 * sig11, reg12 = 2.0f * sig12 * sig12;
    float sig21 = m2.x > cutoff2 ? cutoff2 * v2->k : m2.x * v2->k;
    float sig22 = m2.y > cutoff2 ? cutoff2 * v2->k : m2.y * v2->k;
    float reg21 = 2.0f * sig21 * sig21, reg22 = 2.0f * sig22 * sig22;
    float sim_p1 = fminf(expf(-d11 * d11 / reg11), expf(-d12 * d12 / reg12));
    float sim_p2 = fminf(expf(-d21 * d21 / reg21), expf(-d22 * d22 / reg22));
    return fminf(sim_a, fminf(sim_p1, sim_p2));
}

// everything a thread needs to read the sweep's match store (l3d_sweep.cuh)
struct SwStore {
    const SwView* vt; const L3DPairDev* pairs; const long long* row_off; int num_pairs, knn; const l3d_match_rec* recs;
    const unsigned int* e_val; const unsigned char* e_flag;
};
// depths of the best match of global segment g (its estimate), b = est_best[g]
__device__ __forceinline__ float4 sw_best_depths(const SwStore& M, int view, int b) { return sw_depths(M.e_val[M.vt[view].region_off + b], M.recs); }

__global__ void __launch_bounds__(256)
k_affinity(const L3DViewDev* __restrict__ views, const long long* __restrict__ region_off, const int* __restrict__ order,
           int V, long long K, const long long* __restrict__ kx, const SwStore M, const int* __restrict__ est_best,
           const double* __restrict__ est_P, float two_sigA_sqr, float med_scene_depth_lines, float min_affinity,
           float* __restrict__ sim_out, int* __restrict__ flag_out, long long* __restrict__ gi_out, long long* __restrict__ gj_out)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= K) return;
    const long long x = kx[q];                   // the q-th kept match in emission order
    int flag = 0; float sim = 0.0f; long long gi = -1, gj = -1;
    {
        int lo = 0, hi = V - 1;                  // processing rank whose region contains x
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (region_off[mid] <= x) lo = mid; else hi = mid - 1; }
        const int v1i = order[lo];
        const SwEntry me = sw_decode(M.e_val[x], M.pairs, M.row_off, M.num_pairs, M.knn, M.recs);
        const L3DViewDev* v1 = views + v1i;
        const L3DViewDev* v2 = views + me.tgt_view;
        gi = v1->seg_off + me.seg; gj = v2->seg_off + me.tgt_seg;
        const int b1 = est_best[gi], b2 = est_best[gj];
        if (b1 >= 0 && b2 >= 0) {
            sim = aff_similarity(v1, v2, est_P + 6 * gi, est_P + 6 * gj, sw_best_depths(M, v1i, b1), sw_best_depths(M, me.tgt_view, b2),
                                 two_sigA_sqr, med_scene_depth_lines);
            flag = sim > min_affinity ? 1 : 0;
        }
    }
    sim_out[q] = sim; flag_out[q] = flag; gi_out[q] = gi; gj_out[q] = gj;
}

// ---- affinity candidates WITH collinearity links (line3D.cc:1904-1974), one thread per segment in estimate order ----
// Events of estimate i, in the reference's order: for every kept match with sim > min_affinity the direct edge, then one
// edge to every segment collinear with the match's target (if that one has an estimate and sim > min_affinity); after
// the matches, if any direct edge exists, the edges to the segments collinear with i itself.  Whether an event really
// enters A_ depends on unused() and, for the collinear ones, on their parent having passed unused() - that is resolved
// by l3d_affinity_matrix.  par: -1 direct, >= 0 index of the parent direct event, -2 "any direct event of my source".
// FILL = false counts (evcnt[t]); FILL = true writes at evptr[t].
template <bool FILL>
__global__ void __launch_bounds__(128)
k_aff_events(const L3DViewDev* __restrict__ views, const long long* __restrict__ region_off, const int* __restrict__ order,
             const long long* __restrict__ segrank_off, int V, long long N, const SwStore M,
             const int2* __restrict__ ranges, const int* __restrict__ est_best, const double* __restrict__ est_P,
             long long K, const long long* __restrict__ kx, const float* __restrict__ sim_slot, const int* __restrict__ flag_slot, const long long* __restrict__ cptr,
             const int* __restrict__ cidx, float two_sigA_sqr, float med_scene_depth_lines, float min_affinity, int* __restrict__ evcnt,
             const long long* __restrict__ evptr, long long* __restrict__ out_i, long long* __restrict__ out_j, float* __restrict__ out_w,
             int* __restrict__ out_par)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    int lo = 0, hi = V - 1;                      // processing rank of segment t
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (segrank_off[mid] <= t) lo = mid; else hi = mid - 1; }
    const int v1i = order[lo];
    const L3DViewDev* v1 = views + v1i;
    const long long g = v1->seg_off + (t - segrank_off[lo]);
    const int b1 = est_best[g];
    long long n = 0;
    const long long base = FILL ? evptr[t] : 0;
    if (b1 >= 0) {
        const long long ro = region_off[lo];
        const int2 rng = ranges[g];
        const double* P1 = est_P + 6 * g;
        const float4 m1 = sw_best_depths(M, v1i, b1);
        bool found = false;
        long long q = 0;
        if (rng.x >= 0) {                        // first kept match of this segment in the compact list
            long long a = 0, b = K;
            while (a < b) { const long long mid = (a + b) >> 1; if (kx[mid] < ro + rng.x) a = mid + 1; else b = mid; }
            q = a;
        }
        for (; rng.x >= 0 && q < K && kx[q] <= ro + rng.y; ++q) {
            const long long x = kx[q];
            if (!flag_slot[q]) continue;
            const SwEntry me = sw_decode(M.e_val[x], M.pairs, M.row_off, M.num_pairs, M.knn, M.recs);
            const L3DViewDev* v2 = views + me.tgt_view;
            const long long gj = v2->seg_off + me.tgt_seg;
            const long long parent = base + n;
            if (FILL) { out_i[parent] = g; out_j[parent] = gj; out_w[parent] = sim_slot[q]; out_par[parent] = -1; }
            ++n; found = true;
            for (long long q = cptr[gj]; q < cptr[gj + 1]; ++q) {                  // collinear with the target (line3D.cc:1904-1937)
                const long long g2 = v2->seg_off + cidx[q];
                const int b2 = est_best[g2];
                if (b2 < 0) continue;
                const float s2 = aff_similarity(v1, v2, P1, est_P + 6 * g2, m1, sw_best_depths(M, me.tgt_view, b2), two_sigA_sqr, med_scene_depth_lines);
                if (s2 > min_affinity) {
                    if (FILL) { out_i[base + n] = g; out_j[base + n] = g2; out_w[base + n] = s2; out_par[base + n] = (int)parent; }
                    ++n;
                }
            }
        }
        if (found)
            for (long long q = cptr[g]; q < cptr[g + 1]; ++q) {                    // collinear with the source (line3D.cc:1941-1974)
                const long long g2 = v1->seg_off + cidx[q];
                const int b2 = est_best[g2];
                if (b2 < 0) continue;
                const float s2 = aff_similarity(v1, v1, P1, est_P + 6 * g2, m1, sw_best_depths(M, v1i, b2), two_sigA_sqr, med_scene_depth_lines);
                if (s2 > min_affinity) {
                    if (FILL) { out_i[base + n] = g; out_j[base + n] = g2; out_w[base + n] = s2; out_par[base + n] = -2; }
                    ++n;
                }
            }
    }
    if (!FILL) evcnt[t] = (int)n;
}

__global__ void __launch_bounds__(256)
k_affinity_compact(long long total, const int* __restrict__ flag, const long long* __restrict__ pos, const float* __restrict__ sim,
                   const long long* __restrict__ gi, const long long* __restrict__ gj, long long* __restrict__ out_i,
                   long long* __restrict__ out_j, float* __restrict__ out_w)
{
    const long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= total || !flag[x]) return;
    const long long p = pos[x];
    out_i[p] = gi[x]; out_j[p] = gj[x]; out_w[p] = sim[x];
}

// ================================================================================================ diffusion
__global__ void __launch_bounds__(256)
k_rdd_keys(long long nnz, const int* __restrict__ ei, const int* __restrict__ ej, unsigned long long* __restrict__ krow,
           unsigned long long* __restrict__ kcol, unsigned int* __restrict__ idx)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    krow[e] = ((unsigned long long)(unsigned int)ei[e] << 32) | (unsigned int)ej[e];
    kcol[e] = ((unsigned long long)(unsigned int)ej[e] << 32) | (unsigned int)ei[e];
    idx[e] = (unsigned int)e;
}
// gather sorted values + row/col of each entry, mark row (or col) starts
__global__ void __launch_bounds__(256)
k_rdd_gather(long long nnz, const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ idx,
             const float* __restrict__ w, float* __restrict__ val, int* __restrict__ major, int* __restrict__ minor)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const unsigned long long k = keys[e];
    const int ma = (int)(k >> 32), mi = (int)(k & 0xFFFFFFFFull);
    val[e] = w[idx[e]]; major[e] = ma; minor[e] = mi;
}
// CSR/CSC pointers: ptr[r] = first sorted position whose major index is >= r (r = 0..n), by binary search
__global__ void __launch_bounds__(256) k_rdd_ptr(int n, int nnz, const int* __restrict__ major, int* __restrict__ ptr)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    int lo = 0, hi = nnz;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (major[mid] < r) lo = mid + 1; else hi = mid; }
    ptr[r] = lo;
}
// transposed slot of every row-sorted entry (a,b): first position of (b,a) in row b, -1 if absent
__global__ void __launch_bounds__(256)
k_rdd_tslot(long long nnz, const int* __restrict__ prow, const int* __restrict__ pcol, const int* __restrict__ rowptr, int* __restrict__ tslot)
{
    const long long y = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= nnz) return;
    const int a = prow[y], b = pcol[y];
    int lo = rowptr[b], hi = rowptr[b + 1];
    const int end = hi;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (pcol[mid] < a) lo = mid + 1; else hi = mid; }
    tslot[y] = (lo < end && pcol[lo] == a) ? lo : -1;
}
// ---- sector-aligned rows ----------------------------------------------------------------------------------------------
// The lock-step walk reads a run of P.row(r) and a run of W.col(c) per entry.  With 4-byte loads every thread of a warp
// touches a different 128-byte line on every step (one L1 wavefront per thread and step: the first version was bound by
// exactly that).  Values are therefore kept in PADDED arrays whose rows/columns start on 32-byte boundaries (rp4/cp4 =
// start in float4 units, always even; rows padded with zeros to a multiple of 8); the walk loads float4 and still adds
// the products strictly in k order, so the sums round exactly like the reference's.
__global__ void __launch_bounds__(256) k_rdd_len4(int n, const int* __restrict__ ptr, int* __restrict__ len4)
{
    // rows padded to a multiple of EIGHT floats: every row starts on a 32-byte sector boundary, so the back-to-back float4
    // loads of a walk use whole sectors (with 16-byte alignment half of every gathered sector was wasted L2 bandwidth)
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) len4[r] = 2 * ((ptr[r + 1] - ptr[r] + 7) >> 3);
}
__global__ void __launch_bounds__(256) k_rdd_pad(long long nnz, const int* __restrict__ major, const int* __restrict__ ptr, const int* __restrict__ p4,
                                                 const float* __restrict__ val, float* __restrict__ padded)
{
    const long long y = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= nnz) return;
    const int r = major[y];
    padded[4ll * p4[r] + (y - ptr[r])] = val[y];
}
__global__ void __launch_bounds__(256) k_rdd_unpad(long long nnz, const int* __restrict__ major, const int* __restrict__ ptr, const int* __restrict__ p4,
                                                   const float* __restrict__ padded, float* __restrict__ val)
{
    const long long y = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= nnz) return;
    const int r = major[y];
    val[y] = padded[4ll * p4[r] + (y - ptr[r])];
}
// per-entry walk descriptors, built once (the structure is constant over the iterations): everything the step needs is
// then read with coalesced 16-byte loads instead of eight dependent scattered pointer look-ups per entry
//   x = first float4 of P.row(r)   y = first float4 of W.col(c)   z = walk length min(len_r, len_c)   w = own padded slot
// and dst = padded slot of the transposed entry P'(r,c) (or -1)
__global__ void __launch_bounds__(256)
k_rdd_plan(long long nnz, const int* __restrict__ prow, const int* __restrict__ pcol, const int* __restrict__ rowptr, const int* __restrict__ colptr,
           const int* __restrict__ rp4, const int* __restrict__ cp4, const int* __restrict__ tslot, int4* __restrict__ plan, int* __restrict__ dst)
{
    const long long y = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= nnz) return;
    const int c = prow[y], r = pcol[y];                 // "transpose" (cudawrapper.cu:493-495)
    const int rs = rowptr[r], m = min(rowptr[r + 1] - rs, colptr[c + 1] - colptr[c]);
    plan[y] = make_int4(rp4[r], cp4[c], m, 4 * rp4[c] + (int)(y - rowptr[c]));
    const int t = tslot[y];
    dst[y] = t >= 0 ? 4 * rp4[r] + (t - rs) : -1;
}
// K_sparseMat_diffusion_step (cudawrapper.cu:480-544) on the padded arrays: entry y = (a,b) of P produces P'(b,a).
// One thread per entry, walking in steps of EIGHT values = one whole 32-byte sector of P.row(b) and of W.col(a) per step
// (one 256-bit load each); the first two steps are in flight together, which is the memory-level parallelism this
// latency-bound gather needs at 6 resident CTAs per SM.  No per-element guards: beyond the walk length min(len_r, len_c) at least one of the two factors
// lies in its row's zero padding (rows are padded to 8), the product is +0 and `mul + 0 == mul` exactly, so every entry
// still accumulates exactly the reference's products in the reference's k order (weights and P are finite).
// one whole 32-byte sector per lane and instruction (LDG.E.256, new on sm_100): with divergent addresses the L1 looks up
// about one sector per cycle and SM, so two 16-byte loads of the same sector cost twice as much as one 32-byte load -
// the first float4 version of this kernel sat at 80 % L1 throughput for exactly that reason
#ifndef RDD_MINB
#define RDD_MINB 5
#endif
struct __align__(32) F8 { float v[8]; };
__device__ __forceinline__ F8 ld256(const float* p)
{
    F8 r;
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    return r;
}
__global__ void __launch_bounds__(256, RDD_MINB)
k_rdd_step8(long long nnz, const int4* __restrict__ plan, const int* __restrict__ dst, const float* __restrict__ Pp, const float* __restrict__ Wp,
            float* __restrict__ Pnp)
{
    const long long y = (long long)blockIdx.x * 256 + threadIdx.x;
    if (y >= nnz) return;
    const int4 pl = plan[y];
    const int d = dst[y];
    const int n8 = (pl.z + 7) >> 3;
    const float* pr = Pp + 4ll * pl.x;
    const float* wc = Wp + 4ll * pl.y;
    // the first two steps (16 values: most walks) are requested together, before anything waits on them
    F8 p0, w0, p1, w1;
    if (n8 > 0) { p0 = ld256(pr); w0 = ld256(wc); }
    if (n8 > 1) { p1 = ld256(pr + 8); w1 = ld256(wc + 8); }
    const float own = (pl.z > 0 || d >= 0) ? Pp[pl.w] : 0.0f;
    float m = 0.0f;
    if (n8 > 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) m += p0.v[e] * w0.v[e];
    }
    if (n8 > 1) {
#pragma unroll
        for (int e = 0; e < 8; ++e) m += p1.v[e] * w1.v[e];
    }
    for (int k8 = 2; k8 < n8; ++k8) {
        const F8 pv = ld256(pr + 8 * k8), wv = ld256(wc + 8 * k8);
#pragma unroll
        for (int e = 0; e < 8; ++e) m += pv.v[e] * wv.v[e];
    }
    m *= own;                                           // times P(a,b) itself
    if (m < L3D_EPS_F) m = L3D_EPS_F;
    if (d >= 0) Pnp[d] = m;
}
// K_sparseMat_row_normalization (cudawrapper.cu:432-477), 4 lanes per row, one float4 each per pass: coalesced loads and
// stores, and the row sum is still added strictly in slot order - the running sum is handed from lane to lane (lane j adds
// its four values to what lane j-1 produced).  No guards: padding slots hold +0 (x + 0 == x exactly, 0 / sum == 0).
__global__ void __launch_bounds__(256)
k_rdd_normalize_q(int n, const int* __restrict__ rowptr, const int* __restrict__ rp4, float* __restrict__ Pp)
{
    // control flow is kept warp-uniform (trip count = the longest of the warp's 8 rows, no early exit) so that the shuffles can
    // name the full warp: sub-warp masks compile to a WARPSYNC / collective sequence per shuffle, which made the first version
    // of this kernel instruction-bound
    const int lane = threadIdx.x & 31, gl = lane & 3, g0 = lane & ~3;
    const long long r = ((long long)blockIdx.x * 256 + threadIdx.x) >> 2;
    const int len = r < n ? rowptr[r + 1] - rowptr[r] : 0;
    const int n4 = 2 * ((len + 7) >> 3);                 // float4 per padded row (k_rdd_len4)
    float4* row = reinterpret_cast<float4*>(Pp) + (r < n ? rp4[r] : 0);
    const int n4max = __reduce_max_sync(0xffffffffu, n4);
    float sum = 0.0f;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f);         // this lane's float4 of the first chunk (rows of <= 16 values are done from it)
    for (int c0 = 0; c0 < n4max; c0 += 4) {
        const float4 v = c0 + gl < n4 ? row[c0 + gl] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 == 0) v0 = v;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float t = sum; t += v.x; t += v.y; t += v.z; t += v.w;          // what lane j of the group contributes, in slot order
            sum = __shfl_sync(0xffffffffu, t, g0 + j);
        }
    }
    if (len == 0) return;
    if (sum < L3D_EPS_F) sum = L3D_EPS_F;
    // padding slots hold 0: 0 / sum is 0, but a zero dividend sends the IEEE divide down its slow path (FCHK), which made 52 % of
    // this kernel's instructions - skip them
#define RDD_DIV(x) if (x != 0.0f) { asm volatile(""); x = x / sum; }      /* the empty asm keeps the branch from being if-converted */
    if (n4 <= 4) { if (gl < n4) { RDD_DIV(v0.x) RDD_DIV(v0.y) RDD_DIV(v0.z) RDD_DIV(v0.w) row[gl] = v0; } }
    else for (int c = gl; c < n4; c += 4) { float4 w = row[c]; RDD_DIV(w.x) RDD_DIV(w.y) RDD_DIV(w.z) RDD_DIV(w.w) row[c] = w; }
#undef RDD_DIV
}

__global__ void __launch_bounds__(256) k_iota(long long n, unsigned int* __restrict__ idx)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = (unsigned int)i;
}

// ---- affinity matrix bookkeeping on the device (see l3d_affinity_matrix) ----------------------------------------
static int bits_i64(long long n) { int b = 1; while ((1ll << b) < n) ++b; return b; }
__global__ void __launch_bounds__(256) k_aff_keys(long long ne, const long long* __restrict__ gi, const long long* __restrict__ gj,
                                                  unsigned long long* __restrict__ key, unsigned int* __restrict__ val)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    const unsigned long long a = (unsigned long long)min(gi[e], gj[e]), b = (unsigned long long)max(gi[e], gj[e]);
    key[e] = (a << 32) | b; val[e] = (unsigned int)e;
}
// unused() with conditional events (line3D.cc:1881-1974): event e enters A_ iff its parent did (direct events have none)
// and no EARLIER event of the same unordered pair did.  All dependencies point to earlier events, so iterating
//   acc[e] = parent_ok(e, acc_prev) && no earlier event e' of the same pair with parent_ok(e', acc_prev)
// from acc = 0 reaches the sequential answer after (longest dependency chain) rounds and then stays.  Without
// collinearity links every event is direct and the first round is already final (head of every run of the stable sort).
__device__ __forceinline__ bool aff_parent_ok(int e, const int* __restrict__ par, const long long* __restrict__ gi, const int* __restrict__ acc_prev,
                                              const int* __restrict__ found_prev)
{
    if (!par) return true;
    const int p = par[e];
    return p == -1 ? true : p >= 0 ? acc_prev[p] != 0 : found_prev[gi[e]] != 0;
}
__global__ void __launch_bounds__(256)
k_aff_round(long long ne, const unsigned long long* __restrict__ key, const unsigned int* __restrict__ val, const int* __restrict__ par,
            const long long* __restrict__ gi, const int* __restrict__ acc_prev, const int* __restrict__ found_prev, int* __restrict__ acc_new,
            int* __restrict__ changed)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ne) return;
    const int e = (int)val[j];
    bool a = aff_parent_ok(e, par, gi, acc_prev, found_prev);
    const unsigned long long k = key[j];
    for (long long jj = j - 1; a && jj >= 0 && key[jj] == k; --jj)          // stable sort: earlier position = earlier in emission order
        if (aff_parent_ok((int)val[jj], par, gi, acc_prev, found_prev)) a = false;
    acc_new[e] = a ? 1 : 0;
    if ((acc_prev[e] != 0) != a) *changed = 1;
}
// found_aff of every source segment (line3D.cc:1900): some DIRECT event of it was accepted
__global__ void __launch_bounds__(256)
k_aff_found(long long ne, const int* __restrict__ par, const long long* __restrict__ gi, const int* __restrict__ acc, int* __restrict__ found)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < ne && acc[e] && par[e] == -1) found[gi[e]] = 1;
}
__global__ void __launch_bounds__(256) k_aff_time(long long ne, const int* __restrict__ keep, const long long* __restrict__ q, const long long* __restrict__ gi,
                                                  const long long* __restrict__ gj, unsigned long long* __restrict__ time)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne || !keep[e]) return;
    atomicMin(time + gi[e], 2ull * (unsigned long long)q[e]);          // id1 is looked up before id2 (line3D.cc:1884-1887)
    atomicMin(time + gj[e], 2ull * (unsigned long long)q[e] + 1ull);
}
__global__ void __launch_bounds__(256) k_aff_nodeflags(long long N, const unsigned long long* __restrict__ time, int* __restrict__ flag)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < N) flag[g] = time[g] != ~0ull;
}
__global__ void __launch_bounds__(256) k_aff_nodes(long long N, const int* __restrict__ flag, const long long* __restrict__ pos, const unsigned long long* __restrict__ time,
                                                   unsigned long long* __restrict__ key, unsigned int* __restrict__ val)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N || !flag[g]) return;
    key[pos[g]] = time[g]; val[pos[g]] = (unsigned int)g;
}
__global__ void __launch_bounds__(256) k_aff_ids(long long n, const unsigned int* __restrict__ seg_sorted, int* __restrict__ id_of, long long* __restrict__ l2g)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    id_of[seg_sorted[i]] = (int)i; l2g[i] = (long long)seg_sorted[i];
}
__global__ void __launch_bounds__(256) k_aff_emit(long long ne, const int* __restrict__ keep, const long long* __restrict__ q, const long long* __restrict__ gi,
                                                  const long long* __restrict__ gj, const float* __restrict__ w, const int* __restrict__ id_of,
                                                  int* __restrict__ ei, int* __restrict__ ej, float* __restrict__ ew)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne || !keep[e]) return;
    const long long o = 2 * q[e];
    const int a = id_of[gi[e]], b = id_of[gj[e]];
    ei[o] = a; ej[o] = b; ew[o] = w[e];
    ei[o + 1] = b; ej[o + 1] = a; ew[o + 1] = w[e];
}

struct KeptFlag { const unsigned char* f; __host__ __device__ int operator()(long long x) const { return (f[x] & SW_KEPT) ? 1 : 0; } };

extern "C" {

// device part shared by l3d_affinity_edges / l3d_affinity_matrix: similarities of all kept matches, then the candidates with
// sim > min_affinity compacted IN EMISSION ORDER into S.d_aff_oi / oj / ow.  Returns their number (or < 0).
static long long affinity_candidates(l3d_ctx* c, float two_sigA_sqr, float med_scene_depth_lines, float min_affinity)
{
    SweepState& S = c->sweep;
    const long long total = S.total;
    if (total == 0) return 0;
    const int V = c->num_views;
    int rc;
    DevBuf &d_sim = S.d_aff_sim, &d_flag = S.d_aff_flag, &d_gi = S.d_aff_gi, &d_gj = S.d_aff_gj, &d_pos = S.d_aff_pos;
    cudaStream_t st = c->stream;
    // the kept matches (filterMatches survivors), compacted in emission order: everything below works on this list, not on the
    // whole match store (2.5e9 entries at BASELINE configs[4])
    const long long BLK = 1ll << 30;
    long long K = 0;
    if ((rc = l3d_reserve(c, d_pos, 16, "kept count"))) return rc;
    std::vector<long long> kept_in_block;
    for (long long base = 0; base < total; base += BLK) {
        const int nblk = (int)std::min(BLK, total - base);
        cub::CountingInputIterator<long long> idx(base);
        cub::TransformInputIterator<int, KeptFlag, cub::CountingInputIterator<long long> > flags(idx, KeptFlag{(const unsigned char*)S.d_eflag.p});
        size_t tb = 0;
        cub::DeviceReduce::Sum(nullptr, tb, flags, (int*)d_pos.p, nblk, st);
        if ((rc = l3d_reserve(c, S.d_sort_tmp, tb, "reduce temp"))) return rc;
        tb = S.d_sort_tmp.cap;
        L3D_CUDA(c, cub::DeviceReduce::Sum(S.d_sort_tmp.p, tb, flags, (int*)d_pos.p, nblk, st), "kept count");
        int cnt = 0;
        L3D_CUDA(c, cudaMemcpyAsync(&cnt, d_pos.p, 4, cudaMemcpyDeviceToHost, st), "kept count");
        L3D_CUDA(c, cudaStreamSynchronize(st), "kept count");
        kept_in_block.push_back(cnt); K += cnt;
    }
    S.n_kept = K;
    if (K == 0) return 0;
    if ((rc = l3d_reserve(c, S.d_kx, 8 * (size_t)K, "kept list"))) return rc;
    {
        long long done = 0; size_t bi = 0;
        for (long long base = 0; base < total; base += BLK, ++bi) {
            const int nblk = (int)std::min(BLK, total - base);
            if (kept_in_block[bi] == 0) continue;
            cub::CountingInputIterator<long long> idx(base);
            cub::TransformInputIterator<int, KeptFlag, cub::CountingInputIterator<long long> > flags(idx, KeptFlag{(const unsigned char*)S.d_eflag.p});
            size_t tb = 0;
            cub::DeviceSelect::Flagged(nullptr, tb, idx, flags, (long long*)S.d_kx.p + done, (int*)d_pos.p, nblk, st);
            if ((rc = l3d_reserve(c, S.d_sort_tmp, tb, "select temp"))) return rc;
            tb = S.d_sort_tmp.cap;
            L3D_CUDA(c, cub::DeviceSelect::Flagged(S.d_sort_tmp.p, tb, idx, flags, (long long*)S.d_kx.p + done, (int*)d_pos.p, nblk, st), "kept list");
            done += kept_in_block[bi];
        }
        c->launches += 2 * (long long)kept_in_block.size();
    }
    if ((rc = l3d_reserve(c, d_sim, 4 * (size_t)K, "affinity sims"))) return rc;
    if ((rc = l3d_reserve(c, d_flag, 4 * (size_t)K, "affinity flags"))) return rc;
    if ((rc = l3d_reserve(c, d_gi, 8 * (size_t)K, "affinity gi"))) return rc;
    if ((rc = l3d_reserve(c, d_gj, 8 * (size_t)K, "affinity gj"))) return rc;
    if ((rc = l3d_reserve(c, d_pos, 8 * (size_t)(K + 1), "affinity pos"))) return rc;
    std::vector<int> rank_of_view(V);
    for (int i = 0; i < V; ++i) rank_of_view[S.order[i]] = i;
    if ((rc = l3d_reserve(c, S.d_order, 4 * (size_t)V, "order"))) return rc;
    if ((rc = l3d_reserve(c, S.d_rankofview, 4 * (size_t)V, "rank of view"))) return rc;
    if ((rc = l3d_reserve(c, S.d_region_off, 8 * (size_t)(V + 1), "region offsets"))) return rc;
    L3D_CUDA(c, cudaMemcpyAsync(S.d_order.p, S.order.data(), 4 * (size_t)V, cudaMemcpyHostToDevice, st), "order");
    L3D_CUDA(c, cudaMemcpyAsync(S.d_rankofview.p, rank_of_view.data(), 4 * (size_t)V, cudaMemcpyHostToDevice, st), "rank of view");
    L3D_CUDA(c, cudaMemcpyAsync(S.d_region_off.p, S.region_off.data(), 8 * (size_t)(V + 1), cudaMemcpyHostToDevice, st), "region offsets");
    const unsigned int nb = (unsigned int)((K + 255) / 256);
    SwStore M;
    M.vt = (const SwView*)S.d_vt.p; M.pairs = (const L3DPairDev*)c->d_pairs.p; M.row_off = (const long long*)S.d_rowoff.p; M.num_pairs = c->num_pairs; M.knn = c->knn;
    M.recs = (const l3d_match_rec*)c->d_recs.p; M.e_val = (const unsigned int*)S.d_eval.p; M.e_flag = (const unsigned char*)S.d_eflag.p;
    k_affinity<<<nb, 256, 0, st>>>(c->views(), (const long long*)S.d_region_off.p, (const int*)S.d_order.p, V, K, (const long long*)S.d_kx.p, M, (const int*)S.d_est_best.p,
                                   (const double*)S.d_est_P.p, two_sigA_sqr, med_scene_depth_lines, min_affinity, (float*)d_sim.p, (int*)d_flag.p,
                                   (long long*)d_gi.p, (long long*)d_gj.p);
    if (c->collin.valid) {
        // collinearity links on: events per segment in estimate order (count, scan, fill), each with its parent
        const CollinState& CL = c->collin;
        const long long N = c->total_segs;
        std::vector<long long> segrank_off((size_t)V + 1, 0);
        for (int i = 0; i < V; ++i) segrank_off[i + 1] = segrank_off[i] + c->h_views[S.order[i]].nseg;
        if ((rc = l3d_reserve(c, S.d_segrank_off, 8 * ((size_t)V + 1), "segment rank offsets"))) return rc;
        if ((rc = l3d_reserve(c, S.d_evcnt, 4 * (size_t)(N + 1), "event counts"))) return rc;
        if ((rc = l3d_reserve(c, S.d_evptr, 8 * (size_t)(N + 1), "event offsets"))) return rc;
        L3D_CUDA(c, cudaMemcpyAsync(S.d_segrank_off.p, segrank_off.data(), 8 * ((size_t)V + 1), cudaMemcpyHostToDevice, st), "segment rank offsets");
        L3D_CUDA(c, cudaMemsetAsync(S.d_evcnt.p, 0, 4 * (size_t)(N + 1), st), "event counts");
        const unsigned int nbs = (unsigned int)((N + 127) / 128);
#define EV_ARGS c->views(), (const long long*)S.d_region_off.p, (const int*)S.d_order.p, (const long long*)S.d_segrank_off.p, V, N, M,                      \
                (const int2*)S.d_ranges.p, (const int*)S.d_est_best.p,                                                                                        \
                (const double*)S.d_est_P.p, K, (const long long*)S.d_kx.p, (const float*)d_sim.p, (const int*)d_flag.p, (const long long*)CL.d_ptr.p,             \
                (const int*)CL.d_idx.p, two_sigA_sqr,                                                                                                         \
                med_scene_depth_lines, min_affinity
        k_aff_events<false><<<nbs, 128, 0, st>>>(EV_ARGS, (int*)S.d_evcnt.p, nullptr, nullptr, nullptr, nullptr, nullptr);
        size_t tbe = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tbe, (const int*)nullptr, (long long*)nullptr, N + 1, st);
        if ((rc = l3d_reserve(c, S.d_sort_tmp, tbe, "scan temp"))) return rc;
        tbe = S.d_sort_tmp.cap;
        L3D_CUDA(c, cub::DeviceScan::ExclusiveSum(S.d_sort_tmp.p, tbe, (const int*)S.d_evcnt.p, (long long*)S.d_evptr.p, N + 1, st), "event scan");
        long long nev = 0;
        L3D_CUDA(c, cudaMemcpyAsync(&nev, (long long*)S.d_evptr.p + N, 8, cudaMemcpyDeviceToHost, st), "event count");
        L3D_CUDA(c, cudaStreamSynchronize(st), "affinity events");
        c->launches += 4;
        if (nev == 0) return 0;
        if (nev >= (1ll << 31)) return l3d_fail(c, L3D_ERR_UNSUPPORTED, "affinity: more than 2^31 candidate events");
        if ((rc = l3d_reserve(c, S.d_aff_oi, 8 * (size_t)nev, "edges i"))) return rc;
        if ((rc = l3d_reserve(c, S.d_aff_oj, 8 * (size_t)nev, "edges j"))) return rc;
        if ((rc = l3d_reserve(c, S.d_aff_ow, 4 * (size_t)nev, "edges w"))) return rc;
        if ((rc = l3d_reserve(c, S.d_aff_par, 4 * (size_t)nev, "edge parents"))) return rc;
        k_aff_events<true><<<nbs, 128, 0, st>>>(EV_ARGS, nullptr, (const long long*)S.d_evptr.p, (long long*)S.d_aff_oi.p, (long long*)S.d_aff_oj.p,
                                                (float*)S.d_aff_ow.p, (int*)S.d_aff_par.p);
#undef EV_ARGS
        ++c->launches;
        L3D_CUDA(c, cudaGetLastError(), "k_aff_events");
        S.aff_has_parents = true;
        return nev;
    }
    S.aff_has_parents = false;
    size_t tb = 0;
    if (K >= (1ll << 31)) return l3d_fail(c, L3D_ERR_UNSUPPORTED, "affinity: more than 2^31 kept matches");
    cub::DeviceScan::ExclusiveSum(nullptr, tb, (const int*)d_flag.p, (long long*)d_pos.p, (int)K, st);
    if ((rc = l3d_reserve(c, S.d_sort_tmp, tb, "scan temp"))) return rc;
    tb = S.d_sort_tmp.cap;   // scan over the kept matches; the edge count is pos[K-1] + flag[K-1]
    L3D_CUDA(c, cub::DeviceScan::ExclusiveSum(S.d_sort_tmp.p, tb, (const int*)d_flag.p, (long long*)d_pos.p, (int)K, st), "affinity scan");
    long long last_pos = 0; int last_flag = 0;
    L3D_CUDA(c, cudaMemcpyAsync(&last_pos, (long long*)d_pos.p + K - 1, 8, cudaMemcpyDeviceToHost, st), "count");
    L3D_CUDA(c, cudaMemcpyAsync(&last_flag, (int*)d_flag.p + K - 1, 4, cudaMemcpyDeviceToHost, st), "count");
    L3D_CUDA(c, cudaStreamSynchronize(st), "affinity");
    const long long ne = last_pos + last_flag;
    c->launches += 3;
    if (ne == 0) return 0;
    DevBuf &o_i = S.d_aff_oi, &o_j = S.d_aff_oj, &o_w = S.d_aff_ow;
    if ((rc = l3d_reserve(c, o_i, 8 * (size_t)ne, "edges i"))) return rc;
    if ((rc = l3d_reserve(c, o_j, 8 * (size_t)ne, "edges j"))) return rc;
    if ((rc = l3d_reserve(c, o_w, 4 * (size_t)ne, "edges w"))) return rc;
    k_affinity_compact<<<nb, 256, 0, st>>>(K, (const int*)d_flag.p, (const long long*)d_pos.p, (const float*)d_sim.p, (const long long*)d_gi.p,
                                           (const long long*)d_gj.p, (long long*)o_i.p, (long long*)o_j.p, (float*)o_w.p);
    ++c->launches;
    L3D_CUDA(c, cudaGetLastError(), "k_affinity_compact");
    return ne;
}

// Affinity edges between segments with 3D estimates, in the reference's emission order (estimate order, then match
// list order), BEFORE the "unused" de-duplication and local-id assignment (line3D.cc:1881-1900).
// out_gi/out_gj are GLOBAL segment indices (view seg offset + seg id in l3d_set_views order).  Returns the count.
long long l3d_affinity_edges(l3d_ctx* c, float two_sigA_sqr, float med_scene_depth_lines, float min_affinity,
                             long long* out_gi, long long* out_gj, float* out_w, long long cap)
{
    if (!c) return L3D_ERR_INVALID;
    if (!c->sweep.valid) return l3d_fail(c, L3D_ERR_STATE, "l3d_affinity_edges: call l3d_score_sweep first");
    cudaSetDevice(c->device);
    SweepState& S = c->sweep;
    const long long ne = affinity_candidates(c, two_sigA_sqr, med_scene_depth_lines, min_affinity);
    if (ne <= 0 || !out_gi || ne > cap) return ne;
    cudaStream_t st = c->stream;
    L3D_CUDA(c, cudaMemcpyAsync(out_gi, S.d_aff_oi.p, 8 * (size_t)ne, cudaMemcpyDeviceToHost, st), "download edges");
    L3D_CUDA(c, cudaMemcpyAsync(out_gj, S.d_aff_oj.p, 8 * (size_t)ne, cudaMemcpyDeviceToHost, st), "download edges");
    L3D_CUDA(c, cudaMemcpyAsync(out_w, S.d_aff_ow.p, 4 * (size_t)ne, cudaMemcpyDeviceToHost, st), "download edges");
    L3D_CUDA(c, cudaStreamSynchronize(st), "affinity download");
    return ne;
}

// The complete affinity matrix A_ of computingAffinityMatrix (line3D.cc:1852-1979) built on the device: candidates as
// above, then the reference's bookkeeping reproduced with sorts instead of mutex-protected maps:
//   unused(i,j)   (line3D.cc:1982-2002): of all candidates of an unordered segment pair only the FIRST in emission order
//                 survives                      -> stable radix sort by pair key, keep the head of every run;
//   getLocalID    (line3D.cc:2005-2023): ids in order of first appearance (source of an accepted edge before its target)
//                                               -> atomicMin of the appearance time per segment, sort segments by it.
// Output: the CLEdge list in the reference's order, (id1,id2,w),(id2,id1,w) per accepted edge, and local2global as
// global segment indices.  Returns the number of list entries (2 per edge), even if > cap_edges (then nothing is copied).
long long l3d_affinity_matrix(l3d_ctx* c, float two_sigA_sqr, float med_scene_depth_lines, float min_affinity, int* out_i, int* out_j,
                              float* out_w, long long cap_edges, long long* out_local2global, long long cap_ids, long long* num_ids)
{
    if (!c || !num_ids) return L3D_ERR_INVALID;
    if (!c->sweep.valid) return l3d_fail(c, L3D_ERR_STATE, "l3d_affinity_matrix: call l3d_score_sweep first");
    cudaSetDevice(c->device);
    SweepState& S = c->sweep;
    AffinityState& A = c->aff;
    cudaStream_t st = c->stream;
    int rc;
    if (!A.valid || A.two_sigA_sqr != two_sigA_sqr || A.med != med_scene_depth_lines || A.min_aff != min_affinity) {
        A.valid = false;
        const long long ne = affinity_candidates(c, two_sigA_sqr, med_scene_depth_lines, min_affinity);
        if (ne < 0) return ne;
        A.K = 0; A.n_ids = 0;
        if (ne > 0) {
            if (ne >= (1ll << 32) || c->total_segs >= (1ll << 32)) return l3d_fail(c, L3D_ERR_UNSUPPORTED, "l3d_affinity_matrix: more than 2^32 candidates / segments");
            const long long N = c->total_segs;
#define RES(buf, bytes, what) if ((rc = l3d_reserve(c, buf, (size_t)std::max<long long>((long long)(bytes), 16), what))) return rc
            RES(A.d_key, 8 * ne, "pair keys"); RES(A.d_key2, 8 * ne, "pair keys"); RES(A.d_val, 4 * ne, "pair vals"); RES(A.d_val2, 4 * ne, "pair vals");
            RES(A.d_keep, 4 * ne, "keep flags"); RES(A.d_q, 8 * (ne + 1), "edge positions"); RES(A.d_time, 8 * N, "appearance times");
            RES(A.d_nflag, 4 * N, "node flags"); RES(A.d_npos, 8 * (N + 1), "node positions"); RES(A.d_idof, 4 * N, "id of segment");
            const int bseg = bits_i64(N);
            size_t tb1 = 0, tb2 = 0, tb3 = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tb1, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (unsigned int*)nullptr, (unsigned int*)nullptr, (int)ne, 0, 32 + bseg, st);
            cub::DeviceScan::ExclusiveSum(nullptr, tb2, (const int*)nullptr, (long long*)nullptr, std::max(ne, N), st);
            cub::DeviceRadixSort::SortPairs(nullptr, tb3, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (unsigned int*)nullptr, (unsigned int*)nullptr, (int)std::min<long long>(N, 2 * ne), 0, 64, st);
            RES(S.d_sort_tmp, std::max(tb1, std::max(tb2, tb3)), "sort temp");
            const unsigned int nbe = (unsigned int)((ne + 255) / 256), nbn = (unsigned int)((N + 255) / 256);
            k_aff_keys<<<nbe, 256, 0, st>>>(ne, (const long long*)S.d_aff_oi.p, (const long long*)S.d_aff_oj.p, (unsigned long long*)A.d_key.p, (unsigned int*)A.d_val.p);
            size_t tb = S.d_sort_tmp.cap;
            cub::DeviceRadixSort::SortPairs(S.d_sort_tmp.p, tb, (const unsigned long long*)A.d_key.p, (unsigned long long*)A.d_key2.p, (const unsigned int*)A.d_val.p, (unsigned int*)A.d_val2.p, (int)ne, 0, 32 + bseg, st);
            {   // unused(): fixpoint over the conditional events (one round + one confirming round without collinearity links)
                const int* par = S.aff_has_parents ? (const int*)S.d_aff_par.p : nullptr;
                RES(A.d_keep2, 4 * ne, "keep flags"); RES(A.d_found, 4 * N, "found flags"); RES(A.d_changed, 16, "changed flag");
                L3D_CUDA(c, cudaMemsetAsync(A.d_keep.p, 0, 4 * ne, st), "clear keep");
                L3D_CUDA(c, cudaMemsetAsync(A.d_found.p, 0, 4 * N, st), "clear found");
                int* acc_prev = (int*)A.d_keep.p; int* acc_new = (int*)A.d_keep2.p;
                for (int round = 0;; ++round) {
                    if (round > 100000) return l3d_fail(c, L3D_ERR_UNSUPPORTED, "l3d_affinity_matrix: unused() fixpoint did not converge");
                    L3D_CUDA(c, cudaMemsetAsync(A.d_changed.p, 0, 4, st), "clear changed");
                    k_aff_round<<<nbe, 256, 0, st>>>(ne, (const unsigned long long*)A.d_key2.p, (const unsigned int*)A.d_val2.p, par, (const long long*)S.d_aff_oi.p,
                                                     acc_prev, (const int*)A.d_found.p, acc_new, (int*)A.d_changed.p);
                    std::swap(acc_prev, acc_new);
                    ++c->launches;
                    if (!par) break;                 // all events direct: the head of every run, final after one round
                    L3D_CUDA(c, cudaMemsetAsync(A.d_found.p, 0, 4 * N, st), "clear found");
                    k_aff_found<<<nbe, 256, 0, st>>>(ne, par, (const long long*)S.d_aff_oi.p, acc_prev, (int*)A.d_found.p);
                    int changed = 0;
                    L3D_CUDA(c, cudaMemcpyAsync(&changed, A.d_changed.p, 4, cudaMemcpyDeviceToHost, st), "changed flag");
                    L3D_CUDA(c, cudaStreamSynchronize(st), "affinity fixpoint");
                    ++c->launches;
                    if (!changed) break;
                }
                if (acc_prev != (int*)A.d_keep.p) L3D_CUDA(c, cudaMemcpyAsync(A.d_keep.p, acc_prev, 4 * ne, cudaMemcpyDeviceToDevice, st), "keep flags");
            }
            tb = S.d_sort_tmp.cap;
            cub::DeviceScan::ExclusiveSum(S.d_sort_tmp.p, tb, (const int*)A.d_keep.p, (long long*)A.d_q.p, ne, st);
            L3D_CUDA(c, cudaMemsetAsync(A.d_time.p, 0xFF, 8 * N, st), "init times");
            k_aff_time<<<nbe, 256, 0, st>>>(ne, (const int*)A.d_keep.p, (const long long*)A.d_q.p, (const long long*)S.d_aff_oi.p, (const long long*)S.d_aff_oj.p, (unsigned long long*)A.d_time.p);
            k_aff_nodeflags<<<nbn, 256, 0, st>>>(N, (const unsigned long long*)A.d_time.p, (int*)A.d_nflag.p);
            tb = S.d_sort_tmp.cap;
            cub::DeviceScan::ExclusiveSum(S.d_sort_tmp.p, tb, (const int*)A.d_nflag.p, (long long*)A.d_npos.p, N, st);
            long long lq = 0, lp = 0; int lk = 0, lf = 0;
            L3D_CUDA(c, cudaMemcpyAsync(&lq, (long long*)A.d_q.p + ne - 1, 8, cudaMemcpyDeviceToHost, st), "count");
            L3D_CUDA(c, cudaMemcpyAsync(&lk, (int*)A.d_keep.p + ne - 1, 4, cudaMemcpyDeviceToHost, st), "count");
            L3D_CUDA(c, cudaMemcpyAsync(&lp, (long long*)A.d_npos.p + N - 1, 8, cudaMemcpyDeviceToHost, st), "count");
            L3D_CUDA(c, cudaMemcpyAsync(&lf, (int*)A.d_nflag.p + N - 1, 4, cudaMemcpyDeviceToHost, st), "count");
            L3D_CUDA(c, cudaStreamSynchronize(st), "affinity matrix");
            A.K = lq + lk; A.n_ids = lp + lf;
            RES(A.d_nkey, 8 * A.n_ids, "node keys"); RES(A.d_nkey2, 8 * A.n_ids, "node keys"); RES(A.d_nval, 4 * A.n_ids, "node vals"); RES(A.d_nval2, 4 * A.n_ids, "node vals");
            RES(A.d_l2g, 8 * A.n_ids, "local2global"); RES(A.d_ei, 4 * 2 * A.K, "A i"); RES(A.d_ej, 4 * 2 * A.K, "A j"); RES(A.d_ew, 4 * 2 * A.K, "A w");
#undef RES
            k_aff_nodes<<<nbn, 256, 0, st>>>(N, (const int*)A.d_nflag.p, (const long long*)A.d_npos.p, (const unsigned long long*)A.d_time.p, (unsigned long long*)A.d_nkey.p, (unsigned int*)A.d_nval.p);
            tb = S.d_sort_tmp.cap;
            cub::DeviceRadixSort::SortPairs(S.d_sort_tmp.p, tb, (const unsigned long long*)A.d_nkey.p, (unsigned long long*)A.d_nkey2.p, (const unsigned int*)A.d_nval.p, (unsigned int*)A.d_nval2.p, (int)A.n_ids, 0, 64, st);
            k_aff_ids<<<(unsigned int)((A.n_ids + 255) / 256), 256, 0, st>>>(A.n_ids, (const unsigned int*)A.d_nval2.p, (int*)A.d_idof.p, (long long*)A.d_l2g.p);
            k_aff_emit<<<nbe, 256, 0, st>>>(ne, (const int*)A.d_keep.p, (const long long*)A.d_q.p, (const long long*)S.d_aff_oi.p, (const long long*)S.d_aff_oj.p,
                                            (const float*)S.d_aff_ow.p, (const int*)A.d_idof.p, (int*)A.d_ei.p, (int*)A.d_ej.p, (float*)A.d_ew.p);
            c->launches += 8 + 30;
            L3D_CUDA(c, cudaGetLastError(), "affinity matrix kernels");
        }
        A.valid = true; A.two_sigA_sqr = two_sigA_sqr; A.med = med_scene_depth_lines; A.min_aff = min_affinity;
    }
    *num_ids = A.n_ids;
    const long long nE = 2 * A.K;
    if (nE == 0 || !out_i || nE > cap_edges || A.n_ids > cap_ids) return nE;
    L3D_CUDA(c, cudaMemcpyAsync(out_i, A.d_ei.p, 4 * (size_t)nE, cudaMemcpyDeviceToHost, st), "download A");
    L3D_CUDA(c, cudaMemcpyAsync(out_j, A.d_ej.p, 4 * (size_t)nE, cudaMemcpyDeviceToHost, st), "download A");
    L3D_CUDA(c, cudaMemcpyAsync(out_w, A.d_ew.p, 4 * (size_t)nE, cudaMemcpyDeviceToHost, st), "download A");
    if (out_local2global) L3D_CUDA(c, cudaMemcpyAsync(out_local2global, A.d_l2g.p, 8 * (size_t)A.n_ids, cudaMemcpyDeviceToHost, st), "download ids");
    L3D_CUDA(c, cudaStreamSynchronize(st), "affinity matrix download");
    return nE;
}

// replicator_dynamics_diffusion_GPU (cudawrapper.h:80) on a COO edge list (the CLEdge list A_ of line3D.cc:2030).
// out_*: row-sorted COO of the diffused matrix, nnz entries (same order as the reference's downloaded W).
// kernel_ms (optional): device time of the `iters` diffusion iterations only (CUDA events on the context's stream).
// min(w12, w21) of performRDD (line3D.cc:2039-2071): every entry takes the smaller of itself and its transposed entry
__global__ void __launch_bounds__(256) k_rdd_symmetrise(long long nnz, const float* __restrict__ P, const int* __restrict__ tslot, float* __restrict__ out)
{
    const long long y = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= nnz) return;
    const int t = tslot[y];
    out[y] = t >= 0 ? fminf(P[y], P[t]) : P[y];
}

// the diffusion proper on DEVICE edge arrays d_ei/d_ej/d_ew (in_rdd_bufs: they already are R.d_ei/ej/ew); symmetrise: performRDD's
// min(w12, w21) applied before the download.  Output in row-sorted (i, j) order - the order std::map iterates in (line3D.cc:2063-2071).
static int rdd_core(l3d_ctx* c, int n, long long nnz, const int* d_ei, const int* d_ej, const float* d_ew, int iters, bool symmetrise,
                    int* out_i, int* out_j, float* out_w, float* kernel_ms)
{
    RddState& R = c->rdd;
    int rc;
#define RES(buf, bytes, what) if ((rc = l3d_reserve(c, buf, (size_t)(bytes), what))) return rc
    RES(R.d_krow, 8 * nnz, "rdd keys"); RES(R.d_kcol, 8 * nnz, "rdd keys"); RES(R.d_k2, 8 * nnz, "rdd keys"); RES(R.d_idx, 4 * nnz, "rdd idx"); RES(R.d_idx2, 4 * nnz, "rdd idx");
    RES(R.d_P, 4 * nnz, "rdd P"); RES(R.d_Pn, 4 * nnz, "rdd P'"); RES(R.d_W, 4 * nnz, "rdd W");
    RES(R.d_prow, 4 * nnz, "rdd rows"); RES(R.d_pcol, 4 * nnz, "rdd cols"); RES(R.d_wmaj, 4 * nnz, "rdd cols"); RES(R.d_wmin, 4 * nnz, "rdd rows");
    RES(R.d_rowptr, 4 * ((size_t)n + 1), "rdd rowptr"); RES(R.d_colptr, 4 * ((size_t)n + 1), "rdd colptr"); RES(R.d_tslot, 4 * nnz, "rdd tslot");
    size_t sb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sb, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (unsigned int*)nullptr, (unsigned int*)nullptr, (int)nnz, 0, 64, c->stream);
    RES(R.d_tmp, sb, "rdd sort temp");
#undef RES
    cudaStream_t st = c->stream;
    const unsigned int nb = (unsigned int)((nnz + 255) / 256);
    k_rdd_keys<<<nb, 256, 0, st>>>(nnz, d_ei, d_ej, (unsigned long long*)R.d_krow.p, (unsigned long long*)R.d_kcol.p, (unsigned int*)R.d_idx.p);
    size_t tb = R.d_tmp.cap;
    // P: row-sorted (stable, like std::list::sort with sortCLEdgesByRow, sparsematrix.cc:22-25 / 104-107)
    cub::DeviceRadixSort::SortPairs(R.d_tmp.p, tb, (const unsigned long long*)R.d_krow.p, (unsigned long long*)R.d_k2.p, (const unsigned int*)R.d_idx.p, (unsigned int*)R.d_idx2.p, (int)nnz, 0, 64, st);
    k_rdd_gather<<<nb, 256, 0, st>>>(nnz, (const unsigned long long*)R.d_k2.p, (const unsigned int*)R.d_idx2.p, d_ew, (float*)R.d_P.p, (int*)R.d_prow.p, (int*)R.d_pcol.p);
    // W: col-sorted
    tb = R.d_tmp.cap;
    cub::DeviceRadixSort::SortPairs(R.d_tmp.p, tb, (const unsigned long long*)R.d_kcol.p, (unsigned long long*)R.d_k2.p, (const unsigned int*)R.d_idx.p, (unsigned int*)R.d_idx2.p, (int)nnz, 0, 64, st);
    k_rdd_gather<<<nb, 256, 0, st>>>(nnz, (const unsigned long long*)R.d_k2.p, (const unsigned int*)R.d_idx2.p, d_ew, (float*)R.d_W.p, (int*)R.d_wmaj.p, (int*)R.d_wmin.p);
    k_rdd_ptr<<<(n + 256) / 256, 256, 0, st>>>(n, (int)nnz, (const int*)R.d_prow.p, (int*)R.d_rowptr.p);
    k_rdd_ptr<<<(n + 256) / 256, 256, 0, st>>>(n, (int)nnz, (const int*)R.d_wmaj.p, (int*)R.d_colptr.p);
    k_rdd_tslot<<<nb, 256, 0, st>>>(nnz, (const int*)R.d_prow.p, (const int*)R.d_pcol.p, (const int*)R.d_rowptr.p, (int*)R.d_tslot.p);
    // padded value arrays (rows / columns on float4 boundaries)
    if ((rc = l3d_reserve(c, R.d_len4, 4 * ((size_t)n + 1), "rdd len4"))) return rc;
    if ((rc = l3d_reserve(c, R.d_rp4, 4 * ((size_t)n + 1), "rdd rp4"))) return rc;
    if ((rc = l3d_reserve(c, R.d_cp4, 4 * ((size_t)n + 1), "rdd cp4"))) return rc;
    int tot4[2] = {0, 0};
    for (int pass = 0; pass < 2; ++pass) {
        const int* ptr = (const int*)(pass ? R.d_colptr.p : R.d_rowptr.p);
        int* p4 = (int*)(pass ? R.d_cp4.p : R.d_rp4.p);
        L3D_CUDA(c, cudaMemsetAsync(R.d_len4.p, 0, 4 * ((size_t)n + 1), st), "rdd len4");
        k_rdd_len4<<<(n + 255) / 256, 256, 0, st>>>(n, ptr, (int*)R.d_len4.p);
        tb = R.d_tmp.cap;
        size_t need = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, need, (const int*)R.d_len4.p, p4, n + 1, st);
        if (need > tb) { if ((rc = l3d_reserve(c, R.d_tmp, need, "rdd scan temp"))) return rc; tb = R.d_tmp.cap; }
        L3D_CUDA(c, cub::DeviceScan::ExclusiveSum(R.d_tmp.p, tb, (const int*)R.d_len4.p, p4, n + 1, st), "rdd scan");
        L3D_CUDA(c, cudaMemcpyAsync(&tot4[pass], p4 + n, 4, cudaMemcpyDeviceToHost, st), "rdd padded size");
    }
    L3D_CUDA(c, cudaStreamSynchronize(st), "rdd padded size");
    const size_t pbytes = 16 * (size_t)std::max(tot4[0], 1), wbytes = 16 * (size_t)std::max(tot4[1], 1);
    if ((rc = l3d_reserve(c, R.d_Pp, pbytes, "rdd P padded"))) return rc;
    if ((rc = l3d_reserve(c, R.d_Pnp, pbytes, "rdd P' padded"))) return rc;
    if ((rc = l3d_reserve(c, R.d_Wp, wbytes, "rdd W padded"))) return rc;
    L3D_CUDA(c, cudaMemsetAsync(R.d_Pp.p, 0, pbytes, st), "rdd pad"); L3D_CUDA(c, cudaMemsetAsync(R.d_Wp.p, 0, wbytes, st), "rdd pad");
    k_rdd_pad<<<nb, 256, 0, st>>>(nnz, (const int*)R.d_prow.p, (const int*)R.d_rowptr.p, (const int*)R.d_rp4.p, (const float*)R.d_P.p, (float*)R.d_Pp.p);
    k_rdd_pad<<<nb, 256, 0, st>>>(nnz, (const int*)R.d_wmaj.p, (const int*)R.d_colptr.p, (const int*)R.d_cp4.p, (const float*)R.d_W.p, (float*)R.d_Wp.p);
    if (4ll * std::max(tot4[0], tot4[1]) >= (1ll << 31)) return l3d_fail(c, L3D_ERR_UNSUPPORTED, "l3d_rdd: padded matrix too large for 32-bit slots");
    if ((rc = l3d_reserve(c, R.d_plan, 16 * (size_t)nnz, "rdd plan"))) return rc;
    if ((rc = l3d_reserve(c, R.d_dst, 4 * (size_t)nnz, "rdd dst"))) return rc;
    k_rdd_plan<<<nb, 256, 0, st>>>(nnz, (const int*)R.d_prow.p, (const int*)R.d_pcol.p, (const int*)R.d_rowptr.p, (const int*)R.d_colptr.p,
                                   (const int*)R.d_rp4.p, (const int*)R.d_cp4.p, (const int*)R.d_tslot.p, (int4*)R.d_plan.p, (int*)R.d_dst.p);
    // P' starts as a copy of the un-normalised P (cudawrapper.cu:724), then P is row-normalised (727)
    L3D_CUDA(c, cudaMemcpyAsync(R.d_Pnp.p, R.d_Pp.p, pbytes, cudaMemcpyDeviceToDevice, st), "rdd copy");
    const unsigned int nbn = (unsigned int)((4ll * n + 255) / 256);
    auto normalize = [&](float* Pbuf) { k_rdd_normalize_q<<<nbn, 256, 0, st>>>(n, (const int*)R.d_rowptr.p, (const int*)R.d_rp4.p, Pbuf); };
    normalize((float*)R.d_Pp.p);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (kernel_ms) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
    float* P = (float*)R.d_Pp.p; float* Pn = (float*)R.d_Pnp.p;
    for (int it = 0; it < iters; ++it) {
        k_rdd_step8<<<(unsigned int)((nnz + 255) / 256), 256, 0, st>>>(nnz, (const int4*)R.d_plan.p, (const int*)R.d_dst.p, P, (const float*)R.d_Wp.p, Pn);
        std::swap(P, Pn);
        if (it < iters - 1) normalize(P);       // no normalisation after the last step (cudawrapper.cu:751)
    }
    if (kernel_ms) cudaEventRecord(e1, st);
    k_rdd_unpad<<<nb, 256, 0, st>>>(nnz, (const int*)R.d_prow.p, (const int*)R.d_rowptr.p, (const int*)R.d_rp4.p, P, (float*)R.d_P.p);
    const float* result = (const float*)R.d_P.p;
    if (symmetrise) {
        k_rdd_symmetrise<<<nb, 256, 0, st>>>(nnz, (const float*)R.d_P.p, (const int*)R.d_tslot.p, (float*)R.d_Pn.p);
        result = (const float*)R.d_Pn.p;
        ++c->launches;
    }
    c->launches += 9 + 16 + 10 + 2 * iters;
    L3D_CUDA(c, cudaGetLastError(), "rdd kernels");
    L3D_CUDA(c, cudaMemcpyAsync(out_w, result, 4 * nnz, cudaMemcpyDeviceToHost, st), "rdd download");
    L3D_CUDA(c, cudaMemcpyAsync(out_i, R.d_prow.p, 4 * nnz, cudaMemcpyDeviceToHost, st), "rdd download");
    L3D_CUDA(c, cudaMemcpyAsync(out_j, R.d_pcol.p, 4 * nnz, cudaMemcpyDeviceToHost, st), "rdd download");
    L3D_CUDA(c, cudaStreamSynchronize(st), "rdd");
    if (kernel_ms) { cudaEventElapsedTime(kernel_ms, e0, e1); cudaEventDestroy(e0); cudaEventDestroy(e1); }
    return L3D_OK;
}

int l3d_rdd(l3d_ctx* c, int n, long long nnz, const int* ei, const int* ej, const float* ew, int iters, int* out_i, int* out_j,
            float* out_w, float* kernel_ms)
{
    if (!c || n <= 0 || nnz < 0 || (nnz && (!ei || !ej || !ew || !out_i || !out_j || !out_w))) return l3d_fail(c, L3D_ERR_INVALID, "l3d_rdd: bad arguments");
    if (nnz == 0) return L3D_OK;
    if (nnz >= (1ll << 31) - 2) return l3d_fail(c, L3D_ERR_UNSUPPORTED, "l3d_rdd: more than 2^31 entries");
    cudaSetDevice(c->device);
    RddState& R = c->rdd;
    int rc;
    if ((rc = l3d_reserve(c, R.d_ei, 4 * (size_t)nnz, "rdd i")) || (rc = l3d_reserve(c, R.d_ej, 4 * (size_t)nnz, "rdd j")) || (rc = l3d_reserve(c, R.d_ew, 4 * (size_t)nnz, "rdd w"))) return rc;
    L3D_CUDA(c, cudaMemcpyAsync(R.d_ei.p, ei, 4 * nnz, cudaMemcpyHostToDevice, c->stream), "rdd upload");
    L3D_CUDA(c, cudaMemcpyAsync(R.d_ej.p, ej, 4 * nnz, cudaMemcpyHostToDevice, c->stream), "rdd upload");
    L3D_CUDA(c, cudaMemcpyAsync(R.d_ew.p, ew, 4 * nnz, cudaMemcpyHostToDevice, c->stream), "rdd upload");
    return rdd_core(c, n, nnz, (const int*)R.d_ei.p, (const int*)R.d_ej.p, (const float*)R.d_ew.p, iters, false, out_i, out_j, out_w, kernel_ms);
}

// performRDD (line3D.cc:2026-2076) on the affinity matrix l3d_affinity_matrix left on the device: no download / re-upload of A_, the
// min(w12, w21) symmetrisation on the device, the result in the (i, j) order the reference rebuilds A_ in.  out arrays: 2 * K entries
// (the count l3d_affinity_matrix returned).
long long l3d_rdd_affinity(l3d_ctx* c, int iters, int* out_i, int* out_j, float* out_w, long long cap)
{
    if (!c || !out_i || !out_j || !out_w) return L3D_ERR_INVALID;
    AffinityState& A = c->aff;
    if (!A.valid) return l3d_fail(c, L3D_ERR_STATE, "l3d_rdd_affinity: call l3d_affinity_matrix first");
    const long long nnz = 2 * A.K;
    if (nnz == 0 || nnz > cap) return nnz;
    if (nnz >= (1ll << 31) - 2) return l3d_fail(c, L3D_ERR_UNSUPPORTED, "l3d_rdd_affinity: more than 2^31 entries");
    cudaSetDevice(c->device);
    const int rc = rdd_core(c, (int)A.n_ids, nnz, (const int*)A.d_ei.p, (const int*)A.d_ej.p, (const float*)A.d_ew.p, iters, true, out_i, out_j, out_w, nullptr);
    return rc < 0 ? rc : nnz;
}

// Stable ascending argsort of float keys on the device (the edge-weight sort of performClustering, clustering.cc:13-14:
// std::list::sort by weight is stable).  perm_out[i] = index of the i-th smallest key; equal keys keep their input order.
int l3d_argsort_f32(l3d_ctx* c, long long n, const float* keys, unsigned int* perm_out)
{
    if (!c || n < 0 || (n && (!keys || !perm_out))) return l3d_fail(c, L3D_ERR_INVALID, "l3d_argsort_f32: bad arguments");
    if (n == 0) return L3D_OK;
    if (n >= (1ll << 31)) return l3d_fail(c, L3D_ERR_UNSUPPORTED, "l3d_argsort_f32: more than 2^31 keys");
    cudaSetDevice(c->device);
    RddState& R = c->rdd;
    int rc;
    if ((rc = l3d_reserve(c, R.d_ew, 4 * (size_t)n, "sort keys"))) return rc;
    if ((rc = l3d_reserve(c, R.d_ei, 4 * (size_t)n, "sort keys"))) return rc;
    if ((rc = l3d_reserve(c, R.d_idx, 4 * (size_t)n, "sort idx"))) return rc;
    if ((rc = l3d_reserve(c, R.d_idx2, 4 * (size_t)n, "sort idx"))) return rc;
    size_t sb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sb, (float*)nullptr, (float*)nullptr, (unsigned int*)nullptr, (unsigned int*)nullptr, (int)n, 0, 32, c->stream);
    if ((rc = l3d_reserve(c, R.d_tmp, sb, "sort temp"))) return rc;
    cudaStream_t st = c->stream;
    L3D_CUDA(c, cudaMemcpyAsync(R.d_ew.p, keys, 4 * (size_t)n, cudaMemcpyHostToDevice, st), "sort upload");
    k_iota<<<(unsigned int)((n + 255) / 256), 256, 0, st>>>(n, (unsigned int*)R.d_idx.p);
    size_t tb = R.d_tmp.cap;
    L3D_CUDA(c, cub::DeviceRadixSort::SortPairs(R.d_tmp.p, tb, (const float*)R.d_ew.p, (float*)R.d_ei.p, (const unsigned int*)R.d_idx.p, (unsigned int*)R.d_idx2.p, (int)n, 0, 32, st), "argsort");
    c->launches += 6;
    L3D_CUDA(c, cudaMemcpyAsync(perm_out, R.d_idx2.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, st), "sort download");
    L3D_CUDA(c, cudaStreamSynchronize(st), "argsort");
    return L3D_OK;
}

} // extern "C"
