// l3d_pipeline.cu — the per-view scoring sweep of Line3D::computeMatches (line3D.cc:702-778) on the device.
//
// After l3d_match_pairs has produced the kNN match records of every view pair in one launch, the reference walks the
// views in ascending camID order and, for each view: filters its matches by orientation (checkMatchOrientation,
// line3D.cc:811-858), sorts them by (tgt cam, tgt seg) and flattens them (scoringGPU, 1311-1355), scores them
// (K_score_matches, cudawrapper.cu:256-367), hands the positively scored ones to the not-yet-processed target views as
// inverse matches (storeInverseMatches, 1672-1699) and keeps the good ones + the best 3D estimate per segment
// (filterMatches, 1586-1669).  The order dependence (a view sees the inverse matches of its earlier neighbours) is kept:
// views are processed sequentially, but every step is a kernel over all matches / segments of the view and nothing
// returns to the host in between.
//
// Arithmetic contract as in l3d_device.cuh (-fmad=false): the float parts repeat K_score_matches operation by
// operation (same libdevice expf/acosf), the double parts repeat the host code of view.cc / line3D.cc as restated in
// oracle/l3d_oracle.cc (IEEE double, no contraction).
#include "l3d_ctx.cuh"
#include "l3d_device_f64.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

#define L3D_PI_1_32_F 0.098174771f
#define L3D_PI_31_32_F 3.043417886f
#define L3D_PI_D 3.14159265358979323846

// View::unprojectSegment + Segment3D ctor (view.cc:356-371, segment3D.h:48-66): degenerate segments collapse to zero
struct DSeg { D3 P1, P2, dir; float length; };
__device__ __forceinline__ DSeg dunproject(const L3DViewDev* v, float4 s, float d1, float d2)
{
    D3 C = d3(v->C_d[0], v->C_d[1], v->C_d[2]);
    D3 r1 = dray(v->RtKinv_d, (double)s.x, (double)s.y), r2 = dray(v->RtKinv_d, (double)s.z, (double)s.w);
    D3 P1 = d3(C.x + r1.x * (double)d1, C.y + r1.y * (double)d1, C.z + r1.z * (double)d1);
    D3 P2 = d3(C.x + r2.x * (double)d2, C.y + r2.y * (double)d2, C.z + r2.z * (double)d2);
    DSeg o;
    o.length = (float)dnorm(dsub(P1, P2));
    if (o.length > L3D_EPS_D) { o.P1 = P1; o.P2 = P2; o.dir = dnormalized(dsub(P2, P1)); }
    else { o.P1 = o.P2 = o.dir = d3(0, 0, 0); o.length = 0.0f; }
    return o;
}

// sort key of a candidate match: (segment | rank of the target camera | target segment), packed as tightly as the
// problem allows so that the radix sort touches few bits; bit `end_bit-1` is reserved for "invalid" (all ones)
// REF_CPU semantics (mode 1): scoringCPU does not sort (only scoringGPU calls sortMatches, line3D.cc:1311), so the list
// order of a segment is the order the matches were appended: first the inverse matches stored by the earlier views (in
// those views' processing and list order = their global slot in the match store), then the direct matches in pair / kNN
// order (= record index).  Key = (segment | direct? | origin index).
struct KeyBits { int seg_shift, cam_shift, end_bit; unsigned long long cam_mask, tgt_mask; int mode; };
static int bits_for(long long n) { int b = 1; while ((1ll << b) < n) ++b; return b; }

// ---------------------------------------------------------------------------------------------- kernels
// G1: enumerate the candidate matches of view v (direct records of pairs with src == v, inverse records of pairs with
//     tgt == v whose src was processed earlier and scored them > 0), apply the orientation check, emit sort keys.
__global__ void __launch_bounds__(256)
k_gather(const float4* __restrict__ segs, const L3DViewDev* __restrict__ views, const L3DPairDev* __restrict__ pairs,
         const int* __restrict__ counts, const l3d_match_rec* __restrict__ recs, const float* __restrict__ slot_score,
         const int4* __restrict__ work, int nwork, int v, int knn, const int* __restrict__ cam_rank,
         unsigned long long* __restrict__ keys, unsigned int* __restrict__ vals, int U, int* __restrict__ Mcount, KeyBits kb,
         const int* __restrict__ slot_pos, const long long* __restrict__ region_of_view)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = false;
    unsigned long long key = ~0ull;
    unsigned int val = 0u;
    if (t < U) {
        int w = 0;                                   // work item: (pair, inverse?, first candidate index, -)
        while (w + 1 < nwork && work[w + 1].z <= t) ++w;
        const int4 wk = work[w];
        const L3DPairDev* P = pairs + wk.x;
        const int local = t - wk.z;
        const int r = local / knn, i = local - r * knn;
        const long long row = P->row_off + r;
        if (i < counts[row]) {
            const long long g = row * knn + i;
            const l3d_match_rec rec = recs[g];
            int seg, tgt_view, tgt_seg; float d1, d2;
            bool ok = true;
            if (!wk.y) { seg = r; tgt_view = P->tgt; tgt_seg = (int)rec.tgt_seg; d1 = rec.d_p1; d2 = rec.d_p2; }
            else { ok = slot_score[g] > 0.0f; seg = (int)rec.tgt_seg; tgt_view = P->src; tgt_seg = r; d1 = rec.d_q1; d2 = rec.d_q2; }
            if (ok) {
                const L3DViewDev* V = views + v;
                const float4 s = segs[V->seg_off + seg];
                DSeg S3 = dunproject(V, s, d1, d2);                                  // unprojectMatch (line3D.cc:1556)
                double px = 0.5 * ((double)s.x + (double)s.z), py = 0.5 * ((double)s.y + (double)s.w);
                D3 r1 = dray(V->RtKinv_d, px, py);                                   // segmentQualityAngle (view.cc:466-484)
                double ang = acos(fmin(fmax(ddot(r1, S3.dir), -1.0), 1.0));
                if (ang > (double)L3D_PI_1_32_F && ang < (double)L3D_PI_31_32_F) {
                    valid = true;
                    if (kb.mode == 0)
                        key = ((unsigned long long)seg << kb.seg_shift) | ((unsigned long long)cam_rank[tgt_view] << kb.cam_shift) | (unsigned long long)tgt_seg;
                    else if (!wk.y) key = ((unsigned long long)seg << kb.seg_shift) | (1ull << kb.cam_shift) | (unsigned long long)g;
                    else key = ((unsigned long long)seg << kb.seg_shift) | (unsigned long long)(region_of_view[tgt_view] + slot_pos[g]);
                    val = (unsigned int)g | (wk.y ? 0x80000000u : 0u);
                }
            }
        }
        keys[t] = key; vals[t] = val;
    }
    unsigned int b = __ballot_sync(0xffffffffu, valid);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(Mcount, __popc(b));
}

// G2: per sorted match: decode, fetch payload, target regulariser (double), unprojected 3D direction (float), ranges.
__global__ void __launch_bounds__(256)
k_build(const float4* __restrict__ segs, const float4* __restrict__ cache, const L3DViewDev* __restrict__ views,
        const L3DPairDev* __restrict__ pairs, const l3d_match_rec* __restrict__ recs,
        int v, int knn, const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ vals,
        const int* __restrict__ Mcount, const int* __restrict__ view_of_camrank,
        int4* __restrict__ m_meta, float4* __restrict__ m_dep, float2* __restrict__ m_os, float2* __restrict__ m_reg,
        float4* __restrict__ m_dir, int2* __restrict__ ranges, KeyBits kb, int num_pairs, double4* __restrict__ m_dir64)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int M = *Mcount;
    if (x >= M) return;
    const unsigned long long key = keys[x];
    const int seg = (int)(key >> kb.seg_shift);
    const unsigned int val = vals[x];
    const bool inv = (val & 0x80000000u) != 0u;
    const l3d_match_rec rec = recs[val & 0x7FFFFFFFu];
    int tgt_view, tgt_seg;
    if (kb.mode == 0) { tgt_view = view_of_camrank[(int)((key >> kb.cam_shift) & kb.cam_mask)]; tgt_seg = (int)(key & kb.tgt_mask); }
    else {      // the key carries the list position, not the target: find the record's pair by its row
        const long long row = (long long)(val & 0x7FFFFFFFu) / knn;
        int lo = 0, hi = num_pairs - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (pairs[mid].row_off <= row) lo = mid; else hi = mid - 1; }
        tgt_view = inv ? pairs[lo].src : pairs[lo].tgt;
        tgt_seg = inv ? (int)(row - pairs[lo].row_off) : (int)rec.tgt_seg;
    }
    float4 dep = inv ? make_float4(rec.d_q1, rec.d_q2, rec.d_p1, rec.d_p2) : make_float4(rec.d_p1, rec.d_p2, rec.d_q1, rec.d_q2);
    m_meta[x] = make_int4(seg, tgt_view, tgt_seg, (int)val);
    m_dep[x] = dep;
    m_os[x] = make_float2(rec.overlap, 0.0f);
    const L3DViewDev* V = views + v;
    const L3DViewDev* T = views + tgt_view;
    const float4 s = segs[V->seg_off + seg];
    {   // regularizers_tgt (line3D.cc:1350-1352) == View::regularizerFrom3Dpoint (view.cc:445-448): |P - C_tgt| * k_tgt in double, stored as float
        DSeg S3 = dunproject(V, s, dep.x, dep.y);
        D3 Ct = d3(T->C_d[0], T->C_d[1], T->C_d[2]);
        m_reg[x] = make_float2((float)(dnorm(dsub(S3.P1, Ct)) * (double)T->k), (float)(dnorm(dsub(S3.P2, Ct)) * (double)T->k));
        if (kb.mode) m_dir64[x] = make_double4(S3.dir.x, S3.dir.y, S3.dir.z, (double)S3.length);     // scoringCPU works on the double 3D segment
    }
    if (kb.mode == 0) {   // D_unproject x2 + D_line_direction_3D (cudawrapper.cu:167-171, 40-43) with the cached float rays
        SegRays R = load_rays(cache, V->seg_off + seg);
        float3 C = make_float3(V->C[0], V->C[1], V->C[2]);
        float3 P1 = make_float3(C.x + dep.x * R.r1.x, C.y + dep.x * R.r1.y, C.z + dep.x * R.r1.z);
        float3 P2 = make_float3(C.x + dep.y * R.r2.x, C.y + dep.y * R.r2.y, C.z + dep.y * R.r2.z);
        float3 dir = normalize3(make_float3(P2.x - P1.x, P2.y - P1.y, P2.z - P1.z));
        m_dir[x] = make_float4(dir.x, dir.y, dir.z, 0.f);
    }
    if (x == 0 || (int)(keys[x - 1] >> kb.seg_shift) != seg) ranges[seg].x = x;
    if (x == M - 1 || (int)(keys[x + 1] >> kb.seg_shift) != seg) ranges[seg].y = x;
}

// G3: K_score_matches (cudawrapper.cu:256-367), one thread per match, same operation order.
__global__ void __launch_bounds__(128)
k_score(const L3DViewDev* __restrict__ views, int v, const int* __restrict__ Mcount, const int4* __restrict__ m_meta,
        const float4* __restrict__ m_dep, const float2* __restrict__ m_reg, const float4* __restrict__ m_dir,
        const int2* __restrict__ ranges, float angle_reg, float sim_t, float q_thr, float cos_thr, float2* __restrict__ m_os)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= *Mcount) return;
    const float k = views[v].k;
    const int4 me = m_meta[x];
    const int tgt_cam_src = me.y;
    const float4 dep = m_dep[x];
    const float d1_src = dep.x, d2_src = dep.y;
    const float4 ds = m_dir[x];
    const float3 dir_src = make_float3(ds.x, ds.y, ds.z);
    float sig1 = k * d1_src, sig2 = k * d2_src;
    float pos_reg1 = 2.0f * sig1 * sig1, pos_reg2 = 2.0f * sig2 * sig2;
    const float2 rg = m_reg[x];
    float pos_reg1_tgt = 2.0f * rg.x * rg.x, pos_reg2_tgt = 2.0f * rg.y * rg.y;
    pos_reg1 = 0.5f * (pos_reg1 + pos_reg1_tgt);
    pos_reg2 = 0.5f * (pos_reg2 + pos_reg2_tgt);
    const int2 rng = ranges[me.x];
    float score3D = 0.0f, current_max_sim = 0.0f;
    int current_cam = -1;
    for (int i = rng.x; i <= rng.y; ++i) {
        const int tgt_cam_tgt = m_meta[i].y;
        if (tgt_cam_src != tgt_cam_tgt) {
            const float4 d2 = m_dep[i];
            const float4 dt = m_dir[i];
            const float dp = dir_src.x * dt.x + dir_src.y * dt.y + dir_src.z * dt.z;
            const float e1 = d1_src - d2.x, e2 = d2_src - d2.y;
            float sim;
            // Exact shortcut: sim = min(three terms), then truncated to 0 below sim_t.  If ONE term is certainly below
            // sim_t the result is 0 whatever the others are (fminf ignores NaN).  exp(-q) < sim_t is certain when
            // q > q_thr = 1.01 * -ln(sim_t) (1 % margin >> the rounding of the division and of expf), and the angular
            // term is certainly below sim_t when |cos| < cos_thr (same margin on the angle).  Everything inside the
            // margins takes the full, reference-order path.
            if (e1 * e1 > q_thr * pos_reg1 || e2 * e2 > q_thr * pos_reg2 || fabsf(dp) < cos_thr) sim = 0.0f;
            else {
                // D_undirected_angle_3D_DEG (cudawrapper.cu:46-53): float acos, DOUBLE divide by pi, times 180, back to float
                float angle = (float)((double)acosf(fmaxf(fminf(dp, 1.0f), -1.0f)) / L3D_PI_D * (double)180.0f);
                if (angle > 90.0f) angle = 180.0f - angle;
                float sim_a = expf(-angle * angle / angle_reg);
                float sim_p1 = expf(-e1 * e1 / pos_reg1), sim_p2 = expf(-e2 * e2 / pos_reg2);
                sim = fminf(sim_a, fminf(sim_p1, sim_p2));
                if (sim < sim_t) sim = 0.0f;
            }
            current_max_sim = fmaxf(current_max_sim, sim);
            if (current_cam != tgt_cam_tgt) { score3D += current_max_sim; current_max_sim = 0.0f; current_cam = tgt_cam_tgt; }
        }
    }
    score3D += current_max_sim;
    m_os[x].y = score3D;
}

// G3 (REF_CPU): scoringCPU (line3D.cc:1208-1294) + similarityForScoring (1417-1446) + angleBetweenSeg3D (1571-1583), one
// thread per match.  The score is the sum over target cameras of the best similarity among that camera's matches, built
// with the reference's running update (add the first value of a camera, replace it when a larger one arrives) in list order.
#define SC_MAP 48
__global__ void __launch_bounds__(128)
k_score_cpu(const L3DViewDev* __restrict__ views, int v, const int* __restrict__ Mcount, const int4* __restrict__ m_meta,
            const float4* __restrict__ m_dep, const float2* __restrict__ m_reg, const double4* __restrict__ m_dir64,
            const int2* __restrict__ ranges, float angle_reg, float sim_t, float2* __restrict__ m_os)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= *Mcount) return;
    const float k = views[v].k;
    const int4 me = m_meta[x];
    const float4 dep = m_dep[x];
    const double4 D1 = m_dir64[x];
    const float sig1 = dep.x * k, sig2 = dep.y * k;
    float reg1 = 2.0f * sig1 * sig1, reg2 = 2.0f * sig2 * sig2;
    const float2 rg = m_reg[x];
    reg1 = 0.5f * (reg1 + 2.0f * rg.x * rg.x); reg2 = 0.5f * (reg2 + 2.0f * rg.y * rg.y);
    const int2 rng = ranges[me.x];
    auto sim_of = [&](int i) -> float {
        const double4 D2 = m_dir64[i];
        if ((float)D1.w < L3D_EPS_D || (float)D2.w < L3D_EPS_D) return 0.0f;
        const float dot_p = (float)(D1.x * D2.x + D1.y * D2.y + D1.z * D2.z);
        float angle = (float)((double)acosf(fmaxf(fminf(dot_p, 1.0f), -1.0f)) / L3D_PI_D * (double)180.0f);
        if (angle > 90.0f) angle = 180.0f - angle;
        const float sim_a = expf(-angle * angle / angle_reg);
        const float4 d2 = m_dep[i];
        const float e1 = dep.x - d2.x, e2 = dep.y - d2.y;
        const float sim_p = fminf(expf(-e1 * e1 / reg1), expf(-e2 * e2 / reg2));
        const float sm = fminf(sim_a, sim_p);
        return sm > sim_t ? sm : 0.0f;
    };
    int cams[SC_MAP]; float best[SC_MAP]; int ncam = 0;
    float score3D = 0.0f;
    for (int i = rng.x; i <= rng.y; ++i) {
        const int cam = m_meta[i].y;
        if (cam == me.y) continue;
        const float sim = sim_of(i);
        int slot = -1;
        for (int j = 0; j < ncam; ++j) if (cams[j] == cam) { slot = j; break; }
        if (slot >= 0) { if (sim > best[slot]) { score3D -= best[slot]; score3D += sim; best[slot] = sim; } }
        else if (ncam < SC_MAP) { score3D += sim; cams[ncam] = cam; best[ncam] = sim; ++ncam; }
        else {      // more target cameras than map slots: recover this camera's running maximum from the earlier entries
            bool seen = false; float cur = 0.0f;
            for (int j = rng.x; j < i; ++j) if (m_meta[j].y == cam) { const float sj = sim_of(j); if (!seen) { cur = sj; seen = true; } else if (sj > cur) cur = sj; }
            if (seen) { if (sim > cur) { score3D -= cur; score3D += sim; } } else score3D += sim;
        }
    }
    m_os[x].y = score3D;
}

// G4: publish the scores of direct matches to their record slots (read later by the target views as inverse matches)
//     and reduce the view's maximum score.
__global__ void __launch_bounds__(256)
k_post_score(const int* __restrict__ Mcount, const int4* __restrict__ m_meta, const float2* __restrict__ m_os,
             float* __restrict__ slot_score, int* __restrict__ view_max_bits, int* __restrict__ slot_pos /* REF_CPU only */)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    float s = 0.0f;
    if (x < *Mcount) {
        s = m_os[x].y;
        const unsigned int val = (unsigned int)m_meta[x].w;
        if (!(val & 0x80000000u)) { slot_score[val] = s; if (slot_pos) slot_pos[val] = x; }
    }
    s = fmaxf(s, 0.0f);
    for (int o = 16; o; o >>= 1) s = fmaxf(s, __shfl_xor_sync(0xffffffffu, s, o));
    if ((threadIdx.x & 31) == 0 && s > 0.0f) atomicMax(view_max_bits, __float_as_int(s));
}

// G5: filterMatches (line3D.cc:1586-1669): one thread per segment of the view.
__global__ void __launch_bounds__(256)
k_filter(const float4* __restrict__ segs, const L3DViewDev* __restrict__ views, int v, const int2* __restrict__ ranges,
         const int4* __restrict__ m_meta, const float4* __restrict__ m_dep, float2* __restrict__ m_os,
         const int* __restrict__ view_max_bits, float min_best, float perc, unsigned char* __restrict__ kept,
         int* __restrict__ est_best /*per global seg: index of best match in the view region or -1*/, double* __restrict__ est_P)
{
    const L3DViewDev* V = views + v;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= V->nseg) return;
    const long long gs = V->seg_off + s;
    const int2 rng = ranges[s];
    int best = -1;
    if (rng.x >= 0) {
        const float score_lim = perc * __int_as_float(*view_max_bits);
        float best_score = 0.0f;
        for (int i = rng.x; i <= rng.y; ++i) {
            const float sc = m_os[i].y;
            const bool keep = sc > 0.0f && sc > score_lim;
            kept[i] = keep ? 1 : 0;
            if (keep && sc > best_score) { best_score = sc; best = i; }
        }
        if (!(best_score > min_best)) {
            best = -1;
            for (int i = rng.x; i <= rng.y; ++i) kept[i] = 0;
        }
    }
    est_best[gs] = best;
    if (best >= 0) {
        const float4 dep = m_dep[best];
        DSeg S3 = dunproject(V, segs[gs], dep.x, dep.y);       // unprojectMatch(best_match, true)
        double* o = est_P + 6 * gs;
        o[0] = S3.P1.x; o[1] = S3.P1.y; o[2] = S3.P1.z; o[3] = S3.P2.x; o[4] = S3.P2.y; o[5] = S3.P2.z;
    }
}

// compact list of the segments that have a 3D estimate: their best match (L3DPP::Match layout) and P1,P2
struct HasEstimate { __host__ __device__ long long operator()(int best) const { return best >= 0 ? 1ll : 0ll; } };
__global__ void __launch_bounds__(256)
k_collect_estimates(const L3DViewDev* __restrict__ views, int V, long long N, const long long* __restrict__ reg_of_view,
                    const int* __restrict__ est_best, const long long* __restrict__ pos, const int4* __restrict__ m_meta,
                    const float4* __restrict__ m_dep, const float2* __restrict__ m_os, const double* __restrict__ est_P,
                    l3d_match* __restrict__ out_best, double* __restrict__ out_P)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    const int b = est_best[g];
    if (b < 0) return;
    int lo = 0, hi = V - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (views[mid].seg_off <= g) lo = mid; else hi = mid - 1; }
    const long long x = reg_of_view[lo] + b;
    const int4 me = m_meta[x]; const float4 dep = m_dep[x]; const float2 os = m_os[x];
    l3d_match m;
    m.src_cam = views[lo].cam_id; m.src_seg = (unsigned int)(g - views[lo].seg_off); m.tgt_cam = views[me.y].cam_id; m.tgt_seg = (unsigned int)me.z;
    m.overlap = os.x; m.score3D = os.y; m.d_p1 = dep.x; m.d_p2 = dep.y; m.d_q1 = dep.z; m.d_q2 = dep.w;
    const long long p = pos[g];
    out_best[p] = m;
    for (int i = 0; i < 6; ++i) out_P[6 * p + i] = est_P[6 * g + i];
}

// ---------------------------------------------------------------------------------------------- host driver
extern "C" {

int l3d_score_sweep(l3d_ctx* c, float two_sigA_sqr, float min_similarity, float min_best_score, float min_best_perc)
{
    if (!c) return L3D_ERR_INVALID;
    if (!c->have_matches) return l3d_fail(c, L3D_ERR_STATE, "l3d_score_sweep: call l3d_match_pairs first");
    cudaSetDevice(c->device);
    const int V = c->num_views, NP = c->num_pairs, knn = c->knn;
    if (V >= 65536) return l3d_fail(c, L3D_ERR_UNSUPPORTED, "l3d_score_sweep: more than 65535 views");
    if (c->total_rows * (long long)knn >= (1ll << 31)) return l3d_fail(c, L3D_ERR_UNSUPPORTED, "l3d_score_sweep: more than 2^31 match slots");
    SweepState& S = c->sweep;
    S.valid = false; c->aff.valid = false;
    // processing order = ascending camID (std::map iteration, line3D.cc:704)
    S.order.resize(V);
    std::iota(S.order.begin(), S.order.end(), 0);
    std::stable_sort(S.order.begin(), S.order.end(), [&](int a, int b) { return c->h_views[a].cam_id < c->h_views[b].cam_id; });
    std::vector<int> rank(V), view_of_rank(V);
    for (int i = 0; i < V; ++i) { rank[S.order[i]] = i; view_of_rank[i] = S.order[i]; }
    // per-view work items
    std::vector<std::vector<int4> > work(V);
    S.U.assign(V, 0); S.region_off.assign(V + 1, 0);
    for (int p = 0; p < NP; ++p) {
        const L3DPairDev& P = c->h_pairs[p];
        const int Ns = c->h_views[P.src].nseg;
        work[P.src].push_back(make_int4(p, 0, S.U[P.src], 0));
        S.U[P.src] += Ns * knn;
        if (rank[P.tgt] > rank[P.src]) {      // tgt still unprocessed when src stores its inverse matches (line3D.cc:1680)
            work[P.tgt].push_back(make_int4(p, 1, S.U[P.tgt], 0));
            S.U[P.tgt] += Ns * knn;
        }
    }
    long long Umax = 1, total = 0; int wmax = 1, nseg_max = 1;
    for (int i = 0; i < V; ++i) {
        const int v = S.order[i];
        S.region_off[i] = total; total += S.U[v];
        Umax = std::max<long long>(Umax, S.U[v]); wmax = std::max<int>(wmax, (int)work[v].size());
        nseg_max = std::max(nseg_max, c->h_views[v].nseg);
    }
    S.region_off[V] = total; S.total = total;
    const long long slots = c->total_rows * knn;
    const bool cpu_sem = c->semantics == L3D_SEM_REF_CPU;
    KeyBits kb;
    if (!cpu_sem) {
        const int bt = bits_for(nseg_max), bc = bits_for(V);
        kb.cam_shift = bt; kb.seg_shift = bt + bc; kb.end_bit = bt + bc + bt + 1;
        kb.cam_mask = (1ull << bc) - 1ull; kb.tgt_mask = (1ull << bt) - 1ull; kb.mode = 0;
    } else {
        const int bo = bits_for(std::max(total, slots) + 1), bt = bits_for(nseg_max);
        kb.cam_shift = bo; kb.seg_shift = bo + 1; kb.end_bit = bo + 1 + bt + 1;
        kb.cam_mask = 1ull; kb.tgt_mask = (1ull << bo) - 1ull; kb.mode = 1;
    }
    if (kb.end_bit > 64) return l3d_fail(c, L3D_ERR_UNSUPPORTED, "l3d_score_sweep: views x segments too large for a 64-bit sort key");
    if (Umax >= (1ll << 31)) return l3d_fail(c, L3D_ERR_UNSUPPORTED, "l3d_score_sweep: view with more than 2^31 candidates");

    int rc;
#define RES(buf, bytes, what) if ((rc = l3d_reserve(c, buf, (size_t)std::max<long long>((long long)(bytes), 16), what))) return rc
    RES(S.d_slot_score, sizeof(float) * slots, "slot scores");
    RES(S.d_keys, 8 * Umax, "keys"); RES(S.d_keys2, 8 * Umax, "keys2"); RES(S.d_vals, 4 * Umax, "vals"); RES(S.d_vals2, 4 * Umax, "vals2");
    RES(S.d_reg, 8 * Umax, "reg"); RES(S.d_dir, 16 * Umax, "dir");
    RES(S.d_meta, 16 * total, "match meta"); RES(S.d_dep, 16 * total, "match depths"); RES(S.d_os, 8 * total, "match scores"); RES(S.d_kept, total, "kept flags");
    RES(S.d_ranges, 8 * c->total_segs, "ranges"); RES(S.d_est_best, 4 * c->total_segs, "estimates"); RES(S.d_est_P, 48 * c->total_segs, "estimate points");
    RES(S.d_M, 4 * (size_t)V, "match counts"); RES(S.d_vmax, 4 * (size_t)V, "view max"); RES(S.d_work, 16 * (size_t)wmax, "work items");
    RES(S.d_camrank, 4 * (size_t)V, "cam rank"); RES(S.d_viewofrank, 4 * (size_t)V, "view of rank");
    RES(S.d_reg_of_view, 8 * (size_t)V, "region of view");
    if (cpu_sem) { RES(S.d_slot_pos, 4 * slots, "slot positions"); RES(S.d_dir64, 32 * Umax, "double directions"); }
    size_t sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (unsigned int*)nullptr, (unsigned int*)nullptr, (int)Umax, 0, kb.end_bit, c->stream);
    RES(S.d_sort_tmp, sort_bytes, "sort temp");
#undef RES
    cudaStream_t st = c->stream;
    L3D_CUDA(c, cudaMemsetAsync(S.d_slot_score.p, 0xFF, sizeof(float) * slots, st), "init slot scores");        // NaN: "not scored" (never > 0)
    L3D_CUDA(c, cudaMemsetAsync(S.d_ranges.p, 0xFF, 8 * c->total_segs, st), "init ranges");                      // (-1,-1)
    L3D_CUDA(c, cudaMemsetAsync(S.d_M.p, 0, 4 * (size_t)V, st), "init counts");
    L3D_CUDA(c, cudaMemsetAsync(S.d_kept.p, 0, (size_t)total, st), "init kept flags");     // slots [M_v, U_v) of a region are never written: k_affinity scans them all
    L3D_CUDA(c, cudaMemsetAsync(S.d_vmax.p, 0, 4 * (size_t)V, st), "init maxima");
    L3D_CUDA(c, cudaMemcpyAsync(S.d_camrank.p, rank.data(), 4 * (size_t)V, cudaMemcpyHostToDevice, st), "cam rank");
    L3D_CUDA(c, cudaMemcpyAsync(S.d_viewofrank.p, view_of_rank.data(), 4 * (size_t)V, cudaMemcpyHostToDevice, st), "view of rank");
    std::vector<long long> reg_of_view(V);
    for (int i = 0; i < V; ++i) reg_of_view[S.order[i]] = S.region_off[i];
    L3D_CUDA(c, cudaMemcpyAsync(S.d_reg_of_view.p, reg_of_view.data(), 8 * (size_t)V, cudaMemcpyHostToDevice, st), "region of view");
    L3D_CUDA(c, cudaStreamSynchronize(st), "sync");   // rank/view_of_rank/reg_of_view are stack-scoped

    // shortcut thresholds of k_score (see there); disabled (never true) when sim_t <= 0 or the margins do not apply
    float q_thr = INFINITY, cos_thr = -1.0f;
    if (min_similarity > 0.0f && min_similarity < 1.0f) {
        q_thr = 1.01f * -std::log(min_similarity);
        const double ang = std::sqrt((double)q_thr * (double)two_sigA_sqr);           // degrees
        cos_thr = ang < 89.0 ? (float)(std::cos(ang * L3D_PI_D / 180.0) * (1.0 - 1e-4)) : -1.0f;
    }
    const float4* segs = c->segs(); const float4* cache = (const float4*)c->d_cache.p;
    const L3DViewDev* views = c->views(); const L3DPairDev* pairs = (const L3DPairDev*)c->d_pairs.p;
    const int* counts = (const int*)c->d_counts.p; const l3d_match_rec* recs = (const l3d_match_rec*)c->d_recs.p;
    std::vector<int4> work_flat;   // all views' work items, uploaded once
    std::vector<int> work_off(V + 1, 0);
    for (int i = 0; i < V; ++i) { work_off[i] = (int)work_flat.size(); work_flat.insert(work_flat.end(), work[S.order[i]].begin(), work[S.order[i]].end()); }
    work_off[V] = (int)work_flat.size();
    if ((rc = l3d_reserve(c, S.d_work, 16 * std::max<size_t>(work_flat.size(), 1), "work items"))) return rc;
    if (!work_flat.empty()) L3D_CUDA(c, cudaMemcpyAsync(S.d_work.p, work_flat.data(), 16 * work_flat.size(), cudaMemcpyHostToDevice, st), "work items");
    L3D_CUDA(c, cudaStreamSynchronize(st), "sync");

    for (int i = 0; i < V; ++i) {
        const int v = S.order[i];
        const int U = S.U[v];
        const int nseg = c->h_views[v].nseg;
        if (U == 0 || nseg == 0) continue;
        const long long ro = S.region_off[i];
        int* Mc = (int*)S.d_M.p + i; int* vmax = (int*)S.d_vmax.p + i;
        int4* m_meta = (int4*)S.d_meta.p + ro; float4* m_dep = (float4*)S.d_dep.p + ro; float2* m_os = (float2*)S.d_os.p + ro;
        unsigned char* kept = (unsigned char*)S.d_kept.p + ro;
        int2* ranges = (int2*)S.d_ranges.p + c->h_views[v].seg_off;
        const int nb = (U + 255) / 256;
        k_gather<<<nb, 256, 0, st>>>(segs, views, pairs, counts, recs, (const float*)S.d_slot_score.p, (const int4*)S.d_work.p + work_off[i],
                                     work_off[i + 1] - work_off[i], v, knn, (const int*)S.d_camrank.p, (unsigned long long*)S.d_keys.p,
                                     (unsigned int*)S.d_vals.p, U, Mc, kb, cpu_sem ? (const int*)S.d_slot_pos.p : nullptr, (const long long*)S.d_reg_of_view.p);
        size_t tb = S.d_sort_tmp.cap;
        cub::DeviceRadixSort::SortPairs(S.d_sort_tmp.p, tb, (const unsigned long long*)S.d_keys.p, (unsigned long long*)S.d_keys2.p,
                                        (const unsigned int*)S.d_vals.p, (unsigned int*)S.d_vals2.p, U, 0, kb.end_bit, st);
        k_build<<<nb, 256, 0, st>>>(segs, cache, views, pairs, recs, v, knn, (const unsigned long long*)S.d_keys2.p,
                                    (const unsigned int*)S.d_vals2.p, Mc, (const int*)S.d_viewofrank.p, m_meta, m_dep, m_os,
                                    (float2*)S.d_reg.p, (float4*)S.d_dir.p, ranges, kb, NP, cpu_sem ? (double4*)S.d_dir64.p : nullptr);
        if (cpu_sem)
            k_score_cpu<<<(U + 127) / 128, 128, 0, st>>>(views, v, Mc, m_meta, m_dep, (const float2*)S.d_reg.p, (const double4*)S.d_dir64.p, ranges,
                                                        two_sigA_sqr, min_similarity, m_os);
        else
            k_score<<<(U + 127) / 128, 128, 0, st>>>(views, v, Mc, m_meta, m_dep, (const float2*)S.d_reg.p, (const float4*)S.d_dir.p, ranges,
                                                    two_sigA_sqr, min_similarity, q_thr, cos_thr, m_os);
        k_post_score<<<nb, 256, 0, st>>>(Mc, m_meta, m_os, (float*)S.d_slot_score.p, vmax, cpu_sem ? (int*)S.d_slot_pos.p : nullptr);
        k_filter<<<(nseg + 255) / 256, 256, 0, st>>>(segs, views, v, ranges, m_meta, m_dep, m_os, vmax, min_best_score, min_best_perc, kept,
                                                    (int*)S.d_est_best.p, (double*)S.d_est_P.p);
        c->launches += 5 + 8;   // + cub radix sort passes (histogram + 7 onesweep passes for 64-bit keys)
    }
    L3D_CUDA(c, cudaGetLastError(), "score sweep launch");
    // views without candidates never ran k_filter: no estimates there
    for (int i = 0; i < V; ++i) {
        const int v = S.order[i];
        if ((S.U[v] == 0 || c->h_views[v].nseg == 0) && c->h_views[v].nseg > 0)
            L3D_CUDA(c, cudaMemsetAsync((int*)S.d_est_best.p + c->h_views[v].seg_off, 0xFF, 4 * (size_t)c->h_views[v].nseg, st), "clear estimates");
    }
    S.h_M.resize(V);
    L3D_CUDA(c, cudaMemcpyAsync(S.h_M.data(), S.d_M.p, 4 * (size_t)V, cudaMemcpyDeviceToHost, st), "download counts");
    // compact the estimates on the device: flags -> exclusive scan -> gather (best match record + P1,P2)
    {
        const long long N = c->total_segs;
        if ((rc = l3d_reserve(c, S.d_est_pos, 8 * (size_t)(N + 1), "estimate positions"))) return rc;
        cub::TransformInputIterator<long long, HasEstimate, const int*> flags((const int*)S.d_est_best.p, HasEstimate());
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, flags, (long long*)S.d_est_pos.p, N, st);
        if ((rc = l3d_reserve(c, S.d_sort_tmp, tb, "scan temp"))) return rc;
        tb = S.d_sort_tmp.cap;
        L3D_CUDA(c, cub::DeviceScan::ExclusiveSum(S.d_sort_tmp.p, tb, flags, (long long*)S.d_est_pos.p, N, st), "estimate scan");
        long long last_pos = 0; int last_best = -1;
        L3D_CUDA(c, cudaMemcpyAsync(&last_pos, (long long*)S.d_est_pos.p + N - 1, 8, cudaMemcpyDeviceToHost, st), "estimate count");
        L3D_CUDA(c, cudaMemcpyAsync(&last_best, (int*)S.d_est_best.p + N - 1, 4, cudaMemcpyDeviceToHost, st), "estimate count");
        L3D_CUDA(c, cudaStreamSynchronize(st), "score sweep");
        S.n_est = last_pos + (last_best >= 0 ? 1 : 0);
        if ((rc = l3d_reserve(c, S.d_est_out_best, sizeof(l3d_match) * (size_t)std::max<long long>(S.n_est, 1), "estimate records"))) return rc;
        if ((rc = l3d_reserve(c, S.d_est_out_P, 48 * (size_t)std::max<long long>(S.n_est, 1), "estimate points"))) return rc;
        if (S.n_est > 0) {
            k_collect_estimates<<<(unsigned int)((N + 255) / 256), 256, 0, st>>>(views, V, N, (const long long*)S.d_reg_of_view.p, (const int*)S.d_est_best.p,
                                                                               (const long long*)S.d_est_pos.p, (const int4*)S.d_meta.p, (const float4*)S.d_dep.p,
                                                                               (const float2*)S.d_os.p, (const double*)S.d_est_P.p,
                                                                               (l3d_match*)S.d_est_out_best.p, (double*)S.d_est_out_P.p);
            c->launches += 3;
            L3D_CUDA(c, cudaGetLastError(), "k_collect_estimates");
        }
        L3D_CUDA(c, cudaStreamSynchronize(st), "score sweep");
    }
    S.valid = true;
    return L3D_OK;
}

// matches of one view after scoring, in the reference's list order (segment, then tgt cam, tgt seg).
// kept_only != 0: only the matches that survived filterMatches.  Returns the number of matches (even if > cap).
long long l3d_get_view_matches(l3d_ctx* c, int view, int kept_only, l3d_match* out, long long cap)
{
    if (!c) return L3D_ERR_INVALID;
    if (!c->sweep.valid) return l3d_fail(c, L3D_ERR_STATE, "l3d_get_view_matches: call l3d_score_sweep first");
    if (view < 0 || view >= c->num_views) return l3d_fail(c, L3D_ERR_INVALID, "l3d_get_view_matches: view out of range");
    cudaSetDevice(c->device);
    SweepState& S = c->sweep;
    int i = (int)(std::find(S.order.begin(), S.order.end(), view) - S.order.begin());
    const int M = S.h_M[i];
    if (M == 0) return 0;
    const long long ro = S.region_off[i];
    std::vector<int4> meta(M); std::vector<float4> dep(M); std::vector<float2> os(M); std::vector<unsigned char> kept(M);
    L3D_CUDA(c, cudaMemcpyAsync(meta.data(), (int4*)S.d_meta.p + ro, 16 * (size_t)M, cudaMemcpyDeviceToHost, c->stream), "download");
    L3D_CUDA(c, cudaMemcpyAsync(dep.data(), (float4*)S.d_dep.p + ro, 16 * (size_t)M, cudaMemcpyDeviceToHost, c->stream), "download");
    L3D_CUDA(c, cudaMemcpyAsync(os.data(), (float2*)S.d_os.p + ro, 8 * (size_t)M, cudaMemcpyDeviceToHost, c->stream), "download");
    L3D_CUDA(c, cudaMemcpyAsync(kept.data(), (unsigned char*)S.d_kept.p + ro, (size_t)M, cudaMemcpyDeviceToHost, c->stream), "download");
    L3D_CUDA(c, cudaStreamSynchronize(c->stream), "sync");
    long long n = 0;
    for (int x = 0; x < M; ++x) {
        if (kept_only && !kept[x]) continue;
        if (out && n < cap) {
            l3d_match& m = out[n];
            m.src_cam = c->h_views[view].cam_id; m.src_seg = (uint32_t)meta[x].x;
            m.tgt_cam = c->h_views[meta[x].y].cam_id; m.tgt_seg = (uint32_t)meta[x].z;
            m.overlap = os[x].x; m.score3D = os[x].y;
            m.d_p1 = dep[x].x; m.d_p2 = dep[x].y; m.d_q1 = dep[x].z; m.d_q2 = dep[x].w;
        }
        ++n;
    }
    return n;
}

// best-match 3D estimates (estimated_position3D_, line3D.cc:1635-1647), compacted on the device by k_collect_estimates
// at the end of the sweep, in global segment order (view index, segment).
long long l3d_get_estimates(l3d_ctx* c, l3d_match* best_out, double* p1p2_out, long long cap)
{
    if (!c) return L3D_ERR_INVALID;
    if (!c->sweep.valid) return l3d_fail(c, L3D_ERR_STATE, "l3d_get_estimates: call l3d_score_sweep first");
    cudaSetDevice(c->device);
    SweepState& S = c->sweep;
    const long long n = S.n_est;
    if (n == 0 || n > cap) return n;
    if (best_out) L3D_CUDA(c, cudaMemcpyAsync(best_out, S.d_est_out_best.p, sizeof(l3d_match) * (size_t)n, cudaMemcpyDeviceToHost, c->stream), "download estimates");
    if (p1p2_out) L3D_CUDA(c, cudaMemcpyAsync(p1p2_out, S.d_est_out_P.p, 48 * (size_t)n, cudaMemcpyDeviceToHost, c->stream), "download estimates");
    L3D_CUDA(c, cudaStreamSynchronize(c->stream), "sync");
    return n;
}

} // extern "C"
