// l3d_ctx.cuh — the context object behind the opaque l3d_ctx* of include/l3d_capi.h
#pragma once
#include "l3d_match.cuh"

#include <string>
#include <vector>

struct DevBuf { void* p = nullptr; size_t cap = 0; };   // grow-only device allocation

// state of the per-view scoring sweep (l3d_pipeline.cu)
struct SweepState {
    bool valid = false;
    std::vector<int> order;                 // view indices in processing order (ascending camID)
    std::vector<int> h_M;                   // per processing rank: number of matches in matches_[view] after scoring (active entries)
    std::vector<long long> region_off;      // per processing rank (+1): first entry of the view's region
    long long total = 0;
    // match store (l3d_sweep.cuh): view table, chunk descriptors, chunk offsets, entries
    DevBuf d_vt, d_vp, d_pairc, d_rowoff, d_pinfo, d_bpair, d_ekv, d_rays, d_rmf, d_rflag, d_invpos, d_csize, d_ccur, d_cstart, d_eval, d_escore, d_eflag, d_gstage, d_gpub, d_dir64, d_export, d_export_ok,
        d_ranges, d_est_best, d_est_P, d_M, d_vmax, d_sort_tmp, d_aff_sim, d_aff_flag, d_aff_gi, d_aff_gj, d_aff_pos, d_aff_oi, d_aff_oj,
        d_aff_ow, d_kx, d_order, d_rankofview, d_region_off, d_est_pos, d_est_out_best, d_est_out_P, d_segrank_off, d_evcnt, d_evptr, d_aff_par;
    long long n_est = 0, n_kept = 0;
    float ms_setup = 0.f, ms_chain = 0.f, ms_filter = 0.f;   // device time of the last sweep's three phases
    bool aff_has_parents = false;           // d_aff_par valid: candidates carry parents (collinearity links)
    std::vector<DevBuf*> bufs()
    { return {&d_vt, &d_vp, &d_pairc, &d_rowoff, &d_pinfo, &d_bpair, &d_ekv, &d_rays, &d_rmf, &d_rflag, &d_invpos, &d_csize, &d_ccur, &d_cstart, &d_eval, &d_escore, &d_eflag, &d_gstage, &d_gpub, &d_dir64,
              &d_export, &d_export_ok, &d_ranges, &d_est_best, &d_est_P, &d_M, &d_vmax, &d_sort_tmp, &d_aff_sim, &d_aff_flag, &d_aff_gi, &d_aff_gj,
              &d_aff_pos, &d_aff_oi, &d_aff_oj, &d_aff_ow, &d_kx, &d_order, &d_rankofview, &d_region_off, &d_est_pos,
              &d_est_out_best, &d_est_out_P, &d_segrank_off, &d_evcnt, &d_evptr, &d_aff_par}; }
};

// device state of the affinity-matrix bookkeeping (l3d_affinity.cu)
struct AffinityState {
    bool valid = false; float two_sigA_sqr = 0.f, med = 0.f, min_aff = 0.f; long long K = 0, n_ids = 0;
    DevBuf d_key, d_key2, d_val, d_val2, d_keep, d_keep2, d_found, d_changed, d_q, d_time, d_nflag, d_npos, d_idof, d_nkey, d_nkey2, d_nval, d_nval2, d_l2g, d_ei, d_ej, d_ew;
    std::vector<DevBuf*> bufs()
    { return {&d_key, &d_key2, &d_val, &d_val2, &d_keep, &d_keep2, &d_found, &d_changed, &d_q, &d_time, &d_nflag, &d_npos, &d_idof, &d_nkey, &d_nkey2, &d_nval, &d_nval2, &d_l2g,
              &d_ei, &d_ej, &d_ew}; }
};

// per-view lists of potentially collinear segments (l3d_collinear.cu), CSR over the global segment index
struct CollinState {
    bool valid = false; float dist_t = 0.f; int sem = 0; long long total = 0;
    DevBuf d_tiles, d_cnt, d_ptr, d_idx, d_tmp;
    std::vector<DevBuf*> bufs() { return {&d_tiles, &d_cnt, &d_ptr, &d_idx, &d_tmp}; }
};

// device state of the line bundling (l3d_optimize.cu)
struct OptState {
    DevBuf d_p, d_pout, d_valid, d_resptr, d_rescam, d_resxy, d_obs, d_cams, d_x, d_xc, d_g, d_H, d_S, d_diag, d_free, d_part, d_tot;
    std::vector<DevBuf*> bufs()
    { return {&d_p, &d_pout, &d_valid, &d_resptr, &d_rescam, &d_resxy, &d_obs, &d_cams, &d_x, &d_xc, &d_g, &d_H, &d_S, &d_diag, &d_free, &d_part, &d_tot}; }
};

// device buffers of the diffusion (l3d_affinity.cu)
struct RddState {
    DevBuf d_ei, d_ej, d_ew, d_krow, d_kcol, d_k2, d_idx, d_idx2, d_P, d_Pn, d_W, d_prow, d_pcol, d_wmaj, d_wmin, d_rowptr, d_colptr, d_tslot, d_tmp, d_len4, d_rp4, d_cp4,
        d_Pp, d_Pnp, d_Wp, d_plan, d_dst;
    std::vector<DevBuf*> bufs()
    { return {&d_ei, &d_ej, &d_ew, &d_krow, &d_kcol, &d_k2, &d_idx, &d_idx2, &d_P, &d_Pn, &d_W, &d_prow, &d_pcol, &d_wmaj, &d_wmin, &d_rowptr,
              &d_colptr, &d_tslot, &d_tmp, &d_len4, &d_rp4, &d_cp4, &d_Pp, &d_Pnp, &d_Wp, &d_plan, &d_dst}; }
};

struct l3d_ctx {
    int device = 0, num_sms = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;      // copy_stream: D2H of finished chunks (l3d_match_pairs_host)
    std::vector<cudaEvent_t> events;
    std::string err;
    long long launches = 0;

    // views
    bool have_views = false;
    int num_views = 0;
    long long total_segs = 0;
    std::vector<L3DViewDev> h_views;
    DevBuf d_segs, d_cache, d_cache_d, d_views;
    const float4* segs_ext = nullptr;       // caller-owned device segment array (l3d_set_views_flat on_device)
    void* h_stage = nullptr; size_t h_stage_cap = 0;

    // last match result
    bool have_matches = false;
    int num_pairs = 0, knn = 0;
    float epi = 0.f;
    int semantics = 0;                      // L3D_SEM_* of the last match result; the scoring sweep follows it
    long long total_rows = 0, pair_evals = 0;
    std::vector<L3DPairDev> h_pairs;
    std::vector<int2> h_tiles;
    DevBuf d_pairs, d_tiles, d_counts, d_recs, d_rowptr, d_csr, d_scan_tmp, d_dense_dep, d_dense_ov, d_dense_jobs, d_arcs, d_basis, d_arcraw, d_arckeys, d_arckeys2, d_arcvals, d_arcvals2, d_arctmp;

    SweepState sweep;
    RddState rdd;
    AffinityState aff;
    CollinState collin;
    OptState opt;

    const float4* segs() const { return segs_ext ? segs_ext : (const float4*)d_segs.p; }
    const L3DViewDev* views() const { return (const L3DViewDev*)d_views.p; }
    std::vector<DevBuf*> all_bufs()
    {
        std::vector<DevBuf*> b = {&d_segs, &d_cache, &d_cache_d, &d_views, &d_pairs, &d_tiles, &d_counts, &d_recs, &d_rowptr, &d_csr, &d_scan_tmp, &d_dense_dep, &d_dense_ov, &d_dense_jobs, &d_arcs, &d_basis, &d_arcraw, &d_arckeys, &d_arckeys2, &d_arcvals, &d_arcvals2, &d_arctmp};
        for (DevBuf* x : sweep.bufs()) b.push_back(x);
        for (DevBuf* x : rdd.bufs()) b.push_back(x);
        for (DevBuf* x : aff.bufs()) b.push_back(x);
        for (DevBuf* x : collin.bufs()) b.push_back(x);
        for (DevBuf* x : opt.bufs()) b.push_back(x);
        return b;
    }
};

int l3d_fail(l3d_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess);
int l3d_reserve(l3d_ctx* c, DevBuf& b, size_t bytes, const char* what);

#define L3D_CUDA(ctx, call, what)                                                     \
    do {                                                                              \
        cudaError_t _e = (call);                                                      \
        if (_e != cudaSuccess) return l3d_fail((ctx), L3D_ERR_CUDA, (what), _e);      \
    } while (0)
