// l3d_ctx.cuh — the context object behind the opaque l3d_ctx* of include/l3d_capi.h
#pragma once
#include "l3d_match.cuh"

#include <string>
#include <vector>

struct DevBuf { void* p = nullptr; size_t cap = 0; };   // grow-only device allocation

struct l3d_ctx {
    int device = 0, num_sms = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    long long launches = 0;

    // views
    bool have_views = false;
    int num_views = 0;
    long long total_segs = 0;
    std::vector<L3DViewDev> h_views;
    DevBuf d_segs, d_cache, d_views;
    const float4* segs_ext = nullptr;       // caller-owned device segment array (l3d_set_views_flat on_device)
    void* h_stage = nullptr; size_t h_stage_cap = 0;

    // last match result
    bool have_matches = false;
    int num_pairs = 0, knn = 0;
    float epi = 0.f;
    long long total_rows = 0, pair_evals = 0;
    std::vector<L3DPairDev> h_pairs;
    std::vector<int2> h_tiles;
    DevBuf d_pairs, d_tiles, d_counts, d_recs, d_rowptr, d_csr, d_scan_tmp, d_dense_dep, d_dense_ov;

    const float4* segs() const { return segs_ext ? segs_ext : (const float4*)d_segs.p; }
    const L3DViewDev* views() const { return (const L3DViewDev*)d_views.p; }
    std::vector<DevBuf*> all_bufs()
    { return {&d_segs, &d_cache, &d_views, &d_pairs, &d_tiles, &d_counts, &d_recs, &d_rowptr, &d_csr, &d_scan_tmp, &d_dense_dep, &d_dense_ov}; }
};

int l3d_fail(l3d_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess);
int l3d_reserve(l3d_ctx* c, DevBuf& b, size_t bytes, const char* what);

#define L3D_CUDA(ctx, call, what)                                                     \
    do {                                                                              \
        cudaError_t _e = (call);                                                      \
        if (_e != cudaSuccess) return l3d_fail((ctx), L3D_ERR_CUDA, (what), _e);      \
    } while (0)
