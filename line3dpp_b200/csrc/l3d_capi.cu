// l3d_capi.cu — context, device memory and the C-ABI entry points declared in include/l3d_capi.h.
// Host C++ only orchestrates: every byte of arithmetic on the hot path runs in the kernels of l3d_match.cu /
// l3d_pipeline.cu.  There is no CPU fallback: every entry needs a live CUDA context and fails loudly otherwise.
#include "l3d_ctx.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstring>

int l3d_fail(l3d_ctx* c, int code, const char* what, cudaError_t e)
{
    if (c) {
        char buf[512];
        if (e != cudaSuccess) snprintf(buf, sizeof(buf), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
        else snprintf(buf, sizeof(buf), "%s", what);
        c->err = buf;
    }
    return code;
}

int l3d_reserve(l3d_ctx* c, DevBuf& b, size_t bytes, const char* what)
{
    if (bytes <= b.cap) return L3D_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) { e = cudaMalloc(&b.p, bytes); want = bytes; }
    if (e != cudaSuccess) { b.p = nullptr; return l3d_fail(c, L3D_ERR_NOMEM, what, e); }
    b.cap = want;
    return L3D_OK;
}

struct CastI64 { __host__ __device__ long long operator()(int v) const { return (long long)v; } };

extern "C" {

int l3d_ctx_create(int device, l3d_ctx** out)
{
    if (!out) return L3D_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) return L3D_ERR_CUDA;   // no GPU, no library: no CPU fallback
    l3d_ctx* c = new l3d_ctx();
    c->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete c;
        return L3D_ERR_CUDA;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    c->num_sms = prop.multiProcessorCount;
    e = cudaFuncSetAttribute(k_match_topk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l3d_match_smem_bytes());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_match_all, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l3d_match_smem_bytes());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_match_topk_f64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l3d_match_smem_bytes());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_match_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l3d_dense_smem_bytes());
    if (e != cudaSuccess) { cudaStreamDestroy(c->stream); delete c; return L3D_ERR_CUDA; }
    *out = c;
    return L3D_OK;
}

void l3d_ctx_destroy(l3d_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (DevBuf* b : c->all_bufs()) if (b->p) cudaFree(b->p);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    for (cudaEvent_t e : c->events) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->stream);
    delete c;
}

const char* l3d_last_error(const l3d_ctx* c) { return c ? c->err.c_str() : "null context"; }
void* l3d_stream(l3d_ctx* c) { return c ? (void*)c->stream : nullptr; }
long long l3d_launch_count(const l3d_ctx* c) { return c ? c->launches : 0; }
int l3d_sync(l3d_ctx* c)
{
    if (!c) return L3D_ERR_INVALID;
    cudaError_t e = cudaStreamSynchronize(c->stream);
    return e == cudaSuccess ? L3D_OK : l3d_fail(c, L3D_ERR_CUDA, "l3d_sync", e);
}

// ------------------------------------------------------------------------------------------------ views
static int set_views_common(l3d_ctx* c, int V, const l3d_view_desc* views)
{
    c->h_views.resize(V);
    long long off = 0;
    for (int v = 0; v < V; ++v) {
        if (views[v].nseg < 0) return l3d_fail(c, L3D_ERR_INVALID, "l3d_set_views: negative nseg");
        L3DViewDev& d = c->h_views[v];
        d.seg_off = off; d.nseg = views[v].nseg; d.cam_id = views[v].cam_id; d.width = views[v].width; d.height = views[v].height;
        memcpy(d.RtKinv, views[v].RtKinv, sizeof(d.RtKinv)); memcpy(d.C, views[v].C, sizeof(d.C));
        memcpy(d.RtKinv_d, views[v].RtKinv_d, sizeof(d.RtKinv_d)); memcpy(d.C_d, views[v].C_d, sizeof(d.C_d));
        d.k = views[v].k; d.median_depth = views[v].median_depth;
        off += views[v].nseg;
    }
    c->num_views = V; c->total_segs = off;
    return L3D_OK;
}

static int finish_views(l3d_ctx* c)
{
    int rc;
    if ((rc = l3d_reserve(c, c->d_views, sizeof(L3DViewDev) * (size_t)c->num_views, "views"))) return rc;
    if ((rc = l3d_reserve(c, c->d_cache, sizeof(float4) * 3 * (size_t)c->total_segs, "segment cache"))) return rc;
    L3D_CUDA(c, cudaMemcpyAsync(c->d_views.p, c->h_views.data(), sizeof(L3DViewDev) * c->num_views, cudaMemcpyHostToDevice, c->stream), "upload views");
    if (c->total_segs > 0) {
        unsigned int blocks = (unsigned int)((c->total_segs + 255) / 256);
        k_prep_segments<<<blocks, 256, 0, c->stream>>>(c->segs(), c->views(), c->num_views, c->total_segs, (float4*)c->d_cache.p);
        ++c->launches;
        L3D_CUDA(c, cudaGetLastError(), "k_prep_segments");
    }
    c->have_views = true; c->have_matches = false; c->sweep.valid = false; c->collin.valid = false; c->aff.valid = false;
    return L3D_OK;
}

int l3d_set_views(l3d_ctx* c, int V, const l3d_view_desc* views, const float* const* segs_host)
{
    if (!c || V <= 0 || !views || !segs_host) return l3d_fail(c, L3D_ERR_INVALID, "l3d_set_views: bad arguments");
    cudaSetDevice(c->device);
    int rc = set_views_common(c, V, views);
    if (rc) return rc;
    size_t bytes = sizeof(float4) * (size_t)c->total_segs;
    if ((rc = l3d_reserve(c, c->d_segs, bytes, "segments"))) return rc;
    if (bytes > c->h_stage_cap) {           // pinned staging so the H2D is one full-speed async copy
        if (c->h_stage) cudaFreeHost(c->h_stage);
        c->h_stage = nullptr; c->h_stage_cap = 0;
        L3D_CUDA(c, cudaMallocHost(&c->h_stage, bytes), "pinned staging");
        c->h_stage_cap = bytes;
    }
    L3D_CUDA(c, cudaStreamSynchronize(c->stream), "sync before staging");
    for (int v = 0; v < V; ++v)
        if (views[v].nseg) memcpy((char*)c->h_stage + sizeof(float4) * c->h_views[v].seg_off, segs_host[v], sizeof(float4) * (size_t)views[v].nseg);
    if (bytes) L3D_CUDA(c, cudaMemcpyAsync(c->d_segs.p, c->h_stage, bytes, cudaMemcpyHostToDevice, c->stream), "upload segments");
    c->segs_ext = nullptr;
    return finish_views(c);
}

int l3d_set_views_flat(l3d_ctx* c, int V, const l3d_view_desc* views, const float* segs_flat, int on_device)
{
    if (!c || V <= 0 || !views || !segs_flat) return l3d_fail(c, L3D_ERR_INVALID, "l3d_set_views_flat: bad arguments");
    cudaSetDevice(c->device);
    int rc = set_views_common(c, V, views);
    if (rc) return rc;
    size_t bytes = sizeof(float4) * (size_t)c->total_segs;
    if (on_device) {
        if (((uintptr_t)segs_flat & 15) != 0) return l3d_fail(c, L3D_ERR_INVALID, "l3d_set_views_flat: device array must be 16-byte aligned");
        c->segs_ext = (const float4*)segs_flat;
    } else {
        if ((rc = l3d_reserve(c, c->d_segs, bytes, "segments"))) return rc;
        if (bytes) L3D_CUDA(c, cudaMemcpyAsync(c->d_segs.p, segs_flat, bytes, cudaMemcpyHostToDevice, c->stream), "upload segments");
        c->segs_ext = nullptr;
    }
    return finish_views(c);
}

int l3d_update_view_params(l3d_ctx* c, int V, const l3d_view_desc* views)
{
    if (!c || !views || V != c->num_views || !c->have_views) return l3d_fail(c, L3D_ERR_STATE, "l3d_update_view_params: views not set / count mismatch");
    cudaSetDevice(c->device);
    for (int v = 0; v < V; ++v) {
        L3DViewDev& d = c->h_views[v];
        memcpy(d.RtKinv_d, views[v].RtKinv_d, sizeof(d.RtKinv_d)); memcpy(d.C_d, views[v].C_d, sizeof(d.C_d));
        d.k = views[v].k; d.median_depth = views[v].median_depth;
    }
    c->aff.valid = false;
    L3D_CUDA(c, cudaStreamSynchronize(c->stream), "sync before view update");
    L3D_CUDA(c, cudaMemcpyAsync(c->d_views.p, c->h_views.data(), sizeof(L3DViewDev) * V, cudaMemcpyHostToDevice, c->stream), "upload views");
    return L3D_OK;
}

// ------------------------------------------------------------------------------------------------ matching
int l3d_match_pairs(l3d_ctx* c, int num_pairs, const int32_t* pairs, const float* F, float epi_overlap, int knn)
{ return l3d_match_pairs_range(c, num_pairs, pairs, F, epi_overlap, knn, 0, num_pairs); }

struct HostOut { int32_t* counts; l3d_match_rec* recs; int chunks; };
static int match_impl(l3d_ctx* c, int num_pairs, const int32_t* pairs, const float* F, const double* Fd, float epi_overlap, int knn, int first_pair, int last_pair,
                      const HostOut* host = nullptr);

int l3d_match_pairs_range(l3d_ctx* c, int num_pairs, const int32_t* pairs, const float* F, float epi_overlap, int knn, int first_pair, int last_pair)
{ return match_impl(c, num_pairs, pairs, F, nullptr, epi_overlap, knn, first_pair, last_pair); }

int l3d_match_pairs_f64(l3d_ctx* c, int num_pairs, const int32_t* pairs, const double* Fd, float epi_overlap, int knn, int first_pair, int last_pair)
{
    if (!c) return L3D_ERR_INVALID;
    if (num_pairs > 0 && !Fd) return l3d_fail(c, L3D_ERR_INVALID, "l3d_match_pairs_f64: bad arguments");
    return match_impl(c, num_pairs, pairs, nullptr, Fd, epi_overlap, knn, first_pair, last_pair);
}

int l3d_match_pairs_host(l3d_ctx* c, int num_pairs, const int32_t* pairs, const float* F, float epi_overlap, int knn, int32_t* counts_out,
                         l3d_match_rec* recs_out, int chunks)
{
    if (!c) return L3D_ERR_INVALID;
    if (!counts_out || !recs_out || chunks < 1) return l3d_fail(c, L3D_ERR_INVALID, "l3d_match_pairs_host: bad arguments");
    if (knn <= 0) return l3d_fail(c, L3D_ERR_UNSUPPORTED, "l3d_match_pairs_host: kNN <= 0 has no fixed row stride; use l3d_match_pairs + l3d_get_pair_matches");
    HostOut h = {counts_out, recs_out, std::min(chunks, 64)};
    return match_impl(c, num_pairs, pairs, F, nullptr, epi_overlap, knn, 0, num_pairs, &h);
}

static int match_impl(l3d_ctx* c, int num_pairs, const int32_t* pairs, const float* F, const double* Fd, float epi_overlap, int knn, int first_pair, int last_pair,
                      const HostOut* host)
{
    if (!c) return L3D_ERR_INVALID;
    if (!c->have_views) return l3d_fail(c, L3D_ERR_STATE, "l3d_match_pairs: call l3d_set_views first");
    if (num_pairs < 0 || (num_pairs && (!pairs || (!F && !Fd)))) return l3d_fail(c, L3D_ERR_INVALID, "l3d_match_pairs: bad arguments");
    if (first_pair < 0 || last_pair < first_pair || last_pair > num_pairs) return l3d_fail(c, L3D_ERR_INVALID, "l3d_match_pairs_range: bad pair range");
    // kNN > 32 exceeds the fused kernel's per-row key list: the keep-all passes run instead and every row is cut to its kNN best
    // afterwards (same matches as the reference's priority queue, line3D.cc:999-1006 / cudawrapper.cu:637-645)
    const int big_k = knn > 32 ? knn : 0;
    const bool keep_all = knn <= 0 || big_k > 0;
    if (big_k && host) return l3d_fail(c, L3D_ERR_UNSUPPORTED, "l3d_match_pairs_host: kNN > 32 has no fixed row stride; use l3d_match_pairs + l3d_get_pair_matches");
    cudaSetDevice(c->device);
    L3D_CUDA(c, cudaStreamSynchronize(c->stream), "sync before pair staging");   // h_pairs/h_tiles may still be in flight
    c->h_pairs.resize(num_pairs);
    c->h_tiles.clear();
    long long rows = 0, evals = 0, narcs = 0;
    for (int i = 0; i < num_pairs; ++i) {
        int s = pairs[2 * i], t = pairs[2 * i + 1];
        if (s < 0 || t < 0 || s >= c->num_views || t >= c->num_views) return l3d_fail(c, L3D_ERR_INVALID, "l3d_match_pairs: view index out of range");
        L3DPairDev& p = c->h_pairs[i];
        p.src = s; p.tgt = t; p.row_off = rows;
        if (Fd) {     // REF_CPU: the double matrix decides; the float copy (eigen2dataArray-style cast) only feeds the conservative filter
            memcpy(p.Fd, Fd + 9 * (size_t)i, sizeof(p.Fd));
            for (int k = 0; k < 9; ++k) p.F[k] = (float)p.Fd[k];
        } else { memcpy(p.F, F + 9 * (size_t)i, sizeof(p.F)); memset(p.Fd, 0, sizeof(p.Fd)); }
        int Ns = c->h_views[s].nseg, Nt = c->h_views[t].nseg;
        if (Nt >= (1 << 24)) return l3d_fail(c, L3D_ERR_UNSUPPORTED, "l3d_match_pairs: more than 2^24 segments in one view");
        p.arc_off = narcs;
        if (i >= first_pair && i < last_pair) narcs += Nt;
        if (Nt > 0 && i >= first_pair && i < last_pair)
            for (int r0 = 0; r0 < Ns; r0 += MK_ROWS) c->h_tiles.push_back(make_int2(i, r0));
        rows += Ns;
        if (i >= first_pair && i < last_pair) evals += (long long)Ns * Nt;     // evaluated HERE
    }
    c->num_pairs = num_pairs; c->knn = knn; c->epi = epi_overlap; c->total_rows = rows; c->pair_evals = evals;
    int rc;
    if ((rc = l3d_reserve(c, c->d_pairs, sizeof(L3DPairDev) * (size_t)std::max(num_pairs, 1), "pairs"))) return rc;
    if ((rc = l3d_reserve(c, c->d_tiles, sizeof(int2) * std::max<size_t>(c->h_tiles.size(), 1), "tiles"))) return rc;
    if ((rc = l3d_reserve(c, c->d_arcs, sizeof(uint4) * (size_t)std::max<long long>(narcs, 1), "target arcs"))) return rc;
    if ((rc = l3d_reserve(c, c->d_basis, sizeof(L3DPairBasis) * (size_t)std::max(num_pairs, 1), "pair bases"))) return rc;
    if ((rc = l3d_reserve(c, c->d_counts, sizeof(int) * (size_t)std::max<long long>(rows, 1), "match counts"))) return rc;
    if (!keep_all && (rc = l3d_reserve(c, c->d_recs, sizeof(l3d_match_rec) * (size_t)std::max<long long>(rows, 1) * knn, "match records"))) return rc;
    if (num_pairs) L3D_CUDA(c, cudaMemcpyAsync(c->d_pairs.p, c->h_pairs.data(), sizeof(L3DPairDev) * num_pairs, cudaMemcpyHostToDevice, c->stream), "upload pairs");
    if (!c->h_tiles.empty()) L3D_CUDA(c, cudaMemcpyAsync(c->d_tiles.p, c->h_tiles.data(), sizeof(int2) * c->h_tiles.size(), cudaMemcpyHostToDevice, c->stream), "upload tiles");
    if (rows) L3D_CUDA(c, cudaMemsetAsync(c->d_counts.p, 0, sizeof(int) * rows, c->stream), "clear counts");   // rows of pairs with Nt == 0
    const uint4* arcs = (const uint4*)c->d_arcs.p;
    const L3DPairBasis* basis = (const L3DPairBasis*)c->d_basis.p;
    if (last_pair > first_pair && narcs > 0) {
        // level-1 tables of the pairs matched here (l3d_device.cuh "pencil parameter"): raw arcs + keys, one sort, packed entries in window
        // order.  Below epi_overlap 0.01 the extended arcs would span the whole line: everything is then handed to the float filter.
        if (narcs >= (1ll << 31) - 2) return l3d_fail(c, L3D_ERR_UNSUPPORTED, "l3d_match_pairs: more than 2^31 (pair, target segment) combinations in one call");
        const int npr = last_pair - first_pair;
        const bool on = epi_overlap >= 0.01f && getenv("L3D_NO_LEVEL1") == nullptr;
        int end_bit = 36;                       // key = pair << 36 | class << 32 | start of the arc
        while (end_bit < 64 && (1ll << (end_bit - 36)) < npr) ++end_bit;
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const unsigned int*)nullptr, (unsigned int*)nullptr, (int)narcs, 0, end_bit, c->stream);
        if ((rc = l3d_reserve(c, c->d_arcraw, 16 * (size_t)narcs, "raw arcs")) || (rc = l3d_reserve(c, c->d_arckeys, 8 * (size_t)narcs, "arc keys")) ||
            (rc = l3d_reserve(c, c->d_arckeys2, 8 * (size_t)narcs, "arc keys")) || (rc = l3d_reserve(c, c->d_arcvals, 4 * (size_t)narcs, "arc order")) ||
            (rc = l3d_reserve(c, c->d_arcvals2, 4 * (size_t)narcs, "arc order")) || (rc = l3d_reserve(c, c->d_arctmp, tb, "arc sort temp"))) return rc;
        k_pair_arcs<<<(unsigned int)npr, 256, 0, c->stream>>>(c->segs(), c->views(), (const L3DPairDev*)c->d_pairs.p, first_pair, on ? 1 : 0,
                                                             1.0 / (double)std::max(epi_overlap, 0.01f) + 0.5, (uint4*)c->d_arcraw.p,
                                                             (unsigned long long*)c->d_arckeys.p, (unsigned int*)c->d_arcvals.p, (L3DPairBasis*)c->d_basis.p);
        L3D_CUDA(c, cudaGetLastError(), "k_pair_arcs");
        L3D_CUDA(c, cub::DeviceRadixSort::SortPairs(c->d_arctmp.p, tb, (const unsigned long long*)c->d_arckeys.p, (unsigned long long*)c->d_arckeys2.p,
                                                    (const unsigned int*)c->d_arcvals.p, (unsigned int*)c->d_arcvals2.p, (int)narcs, 0, end_bit, c->stream), "arc sort");
        k_arcs_gather<<<(unsigned int)((narcs + 255) / 256), 256, 0, c->stream>>>(narcs, (const unsigned long long*)c->d_arckeys2.p, (const unsigned int*)c->d_arcvals2.p,
                                                                                 (const uint4*)c->d_arcraw.p, (const L3DPairDev*)c->d_pairs.p, first_pair, (uint4*)c->d_arcs.p);
        L3D_CUDA(c, cudaGetLastError(), "k_arcs_gather");
        c->launches += 2 + 8;
    }
    const double* cache_d = nullptr;
    if (Fd && c->total_segs > 0) {      // matchingCPU's rays / plane normals in double; camera blocks may have been updated since set_views
        if ((rc = l3d_reserve(c, c->d_cache_d, sizeof(double) * 9 * (size_t)c->total_segs, "double segment cache"))) return rc;
        k_prep_segments_f64<<<(unsigned int)((c->total_segs + 255) / 256), 256, 0, c->stream>>>(c->segs(), c->views(), c->num_views, c->total_segs, (double*)c->d_cache_d.p);
        ++c->launches;
        L3D_CUDA(c, cudaGetLastError(), "k_prep_segments_f64");
        cache_d = (const double*)c->d_cache_d.p;
    }
    c->semantics = Fd ? L3D_SEM_REF_CPU : L3D_SEM_REF_GPU;
    if (keep_all) {
        // "keep all matches" (cudawrapper.cu:628-636, line3D.cc:988-996): count pass -> row stride = largest row -> store pass -> sort rows
        int stride = 1;
        if (!c->h_tiles.empty()) {
            k_match_all<<<(unsigned int)c->h_tiles.size(), MK_THREADS, l3d_match_smem_bytes(), c->stream>>>(
                c->segs(), (const float4*)c->d_cache.p, c->views(), (const L3DPairDev*)c->d_pairs.p, (const int2*)c->d_tiles.p, 0,
                epi_overlap, (int*)c->d_counts.p, nullptr, cache_d, arcs, basis);
            ++c->launches;
            L3D_CUDA(c, cudaGetLastError(), "k_match_all (count)");
            size_t tb = 0;
            cub::DeviceReduce::Max(nullptr, tb, (const int*)c->d_counts.p, (int*)nullptr, (int)std::min<long long>(rows, INT_MAX), c->stream);
            if (rows > INT_MAX) return l3d_fail(c, L3D_ERR_UNSUPPORTED, "l3d_match_pairs: more than 2^31 rows with kNN <= 0");
            if ((rc = l3d_reserve(c, c->d_scan_tmp, tb, "reduce temp"))) return rc;
            if ((rc = l3d_reserve(c, c->d_rowptr, 16, "row max"))) return rc;
            int* d_max = (int*)c->d_rowptr.p;
            L3D_CUDA(c, cub::DeviceReduce::Max(c->d_scan_tmp.p, tb, (const int*)c->d_counts.p, d_max, (int)rows, c->stream), "row maximum");
            c->launches += 1;
            L3D_CUDA(c, cudaMemcpyAsync(&stride, d_max, sizeof(int), cudaMemcpyDeviceToHost, c->stream), "row maximum");
            L3D_CUDA(c, cudaStreamSynchronize(c->stream), "sync");
            stride = std::max(stride, 1);
        }
        knn = stride; c->knn = stride;
        if ((rc = l3d_reserve(c, c->d_recs, sizeof(l3d_match_rec) * (size_t)std::max<long long>(rows, 1) * stride, "match records"))) return rc;
        if (!c->h_tiles.empty()) {
            const size_t per_warp = sizeof(l3d_match_rec) * (size_t)stride;
            int wpb = 4;
            while (wpb > 1 && per_warp * wpb > 200 * 1024) wpb >>= 1;
            if (per_warp * wpb > 200 * 1024) return l3d_fail(c, L3D_ERR_UNSUPPORTED, "l3d_match_pairs: a row with more than 8500 matches (kNN <= 0)");
            k_match_all<<<(unsigned int)c->h_tiles.size(), MK_THREADS, l3d_match_smem_bytes(), c->stream>>>(
                c->segs(), (const float4*)c->d_cache.p, c->views(), (const L3DPairDev*)c->d_pairs.p, (const int2*)c->d_tiles.p, stride,
                epi_overlap, (int*)c->d_counts.p, (l3d_match_rec*)c->d_recs.p, cache_d, arcs, basis);
            L3D_CUDA(c, cudaGetLastError(), "k_match_all (store)");
            L3D_CUDA(c, cudaFuncSetAttribute(k_sort_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * wpb)), "k_sort_rows smem");
            const long long blocks = std::min<long long>((rows + wpb - 1) / wpb, (long long)c->num_sms * 16);
            k_sort_rows<<<(unsigned int)blocks, wpb * 32, per_warp * wpb, c->stream>>>((int*)c->d_counts.p, (l3d_match_rec*)c->d_recs.p, stride, rows, big_k);
            L3D_CUDA(c, cudaGetLastError(), "k_sort_rows");
            c->launches += 2;
        }
    } else {
        // One launch, or (l3d_match_pairs_host) `chunks` launches over contiguous tile ranges with the D2H copy of every finished
        // chunk's rows queued on a second stream, so that the PCIe transfer of chunk k hides behind the arithmetic of chunk k+1.
        const size_t ntiles = c->h_tiles.size();
        const int chunks = host ? std::max(1, std::min<int>(host->chunks, (int)std::max<size_t>(ntiles, 1))) : 1;
        if (host) {
            if (!c->copy_stream) L3D_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking), "copy stream");
            while ((int)c->events.size() < chunks + 1) {
                cudaEvent_t e; L3D_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event");
                c->events.push_back(e);
            }
        }
        auto row_of_tile = [&](size_t t) { return t >= ntiles ? rows : c->h_pairs[c->h_tiles[t].x].row_off + c->h_tiles[t].y; };
        long long row_done = 0;
        for (int k = 0; k < chunks; ++k) {
            const size_t t0 = ntiles * (size_t)k / chunks, t1 = ntiles * (size_t)(k + 1) / chunks;
            if (t1 > t0) {
                if (cache_d)
                    k_match_topk_f64<<<(unsigned int)(t1 - t0), MK_THREADS, l3d_match_smem_bytes(), c->stream>>>(
                        c->segs(), (const float4*)c->d_cache.p, c->views(), (const L3DPairDev*)c->d_pairs.p, (const int2*)c->d_tiles.p + t0, knn,
                        epi_overlap, (int*)c->d_counts.p, (l3d_match_rec*)c->d_recs.p, cache_d, arcs, basis);
                else
                    k_match_topk<<<(unsigned int)(t1 - t0), MK_THREADS, l3d_match_smem_bytes(), c->stream>>>(
                        c->segs(), (const float4*)c->d_cache.p, c->views(), (const L3DPairDev*)c->d_pairs.p, (const int2*)c->d_tiles.p + t0, knn,
                        epi_overlap, (int*)c->d_counts.p, (l3d_match_rec*)c->d_recs.p, arcs, basis);
                ++c->launches;
                L3D_CUDA(c, cudaGetLastError(), "k_match_topk");
            }
            if (host) {
                const long long r1 = k == chunks - 1 ? rows : row_of_tile(t1);      // rows [row_done, r1) are final after this launch
                if (r1 > row_done) {
                    L3D_CUDA(c, cudaEventRecord(c->events[k], c->stream), "event record");
                    L3D_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->events[k], 0), "event wait");
                    L3D_CUDA(c, cudaMemcpyAsync(host->counts + row_done, (const int*)c->d_counts.p + row_done, sizeof(int) * (size_t)(r1 - row_done), cudaMemcpyDeviceToHost, c->copy_stream), "download counts");
                    L3D_CUDA(c, cudaMemcpyAsync(host->recs + row_done * knn, (const l3d_match_rec*)c->d_recs.p + row_done * knn,
                                                sizeof(l3d_match_rec) * (size_t)(r1 - row_done) * knn, cudaMemcpyDeviceToHost, c->copy_stream), "download records");
                    row_done = r1;
                }
            }
        }
        if (host) {     // the context's stream (the one callers time and synchronise) completes only after the last copy
            L3D_CUDA(c, cudaEventRecord(c->events[chunks], c->copy_stream), "event record");
            L3D_CUDA(c, cudaStreamWaitEvent(c->stream, c->events[chunks], 0), "event wait");
        }
    }
    c->have_matches = true; c->sweep.valid = false;
    return L3D_OK;
}

int l3d_match_device_buffers(l3d_ctx* c, void** counts_dev, void** recs_dev)
{
    if (!c || !counts_dev || !recs_dev) return L3D_ERR_INVALID;
    if (!c->have_matches) return l3d_fail(c, L3D_ERR_STATE, "l3d_match_device_buffers: no match result");
    *counts_dev = c->d_counts.p; *recs_dev = c->d_recs.p;
    c->sweep.valid = false;                 // the caller may overwrite rows
    return L3D_OK;
}

int l3d_pair_row_offsets(l3d_ctx* c, long long* out)
{
    if (!c || !out) return L3D_ERR_INVALID;
    if (!c->have_matches) return l3d_fail(c, L3D_ERR_STATE, "l3d_pair_row_offsets: no match result");
    for (int i = 0; i < c->num_pairs; ++i) out[i] = c->h_pairs[i].row_off;
    out[c->num_pairs] = c->total_rows;
    return L3D_OK;
}

int l3d_balanced_split(const long long* cost, int n, int parts, int32_t* bounds)
{
    if (n < 0 || parts < 1 || !bounds || (n && !cost)) return L3D_ERR_INVALID;
    long long total = 0;
    for (int i = 0; i < n; ++i) { if (cost[i] < 0) return L3D_ERR_INVALID; total += cost[i]; }
    // item i goes to the part whose cost window contains the midpoint of i's cost interval: monotone, so ranges are contiguous
    bounds[0] = 0;
    int part = 0; long long before = 0;
    for (int i = 0; i < n; ++i) {
        const long double mid = (long double)before + 0.5L * (long double)cost[i];
        int want = total > 0 ? (int)(mid * parts / (long double)total) : (int)((long long)i * parts / n);
        if (want >= parts) want = parts - 1;
        while (part < want) bounds[++part] = i;
        before += cost[i];
    }
    while (part < parts) bounds[++part] = n;
    return L3D_OK;
}

int l3d_match_stride(const l3d_ctx* c) { return c && c->have_matches ? c->knn : L3D_ERR_STATE; }
int l3d_match_semantics(const l3d_ctx* c) { return c && c->have_matches ? c->semantics : L3D_ERR_STATE; }
long long l3d_match_total_rows(const l3d_ctx* c) { return c && c->have_matches ? c->total_rows : -1; }
long long l3d_match_pair_evals(const l3d_ctx* c) { return c && c->have_matches ? c->pair_evals : -1; }

long long l3d_get_match_counts(l3d_ctx* c, int32_t* counts_out)
{
    if (!c || !counts_out) return L3D_ERR_INVALID;
    if (!c->have_matches) return l3d_fail(c, L3D_ERR_STATE, "l3d_get_match_counts: no match result");
    cudaSetDevice(c->device);
    if (c->total_rows) {
        cudaError_t e = cudaMemcpyAsync(counts_out, c->d_counts.p, sizeof(int) * c->total_rows, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) return l3d_fail(c, L3D_ERR_CUDA, "download counts", e);
    }
    long long total = 0;
    for (long long i = 0; i < c->total_rows; ++i) total += counts_out[i];
    return total;
}

// number of matches of the last match result (sum of the per-row counts), reduced on the device: no download of the count array
long long l3d_match_total_matches(l3d_ctx* c)
{
    if (!c) return L3D_ERR_INVALID;
    if (!c->have_matches) return l3d_fail(c, L3D_ERR_STATE, "l3d_match_total_matches: no match result");
    if (c->total_rows == 0) return 0;
    cudaSetDevice(c->device);
    int rc;
    long long total = 0;
    const long long BLK = 1ll << 30;
    if ((rc = l3d_reserve(c, c->d_rowptr, 16, "match total"))) return rc;
    for (long long base = 0; base < c->total_rows; base += BLK) {
        const int n = (int)std::min(BLK, c->total_rows - base);
        size_t tb = 0;
        cub::DeviceReduce::Sum(nullptr, tb, (const int*)c->d_counts.p + base, (long long*)c->d_rowptr.p, n, c->stream);
        if ((rc = l3d_reserve(c, c->d_scan_tmp, tb, "reduce temp"))) return rc;
        tb = c->d_scan_tmp.cap;
        long long part = 0;
        L3D_CUDA(c, cub::DeviceReduce::Sum(c->d_scan_tmp.p, tb, (const int*)c->d_counts.p + base, (long long*)c->d_rowptr.p, n, c->stream), "match total");
        L3D_CUDA(c, cudaMemcpyAsync(&part, c->d_rowptr.p, 8, cudaMemcpyDeviceToHost, c->stream), "match total");
        L3D_CUDA(c, cudaStreamSynchronize(c->stream), "match total");
        total += part;
        ++c->launches;
    }
    return total;
}

int l3d_get_pair_matches(l3d_ctx* c, int pair, int32_t* counts_out, l3d_match_rec* recs_out)
{
    if (!c || !counts_out || !recs_out) return L3D_ERR_INVALID;
    if (!c->have_matches) return l3d_fail(c, L3D_ERR_STATE, "l3d_get_pair_matches: no match result");
    if (pair < 0 || pair >= c->num_pairs) return l3d_fail(c, L3D_ERR_INVALID, "l3d_get_pair_matches: pair out of range");
    cudaSetDevice(c->device);
    const L3DPairDev& p = c->h_pairs[pair];
    int Ns = c->h_views[p.src].nseg;
    if (Ns == 0) return L3D_OK;
    L3D_CUDA(c, cudaMemcpyAsync(counts_out, (const int*)c->d_counts.p + p.row_off, sizeof(int) * Ns, cudaMemcpyDeviceToHost, c->stream), "download pair counts");
    L3D_CUDA(c, cudaMemcpyAsync(recs_out, (const l3d_match_rec*)c->d_recs.p + p.row_off * c->knn, sizeof(l3d_match_rec) * (size_t)Ns * c->knn, cudaMemcpyDeviceToHost, c->stream), "download pair records");
    L3D_CUDA(c, cudaStreamSynchronize(c->stream), "sync");
    return L3D_OK;
}

long long l3d_get_matches_csr(l3d_ctx* c, int64_t* row_ptr_out, l3d_match_rec* recs_out, long long capacity)
{
    if (!c || !row_ptr_out) return L3D_ERR_INVALID;
    if (!c->have_matches) return l3d_fail(c, L3D_ERR_STATE, "l3d_get_matches_csr: no match result");
    cudaSetDevice(c->device);
    const long long rows = c->total_rows;
    if (rows == 0) { row_ptr_out[0] = 0; return 0; }
    int rc;
    if ((rc = l3d_reserve(c, c->d_rowptr, sizeof(long long) * (size_t)(rows + 1), "row_ptr"))) return rc;
    cub::TransformInputIterator<long long, CastI64, const int*> in((const int*)c->d_counts.p, CastI64());
    size_t tmp_bytes = 0;
    // rows+1 items: the element past the end is never read by an exclusive scan's last output
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, (long long*)c->d_rowptr.p, rows, c->stream);
    if ((rc = l3d_reserve(c, c->d_scan_tmp, tmp_bytes, "scan temp"))) return rc;
    L3D_CUDA(c, cub::DeviceScan::ExclusiveSum(c->d_scan_tmp.p, tmp_bytes, in, (long long*)c->d_rowptr.p, rows, c->stream), "row_ptr scan");
    c->launches += 2;   // cub: decoupled look-back init + scan kernels
    long long last_ptr = 0; int last_cnt = 0;
    L3D_CUDA(c, cudaMemcpyAsync(&last_ptr, (const long long*)c->d_rowptr.p + rows - 1, sizeof(long long), cudaMemcpyDeviceToHost, c->stream), "read total");
    L3D_CUDA(c, cudaMemcpyAsync(&last_cnt, (const int*)c->d_counts.p + rows - 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream), "read total");
    L3D_CUDA(c, cudaStreamSynchronize(c->stream), "sync");
    const long long total = last_ptr + last_cnt;
    L3D_CUDA(c, cudaMemcpyAsync((long long*)c->d_rowptr.p + rows, &total, sizeof(long long), cudaMemcpyHostToDevice, c->stream), "write total");
    if ((rc = l3d_reserve(c, c->d_csr, sizeof(l3d_match_rec) * (size_t)std::max<long long>(total, 1), "csr records"))) return rc;
    long long cells = rows * c->knn;
    k_compact_matches<<<(unsigned int)((cells + 255) / 256), 256, 0, c->stream>>>((const int*)c->d_counts.p, (const long long*)c->d_rowptr.p,
                                                                                   (const l3d_match_rec*)c->d_recs.p, c->knn, rows, (l3d_match_rec*)c->d_csr.p);
    ++c->launches;
    L3D_CUDA(c, cudaGetLastError(), "k_compact_matches");
    L3D_CUDA(c, cudaMemcpyAsync(row_ptr_out, c->d_rowptr.p, sizeof(long long) * (rows + 1), cudaMemcpyDeviceToHost, c->stream), "download row_ptr");
    if (recs_out && total <= capacity && total > 0)
        L3D_CUDA(c, cudaMemcpyAsync(recs_out, c->d_csr.p, sizeof(l3d_match_rec) * (size_t)total, cudaMemcpyDeviceToHost, c->stream), "download records");
    L3D_CUDA(c, cudaStreamSynchronize(c->stream), "sync");
    return total;
}

// ------------------------------------------------------------------------------------------------ dense contract
static int match_dense_impl(l3d_ctx* c, int sv, int tv, const float* F, float epi, float* depths, float* overlaps, int on_dev, bool filter)
{
    if (!c) return L3D_ERR_INVALID;
    if (!c->have_views) return l3d_fail(c, L3D_ERR_STATE, "l3d_match_dense: call l3d_set_views first");
    if (!F || !depths || !overlaps || sv < 0 || tv < 0 || sv >= c->num_views || tv >= c->num_views) return l3d_fail(c, L3D_ERR_INVALID, "l3d_match_dense: bad arguments");
    cudaSetDevice(c->device);
    const L3DViewDev& vs = c->h_views[sv]; const L3DViewDev& vt = c->h_views[tv];
    const int Ns = vs.nseg, Nt = vt.nseg;
    if (Ns == 0 || Nt == 0) return L3D_OK;
    const size_t cells = (size_t)Ns * Nt;
    float4* d_dep = (float4*)depths; float* d_ov = overlaps;
    if (!on_dev) {
        int rc;
        if ((rc = l3d_reserve(c, c->d_dense_dep, sizeof(float4) * cells, "dense depths"))) return rc;
        if ((rc = l3d_reserve(c, c->d_dense_ov, sizeof(float) * cells, "dense overlaps"))) return rc;
        d_dep = (float4*)c->d_dense_dep.p; d_ov = (float*)c->d_dense_ov.p;
    }
    L3DMat3 Fm; memcpy(Fm.m, F, sizeof(Fm.m));
    dim3 grid((Nt + DK_THREADS - 1) / DK_THREADS, (Ns + DKN_ROWS - 1) / DKN_ROWS);
    const int R = l3d_dense_rows_per_cta(Ns, Nt, c->num_sms);
    dim3 gridf((Nt + DK_WARPS * DK_T * 32 - 1) / (DK_WARPS * DK_T * 32), (Ns + R - 1) / R);
    const float4* cache = (const float4*)c->d_cache.p;
    if (filter)
        k_match_dense<<<gridf, DK_THREADS, l3d_dense_smem_bytes(), c->stream>>>(c->segs() + vs.seg_off, Ns, c->segs() + vt.seg_off, Nt, cache + 3 * vs.seg_off, cache + 3 * vt.seg_off, Fm,
                                                          make_float3(vs.C[0], vs.C[1], vs.C[2]), make_float3(vt.C[0], vt.C[1], vt.C[2]), epi, d_dep, d_ov, R);
    else
        k_match_dense_nofilter<<<grid, DK_THREADS, 0, c->stream>>>(c->segs() + vs.seg_off, Ns, c->segs() + vt.seg_off, Nt, cache + 3 * vs.seg_off, cache + 3 * vt.seg_off, Fm,
                                                                   make_float3(vs.C[0], vs.C[1], vs.C[2]), make_float3(vt.C[0], vt.C[1], vt.C[2]), epi, d_dep, d_ov);
    ++c->launches;
    L3D_CUDA(c, cudaGetLastError(), "k_match_dense");
    if (!on_dev) {
        L3D_CUDA(c, cudaMemcpyAsync(depths, d_dep, sizeof(float4) * cells, cudaMemcpyDeviceToHost, c->stream), "download dense depths");
        L3D_CUDA(c, cudaMemcpyAsync(overlaps, d_ov, sizeof(float) * cells, cudaMemcpyDeviceToHost, c->stream), "download dense overlaps");
        L3D_CUDA(c, cudaStreamSynchronize(c->stream), "sync");
    }
    return L3D_OK;
}

int l3d_fp32_peak_probe(l3d_ctx* c, double* tflops_out)
{
    if (!c || !tflops_out) return L3D_ERR_INVALID;
    cudaSetDevice(c->device);
    int rc;
    if ((rc = l3d_reserve(c, c->d_scan_tmp, 256, "probe"))) return rc;
    const int iters = 4096, blocks = c->num_sms * 8;
    cudaEvent_t e0, e1;
    L3D_CUDA(c, cudaEventCreate(&e0), "event"); L3D_CUDA(c, cudaEventCreate(&e1), "event");
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, c->stream);
        k_fp32_peak<<<blocks, 256, 0, c->stream>>>((float*)c->d_scan_tmp.p, iters, 1.0000001f, 1e-9f);
        cudaEventRecord(e1, c->stream);
        ++c->launches;
        L3D_CUDA(c, cudaStreamSynchronize(c->stream), "k_fp32_peak");
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        double tf = 2.0 * 16.0 * iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *tflops_out = best;
    return L3D_OK;
}

int l3d_match_dense(l3d_ctx* c, int sv, int tv, const float* F, float epi, float* depths, float* overlaps, int on_dev)
{ return match_dense_impl(c, sv, tv, F, epi, depths, overlaps, on_dev, true); }

// K_match_lines' dense device contract for MANY view pairs in ONE launch; outputs are device pointers, one pair of arrays per view pair
int l3d_match_dense_pairs(l3d_ctx* c, int npairs, const int32_t* pairs, const float* F, float epi, float* const* depths_dev, float* const* overlaps_dev)
{
    if (!c) return L3D_ERR_INVALID;
    if (!c->have_views) return l3d_fail(c, L3D_ERR_STATE, "l3d_match_dense_pairs: call l3d_set_views first");
    if (npairs < 0 || (npairs && (!pairs || !F || !depths_dev || !overlaps_dev))) return l3d_fail(c, L3D_ERR_INVALID, "l3d_match_dense_pairs: bad arguments");
    if (npairs == 0) return L3D_OK;
    cudaSetDevice(c->device);
    std::vector<L3DDenseJob> jobs;
    const float4* cache = (const float4*)c->d_cache.p;
    long long tiles = 0;
    for (int i = 0; i < npairs; ++i) {
        const int sv = pairs[2 * i], tv = pairs[2 * i + 1];
        if (sv < 0 || tv < 0 || sv >= c->num_views || tv >= c->num_views || !depths_dev[i] || !overlaps_dev[i]) return l3d_fail(c, L3D_ERR_INVALID, "l3d_match_dense_pairs: bad pair");
        const L3DViewDev& vs = c->h_views[sv]; const L3DViewDev& vt = c->h_views[tv];
        if (vs.nseg == 0 || vt.nseg == 0) continue;
        L3DDenseJob J;
        J.ssegs = c->segs() + vs.seg_off; J.tsegs = c->segs() + vt.seg_off; J.scache = cache + 3 * vs.seg_off; J.tcache = cache + 3 * vt.seg_off;
        J.depths = (float4*)depths_dev[i]; J.overlaps = overlaps_dev[i];
        memcpy(J.F.m, F + 9 * (size_t)i, sizeof(J.F.m));
        J.Cs = make_float3(vs.C[0], vs.C[1], vs.C[2]); J.Ct = make_float3(vt.C[0], vt.C[1], vt.C[2]);
        J.Ns = vs.nseg; J.Nt = vt.nseg; J.rows_per_cta = DK_ROWS;       // whole waves do not matter inside a batch: the tallest tile amortises the per-tile set-up best
        J.colb = (vt.nseg + DK_WARPS * DK_T * 32 - 1) / (DK_WARPS * DK_T * 32);
        J.tile0 = tiles;
        tiles += (long long)J.colb * ((vs.nseg + DK_ROWS - 1) / DK_ROWS);
        jobs.push_back(J);
    }
    if (jobs.empty()) return L3D_OK;
    if (tiles >= (1ll << 31)) return l3d_fail(c, L3D_ERR_UNSUPPORTED, "l3d_match_dense_pairs: more than 2^31 tiles");
    int rc;
    if ((rc = l3d_reserve(c, c->d_dense_jobs, sizeof(L3DDenseJob) * jobs.size(), "dense jobs"))) return rc;
    L3D_CUDA(c, cudaMemcpyAsync(c->d_dense_jobs.p, jobs.data(), sizeof(L3DDenseJob) * jobs.size(), cudaMemcpyHostToDevice, c->stream), "dense jobs");
    L3D_CUDA(c, cudaStreamSynchronize(c->stream), "dense jobs");       // `jobs` is a local
    k_match_dense_batch<<<(unsigned int)tiles, DK_THREADS, l3d_dense_smem_bytes(), c->stream>>>((const L3DDenseJob*)c->d_dense_jobs.p, (int)jobs.size(), epi);
    ++c->launches;
    L3D_CUDA(c, cudaGetLastError(), "k_match_dense_batch");
    return L3D_OK;
}
int l3d_match_dense_nofilter(l3d_ctx* c, int sv, int tv, const float* F, float epi, float* depths, float* overlaps, int on_dev)
{ return match_dense_impl(c, sv, tv, F, epi, depths, overlaps, on_dev, false); }

} // extern "C"
