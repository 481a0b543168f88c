// oracle/ref_harness.cu — TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Builds the UNMODIFIED reference accelerator code into oracle/_ref/libl3dref_*.so by including the
// reference translation units *where they lie* under /root/reference (the Makefile passes
// -I/root/reference -Ioracle/ref_shim; no reference source is copied into this repository):
//     cudawrapper.cu   (5 kernels + 4 host wrappers, cudawrapper.cu:186-766)
//     sparsematrix.cc  (COO float4 matrix, sparsematrix.cc:8-149)
//     clustering.cc    (Felzenszwalb union-find, clustering.cc:6-48)
// and exposes a flat C ABI that the parity tests / bench drive through ctypes.  What this file adds is
// ONLY the host staging the reference does in line3D.cc around each call:
//     matchingGPU   line3D.cc:1040-1074  (F / RtKinv as pitched 3x3 DataArray<float>, eigen2dataArray 2775-2781)
//     scoringGPU    line3D.cc:1357-1368  (upload ranges/matches/reg_tgt, launch, download scores)
//     performRDD    line3D.cc:2026-2036  (SparseMatrix(A_, n) -> replicator_dynamics_diffusion_GPU -> download)
//     clusterSegments line3D.cc:2089     (performClustering(A_, n, 3.0f))
//     findCollinGPU   view.cc:173-209    (char N x N buffer -> find_collinear_segments_GPU -> download)
#include "cudawrapper.cu"
#include "sparsematrix.cc"
#include "clustering.cc"

#include <chrono>
#include <cstring>

namespace {

// 3x3 row-major host matrix -> pitched DataArray<float>, element (c,r) = M(r,c)   (line3D.cc:2775-2781, view.cc:37-40)
L3DPP::DataArray<float>* mat3_to_da(const float* m)
{
    L3DPP::DataArray<float>* da = new L3DPP::DataArray<float>(3, 3);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            da->dataCPU(c, r)[0] = m[r * 3 + c];
    da->upload();
    return da;
}

L3DPP::DataArray<float4>* lines_to_da(const float* xyzw, int n)
{
    L3DPP::DataArray<float4>* da = new L3DPP::DataArray<float4>(n, 1);
    for (int i = 0; i < n; ++i)
        da->dataCPU(i, 0)[0] = make_float4(xyzw[4 * i], xyzw[4 * i + 1], xyzw[4 * i + 2], xyzw[4 * i + 3]);
    da->upload();
    return da;
}

} // namespace

extern "C" {

// flat mirror of L3DPP::Match (commons.h:186-203), 40 bytes
struct ref_match_t {
    unsigned int src_cam, src_seg, tgt_cam, tgt_seg;
    float overlap, score3D, d_p1, d_p2, d_q1, d_q2;
};

int ref_device_count()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

// Verbatim K_match_lines over the full Ns x Nt grid (one launch; the reference chunks rows only to bound its
// scratch buffer, cudawrapper.cu:570-586).  Outputs are dense row-major [Ns][Nt].
int ref_match_dense(const float* lines_src, int Ns, const float* lines_tgt, int Nt,
                    const float* F, const float* RtKinv_src, const float* RtKinv_tgt,
                    const float* C_src, const float* C_tgt, float epi_overlap,
                    float* depths_out /*4*Ns*Nt*/, float* overlaps_out /*Ns*Nt*/, float* kernel_ms)
{
    L3DPP::DataArray<float4>* ls = lines_to_da(lines_src, Ns);
    L3DPP::DataArray<float4>* lt = lines_to_da(lines_tgt, Nt);
    L3DPP::DataArray<float>* dF = mat3_to_da(F);
    L3DPP::DataArray<float>* dRs = mat3_to_da(RtKinv_src);
    L3DPP::DataArray<float>* dRt = mat3_to_da(RtKinv_tgt);
    L3DPP::DataArray<float4>* buffer = new L3DPP::DataArray<float4>(Nt, Ns, true);
    L3DPP::DataArray<float>* overlaps = new L3DPP::DataArray<float>(Nt, Ns, true);

    dim3 dimBlock(L3DPP::L3D_BLOCK_SIZE, L3DPP::L3D_BLOCK_SIZE);
    dim3 dimGrid(L3DPP::divUp(Nt, dimBlock.x), L3DPP::divUp(Ns, dimBlock.y));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    L3DPP::K_match_lines<<<dimGrid, dimBlock>>>(Nt, Ns, 0, buffer->dataGPU(), buffer->strideGPU(),
                                               overlaps->dataGPU(), overlaps->strideGPU(),
                                               ls->dataGPU(), lt->dataGPU(), dF->dataGPU(), dRs->dataGPU(),
                                               dRt->dataGPU(), dF->strideGPU(),
                                               make_float3(C_src[0], C_src[1], C_src[2]),
                                               make_float3(C_tgt[0], C_tgt[1], C_tgt[2]), epi_overlap);
    cudaEventRecord(e1);
    cudaError_t st = cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (kernel_ms) *kernel_ms = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (depths_out && overlaps_out) {
        buffer->download();
        overlaps->download();
        for (int r = 0; r < Ns; ++r)
            for (int c = 0; c < Nt; ++c) {
                float4 d = buffer->dataCPU(c, r)[0];
                float* o = depths_out + 4 * ((size_t)r * Nt + c);
                o[0] = d.x; o[1] = d.y; o[2] = d.z; o[3] = d.w;
                overlaps_out[(size_t)r * Nt + c] = overlaps->dataCPU(c, r)[0];
            }
    }
    delete ls; delete lt; delete dF; delete dRs; delete dRt; delete buffer; delete overlaps;
    return st == cudaSuccess ? 0 : -(int)st;
}

// Verbatim find_collinear_segments_GPU (cudawrapper.cu:689-705) staged like View::findCollinGPU (view.cc:173-209):
// C_out[i*N + c] = buffer->dataCPU(c,i)[0], the char the reference tests against 1 when it builds collin_[i].
// (every cell is written: thread (x,y), x >= y, stores C[y][x] and its mirror C[x][y], cudawrapper.cu:376-427)
int ref_collinear(const float* lines, int N, float dist_t, unsigned char* C_out, float* kernel_ms)
{
    L3DPP::DataArray<float4>* ls = lines_to_da(lines, N);
    L3DPP::DataArray<char>* buffer = new L3DPP::DataArray<char>(N, N, true);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    L3DPP::find_collinear_segments_GPU(buffer, ls, dist_t);
    cudaEventRecord(e1);
    cudaError_t st = cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (kernel_ms) *kernel_ms = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (C_out) {
        buffer->download();
        for (int i = 0; i < N; ++i)
            for (int c = 0; c < N; ++c) C_out[(size_t)i * N + c] = (unsigned char)buffer->dataCPU(c, i)[0];
    }
    delete ls; delete buffer;
    return st == cudaSuccess ? 0 : -(int)st;
}

// Verbatim match_lines_GPU (kernel + dense D2H + host kNN pass), staged like matchingGPU.
// Fills out[row*cap + i] for the i-th match of src row `row` in the reference's list order; counts[row] = list length
// (matches beyond `cap` are dropped from `out` but still counted).  Returns the reference's own return value.
// wall_ms = host wall clock of the whole reference call incl. uploads (that *is* the reference path).
long long ref_match_lines(const float* lines_src, int Ns, const float* lines_tgt, int Nt,
                          const float* F, const float* RtKinv_src, const float* RtKinv_tgt,
                          const float* C_src, const float* C_tgt, unsigned int srcCamID, unsigned int tgtCamID,
                          float epi_overlap, int kNN, int* counts, ref_match_t* out, int cap, double* wall_ms)
{
    L3DPP::DataArray<float4>* ls = lines_to_da(lines_src, Ns);   // initSrcDataGPU (line3D.cc:1018-1026)
    L3DPP::DataArray<float>* dRs = mat3_to_da(RtKinv_src);
    std::vector<std::list<L3DPP::Match> > matches(Ns);

    auto t0 = std::chrono::steady_clock::now();
    L3DPP::DataArray<float4>* lt = lines_to_da(lines_tgt, Nt);   // matchingGPU (line3D.cc:1049-1057)
    L3DPP::DataArray<float>* dF = mat3_to_da(F);
    L3DPP::DataArray<float>* dRt = mat3_to_da(RtKinv_tgt);
    unsigned int n = L3DPP::match_lines_GPU(ls, lt, dF, dRs, dRt,
                                            make_float3(C_src[0], C_src[1], C_src[2]),
                                            make_float3(C_tgt[0], C_tgt[1], C_tgt[2]),
                                            &matches, srcCamID, tgtCamID, epi_overlap, kNN);
    delete lt; delete dF; delete dRt;                            // line3D.cc:1069-1071
    auto t1 = std::chrono::steady_clock::now();
    if (wall_ms) *wall_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();

    for (int r = 0; r < Ns; ++r) {
        int i = 0;
        for (std::list<L3DPP::Match>::const_iterator it = matches[r].begin(); it != matches[r].end(); ++it, ++i) {
            if (i < cap && out) {
                ref_match_t& o = out[(size_t)r * cap + i];
                o.src_cam = it->src_camID_; o.src_seg = it->src_segID_;
                o.tgt_cam = it->tgt_camID_; o.tgt_seg = it->tgt_segID_;
                o.overlap = it->overlap_score_; o.score3D = it->score3D_;
                o.d_p1 = it->depth_p1_; o.d_p2 = it->depth_p2_; o.d_q1 = it->depth_q1_; o.d_q2 = it->depth_q2_;
            }
        }
        if (counts) counts[r] = i;
    }
    delete ls; delete dRs;
    return (long long)n;
}

// Verbatim score_matches_GPU (cudawrapper.cu:661-684) with the upload/download scoringGPU does around it.
int ref_score_matches(const float* lines, int Ns, const float* matches /*4*M: seg,tgtCam,d1,d2*/, int M,
                      const int* ranges /*2*Ns*/, const float* reg_tgt /*2*M*/, const float* RtKinv, const float* C,
                      float two_sigA_sqr, float k, float min_similarity, float* scores_out, float* kernel_ms)
{
    L3DPP::DataArray<float4>* dl = lines_to_da(lines, Ns);
    L3DPP::DataArray<float>* dR = mat3_to_da(RtKinv);
    L3DPP::DataArray<int2>* dr = new L3DPP::DataArray<int2>(Ns, 1);
    for (int i = 0; i < Ns; ++i) dr->dataCPU(i, 0)[0] = make_int2(ranges[2 * i], ranges[2 * i + 1]);
    L3DPP::DataArray<float4>* dm = new L3DPP::DataArray<float4>(M, 1);
    L3DPP::DataArray<float2>* dg = new L3DPP::DataArray<float2>(M, 1);
    L3DPP::DataArray<float>* ds = new L3DPP::DataArray<float>(M, 1, true);
    for (int i = 0; i < M; ++i) {
        dm->dataCPU(i, 0)[0] = make_float4(matches[4 * i], matches[4 * i + 1], matches[4 * i + 2], matches[4 * i + 3]);
        dg->dataCPU(i, 0)[0] = make_float2(reg_tgt[2 * i], reg_tgt[2 * i + 1]);
    }
    dr->upload(); dm->upload(); dg->upload();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    L3DPP::score_matches_GPU(dl, dm, dr, ds, dg, dR, make_float3(C[0], C[1], C[2]), two_sigA_sqr, k, min_similarity);
    cudaEventRecord(e1);
    cudaError_t st = cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (kernel_ms) *kernel_ms = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    ds->download();
    for (int i = 0; i < M; ++i) scores_out[i] = ds->dataCPU(i, 0)[0];
    delete dl; delete dR; delete dr; delete dm; delete dg; delete ds;
    return st == cudaSuccess ? 0 : -(int)st;
}

// Verbatim SparseMatrix(A_, n) + replicator_dynamics_diffusion_GPU + download (performRDD, line3D.cc:2030-2036).
// Output: row-sorted COO of the diffused matrix P, out_i/out_j/out_w each of length nedges.
int ref_rdd(int nedges, const int* ei, const int* ej, const float* ew, int n,
            int* out_i, int* out_j, float* out_w, double* wall_ms)
{
    std::list<L3DPP::CLEdge> A;
    for (int e = 0; e < nedges; ++e) {
        L3DPP::CLEdge ed; ed.i_ = ei[e]; ed.j_ = ej[e]; ed.w_ = ew[e];
        A.push_back(ed);
    }
    auto t0 = std::chrono::steady_clock::now();
    L3DPP::SparseMatrix* W = new L3DPP::SparseMatrix(A, n);
    std::streambuf* old = std::cout.rdbuf(nullptr);   // the reference prints one line per iteration
    L3DPP::replicator_dynamics_diffusion_GPU(W, std::string(""));
    std::cout.rdbuf(old);
    W->download();
    auto t1 = std::chrono::steady_clock::now();
    if (wall_ms) *wall_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    for (unsigned int i = 0; i < W->entries()->width(); ++i) {
        float4 e = W->entries()->dataCPU(i, 0)[0];
        out_i[i] = (int)e.x; out_j[i] = (int)e.y; out_w[i] = e.z;
    }
    delete W;
    return 0;
}

// Verbatim performClustering (clustering.cc:6-48) + CLUniverse::find for every node (clusterSegments, line3D.cc:2102).
int ref_cluster(int nedges, const int* ei, const int* ej, const float* ew, int n, float c, int* labels_out)
{
    std::list<L3DPP::CLEdge> A;
    for (int e = 0; e < nedges; ++e) {
        L3DPP::CLEdge ed; ed.i_ = ei[e]; ed.j_ = ej[e]; ed.w_ = ew[e];
        A.push_back(ed);
    }
    L3DPP::CLUniverse* u = L3DPP::performClustering(A, n, c);
    if (!u) return -1;
    for (int i = 0; i < n; ++i) labels_out[i] = u->find(i);
    delete u;
    return 0;
}

} // extern "C"
