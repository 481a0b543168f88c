// oracle/ref_full_harness.cu — TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Builds the reference's WHOLE pipeline — line3D.cc + view.cc (+ cudawrapper.cu, sparsematrix.cc, clustering.cc) — VERBATIM
// from /root/reference (included by path, nothing copied) into
//     oracle/_ref/libl3dref_full_cpu.so   g++,  no L3DPP_CUDA: matchingCPU / scoringCPU / findCollinCPU   (runs anywhere)
//     oracle/_ref/libl3dref_full_gpu.so   nvcc, L3DPP_CUDA, -fmad=false: matchingGPU / scoringGPU / diffusion (GPU box)
// against the ~500-line stand-ins for Eigen / OpenCV / Boost in oracle/ref_shim (none of the three is installed).
// Built WITHOUT L3DPP_OPENMP: with OpenMP the reference's own result order depends on thread timing
// (estimated_position3D_.push_back under a mutex, line3D.cc:1639-1647), the single-threaded order is the canonical one.
//
// What this file adds is only a flat C ABI around the reference's public calls and read-only dumps of its private
// state (the headers are included with `private` spelled `public`; no reference line is changed).  State that the
// reference discards on the way (matches right after scoring, A_ before/after diffusion, local ids, clusters) is
// snapshotted from a std::cout stream buffer at the progress lines the reference prints at exactly those points:
//     "scoring: clusterable_segments"   line3D.cc:762  after scoring*(), before storeInverseMatches/filterMatches
//     "A: #entries="                    line3D.cc:1780 after computingAffinityMatrix()
//     "clustering segments..."          line3D.cc:1793 after performRDD() (or straight after the affinity matrix)
//     "computing final 3D lines..."     line3D.cc:1807 after clusterSegments() (+ optimizeClusters()), before clusters3D_.clear()
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <list>
#include <map>
#include <queue>
#include <set>
#include <sstream>
#include <streambuf>
#include <string>
#include <vector>
#include "eigen3/Eigen/Eigen"
#include "opencv2/core.hpp"
#include "opencv2/imgproc.hpp"
#include "boost/filesystem.hpp"
#include "boost/thread/mutex.hpp"

#define private public
#define protected public
#include "line3D.h"
#undef private
#undef protected

#include "line3D.cc"
#include "view.cc"
#ifdef L3DPP_CUDA
#include "cudawrapper.cu"
#include "sparsematrix.cc"
#endif
#include "clustering.cc"

// flat C types of the ABI (outside the anonymous namespace: nvcc gives functions with internal-linkage parameter types internal linkage)
struct rfl_match_t {                      // == orc_match_t / ref_match_t: flat L3DPP::Match (commons.h:186-203)
    uint32_t src_cam, src_seg, tgt_cam, tgt_seg;
    float overlap, score3D, d_p1, d_p2, d_q1, d_q2;
};
struct rfl_seg3d_t { int line; int pad; double p1[3], p2[3]; };
struct rfl_residual_t { int line; uint32_t cam, seg; };

namespace {

struct Edge { int i, j; float w; };

rfl_match_t flat(const L3DPP::Match& m)
{
    rfl_match_t o;
    o.src_cam = m.src_camID_; o.src_seg = m.src_segID_; o.tgt_cam = m.tgt_camID_; o.tgt_seg = m.tgt_segID_;
    o.overlap = m.overlap_score_; o.score3D = m.score3D_;
    o.d_p1 = m.depth_p1_; o.d_p2 = m.depth_p2_; o.d_q1 = m.depth_q1_; o.d_q2 = m.depth_q2_;
    return o;
}

struct Handle {
    L3DPP::Line3D* L = nullptr;
    // snapshots
    std::map<unsigned int, std::vector<rfl_match_t> > scored;
    size_t views_scored = 0;
    std::vector<Edge> A_raw, A_final;
    std::vector<std::pair<uint32_t, uint32_t> > l2g;
    std::vector<L3DPP::LineCluster3D> clusters;
};

Handle* g_active = nullptr;

void snap_edges(const std::list<L3DPP::CLEdge>& A, std::vector<Edge>& out)
{
    out.clear();
    out.reserve(A.size());
    for (std::list<L3DPP::CLEdge>::const_iterator it = A.begin(); it != A.end(); ++it) {
        Edge e; e.i = it->i_; e.j = it->j_; e.w = it->w_;
        out.push_back(e);
    }
}

void on_line(const std::string& s)
{
    Handle* H = g_active;
    if (!H || !H->L) return;
    L3DPP::Line3D* L = H->L;
    if (s.find("scoring: clusterable_segments") != std::string::npos) {
        // computeMatches walks visual_neighbors_ in key order (line3D.cc:704-705): this is the views_scored-th key
        std::map<unsigned int, std::set<unsigned int> >::const_iterator it = L->visual_neighbors_.begin();
        std::advance(it, H->views_scored);
        unsigned int cam = it->first;
        std::vector<rfl_match_t>& v = H->scored[cam];
        v.clear();
        const std::vector<std::list<L3DPP::Match> >& rows = L->matches_[cam];
        for (size_t i = 0; i < rows.size(); ++i)
            for (std::list<L3DPP::Match>::const_iterator m = rows[i].begin(); m != rows[i].end(); ++m) v.push_back(flat(*m));
        ++H->views_scored;
    } else if (s.find("A: #entries=") != std::string::npos) {
        snap_edges(L->A_, H->A_raw);
        H->l2g.clear();
        for (std::map<int, L3DPP::Segment2D>::const_iterator it = L->local2global_.begin(); it != L->local2global_.end(); ++it)
            H->l2g.push_back(std::make_pair(it->second.camID(), it->second.segID()));
    } else if (s.find("clustering segments...") != std::string::npos) {
        snap_edges(L->A_, H->A_final);
    } else if (s.find("computing final 3D lines...") != std::string::npos) {
        H->clusters = L->clusters3D_;
    }
}

// std::cout sink: cuts the stream into lines, calls on_line for each (the reference ends them with std::endl)
class HookBuf : public std::streambuf {
    std::string cur_;
    bool echo_;
public:
    HookBuf() : echo_(std::getenv("RFL_ECHO") != nullptr) {}
protected:
    int overflow(int c) override
    {
        if (c == EOF) return 0;
        if (c == '\n') {
            if (echo_) std::fprintf(stderr, "%s\n", cur_.c_str());
            on_line(cur_);
            cur_.clear();
        } else cur_.push_back(char(c));
        return c;
    }
};

HookBuf g_buf;
std::streambuf* g_prev = nullptr;

struct Scope {       // route std::cout through the hook for the duration of one call
    explicit Scope(Handle* H) { g_active = H; g_prev = std::cout.rdbuf(&g_buf); }
    ~Scope() { std::cout.rdbuf(g_prev); g_active = nullptr; }
};

Eigen::Matrix3d mat3(const double* m)
{
    Eigen::Matrix3d M;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M(r, c) = m[3 * r + c];
    return M;
}

} // namespace

extern "C" {

int rfl_has_cuda()
{
#ifdef L3DPP_CUDA
    return 1;
#else
    return 0;
#endif
}

void* rfl_create(const char* output_folder, int neighbors_by_worldpoints, int use_gpu)
{
    Handle* H = new Handle();
    Scope sc(H);
    // load_segments=false, max_img_width=-1 (no resize), max_line_segments default: irrelevant with explicit segments
    H->L = new L3DPP::Line3D(std::string(output_folder), false, -1, L3D_DEF_MAX_NUM_SEGMENTS, neighbors_by_worldpoints != 0, use_gpu != 0);
    return H;
}

void rfl_destroy(void* h)
{
    Handle* H = (Handle*)h;
    if (!H) return;
    { Scope sc(H); delete H->L; H->L = nullptr; }
    delete H;
}

// addImage with explicit line segments (line3D.cc:112-226)
int rfl_add_view(void* h, uint32_t cam, int width, int height, const double* K, const double* R, const double* t, float median_depth,
                 const uint32_t* wps_or_neighbors, int n_list, const float* segs, int nseg)
{
    Handle* H = (Handle*)h;
    Scope sc(H);
    cv::Mat img(height, width, CV_8U);
    std::list<unsigned int> lst(wps_or_neighbors, wps_or_neighbors + n_list);
    std::vector<cv::Vec4f> ls(nseg);
    for (int i = 0; i < nseg; ++i) ls[i] = cv::Vec4f(segs[4 * i], segs[4 * i + 1], segs[4 * i + 2], segs[4 * i + 3]);
    size_t before = H->L->views_.size();
    H->L->addImage(cam, img, mat3(K), mat3(R), Eigen::Vector3d(t[0], t[1], t[2]), median_depth, lst, ls);
    return H->L->views_.size() == before + 1 ? 0 : -1;
}

int rfl_match_images(void* h, float sigma_p, float sigma_a, uint32_t num_neighbors, float epi_overlap, int kNN, float const_reg_depth)
{
    Handle* H = (Handle*)h;
    Scope sc(H);
    H->scored.clear();
    H->views_scored = 0;
    H->L->matchImages(sigma_p, sigma_a, num_neighbors, epi_overlap, kNN, const_reg_depth);
    return 0;
}

int rfl_reconstruct_opt(void* h, uint32_t visibility_t, int perform_diffusion, float collinearity_t, int use_ceres, uint32_t max_iter_ceres)
{
    Handle* H = (Handle*)h;
    Scope sc(H);
    H->A_raw.clear(); H->A_final.clear(); H->l2g.clear(); H->clusters.clear();
    H->L->reconstruct3Dlines(visibility_t, perform_diffusion != 0, collinearity_t, use_ceres != 0, max_iter_ceres);
    return 0;
}

int rfl_num_views(void* h) { return int(((Handle*)h)->L->views_.size()); }

// matched (src,tgt) pairs: fundamentals_[src][tgt] is filled exactly once per matched pair with the matching direction
// (line3D.cc:861-897); map order = (src asc, tgt asc) = the order computeMatches visits them in.
int rfl_get_pairs(void* h, int* src_tgt, int cap)
{
    L3DPP::Line3D* L = ((Handle*)h)->L;
    int n = 0;
    for (std::map<unsigned int, std::map<unsigned int, Eigen::Matrix3d> >::const_iterator a = L->fundamentals_.begin(); a != L->fundamentals_.end(); ++a)
        for (std::map<unsigned int, Eigen::Matrix3d>::const_iterator b = a->second.begin(); b != a->second.end(); ++b) {
            if (n < cap) { src_tgt[2 * n] = int(a->first); src_tgt[2 * n + 1] = int(b->first); }
            ++n;
        }
    return n;
}

int rfl_get_fundamental(void* h, uint32_t src, uint32_t tgt, double* F9)
{
    L3DPP::Line3D* L = ((Handle*)h)->L;
    if (L->fundamentals_.count(src) == 0 || L->fundamentals_[src].count(tgt) == 0) return -1;
    const Eigen::Matrix3d& F = L->fundamentals_[src][tgt];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) F9[3 * r + c] = F(r, c);
    return 0;
}

int rfl_get_neighbors(void* h, uint32_t cam, uint32_t* out, int cap)
{
    L3DPP::Line3D* L = ((Handle*)h)->L;
    if (L->visual_neighbors_.count(cam) == 0) return -1;
    int n = 0;
    for (std::set<unsigned int>::const_iterator it = L->visual_neighbors_[cam].begin(); it != L->visual_neighbors_[cam].end(); ++it, ++n)
        if (n < cap) out[n] = *it;
    return n;
}

// matches_[cam] as it stands (after matchImages: the filtered matches), row by row in list order
long long rfl_get_matches(void* h, uint32_t cam, rfl_match_t* out, long long cap)
{
    L3DPP::Line3D* L = ((Handle*)h)->L;
    if (L->matches_.count(cam) == 0) return -1;
    const std::vector<std::list<L3DPP::Match> >& rows = L->matches_[cam];
    long long n = 0;
    for (size_t i = 0; i < rows.size(); ++i)
        for (std::list<L3DPP::Match>::const_iterator m = rows[i].begin(); m != rows[i].end(); ++m, ++n)
            if (n < cap) out[n] = flat(*m);
    return n;
}

long long rfl_get_scored(void* h, uint32_t cam, rfl_match_t* out, long long cap)
{
    Handle* H = (Handle*)h;
    if (H->scored.count(cam) == 0) return -1;
    const std::vector<rfl_match_t>& v = H->scored[cam];
    for (long long i = 0; i < (long long)v.size() && i < cap; ++i) out[i] = v[i];
    return (long long)v.size();
}

int rfl_get_view_info(void* h, uint32_t cam, float* k, float* median_depth)
{
    L3DPP::Line3D* L = ((Handle*)h)->L;
    if (L->views_.count(cam) == 0) return -1;
    *k = L->views_[cam]->k();
    *median_depth = L->views_[cam]->median_depth();
    return 0;
}

// View geometry as the reference derived it (view.cc:6-42): RtKinv (row-major), C, k
int rfl_get_view_geometry(void* h, uint32_t cam, double* RtKinv9, double* C3)
{
    L3DPP::Line3D* L = ((Handle*)h)->L;
    if (L->views_.count(cam) == 0) return -1;
    Eigen::Matrix3d M = L->views_[cam]->RtKinv();
    Eigen::Vector3d Cc = L->views_[cam]->C();
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) RtKinv9[3 * r + c] = M(r, c); C3[r] = Cc(r); }
    return 0;
}

long long rfl_get_estimates(void* h, rfl_match_t* best, double* p1p2, long long cap)
{
    L3DPP::Line3D* L = ((Handle*)h)->L;
    long long n = (long long)L->estimated_position3D_.size();
    for (long long i = 0; i < n && i < cap; ++i) {
        const L3DPP::Segment3D& s = L->estimated_position3D_[i].first;
        if (best) best[i] = flat(L->estimated_position3D_[i].second);
        if (p1p2) {
            Eigen::Vector3d a = s.P1(), b = s.P2();
            for (int c = 0; c < 3; ++c) { p1p2[6 * i + c] = a(c); p1p2[6 * i + 3 + c] = b(c); }
        }
    }
    return n;
}

static long long put_edges(const std::vector<Edge>& A, int* ei, int* ej, float* ew, long long cap)
{
    for (long long i = 0; i < (long long)A.size() && i < cap; ++i) { ei[i] = A[i].i; ej[i] = A[i].j; ew[i] = A[i].w; }
    return (long long)A.size();
}
long long rfl_get_affinity(void* h, int* ei, int* ej, float* ew, long long cap) { return put_edges(((Handle*)h)->A_final, ei, ej, ew, cap); }
long long rfl_get_affinity_raw(void* h, int* ei, int* ej, float* ew, long long cap) { return put_edges(((Handle*)h)->A_raw, ei, ej, ew, cap); }

int rfl_get_local2global(void* h, uint32_t* cam_seg, int cap)
{
    Handle* H = (Handle*)h;
    for (int i = 0; i < (int)H->l2g.size() && i < cap; ++i) { cam_seg[2 * i] = H->l2g[i].first; cam_seg[2 * i + 1] = H->l2g[i].second; }
    return (int)H->l2g.size();
}

// clusters3D_ right before computeFinal3Dsegments: 3D line (6 doubles), number of residuals, reference view
int rfl_get_clusters(void* h, double* p1p2, int* nres, uint32_t* ref_view, int cap)
{
    Handle* H = (Handle*)h;
    for (int i = 0; i < (int)H->clusters.size() && i < cap; ++i) {
        L3DPP::Segment3D s = H->clusters[i].seg3D();
        Eigen::Vector3d a = s.P1(), b = s.P2();
        for (int c = 0; c < 3; ++c) { p1p2[6 * i + c] = a(c); p1p2[6 * i + 3 + c] = b(c); }
        nres[i] = (int)H->clusters[i].size();
        ref_view[i] = H->clusters[i].reference_view();
    }
    return (int)H->clusters.size();
}

// View::collin_ as CSR (view.cc:153-263)
long long rfl_get_collinear(void* h, uint32_t cam, long long* row_ptr, int* idx, long long cap)
{
    L3DPP::Line3D* L = ((Handle*)h)->L;
    if (L->views_.count(cam) == 0) return -1;
    L3DPP::View* v = L->views_[cam];
    long long n = 0;
    size_t N = v->num_lines();
    for (size_t r = 0; r < N; ++r) {
        row_ptr[r] = n;
        if (r < v->collin_.size())
            for (std::list<unsigned int>::const_iterator it = v->collin_[r].begin(); it != v->collin_[r].end(); ++it, ++n)
                if (idx && n < cap) idx[n] = int(*it);
    }
    row_ptr[N] = n;
    return n;
}

int rfl_num_lines(void* h)
{
    std::vector<L3DPP::FinalLine3D> res;
    Handle* H = (Handle*)h;
    Scope sc(H);
    H->L->get3Dlines(res);
    return (int)res.size();
}

long long rfl_get_segments3d(void* h, rfl_seg3d_t* out, long long cap)
{
    Handle* H = (Handle*)h;
    Scope sc(H);
    std::vector<L3DPP::FinalLine3D> res;
    H->L->get3Dlines(res);
    long long n = 0;
    for (size_t i = 0; i < res.size(); ++i)
        for (std::list<L3DPP::Segment3D>::const_iterator it = res[i].collinear3Dsegments_.begin(); it != res[i].collinear3Dsegments_.end(); ++it, ++n)
            if (out && n < cap) {
                out[n].line = (int)i; out[n].pad = 0;
                Eigen::Vector3d a = it->P1(), b = it->P2();
                for (int c = 0; c < 3; ++c) { out[n].p1[c] = a(c); out[n].p2[c] = b(c); }
            }
    return n;
}

long long rfl_get_residuals(void* h, rfl_residual_t* out, long long cap)
{
    Handle* H = (Handle*)h;
    Scope sc(H);
    std::vector<L3DPP::FinalLine3D> res;
    H->L->get3Dlines(res);
    long long n = 0;
    for (size_t i = 0; i < res.size(); ++i) {
        const std::list<L3DPP::Segment2D>* r = res[i].underlyingCluster_.residuals();
        for (std::list<L3DPP::Segment2D>::const_iterator it = r->begin(); it != r->end(); ++it, ++n)
            if (out && n < cap) { out[n].line = (int)i; out[n].cam = it->camID(); out[n].seg = it->segID(); }
    }
    return n;
}

// the reference's own writers into `folder`; name_out receives createOutputFilename()
int rfl_save(void* h, const char* folder, int txt, int obj, int stl, char* name_out, int name_cap)
{
    Handle* H = (Handle*)h;
    Scope sc(H);
    if (txt) H->L->save3DLinesAsTXT(folder);
    if (obj) H->L->saveResultAsOBJ(folder);
    if (stl) H->L->saveResultAsSTL(folder);
    std::string n = H->L->createOutputFilename();
    if (name_out && name_cap > 0) { std::strncpy(name_out, n.c_str(), name_cap - 1); name_out[name_cap - 1] = 0; }
    return 0;
}

} // extern "C"
