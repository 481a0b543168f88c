// line3d_host.cc — host side of L3DPP::Line3D (include/line3d.h) on top of the C ABI (include/l3d_capi.h).
//
// What stays on the host is what line3D.cc does around its accelerator calls: camera algebra (View::View, view.cc:6-42),
// translate / untranslate (line3D.cc:500-575), spatial regulariser k (view.cc:301-314), visual neighbours
// (line3D.cc:578-699), the view-pair list and fundamental matrices (702-741, 861-897), then after the device stages the
// "unused"/local-id bookkeeping of the affinity matrix (1881-1900, 1982-2023), the symmetrisation after diffusion
// (2036-2071), graph clustering (clustering.cc:6-48) and the cluster -> 3D segment tail (2079-2452) + writers.
// All matching / scoring / affinity / diffusion arithmetic runs on the GPU; without a GPU the constructor fails.
#include "../../include/line3d.h"
#include <nvtx3/nvToolsExt.h>      // header-only; ranges show up in nsys / ncu timelines
#include "../../include/l3d_capi.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <mutex>
#include <set>
#include <sstream>
#include <stdexcept>
#include <thread>
#include <unordered_map>
#include <unordered_set>

namespace L3DPP {

// Independent per-cluster host work (the reference runs the same loops under `#pragma omp parallel for`, line3D.cc:2125, 2283):
// static partition over a few std::threads, results stored by index, so the output order does not depend on timing.
template <class F> static void parallel_for(size_t n, F&& body)
{
    unsigned int nt = std::min<unsigned int>(std::max(1u, std::thread::hardware_concurrency()), 32u);
    if (const char* e = getenv("L3D_HOST_THREADS")) nt = (unsigned int)std::max(1, atoi(e));
    if (n < 256 || nt <= 1) { for (size_t i = 0; i < n; ++i) body(i); return; }
    nt = (unsigned int)std::min<size_t>(nt, n / 64);
    std::vector<std::thread> th;
    for (unsigned int t = 0; t < nt; ++t)
        th.emplace_back([&, t]() { for (size_t i = n * t / nt; i < n * (t + 1) / nt; ++i) body(i); });
    for (auto& x : th) x.join();
}


namespace {
const double EPS = 1e-12;                    // L3D_EPS commons.h:92
const float MIN_SIMILARITY_3D = 0.50f, MIN_BEST_SCORE_3D = 0.75f, MIN_BEST_SCORE_PERC = 0.10f, MIN_AFFINITY = 0.50f;   // commons.h:58-60,68
const int RDD_MAX_ITER = 10;                 // commons.h:65

inline Vector3d operator+(const Vector3d& a, const Vector3d& b) { return Vector3d(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vector3d operator-(const Vector3d& a, const Vector3d& b) { return Vector3d(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vector3d operator*(const Vector3d& a, double s) { return Vector3d(a.x * s, a.y * s, a.z * s); }
inline double dot(const Vector3d& a, const Vector3d& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double norm(const Vector3d& a) { return std::sqrt(dot(a, a)); }
inline Vector3d unit(const Vector3d& a) { double n2 = dot(a, a); if (n2 > 0) { double n = std::sqrt(n2); return Vector3d(a.x / n, a.y / n, a.z / n); } return a; }
inline Matrix3d matmul(const Matrix3d& a, const Matrix3d& b)
{ Matrix3d r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += a(i, k) * b(k, j); r(i, j) = s; } return r; }
inline Vector3d matvec(const Matrix3d& a, const Vector3d& v)
{ return Vector3d(a(0, 0) * v.x + a(0, 1) * v.y + a(0, 2) * v.z, a(1, 0) * v.x + a(1, 1) * v.y + a(1, 2) * v.z, a(2, 0) * v.x + a(2, 1) * v.y + a(2, 2) * v.z); }
inline Matrix3d transposed(const Matrix3d& a) { Matrix3d r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = a(j, i); return r; }
Matrix3d inverted(const Matrix3d& a)   // adjugate / determinant
{
    Matrix3d r;
    const double c00 = a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1), c01 = a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2), c02 = a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0);
    const double id = 1.0 / (a(0, 0) * c00 + a(0, 1) * c01 + a(0, 2) * c02);
    r(0, 0) = c00 * id; r(1, 0) = c01 * id; r(2, 0) = c02 * id;
    r(0, 1) = (a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2)) * id; r(1, 1) = (a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0)) * id; r(2, 1) = (a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1)) * id;
    r(0, 2) = (a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1)) * id; r(1, 2) = (a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2)) * id; r(2, 2) = (a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0)) * id;
    return r;
}

struct HostView {
    unsigned int id = 0;
    int width = 0, height = 0;
    Matrix3d K, R, RtKinv;
    Vector3d t, C, pp;
    float C_f3[3], RtKinv_f[9];      // frozen at construction like view.cc:35-40 (NOT updated by translate)
    float k = 0.f, median_depth = 0.f, min_line_length = 0.f;
    std::vector<Vec4f> lines;
    std::list<unsigned int> wps_or_neighbors;
    long long seg_off = 0;           // global segment index of segment 0 (views in ascending camID)

    Vector3d ray(double x, double y) const { return unit(matvec(RtKinv, Vector3d(x, y, 1.0))); }
    void shift(const Vector3d& d) { C = C + d; t = matvec(R, C) * -1.0; }             // View::translate view.cc:510-514
    void project(const Vector3d& P, double& u, double& v) const                       // View::project view.cc:374-393
    {
        Vector3d q = matvec(R, P) + t;
        const double xn = (1.0 * q.x + 0.0 * q.z) / q.z, yn = (1.0 * q.y + 0.0 * q.z) / q.z;
        Vector3d h = matvec(K, Vector3d(xn, yn, 1));
        u = h.x / h.z; v = h.y / h.z;
    }
};

struct Edge { int i, j; float w; };   // CLEdge clustering.h:47-51

// Felzenszwalb-Huttenlocher graph segmentation as used by the reference (clustering.cc:6-48, universe.h:49-104):
// edges visited by ascending weight (stable), components merge while w <= both thresholds, threshold = w + c/|C|.
class DisjointSets {
public:
    explicit DisjointSets(int n) : parent_(n), rank_(n, 0), size_(n, 1) { for (int i = 0; i < n; ++i) parent_[i] = i; }
    int find(int x) { int root = x; while (root != parent_[root]) root = parent_[root]; parent_[x] = root; return root; }   // single-step compression like universe.h:70-78
    int unite(int a, int b)
    {
        if (rank_[a] > rank_[b]) { parent_[b] = a; size_[a] += size_[b]; return a; }
        parent_[a] = b; size_[b] += size_[a];
        if (rank_[a] == rank_[b]) ++rank_[b];
        return b;
    }
    int size(int x) const { return size_[x]; }
private:
    std::vector<int> parent_, rank_, size_;
};

// `order` (optional): indices of `edges` by ascending weight, ties in list order (from l3d_argsort_f32 for big lists)
std::vector<int> segment_graph(const std::vector<Edge>& edges, int n, float c, const std::vector<unsigned int>* order)
{
    std::vector<unsigned int> local;
    if (!order) {
        local.resize(edges.size());
        for (size_t i = 0; i < local.size(); ++i) local[i] = (unsigned int)i;
        std::stable_sort(local.begin(), local.end(), [&](unsigned int a, unsigned int b) { return edges[a].w < edges[b].w; });
        order = &local;
    }
    DisjointSets u(n);
    std::vector<float> thr(n, c);
    for (unsigned int idx : *order) {
        const Edge& e = edges[idx];
        int a = u.find(e.i), b = u.find(e.j);
        if (a == b || !(e.w <= thr[a] && e.w <= thr[b])) continue;
        u.unite(a, b);
        a = u.find(a);
        thr[a] = e.w + c / (float)u.size(a);
    }
    std::vector<int> label(n);
    for (int i = 0; i < n; ++i) label[i] = u.find(i);
    return label;
}

// dominant eigenvector of a symmetric positive semi-definite 3x3 matrix by cyclic Jacobi rotations; plays the role of
// the JacobiSVD of the scatter matrix in get3DlineFromCluster (line3D.cc:2196-2211)
Vector3d principal_axis(double A[3][3])
{
    double Q[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 60; ++sweep) {
        const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (std::fabs(A[p][q]) < 1e-300) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double tt = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double cs = 1.0 / std::sqrt(tt * tt + 1.0), sn = tt * cs;
                for (int k = 0; k < 3; ++k) { const double x = A[k][p], y = A[k][q]; A[k][p] = cs * x - sn * y; A[k][q] = sn * x + cs * y; }
                for (int k = 0; k < 3; ++k) { const double x = A[p][k], y = A[q][k]; A[p][k] = cs * x - sn * y; A[q][k] = sn * x + cs * y; }
                for (int k = 0; k < 3; ++k) { const double x = Q[k][p], y = Q[k][q]; Q[k][p] = cs * x - sn * y; Q[k][q] = sn * x + cs * y; }
            }
    }
    int m = 0;
    if (A[1][1] > A[m][m]) m = 1;
    if (A[2][2] > A[m][m]) m = 2;
    return unit(Vector3d(Q[0][m], Q[1][m], Q[2][m]));
}
}  // namespace

// ------------------------------------------------------------------------------------------------ Segment3D
Segment3D::Segment3D() : P1_(0, 0, 0), P2_(0, 0, 0), dir_(0, 0, 0), length_(0.0f), valid_(false) {}
Segment3D::Segment3D(const Vector3d& P1, const Vector3d& P2)
{
    length_ = (float)norm(P1 - P2);
    if (length_ > EPS) { P1_ = P1; P2_ = P2; dir_ = unit(P2 - P1); valid_ = true; }
    else { P1_ = P2_ = dir_ = Vector3d(0, 0, 0); length_ = 0.0f; valid_ = false; }
}
float Segment3D::distance_Point2Line(const Vector3d& P) const
{
    const Vector3d w = P - P1_;
    const double d[3] = {dir_.x, dir_.y, dir_.z}, ww[3] = {w.x, w.y, w.z};
    double h[3];
    for (int i = 0; i < 3; ++i) h[i] = (d[i] * ww[0]) * d[0] + (d[i] * ww[1]) * d[1] + (d[i] * ww[2]) * d[2];
    return (float)norm(Vector3d(P1_.x + h[0], P1_.y + h[1], P1_.z + h[2]) - P);
}
void Segment3D::translate(const Vector3d& t) { P1_ = P1_ + t; P2_ = P2_ + t; }

// ------------------------------------------------------------------------------------------------ Line3D::Impl
struct Line3D::Impl {
    l3d_ctx* ctx = nullptr;
    std::string output_folder, err;
    bool load_segments, by_wps, use_gpu, verbose = false;
    int max_img_width; unsigned int max_line_segments;
    std::mutex mtx;                                       // addImage is thread-safe like line3D.cc:129-226
    std::map<unsigned int, HostView> views;               // ascending camID == processing order (line3D.cc:704)
    std::vector<float> views_avg_depths;
    std::map<unsigned int, std::list<unsigned int> > wps2views;
    std::map<unsigned int, std::set<unsigned int> > visual_neighbors;
    // parameters (matchImages / reconstruct3Dlines)
    unsigned int num_neighbors = 10, visibility_t = 3;
    float sigma_p = 2.5f, sigma_a = 10.f, two_sigA_sqr = 200.f, epi = 0.25f, const_reg_depth = -1.f, med_scene_depth = -1.f,
          med_scene_depth_lines = 0.f, collin_t = -1.f;
    bool use_ceres = false;
    int kNN = 10;
    bool fixed3Dreg = false, perform_RDD = false, matched = false;
    Vector3d translation;
    int shard_rank = 0, shard_world = 1; Line3D::MatchExchangeFn exchange = nullptr; void* exchange_user = nullptr;
    // index maps (rebuilt at matchImages)
    std::vector<HostView*> vlist;                         // by view index (ascending camID)
    std::map<unsigned int, int> index_of;
    std::vector<int32_t> pairs;                           // (src idx, tgt idx) in match order
    // results
    std::vector<l3d_match> est_best; std::vector<double> est_P; std::vector<long long> est_gseg;
    std::unordered_map<long long, size_t> est_of_gseg;
    std::vector<Edge> A, A_raw; std::vector<long long> local2global;
    std::vector<FinalLine3D> lines3D;
    Line3DStats st;

    void fail(const std::string& what) { err = what; if (verbose) std::cerr << "[L3D++/B200] ERROR: " << what << std::endl; }
    bool chk(long long rc, const char* what) { if (rc < 0) { fail(std::string(what) + ": " + l3d_last_error(ctx)); return false; } return true; }
    void log(const std::string& s) { if (verbose) std::cout << "[L3D++/B200] " << s << std::endl; }

    void shift_all(const Vector3d& d)                     // performTranslation line3D.cc:548-575
    {
        for (auto& kv : views) kv.second.shift(d);
        for (FinalLine3D& L : lines3D) {
            for (Segment3D& s : L.collinear3Dsegments_) s.translate(d);
            L.underlyingCluster_.translate(d);
        }
    }
    void translate()                                      // line3D.cc:500-536
    {
        if (views.empty()) return;
        double tr[3] = {0, 0, 0};
        for (int a = 0; a < 3; ++a) {
            std::vector<double> c;
            for (auto& kv : views) { const double v = a == 0 ? kv.second.C.x : (a == 1 ? kv.second.C.y : kv.second.C.z); if (std::fabs(v) > EPS) c.push_back(v); }
            if (!c.empty()) { std::sort(c.begin(), c.end()); tr[a] = c[c.size() / 2]; }
        }
        translation = Vector3d(tr[0], tr[1], tr[2]);
        shift_all(translation * -1.0);
    }
    void untranslate() { shift_all(translation); }        // line3D.cc:539-545

    float spatial_reg(const HostView& v, float r) const   // View::getSpecificSpatialReg view.cc:307-314
    {
        const Vector3d a = v.ray(v.pp.x, v.pp.y), b = v.ray(v.pp.x + r, v.pp.y);
        return (float)std::sin(std::acos(std::fmin(std::fmax(dot(a, b), -1.0), 1.0)));
    }
    void neighbors_from_wps(unsigned int camID);
    Matrix3d fundamental(const HostView& s, const HostView& t) const   // getFundamentalMatrix line3D.cc:874-892
    {
        const Matrix3d R = matmul(t.R, transposed(s.R));
        const Vector3d tt = t.t - matvec(R, s.t);
        Matrix3d T; T(0, 0) = 0; T(0, 1) = -tt.z; T(0, 2) = tt.y; T(1, 0) = tt.z; T(1, 1) = 0; T(1, 2) = -tt.x; T(2, 0) = -tt.y; T(2, 1) = tt.x; T(2, 2) = 0;
        return matmul(matmul(inverted(transposed(t.K)), matmul(T, R)), inverted(s.K));
    }
    std::vector<l3d_view_desc> descs() const
    {
        std::vector<l3d_view_desc> d(vlist.size());
        for (size_t i = 0; i < vlist.size(); ++i) {
            const HostView& v = *vlist[i];
            d[i].cam_id = v.id; d[i].width = v.width; d[i].height = v.height; d[i].nseg = (int32_t)v.lines.size();
            memcpy(d[i].RtKinv, v.RtKinv_f, sizeof(v.RtKinv_f)); memcpy(d[i].C, v.C_f3, sizeof(v.C_f3));
            memcpy(d[i].RtKinv_d, v.RtKinv.m, sizeof(v.RtKinv.m));
            d[i].C_d[0] = v.C.x; d[i].C_d[1] = v.C.y; d[i].C_d[2] = v.C.z;
            d[i].k = v.k; d[i].median_depth = v.median_depth;
        }
        return d;
    }
    Segment2D seg_of(long long g) const
    {
        auto it = std::upper_bound(vlist.begin(), vlist.end(), g, [](long long x, const HostView* v) { return x < v->seg_off; });
        const HostView* v = *(it - 1);
        return Segment2D(v->id, (unsigned int)(g - v->seg_off));
    }
    const Segment3D estimate(const Segment2D& s) const
    {
        const HostView* v = vlist[index_of.at(s.camID())];
        const size_t e = est_of_gseg.at(v->seg_off + s.segID());
        return Segment3D(Vector3d(est_P[6 * e], est_P[6 * e + 1], est_P[6 * e + 2]), Vector3d(est_P[6 * e + 3], est_P[6 * e + 4], est_P[6 * e + 5]));
    }
    bool line_from_cluster(const std::list<Segment2D>& cluster, LineCluster3D& out) const;
    Segment3D project_onto_line(const Segment2D& s2, const Segment3D& s3, bool& ok) const;
    std::list<Segment3D> collinear_segments(const LineCluster3D& cl) const;
};

// findVisualNeighborsFromWPs (line3D.cc:578-699)
void Line3D::Impl::neighbors_from_wps(unsigned int camID)
{
    std::set<unsigned int>& out = visual_neighbors[camID];
    out.clear();
    HostView& v = views[camID];
    std::map<unsigned int, unsigned int> common;
    for (unsigned int wp : v.wps_or_neighbors)
        for (unsigned int other : wps2views[wp]) if (other != camID) ++common[other];
    if (common.empty()) return;
    struct Cand { unsigned int cam; float score, axis, dist; };
    std::vector<Cand> cand;
    for (auto& kv : common) {
        const HostView& o = views[kv.first];
        Cand c; c.cam = kv.first;
        c.score = 2.0f * float(kv.second) / float(v.wps_or_neighbors.size() + o.wps_or_neighbors.size());
        c.axis = (float)std::acos(std::fmin(std::fmax(dot(v.ray(v.pp.x, v.pp.y), o.ray(o.pp.x, o.pp.y)), -1.0), 1.0));      // opticalAxesAngle view.cc:457-463
        const Vector3d ct = matvec(v.R, o.C) + v.t;                                                                        // distanceVisualNeighborScore view.cc:487-501
        c.dist = std::fabs((float)ct.x) + std::fabs((float)ct.y);
        if (c.axis < 1.571f && kv.second > 4) cand.push_back(c);
    }
    std::stable_sort(cand.begin(), cand.end(), [](const Cand& a, const Cand& b) { return a.score > b.score; });
    if (cand.size() > num_neighbors) {
        const std::vector<Cand> all = cand;
        const float score_t = 0.80f * cand.front().score;
        size_t nbig = 0;
        while (nbig < cand.size() && cand[nbig].score > score_t) ++nbig;
        cand.resize(nbig);
        std::stable_sort(cand.begin(), cand.end(), [](const Cand& a, const Cand& b) { return a.dist > b.dist; });
        if (cand.size() > num_neighbors / 2) cand.resize(num_neighbors / 2);
        cand.insert(cand.end(), all.begin(), all.end());
    }
    const float min_baseline = 0.1f;
    auto baseline = [&](unsigned int a, unsigned int b) { return (float)norm(views[a].C - views[b].C); };
    for (size_t i = 0; i < cand.size() && out.size() < num_neighbors; ++i) {
        const unsigned int n = cand[i].cam;
        if (out.count(n) || !(baseline(camID, n) > min_baseline)) continue;
        bool valid = true;
        for (unsigned int u : out) if (!(baseline(camID, u) > min_baseline)) { valid = false; break; }
        if (valid) out.insert(n);
    }
}

// ------------------------------------------------------------------------------------------------ Line3D
Line3D::Line3D(const std::string& output_folder, const bool load_segments, const int max_img_width, const unsigned int max_line_segments,
               const bool neighbors_by_worldpoints, const bool use_GPU, const int cuda_device)
    : p_(new Impl())
{
    p_->output_folder = output_folder; p_->load_segments = load_segments; p_->max_img_width = max_img_width;
    p_->max_line_segments = max_line_segments; p_->by_wps = neighbors_by_worldpoints; p_->use_gpu = use_GPU;
    memset(&p_->st, 0, sizeof(p_->st));
    const int rc = l3d_ctx_create(cuda_device, &p_->ctx);
    if (rc != L3D_OK) { delete p_; p_ = nullptr; throw std::runtime_error("L3DPP::Line3D (B200): no usable CUDA device - this build has no CPU path"); }
}
Line3D::~Line3D() { if (p_) { l3d_ctx_destroy(p_->ctx); delete p_; } }
const Line3DStats& Line3D::stats() const { return p_->st; }
const char* Line3D::lastError() const { return p_->err.c_str(); }
void Line3D::setVerbose(bool v) { p_->verbose = v; }
void Line3D::setShard(int rank, int world, MatchExchangeFn fn, void* user)
{
    Impl& P = *p_;
    std::lock_guard<std::mutex> g(P.mtx);
    P.err.clear();
    if (world < 1 || rank < 0 || rank >= world || (world > 1 && !fn)) { P.fail("setShard: need 0 <= rank < world and an exchange function"); return; }
    P.shard_rank = rank; P.shard_world = world; P.exchange = fn; P.exchange_user = user;
}
size_t Line3D::numImages() { std::lock_guard<std::mutex> g(p_->mtx); return p_->views.size(); }

bool Line3D::addImage(const unsigned int camID, const int w, const int h, const Matrix3d& K, const Matrix3d& R, const Vector3d& t,
                      const float median_depth, const std::list<unsigned int>& wps_or_neighbors, const std::vector<Vec4f>& line_segments)
{
    Impl& P = *p_;
    // everything that needs no shared state first (the reference's frontends call addImage from an OpenMP loop) ...
    HostView v;
    v.id = camID; v.width = w; v.height = h; v.K = K; v.R = R; v.t = t; v.lines = line_segments; v.wps_or_neighbors = wps_or_neighbors;
    v.min_line_length = std::sqrt(float((unsigned)w * (unsigned)w + (unsigned)h * (unsigned)h)) * 0.005f;                           // view.cc:17-18
    v.pp = Vector3d(K(0, 2), K(1, 2), 1.0);
    const Matrix3d Rt = transposed(R);
    v.RtKinv = matmul(Rt, inverted(K));
    v.C = matvec(Rt, t * -1.0);
    v.C_f3[0] = (float)v.C.x; v.C_f3[1] = (float)v.C.y; v.C_f3[2] = (float)v.C.z;
    for (int i = 0; i < 9; ++i) v.RtKinv_f[i] = (float)v.RtKinv.m[i];
    // ... then one critical section for the checks, the insertion and the error string.  A failure is reported through the return
    // value AND stays in lastError() (a later successful addImage does not clear it) until the next non-addImage call.
    std::lock_guard<std::mutex> g(P.mtx);
    auto failed = [&](const std::string& what) { P.fail("addImage [" + std::to_string(camID) + "]: " + what); return false; };
    if (std::max(w, h) < 800) return failed("image is too small for reliable results (larger side should be >= 800px)");   // line3D.cc:119
    if (wps_or_neighbors.empty()) return failed(P.by_wps ? "view has no worldpoints" : "view has no visual neighbors");    // line3D.cc:154
    if (line_segments.empty()) return failed("no line segments given (LSD detection is outside this library's scope)");
    if (P.views.count(camID)) return failed("camera ID already in use");                                                     // line3D.cc:130
    if (P.by_wps) for (unsigned int wp : wps_or_neighbors) P.wps2views[wp].push_back(camID);
    P.views[camID] = v;
    P.visual_neighbors[camID];
    P.views_avg_depths.push_back((float)std::fmax((double)median_depth, EPS));
    return true;
}

void Line3D::matchImages(const float sigma_position, const float sigma_angle, const unsigned int num_neighbors, const float epipolar_overlap,
                         const int kNN, const float const_regularization_depth)
{
    Impl& P = *p_;
    std::lock_guard<std::mutex> g(P.mtx);
    P.err.clear();
    if (P.views.empty()) { P.fail("no images to match"); return; }
    // parameter checks: line3D.cc:394-413
    P.num_neighbors = (unsigned int)std::max(int(num_neighbors), 2);
    P.sigma_p = sigma_position; P.sigma_a = std::fmin(std::fabs(sigma_angle), 90.0f);
    P.two_sigA_sqr = 2.0f * P.sigma_a * P.sigma_a;
    P.epi = std::fmin(std::fabs(epipolar_overlap), 0.99f); P.kNN = kNN; P.const_reg_depth = const_regularization_depth;
    if (P.sigma_p < 0.0f) { P.fixed3Dreg = true; P.sigma_p = std::fabs(P.sigma_p); } else { P.fixed3Dreg = false; P.sigma_p = std::fmax(0.1f, P.sigma_p); }
    P.med_scene_depth = const_regularization_depth;
    if (const_regularization_depth < 0.0f && P.fixed3Dreg && !P.views_avg_depths.empty()) {
        std::sort(P.views_avg_depths.begin(), P.views_avg_depths.end());
        P.med_scene_depth = P.views_avg_depths[P.views_avg_depths.size() / 2];
    }
    P.matched = false; P.est_best.clear(); P.est_P.clear(); P.est_gseg.clear(); P.est_of_gseg.clear();
    P.translate();
    P.vlist.clear(); P.index_of.clear();
    long long off = 0;
    for (auto& kv : P.views) {
        HostView& v = kv.second;
        v.k = P.fixed3Dreg ? P.sigma_p / P.med_scene_depth : P.spatial_reg(v, P.sigma_p);
        v.seg_off = off; off += (long long)v.lines.size();
        P.index_of[v.id] = (int)P.vlist.size(); P.vlist.push_back(&v);
    }
    // visual neighbours (line3D.cc:463-485)
    for (auto& kv : P.views) {
        if (!P.by_wps) {
            std::set<unsigned int>& vn = P.visual_neighbors[kv.first];
            if (vn.empty()) for (unsigned int n : kv.second.wps_or_neighbors) if (P.views.count(n)) vn.insert(n);
        } else P.neighbors_from_wps(kv.first);
    }
    // view pairs in the reference's order (computeMatches line3D.cc:704-741) + float fundamental matrices
    P.pairs.clear();
    std::vector<float> F; std::vector<double> Fdbl;
    std::set<std::pair<unsigned int, unsigned int> > done;
    for (auto& kv : P.visual_neighbors)
        for (unsigned int n : kv.second) {
            const unsigned int s = kv.first;
            if (n == s || done.count(std::make_pair(std::min(s, n), std::max(s, n)))) continue;
            done.insert(std::make_pair(std::min(s, n), std::max(s, n)));
            P.pairs.push_back(P.index_of[s]); P.pairs.push_back(P.index_of[n]);
            const Matrix3d Fd = P.fundamental(P.views[s], P.views[n]);
            for (int i = 0; i < 9; ++i) { F.push_back((float)Fd.m[i]); Fdbl.push_back(Fd.m[i]); }   // eigen2dataArray line3D.cc:2775-2781
        }
    const int npairs = (int)(P.pairs.size() / 2);
    // ---- device: upload, match all pairs, scoring sweep
    std::vector<l3d_view_desc> d = P.descs();
    std::vector<const float*> segp(P.vlist.size());
    for (size_t i = 0; i < P.vlist.size(); ++i) segp[i] = &P.vlist[i]->lines[0].v[0];
    auto t0 = std::chrono::steady_clock::now();
    nvtxRangePushA("l3d:set_views+match");
    bool ok = P.chk(l3d_set_views(P.ctx, (int)d.size(), d.data(), segp.data()), "l3d_set_views");
    // use_GPU selects the reference's semantics, not the processor (line3D.cc:708-757): true = K_match_lines / K_score_matches
    // float arithmetic, false = matchingCPU / scoringCPU double arithmetic.  Both run on the B200.
    auto match_range = [&](int first, int last) {
        return P.use_gpu ? P.chk(l3d_match_pairs_range(P.ctx, npairs, P.pairs.data(), F.data(), P.epi, P.kNN, first, last), "l3d_match_pairs_range")
                         : P.chk(l3d_match_pairs_f64(P.ctx, npairs, P.pairs.data(), Fdbl.data(), P.epi, P.kNN, first, last), "l3d_match_pairs_f64");
    };
    if (ok && P.shard_world > 1 && (P.kNN <= 0 || P.kNN > 32)) { P.fail("matchImages: kNN <= 0 (keep all matches) or kNN > 32 cannot be sharded: the row stride is only known after matching"); ok = false; }
    if (ok && P.shard_world > 1) {
        // this rank's contiguous share of the pair list, balanced by Ns*Nt; the same split on every rank
        std::vector<long long> cost((size_t)npairs), row_off((size_t)npairs + 1), row_bounds((size_t)P.shard_world + 1);
        for (int i = 0; i < npairs; ++i) cost[i] = (long long)P.vlist[P.pairs[2 * i]]->lines.size() * (long long)P.vlist[P.pairs[2 * i + 1]]->lines.size();
        std::vector<int32_t> bounds((size_t)P.shard_world + 1);
        ok = P.chk(l3d_balanced_split(cost.data(), npairs, P.shard_world, bounds.data()), "l3d_balanced_split") &&
             match_range(bounds[P.shard_rank], bounds[P.shard_rank + 1]) &&
             P.chk(l3d_sync(P.ctx), "l3d_sync") && P.chk(l3d_pair_row_offsets(P.ctx, row_off.data()), "l3d_pair_row_offsets");
        void *dc = nullptr, *dr = nullptr;
        ok = ok && P.chk(l3d_match_device_buffers(P.ctx, &dc, &dr), "l3d_match_device_buffers");
        if (ok) {
            for (int r = 0; r <= P.shard_world; ++r) row_bounds[r] = row_off[bounds[r]];
            if (P.exchange(P.exchange_user, dc, dr, row_bounds.data(), P.shard_world, P.kNN) != 0) { P.fail("matchImages: the match exchange callback failed"); ok = false; }
        }
    } else
        ok = ok && match_range(0, npairs) && P.chk(l3d_sync(P.ctx), "l3d_sync");
    auto t1 = std::chrono::steady_clock::now();
    nvtxRangePop(); nvtxRangePushA("l3d:score_sweep");
    ok = ok && P.chk(l3d_score_sweep(P.ctx, P.two_sigA_sqr, MIN_SIMILARITY_3D, MIN_BEST_SCORE_3D, MIN_BEST_SCORE_PERC), "l3d_score_sweep");
    nvtxRangePop();
    auto t2 = std::chrono::steady_clock::now();
    if (ok) {
        const long long ne = l3d_get_estimates(P.ctx, nullptr, nullptr, 0);
        ok = P.chk(ne, "l3d_get_estimates");
        if (ok) {
            P.est_best.resize(ne); P.est_P.resize(6 * (size_t)ne);
            if (ne) ok = P.chk(l3d_get_estimates(P.ctx, P.est_best.data(), P.est_P.data(), ne), "l3d_get_estimates");
        }
    }
    if (ok) {
        // filterMatches' tail (line3D.cc:1657-1668): per-view median of the best matches' depths
        std::map<unsigned int, std::vector<float> > depths;
        P.est_gseg.resize(P.est_best.size());
        for (size_t e = 0; e < P.est_best.size(); ++e) {
            const l3d_match& m = P.est_best[e];
            depths[m.src_cam].push_back(m.d_p1); depths[m.src_cam].push_back(m.d_p2);
            P.est_gseg[e] = P.views[m.src_cam].seg_off + m.src_seg;
            P.est_of_gseg[P.est_gseg[e]] = e;
        }
        for (auto& kv : P.views) {
            float med = (float)EPS;
            std::vector<float>& dv = depths[kv.first];
            if (!dv.empty()) { std::sort(dv.begin(), dv.end()); med = dv[dv.size() / 2]; }
            kv.second.median_depth = med;                                               // update_median_depth view.h:108-121
            if (P.fixed3Dreg) kv.second.k = P.sigma_p / P.med_scene_depth;
        }
        P.st.view_pairs = npairs; P.st.pair_evaluations = l3d_match_pair_evals(P.ctx); P.st.estimates = (long long)P.est_best.size();
        P.st.matches_after_knn = l3d_match_total_matches(P.ctx);      // reduced on the device (was: download of the whole count array)
        P.st.ms_match = std::chrono::duration<double, std::milli>(t1 - t0).count();
        P.st.ms_score = std::chrono::duration<double, std::milli>(t2 - t1).count();
        P.matched = true;
    }
    P.untranslate();
}

// get3DlineFromCluster (line3D.cc:2155-2218)
bool Line3D::Impl::line_from_cluster(const std::list<Segment2D>& cluster, LineCluster3D& out) const
{
    Vector3d cog(0, 0, 0);
    std::vector<Vector3d> pts;
    unsigned int reference_cam = 0; float max_len_2D = 0.0f;
    for (const Segment2D& s : cluster) {
        const Segment3D h = estimate(s);
        cog = cog + h.P1(); cog = cog + h.P2();
        pts.push_back(h.P1()); pts.push_back(h.P2());
        const Vec4f c = vlist[index_of.at(s.camID())]->lines[s.segID()];
        const float l2 = (c.v[0] - c.v[2]) * (c.v[0] - c.v[2]) + (c.v[1] - c.v[3]) * (c.v[1] - c.v[3]);
        if (l2 > max_len_2D) { max_len_2D = l2; reference_cam = s.camID(); }
    }
    const double n = double(pts.size());
    cog = Vector3d(cog.x / n, cog.y / n, cog.z / n);
    double Sc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (const Vector3d& p : pts) {
        const double d[3] = {p.x - cog.x, p.y - cog.y, p.z - cog.z};
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) Sc[a][b] += d[a] * d[b];
    }
    const Vector3d dir = principal_axis(Sc);
    out = LineCluster3D(Segment3D(cog - dir, cog + dir), cluster, reference_cam);
    return out.size() > 0;
}

// project2DsegmentOnto3Dline (line3D.cc:2221-2266): closest points between the 3D line and the two viewing rays
Segment3D Line3D::Impl::project_onto_line(const Segment2D& s2, const Segment3D& s3, bool& ok) const
{
    const HostView& v = *vlist[index_of.at(s2.camID())];
    const Vec4f l = v.lines[s2.segID()];
    const Vector3d Pl = s3.P1(), u = s3.dir(), w = Pl - v.C;
    const Vector3d v1 = v.ray(l.v[0], l.v[1]), v2 = v.ray(l.v[2], l.v[3]);
    const double a = dot(u, u), b1 = dot(u, v1), b2 = dot(u, v2), c1 = dot(v1, v1), c2 = dot(v2, v2), d = dot(u, w), e1 = dot(v1, w), e2 = dot(v2, w);
    const double den1 = a * c1 - b1 * b1, den2 = a * c2 - b2 * b2;
    ok = std::fabs(den1) > EPS && std::fabs(den2) > EPS;
    if (!ok) return Segment3D();
    return Segment3D(Pl + u * ((b1 * e1 - c1 * d) / den1), Pl + u * ((b2 * e2 - c2 * d) / den2));
}

// findCollinearSegments(cluster) (line3D.cc:2342-2452): sweep along the line, a 3D segment is open while >= 3 cameras see it
std::list<Segment3D> Line3D::Impl::collinear_segments(const LineCluster3D& cl) const
{
    std::list<Segment3D> out;
    const Vector3d cog = (cl.seg3D().P1() + cl.seg3D().P2()) * 0.5;
    struct Pt { size_t line, point, cam; float dist; };
    std::vector<Pt> ev;
    std::vector<Vector3d> pts(cl.residuals()->size() * 2);
    float far = 0.0f; Vector3d border(0, 0, 0);
    size_t id = 0, pid = 0;
    for (const Segment2D& s : *cl.residuals()) {
        bool ok; const Segment3D pr = project_onto_line(s, cl.seg3D(), ok);
        if (ok) {
            const Vector3d e[2] = {pr.P1(), pr.P2()};
            for (int k = 0; k < 2; ++k) {
                pts[pid + k] = e[k];
                Pt p = {id, pid + k, s.camID(), 0.0f}; ev.push_back(p);
                const float dd = (float)norm(e[k] - cog);
                if (dd > far) { far = dd; border = e[k]; }
            }
        }
        ++id; pid += 2;
    }
    if (ev.size() < 6) return out;
    for (Pt& p : ev) p.dist = (float)norm(pts[p.point] - border);
    std::stable_sort(ev.begin(), ev.end(), [](const Pt& a, const Pt& b) { return a.dist < b.dist; });
    std::map<size_t, unsigned int> open_cams; std::set<size_t> open_lines;
    bool opened = false; Vector3d start(0, 0, 0);
    for (const Pt& p : ev) {
        if (!open_lines.count(p.line)) { open_lines.insert(p.line); ++open_cams[p.cam]; }
        else { open_lines.erase(p.line); if (--open_cams[p.cam] == 0) open_cams.erase(p.cam); }
        if (opened && open_cams.size() < 3) { out.push_back(Segment3D(start, pts[p.point])); opened = false; }
        else if (!opened && open_cams.size() >= 3) { start = pts[p.point]; opened = true; }
    }
    return out;
}

void Line3D::reconstruct3Dlines(const unsigned int visibility_t, const bool perform_diffusion, const float collinearity_t, const bool use_CERES,
                                const unsigned int max_iter_CERES)
{
    Impl& P = *p_;
    std::lock_guard<std::mutex> g(P.mtx);
    P.err.clear();
    if (!P.matched || P.est_best.empty()) { P.fail("no clusterable segments! forgot to match lines?"); return; }   // line3D.cc:1712-1718
    P.visibility_t = (unsigned int)std::max(int(visibility_t), 3);
    P.lines3D.clear(); P.collin_t = collinearity_t;
    P.perform_RDD = perform_diffusion && P.use_gpu;                                       // line3D.cc:1729
    P.translate();
    // median scene depth for lines (line3D.cc:1759-1774)
    std::vector<float> sd;
    for (auto& kv : P.views) if (kv.second.median_depth > EPS) sd.push_back(kv.second.median_depth);
    if (!sd.empty()) { std::sort(sd.begin(), sd.end()); P.med_scene_depth_lines = sd[sd.size() / 2]; } else P.med_scene_depth_lines = 0.0f;
    std::vector<l3d_view_desc> d = P.descs();
    auto t0 = std::chrono::steady_clock::now();
    bool ok = P.chk(l3d_update_view_params(P.ctx, (int)d.size(), d.data()), "l3d_update_view_params");
    // potentially collinear segments of every view (findCollinearSegments line3D.cc:1751-1756, 1827-1849); the affinity matrix
    // below then adds the collinearity links (line3D.cc:1904-1974).  collinearity_t <= eps switches them off.
    ok = ok && P.chk(l3d_find_collinear(P.ctx, collinearity_t > EPS ? collinearity_t : 0.0f, P.use_gpu ? L3D_SEM_REF_GPU : L3D_SEM_REF_CPU), "l3d_find_collinear");
    P.st.collinear_entries = ok ? l3d_collinear_total(P.ctx) : 0;
    // affinity matrix incl. the "unused" pair filter and first-come local ids (line3D.cc:1881-1900, 1982-2023), on the device
    long long nids = 0;
    long long nA = ok ? l3d_affinity_matrix(P.ctx, P.two_sigA_sqr, P.med_scene_depth_lines, MIN_AFFINITY, nullptr, nullptr, nullptr, 0, nullptr, 0, &nids) : -1;
    ok = ok && P.chk(nA, "l3d_affinity_matrix");
    P.A.clear(); P.local2global.clear();
    if (ok && nA > 0) {
        std::vector<int> ai((size_t)nA), aj((size_t)nA); std::vector<float> aw((size_t)nA);
        P.local2global.resize((size_t)nids);
        ok = P.chk(l3d_affinity_matrix(P.ctx, P.two_sigA_sqr, P.med_scene_depth_lines, MIN_AFFINITY, ai.data(), aj.data(), aw.data(), nA,
                                       P.local2global.data(), nids, &nids), "l3d_affinity_matrix");
        P.A.resize((size_t)nA);
        for (long long e = 0; e < nA; ++e) { P.A[e].i = ai[e]; P.A[e].j = aj[e]; P.A[e].w = aw[e]; }
    }
    if (!ok) { P.untranslate(); return; }
    P.A_raw = P.A;
    auto t1 = std::chrono::steady_clock::now();
    const int n = (int)P.local2global.size();
    P.st.affinity_entries = (long long)P.A.size(); P.st.affinity_rows = n;
    // diffusion (performRDD line3D.cc:2026-2076)
    if (P.perform_RDD && !P.A.empty()) {
        // the matrix is still on the device (l3d_affinity_matrix): diffusion + min(w12, w21) there, one download of the result in the
        // (i, j) order the reference's std::map rebuilds A_ in (line3D.cc:2039-2071)
        const long long nnz = (long long)P.A.size();
        std::vector<int> oi(nnz), oj(nnz); std::vector<float> ow(nnz);
        if (!P.chk(l3d_rdd_affinity(P.ctx, RDD_MAX_ITER, oi.data(), oj.data(), ow.data(), nnz), "l3d_rdd_affinity")) { P.untranslate(); return; }
        for (long long e = 0; e < nnz; ++e) { P.A[e].i = oi[e]; P.A[e].j = oj[e]; P.A[e].w = ow[e]; }
    }
    auto t2 = std::chrono::steady_clock::now();
    // clustering (clusterSegments line3D.cc:2079-2152)
    std::vector<LineCluster3D> clusters;
    P.st.clusters_total = 0;
    if (!P.A.empty()) {
        std::vector<unsigned int> worder;
        if (P.A.size() > (1u << 20)) {          // big list: sort the weights on the device (stable, like std::list::sort)
            std::vector<float> w(P.A.size());
            for (size_t e = 0; e < w.size(); ++e) w[e] = P.A[e].w;
            worder.resize(w.size());
            if (!P.chk(l3d_argsort_f32(P.ctx, (long long)w.size(), w.data(), worder.data()), "l3d_argsort_f32")) { P.untranslate(); return; }
        }
        const std::vector<int> label = segment_graph(P.A, n, 3.0f, worder.empty() ? nullptr : &worder);
        std::map<int, std::list<Segment2D> > members; std::map<int, std::set<unsigned int> > cams; std::vector<int> order;
        for (int id = 0; id < n; ++id) {
            const int cl = label[id];
            if (!members.count(cl)) order.push_back(cl);
            const Segment2D s = P.seg_of(P.local2global[id]);
            members[cl].push_back(s); cams[cl].insert(s.camID());
        }
        P.st.clusters_total = (long long)order.size();
        std::vector<const std::list<Segment2D>*> todo;
        for (int cl : order) if (cams[cl].size() >= P.visibility_t) todo.push_back(&members[cl]);
        std::vector<LineCluster3D> fitted(todo.size()); std::vector<unsigned char> good(todo.size(), 0);
        parallel_for(todo.size(), [&](size_t i) { good[i] = P.line_from_cluster(*todo[i], fitted[i]) ? 1 : 0; });
        for (size_t i = 0; i < todo.size(); ++i) if (good[i]) clusters.push_back(fitted[i]);
    }
    P.st.clusters_valid = (long long)clusters.size();
    // optimizeClusters (line3D.cc:1800-1805, 2269-2275): bundle the cluster lines against their 2D residuals.  The reference
    // needs Ceres for this (optimization.cc); here the same problem is minimised on the GPU (l3d_optimize_lines).
    P.st.opt_iterations = 0; P.st.opt_cost_before = P.st.opt_cost_after = 0.0;
    if (use_CERES && !clusters.empty()) {
        const int Lc = (int)clusters.size(), nc = (int)P.vlist.size();
        std::vector<double> p((size_t)6 * Lc), cams((size_t)16 * nc), xy;
        std::vector<long long> ptr((size_t)Lc + 1, 0);
        std::vector<int> rcam;
        for (int v = 0; v < nc; ++v) {
            const HostView& hv = *P.vlist[v];
            double* cm = &cams[(size_t)16 * v];
            for (int r = 0; r < 3; ++r) for (int c2 = 0; c2 < 3; ++c2) cm[3 * r + c2] = hv.R(r, c2);
            cm[9] = hv.C.x; cm[10] = hv.C.y; cm[11] = hv.C.z;
            cm[12] = hv.K(0, 0); cm[13] = hv.K(1, 1); cm[14] = hv.K(0, 2); cm[15] = hv.K(1, 2);
        }
        for (int i = 0; i < Lc; ++i) {
            const Segment3D s3 = clusters[i].seg3D();
            p[6 * i] = s3.P1().x; p[6 * i + 1] = s3.P1().y; p[6 * i + 2] = s3.P1().z; p[6 * i + 3] = s3.P2().x; p[6 * i + 4] = s3.P2().y; p[6 * i + 5] = s3.P2().z;
            for (const Segment2D& s2 : *clusters[i].residuals()) {
                const int vi = P.index_of.at(s2.camID());
                const Vec4f& ln = P.vlist[vi]->lines[s2.segID()];
                rcam.push_back(vi);
                for (int k = 0; k < 4; ++k) xy.push_back((double)ln.v[k]);
            }
            ptr[i + 1] = (long long)rcam.size();
        }
        std::vector<int> valid((size_t)Lc);
        double summ[8];
        if (!P.chk(l3d_optimize_lines(P.ctx, Lc, p.data(), ptr.data(), rcam.data(), xy.data(), nc, cams.data(), (int)max_iter_CERES, p.data(), valid.data(), summ),
                   "l3d_optimize_lines")) { P.untranslate(); return; }
        std::vector<LineCluster3D> kept;
        for (int i = 0; i < Lc; ++i) {
            if (!valid[i]) continue;                        // optimization.cc:293-298
            clusters[i].update3Dline(Segment3D(Vector3d(p[6 * i], p[6 * i + 1], p[6 * i + 2]), Vector3d(p[6 * i + 3], p[6 * i + 4], p[6 * i + 5])));
            kept.push_back(clusters[i]);
        }
        clusters.swap(kept);
        P.st.opt_iterations = (long long)summ[0]; P.st.opt_cost_before = summ[1]; P.st.opt_cost_after = summ[2];
    }
    P.use_ceres = use_CERES;
    // computeFinal3Dsegments + filterTinySegments (line3D.cc:2278-2339)
    std::vector<std::list<Segment3D> > kept_segs(clusters.size());
    parallel_for(clusters.size(), [&](size_t i) {
        const LineCluster3D& cl = clusters[i];
        std::list<Segment3D> col = P.collinear_segments(cl);
        if (col.empty()) return;
        const HostView& rv = *P.vlist[P.index_of.at(cl.reference_view())];
        for (const Segment3D& s : col) {
            double u1, v1, u2, v2; rv.project(s.P1(), u1, v1); rv.project(s.P2(), u2, v2);
            if (std::sqrt((u1 - u2) * (u1 - u2) + (v1 - v2) * (v1 - v2)) > rv.min_line_length) kept_segs[i].push_back(s);  // projectedLongEnough view.cc:423-428
        }
    });
    for (size_t i = 0; i < clusters.size(); ++i) {
        if (kept_segs[i].empty()) continue;
        FinalLine3D f; f.collinear3Dsegments_ = kept_segs[i]; f.underlyingCluster_ = clusters[i];
        P.lines3D.push_back(f);
    }
    auto t3 = std::chrono::steady_clock::now();
    P.st.lines3D = (long long)P.lines3D.size();
    P.st.ms_affinity = std::chrono::duration<double, std::milli>(t1 - t0).count();
    P.st.ms_diffusion = std::chrono::duration<double, std::milli>(t2 - t1).count();
    P.st.ms_cluster = std::chrono::duration<double, std::milli>(t3 - t2).count();
    P.untranslate();
}

void Line3D::get3Dlines(std::vector<FinalLine3D>& result) { std::lock_guard<std::mutex> g(p_->mtx); result = p_->lines3D; }

Vector4f Line3D::getSegmentCoords2D(const unsigned int camID, const unsigned int segID)
{
    Vector4f c = {{0, 0, 0, 0}};
    auto it = p_->views.find(camID);
    if (it != p_->views.end() && segID < it->second.lines.size()) memcpy(c.v, it->second.lines[segID].v, sizeof(c.v));
    return c;
}
Vector4f Line3D::getSegmentCoords2D(const Segment2D& s) { return getSegmentCoords2D(s.camID(), s.segID()); }

std::string Line3D::createOutputFilename()   // line3D.cc:2855-2894
{
    Impl& P = *p_;
    std::stringstream s;
    s << "Line3D++__";
    if (P.max_img_width > 0) s << "W_" << P.max_img_width << "__"; else s << "W_FULL__";
    s << "N_" << P.num_neighbors << "__" << "sigmaP_" << P.sigma_p << "__" << "sigmaA_" << P.sigma_a << "__" << "epiOverlap_" << P.epi << "__";
    if (P.kNN > 0) s << "kNN_" << P.kNN << "__";
    if (P.collin_t > EPS) s << "COLLIN_" << P.collin_t << "__";
    if (P.use_ceres) s << "OPTIMIZED__";                                   // line3D.cc:2889-2890
    if (P.fixed3Dreg) { s << "FXD_SIGMA_P__"; if (P.const_reg_depth > 0.0f) s << "REG_DEPTH_" << P.const_reg_depth << "__"; }
    if (P.perform_RDD) s << "DIFFUSION__";
    s << "vis_" << P.visibility_t;
    return s.str();
}

void Line3D::save3DLinesAsTXT(const std::string& folder)   // line3D.cc:2631-2687; format README.md:272-277
{
    std::lock_guard<std::mutex> g(p_->mtx);
    p_->err.clear();
    if (p_->lines3D.empty()) { p_->fail("no 3D lines to save!"); return; }
    const std::string path = folder + "/" + createOutputFilename() + ".txt";
    std::ofstream f(path.c_str());
    if (!f) { p_->fail("cannot open " + path + " for writing"); return; }
    for (const FinalLine3D& L : p_->lines3D) {
        if (L.collinear3Dsegments_.empty()) continue;
        f << L.collinear3Dsegments_.size() << " ";
        for (const Segment3D& s : L.collinear3Dsegments_) f << s.P1().x << " " << s.P1().y << " " << s.P1().z << " " << s.P2().x << " " << s.P2().y << " " << s.P2().z << " ";
        f << L.underlyingCluster_.residuals()->size() << " ";
        for (const Segment2D& r : *L.underlyingCluster_.residuals()) {
            const Vector4f c = getSegmentCoords2D(r);
            f << r.camID() << " " << r.segID() << " " << c.v[0] << " " << c.v[1] << " " << c.v[2] << " " << c.v[3] << " ";
        }
        f << std::endl;
    }
}

// Wavefront OBJ as the reference writes it (line3D.cc:2568-2628): all vertices first ("v x y z", two per 3D segment,
// default ostream formatting), then one "l a b" element per segment with 1-based vertex indices.
void Line3D::saveResultAsOBJ(const std::string& folder)
{
    std::lock_guard<std::mutex> g(p_->mtx);
    p_->err.clear();
    if (p_->lines3D.empty()) { p_->fail("no 3D lines to save!"); return; }
    const std::string path = folder + "/" + createOutputFilename() + ".obj";
    std::ofstream f(path.c_str());
    if (!f) { p_->fail("cannot open " + path + " for writing"); return; }
    size_t nseg = 0;
    for (const FinalLine3D& L : p_->lines3D)
        for (const Segment3D& s : L.collinear3Dsegments_) {
            const Vector3d e[2] = {s.P1(), s.P2()};
            for (const Vector3d& v : e) f << "v " << v.x << " " << v.y << " " << v.z << std::endl;
            ++nseg;
        }
    for (size_t i = 0; i < nseg; ++i) f << "l " << 2 * i + 1 << " " << 2 * i + 2 << std::endl;
}

// ASCII STL as the reference writes it (line3D.cc:2465-2528): every 3D segment is a degenerate facet P1,P2,P1 with a
// fixed normal, coordinates printed with %e.
void Line3D::saveResultAsSTL(const std::string& folder)
{
    std::lock_guard<std::mutex> g(p_->mtx);
    p_->err.clear();
    if (p_->lines3D.empty()) { p_->fail("no 3D lines to save!"); return; }
    const std::string path = folder + "/" + createOutputFilename() + ".stl";
    std::ofstream f(path.c_str());
    if (!f) { p_->fail("cannot open " + path + " for writing"); return; }
    auto vertex = [&f](const Vector3d& v) { char b[200]; snprintf(b, sizeof(b), "   vertex %e %e %e", v.x, v.y, v.z); f << b << std::endl; };
    f << "solid lineModel" << std::endl;
    for (const FinalLine3D& L : p_->lines3D)
        for (const Segment3D& s : L.collinear3Dsegments_) {
            f << " facet normal 1.0e+000 0.0e+000 0.0e+000" << std::endl << "  outer loop" << std::endl;
            vertex(s.P1()); vertex(s.P2()); vertex(s.P1());
            f << "  endloop" << std::endl << " endfacet" << std::endl;
        }
    f << "endsolid lineModel" << std::endl;
}

Matrix3d Line3D::rotationFromQ(const double Qw, const double Qx, const double Qy, const double Qz)   // line3D.cc:2737-2754
{
    const double n = std::sqrt(Qw * Qw + Qx * Qx + Qy * Qy + Qz * Qz);
    const double w = Qw / n, x = Qx / n, y = Qy / n, z = Qz / n;
    Matrix3d R;
    R(0, 0) = 1 - 2 * y * y - 2 * z * z; R(0, 1) = 2 * x * y - 2 * z * w; R(0, 2) = 2 * x * z + 2 * y * w;
    R(1, 0) = 2 * x * y + 2 * z * w; R(1, 1) = 1 - 2 * x * x - 2 * z * z; R(1, 2) = 2 * y * z - 2 * x * w;
    R(2, 0) = 2 * x * z - 2 * y * w; R(2, 1) = 2 * y * z + 2 * x * w; R(2, 2) = 1 - 2 * x * x - 2 * y * y;
    return R;
}

}  // namespace L3DPP

// ================================================================================================ C wrapper (tests / Python)
extern "C" {
using namespace L3DPP;
void* l3dpp_create(int neighbors_by_worldpoints, int use_gpu, int device)
{
    try { return new Line3D("", true, -1, 3000, neighbors_by_worldpoints != 0, use_gpu != 0, device); } catch (...) { return nullptr; }
}
void l3dpp_destroy(void* h) { delete (Line3D*)h; }
const char* l3dpp_last_error(void* h) { return ((Line3D*)h)->lastError(); }
void* l3dpp_ctx(void* h) { return ((Line3D*)h)->impl()->ctx; }
int l3dpp_add_image(void* h, unsigned int cam, int w, int hgt, const double* K, const double* R, const double* t, float median_depth,
                    const unsigned int* list, int nlist, const float* segs, int nseg)
{
    Matrix3d Km, Rm; memcpy(Km.m, K, 72); memcpy(Rm.m, R, 72);
    std::vector<Vec4f> s(nseg); if (nseg) memcpy(s.data(), segs, 16 * (size_t)nseg);
    Line3D* L = (Line3D*)h;
    return L->addImage(cam, w, hgt, Km, Rm, Vector3d(t[0], t[1], t[2]), median_depth, std::list<unsigned int>(list, list + nlist), s) ? 0 : -1;
}
int l3dpp_match_images(void* h, float sp, float sa, unsigned int nn, float epi, int knn, float crd)
{ Line3D* L = (Line3D*)h; L->matchImages(sp, sa, nn, epi, knn, crd); return L->lastError()[0] ? -1 : 0; }
int l3dpp_reconstruct(void* h, unsigned int vis, int diffusion, float collin)
{ Line3D* L = (Line3D*)h; L->reconstruct3Dlines(vis, diffusion != 0, collin, false, 0); return L->lastError()[0] ? -1 : 0; }
int l3dpp_reconstruct_opt(void* h, unsigned int vis, int diffusion, float collin, int use_ceres, unsigned int max_iter)
{ Line3D* L = (Line3D*)h; L->reconstruct3Dlines(vis, diffusion != 0, collin, use_ceres != 0, max_iter); return L->lastError()[0] ? -1 : 0; }
int l3dpp_set_shard(void* h, int rank, int world, Line3D::MatchExchangeFn fn, void* user)
{ Line3D* L = (Line3D*)h; L->setShard(rank, world, fn, user); return L->lastError()[0] ? -1 : 0; }
int l3dpp_stats(void* h, Line3DStats* out) { *out = ((Line3D*)h)->stats(); return 0; }
int l3dpp_view_index(void* h, unsigned int cam) { auto& m = ((Line3D*)h)->impl()->index_of; auto it = m.find(cam); return it == m.end() ? -1 : it->second; }
// test hook for the cluster -> 3D segment tail: findCollinearSegments(cluster) (line3D.cc:2342-2452) on an explicit cluster
// (line end points + (camID, segID) residuals of views added with addImage).  Returns the number of 3D segments.
int l3dpp_collinear_from_cluster(void* h, const double* p1p2, int nres, const unsigned int* cams, const unsigned int* segs, double* out6, int cap)
{
    Line3D::Impl& P = *((Line3D*)h)->impl();
    if (P.vlist.size() != P.views.size()) {                       // views are indexed at the start of matchImages: do it here too
        P.vlist.clear(); P.index_of.clear();
        for (auto& kv : P.views) { P.index_of[kv.first] = (int)P.vlist.size(); P.vlist.push_back(&kv.second); }
    }
    std::list<Segment2D> res;
    for (int i = 0; i < nres; ++i) res.push_back(Segment2D(cams[i], segs[i]));
    const LineCluster3D cl(Segment3D(Vector3d(p1p2[0], p1p2[1], p1p2[2]), Vector3d(p1p2[3], p1p2[4], p1p2[5])), res, nres ? cams[0] : 0);
    int n = 0;
    for (const Segment3D& sg : P.collinear_segments(cl)) {
        if (n < cap) { double* o = out6 + 6 * n; o[0] = sg.P1().x; o[1] = sg.P1().y; o[2] = sg.P1().z; o[3] = sg.P2().x; o[4] = sg.P2().y; o[5] = sg.P2().z; }
        ++n;
    }
    return n;
}
int l3dpp_view_info(void* h, unsigned int cam, float* k, float* md)
{ auto& v = ((Line3D*)h)->impl()->views; auto it = v.find(cam); if (it == v.end()) return -1; *k = it->second.k; *md = it->second.median_depth; return 0; }
int l3dpp_get_pairs(void* h, int* out, int cap)   // (src cam, tgt cam) in match order
{
    Line3D::Impl* P = ((Line3D*)h)->impl();
    const int n = (int)(P->pairs.size() / 2);
    for (int i = 0; i < n && i < cap; ++i) { out[2 * i] = (int)P->vlist[P->pairs[2 * i]]->id; out[2 * i + 1] = (int)P->vlist[P->pairs[2 * i + 1]]->id; }
    return n;
}
long long l3dpp_get_affinity(void* h, int raw, int* ei, int* ej, float* ew, long long cap)
{
    Line3D::Impl* P = ((Line3D*)h)->impl();
    const std::vector<L3DPP::Edge>& A = raw ? P->A_raw : P->A;
    for (size_t e = 0; e < A.size() && (long long)e < cap; ++e) { ei[e] = A[e].i; ej[e] = A[e].j; ew[e] = A[e].w; }
    return (long long)A.size();
}
int l3dpp_get_local2global(void* h, unsigned int* cam_seg, int cap)
{
    Line3D::Impl* P = ((Line3D*)h)->impl();
    for (size_t i = 0; i < P->local2global.size() && (int)i < cap; ++i) { const Segment2D s = P->seg_of(P->local2global[i]); cam_seg[2 * i] = s.camID(); cam_seg[2 * i + 1] = s.segID(); }
    return (int)P->local2global.size();
}
struct l3dpp_seg3d { int line; int pad; double p1[3], p2[3]; };
struct l3dpp_residual { int line; unsigned int cam, seg; };
long long l3dpp_get_segments3d(void* h, l3dpp_seg3d* out, long long cap)
{
    std::vector<FinalLine3D> L; ((Line3D*)h)->get3Dlines(L);
    long long n = 0;
    for (size_t i = 0; i < L.size(); ++i)
        for (const Segment3D& s : L[i].collinear3Dsegments_) {
            if (out && n < cap) { out[n].line = (int)i; out[n].pad = 0; out[n].p1[0] = s.P1().x; out[n].p1[1] = s.P1().y; out[n].p1[2] = s.P1().z; out[n].p2[0] = s.P2().x; out[n].p2[1] = s.P2().y; out[n].p2[2] = s.P2().z; }
            ++n;
        }
    return n;
}
long long l3dpp_get_residuals(void* h, l3dpp_residual* out, long long cap)
{
    std::vector<FinalLine3D> L; ((Line3D*)h)->get3Dlines(L);
    long long n = 0;
    for (size_t i = 0; i < L.size(); ++i)
        for (const Segment2D& s : *L[i].underlyingCluster_.residuals()) { if (out && n < cap) { out[n].line = (int)i; out[n].cam = s.camID(); out[n].seg = s.segID(); } ++n; }
    return n;
}
int l3dpp_save_obj(void* h, const char* folder) { Line3D* L = (Line3D*)h; L->saveResultAsOBJ(folder); return L->lastError()[0] ? -1 : 0; }
int l3dpp_save_stl(void* h, const char* folder) { Line3D* L = (Line3D*)h; L->saveResultAsSTL(folder); return L->lastError()[0] ? -1 : 0; }
int l3dpp_output_filename(void* h, char* buf, int cap)
{ const std::string s = ((Line3D*)h)->createOutputFilename(); if ((int)s.size() + 1 > cap) return -1; memcpy(buf, s.c_str(), s.size() + 1); return (int)s.size(); }
// test hook for the writers: replace the result by explicit lines (segments: 6 doubles each; residuals: camID, segID of added views)
int l3dpp_set_lines(void* h, int nlines, const int* nseg, const double* segs6, const int* nres, const unsigned int* res_cam, const unsigned int* res_seg)
{
    Line3D::Impl& P = *((Line3D*)h)->impl();
    std::lock_guard<std::mutex> g(P.mtx);
    P.lines3D.clear();
    size_t so = 0, ro = 0;
    for (int i = 0; i < nlines; ++i) {
        FinalLine3D f;
        for (int k = 0; k < nseg[i]; ++k, ++so) {
            const double* q = segs6 + 6 * so;
            f.collinear3Dsegments_.push_back(Segment3D(Vector3d(q[0], q[1], q[2]), Vector3d(q[3], q[4], q[5])));
        }
        std::list<Segment2D> res;
        for (int k = 0; k < nres[i]; ++k, ++ro) res.push_back(Segment2D(res_cam[ro], res_seg[ro]));
        f.underlyingCluster_ = LineCluster3D(f.collinear3Dsegments_.empty() ? Segment3D() : f.collinear3Dsegments_.front(), res, nres[i] ? res.front().camID() : 0);
        P.lines3D.push_back(f);
    }
    return 0;
}
int l3dpp_save_txt(void* h, const char* folder) { Line3D* L = (Line3D*)h; L->save3DLinesAsTXT(folder); return L->lastError()[0] ? -1 : 0; }
}
