/* include/line3d.h — L3DPP::Line3D on B200: the host-side mirror of the reference's public class (line3D.h:61-233).
 *
 * Same class name, method names, argument meaning, defaults and output types (FinalLine3D / LineCluster3D /
 * Segment3D / Segment2D, segment3D.h:35-178, commons.h:103-133) so an SfM frontend written against the reference
 * (main_vsfm.cpp:306-330 etc.) ports by recompiling.  Everything data-parallel goes through the C ABI of
 * include/l3d_capi.h; this class only does what line3D.cc does on the host around the accelerator calls.
 *
 * Differences forced by this image (no Eigen / OpenCV / Boost headers):
 *   - Matrix3d / Vector3d / Vec4f are small PODs.  When <eigen3/Eigen/Eigen> and <opencv2/core.hpp> are on the include path
 *     (or L3DPP_WITH_EIGEN_OPENCV is defined) the reference's own addImage signature (line3D.h:104-108: cv::Mat&,
 *     Eigen::Matrix3d, Eigen::Vector3d, std::vector<cv::Vec4f>) is available as an inline overload that converts and
 *     forwards; tests/test_cpu.py compiles it against the header stand-ins of oracle/ref_shim.  get3Dlines() returns the POD
 *     types (P1().x is a member, not Eigen's x()): a frontend that reads the result lines needs that one-character change.
 *   - addImage takes the image SIZE and explicit 2D segments (the reference's own line_segments argument,
 *     line3D.h:102-108); LSD detection inside addImage (line3D.cc:249-372) is outside the hot path.
 *   - save3DLinesAsBIN (boost archive) is not provided; TXT / OBJ / STL are.
 *   - use_GPU selects the SCORING SEMANTICS of the reference's two paths; both run on the GPU (there is no CPU path).
 */
#ifndef L3DPP_B200_LINE3D_H_
#define L3DPP_B200_LINE3D_H_

#include <cstddef>
#include <list>
#include <string>
#include <vector>

#if !defined(L3DPP_WITH_EIGEN_OPENCV) && defined(__has_include)
#if __has_include(<eigen3/Eigen/Eigen>) && __has_include(<opencv2/core.hpp>)
#define L3DPP_WITH_EIGEN_OPENCV 1
#endif
#endif
#ifdef L3DPP_WITH_EIGEN_OPENCV
#include <eigen3/Eigen/Eigen>
#include <opencv2/core.hpp>
#endif

namespace L3DPP {

struct Vector3d { double x, y, z; Vector3d() : x(0), y(0), z(0) {} Vector3d(double a, double b, double c) : x(a), y(b), z(c) {} };
struct Matrix3d { double m[9]; double& operator()(int r, int c) { return m[r * 3 + c]; } double operator()(int r, int c) const { return m[r * 3 + c]; } };
struct Vec4f { float v[4]; float operator()(int i) const { return v[i]; } };
struct Vector4f { float v[4]; float operator()(int i) const { return v[i]; } };

/* defaults: commons.h:41-69 */
const int L3D_DEF_MAX_IMG_WIDTH = -1;
const unsigned int L3D_DEF_MAX_NUM_SEGMENTS = 3000;
const bool L3D_DEF_LOAD_AND_STORE_SEGMENTS = true;
const float L3D_DEF_COLLINEARITY_T = -1.0f;
const unsigned int L3D_DEF_MATCHING_NEIGHBORS = 10;
const float L3D_DEF_EPIPOLAR_OVERLAP = 0.25f;
const int L3D_DEF_KNN = 10;
const float L3D_DEF_SCORING_POS_REGULARIZER = 2.5f;
const float L3D_DEF_SCORING_ANG_REGULARIZER = 10.0f;
const bool L3D_DEF_PERFORM_RDD = false;
const unsigned int L3D_DEF_MIN_VISIBILITY_T = 3;
const bool L3D_DEF_USE_CERES = false;
const unsigned int L3D_DEF_CERES_MAX_ITER = 250;

class Segment2D {   /* commons.h:103-133 */
public:
    Segment2D() : camID_(0), segID_(0) {}
    Segment2D(unsigned int camID, unsigned int segID) : camID_(camID), segID_(segID) {}
    unsigned int camID() const { return camID_; }
    unsigned int segID() const { return segID_; }
    bool operator==(const Segment2D& r) const { return camID_ == r.camID_ && segID_ == r.segID_; }
    bool operator<(const Segment2D& r) const { return camID_ < r.camID_ || (camID_ == r.camID_ && segID_ < r.segID_); }
    bool operator!=(const Segment2D& r) const { return !(*this == r); }
private:
    unsigned int camID_, segID_;
};

class Segment3D {   /* segment3D.h:35-94 */
public:
    Segment3D();
    Segment3D(const Vector3d& P1, const Vector3d& P2);
    float distance_Point2Line(const Vector3d& P) const;
    void translate(const Vector3d& t);
    Vector3d P1() const { return P1_; }
    Vector3d P2() const { return P2_; }
    Vector3d dir() const { return dir_; }
    float length() const { return length_; }
    bool valid() const { return valid_; }
private:
    Vector3d P1_, P2_, dir_;
    float length_;
    bool valid_;
};

class LineCluster3D {   /* segment3D.h:124-151 */
public:
    LineCluster3D() : reference_view_(0) {}
    LineCluster3D(const Segment3D& seg3D, const std::list<Segment2D>& residuals, unsigned int ref_view)
        : seg3D_(seg3D), residuals_(residuals), reference_view_(ref_view) {}
    Segment3D seg3D() const { return seg3D_; }
    const std::list<Segment2D>* residuals() const { return &residuals_; }
    size_t size() const { return residuals_.size(); }
    unsigned int reference_view() const { return reference_view_; }
    void update3Dline(const Segment3D& s) { seg3D_ = s; }
    void translate(const Vector3d& t) { seg3D_.translate(t); }
private:
    Segment3D seg3D_;
    std::list<Segment2D> residuals_;
    unsigned int reference_view_;
};

struct FinalLine3D {   /* segment3D.h:155-163 */
    std::list<Segment3D> collinear3Dsegments_;
    LineCluster3D underlyingCluster_;
};

/* counters the reference prints to stdout (SURVEY.md §5 "Metrics / logging"); handy parity checkpoints */
struct Line3DStats {
    long long view_pairs, pair_evaluations, matches_after_knn, estimates, affinity_entries, affinity_rows, clusters_total,
        clusters_valid, lines3D, collinear_entries, opt_iterations;
    double ms_match, ms_score, ms_affinity, ms_diffusion, ms_cluster, opt_cost_before, opt_cost_after;
};

class Line3D {
public:
    /* line3D.h:80-85.  output_folder / load_segments / max_img_width / max_line_segments only matter for the
     * detection cache and the output filename (createOutputFilename). */
    Line3D(const std::string& output_folder, const bool load_segments = L3D_DEF_LOAD_AND_STORE_SEGMENTS,
           const int max_img_width = L3D_DEF_MAX_IMG_WIDTH, const unsigned int max_line_segments = L3D_DEF_MAX_NUM_SEGMENTS,
           const bool neighbors_by_worldpoints = true, const bool use_GPU = true, const int cuda_device = 0);
    ~Line3D();
    Line3D(const Line3D&) = delete;
    Line3D& operator=(const Line3D&) = delete;

    /* line3D.h:104-108 with (image_width, image_height) instead of cv::Mat& image; line_segments must be non-empty */
    /* Returns false (and sets lastError(), which a later successful addImage does not clear) if the view was rejected. */
    bool addImage(const unsigned int camID, const int image_width, const int image_height, const Matrix3d& K, const Matrix3d& R,
                  const Vector3d& t, const float median_depth, const std::list<unsigned int>& wps_or_neighbors,
                  const std::vector<Vec4f>& line_segments);

#ifdef L3DPP_WITH_EIGEN_OPENCV
    /* The reference's exact signature (line3D.h:104-108).  The image only contributes its size; line_segments must be given
     * (line detection is outside this library): with an empty list the call fails like any other rejected view. */
    bool addImage(const unsigned int camID, cv::Mat& image, const Eigen::Matrix3d& K, const Eigen::Matrix3d& R, const Eigen::Vector3d& t,
                  const float median_depth, const std::list<unsigned int>& wps_or_neighbors,
                  const std::vector<cv::Vec4f>& line_segments = std::vector<cv::Vec4f>())
    {
        Matrix3d Kp, Rp;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { Kp(r, c) = K(r, c); Rp(r, c) = R(r, c); }
        std::vector<Vec4f> segs(line_segments.size());
        for (size_t i = 0; i < line_segments.size(); ++i) for (int k = 0; k < 4; ++k) segs[i].v[k] = line_segments[i](k);
        return addImage(camID, image.cols, image.rows, Kp, Rp, Vector3d(t(0), t(1), t(2)), median_depth, wps_or_neighbors, segs);
    }
#endif

    /* line3D.h:140-145 */
    void matchImages(const float sigma_position = L3D_DEF_SCORING_POS_REGULARIZER, const float sigma_angle = L3D_DEF_SCORING_ANG_REGULARIZER,
                     const unsigned int num_neighbors = L3D_DEF_MATCHING_NEIGHBORS, const float epipolar_overlap = L3D_DEF_EPIPOLAR_OVERLAP,
                     const int kNN = L3D_DEF_KNN, const float const_regularization_depth = -1.0f);

    /* line3D.h:162-166.  use_CERES: the bundling of optimization.cc runs on the GPU (l3d_optimize_lines); Ceres is not needed,
     * so the flag is honoured in every build (the reference silently drops it when it was built without Ceres, line3D.cc:1738-1744) */
    void reconstruct3Dlines(const unsigned int visibility_t = L3D_DEF_MIN_VISIBILITY_T, const bool perform_diffusion = L3D_DEF_PERFORM_RDD,
                            const float collinearity_t = L3D_DEF_COLLINEARITY_T, const bool use_CERES = L3D_DEF_USE_CERES,
                            const unsigned int max_iter_CERES = L3D_DEF_CERES_MAX_ITER);

    void get3Dlines(std::vector<FinalLine3D>& result);                 /* line3D.h:173 */
    void saveResultAsSTL(const std::string& output_folder);            /* line3D.h:183-185 */
    void saveResultAsOBJ(const std::string& output_folder);
    void save3DLinesAsTXT(const std::string& output_folder);
    Vector4f getSegmentCoords2D(const Segment2D& seg2D);               /* line3D.h:196-198 */
    Vector4f getSegmentCoords2D(const unsigned int camID, const unsigned int segID);
    size_t numImages();                                                /* line3D.h:205 */
    std::string createOutputFilename();                                /* line3D.h:224 */
    static Matrix3d rotationFromQ(const double Qw, const double Qx, const double Qy, const double Qz);   /* line3D.h:219-220 */

    /* --- additions (not in the reference) --- */
    const Line3DStats& stats() const;
    const char* lastError() const;     /* empty string if the last call succeeded */
    void setVerbose(bool v);
    /* Multi-GPU matching. One Line3D per GPU / process, every one fed the same images. With world > 1
     * matchImages evaluates only this rank's contiguous, cost-balanced share of the view pairs and then calls
     * `exchange(user, counts_dev, recs_dev, row_bounds, world, knn)`: rank r owns rows [row_bounds[r], row_bounds[r+1]) of
     * counts (int32 per row) and recs (knn * 24 B per row), both DEVICE pointers; on return every rank must hold all rows
     * (an NCCL broadcast per owner; line3dpp_b200/dist.py does it with torch.distributed). Scoring, affinity, diffusion and
     * clustering then run replicated, so every rank ends with the single-GPU result bit for bit. Return 0 on success. */
    typedef int (*MatchExchangeFn)(void* user, void* counts_dev, void* recs_dev, const long long* row_bounds, int world, int knn);
    void setShard(int rank, int world, MatchExchangeFn exchange, void* user);
    struct Impl;
    Impl* impl() { return p_; }        /* for the C wrapper used by the tests */
private:
    Impl* p_;
};

}  // namespace L3DPP
#endif
