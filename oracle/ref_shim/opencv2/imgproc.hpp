// oracle/ref_shim/opencv2/imgproc.hpp — TEST INFRASTRUCTURE ONLY (see core.hpp): image-processing entry points abort.
#ifndef L3D_ORACLE_SHIM_OPENCV_IMGPROC_H
#define L3D_ORACLE_SHIM_OPENCV_IMGPROC_H
#include "opencv2/core.hpp"
namespace cv {
enum { INTER_LINEAR = 1, BORDER_CONSTANT = 0, LSD_REFINE_ADV = 2 };
inline void cvtColor(const Mat&, Mat&, int) { shim_unreachable("cv::cvtColor"); }
inline void resize(const Mat&, Mat&, Size, double = 0, double = 0) { shim_unreachable("cv::resize"); }
inline void initUndistortRectifyMap(const Mat&, const Mat&, const Mat&, const Mat&, Size, int, Mat&, Mat&) { shim_unreachable("cv::initUndistortRectifyMap"); }
inline void remap(const Mat&, Mat&, const Mat&, const Mat&, int, int) { shim_unreachable("cv::remap"); }
inline void line(Mat&, Point, Point, const Scalar&, int = 1) { shim_unreachable("cv::line"); }
struct LineSegmentDetector { void detect(const Mat&, std::vector<Vec4f>&) { shim_unreachable("cv::LineSegmentDetector::detect"); } };
inline Ptr<LineSegmentDetector> createLineSegmentDetector(int = 0) { return Ptr<LineSegmentDetector>(new LineSegmentDetector()); }
} // namespace cv
#endif
