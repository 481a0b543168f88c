// oracle/ref_shim/opencv2/core.hpp — TEST INFRASTRUCTURE ONLY.
// Just enough of the OpenCV C++ surface for the reference's line3D.cc / view.cc to COMPILE verbatim
// (oracle/Makefile, target _ref/libl3dref_full_*.so).  Images carry only their size: the parity harness always
// passes explicit line segments (addImage(..., line_segments), line3D.cc:170-186), so every image-processing
// entry point (LSD detection, undistortion, drawing) aborts if it is ever reached.
#ifndef L3D_ORACLE_SHIM_OPENCV_CORE_H
#define L3D_ORACLE_SHIM_OPENCV_CORE_H
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <memory>

#define CV_8U 0
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_64FC1 6
#define CV_RGB2GRAY 7

namespace cv {

inline void shim_unreachable(const char* what) {
    std::fprintf(stderr, "oracle/ref_shim/opencv2: %s is not available in the parity build\n", what);
    std::abort();
}

struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; } };
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
struct Point { int x, y; Point(int xx = 0, int yy = 0) : x(xx), y(yy) {} };

template<class T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(0); }
    Vec(T a, T b, T c, T d) { static_assert(N == 4, "Vec4"); val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    T& operator()(int i) { return val[i]; } const T& operator()(int i) const { return val[i]; }
    T& operator[](int i) { return val[i]; } const T& operator[](int i) const { return val[i]; }
};
typedef Vec<float, 4> Vec4f;

class Mat {
    int type_;
    double dummy_;
public:
    int rows, cols;
    Mat() : type_(CV_8U), dummy_(0), rows(0), cols(0) {}
    Mat(int r, int c, int type, const Scalar& = Scalar()) : type_(type), dummy_(0), rows(r), cols(c) {}
    int type() const { return type_; }
    int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
    bool empty() const { return rows == 0 || cols == 0; }
    Mat clone() const { return *this; }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    template<class T> T& at(int, int = 0) { shim_unreachable("cv::Mat::at"); return *reinterpret_cast<T*>(&dummy_); }
};
template<class T> struct Mat_ : public Mat {
    static Mat eye(int r, int c) { return Mat(r, c, CV_64FC1); }
    static Mat zeros(int r, int c) { return Mat(r, c, CV_64FC1); }
};

template<class T> using Ptr = std::shared_ptr<T>;

} // namespace cv
#endif
