/* include/line3d_io.h — input format of the VisualSfM frontend: what main_vsfm.cpp:143-310 reads from an
 * .nvm file before it calls Line3D::addImage, as a library function so that a frontend built on include/line3d.h needs no
 * parsing code of its own.  Host-only; no CUDA involved. */
#ifndef L3DPP_LINE3D_IO_H_
#define L3DPP_LINE3D_IO_H_
#include "line3d.h"

#include <list>
#include <string>
#include <vector>

namespace L3DPP {

/* one camera of an NVM model, ready for Line3D::addImage (main_vsfm.cpp:165-208, 225-236, 301-309) */
struct NVMCamera {
    std::string image;                     /* file name as written in the .nvm */
    float focal;                           /* cams_focals is a float vector in the reference (main_vsfm.cpp:166, 190) */
    float distortion;                      /* radial distortion coefficient; != 0: the frontend undistorts with (-d, 0, 0) */
    Matrix3d R;                            /* from the quaternion (qw qx qy qz) */
    Vector3d t, C;                         /* t = -R C */
    std::list<unsigned int> worldpoints;   /* ids of the 3D points this camera measures (neighbors_by_worldpoints = true) */
    float median_depth;                    /* median distance of those points to C; 0 if there are none (camera is skipped) */
};

/* Reads an NVM_V3 file like main_vsfm.cpp:143-250.  Returns false and sets *error on an unreadable or empty model. */
bool readNVM(const std::string& path, std::vector<NVMCamera>& cameras, std::string* error = 0);

/* one camera of a Bundler or COLMAP model, ready for Line3D::addImage.  Bundler has no K in the file (focal only: build it with
 * intrinsicsFromFocal once the image size is known, main_bundler.cpp:340-351); COLMAP stores K per camera model. */
struct SfMCamera {
    unsigned int id;                       /* camID handed to addImage: bundler camera index / COLMAP IMAGE_ID */
    std::string image;                     /* image file name (bundler: from the optional list file, may be empty) */
    bool has_K; Matrix3d K;                /* COLMAP only */
    float focal;                           /* bundler only */
    double radial[3], tangential[2];       /* distortion the frontend undistorts with (bundler: k1 k2 0; COLMAP: per camera model) */
    Matrix3d R; Vector3d t, C;             /* x_cam = R X + t,  C = -R^T t */
    std::list<unsigned int> worldpoints;
    float median_depth;                    /* 0 if the camera sees no world point (it is skipped by the frontends) */
};

/* bundle.rd.out (+ optional image list, one file name per line) like main_bundler.cpp:143-262: rows 2 and 3 of R and t.y, t.z are
 * flipped (Bundler looks down -z). */
bool readBundler(const std::string& bundle_file, const std::string& image_list_file, std::vector<SfMCamera>& cameras, std::string* error = 0);

/* cameras.txt / images.txt / points3D.txt of a COLMAP text model like main_colmap.cpp:140-349 (models SIMPLE_PINHOLE, PINHOLE,
 * SIMPLE_RADIAL, RADIAL, OPENCV, FULL_OPENCV; images whose camera is unknown are dropped).  Cameras are returned in file order. */
bool readColmap(const std::string& sfm_folder, std::vector<SfMCamera>& cameras, std::string* error = 0);

/* K of main_vsfm.cpp:266-278: focal on the diagonal, principal point at the image centre (float arithmetic like the reference) */
Matrix3d intrinsicsFromFocal(float focal, int image_width, int image_height);

}  // namespace L3DPP
#endif
