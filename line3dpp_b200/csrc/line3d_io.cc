// line3d_io.cc — .nvm reader for frontends built on include/line3d.h (restates main_vsfm.cpp:143-310; host only)
#include "../../include/line3d_io.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>

namespace L3DPP {

bool readNVM(const std::string& path, std::vector<NVMCamera>& cams, std::string* error)
{
    cams.clear();
    std::ifstream f(path.c_str());
    if (!f) { if (error) *error = "cannot open " + path; return false; }
    std::string line;
    std::getline(f, line);                       // "NVM_V3"
    std::getline(f, line);                       // empty
    std::getline(f, line);
    unsigned int num_cams = 0;
    { std::stringstream s(line); s >> num_cams; }
    if (num_cams == 0) { if (error) *error = "No aligned cameras in NVM file!"; return false; }     // main_vsfm.cpp:157-161
    cams.resize(num_cams);
    for (unsigned int i = 0; i < num_cams; ++i) {
        std::getline(f, line);
        std::stringstream s(line);
        double focal = 0, qw = 1, qx = 0, qy = 0, qz = 0, Cx = 0, Cy = 0, Cz = 0, dist = 0;
        NVMCamera& c = cams[i];
        s >> c.image >> focal >> qw >> qx >> qy >> qz >> Cx >> Cy >> Cz >> dist;
        c.focal = (float)focal; c.distortion = (float)dist;
        Matrix3d& R = c.R;                        // main_vsfm.cpp:193-203 (the quaternion is used as stored, not re-normalised)
        R(0, 0) = 1.0 - 2.0 * qy * qy - 2.0 * qz * qz; R(0, 1) = 2.0 * qx * qy - 2.0 * qz * qw; R(0, 2) = 2.0 * qx * qz + 2.0 * qy * qw;
        R(1, 0) = 2.0 * qx * qy + 2.0 * qz * qw; R(1, 1) = 1.0 - 2.0 * qx * qx - 2.0 * qz * qz; R(1, 2) = 2.0 * qy * qz - 2.0 * qx * qw;
        R(2, 0) = 2.0 * qx * qz - 2.0 * qy * qw; R(2, 1) = 2.0 * qy * qz + 2.0 * qx * qw; R(2, 2) = 1.0 - 2.0 * qx * qx - 2.0 * qy * qy;
        c.C = Vector3d(Cx, Cy, Cz);
        c.t = Vector3d(-(R(0, 0) * Cx + R(0, 1) * Cy + R(0, 2) * Cz), -(R(1, 0) * Cx + R(1, 1) * Cy + R(1, 2) * Cz),
                       -(R(2, 0) * Cx + R(2, 1) * Cy + R(2, 2) * Cz));
        c.median_depth = 0.0f;
    }
    std::getline(f, line);                       // empty
    std::getline(f, line);
    unsigned int num_points = 0;
    { std::stringstream s(line); s >> num_points; }
    std::vector<std::vector<float> > depths(num_cams);
    for (unsigned int i = 0; i < num_points; ++i) {
        if (!std::getline(f, line)) break;
        std::istringstream s(line);
        double px, py, pz, cr, cg, cb;
        s >> px >> py >> pz >> cr >> cg >> cb;
        unsigned int nviews = 0;
        s >> nviews;
        for (unsigned int j = 0; j < nviews; ++j) {
            unsigned int cam = 0, sift = 0; float x, y;
            s >> cam >> sift >> x >> y;
            if (!s || cam >= num_cams) { if (error) *error = "malformed measurement list in " + path; return false; }
            cams[cam].worldpoints.push_back(i);
            const double dx = px - cams[cam].C.x, dy = py - cams[cam].C.y, dz = pz - cams[cam].C.z;
            depths[cam].push_back((float)std::sqrt(dx * dx + dy * dy + dz * dz));       // float vector, main_vsfm.cpp:228, 248
        }
    }
    for (unsigned int i = 0; i < num_cams; ++i)
        if (!depths[i].empty()) { std::sort(depths[i].begin(), depths[i].end()); cams[i].median_depth = depths[i][depths[i].size() / 2]; }   // 301-303
    return true;
}

Matrix3d intrinsicsFromFocal(float focal, int w, int h)
{
    Matrix3d K; std::memset(K.m, 0, sizeof(K.m));
    K(0, 0) = focal; K(1, 1) = focal; K(0, 2) = float(w) / 2.0f; K(1, 2) = float(h) / 2.0f; K(2, 2) = 1.0;
    return K;
}

}  // namespace L3DPP

// C wrapper (tests / Python): flat arrays.  R, t: 9 + 3 doubles per camera; returns the number of cameras or -1.
extern "C" {
void* l3dpp_nvm_open(const char* path, char* err, int errcap)
{
    std::vector<L3DPP::NVMCamera>* v = new std::vector<L3DPP::NVMCamera>();
    std::string e;
    if (!L3DPP::readNVM(path, *v, &e)) { if (err && errcap > 0) { std::strncpy(err, e.c_str(), errcap - 1); err[errcap - 1] = 0; } delete v; return nullptr; }
    return v;
}
void l3dpp_nvm_close(void* h) { delete (std::vector<L3DPP::NVMCamera>*)h; }
int l3dpp_nvm_num_cameras(void* h) { return (int)((std::vector<L3DPP::NVMCamera>*)h)->size(); }
int l3dpp_nvm_camera(void* h, int i, double* R9, double* t3, double* C3, float* focal, float* distortion, float* median_depth, int* num_wps, char* name, int namecap)
{
    const std::vector<L3DPP::NVMCamera>& v = *(std::vector<L3DPP::NVMCamera>*)h;
    if (i < 0 || i >= (int)v.size()) return -1;
    const L3DPP::NVMCamera& c = v[i];
    for (int k = 0; k < 9; ++k) R9[k] = c.R.m[k];
    t3[0] = c.t.x; t3[1] = c.t.y; t3[2] = c.t.z; C3[0] = c.C.x; C3[1] = c.C.y; C3[2] = c.C.z;
    *focal = c.focal; *distortion = c.distortion; *median_depth = c.median_depth; *num_wps = (int)c.worldpoints.size();
    if (name && namecap > 0) { std::strncpy(name, c.image.c_str(), namecap - 1); name[namecap - 1] = 0; }
    return 0;
}
int l3dpp_nvm_worldpoints(void* h, int i, unsigned int* out, int cap)
{
    const std::vector<L3DPP::NVMCamera>& v = *(std::vector<L3DPP::NVMCamera>*)h;
    if (i < 0 || i >= (int)v.size()) return -1;
    int n = 0;
    for (unsigned int w : v[i].worldpoints) { if (n < cap) out[n] = w; ++n; }
    return n;
}
void l3dpp_intrinsics_from_focal(float focal, int w, int h, double* K9) { const L3DPP::Matrix3d K = L3DPP::intrinsicsFromFocal(focal, w, h); for (int k = 0; k < 9; ++k) K9[k] = K.m[k]; }
}
