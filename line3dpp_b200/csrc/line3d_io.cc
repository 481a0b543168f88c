// line3d_io.cc — .nvm reader for frontends built on include/line3d.h (restates main_vsfm.cpp:143-310; host only)
#include "../../include/line3d_io.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <fstream>
#include <map>
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
        double focal = 0, qw = 1, qx = 0, qy = 0, qz = 0, Cx = 0, Cy = 0, Cz = 0, dist = 0;
        NVMCamera& c = cams[i];
        // a truncated or malformed camera line must not become a default camera (the reference has the same hole, main_vsfm.cpp:168-190)
        if (!std::getline(f, line)) { if (error) *error = "unexpected end of file in the camera list of " + path + " (camera " + std::to_string(i) + ")"; return false; }
        std::stringstream s(line);
        if (!(s >> c.image >> focal >> qw >> qx >> qy >> qz >> Cx >> Cy >> Cz >> dist)) {
            if (error) *error = "malformed camera line " + std::to_string(i) + " in " + path;
            return false;
        }
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

static Vector3d center_of(const Matrix3d& R, const Vector3d& t)      // C = R^T (-t)
{
    return Vector3d(-(R(0, 0) * t.x + R(1, 0) * t.y + R(2, 0) * t.z), -(R(0, 1) * t.x + R(1, 1) * t.y + R(2, 1) * t.z),
                    -(R(0, 2) * t.x + R(1, 2) * t.y + R(2, 2) * t.z));
}
static float median_of(std::vector<float>& d) { if (d.empty()) return 0.0f; std::sort(d.begin(), d.end()); return d[d.size() / 2]; }

bool readBundler(const std::string& bundle_file, const std::string& image_list_file, std::vector<SfMCamera>& cams, std::string* error)
{
    cams.clear();
    std::ifstream f(bundle_file.c_str());
    if (!f) { if (error) *error = "cannot open " + bundle_file; return false; }
    std::string line;
    std::getline(f, line);                       // "# Bundle file v0.3"
    std::getline(f, line);
    unsigned int num_cams = 0, num_points = 0;
    { std::stringstream s(line); s >> num_cams >> num_points; }
    if (num_cams == 0 || num_points == 0) { if (error) *error = "No cameras and/or points in bundle file!"; return false; }   // main_bundler.cpp:159-163
    cams.resize(num_cams);
    for (unsigned int i = 0; i < num_cams; ++i) {
        SfMCamera& c = cams[i];
        c.id = i; c.has_K = false; std::memset(c.K.m, 0, sizeof(c.K.m)); c.median_depth = 0.0f;
        double focal = 0, d1 = 0, d2 = 0;
        bool good = (bool)std::getline(f, line);
        if (good) { std::stringstream s(line); good = (bool)(s >> focal >> d1 >> d2); }
        c.focal = (float)focal;
        c.radial[0] = (float)d1; c.radial[1] = (float)d2; c.radial[2] = 0.0; c.tangential[0] = c.tangential[1] = 0.0;   // float pair, main_bundler.cpp:170, 183
        for (int j = 0; j < 3 && good; ++j) { good = (bool)std::getline(f, line); std::stringstream s(line); good = good && (bool)(s >> c.R(j, 0) >> c.R(j, 1) >> c.R(j, 2)); }
        for (int k = 0; k < 3; ++k) { c.R(1, k) *= -1.0; c.R(2, k) *= -1.0; }                                           // main_bundler.cpp:196-198
        good = good && (bool)std::getline(f, line);
        if (good) { std::stringstream s(line); good = (bool)(s >> c.t.x >> c.t.y >> c.t.z); }
        if (!good) { if (error) *error = "truncated or malformed camera " + std::to_string(i) + " in " + bundle_file; return false; }
        c.t.y *= -1.0; c.t.z *= -1.0;                                                                                    // main_bundler.cpp:210-212
        c.C = center_of(c.R, c.t);
    }
    std::vector<std::vector<float> > depths(num_cams);
    for (unsigned int i = 0; i < num_points; ++i) {
        if (!std::getline(f, line)) break;
        double px = 0, py = 0, pz = 0;
        { std::istringstream s(line); s >> px >> py >> pz; }
        std::getline(f, line);                   // colour
        std::getline(f, line);                   // view list
        std::istringstream s(line);
        unsigned int nviews = 0;
        s >> nviews;
        for (unsigned int j = 0; j < nviews; ++j) {
            unsigned int cam = 0, sift = 0; float x, y;
            s >> cam >> sift >> x >> y;
            if (!s || cam >= num_cams) { if (error) *error = "malformed view list in " + bundle_file; return false; }
            cams[cam].worldpoints.push_back(i);
            const double dx = px - cams[cam].C.x, dy = py - cams[cam].C.y, dz = pz - cams[cam].C.z;
            depths[cam].push_back((float)std::sqrt(dx * dx + dy * dy + dz * dz));
        }
    }
    for (unsigned int i = 0; i < num_cams; ++i) cams[i].median_depth = median_of(depths[i]);
    if (!image_list_file.empty()) {              // main_bundler.cpp:264-286: the first token of line i names image i
        std::ifstream l(image_list_file.c_str());
        unsigned int id = 0;
        while (std::getline(l, line)) {
            std::stringstream s(line);
            std::string fname;
            s >> fname;
            if (!fname.empty() && id < num_cams) cams[id].image = fname;
            ++id;
        }
    }
    return true;
}

bool readColmap(const std::string& folder, std::vector<SfMCamera>& cams, std::string* error)
{
    cams.clear();
    std::ifstream fc((folder + "/cameras.txt").c_str()), fi((folder + "/images.txt").c_str()), fp((folder + "/points3D.txt").c_str());
    if (!fc || !fi || !fp) { if (error) *error = "at least one of the colmap result files does not exist in sfm folder: " + folder; return false; }
    struct Intr { Matrix3d K; double radial[3], tangential[2]; };
    std::map<unsigned int, Intr> intr;
    std::string line;
    while (std::getline(fc, line)) {             // main_colmap.cpp:160-234
        if (line.empty() || line[0] == '#') continue;
        std::stringstream s(line);
        unsigned int cam = 0, w = 0, h = 0; std::string model;
        s >> cam >> model >> w >> h;
        double fx = 0, fy = 0, cx = 0, cy = 0, k1 = 0, k2 = 0, k3 = 0, p1 = 0, p2 = 0;
        if (model == "SIMPLE_PINHOLE") { s >> fx >> cx >> cy; fy = fx; }
        else if (model == "PINHOLE") s >> fx >> fy >> cx >> cy;
        else if (model == "SIMPLE_RADIAL") { s >> fx >> cx >> cy >> k1; fy = fx; }
        else if (model == "RADIAL") { s >> fx >> cx >> cy >> k1 >> k2; fy = fx; }
        else if (model == "OPENCV") s >> fx >> fy >> cx >> cy >> k1 >> k2 >> p1 >> p2;
        else if (model == "FULL_OPENCV") s >> fx >> fy >> cx >> cy >> k1 >> k2 >> p1 >> p2 >> k3;
        else { if (error) *error = "camera model " + model + " unknown!"; return false; }
        Intr in; std::memset(in.K.m, 0, sizeof(in.K.m));
        in.K(0, 0) = fx; in.K(0, 2) = cx; in.K(1, 1) = fy; in.K(1, 2) = cy; in.K(2, 2) = 1.0;
        in.radial[0] = k1; in.radial[1] = k2; in.radial[2] = k3; in.tangential[0] = p1; in.tangential[1] = p2;
        intr[cam] = in;
    }
    std::map<unsigned int, Vector3d> wps;        // world points referenced by a kept image; (0,0,0) until points3D.txt fills them in
    bool first = true, keep = false;
    while (std::getline(fi, line)) {             // main_colmap.cpp:252-321: two lines per image
        if (!line.empty() && line[0] == '#') continue;
        std::stringstream s(line);
        if (first) {
            unsigned int img = 0, cam = 0; double qw = 1, qx = 0, qy = 0, qz = 0, tx = 0, ty = 0, tz = 0; std::string name;
            s >> img >> qw >> qx >> qy >> qz >> tx >> ty >> tz >> cam >> name;
            keep = intr.count(cam) != 0;
            if (keep) {
                SfMCamera c;
                c.id = img; c.image = name; c.has_K = true; c.K = intr[cam].K; c.focal = 0.0f; c.median_depth = 0.0f;
                for (int k = 0; k < 3; ++k) c.radial[k] = intr[cam].radial[k];
                c.tangential[0] = intr[cam].tangential[0]; c.tangential[1] = intr[cam].tangential[1];
                c.R = Line3D::rotationFromQ(qw, qx, qy, qz); c.t = Vector3d(tx, ty, tz); c.C = center_of(c.R, c.t);
                cams.push_back(c);
            }
            first = false;
        } else {
            if (keep) {
                double x, y; std::string id;
                while (s >> x >> y >> id) {
                    const int wp = std::atoi(id.c_str());
                    if (wp >= 0) { cams.back().worldpoints.push_back((unsigned int)wp); wps[(unsigned int)wp] = Vector3d(0, 0, 0); }
                }
            }
            first = true;
        }
    }
    while (std::getline(fp, line)) {             // main_colmap.cpp:330-349
        if (line.empty() || line[0] == '#') continue;
        std::stringstream s(line);
        unsigned int id = 0; double X = 0, Y = 0, Z = 0;
        s >> id >> X >> Y >> Z;
        std::map<unsigned int, Vector3d>::iterator it = wps.find(id);
        if (s && it != wps.end()) it->second = Vector3d(X, Y, Z);
    }
    for (size_t i = 0; i < cams.size(); ++i) {   // main_colmap.cpp:386-401
        std::vector<float> d;
        for (unsigned int w : cams[i].worldpoints) {
            const Vector3d& P = wps[w];
            const double dx = cams[i].C.x - P.x, dy = cams[i].C.y - P.y, dz = cams[i].C.z - P.z;
            d.push_back((float)std::sqrt(dx * dx + dy * dy + dz * dz));
        }
        cams[i].median_depth = median_of(d);
    }
    if (cams.empty()) { if (error) *error = "no usable image in " + folder; return false; }
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
void* l3dpp_sfm_open(int kind /*0 bundler, 1 colmap*/, const char* path, const char* aux, char* err, int errcap)
{
    std::vector<L3DPP::SfMCamera>* v = new std::vector<L3DPP::SfMCamera>();
    std::string e;
    const bool ok = kind == 0 ? L3DPP::readBundler(path, aux ? aux : "", *v, &e) : L3DPP::readColmap(path, *v, &e);
    if (!ok) { if (err && errcap > 0) { std::strncpy(err, e.c_str(), errcap - 1); err[errcap - 1] = 0; } delete v; return nullptr; }
    return v;
}
void l3dpp_sfm_close(void* h) { delete (std::vector<L3DPP::SfMCamera>*)h; }
int l3dpp_sfm_num_cameras(void* h) { return (int)((std::vector<L3DPP::SfMCamera>*)h)->size(); }
/* out: id, has_K, focal, median_depth, num_wps as 5 doubles; K R 9 each; t C 3 each; dist = radial[3] + tangential[2] */
int l3dpp_sfm_camera(void* h, int i, double* head5, double* K9, double* R9, double* t3, double* C3, double* dist5, char* name, int namecap)
{
    const std::vector<L3DPP::SfMCamera>& v = *(std::vector<L3DPP::SfMCamera>*)h;
    if (i < 0 || i >= (int)v.size()) return -1;
    const L3DPP::SfMCamera& c = v[i];
    head5[0] = c.id; head5[1] = c.has_K; head5[2] = c.focal; head5[3] = c.median_depth; head5[4] = (double)c.worldpoints.size();
    for (int k = 0; k < 9; ++k) { K9[k] = c.K.m[k]; R9[k] = c.R.m[k]; }
    t3[0] = c.t.x; t3[1] = c.t.y; t3[2] = c.t.z; C3[0] = c.C.x; C3[1] = c.C.y; C3[2] = c.C.z;
    for (int k = 0; k < 3; ++k) dist5[k] = c.radial[k];
    dist5[3] = c.tangential[0]; dist5[4] = c.tangential[1];
    if (name && namecap > 0) { std::strncpy(name, c.image.c_str(), namecap - 1); name[namecap - 1] = 0; }
    return 0;
}
int l3dpp_sfm_worldpoints(void* h, int i, unsigned int* out, int cap)
{
    const std::vector<L3DPP::SfMCamera>& v = *(std::vector<L3DPP::SfMCamera>*)h;
    if (i < 0 || i >= (int)v.size()) return -1;
    int n = 0;
    for (unsigned int w : v[i].worldpoints) { if (n < cap) out[n] = w; ++n; }
    return n;
}
void l3dpp_intrinsics_from_focal(float focal, int w, int h, double* K9) { const L3DPP::Matrix3d K = L3DPP::intrinsicsFromFocal(focal, w, h); for (int k = 0; k < 9; ++k) K9[k] = K.m[k]; }
}
