// Shim for the reference's generated configLIBS.h (configLIBS.h.in:4-7).
// Test infrastructure only: lets oracle/Makefile compile /root/reference sources in place.
//   default                       : the accelerator files alone (ref_harness.cu): CUDA + OpenMP
//   -DL3D_SHIM_FULL               : whole pipeline, CPU code path only (ref_full_harness.cu with g++)
//   -DL3D_SHIM_FULL -DL3D_SHIM_FULL_CUDA : whole pipeline with the CUDA code path (nvcc)
// The full builds leave L3DPP_OPENMP off (single-threaded = the reference's deterministic result order) and
// L3DPP_CERES off (Ceres is not installed).
#ifndef L3D_ORACLE_SHIM_CONFIGLIBS_H
#define L3D_ORACLE_SHIM_CONFIGLIBS_H
#ifdef L3D_SHIM_FULL
#define L3DPP_OPENCV3 1
#ifdef L3D_SHIM_FULL_CUDA
#define L3DPP_CUDA 1
#endif
#else
#define L3DPP_CUDA 1
#define L3DPP_OPENMP 1
#endif
#endif
