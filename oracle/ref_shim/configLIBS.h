// Shim for the reference's generated configLIBS.h (configLIBS.h.in:4-7).
// Test infrastructure only: lets oracle/Makefile compile /root/reference sources in place.
#ifndef L3D_ORACLE_SHIM_CONFIGLIBS_H
#define L3D_ORACLE_SHIM_CONFIGLIBS_H
#define L3DPP_CUDA 1
#define L3DPP_OPENMP 1
#endif
