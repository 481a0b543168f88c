"""Python face of the C++ L3DPP::Line3D mirror (include/line3d.h, csrc/line3d_host.cc) via its C wrapper.

Used by the tests and tools; the arithmetic is all in libl3d_b200.so (GPU) — nothing is computed here.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import capi

SEG3D_DT = np.dtype([("line", "<i4"), ("_pad", "<i4"), ("p1", "<f8", 3), ("p2", "<f8", 3)])
RESID_DT = np.dtype([("line", "<i4"), ("cam", "<u4"), ("seg", "<u4")])


class Stats(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in ("view_pairs", "pair_evaluations", "matches_after_knn", "estimates", "affinity_entries",
                                            "affinity_rows", "clusters_total", "clusters_valid", "lines3D", "collinear_entries", "opt_iterations")] + \
               [(n, C.c_double) for n in ("ms_match", "ms_score", "ms_affinity", "ms_diffusion", "ms_cluster", "opt_cost_before", "opt_cost_after")]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Line3D:
    """addImage / matchImages / reconstruct3Dlines with the reference's argument meaning (line3D.h:80-166)."""

    def __init__(self, neighbors_by_worldpoints=True, use_gpu=True, device=0):
        self.L = capi.lib()
        self.L.l3dpp_create.restype = C.c_void_p
        self.L.l3dpp_ctx.restype = C.c_void_p
        self.L.l3dpp_last_error.restype = C.c_char_p
        for n in ("l3dpp_get_affinity", "l3dpp_get_segments3d", "l3dpp_get_residuals", "l3d_get_view_matches", "l3d_get_estimates"):
            getattr(self.L, n).restype = C.c_longlong
        self.h = C.c_void_p(self.L.l3dpp_create(int(neighbors_by_worldpoints), int(use_gpu), int(device)))
        if not self.h:
            raise capi.L3DError("L3DPP::Line3D: no usable CUDA device (no CPU fallback)")
        self.ctx = C.c_void_p(self.L.l3dpp_ctx(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.L.l3dpp_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _chk(self, rc, what):
        if rc < 0:
            raise capi.L3DError(f"{what}: {self.L.l3dpp_last_error(self.h).decode()}")
        return rc

    def add_image(self, cam, width, height, K, R, t, median_depth, wps_or_neighbors, segs):
        lst = np.ascontiguousarray(wps_or_neighbors, np.uint32)
        segs = np.ascontiguousarray(segs, np.float32)
        f64 = lambda a: np.ascontiguousarray(a, np.float64)
        self._chk(self.L.l3dpp_add_image(self.h, C.c_uint(int(cam)), int(width), int(height), _p(f64(K)), _p(f64(R)), _p(f64(t)),
                                         C.c_float(float(median_depth)), _p(lst), len(lst), _p(segs), len(segs)), "addImage")

    def add_scene(self, scene):
        for i in range(scene.num_views):
            self.add_image(scene.cam_ids[i], scene.width, scene.height, scene.K[i], scene.R[i], scene.t[i], scene.median_depth[i],
                           scene.neighbors[i], scene.segs[i])

    def match_images(self, sigma_p=2.5, sigma_a=10.0, num_neighbors=10, epi_overlap=0.25, knn=10, const_reg_depth=-1.0):
        self._chk(self.L.l3dpp_match_images(self.h, C.c_float(sigma_p), C.c_float(sigma_a), C.c_uint(num_neighbors), C.c_float(epi_overlap),
                                            C.c_int(knn), C.c_float(const_reg_depth)), "matchImages")

    def reconstruct_3d_lines(self, visibility_t=3, perform_diffusion=False, collinearity_t=-1.0, use_ceres=False, max_iter_ceres=250):
        self._chk(self.L.l3dpp_reconstruct_opt(self.h, C.c_uint(visibility_t), int(perform_diffusion), C.c_float(collinearity_t),
                                               int(use_ceres), C.c_uint(max_iter_ceres)), "reconstruct3Dlines")

    # ---- dumps
    def stats(self):
        s = Stats()
        self.L.l3dpp_stats(self.h, C.byref(s))
        return {n: getattr(s, n) for n, _ in Stats._fields_}

    def pairs(self):
        n = self.L.l3dpp_get_pairs(self.h, None, 0)
        a = np.zeros((n, 2), np.int32)
        self.L.l3dpp_get_pairs(self.h, _p(a), n)
        return a

    def view_info(self, cam):
        k, md = C.c_float(0), C.c_float(0)
        self.L.l3dpp_view_info(self.h, C.c_uint(int(cam)), C.byref(k), C.byref(md))
        return k.value, md.value

    def view_matches(self, cam, kept_only):
        vi = self.L.l3dpp_view_index(self.h, C.c_uint(int(cam)))
        n = self.L.l3d_get_view_matches(self.ctx, vi, int(kept_only), None, C.c_longlong(0))
        if n < 0:
            raise capi.L3DError("l3d_get_view_matches failed")
        a = np.zeros(n, capi.MATCH_DT)
        if n:
            self.L.l3d_get_view_matches(self.ctx, vi, int(kept_only), _p(a), C.c_longlong(n))
        return a

    def estimates(self):
        n = self.L.l3d_get_estimates(self.ctx, None, None, C.c_longlong(0))
        best, p = np.zeros(max(n, 0), capi.MATCH_DT), np.zeros((max(n, 0), 6), np.float64)
        if n > 0:
            self.L.l3d_get_estimates(self.ctx, _p(best), _p(p), C.c_longlong(n))
        return best, p

    def ctx_collinear(self, view_index, nseg):
        """View::collin_ of the view_index-th added view as CSR (row_ptr, idx), straight from the C ABI"""
        row_ptr = np.zeros(nseg + 1, np.int64)
        self.L.l3d_get_collinear.restype = C.c_longlong
        n = self.L.l3d_get_collinear(self.ctx, int(view_index), _p(row_ptr), None, C.c_longlong(0))
        if n < 0:
            raise capi.L3DError("l3d_get_collinear failed")
        idx = np.zeros(max(int(n), 1), np.int32)
        if n > 0:
            self.L.l3d_get_collinear(self.ctx, int(view_index), _p(row_ptr), _p(idx), C.c_longlong(int(n)))
        return row_ptr, idx[:int(n)]

    def affinity(self, raw=False):
        n = self.L.l3dpp_get_affinity(self.h, int(raw), None, None, None, C.c_longlong(0))
        ei, ej, ew = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float32)
        if n:
            self.L.l3dpp_get_affinity(self.h, int(raw), _p(ei), _p(ej), _p(ew), C.c_longlong(n))
        return ei, ej, ew

    def local2global(self):
        n = self.L.l3dpp_get_local2global(self.h, None, 0)
        a = np.zeros((n, 2), np.uint32)
        if n:
            self.L.l3dpp_get_local2global(self.h, _p(a), n)
        return a

    def segments3d(self):
        n = self.L.l3dpp_get_segments3d(self.h, None, C.c_longlong(0))
        a = np.zeros(n, SEG3D_DT)
        if n:
            self.L.l3dpp_get_segments3d(self.h, _p(a), C.c_longlong(n))
        return a

    def residuals(self):
        n = self.L.l3dpp_get_residuals(self.h, None, C.c_longlong(0))
        a = np.zeros(n, RESID_DT)
        if n:
            self.L.l3dpp_get_residuals(self.h, _p(a), C.c_longlong(n))
        return a

    def save_txt(self, folder):
        self._chk(self.L.l3dpp_save_txt(self.h, folder.encode()), "save3DLinesAsTXT")
