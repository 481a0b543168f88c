"""ctypes binding of libl3d_b200.so (include/l3d_capi.h).  Thin: no arithmetic happens in Python.

The library is REQUIRED: importing this module builds it if the .so is missing and raises if that fails; creating a
context raises if there is no usable CUDA device.  There is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))


class ViewDesc(C.Structure):
    _fields_ = [("cam_id", C.c_uint32), ("width", C.c_int32), ("height", C.c_int32), ("nseg", C.c_int32),
                ("RtKinv", C.c_float * 9), ("C", C.c_float * 3), ("RtKinv_d", C.c_double * 9), ("C_d", C.c_double * 3),
                ("k", C.c_float), ("median_depth", C.c_float)]


REC_DT = np.dtype([("tgt_seg", "<u4"), ("overlap", "<f4"), ("d_p1", "<f4"), ("d_p2", "<f4"), ("d_q1", "<f4"),
                   ("d_q2", "<f4")])
MATCH_DT = np.dtype([("src_cam", "<u4"), ("src_seg", "<u4"), ("tgt_cam", "<u4"), ("tgt_seg", "<u4"),
                     ("overlap", "<f4"), ("score3D", "<f4"), ("d_p1", "<f4"), ("d_p2", "<f4"),
                     ("d_q1", "<f4"), ("d_q2", "<f4")])

_lib = None


class L3DError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        path = os.environ.get("L3D_LIB") or _build.build()     # L3D_LIB: load a tile-sweep variant instead
        L = C.CDLL(path)
        L.l3d_last_error.restype = C.c_char_p
        L.l3d_stream.restype = C.c_void_p
        for n in ("l3d_launch_count", "l3d_match_total_rows", "l3d_match_pair_evals", "l3d_get_match_counts",
                  "l3d_get_matches_csr", "l3d_collinear_total", "l3d_get_collinear"):
            getattr(L, n).restype = C.c_longlong
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def make_view_descs(cam_ids, widths, heights, nsegs, RtKinv_d, C_untranslated_d, C_work_d, k, median_depth):
    """Pack camera blocks (already computed by the host side, in double) into l3d_view_desc[]."""
    V = len(cam_ids)
    arr = (ViewDesc * V)()
    for v in range(V):
        d = arr[v]
        d.cam_id, d.width, d.height, d.nseg = int(cam_ids[v]), int(widths[v]), int(heights[v]), int(nsegs[v])
        R = np.asarray(RtKinv_d[v], np.float64).reshape(9)
        d.RtKinv[:] = R.astype(np.float32).tolist()
        d.RtKinv_d[:] = R.tolist()
        d.C[:] = np.asarray(C_untranslated_d[v], np.float64).astype(np.float32).tolist()
        d.C_d[:] = np.asarray(C_work_d[v], np.float64).tolist()
        d.k = float(k[v])
        d.median_depth = float(median_depth[v])
    return arr


class Context:
    """One l3d_ctx (one GPU)."""

    def __init__(self, device: int = 0):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.l3d_ctx_create(int(device), C.byref(h))
        if rc != 0 or not h:
            raise L3DError(f"l3d_ctx_create(device={device}) failed with {rc}: no usable CUDA device "
                           "(libl3d_b200 has no CPU fallback)")
        self.h = h
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.l3d_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _chk(self, rc, what):
        if rc < 0:
            raise L3DError(f"{what} failed ({rc}): {self.L.l3d_last_error(self.h).decode()}")
        return rc

    @property
    def stream(self) -> int:
        return int(self.L.l3d_stream(self.h) or 0)

    def sync(self):
        self._chk(self.L.l3d_sync(self.h), "l3d_sync")

    def launch_count(self) -> int:
        return int(self.L.l3d_launch_count(self.h))

    # ---- views
    def set_views(self, descs, segs_list):
        V = len(descs)
        segs = [np.ascontiguousarray(s, np.float32) for s in segs_list]
        ptrs = (C.c_void_p * V)(*[s.ctypes.data for s in segs])
        self._keep = [descs, segs]
        self._chk(self.L.l3d_set_views(self.h, V, descs, ptrs), "l3d_set_views")

    def set_views_flat(self, descs, segs_flat_ptr: int, on_device: bool):
        self._keep = [descs]
        self._chk(self.L.l3d_set_views_flat(self.h, len(descs), descs, C.c_void_p(segs_flat_ptr), int(on_device)),
                  "l3d_set_views_flat")

    def update_view_params(self, descs):
        self._chk(self.L.l3d_update_view_params(self.h, len(descs), descs), "l3d_update_view_params")

    # ---- matching
    def match_pairs(self, pairs, F, epi_overlap=0.25, knn=10):
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        F = np.ascontiguousarray(F, np.float32).reshape(-1, 9)
        assert len(pairs) == len(F)
        self._chk(self.L.l3d_match_pairs(self.h, len(pairs), _p(pairs), _p(F), C.c_float(epi_overlap), int(knn)),
                  "l3d_match_pairs")
        self._knn = int(self.L.l3d_match_stride(self.h))

    def match_pairs_range(self, pairs, F, first, last, epi_overlap=0.25, knn=10):
        """sharded form: stage all pairs, evaluate only [first, last) here (include/l3d_capi.h)"""
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        F = np.ascontiguousarray(F, np.float32).reshape(-1, 9)
        assert len(pairs) == len(F)
        self._chk(self.L.l3d_match_pairs_range(self.h, len(pairs), _p(pairs), _p(F), C.c_float(epi_overlap), int(knn), int(first), int(last)),
                  "l3d_match_pairs_range")
        self._knn = int(self.L.l3d_match_stride(self.h))

    def match_pairs_f64(self, pairs, Fd, epi_overlap=0.25, knn=10, first=0, last=None):
        """REF_CPU semantics: matchingCPU's double arithmetic (line3D.cc:900-1015) with double fundamental matrices"""
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        Fd = np.ascontiguousarray(Fd, np.float64).reshape(-1, 9)
        assert len(pairs) == len(Fd)
        last = len(pairs) if last is None else last
        self._chk(self.L.l3d_match_pairs_f64(self.h, len(pairs), _p(pairs), _p(Fd), C.c_float(epi_overlap), int(knn), int(first), int(last)),
                  "l3d_match_pairs_f64")
        self._knn = int(self.L.l3d_match_stride(self.h))

    def match_pairs_host(self, pairs, F, counts_ptr: int, recs_ptr: int, epi_overlap=0.25, knn=10, chunks=8):
        """match + overlapped download into host memory at the given addresses (pinned for real overlap); asynchronous"""
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        F = np.ascontiguousarray(F, np.float32).reshape(-1, 9)
        assert len(pairs) == len(F)
        self._chk(self.L.l3d_match_pairs_host(self.h, len(pairs), _p(pairs), _p(F), C.c_float(epi_overlap), int(knn),
                                              C.c_void_p(counts_ptr), C.c_void_p(recs_ptr), int(chunks)), "l3d_match_pairs_host")
        self._knn = int(knn)

    def match_device_buffers(self):
        """(counts_ptr, recs_ptr) device addresses of the last match result"""
        a, b = C.c_void_p(), C.c_void_p()
        self._chk(self.L.l3d_match_device_buffers(self.h, C.byref(a), C.byref(b)), "l3d_match_device_buffers")
        return a.value or 0, b.value or 0

    def pair_row_offsets(self, num_pairs):
        out = np.zeros(num_pairs + 1, np.int64)
        self._chk(self.L.l3d_pair_row_offsets(self.h, _p(out)), "l3d_pair_row_offsets")
        return out

    def match_total_rows(self) -> int:
        return int(self.L.l3d_match_total_rows(self.h))

    def match_pair_evals(self) -> int:
        return int(self.L.l3d_match_pair_evals(self.h))

    def match_counts(self):
        out = np.zeros(max(self.match_total_rows(), 0), np.int32)
        total = self._chk(self.L.l3d_get_match_counts(self.h, _p(out)), "l3d_get_match_counts")
        return out, int(total)

    def pair_matches(self, pair: int, ns: int):
        counts = np.zeros(ns, np.int32)
        recs = np.zeros((ns, self._knn), REC_DT)
        self._chk(self.L.l3d_get_pair_matches(self.h, int(pair), _p(counts), _p(recs)), "l3d_get_pair_matches")
        return counts, recs

    def matches_csr(self, row_ptr=None, recs=None):
        rows = self.match_total_rows()
        if row_ptr is None:
            row_ptr = np.zeros(rows + 1, np.int64)
        cap = 0 if recs is None else len(recs)
        total = self._chk(self.L.l3d_get_matches_csr(self.h, _p(row_ptr), _p(recs), C.c_longlong(cap)),
                          "l3d_get_matches_csr")
        if recs is None or total > cap:
            recs = np.zeros(total, REC_DT)
            total = self._chk(self.L.l3d_get_matches_csr(self.h, _p(row_ptr), _p(recs), C.c_longlong(len(recs))),
                              "l3d_get_matches_csr")
        return row_ptr, recs[:total]

    def matches_csr_raw(self, row_ptr_ptr: int, recs_ptr: int, capacity: int) -> int:
        return int(self._chk(self.L.l3d_get_matches_csr(self.h, C.c_void_p(row_ptr_ptr), C.c_void_p(recs_ptr),
                                                        C.c_longlong(capacity)), "l3d_get_matches_csr"))

    def rdd(self, n, ei, ej, ew, iters=10):
        """l3d_rdd: returns (out_i, out_j, out_w, device ms of the diffusion iterations)"""
        ei = np.ascontiguousarray(ei, np.int32); ej = np.ascontiguousarray(ej, np.int32); ew = np.ascontiguousarray(ew, np.float32)
        oi, oj, ow = np.empty_like(ei), np.empty_like(ej), np.empty_like(ew)
        ms = C.c_float(0)
        self._chk(self.L.l3d_rdd(self.h, int(n), C.c_longlong(len(ei)), _p(ei), _p(ej), _p(ew), int(iters), _p(oi), _p(oj), _p(ow),
                                 C.byref(ms)), "l3d_rdd")
        return oi, oj, ow, ms.value

    def find_collinear(self, dist_t: float, semantics: int = 0):
        """l3d_find_collinear for all views (semantics 0 = REF_GPU float kernel, 1 = REF_CPU); dist_t <= 0 switches the links off"""
        self._chk(self.L.l3d_find_collinear(self.h, C.c_float(dist_t), int(semantics)), "l3d_find_collinear")
        return int(self.L.l3d_collinear_total(self.h))

    def collinear(self, view: int, nseg: int):
        """View::collin_ of one view as CSR (row_ptr[nseg+1], idx)"""
        row_ptr = np.zeros(nseg + 1, np.int64)
        n = self._chk(self.L.l3d_get_collinear(self.h, int(view), _p(row_ptr), None, C.c_longlong(0)), "l3d_get_collinear")
        idx = np.zeros(max(int(n), 1), np.int32)
        if n > 0:
            self._chk(self.L.l3d_get_collinear(self.h, int(view), _p(row_ptr), _p(idx), C.c_longlong(int(n))), "l3d_get_collinear")
        return row_ptr, idx[:int(n)]

    def optimize_lines(self, p1p2, res_ptr, res_cam, res_xy, cams, max_iter=250):
        """l3d_optimize_lines: returns (p1p2_out[L,6], valid[L], summary[8])"""
        f64 = lambda a: np.ascontiguousarray(a, np.float64)
        p1p2, res_xy, cams = f64(p1p2), f64(res_xy), f64(cams)
        res_ptr = np.ascontiguousarray(res_ptr, np.int64); res_cam = np.ascontiguousarray(res_cam, np.int32)
        L = len(p1p2)
        out = np.zeros((L, 6), np.float64); valid = np.zeros(L, np.int32); summ = np.zeros(8, np.float64)
        self._chk(self.L.l3d_optimize_lines(self.h, L, _p(p1p2), _p(res_ptr), _p(res_cam), _p(res_xy), len(cams), _p(cams), int(max_iter),
                                            _p(out), _p(valid), _p(summ)), "l3d_optimize_lines")
        return out, valid, summ

    def fp32_peak_tflops(self) -> float:
        v = C.c_double(0)
        self._chk(self.L.l3d_fp32_peak_probe(self.h, C.byref(v)), "l3d_fp32_peak_probe")
        return v.value

    def match_dense_pairs(self, pairs, F, epi_overlap, depth_ptrs, overlap_ptrs):
        """dense contract for many view pairs in one launch; depth_ptrs / overlap_ptrs: device addresses per pair"""
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        F = np.ascontiguousarray(F, np.float32).reshape(-1, 9)
        dp = (C.c_void_p * len(pairs))(*[int(x) for x in depth_ptrs])
        op = (C.c_void_p * len(pairs))(*[int(x) for x in overlap_ptrs])
        self._chk(self.L.l3d_match_dense_pairs(self.h, len(pairs), _p(pairs), _p(F), C.c_float(epi_overlap), dp, op), "l3d_match_dense_pairs")

    def match_dense(self, src_view, tgt_view, F, epi_overlap, ns, nt, nofilter=False, dev_ptrs=None):
        F = np.ascontiguousarray(F, np.float32).reshape(9)
        fn = self.L.l3d_match_dense_nofilter if nofilter else self.L.l3d_match_dense
        if dev_ptrs is not None:
            self._chk(fn(self.h, int(src_view), int(tgt_view), _p(F), C.c_float(epi_overlap), C.c_void_p(dev_ptrs[0]),
                         C.c_void_p(dev_ptrs[1]), 1), "l3d_match_dense")
            return None
        dep = np.empty((ns, nt, 4), np.float32)
        ov = np.empty((ns, nt), np.float32)
        self._chk(fn(self.h, int(src_view), int(tgt_view), _p(F), C.c_float(epi_overlap), _p(dep), _p(ov), 0),
                  "l3d_match_dense")
        return dep, ov
