"""Golden vectors of the UNMODIFIED reference pipeline (line3D.cc + view.cc compiled verbatim into
oracle/_ref/libl3dref_full_cpu.so by oracle/Makefile) on the committed vsfm_result.nvm inputs (BASELINE.json configs[0]):
default parameters, CPU code path, single-threaded.   Run in the build container (needs /root/reference):

    make -f oracle/Makefile ref && python tests/golden/make_ref_full_golden.py

Writes tests/golden/ref_full_nvm_cpu_v1.npz: for every view the number of matches right after scoring and a SHA-256 over
their ids/overlaps/depths, the kept matches in full, k / median depth; the best estimates; the affinity matrix
(local ids, edges); the final 3D segments and residuals; the reference's own .txt output (SHA-256).
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po          # noqa: E402
from tests import nvm_util as nu           # noqa: E402


def digest_matches(m):
    """ids + overlap + depths (everything but the score, which may differ in the last bit between libm builds)"""
    h = hashlib.sha256()
    for f in ("src_cam", "src_seg", "tgt_cam", "tgt_seg", "overlap", "d_p1", "d_p2", "d_q1", "d_q2"):
        h.update(np.ascontiguousarray(m[f]).tobytes())
    return np.frombuffer(h.digest(), np.uint8)


def main():
    inp = nu.load_inputs()
    R = po.RefFullPipeline(True, False, "cpu")
    nu.add_all(R.add_view, inp)
    R.match_images()
    R.reconstruct(3, False)
    out = {}
    V = inp["V"]
    out["pairs"] = R.pairs()
    out["scored_count"] = np.array([len(R.scored(c)) for c in range(V)], np.int64)
    out["scored_sha"] = np.stack([digest_matches(R.scored(c)) for c in range(V)])
    out["scored_score_sum"] = np.array([float(R.scored(c)["score3D"].astype(np.float64).sum()) for c in range(V)])
    kept = np.concatenate([R.matches(c) for c in range(V)])
    out["kept_src_cam"] = kept["src_cam"].astype(np.uint8)
    out["kept_src_seg"] = kept["src_seg"].astype(np.uint16)
    out["kept_tgt_cam"] = kept["tgt_cam"].astype(np.uint8)
    out["kept_tgt_seg"] = kept["tgt_seg"].astype(np.uint16)
    out["kept_score3D"] = kept["score3D"]
    out["kept_sha"] = digest_matches(kept)
    out["view_info"] = np.array([R.view_info(c) for c in range(V)], np.float32)
    best, P = R.estimates()
    out["est_src"] = np.stack([best["src_cam"], best["src_seg"], best["tgt_cam"], best["tgt_seg"]], 1).astype(np.uint16)
    out["est_P"] = P
    ei, ej, ew = R.affinity_raw()
    out["aff_i"], out["aff_j"], out["aff_w"] = ei, ej, ew
    out["l2g"] = R.local2global().astype(np.uint16)
    s = R.segments3d()
    out["seg_line"] = s["line"]
    out["seg_p1p2"] = np.concatenate([s["p1"], s["p2"]], 1)
    r = R.residuals()
    out["res"] = np.stack([r["line"], r["cam"].astype(np.int32), r["seg"].astype(np.int32)], 1)
    with tempfile.TemporaryDirectory() as d:
        name = R.save(d, txt=True)
        txt = open(os.path.join(d, name + ".txt"), "rb").read()
    out["txt_sha"] = np.frombuffer(hashlib.sha256(txt).digest(), np.uint8)
    out["txt_name"] = np.array(name)
    path = os.path.join(ROOT, "tests", "golden", "ref_full_nvm_cpu_v1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", R.num_lines(), "lines,", len(s), "segments,", len(r), "residuals,", len(ei), "edges,", name)


if __name__ == "__main__":
    main()
