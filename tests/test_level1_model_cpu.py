"""The level-1 pre-filter of k_match_topk never rejects a match of the exhaustive CPU oracle: checked on the numpy restatement of its
arithmetic (tests/level1_model.py) for the epipole at infinity, inside the image, near its border and with a rolled camera.  The GPU
suite (test_match_gpu.py::test_level1_prefilter_never_drops) runs the same geometries through the real kernels."""
import numpy as np
import pytest

from tests import level1_model as m
from tests import util


@pytest.mark.parametrize("kind", ["sideways", "forward", "edge", "rolled"])
@pytest.mark.parametrize("epi", [0.25, 0.05])
def test_arc_model_keeps_every_oracle_match(oracle, kind, epi):
    sc = util.two_view_scene(kind, 400, 31)
    checked = narrow_checked = 0
    for s, t in ((0, 1), (1, 0)):
        pi = util.pair_inputs(sc, s, t)
        oc, oo, _, _ = oracle.match_lines(oracle.lib().orc_match_lines_f32, pi["ls"], pi["lt"], pi["F"], pi["Rs"], pi["Rt"], pi["Cs"], pi["Ct"], s, t, epi, 10)
        u, v, _ = m.basis(pi["F"])
        arcs = [m.target_arc(u, v, q, 1.0 / epi + 0.5) for q in pi["lt"]]
        for r in range(len(oc)):
            p = pi["ls"][r]
            k1, off1 = m.line_kappa(u, v, m.epipolar_line(pi["F"], p[0], p[1]))
            k2, off2 = m.line_kappa(u, v, m.epipolar_line(pi["F"], p[2], p[3]))
            for i in range(oc[r]):
                a = arcs[int(oo[r, i]["tgt_seg"])]
                checked += 1
                if a is None or off1 or off2:
                    continue                      # always a candidate
                assert m.arc_may_match(a, k1, k2), (kind, s, t, r, int(oo[r, i]["tgt_seg"]))
                c = m.arc_class(a[1])
                if c < m.NCLS:
                    assert m.in_window(a[0], k1, k2, 1 << (m.CLS0 + c)), (kind, s, t, r, int(oo[r, i]["tgt_seg"]))
                    narrow_checked += 1
        if kind == "sideways" and epi == 0.25:    # ... and it is a filter: most cells of a row are rejected
            passed = total = 0
            for r in range(0, len(oc), 20):
                p = pi["ls"][r]
                k1, _ = m.line_kappa(u, v, m.epipolar_line(pi["F"], p[0], p[1]))
                k2, _ = m.line_kappa(u, v, m.epipolar_line(pi["F"], p[2], p[3]))
                passed += sum(1 for a in arcs if a is None or m.arc_may_match(a, k1, k2))
                total += len(arcs)
            assert passed < 0.3 * total
    assert checked > 1000 and (narrow_checked > 500 or kind == "forward")


def _check_pair(oracle, sc, epi, tag, knn=10):
    """knn <= 0: every cell above epi_overlap is a match (the reference's keep-all mode) - the strongest form of the check"""
    checked = 0
    for s, t in ((0, 1), (1, 0)):
        pi = util.pair_inputs(sc, s, t)
        if len(pi["ls"]) == 0 or len(pi["lt"]) == 0:
            continue
        oc, oo, _, _ = oracle.match_lines(oracle.lib().orc_match_lines_f32, pi["ls"], pi["lt"], pi["F"], pi["Rs"], pi["Rt"], pi["Cs"], pi["Ct"], s, t, epi, knn)
        u, v, _ = m.basis(pi["F"])
        arcs = [m.target_arc(u, v, q, 1.0 / epi + 0.5) for q in pi["lt"]]
        for r in range(len(oc)):
            if oc[r] == 0:
                continue
            p = pi["ls"][r]
            k1, off1 = m.line_kappa(u, v, m.epipolar_line(pi["F"], p[0], p[1]))
            k2, off2 = m.line_kappa(u, v, m.epipolar_line(pi["F"], p[2], p[3]))
            for i in range(oc[r]):
                a = arcs[int(oo[r, i]["tgt_seg"])]
                checked += 1
                if a is None or off1 or off2:
                    continue
                assert m.arc_may_match(a, k1, k2), (tag, s, t, r, int(oo[r, i]["tgt_seg"]))
                c = m.arc_class(a[1])
                if c < m.NCLS:
                    assert m.in_window(a[0], k1, k2, 1 << (m.CLS0 + c)), (tag, s, t, r, int(oo[r, i]["tgt_seg"]))
    return checked


def test_arc_model_random_camera_pairs(oracle):
    """40 random two-view geometries (baselines from pure sideways to pure forward, rotations up to ~35 degrees about a random axis,
    any roll): the model keeps every match of the exhaustive oracle"""
    rng = np.random.default_rng(77)
    total = 0
    for trial in range(40):
        def rot():
            ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
            ang = rng.uniform(0, 0.6) if trial % 4 else rng.uniform(0, 0.05)
            Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
            roll = rng.uniform(-np.pi, np.pi) if trial % 3 == 0 else 0.0
            Rz = np.array([[np.cos(roll), -np.sin(roll), 0], [np.sin(roll), np.cos(roll), 0], [0, 0, 1.0]])
            return Rz @ (np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx)
        C0 = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), -4.2])
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        if trial % 5 == 0: d = np.array([0.0, 0.0, 1.0])          # pure forward motion
        if trial % 5 == 1: d = np.array([1.0, 0.0, 0.0])          # pure sideways motion
        C1 = C0 + d * rng.uniform(0.05, 0.9)
        sc = util.two_view_scene([(np.eye(3), tuple(C0)), (rot(), tuple(C1))], 120, 100 + trial)
        total += _check_pair(oracle, sc, 0.25, trial, knn=0 if trial % 2 else 10)
    assert total > 3000
