"""CPU fuzz of the level-1 pre-filter model (tests/level1_model.py = the arithmetic of k_pair_arcs / arc_may_match / the class windows of
k_match_topk) against the exhaustive CPU oracle on random two-view geometries.  No GPU needed.

    python tools/fuzz_level1_model.py <seed> <trials> [hard]

default: baselines from pure sideways to pure forward motion, rotations up to 40 degrees about a random axis, any roll, epi_overlap 0.25
and 0.1;  hard: baselines down to 1e-4 scene units (F dominated by rounding), image coordinates scaled by 0.25 / 1 / 3.
Every cell above epi_overlap counts (keep-all mode).  Round 2: 1700 geometries, 1.2e7 oracle matches, no match outside its arc test or window."""
import dataclasses
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import pyoracle as oracle                      # noqa: E402
from tests import util                                     # noqa: E402
from tests.test_level1_model_cpu import _check_pair        # noqa: E402

np.seterr(all="ignore")


def default_run():
    seed0=int(sys.argv[1]); ntr=int(sys.argv[2])
    rng=np.random.default_rng(seed0)
    total=0; fails=0
    for trial in range(ntr):
        ax=rng.normal(size=3); ax/=np.linalg.norm(ax)
        ang=rng.uniform(0,0.7) if trial%4 else rng.uniform(0,0.02)
        Kx=np.array([[0,-ax[2],ax[1]],[ax[2],0,-ax[0]],[-ax[1],ax[0],0]])
        roll=rng.uniform(-np.pi,np.pi) if trial%3==0 else 0.0
        Rz=np.array([[np.cos(roll),-np.sin(roll),0],[np.sin(roll),np.cos(roll),0],[0,0,1.0]])
        R=Rz@(np.eye(3)+np.sin(ang)*Kx+(1-np.cos(ang))*Kx@Kx)
        C0=np.array([rng.uniform(-0.3,0.3),rng.uniform(-0.3,0.3),-4.2])
        d=rng.normal(size=3); d/=np.linalg.norm(d)
        if trial%5==0: d=np.array([0,0,1.0])
        if trial%5==1: d=np.array([1.0,0,0])
        if trial%5==2: d=np.array([0,1.0,0])
        C1=C0+d*rng.uniform(0.01,1.2)
        sc=util.two_view_scene([(np.eye(3),tuple(C0)),(R,tuple(C1))],300,seed0*1000+trial)
        for epi in (0.25,0.1):
            try:
                total+=_check_pair(oracle,sc,epi,(trial,epi),knn=0)
            except AssertionError as e:
                fails+=1; print("FAIL",e)
    print("checked",total,"fails",fails)
    


def hard_run():
    seed0=int(sys.argv[1]); ntr=int(sys.argv[2])
    rng=np.random.default_rng(seed0)
    total=0; fails=0
    for trial in range(ntr):
        ax=rng.normal(size=3); ax/=np.linalg.norm(ax)
        ang=rng.uniform(0,0.5)
        Kx=np.array([[0,-ax[2],ax[1]],[ax[2],0,-ax[0]],[-ax[1],ax[0],0]])
        R=np.eye(3)+np.sin(ang)*Kx+(1-np.cos(ang))*Kx@Kx
        C0=np.array([rng.uniform(-0.3,0.3),rng.uniform(-0.3,0.3),-4.2])
        d=rng.normal(size=3); d/=np.linalg.norm(d)
        if trial%3==0: d=np.array([0,0,1.0])
        bl=[1e-4,1e-3,0.01,0.3][trial%4]          # tiny baselines: F ~ noise-dominated
        C1=C0+d*bl
        sc=util.two_view_scene([(np.eye(3),tuple(C0)),(R,tuple(C1))],250,seed0*1000+trial)
        sfac=[0.25,1.0,3.0][trial%3]
        Sm=np.diag([sfac,sfac,1.0])
        sc=dataclasses.replace(sc,K=np.array([Sm@k for k in sc.K]),segs=[np.ascontiguousarray((s_*sfac).astype(np.float32)) for s_ in sc.segs])
        for epi in (0.25,):
            try:
                total+=_check_pair(oracle,sc,epi,(trial,epi,bl,sfac),knn=0)
            except AssertionError as e:
                fails+=1; print("FAIL",e)
    print("checked",total,"fails",fails)
    


if __name__ == "__main__":
    hard_run() if len(sys.argv) > 3 and sys.argv[3] == "hard" else default_run()
