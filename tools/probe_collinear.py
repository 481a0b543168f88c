"""On-GPU probe of l3d_find_collinear (both passes + scan) on a synthetic scene with broken lines; compares the time of the
unmodified reference path (dense char matrix + D2H + host scan, oracle/_ref) on one view."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
from line3dpp_b200 import synth, capi
from tests import util

V, N = int(sys.argv[1]), int(sys.argv[2])
sc = synth.make_scene(V, N, 1004, "ring1", collinear=True)
ctx = capi.Context(0)
ctx.set_views(util.scene_descs(sc), sc.segs)
st = torch.cuda.ExternalStream(ctx.stream)
cells = sum(len(s) ** 2 for s in sc.segs)
for sem in (0, 1):
    for it in range(3):
        ctx.find_collinear(0.0, sem)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); total = ctx.find_collinear(2.0 + it * 1e-3, sem); e1.record(st); ctx.sync()
        ms = e0.elapsed_time(e1)
    print(f"find_collinear sem={sem}: {V} views x {N}: {ms:.3f} ms, {cells/ms*1e3:.3e} cells/s, {total} list entries ({total/cells*100:.4f} % of cells)")
try:
    from oracle import pyoracle as po
    ref = po.ref_lib("default")
    t0 = time.time(); Cm, kms = po.collinear(ref.ref_collinear, sc.segs[0], 2.0); wall = time.time() - t0
    t0 = time.time(); lists = [np.flatnonzero(Cm[i] == 1) for i in range(len(Cm))]; scan = time.time() - t0
    print(f"reference find_collinear_segments_GPU on ONE view of {N}: kernel {kms:.3f} ms, call incl. alloc + D2H {wall*1e3:.1f} ms, numpy row scan {scan*1e3:.1f} ms")
except Exception as e:   # noqa
    print("reference leg skipped:", e)
