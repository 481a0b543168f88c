"""On-GPU probe of l3d_optimize_lines on the reference's own before/after fixture clusters (2489 lines, 17603 residuals),
replicated to larger problems; the oracle (single-threaded restatement of the Ceres algorithm) is timed beside it."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from line3dpp_b200 import capi
from tests import nvm_util as nu

before, after, ptr, res = nu.load_opt_pairs()
cams, shift = nu.optimizer_inputs(nu.load_inputs())
b = before + np.tile(shift, 2)
ctx = capi.Context(0)
for rep in (1, 40):
    p = np.tile(b, (rep, 1))
    n = np.diff(ptr)
    pp = np.concatenate([[0], np.cumsum(np.tile(n, rep))])
    cam = np.tile(res[:, 0].astype(np.int32), rep); xy = np.tile(res[:, 2:6], (rep, 1))
    for it in range(3):
        t0 = time.time(); out, valid, summ = ctx.optimize_lines(p, pp, cam, xy, cams, 250); dt = time.time() - t0
    print(f"l3d_optimize_lines: {len(p)} lines, {len(cam)} residuals: {dt*1e3:.1f} ms wall incl. H2D/D2H, {int(summ[0])} LM iterations, "
          f"{int(summ[7])} kernels, cost {summ[1]:.1f} -> {summ[2]:.1f}")
    if rep == 1:
        from oracle import pyoracle as po
        t0 = time.time(); po.optimize_lines(po.lib().orc_optimize_lines, p, pp, cam, xy, cams, 250); dt = time.time() - t0
        print(f"oracle (1 CPU thread): {dt*1e3:.1f} ms")
