"""How do the GPU box's host cores scale? (oracle matchingCPU port, 1..N threads)"""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np
from line3dpp_b200 import synth
from oracle import pyoracle as po
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    print("cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e:
    print("cpu.max n/a", e)
sc = synth.make_scene(4, 3000, 1004, "ring1")
RtKinv, C = synth.camera_blocks(sc)
F = synth.fundamental(sc.K[0], sc.R[0], sc.t[0], sc.K[1], sc.R[1], sc.t[1])
for thr in (1, 4, 16, 32, 64, 128):
    if thr > (os.cpu_count() or 1): break
    po.set_threads(thr)
    best = 1e9
    for rep in range(4):
        c, o, tot, ms = po.match_lines(po.lib().orc_match_lines_f64, sc.segs[0], sc.segs[1], F, RtKinv[0], RtKinv[1], C[0], C[1], 0, 1, 0.25, 10, f64=True)
        best = min(best, ms)
    print(f"{thr:4d} threads: {best:8.1f} ms  {9e6/best*1e3:.3e} pair-evals/s")
