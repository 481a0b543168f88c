"""Quick on-GPU probe: time k_match_topk / k_match_dense on a synthetic workload (not the bench contract)."""
import sys, time, json
import numpy as np
import torch
sys.path.insert(0, ".")
from line3dpp_b200 import synth, capi
from tests import util

V, N, nb = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
t0 = time.time(); sc = synth.make_scene(V, N, 1004, nb); print("scene s", time.time() - t0, flush=True)
ctx = capi.Context(0)
descs = util.scene_descs(sc)
ctx.set_views(descs, sc.segs)
pairs = synth.view_pairs(sc.neighbors)
F = util.pair_F(sc, pairs)
st = torch.cuda.ExternalStream(ctx.stream)
for it in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st); ctx.match_pairs(pairs, F, 0.25, 10); e1.record(st)
    ctx.sync(); ms = e0.elapsed_time(e1)
    pe = ctx.match_pair_evals()
    print(f"match_pairs: {ms:.2f} ms  pairs={len(pairs)} pair_evals={pe:.3e}  -> {pe/ms*1e3:.3e} pair-evals/s", flush=True)
counts, total = ctx.match_counts()
print("matches", total, "per row", total / max(len(counts), 1), "frac of evals", total / pe)
if len(sys.argv) > 4 and sys.argv[4] == "nodense":
    sys.exit(0)
# dense kernel
Ns, Nt = len(sc.segs[0]), len(sc.segs[1])
dep = torch.empty(Ns * Nt * 4, device="cuda"); ov = torch.empty(Ns * Nt, device="cuda")
for nof in (False, True):
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); 
        for rep in range(10): ctx.match_dense(0, 1, F[0], 0.25, Ns, Nt, nofilter=nof, dev_ptrs=(dep.data_ptr(), ov.data_ptr()))
        e1.record(st); ctx.sync(); ms = e0.elapsed_time(e1) / 10
    print(f"dense nofilter={nof}: {ms:.3f} ms per {Ns}x{Nt} -> {Ns*Nt/ms*1e3:.3e} cells/s, {Ns*Nt*20/ms*1e3/1e9:.1f} GB/s written")
