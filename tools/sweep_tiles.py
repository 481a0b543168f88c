"""Tile-size sweep of k_match_topk (BASELINE.json configs[2]: synthetic 200 views x 2000 segments, dense visibility).

    python tools/sweep_tiles.py build        # here (no GPU): compiles one libl3d_b200_<tag>.so per variant
    python tools/sweep_tiles.py run          # on the GPU box: times every variant, writes gpurun_out/sweep_tiles.json
"""
import json, os, subprocess, sys
sys.path.insert(0, ".")
VARIANTS = {     # round 2: the kernel scans arc windows (no MK_T any more); what is left to vary is rows per warp / CTAs per SM / ring depth
    "base_R8_B3": [],
    "R4_B3": ["MK_RPW=4"],
    "R16_B2": ["MK_RPW=16", "MK_MINB=2"],
    "R8_B2": ["MK_MINB=2"],
    "R8_B4_S2_CAP24": ["MK_MINB=4", "MK_STAGES=2", "MK_CAP=24"],      # measured: 16 % slower (profiles/README.md)
    "R8_B3_S2": ["MK_STAGES=2"],                                        # measured: 13 % slower
}
if sys.argv[1] == "build":
    from line3dpp_b200 import build
    for tag, d in VARIANTS.items():
        print(tag, build.build(defines=d, out=f"libl3d_b200_{tag}.so"), flush=True)
elif sys.argv[1] == "run":
    V, N = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (200, 2000)
    res = {}
    for tag in VARIANTS:
        so = os.path.abspath(f"line3dpp_b200/libl3d_b200_{tag}.so")
        if not os.path.exists(so):
            continue
        env = dict(os.environ, L3D_LIB=so)
        out = subprocess.run([sys.executable, "tools/probe_match.py", str(V), str(N), "dense", "nodense"], env=env, capture_output=True, text=True).stdout
        best = max([float(l.split("->")[1].split()[0]) for l in out.splitlines() if l.startswith("match_pairs")] or [0.0])
        res[tag] = best
        print(f"{tag:24s} {best:.3e} pair-evals/s", flush=True)
    json.dump({"workload": f"{V} views x {N} segments dense", "pair_evals_per_s": res}, open("gpurun_out/sweep_tiles.json", "w"), indent=1)
