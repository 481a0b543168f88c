"""Geometry sweep of the scoring sweep's chain kernel k_sw_score: segments per CTA (L3D_SW_SEGS overrides the per-view choice of
l3d_score_sweep, which picks the fewest segments per CTA that make a view's CTAs ONE resident wave).

    python tools/sweep_score.py            # on the GPU box: chain time per setting at configs[3] -> gpurun_out/sweep_score.json
"""
import json, os, re, subprocess, sys
res = {}
for segs in ["auto", 2, 3, 4, 5, 6, 8, 10, 12, 16]:
    env = dict(os.environ, L3D_SWEEP_TIMING="1")
    if segs != "auto":
        env["L3D_SW_SEGS"] = str(segs)
    r = subprocess.run([sys.executable, "tools/run_pipeline.py", "1000", "3000", "ring5", "0"], env=env, capture_output=True, text=True)
    m = re.findall(r"set-up ([\d.]+) ms, chain ([\d.]+) ms", r.stderr)
    res[str(segs)] = {"setup_ms": float(m[-1][0]), "chain_ms": float(m[-1][1])} if m else {"error": r.stderr[-300:]}
    print("segments per CTA", segs, res[str(segs)], flush=True)
json.dump({"workload": "configs[3]: 1000 views x 3000 segments, ring +-5; 256 threads per CTA", "segments_per_cta": res}, open("gpurun_out/sweep_score.json", "w"), indent=1)
