"""Geometry sweep of the scoring sweep's chain kernel k_sw_score (threads per CTA, segments per CTA, staged entries per CTA).

    python tools/sweep_score.py build      # here (no GPU): one libl3d_b200_sw_<tag>.so per variant
    python tools/sweep_score.py run        # on the GPU box: chain time of every variant at configs[3] -> gpurun_out/sweep_score.json
"""
import json, os, re, subprocess, sys
sys.path.insert(0, ".")
VARIANTS = {
    "T256_S4_C512": [],
    "T256_S3_C384": ["SW_THREADS=256", "SW_SEGS=3", "SW_CAP_GPU=384"],
    "T128_S2_C256": ["SW_THREADS=128", "SW_SEGS=2", "SW_CAP_GPU=256"],
    "T192_S3_C384": ["SW_THREADS=192", "SW_SEGS=3", "SW_CAP_GPU=384"],
    "T320_S4_C512": ["SW_THREADS=320", "SW_SEGS=4", "SW_CAP_GPU=512"],
    "T256_S5_C640": ["SW_THREADS=256", "SW_SEGS=5", "SW_CAP_GPU=640"],
}
if sys.argv[1] == "build":
    from line3dpp_b200 import build
    for tag, d in VARIANTS.items():
        print(tag, build.build(defines=d, out=f"libl3d_b200_sw_{tag}.so"), flush=True)
elif sys.argv[1] == "run":
    res = {}
    for tag in VARIANTS:
        so = os.path.abspath(f"line3dpp_b200/libl3d_b200_sw_{tag}.so")
        if not os.path.exists(so):
            continue
        env = dict(os.environ, L3D_LIB=so, L3D_SWEEP_TIMING="1")
        r = subprocess.run([sys.executable, "tools/run_pipeline.py", "1000", "3000", "ring5", "0"], env=env, capture_output=True, text=True)
        m = re.findall(r"set-up ([\d.]+) ms, chain ([\d.]+) ms", r.stderr)
        res[tag] = {"setup_ms": float(m[-1][0]), "chain_ms": float(m[-1][1])} if m else {"error": r.stderr[-300:]}
        print(tag, res[tag], flush=True)
    json.dump({"workload": "configs[3]: 1000 views x 3000 segments, ring +-5", "variants": res}, open("gpurun_out/sweep_score.json", "w"), indent=1)
