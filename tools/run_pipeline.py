"""Run the whole hot path through L3DPP::Line3D on a synthetic scene and print the stage timings / counters.
usage: python tools/run_pipeline.py V N neighbours [diffusion 0/1] [collinearity_t] [use_ceres 0/1] [collinear scene 0/1]"""
import sys, time, json
sys.path.insert(0, ".")
from line3dpp_b200 import synth, line3d

V, N, nb = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
arg = lambda i, d: sys.argv[i] if len(sys.argv) > i else d
diff, collin, ceres, cscene = arg(4, "0") == "1", float(arg(5, "-1")), arg(6, "0") == "1", arg(7, "0") == "1"
t0 = time.time(); sc = synth.make_scene(V, N, 1004, nb, collinear=cscene); t_scene = time.time() - t0
L = line3d.Line3D(neighbors_by_worldpoints=False, use_gpu=True)
t0 = time.time(); L.add_scene(sc); t_add = time.time() - t0
for rep in range(2):
    t0 = time.time(); L.match_images(); t_match = time.time() - t0
for rep in range(2):
    t0 = time.time(); L.reconstruct_3d_lines(3, diff, collin, ceres); t_rec = time.time() - t0
st = L.stats()
st.update(scene_s=t_scene, add_s=t_add, matchImages_s=t_match, reconstruct_s=t_rec, diffusion=diff, collinearity_t=collin, use_ceres=ceres)
print(json.dumps(st))
