"""Builds tests/golden/ref_writers_v1.npz (run in the build container, where /root/reference exists): SHA-256 of the reference's
own result files testdata/Line3D++_ref/<name>.{txt,obj,stl} and the 3D segments of the .stl at its 7-digit precision.  The
writer test feeds the parsed lines back into L3DPP::Line3D and must reproduce the three files byte for byte."""
import hashlib
import os
import re

import numpy as np

REF = "/root/reference/testdata/Line3D++_ref"
NAME = "Line3D++__W_FULL__N_10__sigmaP_2.5__sigmaA_10__epiOverlap_0.25__kNN_10__vis_3"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    sha = {e: hashlib.sha256(open(os.path.join(REF, NAME + "." + e), "rb").read()).hexdigest() for e in ("txt", "obj", "stl")}
    v = [[float(x) for x in m.groups()] for m in re.finditer(r"vertex (\S+) (\S+) (\S+)", open(os.path.join(REF, NAME + ".stl")).read())]
    v = np.array(v).reshape(-1, 3, 3)
    assert np.array_equal(v[:, 0], v[:, 2])
    segs = np.concatenate([v[:, 0], v[:, 1]], 1)
    print(len(segs), "STL segments", sha)
    np.savez_compressed(os.path.join(OUT, "ref_writers_v1.npz"), name=np.array(NAME), sha_txt=np.array(sha["txt"]), sha_obj=np.array(sha["obj"]),
                        sha_stl=np.array(sha["stl"]), stl_segs=segs)


if __name__ == "__main__":
    main()
