"""On-GPU probe of l3d_rdd on a banded random symmetric graph (same generator as bench.py's roofline leg)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from line3dpp_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
deg = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rng = np.random.default_rng(7)
a = np.repeat(np.arange(n, dtype=np.int64), deg)
b = np.clip(a + rng.integers(-3000, 3001, n * deg), 0, n - 1)
keep = a != b
key = np.unique(np.minimum(a[keep], b[keep]) * n + np.maximum(a[keep], b[keep]))
a, b = (key // n).astype(np.int32), (key % n).astype(np.int32)
w = rng.uniform(0.5, 1.0, len(a)).astype(np.float32)
ei, ej, ew = np.concatenate([a, b]), np.concatenate([b, a]), np.concatenate([w, w])
ctx = capi.Context(0)
for rep in range(3):
    _, _, _, ms = ctx.rdd(n, ei, ej, ew, 10)
    nnz = len(ei)
    print(f"n={n} nnz={nnz} avg deg {nnz/n:.1f}: {ms/10:.3f} ms/iter -> {(20.0*nnz+8.0*n)/(ms/10*1e-3)/1e9:.1f} GB/s (20 B/nnz)")
