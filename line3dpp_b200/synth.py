"""Deterministic synthetic scenes for the BASELINE.json configs.

Input generation only (numpy); nothing here is on the measured path.

Scene: L = 2N random 3D segments (endpoints uniform in the cube [-1,1]^3, length U(0.05, 0.5)); V pinhole cameras
3072x2304, f = 2500 px, pp = (1536, 1152), on a circle of radius 4 in the xz-plane at height U(-0.3, 0.3), looking
at the origin.  Every view observes a random N-subset of the lines that project fully inside the image with a 2D
length >= 0.005*diag = 19.2 px (the reference's own detection filter, line3D.cc:320-360); endpoints get N(0, 0.5 px)
noise.  PRNG: numpy default_rng(seed); seed = 1000 + cfg number.
"""
from __future__ import annotations

import dataclasses
import numpy as np

WIDTH, HEIGHT, FOCAL = 3072, 2304, 2500.0
MIN_LEN_PX = 0.005 * float(np.sqrt(np.float32(WIDTH * WIDTH + HEIGHT * HEIGHT)))


@dataclasses.dataclass
class Scene:
    """What a frontend hands to Line3D::addImage (line3D.h:104-108) for every view."""
    cam_ids: np.ndarray          # (V,) uint32
    K: np.ndarray                # (V,3,3) float64
    R: np.ndarray                # (V,3,3) float64   point2D = K [R | t] point3D
    t: np.ndarray                # (V,3)   float64
    median_depth: np.ndarray     # (V,)    float32
    segs: list                   # V arrays (n_i,4) float32  (x1,y1,x2,y2) px
    neighbors: list              # V arrays of uint32 cam ids (explicit visual neighbours, line3D.h:76-78)
    width: int = WIDTH
    height: int = HEIGHT
    lines3d: np.ndarray | None = None   # (L,6) ground truth
    line_ids: list | None = None        # V arrays (n_i,) index of the 3D line each 2D segment observes

    @property
    def num_views(self) -> int:
        return len(self.cam_ids)


def look_at_camera(C: np.ndarray):
    """World->camera rotation with the optical axis through the origin; t = -R C."""
    z = -C / np.linalg.norm(C)
    up = np.array([0.0, 1.0, 0.0])
    x = np.cross(up, z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z], axis=0)
    return R, -R @ C


def ring_neighbors(V: int, half: int) -> list:
    """Ring +-half neighbours (mutual), cfg4/cfg5."""
    out = []
    for i in range(V):
        nb = [(i + d) % V for d in range(-half, half + 1) if d != 0]
        out.append(np.array(sorted(set(nb) - {i}), dtype=np.uint32))
    return out


def dense_neighbors(V: int) -> list:
    """Every view neighbours every other one (cfg3)."""
    return [np.array([j for j in range(V) if j != i], dtype=np.uint32) for i in range(V)]


def make_scene_views(num_views, segs_per_view, seed, neighbors, views, noise_px: float = 0.5) -> Scene:
    """Same scene as make_scene but only the 2D segments of `views` are materialised (the others are empty arrays):
    every rank of a multi-GPU run builds the identical global scene definition and only its own shard of segments."""
    return make_scene(num_views, segs_per_view, seed, neighbors, noise_px, only_views=set(int(v) for v in views))


def make_scene(num_views: int, segs_per_view: int, seed: int, neighbors: str | int = "ring5",
               noise_px: float = 0.5, only_views=None, collinear: bool = False) -> Scene:
    """collinear=True: every odd 3D line continues the even line before it on the same infinite line after a gap
    (broken edges, the case reconstruct3Dlines' collinearity_t is for); the default scenes are unchanged."""
    rng = np.random.default_rng(seed)
    V, N = num_views, segs_per_view
    L = 2 * N
    # 3D lines
    P1 = rng.uniform(-1.0, 1.0, size=(L, 3))
    d = rng.normal(size=(L, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    length = rng.uniform(0.05, 0.5, size=(L, 1))
    if collinear:
        gap = rng.uniform(0.03, 0.15, size=(L // 2, 1))
        P1 = P1 * 0.6
        d[1::2] = d[0::2]
        P1[1::2] = P1[0::2] + d[0::2] * (length[0::2] + gap)
        P2 = P1 + d * length
    else:
        P2 = np.clip(P1 + d * length, -1.0, 1.0)
    lines3d = np.concatenate([P1, P2], axis=1)

    Kmat = np.array([[FOCAL, 0, WIDTH / 2.0], [0, FOCAL, HEIGHT / 2.0], [0, 0, 1.0]])
    heights = rng.uniform(-0.3, 0.3, size=V)
    Ks = np.repeat(Kmat[None], V, axis=0)
    Rs = np.empty((V, 3, 3))
    ts = np.empty((V, 3))
    segs, line_ids = [], []
    for i in range(V):
        th = 2.0 * np.pi * i / V
        C = np.array([4.0 * np.cos(th), heights[i], 4.0 * np.sin(th)])
        R, t = look_at_camera(C)
        Rs[i], ts[i] = R, t
        if only_views is not None and i not in only_views:
            segs.append(np.zeros((0, 4), np.float32)); line_ids.append(np.zeros(0, np.int64))
            continue
        vrng = np.random.default_rng([seed, i])
        perm = vrng.permutation(L)
        X1 = (R @ P1[perm].T).T + t
        X2 = (R @ P2[perm].T).T + t
        u1 = X1[:, :2] / X1[:, 2:3] * FOCAL + Kmat[:2, 2]
        u2 = X2[:, :2] / X2[:, 2:3] * FOCAL + Kmat[:2, 2]
        u1 = u1 + vrng.normal(scale=noise_px, size=u1.shape)
        u2 = u2 + vrng.normal(scale=noise_px, size=u2.shape)
        ok = (X1[:, 2] > 0.1) & (X2[:, 2] > 0.1)
        for u in (u1, u2):
            ok &= (u[:, 0] >= 0) & (u[:, 0] <= WIDTH - 1) & (u[:, 1] >= 0) & (u[:, 1] <= HEIGHT - 1)
        ok &= np.linalg.norm(u1 - u2, axis=1) >= MIN_LEN_PX
        idx = np.flatnonzero(ok)[:N]
        s = np.concatenate([u1[idx], u2[idx]], axis=1).astype(np.float32)
        segs.append(np.ascontiguousarray(s))
        line_ids.append(perm[idx])
    if neighbors == "dense":
        nb = dense_neighbors(V)
    elif isinstance(neighbors, str) and neighbors.startswith("ring"):
        nb = ring_neighbors(V, int(neighbors[4:]))
    else:
        nb = ring_neighbors(V, int(neighbors))
    return Scene(cam_ids=np.arange(V, dtype=np.uint32), K=Ks, R=Rs, t=ts,
                 median_depth=np.full(V, 4.0, dtype=np.float32), segs=segs, neighbors=nb,
                 lines3d=lines3d, line_ids=line_ids)


# ---- camera quantities a view carries (View::View, view.cc:6-42) -- float64, no translation applied -------------
def camera_blocks(scene: Scene):
    """RtKinv (V,3,3), C (V,3) in double, exactly what View::View derives from K,R,t."""
    Kinv = np.linalg.inv(scene.K)
    Rt = np.transpose(scene.R, (0, 2, 1))
    RtKinv = Rt @ Kinv
    C = np.einsum("vij,vj->vi", Rt, -scene.t)
    return RtKinv, C


def fundamental(K1, R1, t1, K2, R2, t2):
    """Line3D::getFundamentalMatrix (line3D.cc:861-897), double."""
    R = R2 @ R1.T
    t = t2 - R @ t1
    T = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    E = T @ R
    return np.linalg.inv(K2.T) @ E @ np.linalg.inv(K1)


def view_pairs(neighbors: list, cam_ids=None) -> np.ndarray:
    """Deduplicated (src,tgt) view-pair list in the reference's match order (computeMatches, line3D.cc:704-741):
    views ascending, neighbours ascending, a pair is taken the first time either side lists the other."""
    V = len(neighbors)
    ids = np.arange(V) if cam_ids is None else np.asarray(cam_ids)
    order = np.argsort(ids, kind="stable")
    matched = set()
    pairs = []
    for vi in order:
        s = int(ids[vi])
        for n in sorted(int(x) for x in neighbors[vi]):
            key = (min(s, n), max(s, n))
            if n == s or key in matched:
                continue
            matched.add(key)
            pairs.append((s, n))
    return np.array(pairs, dtype=np.int32).reshape(-1, 2)
