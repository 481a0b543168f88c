"""numpy restatement of the level-1 pre-filter of the matching kernel (line3dpp_b200/csrc/l3d_device.cuh "pencil parameter",
l3d_match.cu: pair_basis / target_arc / line_kappa / arc_may_match / the window of k_match_topk).  TEST INFRASTRUCTURE: it lets the
CPU suite check the derivation - no match of the exhaustive oracle may fall outside a row's window or fail the arc test - on
geometries the GPU tests then run through the real kernels."""
import numpy as np

S = 2048.0
TU = 4294967296.0 / np.pi
M32 = 0xFFFFFFFF


def units(k):
    return int(np.rint(k * TU)) & M32


def basis(F):
    F = np.asarray(F, np.float64).reshape(3, 3)
    c = []
    for k in range(3):
        n = np.linalg.norm(F[:, k])
        c.append(F[:, k] / n if n > 0 else np.zeros(3))
    best, E = 0.0, None
    for a in range(3):
        for b in range(a + 1, 3):
            x = np.cross(c[a], c[b])
            if x @ x > best:
                best, E = x @ x, x
    n = np.array([E[0] / S, E[1] / S, E[2]])
    n /= np.linalg.norm(n)
    ax = int(np.argmin(np.abs(n)))
    a = np.zeros(3); a[ax] = 1.0
    u = np.cross(n, a); u /= np.linalg.norm(u)
    return u, np.cross(n, u), E


def target_arc(u, v, q, ext):
    """(A, w, e1, ehi) in units of pi / 2^32, or None when the target has no usable arc (always a candidate)"""
    x1, y1, x2, y2 = [float(t) / S for t in q]
    a1, b1 = x1 * u[0] + y1 * u[1] + u[2], x1 * v[0] + y1 * v[1] + v[2]
    dx, dy = x2 - x1, y2 - y1
    ac, bc = dx * u[0] + dy * u[1], dx * v[0] + dy * v[1]
    ln, rhoc = np.hypot(dx, dy), np.hypot(ac, bc)
    delta = (0.05 + 4e-3 * ln * S) / S
    if not ln > 1e-9 or not rhoc > 1e-3 * ln:
        return None
    l, n = np.array([y1 - y2, x2 - x1, x1 * y2 - y1 * x2]), np.cross(u, v)
    if not abs(l @ n) > 1e-4 * np.linalg.norm(l):
        return None
    kc = units(np.arctan2(ac, -bc))
    al, mg = [], []
    for t in (-ext, 0.0, 1.0, 1.0 + ext):
        a, b = a1 + t * ac, b1 + t * bc
        rho = np.hypot(a, b)
        if not delta < 0.25 * rho:
            return None
        al.append(float((units(np.arctan2(a, -b)) - kc) & M32))
        mg.append((1.2 * delta / rho + 4e-6) * TU)
    up, down = al[0] <= al[1] <= al[2] <= al[3], al[0] >= al[1] >= al[2] >= al[3]
    if not up and not down:
        return None
    i0, i1, i2, i3 = (0, 1, 2, 3) if up else (3, 2, 1, 0)
    top = 4294967295.0
    lo, hi = max(al[i1] - mg[i1], 0.0), min(al[i2] + mg[i2], top)
    lox, hix = min(max(al[i0] - mg[i0], 0.0), lo), max(min(al[i3] + mg[i3], top), hi)
    lo, lox, hi, hix = np.floor(lo), np.floor(lox), np.ceil(hi), np.ceil(hix)
    if not hix - lox < 2147483648.0 - 524288.0:
        return None
    return (kc + int(lo)) & M32, int(hi - lo), int(lo - lox), int(hix - hi)


def epipolar_line(F, x, y):
    """mulmat_h: ((0 + m0*x) + m1*y) + m2 in float"""
    F = np.asarray(F, np.float32).reshape(3, 3)
    o = np.zeros(3, np.float32)
    for i in range(3):
        acc = np.float32(0.0)
        acc = np.float32(acc + np.float32(F[i, 0] * np.float32(x)))
        acc = np.float32(acc + np.float32(F[i, 1] * np.float32(y)))
        o[i] = np.float32(acc + np.float32(F[i, 2] * np.float32(1.0)))
    return o


def line_kappa(u, v, e):
    """(kappa, off): off = the line misses the pencil by more than the row guard allows"""
    x, y, z = S * float(e[0]), S * float(e[1]), float(e[2])
    n = np.cross(u, v)
    cn, exy = x * n[0] + y * n[1] + z * n[2], np.hypot(float(e[0]), float(e[1]))
    k = np.arctan2(x * v[0] + y * v[1] + z * v[2], x * u[0] + y * u[1] + z * u[2])
    return units(k), not 4.0 * abs(cn) <= 0.02 * exy


def r16(x):
    return ((x + 65535) >> 16) << 16


def arc_may_match(arc, k1, k2):
    A, w, e1, eh = arc
    E1 = r16(e1); E2 = E1 + r16(w); E3 = E2 + r16(eh)
    base = (A - E1) & M32
    d1, d2 = (k1 - base) & M32, (k2 - base) & M32
    return max(d1, d2) <= E3 and max(d1, d2) >= E1 and min(d1, d2) <= E2


NCLS, CLS0 = 8, 21


def arc_class(w):
    """class of a target by the length of its arc (l3d_device.cuh arc_class); NCLS = every row looks at it"""
    b = 0 if w <= 1 else (w - 1).bit_length()
    return 0 if b <= CLS0 else min(b - CLS0, NCLS)


def in_window(A, k1, k2, wmax):
    fwd = ((k2 - k1) & M32) < 0x80000000
    ka, kb = (k1, k2) if fwd else (k2, k1)
    ws = (ka - wmax) & M32
    return ws <= A <= kb if ws <= kb else (A >= ws or A <= kb)
