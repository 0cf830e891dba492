"""Host-side scene helpers: the reference's OBJ dialect and the synthetic cloth of BASELINE config C5.

* ``load_obj``   reads the `v x y z` / `f a b c` (1-based) dialect every reference example parses with a
                 char-then-numbers loop (example/AlecTest.cpp:16-62).
* ``cloth``      the 3-layer S-folded cloth recipe of SURVEY.md §8(d): generated ONCE on the host (libm
                 sin/cos/exp through numpy) so CPU checker and GPU see identical doubles.
"""
import numpy as np


def load_obj(path):
    verts = []
    faces = []
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith("v "):
                p = line.split()
                verts.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith("f "):
                p = line.split()
                faces.append((int(p[1].split("/")[0]) - 1, int(p[2].split("/")[0]) - 1, int(p[3].split("/")[0]) - 1))
    return (np.asarray(verts, dtype=np.float64).reshape(-1, 3),
            np.asarray(faces, dtype=np.int32).reshape(-1, 3))


def splitmix64(x):
    """Vectorised splitmix64 on uint64 arrays."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def cloth(n, seed=20261017, cols=None):
    """Synthetic self-colliding cloth: n x n vertices (n=1415 -> 3,998,792 triangles).

    Returns (q0, q1, faces, eta) with q* of shape (V,3) float64, faces (F,3) int32 and
    eta = outerEta = 0.01*h.  `cols=(j0, j1)` keeps only grid columns j0 <= j < j1: a ribbon of the
    same cloth across the whole S-fold (vertex ids renumbered, coordinates and noise identical) —
    used for bounded CPU samples of the full-size workload.
    """
    h = 1.0 / (n - 1)
    gap = 4.0 * h
    r = 2.0 * h
    a = (1.0 - 2.0 * np.pi * r) / 3.0
    j0, j1 = (0, n) if cols is None else cols
    ii, jj = np.meshgrid(np.arange(n), np.arange(j0, j1), indexing="ij")
    u = ii / (n - 1.0)
    v = jj / (n - 1.0)
    s = u
    x = np.empty_like(s)
    z = np.empty_like(s)
    nx = np.empty_like(s)
    nz = np.empty_like(s)
    pr = np.pi * r
    m = s <= a
    x[m] = s[m]; z[m] = 0.0; nx[m] = 0.0; nz[m] = 1.0
    m = (s > a) & (s <= a + pr)
    th = (s[m] - a) / r
    x[m] = a + r * np.sin(th); z[m] = r - r * np.cos(th); nx[m] = -np.sin(th); nz[m] = np.cos(th)
    m = (s > a + pr) & (s <= 2 * a + pr)
    d = s[m] - (a + pr)
    x[m] = a - d; z[m] = gap; nx[m] = 0.0; nz[m] = -1.0
    m = (s > 2 * a + pr) & (s <= 2 * a + 2 * pr)
    th = (s[m] - (2 * a + pr)) / r
    x[m] = -r * np.sin(th); z[m] = gap + r - r * np.cos(th); nx[m] = np.sin(th); nz[m] = -np.cos(th)
    m = s > 2 * a + 2 * pr
    d = s[m] - (2 * a + 2 * pr)
    x[m] = d; z[m] = 2 * gap; nx[m] = 0.0; nz[m] = 1.0
    w = 0.5 * h * np.sin(40.0 * np.pi * u) * np.sin(36.0 * np.pi * v)
    x0 = x + w * nx
    z0 = z + w * nz
    y0 = v
    q0 = np.stack([x0, y0, z0], axis=-1).reshape(-1, 3)
    z1 = gap + (z0 - gap) * (1.0 - 1.3 * np.exp(-(((v - 0.5) / 0.15) ** 2)))
    k = (ii * n + jj).astype(np.uint64).reshape(-1)
    q1 = np.stack([x0, y0, z1], axis=-1).reshape(-1, 3).copy()
    for c in range(3):
        R = (splitmix64(np.uint64(seed) + np.uint64(3) * k + np.uint64(c)) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
        q1[:, c] += (2.0 * R - 1.0) * 0.1 * h
    # faces: all first triangles in row-major cell order, then all second triangles
    nc = j1 - j0
    ci, cj = np.meshgrid(np.arange(n - 1), np.arange(nc - 1), indexing="ij")
    c0 = (ci * nc + cj).reshape(-1)
    c1 = ((ci + 1) * nc + cj).reshape(-1)
    c2 = ((ci + 1) * nc + cj + 1).reshape(-1)
    c3 = (ci * nc + cj + 1).reshape(-1)
    odd = ((ci + (cj + j0)) % 2 == 1).reshape(-1)
    t1 = np.where(odd[:, None], np.stack([c0, c1, c3], -1), np.stack([c0, c1, c2], -1))
    t2 = np.where(odd[:, None], np.stack([c1, c2, c3], -1), np.stack([c0, c2, c3], -1))
    faces = np.concatenate([t1, t2], axis=0).astype(np.int32)
    return np.ascontiguousarray(q0), np.ascontiguousarray(q1), np.ascontiguousarray(faces), 0.01 * h
