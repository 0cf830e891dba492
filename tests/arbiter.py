"""TEST INFRASTRUCTURE — high-precision arbiter for narrowphase flag / TOI mismatches.

Given the reference's own double-precision polynomial coefficients (dumped bit-identically by the C
restatement, oracle/ccd_oracle.c: orc_vf_polys / orc_ee_polys) it evaluates CTCD::vertexFaceCTCD /
edgeEdgeCTCD (src/CTCD.cpp:259-508) with the roots of those polynomials computed to 60 digits and the
midpoint signs evaluated exactly.  It answers "what is the mathematically right result for the
polynomial the reference built", which is what separates

  * reference artefacts — Jenkins-Traub (src/rpoly.h) returning roots that are off by up to ~1e-4 on
    nearly-double root pairs, or not converging at all — from
  * genuine disagreements of the new root isolator.

Also reports how close the decisive sign evaluations are to rounding noise (SURVEY.md §8c class iv).
"""
import ctypes as C
from fractions import Fraction

import mpmath as mp
import numpy as np

mp.mp.dps = 60


def _normalise(op):
    """src/CTCD.cpp:113-133 in double precision (bit-identical to the reference)."""
    op = np.array(op, dtype=np.float64)
    m = np.abs(op).max()
    if m != 0:
        op = op / m
    k = 0
    while k < len(op) - 1 and op[k] == 0:
        k += 1
    return op[k:]


def _could_have_roots(op, pos):
    """src/CTCD.cpp:81-94 (double precision, as the reference)."""
    res = 0.0
    for c in op[:-1]:
        if (pos and c > 0) or (not pos and c < 0):
            res += c
    res += op[-1]
    return not ((pos and res < 0) or (not pos and res > 0))


def _exact_sign(op, t):
    """sign of the polynomial with double coefficients op at the (mp) point t, evaluated in 60 digits"""
    f = mp.mpf(0)
    for c in op:
        f = f * t + mp.mpf(float(c))
    return f


def find_intervals_exact(op, pos):
    """Intervals of [0,1] where the reference's polynomial is >=0 (pos) / <=0, using its exact real roots.
    Returns (list of (l,u) as mp numbers, min |f(mid)| seen relative to sum|c| — the noise margin)."""
    op = _normalise(op)
    deg = len(op) - 1
    margin = mp.inf
    if deg > 2 and not _could_have_roots(op, pos):
        return [], margin
    if deg == 0:
        ok = (op[0] >= 0) if pos else (op[0] <= 0)
        return ([(mp.mpf(0), mp.mpf(1))] if ok else []), margin
    roots = mp.polyroots([mp.mpf(float(c)) for c in op], maxsteps=500, extraprec=400) if deg >= 1 else []
    real = sorted(mp.re(r) for r in roots if abs(mp.im(r)) <= mp.mpf(10) ** -25)
    scale = sum(abs(float(c)) for c in op)
    ivs = []

    def check(a, b):
        nonlocal margin
        a = min(max(a, 0), 1)
        b = min(max(b, 0), 1)
        mid = (a + b) / 2
        f = _exact_sign(op, mid)
        margin = min(margin, abs(f) / scale)
        if (pos and f >= 0) or (not pos and f <= 0):
            ivs.append((min(a, b), max(a, b)))

    if real:
        if real[0] >= 0:
            check(mp.mpf(0), real[0])
        for i in range(len(real) - 1):
            if not ((real[i] < 0 and real[i + 1] < 0) or (real[i] > 1 and real[i + 1] > 1)):
                check(real[i], real[i + 1])
        if real[-1] <= 1:
            check(real[-1], mp.mpf(1))
    else:
        check(mp.mpf(0), mp.mpf(1))
    return ivs, margin


def _overlap(a, b):
    return not (a[0] > b[1] or b[0] > a[1])


def _combine(lists, parallel=()):
    import itertools
    col, mint = False, mp.mpf(1)
    for combo in itertools.product(*lists):
        if all(_overlap(x, y) for i, x in enumerate(combo) for y in combo[i + 1:]):
            il = max([c[0] for c in combo] + [mp.mpf(0)])
            iu = min([c[1] for c in combo] + [mp.mpf(1)])
            if any(_overlap((il, iu), p) for p in parallel):
                continue
            mint = min(mint, il)
            col = True
    return col, mint


class Arbiter(object):
    def __init__(self, port):
        self.lib = port.lib
        self._dp = C.POINTER(C.c_double)
        self.jitter = None   # (rng, relative size): perturb the coefficients at rounding level
        self.push = 0        # +1 / -1: loosest / strictest reading within Horner's rounding-error bound

    def _polys(self, fn, pts, eta, n1, n2):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        if self.jitter is not None:
            # move every input coordinate by about one unit in the last place
            pts = pts * (1.0 + 2.2e-16 * self.jitter[0].choice([-1.0, 0.0, 1.0], pts.shape))
        a = np.zeros(n1)
        b = np.zeros(n2)
        getattr(self.lib, fn)(pts.ctypes.data_as(self._dp), C.c_double(eta), a.ctypes.data_as(self._dp), b.ctypes.data_as(self._dp))
        if self.jitter is not None:
            rng, rel = self.jitter
            a = a * (1.0 + rel * rng.standard_normal(a.shape))
            b = b * (1.0 + rel * rng.standard_normal(b.shape))
        if self.push:
            # Horner's rounding error at t in [0,1] is bounded by ~2*deg*eps*sum|c_i| t^i, itself a polynomial:
            # pushing every coefficient up (down) by that relative amount gives the largest (smallest) value any
            # double-precision evaluation could have reported.  wants_pos tells which direction is "loose".
            def shove(c, wants_pos):
                d = self.push * 16 * 1.1e-16 * np.abs(c)
                return c + d if wants_pos else c - d
            if fn == "orc_vf_polys":
                a, b = shove(a, True), shove(b, False)       # cubics want >= 0, sextic wants <= 0
            else:
                a, b = shove(a, False), shove(b, True)       # sextic wants <= 0, quartics want >= 0
        return a, b

    def noise_sensitive(self, kind, pts, eta, trials=16, rel=4.4e-16, seed=1):
        """True when the hit/miss answer flips when the input coordinates move by one ulp and the
        coefficients by a couple of ulps — the
        "noise-sign" class of SURVEY.md §8(c)(iv): no double-precision evaluation order can be called right
        (the reference itself flips 143 prob17 flags when merely rebuilt with FMA contraction)."""
        fn = self.vertex_face if kind == "vf" else self.edge_edge
        base = fn(pts, eta)[0]
        try:
            self.push = 1
            loose = fn(pts, eta)[0]
            self.push = -1
            strict = fn(pts, eta)[0]
        finally:
            self.push = 0
        if loose != strict:
            return True
        rng = np.random.default_rng(seed)
        try:
            for _ in range(trials):
                self.jitter = (rng, rel)
                if fn(pts, eta)[0] != base:
                    return True
        finally:
            self.jitter = None
        return False

    def classify(self, kind, pts, eta, mine, theirs):
        """Verdict on one flag mismatch (mine != theirs): 'reference-artefact' when the 60-digit answer is
        mine, 'noise-sign' when the decision is not stable at rounding level, else 'unexplained'."""
        fn = self.vertex_face if kind == "vf" else self.edge_edge
        if fn(pts, eta)[0] == bool(mine):
            return "reference-artefact"
        if self.noise_sensitive(kind, pts, eta):
            return "noise-sign"
        if self.degenerate_polynomial(kind, pts, eta):
            return "degenerate-poly"
        return "unexplained"

    def degenerate_polynomial(self, kind, pts, eta):
        """True when one of the stencil's polynomials has a multiple-root cluster inside Horner's rounding-error
        band on [0,1]: its interval STRUCTURE (number of sign changes) differs between the loosest and the
        strictest admissible reading, so any double-precision root finder's output there is arbitrary (typically
        parallel edges / a vertex sliding in the face plane — cases the reference itself warns about,
        include/CTCD.h:33-34)."""
        name, n1, n2 = ("orc_vf_polys", 12, 7) if kind == "vf" else ("orc_ee_polys", 7, 20)
        counts = []
        try:
            for push in (1, -1):
                self.push = push
                a, b = self._polys(name, pts, eta, n1, n2)
                if kind == "vf":
                    polys = [(a[0:4], True), (a[4:8], True), (a[8:12], True), (b, False)]
                else:
                    polys = [(a, False)] + [(b[5 * k:5 * k + 5], True) for k in range(4)]
                counts.append([len(find_intervals_exact(op, pos)[0]) for op, pos in polys])
        finally:
            self.push = 0
        return counts[0] != counts[1]

    def vertex_face(self, pts, eta):
        """pts: 24 doubles (q0..q3 start, q0..q3 end).  Returns (hit, toi, margin)."""
        cubics, sextic = self._polys("orc_vf_polys", pts, eta, 12, 7)
        margin = mp.inf
        lists = []
        for k in range(3):
            iv, m = find_intervals_exact(cubics[4 * k:4 * k + 4], True)
            margin = min(margin, m)
            if not iv:
                return False, None, margin
            lists.append(iv)
        iv, m = find_intervals_exact(sextic, False)
        margin = min(margin, m)
        if not iv:
            return False, None, margin
        col, t = _combine([iv] + lists)
        return col, (float(t) if col else None), margin

    def edge_edge(self, pts, eta):
        """pts: 24 doubles ((q0,p0,q1,p1) start then end)."""
        sextic, quartics = self._polys("orc_ee_polys", pts, eta, 7, 20)
        p = np.asarray(pts, dtype=np.float64).reshape(8, 3)
        q0s, p0s, q1s, p1s = p[0], p[1], p[2], p[3]
        vq0, vp0, vq1, vp1 = p[4] - p[0], p[5] - p[1], p[6] - p[2], p[7] - p[3]
        raw, margin = find_intervals_exact(sextic, False)
        cop, par = [], []
        for (l, u) in raw:
            mid = float((l + u) / 2)
            x10 = (q0s - p0s) + mid * (vq0 - vp0)
            x20 = (q1s - p1s) + mid * (vq1 - vp1)
            (par if np.linalg.norm(np.cross(x10, x20)) < 1e-8 else cop).append((l, u))
        if not cop:
            return False, None, margin
        lists = []
        for k in range(4):
            iv, m = find_intervals_exact(quartics[5 * k:5 * k + 5], True)
            margin = min(margin, m)
            if not iv:
                return False, None, margin
            lists.append(iv)
        col, t = _combine([cop] + lists, par)
        return col, (float(t) if col else None), margin
