#!/usr/bin/env python
"""Small configs for compute-sanitizer (memcheck / racecheck / synccheck): C1 (prob3_402: both broadphases, both narrowphases,
multi-entry History) and the 41 x 41 cloth through the fused step, single and sharded.  Run as
    compute-sanitizer --tool memcheck python scripts/sanitize.py
Results are only checked for sanity here (parity is the tests' job)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from collisiondetection_b200 import api, scenes  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
ctx = api.Context(0)
g = np.load(os.path.join(G, "alec_prob3_402.npz"))
q0, q1, f = g["q0"], g["q1"], g["faces"]
H = api.single_step_history(q0, q1)
for kind in (api.KDOP, api.AABB):
    vf, ee = ctx.findCollisionCandidatesStep(kind, f, q0, q1, 1e-8)
    vf2, ee2 = ctx.findCollisionCandidates(kind, f, *H, 1e-8)
    assert np.array_equal(vf, vf2) and np.array_equal(ee, ee2)
out = ctx.findCollisions(*H, vf, 1e-8, ee, 1e-8)
sp = ctx.findCollisionsSeparatingPlane(*H, vf, 1e-8, ee, 1e-8)
rng = np.random.default_rng(1)
out2 = ctx.findCollisions(*H, vf, rng.uniform(1e-6, 1e-3, len(vf)), ee, rng.uniform(1e-6, 1e-3, len(ee)))
h = np.load(os.path.join(G, "history_prob3_402.npz"))
Hm = (h["hoff"], h["htime"], h["hpos"])
vfm, eem = ctx.findCollisionCandidates(api.KDOP, h["faces"], *Hm, float(h["outer_eta"]) if "outer_eta" in h.files else 1e-8)
outm = ctx.findCollisions(*Hm, vfm, 1e-8, eem, 1e-8)
q0, q1, f, eta = scenes.cloth(41)
r = ctx.step(api.KDOP, f, q0, q1, eta, eta)
tot = 0
for rank in range(3):
    rr = ctx.step(api.KDOP, f, q0, q1, eta, eta, None, rank, 3)
    tot += rr["n_vf_hits"] + rr["n_ee_hits"]
assert tot == r["n_vf_hits"] + r["n_ee_hits"], (tot, r["n_vf_hits"] + r["n_ee_hits"])
d = ctx.meshSelfDistance(q0, f)
# PenaltyGroup::addForce (penalty.cu), the staged multi-entry / SeparatingPlane paths ran above (outm, sp); their one-thread forms:
pg = np.load(os.path.join(G, "penalty.npz"))
dt_, outer_, inner_, k_, cor_ = [float(x) for x in pg["thick_params"]]
ctx.penaltyGroupAddForce(pg["thick_q"], pg["thick_v"], pg["thick_vf"], pg["thick_ee"], dt_, outer_, inner_, k_, cor_, pg["thick_F0"])
os.environ["CCD_HISTORY_ONE_THREAD"] = "1"
os.environ["CCD_SEPPLANE_ONE_THREAD"] = "1"
ctx.findCollisions(*Hm, vfm, 1e-8, eem, 1e-8)
ctx.findCollisionsSeparatingPlane(*Hm, vfm, 1e-8, eem, 1e-8)
del os.environ["CCD_HISTORY_ONE_THREAD"], os.environ["CCD_SEPPLANE_ONE_THREAD"]
spm = ctx.findCollisionsSeparatingPlane(*Hm, vfm, 1e-8, eem, 1e-8)
print("sanitize run ok: %d+%d candidates, %d+%d hits, cloth41 hits %d, self distance %r" % (
    len(vf), len(ee), out["n_vf_hits"], out["n_ee_hits"], tot, d))
ctx.close()
