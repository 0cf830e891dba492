import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from conftest import golden
from collisiondetection_b200 import api
from oracle import bind
g = golden("alec_prob3_402_thick.npz")
rng = np.random.default_rng(3)
vf, ee = g["ref_vf"], g["ref_ee"]
ve = rng.uniform(1e-6, 2e-3, len(vf)); ee_eta = rng.uniform(1e-6, 2e-3, len(ee))
H = bind.single_step_history(g["q0"], g["q1"])
port = bind.Port(); ctx = api.Context(0)
p = port.narrowphase(*H, vf, ve, ee, ee_eta)
for rep in range(3):
    out = ctx.findCollisions(*H, vf, ve, ee, ee_eta)
    for k in ("vf", "ee"):
        bad = np.nonzero(out[k + "_toi"].view(np.uint64) != p[k + "_toi"].view(np.uint64))[0]
        print(rep, k, "toi mismatches", len(bad), bad[:5], [(out[k+"_toi"][b], p[k+"_toi"][b], int(out[k+"_stage"][b]), int(p[k+"_stage"][b]), int(out[k+"_hit"][b])) for b in bad[:3]])
