#!/usr/bin/env python
"""Golden vectors for PenaltyGroup::addForce, computed by the UNMODIFIED reference (oracle/_ref/libccdvf.so:
src/PenaltyGroup.cpp + src/PenaltyPotential.cpp + include/Distance.h through oracle/ref_recorder.cpp's
ref_penalty_group_force).  Runs in the build container only (needs /root/reference); writes tests/golden/penalty.npz.

Scenes: the candidate stencils of golden scenes (the lists a PenaltyGroup would hold, in set order) with
q = a point of the step where the meshes are close, v = the step's velocity, and thickness parameters in the range
ActiveLayers::addGroups produces (src/ActiveLayers.cpp:67-85)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import bind  # noqa: E402


def main():
    ref = bind.RefVF()
    rng = np.random.default_rng(20261017)
    out = {}
    cases = []
    for name, scene, tq, outer, inner in (("thick", "alec_prob3_402_thick.npz", 1.0, 4e-3, 1e-4),
                                          ("thick_mid", "alec_prob3_402_thick.npz", 0.5, 2e-2, 2e-3),
                                          ("prob11", "alec_prob11_835.npz", 1.0, 3e-3, 1e-5)):
        g = np.load(os.path.join(HERE, scene))
        q0, q1 = g["q0"].reshape(-1), g["q1"].reshape(-1)
        q = (1 - tq) * q0 + tq * q1
        v = q1 - q0
        cases.append((name, q, v, g["ref_vf"], g["ref_ee"], outer, inner))
    for name, q, v, vf, ee, outer, inner in cases:
        F0 = rng.standard_normal(q.size) * 1e-3
        F0[::7] = 0.0
        F0[3::11] = -0.0
        for tag, rolled in (("", False), ("_rb", True)):
            F, fired, newused = ref.penalty_group_force(q, v, vf, ee, 1e-3, outer, inner, 1e3, 0.5, F0, rolled_back=rolled)
            print(name + tag, "stencils", len(vf), len(ee), "fired", int(fired[:len(vf)].sum()), int(fired[len(vf):].sum()), "newused", newused,
                  "touched coords", int((F != F0).sum()))
            out[name + tag + "_F"] = F
            out[name + tag + "_newused"] = np.array(newused)
        out[name + "_q"], out[name + "_v"], out[name + "_vf"], out[name + "_ee"] = q, v, vf, ee
        out[name + "_F0"], out[name + "_fired"] = F0, fired
        out[name + "_params"] = np.array([1e-3, outer, inner, 1e3, 0.5])
    out["cases"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(HERE, "penalty.npz"), **out)


if __name__ == "__main__":
    main()
