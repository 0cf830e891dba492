#!/usr/bin/env python
"""Golden hit flags of the UNMODIFIED reference's SeparatingPlaneNarrowPhase (oracle/_ref, which=1) on the scenes that
already have fixtures here.  Run where /root/reference exists (oracle/_ref built): python tests/golden/make_golden_sepplane.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import bind  # noqa: E402

SCENES = ["alec_prob3_402", "alec_prob11_835", "alec_prob18_834", "history_prob3_402", "alec_prob3_402_thick", "alec_prob3_402_fixed"]


def main():
    ref = bind.Ref()
    out = {}
    for name in SCENES:
        g = np.load(os.path.join(HERE, name + ".npz"))
        H = (g["hoff"], g["htime"], g["hpos"]) if "hoff" in g.files else bind.single_step_history(g["q0"], g["q1"])
        eta = float(g["eta"])
        r = ref.narrowphase(*H, g["ref_vf"], eta, g["ref_ee"], eta, which=1)
        out[name + "_vf"] = r["vf_hit"].astype(np.uint8)
        out[name + "_ee"] = r["ee_hit"].astype(np.uint8)
        print(name, int(r["vf_hit"].sum()), int(r["ee_hit"].sum()))
    np.savez_compressed(os.path.join(HERE, "sepplane.npz"), **out)


if __name__ == "__main__":
    main()
