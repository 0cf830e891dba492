#!/usr/bin/env python
"""Golden vectors for BASELINE config C3 (meshes/Model1_flow as example/testNewSequence.cpp:150-203 assembles it): the
coarse mesh (dynamic) joined with frame k of the fine sequence (fixed vertices, invmass 0) -> frame k+1, outer radius
1e-3, inner radius 1e-4, KDOP broadphase with the fixed-vertex filter, then both narrowphases of the UNMODIFIED
reference (oracle/_ref).  Run where /root/reference exists: python tests/golden/make_golden_model1.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from collisiondetection_b200 import scenes  # noqa: E402
from oracle import bind  # noqa: E402

DIR = "/root/reference/meshes/Model1_flow"
FRAMES = [0, 1, 60, 135]


def scene(k):
    qc, fc = scenes.load_obj(os.path.join(DIR, "Model1_coarse.obj"))
    qa, fa = scenes.load_obj(os.path.join(DIR, "Model1_%d.obj" % k))
    qb, fb = scenes.load_obj(os.path.join(DIR, "Model1_%d.obj" % (k + 1)))
    assert qa.shape == qb.shape and np.array_equal(fa, fb)
    nc = len(qc)
    q0 = np.concatenate([qc, qa])
    q1 = np.concatenate([qc, qb])
    faces = np.concatenate([fc, fa + nc]).astype(np.int32)
    fixed = np.zeros(len(q0), np.uint8)
    fixed[nc:] = 1      # invmasses == 0 -> fixedVerts (src/VelocityFilter.cpp)
    return q0, q1, faces, fixed


def main():
    ref = bind.Ref()
    out = {}
    for k in FRAMES:
        q0, q1, faces, fixed = scene(k)
        H = bind.single_step_history(q0, q1)
        vf, ee, _ = ref.broadphase(13, faces, *H, 1e-3, fixed)
        a = ref.narrowphase(*H, vf, 1e-4, ee, 1e-4)
        b = ref.narrowphase(*H, vf, 1e-4, ee, 1e-4, which=1)
        p = "f%d_" % k
        out.update({p + "q0": q0, p + "q1": q1, p + "faces": faces, p + "fixed": fixed, p + "vf": vf, p + "ee": ee,
                    p + "vf_hit": a["vf_hit"], p + "ee_hit": a["ee_hit"], p + "vf_toi": a["vf_toi"], p + "ee_toi": a["ee_toi"],
                    p + "vf_stage": a["vf_stage"].astype(np.uint8), p + "ee_stage": a["ee_stage"].astype(np.uint8),
                    p + "sp_vf_hit": b["vf_hit"], p + "sp_ee_hit": b["ee_hit"]})
        print("frame", k, "V", len(q0), "F", len(faces), "cand", len(vf), len(ee), "ctcd hits", int(a["vf_hit"].sum()), int(a["ee_hit"].sum()),
              "sepplane hits", int(b["vf_hit"].sum()), int(b["ee_hit"].sum()))
    out["frames"] = np.array(FRAMES)
    out["outer_eta"] = 1e-3
    out["eta"] = 1e-4
    np.savez_compressed(os.path.join(HERE, "model1_flow.npz"), **out)


if __name__ == "__main__":
    main()
