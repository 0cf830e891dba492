#!/usr/bin/env python
"""Golden vectors for BASELINE config C4: the detection calls ActiveLayers makes INSIDE the unmodified
VelocityFilter::velocityFilter (src/VelocityFilter.cpp, src/ActiveLayers.cpp:189-215) — multi-entry Histories that grow
from pass to pass, every broadphase candidate with its per-stencil thickness, and the hits of the reference's
SeparatingPlaneNarrowPhase.  Recorded by oracle/ref_recorder.{h,cpp} (a subclass of the reference's narrowphase that calls
the reference's own code); the candidate sets are re-derived with the reference's KDOPBroadPhase on the recorded History.
Run where /root/reference exists:  make -C oracle ref && python tests/golden/make_golden_velocityfilter.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from collisiondetection_b200 import scenes  # noqa: E402
from oracle import bind  # noqa: E402

MESHES = "/root/reference/meshes"


def record(lib, q1, q2, faces, invmass, outer, inner, npasses):
    q1 = np.ascontiguousarray(q1, np.float64).reshape(-1)
    q2 = np.ascontiguousarray(q2, np.float64).reshape(-1)
    faces = np.ascontiguousarray(faces, np.int32).reshape(-1)
    invmass = np.ascontiguousarray(invmass, np.float64)
    V, F = len(q1) // 3, len(faces) // 3
    lib.vfrec_run.restype = C.c_int
    n = lib.vfrec_run(C.c_int(V), C.c_int(F), q1.ctypes.data_as(C.c_void_p), q2.ctypes.data_as(C.c_void_p), faces.ctypes.data_as(C.c_void_p),
                      invmass.ctypes.data_as(C.c_void_p), C.c_double(outer), C.c_double(inner), C.c_int(npasses))
    passes = []
    for p in range(n):
        sz = np.zeros(8, np.int64)
        lib.vfrec_sizes(C.c_int(p), sz.ctypes.data_as(C.c_void_p))
        N, nvf, nee, Vr = (int(x) for x in sz[:4])
        assert Vr == V
        d = dict(hoff=np.zeros(V + 1, np.int64), htime=np.zeros(N), hpos=np.zeros(3 * N), vf=np.zeros((nvf, 4), np.int32), ee=np.zeros((nee, 4), np.int32),
                 vf_eta=np.zeros(nvf), ee_eta=np.zeros(nee), vf_hit=np.zeros(nvf, np.uint8), ee_hit=np.zeros(nee, np.uint8))
        lib.vfrec_get(C.c_int(p), *[d[k].ctypes.data_as(C.c_void_p) for k in ("hoff", "htime", "hpos", "vf", "ee", "vf_eta", "ee_eta", "vf_hit", "ee_hit")])
        passes.append(d)
    return passes


def main():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libccdvf.so"))
    ref = bind.Ref()
    out = {}
    # (a) example/testVelocityFilter.cpp: mesh1 -> mesh2, no infinite masses, radii 2e-8 / 1e-8
    qa, fa = scenes.load_obj(os.path.join(MESHES, "mesh1.obj"))
    qb, _ = scenes.load_obj(os.path.join(MESHES, "mesh2.obj"))
    cases = [("mesh12", qa, qb, fa, np.ones(len(qa)), 2e-8, 1e-8, 5)]
    for name, q1, q2, faces, inv, outer, inner, npasses in cases:
        passes = record(lib, q1, q2, faces, inv, outer, inner, npasses)
        fixed = (inv == 0).astype(np.uint8)
        out[name + "_faces"] = faces.astype(np.int32)
        out[name + "_fixed"] = fixed
        out[name + "_outer"] = outer
        out[name + "_npasses"] = len(passes)
        for p, d in enumerate(passes):
            # the candidates the narrowphase was given ARE the broadphase output (src/ActiveLayers.cpp:193-212): check with the reference's own class
            vf, ee, _ = ref.broadphase(13, faces, d["hoff"], d["htime"], d["hpos"], outer, fixed if fixed.any() else None)
            assert np.array_equal(vf, d["vf"]) and np.array_equal(ee, d["ee"])
            for k, v in d.items():
                out["%s_p%d_%s" % (name, p, k)] = v
            print(name, "pass", p, "history entries", len(d["htime"]), "cand", len(d["vf"]), len(d["ee"]), "hits", int(d["vf_hit"].sum()), int(d["ee_hit"].sum()),
                  "eta range", d["vf_eta"].min() if len(d["vf_eta"]) else None, d["vf_eta"].max() if len(d["vf_eta"]) else None)
    np.savez_compressed(os.path.join(HERE, "velocityfilter.npz"), **out)


if __name__ == "__main__":
    main()
