#!/usr/bin/env python
"""Golden output of the reference's UNMODIFIED example drivers (example/AlecTest.cpp, testVelocityFilter.cpp,
testNewSequence.cpp) built against its own CPU detection classes (oracle/_ref/<driver>_cpu), with the meshes they ran on.
tests/test_gpu_parity.py::test_reference_drivers_with_gpu_detection feeds the same meshes to the same drivers linked against
the GPU library (oracle/_ref/<driver>_gpu, examples/dropin_link.cpp) and compares what they print, line by line.
Run where /root/reference exists:  make -C oracle ref && python tests/golden/make_golden_drivers.py"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from collisiondetection_b200 import scenes  # noqa: E402

MESHES = "/root/reference/meshes"
REF = os.path.join(ROOT, "oracle", "_ref")
VF_SECONDS = 25      # testVelocityFilter on mesh1 -> mesh2 does not finish in minutes: a prefix is kept


def write_obj(path, q, f):
    with open(path, "w") as fh:
        for p in np.asarray(q).reshape(-1, 3):
            fh.write("v %.17g %.17g %.17g\n" % tuple(p))
        for t in np.asarray(f).reshape(-1, 3):
            fh.write("f %d %d %d\n" % (t[0] + 1, t[1] + 1, t[2] + 1))


def run(exe, args, cwd, seconds=None):
    try:
        p = subprocess.run([os.path.join(REF, exe)] + args, cwd=cwd, capture_output=True, text=True, timeout=seconds)
        return p.stdout
    except subprocess.TimeoutExpired as e:
        out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
        return out[:out.rfind("\n") + 1]      # whole lines only


def main():
    out = {}
    with tempfile.TemporaryDirectory() as d:
        # example/AlecTest.cpp on prob11 (a case where the two narrowphases differ: 2210 / 2212 edge-edge hits)
        q0, f = scenes.load_obj(os.path.join(MESHES, "V0_prob11_835.obj"))
        q1, _ = scenes.load_obj(os.path.join(MESHES, "V1_prob11_835.obj"))
        write_obj(os.path.join(d, "V0.obj"), q0, f)
        write_obj(os.path.join(d, "V1.obj"), q1, f)
        out["alec_q0"], out["alec_q1"], out["alec_f"] = q0, q1, f.astype(np.int32)
        out["alec_stdout"] = run("AlecTest_cpu", ["V0.obj", "V1.obj"], d)
        # example/testVelocityFilter.cpp: mesh1 -> mesh2, no infinite masses
        qa, fa = scenes.load_obj(os.path.join(MESHES, "mesh1.obj"))
        qb, _ = scenes.load_obj(os.path.join(MESHES, "mesh2.obj"))
        write_obj(os.path.join(d, "mesh1.obj"), qa, fa)
        write_obj(os.path.join(d, "mesh2.obj"), qb, fa)
        out["vf_q1"], out["vf_q2"], out["vf_f"] = qa, qb, fa.astype(np.int32)
        out["vf_stdout"] = run("testVelocityFilter_cpu", ["mesh1.obj", "mesh2.obj", "0"], d, VF_SECONDS)
        # example/testNewSequence.cpp: coarse mesh against the first frames of Model1_flow
        qc, fc = scenes.load_obj(os.path.join(MESHES, "Model1_flow", "Model1_coarse.obj"))
        write_obj(os.path.join(d, "coarse.obj"), qc, fc)
        out["seq_coarse_q"], out["seq_coarse_f"] = qc, fc.astype(np.int32)
        nframes = 4
        for k in range(nframes):
            qf, ff = scenes.load_obj(os.path.join(MESHES, "Model1_flow", "Model1_%d.obj" % k))
            write_obj(os.path.join(d, "fine_%d.obj" % k), qf, ff)
            out["seq_fine_q%d" % k] = qf
            out["seq_fine_f"] = ff.astype(np.int32)
        out["seq_nframes"] = nframes
        out["seq_stdout"] = run("testNewSequence_cpu", ["1e-3", "1e-4", "coarse.obj", "fine_"], d)
    for k in ("alec_stdout", "vf_stdout", "seq_stdout"):
        print("== %s: %d lines" % (k, out[k].count("\n")))
        print(out[k][:1500])
    np.savez_compressed(os.path.join(HERE, "drivers.npz"), **out)


if __name__ == "__main__":
    main()
