#!/usr/bin/env python
"""Wall time of the reference's UNMODIFIED example drivers, all-CPU build against GPU-detection build (oracle/_ref/<driver>_cpu /
_gpu, see oracle/Makefile and examples/dropin_link.cpp): the drop-in boundary end to end, std::set marshalling, context
creation and process start included.  One JSON line per driver.  Needs the binaries (built where /root/reference exists)."""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
G = os.path.join(ROOT, "tests", "golden")


def write_obj(path, q, f):
    with open(path, "w") as fh:
        for p in np.asarray(q).reshape(-1, 3):
            fh.write("v %.17g %.17g %.17g\n" % tuple(p))
        for t in np.asarray(f).reshape(-1, 3):
            fh.write("f %d %d %d\n" % (t[0] + 1, t[1] + 1, t[2] + 1))


def timed(exe, args, cwd, repeat=1):
    best, out = None, ""
    for _ in range(repeat):
        t0 = time.perf_counter()
        p = subprocess.run([os.path.join(REF, exe)] + args, cwd=cwd, capture_output=True, text=True, timeout=1800)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        out = p.stdout
    return best, out


def main():
    with tempfile.TemporaryDirectory() as d:
        for name in ("alec_prob3_402", "alec_prob17_30957"):
            g = np.load(os.path.join(G, name + ".npz"))
            write_obj(os.path.join(d, "V0.obj"), g["q0"], g["faces"])
            write_obj(os.path.join(d, "V1.obj"), g["q1"], g["faces"])
            tc, oc = timed("AlecTest_cpu", ["V0.obj", "V1.obj"], d)
            tg, og = timed("AlecTest_gpu", ["V0.obj", "V1.obj"], d, 2)
            print(json.dumps(dict(driver="example/AlecTest.cpp (unmodified)", workload=name, triangles=int(len(g["faces"])), cpu_build_s=tc, gpu_build_s=tg,
                                  speedup=tc / tg, cpu_lines=oc.splitlines()[2:], gpu_lines=og.splitlines()[2:])))
        g = np.load(os.path.join(G, "drivers.npz"))
        write_obj(os.path.join(d, "coarse.obj"), g["seq_coarse_q"], g["seq_coarse_f"])
        for k in range(int(g["seq_nframes"])):
            write_obj(os.path.join(d, "fine_%d.obj" % k), g["seq_fine_q%d" % k], g["seq_fine_f"])
        a = ["1e-3", "1e-4", "coarse.obj", "fine_"]
        tc, oc = timed("testNewSequence_cpu", a, d)
        tg, og = timed("testNewSequence_gpu", a, d, 2)
        print(json.dumps(dict(driver="example/testNewSequence.cpp (unmodified)", workload="Model1_flow frames 0..3", cpu_build_s=tc, gpu_build_s=tg,
                              speedup=tc / tg, same_log=oc == og,
                              note="518 triangles: the detection passes are microseconds of work; process start, context creation and the CPU penalty-layer simulation dominate")))


if __name__ == "__main__":
    main()
