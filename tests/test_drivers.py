"""The reference's unmodified example drivers built against its own CPU detection classes (oracle/_ref/<driver>_cpu) still
print what tests/golden/drivers.npz holds — keeps the golden of the GPU test test_reference_drivers_with_gpu_detection honest.
Runs only where oracle/_ref was built (a container with /root/reference)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden


def _write_obj(path, q, f):
    with open(path, "w") as fh:
        for p in np.asarray(q).reshape(-1, 3):
            fh.write("v %.17g %.17g %.17g\n" % tuple(p))
        for t in np.asarray(f).reshape(-1, 3):
            fh.write("f %d %d %d\n" % (t[0] + 1, t[1] + 1, t[2] + 1))


@pytest.mark.parametrize("driver", ["AlecTest", "testNewSequence"])
def test_cpu_drivers_reproduce_golden(tmp_path, driver):
    exe = os.path.join(ROOT, "oracle", "_ref", driver + "_cpu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/%s_cpu not built" % driver)
    g = golden("drivers.npz")
    d = str(tmp_path)
    if driver == "AlecTest":
        _write_obj(os.path.join(d, "V0.obj"), g["alec_q0"], g["alec_f"])
        _write_obj(os.path.join(d, "V1.obj"), g["alec_q1"], g["alec_f"])
        args, want = ["V0.obj", "V1.obj"], str(g["alec_stdout"])
    else:
        _write_obj(os.path.join(d, "coarse.obj"), g["seq_coarse_q"], g["seq_coarse_f"])
        for k in range(int(g["seq_nframes"])):
            _write_obj(os.path.join(d, "fine_%d.obj" % k), g["seq_fine_q%d" % k], g["seq_fine_f"])
        args, want = ["1e-3", "1e-4", "coarse.obj", "fine_"], str(g["seq_stdout"])
    out = subprocess.run([exe] + args, cwd=d, capture_output=True, text=True, timeout=300)
    assert out.stdout == want
