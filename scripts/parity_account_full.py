#!/usr/bin/env python
"""The full parity account of the CPU restatement (= the GPU, bit for bit) against the unmodified reference's golden vectors,
written to profiles/r02_parity_account.json: per config and stencil type the classified flag differences and EVERY hit whose
time of impact is more than 1e-9 (relative) from the reference's, arbitrated with 60-digit arithmetic (tests/parity_account.py,
tests/arbiter.py).  prob17 alone takes ~13 minutes.  CPU only."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import bind  # noqa: E402
import parity_account as PA  # noqa: E402
from arbiter import Arbiter  # noqa: E402
import test_parity_account as T  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def main():
    port = bind.Port()
    arb = Arbiter(port)
    out = {}
    for name in ("alec_prob3_402", "alec_prob18_834", "alec_prob11_835", "alec_prob3_402_thick"):
        g = np.load(os.path.join(G, name + ".npz"))
        eta = float(g["eta"])
        H = bind.single_step_history(g["q0"], g["q1"])
        o = port.narrowphase(*H, g["ref_vf"], eta, g["ref_ee"], eta)
        out[name] = {}
        for k in ("vf", "ee"):
            out[name][k] = PA.account(arb, port, k, g["q0"], g["q1"], g["ref_" + k], eta, o[k + "_hit"], o[k + "_toi"], o[k + "_stage"],
                                      g["ref_%s_hit" % k], g["ref_%s_toi" % k], g["ref_%s_stage" % k])
        print(name, json.dumps(out[name]), flush=True)
    t0 = time.time()
    out["alec_prob17_30957"] = T._prob17(port, max_toi=None)
    out["alec_prob17_30957"]["seconds"] = time.time() - t0
    print("prob17", json.dumps(out["alec_prob17_30957"]), flush=True)
    g = np.load(os.path.join(G, "cloth_1415.npz"))
    out["cloth_1415"] = dict(n_vf=int(g["n_vf"]), n_ee=int(g["n_ee"]), ref_vf_hits=int(g["ref_vf_n_hits"]), ref_ee_hits=int(g["ref_ee_n_hits"]),
                             vf_flag_classes=sorted(map(str, g["vf_mismatch_class"])), ee_flag_classes=sorted(map(str, g["ee_mismatch_class"])),
                             vf_toi_out_of_1e9=int(g["vf_toi_out_of_1e9"]), ee_toi_out_of_1e9=int(g["ee_toi_out_of_1e9"]),
                             toi_sample_classes=sorted(map(str, list(g["vf_toi_sample_class"]) + list(g["ee_toi_sample_class"]))),
                             source="tests/golden/cloth_1415.npz (make_golden_c5.py: classified when the golden was made)")
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_parity_account.json"), "w"), indent=1)
    print("written")


if __name__ == "__main__":
    main()
