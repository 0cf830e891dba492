#!/usr/bin/env python
"""GPU time per detection step for BASELINE configs C1-C4, the unmodified reference's CPU time beside it (same box, one
core: the reference has no threading).  One JSON line per config on stdout; kept under profiles/ per round.

  C1  meshes/V0_prob3_402  -> V1   single step, fused ccd_step_shard (host buffers in, hit lists out)
  C2  meshes/V0_prob17_30957 -> V1 single step, same call; plus the two-call plugin boundary (ccd_broadphase_step +
                                   ccd_narrowphase: candidates to the host and back, what the C++ adapters do)
  C3  Model1_flow frames k -> k+1  broadphase with the fixed-vertex filter + CTCD narrowphase + SeparatingPlane narrowphase
                                   (example/testNewSequence.cpp:150-203), frames 0, 1, 60, 135 of the golden fixture
  C4  VelocityFilter passes        multi-entry History broadphase + SeparatingPlane narrowphase with per-stencil thickness
                                   (src/ActiveLayers.cpp:188-215), the five recorded passes of mesh1 -> mesh2

Inputs come from tests/golden/*.npz (generated from the reference's own meshes by tests/golden/make_golden*.py).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from collisiondetection_b200 import api  # noqa: E402
from oracle import bind  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def timed(fn, reps=7, warm=2):
    import torch
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3


def main():
    ctx = api.Context(0)
    ref = bind.Ref() if bind.have_ref() else None
    lines = []

    for cfg, name, eta in (("C1", "alec_prob3_402.npz", 1e-8), ("C2", "alec_prob17_30957.npz", 1e-8)):
        g = np.load(os.path.join(G, name))
        q0, q1, f = g["q0"], g["q1"], g["faces"]
        H = bind.single_step_history(q0, q1)
        r = ctx.step(13, f, q0, q1, eta, eta)
        nst = r["n_vf_candidates"] + r["n_ee_candidates"]
        ms_fused = timed(lambda: ctx.step(13, f, q0, q1, eta, eta, copy=False))

        def two_calls():
            vf, ee = ctx.findCollisionCandidatesStep(13, f, q0, q1, eta)
            ctx.findCollisions(*H, vf, eta, ee, eta)
        ms_two = timed(two_calls, reps=5)
        d = dict(config=cfg, workload=name[:-4], triangles=int(len(f)), stencils=int(nst), hits=int(r["n_vf_hits"] + r["n_ee_hits"]),
                 gpu_ms_fused_step_host_buffers=ms_fused, gpu_ms_device=r["ms_broadphase"] + r["ms_narrowphase"],
                 gpu_ms_two_call_boundary=ms_two, gpu_stencils_per_s=nst / (ms_fused * 1e-3))
        if ref is not None:
            t0 = time.perf_counter()
            vf, ee, bp_s = ref.broadphase(13, f, *H, eta)
            o = ref.narrowphase(*H, vf, eta, ee, eta)
            d.update(cpu_reference_s=time.perf_counter() - t0, cpu_broadphase_s=bp_s, cpu_narrowphase_s=o["seconds"], cpu_cores=1)
            d["speedup_vs_reference_one_core"] = d["cpu_reference_s"] * 1e3 / ms_fused
        lines.append(d)

    g = np.load(os.path.join(G, "model1_flow.npz"))
    eta, outer = float(g["eta"]), float(g["outer_eta"])
    for k in (0, 1, 60, 135):
        p = "f%d_" % k
        q0, q1, f, fixed = g[p + "q0"], g[p + "q1"], g[p + "faces"], g[p + "fixed"]
        H = bind.single_step_history(q0, q1)
        vf, ee = ctx.findCollisionCandidatesStep(13, f, q0, q1, outer, fixed)
        ms_bp = timed(lambda: ctx.findCollisionCandidatesStep(13, f, q0, q1, outer, fixed))
        ms_ctcd = timed(lambda: ctx.findCollisions(*H, vf, eta, ee, eta))
        ms_sp = timed(lambda: ctx.findCollisionsSeparatingPlane(*H, vf, eta, ee, eta))
        d = dict(config="C3", workload="Model1_flow frame %d -> %d (coarse dynamic + fine fixed)" % (k, k + 1), triangles=int(len(f)),
                 stencils=int(len(vf) + len(ee)), gpu_ms_broadphase=ms_bp, gpu_ms_ctcd_narrowphase=ms_ctcd, gpu_ms_sepplane_narrowphase=ms_sp)
        if ref is not None:
            _, _, bp_s = ref.broadphase(13, f, *H, outer, fixed)
            c = ref.narrowphase(*H, vf, eta, ee, eta)
            s = ref.narrowphase(*H, vf, eta, ee, eta, which=1)
            d.update(cpu_broadphase_s=bp_s, cpu_ctcd_narrowphase_s=c["seconds"], cpu_sepplane_narrowphase_s=s["seconds"], cpu_cores=1)
        lines.append(d)

    g = np.load(os.path.join(G, "velocityfilter.npz"))
    f, outer = g["mesh12_faces"], float(g["mesh12_outer"])
    for p in range(int(g["mesh12_npasses"])):
        k = "mesh12_p%d_" % p
        H = (g[k + "hoff"], g[k + "htime"], g[k + "hpos"])
        vf, ee = ctx.findCollisionCandidates(13, f, *H, outer)
        ve, ee_eta = g[k + "vf_eta"], g[k + "ee_eta"]
        ms_bp = timed(lambda: ctx.findCollisionCandidates(13, f, *H, outer))
        ms_sp = timed(lambda: ctx.findCollisionsSeparatingPlane(*H, vf, ve, ee, ee_eta))
        ms_ctcd = timed(lambda: ctx.findCollisions(*H, vf, ve, ee, ee_eta))
        os.environ["CCD_HISTORY_ONE_THREAD"] = "1"      # the one-thread-per-stencil kernels the staged paths replaced (same bits)
        os.environ["CCD_SEPPLANE_ONE_THREAD"] = "1"
        ms_ctcd_1t = timed(lambda: ctx.findCollisions(*H, vf, ve, ee, ee_eta), reps=3, warm=1)
        ms_sp_1t = timed(lambda: ctx.findCollisionsSeparatingPlane(*H, vf, ve, ee, ee_eta), reps=3, warm=1)
        del os.environ["CCD_HISTORY_ONE_THREAD"]
        del os.environ["CCD_SEPPLANE_ONE_THREAD"]
        d = dict(config="C4", workload="VelocityFilter mesh1 -> mesh2, detection pass %d (History of %d entries)" % (p, len(H[1])),
                 triangles=int(len(f)), stencils=int(len(vf) + len(ee)), gpu_ms_broadphase=ms_bp, gpu_ms_sepplane_narrowphase=ms_sp,
                 gpu_ms_ctcd_narrowphase_multi_entry=ms_ctcd, gpu_ms_ctcd_narrowphase_multi_entry_one_thread_per_stencil=ms_ctcd_1t,
                 gpu_ms_sepplane_narrowphase_one_thread_per_stencil=ms_sp_1t)
        if ref is not None:
            _, _, bp_s = ref.broadphase(13, f, *H, outer)
            s = ref.narrowphase(*H, vf, ve, ee, ee_eta, which=1)
            c = ref.narrowphase(*H, vf, ve, ee, ee_eta)
            d.update(cpu_broadphase_s=bp_s, cpu_sepplane_narrowphase_s=s["seconds"], cpu_ctcd_narrowphase_s=c["seconds"], cpu_cores=1)
        lines.append(d)

    for d in lines:
        print(json.dumps(d))


if __name__ == "__main__":
    main()
