"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libccdref.so).

Run in the build container (needs /root/reference for the mesh fixtures and the compiled reference):
    python tests/golden/make_golden.py
The GPU box has neither; tests there read only the committed .npz files.

Every array named ref_* was computed by the reference's own object code through oracle/ref_harness.cpp.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bind  # noqa: E402
from collisiondetection_b200 import scenes  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
MESHES = "/root/reference/meshes"


def mesh_pair(name):
    q0, f = scenes.load_obj(os.path.join(MESHES, "V0_%s.obj" % name))
    q1, _ = scenes.load_obj(os.path.join(MESHES, "V1_%s.obj" % name))
    return q0, q1, f


def full_case(ref, q0, q1, faces, outer_eta, eta, kind=13, fixed=None):
    H = bind.single_step_history(q0, q1)
    vf, ee, _ = ref.broadphase(kind, faces, *H, outer_eta, fixed)
    np_ = ref.narrowphase(*H, vf, eta, ee, eta)
    assert np_["disagree"] == 0
    return dict(q0=q0, q1=q1, faces=faces, outer_eta=outer_eta, eta=eta, kind=kind,
                fixed=np.zeros(0, np.uint8) if fixed is None else fixed,
                ref_vf=vf, ref_ee=ee, ref_vf_hit=np_["vf_hit"], ref_ee_hit=np_["ee_hit"], ref_vf_toi=np_["vf_toi"],
                ref_ee_toi=np_["ee_toi"], ref_vf_stage=np_["vf_stage"].astype(np.uint8), ref_ee_stage=np_["ee_stage"].astype(np.uint8))


def summary_case(ref, q0, q1, faces, outer_eta, eta, kind=13, keep_inputs=True):
    """Large meshes: inputs + counts, set hashes and the hit lists (candidate arrays are too big to commit)."""
    H = bind.single_step_history(q0, q1)
    vf, ee, _ = ref.broadphase(kind, faces, *H, outer_eta)
    np_ = ref.narrowphase(*H, vf, eta, ee, eta)
    assert np_["disagree"] == 0
    vh, eh = np_["vf_hit"] > 0, np_["ee_hit"] > 0
    d = dict(outer_eta=outer_eta, eta=eta, kind=kind, n_vf=len(vf), n_ee=len(ee),
             vf_fnv=bind.fnv1a64(vf), ee_fnv=bind.fnv1a64(ee),
             ref_vf_hits=vf[vh], ref_ee_hits=ee[eh], ref_vf_hit_toi=np_["vf_toi"][vh], ref_ee_hit_toi=np_["ee_toi"][eh])
    if keep_inputs:
        d.update(q0=q0, q1=q1, faces=faces)
    return d


def rand_prims(ref, rng, n):
    """Random primitive configurations around unit-scale geometry, with a sprinkle of degenerate ones."""
    out = {}
    for name, npts, fn in (("vf", 8, ref.vf_batch), ("ee", 8, ref.ee_batch), ("ve", 6, ref.ve_batch), ("vv", 4, ref.vv_batch)):
        half = npts // 2
        start = rng.uniform(-1, 1, (n, half, 3))
        vel = rng.uniform(-1.5, 1.5, (n, half, 3))
        # a third of the items: small motions (few roots), another third: some static vertices (degree drops)
        vel[: n // 3] *= 0.05
        static = rng.random((n, half)) < 0.25
        static[: 2 * n // 3] = False
        vel[static] = 0.0
        # a few exactly coplanar / coincident starts
        start[-20:, :, 2] = 0.0
        pts = np.concatenate([start, start + vel], axis=1).reshape(n, -1)
        eta = np.where(rng.random(n) < 0.5, 1e-6, 1e-2)
        hit, t = fn(pts, eta)
        out[name + "_pts"] = pts
        out[name + "_eta"] = eta
        out["ref_" + name + "_hit"] = hit
        out["ref_" + name + "_t"] = t
    return out


def rand_dist(ref, rng, n):
    pts = rng.uniform(-1, 1, (n, 12))
    pts[: n // 4, 9:12] = pts[: n // 4, 6:9] + rng.uniform(-1e-3, 1e-3, (n // 4, 3))   # near-degenerate
    pts[-8:, 3:6] = pts[-8:, 0:3]                                                      # zero-length first edge
    eta = rng.uniform(0.01, 0.5, n)
    vvec, vbary = ref.dist_vf_batch(pts)
    evec, ebary = ref.dist_ee_batch(pts)
    return dict(pts=pts, eta=eta, ref_vf_vec=vvec, ref_vf_bary=vbary, ref_ee_vec=evec, ref_ee_bary=ebary,
                ref_plane_lt=ref.dist_plane_lt_batch(pts, eta), ref_line_lt=ref.dist_line_lt_batch(pts, eta))


def multi_entry_history(q0, q1, rng, max_extra=3):
    """A History with 2..2+max_extra entries per vertex: random breakpoints, positions off the straight line."""
    V = q0.shape[0]
    extra = rng.integers(0, max_extra + 1, V)
    hoff = np.zeros(V + 1, np.int64)
    hoff[1:] = np.cumsum(extra + 2)
    htime = np.zeros(hoff[-1])
    hpos = np.zeros((hoff[-1], 3))
    scale = np.abs(q1 - q0).max() + 1e-9
    for v in range(V):
        ts = np.sort(rng.uniform(0.05, 0.95, extra[v]))
        a = hoff[v]
        htime[a] = 0.0
        hpos[a] = q0[v]
        for k, t in enumerate(ts):
            htime[a + 1 + k] = t
            hpos[a + 1 + k] = (1 - t) * q0[v] + t * q1[v] + rng.uniform(-0.05, 0.05, 3) * scale
        htime[hoff[v + 1] - 1] = 1.0
        hpos[hoff[v + 1] - 1] = q1[v]
    return hoff, htime, np.ascontiguousarray(hpos.reshape(-1))


def main():
    ref = bind.Ref()
    rng = np.random.default_rng(20261017)

    # known answers of example/testCTCD.cpp:10-68 (eta = 1e-6)
    ee = [0, 0, 0, 1, 0, 0, 2, 0, 0, 2, 1, 0, 0, 0, 0, 1, 0, 0, 2, 1, 0, -1, -1, 0]
    vf = [.5, .5, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, .5, .5, 1, 0, 0, 0, 1, 0, 3, 0, 1, 0]
    ve = [.5, 0, 1, 0, 0, 0, 1, 0, 0, .5, 0, -3, 0, 0, 0, 1, 0, 0]
    vv = [0, 0, 0, 1, 0, 0, 6, 0, 0, 1, 0, 0]
    ka = {}
    for name, pts, fn in (("ee", ee, ref.ee_batch), ("vf", vf, ref.vf_batch), ("ve", ve, ref.ve_batch), ("vv", vv, ref.vv_batch)):
        h, t = fn(np.array(pts, float), 1e-6)
        ka[name + "_pts"] = np.array(pts, float)
        ka["ref_" + name + "_hit"] = h
        ka["ref_" + name + "_t"] = t
        print("testCTCD", name, int(h[0]), "%.17g" % t[0])
    np.savez_compressed(os.path.join(OUT, "testctcd.npz"), **ka)

    # AlecTest flow (example/AlecTest.cpp:86-111) on the three small mesh pairs: everything kept
    for name in ("prob3_402", "prob11_835", "prob18_834"):
        q0, q1, f = mesh_pair(name)
        c = full_case(ref, q0, q1, f, 1e-8, 1e-8)
        print(name, len(c["ref_vf"]), len(c["ref_ee"]), int(c["ref_vf_hit"].sum()), int(c["ref_ee_hit"].sum()))
        np.savez_compressed(os.path.join(OUT, "alec_%s.npz" % name), **c)
    # AABB variant and fixed vertices on prob3
    q0, q1, f = mesh_pair("prob3_402")
    c = full_case(ref, q0, q1, f, 1e-8, 1e-8, kind=3)
    np.savez_compressed(os.path.join(OUT, "alec_prob3_402_aabb.npz"), **c)
    fixed = (rng.random(q0.shape[0]) < 0.6).astype(np.uint8)
    c = full_case(ref, q0, q1, f, 1e-8, 1e-8, kind=13, fixed=fixed)
    print("fixed", len(c["ref_vf"]), len(c["ref_ee"]))
    np.savez_compressed(os.path.join(OUT, "alec_prob3_402_fixed.npz"), **c)
    # a thicker eta on prob3 (more hits, VE/VV stages exercised)
    c = full_case(ref, q0, q1, f, 2e-3, 1e-3)
    print("thick", len(c["ref_vf"]), len(c["ref_ee"]), int(c["ref_vf_hit"].sum()), int(c["ref_ee_hit"].sum()),
          np.bincount(c["ref_vf_stage"]), np.bincount(c["ref_ee_stage"]))
    np.savez_compressed(os.path.join(OUT, "alec_prob3_402_thick.npz"), **c)

    # multi-entry History on prob3 (leaf boxes span all entries; narrowphase walks stitched segments)
    H = multi_entry_history(q0, q1, rng)
    vf, ee, _ = ref.broadphase(13, f, *H, 1e-4)
    np_ = ref.narrowphase(*H, vf, 5e-5, ee, 5e-5)
    assert np_["disagree"] == 0
    print("history", H[0][-1], len(vf), len(ee), int(np_["vf_hit"].sum()), int(np_["ee_hit"].sum()))
    np.savez_compressed(os.path.join(OUT, "history_prob3_402.npz"), faces=f, hoff=H[0], htime=H[1], hpos=H[2], outer_eta=1e-4,
                        eta=5e-5, ref_vf=vf, ref_ee=ee, ref_vf_hit=np_["vf_hit"], ref_ee_hit=np_["ee_hit"],
                        ref_vf_toi=np_["vf_toi"], ref_ee_toi=np_["ee_toi"],
                        ref_vf_stage=np_["vf_stage"].astype(np.uint8), ref_ee_stage=np_["ee_stage"].astype(np.uint8))

    # BASELINE config C2: prob17_30957 — inputs + summary
    q0, q1, f = mesh_pair("prob17_30957")
    c = summary_case(ref, q0, q1, f, 1e-8, 1e-8)
    print("prob17", c["n_vf"], c["n_ee"], len(c["ref_vf_hits"]), len(c["ref_ee_hits"]), c["vf_fnv"], c["ee_fnv"])
    np.savez_compressed(os.path.join(OUT, "alec_prob17_30957.npz"), **c)

    # synthetic cloth twin n=101 (SURVEY.md §8d calibration: 57,464 VF + 97,726 EE candidates)
    q0, q1, f, eta = scenes.cloth(101)
    c = summary_case(ref, q0, q1, f, eta, eta, keep_inputs=False)
    print("cloth101", c["n_vf"], c["n_ee"], len(c["ref_vf_hits"]), len(c["ref_ee_hits"]), c["vf_fnv"], c["ee_fnv"])
    np.savez_compressed(os.path.join(OUT, "cloth_101.npz"), **c)

    # random primitive and distance batches
    np.savez_compressed(os.path.join(OUT, "prims_random.npz"), **rand_prims(ref, rng, 3000))
    np.savez_compressed(os.path.join(OUT, "dist_random.npz"), **rand_dist(ref, rng, 2000))

    # Distance::meshSelfDistance (src/Distance.cpp:12-66)
    msd = {}
    for name in ("prob3_402", "prob11_835"):
        q0, _, f = mesh_pair(name)
        d, _ = ref.mesh_self_distance(q0, f)
        msd[name] = d
        print("meshSelfDistance", name, "%.17g" % d)
    np.savez_compressed(os.path.join(OUT, "mesh_self_distance.npz"), **msd)


if __name__ == "__main__":
    main()
