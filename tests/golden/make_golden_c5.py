"""Generates tests/golden/cloth_1415.npz — BASELINE config C5 (the 4M-triangle cloth step) — from the UNMODIFIED
reference (oracle/_ref/libccdref.so), plus the comparison of the plain-C restatement (= the GPU, bit for bit) with it.

    python tests/golden/make_golden_c5.py ref        # ~10 min, ~6 GB: reference KDOPBroadPhase + CTCDNarrowPhase, raw -> /tmp
    python tests/golden/make_golden_c5.py port       # ~10 min: restatement narrowphase on the same stencils, raw -> /tmp
    python tests/golden/make_golden_c5.py pack       # classifies every difference (tests/arbiter.py) and writes the .npz

The candidate arrays (488 MB) are not committed: the golden holds their counts and FNV-1a-64, the reference's hit
flags as bit masks over the sorted candidate lists (so the full hit lists are recoverable from any bit-exact
candidate list), the stage of every hit, every 16th hit's time of impact in full precision, the earliest time of
impact, and the classified list of every stencil on which the restatement's flag differs from the reference's.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import bind  # noqa: E402
from collisiondetection_b200 import scenes  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cloth_1415.npz")
N = int(os.environ.get("C5_N", "1415"))
TMP = os.environ.get("C5_TMP", "/tmp/c5_golden_%d" % N)
if N != 1415:
    OUT = OUT.replace("cloth_1415", "cloth_%d" % N)


def stage_ref():
    os.makedirs(TMP, exist_ok=True)
    q0, q1, f, eta = scenes.cloth(N)
    H = bind.single_step_history(q0, q1)
    ref = bind.Ref()
    t0 = time.time()
    vf, ee, bp_s = ref.broadphase(13, f, *H, eta)
    print("reference broadphase: %d VF + %d EE, %.1f s (call %.1f s)" % (len(vf), len(ee), bp_s, time.time() - t0), flush=True)
    np.save(os.path.join(TMP, "vf.npy"), vf)
    np.save(os.path.join(TMP, "ee.npy"), ee)
    t0 = time.time()
    r = ref.narrowphase(*H, vf, eta, ee, eta)
    print("reference narrowphase: %d + %d hits, %.1f s (call %.1f s)" % (r["vf_hit"].sum(), r["ee_hit"].sum(), r["seconds"], time.time() - t0), flush=True)
    assert r["disagree"] == 0
    for k in ("vf", "ee"):
        np.save(os.path.join(TMP, "ref_%s_hit.npy" % k), r[k + "_hit"])
        np.save(os.path.join(TMP, "ref_%s_toi.npy" % k), r[k + "_toi"])
        np.save(os.path.join(TMP, "ref_%s_stage.npy" % k), r[k + "_stage"].astype(np.uint8))
    np.save(os.path.join(TMP, "ref_seconds.npy"), np.array([bp_s, r["seconds"]]))


def stage_port():
    q0, q1, f, eta = scenes.cloth(N)
    H = bind.single_step_history(q0, q1)
    vf = np.load(os.path.join(TMP, "vf.npy"))
    ee = np.load(os.path.join(TMP, "ee.npy"))
    port = bind.Port()
    t0 = time.time()
    p = port.narrowphase(*H, vf, eta, ee, eta)
    print("restatement narrowphase: %d + %d hits, %.1f s" % (p["vf_hit"].sum(), p["ee_hit"].sum(), time.time() - t0), flush=True)
    for k in ("vf", "ee"):
        np.save(os.path.join(TMP, "port_%s_hit.npy" % k), p[k + "_hit"])
        np.save(os.path.join(TMP, "port_%s_toi.npy" % k), p[k + "_toi"])
        np.save(os.path.join(TMP, "port_%s_stage.npy" % k), p[k + "_stage"].astype(np.uint8))


def stage_pack():
    from arbiter import Arbiter
    import parity_account as PA
    q0, q1, f, eta = scenes.cloth(N)
    port = bind.Port()
    arb = Arbiter(port)
    d = dict(n=N, eta=eta, outer_eta=eta, kind=13)
    secs = np.load(os.path.join(TMP, "ref_seconds.npy"))
    d["ref_seconds_broadphase"], d["ref_seconds_narrowphase"] = float(secs[0]), float(secs[1])
    earliest = np.inf
    for k in ("vf", "ee"):
        st = np.load(os.path.join(TMP, "%s.npy" % k))
        rh = np.load(os.path.join(TMP, "ref_%s_hit.npy" % k)) > 0
        rt = np.load(os.path.join(TMP, "ref_%s_toi.npy" % k))
        rs = np.load(os.path.join(TMP, "ref_%s_stage.npy" % k))
        ph = np.load(os.path.join(TMP, "port_%s_hit.npy" % k)) > 0
        pt = np.load(os.path.join(TMP, "port_%s_toi.npy" % k))
        ps = np.load(os.path.join(TMP, "port_%s_stage.npy" % k))
        d["n_" + k] = len(st)
        d[k + "_fnv"] = bind.fnv1a64(st)
        d["ref_%s_hit_bits" % k] = np.packbits(rh)
        d["ref_%s_n_hits" % k] = int(rh.sum())
        d["ref_%s_hit_stage" % k] = rs[rh]
        hit_toi = rt[rh]
        d["ref_%s_hit_toi_16th" % k] = hit_toi[::16].copy()
        earliest = min(earliest, float(hit_toi.min()))
        # ---- flags: every difference classified by the 60-digit arbiter
        mism = np.nonzero(rh != ph)[0]
        classes = []
        for i in mism:
            pts = np.concatenate([q0[st[i]].reshape(-1), q1[st[i]].reshape(-1)])
            classes.append(PA.classify_flag(arb, k, pts, eta, bool(ph[i]), bool(rh[i]), int(ps[i]), int(rs[i])))
        d["%s_mismatch_index" % k] = mism.astype(np.int64)
        d["%s_mismatch_stencil" % k] = st[mism]
        d["%s_mismatch_mine" % k] = ph[mism].astype(np.uint8)
        d["%s_mismatch_class" % k] = np.array(classes, dtype="U24")
        print(k, "flag mismatches:", len(mism), dict(zip(*np.unique(np.array(classes, dtype="U24"), return_counts=True))) if len(mism) else {}, flush=True)
        # ---- time of impact on the stencils both report as hits
        both = np.nonzero(rh & ph)[0]
        rel = np.abs(pt[both] - rt[both]) / np.maximum(np.abs(rt[both]), 1e-300)
        out_tol = both[rel > 1e-9]
        d["%s_toi_both" % k] = len(both)
        d["%s_toi_out_of_1e9" % k] = len(out_tol)
        d["%s_toi_rel_max" % k] = float(rel.max(initial=0))
        d["%s_toi_rel_median" % k] = float(np.median(rel)) if len(rel) else 0.0
        # both hit, but through different sub-tests (e.g. the reference's primitive missed and a vertex-edge test caught
        # it): the earlier of the two sub-tests is where the runs disagree -> that disagreement is classified
        sd = both[ps[both] != rs[both]]
        sdc = []
        for i in sd:
            pts = np.concatenate([q0[st[i]].reshape(-1), q1[st[i]].reshape(-1)])
            first = int(min(ps[i], rs[i]))
            sdc.append(PA.classify_flag(arb, k, pts, eta, bool(ps[i] == first), bool(rs[i] == first), first, first))
        d["%s_stage_differs" % k] = len(sd)
        d["%s_stage_differs_index" % k] = sd.astype(np.int64)
        d["%s_stage_differs_class" % k] = np.array(sdc, dtype="U24")
        print(k, "both hit, different sub-test:", len(sd), dict(zip(*np.unique(np.array(sdc, dtype="U24"), return_counts=True))) if len(sd) else {}, flush=True)
        # the out-of-tolerance ones: a deterministic sample of at most 2000 is arbitrated (all of them when fewer)
        pick = out_tol if len(out_tol) <= 2000 else out_tol[np.linspace(0, len(out_tol) - 1, 2000).astype(np.int64)]
        tcls = []
        for i in pick:
            pts = np.concatenate([q0[st[i]].reshape(-1), q1[st[i]].reshape(-1)])
            tcls.append(PA.classify_toi(arb, port, k, pts, eta, float(pt[i]), float(rt[i]), int(ps[i])))
        d["%s_toi_sample_index" % k] = pick.astype(np.int64)
        d["%s_toi_sample_class" % k] = np.array(tcls, dtype="U24")
        d["%s_toi_sample_mine" % k] = pt[pick]
        d["%s_toi_sample_ref" % k] = rt[pick]
        print(k, "TOI: both", len(both), "out of 1e-9:", len(out_tol), "max rel", d["%s_toi_rel_max" % k],
              dict(zip(*np.unique(np.array(tcls, dtype="U24"), return_counts=True))) if len(pick) else {}, flush=True)
    d["ref_earliest_toi"] = earliest
    np.savez_compressed(OUT, **d)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("ref", "all"):
        stage_ref()
    if what in ("port", "all"):
        stage_port()
    if what in ("pack", "all"):
        stage_pack()
