"""CPU checks of the C restatement (oracle/ccd_oracle.c) against the golden vectors that
tests/golden/make_golden.py produced with the unmodified reference, and — when oracle/_ref is built —
against the reference itself.  No GPU involved."""
import numpy as np
import pytest

from conftest import golden
from oracle import bind


def _single(g):
    return bind.single_step_history(g["q0"], g["q1"])


def test_known_answers_testctcd(port):
    """example/testCTCD.cpp:10-68 — four hard-coded cases, eta 1e-6; bit-identical t."""
    g = golden("testctcd.npz")
    for name, fn in (("ee", port.ee_batch), ("vf", port.vf_batch), ("ve", port.ve_batch), ("vv", port.vv_batch)):
        hit, t = fn(g[name + "_pts"], 1e-6)
        assert hit[0] == 1 and g["ref_" + name + "_hit"][0] == 1
        assert t[0] == g["ref_" + name + "_t"][0], name
    assert "%.17g" % g["ref_vf_t"][0] == "0.66666517599038733"
    assert "%.17g" % g["ref_ee_t"][0] == "0.50000000000000022"
    assert "%.17g" % g["ref_ve_t"][0] == "0.24999975000276525"
    assert "%.17g" % g["ref_vv_t"][0] == "0.1666665000090396"


@pytest.mark.parametrize("name", ["alec_prob3_402", "alec_prob11_835", "alec_prob18_834", "alec_prob3_402_aabb",
                                  "alec_prob3_402_fixed", "alec_prob3_402_thick"])
def test_candidate_sets_bit_exact(port, name):
    g = golden(name + ".npz")
    fixed = g["fixed"] if g["fixed"].size else None
    vf, ee, _ = port.broadphase(int(g["kind"]), g["faces"], *_single(g), float(g["outer_eta"]), fixed)
    assert np.array_equal(vf, g["ref_vf"])
    assert np.array_equal(ee, g["ref_ee"])


# (name, allowed flag mismatches vs rpoly): every mismatch must be decided for the restatement by the arbiter
@pytest.mark.parametrize("name,max_mismatch", [("alec_prob3_402", 0), ("alec_prob11_835", 2), ("alec_prob18_834", 0),
                                                ("alec_prob3_402_thick", 0)])
def test_flags_and_toi_against_reference(port, name, max_mismatch):
    from arbiter import Arbiter
    g = golden(name + ".npz")
    eta = float(g["eta"])
    out = port.narrowphase(*_single(g), g["ref_vf"], eta, g["ref_ee"], eta)
    arb = Arbiter(port)
    total_mismatch = 0
    for k, fn in (("vf", arb.vertex_face), ("ee", arb.edge_edge)):
        st = g["ref_" + k]
        ph, rh = out[k + "_hit"], g["ref_%s_hit" % k]
        mism = np.nonzero(ph != rh)[0]
        total_mismatch += len(mism)
        for i in mism:
            pts = np.concatenate([g["q0"][st[i]].reshape(-1), g["q1"][st[i]].reshape(-1)])
            ah, _, _ = fn(pts, eta)
            assert ah == bool(ph[i]), "stencil %s: restatement disagrees with reference AND with the 60-digit arbiter" % (st[i],)
        both = (ph > 0) & (rh > 0)
        assert np.array_equal(out[k + "_stage"][both], g["ref_%s_stage" % k][both])
        # degenerate-stage TOIs (closed-form quadratics) and well-conditioned roots agree to 1e-9 relative;
        # rpoly is off by up to ~1e-4 on nearly-double root pairs (tests/arbiter.py), hence the looser bound here.
        rt, pt = g["ref_%s_toi" % k][both], out[k + "_toi"][both]
        rel = np.abs(rt - pt) / np.maximum(np.abs(rt), 1e-300)
        assert rel.max(initial=0) < 2e-3
        assert np.median(rel) < 1e-9 if rel.size else True
    assert total_mismatch <= max_mismatch


def test_toi_accuracy_against_arbiter(port):
    """Where the restatement and rpoly differ by > 1e-9, the restatement is the one within ~1e-8 of the exact root."""
    from arbiter import Arbiter
    g = golden("alec_prob11_835.npz")
    eta = float(g["eta"])
    out = port.narrowphase(*_single(g), g["ref_vf"], eta, g["ref_ee"], eta)
    arb = Arbiter(port)
    st = g["ref_ee"]
    both = np.nonzero((out["ee_hit"] > 0) & (g["ref_ee_hit"] > 0) & (out["ee_stage"] == 1))[0]
    rel = np.abs(out["ee_toi"][both] - g["ref_ee_toi"][both]) / np.maximum(np.abs(g["ref_ee_toi"][both]), 1e-300)
    worst = both[np.argsort(-rel)][:15]
    for i in worst:
        pts = np.concatenate([g["q0"][st[i]].reshape(-1), g["q1"][st[i]].reshape(-1)])
        ah, at, _ = arb.edge_edge(pts, eta)
        assert ah
        assert abs(out["ee_toi"][i] - at) <= 5e-8 * abs(at)
        assert abs(out["ee_toi"][i] - at) <= abs(g["ref_ee_toi"][i] - at)


def test_multi_entry_history(port):
    g = golden("history_prob3_402.npz")
    H = (g["hoff"], g["htime"], g["hpos"])
    vf, ee, _ = port.broadphase(13, g["faces"], *H, float(g["outer_eta"]))
    assert np.array_equal(vf, g["ref_vf"]) and np.array_equal(ee, g["ref_ee"])
    out = port.narrowphase(*H, vf, float(g["eta"]), ee, float(g["eta"]))
    for k in ("vf", "ee"):
        assert np.array_equal(out[k + "_hit"], g["ref_%s_hit" % k])
        both = out[k + "_hit"] > 0
        assert np.array_equal(out[k + "_stage"][both], g["ref_%s_stage" % k][both])
        # TOI in History time (ta + t (tb - ta) of the stitched segment that hit), north_star's 1e-9 relative
        rel = np.abs(out[k + "_toi"][both] - g["ref_%s_toi" % k][both]) / np.maximum(np.abs(g["ref_%s_toi" % k][both]), 1e-300)
        assert rel.max(initial=0) < 1e-9
        assert np.all((out[k + "_toi"][both] >= 0) & (out[k + "_toi"][both] <= 1))


def test_random_primitives(port):
    g = golden("prims_random.npz")
    for k, fn in (("vf", port.vf_batch), ("ee", port.ee_batch), ("ve", port.ve_batch), ("vv", port.vv_batch)):
        hit, t = fn(g[k + "_pts"], g[k + "_eta"])
        rh, rt = g["ref_%s_hit" % k], g["ref_%s_t" % k]
        n_mis = int((hit != rh).sum())
        assert n_mis <= (0 if k in ("vv",) else 3), (k, n_mis)
        both = (hit > 0) & (rh > 0)
        rel = np.abs(t[both] - rt[both]) / np.maximum(np.abs(rt[both]), 1e-300)
        if k == "vv":
            assert np.array_equal(t[both], rt[both])        # closed form, same operations -> same bits
        else:
            # rpoly returns nearly-double roots only to ~1e-5 (tests/arbiter.py); the typical case agrees to 1e-9
            assert np.median(rel) < 1e-9 and rel.max() < 2e-3, (k, rel.max())
        if k in ("vf", "ee"):
            # on the worst disagreements the restatement must be the one next to the exact root
            from arbiter import Arbiter
            arb = Arbiter(port)
            idx = np.nonzero(both)[0][np.argsort(-rel)][:5]
            for i in idx:
                ah, at, _ = (arb.vertex_face if k == "vf" else arb.edge_edge)(g[k + "_pts"][i], float(g[k + "_eta"][i]))
                assert ah and abs(t[i] - at) <= max(1e-7 * abs(at), abs(rt[i] - at)), (k, i, t[i], rt[i], at)


def test_distance_queries_bit_exact(port):
    g = golden("dist_random.npz")
    vec, bary = port.dist_vf_batch(g["pts"])
    assert np.array_equal(vec, g["ref_vf_vec"]) and np.array_equal(bary, g["ref_vf_bary"])
    vec, bary = port.dist_ee_batch(g["pts"])
    assert np.array_equal(vec, g["ref_ee_vec"]) and np.array_equal(bary, g["ref_ee_bary"])
    assert np.array_equal(port.dist_plane_lt_batch(g["pts"], g["eta"]), g["ref_plane_lt"])
    assert np.array_equal(port.dist_line_lt_batch(g["pts"], g["eta"]), g["ref_line_lt"])


def test_mesh_self_distance(port):
    g = golden("mesh_self_distance.npz")
    a = golden("alec_prob3_402.npz")
    d, _ = port.mesh_self_distance(a["q0"], a["faces"])
    assert d == float(g["prob3_402"]) == 1.4204876861184471e-13


def test_cloth_twin_candidates(port):
    """SURVEY.md §8(d) calibration: n=101 -> 57,464 VF + 97,726 EE candidates; sets hashed against the reference's."""
    from collisiondetection_b200 import scenes
    g = golden("cloth_101.npz")
    q0, q1, f, eta = scenes.cloth(101)
    vf, ee, _ = port.broadphase(13, f, *bind.single_step_history(q0, q1), eta)
    assert (len(vf), len(ee)) == (57464, 97726) == (int(g["n_vf"]), int(g["n_ee"]))
    assert bind.fnv1a64(vf) == str(g["vf_fnv"]) and bind.fnv1a64(ee) == str(g["ee_fnv"])


def test_port_against_live_reference(port, ref):
    """When the compiled reference is present: same candidate sets and flags on a fresh random scene."""
    from collisiondetection_b200 import scenes
    q0, q1, f, eta = scenes.cloth(41)
    H = bind.single_step_history(q0, q1)
    for kind in (13, 3):
        a, b, _ = ref.broadphase(kind, f, *H, eta)
        a2, b2, _ = port.broadphase(kind, f, *H, eta)
        assert np.array_equal(a, a2) and np.array_equal(b, b2)
    r = ref.narrowphase(*H, a, eta, b, eta)
    p = port.narrowphase(*H, a, eta, b, eta)
    assert r["disagree"] == 0
    from arbiter import Arbiter
    arb = Arbiter(port)
    verdicts = {}
    for k, st in (("vf", a), ("ee", b)):
        for i in np.nonzero(r[k + "_hit"] != p[k + "_hit"])[0]:
            pts = np.concatenate([q0[st[i]].reshape(-1), q1[st[i]].reshape(-1)])
            v = arb.classify(k, pts, eta, p[k + "_hit"][i], r[k + "_hit"][i])
            verdicts[v] = verdicts.get(v, 0) + 1
    assert verdicts.get("unexplained", 0) == 0, verdicts
    assert sum(verdicts.values()) <= 20, verdicts


@pytest.mark.slow
def test_prob17_full_size(port):
    """BASELINE config C2 (V0/V1_prob17_30957, 77,386 faces): candidate sets bit-exact (count + FNV of the sorted set);
    flags equal to the reference's except stencils whose decision is a reference artefact (rpoly's inexact or missing
    roots; the 60-digit arbiter sides with the restatement) or not stable at rounding level (noise-sign)."""
    from arbiter import Arbiter
    g = golden("alec_prob17_30957.npz")
    q0, q1 = g["q0"], g["q1"]
    H = bind.single_step_history(q0, q1)
    vf, ee, _ = port.broadphase(13, g["faces"], *H, 1e-8)
    assert (len(vf), len(ee)) == (1330564, 2370945)
    assert bind.fnv1a64(vf) == str(g["vf_fnv"]) and bind.fnv1a64(ee) == str(g["ee_fnv"])
    out = port.narrowphase_flat(*H, vf, 1e-8, ee, 1e-8)
    arb = Arbiter(port)
    verdicts = {}
    for k, st in (("vf", vf), ("ee", ee)):
        refhits = set(map(tuple, g["ref_%s_hits" % k].tolist()))
        mine = out[k + "_hit"] > 0
        minehits = set(map(tuple, st[mine].tolist()))
        assert len(refhits) == (23021 if k == "vf" else 63521)
        for s in sorted(refhits ^ minehits):
            pts = np.concatenate([q0[list(s)].reshape(-1), q1[list(s)].reshape(-1)])
            v = arb.classify(k, pts, 1e-8, s in minehits, s in refhits)
            verdicts[v] = verdicts.get(v, 0) + 1
    print("prob17 flag mismatches vs reference:", verdicts)
    assert verdicts.get("unexplained", 0) == 0, verdicts
    assert sum(verdicts.values()) <= 200, verdicts


SP_SCENES = ["alec_prob3_402", "alec_prob11_835", "alec_prob18_834", "history_prob3_402", "alec_prob3_402_thick", "alec_prob3_402_fixed"]


@pytest.mark.parametrize("name", SP_SCENES)
def test_separating_plane_restatement_matches_golden(port, name):
    """SeparatingPlaneNarrowPhase (src/SeparatingPlaneNarrowPhase.cpp:11-278) restated in plain C (oracle/ccd_oracle.c:
    orc_narrowphase_sepplane) against the flags of the unmodified reference class (tests/golden/sepplane.npz, made by
    tests/golden/make_golden_sepplane.py): identical on every candidate.  prob11: 840 VF / 2,212 EE."""
    g = golden(name + ".npz")
    sp = golden("sepplane.npz")
    H = (g["hoff"], g["htime"], g["hpos"]) if "hoff" in g.files else _single(g)
    eta = float(g["eta"])
    p = port.narrowphase(*H, g["ref_vf"], eta, g["ref_ee"], eta, which=1)
    assert np.array_equal(p["vf_hit"], sp[name + "_vf"]) and np.array_equal(p["ee_hit"], sp[name + "_ee"])
    if name == "alec_prob11_835":
        assert int(p["vf_hit"].sum()) == 840 and int(p["ee_hit"].sum()) == 2212


@pytest.mark.parametrize("k", [0, 1, 60, 135])
def test_model1_flow_frames(port, k):
    """BASELINE config C3: Model1_flow frame k -> k+1 as example/testNewSequence.cpp:150-203 assembles it (coarse mesh
    dynamic, fine frame fixed, outer / inner radius 1e-3 / 1e-4).  Restatement against the unmodified reference
    (tests/golden/model1_flow.npz): candidate sets with the fixed-vertex filter, CTCD flags / TOI / stage, SeparatingPlane flags."""
    g = golden("model1_flow.npz")
    p = "f%d_" % k
    H = bind.single_step_history(g[p + "q0"], g[p + "q1"])
    vf, ee, _ = port.broadphase(13, g[p + "faces"], *H, float(g["outer_eta"]), g[p + "fixed"])
    assert np.array_equal(vf, g[p + "vf"]) and np.array_equal(ee, g[p + "ee"])
    eta = float(g["eta"])
    a = port.narrowphase(*H, vf, eta, ee, eta)
    for nm in ("vf", "ee"):
        assert np.array_equal(a[nm + "_hit"], g[p + nm + "_hit"])
        hit = g[p + nm + "_hit"] > 0
        assert np.array_equal(a[nm + "_stage"][hit], g[p + nm + "_stage"][hit])
        assert np.allclose(a[nm + "_toi"][hit], g[p + nm + "_toi"][hit], rtol=1e-9, atol=0)
    b = port.narrowphase(*H, vf, eta, ee, eta, which=1)
    assert np.array_equal(b["vf_hit"], g[p + "sp_vf_hit"]) and np.array_equal(b["ee_hit"], g[p + "sp_ee_hit"])


@pytest.mark.parametrize("p", [0, 1, 2, 3, 4])
def test_velocityfilter_detection_passes(port, p):
    """BASELINE config C4: detection pass p of ActiveLayers inside the unmodified VelocityFilter::velocityFilter on
    mesh1.obj -> mesh2.obj (example/testVelocityFilter.cpp, radii 2e-8 / 1e-8), recorded by oracle/ref_recorder.* into
    tests/golden/velocityfilter.npz: a History that grows from 394 to 1,038 entries, every broadphase candidate with the
    per-stencil thickness ActiveLayers assigns, the SeparatingPlane hits.  SURVEY.md 8(c) lists the same counts:
    (394; 2299/3682 -> 30/23), (858; 2304/3691 -> 30/11), (1022; 2287/3658 -> 18/8), (1038; ... -> 6/0), (1038; ... -> 3/0)."""
    g = golden("velocityfilter.npz")
    k = "mesh12_p%d_" % p
    H = (g[k + "hoff"], g[k + "htime"], g[k + "hpos"])
    expect = [(394, 2299, 3682, 30, 23), (858, 2304, 3691, 30, 11), (1022, 2287, 3658, 18, 8), (1038, 2287, 3658, 6, 0), (1038, 2287, 3658, 3, 0)][p]
    assert (len(H[1]), len(g[k + "vf"]), len(g[k + "ee"]), int(g[k + "vf_hit"].sum()), int(g[k + "ee_hit"].sum())) == expect
    vf, ee, _ = port.broadphase(13, g["mesh12_faces"], *H, float(g["mesh12_outer"]))
    assert np.array_equal(vf, g[k + "vf"]) and np.array_equal(ee, g[k + "ee"])
    r = port.narrowphase(*H, vf, g[k + "vf_eta"], ee, g[k + "ee_eta"], which=1)
    assert np.array_equal(r["vf_hit"], g[k + "vf_hit"]) and np.array_equal(r["ee_hit"], g[k + "ee_hit"])
