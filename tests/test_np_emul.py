"""Host emulation of the single-step narrowphase pipeline (tests/np_emul.cu: the per-item functions of the stage kernels,
compiled for the CPU) against the CPU checker: flags and TOI bits must be identical.  Runs without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import bind

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "_build", "libnp_emul.so")
SRC = os.path.join(HERE, "np_emul.cu")
CSRC = os.path.join(ROOT, "collisiondetection_b200", "csrc")


def _build():
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["nvcc", "-O2", "--fmad=false", "-std=c++17", "-Wno-deprecated-gpu-targets", "-diag-suppress", "128", "-Xcompiler", "-ffp-contract=off,-fPIC",
                           "-shared", SRC, "-o", SO])


@pytest.fixture(scope="module")
def emul():
    _build()
    lib = C.CDLL(SO)
    lib.np_emul.restype = C.c_int
    lib.np_emul_roots.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def port():
    if not bind.have_port():
        bind.build(ref=False, port=True)
    return bind.Port()


def run_emul(lib, is_vf, st, q0, q1, eta):
    st = np.ascontiguousarray(st, dtype=np.int32)
    n = len(st)
    hit = np.zeros(n, dtype=np.uint8)
    toi = np.zeros(n, dtype=np.float64)
    stage = np.zeros(n, dtype=np.uint8)
    stats = np.zeros(11, dtype=np.int64)
    q0 = np.ascontiguousarray(q0, dtype=np.float64).reshape(-1)
    q1 = np.ascontiguousarray(q1, dtype=np.float64).reshape(-1)
    lib.np_emul(C.c_int(int(is_vf)), C.c_longlong(n), st.ctypes.data_as(C.c_void_p), q0.ctypes.data_as(C.c_void_p), q1.ctypes.data_as(C.c_void_p),
                C.c_int(3), None, C.c_double(eta), hit.ctypes.data_as(C.c_void_p), toi.ctypes.data_as(C.c_void_p), stage.ctypes.data_as(C.c_void_p),
                stats.ctypes.data_as(C.c_void_p))
    return hit, toi, stage, stats


def check_scene(emul, port, q0, q1, vf, ee, eta):
    H = bind.single_step_history(q0, q1)
    p = port.narrowphase(*H, vf, eta, ee, eta)
    out = {}
    for name, is_vf, st in (("vf", True, vf), ("ee", False, ee)):
        hit, toi, stage, stats = run_emul(emul, is_vf, st, q0, q1, eta)
        assert stats[0] == 0, "pipeline inconsistency"
        bad = np.nonzero(hit != p[name + "_hit"])[0]
        assert len(bad) == 0, "%s: %d flag mismatches, first %s" % (name, len(bad), bad[:5])
        tb = np.nonzero(toi.view(np.uint64) != p[name + "_toi"].view(np.uint64))[0]
        assert len(tb) == 0, "%s: %d TOI mismatches, first %s" % (name, len(tb), tb[:5])
        if name + "_stage" in p:
            assert np.array_equal(stage, p[name + "_stage"])
        out[name] = stats
    return out


@pytest.mark.parametrize("name", ["alec_prob3_402", "alec_prob11_835", "alec_prob18_834", "alec_prob3_402_thick"])
def test_emul_matches_checker_on_golden_scenes(emul, port, name):
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    eta = float(g["eta"]) if "eta" in g.files else 1e-8
    stats = check_scene(emul, port, g["q0"], g["q1"], g["ref_vf"], g["ref_ee"], eta)
    assert stats["vf"][3] + stats["ee"][3] <= 0.02 * (len(g["ref_vf"]) + len(g["ref_ee"])) + 8, "general routine used too often: %s" % stats


def test_emul_matches_checker_on_cloth(emul, port):
    from collisiondetection_b200 import scenes
    q0, q1, f, eta = scenes.cloth(61)
    H = bind.single_step_history(q0, q1)
    vf, ee, _ = port.broadphase(13, f, *H, eta)
    stats = check_scene(emul, port, q0, q1, vf, ee, eta)
    assert stats["vf"][3] + stats["ee"][3] <= 0.01 * (len(vf) + len(ee)) + 8


def test_lane_isolator_matches_checker(emul, port):
    """RootLane (ccd_solve.cuh) against orc_roots01 on random and clustered-root polynomials."""
    rng = np.random.default_rng(7)
    lib = port.lib
    lib.orc_roots01.restype = C.c_int
    for d in (3, 4, 5, 6):
        for trial in range(4000):
            if trial % 2:
                r = rng.uniform(-0.3, 1.3, size=d)
                if trial % 4 == 1:
                    r[1] = r[0] + rng.normal() * 1e-7
                c = np.poly(r) * rng.uniform(0.5, 2.0)
            else:
                c = rng.normal(size=d + 1)
            c = np.ascontiguousarray(c / np.abs(c).max(), dtype=np.float64)
            if c[0] == 0:
                continue
            a = np.zeros(8)
            b = np.zeros(8)
            na = emul.np_emul_roots(C.c_int(d), c.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p))
            nb = lib.orc_roots01(c.ctypes.data_as(C.c_void_p), C.c_int(d), b.ctypes.data_as(C.c_void_p))
            assert na == nb, (d, trial, c, a, b)
            assert np.array_equal(a[:na].view(np.uint64), b[:nb].view(np.uint64)), (d, trial, c, a, b)


def _sepplane_emul(lib, is_vf, st, eta, H):
    st = np.ascontiguousarray(st, dtype=np.int32)
    hoff, htime, hpos = (np.ascontiguousarray(H[0], dtype=np.int64), np.ascontiguousarray(H[1], dtype=np.float64), np.ascontiguousarray(H[2], dtype=np.float64).reshape(-1))
    gaps = [np.diff(htime[hoff[v]:hoff[v + 1]]).min() for v in range(len(hoff) - 1) if hoff[v + 1] - hoff[v] > 1]
    eps = min(1.0, min(gaps)) / 4.0
    eta = np.ascontiguousarray(np.broadcast_to(np.asarray(eta, dtype=np.float64), (len(st),)))
    hit = np.zeros(len(st), np.uint8)
    lib.np_emul_sepplane.restype = C.c_int
    bad = lib.np_emul_sepplane(C.c_int(int(is_vf)), C.c_longlong(len(st)), st.ctypes.data_as(C.c_void_p), eta.ctypes.data_as(C.c_void_p),
                               hoff.ctypes.data_as(C.c_void_p), htime.ctypes.data_as(C.c_void_p), hpos.ctypes.data_as(C.c_void_p), C.c_double(eps),
                               hit.ctypes.data_as(C.c_void_p))
    assert bad == 0
    return hit


@pytest.mark.skipif(not bind.have_ref(), reason="needs oracle/_ref (the unmodified reference; built where /root/reference exists)")
@pytest.mark.parametrize("name,allowed", [("alec_prob3_402", 0), ("alec_prob11_835", 0), ("alec_prob18_834", 0), ("history_prob3_402", 0)])
def test_sepplane_emul_matches_reference(emul, name, allowed):
    """SeparatingPlaneNarrowPhase (ccd_sepplane.cuh, host build) against the UNMODIFIED reference class on the golden scenes.
    Measured: 0 mismatches on every scene (rpoly and the new root isolator agree on all the short intervals these scenes
    produce); prob11's reference counts are 840 VF / 2,212 EE."""
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    ref = bind.Ref()
    if "hoff" in g.files:
        H = (g["hoff"], g["htime"], g["hpos"])
    else:
        H = bind.single_step_history(g["q0"], g["q1"])
    eta = float(g["eta"])
    vf, ee = g["ref_vf"], g["ref_ee"]
    r = ref.narrowphase(*H, vf, eta, ee, eta, which=1)
    mism = 0
    for nm, is_vf, st in (("vf", True, vf), ("ee", False, ee)):
        hit = _sepplane_emul(emul, is_vf, st, eta, H)
        mism += int((hit != r[nm + "_hit"]).sum())
    assert mism <= allowed, mism
    if name == "alec_prob11_835":
        assert int(r["vf_hit"].sum()) == 840 and int(r["ee_hit"].sum()) == 2212


@pytest.mark.parametrize("variant", ["static", "reverse", "scaled1e3", "half-static", "big-eta"])
def test_emul_matches_checker_on_degenerate_variants(emul, port, variant):
    """Shapes the staged pipeline has special paths for: no motion at all (every polynomial degenerates to a constant), half of
    the vertices static (exactly-zero leading coefficients, closed-form records), coordinates scaled by 1e3, the step run
    backwards, and a thickness of 5e-2 (nearly everything is deferred, many hits) — flags, TOI bits and stage must still
    equal the checker's on every candidate of prob11."""
    g = np.load(os.path.join(HERE, "golden", "alec_prob11_835.npz"))
    q0, q1, vf, ee = g["q0"], g["q1"], g["ref_vf"], g["ref_ee"]
    even = (np.arange(len(q0)) % 2 == 0)[:, None]
    a, b, eta = {"static": (q0, q0.copy(), 1e-3), "reverse": (q1, q0, 1e-8), "scaled1e3": (q0 * 1e3, q1 * 1e3, 1e-5),
                 "half-static": (q0, np.where(even, q0, q1), 1e-6), "big-eta": (q0, q1, 5e-2)}[variant]
    check_scene(emul, port, a, b, vf, ee, eta)


@pytest.mark.parametrize("p", [0, 2, 4])
def test_sepplane_emul_on_velocityfilter_passes(emul, p):
    """The host build of ccd_sepplane.cuh on the growing Histories of the reference's own VelocityFilter run (config C4,
    tests/golden/velocityfilter.npz): per-stencil thickness, up to ~10 entries per vertex, eps = minimum gap / 4."""
    g = np.load(os.path.join(HERE, "golden", "velocityfilter.npz"))
    k = "mesh12_p%d_" % p
    H = (g[k + "hoff"], g[k + "htime"], g[k + "hpos"])
    for nm, is_vf in (("vf", True), ("ee", False)):
        hit = _sepplane_emul(emul, is_vf, g[k + nm], g[k + nm + "_eta"], H)
        assert np.array_equal(hit, g[k + nm + "_hit"]), nm
