"""PenaltyGroup::addForce (src/PenaltyGroup.cpp:34-52, src/PenaltyPotential.cpp:7-64) — SURVEY §8(f) rank 4.

Golden: tests/golden/penalty.npz, computed by the reference's own PenaltyGroup object code (make_golden_penalty.py).
CPU: the plain-C restatement against the golden, bit for bit.  GPU: ccd_penalty_group_force against the golden, bit for bit
(forces are accumulated per vertex in list order, so even the rounding of the sums is the reference's)."""
import numpy as np
import pytest

from conftest import golden

CASES = ["thick", "thick_mid", "prob11"]


def _case(g, name):
    dt, outer, inner, k, cor = [float(x) for x in g[name + "_params"]]
    return (g[name + "_q"], g[name + "_v"], g[name + "_vf"], g[name + "_ee"], dt, outer, inner, k, cor, g[name + "_F0"])


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


@pytest.mark.parametrize("name", CASES)
def test_restatement_matches_reference(port, name):
    g = golden("penalty.npz")
    args = _case(g, name)
    F, fired, newused = port.penalty_group_force(*args)
    assert np.array_equal(fired, g[name + "_fired"])
    assert np.array_equal(_bits(F), _bits(g[name + "_F"]))      # includes -0 + 0 = +0 on untouched coordinates
    assert newused == bool(g[name + "_newused"])
    n_vf = len(args[2])
    # after rollback() no stencil is new: same force, newused false
    F2, _, nu2 = port.penalty_group_force(*args, vf_isnew=np.zeros(n_vf, np.uint8), ee_isnew=np.zeros(len(args[3]), np.uint8))
    assert np.array_equal(_bits(F2), _bits(g[name + "_rb_F"])) and nu2 == bool(g[name + "_rb_newused"]) and not nu2


def test_restatement_matches_reference_live(port):
    """When oracle/_ref travelled: a fresh random group straight against the reference's object code."""
    from oracle import bind
    if not bind.have_refvf():
        pytest.skip("oracle/_ref/libccdvf.so not built")
    ref = bind.RefVF()
    g = golden("penalty.npz")
    q, v, vf, ee = g["thick_q"], g["thick_v"], g["thick_vf"], g["thick_ee"]
    rng = np.random.default_rng(5)
    sel_v = rng.permutation(len(vf))[:3000]      # list order need not be set order
    sel_e = rng.permutation(len(ee))[:5000]
    F0 = rng.standard_normal(q.size)
    a = ref.penalty_group_force(q, v, vf[sel_v], ee[sel_e], 2e-3, 6e-3, 5e-4, 250.0, 0.3, F0)
    b = port.penalty_group_force(q, v, vf[sel_v], ee[sel_e], 2e-3, 6e-3, 5e-4, 250.0, 0.3, F0)
    assert int(a[1].sum()) > 100
    assert np.array_equal(a[1], b[1]) and np.array_equal(_bits(a[0]), _bits(b[0])) and a[2] == b[2]


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_penalty_group_force(ctx, port, name):
    g = golden("penalty.npz")
    args = _case(g, name)
    F, fired, newused, nf = ctx.penaltyGroupAddForce(*args)
    assert np.array_equal(fired, g[name + "_fired"]) and nf == int(g[name + "_fired"].sum())
    assert np.array_equal(_bits(F), _bits(g[name + "_F"]))
    assert newused == bool(g[name + "_newused"])
    # mixed isnew flags and a shuffled list order against the restatement
    rng = np.random.default_rng(11)
    q, v, vf, ee = args[:4]
    pv, pe = rng.permutation(len(vf)), rng.permutation(len(ee))
    vn, en = (rng.random(len(vf)) < 0.01).astype(np.uint8), np.zeros(len(ee), np.uint8)
    a = ctx.penaltyGroupAddForce(q, v, vf[pv], ee[pe], *args[4:], vf_isnew=vn, ee_isnew=en)
    b = port.penalty_group_force(q, v, vf[pv], ee[pe], *args[4:], vf_isnew=vn, ee_isnew=en)
    assert np.array_equal(a[1], b[1]) and np.array_equal(_bits(a[0]), _bits(b[0])) and a[2] == b[2]


@pytest.mark.gpu
def test_gpu_penalty_empty_group(ctx):
    F0 = np.array([1.0, -0.0, 2.0, 0.0, 0.0, 0.0])
    F, fired, newused, nf = ctx.penaltyGroupAddForce(np.zeros(6), np.zeros(6), np.zeros((0, 4), np.int32), np.zeros((0, 4), np.int32), 1e-3, 1e-2, 1e-3, 1.0, 0.5, F0)
    assert nf == 0 and not newused and len(fired) == 0
    assert np.array_equal(_bits(F), _bits(F0 + 0.0))
