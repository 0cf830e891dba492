"""TEST INFRASTRUCTURE — ctypes bindings for the two CPU checkers.

* ``Ref``    : oracle/_ref/libccdref.so — the UNMODIFIED reference compiled against the Eigen
               shim (oracle/Makefile, oracle/ref_harness.cpp).
* ``Port``   : oracle/libccd_oracle.so — the plain-C restatement (oracle/ccd_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (collisiondetection_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libccdref.so")
PORT_SO = os.path.join(HERE, "libccd_oracle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_longlong)
_bp = C.POINTER(C.c_ubyte)


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _l(a):
    return a.ctypes.data_as(_lp) if a is not None else None


def _b(a):
    return a.ctypes.data_as(_bp) if a is not None else None


def build(ref=True, port=True):
    """Build the checkers (oracle/Makefile).  `ref` is skipped when /root/reference is absent."""
    targets = []
    if port:
        targets.append("oracle")
    if ref:
        targets.append("ref")
    if targets:
        subprocess.check_call(["make", "-s", "-C", HERE] + targets)


def single_step_history(q0, q1):
    """CSR history for the 2-entries-per-vertex case (src/History.cpp:8-39)."""
    q0 = np.ascontiguousarray(q0, dtype=np.float64).reshape(-1, 3)
    q1 = np.ascontiguousarray(q1, dtype=np.float64).reshape(-1, 3)
    V = q0.shape[0]
    hoff = np.arange(0, 2 * V + 1, 2, dtype=np.int64)
    htime = np.tile(np.array([0.0, 1.0]), V)
    hpos = np.empty((2 * V, 3))
    hpos[0::2] = q0
    hpos[1::2] = q1
    return hoff, htime, np.ascontiguousarray(hpos.reshape(-1))


def fnv1a64(stencils):
    """FNV-1a-64 over sorted stencils, each 4 little-endian int32."""
    data = np.ascontiguousarray(stencils, dtype="<i4").tobytes()
    h = 0xCBF29CE484222325
    # chunked pure-python loop is fine for the small golden sets; numpy path for big ones
    if len(data) > (1 << 16):
        return _fnv_big(data)
    for b in data:
        h ^= b
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def _fnv_big(data):
    lib = _fnvlib()
    return "%016x" % lib.fnv1a64(data, C.c_size_t(len(data)))


_FNV = None


def _fnvlib():
    global _FNV
    if _FNV is None:
        src = os.path.join(HERE, "fnv.c")
        so = os.path.join(HERE, "libfnv.so")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src])
        _FNV = C.CDLL(so)
        _FNV.fnv1a64.restype = C.c_uint64
        _FNV.fnv1a64.argtypes = [C.c_char_p, C.c_size_t]
    return _FNV


class _Lib(object):
    prefix = ""
    path = ""

    def __init__(self):
        if not os.path.exists(self.path):
            raise FileNotFoundError(self.path + " not built (run `make -C oracle`)")
        self.lib = C.CDLL(self.path)
        self.lib[self.prefix + "_free"].argtypes = [C.c_void_p]
        self.lib[self.prefix + "_free"].restype = None

    def f(self, name, restype=None):
        fn = self.lib[self.prefix + "_" + name]
        fn.restype = restype
        return fn

    # -- broadphase ----------------------------------------------------------------------
    def broadphase(self, kind, faces, hoff, htime, hpos, outer_eta, fixed=None):
        faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1)
        F = faces.size // 3
        V = hoff.size - 1
        vf = C.POINTER(C.c_int)()
        ee = C.POINTER(C.c_int)()
        nvf = C.c_longlong()
        nee = C.c_longlong()
        secs = C.c_double()
        fm = None if fixed is None else np.ascontiguousarray(fixed, dtype=np.uint8)
        rc = self.f("broadphase", C.c_int)(
            C.c_int(kind), C.c_int(V), C.c_int(F), _i(faces), _l(hoff), _d(htime), _d(hpos),
            C.c_double(outer_eta), _b(fm), C.byref(vf), C.byref(nvf), C.byref(ee), C.byref(nee), C.byref(secs))
        assert rc == 0
        a = np.ctypeslib.as_array(vf, shape=(max(nvf.value, 1), 4))[: nvf.value].copy()
        b = np.ctypeslib.as_array(ee, shape=(max(nee.value, 1), 4))[: nee.value].copy()
        self.lib[self.prefix + "_free"](vf)
        self.lib[self.prefix + "_free"](ee)
        return a, b, secs.value

    # -- narrowphase ---------------------------------------------------------------------
    def narrowphase(self, hoff, htime, hpos, vf, vf_eta, ee, ee_eta, which=0):
        V = hoff.size - 1
        vf = np.ascontiguousarray(vf, dtype=np.int32).reshape(-1, 4)
        ee = np.ascontiguousarray(ee, dtype=np.int32).reshape(-1, 4)
        nvf, nee = vf.shape[0], ee.shape[0]
        vf_eta = np.ascontiguousarray(np.broadcast_to(np.asarray(vf_eta, dtype=np.float64), (nvf,)))
        ee_eta = np.ascontiguousarray(np.broadcast_to(np.asarray(ee_eta, dtype=np.float64), (nee,)))
        out = dict(
            vf_hit=np.zeros(nvf, np.uint8), vf_toi=np.zeros(nvf), vf_stage=np.zeros(nvf, np.int32),
            ee_hit=np.zeros(nee, np.uint8), ee_toi=np.zeros(nee), ee_stage=np.zeros(nee, np.int32))
        secs = C.c_double()
        rc = self.f("narrowphase", C.c_int)(
            C.c_int(which), C.c_int(V), _l(hoff), _d(htime), _d(hpos), C.c_longlong(nvf), _i(vf), _d(vf_eta),
            C.c_longlong(nee), _i(ee), _d(ee_eta), _b(out["vf_hit"]), _d(out["vf_toi"]), _i(out["vf_stage"]),
            _b(out["ee_hit"]), _d(out["ee_toi"]), _i(out["ee_stage"]), C.byref(secs))
        out["disagree"] = rc
        out["seconds"] = secs.value
        return out

    def narrowphase_flat(self, hoff, htime, hpos, vf, vf_eta, ee, ee_eta):
        V = hoff.size - 1
        vf = np.ascontiguousarray(vf, dtype=np.int32).reshape(-1, 4)
        ee = np.ascontiguousarray(ee, dtype=np.int32).reshape(-1, 4)
        nvf, nee = vf.shape[0], ee.shape[0]
        vf_eta = np.ascontiguousarray(np.broadcast_to(np.asarray(vf_eta, dtype=np.float64), (nvf,)))
        ee_eta = np.ascontiguousarray(np.broadcast_to(np.asarray(ee_eta, dtype=np.float64), (nee,)))
        out = dict(vf_hit=np.zeros(nvf, np.uint8), vf_toi=np.zeros(nvf),
                   ee_hit=np.zeros(nee, np.uint8), ee_toi=np.zeros(nee))
        self.f("narrowphase_flat")(
            C.c_int(V), _l(hoff), _d(htime), _d(hpos), C.c_longlong(nvf), _i(vf), _d(vf_eta),
            C.c_longlong(nee), _i(ee), _d(ee_eta), _b(out["vf_hit"]), _d(out["vf_toi"]),
            _b(out["ee_hit"]), _d(out["ee_toi"]))
        return out

    # -- primitives ----------------------------------------------------------------------
    def _prim(self, name, npts, pts, eta):
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, npts * 3)
        n = pts.shape[0]
        eta = np.ascontiguousarray(np.broadcast_to(np.asarray(eta, dtype=np.float64), (n,)))
        hit = np.zeros(n, np.uint8)
        t = np.zeros(n)
        self.f(name)(C.c_longlong(n), _d(pts), _d(eta), _b(hit), _d(t))
        return hit, t

    def vf_batch(self, pts, eta):
        return self._prim("vf_batch", 8, pts, eta)

    def ee_batch(self, pts, eta):
        return self._prim("ee_batch", 8, pts, eta)

    def ve_batch(self, pts, eta):
        return self._prim("ve_batch", 6, pts, eta)

    def vv_batch(self, pts, eta):
        return self._prim("vv_batch", 4, pts, eta)

    # -- distance ------------------------------------------------------------------------
    def dist_vf_batch(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 12)
        n = pts.shape[0]
        vec = np.zeros((n, 3))
        bary = np.zeros((n, 3))
        self.f("dist_vf_batch")(C.c_longlong(n), _d(pts), _d(vec), _d(bary))
        return vec, bary

    def dist_ee_batch(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 12)
        n = pts.shape[0]
        vec = np.zeros((n, 3))
        bary = np.zeros((n, 4))
        self.f("dist_ee_batch")(C.c_longlong(n), _d(pts), _d(vec), _d(bary))
        return vec, bary

    def _lt(self, name, pts, eta):
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 12)
        n = pts.shape[0]
        eta = np.ascontiguousarray(np.broadcast_to(np.asarray(eta, dtype=np.float64), (n,)))
        out = np.zeros(n, np.uint8)
        self.f(name)(C.c_longlong(n), _d(pts), _d(eta), _b(out))
        return out

    def dist_plane_lt_batch(self, pts, eta):
        return self._lt("dist_plane_lt_batch", pts, eta)

    def dist_line_lt_batch(self, pts, eta):
        return self._lt("dist_line_lt_batch", pts, eta)

    def mesh_self_distance(self, verts, faces, fixed=None):
        verts = np.ascontiguousarray(verts, dtype=np.float64).reshape(-1)
        faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1)
        fm = None if fixed is None else np.ascontiguousarray(fixed, dtype=np.uint8)
        secs = C.c_double()
        d = self.f("mesh_self_distance", C.c_double)(
            C.c_int(verts.size // 3), _d(verts), C.c_int(faces.size // 3), _i(faces), _b(fm), C.byref(secs))
        return d, secs.value


class Ref(_Lib):
    prefix = "ref"
    path = REF_SO

    def rpoly(self, op):
        op = np.ascontiguousarray(op, dtype=np.float64)
        deg = op.size - 1
        zr = np.zeros(8)
        zi = np.zeros(8)
        n = self.f("rpoly", C.c_int)(_d(op), C.c_int(deg), _d(zr), _d(zi))
        return n, zr[:max(n, 0)], zi[:max(n, 0)]

    def stitch(self, hoff, htime, hpos, verts4, cap=4096):
        verts4 = np.ascontiguousarray(verts4, dtype=np.int32)
        times = np.zeros(cap)
        pos = np.zeros((cap, 12))
        n = self.f("stitch", C.c_int)(C.c_int(hoff.size - 1), _l(hoff), _d(htime), _d(hpos), _i(verts4),
                                      C.c_int(cap), _d(times), _d(pos))
        return times[:n], pos[:n]


def _penalty_args(q, v, vf, ee, F):
    q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1)
    v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1)
    vf = np.ascontiguousarray(vf, dtype=np.int32).reshape(-1, 4)
    ee = np.ascontiguousarray(ee, dtype=np.int32).reshape(-1, 4)
    F = np.array(F, dtype=np.float64).reshape(-1)      # copy: in/out
    return q, v, vf, ee, F


class Port(_Lib):
    prefix = "orc"
    path = PORT_SO

    def penalty_group_force(self, q, v, vf, ee, dt, outer_eta, inner_eta, stiffness, cor, F, vf_isnew=None, ee_isnew=None):
        """PenaltyGroup::addForce restated (src/PenaltyGroup.cpp:34-52).  Returns (F, fired flags, newused)."""
        q, v, vf, ee, F = _penalty_args(q, v, vf, ee, F)
        fired = np.zeros(len(vf) + len(ee), dtype=np.uint8)
        vn = None if vf_isnew is None else np.ascontiguousarray(vf_isnew, dtype=np.uint8)
        en = None if ee_isnew is None else np.ascontiguousarray(ee_isnew, dtype=np.uint8)
        nf = C.c_longlong()
        r = self.f("penalty_group_force", C.c_int)(
            C.c_int(q.size // 3), _d(q), _d(v), C.c_longlong(len(vf)), _i(vf), _b(vn), C.c_longlong(len(ee)), _i(ee), _b(en),
            C.c_double(dt), C.c_double(outer_eta), C.c_double(inner_eta), C.c_double(stiffness), C.c_double(cor), _d(F), _b(fired), C.byref(nf))
        return F, fired, bool(r)


VF_SO = os.path.join(HERE, "_ref", "libccdvf.so")


class RefVF(object):
    """oracle/_ref/libccdvf.so: the reference's VelocityFilter / ActiveLayers / PenaltyGroup objects (oracle/ref_recorder.cpp)."""

    def __init__(self):
        if not os.path.exists(VF_SO):
            raise FileNotFoundError(VF_SO + " not built (run `make -C oracle ref`)")
        self.lib = C.CDLL(VF_SO)

    def penalty_group_force(self, q, v, vf, ee, dt, outer_eta, inner_eta, stiffness, cor, F, rolled_back=False):
        """The reference's own PenaltyGroup::addForce over the lists (all stencils new, or none after rollback())."""
        q, v, vf, ee, F = _penalty_args(q, v, vf, ee, F)
        fired = np.zeros(len(vf) + len(ee), dtype=np.uint8)
        fn = self.lib.ref_penalty_group_force
        fn.restype = C.c_int
        r = fn(C.c_int(q.size // 3), _d(q), _d(v), C.c_longlong(len(vf)), _i(vf), C.c_longlong(len(ee)), _i(ee), C.c_int(int(rolled_back)),
               C.c_double(dt), C.c_double(outer_eta), C.c_double(inner_eta), C.c_double(stiffness), C.c_double(cor), _d(F), _b(fired))
        return F, fired, bool(r)


def have_refvf():
    return os.path.exists(VF_SO)


def have_ref():
    return os.path.exists(REF_SO)


def have_port():
    return os.path.exists(PORT_SO)
