"""ctypes binding of the C ABI (include/ccd_b200.h) — the Python face of libccd_b200.so.

Mirrors the reference's plugin interface for the CCD path with the same names and argument meaning:
``BroadPhase.findCollisionCandidates`` (src/RetrospectiveDetection.h:10-15),
``NarrowPhase.findCollisions`` (:17-23), the ``CTCD`` statics (include/CTCD.h:36-79) and the
``Distance`` statics (include/Distance.h:14-180).  Sets become sorted (n,4) int32 arrays.

There is no CPU fallback: if the shared library or a CUDA device is missing, construction raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libccd_b200.so")

KDOP = 13
AABB = 3

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_bp = C.POINTER(C.c_uint8)


class CcdError(RuntimeError):
    pass


class NpSummary(C.Structure):
    _fields_ = [("n_vf_hits", C.c_int64), ("n_ee_hits", C.c_int64), ("earliest_toi", C.c_double)]


class StepResult(C.Structure):
    _fields_ = [("n_vf_candidates", C.c_int64), ("n_ee_candidates", C.c_int64), ("n_vf_hits", C.c_int64),
                ("n_ee_hits", C.c_int64), ("earliest_toi", C.c_double), ("vf_hits", _ip), ("vf_hit_toi", _dp),
                ("ee_hits", _ip), ("ee_hit_toi", _dp), ("ms_broadphase", C.c_float), ("ms_narrowphase", C.c_float)]


class DeviceResult(C.Structure):
    _fields_ = [("n_vf_candidates", C.c_int64), ("n_ee_candidates", C.c_int64), ("n_vf_hits", C.c_int64),
                ("n_ee_hits", C.c_int64), ("earliest_toi", C.c_double), ("d_vf", C.c_void_p), ("d_ee", C.c_void_p),
                ("d_vf_hit", C.c_void_p), ("d_ee_hit", C.c_void_p), ("d_vf_toi", C.c_void_p), ("d_ee_toi", C.c_void_p),
                ("n_face_pairs", C.c_int64), ("n_tree_candidates", C.c_int64), ("ms_broadphase", C.c_float),
                ("ms_narrowphase", C.c_float), ("n_launches", C.c_int32), ("n_vf_deferred", C.c_int64), ("n_ee_deferred", C.c_int64)]


EXPORTS = [
    "ccd_create", "ccd_destroy", "ccd_last_error", "ccd_free_host", "ccd_version", "ccd_broadphase",
    "ccd_broadphase_step", "ccd_narrowphase", "ccd_step", "ccd_step_result_free", "ccd_step_device", "ccd_vf_batch",
    "ccd_ee_batch", "ccd_ve_batch", "ccd_vv_batch", "ccd_find_intervals_batch", "ccd_dist_vf_batch",
    "ccd_dist_ee_batch", "ccd_dist_plane_lt_batch", "ccd_dist_line_lt_batch", "ccd_mesh_self_distance",
    "ccd_memcpy_d2h", "ccd_fp64_peak", "ccd_stage_times", "ccd_step_shard", "ccd_set_shard_partition", "ccd_shard_histogram", "ccd_step_device_hits", "ccd_wait_stream", "ccd_narrowphase_sepplane",
    "ccd_penalty_group_force", "ccd_penalty_group_force_device",
]

_LIB = None


def load_library():
    """Load libccd_b200.so (built by __graft_entry__.build / csrc/Makefile).  Fails loudly if absent."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise CcdError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.ccd_last_error.restype = C.c_char_p
        lib.ccd_last_error.argtypes = [C.c_void_p]
        lib.ccd_version.restype = C.c_char_p
        lib.ccd_free_host.argtypes = [C.c_void_p]
        lib.ccd_free_host.restype = None
        lib.ccd_destroy.argtypes = [C.c_void_p]
        lib.ccd_destroy.restype = None
        lib.ccd_step_result_free.argtypes = [C.POINTER(StepResult)]
        lib.ccd_step_result_free.restype = None
        _LIB = lib
    return _LIB


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(t)


def single_step_history(q0, q1):
    """History(q0) + finishHistory(q1) as CSR (src/History.cpp:8-39)."""
    q0 = _f64(q0).reshape(-1, 3)
    q1 = _f64(q1).reshape(-1, 3)
    V = q0.shape[0]
    hoff = np.arange(0, 2 * V + 1, 2, dtype=np.int64)
    htime = np.tile(np.array([0.0, 1.0]), V)
    hpos = np.empty((2 * V, 3))
    hpos[0::2] = q0
    hpos[1::2] = q1
    return hoff, htime, np.ascontiguousarray(hpos.reshape(-1))


class Context(object):
    """One per process / GPU (ccd_create)."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.ccd_create(C.byref(self.h), C.c_int(device))
        if rc != 0:
            raise CcdError("ccd_create(device=%d) failed with code %d (no CUDA device? there is no CPU fallback)" % (device, rc))

    def close(self):
        if self.h:
            self.lib.ccd_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise CcdError("%s failed (%d): %s" % (what, rc, self.lib.ccd_last_error(self.h).decode()))

    def _take(self, ptr, n):
        out = np.ctypeslib.as_array(ptr, shape=(max(n, 1), 4))[:n].copy() if n > 0 else np.zeros((0, 4), np.int32)
        self.lib.ccd_free_host(ptr)
        return out

    # ---- BroadPhase::findCollisionCandidates --------------------------------------------------
    def findCollisionCandidates(self, kind, faces, hoff, htime, hpos, outerEta, fixedMask=None):
        faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1)
        hoff = np.ascontiguousarray(hoff, dtype=np.int64)
        htime = _f64(htime)
        hpos = _f64(hpos).reshape(-1)
        V, F = hoff.size - 1, faces.size // 3
        fm = None if fixedMask is None else np.ascontiguousarray(fixedMask, dtype=np.uint8)
        vf, ee, nvf, nee = _ip(), _ip(), C.c_int64(), C.c_int64()
        rc = self.lib.ccd_broadphase(self.h, C.c_int(kind), C.c_int(V), C.c_int(F), _ptr(faces, _ip), _ptr(hoff, _lp),
                                     _ptr(htime, _dp), _ptr(hpos, _dp), C.c_double(outerEta), _ptr(fm, _bp),
                                     C.byref(vf), C.byref(nvf), C.byref(ee), C.byref(nee))
        self._check(rc, "ccd_broadphase")
        return self._take(vf, nvf.value), self._take(ee, nee.value)

    def findCollisionCandidatesStep(self, kind, faces, q0, q1, outerEta, fixedMask=None):
        faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1)
        q0 = _f64(q0).reshape(-1)
        q1 = _f64(q1).reshape(-1)
        V, F = q0.size // 3, faces.size // 3
        fm = None if fixedMask is None else np.ascontiguousarray(fixedMask, dtype=np.uint8)
        vf, ee, nvf, nee = _ip(), _ip(), C.c_int64(), C.c_int64()
        rc = self.lib.ccd_broadphase_step(self.h, C.c_int(kind), C.c_int(V), C.c_int(F), _ptr(faces, _ip), _ptr(q0, _dp),
                                          _ptr(q1, _dp), C.c_double(outerEta), _ptr(fm, _bp), C.byref(vf), C.byref(nvf),
                                          C.byref(ee), C.byref(nee))
        self._check(rc, "ccd_broadphase_step")
        return self._take(vf, nvf.value), self._take(ee, nee.value)

    # ---- NarrowPhase::findCollisions (CTCDNarrowPhase) ----------------------------------------
    def findCollisions(self, hoff, htime, hpos, vf, vf_eta, ee, ee_eta):
        hoff = np.ascontiguousarray(hoff, dtype=np.int64)
        htime = _f64(htime)
        hpos = _f64(hpos).reshape(-1)
        vf = np.ascontiguousarray(vf, dtype=np.int32).reshape(-1, 4)
        ee = np.ascontiguousarray(ee, dtype=np.int32).reshape(-1, 4)
        nvf, nee = vf.shape[0], ee.shape[0]
        vf_eta = np.ascontiguousarray(np.broadcast_to(np.asarray(vf_eta, dtype=np.float64), (nvf,)))
        ee_eta = np.ascontiguousarray(np.broadcast_to(np.asarray(ee_eta, dtype=np.float64), (nee,)))
        out = dict(vf_hit=np.zeros(nvf, np.uint8), vf_toi=np.zeros(nvf), vf_stage=np.zeros(nvf, np.uint8),
                   ee_hit=np.zeros(nee, np.uint8), ee_toi=np.zeros(nee), ee_stage=np.zeros(nee, np.uint8))
        s = NpSummary()
        rc = self.lib.ccd_narrowphase(self.h, C.c_int(hoff.size - 1), _ptr(hoff, _lp), _ptr(htime, _dp), _ptr(hpos, _dp),
                                      C.c_int64(nvf), _ptr(vf, _ip), _ptr(vf_eta, _dp), C.c_int64(nee), _ptr(ee, _ip),
                                      _ptr(ee_eta, _dp), _ptr(out["vf_hit"], _bp), _ptr(out["vf_toi"], _dp),
                                      _ptr(out["vf_stage"], _bp), _ptr(out["ee_hit"], _bp), _ptr(out["ee_toi"], _dp),
                                      _ptr(out["ee_stage"], _bp), C.byref(s))
        self._check(rc, "ccd_narrowphase")
        out.update(n_vf_hits=s.n_vf_hits, n_ee_hits=s.n_ee_hits, earliest_toi=s.earliest_toi)
        return out

    # ---- whole step (example/AlecTest.cpp:86-111) ---------------------------------------------
    STAGES = ("topology", "leaf_boxes", "tree_build", "traverse_exact", "adjacency", "emit_count", "emit_write", "np_ee", "np_vf")

    def findCollisionsSeparatingPlane(self, hoff, htime, hpos, vf, vf_eta, ee, ee_eta):
        """SeparatingPlaneNarrowPhase::findCollisions (ccd_narrowphase_sepplane): hit flags per candidate."""
        hoff = np.ascontiguousarray(hoff, dtype=np.int64)
        htime = _f64(htime)
        hpos = _f64(hpos).reshape(-1)
        V = hoff.size - 1
        vf = np.ascontiguousarray(vf, dtype=np.int32).reshape(-1, 4)
        ee = np.ascontiguousarray(ee, dtype=np.int32).reshape(-1, 4)
        nvf, nee = len(vf), len(ee)
        vf_eta = np.ascontiguousarray(np.broadcast_to(np.asarray(vf_eta, dtype=np.float64), (nvf,)))
        ee_eta = np.ascontiguousarray(np.broadcast_to(np.asarray(ee_eta, dtype=np.float64), (nee,)))
        vf_hit = np.zeros(nvf, np.uint8)
        ee_hit = np.zeros(nee, np.uint8)
        nh = (C.c_int64(), C.c_int64())
        rc = self.lib.ccd_narrowphase_sepplane(self.h, C.c_int(V), _ptr(hoff, _lp), _ptr(htime, _dp), _ptr(hpos, _dp), C.c_int64(nvf), _ptr(vf, _ip),
                                               _ptr(vf_eta, _dp), C.c_int64(nee), _ptr(ee, _ip), _ptr(ee_eta, _dp), _ptr(vf_hit, _bp), _ptr(ee_hit, _bp),
                                               C.byref(nh[0]), C.byref(nh[1]))
        self._check(rc, "ccd_narrowphase_sepplane")
        return dict(vf_hit=vf_hit, ee_hit=ee_hit, n_vf_hits=nh[0].value, n_ee_hits=nh[1].value)

    def stage_times(self):
        """Device ms per stage of the last step (ccd_stage_times)."""
        ms = (C.c_float * 9)()
        self._check(self.lib.ccd_stage_times(self.h, ms, C.c_int(9)), "ccd_stage_times")
        return dict(zip(self.STAGES, [float(x) for x in ms]))

    def step(self, kind, faces, q0, q1, outerEta, eta, fixedMask=None, shard_rank=0, shard_world=1, copy=True):
        """ccd_step_shard.  The hit lists come back in pinned host buffers owned by the context (valid until its next
        call): copy=False returns views of them — what a C caller gets — instead of fresh numpy arrays."""
        faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1)
        q0 = _f64(q0).reshape(-1)
        q1 = _f64(q1).reshape(-1)
        V, F = q0.size // 3, faces.size // 3
        fm = None if fixedMask is None else np.ascontiguousarray(fixedMask, dtype=np.uint8)
        r = StepResult()
        rc = self.lib.ccd_step_shard(self.h, C.c_int(kind), C.c_int(V), C.c_int(F), _ptr(faces, _ip), _ptr(q0, _dp), _ptr(q1, _dp),
                                     C.c_double(outerEta), C.c_double(eta), _ptr(fm, _bp), C.c_int(shard_rank), C.c_int(shard_world),
                                     C.byref(r))
        self._check(rc, "ccd_step_shard")

        def arr(p, n, w, dt):
            if n == 0:
                return np.zeros((0, w) if w > 1 else (0,), dt)
            a = np.ctypeslib.as_array(p, shape=(n * w,))
            if copy:
                a = a.copy()
            return a.reshape(n, w) if w > 1 else a

        out = dict(n_vf_candidates=r.n_vf_candidates, n_ee_candidates=r.n_ee_candidates, n_vf_hits=r.n_vf_hits,
                   n_ee_hits=r.n_ee_hits, earliest_toi=r.earliest_toi, ms_broadphase=r.ms_broadphase,
                   ms_narrowphase=r.ms_narrowphase,
                   vf_hits=arr(r.vf_hits, r.n_vf_hits, 4, np.int32), vf_hit_toi=arr(r.vf_hit_toi, r.n_vf_hits, 1, np.float64),
                   ee_hits=arr(r.ee_hits, r.n_ee_hits, 4, np.int32), ee_hit_toi=arr(r.ee_hit_toi, r.n_ee_hits, 1, np.float64))
        self.lib.ccd_step_result_free(C.byref(r))
        return out

    def step_device(self, kind, V, F, d_faces, d_q0, d_q1, outerEta, eta, d_fixed=0, shard_rank=0, shard_world=1):
        """Device-resident step; d_* are raw device addresses (e.g. torch tensor .data_ptr())."""
        r = DeviceResult()
        rc = self.lib.ccd_step_device(self.h, C.c_int(kind), C.c_int(V), C.c_int(F), C.c_void_p(d_faces), C.c_void_p(d_q0),
                                      C.c_void_p(d_q1), C.c_double(outerEta), C.c_double(eta),
                                      C.c_void_p(d_fixed) if d_fixed else None, C.c_int(shard_rank), C.c_int(shard_world),
                                      C.byref(r))
        self._check(rc, "ccd_step_device")
        return r

    def wait_stream(self, cuda_stream):
        """ccd_wait_stream: order the context's next calls after what is enqueued on `cuda_stream` (a raw cudaStream_t
        address, e.g. torch.cuda.current_stream().cuda_stream)."""
        self._check(self.lib.ccd_wait_stream(self.h, C.c_void_p(cuda_stream)), "ccd_wait_stream")

    def step_device_hits(self, kind, V, F, d_faces, d_q0, d_q1, outerEta, eta, d_fixed=0, shard_rank=0, shard_world=1):
        """ccd_step_device_hits: device inputs, hit lists in the context's pinned host buffers (views, valid until the next call)."""
        r = StepResult()
        rc = self.lib.ccd_step_device_hits(self.h, C.c_int(kind), C.c_int(V), C.c_int(F), C.c_void_p(d_faces), C.c_void_p(d_q0), C.c_void_p(d_q1),
                                           C.c_double(outerEta), C.c_double(eta), C.c_void_p(d_fixed) if d_fixed else None,
                                           C.c_int(shard_rank), C.c_int(shard_world), C.byref(r))
        self._check(rc, "ccd_step_device_hits")

        def arr(p, n, w, dt):
            if n == 0:
                return np.zeros((0, w) if w > 1 else (0,), dt)
            a = np.ctypeslib.as_array(p, shape=(n * w,))
            return a.reshape(n, w) if w > 1 else a

        return dict(n_vf_candidates=r.n_vf_candidates, n_ee_candidates=r.n_ee_candidates, n_vf_hits=r.n_vf_hits, n_ee_hits=r.n_ee_hits,
                    earliest_toi=r.earliest_toi, vf_hits=arr(r.vf_hits, r.n_vf_hits, 4, np.int32), vf_hit_toi=arr(r.vf_hit_toi, r.n_vf_hits, 1, np.float64),
                    ee_hits=arr(r.ee_hits, r.n_ee_hits, 4, np.int32), ee_hit_toi=arr(r.ee_hit_toi, r.n_ee_hits, 1, np.float64))

    SHARD_BUCKETS = 1024

    def shard_histogram(self):
        """Load profile of the last sharded step (ccd_shard_histogram): (vf_hist, ee_hist, n_positions)."""
        vf = np.zeros(self.SHARD_BUCKETS, np.int64)
        ee = np.zeros(self.SHARD_BUCKETS, np.int64)
        npos = C.c_int32()
        self._check(self.lib.ccd_shard_histogram(self.h, vf.ctypes.data_as(C.c_void_p), ee.ctypes.data_as(C.c_void_p), C.byref(npos)),
                    "ccd_shard_histogram")
        return vf, ee, npos.value

    def set_shard_partition(self, pbounds):
        """Ownership ranges of the next sharded steps (ccd_set_shard_partition): world+1 ascending bounds over the faces'
        sorted (Morton) positions."""
        pb = np.ascontiguousarray(pbounds, dtype=np.int32)
        self._check(self.lib.ccd_set_shard_partition(self.h, C.c_int(len(pb) - 1), pb.ctypes.data_as(C.c_void_p)), "ccd_set_shard_partition")

    def download(self, d_ptr, shape, dtype):
        """Copy a device array owned by this context (e.g. DeviceResult.d_vf) to a new numpy array."""
        out = np.empty(shape, dtype)
        if out.nbytes:
            self._check(self.lib.ccd_memcpy_d2h(self.h, out.ctypes.data_as(C.c_void_p), C.c_void_p(d_ptr), C.c_uint64(out.nbytes)), "ccd_memcpy_d2h")
        return out

    def fp64_peak_tflops(self):
        t = C.c_double()
        self._check(self.lib.ccd_fp64_peak(self.h, C.byref(t)), "ccd_fp64_peak")
        return t.value

    # ---- CTCD statics, batched ----------------------------------------------------------------
    def _prim(self, fn, npts, pts, eta):
        pts = _f64(pts).reshape(-1, npts * 3)
        n = pts.shape[0]
        eta = np.ascontiguousarray(np.broadcast_to(np.asarray(eta, dtype=np.float64), (n,)))
        hit = np.zeros(n, np.uint8)
        t = np.zeros(n)
        self._check(fn(self.h, C.c_int64(n), _ptr(pts, _dp), _ptr(eta, _dp), _ptr(hit, _bp), _ptr(t, _dp)), "ccd_*_batch")
        return hit, t

    def vertexFaceCTCD(self, pts, eta):
        return self._prim(self.lib.ccd_vf_batch, 8, pts, eta)

    def edgeEdgeCTCD(self, pts, eta):
        return self._prim(self.lib.ccd_ee_batch, 8, pts, eta)

    def vertexEdgeCTCD(self, pts, eta):
        return self._prim(self.lib.ccd_ve_batch, 6, pts, eta)

    def vertexVertexCTCD(self, pts, eta):
        return self._prim(self.lib.ccd_vv_batch, 4, pts, eta)

    def findIntervals(self, coeffs, degree, pos):
        coeffs = _f64(coeffs)
        n = coeffs.shape[0]
        c7 = np.zeros((n, 7))
        c7[:, : degree + 1] = coeffs[:, : degree + 1]
        cnt = np.zeros(n, np.int32)
        lo = np.zeros((n, 7))
        hi = np.zeros((n, 7))
        self._check(self.lib.ccd_find_intervals_batch(self.h, C.c_int64(n), C.c_int(degree), C.c_int(1 if pos else 0),
                                                      _ptr(c7, _dp), _ptr(cnt, _ip), _ptr(lo, _dp), _ptr(hi, _dp)),
                    "ccd_find_intervals_batch")
        return cnt, lo, hi

    # ---- Distance statics, batched ------------------------------------------------------------
    def vertexFaceDistance(self, pts):
        pts = _f64(pts).reshape(-1, 12)
        n = pts.shape[0]
        vec, bary = np.zeros((n, 3)), np.zeros((n, 3))
        self._check(self.lib.ccd_dist_vf_batch(self.h, C.c_int64(n), _ptr(pts, _dp), _ptr(vec, _dp), _ptr(bary, _dp)), "ccd_dist_vf_batch")
        return vec, bary

    def edgeEdgeDistance(self, pts):
        pts = _f64(pts).reshape(-1, 12)
        n = pts.shape[0]
        vec, bary = np.zeros((n, 3)), np.zeros((n, 4))
        self._check(self.lib.ccd_dist_ee_batch(self.h, C.c_int64(n), _ptr(pts, _dp), _ptr(vec, _dp), _ptr(bary, _dp)), "ccd_dist_ee_batch")
        return vec, bary

    def _lt(self, fn, pts, eta):
        pts = _f64(pts).reshape(-1, 12)
        n = pts.shape[0]
        eta = np.ascontiguousarray(np.broadcast_to(np.asarray(eta, dtype=np.float64), (n,)))
        out = np.zeros(n, np.uint8)
        self._check(fn(self.h, C.c_int64(n), _ptr(pts, _dp), _ptr(eta, _dp), _ptr(out, _bp)), "ccd_dist_*_lt_batch")
        return out

    def vertexPlaneDistanceLessThan(self, pts, eta):
        return self._lt(self.lib.ccd_dist_plane_lt_batch, pts, eta)

    def lineLineDistanceLessThan(self, pts, eta):
        return self._lt(self.lib.ccd_dist_line_lt_batch, pts, eta)

    def penaltyGroupAddForce(self, q, v, vf, ee, dt, outerEta, innerEta, stiffness, CoR, F, vf_isnew=None, ee_isnew=None):
        """PenaltyGroup::addForce (src/PenaltyGroup.cpp:34-52) over the group's stencil lists: returns (F + dt * group force,
        fired flags of the vertex-face then the edge-edge stencils, newused)."""
        q = _f64(q).reshape(-1)
        v = _f64(v).reshape(-1)
        F = np.array(F, dtype=np.float64).reshape(-1)
        vf = np.ascontiguousarray(vf, dtype=np.int32).reshape(-1, 4)
        ee = np.ascontiguousarray(ee, dtype=np.int32).reshape(-1, 4)
        vn = None if vf_isnew is None else np.ascontiguousarray(vf_isnew, dtype=np.uint8)
        en = None if ee_isnew is None else np.ascontiguousarray(ee_isnew, dtype=np.uint8)
        vfired = np.zeros(len(vf), dtype=np.uint8)
        efired = np.zeros(len(ee), dtype=np.uint8)
        nf, nu = C.c_int64(), C.c_int()
        self._check(self.lib.ccd_penalty_group_force(
            self.h, C.c_int(q.size // 3), _ptr(q, _dp), _ptr(v, _dp), C.c_int64(len(vf)), _ptr(vf, _ip), _ptr(vn, _bp), C.c_int64(len(ee)),
            _ptr(ee, _ip), _ptr(en, _bp), C.c_double(dt), C.c_double(outerEta), C.c_double(innerEta), C.c_double(stiffness), C.c_double(CoR),
            _ptr(F, _dp), _ptr(vfired, _bp), _ptr(efired, _bp), C.byref(nf), C.byref(nu)), "ccd_penalty_group_force")
        return F, np.concatenate([vfired, efired]), bool(nu.value), int(nf.value)

    def meshSelfDistance(self, verts, faces, fixedMask=None):
        verts = _f64(verts).reshape(-1)
        faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1)
        fm = None if fixedMask is None else np.ascontiguousarray(fixedMask, dtype=np.uint8)
        d = C.c_double()
        nvf, nee = C.c_int64(), C.c_int64()
        self._check(self.lib.ccd_mesh_self_distance(self.h, C.c_int(verts.size // 3), _ptr(verts, _dp), C.c_int(faces.size // 3),
                                                    _ptr(faces, _ip), _ptr(fm, _bp), C.byref(d), C.byref(nvf), C.byref(nee)),
                    "ccd_mesh_self_distance")
        return d.value, nvf.value, nee.value
