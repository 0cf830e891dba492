// C++ drop-in adapters over the C ABI (include/ccd_b200.h) for evouga/collisiondetection.
//
// Compile this header inside the reference tree (it includes the reference's own headers) and swap one type name
// at the call sites of the hot path (or change no source line at all and link examples/dropin_link.cpp in place of the
// reference's detection objects):
//
//   src/ActiveLayers.cpp:26     bp_ = new KDOPBroadPhase();             ->  new ccdgpu::KDOPBroadPhase();
//   src/ActiveLayers.cpp:28     np_ = new SeparatingPlaneNarrowPhase(); ->  new ccdgpu::SeparatingPlaneNarrowPhase();
//                               (ccdgpu::CTCDNarrowPhase has the same interface but is the OTHER algorithm: swapping it in changes
//                               the hit set of the VelocityFilter path, e.g. 2210 instead of 2212 edge-edge hits on prob11)
//   src/Distance.cpp:43         AABBBroadPhase bp;                      ->  ccdgpu::AABBBroadPhase bp;
//   example/AlecTest.cpp:97,111 KDOPBroadPhase() / CTCDNarrowPhase()    ->  ccdgpu::...
//
// and, for the static entry points, `CTCD::vertexFaceCTCD(...)` -> `ccdgpu::CTCD::vertexFaceCTCD(...)`, same for
// edgeEdge/vertexEdge/vertexVertex (include/CTCD.h:36-79) and `Distance::...` (include/Distance.h:14-180).
// Argument meaning, ownership and set semantics are the reference's: the broadphase CLEARS its output sets
// (src/KDOPBroadPhase.cpp:36-37), the narrowphase APPENDS (src/CTCDNarrowPhase.cpp:14,19), `t` is written only on
// a hit.  There is no CPU fallback: if no CUDA device is present the constructors throw.
#ifndef CCD_B200_ADAPTERS_HPP
#define CCD_B200_ADAPTERS_HPP

#include <cstdlib>
#include <iostream>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include <Eigen/Core>
#include "RetrospectiveDetection.h" // BroadPhase, NarrowPhase, Stencils.h, History.h
#include "Mesh.h"

#include "ccd_b200.h"

namespace ccdgpu {

// One context per process, created on first use (device from CCD_B200_DEVICE, default 0).
inline ccd_context *create_context()
{
    ccd_context *ctx = 0;
    int dev = 0;
    if (const char *e = getenv("CCD_B200_DEVICE"))
        dev = atoi(e);
    int rc = ccd_create(&ctx, dev);
    if (rc != CCD_OK)
        throw std::runtime_error("ccd_create failed (" + std::to_string(rc) + "): a CUDA device is required, there is no CPU fallback");
    return ctx;
}
inline ccd_context *context()
{
    static ccd_context *ctx = create_context();      // initialised once, thread-safe (C++11 function-local static)
    return ctx;
}

inline void check(int rc, const char *what)
{
    if (rc != CCD_OK)
        throw std::runtime_error(std::string(what) + " failed: " + ccd_last_error(context()));
}

// History (src/History.h:28-42) -> CSR arrays
struct FlatHistory
{
    std::vector<int64_t> off;
    std::vector<double> time, pos;
    FlatHistory(const History &h, int nverts)
    {
        off.resize(nverts + 1);
        off[0] = 0;
        for (int v = 0; v < nverts; v++)
            off[v + 1] = off[v] + (int64_t)h.getVertexHistory(v).size();
        time.reserve(off[nverts]);
        pos.reserve(3 * off[nverts]);
        for (int v = 0; v < nverts; v++)
        {
            const std::vector<HistoryEntry> &hv = h.getVertexHistory(v);
            for (size_t e = 0; e < hv.size(); e++)
            {
                time.push_back(hv[e].time);
                pos.push_back(hv[e].pos[0]);
                pos.push_back(hv[e].pos[1]);
                pos.push_back(hv[e].pos[2]);
            }
        }
    }
};

class BroadPhaseBase : public ::BroadPhase
{
public:
    explicit BroadPhaseBase(int kind) : kind_(kind) { context(); }

    virtual void findCollisionCandidates(const History &h, const Mesh &m, double outerEta, std::set<VertexFaceStencil> &vfs,
                                         std::set<EdgeEdgeStencil> &ees, const std::set<int> &fixedVerts)
    {
        vfs.clear();
        ees.clear();
        const int V = (int)(m.vertices.size() / 3), F = (int)m.faces.cols();
        FlatHistory fh(h, V);
        std::vector<int32_t> faces(3 * (size_t)F);
        for (int f = 0; f < F; f++)
            for (int j = 0; j < 3; j++)
                faces[3 * (size_t)f + j] = m.faces.coeff(j, f);
        std::vector<uint8_t> fixedMask;
        if (!fixedVerts.empty())
        {
            fixedMask.assign(V, 0);
            for (std::set<int>::const_iterator it = fixedVerts.begin(); it != fixedVerts.end(); ++it)
                if (*it >= 0 && *it < V)
                    fixedMask[*it] = 1;
        }
        int32_t *vf = 0, *ee = 0;
        int64_t nvf = 0, nee = 0;
        check(ccd_broadphase(context(), kind_, V, F, faces.data(), fh.off.data(), fh.time.data(), fh.pos.data(), outerEta,
                             fixedMask.empty() ? 0 : fixedMask.data(), &vf, &nvf, &ee, &nee),
              "ccd_broadphase");
        // arrays arrive in std::set order: hinted insertion at end() is O(1) per element
        for (int64_t i = 0; i < nvf; i++)
            vfs.insert(vfs.end(), VertexFaceStencil(vf[4 * i], vf[4 * i + 1], vf[4 * i + 2], vf[4 * i + 3]));
        for (int64_t i = 0; i < nee; i++)
            ees.insert(ees.end(), EdgeEdgeStencil(ee[4 * i], ee[4 * i + 1], ee[4 * i + 2], ee[4 * i + 3]));
        ccd_free_host(vf);
        ccd_free_host(ee);
    }

private:
    int kind_;
};

class KDOPBroadPhase : public BroadPhaseBase
{
public:
    KDOPBroadPhase() : BroadPhaseBase(CCD_KDOP) {}
};
class AABBBroadPhase : public BroadPhaseBase
{
public:
    AABBBroadPhase() : BroadPhaseBase(CCD_AABB) {}
};

class CTCDNarrowPhase : public ::NarrowPhase
{
public:
    CTCDNarrowPhase() : earliestTOI(1.0 / 0.0) { context(); }

    virtual void findCollisions(const History &h, const std::set<std::pair<VertexFaceStencil, double> > &candidateVFS,
                                const std::set<std::pair<EdgeEdgeStencil, double> > &candidateEES, std::set<VertexFaceStencil> &vfs,
                                std::set<EdgeEdgeStencil> &ees)
    {
        // vertex count: one past the largest index referenced (History does not expose its size)
        int V = 0;
        std::vector<int32_t> vf, ee;
        std::vector<double> vfe, eee;
        vf.reserve(4 * candidateVFS.size());
        ee.reserve(4 * candidateEES.size());
        for (std::set<std::pair<VertexFaceStencil, double> >::const_iterator it = candidateVFS.begin(); it != candidateVFS.end(); ++it)
        {
            const int s[4] = {it->first.p, it->first.q0, it->first.q1, it->first.q2};
            for (int k = 0; k < 4; k++) { vf.push_back(s[k]); if (s[k] + 1 > V) V = s[k] + 1; }
            vfe.push_back(it->second);
        }
        for (std::set<std::pair<EdgeEdgeStencil, double> >::const_iterator it = candidateEES.begin(); it != candidateEES.end(); ++it)
        {
            const int s[4] = {it->first.p0, it->first.p1, it->first.q0, it->first.q1};
            for (int k = 0; k < 4; k++) { ee.push_back(s[k]); if (s[k] + 1 > V) V = s[k] + 1; }
            eee.push_back(it->second);
        }
        FlatHistory fh(h, V);
        const int64_t nvf = (int64_t)vfe.size(), nee = (int64_t)eee.size();
        vfHit.assign(nvf, 0);
        eeHit.assign(nee, 0);
        vfTOI.assign(nvf, 0.0);
        eeTOI.assign(nee, 0.0);
        ccd_np_summary sum;
        check(ccd_narrowphase(context(), V, fh.off.data(), fh.time.data(), fh.pos.data(), nvf, vf.data(), vfe.data(), nee, ee.data(),
                              eee.data(), vfHit.data(), vfTOI.data(), 0, eeHit.data(), eeTOI.data(), 0, &sum),
              "ccd_narrowphase");
        earliestTOI = sum.earliest_toi;
        int64_t i = 0;
        for (std::set<std::pair<VertexFaceStencil, double> >::const_iterator it = candidateVFS.begin(); it != candidateVFS.end(); ++it, ++i)
            if (vfHit[i])
                vfs.insert(it->first);
        i = 0;
        for (std::set<std::pair<EdgeEdgeStencil, double> >::const_iterator it = candidateEES.begin(); it != candidateEES.end(); ++it, ++i)
            if (eeHit[i])
                ees.insert(it->first);
    }

    // what the reference computes and throws away (src/CTCDNarrowPhase.cpp:42-48): per-candidate flags / TOIs of the
    // last call, in the iteration order of the candidate sets, and the earliest TOI over all hits
    std::vector<uint8_t> vfHit, eeHit;
    std::vector<double> vfTOI, eeTOI;
    double earliestTOI;
};

// src/SeparatingPlaneNarrowPhase.h:8-22 — the narrowphase ActiveLayers owns (src/ActiveLayers.cpp:28)
class SeparatingPlaneNarrowPhase : public ::NarrowPhase
{
public:
    SeparatingPlaneNarrowPhase() { context(); }

    virtual void findCollisions(const History &h, const std::set<std::pair<VertexFaceStencil, double> > &candidateVFS,
                                const std::set<std::pair<EdgeEdgeStencil, double> > &candidateEES, std::set<VertexFaceStencil> &vfs,
                                std::set<EdgeEdgeStencil> &ees)
    {
        int V = 0;
        std::vector<int32_t> vf, ee;
        std::vector<double> vfe, eee;
        vf.reserve(4 * candidateVFS.size());
        ee.reserve(4 * candidateEES.size());
        for (std::set<std::pair<VertexFaceStencil, double> >::const_iterator it = candidateVFS.begin(); it != candidateVFS.end(); ++it)
        {
            const int s[4] = {it->first.p, it->first.q0, it->first.q1, it->first.q2};
            for (int k = 0; k < 4; k++) { vf.push_back(s[k]); if (s[k] + 1 > V) V = s[k] + 1; }
            vfe.push_back(it->second);
        }
        for (std::set<std::pair<EdgeEdgeStencil, double> >::const_iterator it = candidateEES.begin(); it != candidateEES.end(); ++it)
        {
            const int s[4] = {it->first.p0, it->first.p1, it->first.q0, it->first.q1};
            for (int k = 0; k < 4; k++) { ee.push_back(s[k]); if (s[k] + 1 > V) V = s[k] + 1; }
            eee.push_back(it->second);
        }
        FlatHistory fh(h, V);
        const int64_t nvf = (int64_t)vfe.size(), nee = (int64_t)eee.size();
        vfHit.assign(nvf, 0);
        eeHit.assign(nee, 0);
        check(ccd_narrowphase_sepplane(context(), V, fh.off.data(), fh.time.data(), fh.pos.data(), nvf, vf.data(), vfe.data(), nee, ee.data(),
                                       eee.data(), vfHit.data(), eeHit.data(), 0, 0),
              "ccd_narrowphase_sepplane");
        int64_t i = 0;
        for (std::set<std::pair<VertexFaceStencil, double> >::const_iterator it = candidateVFS.begin(); it != candidateVFS.end(); ++it, ++i)
            if (vfHit[i])
                vfs.insert(it->first);
        i = 0;
        for (std::set<std::pair<EdgeEdgeStencil, double> >::const_iterator it = candidateEES.begin(); it != candidateEES.end(); ++it, ++i)
            if (eeHit[i])
                ees.insert(it->first);
    }

    std::vector<uint8_t> vfHit, eeHit;      // per-candidate flags of the last call, in the iteration order of the candidate sets
};

// include/CTCD.h:36-79 — a single call is a batch of one on the GPU
struct CTCD
{
    static bool run(int (*fn)(ccd_context *, int64_t, const double *, const double *, uint8_t *, double *), const Eigen::Vector3d *const *p,
                    int np, double eta, double &t)
    {
        double pts[24];
        for (int k = 0; k < np; k++)
            for (int c = 0; c < 3; c++)
                pts[3 * k + c] = (*p[k])[c];
        uint8_t hit = 0;
        double tt = t;
        check(fn(context(), 1, pts, &eta, &hit, &tt), "ccd_*_batch");
        if (hit)
            t = tt;
        return hit != 0;
    }
    static bool edgeEdgeCTCD(const Eigen::Vector3d &q0start, const Eigen::Vector3d &p0start, const Eigen::Vector3d &q1start,
                             const Eigen::Vector3d &p1start, const Eigen::Vector3d &q0end, const Eigen::Vector3d &p0end,
                             const Eigen::Vector3d &q1end, const Eigen::Vector3d &p1end, double eta, double &t)
    {
        const Eigen::Vector3d *p[8] = {&q0start, &p0start, &q1start, &p1start, &q0end, &p0end, &q1end, &p1end};
        return run(ccd_ee_batch, p, 8, eta, t);
    }
    static bool vertexFaceCTCD(const Eigen::Vector3d &q0start, const Eigen::Vector3d &q1start, const Eigen::Vector3d &q2start,
                               const Eigen::Vector3d &q3start, const Eigen::Vector3d &q0end, const Eigen::Vector3d &q1end,
                               const Eigen::Vector3d &q2end, const Eigen::Vector3d &q3end, double eta, double &t)
    {
        const Eigen::Vector3d *p[8] = {&q0start, &q1start, &q2start, &q3start, &q0end, &q1end, &q2end, &q3end};
        return run(ccd_vf_batch, p, 8, eta, t);
    }
    static bool vertexEdgeCTCD(const Eigen::Vector3d &q0start, const Eigen::Vector3d &q1start, const Eigen::Vector3d &q2start,
                               const Eigen::Vector3d &q0end, const Eigen::Vector3d &q1end, const Eigen::Vector3d &q2end, double eta,
                               double &t)
    {
        const Eigen::Vector3d *p[6] = {&q0start, &q1start, &q2start, &q0end, &q1end, &q2end};
        return run(ccd_ve_batch, p, 6, eta, t);
    }
    static bool vertexVertexCTCD(const Eigen::Vector3d &q1start, const Eigen::Vector3d &q2start, const Eigen::Vector3d &q1end,
                                 const Eigen::Vector3d &q2end, double eta, double &t)
    {
        const Eigen::Vector3d *p[4] = {&q1start, &q2start, &q1end, &q2end};
        return run(ccd_vv_batch, p, 4, eta, t);
    }
};

// include/Distance.h:14-180
struct Distance
{
    static void pack(const Eigen::Vector3d &a, const Eigen::Vector3d &b, const Eigen::Vector3d &c, const Eigen::Vector3d &d, double *pts)
    {
        const Eigen::Vector3d *p[4] = {&a, &b, &c, &d};
        for (int k = 0; k < 4; k++)
            for (int j = 0; j < 3; j++)
                pts[3 * k + j] = (*p[k])[j];
    }
    static bool vertexPlaneDistanceLessThan(const Eigen::Vector3d &p, const Eigen::Vector3d &q0, const Eigen::Vector3d &q1,
                                            const Eigen::Vector3d &q2, double eta)
    {
        double pts[12];
        uint8_t out = 0;
        pack(p, q0, q1, q2, pts);
        check(ccd_dist_plane_lt_batch(context(), 1, pts, &eta, &out), "ccd_dist_plane_lt_batch");
        return out != 0;
    }
    static bool lineLineDistanceLessThan(const Eigen::Vector3d &p0, const Eigen::Vector3d &p1, const Eigen::Vector3d &q0,
                                         const Eigen::Vector3d &q1, double eta)
    {
        double pts[12];
        uint8_t out = 0;
        pack(p0, p1, q0, q1, pts);
        check(ccd_dist_line_lt_batch(context(), 1, pts, &eta, &out), "ccd_dist_line_lt_batch");
        return out != 0;
    }
    static Eigen::Vector3d vertexFaceDistance(const Eigen::Vector3d &p, const Eigen::Vector3d &q0, const Eigen::Vector3d &q1,
                                              const Eigen::Vector3d &q2, double &q0bary, double &q1bary, double &q2bary)
    {
        double pts[12], vec[3], bary[3];
        pack(p, q0, q1, q2, pts);
        check(ccd_dist_vf_batch(context(), 1, pts, vec, bary), "ccd_dist_vf_batch");
        q0bary = bary[0]; q1bary = bary[1]; q2bary = bary[2];
        return Eigen::Vector3d(vec[0], vec[1], vec[2]);
    }
    static Eigen::Vector3d edgeEdgeDistance(const Eigen::Vector3d &p0, const Eigen::Vector3d &p1, const Eigen::Vector3d &q0,
                                            const Eigen::Vector3d &q1, double &p0bary, double &p1bary, double &q0bary, double &q1bary)
    {
        double pts[12], vec[3], bary[4];
        pack(p0, p1, q0, q1, pts);
        check(ccd_dist_ee_batch(context(), 1, pts, vec, bary), "ccd_dist_ee_batch");
        p0bary = bary[0]; p1bary = bary[1]; q0bary = bary[2]; q1bary = bary[3];
        return Eigen::Vector3d(vec[0], vec[1], vec[2]);
    }
    static double meshSelfDistance(const Eigen::VectorXd &verts, const Eigen::Matrix3Xi &faces, const std::set<int> &fixedVerts)
    {
        const int V = (int)(verts.size() / 3), F = (int)faces.cols();
        std::vector<double> q(3 * (size_t)V);
        for (size_t i = 0; i < q.size(); i++)
            q[i] = verts[i];
        std::vector<int32_t> f(3 * (size_t)F);
        for (int k = 0; k < F; k++)
            for (int j = 0; j < 3; j++)
                f[3 * (size_t)k + j] = faces.coeff(j, k);
        std::vector<uint8_t> fixedMask;
        if (!fixedVerts.empty())
        {
            fixedMask.assign(V, 0);
            for (std::set<int>::const_iterator it = fixedVerts.begin(); it != fixedVerts.end(); ++it)
                if (*it >= 0 && *it < V)
                    fixedMask[*it] = 1;
        }
        double d = 0;
        int64_t nvf = 0, nee = 0;
        check(ccd_mesh_self_distance(context(), V, q.data(), F, f.data(), fixedMask.empty() ? 0 : fixedMask.data(), &d, &nvf, &nee),
              "ccd_mesh_self_distance");
        std::cout << "Checking " << nvf << " vertex-face and " << nee << " edge-edge stencils" << std::endl;   // src/Distance.cpp:47
        return d;
    }
};

} // namespace ccdgpu

#endif
