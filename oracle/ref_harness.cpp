// TEST INFRASTRUCTURE — not product code.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the library built from this.
//
// C-ABI harness around the UNMODIFIED reference (evouga/collisiondetection) compiled from
// /root/reference/src/*.cpp against oracle/eigen_shim (see oracle/Makefile).  It drives the
// reference's own classes exactly the way example/AlecTest.cpp:86-111 and
// src/Distance.cpp:12-66 do and flattens std::set / History to plain arrays.
//
// Nothing in here re-implements reference arithmetic; every number returned was computed
// by the reference's object code.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <set>
#include <vector>
#include <utility>

#include <Eigen/Core>
#include "CTCD.h"
#include "Distance.h"
#include "History.h"
#include "Mesh.h"
#include "Stencils.h"
#include "RetrospectiveDetection.h"
#include "CTCDNarrowPhase.h"
#include "SeparatingPlaneNarrowPhase.h"
#include "TrivialBroadPhase.h"
#include "rpoly.h"

using namespace Eigen;

// KDOPBroadPhase.h and AABBBroadPhase.h both define `NodeComparator`, so each is wrapped in
// its own translation unit (ref_harness_kdop.cpp / ref_harness_aabb.cpp).
BroadPhase *ref_make_kdop();
BroadPhase *ref_make_aabb();

typedef std::chrono::steady_clock Clock;
static double secs(Clock::time_point a, Clock::time_point b)
{
    return std::chrono::duration<double>(b - a).count();
}

// History from CSR arrays.  Entry 0 of each vertex must have time 0 and the last time 1
// (src/History.cpp:8-39): History(q at first entry) + addHistory(middle) + finishHistory(last).
static History *make_history(int V, const long long *hoff, const double *htime, const double *hpos)
{
    VectorXd q0(3 * (long)V), q1(3 * (long)V);
    for (int v = 0; v < V; v++)
    {
        long long a = hoff[v], b = hoff[v + 1] - 1;
        for (int c = 0; c < 3; c++)
        {
            q0[3 * (long)v + c] = hpos[3 * a + c];
            q1[3 * (long)v + c] = hpos[3 * b + c];
        }
    }
    History *h = new History(q0);
    for (int v = 0; v < V; v++)
        for (long long e = hoff[v] + 1; e < hoff[v + 1] - 1; e++)
            h->addHistory(v, htime[e], Vector3d(hpos[3 * e], hpos[3 * e + 1], hpos[3 * e + 2]));
    h->finishHistory(q1);
    return h;
}

static void make_mesh(Mesh &m, int V, int F, const int *faces, const double *verts)
{
    m.vertices.resize(3 * (long)V);
    if (verts)
        for (long i = 0; i < 3 * (long)V; i++)
            m.vertices[i] = verts[i];
    else
        m.vertices.setZero();
    m.faces.resize(3, F);
    for (long i = 0; i < F; i++)
        for (int j = 0; j < 3; j++)
            m.faces.coeffRef(j, i) = faces[3 * i + j];
}

extern "C" {

void ref_free(void *p) { free(p); }

// kind: 13 = KDOPBroadPhase, 3 = AABBBroadPhase, 0 = TrivialBroadPhase.
// Outputs are malloc'ed 4-int tuples in std::set iteration (= lexicographic) order.
int ref_broadphase(int kind, int V, int F, const int *faces, const long long *hoff, const double *htime,
                   const double *hpos, double outerEta, const unsigned char *fixedMask, int **vf_out,
                   long long *nvf, int **ee_out, long long *nee, double *seconds)
{
    Mesh m;
    make_mesh(m, V, F, faces, 0);
    History *h = make_history(V, hoff, htime, hpos);
    std::set<int> fixedVerts;
    if (fixedMask)
        for (int v = 0; v < V; v++)
            if (fixedMask[v])
                fixedVerts.insert(v);
    BroadPhase *bp = kind == 13 ? ref_make_kdop() : kind == 3 ? ref_make_aabb() : (BroadPhase *)new TrivialBroadPhase();
    std::set<VertexFaceStencil> vfs;
    std::set<EdgeEdgeStencil> ees;
    Clock::time_point t0 = Clock::now();
    bp->findCollisionCandidates(*h, m, outerEta, vfs, ees, fixedVerts);
    Clock::time_point t1 = Clock::now();
    if (seconds)
        *seconds = secs(t0, t1);
    delete bp;
    delete h;
    *nvf = (long long)vfs.size();
    *nee = (long long)ees.size();
    int *vf = (int *)malloc(sizeof(int) * 4 * (vfs.size() + 1));
    int *ee = (int *)malloc(sizeof(int) * 4 * (ees.size() + 1));
    long k = 0;
    for (std::set<VertexFaceStencil>::iterator it = vfs.begin(); it != vfs.end(); ++it, ++k)
    {
        vf[4 * k] = it->p; vf[4 * k + 1] = it->q0; vf[4 * k + 2] = it->q1; vf[4 * k + 3] = it->q2;
    }
    k = 0;
    for (std::set<EdgeEdgeStencil>::iterator it = ees.begin(); it != ees.end(); ++it, ++k)
    {
        ee[4 * k] = it->p0; ee[4 * k + 1] = it->p1; ee[4 * k + 2] = it->q0; ee[4 * k + 3] = it->q1;
    }
    *vf_out = vf;
    *ee_out = ee;
    return 0;
}

// Per-stencil replay of CTCDNarrowPhase::checkVFS / checkEES (src/CTCDNarrowPhase.cpp:24-135)
// through the PUBLIC CTCD:: statics so the time of impact the reference has in hand when it
// returns true is kept (mapped from the stitched segment's parameter to History time).  `stage` = 0 miss, 1 VF/EE primitive, 2..4 / 2..5 VE, then VV.
static bool replay_vf(const History &h, const int *s, double eta, double *toi, int *stage)
{
    std::vector<int> verts(s, s + 4);
    std::vector<StitchedEntry> sh;
    h.stitchCommonHistory(verts, sh);
    for (size_t i = 0; i + 1 < sh.size(); i++)
    {
        const std::vector<Vector3d> &a = sh[i].pos, &b = sh[i + 1].pos;
        double t;
        if (CTCD::vertexFaceCTCD(a[0], a[1], a[2], a[3], b[0], b[1], b[2], b[3], eta, t)) { *toi = sh[i].time + t * (sh[i + 1].time - sh[i].time); *stage = 1; return true; }
        for (int e = 0; e < 3; e++)
            if (CTCD::vertexEdgeCTCD(a[0], a[1 + (e % 3)], a[1 + ((e + 1) % 3)], b[0], b[1 + (e % 3)], b[1 + ((e + 1) % 3)], eta, t)) { *toi = sh[i].time + t * (sh[i + 1].time - sh[i].time); *stage = 2 + e; return true; }
        for (int v = 0; v < 3; v++)
            if (CTCD::vertexVertexCTCD(a[0], a[1 + v], b[0], b[1 + v], eta, t)) { *toi = sh[i].time + t * (sh[i + 1].time - sh[i].time); *stage = 5 + v; return true; }
    }
    return false;
}

static bool replay_ee(const History &h, const int *s, double eta, double *toi, int *stage)
{
    std::vector<int> verts(s, s + 4);
    std::vector<StitchedEntry> sh;
    h.stitchCommonHistory(verts, sh);
    static const int ve[4][3] = {{0, 2, 3}, {1, 2, 3}, {2, 0, 1}, {3, 0, 1}};
    static const int vv[4][2] = {{0, 2}, {0, 3}, {1, 2}, {1, 3}};
    for (size_t i = 0; i + 1 < sh.size(); i++)
    {
        const std::vector<Vector3d> &a = sh[i].pos, &b = sh[i + 1].pos;
        double t;
        if (CTCD::edgeEdgeCTCD(a[0], a[1], a[2], a[3], b[0], b[1], b[2], b[3], eta, t)) { *toi = sh[i].time + t * (sh[i + 1].time - sh[i].time); *stage = 1; return true; }
        for (int e = 0; e < 4; e++)
            if (CTCD::vertexEdgeCTCD(a[ve[e][0]], a[ve[e][1]], a[ve[e][2]], b[ve[e][0]], b[ve[e][1]], b[ve[e][2]], eta, t)) { *toi = sh[i].time + t * (sh[i + 1].time - sh[i].time); *stage = 2 + e; return true; }
        for (int v = 0; v < 4; v++)
            if (CTCD::vertexVertexCTCD(a[vv[v][0]], a[vv[v][1]], b[vv[v][0]], b[vv[v][1]], eta, t)) { *toi = sh[i].time + t * (sh[i + 1].time - sh[i].time); *stage = 6 + v; return true; }
    }
    return false;
}

// Runs the reference's CTCDNarrowPhase::findCollisions (timed; flags = membership in its output
// sets), then replays every stencil for TOI / stage.  Returns the number of stencils whose
// replayed flag disagrees with findCollisions (must be 0).
// which: 0 = CTCDNarrowPhase, 1 = SeparatingPlaneNarrowPhase (no TOI replay for 1).
int ref_narrowphase(int which, int V, const long long *hoff, const double *htime, const double *hpos,
                    long long nvf, const int *vf, const double *vf_eta, long long nee, const int *ee,
                    const double *ee_eta, unsigned char *vf_hit, double *vf_toi, int *vf_stage,
                    unsigned char *ee_hit, double *ee_toi, int *ee_stage, double *seconds)
{
    History *h = make_history(V, hoff, htime, hpos);
    std::set<std::pair<VertexFaceStencil, double> > cvf;
    std::set<std::pair<EdgeEdgeStencil, double> > cee;
    for (long long i = 0; i < nvf; i++)
        cvf.insert(std::make_pair(VertexFaceStencil(vf[4 * i], vf[4 * i + 1], vf[4 * i + 2], vf[4 * i + 3]), vf_eta[i]));
    for (long long i = 0; i < nee; i++)
        cee.insert(std::make_pair(EdgeEdgeStencil(ee[4 * i], ee[4 * i + 1], ee[4 * i + 2], ee[4 * i + 3]), ee_eta[i]));
    std::set<VertexFaceStencil> ovf;
    std::set<EdgeEdgeStencil> oee;
    NarrowPhase *np = which == 0 ? (NarrowPhase *)new CTCDNarrowPhase() : (NarrowPhase *)new SeparatingPlaneNarrowPhase();
    Clock::time_point t0 = Clock::now();
    np->findCollisions(*h, cvf, cee, ovf, oee);
    Clock::time_point t1 = Clock::now();
    if (seconds)
        *seconds = secs(t0, t1);
    delete np;
    int disagree = 0;
    for (long long i = 0; i < nvf; i++)
    {
        bool in = ovf.count(VertexFaceStencil(vf[4 * i], vf[4 * i + 1], vf[4 * i + 2], vf[4 * i + 3])) != 0;
        vf_hit[i] = in;
        if (which == 0 && vf_toi)
        {
            double t = 0; int st = 0;
            bool r = replay_vf(*h, vf + 4 * i, vf_eta[i], &t, &st);
            vf_toi[i] = r ? t : 0.0;
            if (vf_stage) vf_stage[i] = st;
            if (r != in) disagree++;
        }
    }
    for (long long i = 0; i < nee; i++)
    {
        bool in = oee.count(EdgeEdgeStencil(ee[4 * i], ee[4 * i + 1], ee[4 * i + 2], ee[4 * i + 3])) != 0;
        ee_hit[i] = in;
        if (which == 0 && ee_toi)
        {
            double t = 0; int st = 0;
            bool r = replay_ee(*h, ee + 4 * i, ee_eta[i], &t, &st);
            ee_toi[i] = r ? t : 0.0;
            if (ee_stage) ee_stage[i] = st;
            if (r != in) disagree++;
        }
    }
    delete h;
    return disagree;
}

// Flat-array replay only (no std::set, no findCollisions): the "primitives over a flat candidate
// array" CPU baseline of SURVEY.md §8(d).  Stencil range [begin,end) so callers can thread it.
void ref_narrowphase_flat(int V, const long long *hoff, const double *htime, const double *hpos,
                          long long nvf, const int *vf, const double *vf_eta, long long nee, const int *ee,
                          const double *ee_eta, unsigned char *vf_hit, double *vf_toi, unsigned char *ee_hit,
                          double *ee_toi)
{
    History *h = make_history(V, hoff, htime, hpos);
    for (long long i = 0; i < nvf; i++)
    {
        double t = 0; int st = 0;
        vf_hit[i] = replay_vf(*h, vf + 4 * i, vf_eta[i], &t, &st);
        vf_toi[i] = vf_hit[i] ? t : 0.0;
    }
    for (long long i = 0; i < nee; i++)
    {
        double t = 0; int st = 0;
        ee_hit[i] = replay_ee(*h, ee + 4 * i, ee_eta[i], &t, &st);
        ee_toi[i] = ee_hit[i] ? t : 0.0;
    }
    delete h;
}

// ---- the four public primitives, batched.  Layout per item: start points then end points,
// each xyz, in the argument order of include/CTCD.h:36-79.  t is written only on a hit.
static Vector3d V3(const double *p) { return Vector3d(p[0], p[1], p[2]); }

void ref_vf_batch(long long n, const double *pts, const double *eta, unsigned char *hit, double *t)
{
    for (long long i = 0; i < n; i++)
    {
        const double *p = pts + 24 * i;
        double tt = 0;
        hit[i] = CTCD::vertexFaceCTCD(V3(p), V3(p + 3), V3(p + 6), V3(p + 9), V3(p + 12), V3(p + 15), V3(p + 18), V3(p + 21), eta[i], tt);
        if (hit[i]) t[i] = tt;
    }
}
void ref_ee_batch(long long n, const double *pts, const double *eta, unsigned char *hit, double *t)
{
    for (long long i = 0; i < n; i++)
    {
        const double *p = pts + 24 * i;
        double tt = 0;
        hit[i] = CTCD::edgeEdgeCTCD(V3(p), V3(p + 3), V3(p + 6), V3(p + 9), V3(p + 12), V3(p + 15), V3(p + 18), V3(p + 21), eta[i], tt);
        if (hit[i]) t[i] = tt;
    }
}
void ref_ve_batch(long long n, const double *pts, const double *eta, unsigned char *hit, double *t)
{
    for (long long i = 0; i < n; i++)
    {
        const double *p = pts + 18 * i;
        double tt = 0;
        hit[i] = CTCD::vertexEdgeCTCD(V3(p), V3(p + 3), V3(p + 6), V3(p + 9), V3(p + 12), V3(p + 15), eta[i], tt);
        if (hit[i]) t[i] = tt;
    }
}
void ref_vv_batch(long long n, const double *pts, const double *eta, unsigned char *hit, double *t)
{
    for (long long i = 0; i < n; i++)
    {
        const double *p = pts + 12 * i;
        double tt = 0;
        hit[i] = CTCD::vertexVertexCTCD(V3(p), V3(p + 3), V3(p + 6), V3(p + 9), eta[i], tt);
        if (hit[i]) t[i] = tt;
    }
}

// ---- include/Distance.h:14-174, batched.  pts: 4 points xyz per item.
void ref_dist_vf_batch(long long n, const double *pts, double *vec, double *bary)
{
    for (long long i = 0; i < n; i++)
    {
        const double *p = pts + 12 * i;
        Vector3d r = Distance::vertexFaceDistance(V3(p), V3(p + 3), V3(p + 6), V3(p + 9), bary[3 * i], bary[3 * i + 1], bary[3 * i + 2]);
        vec[3 * i] = r[0]; vec[3 * i + 1] = r[1]; vec[3 * i + 2] = r[2];
    }
}
void ref_dist_ee_batch(long long n, const double *pts, double *vec, double *bary)
{
    for (long long i = 0; i < n; i++)
    {
        const double *p = pts + 12 * i;
        Vector3d r = Distance::edgeEdgeDistance(V3(p), V3(p + 3), V3(p + 6), V3(p + 9), bary[4 * i], bary[4 * i + 1], bary[4 * i + 2], bary[4 * i + 3]);
        vec[3 * i] = r[0]; vec[3 * i + 1] = r[1]; vec[3 * i + 2] = r[2];
    }
}
void ref_dist_plane_lt_batch(long long n, const double *pts, const double *eta, unsigned char *out)
{
    for (long long i = 0; i < n; i++)
    {
        const double *p = pts + 12 * i;
        out[i] = Distance::vertexPlaneDistanceLessThan(V3(p), V3(p + 3), V3(p + 6), V3(p + 9), eta[i]);
    }
}
void ref_dist_line_lt_batch(long long n, const double *pts, const double *eta, unsigned char *out)
{
    for (long long i = 0; i < n; i++)
    {
        const double *p = pts + 12 * i;
        out[i] = Distance::lineLineDistanceLessThan(V3(p), V3(p + 3), V3(p + 6), V3(p + 9), eta[i]);
    }
}

// src/Distance.cpp:12-66.  (Prints its "Checking ..." line to stdout like the reference.)
double ref_mesh_self_distance(int V, const double *verts, int F, const int *faces, const unsigned char *fixedMask, double *seconds)
{
    VectorXd q(3 * (long)V);
    for (long i = 0; i < 3 * (long)V; i++)
        q[i] = verts[i];
    Matrix3Xi f;
    f.resize(3, F);
    for (long i = 0; i < F; i++)
        for (int j = 0; j < 3; j++)
            f.coeffRef(j, i) = faces[3 * i + j];
    std::set<int> fixedVerts;
    if (fixedMask)
        for (int v = 0; v < V; v++)
            if (fixedMask[v])
                fixedVerts.insert(v);
    Clock::time_point t0 = Clock::now();
    double d = Distance::meshSelfDistance(q, f, fixedVerts);
    Clock::time_point t1 = Clock::now();
    if (seconds)
        *seconds = secs(t0, t1);
    return d;
}

// Direct access to the vendored Jenkins-Traub (src/rpoly.h:49) for the mismatch classifier:
// returns the number of roots it reports (< degree means "did not converge").
int ref_rpoly(const double *op, int degree, double *zeror, double *zeroi)
{
    double c[7];
    for (int i = 0; i <= degree; i++)
        c[i] = op[i];
    RootFinder rf;
    return rf.rpoly(c, degree, zeror, zeroi);
}

// History::stitchCommonHistory (src/History.cpp:98-140) for 4 vertices -> times[ns], pos[ns*12].
int ref_stitch(int V, const long long *hoff, const double *htime, const double *hpos, const int *verts4,
               int cap, double *times, double *pos)
{
    History *h = make_history(V, hoff, htime, hpos);
    std::vector<int> verts(verts4, verts4 + 4);
    std::vector<StitchedEntry> sh;
    h->stitchCommonHistory(verts, sh);
    int n = (int)sh.size();
    for (int i = 0; i < n && i < cap; i++)
    {
        times[i] = sh[i].time;
        for (int j = 0; j < 4; j++)
            for (int c = 0; c < 3; c++)
                pos[12 * i + 3 * j + c] = sh[i].pos[j][c];
    }
    delete h;
    return n;
}

} // extern "C"
