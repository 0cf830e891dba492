// SeparatingPlaneNarrowPhase (src/SeparatingPlaneNarrowPhase.cpp:11-278) — the narrowphase ActiveLayers instantiates
// (src/ActiveLayers.cpp:28) — for one stencil, host+device.  The reference recurses on time intervals: a short interval
// inside one History segment is handed to the CTCD primitives; otherwise the stencil is evaluated at the interval's
// midpoint (closest points, Distance.h), a hit if closer than eta, else the plane between the closest points, moving
// with their mean velocity, is followed backwards and forwards along the four trajectories until one of them comes
// within eta/2 of it: only the times beyond those crossings can still collide and are searched recursively.
// Here the recursion is an explicit stack of intervals (same intervals; the result is a disjunction, so their order
// does not matter).  History in CSR form: hoff[V+1], htime[N], hpos[3N].
#pragma once
#include "ccd_stencil.cuh"
#include "ccd_distance.cuh"

namespace ccd {

struct HistView
{
    const long long *hoff;
    const double *htime;
    const double *hpos;
};

// History::getPosAtTime (src/History.cpp:49-80): idx = last entry with time <= t (the bisection there only narrows the
// linear scan's start), position interpolated inside [idx, idx+1]
CCD_FN void hist_pos_at(const HistView &H, int vert, double time, V3 &pos, int &idx)
{
    const long long b = H.hoff[vert], e = H.hoff[vert + 1];
    long long lo = b, hi = e;      // first entry with htime > time
    while (lo < hi)
    {
        const long long mid = (lo + hi) >> 1;
        if (H.htime[mid] <= time) lo = mid + 1; else hi = mid;
    }
    const long long next = lo, prev = next - 1;
    if (next == e)
    {
        pos = ldv(H.hpos + 3 * prev);
        idx = (int)(prev - b);
        return;
    }
    const double dt = H.htime[next] - H.htime[prev];
    const double alpha = (time - H.htime[prev]) / dt;
    pos = (1.0 - alpha) * ldv(H.hpos + 3 * prev) + alpha * ldv(H.hpos + 3 * next);
    idx = (int)(prev - b);
}

// SeparatingPlaneNarrowPhase::planeIntersect (:266-276)
CCD_FN double sp_plane_intersect(V3 planePos, V3 planeVel, V3 planeNormal, V3 ptOld, V3 ptNew, double ptdt, double eta)
{
    const double numerator = 0.5 * eta - dot(ptOld - planePos, planeNormal);
    double denom = -dot(planeVel, planeNormal);
    if (ptdt != 0.0)
        denom += dot(ptNew - ptOld, planeNormal) / ptdt;
    return numerator / denom;
}

// SeparatingPlaneNarrowPhase::planeTrajectoryIntersect (:220-264) for vertex `vert`
CCD_FN double sp_trajectory_intersect(const HistView &H, int vert, int startidx, bool forward, V3 planePos, V3 planeVel, V3 planeNormal, V3 ptstart,
                                      double timestart, double eta)
{
    const long long b = H.hoff[vert];
    const int size = (int)(H.hoff[vert + 1] - b);
    const double *ht = H.htime + b;
    const double *hp = H.hpos + 3 * b;
    {
        const double dt = forward ? ht[startidx] - timestart : timestart - ht[startidx];
        const V3 newpos = ldv(hp + 3 * startidx);
        const double t = sp_plane_intersect(planePos, (forward ? 1.0 : -1.0) * planeVel, planeNormal, ptstart, newpos, dt, eta);
        if (t >= 0 && t <= dt)
            return t;
    }
    if (forward)
    {
        for (int i = startidx; i < size; i++)
        {
            if (i == size - 1)
                return INFINITY;
            const int next = i + 1;
            const V3 newplanepos = planePos + (ht[i] - timestart) * planeVel;
            const double dt = ht[next] - ht[i];
            const double t = sp_plane_intersect(newplanepos, planeVel, planeNormal, ldv(hp + 3 * i), ldv(hp + 3 * next), dt, eta);
            if (t >= 0 && t <= dt)
                return ht[i] + t - timestart;
        }
    }
    else
    {
        for (int i = startidx; i >= 0; i--)
        {
            if (i == 0)
                return INFINITY;
            const int prev = i - 1;
            const V3 newplanepos = planePos - (timestart - ht[i]) * planeVel;
            const double dt = ht[i] - ht[prev];
            const double t = sp_plane_intersect(newplanepos, -1.0 * planeVel, planeNormal, ldv(hp + 3 * i), ldv(hp + 3 * prev), dt, eta);
            if (t >= 0 && t <= dt)
                return timestart - (ht[i] - t);
        }
    }
    return INFINITY;
}

// One step of SeparatingPlaneNarrowPhase::checkInterval (:47-218) on the interval [mint, maxt] of one stencil — the body of
// the recursion, free of any stack or queue (sp_check_stencil below keeps the intervals on a per-thread stack; the staged
// kernels of narrowphase.cu keep them in a global queue shared by all stencils).  Outcome:
//   SP_DROP   nothing left to do for this interval
//   SP_HIT    closer than eta at the midpoint
//   SP_LEAF   shorter than 3 eps inside one History segment: oldpos / newpos go to the CTCD primitives
//   SP_SPLIT  nsub (0..2) sub-intervals sub_lo / sub_hi remain to be searched
enum { SP_DROP = 0, SP_HIT = 1, SP_LEAF = 2, SP_SPLIT = 3 };
template <bool IS_VF>
CCD_FN int sp_interval_step(const HistView &H, const int *verts, double eta, double eps, double mint, double maxt, V3 *oldpos, V3 *newpos, int &nsub,
                            double *sub_lo, double *sub_hi)
{
    nsub = 0;
    if (maxt < mint)
        return SP_DROP;
    if (maxt - mint < 3 * eps)
    {
        bool ok = true;
        for (int i = 0; i < 4; i++)
        {
            int oldidx, newidx;
            hist_pos_at(H, verts[i], mint, oldpos[i], oldidx);
            hist_pos_at(H, verts[i], maxt, newpos[i], newidx);
            if (oldidx != newidx)
            {
                ok = false;
                break;
            }
        }
        if (ok)
            return SP_LEAF;
    }
    const double midt = 0.5 * (mint + maxt);
    V3 midpos[4];
    int mididx[4];
    for (int i = 0; i < 4; i++)
        hist_pos_at(H, verts[i], midt, midpos[i], mididx[i]);
    double bary1, bary2, bary3, bary4 = 0;
    V3 closest;
    if (IS_VF)
        closest = dist_vf(midpos[0], midpos[1], midpos[2], midpos[3], bary1, bary2, bary3);
    else
        closest = dist_ee(midpos[0], midpos[1], midpos[2], midpos[3], bary1, bary2, bary3, bary4);
    const double distsq = dot(closest, closest);
    if (distsq < eta * eta)
        return SP_HIT;
    // separating plane (:167-189): bary * (next - prev) / dt, term by term as the reference writes it
    const double w[4] = {IS_VF ? 1.0 : bary1, IS_VF ? bary1 : bary2, IS_VF ? bary2 : bary3, IS_VF ? bary3 : bary4};
    V3 term[4], wpos[4];
    for (int i = 0; i < 4; i++)
    {
        const long long e0 = H.hoff[verts[i]] + mididx[i];
        const V3 prevpos = ldv(H.hpos + 3 * e0), nextpos = ldv(H.hpos + 3 * (e0 + 1));
        const double dts = H.htime[e0 + 1] - H.htime[e0];
        V3 d = nextpos - prevpos;
        if (!(IS_VF && i == 0)) d = w[i] * d;
        term[i] = mk(d.x / dts, d.y / dts, d.z / dts);
        wpos[i] = (IS_VF && i == 0) ? midpos[0] : w[i] * midpos[i];
    }
    const V3 planepos = 0.5 * (((wpos[0] + wpos[1]) + wpos[2]) + wpos[3]);
    const V3 planevel = 0.5 * (((term[0] + term[1]) + term[2]) + term[3]);
    const double cn = sqrt(dot(closest, closest));
    const V3 normal = mk(closest.x / cn, closest.y / cn, closest.z / cn);
    double tb = INFINITY, tf = INFINITY;
    for (int vert = 0; vert < 4; vert++)
    {
        const double sign = (vert == 0 || (vert == 1 && !IS_VF)) ? -1.0 : 1.0;
        tb = smin(tb, sp_trajectory_intersect(H, verts[vert], mididx[vert], false, planepos, planevel, sign * normal, midpos[vert], midt, eta));
        tf = smin(tf, sp_trajectory_intersect(H, verts[vert], mididx[vert] + 1, true, planepos, planevel, sign * normal, midpos[vert], midt, eta));
    }
    if (tf < 1.0)
    {
        sub_lo[nsub] = midt + smax(0.0, tf - eps);
        sub_hi[nsub] = maxt;
        nsub++;
    }
    if (tb < 1.0)
    {
        sub_lo[nsub] = mint;
        sub_hi[nsub] = midt - smax(0.0, tb - eps);
        nsub++;
    }
    return SP_SPLIT;
}

#define SP_STACK 96
// SeparatingPlaneNarrowPhase::checkInterval (:47-218) over [0,1].  Returns 1 hit, 0 miss, -1 when the interval stack
// overflows (the caller reports an error; with eps = minimum gap / 4 the search is ~log2(1/eps) deep).
template <bool IS_VF> CCD_FN int sp_check_stencil(const HistView &H, const int *verts, double eta, double eps)
{
    double smin_[SP_STACK], smax_[SP_STACK];
    int sp = 0;
    smin_[sp] = 0.0;
    smax_[sp] = 1.0;
    sp++;
    while (sp > 0)
    {
        sp--;
        V3 oldpos[4], newpos[4];
        int nsub;
        double lo[2], hi[2];
        const int r = sp_interval_step<IS_VF>(H, verts, eta, eps, smin_[sp], smax_[sp], oldpos, newpos, nsub, lo, hi);
        if (r == SP_HIT)
            return 1;
        if (r == SP_LEAF)
        {
            // the primitive, then the vertex-edge and vertex-vertex tests in the reference's order (:76-140): the same
            // sequence as CTCDNarrowPhase's on this linear piece
            double t;
            if (stencil_segment_full<IS_VF>(oldpos, newpos, eta, t) != 0)
                return 1;
            continue;
        }
        if (r == SP_DROP)
            continue;
        if (sp + 2 > SP_STACK)
            return -1;
        for (int k = 0; k < nsub; k++)
        {
            smin_[sp] = lo[k];
            smax_[sp] = hi[k];
            sp++;
        }
    }
    return 0;
}

} // namespace ccd
