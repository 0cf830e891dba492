// Distance.h queries (include/Distance.h:14-174) as host+device functions: operation order follows the reference so the
// results are bit-identical to its CPU build (translation units compiled with --fmad=false).  Used by distance.cu
// (batched kernels, meshSelfDistance) and by the separating-plane narrowphase (ccd_sepplane.cuh).
#pragma once
#include "ccd_math.cuh"

namespace ccd {

CCD_FN double clamp01(double u) { return smin(1.0, smax(u, 0.0)); }

// Distance::vertexPlaneDistanceLessThan, include/Distance.h:14-19
CCD_FN bool plane_lt(V3 p, V3 q0, V3 q1, V3 q2, double eta)
{
    V3 c = cross(q1 - q0, q2 - q0);
    return dot(c, p - q0) * dot(c, p - q0) < eta * eta * dot(c, c);
}
// Distance::lineLineDistanceLessThan, include/Distance.h:23-28
CCD_FN bool line_lt(V3 p0, V3 p1, V3 q0, V3 q1, double eta)
{
    V3 c = cross(p1 - p0, q1 - q0);
    return dot(c, q0 - p0) * dot(c, q0 - p0) < eta * eta * dot(c, c);
}

// Distance::vertexFaceDistance, include/Distance.h:32-115 (closest point on triangle, region by region)
CCD_FN V3 dist_vf(V3 p, V3 q0, V3 q1, V3 q2, double &b0, double &b1, double &b2)
{
    V3 ab = q1 - q0, ac = q2 - q0, ap = p - q0;
    double d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0 && d2 <= 0) { b0 = 1.0; b1 = 0.0; b2 = 0.0; return q0 - p; }
    V3 bp = p - q1;
    double d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0 && d4 <= d3) { b0 = 0.0; b1 = 1.0; b2 = 0.0; return q1 - p; }
    double vc = d1 * d4 - d3 * d2;
    if ((vc <= 0) && (d1 >= 0) && (d3 <= 0))
    {
        double v = d1 / (d1 - d3);
        b0 = 1.0 - v; b1 = v; b2 = 0;
        return (q0 + v * ab) - p;
    }
    V3 cp = p - q2;
    double d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0 && d5 <= d6) { b0 = 0; b1 = 0; b2 = 1.0; return q2 - p; }
    double vb = d5 * d2 - d1 * d6;
    if ((vb <= 0) && (d2 >= 0) && (d6 <= 0))
    {
        double w = d2 / (d2 - d6);
        b0 = 1 - w; b1 = 0; b2 = w;
        return (q0 + w * ac) - p;
    }
    double va = d3 * d6 - d5 * d4;
    if ((va <= 0) && (d4 - d3 >= 0) && (d5 - d6 >= 0))
    {
        double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        b0 = 0; b1 = 1.0 - w; b2 = w;
        return (q1 + w * (q2 - q1)) - p;
    }
    double denom = 1.0 / (va + vb + vc);
    double v = vb * denom;
    double w = vc * denom;
    double u = 1.0 - v - w;
    b0 = u; b1 = v; b2 = w;
    return ((u * q0 + v * q1) + w * q2) - p;
}

// Distance::edgeEdgeDistance, include/Distance.h:119-174
CCD_FN V3 dist_ee(V3 p0, V3 p1, V3 q0, V3 q1, double &bp0, double &bp1, double &bq0, double &bq1)
{
    V3 d1 = p1 - p0, d2 = q1 - q0, r = p0 - q0;
    double a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
    double s, t;
    double c = dot(d1, r), b = dot(d1, d2);
    double denom = a * e - b * b;
    if (denom != 0.0) s = clamp01((b * f - c * e) / denom);
    else s = 0;
    double tnom = b * s + f;
    if (tnom < 0 || e == 0)
    {
        t = 0;
        if (a == 0) s = 0; else s = clamp01(-c / a);
    }
    else if (tnom > e)
    {
        t = 1.0;
        if (a == 0) s = 0; else s = clamp01((b - c) / a);
    }
    else
        t = tnom / e;
    V3 c1 = p0 + s * d1;
    V3 c2 = q0 + t * d2;
    bp0 = 1.0 - s; bp1 = s; bq0 = 1.0 - t; bq1 = t;
    return c2 - c1;
}

} // namespace ccd
