/* TEST INFRASTRUCTURE — not product code (see ccd_oracle.h).
 *
 * Plain-C restatement of the reference's CCD hot path.  Citations are file:line into the
 * reference tree (evouga/collisiondetection).  Arithmetic follows the reference's operation
 * order term by term and the file is compiled with -ffp-contract=off, so polynomial
 * coefficients and leaf boxes are bit-identical to the reference built in oracle/_ref.
 *
 * Root finding: src/rpoly.h (Jenkins-Traub) is NOT restated.  BASELINE.json's north_star
 * replaces it with an interval method; `orc_roots01` is the CPU statement of that method and
 * the CUDA kernels implement the same steps with the same fused multiply-adds, so GPU results
 * are bit-identical to this file while differences against rpoly are confined to the classes
 * listed in SURVEY.md §8(c) (counted by tests/test_oracle_vs_reference.py).
 */
#include "ccd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------
 * 3-vectors with the evaluation order fixed in oracle/eigen_shim/Eigen/Core
 * ---------------------------------------------------------------------------------------- */
typedef struct { double x, y, z; } v3;

static v3 mk(double x, double y, double z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
static v3 ld(const double *p) { return mk(p[0], p[1], p[2]); }
static v3 sub(v3 a, v3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 add(v3 a, v3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 scl(double s, v3 a) { return mk(s * a.x, s * a.y, s * a.z); }
static double dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static v3 cross(v3 a, v3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

/* std::max(a,b) / std::min(a,b) including their NaN behaviour */
static double smax(double a, double b) { return (a < b) ? b : a; }
static double smin(double a, double b) { return (b < a) ? b : a; }

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void orc_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------
 * Real-root isolation on [0,1]  (replaces src/rpoly.h:49-261; north_star item (2))
 *
 * Levels q_d = p, q_{m-1} = q_m'.  The number of sign variations V_m of the Bernstein
 * coefficients of q_m on [0,1] bounds its roots in (0,1): V_m = 0 -> none, V_m = 1 -> exactly
 * one.  Descend from m = d until a level is decided that way (or the quadratic is reached and
 * solved in closed form), then climb back: the roots of q_{m-1} split [0,1] into pieces on
 * which q_m is monotone, so each piece holds at most one root, found by a bracketed Newton
 * iteration.  Every multiply-add below that is written fma() is a single fused operation.
 * ---------------------------------------------------------------------------------------- */
static const double RBINOM[7][7] = {
    {1.0, 0, 0, 0, 0, 0, 0},
    {1.0, 1.0, 0, 0, 0, 0, 0},
    {1.0, 1.0 / 2.0, 1.0, 0, 0, 0, 0},
    {1.0, 1.0 / 3.0, 1.0 / 3.0, 1.0, 0, 0, 0},
    {1.0, 1.0 / 4.0, 1.0 / 6.0, 1.0 / 4.0, 1.0, 0, 0},
    {1.0, 1.0 / 5.0, 1.0 / 10.0, 1.0 / 10.0, 1.0 / 5.0, 1.0, 0},
    {1.0, 1.0 / 6.0, 1.0 / 15.0, 1.0 / 20.0, 1.0 / 15.0, 1.0 / 6.0, 1.0}};

static int sign_variations(const double *b, int count)
{
    int v = 0, last = 0, i;
    for (i = 0; i < count; i++)
    {
        int s = (b[i] > 0.0) - (b[i] < 0.0);
        if (s != 0)
        {
            if (last != 0 && s != last)
                v++;
            last = s;
        }
    }
    return v;
}

static double horner_fma(const double *c, int m, double x)
{
    double f = c[0];
    int i;
    for (i = 1; i <= m; i++)
        f = fma(f, x, c[i]);
    return f;
}

/* Root of the degree-m polynomial c in (lo,hi); f(lo) has the sign of flo, f(hi) the opposite. */
static double solve_bracket(const double *c, int m, double lo, double hi, double flo)
{
    double x = 0.5 * (lo + hi);
    double dxold = hi - lo, dx = dxold;
    int it, i;
    for (it = 0; it < 128; it++)
    {
        double f = c[0], df = 0.0, xn, step;
        int bisect = 0;
        for (i = 1; i <= m; i++)
        {
            df = fma(df, x, f);
            f = fma(f, x, c[i]);
        }
        if (f == 0.0)
            return x;
        if ((f < 0.0) == (flo < 0.0))
            lo = x;
        else
            hi = x;
        step = f / df;
        xn = x - step;
        if (!(xn > lo && xn < hi))
            bisect = 1;
        else if (fabs(2.0 * f) > fabs(dxold * df))
            bisect = 1;
        dxold = dx;
        if (bisect)
        {
            dx = 0.5 * (hi - lo);
            xn = lo + dx;
            if (!(xn > lo && xn < hi))
                return xn;
        }
        else
            dx = step;
        if (fabs(xn - x) <= 8.9e-16 * fabs(xn))
            return xn;
        x = xn;
    }
    return x;
}

int orc_roots01(const double *c, int d, double *roots)
{
    double P[7][7];  /* P[m][0..m]: descending power coefficients of q_m */
    double B[7][7];  /* B[m][0..m]: Bernstein coefficients of q_m (up to a positive factor) */
    double brk[8], fv[8], cur[6];
    int m, i, k, m0, ncur;

    for (i = 0; i <= d; i++)
        P[d][i] = c[i];
    for (m = d; m > 2; m--)
        for (i = 0; i < m; i++)
            P[m - 1][i] = P[m][i] * (double)(m - i);

    for (i = 0; i <= d; i++)
        B[d][i] = c[d - i] * RBINOM[d][i];
    for (k = 1; k <= d; k++)
        for (i = d; i >= k; i--)
            B[d][i] = B[d][i] + B[d][i - 1];
    for (m = d; m > 2; m--)
        for (i = 0; i < m; i++)
            B[m - 1][i] = B[m][i + 1] - B[m][i];

    /* lowest level that has to be looked at */
    ncur = 0;
    for (m0 = d; m0 >= 2; m0--)
    {
        if (B[m0][0] != 0.0 && B[m0][m0] != 0.0)
        {
            int v = sign_variations(B[m0], m0 + 1);
            if (v == 0)
                break;
            if (v == 1)
            {
                double f0 = P[m0][m0], f1 = horner_fma(P[m0], m0, 1.0);
                if ((f0 < 0.0 && f1 > 0.0) || (f0 > 0.0 && f1 < 0.0))
                    cur[ncur++] = solve_bracket(P[m0], m0, 0.0, 1.0, f0);
                break;
            }
        }
        if (m0 == 2)
        {
            /* both critical points may lie inside: closed form, cancellation-free */
            double a = P[2][0], b = P[2][1], cc = P[2][2];
            double D = fma(b, b, -4.0 * a * cc);
            if (D >= 0.0)
            {
                double q = -0.5 * (b + (b < 0.0 ? -sqrt(D) : sqrt(D)));
                double r0 = q / a, r1 = (q != 0.0) ? cc / q : r0;
                if (r0 > r1) { double t = r0; r0 = r1; r1 = t; }
                if (r0 > 0.0 && r0 < 1.0) cur[ncur++] = r0;
                if (r1 > 0.0 && r1 < 1.0 && r1 != r0) cur[ncur++] = r1;
            }
            break;
        }
    }
    if (m0 == d)
    {
        /* decided at the top level: interior roots only (end coefficients are non-zero) */
        for (i = 0; i < ncur; i++)
            roots[i] = cur[i];
        return ncur;
    }

    /* climb: cur = roots of q_{m-1} strictly inside (0,1) */
    for (m = m0 + 1; m <= d; m++)
    {
        int nb = 0, nr = 0, last = (m == d);
        double out[7];
        brk[nb++] = 0.0;
        for (i = 0; i < ncur; i++)
            brk[nb++] = cur[i];
        brk[nb++] = 1.0;
        for (i = 0; i < nb; i++)
            fv[i] = horner_fma(P[m], m, brk[i]);
        for (i = 0; i + 1 < nb; i++)
        {
            if (fv[i] == 0.0)
            {
                if ((i > 0 || last) && (nr == 0 || out[nr - 1] != brk[i]))
                    out[nr++] = brk[i];
            }
            else if ((fv[i] < 0.0 && fv[i + 1] > 0.0) || (fv[i] > 0.0 && fv[i + 1] < 0.0))
            {
                double r = solve_bracket(P[m], m, brk[i], brk[i + 1], fv[i]);
                if (nr == 0 || out[nr - 1] != r)
                    out[nr++] = r;
            }
        }
        if (last && fv[nb - 1] == 0.0 && (nr == 0 || out[nr - 1] != 1.0))
            out[nr++] = 1.0;
        ncur = 0;
        for (i = 0; i < nr; i++)
            if (last || (out[i] > 0.0 && out[i] < 1.0))
                cur[ncur++] = out[i];
    }
    for (i = 0; i < ncur; i++)
        roots[i] = cur[i];
    return ncur;
}

/* ------------------------------------------------------------------------------------------
 * CTCD::findIntervals and helpers  (src/CTCD.cpp:38-177, include/CTCD.h:7-26)
 * ---------------------------------------------------------------------------------------- */
typedef struct { int n; double l[8], u[8]; } ivals;

/* TimeInterval ctor, include/CTCD.h:9-14 */
static void push_interval(ivals *iv, double tl, double tu)
{
    double l = tl, u = tu;
    if (l > u) { double t = l; l = u; u = t; }
    l = smax(l, 0.0);
    u = smin(u, 1.0);
    iv->l[iv->n] = l;
    iv->u[iv->n] = u;
    iv->n++;
}

/* CTCD::getQuadRoots, src/CTCD.cpp:38-56 */
static int quad_roots(double a, double b, double c, double *t0, double *t1)
{
    int roots = 0;
    int sign = 1;
    double D;
    if (b < 0)
        sign = -1;
    D = b * b - 4 * a * c;
    if (D >= 0)
    {
        double q = -0.5 * (b + sign * sqrt(D));
        roots = 2;
        *t0 = q / a;
        *t1 = c / q;
        if (*t0 > *t1) { double t = *t0; *t0 = *t1; *t1 = t; }
    }
    return roots;
}

/* CTCD::checkInterval, src/CTCD.cpp:59-79: plain (unfused) Horner at the clamped midpoint */
static void check_interval(double t1, double t2, const double *op, int degree, ivals *iv, int pos)
{
    double tmid, f;
    int i;
    t1 = smax(0.0, t1);
    t2 = smax(0.0, t2);
    t1 = smin(1.0, t1);
    t2 = smin(1.0, t2);
    tmid = (t2 + t1) / 2;
    f = op[0];
    for (i = 1; i <= degree; i++)
    {
        f *= tmid;
        f += op[i];
    }
    if (pos && f >= 0)
        push_interval(iv, t1, t2);
    else if (!pos && f <= 0)
        push_interval(iv, t1, t2);
}

/* CTCD::couldHaveRoots, src/CTCD.cpp:81-94 */
static int could_have_roots(const double *op, int degree, int pos)
{
    double result = 0;
    int i;
    if ((pos && op[0] > 0) || (!pos && op[0] < 0))
        result = op[0];
    for (i = 1; i < degree; i++)
    {
        result *= 1.0;
        if ((pos && op[i] > 0) || (!pos && op[i] < 0))
            result += op[i];
    }
    result *= 1.0;
    result += op[degree];
    return !((pos && result < 0) || (!pos && result > 0));
}

/* CTCD::findIntervals, src/CTCD.cpp:98-177 */
static void find_intervals(double *op, int n, ivals *iv, int pos)
{
    int roots = 0, reducedDegree = n, i;
    double time[6];
    double maxval = 0;

    for (i = 0; i <= n; i++)
        maxval = smax(maxval, fabs(op[i]));
    if (maxval != 0)
        for (i = 0; i <= n; i++)
            op[i] /= maxval;

    for (i = 0; i < n; i++)
    {
        if (op[i] == 0)
            reducedDegree--;
        else
            break;
    }
    if (reducedDegree < n)
        for (i = 0; i <= reducedDegree; i++)
            op[i] = op[i + n - reducedDegree];

    if (reducedDegree > 2)
    {
        if (!could_have_roots(op, reducedDegree, pos))
            return;
        roots = orc_roots01(op, reducedDegree, time);      /* was: RootFinder::rpoly */
    }
    else if (reducedDegree == 2)
        roots = quad_roots(op[0], op[1], op[2], &time[0], &time[1]);
    else if (reducedDegree == 1)
    {
        time[0] = -op[1] / op[0];
        roots = 1;
    }
    else
    {
        if ((!pos && op[0] <= 0) || (pos && op[0] >= 0))
            push_interval(iv, 0, 1.0);
        return;
    }

    if (roots > 0)
    {
        /* time[] is already ascending (quad_roots swaps; orc_roots01 returns sorted) */
        if (time[0] >= 0)
            check_interval(0, time[0], op, reducedDegree, iv, pos);
        for (i = 0; i < roots - 1; i++)
            if (!((time[i] < 0 && time[i + 1] < 0) || (time[i] > 1.0 && time[i + 1] > 1.0)))
                check_interval(time[i], time[i + 1], op, reducedDegree, iv, pos);
        if (time[roots - 1] <= 1.0)
            check_interval(time[roots - 1], 1.0, op, reducedDegree, iv, pos);
    }
    else
        check_interval(0.0, 1.0, op, reducedDegree, iv, pos);
}

int orc_find_intervals(double *op, int n, int pos, double *l, double *u)
{
    ivals iv;
    int i;
    iv.n = 0;
    find_intervals(op, n, &iv, pos);
    for (i = 0; i < iv.n; i++) { l[i] = iv.l[i]; u[i] = iv.u[i]; }
    return iv.n;
}

/* ------------------------------------------------------------------------------------------
 * Coefficient builders  (src/CTCD.cpp:179-257)
 * ---------------------------------------------------------------------------------------- */
/* planePoly3D coefficients, src/CTCD.cpp:222-227 */
static void plane_coeffs(v3 x10, v3 x20, v3 x30, v3 v10, v3 v20, v3 v30, double *op)
{
    op[0] = dot(v10, cross(v20, v30));
    op[1] = dot(x10, cross(v20, v30)) + dot(v10, cross(x20, v30)) + dot(v10, cross(v20, x30));
    op[2] = dot(x10, cross(x20, v30)) + dot(x10, cross(v20, x30)) + dot(v10, cross(x20, x30));
    op[3] = dot(x10, cross(x20, x30));
}

/* distancePoly3D coefficients, src/CTCD.cpp:240-255 */
static void distance_coeffs(v3 x10, v3 x20, v3 x30, v3 v10, v3 v20, v3 v30, double minDSquared, double *op)
{
    double abcd[4];
    double A, B, C, D;
    v3 E, F, G;
    plane_coeffs(x10, x20, x30, v10, v20, v30, abcd);
    A = abcd[0]; B = abcd[1]; C = abcd[2]; D = abcd[3];
    E = cross(x20, x30);
    F = add(cross(x20, v30), cross(v20, x30));
    G = cross(v20, v30);
    op[0] = A * A;
    op[1] = 2 * A * B;
    op[2] = B * B + 2 * A * C - dot(G, G) * minDSquared;
    op[3] = 2 * A * D + 2 * B * C - 2 * dot(G, F) * minDSquared;
    op[4] = 2 * B * D + C * C - (2 * dot(G, E) + dot(F, F)) * minDSquared;
    op[5] = 2 * C * D - 2 * dot(F, E) * minDSquared;
    op[6] = D * D - dot(E, E) * minDSquared;
}

/* barycentricPoly3D coefficients, src/CTCD.cpp:187-210 */
static void barycentric_coeffs(v3 x10, v3 x20, v3 x30, v3 v10, v3 v20, v3 v30, double *op)
{
    double A = dot(x10, x10);
    double B = 2 * dot(x10, v10);
    double C = dot(v10, v10);
    double D = dot(x20, x10);
    double E = dot(x20, v10) + dot(v20, x10);
    double F = dot(v20, v10);
    double G = dot(x30, x20);
    double H = dot(x30, v20) + dot(v30, x20);
    double I = dot(v30, v20);
    double J = dot(x30, x10);
    double K = dot(x30, v10) + dot(v30, x10);
    double L = dot(v30, v10);
    op[0] = F * L - C * I;
    op[1] = F * K + E * L - C * H - B * I;
    op[2] = F * J + D * L + E * K - C * G - A * I - B * H;
    op[3] = D * K + E * J - A * H - B * G;
    op[4] = D * J - A * G;
}

/* TimeInterval::overlap (pair), src/CTCD.cpp:10-13 */
static int overlap2(double al, double au, double bl, double bu) { return !(al > bu || bl > au); }

/* ------------------------------------------------------------------------------------------
 * The four primitives  (src/CTCD.cpp:259-692)
 * ---------------------------------------------------------------------------------------- */
/* CTCD::vertexFaceCTCD, src/CTCD.cpp:413-508.  p: q0..q3 start then q0..q3 end */
static int vertex_face(const double *p, double eta, double *t)
{
    double minD = eta * eta;
    v3 q0s = ld(p), q1s = ld(p + 3), q2s = ld(p + 6), q3s = ld(p + 9);
    v3 v0 = sub(ld(p + 12), q0s), v1 = sub(ld(p + 15), q1s), v2 = sub(ld(p + 18), q2s), v3_ = sub(ld(p + 21), q3s);
    ivals cop, e1, e2, e3;
    double op[7];
    v3 x10, x20, x30, v10, v20, v30;
    int i, j, k, l, col = 0;
    double mint = 1.0;
    cop.n = e1.n = e2.n = e3.n = 0;

    x10 = sub(q0s, q1s); v10 = sub(v0, v1);
    x20 = cross(sub(q3s, q1s), sub(q2s, q1s)); v20 = cross(sub(v3_, v1), sub(v2, v1));
    x30 = sub(q3s, q1s); v30 = sub(v3_, v1);
    plane_coeffs(x10, x20, x30, v10, v20, v30, op);
    find_intervals(op, 3, &e1, 1);
    if (e1.n == 0) return 0;

    x10 = sub(q0s, q2s); v10 = sub(v0, v2);
    x20 = cross(sub(q1s, q2s), sub(q3s, q2s)); v20 = cross(sub(v1, v2), sub(v3_, v2));
    x30 = sub(q1s, q2s); v30 = sub(v1, v2);
    plane_coeffs(x10, x20, x30, v10, v20, v30, op);
    find_intervals(op, 3, &e2, 1);
    if (e2.n == 0) return 0;

    x10 = sub(q0s, q3s); v10 = sub(v0, v3_);
    x20 = cross(sub(q2s, q3s), sub(q1s, q3s)); v20 = cross(sub(v2, v3_), sub(v1, v3_));
    x30 = sub(q2s, q3s); v30 = sub(v2, v3_);
    plane_coeffs(x10, x20, x30, v10, v20, v30, op);
    find_intervals(op, 3, &e3, 1);
    if (e3.n == 0) return 0;

    x10 = sub(q0s, q1s); x20 = sub(q2s, q1s); x30 = sub(q3s, q1s);
    v10 = sub(v0, v1); v20 = sub(v2, v1); v30 = sub(v3_, v1);
    distance_coeffs(x10, x20, x30, v10, v20, v30, minD, op);
    find_intervals(op, 6, &cop, 0);
    if (cop.n == 0) return 0;

    for (i = 0; i < cop.n; i++)
        for (j = 0; j < e1.n; j++)
            for (k = 0; k < e2.n; k++)
                for (l = 0; l < e3.n; l++)
                {
                    double L4[4], U4[4];
                    int a, b, ok = 1;
                    L4[0] = cop.l[i]; U4[0] = cop.u[i];
                    L4[1] = e1.l[j]; U4[1] = e1.u[j];
                    L4[2] = e2.l[k]; U4[2] = e2.u[k];
                    L4[3] = e3.l[l]; U4[3] = e3.u[l];
                    for (a = 0; a < 4 && ok; a++)
                        for (b = a + 1; b < 4; b++)
                            if (!overlap2(L4[a], U4[a], L4[b], U4[b])) { ok = 0; break; }
                    if (ok)
                    {
                        double il = 0.0;   /* TimeInterval::intersect starts from (0,1), src/CTCD.cpp:28-36 */
                        for (a = 0; a < 4; a++) il = smax(L4[a], il);
                        mint = smin(il, mint);
                        col = 1;
                    }
                }
    if (col) { *t = mint; return 1; }
    return 0;
}

/* CTCD::edgeEdgeCTCD, src/CTCD.cpp:259-411.  p: (q0,p0,q1,p1) start then end */
static int edge_edge(const double *p, double eta, double *t)
{
    double minD = eta * eta;
    v3 q0s = ld(p), p0s = ld(p + 3), q1s = ld(p + 6), p1s = ld(p + 9);
    v3 vq0 = sub(ld(p + 12), q0s), vp0 = sub(ld(p + 15), p0s), vq1 = sub(ld(p + 18), q1s), vp1 = sub(ld(p + 21), p1s);
    ivals raw, cop, par, a0, a1, b0, b1;
    double op[7];
    v3 x10, x20, x30, v10, v20, v30;
    int i, j, k, l, m, q, col = 0;
    double mint = 1.0;
    raw.n = cop.n = par.n = a0.n = a1.n = b0.n = b1.n = 0;

    x10 = sub(p0s, p1s); x20 = sub(p0s, q0s); x30 = sub(p1s, q1s);
    v10 = sub(vp0, vp1); v20 = sub(vp0, vq0); v30 = sub(vp1, vq1);
    distance_coeffs(x10, x20, x30, v10, v20, v30, minD, op);
    find_intervals(op, 6, &raw, 0);

    /* parallel-edge classification, src/CTCD.cpp:290-308 */
    for (i = 0; i < raw.n; i++)
    {
        double midt = (raw.u[i] + raw.l[i]) / 2;
        v3 c;
        x10 = add(sub(q0s, p0s), scl(midt, sub(vq0, vp0)));
        x20 = add(sub(q1s, p1s), scl(midt, sub(vq1, vp1)));
        c = cross(x10, x20);
        if (sqrt(dot(c, c)) < 1e-8) { par.l[par.n] = raw.l[i]; par.u[par.n] = raw.u[i]; par.n++; }
        else { cop.l[cop.n] = raw.l[i]; cop.u[cop.n] = raw.u[i]; cop.n++; }
    }
    if (cop.n == 0) return 0;

    x10 = sub(p1s, q1s); v10 = sub(vp1, vq1);
    x20 = sub(p0s, q0s); v20 = sub(vp0, vq0);
    x30 = sub(q0s, q1s); v30 = sub(vq0, vq1);
    barycentric_coeffs(x10, x20, x30, v10, v20, v30, op);
    find_intervals(op, 4, &a0, 1);
    if (a0.n == 0) return 0;

    x20 = sub(q0s, p0s); v20 = sub(vq0, vp0);
    x30 = sub(p0s, q1s); v30 = sub(vp0, vq1);
    barycentric_coeffs(x10, x20, x30, v10, v20, v30, op);
    find_intervals(op, 4, &a1, 1);
    if (a1.n == 0) return 0;

    x10 = sub(p0s, q0s); v10 = sub(vp0, vq0);
    x20 = sub(p1s, q1s); v20 = sub(vp1, vq1);
    x30 = sub(q1s, q0s); v30 = sub(vq1, vq0);
    barycentric_coeffs(x10, x20, x30, v10, v20, v30, op);
    find_intervals(op, 4, &b0, 1);
    if (b0.n == 0) return 0;

    x20 = sub(q1s, p1s); v20 = sub(vq1, vp1);
    x30 = sub(p1s, q0s); v30 = sub(vp1, vq0);
    barycentric_coeffs(x10, x20, x30, v10, v20, v30, op);
    find_intervals(op, 4, &b1, 1);
    if (b1.n == 0) return 0;

    for (i = 0; i < cop.n; i++)
        for (j = 0; j < a0.n; j++)
            for (k = 0; k < a1.n; k++)
                for (l = 0; l < b0.n; l++)
                    for (m = 0; m < b1.n; m++)
                    {
                        double L5[5], U5[5];
                        int a, b, ok = 1;
                        L5[0] = cop.l[i]; U5[0] = cop.u[i];
                        L5[1] = a0.l[j]; U5[1] = a0.u[j];
                        L5[2] = a1.l[k]; U5[2] = a1.u[k];
                        L5[3] = b0.l[l]; U5[3] = b0.u[l];
                        L5[4] = b1.l[m]; U5[4] = b1.u[m];
                        for (a = 0; a < 5 && ok; a++)
                            for (b = a + 1; b < 5; b++)
                                if (!overlap2(L5[a], U5[a], L5[b], U5[b])) { ok = 0; break; }
                        if (ok)
                        {
                            double il = 0.0, iu = 1.0;
                            int skip = 0;
                            for (a = 0; a < 5; a++) { il = smax(L5[a], il); iu = smin(U5[a], iu); }
                            for (q = 0; q < par.n; q++)
                                if (overlap2(il, iu, par.l[q], par.u[q])) { skip = 1; break; }
                            if (!skip) { mint = smin(mint, il); col = 1; }
                        }
                    }
    if (col) { *t = mint; return 1; }
    return 0;
}

/* CTCD::vertexEdgeCTCD, src/CTCD.cpp:511-602.  p: q0,q1,q2 start then end */
static int vertex_edge(const double *p, double eta, double *t)
{
    double op[5];
    double minD = eta * eta;
    v3 q0s = ld(p), q1s = ld(p + 3), q2s = ld(p + 6);
    v3 v0 = sub(ld(p + 9), q0s), v1 = sub(ld(p + 12), q1s), v2 = sub(ld(p + 15), q2s);
    ivals colin, e1, e2;
    v3 ab = sub(q2s, q1s), ac = sub(q0s, q1s), cb = sub(q2s, q0s);
    v3 vab = sub(v2, v1), vac = sub(v0, v1), vcb = sub(v2, v0);
    double a, b, c, A, B, C, D, E, F, G, H, I, mint = 1.0;
    int i, j, k, col = 0;
    colin.n = e1.n = e2.n = 0;

    c = dot(ab, ac);
    b = dot(ac, vab) + dot(ab, vac);
    a = dot(vab, vac);
    op[0] = a; op[1] = b; op[2] = c;
    find_intervals(op, 2, &e1, 1);
    if (e1.n == 0) return 0;

    c = dot(ab, cb);
    b = dot(cb, vab) + dot(ab, vcb);
    a = dot(vab, vcb);
    op[0] = a; op[1] = b; op[2] = c;
    find_intervals(op, 2, &e2, 1);
    if (e2.n == 0) return 0;

    A = dot(ab, ab);
    B = 2 * dot(ab, vab);
    C = dot(vab, vab);
    D = dot(ac, ac);
    E = 2 * dot(ac, vac);
    F = dot(vac, vac);
    G = dot(ac, ab);
    H = dot(vab, ac) + dot(vac, ab);
    I = dot(vab, vac);
    op[4] = A * D - G * G - minD * A;
    op[3] = B * D + A * E - 2 * G * H - minD * B;
    op[2] = B * E + A * F + C * D - H * H - 2 * G * I - minD * C;
    op[1] = B * F + C * E - 2 * H * I;
    op[0] = C * F - I * I;
    find_intervals(op, 4, &colin, 0);
    if (colin.n == 0) return 0;

    for (i = 0; i < colin.n; i++)
        for (j = 0; j < e1.n; j++)
            for (k = 0; k < e2.n; k++)
                if (overlap2(colin.l[i], colin.u[i], e1.l[j], e1.u[j]) &&
                    overlap2(colin.l[i], colin.u[i], e2.l[k], e2.u[k]) &&
                    overlap2(e1.l[j], e1.u[j], e2.l[k], e2.u[k]))
                {
                    double il = smax(e2.l[k], smax(e1.l[j], smax(colin.l[i], 0.0)));
                    mint = smin(il, mint);
                    col = 1;
                }
    if (col) { *t = mint; return 1; }
    return 0;
}

/* one-shot checkInterval for vertexVertexCTCD: returns 1 if an interval would be pushed */
static int check_once(double t1, double t2, const double *op)
{
    ivals iv;
    iv.n = 0;
    check_interval(t1, t2, op, 2, &iv, 0);
    return iv.n != 0;
}

/* CTCD::vertexVertexCTCD, src/CTCD.cpp:604-692.  p: q1,q2 start then end */
static int vertex_vertex(const double *p, double eta, double *t)
{
    int roots = 0;
    double min_d = eta * eta;
    double t1 = 0, t2 = 0;
    v3 q1s = ld(p), q2s = ld(p + 3);
    v3 v1 = sub(ld(p + 6), q1s), v2 = sub(ld(p + 9), q2s);
    double op[3];
    double a = dot(v1, v1) + dot(v2, v2) - 2 * dot(v1, v2);
    double b = 2 * (dot(v1, q1s) - dot(v2, q1s) - dot(v1, q2s) + dot(v2, q2s));
    double c = dot(q1s, q1s) + dot(q2s, q2s) - 2 * dot(q1s, q2s) - min_d;
    if (a != 0)
        roots = quad_roots(a, b, c, &t1, &t2);
    else if (b != 0)
    {
        t1 = -c / b;
        roots = 1;
    }
    else
    {
        if (c <= 0) { *t = 0; return 1; }
        return 0;
    }
    op[0] = a; op[1] = b; op[2] = c;
    if (roots == 2)
    {
        if (check_once(0, t1, op)) { *t = 0; return 1; }
        if (check_once(t1, t2, op)) { *t = t1; return 1; }
        if (check_once(t2, 1.0, op)) { *t = t2; return 1; }
        return 0;
    }
    else if (roots == 1)
    {
        if (check_once(0, t1, op)) { *t = 0; return 1; }
        if (check_once(t1, 1.0, op)) { *t = t1; return 1; }
        return 0;
    }
    if (check_once(0, 1.0, op)) { *t = 0; return 1; }
    return 0;
}

void orc_vf_batch(long long n, const double *pts, const double *eta, unsigned char *hit, double *t)
{
    long long i;
    for (i = 0; i < n; i++) { double tt = 0; hit[i] = (unsigned char)vertex_face(pts + 24 * i, eta[i], &tt); if (hit[i]) t[i] = tt; }
}
void orc_ee_batch(long long n, const double *pts, const double *eta, unsigned char *hit, double *t)
{
    long long i;
    for (i = 0; i < n; i++) { double tt = 0; hit[i] = (unsigned char)edge_edge(pts + 24 * i, eta[i], &tt); if (hit[i]) t[i] = tt; }
}
void orc_ve_batch(long long n, const double *pts, const double *eta, unsigned char *hit, double *t)
{
    long long i;
    for (i = 0; i < n; i++) { double tt = 0; hit[i] = (unsigned char)vertex_edge(pts + 18 * i, eta[i], &tt); if (hit[i]) t[i] = tt; }
}
void orc_vv_batch(long long n, const double *pts, const double *eta, unsigned char *hit, double *t)
{
    long long i;
    for (i = 0; i < n; i++) { double tt = 0; hit[i] = (unsigned char)vertex_vertex(pts + 12 * i, eta[i], &tt); if (hit[i]) t[i] = tt; }
}

void orc_vf_polys(const double *p, double eta, double *cubics, double *sextic)
{
    v3 q0s = ld(p), q1s = ld(p + 3), q2s = ld(p + 6), q3s = ld(p + 9);
    v3 v0 = sub(ld(p + 12), q0s), v1 = sub(ld(p + 15), q1s), v2 = sub(ld(p + 18), q2s), v3_ = sub(ld(p + 21), q3s);
    plane_coeffs(sub(q0s, q1s), cross(sub(q3s, q1s), sub(q2s, q1s)), sub(q3s, q1s),
                 sub(v0, v1), cross(sub(v3_, v1), sub(v2, v1)), sub(v3_, v1), cubics);
    plane_coeffs(sub(q0s, q2s), cross(sub(q1s, q2s), sub(q3s, q2s)), sub(q1s, q2s),
                 sub(v0, v2), cross(sub(v1, v2), sub(v3_, v2)), sub(v1, v2), cubics + 4);
    plane_coeffs(sub(q0s, q3s), cross(sub(q2s, q3s), sub(q1s, q3s)), sub(q2s, q3s),
                 sub(v0, v3_), cross(sub(v2, v3_), sub(v1, v3_)), sub(v2, v3_), cubics + 8);
    distance_coeffs(sub(q0s, q1s), sub(q2s, q1s), sub(q3s, q1s), sub(v0, v1), sub(v2, v1), sub(v3_, v1), eta * eta, sextic);
}

void orc_ee_polys(const double *p, double eta, double *sextic, double *quartics)
{
    v3 q0s = ld(p), p0s = ld(p + 3), q1s = ld(p + 6), p1s = ld(p + 9);
    v3 vq0 = sub(ld(p + 12), q0s), vp0 = sub(ld(p + 15), p0s), vq1 = sub(ld(p + 18), q1s), vp1 = sub(ld(p + 21), p1s);
    distance_coeffs(sub(p0s, p1s), sub(p0s, q0s), sub(p1s, q1s), sub(vp0, vp1), sub(vp0, vq0), sub(vp1, vq1), eta * eta, sextic);
    barycentric_coeffs(sub(p1s, q1s), sub(p0s, q0s), sub(q0s, q1s), sub(vp1, vq1), sub(vp0, vq0), sub(vq0, vq1), quartics);
    barycentric_coeffs(sub(p1s, q1s), sub(q0s, p0s), sub(p0s, q1s), sub(vp1, vq1), sub(vq0, vp0), sub(vp0, vq1), quartics + 5);
    barycentric_coeffs(sub(p0s, q0s), sub(p1s, q1s), sub(q1s, q0s), sub(vp0, vq0), sub(vp1, vq1), sub(vq1, vq0), quartics + 10);
    barycentric_coeffs(sub(p0s, q0s), sub(q1s, p1s), sub(p1s, q0s), sub(vp0, vq0), sub(vq1, vp1), sub(vp1, vq0), quartics + 15);
}

/* ------------------------------------------------------------------------------------------
 * History::stitchCommonHistory for four vertices on CSR arrays  (src/History.cpp:98-140)
 * ---------------------------------------------------------------------------------------- */
/* Writes stitched positions segment by segment through the callback-free pattern used below:
 * caller iterates with stitch_begin / stitch_next. */
typedef struct
{
    const long long *hoff; const double *htime; const double *hpos;
    int verts[4];
    long long it[4];
    double curtime;
    double postime; /* History time of the entry the last stitch_next produced */
} stitcher;

static void stitch_begin(stitcher *s, const long long *hoff, const double *htime, const double *hpos, const int *verts)
{
    int i;
    s->hoff = hoff; s->htime = htime; s->hpos = hpos;
    for (i = 0; i < 4; i++) { s->verts[i] = verts[i]; s->it[i] = hoff[verts[i]]; }
    s->curtime = 0;
}

/* produces the next stitched entry (positions of the 4 vertices, xyz each); returns 0 when done */
static int stitch_next(stitcher *s, double *pos12)
{
    double newtime = INFINITY;
    int i;
    if (!(s->curtime <= 1.0))
        return 0;
    s->postime = s->curtime;
    for (i = 0; i < 4; i++)
    {
        long long end = s->hoff[s->verts[i] + 1];
        long long next = s->it[i] + 1;
        v3 oldpos;
        while (next != end && s->htime[next] <= s->curtime)
        {
            ++next;
            ++s->it[i];
        }
        oldpos = ld(s->hpos + 3 * s->it[i]);
        if (next == end)
        {
            pos12[3 * i] = oldpos.x; pos12[3 * i + 1] = oldpos.y; pos12[3 * i + 2] = oldpos.z;
        }
        else
        {
            v3 newpos = ld(s->hpos + 3 * next);
            double dt = s->htime[next] - s->htime[s->it[i]];
            double a = s->curtime - s->htime[s->it[i]];
            double b = s->htime[next] - s->curtime;
            v3 r = add(scl(b, oldpos), scl(a, newpos));
            pos12[3 * i] = r.x / dt; pos12[3 * i + 1] = r.y / dt; pos12[3 * i + 2] = r.z / dt;
            newtime = smin(newtime, s->htime[next]);
        }
    }
    s->curtime = newtime;
    return 1;
}

/* CTCDNarrowPhase::checkVFS, src/CTCDNarrowPhase.cpp:24-72 (TOI/stage kept, see ref_harness.cpp; the TOI is mapped from the
 * stitched segment's own parameter back to History time: ta + t (tb - ta), which is t itself for a single step) */
static int check_vfs(const long long *hoff, const double *htime, const double *hpos, const int *s, double eta,
                     double *toi, int *stage)
{
    stitcher st;
    double a[12], b[12], p[24];
    int e, v, c;
    double ta, tb;
    stitch_begin(&st, hoff, htime, hpos, s);
    if (!stitch_next(&st, a))
        return 0;
    ta = st.postime;
    while (stitch_next(&st, b))
    {
        double t;
        tb = st.postime;
        memcpy(p, a, sizeof(a));
        memcpy(p + 12, b, sizeof(b));
        if (vertex_face(p, eta, &t)) { *toi = ta + t * (tb - ta); *stage = 1; return 1; }
        for (e = 0; e < 3; e++)
        {
            int i1 = 1 + (e % 3), i2 = 1 + ((e + 1) % 3);
            double q[18];
            for (c = 0; c < 3; c++)
            {
                q[c] = a[c]; q[3 + c] = a[3 * i1 + c]; q[6 + c] = a[3 * i2 + c];
                q[9 + c] = b[c]; q[12 + c] = b[3 * i1 + c]; q[15 + c] = b[3 * i2 + c];
            }
            if (vertex_edge(q, eta, &t)) { *toi = ta + t * (tb - ta); *stage = 2 + e; return 1; }
        }
        for (v = 0; v < 3; v++)
        {
            double q[12];
            for (c = 0; c < 3; c++)
            {
                q[c] = a[c]; q[3 + c] = a[3 * (1 + v) + c];
                q[6 + c] = b[c]; q[9 + c] = b[3 * (1 + v) + c];
            }
            if (vertex_vertex(q, eta, &t)) { *toi = ta + t * (tb - ta); *stage = 5 + v; return 1; }
        }
        memcpy(a, b, sizeof(a));
        ta = tb;
    }
    return 0;
}

/* CTCDNarrowPhase::checkEES, src/CTCDNarrowPhase.cpp:74-135 */
static int check_ees(const long long *hoff, const double *htime, const double *hpos, const int *s, double eta,
                     double *toi, int *stage)
{
    static const int ve[4][3] = {{0, 2, 3}, {1, 2, 3}, {2, 0, 1}, {3, 0, 1}};
    static const int vv[4][2] = {{0, 2}, {0, 3}, {1, 2}, {1, 3}};
    stitcher st;
    double a[12], b[12], p[24];
    int e, v, c;
    double ta, tb;
    stitch_begin(&st, hoff, htime, hpos, s);
    if (!stitch_next(&st, a))
        return 0;
    ta = st.postime;
    while (stitch_next(&st, b))
    {
        double t;
        tb = st.postime;
        memcpy(p, a, sizeof(a));
        memcpy(p + 12, b, sizeof(b));
        if (edge_edge(p, eta, &t)) { *toi = ta + t * (tb - ta); *stage = 1; return 1; }
        for (e = 0; e < 4; e++)
        {
            double q[18];
            for (c = 0; c < 3; c++)
            {
                q[c] = a[3 * ve[e][0] + c]; q[3 + c] = a[3 * ve[e][1] + c]; q[6 + c] = a[3 * ve[e][2] + c];
                q[9 + c] = b[3 * ve[e][0] + c]; q[12 + c] = b[3 * ve[e][1] + c]; q[15 + c] = b[3 * ve[e][2] + c];
            }
            if (vertex_edge(q, eta, &t)) { *toi = ta + t * (tb - ta); *stage = 2 + e; return 1; }
        }
        for (v = 0; v < 4; v++)
        {
            double q[12];
            for (c = 0; c < 3; c++)
            {
                q[c] = a[3 * vv[v][0] + c]; q[3 + c] = a[3 * vv[v][1] + c];
                q[6 + c] = b[3 * vv[v][0] + c]; q[9 + c] = b[3 * vv[v][1] + c];
            }
            if (vertex_vertex(q, eta, &t)) { *toi = ta + t * (tb - ta); *stage = 6 + v; return 1; }
        }
        memcpy(a, b, sizeof(a));
        ta = tb;
    }
    return 0;
}

void orc_narrowphase_flat(int V, const long long *hoff, const double *htime, const double *hpos,
                          long long nvf, const int *vf, const double *vf_eta, long long nee, const int *ee,
                          const double *ee_eta, unsigned char *vf_hit, double *vf_toi, unsigned char *ee_hit,
                          double *ee_toi)
{
    long long i;
    (void)V;
    for (i = 0; i < nvf; i++)
    {
        double t = 0; int st = 0;
        vf_hit[i] = (unsigned char)check_vfs(hoff, htime, hpos, vf + 4 * i, vf_eta[i], &t, &st);
        vf_toi[i] = vf_hit[i] ? t : 0.0;
    }
    for (i = 0; i < nee; i++)
    {
        double t = 0; int st = 0;
        ee_hit[i] = (unsigned char)check_ees(hoff, htime, hpos, ee + 4 * i, ee_eta[i], &t, &st);
        ee_toi[i] = ee_hit[i] ? t : 0.0;
    }
}

int orc_narrowphase_sepplane(int V, const long long *hoff, const double *htime, const double *hpos, long long nvf, const int *vf,
                             const double *vf_eta, long long nee, const int *ee, const double *ee_eta, unsigned char *vf_hit,
                             unsigned char *ee_hit);

/* CTCDNarrowPhase::findCollisions, src/CTCDNarrowPhase.cpp:9-22 (flat arrays instead of sets) */
int orc_narrowphase(int which, int V, const long long *hoff, const double *htime, const double *hpos,
                    long long nvf, const int *vf, const double *vf_eta, long long nee, const int *ee,
                    const double *ee_eta, unsigned char *vf_hit, double *vf_toi, int *vf_stage,
                    unsigned char *ee_hit, double *ee_toi, int *ee_stage, double *seconds)
{
    long long i;
    double t0 = now_s();
    (void)V;
    if (which != 0)
    {
        /* SeparatingPlaneNarrowPhase: flags only (no TOI / stage) */
        int rc = orc_narrowphase_sepplane(V, hoff, htime, hpos, nvf, vf, vf_eta, nee, ee, ee_eta, vf_hit, ee_hit);
        if (seconds)
            *seconds = now_s() - t0;
        return rc;
    }
    for (i = 0; i < nvf; i++)
    {
        double t = 0; int st = 0;
        vf_hit[i] = (unsigned char)check_vfs(hoff, htime, hpos, vf + 4 * i, vf_eta[i], &t, &st);
        if (vf_toi) vf_toi[i] = vf_hit[i] ? t : 0.0;
        if (vf_stage) vf_stage[i] = st;
    }
    for (i = 0; i < nee; i++)
    {
        double t = 0; int st = 0;
        ee_hit[i] = (unsigned char)check_ees(hoff, htime, hpos, ee + 4 * i, ee_eta[i], &t, &st);
        if (ee_toi) ee_toi[i] = ee_hit[i] ? t : 0.0;
        if (ee_stage) ee_stage[i] = st;
    }
    if (seconds)
        *seconds = now_s() - t0;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Broadphase  (src/KDOPBroadPhase.cpp:10-185, src/AABBBroadPhase.cpp:9-152, src/Stencils.h)
 * ---------------------------------------------------------------------------------------- */
static void kdop_axes(double ax[13][3])
{
    /* src/KDOPBroadPhase.cpp:12-31: direction then "/= norm()" per component */
    static const double raw[13][3] = {
        {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {1, -1, 0}, {0, 1, 1}, {0, 1, -1},
        {1, 0, 1}, {1, 0, -1}, {1, 1, 1}, {1, 1, -1}, {1, -1, 1}, {1, -1, -1}};
    int i, c;
    for (i = 0; i < 13; i++)
    {
        double nrm = sqrt((raw[i][0] * raw[i][0] + raw[i][1] * raw[i][1]) + raw[i][2] * raw[i][2]);
        for (c = 0; c < 3; c++)
            ax[i][c] = raw[i][c] / nrm;
    }
}

/* leaf boxes: src/KDOPBroadPhase.cpp:45-72 (K=13) and src/AABBBroadPhase.cpp:21-44 (K=3) */
void orc_leaf_boxes(int kind, int F, const int *faces, const long long *hoff, const double *hpos,
                    double outerEta, double *boxes)
{
    double ax[13][3];
    int f, j, k;
    kdop_axes(ax);
    for (f = 0; f < F; f++)
    {
        double *mins = boxes + (size_t)f * 2 * kind, *maxs = mins + kind;
        for (k = 0; k < kind; k++) { mins[k] = INFINITY; maxs[k] = -INFINITY; }
        for (j = 0; j < 3; j++)
        {
            int v = faces[3 * f + j];
            long long e;
            for (e = hoff[v]; e < hoff[v + 1]; e++)
            {
                v3 pos = ld(hpos + 3 * e);
                for (k = 0; k < kind; k++)
                {
                    double d = (kind == 13) ? dot(pos, mk(ax[k][0], ax[k][1], ax[k][2])) : hpos[3 * e + k];
                    mins[k] = smin(d - outerEta, mins[k]);
                    maxs[k] = smax(d + outerEta, maxs[k]);
                }
            }
        }
    }
}

typedef struct { int *d; long long n, cap; } ivec;
static void ivec_push4(ivec *v, int a, int b, int c, int d)
{
    if (v->n + 4 > v->cap)
    {
        v->cap = v->cap ? 2 * v->cap : 4096;
        v->d = (int *)realloc(v->d, sizeof(int) * (size_t)v->cap);
    }
    v->d[v->n++] = a; v->d[v->n++] = b; v->d[v->n++] = c; v->d[v->n++] = d;
}

/* VertexFaceStencil ctor, src/Stencils.h:13-57: face indices ascending */
static void push_vfs(ivec *v, int p, int a, int b, int c)
{
    int t;
    if (a > b) { t = a; a = b; b = t; }
    if (b > c) { t = b; b = c; c = t; }
    if (a > b) { t = a; a = b; b = t; }
    ivec_push4(v, p, a, b, c);
}

/* EdgeEdgeStencil ctor, src/Stencils.h:86-117 */
static void push_ees(ivec *v, int p0, int p1, int q0, int q1)
{
    int t;
    if (p0 > p1) { t = p0; p0 = p1; p1 = t; }
    if (q0 > q1) { t = q0; q0 = q1; q1 = t; }
    if (p0 > q0) { t = p0; p0 = q0; q0 = t; t = p1; p1 = q1; q1 = t; }
    ivec_push4(v, p0, p1, q0, q1);
}

static int cmp4(const void *a, const void *b)
{
    const int *x = (const int *)a, *y = (const int *)b;
    int i;
    for (i = 0; i < 4; i++)
    {
        if (x[i] < y[i]) return -1;
        if (x[i] > y[i]) return 1;
    }
    return 0;
}

static long long sort_unique4(int *d, long long n4)
{
    long long n = n4 / 4, i, w = 0;
    if (n == 0) return 0;
    qsort(d, (size_t)n, 4 * sizeof(int), cmp4);
    for (i = 1; i < n; i++)
        if (cmp4(d + 4 * w, d + 4 * i) != 0)
        {
            w++;
            if (w != i) memcpy(d + 4 * w, d + 4 * i, 4 * sizeof(int));
        }
    return w + 1;
}

typedef struct
{
    int kind, F;
    const int *faces;
    const double *leaf;     /* F x 2*kind */
    double *node;           /* interior boxes, (F-1) x 2*kind */
    int *left, *right;      /* child ids: >= 0 interior, < 0 leaf ~(face) */
    int nnodes;
    const unsigned char *fixed;
    ivec vf, ee;
    long long vf_limit, ee_limit;   /* compaction thresholds (ints) */
} bvh;

static const double *box_of(const bvh *t, int id)
{
    return id >= 0 ? t->node + (size_t)id * 2 * t->kind : t->leaf + (size_t)(~id) * 2 * t->kind;
}

static int g_axis, g_kind;
static const double *g_leaf;
static int cmp_axis(const void *a, const void *b)
{
    /* NodeComparator, src/KDOPBroadPhase.h:14-26: order by mins[axis] */
    double x = g_leaf[(size_t)(*(const int *)a) * 2 * g_kind + g_axis];
    double y = g_leaf[(size_t)(*(const int *)b) * 2 * g_kind + g_axis];
    return (x < y) ? -1 : (x > y) ? 1 : 0;
}

/* buildKDOPInterior / buildAABBInterior, src/KDOPBroadPhase.cpp:76-128: union box, longest axis,
 * sort by mins, first half left.  (Tree shape does not influence the candidate set.) */
static int build_interior(bvh *t, int *ids, int n)
{
    int K = t->kind, id, i, k, greatest = -1;
    double *box, greatestlen = 0;
    if (n == 1)
        return ~ids[0];
    id = t->nnodes++;
    box = t->node + (size_t)id * 2 * K;
    for (k = 0; k < K; k++) { box[k] = INFINITY; box[K + k] = -INFINITY; }
    for (i = 0; i < n; i++)
    {
        const double *lb = t->leaf + (size_t)ids[i] * 2 * K;
        for (k = 0; k < K; k++)
        {
            box[k] = smin(lb[k], box[k]);
            box[K + k] = smax(lb[K + k], box[K + k]);
        }
    }
    for (k = 0; k < K; k++)
        if (box[K + k] - box[k] > greatestlen) { greatestlen = box[K + k] - box[k]; greatest = k; }
    if (greatest >= 0)
    {
        g_axis = greatest; g_kind = K; g_leaf = t->leaf;
        qsort(ids, (size_t)n, sizeof(int), cmp_axis);
    }
    t->left[id] = build_interior(t, ids, n / 2);
    t->right[id] = build_interior(t, ids + n / 2, n - n / 2);
    return id;
}

static int is_fixed(const bvh *t, int v) { return t->fixed && t->fixed[v]; }

/* KDOPBroadPhase::intersect, src/KDOPBroadPhase.cpp:130-185 */
static void intersect(bvh *t, int L, int R)
{
    const double *a = box_of(t, L), *b = box_of(t, R);
    int K = t->kind, axis;
    for (axis = 0; axis < K; axis++)
        if (a[K + axis] < b[axis] || b[K + axis] < a[axis])
            return;
    if (L >= 0)
    {
        intersect(t, t->left[L], R);
        intersect(t, t->right[L], R);
    }
    else if (R >= 0)
    {
        intersect(t, L, t->left[R]);
        intersect(t, L, t->right[R]);
    }
    else
    {
        const int *fl = t->faces + 3 * (~L), *fr = t->faces + 3 * (~R);
        int i, j;
        for (i = 0; i < 3; i++)           /* Mesh::neighboringFaces, src/Mesh.cpp:13-20 */
            for (j = 0; j < 3; j++)
                if (fl[i] == fr[j])
                    return;
        for (i = 0; i < 3; i++)
        {
            int alll = is_fixed(t, fl[i]), allr = is_fixed(t, fr[i]);
            for (j = 0; j < 3; j++)
            {
                alll = alll && is_fixed(t, fr[j]);
                allr = allr && is_fixed(t, fl[j]);
            }
            if (!alll) push_vfs(&t->vf, fl[i], fr[0], fr[1], fr[2]);
            if (!allr) push_vfs(&t->vf, fr[i], fl[0], fl[1], fl[2]);
            for (j = 0; j < 3; j++)
            {
                int alle = is_fixed(t, fl[i]) && is_fixed(t, fl[(i + 1) % 3]) && is_fixed(t, fr[j]) && is_fixed(t, fr[(j + 1) % 3]);
                if (!alle) push_ees(&t->ee, fl[i], fl[(i + 1) % 3], fr[j], fr[(j + 1) % 3]);
            }
        }
        /* keep memory bounded: compact now and then (std::set semantics = unique) */
        if (t->vf.n > t->vf_limit)
        {
            t->vf.n = 4 * sort_unique4(t->vf.d, t->vf.n);
            if (2 * t->vf.n > t->vf_limit) t->vf_limit = 2 * t->vf.n;
        }
        if (t->ee.n > t->ee_limit)
        {
            t->ee.n = 4 * sort_unique4(t->ee.d, t->ee.n);
            if (2 * t->ee.n > t->ee_limit) t->ee_limit = 2 * t->ee.n;
        }
    }
}

int orc_broadphase(int kind, int V, int F, const int *faces, const long long *hoff, const double *htime,
                   const double *hpos, double outerEta, const unsigned char *fixedMask, int **vf_out,
                   long long *nvf, int **ee_out, long long *nee, double *seconds)
{
    bvh t;
    double *leaf;
    int *ids, i, root;
    double t0 = now_s();
    (void)V; (void)htime;
    if (kind != 13 && kind != 3)
        return -1;
    memset(&t, 0, sizeof(t));
    leaf = (double *)malloc(sizeof(double) * 2 * (size_t)kind * (size_t)(F > 0 ? F : 1));
    orc_leaf_boxes(kind, F, faces, hoff, hpos, outerEta, leaf);
    t.kind = kind; t.F = F; t.faces = faces; t.leaf = leaf; t.fixed = fixedMask;
    t.vf_limit = t.ee_limit = 1LL << 25;
    t.node = (double *)malloc(sizeof(double) * 2 * (size_t)kind * (size_t)(F > 1 ? F - 1 : 1));
    t.left = (int *)malloc(sizeof(int) * (size_t)(F > 1 ? F - 1 : 1));
    t.right = (int *)malloc(sizeof(int) * (size_t)(F > 1 ? F - 1 : 1));
    ids = (int *)malloc(sizeof(int) * (size_t)(F > 0 ? F : 1));
    for (i = 0; i < F; i++) ids[i] = i;
    if (F > 0)
    {
        root = build_interior(&t, ids, F);
        intersect(&t, root, root);
    }
    *nvf = sort_unique4(t.vf.d, t.vf.n);
    *nee = sort_unique4(t.ee.d, t.ee.n);
    if (!t.vf.d) t.vf.d = (int *)malloc(16);
    if (!t.ee.d) t.ee.d = (int *)malloc(16);
    *vf_out = t.vf.d;
    *ee_out = t.ee.d;
    free(leaf); free(t.node); free(t.left); free(t.right); free(ids);
    if (seconds)
        *seconds = now_s() - t0;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Distance queries  (include/Distance.h:14-185, src/Distance.cpp:12-66)
 * ---------------------------------------------------------------------------------------- */
static double clamp01(double u) { return smin(1.0, smax(u, 0.0)); }

/* Distance::vertexPlaneDistanceLessThan, include/Distance.h:14-19 */
static int plane_lt(v3 p, v3 q0, v3 q1, v3 q2, double eta)
{
    v3 c = cross(sub(q1, q0), sub(q2, q0));
    return dot(c, sub(p, q0)) * dot(c, sub(p, q0)) < eta * eta * dot(c, c);
}

/* Distance::lineLineDistanceLessThan, include/Distance.h:23-28 */
static int line_lt(v3 p0, v3 p1, v3 q0, v3 q1, double eta)
{
    v3 c = cross(sub(p1, p0), sub(q1, q0));
    return dot(c, sub(q0, p0)) * dot(c, sub(q0, p0)) < eta * eta * dot(c, c);
}

/* Distance::vertexFaceDistance, include/Distance.h:32-115 */
static v3 dist_vf(v3 p, v3 q0, v3 q1, v3 q2, double *b0, double *b1, double *b2)
{
    v3 ab = sub(q1, q0), ac = sub(q2, q0), ap = sub(p, q0), bp, cp;
    double d1 = dot(ab, ap), d2 = dot(ac, ap), d3, d4, d5, d6, vc, vb, va, denom, v, w, u;
    if (d1 <= 0 && d2 <= 0) { *b0 = 1.0; *b1 = 0.0; *b2 = 0.0; return sub(q0, p); }
    bp = sub(p, q1);
    d3 = dot(ab, bp); d4 = dot(ac, bp);
    if (d3 >= 0 && d4 <= d3) { *b0 = 0.0; *b1 = 1.0; *b2 = 0.0; return sub(q1, p); }
    vc = d1 * d4 - d3 * d2;
    if ((vc <= 0) && (d1 >= 0) && (d3 <= 0))
    {
        v = d1 / (d1 - d3);
        *b0 = 1.0 - v; *b1 = v; *b2 = 0;
        return sub(add(q0, scl(v, ab)), p);
    }
    cp = sub(p, q2);
    d5 = dot(ab, cp); d6 = dot(ac, cp);
    if (d6 >= 0 && d5 <= d6) { *b0 = 0; *b1 = 0; *b2 = 1.0; return sub(q2, p); }
    vb = d5 * d2 - d1 * d6;
    if ((vb <= 0) && (d2 >= 0) && (d6 <= 0))
    {
        w = d2 / (d2 - d6);
        *b0 = 1 - w; *b1 = 0; *b2 = w;
        return sub(add(q0, scl(w, ac)), p);
    }
    va = d3 * d6 - d5 * d4;
    if ((va <= 0) && (d4 - d3 >= 0) && (d5 - d6 >= 0))
    {
        w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        *b0 = 0; *b1 = 1.0 - w; *b2 = w;
        return sub(add(q1, scl(w, sub(q2, q1))), p);
    }
    denom = 1.0 / (va + vb + vc);
    v = vb * denom;
    w = vc * denom;
    u = 1.0 - v - w;
    *b0 = u; *b1 = v; *b2 = w;
    return sub(add(add(scl(u, q0), scl(v, q1)), scl(w, q2)), p);
}

/* Distance::edgeEdgeDistance, include/Distance.h:119-174 */
static v3 dist_ee(v3 p0, v3 p1, v3 q0, v3 q1, double *bp0, double *bp1, double *bq0, double *bq1)
{
    v3 d1 = sub(p1, p0), d2 = sub(q1, q0), r = sub(p0, q0), c1, c2;
    double a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r), s, t;
    double c = dot(d1, r), b = dot(d1, d2), denom = a * e - b * b, tnom;
    if (denom != 0.0)
        s = clamp01((b * f - c * e) / denom);
    else
        s = 0;
    tnom = b * s + f;
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
    c1 = add(p0, scl(s, d1));
    c2 = add(q0, scl(t, d2));
    *bp0 = 1.0 - s; *bp1 = s; *bq0 = 1.0 - t; *bq1 = t;
    return sub(c2, c1);
}

/* ------------------------------------------------------------------------------------------
 * PenaltyGroup::addForce (src/PenaltyGroup.cpp:34-52): the group's vertex-face stencils, then its edge-edge stencils,
 * each through its potential (src/PenaltyPotential.cpp:7-64); F += groupforce * dt.  Sequential, like the reference.
 * fired (may be NULL): nvf + nee flags.  Returns the reference's return value (a stencil that is new fired).
 * ---------------------------------------------------------------------------------------- */
int orc_penalty_group_force(int V, const double *q, const double *v, long long nvf, const int *vf, const unsigned char *vf_isnew,
                            long long nee, const int *ee, const unsigned char *ee_isnew, double dt, double outerEta, double innerEta,
                            double stiffness0, double CoR, double *F, unsigned char *fired, long long *n_fired)
{
    double *g = (double *)calloc(3 * (size_t)V + 1, sizeof(double));
    long long i, nf = 0;
    int newused = 0, k, c;
    for (i = 0; i < nvf + nee; i++)
    {
        const int is_vf = i < nvf;
        const int *s = is_vf ? vf + 4 * i : ee + 4 * (i - nvf);
        v3 x[4], vel[4], cv, relvel, localF, t;
        double w[4] = {0, 0, 0, 0}, dist, stiffness = stiffness0, sc;
        if (fired) fired[i] = 0;
        for (k = 0; k < 4; k++) { x[k] = ld(q + 3 * (size_t)s[k]); vel[k] = ld(v + 3 * (size_t)s[k]); }
        if (!(is_vf ? plane_lt(x[0], x[1], x[2], x[3], outerEta) : line_lt(x[0], x[1], x[2], x[3], outerEta)))
            continue;                                                            /* PenaltyPotential.cpp:14-15 / :44-45 */
        if (is_vf) cv = dist_vf(x[0], x[1], x[2], x[3], &w[1], &w[2], &w[3]);
        else cv = dist_ee(x[0], x[1], x[2], x[3], &w[0], &w[1], &w[2], &w[3]);
        dist = sqrt(dot(cv, cv));
        if (dist >= outerEta || dist < innerEta)
            continue;                                                            /* :19-21 / :50-52 */
        if (is_vf)                                                               /* :23 */
            relvel = add(add(add(mk(-vel[0].x, -vel[0].y, -vel[0].z), scl(w[1], vel[1])), scl(w[2], vel[2])), scl(w[3], vel[3]));
        else                                                                     /* :54 */
            relvel = add(add(sub(scl(-w[0], vel[0]), scl(w[1], vel[1])), scl(w[2], vel[2])), scl(w[3], vel[3]));
        if (dot(relvel, cv) > 0)
            stiffness *= CoR;
        sc = stiffness * (outerEta - dist) / (outerEta - innerEta);             /* :27 / :58 */
        t = scl(sc, cv);
        localF = mk(t.x / dist, t.y / dist, t.z / dist);
        for (k = 0; k < 4; k++)
        {
            /* VF: F[p] -= localF, F[q_k] += bary_k localF;  EE: F[p_k] -= baryp_k localF, F[q_k] += baryq_k localF */
            const int minus = is_vf ? k == 0 : k < 2;
            const v3 f = (is_vf && k == 0) ? localF : scl(w[k], localF);
            double *gv = g + 3 * (size_t)s[k];
            const double fc[3] = {f.x, f.y, f.z};
            for (c = 0; c < 3; c++) gv[c] = minus ? gv[c] - fc[c] : gv[c] + fc[c];
        }
        if (fired) fired[i] = 1;
        nf++;
        if (is_vf ? (vf_isnew ? vf_isnew[i] : 1) : (ee_isnew ? ee_isnew[i - nvf] : 1)) newused = 1;
    }
    for (i = 0; i < 3 * (long long)V; i++) F[i] = F[i] + g[i] * dt;              /* PenaltyGroup.cpp:50 */
    free(g);
    if (n_fired) *n_fired = nf;
    return newused;
}

void orc_dist_vf_batch(long long n, const double *pts, double *vec, double *bary)
{
    long long i;
    for (i = 0; i < n; i++)
    {
        const double *p = pts + 12 * i;
        v3 r = dist_vf(ld(p), ld(p + 3), ld(p + 6), ld(p + 9), bary + 3 * i, bary + 3 * i + 1, bary + 3 * i + 2);
        vec[3 * i] = r.x; vec[3 * i + 1] = r.y; vec[3 * i + 2] = r.z;
    }
}
void orc_dist_ee_batch(long long n, const double *pts, double *vec, double *bary)
{
    long long i;
    for (i = 0; i < n; i++)
    {
        const double *p = pts + 12 * i;
        v3 r = dist_ee(ld(p), ld(p + 3), ld(p + 6), ld(p + 9), bary + 4 * i, bary + 4 * i + 1, bary + 4 * i + 2, bary + 4 * i + 3);
        vec[3 * i] = r.x; vec[3 * i + 1] = r.y; vec[3 * i + 2] = r.z;
    }
}
void orc_dist_plane_lt_batch(long long n, const double *pts, const double *eta, unsigned char *out)
{
    long long i;
    for (i = 0; i < n; i++)
    {
        const double *p = pts + 12 * i;
        out[i] = (unsigned char)plane_lt(ld(p), ld(p + 3), ld(p + 6), ld(p + 9), eta[i]);
    }
}
void orc_dist_line_lt_batch(long long n, const double *pts, const double *eta, unsigned char *out)
{
    long long i;
    for (i = 0; i < n; i++)
    {
        const double *p = pts + 12 * i;
        out[i] = (unsigned char)line_lt(ld(p), ld(p + 3), ld(p + 6), ld(p + 9), eta[i]);
    }
}

/* Distance::meshSelfDistance, src/Distance.cpp:12-66 */
double orc_mesh_self_distance(int V, const double *verts, int F, const int *faces,
                              const unsigned char *fixedMask, double *seconds)
{
    double closest = INFINITY, t0 = now_s();
    long long *hoff, nvf = 0, nee = 0, s;
    double *htime, *hpos;
    int *vf = 0, *ee = 0, i, j, k;
    for (i = 0; i < V; i++)
        for (j = 0; j < F; j++)
        {
            const int *f = faces + 3 * j;
            if (f[0] == i || f[1] == i || f[2] == i)      /* Mesh::vertexOfFace, src/Mesh.cpp:5-11 */
                continue;
            for (k = 0; k < 3; k++)
            {
                v3 d = sub(ld(verts + 3 * i), ld(verts + 3 * f[k]));
                double dist = dot(d, d);
                if (dist < closest)
                    closest = dist;
            }
        }
    closest = sqrt(closest);

    /* static history q -> q, AABB broadphase with outerEta = closest (src/Distance.cpp:37-45) */
    hoff = (long long *)malloc(sizeof(long long) * ((size_t)V + 1));
    htime = (double *)malloc(sizeof(double) * 2 * (size_t)(V > 0 ? V : 1));
    hpos = (double *)malloc(sizeof(double) * 6 * (size_t)(V > 0 ? V : 1));
    for (i = 0; i <= V; i++) hoff[i] = 2LL * i;
    for (i = 0; i < V; i++)
    {
        htime[2 * i] = 0; htime[2 * i + 1] = 1.0;
        for (k = 0; k < 3; k++) { hpos[6 * i + k] = verts[3 * i + k]; hpos[6 * i + 3 + k] = verts[3 * i + k]; }
    }
    orc_broadphase(3, V, F, faces, hoff, htime, hpos, closest, fixedMask, &vf, &nvf, &ee, &nee, 0);
    free(hoff); free(htime); free(hpos);

    closest = INFINITY;
    for (s = 0; s < nvf; s++)
    {
        double b0, b1, b2, dist;
        v3 r = dist_vf(ld(verts + 3 * vf[4 * s]), ld(verts + 3 * vf[4 * s + 1]), ld(verts + 3 * vf[4 * s + 2]), ld(verts + 3 * vf[4 * s + 3]), &b0, &b1, &b2);
        dist = sqrt(dot(r, r));
        if (dist < closest) closest = dist;
    }
    for (s = 0; s < nee; s++)
    {
        double b0, b1, b2, b3, dist;
        v3 r = dist_ee(ld(verts + 3 * ee[4 * s]), ld(verts + 3 * ee[4 * s + 1]), ld(verts + 3 * ee[4 * s + 2]), ld(verts + 3 * ee[4 * s + 3]), &b0, &b1, &b2, &b3);
        dist = sqrt(dot(r, r));
        if (dist < closest) closest = dist;
    }
    free(vf); free(ee);
    if (seconds)
        *seconds = now_s() - t0;
    return closest;
}

/* ------------------------------------------------------------------------------------------
 * SeparatingPlaneNarrowPhase  (src/SeparatingPlaneNarrowPhase.cpp:11-278), recursive like the reference
 * ---------------------------------------------------------------------------------------- */
typedef struct
{
    const long long *hoff;
    const double *htime;
    const double *hpos;
} hist_t;

/* History::getPosAtTime, src/History.cpp:49-80 (the bisection there only moves the start of this scan) */
static void sp_pos_at(const hist_t *H, int vert, double time, v3 *pos, int *idx)
{
    const long long b = H->hoff[vert], e = H->hoff[vert + 1];
    long long next = b, prev;
    while (next < e && H->htime[next] <= time)
        next++;
    prev = next - 1;
    if (next == e)
    {
        *pos = ld(H->hpos + 3 * prev);
        *idx = (int)(prev - b);
        return;
    }
    {
        double dt = H->htime[next] - H->htime[prev];
        double alpha = (time - H->htime[prev]) / dt;
        *pos = add(scl(1.0 - alpha, ld(H->hpos + 3 * prev)), scl(alpha, ld(H->hpos + 3 * next)));
        *idx = (int)(prev - b);
    }
}

/* planeIntersect, :266-276 */
static double sp_plane_intersect(v3 planePos, v3 planeVel, v3 planeNormal, v3 ptOld, v3 ptNew, double ptdt, double eta)
{
    double numerator = 0.5 * eta - dot(sub(ptOld, planePos), planeNormal);
    double denom = -dot(planeVel, planeNormal);
    if (ptdt != 0.0)
        denom += dot(sub(ptNew, ptOld), planeNormal) / ptdt;
    return numerator / denom;
}

/* planeTrajectoryIntersect, :220-264 */
static double sp_trajectory(const hist_t *H, int vert, int startidx, int forward, v3 planePos, v3 planeVel, v3 planeNormal,
                            v3 ptstart, double timestart, double eta)
{
    const long long b = H->hoff[vert];
    const int size = (int)(H->hoff[vert + 1] - b);
    const double *ht = H->htime + b, *hp = H->hpos + 3 * b;
    int i;
    {
        double dt = forward ? ht[startidx] - timestart : timestart - ht[startidx];
        double t = sp_plane_intersect(planePos, scl(forward ? 1.0 : -1.0, planeVel), planeNormal, ptstart, ld(hp + 3 * startidx), dt, eta);
        if (t >= 0 && t <= dt)
            return t;
    }
    if (forward)
    {
        for (i = startidx; i < size; i++)
        {
            double dt, t;
            if (i == size - 1)
                return INFINITY;
            dt = ht[i + 1] - ht[i];
            t = sp_plane_intersect(add(planePos, scl(ht[i] - timestart, planeVel)), planeVel, planeNormal, ld(hp + 3 * i), ld(hp + 3 * (i + 1)), dt, eta);
            if (t >= 0 && t <= dt)
                return ht[i] + t - timestart;
        }
    }
    else
    {
        for (i = startidx; i >= 0; i--)
        {
            double dt, t;
            if (i == 0)
                return INFINITY;
            dt = ht[i] - ht[i - 1];
            t = sp_plane_intersect(sub(planePos, scl(timestart - ht[i], planeVel)), scl(-1.0, planeVel), planeNormal, ld(hp + 3 * i), ld(hp + 3 * (i - 1)), dt, eta);
            if (t >= 0 && t <= dt)
                return timestart - (ht[i] - t);
        }
    }
    return INFINITY;
}

/* the CTCD tests of a short interval inside one History segment, :76-140 */
static int sp_segment(int is_vf, const v3 *oldpos, const v3 *newpos, double eta)
{
    static const int ve_ee[4][3] = {{0, 2, 3}, {1, 2, 3}, {2, 0, 1}, {3, 0, 1}};
    static const int vv_ee[4][2] = {{0, 2}, {0, 3}, {1, 2}, {1, 3}};
    double p[24], q[18], t;
    int i, e, v;
    for (i = 0; i < 4; i++)
    {
        p[3 * i] = oldpos[i].x; p[3 * i + 1] = oldpos[i].y; p[3 * i + 2] = oldpos[i].z;
        p[12 + 3 * i] = newpos[i].x; p[12 + 3 * i + 1] = newpos[i].y; p[12 + 3 * i + 2] = newpos[i].z;
    }
    if (is_vf ? vertex_face(p, eta, &t) : edge_edge(p, eta, &t))
        return 1;
    for (e = 0; e < (is_vf ? 3 : 4); e++)
    {
        int iv = is_vf ? 0 : ve_ee[e][0], i1 = is_vf ? 1 + (e % 3) : ve_ee[e][1], i2 = is_vf ? 1 + ((e + 1) % 3) : ve_ee[e][2];
        memcpy(q, p + 3 * iv, 3 * sizeof(double)); memcpy(q + 3, p + 3 * i1, 3 * sizeof(double)); memcpy(q + 6, p + 3 * i2, 3 * sizeof(double));
        memcpy(q + 9, p + 12 + 3 * iv, 3 * sizeof(double)); memcpy(q + 12, p + 12 + 3 * i1, 3 * sizeof(double)); memcpy(q + 15, p + 12 + 3 * i2, 3 * sizeof(double));
        if (vertex_edge(q, eta, &t))
            return 1;
    }
    for (v = 0; v < (is_vf ? 3 : 4); v++)
    {
        int i1 = is_vf ? 0 : vv_ee[v][0], i2 = is_vf ? 1 + v : vv_ee[v][1];
        memcpy(q, p + 3 * i1, 3 * sizeof(double)); memcpy(q + 3, p + 3 * i2, 3 * sizeof(double));
        memcpy(q + 6, p + 12 + 3 * i1, 3 * sizeof(double)); memcpy(q + 9, p + 12 + 3 * i2, 3 * sizeof(double));
        if (vertex_vertex(q, eta, &t))
            return 1;
    }
    return 0;
}

/* checkInterval, :47-218 */
static int sp_check(int is_vf, const hist_t *H, const int *verts, double eta, double mint, double maxt, double eps)
{
    v3 midpos[4], closest, planepos, planevel, term[4], wpos[4];
    int mididx[4], i, vert;
    double midt, b1, b2, b3, b4 = 0, w[4], t;
    if (maxt < mint)
        return 0;
    if (maxt - mint < 3 * eps)
    {
        int ok = 1;
        v3 oldpos[4], newpos[4];
        for (i = 0; i < 4; i++)
        {
            int oldidx, newidx;
            sp_pos_at(H, verts[i], mint, &oldpos[i], &oldidx);
            sp_pos_at(H, verts[i], maxt, &newpos[i], &newidx);
            if (oldidx != newidx) { ok = 0; break; }
        }
        if (ok)
            return sp_segment(is_vf, oldpos, newpos, eta);
    }
    midt = 0.5 * (mint + maxt);
    for (i = 0; i < 4; i++)
        sp_pos_at(H, verts[i], midt, &midpos[i], &mididx[i]);
    if (is_vf)
        closest = dist_vf(midpos[0], midpos[1], midpos[2], midpos[3], &b1, &b2, &b3);
    else
        closest = dist_ee(midpos[0], midpos[1], midpos[2], midpos[3], &b1, &b2, &b3, &b4);
    if (dot(closest, closest) < eta * eta)
        return 1;
    /* separating plane, :167-189: bary * (next - prev) / dt term by term */
    w[0] = is_vf ? 1.0 : b1; w[1] = is_vf ? b1 : b2; w[2] = is_vf ? b2 : b3; w[3] = is_vf ? b3 : b4;
    for (i = 0; i < 4; i++)
    {
        const long long e0 = H->hoff[verts[i]] + mididx[i];
        double dts = H->htime[e0 + 1] - H->htime[e0];
        v3 d = sub(ld(H->hpos + 3 * (e0 + 1)), ld(H->hpos + 3 * e0));
        if (!(is_vf && i == 0)) d = scl(w[i], d);
        term[i] = mk(d.x / dts, d.y / dts, d.z / dts);
        wpos[i] = (is_vf && i == 0) ? midpos[0] : scl(w[i], midpos[i]);
    }
    planepos = scl(0.5, add(add(add(wpos[0], wpos[1]), wpos[2]), wpos[3]));
    planevel = scl(0.5, add(add(add(term[0], term[1]), term[2]), term[3]));
    {
        double cn = sqrt(dot(closest, closest));
        closest = mk(closest.x / cn, closest.y / cn, closest.z / cn);
    }
    t = INFINITY;
    for (vert = 0; vert < 4; vert++)
    {
        double sign = (vert == 0 || (vert == 1 && !is_vf)) ? -1.0 : 1.0;
        t = smin(t, sp_trajectory(H, verts[vert], mididx[vert], 0, planepos, planevel, scl(sign, closest), midpos[vert], midt, eta));
    }
    if (t < 1.0)
        if (sp_check(is_vf, H, verts, eta, mint, midt - smax(0.0, t - eps), eps))
            return 1;
    t = INFINITY;
    for (vert = 0; vert < 4; vert++)
    {
        double sign = (vert == 0 || (vert == 1 && !is_vf)) ? -1.0 : 1.0;
        t = smin(t, sp_trajectory(H, verts[vert], mididx[vert] + 1, 1, planepos, planevel, scl(sign, closest), midpos[vert], midt, eta));
    }
    if (t < 1.0)
        if (sp_check(is_vf, H, verts, eta, midt + smax(0.0, t - eps), maxt, eps))
            return 1;
    return 0;
}

/* SeparatingPlaneNarrowPhase::findCollisions, :11-25: eps = History::computeMinimumGap() / 4 (src/History.cpp:82-96) */
int orc_narrowphase_sepplane(int V, const long long *hoff, const double *htime, const double *hpos, long long nvf, const int *vf,
                             const double *vf_eta, long long nee, const int *ee, const double *ee_eta, unsigned char *vf_hit,
                             unsigned char *ee_hit)
{
    hist_t H;
    double gap = 1.0, eps;
    long long i, j;
    int v;
    H.hoff = hoff; H.htime = htime; H.hpos = hpos;
    for (v = 0; v < V; v++)
        for (j = hoff[v] + 1; j < hoff[v + 1]; j++)
            gap = smin(htime[j] - htime[j - 1], gap);
    eps = gap / 4.0;
    for (i = 0; i < nvf; i++)
        vf_hit[i] = (unsigned char)sp_check(1, &H, vf + 4 * i, vf_eta[i], 0.0, 1.0, eps);
    for (i = 0; i < nee; i++)
        ee_hit[i] = (unsigned char)sp_check(0, &H, ee + 4 * i, ee_eta[i], 0.0, 1.0, eps);
    return 0;
}
