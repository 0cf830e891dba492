// Device-side FP64 math for the CTCD primitives (sm_100a).
//
// Semantics follow evouga/collisiondetection src/CTCD.cpp (file:line cited per function); the
// polynomial coefficients are built with the reference's operation order and the translation
// unit is compiled with --fmad=false, so every product and sum is individually rounded exactly
// like the reference's CPU build.  Fused multiply-adds appear only where written as fma().
//
// Root finding: the reference's Jenkins-Traub search (src/rpoly.h) is replaced by a real-root
// isolator on [0,1] built from Bernstein sign-variation counts down the derivative chain plus
// bracketed Newton (roots01).
//
// Code shape: ONE copy of every routine (runtime degree, __noinline__, single call sites inside
// loops over the polynomials of a primitive).  The first version instantiated the isolator per
// degree and per call site; ncu showed the narrowphase stalled 58 % of the time on instruction
// fetch with 5.7 of 32 lanes active (profiles/).  Lanes of a warp now walk the same loop and meet
// again at the same call, and the whole kernel fits the instruction cache.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

// Every routine here is plain arithmetic and is also compiled for the host (tests/np_emul.cu runs the same code on the
// CPU against the checker); device-only intrinsics are wrapped.
#define CCD_HD __host__ __device__
#define CCD_FN CCD_HD __forceinline__

namespace ccd {

CCD_FN int ccd_ffs(unsigned x)
{
#ifdef __CUDA_ARCH__
    return __ffs(x);
#else
    return __builtin_ffs((int)x);
#endif
}
CCD_FN int ccd_popc(unsigned x)
{
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

struct V3 { double x, y, z; };

CCD_FN V3 mk(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
CCD_FN V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
CCD_FN V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
CCD_FN V3 operator*(double s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
// dot = (x0*y0 + x1*y1) + x2*y2 ; cross in the usual component order — the order the CPU checker uses
CCD_FN double dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
CCD_FN V3 cross(V3 a, V3 b)
{
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
CCD_FN V3 ldv(const double *p) { return mk(p[0], p[1], p[2]); }

// std::max / std::min argument-order semantics (matters only for NaN)
CCD_FN double smax(double a, double b) { return (a < b) ? b : a; }
CCD_FN double smin(double a, double b) { return (b < a) ? b : a; }

// result codes of the primitives / stencil tests
enum { R_MISS = 0, R_HIT = 1, R_DEFER = 2 };
// how a primitive treats polynomials that need the iterative root isolator:
//   FULL   isolate in place;   DEFER  export them (Pend) and return R_DEFER;
//   RESUME take their roots from task records computed by the root kernel (narrowphase.cu)
enum { MODE_FULL = 0, MODE_DEFER = 1, MODE_RESUME = 2 };

// the polynomials of one primitive that still need root isolation: normalised, reduced coefficients
struct Pend
{
    double ops[5][7];
    int rds[5];
    unsigned mask;
};

// ------------------------------------------------------------------------------------------
// closed time intervals (include/CTCD.h:7-26); at most 7 per polynomial (<= 6 breakpoints)
// ------------------------------------------------------------------------------------------
struct Ivals
{
    int n;
    double l[7], u[7];
};

// TimeInterval ctor, include/CTCD.h:9-14
CCD_FN void push_interval(Ivals &iv, double tl, double tu)
{
    double l = tl, u = tu;
    if (l > u) { double t = l; l = u; u = t; }
    l = smax(l, 0.0);
    u = smin(u, 1.0);
    iv.l[iv.n] = l;
    iv.u[iv.n] = u;
    iv.n++;
}

CCD_FN bool overlap2(double al, double au, double bl, double bu) { return !(al > bu || bl > au); }

// ------------------------------------------------------------------------------------------
// Real roots on [0,1] of c[0] t^d + ... + c[d], c[0] != 0, 3 <= d <= 6.
// Same steps, same fused operations as oracle/ccd_oracle.c: orc_roots01.
// ------------------------------------------------------------------------------------------
CCD_FN double horner_fma(const double *c, int m, double x)
{
    double f = c[0];
    for (int i = 1; i <= m; i++)
        f = fma(f, x, c[i]);
    return f;
}

// root of the degree-m polynomial p in (lo,hi); f(lo) has the sign of flo, f(hi) the opposite
static CCD_HD __noinline__ double solve_bracket(const double *p, int m, double lo, double hi, double flo)
{
    double c[7];
    for (int i = 0; i <= m; i++)
        c[i] = p[i];
    double x = 0.5 * (lo + hi);
    double dxold = hi - lo, dx = dxold;
    const bool lo_neg = flo < 0.0;
    for (int it = 0; it < 128; it++)
    {
        double f = c[0], df = 0.0;
        for (int i = 1; i <= m; i++)
        {
            df = fma(df, x, f);
            f = fma(f, x, c[i]);
        }
        if (f == 0.0)
            return x;
        if ((f < 0.0) == lo_neg)
            lo = x;
        else
            hi = x;
        double step = f / df;
        double xn = x - step;
        bool bisect = !(xn > lo && xn < hi);
        if (!bisect && fabs(2.0 * f) > fabs(dxold * df))
            bisect = true;
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

CCD_FN int sign_variations(const double *b, int count)
{
    int v = 0, last = 0;
    for (int i = 0; i < count; i++)
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

// reciprocal binomials 1/C(d,i), the same constants (and roundings) as the checker's table
CCD_FN double rbinom(int d, int i)
{
    const double R3[4] = {1.0, 1.0 / 3.0, 1.0 / 3.0, 1.0};
    const double R4[5] = {1.0, 1.0 / 4.0, 1.0 / 6.0, 1.0 / 4.0, 1.0};
    const double R5[6] = {1.0, 1.0 / 5.0, 1.0 / 10.0, 1.0 / 10.0, 1.0 / 5.0, 1.0};
    const double R6[7] = {1.0, 1.0 / 6.0, 1.0 / 15.0, 1.0 / 20.0, 1.0 / 15.0, 1.0 / 6.0, 1.0};
    return d == 3 ? R3[i] : d == 4 ? R4[i] : d == 5 ? R5[i] : R6[i];
}

// Bernstein coefficients on [0,1] of the degree-d polynomial c (descending): scaled power coefficients,
// then the binomial transform by repeated adjacent sums
CCD_FN void bernstein(const double *c, int d, double *b)
{
    for (int i = 0; i <= d; i++)
        b[i] = c[d - i] * rbinom(d, i);
    for (int k = 1; k <= d; k++)
        for (int i = d; i >= k; i--)
            b[i] = b[i] + b[i - 1];
}

// True when the top level already decides "no root in [0,1]": end coefficients non-zero, no sign variation.
CCD_FN bool no_root_at_top(const double *b, int d)
{
    return b[0] != 0.0 && b[d] != 0.0 && sign_variations(b, d + 1) == 0;
}

// power coefficients of derivative level m of the degree-d polynomial c (successive q' steps, rounded like the checker)
CCD_FN void deriv_level(const double *c, int d, int m, double *p)
{
    for (int i = 0; i <= d; i++)
        p[i] = c[i];
    for (int k = d; k > m; k--)
        for (int i = 0; i < k; i++)
            p[i] = p[i] * (double)(k - i);
}

// b holds the Bernstein coefficients of c (degree d) on entry and is used as scratch
static CCD_HD __noinline__ int roots01(const double *c, int d, double *b, double *roots)
{
    double p[7], cur[6];
    int ncur = 0, m0;
    bool one = false;      // level m0 has exactly one sign variation: one root, bracketed by [0,1]
    // descend to the first level that can be decided
    for (m0 = d; m0 >= 2; m0--)
    {
        if (m0 < d)
            for (int i = 0; i <= m0; i++)          // Bernstein coefficients of the derivative: forward differences
                b[i] = b[i + 1] - b[i];
        if (b[0] != 0.0 && b[m0] != 0.0)
        {
            int v = sign_variations(b, m0 + 1);
            if (v == 0)
                break;
            if (v == 1)
            {
                one = true;
                break;
            }
        }
        if (m0 == 2)
        {
            // both critical points may lie inside: closed form, cancellation-free
            deriv_level(c, d, 2, p);
            double a = p[0], bb = p[1], cc = p[2];
            double D = fma(bb, bb, -4.0 * a * cc);
            if (D >= 0.0)
            {
                double q = -0.5 * (bb + (bb < 0.0 ? -sqrt(D) : sqrt(D)));
                double r0 = q / a, r1 = (q != 0.0) ? cc / q : r0;
                if (r0 > r1) { double t = r0; r0 = r1; r1 = t; }
                if (r0 > 0.0 && r0 < 1.0) cur[ncur++] = r0;
                if (r1 > 0.0 && r1 < 1.0 && r1 != r0) cur[ncur++] = r1;
            }
            break;
        }
    }
    // climb: cur = roots of q_{m-1} strictly inside (0,1); pieces between them are monotone for q_m.
    // A level decided by "one sign variation" is handled by the same loop as a single piece [0,1] in which only a
    // strict sign change counts (zeros at the ends are not roots there), so every bracketed solve of every lane of
    // the warp is issued from the one call below.
    double brk[8], fv[8], out[7];
    for (int m = one ? m0 : m0 + 1; m <= d; m++)
    {
        const bool plain = one && m == m0;
        const bool last = (m == d);
        int nb = 0, nr = 0;
        deriv_level(c, d, m, p);
        brk[nb++] = 0.0;
        if (!plain)
            for (int i = 0; i < ncur; i++)
                brk[nb++] = cur[i];
        brk[nb++] = 1.0;
        for (int i = 0; i < nb; i++)
            fv[i] = horner_fma(p, m, brk[i]);
        for (int i = 0; i + 1 < nb; i++)
        {
            if (fv[i] == 0.0)
            {
                if (!plain && (i > 0 || last) && (nr == 0 || out[nr - 1] != brk[i]))
                    out[nr++] = brk[i];
            }
            else if ((fv[i] < 0.0 && fv[i + 1] > 0.0) || (fv[i] > 0.0 && fv[i + 1] < 0.0))
            {
                double r = solve_bracket(p, m, brk[i], brk[i + 1], fv[i]);
                if (nr == 0 || out[nr - 1] != r)
                    out[nr++] = r;
            }
        }
        if (!plain && last && fv[nb - 1] == 0.0 && (nr == 0 || out[nr - 1] != 1.0))
            out[nr++] = 1.0;
        ncur = 0;
        for (int i = 0; i < nr; i++)
            if ((last && !plain) || (out[i] > 0.0 && out[i] < 1.0))
                cur[ncur++] = out[i];
    }
    for (int i = 0; i < ncur; i++)
        roots[i] = cur[i];
    return ncur;
}

// ------------------------------------------------------------------------------------------
// CTCD::findIntervals (src/CTCD.cpp:98-177)
// ------------------------------------------------------------------------------------------
// CTCD::checkInterval, src/CTCD.cpp:59-79 — unfused Horner at the clamped midpoint
static CCD_HD __noinline__ void check_interval(double t1, double t2, const double *op, int degree, Ivals &iv, bool pos)
{
    t1 = smax(0.0, t1);
    t2 = smax(0.0, t2);
    t1 = smin(1.0, t1);
    t2 = smin(1.0, t2);
    double tmid = (t2 + t1) / 2;
    double f = op[0];
    for (int i = 1; i <= degree; i++)
    {
        f *= tmid;
        f += op[i];
    }
    if (pos ? (f >= 0) : (f <= 0))
        push_interval(iv, t1, t2);
}

// CTCD::getQuadRoots, src/CTCD.cpp:38-56
static CCD_HD __noinline__ int quad_roots(double a, double b, double c, double &t0, double &t1)
{
    int roots = 0;
    double sign = (b < 0) ? -1.0 : 1.0;
    double D = b * b - 4 * a * c;
    if (D >= 0)
    {
        roots = 2;
        double q = -0.5 * (b + sign * sqrt(D));
        t0 = q / a;
        t1 = c / q;
        if (t0 > t1) { double t = t0; t0 = t1; t1 = t; }
    }
    return roots;
}

// First half of findIntervals: normalise, drop exactly-zero leading coefficients, and settle everything that
// needs no iteration (quick reject, closed forms for degree <= 2, Bernstein "no root in [0,1]").
// op[0..n] is modified exactly like the reference does.  Returns 0 when iv is final, 1 when the reduced
// polynomial op[0..rd] (rd >= 3) still needs the root isolator (finish_poly).
static CCD_HD __noinline__ int prepare_poly(double *op, int n, Ivals &iv, bool pos, int &rd_out)
{
    // normalise, src/CTCD.cpp:113-119
    double maxval = 0;
    for (int i = 0; i <= n; i++)
        maxval = smax(maxval, fabs(op[i]));
    if (maxval != 0)
        for (int i = 0; i <= n; i++)
            op[i] = op[i] / maxval;
    // exactly-zero leading coefficients, src/CTCD.cpp:121-133
    int rd = n;
    for (int i = 0; i < n; i++)
    {
        if (op[i] == 0)
            rd--;
        else
            break;
    }
    if (rd < n)
        for (int i = 0; i <= rd; i++)
            op[i] = op[i + n - rd];
    rd_out = rd;

    double time[2];
    int roots = 0;
    if (rd > 2)
    {
        // CTCD::couldHaveRoots, src/CTCD.cpp:81-94
        double result = 0;
        for (int i = 0; i < rd; i++)
            if (pos ? (op[i] > 0) : (op[i] < 0))
                result += op[i];
        result += op[rd];
        if (pos ? (result < 0) : (result > 0))
            return 0;
        double b[7];
        bernstein(op, rd, b);
        if (!no_root_at_top(b, rd))
            return 1;
    }
    else if (rd == 2)
        roots = quad_roots(op[0], op[1], op[2], time[0], time[1]);
    else if (rd == 1)
    {
        time[0] = -op[1] / op[0];
        roots = 1;
    }
    else
    {
        if (pos ? (op[0] >= 0) : (op[0] <= 0))
            push_interval(iv, 0, 1.0);
        return 0;
    }
    // src/CTCD.cpp:161-176 (at most two breakpoints here)
    if (roots > 0)
    {
        if (time[0] >= 0)
            check_interval(0, time[0], op, rd, iv, pos);
        if (roots == 2 && !((time[0] < 0 && time[1] < 0) || (time[0] > 1.0 && time[1] > 1.0)))
            check_interval(time[0], time[1], op, rd, iv, pos);
        if (time[roots - 1] <= 1.0)
            check_interval(time[roots - 1], 1.0, op, rd, iv, pos);
    }
    else
        check_interval(0.0, 1.0, op, rd, iv, pos);
    return 0;
}

// the interval rules of src/CTCD.cpp:161-176 for the prepared polynomial and its real roots in [0,1]
CCD_FN void intervals_from_roots(const double *op, int rd, const double *time, int roots, Ivals &iv, bool pos)
{
    if (roots > 0)
    {
        if (time[0] >= 0)
            check_interval(0, time[0], op, rd, iv, pos);
        for (int i = 0; i < roots - 1; i++)
            if (!((time[i] < 0 && time[i + 1] < 0) || (time[i] > 1.0 && time[i + 1] > 1.0)))
                check_interval(time[i], time[i + 1], op, rd, iv, pos);
        if (time[roots - 1] <= 1.0)
            check_interval(time[roots - 1], 1.0, op, rd, iv, pos);
    }
    else
        check_interval(0.0, 1.0, op, rd, iv, pos);
}

// Second half of findIntervals: real roots of the prepared polynomial in [0,1], then the interval rules.
static CCD_HD __noinline__ void finish_poly(const double *op, int rd, Ivals &iv, bool pos)
{
    double time[6], b[7];
    bernstein(op, rd, b);
    const int roots = roots01(op, rd, b, time);
    intervals_from_roots(op, rd, time, roots, iv, pos);
}

// CTCD::findIntervals for a single polynomial (FULL mode).
CCD_FN void find_intervals(double *op, int n, Ivals &iv, bool pos)
{
    int rd;
    if (prepare_poly(op, n, iv, pos, rd))
        finish_poly(op, rd, iv, pos);
}

// Settle the pending polynomials of a primitive: lists[k] receives the intervals of polynomial k (bit k of P.mask),
// posmask bit k = the reference's `pos` flag.  RESUME reads 64-byte task records {roots[6], -, count} in bit order.
// Returns false as soon as a list comes out empty (the primitive misses), else true.
template <int MODE> static CCD_HD __noinline__ bool resolve_pending(Pend &P, Ivals *lists, unsigned posmask, const double *trec)
{
    int j = 0;
    while (P.mask)
    {
        const int k = ccd_ffs(P.mask) - 1;
        P.mask &= P.mask - 1;
        const bool pos = (posmask >> k) & 1u;
        if (MODE == MODE_RESUME)
        {
            double r[6];
            const double *rec = trec + 8 * j;
            j++;
            const int nr = (int)rec[7];
            for (int i = 0; i < nr; i++) r[i] = rec[i];
            intervals_from_roots(P.ops[k], P.rds[k], r, nr, lists[k], pos);
        }
        else
            finish_poly(P.ops[k], P.rds[k], lists[k], pos);
        if (lists[k].n == 0)
            return false;
    }
    return true;
}

// ------------------------------------------------------------------------------------------
// coefficient builders (src/CTCD.cpp:179-257)
// ------------------------------------------------------------------------------------------
// planePoly3D, src/CTCD.cpp:222-227
CCD_FN void plane_coeffs(V3 x10, V3 x20, V3 x30, V3 v10, V3 v20, V3 v30, double *op)
{
    op[0] = dot(v10, cross(v20, v30));
    op[1] = dot(x10, cross(v20, v30)) + dot(v10, cross(x20, v30)) + dot(v10, cross(v20, x30));
    op[2] = dot(x10, cross(x20, v30)) + dot(x10, cross(v20, x30)) + dot(v10, cross(x20, x30));
    op[3] = dot(x10, cross(x20, x30));
}

// distancePoly3D, src/CTCD.cpp:240-255
static CCD_HD __noinline__ void distance_coeffs(V3 x10, V3 x20, V3 x30, V3 v10, V3 v20, V3 v30, double m, double *op)
{
    double abcd[4];
    plane_coeffs(x10, x20, x30, v10, v20, v30, abcd);
    double A = abcd[0], B = abcd[1], C = abcd[2], D = abcd[3];
    V3 E = cross(x20, x30);
    V3 F = cross(x20, v30) + cross(v20, x30);
    V3 G = cross(v20, v30);
    op[0] = A * A;
    op[1] = 2 * A * B;
    op[2] = B * B + 2 * A * C - dot(G, G) * m;
    op[3] = 2 * A * D + 2 * B * C - 2 * dot(G, F) * m;
    op[4] = 2 * B * D + C * C - (2 * dot(G, E) + dot(F, F)) * m;
    op[5] = 2 * C * D - 2 * dot(F, E) * m;
    op[6] = D * D - dot(E, E) * m;
}

// barycentricPoly3D, src/CTCD.cpp:187-210
CCD_FN void barycentric_coeffs(V3 x10, V3 x20, V3 x30, V3 v10, V3 v20, V3 v30, double *op)
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

// polynomial k of the VF primitive: k = 0,1,2 the inside cubics e1,e2,e3 — one formula under a rotation of the face,
// base = 1+k, A = 1+(k+2)%3, B = 1+(k+1)%3: x10 = q0-base, x20 = (A-base) x (B-base), x30 = A-base (src/CTCD.cpp:433-464);
// k = 3 the coplanarity sextic (src/CTCD.cpp:466-472)
CCD_FN void build_vf_poly(int k, const V3 *s, const V3 *v, double eta, double *op)
{
    if (k < 3)
    {
        const int ib = 1 + k, ia = 1 + (k + 2) % 3, ic = 1 + (k + 1) % 3;
        const V3 xa = s[ia] - s[ib], xc = s[ic] - s[ib], va = v[ia] - v[ib], vc = v[ic] - v[ib];
        plane_coeffs(s[0] - s[ib], cross(xa, xc), xa, v[0] - v[ib], cross(va, vc), va, op);
    }
    else
        distance_coeffs(s[0] - s[1], s[2] - s[1], s[3] - s[1], v[0] - v[1], v[2] - v[1], v[3] - v[1], eta * eta, op);
}

// polynomial k of the EE primitive on points (q0,p0,q1,p1) = s[0..3]: k = 0..3 the barycentric quartics a0,a1,b0,b1
// (src/CTCD.cpp:313-348: x10 = P[a]-P[b], x20 = P[c]-P[d], x30 = P[e]-P[f]); k = 4 the line-distance sextic (:266-288)
CCD_FN void build_ee_poly(int k, const V3 *s, const V3 *v, double eta, double *op)
{
    if (k < 4)
    {
        // packed point indices, 2 bits each: a b c d e f
        const unsigned tab = (k == 0) ? 0x0E42u /*3,2 | 1,0 | 0,2*/ : (k == 1) ? 0x0E16u /*3,2 | 0,1 | 1,2*/
                           : (k == 2) ? 0x04E8u /*1,0 | 3,2 | 2,0*/ : 0x04BCu /*1,0 | 2,3 | 3,0*/;
        const int a = (tab >> 10) & 3, b = (tab >> 8) & 3, c = (tab >> 6) & 3, d = (tab >> 4) & 3, e = (tab >> 2) & 3, f = tab & 3;
        barycentric_coeffs(s[a] - s[b], s[c] - s[d], s[e] - s[f], v[a] - v[b], v[c] - v[d], v[e] - v[f], op);
    }
    else
        distance_coeffs(s[1] - s[3], s[1] - s[0], s[3] - s[2], v[1] - v[3], v[1] - v[0], v[3] - v[2], eta * eta, op);
}

// ------------------------------------------------------------------------------------------
// the four primitives (src/CTCD.cpp:259-692).  s[] = start positions, v[] = end - start.
// Return R_MISS / R_HIT (t written) / R_DEFER (only with defer=true: needs the iterative isolator).
// ------------------------------------------------------------------------------------------
// CTCD::vertexFaceCTCD, src/CTCD.cpp:413-508 — vertex s[0] against face (s[1],s[2],s[3])
template <int MODE> static CCD_HD __noinline__ int vertex_face(const V3 *s, const V3 *v, double eta, double &t, Pend &P, const double *trec)
{
    Ivals iv[4];       // e1, e2, e3, coplane
    P.mask = 0;
    // the three inside tests are one formula under a rotation of the face: base = 1+k, A = 1+(k+2)%3, B = 1+(k+1)%3
    //   x10 = q0-base, x20 = (A-base) x (B-base), x30 = A-base            (src/CTCD.cpp:433-464)
    // Every polynomial is built and settled as far as straight-line code goes before any root is isolated: an empty
    // list anywhere is a miss (src/CTCD.cpp:441,452,463,474), and the lanes of a warp then run the isolator together.
    for (int k = 0; k < 3; k++)
    {
        build_vf_poly(k, s, v, eta, P.ops[k]);
        iv[k].n = 0;
        if (prepare_poly(P.ops[k], 3, iv[k], true, P.rds[k]))
            P.mask |= 1u << k;
        else if (iv[k].n == 0)
            return R_MISS;
    }
    build_vf_poly(3, s, v, eta, P.ops[3]);
    iv[3].n = 0;
    if (prepare_poly(P.ops[3], 6, iv[3], false, P.rds[3]))
        P.mask |= 8u;
    else if (iv[3].n == 0)
        return R_MISS;
    if (P.mask)
    {
        if (MODE == MODE_DEFER)
            return R_DEFER;
        if (!resolve_pending<MODE>(P, iv, 0x7u, trec))
            return R_MISS;
    }
    const Ivals &cop = iv[3], &e1 = iv[0], &e2 = iv[1], &e3 = iv[2];
    bool col = false;
    double mint = 1.0;
    for (int i = 0; i < cop.n; i++)
        for (int j = 0; j < e1.n; j++)
        {
            if (!overlap2(cop.l[i], cop.u[i], e1.l[j], e1.u[j])) continue;
            for (int k = 0; k < e2.n; k++)
            {
                if (!overlap2(cop.l[i], cop.u[i], e2.l[k], e2.u[k]) || !overlap2(e1.l[j], e1.u[j], e2.l[k], e2.u[k])) continue;
                for (int l = 0; l < e3.n; l++)
                {
                    if (!overlap2(cop.l[i], cop.u[i], e3.l[l], e3.u[l]) || !overlap2(e1.l[j], e1.u[j], e3.l[l], e3.u[l]) ||
                        !overlap2(e2.l[k], e2.u[k], e3.l[l], e3.u[l]))
                        continue;
                    double il = smax(cop.l[i], 0.0);
                    il = smax(e1.l[j], il);
                    il = smax(e2.l[k], il);
                    il = smax(e3.l[l], il);
                    mint = smin(il, mint);
                    col = true;
                }
            }
        }
    if (col) t = mint;
    return col ? R_HIT : R_MISS;
}

// CTCD::edgeEdgeCTCD, src/CTCD.cpp:259-411 — points (q0,p0,q1,p1) = s[0..3]: edges (q0,p0) and (q1,p1)
template <int MODE> static CCD_HD __noinline__ int edge_edge(const V3 *s, const V3 *v, double eta, double &t, Pend &P, const double *trec)
{
    Ivals cop, par, q[5];      // q[0..3] = a0,a1,b0,b1 ; q[4] = raw coplanarity intervals
    P.mask = 0;
    cop.n = par.n = 0;
    build_ee_poly(4, s, v, eta, P.ops[4]);
    q[4].n = 0;
    if (prepare_poly(P.ops[4], 6, q[4], false, P.rds[4]))
        P.mask |= 16u;
    else if (q[4].n == 0)
        return R_MISS;
    // the four barycentric quartics a0,a1,b0,b1 (src/CTCD.cpp:313-348): x10 = P[a]-P[b], x20 = P[c]-P[d], x30 = P[e]-P[f]
    for (int k = 0; k < 4; k++)
    {
        build_ee_poly(k, s, v, eta, P.ops[k]);
        q[k].n = 0;
        if (prepare_poly(P.ops[k], 4, q[k], true, P.rds[k]))
            P.mask |= 1u << k;
        else if (q[k].n == 0)
            return R_MISS;
    }
    if (P.mask)
    {
        if (MODE == MODE_DEFER)
            return R_DEFER;
        if (!resolve_pending<MODE>(P, q, 0xFu, trec))
            return R_MISS;
    }
    {
        // parallel-edge classification at each interval midpoint, src/CTCD.cpp:290-308
        const Ivals &raw = q[4];
        for (int i = 0; i < raw.n; i++)
        {
            double midt = (raw.u[i] + raw.l[i]) / 2;
            V3 x10 = (s[0] - s[1]) + midt * (v[0] - v[1]);
            V3 x20 = (s[2] - s[3]) + midt * (v[2] - v[3]);
            V3 c = cross(x10, x20);
            if (sqrt(dot(c, c)) < 1e-8) { par.l[par.n] = raw.l[i]; par.u[par.n] = raw.u[i]; par.n++; }
            else { cop.l[cop.n] = raw.l[i]; cop.u[cop.n] = raw.u[i]; cop.n++; }
        }
        if (cop.n == 0) return R_MISS;
    }
    const Ivals &a0 = q[0], &a1 = q[1], &b0 = q[2], &b1 = q[3];
    bool col = false;
    double mint = 1.0;
    for (int i = 0; i < cop.n; i++)
        for (int j = 0; j < a0.n; j++)
        {
            if (!overlap2(cop.l[i], cop.u[i], a0.l[j], a0.u[j])) continue;
            for (int k = 0; k < a1.n; k++)
            {
                if (!overlap2(cop.l[i], cop.u[i], a1.l[k], a1.u[k]) || !overlap2(a0.l[j], a0.u[j], a1.l[k], a1.u[k])) continue;
                for (int l = 0; l < b0.n; l++)
                {
                    if (!overlap2(cop.l[i], cop.u[i], b0.l[l], b0.u[l]) || !overlap2(a0.l[j], a0.u[j], b0.l[l], b0.u[l]) ||
                        !overlap2(a1.l[k], a1.u[k], b0.l[l], b0.u[l]))
                        continue;
                    for (int m = 0; m < b1.n; m++)
                    {
                        if (!overlap2(cop.l[i], cop.u[i], b1.l[m], b1.u[m]) || !overlap2(a0.l[j], a0.u[j], b1.l[m], b1.u[m]) ||
                            !overlap2(a1.l[k], a1.u[k], b1.l[m], b1.u[m]) || !overlap2(b0.l[l], b0.u[l], b1.l[m], b1.u[m]))
                            continue;
                        double il = smax(cop.l[i], 0.0), iu = smin(cop.u[i], 1.0);
                        il = smax(a0.l[j], il); iu = smin(a0.u[j], iu);
                        il = smax(a1.l[k], il); iu = smin(a1.u[k], iu);
                        il = smax(b0.l[l], il); iu = smin(b0.u[l], iu);
                        il = smax(b1.l[m], il); iu = smin(b1.u[m], iu);
                        bool skip = false;
                        for (int p = 0; p < par.n; p++)
                            if (overlap2(il, iu, par.l[p], par.u[p])) { skip = true; break; }
                        if (!skip) { mint = smin(mint, il); col = true; }
                    }
                }
            }
        }
    if (col) t = mint;
    return col ? R_HIT : R_MISS;
}

// CTCD::vertexEdgeCTCD, src/CTCD.cpp:511-602 — vertex q0 against segment (q1,q2); v* = end - start
template <int MODE> static CCD_HD __noinline__ int vertex_edge(V3 q0s, V3 q1s, V3 q2s, V3 v0, V3 v1, V3 v2, double eta, double &t, Pend &P, const double *trec)
{
    const double minD = eta * eta;
    const V3 ab = q2s - q1s, ac = q0s - q1s, cb = q2s - q0s;
    const V3 vab = v2 - v1, vac = v0 - v1, vcb = v2 - v0;
    Ivals colin, e1, e2;
    double *op = P.ops[0];
    P.mask = 0;
    colin.n = e1.n = e2.n = 0;
    op[2] = dot(ab, ac);
    op[1] = dot(ac, vab) + dot(ab, vac);
    op[0] = dot(vab, vac);
    find_intervals(op, 2, e1, true);
    if (e1.n == 0) return R_MISS;
    op[2] = dot(ab, cb);
    op[1] = dot(cb, vab) + dot(ab, vcb);
    op[0] = dot(vab, vcb);
    find_intervals(op, 2, e2, true);
    if (e2.n == 0) return R_MISS;
    {
        double A = dot(ab, ab);
        double B = 2 * dot(ab, vab);
        double C = dot(vab, vab);
        double D = dot(ac, ac);
        double E = 2 * dot(ac, vac);
        double F = dot(vac, vac);
        double G = dot(ac, ab);
        double H = dot(vab, ac) + dot(vac, ab);
        double I = dot(vab, vac);
        op[4] = A * D - G * G - minD * A;
        op[3] = B * D + A * E - 2 * G * H - minD * B;
        op[2] = B * E + A * F + C * D - H * H - 2 * G * I - minD * C;
        op[1] = B * F + C * E - 2 * H * I;
        op[0] = C * F - I * I;
    }
    if (prepare_poly(op, 4, colin, false, P.rds[0]))
    {
        P.mask = 1u;
        if (MODE == MODE_DEFER)
            return R_DEFER;
        resolve_pending<MODE>(P, &colin, 0u, trec);
    }
    if (colin.n == 0) return R_MISS;
    bool col = false;
    double mint = 1.0;
    for (int i = 0; i < colin.n; i++)
        for (int j = 0; j < e1.n; j++)
            for (int k = 0; k < e2.n; k++)
                if (overlap2(colin.l[i], colin.u[i], e1.l[j], e1.u[j]) && overlap2(colin.l[i], colin.u[i], e2.l[k], e2.u[k]) &&
                    overlap2(e1.l[j], e1.u[j], e2.l[k], e2.u[k]))
                {
                    double il = smax(e2.l[k], smax(e1.l[j], smax(colin.l[i], 0.0)));
                    mint = smin(il, mint);
                    col = true;
                }
    if (col) t = mint;
    return col ? R_HIT : R_MISS;
}

// checkInterval with a throw-away list, as CTCD::vertexVertexCTCD uses it (src/CTCD.cpp:645-690)
CCD_FN bool check_once(double t1, double t2, const double *op)
{
    Ivals iv;
    iv.n = 0;
    check_interval(t1, t2, op, 2, iv, false);
    return iv.n != 0;
}

// CTCD::vertexVertexCTCD, src/CTCD.cpp:604-692 — closed form, never deferred
static CCD_HD __noinline__ int vertex_vertex(V3 q1s, V3 q2s, V3 v1, V3 v2, double eta, double &t)
{
    int roots = 0;
    const double min_d = eta * eta;
    double t1 = 0, t2 = 0;
    double a = dot(v1, v1) + dot(v2, v2) - 2 * dot(v1, v2);
    double b = 2 * (dot(v1, q1s) - dot(v2, q1s) - dot(v1, q2s) + dot(v2, q2s));
    double c = dot(q1s, q1s) + dot(q2s, q2s) - 2 * dot(q1s, q2s) - min_d;
    if (a != 0)
        roots = quad_roots(a, b, c, t1, t2);
    else if (b != 0)
    {
        t1 = -c / b;
        roots = 1;
    }
    else
    {
        if (c <= 0) { t = 0; return R_HIT; }
        return R_MISS;
    }
    double op[3] = {a, b, c};
    if (roots == 2)
    {
        if (check_once(0, t1, op)) { t = 0; return R_HIT; }
        if (check_once(t1, t2, op)) { t = t1; return R_HIT; }
        if (check_once(t2, 1.0, op)) { t = t2; return R_HIT; }
        return R_MISS;
    }
    else if (roots == 1)
    {
        if (check_once(0, t1, op)) { t = 0; return R_HIT; }
        if (check_once(t1, 1.0, op)) { t = t1; return R_HIT; }
        return R_MISS;
    }
    if (check_once(0, 1.0, op)) { t = 0; return R_HIT; }
    return R_MISS;
}

} // namespace ccd
