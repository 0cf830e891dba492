// Device-side FP64 math for the CTCD primitives (sm_100a).
//
// Semantics follow evouga/collisiondetection src/CTCD.cpp (file:line cited per function); the
// polynomial coefficients are built with the reference's operation order and the translation
// unit is compiled with --fmad=false, so every product and sum is individually rounded exactly
// like the reference's CPU build.  Fused multiply-adds appear only where written as fma().
//
// Root finding: the reference's Jenkins-Traub search (src/rpoly.h) is replaced by a real-root
// isolator on [0,1] built from Bernstein sign-variation counts down the derivative chain plus
// bracketed Newton (see roots01<D>).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace ccd {

struct V3 { double x, y, z; };

__device__ __forceinline__ V3 mk(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
// dot = (x0*y0 + x1*y1) + x2*y2 ; cross in the usual component order — the order the CPU checker uses
__device__ __forceinline__ double dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b)
{
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ V3 ldv(const double *p) { return mk(p[0], p[1], p[2]); }

// std::max / std::min argument-order semantics (matters only for NaN)
__device__ __forceinline__ double smax(double a, double b) { return (a < b) ? b : a; }
__device__ __forceinline__ double smin(double a, double b) { return (b < a) ? b : a; }

// ------------------------------------------------------------------------------------------
// closed time intervals (include/CTCD.h:7-26); at most 7 per polynomial (<= 6 breakpoints)
// ------------------------------------------------------------------------------------------
struct Ivals
{
    int n;
    double l[7], u[7];
};

// TimeInterval ctor, include/CTCD.h:9-14
__device__ __forceinline__ void push_interval(Ivals &iv, double tl, double tu)
{
    double l = tl, u = tu;
    if (l > u) { double t = l; l = u; u = t; }
    l = smax(l, 0.0);
    u = smin(u, 1.0);
    iv.l[iv.n] = l;
    iv.u[iv.n] = u;
    iv.n++;
}

__device__ __forceinline__ bool overlap2(double al, double au, double bl, double bu) { return !(al > bu || bl > au); }

// ------------------------------------------------------------------------------------------
// Real roots on [0,1] of a degree-D polynomial, c[0] != 0 (descending powers).
// Same steps, same fused operations as oracle/ccd_oracle.c: orc_roots01.
// ------------------------------------------------------------------------------------------
template <int M> __device__ __forceinline__ double horner_fma(const double (&c)[M + 1], double x)
{
    double f = c[0];
#pragma unroll
    for (int i = 1; i <= M; i++)
        f = fma(f, x, c[i]);
    return f;
}

// root of c in (lo,hi); f(lo) has the sign of flo, f(hi) the opposite
template <int M> __device__ __forceinline__ double solve_bracket(const double (&c)[M + 1], double lo, double hi, double flo)
{
    double x = 0.5 * (lo + hi);
    double dxold = hi - lo, dx = dxold;
    const bool lo_neg = flo < 0.0;
    for (int it = 0; it < 128; it++)
    {
        double f = c[0], df = 0.0;
#pragma unroll
        for (int i = 1; i <= M; i++)
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

__device__ __forceinline__ int sgn(double v) { return (v > 0.0) - (v < 0.0); }

// Derivative level M of the degree-D polynomial c: successive q' steps, each rounding like the checker
template <int D, int M> __device__ __forceinline__ void deriv_level(const double (&c)[D + 1], double (&p)[M + 1])
{
    double t[D + 1];
#pragma unroll
    for (int i = 0; i <= D; i++)
        t[i] = c[i];
#pragma unroll
    for (int m = D; m > M; m--)
    {
#pragma unroll
        for (int i = 0; i < m; i++)
            t[i] = t[i] * (double)(m - i);
    }
#pragma unroll
    for (int i = 0; i <= M; i++)
        p[i] = t[i];
}

template <int D> struct Climb
{
    // one climb step at level M (roots of q_{M-1} in cur -> roots of q_M), M in (m0, D]
    template <int M> static __device__ __forceinline__ void step(const double (&c)[D + 1], int m0, double *cur, int &ncur)
    {
        if (M <= m0)
            return;
        double p[M + 1];
        deriv_level<D, M>(c, p);
        const bool last = (M == D);
        double brk[8], fv[8], out[7];
        int nb = 0, nr = 0;
        brk[nb++] = 0.0;
        for (int i = 0; i < ncur; i++)
            brk[nb++] = cur[i];
        brk[nb++] = 1.0;
        for (int i = 0; i < nb; i++)
            fv[i] = horner_fma<M>(p, brk[i]);
        for (int i = 0; i + 1 < nb; i++)
        {
            if (fv[i] == 0.0)
            {
                if ((i > 0 || last) && (nr == 0 || out[nr - 1] != brk[i]))
                    out[nr++] = brk[i];
            }
            else if ((fv[i] < 0.0 && fv[i + 1] > 0.0) || (fv[i] > 0.0 && fv[i + 1] < 0.0))
            {
                double r = solve_bracket<M>(p, brk[i], brk[i + 1], fv[i]);
                if (nr == 0 || out[nr - 1] != r)
                    out[nr++] = r;
            }
        }
        if (last && fv[nb - 1] == 0.0 && (nr == 0 || out[nr - 1] != 1.0))
            out[nr++] = 1.0;
        ncur = 0;
        for (int i = 0; i < nr; i++)
            if (last || (out[i] > 0.0 && out[i] < 1.0))
                cur[ncur++] = out[i];
    }
};

// Decide level M from its Bernstein coefficients b[0..M]; returns true when the descent stops here.
template <int D, int M>
__device__ __forceinline__ bool decide_level(const double (&c)[D + 1], const double *b, double *cur, int &ncur)
{
    if (b[0] != 0.0 && b[M] != 0.0)
    {
        int v = 0, last = 0;
#pragma unroll
        for (int i = 0; i <= M; i++)
        {
            int s = sgn(b[i]);
            if (s != 0)
            {
                if (last != 0 && s != last)
                    v++;
                last = s;
            }
        }
        if (v == 0)
            return true;
        if (v == 1)
        {
            double p[M + 1];
            deriv_level<D, M>(c, p);
            double f0 = p[M], f1 = horner_fma<M>(p, 1.0);
            if ((f0 < 0.0 && f1 > 0.0) || (f0 > 0.0 && f1 < 0.0))
                cur[ncur++] = solve_bracket<M>(p, 0.0, 1.0, f0);
            return true;
        }
    }
    if (M == 2)
    {
        double p[M + 1];
        deriv_level<D, M>(c, p);
        double a = p[0], bb = p[1], cc = p[2];
        double Dd = fma(bb, bb, -4.0 * a * cc);
        if (Dd >= 0.0)
        {
            double q = -0.5 * (bb + (bb < 0.0 ? -sqrt(Dd) : sqrt(Dd)));
            double r0 = q / a, r1 = (q != 0.0) ? cc / q : r0;
            if (r0 > r1) { double t = r0; r0 = r1; r1 = t; }
            if (r0 > 0.0 && r0 < 1.0) cur[ncur++] = r0;
            if (r1 > 0.0 && r1 < 1.0 && r1 != r0) cur[ncur++] = r1;
        }
        return true;
    }
    return false;
}

template <int D> __device__ __noinline__ int roots01(const double *cin, double *roots)
{
    static_assert(D >= 3 && D <= 6, "degree 3..6");
    double c[D + 1];
#pragma unroll
    for (int i = 0; i <= D; i++)
        c[i] = cin[i];

    // Bernstein coefficients of q_D on [0,1]: scaled power coefficients, then the binomial transform
    double b[D + 1];
    {
        constexpr double RB[7][7] = {
            {1.0, 0, 0, 0, 0, 0, 0},
            {1.0, 1.0, 0, 0, 0, 0, 0},
            {1.0, 1.0 / 2.0, 1.0, 0, 0, 0, 0},
            {1.0, 1.0 / 3.0, 1.0 / 3.0, 1.0, 0, 0, 0},
            {1.0, 1.0 / 4.0, 1.0 / 6.0, 1.0 / 4.0, 1.0, 0, 0},
            {1.0, 1.0 / 5.0, 1.0 / 10.0, 1.0 / 10.0, 1.0 / 5.0, 1.0, 0},
            {1.0, 1.0 / 6.0, 1.0 / 15.0, 1.0 / 20.0, 1.0 / 15.0, 1.0 / 6.0, 1.0}};
#pragma unroll
        for (int i = 0; i <= D; i++)
            b[i] = c[D - i] * RB[D][i];
#pragma unroll
        for (int k = 1; k <= D; k++)
#pragma unroll
            for (int i = D; i >= k; i--)
                b[i] = b[i] + b[i - 1];
    }

    double cur[6];
    int ncur = 0, m0 = D;
    bool done = decide_level<D, D>(c, b, cur, ncur);
    if (done)
    {
        for (int i = 0; i < ncur; i++)
            roots[i] = cur[i];
        return ncur;
    }
    // descend: Bernstein coefficients of the derivative are forward differences
#define CCD_DESCEND(M)                                              \
    if (!done && D > M)                                             \
    {                                                               \
        _Pragma("unroll") for (int i = 0; i <= M; i++) b[i] = b[i + 1] - b[i]; \
        m0 = M;                                                     \
        done = decide_level<D, (M < D ? M : D)>(c, b, cur, ncur);   \
    }
    CCD_DESCEND(5)
    CCD_DESCEND(4)
    CCD_DESCEND(3)
    CCD_DESCEND(2)
#undef CCD_DESCEND

    if (D >= 3) Climb<D>::template step<3>(c, m0, cur, ncur);
    if (D >= 4) Climb<D>::template step<(D >= 4 ? 4 : D)>(c, m0, cur, ncur);
    if (D >= 5) Climb<D>::template step<(D >= 5 ? 5 : D)>(c, m0, cur, ncur);
    if (D >= 6) Climb<D>::template step<(D >= 6 ? 6 : D)>(c, m0, cur, ncur);
    for (int i = 0; i < ncur; i++)
        roots[i] = cur[i];
    return ncur;
}

// ------------------------------------------------------------------------------------------
// CTCD::findIntervals (src/CTCD.cpp:98-177) on a fixed-size coefficient array op[0..N].
// Leading zeros are kept in place: Horner and the couldHaveRoots sum give bit-identical values
// with or without them (0*t+x == x), so only the root finder sees the reduced polynomial.
// ------------------------------------------------------------------------------------------
// CTCD::checkInterval, src/CTCD.cpp:59-79 — unfused Horner at the clamped midpoint
template <int N> __device__ __forceinline__ void check_interval(double t1, double t2, const double (&op)[N + 1], Ivals &iv, bool pos)
{
    t1 = smax(0.0, t1);
    t2 = smax(0.0, t2);
    t1 = smin(1.0, t1);
    t2 = smin(1.0, t2);
    double tmid = (t2 + t1) / 2;
    double f = op[0];
#pragma unroll
    for (int i = 1; i <= N; i++)
    {
        f *= tmid;
        f += op[i];
    }
    if (pos ? (f >= 0) : (f <= 0))
        push_interval(iv, t1, t2);
}

// CTCD::getQuadRoots, src/CTCD.cpp:38-56
__device__ __forceinline__ int quad_roots(double a, double b, double c, double &t0, double &t1)
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

template <int N> __device__ __forceinline__ void find_intervals(double (&op)[N + 1], Ivals &iv, bool pos)
{
    // normalise, src/CTCD.cpp:113-119
    double maxval = 0;
#pragma unroll
    for (int i = 0; i <= N; i++)
        maxval = smax(maxval, fabs(op[i]));
    if (maxval != 0)
    {
#pragma unroll
        for (int i = 0; i <= N; i++)
            op[i] = op[i] / maxval;
    }
    // exactly-zero leading coefficients, src/CTCD.cpp:121-127
    int rd = N;
    {
        bool lead = true;
#pragma unroll
        for (int i = 0; i < N; i++)
        {
            lead = lead && (op[i] == 0);
            if (lead)
                rd--;
        }
    }
    double time[6];
    int roots = 0;
    if (rd > 2)
    {
        // CTCD::couldHaveRoots, src/CTCD.cpp:81-94 (zeros contribute nothing)
        double result = 0;
#pragma unroll
        for (int i = 0; i < N; i++)
            if (pos ? (op[i] > 0) : (op[i] < 0))
                result += op[i];
        result += op[N];
        if (pos ? (result < 0) : (result > 0))
            return;
        if constexpr (N >= 3)
        {
            if (N >= 6 && rd == 6) roots = roots01<(N >= 6 ? 6 : 3)>(&op[N >= 6 ? N - 6 : 0], time);
            else if (N >= 5 && rd == 5) roots = roots01<(N >= 5 ? 5 : 3)>(&op[N >= 5 ? N - 5 : 0], time);
            else if (N >= 4 && rd == 4) roots = roots01<(N >= 4 ? 4 : 3)>(&op[N >= 4 ? N - 4 : 0], time);
            else roots = roots01<3>(&op[N - 3], time);
        }
    }
    else if (rd == 2)
        roots = quad_roots(op[N - 2], op[N - 1], op[N], time[0], time[1]);
    else if (rd == 1)
    {
        time[0] = -op[N] / op[N - 1];
        roots = 1;
    }
    else
    {
        if (pos ? (op[N] >= 0) : (op[N] <= 0))
            push_interval(iv, 0, 1.0);
        return;
    }
    // src/CTCD.cpp:161-176
    if (roots > 0)
    {
        if (time[0] >= 0)
            check_interval<N>(0, time[0], op, iv, pos);
        for (int i = 0; i < roots - 1; i++)
            if (!((time[i] < 0 && time[i + 1] < 0) || (time[i] > 1.0 && time[i + 1] > 1.0)))
                check_interval<N>(time[i], time[i + 1], op, iv, pos);
        if (time[roots - 1] <= 1.0)
            check_interval<N>(time[roots - 1], 1.0, op, iv, pos);
    }
    else
        check_interval<N>(0.0, 1.0, op, iv, pos);
}

// ------------------------------------------------------------------------------------------
// coefficient builders (src/CTCD.cpp:179-257)
// ------------------------------------------------------------------------------------------
// planePoly3D, src/CTCD.cpp:222-227
__device__ __forceinline__ void plane_coeffs(V3 x10, V3 x20, V3 x30, V3 v10, V3 v20, V3 v30, double (&op)[4])
{
    op[0] = dot(v10, cross(v20, v30));
    op[1] = dot(x10, cross(v20, v30)) + dot(v10, cross(x20, v30)) + dot(v10, cross(v20, x30));
    op[2] = dot(x10, cross(x20, v30)) + dot(x10, cross(v20, x30)) + dot(v10, cross(x20, x30));
    op[3] = dot(x10, cross(x20, x30));
}

// distancePoly3D, src/CTCD.cpp:240-255
__device__ __forceinline__ void distance_coeffs(V3 x10, V3 x20, V3 x30, V3 v10, V3 v20, V3 v30, double m, double (&op)[7])
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
__device__ __forceinline__ void barycentric_coeffs(V3 x10, V3 x20, V3 x30, V3 v10, V3 v20, V3 v30, double (&op)[5])
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

// ------------------------------------------------------------------------------------------
// the four primitives (src/CTCD.cpp:259-692); positions: *s = start, *e = end of the linear move
// ------------------------------------------------------------------------------------------
// CTCD::vertexFaceCTCD, src/CTCD.cpp:413-508
static __device__ __noinline__ bool vertex_face(V3 q0s, V3 q1s, V3 q2s, V3 q3s, V3 q0e, V3 q1e, V3 q2e, V3 q3e, double eta, double &t)
{
    const double minD = eta * eta;
    const V3 v0 = q0e - q0s, v1 = q1e - q1s, v2 = q2e - q2s, v3 = q3e - q3s;
    Ivals cop, e1, e2, e3;
    cop.n = e1.n = e2.n = e3.n = 0;
    {
        double op[4];
        plane_coeffs(q0s - q1s, cross(q3s - q1s, q2s - q1s), q3s - q1s, v0 - v1, cross(v3 - v1, v2 - v1), v3 - v1, op);
        find_intervals<3>(op, e1, true);
        if (e1.n == 0) return false;
        plane_coeffs(q0s - q2s, cross(q1s - q2s, q3s - q2s), q1s - q2s, v0 - v2, cross(v1 - v2, v3 - v2), v1 - v2, op);
        find_intervals<3>(op, e2, true);
        if (e2.n == 0) return false;
        plane_coeffs(q0s - q3s, cross(q2s - q3s, q1s - q3s), q2s - q3s, v0 - v3, cross(v2 - v3, v1 - v3), v2 - v3, op);
        find_intervals<3>(op, e3, true);
        if (e3.n == 0) return false;
    }
    {
        double op[7];
        distance_coeffs(q0s - q1s, q2s - q1s, q3s - q1s, v0 - v1, v2 - v1, v3 - v1, minD, op);
        find_intervals<6>(op, cop, false);
        if (cop.n == 0) return false;
    }
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
    return col;
}

// CTCD::edgeEdgeCTCD, src/CTCD.cpp:259-411 — edges (q0,p0) and (q1,p1)
static __device__ __noinline__ bool edge_edge(V3 q0s, V3 p0s, V3 q1s, V3 p1s, V3 q0e, V3 p0e, V3 q1e, V3 p1e, double eta, double &t)
{
    const double minD = eta * eta;
    const V3 vq0 = q0e - q0s, vp0 = p0e - p0s, vq1 = q1e - q1s, vp1 = p1e - p1s;
    Ivals cop, par, a0, a1, b0, b1;
    cop.n = par.n = a0.n = a1.n = b0.n = b1.n = 0;
    {
        Ivals raw;
        raw.n = 0;
        double op[7];
        distance_coeffs(p0s - p1s, p0s - q0s, p1s - q1s, vp0 - vp1, vp0 - vq0, vp1 - vq1, minD, op);
        find_intervals<6>(op, raw, false);
        // parallel-edge classification at each interval midpoint, src/CTCD.cpp:290-308
        for (int i = 0; i < raw.n; i++)
        {
            double midt = (raw.u[i] + raw.l[i]) / 2;
            V3 x10 = (q0s - p0s) + midt * (vq0 - vp0);
            V3 x20 = (q1s - p1s) + midt * (vq1 - vp1);
            V3 c = cross(x10, x20);
            if (sqrt(dot(c, c)) < 1e-8) { par.l[par.n] = raw.l[i]; par.u[par.n] = raw.u[i]; par.n++; }
            else { cop.l[cop.n] = raw.l[i]; cop.u[cop.n] = raw.u[i]; cop.n++; }
        }
        if (cop.n == 0) return false;
    }
    {
        double op[5];
        barycentric_coeffs(p1s - q1s, p0s - q0s, q0s - q1s, vp1 - vq1, vp0 - vq0, vq0 - vq1, op);
        find_intervals<4>(op, a0, true);
        if (a0.n == 0) return false;
        barycentric_coeffs(p1s - q1s, q0s - p0s, p0s - q1s, vp1 - vq1, vq0 - vp0, vp0 - vq1, op);
        find_intervals<4>(op, a1, true);
        if (a1.n == 0) return false;
        barycentric_coeffs(p0s - q0s, p1s - q1s, q1s - q0s, vp0 - vq0, vp1 - vq1, vq1 - vq0, op);
        find_intervals<4>(op, b0, true);
        if (b0.n == 0) return false;
        barycentric_coeffs(p0s - q0s, q1s - p1s, p1s - q0s, vp0 - vq0, vq1 - vp1, vp1 - vq0, op);
        find_intervals<4>(op, b1, true);
        if (b1.n == 0) return false;
    }
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
                        for (int q = 0; q < par.n; q++)
                            if (overlap2(il, iu, par.l[q], par.u[q])) { skip = true; break; }
                        if (!skip) { mint = smin(mint, il); col = true; }
                    }
                }
            }
        }
    if (col) t = mint;
    return col;
}

// CTCD::vertexEdgeCTCD, src/CTCD.cpp:511-602 — vertex q0 against segment (q1,q2)
static __device__ __noinline__ bool vertex_edge(V3 q0s, V3 q1s, V3 q2s, V3 q0e, V3 q1e, V3 q2e, double eta, double &t)
{
    const double minD = eta * eta;
    const V3 v0 = q0e - q0s, v1 = q1e - q1s, v2 = q2e - q2s;
    const V3 ab = q2s - q1s, ac = q0s - q1s, cb = q2s - q0s;
    const V3 vab = v2 - v1, vac = v0 - v1, vcb = v2 - v0;
    Ivals colin, e1, e2;
    colin.n = e1.n = e2.n = 0;
    {
        double op[3];
        op[2] = dot(ab, ac);
        op[1] = dot(ac, vab) + dot(ab, vac);
        op[0] = dot(vab, vac);
        find_intervals<2>(op, e1, true);
        if (e1.n == 0) return false;
        op[2] = dot(ab, cb);
        op[1] = dot(cb, vab) + dot(ab, vcb);
        op[0] = dot(vab, vcb);
        find_intervals<2>(op, e2, true);
        if (e2.n == 0) return false;
    }
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
        double op[5];
        op[4] = A * D - G * G - minD * A;
        op[3] = B * D + A * E - 2 * G * H - minD * B;
        op[2] = B * E + A * F + C * D - H * H - 2 * G * I - minD * C;
        op[1] = B * F + C * E - 2 * H * I;
        op[0] = C * F - I * I;
        find_intervals<4>(op, colin, false);
        if (colin.n == 0) return false;
    }
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
    return col;
}

// checkInterval with a throw-away list, as CTCD::vertexVertexCTCD uses it (src/CTCD.cpp:645-690)
__device__ __forceinline__ bool check_once(double t1, double t2, const double (&op)[3])
{
    Ivals iv;
    iv.n = 0;
    check_interval<2>(t1, t2, op, iv, false);
    return iv.n != 0;
}

// CTCD::vertexVertexCTCD, src/CTCD.cpp:604-692
static __device__ __noinline__ bool vertex_vertex(V3 q1s, V3 q2s, V3 q1e, V3 q2e, double eta, double &t)
{
    int roots = 0;
    const double min_d = eta * eta;
    double t1 = 0, t2 = 0;
    const V3 v1 = q1e - q1s, v2 = q2e - q2s;
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
        if (c <= 0) { t = 0; return true; }
        return false;
    }
    double op[3] = {a, b, c};
    if (roots == 2)
    {
        if (check_once(0, t1, op)) { t = 0; return true; }
        if (check_once(t1, t2, op)) { t = t1; return true; }
        if (check_once(t2, 1.0, op)) { t = t2; return true; }
        return false;
    }
    else if (roots == 1)
    {
        if (check_once(0, t1, op)) { t = 0; return true; }
        if (check_once(t1, 1.0, op)) { t = t1; return true; }
        return false;
    }
    if (check_once(0, 1.0, op)) { t = 0; return true; }
    return false;
}

} // namespace ccd
