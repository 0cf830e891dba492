// Degree-specialised, register-resident form of the root isolator (same steps, same roundings as roots01 in
// ccd_math.cuh and orc_roots01 in the CPU checker) for the dense root kernels: every array has compile-time size D+1
// and is only indexed by unrolled loop counters, so nothing lives in local memory; the loops over derivative levels
// and over monotone pieces stay rolled, so lanes at different levels of different polynomials still share instructions.
//
// Level m of the derivative chain is stored RIGHT-ALIGNED in D+1 slots (D-m leading zeros).  Horner over all D+1 slots
// gives bit-identical values (0*x+0 = 0 exactly until the first real coefficient enters), and one derivative step is
// new[j+1] = old[j] * (D-j) for every j — independent of the level.
#pragma once
#include "ccd_math.cuh"

namespace ccd {

template <int D> CCD_FN void horner2_padded(const double (&p)[D + 1], double x, double &f, double &df)
{
    f = p[0];
    df = 0.0;
#pragma unroll
    for (int i = 1; i <= D; i++)
    {
        df = fma(df, x, f);
        f = fma(f, x, p[i]);
    }
}

template <int D> CCD_FN double horner_padded(const double (&p)[D + 1], double x)
{
    double f = p[0];
#pragma unroll
    for (int i = 1; i <= D; i++)
        f = fma(f, x, p[i]);
    return f;
}

// same iteration as solve_bracket() on the padded polynomial
template <int D> CCD_FN double solve_bracket_t(const double (&p)[D + 1], double lo, double hi, double flo)
{
    double x = 0.5 * (lo + hi);
    double dxold = hi - lo, dx = dxold;
    const bool lo_neg = flo < 0.0;
    for (int it = 0; it < 128; it++)
    {
        double f, df;
        horner2_padded<D>(p, x, f, df);
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

// level m of c (degree D) right-aligned in p
template <int D> CCD_FN void deriv_level_padded(const double (&c)[D + 1], int m, double (&p)[D + 1])
{
#pragma unroll
    for (int i = 0; i <= D; i++)
        p[i] = c[i];
    for (int k = D; k > m; k--)
    {
#pragma unroll
        for (int j = D - 1; j >= 0; j--)
            p[j + 1] = p[j] * (double)(D - j);
        p[0] = 0.0;
    }
}

// real roots in [0,1] of c[0] t^D + ... + c[D], c[0] != 0; returns the count, roots ascending in r[]
template <int D> CCD_FN int roots01_t(const double (&c)[D + 1], double (&r)[6])
{
    static_assert(D >= 3 && D <= 6, "degree 3..6");
    double b[D + 1], p[D + 1];
    // Bernstein coefficients: scaled power coefficients, then the binomial transform
#pragma unroll
    for (int i = 0; i <= D; i++)
        b[i] = c[D - i] * rbinom(D, i);
#pragma unroll
    for (int k = 1; k <= D; k++)
#pragma unroll
        for (int i = D; i >= k; i--)
            b[i] = b[i] + b[i - 1];

    double cur[D];        // roots of the level below, ascending (at most m-1 <= D-1 used; slot D-1 spare)
    int ncur = 0, m0;
    bool one = false;
    for (m0 = D; m0 >= 2; m0--)
    {
        if (m0 < D)
        {
#pragma unroll
            for (int i = 0; i < D; i++)
                if (i <= m0) b[i] = b[i + 1] - b[i];
        }
        double bend = b[0];
#pragma unroll
        for (int i = 1; i <= D; i++)
            if (i == m0) bend = b[i];
        if (b[0] != 0.0 && bend != 0.0)
        {
            int v = 0, last = 0;
#pragma unroll
            for (int i = 0; i <= D; i++)
                if (i <= m0)
                {
                    const int s = (b[i] > 0.0) - (b[i] < 0.0);
                    if (s != 0)
                    {
                        if (last != 0 && s != last) v++;
                        last = s;
                    }
                }
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
            deriv_level_padded<D>(c, 2, p);
            const double a = p[D - 2], bb = p[D - 1], cc = p[D];
            const double Dq = fma(bb, bb, -4.0 * a * cc);
            if (Dq >= 0.0)
            {
                const double q = -0.5 * (bb + (bb < 0.0 ? -sqrt(Dq) : sqrt(Dq)));
                double r0 = q / a, r1 = (q != 0.0) ? cc / q : r0;
                if (r0 > r1) { double t = r0; r0 = r1; r1 = t; }
                if (r0 > 0.0 && r0 < 1.0) { cur[0] = r0; ncur = 1; }
                if (r1 > 0.0 && r1 < 1.0 && r1 != r0)
                {
                    if (ncur == 0) cur[0] = r1; else cur[1] = r1;
                    ncur++;
                }
            }
            break;
        }
    }
    for (int m = one ? m0 : m0 + 1; m <= D; m++)
    {
        const bool plain = one && m == m0;
        const bool last = (m == D);
        deriv_level_padded<D>(c, m, p);
        // breakpoints 0, cur[0..ncur), 1 and the values there
        const int nb = plain ? 2 : ncur + 2;
        double brk[D + 1], fv[D + 1], out[D];
#pragma unroll
        for (int j = 0; j <= D; j++)
        {
            double x = 1.0;
            if (j == 0) x = 0.0;
            else if (!plain && j <= ncur) x = cur[j - 1 < D ? j - 1 : D - 1];
            brk[j] = x;
            fv[j] = (j < nb) ? horner_padded<D>(p, x) : 0.0;
        }
        int nr = 0;
        double prev = 0.0;      // out[nr-1]
#pragma unroll
        for (int i = 0; i < D; i++)
        {
            if (i + 1 < nb)
            {
                double root = 0.0;
                bool push = false;
                if (fv[i] == 0.0)
                {
                    if (!plain && (i > 0 || last) && (nr == 0 || prev != brk[i])) { root = brk[i]; push = true; }
                }
                else if ((fv[i] < 0.0 && fv[i + 1] > 0.0) || (fv[i] > 0.0 && fv[i + 1] < 0.0))
                {
                    root = solve_bracket_t<D>(p, brk[i], brk[i + 1], fv[i]);
                    push = (nr == 0 || prev != root);
                }
                if (push)
                {
#pragma unroll
                    for (int k = 0; k < D; k++)
                        if (k == nr) out[k] = root;
                    prev = root;
                    nr++;
                }
            }
        }
        {
            double fend = fv[0];
#pragma unroll
            for (int j = 1; j <= D; j++)
                if (j == nb - 1) fend = fv[j];
            if (!plain && last && fend == 0.0 && (nr == 0 || prev != 1.0))
            {
#pragma unroll
                for (int k = 0; k < D; k++)
                    if (k == nr) out[k] = 1.0;
                nr++;
            }
        }
        ncur = 0;
#pragma unroll
        for (int i = 0; i < D; i++)
            if (i < nr && ((last && !plain) || (out[i] > 0.0 && out[i] < 1.0)))
            {
#pragma unroll
                for (int k = 0; k < D; k++)
                    if (k == ncur) cur[k] = out[i];
                ncur++;
            }
    }
#pragma unroll
    for (int i = 0; i < D && i < 6; i++)
        r[i] = cur[i];
    return ncur;
}

} // namespace ccd
