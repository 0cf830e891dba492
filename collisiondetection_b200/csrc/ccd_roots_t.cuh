// Degree-specialised helpers of the root isolator (ccd_solve.cuh): every array has compile-time size D+1 and is only indexed
// by unrolled loop counters, so nothing lives in local memory.
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

} // namespace ccd
