// Register-resident first half of findIntervals for the dense pass-1 kernels: classify one polynomial as
// EMPTY (no interval of [0,1] satisfies it), DECIDED (some interval does, known without iteration) or PENDING
// (needs the root isolator) — the same decisions, from the same roundings, as prepare_poly() in ccd_math.cuh, with
// every array of compile-time size and indexed only by unrolled counters (nothing in local memory).
#pragma once
#include "ccd_math.cuh"

namespace ccd {

enum { PC_EMPTY = 0, PC_DECIDED = 1, PC_PENDING = 2 };

// "no root in [0,1]" from the Bernstein coefficients of the reduced polynomial c[0..RD] (descending): end
// coefficients non-zero and no sign variation (no_root_at_top)
template <int RD> CCD_FN bool no_root_static(const double *c)
{
    double b[RD + 1];
#pragma unroll
    for (int i = 0; i <= RD; i++)
        b[i] = c[RD - i] * rbinom(RD, i);
#pragma unroll
    for (int k = 1; k <= RD; k++)
#pragma unroll
        for (int i = RD; i >= k; i--)
            b[i] = b[i] + b[i - 1];
    if (b[0] == 0.0 || b[RD] == 0.0)
        return false;
    int v = 0, last = 0;
#pragma unroll
    for (int i = 0; i <= RD; i++)
    {
        const int s = (b[i] > 0.0) - (b[i] < 0.0);
        if (s != 0)
        {
            if (last != 0 && s != last) v++;
            last = s;
        }
    }
    return v == 0;
}

// ---- dyadic occupancy masks ------------------------------------------------------------------------------------
// A PENDING polynomial would normally send the whole stencil to the root isolator.  Most such stencils still miss:
// their polynomials are satisfied on parts of [0,1] that do not meet.  Three levels of de Casteljau subdivision give,
// per polynomial, an 8-bit mask of the eighths of [0,1] on which it can possibly be satisfied (an eighth is ruled out
// only when ALL its Bernstein coefficients have the wrong sign by more than 1e-12 of the normalised scale — the curve
// lies in their convex hull, so the polynomial has that sign on the whole closed eighth).  If the masks of a
// primitive's polynomials have no common bit, no time satisfies all of them, whatever the exact roots are: the
// interval lists the reference would build cannot overlap, and the primitive misses without any root being isolated.
// want: the eighths (bits of this subtree, lowest first) the caller will look at — the others are reported as 0 without
// being examined (the masks are only ever AND-ed into the running occupancy, so a stencil whose earlier polynomials
// already ruled most of [0,1] out pays for the few eighths that are left)
template <int RD, int LEVEL> CCD_FN unsigned dyadic_sub(const double (&b)[RD + 1], bool pos, unsigned want = 0xffu)
{
    if (!(want & ((1u << (1 << LEVEL)) - 1u)))
        return 0u;
    bool possible = false, everywhere = true;
#pragma unroll
    for (int i = 0; i <= RD; i++)
    {
        const bool ok = pos ? (b[i] >= -1e-12) : (b[i] <= 1e-12);
        possible = possible || ok;
        everywhere = everywhere && ok;
    }
    if (!possible)
        return 0u;
    // every coefficient within the margin: so is every coefficient of every sub-interval (de Casteljau halving takes
    // 0.5 * (a + b) of neighbours, and rounding is monotone: a, b >= t gives fl(a + b) >= 2 t exactly), i.e. every eighth
    // below this node would report "possible" — no need to split any further
    if (LEVEL == 0 || everywhere)
        return (1u << (1 << LEVEL)) - 1u;
    if constexpr (LEVEL > 0)
    {
    double l[RD + 1], r[RD + 1], t[RD + 1];
#pragma unroll
    for (int i = 0; i <= RD; i++)
        t[i] = b[i];
#pragma unroll
    for (int k = 0; k <= RD; k++)
    {
        l[k] = t[0];
        r[RD - k] = t[RD - k];
#pragma unroll
        for (int i = 0; i < RD - k; i++)
            t[i] = 0.5 * (t[i] + t[i + 1]);
    }
    const unsigned ml = dyadic_sub<RD, (LEVEL > 0 ? LEVEL - 1 : 0)>(l, pos, want);
    const unsigned mr = dyadic_sub<RD, (LEVEL > 0 ? LEVEL - 1 : 0)>(r, pos, want >> (1 << (LEVEL > 0 ? LEVEL - 1 : 0)));
    return ml | (mr << (1 << (LEVEL > 0 ? LEVEL - 1 : 0)));
    }
    return 1u;      // (LEVEL == 0 returned above)
}

template <int RD> CCD_FN unsigned dyadic_mask(const double *c, bool pos, unsigned want = 0xffu)
{
    double b[RD + 1];
#pragma unroll
    for (int i = 0; i <= RD; i++)
        b[i] = c[RD - i] * rbinom(RD, i);
#pragma unroll
    for (int k = 1; k <= RD; k++)
#pragma unroll
        for (int i = RD; i >= k; i--)
            b[i] = b[i] + b[i - 1];
    return dyadic_sub<RD, 3>(b, pos, want);
}

// mask for the normalised polynomial op[0..N] of reduced degree rd >= 3
template <int N> CCD_FN unsigned dyadic_mask_reduced(const double (&op)[N + 1], int rd, bool pos, unsigned want = 0xffu)
{
    if (N >= 6 && rd == 6) return dyadic_mask<(N >= 6 ? 6 : 3)>(&op[N >= 6 ? N - 6 : 0], pos, want);
    if (N >= 5 && rd == 5) return dyadic_mask<(N >= 5 ? 5 : 3)>(&op[N >= 5 ? N - 5 : 0], pos, want);
    if (N >= 4 && rd == 4) return dyadic_mask<(N >= 4 ? 4 : 3)>(&op[N >= 4 ? N - 4 : 0], pos, want);
    return dyadic_mask<3>(&op[N - 3], pos, want);
}

// would CTCD::checkInterval(t1,t2) push an interval?  Unfused Horner over op[0..N] (exactly-zero leading coefficients
// change nothing: 0*t + x == x)
template <int N> CCD_FN bool interval_ok(double t1, double t2, const double (&op)[N + 1], bool pos)
{
    t1 = smax(0.0, t1);
    t2 = smax(0.0, t2);
    t1 = smin(1.0, t1);
    t2 = smin(1.0, t2);
    const double tmid = (t2 + t1) / 2;
    double f = op[0];
#pragma unroll
    for (int i = 1; i <= N; i++)
    {
        f *= tmid;
        f += op[i];
    }
    return pos ? (f >= 0) : (f <= 0);
}

// op[0..N] raw coefficients in; normalised in place (leading zeros stay where they are); rd = reduced degree out
template <int N> CCD_FN int classify_poly(double (&op)[N + 1], bool pos, int &rd_out)
{
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
    int rd = N;
    {
        bool lead = true;
#pragma unroll
        for (int i = 0; i < N; i++)
        {
            lead = lead && (op[i] == 0);
            if (lead) rd--;
        }
    }
    rd_out = rd;
    if constexpr (N >= 3)
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
            return PC_EMPTY;
        bool noroot;
        if (N >= 6 && rd == 6) noroot = no_root_static<(N >= 6 ? 6 : 3)>(&op[N >= 6 ? N - 6 : 0]);
        else if (N >= 5 && rd == 5) noroot = no_root_static<(N >= 5 ? 5 : 3)>(&op[N >= 5 ? N - 5 : 0]);
        else if (N >= 4 && rd == 4) noroot = no_root_static<(N >= 4 ? 4 : 3)>(&op[N >= 4 ? N - 4 : 0]);
        else noroot = no_root_static<3>(&op[N - 3]);
        if (!noroot)
            return PC_PENDING;
        return interval_ok<N>(0.0, 1.0, op, pos) ? PC_DECIDED : PC_EMPTY;
    }
    if (rd == 2)
    {
        // CTCD::getQuadRoots + the interval rules of src/CTCD.cpp:161-176
        const double a = op[N - 2], b = op[N - 1], c = op[N];
        const double sign = (b < 0) ? -1.0 : 1.0;
        const double D = b * b - 4 * a * c;
        if (D >= 0)
        {
            const double q = -0.5 * (b + sign * sqrt(D));
            double t0 = q / a, t1 = c / q;
            if (t0 > t1) { double t = t0; t0 = t1; t1 = t; }
            bool ok = false;
            if (t0 >= 0) ok = interval_ok<N>(0, t0, op, pos);
            if (!ok && !((t0 < 0 && t1 < 0) || (t0 > 1.0 && t1 > 1.0))) ok = interval_ok<N>(t0, t1, op, pos);
            if (!ok && t1 <= 1.0) ok = interval_ok<N>(t1, 1.0, op, pos);
            return ok ? PC_DECIDED : PC_EMPTY;
        }
        return interval_ok<N>(0.0, 1.0, op, pos) ? PC_DECIDED : PC_EMPTY;
    }
    if (rd == 1)
    {
        const double t0 = -op[N] / op[N - 1];
        bool ok = false;
        if (t0 >= 0) ok = interval_ok<N>(0, t0, op, pos);
        if (!ok && t0 <= 1.0) ok = interval_ok<N>(t0, 1.0, op, pos);
        return ok ? PC_DECIDED : PC_EMPTY;
    }
    return (pos ? (op[N] >= 0) : (op[N] <= 0)) ? PC_DECIDED : PC_EMPTY;
}

// eighths of [0,1] touched by the closed interval [l,u] (already clamped to [0,1])
CCD_FN unsigned interval_mask(double l, double u)
{
    int a = (int)(l * 8.0), b = (int)(u * 8.0);
    if (a > 0 && (double)a * 0.125 == l) a--;       // a boundary point belongs to both neighbours
    if (a > 7) a = 7;
    if (b > 7) b = 7;
    return (0xffu >> (7 - b)) & (0xffu << a);
}

} // namespace ccd
