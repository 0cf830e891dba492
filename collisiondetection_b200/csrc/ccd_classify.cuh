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
template <int RD> __device__ __forceinline__ bool no_root_static(const double *c)
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

// would CTCD::checkInterval(t1,t2) push an interval?  Unfused Horner over op[0..N] (exactly-zero leading coefficients
// change nothing: 0*t + x == x)
template <int N> __device__ __forceinline__ bool interval_ok(double t1, double t2, const double (&op)[N + 1], bool pos)
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
template <int N> __device__ __forceinline__ int classify_poly(double (&op)[N + 1], bool pos, int &rd_out)
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

// Straight-line classification of a whole VF / EE primitive: false when some polynomial is EMPTY (the primitive
// misses), else true with the mask of pending polynomials (bit k as in build_vf_poly / build_ee_poly).  pendmask == 0 means every list is known
// and non-empty: the caller runs the full primitive for the interval combination (rare).
template <bool IS_VF> __device__ __forceinline__ bool classify_primitive(const V3 *s, const V3 *v, double eta, unsigned &pendmask)
{
    pendmask = 0;
    int rd;
    if (IS_VF)
    {
        for (int k = 0; k < 3; k++)
        {
            double op[4];
            build_vf_poly(k, s, v, eta, op);
            const int r = classify_poly<3>(op, true, rd);
            if (r == PC_EMPTY) return false;
            if (r == PC_PENDING) pendmask |= 1u << k;
        }
        double op[7];
        build_vf_poly(3, s, v, eta, op);
        const int r = classify_poly<6>(op, false, rd);
        if (r == PC_EMPTY) return false;
        if (r == PC_PENDING) pendmask |= 8u;
    }
    else
    {
        {
            double op[7];
            build_ee_poly(4, s, v, eta, op);
            const int r = classify_poly<6>(op, false, rd);
            if (r == PC_EMPTY) return false;
            if (r == PC_PENDING) pendmask |= 16u;
        }
        for (int k = 0; k < 4; k++)
        {
            double op[5];
            build_ee_poly(k, s, v, eta, op);
            const int r = classify_poly<4>(op, true, rd);
            if (r == PC_EMPTY) return false;
            if (r == PC_PENDING) pendmask |= 1u << k;
        }
    }
    return true;
}

// Rebuild pending polynomial k of the primitive and write its task record: normalised coefficients of the reduced
// polynomial + reduced degree — the same values prepare_poly() leaves in Pend.
template <bool IS_VF> __device__ __forceinline__ void export_poly(int k, const V3 *s, const V3 *v, double eta, double *rec)
{
    double op[7];
    int n;
    if (IS_VF) { build_vf_poly(k, s, v, eta, op); n = (k < 3) ? 3 : 6; }
    else { build_ee_poly(k, s, v, eta, op); n = (k < 4) ? 4 : 6; }
    double maxval = 0;
    for (int i = 0; i <= n; i++)
        maxval = smax(maxval, fabs(op[i]));
    if (maxval != 0)
        for (int i = 0; i <= n; i++)
            op[i] = op[i] / maxval;
    int rd = n;
    for (int i = 0; i < n; i++)
    {
        if (op[i] == 0) rd--;
        else break;
    }
    for (int i = 0; i <= rd; i++)
        rec[i] = op[i + n - rd];
    rec[7] = (double)rd;
}

} // namespace ccd
