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
template <int RD, int LEVEL> CCD_FN unsigned dyadic_sub(const double (&b)[RD + 1], bool pos)
{
    bool possible = false;
#pragma unroll
    for (int i = 0; i <= RD; i++)
        possible = possible || (pos ? (b[i] >= -1e-12) : (b[i] <= 1e-12));
    if (!possible)
        return 0u;
    if (LEVEL == 0)
        return 1u;
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
    const unsigned ml = dyadic_sub<RD, (LEVEL > 0 ? LEVEL - 1 : 0)>(l, pos);
    const unsigned mr = dyadic_sub<RD, (LEVEL > 0 ? LEVEL - 1 : 0)>(r, pos);
    return ml | (mr << (1 << (LEVEL > 0 ? LEVEL - 1 : 0)));
}

template <int RD> CCD_FN unsigned dyadic_mask(const double *c, bool pos)
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
    return dyadic_sub<RD, 3>(b, pos);
}

// mask for the normalised polynomial op[0..N] of reduced degree rd >= 3
template <int N> CCD_FN unsigned dyadic_mask_reduced(const double (&op)[N + 1], int rd, bool pos)
{
    if (N >= 6 && rd == 6) return dyadic_mask<(N >= 6 ? 6 : 3)>(&op[N >= 6 ? N - 6 : 0], pos);
    if (N >= 5 && rd == 5) return dyadic_mask<(N >= 5 ? 5 : 3)>(&op[N >= 5 ? N - 5 : 0], pos);
    if (N >= 4 && rd == 4) return dyadic_mask<(N >= 4 ? 4 : 3)>(&op[N >= 4 ? N - 4 : 0], pos);
    return dyadic_mask<3>(&op[N - 3], pos);
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

// Straight-line classification of a whole VF / EE primitive: false when some polynomial is EMPTY (the primitive
// misses), else true with the mask of pending polynomials (bit k as in build_vf_poly / build_ee_poly).  pendmask == 0 means every list is known
// and non-empty: the caller runs the full primitive for the interval combination (rare).
template <bool IS_VF> CCD_FN bool classify_primitive(const V3 *s, const V3 *v, double eta, unsigned &pendmask)
{
    pendmask = 0;
    unsigned occupancy = 0xffu;      // eighths of [0,1] on which every pending polynomial seen so far can be satisfied
    int rd;
    if (IS_VF)
    {
        for (int k = 0; k < 3; k++)
        {
            double op[4];
            build_vf_poly(k, s, v, eta, op);
            const int r = classify_poly<3>(op, true, rd);
            if (r == PC_EMPTY) return false;
            if (r == PC_PENDING) { pendmask |= 1u << k; occupancy &= dyadic_mask_reduced<3>(op, rd, true); }
        }
        double op[7];
        build_vf_poly(3, s, v, eta, op);
        const int r = classify_poly<6>(op, false, rd);
        if (r == PC_EMPTY) return false;
        if (r == PC_PENDING) { pendmask |= 8u; occupancy &= dyadic_mask_reduced<6>(op, rd, false); }
    }
    else
    {
        {
            double op[7];
            build_ee_poly(4, s, v, eta, op);
            const int r = classify_poly<6>(op, false, rd);
            if (r == PC_EMPTY) return false;
            if (r == PC_PENDING) { pendmask |= 16u; occupancy &= dyadic_mask_reduced<6>(op, rd, false); }
        }
        for (int k = 0; k < 4; k++)
        {
            double op[5];
            build_ee_poly(k, s, v, eta, op);
            const int r = classify_poly<4>(op, true, rd);
            if (r == PC_EMPTY) return false;
            if (r == PC_PENDING) { pendmask |= 1u << k; occupancy &= dyadic_mask_reduced<4>(op, rd, true); }
            if (occupancy == 0) return false;
        }
    }
    return occupancy != 0;
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

// occupancy mask of a polynomial of degree <= 2 (op[0..2], not yet normalised; normalised in place): the intervals the
// reference's rules give (src/CTCD.cpp:145-176), each widened to the eighths it touches.  0 = EMPTY.
CCD_FN unsigned lowdeg_mask(double (&op)[3], bool pos)
{
    double maxval = smax(smax(smax(0.0, fabs(op[0])), fabs(op[1])), fabs(op[2]));
    if (maxval != 0) { op[0] = op[0] / maxval; op[1] = op[1] / maxval; op[2] = op[2] / maxval; }
    unsigned m = 0;
    auto add = [&](double t1, double t2) {
        t1 = smin(1.0, smax(0.0, t1));
        t2 = smin(1.0, smax(0.0, t2));
        if (interval_ok<2>(t1, t2, op, pos)) m |= interval_mask(fmin(t1, t2), fmax(t1, t2));
    };
    if (op[0] != 0)
    {
        const double a = op[0], b = op[1], c = op[2];
        const double sign = (b < 0) ? -1.0 : 1.0;
        const double D = b * b - 4 * a * c;
        if (D >= 0)
        {
            const double q = -0.5 * (b + sign * sqrt(D));
            double t0 = q / a, t1 = c / q;
            if (t0 > t1) { double t = t0; t0 = t1; t1 = t; }
            if (!(t0 == t0) || !(t1 == t1)) return 0xffu;      // NaN roots (0/0): stay conservative, the general routine decides
            if (t0 >= 0) add(0, t0);
            if (!((t0 < 0 && t1 < 0) || (t0 > 1.0 && t1 > 1.0))) add(t0, t1);
            if (t1 <= 1.0) add(t1, 1.0);
        }
        else
            add(0.0, 1.0);
    }
    else if (op[1] != 0)
    {
        const double t0 = -op[2] / op[1];
        if (t0 >= 0) add(0, t0);
        if (t0 <= 1.0) add(t0, 1.0);
    }
    else
        m = (pos ? (op[2] >= 0) : (op[2] <= 0)) ? 0xffu : 0u;
    return m;
}

enum { VE_MISS = 0, VE_PENDING = 1, VE_FULL = 2 };

// Straight-line classification of CTCD::vertexEdgeCTCD (src/CTCD.cpp:511-602): VE_MISS when one of its three lists is
// empty or their occupancy masks share no eighth of [0,1]; VE_PENDING when the distance quartic needs the root isolator
// (rec receives its task record); VE_FULL when every list is known and the general routine has to combine them.
CCD_FN int classify_ve(V3 q0s, V3 q1s, V3 q2s, V3 v0, V3 v1, V3 v2, double eta, double *rec_out, bool &want_rec)
{
    const double minD = eta * eta;
    const V3 ab = q2s - q1s, ac = q0s - q1s, cb = q2s - q0s;
    const V3 vab = v2 - v1, vac = v0 - v1, vcb = v2 - v0;
    want_rec = false;
    unsigned occ;
    {
        double op[3];
        op[2] = dot(ab, ac);
        op[1] = dot(ac, vab) + dot(ab, vac);
        op[0] = dot(vab, vac);
        occ = lowdeg_mask(op, true);
        if (!occ) return VE_MISS;
        op[2] = dot(ab, cb);
        op[1] = dot(cb, vab) + dot(ab, vcb);
        op[0] = dot(vab, vcb);
        occ &= lowdeg_mask(op, true);
        if (!occ) return VE_MISS;
    }
    double op[5];
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
    int rd;
    const int r = classify_poly<4>(op, false, rd);
    if (r == PC_EMPTY) return VE_MISS;
    if (r == PC_DECIDED) return VE_FULL;
    occ &= dyadic_mask_reduced<4>(op, rd, false);
    if (!occ) return VE_MISS;
    want_rec = true;
    if (rd == 4) { rec_out[0] = op[0]; rec_out[1] = op[1]; rec_out[2] = op[2]; rec_out[3] = op[3]; rec_out[4] = op[4]; }
    else { rec_out[0] = op[1]; rec_out[1] = op[2]; rec_out[2] = op[3]; rec_out[3] = op[4]; }      // rd == 3
    rec_out[7] = (double)rd;
    return VE_PENDING;
}

// Rebuild pending polynomial k of the primitive and write its task record: normalised coefficients of the reduced
// polynomial + reduced degree — the same values prepare_poly() leaves in Pend.
template <bool IS_VF> CCD_FN void export_poly(int k, const V3 *s, const V3 *v, double eta, double *rec)
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
