// Register-resident finish of a deferred sub-test (narrowphase pass 3): same decisions as vertex_face<RESUME> /
// edge_edge<RESUME> / vertex_edge<RESUME> in ccd_math.cuh, without their local-memory interval lists.
//
// The reference combines its interval lists by testing every combination for pairwise overlap and taking the smallest
// "largest lower end" (src/CTCD.cpp:352-410, :477-507).  For closed intervals on a line that is exactly: the primitive
// hits iff the intersection of the lists' unions is non-empty, and t is the smallest point of that intersection.  So a
// running intersection R (at most RCAP disjoint pieces, in registers) is narrowed polynomial by polynomial, each
// polynomial's intervals being generated on the fly from its breakpoints by the reference's rules (src/CTCD.cpp:161-176)
// with the reference's unfused midpoint evaluation.  Anything unusual — more pieces than RCAP, a NaN, or an edge-edge
// "parallel" interval (which vetoes combinations individually, src/CTCD.cpp:374-390) — returns RS_FALLBACK and the
// caller runs the general routine.
#pragma once
#include "ccd_classify.cuh"

namespace ccd {

enum { RS_MISS = 0, RS_HIT = 1, RS_FALLBACK = 2 };
#define RCAP 4

#define PCAP 3
struct RunSet
{
    double l[RCAP], u[RCAP];
    double pl[PCAP], pu[PCAP];     // edge-edge only: raw coplanarity intervals classified "parallel"
    int n, np;
    bool bad;      // overflow or NaN seen
};

__device__ __forceinline__ void runset_full(RunSet &R)
{
    R.n = 1;
    R.np = 0;
    R.bad = false;
#pragma unroll
    for (int j = 0; j < RCAP; j++) { R.l[j] = 0.0; R.u[j] = 0.0; }
#pragma unroll
    for (int j = 0; j < PCAP; j++) { R.pl[j] = 0.0; R.pu[j] = 0.0; }
    R.l[0] = 0.0;
    R.u[0] = 1.0;
}

// Narrow R by the polynomial op[0..N] (normalised; exactly-zero leading coefficients in place) whose breakpoints are
// time[0..nroots): every interval the reference's rules accept is intersected with the pieces of R.
// EE_SEXTIC: the interval is the raw coplanarity interval of edgeEdgeCTCD; a "parallel" one is set aside.
// One out-of-line copy per (N, EE_SEXTIC) and a rolled loop over the candidate intervals keep the resume kernel small
// (the unrolled version was 19.5 K instructions and spent its time waiting for instruction fetch, profiles/).
template <int N, bool EE_SEXTIC>
static __device__ __noinline__ void narrow(RunSet &R, const double *op_in, bool pos, const double *time, int nroots, V3 ex0, V3 ev0,
                                          V3 ex1, V3 ev1)
{
    double op[N + 1];
#pragma unroll
    for (int i = 0; i <= N; i++) op[i] = op_in[i];
    RunSet Q;
    Q.n = 0;
    Q.np = R.np;
    Q.bad = R.bad;
#pragma unroll
    for (int j = 0; j < RCAP; j++) { Q.l[j] = 0.0; Q.u[j] = 0.0; }
#pragma unroll
    for (int j = 0; j < PCAP; j++) { Q.pl[j] = R.pl[j]; Q.pu[j] = R.pu[j]; }
    // candidate intervals in the reference's order: [0,r0], [r0,r1], ..., [r_last,1]  (src/CTCD.cpp:161-176)
    const int ncand = nroots > 0 ? nroots + 1 : 1;
    for (int i = 0; i < ncand; i++)
    {
        double t1 = 0.0, t2 = 1.0;
        if (nroots > 0)
        {
            if (i == 0)
            {
                t2 = time[0];
                if (!(t2 >= 0)) continue;
            }
            else if (i == nroots)
            {
                t1 = time[nroots - 1];
                if (!(t1 <= 1.0)) continue;
            }
            else
            {
                t1 = time[i - 1];
                t2 = time[i];
                if ((t1 < 0 && t2 < 0) || (t1 > 1.0 && t2 > 1.0)) continue;
            }
        }
        // CTCD::checkInterval + TimeInterval ctor
        t1 = smax(0.0, t1);
        t2 = smax(0.0, t2);
        t1 = smin(1.0, t1);
        t2 = smin(1.0, t2);
        const double tmid = (t2 + t1) / 2;
        double f = op[0];
#pragma unroll
        for (int k = 1; k <= N; k++)
        {
            f *= tmid;
            f += op[k];
        }
        if (!(pos ? (f >= 0) : (f <= 0)))
            continue;
        double l = t1, u = t2;
        if (l > u) { double t = l; l = u; u = t; }
        if (!(l == l) || !(u == u)) { Q.bad = true; continue; }
        if (EE_SEXTIC)
        {
            // parallel-edge classification at the interval midpoint, src/CTCD.cpp:290-308
            const double midt = (u + l) / 2;
            const V3 x10 = ex0 + midt * ev0, x20 = ex1 + midt * ev1;
            const V3 c = cross(x10, x20);
            if (sqrt(dot(c, c)) < 1e-8)
            {
                // a parallel interval never joins a combination; it vetoes the combinations it overlaps (runset_result)
                if (Q.np >= PCAP) Q.bad = true;
#pragma unroll
                for (int k = 0; k < PCAP; k++)
                    if (k == Q.np) { Q.pl[k] = l; Q.pu[k] = u; }
                Q.np++;
                continue;
            }
        }
#pragma unroll
        for (int j = 0; j < RCAP; j++)
            if (j < R.n && !(R.l[j] > u || l > R.u[j]))
            {
                const double nl = smax(R.l[j], l), nu = smin(R.u[j], u);
                if (Q.n >= RCAP) Q.bad = true;
#pragma unroll
                for (int k = 0; k < RCAP; k++)
                    if (k == Q.n) { Q.l[k] = nl; Q.u[k] = nu; }
                Q.n++;
            }
    }
    if (Q.n > RCAP) Q.n = RCAP;
    if (Q.np > PCAP) Q.np = PCAP;
    R = Q;
}

// breakpoints of a normalised polynomial that needs no isolator (classify_poly gave DECIDED): closed forms for reduced
// degree <= 2, none otherwise
template <int N> __device__ __forceinline__ int decided_breakpoints(const double (&op)[N + 1], int rd, double (&time)[6])
{
    if (rd == 2)
    {
        const double a = op[N - 2], b = op[N - 1], c = op[N];
        const double sign = (b < 0) ? -1.0 : 1.0;
        const double D = b * b - 4 * a * c;
        if (D >= 0)
        {
            const double q = -0.5 * (b + sign * sqrt(D));
            double t0 = q / a, t1 = c / q;
            if (t0 > t1) { double t = t0; t0 = t1; t1 = t; }
            time[0] = t0;
            time[1] = t1;
            return 2;
        }
        return 0;
    }
    if (rd == 1)
    {
        time[0] = -op[N] / op[N - 1];
        return 1;
    }
    return 0;
}

// One polynomial of a resumed sub-test: classify (same code as pass 1), fetch or derive its breakpoints, narrow R.
// `rec` walks the task records of the sub-test (one per pending polynomial, in polynomial order).
// Returns false when the polynomial's list is empty (the sub-test misses).
template <int N, bool EE_SEXTIC>
static __device__ __noinline__ bool resume_poly(RunSet &R, double (&op)[N + 1], bool pos, const double *&rec, V3 ex0, V3 ev0, V3 ex1, V3 ev1)
{
    int rd;
    const int cls = classify_poly<N>(op, pos, rd);
    if (cls == PC_EMPTY)
        return false;
    double time[6] = {0, 0, 0, 0, 0, 0};
    int nroots;
    if (cls == PC_PENDING)
    {
        nroots = (int)rec[7];
#pragma unroll
        for (int i = 0; i < 6; i++)
            if (i < nroots) time[i] = rec[i];
        rec += 8;
    }
    else
        nroots = (rd <= 2 && rd >= 1) ? decided_breakpoints<N>(op, rd, time) : 0;
    if (rd == 0)
    {
        // constant polynomial: the reference pushes [0,1] when its sign fits (src/CTCD.cpp:151-158) — nothing to narrow
        if (EE_SEXTIC) R.bad = true;      // its parallel classification is left to the general routine
        return true;
    }
    narrow<N, EE_SEXTIC>(R, op, pos, time, nroots, ex0, ev0, ex1, ev1);
    return R.n > 0 || R.bad;
}

__device__ __forceinline__ int runset_result(const RunSet &R, double &t)
{
    if (R.bad) return RS_FALLBACK;
    // every piece of R is the intersection of one interval per list, i.e. one of the reference's combinations; those
    // that overlap a parallel interval are skipped (src/CTCD.cpp:374-390)
    bool col = false;
    double m = 1.0;
#pragma unroll
    for (int j = 0; j < RCAP; j++)
        if (j < R.n)
        {
            bool skip = false;
#pragma unroll
            for (int k = 0; k < PCAP; k++)
                if (k < R.np && !(R.l[j] > R.pu[k] || R.pl[k] > R.u[j])) skip = true;
            if (!skip) { m = smin(smax(R.l[j], 0.0), m); col = true; }
        }
    if (!col) return RS_MISS;
    t = m;
    return RS_HIT;
}

// the VF / EE primitive with the roots of its pending polynomials in the task records at rec
template <bool IS_VF> __device__ __forceinline__ int resume_primitive(const V3 *s, const V3 *v, double eta, const double *rec, double &t)
{
    RunSet R;
    runset_full(R);
    const V3 z = mk(0, 0, 0);
    if (IS_VF)
    {
        for (int k = 0; k < 3; k++)
        {
            double op[4];
            build_vf_poly(k, s, v, eta, op);
            if (!resume_poly<3, false>(R, op, true, rec, z, z, z, z)) return RS_MISS;
        }
        double op[7];
        build_vf_poly(3, s, v, eta, op);
        if (!resume_poly<6, false>(R, op, false, rec, z, z, z, z)) return RS_MISS;
    }
    else
    {
        // task records are in polynomial-index order (quartics 0..3, then the sextic = index 4), while the walk below
        // starts with the sextic: count the pending quartics first to find the sextic's record
        // -> simpler: evaluate in index order, the sextic last; the parallel classification only needs the sextic's
        //    own intervals, and the intersection is order-independent.
        for (int k = 0; k < 4; k++)
        {
            double op[5];
            build_ee_poly(k, s, v, eta, op);
            if (!resume_poly<4, false>(R, op, true, rec, z, z, z, z)) return RS_MISS;
        }
        double op[7];
        build_ee_poly(4, s, v, eta, op);
        if (!resume_poly<6, true>(R, op, false, rec, s[0] - s[1], v[0] - v[1], s[2] - s[3], v[2] - v[3])) return RS_MISS;
    }
    return runset_result(R, t);
}

// CTCD::vertexEdgeCTCD with the roots of its distance quartic in the task record at rec
__device__ __forceinline__ int resume_ve(V3 q0s, V3 q1s, V3 q2s, V3 v0, V3 v1, V3 v2, double eta, const double *rec, double &t)
{
    const double minD = eta * eta;
    const V3 ab = q2s - q1s, ac = q0s - q1s, cb = q2s - q0s;
    const V3 vab = v2 - v1, vac = v0 - v1, vcb = v2 - v0;
    const V3 z = mk(0, 0, 0);
    RunSet R;
    runset_full(R);
    {
        double op[3];
        op[2] = dot(ab, ac);
        op[1] = dot(ac, vab) + dot(ab, vac);
        op[0] = dot(vab, vac);
        if (!resume_poly<2, false>(R, op, true, rec, z, z, z, z)) return RS_MISS;
        op[2] = dot(ab, cb);
        op[1] = dot(cb, vab) + dot(ab, vcb);
        op[0] = dot(vab, vcb);
        if (!resume_poly<2, false>(R, op, true, rec, z, z, z, z)) return RS_MISS;
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
    if (!resume_poly<4, false>(R, op, false, rec, z, z, z, z)) return RS_MISS;
    return runset_result(R, t);
}

} // namespace ccd
