// Per-item bodies of the narrowphase stage kernels (narrowphase.cu).  Plain arithmetic, host+device: the kernels add only
// the queue plumbing, and tests/np_emul.cu runs the very same functions on the CPU against the checker.
//
// A sub-test (the VF / EE primitive or one vertex-edge test) is a conjunction: every one of its polynomials must be
// satisfied (>= 0 or <= 0) at a common time in [0,1] (src/CTCD.cpp:352-410, :477-507, :575-600).  The pipeline settles it
// polynomial by polynomial:
//   stage_item     one polynomial of the primitive for one stencil: EMPTY (no interval: the sub-test misses, the stencil
//                  leaves the pipeline), whole [0,1] (no constraint), or "needs a record" — its intervals come from
//                  closed forms (reduced degree <= 2) or from the root isolator (degree 3..6).  A 3-level dyadic
//                  occupancy mask per polynomial is AND-ed along the way: when no eighth of [0,1] can satisfy all
//                  polynomials seen so far the sub-test misses without any root being isolated.
//   export_item    rebuilds one polynomial of a surviving stencil and writes its 64-byte record.
//   ve_item        the three polynomials of CTCD::vertexEdgeCTCD in one go (two quadratics + the distance quartic).
//   combine_records  intersection of the interval lists of a sub-test -> hit / time of impact, with the edge-edge
//                  "parallel" veto of src/CTCD.cpp:290-308, :374-390.
#pragma once
#include "ccd_classify.cuh"
#include "ccd_solve.cuh"

namespace ccd {

enum { RS_MISS = 0, RS_HIT = 1, RS_FALLBACK = 2 };
// status codes of a sub-test after pass 1 (2 bits each in the stencil's status word)
enum { SC_MISS = 0, SC_HIT = 1, SC_DEFERRED = 2, SC_GENERAL = 3 };

template <bool IS_VF> struct Prim
{
    static constexpr int NST = IS_VF ? 4 : 5;      // polynomials = stages
    // polynomial handled by stage s: VF e1,e2,e3,coplanarity (src/CTCD.cpp:433-475); EE a0,a1,b0,b1 and the distance sextic
    // LAST (the reference tests it first, src/CTCD.cpp:266-350, but the verdict is a conjunction: once the swept-box cull
    // has run the sextic rejects only ~6 % of the stencils while each quartic rejects ~20 %, so the expensive polynomial
    // is built for half as many stencils this way — and its record is written by the stage itself, not rebuilt)
    static CCD_HD constexpr int poly(int s) { return s; }
    static CCD_HD constexpr int degree(int k) { return IS_VF ? (k < 3 ? 3 : 6) : (k < 4 ? 4 : 6); }
    static CCD_HD constexpr bool pos(int k) { return IS_VF ? (k < 3) : (k < 4); }
};

template <int N> CCD_FN void pending_record(const double (&op)[N + 1], int rd, unsigned tag, double (&rec)[8])
{
#pragma unroll
    for (int c = 0; c <= N; c++)
    {
        // rec[c] = op[c + N - rd] for c <= rd
        double x = 0.0;
#pragma unroll
        for (int j = 0; j <= N; j++)
            if (j == c + N - rd) x = op[j];
        rec[c] = x;
    }
#pragma unroll
    for (int c = N + 1; c < 7; c++) rec[c] = 0.0;
    rec[7] = rec_tag(tag);
}

// intervals of a normalised polynomial of reduced degree 1 or 2 (closed-form breakpoints, src/CTCD.cpp:145-176)
template <int N> CCD_FN void lowdeg_intervals(const double (&op)[N + 1], int rd, bool pos, Ivl3 &o)
{
    double time[6] = {0, 0, 0, 0, 0, 0};
    int nroots = 0;
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
            nroots = 2;
            if (!(t0 == t0) || !(t1 == t1))
            {
                // 0/0 somewhere: the reference's comparisons with NaN are left to the general routine
                o.n = 0;
                o.bad = true;
                return;
            }
        }
    }
    else
    {
        time[0] = -op[N] / op[N - 1];
        nroots = 1;
        if (!(time[0] == time[0])) { o.n = 0; o.bad = true; return; }
    }
    intervals_from_breakpoints<N>(op, pos, time, nroots, o);
}

CCD_FN bool ivl_is_whole(const Ivl3 &o) { return !o.bad && o.n == 1 && o.l[0] == 0.0 && o.u[0] == 1.0; }

CCD_FN unsigned ivl_mask(const Ivl3 &o)
{
    if (o.bad) return 0xffu;
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (k < o.n) m |= interval_mask(o.l[k], o.u[k]);
    return m;
}

// One polynomial of the primitive.  state: bits 0-4 = polynomials that need a record, bits 8-15 = occupancy mask.
// Returns false when the stencil's primitive is known to miss.  WANT_REC: also produce this polynomial's record
// (has_rec tells whether it has one).
template <bool IS_VF, int S, bool WANT_REC>
CCD_FN bool stage_item(const V3 *a, const V3 *v, double eta, unsigned state_in, unsigned &state_out, double (&rec)[8], bool &has_rec)
{
    constexpr int K = Prim<IS_VF>::poly(S), N = Prim<IS_VF>::degree(K);
    constexpr bool POS = Prim<IS_VF>::pos(K);
    double op[N + 1];
    if (IS_VF) build_vf_poly(K, a, v, eta, op); else build_ee_poly(K, a, v, eta, op);
    int rd;
    const int r = classify_poly<N>(op, POS, rd);
    has_rec = false;
    if (r == PC_EMPTY) return false;
    unsigned need = state_in & 0x1fu, occ = (state_in >> 8) & 0xffu;
    if (r == PC_PENDING)
    {
        need |= 1u << K;
        occ &= dyadic_mask_reduced<N>(op, rd, POS, occ);
        if (!occ) return false;
        if (WANT_REC) { pending_record<N>(op, rd, make_tag(K, POS, rd), rec); has_rec = true; }
    }
    else if (rd == 1 || rd == 2)
    {
        Ivl3 o;
        lowdeg_intervals<N>(op, rd, POS, o);
        if (!ivl_is_whole(o))
        {
            need |= 1u << K;
            occ &= ivl_mask(o);
            if (!occ) return false;
            if (WANT_REC) { write_final_record(rec, o, make_tag(K, POS, rd)); has_rec = true; }
        }
    }
    state_out = need | (occ << 8);
    return true;
}

// the record of polynomial K for a stencil that survived every stage with bit K set in its need mask
template <bool IS_VF, int K> CCD_FN void export_item(const V3 *a, const V3 *v, double eta, double (&rec)[8])
{
    constexpr int N = Prim<IS_VF>::degree(K);
    constexpr bool POS = Prim<IS_VF>::pos(K);
    double op[N + 1];
    if (IS_VF) build_vf_poly(K, a, v, eta, op); else build_ee_poly(K, a, v, eta, op);
    int rd;
    const int r = classify_poly<N>(op, POS, rd);
    if (r == PC_PENDING)
        pending_record<N>(op, rd, make_tag(K, POS, rd), rec);
    else
    {
        Ivl3 o;
        o.n = 0;
        o.bad = true;
        if (r == PC_DECIDED && (rd == 1 || rd == 2)) lowdeg_intervals<N>(op, rd, POS, o);
        write_final_record(rec, o, make_tag(K, POS, rd));
    }
}

// ---- interval combination ------------------------------------------------------------------------------------
// The reference tests every combination of one interval per list for pairwise overlap and takes the smallest "largest
// lower end" (src/CTCD.cpp:352-410, :477-507).  For closed intervals on a line that is exactly: hit iff the
// intersection of the lists' unions is non-empty, t = its smallest point.  A running intersection R (at most RCAP
// disjoint pieces, in registers) is narrowed list by list.  Anything unusual — more pieces than RCAP, more parallel
// intervals than PCAP, a BAD record — gives RS_FALLBACK and the stencil is redone by the general routine.
#define RCAP 4
#define PCAP 3
struct RunSet
{
    double l[RCAP], u[RCAP];
    double pl[PCAP], pu[PCAP];     // edge-edge only: raw coplanarity intervals classified "parallel"
    int n, np;
    bool bad;
};

CCD_FN void runset_full(RunSet &R)
{
    R.n = 1;
    R.np = 0;
    R.bad = false;
#pragma unroll
    for (int j = 0; j < RCAP; j++) { R.l[j] = 0.0; R.u[j] = 0.0; }
#pragma unroll
    for (int j = 0; j < PCAP; j++) { R.pl[j] = 0.0; R.pu[j] = 0.0; }
    R.u[0] = 1.0;
}

// R := R ∩ (union of iv).  EE_SEXTIC: iv are the raw coplanarity intervals of edgeEdgeCTCD; those at whose midpoint the
// edges are parallel (|| x10 x x20 || < 1e-8, src/CTCD.cpp:294-308) are set aside instead.
CCD_FN void narrow_ivl(RunSet &R, const Ivl3 &iv, bool EE_SEXTIC, V3 ex0, V3 ev0, V3 ex1, V3 ev1)
{
    RunSet Q;
    Q.n = 0;
    Q.np = R.np;
    Q.bad = R.bad || iv.bad;
#pragma unroll
    for (int j = 0; j < RCAP; j++) { Q.l[j] = 0.0; Q.u[j] = 0.0; }
#pragma unroll
    for (int j = 0; j < PCAP; j++) { Q.pl[j] = R.pl[j]; Q.pu[j] = R.pu[j]; }
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
        if (i >= iv.n) continue;
        const double l = iv.l[i], u = iv.u[i];
        if (EE_SEXTIC)
        {
            const double midt = (u + l) / 2;
            const V3 x10 = ex0 + midt * ev0, x20 = ex1 + midt * ev1;
            const V3 c = cross(x10, x20);
            if (sqrt(dot(c, c)) < 1e-8)
            {
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

CCD_FN int runset_result(const RunSet &R, double &t)
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

// nrec final records of one sub-test.  ee_prim: the sub-test is the edge-edge primitive on points s[0..3] =
// (q0,p0,q1,p1); its sextic (polynomial 4) carries the raw coplanarity intervals — the whole [0,1] when it has no record.
CCD_FN int combine_records(const double *rec, int nrec, bool ee_prim, const V3 *s, const V3 *v, double &t)
{
    RunSet R;
    runset_full(R);
    const V3 z = mk(0, 0, 0);
    V3 ex0 = z, ev0 = z, ex1 = z, ev1 = z;
    if (ee_prim) { ex0 = s[0] - s[1]; ev0 = v[0] - v[1]; ex1 = s[2] - s[3]; ev1 = v[2] - v[3]; }
    bool saw_sextic = false;
    for (int j = 0; j < nrec; j++)
    {
        Ivl3 iv;
        unsigned tag;
        read_final_record(rec + REC_STRIDE * j, iv, tag);
        if (iv.bad) return RS_FALLBACK;
        const bool sext = ee_prim && (tag & 7u) == 4u;
        saw_sextic = saw_sextic || sext;
        narrow_ivl(R, iv, sext, ex0, ev0, ex1, ev1);
        if (R.n == 0 && !R.bad) return RS_MISS;
    }
    if (ee_prim && !saw_sextic)
    {
        Ivl3 whole;
        whole.n = 1;
        whole.bad = false;
        whole.l[0] = 0.0; whole.u[0] = 1.0;
        whole.l[1] = whole.u[1] = whole.l[2] = whole.u[2] = 0.0;
        narrow_ivl(R, whole, true, ex0, ev0, ex1, ev1);
    }
    return runset_result(R, t);
}

// ---- window stage -----------------------------------------------------------------------------------------------
// After the distance polynomial of a primitive (VF coplanarity sextic / EE line-distance sextic) has been solved, the
// only times that can still matter are its few, usually very short, intervals W_j ("windows").  An inside polynomial
// q (cubic / quartic, ">= 0 wanted") that keeps one sign on a window needs no root isolation there: its Bernstein
// coefficients on W_j (two de Casteljau splits) all beyond +-1e-12 of the normalised scale prove the sign on the whole
// closed window (convex hull property; the margin is ~10^3 rounding errors of the splits).  If no window is mixed, q's
// record is finalised with the windows on which it is satisfied AS ITS INTERVALS: intersecting with them keeps exactly
// those windows, bit for bit, which is all the exact interval list of q could do to them (it covers a window where q > 0
// and misses a window where q < 0).  Only polynomials with a sign change near a window still go to the root isolator.
// Bernstein coefficients b on some interval -> on its sub-interval [wl, wu] (in the interval's own [0,1] parameter)
template <int RD> CCD_FN void bernstein_restrict(double (&b)[RD + 1], double wl, double wu)
{
    // keep [wl, 1]: after step k, b[0..RD-k] are the level-k points; the right part collects the last point of each level
    {
        double r[RD + 1];
#pragma unroll
        for (int k = 0; k <= RD; k++)
        {
            r[RD - k] = b[RD - k];
#pragma unroll
            for (int i = 0; i < RD - k; i++)
                b[i] = b[i] + wl * (b[i + 1] - b[i]);
        }
#pragma unroll
        for (int i = 0; i <= RD; i++) b[i] = r[i];
    }
    // keep the first (wu-wl)/(1-wl) of it: the left part collects the first point of each level
    {
        const double den = 1.0 - wl;
        double s = den > 0.0 ? (wu - wl) / den : 0.0;
        s = smin(smax(s, 0.0), 1.0);
        double l[RD + 1];
#pragma unroll
        for (int k = 0; k <= RD; k++)
        {
            l[k] = b[0];
#pragma unroll
            for (int i = 0; i < RD - k; i++)
                b[i] = b[i] + s * (b[i + 1] - b[i]);
        }
#pragma unroll
        for (int i = 0; i <= RD; i++) b[i] = l[i];
    }
}

// Bernstein coefficients of c (degree RD, descending power coefficients) on the window [wl, wu] of [0,1]
template <int RD> CCD_FN void window_bernstein(const double (&c)[RD + 1], double wl, double wu, double (&b)[RD + 1])
{
#pragma unroll
    for (int i = 0; i <= RD; i++)
        b[i] = c[RD - i] * rbinom(RD, i);
#pragma unroll
    for (int k = 1; k <= RD; k++)
#pragma unroll
        for (int i = RD; i >= k; i--)
            b[i] = b[i] + b[i - 1];
    bernstein_restrict<RD>(b, wl, wu);
}

template <int RD> CCD_FN int window_class(const double (&c)[RD + 1], bool pos, double wl, double wu)
{
    double b[RD + 1];
    window_bernstein<RD>(c, wl, wu, b);
    bool allp = true, alln = true;
#pragma unroll
    for (int i = 0; i <= RD; i++)
    {
        allp = allp && (b[i] > 1e-12);
        alln = alln && (b[i] < -1e-12);
    }
    if (pos ? allp : alln) return 1;       // satisfied on the whole window
    if (pos ? alln : allp) return -1;      // violated on the whole window
    return 0;                              // mixed / too close to call
}

// true when c cannot be satisfied anywhere on [wl, wu]: the Bernstein hull has the wrong sign by more than 1e-12 on every
// 64th of the window (three dyadic levels, then three more inside each eighth that is still possible)
template <int RD> CCD_FN bool window_excluded(const double (&c)[RD + 1], bool pos, double wl, double wu)
{
    double b[RD + 1];
    window_bernstein<RD>(c, wl, wu, b);
    unsigned m = dyadic_sub<RD, 3>(b, pos);
    while (m)
    {
        const int k = ccd_ffs(m) - 1;
        m &= m - 1;
        double bb[RD + 1];
#pragma unroll
        for (int i = 0; i <= RD; i++) bb[i] = b[i];
        bernstein_restrict<RD>(bb, 0.125 * (double)k, 0.125 * (double)(k + 1));
        if (dyadic_sub<RD, 3>(bb, pos)) return false;
    }
    return true;
}

CCD_FN int window_class_rd(const double *rec, int rd, bool pos, double wl, double wu)
{
    if (rd == 3)
    {
        double c[4] = {rec[0], rec[1], rec[2], rec[3]};
        return window_class<3>(c, pos, wl, wu);
    }
    if (rd == 4)
    {
        double c[5] = {rec[0], rec[1], rec[2], rec[3], rec[4]};
        return window_class<4>(c, pos, wl, wu);
    }
    return 0;
}

// The records of one deferred primitive, in place: rec[0..nrec) with the distance polynomial (index KD) last.  Nothing
// happens unless that record is final and clean.
CCD_FN void window_item(double *rec, int nrec, int KD)
{
    if (nrec < 2) return;
    Ivl3 W;
    unsigned wtag;
    read_final_record(rec + REC_STRIDE * (nrec - 1), W, wtag);
    if (W.bad || (int)(wtag & 7u) != KD) return;
    bool alive[3] = {W.n > 0, W.n > 1, W.n > 2};
    for (int j = 0; j < nrec - 1; j++)
    {
        double *r = rec + REC_STRIDE * j;
        const unsigned tag = rec_untag(r[7]);
        if (tag & REC_FINAL) continue;
        const int rd = (int)((tag >> 4) & 7u);
        const bool pos = (tag & REC_POS) != 0;
        if (rd != 3 && rd != 4) continue;
        int cls[3] = {-1, -1, -1};
        bool mixed = false;
#pragma unroll
        for (int k = 0; k < 3; k++)
            if (alive[k])
            {
                cls[k] = window_class_rd(r, rd, pos, W.l[k], W.u[k]);
                mixed = mixed || cls[k] == 0;
            }
        Ivl3 o;
        o.n = 0;
        o.bad = false;
#pragma unroll
        for (int k = 0; k < 3; k++) { o.l[k] = 0.0; o.u[k] = 0.0; }
#pragma unroll
        for (int k = 0; k < 3; k++)
        {
            if (alive[k] && cls[k] < 0) alive[k] = false;      // this polynomial rules the window out, solved or not
            if (alive[k] && cls[k] > 0 && !mixed) ivl_push(o, W.l[k], W.u[k]);
        }
        if (!mixed) write_final_record(r, o, tag);
    }
}

// CTCD::vertexEdgeCTCD (src/CTCD.cpp:511-602), vertex q0 against segment (q1,q2); v* = end - start.
// Returns SC_MISS, SC_DEFERRED (nrec records in recs: polynomial 0,1 = the two inside quadratics, 2 = the distance
// quartic; whole-[0,1] lists have no record — possibly none at all).
CCD_FN int ve_item(V3 q0s, V3 q1s, V3 q2s, V3 v0, V3 v1, V3 v2, double eta, double (&recs)[3][8], int &nrec)
{
    const double minD = eta * eta;
    const V3 ab = q2s - q1s, ac = q0s - q1s, cb = q2s - q0s;
    const V3 vab = v2 - v1, vac = v0 - v1, vcb = v2 - v0;
    nrec = 0;
    unsigned occ = 0xffu;
    int rd;
    {
        double op[3];
        op[2] = dot(ab, ac);
        op[1] = dot(ac, vab) + dot(ab, vac);
        op[0] = dot(vab, vac);
        if (classify_poly<2>(op, true, rd) == PC_EMPTY) return SC_MISS;
        if (rd >= 1)
        {
            Ivl3 o;
            lowdeg_intervals<2>(op, rd, true, o);
            if (!ivl_is_whole(o))
            {
                occ &= ivl_mask(o);
                write_final_record(recs[nrec], o, make_tag(0, true, rd));
                nrec++;
            }
        }
    }
    {
        double op[3];
        op[2] = dot(ab, cb);
        op[1] = dot(cb, vab) + dot(ab, vcb);
        op[0] = dot(vab, vcb);
        if (classify_poly<2>(op, true, rd) == PC_EMPTY) return SC_MISS;
        if (rd >= 1)
        {
            Ivl3 o;
            lowdeg_intervals<2>(op, rd, true, o);
            if (!ivl_is_whole(o))
            {
                occ &= ivl_mask(o);
                if (!occ) return SC_MISS;
                if (nrec == 0) write_final_record(recs[0], o, make_tag(1, true, rd));
                else write_final_record(recs[1], o, make_tag(1, true, rd));
                nrec++;
            }
        }
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
    const int r = classify_poly<4>(op, false, rd);
    if (r == PC_EMPTY) return SC_MISS;
    if (r == PC_PENDING)
    {
        occ &= dyadic_mask_reduced<4>(op, rd, false, occ);
        if (!occ) return SC_MISS;
        double rec[8];
        pending_record<4>(op, rd, make_tag(2, false, rd), rec);
        rec[5] = (double)nrec;      // sibling records (the inside quadratics) right before this one: see ve_refine_item
#pragma unroll
        for (int k = 0; k < 3; k++)
            if (k == nrec)
            {
#pragma unroll
                for (int c = 0; c < 8; c++) recs[k][c] = rec[c];
            }
        nrec++;
    }
    else if (rd == 1 || rd == 2)
    {
        Ivl3 o;
        lowdeg_intervals<4>(op, rd, false, o);
        if (!ivl_is_whole(o))
        {
            occ &= ivl_mask(o);
            if (!occ) return SC_MISS;
            double rec[8];
            write_final_record(rec, o, make_tag(2, false, rd));
#pragma unroll
            for (int k = 0; k < 3; k++)
                if (k == nrec)
                {
#pragma unroll
                    for (int c = 0; c < 8; c++) recs[k][c] = rec[c];
                }
            nrec++;
        }
    }
    return SC_DEFERRED;      // nrec == 0: nothing constrains the test — combining zero records gives the whole [0,1], a hit at t = 0
}

// A pending vertex-edge distance quartic (record `rec`, its sub-test's inside-quadratic records right before it).  The
// quartic dips wherever the vertex sweeps over the edge's LINE, usually without coming within eta of it, and on [0,1] its
// Bernstein hull is rarely conclusive.  Only the times at which the vertex projects inside the segment count ([0,1]
// narrowed by the quadratics' intervals), and on those windows the hull, refined down to 64ths, usually proves "further
// than eta": the record is then finalised with no interval and no root is isolated.  Returns true in that case.
template <int STRIDE> CCD_FN bool ve_refine_item_s(double *rec)
{
    const unsigned tag = rec_untag(rec[7]);
    const int rd = (int)((tag >> 4) & 7u);
    const int nsib = (int)rec[5];
    if ((tag & (REC_FINAL | REC_POS)) || (rd != 3 && rd != 4) || nsib < 0 || nsib > 2) return false;
    RunSet R;
    runset_full(R);
    const V3 z = mk(0, 0, 0);
    for (int s = nsib; s >= 1; s--)
    {
        Ivl3 o;
        unsigned t;
        read_final_record(rec - STRIDE * s, o, t);
        if (o.bad) return false;
        narrow_ivl(R, o, false, z, z, z, z);
    }
    if (R.bad) return false;
    bool excluded = true;
    for (int j = 0; j < RCAP && excluded; j++)
        if (j < R.n)
        {
            double lj = R.l[0], uj = R.u[0];
#pragma unroll
            for (int k = 1; k < RCAP; k++)
                if (k == j) { lj = R.l[k]; uj = R.u[k]; }
            if (rd == 4)
            {
                const double c4[5] = {rec[0], rec[1], rec[2], rec[3], rec[4]};
                excluded = window_excluded<4>(c4, false, lj, uj);
            }
            else
            {
                const double c3[4] = {rec[0], rec[1], rec[2], rec[3]};
                excluded = window_excluded<3>(c3, false, lj, uj);
            }
        }
    if (!excluded) return false;
    Ivl3 none;
    none.n = 0;
    none.bad = false;
#pragma unroll
    for (int k = 0; k < 3; k++) { none.l[k] = 0.0; none.u[k] = 0.0; }
    write_final_record(rec, none, tag);
    return true;
}

CCD_FN bool ve_refine_item(double *rec) { return ve_refine_item_s<REC_STRIDE>(rec); }

// ve_item followed at once by the refinement of its pending distance quartic on the windows its inside quadratics leave:
// a test whose quartic is proven "further than eta" there has no common time at all — combining its records would give
// RS_MISS — so it is reported as SC_MISS and none of its records is ever written (four out of five deferred vertex-edge
// tests of the 4M-triangle cloth end this way).
CCD_FN int ve_item_refined(V3 q0s, V3 q1s, V3 q2s, V3 v0, V3 v1, V3 v2, double eta, double (&recs)[3][8], int &nrec)
{
    const int code = ve_item(q0s, q1s, q2s, v0, v1, v2, eta, recs, nrec);
    if (code != SC_DEFERRED || nrec == 0) return code;
    bool gone = false;
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (k == nrec - 1)
        {
            const unsigned tag = rec_untag(recs[k][7]);
            if (!(tag & (REC_FINAL | REC_POS)) && (tag & 7u) == 2u) gone = ve_refine_item_s<8>(&recs[k][0]);
        }
    return gone ? SC_MISS : code;
}

// turn a pending record into a final one: isolate the roots of its polynomial, apply the interval rules
template <int D> CCD_FN void finalize_record(const double (&c)[D + 1], unsigned tag, const double (&roots)[6], int nr, double *rec)
{
    Ivl3 o;
    intervals_from_breakpoints<D>(c, (tag & REC_POS) != 0, roots, nr, o);
    write_final_record(rec, o, tag);
}

} // namespace ccd
