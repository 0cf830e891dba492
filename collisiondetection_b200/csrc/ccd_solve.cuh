// Root isolation as a per-lane state machine + the interval rules of CTCD::findIntervals, for the dense solve kernels.
//
// Same steps and the same roundings as roots01 (ccd_math.cuh) and orc_roots01 in the CPU checker — only the control flow
// differs.  A form that inlines one bracketed Newton solve per monotone piece and per derivative level never lets lanes on
// different pieces of different levels share instructions (ncu: 6.7 of 32 lanes active in roots_kernel<6>,
// profiles/r01_narrowphase_v9_c5.txt).  Here a lane alternates between
//     advance()      walk its pieces / levels until the next sign change that needs a solve (or the end), and
//     newton_step()  one iteration of that solve,
// so a warp runs ONE copy of the Newton iteration for all lanes that have a solve pending, whatever level or piece
// each of them is at, and lanes that finish a polynomial pick up the next one (solve_kernel in narrowphase.cu).
#pragma once
#include "ccd_roots_t.cuh"

namespace ccd {

// ---- task / interval records (REC_STRIDE doubles = 128 bytes; the first 8 are "the record") -------------------------
// scratch of a pending record (second half): after `prepare` d[8], d[9] = roots of the quadratic level, d[10] = packed
// start state; after `solve` d[8..13] = roots in [0,1], d[14] = their number
// pending: d[0..rd] = normalised coefficients of the reduced polynomial (descending), d[7] = tag
// final:   d[0] = number of intervals n (<= 3), d[1+2j], d[2+2j] = [l_j, u_j] (closed, clamped to [0,1], in the
//          reference's order), d[7] = tag | REC_FINAL (| REC_BAD when n > 3 or a NaN was seen: the stencil is then
//          redone by the general routine)
// tag bits: 0-2 polynomial index inside its sub-test, 3 REC_POS (intervals where the polynomial is >= 0, else <= 0),
//           4-6 reduced degree, 8 REC_FINAL, 9 REC_BAD
enum { REC_POS = 1 << 3, REC_FINAL = 1 << 8, REC_BAD = 1 << 9 };
enum { REC_STRIDE = 16 };
CCD_FN double rec_tag(unsigned t) { return (double)t; }
CCD_FN unsigned rec_untag(double d) { return (unsigned)d; }
CCD_FN unsigned make_tag(int k, bool pos, int rd) { return (unsigned)k | (pos ? (unsigned)REC_POS : 0u) | ((unsigned)rd << 4); }

struct Ivl3
{
    double l[3], u[3];
    int n;
    bool bad;
};

CCD_FN void ivl_push(Ivl3 &o, double t1, double t2)
{
    // TimeInterval ctor, include/CTCD.h:9-14 (t1,t2 already clamped by the caller as checkInterval does)
    double l = t1, u = t2;
    if (l > u) { double t = l; l = u; u = t; }
    if (!(l == l) || !(u == u)) { o.bad = true; return; }
    if (o.n >= 3) { o.bad = true; return; }
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (k == o.n) { o.l[k] = l; o.u[k] = u; }
    o.n++;
}

// The interval rules of src/CTCD.cpp:161-176 for the normalised polynomial op[0..N] (exactly-zero leading coefficients
// may stay in place: 0*t + x == x) and its breakpoints time[0..nroots): every piece whose clamped midpoint satisfies
// the sign test (CTCD::checkInterval, unfused Horner) becomes an interval.
template <int N> CCD_FN void intervals_from_breakpoints(const double (&op)[N + 1], bool pos, const double (&time)[6], int nroots, Ivl3 &o)
{
    o.n = 0;
    o.bad = false;
#pragma unroll
    for (int k = 0; k < 3; k++) { o.l[k] = 0.0; o.u[k] = 0.0; }
    const int ncand = nroots > 0 ? nroots + 1 : 1;
    for (int i = 0; i < ncand; i++)
    {
        double t1 = 0.0, t2 = 1.0;
        if (nroots > 0)
        {
            double ta = 0.0, tb = 0.0;      // time[i-1], time[i]
#pragma unroll
            for (int k = 0; k < 6; k++)
            {
                if (k == i - 1) ta = time[k];
                if (k == i) tb = time[k];
            }
            if (i == 0)
            {
                t2 = tb;
                if (!(t2 >= 0)) continue;
            }
            else if (i == nroots)
            {
                t1 = ta;
                if (!(t1 <= 1.0)) continue;
            }
            else
            {
                t1 = ta;
                t2 = tb;
                if ((t1 < 0 && t2 < 0) || (t1 > 1.0 && t2 > 1.0)) continue;
            }
        }
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
        if (pos ? (f >= 0) : (f <= 0))
            ivl_push(o, t1, t2);
    }
}

CCD_FN void write_final_record(double *rec, const Ivl3 &o, unsigned tag)
{
    rec[0] = (double)o.n;
    rec[1] = o.l[0]; rec[2] = o.u[0];
    rec[3] = o.l[1]; rec[4] = o.u[1];
    rec[5] = o.l[2]; rec[6] = o.u[2];
    rec[7] = rec_tag(tag | REC_FINAL | (o.bad ? (unsigned)REC_BAD : 0u));
}

CCD_FN void read_final_record(const double *rec, Ivl3 &o, unsigned &tag)
{
    tag = rec_untag(rec[7]);
    o.n = (int)rec[0];
    o.l[0] = rec[1]; o.u[0] = rec[2];
    o.l[1] = rec[3]; o.u[1] = rec[4];
    o.l[2] = rec[5]; o.u[2] = rec[6];
    o.bad = (tag & REC_BAD) != 0 || !(tag & REC_FINAL);
}

// ---- the state machine ---------------------------------------------------------------------------------------
// STRIDE == 0: the polynomial and the two root lists live in the struct (registers, indexed through unrolled selects);
// STRIDE > 0: they live in memory the caller binds (shared memory in solve_kernel, element k of a lane at [k * STRIDE]),
// indexed directly — fewer registers (more resident warps for the latency-bound Newton chains) and no select chains.
template <int D, int STRIDE = 0> struct RootLane
{
    double c[D + 1];      // the polynomial
    double p[D + 1];      // derivative level m, right-aligned (deriv_level_padded)
    double cur[D];        // roots of the level below, ascending
    double out[D];        // roots found at this level so far
    double *mc, *mcur, *mout;      // STRIDE > 0
    double prev;          // out[nr-1]
    double brk_lo, f_lo, x_hi, f_hi;
    double lo, hi, x, dx, dxold;      // the solve in flight
    int ncur, nr, m, m0, i, nb, it;
    bool one, plain, last, lo_neg, solving, done, got_root;      // got_root: x holds a converged root not yet recorded

    CCD_FN void bind(double *mem)
    {
        mc = mem;
        mcur = mem + (D + 1) * STRIDE;
        mout = mem + (2 * D + 1) * STRIDE;
    }
    CCD_FN double cur_at(int j) const
    {
        if (STRIDE) return mcur[j * STRIDE];
        double r = cur[0];
#pragma unroll
        for (int k = 1; k < D; k++)
            if (k == j) r = cur[k];
        return r;
    }
    CCD_FN void set_cur(int j, double r)
    {
        if (STRIDE) { mcur[j * STRIDE] = r; return; }
#pragma unroll
        for (int k = 0; k < D; k++)
            if (k == j) cur[k] = r;
    }
    CCD_FN double out_at(int j) const
    {
        if (STRIDE) return mout[j * STRIDE];
        double r = out[0];
#pragma unroll
        for (int k = 1; k < D; k++)
            if (k == j) r = out[k];
        return r;
    }
    CCD_FN void push_out(double r)
    {
        if (STRIDE)
        {
            if (nr < D) mout[nr * STRIDE] = r;
        }
        else
        {
#pragma unroll
            for (int k = 0; k < D; k++)
                if (k == nr) out[k] = r;
        }
        prev = r;
        nr++;
    }
    CCD_FN void load_c(double (&t)[D + 1]) const
    {
#pragma unroll
        for (int k = 0; k <= D; k++) t[k] = STRIDE ? mc[k * STRIDE] : c[k];
    }
    CCD_FN void setup_level()
    {
        plain = one && m == m0;
        last = (m == D);
        {
            double t[D + 1];
            load_c(t);
            deriv_level_padded<D>(t, m, p);
        }
        nb = plain ? 2 : ncur + 2;
        i = 0;
        brk_lo = 0.0;
        f_lo = horner_padded<D>(p, 0.0);
        nr = 0;
        prev = 0.0;
    }
    CCD_FN void next_piece()
    {
        i++;
        brk_lo = x_hi;
        f_lo = f_hi;
    }

    // Bernstein sign variations down the derivative chain (roots01, first half): finds the first level to climb
    CCD_FN void prepare(const double (&coef)[D + 1])
    {
        double b[D + 1];
#pragma unroll
        for (int k = 0; k <= D; k++) c[k] = coef[k];
#pragma unroll
        for (int k = 0; k <= D; k++)
            b[k] = c[D - k] * rbinom(D, k);
#pragma unroll
        for (int k = 1; k <= D; k++)
#pragma unroll
            for (int j = D; j >= k; j--)
                b[j] = b[j] + b[j - 1];
#pragma unroll
        for (int k = 0; k < D; k++) { cur[k] = 0.0; out[k] = 0.0; }
        ncur = 0;
        one = false;
        solving = false;
        done = false;
        got_root = false;
        for (m0 = D; m0 >= 2; m0--)
        {
            if (m0 < D)
            {
#pragma unroll
                for (int k = 0; k < D; k++)
                    if (k <= m0) b[k] = b[k + 1] - b[k];
            }
            double bend = b[0];
#pragma unroll
            for (int k = 1; k <= D; k++)
                if (k == m0) bend = b[k];
            if (b[0] != 0.0 && bend != 0.0)
            {
                int v = 0, lastsign = 0;
#pragma unroll
                for (int k = 0; k <= D; k++)
                    if (k <= m0)
                    {
                        const int s = (b[k] > 0.0) - (b[k] < 0.0);
                        if (s != 0)
                        {
                            if (lastsign != 0 && s != lastsign) v++;
                            lastsign = s;
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
    }
    // first level to climb after prepare() / load_start()
    CCD_FN void start()
    {
        solving = false;
        done = false;
        got_root = false;
        m = one ? m0 : m0 + 1;
        if (m > D) { done = true; return; }
        setup_level();
    }
    CCD_FN void begin(const double (&coef)[D + 1])
    {
        prepare(coef);
        start();
    }
    // the outcome of prepare() in three doubles, and back
    CCD_FN void save_start(double *aux) const
    {
        aux[0] = cur[0];
        aux[1] = cur[1];
        aux[2] = (double)(m0 | (one ? 16 : 0) | (ncur << 5));
    }
    CCD_FN void load_start(const double (&coef)[D + 1], const double *aux)
    {
        if (STRIDE)
        {
#pragma unroll
            for (int k = 0; k <= D; k++) mc[k * STRIDE] = coef[k];
#pragma unroll
            for (int k = 0; k < D; k++) { mcur[k * STRIDE] = 0.0; mout[k * STRIDE] = 0.0; }
            mcur[0] = aux[0];
            mcur[STRIDE] = aux[1];
        }
        else
        {
#pragma unroll
            for (int k = 0; k <= D; k++) c[k] = coef[k];
#pragma unroll
            for (int k = 0; k < D; k++) { cur[k] = 0.0; out[k] = 0.0; }
            cur[0] = aux[0];
            cur[1] = aux[1];
        }
        const int pk = (int)aux[2];
        m0 = pk & 15;
        one = (pk & 16) != 0;
        ncur = pk >> 5;
        start();
    }

    // until a solve is pending or the polynomial is finished
    CCD_FN void advance()
    {
        if (got_root)
        {
            // the solve that just converged (recorded here, outside the Newton loop, so that loop stays small)
            if (nr == 0 || prev != x) push_out(x);
            next_piece();
            got_root = false;
        }
        while (!solving && !done)
        {
            if (i + 1 < nb)
            {
                x_hi = (i + 1 == nb - 1) ? 1.0 : cur_at(i);
                f_hi = horner_padded<D>(p, x_hi);
                if (f_lo == 0.0)
                {
                    if (!plain && (i > 0 || last) && (nr == 0 || prev != brk_lo)) push_out(brk_lo);
                    next_piece();
                }
                else if ((f_lo < 0.0 && f_hi > 0.0) || (f_lo > 0.0 && f_hi < 0.0))
                {
                    lo = brk_lo;
                    hi = x_hi;
                    lo_neg = f_lo < 0.0;
                    x = 0.5 * (lo + hi);
                    dxold = hi - lo;
                    dx = dxold;
                    it = 0;
                    solving = true;
                }
                else
                    next_piece();
            }
            else
            {
                // level finished: f_lo is the value at the last breakpoint (t = 1)
                if (!plain && last && f_lo == 0.0 && (nr == 0 || prev != 1.0)) push_out(1.0);
                const bool keep_all = last && !plain;
                int n2 = 0;
                if (STRIDE)
                {
                    const int nlim = nr < D ? nr : D;
                    for (int k = 0; k < nlim; k++)
                    {
                        const double o = mout[k * STRIDE];
                        if (keep_all || (o > 0.0 && o < 1.0)) mcur[(n2++) * STRIDE] = o;
                    }
                }
                else
                {
#pragma unroll
                    for (int k = 0; k < D; k++)
                        if (k < nr && (keep_all || (out[k] > 0.0 && out[k] < 1.0)))
                        {
#pragma unroll
                            for (int j = 0; j < D; j++)
                                if (j == n2) cur[j] = out[k];
                            n2++;
                        }
                }
                ncur = n2;
                if (m == D) { done = true; return; }
                m++;
                setup_level();
            }
        }
    }

    // one iteration of solve_bracket (ccd_math.cuh; same operations in the same order; the exits are folded into one flag)
    CCD_FN void newton_step()
    {
        double f, df;
        horner2_padded<D>(p, x, f, df);
        bool conv = (f == 0.0);
        if (!conv)
        {
            if ((f < 0.0) == lo_neg)
                lo = x;
            else
                hi = x;
            const double step = f / df;
            double xn = x - step;
            const bool bisect = !(xn > lo && xn < hi) || fabs(2.0 * f) > fabs(dxold * df);
            dxold = dx;
            if (bisect)
            {
                dx = 0.5 * (hi - lo);
                xn = lo + dx;
                conv = !(xn > lo && xn < hi);
            }
            else
                dx = step;
            if (!conv) conv = fabs(xn - x) <= 8.9e-16 * fabs(xn);
            x = xn;
            it++;
            if (it == 128) conv = true;
        }
        if (conv)
        {
            solving = false;
            got_root = true;
        }
    }

    // roots in [0,1] (ascending) once done
    CCD_FN int result(double (&r)[6]) const
    {
#pragma unroll
        for (int k = 0; k < 6; k++) r[k] = (k < D) ? (STRIDE ? mcur[(k < D ? k : 0) * STRIDE] : cur[k < D ? k : 0]) : 0.0;
        return ncur;
    }
};

// host / reference use: the whole polynomial in one go
template <int D> CCD_FN int roots01_lane(const double (&c)[D + 1], double (&r)[6])
{
    RootLane<D> L;
    L.begin(c);
    while (!L.done)
    {
        L.advance();
        while (L.solving) L.newton_step();
    }
    return L.result(r);
}

} // namespace ccd
