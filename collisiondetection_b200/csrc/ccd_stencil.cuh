// Per-stencil structure of the CTCD narrowphase (src/CTCDNarrowPhase.cpp:24-135): numbering of the sub-tests, the
// swept-box culling of the degenerate sub-tests, and the whole reference-order sequence for one linear segment
// (stencil_segment_full: used for multi-entry Histories and as the general routine of the single-step pipeline).
// Host+device: tests/np_emul.cu runs the same code on the CPU.
#pragma once
#include "ccd_math.cuh"

namespace ccd {

struct Box { V3 lo, hi; };

CCD_FN Box swept_box(V3 s, V3 e)
{
    Box b;
    b.lo = mk(fmin(s.x, e.x), fmin(s.y, e.y), fmin(s.z, e.z));
    b.hi = mk(fmax(s.x, e.x), fmax(s.y, e.y), fmax(s.z, e.z));
    return b;
}
CCD_FN Box join(const Box &a, const Box &b)
{
    Box r;
    r.lo = mk(fmin(a.lo.x, b.lo.x), fmin(a.lo.y, b.lo.y), fmin(a.lo.z, b.lo.z));
    r.hi = mk(fmax(a.hi.x, b.hi.x), fmax(a.hi.y, b.hi.y), fmax(a.hi.z, b.hi.z));
    return r;
}
// true when the boxes stay more than m apart along some axis
CCD_FN bool apart(const Box &a, const Box &b, double m)
{
    return a.hi.x + m < b.lo.x || b.hi.x + m < a.lo.x || a.hi.y + m < b.lo.y || b.hi.y + m < a.lo.y || a.hi.z + m < b.lo.z ||
           b.hi.z + m < a.lo.z;
}
CCD_FN double box_scale(const Box &b)
{
    return fmax(fmax(fmax(fabs(b.lo.x), fabs(b.hi.x)), fmax(fabs(b.lo.y), fabs(b.hi.y))), fmax(fabs(b.lo.z), fabs(b.hi.z)));
}

// ---- per-segment stencil tests --------------------------------------------------------------
// Sub-tests are numbered in the reference's order: 0 = the VF / EE primitive, 1.. = the vertex-edge tests (3 for a VF
// stencil: face edges (1,2),(2,3),(3,1), src/CTCDNarrowPhase.cpp:51-59; 4 for an EE stencil: p0|p1 against (q0,q1),
// q0|q1 against (p0,p1), :99-114), then the vertex-vertex tests (:61-69, :117-132).  stage = sub-test index + 1.
// a[0..3] start positions of (p,q0,q1,q2) / (p0,p1,q0,q1), v[] = end - start.
template <bool IS_VF> struct Subs
{
    static constexpr int NVE = IS_VF ? 3 : 4;
    static constexpr int NVV = IS_VF ? 3 : 4;
    // vertex / edge endpoints of vertex-edge sub-test `sub` (1-based)
    static CCD_FN void ve(int sub, int &iv, int &i1, int &i2)
    {
        if (IS_VF) { iv = 0; i1 = sub; i2 = 1 + (sub % 3); }
        else { iv = sub - 1; i1 = (sub <= 2) ? 2 : 0; i2 = i1 + 1; }
    }
    static CCD_FN void vv(int k, int &i1, int &i2)
    {
        if (IS_VF) { i1 = 0; i2 = 1 + k; }
        else { i1 = k >> 1; i2 = 2 + (k & 1); }
    }
};

// the primitive or one vertex-edge test (sub <= NVE)
template <bool IS_VF, int MODE>
CCD_FN int eval_sub(int sub, const V3 *a, const V3 *v, double eta, double &t, Pend &P, const double *trec)
{
    if (sub == 0)
        return IS_VF ? vertex_face<MODE>(a, v, eta, t, P, trec) : edge_edge<MODE>(a, v, eta, t, P, trec);
    int iv, i1, i2;
    Subs<IS_VF>::ve(sub, iv, i1, i2);
    return vertex_edge<MODE>(a[iv], a[i1], a[i2], v[iv], v[i1], v[i2], eta, t, P, trec);
}

// swept boxes of the four vertices and the culling margin of the stencil
template <bool IS_VF> struct Cull
{
    Box bx[4], g0, g1;      // VF: g0 = vertex, g1 = face ; EE: g0 = edge (0,1), g1 = edge (2,3)
    double m;
    CCD_FN void init(const V3 *a, const V3 *b, double eta)
    {
        for (int i = 0; i < 4; i++) bx[i] = swept_box(a[i], b[i]);
        if (IS_VF) { g0 = bx[0]; g1 = join(join(bx[1], bx[2]), bx[3]); }
        else { g0 = join(bx[0], bx[1]); g1 = join(bx[2], bx[3]); }
        m = eta + 4e-5 * fmax(box_scale(g0), box_scale(g1));
    }
    CCD_FN bool stencil_apart() const { return apart(g0, g1, m); }
    CCD_FN bool ve_apart(int sub) const
    {
        int iv, i1, i2;
        Subs<IS_VF>::ve(sub, iv, i1, i2);
        return apart(bx[iv], join(bx[i1], bx[i2]), m);
    }
    CCD_FN bool vv_apart(int k) const
    {
        int i1, i2;
        Subs<IS_VF>::vv(k, i1, i2);
        return apart(bx[i1], bx[i2], m);
    }
};

// ---- the same culling on float swept boxes rounded OUTWARD (one 32-byte record per vertex, built once per step) -----
// Conservative with respect to the double-precision rule above: a float box contains the exact one and sums are rounded
// up, so "apart" here implies apart there.  It culls marginally less and never more; results cannot change, and the
// dense pass over all stencils reads 128 bytes per stencil instead of 512 and touches no FP64 unit.
struct BoxF { float lo[3], hi[3]; };

CCD_FN float f_down(double x)
{
#ifdef __CUDA_ARCH__
    return __double2float_rd(x);
#else
    float f = (float)x;
    return ((double)f > x) ? nextafterf(f, -INFINITY) : f;
#endif
}
CCD_FN float f_up(double x)
{
#ifdef __CUDA_ARCH__
    return __double2float_ru(x);
#else
    float f = (float)x;
    return ((double)f < x) ? nextafterf(f, INFINITY) : f;
#endif
}
CCD_FN float f_add_up(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fadd_ru(a, b);
#else
    return f_up((double)a + (double)b);
#endif
}
CCD_FN float f_mul_up(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fmul_ru(a, b);
#else
    return f_up((double)a * (double)b);
#endif
}
CCD_FN BoxF swept_box_f(V3 s, V3 e)
{
    BoxF b;
    b.lo[0] = f_down(fmin(s.x, e.x)); b.lo[1] = f_down(fmin(s.y, e.y)); b.lo[2] = f_down(fmin(s.z, e.z));
    b.hi[0] = f_up(fmax(s.x, e.x)); b.hi[1] = f_up(fmax(s.y, e.y)); b.hi[2] = f_up(fmax(s.z, e.z));
    return b;
}
CCD_FN BoxF join_f(const BoxF &a, const BoxF &b)
{
    BoxF r;
#pragma unroll
    for (int k = 0; k < 3; k++) { r.lo[k] = fminf(a.lo[k], b.lo[k]); r.hi[k] = fmaxf(a.hi[k], b.hi[k]); }
    return r;
}
CCD_FN bool apart_f(const BoxF &a, const BoxF &b, float m)
{
    bool r = false;
#pragma unroll
    for (int k = 0; k < 3; k++) r = r || f_add_up(a.hi[k], m) < b.lo[k] || f_add_up(b.hi[k], m) < a.lo[k];
    return r;
}
CCD_FN float box_scale_f(const BoxF &b)
{
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) r = fmaxf(r, fmaxf(fabsf(b.lo[k]), fabsf(b.hi[k])));
    return r;
}

// Extent of a group of vertices along the ten diagonal k-DOP directions (x+-y, x+-z, y+-z, x+-y+-z; NOT normalised: lengths
// sqrt 2 and sqrt 3), in float, bounded from the vertices' swept BOXES: a vertex moves on a segment inside its box, so its
// projection on x+y stays within [lo.x + lo.y, hi.x + hi.y], and so on: no position is read.  Two groups separated along a
// direction by more than 2 m (> m times the direction's length) stay further than m apart: the same argument as for the
// coordinate axes, and with m >= 4e-5 of the coordinate scale the float roundings of the sums (~1e-7 of the scale) are two
// orders below the margin.  At the 4M-triangle cloth half of the edge pairs whose swept boxes overlap are separated along one
// of these directions (bounding the projections by the box corners instead of the two end points costs 1 % of that).
struct DopF { float lo[10], hi[10]; };
CCD_FN void dopf_init(DopF &d)
{
#pragma unroll
    for (int k = 0; k < 10; k++) { d.lo[k] = INFINITY; d.hi[k] = -INFINITY; }
}
CCD_FN void dopf_add_box(DopF &d, const BoxF &b)
{
    const float lx = b.lo[0], ly = b.lo[1], lz = b.lo[2], hx = b.hi[0], hy = b.hi[1], hz = b.hi[2];
    const float lxy = lx + ly, lxmy = lx - hy, hxy = hx + hy, hxmy = hx - ly;
    const float lo[10] = {lxy, lxmy, lx + lz, lx - hz, ly + lz, ly - hz, lxy + lz, lxy - hz, lxmy + lz, lxmy - hz};
    const float hi[10] = {hxy, hxmy, hx + hz, hx - lz, hy + hz, hy - lz, hxy + hz, hxy - lz, hxmy + hz, hxmy - lz};
#pragma unroll
    for (int k = 0; k < 10; k++) { d.lo[k] = fminf(d.lo[k], lo[k]); d.hi[k] = fmaxf(d.hi[k], hi[k]); }
}
CCD_FN bool dopf_apart(const DopF &a, const DopF &b, float m2)
{
    bool r = false;
#pragma unroll
    for (int k = 0; k < 10; k++) r = r || (b.lo[k] - a.hi[k] > m2) || (a.lo[k] - b.hi[k] > m2);
    return r;
}

template <bool IS_VF> struct CullF
{
    BoxF bx[4], g0, g1;
    float m;
    bool far13 = false;      // the two parts are separated along a diagonal direction (init13)
    CCD_FN void init(double eta)
    {
        if (IS_VF) { g0 = bx[0]; g1 = join_f(join_f(bx[1], bx[2]), bx[3]); }
        else { g0 = join_f(bx[0], bx[1]); g1 = join_f(bx[2], bx[3]); }
        m = f_add_up(f_up(eta), f_mul_up(4.0001e-5f, fmaxf(box_scale_f(g0), box_scale_f(g1))));
    }
    CCD_FN bool stencil_apart() const { return apart_f(g0, g1, m); }
    // call after init(); only worth it when !stencil_apart()
    CCD_FN void init13()
    {
        DopF d0, d1;
        dopf_init(d0);
        dopf_init(d1);
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            const bool first = IS_VF ? (i == 0) : (i < 2);
            if (first) dopf_add_box(d0, bx[i]); else dopf_add_box(d1, bx[i]);
        }
        far13 = dopf_apart(d0, d1, f_mul_up(2.0f, m));
    }
    CCD_FN bool ve_apart(int sub) const
    {
        int iv, i1, i2;
        Subs<IS_VF>::ve(sub, iv, i1, i2);
        return apart_f(bx[iv], join_f(bx[i1], bx[i2]), m);
    }
    CCD_FN bool vv_apart(int k) const
    {
        int i1, i2;
        Subs<IS_VF>::vv(k, i1, i2);
        return apart_f(bx[i1], bx[i2], m);
    }
    // sub-tests that have to be evaluated: bit 0 the primitive, then the vertex-edge and vertex-vertex tests
    CCD_FN unsigned todo() const
    {
        constexpr int NVE = Subs<IS_VF>::NVE, NVV = Subs<IS_VF>::NVV;
        const bool far = stencil_apart() || far13;
        unsigned t = (IS_VF || !far) ? 1u : 0u;
        if (!far)
        {
            for (int sub = 1; sub <= NVE; sub++)
                if (!ve_apart(sub)) t |= 1u << sub;
            for (int k = 0; k < NVV; k++)
                if (!vv_apart(k)) t |= 1u << (NVE + 1 + k);
        }
        return t;
    }
};

// Whole sequence in place (FULL mode): multi-entry History segments and the reference-order semantics in one walk.
// Returns the stage that hit, 0 for a miss.
template <bool IS_VF> static CCD_HD __noinline__ int stencil_segment_full(const V3 *a, const V3 *b, double eta, double &t)
{
    V3 v[4];
    for (int i = 0; i < 4; i++) v[i] = b[i] - a[i];
    Cull<IS_VF> c;
    c.init(a, b, eta);
    Pend P;
    if (!IS_VF && c.stencil_apart()) return 0;      // every EE sub-test is a distance between parts of the two edges
    if (eval_sub<IS_VF, MODE_FULL>(0, a, v, eta, t, P, nullptr) == R_HIT) return 1;
    if (IS_VF && c.stencil_apart()) return 0;
    for (int sub = 1; sub <= Subs<IS_VF>::NVE; sub++)
    {
        if (c.ve_apart(sub)) continue;
        if (eval_sub<IS_VF, MODE_FULL>(sub, a, v, eta, t, P, nullptr) == R_HIT) return sub + 1;
    }
    for (int k = 0; k < Subs<IS_VF>::NVV; k++)
    {
        if (c.vv_apart(k)) continue;
        int i1, i2;
        Subs<IS_VF>::vv(k, i1, i2);
        if (vertex_vertex(a[i1], a[i2], v[i1], v[i2], eta, t) == R_HIT) return Subs<IS_VF>::NVE + 2 + k;
    }
    return 0;
}

} // namespace ccd
