// CTCD narrowphase and batched primitive kernels (sm_100a, FP64, no tensor cores: nothing in this
// path is a dense contraction).
//
// One thread per candidate stencil runs the reference's per-stencil sequence
// (src/CTCDNarrowPhase.cpp:24-135): the VF (or EE) primitive, then the degenerate vertex-edge and
// vertex-vertex tests, first hit wins; over a multi-entry History the sequence repeats per
// stitched linear segment (src/History.cpp:98-140).
//
// Three kernels keep the warps converged (profiles/: the one-pass version ran with 5.7 of 32 lanes active):
//   pass 1  every stencil, straight-line work only: coefficient construction, the reference's own quick
//           rejects, closed-form quadratics and the Bernstein "no root in [0,1]" decision.  A stencil whose
//           next sub-test needs the iterative root isolator goes on a work list, with one 64-byte task
//           record per pending polynomial (warp-aggregated allocation);
//   roots   one thread per pending polynomial: dense warps, every lane inside the isolator;
//           (the walk goes on past a deferring sub-test, so one round collects everything a stencil needs);
//   pass 3  the work list only: re-evaluates just the deferred sub-tests, in order, with their roots.
// Degenerate sub-tests (and the EE primitive) are skipped when the swept boxes of the two parts stay further
// apart than eta plus a safety margin: those tests measure true distances (include/CTCD.h:31-79), so the
// reference cannot report a hit there; the margin (4e-5 of the coordinate scale) is ~1e3 times the distance
// at which its double-precision polynomials could change sign.  The VF primitive is never culled this way —
// its inside tests are not geometric distances (src/CTCD.cpp:433-464).
// Hit flag, time of impact and the index of the sub-test that fired are written per stencil; earliest TOI and
// hit counts are reduced per warp and merged with one atomic pair per warp.
#include "ccd_kernels.h"
#ifndef NP_MINB
#define NP_MINB 4      // resident blocks of 128 threads per SM the stencil kernels are compiled for
#endif
#include "ccd_math.cuh"
#include "ccd_roots_t.cuh"
#include "ccd_classify.cuh"
#include "ccd_resume.cuh"
#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>
namespace cg = cooperative_groups;

namespace ccd {

struct Box { V3 lo, hi; };

__device__ __forceinline__ Box swept_box(V3 s, V3 e)
{
    Box b;
    b.lo = mk(fmin(s.x, e.x), fmin(s.y, e.y), fmin(s.z, e.z));
    b.hi = mk(fmax(s.x, e.x), fmax(s.y, e.y), fmax(s.z, e.z));
    return b;
}
__device__ __forceinline__ Box join(const Box &a, const Box &b)
{
    Box r;
    r.lo = mk(fmin(a.lo.x, b.lo.x), fmin(a.lo.y, b.lo.y), fmin(a.lo.z, b.lo.z));
    r.hi = mk(fmax(a.hi.x, b.hi.x), fmax(a.hi.y, b.hi.y), fmax(a.hi.z, b.hi.z));
    return r;
}
// true when the boxes stay more than m apart along some axis
__device__ __forceinline__ bool apart(const Box &a, const Box &b, double m)
{
    return a.hi.x + m < b.lo.x || b.hi.x + m < a.lo.x || a.hi.y + m < b.lo.y || b.hi.y + m < a.lo.y || a.hi.z + m < b.lo.z ||
           b.hi.z + m < a.lo.z;
}
__device__ __forceinline__ double box_scale(const Box &b)
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
    static __device__ __forceinline__ void ve(int sub, int &iv, int &i1, int &i2)
    {
        if (IS_VF) { iv = 0; i1 = sub; i2 = 1 + (sub % 3); }
        else { iv = sub - 1; i1 = (sub <= 2) ? 2 : 0; i2 = i1 + 1; }
    }
    static __device__ __forceinline__ void vv(int k, int &i1, int &i2)
    {
        if (IS_VF) { i1 = 0; i2 = 1 + k; }
        else { i1 = k >> 1; i2 = 2 + (k & 1); }
    }
};

// the primitive or one vertex-edge test (sub <= NVE)
template <bool IS_VF, int MODE>
__device__ __forceinline__ int eval_sub(int sub, const V3 *a, const V3 *v, double eta, double &t, Pend &P, const double *trec)
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
    __device__ __forceinline__ void init(const V3 *a, const V3 *b, double eta)
    {
        for (int i = 0; i < 4; i++) bx[i] = swept_box(a[i], b[i]);
        if (IS_VF) { g0 = bx[0]; g1 = join(join(bx[1], bx[2]), bx[3]); }
        else { g0 = join(bx[0], bx[1]); g1 = join(bx[2], bx[3]); }
        m = eta + 4e-5 * fmax(box_scale(g0), box_scale(g1));
    }
    __device__ __forceinline__ bool stencil_apart() const { return apart(g0, g1, m); }
    __device__ __forceinline__ bool ve_apart(int sub) const
    {
        int iv, i1, i2;
        Subs<IS_VF>::ve(sub, iv, i1, i2);
        return apart(bx[iv], join(bx[i1], bx[i2]), m);
    }
    __device__ __forceinline__ bool vv_apart(int k) const
    {
        int i1, i2;
        Subs<IS_VF>::vv(k, i1, i2);
        return apart(bx[i1], bx[i2], m);
    }
};

// Whole sequence in place (FULL mode): multi-entry History segments and the reference-order semantics in one walk.
// Returns the stage that hit, 0 for a miss.
template <bool IS_VF> static __device__ __noinline__ int stencil_segment_full(const V3 *a, const V3 *b, double eta, double &t)
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

// ---- History::stitchCommonHistory for four vertices over the CSR history (src/History.cpp:98-140)
struct Stitcher
{
    const long long *hoff;
    const double *htime;
    const double *hpos;
    long long it[4], end[4];
    double curtime;

    __device__ void begin(const long long *ho, const double *ht, const double *hp, const int *verts)
    {
        hoff = ho; htime = ht; hpos = hp;
        for (int i = 0; i < 4; i++) { it[i] = ho[verts[i]]; end[i] = ho[verts[i] + 1]; }
        curtime = 0;
    }
    __device__ bool next(V3 *pos)
    {
        if (!(curtime <= 1.0))
            return false;
        double newtime = INFINITY;
        for (int i = 0; i < 4; i++)
        {
            long long nx = it[i] + 1;
            while (nx != end[i] && htime[nx] <= curtime) { ++nx; ++it[i]; }
            V3 oldpos = ldv(hpos + 3 * it[i]);
            if (nx == end[i])
                pos[i] = oldpos;
            else
            {
                V3 newpos = ldv(hpos + 3 * nx);
                double dt = htime[nx] - htime[it[i]];
                double a = curtime - htime[it[i]];
                double b = htime[nx] - curtime;
                V3 r = b * oldpos + a * newpos;
                pos[i] = mk(r.x / dt, r.y / dt, r.z / dt);
                newtime = smin(newtime, htime[nx]);
            }
        }
        curtime = newtime;
        return true;
    }
};

// ---- reductions -----------------------------------------------------------------------------
// TOI >= 0, so the IEEE bit pattern orders like the value: min over unsigned 64-bit.  Warp-level only.
__device__ __forceinline__ void reduce_warp(bool hit, double toi, unsigned long long *earliest_bits, unsigned long long *nhit)
{
    unsigned long long bits = hit ? (unsigned long long)__double_as_longlong(toi) : 0xFFFFFFFFFFFFFFFFull;
    unsigned cnt = hit ? 1u : 0u;
    const unsigned mask = __activemask();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        unsigned long long ob = __shfl_xor_sync(mask, bits, o);
        unsigned oc = __shfl_xor_sync(mask, cnt, o);
        bits = ob < bits ? ob : bits;
        cnt += oc;
    }
    if ((threadIdx.x & 31) == 0 && cnt)
    {
        atomicMin(earliest_bits, bits);
        atomicAdd(nhit, (unsigned long long)cnt);
    }
}

struct NpArgs
{
    long long n;
    const int *stencils;
    const double *eta_arr;
    double eta_all;
    const double *q0, *q1;
    int vstride;
    const long long *hoff;
    const double *htime, *hpos;
    unsigned char *hit;
    double *toi;
    unsigned char *stage;
    unsigned long long *earliest_bits, *nhit;
    // deferred stencils: entry w = {stencil index, meta = deferred-sub-test mask | decided later hit << 8, first task
    // record of each deferred sub-test (<= 5: the primitive and the vertex-edge tests)}
    int *w_stencil;
    int *w_meta;
    int *w_base;                // 5 ints per entry
    unsigned long long *nwork;
    double *tasks;              // 64-byte records: coefficients + reduced degree in, roots + count out
    unsigned long long *ntask;  // running number of task records
    unsigned long long task_cap;
};

struct StencilIn
{
    V3 a[4], b[4];
    double eta;
};

template <bool IS_VF> __device__ __forceinline__ void load_single(const NpArgs &A, long long i, StencilIn &S)
{
    const int4 s = reinterpret_cast<const int4 *>(A.stencils)[i];
    S.eta = A.eta_arr ? A.eta_arr[i] : A.eta_all;
    const long long vs = A.vstride;
    S.a[0] = ldv(A.q0 + vs * s.x); S.a[1] = ldv(A.q0 + vs * s.y); S.a[2] = ldv(A.q0 + vs * s.z); S.a[3] = ldv(A.q0 + vs * s.w);
    S.b[0] = ldv(A.q1 + vs * s.x); S.b[1] = ldv(A.q1 + vs * s.y); S.b[2] = ldv(A.q1 + vs * s.z); S.b[3] = ldv(A.q1 + vs * s.w);
}

__device__ __forceinline__ void store_result(const NpArgs &A, long long i, int stage, double toi)
{
    A.hit[i] = stage != 0;
    A.toi[i] = stage ? toi : 0.0;
    if (A.stage) A.stage[i] = (unsigned char)stage;
}

// Reserve n consecutive task records; returns the index of the first.  Called from divergent code: the lanes that
// happen to be here together share one atomic (coalesced group).
__device__ __forceinline__ unsigned long long alloc_task_slots(const NpArgs &A, int n)
{
    cg::coalesced_group g = cg::coalesced_threads();
    const int pre = cg::exclusive_scan(g, n);
    unsigned long long base = 0;
    if (g.thread_rank() == g.size() - 1) base = atomicAdd(A.ntask, (unsigned long long)(pre + n));
    base = g.shfl(base, g.size() - 1);
    return base + (unsigned long long)pre;
}

// One task record per pending polynomial of P (coefficients + reduced degree); returns the first record's index.
__device__ __forceinline__ int alloc_tasks(const NpArgs &A, const Pend &P)
{
    const unsigned long long t0 = alloc_task_slots(A, __popc(P.mask));
    int j = 0;
    unsigned mask = P.mask;
    while (mask)
    {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        if (t0 + j < A.task_cap)
        {
            double *rec = A.tasks + 8 * (t0 + j);
            const int rd = P.rds[k];
            for (int c = 0; c <= rd; c++) rec[c] = P.ops[k][c];
            rec[7] = (double)rd;
        }
        j++;
    }
    return (int)t0;
}

// pass 1: every stencil, straight-line work only, as a chain of dense kernels connected by queues in HBM, so that the
// lanes of a warp do the same kind of work at the same time and every kernel is small:
//   cull     one thread per stencil: gather the 8 positions, swept-box culling -> the sub-tests that have to be
//            evaluated at all become items of three queues (primitive / vertex-edge / vertex-vertex);
//   prim, ve, vv   one thread per queue item: the item ends as miss, hit or deferred (its pending polynomials exported
//            as task records); 2 bits per sub-test are OR-ed into the stencil's status word;
//   decide   one thread per stencil: first hit in the reference's order wins if nothing before it is deferred; a stencil
//            with deferred sub-tests goes on the work list with the later hit (if any) pass 1 already knows.
// Evaluating every culling survivor instead of stopping at the first hit costs little (hits are ~10 %) and changes no
// result: the winner is still the first hit in order.
struct P1Args
{
    NpArgs A;
    unsigned *status;            // per stencil: 2 bits per sub-test (1 hit, 2 deferred)
    int *sbase;                  // per stencil: first task record of deferred sub-test 0..4
    int *qprim, *qve, *qvv;      // queues: stencil index | sub-test << 28
    unsigned long long *nq;      // queue lengths: prim, ve, vv
};

__device__ __forceinline__ void queue_push(bool want, int value, int *queue, unsigned long long *count)
{
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (want) queue[base + __popc(m & ((1u << lane) - 1))] = value;
}

template <bool IS_VF> __global__ void __launch_bounds__(256) np_cull_kernel(P1Args Q)
{
    const NpArgs &A = Q.A;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int NVE = Subs<IS_VF>::NVE, NVV = Subs<IS_VF>::NVV;
    unsigned todo = 0;
    if (i < A.n)
    {
        StencilIn S;
        load_single<IS_VF>(A, i, S);
        Cull<IS_VF> c;
        c.init(S.a, S.b, S.eta);
        const bool far = c.stencil_apart();
        if (IS_VF || !far) todo = 1u;
        if (!far)
        {
            for (int sub = 1; sub <= NVE; sub++)
                if (!c.ve_apart(sub)) todo |= 1u << sub;
            for (int k = 0; k < NVV; k++)
                if (!c.vv_apart(k)) todo |= 1u << (NVE + 1 + k);
        }
        Q.status[i] = 0u;
    }
    queue_push((todo & 1u) != 0, (int)i, Q.qprim, Q.nq + 0);
#pragma unroll
    for (int sub = 1; sub <= NVE; sub++)
        queue_push((todo >> sub) & 1u, (int)i | (sub << 28), Q.qve, Q.nq + 1);
#pragma unroll
    for (int k = 0; k < NVV; k++)
        queue_push((todo >> (NVE + 1 + k)) & 1u, (int)i | (k << 28), Q.qvv, Q.nq + 2);
}

template <bool IS_VF> __global__ void __launch_bounds__(128, NP_MINB) np_prim_kernel(P1Args Q)
{
    const NpArgs &A = Q.A;
    const unsigned long long n = Q.nq[0];
    for (unsigned long long it = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; it < n; it += (unsigned long long)gridDim.x * blockDim.x)
    {
        // register-resident classification of the primitive's polynomials (ccd_classify.cuh); only when every list is
        // already known and non-empty (rare) is the general routine run for the interval combination
        const long long i = Q.qprim[it];
        StencilIn S;
        load_single<IS_VF>(A, i, S);
        V3 v[4];
        for (int k = 0; k < 4; k++) v[k] = S.b[k] - S.a[k];
        unsigned pend;
        if (!classify_primitive<IS_VF>(S.a, v, S.eta, pend))
            continue;
        if (pend == 0)
        {
            double t = 0.0;
            Pend P;
            if (eval_sub<IS_VF, MODE_DEFER>(0, S.a, v, S.eta, t, P, nullptr) == R_HIT) atomicOr(&Q.status[i], 1u);
            continue;
        }
        const unsigned long long t0 = alloc_task_slots(A, __popc(pend));
        int j = 0;
        while (pend)
        {
            const int k = __ffs(pend) - 1;
            pend &= pend - 1;
            if (t0 + j < A.task_cap) export_poly<IS_VF>(k, S.a, v, S.eta, A.tasks + 8 * (t0 + j));
            j++;
        }
        Q.sbase[5 * i + 0] = (int)t0;
        atomicOr(&Q.status[i], 2u);
    }
}

template <bool IS_VF> __global__ void __launch_bounds__(128, NP_MINB) np_ve_kernel(P1Args Q)
{
    const NpArgs &A = Q.A;
    const unsigned long long n = Q.nq[1];
    for (unsigned long long it = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; it < n; it += (unsigned long long)gridDim.x * blockDim.x)
    {
        const int item = Q.qve[it];
        const long long i = item & 0x0fffffff;
        const int sub = (unsigned)item >> 28;
        int iv, i1, i2;
        Subs<IS_VF>::ve(sub, iv, i1, i2);
        const int4 s4 = reinterpret_cast<const int4 *>(A.stencils)[i];
        const int idx[4] = {s4.x, s4.y, s4.z, s4.w};
        const double eta = A.eta_arr ? A.eta_arr[i] : A.eta_all;
        const long long vs = A.vstride;
        const V3 a0 = ldv(A.q0 + vs * idx[iv]), a1 = ldv(A.q0 + vs * idx[i1]), a2 = ldv(A.q0 + vs * idx[i2]);
        const V3 v0 = ldv(A.q1 + vs * idx[iv]) - a0, v1 = ldv(A.q1 + vs * idx[i1]) - a1, v2 = ldv(A.q1 + vs * idx[i2]) - a2;
        double rec[8];
        bool want_rec;
        const int c = classify_ve(a0, a1, a2, v0, v1, v2, eta, rec, want_rec);
        if (c == VE_MISS)
            continue;
        if (c == VE_FULL)
        {
            double t = 0.0;
            Pend P;
            if (vertex_edge<MODE_DEFER>(a0, a1, a2, v0, v1, v2, eta, t, P, nullptr) == R_HIT) atomicOr(&Q.status[i], 1u << (2 * sub));
            continue;
        }
        const unsigned long long t0 = alloc_task_slots(A, 1);
        if (t0 < A.task_cap)
        {
            double *out = A.tasks + 8 * t0;
            const int rd = (int)rec[7];
            out[0] = rec[0]; out[1] = rec[1]; out[2] = rec[2]; out[3] = rec[3];
            if (rd == 4) out[4] = rec[4];
            out[7] = rec[7];
        }
        Q.sbase[5 * i + sub] = (int)t0;
        atomicOr(&Q.status[i], 2u << (2 * sub));
    }
}

template <bool IS_VF> __global__ void __launch_bounds__(256) np_vv_kernel(P1Args Q)
{
    const NpArgs &A = Q.A;
    const unsigned long long n = Q.nq[2];
    constexpr int NVE = Subs<IS_VF>::NVE;
    for (unsigned long long it = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; it < n; it += (unsigned long long)gridDim.x * blockDim.x)
    {
        const int item = Q.qvv[it];
        const long long i = item & 0x0fffffff;
        const int k = (unsigned)item >> 28;
        int i1, i2;
        Subs<IS_VF>::vv(k, i1, i2);
        const int4 s4 = reinterpret_cast<const int4 *>(A.stencils)[i];
        const int idx[4] = {s4.x, s4.y, s4.z, s4.w};
        const double eta = A.eta_arr ? A.eta_arr[i] : A.eta_all;
        const long long vs = A.vstride;
        const V3 a1 = ldv(A.q0 + vs * idx[i1]), a2 = ldv(A.q0 + vs * idx[i2]);
        const V3 v1 = ldv(A.q1 + vs * idx[i1]) - a1, v2 = ldv(A.q1 + vs * idx[i2]) - a2;
        double t = 0.0;
        if (vertex_vertex(a1, a2, v1, v2, eta, t) == R_HIT) atomicOr(&Q.status[i], 1u << (2 * (NVE + 1 + k)));
    }
}

// the t of a sub-test pass 1 decided as a hit (recomputed: only the first such hit of a stencil is ever needed)
template <bool IS_VF> static __device__ __noinline__ double decided_toi(const StencilIn &S, int sub)
{
    V3 v[4];
    for (int k = 0; k < 4; k++) v[k] = S.b[k] - S.a[k];
    double t = 0.0;
    if (sub <= Subs<IS_VF>::NVE)
    {
        Pend P;
        eval_sub<IS_VF, MODE_DEFER>(sub, S.a, v, S.eta, t, P, nullptr);
    }
    else
    {
        int i1, i2;
        Subs<IS_VF>::vv(sub - Subs<IS_VF>::NVE - 1, i1, i2);
        vertex_vertex(S.a[i1], S.a[i2], v[i1], v[i2], S.eta, t);
    }
    return t;
}

template <bool IS_VF> __global__ void __launch_bounds__(128, NP_MINB) np_decide_kernel(P1Args Q)
{
    const NpArgs &A = Q.A;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int NSUB = 1 + Subs<IS_VF>::NVE + Subs<IS_VF>::NVV;
    int stage = 0, meta = 0, nd = 0;
    int base[5] = {0, 0, 0, 0, 0};
    double toi = 0.0;
    if (i < A.n)
    {
        const unsigned st = Q.status[i];
        unsigned submask = 0;
        int later_hit = 255;
        if (st)
        {
            for (int sub = 0; sub < NSUB; sub++)
            {
                const unsigned r = (st >> (2 * sub)) & 3u;
                if (r == 1u)
                {
                    if (nd == 0) stage = sub + 1; else later_hit = sub;
                    break;
                }
                if (r == 2u) { base[nd++] = Q.sbase[5 * i + sub]; submask |= 1u << sub; }
            }
        }
        if (nd == 0)
        {
            if (stage)
            {
                StencilIn S;
                load_single<IS_VF>(A, i, S);
                toi = decided_toi<IS_VF>(S, stage - 1);
            }
            store_result(A, i, stage, toi);
        }
        else { stage = -1; meta = (int)(submask | ((unsigned)later_hit << 8)); }
    }
    // warp-aggregated append to the work list
    const bool deferred = stage < 0;
    const unsigned m = __ballot_sync(0xffffffffu, deferred);
    if (m)
    {
        const int lane = threadIdx.x & 31;
        unsigned long long wbase = 0;
        if (lane == 0) wbase = atomicAdd(A.nwork, (unsigned long long)__popc(m));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (deferred)
        {
            const unsigned long long w = wbase + __popc(m & ((1u << lane) - 1));
            A.w_stencil[w] = (int)i;
            A.w_meta[w] = meta;
            for (int j = 0; j < 5; j++) A.w_base[5 * w + j] = base[j];
        }
    }
    reduce_warp(stage > 0, toi, A.earliest_bits, A.nhit);
}

// Root stage.  Task records are bucketed by reduced degree (3..6) so that each root kernel is specialised for one degree:
// compile-time array sizes, everything in registers (ccd_roots_t.cuh), dense warps of identical work.
__global__ void __launch_bounds__(256) bucket_tasks_kernel(const double *__restrict__ tasks, const unsigned long long *ntask_ptr,
                                                           unsigned long long cap, int *__restrict__ lists, unsigned long long *counts)
{
    unsigned long long nt = *ntask_ptr;
    if (nt > cap) nt = cap;
    const unsigned long long nround = (nt + 31ull) & ~31ull;
    for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < nround;
         j += (unsigned long long)gridDim.x * blockDim.x)
    {
        const int rd = (j < nt) ? (int)tasks[8 * j + 7] : 0;
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int d = 3; d <= 6; d++)
        {
            const unsigned m = __ballot_sync(0xffffffffu, rd == d);
            if (m)
            {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(&counts[d - 3], (unsigned long long)__popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (rd == d) lists[(size_t)(d - 3) * cap + base + __popc(m & ((1u << lane) - 1))] = (int)j;
            }
        }
    }
}

// one thread per pending polynomial of degree D; the record's coefficients are replaced by its roots in [0,1]
template <int D>
__global__ void __launch_bounds__(128) roots_kernel(double *tasks, const int *__restrict__ list, const unsigned long long *count_ptr)
{
    const unsigned long long nt = *count_ptr;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < nt;
         w += (unsigned long long)gridDim.x * blockDim.x)
    {
        double *rec = tasks + 8ll * list[w];
        double op[D + 1], roots[6];
#pragma unroll
        for (int c = 0; c <= D; c++) op[c] = rec[c];
        const int nr = roots01_t<D>(op, roots);
#pragma unroll
        for (int c = 0; c < 6; c++)
            if (c < nr) rec[c] = roots[c];
        rec[7] = (double)nr;
    }
}

// pass 3: the deferred stencils.  Only the sub-tests that exported polynomials are evaluated again, in order, with
// their roots read from the task records; the first one that hits wins, else the hit pass 1 found later (if any).
// GENERAL = false: register-resident finish (ccd_resume.cuh); returns -1 when a sub-test needs the general routine —
// the stencil is then redone by the GENERAL = true instance in a separate, rarely needed kernel, so that the hot
// kernel does not carry the general code.
template <bool IS_VF, bool GENERAL>
static __device__ __noinline__ int resume_walk(const NpArgs &A, const StencilIn &S, double &t, int meta, const int *base)
{
    V3 v[4];
    for (int i = 0; i < 4; i++) v[i] = S.b[i] - S.a[i];
    unsigned submask = (unsigned)meta & 0xffu;
    const int later_hit = (meta >> 8) & 0xff;
    int j = 0;
    while (submask)
    {
        const int sub = __ffs(submask) - 1;
        submask &= submask - 1;
        const double *rec = A.tasks + 8ll * base[j];
        int r;
        if (GENERAL)
        {
            Pend P;
            r = (eval_sub<IS_VF, MODE_RESUME>(sub, S.a, v, S.eta, t, P, rec) == R_HIT) ? RS_HIT : RS_MISS;
        }
        else if (sub == 0)
            r = resume_primitive<IS_VF>(S.a, v, S.eta, rec, t);
        else
        {
            int iv, i1, i2;
            Subs<IS_VF>::ve(sub, iv, i1, i2);
            r = resume_ve(S.a[iv], S.a[i1], S.a[i2], v[iv], v[i1], v[i2], S.eta, rec, t);
        }
        if (r == RS_FALLBACK) return -1;
        if (r == RS_HIT) return sub + 1;
        j++;
    }
    if (later_hit == 255) return 0;
    if (!GENERAL) return -1;        // a decided later hit: its t is recomputed by the general routine (rare)
    t = decided_toi<IS_VF>(S, later_hit);
    return later_hit + 1;
}

// work: list of work-list entries to process (nullptr = all nwork entries); fb_list/fb_count: entries to redo generally
template <bool IS_VF, bool GENERAL>
__global__ void __launch_bounds__(128, NP_MINB) stencil_resume_kernel(NpArgs A, const int *work, const unsigned long long *work_count,
                                                                      int *fb_list, unsigned long long *fb_count)
{
    const unsigned long long nw = *work_count;
    const unsigned long long nround = (nw + 31ull) & ~31ull;
    for (unsigned long long x = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; x < nround;
         x += (unsigned long long)gridDim.x * blockDim.x)
    {
        int stage = 0;
        double toi = 0.0;
        bool fb = false;
        unsigned long long w = 0;
        if (x < nw)
        {
            w = work ? (unsigned long long)work[x] : x;
            int base[5];
            bool ok = true;      // records past the capacity were never written: the caller grows the buffer and reruns
            for (int j = 0; j < 5; j++) { base[j] = A.w_base[5 * w + j]; ok = ok && ((unsigned long long)base[j] + 5ull <= A.task_cap); }
            if (ok)
            {
                const long long i = A.w_stencil[w];
                StencilIn S;
                load_single<IS_VF>(A, i, S);
                stage = resume_walk<IS_VF, GENERAL>(A, S, toi, A.w_meta[w], base);
                if (stage >= 0) store_result(A, i, stage, toi);
                else { fb = true; stage = 0; }
            }
        }
        if (!GENERAL) queue_push(fb, (int)w, fb_list, fb_count);
        reduce_warp(stage > 0, toi, A.earliest_bits, A.nhit);
    }
}

// multi-entry History: stitched segments, full algorithm in one pass
template <bool IS_VF> __global__ void __launch_bounds__(128) stencil_history_kernel(NpArgs A)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int stage = 0;
    double toi = 0.0;
    if (i < A.n)
    {
        const int4 s = reinterpret_cast<const int4 *>(A.stencils)[i];
        const double eta = A.eta_arr ? A.eta_arr[i] : A.eta_all;
        int verts[4] = {s.x, s.y, s.z, s.w};
        V3 a[4], b[4];
        Stitcher st;
        st.begin(A.hoff, A.htime, A.hpos, verts);
        if (st.next(a))
            while (st.next(b))
            {
                stage = stencil_segment_full<IS_VF>(a, b, eta, toi);
                if (stage) break;
                for (int k = 0; k < 4; k++) a[k] = b[k];
            }
        store_result(A, i, stage, toi);
    }
    reduce_warp(stage != 0, toi, A.earliest_bits, A.nhit);
}

// ---- batched public primitives (include/CTCD.h:36-79): pts = start points then end points ------
__global__ void __launch_bounds__(128) prim_kernel(int kind, long long n, const double *__restrict__ pts, const double *__restrict__ eta,
                                                   unsigned char *__restrict__ hit, double *__restrict__ t)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double tt = 0;
    int r;
    const int np = (kind <= 1) ? 4 : (kind == 2 ? 3 : 2);
    const double *p = pts + 6ll * np * i;
    V3 s[4], v[4];
    for (int k = 0; k < np; k++) { s[k] = ldv(p + 3 * k); v[k] = ldv(p + 3 * (np + k)) - s[k]; }
    Pend P;
    if (kind == 0) r = vertex_face<MODE_FULL>(s, v, eta[i], tt, P, nullptr);
    else if (kind == 1) r = edge_edge<MODE_FULL>(s, v, eta[i], tt, P, nullptr);
    else if (kind == 2) r = vertex_edge<MODE_FULL>(s[0], s[1], s[2], v[0], v[1], v[2], eta[i], tt, P, nullptr);
    else r = vertex_vertex(s[0], s[1], v[0], v[1], eta[i], tt);
    hit[i] = r == R_HIT;
    if (r == R_HIT) t[i] = tt;       // t is written only on a hit, like the reference
}

// raw interval finder, exposed for parity tests
__global__ void find_intervals_kernel(long long n, int degree, int pos, const double *__restrict__ coeffs, int *__restrict__ cnt,
                                      double *__restrict__ lo, double *__restrict__ hi)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ivals iv;
    iv.n = 0;
    double op[7];
    for (int k = 0; k <= degree; k++) op[k] = coeffs[7 * i + k];
    find_intervals(op, degree, iv, pos != 0);
    cnt[i] = iv.n;
    for (int k = 0; k < iv.n; k++) { lo[7 * i + k] = iv.l[k]; hi[7 * i + k] = iv.u[k]; }
}

} // namespace ccd

// ---- launchers --------------------------------------------------------------------------------
using namespace ccd;

static inline unsigned grid_for(long long n, int block) { return (unsigned)((n + block - 1) / block); }

// Buffers: work list {w_stencil: n ints, w_meta: n ints, w_base: 5n ints}, tasks (task_cap records of 8 doubles),
// tlists (4 x task_cap ints: task indices by degree), pass-1 scratch {status: n u32, sbase: 5n ints, queues: 9n ints},
// counters ctr[0] = work-list entries, ctr[1] = task records, ctr[2..5] = tasks of degree 3..6, ctr[6..8] = queue
// lengths, ctr[9] = stencils for the general resume kernel (all zeroed here).  Returns the number of kernels launched.
// If ctr[1] ends above task_cap - 5 the caller must grow the task buffer and call again.
int ccdk_narrowphase(cudaStream_t st, bool is_vf, long long n, const int *stencils, const double *eta_arr, double eta_all,
                     const double *q0, const double *q1, int vstride, const long long *hoff, const double *htime, const double *hpos,
                     unsigned char *hit, double *toi, unsigned char *stage, unsigned long long *earliest_bits,
                     unsigned long long *nhit, int *w_stencil, int *w_meta, int *w_base, double *tasks, int *tlists,
                     unsigned long long task_cap, unsigned *status, int *sbase, int *queues, unsigned long long *ctr)
{
    if (n <= 0) return 0;
    NpArgs A;
    A.n = n; A.stencils = stencils; A.eta_arr = eta_arr; A.eta_all = eta_all; A.q0 = q0; A.q1 = q1; A.vstride = vstride;
    A.hoff = hoff; A.htime = htime; A.hpos = hpos; A.hit = hit; A.toi = toi; A.stage = stage;
    A.earliest_bits = earliest_bits; A.nhit = nhit;
    A.w_stencil = w_stencil; A.w_meta = w_meta; A.w_base = w_base; A.nwork = ctr + 0;
    A.tasks = tasks; A.ntask = ctr + 1; A.task_cap = task_cap;
    const int B = 128;
    if (q0 == nullptr)
    {
        if (is_vf) stencil_history_kernel<true><<<grid_for(n, B), B, 0, st>>>(A);
        else stencil_history_kernel<false><<<grid_for(n, B), B, 0, st>>>(A);
        return 1;
    }
    cudaMemsetAsync(ctr, 0, 10 * sizeof(unsigned long long), st);
    const unsigned g2 = (unsigned)min((long long)148 * 32, (long long)grid_for(n, B));
    P1Args Q;
    Q.A = A; Q.status = status; Q.sbase = sbase;
    Q.qprim = queues; Q.qve = queues + n; Q.qvv = queues + 5 * n; Q.nq = ctr + 6;
    const unsigned gq = 148 * 16;
    if (is_vf)
    {
        np_cull_kernel<true><<<grid_for(n, 256), 256, 0, st>>>(Q);
        np_prim_kernel<true><<<gq, B, 0, st>>>(Q);
        np_ve_kernel<true><<<gq, B, 0, st>>>(Q);
        np_vv_kernel<true><<<gq, 256, 0, st>>>(Q);
        np_decide_kernel<true><<<grid_for(n, B), B, 0, st>>>(Q);
    }
    else
    {
        np_cull_kernel<false><<<grid_for(n, 256), 256, 0, st>>>(Q);
        np_prim_kernel<false><<<gq, B, 0, st>>>(Q);
        np_ve_kernel<false><<<gq, B, 0, st>>>(Q);
        np_vv_kernel<false><<<gq, 256, 0, st>>>(Q);
        np_decide_kernel<false><<<grid_for(n, B), B, 0, st>>>(Q);
    }
    bucket_tasks_kernel<<<148 * 4, 256, 0, st>>>(tasks, ctr + 1, task_cap, tlists, ctr + 2);
    roots_kernel<3><<<g2, B, 0, st>>>(tasks, tlists + 0 * task_cap, ctr + 2);
    roots_kernel<4><<<g2, B, 0, st>>>(tasks, tlists + 1 * task_cap, ctr + 3);
    roots_kernel<5><<<g2, B, 0, st>>>(tasks, tlists + 2 * task_cap, ctr + 4);
    roots_kernel<6><<<g2, B, 0, st>>>(tasks, tlists + 3 * task_cap, ctr + 5);
    // the queues of pass 1 are free again: their buffer holds the (short) list of stencils for the general routine
    if (is_vf)
    {
        stencil_resume_kernel<true, false><<<g2, B, 0, st>>>(A, nullptr, ctr + 0, queues, ctr + 9);
        stencil_resume_kernel<true, true><<<148 * 2, B, 0, st>>>(A, queues, ctr + 9, nullptr, nullptr);
    }
    else
    {
        stencil_resume_kernel<false, false><<<g2, B, 0, st>>>(A, nullptr, ctr + 0, queues, ctr + 9);
        stencil_resume_kernel<false, true><<<148 * 2, B, 0, st>>>(A, queues, ctr + 9, nullptr, nullptr);
    }
    return 12;
}

void ccdk_prim_batch(cudaStream_t st, int kind, long long n, const double *pts, const double *eta, unsigned char *hit, double *t)
{
    if (n <= 0) return;
    prim_kernel<<<grid_for(n, 128), 128, 0, st>>>(kind, n, pts, eta, hit, t);
}

void ccdk_find_intervals(cudaStream_t st, long long n, int degree, int pos, const double *coeffs, int *cnt, double *lo, double *hi)
{
    if (n <= 0) return;
    find_intervals_kernel<<<grid_for(n, 128), 128, 0, st>>>(n, degree, pos, coeffs, cnt, lo, hi);
}
