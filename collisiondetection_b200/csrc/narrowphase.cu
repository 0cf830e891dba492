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
//   resume  the work list only: continues at the deferred sub-test with its roots; a later sub-test that needs
//           roots again is deferred once more (second, tiny round) so that only the last, rare-work kernel
//           carries the isolator inline — the big kernels stay small enough for the instruction cache.
// Degenerate sub-tests (and the EE primitive) are skipped when the swept boxes of the two parts stay further
// apart than eta plus a safety margin: those tests measure true distances (include/CTCD.h:31-79), so the
// reference cannot report a hit there; the margin (4e-5 of the coordinate scale) is ~1e3 times the distance
// at which its double-precision polynomials could change sign.  The VF primitive is never culled this way —
// its inside tests are not geometric distances (src/CTCD.cpp:433-464).
// Hit flag, time of impact and the index of the sub-test that fired are written per stencil; earliest TOI and
// hit counts are reduced per warp and merged with one atomic pair per warp.
#include "ccd_kernels.h"
#include "ccd_math.cuh"

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
// Sub-tests are numbered in the reference's order: 0 = the VF / EE primitive, then the vertex-edge tests, then
// the vertex-vertex tests.  a[0..3] start positions, b[0..3] end positions.
// Return: stage that hit (sub-test index + 1), 0 for a miss, -(sub+1) when sub-test `sub` was deferred (its pending
// polynomials are in P).  The walk starts at sub-test `start` (everything before it is known to miss); that sub-test
// runs in mode FIRST (RESUME: its roots come from trec), the ones after it in mode LATER.
template <int FIRST, int LATER>
static __device__ __noinline__ int vf_stencil_segment(const V3 *a, const V3 *b, double eta, double &t, int start, Pend &P, const double *trec)
{
    V3 v[4];
    Box bx[4];
    for (int i = 0; i < 4; i++) { v[i] = b[i] - a[i]; bx[i] = swept_box(a[i], b[i]); }
    int r;
    if (start <= 0)
    {
        r = vertex_face<FIRST>(a, v, eta, t, P, trec);
        if (r == R_HIT) return 1;
        if (r == R_DEFER) return -1;
    }
    const Box face = join(join(bx[1], bx[2]), bx[3]);
    const double m = eta + 4e-5 * fmax(box_scale(bx[0]), box_scale(face));
    if (apart(bx[0], face, m)) return 0;
    // vertex against the three face edges (1,2),(2,3),(3,1): src/CTCDNarrowPhase.cpp:51-59
    for (int e = 0; e < 3; e++)
    {
        const int sub = 1 + e, i1 = 1 + e, i2 = 1 + ((e + 1) % 3);
        if (sub < start) continue;
        if (apart(bx[0], join(bx[i1], bx[i2]), m)) continue;
        r = (FIRST != LATER && sub == start) ? vertex_edge<FIRST>(a[0], a[i1], a[i2], v[0], v[i1], v[i2], eta, t, P, trec)
                                              : vertex_edge<LATER>(a[0], a[i1], a[i2], v[0], v[i1], v[i2], eta, t, P, trec);
        if (r == R_HIT) return sub + 1;
        if (r == R_DEFER) return -(sub + 1);
    }
    // vertex against the three face vertices: src/CTCDNarrowPhase.cpp:61-69
    for (int k = 0; k < 3; k++)
    {
        if (apart(bx[0], bx[1 + k], m)) continue;
        if (vertex_vertex(a[0], a[1 + k], v[0], v[1 + k], eta, t) == R_HIT) return 5 + k;
    }
    return 0;
}

// a/b: (p0, p1, q0, q1); edgeEdgeCTCD takes (q0,p0,q1,p1) = (pos0,pos1,pos2,pos3): src/CTCDNarrowPhase.cpp:91
template <int FIRST, int LATER>
static __device__ __noinline__ int ee_stencil_segment(const V3 *a, const V3 *b, double eta, double &t, int start, Pend &P, const double *trec)
{
    V3 v[4];
    Box bx[4];
    for (int i = 0; i < 4; i++) { v[i] = b[i] - a[i]; bx[i] = swept_box(a[i], b[i]); }
    const Box e0 = join(bx[0], bx[1]), e1 = join(bx[2], bx[3]);
    const double m = eta + 4e-5 * fmax(box_scale(e0), box_scale(e1));
    if (apart(e0, e1, m)) return 0;       // every sub-test below is a distance between parts of these two edges
    int r;
    if (start <= 0)
    {
        r = edge_edge<FIRST>(a, v, eta, t, P, trec);
        if (r == R_HIT) return 1;
        if (r == R_DEFER) return -1;
    }
    // src/CTCDNarrowPhase.cpp:99-114: p0|p1 against (q0,q1), q0|q1 against (p0,p1)
    for (int k = 0; k < 4; k++)
    {
        const int sub = 1 + k, iv = k, i1 = (k < 2) ? 2 : 0, i2 = i1 + 1;
        if (sub < start) continue;
        if (apart(bx[iv], (k < 2) ? e1 : e0, m)) continue;
        r = (FIRST != LATER && sub == start) ? vertex_edge<FIRST>(a[iv], a[i1], a[i2], v[iv], v[i1], v[i2], eta, t, P, trec)
                                              : vertex_edge<LATER>(a[iv], a[i1], a[i2], v[iv], v[i1], v[i2], eta, t, P, trec);
        if (r == R_HIT) return sub + 1;
        if (r == R_DEFER) return -(sub + 1);
    }
    // src/CTCDNarrowPhase.cpp:117-132: (p0,q0) (p0,q1) (p1,q0) (p1,q1)
    for (int k = 0; k < 4; k++)
    {
        const int i1 = k >> 1, i2 = 2 + (k & 1);
        if (apart(bx[i1], bx[i2], m)) continue;
        if (vertex_vertex(a[i1], a[i2], v[i1], v[i2], eta, t) == R_HIT) return 6 + k;
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

// one list of deferred stencils: entry w = {stencil index, first task record, sub-test that deferred}
struct WorkList
{
    int *stencil;
    int *task;
    unsigned char *sub;
    unsigned long long *count;
};

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
    WorkList in, out;           // pass 1 fills `out`; a resume pass reads `in` and may fill `out`
    double *tasks;              // 64-byte records: coefficients + reduced degree in, roots + count out
    unsigned long long *ntask;  // running number of task records
    unsigned long long task_cap;
};

// one stencil, single linear segment (two History entries per vertex)
template <bool IS_VF, int FIRST, int LATER>
__device__ __forceinline__ int run_single(const NpArgs &A, long long i, double &toi, int start, Pend &P, const double *trec)
{
    const int4 s = reinterpret_cast<const int4 *>(A.stencils)[i];
    const double eta = A.eta_arr ? A.eta_arr[i] : A.eta_all;
    const long long vs = A.vstride;
    V3 a[4], b[4];
    a[0] = ldv(A.q0 + vs * s.x); a[1] = ldv(A.q0 + vs * s.y); a[2] = ldv(A.q0 + vs * s.z); a[3] = ldv(A.q0 + vs * s.w);
    b[0] = ldv(A.q1 + vs * s.x); b[1] = ldv(A.q1 + vs * s.y); b[2] = ldv(A.q1 + vs * s.z); b[3] = ldv(A.q1 + vs * s.w);
    return IS_VF ? vf_stencil_segment<FIRST, LATER>(a, b, eta, toi, start, P, trec) : ee_stencil_segment<FIRST, LATER>(a, b, eta, toi, start, P, trec);
}

__device__ __forceinline__ void store_result(const NpArgs &A, long long i, int stage, double toi)
{
    A.hit[i] = stage != 0;
    A.toi[i] = stage ? toi : 0.0;
    if (A.stage) A.stage[i] = (unsigned char)stage;
}

// Whole-warp call.  Lanes with `deferred` set put stencil i on the out list with one task record per pending polynomial
// of P (coefficients + reduced degree): one atomic for the list slots and one for the records per warp.
__device__ __forceinline__ void defer_stencil(const NpArgs &A, bool deferred, long long i, int sub, const Pend &P)
{
    const int lane = threadIdx.x & 31;
    const unsigned m = __ballot_sync(0xffffffffu, deferred);
    if (!m) return;
    const int ntask = deferred ? __popc(P.mask) : 0;
    int pre = ntask;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        int x = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += x;
    }
    const int total = __shfl_sync(0xffffffffu, pre, 31);
    unsigned long long wbase = 0, tbase = 0;
    if (lane == 0)
    {
        wbase = atomicAdd(A.out.count, (unsigned long long)__popc(m));
        tbase = atomicAdd(A.ntask, (unsigned long long)total);
    }
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    tbase = __shfl_sync(0xffffffffu, tbase, 0);
    if (!deferred) return;
    const unsigned long long w = wbase + __popc(m & ((1u << lane) - 1));
    const unsigned long long t0 = tbase + (unsigned long long)(pre - ntask);
    A.out.stencil[w] = (int)i;
    A.out.task[w] = (int)t0;
    A.out.sub[w] = (unsigned char)sub;
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
}

// pass 1: all stencils, straight-line work only.  A stencil whose next sub-test needs the root isolator is deferred.
template <bool IS_VF> __global__ void __launch_bounds__(128) stencil_pass1_kernel(NpArgs A)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int stage = 0;
    double toi = 0.0;
    Pend P;
    P.mask = 0;
    if (i < A.n)
    {
        stage = run_single<IS_VF, MODE_DEFER, MODE_DEFER>(A, i, toi, 0, P, nullptr);
        if (stage >= 0) store_result(A, i, stage, toi);
    }
    defer_stencil(A, stage < 0, i, -stage - 1, P);
    reduce_warp(stage > 0, toi, A.earliest_bits, A.nhit);
}

// root kernel: one thread per pending polynomial of records [*begin, *end); coefficients are replaced by the roots in [0,1]
__global__ void __launch_bounds__(128) roots_kernel(double *tasks, const unsigned long long *begin_ptr, const unsigned long long *end_ptr,
                                                    unsigned long long cap)
{
    unsigned long long nt = *end_ptr;
    const unsigned long long t0 = begin_ptr ? *begin_ptr : 0ull;
    if (nt > cap) nt = cap;
    for (unsigned long long j = t0 + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < nt;
         j += (unsigned long long)gridDim.x * blockDim.x)
    {
        double *rec = tasks + 8 * j;
        double op[7], b[7], roots[6];
        const int rd = (int)rec[7];
        for (int c = 0; c <= rd; c++) op[c] = rec[c];
        bernstein(op, rd, b);
        const int nr = roots01(op, rd, b, roots);
        for (int c = 0; c < nr; c++) rec[c] = roots[c];
        rec[7] = (double)nr;
    }
}

// resume pass: the stencils of the `in` list continue at the sub-test that deferred, with its roots read from the task
// records.  LATER = DEFER: a later sub-test that needs the isolator again goes on the `out` list (second round);
// LATER = FULL: everything is finished in place (last round; rare work, the only kernel that carries the isolator inline).
template <bool IS_VF, int LATER> __global__ void __launch_bounds__(128) stencil_resume_kernel(NpArgs A)
{
    const unsigned long long nw = *A.in.count;
    const unsigned long long nround = (nw + 31ull) & ~31ull;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < nround;
         w += (unsigned long long)gridDim.x * blockDim.x)
    {
        int stage = 0;
        long long i = 0;
        double toi = 0.0;
        Pend P;
        P.mask = 0;
        // records past the capacity were never written: the caller grows the buffer and reruns the whole narrowphase
        if (w < nw && (unsigned long long)A.in.task[w] + 5ull <= A.task_cap)
        {
            i = A.in.stencil[w];
            stage = run_single<IS_VF, MODE_RESUME, LATER>(A, i, toi, A.in.sub[w], P, A.tasks + 8ll * A.in.task[w]);
            if (stage >= 0) store_result(A, i, stage, toi);
        }
        if (LATER == MODE_DEFER) defer_stencil(A, stage < 0, i, -stage - 1, P);
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
        Pend P;
        Stitcher st;
        st.begin(A.hoff, A.htime, A.hpos, verts);
        if (st.next(a))
            while (st.next(b))
            {
                stage = IS_VF ? vf_stencil_segment<MODE_FULL, MODE_FULL>(a, b, eta, toi, 0, P, nullptr) : ee_stencil_segment<MODE_FULL, MODE_FULL>(a, b, eta, toi, 0, P, nullptr);
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

// Buffers: two work lists w1/w2 = {n ints, n ints, n bytes}, tasks (task_cap records of 8 doubles), counters
// ctr[0] = entries of list 1, ctr[1] = entries of list 2, ctr[2] = task records, ctr[3] = task records after pass 1
// (all zeroed here).  Returns the number of kernels launched.  If ctr[2] ends above task_cap the caller must grow the
// task buffer and call again.
int ccdk_narrowphase(cudaStream_t st, bool is_vf, long long n, const int *stencils, const double *eta_arr, double eta_all,
                     const double *q0, const double *q1, int vstride, const long long *hoff, const double *htime, const double *hpos,
                     unsigned char *hit, double *toi, unsigned char *stage, unsigned long long *earliest_bits,
                     unsigned long long *nhit, int *w1_stencil, int *w1_task, unsigned char *w1_sub, int *w2_stencil, int *w2_task,
                     unsigned char *w2_sub, double *tasks, unsigned long long task_cap, unsigned long long *ctr)
{
    if (n <= 0) return 0;
    NpArgs A;
    A.n = n; A.stencils = stencils; A.eta_arr = eta_arr; A.eta_all = eta_all; A.q0 = q0; A.q1 = q1; A.vstride = vstride;
    A.hoff = hoff; A.htime = htime; A.hpos = hpos; A.hit = hit; A.toi = toi; A.stage = stage;
    A.earliest_bits = earliest_bits; A.nhit = nhit;
    A.tasks = tasks; A.ntask = ctr + 2; A.task_cap = task_cap;
    const WorkList L1 = {w1_stencil, w1_task, w1_sub, ctr + 0}, L2 = {w2_stencil, w2_task, w2_sub, ctr + 1}, none = {nullptr, nullptr, nullptr, nullptr};
    const int B = 128;
    if (q0 == nullptr)
    {
        A.in = none; A.out = none;
        if (is_vf) stencil_history_kernel<true><<<grid_for(n, B), B, 0, st>>>(A);
        else stencil_history_kernel<false><<<grid_for(n, B), B, 0, st>>>(A);
        return 1;
    }
    cudaMemsetAsync(ctr, 0, 4 * sizeof(unsigned long long), st);
    const unsigned g2 = (unsigned)min((long long)148 * 32, (long long)grid_for(n, B));
    // pass 1 -> list 1
    A.in = none; A.out = L1;
    if (is_vf) stencil_pass1_kernel<true><<<grid_for(n, B), B, 0, st>>>(A);
    else stencil_pass1_kernel<false><<<grid_for(n, B), B, 0, st>>>(A);
    cudaMemcpyAsync(ctr + 3, ctr + 2, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st);
    roots_kernel<<<g2, B, 0, st>>>(tasks, nullptr, ctr + 3, task_cap);
    // resume list 1; sub-tests that need roots again -> list 2
    A.in = L1; A.out = L2;
    if (is_vf) stencil_resume_kernel<true, MODE_DEFER><<<g2, B, 0, st>>>(A);
    else stencil_resume_kernel<false, MODE_DEFER><<<g2, B, 0, st>>>(A);
    roots_kernel<<<148, B, 0, st>>>(tasks, ctr + 3, ctr + 2, task_cap);
    // resume list 2 and finish in place
    A.in = L2; A.out = none;
    if (is_vf) stencil_resume_kernel<true, MODE_FULL><<<148 * 4, B, 0, st>>>(A);
    else stencil_resume_kernel<false, MODE_FULL><<<148 * 4, B, 0, st>>>(A);
    return 5;
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
