// CTCD narrowphase and batched primitive kernels (sm_100a, FP64, no tensor cores: nothing in this
// path is a dense contraction).
//
// One thread per candidate stencil runs the reference's per-stencil sequence
// (src/CTCDNarrowPhase.cpp:24-135): the VF (or EE) primitive, then the degenerate vertex-edge and
// vertex-vertex tests, first hit wins; over a multi-entry History the sequence repeats per
// stitched linear segment (src/History.cpp:98-140).
//
// The single-step path is a chain of dense kernels connected by queues (see "Single-step pipeline" below): the
// one-thread-per-stencil walk ran with 5.7 of 32 lanes active, and even the first staged version (one kernel per
// sub-test kind) stayed at 7-12 lanes because stencils leave at different polynomials and every root solve was an
// inlined copy (profiles/).  Now every kernel handles ONE polynomial (or one kind of step) for all its lanes.
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
#ifndef NP_MINB_STAGE
#define NP_MINB_STAGE 8      // ... and the per-polynomial stage / export kernels (latency-bound gathers: occupancy wins, measured)
#endif
#include "ccd_math.cuh"
#include "ccd_roots_t.cuh"
#include "ccd_stencil.cuh"
#include "ccd_classify.cuh"
#include "ccd_solve.cuh"
#include "ccd_stages.cuh"
#include "ccd_sepplane.cuh"
#include <stdio.h>
#include <stdlib.h>

namespace ccd {

// ---- History::stitchCommonHistory for four vertices over the CSR history (src/History.cpp:98-140)
struct Stitcher
{
    const long long *hoff;
    const double *htime;
    const double *hpos;
    long long it[4], end[4];
    double curtime;
    double postime;      // History time of the positions the last next() produced

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
        postime = curtime;
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

// ---- block-level aggregation --------------------------------------------------------------------------------
// Appending to a global queue with one atomic per WARP (and waiting for the returned base) serialises the whole grid on
// one L2 address: measured 1.97 ms of a 2.0 ms kernel.  These helpers spend ONE atomic per block.  They contain
// __syncthreads(): every thread of the block must reach them, so the loops around them use block_rounded() trip counts.
__device__ __forceinline__ unsigned long long block_rounded(unsigned long long n) { return (n + blockDim.x - 1) / blockDim.x * blockDim.x; }

// reserve c (>= 0) consecutive slots of *counter for this thread; returns the index of its first slot
__device__ __forceinline__ unsigned long long block_alloc(unsigned c, unsigned long long *counter)
{
    __shared__ unsigned s_w[32];
    __shared__ unsigned long long s_b;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    unsigned incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) s_w[wib] = incl;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        unsigned tot = 0;
        for (int w = 0; w < nw; w++) { const unsigned x = s_w[w]; s_w[w] = tot; tot += x; }
        s_b = tot ? atomicAdd(counter, (unsigned long long)tot) : 0ull;
    }
    __syncthreads();
    const unsigned long long r = s_b + s_w[wib] + (incl - c);
    __syncthreads();
    return r;
}

// earliest TOI / hit count of the block merged with one atomic pair
__device__ __forceinline__ void reduce_block(bool hit, double toi, unsigned long long *earliest_bits, unsigned long long *nhit)
{
    __shared__ unsigned long long s_bits[32];
    __shared__ unsigned s_cnt[32];
    unsigned long long bits = hit ? (unsigned long long)__double_as_longlong(toi) : 0xFFFFFFFFFFFFFFFFull;
    unsigned cnt = hit ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        const unsigned long long ob = __shfl_xor_sync(0xffffffffu, bits, o);
        const unsigned oc = __shfl_xor_sync(0xffffffffu, cnt, o);
        bits = ob < bits ? ob : bits;
        cnt += oc;
    }
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) { s_bits[wib] = bits; s_cnt[wib] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int w = 1; w < nw; w++) { bits = s_bits[w] < bits ? s_bits[w] : bits; cnt += s_cnt[w]; }
        if (cnt)
        {
            atomicMin(earliest_bits, bits);
            atomicAdd(nhit, (unsigned long long)cnt);
        }
    }
    __syncthreads();
}

struct NpArgs
{
    long long n;
    const int *stencils;
    const double *eta_arr;
    double eta_all;
    const double *q0, *q1;
    int vstride;
    const float *vbox;          // single step: per-vertex float swept boxes {lo.xyz, hi.xyz, -, -} (pack_positions_kernel)
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
    double *tasks;              // REC_STRIDE doubles per record
    const double *tasks_other;  // records of the run before this one (vertex-edge tests shared across the two runs: sbase bit 31)
    int *rtag;                  // per record: its tag, kept in a compact array for the bucket kernel (a tag read out of the
                                // 128-byte records costs a DRAM sector per record and pass)
    unsigned long long *ntask;  // running number of task records
    unsigned long long task_cap;
};

struct StencilIn
{
    V3 a[4], b[4];
    double eta;
};

// Single-step positions are packed per vertex as {x0,y0,z0,-,x1,y1,z1,-} (64 bytes, pack_positions_kernel): the start
// and end position of a vertex come from one aligned 64-byte chunk with four 16-byte loads, instead of six 8-byte loads
// from two arrays whose 24-byte elements straddle sectors.
__device__ __forceinline__ void ldpair(const double *qpack, int idx, V3 &a, V3 &b)
{
    const double2 *p = reinterpret_cast<const double2 *>(qpack + 8ll * idx);
    const double2 p0 = __ldg(p), p1 = __ldg(p + 1), p2 = __ldg(p + 2), p3 = __ldg(p + 3);
    a = mk(p0.x, p0.y, p1.x);
    b = mk(p2.x, p2.y, p3.x);
}

template <bool IS_VF> __device__ __forceinline__ void load_single(const NpArgs &A, long long i, StencilIn &S)
{
    const int4 s = reinterpret_cast<const int4 *>(A.stencils)[i];
    S.eta = A.eta_arr ? A.eta_arr[i] : A.eta_all;
    ldpair(A.q0, s.x, S.a[0], S.b[0]);
    ldpair(A.q0, s.y, S.a[1], S.b[1]);
    ldpair(A.q0, s.z, S.a[2], S.b[2]);
    ldpair(A.q0, s.w, S.a[3], S.b[3]);
}

__global__ void __launch_bounds__(256) pack_positions_kernel(int V, const double *__restrict__ q0, const double *__restrict__ q1, int vstride,
                                                             double *__restrict__ qpack, float *__restrict__ vbox)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const double *a = q0 + (long long)vstride * v, *b = q1 + (long long)vstride * v;
    double2 *o = reinterpret_cast<double2 *>(qpack + 8ll * v);
    o[0] = make_double2(a[0], a[1]);
    o[1] = make_double2(a[2], 0.0);
    o[2] = make_double2(b[0], b[1]);
    o[3] = make_double2(b[2], 0.0);
    const BoxF bx = swept_box_f(mk(a[0], a[1], a[2]), mk(b[0], b[1], b[2]));
    float4 *f = reinterpret_cast<float4 *>(vbox + 8ll * v);
    f[0] = make_float4(bx.lo[0], bx.lo[1], bx.lo[2], bx.hi[0]);
    f[1] = make_float4(bx.hi[1], bx.hi[2], 0.f, 0.f);
}

__device__ __forceinline__ void store_result(const NpArgs &A, long long i, int stage, double toi)
{
    A.hit[i] = stage != 0;
    A.toi[i] = stage ? toi : 0.0;
    if (A.stage) A.stage[i] = (unsigned char)stage;
}


// ================================================================================================================
// Single-step pipeline.  Every kernel does one kind of work for all its lanes; kernels are connected by queues in HBM
// (indices only: positions are re-gathered, they live in L2).
//   cull      one thread per stencil: swept-box culling -> the sub-tests that have to be evaluated at all become items
//             of three queues (primitive / vertex-edge / vertex-vertex);
//   stage<S>  one thread per surviving stencil, ONE polynomial of the primitive per kernel (VF: e1,e2,e3,coplanarity;
//             EE: a0,a1,b0,b1, distance sextic): most stencils leave at the first or second polynomial, and the
//             survivors are compacted before the next one, so warps stay full.  The last stage reserves the records of
//             a surviving stencil and feeds one export queue per polynomial;
//   export<K> rebuilds polynomial K for the survivors that need its record (all lanes build the same polynomial);
//   ve, vv    one thread per queued vertex-edge / vertex-vertex test;
//   decide    one thread per stencil: final when pass 1 settled everything before the first hit, else the stencil goes
//             on the work list (deferred sub-tests) or to the general routine (rare shapes);
//   bucket    pending records by reduced degree 3..6;
//   solve<D>  root isolation + interval rules, one lane per record, lanes refilled from a shared cursor;
//             two rounds: first the distance polynomials, then — after `window` has finalised every inside polynomial
//             that keeps one sign on the (short) intervals of its distance polynomial — the few that are left;
//   combine   work list: interval lists of each deferred sub-test intersected in the reference's order of sub-tests;
//   general   the few stencils left: the reference's whole sequence in one thread (stencil_segment_full).
// ================================================================================================================
enum
{
    K_NWORK = 0, K_NTASK = 1, K_NDEG = 2 /* 4 */, K_NQ = 6 /* prim, ve, vv */, K_NGEN = 9, K_NSQ = 10 /* after stage 0..3 */,
    K_NXQ = 14 /* polynomial 0..4 */, K_CURSOR = 19 /* degree 3..6 */, K_NDEG2 = 23 /* second solve round */, K_CURSOR2 = 27, K_NVE2 = 31 /* unique vertex-edge tests that need records */, K_NVEU = CCD_NP_KVEU /* unique vertex-edge tests */, K_COUNT = 33
};

struct P1Args
{
    NpArgs A;
    unsigned *status;            // per stencil: 2 bits per sub-test (SC_*)
    int *sbase;                  // per stencil, per deferred sub-test 0..4: first record | number of records << 28
    int *qprim, *qve, *qvv;      // queues out of cull: stencil index (| sub-test << 28)
    int *qve2;                   // unique vertex-edge tests whose first look did not settle them (np_ve_kernel -> np_ve_rec_kernel)
    // vertex-edge de-duplication (np_ve_key_kernel): the same (vertex, edge) meets in ~6 stencils
    unsigned long long *vkeys;   // hash set of packed (vertex, edge vertex, edge vertex) triples, vslots entries (power of two)
    int *vslotval;               // per slot: index of the unique test
    int *vitem;                  // per queued item: its unique test (| VE_DIRECT) or the slot that holds it
    int *vulist;                 // unique tests: a representative item (stencil | sub-test << 28)
    int *vures;                  // per unique test: -1 miss, else first record | number of records << 28
    unsigned vslots;
    // the run before this one (edge-edge, when this is the vertex-face run of the same step with the same eta): its hash
    // set is consulted first, and a test found there is never evaluated again — its records are read out of that run's
    // record buffer (sbase bit 31)
    const unsigned long long *pkeys;
    const int *pslotval, *pures;
    unsigned pslots;             // 0: no previous run to consult
    bool vdedup;                 // false: per-stencil eta or vertex ids beyond 21 bits -> every item is its own test
    int2 *sq[2];                 // stage queues (ping-pong): {stencil, state}
    int2 *xq[5];                 // export queues per polynomial: {stencil, record slot}
    int *qgen;                   // stencils for the general routine
    unsigned long long *ctr;     // counters, K_* above
};

__device__ __forceinline__ void store_record(double *dst, const double (&rec)[8])
{
    double2 *d = reinterpret_cast<double2 *>(dst);
    d[0] = make_double2(rec[0], rec[1]);
    d[1] = make_double2(rec[2], rec[3]);
    d[2] = make_double2(rec[4], rec[5]);
    d[3] = make_double2(rec[6], rec[7]);
}

__device__ __forceinline__ BoxF ldbox(const float *vbox, int idx)
{
    const float4 *p = reinterpret_cast<const float4 *>(vbox + 8ll * idx);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    BoxF r;
    r.lo[0] = a.x; r.lo[1] = a.y; r.lo[2] = a.z;
    r.hi[0] = a.w; r.hi[1] = b.x; r.hi[2] = b.y;
    return r;
}

// one thread per stencil, float swept boxes only (CullF): which sub-tests have to be evaluated at all
template <bool IS_VF> __global__ void __launch_bounds__(256) np_cull_kernel(P1Args Q)
{
    const NpArgs &A = Q.A;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int NVE = Subs<IS_VF>::NVE, NVV = Subs<IS_VF>::NVV;
    unsigned todo = 0;
    if (i < A.n)
    {
        const int4 s = reinterpret_cast<const int4 *>(A.stencils)[i];
        CullF<IS_VF> c;
        c.bx[0] = ldbox(A.vbox, s.x);
        c.bx[1] = ldbox(A.vbox, s.y);
        c.bx[2] = ldbox(A.vbox, s.z);
        c.bx[3] = ldbox(A.vbox, s.w);
        c.init(A.eta_arr ? A.eta_arr[i] : A.eta_all);
        if (!c.stencil_apart()) c.init13();      // the swept boxes overlap: the ten diagonal directions, still from the boxes alone
        todo = c.todo();
        Q.status[i] = 0u;
    }
    // Queue appends with ONE atomic per block and queue: every warp of the grid appending to the same three counters
    // with its own atomic (and waiting for the returned base) was what this kernel spent its time on.
    __shared__ unsigned s_warp[3][8];
    __shared__ unsigned long long s_base[3];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned ve_bits = (todo >> 1) & ((1u << NVE) - 1u), vv_bits = (todo >> (NVE + 1)) & ((1u << NVV) - 1u);
    unsigned c[3] = {IS_VF ? 0u : (todo & 1u), (unsigned)__popc(ve_bits), (unsigned)__popc(vv_bits)};      // VF: every stencil runs its primitive
    unsigned pre[3];
#pragma unroll
    for (int q = 0; q < 3; q++)
    {
        unsigned incl = c[q];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        pre[q] = incl - c[q];
        if (lane == 31) s_warp[q][wib] = incl;
    }
    __syncthreads();
    if (threadIdx.x < 3)
    {
        unsigned tot = 0;
        for (int w = 0; w < 8; w++) { const unsigned x = s_warp[threadIdx.x][w]; s_warp[threadIdx.x][w] = tot; tot += x; }
        s_base[threadIdx.x] = tot ? atomicAdd(Q.ctr + K_NQ + threadIdx.x, (unsigned long long)tot) : 0ull;
    }
    __syncthreads();
    if (c[0]) Q.qprim[s_base[0] + s_warp[0][wib] + pre[0]] = (int)i;
    {
        unsigned long long o = s_base[1] + s_warp[1][wib] + pre[1];
#pragma unroll
        for (int sub = 1; sub <= NVE; sub++)
            if ((todo >> sub) & 1u) Q.qve[o++] = (int)i | (sub << 28);
    }
    {
        unsigned long long o = s_base[2] + s_warp[2][wib] + pre[2];
#pragma unroll
        for (int k = 0; k < NVV; k++)
            if ((todo >> (NVE + 1 + k)) & 1u) Q.qvv[o++] = (int)i | (k << 28);
    }
}

// Stage S (and, with TWO, stage S+1 for the stencils that survive S: the edge-edge quartics reject only ~20 % each, so a
// pair of them shares one gather of the positions).  in_q / out_q: which of the two stage queues is read / written.
template <bool IS_VF, int S, bool TWO> __global__ void __launch_bounds__(128, NP_MINB_STAGE) np_stage_kernel(P1Args Q, int in_q, int out_q)
{
    const NpArgs &A = Q.A;
    constexpr int NST = Prim<IS_VF>::NST;
    constexpr int SL = S + (TWO ? 1 : 0);      // the last stage this kernel runs
    constexpr bool LAST = (SL == NST - 1);
    constexpr int KOWN = Prim<IS_VF>::poly(SL);
    const unsigned long long n = (S == 0) ? (IS_VF ? (unsigned long long)A.n : Q.ctr[K_NQ]) : Q.ctr[K_NSQ + S - 1];
    const unsigned long long nround = block_rounded(n);
    const int2 *in = Q.sq[in_q];
    int2 *out = Q.sq[out_q];
    for (unsigned long long it = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; it < nround; it += (unsigned long long)gridDim.x * blockDim.x)
    {
        bool alive = false, has_rec = false;
        unsigned state = 0xff00u;
        long long i = 0;
        double rec[8];
        if (it < n)
        {
            if (S == 0) i = IS_VF ? (long long)it : (long long)Q.qprim[it];
            else { const int2 e = in[it]; i = e.x; state = (unsigned)e.y; }
            StencilIn St;
            load_single<IS_VF>(A, i, St);
            V3 v[4];
            for (int k = 0; k < 4; k++) v[k] = St.b[k] - St.a[k];
            unsigned so = state;
            if (TWO)
            {
                double dummy[8];
                bool h;
                alive = stage_item<IS_VF, S, false>(St.a, v, St.eta, state, so, dummy, h);
                state = so;
                if (alive) alive = stage_item<IS_VF, SL, LAST>(St.a, v, St.eta, state, so, rec, has_rec);
            }
            else
                alive = stage_item<IS_VF, S, LAST>(St.a, v, St.eta, state, so, rec, has_rec);
            state = so;
        }
        if (!LAST)
        {
            const unsigned long long o = block_alloc(alive ? 1u : 0u, Q.ctr + K_NSQ + SL);
            if (alive) out[o] = make_int2((int)i, (int)state);
        }
        else
        {
            // a survivor is deferred to the combine kernel with the records of its constrained polynomials — possibly
            // none: every list is the whole [0,1] then, and combining zero records gives exactly that (hit at t = 0,
            // subject to the edge-edge parallel test)
            const unsigned need = state & 0x1fu;
            const bool def = alive;
            // records of the surviving stencils of this block: one reservation
            const int cnt = def ? __popc(need) : 0;
            const unsigned long long t0 = block_alloc((unsigned)cnt, A.ntask);
            const bool fits = t0 + (unsigned long long)cnt <= A.task_cap && t0 + (unsigned long long)cnt < (1ull << 28);
            if (def)
            {
                Q.sbase[5 * i + 0] = (int)((t0 < 0x0fffffffull ? (unsigned)t0 : 0x0fffffffu) | ((unsigned)cnt << 28));
                atomicOr(&Q.status[i], (unsigned)SC_DEFERRED);
                if (has_rec && fits)
                {
                    const unsigned long long slot = t0 + __popc(need & ((1u << KOWN) - 1u));
                    store_record(A.tasks + (unsigned long long)REC_STRIDE * slot, rec);
                    A.rtag[slot] = (int)rec_untag(rec[7]);
                }
            }
#pragma unroll
            for (int k = 0; k < 5; k++)
                if (k != KOWN && k < NST)
                {
                    const bool want = def && fits && ((need >> k) & 1u);
                    const unsigned long long o = block_alloc(want ? 1u : 0u, Q.ctr + K_NXQ + k);
                    if (want) Q.xq[k][o] = make_int2((int)i, (int)(t0 + __popc(need & ((1u << k) - 1u))));
                }
        }
    }
}

template <bool IS_VF, int K> __global__ void __launch_bounds__(128, NP_MINB_STAGE) np_export_kernel(P1Args Q)
{
    const NpArgs &A = Q.A;
    const unsigned long long n = Q.ctr[K_NXQ + K];
    for (unsigned long long it = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; it < n; it += (unsigned long long)gridDim.x * blockDim.x)
    {
        const int2 e = Q.xq[K][it];
        StencilIn St;
        load_single<IS_VF>(A, e.x, St);
        V3 v[4];
        for (int k = 0; k < 4; k++) v[k] = St.b[k] - St.a[k];
        double rec[8];
        export_item<IS_VF, K>(St.a, v, St.eta, rec);
        store_record(A.tasks + (unsigned long long)REC_STRIDE * (unsigned)e.y, rec);
        A.rtag[(unsigned)e.y] = (int)rec_untag(rec[7]);
    }
}

// Vertex-edge tests.  The test of vertex v against edge (a, b) depends on nothing else of the stencil, and the same triple
// comes up in every stencil that pairs an edge at v with (a, b) and in both faces at (a, b): ~5 times at the 4M-triangle
// cloth.  np_ve_key_kernel therefore folds the queued items into UNIQUE tests through a hash set of packed triples (one
// 64-bit compare-and-swap per item; an item that cannot be placed within VE_PROBES probes simply becomes its own test), the
// two dense kernels below run once per unique test, and np_ve_resolve_kernel hands every item the result of its test —
// the records of a deferred test are shared by all its stencils (they are read-only once solved).
//   np_ve_kernel      classifies a unique test (miss / needs records) and compacts the survivors;
//   np_ve_rec_kernel  rebuilds those (the gather hits L2, the arithmetic is ~200 flop), refines the pending distance
//                     quartic on the windows of its inside quadratics (ve_item_refined: four out of five end here as a
//                     miss) and writes records only for the rest.
#define VE_DIRECT 0x80000000u
#define VE_PREV 0x40000000u      // the item's test lives in the previous run's table (slot in the low bits)
#define VE_PROBES 48
#define VE_EMPTY 0xFFFFFFFFFFFFFFFFull

template <bool IS_VF> __device__ __forceinline__ void ve_item_verts(const NpArgs &A, int item, long long &i, int &sub, int &v, int &a, int &b)
{
    i = item & 0x0fffffff;
    sub = (unsigned)item >> 28;
    int iv, i1, i2;
    Subs<IS_VF>::ve(sub, iv, i1, i2);
    const int4 s4 = reinterpret_cast<const int4 *>(A.stencils)[i];
    const int idx[4] = {s4.x, s4.y, s4.z, s4.w};
    v = idx[iv]; a = idx[i1]; b = idx[i2];
}

template <bool IS_VF> __global__ void __launch_bounds__(256) np_ve_key_kernel(P1Args Q)
{
    const NpArgs &A = Q.A;
    const unsigned long long n = Q.ctr[K_NQ + 1];
    const unsigned long long nround = block_rounded(n);
    const unsigned mask = Q.vslots - 1u;
    for (unsigned long long it = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; it < nround; it += (unsigned long long)gridDim.x * blockDim.x)
    {
        bool mine = false, placed = false, prev = false;
        int item = 0;
        unsigned h = 0;
        if (it < n)
        {
            item = Q.qve[it];
            mine = true;
            if (Q.vdedup)
            {
                long long i;
                int sub, v, a, b;
                ve_item_verts<IS_VF>(A, item, i, sub, v, a, b);
                const unsigned long long key = (unsigned long long)(unsigned)v | ((unsigned long long)(unsigned)a << 21) | ((unsigned long long)(unsigned)b << 42);
                unsigned long long x = key * 0x9E3779B97F4A7C15ull;
                x ^= x >> 29;
                if (Q.pslots)
                {
                    const unsigned pmask = Q.pslots - 1u;
                    unsigned hp = (unsigned)x & pmask;
                    for (int probe = 0; probe < VE_PROBES; probe++)
                    {
                        const unsigned long long old = Q.pkeys[hp];
                        if (old == VE_EMPTY) break;
                        if (old == key) { mine = false; prev = true; h = hp; break; }
                        hp = (hp + 1u) & pmask;
                    }
                }
                if (!prev) h = (unsigned)x & mask;
                for (int probe = 0; probe < VE_PROBES && !prev; probe++)
                {
                    const unsigned long long old = atomicCAS(&Q.vkeys[h], VE_EMPTY, key);
                    if (old == VE_EMPTY) { placed = true; break; }
                    if (old == key) { mine = false; break; }
                    h = (h + 1u) & mask;
                }
            }
        }
        const unsigned long long u = block_alloc(mine ? 1u : 0u, Q.ctr + K_NVEU);
        if (mine)
        {
            Q.vulist[u] = item;
            if (placed) Q.vslotval[h] = (int)u;
        }
        if (it < n) Q.vitem[it] = mine ? (int)((unsigned)u | VE_DIRECT) : (int)(prev ? (h | VE_PREV) : h);
    }
}

template <bool IS_VF> __global__ void __launch_bounds__(128, NP_MINB) np_ve_kernel(P1Args Q)
{
    const NpArgs &A = Q.A;
    const unsigned long long n = Q.ctr[K_NVEU];
    const unsigned long long nround = block_rounded(n);
    for (unsigned long long it = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; it < nround; it += (unsigned long long)gridDim.x * blockDim.x)
    {
        int code = SC_MISS;
        if (it < n)
        {
            long long i;
            int sub, iv, i1, i2;
            ve_item_verts<IS_VF>(A, Q.vulist[it], i, sub, iv, i1, i2);
            const double eta = A.eta_arr ? A.eta_arr[i] : A.eta_all;
            V3 a0, a1, a2, b0, b1, b2;
            ldpair(A.q0, iv, a0, b0);
            ldpair(A.q0, i1, a1, b1);
            ldpair(A.q0, i2, a2, b2);
            double recs[3][8];
            int nrec = 0;
            code = ve_item(a0, a1, a2, b0 - a0, b1 - a1, b2 - a2, eta, recs, nrec);
            Q.vures[it] = -1;
        }
        const unsigned long long o = block_alloc(code == SC_DEFERRED ? 1u : 0u, Q.ctr + K_NVE2);
        if (code == SC_DEFERRED) Q.qve2[o] = (int)it;
    }
}

template <bool IS_VF> __global__ void __launch_bounds__(128, NP_MINB) np_ve_rec_kernel(P1Args Q)
{
    const NpArgs &A = Q.A;
    const unsigned long long n = Q.ctr[K_NVE2];
    const unsigned long long nround = block_rounded(n);
    for (unsigned long long it = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; it < nround; it += (unsigned long long)gridDim.x * blockDim.x)
    {
        int code = SC_MISS, nrec = 0, u = 0;
        double recs[3][8];
        if (it < n)
        {
            u = Q.qve2[it];
            long long i;
            int sub, iv, i1, i2;
            ve_item_verts<IS_VF>(A, Q.vulist[u], i, sub, iv, i1, i2);
            const double eta = A.eta_arr ? A.eta_arr[i] : A.eta_all;
            V3 a0, a1, a2, b0, b1, b2;
            ldpair(A.q0, iv, a0, b0);
            ldpair(A.q0, i1, a1, b1);
            ldpair(A.q0, i2, a2, b2);
            code = ve_item_refined(a0, a1, a2, b0 - a0, b1 - a1, b2 - a2, eta, recs, nrec);
        }
        const unsigned c = code == SC_DEFERRED ? (unsigned)nrec : 0u;
        const unsigned long long t0 = block_alloc(c, A.ntask);
        if (code == SC_DEFERRED)      // also with nrec == 0: the combine kernel reads the record count from sbase
        {
            if (c && t0 + c <= A.task_cap && t0 + c < (1ull << 28))
            {
#pragma unroll
                for (int k = 0; k < 3; k++)
                    if (k < nrec)
                    {
                        store_record(A.tasks + (unsigned long long)REC_STRIDE * (t0 + k), recs[k]);
                        A.rtag[t0 + k] = (int)rec_untag(recs[k][7]);
                    }
            }
            Q.vures[u] = (int)((t0 < 0x0fffffffull ? (unsigned)t0 : 0x0fffffffu) | ((unsigned)nrec << 28));
        }
    }
}

// every queued item takes the result of its unique test
template <bool IS_VF> __global__ void __launch_bounds__(256) np_ve_resolve_kernel(P1Args Q)
{
    const unsigned long long n = Q.ctr[K_NQ + 1];
    for (unsigned long long it = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; it < n; it += (unsigned long long)gridDim.x * blockDim.x)
    {
        const unsigned ref = (unsigned)Q.vitem[it];
        int res;
        if (ref & VE_DIRECT) res = Q.vures[ref & ~VE_DIRECT];
        else if (ref & VE_PREV)
        {
            res = Q.pures[Q.pslotval[ref & ~VE_PREV]];
            if (res >= 0) res = (int)((unsigned)res | 0x80000000u);      // records in the previous run's buffer
            else res = -1;
        }
        else res = Q.vures[Q.vslotval[ref]];
        if (res == -1) continue;
        const int item = Q.qve[it];
        const long long i = item & 0x0fffffff;
        const int sub = (unsigned)item >> 28;
        Q.sbase[5 * i + sub] = res;
        atomicOr(&Q.status[i], (unsigned)SC_DEFERRED << (2 * sub));
    }
}

template <bool IS_VF> __global__ void __launch_bounds__(256) np_vv_kernel(P1Args Q)
{
    const NpArgs &A = Q.A;
    const unsigned long long n = Q.ctr[K_NQ + 2];
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
        V3 a1, a2, b1, b2;
        ldpair(A.q0, idx[i1], a1, b1);
        ldpair(A.q0, idx[i2], a2, b2);
        const V3 v1 = b1 - a1, v2 = b2 - a2;
        double t = 0.0;
        if (vertex_vertex(a1, a2, v1, v2, eta, t) == R_HIT) atomicOr(&Q.status[i], (unsigned)SC_HIT << (2 * (NVE + 1 + k)));
    }
}

// time of impact of vertex-vertex sub-test `sub` (pass 1 only recorded that it hits)
template <bool IS_VF> __device__ __forceinline__ double vv_toi(const StencilIn &S, int sub)
{
    int i1, i2;
    Subs<IS_VF>::vv(sub - Subs<IS_VF>::NVE - 1, i1, i2);
    double t = 0.0;
    vertex_vertex(S.a[i1], S.a[i2], S.b[i1] - S.a[i1], S.b[i2] - S.a[i2], S.eta, t);
    return t;
}

// Persistent grid-stride kernel, FOUR consecutive stencils per thread: the status words arrive as one 16-byte load, the
// results leave as 4-byte / 32-byte stores, and the block's barriers (work-list reservation, reduction) are paid once per
// 1024 stencils.  (Pass 1 no longer produces SC_GENERAL: unconstrained sub-tests are deferred with zero records.)
template <bool IS_VF> __global__ void __launch_bounds__(256) np_decide_kernel(P1Args Q)
{
    const NpArgs &A = Q.A;
    constexpr int NSUB = 1 + Subs<IS_VF>::NVE + Subs<IS_VF>::NVV;
    const unsigned long long n = (unsigned long long)A.n, nquad = (n + 3) / 4, nround = block_rounded(nquad);
    for (unsigned long long it = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; it < nround; it += (unsigned long long)gridDim.x * blockDim.x)
    {
        const long long i0 = 4ll * (long long)it;
        unsigned st4[4] = {0, 0, 0, 0};
        if (i0 + 3 < (long long)n)
        {
            const uint4 s = reinterpret_cast<const uint4 *>(Q.status)[it];
            st4[0] = s.x; st4[1] = s.y; st4[2] = s.z; st4[3] = s.w;
        }
        else
            for (int k = 0; k < 4; k++)
                if (i0 + k < (long long)n) st4[k] = Q.status[i0 + k];
        int stage[4] = {0, 0, 0, 0}, meta[4] = {0, 0, 0, 0};
        double toi[4] = {0.0, 0.0, 0.0, 0.0};
        unsigned ndef = 0;
        bool any_hit = false;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const unsigned st = st4[k];
            if (!st) continue;
            unsigned submask = 0;
            int later_hit = 255, nd = 0;
            for (int sub = 0; sub < NSUB; sub++)
            {
                const unsigned r = (st >> (2 * sub)) & 3u;
                if (r == (unsigned)SC_HIT)
                {
                    if (nd == 0) stage[k] = sub + 1; else later_hit = sub;
                    break;
                }
                if (r == (unsigned)SC_DEFERRED) { nd++; submask |= 1u << sub; }
            }
            if (nd == 0)
            {
                if (stage[k])
                {
                    StencilIn S;
                    load_single<IS_VF>(A, i0 + k, S);
                    toi[k] = vv_toi<IS_VF>(S, stage[k] - 1);
                    any_hit = true;
                }
            }
            else { stage[k] = -1; meta[k] = (int)(submask | ((unsigned)later_hit << 8)); ndef++; }
        }
        // results of the stencils that are settled
        if (i0 + 3 < (long long)n)
        {
            reinterpret_cast<uchar4 *>(A.hit)[it] = make_uchar4(stage[0] > 0, stage[1] > 0, stage[2] > 0, stage[3] > 0);
            if (A.stage) reinterpret_cast<uchar4 *>(A.stage)[it] = make_uchar4((unsigned char)max(stage[0], 0), (unsigned char)max(stage[1], 0),
                                                                              (unsigned char)max(stage[2], 0), (unsigned char)max(stage[3], 0));
            double2 *t2 = reinterpret_cast<double2 *>(A.toi + i0);
            t2[0] = make_double2(toi[0], toi[1]);
            t2[1] = make_double2(toi[2], toi[3]);
        }
        else
            for (int k = 0; k < 4; k++)
                if (i0 + k < (long long)n) store_result(A, i0 + k, max(stage[k], 0), toi[k]);
        unsigned long long w = block_alloc(ndef, A.nwork);
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (stage[k] < 0)
            {
                const long long i = i0 + k;
                A.w_stencil[w] = (int)i;
                A.w_meta[w] = meta[k];
                int j = 0;
                for (unsigned m = (unsigned)meta[k] & 0xffu; m; m &= m - 1) A.w_base[5 * w + (j++)] = Q.sbase[5 * i + (__ffs(m) - 1)];
                for (; j < 5; j++) A.w_base[5 * w + j] = 0;
                w++;
            }
        if (__syncthreads_or(any_hit))      // vertex-vertex hits only: rare
            for (int k = 0; k < 4; k++) reduce_block(stage[k] > 0, toi[k], A.earliest_bits, A.nhit);
    }
}

// pending records by reduced degree (3..6); final records (closed forms) are skipped
// phase 0: the distance polynomials ("<= 0 wanted": VF / EE sextic, vertex-edge quartic); phase 1: the inside
// polynomials (">= 0 wanted") the window stage left pending
__global__ void __launch_bounds__(256) bucket_tasks_kernel(const int *__restrict__ rtag, const unsigned long long *ntask_ptr,
                                                           unsigned long long cap, int *__restrict__ lists, unsigned long long *counts, int phase)
{
    // the four degrees share ONE block-wide scan: their running counts are packed 16 bits each into a 64-bit word (a block
    // of 256 threads adds at most 256 to any of them)
    __shared__ unsigned long long s_w[8];
    __shared__ unsigned long long s_base[4];
    unsigned long long nt = *ntask_ptr;
    if (nt > cap) nt = cap;
    const unsigned long long nround = block_rounded(nt);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < nround;
         j += (unsigned long long)gridDim.x * blockDim.x)
    {
        int rd = 0;
        if (j < nt)
        {
            const unsigned tag = (unsigned)rtag[j];
            if (!(tag & REC_FINAL) && ((phase == 1) == ((tag & REC_POS) != 0))) rd = (int)((tag >> 4) & 7u);
        }
        const unsigned long long mine = (rd >= 3 && rd <= 6) ? 1ull << (16 * (rd - 3)) : 0ull;
        unsigned long long incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) s_w[wib] = incl;
        __syncthreads();
        if (threadIdx.x == 0)
        {
            unsigned long long tot = 0;
            for (int w = 0; w < 8; w++) { const unsigned long long x = s_w[w]; s_w[w] = tot; tot += x; }
#pragma unroll
            for (int d = 0; d < 4; d++)
            {
                const unsigned long long c = (tot >> (16 * d)) & 0xffffull;
                s_base[d] = c ? atomicAdd(&counts[d], c) : 0ull;
            }
        }
        __syncthreads();
        if (mine)
        {
            const int d = rd - 3;
            const unsigned long long before = ((s_w[wib] + incl - mine) >> (16 * d)) & 0xffffull;
            lists[(size_t)d * cap + s_base[d] + before] = (int)j;
        }
        __syncthreads();
    }
}

// The root isolator runs as three kernels per degree so that each of them is uniform work:
//   prepare<D>   one lane per pending record: Bernstein sign variations down the derivative chain (RootLane::prepare);
//   solve<D>     the climb: all lanes of a warp with a solve in flight iterate ONE Newton loop together, lanes walk their
//                pieces / levels in between, and a lane that has finished its polynomial takes the next record from
//                the shared cursor; roots go to the record's scratch half;
//   finalize<D>  one lane per record: the interval rules (checkInterval midpoints), record rewritten as a final one.
template <int D> __global__ void __launch_bounds__(128) prepare_kernel(double *tasks, const int *__restrict__ list, const unsigned long long *count_ptr)
{
    const unsigned long long nt = *count_ptr;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < nt; w += (unsigned long long)gridDim.x * blockDim.x)
    {
        double *rec = tasks + (long long)REC_STRIDE * list[w];
        double c[D + 1];
#pragma unroll
        for (int k = 0; k <= D; k++) c[k] = rec[k];
        RootLane<D> L;
        L.prepare(c);
        L.save_start(rec + 8);
    }
}

#ifndef SOLVE_MINB6
#define SOLVE_MINB6 6      // sextic solve: 6 resident blocks (85 registers) beat 5 (86 needed, 102 allowed) by 4 %, measured
#endif
#define SOLVE_MINB(D) ((D) <= 4 ? 6 : SOLVE_MINB6)
#ifndef SOLVE_CHUNK
#define SOLVE_CHUNK 64
#endif
#ifndef SOLVE_KEEP_NUM
#define SOLVE_KEEP_NUM 2      // a Newton round ends when fewer than SOLVE_KEEP_NUM/4 of its lanes are still iterating
#endif
template <int D>
__global__ void __launch_bounds__(128, SOLVE_MINB(D)) solve_kernel(double *tasks, const int *__restrict__ list, const unsigned long long *count_ptr,
                                                                   unsigned long long *cursor)
{
    const unsigned long long nt = *count_ptr;
    const int lane = threadIdx.x & 31;
    __shared__ double s_lane[(3 * D + 1) * 128];      // polynomial + two root lists of every lane, element k at [k * 128 + thread]
    RootLane<D, 128> L;
    L.bind(s_lane + threadIdx.x);
    L.done = true;
    L.solving = false;
    L.got_root = false;
    bool have = false, exhausted = false, gex = false;
    unsigned long long wnext = 0, wend = 0;      // the warp's current chunk of the list (warp-uniform)
    double *rec = nullptr;
    for (;;)
    {
        const unsigned need = __ballot_sync(0xffffffffu, !have && !exhausted);
        if (need)
        {
            // records are taken from a chunk the warp reserved with one atomic (SOLVE_CHUNK at a time): a refill per
            // round per warp on the one shared cursor would serialise all warps of the grid on that address
            if (wnext >= wend && !gex)
            {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(cursor, (unsigned long long)SOLVE_CHUNK);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= nt) { gex = true; wnext = wend = 0; }
                else { wnext = base; wend = base + SOLVE_CHUNK < nt ? base + SOLVE_CHUNK : nt; }
            }
            if (!have && !exhausted)
            {
                const unsigned long long my = wnext + __popc(need & ((1u << lane) - 1));
                if (my < wend)
                {
                    rec = tasks + (long long)REC_STRIDE * list[my];
                    double c[D + 1];
#pragma unroll
                    for (int k = 0; k <= D; k++) c[k] = rec[k];
                    const double aux[3] = {rec[8], rec[9], rec[10]};
                    L.load_start(c, aux);
                    have = true;
                }
                else if (gex)
                    exhausted = true;
            }
            const unsigned long long adv = wnext + (unsigned long long)__popc(need);
            wnext = adv < wend ? adv : wend;
        }
        if (!__any_sync(0xffffffffu, have))
            break;
        if (have && !L.solving) L.advance();
        if (have && L.done)
        {
            double r[6];
            const int nr = L.result(r);
#pragma unroll
            for (int k = 0; k < 6; k++) rec[8 + k] = r[k];
            rec[14] = (double)nr;
            have = false;
        }
        // Newton rounds for the lanes with a solve in flight.  Iteration counts have a long tail (near-double roots are
        // the rule for the distance sextics), so the round ends as soon as half of its lanes have converged: those go
        // back to advance() / take new records while the slow solves simply continue in the next round.
        unsigned ms = __ballot_sync(0xffffffffu, have && L.solving);
        if (ms)
        {
            const int thresh = max(1, (__popc(ms) * SOLVE_KEEP_NUM) >> 2);
            do
            {
                if (have && L.solving) L.newton_step();
                ms = __ballot_sync(0xffffffffu, have && L.solving);
            } while (__popc(ms) >= thresh);
        }
    }
}

template <int D> __global__ void __launch_bounds__(128) finalize_kernel(double *tasks, const int *__restrict__ list, const unsigned long long *count_ptr)
{
    const unsigned long long nt = *count_ptr;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < nt; w += (unsigned long long)gridDim.x * blockDim.x)
    {
        double *rec = tasks + (long long)REC_STRIDE * list[w];
        double c[D + 1], r[6];
#pragma unroll
        for (int k = 0; k <= D; k++) c[k] = rec[k];
#pragma unroll
        for (int k = 0; k < 6; k++) r[k] = rec[8 + k];
        finalize_record<D>(c, rec_untag(rec[7]), r, (int)rec[14], rec);
    }
}

// work list, between the two solve rounds: inside polynomials that keep one sign on every window of the solved distance
// polynomial are finalised without root isolation (window_item, ccd_stages.cuh)
template <bool IS_VF> __global__ void __launch_bounds__(128) np_window_kernel(P1Args Q)
{
    const NpArgs &A = Q.A;
    const unsigned long long nw = *A.nwork;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < nw; w += (unsigned long long)gridDim.x * blockDim.x)
    {
        if (!(A.w_meta[w] & 1)) continue;      // primitive not deferred
        const unsigned b = (unsigned)A.w_base[5 * w];
        const unsigned long long t0 = b & 0x0fffffffu;
        const int nrec = (int)((b >> 28) & 7u);
        if (t0 + (unsigned long long)nrec > A.task_cap) continue;
        window_item(A.tasks + (unsigned long long)REC_STRIDE * t0, nrec, IS_VF ? 3 : 4);
        for (int j = 0; j + 1 < nrec; j++) A.rtag[t0 + j] = (int)rec_untag(A.tasks[(unsigned long long)REC_STRIDE * (t0 + j) + 7]);
    }
}

// work list: the deferred sub-tests of a stencil in the reference's order; first hit wins, else the vertex-vertex hit
// pass 1 found behind them (if any)
template <bool IS_VF> __global__ void __launch_bounds__(128, NP_MINB) np_combine_kernel(P1Args Q)
{
    const NpArgs &A = Q.A;
    const unsigned long long nw = *A.nwork;
    const unsigned long long nround = block_rounded(nw);
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < nround;
         w += (unsigned long long)gridDim.x * blockDim.x)
    {
        int stage = 0;
        double toi = 0.0;
        bool fb = false;
        long long i = 0;
        if (w < nw)
        {
            i = A.w_stencil[w];
            const int meta = A.w_meta[w];
            unsigned submask = (unsigned)meta & 0xffu;
            const int later_hit = (meta >> 8) & 0xff;
            // the positions are needed only by the edge-edge primitive's parallel test and by a vertex-vertex hit behind
            // the deferred sub-tests
            StencilIn S;
            V3 v[4];
            const bool need_pos = (!IS_VF && (submask & 1u)) || later_hit != 255;
            if (need_pos)
            {
                load_single<IS_VF>(A, i, S);
                for (int k = 0; k < 4; k++) v[k] = S.b[k] - S.a[k];
            }
            else
                for (int k = 0; k < 4; k++) { S.a[k] = mk(0, 0, 0); S.b[k] = mk(0, 0, 0); v[k] = mk(0, 0, 0); }
            int j = 0;
            bool settled = false;
            while (submask)
            {
                const int sub = __ffs(submask) - 1;
                submask &= submask - 1;
                const unsigned b = (unsigned)A.w_base[5 * w + j];
                j++;
                const unsigned long long t0 = b & 0x0fffffffu;
                const int nrec = (int)((b >> 28) & 7u);
                const bool other = (b & 0x80000000u) != 0;
                if (!other && t0 + (unsigned long long)nrec > A.task_cap) { settled = true; break; }      // never written: the caller grows the buffer and reruns
                const double *tb = other ? A.tasks_other : A.tasks;
                const int r = combine_records(tb + (unsigned long long)REC_STRIDE * t0, nrec, !IS_VF && sub == 0, S.a, v, toi);
                if (r == RS_FALLBACK) { fb = true; settled = true; break; }
                if (r == RS_HIT) { stage = sub + 1; settled = true; break; }
            }
            if (!settled && later_hit != 255)
            {
                stage = later_hit + 1;
                toi = vv_toi<IS_VF>(S, later_hit);
            }
            if (!fb) store_result(A, i, stage, toi);
            else stage = 0;
        }
        {
            const unsigned long long o = block_alloc(fb ? 1u : 0u, Q.ctr + K_NGEN);
            if (fb) Q.qgen[o] = (int)i;
        }
        reduce_block(stage > 0, toi, A.earliest_bits, A.nhit);
    }
}

// ONE stencil per warp (lane 0): the general routine is long, branchy, local-memory code; 32 of them in one warp run one
// after the other in divergent lock-step, and the few hundred stencils that need it set the kernel's duration
template <bool IS_VF> __global__ void __launch_bounds__(128) np_general_kernel(P1Args Q)
{
    const NpArgs &A = Q.A;
    const unsigned long long n = Q.ctr[K_NGEN];
    const unsigned long long wpb = blockDim.x >> 5, nwarps = (unsigned long long)gridDim.x * wpb;
    const unsigned long long nround = (n + nwarps - 1) / nwarps * nwarps;      // same trip count for every warp of the grid
    const int lane = threadIdx.x & 31;
    for (unsigned long long x = (unsigned long long)blockIdx.x * wpb + (threadIdx.x >> 5); x < nround; x += nwarps)
    {
        int stage = 0;
        double toi = 0.0;
        if (x < n && lane == 0)
        {
            const long long i = Q.qgen[x];
            StencilIn S;
            load_single<IS_VF>(A, i, S);
            stage = stencil_segment_full<IS_VF>(S.a, S.b, S.eta, toi);
            store_result(A, i, stage, toi);
        }
        reduce_block(stage > 0, toi, A.earliest_bits, A.nhit);
    }
}

// SeparatingPlaneNarrowPhase::findCollisions (src/SeparatingPlaneNarrowPhase.cpp:11-25): one thread per stencil runs the
// interval search of ccd_sepplane.cuh; flag = 1 hit, 0 miss.  err counts stencils whose interval stack overflowed.
template <bool IS_VF>
__global__ void __launch_bounds__(128) sepplane_kernel(long long n, const int *__restrict__ stencils, const double *__restrict__ eta_arr, HistView H,
                                                       double eps, unsigned char *__restrict__ hit, unsigned long long *nhit, unsigned long long *err)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int r = 0;
    if (i < n)
    {
        const int4 s = reinterpret_cast<const int4 *>(stencils)[i];
        const int verts[4] = {s.x, s.y, s.z, s.w};
        r = sp_check_stencil<IS_VF>(H, verts, eta_arr[i], eps);
        hit[i] = r > 0;
        if (r < 0) atomicAdd(err, 1ull);
    }
    const unsigned long long o = block_alloc(r > 0 ? 1u : 0u, nhit);
    (void)o;
}

// ---- SeparatingPlaneNarrowPhase, staged ----------------------------------------------------------------------------------
// The reference recurses on time intervals per stencil (src/SeparatingPlaneNarrowPhase.cpp:47-218); sepplane_kernel above keeps
// that recursion on a 96-deep per-thread stack, one thread per stencil, and the stencil with the deepest search sets the
// kernel's duration.  Staged: the open intervals of ALL stencils sit in one global queue and are processed in rounds — one
// thread per interval runs one step of the search (sp_interval_step: midpoint distance, separating plane, first crossings
// along the four trajectories), appends the sub-intervals that remain to the next round's queue, marks a stencil that comes
// within eta at a midpoint, and turns an interval short enough for the CTCD primitives into a LEAF: a virtual single-step
// stencil on four virtual vertices.  When the queue runs empty the dense single-step pipeline tests all leaves at once and
// sp_leaf_or_kernel folds their flags into the stencils' (the result is a disjunction over intervals: order is irrelevant).
struct SpRoundArgs
{
    HistView H;
    const int *stencils[2];               // [0] vertex-face, [1] edge-edge
    const double *eta_arr[2];
    double eps;
    long long n[2];                       // round 0: one interval [0,1] per stencil, vertex-face stencils first
    const int *in_st;                     // later rounds: the queue the round before wrote; stencil index | type << 31
    const double *in_lo, *in_hi;
    const unsigned long long *nin;
    int *out_st;
    double *out_lo, *out_hi;
    unsigned long long *nout;
    unsigned long long qcap;              // counts keep running past the capacities (writes are guarded): the caller grows and reruns
    unsigned char *hit[2];                // per stencil
    unsigned long long *nleaf;            // [2]
    unsigned long long leaf_cap[2];
    long long vbase[2];                   // first virtual stencil of each type (two regions of the leaf arrays)
    double *q0v, *q1v;                    // 12 doubles per leaf each
    int *vst;                             // 4 virtual vertex ids per leaf
    double *veta;
    int *leaf_stencil;
};

template <bool FIRST> __global__ void __launch_bounds__(128) sp_round_kernel(SpRoundArgs G)
{
    const unsigned long long n = FIRST ? (unsigned long long)(G.n[0] + G.n[1]) : (*G.nin < G.qcap ? *G.nin : G.qcap);
    const unsigned long long nround = block_rounded(n);
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < nround; w += (unsigned long long)gridDim.x * blockDim.x)
    {
        int r = SP_DROP, nsub = 0, st = 0, ty = 0;
        double lo[2] = {0, 0}, hi[2] = {0, 0}, eta = 0;
        V3 oldpos[4], newpos[4];
        if (w < n)
        {
            if (FIRST)
            {
                ty = (long long)w >= G.n[0] ? 1 : 0;
                st = (int)((long long)w - (ty ? G.n[0] : 0));
            }
            else
            {
                const unsigned e = (unsigned)G.in_st[w];
                ty = (int)(e >> 31);
                st = (int)(e & 0x7fffffffu);
            }
            if (!G.hit[ty][st])      // (a stale 0 only costs work: the flag is a disjunction)
            {
                const int4 s4 = reinterpret_cast<const int4 *>(G.stencils[ty])[st];
                const int verts[4] = {s4.x, s4.y, s4.z, s4.w};
                eta = G.eta_arr[ty][st];
                const double mint = FIRST ? 0.0 : G.in_lo[w], maxt = FIRST ? 1.0 : G.in_hi[w];
                if (ty == 0) r = sp_interval_step<true>(G.H, verts, eta, G.eps, mint, maxt, oldpos, newpos, nsub, lo, hi);
                else r = sp_interval_step<false>(G.H, verts, eta, G.eps, mint, maxt, oldpos, newpos, nsub, lo, hi);
                if (r == SP_HIT) G.hit[ty][st] = 1;
            }
        }
        // (nothing to append anywhere in this block: skip the three reservations' barriers)
        if (!__syncthreads_or(nsub != 0 || r == SP_LEAF)) continue;
        const unsigned long long o = block_alloc((unsigned)nsub, G.nout);
        const unsigned long long l0 = block_alloc((r == SP_LEAF && ty == 0) ? 1u : 0u, G.nleaf + 0);
        const unsigned long long l1 = block_alloc((r == SP_LEAF && ty == 1) ? 1u : 0u, G.nleaf + 1);
        for (int k = 0; k < nsub; k++)
            if (o + k < G.qcap)
            {
                G.out_st[o + k] = (int)((unsigned)st | ((unsigned)ty << 31));
                G.out_lo[o + k] = lo[k];
                G.out_hi[o + k] = hi[k];
            }
        const unsigned long long l = ty ? l1 : l0;
        if (r == SP_LEAF && l < G.leaf_cap[ty])
        {
            const long long sg = G.vbase[ty] + (long long)l;
            for (int j = 0; j < 4; j++)
            {
                double *o0 = G.q0v + 3 * (4 * sg + j), *o1 = G.q1v + 3 * (4 * sg + j);
                o0[0] = oldpos[j].x; o0[1] = oldpos[j].y; o0[2] = oldpos[j].z;
                o1[0] = newpos[j].x; o1[1] = newpos[j].y; o1[2] = newpos[j].z;
                G.vst[4 * sg + j] = (int)(4 * sg + j);
            }
            G.veta[sg] = eta;
            G.leaf_stencil[sg] = st;
        }
    }
}

// a few leaves (a small mesh): the whole CTCD sequence per leaf in one kernel beats the ~45 launches of the dense pipeline
template <bool IS_VF>
__global__ void __launch_bounds__(128) sp_leaf_direct_kernel(long long nleaf, long long vbase, const double *__restrict__ q0v, const double *__restrict__ q1v,
                                                             const double *__restrict__ veta, const int *__restrict__ leaf_stencil, unsigned char *hit)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nleaf) return;
    const long long sg = vbase + k;
    V3 a[4], b[4];
    for (int j = 0; j < 4; j++) { a[j] = ldv(q0v + 3 * (4 * sg + j)); b[j] = ldv(q1v + 3 * (4 * sg + j)); }
    double t;
    if (stencil_segment_full<IS_VF>(a, b, veta[sg], t) != 0) hit[leaf_stencil[sg]] = 1;
}

__global__ void __launch_bounds__(256) sp_leaf_or_kernel(long long nleaf, const int *__restrict__ leaf_stencil, const unsigned char *__restrict__ hitv,
                                                         unsigned char *hit)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nleaf && hitv[k]) hit[leaf_stencil[k]] = 1;
}

__global__ void __launch_bounds__(256) count_flags_kernel(long long n, const unsigned char *__restrict__ flag, unsigned long long *count)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long o = block_alloc((i < n && flag[i]) ? 1u : 0u, count);
    (void)o;
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
        {
            double ta = st.postime;
            while (st.next(b))
            {
                stage = stencil_segment_full<IS_VF>(a, b, eta, toi);
                // the primitives return the parameter inside the stitched segment: map it back to History time
                if (stage) { toi = ta + toi * (st.postime - ta); break; }
                for (int k = 0; k < 4; k++) a[k] = b[k];
                ta = st.postime;
            }
        }
        store_result(A, i, stage, toi);
    }
    reduce_block(stage != 0, toi, A.earliest_bits, A.nhit);
}

// ---- multi-entry History, staged --------------------------------------------------------------------------------------
// Every stitched linear segment of a stencil (src/History.cpp:98-140) is an independent single-step test, and the stencil's
// answer is the first segment (in time) that hits.  So instead of one thread walking all segments of its stencil with the
// whole algorithm inlined (stencil_history_kernel: the one-thread-per-stencil shape the single-step path left behind at
// 5.7 of 32 lanes), the (stencil, segment) pairs are EXPANDED into virtual single-step stencils — four virtual vertices
// each, holding the stitched start and end positions — the dense single-step pipeline above runs on all of them at once,
// and hist_reduce_kernel takes, per stencil, its first hitting segment and maps the time of impact back to History time.
//   pass COUNT: segments per stencil;  pass fill (after the scan): positions, virtual stencils, thickness, segment times
template <bool COUNT>
__global__ void __launch_bounds__(128) hist_expand_kernel(long long n, const int *__restrict__ stencils, const double *__restrict__ eta_arr, double eta_all,
                                                          const long long *__restrict__ hoff, const double *__restrict__ htime, const double *__restrict__ hpos,
                                                          int *__restrict__ seg_count, const long long *__restrict__ seg_off, long long vbase,
                                                          double *__restrict__ q0v, double *__restrict__ q1v, int *__restrict__ vst, double *__restrict__ veta,
                                                          double *__restrict__ vtime)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 s = reinterpret_cast<const int4 *>(stencils)[i];
    const int verts[4] = {s.x, s.y, s.z, s.w};
    V3 a[4], b[4];
    Stitcher st;
    st.begin(hoff, htime, hpos, verts);
    int k = 0;
    if (st.next(a))
    {
        double ta = st.postime;
        while (st.next(b))
        {
            if (!COUNT)
            {
                const long long sl = seg_off[i] + k, sg = vbase + sl;      // index inside this type's run / among all virtual stencils
                for (int j = 0; j < 4; j++)
                {
                    double *o0 = q0v + 3 * (4 * sg + j), *o1 = q1v + 3 * (4 * sg + j);
                    o0[0] = a[j].x; o0[1] = a[j].y; o0[2] = a[j].z;
                    o1[0] = b[j].x; o1[1] = b[j].y; o1[2] = b[j].z;
                    vst[4 * sl + j] = (int)(4 * sg + j);
                }
                veta[sl] = eta_arr ? eta_arr[i] : eta_all;
                vtime[2 * sl] = ta;
                vtime[2 * sl + 1] = st.postime;
            }
            k++;
            for (int j = 0; j < 4; j++) a[j] = b[j];
            ta = st.postime;
        }
    }
    if (COUNT) seg_count[i] = k;
}

__global__ void __launch_bounds__(128) hist_reduce_kernel(long long n, const long long *__restrict__ seg_off, const unsigned char *__restrict__ hitv,
                                                          const double *__restrict__ toiv, const unsigned char *__restrict__ stagev,
                                                          const double *__restrict__ vtime, unsigned char *__restrict__ hit, double *__restrict__ toi,
                                                          unsigned char *__restrict__ stage, unsigned long long *earliest_bits, unsigned long long *nhit)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int sg = 0;
    double t = 0.0;
    if (i < n)
    {
        for (long long k = seg_off[i]; k < seg_off[i + 1]; k++)
            if (hitv[k])
            {
                const double ta = vtime[2 * k], tb = vtime[2 * k + 1];
                t = ta + toiv[k] * (tb - ta);      // the primitives return the parameter inside the stitched segment
                sg = stagev[k];
                break;
            }
        hit[i] = sg != 0;
        toi[i] = sg ? t : 0.0;
        if (stage) stage[i] = (unsigned char)sg;
    }
    reduce_block(sg != 0, t, earliest_bits, nhit);
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

// Diagnostics: CCD_NP_TRACE=1 prints the device time of each kernel group of the single-step pipeline to stderr
// (CUDA events on the launching stream; adds a synchronisation per call, so never set it for measurements of the step).
struct NpTrace
{
    bool on = false;
    int n = 0;
    cudaEvent_t ev[40];
    const char *name[40];
    NpTrace()
    {
        const char *e = getenv("CCD_NP_TRACE");
        on = e && e[0] == '1';
        if (on) for (int i = 0; i < 40; i++) cudaEventCreate(&ev[i]);
    }
    void mark(cudaStream_t st, const char *what)
    {
        if (!on || n >= 40) return;
        name[n] = what;
        cudaEventRecord(ev[n++], st);
    }
    void flush(const char *title, const unsigned long long *ctr = nullptr, long long nst = 0)
    {
        if (!on || n == 0) return;
        cudaEventSynchronize(ev[n - 1]);
        if (ctr)
        {
            unsigned long long h[CCD_NP_COUNTERS];
            cudaMemcpy(h, ctr, sizeof(h), cudaMemcpyDeviceToHost);
            fprintf(stderr, "[np counts %s] stencils %lld", title, nst);
            for (int i = 0; i < CCD_NP_COUNTERS; i++) fprintf(stderr, " %llu", h[i]);
            fprintf(stderr, "\n");
        }
        fprintf(stderr, "[np trace %s]", title);
        for (int i = 1; i < n; i++)
        {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
            fprintf(stderr, " %s %.3f", name[i], ms);
        }
        fprintf(stderr, "\n");
        n = 0;
    }
};
static NpTrace g_trace;

template <int D> static void launch_solve(cudaStream_t st, double *tasks, const int *list, const unsigned long long *count, unsigned long long *cursor)
{
    prepare_kernel<D><<<148 * 8, 128, 0, st>>>(tasks, list, count);
    solve_kernel<D><<<148 * SOLVE_MINB(D), 128, 0, st>>>(tasks, list, count, cursor);
    finalize_kernel<D><<<148 * 8, 128, 0, st>>>(tasks, list, count);
}

// sync / sync_role: two runs of one step on two streams that share vertex-edge tests.  Role 1 (the run whose table is consulted)
// records sync[0] once its hash set is complete, sync[1] once the results of its unique tests are known, sync[2] once their
// records are final; role 2 (the run that consults) waits for them right before it needs them.  Role 0: nothing.
// skip_mask: bit 4 * phase + (degree - 3) set = the root-isolation kernels of that degree and phase are not launched (the caller
// saw no such record in the same run of its previous, equally sized call).  Should records of that kind turn up after all
// they stay pending, the combine step sees a record that is not final and hands the stencil to the general routine: slower,
// same bits.
template <bool IS_VF> static int launch_single_step(cudaStream_t st, const P1Args &Q, long long n, int *tlists, cudaStream_t side, cudaEvent_t ev_fork,
                                                    cudaEvent_t ev_join, cudaEvent_t *sync, int sync_role, unsigned skip_mask)
{
    const NpArgs &A = Q.A;
    const int B = 128;
    const unsigned gq = 148 * 16;
    g_trace.mark(st, "start");
    np_cull_kernel<IS_VF><<<grid_for(n, 256), 256, 0, st>>>(Q);
    g_trace.mark(st, "cull");
    int nl = 1;
    if (IS_VF)
    {
        np_stage_kernel<true, 0, false><<<gq, B, 0, st>>>(Q, 0, 0);
        g_trace.mark(st, "stage0");
        np_stage_kernel<true, 1, false><<<gq, B, 0, st>>>(Q, 0, 1);
        np_stage_kernel<true, 2, false><<<gq, B, 0, st>>>(Q, 1, 0);
        np_stage_kernel<true, 3, false><<<gq, B, 0, st>>>(Q, 0, 1);
        nl += 4;
    }
    else
    {
        np_stage_kernel<false, 0, true><<<gq, B, 0, st>>>(Q, 0, 0);      // a0, a1
        g_trace.mark(st, "stage0");
        np_stage_kernel<false, 2, true><<<gq, B, 0, st>>>(Q, 0, 1);      // b0, b1
        np_stage_kernel<false, 4, false><<<gq, B, 0, st>>>(Q, 1, 0);     // distance sextic (last: writes its own record)
        np_export_kernel<false, 3><<<gq, B, 0, st>>>(Q);
        nl += 4;
    }
    g_trace.mark(st, "stages");
    np_export_kernel<IS_VF, 0><<<gq, B, 0, st>>>(Q);
    np_export_kernel<IS_VF, 1><<<gq, B, 0, st>>>(Q);
    np_export_kernel<IS_VF, 2><<<gq, B, 0, st>>>(Q);
    g_trace.mark(st, "export");
    if (sync_role == 2) cudaStreamWaitEvent(st, sync[0], 0);
    np_ve_key_kernel<IS_VF><<<148 * 8, 256, 0, st>>>(Q);
    if (sync_role == 1) cudaEventRecord(sync[0], st);
    g_trace.mark(st, "ve_key");
    np_ve_kernel<IS_VF><<<gq, B, 0, st>>>(Q);
    np_ve_rec_kernel<IS_VF><<<gq, B, 0, st>>>(Q);
    if (sync_role == 1) cudaEventRecord(sync[1], st);
    if (sync_role == 2) cudaStreamWaitEvent(st, sync[1], 0);
    np_ve_resolve_kernel<IS_VF><<<148 * 8, 256, 0, st>>>(Q);
    g_trace.mark(st, "ve");
    np_vv_kernel<IS_VF><<<gq, 256, 0, st>>>(Q);
    np_decide_kernel<IS_VF><<<148 * 8, 256, 0, st>>>(Q);
    g_trace.mark(st, "vv+decide");
    for (int phase = 0; phase < 2; phase++)
    {
        unsigned long long *nd = Q.ctr + (phase ? K_NDEG2 : K_NDEG), *cu = Q.ctr + (phase ? K_CURSOR2 : K_CURSOR);
        bucket_tasks_kernel<<<148 * 4, 256, 0, st>>>(A.rtag, A.ntask, A.task_cap, tlists, nd, phase);
        g_trace.mark(st, "bucket");
        if (!((skip_mask >> (4 * phase + 0)) & 1u)) launch_solve<3>(st, A.tasks, tlists + 0 * A.task_cap, nd + 0, cu + 0);
        g_trace.mark(st, "solve3");
        if (!((skip_mask >> (4 * phase + 1)) & 1u)) launch_solve<4>(st, A.tasks, tlists + 1 * A.task_cap, nd + 1, cu + 1);
        g_trace.mark(st, "solve4");
        if (!((skip_mask >> (4 * phase + 2)) & 1u)) launch_solve<5>(st, A.tasks, tlists + 2 * A.task_cap, nd + 2, cu + 2);
        g_trace.mark(st, "solve5");
        if (!((skip_mask >> (4 * phase + 3)) & 1u)) launch_solve<6>(st, A.tasks, tlists + 3 * A.task_cap, nd + 3, cu + 3);
        g_trace.mark(st, "solve6");
        if (phase == 0 && sync_role == 1) cudaEventRecord(sync[2], st);      // the distance polynomials (vertex-edge quartics among them) are final
        if (phase == 0) { np_window_kernel<IS_VF><<<gq, B, 0, st>>>(Q); g_trace.mark(st, "window"); }
    }
    if (sync_role == 2) cudaStreamWaitEvent(st, sync[2], 0);
    np_combine_kernel<IS_VF><<<gq, B, 0, st>>>(Q);
    g_trace.mark(st, "combine");
    if (side && !g_trace.on)
    {
        // the few stencils of the general routine are long single-lane walks (0.3 ms for ~300 stencils at the 4M-triangle
        // cloth): they run beside whatever the caller launches next and are joined through ev_join
        cudaEventRecord(ev_fork, st);
        cudaStreamWaitEvent(side, ev_fork, 0);
        np_general_kernel<IS_VF><<<148 * 4, B, 0, side>>>(Q);
        cudaEventRecord(ev_join, side);
    }
    else
        np_general_kernel<IS_VF><<<148 * 4, B, 0, st>>>(Q);
    g_trace.mark(st, "general");
    g_trace.flush(IS_VF ? "VF" : "EE", Q.ctr, n);
    return nl + 3 + 6 + 2 * 13 + 1 + 2;
}

// Single step: q0 = positions packed by ccdk_pack_positions (q1, vstride unused); multi-entry History: q0 == nullptr.
// Scratch (sizes in elements, n = number of stencils): work list {w_stencil: n ints, w_meta: n ints, w_base: 5n ints},
// tasks (task_cap records of REC_STRIDE doubles, task_cap < 2^28), tlists (6 x task_cap ints: pending records by degree, refined list, record tags),
// status (n u32), sbase (5n ints), queues (13n ints: primitive / vertex-edge / vertex-vertex items / vertex-edge items that need records), sq (2 x n int2: stage
// queues), xq (5 x n int2: export queues), ctr (CCD_NP_COUNTERS counters, zeroed here; ctr[0] = work-list entries,
// ctr[1] = records; tlists region 4 holds the general routine's list), ve_scratch (12 bytes x ve_slots + 48 bytes x n; ve_slots a power of two, see ccdk_np_ve_slots).  Returns the number of kernels launched.  If ctr[1] ends above task_cap the caller must grow the
// record buffer and call again.  side != nullptr: the general routine is launched on that stream (after ev_fork) and
// signals ev_join; the caller must make its stream wait for ev_join before it reads the results.
// prev_*: the vertex-edge scratch, stencil count and record buffer of the run just before this one when both belong to the
// same step and use the same eta (nullptr otherwise): tests already made there are looked up instead of repeated.
int ccdk_narrowphase(cudaStream_t st, bool is_vf, long long n, const int *stencils, const double *eta_arr, double eta_all,
                     const double *q0, const double *q1, int vstride, const float *vbox, const long long *hoff, const double *htime, const double *hpos,
                     unsigned char *hit, double *toi, unsigned char *stage, unsigned long long *earliest_bits,
                     unsigned long long *nhit, int *w_stencil, int *w_meta, int *w_base, double *tasks, int *tlists,
                     unsigned long long task_cap, unsigned *status, int *sbase, int *queues, int *sq, int *xq, unsigned long long *ctr,
                     void *ve_scratch, unsigned ve_slots, int V, cudaStream_t side, cudaEvent_t ev_fork, cudaEvent_t ev_join,
                     const void *prev_ve_scratch, unsigned prev_ve_slots, long long prev_n, const double *prev_tasks, cudaEvent_t *sync, int sync_role, unsigned skip_mask)
{
    static_assert(K_COUNT <= CCD_NP_COUNTERS, "counter block too small");
    if (n <= 0) return 0;
    NpArgs A;
    A.n = n; A.stencils = stencils; A.eta_arr = eta_arr; A.eta_all = eta_all; A.q0 = q0; A.q1 = q1; A.vstride = vstride; A.vbox = vbox;
    A.hoff = hoff; A.htime = htime; A.hpos = hpos; A.hit = hit; A.toi = toi; A.stage = stage;
    A.earliest_bits = earliest_bits; A.nhit = nhit;
    A.w_stencil = w_stencil; A.w_meta = w_meta; A.w_base = w_base; A.nwork = ctr + K_NWORK;
    A.tasks = tasks; A.tasks_other = nullptr; A.ntask = ctr + K_NTASK; A.task_cap = task_cap; A.rtag = tlists + 5 * task_cap;
    if (q0 == nullptr)
    {
        if (is_vf) stencil_history_kernel<true><<<grid_for(n, 128), 128, 0, st>>>(A);
        else stencil_history_kernel<false><<<grid_for(n, 128), 128, 0, st>>>(A);
        return 1;
    }
    cudaMemsetAsync(ctr, 0, CCD_NP_COUNTERS * sizeof(unsigned long long), st);
    P1Args Q;
    Q.A = A; Q.status = status; Q.sbase = sbase;
    Q.qprim = queues; Q.qve = queues + n; Q.qvv = queues + 5 * n; Q.qve2 = queues + 9 * n;
    Q.sq[0] = reinterpret_cast<int2 *>(sq); Q.sq[1] = reinterpret_cast<int2 *>(sq) + n;
    for (int k = 0; k < 5; k++) Q.xq[k] = reinterpret_cast<int2 *>(xq) + (size_t)k * n;
    Q.qgen = tlists + 4 * task_cap;      // a list of its own: the general routine may still be running when the next run reuses the queues
    Q.ctr = ctr;
    // vertex-edge de-duplication scratch: ve_slots keys (8 bytes) + ve_slots slot values + 3 x 4n ints
    Q.vslots = ve_slots;
    Q.vkeys = reinterpret_cast<unsigned long long *>(ve_scratch);
    Q.vslotval = reinterpret_cast<int *>(Q.vkeys + ve_slots);
    Q.vitem = Q.vslotval + ve_slots;
    Q.vulist = Q.vitem + 4 * n;
    Q.vures = Q.vulist + 4 * n;
    Q.vdedup = eta_arr == nullptr && V <= (1 << 21) && ve_slots >= 2;
    if (Q.vdedup) cudaMemsetAsync(Q.vkeys, 0xff, sizeof(unsigned long long) * ve_slots, st);
    Q.pslots = 0; Q.pkeys = nullptr; Q.pslotval = nullptr; Q.pures = nullptr;
    Q.A.tasks_other = prev_tasks;
    if (Q.vdedup && prev_ve_scratch && prev_ve_slots >= 2 && prev_tasks)
    {
        Q.pslots = prev_ve_slots;
        Q.pkeys = reinterpret_cast<const unsigned long long *>(prev_ve_scratch);
        Q.pslotval = reinterpret_cast<const int *>(Q.pkeys + prev_ve_slots);
        Q.pures = Q.pslotval + prev_ve_slots + 8 * prev_n;      // behind that run's vitem and vulist (4 n ints each)
    }
    return is_vf ? launch_single_step<true>(st, Q, n, tlists, side, ev_fork, ev_join, sync, sync_role, skip_mask) : launch_single_step<false>(st, Q, n, tlists, side, ev_fork, ev_join, sync, sync_role, skip_mask);
}

// hash-set size for the vertex-edge de-duplication of a run over n stencils: the unique tests number ~0.15 n (4M-triangle
// cloth), at most 4 n; a set that fills up only loses de-duplication (np_ve_key_kernel), never correctness
unsigned ccdk_np_ve_slots(long long n)
{
    unsigned s = 1u << 12;
    while ((long long)s < n && s < (1u << 30)) s <<= 1;
    return s;
}

void ccdk_sepplane(cudaStream_t st, bool is_vf, long long n, const int *stencils, const double *eta_arr, const long long *hoff, const double *htime,
                   const double *hpos, double eps, unsigned char *hit, unsigned long long *nhit, unsigned long long *err)
{
    if (n <= 0) return;
    HistView H;
    H.hoff = hoff; H.htime = htime; H.hpos = hpos;
    if (is_vf) sepplane_kernel<true><<<grid_for(n, 128), 128, 0, st>>>(n, stencils, eta_arr, H, eps, hit, nhit, err);
    else sepplane_kernel<false><<<grid_for(n, 128), 128, 0, st>>>(n, stencils, eta_arr, H, eps, hit, nhit, err);
}

void ccdk_hist_count(cudaStream_t st, long long n, const int *stencils, const long long *hoff, const double *htime, const double *hpos, int *seg_count)
{
    if (n > 0) hist_expand_kernel<true><<<grid_for(n, 128), 128, 0, st>>>(n, stencils, nullptr, 0.0, hoff, htime, hpos, seg_count, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
}

void ccdk_hist_fill(cudaStream_t st, long long n, const int *stencils, const double *eta_arr, double eta_all, const long long *hoff, const double *htime,
                    const double *hpos, const long long *seg_off, long long vbase, double *q0v, double *q1v, int *vst, double *veta, double *vtime)
{
    if (n > 0) hist_expand_kernel<false><<<grid_for(n, 128), 128, 0, st>>>(n, stencils, eta_arr, eta_all, hoff, htime, hpos, nullptr, seg_off, vbase, q0v, q1v, vst, veta, vtime);
}

void ccdk_hist_reduce(cudaStream_t st, long long n, const long long *seg_off, const unsigned char *hitv, const double *toiv, const unsigned char *stagev,
                      const double *vtime, unsigned char *hit, double *toi, unsigned char *stage, unsigned long long *earliest_bits, unsigned long long *nhit)
{
    if (n > 0) hist_reduce_kernel<<<grid_for(n, 128), 128, 0, st>>>(n, seg_off, hitv, toiv, stagev, vtime, hit, toi, stage, earliest_bits, nhit);
}

// one round of the staged SeparatingPlane search (first: round 0 over the stencils, vertex-face then edge-edge); see SpRoundArgs
void ccdk_sp_round(cudaStream_t st, bool first, long long nvf, long long nee, const int *vf, const int *ee, const double *vf_eta, const double *ee_eta,
                   const long long *hoff, const double *htime, const double *hpos, double eps, const int *in_st, const double *in_lo, const double *in_hi,
                   const unsigned long long *nin, int *out_st, double *out_lo, double *out_hi, unsigned long long *nout, unsigned long long qcap,
                   unsigned char *hit_vf, unsigned char *hit_ee, unsigned long long *nleaf, unsigned long long leaf_cap_vf, unsigned long long leaf_cap_ee,
                   double *q0v, double *q1v, int *vst, double *veta, int *leaf_stencil)
{
    SpRoundArgs G;
    G.H.hoff = hoff; G.H.htime = htime; G.H.hpos = hpos;
    G.stencils[0] = vf; G.stencils[1] = ee; G.eta_arr[0] = vf_eta; G.eta_arr[1] = ee_eta; G.eps = eps; G.n[0] = nvf; G.n[1] = nee;
    G.in_st = in_st; G.in_lo = in_lo; G.in_hi = in_hi; G.nin = nin;
    G.out_st = out_st; G.out_lo = out_lo; G.out_hi = out_hi; G.nout = nout; G.qcap = qcap; G.hit[0] = hit_vf; G.hit[1] = hit_ee; G.nleaf = nleaf;
    G.leaf_cap[0] = leaf_cap_vf; G.leaf_cap[1] = leaf_cap_ee; G.vbase[0] = 0; G.vbase[1] = (long long)leaf_cap_vf;
    G.q0v = q0v; G.q1v = q1v; G.vst = vst; G.veta = veta; G.leaf_stencil = leaf_stencil;
    if (first) sp_round_kernel<true><<<grid_for(nvf + nee > 0 ? nvf + nee : 1, 128), 128, 0, st>>>(G);
    else
    {
        // the queue of a later round is rarely longer than the stencil list: no more blocks than that needs (a small mesh runs ~60
        // rounds of a few thousand intervals each, and every block pays the round's barriers)
        const unsigned want = grid_for(nvf + nee > 0 ? nvf + nee : 1, 128);
        sp_round_kernel<false><<<want < 148u * 4u ? want : 148u * 4u, 128, 0, st>>>(G);
    }
}

void ccdk_sp_leaf_direct(cudaStream_t st, bool is_vf, long long nleaf, long long vbase, const double *q0v, const double *q1v, const double *veta,
                         const int *leaf_stencil, unsigned char *hit)
{
    if (nleaf <= 0) return;
    if (is_vf) sp_leaf_direct_kernel<true><<<grid_for(nleaf, 128), 128, 0, st>>>(nleaf, vbase, q0v, q1v, veta, leaf_stencil, hit);
    else sp_leaf_direct_kernel<false><<<grid_for(nleaf, 128), 128, 0, st>>>(nleaf, vbase, q0v, q1v, veta, leaf_stencil, hit);
}

void ccdk_sp_leaf_or(cudaStream_t st, long long nleaf, const int *leaf_stencil, const unsigned char *hitv, unsigned char *hit)
{
    if (nleaf > 0) sp_leaf_or_kernel<<<grid_for(nleaf, 256), 256, 0, st>>>(nleaf, leaf_stencil, hitv, hit);
}

void ccdk_count_flags(cudaStream_t st, long long n, const unsigned char *flag, unsigned long long *count)
{
    if (n > 0) count_flags_kernel<<<grid_for(n, 256), 256, 0, st>>>(n, flag, count);
}

void ccdk_pack_positions(cudaStream_t st, int V, const double *q0, const double *q1, int vstride, double *qpack, float *vbox)
{
    if (V > 0) pack_positions_kernel<<<grid_for(V, 256), 256, 0, st>>>(V, q0, q1, vstride, qpack, vbox);
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
