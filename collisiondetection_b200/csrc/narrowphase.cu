// CTCD narrowphase and batched primitive kernels (sm_100a, FP64, no tensor cores: nothing in this
// path is a dense contraction).
//
// One thread per candidate stencil runs the reference's per-stencil sequence
// (src/CTCDNarrowPhase.cpp:24-135): the VF (or EE) primitive, then the degenerate vertex-edge and
// vertex-vertex tests, first hit wins; over a multi-entry History the sequence repeats per
// stitched linear segment (src/History.cpp:98-140).  Hit flag, time of impact and the index of
// the sub-test that fired are written per stencil; earliest TOI and hit counts are reduced
// warp -> block -> one atomic per block.
#include "ccd_kernels.h"
#include "ccd_math.cuh"

namespace ccd {

// ---- per-segment stencil tests --------------------------------------------------------------
// a[0..3] start positions, b[0..3] end positions of (p, q0, q1, q2)
__device__ __forceinline__ int vf_stencil_segment(const V3 *a, const V3 *b, double eta, double &t)
{
    if (vertex_face(a[0], a[1], a[2], a[3], b[0], b[1], b[2], b[3], eta, t)) return 1;
    // vertex against the three face edges (1,2),(2,3),(3,1): src/CTCDNarrowPhase.cpp:51-59
    for (int e = 0; e < 3; e++)
    {
        int i1 = 1 + e, i2 = 1 + ((e + 1) % 3);
        if (vertex_edge(a[0], a[i1], a[i2], b[0], b[i1], b[i2], eta, t)) return 2 + e;
    }
    // vertex against the three face vertices: src/CTCDNarrowPhase.cpp:61-69
    for (int v = 0; v < 3; v++)
        if (vertex_vertex(a[0], a[1 + v], b[0], b[1 + v], eta, t)) return 5 + v;
    return 0;
}

// a/b: (p0, p1, q0, q1); edgeEdgeCTCD takes (q0,p0,q1,p1) = (pos0,pos1,pos2,pos3): src/CTCDNarrowPhase.cpp:91
__device__ __forceinline__ int ee_stencil_segment(const V3 *a, const V3 *b, double eta, double &t)
{
    if (edge_edge(a[0], a[1], a[2], a[3], b[0], b[1], b[2], b[3], eta, t)) return 1;
    // src/CTCDNarrowPhase.cpp:99-114
    if (vertex_edge(a[0], a[2], a[3], b[0], b[2], b[3], eta, t)) return 2;
    if (vertex_edge(a[1], a[2], a[3], b[1], b[2], b[3], eta, t)) return 3;
    if (vertex_edge(a[2], a[0], a[1], b[2], b[0], b[1], eta, t)) return 4;
    if (vertex_edge(a[3], a[0], a[1], b[3], b[0], b[1], eta, t)) return 5;
    // src/CTCDNarrowPhase.cpp:117-132
    if (vertex_vertex(a[0], a[2], b[0], b[2], eta, t)) return 6;
    if (vertex_vertex(a[0], a[3], b[0], b[3], eta, t)) return 7;
    if (vertex_vertex(a[1], a[2], b[1], b[2], eta, t)) return 8;
    if (vertex_vertex(a[1], a[3], b[1], b[3], eta, t)) return 9;
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
// TOI >= 0, so the IEEE bit pattern orders like the value: min over unsigned 64-bit.
__device__ __forceinline__ void reduce_block(bool hit, double toi, unsigned long long *earliest_bits, unsigned long long *nhit)
{
    unsigned long long bits = hit ? (unsigned long long)__double_as_longlong(toi) : 0xFFFFFFFFFFFFFFFFull;
    unsigned cnt = hit ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        unsigned long long ob = __shfl_xor_sync(0xffffffffu, bits, o);
        bits = ob < bits ? ob : bits;
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    __shared__ unsigned long long s_bits[32];
    __shared__ unsigned s_cnt[32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { s_bits[w] = bits; s_cnt[w] = cnt; }
    __syncthreads();
    if (w == 0)
    {
        int nw = (blockDim.x + 31) >> 5;
        bits = lane < nw ? s_bits[lane] : 0xFFFFFFFFFFFFFFFFull;
        cnt = lane < nw ? s_cnt[lane] : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            unsigned long long ob = __shfl_xor_sync(0xffffffffu, bits, o);
            bits = ob < bits ? ob : bits;
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        if (lane == 0 && cnt)
        {
            atomicMin(earliest_bits, bits);
            atomicAdd(nhit, (unsigned long long)cnt);
        }
    }
}

// ---- stencil kernels --------------------------------------------------------------------------
// IS_VF: vertex-face stencils (p,q0,q1,q2) else edge-edge (p0,p1,q0,q1).  SINGLE: two entries per
// vertex (q0 -> q1, xyz-interleaved) instead of the CSR history.
template <bool IS_VF, bool SINGLE>
__global__ void __launch_bounds__(128) stencil_kernel(long long n, const int *__restrict__ stencils,
                                                      const double *__restrict__ eta_arr, double eta_all,
                                                      const double *__restrict__ q0, const double *__restrict__ q1, int vstride,
                                                      const long long *__restrict__ hoff, const double *__restrict__ htime,
                                                      const double *__restrict__ hpos, unsigned char *__restrict__ hit_out,
                                                      double *__restrict__ toi_out, unsigned char *__restrict__ stage_out,
                                                      unsigned long long *earliest_bits, unsigned long long *nhit)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int stage = 0;
    double toi = 0.0;
    if (i < n)
    {
        const int4 s = reinterpret_cast<const int4 *>(stencils)[i];
        const double eta = eta_arr ? eta_arr[i] : eta_all;
        V3 a[4], b[4];
        if (SINGLE)
        {
            const long long vs = vstride;
            a[0] = ldv(q0 + vs * s.x); a[1] = ldv(q0 + vs * s.y); a[2] = ldv(q0 + vs * s.z); a[3] = ldv(q0 + vs * s.w);
            b[0] = ldv(q1 + vs * s.x); b[1] = ldv(q1 + vs * s.y); b[2] = ldv(q1 + vs * s.z); b[3] = ldv(q1 + vs * s.w);
            stage = IS_VF ? vf_stencil_segment(a, b, eta, toi) : ee_stencil_segment(a, b, eta, toi);
        }
        else
        {
            int verts[4] = {s.x, s.y, s.z, s.w};
            Stitcher st;
            st.begin(hoff, htime, hpos, verts);
            if (st.next(a))
                while (st.next(b))
                {
                    stage = IS_VF ? vf_stencil_segment(a, b, eta, toi) : ee_stencil_segment(a, b, eta, toi);
                    if (stage) break;
                    for (int k = 0; k < 4; k++) a[k] = b[k];
                }
        }
        hit_out[i] = stage != 0;
        toi_out[i] = stage ? toi : 0.0;
        if (stage_out) stage_out[i] = (unsigned char)stage;
    }
    reduce_block(stage != 0, toi, earliest_bits, nhit);
}

// ---- batched public primitives (include/CTCD.h:36-79): pts = start points then end points ------
template <int KIND> __global__ void __launch_bounds__(128) prim_kernel(long long n, const double *__restrict__ pts,
                                                                       const double *__restrict__ eta,
                                                                       unsigned char *__restrict__ hit, double *__restrict__ t)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double tt = 0;
    bool h;
    if (KIND == 0)
    {
        const double *p = pts + 24 * i;
        h = vertex_face(ldv(p), ldv(p + 3), ldv(p + 6), ldv(p + 9), ldv(p + 12), ldv(p + 15), ldv(p + 18), ldv(p + 21), eta[i], tt);
    }
    else if (KIND == 1)
    {
        const double *p = pts + 24 * i;
        h = edge_edge(ldv(p), ldv(p + 3), ldv(p + 6), ldv(p + 9), ldv(p + 12), ldv(p + 15), ldv(p + 18), ldv(p + 21), eta[i], tt);
    }
    else if (KIND == 2)
    {
        const double *p = pts + 18 * i;
        h = vertex_edge(ldv(p), ldv(p + 3), ldv(p + 6), ldv(p + 9), ldv(p + 12), ldv(p + 15), eta[i], tt);
    }
    else
    {
        const double *p = pts + 12 * i;
        h = vertex_vertex(ldv(p), ldv(p + 3), ldv(p + 6), ldv(p + 9), eta[i], tt);
    }
    hit[i] = h;
    if (h) t[i] = tt;       // t is written only on a hit, like the reference
}

// raw root isolator / interval finder, exposed for parity tests
__global__ void find_intervals_kernel(long long n, int degree, int pos, const double *__restrict__ coeffs, int *__restrict__ cnt,
                                      double *__restrict__ lo, double *__restrict__ hi)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ivals iv;
    iv.n = 0;
    const double *c = coeffs + 7 * i;
    if (degree == 6) { double op[7]; for (int k = 0; k < 7; k++) op[k] = c[k]; find_intervals<6>(op, iv, pos != 0); }
    else if (degree == 4) { double op[5]; for (int k = 0; k < 5; k++) op[k] = c[k]; find_intervals<4>(op, iv, pos != 0); }
    else if (degree == 3) { double op[4]; for (int k = 0; k < 4; k++) op[k] = c[k]; find_intervals<3>(op, iv, pos != 0); }
    else { double op[3]; for (int k = 0; k < 3; k++) op[k] = c[k]; find_intervals<2>(op, iv, pos != 0); }
    cnt[i] = iv.n;
    for (int k = 0; k < iv.n; k++) { lo[7 * i + k] = iv.l[k]; hi[7 * i + k] = iv.u[k]; }
}

} // namespace ccd

// ---- launchers --------------------------------------------------------------------------------
using namespace ccd;

static inline unsigned grid_for(long long n, int block) { return (unsigned)((n + block - 1) / block); }

void ccdk_narrowphase(cudaStream_t st, bool is_vf, long long n, const int *stencils, const double *eta_arr, double eta_all,
                      const double *q0, const double *q1, int vstride, const long long *hoff, const double *htime, const double *hpos,
                      unsigned char *hit, double *toi, unsigned char *stage, unsigned long long *earliest_bits,
                      unsigned long long *nhit)
{
    if (n <= 0) return;
    const int B = 128;
    const bool single = (q0 != nullptr);
    if (is_vf)
    {
        if (single) stencil_kernel<true, true><<<grid_for(n, B), B, 0, st>>>(n, stencils, eta_arr, eta_all, q0, q1, vstride, hoff, htime, hpos, hit, toi, stage, earliest_bits, nhit);
        else stencil_kernel<true, false><<<grid_for(n, B), B, 0, st>>>(n, stencils, eta_arr, eta_all, q0, q1, vstride, hoff, htime, hpos, hit, toi, stage, earliest_bits, nhit);
    }
    else
    {
        if (single) stencil_kernel<false, true><<<grid_for(n, B), B, 0, st>>>(n, stencils, eta_arr, eta_all, q0, q1, vstride, hoff, htime, hpos, hit, toi, stage, earliest_bits, nhit);
        else stencil_kernel<false, false><<<grid_for(n, B), B, 0, st>>>(n, stencils, eta_arr, eta_all, q0, q1, vstride, hoff, htime, hpos, hit, toi, stage, earliest_bits, nhit);
    }
}

void ccdk_prim_batch(cudaStream_t st, int kind, long long n, const double *pts, const double *eta, unsigned char *hit, double *t)
{
    if (n <= 0) return;
    const int B = 128;
    switch (kind)
    {
    case 0: prim_kernel<0><<<grid_for(n, B), B, 0, st>>>(n, pts, eta, hit, t); break;
    case 1: prim_kernel<1><<<grid_for(n, B), B, 0, st>>>(n, pts, eta, hit, t); break;
    case 2: prim_kernel<2><<<grid_for(n, B), B, 0, st>>>(n, pts, eta, hit, t); break;
    default: prim_kernel<3><<<grid_for(n, B), B, 0, st>>>(n, pts, eta, hit, t); break;
    }
}

void ccdk_find_intervals(cudaStream_t st, long long n, int degree, int pos, const double *coeffs, int *cnt, double *lo, double *hi)
{
    if (n <= 0) return;
    find_intervals_kernel<<<grid_for(n, 128), 128, 0, st>>>(n, degree, pos, coeffs, cnt, lo, hi);
}
