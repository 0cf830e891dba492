// Distance.h queries as batched FP64 kernels and the reductions of Distance::meshSelfDistance
// (include/Distance.h:14-174, src/Distance.cpp:12-66).  Operation order follows the reference so the
// results are bit-identical to its CPU build (translation unit compiled with --fmad=false).
#include "ccd_kernels.h"
#include "ccd_math.cuh"
#include "ccd_distance.cuh"

namespace ccd {

__global__ void dist_batch_kernel(int which, long long n, const double *__restrict__ pts, const double *__restrict__ eta,
                                  double *__restrict__ vec, double *__restrict__ bary, unsigned char *__restrict__ flag)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *p = pts + 12 * i;
    V3 a = ldv(p), b = ldv(p + 3), c = ldv(p + 6), d = ldv(p + 9);
    if (which == 0)
    {
        double b0, b1, b2;
        V3 r = dist_vf(a, b, c, d, b0, b1, b2);
        vec[3 * i] = r.x; vec[3 * i + 1] = r.y; vec[3 * i + 2] = r.z;
        bary[3 * i] = b0; bary[3 * i + 1] = b1; bary[3 * i + 2] = b2;
    }
    else if (which == 1)
    {
        double b0, b1, b2, b3;
        V3 r = dist_ee(a, b, c, d, b0, b1, b2, b3);
        vec[3 * i] = r.x; vec[3 * i + 1] = r.y; vec[3 * i + 2] = r.z;
        bary[4 * i] = b0; bary[4 * i + 1] = b1; bary[4 * i + 2] = b2; bary[4 * i + 3] = b3;
    }
    else if (which == 2)
        flag[i] = plane_lt(a, b, c, d, eta[i]);
    else
        flag[i] = line_lt(a, b, c, d, eta[i]);
}

// block min of non-negative doubles through their bit patterns -> one atomicMin per block
__device__ __forceinline__ void block_min_bits(double v, unsigned long long *out)
{
    unsigned long long bits = (unsigned long long)__double_as_longlong(v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        unsigned long long ob = __shfl_xor_sync(0xffffffffu, bits, o);
        bits = ob < bits ? ob : bits;
    }
    __shared__ unsigned long long s[32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s[w] = bits;
    __syncthreads();
    if (w == 0)
    {
        int nw = (blockDim.x + 31) >> 5;
        bits = lane < nw ? s[lane] : 0x7FF0000000000000ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            unsigned long long ob = __shfl_xor_sync(0xffffffffu, bits, o);
            bits = ob < bits ? ob : bits;
        }
        if (lane == 0) atomicMin(out, bits);
    }
}

// src/Distance.cpp:21-35: min |v_i - v_k|^2 over vertices k of faces that do not contain i.
// One thread per vertex, faces streamed through shared-memory tiles.
#define VMD_TILE 256
__global__ void __launch_bounds__(VMD_TILE) vertex_min_dist2_kernel(int V, int F, const double *__restrict__ verts,
                                                                    const int *__restrict__ faces, unsigned long long *out)
{
    __shared__ int s_idx[VMD_TILE * 3];
    __shared__ double s_pos[VMD_TILE * 9];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    V3 me = mk(0, 0, 0);
    if (i < V) me = ldv(verts + 3ll * i);
    double best = INFINITY;
    for (int base = 0; base < F; base += VMD_TILE)
    {
        int nt = min(VMD_TILE, F - base);
        __syncthreads();
        if ((int)threadIdx.x < nt)
        {
            int f = base + threadIdx.x;
#pragma unroll
            for (int k = 0; k < 3; k++)
            {
                int v = faces[3 * f + k];
                s_idx[3 * threadIdx.x + k] = v;
                s_pos[9 * threadIdx.x + 3 * k] = verts[3ll * v];
                s_pos[9 * threadIdx.x + 3 * k + 1] = verts[3ll * v + 1];
                s_pos[9 * threadIdx.x + 3 * k + 2] = verts[3ll * v + 2];
            }
        }
        __syncthreads();
        if (i < V)
            for (int t = 0; t < nt; t++)
            {
                if (s_idx[3 * t] == i || s_idx[3 * t + 1] == i || s_idx[3 * t + 2] == i) continue;   // Mesh::vertexOfFace
#pragma unroll
                for (int k = 0; k < 3; k++)
                {
                    V3 d = me - mk(s_pos[9 * t + 3 * k], s_pos[9 * t + 3 * k + 1], s_pos[9 * t + 3 * k + 2]);
                    double dist = dot(d, d);
                    if (dist < best) best = dist;
                }
            }
    }
    block_min_bits(best, out);
}

// src/Distance.cpp:49-64: min closest-point distance over candidate stencils
template <bool IS_VF>
__global__ void __launch_bounds__(256) stencil_min_dist_kernel(long long n, const int *__restrict__ stencils, const double *__restrict__ verts,
                                                               unsigned long long *out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double best = INFINITY;
    if (i < n)
    {
        const int4 s = reinterpret_cast<const int4 *>(stencils)[i];
        V3 a = ldv(verts + 3ll * s.x), b = ldv(verts + 3ll * s.y), c = ldv(verts + 3ll * s.z), d = ldv(verts + 3ll * s.w);
        double t0, t1, t2, t3;
        V3 r = IS_VF ? dist_vf(a, b, c, d, t0, t1, t2) : dist_ee(a, b, c, d, t0, t1, t2, t3);
        best = sqrt(dot(r, r));
    }
    block_min_bits(best, out);
}

// FP64 pipe roofline probe: 16 independent DFMA chains per thread, 148*8 blocks of 256 threads
__global__ void __launch_bounds__(256) fp64_peak_kernel(int iters, double *sink)
{
    double a[16];
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = 1.0 + 1e-9 * (threadIdx.x + k);
    const double m = 1.0000001, b = 1e-12;
    for (int i = 0; i < iters; i++)
    {
#pragma unroll
        for (int k = 0; k < 16; k++) a[k] = fma(a[k], m, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) s += a[k];
    sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

} // namespace ccd

using namespace ccd;

void ccdk_fp64_peak(cudaStream_t st, int iters, double *sink) { fp64_peak_kernel<<<148 * 8, 256, 0, st>>>(iters, sink); }

void ccdk_dist_batch(cudaStream_t st, int which, long long n, const double *pts, const double *eta, double *vec, double *bary, unsigned char *flag)
{
    if (n <= 0) return;
    dist_batch_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(which, n, pts, eta, vec, bary, flag);
}
void ccdk_vertex_min_dist2(cudaStream_t st, int V, int F, const double *verts, const int *faces, unsigned long long *out_bits)
{
    if (V <= 0) return;
    vertex_min_dist2_kernel<<<(unsigned)((V + VMD_TILE - 1) / VMD_TILE), VMD_TILE, 0, st>>>(V, F, verts, faces, out_bits);
}
void ccdk_stencil_min_dist(cudaStream_t st, bool is_vf, long long n, const int *stencils, const double *verts, unsigned long long *out_bits)
{
    if (n <= 0) return;
    if (is_vf) stencil_min_dist_kernel<true><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, stencils, verts, out_bits);
    else stencil_min_dist_kernel<false><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, stencils, verts, out_bits);
}
