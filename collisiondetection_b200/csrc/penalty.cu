// PenaltyGroup::addForce (src/PenaltyGroup.cpp:34-52) over a group's stencil lists, with the per-stencil potentials
// VertexFacePenaltyPotential::addForce / EdgeEdgePenaltyPotential::addForce (src/PenaltyPotential.cpp:7-64).
//
// The reference walks the vertex-face list, then the edge-edge list, and scatters four contributions per fired stencil into
// one `groupforce` vector; F += groupforce * dt at the end.  The sum a vertex receives therefore depends on the ORDER of the
// lists.  Here every stencil computes its four contributions in parallel (penalty_stencil_kernel), the (vertex, item) pairs
// of the fired stencils are sorted by vertex with a STABLE radix sort (items keep list order inside a vertex), and one thread
// per vertex run adds its contributions in that order (penalty_gather_kernel) — bit-identical to the sequential loop, with
// no floating-point atomics.  Translation unit compiled with --fmad=false like the rest of the library.
#include "ccd_kernels.h"
#include "ccd_math.cuh"
#include "ccd_distance.cuh"
#include <cub/cub.cuh>

namespace ccd {

static __device__ __forceinline__ V3 neg(V3 a) { return mk(-a.x, -a.y, -a.z); }

// one thread per stencil (VF list first, then EE): fired flag, contributions item 4 i + k for the stencil's k-th vertex in
// the order the reference updates F (p, q0, q1, q2 | p0, p1, q0, q1); key = vertex id for fired stencils, V (sentinel,
// sorted behind every vertex) otherwise
__global__ void __launch_bounds__(128) penalty_stencil_kernel(long long nvf, long long nee, const int *__restrict__ vf, const int *__restrict__ ee,
                                                              const double *__restrict__ q, const double *__restrict__ v, int V, double outerEta,
                                                              double innerEta, double stiffness, double CoR, double *__restrict__ contrib,
                                                              unsigned *__restrict__ keys, unsigned *__restrict__ items, unsigned char *__restrict__ fired_out,
                                                              const unsigned char *__restrict__ vf_isnew, const unsigned char *__restrict__ ee_isnew,
                                                              unsigned long long *__restrict__ ctr)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvf + nee) return;
    const bool is_vf = i < nvf;
    const int4 s = is_vf ? reinterpret_cast<const int4 *>(vf)[i] : reinterpret_cast<const int4 *>(ee)[i - nvf];
    const int id[4] = {s.x, s.y, s.z, s.w};
    const V3 a = ldv(q + 3 * (size_t)id[0]), b = ldv(q + 3 * (size_t)id[1]), c = ldv(q + 3 * (size_t)id[2]), d = ldv(q + 3 * (size_t)id[3]);
    bool fired = false;
    double w[4] = {0, 0, 0, 0};
    V3 localF = mk(0, 0, 0);
    if (is_vf ? plane_lt(a, b, c, d, outerEta) : line_lt(a, b, c, d, outerEta))
    {
        V3 cv;
        if (is_vf)
        {
            double b0, b1, b2;
            cv = dist_vf(a, b, c, d, b0, b1, b2);
            w[1] = b0; w[2] = b1; w[3] = b2;
        }
        else
            cv = dist_ee(a, b, c, d, w[0], w[1], w[2], w[3]);
        const double dist = sqrt(dot(cv, cv));
        if (!(dist >= outerEta || dist < innerEta))
        {
            fired = true;
            const V3 va = ldv(v + 3 * (size_t)id[0]), vb = ldv(v + 3 * (size_t)id[1]), vc = ldv(v + 3 * (size_t)id[2]), vd = ldv(v + 3 * (size_t)id[3]);
            // src/PenaltyPotential.cpp:22 / :53, coefficient-wise, left to right
            V3 relvel;
            if (is_vf)
                relvel = ((neg(va) + w[1] * vb) + w[2] * vc) + w[3] * vd;
            else
                relvel = ((neg(w[0] * va) - w[1] * vb) + w[2] * vc) + w[3] * vd;
            double k = stiffness;
            if (dot(relvel, cv) > 0) k *= CoR;
            const double sc = k * (outerEta - dist) / (outerEta - innerEta);
            const V3 t = sc * cv;
            localF = mk(t.x / dist, t.y / dist, t.z / dist);
        }
    }
    if (fired_out) fired_out[i] = fired;
    V3 f[4];
    if (is_vf) { f[0] = neg(localF); f[1] = w[1] * localF; f[2] = w[2] * localF; f[3] = w[3] * localF; }
    else       { f[0] = neg(w[0] * localF); f[1] = neg(w[1] * localF); f[2] = w[2] * localF; f[3] = w[3] * localF; }
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        const size_t it = 4 * (size_t)i + k;
        keys[it] = fired ? (unsigned)id[k] : (unsigned)V;
        items[it] = (unsigned)it;
        if (fired) { contrib[3 * it] = f[k].x; contrib[3 * it + 1] = f[k].y; contrib[3 * it + 2] = f[k].z; }
    }
    if (fired)
    {
        atomicAdd(ctr, 1ull);
        const bool isnew = is_vf ? (vf_isnew ? vf_isnew[i] != 0 : true) : (ee_isnew ? ee_isnew[i - nvf] != 0 : true);
        if (isnew) atomicOr(reinterpret_cast<unsigned *>(ctr + 1), 1u);
    }
}

// one thread per sorted item; the head of a vertex's run adds the run in order: groupforce[v] = (((0 + c0) + c1) + ...)
__global__ void __launch_bounds__(128) penalty_gather_kernel(long long nitems, int V, const unsigned *__restrict__ keys, const unsigned *__restrict__ items,
                                                             const double *__restrict__ contrib, double *__restrict__ group)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nitems) return;
    const unsigned key = keys[j];
    if (key >= (unsigned)V) return;
    if (j > 0 && keys[j - 1] == key) return;
    double gx = 0.0, gy = 0.0, gz = 0.0;
    for (long long k = j; k < nitems && keys[k] == key; k++)
    {
        const size_t it = items[k];
        gx = gx + contrib[3 * it]; gy = gy + contrib[3 * it + 1]; gz = gz + contrib[3 * it + 2];
    }
    group[3 * (size_t)key] = gx; group[3 * (size_t)key + 1] = gy; group[3 * (size_t)key + 2] = gz;
}

// F += groupforce * dt over every coordinate (src/PenaltyGroup.cpp:50), untouched vertices included (-0 + 0 = +0)
__global__ void penalty_axpy_kernel(long long n, const double *__restrict__ group, double dt, double *__restrict__ F)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) F[i] = F[i] + group[i] * dt;
}

} // namespace ccd

using namespace ccd;

size_t ccdk_penalty_temp_bytes(long long nitems)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned *)nullptr, (unsigned *)nullptr, (const unsigned *)nullptr, (unsigned *)nullptr, (int)nitems);
    return bytes;
}

// Device pointers throughout.  contrib: 12 doubles per stencil; keysA/B, itemsA/B: 4 unsigned per stencil each; group: 3 V
// doubles; ctr: two counters (fired stencils, "a new stencil fired" bit), zeroed here.  Returns the number of launches.
int ccdk_penalty_group_force(cudaStream_t st, int V, const double *q, const double *v, long long nvf, const int *vf, const unsigned char *vf_isnew,
                             long long nee, const int *ee, const unsigned char *ee_isnew, double dt, double outerEta, double innerEta, double stiffness,
                             double CoR, double *F, unsigned char *fired, double *contrib, unsigned *keysA, unsigned *keysB, unsigned *itemsA,
                             unsigned *itemsB, void *temp, size_t temp_bytes, double *group, unsigned long long *ctr)
{
    const long long n = nvf + nee, nitems = 4 * n;
    int launches = 0;
    cudaMemsetAsync(ctr, 0, 2 * sizeof(unsigned long long), st);
    cudaMemsetAsync(group, 0, sizeof(double) * 3 * (size_t)V, st);
    if (n > 0)
    {
        penalty_stencil_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(nvf, nee, vf, ee, q, v, V, outerEta, innerEta, stiffness, CoR, contrib, keysA,
                                                                            itemsA, fired, vf_isnew, ee_isnew, ctr);
        int bits = 1;
        while (bits < 32 && (1ll << bits) <= (long long)V) bits++;      // keys are 0..V
        cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keysA, keysB, itemsA, itemsB, (int)nitems, 0, bits, st);
        penalty_gather_kernel<<<(unsigned)((nitems + 127) / 128), 128, 0, st>>>(nitems, V, keysB, itemsB, contrib, group);
        launches += 2 + 3;
    }
    if (V > 0)
    {
        penalty_axpy_kernel<<<(unsigned)((3ll * V + 255) / 256), 256, 0, st>>>(3ll * V, group, dt, F);
        launches++;
    }
    return launches;
}
