/* ccd_b200 — C ABI of the B200-native continuous-time collision-detection hot path.
 *
 * Drop-in boundary for evouga/collisiondetection's CCD step.  Every entry point below replaces one
 * reference interface (cited file:line into the reference tree); INTEGRATION.md shows the C++
 * adapter a maintainer would compile against this header (include/ccd_b200_adapters.hpp).
 *
 * Conventions
 *   - plain pointers and sizes only; all arithmetic is IEEE double, indices are 32-bit int;
 *   - `faces` is 3 contiguous int32 per face (Mesh::faces is a column-major Matrix3Xi, src/Mesh.h:9);
 *   - positions are xyz-interleaved doubles (Mesh::vertices / History positions, src/Mesh.h:8);
 *   - a History (src/History.h:28-42) crosses the boundary as CSR: hoff[V+1] entry offsets, htime[N],
 *     hpos[3N]; the first entry of a vertex has time 0, the last time 1 (src/History.cpp:8-39);
 *   - stencils are 4 x int32 in the reference's canonical form (src/Stencils.h:7-138) and the
 *     candidate arrays come back in std::set iteration order (lexicographic);
 *   - `kind` is the number of box axes: 13 = KDOPBroadPhase, 3 = AABBBroadPhase;
 *   - every function returns 0 on success or a negative CCD_ERR_* code; ccd_last_error() gives text.
 *     There is no CPU fallback: without a CUDA device ccd_create() fails.
 *   - functions taking `const T *d_...` expect DEVICE pointers on the context's GPU; all others take
 *     HOST pointers and perform the host<->device copies themselves.
 */
#ifndef CCD_B200_H
#define CCD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CCD_OK 0
#define CCD_ERR_CUDA (-1)      /* CUDA runtime failure (message in ccd_last_error) */
#define CCD_ERR_ARG (-2)       /* invalid argument */
#define CCD_ERR_NOMEM (-3)     /* host or device allocation failed */
#define CCD_ERR_NODEVICE (-4)  /* no usable CUDA device: there is no CPU fallback */

#define CCD_KDOP 13
#define CCD_AABB 3

typedef struct ccd_context ccd_context;

/* One context per process / GPU: owns its streams and a growable pool of device buffers that is reused
 * across calls (the reference keeps no state between calls; neither do results here).  Every call on a context makes
 * the context's device the calling thread's current CUDA device (cudaSetDevice) and leaves it so.  Not thread-safe:
 * one call at a time per context. */
int ccd_create(ccd_context **ctx, int device);
void ccd_destroy(ccd_context *ctx);
const char *ccd_last_error(const ccd_context *ctx);
void ccd_free_host(void *p);
/* ABI / build identification: "ccd_b200 <version> sm_100a" */
const char *ccd_version(void);

/* ---- BroadPhase::findCollisionCandidates (src/RetrospectiveDetection.h:10-15), implemented by
 * KDOPBroadPhase (src/KDOPBroadPhase.cpp:34-41) and AABBBroadPhase (src/AABBBroadPhase.cpp:9-16).
 * fixedMask: V bytes, non-zero = vertex is in fixedVerts (NULL = none).
 * *vf / *ee receive malloc'ed arrays of 4*n ints (release with ccd_free_host). */
int ccd_broadphase(ccd_context *ctx, int kind, int V, int F, const int32_t *faces, const int64_t *hoff,
                   const double *htime, const double *hpos, double outerEta, const uint8_t *fixedMask,
                   int32_t **vf, int64_t *nvf, int32_t **ee, int64_t *nee);

/* Single-step fast path of the same call: History(q0) + finishHistory(q1) (example/AlecTest.cpp:86-97). */
int ccd_broadphase_step(ccd_context *ctx, int kind, int V, int F, const int32_t *faces, const double *q0,
                        const double *q1, double outerEta, const uint8_t *fixedMask, int32_t **vf, int64_t *nvf,
                        int32_t **ee, int64_t *nee);

/* ---- NarrowPhase::findCollisions (src/RetrospectiveDetection.h:17-23) as implemented by
 * CTCDNarrowPhase (src/CTCDNarrowPhase.cpp:9-135).  Per stencil: hit flag, the time of impact the
 * reference has in hand when it returns true (it discards it; this ABI surfaces it) and the index of
 * the sub-test that fired (1 = VF/EE primitive, 2.. = vertex-edge, then vertex-vertex; 0 = miss).
 * vf_eta / ee_eta: per-stencil thickness (the `.second` of the reference's pairs).
 * The time of impact is in History time: for a multi-entry History the primitive's parameter inside the stitched
 * segment [ta, tb] that hit is mapped back as ta + t (tb - ta) (which is t itself for a single step, ta = 0, tb = 1).
 * Any of the *_toi / *_stage outputs may be NULL. */
typedef struct
{
    int64_t n_vf_hits, n_ee_hits;
    double earliest_toi; /* min TOI over all hits; +inf when nothing hit */
} ccd_np_summary;

int ccd_narrowphase(ccd_context *ctx, int V, const int64_t *hoff, const double *htime, const double *hpos,
                    int64_t nvf, const int32_t *vf, const double *vf_eta, int64_t nee, const int32_t *ee,
                    const double *ee_eta, uint8_t *vf_hit, double *vf_toi, uint8_t *vf_stage, uint8_t *ee_hit,
                    double *ee_toi, uint8_t *ee_stage, ccd_np_summary *summary);

/* NarrowPhase::findCollisions as SeparatingPlaneNarrowPhase (src/SeparatingPlaneNarrowPhase.cpp:11-278), the narrowphase
 * ActiveLayers instantiates (src/ActiveLayers.cpp:28): recursive interval culling with the plane between the closest
 * points at an interval's midpoint, CTCD primitives on the short intervals that remain.  Same inputs as ccd_narrowphase;
 * outputs are the hit flags (the reference returns nothing else).  eps = History::computeMinimumGap() / 4. */
int ccd_narrowphase_sepplane(ccd_context *ctx, int V, const int64_t *hoff, const double *htime, const double *hpos,
                             int64_t nvf, const int32_t *vf, const double *vf_eta, int64_t nee, const int32_t *ee,
                             const double *ee_eta, uint8_t *vf_hit, uint8_t *ee_hit, int64_t *n_vf_hits, int64_t *n_ee_hits);

/* ---- one whole CCD step, the flow of example/AlecTest.cpp:86-111: broadphase with outerEta, every
 * candidate wrapped with thickness `eta`, CTCD narrowphase.  Returns the colliding stencils (sorted)
 * with their TOIs; candidate arrays are returned only when the pointers are non-NULL. */
typedef struct
{
    int64_t n_vf_candidates, n_ee_candidates;
    int64_t n_vf_hits, n_ee_hits;
    double earliest_toi;
    int32_t *vf_hits;   /* 4*n_vf_hits; pinned host memory owned by the context, valid until its next call */
    double *vf_hit_toi; /* n_vf_hits */
    int32_t *ee_hits;
    double *ee_hit_toi;
    float ms_broadphase, ms_narrowphase; /* device time of the two phases (CUDA events) */
} ccd_step_result;

int ccd_step(ccd_context *ctx, int kind, int V, int F, const int32_t *faces, const double *q0, const double *q1,
             double outerEta, double eta, const uint8_t *fixedMask, ccd_step_result *out);
void ccd_step_result_free(ccd_step_result *r);
/* Same with the sharding of ccd_step_device (below): this rank's part of the step, host buffers in and out. */
int ccd_step_shard(ccd_context *ctx, int kind, int V, int F, const int32_t *faces, const double *q0, const double *q1,
                   double outerEta, double eta, const uint8_t *fixedMask, int shard_rank, int shard_world,
                   ccd_step_result *out);

/* Device-resident variant: inputs already in HBM, results stay in HBM (pointers valid until the next
 * call on this context).  shard_rank / shard_world partition the step across GPUs of one job: rank r owns the
 * vertices (its VF stencils) and unique edges (its EE stencils) anchored in a contiguous range of the faces' Morton
 * order, i.e. a subtree range of the step's LBVH.  Every rank builds the (replicated) tree, but intersects it, tests
 * face pairs and counts stencils only around the faces its own range touches.  The ranges are an equal split until
 * the caller installs better ones with ccd_set_shard_partition — normally computed from the load profiles of the
 * previous step (ccd_shard_histogram, summed over ranks).  Whatever the ranges, the ranks' stencil lists are sorted,
 * pairwise disjoint, and their union is the single-GPU list exactly.  world = 1 means the whole step. */
typedef struct
{
    int64_t n_vf_candidates, n_ee_candidates;
    int64_t n_vf_hits, n_ee_hits;
    double earliest_toi;
    const int32_t *d_vf; /* device: 4*n_vf_candidates */
    const int32_t *d_ee;
    const uint8_t *d_vf_hit;
    const uint8_t *d_ee_hit;
    const double *d_vf_toi;
    const double *d_ee_toi;
    int64_t n_face_pairs;      /* overlapping non-neighbour face pairs found by this rank */
    int64_t n_tree_candidates; /* pairs that survived the float AABB tree before the exact test */
    float ms_broadphase, ms_narrowphase;
    int32_t n_launches; /* kernels launched by this call (ours + CUB passes) */
    int64_t n_vf_deferred, n_ee_deferred; /* stencils that needed the iterative root isolator (narrowphase pass 2) */
} ccd_device_result;

int ccd_step_device(ccd_context *ctx, int kind, int V, int F, const int32_t *d_faces, const double *d_q0,
                    const double *d_q1, double outerEta, double eta, const uint8_t *d_fixedMask, int shard_rank,
                    int shard_world, ccd_device_result *out);

/* ccd_step_device with the hit lists copied to (pinned) host memory like ccd_step_shard: device inputs in, host results
 * out — the end-to-end call of a multi-GPU job whose positions reach the GPUs by an all-gather over NVLink. */
int ccd_step_device_hits(ccd_context *ctx, int kind, int V, int F, const int32_t *d_faces, const double *d_q0,
                         const double *d_q1, double outerEta, double eta, const uint8_t *d_fixedMask, int shard_rank,
                         int shard_world, ccd_step_result *out);
/* The context runs its kernels on a private non-blocking stream.  Device inputs (d_faces, d_q0, d_q1, d_fixedMask) that
 * are still being produced on another stream — a copy, an NCCL broadcast / all-gather — must be ordered before the next
 * call: ccd_wait_stream(ctx, stream) makes everything the context enqueues afterwards wait for what is on `stream`
 * (a cudaStream_t; NULL = the legacy default stream) at the time of the call.  No host synchronisation.  Alternatively
 * synchronise the producer stream yourself. */
int ccd_wait_stream(ccd_context *ctx, void *producer_stream);

/* Ownership ranges for sharded steps on this context: pbounds has world+1 ascending entries over the F sorted (Morton)
 * positions of the faces, pbounds[0] = 0, pbounds[world] = F.  Rank r owns the vertices and unique edges whose anchor face
 * (the first face of the vertex's star / of the edge's face list) sorts into [pbounds[r], pbounds[r+1]) — a region of
 * space, whatever the mesh numbering (BASELINE north_star: "pairs are partitioned by BVH subtree").  Ignored (equal
 * split) when the last bound does not match F of a later call. */
int ccd_set_shard_partition(ccd_context *ctx, int world, const int32_t *pbounds);
/* Load profile of the last sharded step on this context: stencils owned by this rank per bucket of sorted positions,
 * CCD_SHARD_BUCKETS equal-width buckets (position p falls in bucket p * CCD_SHARD_BUCKETS / n_positions): vf_hist counts
 * VF stencils by the anchor of their vertex, ee_hist EE stencils by the anchor of their first edge.  Summing over ranks
 * gives the whole step's profile.  n_positions = F. */
#define CCD_SHARD_BUCKETS 1024
int ccd_shard_histogram(ccd_context *ctx, int64_t *vf_hist, int64_t *ee_hist, int32_t *n_positions);

/* Utilities for bindings that hold device results: copy `bytes` from a device pointer of this context to host
 * memory, and a measured FP64 roofline denominator (dependent-free DFMA streams on every SM) in TFLOP/s. */
int ccd_memcpy_d2h(ccd_context *ctx, void *h_dst, const void *d_src, uint64_t bytes);
/* Device time (ms, CUDA events on the context's stream) of the stages of the last ccd_step_device / ccd_step call:
 * 0 topology tables (cached after the first call on a mesh; a hash of the faces is checked every call), 1 face boxes,
 * 2 Morton codes + radix sort + cluster tree, 3 pair traversal + cluster-pair / k-DOP face-pair tests, 4 adjacency CSR,
 * 5 stencil count pass, 6 stencil write pass, 7 and 8 the narrowphase: the edge-edge run (8) and the vertex-face run (7).
 * When both stencil types are present the two runs execute side by side on two streams: 8 then is the time until the
 * edge-edge run has finished and 7 what is left of the vertex-face run after that; their sum is the narrowphase.
 * Zero after any call that was not a complete step.  n must be >= 9. */
#define CCD_N_STAGE_TIMES 9
int ccd_stage_times(ccd_context *ctx, float *ms, int n);
int ccd_fp64_peak(ccd_context *ctx, double *tflops);

/* ---- the four public primitives, batched (include/CTCD.h:36-79, src/CTCD.cpp:259-692).
 * pts holds, per item, the start points then the end points in the reference's argument order:
 *   vf: q0 q1 q2 q3 | q0end..q3end (24 doubles)      ee: q0 p0 q1 p1 | ends (24)
 *   ve: q0 q1 q2 | ends (18)                          vv: q1 q2 | ends (12)
 * t[i] is written only when hit[i] != 0, like the reference's `double &t`. */
int ccd_vf_batch(ccd_context *ctx, int64_t n, const double *pts, const double *eta, uint8_t *hit, double *t);
int ccd_ee_batch(ccd_context *ctx, int64_t n, const double *pts, const double *eta, uint8_t *hit, double *t);
int ccd_ve_batch(ccd_context *ctx, int64_t n, const double *pts, const double *eta, uint8_t *hit, double *t);
int ccd_vv_batch(ccd_context *ctx, int64_t n, const double *pts, const double *eta, uint8_t *hit, double *t);

/* CTCD::findIntervals (src/CTCD.cpp:98-177) batched, for parity tests of the root isolator:
 * coeffs is n x 7 (descending powers, first degree+1 used); degree in {2,3,4,6}; pos as the reference's
 * flag.  cnt[i] intervals are written to lo/hi[7*i ...]. */
int ccd_find_intervals_batch(ccd_context *ctx, int64_t n, int degree, int pos, const double *coeffs, int32_t *cnt,
                             double *lo, double *hi);

/* ---- include/Distance.h:14-174, batched; pts = 4 points xyz per item in the reference's argument order.
 *   dist_vf: p q0 q1 q2 -> vec[3] (closest point minus p), bary[3]
 *   dist_ee: p0 p1 q0 q1 -> vec[3] (c2 - c1), bary[4] */
int ccd_dist_vf_batch(ccd_context *ctx, int64_t n, const double *pts, double *vec, double *bary);
int ccd_dist_ee_batch(ccd_context *ctx, int64_t n, const double *pts, double *vec, double *bary);
int ccd_dist_plane_lt_batch(ccd_context *ctx, int64_t n, const double *pts, const double *eta, uint8_t *out);
int ccd_dist_line_lt_batch(ccd_context *ctx, int64_t n, const double *pts, const double *eta, uint8_t *out);

/* ---- PenaltyGroup::addForce (src/PenaltyGroup.cpp:34-52) with VertexFacePenaltyPotential::addForce and
 * EdgeEdgePenaltyPotential::addForce (src/PenaltyPotential.cpp:7-64) per stencil: F += dt * (sum of the fired stencils'
 * forces), every vertex's sum taken in the order of the lists (vertex-face stencils, then edge-edge stencils) as the
 * reference's sequential loop does, so F is bit-identical to the CPU result.  q, v, F: 3 V doubles (F in/out).
 * vf / ee: canonical stencils (4 ints each); *_isnew (may be NULL = all new): the stencils' `isnew` members;
 * *_fired (may be NULL): whether the potential fired; *n_fired: their number; *newused: the reference's return value
 * (a stencil that is new fired).  The _device form takes device pointers (d_fired: nvf + nee bytes, may be NULL) so that
 * positions, velocities and forces can stay in HBM between the passes of a VelocityFilter-style loop. */
int ccd_penalty_group_force(ccd_context *ctx, int V, const double *q, const double *v, int64_t nvf, const int32_t *vf,
                            const uint8_t *vf_isnew, int64_t nee, const int32_t *ee, const uint8_t *ee_isnew, double dt,
                            double outerEta, double innerEta, double stiffness, double CoR, double *F, uint8_t *vf_fired,
                            uint8_t *ee_fired, int64_t *n_fired, int *newused);
int ccd_penalty_group_force_device(ccd_context *ctx, int V, const double *d_q, const double *d_v, int64_t nvf,
                                   const int32_t *d_vf, const uint8_t *d_vf_isnew, int64_t nee, const int32_t *d_ee,
                                   const uint8_t *d_ee_isnew, double dt, double outerEta, double innerEta, double stiffness,
                                   double CoR, double *d_F, uint8_t *d_fired, int64_t *n_fired, int *newused);

/* Distance::meshSelfDistance (src/Distance.cpp:12-66); n_vf / n_ee (may be NULL) receive the stencil
 * counts the reference prints ("Checking N vertex-face and M edge-edge stencils"). */
int ccd_mesh_self_distance(ccd_context *ctx, int V, const double *verts, int F, const int32_t *faces,
                           const uint8_t *fixedMask, double *distance, int64_t *n_vf, int64_t *n_ee);

#ifdef __cplusplus
}
#endif
#endif /* CCD_B200_H */
