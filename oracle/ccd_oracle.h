/* TEST INFRASTRUCTURE — not product code.
 *
 * Plain-C CPU restatement of the continuous-time collision-detection hot path of
 * evouga/collisiondetection (see ccd_oracle.c for per-function reference citations).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * Parity status: PINNED against the unmodified reference built into oracle/_ref
 * (tests/test_oracle_vs_reference.py) and against tests/golden/ fixtures generated from it.
 * The one deliberate difference is the polynomial root finder: the reference's Jenkins-Traub
 * (src/rpoly.h) is replaced, as BASELINE.json's north_star prescribes, by a real-root isolator
 * on [0,1]; see `orc_roots01`.
 *
 * All entry points mirror oracle/ref_harness.cpp one-to-one (prefix orc_ instead of ref_).
 */
#ifndef CCD_ORACLE_H
#define CCD_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

void orc_free(void *p);

int orc_broadphase(int kind, int V, int F, const int *faces, const long long *hoff, const double *htime,
                   const double *hpos, double outerEta, const unsigned char *fixedMask, int **vf_out,
                   long long *nvf, int **ee_out, long long *nee, double *seconds);

int orc_narrowphase(int which, int V, const long long *hoff, const double *htime, const double *hpos,
                    long long nvf, const int *vf, const double *vf_eta, long long nee, const int *ee,
                    const double *ee_eta, unsigned char *vf_hit, double *vf_toi, int *vf_stage,
                    unsigned char *ee_hit, double *ee_toi, int *ee_stage, double *seconds);

void orc_narrowphase_flat(int V, const long long *hoff, const double *htime, const double *hpos,
                          long long nvf, const int *vf, const double *vf_eta, long long nee, const int *ee,
                          const double *ee_eta, unsigned char *vf_hit, double *vf_toi, unsigned char *ee_hit,
                          double *ee_toi);

void orc_vf_batch(long long n, const double *pts, const double *eta, unsigned char *hit, double *t);
void orc_ee_batch(long long n, const double *pts, const double *eta, unsigned char *hit, double *t);
void orc_ve_batch(long long n, const double *pts, const double *eta, unsigned char *hit, double *t);
void orc_vv_batch(long long n, const double *pts, const double *eta, unsigned char *hit, double *t);

void orc_dist_vf_batch(long long n, const double *pts, double *vec, double *bary);
void orc_dist_ee_batch(long long n, const double *pts, double *vec, double *bary);
void orc_dist_plane_lt_batch(long long n, const double *pts, const double *eta, unsigned char *out);
void orc_dist_line_lt_batch(long long n, const double *pts, const double *eta, unsigned char *out);
double orc_mesh_self_distance(int V, const double *verts, int F, const int *faces,
                              const unsigned char *fixedMask, double *seconds);

/* PenaltyGroup::addForce restated (src/PenaltyGroup.cpp:34-52, src/PenaltyPotential.cpp:7-64); F is in/out */
int orc_penalty_group_force(int V, const double *q, const double *v, long long nvf, const int *vf, const unsigned char *vf_isnew,
                            long long nee, const int *ee, const unsigned char *ee_isnew, double dt, double outerEta, double innerEta,
                            double stiffness, double CoR, double *F, unsigned char *fired, long long *n_fired);

/* Real roots in [0,1] of c[0] t^d + ... + c[d] (c[0] != 0, 3 <= d <= 6), ascending; returns count. */
int orc_roots01(const double *c, int d, double *roots);

/* CTCD::findIntervals restated: returns the number of closed intervals written to l[], u[] (<= 8).
 * op is modified (normalised / shifted) exactly like the reference does. */
int orc_find_intervals(double *op, int n, int pos, double *l, double *u);

/* Leaf boxes: kind 13 or 3; out is F x 2*kind doubles (mins then maxs per face). */
void orc_leaf_boxes(int kind, int F, const int *faces, const long long *hoff, const double *hpos,
                    double outerEta, double *boxes);

/* Polynomial coefficient dumps for the mismatch classifier (start/end points as in *_batch). */
void orc_vf_polys(const double *pts, double eta, double *cubics /*3x4*/, double *sextic /*7*/);
void orc_ee_polys(const double *pts, double eta, double *sextic /*7*/, double *quartics /*4x5*/);

#ifdef __cplusplus
}
#endif
#endif
