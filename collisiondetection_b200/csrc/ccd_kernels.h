// Internal launcher prototypes shared between the .cu translation units and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// narrowphase.cu
int ccdk_narrowphase(cudaStream_t st, bool is_vf, long long n, const int *stencils, const double *eta_arr, double eta_all,
                     const double *q0, const double *q1, int vstride, const float *vbox, const long long *hoff, const double *htime, const double *hpos,
                     unsigned char *hit, double *toi, unsigned char *stage, unsigned long long *earliest_bits,
                     unsigned long long *nhit, int *w_stencil, int *w_meta, int *w_base, double *tasks, int *tlists,
                     unsigned long long task_cap, unsigned *status, int *sbase, int *queues, int *sq, int *xq, unsigned long long *ctr,
                     void *ve_scratch, unsigned ve_slots, int V, cudaStream_t side, cudaEvent_t ev_fork, cudaEvent_t ev_join,
                     const void *prev_ve_scratch, unsigned prev_ve_slots, long long prev_n, const double *prev_tasks, cudaEvent_t *sync, int sync_role, unsigned skip_mask);
unsigned ccdk_np_ve_slots(long long n);
// multi-entry History, staged: (stencil, stitched segment) pairs -> virtual single-step stencils -> per-stencil first hit
void ccdk_hist_count(cudaStream_t st, long long n, const int *stencils, const long long *hoff, const double *htime, const double *hpos, int *seg_count);
void ccdk_hist_fill(cudaStream_t st, long long n, const int *stencils, const double *eta_arr, double eta_all, const long long *hoff, const double *htime,
                    const double *hpos, const long long *seg_off, long long vbase, double *q0v, double *q1v, int *vst, double *veta, double *vtime);
void ccdk_hist_reduce(cudaStream_t st, long long n, const long long *seg_off, const unsigned char *hitv, const double *toiv, const unsigned char *stagev,
                      const double *vtime, unsigned char *hit, double *toi, unsigned char *stage, unsigned long long *earliest_bits, unsigned long long *nhit);
#define CCD_NP_COUNTERS 40      // counters per narrowphase run (narrowphase.cu: K_*)
#define CCD_NP_KNDEG 2          // pending records of degree 3..6, first root-isolation phase; second phase: CCD_NP_KNDEG2
#define CCD_NP_KNDEG2 23
#define CCD_NP_KVEU 32          // unique vertex-edge tests of the run
// {x0,y0,z0,-,x1,y1,z1,-} per vertex (8 doubles) for the single-step narrowphase kernels
void ccdk_pack_positions(cudaStream_t st, int V, const double *q0, const double *q1, int vstride, double *qpack, float *vbox);
void ccdk_prim_batch(cudaStream_t st, int kind, long long n, const double *pts, const double *eta, unsigned char *hit, double *t);
void ccdk_find_intervals(cudaStream_t st, long long n, int degree, int pos, const double *coeffs, int *cnt, double *lo, double *hi);
// SeparatingPlaneNarrowPhase: hit flags only; *nhit += hits, *err += stencils whose interval stack overflowed
void ccdk_sepplane(cudaStream_t st, bool is_vf, long long n, const int *stencils, const double *eta_arr, const long long *hoff, const double *htime,
                   const double *hpos, double eps, unsigned char *hit, unsigned long long *nhit, unsigned long long *err);
// staged SeparatingPlaneNarrowPhase (narrowphase.cu): interval queue rounds, leaves -> single-step pipeline, OR
void ccdk_sp_round(cudaStream_t st, bool first, long long nvf, long long nee, const int *vf, const int *ee, const double *vf_eta, const double *ee_eta,
                   const long long *hoff, const double *htime, const double *hpos, double eps, const int *in_st, const double *in_lo, const double *in_hi,
                   const unsigned long long *nin, int *out_st, double *out_lo, double *out_hi, unsigned long long *nout, unsigned long long qcap,
                   unsigned char *hit_vf, unsigned char *hit_ee, unsigned long long *nleaf, unsigned long long leaf_cap_vf, unsigned long long leaf_cap_ee,
                   double *q0v, double *q1v, int *vst, double *veta, int *leaf_stencil);
void ccdk_sp_leaf_direct(cudaStream_t st, bool is_vf, long long nleaf, long long vbase, const double *q0v, const double *q1v, const double *veta,
                         const int *leaf_stencil, unsigned char *hit);
void ccdk_sp_leaf_or(cudaStream_t st, long long nleaf, const int *leaf_stencil, const unsigned char *hitv, unsigned char *hit);
void ccdk_count_flags(cudaStream_t st, long long n, const unsigned char *flag, unsigned long long *count);
// penalty.cu: PenaltyGroup::addForce over device arrays (see the file header for the scratch layout)
size_t ccdk_penalty_temp_bytes(long long nitems);
int ccdk_penalty_group_force(cudaStream_t st, int V, const double *q, const double *v, long long nvf, const int *vf, const unsigned char *vf_isnew,
                             long long nee, const int *ee, const unsigned char *ee_isnew, double dt, double outerEta, double innerEta, double stiffness,
                             double CoR, double *F, unsigned char *fired, double *contrib, unsigned *keysA, unsigned *keysB, unsigned *itemsA,
                             unsigned *itemsB, void *temp, size_t temp_bytes, double *group, unsigned long long *ctr);
