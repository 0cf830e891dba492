// extern "C" layer of ccd_b200 (include/ccd_b200.h): context, device-buffer pool, host<->device
// staging and the orchestration of the broadphase / narrowphase kernels.  No CPU fallback: every
// entry point runs CUDA kernels or fails with an error code.
#include "../../include/ccd_b200.h"
#include "ccd_kernels.h"

#include <cub/cub.cuh>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

// per-stage device timing (CUDA events on the context's stream); see ccd_stage_times()
enum { ST_TOPOLOGY = 0, ST_BOXES, ST_TREE, ST_TRAVERSE_EXACT, ST_ADJACENCY, ST_EMIT_COUNT, ST_EMIT_WRITE, ST_NP_EE, ST_NP_VF, CCD_N_STAGES };

struct DBuf
{
    void *p = nullptr;
    size_t cap = 0;
};

struct ccd_context
{
    int device = 0;
    cudaStream_t st = nullptr;
    cudaStream_t st2 = nullptr;      // side stream: the edge-edge run's general routine overlaps the vertex-face run
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    cudaStream_t st3 = nullptr;      // the vertex-face narrowphase run, beside the edge-edge run (narrowphase_device)
    cudaEvent_t evPack = nullptr, evVf = nullptr;
    cudaEvent_t evVe[3] = {nullptr, nullptr, nullptr};      // vertex-edge tests shared by the two concurrent runs (narrowphase.cu: launch_single_step)
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t sev[CCD_N_STAGES + 1];
    bool stage_valid = false;
    std::string err;
    int launches = 0;

    // staging of host inputs
    DBuf faces, q0, q1, hoff, htime, hpos, fixed, vf_in, ee_in, vf_eta, ee_eta, pts, eta;
    // broadphase
    DBuf boxes, faabb, bounds, keysA, keysB, valsA, valsB, temp, facePos, cmark;
    DBuf cand, counters, pairL, pairR, deg, adjOff, cursor, adj, heap, srec, sbox, sfaces, unsure, frontA, frontB, bigV, bigE;
    size_t frontCap = 0, unsureCap = 0;
    DBuf k32A, k32B, scanFlags, scanIds;
    // topology cache (function of `faces` only)
    DBuf edgeVerts, edgeStart, faceEdge, heFace, faceRank, rankFace, vdeg, starOff, starCur, star, topoHash;
    int topoF = -1, topoV = -1, nEdges = 0;
    unsigned long long topoHashVal = 0;
    // emission
    DBuf vfCounts, vfOffsets, eeCounts, eeOffsets, vfOut, eeOut;
    // narrowphase
    DBuf vfHit, eeHit, vfToi, eeToi, vfStage, eeStage, workVf, workEe, workTaskVf, workTaskEe, workSubVf, workSubEe, tasksVf, tasksEe, tlistVf, tlistEe, p1Status, p1Sbase, p1Queues, p1Sq, p1Xq, tempB, p1StatusB, p1SbaseB, p1QueuesB, p1SqB, p1XqB, p1Ve, qpack, histCntVf, histCntEe, histOffVf, histOffEe, histQ0, histQ1, histVst, histEta, histTime, histHit, histStage, histToi, spCtr, spQst, spQlo, spQhi, spLeafSt, selTmp, selA, selB, selC, selD, selCount;
    // penalty forces (penalty.cu)
    DBuf penF, penGroup, penContrib, penKeysA, penKeysB, penItemsA, penItemsB, penFired, penNewVf, penNewEe, penCtr;
    // pinned host scratch
    unsigned long long *h_counters = nullptr; // C_TOTAL entries
    // pinned host buffers for the hit lists returned by ccd_step (valid until the next call on the context)
    void *h_res[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t h_res_cap[4] = {0, 0, 0, 0};
    size_t candCap = 0, pairCap = 0, taskCapVf = 0, taskCapEe = 0, spQCap = 0, spLeafCapVf = 0, spLeafCapEe = 0;
    unsigned npSkipVf = 0, npSkipEe = 0;      // degrees / phases without records in the last single-step call (of npLastVf + npLastEe stencils)
    long long npLastVf = -1, npLastEe = -1;
    long long veUniqueEe = 0, veUniqueVf = 0;      // unique vertex-edge tests of the last single-step narrowphase call
    // sharding: ownership ranges chosen by the caller (ccd_set_shard_partition), load profile of the last sharded step
    std::vector<int> partP;      // ownership bounds over the sorted (Morton) positions of a sharded step (ccd_set_shard_partition)
    DBuf qlist, hist, needed, neededPre, alistV, alistE, kstartV, kstartE, keysV, keysE;
    unsigned long long *h_hist = nullptr;      // pinned, 2 * CCD_SHARD_BUCKETS
    int histF = 0;
    bool hist_valid = false;
};

// device counters layout (unsigned long long each)
enum { C_NCAND = 0, C_NPAIRS = 1, C_EARLY_VF = 2, C_NHIT_VF = 3, C_EARLY_EE = 4, C_NHIT_EE = 5, C_HASH = 6, C_NQUERY = 7, C_NA_VF = 8, C_NA_EE = 9, C_KCUR_VF = 10, C_KCUR_EE = 11, C_FRONT = 12 /* 4 counters: frontier sizes (ping-pong), their maximum, undecided face pairs */, C_NP_VF = 16 /* CCD_NP_COUNTERS counters per run: work-list entries, records, ... (narrowphase.cu K_*) */, C_NP_EE = 16 + CCD_NP_COUNTERS, C_NBIG = 16 + 2 * CCD_NP_COUNTERS /* 2 counters: items of the block-per-item emission (VF, EE) */, C_TOTAL = 16 + 2 * CCD_NP_COUNTERS + 2 };
enum { C_NWORK_VF = C_NP_VF, C_NTASK_VF = C_NP_VF + 1, C_NWORK_EE = C_NP_EE, C_NTASK_EE = C_NP_EE + 1 };

#define CK(call)                                                                                      \
    do                                                                                                \
    {                                                                                                 \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
        {                                                                                             \
            c->err = std::string(#call) + ": " + cudaGetErrorString(e_);                              \
            return CCD_ERR_CUDA;                                                                      \
        }                                                                                             \
    } while (0)

#define CKR(expr)                \
    do                           \
    {                            \
        int r_ = (expr);         \
        if (r_ != CCD_OK)        \
            return r_;           \
    } while (0)

static int ensure(ccd_context *c, DBuf &b, size_t bytes)
{
    if (bytes <= b.cap && b.p)
        return CCD_OK;
    if (b.p)
    {
        cudaStreamSynchronize(c->st);
        cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    }
    size_t ncap = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&b.p, ncap);
    if (e != cudaSuccess)
    {
        c->err = std::string("cudaMalloc failed: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return CCD_ERR_NOMEM;
    }
    b.cap = ncap;
    return CCD_OK;
}

template <typename T> static T *P(DBuf &b) { return (T *)b.p; }

static int upload(ccd_context *c, DBuf &b, const void *host, size_t bytes)
{
    CKR(ensure(c, b, bytes ? bytes : 8));
    if (bytes)
        CK(cudaMemcpyAsync(b.p, host, bytes, cudaMemcpyHostToDevice, c->st));
    return CCD_OK;
}

static void kdop_axes(double ax[13][3])
{
    // src/KDOPBroadPhase.cpp:12-31: thirteen directions, each divided by its norm
    static const double raw[13][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {1, -1, 0}, {0, 1, 1}, {0, 1, -1},
                                      {1, 0, 1}, {1, 0, -1}, {1, 1, 1}, {1, 1, -1}, {1, -1, 1}, {1, -1, -1}};
    for (int i = 0; i < 13; i++)
    {
        volatile double s01 = raw[i][0] * raw[i][0] + raw[i][1] * raw[i][1];
        volatile double s = s01 + raw[i][2] * raw[i][2];
        double nrm = sqrt(s);
        for (int k = 0; k < 3; k++)
            ax[i][k] = raw[i][k] / nrm;
    }
}

// launcher prototypes (broadphase.cu / distance.cu)
void ccdk_set_axes(const double axes[13][3]);
void ccdk_leaf_boxes(cudaStream_t st, int kind, int F, const int *faces, const double *q0, const double *q1, const long long *hoff,
                     const double *hpos, double eta, double *boxes, float *faabb, float *fkdop);
size_t ccdk_sort_temp_bytes(int n);
void ccdk_exclusive_sum64(cudaStream_t st, void *temp, size_t temp_bytes, int n_plus_1, const int *in, long long *out);
void ccdk_shard_queries(cudaStream_t st, int F, const unsigned *sortedFace, const int *faces, const int *faceEdge, int *facePos, const long long *starOff,
                        const int *star, const int *edgeStart, const int *heFace, int p0, int p1, int *qlist, unsigned long long *count, int *needed,
                        int *neededPre, void *temp, size_t temp_bytes);
void ccdk_shard_hist(cudaStream_t st, const int *counts, int n_items, const long long *segOff64, const int *segOff32, const int *segFace,
                     const int *facePos, int F, int nb, unsigned long long *hist);
void ccdk_adjacency_fill(cudaStream_t st, const unsigned long long *npairs, const int *pairL, const int *pairR, const long long *adjOff,
                         int *cursor, int *adj);
void ccdk_topology_edges(cudaStream_t st, int F, const int *faces, unsigned long long *keys_in, unsigned long long *keys_sorted,
                         unsigned *vals_in, unsigned *vals_sorted, void *temp, size_t temp_bytes, int *flags, int *ids, void *edgeVerts,
                         int *edgeStart, int *faceEdge, int *heFace);
void ccdk_topology_faceranks(cudaStream_t st, int F, const int *faces, unsigned *k32_in, unsigned *k32_out, unsigned *vals_a, unsigned *vals_b,
                             unsigned long long *k64_in, unsigned long long *k64_out, void *temp, size_t temp_bytes, int *flags, int *ids,
                             int *faceRank, int *rankFace);
void ccdk_topology_star(cudaStream_t st, int V, int F, const int *faces, int *vdeg, long long *starOff, int *cursor, int *star, void *temp,
                        size_t temp_bytes);
void ccdk_hash_ints(cudaStream_t st, long long n, const int *d, unsigned long long *out);
void ccdk_active_list(cudaStream_t st, int begin, int end, const long long *segOff64, const int *segOff32, const int *segFace, const int *deg,
                      int *counts, int *alist, unsigned long long *na, const int *facePos, int p0, int p1);
void ccdk_emit_sort(cudaStream_t st, bool is_vf, const int *alist, const unsigned long long *na, const int *faces, const long long *starOff,
                    const int *star, const int *edgeStart, const int *heFace, const long long *adjOff, const int *adj, const int *faceRank,
                    const int *rankFace, const int *faceEdge, const void *edgeVerts, const unsigned char *fixed, int *counts, long long *kstart,
                    int *keys_out, unsigned long long *kcursor, int *biglist, unsigned long long *nbig);
void ccdk_emit_write(cudaStream_t st, bool is_vf, const int *alist, const unsigned long long *na, int begin, const int *faces, const long long *starOff,
                     const int *star, const int *edgeStart, const int *heFace, const long long *adjOff, const int *adj, const int *faceRank,
                     const int *rankFace, const int *faceEdge, const void *edgeVerts, const unsigned char *fixed, const int *counts,
                     const long long *kstart, const int *keys_in, const long long *offsets, int *out);
void ccdk_face_aabb(cudaStream_t st, int F, const int *faces, const double *q0, const double *q1, double eta, float *faabb);
int ccdk_cluster_count_pow2(int F);
int ccdk_cluster_size(void);
void ccdk_build_cluster_tree(cudaStream_t st, int F, const int *faces, const float *faabb, unsigned *bounds, unsigned long long *keys_in,
                             unsigned long long *keys_sorted, unsigned *vals_in, unsigned *sortedFace, void *temp, size_t temp_bytes, float *heap,
                             float *sbox, void *sfaces);
int ccdk_leaf_record_words(void);
int ccdk_pair_traversal(cudaStream_t st, int kind, int F, const int *faces, const unsigned *sortedFace, const float *heap, float *rec, unsigned char *cmark,
                        const float *sbox, const void *sfaces, const double *boxes,
                        const double *q0, const double *q1, double eta, void *fr0, void *fr1, unsigned long long fcap, unsigned long long *fcount,
                        void *unsure, unsigned long long ucap, int *pairL, int *pairR, unsigned long long pcap, unsigned long long *npairs, int *deg,
                        const int *own, const int *ownPre, void *cand, unsigned long long ccap, unsigned long long *ncand);
void ccdk_dist_batch(cudaStream_t st, int which, long long n, const double *pts, const double *eta, double *vec, double *bary, unsigned char *flag);
void ccdk_vertex_min_dist2(cudaStream_t st, int V, int F, const double *verts, const int *faces, unsigned long long *out_bits);
void ccdk_stencil_min_dist(cudaStream_t st, bool is_vf, long long n, const int *stencils, const double *verts, unsigned long long *out_bits);
void ccdk_fp64_peak(cudaStream_t st, int iters, double *sink);

extern "C" {

const char *ccd_version(void) { return "ccd_b200 0.1 sm_100a"; }

int ccd_create(ccd_context **out, int device)
{
    if (!out)
        return CCD_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        cudaGetLastError();
        return CCD_ERR_NODEVICE;
    }
    if (device < 0 || device >= ndev)
        return CCD_ERR_ARG;
    ccd_context *c = new ccd_context();
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess)
    {
        delete c;
        return CCD_ERR_CUDA;
    }
    for (int i = 0; i <= CCD_N_STAGES; i++) c->sev[i] = nullptr;
    bool ok = cudaStreamCreateWithFlags(&c->st2, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&c->st3, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->evPack, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&c->evVf, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 3; i++) ok = ok && cudaEventCreateWithFlags(&c->evVe[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 4 && ok; i++)
        ok = cudaEventCreate(&c->ev[i]) == cudaSuccess;
    for (int i = 0; i <= CCD_N_STAGES && ok; i++)
        ok = cudaEventCreate(&c->sev[i]) == cudaSuccess;
    if (!ok)
    {
        cudaGetLastError();
        ccd_destroy(c);
        return CCD_ERR_CUDA;
    }
    if (cudaMallocHost((void **)&c->h_counters, sizeof(unsigned long long) * C_TOTAL) != cudaSuccess ||
        cudaMallocHost((void **)&c->h_hist, sizeof(unsigned long long) * 2 * CCD_SHARD_BUCKETS) != cudaSuccess)
    {
        cudaGetLastError();
        ccd_destroy(c);
        return CCD_ERR_NOMEM;
    }
    double ax[13][3];
    kdop_axes(ax);
    ccdk_set_axes(ax);
    if (ensure(c, c->counters, sizeof(unsigned long long) * C_TOTAL) != CCD_OK)
    {
        ccd_destroy(c);
        return CCD_ERR_NOMEM;
    }
    *out = c;
    return CCD_OK;
}

void ccd_destroy(ccd_context *c)
{
    if (!c)
        return;
    cudaSetDevice(c->device);
    if (c->st) cudaStreamSynchronize(c->st);
    if (c->st2) cudaStreamSynchronize(c->st2);
    if (c->st3) cudaStreamSynchronize(c->st3);
    DBuf *all[] = {&c->faces, &c->q0, &c->q1, &c->hoff, &c->htime, &c->hpos, &c->fixed, &c->vf_in, &c->ee_in, &c->vf_eta, &c->ee_eta,
                   &c->pts, &c->eta, &c->boxes, &c->faabb, &c->bounds, &c->keysA, &c->keysB, &c->valsA, &c->valsB, &c->temp, &c->facePos, &c->cmark,
                   &c->cand, &c->counters, &c->pairL, &c->pairR, &c->deg, &c->adjOff,
                   &c->cursor, &c->adj, &c->k32A, &c->k32B, &c->scanFlags, &c->scanIds, &c->edgeVerts, &c->edgeStart, &c->faceEdge,
                   &c->heFace, &c->faceRank, &c->rankFace, &c->vdeg, &c->starOff, &c->starCur, &c->star, &c->topoHash, &c->vfCounts,
                   &c->vfOffsets, &c->eeCounts, &c->eeOffsets, &c->vfOut, &c->eeOut, &c->vfHit, &c->eeHit, &c->vfToi, &c->eeToi,
                   &c->vfStage, &c->eeStage, &c->workVf, &c->workEe, &c->workTaskVf, &c->workTaskEe, &c->workSubVf, &c->workSubEe, &c->tasksVf, &c->tlistVf, &c->tlistEe, &c->p1Status, &c->p1Sbase, &c->p1Queues, &c->p1Sq, &c->p1Xq, &c->tempB, &c->p1StatusB, &c->p1SbaseB, &c->p1QueuesB, &c->p1SqB, &c->p1XqB, &c->p1Ve, &c->qpack, &c->histCntVf, &c->histCntEe, &c->histOffVf, &c->histOffEe, &c->histQ0, &c->histQ1, &c->histVst, &c->histEta, &c->histTime, &c->histHit, &c->histStage, &c->histToi, &c->spCtr, &c->spQst, &c->spQlo, &c->spQhi, &c->spLeafSt, &c->qlist, &c->hist, &c->needed, &c->neededPre, &c->alistV, &c->alistE, &c->kstartV, &c->kstartE, &c->keysV, &c->keysE, &c->tasksEe, &c->selTmp, &c->selA, &c->selB, &c->selC, &c->selD, &c->selCount, &c->heap, &c->srec, &c->sbox, &c->sfaces, &c->unsure, &c->frontA, &c->frontB, &c->bigV, &c->bigE,
                   &c->penF, &c->penGroup, &c->penContrib, &c->penKeysA, &c->penKeysB, &c->penItemsA, &c->penItemsB, &c->penFired, &c->penNewVf, &c->penNewEe, &c->penCtr};
    for (DBuf *b : all)
        if (b->p)
            cudaFree(b->p);
    if (c->h_counters)
        cudaFreeHost(c->h_counters);
    if (c->h_hist)
        cudaFreeHost(c->h_hist);
    for (int i = 0; i < 4; i++)
        if (c->h_res[i])
            cudaFreeHost(c->h_res[i]);
    for (int i = 0; i < 4; i++)
        if (c->ev[i])
            cudaEventDestroy(c->ev[i]);
    for (int i = 0; i <= CCD_N_STAGES; i++)
        if (c->sev[i])
            cudaEventDestroy(c->sev[i]);
    if (c->evFork) cudaEventDestroy(c->evFork);
    if (c->evJoin) cudaEventDestroy(c->evJoin);
    for (int i = 0; i < 3; i++)
        if (c->evVe[i]) cudaEventDestroy(c->evVe[i]);
    if (c->evPack) cudaEventDestroy(c->evPack);
    if (c->evVf) cudaEventDestroy(c->evVf);
    if (c->st3) cudaStreamDestroy(c->st3);
    if (c->st2) cudaStreamDestroy(c->st2);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}

const char *ccd_last_error(const ccd_context *c) { return c ? c->err.c_str() : "null context"; }
void ccd_free_host(void *p) { free(p); }

} // extern "C"

// ---------------------------------------------------------------------------------------------
// device-side orchestration
// ---------------------------------------------------------------------------------------------
struct BpResult
{
    long long nvf = 0, nee = 0, npairs = 0, ncand = 0;
};

static int sync_counters(ccd_context *c)
{
    CK(cudaMemcpyAsync(c->h_counters, c->counters.p, sizeof(unsigned long long) * C_TOTAL, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return CCD_OK;
}

// mesh tables that depend on `faces` only; rebuilt when the face array changes
static int ensure_topology(ccd_context *c, int V, int F, const int *d_faces)
{
    unsigned long long *ctr = P<unsigned long long>(c->counters);
    ccdk_hash_ints(c->st, 3ll * F, d_faces, ctr + C_HASH);
    c->launches += 1;
    CKR(sync_counters(c));
    unsigned long long h = c->h_counters[C_HASH];
    if (c->topoF == F && c->topoV == V && c->topoHashVal == h)
        return CCD_OK;
    const int n3 = 3 * F;
    size_t tb = ccdk_sort_temp_bytes(n3 > V + 1 ? n3 : V + 1);
    CKR(ensure(c, c->temp, tb));
    CKR(ensure(c, c->keysA, sizeof(unsigned long long) * (size_t)(n3 + 1)));
    CKR(ensure(c, c->keysB, sizeof(unsigned long long) * (size_t)(n3 + 1)));
    CKR(ensure(c, c->valsA, sizeof(unsigned) * (size_t)(n3 + 1)));
    CKR(ensure(c, c->valsB, sizeof(unsigned) * (size_t)(n3 + 1)));
    CKR(ensure(c, c->k32A, sizeof(unsigned) * (size_t)(F + 1)));
    CKR(ensure(c, c->k32B, sizeof(unsigned) * (size_t)(F + 1)));
    CKR(ensure(c, c->scanFlags, sizeof(int) * (size_t)(n3 + 1)));
    CKR(ensure(c, c->scanIds, sizeof(int) * (size_t)(n3 + 1)));
    CKR(ensure(c, c->edgeVerts, sizeof(int) * 2 * (size_t)(n3 + 1)));
    CKR(ensure(c, c->edgeStart, sizeof(int) * (size_t)(n3 + 2)));
    CKR(ensure(c, c->faceEdge, sizeof(int) * (size_t)(n3 + 1)));
    CKR(ensure(c, c->heFace, sizeof(int) * (size_t)(n3 + 1)));
    CKR(ensure(c, c->faceRank, sizeof(int) * (size_t)(F + 1)));
    CKR(ensure(c, c->rankFace, sizeof(int) * (size_t)(F + 1)));
    CKR(ensure(c, c->vdeg, sizeof(int) * (size_t)(V + 2)));
    CKR(ensure(c, c->starCur, sizeof(int) * (size_t)(V + 2)));
    CKR(ensure(c, c->starOff, sizeof(long long) * (size_t)(V + 2)));
    CKR(ensure(c, c->star, sizeof(int) * (size_t)(n3 + 1)));
    c->nEdges = 0;
    if (F > 0)
    {
        ccdk_topology_edges(c->st, F, d_faces, P<unsigned long long>(c->keysA), P<unsigned long long>(c->keysB), P<unsigned>(c->valsA),
                            P<unsigned>(c->valsB), c->temp.p, c->temp.cap, P<int>(c->scanFlags), P<int>(c->scanIds), c->edgeVerts.p,
                            P<int>(c->edgeStart), P<int>(c->faceEdge), P<int>(c->heFace));
        int ne = 0;
        CK(cudaMemcpyAsync(&ne, P<int>(c->scanIds) + (n3 - 1), sizeof(int), cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        c->nEdges = ne;
        ccdk_topology_faceranks(c->st, F, d_faces, P<unsigned>(c->k32A), P<unsigned>(c->k32B), P<unsigned>(c->valsA), P<unsigned>(c->valsB),
                                P<unsigned long long>(c->keysA), P<unsigned long long>(c->keysB), c->temp.p, c->temp.cap,
                                P<int>(c->scanFlags), P<int>(c->scanIds), P<int>(c->faceRank), P<int>(c->rankFace));
        c->launches += 5 + 8 + 7 + 8 + 8;
    }
    ccdk_topology_star(c->st, V, F, d_faces, P<int>(c->vdeg), P<long long>(c->starOff), P<int>(c->starCur), P<int>(c->star), c->temp.p,
                       c->temp.cap);
    c->launches += 3;
    CK(cudaGetLastError());
    c->topoF = F;
    c->topoV = V;
    c->topoHashVal = h;
    return CCD_OK;
}

// Broadphase on device-resident inputs.  Either (d_q0,d_q1) or (d_hoff,d_hpos) describes the History.
// Leaves the sorted stencils in c->vfOut / c->eeOut.
static int broadphase_device(ccd_context *c, int kind, int V, int F, const int *d_faces, const double *d_q0, const double *d_q1,
                             const long long *d_hoff, const double *d_hpos, double outerEta, const unsigned char *d_fixed,
                             int shard_rank, int shard_world, BpResult *res)
{
    if ((kind != CCD_KDOP && kind != CCD_AABB) || V < 0 || F < 0 || shard_world < 1 || shard_rank < 0 || shard_rank >= shard_world)
    {
        c->err = "broadphase: invalid argument";
        return CCD_ERR_ARG;
    }
    *res = BpResult();
    c->stage_valid = false;      // only a complete step (ccd_step_device) leaves a consistent set of stage events
    CKR(ensure(c, c->vfOut, 16));
    CKR(ensure(c, c->eeOut, 16));
    if (F == 0 || V == 0)
        return CCD_OK;
    unsigned long long *ctr = P<unsigned long long>(c->counters);
    cudaEventRecord(c->sev[ST_TOPOLOGY], c->st);
    CKR(ensure_topology(c, V, F, d_faces));

    const bool lazy_boxes = d_q0 != nullptr;      // single step: exact boxes are recomputed where they are needed
    if (!lazy_boxes) CKR(ensure(c, c->boxes, sizeof(double) * 2 * (size_t)kind * (size_t)F));
    CKR(ensure(c, c->faabb, sizeof(float) * 6 * (size_t)F));
    CKR(ensure(c, c->bounds, 64));
    CKR(ensure(c, c->temp, ccdk_sort_temp_bytes(3 * F > V + 1 ? 3 * F : V + 1)));
    CKR(ensure(c, c->keysA, sizeof(unsigned long long) * (size_t)(3 * F + 1)));
    CKR(ensure(c, c->keysB, sizeof(unsigned long long) * (size_t)(3 * F + 1)));
    CKR(ensure(c, c->valsA, sizeof(unsigned) * (size_t)(3 * F + 1)));
    CKR(ensure(c, c->valsB, sizeof(unsigned) * (size_t)(3 * F + 1)));
    CKR(ensure(c, c->deg, sizeof(int) * (size_t)(F + 2)));
    CKR(ensure(c, c->cursor, sizeof(int) * (size_t)(F + 2)));
    CKR(ensure(c, c->adjOff, sizeof(long long) * (size_t)(F + 2)));
    if (c->candCap == 0)
        c->candCap = (size_t)F * 24 + (1u << 16);
    if (c->pairCap == 0)
        c->pairCap = (size_t)F * 12 + (1u << 16);

    cudaEventRecord(c->sev[ST_BOXES], c->st);
    if (lazy_boxes) ccdk_face_aabb(c->st, F, d_faces, d_q0, d_q1, outerEta, P<float>(c->faabb));
    else ccdk_leaf_boxes(c->st, kind, F, d_faces, d_q0, d_q1, d_hoff, d_hpos, outerEta, P<double>(c->boxes), P<float>(c->faabb), nullptr);
    cudaEventRecord(c->sev[ST_TREE], c->st);
    {
        const size_t NLc = (size_t)ccdk_cluster_count_pow2(F), NLf = NLc * (size_t)ccdk_cluster_size();
        CKR(ensure(c, c->heap, sizeof(float) * 6 * (2 * NLc)));
        CKR(ensure(c, c->srec, sizeof(float) * (size_t)ccdk_leaf_record_words() * NLf));
        CKR(ensure(c, c->sbox, sizeof(float) * 6 * NLf));
        CKR(ensure(c, c->sfaces, sizeof(int) * 4 * NLf));
        CKR(ensure(c, c->cmark, NLc + 16));
        ccdk_build_cluster_tree(c->st, F, d_faces, P<float>(c->faabb), P<unsigned>(c->bounds), P<unsigned long long>(c->keysA), P<unsigned long long>(c->keysB),
                                P<unsigned>(c->valsA), P<unsigned>(c->valsB), c->temp.p, c->temp.cap, P<float>(c->heap), P<float>(c->sbox), c->sfaces.p);
        c->launches += 1 + 2 + 8 + 2 + 3;
    }
    const unsigned *sortedFace = P<unsigned>(c->valsB);
    cudaEventRecord(c->sev[ST_TRAVERSE_EXACT], c->st);

    // Ownership of a shard: the vertices (VF stencils) and unique edges (EE stencils) anchored in its range [p0, p1) of
    // sorted (Morton) positions — see shard_queries_kernel.  With one rank the range is everything; with several the
    // ranges come from the caller (ccd_set_shard_partition, balanced on the previous step's load profile; equal split
    // until then) and the rank traverses, tests and counts only for the faces its own emission reads.
    const bool sharded = shard_world > 1;
    const int E = c->nEdges;
    int p0 = 0, p1 = F;
    if (sharded)
    {
        if ((int)c->partP.size() == shard_world + 1 && c->partP[shard_world] == F)
        {
            p0 = c->partP[shard_rank]; p1 = c->partP[shard_rank + 1];
        }
        else
        {
            p0 = (int)(((long long)F * shard_rank) / shard_world); p1 = (int)(((long long)F * (shard_rank + 1)) / shard_world);
        }
        CKR(ensure(c, c->qlist, sizeof(int) * (size_t)(F + 32)));
        CKR(ensure(c, c->needed, sizeof(int) * (size_t)(F + 2)));
        CKR(ensure(c, c->neededPre, sizeof(int) * (size_t)(F + 2)));
        CKR(ensure(c, c->facePos, sizeof(int) * (size_t)(F + 2)));
    }
    const int *facePos = sharded ? P<int>(c->facePos) : nullptr;
    for (int attempt = 0; attempt < 8; attempt++)
    {
        CKR(ensure(c, c->cand, sizeof(int) * 2 * c->candCap));
        CKR(ensure(c, c->pairL, sizeof(int) * c->pairCap));
        CKR(ensure(c, c->pairR, sizeof(int) * c->pairCap));
        CK(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long) * 2, c->st));
        CK(cudaMemsetAsync(c->deg.p, 0, sizeof(int) * (size_t)(F + 2), c->st));
        if (sharded && attempt == 0)
        {
            CK(cudaMemsetAsync(ctr + C_NQUERY, 0, sizeof(unsigned long long), c->st));
            CK(cudaMemsetAsync(P<int>(c->needed) + F, 0, sizeof(int), c->st));
            ccdk_shard_queries(c->st, F, sortedFace, d_faces, P<int>(c->faceEdge), P<int>(c->facePos), P<long long>(c->starOff), P<int>(c->star),
                               P<int>(c->edgeStart), P<int>(c->heFace), p0, p1, P<int>(c->qlist), ctr + C_NQUERY, P<int>(c->needed), P<int>(c->neededPre),
                               c->temp.p, c->temp.cap);
            c->launches += 4;
        }
        // simultaneous pair traversal of the cluster tree, then the cluster pairs' face tests: pairs + degrees
        if (c->frontCap == 0) c->frontCap = (size_t)F * 2 + (1u << 16);
        CKR(ensure(c, c->frontA, sizeof(int) * 2 * c->frontCap));
        CKR(ensure(c, c->frontB, sizeof(int) * 2 * c->frontCap));
        if (c->unsureCap == 0) c->unsureCap = (size_t)F / 16 + (1u << 14);
        CKR(ensure(c, c->unsure, sizeof(int) * 2 * c->unsureCap));
        c->launches += ccdk_pair_traversal(c->st, kind, F, d_faces, sortedFace, P<float>(c->heap), P<float>(c->srec), P<unsigned char>(c->cmark), P<float>(c->sbox), c->sfaces.p, lazy_boxes ? nullptr : P<double>(c->boxes),
                                           lazy_boxes ? d_q0 : nullptr, d_q1, outerEta, c->frontA.p, c->frontB.p, c->frontCap, ctr + C_FRONT, c->unsure.p,
                                           c->unsureCap, P<int>(c->pairL), P<int>(c->pairR), c->pairCap, ctr + C_NPAIRS, P<int>(c->deg),
                                           sharded ? P<int>(c->needed) : nullptr, sharded ? P<int>(c->neededPre) : nullptr, c->cand.p, c->candCap, ctr + C_NCAND);
        CKR(sync_counters(c));
        const unsigned long long ncand = c->h_counters[C_NCAND], npairs = c->h_counters[C_NPAIRS];
        // counts keep running past the capacities (writes are guarded), so an overflow tells the size to retry with
        const bool front_over = c->h_counters[C_FRONT + 2] > c->frontCap;
        if (front_over) c->frontCap = (size_t)(c->h_counters[C_FRONT + 2] * 2 + 1024);      // a truncated level hides the size of the next: leave room
        const bool unsure_over = c->h_counters[C_FRONT + 3] > c->unsureCap;
        if (unsure_over) c->unsureCap = (size_t)(c->h_counters[C_FRONT + 3] * 2 + 1024);
        const bool cand_over = ncand > c->candCap, pair_over = npairs > c->pairCap;
        const bool again = cand_over || pair_over || front_over || unsure_over;
        if (cand_over) c->candCap = (size_t)(ncand + ncand / 4 + 1024);
        if (again)
        {
            // pairs found from a truncated candidate list are a lower bound only: leave generous room
            size_t want = (size_t)(npairs + npairs / 4 + 1024);
            if (cand_over || front_over) want = want > c->candCap / 2 ? want : c->candCap / 2;
            if (want > c->pairCap) c->pairCap = want;
        }
        if (!again)
        {
            res->ncand = (long long)c->h_counters[C_FRONT + 2];      // the largest frontier (= cluster pairs)
            res->npairs = (long long)npairs;
            if (getenv("CCD_BP_TRACE")) fprintf(stderr, "[bp] cluster pairs %llu, face-pair candidates %llu, undecided %llu, face pairs %llu\n", c->h_counters[C_FRONT + 2], ncand, c->h_counters[C_FRONT + 3], npairs);
            break;
        }
        if (attempt == 7)
        {
            c->err = "broadphase: candidate buffers kept overflowing";
            return CCD_ERR_NOMEM;
        }
    }
    CK(cudaGetLastError());

    // face adjacency CSR (both directions)
    cudaEventRecord(c->sev[ST_ADJACENCY], c->st);
    ccdk_exclusive_sum64(c->st, c->temp.p, c->temp.cap, F + 1, P<int>(c->deg), P<long long>(c->adjOff));
    CKR(ensure(c, c->adj, sizeof(int) * (size_t)(2 * res->npairs + 1)));
    CK(cudaMemsetAsync(c->cursor.p, 0, sizeof(int) * (size_t)(F + 2), c->st));
    ccdk_adjacency_fill(c->st, ctr + C_NPAIRS, P<int>(c->pairL), P<int>(c->pairR), P<long long>(c->adjOff), P<int>(c->cursor), P<int>(c->adj));
    c->launches += 3;

    // Stencil counts of the owned vertices / unique edges (zero for everything else), scanned into write offsets
    CKR(ensure(c, c->vfCounts, sizeof(int) * (size_t)(V + 2)));
    CKR(ensure(c, c->vfOffsets, sizeof(long long) * (size_t)(V + 2)));
    CKR(ensure(c, c->eeCounts, sizeof(int) * (size_t)(E + 2)));
    CKR(ensure(c, c->eeOffsets, sizeof(long long) * (size_t)(E + 2)));
    cudaEventRecord(c->sev[ST_EMIT_COUNT], c->st);
    {
        // warp-cooperative emission (broadphase.cu section 6b): active items -> unique sorted keys + counts
        const size_t ordered = (size_t)res->npairs * 2;      // upper bound of the (face, neighbour) entries in the adjacency lists
        CKR(ensure(c, c->alistV, sizeof(int) * (size_t)(V + 32)));
        CKR(ensure(c, c->alistE, sizeof(int) * (size_t)(E + 32)));
        CKR(ensure(c, c->kstartV, sizeof(long long) * (size_t)(V + 2)));
        CKR(ensure(c, c->kstartE, sizeof(long long) * (size_t)(E + 2)));
        CKR(ensure(c, c->keysV, sizeof(int) * (3 * ordered + 64)));      // every adjacency entry is seen by the face's 3 vertices
        CKR(ensure(c, c->keysE, sizeof(int) * (9 * ordered + 64)));      // ... and by its 3 edges, each against the neighbour's 3 edges
        CK(cudaMemsetAsync(ctr + C_NA_VF, 0, sizeof(unsigned long long) * 4, c->st));
        CK(cudaMemsetAsync(ctr + C_NBIG, 0, sizeof(unsigned long long) * 2, c->st));
        CKR(ensure(c, c->bigV, sizeof(int) * (size_t)(V + 32)));
        CKR(ensure(c, c->bigE, sizeof(int) * (size_t)(E + 32)));
        // the vertex-face chain (active list, key sets, scan of the counts) on the second stream beside the edge-edge chain:
        // nothing on one GPU at full size, a shorter stage on a shard (small grids leave most of the machine idle)
        CK(cudaEventRecord(c->evPack, c->st));
        CK(cudaStreamWaitEvent(c->st3, c->evPack, 0));
        ccdk_active_list(c->st3, 0, V, P<long long>(c->starOff), nullptr, P<int>(c->star), P<int>(c->deg), P<int>(c->vfCounts), P<int>(c->alistV), ctr + C_NA_VF, facePos, p0, p1);
        ccdk_active_list(c->st, 0, E, nullptr, P<int>(c->edgeStart), P<int>(c->heFace), P<int>(c->deg), P<int>(c->eeCounts), P<int>(c->alistE), ctr + C_NA_EE, facePos, p0, p1);
        ccdk_emit_sort(c->st3, true, P<int>(c->alistV), ctr + C_NA_VF, d_faces, P<long long>(c->starOff), P<int>(c->star), P<int>(c->edgeStart),
                       P<int>(c->heFace), P<long long>(c->adjOff), P<int>(c->adj), P<int>(c->faceRank), P<int>(c->rankFace), P<int>(c->faceEdge),
                       c->edgeVerts.p, d_fixed, P<int>(c->vfCounts), P<long long>(c->kstartV), P<int>(c->keysV), ctr + C_KCUR_VF, P<int>(c->bigV), ctr + C_NBIG);
        ccdk_emit_sort(c->st, false, P<int>(c->alistE), ctr + C_NA_EE, d_faces, P<long long>(c->starOff), P<int>(c->star), P<int>(c->edgeStart),
                       P<int>(c->heFace), P<long long>(c->adjOff), P<int>(c->adj), P<int>(c->faceRank), P<int>(c->rankFace), P<int>(c->faceEdge),
                       c->edgeVerts.p, d_fixed, P<int>(c->eeCounts), P<long long>(c->kstartE), P<int>(c->keysE), ctr + C_KCUR_EE, P<int>(c->bigE), ctr + C_NBIG + 1);
        c->launches += 6;
    }
    // the scan reads one element past the range (never added to anything it outputs)
    CKR(ensure(c, c->tempB, ccdk_sort_temp_bytes(V + 1)));
    ccdk_exclusive_sum64(c->st3, c->tempB.p, c->tempB.cap, V + 1, P<int>(c->vfCounts), P<long long>(c->vfOffsets));
    ccdk_exclusive_sum64(c->st, c->temp.p, c->temp.cap, E + 1, P<int>(c->eeCounts), P<long long>(c->eeOffsets));
    CK(cudaEventRecord(c->evVf, c->st3));
    CK(cudaStreamWaitEvent(c->st, c->evVf, 0));
    c->launches += 4;
    c->hist_valid = false;
    if (sharded)
    {
        CKR(ensure(c, c->hist, sizeof(unsigned long long) * 2 * CCD_SHARD_BUCKETS));
        CK(cudaMemsetAsync(c->hist.p, 0, sizeof(unsigned long long) * 2 * CCD_SHARD_BUCKETS, c->st));
        ccdk_shard_hist(c->st, P<int>(c->vfCounts), V, P<long long>(c->starOff), nullptr, P<int>(c->star), facePos, F, CCD_SHARD_BUCKETS, P<unsigned long long>(c->hist));
        ccdk_shard_hist(c->st, P<int>(c->eeCounts), E, nullptr, P<int>(c->edgeStart), P<int>(c->heFace), facePos, F, CCD_SHARD_BUCKETS, P<unsigned long long>(c->hist) + CCD_SHARD_BUCKETS);
        CK(cudaMemcpyAsync(c->h_hist, c->hist.p, sizeof(unsigned long long) * 2 * CCD_SHARD_BUCKETS, cudaMemcpyDeviceToHost, c->st));
        c->histF = F;
        c->hist_valid = true;
        c->launches += 2;
    }
    long long range[2] = {0, 0};
    CK(cudaMemcpyAsync(&range[0], P<long long>(c->vfOffsets) + V, sizeof(long long), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(&range[1], P<long long>(c->eeOffsets) + E, sizeof(long long), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    res->nvf = range[0];
    res->nee = range[1];
    CKR(ensure(c, c->vfOut, sizeof(int) * 4 * (size_t)(res->nvf + 1)));
    CKR(ensure(c, c->eeOut, sizeof(int) * 4 * (size_t)(res->nee + 1)));
    cudaEventRecord(c->sev[ST_EMIT_WRITE], c->st);
    CK(cudaEventRecord(c->evPack, c->st));
    CK(cudaStreamWaitEvent(c->st3, c->evPack, 0));
    ccdk_emit_write(c->st3, true, P<int>(c->alistV), ctr + C_NA_VF, 0, d_faces, P<long long>(c->starOff), P<int>(c->star), P<int>(c->edgeStart),
                    P<int>(c->heFace), P<long long>(c->adjOff), P<int>(c->adj), P<int>(c->faceRank), P<int>(c->rankFace), P<int>(c->faceEdge),
                    c->edgeVerts.p, d_fixed, P<int>(c->vfCounts), P<long long>(c->kstartV), P<int>(c->keysV), P<long long>(c->vfOffsets), P<int>(c->vfOut));
    ccdk_emit_write(c->st, false, P<int>(c->alistE), ctr + C_NA_EE, 0, d_faces, P<long long>(c->starOff), P<int>(c->star), P<int>(c->edgeStart),
                    P<int>(c->heFace), P<long long>(c->adjOff), P<int>(c->adj), P<int>(c->faceRank), P<int>(c->rankFace), P<int>(c->faceEdge),
                    c->edgeVerts.p, d_fixed, P<int>(c->eeCounts), P<long long>(c->kstartE), P<int>(c->keysE), P<long long>(c->eeOffsets), P<int>(c->eeOut));
    CK(cudaEventRecord(c->evVf, c->st3));
    CK(cudaStreamWaitEvent(c->st, c->evVf, 0));
    c->launches += 2;
    CK(cudaGetLastError());
    return CCD_OK;
}

// Narrowphase over device-resident stencils; results left in c->vfHit/... ; summary via counters.
static int narrowphase_device(ccd_context *c, int V, long long nvf, const int *d_vf, const double *d_vf_eta, long long nee, const int *d_ee,
                              const double *d_ee_eta, double eta_all_vf, double eta_all_ee, const double *d_q0, const double *d_q1, int vstride, const long long *d_hoff,
                              const double *d_htime, const double *d_hpos, ccd_np_summary *sum)
{
    unsigned long long *ctr = P<unsigned long long>(c->counters);
    c->stage_valid = false;
    CKR(ensure(c, c->vfHit, (size_t)nvf + 16));
    CKR(ensure(c, c->eeHit, (size_t)nee + 16));
    CKR(ensure(c, c->vfStage, (size_t)nvf + 16));
    CKR(ensure(c, c->eeStage, (size_t)nee + 16));
    CKR(ensure(c, c->vfToi, sizeof(double) * ((size_t)nvf + 2)));
    CKR(ensure(c, c->eeToi, sizeof(double) * ((size_t)nee + 2)));
    CKR(ensure(c, c->workVf, sizeof(int) * ((size_t)nvf + 32)));
    CKR(ensure(c, c->workEe, sizeof(int) * ((size_t)nee + 32)));
    CKR(ensure(c, c->workTaskVf, sizeof(int) * ((size_t)nvf + 32)));
    CKR(ensure(c, c->workTaskEe, sizeof(int) * ((size_t)nee + 32)));
    CKR(ensure(c, c->workSubVf, sizeof(int) * 5 * ((size_t)nvf + 32)));
    CKR(ensure(c, c->workSubEe, sizeof(int) * 5 * ((size_t)nee + 32)));
    // Single step with both stencil types: the vertex-face run executes on a stream of its own BESIDE the edge-edge run (most
    // kernels of either pipeline are latency- or bandwidth-bound and leave issue slots, registers and block slots free —
    // profiles/) with pass-1 scratch of its own.  CCD_NP_SEQUENTIAL=1 or CCD_NP_TRACE=1: one after the other.
    const bool concurrent = d_q0 != nullptr && nvf > 0 && nee > 0 && !getenv("CCD_NP_SEQUENTIAL") && !getenv("CCD_NP_TRACE");
    {
        // pass-1 scratch: shared by the VF and the EE run when they execute one after the other
        const size_t nmax = (size_t)(concurrent ? nee : (nvf > nee ? nvf : nee)) + 32;
        CKR(ensure(c, c->p1Status, sizeof(unsigned) * nmax));
        CKR(ensure(c, c->p1Sbase, sizeof(int) * 5 * nmax));
        CKR(ensure(c, c->p1Queues, sizeof(int) * 13 * nmax));
        CKR(ensure(c, c->p1Sq, sizeof(int) * 4 * nmax));
        CKR(ensure(c, c->p1Xq, sizeof(int) * 10 * nmax));
        if (concurrent)
        {
            const size_t nv = (size_t)nvf + 32;
            CKR(ensure(c, c->p1StatusB, sizeof(unsigned) * nv));
            CKR(ensure(c, c->p1SbaseB, sizeof(int) * 5 * nv));
            CKR(ensure(c, c->p1QueuesB, sizeof(int) * 13 * nv));
            CKR(ensure(c, c->p1SqB, sizeof(int) * 4 * nv));
            CKR(ensure(c, c->p1XqB, sizeof(int) * 10 * nv));
        }
        // vertex-edge de-duplication scratch: one region per run (the vertex-face run consults the edge-edge run's table)
        CKR(ensure(c, c->p1Ve, 12 * (size_t)ccdk_np_ve_slots(nee) + 48 * ((size_t)nee + 32) + 12 * (size_t)ccdk_np_ve_slots(nvf) + 48 * ((size_t)nvf + 32) + 64));
    }
    const float *d_vbox = nullptr;
    if (d_q0)
    {
        CKR(ensure(c, c->qpack, (sizeof(double) * 8 + sizeof(float) * 8) * ((size_t)V + 1)));
        d_vbox = reinterpret_cast<const float *>(P<double>(c->qpack) + 8 * ((size_t)V + 1));
        ccdk_pack_positions(c->st, V, d_q0, d_q1, vstride, P<double>(c->qpack), const_cast<float *>(d_vbox));
        d_q0 = P<double>(c->qpack);
        d_q1 = d_q0 + 4;
        vstride = 8;
        c->launches += 1;
    }
    int nl = 0;
    for (int attempt = 0; attempt < 4; attempt++)
    {
        // task records: one per polynomial that needs the root isolator; grown on demand (count known after pass 1)
        if (c->taskCapVf < (size_t)nvf / 2 + 1024) c->taskCapVf = (size_t)nvf / 2 + 1024;
        if (c->taskCapEe < (size_t)nee + 1024) c->taskCapEe = (size_t)nee + 1024;
        CKR(ensure(c, c->tasksVf, 128 * c->taskCapVf));
        CKR(ensure(c, c->tasksEe, 128 * c->taskCapEe));
        CKR(ensure(c, c->tlistVf, sizeof(int) * 6 * c->taskCapVf));
        CKR(ensure(c, c->tlistEe, sizeof(int) * 6 * c->taskCapEe));
        // (a run with no stencils of one type launches nothing for it: its counters must not keep an earlier call's values)
        CK(cudaMemsetAsync(ctr + C_NP_VF, 0, sizeof(unsigned long long) * 2 * CCD_NP_COUNTERS, c->st));
        unsigned long long init[4] = {0xFFFFFFFFFFFFFFFFull, 0ull, 0xFFFFFFFFFFFFFFFFull, 0ull};
        memcpy(c->h_counters + 8, init, sizeof(init));
        CK(cudaMemcpyAsync(ctr + C_EARLY_VF, c->h_counters + 8, sizeof(init), cudaMemcpyHostToDevice, c->st));
        // the edge-edge run goes first: its general routine (a few hundred long single-lane walks) then runs on the side
        // stream beside the vertex-face run
        const bool single_step = d_q0 != nullptr;
        // hash-set sizes of the vertex-edge de-duplication: 4 x the unique tests the previous call saw (the set then fits in
        // L2: 2.1M unique tests of 19M edge-edge stencils at the 4M-triangle cloth), never more than the first-call size
        unsigned slotsEe = ccdk_np_ve_slots(nee), slotsVf = ccdk_np_ve_slots(nvf);
        if (c->veUniqueEe > 0) { const unsigned s2 = ccdk_np_ve_slots(4 * c->veUniqueEe); if (s2 < slotsEe) slotsEe = s2; }
        if (c->veUniqueVf > 0) { const unsigned s2 = ccdk_np_ve_slots(4 * c->veUniqueVf); if (s2 < slotsVf) slotsVf = s2; }
        const size_t ve_off_vf = (12 * (size_t)ccdk_np_ve_slots(nee) + 48 * ((size_t)nee + 32) + 15) & ~(size_t)15;
        // the vertex-face run looks its vertex-edge tests up in the edge-edge run's table first (same eta): side by side, the
        // runs are then chained through three events at the points where the table, the results and the records are needed
        const bool share_ve = single_step && nee > 0 && nvf > 0 && !d_vf_eta && !d_ee_eta && eta_all_vf == eta_all_ee && !getenv("CCD_NO_VE_SHARE");
        const int role_ee = (concurrent && share_ve) ? 1 : 0, role_vf = (concurrent && share_ve) ? 2 : 0;
        // root-isolation kernels of a degree and phase that had no record in the previous call of (nearly) the same size are not launched
        // (3 launches each; never the distance sextics of the first phase): see launch_single_step
        unsigned skipEe = 0, skipVf = 0;
        auto near = [](long long a, long long b) { const long long d = a > b ? a - b : b - a; return b >= 0 && d <= b / 32; };      // a shard's lists move a little with every rebalancing
        if (single_step && near(nvf, c->npLastVf) && near(nee, c->npLastEe) && !getenv("CCD_NP_NO_SKIP")) { skipEe = c->npSkipEe; skipVf = c->npSkipVf; }
        if (const char *fs = getenv("CCD_NP_FORCE_SKIP")) skipEe = skipVf = (unsigned)strtoul(fs, nullptr, 0);      // tests: the fall-back of records left pending
        cudaStream_t svf = concurrent ? c->st3 : c->st;
        if (concurrent)
        {
            // the vertex-face stream starts once the packed positions (and the counters reset above) are in place
            CK(cudaEventRecord(c->evPack, c->st));
            CK(cudaStreamWaitEvent(c->st3, c->evPack, 0));
        }
        cudaEventRecord(c->sev[ST_NP_EE], c->st);
        nl += ccdk_narrowphase(c->st, false, nee, d_ee, d_ee_eta, eta_all_ee, d_q0, d_q1, vstride, d_vbox, d_hoff, d_htime, d_hpos, P<unsigned char>(c->eeHit),
                               P<double>(c->eeToi), P<unsigned char>(c->eeStage), ctr + C_EARLY_EE, ctr + C_NHIT_EE, P<int>(c->workEe),
                               P<int>(c->workTaskEe), P<int>(c->workSubEe), P<double>(c->tasksEe), P<int>(c->tlistEe), c->taskCapEe, P<unsigned>(c->p1Status), P<int>(c->p1Sbase),
                               P<int>(c->p1Queues), P<int>(c->p1Sq), P<int>(c->p1Xq), ctr + C_NP_EE, c->p1Ve.p, slotsEe, V, c->st2, c->evFork, c->evJoin, nullptr, 0, 0, nullptr, c->evVe, role_ee, skipEe);
        cudaEventRecord(c->sev[ST_NP_VF], c->st);      // concurrent runs: "np_ee" is the edge-edge run, "np_vf" what is left of the vertex-face run after it
        nl += ccdk_narrowphase(svf, true, nvf, d_vf, d_vf_eta, eta_all_vf, d_q0, d_q1, vstride, d_vbox, d_hoff, d_htime, d_hpos, P<unsigned char>(c->vfHit),
                               P<double>(c->vfToi), P<unsigned char>(c->vfStage), ctr + C_EARLY_VF, ctr + C_NHIT_VF, P<int>(c->workVf),
                               P<int>(c->workTaskVf), P<int>(c->workSubVf), P<double>(c->tasksVf), P<int>(c->tlistVf), c->taskCapVf,
                               concurrent ? P<unsigned>(c->p1StatusB) : P<unsigned>(c->p1Status), concurrent ? P<int>(c->p1SbaseB) : P<int>(c->p1Sbase),
                               concurrent ? P<int>(c->p1QueuesB) : P<int>(c->p1Queues), concurrent ? P<int>(c->p1SqB) : P<int>(c->p1Sq),
                               concurrent ? P<int>(c->p1XqB) : P<int>(c->p1Xq), ctr + C_NP_VF, (char *)c->p1Ve.p + ve_off_vf, slotsVf, V, nullptr, nullptr, nullptr,
                               share_ve ? c->p1Ve.p : nullptr, share_ve ? slotsEe : 0u, nee, share_ve ? P<double>(c->tasksEe) : nullptr, c->evVe, role_vf, skipVf);
        if (concurrent)
        {
            CK(cudaEventRecord(c->evVf, c->st3));
            CK(cudaStreamWaitEvent(c->st, c->evVf, 0));
        }
        if (single_step && nee > 0 && !getenv("CCD_NP_TRACE")) CK(cudaStreamWaitEvent(c->st, c->evJoin, 0));
        cudaEventRecord(c->sev[CCD_N_STAGES], c->st);
        CK(cudaGetLastError());
        CKR(sync_counters(c));
        const bool single = d_q0 != nullptr;
        if (single)
        {
            c->veUniqueEe = nee > 0 ? (long long)c->h_counters[C_NP_EE + CCD_NP_KVEU] : 0;
            c->veUniqueVf = nvf > 0 ? (long long)c->h_counters[C_NP_VF + CCD_NP_KVEU] : 0;
            unsigned m[2] = {0, 0};
            for (int run = 0; run < 2; run++)
                for (int phase = 0; phase < 2; phase++)
                    for (int d = 0; d < 4; d++)
                        if (!(phase == 0 && d == 3) && c->h_counters[(run ? C_NP_EE : C_NP_VF) + (phase ? CCD_NP_KNDEG2 : CCD_NP_KNDEG) + d] == 0)
                            m[run] |= 1u << (4 * phase + d);
            c->npSkipVf = m[0]; c->npSkipEe = m[1];
            c->npLastVf = nvf; c->npLastEe = nee;
        }
        const unsigned long long tv = single ? c->h_counters[C_NTASK_VF] : 0, te = single ? c->h_counters[C_NTASK_EE] : 0;
        if (tv + 5 <= c->taskCapVf && te + 5 <= c->taskCapEe)
            break;
        if (tv + 5 > c->taskCapVf) c->taskCapVf = (size_t)(tv + tv / 8 + 1024);
        if (te + 5 > c->taskCapEe) c->taskCapEe = (size_t)(te + te / 8 + 1024);
        if (c->taskCapVf >= (1u << 28) - 8 || c->taskCapEe >= (1u << 28) - 8)
        {
            c->err = "narrowphase: more than 2^28 polynomial records in one call; split the stencil list";
            return CCD_ERR_NOMEM;
        }
    }
    c->launches += nl;
    if (sum)
    {
        sum->n_vf_hits = (int64_t)c->h_counters[C_NHIT_VF];
        sum->n_ee_hits = (int64_t)c->h_counters[C_NHIT_EE];
        unsigned long long b = c->h_counters[C_EARLY_VF] < c->h_counters[C_EARLY_EE] ? c->h_counters[C_EARLY_VF] : c->h_counters[C_EARLY_EE];
        double t;
        memcpy(&t, &b, sizeof(t));
        sum->earliest_toi = (b == 0xFFFFFFFFFFFFFFFFull) ? INFINITY : t;
    }
    return CCD_OK;
}

// Narrowphase over a multi-entry History (narrowphase.cu, "multi-entry History, staged"): every (stencil, stitched segment)
// pair becomes a virtual single-step stencil on four virtual vertices, the single-step pipeline runs on all of them, and
// each stencil takes its first hitting segment.  CCD_HISTORY_ONE_THREAD=1 keeps the one-thread-per-stencil kernel
// (diagnostics: both give the same bits).
static int narrowphase_history_device(ccd_context *c, int V, long long nvf, const int *d_vf, const double *d_vf_eta, long long nee, const int *d_ee,
                                      const double *d_ee_eta, double eta_all_vf, double eta_all_ee, const long long *d_hoff, const double *d_htime,
                                      const double *d_hpos, ccd_np_summary *sum)
{
    if (getenv("CCD_HISTORY_ONE_THREAD") || nvf + nee == 0)
        return narrowphase_device(c, V, nvf, d_vf, d_vf_eta, nee, d_ee, d_ee_eta, eta_all_vf, eta_all_ee, nullptr, nullptr, 0, d_hoff, d_htime, d_hpos, sum);
    unsigned long long *ctr = P<unsigned long long>(c->counters);
    const long long nmax = nvf > nee ? nvf : nee;
    CKR(ensure(c, c->histCntVf, sizeof(int) * ((size_t)nvf + 2)));
    CKR(ensure(c, c->histCntEe, sizeof(int) * ((size_t)nee + 2)));
    CKR(ensure(c, c->histOffVf, sizeof(long long) * ((size_t)nvf + 2)));
    CKR(ensure(c, c->histOffEe, sizeof(long long) * ((size_t)nee + 2)));
    CKR(ensure(c, c->temp, ccdk_sort_temp_bytes((int)(nmax + 2))));
    CK(cudaMemsetAsync(P<int>(c->histCntVf) + nvf, 0, sizeof(int), c->st));
    CK(cudaMemsetAsync(P<int>(c->histCntEe) + nee, 0, sizeof(int), c->st));
    ccdk_hist_count(c->st, nvf, d_vf, d_hoff, d_htime, d_hpos, P<int>(c->histCntVf));
    ccdk_hist_count(c->st, nee, d_ee, d_hoff, d_htime, d_hpos, P<int>(c->histCntEe));
    ccdk_exclusive_sum64(c->st, c->temp.p, c->temp.cap, (int)nvf + 1, P<int>(c->histCntVf), P<long long>(c->histOffVf));
    ccdk_exclusive_sum64(c->st, c->temp.p, c->temp.cap, (int)nee + 1, P<int>(c->histCntEe), P<long long>(c->histOffEe));
    long long ns[2] = {0, 0};
    CK(cudaMemcpyAsync(&ns[0], P<long long>(c->histOffVf) + nvf, sizeof(long long), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(&ns[1], P<long long>(c->histOffEe) + nee, sizeof(long long), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    const long long nsVf = ns[0], nsEe = ns[1], nsAll = nsVf + nsEe;
    if (4 * nsAll >= (1ll << 31) - 8)
    {
        c->err = "narrowphase: more than 2^29 stitched segments in one call; split the stencil list";
        return CCD_ERR_NOMEM;
    }
    c->launches += 4 + 4;
    CKR(ensure(c, c->histQ0, sizeof(double) * 12 * ((size_t)nsAll + 1)));
    CKR(ensure(c, c->histQ1, sizeof(double) * 12 * ((size_t)nsAll + 1)));
    CKR(ensure(c, c->histVst, sizeof(int) * 4 * ((size_t)nsAll + 1)));
    CKR(ensure(c, c->histEta, sizeof(double) * ((size_t)nsAll + 1)));
    CKR(ensure(c, c->histTime, sizeof(double) * 2 * ((size_t)nsAll + 1)));
    int *vstVf = P<int>(c->histVst), *vstEe = vstVf + 4 * nsVf;
    double *etaVf = P<double>(c->histEta), *etaEe = etaVf + nsVf;
    double *timeVf = P<double>(c->histTime), *timeEe = timeVf + 2 * nsVf;
    ccdk_hist_fill(c->st, nvf, d_vf, d_vf_eta, eta_all_vf, d_hoff, d_htime, d_hpos, P<long long>(c->histOffVf), 0, P<double>(c->histQ0), P<double>(c->histQ1),
                   vstVf, etaVf, timeVf);
    ccdk_hist_fill(c->st, nee, d_ee, d_ee_eta, eta_all_ee, d_hoff, d_htime, d_hpos, P<long long>(c->histOffEe), nsVf, P<double>(c->histQ0), P<double>(c->histQ1),
                   vstEe, etaEe, timeEe);
    // the result buffers are shared with the virtual run: size them for both before anything is written
    CKR(ensure(c, c->vfHit, (size_t)(nvf > nsVf ? nvf : nsVf) + 16));
    CKR(ensure(c, c->eeHit, (size_t)(nee > nsEe ? nee : nsEe) + 16));
    CKR(ensure(c, c->vfStage, (size_t)(nvf > nsVf ? nvf : nsVf) + 16));
    CKR(ensure(c, c->eeStage, (size_t)(nee > nsEe ? nee : nsEe) + 16));
    CKR(ensure(c, c->vfToi, sizeof(double) * ((size_t)(nvf > nsVf ? nvf : nsVf) + 2)));
    CKR(ensure(c, c->eeToi, sizeof(double) * ((size_t)(nee > nsEe ? nee : nsEe) + 2)));
    CKR(ensure(c, c->histHit, (size_t)(nvf + nee) + 16));
    CKR(ensure(c, c->histStage, (size_t)(nvf + nee) + 16));
    CKR(ensure(c, c->histToi, sizeof(double) * ((size_t)(nvf + nee) + 2)));
    const int launches_before = c->launches;
    ccd_np_summary virt;
    CKR(narrowphase_device(c, (int)(4 * nsAll), nsVf, vstVf, etaVf, nsEe, vstEe, etaEe, 0.0, 0.0, P<double>(c->histQ0), P<double>(c->histQ1), 3, nullptr, nullptr,
                           nullptr, &virt));
    (void)launches_before;
    // per stencil: the first hitting segment; the counters of the virtual run are replaced by the stencils' own
    unsigned long long init[4] = {0xFFFFFFFFFFFFFFFFull, 0ull, 0xFFFFFFFFFFFFFFFFull, 0ull};
    memcpy(c->h_counters + 8, init, sizeof(init));
    CK(cudaMemcpyAsync(ctr + C_EARLY_VF, c->h_counters + 8, sizeof(init), cudaMemcpyHostToDevice, c->st));
    unsigned char *hHit = P<unsigned char>(c->histHit), *hStage = P<unsigned char>(c->histStage);
    double *hToi = P<double>(c->histToi);
    ccdk_hist_reduce(c->st, nvf, P<long long>(c->histOffVf), P<unsigned char>(c->vfHit), P<double>(c->vfToi), P<unsigned char>(c->vfStage), timeVf, hHit, hToi,
                     hStage, ctr + C_EARLY_VF, ctr + C_NHIT_VF);
    ccdk_hist_reduce(c->st, nee, P<long long>(c->histOffEe), P<unsigned char>(c->eeHit), P<double>(c->eeToi), P<unsigned char>(c->eeStage), timeEe, hHit + nvf,
                     hToi + nvf, hStage + nvf, ctr + C_EARLY_EE, ctr + C_NHIT_EE);
    c->launches += 2;
    if (nvf > 0)
    {
        CK(cudaMemcpyAsync(c->vfHit.p, hHit, (size_t)nvf, cudaMemcpyDeviceToDevice, c->st));
        CK(cudaMemcpyAsync(c->vfStage.p, hStage, (size_t)nvf, cudaMemcpyDeviceToDevice, c->st));
        CK(cudaMemcpyAsync(c->vfToi.p, hToi, sizeof(double) * (size_t)nvf, cudaMemcpyDeviceToDevice, c->st));
    }
    if (nee > 0)
    {
        CK(cudaMemcpyAsync(c->eeHit.p, hHit + nvf, (size_t)nee, cudaMemcpyDeviceToDevice, c->st));
        CK(cudaMemcpyAsync(c->eeStage.p, hStage + nvf, (size_t)nee, cudaMemcpyDeviceToDevice, c->st));
        CK(cudaMemcpyAsync(c->eeToi.p, hToi + nvf, sizeof(double) * (size_t)nee, cudaMemcpyDeviceToDevice, c->st));
    }
    cudaEventRecord(c->sev[CCD_N_STAGES], c->st);
    CKR(sync_counters(c));
    if (sum)
    {
        sum->n_vf_hits = (int64_t)c->h_counters[C_NHIT_VF];
        sum->n_ee_hits = (int64_t)c->h_counters[C_NHIT_EE];
        unsigned long long b = c->h_counters[C_EARLY_VF] < c->h_counters[C_EARLY_EE] ? c->h_counters[C_EARLY_VF] : c->h_counters[C_EARLY_EE];
        double t;
        memcpy(&t, &b, sizeof(t));
        sum->earliest_toi = (b == 0xFFFFFFFFFFFFFFFFull) ? INFINITY : t;
    }
    return CCD_OK;
}

static int download_stencils(ccd_context *c, const DBuf &src, long long n, int32_t **out)
{
    int32_t *h = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)(n > 0 ? n : 1));
    if (!h)
        return CCD_ERR_NOMEM;
    if (n > 0)
    {
        cudaError_t e = cudaMemcpyAsync(h, src.p, sizeof(int32_t) * 4 * (size_t)n, cudaMemcpyDeviceToHost, c->st);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(c->st);
        if (e != cudaSuccess)
        {
            free(h);
            c->err = std::string("download: ") + cudaGetErrorString(e);
            return CCD_ERR_CUDA;
        }
    }
    *out = h;
    return CCD_OK;
}

static bool history_is_single_step(int V, const int64_t *hoff)
{
    for (int v = 0; v <= V; v++)
        if (hoff[v] != 2ll * v)
            return false;
    return true;
}

extern "C" {

int ccd_broadphase(ccd_context *c, int kind, int V, int F, const int32_t *faces, const int64_t *hoff, const double *htime,
                   const double *hpos, double outerEta, const uint8_t *fixedMask, int32_t **vf, int64_t *nvf, int32_t **ee, int64_t *nee)
{
    if (!c || !vf || !nvf || !ee || !nee || (F > 0 && !faces) || (V > 0 && (!hoff || !hpos)))
        return CCD_ERR_ARG;
    (void)htime;
    CK(cudaSetDevice(c->device));
    c->launches = 0;
    const long long N = V > 0 ? (long long)hoff[V] : 0;
    CKR(upload(c, c->faces, faces, sizeof(int32_t) * 3 * (size_t)F));
    CKR(upload(c, c->hoff, hoff, sizeof(int64_t) * ((size_t)V + 1)));
    CKR(upload(c, c->hpos, hpos, sizeof(double) * 3 * (size_t)N));
    if (fixedMask)
        CKR(upload(c, c->fixed, fixedMask, (size_t)V));
    BpResult r;
    CKR(broadphase_device(c, kind, V, F, P<int>(c->faces), nullptr, nullptr, P<long long>(c->hoff), P<double>(c->hpos), outerEta,
                          fixedMask ? P<unsigned char>(c->fixed) : nullptr, 0, 1, &r));
    CKR(download_stencils(c, c->vfOut, r.nvf, vf));
    CKR(download_stencils(c, c->eeOut, r.nee, ee));
    *nvf = r.nvf;
    *nee = r.nee;
    return CCD_OK;
}

int ccd_broadphase_step(ccd_context *c, int kind, int V, int F, const int32_t *faces, const double *q0, const double *q1, double outerEta,
                        const uint8_t *fixedMask, int32_t **vf, int64_t *nvf, int32_t **ee, int64_t *nee)
{
    if (!c || !vf || !nvf || !ee || !nee || (F > 0 && !faces) || (V > 0 && (!q0 || !q1)))
        return CCD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    c->launches = 0;
    CKR(upload(c, c->faces, faces, sizeof(int32_t) * 3 * (size_t)F));
    CKR(upload(c, c->q0, q0, sizeof(double) * 3 * (size_t)V));
    CKR(upload(c, c->q1, q1, sizeof(double) * 3 * (size_t)V));
    if (fixedMask)
        CKR(upload(c, c->fixed, fixedMask, (size_t)V));
    BpResult r;
    CKR(broadphase_device(c, kind, V, F, P<int>(c->faces), P<double>(c->q0), P<double>(c->q1), nullptr, nullptr, outerEta,
                          fixedMask ? P<unsigned char>(c->fixed) : nullptr, 0, 1, &r));
    CKR(download_stencils(c, c->vfOut, r.nvf, vf));
    CKR(download_stencils(c, c->eeOut, r.nee, ee));
    *nvf = r.nvf;
    *nee = r.nee;
    return CCD_OK;
}

int ccd_narrowphase(ccd_context *c, int V, const int64_t *hoff, const double *htime, const double *hpos, int64_t nvf, const int32_t *vf,
                    const double *vf_eta, int64_t nee, const int32_t *ee, const double *ee_eta, uint8_t *vf_hit, double *vf_toi,
                    uint8_t *vf_stage, uint8_t *ee_hit, double *ee_toi, uint8_t *ee_stage, ccd_np_summary *summary)
{
    if (!c || V < 0 || nvf < 0 || nee < 0 || (V > 0 && (!hoff || !htime || !hpos)) || (nvf > 0 && (!vf || !vf_eta || !vf_hit)) ||
        (nee > 0 && (!ee || !ee_eta || !ee_hit)))
        return CCD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    c->launches = 0;
    const long long N = V > 0 ? (long long)hoff[V] : 0;
    CKR(upload(c, c->hoff, hoff, sizeof(int64_t) * ((size_t)V + 1)));
    CKR(upload(c, c->htime, htime, sizeof(double) * (size_t)N));
    CKR(upload(c, c->hpos, hpos, sizeof(double) * 3 * (size_t)N));
    CKR(upload(c, c->vf_in, vf, sizeof(int32_t) * 4 * (size_t)nvf));
    CKR(upload(c, c->ee_in, ee, sizeof(int32_t) * 4 * (size_t)nee));
    // one thickness for every stencil of a type (the common case: include/CTCD.h callers pass one eta) is passed by value:
    // no per-stencil array to upload or read, and the vertex-edge de-duplication applies (it needs a shared eta)
    auto uniform = [](const double *a, int64_t n) {
        if (n <= 0) return false;
        for (int64_t i = 1; i < n; i++)
            if (a[i] != a[0]) return false;
        return true;
    };
    const bool uvf = uniform(vf_eta, nvf), uee = uniform(ee_eta, nee);
    if (!uvf) CKR(upload(c, c->vf_eta, vf_eta, sizeof(double) * (size_t)nvf));
    if (!uee) CKR(upload(c, c->ee_eta, ee_eta, sizeof(double) * (size_t)nee));
    ccd_np_summary s;
    // two entries per vertex = one linear segment: read q0/q1 straight out of hpos (stride 6)
    const bool single = history_is_single_step(V, hoff);
    // a few thousand stencils (Model1_flow frames, one VelocityFilter pass on a small mesh): the ~90 launches of the dense
    // pipeline cost more than the work; one thread per stencil runs the whole sequence in a single kernel (same bits)
    const bool tiny = nvf + nee <= 8192 && !getenv("CCD_NP_NO_SMALL");
    if (single && tiny)
        CKR(narrowphase_device(c, V, nvf, P<int>(c->vf_in), uvf ? nullptr : P<double>(c->vf_eta), nee, P<int>(c->ee_in), uee ? nullptr : P<double>(c->ee_eta),
                               uvf ? vf_eta[0] : 0.0, uee ? ee_eta[0] : 0.0, nullptr, nullptr, 0, P<long long>(c->hoff), P<double>(c->htime),
                               P<double>(c->hpos), &s));
    else if (single)
        CKR(narrowphase_device(c, V, nvf, P<int>(c->vf_in), uvf ? nullptr : P<double>(c->vf_eta), nee, P<int>(c->ee_in), uee ? nullptr : P<double>(c->ee_eta),
                               uvf ? vf_eta[0] : 0.0, uee ? ee_eta[0] : 0.0, P<double>(c->hpos), P<double>(c->hpos) + 3, 6, P<long long>(c->hoff),
                               P<double>(c->htime), P<double>(c->hpos), &s));
    else
        CKR(narrowphase_history_device(c, V, nvf, P<int>(c->vf_in), uvf ? nullptr : P<double>(c->vf_eta), nee, P<int>(c->ee_in),
                                       uee ? nullptr : P<double>(c->ee_eta), uvf ? vf_eta[0] : 0.0, uee ? ee_eta[0] : 0.0, P<long long>(c->hoff),
                                       P<double>(c->htime), P<double>(c->hpos), &s));
    if (nvf > 0)
    {
        CK(cudaMemcpyAsync(vf_hit, c->vfHit.p, (size_t)nvf, cudaMemcpyDeviceToHost, c->st));
        if (vf_toi) CK(cudaMemcpyAsync(vf_toi, c->vfToi.p, sizeof(double) * (size_t)nvf, cudaMemcpyDeviceToHost, c->st));
        if (vf_stage) CK(cudaMemcpyAsync(vf_stage, c->vfStage.p, (size_t)nvf, cudaMemcpyDeviceToHost, c->st));
    }
    if (nee > 0)
    {
        CK(cudaMemcpyAsync(ee_hit, c->eeHit.p, (size_t)nee, cudaMemcpyDeviceToHost, c->st));
        if (ee_toi) CK(cudaMemcpyAsync(ee_toi, c->eeToi.p, sizeof(double) * (size_t)nee, cudaMemcpyDeviceToHost, c->st));
        if (ee_stage) CK(cudaMemcpyAsync(ee_stage, c->eeStage.p, (size_t)nee, cudaMemcpyDeviceToHost, c->st));
    }
    CK(cudaStreamSynchronize(c->st));
    if (summary)
        *summary = s;
    return CCD_OK;
}

// Staged SeparatingPlaneNarrowPhase (narrowphase.cu): rounds over a global queue of open intervals, leaves through the
// single-step pipeline, flags OR-ed per stencil.  Inputs are the uploaded c->vf_in / ee_in / vf_eta / ee_eta / hoff / htime /
// hpos; per-stencil flags end in c->histHit (vertex-face first), hit counts in the C_NHIT_* counters.
#define SP_MAX_ROUNDS 192      // interval lengths at least halve per round; a search that straddles a History breakpoint runs down to ~2^-52
#define SP_BATCH 12            // rounds launched between two looks at the queue counters
#define SP_DIRECT_LEAVES 32768 // up to this many leaves the CTCD sequence runs in one kernel per type instead of the dense pipeline
static int sepplane_staged(ccd_context *c, long long nvf, long long nee, double eps)
{
    unsigned long long *ctr = P<unsigned long long>(c->counters);
    for (int attempt = 0; attempt < 8; attempt++)
    {
        if (c->spQCap < (size_t)(2 * (nvf + nee) + 4096)) c->spQCap = (size_t)(2 * (nvf + nee) + 4096);
        if (c->spLeafCapVf < (size_t)(2 * nvf + 2048)) c->spLeafCapVf = (size_t)(2 * nvf + 2048);
        if (c->spLeafCapEe < (size_t)(2 * nee + 2048)) c->spLeafCapEe = (size_t)(2 * nee + 2048);
        const size_t qcap = c->spQCap, lcv = c->spLeafCapVf, lce = c->spLeafCapEe, lcap = lcv + lce;
        CKR(ensure(c, c->spCtr, sizeof(unsigned long long) * (SP_MAX_ROUNDS + 2 + 8)));
        CKR(ensure(c, c->spQst, sizeof(int) * 2 * qcap));
        CKR(ensure(c, c->spQlo, sizeof(double) * 2 * qcap));
        CKR(ensure(c, c->spQhi, sizeof(double) * 2 * qcap));
        CKR(ensure(c, c->histQ0, sizeof(double) * 12 * (lcap + 1)));
        CKR(ensure(c, c->histQ1, sizeof(double) * 12 * (lcap + 1)));
        CKR(ensure(c, c->histVst, sizeof(int) * 4 * (lcap + 1)));
        CKR(ensure(c, c->histEta, sizeof(double) * (lcap + 1)));
        CKR(ensure(c, c->spLeafSt, sizeof(int) * (lcap + 1)));
        CKR(ensure(c, c->histHit, (size_t)(nvf + nee) + 16));
        unsigned long long *rc = P<unsigned long long>(c->spCtr);      // rc[r] = intervals entering round r
        CK(cudaMemsetAsync(rc, 0, sizeof(unsigned long long) * (SP_MAX_ROUNDS + 2 + 8), c->st));
        CK(cudaMemsetAsync(c->histHit.p, 0, (size_t)(nvf + nee) + 16, c->st));
        unsigned long long *leafCtr = rc + SP_MAX_ROUNDS + 2;
        unsigned char *hitVf = P<unsigned char>(c->histHit), *hitEe = hitVf + nvf;
        int *qst = P<int>(c->spQst);
        double *qlo = P<double>(c->spQlo), *qhi = P<double>(c->spQhi);
        std::vector<unsigned long long> h(SP_MAX_ROUNDS + 2 + 8);
        bool overflow = false, done = false;
        int r = 0;
        while (!done && r < SP_MAX_ROUNDS)
        {
            for (int b = 0; b < SP_BATCH && r < SP_MAX_ROUNDS; b++, r++)
            {
                const size_t in = (size_t)((r + 1) & 1) * qcap, out = (size_t)(r & 1) * qcap;
                ccdk_sp_round(c->st, r == 0, nvf, nee, P<int>(c->vf_in), P<int>(c->ee_in), P<double>(c->vf_eta), P<double>(c->ee_eta), P<long long>(c->hoff),
                              P<double>(c->htime), P<double>(c->hpos), eps, qst + in, qlo + in, qhi + in, rc + r, qst + out, qlo + out, qhi + out, rc + r + 1,
                              qcap, hitVf, hitEe, leafCtr, lcv, lce, P<double>(c->histQ0), P<double>(c->histQ1), P<int>(c->histVst), P<double>(c->histEta),
                              P<int>(c->spLeafSt));
                c->launches += 1;
            }
            CK(cudaMemcpyAsync(h.data(), rc, sizeof(unsigned long long) * (SP_MAX_ROUNDS + 2 + 8), cudaMemcpyDeviceToHost, c->st));
            CK(cudaStreamSynchronize(c->st));
            for (int k = 1; k <= r; k++)
                if (h[k] > qcap) { overflow = true; if (h[k] + h[k] / 2 > c->spQCap) c->spQCap = (size_t)(h[k] + h[k] / 2 + 4096); }
            const unsigned long long lv = h[SP_MAX_ROUNDS + 2], le = h[SP_MAX_ROUNDS + 3];
            if (lv > lcv) { overflow = true; c->spLeafCapVf = (size_t)(lv + lv / 2 + 2048); }
            if (le > lce) { overflow = true; c->spLeafCapEe = (size_t)(le + le / 2 + 2048); }
            done = overflow || h[r] == 0;
        }
        if (getenv("CCD_SP_TRACE"))
        {
            fprintf(stderr, "[sp] rounds %d, leaves %llu + %llu, intervals per round:", r, h[SP_MAX_ROUNDS + 2], h[SP_MAX_ROUNDS + 3]);
            for (int k = 1; k <= r && (k == 1 || h[k - 1]); k++) fprintf(stderr, " %llu", h[k]);
            fprintf(stderr, "\n");
        }
        if (overflow) continue;
        if (!done)
        {
            c->err = "narrowphase_sepplane: interval search did not finish (History with extremely small time gaps)";
            return CCD_ERR_NOMEM;
        }
        // the leaves: CTCD on short linear pieces
        const long long nlv = (long long)h[SP_MAX_ROUNDS + 2], nle = (long long)h[SP_MAX_ROUNDS + 3];
        if (nlv + nle > 0 && nlv + nle <= SP_DIRECT_LEAVES)
        {
            ccdk_sp_leaf_direct(c->st, true, nlv, 0, P<double>(c->histQ0), P<double>(c->histQ1), P<double>(c->histEta), P<int>(c->spLeafSt), hitVf);
            ccdk_sp_leaf_direct(c->st, false, nle, (long long)lcv, P<double>(c->histQ0), P<double>(c->histQ1), P<double>(c->histEta), P<int>(c->spLeafSt), hitEe);
            c->launches += 2;
        }
        else if (nlv + nle > 0)
        {
            // through the dense single-step pipeline as virtual stencils (vertex-face leaves in [0, nlv), edge-edge leaves from lcv on)
            if (4 * ((long long)lcv + nle) >= (1ll << 31) - 8) { c->err = "narrowphase_sepplane: too many leaf intervals in one call"; return CCD_ERR_NOMEM; }
            ccd_np_summary virt;
            CKR(narrowphase_device(c, (int)(4 * ((long long)lcv + nle)), nlv, P<int>(c->histVst), P<double>(c->histEta), nle, P<int>(c->histVst) + 4 * lcv,
                                   P<double>(c->histEta) + lcv, 0.0, 0.0, P<double>(c->histQ0), P<double>(c->histQ1), 3, nullptr, nullptr, nullptr, &virt));
            ccdk_sp_leaf_or(c->st, nlv, P<int>(c->spLeafSt), P<unsigned char>(c->vfHit), hitVf);
            ccdk_sp_leaf_or(c->st, nle, P<int>(c->spLeafSt) + lcv, P<unsigned char>(c->eeHit), hitEe);
            c->launches += 2;
        }
        CK(cudaMemsetAsync(ctr + C_EARLY_VF, 0, sizeof(unsigned long long) * 4, c->st));
        ccdk_count_flags(c->st, nvf, hitVf, ctr + C_NHIT_VF);
        ccdk_count_flags(c->st, nee, hitEe, ctr + C_NHIT_EE);
        c->launches += 2;
        CK(cudaGetLastError());
        return CCD_OK;
    }
    c->err = "narrowphase_sepplane: interval queues kept overflowing";
    return CCD_ERR_NOMEM;
}

// SeparatingPlaneNarrowPhase::findCollisions (src/SeparatingPlaneNarrowPhase.cpp:11-25)
int ccd_narrowphase_sepplane(ccd_context *c, int V, const int64_t *hoff, const double *htime, const double *hpos, int64_t nvf, const int32_t *vf,
                             const double *vf_eta, int64_t nee, const int32_t *ee, const double *ee_eta, uint8_t *vf_hit, uint8_t *ee_hit,
                             int64_t *n_vf_hits, int64_t *n_ee_hits)
{
    if (!c || V < 0 || nvf < 0 || nee < 0 || (V > 0 && (!hoff || !htime || !hpos)) || (nvf > 0 && (!vf || !vf_eta || !vf_hit)) ||
        (nee > 0 && (!ee || !ee_eta || !ee_hit)))
        return CCD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    c->launches = 0;
    const long long N = V > 0 ? (long long)hoff[V] : 0;
    // History::computeMinimumGap (src/History.cpp:82-96) / 4
    double gap = 1.0;
    for (int v = 0; v < V; v++)
        for (long long j = hoff[v] + 1; j < hoff[v + 1]; j++)
        {
            const double g = htime[j] - htime[j - 1];
            gap = g < gap ? g : gap;
        }
    const double eps = gap / 4.0;
    CKR(upload(c, c->hoff, hoff, sizeof(int64_t) * ((size_t)V + 1)));
    CKR(upload(c, c->htime, htime, sizeof(double) * (size_t)N));
    CKR(upload(c, c->hpos, hpos, sizeof(double) * 3 * (size_t)N));
    CKR(upload(c, c->vf_in, vf, sizeof(int32_t) * 4 * (size_t)nvf));
    CKR(upload(c, c->ee_in, ee, sizeof(int32_t) * 4 * (size_t)nee));
    CKR(upload(c, c->vf_eta, vf_eta, sizeof(double) * (size_t)nvf));
    CKR(upload(c, c->ee_eta, ee_eta, sizeof(double) * (size_t)nee));
    if (!getenv("CCD_SEPPLANE_ONE_THREAD"))
    {
        // staged: interval queue across all stencils (CCD_SEPPLANE_ONE_THREAD=1 keeps the one-thread-per-stencil kernel; same flags)
        CKR(sepplane_staged(c, nvf, nee, eps));
        if (nvf > 0) CK(cudaMemcpyAsync(vf_hit, c->histHit.p, (size_t)nvf, cudaMemcpyDeviceToHost, c->st));
        if (nee > 0) CK(cudaMemcpyAsync(ee_hit, P<unsigned char>(c->histHit) + nvf, (size_t)nee, cudaMemcpyDeviceToHost, c->st));
        CKR(sync_counters(c));
        if (n_vf_hits) *n_vf_hits = (int64_t)c->h_counters[C_NHIT_VF];
        if (n_ee_hits) *n_ee_hits = (int64_t)c->h_counters[C_NHIT_EE];
        return CCD_OK;
    }
    CKR(ensure(c, c->vfHit, (size_t)nvf + 16));
    CKR(ensure(c, c->eeHit, (size_t)nee + 16));
    unsigned long long *ctr = P<unsigned long long>(c->counters);
    CK(cudaMemsetAsync(ctr + C_EARLY_VF, 0, sizeof(unsigned long long) * 4, c->st));
    ccdk_sepplane(c->st, true, nvf, P<int>(c->vf_in), P<double>(c->vf_eta), P<long long>(c->hoff), P<double>(c->htime), P<double>(c->hpos), eps,
                  P<unsigned char>(c->vfHit), ctr + C_NHIT_VF, ctr + C_EARLY_VF);
    ccdk_sepplane(c->st, false, nee, P<int>(c->ee_in), P<double>(c->ee_eta), P<long long>(c->hoff), P<double>(c->htime), P<double>(c->hpos), eps,
                  P<unsigned char>(c->eeHit), ctr + C_NHIT_EE, ctr + C_EARLY_EE);
    c->launches += 2;
    CK(cudaGetLastError());
    if (nvf > 0) CK(cudaMemcpyAsync(vf_hit, c->vfHit.p, (size_t)nvf, cudaMemcpyDeviceToHost, c->st));
    if (nee > 0) CK(cudaMemcpyAsync(ee_hit, c->eeHit.p, (size_t)nee, cudaMemcpyDeviceToHost, c->st));
    CKR(sync_counters(c));
    if (c->h_counters[C_EARLY_VF] || c->h_counters[C_EARLY_EE])
    {
        c->err = "narrowphase_sepplane: interval stack overflow (History with extremely small time gaps)";
        return CCD_ERR_NOMEM;
    }
    if (n_vf_hits) *n_vf_hits = (int64_t)c->h_counters[C_NHIT_VF];
    if (n_ee_hits) *n_ee_hits = (int64_t)c->h_counters[C_NHIT_EE];
    return CCD_OK;
}

int ccd_step_device(ccd_context *c, int kind, int V, int F, const int32_t *d_faces, const double *d_q0, const double *d_q1, double outerEta,
                    double eta, const uint8_t *d_fixedMask, int shard_rank, int shard_world, ccd_device_result *out)
{
    if (!c || !out)
        return CCD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    c->launches = 0;
    memset(out, 0, sizeof(*out));
    BpResult r;
    CK(cudaEventRecord(c->ev[0], c->st));
    CKR(broadphase_device(c, kind, V, F, d_faces, d_q0, d_q1, nullptr, nullptr, outerEta, d_fixedMask, shard_rank, shard_world, &r));
    CK(cudaEventRecord(c->ev[1], c->st));
    ccd_np_summary s;
    CKR(narrowphase_device(c, V, r.nvf, P<int>(c->vfOut), nullptr, r.nee, P<int>(c->eeOut), nullptr, eta, eta, d_q0, d_q1, 3, nullptr, nullptr, nullptr, &s));
    CK(cudaEventRecord(c->ev[2], c->st));
    CK(cudaEventSynchronize(c->ev[2]));
    cudaEventElapsedTime(&out->ms_broadphase, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&out->ms_narrowphase, c->ev[1], c->ev[2]);
    out->n_vf_candidates = r.nvf;
    out->n_ee_candidates = r.nee;
    out->n_vf_hits = s.n_vf_hits;
    out->n_ee_hits = s.n_ee_hits;
    out->earliest_toi = s.earliest_toi;
    out->d_vf = P<int32_t>(c->vfOut);
    out->d_ee = P<int32_t>(c->eeOut);
    out->d_vf_hit = P<uint8_t>(c->vfHit);
    out->d_ee_hit = P<uint8_t>(c->eeHit);
    out->d_vf_toi = P<double>(c->vfToi);
    out->d_ee_toi = P<double>(c->eeToi);
    out->n_face_pairs = r.npairs;
    out->n_tree_candidates = r.ncand;
    out->n_launches = c->launches;
    out->n_vf_deferred = (int64_t)c->h_counters[C_NWORK_VF];
    out->n_ee_deferred = (int64_t)c->h_counters[C_NWORK_EE];
    c->stage_valid = (F > 0 && V > 0);
    return CCD_OK;
}

static int ensure_pinned(ccd_context *c, int slot, size_t bytes)
{
    if (bytes <= c->h_res_cap[slot] && c->h_res[slot])
        return CCD_OK;
    if (c->h_res[slot])
        cudaFreeHost(c->h_res[slot]);
    c->h_res[slot] = nullptr;
    c->h_res_cap[slot] = 0;
    const size_t ncap = bytes + bytes / 4 + 4096;
    if (cudaMallocHost(&c->h_res[slot], ncap) != cudaSuccess)
    {
        cudaGetLastError();
        c->err = "cudaMallocHost failed";
        return CCD_ERR_NOMEM;
    }
    c->h_res_cap[slot] = ncap;
    return CCD_OK;
}

// hit compaction in std::set order (stable) into the context's pinned host buffers `slot`, `slot+1`
static int select_hits(ccd_context *c, int slot, long long n, const int *d_st, const double *d_toi, const unsigned char *d_flag, long long nh,
                       int32_t **h_st, double **h_toi)
{
    CKR(ensure_pinned(c, slot, sizeof(int32_t) * 4 * (size_t)(nh > 0 ? nh : 1)));
    CKR(ensure_pinned(c, slot + 1, sizeof(double) * (size_t)(nh > 0 ? nh : 1)));
    *h_st = (int32_t *)c->h_res[slot];
    *h_toi = (double *)c->h_res[slot + 1];
    if (n == 0 || nh == 0)
        return CCD_OK;
    size_t tb1 = 0, tb2 = 0;
    cub::DeviceSelect::Flagged(nullptr, tb1, (const int4 *)nullptr, (const unsigned char *)nullptr, (int4 *)nullptr, (int *)nullptr, (int)n);
    cub::DeviceSelect::Flagged(nullptr, tb2, (const double *)nullptr, (const unsigned char *)nullptr, (double *)nullptr, (int *)nullptr, (int)n);
    CKR(ensure(c, c->selTmp, (tb1 > tb2 ? tb1 : tb2) + 256));
    CKR(ensure(c, c->selA, sizeof(int) * 4 * (size_t)(nh + 1)));
    CKR(ensure(c, c->selB, sizeof(double) * (size_t)(nh + 1)));
    CKR(ensure(c, c->selCount, 64));
    cub::DeviceSelect::Flagged(c->selTmp.p, tb1, (const int4 *)d_st, d_flag, (int4 *)c->selA.p, (int *)c->selCount.p, (int)n, c->st);
    cub::DeviceSelect::Flagged(c->selTmp.p, tb2, d_toi, d_flag, (double *)c->selB.p, (int *)c->selCount.p, (int)n, c->st);
    c->launches += 4;
    CK(cudaMemcpyAsync(*h_st, c->selA.p, sizeof(int32_t) * 4 * (size_t)nh, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(*h_toi, c->selB.p, sizeof(double) * (size_t)nh, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return CCD_OK;
}

int ccd_step(ccd_context *c, int kind, int V, int F, const int32_t *faces, const double *q0, const double *q1, double outerEta, double eta,
             const uint8_t *fixedMask, ccd_step_result *out)
{
    return ccd_step_shard(c, kind, V, F, faces, q0, q1, outerEta, eta, fixedMask, 0, 1, out);
}

int ccd_step_shard(ccd_context *c, int kind, int V, int F, const int32_t *faces, const double *q0, const double *q1, double outerEta,
                   double eta, const uint8_t *fixedMask, int shard_rank, int shard_world, ccd_step_result *out)
{
    if (!c || !out || (F > 0 && !faces) || (V > 0 && (!q0 || !q1)))
        return CCD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    memset(out, 0, sizeof(*out));
    CKR(upload(c, c->faces, faces, sizeof(int32_t) * 3 * (size_t)F));
    CKR(upload(c, c->q0, q0, sizeof(double) * 3 * (size_t)V));
    CKR(upload(c, c->q1, q1, sizeof(double) * 3 * (size_t)V));
    if (fixedMask)
        CKR(upload(c, c->fixed, fixedMask, (size_t)V));
    ccd_device_result d;
    CKR(ccd_step_device(c, kind, V, F, P<int>(c->faces), P<double>(c->q0), P<double>(c->q1), outerEta, eta,
                        fixedMask ? P<unsigned char>(c->fixed) : nullptr, shard_rank, shard_world, &d));
    out->n_vf_candidates = d.n_vf_candidates;
    out->n_ee_candidates = d.n_ee_candidates;
    out->n_vf_hits = d.n_vf_hits;
    out->n_ee_hits = d.n_ee_hits;
    out->earliest_toi = d.earliest_toi;
    out->ms_broadphase = d.ms_broadphase;
    out->ms_narrowphase = d.ms_narrowphase;
    CKR(select_hits(c, 0, d.n_vf_candidates, d.d_vf, d.d_vf_toi, d.d_vf_hit, d.n_vf_hits, &out->vf_hits, &out->vf_hit_toi));
    CKR(select_hits(c, 2, d.n_ee_candidates, d.d_ee, d.d_ee_toi, d.d_ee_hit, d.n_ee_hits, &out->ee_hits, &out->ee_hit_toi));
    return CCD_OK;
}

int ccd_step_device_hits(ccd_context *c, int kind, int V, int F, const int32_t *d_faces, const double *d_q0, const double *d_q1, double outerEta,
                         double eta, const uint8_t *d_fixedMask, int shard_rank, int shard_world, ccd_step_result *out)
{
    if (!c || !out)
        return CCD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    memset(out, 0, sizeof(*out));
    ccd_device_result d;
    CKR(ccd_step_device(c, kind, V, F, d_faces, d_q0, d_q1, outerEta, eta, d_fixedMask, shard_rank, shard_world, &d));
    out->n_vf_candidates = d.n_vf_candidates;
    out->n_ee_candidates = d.n_ee_candidates;
    out->n_vf_hits = d.n_vf_hits;
    out->n_ee_hits = d.n_ee_hits;
    out->earliest_toi = d.earliest_toi;
    out->ms_broadphase = d.ms_broadphase;
    out->ms_narrowphase = d.ms_narrowphase;
    CKR(select_hits(c, 0, d.n_vf_candidates, d.d_vf, d.d_vf_toi, d.d_vf_hit, d.n_vf_hits, &out->vf_hits, &out->vf_hit_toi));
    CKR(select_hits(c, 2, d.n_ee_candidates, d.d_ee, d.d_ee_toi, d.d_ee_hit, d.n_ee_hits, &out->ee_hits, &out->ee_hit_toi));
    return CCD_OK;
}

int ccd_wait_stream(ccd_context *c, void *producer_stream)
{
    if (!c)
        return CCD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    // everything enqueued so far on the caller's stream (the copy / broadcast that fills the device inputs) is ordered before
    // everything this context enqueues from now on; no host synchronisation
    CK(cudaEventRecord(c->evFork, (cudaStream_t)producer_stream));
    CK(cudaStreamWaitEvent(c->st, c->evFork, 0));
    return CCD_OK;
}

void ccd_step_result_free(ccd_step_result *r)
{
    if (!r)
        return;
    // the arrays live in pinned buffers owned by the context: nothing to release, just forget them
    r->vf_hits = r->ee_hits = nullptr;
    r->vf_hit_toi = r->ee_hit_toi = nullptr;
}

int ccd_set_shard_partition(ccd_context *c, int world, const int32_t *pbounds)
{
    if (!c || world < 1 || !pbounds || pbounds[0] != 0)
        return CCD_ERR_ARG;
    for (int r = 0; r < world; r++)
        if (pbounds[r + 1] < pbounds[r])
            return CCD_ERR_ARG;
    c->partP.assign(pbounds, pbounds + world + 1);
    return CCD_OK;
}

int ccd_shard_histogram(ccd_context *c, int64_t *vf_hist, int64_t *ee_hist, int32_t *n_positions)
{
    if (!c || !vf_hist || !ee_hist)
        return CCD_ERR_ARG;
    if (!c->hist_valid)
    {
        c->err = "ccd_shard_histogram: the last step on this context was not sharded";
        return CCD_ERR_ARG;
    }
    for (int i = 0; i < CCD_SHARD_BUCKETS; i++)
    {
        vf_hist[i] = (int64_t)c->h_hist[i];
        ee_hist[i] = (int64_t)c->h_hist[CCD_SHARD_BUCKETS + i];
    }
    if (n_positions) *n_positions = c->histF;
    return CCD_OK;
}

int ccd_stage_times(ccd_context *c, float *ms, int n)
{
    if (!c || !ms || n < CCD_N_STAGES)
        return CCD_ERR_ARG;
    for (int i = 0; i < n; i++)
        ms[i] = 0.f;
    if (!c->stage_valid)
        return CCD_OK;
    CK(cudaSetDevice(c->device));
    CK(cudaEventSynchronize(c->sev[CCD_N_STAGES]));
    for (int i = 0; i < CCD_N_STAGES; i++)
        cudaEventElapsedTime(&ms[i], c->sev[i], c->sev[i + 1]);
    cudaGetLastError();
    return CCD_OK;
}

int ccd_memcpy_d2h(ccd_context *c, void *h_dst, const void *d_src, uint64_t bytes)
{
    if (!c || (bytes && (!h_dst || !d_src)))
        return CCD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    if (bytes)
    {
        CK(cudaMemcpyAsync(h_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
    }
    return CCD_OK;
}

int ccd_fp64_peak(ccd_context *c, double *tflops)
{
    if (!c || !tflops)
        return CCD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    CKR(ensure(c, c->selD, sizeof(double) * 148 * 8 * 256));
    const int iters = 4096;
    double best = 0;
    for (int rep = 0; rep < 5; rep++)
    {
        CK(cudaEventRecord(c->ev[0], c->st));
        ccdk_fp64_peak(c->st, iters, P<double>(c->selD));
        CK(cudaEventRecord(c->ev[1], c->st));
        CK(cudaEventSynchronize(c->ev[1]));
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
        double flops = 2.0 * 16.0 * (double)iters * 148.0 * 8.0 * 256.0;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best)
            best = tf;
    }
    CK(cudaGetLastError());
    *tflops = best;
    return CCD_OK;
}

static int prim_batch(ccd_context *c, int kind, int npts, int64_t n, const double *pts, const double *eta, uint8_t *hit, double *t)
{
    if (!c || n < 0 || (n > 0 && (!pts || !eta || !hit || !t)))
        return CCD_ERR_ARG;
    if (n == 0)
        return CCD_OK;
    CK(cudaSetDevice(c->device));
    CKR(upload(c, c->pts, pts, sizeof(double) * 3 * (size_t)npts * (size_t)n));
    CKR(upload(c, c->eta, eta, sizeof(double) * (size_t)n));
    CKR(ensure(c, c->vfHit, (size_t)n + 16));
    CKR(ensure(c, c->vfToi, sizeof(double) * ((size_t)n + 2)));
    // t is only written on a hit: start from the caller's values
    CK(cudaMemcpyAsync(c->vfToi.p, t, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->st));
    ccdk_prim_batch(c->st, kind, n, P<double>(c->pts), P<double>(c->eta), P<unsigned char>(c->vfHit), P<double>(c->vfToi));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(hit, c->vfHit.p, (size_t)n, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(t, c->vfToi.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return CCD_OK;
}

int ccd_vf_batch(ccd_context *c, int64_t n, const double *pts, const double *eta, uint8_t *hit, double *t) { return prim_batch(c, 0, 8, n, pts, eta, hit, t); }
int ccd_ee_batch(ccd_context *c, int64_t n, const double *pts, const double *eta, uint8_t *hit, double *t) { return prim_batch(c, 1, 8, n, pts, eta, hit, t); }
int ccd_ve_batch(ccd_context *c, int64_t n, const double *pts, const double *eta, uint8_t *hit, double *t) { return prim_batch(c, 2, 6, n, pts, eta, hit, t); }
int ccd_vv_batch(ccd_context *c, int64_t n, const double *pts, const double *eta, uint8_t *hit, double *t) { return prim_batch(c, 3, 4, n, pts, eta, hit, t); }

int ccd_find_intervals_batch(ccd_context *c, int64_t n, int degree, int pos, const double *coeffs, int32_t *cnt, double *lo, double *hi)
{
    if (!c || n < 0 || !(degree == 2 || degree == 3 || degree == 4 || degree == 6) || (n > 0 && (!coeffs || !cnt || !lo || !hi)))
        return CCD_ERR_ARG;
    if (n == 0)
        return CCD_OK;
    CK(cudaSetDevice(c->device));
    CKR(upload(c, c->pts, coeffs, sizeof(double) * 7 * (size_t)n));
    CKR(ensure(c, c->selA, sizeof(int) * (size_t)n));
    CKR(ensure(c, c->selB, sizeof(double) * 7 * (size_t)n));
    CKR(ensure(c, c->selC, sizeof(double) * 7 * (size_t)n));
    CK(cudaMemsetAsync(c->selB.p, 0, sizeof(double) * 7 * (size_t)n, c->st));
    CK(cudaMemsetAsync(c->selC.p, 0, sizeof(double) * 7 * (size_t)n, c->st));
    ccdk_find_intervals(c->st, n, degree, pos, P<double>(c->pts), P<int>(c->selA), P<double>(c->selB), P<double>(c->selC));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(cnt, c->selA.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(lo, c->selB.p, sizeof(double) * 7 * (size_t)n, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(hi, c->selC.p, sizeof(double) * 7 * (size_t)n, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return CCD_OK;
}

// ---- Distance.h ------------------------------------------------------------------------------
static int dist_batch(ccd_context *c, int which, int64_t n, const double *pts, const double *eta, double *vec, double *bary, int nb, uint8_t *flag)
{
    if (!c || n < 0 || (n > 0 && !pts))
        return CCD_ERR_ARG;
    if (n == 0)
        return CCD_OK;
    CK(cudaSetDevice(c->device));
    CKR(upload(c, c->pts, pts, sizeof(double) * 12 * (size_t)n));
    if (eta)
        CKR(upload(c, c->eta, eta, sizeof(double) * (size_t)n));
    CKR(ensure(c, c->selB, sizeof(double) * 3 * (size_t)n));
    CKR(ensure(c, c->selC, sizeof(double) * 4 * (size_t)n));
    CKR(ensure(c, c->vfHit, (size_t)n + 16));
    ccdk_dist_batch(c->st, which, n, P<double>(c->pts), eta ? P<double>(c->eta) : nullptr, P<double>(c->selB), P<double>(c->selC), P<unsigned char>(c->vfHit));
    CK(cudaGetLastError());
    if (vec) CK(cudaMemcpyAsync(vec, c->selB.p, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost, c->st));
    if (bary) CK(cudaMemcpyAsync(bary, c->selC.p, sizeof(double) * (size_t)nb * (size_t)n, cudaMemcpyDeviceToHost, c->st));
    if (flag) CK(cudaMemcpyAsync(flag, c->vfHit.p, (size_t)n, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return CCD_OK;
}

int ccd_dist_vf_batch(ccd_context *c, int64_t n, const double *pts, double *vec, double *bary)
{
    if (n > 0 && (!vec || !bary)) return CCD_ERR_ARG;
    return dist_batch(c, 0, n, pts, nullptr, vec, bary, 3, nullptr);
}
int ccd_dist_ee_batch(ccd_context *c, int64_t n, const double *pts, double *vec, double *bary)
{
    if (n > 0 && (!vec || !bary)) return CCD_ERR_ARG;
    return dist_batch(c, 1, n, pts, nullptr, vec, bary, 4, nullptr);
}
int ccd_dist_plane_lt_batch(ccd_context *c, int64_t n, const double *pts, const double *eta, uint8_t *out)
{
    if (n > 0 && (!eta || !out)) return CCD_ERR_ARG;
    return dist_batch(c, 2, n, pts, eta, nullptr, nullptr, 0, out);
}
int ccd_dist_line_lt_batch(ccd_context *c, int64_t n, const double *pts, const double *eta, uint8_t *out)
{
    if (n > 0 && (!eta || !out)) return CCD_ERR_ARG;
    return dist_batch(c, 3, n, pts, eta, nullptr, nullptr, 0, out);
}

// ---- PenaltyGroup::addForce, src/PenaltyGroup.cpp:34-52 -----------------------------------------
// device pointers: q, v, F (3 V doubles each), stencil lists, optional isnew flags, optional fired flags (nvf + nee bytes)
static int penalty_device(ccd_context *c, int V, const double *d_q, const double *d_v, int64_t nvf, const int *d_vf, const unsigned char *d_vf_isnew,
                          int64_t nee, const int *d_ee, const unsigned char *d_ee_isnew, double dt, double outerEta, double innerEta, double stiffness,
                          double CoR, double *d_F, unsigned char *d_fired, int64_t *n_fired, int *newused)
{
    const int64_t n = nvf + nee;
    if (4 * n >= (int64_t)1 << 31) return CCD_ERR_ARG;
    const size_t tb = n > 0 ? ccdk_penalty_temp_bytes(4 * n) : 0;
    CKR(ensure(c, c->temp, tb + 16));
    CKR(ensure(c, c->penGroup, sizeof(double) * 3 * (size_t)V + 16));
    CKR(ensure(c, c->penContrib, sizeof(double) * 12 * (size_t)n + 16));
    CKR(ensure(c, c->penKeysA, sizeof(unsigned) * 4 * (size_t)n + 16));
    CKR(ensure(c, c->penKeysB, sizeof(unsigned) * 4 * (size_t)n + 16));
    CKR(ensure(c, c->penItemsA, sizeof(unsigned) * 4 * (size_t)n + 16));
    CKR(ensure(c, c->penItemsB, sizeof(unsigned) * 4 * (size_t)n + 16));
    CKR(ensure(c, c->penCtr, 2 * sizeof(unsigned long long)));
    c->launches += ccdk_penalty_group_force(c->st, V, d_q, d_v, nvf, d_vf, d_vf_isnew, nee, d_ee, d_ee_isnew, dt, outerEta, innerEta, stiffness, CoR, d_F,
                                            d_fired, P<double>(c->penContrib), P<unsigned>(c->penKeysA), P<unsigned>(c->penKeysB), P<unsigned>(c->penItemsA),
                                            P<unsigned>(c->penItemsB), c->temp.p, tb, P<double>(c->penGroup), P<unsigned long long>(c->penCtr));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(c->h_counters, c->penCtr.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    if (n_fired) *n_fired = (int64_t)c->h_counters[0];
    if (newused) *newused = c->h_counters[1] != 0;
    return CCD_OK;
}

int ccd_penalty_group_force(ccd_context *c, int V, const double *q, const double *v, int64_t nvf, const int32_t *vf, const uint8_t *vf_isnew,
                            int64_t nee, const int32_t *ee, const uint8_t *ee_isnew, double dt, double outerEta, double innerEta, double stiffness,
                            double CoR, double *F, uint8_t *vf_fired, uint8_t *ee_fired, int64_t *n_fired, int *newused)
{
    if (!c || V < 0 || nvf < 0 || nee < 0 || (V > 0 && (!q || !v || !F)) || (nvf > 0 && !vf) || (nee > 0 && !ee))
        return CCD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    c->launches = 0;
    const size_t vb = sizeof(double) * 3 * (size_t)V;
    CKR(upload(c, c->q0, q, vb));
    CKR(upload(c, c->q1, v, vb));
    CKR(upload(c, c->penF, F, vb));
    CKR(upload(c, c->vf_in, vf, sizeof(int32_t) * 4 * (size_t)nvf));
    CKR(upload(c, c->ee_in, ee, sizeof(int32_t) * 4 * (size_t)nee));
    if (vf_isnew) CKR(upload(c, c->penNewVf, vf_isnew, (size_t)nvf));
    if (ee_isnew) CKR(upload(c, c->penNewEe, ee_isnew, (size_t)nee));
    CKR(ensure(c, c->penFired, (size_t)(nvf + nee) + 16));
    CKR(penalty_device(c, V, P<double>(c->q0), P<double>(c->q1), nvf, P<int>(c->vf_in), vf_isnew ? P<unsigned char>(c->penNewVf) : nullptr, nee,
                       P<int>(c->ee_in), ee_isnew ? P<unsigned char>(c->penNewEe) : nullptr, dt, outerEta, innerEta, stiffness, CoR, P<double>(c->penF),
                       P<unsigned char>(c->penFired), n_fired, newused));
    if (vb) CK(cudaMemcpyAsync(F, c->penF.p, vb, cudaMemcpyDeviceToHost, c->st));
    if (vf_fired && nvf) CK(cudaMemcpyAsync(vf_fired, c->penFired.p, (size_t)nvf, cudaMemcpyDeviceToHost, c->st));
    if (ee_fired && nee) CK(cudaMemcpyAsync(ee_fired, P<unsigned char>(c->penFired) + nvf, (size_t)nee, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return CCD_OK;
}

int ccd_penalty_group_force_device(ccd_context *c, int V, const double *d_q, const double *d_v, int64_t nvf, const int32_t *d_vf, const uint8_t *d_vf_isnew,
                                   int64_t nee, const int32_t *d_ee, const uint8_t *d_ee_isnew, double dt, double outerEta, double innerEta,
                                   double stiffness, double CoR, double *d_F, uint8_t *d_fired, int64_t *n_fired, int *newused)
{
    if (!c || V < 0 || nvf < 0 || nee < 0 || (V > 0 && (!d_q || !d_v || !d_F)) || (nvf > 0 && !d_vf) || (nee > 0 && !d_ee))
        return CCD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    c->launches = 0;
    return penalty_device(c, V, d_q, d_v, nvf, d_vf, d_vf_isnew, nee, d_ee, d_ee_isnew, dt, outerEta, innerEta, stiffness, CoR, d_F, d_fired, n_fired, newused);
}

// Distance::meshSelfDistance, src/Distance.cpp:12-66
int ccd_mesh_self_distance(ccd_context *c, int V, const double *verts, int F, const int32_t *faces, const uint8_t *fixedMask,
                           double *distance, int64_t *n_vf, int64_t *n_ee)
{
    if (!c || !distance || V < 0 || F < 0 || (V > 0 && !verts) || (F > 0 && !faces))
        return CCD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    c->launches = 0;
    CKR(upload(c, c->faces, faces, sizeof(int32_t) * 3 * (size_t)F));
    CKR(upload(c, c->q0, verts, sizeof(double) * 3 * (size_t)V));
    if (fixedMask)
        CKR(upload(c, c->fixed, fixedMask, (size_t)V));
    unsigned long long *ctr = P<unsigned long long>(c->counters);
    // (1) min squared distance from every vertex to the vertices of every face not containing it (:21-35)
    c->h_counters[8] = 0x7FF0000000000000ull; // +inf
    CK(cudaMemcpyAsync(ctr + C_EARLY_VF, c->h_counters + 8, sizeof(unsigned long long), cudaMemcpyHostToDevice, c->st));
    ccdk_vertex_min_dist2(c->st, V, F, P<double>(c->q0), P<int>(c->faces), ctr + C_EARLY_VF);
    CK(cudaGetLastError());
    CKR(sync_counters(c));
    double d2;
    memcpy(&d2, &c->h_counters[C_EARLY_VF], sizeof(d2));
    double closest = sqrt(d2);
    // (2) static history q -> q through the AABB broadphase with outerEta = closest (:37-45)
    BpResult r;
    CKR(broadphase_device(c, CCD_AABB, V, F, P<int>(c->faces), P<double>(c->q0), P<double>(c->q0), nullptr, nullptr, closest,
                          fixedMask ? P<unsigned char>(c->fixed) : nullptr, 0, 1, &r));
    // (3) min closest-point distance over the candidate stencils (:49-65)
    c->h_counters[8] = 0x7FF0000000000000ull;
    CK(cudaMemcpyAsync(ctr + C_EARLY_VF, c->h_counters + 8, sizeof(unsigned long long), cudaMemcpyHostToDevice, c->st));
    ccdk_stencil_min_dist(c->st, true, r.nvf, P<int>(c->vfOut), P<double>(c->q0), ctr + C_EARLY_VF);
    ccdk_stencil_min_dist(c->st, false, r.nee, P<int>(c->eeOut), P<double>(c->q0), ctr + C_EARLY_VF);
    CK(cudaGetLastError());
    CKR(sync_counters(c));
    memcpy(distance, &c->h_counters[C_EARLY_VF], sizeof(double));
    if (n_vf) *n_vf = r.nvf;
    if (n_ee) *n_ee = r.nee;
    return CCD_OK;
}

} // extern "C"
