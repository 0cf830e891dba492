// TEST INFRASTRUCTURE — host-side emulation of the single-step narrowphase pipeline.
//
// The stage kernels of collisiondetection_b200/csrc/narrowphase.cu are thin queue plumbing around per-item functions
// that are compiled for host and device (ccd_stencil.cuh, ccd_classify.cuh, ccd_solve.cuh, ccd_stages.cuh).  This file
// runs exactly those functions on the CPU, one stencil at a time, in the order the kernels apply them (cull -> stages ->
// export -> ve / vv -> decide -> solve -> combine -> general), so the pipeline's arithmetic and decision logic can be
// compared bit for bit with the CPU checker (oracle/) on a machine without a GPU.  It is never part of the product:
// nothing in collisiondetection_b200/ links or loads it.  Built by tests/test_np_emul.py with
//   nvcc -O2 --fmad=false -Xcompiler -ffp-contract=off,-fPIC -shared tests/np_emul.cu
// (x86 fma() from libm is the correctly rounded fused operation, like the device's).
#include "../collisiondetection_b200/csrc/ccd_stencil.cuh"
#include "../collisiondetection_b200/csrc/ccd_classify.cuh"
#include "../collisiondetection_b200/csrc/ccd_solve.cuh"
#include "../collisiondetection_b200/csrc/ccd_stages.cuh"
#include "../collisiondetection_b200/csrc/ccd_sepplane.cuh"
#include <vector>
#include <string.h>

using namespace ccd;

namespace {

struct Rec { double d[REC_STRIDE]; };
long long *g_stats = nullptr;      // funnel counters (np_emul's stats array)

template <int D> void solve_one(Rec &r)
{
    double c[D + 1];
    for (int k = 0; k <= D; k++) c[k] = r.d[k];
    const unsigned tag = rec_untag(r.d[7]);
    // prepare -> (start state through the record's scratch half, as the kernels pass it) -> climb -> finalize
    RootLane<D> P;
    P.prepare(c);
    P.save_start(r.d + 8);
    RootLane<D> L;
    const double aux[3] = {r.d[8], r.d[9], r.d[10]};
    L.load_start(c, aux);
    while (!L.done)
    {
        L.advance();
        while (L.solving) L.newton_step();
    }
    double roots[6];
    const int nr = L.result(roots);
    finalize_record<D>(c, tag, roots, nr, r.d);
}

void solve_records(std::vector<Rec> &recs, long long *stats, int phase)
{
    for (auto &r : recs)
    {
        const unsigned tag = rec_untag(r.d[7]);
        if (tag & REC_FINAL) continue;
        if (phase == 0 && (tag & REC_POS)) continue;
        const int rd = (int)((tag >> 4) & 7u);
        stats[4 + rd]++;
        if (rd == 3) solve_one<3>(r);
        else if (rd == 4) solve_one<4>(r);
        else if (rd == 5) solve_one<5>(r);
        else if (rd == 6) solve_one<6>(r);
    }
}

template <bool IS_VF, int S> bool run_stage(const V3 *a, const V3 *v, double eta, unsigned &state, Rec &own, bool &has)
{
    unsigned so = state;
    bool alive;
    if (S == Prim<IS_VF>::NST - 1) { double d8[8]; alive = stage_item<IS_VF, S, true>(a, v, eta, state, so, d8, has); memset(own.d, 0, sizeof(own.d)); memcpy(own.d, d8, sizeof(d8)); }
    else { double dummy[8]; bool h; alive = stage_item<IS_VF, S, false>(a, v, eta, state, so, dummy, h); }
    state = so;
    return alive;
}

template <bool IS_VF, int K> void run_export(const V3 *a, const V3 *v, double eta, Rec &r) { { double d8[8]; export_item<IS_VF, K>(a, v, eta, d8); memset(r.d, 0, sizeof(r.d)); memcpy(r.d, d8, sizeof(d8)); } }

// returns the status code of the primitive and appends its records
template <bool IS_VF> int primitive(const V3 *a, const V3 *v, double eta, std::vector<Rec> &recs, int &nrec)
{
    unsigned state = 0xff00u;
    Rec own;
    bool has = false;
    nrec = 0;
    if (!run_stage<IS_VF, 0>(a, v, eta, state, own, has)) return SC_MISS;
    if (g_stats) g_stats[20]++;
    if (!run_stage<IS_VF, 1>(a, v, eta, state, own, has)) return SC_MISS;
    if (g_stats) g_stats[21]++;
    if (!run_stage<IS_VF, 2>(a, v, eta, state, own, has)) return SC_MISS;
    if (g_stats) g_stats[22]++;
    if (!run_stage<IS_VF, 3>(a, v, eta, state, own, has)) return SC_MISS;
    if (g_stats) g_stats[23]++;
    if (!IS_VF)
        if (!run_stage<false, 4>(a, v, eta, state, own, has)) return SC_MISS;
    if (g_stats) g_stats[24]++;
    const unsigned need = state & 0x1fu;      // may be 0: deferred with no record
    if (g_stats) for (int k = 0; k < 5; k++) g_stats[25 + k] += (need >> k) & 1u;
    constexpr int KOWN = Prim<IS_VF>::poly(Prim<IS_VF>::NST - 1);
    for (int k = 0; k < 5; k++)
    {
        if (!((need >> k) & 1u)) continue;
        Rec r;
        if (k == KOWN) r = own;      // has must be true
        else if (k == 0) run_export<IS_VF, 0>(a, v, eta, r);
        else if (k == 1) run_export<IS_VF, 1>(a, v, eta, r);
        else if (k == 2) run_export<IS_VF, 2>(a, v, eta, r);
        else if (k == 3) run_export<false, 3>(a, v, eta, r);
        recs.push_back(r);
        nrec++;
    }
    if (((need >> KOWN) & 1u) && !has) return -1;
    return SC_DEFERRED;
}

template <bool IS_VF>
void run(long long n, const int *stencils, const double *q0, const double *q1, int vstride, const double *eta_arr, double eta_all,
         unsigned char *hit, double *toi, unsigned char *stage_out, long long *stats)
{
    constexpr int NVE = Subs<IS_VF>::NVE, NVV = Subs<IS_VF>::NVV, NSUB = 1 + NVE + NVV;
    std::vector<Rec> recs;
    for (long long i = 0; i < n; i++)
    {
        const int *s = stencils + 4 * i;
        const double eta = eta_arr ? eta_arr[i] : eta_all;
        V3 a[4], b[4], v[4];
        for (int k = 0; k < 4; k++)
        {
            a[k] = ldv(q0 + (long long)vstride * s[k]);
            b[k] = ldv(q1 + (long long)vstride * s[k]);
            v[k] = b[k] - a[k];
        }
        // cull (float swept boxes, as np_cull_kernel)
        CullF<IS_VF> c;
        for (int k = 0; k < 4; k++) c.bx[k] = swept_box_f(a[k], b[k]);
        c.init(eta);
        if (!c.stencil_apart()) c.init13();      // as np_cull_kernel
        const unsigned todo = c.todo();
        if (g_stats)
        {
            g_stats[16] += (todo & 1u);
            g_stats[17] += ccd_popc((todo >> 1) & ((1u << NVE) - 1u));
            g_stats[18] += ccd_popc(todo >> (NVE + 1));
            g_stats[19] += (todo == (IS_VF ? 1u : 0u));
        }
        // pass 1
        recs.clear();
        int code[NSUB], first[NSUB], cnt[NSUB];
        for (int sub = 0; sub < NSUB; sub++) { code[sub] = SC_MISS; first[sub] = 0; cnt[sub] = 0; }
        if (todo & 1u)
        {
            first[0] = (int)recs.size();
            code[0] = primitive<IS_VF>(a, v, eta, recs, cnt[0]);
            if (code[0] < 0) { stats[0]++; code[0] = SC_GENERAL; }      // internal inconsistency (counted)
            stats[1 + 0] += (code[0] == SC_DEFERRED);
        }
        for (int sub = 1; sub <= NVE; sub++)
            if ((todo >> sub) & 1u)
            {
                int iv, i1, i2;
                Subs<IS_VF>::ve(sub, iv, i1, i2);
                double r3[3][8];
                int nrec;
                code[sub] = ve_item(a[iv], a[i1], a[i2], v[iv], v[i1], v[i2], eta, r3, nrec);      // np_ve_kernel: classify
                if (code[sub] == SC_DEFERRED)                                                       // np_ve_rec_kernel: rebuild + refine
                {
                    code[sub] = ve_item_refined(a[iv], a[i1], a[i2], v[iv], v[i1], v[i2], eta, r3, nrec);
                    if (g_stats) g_stats[31] += (code[sub] == SC_MISS);
                }
                if (code[sub] == SC_DEFERRED)
                {
                    first[sub] = (int)recs.size();
                    cnt[sub] = nrec;
                    for (int k = 0; k < nrec; k++) { Rec r; memset(r.d, 0, sizeof(r.d)); memcpy(r.d, r3[k], 8 * sizeof(double)); recs.push_back(r); }
                    stats[2]++;
                }
            }
        for (int k = 0; k < NVV; k++)
            if ((todo >> (NVE + 1 + k)) & 1u)
            {
                int i1, i2;
                Subs<IS_VF>::vv(k, i1, i2);
                double t = 0;
                if (vertex_vertex(a[i1], a[i2], v[i1], v[i2], eta, t) == R_HIT) code[NVE + 1 + k] = SC_HIT;
            }
        // decide
        int stage = 0, nd = 0, later_hit = 255, dsub[5];
        bool general = false;
        double t = 0.0;
        for (int sub = 0; sub < NSUB; sub++)
        {
            if (code[sub] == SC_GENERAL) { general = true; break; }
            if (code[sub] == SC_HIT)
            {
                if (nd == 0) stage = sub + 1; else later_hit = sub;
                break;
            }
            if (code[sub] == SC_DEFERRED) dsub[nd++] = sub;
        }
        auto vv_time = [&](int sub) {
            int i1, i2;
            Subs<IS_VF>::vv(sub - NVE - 1, i1, i2);
            double tt = 0.0;
            vertex_vertex(a[i1], a[i2], v[i1], v[i2], eta, tt);
            return tt;
        };
        if (!general && nd == 0)
        {
            if (stage) t = vv_time(stage - 1);
        }
        else if (!general)
        {
                  // no-op for anything but pending VE quartics
            solve_records(recs, stats, 0);
            if (code[0] == SC_DEFERRED && cnt[0] > 0) window_item(recs[first[0]].d, cnt[0], IS_VF ? 3 : 4);
            solve_records(recs, stats, 1);
            stage = 0;
            bool settled = false;
            for (int j = 0; j < nd && !settled; j++)
            {
                const int sub = dsub[j];
                const int r = combine_records(cnt[sub] > 0 ? recs[first[sub]].d : nullptr, cnt[sub], !IS_VF && sub == 0, a, v, t);
                if (r == RS_FALLBACK) { general = true; settled = true; }
                else if (r == RS_HIT) { stage = sub + 1; settled = true; }
            }
            if (!settled && later_hit != 255) { stage = later_hit + 1; t = vv_time(later_hit); }
            if (g_stats) { g_stats[32]++; g_stats[33] += (stage != 0); g_stats[40 + (stage < 10 ? stage : 9)]++; }
        }
        if (general)
        {
            stats[3]++;
            stage = stencil_segment_full<IS_VF>(a, b, eta, t);
        }
        hit[i] = stage != 0;
        toi[i] = stage ? t : 0.0;
        if (stage_out) stage_out[i] = (unsigned char)stage;
    }
}

} // namespace

// stats (11 counters): 0 inconsistencies, 1 deferred primitives, 2 deferred vertex-edge tests, 3 stencils redone by the
// general routine, 4+d pending records of degree d (3..6)
extern "C" int np_emul(int is_vf, long long n, const int *stencils, const double *q0, const double *q1, int vstride, const double *eta_arr,
                       double eta_all, unsigned char *hit, double *toi, unsigned char *stage, long long *stats)
{
    for (int k = 0; k < 11; k++) stats[k] = 0;
    g_stats = nullptr;
    if (is_vf) run<true>(n, stencils, q0, q1, vstride, eta_arr, eta_all, hit, toi, stage, stats);
    else run<false>(n, stencils, q0, q1, vstride, eta_arr, eta_all, hit, toi, stage, stats);
    return 0;
}

// root isolator alone: coefficients (descending, degree d in 3..6, c[0] != 0) -> roots in [0,1]
extern "C" int np_emul_roots(int d, const double *c, double *roots)
{
    double r[6];
    int n = 0;
    if (d == 3) { double cc[4]; for (int k = 0; k < 4; k++) cc[k] = c[k]; n = roots01_lane<3>(cc, r); }
    else if (d == 4) { double cc[5]; for (int k = 0; k < 5; k++) cc[k] = c[k]; n = roots01_lane<4>(cc, r); }
    else if (d == 5) { double cc[6]; for (int k = 0; k < 6; k++) cc[k] = c[k]; n = roots01_lane<5>(cc, r); }
    else if (d == 6) { double cc[7]; for (int k = 0; k < 7; k++) cc[k] = c[k]; n = roots01_lane<6>(cc, r); }
    else return -1;
    for (int k = 0; k < n; k++) roots[k] = r[k];
    return n;
}

// SeparatingPlaneNarrowPhase (ccd_sepplane.cuh) for a list of stencils over a CSR History; returns the number of stencils
// whose interval stack overflowed
extern "C" int np_emul_sepplane(int is_vf, long long n, const int *stencils, const double *eta, const long long *hoff, const double *htime,
                                const double *hpos, double eps, unsigned char *hit)
{
    HistView H;
    H.hoff = hoff; H.htime = htime; H.hpos = hpos;
    int bad = 0;
    for (long long i = 0; i < n; i++)
    {
        const int r = is_vf ? sp_check_stencil<true>(H, stencils + 4 * i, eta[i], eps) : sp_check_stencil<false>(H, stencils + 4 * i, eta[i], eps);
        hit[i] = r > 0;
        bad += r < 0;
    }
    return bad;
}

// same as np_emul with the funnel counters: stats has 64 entries (0..10 as np_emul; 16 primitives after the cull, 17 / 18
// vertex-edge / vertex-vertex items after the cull, 19 stencils with nothing to do, 20..24 survivors of stage 0..4,
// 25..29 records needed per polynomial, 31 vertex-edge quartics settled by ve_refine, 32 deferred stencils, 33 of them hits,
// 40..49 deferred stencils by final stage)
extern "C" int np_emul_funnel(int is_vf, long long n, const int *stencils, const double *q0, const double *q1, int vstride, const double *eta_arr,
                              double eta_all, unsigned char *hit, double *toi, unsigned char *stage, long long *stats)
{
    for (int k = 0; k < 64; k++) stats[k] = 0;
    g_stats = stats;
    if (is_vf) run<true>(n, stencils, q0, q1, vstride, eta_arr, eta_all, hit, toi, stage, stats);
    else run<false>(n, stencils, q0, q1, vstride, eta_arr, eta_all, hit, toi, stage, stats);
    g_stats = nullptr;
    return 0;
}
