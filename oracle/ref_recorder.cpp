// TEST INFRASTRUCTURE — see ref_recorder.h.  C entry points: run the UNMODIFIED VelocityFilter::velocityFilter
// (src/VelocityFilter.cpp) until `npasses` detection passes of its ActiveLayers have been recorded, then read them back.
#include "ref_recorder.h"
#include "VelocityFilter.h"
#include "PenaltyGroup.h"
#include "PenaltyPotential.h"
#include <cstring>
#include <iostream>
#include <sstream>

using namespace Eigen;

std::vector<RecordedPass> &recorder_passes() { static std::vector<RecordedPass> p; return p; }
int &recorder_limit() { static int n = 0; return n; }
int &recorder_nverts() { static int n = 0; return n; }

void RecordingNarrowPhase::findCollisions(const History &h, const std::set<std::pair<VertexFaceStencil, double> > &candidateVFS,
                                          const std::set<std::pair<EdgeEdgeStencil, double> > &candidateEES, std::set<VertexFaceStencil> &vfs,
                                          std::set<EdgeEdgeStencil> &ees)
{
    std::set<VertexFaceStencil> hv;
    std::set<EdgeEdgeStencil> he;
    SeparatingPlaneNarrowPhase::findCollisions(h, candidateVFS, candidateEES, hv, he);
    {
        recorder_passes().push_back(RecordedPass());
        RecordedPass &P = recorder_passes().back();
        const int V = recorder_nverts();
        P.hoff.push_back(0);
        for (int v = 0; v < V; v++)
        {
            const std::vector<HistoryEntry> &hvv = h.getVertexHistory(v);
            for (size_t k = 0; k < hvv.size(); k++)
            {
                P.htime.push_back(hvv[k].time);
                for (int cc = 0; cc < 3; cc++) P.hpos.push_back(hvv[k].pos[cc]);
            }
            P.hoff.push_back((long long)P.htime.size());
        }
        for (std::set<std::pair<VertexFaceStencil, double> >::const_iterator it = candidateVFS.begin(); it != candidateVFS.end(); ++it)
        {
            P.np_vf.push_back(it->first.p); P.np_vf.push_back(it->first.q0); P.np_vf.push_back(it->first.q1); P.np_vf.push_back(it->first.q2);
            P.np_vf_eta.push_back(it->second);
            P.np_vf_hit.push_back(hv.count(it->first) ? 1 : 0);
        }
        for (std::set<std::pair<EdgeEdgeStencil, double> >::const_iterator it = candidateEES.begin(); it != candidateEES.end(); ++it)
        {
            P.np_ee.push_back(it->first.p0); P.np_ee.push_back(it->first.p1); P.np_ee.push_back(it->first.q0); P.np_ee.push_back(it->first.q1);
            P.np_ee_eta.push_back(it->second);
            P.np_ee_hit.push_back(he.count(it->first) ? 1 : 0);
        }
    }
    vfs.insert(hv.begin(), hv.end());
    ees.insert(he.begin(), he.end());
    if ((int)recorder_passes().size() >= recorder_limit())
        throw RecorderStop();
}

extern "C" {

// returns the number of recorded passes (each = one broadphase + one narrowphase call of ActiveLayers)
int vfrec_run(int V, int F, const double *q1, const double *q2, const int *faces, const double *invmass_per_vertex, double outerEta,
              double innerEta, int npasses)
{
    VectorXd a(3 * (long)V), b(3 * (long)V), im(3 * (long)V);
    for (long i = 0; i < 3 * (long)V; i++) { a[i] = q1[i]; b[i] = q2[i]; im[i] = invmass_per_vertex[i / 3]; }
    Matrix3Xi f;
    f.resize(3, F);
    for (long i = 0; i < F; i++)
        for (int j = 0; j < 3; j++) f.coeffRef(j, i) = faces[3 * i + j];
    recorder_passes().clear();
    recorder_limit() = npasses;
    recorder_nverts() = V;
    std::streambuf *old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());      // the reference prints progress lines
    try { VelocityFilter::velocityFilter(a, b, f, im, outerEta, innerEta); }
    catch (const RecorderStop &) { }
    std::cout.rdbuf(old);
    return (int)recorder_passes().size();
}

// sizes of pass p: N history entries, nvf, nee narrowphase inputs, V
void vfrec_sizes(int p, long long *out)
{
    const RecordedPass &P = recorder_passes()[p];
    out[0] = (long long)P.htime.size(); out[1] = (long long)P.np_vf.size() / 4; out[2] = (long long)P.np_ee.size() / 4;
    out[3] = (long long)P.hoff.size() - 1;
}

void vfrec_get(int p, long long *hoff, double *htime, double *hpos, int *np_vf, int *np_ee, double *np_vf_eta, double *np_ee_eta,
               unsigned char *np_vf_hit, unsigned char *np_ee_hit)
{
    const RecordedPass &P = recorder_passes()[p];
#define CP(dst, v) if (!(v).empty()) memcpy(dst, (v).data(), sizeof((v)[0]) * (v).size())
    CP(hoff, P.hoff); CP(htime, P.htime); CP(hpos, P.hpos);
    CP(np_vf, P.np_vf); CP(np_ee, P.np_ee); CP(np_vf_eta, P.np_vf_eta); CP(np_ee_eta, P.np_ee_eta);
    CP(np_vf_hit, P.np_vf_hit); CP(np_ee_hit, P.np_ee_hit);
#undef CP
}

// The reference's own PenaltyGroup (src/PenaltyGroup.cpp) over given stencil lists: addVFStencil / addEEStencil in list order,
// optional rollback() (clears every stencil's isnew), then addForce(q, v, F).  fired: VertexFacePenaltyPotential::addForce /
// EdgeEdgePenaltyPotential::addForce called once more per stencil on a scratch force vector (their return value).
int ref_penalty_group_force(int V, const double *q, const double *v, long long nvf, const int *vf, long long nee, const int *ee, int rolled_back,
                            double dt, double outerEta, double innerEta, double stiffness, double CoR, double *F, unsigned char *fired)
{
    VectorXd Q(3 * (long)V), Vv(3 * (long)V), Ff(3 * (long)V), scratch(3 * (long)V);
    for (long i = 0; i < 3 * (long)V; i++) { Q[i] = q[i]; Vv[i] = v[i]; Ff[i] = F[i]; }
    scratch.setZero();
    PenaltyGroup g(dt, outerEta, innerEta, stiffness, CoR);
    for (long long i = 0; i < nvf; i++) g.addVFStencil(VertexFaceStencil(vf[4 * i], vf[4 * i + 1], vf[4 * i + 2], vf[4 * i + 3]));
    for (long long i = 0; i < nee; i++) g.addEEStencil(EdgeEdgeStencil(ee[4 * i], ee[4 * i + 1], ee[4 * i + 2], ee[4 * i + 3]));
    if (rolled_back) g.rollback();
    const bool newused = g.addForce(Q, Vv, Ff);
    for (long i = 0; i < 3 * (long)V; i++) F[i] = Ff[i];
    if (fired)
    {
        for (long long i = 0; i < nvf; i++)
            fired[i] = VertexFacePenaltyPotential::addForce(Q, Vv, scratch, VertexFaceStencil(vf[4 * i], vf[4 * i + 1], vf[4 * i + 2], vf[4 * i + 3]), outerEta, innerEta, stiffness, CoR);
        for (long long i = 0; i < nee; i++)
            fired[nvf + i] = EdgeEdgePenaltyPotential::addForce(Q, Vv, scratch, EdgeEdgeStencil(ee[4 * i], ee[4 * i + 1], ee[4 * i + 2], ee[4 * i + 3]), outerEta, innerEta, stiffness, CoR);
    }
    return newused ? 1 : 0;
}

} // extern "C"
