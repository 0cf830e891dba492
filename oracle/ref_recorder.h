// TEST INFRASTRUCTURE — recording wrappers for the detection calls ActiveLayers makes.
//
// Force-included (-include) when the UNMODIFIED src/ActiveLayers.cpp is compiled for oracle/_ref/libccdvf.so: the real
// header is included first, then the narrowphase class ActiveLayers instantiates (src/ActiveLayers.cpp:28) is mapped to a
// subclass that calls the reference's own implementation and keeps a copy of what went in (the History and every
// broadphase candidate with its thickness) and what came out.
// Nothing is re-implemented: every recorded number was computed by the reference's object code.
#ifndef CCD_REF_RECORDER_H
#define CCD_REF_RECORDER_H
#include <set>
#include <vector>
#include <utility>
#include "RetrospectiveDetection.h"
#include "SeparatingPlaneNarrowPhase.h"
#include "History.h"
#include "Mesh.h"

struct RecordedPass
{
    // History as CSR (every vertex of the mesh), the broadphase call, then the narrowphase call of the same pass
    std::vector<long long> hoff;
    std::vector<double> htime, hpos;
    std::vector<int> np_vf, np_ee;         // what the narrowphase was given: every broadphase candidate (src/ActiveLayers.cpp:193-212), set order
    std::vector<double> np_vf_eta, np_ee_eta;
    std::vector<unsigned char> np_vf_hit, np_ee_hit;
};

struct RecorderStop { };      // thrown once the requested number of passes has been recorded
std::vector<RecordedPass> &recorder_passes();
int &recorder_limit();
int &recorder_nverts();

class RecordingNarrowPhase : public SeparatingPlaneNarrowPhase
{
public:
    virtual void findCollisions(const History &h, const std::set<std::pair<VertexFaceStencil, double> > &candidateVFS,
                                const std::set<std::pair<EdgeEdgeStencil, double> > &candidateEES, std::set<VertexFaceStencil> &vfs,
                                std::set<EdgeEdgeStencil> &ees);
};

#ifdef CCD_RECORDER_SUBSTITUTE
#define SeparatingPlaneNarrowPhase RecordingNarrowPhase
#endif
#endif
