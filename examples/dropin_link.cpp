// Link-level drop-in for evouga/collisiondetection: definitions of the reference's OWN classes' entry points
// (KDOPBroadPhase / AABBBroadPhase::findCollisionCandidates, CTCDNarrowPhase / SeparatingPlaneNarrowPhase::findCollisions,
// Distance::meshSelfDistance — src/RetrospectiveDetection.h:10-23, include/Distance.h:180) that forward to the GPU library
// through the adapters of include/ccd_b200_adapters.hpp.  Link these objects INSTEAD of the reference's
// KDOPBroadPhase.o, AABBBroadPhase.o, CTCDNarrowPhase.o, SeparatingPlaneNarrowPhase.o and Distance.o and every caller —
// ActiveLayers (src/ActiveLayers.cpp:26-28, :193, :212), VelocityFilter (src/VelocityFilter.cpp:58), the example drivers
// AlecTest / testVelocityFilter / testNewSequence — runs its detection passes on the B200 without a changed source line.
// oracle/Makefile builds the reference's unmodified drivers both ways (_ref/<driver>_cpu, _ref/<driver>_gpu); the GPU test
// tests/test_gpu_parity.py::test_reference_drivers_with_gpu_detection compares what they print.
//
// The reference's two broadphase headers both define a class NodeComparator, so the AABB entry point is compiled as a
// translation unit of its own (-DCCD_DROPIN_AABB).
#include "ccd_b200_adapters.hpp"

#ifdef CCD_DROPIN_AABB
#include "AABBBroadPhase.h"

void AABBBroadPhase::findCollisionCandidates(const History &h, const Mesh &m, double outerEta, std::set<VertexFaceStencil> &vfs,
                                             std::set<EdgeEdgeStencil> &ees, const std::set<int> &fixedVerts)
{
    ccdgpu::AABBBroadPhase().findCollisionCandidates(h, m, outerEta, vfs, ees, fixedVerts);
}

#else
#include "KDOPBroadPhase.h"
#include "CTCDNarrowPhase.h"
#include "SeparatingPlaneNarrowPhase.h"
#include "Distance.h"

KDOPBroadPhase::KDOPBroadPhase() {}      // the reference fills DOPaxis here (src/KDOPBroadPhase.cpp:9-32); the axes live in the library

void KDOPBroadPhase::findCollisionCandidates(const History &h, const Mesh &m, double outerEta, std::set<VertexFaceStencil> &vfs,
                                             std::set<EdgeEdgeStencil> &ees, const std::set<int> &fixedVerts)
{
    ccdgpu::KDOPBroadPhase().findCollisionCandidates(h, m, outerEta, vfs, ees, fixedVerts);
}

void CTCDNarrowPhase::findCollisions(const History &h, const std::set<std::pair<VertexFaceStencil, double> > &candidateVFS,
                                     const std::set<std::pair<EdgeEdgeStencil, double> > &candidateEES, std::set<VertexFaceStencil> &vfs,
                                     std::set<EdgeEdgeStencil> &ees)
{
    ccdgpu::CTCDNarrowPhase().findCollisions(h, candidateVFS, candidateEES, vfs, ees);
}

void SeparatingPlaneNarrowPhase::findCollisions(const History &h, const std::set<std::pair<VertexFaceStencil, double> > &candidateVFS,
                                                const std::set<std::pair<EdgeEdgeStencil, double> > &candidateEES, std::set<VertexFaceStencil> &vfs,
                                                std::set<EdgeEdgeStencil> &ees)
{
    ccdgpu::SeparatingPlaneNarrowPhase().findCollisions(h, candidateVFS, candidateEES, vfs, ees);
}

// src/Distance.cpp:12-66, including the line it prints (:47)
double Distance::meshSelfDistance(const Eigen::VectorXd &verts, const Eigen::Matrix3Xi &faces, const std::set<int> &fixedVerts)
{
    return ccdgpu::Distance::meshSelfDistance(verts, faces, fixedVerts);
}
#endif
