// Single-step mesh CCD through the drop-in adapters: the flow of the reference's example/AlecTest.cpp:64-116 with
// ccdgpu::KDOPBroadPhase / ccdgpu::CTCDNarrowPhase in place of the CPU classes.  Built against the reference's own
// History/Mesh/Stencils sources (oracle/Makefile, target _ref/alec_gpu) — it is the integration proof, not product code.
//
//   alec_gpu V0.obj V1.obj [eta]
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "ccd_b200_adapters.hpp"

static bool readObj(const char *path, Eigen::VectorXd &verts, Eigen::Matrix3Xi &faces)
{
    std::ifstream in(path);
    if (!in)
        return false;
    std::vector<double> v;
    std::vector<int> f;
    std::string line;
    while (std::getline(in, line))
    {
        std::istringstream ss(line);
        std::string tag;
        ss >> tag;
        if (tag == "v")
        {
            double x, y, z;
            ss >> x >> y >> z;
            v.push_back(x); v.push_back(y); v.push_back(z);
        }
        else if (tag == "f")
        {
            int a, b, c;
            ss >> a >> b >> c;
            f.push_back(a - 1); f.push_back(b - 1); f.push_back(c - 1);
        }
    }
    verts.resize((long)v.size());
    for (size_t i = 0; i < v.size(); i++)
        verts[(long)i] = v[i];
    faces.resize(3, (long)(f.size() / 3));
    for (size_t i = 0; i < f.size() / 3; i++)
        for (int j = 0; j < 3; j++)
            faces.coeffRef(j, (long)i) = f[3 * i + j];
    return true;
}

int main(int argc, char **argv)
{
    if (argc < 3)
    {
        std::cerr << "Usage: alec_gpu (initial mesh) (final mesh) [eta]" << std::endl;
        return -1;
    }
    const double eta = argc > 3 ? atof(argv[3]) : 1e-8;
    Eigen::VectorXd q0, q1;
    Eigen::Matrix3Xi faces;
    if (!readObj(argv[1], q0, faces) || !readObj(argv[2], q1, faces))
    {
        std::cerr << "cannot read meshes" << std::endl;
        return -1;
    }
    std::cout << "Loaded " << q0.size() / 3 << " vertices and " << faces.cols() << " faces" << std::endl;
    Mesh m;
    m.faces = faces;
    m.vertices = q0;
    History h(q0);
    h.finishHistory(q1);

    std::set<VertexFaceStencil> can_vfs;
    std::set<EdgeEdgeStencil> can_ees;
    const std::set<int> fixedVerts;
    ccdgpu::KDOPBroadPhase().findCollisionCandidates(h, m, eta, can_vfs, can_ees, fixedVerts);
    std::cout << "Broad phase found " << can_vfs.size() << " vertex-face and " << can_ees.size() << " edge-edge candidates" << std::endl;

    std::set<std::pair<VertexFaceStencil, double> > vfs_eta;
    std::set<std::pair<EdgeEdgeStencil, double> > ees_eta;
    for (std::set<VertexFaceStencil>::iterator it = can_vfs.begin(); it != can_vfs.end(); ++it)
        vfs_eta.insert(vfs_eta.end(), std::make_pair(*it, eta));
    for (std::set<EdgeEdgeStencil>::iterator it = can_ees.begin(); it != can_ees.end(); ++it)
        ees_eta.insert(ees_eta.end(), std::make_pair(*it, eta));
    std::set<VertexFaceStencil> vfs;
    std::set<EdgeEdgeStencil> ees;
    ccdgpu::CTCDNarrowPhase np;
    np.findCollisions(h, vfs_eta, ees_eta, vfs, ees);
    std::cout << "CTCDNarrowPhase:" << std::endl;
    std::cout << "  " << vfs.size() << " vf collisions" << std::endl;
    std::cout << "  " << ees.size() << " ee collisions" << std::endl;
    printf("  earliest TOI %.17g\n", np.earliestTOI);

    // the static entry points, one call each (example/testCTCD.cpp:25-40)
    double t = 0;
    bool hit = ccdgpu::CTCD::vertexFaceCTCD(Eigen::Vector3d(0.5, 0.5, 1), Eigen::Vector3d(0, 0, 0), Eigen::Vector3d(1, 0, 0), Eigen::Vector3d(0, 1, 0),
                                            Eigen::Vector3d(0.5, 0.5, 1), Eigen::Vector3d(0, 0, 0), Eigen::Vector3d(1, 0, 3), Eigen::Vector3d(0, 1, 0), 1e-6, t);
    printf("vertexFaceCTCD %d %.17g\n", (int)hit, t);
    printf("meshSelfDistance %.17g\n", ccdgpu::Distance::meshSelfDistance(q0, faces, fixedVerts));
    return 0;
}
