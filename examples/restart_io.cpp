// Restart files at the edges of the hot loop through the mrmd::io mirror (include/mrmd/io/): restore a .gro written in
// the reference's format, run a few NVE steps of the examples/02 loop, dump it again, and take a thermodynamic-force
// profile through dumpThermoForce / restoreThermoForce (the round trips of mrmd/io/GRO.test.cpp and ThermoForce.test.cpp).
//
//   ./a.out <in.gro> <out.gro> <thermoForce.txt> <steps>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "action/LennardJones.hpp"
#include "action/ThermodynamicForce.hpp"
#include "action/VelocityVerlet.hpp"
#include "communication/GhostLayer.hpp"
#include "data/Atoms.hpp"
#include "data/Subdomain.hpp"
#include "datatypes.hpp"
#include "io/DumpCSV.hpp"
#include "io/DumpGRO.hpp"
#include "io/DumpThermoForce.hpp"
#include "io/RestoreGRO.hpp"
#include "io/RestoreThermoForce.hpp"

using namespace mrmd;

int main(int argc, char* argv[])
{
    if (argc < 5)
    {
        std::fprintf(stderr, "usage: %s <in.gro> <out.gro> <thermoForce.txt> <steps>\n", argv[0]);
        return 2;
    }
    const idx_t nsteps = std::atoll(argv[4]);
    constexpr real_t dt = 0.002, rc = 2.5, skin = 0.1, neighborCutoff = rc + skin;

    auto subdomain = data::Subdomain({0_r, 0_r, 0_r}, {0_r, 0_r, 0_r}, neighborCutoff);  // examples/02:91
    auto atoms = data::Atoms(0);
    io::restoreGRO(argv[1], subdomain, atoms);

    data::HostAtoms h_atoms(0);
    data::deep_copy(h_atoms, atoms);
    real_t sumPos = 0_r, sumVel = 0_r;
    for (idx_t i = 0; i < atoms.numLocalAtoms; ++i)
        for (int d = 0; d < 3; ++d)
        {
            sumPos += h_atoms.getPos()(i, d);
            sumVel += h_atoms.getVel()(i, d);
        }

    communication::GhostLayer ghostLayer;
    HalfVerletList verletList;
    action::LennardJones lennardJones(rc, 1_r, 1_r, 0.7_r);
    real_t maxAtomDisplacement = std::numeric_limits<real_t>::max();
    for (idx_t step = 0; step < nsteps; ++step)
    {
        maxAtomDisplacement += action::VelocityVerlet::preForceIntegrate(atoms, dt);
        if (maxAtomDisplacement >= skin * 0.5_r)
        {
            maxAtomDisplacement = 0_r;
            ghostLayer.exchangeRealAtoms(atoms, subdomain);
            ghostLayer.createGhostAtoms(atoms, subdomain);
            verletList.build(atoms.getPos(), 0, atoms.numLocalAtoms, neighborCutoff, 1_r, subdomain.minGhostCorner.data(),
                             subdomain.maxGhostCorner.data(), 60);
        }
        else
            ghostLayer.updateGhostAtoms(atoms, subdomain);
        atoms.setForce(0_r);
        lennardJones.apply(atoms, verletList);
        ghostLayer.contributeBackGhostToReal(atoms);
        action::VelocityVerlet::postForceIntegrate(atoms, dt);
    }
    io::dumpGRO(argv[2], atoms, subdomain, real_c(nsteps) * dt, "restart_io", "Argon", {"Ar"}, false, true);
    io::dumpCSV(std::string(argv[2]) + ".csv", atoms, false);

    // thermodynamic force profile: forces(i, j) = (i + 1)(j + 1) as in mrmd/io/ThermoForce.test.cpp:38-44
    const idx_t numBins = 100, numForces = 2;
    data::Subdomain tfDomain({1_r, 2_r, 3_r}, {4_r, 6_r, 8_r}, 0.5_r);
    const real_t binWidth = (tfDomain.maxCorner[0] - tfDomain.minCorner[0]) / real_c(numBins);
    action::ThermodynamicForce tf1({1_r, 1_r}, tfDomain, binWidth, {1_r, 1_r});
    std::vector<real_t> forces(static_cast<size_t>(tf1.numBins() * numForces));
    for (idx_t i = 0; i < tf1.numBins(); ++i)
        for (idx_t j = 0; j < numForces; ++j) forces[static_cast<size_t>(i * numForces + j)] = (real_c(i) + 1_r) * (real_c(j) + 1_r);
    tf1.setForce(forces);
    io::dumpThermoForce(argv[3], tf1);
    auto tf2 = io::restoreThermoForce(argv[3], tfDomain, {1_r, 1_r}, {1_r, 1_r});
    const auto back = tf2.getForce().data.toHost();  // data::MultiHistogram, numBins x numTypes row-major
    const auto grid1 = tf1.createGrid(), grid2 = tf2.createGrid();
    real_t maxForceDiff = 0_r, maxGridDiff = 0_r;
    for (size_t k = 0; k < forces.size(); ++k) maxForceDiff = std::max(maxForceDiff, std::abs(back[k] - forces[k]));
    for (size_t k = 0; k < grid1.size(); ++k) maxGridDiff = std::max(maxGridDiff, std::abs(grid1[k] - grid2[k]));

    std::printf("{\"atoms\": %lld, \"box\": [%.17g, %.17g, %.17g], \"sumPos\": %.17g, \"sumVel\": %.17g, \"E0\": %.17g, "
                "\"tfBins\": %lld, \"tfBinsRestored\": %lld, \"maxForceDiff\": %.3g, \"maxGridDiff\": %.3g}\n",
                static_cast<long long>(atoms.numLocalAtoms), subdomain.diameter[0], subdomain.diameter[1], subdomain.diameter[2],
                sumPos, sumVel, lennardJones.getEnergy(), static_cast<long long>(tf1.numBins()),
                static_cast<long long>(tf2.numBins()), maxForceDiff, maxGridDiff);
    return 0;
}
