// NVE Lennard-Jones driver written against the reference's public API (include/mrmd/), the loop of
// examples/02_LennardJones_NVE/02_LennardJones_NVE.cpp:135-216 of XzzX/mrmd: preForceIntegrate -> displacement
// check -> (exchangeRealAtoms + createGhostAtoms + verletList.build | updateGhostAtoms) -> zero force ->
// LennardJones::apply -> contributeBackGhostToReal -> postForceIntegrate.
//
// Instead of a .gro restart file the start configuration is a jittered simple-cubic lattice generated from a
// fixed LCG so that tests/test_gpu_cpp_mirror.py can rebuild the identical system in numpy.
//
//   g++ -std=c++20 -O2 -Iinclude/mrmd examples/lennard_jones_nve.cpp -Lmrmd_b200 -lmrmd_b200 -Wl,-rpath,$PWD/mrmd_b200
//   ./a.out <sites per edge> <steps>
#include <cstdio>
#include <cstdlib>

#include "action/LennardJones.hpp"
#include "action/VelocityVerlet.hpp"
#include "analysis/KineticEnergy.hpp"
#include "analysis/MeanSquareDisplacement.hpp"
#include "analysis/Pressure.hpp"
#include "analysis/SystemMomentum.hpp"
#include "communication/GhostLayer.hpp"
#include "data/Atoms.hpp"
#include "data/Subdomain.hpp"
#include "datatypes.hpp"

using namespace mrmd;

struct Config
{
    idx_t nsteps = 100;
    static constexpr real_t dt = 0.002;
    static constexpr real_t sigma = 1_r;
    static constexpr real_t epsilon = 1_r;
    static constexpr real_t r_cut = 2.5_r * sigma;
    static constexpr real_t r_cap = 0.7_r * sigma;
    static constexpr real_t skin = 0.1_r * sigma;
    static constexpr real_t neighborCutoff = r_cut + skin;
    static constexpr real_t cell_ratio = 1_r;
    static constexpr idx_t estimatedMaxNeighbors = 60;
    static constexpr real_t spacing = 1.25_r;
};

/// 48-bit LCG (the drand48 recurrence), uniform in [0, 1)
struct Lcg
{
    uint64_t s = 0x1234ABCD330EULL;
    real_t operator()()
    {
        s = (s * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
        return real_c(s) / real_c(1ULL << 48);
    }
};

int main(int argc, char* argv[])
{
    Kokkos::ScopeGuard scope_guard(argc, argv);  // examples/02:242
    Kokkos::Timer timer;                         // examples/02:117
    Config config;
    const idx_t sites = argc > 1 ? std::atoll(argv[1]) : 16;
    if (argc > 2) config.nsteps = std::atoll(argv[2]);

    const real_t L = real_c(sites) * config.spacing;
    auto subdomain = data::Subdomain({0_r, 0_r, 0_r}, {L, L, L}, config.neighborCutoff);

    const idx_t n = sites * sites * sites;
    data::HostAtoms h_atoms(n);
    {
        Lcg rnd;
        auto pos = h_atoms.getPos();
        auto vel = h_atoms.getVel();
        auto mass = h_atoms.getMass();
        auto relMass = h_atoms.getRelativeMass();
        idx_t idx = 0;
        for (idx_t i = 0; i < sites; ++i)
            for (idx_t j = 0; j < sites; ++j)
                for (idx_t k = 0; k < sites; ++k, ++idx)
                {
                    const idx_t cell[3] = {i, j, k};
                    for (int d = 0; d < 3; ++d) pos(idx, d) = (real_c(cell[d]) + 0.5_r) * config.spacing + (rnd() - 0.5_r) * 0.4_r;
                    for (int d = 0; d < 3; ++d) vel(idx, d) = rnd() - 0.5_r;
                    mass(idx) = 1_r;
                    relMass(idx) = 1_r;
                }
        h_atoms.numLocalAtoms = n;
        h_atoms.numGhostAtoms = 0;
    }
    auto atoms = data::Atoms(n);
    data::deep_copy(atoms, h_atoms);

    communication::GhostLayer ghostLayer;
    HalfVerletList verletList;
    real_t maxAtomDisplacement = std::numeric_limits<real_t>::max();
    idx_t rebuildCounter = 0;
    action::LennardJones lennardJones(config.r_cut, config.sigma, config.epsilon, config.r_cap);
    analysis::MeanSquareDisplacement meanSquareDisplacement;
    meanSquareDisplacement.reset(atoms);

    for (idx_t step = 0; step < config.nsteps; ++step)
    {
        maxAtomDisplacement += action::VelocityVerlet::preForceIntegrate(atoms, config.dt);
        if (maxAtomDisplacement >= config.skin * 0.5_r)
        {
            maxAtomDisplacement = 0_r;
            ghostLayer.exchangeRealAtoms(atoms, subdomain);
            ghostLayer.createGhostAtoms(atoms, subdomain);
            verletList.build(atoms.getPos(), 0, atoms.numLocalAtoms, config.neighborCutoff, config.cell_ratio,
                             subdomain.minGhostCorner.data(), subdomain.maxGhostCorner.data(), config.estimatedMaxNeighbors);
            ++rebuildCounter;
        }
        else
        {
            ghostLayer.updateGhostAtoms(atoms, subdomain);
        }
        auto force = atoms.getForce();
        Cabana::deep_copy(force, 0_r);
        lennardJones.apply(atoms, verletList);
        ghostLayer.contributeBackGhostToReal(atoms);
        action::VelocityVerlet::postForceIntegrate(atoms, config.dt);
    }

    // the statistics of the reference's table (examples/02:190-199)
    const auto E0 = lennardJones.getEnergy();
    const auto Ek = analysis::getKineticEnergy(atoms);
    const auto T = (2_r / 3_r) * analysis::getMeanKineticEnergy(atoms);
    const auto p = analysis::getPressure(atoms, subdomain);
    const auto systemMomentum = analysis::getSystemMomentum(atoms);
    const auto msd = meanSquareDisplacement.calc(atoms, subdomain);

    // neighbours of the first atom through the reference's accessors (HalfNeighborList, datatypes.hpp:193)
    idx_t neighborSum = 0;
    for (idx_t n = 0; n < HalfNeighborList::numNeighbor(verletList, 0); ++n) neighborSum += HalfNeighborList::getNeighbor(verletList, 0, n);

    data::deep_copy(h_atoms, atoms);
    auto pos = h_atoms.getPos();
    std::printf("{\"seconds\": %.6f, \"neighbors0\": %lld, \"neighborSum0\": %lld, ", timer.seconds(),
                static_cast<long long>(HalfNeighborList::numNeighbor(verletList, 0)), static_cast<long long>(neighborSum));
    std::printf("\"atoms\": %lld, \"ghosts\": %lld, \"steps\": %lld, \"rebuilds\": %lld, \"pairs\": %zu, "
                "\"E0\": %.17g, \"Ek\": %.17g, \"T\": %.17g, \"p\": %.17g, \"msd\": %.17g, \"momentum\": [%.17g, %.17g, %.17g], "
                "\"x0\": [%.17g, %.17g, %.17g]}\n",
                static_cast<long long>(atoms.numLocalAtoms), static_cast<long long>(atoms.numGhostAtoms),
                static_cast<long long>(config.nsteps), static_cast<long long>(rebuildCounter), verletList.totalPairs(), E0,
                Ek, T, p, msd, systemMomentum[0], systemMomentum[1], systemMomentum[2], pos(0, 0), pos(0, 1), pos(0, 2));
    return 0;
}
