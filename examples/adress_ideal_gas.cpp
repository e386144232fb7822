// AdResS LJ <-> ideal-gas driver written against the reference's public API (include/mrmd/): the step SURVEY.md
// section 3.5 assembles from action::UpdateMolecules, action::LJ_IdealGas, action::ThermodynamicForce,
// action::ContributeMoleculeForceToAtoms and communication::MultiResGhostLayer, with the local capping of
// examples/04_LennardJones_IdealGas_LocalCap (apply_if with the either()/both() predicates) available as an
// option.  One molecule per atom (data::createMoleculeForEachAtom).
//
//   g++ -std=c++20 -O2 -Iinclude/mrmd examples/adress_ideal_gas.cpp -Lmrmd_b200 -lmrmd_b200 -Wl,-rpath,$PWD/mrmd_b200
//   ./a.out <sites per edge> <steps>
#include <cstdio>
#include <cstdlib>

#include "action/ContributeMoleculeForceToAtoms.hpp"
#include "action/LJ_IdealGas.hpp"
#include "action/LennardJones.hpp"
#include "action/ThermodynamicForce.hpp"
#include "action/UpdateMolecules.hpp"
#include "action/VelocityVerletLangevinThermostat.hpp"
#include "communication/MultiResGhostLayer.hpp"
#include "data/Atoms.hpp"
#include "data/Molecules.hpp"
#include "data/MoleculesFromAtoms.hpp"
#include "data/Subdomain.hpp"
#include "datatypes.hpp"
#include "util/IsInSymmetricSlab.hpp"
#include "weighting_function/Slab.hpp"

using namespace mrmd;

struct Config
{
    idx_t nsteps = 200;
    static constexpr real_t dt = 0.002;
    static constexpr real_t sigma = 1_r;
    static constexpr real_t epsilon = 1_r;
    static constexpr real_t r_cut = 2.5_r;
    static constexpr real_t r_cap = 0.7_r;
    static constexpr real_t skin = 0.3_r;
    static constexpr real_t neighborCutoff = r_cut + skin;
    static constexpr real_t cell_ratio = 0.5_r;
    static constexpr idx_t estimatedMaxNeighbors = 60;
    static constexpr real_t spacing = 1.35_r;
    static constexpr real_t temperature = 1.5_r;
    static constexpr real_t gamma = 10_r;
    // thermodynamic force
    static constexpr real_t densityBinWidth = 0.5_r;
    static constexpr real_t thermodynamicForceModulation = 2_r;
    static constexpr idx_t densitySamplingInterval = 10;
    static constexpr idx_t densityUpdateInterval = 100;
    static constexpr real_t smoothingSigma = 2_r;
    static constexpr real_t smoothingIntensity = 2_r;
};

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
    Config config;
    const idx_t sites = argc > 1 ? std::atoll(argv[1]) : 16;
    if (argc > 2) config.nsteps = std::atoll(argv[2]);

    const real_t L = real_c(sites) * config.spacing;
    auto subdomain = data::Subdomain({0_r, 0_r, 0_r}, {L, L, L}, config.neighborCutoff);
    const auto center = subdomain.getCenter();

    const idx_t n = sites * sites * sites;
    data::HostAtoms h_atoms(n);
    {
        Lcg rnd;
        auto pos = h_atoms.getPos();
        auto vel = h_atoms.getVel();
        idx_t idx = 0;
        for (idx_t i = 0; i < sites; ++i)
            for (idx_t j = 0; j < sites; ++j)
                for (idx_t k = 0; k < sites; ++k, ++idx)
                {
                    const idx_t cell[3] = {i, j, k};
                    for (int d = 0; d < 3; ++d) pos(idx, d) = (real_c(cell[d]) + 0.5_r) * config.spacing + (rnd() - 0.5_r) * 0.4_r;
                    for (int d = 0; d < 3; ++d) vel(idx, d) = rnd() - 0.5_r;
                    h_atoms.getMass()(idx) = 1_r;
                    h_atoms.getRelativeMass()(idx) = 1_r;
                }
        h_atoms.numLocalAtoms = n;
    }
    auto atoms = data::Atoms(n);
    data::deep_copy(atoms, h_atoms);
    auto molecules = data::createMoleculeForEachAtom(atoms);
    const real_t rho = real_c(n) / subdomain.getVolume();

    // atomistic slab of width L/4 around the centre, hybrid regions of width L/8 on both sides
    auto weightingFunction = weighting_function::Slab(center, 0.25_r * L, 0.125_r * L, 7);
    util::IsInSymmetricSlab applicationRegion(center, 0.125_r * L - 1_r, 0.25_r * L + 1_r);

    communication::MultiResGhostLayer ghostLayer;
    HalfVerletList moleculesVerletList;
    action::LJ_IdealGas LJ(config.r_cap, config.r_cut, config.sigma, config.epsilon, true);
    action::ThermodynamicForce thermodynamicForce(rho, subdomain, config.densityBinWidth, config.thermodynamicForceModulation);
    action::VelocityVerletLangevinThermostat integrator(config.gamma, config.temperature);

    real_t maxAtomDisplacement = std::numeric_limits<real_t>::max();
    idx_t rebuildCounter = 0;
    real_t energy = 0_r;
    for (idx_t step = 0; step < config.nsteps; ++step)
    {
        maxAtomDisplacement += integrator.preForceIntegrate(atoms, config.dt);
        if (maxAtomDisplacement >= config.skin * 0.5_r)
        {
            maxAtomDisplacement = 0_r;
            ghostLayer.exchangeRealAtoms(molecules, atoms, subdomain);
            ghostLayer.createGhostAtoms(molecules, atoms, subdomain);
            moleculesVerletList.build(molecules.getPos(), 0, molecules.numLocalMolecules, config.neighborCutoff, config.cell_ratio,
                                      subdomain.minGhostCorner.data(), subdomain.maxGhostCorner.data(), config.estimatedMaxNeighbors);
            ++rebuildCounter;
        }
        else
        {
            ghostLayer.updateGhostAtoms(atoms, subdomain);
        }
        action::UpdateMolecules::update(molecules, atoms, weightingFunction);

        atoms.setForce(0_r);
        molecules.setForce(0_r);

        if (step % config.densitySamplingInterval == 0) thermodynamicForce.sample(atoms);
        if (step % config.densityUpdateInterval == 0 && step > 0) thermodynamicForce.update(config.smoothingSigma, config.smoothingIntensity);
        thermodynamicForce.apply_if(atoms, applicationRegion);

        energy = LJ.run(molecules, moleculesVerletList, atoms);
        action::ContributeMoleculeForceToAtoms::update(molecules, atoms);
        ghostLayer.contributeBackGhostToReal(atoms);
        integrator.postForceIntegrate(atoms, config.dt);
    }

    data::HostMolecules h_molecules(0);
    data::deep_copy(h_molecules, molecules);
    idx_t numAT = 0, numHY = 0;
    for (idx_t i = 0; i < molecules.numLocalMolecules; ++i)
    {
        const auto lambda = h_molecules.getLambda()(i);
        numAT += weighting_function::isInATRegion(lambda);
        numHY += weighting_function::isInHYRegion(lambda);
    }
    const auto muLeft = thermodynamicForce.getMuLeft();
    std::printf("{\"atoms\": %lld, \"ghosts\": %lld, \"steps\": %lld, \"rebuilds\": %lld, \"pairs\": %zu, \"E\": %.17g, "
                "\"numAT\": %lld, \"numHY\": %lld, \"densitySamples\": %lld, \"muLeft\": %.17g}\n",
                static_cast<long long>(atoms.numLocalAtoms), static_cast<long long>(atoms.numGhostAtoms),
                static_cast<long long>(config.nsteps), static_cast<long long>(rebuildCounter), moleculesVerletList.totalPairs(), energy,
                static_cast<long long>(numAT), static_cast<long long>(numHY),
                static_cast<long long>(thermodynamicForce.getNumberOfDensityProfileSamples()), muLeft[0]);
    return 0;
}
