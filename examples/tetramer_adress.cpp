// AdResS Lennard-Jones tetramers with a spherical atomistic region (BASELINE.json configs[3]) written against the
// mirror of the reference API (include/mrmd/): the AdResS step of SURVEY.md section 3.5 (UpdateMolecules, LJ_IdealGas,
// ContributeMoleculeForceToAtoms, MultiResGhostLayer) with MoleculeConstraints around the Langevin integrator in the
// order of the reference's tests/Constraints/Constraints.cpp:53-64.  Four atoms per molecule, atoms of a molecule
// contiguous.
//
//   g++ -std=c++20 -O2 -Iinclude/mrmd examples/tetramer_adress.cpp -Lmrmd_b200 -lmrmd_b200 -Wl,-rpath,$PWD/mrmd_b200
//   ./a.out <molecules per edge> <steps>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>

#include "action/ContributeMoleculeForceToAtoms.hpp"
#include "action/LJ_IdealGas.hpp"
#include "action/Shake.hpp"
#include "action/UpdateMolecules.hpp"
#include "action/VelocityVerletLangevinThermostat.hpp"
#include "communication/MultiResGhostLayer.hpp"
#include "data/Atoms.hpp"
#include "data/Bond.hpp"
#include "data/Molecules.hpp"
#include "data/Subdomain.hpp"
#include "datatypes.hpp"
#include "weighting_function/Spherical.hpp"

using namespace mrmd;

struct Config
{
    idx_t nsteps = 60;
    static constexpr real_t dt = 0.002_r;
    static constexpr real_t sigma = 1_r;
    static constexpr real_t epsilon = 1_r;
    static constexpr real_t r_cut = 2.5_r;
    static constexpr real_t r_cap = 0.7_r;
    static constexpr real_t skin = 0.1_r;
    static constexpr real_t neighborCutoff = r_cut + skin;
    static constexpr idx_t estimatedMaxNeighbors = 40;
    static constexpr real_t spacing = 1.98425_r;  ///< molecule density 0.128, atom density 0.512
    static constexpr real_t bondLength = 1_r;
    static constexpr idx_t atomsPerMolecule = 4;
    static constexpr idx_t constraintIterations = 3;
    static constexpr real_t temperature = 1.5_r;
    static constexpr real_t gamma = 20_r;
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
    const idx_t sites = argc > 1 ? std::atoll(argv[1]) : 10;
    if (argc > 2) config.nsteps = std::atoll(argv[2]);

    const real_t L = real_c(sites) * config.spacing;
    auto subdomain = data::Subdomain({0_r, 0_r, 0_r}, {L, L, L}, config.neighborCutoff);
    const idx_t numMolecules = sites * sites * sites;
    const idx_t numAtoms = config.atomsPerMolecule * numMolecules;
    const real_t grown = (L + 2_r * config.neighborCutoff) / L;
    const idx_t molCapacity = idx_c(real_c(numMolecules) * grown * grown * grown * 1.3_r) + 1024;

    data::HostAtoms h_atoms(config.atomsPerMolecule * molCapacity);
    data::HostMolecules h_molecules(molCapacity);
    {
        // regular tetrahedron of edge 1 around every lattice site, one velocity per molecule
        const real_t a = 1_r / (2_r * std::sqrt(2_r));
        const real_t tet[4][3] = {{a, a, a}, {a, -a, -a}, {-a, a, -a}, {-a, -a, a}};
        Lcg rnd;
        auto pos = h_atoms.getPos();
        auto vel = h_atoms.getVel();
        idx_t m = 0;
        for (idx_t i = 0; i < sites; ++i)
            for (idx_t j = 0; j < sites; ++j)
                for (idx_t k = 0; k < sites; ++k, ++m)
                {
                    const idx_t cell[3] = {i, j, k};
                    real_t v[3];
                    for (int d = 0; d < 3; ++d) v[d] = rnd() - 0.5_r;
                    for (idx_t t = 0; t < 4; ++t)
                    {
                        const idx_t idx = 4 * m + t;
                        for (int d = 0; d < 3; ++d)
                        {
                            pos(idx, d) = (real_c(cell[d]) + 0.5_r) * config.spacing + tet[t][d];
                            vel(idx, d) = v[d];
                        }
                        h_atoms.getMass()(idx) = 1_r;
                        h_atoms.getRelativeMass()(idx) = 0.25_r;
                    }
                    h_molecules.getAtomsOffset()(m) = 4 * m;
                    h_molecules.getNumAtoms()(m) = 4;
                }
        h_atoms.numLocalAtoms = numAtoms;
        h_molecules.numLocalMolecules = numMolecules;
    }
    data::Atoms atoms(config.atomsPerMolecule * molCapacity);
    data::deep_copy(atoms, h_atoms);
    data::Molecules molecules(molCapacity);
    data::deep_copy(molecules, h_molecules);

    // config 4: R = 60, h = 30 in the 317.48 box, scaled with the box
    auto weightingFunction = weighting_function::Spherical(subdomain.getCenter(), 60_r / 317.48_r * L, 30_r / 317.48_r * L, 2);
    communication::MultiResGhostLayer ghostLayer;
    HalfVerletList moleculesVerletList;
    action::LJ_IdealGas LJ(config.r_cap, config.r_cut, config.sigma, config.epsilon, true);
    LJ.setAtomsPerMolecule(config.atomsPerMolecule);
    action::VelocityVerletLangevinThermostat integrator(config.gamma, config.temperature);
    data::BondView bonds(6);
    {
        idx_t b = 0;
        for (idx_t i = 0; i < 4; ++i)
            for (idx_t j = i + 1; j < 4; ++j, ++b)
            {
                bonds[b].idx = i;
                bonds[b].jdx = j;
                bonds[b].eqDistance = config.bondLength;
            }
    }
    action::MoleculeConstraints constraints(config.atomsPerMolecule, config.constraintIterations);
    constraints.setConstraints(bonds);

    real_t maxAtomDisplacement = std::numeric_limits<real_t>::max();
    idx_t rebuildCounter = 0;
    real_t energy = 0_r;
    for (idx_t step = 0; step < config.nsteps; ++step)
    {
        constraints.enforcePositionalConstraints(molecules, atoms, config.dt);
        maxAtomDisplacement += integrator.preForceIntegrate(atoms, config.dt);
        if (maxAtomDisplacement >= config.skin * 0.5_r)
        {
            maxAtomDisplacement = 0_r;
            action::UpdateMolecules::update(molecules, atoms, weightingFunction);
            ghostLayer.exchangeRealAtoms(molecules, atoms, subdomain);
            ghostLayer.createGhostAtoms(molecules, atoms, subdomain);
            action::UpdateMolecules::update(molecules, atoms, weightingFunction);
            moleculesVerletList.build(molecules.getPos(), 0, molecules.numLocalMolecules, config.neighborCutoff, 1_r,
                                      subdomain.minGhostCorner.data(), subdomain.maxGhostCorner.data(), config.estimatedMaxNeighbors);
            ++rebuildCounter;
        }
        else
        {
            ghostLayer.updateGhostAtoms(atoms, subdomain);
            action::UpdateMolecules::update(molecules, atoms, weightingFunction);
        }
        atoms.setForce(0_r);
        molecules.setForce(0_r);
        energy = LJ.run(molecules, moleculesVerletList, atoms);
        action::ContributeMoleculeForceToAtoms::update(molecules, atoms);
        ghostLayer.contributeBackGhostToReal(atoms);
        integrator.postForceIntegrate(atoms, config.dt);
        constraints.enforceVelocityConstraints(molecules, atoms, config.dt);
    }

    data::deep_copy(h_atoms, atoms);
    auto pos = h_atoms.getPos();
    auto vel = h_atoms.getVel();
    real_t maxBondError = 0_r;
    for (idx_t m = 0; m < numMolecules; ++m)
        for (idx_t i = 0; i < 4; ++i)
            for (idx_t j = i + 1; j < 4; ++j)
            {
                real_t d2 = 0_r;
                for (int d = 0; d < 3; ++d) d2 += (pos(4 * m + i, d) - pos(4 * m + j, d)) * (pos(4 * m + i, d) - pos(4 * m + j, d));
                maxBondError = std::max(maxBondError, std::abs(std::sqrt(d2) - config.bondLength));
            }
    std::printf(
        "{\"atoms\": %lld, \"steps\": %lld, \"rebuilds\": %lld, \"ghostAtoms\": %lld, \"E\": %.17g, \"maxBondError\": %.17g, "
        "\"x0\": [%.17g, %.17g, %.17g], \"v0\": [%.17g, %.17g, %.17g]}\n",
        static_cast<long long>(numAtoms), static_cast<long long>(config.nsteps), static_cast<long long>(rebuildCounter),
        static_cast<long long>(atoms.numGhostAtoms), energy, maxBondError, pos(0, 0), pos(0, 1), pos(0, 2), vel(0, 0), vel(0, 1),
        vel(0, 2));
    return 0;
}
