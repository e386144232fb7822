// Constrained SPC water in a periodic box through the mirror of the reference API (include/mrmd/): action::SPC
// (mrmd/action/SPC.hpp) driven in the call order of the reference's tests/Constraints/Constraints.cpp:25-72 (SHAKE
// before the integrator, RATTLE after it) around the molecule loop of SURVEY.md section 3.5 (MultiResGhostLayer,
// UpdateMolecules, half Verlet list of molecules).
//
//   g++ -std=c++20 -O2 -Iinclude/mrmd examples/spc_water.cpp -Lmrmd_b200 -lmrmd_b200 -Wl,-rpath,$PWD/mrmd_b200
//   ./a.out <molecules per edge> <steps>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>

#include "action/SPC.hpp"
#include "action/UpdateMolecules.hpp"
#include "action/VelocityVerlet.hpp"
#include "communication/MultiResGhostLayer.hpp"
#include "data/Atoms.hpp"
#include "data/Molecules.hpp"
#include "data/Subdomain.hpp"
#include "datatypes.hpp"
#include "weighting_function/Slab.hpp"

using namespace mrmd;

struct Config
{
    idx_t nsteps = 30;
    static constexpr real_t dt = 0.0005_r;  ///< unit: ps
    static constexpr real_t skin = 0.02_r;  ///< unit: nm
    static constexpr real_t neighborCutoff = action::SPC::rc + skin;
    static constexpr real_t cell_ratio = 1_r;
    static constexpr idx_t estimatedMaxNeighbors = 220;
    static constexpr real_t spacing = 0.31_r;  ///< unit: nm, about 33.5 molecules / nm^3
    static constexpr real_t jitter = 0.02_r;
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
    const idx_t numAtoms = 3 * numMolecules;
    // room for the ghost layer
    const real_t grown = (L + 2_r * config.neighborCutoff) / L;
    const idx_t molCapacity = idx_c(real_c(numMolecules) * grown * grown * grown * 1.3_r) + 1024;

    data::HostAtoms h_atoms(3 * molCapacity);
    data::HostMolecules h_molecules(molCapacity);
    {
        Lcg rnd;
        auto pos = h_atoms.getPos();
        auto vel = h_atoms.getVel();
        const real_t massMolecule = action::SPC::massO + 2_r * action::SPC::massH;
        idx_t m = 0;
        for (idx_t i = 0; i < sites; ++i)
            for (idx_t j = 0; j < sites; ++j)
                for (idx_t k = 0; k < sites; ++k, ++m)
                {
                    const idx_t cell[3] = {i, j, k};
                    const idx_t o = 3 * m;
                    for (int d = 0; d < 3; ++d) pos(o, d) = (real_c(cell[d]) + 0.5_r) * config.spacing + (rnd() - 0.5_r) * config.jitter;
                    const real_t phi = rnd() * 2_r * M_PI;
                    const real_t angles[2] = {phi, phi + action::SPC::angleHOH};
                    for (int h = 0; h < 2; ++h)
                    {
                        pos(o + 1 + h, 0) = pos(o, 0) + action::SPC::eqDistanceHO * std::cos(angles[h]);
                        pos(o + 1 + h, 1) = pos(o, 1) + action::SPC::eqDistanceHO * std::sin(angles[h]);
                        pos(o + 1 + h, 2) = pos(o, 2);
                    }
                    for (int d = 0; d < 3; ++d)
                    {
                        const real_t v = (rnd() - 0.5_r) * 0.5_r;  // rigid translation: no velocity along the bonds
                        for (int a = 0; a < 3; ++a) vel(o + a, d) = v;
                    }
                    for (int a = 0; a < 3; ++a)
                    {
                        const bool oxygen = a == 0;
                        h_atoms.getType()(o + a) = oxygen ? 0 : 1;
                        h_atoms.getMass()(o + a) = oxygen ? action::SPC::massO : action::SPC::massH;
                        h_atoms.getCharge()(o + a) = oxygen ? action::SPC::chargeO : action::SPC::chargeH;
                        h_atoms.getRelativeMass()(o + a) = h_atoms.getMass()(o + a) / massMolecule;
                    }
                    h_molecules.getAtomsOffset()(m) = o;
                    h_molecules.getNumAtoms()(m) = 3;
                }
        h_atoms.numLocalAtoms = numAtoms;
        h_molecules.numLocalMolecules = numMolecules;
    }
    data::Atoms atoms(3 * molCapacity);
    data::deep_copy(atoms, h_atoms);
    data::Molecules molecules(molCapacity);
    data::deep_copy(molecules, h_molecules);

    // lambda = 1 everywhere: UpdateMolecules only supplies the centres of mass the list is built on
    auto weightingFunction = weighting_function::Slab(subdomain.getCenter(), 10_r * L, 1_r, 7);
    communication::MultiResGhostLayer ghostLayer;
    HalfVerletList moleculesVerletList;
    action::SPC spc;

    real_t maxAtomDisplacement = std::numeric_limits<real_t>::max();
    idx_t rebuildCounter = 0;
    for (idx_t step = 0; step < config.nsteps; ++step)
    {
        spc.enforcePositionalConstraints(molecules, atoms, config.dt);
        maxAtomDisplacement += action::VelocityVerlet::preForceIntegrate(atoms, config.dt);
        if (maxAtomDisplacement >= config.skin * 0.5_r)
        {
            maxAtomDisplacement = 0_r;
            action::UpdateMolecules::update(molecules, atoms, weightingFunction);
            ghostLayer.exchangeRealAtoms(molecules, atoms, subdomain);
            ghostLayer.createGhostAtoms(molecules, atoms, subdomain);
            action::UpdateMolecules::update(molecules, atoms, weightingFunction);
            moleculesVerletList.build(molecules.getPos(), 0, molecules.numLocalMolecules, config.neighborCutoff, config.cell_ratio,
                                      subdomain.minGhostCorner.data(), subdomain.maxGhostCorner.data(), config.estimatedMaxNeighbors);
            ++rebuildCounter;
        }
        else
        {
            ghostLayer.updateGhostAtoms(atoms, subdomain);
            action::UpdateMolecules::update(molecules, atoms, weightingFunction);
        }
        atoms.setForce(0_r);
        spc.applyForces(molecules, moleculesVerletList, atoms);
        ghostLayer.contributeBackGhostToReal(atoms);
        action::VelocityVerlet::postForceIntegrate(atoms, config.dt);
        spc.enforceVelocityConstraints(molecules, atoms, config.dt);
    }
    const real_t bondEnergy = spc.calcBondEnergy(molecules, atoms, 1000_r);

    data::deep_copy(h_atoms, atoms);
    auto pos = h_atoms.getPos();
    real_t maxBondError = 0_r;
    for (idx_t m = 0; m < numMolecules; ++m)
        for (int h = 1; h <= 2; ++h)
        {
            real_t d2 = 0_r;
            for (int d = 0; d < 3; ++d) d2 += (pos(3 * m, d) - pos(3 * m + h, d)) * (pos(3 * m, d) - pos(3 * m + h, d));
            maxBondError = std::max(maxBondError, std::abs(std::sqrt(d2) - action::SPC::eqDistanceHO));
        }
    std::printf(
        "{\"molecules\": %lld, \"steps\": %lld, \"rebuilds\": %lld, \"ghostAtoms\": %lld, \"ghostMolecules\": %lld, "
        "\"ELJ\": %.17g, \"ECoulomb\": %.17g, \"bondEnergy\": %.17g, \"maxBondError\": %.17g, \"x0\": [%.17g, %.17g, %.17g]}\n",
        static_cast<long long>(numMolecules), static_cast<long long>(config.nsteps), static_cast<long long>(rebuildCounter),
        static_cast<long long>(atoms.numGhostAtoms), static_cast<long long>(molecules.numGhostMolecules), spc.getEnergyLJ(),
        spc.getEnergyCoulomb(), bondEnergy, maxBondError, pos(0, 0), pos(0, 1), pos(0, 2));
    return 0;
}
