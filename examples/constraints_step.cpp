// One constrained velocity-Verlet step through the mirror of the reference API (include/mrmd/): the flow of the
// reference's tests/Constraints/Constraints.cpp:25-72 (SHAKE before the integrator, RATTLE after it) on the diamond
// fixture (mrmd/test/DiamondFixture.hpp), followed by a Berendsen thermostat and barostat call as tests/NVT and tests/NPT
// place them.
//
//   g++ -std=c++20 -O2 -Iinclude/mrmd examples/constraints_step.cpp -Lmrmd_b200 -lmrmd_b200 -Wl,-rpath,$PWD/mrmd_b200
#include <cmath>
#include <cstdio>

#include "action/BerendsenBarostat.hpp"
#include "action/BerendsenThermostat.hpp"
#include "action/LimitAcceleration.hpp"
#include "action/LimitVelocity.hpp"
#include "action/Shake.hpp"
#include "action/VelocityVerlet.hpp"
#include "analysis/KineticEnergy.hpp"
#include "analysis/Pressure.hpp"
#include "data/Atoms.hpp"
#include "data/Bond.hpp"
#include "data/Molecules.hpp"
#include "data/Subdomain.hpp"
#include "datatypes.hpp"
#include "util/ExponentialMovingAverage.hpp"

using namespace mrmd;

int main()
{
    // 2 molecules with 2 atoms each (DiamondFixture)
    data::HostAtoms h_atoms(4);
    const real_t positions[4][3] = {{1_r, 0_r, 0_r}, {0_r, 1_r, 0_r}, {-1_r, 0_r, 0_r}, {0_r, -1_r, 0_r}};
    const real_t masses[4] = {1_r, 3_r, 1_r, 3_r};
    for (idx_t i = 0; i < 4; ++i)
    {
        for (int d = 0; d < 3; ++d) h_atoms.getPos()(i, d) = positions[i][d];
        h_atoms.getMass()(i) = masses[i];
        h_atoms.getRelativeMass()(i) = masses[i] / 4_r;
    }
    h_atoms.numLocalAtoms = 4;
    data::Atoms atoms(4);
    data::deep_copy(atoms, h_atoms);

    data::HostMolecules h_molecules(2);
    h_molecules.getAtomsOffset()(0) = 0;
    h_molecules.getAtomsOffset()(1) = 2;
    h_molecules.getNumAtoms()(0) = 2;
    h_molecules.getNumAtoms()(1) = 2;
    h_molecules.numLocalMolecules = 2;
    data::Molecules molecules(2);
    data::deep_copy(molecules, h_molecules);

    const auto dt = 0.1_r;
    data::BondView bonds(1);
    bonds[0].idx = 0;
    bonds[0].jdx = 1;
    bonds[0].eqDistance = 1_r;
    action::MoleculeConstraints mc(2, 1);
    mc.setConstraints(bonds);

    mc.enforcePositionalConstraints(molecules, atoms, dt);
    action::VelocityVerlet::preForceIntegrate(atoms, dt);
    atoms.setForce(0_r);
    // the limiters every reference driver includes (examples/02:28-29); wide open here
    action::limitAccelerationPerComponent(atoms, 1e9_r);
    action::limitVelocityPerComponent(atoms, 1e9_r);
    action::VelocityVerlet::postForceIntegrate(atoms, dt);
    mc.enforceVelocityConstraints(molecules, atoms, dt);

    data::deep_copy(h_atoms, atoms);
    auto pos = h_atoms.getPos();
    auto vel = h_atoms.getVel();
    real_t dx[3], dv[3];
    for (int d = 0; d < 3; ++d)
    {
        dx[d] = pos(0, d) - pos(1, d);
        dv[d] = vel(0, d) - vel(1, d);
    }
    const real_t dist = std::sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]);
    const real_t relVel = (dx[0] * dv[0] + dx[1] * dv[1] + dx[2] * dv[2]) / dist;

    data::Subdomain subdomain({-2_r, -2_r, -2_r}, {2_r, 2_r, 2_r}, 0.5_r);
    util::ExponentialMovingAverage averageT(0.5_r);  // tests/NVT/NVT.cpp:132
    averageT << analysis::getMeanKineticEnergy(atoms) * (2_r / 3_r);
    const real_t T0 = averageT;
    action::BerendsenThermostat::apply(atoms, T0, 2_r * T0, 1_r);
    const real_t T1 = analysis::getMeanKineticEnergy(atoms) * (2_r / 3_r);
    const real_t p = analysis::getPressure(atoms, subdomain);
    action::BerendsenBarostat::apply(atoms, 2_r, 1_r, 1_r, subdomain);

    std::printf("{\"dist\": %.17g, \"relVel\": %.17g, \"T0\": %.17g, \"T1\": %.17g, \"p\": %.17g, \"maxCorner\": %.17g}\n", dist,
                relVel, T0, T1, p, subdomain.maxCorner[0]);
    return 0;
}
