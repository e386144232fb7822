"""Berendsen thermostat / barostat and SHAKE / RATTLE molecule constraints on the device
against the reference's KATs (mrmd/action/BerendsenThermostat.test.cpp, BerendsenBarostat.test.cpp, Shake.test.cpp) and
the oracle."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from mrmd_b200 import api as a

    assert a.L().mrmd_b200_device_count() > 0
    return a


def single_atom(api):
    """mrmd/test/SingleAtom.hpp:28-55"""
    atoms = api.Atoms.from_arrays(np.array([(2.0, 3.0, 4.0)]), np.array([(7.0, 5.0, 3.0)]), mass=1.5)
    atoms.set("force", np.array([(9.0, 7.0, 8.0)]))
    return atoms


def diamond(api):
    """mrmd/test/DiamondFixture.hpp:26-106"""
    atoms = api.Atoms.from_arrays(np.array([(1.0, 0, 0), (0, 1.0, 0), (-1.0, 0, 0), (0, -1.0, 0)]),
                                  mass=np.array([1.0, 3.0, 1.0, 3.0]), relativeMass=np.array([0.25, 0.75, 0.25, 0.75]))
    mols = api.Molecules(2)
    mols.set("atomsOffset", [0, 2])
    mols.set("numAtoms", [2, 2])
    mols.numLocalMolecules = 2
    return atoms, mols


def whole_molecule(api):
    mols = api.Molecules(1)
    mols.set("atomsOffset", [0])
    mols.set("numAtoms", [4])
    mols.numLocalMolecules = 1
    return mols


def integrate_position(atoms, dt):
    """integratePosition of mrmd/action/Shake.test.cpp:27-46 (host side)"""
    n = atoms.numLocalAtoms
    pos, vel, force, mass = atoms.getPos()[:n], atoms.getVel()[:n], atoms.getForce()[:n], atoms.getMass()[:n]
    return pos + dt * vel + (0.5 * dt * dt / mass)[:, None] * force


def temperature(api, atoms):
    return api.analysis.getMeanKineticEnergy(atoms) * (2.0 / 3.0)


def test_berendsen_thermostat_kats(api):
    atoms = single_atom(api)
    api.BerendsenThermostat.apply(atoms, temperature(api, atoms), 3.8, 0.0)
    assert temperature(api, atoms) == pytest.approx(41.5, rel=1e-7)
    api.BerendsenThermostat.apply(atoms, temperature(api, atoms), 3.8, 1.0)
    assert temperature(api, atoms) == pytest.approx(3.8, rel=1e-7)
    api.BerendsenThermostat.apply(atoms, 0.0, 3.8, 1.0)  # T <= 0: no-op
    assert temperature(api, atoms) == pytest.approx(3.8, rel=1e-7)
    with pytest.raises(RuntimeError):
        api.BerendsenThermostat.apply(atoms, 1.0, 0.0, 1.0)


def test_berendsen_barostat_kats(api):
    atoms = single_atom(api)
    sub = api.Subdomain([0, 0, 0], [1, 1, 1], 0.1)
    api.BerendsenBarostat.apply(atoms, 1.0, 3.8, 0.0, sub)
    assert np.allclose(list(sub.maxCorner), 1.0) and np.allclose(atoms.getPos()[0], [2, 3, 4])
    api.BerendsenBarostat.apply(atoms, 2.0, 1.0, 1.0, sub)
    mu = 2.0 ** (1.0 / 3.0)
    assert np.allclose(list(sub.maxCorner), mu, rtol=1e-7) and np.allclose(atoms.getPos()[0], np.array([2, 3, 4]) * mu, rtol=1e-7)
    api.BerendsenBarostat.apply(atoms, 2.0, 1.0, 1.0, sub, stretchX=True, stretchY=False, stretchZ=False)
    assert np.allclose(list(sub.maxCorner), [mu * mu, mu, mu], rtol=1e-7)
    assert np.allclose(atoms.getPos()[0], [2 * mu * mu, 3 * mu, 4 * mu], rtol=1e-7)


@pytest.mark.parametrize("eq,sign", [(1.0, -1.0), (2.0, 1.0)])
def test_shake_single_constraint(api, eq, sign):
    """Shake.test.cpp:59-113 (Attraction / Repulsion)"""
    atoms, _ = diamond(api)
    mc = api.MoleculeConstraints(4, 1)
    mc.setConstraints([(0, 1, eq)])
    mc.enforcePositionalConstraints(whole_molecule(api), atoms, 0.1)
    f = atoms.getForce()
    assert np.allclose(f[0], -f[1]) and sign * f[0][0] > 0 and sign * f[0][1] < 0
    new = integrate_position(atoms, 0.1)
    assert np.linalg.norm(new[0] - new[1]) == pytest.approx(eq, rel=1e-7)


@pytest.mark.parametrize("eq", [1.0, 2.0])
def test_shake_ring(api, eq):
    """Shake.test.cpp:132-221 (Shrink / Grow): ten iterations over the ring 0-1-2-3-0"""
    atoms, _ = diamond(api)
    mc = api.MoleculeConstraints(4, 10)
    mc.setConstraints([(0, 1, eq), (1, 2, eq), (2, 3, eq), (3, 0, eq)])
    mc.enforcePositionalConstraints(whole_molecule(api), atoms, 0.1)
    new = integrate_position(atoms, 0.1)
    for i in range(4):
        assert np.linalg.norm(new[i] - new[(i + 1) % 4]) == pytest.approx(eq, rel=1e-6)


def test_shake_molecules_kat(api):
    """Shake.test.cpp:223-249"""
    atoms, mols = diamond(api)
    mc = api.MoleculeConstraints(2, 1)
    mc.setConstraints([(0, 1, 1.0)])
    mc.enforcePositionalConstraints(mols, atoms, 0.1)
    f = atoms.getForce()
    assert np.allclose(f[0], -f[1]) and f[0][0] < 0 < f[0][1]
    assert np.allclose(f[2], -f[3]) and f[2][0] > 0 > f[2][1]
    mc.setConstraints([(0, 2, 1.0)])  # not enough atoms in molecule to satisfy bond
    with pytest.raises(RuntimeError, match="not enough atoms"):
        mc.enforcePositionalConstraints(mols, atoms, 0.1)


@pytest.mark.parametrize("a_per,order", [(4, "lexicographic"), (4, "reversed"), (6, "lexicographic")])
def test_constraints_vs_oracle(api, oracle, a_per, order):
    """20 000 tetramers with six bonds each, three SHAKE iterations and the RATTLE projection, against the oracle.
    a_per = 4 with the six bonds listed as all pairs in lexicographic order takes the register-resident all-pairs kernels,
    the same bonds listed backwards the fused shared-memory kernels (any bonds among the first four atoms of a molecule);
    a_per = 6 hangs two more atoms on every tetramer (bonds 3-4, 4-5) and takes the kernel sequence of the reference."""
    rng = np.random.default_rng(8)
    M = 20000
    N = M * a_per
    centres = rng.random((M, 3)) * 60.0
    tet = np.array([(1, 1, 1), (1, -1, -1), (-1, 1, -1), (-1, -1, 1), (-1.8, -1.9, 2.1), (-2.9, -2.8, 3.2)]) * (0.5 / np.sqrt(2.0))
    pos = (centres[:, None, :] + tet[None, :a_per, :] * (1.0 + 0.1 * rng.random((M, a_per, 1)))).reshape(-1, 3)
    vel, force = rng.normal(size=(N, 3)), rng.normal(size=(N, 3)) * 3
    mass = 0.5 + rng.random(N)
    bonds = [(i, j, 1.0) for i in range(4) for j in range(i + 1, 4)] + [(i, i + 1, 0.6) for i in range(3, a_per - 1)]
    if order == "reversed":
        bonds = bonds[::-1]
    dt = 0.002

    atoms = api.Atoms.from_arrays(pos, vel, mass=mass)
    atoms.set("force", force)
    mols = api.Molecules(M)
    mols.set("atomsOffset", np.arange(M) * a_per)
    mols.set("numAtoms", np.full(M, a_per))
    mols.numLocalMolecules = M
    mc = api.MoleculeConstraints(a_per, 3)
    mc.setConstraints(bonds)
    mc.enforcePositionalConstraints(mols, atoms, dt)

    L = oracle.lib()
    oa = np.zeros(N, dtype=oracle.ATOM)
    oa["pos"], oa["vel"], oa["force"], oa["mass"] = pos, vel, force, mass
    om = np.zeros(M, dtype=oracle.MOLECULE)
    om["atomsOffset"], om["numAtoms"] = np.arange(M) * a_per, a_per
    bidx = np.ascontiguousarray([[b[0], b[1]] for b in bonds], dtype=np.int64)
    beq = np.ascontiguousarray([b[2] for b in bonds])
    assert L.or_shake_positional(om.ctypes.data, M, oa.ctypes.data, N, bidx.ctypes.data, beq.ctypes.data, len(bonds), 3, dt) == 0
    f = atoms.getForce()
    assert np.abs(f - oa["force"]).max() <= 1e-9 * np.abs(oa["force"]).max()

    mc.enforceVelocityConstraints(mols, atoms, dt)
    assert L.or_shake_velocity(om.ctypes.data, M, oa.ctypes.data, bidx.ctypes.data, len(bonds)) == 0
    assert np.abs(atoms.getVel() - oa["vel"]).max() <= 1e-12 * np.abs(oa["vel"]).max()

    api.BerendsenThermostat.apply(atoms, 1.3, 1.5, 0.1)
    L.or_berendsen_thermostat(oa.ctypes.data, N, 1.3, 1.5, 0.1)
    assert np.array_equal(atoms.getVel(), oa["vel"]) or np.abs(atoms.getVel() - oa["vel"]).max() < 1e-13
    sub, osub = api.Subdomain([0, 0, 0], [60, 60, 60], 2.6), oracle.subdomain([0, 0, 0], [60, 60, 60], 2.6)
    api.BerendsenBarostat.apply(atoms, 0.8, 1.0, 0.05, sub)
    L.or_berendsen_barostat(oa.ctypes.data, N, 0.8, 1.0, 0.05, C.byref(osub), 1, 1, 1)
    assert np.abs(atoms.getPos() - oa["pos"]).max() < 1e-12
    assert np.allclose(list(sub.maxCorner), list(osub.maxCorner), rtol=1e-15)
    assert np.allclose(list(sub.minGhostCorner), list(osub.minGhostCorner), rtol=1e-15)


def test_limit_acceleration_and_velocity(api, oracle):
    """LimitAcceleration.test.cpp:27-38, LimitVelocity.test.cpp:27-38, then 100 000 random atoms against the oracle"""
    atoms = single_atom(api)
    api.limitAccelerationPerComponent(atoms, 0.5)
    assert np.allclose(atoms.getForce()[0], 0.75, rtol=1e-7)  # EXPECT_FLOAT_EQ(force(0, d), 0.75_r)
    api.limitVelocityPerComponent(atoms, 0.5)
    assert np.array_equal(atoms.getVel()[0], [0.5, 0.5, 0.5])

    rng = np.random.default_rng(21)
    n = 100000
    pos, vel, force = rng.random((n, 3)), rng.normal(size=(n, 3)) * 2, rng.normal(size=(n, 3)) * 5
    mass = 0.5 + rng.random(n)
    atoms = api.Atoms.from_arrays(pos, vel, mass=mass)
    atoms.set("force", force)
    oa = np.zeros(n, dtype=oracle.ATOM)
    oa["pos"], oa["vel"], oa["force"], oa["mass"] = pos, vel, force, mass
    api.limitAccelerationPerComponent(atoms, 1.5)
    api.limitVelocityPerComponent(atoms, 1.0)
    oracle.lib().or_limit_acceleration(oa.ctypes.data, n, 1.5)
    oracle.lib().or_limit_velocity(oa.ctypes.data, n, 1.0)
    assert np.array_equal(atoms.getVel()[:n], oa["vel"])
    assert np.abs(atoms.getForce()[:n] - oa["force"]).max() <= 1e-15 * np.abs(oa["force"]).max()
    assert np.abs(atoms.getForce()[:n] / mass[:, None]).max() <= 1.5 * (1 + 1e-15)
