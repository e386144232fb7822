"""analysis:: diagnostics on the device against the reference's KATs and the oracle."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from mrmd_b200 import api as a

    assert a.L().mrmd_b200_device_count() > 0
    return a


def test_kinetic_energy_kat(api):
    """mrmd/analysis/KineticEnergy.test.cpp:23-53"""
    atoms = api.Atoms.from_arrays(np.zeros((3, 3)), np.array([(2, 0, 0), (0, -8, 0), (0, 0, 16.0)]),
                                  mass=np.array([1.0, 2.0, 0.5]))
    assert api.analysis.getKineticEnergy(atoms) == pytest.approx((4 + 2 * 64 + 0.5 * 256) * 0.5, rel=1e-7)
    assert api.analysis.getMeanKineticEnergy(atoms) == pytest.approx((4 + 2 * 64 + 0.5 * 256) * 0.5 / 3, rel=1e-7)


def test_system_momentum_kat(api):
    """mrmd/analysis/SystemMomentum.test.cpp:23-42"""
    atoms = api.Atoms.from_arrays(np.zeros((2, 3)), np.array([(2, 3, 4), (-4, -8, -16.0)]))
    assert np.allclose(api.analysis.getSystemMomentum(atoms), [-2, -5, -12], rtol=1e-7)


def test_msd_no_displacement_kat(api):
    """mrmd/analysis/MeanSquareDisplacement.test.cpp:28-35"""
    atoms = api.Atoms.from_arrays(np.array([(1.0, 2.0, 3.0)]))
    sub = api.Subdomain([0, 0, 0], [10, 10, 10], 1.0)
    msd = api.analysis.MeanSquareDisplacement()
    msd.reset(atoms)
    assert msd.calc(atoms, sub) == 0.0


def test_diagnostics_vs_oracle(api, oracle):
    rng = np.random.default_rng(77)
    n, ng = 50000, 3000
    box = np.array([30.0, 20.0, 25.0])
    pos = rng.random((n + ng, 3)) * box
    vel = rng.normal(size=(n + ng, 3))
    force = rng.normal(size=(n + ng, 3)) * 5
    mass = 0.5 + rng.random(n + ng)
    atoms = api.Atoms.from_arrays(pos, vel, mass=mass)
    atoms.set("force", force)
    atoms.numLocalAtoms, atoms.numGhostAtoms = n, ng
    sub = api.Subdomain([0, 0, 0], box, 2.6)
    osub = oracle.subdomain([0, 0, 0], box, 2.6)
    oa = np.zeros(n + ng, dtype=oracle.ATOM)
    oa["pos"], oa["vel"], oa["force"], oa["mass"] = pos, vel, force, mass
    L = oracle.lib()
    ke = L.or_kinetic_energy(oa.ctypes.data, n)
    assert abs(api.analysis.getKineticEnergy(atoms) - ke) <= 1e-12 * abs(ke)
    mom = np.zeros(3)
    L.or_system_momentum(oa.ctypes.data, n, mom.ctypes.data)
    assert np.abs(api.analysis.getSystemMomentum(atoms) - mom).max() <= 1e-11 * np.abs(vel[:n]).sum(axis=0).max()
    p = L.or_pressure(oa.ctypes.data, n + ng, C.byref(osub))
    assert abs(api.analysis.getPressure(atoms, sub) - p) <= 1e-11 * (np.abs(force * pos).sum() / (3 * box.prod()))
    # mean square displacement: moves up to +-0.6 box lengths exercise the reference's fold
    msd = api.analysis.MeanSquareDisplacement()
    msd.reset(atoms)
    moved = pos.copy()
    moved[:n] += (rng.random((n, 3)) - 0.5) * 1.2 * box
    atoms.set("pos", moved)
    oa["pos"] = moved
    init = np.ascontiguousarray(pos[:n])
    want = L.or_msd(oa.ctypes.data, init.ctypes.data, n, C.byref(osub))
    assert abs(msd.calc(atoms, sub) - want) <= 1e-12 * want
    # the number of items must not change between reset and calc (MRMD_HOST_CHECK_EQUAL)
    atoms.numLocalAtoms = n - 1
    with pytest.raises(RuntimeError):
        msd.calc(atoms, sub)
