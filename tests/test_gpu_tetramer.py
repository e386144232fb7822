"""BASELINE.json configs[3] at test size: AdResS Lennard-Jones tetramers with a spherical atomistic region, Langevin
NVT, through the step-loop driver (mrmd_b200_md_* with atomsPerMolecule = 4) against the oracle's loop of the same
operators (oracle/md_loop.py:OracleAdressMD), with and without SHAKE / RATTLE on the six bonds of every tetramer."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from mrmd_b200 import api as a

    assert a.L().mrmd_b200_device_count() > 0
    return a


@pytest.mark.parametrize("iterations", [0, 3])
def test_tetramer_md_vs_oracle_loop(api, oracle, iterations):
    from mrmd_b200.workloads import tetramer_system
    from oracle.md_loop import OracleAdressMD

    pos, vel, box = tetramer_system(10, seed=77)
    n, steps = len(pos), 60
    sub = api.Subdomain([0, 0, 0], box, 2.6)
    w = api.Spherical(0.5 * box, 4.0, 3.0, 2)
    ow = oracle.make_weight(oracle.WEIGHT_SPHERICAL, 0.5 * box, 4.0, 3.0, 2)
    atoms = api.Atoms.from_arrays(pos, vel, mass=1.0, relativeMass=0.25)
    md = api.MolecularDynamics(atoms, sub, langevin=True, zeta=20.0, temperature=1.5, seed=5, cellSort=False, fullList=0,
                               adress=True, weight=w, maxNeighbors=40, atomsPerMolecule=4,
                               numConstraintIterations=iterations, bondLength=1.0)
    omd = OracleAdressMD(pos, vel, box, ow, langevin=True, zeta=20.0, temperature=1.5, seed=5, max_neigh=40,
                         atoms_per_mol=4, constraint_iterations=iterations, bond_length=1.0)
    st, res = md.run(steps), omd.run(steps)
    assert st["rebuilds"] == res["rebuilds"] and st["rebuilds"] >= 2
    assert st["pairInteractions"] == res["pairInteractions"] > 0
    assert st["numGhost"] == omd.ng
    assert abs(st["energy"] - res["energy"]) <= 1e-8 * abs(res["energy"])
    assert np.abs(atoms.getPos()[:n] - omd.atoms["pos"][:n]).max() < 1e-9
    assert np.abs(atoms.getVel()[:n] - omd.atoms["vel"][:n]).max() < 1e-8
    if iterations:
        p = atoms.getPos()[:n].reshape(-1, 4, 3)
        d = np.linalg.norm(p[:, :, None, :] - p[:, None, :, :], axis=-1)[:, np.triu_indices(4, 1)[0], np.triu_indices(4, 1)[1]]
        assert np.abs(d - 1.0).max() < 5e-3  # three SHAKE iterations hold the six bonds to a fraction of a percent


def test_tetramer_config_is_validated(api):
    from mrmd_b200.workloads import tetramer_system

    pos, vel, box = tetramer_system(4)
    sub = api.Subdomain([0, 0, 0], box, 2.6)
    atoms = api.Atoms.from_arrays(pos, vel, mass=1.0, relativeMass=0.25)
    w = api.Spherical(0.5 * box, 1.0, 1.0, 2)
    with pytest.raises(RuntimeError, match="multi-atom molecules"):
        api.MolecularDynamics(atoms, sub, fullList=2, adress=True, weight=w, atomsPerMolecule=4)
    with pytest.raises(RuntimeError, match="multi-atom molecules"):
        api.MolecularDynamics(atoms, sub, fullList=0, cellSort=True, adress=True, weight=w, atomsPerMolecule=4)
    with pytest.raises(RuntimeError, match="multiple of atomsPerMolecule"):
        api.MolecularDynamics(atoms, sub, fullList=0, cellSort=False, adress=True, weight=w, atomsPerMolecule=3)
