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


@pytest.mark.parametrize("iterations", [0, 3])
def test_tetramer_tiled_md_vs_oracle_loop(api, oracle, iterations):
    """the tiled molecule path (fullList = 2: cell sort on the centres of mass, tiled list on them, one force kernel with
    four lanes per molecule, no ghost molecules) against the same oracle loop; the molecules are re-ordered by the
    sort, atoms are matched through their global ids"""
    from mrmd_b200.workloads import tetramer_system
    from oracle.md_loop import OracleAdressMD

    pos, vel, box = tetramer_system(10, seed=77)
    n, steps = len(pos), 60
    sub = api.Subdomain([0, 0, 0], box, 2.6)
    w = api.Spherical(0.5 * box, 4.0, 3.0, 2)
    ow = oracle.make_weight(oracle.WEIGHT_SPHERICAL, 0.5 * box, 4.0, 3.0, 2)
    atoms = api.Atoms.from_arrays(pos, vel, mass=1.0, relativeMass=0.25)
    md = api.MolecularDynamics(atoms, sub, langevin=True, zeta=20.0, temperature=1.5, seed=5, cellSort=True, fullList=2,
                               adress=True, weight=w, maxNeighbors=40, atomsPerMolecule=4,
                               numConstraintIterations=iterations, bondLength=1.0)
    omd = OracleAdressMD(pos, vel, box, ow, langevin=True, zeta=20.0, temperature=1.5, seed=5, max_neigh=40,
                         atoms_per_mol=4, constraint_iterations=iterations, bond_length=1.0)
    st, res = md.run(steps), omd.run(steps)
    assert st["rebuilds"] == res["rebuilds"] and st["rebuilds"] >= 2
    assert st["pairInteractions"] == res["pairInteractions"] > 0
    assert st["numGhost"] == 0
    assert abs(st["energy"] - res["energy"]) <= 1e-8 * abs(res["energy"])
    ids = atoms.get("id")[:n]
    assert np.array_equal(np.sort(ids), np.arange(n))
    assert np.array_equal(ids.reshape(-1, 4)[:, 0] % 4, np.zeros(n // 4, dtype=ids.dtype))  # molecules stay blocks
    d = atoms.getPos()[:n] - omd.atoms["pos"][ids]
    d -= box * np.round(d / box)
    assert np.abs(d).max() < 1e-9
    assert np.abs(atoms.getVel()[:n] - omd.atoms["vel"][ids]).max() < 1e-8


def test_tetramer_tiled_list_and_force_vs_generic(api):
    """one force evaluation: the tiled molecule kernel against the operator sequence on the Cabana-layout list
    (UpdateMolecules, LJ_IdealGas::run on the half list over MultiResGhostLayer ghosts, ContributeMoleculeForceToAtoms,
    contributeBackGhostToReal), with the compensation histograms sampled, and the stored pair sets equal"""
    import ctypes as C

    from mrmd_b200 import _lib
    from mrmd_b200.workloads import tetramer_system

    pos, vel, box = tetramer_system(9, seed=3)
    rng = np.random.default_rng(11)
    pos = pos + np.repeat((rng.random((len(pos) // 4, 3)) - 0.5) * 0.6, 4, axis=0)  # off the lattice, molecules rigid
    n, nm, cut = len(pos), len(pos) // 4, 2.6
    sub = api.Subdomain([0, 0, 0], box, cut)
    w = api.Spherical(0.5 * box, 3.0, 3.0, 2)
    L = api.L()

    def molecules_of(atoms):
        m = api.Molecules(2 * nm)
        m.resize(nm)
        m.set("atomsOffset", np.arange(nm, dtype=np.int64) * 4)
        m.set("numAtoms", np.full(nm, 4, dtype=np.int64))
        m._set_counts(nm, 0)
        return m

    # generic path
    ga = api.Atoms.from_arrays(pos, vel, mass=1.0, relativeMass=0.25, capacity=4 * n)
    gm = molecules_of(ga)
    api.UpdateMolecules.update(gm, ga, w)
    ghost = api.MultiResGhostLayer()
    ghost.exchangeRealAtoms(gm, ga, sub)
    ghost.createGhostAtoms(gm, ga, sub)
    api.UpdateMolecules.update(gm, ga, w)
    hv = api.HalfVerletList()
    hv.build(gm, 0, nm, cut, 1.0, sub.minGhostCorner, sub.maxGhostCorner, 40)
    gl = api.LJ_IdealGas(0.7, 2.5, 1.0, 1.0, True)
    ga.setForce(0.0)
    gm.setForce(0.0)
    e_ref = gl.run(gm, hv, ga)
    api.ContributeMoleculeForceToAtoms.update(gm, ga)
    ghost.contributeBackGhostToReal(ga)
    f_ref = ga.getForce()[:n]

    # tiled path
    ta = api.Atoms.from_arrays(pos, vel, mass=1.0, relativeMass=0.25)
    tm = molecules_of(ta)
    api.UpdateMolecules.update(tm, ta, w)
    api.MultiResGhostLayer().exchangeRealAtoms(tm, ta, sub)
    api.UpdateMolecules.update(tm, ta, w)
    delta = np.array([cut, cut, 0.25 * cut])
    lo, hi = np.asarray(sub.minCorner, dtype=np.float64), np.asarray(sub.maxCorner, dtype=np.float64)
    _lib.check(L.mrmd_b200_molecules_cell_sort_with_atoms(tm.h, ta.h, 4, delta.ctypes.data, lo.ctypes.data, hi.ctypes.data, None))
    fv = api.FullVerletList()
    _lib.check(L.mrmd_b200_verlet_build_periodic_molecules(fv.h, tm.h, C.byref(sub), cut, 1.0, 40, 4, None))
    assert fv.info()["totalPairs"] == 2 * int(hv.to_host()[0].sum())
    tl = api.LJ_IdealGas(0.7, 2.5, 1.0, 1.0, True)
    ta.setForce(0.0)
    api.UpdateMolecules.update(tm, ta, w)
    energy, pairs = C.c_double(), C.c_int64()
    _lib.check(L.mrmd_b200_adress_run_periodic_molecules(tl.h, tm.h, ta.h, fv.h, C.byref(w), 4, C.byref(energy), C.byref(pairs), None))
    ids = ta.get("id")[:n]
    f = ta.getForce()[:n]
    scale = np.abs(f_ref).max()
    assert scale > 0
    assert np.abs(f - f_ref[ids]).max() <= 1e-10 * scale
    assert abs(energy.value - e_ref) <= 1e-11 * abs(e_ref)
    for kind in (1, 2):  # compensationEnergy and its counter were sampled (run 0 samples)
        assert np.allclose(tl._hist(kind), gl._hist(kind), rtol=1e-10, atol=1e-12)


def test_tetramer_config_is_validated(api):
    from mrmd_b200.workloads import tetramer_system

    pos, vel, box = tetramer_system(4)
    sub = api.Subdomain([0, 0, 0], box, 2.6)
    atoms = api.Atoms.from_arrays(pos, vel, mass=1.0, relativeMass=0.25)
    w = api.Spherical(0.5 * box, 1.0, 1.0, 2)
    with pytest.raises(RuntimeError, match="half list"):
        api.MolecularDynamics(atoms, sub, fullList=1, adress=True, weight=w, atomsPerMolecule=4)
    with pytest.raises(RuntimeError, match="multi-atom molecules"):
        api.MolecularDynamics(atoms, sub, fullList=2, adress=True, weight=w, atomsPerMolecule=2)
    with pytest.raises(RuntimeError, match="multi-atom molecules"):
        api.MolecularDynamics(atoms, sub, fullList=0, cellSort=True, adress=True, weight=w, atomsPerMolecule=4)
    with pytest.raises(RuntimeError, match="multiple of atomsPerMolecule"):
        api.MolecularDynamics(atoms, sub, fullList=0, cellSort=False, adress=True, weight=w, atomsPerMolecule=3)
