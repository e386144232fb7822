"""SPC water and the Coulomb pair potentials on the device against the reference's KATs
(mrmd/action/CoulombDSF.test.cpp, SPC.test.cpp) and the oracle."""
import ctypes as C
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from mrmd_b200 import api as a

    assert a.L().mrmd_b200_device_count() > 0
    return a


def water(api, pos, vel, mass, charge, rel_mass, typ, capacity=None):
    n = len(pos)
    atoms = api.Atoms.from_arrays(pos, vel, mass=mass, type=typ, relativeMass=rel_mass, capacity=capacity)
    atoms.set("charge", charge)
    cap_m = (capacity or n) // 3
    mols = api.Molecules(cap_m)
    mols.resize(cap_m)
    mols.set("atomsOffset", np.arange(n // 3) * 3)
    mols.set("numAtoms", np.full(n // 3, 3))
    mols.numLocalMolecules = n // 3
    return atoms, mols


def single_molecule(api):
    """mrmd/action/SPC.test.cpp:29-110"""
    eq, ang = 0.1, 109.47 / 180.0 * math.pi
    pos = np.array([(0, 0, 0), (eq, 0, 0), (eq * math.cos(ang), eq * math.sin(ang), 0)])
    m = np.array([15.999, 1.008, 1.008])
    return water(api, pos, np.zeros((3, 3)), m, np.array([-0.82, 0.41, 0.41]), m / m.sum(), np.array([0, 1, 1]))


def test_coulomb_dsf_kats(api, oracle):
    """CoulombDSF.test.cpp:63-117 on the device; device == oracle to 1e-12 of the largest term"""
    isp = 1.0 / math.sqrt(math.pi)
    rc, alpha, q1, q2 = 5.0, 0.1, 1.2, -1.3
    dsf = api.CoulombDSF(rc, alpha)
    xs = np.arange(1e-8, rc, 0.01)
    f, e = dsf.computeForce(xs * xs, q1, q2), dsf.computeEnergy(xs * xs, q1, q2)
    erfc = np.array([math.erfc(alpha * x) for x in xs])
    shift_f = math.erfc(alpha * rc) / (rc * rc) + 2 * alpha * isp * math.exp(-alpha * alpha * rc * rc) / rc
    f_ref = 138.935458 * q1 * q2 * (erfc / xs ** 2 + 2 * alpha * isp * np.exp(-alpha * alpha * xs * xs) / xs - shift_f) / xs
    e_ref = 138.935458 * q1 * q2 * (erfc / xs - math.erfc(alpha * rc) / rc + shift_f * (xs - rc))
    assert np.all(np.abs((f - f_ref) / f_ref) < 1e-4)
    sel = xs < rc - 1.0
    assert np.all(np.abs((e[sel] - e_ref[sel]) / e_ref[sel]) < 1e-5)
    fo, eo = np.zeros_like(xs), np.zeros_like(xs)
    d = np.ascontiguousarray(xs * xs)
    oracle.lib().or_coulomb_eval(1, rc, alpha, d.ctypes.data, len(d), q1, q2, fo.ctypes.data, eo.ctypes.data)
    big = 138.935458 * abs(q1 * q2)
    assert np.all(np.abs(f - fo) <= 1e-12 * big * (1.0 / xs ** 3 + 1.0)) and np.all(np.abs(e - eo) <= 1e-12 * big * (1.0 / xs + 1.0))
    sym = api.CoulombDSF(1.0, 0.1)
    pp, mm, pm, mp = (sym.computeForce(1.0, a, b) for a, b in ((1, 1), (-1, -1), (1, -1), (-1, 1)))
    assert pp == mm and pm == mp and pp == -pm  # Symmetry
    assert abs(api.CoulombDSF(1.5, 0.1).computeForce(1.5 * 1.5, 1.0, 1.0)) < 1e-5  # shift
    plain = api.Coulomb()
    assert plain.computeForce(4.0, 0.5, -0.25) == pytest.approx(138.935458 * 0.5 * -0.25 / 4.0, rel=1e-15)
    assert plain.computeEnergy(4.0, 0.5, -0.25) == pytest.approx(138.935458 * 0.5 * -0.25 / 2.0, rel=1e-15)


def test_spc_check_constraints(api):
    """SPC.test.cpp:131-152: the equilibrium molecule needs no constraint force"""
    atoms, mols = single_molecule(api)
    spc = api.SPC()
    spc.enforcePositionalConstraints(mols, atoms, 0.1)
    f = atoms.getForce()[:3]
    assert np.all(np.abs(f) < 1e-7)  # EXPECT_FLOAT_EQ(force + 1, 1)
    assert abs(spc.calcBondEnergy(mols, atoms, 1000.0)) < 1e-20
    spc.enforceVelocityConstraints(mols, atoms, 0.1)
    assert np.all(atoms.getVel()[:3] == 0.0)


@pytest.mark.parametrize("kind", [0, 1])
def test_spc_apply_forces_vs_oracle(api, oracle, kind):
    """4096 molecules, open boundaries: forces, both energies and the bond energy against the oracle on the same list"""
    from oracle.md_loop import spc_water_box

    pos, vel, mass, q, rm, typ, box = spc_water_box(16, jitter=0.05, seed=11)
    rng = np.random.default_rng(3)
    pos += (rng.random(pos.shape) - 0.5) * 0.01  # bonds off their equilibrium lengths: a non-zero bond energy
    n, nm = len(pos), len(pos) // 3
    atoms, mols = water(api, pos, vel, mass, q, rm, typ)
    w = api.Slab(0.5 * box, 100.0, 1.0, 7)
    api.UpdateMolecules.update(mols, atoms, w)
    vl = api.HalfVerletList()
    vl.build(mols, 0, nm, 1.3, 1.0, np.full(3, -0.2), box + 0.2, 230)
    counts, neigh = vl.to_host()
    spc = api.SPC(kind)
    atoms.setForce(0.0)
    spc.applyForces(mols, vl, atoms)

    L = oracle.lib()
    oa = np.zeros(n, dtype=oracle.ATOM)
    oa["pos"], oa["charge"], oa["mass"], oa["relMass"], oa["type"] = pos, q, mass, rm, typ
    om = np.zeros(nm, dtype=oracle.MOLECULE)
    om["atomsOffset"], om["numAtoms"] = np.arange(nm) * 3, 3
    oc, on = np.ascontiguousarray(counts[:nm], dtype=np.int32), np.ascontiguousarray(neigh[:nm], dtype=np.int32)
    en = np.zeros(2)
    L.or_spc_apply_forces(om.ctypes.data, nm, oc.ctypes.data, on.ctypes.data, on.shape[1], oa.ctypes.data, kind, en.ctypes.data)
    f = atoms.getForce()[:n]
    scale = np.abs(oa["force"]).max()
    assert np.abs(f - oa["force"]).max() <= 1e-10 * scale  # FP64 forces: summation order is the only difference
    assert abs(spc.getEnergyLJ() - en[0]) <= 1e-10 * abs(en[0])
    assert abs(spc.getEnergyCoulomb() - en[1]) <= 1e-10 * max(abs(en[1]), np.abs(oa["force"]).sum() * 1e-3)
    be = L.or_spc_bond_energy(om.ctypes.data, nm, oa.ctypes.data, n, 1000.0)
    assert be > 0 and abs(spc.calcBondEnergy(mols, atoms, 1000.0) - be) <= 1e-12 * be


def test_spc_generic_molecules_vs_oracle(api, oracle):
    """molecules with 1-4 atoms take the general kernel (the first atom carries the Lennard-Jones site)"""
    rng = np.random.default_rng(17)
    nm = 3000
    na = rng.integers(1, 5, nm)
    off = np.concatenate([[0], np.cumsum(na)[:-1]])
    n = int(na.sum())
    centres = rng.random((nm, 3)) * 5.0
    pos = np.repeat(centres, na, axis=0) + (rng.random((n, 3)) - 0.5) * 0.1
    q = rng.normal(size=n) * 0.5
    mass = 1.0 + rng.random(n)
    rel = mass / np.repeat(np.add.reduceat(mass, off), na)
    atoms = api.Atoms.from_arrays(pos, np.zeros((n, 3)), mass=mass, relativeMass=rel)
    atoms.set("charge", q)
    mols = api.Molecules(nm)
    mols.resize(nm)
    mols.set("atomsOffset", off)
    mols.set("numAtoms", na)
    mols.numLocalMolecules = nm
    api.UpdateMolecules.update(mols, atoms, api.Slab([2.5, 2.5, 2.5], 100.0, 1.0, 7))
    vl = api.HalfVerletList()
    vl.build(mols, 0, nm, 1.3, 1.0, np.full(3, -0.2), np.full(3, 5.2), 200)
    counts, neigh = vl.to_host()
    spc = api.SPC()
    atoms.setForce(0.0)
    spc.applyForces(mols, vl, atoms)
    oa = np.zeros(n, dtype=oracle.ATOM)
    oa["pos"], oa["charge"], oa["mass"], oa["relMass"] = pos, q, mass, rel
    om = np.zeros(nm, dtype=oracle.MOLECULE)
    om["atomsOffset"], om["numAtoms"] = off, na
    oc, on = np.ascontiguousarray(counts[:nm], dtype=np.int32), np.ascontiguousarray(neigh[:nm], dtype=np.int32)
    en = np.zeros(2)
    oracle.lib().or_spc_apply_forces(om.ctypes.data, nm, oc.ctypes.data, on.ctypes.data, on.shape[1], oa.ctypes.data, 0,
                                     en.ctypes.data)
    assert np.abs(atoms.getForce()[:n] - oa["force"]).max() <= 1e-10 * np.abs(oa["force"]).max()
    assert abs(spc.getEnergyLJ() - en[0]) <= 1e-10 * abs(en[0])
    assert abs(spc.getEnergyCoulomb() - en[1]) <= 1e-9 * np.abs(oa["force"]).max()


def test_spc_constrained_md_vs_oracle(api):
    """1000 molecules in a periodic box, 30 constrained steps with list rebuilds: SHAKE -> velocity Verlet ->
    MultiResGhostLayer -> UpdateMolecules -> SPC::applyForces -> RATTLE against the oracle's run of the same loop"""
    from oracle.md_loop import OracleSpcMD, spc_water_box

    pos, vel, mass, q, rm, typ, box = spc_water_box(10)
    steps, dt, skin = 30, 0.0005, 0.02  # a thin skin, so that the list is rebuilt a few times
    ref = OracleSpcMD(pos, vel, mass, q, rm, typ, box, dt=dt, skin=skin)
    st = ref.run(steps)
    assert st["rebuilds"] >= 2

    n, nm = len(pos), len(pos) // 3
    cap = 3 * len(ref.mols)
    atoms, mols = water(api, pos, vel, mass, q, rm, typ, capacity=cap)
    cutoff = 1.2 + skin
    sub = api.Subdomain([0, 0, 0], box, cutoff)
    w = api.Slab(0.5 * box, 10.0 * box[0], 1.0, 7)
    ghost, vl, spc = api.MultiResGhostLayer(), api.HalfVerletList(), api.SPC()
    max_disp, rebuilds = np.finfo(np.float64).max, 0
    for _ in range(steps):
        spc.enforcePositionalConstraints(mols, atoms, dt)
        max_disp += api.VelocityVerlet.preForceIntegrate(atoms, dt)
        if max_disp >= skin * 0.5:
            max_disp = 0.0
            api.UpdateMolecules.update(mols, atoms, w)
            ghost.exchangeRealAtoms(mols, atoms, sub)
            ghost.createGhostAtoms(mols, atoms, sub)
            api.UpdateMolecules.update(mols, atoms, w)
            vl.build(mols, 0, mols.numLocalMolecules, cutoff, 1.0, list(sub.minGhostCorner), list(sub.maxGhostCorner), 220)
            rebuilds += 1
        else:
            ghost.updateGhostAtoms(atoms, sub)
            api.UpdateMolecules.update(mols, atoms, w)
        atoms.setForce(0.0)
        spc.applyForces(mols, vl, atoms)
        ghost.contributeBackGhostToReal(atoms)
        api.VelocityVerlet.postForceIntegrate(atoms, dt)
        spc.enforceVelocityConstraints(mols, atoms, dt)
    assert rebuilds == st["rebuilds"]
    assert atoms.numGhostAtoms == ref.ng and mols.numGhostMolecules == ref.mg
    assert abs(spc.getEnergyLJ() - st["energyLJ"]) <= 1e-9 * abs(st["energyLJ"])
    assert abs(spc.getEnergyCoulomb() - st["energyCoulomb"]) <= 1e-9 * abs(st["energyCoulomb"])
    assert np.abs(atoms.getPos()[:n] - ref.atoms["pos"][:n]).max() <= 1e-9
    assert np.abs(atoms.getVel()[:n] - ref.atoms["vel"][:n]).max() <= 1e-8 * np.abs(ref.atoms["vel"][:n]).max()
    p = atoms.getPos()[:n].reshape(-1, 3, 3)
    assert np.abs(np.linalg.norm(p[:, 0] - p[:, 1], axis=1) - 0.1).max() < 1e-5  # SHAKE holds the O-H bonds
