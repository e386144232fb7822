"""Parity at BASELINE.json's full sizes.

configs[1] (1 M atoms): the oracle finishes one force evaluation of the full system in seconds, so the tiled fast path
is compared with it directly -- stored neighbour sets bit exact (as sorted (partner, image) sets per atom), pair count
exact, forces to 1e-10, energy and virial to 1e-12 -- and through size-independent properties of a 200-step NVE run
(momentum conservation, bounded energy drift, Newton's third law).
configs[2] (8 M atoms, AdResS slab): the fused tiled kernel against the operator sequence of the reference semantics
(UpdateMolecules -> LJ_IdealGas on a half list of molecules over ghost molecules -> ContributeMoleculeForceToAtoms),
which the small-size tests pin to the oracle."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RC, SKIN, CAP = 2.5, 0.1, 0.7


@pytest.fixture(scope="module")
def api():
    from mrmd_b200 import api as a

    assert a.L().mrmd_b200_device_count() > 0
    return a


def jittered_lattice(n_side, jitter, seed, spacing=1.25):
    rng = np.random.default_rng(seed)
    g = np.stack(np.meshgrid(*[np.arange(n_side)] * 3, indexing="ij"), axis=-1).reshape(-1, 3)
    box = np.full(3, n_side * spacing)
    pos = np.mod((g + 0.5) * spacing + (rng.random(g.shape) - 0.5) * jitter, box)
    vel = rng.random(g.shape) - 0.5
    return pos, vel - vel.mean(axis=0), box


def sorted_rows(keys, counts, width):
    """rows of a neighbour table as sorted int32 key arrays padded with INT32_MAX"""
    out = np.full((len(counts), width), np.iinfo(np.int32).max, dtype=np.int32)
    mask = np.arange(keys.shape[1])[None, :] < counts[:, None]
    w = min(width, keys.shape[1])
    out[:, :w] = np.where(mask[:, :w], keys[:, :w], np.iinfo(np.int32).max)
    assert not mask[:, w:].any()
    out.sort(axis=1)
    return out


def test_lj_1m_step_vs_oracle(api, oracle):
    pos, vel, box = jittered_lattice(100, 0.6, 3)
    n = len(pos)
    cutoff = RC + SKIN
    sub = api.Subdomain([0, 0, 0], box, cutoff)
    atoms = api.Atoms.from_arrays(pos, vel)
    api.GhostLayer().exchangeRealAtoms(atoms, sub)
    atoms.permute(api.LinkedCellList(0, n, [cutoff, cutoff, 0.25 * cutoff], sub.minCorner, sub.maxCorner))
    sorted_pos = atoms.getPos()[:n]
    vl = api.FullVerletList()
    vl.build_periodic(atoms, sub, cutoff, 1.0, 60)
    lj = api.LennardJones(RC, 1.0, 1.0, CAP)
    atoms.setForce(0.0)
    lj.apply(atoms, vl)
    f = atoms.getForce()[:n]
    e, v, p = lj._get()

    L = oracle.lib()
    osub = oracle.subdomain([0, 0, 0], box, cutoff)
    oa = np.zeros(int(1.2 * n) + 1024, dtype=oracle.ATOM)
    oa["pos"][:n], oa["mass"][:n] = sorted_pos, 1.0
    corr = np.zeros(len(oa), dtype=np.int64)
    ng = L.or_ghost_create_xyz(oa.ctypes.data, n, len(oa), C.byref(osub), corr.ctypes.data)
    assert ng > 0
    gmin, gmax = np.array(osub.minGhostCorner), np.array(osub.maxGhostCorner)

    # stored neighbour sets: the reference's full list over local + ghost atoms, partners mapped to (real atom, image)
    fc, fn = oracle.verlet_build(oa, 13, n + ng, 0, n, cutoff, 1.0, gmin, gmax, half=False, width=80)
    okeys = np.empty((n, fn.shape[1]), dtype=np.int32)
    for lo in range(0, n, 65536):  # row blocks keep the temporaries below 1 GB
        hi = min(n, lo + 65536)
        nb = np.where(fn[lo:hi] >= 0, fn[lo:hi], 0).astype(np.int64)
        real = np.where(nb < n, nb, corr[nb])
        shift = np.rint((oa["pos"][nb] - oa["pos"][real]) / box).astype(np.int64)
        okeys[lo:hi] = real * 27 + (shift[..., 0] + 1) + 3 * (shift[..., 1] + 1) + 9 * (shift[..., 2] + 1)
    del nb, real, shift
    gc, gp, gcode = vl.to_host_periodic(atoms)
    assert np.array_equal(gc[:n], fc[:n])
    assert vl.info()["totalPairs"] == int(fc[:n].sum())
    gkeys = (gp.astype(np.int64) * 27 + gcode).astype(np.int32)
    assert np.array_equal(sorted_rows(gkeys, gc[:n], 80), sorted_rows(okeys, fc[:n], 80))  # bit exact, 3.5e7 entries
    del gkeys, okeys, gp, gcode, fn

    # forces, energy, virial, pair count: half list + ghost fold on the oracle side
    oc, on = oracle.verlet_build(oa, 13, n + ng, 0, n, cutoff, 1.0, gmin, gmax, half=True, width=64)
    table = oracle.lj_table(CAP, RC, 1.0, 1.0)
    ev = np.zeros(2)
    pairs = L.or_lj_apply(oa.ctypes.data, n, oc.ctypes.data, on.ctypes.data, on.shape[1], C.addressof(table), RC * RC, 1,
                          None, ev.ctypes.data)
    L.or_ghost_fold_force(oa.ctypes.data, n, ng, corr.ctypes.data)
    assert p == pairs
    assert np.abs(f - oa["force"][:n]).max() <= 1e-10 * np.abs(oa["force"][:n]).max()
    # and atom by atom against the atom's own force (floor: 1e-3 of the largest force in the system)
    norm, err = np.linalg.norm(oa["force"][:n], axis=1), np.linalg.norm(f - oa["force"][:n], axis=1)
    assert np.all(err <= 1e-10 * np.maximum(norm, 1e-3 * norm.max()))
    assert abs(e - ev[0]) <= 1e-12 * abs(ev[0]) and abs(v - ev[1]) <= 1e-12 * abs(ev[1])
    assert np.abs(f.sum(axis=0)).max() <= 1e-9 * np.abs(f).sum()  # Newton's third law over the periodic box


def test_lj_1m_nve_invariants(api):
    pos, vel, box = jittered_lattice(100, 0.3, 5)
    n = len(pos)
    sub = api.Subdomain([0, 0, 0], box, RC + SKIN)
    atoms = api.Atoms.from_arrays(pos, vel)
    # melt the lattice under the Langevin thermostat first, then switch the thermostat off
    api.MolecularDynamics(atoms, sub, langevin=True, zeta=20.0, temperature=1.5, seed=7, cellSort=True, fullList=2).run(300)
    md = api.MolecularDynamics(atoms, sub, langevin=False, cellSort=True, fullList=2)
    p0 = atoms.getVel()[:n].sum(axis=0)

    def total_energy():
        st = md.run(1)
        return st["energy"] + api.analysis.getKineticEnergy(atoms)

    e0 = total_energy()
    st = md.run(198)
    e1 = total_energy()
    assert st["rebuilds"] >= 10
    p1 = atoms.getVel()[:n].sum(axis=0)
    assert np.abs(p1 - p0).max() <= 1e-9 * np.abs(atoms.getVel()[:n]).sum()  # momentum conservation
    # velocity Verlet with dt = 0.002 on the capped, unshifted potential (every pair crossing r_c moves the truncated
    # Hamiltonian by 0.016 epsilon, in the reference too): the drift stays two orders below the kinetic energy
    ekin = api.analysis.getKineticEnergy(atoms)
    assert abs(e1 - e0) <= 1e-2 * ekin, (e0, e1, ekin)


def test_adress_8m_tiled_vs_operator_sequence(api):
    """configs[2] size: 200^3 atoms, Slab(AT 0.2 L, HY 0.1 L): one force evaluation of the fused tiled kernel against
    the generic operators on the same cell-sorted atoms"""
    pos, vel, box = jittered_lattice(200, 0.6, 9)
    n = len(pos)
    cutoff = RC + SKIN
    sub = api.Subdomain([0, 0, 0], box, cutoff)
    w = api.Slab(0.5 * box, 0.2 * box[0], 0.1 * box[0], 1)
    atoms = api.Atoms.from_arrays(pos, vel, mass=1.0, relativeMass=1.0, capacity=int(1.08 * n))
    api.GhostLayer().exchangeRealAtoms(atoms, sub)
    atoms.permute(api.LinkedCellList(0, n, [cutoff, cutoff, 0.25 * cutoff], sub.minCorner, sub.maxCorner))
    vl = api.FullVerletList()
    vl.build_periodic(atoms, sub, cutoff, 1.0, 60)
    tiled = api.LJ_IdealGas(CAP, RC, 1.0, 1.0, True)
    atoms.setForce(0.0)
    e_tiled = tiled.run_periodic(atoms, vl, w)
    pairs_tiled = tiled.lastNumPairs
    f_tiled = atoms.getForce()[:n]
    del vl

    mols = api.createMoleculeForEachAtom(atoms)
    api.UpdateMolecules.update(mols, atoms, w)
    ghost = api.MultiResGhostLayer()
    ghost.createGhostAtoms(mols, atoms, sub)
    api.UpdateMolecules.update(mols, atoms, w)
    hl = api.HalfVerletList()
    hl.build(mols, 0, mols.numLocalMolecules, cutoff, 1.0, list(sub.minGhostCorner), list(sub.maxGhostCorner), 60)
    generic = api.LJ_IdealGas(CAP, RC, 1.0, 1.0, True)
    atoms.setForce(0.0)
    mols.setForce(0.0)
    e_generic = generic.run(mols, hl, atoms)
    api.ContributeMoleculeForceToAtoms.update(mols, atoms)
    ghost.contributeBackGhostToReal(atoms)
    f_generic = atoms.getForce()[:n]
    assert pairs_tiled == generic.lastNumPairs > 0
    assert abs(e_tiled - e_generic) <= 1e-11 * abs(e_generic)
    assert np.abs(f_tiled - f_generic).max() <= 1e-10 * np.abs(f_generic).max()


def test_adress_8m_tiled_vs_oracle(api, oracle):
    """configs[2] size against the ORACLE (not against another CUDA path): 200^3 atoms, Slab(AT 0.2 L, HY 0.1 L), one
    evaluation of UpdateMolecules + LJ_IdealGas::run + ContributeMoleculeForceToAtoms.  The oracle runs the reference's
    step over ghost molecules (half list, forces folded back) on the atoms in the GPU's sorted order.  Forces are compared
    atom by atom against the atom's own force (with a floor of 1e-3 of the largest), not only against the largest."""
    import ctypes as C

    pos, vel, box = jittered_lattice(200, 0.6, 9)
    n = len(pos)
    cutoff = RC + SKIN
    sub = api.Subdomain([0, 0, 0], box, cutoff)
    centre = [0.5 * float(b) for b in box]
    w = api.Slab(0.5 * box, 0.2 * box[0], 0.1 * box[0], 1)
    ow = oracle.make_weight(oracle.WEIGHT_SLAB, centre, 0.2 * float(box[0]), 0.1 * float(box[0]), 1)
    atoms = api.Atoms.from_arrays(pos, vel, mass=1.0, relativeMass=1.0, capacity=int(1.08 * n))
    del pos, vel
    api.GhostLayer().exchangeRealAtoms(atoms, sub)
    atoms.permute(api.LinkedCellList(0, n, [cutoff, cutoff, 0.25 * cutoff], sub.minCorner, sub.maxCorner))
    spos = atoms.getPos()[:n]
    vl = api.FullVerletList()
    vl.build_periodic(atoms, sub, cutoff, 1.0, 60)
    tiled = api.LJ_IdealGas(CAP, RC, 1.0, 1.0, True)
    atoms.setForce(0.0)
    e_tiled = tiled.run_periodic(atoms, vl, w)
    pairs_tiled = tiled.lastNumPairs
    total_pairs = vl.info()["totalPairs"]
    f = atoms.getForce()[:n]
    del vl, atoms

    L = oracle.lib()
    osub = oracle.subdomain([0, 0, 0], [float(b) for b in box], cutoff)
    cap_rows = int(1.10 * n)
    oa = np.zeros(cap_rows, dtype=oracle.ATOM)
    oa["pos"][:n], oa["mass"][:n], oa["relMass"][:n] = spos, 1.0, 1.0
    om = np.zeros(cap_rows, dtype=oracle.MOLECULE)
    om["atomsOffset"][:n], om["numAtoms"][:n] = np.arange(n), 1
    L.or_update_molecules(om.ctypes.data, n, oa.ctypes.data, C.byref(ow))
    corr = np.zeros(cap_rows, dtype=np.int64)
    out = np.zeros(2, dtype=np.int64)
    assert L.or_mr_ghost_create_xyz(om.ctypes.data, n, cap_rows, oa.ctypes.data, n, cap_rows, C.byref(osub),
                                    corr.ctypes.data, out.ctypes.data) == 0
    mg, ag = int(out[0]), int(out[1])
    oc, on = oracle.verlet_build(om, 13, n + mg, 0, n, cutoff, 1.0, np.array(osub.minGhostCorner),
                                 np.array(osub.maxGhostCorner), half=True, width=40)
    assert total_pairs == 2 * int(oc[:n].sum())
    one = np.ones(1)
    arrs = [np.ascontiguousarray(x) for x in (CAP * one, RC * one, one, one)]
    oh = L.or_adress_create(*[x.ctypes.data for x in arrs], 1, 1)
    L.or_update_molecules(om.ctypes.data, n + mg, oa.ctypes.data, C.byref(ow))
    nact = C.c_int64()
    oe = L.or_adress_run(oh, om.ctypes.data, n, oc.ctypes.data, on.ctypes.data, on.shape[1], oa.ctypes.data, C.byref(nact))
    L.or_contribute_molecule_force(om.ctypes.data, n + mg, oa.ctypes.data)
    L.or_ghost_fold_force(oa.ctypes.data, n, ag, corr.ctypes.data)
    L.or_adress_destroy(oh)
    f_ref = oa["force"][:n]
    assert pairs_tiled == nact.value > 0
    assert abs(e_tiled - oe) <= 1e-11 * abs(oe)
    fmax = np.abs(f_ref).max()
    assert np.abs(f - f_ref).max() <= 1e-10 * fmax
    norm = np.linalg.norm(f_ref, axis=1)
    err = np.linalg.norm(f - f_ref, axis=1)
    assert np.all(err <= 1e-10 * np.maximum(norm, 1e-3 * fmax))
