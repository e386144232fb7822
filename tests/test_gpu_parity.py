"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same inputs and
against the reference's golden numbers.  Integer / index work is bit exact; floating point within the
tolerance written in each test (north star: 1e-10 relative for single-step FP64 forces)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FORCE_RTOL = 1e-10   # relative to the largest force component in the system
SCALAR_RTOL = 1e-12  # energy / virial


@pytest.fixture(scope="module")
def api():
    from mrmd_b200 import api as a

    assert a.L().mrmd_b200_device_count() > 0, "no CUDA device: the gpu tests need a B200"
    return a


def float_eq(a, b):
    a32, b32 = np.float32(a), np.float32(b)
    return abs(float(a32) - float(b32)) <= 4 * np.spacing(max(abs(a32), abs(b32), np.float32(1e-30)))


def oracle_atoms(orc, pos, vel=None, capacity=None, mass=1.0):
    n = len(pos)
    a = np.zeros(capacity or n, dtype=orc.ATOM)
    a["pos"][:n] = pos
    if vel is not None:
        a["vel"][:n] = vel
    a["mass"][:n] = mass
    a["relMass"][:n] = 1.0
    return a


def lj_system(n_side, spacing, jitter, seed):
    rng = np.random.default_rng(seed)
    g = (np.stack(np.meshgrid(*[np.arange(n_side)] * 3, indexing="ij"), axis=-1).reshape(-1, 3) + 0.5) * spacing
    pos = g + (rng.random(g.shape) - 0.5) * jitter
    box = n_side * spacing
    pos = np.mod(pos, box)
    vel = (rng.random(g.shape) - 0.5)
    return pos, vel, box


def sorted_rows(counts, neigh, n):
    out = np.full_like(neigh[:n], np.iinfo(np.int32).max)
    mask = np.arange(neigh.shape[1])[None, :] < counts[:n, None]
    out[mask] = neigh[:n][mask]
    out.sort(axis=1)
    return out


def compare_lists(orc, gc, gn, oc, on, n):
    assert np.array_equal(gc[:n], oc[:n])
    w = max(gn.shape[1], on.shape[1])

    def pad(a):
        return np.pad(a, ((0, 0), (0, w - a.shape[1])), constant_values=np.iinfo(np.int32).max)
    assert np.array_equal(pad(sorted_rows(gc, gn, n)), pad(sorted_rows(oc, on, n)))


# ---------------------------------------------------------------------------------------------------
def test_espp_golden_through_cuda(api, oracle, golden_dir):
    """tests/LennardJones/LennardJones.cpp:82-143 run on the device."""
    g = np.load(f"{golden_dir}/espp_positions.npz")
    n = int(g["espp_real"])
    rc, skin = float(g["rc"]), float(g["skin"])
    sub = api.Subdomain([0, 0, 0], g["box"], rc + skin)
    atoms = api.Atoms.from_arrays(g["pos"])
    ghost = api.GhostLayer()
    ghost.createGhostAtoms(atoms, sub)
    assert atoms.numLocalAtoms == n and atoms.numGhostAtoms == int(g["espp_ghost"])
    vl = api.HalfVerletList()
    vl.build(atoms, 0, n, rc + skin, 1.0, sub.minGhostCorner, sub.maxGhostCorner, 60)
    gc, gn = vl.to_host()
    assert int(gc.sum()) == int(g["espp_neighbors"]) == vl.info()["totalPairs"]
    lj = api.LennardJones(rc, 1.0, 1.0)
    lj.apply(atoms, vl)
    assert float_eq(lj.getEnergy(), float(g["espp_initial_energy"]))

    # same path in the oracle: ghosts, list (sorted pair sets bit exact), forces
    L = oracle.lib()
    osub = oracle.subdomain([0, 0, 0], g["box"], rc + skin)
    oa = oracle_atoms(oracle, g["pos"], capacity=2 * n)
    corr = np.zeros(2 * n, dtype=np.int64)
    ng = L.or_ghost_create_xyz(oa.ctypes.data, n, 2 * n, C.byref(osub), corr.ctypes.data)
    assert np.array_equal(atoms.getPos()[:n + ng], oa["pos"][:n + ng])  # ghost order and shifted positions
    assert np.array_equal(ghost.correspondingRealAtom(n + ng), corr[:n + ng])
    oc, on = oracle.verlet_build(oa, 13, n + ng, 0, n, rc + skin, 1.0, np.array(osub.minGhostCorner),
                                 np.array(osub.maxGhostCorner), half=True, width=80)
    compare_lists(oracle, gc, gn, oc, on, n)
    table = oracle.lj_table(0.0, rc, 1.0, 1.0)
    ev = np.zeros(2)
    pairs = L.or_lj_apply(oa.ctypes.data, n, oc.ctypes.data, on.ctypes.data, on.shape[1], C.addressof(table), rc * rc, 1,
                          None, ev.ctypes.data)
    f = atoms.getForce()[:n + ng]
    scale = np.abs(oa["force"]).max()
    assert np.abs(f - oa["force"][:n + ng]).max() <= FORCE_RTOL * scale
    e, v, p = lj._get()
    assert abs(e - ev[0]) <= SCALAR_RTOL * abs(ev[0]) and abs(v - ev[1]) <= SCALAR_RTOL * abs(ev[1]) and p == pairs
    # fold back and compare again; full list gives the same folded forces with no scatter
    ghost.contributeBackGhostToReal(atoms)
    L.or_ghost_fold_force(oa.ctypes.data, n, ng, corr.ctypes.data)
    folded = atoms.getForce()
    assert np.abs(folded[:n] - oa["force"][:n]).max() <= FORCE_RTOL * scale and np.all(folded[n:n + ng] == 0)
    fl = api.FullVerletList()
    fl.build(atoms, 0, n, rc + skin, 1.0, sub.minGhostCorner, sub.maxGhostCorner, 100)
    fc, fn = fl.to_host()
    ofc, ofn = oracle.verlet_build(oa, 13, n + ng, 0, n, rc + skin, 1.0, np.array(osub.minGhostCorner),
                                   np.array(osub.maxGhostCorner), half=False, width=120)
    compare_lists(oracle, fc, fn, ofc, ofn, n)
    atoms.setForce(0.0)
    lj.apply(atoms, fl)
    full = atoms.getForce()
    assert np.abs(full[:n] - oa["force"][:n]).max() <= FORCE_RTOL * scale and np.all(full[n:] == 0)
    e2, v2, p2 = lj._get()
    assert abs(e2 - ev[0]) <= SCALAR_RTOL * abs(ev[0]) and abs(v2 - ev[1]) <= SCALAR_RTOL * abs(ev[1]) and p2 == pairs


@pytest.mark.parametrize("n_side,jitter,half", [(12, 0.6, True), (12, 0.6, False), (20, 0.0, True), (20, 0.0, False),
                                                 (9, 1.2, True)])
def test_verlet_and_force_vs_oracle(api, oracle, n_side, jitter, half):
    """Random and perfect-lattice systems (the lattice puts the 4th shell exactly on rc = 2.5: the strict
    `distSqr > rcSqr` skip and the coordinate tie-break of the half criterion decide)."""
    pos, vel, box = lj_system(n_side, 1.25, jitter, 11 + n_side)
    n = len(pos)
    rc, skin, cap = 2.5, 0.1, 0.7
    sub = api.Subdomain([0, 0, 0], [box] * 3, rc + skin)
    atoms = api.Atoms.from_arrays(pos, vel)
    ghost = api.GhostLayer()
    ghost.exchangeRealAtoms(atoms, sub)
    ghost.createGhostAtoms(atoms, sub)
    vl = api.HalfVerletList() if half else api.FullVerletList()
    vl.build(atoms, 0, n, rc + skin, 1.0, sub.minGhostCorner, sub.maxGhostCorner, 40 if half else 60)
    lj = api.LennardJones(rc, 1.0, 1.0, cap)
    atoms.setForce(0.0)
    lj.apply(atoms, vl)
    ghost.contributeBackGhostToReal(atoms)

    L = oracle.lib()
    osub = oracle.subdomain([0, 0, 0], [box] * 3, rc + skin)
    oa = oracle_atoms(oracle, pos, vel, capacity=4 * n)
    L.or_periodic_map(oa.ctypes.data, n, C.byref(osub))
    corr = np.zeros(4 * n, dtype=np.int64)
    ng = L.or_ghost_create_xyz(oa.ctypes.data, n, 4 * n, C.byref(osub), corr.ctypes.data)
    assert atoms.numGhostAtoms == ng
    assert np.array_equal(atoms.getPos()[:n + ng], oa["pos"][:n + ng])
    oc, on = oracle.verlet_build(oa, 13, n + ng, 0, n, rc + skin, 1.0, np.array(osub.minGhostCorner),
                                 np.array(osub.maxGhostCorner), half=half, width=64)
    gc, gn = vl.to_host()
    compare_lists(oracle, gc, gn, oc, on, n)
    # forces: the oracle always walks a half list (reference semantics)
    hc, hn = (oc, on) if half else oracle.verlet_build(oa, 13, n + ng, 0, n, rc + skin, 1.0,
                                                       np.array(osub.minGhostCorner), np.array(osub.maxGhostCorner),
                                                       half=True, width=64)
    table = oracle.lj_table(cap, rc, 1.0, 1.0)
    ev = np.zeros(2)
    pairs = L.or_lj_apply(oa.ctypes.data, n, hc.ctypes.data, hn.ctypes.data, hn.shape[1], C.addressof(table), rc * rc, 1,
                          None, ev.ctypes.data)
    L.or_ghost_fold_force(oa.ctypes.data, n, ng, corr.ctypes.data)
    f = atoms.getForce()[:n]
    scale = max(np.abs(oa["force"][:n]).max(), 1e-300)
    if jitter == 0.0:
        # perfect lattice: forces cancel to rounding; compare against the size of a single pair force
        scale = 1.0
    assert np.abs(f - oa["force"][:n]).max() <= FORCE_RTOL * scale
    e, v, p = lj._get()
    assert p == pairs
    assert abs(e - ev[0]) <= SCALAR_RTOL * abs(ev[0]) and abs(v - ev[1]) <= SCALAR_RTOL * abs(ev[1])


def test_apply_if_predicates(api, oracle):
    """examples/04: bare LJ where either atom is in the inner slab, capped LJ where both are in the outer slab."""
    pos, vel, box = lj_system(14, 1.25, 0.8, 5)
    n = len(pos)
    rc, skin = 2.5, 0.1
    sub = api.Subdomain([0, 0, 0], [box] * 3, [0.0, rc + skin, rc + skin])  # no x ghosts, examples/04:96-98
    atoms = api.Atoms.from_arrays(pos, vel)
    ghost = api.GhostLayer()
    ghost.createGhostAtoms(atoms, sub)
    vl = api.HalfVerletList()
    vl.build(atoms, 0, n, rc + skin, 1.0, sub.minGhostCorner, sub.maxGhostCorner, 60)
    center = sub.getCenter()
    inner = api.IsInSymmetricSlab(center, 0.0, 3.0)
    outer = api.IsInSymmetricSlab(center, 3.0, 6.0)
    lj = api.LennardJones(rc, 1.0, 1.0, 0.0)
    ljcap = api.LennardJones(rc, 1.0, 1.0, 0.7)
    atoms.setForce(0.0)
    lj.apply_if(atoms, vl, inner.either())
    ljcap.apply_if(atoms, vl, outer.both())

    L = oracle.lib()
    osub = oracle.subdomain([0, 0, 0], [box] * 3, [0.0, rc + skin, rc + skin])
    oa = oracle_atoms(oracle, pos, vel, capacity=3 * n)
    corr = np.zeros(3 * n, dtype=np.int64)
    ng = L.or_ghost_create_xyz(oa.ctypes.data, n, 3 * n, C.byref(osub), corr.ctypes.data)
    assert ng == atoms.numGhostAtoms
    oc, on = oracle.verlet_build(oa, 13, n + ng, 0, n, rc + skin, 1.0, np.array(osub.minGhostCorner),
                                 np.array(osub.maxGhostCorner), half=True, width=64)
    t0, t1 = oracle.lj_table(0.0, rc, 1.0, 1.0), oracle.lj_table(0.7, rc, 1.0, 1.0)
    p0 = oracle.make_pred(oracle.PRED_SLAB_EITHER, 0, center[0], 0.0, 3.0)
    p1 = oracle.make_pred(oracle.PRED_SLAB_BOTH, 0, center[0], 3.0, 6.0)
    ev0, ev1 = np.zeros(2), np.zeros(2)
    n0 = L.or_lj_apply(oa.ctypes.data, n, oc.ctypes.data, on.ctypes.data, on.shape[1], C.addressof(t0), rc * rc, 1,
                       C.byref(p0), ev0.ctypes.data)
    n1 = L.or_lj_apply(oa.ctypes.data, n, oc.ctypes.data, on.ctypes.data, on.shape[1], C.addressof(t1), rc * rc, 1,
                       C.byref(p1), ev1.ctypes.data)
    assert 0 < n0 and 0 < n1 and lj.getNumPairs() == n0 and ljcap.getNumPairs() == n1
    f = atoms.getForce()[:n + ng]
    assert np.abs(f - oa["force"][:n + ng]).max() <= FORCE_RTOL * np.abs(oa["force"]).max()
    assert abs(lj.getEnergy() - ev0[0]) <= SCALAR_RTOL * abs(ev0[0])
    assert abs(ljcap.getEnergy() - ev1[0]) <= SCALAR_RTOL * abs(ev1[0])


def test_lj_potential_kat(api):
    """mrmd/action/LennardJones.test.cpp:46-81 on the device."""
    eps, sigma = 2.0, 3.01
    rc, cap = 2.5 * sigma, 0.1
    lj = api.LennardJones(rc, sigma, eps, cap, isShifted=True)
    x = cap + 0.1 + np.arange(100) * 0.1
    ff, e = lj.computeForceAndEnergy(x * x)
    cutoff_pot = 4 * eps * ((sigma / rc) ** 12 - (sigma / rc) ** 6)
    for i in range(100):
        pot = 4 * eps * ((sigma / x[i]) ** 12 - (sigma / x[i]) ** 6) - cutoff_pot
        force = 4 * eps * (-12 * (sigma / x[i]) ** 12 + 6 * (sigma / x[i]) ** 6) * x[i] / (x[i] * x[i])
        assert float_eq(e[i], pot) and float_eq(-x[i] * ff[i], force)
    ljc = api.LennardJones(2.5, 1.0, 1.0, 0.7)
    ff, e = ljc.computeForceAndEnergy(np.array([0.49 * (1 + 1e-12), 0.49 * (1 - 1e-12), 0.25]))
    assert abs(ff[0] - ff[1]) < 1e-6 * abs(ff[0]) and abs(ff[2] * 0.5 - ff[0] * 0.7) < 1e-6 * abs(ff[0])


def test_cell_sort_bit_exact(api, oracle):
    """LinkedCellList + permute (tests/NVT/NVT.cpp:136-144): cell ids bit equal, stable in-cell order."""
    rng = np.random.default_rng(17)
    n = 70000
    box = 33.0
    pos = rng.random((n, 3)) * box
    pos[:50] = np.floor(pos[:50] / 2.6) * (box / 12)  # points exactly on cell faces
    vel = rng.random((n, 3))
    atoms = api.Atoms.from_arrays(pos, vel, mass=np.arange(n) + 1.0, type=np.arange(n) % 3)
    atoms.set("charge", rng.random(n))
    atoms.set("relativeMass", rng.random(n))
    atoms.set("force", rng.random((n, 3)))
    before = {k: atoms.get(k) for k in api.ATOM_FIELDS}
    import torch  # device memory for the optional cell-id output

    cid = torch.zeros(n, dtype=torch.int32, device="cuda")
    delta, lo, hi = np.full(3, 2.6), np.zeros(3), np.full(3, box)
    from mrmd_b200._lib import check

    check(api.L().mrmd_b200_atoms_cell_sort(atoms.h, 0, n, delta.ctypes.data, lo.ctypes.data, hi.ctypes.data,
                                            cid.data_ptr(), None))
    api.sync()
    oa = oracle_atoms(oracle, pos)
    ocid = np.zeros(n, dtype=np.int32)
    dims = np.zeros(3, dtype=np.int32)
    nc = oracle.lib().or_cell_ids(oa.ctypes.data, 13, 0, n, delta.ctypes.data, lo.ctypes.data, hi.ctypes.data,
                                  ocid.ctypes.data, dims.ctypes.data)
    assert np.array_equal(cid.cpu().numpy(), ocid)
    perm = np.zeros(n, dtype=np.int64)
    off = np.zeros(nc + 1, dtype=np.int64)
    oracle.lib().or_cell_perm(ocid.ctypes.data, 0, n, nc, perm.ctypes.data, off.ctypes.data)
    for k in api.ATOM_FIELDS:
        assert np.array_equal(atoms.get(k), before[k][perm]), k
    # partial range: only [begin, end) moves
    atoms2 = api.Atoms.from_arrays(pos, vel)
    atoms2.permute(api.LinkedCellList(1000, 40000, delta, lo, hi))
    p2 = atoms2.getPos()
    assert np.array_equal(p2[:1000], pos[:1000]) and np.array_equal(p2[40000:], pos[40000:])
    sub_ids = ocid[1000:40000]
    order = np.argsort(sub_ids, kind="stable")
    assert np.array_equal(p2[1000:40000], pos[1000:40000][order])


def test_ghost_layer_kats(api):
    """GhostExchange.test.cpp:95-153, UpdateGhostAtoms.test.cpp:69-109, AccumulateForce.test.cpp:27-49,
    PeriodicMapping.test.cpp:46-87 on the device."""
    sub = api.Subdomain([0, 0, 0], [3, 3, 3], 0.7)
    grid = np.array([(x + .5, y + .5, z + .5) for x in range(3) for y in range(3) for z in range(3)])
    for axis in range(3):
        atoms = api.Atoms(200)
        atoms.set("pos", grid)
        atoms.numLocalAtoms, atoms.numGhostAtoms = 27, 0
        atoms.resize(27)
        ghost = api.GhostLayer()
        ghost.resetCorrespondingRealAtoms(atoms)
        ghost.createGhostAtoms(atoms, sub, axis=axis)
        assert atoms.numGhostAtoms == 18 and atoms.size() == 45
        corr = ghost.correspondingRealAtom(45)
        assert np.all(corr[:27] == -1) and np.all((corr[27:] >= 0) & (corr[27:] < 27))
        p = atoms.getPos()
        assert np.all(p[27:36, axis] > 3.0) and np.all(p[36:45, axis] < 0.0)
    atoms = api.Atoms.from_arrays(grid)
    ghost = api.GhostLayer()
    ghost.createGhostAtoms(atoms, sub)
    assert atoms.numGhostAtoms == 98
    corr = ghost.correspondingRealAtom(125)
    assert np.all(corr[:27] == -1) and np.all((corr[27:] >= 0) & (corr[27:] < 27))

    cases = [((0.4, 0.5, 0.6), (0.4, 0.5, 0.6)), ((1.1, 0.5, 0.6), (0.1, 0.5, 0.6)), ((-0.1, 0.5, 0.6), (0.9, 0.5, 0.6)),
             ((0.4, 1.1, 0.6), (0.4, 0.1, 0.6)), ((0.4, -0.1, 0.6), (0.4, 0.9, 0.6)), ((0.4, 0.5, 1.1), (0.4, 0.5, 0.1)),
             ((0.4, 0.5, -0.1), (0.4, 0.5, 0.9)), ((1.1, 1.2, 1.3), (0.1, 0.2, 0.3)), ((-0.3, -0.2, -0.1), (0.7, 0.8, 0.9)),
             ((1.0, -1e-18, 0.0), (0.0, 0.0, 0.0))]
    unit = api.Subdomain([0, 0, 0], [1, 1, 1], 0.0)
    a = api.Atoms.from_arrays(np.array([c[0] for c in cases]))
    api.GhostLayer().exchangeRealAtoms(a, unit)
    got = a.getPos()
    for row, (_, want) in zip(got, cases):
        assert all(float_eq(x, y) for x, y in zip(row, want))

    unit = api.Subdomain([0, 0, 0], [1, 1, 1], 0.1)
    for delta, final in [((0.2, 0, 0), (1, 0, 0)), ((-0.2, 0, 0), (-1, 0, 0)), ((0, 0.2, 0), (0, 1, 0)),
                         ((0, -0.2, 0), (0, -1, 0)), ((0, 0, 0.2), (0, 0, 1)), ((0, 0, -0.2), (0, 0, -1)),
                         ((0.2, 0.2, 0.2), (1, 1, 1)), ((-0.2, -0.2, -0.2), (-1, -1, -1))]:
        a = api.Atoms.from_arrays(np.array([[0.5, 0.5, 0.5], 0.5 + np.array(delta)]))
        a.numLocalAtoms, a.numGhostAtoms = 1, 1
        gl = api.GhostLayer()
        gl.setCorrespondingRealAtom([-1, 0])
        gl.updateGhostAtoms(a, unit)
        assert all(float_eq(x, y) for x, y in zip(a.getPos()[1], 0.5 + np.array(final)))

    a = api.Atoms(101)
    a.numLocalAtoms, a.numGhostAtoms = 1, 100
    a.fill("force", 1.0)
    gl = api.GhostLayer()
    gl.setCorrespondingRealAtom(np.concatenate([[-1], np.zeros(100, dtype=np.int64)]))
    gl.contributeBackGhostToReal(a)
    f = a.getForce()
    assert tuple(f[0]) == (101.0, 101.0, 101.0) and np.all(f[1:] == 0.0)


def test_integrators(api, oracle):
    """VelocityVerlet.test.cpp:27-71, VelocityVerletLangevinThermostat.test.cpp:44-121,
    tests/LangevinThermostat/LangevinThermostat.cpp:82-136 and parity with the oracle on 50k atoms."""
    def single():
        a = api.Atoms.from_arrays(np.array([[2.0, 3, 4]]), np.array([[7.0, 5, 3]]), mass=1.5)
        a.set("force", np.array([[9.0, 7, 8]]))
        return a
    a = single()
    d = api.VelocityVerlet.preForceIntegrate(a, 4.0)
    assert all(float_eq(x, y) for x, y in zip(a.getVel()[0], (19, 14.333333, 13.666667)))
    assert all(float_eq(x, y) for x, y in zip(a.getPos()[0], (78, 60.333332, 58.666668)))
    assert tuple(a.getForce()[0]) == (9.0, 7.0, 8.0)
    assert abs(d - np.linalg.norm(a.getPos()[0] - np.array([2.0, 3, 4]))) < 1e-12
    b = single()
    api.VelocityVerlet.postForceIntegrate(b, 4.0)
    assert all(float_eq(x, y) for x, y in zip(b.getVel()[0], (19, 14.333333, 13.666667)))
    assert tuple(b.getPos()[0]) == (2.0, 3.0, 4.0)
    c, e = single(), single()
    lv = api.VelocityVerletLangevinThermostat(0.5, 1.0)
    lv.preForceIntegrate_apply_if(c, 4.0, api.never_pred())
    assert np.allclose(c.getPos(), a.getPos(), rtol=1e-15) and np.allclose(c.getVel(), a.getVel(), rtol=1e-15)
    lv.preForceIntegrate(e, 4.0)
    assert not np.allclose(e.getVel(), a.getVel()) and not np.allclose(e.getPos(), a.getPos())

    rng = np.random.default_rng(3)
    n = 50000
    pos, vel, frc = rng.random((n, 3)) * 30, rng.normal(size=(n, 3)), rng.normal(size=(n, 3)) * 5
    mass = 0.5 + rng.random(n)
    for langevin in (False, True):
        ga = api.Atoms.from_arrays(pos, vel, mass=mass)
        ga.set("force", frc)
        oa = oracle_atoms(oracle, pos, vel, mass=mass)
        oa["force"] = frc
        if langevin:
            lv = api.VelocityVerletLangevinThermostat(20.0, 1.5, seed=1234)
            lv.step = 77
            slab = api.IsInSymmetricSlab([15.0, 0, 0], 0.0, 8.0)
            gd = lv.preForceIntegrate_apply_if(ga, 0.002, slab)
            op = oracle.make_pred(oracle.PRED_SLAB, 0, 15.0, 0.0, 8.0)
            od = oracle.lib().or_langevin_pre(oa.ctypes.data, n, 0.002, 20.0, 1.5, 1234, 77, C.byref(op))
        else:
            gd = api.VelocityVerlet.preForceIntegrate(ga, 0.002)
            od = oracle.lib().or_vv_pre(oa.ctypes.data, n, 0.002)
        # tolerance: device FMA contraction and libm-vs-CUDA log/sincos differ by a few ulp
        assert np.abs(ga.getPos() - oa["pos"]).max() <= 1e-13 * 30
        assert np.abs(ga.getVel() - oa["vel"]).max() <= 1e-12 * np.abs(oa["vel"]).max()
        assert abs(gd - od) <= 1e-12 * od
        api.VelocityVerlet.postForceIntegrate(ga, 0.002)
        oracle.lib().or_vv_post(oa.ctypes.data, n, 0.002)
        assert np.abs(ga.getVel() - oa["vel"]).max() <= 1e-12 * np.abs(oa["vel"]).max()

    # Langevin statistics: 100 000 free atoms, zeta 1e5, 21 steps, T = 1.12 +- 0.01
    n = 100000
    ga = api.Atoms.from_arrays(rng.random((n, 3)) * 10, None, mass=1.0)
    lv = api.VelocityVerletLangevinThermostat(1e5, 1.12)
    for _ in range(21):
        lv.preForceIntegrate(ga, 0.001)
        lv.postForceIntegrate(ga, 0.001)
    ekin = 0.5 * (ga.getVel() ** 2).sum() / n
    assert abs((2.0 / 3.0) * ekin - 1.12) < 0.01


def test_nve_trajectory_vs_oracle(api, oracle, golden_dir):
    """examples/02 loop (config 1, as shipped: lennardJonesNVT_final.gro) for 60 steps on both sides: same
    rebuild steps, same ghost counts, trajectories within a divergence bound."""
    g = np.load(f"{golden_dir}/lj_nvt_final.npz")
    n = len(g["pos"])
    rc, skin, cap, dt = 2.5, 0.1, 0.7, 0.002
    sub = api.Subdomain([0, 0, 0], g["box"], rc + skin)
    atoms = api.Atoms.from_arrays(g["pos"], g["vel"])
    ghost, vl, lj = api.GhostLayer(), api.HalfVerletList(), api.LennardJones(rc, 1.0, 1.0, cap)
    L = oracle.lib()
    osub = oracle.subdomain([0, 0, 0], g["box"], rc + skin)
    oa = oracle_atoms(oracle, g["pos"], g["vel"], capacity=4 * n)
    corr = np.zeros(4 * n, dtype=np.int64)
    table = oracle.lj_table(cap, rc, 1.0, 1.0)
    gmax = omax = np.finfo(np.float64).max
    ng = 0
    oc = on = None
    rebuilds = []
    for step in range(60):
        gmax += api.VelocityVerlet.preForceIntegrate(atoms, dt)
        omax += L.or_vv_pre(oa.ctypes.data, n, dt)
        assert (gmax >= skin * 0.5) == (omax >= skin * 0.5), step
        if gmax >= skin * 0.5:
            gmax = omax = 0.0
            rebuilds.append(step)
            ghost.exchangeRealAtoms(atoms, sub)
            ghost.createGhostAtoms(atoms, sub)
            vl.build(atoms, 0, n, rc + skin, 1.0, sub.minGhostCorner, sub.maxGhostCorner, 60)
            L.or_periodic_map(oa.ctypes.data, n, C.byref(osub))
            ng = L.or_ghost_create_xyz(oa.ctypes.data, n, 4 * n, C.byref(osub), corr.ctypes.data)
            oc, on = oracle.verlet_build(oa, 13, n + ng, 0, n, rc + skin, 1.0, np.array(osub.minGhostCorner),
                                         np.array(osub.maxGhostCorner), half=True, width=60)
            assert atoms.numGhostAtoms == ng, step
        else:
            ghost.updateGhostAtoms(atoms, sub)
            L.or_ghost_update_pos(oa.ctypes.data, n, ng, corr.ctypes.data, C.byref(osub))
        atoms.setForce(0.0)
        oa["force"] = 0.0
        lj.apply(atoms, vl)
        ev = np.zeros(2)
        L.or_lj_apply(oa.ctypes.data, n, oc.ctypes.data, on.ctypes.data, on.shape[1], C.addressof(table), rc * rc, 1,
                      None, ev.ctypes.data)
        ghost.contributeBackGhostToReal(atoms)
        L.or_ghost_fold_force(oa.ctypes.data, n, ng, corr.ctypes.data)
        api.VelocityVerlet.postForceIntegrate(atoms, dt)
        L.or_vv_post(oa.ctypes.data, n, dt)
        assert abs(lj.getEnergy() - ev[0]) <= 1e-9 * abs(ev[0]), step
    assert len(rebuilds) >= 5 and rebuilds[0] == 0
    # chaotic divergence bound after 60 steps from ~1e-16 seeds
    assert np.abs(atoms.getPos()[:n] - oa["pos"][:n]).max() < 1e-9
    assert np.abs(atoms.getVel()[:n] - oa["vel"][:n]).max() < 1e-8


def test_langevin_noise_follows_the_global_atom_id(api):
    """Philox is keyed by (seed, step, global atom id): the same system handed over in another atom order (ids given)
    runs the same thermostatted trajectory atom by atom, through spatial sorts and rebuilds"""
    from mrmd_b200.workloads import lattice_system

    pos, vel, box = lattice_system(12, seed=9)
    n = len(pos)
    sub = api.Subdomain([0, 0, 0], box, 2.6)
    kw = dict(langevin=True, zeta=20.0, temperature=1.5, seed=77, cellSort=True, fullList=2)
    a = api.Atoms.from_arrays(pos, vel, mass=1.0)
    api.MolecularDynamics(a, sub, **kw).run(30)
    perm = np.random.default_rng(3).permutation(n)
    b = api.Atoms.from_arrays(pos[perm], vel[perm], mass=1.0, ids=perm)
    api.MolecularDynamics(b, sub, **kw).run(30)
    ia, ib = a.get("id")[:n], b.get("id")[:n]
    assert np.array_equal(np.sort(ia), np.arange(n)) and np.array_equal(np.sort(ib), np.arange(n))
    pa, pb = a.getPos()[:n][np.argsort(ia)], b.getPos()[:n][np.argsort(ib)]
    va, vb = a.getVel()[:n][np.argsort(ia)], b.getVel()[:n][np.argsort(ib)]
    # the neighbour rows are summed in another order: equal up to rounding that 30 steps do not amplify visibly
    assert np.abs(pa - pb).max() < 1e-10 and np.abs(va - vb).max() < 1e-9


@pytest.mark.parametrize("adress", [False, True])
def test_steps_queued_ahead_of_the_host_change_nothing(api, adress, monkeypatch):
    """mrmd_b200_md_run queues steps ahead of the host with the rebuild criterion evaluated on the device
    (md.cu:runQueued); the trajectory, the rebuild steps and the statistics are those of the step-by-step loop, bit for
    bit"""
    from mrmd_b200.workloads import lattice_system

    pos, vel, box = lattice_system(14, seed=4)
    n = len(pos)
    sub = api.Subdomain([0, 0, 0], box, 2.6)
    extra = {}
    if adress:
        extra = dict(adress=True, weight=api.Slab(0.5 * box, 0.25 * box[0], 0.12 * box[0], 1),
                     thermo=dict(targetDensity=0.512, binWidth=0.5, modulation=2.0, sampleInterval=4, updateInterval=12,
                                 sigma=2.0, range=2.0))
    out = []
    for queued in (True, False):
        if queued:
            monkeypatch.delenv("MRMD_B200_NO_QUEUED_STEPS", raising=False)
        else:
            monkeypatch.setenv("MRMD_B200_NO_QUEUED_STEPS", "1")
        atoms = api.Atoms.from_arrays(pos, vel, mass=1.0)
        md = api.MolecularDynamics(atoms, sub, langevin=True, zeta=20.0, temperature=1.5, seed=5, cellSort=True, fullList=2,
                                   **extra)
        stats = [md.run(k) for k in (1, 17, 30)]  # a first step alone (always rebuilds), then runs of uneven length
        out.append((stats, atoms.getPos()[:n], atoms.getVel()[:n], atoms.get("id")[:n]))
    (sa, pa, va, ia), (sb, pb, vb, ib) = out
    for x, y in zip(sa, sb):
        assert x["rebuilds"] == y["rebuilds"] and x["pairInteractions"] == y["pairInteractions"]
        assert x["energy"] == y["energy"] and x["storedPairs"] == y["storedPairs"] and x["steps"] == y["steps"]
    assert sum(s["rebuilds"] for s in sa) >= 5
    assert np.array_equal(ia, ib) and np.array_equal(pa, pb) and np.array_equal(va, vb)
