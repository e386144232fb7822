"""Pins the CPU oracle against every golden vector / known-answer test the reference holds for the
hot path.  CPU only."""
import ctypes as C
import math

import numpy as np
import pytest


def float_eq(a, b):
    """gtest EXPECT_FLOAT_EQ: within 4 float ULPs."""
    a32, b32 = np.float32(a), np.float32(b)
    return abs(float(a32) - float(b32)) <= 4 * np.spacing(max(abs(a32), abs(b32), np.float32(1e-30)))


def atoms_from(oracle, pos, capacity=None):
    n = len(pos)
    a = np.zeros(capacity or n, dtype=oracle.ATOM)
    a["pos"][:n] = pos
    a["mass"][:n] = 1.0
    return a


# --- tests/LennardJones/LennardJones.cpp:82-143 (primary golden) -------------------------------------
def test_espp_comparison(oracle, golden_dir):
    g = np.load(f"{golden_dir}/espp_positions.npz")
    L = oracle.lib()
    pos = g["pos"]
    n = int(g["espp_real"])
    assert pos.shape[0] == n
    rc, skin = float(g["rc"]), float(g["skin"])
    box = np.ascontiguousarray(g["box"])
    sub = oracle.subdomain([0, 0, 0], box, rc + skin)
    cap = 2 * n
    atoms = atoms_from(oracle, pos, cap)

    assert L.or_count_within_cutoff(atoms.ctypes.data, 13, n, n, rc + skin, box.ctypes.data, 1) == int(g["espp_neighbors"])

    corr = np.zeros(cap, dtype=np.int64)
    ng = L.or_ghost_create_xyz(atoms.ctypes.data, n, cap, C.byref(sub), corr.ctypes.data)
    assert ng == int(g["espp_ghost"])
    assert np.all(corr[:n] == -1) and np.all(corr[n:n + ng] >= 0) and np.all(corr[n:n + ng] < n)
    assert L.or_count_within_cutoff(atoms.ctypes.data, 13, n, n + ng, rc + skin, box.ctypes.data, 0) == int(
        g["nonperiodic_pairs_with_ghosts"])

    counts, neigh = oracle.verlet_build(atoms, 13, n + ng, 0, n, rc + skin, 1.0, np.array(sub.minGhostCorner),
                                        np.array(sub.maxGhostCorner), half=True, width=80)
    assert int(counts.sum()) == int(g["espp_neighbors"])
    # the full list holds every pair twice minus ghost rows: local-local pairs twice, local-ghost once
    cf, nf = oracle.verlet_build(atoms, 13, n + ng, 0, n, rc + skin, 1.0, np.array(sub.minGhostCorner),
                                 np.array(sub.maxGhostCorner), half=False, width=120)
    local_local = int((nf[np.arange(nf.shape[1])[None, :] < cf[:, None]] < n).sum())
    assert local_local % 2 == 0
    assert int(cf.sum()) - local_local // 2 == int(g["nonperiodic_pairs_with_ghosts"])

    table = oracle.lj_table(0.0, rc, 1.0, 1.0)
    ev = np.zeros(2)
    L.or_lj_apply(atoms.ctypes.data, n, counts.ctypes.data, neigh.ctypes.data, neigh.shape[1], C.addressof(table),
                  rc * rc, 1, None, ev.ctypes.data)
    assert float_eq(ev[0], float(g["espp_initial_energy"]))
    assert abs(ev[0] - (-94795.927257)) < 1e-4  # survey-time probe, SURVEY.md section 0.7


# --- mrmd/action/LennardJones.test.cpp:46-81 ----------------------------------------------------------
def test_lj_explicit_comparison(oracle):
    eps, sigma = 2.0, 3.01
    rc, cap = 2.5 * sigma, 0.1
    table = oracle.lj_table(cap, rc, sigma, eps, shifted=True)
    cutoff_pot = 4 * eps * ((sigma / rc) ** 12 - (sigma / rc) ** 6)
    ff, e = C.c_double(), C.c_double()
    for idx in range(100):
        x = cap + 0.1 + idx * 0.1
        oracle.lib().or_lj_force_energy(C.addressof(table), 0, x * x, C.byref(ff), C.byref(e))
        pot = 4 * eps * ((sigma / x) ** 12 - (sigma / x) ** 6) - cutoff_pot
        force = 4 * eps * (-12 * (sigma / x) ** 12 + 6 * (sigma / x) ** 6) * x / (x * x)
        assert float_eq(e.value, pot)
        assert float_eq(-x * ff.value, force)


def test_lj_capping_continuity(oracle):
    table = oracle.lj_table(0.7, 2.5, 1.0, 1.0)
    ff0, e0, ff1, e1 = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    r = 0.7
    oracle.lib().or_lj_force_energy(C.addressof(table), 0, r * r * (1 + 1e-12), C.byref(ff0), C.byref(e0))
    oracle.lib().or_lj_force_energy(C.addressof(table), 0, r * r * (1 - 1e-12), C.byref(ff1), C.byref(e1))
    assert abs(ff0.value - ff1.value) < 1e-6 * abs(ff0.value)
    assert abs(e0.value - e1.value) < 1e-6 * abs(e0.value)
    # below the cap the force magnitude |ff * r| is constant
    oracle.lib().or_lj_force_energy(C.addressof(table), 0, 0.25, C.byref(ff1), C.byref(e1))
    assert abs(ff1.value * 0.5 - ff0.value * r) < 1e-6 * abs(ff0.value * r)


# --- mrmd/action/VelocityVerlet.test.cpp:27-71 --------------------------------------------------------
def single_atom(oracle):
    a = np.zeros(1, dtype=oracle.ATOM)
    a["pos"][0] = (2, 3, 4)
    a["vel"][0] = (7, 5, 3)
    a["force"][0] = (9, 7, 8)
    a["mass"][0] = 1.5
    a["charge"][0] = 0.5
    return a


def test_velocity_verlet_kat(oracle):
    a = single_atom(oracle)
    disp = oracle.lib().or_vv_pre(a.ctypes.data, 1, 4.0)
    for got, want in zip(a["vel"][0], (19, 14.333333, 13.666667)):
        assert float_eq(got, want)
    for got, want in zip(a["pos"][0], (78, 60.333332, 58.666668)):
        assert float_eq(got, want)
    assert np.all(a["force"][0] == (9, 7, 8))
    assert abs(disp - np.linalg.norm(a["pos"][0] - np.array([2, 3, 4.0]))) < 1e-12
    b = single_atom(oracle)
    oracle.lib().or_vv_post(b.ctypes.data, 1, 4.0)
    for got, want in zip(b["vel"][0], (19, 14.333333, 13.666667)):
        assert float_eq(got, want)
    assert np.all(b["pos"][0] == (2, 3, 4))


# --- VelocityVerletLangevinThermostat.test.cpp:44-121 -------------------------------------------------
def test_langevin_predicate(oracle):
    a, b, c = single_atom(oracle), single_atom(oracle), single_atom(oracle)
    never = oracle.make_pred(oracle.PRED_NEVER)
    d0 = oracle.lib().or_vv_pre(a.ctypes.data, 1, 4.0)
    d1 = oracle.lib().or_langevin_pre(b.ctypes.data, 1, 4.0, 0.5, 1.0, 1234, 0, C.byref(never))
    assert np.allclose(a["pos"], b["pos"], rtol=1e-15) and np.allclose(a["vel"], b["vel"], rtol=1e-15)
    assert abs(d0 - d1) < 1e-12
    oracle.lib().or_langevin_pre(c.ctypes.data, 1, 4.0, 0.5, 1.0, 1234, 0, None)
    assert not np.allclose(a["vel"], c["vel"]) and not np.allclose(a["pos"], c["pos"])


def test_philox_known_answers(oracle):
    """Random123 kat_vectors for philox4x32-10."""
    cases = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in cases:
        c = np.array(ctr, dtype=np.uint32)
        k = np.array(key, dtype=np.uint32)
        out = np.zeros(4, dtype=np.uint32)
        oracle.lib().or_philox4x32(c.ctypes.data, k.ctypes.data, out.ctypes.data)
        assert tuple(int(v) for v in out) == want


# --- tests/LangevinThermostat/LangevinThermostat.cpp:82-136 -------------------------------------------
def test_langevin_statistics(oracle):
    n = 100000
    rng = np.random.default_rng(7)
    a = np.zeros(n, dtype=oracle.ATOM)
    a["pos"] = rng.random((n, 3)) * 10
    a["mass"] = 1.0
    target = 1.12
    for step in range(21):
        oracle.lib().or_langevin_pre(a.ctypes.data, n, 0.001, 1e5, target, 1234, step, None)
        oracle.lib().or_vv_post(a.ctypes.data, n, 0.001)
    ekin = 0.5 * (a["mass"][:, None] * a["vel"] ** 2).sum() / n
    assert abs((2.0 / 3.0) * ekin - target) < 0.01
    normals = np.zeros((20000, 4))
    for i in range(20000):
        oracle.lib().or_philox_normals(1234, 5, i, normals[i].ctypes.data)
    assert abs(normals.mean()) < 0.02 and abs(normals.std() - 1.0) < 0.02


# --- communication tests -----------------------------------------------------------------------------
def grid27(oracle, capacity=200):
    pos = np.array([(x + 0.5, y + 0.5, z + 0.5) for x in range(3) for y in range(3) for z in range(3)], dtype=float)
    return atoms_from(oracle, pos, capacity)


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_ghost_exchange_axis(oracle, axis):
    sub = oracle.subdomain([0, 0, 0], [3, 3, 3], 0.7)
    a = grid27(oracle)
    corr = np.full(200, -1, dtype=np.int64)
    ng = oracle.lib().or_ghost_create_axis(a.ctypes.data, 27, 0, 200, C.byref(sub), axis, corr.ctypes.data)
    assert ng == 18  # GhostExchange.test.cpp:102
    assert np.all(corr[:27] == -1) and np.all((corr[27:45] >= 0) & (corr[27:45] < 27))
    # low-side atoms (+L) are appended first
    assert np.all(a["pos"][27:36, axis] > 3.0) and np.all(a["pos"][36:45, axis] < 0.0)


def test_ghost_exchange_xyz_and_pairs(oracle):
    L = oracle.lib()
    sub = oracle.subdomain([0, 0, 0], [3, 3, 3], 0.7)
    a = grid27(oracle)
    box = np.array([3.0, 3.0, 3.0])
    assert L.or_count_within_cutoff(a.ctypes.data, 13, 27, 27, 1.1, box.ctypes.data, 0) == 12 * 3 + 9 * 2
    assert L.or_count_within_cutoff(a.ctypes.data, 13, 27, 27, 1.1, box.ctypes.data, 1) == 27 * 6 // 2
    corr = np.zeros(200, dtype=np.int64)
    ng = L.or_ghost_create_xyz(a.ctypes.data, 27, 200, C.byref(sub), corr.ctypes.data)
    assert ng == 98  # GhostExchange.test.cpp:124
    assert np.all(corr[:27] == -1) and np.all((corr[27:125] >= 0) & (corr[27:125] < 27))
    assert L.or_count_within_cutoff(a.ctypes.data, 13, 27, 125, 1.1, box.ctypes.data, 0) == 108


@pytest.mark.parametrize("initial,mapped", [
    ((0.4, 0.5, 0.6), (0.4, 0.5, 0.6)), ((1.1, 0.5, 0.6), (0.1, 0.5, 0.6)), ((-0.1, 0.5, 0.6), (0.9, 0.5, 0.6)),
    ((0.4, 1.1, 0.6), (0.4, 0.1, 0.6)), ((0.4, -0.1, 0.6), (0.4, 0.9, 0.6)), ((0.4, 0.5, 1.1), (0.4, 0.5, 0.1)),
    ((0.4, 0.5, -0.1), (0.4, 0.5, 0.9)), ((1.1, 1.2, 1.3), (0.1, 0.2, 0.3)), ((-0.3, -0.2, -0.1), (0.7, 0.8, 0.9))])
def test_periodic_mapping(oracle, initial, mapped):
    sub = oracle.subdomain([0, 0, 0], [1, 1, 1], 0.0)
    a = atoms_from(oracle, np.array([initial], dtype=float))
    oracle.lib().or_periodic_map(a.ctypes.data, 1, C.byref(sub))
    for got, want in zip(a["pos"][0], mapped):
        assert float_eq(got, want)


def test_periodic_mapping_edges(oracle):
    sub = oracle.subdomain([0, 0, 0], [1, 1, 1], 0.0)
    a = atoms_from(oracle, np.array([[1.0, -1e-18, 0.0]]))
    oracle.lib().or_periodic_map(a.ctypes.data, 1, C.byref(sub))
    assert tuple(a["pos"][0]) == (0.0, 0.0, 0.0)  # max->min; tiny negative +L rounds to max -> clamped to min


@pytest.mark.parametrize("delta,final", [((0.2, 0, 0), (1, 0, 0)), ((-0.2, 0, 0), (-1, 0, 0)), ((0, 0.2, 0), (0, 1, 0)),
                                         ((0, -0.2, 0), (0, -1, 0)), ((0, 0, 0.2), (0, 0, 1)), ((0, 0, -0.2), (0, 0, -1)),
                                         ((0.2, 0.2, 0.2), (1, 1, 1)), ((-0.2, -0.2, -0.2), (-1, -1, -1))])
def test_update_ghost_atoms(oracle, delta, final):
    sub = oracle.subdomain([0, 0, 0], [1, 1, 1], 0.1)
    a = atoms_from(oracle, np.array([[0.5, 0.5, 0.5], 0.5 + np.array(delta)]))
    corr = np.array([-1, 0], dtype=np.int64)
    oracle.lib().or_ghost_update_pos(a.ctypes.data, 1, 1, corr.ctypes.data, C.byref(sub))
    for got, want in zip(a["pos"][1], 0.5 + np.array(final)):
        assert float_eq(got, want)


def test_accumulate_force(oracle):
    a = np.zeros(101, dtype=oracle.ATOM)
    a["force"] = 1.0
    corr = np.zeros(101, dtype=np.int64)
    corr[0] = -1
    oracle.lib().or_ghost_fold_force(a.ctypes.data, 1, 100, corr.ctypes.data)
    assert tuple(a["force"][0]) == (101.0, 101.0, 101.0)
    assert np.all(a["force"][1:] == 0.0)


# --- multi-resolution ghost layer (GridFixture.hpp:27-86) ----------------------------------------------
def grid_fixture(oracle, atoms_per_molecule=2):
    mols = np.zeros(27 * 10, dtype=oracle.MOLECULE)
    atoms = np.zeros(27 * atoms_per_molecule * 10, dtype=oracle.ATOM)
    idx = 0
    for x in range(3):
        for y in range(3):
            for z in range(3):
                mols["pos"][idx] = (x + 0.5, y + 0.5, z + 0.5)
                mols["atomsOffset"][idx] = idx * atoms_per_molecule
                mols["numAtoms"][idx] = atoms_per_molecule
                for i in range(atoms_per_molecule):
                    atoms["pos"][idx * atoms_per_molecule + i] = (x + 0.5 + 0.1 * i, y + 0.5 + 0.2 * i, z + 0.5 + 0.3 * i)
                idx += 1
    return mols, atoms


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_multires_ghost_axis(oracle, axis):
    sub = oracle.subdomain([0, 0, 0], [3, 3, 3], 0.7)
    mols, atoms = grid_fixture(oracle)
    corr = np.full(len(atoms), -1, dtype=np.int64)
    out = np.zeros(2, dtype=np.int64)
    rc = oracle.lib().or_mr_ghost_create_axis(mols.ctypes.data, 27, 0, len(mols), atoms.ctypes.data, 54, 0, len(atoms),
                                              C.byref(sub), axis, corr.ctypes.data, out.ctypes.data)
    assert rc == 0 and tuple(out) == (18, 36)  # MultiResPeriodicGhostExchange.test.cpp:36-37
    # high-side molecules (-L) are appended first (the opposite of the atom variant)
    assert np.all(mols["pos"][27:36, axis] < 0.0) and np.all(mols["pos"][36:45, axis] > 3.0)
    for m in range(27, 45):
        off, cnt = mols["atomsOffset"][m], mols["numAtoms"][m]
        assert cnt == 2 and 54 <= off < 90
        for a in range(off, off + cnt):
            src = corr[a]
            shift = atoms["pos"][a] - atoms["pos"][src]
            assert abs(abs(shift[axis]) - 3.0) < 1e-12


def test_multires_ghost_xyz(oracle):
    sub = oracle.subdomain([0, 0, 0], [3, 3, 3], 0.7)
    mols, atoms = grid_fixture(oracle)
    corr = np.zeros(len(atoms), dtype=np.int64)
    out = np.zeros(2, dtype=np.int64)
    rc = oracle.lib().or_mr_ghost_create_xyz(mols.ctypes.data, 27, len(mols), atoms.ctypes.data, 54, len(atoms),
                                             C.byref(sub), corr.ctypes.data, out.ctypes.data)
    assert rc == 0 and tuple(out) == (98, 196)  # :60-61
    assert np.all(corr[:54] == -1) and np.all((corr[54:54 + 196] >= 0) & (corr[54:54 + 196] < 54))


def test_multires_real_exchange(oracle):
    """MultiResRealAtomsExchange.test.cpp:28-88: molecule + its atoms are shifted together, no clamp."""
    sub = oracle.subdomain([0, 0, 0], [1, 1, 1], 0.1)
    mols = np.zeros(1, dtype=oracle.MOLECULE)
    atoms = np.zeros(2, dtype=oracle.ATOM)
    mols["pos"][0] = (1.1, -0.2, 0.5)
    mols["atomsOffset"][0] = 0
    mols["numAtoms"][0] = 2
    atoms["pos"][0] = (1.05, -0.25, 0.45)
    atoms["pos"][1] = (1.15, -0.15, 0.55)
    oracle.lib().or_mr_periodic_map(mols.ctypes.data, 1, atoms.ctypes.data, C.byref(sub))
    assert np.allclose(mols["pos"][0], (0.1, 0.8, 0.5), atol=1e-14)
    assert np.allclose(atoms["pos"][0], (0.05, 0.75, 0.45), atol=1e-14)
    assert np.allclose(atoms["pos"][1], (0.15, 0.85, 0.55), atol=1e-14)


# --- AdResS -------------------------------------------------------------------------------------------
def lj_idealgas_fixture(oracle):
    mols = np.zeros(2, dtype=oracle.MOLECULE)
    mols["pos"][0] = (-0.5, 0, 0)
    mols["pos"][1] = (+0.5, 0, 0)
    mols["atomsOffset"] = (0, 2)
    mols["numAtoms"] = (2, 2)
    atoms = np.zeros(4, dtype=oracle.ATOM)
    atoms["pos"] = [(-0.5, -0.5, 0), (-0.5, 0.5, 0), (0.5, -0.5, 0), (0.5, 0.5, 0)]
    atoms["relMass"] = 0.5
    counts, neigh = oracle.verlet_build(mols, 13, 2, 0, 2, 2.0, 1.0, [-1, -1, -1], [1, 1, 1], half=True, width=4)
    assert tuple(counts) == (1, 0) and neigh[0, 0] == 1
    return mols, atoms, counts, neigh


def adress_create(oracle, cap, rc, sigma, eps, shift=True):
    arrs = [np.array([float(v)]) for v in (cap, rc, sigma, eps)]  # keep alive across the call
    return oracle.lib().or_adress_create(*[a.ctypes.data for a in arrs], 1, int(shift))


@pytest.mark.parametrize("lam,scale", [(0.0, 0.0), (0.5, 0.5), (1.0, 1.0)])
def test_lj_idealgas_kat(oracle, lam, scale):
    """mrmd/action/LJ_IdealGas.test.cpp:140-220"""
    mols, atoms, counts, neigh = lj_idealgas_fixture(oracle)
    mols["modLambda"] = lam
    base = 2.0 if lam == 0.0 else 0.0
    atoms["force"] = base
    eps, sigma = 2.0, 0.9
    h = adress_create(oracle, 0.0, 2.5 * sigma, sigma, eps)
    oracle.lib().or_adress_run(h, mols.ctypes.data, 2, counts.ctypes.data, neigh.ctypes.data, neigh.shape[1],
                               atoms.ctypes.data, None)
    oracle.lib().or_adress_destroy(h)
    xf, yf = 0.22156665 * scale, 1.3825009 * scale
    want = np.array([(-xf, +yf, 0), (-xf, -yf, 0), (+xf, +yf, 0), (+xf, -yf, 0)]) + base
    for i in range(4):
        for d in range(3):
            assert float_eq(atoms["force"][i, d], want[i, d]), (i, d, atoms["force"][i, d], want[i, d])


def test_lj_idealgas_drift_and_compensation(oracle):
    """Hybrid pair: drift force -V grad(lambda) on molecules and compensation-energy sampling
    (LJ_IdealGas.cpp:159-220, updateMeanCompensationEnergy :21-50)."""
    mols, atoms, counts, neigh = lj_idealgas_fixture(oracle)
    mols["modLambda"] = (0.5, 0.25)
    mols["lambda"] = (0.5, 0.3)
    mols["gradLambda"][0] = (0.1, 0.0, 0.0)
    mols["gradLambda"][1] = (-0.2, 0.3, 0.0)
    eps, sigma = 2.0, 0.9
    h = adress_create(oracle, 0.0, 2.5 * sigma, sigma, eps)
    n_act = C.c_int64()
    e = oracle.lib().or_adress_run(h, mols.ctypes.data, 2, counts.ctypes.data, neigh.ctypes.data, neigh.shape[1],
                                   atoms.ctypes.data, C.byref(n_act))
    assert n_act.value == 4
    table = oracle.lj_table(0.0, 2.5 * sigma, sigma, eps, shifted=True)
    ff, en = C.c_double(), C.c_double()
    esum = 0.0
    for i in (0, 1):
        for j in (2, 3):
            d2 = float(((atoms["pos"][i] - atoms["pos"][j]) ** 2).sum())
            oracle.lib().or_lj_force_energy(C.addressof(table), 0, d2, C.byref(ff), C.byref(en))
            esum += en.value
    w = 0.5 * (0.5 + 0.25)
    assert abs(e - esum * w) < 1e-13 * abs(esum)
    assert np.allclose(mols["force"][0], -0.5 * esum * mols["gradLambda"][0], rtol=1e-13)
    assert np.allclose(mols["force"][1], -0.5 * esum * mols["gradLambda"][1], rtol=1e-13)
    a = h.contents
    assert a.runCounter == 1
    # run 0 is a sampling run AND an update run: histograms were folded into the mean (factor 10) and reset
    mean = np.ctypeslib.as_array(a.meanCompensationEnergy, shape=(200,))
    assert abs(mean[100] - (0.5 * esum / 2) / 11.0) < 1e-14  # bin(0.5)=100: sum Vij = esum/2 over 2 atoms
    assert abs(mean[60] - (0.5 * esum / 2) / 11.0) < 1e-14   # bin(0.3)=60
    assert np.count_nonzero(mean) == 2
    assert np.all(np.ctypeslib.as_array(a.compensationEnergy, shape=(200,)) == 0)
    oracle.lib().or_adress_destroy(h)


def test_update_molecules_and_force_scatter(oracle):
    """DiamondFixture (mrmd/test/DiamondFixture.hpp:54-107); UpdateMolecules.test.cpp:43-60;
    ContributeMoleculeForceToAtoms.test.cpp:27-51."""
    atoms = np.zeros(4, dtype=oracle.ATOM)
    atoms["pos"] = [(0, 0, 0), (1 / 3, 1, 0), (-1 / 3, -1, 0), (0, 0, 0)]
    atoms["relMass"] = (0.25, 0.75, 0.75, 0.25)
    mols = np.zeros(2, dtype=oracle.MOLECULE)
    mols["atomsOffset"] = (0, 2)
    mols["numAtoms"] = (2, 2)
    w = oracle.make_weight(oracle.WEIGHT_SLAB, (0, 0, 0), 1.0, 1.0, 1)
    oracle.lib().or_update_molecules(mols.ctypes.data, 2, atoms.ctypes.data, C.byref(w))
    assert np.allclose(mols["pos"][0], (0.25, 0.75, 0)) and np.allclose(mols["pos"][1], (-0.25, -0.75, 0))
    assert np.all(mols["lambda"] == 1.0) and np.all(mols["modLambda"] == 1.0)
    mols["force"][0] = (1, 2, 3)
    mols["force"][1] = (-4, -5, -6)
    oracle.lib().or_contribute_molecule_force(mols.ctypes.data, 2, atoms.ctypes.data)
    assert np.allclose(atoms["force"][0], 0.25 * np.array([1, 2, 3.0]))
    assert np.allclose(atoms["force"][1], 0.75 * np.array([1, 2, 3.0]))
    assert np.allclose(atoms["force"][2], 0.75 * np.array([-4, -5, -6.0]))
    assert np.allclose(atoms["force"][3], 0.25 * np.array([-4, -5, -6.0]))


def weight(oracle, w, x, y=0.0, z=0.0):
    lam, mod = C.c_double(), C.c_double()
    grad = np.zeros(3)
    oracle.lib().or_weight_eval(C.byref(w), x, y, z, C.byref(lam), C.byref(mod), grad.ctypes.data)
    return lam.value, mod.value, grad


def test_slab_weighting(oracle):
    """weighting_function/Slab.test.cpp:23-80 + finite-difference check of the gradient."""
    w = oracle.make_weight(oracle.WEIGHT_SLAB, (2, 3, 4), 2.0, 2.0, 1)  # AT |dx|<1, HY 1..3
    assert weight(oracle, w, 2.5)[:2] == (1.0, 1.0)
    assert weight(oracle, w, 5.5)[:2] == (0.0, 0.0) and weight(oracle, w, -1.5)[:2] == (0.0, 0.0)
    prev = 1.0
    for x in np.linspace(3.0, 5.0, 41):
        lam, mod, g = weight(oracle, w, x)
        assert lam <= prev + 1e-15 and 0 <= lam <= 1
        prev = lam
        assert abs(lam - mod) < 1e-15  # nu = 1
    assert abs(weight(oracle, w, 3.0)[0] - 1.0) < 1e-15 and abs(weight(oracle, w, 5.0)[0]) < 1e-15
    w2 = oracle.make_weight(oracle.WEIGHT_SLAB, (2, 3, 4), 2.0, 2.0, 3)
    h = 1e-6
    for x in (3.3, 4.1, 0.4, -0.7):
        _, m0, _ = weight(oracle, w2, x - h)
        _, m1, _ = weight(oracle, w2, x + h)
        lam, mod, g = weight(oracle, w2, x)
        assert abs((m1 - m0) / (2 * h) - g[0]) < 1e-6
        assert g[1] == 0 and g[2] == 0
        assert abs(mod - lam ** 3) < 1e-14
    wa = oracle.make_weight(oracle.WEIGHT_SLAB, (2, 3, 4), 2.0, 2.0, 1, abrupt=True)
    assert weight(oracle, wa, 4.0)[:2] == (1.0, 1.0) and weight(oracle, wa, 5.5)[:2] == (0.0, 0.0)


def test_spherical_weighting(oracle):
    """weighting_function/Spherical.test.cpp:23-85."""
    w = oracle.make_weight(oracle.WEIGHT_SPHERICAL, (2, 3, 4), 2.0, 2.0, 7)
    assert weight(oracle, w, 2.5, 3.5, 4.5)[0] == 1.0
    assert weight(oracle, w, 6.5, 3.0, 4.0)[0] == 0.0
    h = 1e-6
    p = np.array([4.2, 4.1, 4.9])
    lam, mod, g = weight(oracle, w, *p)
    assert 0 < lam < 1 and lam == mod
    for d in range(3):
        e = np.zeros(3)
        e[d] = h
        fd = (weight(oracle, w, *(p + e))[0] - weight(oracle, w, *(p - e))[0]) / (2 * h)
        assert abs(fd - g[d]) < 1e-6


# --- MultiHistogram / thermodynamic force -------------------------------------------------------------
def thermo_fixture(oracle, ramp=True):
    sub = oracle.subdomain([0, 0, 0], [10, 1, 1], 1.0)
    one = np.array([1.0])
    t = oracle.lib().or_thermo_create(one.ctypes.data, 1, C.byref(sub), 1.0, one.ctypes.data, 0, 0)
    assert t.contents.numBins == 10
    f = np.ctypeslib.as_array(t.contents.force, shape=(10,))
    if ramp:
        f[:] = np.arange(10.0)
    return t, f


def atoms_uniform(oracle):
    return atoms_from(oracle, np.stack([np.arange(100) / 10.0, np.zeros(100), np.zeros(100)], axis=1))


def atoms_nonuniform(oracle):
    xs = [i + 0.5 for i in range(10) for _ in range(i)]
    return atoms_from(oracle, np.stack([np.array(xs), np.zeros(45), np.zeros(45)], axis=1))


def test_thermo_sample(oracle):
    t, f = thermo_fixture(oracle)
    a = atoms_uniform(oracle)
    oracle.lib().or_thermo_sample(t, a.ctypes.data, 100)
    assert np.all(np.ctypeslib.as_array(t.contents.density, shape=(10,)) == 10.0)
    oracle.lib().or_thermo_destroy(t)
    t, f = thermo_fixture(oracle)
    a = atoms_nonuniform(oracle)
    oracle.lib().or_thermo_sample(t, a.ctypes.data, 45)
    assert np.all(np.ctypeslib.as_array(t.contents.density, shape=(10,)) == np.arange(10.0))
    assert t.contents.samples == 1
    oracle.lib().or_thermo_destroy(t)


def test_thermo_apply_variants(oracle):
    t, f = thermo_fixture(oracle)
    a = atoms_uniform(oracle)
    oracle.lib().or_thermo_apply(t, a.ctypes.data, 100, None, 0)
    assert np.all(a["force"][:, 0] == np.floor(a["pos"][:, 0])) and np.all(a["force"][:, 1:] == 0)
    pred = oracle.make_pred(oracle.PRED_INTERVAL, 0, 0.0, 0.5, 4.5)
    for interp in (0, 1):
        a = atoms_uniform(oracle)
        oracle.lib().or_thermo_apply(t, a.ctypes.data, 100, C.byref(pred), interp)
        x = a["pos"][:, 0]
        inside = (x > 0.5) & (x < 4.5)
        want = np.where(inside, x - 0.5 if interp else np.floor(x), 0.0)
        assert np.allclose(a["force"][:, 0], want, rtol=0, atol=1e-12)
    oracle.lib().or_thermo_destroy(t)


def test_thermo_update(oracle):
    t, f = thermo_fixture(oracle)
    a = atoms_nonuniform(oracle)
    oracle.lib().or_thermo_sample(t, a.ctypes.data, 45)
    oracle.lib().or_thermo_update(t, 1.0, 0.0, None)
    assert np.allclose(f, np.arange(10.0) - 1.0, atol=1e-13)
    assert t.contents.samples == 0 and np.all(np.ctypeslib.as_array(t.contents.density, shape=(10,)) == 0)
    oracle.lib().or_thermo_destroy(t)
    t, f = thermo_fixture(oracle)
    oracle.lib().or_thermo_sample(t, a.ctypes.data, 45)
    pred = oracle.make_pred(oracle.PRED_INTERVAL, 0, 0.0, 0.51, 4.49)
    oracle.lib().or_thermo_update(t, 1.0, 0.0, C.byref(pred))
    grid = np.arange(10) + 0.5
    want = np.where((grid > 0.51) & (grid < 4.49), np.arange(10.0) - 1.0, np.arange(10.0))
    assert np.allclose(f, want, atol=1e-13)
    mu_l, mu_r = np.zeros(1), np.zeros(1)
    oracle.lib().or_thermo_mu(t, mu_l.ctypes.data, mu_r.ctypes.data)
    assert abs(mu_l[0] - want[:5].sum()) < 1e-12 and abs(mu_r[0] - want[5:].sum()) < 1e-12
    oracle.lib().or_thermo_destroy(t)


def test_multihistogram_ops(oracle):
    """data/MultiHistogram.test.cpp: scale, makeSymmetric, gradient, smoothen properties."""
    L = oracle.lib()
    nb, nh = 10, 2
    d = np.arange(nb * nh, dtype=float).reshape(nb, nh).copy()
    L.or_hist_scale(d.ctypes.data, nb, nh, 0.5)
    assert np.all(d == np.arange(nb * nh).reshape(nb, nh) * 0.5)
    fac = np.array([2.0, 3.0])
    L.or_hist_scale_per_hist(d.ctypes.data, nb, nh, fac.ctypes.data)
    assert np.all(d[:, 1] == np.arange(1, 20, 2) * 1.5)
    s = np.arange(nb * nh, dtype=float).reshape(nb, nh).copy()
    L.or_hist_make_symmetric(s.ctypes.data, nb, nh)
    assert np.all(s == s[::-1]) and np.all(s[:, 0] == 9.0)
    x = np.arange(nb, dtype=float)[:, None] * np.array([[1.0, 2.0]])
    g = np.zeros_like(x)
    L.or_hist_gradient(np.ascontiguousarray(x).ctypes.data, g.ctypes.data, 0.0, 5.0, nb, nh, 0)
    assert np.allclose(g[:, 0], 2.0) and np.allclose(g[:, 1], 4.0)  # binSize 0.5
    L.or_hist_gradient(np.ascontiguousarray(x).ctypes.data, g.ctypes.data, 0.0, 5.0, nb, nh, 1)
    assert np.isclose(g[0, 0], (1 - 9) / 1.0) and np.isclose(g[9, 0], (0 - 8) / 1.0)
    c = np.full((nb, nh), 3.0)
    out = np.zeros_like(c)
    for periodic in (0, 1):
        L.or_hist_smoothen(c.ctypes.data, out.ctypes.data, 0.0, 5.0, nb, nh, 1.0, 2.0, periodic)
        assert np.allclose(out, 3.0)  # constants are preserved
    bump = np.zeros((nb, 1))
    bump[4] = bump[5] = 1.0
    out = np.zeros_like(bump)
    L.or_hist_smoothen(bump.ctypes.data, out.ctypes.data, 0.0, 5.0, nb, 1, 1.0, 2.0, 0)
    assert np.allclose(out, out[::-1])  # symmetric input stays symmetric
    assert L.or_hist_get_bin(0.0, 10.0, 10, 9.999) == 9 and L.or_hist_get_bin(0.0, 10.0, 10, 10.0) == -1
    assert L.or_hist_get_bin(0.0, 10.0, 10, -0.001) == -1 and L.or_hist_get_bin(0.0, 10.0, 10, 0.0) == 0


def test_subdomain(oracle):
    """data/Subdomain.test.cpp: derived corners and scaleDim."""
    s = oracle.subdomain([1, 2, 3], [3, 6, 9], [0.1, 0.2, 0.3])
    assert np.allclose(s.minGhostCorner, (0.9, 1.8, 2.7)) and np.allclose(s.maxGhostCorner, (3.1, 6.2, 9.3))
    assert np.allclose(s.minInnerCorner, (1.1, 2.2, 3.3)) and np.allclose(s.maxInnerCorner, (2.9, 5.8, 8.7))
    assert np.allclose(s.diameter, (2, 4, 6)) and np.allclose(s.diameterWithGhostLayer, (2.2, 4.4, 6.6))
    oracle.lib().or_subdomain_scale_dim(C.byref(s), 2.0, 1)
    assert np.allclose(s.minCorner, (1, 4, 3)) and np.allclose(s.maxCorner, (3, 12, 9)) and np.allclose(s.diameter, (2, 8, 6))


def test_cell_sort(oracle):
    """LinkedCellList + permute as used in tests/NVT/NVT.cpp:136-144."""
    rng = np.random.default_rng(3)
    n = 5000
    a = atoms_from(oracle, rng.random((n, 3)) * 10.0)
    a["vel"][:, 0] = np.arange(n)  # tag
    delta = np.array([2.6, 2.6, 2.6])
    lo, hi = np.zeros(3), np.full(3, 10.0)
    cid = np.zeros(n, dtype=np.int32)
    dims = np.zeros(3, dtype=np.int32)
    nc = oracle.lib().or_cell_ids(a.ctypes.data, 13, 0, n, delta.ctypes.data, lo.ctypes.data, hi.ctypes.data,
                                  cid.ctypes.data, dims.ctypes.data)
    assert tuple(dims) == (3, 3, 3) and nc == 27
    dx = 10.0 / 3
    want = (np.floor(a["pos"] * (1.0 / dx)).astype(int) * np.array([9, 3, 1])).sum(axis=1)
    assert np.all(cid == want)
    perm = np.zeros(n, dtype=np.int64)
    off = np.zeros(nc + 1, dtype=np.int64)
    oracle.lib().or_cell_perm(cid.ctypes.data, 0, n, nc, perm.ctypes.data, off.ctypes.data)
    oracle.lib().or_permute_atoms(a.ctypes.data, 0, n, perm.ctypes.data)
    cid2 = np.zeros(n, dtype=np.int32)
    oracle.lib().or_cell_ids(a.ctypes.data, 13, 0, n, delta.ctypes.data, lo.ctypes.data, hi.ctypes.data,
                             cid2.ctypes.data, None)
    assert np.all(np.diff(cid2) >= 0)
    same = np.diff(cid2) == 0
    assert np.all(np.diff(a["vel"][:, 0])[same] > 0)  # stable within a cell
    assert sorted(a["vel"][:, 0].astype(int)) == list(range(n))


def test_analysis_kats(oracle):
    """analysis:: diagnostics (row (f)1): mrmd/analysis/KineticEnergy.test.cpp:23-53, SystemMomentum.test.cpp:23-42,
    MeanSquareDisplacement.test.cpp:28-35; pressure has no reference test: hand-computed from Pressure.cpp:23-51"""
    import ctypes as C

    L = oracle.lib()
    a = np.zeros(3, dtype=oracle.ATOM)
    a["vel"] = [(2, 0, 0), (0, -8, 0), (0, 0, 16)]
    a["mass"] = [1.0, 2.0, 0.5]
    assert float_eq(L.or_kinetic_energy(a.ctypes.data, 3), (4 + 2 * 64 + 0.5 * 256) * 0.5)
    b = np.zeros(2, dtype=oracle.ATOM)
    b["vel"] = [(2, 3, 4), (-4, -8, -16)]
    mom = np.zeros(3)
    L.or_system_momentum(b.ctypes.data, 2, mom.ctypes.data)
    assert all(float_eq(mom[d], w) for d, w in enumerate((-2.0, -5.0, -12.0)))
    sub = oracle.subdomain([0, 0, 0], [10, 10, 10], 1.0)
    c = np.zeros(1, dtype=oracle.ATOM)
    c["pos"] = [(1, 2, 3)]
    init = np.array([1.0, 2.0, 3.0])
    assert L.or_msd(c.ctypes.data, init.ctypes.data, 1, C.byref(sub)) == 0.0
    # one fold: |dx| = 6 > L/2 -> 6 - 10 = -4
    init = np.array([7.0, 2.0, 3.0])
    assert float_eq(L.or_msd(c.ctypes.data, init.ctypes.data, 1, C.byref(sub)), 16.0)
    c["vel"], c["force"], c["mass"] = [(1, 1, 1)], [(1, 0, -1)], 2.0
    assert float_eq(L.or_pressure(c.ctypes.data, 1, C.byref(sub)), (2 * 3 + (1 - 3)) / 3000.0)


def _diamond(oracle):
    """mrmd/test/DiamondFixture.hpp:26-106: two molecules of two atoms each"""
    a = np.zeros(4, dtype=oracle.ATOM)
    a["pos"] = [(1, 0, 0), (0, 1, 0), (-1, 0, 0), (0, -1, 0)]
    a["mass"] = [1, 3, 1, 3]
    a["relMass"] = [0.25, 0.75, 0.25, 0.75]
    m = np.zeros(2, dtype=oracle.MOLECULE)
    m["atomsOffset"], m["numAtoms"] = [0, 2], [2, 2]
    return a, m


def _integrate_position(a, dt):
    """integratePosition of mrmd/action/Shake.test.cpp:27-46"""
    a["pos"] += dt * a["vel"] + (0.5 * dt * dt / a["mass"])[:, None] * a["force"]


def test_berendsen_and_shake_kats(oracle):
    """row (f)2: mrmd/action/BerendsenThermostat.test.cpp:29-47, BerendsenBarostat.test.cpp:28-64,
    Shake.test.cpp:59-249"""
    import ctypes as C

    L = oracle.lib()
    one = np.zeros(1, dtype=oracle.ATOM)  # test::SingleAtom
    one["pos"], one["vel"], one["force"], one["mass"] = [(2, 3, 4)], [(7, 5, 3)], [(9, 7, 8)], 1.5

    def temperature(a):
        return L.or_kinetic_energy(a.ctypes.data, 1) * (2.0 / 3.0)

    a = one.copy()
    L.or_berendsen_thermostat(a.ctypes.data, 1, temperature(a), 3.8, 0.0)
    assert float_eq(temperature(a), 41.5)
    L.or_berendsen_thermostat(a.ctypes.data, 1, temperature(a), 3.8, 1.0)
    assert float_eq(temperature(a), 3.8)

    a, sub = one.copy(), oracle.subdomain([0, 0, 0], [1, 1, 1], 0.1)
    L.or_berendsen_barostat(a.ctypes.data, 1, 1.0, 3.8, 0.0, C.byref(sub), 1, 1, 1)
    assert [float_eq(sub.maxCorner[d], 1.0) for d in range(3)] == [True] * 3 and np.allclose(a["pos"][0], [2, 3, 4])
    L.or_berendsen_barostat(a.ctypes.data, 1, 2.0, 1.0, 1.0, C.byref(sub), 1, 1, 1)
    mu = 2.0 ** (1.0 / 3.0)
    assert all(float_eq(sub.maxCorner[d], mu) for d in range(3))
    assert all(float_eq(a["pos"][0][d], w * mu) for d, w in enumerate((2, 3, 4)))

    dt = 0.1
    whole = np.zeros(1, dtype=oracle.MOLECULE)
    whole["atomsOffset"], whole["numAtoms"] = 0, 4

    def dist(a, i, j):
        return float(np.linalg.norm(a["pos"][i] - a["pos"][j]))

    for eq, sign in ((1.0, -1.0), (2.0, 1.0)):  # Attraction / Repulsion: one constraint between atoms 0 and 1
        a, _ = _diamond(oracle)
        bonds, eqs = np.array([0, 1], dtype=np.int64), np.array([eq])
        assert L.or_shake_positional(whole.ctypes.data, 1, a.ctypes.data, 4, bonds.ctypes.data, eqs.ctypes.data, 1, 1, dt) == 0
        assert np.allclose(a["force"][0], -a["force"][1]) and sign * a["force"][0][0] > 0 and sign * a["force"][0][1] < 0
        _integrate_position(a, dt)
        assert float_eq(dist(a, 0, 1), eq)
    for eq in (1.0, 2.0):  # Shrink / Grow: the ring 0-1-2-3-0, ten iterations
        a, _ = _diamond(oracle)
        bonds, eqs = np.array([0, 1, 1, 2, 2, 3, 3, 0], dtype=np.int64), np.full(4, eq)
        assert L.or_shake_positional(whole.ctypes.data, 1, a.ctypes.data, 4, bonds.ctypes.data, eqs.ctypes.data, 4, 10, dt) == 0
        _integrate_position(a, dt)
        assert all(float_eq(dist(a, i, (i + 1) % 4), eq) for i in range(4))
    a, m = _diamond(oracle)  # Molecules: the bond 0-1 in both molecules
    bonds, eqs = np.array([0, 1], dtype=np.int64), np.array([1.0])
    assert L.or_shake_positional(m.ctypes.data, 2, a.ctypes.data, 4, bonds.ctypes.data, eqs.ctypes.data, 1, 1, dt) == 0
    f = a["force"]
    assert np.allclose(f[0], -f[1]) and f[0][0] < 0 < f[0][1] and np.allclose(f[2], -f[3]) and f[2][0] > 0 > f[2][1]
    # RATTLE: after the projection the relative velocity has no component along the bond
    a["vel"] = [(1, 2, 3), (-1, 0.5, 0), (0, 0, 1), (2, 2, 2)]
    assert L.or_shake_velocity(m.ctypes.data, 2, a.ctypes.data, bonds.ctypes.data, 1) == 0
    for i, j in ((0, 1), (2, 3)):
        assert abs(np.dot(a["vel"][i] - a["vel"][j], a["pos"][i] - a["pos"][j])) < 1e-14
    bad = np.array([0, 2], dtype=np.int64)  # MRMD_DEVICE_ASSERT_LESS: not enough atoms in molecule to satisfy bond
    assert L.or_shake_velocity(m.ctypes.data, 2, a.ctypes.data, bad.ctypes.data, 1) == -1


# ---------------------------------------------------------------------------------------------
# row (f)4: Coulomb / CoulombDSF / SPC water
def spc_water(oracle, origins):
    """the molecule of mrmd/action/SPC.test.cpp:51-96, once per origin"""
    eq_ho, angle = 0.1, 109.47 / 180.0 * math.pi
    n = len(origins)
    atoms = np.zeros(3 * n, dtype=oracle.ATOM)
    mols = np.zeros(n, dtype=oracle.MOLECULE)
    m_tot = 15.999 + 2 * 1.008
    for i, o in enumerate(np.asarray(origins, dtype=np.float64)):
        atoms["pos"][3 * i] = o
        atoms["pos"][3 * i + 1] = o + (eq_ho, 0, 0)
        atoms["pos"][3 * i + 2] = o + (eq_ho * math.cos(angle), eq_ho * math.sin(angle), 0)
        atoms["type"][3 * i:3 * i + 3] = (0, 1, 1)
        atoms["mass"][3 * i:3 * i + 3] = (15.999, 1.008, 1.008)
        atoms["charge"][3 * i:3 * i + 3] = (-0.82, 0.41, 0.41)
        atoms["relMass"][3 * i:3 * i + 3] = np.array((15.999, 1.008, 1.008)) / m_tot
        mols["pos"][i] = o
        mols["atomsOffset"][i], mols["numAtoms"][i] = 3 * i, 3
    return atoms, mols


def test_approx_erfc_and_coulomb_dsf_kats(oracle):
    """mrmd/util/math.test.cpp:45-51, mrmd/action/CoulombDSF.test.cpp:63-117"""
    L = oracle.lib()
    for x in np.arange(0.1, 3.0, 0.1):
        assert abs(math.erfc(x) - L.or_approx_erfc(float(x))) < 1e-6
    isp = 1.0 / math.sqrt(math.pi)

    def energy_dsf(r, q1, q2, alpha, rc):
        b = math.erfc(alpha * r) / r - math.erfc(alpha * rc) / rc
        b += (math.erfc(alpha * rc) / (rc * rc) + 2 * alpha * isp * math.exp(-alpha * alpha * rc * rc) / rc) * (r - rc)
        return 138.935458 * q1 * q2 * b

    def force_dsf(r, q1, q2, alpha, rc):
        b = math.erfc(alpha * r) / (r * r) + 2 * alpha * isp * math.exp(-alpha * alpha * r * r) / r
        b -= math.erfc(alpha * rc) / (rc * rc) + 2 * alpha * isp * math.exp(-alpha * alpha * rc * rc) / rc
        return 138.935458 * q1 * q2 * b / r

    def evaluate(kind, rc, alpha, x, q1, q2):
        d = np.ascontiguousarray(np.asarray(x, dtype=np.float64) ** 2)
        f, e = np.zeros_like(d), np.zeros_like(d)
        L.or_coulomb_eval(kind, rc, alpha, d.ctypes.data, len(d), q1, q2, f.ctypes.data, e.ctypes.data)
        return f, e

    rc, alpha, q1, q2 = 5.0, 0.1, 1.2, -1.3
    xs = np.arange(1e-8, rc, 0.01)
    f, e = evaluate(1, rc, alpha, xs, q1, q2)
    for x, fi, ei in zip(xs, f, e):
        assert abs((fi - force_dsf(x, q1, q2, alpha, rc)) / force_dsf(x, q1, q2, alpha, rc)) < 1e-4  # ForceExplicitComparison
        if x < rc - 1.0:
            assert abs((ei - energy_dsf(x, q1, q2, alpha, rc)) / energy_dsf(x, q1, q2, alpha, rc)) < 1e-5  # Energy...
    pp, mm, pm, mp = (evaluate(1, 1.0, 0.1, [1.0], a, b)[0][0] for a, b in ((1, 1), (-1, -1), (1, -1), (-1, 1)))
    assert float_eq(pp, mm) and float_eq(pm, mp) and float_eq(pp, -pm)  # Symmetry
    assert abs(evaluate(1, 1.5, 0.1, [1.5], 1.0, 1.0)[0][0]) < 1e-5  # shift
    f, e = evaluate(0, 0.0, 0.0, [2.0], 0.5, -0.25)  # action/Coulomb.hpp:30-43
    assert f[0] == 138.935458 * 0.5 * -0.25 / 4.0 and e[0] == 138.935458 * 0.5 * -0.25 / 2.0


def test_spc_kats(oracle):
    """mrmd/action/SPC.test.cpp:131-152 (an equilibrium molecule needs no constraint force) and SPC::applyForces /
    calcBondEnergy (SPC.hpp:143-344) for two molecules against the formulas written out in numpy"""
    L = oracle.lib()
    eq_ho, angle = 0.1, 109.47 / 180.0 * math.pi
    eq_hh = eq_ho * math.sqrt(2.0 - 2.0 * math.cos(angle))
    atoms, mols = spc_water(oracle, [(0, 0, 0)])
    bonds, eqs = np.array([0, 1, 0, 2, 1, 2], dtype=np.int64), np.array([eq_ho, eq_ho, eq_hh])
    assert L.or_shake_positional(mols.ctypes.data, 1, atoms.ctypes.data, 3, bonds.ctypes.data, eqs.ctypes.data, 3, 20, 0.1) == 0
    assert all(float_eq(v + 1.0, 1.0) for v in atoms["force"].ravel())
    assert abs(L.or_spc_bond_energy(mols.ctypes.data, 1, atoms.ctypes.data, 3, 1000.0)) < 1e-25

    atoms, mols = spc_water(oracle, [(0, 0, 0), (0.3, 0.1, -0.05)])
    counts, neigh = np.array([1, 0], dtype=np.int32), np.array([[1], [0]], dtype=np.int32)
    en = np.zeros(2)
    L.or_spc_apply_forces(mols.ctypes.data, 2, counts.ctypes.data, neigh.ctypes.data, 1, atoms.ctypes.data, 0, en.ctypes.data)
    sigma, eps, rcut = 0.31655578901998815, 0.6501695808187486, 1.2
    table = oracle.lj_table(0.7 * sigma, rcut, sigma, eps, 1, True)
    force = np.zeros((6, 3))
    d = atoms["pos"][0] - atoms["pos"][3]
    ff, e = C.c_double(), C.c_double()
    L.or_lj_force_energy(C.addressof(table), 0, float(d @ d), C.byref(ff), C.byref(e))
    force[0] += d * ff.value
    force[3] -= d * ff.value
    e_c = 0.0
    for i in range(3):
        for j in range(3, 6):
            d = atoms["pos"][i] - atoms["pos"][j]
            pre = 138.935458 * atoms["charge"][i] * atoms["charge"][j]
            force[i] += d * pre / (d @ d)
            force[j] -= d * pre / (d @ d)
            e_c += pre / math.sqrt(d @ d)
    assert abs(en[0] - e.value) <= 1e-14 * abs(e.value) and abs(en[1] - e_c) <= 1e-13 * abs(e_c)
    assert np.allclose(atoms["force"], force, rtol=1e-13, atol=1e-12)
    assert np.abs(atoms["force"].sum(axis=0)).max() < 1e-10  # Newton's third law
    # the DSF flavour vanishes at the cutoff, the plain one does not
    far, mfar = spc_water(oracle, [(0, 0, 0), (1.25, 0, 0)])
    L.or_spc_apply_forces(mfar.ctypes.data, 2, counts.ctypes.data, neigh.ctypes.data, 1, far.ctypes.data, 1, en.ctypes.data)
    assert en[0] == 0.0 and abs(en[1]) < 1e-2
    plain = np.zeros(2)
    far["force"] = 0.0
    L.or_spc_apply_forces(mfar.ctypes.data, 2, counts.ctypes.data, neigh.ctypes.data, 1, far.ctypes.data, 0, plain.ctypes.data)
    assert abs(plain[1]) > 1.0


def test_oracle_molecule_loops_hold_their_bonds(oracle):
    """the oracle's constrained step loops (oracle/md_loop.py: SPC water, AdResS tetramers) keep the bond lengths that
    SHAKE / RATTLE enforce and stay finite: the GPU parity tests of rows (f)2 / (f)4 and configs[3] compare against them"""
    from mrmd_b200.workloads import tetramer_system
    from oracle.md_loop import OracleAdressMD, OracleSpcMD, spc_water_box

    pos, vel, mass, q, rm, typ, box = spc_water_box(6)
    md = OracleSpcMD(pos, vel, mass, q, rm, typ, box, dt=0.0005, skin=0.02)
    st = md.run(8)
    p = md.atoms["pos"][:md.n].reshape(-1, 3, 3)
    assert np.abs(np.linalg.norm(p[:, 0] - p[:, 1], axis=1) - 0.1).max() < 1e-5
    assert np.isfinite(st["energyCoulomb"]) and st["energyLJ"] < 0 and st["rebuilds"] >= 1

    pos, vel, box = tetramer_system(6)
    w = oracle.make_weight(oracle.WEIGHT_SPHERICAL, 0.5 * box, 3.0, 2.0, 2)
    md = OracleAdressMD(pos, vel, box, w, atoms_per_mol=4, constraint_iterations=3, langevin=True, max_neigh=40)
    st = md.run(20)
    p = md.atoms["pos"][:md.n].reshape(-1, 4, 3)
    iu = np.triu_indices(4, 1)
    d = np.linalg.norm(p[:, :, None, :] - p[:, None, :, :], axis=-1)[:, iu[0], iu[1]]
    assert np.abs(d - 1.0).max() < 5e-3 and st["rebuilds"] >= 2 and st["pairInteractions"] > 0


def test_limit_acceleration_and_velocity_kats(oracle):
    """mrmd/action/LimitAcceleration.test.cpp:27-38, LimitVelocity.test.cpp:27-38 on mrmd/test/SingleAtom.hpp:28-55"""
    L = oracle.lib()
    a = np.zeros(1, dtype=oracle.ATOM)
    a["pos"], a["vel"], a["force"], a["mass"] = (2.0, 3.0, 4.0), (7.0, 5.0, 3.0), (9.0, 7.0, 8.0), 1.5
    L.or_limit_acceleration(a.ctypes.data, 1, 0.5)
    assert all(float_eq(v, 0.75) for v in a["force"][0])
    L.or_limit_velocity(a.ctypes.data, 1, 0.5)
    assert all(float_eq(v, 0.5) for v in a["vel"][0])
    a["vel"], a["force"] = (-7.0, 0.25, -0.5), (-9.0, 0.3, -0.75)
    L.or_limit_acceleration(a.ctypes.data, 1, 0.5)
    L.or_limit_velocity(a.ctypes.data, 1, 0.5)
    assert np.allclose(a["force"][0], (-0.75, 0.3, -0.75), rtol=1e-15) and np.array_equal(a["vel"][0], (-0.5, 0.25, -0.5))
