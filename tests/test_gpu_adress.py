"""AdResS + thermodynamic-force parity: CUDA path vs the oracle and the reference's KATs."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FORCE_RTOL = 1e-10


@pytest.fixture(scope="module")
def api():
    from mrmd_b200 import api as a

    assert a.L().mrmd_b200_device_count() > 0
    return a


def float_eq(a, b):
    a32, b32 = np.float32(a), np.float32(b)
    return abs(float(a32) - float(b32)) <= 4 * np.spacing(max(abs(a32), abs(b32), np.float32(1e-30)))


def kat_fixture(api):
    """mrmd/action/LJ_IdealGas.test.cpp:34-125"""
    mols = api.Molecules(2)
    mols.set("pos", [(-0.5, 0, 0), (0.5, 0, 0)])
    mols.set("atomsOffset", [0, 2])
    mols.set("numAtoms", [2, 2])
    mols.numLocalMolecules = 2
    atoms = api.Atoms.from_arrays(np.array([(-0.5, -0.5, 0), (-0.5, 0.5, 0), (0.5, -0.5, 0), (0.5, 0.5, 0)]),
                                  relativeMass=0.5)
    vl = api.HalfVerletList()
    vl.build(mols, 0, 2, 2.0, 1.0, [-1, -1, -1], [1, 1, 1], 4)
    c, nb = vl.to_host()
    assert tuple(c) == (1, 0) and nb[0, 0] == 1
    return mols, atoms, vl


@pytest.mark.parametrize("lam,scale", [(0.0, 0.0), (0.5, 0.5), (1.0, 1.0)])
def test_lj_idealgas_kat(api, lam, scale):
    mols, atoms, vl = kat_fixture(api)
    mols.fill("modulatedLambda", lam)
    base = 2.0 if lam == 0.0 else 0.0
    atoms.fill("force", base)
    lj = api.LJ_IdealGas(0.0, 2.5 * 0.9, 0.9, 2.0, True)
    lj.run(mols, vl, atoms)
    xf, yf = 0.22156665 * scale, 1.3825009 * scale
    want = np.array([(-xf, +yf, 0), (-xf, -yf, 0), (+xf, +yf, 0), (+xf, -yf, 0)]) + base
    f = atoms.getForce()
    for i in range(4):
        for d in range(3):
            assert float_eq(f[i, d], want[i, d]), (i, d, f[i, d], want[i, d])


def test_update_molecules_and_scatter(api):
    """DiamondFixture; UpdateMolecules.test.cpp:43-60; ContributeMoleculeForceToAtoms.test.cpp:27-51"""
    atoms = api.Atoms.from_arrays(np.array([(0, 0, 0), (1 / 3, 1, 0), (-1 / 3, -1, 0), (0, 0, 0)]),
                                  relativeMass=np.array([0.25, 0.75, 0.75, 0.25]))
    mols = api.Molecules(2)
    mols.set("atomsOffset", [0, 2])
    mols.set("numAtoms", [2, 2])
    mols.numLocalMolecules = 2
    api.UpdateMolecules.update(mols, atoms, api.Slab((0, 0, 0), 1.0, 1.0, 1))
    p = mols.get("pos")
    assert np.allclose(p[0], (0.25, 0.75, 0)) and np.allclose(p[1], (-0.25, -0.75, 0))
    assert np.all(mols.get("lambda") == 1.0) and np.all(mols.get("modulatedLambda") == 1.0)
    mols.set("force", [(1, 2, 3), (-4, -5, -6)])
    api.ContributeMoleculeForceToAtoms.update(mols, atoms)
    f = atoms.getForce()
    assert np.allclose(f[0], 0.25 * np.array([1, 2, 3.0])) and np.allclose(f[1], 0.75 * np.array([1, 2, 3.0]))
    assert np.allclose(f[2], 0.75 * np.array([-4, -5, -6.0])) and np.allclose(f[3], 0.25 * np.array([-4, -5, -6.0]))


def test_weighting_functions_vs_oracle(api, oracle):
    rng = np.random.default_rng(9)
    pts = rng.random((4000, 3)) * 12 - 2
    pts[:4] = [(3.0, 3, 4), (5.0, 3, 4), (1.0, 3, 4), (-1.0, 3, 4)]  # region boundaries of the slab
    for w, ow in [
        (api.Slab((2, 3, 4), 2.0, 2.0, 1), oracle.make_weight(oracle.WEIGHT_SLAB, (2, 3, 4), 2.0, 2.0, 1)),
        (api.Slab((2, 3, 4), 2.0, 2.0, 3), oracle.make_weight(oracle.WEIGHT_SLAB, (2, 3, 4), 2.0, 2.0, 3)),
        (api.Slab((2, 3, 4), 2.0, 2.0, 1, abrupt=True), oracle.make_weight(oracle.WEIGHT_SLAB, (2, 3, 4), 2.0, 2.0, 1, True)),
        (api.Spherical((2, 3, 4), 2.0, 2.0, 7), oracle.make_weight(oracle.WEIGHT_SPHERICAL, (2, 3, 4), 2.0, 2.0, 7)),
    ]:
        lam, mod, grad = api.weight_eval(w, pts)
        ol, om, og = np.zeros(len(pts)), np.zeros(len(pts)), np.zeros((len(pts), 3))
        l_, m_ = C.c_double(), C.c_double()
        g_ = np.zeros(3)
        for i, p in enumerate(pts):
            oracle.lib().or_weight_eval(C.byref(ow), p[0], p[1], p[2], C.byref(l_), C.byref(m_), g_.ctypes.data)
            ol[i], om[i], og[i] = l_.value, m_.value, g_
        # region decisions are exact; values differ by CUDA-vs-libm cos/sin ulps
        assert np.array_equal(lam == 1.0, ol == 1.0) and np.array_equal(lam == 0.0, ol == 0.0)
        assert np.abs(lam - ol).max() < 1e-14 and np.abs(mod - om).max() < 1e-14
        assert np.abs(grad - og).max() < 1e-13 * max(1.0, np.abs(og).max())


def adress_system(n_side, atoms_per_mol, seed):
    rng = np.random.default_rng(seed)
    spacing = 1.6 if atoms_per_mol > 1 else 1.25
    sites = (np.stack(np.meshgrid(*[np.arange(n_side)] * 3, indexing="ij"), axis=-1).reshape(-1, 3) + 0.5) * spacing
    sites = sites + (rng.random(sites.shape) - 0.5) * 0.5
    box = n_side * spacing
    m = len(sites)
    if atoms_per_mol == 1:
        apos = sites.copy()
    else:
        tet = np.array([(1, 1, 1), (1, -1, -1), (-1, 1, -1), (-1, -1, 1)]) * (0.5 / np.sqrt(2.0))
        apos = (sites[:, None, :] + tet[None, :atoms_per_mol, :]).reshape(-1, 3)
    return sites, apos, box, m


@pytest.mark.parametrize("atoms_per_mol,weight_kind,lanes", [(1, "slab", 0), (4, "slab", 0), (4, "spherical", 0),
                                                              (4, "slab", 4), (4, "spherical", 4)])
def test_adress_step_vs_oracle(api, oracle, atoms_per_mol, weight_kind, lanes):
    """The assembled AdResS force step of SURVEY.md section 3.5 on both sides, two runs (run 0 samples and
    updates the compensation histograms; run 1 applies the mean compensation energy).  lanes = 4: the kernel with four
    lanes per four-atom molecule (mrmd_b200_adress_set_atoms_per_molecule)."""
    sites, apos, box, M = adress_system(10, atoms_per_mol, 21)
    a_per = atoms_per_mol
    N = M * a_per
    rc, skin = 2.5, 0.1
    types = (np.arange(N) % 2).astype(np.int64)
    relmass = np.full(N, 1.0 / a_per)
    center = [box / 2] * 3
    if weight_kind == "slab":
        w = api.Slab(center, 4.0, 3.0, 2)
        ow = oracle.make_weight(oracle.WEIGHT_SLAB, center, 4.0, 3.0, 2)
    else:
        w = api.Spherical(center, 3.0, 3.0, 2)
        ow = oracle.make_weight(oracle.WEIGHT_SPHERICAL, center, 3.0, 3.0, 2)
    nt = 2
    capv, rcv = np.full(4, 0.7), np.full(4, rc)
    sig, eps = np.array([1.0, 0.9, 0.9, 0.8]), np.array([1.0, 1.1, 1.1, 1.2])

    sub = api.Subdomain([0, 0, 0], [box] * 3, rc + skin)
    atoms = api.Atoms.from_arrays(apos, None, mass=1.0, type=types, relativeMass=relmass)
    mols = api.Molecules(M)
    mols.set("atomsOffset", np.arange(M) * a_per)
    mols.set("numAtoms", np.full(M, a_per))
    mols.numLocalMolecules = M
    api.UpdateMolecules.update(mols, atoms, w)  # COM needed by the exchange
    ghost = api.MultiResGhostLayer()
    ghost.exchangeRealAtoms(mols, atoms, sub)
    ghost.createGhostAtoms(mols, atoms, sub)
    vl = api.HalfVerletList()
    vl.build(mols, 0, M, rc + skin, 1.0, sub.minGhostCorner, sub.maxGhostCorner, 40)
    lj = api.LJ_IdealGas(capv, rcv, sig, eps, True, numTypes=nt)
    lj.setAtomsPerMolecule(lanes)

    L = oracle.lib()
    osub = oracle.subdomain([0, 0, 0], [box] * 3, rc + skin)
    oa = np.zeros(5 * N, dtype=oracle.ATOM)
    oa["pos"][:N], oa["mass"][:N], oa["type"][:N], oa["relMass"][:N] = apos, 1.0, types, relmass
    om = np.zeros(5 * M, dtype=oracle.MOLECULE)
    om["atomsOffset"][:M], om["numAtoms"][:M] = np.arange(M) * a_per, a_per
    L.or_update_molecules(om.ctypes.data, M, oa.ctypes.data, C.byref(ow))
    L.or_mr_periodic_map(om.ctypes.data, M, oa.ctypes.data, C.byref(osub))
    corr = np.zeros(len(oa), dtype=np.int64)
    out = np.zeros(2, dtype=np.int64)
    assert L.or_mr_ghost_create_xyz(om.ctypes.data, M, len(om), oa.ctypes.data, N, len(oa), C.byref(osub),
                                    corr.ctypes.data, out.ctypes.data) == 0
    mg, ag = int(out[0]), int(out[1])
    assert (mols.numGhostMolecules, atoms.numGhostAtoms) == (mg, ag)
    assert np.array_equal(atoms.getPos()[:N + ag], oa["pos"][:N + ag])
    assert np.array_equal(mols.get("atomsOffset")[:M + mg], om["atomsOffset"][:M + mg])
    assert np.array_equal(ghost.correspondingRealAtom(N + ag), corr[:N + ag])
    oc, on = oracle.verlet_build(om, 13, M + mg, 0, M, rc + skin, 1.0, np.array(osub.minGhostCorner),
                                 np.array(osub.maxGhostCorner), half=True, width=40)
    gc, gn = vl.to_host()
    assert np.array_equal(gc[:M], oc[:M])
    arrs = [np.ascontiguousarray(x) for x in (capv, rcv, sig, eps)]
    oh = L.or_adress_create(*[x.ctypes.data for x in arrs], nt, 1)

    for run in range(2):
        api.UpdateMolecules.update(mols, atoms, w)
        L.or_update_molecules(om.ctypes.data, M + mg, oa.ctypes.data, C.byref(ow))
        assert np.abs(mols.get("pos")[:M + mg] - om["pos"][:M + mg]).max() < 1e-13
        assert np.abs(mols.get("modulatedLambda")[:M + mg] - om["modLambda"][:M + mg]).max() < 1e-14
        assert np.abs(mols.get("gradLambda")[:M + mg] - om["gradLambda"][:M + mg]).max() < 1e-13
        # identical weights on both sides so that region decisions (and histogram bins) cannot flip on an ulp
        mols.set("lambda", om["lambda"][:M + mg])
        mols.set("modulatedLambda", om["modLambda"][:M + mg])
        mols.set("gradLambda", om["gradLambda"][:M + mg])
        atoms.setForce(0.0)
        mols.setForce(0.0)
        oa["force"] = 0.0
        om["force"] = 0.0
        e = lj.run(mols, vl, atoms)
        nact = C.c_int64()
        oe = L.or_adress_run(oh, om.ctypes.data, M, oc.ctypes.data, on.ctypes.data, on.shape[1], oa.ctypes.data,
                             C.byref(nact))
        assert lj.lastNumPairs == nact.value and nact.value > 0
        assert abs(e - oe) <= 1e-12 * abs(oe)
        fa, fm = atoms.getForce()[:N + ag], mols.get("force")[:M + mg]
        assert np.abs(fa - oa["force"][:N + ag]).max() <= FORCE_RTOL * np.abs(oa["force"]).max()
        assert np.abs(fm - om["force"][:M + mg]).max() <= FORCE_RTOL * max(np.abs(om["force"]).max(), 1e-30)
        mean = lj.getMeanCompensationEnergy()
        omean = np.ctypeslib.as_array(oh.contents.meanCompensationEnergy, shape=(200, nt))
        assert np.count_nonzero(omean) > 0
        assert np.abs(mean - omean).max() <= 1e-12 * np.abs(omean).max()
        api.ContributeMoleculeForceToAtoms.update(mols, atoms)
        L.or_contribute_molecule_force(om.ctypes.data, M + mg, oa.ctypes.data)
        ghost.contributeBackGhostToReal(atoms)
        L.or_ghost_fold_force(oa.ctypes.data, N, ag, corr.ctypes.data)
        assert np.abs(atoms.getForce()[:N] - oa["force"][:N]).max() <= FORCE_RTOL * np.abs(oa["force"][:N]).max()
    L.or_adress_destroy(oh)


# --- thermodynamic force ---------------------------------------------------------------------------
def thermo_fixture(api):
    sub = api.Subdomain([0, 0, 0], [10, 1, 1], 1.0)
    t = api.ThermodynamicForce([1.0], sub, 1.0, [1.0])
    assert t.numBins == 10
    t.setForce(np.arange(10.0).reshape(10, 1))
    return t


def atoms_uniform(api):
    return api.Atoms.from_arrays(np.stack([np.arange(100) / 10.0, np.zeros(100), np.zeros(100)], axis=1))


def atoms_nonuniform(api):
    xs = [i + 0.5 for i in range(10) for _ in range(i)]
    return api.Atoms.from_arrays(np.stack([np.array(xs), np.zeros(45), np.zeros(45)], axis=1))


def test_thermodynamic_force_kats(api):
    """mrmd/action/ThermodynamicForce.test.cpp:96-245"""
    t = thermo_fixture(api)
    t.sample(atoms_uniform(api))
    assert np.all(t.getDensityProfile(0) == 10.0) and t.getNumberOfDensityProfileSamples() == 1
    t = thermo_fixture(api)
    t.sample(atoms_nonuniform(api))
    assert np.all(t.getDensityProfile(0) == np.arange(10.0))

    t = thermo_fixture(api)
    a = atoms_uniform(api)
    t.apply(a)
    assert np.all(a.getForce()[:, 0] == np.floor(a.getPos()[:, 0])) and np.all(a.getForce()[:, 1:] == 0)
    pred = api.interval_pred(0.5, 4.5)
    for interp in (False, True):
        a = atoms_uniform(api)
        (t.applyInterpolated_if if interp else t.apply_if)(a, pred)
        x = a.getPos()[:, 0]
        inside = (x > 0.5) & (x < 4.5)
        want = np.where(inside, x - 0.5 if interp else np.floor(x), 0.0)
        assert np.allclose(a.getForce()[:, 0], want, rtol=0, atol=1e-12)

    t = thermo_fixture(api)
    t.sample(atoms_nonuniform(api))
    t.update(1.0, 0.0)
    assert np.allclose(t.getForce(0), np.arange(10.0) - 1.0, atol=1e-13)
    assert t.getNumberOfDensityProfileSamples() == 0 and np.all(t.getDensityProfile() == 0)
    t = thermo_fixture(api)
    t.sample(atoms_nonuniform(api))
    t.update_if(1.0, 0.0, api.interval_pred(0.51, 4.49))
    grid = np.arange(10) + 0.5
    want = np.where((grid > 0.51) & (grid < 4.49), np.arange(10.0) - 1.0, np.arange(10.0))
    assert np.allclose(t.getForce(0), want, atol=1e-13)
    assert abs(t.getMuLeft()[0] - want[:5].sum()) < 1e-12 and abs(t.getMuRight()[0] - want[5:].sum()) < 1e-12


@pytest.mark.parametrize("symmetric,periodic", [(False, False), (True, True)])
def test_thermodynamic_force_vs_oracle(api, oracle, symmetric, periodic):
    rng = np.random.default_rng(4)
    n = 200000
    box = np.array([50.0, 8.0, 8.0])
    pos = rng.random((n, 3)) * box
    pos[:, 0] = np.clip(pos[:, 0] + 3.0 * np.sin(pos[:, 0] * 0.4), 0, 49.999999)
    types = (rng.random(n) < 0.3).astype(np.int64)
    sub = api.Subdomain([0, 0, 0], box, 1.0)
    osub = oracle.subdomain([0, 0, 0], box, 1.0)
    td, mod = np.array([0.5, 0.2]), np.array([2.0, 1.5])
    t = api.ThermodynamicForce(td, sub, 0.25, mod, symmetric, periodic)
    ot = oracle.lib().or_thermo_create(td.ctypes.data, 2, C.byref(osub), 0.25, mod.ctypes.data, int(symmetric),
                                       int(periodic))
    nb = t.numBins
    assert nb == ot.contents.numBins == 200
    atoms = api.Atoms.from_arrays(pos, type=types)
    oa = np.zeros(n, dtype=oracle.ATOM)
    oa["pos"], oa["type"] = pos, types
    for _ in range(3):
        t.sample(atoms)
        oracle.lib().or_thermo_sample(ot, oa.ctypes.data, n)
    od = np.ctypeslib.as_array(ot.contents.density, shape=(nb, 2))
    assert np.array_equal(t.getDensityProfile(), od)  # histogram counts are exact
    t.update(2.0, 2.0)
    oracle.lib().or_thermo_update(ot, 2.0, 2.0, None)
    of = np.ctypeslib.as_array(ot.contents.force, shape=(nb, 2))
    assert np.abs(t.getForce() - of).max() <= 1e-12 * np.abs(of).max()
    t.setForce(of)  # identical tables for the apply comparison
    pred = api.IsInSymmetricSlab([25.0, 0, 0], 5.0, 15.0)
    op = oracle.make_pred(oracle.PRED_SLAB, 0, 25.0, 5.0, 15.0)
    for interp in (0, 1):
        atoms.setForce(0.0)
        oa["force"] = 0.0
        (t.applyInterpolated_if if interp else t.apply_if)(atoms, pred)
        oracle.lib().or_thermo_apply(ot, oa.ctypes.data, n, C.byref(op), interp)
        f = atoms.getForce()
        assert np.abs(f - oa["force"]).max() <= 1e-13 * np.abs(oa["force"]).max()
        assert np.count_nonzero(f[:, 0]) > 0 and np.all(f[:, 1:] == 0)
    oracle.lib().or_thermo_destroy(ot)


def test_atoms_per_molecule_promise_is_checked(api):
    """a molecule that breaks mrmd_b200_adress_set_atoms_per_molecule(4) makes the run fail loudly"""
    sites, apos, box, M = adress_system(6, 4, 3)
    atoms = api.Atoms.from_arrays(apos, None, mass=1.0, relativeMass=0.25)
    mols = api.Molecules(M)
    na = np.full(M, 4)
    na[M // 2] = 3
    mols.set("atomsOffset", np.arange(M) * 4)
    mols.set("numAtoms", na)
    mols.numLocalMolecules = M
    w = api.Slab([box / 2] * 3, 4.0, 3.0, 2)
    api.UpdateMolecules.update(mols, atoms, w)
    vl = api.HalfVerletList()
    vl.build(mols, 0, M, 2.6, 1.0, [-0.5] * 3, [box + 0.5] * 3, 40)
    lj = api.LJ_IdealGas(0.7, 2.5, 1.0, 1.0, True)
    lj.setAtomsPerMolecule(4)
    with pytest.raises(RuntimeError, match="atom count promised"):
        lj.run(mols, vl, atoms)
