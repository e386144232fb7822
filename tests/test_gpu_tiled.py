"""The B200 fast path (mrmd_b200_verlet_build_periodic + tiled LJ force, mrmd_b200/csrc/tiled.cu) against the
oracle: the decoded pair set must equal the reference's list over local + ghost atoms mapped through
correspondingRealAtom (bit exact), forces / energy / virial within the FP64 tolerances."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FORCE_RTOL = 1e-10
SCALAR_RTOL = 1e-12


@pytest.fixture(scope="module")
def api():
    from mrmd_b200 import api as a

    assert a.L().mrmd_b200_device_count() > 0
    return a


def system(n_side, jitter, seed, spacing=1.25, box_scale=(1, 1, 1)):
    rng = np.random.default_rng(seed)
    dims = [n_side * s for s in box_scale]
    g = np.stack(np.meshgrid(*[np.arange(d) for d in dims], indexing="ij"), axis=-1).reshape(-1, 3)
    pos = (g + 0.5) * spacing + (rng.random(g.shape) - 0.5) * jitter
    box = np.array(dims) * spacing
    return np.mod(pos, box), rng.random(g.shape) - 0.5, box


def reference_pairs(oracle, pos_sorted, box, thickness, radius, half, width=128):
    """oracle: ghosts + Cabana list, partners mapped to (real atom, image shift code)"""
    L = oracle.lib()
    n = len(pos_sorted)
    sub = oracle.subdomain([0, 0, 0], box, thickness)
    oa = np.zeros(8 * n + 64, dtype=oracle.ATOM)
    oa["pos"][:n] = pos_sorted
    oa["mass"][:n] = 1.0
    corr = np.zeros(len(oa), dtype=np.int64)
    ng = L.or_ghost_create_xyz(oa.ctypes.data, n, len(oa), C.byref(sub), corr.ctypes.data)
    assert ng >= 0
    counts, neigh = oracle.verlet_build(oa, 13, n + ng, 0, n, radius, 1.0, np.array(sub.minGhostCorner),
                                        np.array(sub.maxGhostCorner), half=half, width=width)
    rows = []
    for i in range(n):
        nb = neigh[i, :counts[i]].astype(np.int64)
        real = np.where(nb < n, nb, corr[nb])
        shift = np.rint((oa["pos"][nb] - oa["pos"][real]) / box).astype(np.int64)
        code = (shift[:, 0] + 1) + 3 * (shift[:, 1] + 1) + 9 * (shift[:, 2] + 1)
        rows.append(sorted(zip(real.tolist(), code.tolist())))
    return oa, corr, ng, counts, neigh, rows, sub


@pytest.mark.parametrize("n_side,jitter,thickness,half,zfine", [
    (12, 0.7, 2.6, False, 1), (12, 0.7, 2.6, True, 1), (16, 0.0, 2.6, False, 1), (10, 1.2, [0.0, 2.6, 2.6], False, 1),
    (9, 0.9, 2.6, False, 1), (12, 0.7, 2.6, False, 4), (12, 0.7, 2.6, True, 4), (16, 0.0, 2.6, False, 4),
    (10, 1.2, [2.6, 2.6, 0.0], False, 4), (9, 0.9, 2.6, False, 2.5), (20, 1.0, 2.6, False, 4)])
def test_periodic_list_equals_ghost_list(api, oracle, n_side, jitter, thickness, half, zfine):
    """zfine > 1: LinkedCellList gridDelta (r, r, r / zfine) -- the drivers' sort grid; the builder then scans only the
    z interval of each column that the cutoff sphere reaches"""
    pos, vel, box = system(n_side, jitter, 100 + n_side)
    n = len(pos)
    radius = 2.6
    sub = api.Subdomain([0, 0, 0], box, thickness)
    atoms = api.Atoms.from_arrays(pos, vel)
    api.GhostLayer().exchangeRealAtoms(atoms, sub)
    atoms.permute(api.LinkedCellList(0, n, [radius, radius, radius / zfine], sub.minCorner, sub.maxCorner))
    sorted_pos = atoms.getPos()[:n]
    vl = api.HalfVerletList() if half else api.FullVerletList()
    vl.build_periodic(atoms, sub, radius, 1.0, 30 if half else 50)
    gc, partner, code = vl.to_host_periodic(atoms)
    oa, corr, ng, oc, on, rows, osub = reference_pairs(oracle, sorted_pos, box, thickness, radius, half)
    assert np.array_equal(gc, oc[:n])
    assert vl.info()["totalPairs"] == int(oc[:n].sum())
    for i in range(n):
        got = sorted(zip(partner[i, :gc[i]].tolist(), code[i, :gc[i]].tolist()))
        assert got == rows[i], i


def test_dense_region_tiles_take_several_passes(api, oracle):
    """A third of the box holds all atoms at almost four times the mean density: the tiles of that region hold more home
    atoms than the builder has lanes (several passes per tile), and the rows are wider than 64 entries (the words of a row
    behind the prefetched ones in the builder's copy-out and in the force kernel).  Pair sets bit-exact, forces 1e-10."""
    rng = np.random.default_rng(77)
    box = np.array([20.8, 20.8, 83.2])
    n = 23040
    pos = rng.random((n, 3)) * np.array([box[0], box[1], box[2] / 3.0])
    # keep pairs apart (the force law is capped, but the comparison is relative to the largest force)
    pos = pos[np.argsort(pos[:, 2])]
    vel = np.zeros_like(pos)
    radius, rc, cap = 2.6, 2.5, 0.7
    sub = api.Subdomain([0, 0, 0], box, radius)
    atoms = api.Atoms.from_arrays(pos, vel)
    api.GhostLayer().exchangeRealAtoms(atoms, sub)
    atoms.permute(api.LinkedCellList(0, n, [radius, radius, radius / 4], sub.minCorner, sub.maxCorner))
    sorted_pos = atoms.getPos()[:n]
    vl = api.FullVerletList()
    vl.build_periodic(atoms, sub, radius, 1.0, 250)
    gc, partner, code = vl.to_host_periodic(atoms)
    oa, corr, ng, oc, on, rows, osub = reference_pairs(oracle, sorted_pos, box, radius, radius, False, width=320)
    assert gc.max() > 128 and np.array_equal(gc, oc[:n])
    for i in range(n):
        got = sorted(zip(partner[i, :gc[i]].tolist(), code[i, :gc[i]].tolist()))
        assert got == rows[i], i
    lj = api.LennardJones(rc, 1.0, 1.0, cap)
    atoms.setForce(0.0)
    lj.apply(atoms, vl)
    # the reference path: half list over local + ghost atoms, ghost forces folded back
    oa, corr, ng, oc, on, rows, osub = reference_pairs(oracle, sorted_pos, box, radius, radius, True, width=320)
    table = oracle.lj_table(cap, rc, 1.0, 1.0)
    ev = np.zeros(2)
    pairs = oracle.lib().or_lj_apply(oa.ctypes.data, n, oc.ctypes.data, on.ctypes.data, on.shape[1],
                                     C.addressof(table), rc * rc, 1, None, ev.ctypes.data)
    oracle.lib().or_ghost_fold_force(oa.ctypes.data, n, ng, corr.ctypes.data)
    f, f_ref = atoms.getForce()[:n], oa["force"][:n]
    assert np.abs(f - f_ref).max() <= 1e-10 * np.abs(f_ref).max()
    e, v, p = lj._get()
    assert p == pairs and abs(e - ev[0]) <= 1e-11 * abs(ev[0])


@pytest.mark.parametrize("n_side,jitter,box_scale", [(12, 0.7, (1, 1, 1)), (20, 0.0, (1, 1, 1)), (8, 0.9, (3, 1, 2))])
def test_tiled_force_vs_oracle(api, oracle, n_side, jitter, box_scale):
    pos, vel, box = system(n_side, jitter, 7 + n_side, box_scale=box_scale)
    n = len(pos)
    rc, skin, cap = 2.5, 0.1, 0.7
    sub = api.Subdomain([0, 0, 0], box, rc + skin)
    atoms = api.Atoms.from_arrays(pos, vel)
    ghost = api.GhostLayer()
    ghost.exchangeRealAtoms(atoms, sub)
    atoms.permute(api.LinkedCellList(0, n, [rc + skin] * 3, sub.minCorner, sub.maxCorner))
    sorted_pos = atoms.getPos()[:n]
    ghost.createGhostAtoms(atoms, sub)
    vl = api.FullVerletList()
    vl.build_periodic(atoms, sub, rc + skin, 1.0, 60)
    lj = api.LennardJones(rc, 1.0, 1.0, cap)
    atoms.setForce(0.0)
    lj.apply(atoms, vl)

    oa, corr, ng, oc, on, rows, osub = reference_pairs(oracle, sorted_pos, box, rc + skin, rc + skin, True)
    assert atoms.numGhostAtoms == ng
    table = oracle.lj_table(cap, rc, 1.0, 1.0)
    ev = np.zeros(2)
    pairs = oracle.lib().or_lj_apply(oa.ctypes.data, n, oc.ctypes.data, on.ctypes.data, on.shape[1],
                                     C.addressof(table), rc * rc, 1, None, ev.ctypes.data)
    oracle.lib().or_ghost_fold_force(oa.ctypes.data, n, ng, corr.ctypes.data)
    f = atoms.getForce()
    scale = np.abs(oa["force"][:n]).max() if jitter > 0 else 1.0
    assert np.abs(f[:n] - oa["force"][:n]).max() <= FORCE_RTOL * scale
    assert np.all(f[n:n + ng] == 0.0)
    e, v, p = lj._get()
    assert p == pairs
    assert abs(e - ev[0]) <= SCALAR_RTOL * abs(ev[0]) and abs(v - ev[1]) <= SCALAR_RTOL * abs(ev[1])
    # a second apply accumulates (reference semantics of LennardJones::apply)
    lj.apply(atoms, vl)
    assert np.abs(atoms.getForce()[:n] - 2 * oa["force"][:n]).max() <= 2 * FORCE_RTOL * scale


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_md_driver_vs_oracle_loop(api, oracle, golden_dir, mode):
    """mrmd_b200_md_run (C++ step loop) in all three list modes against the oracle's loop on the shipped
    4096-atom configuration, Langevin integrator, 40 steps: same rebuild count, ghost count, energies and
    trajectories within a divergence bound."""
    from oracle.md_loop import OracleMD

    g = np.load(f"{golden_dir}/lj_nvt_final.npz")
    n = len(g["pos"])
    sub = api.Subdomain([0, 0, 0], g["box"], 2.6)
    atoms = api.Atoms.from_arrays(g["pos"], g["vel"])
    md = api.MolecularDynamics(atoms, sub, langevin=True, zeta=20.0, temperature=1.5, seed=1234, cellSort=True,
                               fullList=mode)
    omd = OracleMD(g["pos"], g["vel"], g["box"], langevin=True, zeta=20.0, temperature=1.5, seed=1234, cell_sort=True)
    steps = 40
    st = md.run(steps)
    res = omd.run(steps)
    assert st["rebuilds"] == res["rebuilds"]
    assert st["pairInteractions"] == res["pairInteractions"]
    assert abs(st["energy"] - res["energy"]) <= 1e-8 * abs(res["energy"])
    assert np.abs(atoms.getPos()[:n] - omd.atoms["pos"][:n]).max() < 1e-9
    assert np.abs(atoms.getVel()[:n] - omd.atoms["vel"][:n]).max() < 1e-8
    assert np.abs(atoms.getForce()[:n] - omd.atoms["force"][:n]).max() < 1e-7 * np.abs(omd.atoms["force"][:n]).max()
    if mode == 2:
        # tiled fast path: periodic images are generated while staging, no ghost atom is ever materialised
        assert st["numGhost"] == 0 and atoms.numGhostAtoms == 0
        return
    # the container is left as the reference loop leaves it: ghosts at their images, no force on ghosts
    ng = omd.ng
    assert st["numGhost"] == ng
    assert np.abs(atoms.getPos()[n:n + ng] - omd.atoms["pos"][n:n + ng]).max() < 1e-9
    assert np.all(atoms.getForce()[n:n + ng] == 0.0)


def test_md_host_buffer_path(api, golden_dir):
    """mrmd_b200_md_run_host (host buffers in, host buffers out) reproduces the device-resident run."""
    g = np.load(f"{golden_dir}/lj_nvt_final.npz")
    n = len(g["pos"])
    sub = api.Subdomain([0, 0, 0], g["box"], 2.6)
    a1, a2 = api.Atoms.from_arrays(g["pos"], g["vel"]), api.Atoms.from_arrays(g["pos"], g["vel"])
    md1 = api.MolecularDynamics(a1, sub, langevin=True, fullList=2)
    md2 = api.MolecularDynamics(a2, sub, langevin=True, fullList=2)
    md1.run(25)
    hp, hv, hs = api.PinnedBuffer((n, 3)), api.PinnedBuffer((n, 3)), api.PinnedBuffer((3,))
    hp.array[:], hv.array[:] = g["pos"], g["vel"]
    md2.run_host(25, hp.ptr, hv.ptr, hs.ptr)
    assert np.array_equal(hp.array, a1.getPos()[:n]) and np.array_equal(hv.array, a1.getVel()[:n])
    assert hs.array[0] < 0 and np.isfinite(hs.array).all()
