"""The reference's step loop driven through the CPU oracle (TEST INFRASTRUCTURE / CPU baseline only).

Same loop as mrmd_b200/csrc/md.cu: examples/02_LennardJones_NVE.cpp:135-216 with the Langevin integrator of
examples/01 and the LinkedCellList + permute of tests/NVT/NVT.cpp:136-144 at every rebuild.  Used by
bench.py (cpu_baseline leg and --impl reference) and by tests/; never by the product.
"""
import ctypes as C
import time

import numpy as np

from . import pyoracle as orc


class OracleMD:
    def __init__(self, pos, vel, box, dt=0.002, rc=2.5, skin=0.1, sigma=1.0, epsilon=1.0, cap=0.7, max_neigh=60,
                 langevin=False, zeta=20.0, temperature=1.5, seed=1234, cell_sort=True, ghost_capacity_factor=None,
                 ids=None):
        self.L = orc.lib()
        self.n = n = len(pos)
        self.box = np.asarray(box, dtype=np.float64)
        self.cutoff = rc + skin
        self.sub = orc.subdomain([0, 0, 0], self.box, self.cutoff)
        frac = float(np.prod(self.box + 2 * self.cutoff) / np.prod(self.box)) - 1.0
        cap_atoms = int(n * (1.0 + (ghost_capacity_factor or (1.3 * frac + 0.05)))) + 1024
        self.atoms = np.zeros(cap_atoms, dtype=orc.ATOM)
        self.atoms["pos"][:n] = pos
        self.atoms["vel"][:n] = vel
        self.atoms["mass"][:n] = 1.0
        self.atoms["relMass"][:n] = 1.0
        self.corr = np.full(cap_atoms, -1, dtype=np.int64)
        self.table = orc.lj_table(cap, rc, sigma, epsilon)
        self.rc, self.skin, self.dt = rc, skin, dt
        self.langevin, self.zeta, self.temperature, self.seed = langevin, zeta, temperature, seed
        self.cell_sort, self.max_neigh = cell_sort, max_neigh
        self.max_disp = np.finfo(np.float64).max
        self.step = 0
        self.ng = 0
        self.counts = self.neigh = None
        self.rebuilds = 0
        self.pairs = 0
        self.energy_virial = np.zeros(2)
        self._cid = np.zeros(n, dtype=np.int32)
        self._perm = np.zeros(n, dtype=np.int64)
        # global atom ids: the Philox counter of the Langevin noise, permuted with the records
        self.ids = np.arange(n, dtype=np.int64) if ids is None else np.ascontiguousarray(ids, dtype=np.int64).copy()

    def _rebuild(self):
        L, a, n = self.L, self.atoms, self.n
        L.or_periodic_map(a.ctypes.data, n, C.byref(self.sub))
        if self.cell_sort:
            delta = np.array([self.cutoff, self.cutoff, 0.25 * self.cutoff])  # the driver's LinkedCellList gridDelta
            lo, hi = np.zeros(3), self.box
            nc = L.or_cell_ids(a.ctypes.data, 13, 0, n, delta.ctypes.data, lo.ctypes.data, hi.ctypes.data,
                               self._cid.ctypes.data, None)
            off = np.zeros(nc + 1, dtype=np.int64)
            L.or_cell_perm(self._cid.ctypes.data, 0, n, nc, self._perm.ctypes.data, off.ctypes.data)
            L.or_permute_atoms(a.ctypes.data, 0, n, self._perm.ctypes.data)
            self.ids = self.ids[self._perm]
        self.ng = L.or_ghost_create_xyz(a.ctypes.data, n, len(a), C.byref(self.sub), self.corr.ctypes.data)
        assert self.ng >= 0, "oracle ghost capacity exceeded"
        self.counts, self.neigh = orc.verlet_build(a, 13, n + self.ng, 0, n, self.cutoff, 1.0,
                                                   np.array(self.sub.minGhostCorner), np.array(self.sub.maxGhostCorner),
                                                   half=True, width=self.max_neigh)
        self.rebuilds += 1

    def one_step(self):
        L, a, n = self.L, self.atoms, self.n
        if self.langevin:
            d = L.or_langevin_pre_ids(a.ctypes.data, n, self.dt, self.zeta, self.temperature, self.seed, self.step, None,
                                      self.ids.ctypes.data)
        else:
            d = L.or_vv_pre(a.ctypes.data, n, self.dt)
        self.max_disp += d
        if self.max_disp >= self.skin * 0.5:
            self.max_disp = 0.0
            self._rebuild()
        else:
            L.or_ghost_update_pos(a.ctypes.data, n, self.ng, self.corr.ctypes.data, C.byref(self.sub))
        L.or_zero_force(a.ctypes.data, n + self.ng)
        self.pairs += L.or_lj_apply(a.ctypes.data, n, self.counts.ctypes.data, self.neigh.ctypes.data,
                                    self.neigh.shape[1], C.addressof(self.table), self.rc * self.rc, 1, None,
                                    self.energy_virial.ctypes.data)
        L.or_ghost_fold_force(a.ctypes.data, n, self.ng, self.corr.ctypes.data)
        L.or_vv_post(a.ctypes.data, n, self.dt)
        self.step += 1

    def run(self, nsteps):
        t0 = time.perf_counter()
        p0, r0 = self.pairs, self.rebuilds
        for _ in range(nsteps):
            self.one_step()
        dt = time.perf_counter() - t0
        return {"seconds": dt, "steps": nsteps, "pairInteractions": self.pairs - p0, "rebuilds": self.rebuilds - r0,
                "energy": float(self.energy_virial[0])}


class OracleAdressMD:
    """The AdResS step of SURVEY.md section 3.5 (one molecule per atom) driven through the CPU oracle, the loop of
    mrmd_b200/csrc/md.cu in AdResS mode: UpdateMolecules -> LJ_IdealGas -> ThermodynamicForce ->
    ContributeMoleculeForceToAtoms -> MultiResGhostLayer, Langevin integrator, spatial sort at every rebuild."""

    def __init__(self, pos, vel, box, weight, dt=0.002, rc=2.5, skin=0.1, sigma=1.0, epsilon=1.0, cap=0.7,
                 max_neigh=60, langevin=True, zeta=20.0, temperature=1.5, seed=1234, thermo=None, do_shift=True,
                 atoms_per_mol=1, constraint_iterations=0, bond_length=1.0):
        self.L = orc.lib()
        self.n = n = len(pos)
        self.apm = apm = atoms_per_mol
        self.nm = nm = n // apm
        self.box = np.asarray(box, dtype=np.float64)
        self.cutoff = rc + skin
        self.sub = orc.subdomain([0, 0, 0], self.box, self.cutoff)
        frac = float(np.prod(self.box + 2 * self.cutoff) / np.prod(self.box)) - 1.0
        cap_atoms = int(n * (1.0 + 1.3 * frac + 0.05)) + 1024
        self.atoms = np.zeros(cap_atoms, dtype=orc.ATOM)
        self.atoms["pos"][:n], self.atoms["vel"][:n] = pos, vel
        self.atoms["mass"][:n], self.atoms["relMass"][:n] = 1.0, 1.0 / apm
        self.mols = np.zeros(cap_atoms // apm + 1, dtype=orc.MOLECULE)
        self.mols["atomsOffset"][:nm], self.mols["numAtoms"][:nm] = np.arange(nm) * apm, apm
        self.corr = np.full(cap_atoms, -1, dtype=np.int64)
        # MoleculeConstraints(apm, constraint_iterations) with a bond between every two atoms of a molecule
        self.constraint_iterations = constraint_iterations
        pairs = [(i, j) for i in range(apm) for j in range(i + 1, apm)]
        self.bond_idx = np.ascontiguousarray(pairs, dtype=np.int64).reshape(-1)
        self.bond_eq = np.full(len(pairs), float(bond_length))
        self.weight = weight
        one = [np.array([float(v)]) for v in (cap, rc, sigma, epsilon)]
        self._keep = one
        self.adress = self.L.or_adress_create(*[x.ctypes.data for x in one], 1, int(do_shift))
        self.thermo, self.thermo_cfg = None, thermo
        if thermo is not None:
            td, tm = np.array([thermo["targetDensity"]]), np.array([thermo["modulation"]])
            self._keep += [td, tm]
            self.thermo = self.L.or_thermo_create(td.ctypes.data, 1, C.byref(self.sub), thermo["binWidth"], tm.ctypes.data, 0, 0)
        self.rc, self.skin, self.dt = rc, skin, dt
        self.langevin, self.zeta, self.temperature, self.seed, self.max_neigh = langevin, zeta, temperature, seed, max_neigh
        self.max_disp = np.finfo(np.float64).max
        self.step = self.ng = self.mg = self.rebuilds = self.pairs = 0
        self.energy = 0.0
        self._cid, self._perm = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int64)
        self.ids = np.arange(n, dtype=np.int64)

    def _rebuild(self):
        L, a, m, n, nm = self.L, self.atoms, self.mols, self.n, self.nm
        L.or_update_molecules(m.ctypes.data, nm, a.ctypes.data, C.byref(self.weight))
        L.or_mr_periodic_map(m.ctypes.data, nm, a.ctypes.data, C.byref(self.sub))
        if self.apm == 1:
            delta = np.array([self.cutoff, self.cutoff, 0.25 * self.cutoff])
            lo, hi = np.zeros(3), self.box
            nc = L.or_cell_ids(a.ctypes.data, 13, 0, n, delta.ctypes.data, lo.ctypes.data, hi.ctypes.data,
                               self._cid.ctypes.data, None)
            off = np.zeros(nc + 1, dtype=np.int64)
            L.or_cell_perm(self._cid.ctypes.data, 0, n, nc, self._perm.ctypes.data, off.ctypes.data)
            L.or_permute_atoms(a.ctypes.data, 0, n, self._perm.ctypes.data)  # one atom per molecule: offsets stay i -> i
            self.ids = self.ids[self._perm]
            L.or_update_molecules(m.ctypes.data, nm, a.ctypes.data, C.byref(self.weight))
        out = np.zeros(2, dtype=np.int64)
        rc = L.or_mr_ghost_create_xyz(m.ctypes.data, nm, len(m), a.ctypes.data, n, len(a), C.byref(self.sub),
                                      self.corr.ctypes.data, out.ctypes.data)
        assert rc == 0, "oracle ghost capacity exceeded"
        self.mg, self.ng = int(out[0]), int(out[1])
        L.or_update_molecules(m.ctypes.data, nm + self.mg, a.ctypes.data, C.byref(self.weight))
        self.counts, self.neigh = orc.verlet_build(m, 13, nm + self.mg, 0, nm, self.cutoff, 1.0,
                                                   np.array(self.sub.minGhostCorner), np.array(self.sub.maxGhostCorner),
                                                   half=True, width=self.max_neigh)
        self.rebuilds += 1

    def one_step(self):
        L, a, m, n, nm = self.L, self.atoms, self.mols, self.n, self.nm
        if self.constraint_iterations > 0:
            assert L.or_shake_positional(m.ctypes.data, nm, a.ctypes.data, n + self.ng, self.bond_idx.ctypes.data,
                                         self.bond_eq.ctypes.data, len(self.bond_eq), self.constraint_iterations, self.dt) == 0
        if self.langevin:
            d = L.or_langevin_pre_ids(a.ctypes.data, n, self.dt, self.zeta, self.temperature, self.seed, self.step, None,
                                      self.ids.ctypes.data)
        else:
            d = L.or_vv_pre(a.ctypes.data, n, self.dt)
        self.max_disp += d
        if self.max_disp >= self.skin * 0.5:
            self.max_disp = 0.0
            self._rebuild()
        else:
            L.or_ghost_update_pos(a.ctypes.data, n, self.ng, self.corr.ctypes.data, C.byref(self.sub))
            L.or_update_molecules(m.ctypes.data, nm + self.mg, a.ctypes.data, C.byref(self.weight))
        L.or_zero_force(a.ctypes.data, n + self.ng)
        m["force"][:nm + self.mg] = 0.0
        nact = C.c_int64()
        self.energy = L.or_adress_run(self.adress, m.ctypes.data, nm, self.counts.ctypes.data, self.neigh.ctypes.data,
                                      self.neigh.shape[1], a.ctypes.data, C.byref(nact))
        self.pairs += nact.value
        if self.thermo is not None:
            t = self.thermo_cfg
            if self.step % t["sampleInterval"] == 0:
                L.or_thermo_sample(self.thermo, a.ctypes.data, n)
            if self.step > 0 and self.step % t["updateInterval"] == 0 and self.thermo.contents.samples > 0:
                L.or_thermo_update(self.thermo, t["sigma"], t["range"], None)
            L.or_thermo_apply(self.thermo, a.ctypes.data, n, None, 0)
        L.or_contribute_molecule_force(m.ctypes.data, nm + self.mg, a.ctypes.data)
        L.or_ghost_fold_force(a.ctypes.data, n, self.ng, self.corr.ctypes.data)
        L.or_vv_post(a.ctypes.data, n, self.dt)
        if self.constraint_iterations > 0:
            assert L.or_shake_velocity(m.ctypes.data, nm, a.ctypes.data, self.bond_idx.ctypes.data, len(self.bond_eq)) == 0
        self.step += 1

    def run(self, nsteps):
        t0 = time.perf_counter()
        p0, r0 = self.pairs, self.rebuilds
        for _ in range(nsteps):
            self.one_step()
        dt = time.perf_counter() - t0
        return {"seconds": dt, "steps": nsteps, "pairInteractions": self.pairs - p0, "rebuilds": self.rebuilds - r0,
                "energy": float(self.energy)}


def spc_water_box(sites, spacing=0.31, jitter=0.02, seed=5):
    """sites^3 SPC molecules in their equilibrium geometry (mrmd/action/SPC.test.cpp:51-96) on a jittered lattice, each
    turned about z by a random angle; returns pos[3M,3], vel[3M,3], mass, charge, relMass, type, box"""
    rng = np.random.default_rng(seed)
    m = sites ** 3
    eq_ho, angle = 0.1, 109.47 / 180.0 * np.pi
    g = (np.stack(np.meshgrid(*[np.arange(sites)] * 3, indexing="ij"), axis=-1).reshape(-1, 3) + 0.5) * spacing
    o = g + (rng.random((m, 3)) - 0.5) * jitter
    phi = rng.random(m) * 2 * np.pi
    h0 = o + eq_ho * np.stack([np.cos(phi), np.sin(phi), np.zeros(m)], axis=1)
    h1 = o + eq_ho * np.stack([np.cos(phi + angle), np.sin(phi + angle), np.zeros(m)], axis=1)
    pos = np.stack([o, h0, h1], axis=1).reshape(-1, 3)
    vel = np.repeat((rng.random((m, 3)) - 0.5) * 0.5, 3, axis=0)  # rigid translation: no velocity along the bonds
    mass = np.tile([15.999, 1.008, 1.008], m)
    charge = np.tile([-0.82, 0.41, 0.41], m)
    typ = np.tile([0, 1, 1], m).astype(np.int64)
    return pos, vel, mass, charge, mass / (15.999 + 2 * 1.008), typ, np.full(3, sites * spacing)


class OracleSpcMD:
    """A constrained SPC water step driven through the CPU oracle: SHAKE -> velocity Verlet -> MultiResGhostLayer ->
    UpdateMolecules -> SPC::applyForces -> RATTLE, the call order of the reference's tests/Constraints/Constraints.cpp:
    25-72 around the AdResS-style molecule loop.  Three atoms per molecule, no spatial sort."""

    SPC_RC = 1.2

    def __init__(self, pos, vel, mass, charge, rel_mass, typ, box, dt=0.0005, skin=0.1, max_neigh=220, coulomb_kind=0):
        self.L = orc.lib()
        self.n = n = len(pos)
        self.nm = nm = n // 3
        self.box = np.asarray(box, dtype=np.float64)
        self.cutoff = self.SPC_RC + skin
        self.sub = orc.subdomain([0, 0, 0], self.box, self.cutoff)
        frac = float(np.prod(self.box + 2 * self.cutoff) / np.prod(self.box)) - 1.0
        cap_mols = int(nm * (1.0 + 1.3 * frac + 0.05)) + 1024
        self.atoms = np.zeros(3 * cap_mols, dtype=orc.ATOM)
        a = self.atoms
        a["pos"][:n], a["vel"][:n], a["mass"][:n], a["charge"][:n] = pos, vel, mass, charge
        a["relMass"][:n], a["type"][:n] = rel_mass, typ
        self.mols = np.zeros(cap_mols, dtype=orc.MOLECULE)
        self.mols["atomsOffset"][:nm], self.mols["numAtoms"][:nm] = np.arange(nm) * 3, 3
        self.corr = np.full(3 * cap_mols, -1, dtype=np.int64)
        self.weight = orc.make_weight(orc.WEIGHT_SLAB, 0.5 * self.box, 10.0 * float(self.box[0]), 1.0, 7)  # lambda = 1
        eq_ho, angle = 0.1, 109.47 / 180.0 * np.pi
        self.bond_idx = np.array([0, 1, 0, 2, 1, 2], dtype=np.int64)
        self.bond_eq = np.array([eq_ho, eq_ho, eq_ho * np.sqrt(2.0 - 2.0 * np.cos(angle))])
        self.dt, self.skin, self.max_neigh, self.kind = dt, skin, max_neigh, coulomb_kind
        self.max_disp = np.finfo(np.float64).max
        self.step = self.ng = self.mg = self.rebuilds = 0
        self.energies = np.zeros(2)

    def _rebuild(self):
        L, a, m, n, nm = self.L, self.atoms, self.mols, self.n, self.nm
        L.or_update_molecules(m.ctypes.data, nm, a.ctypes.data, C.byref(self.weight))
        L.or_mr_periodic_map(m.ctypes.data, nm, a.ctypes.data, C.byref(self.sub))
        out = np.zeros(2, dtype=np.int64)
        rc = L.or_mr_ghost_create_xyz(m.ctypes.data, nm, len(m), a.ctypes.data, n, len(a), C.byref(self.sub),
                                      self.corr.ctypes.data, out.ctypes.data)
        assert rc == 0, "oracle ghost capacity exceeded"
        self.mg, self.ng = int(out[0]), int(out[1])
        L.or_update_molecules(m.ctypes.data, nm + self.mg, a.ctypes.data, C.byref(self.weight))
        self.counts, self.neigh = orc.verlet_build(m, 13, nm + self.mg, 0, nm, self.cutoff, 1.0,
                                                   np.array(self.sub.minGhostCorner), np.array(self.sub.maxGhostCorner),
                                                   half=True, width=self.max_neigh)
        self.rebuilds += 1

    def one_step(self):
        L, a, m, n, nm = self.L, self.atoms, self.mols, self.n, self.nm
        assert L.or_shake_positional(m.ctypes.data, nm, a.ctypes.data, n + self.ng, self.bond_idx.ctypes.data,
                                     self.bond_eq.ctypes.data, 3, 20, self.dt) == 0
        self.max_disp += L.or_vv_pre(a.ctypes.data, n, self.dt)
        if self.max_disp >= self.skin * 0.5:
            self.max_disp = 0.0
            self._rebuild()
        else:
            L.or_ghost_update_pos(a.ctypes.data, n, self.ng, self.corr.ctypes.data, C.byref(self.sub))
            L.or_update_molecules(m.ctypes.data, nm + self.mg, a.ctypes.data, C.byref(self.weight))
        L.or_zero_force(a.ctypes.data, n + self.ng)
        L.or_spc_apply_forces(m.ctypes.data, nm, self.counts.ctypes.data, self.neigh.ctypes.data, self.neigh.shape[1],
                              a.ctypes.data, self.kind, self.energies.ctypes.data)
        L.or_ghost_fold_force(a.ctypes.data, n, self.ng, self.corr.ctypes.data)
        L.or_vv_post(a.ctypes.data, n, self.dt)
        assert L.or_shake_velocity(m.ctypes.data, nm, a.ctypes.data, self.bond_idx.ctypes.data, 3) == 0
        self.step += 1

    def run(self, nsteps):
        t0 = time.perf_counter()
        for _ in range(nsteps):
            self.one_step()
        return {"seconds": time.perf_counter() - t0, "steps": nsteps, "rebuilds": self.rebuilds,
                "energyLJ": float(self.energies[0]), "energyCoulomb": float(self.energies[1])}
