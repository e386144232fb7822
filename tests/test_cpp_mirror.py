"""include/mrmd/: the C++20 mirror of the reference API.

CPU: both example drivers compile and link against the C ABI with -Wall -Wextra -Werror, every reference header
path a hot-path driver includes exists, and without a GPU the binary aborts with the library's message.
GPU: the NVE driver written against the mirror reproduces the oracle's run of the same loop."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include", "mrmd")
LIBDIR = os.path.join(ROOT, "mrmd_b200")

REFERENCE_HEADERS = [
    "datatypes.hpp", "data/Atoms.hpp", "data/Molecules.hpp", "data/MoleculesFromAtoms.hpp", "data/Subdomain.hpp",
    "action/LennardJones.hpp", "action/LJ_IdealGas.hpp", "action/ThermodynamicForce.hpp", "action/UpdateMolecules.hpp",
    "action/ContributeMoleculeForceToAtoms.hpp", "action/VelocityVerlet.hpp",
    "action/VelocityVerletLangevinThermostat.hpp", "communication/GhostLayer.hpp",
    "communication/MultiResGhostLayer.hpp", "util/IsInSymmetricSlab.hpp", "weighting_function/Slab.hpp",
    "weighting_function/Spherical.hpp", "weighting_function/CheckRegion.hpp",
    "analysis/KineticEnergy.hpp", "analysis/SystemMomentum.hpp", "analysis/Pressure.hpp",
    "analysis/MeanSquareDisplacement.hpp", "io/RestoreGRO.hpp", "io/DumpGRO.hpp", "io/RestoreTXT.hpp",
    "io/DumpThermoForce.hpp", "io/RestoreThermoForce.hpp", "action/BerendsenThermostat.hpp",
    "action/BerendsenBarostat.hpp", "action/Shake.hpp", "data/Bond.hpp", "action/SPC.hpp", "action/Coulomb.hpp",
    "action/CoulombDSF.hpp", "io/DumpCSV.hpp", "action/LimitAcceleration.hpp", "action/LimitVelocity.hpp",
    "util/ExponentialMovingAverage.hpp", "Cabana_NeighborList.hpp", "data/MultiHistogram.hpp",
]


def _compile(name, out_dir):
    from mrmd_b200 import build

    build.build()  # makes sure libmrmd_b200.so exists
    exe = os.path.join(str(out_dir), name)
    cmd = ["/usr/bin/g++", "-std=c++20", "-O2", "-Wall", "-Wextra", "-Werror", "-I" + INC,
           os.path.join(ROOT, "examples", name + ".cpp"), "-L" + LIBDIR, "-lmrmd_b200", "-Wl,-rpath," + LIBDIR, "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-4000:]
    return exe


def test_reference_header_paths_exist():
    for h in REFERENCE_HEADERS:
        assert os.path.exists(os.path.join(INC, h)), h


@pytest.mark.parametrize("name", ["lennard_jones_nve", "adress_ideal_gas", "restart_io", "constraints_step", "spc_water",
                                  "tetramer_adress", "multi_histogram"])
def test_example_compiles_and_fails_loudly_without_gpu(name, tmp_path):
    import torch

    exe = _compile(name, tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    args = ["4", "1"] if name != "restart_io" else [str(tmp_path / "a.gro"), str(tmp_path / "b.gro"), str(tmp_path / "t"), "1"]
    res = subprocess.run([exe] + args, capture_output=True, text=True, timeout=60)
    assert res.returncode != 0
    assert "no CUDA device" in res.stderr


def _lcg_system(sites, spacing=1.25):
    """the start configuration of examples/lennard_jones_nve.cpp (48-bit LCG, same draw order)"""
    s = 0x1234ABCD330E
    n = sites ** 3
    pos = np.zeros((n, 3))
    vel = np.zeros((n, 3))
    idx = 0
    for i in range(sites):
        for j in range(sites):
            for k in range(sites):
                cell = (i, j, k)
                for d in range(3):
                    s = (s * 0x5DEECE66D + 0xB) & ((1 << 48) - 1)
                    pos[idx, d] = (cell[d] + 0.5) * spacing + (s / float(1 << 48) - 0.5) * 0.4
                for d in range(3):
                    s = (s * 0x5DEECE66D + 0xB) & ((1 << 48) - 1)
                    vel[idx, d] = s / float(1 << 48) - 0.5
                idx += 1
    return pos, vel


@pytest.mark.gpu
def test_nve_driver_matches_oracle(tmp_path):
    from oracle.md_loop import OracleMD

    sites, steps = 12, 60
    exe = _compile("lennard_jones_nve", tmp_path)
    res = subprocess.run([exe, str(sites), str(steps)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])

    pos, vel = _lcg_system(sites)
    box = np.full(3, sites * 1.25)
    md = OracleMD(pos, vel, box, langevin=False, cell_sort=False)
    st = md.run(steps)
    assert out["atoms"] == sites ** 3
    assert out["rebuilds"] == st["rebuilds"]
    assert out["ghosts"] == md.ng
    assert out["pairs"] == int(md.counts.sum())
    # HalfNeighborList::numNeighbor / getNeighbor (host accessors of the mirror) on the first atom's row
    assert out["neighbors0"] == int(md.counts[0]) and out["neighborSum0"] == int(md.neigh[0, :md.counts[0]].sum())
    assert out["seconds"] > 0
    # NVE trajectories through the same neighbour lists: summation order is the only difference
    assert abs(out["E0"] - st["energy"]) <= 1e-9 * abs(st["energy"])
    v = md.atoms["vel"][:md.n]
    ek = 0.5 * float((v * v).sum())
    assert abs(out["Ek"] - ek) <= 1e-9 * ek
    assert np.allclose(out["x0"], md.atoms["pos"][0], rtol=0, atol=1e-9)
    # the statistics line through the analysis:: mirror
    import ctypes as C

    from oracle import pyoracle as orc

    L, n = orc.lib(), md.n
    assert abs(out["T"] - (2.0 / 3.0) * ek / n) <= 1e-9 * ek / n
    p = L.or_pressure(md.atoms.ctypes.data, n + md.ng, C.byref(md.sub))
    assert abs(out["p"] - p) <= 1e-8 * abs(p)
    init = np.ascontiguousarray(pos)
    msd = L.or_msd(md.atoms.ctypes.data, init.ctypes.data, n, C.byref(md.sub))
    assert abs(out["msd"] - msd) <= 1e-7 * msd
    assert np.allclose(out["momentum"], v.sum(axis=0), rtol=0, atol=1e-9)


def _write_gro(path, pos, vel, box):
    """a .gro in the reference's format (mrmd/io/DumpGRO.cpp:26-89 with velocities)"""
    with open(path, "w") as f:
        f.write("golden, t=0\n%d\n" % len(pos))
        for i, (p, v) in enumerate(zip(pos, vel)):
            f.write("%5d%-5s%5s%5d%8.3f%8.3f%8.3f%8.4f%8.4f%8.4f\n" % (i + 1, "Argon", "Ar", i + 1, *p, *v))
        f.write("    %g %g %g\n" % tuple(box))


def _read_gro(path):
    lines = open(path).read().splitlines()
    n = int(lines[1])
    rows = np.array([[float(ln[20 + 8 * k:28 + 8 * k]) for k in range(6)] for ln in lines[2:2 + n]])
    return rows[:, :3], rows[:, 3:], np.array([float(x) for x in lines[2 + n].split()])


@pytest.mark.gpu
def test_gro_restart_and_thermo_force_files(tmp_path, golden_dir):
    """io:: mirror (row (f)3): a reference-format .gro is restored, continued for 20 NVE steps and dumped; the dump equals
    the oracle's continuation of the same (rounded) file contents to the file's precision; the thermodynamic-force profile
    survives dumpThermoForce / restoreThermoForce (mrmd/io/GRO.test.cpp, ThermoForce.test.cpp)"""
    from oracle.md_loop import OracleMD

    g = np.load(f"{golden_dir}/lj_nvt_final.npz")
    gin, gout, tf = str(tmp_path / "in.gro"), str(tmp_path / "out.gro"), str(tmp_path / "tf.txt")
    _write_gro(gin, g["pos"], g["vel"], g["box"])
    pos, vel, box = _read_gro(gin)  # what the file holds after rounding to %8.3f / %8.4f
    steps = 20
    exe = _compile("restart_io", tmp_path)
    res = subprocess.run([exe, gin, gout, tf, str(steps)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert out["atoms"] == len(pos) and np.allclose(out["box"], box)
    assert abs(out["sumPos"] - pos.sum()) < 1e-9 * abs(pos.sum()) and abs(out["sumVel"] - vel.sum()) < 1e-9
    md = OracleMD(pos, vel, box, langevin=False, cell_sort=False)
    st = md.run(steps)
    assert abs(out["E0"] - st["energy"]) <= 1e-9 * abs(st["energy"])
    p2, v2, b2 = _read_gro(gout)
    assert np.allclose(b2, box)
    assert np.abs(p2 - md.atoms["pos"][:md.n]).max() <= 0.5e-3 + 1e-9
    assert np.abs(v2 - md.atoms["vel"][:md.n]).max() <= 0.5e-4 + 1e-9
    csv = open(gout + ".csv").read().splitlines()  # io::dumpCSV (mrmd/io/DumpCSV.cpp:26-52), local atoms only
    assert csv[0] == "idx, mol, type, ghost, pos_x, pos_y, pos_z, vel_x, vel_y, vel_z" and len(csv) == len(pos) + 1
    row = [float(x) for x in csv[6].split(",")]
    assert row[:4] == [5.0, 1.0, 0.0, 0.0] and np.allclose(row[4:7], md.atoms["pos"][5], rtol=1e-5, atol=1e-5)
    assert np.allclose(row[7:], md.atoms["vel"][5], rtol=1e-5, atol=1e-5)
    assert out["tfBins"] == out["tfBinsRestored"] == 100
    assert out["maxForceDiff"] <= 1e-3 and out["maxGridDiff"] <= 1e-5  # default ostream precision: 6 significant digits
    lines = open(tf).read().splitlines()
    assert len(lines) == 3 and len(lines[0].split()) == 100


@pytest.mark.gpu
def test_constrained_step(tmp_path):
    """tests/Constraints/Constraints.cpp:25-72 of the reference through the mirror: after SHAKE + velocity Verlet +
    RATTLE the bond has its length and no relative velocity along it; Berendsen calls as in tests/NVT, tests/NPT"""
    exe = _compile("constraints_step", tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert abs(out["dist"] - 1.0) < 1e-6  # EXPECT_FLOAT_EQ(calcDist(0, 1), 1_r)
    assert abs(out["relVel"]) < 1e-6      # EXPECT_FLOAT_EQ(calcRelVel(0, 1) + 1_r, 1_r)
    assert out["T0"] > 0 and abs(out["T1"] - 2.0 * out["T0"]) < 1e-12 * out["T0"]
    assert np.isfinite(out["p"]) and abs(out["maxCorner"] - 2.0 * 2.0 ** (1.0 / 3.0)) < 1e-12


@pytest.mark.gpu
def test_adress_driver_runs(tmp_path):
    sites, steps = 12, 120
    exe = _compile("adress_ideal_gas", tmp_path)
    res = subprocess.run([exe, str(sites), str(steps)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert out["atoms"] == sites ** 3 and out["rebuilds"] >= 1
    assert out["numAT"] > 0 and out["numHY"] > 0 and out["numAT"] + out["numHY"] < out["atoms"]
    assert out["densitySamples"] == len(range(110, steps, 10))  # samples since the update at step 100
    assert np.isfinite(out["E"]) and np.isfinite(out["muLeft"])


def _lcg_water(sites, spacing=0.31, jitter=0.02):
    """the start configuration of examples/spc_water.cpp (48-bit LCG, same draw order)"""
    state = [0x1234ABCD330E]

    def rnd():
        state[0] = (state[0] * 0x5DEECE66D + 0xB) & ((1 << 48) - 1)
        return state[0] / float(1 << 48)

    eq, ang = 0.1, 109.47 / 180.0 * 3.14159265358979323846
    m = sites ** 3
    pos, vel = np.zeros((3 * m, 3)), np.zeros((3 * m, 3))
    idx = 0
    for i in range(sites):
        for j in range(sites):
            for k in range(sites):
                o = 3 * idx
                for d, c in enumerate((i, j, k)):
                    pos[o, d] = (c + 0.5) * spacing + (rnd() - 0.5) * jitter
                phi = rnd() * 2.0 * np.pi
                for h, a in enumerate((phi, phi + ang)):
                    pos[o + 1 + h] = pos[o] + (eq * np.cos(a), eq * np.sin(a), 0.0)
                for d in range(3):
                    vel[o:o + 3, d] = (rnd() - 0.5) * 0.5
                idx += 1
    mass = np.tile([15.999, 1.008, 1.008], m)
    return pos, vel, mass, np.tile([-0.82, 0.41, 0.41], m), mass / (15.999 + 2 * 1.008), np.tile([0, 1, 1], m), np.full(3, sites * spacing)


@pytest.mark.gpu
def test_spc_water_driver_matches_oracle(tmp_path):
    """row (f)4 through the mirror: constrained SPC water, 30 steps with rebuilds, against the oracle's run"""
    from oracle.md_loop import OracleSpcMD

    sites, steps = 10, 30
    exe = _compile("spc_water", tmp_path)
    res = subprocess.run([exe, str(sites), str(steps)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    md = OracleSpcMD(*_lcg_water(sites), dt=0.0005, skin=0.02)
    st = md.run(steps)
    assert out["molecules"] == sites ** 3 and out["rebuilds"] == st["rebuilds"] >= 2
    assert out["ghostAtoms"] == md.ng and out["ghostMolecules"] == md.mg
    assert abs(out["ELJ"] - st["energyLJ"]) <= 1e-8 * abs(st["energyLJ"])
    assert abs(out["ECoulomb"] - st["energyCoulomb"]) <= 1e-8 * abs(st["energyCoulomb"])
    assert np.allclose(out["x0"], md.atoms["pos"][0], rtol=0, atol=1e-9)
    assert out["maxBondError"] < 1e-5 and 0 <= out["bondEnergy"] < 1e-6


def _lcg_tetramers(sites, spacing=1.98425):
    """the start configuration of examples/tetramer_adress.cpp (48-bit LCG, same draw order)"""
    state = [0x1234ABCD330E]

    def rnd():
        state[0] = (state[0] * 0x5DEECE66D + 0xB) & ((1 << 48) - 1)
        return state[0] / float(1 << 48)

    a = 1.0 / (2.0 * np.sqrt(2.0))
    tet = np.array([(a, a, a), (a, -a, -a), (-a, a, -a), (-a, -a, a)])
    m = sites ** 3
    pos, vel = np.zeros((4 * m, 3)), np.zeros((4 * m, 3))
    idx = 0
    for i in range(sites):
        for j in range(sites):
            for k in range(sites):
                v = [rnd() - 0.5 for _ in range(3)]
                pos[4 * idx:4 * idx + 4] = (np.array([i, j, k]) + 0.5) * spacing + tet
                vel[4 * idx:4 * idx + 4] = v
                idx += 1
    return pos, vel, np.full(3, sites * spacing)


@pytest.mark.gpu
def test_tetramer_driver_matches_oracle(tmp_path):
    """BASELINE.json configs[3] through the mirror: AdResS tetramers, spherical region, Langevin, SHAKE / RATTLE; 60 steps
    against the oracle's loop of the same operators"""
    from oracle import pyoracle as orc
    from oracle.md_loop import OracleAdressMD

    sites, steps = 10, 60
    exe = _compile("tetramer_adress", tmp_path)
    res = subprocess.run([exe, str(sites), str(steps)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    pos, vel, box = _lcg_tetramers(sites)
    w = orc.make_weight(orc.WEIGHT_SPHERICAL, 0.5 * box, 60.0 / 317.48 * box[0], 30.0 / 317.48 * box[0], 2)
    md = OracleAdressMD(pos, vel, box, w, langevin=True, zeta=20.0, temperature=1.5, seed=1234, max_neigh=40,
                        atoms_per_mol=4, constraint_iterations=3, bond_length=1.0)
    st = md.run(steps)
    assert out["atoms"] == 4 * sites ** 3 and out["rebuilds"] == st["rebuilds"] >= 2
    assert out["ghostAtoms"] == md.ng
    assert abs(out["E"] - st["energy"]) <= 1e-8 * abs(st["energy"])
    assert np.allclose(out["x0"], md.atoms["pos"][0], rtol=0, atol=1e-9)
    assert np.allclose(out["v0"], md.atoms["vel"][0], rtol=0, atol=1e-8)
    assert out["maxBondError"] < 5e-3


@pytest.mark.gpu
def test_multi_histogram_reference_tests_through_the_mirror(tmp_path):
    """row a26: every TEST of mrmd/data/MultiHistogram.test.cpp (getBin, getBinPosition, createGrid, scale, make_symmetric,
    gradient, the four operators, smoothen, replace_if_bin_position) plus cumulativeMovingAverage, the copy constructor and
    ThermodynamicForce::getForce() returning a data::MultiHistogram, written against include/mrmd/data/MultiHistogram.hpp"""
    exe = _compile("multi_histogram", tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert res.returncode == 0 and out["failures"] == 0, (out, res.stderr[-1000:])
