"""ctypes binding of the CPU parity oracle (TEST INFRASTRUCTURE -- see oracle/mrmd_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package mrmd_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmrmd_oracle.so")

ATOM = np.dtype(
    [("pos", "f8", 3), ("vel", "f8", 3), ("force", "f8", 3), ("type", "i8"), ("mass", "f8"), ("charge", "f8"),
     ("relMass", "f8")]
)
MOLECULE = np.dtype(
    [("pos", "f8", 3), ("force", "f8", 3), ("lambda", "f8"), ("modLambda", "f8"), ("gradLambda", "f8", 3),
     ("atomsOffset", "i8"), ("numAtoms", "i8")]
)
assert ATOM.itemsize == 104 and MOLECULE.itemsize == 104


class Subdomain(C.Structure):
    _fields_ = [(n, C.c_double * 3) for n in (
        "minCorner", "maxCorner", "ghostLayerThickness", "minGhostCorner", "maxGhostCorner", "minInnerCorner",
        "maxInnerCorner", "diameter", "diameterWithGhostLayer")]


class LJType(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "ff1", "ff2", "ef1", "ef2", "rcSqr", "cappingDistance", "cappingDistanceSqr", "cappingCoeff", "shift",
        "energyAtCappingPoint")]


class Pred(C.Structure):
    _fields_ = [("kind", C.c_int32), ("axis", C.c_int32), ("center", C.c_double), ("slabMin", C.c_double),
                ("slabMax", C.c_double), ("tolerance", C.c_double)]


PRED_ALWAYS, PRED_NEVER, PRED_SLAB, PRED_SLAB_EITHER, PRED_SLAB_BOTH, PRED_INTERVAL = range(6)


class Weight(C.Structure):
    _fields_ = [("kind", C.c_int32), ("abrupt", C.c_int32), ("center", C.c_double * 3), ("atRegion", C.c_double),
                ("hyRegion", C.c_double), ("exponent", C.c_int64)]


WEIGHT_SLAB, WEIGHT_SPHERICAL = 0, 1


class Adress(C.Structure):
    _fields_ = [("numTypes", C.c_int64), ("rcSqr", C.c_double), ("numBins", C.c_int64), ("runCounter", C.c_int64),
                ("samplingInterval", C.c_int64), ("updateInterval", C.c_int64), ("table", C.POINTER(LJType)),
                ("compensationEnergy", C.POINTER(C.c_double)), ("compensationEnergyCounter", C.POINTER(C.c_double)),
                ("meanCompensationEnergy", C.POINTER(C.c_double))]


class Thermo(C.Structure):
    _fields_ = [("min", C.c_double), ("max", C.c_double), ("numBins", C.c_int64), ("numTypes", C.c_int64),
                ("binSize", C.c_double), ("inverseBinSize", C.c_double), ("binVolume", C.c_double),
                ("samples", C.c_int64), ("enforceSymmetry", C.c_int), ("usePeriodicity", C.c_int),
                ("force", C.POINTER(C.c_double)), ("density", C.POINTER(C.c_double)),
                ("forceFactor", C.POINTER(C.c_double))]


def build(force=False):
    """Compile oracle/libmrmd_oracle.so (building the checker is not using it)."""
    src = os.path.join(_HERE, "mrmd_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmrmd_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    dp = C.POINTER(C.c_double)
    vp = C.c_void_p
    i64 = C.c_int64
    sigs = {
        "or_set_threads": (None, [C.c_int]),
        "or_get_max_threads": (C.c_int, []),
        "or_subdomain_init": (None, [C.POINTER(Subdomain), vp, vp, vp]),
        "or_subdomain_scale_dim": (None, [C.POINTER(Subdomain), C.c_double, C.c_int]),
        "or_lj_init": (None, [vp, vp, vp, vp, vp, i64, C.c_int]),
        "or_lj_force_energy": (None, [vp, i64, C.c_double, dp, dp]),
        "or_lj_apply": (i64, [vp, i64, vp, vp, i64, vp, C.c_double, i64, C.POINTER(Pred), vp]),
        "or_cell_ids": (i64, [vp, i64, i64, i64, vp, vp, vp, vp, vp]),
        "or_cell_perm": (None, [vp, i64, i64, i64, vp, vp]),
        "or_permute_atoms": (None, [vp, i64, i64, vp]),
        "or_permute_molecules": (None, [vp, i64, i64, vp]),
        "or_verlet_build": (i64, [vp, i64, i64, i64, i64, C.c_double, C.c_double, vp, vp, C.c_int, i64, vp, vp]),
        "or_count_within_cutoff": (i64, [vp, i64, i64, i64, C.c_double, vp, C.c_int]),
        "or_periodic_map": (None, [vp, i64, C.POINTER(Subdomain)]),
        "or_ghost_create_axis": (i64, [vp, i64, i64, i64, C.POINTER(Subdomain), C.c_int, vp]),
        "or_ghost_create_xyz": (i64, [vp, i64, i64, C.POINTER(Subdomain), vp]),
        "or_ghost_update_pos": (None, [vp, i64, i64, vp, C.POINTER(Subdomain)]),
        "or_ghost_fold_force": (None, [vp, i64, i64, vp]),
        "or_zero_force": (None, [vp, i64]),
        "or_mr_periodic_map": (None, [vp, i64, vp, C.POINTER(Subdomain)]),
        "or_mr_ghost_create_axis": (C.c_int, [vp, i64, i64, i64, vp, i64, i64, i64, C.POINTER(Subdomain), C.c_int,
                                              vp, vp]),
        "or_mr_ghost_create_xyz": (C.c_int, [vp, i64, i64, vp, i64, i64, C.POINTER(Subdomain), vp, vp]),
        "or_vv_pre": (C.c_double, [vp, i64, C.c_double]),
        "or_vv_post": (None, [vp, i64, C.c_double]),
        "or_langevin_pre": (C.c_double, [vp, i64, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_uint64,
                                         C.POINTER(Pred)]),
        "or_langevin_pre_ids": (C.c_double, [vp, i64, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_uint64,
                                             C.POINTER(Pred), vp]),
        "or_philox_normals": (None, [C.c_uint64, C.c_uint64, C.c_uint64, vp]),
        "or_philox4x32": (None, [vp, vp, vp]),
        "or_weight_eval": (None, [C.POINTER(Weight), C.c_double, C.c_double, C.c_double, dp, dp, vp]),
        "or_update_molecules": (None, [vp, i64, vp, C.POINTER(Weight)]),
        "or_contribute_molecule_force": (None, [vp, i64, vp]),
        "or_adress_create": (C.POINTER(Adress), [vp, vp, vp, vp, i64, C.c_int]),
        "or_adress_destroy": (None, [C.POINTER(Adress)]),
        "or_adress_run": (C.c_double, [C.POINTER(Adress), vp, i64, vp, vp, i64, vp, C.POINTER(i64)]),
        "or_hist_get_bin": (i64, [C.c_double, C.c_double, i64, C.c_double]),
        "or_hist_scale": (None, [vp, i64, i64, C.c_double]),
        "or_hist_scale_per_hist": (None, [vp, i64, i64, vp]),
        "or_hist_make_symmetric": (None, [vp, i64, i64]),
        "or_hist_gradient": (None, [vp, vp, C.c_double, C.c_double, i64, i64, C.c_int]),
        "or_hist_smoothen": (None, [vp, vp, C.c_double, C.c_double, i64, i64, C.c_double, C.c_double, C.c_int]),
        "or_limit_acceleration": (None, [vp, i64, C.c_double]),
        "or_limit_velocity": (None, [vp, i64, C.c_double]),
        "or_berendsen_thermostat": (None, [vp, i64, C.c_double, C.c_double, C.c_double]),
        "or_berendsen_barostat": (None, [vp, i64, C.c_double, C.c_double, C.c_double, C.POINTER(Subdomain), C.c_int,
                                        C.c_int, C.c_int]),
        "or_shake_positional": (C.c_int, [vp, i64, vp, i64, vp, vp, i64, i64, C.c_double]),
        "or_shake_velocity": (C.c_int, [vp, i64, vp, vp, i64]),
        "or_approx_erfc": (C.c_double, [C.c_double]),
        "or_coulomb_eval": (None, [C.c_int, C.c_double, C.c_double, vp, i64, C.c_double, C.c_double, vp, vp]),
        "or_spc_apply_forces": (None, [vp, i64, vp, vp, i64, vp, C.c_int, vp]),
        "or_spc_bond_energy": (C.c_double, [vp, i64, vp, i64, C.c_double]),
        "or_kinetic_energy": (C.c_double, [vp, i64]),
        "or_system_momentum": (None, [vp, i64, vp]),
        "or_pressure": (C.c_double, [vp, i64, vp]),
        "or_msd": (C.c_double, [vp, vp, i64, vp]),
        "or_density_profile": (None, [vp, i64, i64, C.c_double, C.c_double, i64, C.c_int, vp]),
        "or_thermo_create": (C.POINTER(Thermo), [vp, i64, C.POINTER(Subdomain), C.c_double, vp, C.c_int, C.c_int]),
        "or_thermo_destroy": (None, [C.POINTER(Thermo)]),
        "or_thermo_sample": (None, [C.POINTER(Thermo), vp, i64]),
        "or_thermo_update": (None, [C.POINTER(Thermo), C.c_double, C.c_double, C.POINTER(Pred)]),
        "or_thermo_apply": (None, [C.POINTER(Thermo), vp, i64, C.POINTER(Pred), C.c_int]),
        "or_thermo_mu": (None, [C.POINTER(Thermo), vp, vp]),
    }
    for name, (res, args) in sigs.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def ptr(a):
    """Raw data pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle arrays must be C-contiguous"
    return a.ctypes.data


def d3(v):
    return np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (3,)))


def subdomain(min_corner, max_corner, thickness):
    s = Subdomain()
    a, b, c = d3(min_corner), d3(max_corner), d3(thickness)
    lib().or_subdomain_init(C.byref(s), ptr(a), ptr(b), ptr(c))
    return s


def make_pred(kind=PRED_ALWAYS, axis=0, center=0.0, slab_min=0.0, slab_max=0.0, tol=0.0):
    return Pred(kind, axis, center, slab_min, slab_max, tol)


def make_weight(kind, center, at_region, hy_region, exponent, abrupt=False):
    w = Weight()
    w.kind = kind
    w.abrupt = int(abrupt)
    for d in range(3):
        w.center[d] = float(center[d])
    w.atRegion = at_region
    w.hyRegion = hy_region
    w.exponent = exponent
    return w


def lj_table(cap, rc, sigma, eps, num_types=1, shifted=False):
    n = num_types * num_types
    arrs = [np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64).ravel(), (n,))) for v in
            (cap, rc, sigma, eps)]
    table = (LJType * n)()
    lib().or_lj_init(C.addressof(table), *[ptr(a) for a in arrs], num_types, int(shifted))
    return table


def verlet_build(pos_arr, stride, n_all, begin, end, radius, ratio, gmin, gmax, half=True, width=64):
    """Returns (counts[n_all], neigh[n_all,width]); widens and refills on overflow like Cabana."""
    gmin, gmax = d3(gmin), d3(gmax)
    while True:
        counts = np.zeros(n_all, dtype=np.int32)
        neigh = np.full((n_all, max(width, 1)), -1, dtype=np.int32)
        mx = lib().or_verlet_build(ptr(pos_arr), stride, n_all, begin, end, radius, ratio, ptr(gmin), ptr(gmax),
                                   int(half), neigh.shape[1], ptr(counts), ptr(neigh))
        if mx <= neigh.shape[1]:
            return counts, neigh
        width = int(mx)


def pair_set(counts, neigh, n_rows):
    """Sorted (i, j) pair array of a 2-D neighbour table."""
    rows = np.repeat(np.arange(n_rows, dtype=np.int64), counts[:n_rows])
    mask = np.arange(neigh.shape[1])[None, :] < counts[:n_rows, None]
    cols = neigh[:n_rows][mask].astype(np.int64)
    pairs = np.stack([rows, cols], axis=1)
    order = np.lexsort((pairs[:, 1], pairs[:, 0]))
    return pairs[order]
