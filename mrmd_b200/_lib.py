"""ctypes loader for libmrmd_b200.so (the C ABI declared in include/mrmd_b200.h).

There is no CPU fallback: if the library is missing the import fails, and every compute entry point
returns MRMD_B200_ENODEVICE when no CUDA device is present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmrmd_b200.so")
# measurement only (profiles/variants.py): a side-by-side build of the same sources with other tuning constants
if os.environ.get("MRMD_B200_LIB_VARIANT"):
    LIB_PATH = os.path.join(_HERE, "variants", f"libmrmd_b200_{os.environ['MRMD_B200_LIB_VARIANT']}.so")


class Subdomain(C.Structure):
    _fields_ = [(n, C.c_double * 3) for n in (
        "minCorner", "maxCorner", "ghostLayerThickness", "minGhostCorner", "maxGhostCorner", "minInnerCorner",
        "maxInnerCorner", "diameter", "diameterWithGhostLayer")]


class Pred(C.Structure):
    _fields_ = [("kind", C.c_int32), ("axis", C.c_int32), ("center", C.c_double), ("slabMin", C.c_double),
                ("slabMax", C.c_double), ("tolerance", C.c_double)]


class Weight(C.Structure):
    _fields_ = [("kind", C.c_int32), ("abrupt", C.c_int32), ("center", C.c_double * 3), ("atRegion", C.c_double),
                ("hyRegion", C.c_double), ("exponent", C.c_int64)]


class MdConfig(C.Structure):
    _fields_ = [("dt", C.c_double), ("rc", C.c_double), ("skin", C.c_double), ("sigma", C.c_double),
                ("epsilon", C.c_double), ("cappingDistance", C.c_double), ("maxNeighbors", C.c_int64),
                ("integrator", C.c_int32), ("cellSort", C.c_int32), ("fullList", C.c_int32), ("adress", C.c_int32),
                ("zeta", C.c_double), ("temperature", C.c_double), ("seed", C.c_uint64), ("weight", Weight),
                ("doShift", C.c_int32), ("useThermoForce", C.c_int32), ("thermoTargetDensity", C.c_double),
                ("thermoBinWidth", C.c_double), ("thermoModulation", C.c_double),
                ("thermoSampleInterval", C.c_int64), ("thermoUpdateInterval", C.c_int64),
                ("thermoSmoothingSigma", C.c_double), ("thermoSmoothingIntensity", C.c_double),
                ("atomsPerMolecule", C.c_int64), ("numConstraintIterations", C.c_int64), ("bondLength", C.c_double),
                ("energyEveryStep", C.c_int32), ("reserved0", C.c_int32)]


class MdStats(C.Structure):
    _fields_ = [("steps", C.c_int64), ("rebuilds", C.c_int64), ("pairInteractions", C.c_int64),
                ("storedPairs", C.c_int64), ("numLocal", C.c_int64), ("numGhost", C.c_int64), ("energy", C.c_double),
                ("virial", C.c_double), ("forceKernelMs", C.c_double), ("maxDisplacement", C.c_double), ("activePairs", C.c_int64)]


vp = C.c_void_p
i64 = C.c_int64
dbl = C.c_double
pi64 = C.POINTER(C.c_int64)
pdbl = C.POINTER(C.c_double)
pint = C.POINTER(C.c_int)
pvp = C.POINTER(C.c_void_p)
pSub = C.POINTER(Subdomain)
pPred = C.POINTER(Pred)
pWeight = C.POINTER(Weight)

# name -> (restype, argtypes); mirrors include/mrmd_b200.h one to one
SIGNATURES = {
    "mrmd_b200_last_error": (C.c_char_p, []),
    "mrmd_b200_device_count": (C.c_int, []),
    "mrmd_b200_set_device": (C.c_int, [C.c_int]),
    "mrmd_b200_sync": (C.c_int, [vp]),
    "mrmd_b200_launch_count": (i64, []),
    "mrmd_b200_subdomain_init": (None, [pSub, vp, vp, vp]),
    "mrmd_b200_subdomain_scale_dim": (None, [pSub, dbl, C.c_int]),
    "mrmd_b200_atoms_create": (C.c_int, [pvp, i64]),
    "mrmd_b200_atoms_destroy": (C.c_int, [vp]),
    "mrmd_b200_atoms_resize": (C.c_int, [vp, i64, vp]),
    "mrmd_b200_atoms_reserve": (C.c_int, [vp, i64, vp]),
    "mrmd_b200_atoms_size": (i64, [vp]),
    "mrmd_b200_atoms_set_counts": (C.c_int, [vp, i64, i64]),
    "mrmd_b200_atoms_get_counts": (C.c_int, [vp, pi64, pi64]),
    "mrmd_b200_atoms_write": (C.c_int, [vp, C.c_int, vp, i64, i64, i64, i64, C.c_int, vp]),
    "mrmd_b200_atoms_read": (C.c_int, [vp, C.c_int, vp, i64, i64, i64, i64, C.c_int, vp]),
    "mrmd_b200_atoms_fill": (C.c_int, [vp, C.c_int, dbl, vp]),
    "mrmd_b200_atoms_copy": (C.c_int, [vp, vp, vp]),
    "mrmd_b200_molecules_create": (C.c_int, [pvp, i64]),
    "mrmd_b200_molecules_destroy": (C.c_int, [vp]),
    "mrmd_b200_molecules_resize": (C.c_int, [vp, i64, vp]),
    "mrmd_b200_molecules_size": (i64, [vp]),
    "mrmd_b200_molecules_set_counts": (C.c_int, [vp, i64, i64]),
    "mrmd_b200_molecules_get_counts": (C.c_int, [vp, pi64, pi64]),
    "mrmd_b200_molecules_write": (C.c_int, [vp, C.c_int, vp, i64, i64, i64, i64, C.c_int, vp]),
    "mrmd_b200_molecules_read": (C.c_int, [vp, C.c_int, vp, i64, i64, i64, i64, C.c_int, vp]),
    "mrmd_b200_molecules_fill": (C.c_int, [vp, C.c_int, dbl, vp]),
    "mrmd_b200_molecules_for_each_atom": (C.c_int, [pvp, vp, vp]),
    "mrmd_b200_vv_pre": (C.c_int, [vp, dbl, pdbl, vp]),
    "mrmd_b200_vv_post": (C.c_int, [vp, dbl, vp]),
    "mrmd_b200_langevin_pre": (C.c_int, [vp, dbl, dbl, dbl, C.c_uint64, C.c_uint64, pPred, pdbl, vp]),
    "mrmd_b200_ghost_create": (C.c_int, [pvp]),
    "mrmd_b200_ghost_destroy": (C.c_int, [vp]),
    "mrmd_b200_ghost_map_into_domain": (C.c_int, [vp, pSub, vp]),
    "mrmd_b200_ghost_create_atoms": (C.c_int, [vp, vp, pSub, C.c_int, vp]),
    "mrmd_b200_ghost_reset": (C.c_int, [vp, vp, vp]),
    "mrmd_b200_ghost_update": (C.c_int, [vp, vp, pSub, vp]),
    "mrmd_b200_ghost_contribute_back": (C.c_int, [vp, vp, vp]),
    "mrmd_b200_ghost_read_corresponding": (C.c_int, [vp, vp, i64, i64, C.c_int, vp]),
    "mrmd_b200_ghost_write_corresponding": (C.c_int, [vp, vp, i64, i64, C.c_int, vp]),
    "mrmd_b200_ghost_mr_map_into_domain": (C.c_int, [vp, vp, pSub, vp]),
    "mrmd_b200_ghost_mr_create_atoms": (C.c_int, [vp, vp, vp, pSub, C.c_int, vp]),
    "mrmd_b200_atoms_cell_sort": (C.c_int, [vp, i64, i64, vp, vp, vp, vp, vp]),
    "mrmd_b200_molecules_cell_sort": (C.c_int, [vp, i64, i64, vp, vp, vp, vp]),
    "mrmd_b200_verlet_create": (C.c_int, [pvp, C.c_int]),
    "mrmd_b200_verlet_destroy": (C.c_int, [vp]),
    "mrmd_b200_verlet_build_atoms": (C.c_int, [vp, vp, i64, i64, dbl, dbl, vp, vp, i64, vp]),
    "mrmd_b200_verlet_build_molecules": (C.c_int, [vp, vp, i64, i64, dbl, dbl, vp, vp, i64, vp]),
    "mrmd_b200_verlet_build_periodic": (C.c_int, [vp, vp, pSub, dbl, dbl, i64, vp]),
    "mrmd_b200_verlet_read_periodic": (C.c_int, [vp, vp, vp, vp, vp, vp]),
    "mrmd_b200_verlet_info": (C.c_int, [vp, pi64, pi64, pi64, pint]),
    "mrmd_b200_verlet_read": (C.c_int, [vp, vp, vp, C.c_int, vp]),
    "mrmd_b200_lj_create": (C.c_int, [pvp, vp, vp, vp, vp, i64, C.c_int]),
    "mrmd_b200_lj_destroy": (C.c_int, [vp]),
    "mrmd_b200_lj_apply": (C.c_int, [vp, vp, vp, pPred, vp]),
    "mrmd_b200_lj_get": (C.c_int, [vp, pdbl, pdbl, pi64, vp]),
    "mrmd_b200_lj_eval": (C.c_int, [vp, i64, vp, i64, vp, vp, vp]),
    "mrmd_b200_molecules_update": (C.c_int, [vp, vp, pWeight, vp]),
    "mrmd_b200_molecules_contribute_force": (C.c_int, [vp, vp, vp]),
    "mrmd_b200_weight_eval": (C.c_int, [pWeight, vp, i64, vp, vp, vp, vp]),
    "mrmd_b200_adress_create": (C.c_int, [pvp, vp, vp, vp, vp, i64, C.c_int]),
    "mrmd_b200_adress_destroy": (C.c_int, [vp]),
    "mrmd_b200_adress_set_intervals": (C.c_int, [vp, i64, i64]),
    "mrmd_b200_adress_set_atoms_per_molecule": (C.c_int, [vp, i64]),
    "mrmd_b200_adress_run": (C.c_int, [vp, vp, vp, vp, pdbl, pi64, vp]),
    "mrmd_b200_adress_run_periodic": (C.c_int, [vp, vp, vp, vp, pdbl, pi64, vp]),
    "mrmd_b200_adress_read_histogram": (C.c_int, [vp, C.c_int, vp, vp]),
    "mrmd_b200_thermo_create": (C.c_int, [pvp, vp, i64, pSub, dbl, vp, C.c_int, C.c_int]),
    "mrmd_b200_thermo_destroy": (C.c_int, [vp]),
    "mrmd_b200_thermo_info": (C.c_int, [vp, pi64, pi64, pdbl, pi64]),
    "mrmd_b200_thermo_sample": (C.c_int, [vp, vp, vp]),
    "mrmd_b200_thermo_update": (C.c_int, [vp, dbl, dbl, pPred, vp]),
    "mrmd_b200_thermo_apply": (C.c_int, [vp, vp, pPred, C.c_int, vp]),
    "mrmd_b200_thermo_read": (C.c_int, [vp, C.c_int, vp, vp]),
    "mrmd_b200_thermo_write_force": (C.c_int, [vp, vp, vp]),
    "mrmd_b200_thermo_density_ptr": (C.c_int, [vp, pvp, pi64]),
    "mrmd_b200_thermo_mu": (C.c_int, [vp, vp, vp, vp]),
    "mrmd_b200_kinetic_energy": (C.c_int, [vp, pdbl, vp]),
    "mrmd_b200_system_momentum": (C.c_int, [vp, vp, vp]),
    "mrmd_b200_pressure": (C.c_int, [vp, pSub, pdbl, vp]),
    "mrmd_b200_msd_create": (C.c_int, [pvp]),
    "mrmd_b200_msd_destroy": (C.c_int, [vp]),
    "mrmd_b200_msd_reset_atoms": (C.c_int, [vp, vp, vp]),
    "mrmd_b200_msd_reset_molecules": (C.c_int, [vp, vp, vp]),
    "mrmd_b200_msd_calc_atoms": (C.c_int, [vp, vp, pSub, pdbl, vp]),
    "mrmd_b200_msd_calc_molecules": (C.c_int, [vp, vp, pSub, pdbl, vp]),
    "mrmd_b200_berendsen_thermostat": (C.c_int, [vp, dbl, dbl, dbl, vp]),
    "mrmd_b200_berendsen_barostat": (C.c_int, [vp, dbl, dbl, dbl, pSub, C.c_int, C.c_int, C.c_int, vp]),
    "mrmd_b200_limit_acceleration": (C.c_int, [vp, dbl, vp]),
    "mrmd_b200_limit_velocity": (C.c_int, [vp, dbl, vp]),
    "mrmd_b200_constraints_create": (C.c_int, [pvp, i64, i64]),
    "mrmd_b200_constraints_destroy": (C.c_int, [vp]),
    "mrmd_b200_constraints_set": (C.c_int, [vp, vp, vp, vp, i64]),
    "mrmd_b200_constraints_enforce_positional": (C.c_int, [vp, vp, vp, dbl, vp]),
    "mrmd_b200_constraints_enforce_velocity": (C.c_int, [vp, vp, vp, dbl, vp]),
    "mrmd_b200_coulomb_eval": (C.c_int, [C.c_int, dbl, dbl, vp, i64, dbl, dbl, vp, vp, vp]),
    "mrmd_b200_spc_create": (C.c_int, [pvp, C.c_int]),
    "mrmd_b200_spc_destroy": (C.c_int, [vp]),
    "mrmd_b200_spc_apply_forces": (C.c_int, [vp, vp, vp, vp, pdbl, pdbl, vp]),
    "mrmd_b200_spc_calc_bond_energy": (C.c_int, [vp, vp, vp, dbl, pdbl, vp]),
    "mrmd_b200_spc_enforce_positional_constraints": (C.c_int, [vp, vp, vp, dbl, vp]),
    "mrmd_b200_spc_enforce_velocity_constraints": (C.c_int, [vp, vp, vp, dbl, vp]),
    "mrmd_b200_md_create": (C.c_int, [pvp, C.POINTER(MdConfig), pSub, vp]),
    "mrmd_b200_md_destroy": (C.c_int, [vp]),
    "mrmd_b200_md_run": (C.c_int, [vp, i64, C.c_int, C.POINTER(MdStats), vp]),
    "mrmd_b200_md_run_host": (C.c_int, [vp, i64, vp, vp, vp, C.POINTER(MdStats), vp]),
    "mrmd_b200_molecules_cell_sort_with_atoms": (C.c_int, [vp, vp, C.c_int, vp, vp, vp, vp]),
    "mrmd_b200_verlet_build_periodic_molecules": (C.c_int, [vp, vp, pSub, dbl, dbl, i64, C.c_int, vp]),
    "mrmd_b200_verlet_read_periodic_molecules": (C.c_int, [vp, vp, vp, vp, vp, vp]),
    "mrmd_b200_adress_run_periodic_molecules": (C.c_int, [vp, vp, vp, vp, pWeight, C.c_int, pdbl, pi64, vp]),
    "mrmd_b200_hist_create": (C.c_int, [pvp, dbl, dbl, i64, i64]),
    "mrmd_b200_hist_clone": (C.c_int, [pvp, vp, vp]),
    "mrmd_b200_hist_destroy": (C.c_int, [vp]),
    "mrmd_b200_hist_info": (C.c_int, [vp, pdbl, pdbl, pi64, pi64, pdbl, pdbl]),
    "mrmd_b200_hist_device_data": (vp, [vp]),
    "mrmd_b200_hist_write": (C.c_int, [vp, vp, C.c_int, vp]),
    "mrmd_b200_hist_read": (C.c_int, [vp, vp, C.c_int, vp]),
    "mrmd_b200_hist_get_bin": (i64, [vp, dbl]),
    "mrmd_b200_hist_get_bin_position": (dbl, [vp, i64]),
    "mrmd_b200_hist_transform": (C.c_int, [vp, vp, C.c_int, vp]),
    "mrmd_b200_hist_scale": (C.c_int, [vp, dbl, vp]),
    "mrmd_b200_hist_scale_per_histogram": (C.c_int, [vp, vp, i64, vp]),
    "mrmd_b200_hist_make_symmetric": (C.c_int, [vp, vp]),
    "mrmd_b200_hist_cumulative_moving_average": (C.c_int, [vp, vp, dbl, vp]),
    "mrmd_b200_hist_gradient": (C.c_int, [pvp, vp, C.c_int, vp]),
    "mrmd_b200_hist_smoothen": (C.c_int, [pvp, vp, dbl, dbl, C.c_int, vp]),
    "mrmd_b200_hist_replace_if_bin_position": (C.c_int, [vp, pPred, dbl, vp]),
    "mrmd_b200_hist_create_grid": (C.c_int, [vp, vp, vp]),
    "mrmd_b200_thermo_get_hist": (C.c_int, [vp, C.c_int, pvp, vp]),
    "mrmd_b200_md_set_energy_every_step": (C.c_int, [vp, C.c_int]),
    "mrmd_b200_slab_set_energy_every_step": (C.c_int, [vp, C.c_int]),
    "mrmd_b200_slab_run_host": (C.c_int, [vp, i64, vp, vp, vp, C.POINTER(MdStats), vp]),
    "mrmd_b200_nccl_unique_id": (C.c_int, [vp]),
    "mrmd_b200_slab_create": (C.c_int, [pvp, C.POINTER(MdConfig), vp, vp, C.c_int, C.c_int, vp, vp, vp]),
    "mrmd_b200_slab_create_cuts": (C.c_int, [pvp, C.POINTER(MdConfig), vp, vp, vp, C.c_int, C.c_int, vp, vp, vp]),
    "mrmd_b200_slab_destroy": (C.c_int, [vp]),
    "mrmd_b200_slab_run": (C.c_int, [vp, i64, C.c_int, C.POINTER(MdStats), vp]),
    "mrmd_b200_host_alloc": (C.c_int, [pvp, i64]),
    "mrmd_b200_host_free": (C.c_int, [vp]),
}


class MrmdB200Error(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MrmdB200Error(
            f"{LIB_PATH} is missing: build it with `python -m mrmd_b200.build` (or __graft_entry__.build()); "
            "the mrmd_b200 hot path has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().mrmd_b200_last_error().decode(errors="replace")
        raise MrmdB200Error(f"mrmd_b200 call failed (code {rc}): {msg}")
