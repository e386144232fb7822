"""Python host-side mirror of the reference's operator interface for the hot path.

Same class / method names and argument meaning as namespace mrmd (data::Atoms, data::Molecules,
data::Subdomain, HalfVerletList / FullVerletList, action::LennardJones, action::LJ_IdealGas,
action::UpdateMolecules, action::ContributeMoleculeForceToAtoms, action::ThermodynamicForce,
action::VelocityVerlet, action::VelocityVerletLangevinThermostat, communication::GhostLayer /
MultiResGhostLayer, weighting_function::Slab / Spherical, util::IsInSymmetricSlab), each forwarding to the
C ABI of include/mrmd_b200.h.  Used by tests/ and bench.py; the C++20 mirror lives in include/mrmd/.
Errors follow the reference (message + abort) as exceptions: MrmdB200Error.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Pred, Weight, check

HOST, DEVICE = 0, 1
ATOM_FIELDS = {"pos": (0, 3, np.float64), "vel": (1, 3, np.float64), "force": (2, 3, np.float64),
               "type": (3, 1, np.int64), "mass": (4, 1, np.float64), "charge": (5, 1, np.float64),
               "relativeMass": (6, 1, np.float64), "id": (7, 1, np.int64)}
MOL_FIELDS = {"pos": (0, 3, np.float64), "force": (1, 3, np.float64), "lambda": (2, 1, np.float64),
              "modulatedLambda": (3, 1, np.float64), "gradLambda": (4, 3, np.float64),
              "atomsOffset": (5, 1, np.int64), "numAtoms": (6, 1, np.int64)}

PRED_ALWAYS, PRED_NEVER, PRED_SLAB, PRED_SLAB_EITHER, PRED_SLAB_BOTH, PRED_INTERVAL = range(6)


def L():
    return _lib.load()


def _d3(v):
    return np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (3,)))


def _stream(stream):
    return C.c_void_p(stream) if stream else None


class Subdomain(_lib.Subdomain):
    """data::Subdomain (data/Subdomain.hpp:37-110)."""

    def __init__(self, minCorner=(0, 0, 0), maxCorner=(0, 0, 0), ghostLayerThickness=0.0):
        super().__init__()
        a, b, c = _d3(minCorner), _d3(maxCorner), _d3(ghostLayerThickness)
        L().mrmd_b200_subdomain_init(C.byref(self), a.ctypes.data, b.ctypes.data, c.ctypes.data)

    def scaleDim(self, factor, axis):
        L().mrmd_b200_subdomain_scale_dim(C.byref(self), factor, axis)

    def scale(self, factor):
        for axis in range(3):
            self.scaleDim(factor, axis)

    def getVolume(self):
        return self.diameter[0] * self.diameter[1] * self.diameter[2]

    def getCenter(self):
        return [(self.minCorner[d] + self.maxCorner[d]) * 0.5 for d in range(3)]


class IsInSymmetricSlab(Pred):
    """util::IsInSymmetricSlab (util/IsInSymmetricSlab.hpp:24-64) as a parametric predicate."""

    def __init__(self, center, slabMin, slabMax, axis=0, tolerance=0.0, kind=PRED_SLAB):
        c = center[axis] if np.ndim(center) else center
        super().__init__(kind, axis, float(c), float(slabMin), float(slabMax), float(tolerance))

    def either(self):
        """slab(p1) || slab(p2), examples/04:209-216"""
        return Pred(PRED_SLAB_EITHER, self.axis, self.center, self.slabMin, self.slabMax, self.tolerance)

    def both(self):
        """slab(p1) && slab(p2), examples/04:220-227"""
        return Pred(PRED_SLAB_BOTH, self.axis, self.center, self.slabMin, self.slabMax, self.tolerance)


def interval_pred(lower, upper, axis=0):
    return Pred(PRED_INTERVAL, axis, 0.0, float(lower), float(upper), 0.0)


def never_pred():
    return Pred(PRED_NEVER, 0, 0.0, 0.0, 0.0, 0.0)


def _pref(pred):
    return C.byref(pred) if pred is not None else None


class _Container:
    _fields = None
    _prefix = None

    def _call(self, name):
        return getattr(L(), f"mrmd_b200_{self._prefix}_{name}")

    def size(self):
        return int(self._call("size")(self.h))

    def resize(self, n, stream=None):
        check(self._call("resize")(self.h, n, _stream(stream)))

    def set(self, field, values, first=0, stream=None):
        fid, ncomp, dtype = self._fields[field]
        arr = np.ascontiguousarray(values, dtype=dtype)
        count = arr.size // ncomp
        check(self._call("write")(self.h, fid, arr.ctypes.data, first, count, ncomp, 1, HOST, _stream(stream)))

    def get(self, field, first=0, count=None, stream=None):
        fid, ncomp, dtype = self._fields[field]
        if count is None:
            count = self.size() - first
        out = np.zeros((count, ncomp) if ncomp > 1 else (count,), dtype=dtype)
        check(self._call("read")(self.h, fid, out.ctypes.data, first, count, ncomp, 1, HOST, _stream(stream)))
        return out

    def write_ptr(self, field, ptr, first, count, stride, vlen=1, mem=HOST, stream=None):
        check(self._call("write")(self.h, self._fields[field][0], ptr, first, count, stride, vlen, mem, _stream(stream)))

    def read_ptr(self, field, ptr, first, count, stride, vlen=1, mem=HOST, stream=None):
        check(self._call("read")(self.h, self._fields[field][0], ptr, first, count, stride, vlen, mem, _stream(stream)))

    def fill(self, field, value, stream=None):
        check(self._call("fill")(self.h, self._fields[field][0], float(value), _stream(stream)))

    def _counts(self):
        a, b = C.c_int64(), C.c_int64()
        check(self._call("get_counts")(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def _set_counts(self, nl, ng):
        check(self._call("set_counts")(self.h, nl, ng))

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                self._call("destroy")(h)
            except Exception:
                pass


class Atoms(_Container):
    """data::Atoms (data/Atoms.hpp:33-147): locals at [0, numLocalAtoms), ghosts behind them."""
    _fields = ATOM_FIELDS
    _prefix = "atoms"

    def __init__(self, numAtoms):
        self.h = C.c_void_p()
        check(L().mrmd_b200_atoms_create(C.byref(self.h), numAtoms))

    numLocalAtoms = property(lambda s: s._counts()[0], lambda s, v: s._set_counts(v, s._counts()[1]))
    numGhostAtoms = property(lambda s: s._counts()[1], lambda s, v: s._set_counts(s._counts()[0], v))

    def getPos(self):
        return self.get("pos")

    def getVel(self):
        return self.get("vel")

    def getForce(self):
        return self.get("force")

    def getType(self):
        return self.get("type")

    def getMass(self):
        return self.get("mass")

    def getId(self):
        return self.get("id")

    def setForce(self, value, stream=None):
        """Atoms::setForce / Cabana::deep_copy(force, value)"""
        self.fill("force", value, stream)

    def removeGhostAtoms(self):
        nl, _ = self._counts()
        self._set_counts(nl, 0)
        self.resize(nl)

    def permute(self, linkedCellList, stream=None):
        """atoms.permute(LinkedCellList) (data/Atoms.hpp:95)"""
        lc = linkedCellList
        check(L().mrmd_b200_atoms_cell_sort(self.h, lc.begin, lc.end, lc.delta.ctypes.data, lc.gmin.ctypes.data,
                                            lc.gmax.ctypes.data, None, _stream(stream)))

    @classmethod
    def from_arrays(cls, pos, vel=None, mass=1.0, type=None, relativeMass=None, capacity=None, ids=None):
        n = len(pos)
        a = cls(capacity or n)
        a.resize(n)
        a.set("pos", pos)
        if vel is not None:
            a.set("vel", vel)
        a.set("mass", np.broadcast_to(np.asarray(mass, dtype=np.float64), (n,)))
        if type is not None:
            a.set("type", type)
        if relativeMass is not None:
            a.set("relativeMass", np.broadcast_to(np.asarray(relativeMass, dtype=np.float64), (n,)))
        if ids is not None:  # global atom ids (default: the index), the Philox counter of the Langevin thermostat
            a.set("id", ids)
        a.numLocalAtoms = n
        return a


class Molecules(_Container):
    """data::Molecules (data/Molecules.hpp:27-148)."""
    _fields = MOL_FIELDS
    _prefix = "molecules"

    def __init__(self, numMolecules, _handle=None):
        if _handle is not None:
            self.h = _handle
            return
        self.h = C.c_void_p()
        check(L().mrmd_b200_molecules_create(C.byref(self.h), numMolecules))

    numLocalMolecules = property(lambda s: s._counts()[0], lambda s, v: s._set_counts(v, s._counts()[1]))
    numGhostMolecules = property(lambda s: s._counts()[1], lambda s, v: s._set_counts(s._counts()[0], v))

    def setForce(self, value, stream=None):
        self.fill("force", value, stream)

    def permute(self, linkedCellList, stream=None):
        lc = linkedCellList
        check(L().mrmd_b200_molecules_cell_sort(self.h, lc.begin, lc.end, lc.delta.ctypes.data, lc.gmin.ctypes.data,
                                                lc.gmax.ctypes.data, _stream(stream)))


def createMoleculeForEachAtom(atoms, stream=None):
    """data::createMoleculeForEachAtom (data/MoleculesFromAtoms.cpp:19-39)"""
    h = C.c_void_p()
    check(L().mrmd_b200_molecules_for_each_atom(C.byref(h), atoms.h, _stream(stream)))
    return Molecules(0, _handle=h)


class LinkedCellList:
    """Cabana::LinkedCellList(pos, begin, end, gridDelta, gridMin, gridMax) as used in tests/NVT/NVT.cpp:136-144.
    The binning itself runs inside permute() (one fused sort + gather on the device)."""

    def __init__(self, begin, end, gridDelta, gridMin, gridMax):
        self.begin, self.end = int(begin), int(end)
        self.delta, self.gmin, self.gmax = _d3(gridDelta), _d3(gridMin), _d3(gridMax)


class _VerletList:
    half = True

    def __init__(self):
        self.h = C.c_void_p()
        check(L().mrmd_b200_verlet_create(C.byref(self.h), int(self.half)))

    def build(self, particles, begin, end, radius, cellRatio, gridMin, gridMax, maxNeigh=60, stream=None):
        """VerletList::build(pos, begin, end, radius, ratio, gridMin, gridMax, maxNeigh); `particles` is the
        Atoms or Molecules container whose position slice the reference would pass."""
        gmin, gmax = _d3(gridMin), _d3(gridMax)
        fn = L().mrmd_b200_verlet_build_atoms if isinstance(particles, Atoms) else L().mrmd_b200_verlet_build_molecules
        check(fn(self.h, particles.h, begin, end, radius, cellRatio, gmin.ctypes.data, gmax.ctypes.data, maxNeigh,
                 _stream(stream)))

    def build_periodic(self, atoms, subdomain, radius, cellRatio=1.0, maxNeigh=60, stream=None):
        """B200 fast path: createGhostAtoms + build over the ghosts in one tiled pass (atoms must be cell-sorted
        over [0, numLocalAtoms) on the subdomain grid)."""
        check(L().mrmd_b200_verlet_build_periodic(self.h, atoms.h, C.byref(subdomain), radius, cellRatio, maxNeigh,
                                                  _stream(stream)))

    def to_host_periodic(self, atoms):
        """(counts[n], partner[n, width], shiftCode[n, width]) of a periodic (tiled) list"""
        i = self.info()
        n, w = i["numParticles"], max(i["width"], 1)
        counts = np.zeros(n, dtype=np.int32)
        partner = np.full((n, w), -1, dtype=np.int32)
        code = np.full((n, w), -1, dtype=np.int32)
        check(L().mrmd_b200_verlet_read_periodic(self.h, atoms.h, counts.ctypes.data, partner.ctypes.data,
                                                 code.ctypes.data, None))
        return counts, partner, code

    def info(self):
        n, w, t, h = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
        check(L().mrmd_b200_verlet_info(self.h, C.byref(n), C.byref(w), C.byref(t), C.byref(h)))
        return {"numParticles": n.value, "width": w.value, "totalPairs": t.value, "half": bool(h.value)}

    def to_host(self):
        """(counts[numParticles], neighbors[numParticles, width]) in Cabana's VerletLayout2D"""
        i = self.info()
        counts = np.zeros(i["numParticles"], dtype=np.int32)
        neigh = np.full((i["numParticles"], max(i["width"], 1)), -1, dtype=np.int32)
        check(L().mrmd_b200_verlet_read(self.h, counts.ctypes.data, neigh.ctypes.data, HOST, None))
        return counts, neigh

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                L().mrmd_b200_verlet_destroy(h)
            except Exception:
                pass


class HalfVerletList(_VerletList):
    half = True


class FullVerletList(_VerletList):
    half = False


def _arr(v, n):
    return np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64).ravel(), (n,)))


class LennardJones:
    """action::LennardJones (action/LennardJones.hpp:98-133)."""

    def __init__(self, rc, sigma, epsilon, cappingDistance=0.0, numTypes=1, isShifted=False):
        n = numTypes * numTypes
        arrs = [_arr(v, n) for v in (cappingDistance, rc, sigma, epsilon)]
        self.h = C.c_void_p()
        check(L().mrmd_b200_lj_create(C.byref(self.h), *[a.ctypes.data for a in arrs], numTypes, int(isShifted)))

    def apply(self, atoms, verletList, stream=None):
        check(L().mrmd_b200_lj_apply(self.h, atoms.h, verletList.h, None, _stream(stream)))

    def apply_if(self, atoms, verletList, pred, stream=None):
        check(L().mrmd_b200_lj_apply(self.h, atoms.h, verletList.h, _pref(pred), _stream(stream)))

    def _get(self, stream=None):
        e, v, p = C.c_double(), C.c_double(), C.c_int64()
        check(L().mrmd_b200_lj_get(self.h, C.byref(e), C.byref(v), C.byref(p), _stream(stream)))
        return e.value, v.value, p.value

    def getEnergy(self):
        return self._get()[0]

    def getVirial(self):
        return self._get()[1]

    def getNumPairs(self):
        return self._get()[2]

    def computeForceAndEnergy(self, distSqr, typeIdx=0):
        d = np.ascontiguousarray(distSqr, dtype=np.float64).ravel()
        ff, e = np.zeros_like(d), np.zeros_like(d)
        check(L().mrmd_b200_lj_eval(self.h, typeIdx, d.ctypes.data, d.size, ff.ctypes.data, e.ctypes.data, None))
        return ff, e

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                L().mrmd_b200_lj_destroy(h)
            except Exception:
                pass


class VelocityVerlet:
    """action::VelocityVerlet (action/VelocityVerlet.hpp)."""

    @staticmethod
    def preForceIntegrate(atoms, dt, stream=None, fetch=True):
        out = C.c_double()
        check(L().mrmd_b200_vv_pre(atoms.h, dt, C.byref(out) if fetch else None, _stream(stream)))
        return out.value

    @staticmethod
    def postForceIntegrate(atoms, dt, stream=None):
        check(L().mrmd_b200_vv_post(atoms.h, dt, _stream(stream)))


class VelocityVerletLangevinThermostat:
    """action::VelocityVerletLangevinThermostat (VelocityVerletLangevinThermostat.hpp:29-61).  The pool seed
    of the reference (1234, :32) keys a Philox4x32-10 stream; every call advances the step counter."""

    def __init__(self, zeta, temperature, seed=1234):
        self.set(zeta, temperature)
        self.seed = seed
        self.step = 0

    def set(self, zeta, temperature):
        self.zeta, self.temperature = float(zeta), float(temperature)

    def preForceIntegrate(self, atoms, dt, stream=None, fetch=True):
        return self.preForceIntegrate_apply_if(atoms, dt, None, stream, fetch)

    def preForceIntegrate_apply_if(self, atoms, dt, pred, stream=None, fetch=True):
        out = C.c_double()
        check(L().mrmd_b200_langevin_pre(atoms.h, dt, self.zeta, self.temperature, self.seed, self.step, _pref(pred),
                                         C.byref(out) if fetch else None, _stream(stream)))
        self.step += 1
        return out.value

    def postForceIntegrate(self, atoms, dt, stream=None):
        VelocityVerlet.postForceIntegrate(atoms, dt, stream)


class GhostLayer:
    """communication::GhostLayer (communication/GhostLayer.hpp:28-56)."""

    def __init__(self):
        self.h = C.c_void_p()
        check(L().mrmd_b200_ghost_create(C.byref(self.h)))

    def exchangeRealAtoms(self, atoms, subdomain, stream=None):
        check(L().mrmd_b200_ghost_map_into_domain(atoms.h, C.byref(subdomain), _stream(stream)))

    def createGhostAtoms(self, atoms, subdomain, stream=None, axis=-1):
        check(L().mrmd_b200_ghost_create_atoms(self.h, atoms.h, C.byref(subdomain), axis, _stream(stream)))

    def resetCorrespondingRealAtoms(self, atoms, stream=None):
        check(L().mrmd_b200_ghost_reset(self.h, atoms.h, _stream(stream)))

    def updateGhostAtoms(self, atoms, subdomain, stream=None):
        check(L().mrmd_b200_ghost_update(self.h, atoms.h, C.byref(subdomain), _stream(stream)))

    def contributeBackGhostToReal(self, atoms, stream=None):
        check(L().mrmd_b200_ghost_contribute_back(self.h, atoms.h, _stream(stream)))

    def correspondingRealAtom(self, count, first=0):
        out = np.zeros(count, dtype=np.int64)
        check(L().mrmd_b200_ghost_read_corresponding(self.h, out.ctypes.data, first, count, HOST, None))
        return out

    def setCorrespondingRealAtom(self, values, first=0):
        v = np.ascontiguousarray(values, dtype=np.int64)
        check(L().mrmd_b200_ghost_write_corresponding(self.h, v.ctypes.data, first, v.size, HOST, None))

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                L().mrmd_b200_ghost_destroy(h)
            except Exception:
                pass


class MultiResGhostLayer(GhostLayer):
    """communication::MultiResGhostLayer (communication/MultiResGhostLayer.hpp:29-61)."""

    def exchangeRealAtoms(self, molecules, atoms, subdomain, stream=None):
        check(L().mrmd_b200_ghost_mr_map_into_domain(molecules.h, atoms.h, C.byref(subdomain), _stream(stream)))

    def createGhostAtoms(self, molecules, atoms, subdomain, stream=None, axis=-1):
        check(L().mrmd_b200_ghost_mr_create_atoms(self.h, molecules.h, atoms.h, C.byref(subdomain), axis, _stream(stream)))


class Slab(Weight):
    """weighting_function::Slab(center, atomisticRegionDiameter, hybridRegionDiameter, nu, interfaceType)"""

    def __init__(self, center, atomisticRegionDiameter, hybridRegionDiameter, nu, abrupt=False):
        super().__init__()
        self.kind, self.abrupt = 0, int(abrupt)
        for d in range(3):
            self.center[d] = float(center[d])
        self.atRegion, self.hyRegion, self.exponent = atomisticRegionDiameter, hybridRegionDiameter, nu


class Spherical(Weight):
    """weighting_function::Spherical(center, atomisticRadius, hybridRegionDiameter, exponent); lambda^mod := lambda"""

    def __init__(self, center, atomisticRadius, hybridRegionDiameter, exponent):
        super().__init__()
        self.kind, self.abrupt = 1, 0
        for d in range(3):
            self.center[d] = float(center[d])
        self.atRegion, self.hyRegion, self.exponent = atomisticRadius, hybridRegionDiameter, exponent


def weight_eval(weight, pos):
    p = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
    n = len(p)
    lam, mod, grad = np.zeros(n), np.zeros(n), np.zeros((n, 3))
    check(L().mrmd_b200_weight_eval(C.byref(weight), p.ctypes.data, n, lam.ctypes.data, mod.ctypes.data,
                                    grad.ctypes.data, None))
    return lam, mod, grad


class UpdateMolecules:
    """action::UpdateMolecules::update (action/UpdateMolecules.hpp:24-70)"""

    @staticmethod
    def update(molecules, atoms, weight, stream=None):
        check(L().mrmd_b200_molecules_update(molecules.h, atoms.h, C.byref(weight), _stream(stream)))


class ContributeMoleculeForceToAtoms:
    """action::ContributeMoleculeForceToAtoms::update (action/ContributeMoleculeForceToAtoms.cpp:23-48)"""

    @staticmethod
    def update(molecules, atoms, stream=None):
        check(L().mrmd_b200_molecules_contribute_force(molecules.h, atoms.h, _stream(stream)))


class LJ_IdealGas:
    """action::LJ_IdealGas (action/LJ_IdealGas.hpp:34-104)."""

    def __init__(self, cappingDistance, rc, sigma, epsilon, doShift, numTypes=1):
        n = numTypes * numTypes
        arrs = [_arr(v, n) for v in (cappingDistance, rc, sigma, epsilon)]
        self.h = C.c_void_p()
        self.numTypes = numTypes
        check(L().mrmd_b200_adress_create(C.byref(self.h), *[a.ctypes.data for a in arrs], numTypes, int(doShift)))
        self._sampling, self._update = 200, 20000

    def setAtomsPerMolecule(self, atomsPerMolecule):
        """mrmd_b200_adress_set_atoms_per_molecule (extension): 4 selects the four-lanes-per-molecule kernel"""
        check(L().mrmd_b200_adress_set_atoms_per_molecule(self.h, int(atomsPerMolecule)))

    def setCompensationEnergySamplingInterval(self, interval):
        self._sampling = interval
        check(L().mrmd_b200_adress_set_intervals(self.h, self._sampling, self._update))

    def setCompensationEnergyUpdateInterval(self, interval):
        self._update = interval
        check(L().mrmd_b200_adress_set_intervals(self.h, self._sampling, self._update))

    def run(self, molecules, verletList, atoms, stream=None, fetch=True):
        e, p = C.c_double(), C.c_int64()
        check(L().mrmd_b200_adress_run(self.h, molecules.h, verletList.h, atoms.h, C.byref(e) if fetch else None,
                                       C.byref(p) if fetch else None, _stream(stream)))
        self.lastNumPairs = p.value
        return e.value

    def run_periodic(self, atoms, verletList, weight, stream=None, fetch=True):
        """B200 fast path for one-atom molecules on a build_periodic list: UpdateMolecules + run +
        ContributeMoleculeForceToAtoms in one kernel (mrmd_b200_adress_run_periodic)."""
        e, p = C.c_double(), C.c_int64()
        check(L().mrmd_b200_adress_run_periodic(self.h, atoms.h, verletList.h, C.byref(weight),
                                                C.byref(e) if fetch else None, C.byref(p) if fetch else None,
                                                _stream(stream)))
        self.lastNumPairs = p.value
        return e.value

    def _hist(self, kind):
        out = np.zeros((200, self.numTypes))
        check(L().mrmd_b200_adress_read_histogram(self.h, kind, out.ctypes.data, None))
        return out

    def getMeanCompensationEnergy(self):
        return self._hist(0)

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                L().mrmd_b200_adress_destroy(h)
            except Exception:
                pass


class MultiHistogram:
    """data::MultiHistogram (data/MultiHistogram.hpp:29-92) on the device; `data` is a host copy (numBins x numHistograms),
    `set_data` uploads one."""

    def __init__(self, label, min, max, numBins, numHistograms, _handle=None):
        self.label = label
        self.h = _handle if _handle is not None else C.c_void_p()
        if _handle is None:
            check(L().mrmd_b200_hist_create(C.byref(self.h), float(min), float(max), int(numBins), int(numHistograms)))
        a, b, g, s = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        nb, nh = C.c_int64(), C.c_int64()
        check(L().mrmd_b200_hist_info(self.h, C.byref(a), C.byref(b), C.byref(nb), C.byref(nh), C.byref(g), C.byref(s)))
        self.min, self.max, self.numBins, self.numHistograms = a.value, b.value, nb.value, nh.value
        self.binSize, self.inverseBinSize = g.value, s.value

    @classmethod
    def _wrap(cls, handle, label="histogram"):
        return cls(label, 0, 1, 0, 0, _handle=handle)

    def copy(self, label="copy"):
        h = C.c_void_p()
        check(L().mrmd_b200_hist_clone(C.byref(h), self.h, None))
        return MultiHistogram._wrap(h, label)

    @property
    def data(self):
        out = np.zeros((self.numBins, self.numHistograms))
        check(L().mrmd_b200_hist_read(self.h, out.ctypes.data, HOST, None))
        return out

    def set_data(self, values):
        arr = np.ascontiguousarray(values, dtype=np.float64).reshape(self.numBins, self.numHistograms)
        check(L().mrmd_b200_hist_write(self.h, arr.ctypes.data, HOST, None))

    def getBin(self, val):
        return int(L().mrmd_b200_hist_get_bin(self.h, float(val)))

    def getBinPosition(self, binIdx):
        return float(L().mrmd_b200_hist_get_bin_position(self.h, int(binIdx)))

    def _op(self, rhs, op):
        check(L().mrmd_b200_hist_transform(self.h, rhs.h, op, None))
        return self

    def __iadd__(self, rhs):
        return self._op(rhs, 0)

    def __isub__(self, rhs):
        return self._op(rhs, 1)

    def __imul__(self, rhs):
        return self._op(rhs, 2)

    def __itruediv__(self, rhs):
        return self._op(rhs, 3)

    def scale(self, factor):
        """scale(real_t) or scale(ScalarView): one factor, or one per histogram"""
        if np.ndim(factor) == 0:
            check(L().mrmd_b200_hist_scale(self.h, float(factor), None))
        else:
            f = np.ascontiguousarray(factor, dtype=np.float64)
            check(L().mrmd_b200_hist_scale_per_histogram(self.h, f.ctypes.data, f.size, None))

    def makeSymmetric(self):
        check(L().mrmd_b200_hist_make_symmetric(self.h, None))

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                L().mrmd_b200_hist_destroy(h)
            except Exception:
                pass


def cumulativeMovingAverage(average, current, movingAverageFactor=10.0):
    check(L().mrmd_b200_hist_cumulative_moving_average(average.h, current.h, float(movingAverageFactor), None))


def gradient(histogram, periodic=False):
    h = C.c_void_p()
    check(L().mrmd_b200_hist_gradient(C.byref(h), histogram.h, int(periodic), None))
    return MultiHistogram._wrap(h, "gradient")


def smoothen(histogram, sigma, range, periodic=False):
    h = C.c_void_p()
    check(L().mrmd_b200_hist_smoothen(C.byref(h), histogram.h, float(sigma), float(range), int(periodic), None))
    return MultiHistogram._wrap(h, "smooth-input")


def createGrid(histogram):
    out = np.zeros(histogram.numBins)
    check(L().mrmd_b200_hist_create_grid(histogram.h, out.ctypes.data, None))
    return out


def replace_if_bin_position(histogram, pred, newValue):
    """pred: a parametric one-coordinate predicate (IsInSymmetricSlab, interval_pred, ...) evaluated at the bin position"""
    check(L().mrmd_b200_hist_replace_if_bin_position(histogram.h, C.byref(pred), float(newValue), None))


class ThermodynamicForce:
    """action::ThermodynamicForce (action/ThermodynamicForce.hpp:32-96)."""

    def __init__(self, targetDensity, subdomain, requestedDensityBinWidth, thermodynamicForceModulation,
                 enforceSymmetry=False, usePeriodicity=False):
        td = np.ascontiguousarray(np.atleast_1d(targetDensity), dtype=np.float64)
        mod = np.ascontiguousarray(np.atleast_1d(thermodynamicForceModulation), dtype=np.float64)
        assert td.size == mod.size
        self.h = C.c_void_p()
        check(L().mrmd_b200_thermo_create(C.byref(self.h), td.ctypes.data, td.size, C.byref(subdomain),
                                          requestedDensityBinWidth, mod.ctypes.data, int(enforceSymmetry),
                                          int(usePeriodicity)))

    def _info(self):
        nb, nt, bs, s = C.c_int64(), C.c_int64(), C.c_double(), C.c_int64()
        check(L().mrmd_b200_thermo_info(self.h, C.byref(nb), C.byref(nt), C.byref(bs), C.byref(s)))
        return nb.value, nt.value, bs.value, s.value

    numBins = property(lambda s: s._info()[0])
    binSize = property(lambda s: s._info()[2])

    def getNumberOfDensityProfileSamples(self):
        return self._info()[3]

    def _read(self, kind):
        nb, nt, _, _ = self._info()
        out = np.zeros((nb, nt))
        check(L().mrmd_b200_thermo_read(self.h, kind, out.ctypes.data, None))
        return out

    def getForce(self, typeId=None):
        f = self._read(0)
        return f if typeId is None else f[:, typeId]

    def getForceHistogram(self):
        """getForce() as the data::MultiHistogram the reference returns (ThermodynamicForce.hpp:57)"""
        h = C.c_void_p()
        check(L().mrmd_b200_thermo_get_hist(self.h, 0, C.byref(h), None))
        return MultiHistogram._wrap(h, "thermodynamic-force")

    def getDensityProfileHistogram(self):
        h = C.c_void_p()
        check(L().mrmd_b200_thermo_get_hist(self.h, 1, C.byref(h), None))
        return MultiHistogram._wrap(h, "density-profile")

    def getDensityProfile(self, typeId=None):
        d = self._read(1)
        return d if typeId is None else d[:, typeId]

    def setForce(self, forces):
        f = np.ascontiguousarray(forces, dtype=np.float64)
        check(L().mrmd_b200_thermo_write_force(self.h, f.ctypes.data, None))

    def sample(self, atoms, stream=None):
        check(L().mrmd_b200_thermo_sample(self.h, atoms.h, _stream(stream)))

    def update(self, smoothingSigma, smoothingIntensity, stream=None):
        check(L().mrmd_b200_thermo_update(self.h, smoothingSigma, smoothingIntensity, None, _stream(stream)))

    def update_if(self, smoothingSigma, smoothingIntensity, pred, stream=None):
        check(L().mrmd_b200_thermo_update(self.h, smoothingSigma, smoothingIntensity, _pref(pred), _stream(stream)))

    def apply(self, atoms, stream=None):
        check(L().mrmd_b200_thermo_apply(self.h, atoms.h, None, 0, _stream(stream)))

    def apply_if(self, atoms, pred, stream=None):
        check(L().mrmd_b200_thermo_apply(self.h, atoms.h, _pref(pred), 0, _stream(stream)))

    def applyInterpolated_if(self, atoms, pred, stream=None):
        check(L().mrmd_b200_thermo_apply(self.h, atoms.h, _pref(pred), 1, _stream(stream)))

    def densityPointer(self):
        p, n = C.c_void_p(), C.c_int64()
        check(L().mrmd_b200_thermo_density_ptr(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def _mu(self):
        nt = self._info()[1]
        l, r = np.zeros(nt), np.zeros(nt)
        check(L().mrmd_b200_thermo_mu(self.h, l.ctypes.data, r.ctypes.data, None))
        return l, r

    def getMuLeft(self):
        return self._mu()[0]

    def getMuRight(self):
        return self._mu()[1]

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                L().mrmd_b200_thermo_destroy(h)
            except Exception:
                pass


class BerendsenThermostat:
    """action::BerendsenThermostat::apply (action/BerendsenThermostat.cpp:25-50)"""

    @staticmethod
    def apply(atoms, currentTemperature, targetTemperature, gamma, stream=None):
        check(L().mrmd_b200_berendsen_thermostat(atoms.h, currentTemperature, targetTemperature, gamma, _stream(stream)))


class BerendsenBarostat:
    """action::BerendsenBarostat::apply (action/BerendsenBarostat.cpp:23-50); scales `subdomain` in place"""

    @staticmethod
    def apply(atoms, currentPressure, targetPressure, gamma, subdomain, stretchX=True, stretchY=True, stretchZ=True,
              stream=None):
        check(L().mrmd_b200_berendsen_barostat(atoms.h, currentPressure, targetPressure, gamma, C.byref(subdomain),
                                               int(stretchX), int(stretchY), int(stretchZ), _stream(stream)))


def limitAccelerationPerComponent(atoms, maxAccelerationPerComponent, stream=None):
    """action::limitAccelerationPerComponent (action/LimitAcceleration.cpp:21-45)"""
    check(L().mrmd_b200_limit_acceleration(atoms.h, maxAccelerationPerComponent, _stream(stream)))


def limitVelocityPerComponent(atoms, maxVelocityPerComponent, stream=None):
    """action::limitVelocityPerComponent (action/LimitVelocity.cpp:23-43)"""
    check(L().mrmd_b200_limit_velocity(atoms.h, maxVelocityPerComponent, _stream(stream)))


class MoleculeConstraints:
    """action::MoleculeConstraints (action/Shake.hpp:159-251): SHAKE / RATTLE over the bonds of every local molecule"""

    def __init__(self, atomsPerMolecule, numConstraintIterations):
        self.h = C.c_void_p()
        check(L().mrmd_b200_constraints_create(C.byref(self.h), atomsPerMolecule, numConstraintIterations))

    def setConstraints(self, bonds):
        """bonds: iterable of (idx, jdx, eqDistance) (data::Bond)"""
        b = list(bonds)
        idx = np.ascontiguousarray([x[0] for x in b], dtype=np.int64)
        jdx = np.ascontiguousarray([x[1] for x in b], dtype=np.int64)
        eq = np.ascontiguousarray([x[2] for x in b], dtype=np.float64)
        check(L().mrmd_b200_constraints_set(self.h, idx.ctypes.data, jdx.ctypes.data, eq.ctypes.data, len(b)))

    def enforcePositionalConstraints(self, molecules, atoms, dt, stream=None):
        check(L().mrmd_b200_constraints_enforce_positional(self.h, molecules.h, atoms.h, dt, _stream(stream)))

    def enforceVelocityConstraints(self, molecules, atoms, dt, stream=None):
        check(L().mrmd_b200_constraints_enforce_velocity(self.h, molecules.h, atoms.h, dt, _stream(stream)))

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                L().mrmd_b200_constraints_destroy(h)
            except Exception:
                pass


class Coulomb:
    """action::impl::Coulomb (action/Coulomb.hpp:27-46)"""

    kind = 0

    def __init__(self):
        self.rc, self.alpha = 0.0, 0.0

    def _eval(self, distSqr, q1, q2, stream=None):
        d = np.ascontiguousarray(np.atleast_1d(np.asarray(distSqr, dtype=np.float64)))
        f, e = np.zeros_like(d), np.zeros_like(d)
        check(L().mrmd_b200_coulomb_eval(self.kind, self.rc, self.alpha, d.ctypes.data, d.size, q1, q2, f.ctypes.data,
                                         e.ctypes.data, _stream(stream)))
        return f, e

    def computeForce(self, distSqr, q1, q2):
        f = self._eval(distSqr, q1, q2)[0]
        return f if np.ndim(distSqr) else float(f[0])

    def computeEnergy(self, distSqr, q1, q2):
        e = self._eval(distSqr, q1, q2)[1]
        return e if np.ndim(distSqr) else float(e[0])


class CoulombDSF(Coulomb):
    """action::impl::CoulombDSF(rc, alpha) (action/CoulombDSF.hpp:42-84)"""

    kind = 1

    def __init__(self, rc, alpha):
        self.rc, self.alpha = float(rc), float(alpha)


class SPC:
    """action::SPC (action/SPC.hpp:61-362).  coulombKind 0 is the reference's impl::Coulomb member, 1 evaluates the
    charges with CoulombDSF(rc, alpha)."""

    massO, chargeO = 15.999, -0.82
    massH, chargeH = 1.008, +0.41
    sigma, epsilon, rc = 0.31655578901998815, 0.6501695808187486, 1.2
    alpha = 2.0
    eqDistanceHO = 0.1
    angleHOH = 109.47 / 180.0 * np.pi

    def __init__(self, coulombKind=0):
        self.h = C.c_void_p()
        check(L().mrmd_b200_spc_create(C.byref(self.h), int(coulombKind)))
        self.eqDistanceHH = self.eqDistanceHO * float(np.sqrt(2.0 - 2.0 * np.cos(self.angleHOH)))
        self.sumEnergyLJ_ = 0.0
        self.sumEnergyCoulomb_ = 0.0

    def applyForces(self, molecules, verletList, atoms, stream=None):
        eLJ, eC = C.c_double(), C.c_double()
        check(L().mrmd_b200_spc_apply_forces(self.h, molecules.h, verletList.h, atoms.h, C.byref(eLJ), C.byref(eC),
                                             _stream(stream)))
        self.sumEnergyLJ_, self.sumEnergyCoulomb_ = eLJ.value, eC.value

    def getEnergyLJ(self):
        return self.sumEnergyLJ_

    def getEnergyCoulomb(self):
        return self.sumEnergyCoulomb_

    def calcBondEnergy(self, molecules, atoms, harmonicPreFactor, stream=None):
        e = C.c_double()
        check(L().mrmd_b200_spc_calc_bond_energy(self.h, molecules.h, atoms.h, harmonicPreFactor, C.byref(e), _stream(stream)))
        return e.value

    def enforcePositionalConstraints(self, molecules, atoms, dt, stream=None):
        check(L().mrmd_b200_spc_enforce_positional_constraints(self.h, molecules.h, atoms.h, dt, _stream(stream)))

    def enforceVelocityConstraints(self, molecules, atoms, dt, stream=None):
        check(L().mrmd_b200_spc_enforce_velocity_constraints(self.h, molecules.h, atoms.h, dt, _stream(stream)))

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                L().mrmd_b200_spc_destroy(h)
            except Exception:
                pass


class analysis:
    """namespace mrmd::analysis: the diagnostics of the drivers' statistics lines (examples/02:190-199)"""

    @staticmethod
    def getKineticEnergy(atoms, stream=None):
        """analysis/KineticEnergy.hpp:26-39"""
        e = C.c_double()
        check(L().mrmd_b200_kinetic_energy(atoms.h, C.byref(e), _stream(stream)))
        return e.value

    @staticmethod
    def getMeanKineticEnergy(atoms, stream=None):
        """analysis/KineticEnergy.hpp:44-47"""
        return analysis.getKineticEnergy(atoms, stream) / float(atoms.numLocalAtoms)

    @staticmethod
    def getSystemMomentum(atoms, stream=None):
        """analysis/SystemMomentum.cpp:21-50 (sums velocities)"""
        out = np.zeros(3)
        check(L().mrmd_b200_system_momentum(atoms.h, out.ctypes.data, _stream(stream)))
        return out

    @staticmethod
    def getPressure(atoms, subdomain, stream=None):
        """analysis/Pressure.cpp:23-51"""
        p = C.c_double()
        check(L().mrmd_b200_pressure(atoms.h, C.byref(subdomain), C.byref(p), _stream(stream)))
        return p.value

    class MeanSquareDisplacement:
        """analysis/MeanSquareDisplacement.hpp:24-49"""

        def __init__(self):
            self.h = C.c_void_p()
            check(L().mrmd_b200_msd_create(C.byref(self.h)))

        def reset(self, items, stream=None):
            fn = L().mrmd_b200_msd_reset_atoms if isinstance(items, Atoms) else L().mrmd_b200_msd_reset_molecules
            check(fn(self.h, items.h, _stream(stream)))

        def calc(self, items, subdomain, stream=None):
            fn = L().mrmd_b200_msd_calc_atoms if isinstance(items, Atoms) else L().mrmd_b200_msd_calc_molecules
            out = C.c_double()
            check(fn(self.h, items.h, C.byref(subdomain), C.byref(out), _stream(stream)))
            return out.value

        def __del__(self):
            h, self.h = getattr(self, "h", None), None
            if h:
                try:
                    L().mrmd_b200_msd_destroy(h)
                except Exception:
                    pass


class MolecularDynamics:
    """The step loop of examples/02 (rebuild policy), examples/01 (Langevin) and tests/NVT (spatial sort at
    rebuild) as one C++ driver inside the library: mrmd_b200_md_* in include/mrmd_b200.h."""

    def __init__(self, atoms, subdomain, dt=0.002, rc=2.5, skin=0.1, sigma=1.0, epsilon=1.0, cappingDistance=0.7,
                 maxNeighbors=60, langevin=False, zeta=20.0, temperature=1.5, seed=1234, cellSort=True, fullList=False,
                 adress=False, weight=None, doShift=True, thermo=None, atomsPerMolecule=1, numConstraintIterations=0,
                 bondLength=1.0):
        cfg = _lib.MdConfig()
        cfg.atomsPerMolecule, cfg.numConstraintIterations, cfg.bondLength = atomsPerMolecule, numConstraintIterations, bondLength
        cfg.dt, cfg.rc, cfg.skin, cfg.sigma, cfg.epsilon, cfg.cappingDistance = dt, rc, skin, sigma, epsilon, cappingDistance
        cfg.maxNeighbors, cfg.integrator, cfg.cellSort, cfg.fullList = maxNeighbors, int(langevin), int(cellSort), int(fullList)
        cfg.adress, cfg.zeta, cfg.temperature, cfg.seed, cfg.doShift = int(adress), zeta, temperature, seed, int(doShift)
        if weight is not None:
            C.memmove(C.byref(cfg.weight), C.byref(weight), C.sizeof(Weight))
        if thermo is not None:
            cfg.useThermoForce = 1
            cfg.thermoTargetDensity, cfg.thermoBinWidth, cfg.thermoModulation = thermo["targetDensity"], thermo["binWidth"], thermo["modulation"]
            cfg.thermoSampleInterval, cfg.thermoUpdateInterval = thermo["sampleInterval"], thermo["updateInterval"]
            cfg.thermoSmoothingSigma, cfg.thermoSmoothingIntensity = thermo["sigma"], thermo["range"]
        self.cfg, self.atoms, self.subdomain = cfg, atoms, subdomain
        self.h = C.c_void_p()
        check(L().mrmd_b200_md_create(C.byref(self.h), C.byref(cfg), C.byref(subdomain), atoms.h))

    def run(self, nsteps, timeForceKernel=False, stream=None):
        st = _lib.MdStats()
        check(L().mrmd_b200_md_run(self.h, nsteps, int(timeForceKernel), C.byref(st), _stream(stream)))
        return {f: getattr(st, f) for f, _ in st._fields_}

    def setEnergyEveryStep(self, enabled):
        """reduce energy and virial on every step (LennardJones.hpp:187-188) instead of on a run's last step only"""
        check(L().mrmd_b200_md_set_energy_every_step(self.h, int(enabled)))

    def run_host(self, nsteps, posHostPtr, velHostPtr, scalarsHostPtr=None, stream=None):
        st = _lib.MdStats()
        check(L().mrmd_b200_md_run_host(self.h, nsteps, posHostPtr, velHostPtr, scalarsHostPtr, C.byref(st), _stream(stream)))
        return {f: getattr(st, f) for f, _ in st._fields_}

    def close(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            L().mrmd_b200_md_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedBuffer:
    """cudaMallocHost-backed numpy array (host side of the host-buffer path)."""

    def __init__(self, shape, dtype=np.float64):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = C.c_void_p()
        check(L().mrmd_b200_host_alloc(C.byref(self.ptr), self.nbytes))
        buf = (C.c_char * self.nbytes).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def __del__(self):
        p, self.ptr = getattr(self, "ptr", None), None
        if p:
            try:
                self.array = None
                L().mrmd_b200_host_free(p)
            except Exception:
                pass


def sync(stream=None):
    check(L().mrmd_b200_sync(_stream(stream)))


def launch_count():
    return int(L().mrmd_b200_launch_count())
