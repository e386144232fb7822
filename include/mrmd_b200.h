/*
 * mrmd_b200.h -- C ABI of the B200-native MRMD force + neighbour hot path.
 *
 * The reference (XzzX/mrmd) has no FFI: its boundary is the public C++ API of
 * namespace mrmd.  Every entry point below replaces the
 * Kokkos/Cabana kernel(s) behind one reference member function; the citation
 * after "replaces" is the reference file:line.  include/mrmd/ holds the C++20
 * mirror of the reference classes that forwards to these functions and
 * INTEGRATION.md shows the binding a maintainer would add on the reference side.
 *
 * Conventions
 *   - plain C types only; all handles are opaque; no torch / Kokkos types.
 *   - every function returns 0 on success, a positive cudaError_t, or a negative
 *     MRMD_B200_E* code; mrmd_b200_last_error() gives the message.
 *   - "stream" is a cudaStream_t passed as void* (NULL = default stream).  Work is
 *     enqueued asynchronously unless the function returns a host scalar (the
 *     reference's equivalents all fence: every Kokkos kernel is followed by
 *     Kokkos::fence()).
 *   - device memory is owned by the handles; the library never frees caller memory.
 *   - there is NO CPU fallback: without a CUDA device every call fails with
 *     MRMD_B200_ENODEVICE.
 */
#ifndef MRMD_B200_H
#define MRMD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRMD_B200_EINVAL (-1)
#define MRMD_B200_ENODEVICE (-2)
#define MRMD_B200_ECAPACITY (-3)
#define MRMD_B200_ENOMEM (-4)

typedef struct mrmd_b200_atoms mrmd_b200_atoms;         /* data::Atoms            (data/Atoms.hpp:33-147) */
typedef struct mrmd_b200_molecules mrmd_b200_molecules; /* data::Molecules        (data/Molecules.hpp:27-148) */
typedef struct mrmd_b200_verlet mrmd_b200_verlet;       /* Half/FullVerletList    (datatypes.hpp:184-191) */
typedef struct mrmd_b200_ghost mrmd_b200_ghost;         /* communication::GhostLayer / MultiResGhostLayer */
typedef struct mrmd_b200_lj mrmd_b200_lj;               /* action::LennardJones   (action/LennardJones.hpp:98-133) */
typedef struct mrmd_b200_adress mrmd_b200_adress;       /* action::LJ_IdealGas    (action/LJ_IdealGas.hpp:34-104) */
typedef struct mrmd_b200_thermo mrmd_b200_thermo;       /* action::ThermodynamicForce (ThermodynamicForce.hpp:32-96) */

/* data::Subdomain (data/Subdomain.hpp:37-110); fill with mrmd_b200_subdomain_init */
typedef struct
{
    double minCorner[3];
    double maxCorner[3];
    double ghostLayerThickness[3];
    double minGhostCorner[3];
    double maxGhostCorner[3];
    double minInnerCorner[3];
    double maxInnerCorner[3];
    double diameter[3];
    double diameterWithGhostLayer[3];
} mrmd_b200_subdomain;

/* Parametric stand-in for the reference's predicate lambdas (device lambdas cannot cross a C ABI):
 * util::IsInSymmetricSlab (util/IsInSymmetricSlab.hpp:24-64) alone or combined over two positions as in
 * examples/04_LennardJones_IdealGas_LocalCap.cpp:206-227. */
enum
{
    MRMD_B200_PRED_ALWAYS = 0,
    MRMD_B200_PRED_NEVER = 1,
    MRMD_B200_PRED_SLAB = 2,        /* slabMin-tol <= |x[axis]-center| <= slabMax+tol            */
    MRMD_B200_PRED_SLAB_EITHER = 3, /* two positions: slab(p1) || slab(p2)                       */
    MRMD_B200_PRED_SLAB_BOTH = 4,   /* two positions: slab(p1) && slab(p2)                       */
    MRMD_B200_PRED_INTERVAL = 5     /* slabMin < x[axis] < slabMax (ThermodynamicForce.test.cpp:29-40) */
};
typedef struct
{
    int32_t kind;
    int32_t axis;
    double center;
    double slabMin;
    double slabMax;
    double tolerance;
} mrmd_b200_pred;

/* weighting_function::Slab (Slab.hpp:27-201) / Spherical (Spherical.hpp:25-99, lambda^mod := lambda) */
enum
{
    MRMD_B200_WEIGHT_SLAB = 0,
    MRMD_B200_WEIGHT_SPHERICAL = 1
};
typedef struct
{
    int32_t kind;
    int32_t abrupt;   /* Slab::InterfaceType::ABRUPT */
    double center[3];
    double atRegion;  /* Slab: atomistic region DIAMETER; Spherical: atomistic RADIUS */
    double hyRegion;  /* hybrid region width */
    int64_t exponent; /* Slab: nu (exponent 2 nu); Spherical: exponent */
} mrmd_b200_weight;

enum
{
    MRMD_B200_ATOM_POS = 0,
    MRMD_B200_ATOM_VEL = 1,
    MRMD_B200_ATOM_FORCE = 2,
    MRMD_B200_ATOM_TYPE = 3, /* int64 */
    MRMD_B200_ATOM_MASS = 4,
    MRMD_B200_ATOM_CHARGE = 5,
    MRMD_B200_ATOM_RELATIVE_MASS = 6,
    /* int64; not a reference field: the global atom id that keys the Langevin noise (Philox counter), so that a run
       gives the same trajectory whatever the atom order and however many GPUs share it.  A new container numbers its
       atoms 0, 1, ...; the id travels with the atom through permute, ghost creation and slab migration. */
    MRMD_B200_ATOM_ID = 7
};
enum
{
    MRMD_B200_MOL_POS = 0,
    MRMD_B200_MOL_FORCE = 1,
    MRMD_B200_MOL_LAMBDA = 2,
    MRMD_B200_MOL_MODULATED_LAMBDA = 3,
    MRMD_B200_MOL_GRAD_LAMBDA = 4,
    MRMD_B200_MOL_ATOMS_OFFSET = 5, /* int64 */
    MRMD_B200_MOL_NUM_ATOMS = 6     /* int64 */
};
enum
{
    MRMD_B200_MEM_HOST = 0,
    MRMD_B200_MEM_DEVICE = 1
};

const char* mrmd_b200_last_error(void);
int mrmd_b200_device_count(void);
int mrmd_b200_set_device(int device);
int mrmd_b200_sync(void* stream);
/* number of kernels this library launched since load (bench.py's gpu_launches) */
int64_t mrmd_b200_launch_count(void);

void mrmd_b200_subdomain_init(mrmd_b200_subdomain* s, const double* minCorner, const double* maxCorner,
                              const double* ghostLayerThickness);
/* Subdomain::scaleDim (data/Subdomain.cpp:24-32) */
void mrmd_b200_subdomain_scale_dim(mrmd_b200_subdomain* s, double factor, int axis);

/* ---- data::Atoms ---------------------------------------------------------------------------- */
/* replaces GeneralAtoms(numAtoms) (data/Atoms.hpp:123-133): zero-filled container of `size` atoms */
int mrmd_b200_atoms_create(mrmd_b200_atoms** out, int64_t size);
int mrmd_b200_atoms_destroy(mrmd_b200_atoms* a);
/* replaces resize() (:89-93); contents are preserved, new entries are zero */
int mrmd_b200_atoms_resize(mrmd_b200_atoms* a, int64_t size, void* stream);
int mrmd_b200_atoms_reserve(mrmd_b200_atoms* a, int64_t capacity, void* stream);
int64_t mrmd_b200_atoms_size(const mrmd_b200_atoms* a);
/* numLocalAtoms / numGhostAtoms (:120-121) */
int mrmd_b200_atoms_set_counts(mrmd_b200_atoms* a, int64_t numLocal, int64_t numGhost);
int mrmd_b200_atoms_get_counts(const mrmd_b200_atoms* a, int64_t* numLocal, int64_t* numGhost);
/* slice transfer of atoms [first, first + count).  The caller's buffer starts at atom `first`: element
 * (j, d), j = 0 .. count-1 (atom first + j), sits at
 *   buf[(j / vlen) * stride + d * vlen + (j % vlen)]
 * which is a Cabana slice (data(), stride(0), vector length) offset to tuple `first` (first a multiple
 * of vlen); a dense (n, ncomp) array is
 * stride = ncomp, vlen = 1; the reference's default AoSoA (MRMD_VECTOR_LENGTH=1) is stride 13.
 * Replaces deep_copy host<->device (data/Atoms.hpp:150-159). */
int mrmd_b200_atoms_write(mrmd_b200_atoms* a, int field, const void* src, int64_t first, int64_t count,
                          int64_t stride, int64_t vlen, int memKind, void* stream);
int mrmd_b200_atoms_read(const mrmd_b200_atoms* a, int field, void* dst, int64_t first, int64_t count,
                         int64_t stride, int64_t vlen, int memKind, void* stream);
/* Cabana::deep_copy(slice, value) / Atoms::setForce (data/Atoms.hpp:72, examples/02:174-175) */
int mrmd_b200_atoms_fill(mrmd_b200_atoms* a, int field, double value, void* stream);
/* deep_copy(dst, src) between two device containers */
int mrmd_b200_atoms_copy(mrmd_b200_atoms* dst, const mrmd_b200_atoms* src, void* stream);

/* ---- data::Molecules ------------------------------------------------------------------------ */
int mrmd_b200_molecules_create(mrmd_b200_molecules** out, int64_t size);
int mrmd_b200_molecules_destroy(mrmd_b200_molecules* m);
int mrmd_b200_molecules_resize(mrmd_b200_molecules* m, int64_t size, void* stream);
int64_t mrmd_b200_molecules_size(const mrmd_b200_molecules* m);
int mrmd_b200_molecules_set_counts(mrmd_b200_molecules* m, int64_t numLocal, int64_t numGhost);
int mrmd_b200_molecules_get_counts(const mrmd_b200_molecules* m, int64_t* numLocal, int64_t* numGhost);
int mrmd_b200_molecules_write(mrmd_b200_molecules* m, int field, const void* src, int64_t first, int64_t count,
                              int64_t stride, int64_t vlen, int memKind, void* stream);
int mrmd_b200_molecules_read(const mrmd_b200_molecules* m, int field, void* dst, int64_t first, int64_t count,
                             int64_t stride, int64_t vlen, int memKind, void* stream);
int mrmd_b200_molecules_fill(mrmd_b200_molecules* m, int field, double value, void* stream);
/* replaces data::createMoleculeForEachAtom (data/MoleculesFromAtoms.cpp:19-39) */
int mrmd_b200_molecules_for_each_atom(mrmd_b200_molecules** out, const mrmd_b200_atoms* a, void* stream);

/* ---- integrators ---------------------------------------------------------------------------- */
/* replaces VelocityVerlet::preForceIntegrate (action/VelocityVerlet.cpp:26-67).  The maximum
 * displacement is accumulated in a device scalar; *maxDisplacement (host, may be NULL) receives
 * sqrt(max |dx|^2) after a stream sync, exactly what the reference returns. */
int mrmd_b200_vv_pre(mrmd_b200_atoms* a, double dt, double* maxDisplacement, void* stream);
/* replaces VelocityVerlet::postForceIntegrate (action/VelocityVerlet.cpp:69-90) */
int mrmd_b200_vv_post(mrmd_b200_atoms* a, double dt, void* stream);
/* replaces VelocityVerletLangevinThermostat::preForceIntegrate(_apply_if)
 * (action/VelocityVerletLangevinThermostat.hpp:48-52,63-136).  Random numbers: Philox4x32-10,
 * key = seed, counter = (atom index, step); the reference's XorShift1024 pool stream is scheduling
 * dependent and unpinned, so only the statistics are comparable. */
int mrmd_b200_langevin_pre(mrmd_b200_atoms* a, double dt, double zeta, double temperature, uint64_t seed,
                           uint64_t step, const mrmd_b200_pred* pred, double* maxDisplacement, void* stream);

/* ---- communication::GhostLayer -------------------------------------------------------------- */
int mrmd_b200_ghost_create(mrmd_b200_ghost** out);
int mrmd_b200_ghost_destroy(mrmd_b200_ghost* g);
/* replaces GhostLayer::exchangeRealAtoms -> PeriodicMapping::mapIntoDomain (PeriodicMapping.cpp:30-58) */
int mrmd_b200_ghost_map_into_domain(mrmd_b200_atoms* a, const mrmd_b200_subdomain* s, void* stream);
/* replaces GhostExchange::resetCorrespondingRealAtoms + createGhostAtoms(axis) (GhostExchange.cpp:59-169,183-187).
 * axis < 0: reset + X, Y, Z = createGhostAtomsXYZ (:171-181) = GhostLayer::createGhostAtoms.  Grows `a`. */
int mrmd_b200_ghost_create_atoms(mrmd_b200_ghost* g, mrmd_b200_atoms* a, const mrmd_b200_subdomain* s, int axis,
                                 void* stream);
int mrmd_b200_ghost_reset(mrmd_b200_ghost* g, mrmd_b200_atoms* a, void* stream);
/* replaces GhostLayer::updateGhostAtoms -> UpdateGhostAtoms::updateOnlyPos (UpdateGhostAtoms.cpp:31-68) */
int mrmd_b200_ghost_update(mrmd_b200_ghost* g, mrmd_b200_atoms* a, const mrmd_b200_subdomain* s, void* stream);
/* replaces GhostLayer::contributeBackGhostToReal -> AccumulateForce::ghostToReal (AccumulateForce.cpp:25-47) */
int mrmd_b200_ghost_contribute_back(mrmd_b200_ghost* g, mrmd_b200_atoms* a, void* stream);
/* correspondingRealAtom view (int64, -1 for real atoms); count entries starting at first */
int mrmd_b200_ghost_read_corresponding(const mrmd_b200_ghost* g, int64_t* dst, int64_t first, int64_t count,
                                       int memKind, void* stream);
int mrmd_b200_ghost_write_corresponding(mrmd_b200_ghost* g, const int64_t* src, int64_t first, int64_t count,
                                        int memKind, void* stream);
/* MultiResGhostLayer (communication/MultiResGhostLayer.hpp:29-61):
 * replaces realAtomsExchange (MultiResRealAtomsExchange.cpp:23-73) */
int mrmd_b200_ghost_mr_map_into_domain(mrmd_b200_molecules* m, mrmd_b200_atoms* a, const mrmd_b200_subdomain* s,
                                       void* stream);
/* replaces MultiResPeriodicGhostExchange::createGhostAtoms / XYZ (MultiResPeriodicGhostExchange.cpp:60-265) */
int mrmd_b200_ghost_mr_create_atoms(mrmd_b200_ghost* g, mrmd_b200_molecules* m, mrmd_b200_atoms* a,
                                    const mrmd_b200_subdomain* s, int axis, void* stream);

/* ---- Cabana::LinkedCellList + permute, Cabana::VerletList ----------------------------------- */
/* replaces LinkedCellList(pos, begin, end, delta, min, max) + atoms.permute(list)
 * (tests/NVT/NVT.cpp:136-144, data/Atoms.hpp:95): stable, atomic-free radix sort by cell index, all
 * members reordered.  cellIdOut (device, int32[end], optional) receives the cell index of every particle
 * BEFORE the permutation (for parity checks). */
int mrmd_b200_atoms_cell_sort(mrmd_b200_atoms* a, int64_t begin, int64_t end, const double* delta,
                              const double* gridMin, const double* gridMax, int32_t* cellIdOut, void* stream);
/* same for molecules (data/Molecules.hpp:91-95); atoms are not moved */
int mrmd_b200_molecules_cell_sort(mrmd_b200_molecules* m, int64_t begin, int64_t end, const double* delta,
                                  const double* gridMin, const double* gridMax, void* stream);

int mrmd_b200_verlet_create(mrmd_b200_verlet** out, int half);
int mrmd_b200_verlet_destroy(mrmd_b200_verlet* v);
/* replaces VerletList::build(pos, begin, end, radius, cellRatio, gridMin, gridMax, maxNeigh)
 * (call sites examples/02_LennardJones_NVE.cpp:156-163, tests/LennardJones/LennardJones.cpp:112-118,
 * action/LJ_IdealGas.test.cpp:113-120).  All size() particles are candidates, rows exist for [begin,end).
 * If a row overflows maxNeigh the table is widened and refilled (as Cabana does). */
int mrmd_b200_verlet_build_atoms(mrmd_b200_verlet* v, const mrmd_b200_atoms* a, int64_t begin, int64_t end,
                                 double radius, double cellRatio, const double* gridMin, const double* gridMax,
                                 int64_t maxNeigh, void* stream);
int mrmd_b200_verlet_build_molecules(mrmd_b200_verlet* v, const mrmd_b200_molecules* m, int64_t begin, int64_t end,
                                     double radius, double cellRatio, const double* gridMin, const double* gridMax,
                                     int64_t maxNeigh, void* stream);
/* B200 fast path for periodic subdomains: equivalent to GhostLayer::createGhostAtoms followed by
 * VerletList::build(pos, 0, numLocalAtoms, radius, cellRatio, minGhostCorner, maxGhostCorner, maxNeigh) as far as
 * the local atoms are concerned -- the same pairs (partner = a local atom or one of its periodic images, which
 * are exactly the reference's ghost atoms) -- but built on shared-memory tiles of the cell-sorted local atoms.
 * Requires LinkedCellList + permute over [0, numLocalAtoms) with the subdomain corners as grid (tests/NVT/NVT.cpp:136-144)
 * since the last change of the atom order.  The list is stored as 16-bit tile slots; mrmd_b200_lj_apply
 * recognises it and runs the tiled force kernel (full list: forces on row owners only). */
int mrmd_b200_verlet_build_periodic(mrmd_b200_verlet* v, const mrmd_b200_atoms* a, const mrmd_b200_subdomain* s,
                                    double radius, double cellRatio, int64_t maxNeigh, void* stream);
/* decodes a periodic list to Cabana's row-major layout: partner[i][n] = local index of the partner,
 * shiftCode[i][n] = (sx+1) + 3 (sy+1) + 9 (sz+1) with the image shift s_d in {-1,0,+1} (x subdomain.diameter);
 * host buffers of numLocal (x width) int32, -1 padded */
int mrmd_b200_verlet_read_periodic(const mrmd_b200_verlet* v, const mrmd_b200_atoms* a, int32_t* countsHost,
                                   int32_t* partnerHost, int32_t* shiftCodeHost, void* stream);
/* The periodic list for molecules of atomsPerMolecule consecutive atoms (atomsOffset = atomsPerMolecule * m): the reference
 * builds the molecule list on the centres of mass over MultiResGhostLayer ghosts; here
 * mrmd_b200_molecules_cell_sort_with_atoms is LinkedCellList + permute on the centres of mass of the local molecules (the
 * atoms move in blocks of atomsPerMolecule, all members; the centres of mass move along), and the list is built on
 * tiles of the sorted centres of mass with the periodic images generated on the fly, like
 * mrmd_b200_verlet_build_periodic.  UpdateMolecules::update must have filled the centres of mass before the sort. */
int mrmd_b200_molecules_cell_sort_with_atoms(mrmd_b200_molecules* m, mrmd_b200_atoms* a, int atomsPerMolecule,
                                             const double* delta, const double* gridMin, const double* gridMax, void* stream);
int mrmd_b200_verlet_build_periodic_molecules(mrmd_b200_verlet* v, mrmd_b200_molecules* m, const mrmd_b200_subdomain* s,
                                              double radius, double cellRatio, int64_t maxNeigh, int atomsPerMolecule,
                                              void* stream);
int mrmd_b200_verlet_read_periodic_molecules(const mrmd_b200_verlet* v, const mrmd_b200_molecules* m, int32_t* countsHost,
                                             int32_t* partnerHost, int32_t* shiftCodeHost, void* stream);
/* list._data.counts / neighbors as a Cabana VerletLayout2D table: counts[numParticles] int32 and
 * neighbors[numParticles][width] int32 (row-major).  Pass NULL to query sizes only. */
int mrmd_b200_verlet_info(const mrmd_b200_verlet* v, int64_t* numParticles, int64_t* width, int64_t* totalPairs,
                          int* half);
int mrmd_b200_verlet_read(const mrmd_b200_verlet* v, int32_t* counts, int32_t* neighbors, int memKind, void* stream);

/* ---- action::LennardJones ------------------------------------------------------------------- */
/* replaces LennardJones(cappingDistance[], rc[], sigma[], epsilon[], numTypes, isShifted)
 * (action/LennardJones.cpp:46-56, :81-116); arrays hold numTypes^2 entries.
 * Reference quirk kept on purpose: LennardJones::numTypes_ is initialised to 1 (LennardJones.cpp:52), so apply() indexes
 * the table with  type_i * 1 + type_j : with more than one type, pair (1, 1) reads the entry of (1, 0) and so on
 * (LJ_IdealGas indexes type_i * numTypes + type_j and does not have the quirk).  mrmd_b200_lj_create reproduces that for
 * parity and leaves a note in mrmd_b200_last_error() when numTypes > 1; single-type systems are unaffected.
 * Handles are single-stream objects: an operator handle (lj, adress, thermo, verlet, md, slab) owns its reduction
 * scratch, ticket counter and result buffers, so two calls on the same handle must be ordered on one stream (or by an
 * event); different handles may run on different streams concurrently. */
int mrmd_b200_lj_create(mrmd_b200_lj** out, const double* cappingDistance, const double* rc, const double* sigma,
                        const double* epsilon, int64_t numTypes, int isShifted);
int mrmd_b200_lj_destroy(mrmd_b200_lj* lj);
/* replaces LennardJones::apply / apply_if (action/LennardJones.hpp:135-206).  pred == NULL: apply.
 * With a half list forces are scattered to both partners (ghosts included, fold them back with
 * mrmd_b200_ghost_contribute_back); with a full list only row owners receive force and energy/virial
 * are halved (extension, same post-fold-back result).  Energy and virial are kept on the device until
 * mrmd_b200_lj_get is called. */
int mrmd_b200_lj_apply(mrmd_b200_lj* lj, mrmd_b200_atoms* a, const mrmd_b200_verlet* v, const mrmd_b200_pred* pred,
                       void* stream);
/* getEnergy()/getVirial() (action/LennardJones.cpp:35-36); numPairs = pairs that reached the force
 * evaluation in the last apply (the pair-interactions/s numerator).  Syncs the stream. */
int mrmd_b200_lj_get(mrmd_b200_lj* lj, double* energy, double* virial, int64_t* numPairs, void* stream);
/* CappedLennardJonesPotential::computeForceAndEnergy on the device for n squared distances (unit tests) */
int mrmd_b200_lj_eval(const mrmd_b200_lj* lj, int64_t typeIdx, const double* distSqrHost, int64_t n,
                      double* forceFactorHost, double* energyHost, void* stream);

/* ---- AdResS --------------------------------------------------------------------------------- */
/* replaces UpdateMolecules::update (action/UpdateMolecules.hpp:24-70) */
int mrmd_b200_molecules_update(mrmd_b200_molecules* m, const mrmd_b200_atoms* a, const mrmd_b200_weight* w,
                               void* stream);
/* replaces ContributeMoleculeForceToAtoms::update (action/ContributeMoleculeForceToAtoms.cpp:23-48) */
int mrmd_b200_molecules_contribute_force(const mrmd_b200_molecules* m, mrmd_b200_atoms* a, void* stream);
/* evaluates the weighting function on the device for n host positions (unit tests) */
int mrmd_b200_weight_eval(const mrmd_b200_weight* w, const double* posHost, int64_t n, double* lambdaHost,
                          double* modLambdaHost, double* gradHost, void* stream);
/* replaces LJ_IdealGas(...) ctors (action/LJ_IdealGas.cpp:262-292) */
int mrmd_b200_adress_create(mrmd_b200_adress** out, const double* cappingDistance, const double* rc,
                            const double* sigma, const double* epsilon, int64_t numTypes, int doShift);
int mrmd_b200_adress_destroy(mrmd_b200_adress* ad);
/* setCompensationEnergySamplingInterval / UpdateInterval (action/LJ_IdealGas.hpp:73-80) */
int mrmd_b200_adress_set_intervals(mrmd_b200_adress* ad, int64_t samplingInterval, int64_t updateInterval);
/* optional promise that every molecule has exactly atomsPerMolecule atoms (0: unknown, the default).  4 selects a kernel
 * with four lanes per molecule (the tetramers of BASELINE.json configs[3]); a molecule that breaks the promise makes the
 * next call that reads results back fail. */
int mrmd_b200_adress_set_atoms_per_molecule(mrmd_b200_adress* ad, int64_t atomsPerMolecule);
/* replaces LJ_IdealGas::run (action/LJ_IdealGas.cpp:227-260); *energy (host, optional) after a sync */
int mrmd_b200_adress_run(mrmd_b200_adress* ad, mrmd_b200_molecules* m, const mrmd_b200_verlet* v,
                         mrmd_b200_atoms* a, double* energy, int64_t* numPairs, void* stream);
/* B200 fast path for one-atom molecules (data::createMoleculeForEachAtom, relativeMass 1) on a list from
 * mrmd_b200_verlet_build_periodic: UpdateMolecules::update (the weighting function is evaluated inline at the
 * atom = molecule position, images at their image position), LJ_IdealGas::run and
 * ContributeMoleculeForceToAtoms::update in one kernel without atomics on forces; the atoms' force is
 * accumulated (+=).  Requires the AT + HY region to stay at least the ghost layer thickness away from the
 * periodic faces the weight depends on (x for Slab, all for Spherical); otherwise MRMD_B200_EINVAL. */
int mrmd_b200_adress_run_periodic(mrmd_b200_adress* ad, mrmd_b200_atoms* a, const mrmd_b200_verlet* v,
                                  const mrmd_b200_weight* w, double* energy, int64_t* numPairs, void* stream);
/* The same fast path for molecules of atomsPerMolecule (4: the tetramers of BASELINE.json configs[3]) consecutive atoms
 * with atomsOffset = atomsPerMolecule * m: v is a list from mrmd_b200_verlet_build_periodic_molecules on the centres of
 * mass in m (UpdateMolecules::update must have run: pos holds them), the force kernel stages all atoms of a molecule
 * per list slot, evaluates the weight at the (image) centre of mass and runs LJ_IdealGas::run (LJ_IdealGas.cpp:96-225)
 * and ContributeMoleculeForceToAtoms::update from the row owner's side: no atomics on forces, no ghost molecules. */
int mrmd_b200_adress_run_periodic_molecules(mrmd_b200_adress* ad, const mrmd_b200_molecules* m, mrmd_b200_atoms* a,
                                            const mrmd_b200_verlet* v, const mrmd_b200_weight* w, int atomsPerMolecule,
                                            double* energy, int64_t* numPairs, void* stream);
/* getMeanCompensationEnergy() and the two accumulation histograms, 200 x numTypes doubles each
 * (kind 0 mean, 1 compensationEnergy, 2 compensationEnergyCounter) */
int mrmd_b200_adress_read_histogram(const mrmd_b200_adress* ad, int kind, double* dstHost, void* stream);

/* ---- data::MultiHistogram ---------------------------------------------------------------------
 * data/MultiHistogram.hpp:29-194, MultiHistogram.cpp:27-227: numBins x numHistograms doubles over [min, max) on the
 * device, data(bin, histogram) row-major.  Every entry point replaces the member / free function it is named after. */
typedef struct mrmd_b200_hist mrmd_b200_hist;
/* MultiHistogram(label, min, max, numBins, numHistograms): zero filled (:31-47) */
int mrmd_b200_hist_create(mrmd_b200_hist** out, double min, double max, int64_t numBins, int64_t numHistograms);
/* MultiHistogram(label, histogram): deep copy (:49-54) */
int mrmd_b200_hist_clone(mrmd_b200_hist** out, const mrmd_b200_hist* src, void* stream);
int mrmd_b200_hist_destroy(mrmd_b200_hist* h);
/* min, max, numBins, numHistograms, binSize, inverseBinSize (:76-81); NULL skips a field */
int mrmd_b200_hist_info(const mrmd_b200_hist* h, double* min, double* max, int64_t* numBins, int64_t* numHistograms,
                        double* binSize, double* inverseBinSize);
/* the MultiView `data`: device pointer, and copies from / to a dense row-major buffer (memKind as for the slices) */
void* mrmd_b200_hist_device_data(mrmd_b200_hist* h);
int mrmd_b200_hist_write(mrmd_b200_hist* h, const double* src, int memKind, void* stream);
int mrmd_b200_hist_read(const mrmd_b200_hist* h, double* dst, int memKind, void* stream);
/* getBin (:60-66, -1 outside the range) and getBinPosition (:68-74): host arithmetic */
int64_t mrmd_b200_hist_get_bin(const mrmd_b200_hist* h, double val);
double mrmd_b200_hist_get_bin_position(const mrmd_b200_hist* h, int64_t binIdx);
/* operator+= (op 0), -= (1), *= (2), /= (3) (MultiHistogram.cpp:27-46) */
int mrmd_b200_hist_transform(mrmd_b200_hist* h, const mrmd_b200_hist* rhs, int op, void* stream);
/* scale(real_t) (:48-59) and scale(ScalarView): one factor per histogram, numFactors >= numHistograms (:61-74) */
int mrmd_b200_hist_scale(mrmd_b200_hist* h, double factor, void* stream);
int mrmd_b200_hist_scale_per_histogram(mrmd_b200_hist* h, const double* factorsHost, int64_t numFactors, void* stream);
/* makeSymmetric (:76-90) */
int mrmd_b200_hist_make_symmetric(mrmd_b200_hist* h, void* stream);
/* cumulativeMovingAverage(average, current, factor) (:92-111) */
int mrmd_b200_hist_cumulative_moving_average(mrmd_b200_hist* average, const mrmd_b200_hist* current,
                                             double movingAverageFactor, void* stream);
/* gradient(input, periodic) (:113-161) and smoothen(input, sigma, range, periodic) (:163-212): a new histogram */
int mrmd_b200_hist_gradient(mrmd_b200_hist** out, const mrmd_b200_hist* in, int periodic, void* stream);
int mrmd_b200_hist_smoothen(mrmd_b200_hist** out, const mrmd_b200_hist* in, double sigma, double range, int periodic,
                            void* stream);
/* replace_if_bin_position(hist, pred, newValue) (MultiHistogram.hpp:176-191); the predicate is parametric (its axis
 * coordinate is the bin position): device lambdas cannot cross a C ABI */
int mrmd_b200_hist_replace_if_bin_position(mrmd_b200_hist* h, const mrmd_b200_pred* pred, double newValue, void* stream);
/* createGrid (:214-227): the numBins bin positions, to the host */
int mrmd_b200_hist_create_grid(const mrmd_b200_hist* h, double* gridHost, void* stream);

/* ---- action::ThermodynamicForce ------------------------------------------------------------- */
/* replaces the ctor (action/ThermodynamicForce.cpp:25-57) */
int mrmd_b200_thermo_create(mrmd_b200_thermo** out, const double* targetDensity, int64_t numTypes,
                            const mrmd_b200_subdomain* s, double requestedDensityBinWidth,
                            const double* modulation, int enforceSymmetry, int usePeriodicity);
int mrmd_b200_thermo_destroy(mrmd_b200_thermo* t);
int mrmd_b200_thermo_info(const mrmd_b200_thermo* t, int64_t* numBins, int64_t* numTypes, double* binSize,
                          int64_t* samples);
/* replaces sample() (ThermodynamicForce.cpp:74-86) = analysis::getAxialDensityProfile (AxialDensityProfile.cpp:21-51) */
int mrmd_b200_thermo_sample(mrmd_b200_thermo* t, const mrmd_b200_atoms* a, void* stream);
/* replaces update / update_if (ThermodynamicForce.hpp:124-152, .cpp:88-91); pred on the bin centre */
int mrmd_b200_thermo_update(mrmd_b200_thermo* t, double smoothingSigma, double smoothingIntensity,
                            const mrmd_b200_pred* pred, void* stream);
/* replaces apply / apply_if (ThermodynamicForce.hpp:98-122) and applyInterpolated_if (:154-216) */
int mrmd_b200_thermo_apply(const mrmd_b200_thermo* t, mrmd_b200_atoms* a, const mrmd_b200_pred* pred,
                           int interpolated, void* stream);
/* getForce()/setForce()/getDensityProfile(): numBins x numTypes doubles, host buffers (kind 0 force, 1 density) */
int mrmd_b200_thermo_read(const mrmd_b200_thermo* t, int kind, double* dstHost, void* stream);
/* getForce() (kind 0) / getDensityProfile() (kind 1) as the data::MultiHistogram the reference returns
 * (ThermodynamicForce.hpp:57-63): a copy over [minCorner_x, maxCorner_x) with numTypes histograms */
int mrmd_b200_thermo_get_hist(const mrmd_b200_thermo* t, int kind, mrmd_b200_hist** out, void* stream);
int mrmd_b200_thermo_write_force(mrmd_b200_thermo* t, const double* srcHost, void* stream);
/* all-reduce hooks for the x-slab decomposition: raw device pointer of the density histogram */
int mrmd_b200_thermo_density_ptr(mrmd_b200_thermo* t, double** devicePtr, int64_t* count);
/* getMuLeft / getMuRight (ThermodynamicForce.cpp:98-130) */
int mrmd_b200_thermo_mu(const mrmd_b200_thermo* t, double* muLeftHost, double* muRightHost, void* stream);

/* ---- analysis:: diagnostics of the drivers' statistics lines (examples/02:190-199) ------------- */
typedef struct mrmd_b200_msd mrmd_b200_msd; /* analysis::MeanSquareDisplacement (MeanSquareDisplacement.hpp:24-49) */
/* replaces analysis::getKineticEnergy (analysis/KineticEnergy.hpp:26-39): 0.5 sum m v^2 over the local atoms;
 * getMeanKineticEnergy (:44-47) is this / numLocalAtoms */
int mrmd_b200_kinetic_energy(const mrmd_b200_atoms* a, double* kineticEnergy, void* stream);
/* replaces analysis::getSystemMomentum (analysis/SystemMomentum.cpp:21-50): the sum of the local atoms'
 * VELOCITIES per component (the reference does not weight by mass) */
int mrmd_b200_system_momentum(const mrmd_b200_atoms* a, double* momentum3, void* stream);
/* replaces analysis::getPressure (analysis/Pressure.cpp:23-51): sum over local AND ghost atoms of
 * m v^2 + F . x, divided by 3 V */
int mrmd_b200_pressure(const mrmd_b200_atoms* a, const mrmd_b200_subdomain* s, double* pressure, void* stream);
int mrmd_b200_msd_create(mrmd_b200_msd** out);
int mrmd_b200_msd_destroy(mrmd_b200_msd* m);
/* replaces MeanSquareDisplacement::reset (MeanSquareDisplacement.cpp:23-57) */
int mrmd_b200_msd_reset_atoms(mrmd_b200_msd* m, const mrmd_b200_atoms* a, void* stream);
int mrmd_b200_msd_reset_molecules(mrmd_b200_msd* m, const mrmd_b200_molecules* mol, void* stream);
/* replaces MeanSquareDisplacement::calc (:59-113): |dx| folded by one box length when larger than half of it;
 * items are matched by index, MRMD_B200_EINVAL if their number changed since reset */
int mrmd_b200_msd_calc_atoms(const mrmd_b200_msd* m, const mrmd_b200_atoms* a, const mrmd_b200_subdomain* s,
                             double* meanSquareDisplacement, void* stream);
int mrmd_b200_msd_calc_molecules(const mrmd_b200_msd* m, const mrmd_b200_molecules* mol, const mrmd_b200_subdomain* s,
                                 double* meanSquareDisplacement, void* stream);

/* ---- thermostat / barostat / constraints around the loop (tests/NVT, tests/NPT, tests/Constraints) ---- */
typedef struct mrmd_b200_constraints mrmd_b200_constraints; /* action::MoleculeConstraints (action/Shake.hpp:159-251) */
/* replaces BerendsenThermostat::apply (action/BerendsenThermostat.cpp:25-50): v *= sqrt(1 + gamma (T_target / T - 1))
 * for the local atoms; no-op for T <= 0 */
int mrmd_b200_berendsen_thermostat(mrmd_b200_atoms* a, double currentTemperature, double targetTemperature, double gamma,
                                   void* stream);
/* replaces BerendsenBarostat::apply (action/BerendsenBarostat.cpp:23-50): mu = cbrt(1 + gamma (P - P_target)) scales
 * the chosen axes of the subdomain (Subdomain::scaleDim) and of the local atoms' positions */
int mrmd_b200_berendsen_barostat(mrmd_b200_atoms* a, double currentPressure, double targetPressure, double gamma,
                                 mrmd_b200_subdomain* s, int stretchX, int stretchY, int stretchZ, void* stream);
/* replace limitAccelerationPerComponent (action/LimitAcceleration.cpp:21-45) and limitVelocityPerComponent
 * (action/LimitVelocity.cpp:23-43): per-component clamps of force / mass and of the velocity of the local atoms (the
 * headers every reference driver includes) */
int mrmd_b200_limit_acceleration(mrmd_b200_atoms* a, double maxAccelerationPerComponent, void* stream);
int mrmd_b200_limit_velocity(mrmd_b200_atoms* a, double maxVelocityPerComponent, void* stream);
/* MoleculeConstraints(atomsPerMolecule, numConstraintIterations) (:247-250) */
int mrmd_b200_constraints_create(mrmd_b200_constraints** out, int64_t atomsPerMolecule, int64_t numConstraintIterations);
int mrmd_b200_constraints_destroy(mrmd_b200_constraints* c);
/* setConstraints (:236-245): bonds {idx, jdx, eqDistance} (data/Bond.hpp:23-28), indices relative to the molecule's
 * first atom; host arrays */
int mrmd_b200_constraints_set(mrmd_b200_constraints* c, const int64_t* idx, const int64_t* jdx, const double* eqDistance,
                              int64_t numBonds);
/* replaces enforcePositionalConstraints (:167-201): numConstraintIterations x (unconstrained update of all atoms,
 * SHAKE force correction per bond of every local molecule, impl::Shake :84-137) */
int mrmd_b200_constraints_enforce_positional(mrmd_b200_constraints* c, const mrmd_b200_molecules* m, mrmd_b200_atoms* a,
                                             double dt, void* stream);
/* replaces enforceVelocityConstraints (:203-233): RATTLE velocity projection per bond (impl::Shake :56-82) */
int mrmd_b200_constraints_enforce_velocity(mrmd_b200_constraints* c, const mrmd_b200_molecules* m, mrmd_b200_atoms* a,
                                           double dt, void* stream);

/* ---- SPC water and the Coulomb pair potentials (action/SPC.hpp, Coulomb.hpp, CoulombDSF.hpp) ------------------- */
typedef struct mrmd_b200_spc mrmd_b200_spc; /* action::SPC (action/SPC.hpp:61-362) */
#define MRMD_B200_COULOMB_PLAIN 0 /* impl::Coulomb    (action/Coulomb.hpp:27-46): the member SPC holds */
#define MRMD_B200_COULOMB_DSF 1   /* impl::CoulombDSF (action/CoulombDSF.hpp:42-84), approxErfc of util/math.hpp:57-76 */
/* computeForce / computeEnergy of the chosen potential for n squared distances (host arrays), evaluated on the device;
 * rc and alpha are the CoulombDSF constructor's arguments (ignored for the plain potential) */
int mrmd_b200_coulomb_eval(int kind, double rc, double alpha, const double* distSqrHost, int64_t n, double q1, double q2,
                           double* forceHost, double* energyHost, void* stream);
/* SPC() (:346-362): capped (0.7 sigma), shifted O-O Lennard-Jones with rc = 1.2 nm, the three bonds H-O, H-O, H-H.
 * coulombKind MRMD_B200_COULOMB_PLAIN reproduces the reference; MRMD_B200_COULOMB_DSF evaluates the charges with
 * CoulombDSF(rc, alpha = 2 / nm) (the parameters SPC.hpp:110 declares) instead. */
int mrmd_b200_spc_create(mrmd_b200_spc** out, int coulombKind);
int mrmd_b200_spc_destroy(mrmd_b200_spc* spc);
/* replaces SPC::applyForces(molecules, verletList, atoms) (:252-282, kernel :143-236): for every pair of the half
 * Verlet list of molecules O-O Lennard-Jones between the first atoms (distSqr < rc^2) and Coulomb between all atom
 * pairs (skipped for distSqr > rc^2), accumulated into the atoms' forces; getEnergyLJ / getEnergyCoulomb come back
 * through the pointers (either may be NULL; both NULL skips the read-back and the fence) */
int mrmd_b200_spc_apply_forces(mrmd_b200_spc* spc, const mrmd_b200_molecules* m, const mrmd_b200_verlet* v,
                               mrmd_b200_atoms* a, double* energyLJ, double* energyCoulomb, void* stream);
/* replaces SPC::calcBondEnergy (:284-344): harmonicPreFactor * sum over local + ghost molecules of the squared
 * deviations of |O-H0|, |O-H1|, |H0-H1| from their equilibrium lengths / (local + ghost atoms) */
int mrmd_b200_spc_calc_bond_energy(mrmd_b200_spc* spc, const mrmd_b200_molecules* m, const mrmd_b200_atoms* a,
                                   double harmonicPreFactor, double* bondEnergy, void* stream);
/* replace SPC::enforcePositionalConstraints / enforceVelocityConstraints (:238-250): MoleculeConstraints(3, 20) */
int mrmd_b200_spc_enforce_positional_constraints(mrmd_b200_spc* spc, const mrmd_b200_molecules* m, mrmd_b200_atoms* a,
                                                 double dt, void* stream);
int mrmd_b200_spc_enforce_velocity_constraints(mrmd_b200_spc* spc, const mrmd_b200_molecules* m, mrmd_b200_atoms* a,
                                               double dt, void* stream);

/* ---- step loop of the reference's drivers ----------------------------------------------------
 * The hot loop of examples/02_LennardJones_NVE.cpp:135-216 (rebuild policy :141-171), with the Langevin
 * integrator of examples/01_LennardJones_NVT.cpp:121,142 and the LinkedCellList + permute spatial sort of
 * tests/NVT/NVT.cpp:136-144 at every rebuild, as one host-side C++ driver so that the only per-step
 * host<->device traffic is the displacement scalar the reference reads too.  AdResS mode assembles the
 * step of SURVEY.md section 3.5 (UpdateMolecules -> LJ_IdealGas -> ThermodynamicForce ->
 * ContributeMoleculeForceToAtoms -> MultiResGhostLayer) from the same operators. */
typedef struct mrmd_b200_md mrmd_b200_md;
typedef struct
{
    double dt;
    double rc, skin;                /* neighborCutoff = rc + skin */
    double sigma, epsilon, cappingDistance;
    int64_t maxNeighbors;           /* estimatedMaxNeighbors */
    int32_t integrator;             /* 0 VelocityVerlet, 1 VelocityVerletLangevinThermostat */
    int32_t cellSort;               /* 1: LinkedCellList + permute at every rebuild (tests/NVT) */
    int32_t fullList;               /* 0: HalfVerletList (reference), 1: FullVerletList over ghost atoms,
                                       2: tiled periodic list (mrmd_b200_verlet_build_periodic), no ghost atoms */
    int32_t adress;                 /* 0: LennardJones::apply, 1: AdResS step (one molecule per atom) */
    double zeta, temperature;       /* Langevin: gamma and T */
    uint64_t seed;
    /* AdResS only */
    mrmd_b200_weight weight;
    int32_t doShift;
    int32_t useThermoForce;
    double thermoTargetDensity, thermoBinWidth, thermoModulation;
    int64_t thermoSampleInterval, thermoUpdateInterval;
    double thermoSmoothingSigma, thermoSmoothingIntensity;
    /* AdResS with the half list (fullList 0) only: molecules of atomsPerMolecule consecutive atoms (atomsOffset =
     * atomsPerMolecule * m, relative masses taken from the atoms; 0 or 1: one molecule per atom) -- the tetramers of
     * BASELINE.json configs[3].  numConstraintIterations > 0 adds MoleculeConstraints(atomsPerMolecule,
     * numConstraintIterations) with a bond of length bondLength between every two atoms of a molecule: SHAKE in front
     * of preForceIntegrate, RATTLE after postForceIntegrate (tests/Constraints/Constraints.cpp:53-64).  cellSort must
     * be 0 for multi-atom molecules. */
    int64_t atomsPerMolecule;
    int64_t numConstraintIterations;
    double bondLength;
    /* 0: energy and virial are reduced on the last step of a run only (they are not observable earlier; the force and
     * the pair count are the same); 1: on every step, like LennardJones::apply (LennardJones.hpp:187-188) */
    int32_t energyEveryStep;
    int32_t reserved0;
} mrmd_b200_md_config;
typedef struct
{
    int64_t steps;          /* steps executed by this call */
    int64_t rebuilds;       /* neighbour rebuilds in this call */
    int64_t pairInteractions; /* pairs that reached the force evaluation, summed over the steps */
    int64_t storedPairs;    /* stored list pairs summed over the steps (roofline P) */
    int64_t numLocal, numGhost;
    double energy, virial;  /* of the last step */
    double forceKernelMs;   /* CUDA-event time of the force kernel summed over the steps (0 if not timed) */
    double maxDisplacement;
    int64_t activePairs;    /* AdResS: stored half pairs that were not skipped as CG-CG, summed over the steps */
} mrmd_b200_md_stats;

int mrmd_b200_md_create(mrmd_b200_md** out, const mrmd_b200_md_config* cfg, const mrmd_b200_subdomain* s,
                        mrmd_b200_atoms* atoms);
int mrmd_b200_md_destroy(mrmd_b200_md* md);
/* nsteps device-resident steps.  timeForceKernel != 0 brackets every force launch with CUDA events. */
int mrmd_b200_md_run(mrmd_b200_md* md, int64_t nsteps, int timeForceKernel, mrmd_b200_md_stats* stats,
                     void* stream);
/* nsteps steps through HOST buffers: every step copies pos and vel (numLocal x 3 doubles each, pinned or
 * pageable) to the device, runs one step and copies pos, vel and {energy, virial, maxDisplacement} back. */
int mrmd_b200_md_run_host(mrmd_b200_md* md, int64_t nsteps, double* posHost, double* velHost,
                          double* scalarsHost, mrmd_b200_md_stats* stats, void* stream);
/* mrmd_b200_md_config::energyEveryStep of an existing driver (LennardJones.hpp:187-188 reduces on every apply) */
int mrmd_b200_md_set_energy_every_step(mrmd_b200_md* md, int enabled);
/* ---- x-slab decomposition over the GPUs of one node ---------------------------------------------
 * New functionality (the reference's communication layer is single-process periodic self-ghosting,
 * communication/MultiResRealAtomsExchange.hpp:26): one process per GPU, rank r owns the slab
 * [xmin + r W, xmin + (r+1) W) of the global box.  y / z stay locally periodic, the x pass of
 * GhostExchange::createGhostAtoms (communication/GhostExchange.cpp:59-169) becomes an NCCL halo over NVLink:
 * full records migrate at a rebuild, positions of the face atoms are sent every step, no reverse force halo
 * (full list).  The displacement criterion of examples/02:141-143 is evaluated on the ncclAllReduce(max).
 * uniqueId128: the 128 bytes of mrmd_b200_nccl_unique_id from rank 0, broadcast by the caller. */
typedef struct mrmd_b200_slab mrmd_b200_slab;
int mrmd_b200_nccl_unique_id(void* out128);
int mrmd_b200_slab_create(mrmd_b200_slab** out, const mrmd_b200_md_config* cfg, const double* globalMin,
                          const double* globalMax, int rank, int nranks, const void* uniqueId128, mrmd_b200_atoms* atoms,
                          void* stream);
/* same with caller-chosen slab boundaries: cuts[0] = globalMin[0] < cuts[1] < ... < cuts[nranks] = globalMax[0], rank r
 * owns [cuts[r], cuts[r+1]); every slab at least rc + skin wide.  For cost-balanced slabs (narrow over the AT / HY
 * region, wide over the coarse-grained region).  cuts == NULL: equal widths. */
int mrmd_b200_slab_create_cuts(mrmd_b200_slab** out, const mrmd_b200_md_config* cfg, const double* globalMin,
                               const double* globalMax, const double* cuts, int rank, int nranks,
                               const void* uniqueId128, mrmd_b200_atoms* atoms, void* stream);
int mrmd_b200_slab_destroy(mrmd_b200_slab* sl);
/* nsteps collective steps; stats: energy, virial and pairInteractions are summed over the ranks (a pair across
 * a slab face counts one half on either side), the other fields are per rank */
int mrmd_b200_slab_run(mrmd_b200_slab* sl, int64_t nsteps, int timeForceKernel, mrmd_b200_md_stats* stats,
                       void* stream);
int mrmd_b200_slab_set_energy_every_step(mrmd_b200_slab* sl, int enabled);
/* mrmd_b200_md_run_host on a slab (collective): every step copies this rank's pos and vel (numLocal x 3 doubles each,
 * in the rank's current atom order) from the host buffers, runs one step and copies pos, vel and scalarsHost =
 * {this rank's energy, virial, maxDisplacement, numLocal after the step} back.  The buffers must hold the largest
 * number of resident atoms (atoms migrate at a rebuild); stats: steps, rebuilds, storedPairs, numLocal, numGhost. */
int mrmd_b200_slab_run_host(mrmd_b200_slab* sl, int64_t nsteps, double* posHost, double* velHost, double* scalarsHost,
                            mrmd_b200_md_stats* stats, void* stream);

/* pinned host memory for the host-buffer path */
int mrmd_b200_host_alloc(void** ptr, int64_t bytes);
int mrmd_b200_host_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* MRMD_B200_H */
