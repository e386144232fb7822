/*
 * mrmd_oracle.h -- C interface of the CPU parity oracle.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  This is a plain C++20/OpenMP
 * restatement of the MRMD per-step force + neighbour path (reference:
 * XzzX/mrmd, files cited per function in mrmd_oracle.cpp).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product (mrmd_b200/) never links or calls it.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   - MRMD-owned arithmetic (LJ, AdResS, integrators, ghost layer, weighting
 *     functions, histograms): pinned by the reference's golden vectors
 *     (tests/golden/, tests/test_oracle_golden.py).
 *   - Cabana 0.7 (7914d28) Verlet/linked-cell arithmetic is NOT in the
 *     reference tree: restated from its published algorithm.  The half-list
 *     criterion and pair counts are pinned by the 1 310 403-pair golden; the
 *     inclusive list cutoff (<=), the grid arithmetic and in-cell ordering are
 *     "parity unpinned".  The Langevin random stream is parity unpinned as well
 *     (Kokkos XorShift1024 pool is scheduling dependent); we define Philox4x32-10,
 *     pinned statistically by the reference's thermostat integration test
 *     (tests/test_oracle_integration.py).
 *
 * Layouts follow the reference's default build (MRMD_VECTOR_LENGTH=1, i.e. an
 * array of 104-byte records, mrmd/data/Atoms.hpp:47-53, Molecules.hpp:41-47).
 */
#ifndef MRMD_ORACLE_H
#define MRMD_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct
{
    double pos[3];
    double vel[3];
    double force[3];
    int64_t type;
    double mass;
    double charge;
    double relMass;
} or_atom_t; /* 104 bytes */

typedef struct
{
    double pos[3];
    double force[3];
    double lambda;
    double modLambda;
    double gradLambda[3];
    int64_t atomsOffset;
    int64_t numAtoms;
} or_molecule_t; /* 104 bytes */

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
} or_subdomain_t;

typedef struct
{
    double ff1, ff2, ef1, ef2;
    double rcSqr;
    double cappingDistance, cappingDistanceSqr, cappingCoeff;
    double shift;
    double energyAtCappingPoint;
} or_lj_type_t; /* 80 bytes, LennardJones.hpp:27-39 */

/* parametric stand-in for the reference's predicate lambdas */
enum
{
    OR_PRED_ALWAYS = 0,
    OR_PRED_NEVER = 1,
    OR_PRED_SLAB = 2,        /* one position: IsInSymmetricSlab */
    OR_PRED_SLAB_EITHER = 3, /* two positions: slab(p1) || slab(p2) */
    OR_PRED_SLAB_BOTH = 4,   /* two positions: slab(p1) && slab(p2) */
    OR_PRED_INTERVAL = 5     /* one position: slabMin < x[axis] < slabMax (open; the TestPredicate of
                                ThermodynamicForce.test.cpp:29-40) */
};
typedef struct
{
    int32_t kind;
    int32_t axis;
    double center;
    double slabMin;
    double slabMax;
    double tolerance;
} or_pred_t;

enum
{
    OR_WEIGHT_SLAB = 0,
    OR_WEIGHT_SPHERICAL = 1
};
typedef struct
{
    int32_t kind;
    int32_t abrupt;      /* Slab::InterfaceType::ABRUPT */
    double center[3];
    double atRegion;     /* Slab: atomistic region DIAMETER; Spherical: atomistic RADIUS */
    double hyRegion;     /* hybrid region width */
    int64_t exponent;    /* Slab: nu (exponent = 2 nu); Spherical: exponent */
} or_weight_t;

void or_set_threads(int n);
int or_get_max_threads(void);

void or_subdomain_init(or_subdomain_t* s, const double* minCorner, const double* maxCorner, const double* thickness);
void or_subdomain_scale_dim(or_subdomain_t* s, double factor, int axis);

/* --- Lennard-Jones ------------------------------------------------------ */
void or_lj_init(or_lj_type_t* out, const double* cappingDistance, const double* rc, const double* sigma,
                const double* epsilon, int64_t numTypes, int isShifted);
void or_lj_force_energy(const or_lj_type_t* table, int64_t typeIdx, double distSqr, double* forceFactor,
                        double* energy);
/* LennardJones::apply_if; numTypesQuirk is the value multiplied with type(i) (the
 * reference hard-wires 1, LennardJones.cpp:52).  Returns #pairs that reached the
 * force evaluation. */
int64_t or_lj_apply(or_atom_t* atoms, int64_t numLocal, const int32_t* counts, const int32_t* neigh,
                    int64_t width, const or_lj_type_t* table, double rcSqr, int64_t numTypesQuirk,
                    const or_pred_t* pred, double* energyVirial);

/* --- Cabana-style neighbour structures ---------------------------------- */
/* cell id of every particle in [begin,end) (others untouched); returns #cells, fills dims[3] */
int64_t or_cell_ids(const double* pos, int64_t stride, int64_t begin, int64_t end, const double* delta,
                    const double* gmin, const double* gmax, int32_t* cellId, int32_t* dims);
/* stable counting sort by cell id: perm[k] = source index (relative to 0) for slot begin+k;
 * cellOffsets has numCells+1 entries */
void or_cell_perm(const int32_t* cellId, int64_t begin, int64_t end, int64_t numCells, int64_t* perm,
                  int64_t* cellOffsets);
void or_permute_atoms(or_atom_t* atoms, int64_t begin, int64_t end, const int64_t* perm);
void or_permute_molecules(or_molecule_t* mols, int64_t begin, int64_t end, const int64_t* perm);
/* VerletList build. half=1 HalfNeighborTag, 0 FullNeighborTag. counts has nAll entries,
 * neigh nAll*width.  Returns the maximum row count (if > width the table is
 * truncated: caller re-runs with a wider table, like Cabana's refill). */
int64_t or_verlet_build(const double* pos, int64_t stride, int64_t nAll, int64_t begin, int64_t end,
                        double radius, double ratio, const double* gmin, const double* gmax, int half,
                        int64_t width, int32_t* counts, int32_t* neigh);
/* brute-force pair counter of the reference's tests (tests/LennardJones/LennardJones.cpp:40-70) */
int64_t or_count_within_cutoff(const double* pos, int64_t stride, int64_t numLocal, int64_t numAll,
                               double cutoff, const double* box, int periodic);

/* --- ghost layer --------------------------------------------------------- */
void or_periodic_map(or_atom_t* atoms, int64_t numLocal, const or_subdomain_t* s);
/* appends ghosts behind numLocal (capacity must suffice: capacity checked, returns -1 if not);
 * corr has capacity entries.  Returns numGhost. */
int64_t or_ghost_create_axis(or_atom_t* atoms, int64_t numLocal, int64_t numGhost, int64_t capacity,
                             const or_subdomain_t* s, int axis, int64_t* corr);
int64_t or_ghost_create_xyz(or_atom_t* atoms, int64_t numLocal, int64_t capacity, const or_subdomain_t* s,
                            int64_t* corr);
void or_ghost_update_pos(or_atom_t* atoms, int64_t numLocal, int64_t numGhost, const int64_t* corr,
                         const or_subdomain_t* s);
void or_ghost_fold_force(or_atom_t* atoms, int64_t numLocal, int64_t numGhost, const int64_t* corr);
void or_zero_force(or_atom_t* atoms, int64_t n);

/* multi-resolution (molecule granular) variants */
void or_mr_periodic_map(or_molecule_t* mols, int64_t numLocalMols, or_atom_t* atoms, const or_subdomain_t* s);
/* out[0]=numGhostMolecules out[1]=numGhostAtoms; returns 0 or -1 when capacity is exceeded */
int or_mr_ghost_create_axis(or_molecule_t* mols, int64_t numLocalMols, int64_t numGhostMols, int64_t molCapacity,
                            or_atom_t* atoms, int64_t numLocalAtoms, int64_t numGhostAtoms, int64_t atomCapacity,
                            const or_subdomain_t* s, int axis, int64_t* corrAtoms, int64_t* out);
int or_mr_ghost_create_xyz(or_molecule_t* mols, int64_t numLocalMols, int64_t molCapacity, or_atom_t* atoms,
                           int64_t numLocalAtoms, int64_t atomCapacity, const or_subdomain_t* s,
                           int64_t* corrAtoms, int64_t* out);

/* --- integrators --------------------------------------------------------- */
double or_vv_pre(or_atom_t* atoms, int64_t numLocal, double dt);
void or_vv_post(or_atom_t* atoms, int64_t numLocal, double dt);
/* Langevin BAOAB pre-force step; random numbers: Philox4x32-10, key=(seed lo,hi),
 * counter=(idx lo, idx hi, step lo, step hi), 32-bit uniforms, Box-Muller. */
double or_langevin_pre(or_atom_t* atoms, int64_t numLocal, double dt, double zeta, double temperature,
                       uint64_t seed, uint64_t step, const or_pred_t* pred);
/* the same with the counter taken from ids[idx] (global atom ids; NULL: the index) */
double or_langevin_pre_ids(or_atom_t* atoms, int64_t numLocal, double dt, double zeta, double temperature,
                           uint64_t seed, uint64_t step, const or_pred_t* pred, const int64_t* ids);
void or_philox_normals(uint64_t seed, uint64_t step, uint64_t idx, double* out4);
void or_philox4x32(const uint32_t* ctr, const uint32_t* key, uint32_t* out);

/* --- AdResS --------------------------------------------------------------- */
void or_weight_eval(const or_weight_t* w, double x, double y, double z, double* lambda, double* modLambda,
                    double* grad);
void or_update_molecules(or_molecule_t* mols, int64_t numAllMols, const or_atom_t* atoms, const or_weight_t* w);
void or_contribute_molecule_force(const or_molecule_t* mols, int64_t numAllMols, or_atom_t* atoms);

typedef struct
{
    int64_t numTypes;
    double rcSqr;
    int64_t numBins;        /* 200 */
    int64_t runCounter;
    int64_t samplingInterval; /* 200 */
    int64_t updateInterval;   /* 20000 */
    or_lj_type_t* table;      /* numTypes^2 */
    double* compensationEnergy;        /* numBins x numTypes */
    double* compensationEnergyCounter; /* numBins x numTypes */
    double* meanCompensationEnergy;    /* numBins x numTypes */
} or_adress_t;
or_adress_t* or_adress_create(const double* cappingDistance, const double* rc, const double* sigma,
                              const double* epsilon, int64_t numTypes, int doShift);
void or_adress_destroy(or_adress_t* a);
/* LJ_IdealGas::run; returns the energy; numActive (optional) receives #atom pairs evaluated */
double or_adress_run(or_adress_t* a, or_molecule_t* mols, int64_t numLocalMols, const int32_t* counts,
                     const int32_t* neigh, int64_t width, or_atom_t* atoms, int64_t* numActive);

/* --- MultiHistogram / thermodynamic force ---------------------------------- */
int64_t or_hist_get_bin(double min, double max, int64_t numBins, double val);
void or_hist_scale(double* data, int64_t numBins, int64_t numHist, double factor);
void or_hist_scale_per_hist(double* data, int64_t numBins, int64_t numHist, const double* factors);
void or_hist_make_symmetric(double* data, int64_t numBins, int64_t numHist);
void or_hist_gradient(const double* in, double* out, double min, double max, int64_t numBins, int64_t numHist,
                      int periodic);
void or_hist_smoothen(const double* in, double* out, double min, double max, int64_t numBins, int64_t numHist,
                      double sigma, double range, int periodic);
void or_limit_acceleration(or_atom_t* atoms, int64_t numLocal, double maxAccelerationPerComponent);
void or_limit_velocity(or_atom_t* atoms, int64_t numLocal, double maxVelocityPerComponent);
void or_berendsen_thermostat(or_atom_t* atoms, int64_t numLocal, double currentTemperature, double targetTemperature,
                             double gamma);
void or_berendsen_barostat(or_atom_t* atoms, int64_t numLocal, double currentPressure, double targetPressure, double gamma,
                           or_subdomain_t* s, int stretchX, int stretchY, int stretchZ);
int or_shake_positional(const or_molecule_t* mols, int64_t numLocalMols, or_atom_t* atoms, int64_t numAllAtoms,
                        const int64_t* bondIdx, const double* eqDistance, int64_t numBonds, int64_t numIterations,
                        double dt);
int or_shake_velocity(const or_molecule_t* mols, int64_t numLocalMols, or_atom_t* atoms, const int64_t* bondIdx,
                      int64_t numBonds);
/* util/math.hpp:57-76, action/Coulomb.hpp:27-46 (kind 0), action/CoulombDSF.hpp:42-84 (kind 1), action/SPC.hpp:143-344 */
double or_approx_erfc(double x);
void or_coulomb_eval(int kind, double rc, double alpha, const double* distSqr, int64_t n, double q1, double q2,
                     double* force, double* energy);
void or_spc_apply_forces(const or_molecule_t* mols, int64_t numLocalMols, const int32_t* counts, const int32_t* neigh,
                         int64_t width, or_atom_t* atoms, int coulombKind, double* energies);
double or_spc_bond_energy(const or_molecule_t* mols, int64_t numAllMols, const or_atom_t* atoms, int64_t numAllAtoms,
                          double harmonicPreFactor);
double or_kinetic_energy(const or_atom_t* atoms, int64_t numLocal);
void or_system_momentum(const or_atom_t* atoms, int64_t numLocal, double* out3);
double or_pressure(const or_atom_t* atoms, int64_t numAll, const or_subdomain_t* s);
double or_msd(const or_atom_t* atoms, const double* initialPos, int64_t numItems, const or_subdomain_t* s);
void or_density_profile(const or_atom_t* atoms, int64_t numAtoms, int64_t numTypes, double min, double max,
                        int64_t numBins, int axis, double* hist);

typedef struct
{
    double min, max;
    int64_t numBins, numTypes;
    double binSize, inverseBinSize;
    double binVolume;
    int64_t samples;
    int enforceSymmetry, usePeriodicity;
    double* force;       /* numBins x numTypes */
    double* density;     /* numBins x numTypes */
    double* forceFactor; /* numTypes */
} or_thermo_t;
or_thermo_t* or_thermo_create(const double* targetDensity, int64_t numTypes, const or_subdomain_t* s,
                              double requestedBinWidth, const double* modulation, int enforceSymmetry,
                              int usePeriodicity);
void or_thermo_destroy(or_thermo_t* t);
void or_thermo_sample(or_thermo_t* t, const or_atom_t* atoms, int64_t numLocal);
void or_thermo_update(or_thermo_t* t, double smoothingSigma, double smoothingIntensity, const or_pred_t* pred);
void or_thermo_apply(const or_thermo_t* t, or_atom_t* atoms, int64_t numLocal, const or_pred_t* pred,
                     int interpolated);
void or_thermo_mu(const or_thermo_t* t, double* muLeft, double* muRight);

#ifdef __cplusplus
}
#endif
#endif
