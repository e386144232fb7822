// handles.cuh -- host-side state behind the opaque operator handles of include/mrmd_b200.h.
#pragma once

#include "common.cuh"

struct mrmd_b200_ghost
{
    int64_t* corr = nullptr;  // correspondingRealAtom, -1 for real atoms
    int64_t corrCapacity = 0;
    mrmd_b200::DevBuf blockCounts;  // int64[4 * numBlocks]
    int64_t* dTotals = nullptr;     // int64[4]
    int64_t* hTotals = nullptr;     // pinned
};

struct mrmd_b200_lj
{
    int64_t numTypes = 1;
    int64_t numTypesQuirk = 1;  // LennardJones::numTypes_ is hard-wired to 1 (LennardJones.cpp:52)
    double rcSqr = 0.0;
    mrmd_b200::LJTable table{};
    mrmd_b200::DevBuf partials;  // double[3 * maxBlocks]
    double* dResult = nullptr;   // [0..2] energy, virial, pairs of the last apply; [3..5] running sums
    unsigned int* dTicket = nullptr;
    double* hResult = nullptr;  // pinned, 6 doubles
};

struct mrmd_b200_adress
{
    int64_t numTypes = 1;
    double rcSqr = 0.0;
    mrmd_b200::LJTable table{};
    int64_t runCounter = 0;
    int64_t samplingInterval = 200;  // LJ_IdealGas.hpp:69
    int64_t updateInterval = 20000;  // :70
    double* hist = nullptr;          // 3 x (200 x numTypes): compensationEnergy, counter, mean
    // x-slab decomposition: called on {compensationEnergy, counter} (count doubles) right before the mean update
    // so that the ranks can sum their samples
    int (*preUpdateHook)(void* ctx, double* sums, int64_t count, cudaStream_t st) = nullptr;
    void* hookCtx = nullptr;
    mrmd_b200::DevBuf partials;
    mrmd_b200::DevBuf activeList;  // int32[numLocalMolecules] + the count: molecules with work in the current run
    double* dResult = nullptr;  // [0..2] energy, pairs, - of the last run; [3..5] running sums
    unsigned int* dTicket = nullptr;
    double* hResult = nullptr;
    // mrmd_b200_adress_set_atoms_per_molecule: 4 selects the four-lanes-per-molecule kernel; a molecule with another
    // atom count raises *dErr, reported by the next call that reads results back
    int64_t uniformAtoms = 0;
    int* dErr = nullptr;
    int* hErr = nullptr;  // pinned
};

struct mrmd_b200_thermo
{
    double min = 0, max = 0;
    int64_t numBins = 0, numTypes = 0;
    double binSize = 0, inverseBinSize = 0;
    double binVolume = 0;
    int64_t samples = 0;
    int enforceSymmetry = 0, usePeriodicity = 0;
    double* force = nullptr;        // numBins x numTypes
    double* density = nullptr;      // numBins x numTypes
    double* tmpA = nullptr;
    double* tmpB = nullptr;
    double* forceFactor = nullptr;  // numTypes
};

namespace mrmd_b200
{
// integrate.cu: preForceIntegrate, optionally with the previous step's postForceIntegrate fused in front; the
// squared maximum displacement is left in a->dMaxDisp
// stop (device int, optional): the kernel returns at once when it is set -- for steps queued ahead of the host
int integratePre(mrmd_b200_atoms* a, double dt, bool langevin, double zeta, double temperature, uint64_t seed,
                 uint64_t step, const mrmd_b200_pred* pred, bool fusedPost, cudaStream_t st, const int* stop = nullptr);
// integrate.cu: the displacement criterion of examples/02:138-143 evaluated on the device (see there)
int displacementDecision(const double* dMaxDispSqr, double* dAccum, double threshold, int* dStop, int localStep, cudaStream_t st);
// containers.cu: dense (n x ncomp) device buffer <-> one field of the container
int atomsFieldToDense(const mrmd_b200_atoms* a, int field, double* devBuf, int64_t n, cudaStream_t st);
int atomsFieldFromDense(mrmd_b200_atoms* a, int field, const double* devBuf, int64_t n, cudaStream_t st);
// tiled.cu: LennardJones::apply over a tiled (periodic, shared-memory staged) full list
int ljApplyTiled(mrmd_b200_lj* lj, mrmd_b200_atoms* a, const mrmd_b200_verlet* v, bool accumulate, bool energy,
                 cudaStream_t st, const int* stop = nullptr);
int adressApplyTiled(mrmd_b200_adress* ad, mrmd_b200_atoms* a, const mrmd_b200_verlet* v, const mrmd_b200_weight* w,
                     bool sampling, bool energy, cudaStream_t st, const int* stop = nullptr);
// tiled.cu / adress.cu: LJ_IdealGas over a tiled list of the centres of mass of molecules of atomsPerMolecule atoms
int moleculeApplyTiled(mrmd_b200_adress* ad, const mrmd_b200_molecules* m, mrmd_b200_atoms* a, const mrmd_b200_verlet* v,
                       const mrmd_b200_weight* w, int atomsPerMolecule, bool sampling, bool energy, cudaStream_t st);
int adressRunPeriodicMolecules(mrmd_b200_adress* ad, const mrmd_b200_molecules* m, mrmd_b200_atoms* a,
                               const mrmd_b200_verlet* v, const mrmd_b200_weight* w, int atomsPerMolecule, bool energy,
                               cudaStream_t st);
// tiled.cu: the tiled list on the cell-sorted centres of mass (moleculesCellSortWithAtoms); halo arguments as verletBuildTiled
int verletBuildTiledMolecules(mrmd_b200_verlet* v, mrmd_b200_molecules* m, const mrmd_b200_subdomain* s, double radius,
                              double cellRatio, int64_t maxNeigh, int atomsPerMolecule, const int32_t* haloLeft,
                              const int32_t* haloRight, cudaStream_t st);
// adress.cu: mrmd_b200_adress_run_periodic with the energy accumulation optional (the drivers want it on the last
// step of a run only)
// adress.cu: reads the atoms-per-molecule flag of the four-lane kernel back (synchronises the stream)
int adressCheckUniform(mrmd_b200_adress* ad, cudaStream_t st);
int adressRunPeriodic(mrmd_b200_adress* ad, mrmd_b200_atoms* a, const mrmd_b200_verlet* v, const mrmd_b200_weight* w,
                      bool energy, cudaStream_t st, const int* stop = nullptr);
// constraints.cu: SHAKE / RATTLE launches without the bond-range read-back and its stream synchronisation
int constraintsEnforcePositional(mrmd_b200_constraints* c, const mrmd_b200_molecules* m, mrmd_b200_atoms* a, double dt,
                                 cudaStream_t st);
int constraintsEnforceVelocity(mrmd_b200_constraints* c, const mrmd_b200_molecules* m, mrmd_b200_atoms* a, cudaStream_t st,
                               double postDt);
// promise of the step-loop drivers: uniform molecules that own every local atom (lets postForceIntegrate ride along)
void constraintsSetUniformMolecules(mrmd_b200_constraints* c, bool uniform);
// neighbor.cu: LinkedCellList + permute on the centres of mass of molecules of apm consecutive atoms (see there)
int moleculesCellSortWithAtoms(mrmd_b200_molecules* m, mrmd_b200_atoms* a, int64_t count, int apm, const double* delta,
                               const double* gridMin, const double* gridMax, const signed char* dropFlags, cudaStream_t st);
int verletBuildTiled(mrmd_b200_verlet* v, const mrmd_b200_atoms* a, const mrmd_b200_subdomain* s, double radius,
                     double cellRatio, int64_t maxNeigh, const int32_t* haloLeft, const int32_t* haloRight,
                     cudaStream_t st);
}  // namespace mrmd_b200
