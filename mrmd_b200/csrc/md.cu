// md.cu -- the step loop of the reference's drivers as one host-side C++ driver over the operators of
// this library (no new arithmetic).  Reference: examples/02_LennardJones_NVE/02_LennardJones_NVE.cpp:135-216
// (rebuild policy :141-171), examples/01_LennardJones_NVT/01_LennardJones_NVT.cpp:121,142 (Langevin),
// tests/NVT/NVT.cpp:136-144 (LinkedCellList + permute at rebuild), and the AdResS step assembled from the
// unit-test usage of the operators.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "handles.cuh"
#include "hostpipe.cuh"

struct mrmd_b200_md
{
    mrmd_b200_md_config cfg{};
    mrmd_b200_subdomain sub{};
    mrmd_b200_atoms* atoms = nullptr;  // not owned
    mrmd_b200_molecules* mols = nullptr;
    mrmd_b200_ghost* ghost = nullptr;
    mrmd_b200_verlet* list = nullptr;
    mrmd_b200_lj* lj = nullptr;
    mrmd_b200_adress* adress = nullptr;
    mrmd_b200_thermo* thermo = nullptr;
    mrmd_b200_constraints* constraints = nullptr;  // multi-atom molecules with numConstraintIterations > 0
    double maxDisplacement = DBL_MAX;  // examples/02:110
    int64_t step = 0;
    int64_t rebuilds = 0;
    int64_t storedPairsNow = 0;
    double active0 = 0.0;  // running active-pair count when the current run started
    bool postPending = false;  // the last step's postForceIntegrate is fused into the next preForceIntegrate
    std::vector<cudaEvent_t> events;
    mrmd_b200::HostPipe hp;  // host-buffer path (mrmd_b200_md_run_host)
    // steps queued ahead of the host (runQueued): the displacement criterion is evaluated on the device
    int* dStop = nullptr;      // {stop, local step that stopped, non-finite displacement}
    double* dAccum = nullptr;  // accumulated displacement (the device copy of maxDisplacement)
    int* hStop = nullptr;      // pinned: dStop, and behind it the accumulated displacement
    int64_t stepsSinceRebuild = 0, lastRebuildInterval = 4;
};

namespace mrmd_b200
{
__global__ void moleculePerAtomInitKernel(MolsView m, int64_t n, int64_t atomsPerMolecule)
{
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i < n) m.oc[i] = make_longlong2(i * atomsPerMolecule, atomsPerMolecule);
}

static int rebuild(mrmd_b200_md* md, cudaStream_t st)
{
    const mrmd_b200_md_config& c = md->cfg;
    mrmd_b200_atoms* a = md->atoms;
    const double cutoff = c.rc + c.skin;
    // LinkedCellList gridDelta of the spatial sort: one list radius along x and y, a quarter along z (atoms of a
    // cell column end up in fine z order, which the tiled neighbour build exploits)
    const double delta[3] = {cutoff, cutoff, 0.25 * cutoff};
    if (c.adress)
    {
        // MultiResGhostLayer::exchangeRealAtoms (needs the current centres of mass),
        // optional spatial sort, MultiResGhostLayer::createGhostAtoms, molecule Verlet list on the COMs.
        mrmd_b200_molecules* m = md->mols;
        m->numGhost = 0;
        m->size = m->numLocal;
        a->numGhost = 0;
        a->size = a->numLocal;
        MB_TRY(mrmd_b200_molecules_update(m, a, &c.weight, st));
        MB_TRY(mrmd_b200_ghost_mr_map_into_domain(m, a, &md->sub, st));
        if (c.fullList == 2 && c.atomsPerMolecule > 1)
        {
            // fast path for molecules of atomsPerMolecule consecutive atoms: the centres of mass of the wrapped atoms
            // (UpdateMolecules::update, as the reference recomputes them before its list build), LinkedCellList +
            // permute on them with the atoms moving in blocks, the tiled list on the sorted centres of mass
            const int apm = static_cast<int>(c.atomsPerMolecule);
            MB_TRY(mrmd_b200_molecules_update(m, a, &c.weight, st));
            MB_TRY(moleculesCellSortWithAtoms(m, a, m->numLocal, apm, delta, md->sub.minCorner, md->sub.maxCorner, nullptr, st));
            MB_TRY(verletBuildTiledMolecules(md->list, m, &md->sub, cutoff, 1.0, c.maxNeighbors, apm, nullptr, nullptr, st));
            int64_t total = 0;
            MB_TRY(mrmd_b200_verlet_info(md->list, nullptr, nullptr, &total, nullptr));
            md->storedPairsNow = total;
            md->rebuilds += 1;
            return 0;
        }
        if (c.fullList == 2)
        {
            // fast path: one-atom molecules are their atoms, the tiled list over the sorted atoms is the molecule
            // list, images are generated while staging: no ghost molecules / atoms are materialised
            MB_TRY(mrmd_b200_atoms_cell_sort(a, 0, a->numLocal, delta, md->sub.minCorner, md->sub.maxCorner, nullptr, st));
            MB_TRY(mrmd_b200_verlet_build_periodic(md->list, a, &md->sub, cutoff, 1.0, c.maxNeighbors, st));
            int64_t total = 0;
            MB_TRY(mrmd_b200_verlet_info(md->list, nullptr, nullptr, &total, nullptr));
            md->storedPairsNow = total;
            md->rebuilds += 1;
            return 0;
        }
        if (c.cellSort)
        {
            // one atom per molecule (data::createMoleculeForEachAtom): sorting the atoms sorts the molecules
            MB_TRY(mrmd_b200_atoms_cell_sort(a, 0, a->numLocal, delta, md->sub.minCorner, md->sub.maxCorner, nullptr, st));
            MB_TRY(mrmd_b200_molecules_update(m, a, &c.weight, st));
        }
        MB_TRY(mrmd_b200_ghost_mr_create_atoms(md->ghost, m, a, &md->sub, -1, st));
        MB_TRY(mrmd_b200_molecules_update(m, a, &c.weight, st));
        MB_TRY(mrmd_b200_verlet_build_molecules(md->list, m, 0, m->numLocal, cutoff, 1.0, md->sub.minGhostCorner,
                                                md->sub.maxGhostCorner, c.maxNeighbors, st));
    }
    else
    {
        MB_TRY(mrmd_b200_ghost_map_into_domain(a, &md->sub, st));  // ghostLayer.exchangeRealAtoms (examples/02:150)
        a->numGhost = 0;
        a->size = a->numLocal;
        if (c.cellSort || c.fullList == 2)  // tests/NVT/NVT.cpp:136-144; the tiled list is built from the fresh sort
            MB_TRY(mrmd_b200_atoms_cell_sort(a, 0, a->numLocal, delta, md->sub.minCorner, md->sub.maxCorner, nullptr, st));
        if (c.fullList == 2)
        {
            // fast path: the same pair set as the list over local + ghost atoms, built on shared-memory tiles
            // of the freshly sorted local atoms with the periodic images generated on the fly (tiled.cu); ghost
            // atoms are never materialised: the container holds the local atoms only
            MB_TRY(mrmd_b200_verlet_build_periodic(md->list, a, &md->sub, cutoff, 1.0, c.maxNeighbors, st));
            int64_t total = 0;
            MB_TRY(mrmd_b200_verlet_info(md->list, nullptr, nullptr, &total, nullptr));
            md->storedPairsNow = total;
            md->rebuilds += 1;
            return 0;
        }
        MB_TRY(mrmd_b200_ghost_create_atoms(md->ghost, a, &md->sub, -1, st));  // examples/02:153
        MB_TRY(mrmd_b200_verlet_build_atoms(md->list, a, 0, a->numLocal, cutoff, 1.0, md->sub.minGhostCorner,
                                            md->sub.maxGhostCorner, c.maxNeighbors, st));  // examples/02:156-163
    }
    int64_t total = 0;
    MB_TRY(mrmd_b200_verlet_info(md->list, nullptr, nullptr, &total, nullptr));
    md->storedPairsNow = total;
    md->rebuilds += 1;
    return 0;
}

// one step; evStart/evStop (optional) bracket the force kernel
static int postIntegrate(mrmd_b200_md* md, bool deferPost, cudaStream_t st)
{
    if (md->constraints != nullptr)
    {
        // RATTLE needs the kicked velocities: no deferral (tests/Constraints/Constraints.cpp:62-64); the kick rides
        // along in the RATTLE kernel when the molecules fit it
        return constraintsEnforceVelocity(md->constraints, md->mols, md->atoms, st, md->cfg.dt);
    }
    if (deferPost)
    {
        md->postPending = true;
        return 0;
    }
    return mrmd_b200_vv_post(md->atoms, md->cfg.dt, st);
}

static int stepAfterPre(mrmd_b200_md* md, cudaStream_t st, cudaEvent_t evStart, cudaEvent_t evStop, bool deferPost,
                        bool wantEnergy, cudaEvent_t evPosReady, bool rebuildNow, const int* stop);

// deferPost: leave postForceIntegrate to the next step's fused kernel (the caller flushes it when the run ends)
static int oneStep(mrmd_b200_md* md, cudaStream_t st, cudaEvent_t evStart, cudaEvent_t evStop, bool deferPost,
                   bool wantEnergy, cudaEvent_t evPosReady = nullptr)
{
    const mrmd_b200_md_config& c = md->cfg;
    mrmd_b200_atoms* a = md->atoms;
    if (md->constraints != nullptr)  // tests/Constraints/Constraints.cpp:53-54
        MB_TRY(constraintsEnforcePositional(md->constraints, md->mols, a, c.dt, st));
    MB_TRY(integratePre(a, c.dt, c.integrator == 1, c.zeta, c.temperature, c.seed, uint64_t(md->step), nullptr,
                        md->postPending, st));
    md->postPending = false;
    MB_CUDA(cudaMemcpyAsync(a->hMaxDisp, a->dMaxDisp, 8, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    MB_REQUIRE(std::isfinite(*a->hMaxDisp), "md_run: non-finite position, velocity or force (the system blew up)");
    md->maxDisplacement += std::sqrt(*a->hMaxDisp);  // examples/02:138, VelocityVerlet.cpp:66
    return stepAfterPre(md, st, evStart, evStop, deferPost, wantEnergy, evPosReady, md->maxDisplacement >= c.skin * 0.5, nullptr);
}

// the step behind preForceIntegrate: rebuild or ghost refresh, force, postForceIntegrate.  stop (device, optional): the
// step is queued ahead of the host and its kernels return at once when an earlier queued step asked for a rebuild.
static int stepAfterPre(mrmd_b200_md* md, cudaStream_t st, cudaEvent_t evStart, cudaEvent_t evStop, bool deferPost,
                        bool wantEnergy, cudaEvent_t evPosReady, bool rebuildNow, const int* stop)
{
    const mrmd_b200_md_config& c = md->cfg;
    mrmd_b200_atoms* a = md->atoms;
    if (rebuildNow)  // :141-143
    {
        md->maxDisplacement = 0.0;
        if (md->stepsSinceRebuild > 0) md->lastRebuildInterval = md->stepsSinceRebuild;
        md->stepsSinceRebuild = 0;
        MB_TRY(rebuild(md, st));
    }
    else if (c.fullList != 2)
    {
        MB_TRY(mrmd_b200_ghost_update(md->ghost, a, &md->sub, st));  // :170
        if (c.adress) MB_TRY(mrmd_b200_molecules_update(md->mols, a, &c.weight, st));
    }
    if (evPosReady != nullptr) MB_CUDA(cudaEventRecord(evPosReady, st));  // positions and atom order are final
    if (c.fullList == 2 && c.adress && c.atomsPerMolecule > 1)
    {
        // tiled AdResS step for multi-atom molecules: centres of mass, then LJ_IdealGas + ContributeMoleculeForceToAtoms
        // as one kernel over the molecule list (mrmd_b200_adress_run_periodic_molecules)
        MB_TRY(mrmd_b200_molecules_update(md->mols, a, &c.weight, st));
        MB_TRY(mrmd_b200_atoms_fill(a, MRMD_B200_ATOM_FORCE, 0.0, st));
        if (evStart) MB_CUDA(cudaEventRecord(evStart, st));
        MB_TRY(adressRunPeriodicMolecules(md->adress, md->mols, a, md->list, &c.weight, static_cast<int>(c.atomsPerMolecule),
                                          wantEnergy, st));
        if (evStop) MB_CUDA(cudaEventRecord(evStop, st));
        MB_TRY(postIntegrate(md, deferPost, st));
        md->step += 1;
        return 0;
    }
    if (c.fullList == 2 && c.adress)
    {
        // tiled AdResS step: thermodynamic force on the zeroed force, then UpdateMolecules + LJ_IdealGas +
        // ContributeMoleculeForceToAtoms as one kernel (mrmd_b200_adress_run_periodic)
        MB_TRY(mrmd_b200_atoms_fill(a, MRMD_B200_ATOM_FORCE, 0.0, st));
        if (md->thermo != nullptr)
        {
            if (c.thermoSampleInterval > 0 && md->step % c.thermoSampleInterval == 0)
                MB_TRY(mrmd_b200_thermo_sample(md->thermo, a, st));
            if (c.thermoUpdateInterval > 0 && md->step > 0 && md->step % c.thermoUpdateInterval == 0 &&
                md->thermo->samples > 0)
                MB_TRY(mrmd_b200_thermo_update(md->thermo, c.thermoSmoothingSigma, c.thermoSmoothingIntensity, nullptr, st));
            MB_TRY(mrmd_b200_thermo_apply(md->thermo, a, nullptr, 0, st));
        }
        if (evStart) MB_CUDA(cudaEventRecord(evStart, st));
        MB_TRY(adressRunPeriodic(md->adress, a, md->list, &c.weight, wantEnergy, st, stop));
        if (evStop) MB_CUDA(cudaEventRecord(evStop, st));
        MB_TRY(postIntegrate(md, deferPost, st));
        md->step += 1;
        md->stepsSinceRebuild += 1;
        return 0;
    }
    if (c.fullList == 2)
    {
        // tiled fast path: images come from the staged local atoms, the force is stored (not accumulated) and
        // no ghost atom receives force: no ghost refresh, no force reset, no fold-back inside the loop
        if (evStart) MB_CUDA(cudaEventRecord(evStart, st));
        // energy and virial are only observable after the run returns: accumulate them on its last step
        MB_TRY(ljApplyTiled(md->lj, a, md->list, false, wantEnergy, st, stop));
        {
            // measurement knob (profiles/r02_build_experiments.md): the force is stored, not accumulated, so repeating the
            // launch changes nothing but the time -- the marginal cost of the kernel inside the step loop
            static const int repeat = std::getenv("MRMD_B200_REPEAT_FORCE") ? std::atoi(std::getenv("MRMD_B200_REPEAT_FORCE")) : 1;
            for (int r = 1; r < repeat; ++r) MB_TRY(ljApplyTiled(md->lj, a, md->list, false, wantEnergy, st, stop));
        }
        if (evStop) MB_CUDA(cudaEventRecord(evStop, st));
        MB_TRY(postIntegrate(md, deferPost, st));
        md->step += 1;
        md->stepsSinceRebuild += 1;
        return 0;
    }
    MB_TRY(mrmd_b200_atoms_fill(a, MRMD_B200_ATOM_FORCE, 0.0, st));  // :174-175
    if (c.adress)
    {
        MB_TRY(mrmd_b200_molecules_fill(md->mols, MRMD_B200_MOL_FORCE, 0.0, st));
        if (evStart) MB_CUDA(cudaEventRecord(evStart, st));
        MB_TRY(mrmd_b200_adress_run(md->adress, md->mols, md->list, a, nullptr, nullptr, st));
        if (evStop) MB_CUDA(cudaEventRecord(evStop, st));
        if (md->thermo != nullptr)
        {
            if (c.thermoSampleInterval > 0 && md->step % c.thermoSampleInterval == 0)
                MB_TRY(mrmd_b200_thermo_sample(md->thermo, a, st));
            if (c.thermoUpdateInterval > 0 && md->step > 0 && md->step % c.thermoUpdateInterval == 0 &&
                md->thermo->samples > 0)
                MB_TRY(mrmd_b200_thermo_update(md->thermo, c.thermoSmoothingSigma, c.thermoSmoothingIntensity, nullptr, st));
            MB_TRY(mrmd_b200_thermo_apply(md->thermo, a, nullptr, 0, st));
        }
        MB_TRY(mrmd_b200_molecules_contribute_force(md->mols, a, st));
        MB_TRY(mrmd_b200_ghost_contribute_back(md->ghost, a, st));
    }
    else
    {
        if (evStart) MB_CUDA(cudaEventRecord(evStart, st));
        MB_TRY(mrmd_b200_lj_apply(md->lj, a, md->list, nullptr, st));  // :178
        if (evStop) MB_CUDA(cudaEventRecord(evStop, st));
        if (!c.fullList) MB_TRY(mrmd_b200_ghost_contribute_back(md->ghost, a, st));  // :181 (no-op for a full list)
    }
    MB_TRY(postIntegrate(md, deferPost, st));  // :184
    md->step += 1;
    md->stepsSinceRebuild += 1;
    return 0;
}

// Steps queued ahead of the host.  The loop of examples/02:135-216 asks the host after every preForceIntegrate whether to
// rebuild (one device -> host scalar and a synchronisation per step: the GPU idles while the answer travels).  Here up to
// `count` steps are enqueued back to back with the criterion evaluated on the device (displacementDecisionKernel); the
// step that reaches skin / 2 raises a flag and every kernel queued behind it returns at once.  The host synchronises once
// per chunk: *done steps ran completely; if *stopped, the positions are those after preForceIntegrate of step *done and
// the caller finishes that step the usual way (rebuild, force).  Only for the tiled paths without per-step host
// bookkeeping (no constraints, no thermodynamic-force sampling / update and no compensation-energy sampling / update in
// the chunk: canQueue).
static bool canQueue(const mrmd_b200_md* md, int64_t ahead)
{
    const mrmd_b200_md_config& c = md->cfg;
    if (c.fullList != 2 || md->constraints != nullptr || c.atomsPerMolecule > 1 || md->rebuilds == 0) return false;
    if (std::getenv("MRMD_B200_NO_QUEUED_STEPS") != nullptr) return false;
    const int64_t step = md->step + ahead;
    if (c.adress)
    {
        if (md->thermo != nullptr)
        {
            if (c.thermoSampleInterval > 0 && step % c.thermoSampleInterval == 0) return false;
            if (c.thermoUpdateInterval > 0 && step > 0 && step % c.thermoUpdateInterval == 0) return false;
        }
        const int64_t run = md->adress->runCounter + ahead;
        if (run % md->adress->samplingInterval == 0 || run % md->adress->updateInterval == 0) return false;
    }
    return true;
}

static int runQueued(mrmd_b200_md* md, int64_t count, cudaStream_t st, cudaEvent_t* events, bool energyOnLast, bool energyAlways,
                     int64_t* done, bool* stopped)
{
    const mrmd_b200_md_config& c = md->cfg;
    mrmd_b200_atoms* a = md->atoms;
    if (md->dStop == nullptr)
    {
        MB_CUDA(cudaMalloc(&md->dStop, 16));
        MB_CUDA(cudaMalloc(&md->dAccum, 8));
        MB_CUDA(cudaMallocHost(&md->hStop, 32));
    }
    double* hAccum = reinterpret_cast<double*>(md->hStop + 4);
    *hAccum = md->maxDisplacement;
    MB_CUDA(cudaMemsetAsync(md->dStop, 0, 16, st));
    MB_CUDA(cudaMemcpyAsync(md->dAccum, hAccum, 8, cudaMemcpyHostToDevice, st));
    const int64_t step0 = md->step, since0 = md->stepsSinceRebuild, run0 = c.adress ? md->adress->runCounter : 0;
    const bool pending0 = md->postPending;
    for (int64_t k = 0; k < count; ++k)
    {
        MB_TRY(integratePre(a, c.dt, c.integrator == 1, c.zeta, c.temperature, c.seed, uint64_t(md->step), nullptr,
                            md->postPending, st, md->dStop));
        md->postPending = false;
        MB_TRY(displacementDecision(a->dMaxDisp, md->dAccum, c.skin * 0.5, md->dStop, static_cast<int>(k), st));
        const bool wantEnergy = energyAlways || (energyOnLast && k == count - 1);
        MB_TRY(stepAfterPre(md, st, events ? events[2 * k] : nullptr, events ? events[2 * k + 1] : nullptr, true, wantEnergy,
                            nullptr, false, md->dStop));
    }
    MB_CUDA(cudaMemcpyAsync(md->hStop, md->dStop, 16, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaMemcpyAsync(hAccum, md->dAccum, 8, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    MB_REQUIRE(md->hStop[2] == 0, "md_run: non-finite position, velocity or force (the system blew up)");
    *stopped = md->hStop[0] != 0;
    *done = *stopped ? md->hStop[1] : count;
    md->maxDisplacement = *hAccum;
    // the host-side counters follow the steps that really ran
    md->step = step0 + *done;
    md->stepsSinceRebuild = since0 + *done;
    if (c.adress) md->adress->runCounter = run0 + *done;
    // the stopped step's preForceIntegrate ran (with the previous step's kick fused in front if one was pending)
    md->postPending = *stopped ? false : (count > 0 ? true : pending0);
    return 0;
}

static int collectStats(mrmd_b200_md* md, int64_t nsteps, int64_t rebuilds0, int64_t storedSum, double pairs0,
                        int nTimed, mrmd_b200_md_stats* stats, cudaStream_t st)
{
    double* dRes = md->cfg.adress ? md->adress->dResult : md->lj->dResult;
    double* hRes = md->cfg.adress ? md->adress->hResult : md->lj->hResult;
    MB_CUDA(cudaMemcpyAsync(hRes, dRes, 48, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    if (md->cfg.adress && md->adress->uniformAtoms > 0) MB_TRY(adressCheckUniform(md->adress, st));
    if (stats == nullptr) return 0;
    *stats = mrmd_b200_md_stats{};
    stats->steps = nsteps;
    stats->rebuilds = md->rebuilds - rebuilds0;
    stats->storedPairs = storedSum;
    stats->numLocal = md->atoms->numLocal;
    stats->numGhost = md->atoms->numGhost;
    stats->maxDisplacement = md->maxDisplacement;
    stats->activePairs = 0;
    if (nsteps == 0)
    {
        // nothing ran: no energy / virial / pair counts of an earlier run are reported as this run's
        stats->forceKernelMs = 0.0;
        return 0;
    }
    if (md->cfg.adress)
    {
        stats->energy = hRes[0];
        stats->virial = 0.0;
        stats->pairInteractions = static_cast<int64_t>(hRes[4] - pairs0 + 0.5);
        stats->activePairs = static_cast<int64_t>(hRes[5] - md->active0 + 0.5);
    }
    else
    {
        stats->energy = hRes[0];
        stats->virial = hRes[1];
        stats->pairInteractions = static_cast<int64_t>(hRes[5] - pairs0 + 0.5);
    }
    double ms = 0.0;
    for (int i = 0; i < nTimed; ++i)
    {
        float t = 0.f;
        MB_CUDA(cudaEventElapsedTime(&t, md->events[2 * i], md->events[2 * i + 1]));
        ms += t;
    }
    stats->forceKernelMs = ms;
    return 0;
}

static int runningPairs(mrmd_b200_md* md, double* out, cudaStream_t st)
{
    double* dRes = md->cfg.adress ? md->adress->dResult : md->lj->dResult;
    double* hRes = md->cfg.adress ? md->adress->hResult : md->lj->hResult;
    MB_CUDA(cudaMemcpyAsync(hRes, dRes, 48, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    *out = md->cfg.adress ? hRes[4] : hRes[5];
    md->active0 = md->cfg.adress ? hRes[5] : 0.0;
    return 0;
}
}  // namespace mrmd_b200

using namespace mrmd_b200;

extern "C" {

int mrmd_b200_md_create(mrmd_b200_md** out, const mrmd_b200_md_config* cfg, const mrmd_b200_subdomain* s,
                        mrmd_b200_atoms* atoms)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(out != nullptr && cfg != nullptr && s != nullptr && atoms != nullptr, "md_create");
    MB_REQUIRE(cfg->dt > 0.0 && cfg->rc > 0.0 && cfg->skin >= 0.0 && cfg->maxNeighbors > 0, "md_create: bad config");
    MB_REQUIRE(!(cfg->adress && cfg->fullList == 1), "md_create: LJ_IdealGas takes a half list (0) or the tiled list (2)");
    const int64_t apm = std::max<int64_t>(cfg->atomsPerMolecule, 1);
    MB_REQUIRE(apm == 1 || (cfg->adress && ((cfg->fullList == 0 && cfg->cellSort == 0) || (cfg->fullList == 2 && apm == 4))),
               "md_create: multi-atom molecules need adress = 1 and fullList = 0 with cellSort = 0, or fullList = 2 (tiled "
               "list on the centres of mass) with four atoms per molecule");
    MB_REQUIRE(!(apm > 1 && cfg->fullList == 2 && cfg->useThermoForce), "md_create: no thermodynamic force on the tiled molecule path");
    MB_REQUIRE(atoms->numLocal % apm == 0, "md_create: local atoms are not a multiple of atomsPerMolecule");
    MB_REQUIRE(cfg->numConstraintIterations >= 0 && (cfg->numConstraintIterations == 0 || (apm > 1 && cfg->bondLength > 0.0)),
               "md_create: constraints need multi-atom molecules and a positive bond length");
    auto* md = new mrmd_b200_md;
    md->cfg = *cfg;
    md->sub = *s;
    md->atoms = atoms;
    int rc = mrmd_b200_ghost_create(&md->ghost);
    if (rc == 0) rc = mrmd_b200_verlet_create(&md->list, cfg->fullList ? 0 : 1);
    if (rc == 0 && cfg->adress && cfg->fullList == 2)
    {
        // tiles in the coarse-grained region hold ideal-gas pairs only: the tiled build leaves their rows empty
        md->list->tiledCgSkip = true;
        md->list->tiledCgWeight = cfg->weight;
    }
    if (rc == 0 && !cfg->adress)
        rc = mrmd_b200_lj_create(&md->lj, &cfg->cappingDistance, &cfg->rc, &cfg->sigma, &cfg->epsilon, 1, 0);
    if (rc == 0 && cfg->adress)
    {
        rc = mrmd_b200_adress_create(&md->adress, &cfg->cappingDistance, &cfg->rc, &cfg->sigma, &cfg->epsilon, 1,
                                     cfg->doShift);
        // four-lane kernel; MRMD_B200_ADRESS_NO_LANES=1 keeps the thread-per-molecule kernel (diagnostics)
        if (rc == 0 && apm == 4 && cfg->fullList == 0 && std::getenv("MRMD_B200_ADRESS_NO_LANES") == nullptr)
            rc = mrmd_b200_adress_set_atoms_per_molecule(md->adress, 4);
        const int64_t numMols = atoms->numLocal / apm;
        if (rc == 0) rc = mrmd_b200_molecules_create(&md->mols, std::max<int64_t>(numMols, 1));
        if (rc == 0 && numMols > 0)
        {
            // data::createMoleculeForEachAtom (data/MoleculesFromAtoms.cpp:19-39) for the local atoms, or molecules
            // of atomsPerMolecule consecutive atoms
            md->mols->size = numMols;
            md->mols->numLocal = numMols;
            moleculePerAtomInitKernel<<<gridFor(numMols, 256), 256>>>(md->mols->v, numMols, apm);
            g_launchCount.fetch_add(1);
            if (cudaDeviceSynchronize() != cudaSuccess) rc = MRMD_B200_EINVAL;
        }
        if (rc == 0 && cfg->numConstraintIterations > 0)
        {
            rc = mrmd_b200_constraints_create(&md->constraints, apm, cfg->numConstraintIterations);
            std::vector<int64_t> bi, bj;
            std::vector<double> eq;
            for (int64_t i = 0; i < apm; ++i)
                for (int64_t j = i + 1; j < apm; ++j)
                {
                    bi.push_back(i);
                    bj.push_back(j);
                    eq.push_back(cfg->bondLength);
                }
            if (rc == 0) rc = mrmd_b200_constraints_set(md->constraints, bi.data(), bj.data(), eq.data(), int64_t(eq.size()));
            if (rc == 0) constraintsSetUniformMolecules(md->constraints, true);
        }
        if (rc == 0 && cfg->useThermoForce)
            rc = mrmd_b200_thermo_create(&md->thermo, &cfg->thermoTargetDensity, 1, s, cfg->thermoBinWidth,
                                         &cfg->thermoModulation, 0, 0);
    }
    if (rc != 0)
    {
        mrmd_b200_md_destroy(md);
        return rc;
    }
    *out = md;
    return 0;
}

int mrmd_b200_md_destroy(mrmd_b200_md* md)
{
    if (md == nullptr) return 0;
    cudaDeviceSynchronize();
    for (auto e : md->events) cudaEventDestroy(e);
    md->hp.destroy();
    if (md->dStop != nullptr) cudaFree(md->dStop);
    if (md->dAccum != nullptr) cudaFree(md->dAccum);
    if (md->hStop != nullptr) cudaFreeHost(md->hStop);
    mrmd_b200_ghost_destroy(md->ghost);
    mrmd_b200_verlet_destroy(md->list);
    mrmd_b200_lj_destroy(md->lj);
    mrmd_b200_adress_destroy(md->adress);
    mrmd_b200_thermo_destroy(md->thermo);
    mrmd_b200_constraints_destroy(md->constraints);
    mrmd_b200_molecules_destroy(md->mols);
    delete md;
    return 0;
}

int mrmd_b200_md_run(mrmd_b200_md* md, int64_t nsteps, int timeForceKernel, mrmd_b200_md_stats* stats, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(md != nullptr && nsteps >= 0, "md_run");
    cudaStream_t st = S(stream);
    const int64_t rebuilds0 = md->rebuilds;
    double pairs0 = 0.0;
    MB_TRY(runningPairs(md, &pairs0, st));
    const int nTimed = timeForceKernel ? static_cast<int>(std::min<int64_t>(nsteps, 1 << 16)) : 0;
    while (static_cast<int>(md->events.size()) < 2 * nTimed)
    {
        cudaEvent_t e;
        MB_CUDA(cudaEventCreate(&e));
        md->events.push_back(e);
    }
    int64_t storedSum = 0;
    const bool energyAlways = md->cfg.energyEveryStep != 0;
    for (int64_t i = 0; i < nsteps;)
    {
        cudaEvent_t e0 = (i < nTimed) ? md->events[2 * i] : nullptr;
        cudaEvent_t e1 = (i < nTimed) ? md->events[2 * i + 1] : nullptr;
        // as many steps as the last rebuild interval suggests are queued ahead of the host (see runQueued)
        int64_t count = 0;
        const int64_t want = std::max<int64_t>(1, std::min<int64_t>(16, md->lastRebuildInterval - md->stepsSinceRebuild + 1));
        while (count < want && i + count < nsteps && canQueue(md, count)) ++count;
        if (count >= 1)
        {
            int64_t done = 0;
            bool stopped = false;
            MB_TRY(runQueued(md, count, st, (i + count <= nTimed) ? md->events.data() + 2 * i : nullptr, i + count == nsteps,
                             energyAlways, &done, &stopped));
            storedSum += md->storedPairsNow * done;
            i += done;
            if (stopped)
            {
                // step i: preForceIntegrate ran on the device and reached skin / 2 -> rebuild, force, postForceIntegrate
                e0 = (i < nTimed) ? md->events[2 * i] : nullptr;
                e1 = (i < nTimed) ? md->events[2 * i + 1] : nullptr;
                MB_TRY(stepAfterPre(md, st, e0, e1, true, i == nsteps - 1 || energyAlways, nullptr, true, nullptr));
                storedSum += md->storedPairsNow;
                i += 1;
            }
            continue;
        }
        MB_TRY(oneStep(md, st, e0, e1, true, i == nsteps - 1 || energyAlways));
        storedSum += md->storedPairsNow;
        ++i;
    }
    if (md->postPending)
    {
        MB_TRY(mrmd_b200_vv_post(md->atoms, md->cfg.dt, st));
        md->postPending = false;
    }
    return collectStats(md, nsteps, rebuilds0, storedSum, pairs0, nTimed, stats, st);
}

// Host-buffer path (hostpipe.cuh): per step pos and vel come from the host buffers, one step runs, pos, vel and
// {energy, virial, maxDisplacement} go back, copies chunked on two extra streams so that both PCIe directions stay busy.
int mrmd_b200_md_run_host(mrmd_b200_md* md, int64_t nsteps, double* posHost, double* velHost, double* scalarsHost,
                          mrmd_b200_md_stats* stats, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(md != nullptr && nsteps >= 0 && posHost != nullptr && velHost != nullptr, "md_run_host");
    cudaStream_t st = S(stream);
    const int64_t rebuilds0 = md->rebuilds;
    double pairs0 = 0.0;
    MB_TRY(runningPairs(md, &pairs0, st));
    int64_t storedSum = 0;
    const double* dRes = md->cfg.adress ? md->adress->dResult : md->lj->dResult;
    MB_TRY(hostPipeRun(md->hp, md->atoms, nsteps, posHost, velHost, scalarsHost, dRes, &md->maxDisplacement, false, st,
                       [&](cudaEvent_t evPosReady) -> int
                       {
                           MB_TRY(oneStep(md, st, nullptr, nullptr, false, true, evPosReady));
                           storedSum += md->storedPairsNow;
                           return 0;
                       }));
    return collectStats(md, nsteps, rebuilds0, storedSum, pairs0, 0, stats, st);
}

int mrmd_b200_md_set_energy_every_step(mrmd_b200_md* md, int enabled)
{
    MB_REQUIRE(md != nullptr, "md_set_energy_every_step");
    md->cfg.energyEveryStep = enabled ? 1 : 0;
    return 0;
}

int mrmd_b200_host_alloc(void** ptr, int64_t bytes)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(ptr != nullptr && bytes >= 0, "host_alloc");
    MB_CUDA(cudaMallocHost(ptr, size_t(std::max<int64_t>(bytes, 1))));
    return 0;
}

int mrmd_b200_host_free(void* ptr)
{
    if (ptr != nullptr) cudaFreeHost(ptr);
    return 0;
}

}  // extern "C"
