// lj.cu -- capped / shifted 12-6 Lennard-Jones force over a Verlet list (the dominant kernel).
// Reference: mrmd/action/LennardJones.hpp:24-206, mrmd/action/LennardJones.cpp:24-116.
//
// One thread per local atom walks its slot-major neighbour row: the index loads of a warp are one
// coalesced 128-byte line per slot, each partner position is a single 256-bit gather of the packed
// {x,y,z,type} record (L1/L2 resident for cell-sorted atoms).  Half list: partner forces are scattered
// with fp64 RED atomics exactly as the reference's atomic_access_slice; full list: no scatter at all.
// Energy, virial and the pair count are reduced warp -> block -> per-block partial; the last block to
// finish sums the partials in a fixed order, so the scalars are reproducible run to run.
// Algorithmic bytes per launch: 60*N + 52*P_stored.
#include <algorithm>
#include <vector>

#include "handles.cuh"

namespace mrmd_b200
{
// host restatement of CappedLennardJonesPotential's ctor + init kernel (LennardJones.cpp:59-116)
static void hostForceEnergy(const LJType& t, double distSqr, double& ff, double& e)
{
    if (distSqr >= t.cappingDistanceSqr)
    {
        const double frac2 = 1.0 / distSqr;
        const double frac6 = frac2 * frac2 * frac2;
        ff = frac6 * (t.ff1 * frac6 - t.ff2) * frac2;
        e = frac6 * (t.ef1 * frac6 - t.ef2) - t.shift;
        return;
    }
    const double dist = std::sqrt(distSqr);
    ff = t.cappingCoeff / dist;
    e = t.energyAtCappingPoint - (dist - t.cappingDistance) * t.cappingCoeff - t.shift;
}

int buildLJTable(LJTable& table, const double* cappingDistance, const double* rc, const double* sigma,
                 const double* epsilon, int64_t numTypes, int isShifted, double* rcSqrMax)
{
    MB_REQUIRE(numTypes >= 1 && numTypes <= MAX_LJ_TYPES, "number of atom types must be in [1, 4]");
    MB_REQUIRE(cappingDistance && rc && sigma && epsilon, "null LJ parameter array");
    double rcMax = 0.0;
    for (int64_t i = 0; i < numTypes * numTypes; ++i)
    {
        LJType v{};
        const double sig2 = sigma[i] * sigma[i];
        const double sig6 = sig2 * sig2 * sig2;
        v.ff1 = 48.0 * epsilon[i] * sig6 * sig6;
        v.ff2 = 24.0 * epsilon[i] * sig6;
        v.ef1 = 4.0 * epsilon[i] * sig6 * sig6;
        v.ef2 = 4.0 * epsilon[i] * sig6;
        v.rcSqr = rc[i] * rc[i];
        const double capDist = cappingDistance[i];
        v.cappingDistance = 0.0;
        v.cappingDistanceSqr = 0.0;
        double ff, e;
        hostForceEnergy(v, capDist * capDist, ff, e);
        v.cappingCoeff = ff * capDist;
        v.energyAtCappingPoint = e;
        v.cappingDistance = capDist;
        v.cappingDistanceSqr = capDist * capDist;
        if (isShifted)
        {
            hostForceEnergy(v, v.rcSqr, ff, e);
            v.shift = e;
        }
        table.t[i] = v;
        rcMax = std::max(rcMax, rc[i]);
    }
    *rcSqrMax = rcMax * rcMax;
    return 0;
}

constexpr int LJ_THREADS = 128;

template <bool HALF, bool PRED, bool SINGLE_TYPE>
__global__ void __launch_bounds__(LJ_THREADS)
    ljForceKernel(AtomsView a, int64_t numLocal, const int32_t* __restrict__ counts, const int32_t* __restrict__ neigh,
                  int64_t pitch, LJTable table, double rcSqr, int64_t numTypesQuirk, mrmd_b200_pred pred,
                  double* partials, double* result, unsigned int* ticket)
{
    const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    double energy = 0.0, virial = 0.0, pairs = 0.0;
    if (idx < numLocal)
    {
        const double4 pi = ld4nc(a.pos + idx);
        const int64_t typeI = typeOf(pi);
        const LJType t0 = table.t[0];
        double fx = 0.0, fy = 0.0, fz = 0.0;
        const int numNeighbors = counts[idx];
        const int32_t* row = neigh + idx;
        for (int n = 0; n < numNeighbors; ++n)
        {
            const int64_t jdx = row[int64_t(n) * pitch];
            const double4 pj = ld4nc(a.pos + jdx);
            if (PRED && !pred2(pred, pi.x, pi.y, pi.z, pj.x, pj.y, pj.z)) continue;  // LennardJones.hpp:173
            const double dx = pi.x - pj.x;
            const double dy = pi.y - pj.y;
            const double dz = pi.z - pj.z;
            const double distSqr = distSqrExact(dx, dy, dz);
            if (distSqr > rcSqr) continue;  // :182
            double ff, e;
            if (SINGLE_TYPE)
                ljForceEnergy(t0, distSqr, ff, e);
            else
                ljForceEnergy(table.t[typeI * numTypesQuirk + typeOf(pj)], distSqr, ff, e);  // :184
            energy += e;
            virial -= 0.5 * ff * distSqr;
            pairs += 1.0;
            fx += dx * ff;
            fy += dy * ff;
            fz += dz * ff;
            if (HALF)
            {
                atomicAdd(a.force[0] + jdx, -(dx * ff));  // :194-196, atomic_access_slice
                atomicAdd(a.force[1] + jdx, -(dy * ff));
                atomicAdd(a.force[2] + jdx, -(dz * ff));
            }
        }
        if (HALF)
        {
            atomicAdd(a.force[0] + idx, fx);  // :199-201
            atomicAdd(a.force[1] + idx, fy);
            atomicAdd(a.force[2] + idx, fz);
        }
        else
        {
            a.force[0][idx] += fx;  // nobody else writes row owners with a full list
            a.force[1][idx] += fy;
            a.force[2][idx] += fz;
        }
    }
    if (!HALF)
    {
        energy *= 0.5;  // every pair is visited from both sides
        virial *= 0.5;
        pairs *= 0.5;
    }
    gridReduce3<LJ_THREADS>(energy, virial, pairs, partials, result, ticket);
}

__global__ void ljEvalKernel(LJTable table, int64_t typeIdx, const double* distSqr, int64_t n, double* ff, double* e)
{
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    double f, en;
    ljForceEnergy(table.t[typeIdx], distSqr[i], f, en);
    ff[i] = f;
    e[i] = en;
}
}  // namespace mrmd_b200

using namespace mrmd_b200;

extern "C" {

int mrmd_b200_lj_create(mrmd_b200_lj** out, const double* cappingDistance, const double* rc, const double* sigma,
                        const double* epsilon, int64_t numTypes, int isShifted)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(out != nullptr, "lj_create");
    auto* lj = new mrmd_b200_lj;
    int rc_ = buildLJTable(lj->table, cappingDistance, rc, sigma, epsilon, numTypes, isShifted, &lj->rcSqr);
    if (rc_ == 0 && cudaMalloc(&lj->dResult, 48) != cudaSuccess) rc_ = MRMD_B200_ENOMEM;
    if (rc_ == 0 && cudaMalloc(&lj->dTicket, 4) != cudaSuccess) rc_ = MRMD_B200_ENOMEM;
    if (rc_ == 0 && cudaMallocHost(&lj->hResult, 48) != cudaSuccess) rc_ = MRMD_B200_ENOMEM;
    if (rc_ != 0)
    {
        delete lj;
        return rc_;
    }
    cudaMemset(lj->dResult, 0, 48);
    cudaMemset(lj->dTicket, 0, 4);
    lj->numTypes = numTypes;
    if (numTypes > 1)
        setLastError("note: LennardJones indexes its table with type_i + type_j (numTypes_ is 1 in the reference, "
                     "LennardJones.cpp:52): with several types pair (1, 1) uses the parameters of (1, 0)");
    *out = lj;
    return 0;
}

int mrmd_b200_lj_destroy(mrmd_b200_lj* lj)
{
    if (lj == nullptr) return 0;
    cudaDeviceSynchronize();
    if (lj->dResult) cudaFree(lj->dResult);
    if (lj->dTicket) cudaFree(lj->dTicket);
    if (lj->hResult) cudaFreeHost(lj->hResult);
    lj->partials.release();
    delete lj;
    return 0;
}

int mrmd_b200_lj_apply(mrmd_b200_lj* lj, mrmd_b200_atoms* a, const mrmd_b200_verlet* v, const mrmd_b200_pred* pred,
                       void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(lj != nullptr && a != nullptr && v != nullptr, "lj_apply");
    cudaStream_t st = S(stream);
    if (v->tiled)
    {
        MB_REQUIRE(pred == nullptr || pred->kind == MRMD_B200_PRED_ALWAYS, "lj_apply: predicates need a generic Verlet list");
        return ljApplyTiled(lj, a, v, true, true, st);
    }
    const int32_t* counts = v->counts.as<int32_t>();
    const int32_t* neigh = v->neigh.as<int32_t>();
    const int64_t pitch = v->pitch, numParticles = v->numParticles;
    const int half = v->half;
    const int64_t numLocal = a->numLocal;
    MB_REQUIRE(numLocal <= numParticles || numLocal == 0, "lj_apply: Verlet list has fewer rows than local atoms");
    // energyAndVirial_ = EnergyAndVirialReducer(), LennardJones.hpp:140
    MB_CUDA(cudaMemsetAsync(lj->dResult, 0, 24, st));
    if (numLocal == 0) return 0;
    const int blocks = gridFor(numLocal, LJ_THREADS);
    MB_TRY(lj->partials.reserve(size_t(blocks) * 3 * 8));
    mrmd_b200_pred p{};
    const bool usePred = (pred != nullptr && pred->kind != MRMD_B200_PRED_ALWAYS);
    if (usePred) p = *pred;
    const bool single = (lj->numTypes == 1);
#define LJ_LAUNCH(H, P, S1)                                                                                      \
    ljForceKernel<H, P, S1><<<blocks, LJ_THREADS, 0, st>>>(a->v, numLocal, counts, neigh, pitch, lj->table, lj->rcSqr, \
                                                           lj->numTypesQuirk, p, lj->partials.as<double>(),       \
                                                           lj->dResult, lj->dTicket)
    if (half)
    {
        if (usePred) { if (single) LJ_LAUNCH(true, true, true); else LJ_LAUNCH(true, true, false); }
        else { if (single) LJ_LAUNCH(true, false, true); else LJ_LAUNCH(true, false, false); }
    }
    else
    {
        if (usePred) { if (single) LJ_LAUNCH(false, true, true); else LJ_LAUNCH(false, true, false); }
        else { if (single) LJ_LAUNCH(false, false, true); else LJ_LAUNCH(false, false, false); }
    }
#undef LJ_LAUNCH
    MB_LAUNCHED();
    return 0;
}

int mrmd_b200_lj_get(mrmd_b200_lj* lj, double* energy, double* virial, int64_t* numPairs, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(lj != nullptr, "lj_get");
    MB_CUDA(cudaMemcpyAsync(lj->hResult, lj->dResult, 24, cudaMemcpyDeviceToHost, S(stream)));
    MB_CUDA(cudaStreamSynchronize(S(stream)));
    if (energy) *energy = lj->hResult[0];
    if (virial) *virial = lj->hResult[1];
    if (numPairs) *numPairs = static_cast<int64_t>(lj->hResult[2] + 0.5);
    return 0;
}

int mrmd_b200_lj_eval(const mrmd_b200_lj* lj, int64_t typeIdx, const double* distSqrHost, int64_t n,
                      double* forceFactorHost, double* energyHost, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(lj != nullptr && distSqrHost && forceFactorHost && energyHost && n >= 0, "lj_eval");
    MB_REQUIRE(typeIdx >= 0 && typeIdx < lj->numTypes * lj->numTypes, "lj_eval: type index out of range");
    if (n == 0) return 0;
    cudaStream_t st = S(stream);
    double* d = nullptr;
    MB_CUDA(cudaMalloc(&d, size_t(n) * 24));
    MB_CUDA(cudaMemcpyAsync(d, distSqrHost, size_t(n) * 8, cudaMemcpyHostToDevice, st));
    ljEvalKernel<<<gridFor(n, 128), 128, 0, st>>>(lj->table, typeIdx, d, n, d + n, d + 2 * n);
    g_launchCount.fetch_add(1);
    cudaMemcpyAsync(forceFactorHost, d + n, size_t(n) * 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(energyHost, d + 2 * n, size_t(n) * 8, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(d);
    MB_CUDA(e);
    return 0;
}

}  // extern "C"
