// adress.cu -- AdResS path: molecule update with inlined weighting function, lambda-weighted LJ / ideal-gas
// force interpolation with drift force and drift-compensation histograms, molecule -> atom force scatter
// (SURVEY.md K13-K16).
// Reference: mrmd/action/UpdateMolecules.hpp:24-70, mrmd/weighting_function/Slab.hpp:27-201,
//            Spherical.hpp:25-99, CheckRegion.hpp:25-38, mrmd/action/LJ_IdealGas.hpp:34-104,
//            LJ_IdealGas.cpp:21-292, mrmd/action/ContributeMoleculeForceToAtoms.cpp:23-48, util/math.hpp:31-46.
#include <algorithm>

#include "handles.cuh"
#include "weight.cuh"

namespace mrmd_b200
{
int buildLJTable(LJTable& table, const double* cappingDistance, const double* rc, const double* sigma,
                 const double* epsilon, int64_t numTypes, int isShifted, double* rcSqrMax);

constexpr int COMPENSATION_ENERGY_BINS = 200;  // LJ_IdealGas.hpp:54
constexpr int AD_THREADS = 128;

// UpdateMolecules::update, UpdateMolecules.hpp:41-67
__global__ void updateMoleculesKernel(MolsView m, AtomsView a, int64_t numAll, mrmd_b200_weight w)
{
    const int64_t mi = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (mi >= numAll) return;
    const longlong2 oc = m.oc[mi];
    double X = 0.0, Y = 0.0, Z = 0.0;
    for (long long ai = oc.x; ai < oc.x + oc.y; ++ai)
    {
        const double4 p = ld4nc(a.pos + ai);
        const double rm = a.relMass[ai];
        X += p.x * rm;
        Y += p.y * rm;
        Z += p.z * rm;
    }
    double lambda, modLambda, gx, gy, gz;
    weightEval(w, X, Y, Z, lambda, modLambda, gx, gy, gz);
    st4(m.pos + mi, make_double4(X, Y, Z, 0.0));
    st4(m.w + mi, make_double4(modLambda, gx, gy, gz));
    m.lambda[mi] = lambda;
}

// ContributeMoleculeForceToAtoms::update, ContributeMoleculeForceToAtoms.cpp:34-45
__global__ void contributeMoleculeForceKernel(MolsView m, AtomsView a, int64_t numAll)
{
    const int64_t mi = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (mi >= numAll) return;
    const longlong2 oc = m.oc[mi];
    const double fx = m.force[0][mi], fy = m.force[1][mi], fz = m.force[2][mi];
    if (fx == 0.0 && fy == 0.0 && fz == 0.0) return;  // the coarse-grained bulk: nothing to distribute
    for (long long ai = oc.x; ai < oc.x + oc.y; ++ai)
    {
        const double rm = a.relMass[ai];
        a.force[0][ai] += rm * fx;
        a.force[1][ai] += rm * fy;
        a.force[2][ai] += rm * fz;
    }
}

__global__ void weightEvalKernel(mrmd_b200_weight w, const double* pos, int64_t n, double* lambda, double* modLambda,
                                 double* grad)
{
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    double l, ml, gx, gy, gz;
    weightEval(w, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], l, ml, gx, gy, gz);
    lambda[i] = l;
    modLambda[i] = ml;
    grad[3 * i] = gx;
    grad[3 * i + 1] = gy;
    grad[3 * i + 2] = gz;
}

// Molecules whose row holds work: alpha outside the coarse-grained region, or any partner outside it (a CG-CG pair is
// ideal gas, LJ_IdealGas.cpp:102-107).  Their indices are appended warp by warp to activeList, so that the force
// kernels run on densely packed warps of working molecules instead of dragging the coarse-grained bulk along.
__global__ void adressActiveMoleculesKernel(MolsView m, int64_t numLocalMols, const int32_t* __restrict__ counts,
                                            const int32_t* __restrict__ neigh, int64_t pitch, int32_t* activeList,
                                            int* activeCount)
{
    const int64_t alpha = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    bool active = false;
    if (alpha < numLocalMols)
    {
        active = !inCG(reinterpret_cast<const double*>(m.w + alpha)[0]);
        if (!active)
        {
            const int numNeighbors = counts[alpha];
            const int32_t* row = neigh + alpha;
            for (int n = 0; n < numNeighbors && !active; ++n)
                active = !inCG(reinterpret_cast<const double*>(m.w + row[int64_t(n) * pitch])[0]);
        }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, active);
    if (ballot == 0u) return;
    const int lane = threadIdx.x & 31, leader = __ffs(ballot) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(activeCount, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (active) activeList[base + __popc(ballot & ((1u << lane) - 1u))] = static_cast<int32_t>(alpha);
}

// LJ_IdealGas::operator()(alpha, sumEnergy), LJ_IdealGas.cpp:52-225.  One thread per working molecule (activeList).
// The partner's {lambda^mod, grad lambda} come in one 256-bit gather; CG-CG pairs leave after it.
// hist: [0] compensationEnergy, [1] compensationEnergyCounter, [2] meanCompensationEnergy.
template <bool SAMPLING>
__global__ void __launch_bounds__(AD_THREADS)
    adressForceKernel(MolsView m, AtomsView a, int64_t numLocalMols, const int32_t* __restrict__ counts,
                      const int32_t* __restrict__ neigh, int64_t pitch, LJTable table, double rcSqr, int64_t numTypes,
                      double* hist, double* partials, double* result, unsigned int* ticket,
                      const int32_t* __restrict__ activeList, const int* __restrict__ activeCount)
{
    // the grid is sized for the case that every molecule has work; blocks past the list leave at once and stay out
    // of the reduction
    const int64_t numActive = min(numLocalMols, int64_t(*activeCount));
    const unsigned int workBlocks = static_cast<unsigned int>((numActive + AD_THREADS - 1) / AD_THREADS);
    if (blockIdx.x >= workBlocks) return;
    const int64_t slot = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    double sumEnergy = 0.0, pairs = 0.0, activePairs = 0.0;
    if (slot < numActive)
    {
        const int64_t alpha = activeList[slot];
        const int64_t T = numTypes;
        double* compensationEnergy = hist;
        double* compensationEnergyCounter = hist + COMPENSATION_ENERGY_BINS * T;
        const double* meanCompensationEnergy = hist + 2 * COMPENSATION_ENERGY_BINS * T;
        const double inverseBinSize = 1.0 / ((1.0 - 0.0) / double(COMPENSATION_ENERGY_BINS));

        double fAx = 0.0, fAy = 0.0, fAz = 0.0;
        const double4 wA = ld4nc(m.w + alpha);
        const double modLambdaAlpha = wA.x;
        const bool hyAlpha = inHY(modLambdaAlpha);
        const bool cgAlpha = inCG(modLambdaAlpha);
        long long binAlpha = -1;
        if (hyAlpha) binAlpha = histBin(0.0, inverseBinSize, COMPENSATION_ENERGY_BINS, m.lambda[alpha]);
        const longlong2 ocA = m.oc[alpha];
        const long long startAlpha = ocA.x, endAlpha = ocA.x + ocA.y;

        const int numNeighbors = counts[alpha];
        const int32_t* row = neigh + alpha;
        for (int n = 0; n < numNeighbors; ++n)
        {
            const int64_t beta = row[int64_t(n) * pitch];
            const double4 wB = ld4nc(m.w + beta);
            const double modLambdaBeta = wB.x;
            if (cgAlpha && inCG(modLambdaBeta)) continue;  // ideal gas, :102-107
            activePairs += 1.0;
            const double weighting = 0.5 * (modLambdaAlpha + modLambdaBeta);
            const bool hyBeta = inHY(modLambdaBeta);
            const bool drift = hyAlpha || hyBeta;
            double fBx = 0.0, fBy = 0.0, fBz = 0.0;
            const longlong2 ocB = m.oc[beta];
            const long long startBeta = ocB.x, endBeta = ocB.x + ocB.y;
            long long binBeta = -1;
            if (SAMPLING && hyBeta) binBeta = histBin(0.0, inverseBinSize, COMPENSATION_ENERGY_BINS, m.lambda[beta]);

            for (long long idx = startAlpha; idx < endAlpha; ++idx)
            {
                const double4 pi = ld4nc(a.pos + idx);
                double fx = 0.0, fy = 0.0, fz = 0.0;
                for (long long jdx = startBeta; jdx < endBeta; ++jdx)
                {
                    const double4 pj = ld4nc(a.pos + jdx);
                    const double dx = pi.x - pj.x;
                    const double dy = pi.y - pj.y;
                    const double dz = pi.z - pj.z;
                    const double distSqr = distSqrExact(dx, dy, dz);
                    if (distSqr > rcSqr) continue;  // :137
                    double ff, e;
                    ljForceEnergy(table.t[typeOf(pi) * T + typeOf(pj)], distSqr, ff, e);
                    const double ffactor = ff * weighting;
                    pairs += 1.0;
                    fx += dx * ffactor;
                    fy += dy * ffactor;
                    fz += dz * ffactor;
                    atomicAdd(a.force[0] + jdx, -(dx * ffactor));
                    atomicAdd(a.force[1] + jdx, -(dy * ffactor));
                    atomicAdd(a.force[2] + jdx, -(dz * ffactor));
                    sumEnergy += e * weighting;
                    const double Vij = 0.5 * e;
                    if (drift)
                    {
                        fAx += -Vij * wA.y;  // drift force, :163-169
                        fAy += -Vij * wA.z;
                        fAz += -Vij * wA.w;
                        fBx += -Vij * wB.y;
                        fBy += -Vij * wB.z;
                        fBz += -Vij * wB.w;
                        if (SAMPLING)
                        {
                            if (hyAlpha && binAlpha != -1) atomicAdd(compensationEnergy + binAlpha * T + typeOf(pi), Vij);
                            if (hyBeta && binBeta != -1) atomicAdd(compensationEnergy + binBeta * T + typeOf(pj), Vij);
                        }
                    }
                }
                atomicAdd(a.force[0] + idx, fx);
                atomicAdd(a.force[1] + idx, fy);
                atomicAdd(a.force[2] + idx, fz);
            }
            if (fBx != 0.0 || fBy != 0.0 || fBz != 0.0)
            {
                atomicAdd(m.force[0] + beta, fBx);
                atomicAdd(m.force[1] + beta, fBy);
                atomicAdd(m.force[2] + beta, fBz);
            }
        }
        if (SAMPLING && hyAlpha && binAlpha != -1)
        {
            // the reference increments non-atomically from a parallel loop (a race, LJ_IdealGas.cpp:207);
            // atomics give the serial result
            for (long long ai = startAlpha; ai < endAlpha; ++ai)
                atomicAdd(compensationEnergyCounter + binAlpha * T + typeOf(ld4nc(a.pos + ai)), 1.0);
        }
        if (hyAlpha && binAlpha != -1)
        {
            const double mean = meanCompensationEnergy[binAlpha * T + typeOf(ld4nc(a.pos + startAlpha))];  // :212-220
            fAx += mean * wA.y;
            fAy += mean * wA.z;
            fAz += mean * wA.w;
        }
        if (fAx != 0.0 || fAy != 0.0 || fAz != 0.0)
        {
            atomicAdd(m.force[0] + alpha, fAx);
            atomicAdd(m.force[1] + alpha, fAy);
            atomicAdd(m.force[2] + alpha, fAz);
        }
    }
    gridReduce3<AD_THREADS>(sumEnergy, pairs, activePairs, partials, result, ticket, workBlocks);
}

// The same operator for molecules of exactly four atoms (the tetramers of BASELINE.json configs[3]): four lanes share
// one molecule alpha, lane i owns alpha's atom i (position and force accumulator in registers) and walks the four atoms
// of the partner; the partner's forces are reduce-scattered over the four lanes (9 shuffles) so that lane j adds the
// force on the partner's atom j: 4x the parallelism of the thread-per-molecule kernel inside the small atomistic
// region and 8 instead of 20 force atomics per molecule pair.  A molecule with another atom count raises *error.
constexpr int AD_LANES = 4;
template <bool SAMPLING>
__global__ void __launch_bounds__(AD_THREADS, 5)
    adressForceLanes4Kernel(MolsView m, AtomsView a, int64_t numLocalMols, const int32_t* __restrict__ counts,
                            const int32_t* __restrict__ neigh, int64_t pitch, LJTable table, double rcSqr, int64_t numTypes,
                            double* hist, double* partials, double* result, unsigned int* ticket, int* error,
                            const int32_t* __restrict__ activeList, const int* __restrict__ activeCount)
{
    const int64_t numActive = min(numLocalMols, int64_t(*activeCount));
    const unsigned int workBlocks = static_cast<unsigned int>((numActive * AD_LANES + AD_THREADS - 1) / AD_THREADS);
    if (blockIdx.x >= workBlocks) return;  // see adressForceKernel
    const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    const int64_t slot = t / AD_LANES;
    const int li = int(t % AD_LANES);
    const unsigned gmask = 0xFu << ((threadIdx.x & 31) & ~3);
    double sumEnergy = 0.0, pairs = 0.0, activePairs = 0.0;
    bool valid = slot < numActive;
    const int64_t alpha = valid ? activeList[slot] : 0;
    longlong2 ocA = make_longlong2(0, 0);
    if (valid)
    {
        ocA = m.oc[alpha];
        if (ocA.y != AD_LANES)
        {
            *error = 1;
            valid = false;
        }
    }
    if (valid)
    {
        const int64_t T = numTypes;
        double* compensationEnergy = hist;
        double* compensationEnergyCounter = hist + COMPENSATION_ENERGY_BINS * T;
        const double* meanCompensationEnergy = hist + 2 * COMPENSATION_ENERGY_BINS * T;
        const double inverseBinSize = 1.0 / ((1.0 - 0.0) / double(COMPENSATION_ENERGY_BINS));

        const double4 wA = ld4nc(m.w + alpha);
        const double modLambdaAlpha = wA.x;
        const bool hyAlpha = inHY(modLambdaAlpha);
        const bool cgAlpha = inCG(modLambdaAlpha);
        long long binAlpha = -1;
        if (hyAlpha) binAlpha = histBin(0.0, inverseBinSize, COMPENSATION_ENERGY_BINS, m.lambda[alpha]);
        const double4 pi = ld4nc(a.pos + ocA.x + li);
        const int64_t typeI = typeOf(pi);
        double fIx = 0.0, fIy = 0.0, fIz = 0.0;  // force on alpha's atom li
        double sumVA = 0.0;                       // this lane's share of sum Vij over the drift pairs of alpha

        const int numNeighbors = counts[alpha];
        const int32_t* row = neigh + alpha;
        for (int n = 0; n < numNeighbors; ++n)
        {
            const int64_t beta = row[int64_t(n) * pitch];
            const double4 wB = ld4nc(m.w + beta);
            const double modLambdaBeta = wB.x;
            if (cgAlpha && inCG(modLambdaBeta)) continue;  // ideal gas, LJ_IdealGas.cpp:102-107
            const longlong2 ocB = m.oc[beta];
            if (ocB.y != AD_LANES)
            {
                *error = 1;
                continue;
            }
            if (li == 0) activePairs += 1.0;
            const double weighting = 0.5 * (modLambdaAlpha + modLambdaBeta);
            const bool hyBeta = inHY(modLambdaBeta);
            const bool drift = hyAlpha || hyBeta;
            long long binBeta = -1;
            if (SAMPLING && hyBeta) binBeta = histBin(0.0, inverseBinSize, COMPENSATION_ENERGY_BINS, m.lambda[beta]);
            double fb[AD_LANES][3];
            double sumV = 0.0;
#pragma unroll
            for (int j = 0; j < AD_LANES; ++j)
            {
                const double4 pj = ld4nc(a.pos + ocB.x + j);
                const double dx = pi.x - pj.x;
                const double dy = pi.y - pj.y;
                const double dz = pi.z - pj.z;
                const double distSqr = distSqrExact(dx, dy, dz);
                fb[j][0] = fb[j][1] = fb[j][2] = 0.0;
                if (distSqr > rcSqr) continue;  // :137
                double ff, e;
                ljForceEnergy(table.t[typeI * T + typeOf(pj)], distSqr, ff, e);
                const double ffactor = ff * weighting;
                pairs += 1.0;
                fIx += dx * ffactor;
                fIy += dy * ffactor;
                fIz += dz * ffactor;
                fb[j][0] = -(dx * ffactor);
                fb[j][1] = -(dy * ffactor);
                fb[j][2] = -(dz * ffactor);
                sumEnergy += e * weighting;
                const double Vij = 0.5 * e;
                if (drift)
                {
                    sumV += Vij;  // drift force, :163-169
                    if (SAMPLING && hyBeta && binBeta != -1) atomicAdd(compensationEnergy + binBeta * T + typeOf(pj), Vij);
                }
            }
            // reduce-scatter of the partner's forces: lane j ends up with the sum over alpha's atoms for partner atom j
            const bool hi = (li & 2) != 0, odd = (li & 1) != 0;
            double keep[2][3];
#pragma unroll
            for (int d = 0; d < 3; ++d)
            {
                const double s0 = hi ? fb[0][d] : fb[2][d], s1 = hi ? fb[1][d] : fb[3][d];
                keep[0][d] = (hi ? fb[2][d] : fb[0][d]) + __shfl_xor_sync(gmask, s0, 2);
                keep[1][d] = (hi ? fb[3][d] : fb[1][d]) + __shfl_xor_sync(gmask, s1, 2);
            }
#pragma unroll
            for (int d = 0; d < 3; ++d)
            {
                const double s = odd ? keep[0][d] : keep[1][d];
                const double mine = (odd ? keep[1][d] : keep[0][d]) + __shfl_xor_sync(gmask, s, 1);
                if (mine != 0.0) atomicAdd(a.force[d] + ocB.x + li, mine);
            }
            sumVA += sumV;
            if (drift)
            {
                double v = sumV;
                v += __shfl_xor_sync(gmask, v, 1);
                v += __shfl_xor_sync(gmask, v, 2);
                if (li == 0 && v != 0.0)
                {
                    atomicAdd(m.force[0] + beta, -v * wB.y);
                    atomicAdd(m.force[1] + beta, -v * wB.z);
                    atomicAdd(m.force[2] + beta, -v * wB.w);
                }
            }
        }
        if (fIx != 0.0 || fIy != 0.0 || fIz != 0.0)
        {
            atomicAdd(a.force[0] + ocA.x + li, fIx);
            atomicAdd(a.force[1] + ocA.x + li, fIy);
            atomicAdd(a.force[2] + ocA.x + li, fIz);
        }
        if (SAMPLING && hyAlpha && binAlpha != -1)
        {
            if (sumVA != 0.0) atomicAdd(compensationEnergy + binAlpha * T + typeI, sumVA);
            atomicAdd(compensationEnergyCounter + binAlpha * T + typeI, 1.0);
        }
        double vA = sumVA;
        vA += __shfl_xor_sync(gmask, vA, 1);
        vA += __shfl_xor_sync(gmask, vA, 2);
        if (li == 0)
        {
            double fAx = -vA * wA.y, fAy = -vA * wA.z, fAz = -vA * wA.w;
            if (hyAlpha && binAlpha != -1)
            {
                const double mean = meanCompensationEnergy[binAlpha * T + typeI];  // type of alpha's first atom, :212-220
                fAx += mean * wA.y;
                fAy += mean * wA.z;
                fAz += mean * wA.w;
            }
            if (fAx != 0.0 || fAy != 0.0 || fAz != 0.0)
            {
                atomicAdd(m.force[0] + alpha, fAx);
                atomicAdd(m.force[1] + alpha, fAy);
                atomicAdd(m.force[2] + alpha, fAz);
            }
        }
    }
    gridReduce3<AD_THREADS>(sumEnergy, pairs, activePairs, partials, result, ticket, workBlocks);
}

// updateMeanCompensationEnergy, LJ_IdealGas.cpp:21-50 (runningAverageFactor = 10)
__global__ void updateMeanCompensationKernel(double* hist, int64_t n, double runningAverageFactor)
{
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    double* energy = hist;
    double* counter = hist + n;
    double* mean = hist + 2 * n;
    if (counter[i] < 0.5) return;
    const double e = energy[i] / counter[i];
    mean[i] = (runningAverageFactor * mean[i] + e) / (runningAverageFactor + 1.0);
    energy[i] = 0.0;
    counter[i] = 0.0;
}
}  // namespace mrmd_b200

using namespace mrmd_b200;

extern "C" {

int mrmd_b200_molecules_update(mrmd_b200_molecules* m, const mrmd_b200_atoms* a, const mrmd_b200_weight* w, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(m != nullptr && a != nullptr && w != nullptr, "molecules_update");
    const int64_t n = m->numLocal + m->numGhost;
    if (n == 0) return 0;
    updateMoleculesKernel<<<gridFor(n, 256), 256, 0, S(stream)>>>(m->v, a->v, n, *w);
    MB_LAUNCHED();
    return 0;
}

int mrmd_b200_molecules_contribute_force(const mrmd_b200_molecules* m, mrmd_b200_atoms* a, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(m != nullptr && a != nullptr, "molecules_contribute_force");
    const int64_t n = m->numLocal + m->numGhost;
    if (n == 0) return 0;
    contributeMoleculeForceKernel<<<gridFor(n, 256), 256, 0, S(stream)>>>(m->v, a->v, n);
    MB_LAUNCHED();
    return 0;
}

int mrmd_b200_weight_eval(const mrmd_b200_weight* w, const double* posHost, int64_t n, double* lambdaHost,
                          double* modLambdaHost, double* gradHost, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(w && posHost && lambdaHost && modLambdaHost && gradHost && n >= 0, "weight_eval");
    if (n == 0) return 0;
    cudaStream_t st = S(stream);
    double* d = nullptr;
    MB_CUDA(cudaMalloc(&d, size_t(n) * 8 * 8));
    MB_CUDA(cudaMemcpyAsync(d, posHost, size_t(n) * 24, cudaMemcpyHostToDevice, st));
    weightEvalKernel<<<gridFor(n, 128), 128, 0, st>>>(*w, d, n, d + 3 * n, d + 4 * n, d + 5 * n);
    g_launchCount.fetch_add(1);
    cudaMemcpyAsync(lambdaHost, d + 3 * n, size_t(n) * 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(modLambdaHost, d + 4 * n, size_t(n) * 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(gradHost, d + 5 * n, size_t(n) * 24, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(d);
    MB_CUDA(e);
    return 0;
}

int mrmd_b200_adress_create(mrmd_b200_adress** out, const double* cappingDistance, const double* rc,
                            const double* sigma, const double* epsilon, int64_t numTypes, int doShift)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(out != nullptr, "adress_create");
    auto* ad = new mrmd_b200_adress;
    int rc_ = buildLJTable(ad->table, cappingDistance, rc, sigma, epsilon, numTypes, doShift, &ad->rcSqr);
    const size_t histBytes = size_t(3) * COMPENSATION_ENERGY_BINS * size_t(std::max<int64_t>(numTypes, 1)) * 8;
    if (rc_ == 0 && cudaMalloc(&ad->hist, histBytes) != cudaSuccess) rc_ = MRMD_B200_ENOMEM;
    if (rc_ == 0 && cudaMalloc(&ad->dResult, 48) != cudaSuccess) rc_ = MRMD_B200_ENOMEM;
    if (rc_ == 0 && cudaMalloc(&ad->dTicket, 4) != cudaSuccess) rc_ = MRMD_B200_ENOMEM;
    if (rc_ == 0 && cudaMallocHost(&ad->hResult, 48) != cudaSuccess) rc_ = MRMD_B200_ENOMEM;
    if (rc_ == 0 && cudaMalloc(&ad->dErr, 4) != cudaSuccess) rc_ = MRMD_B200_ENOMEM;
    if (rc_ == 0 && cudaMallocHost(&ad->hErr, 4) != cudaSuccess) rc_ = MRMD_B200_ENOMEM;
    if (rc_ != 0)
    {
        delete ad;
        return rc_;
    }
    cudaMemset(ad->hist, 0, histBytes);
    cudaMemset(ad->dResult, 0, 48);
    cudaMemset(ad->dTicket, 0, 4);
    cudaMemset(ad->dErr, 0, 4);
    ad->numTypes = numTypes;
    *out = ad;
    return 0;
}

int mrmd_b200_adress_destroy(mrmd_b200_adress* ad)
{
    if (ad == nullptr) return 0;
    cudaDeviceSynchronize();
    if (ad->hist) cudaFree(ad->hist);
    if (ad->dResult) cudaFree(ad->dResult);
    if (ad->dTicket) cudaFree(ad->dTicket);
    if (ad->hResult) cudaFreeHost(ad->hResult);
    if (ad->dErr) cudaFree(ad->dErr);
    if (ad->hErr) cudaFreeHost(ad->hErr);
    ad->partials.release();
    ad->activeList.release();
    delete ad;
    return 0;
}

int mrmd_b200_adress_set_intervals(mrmd_b200_adress* ad, int64_t samplingInterval, int64_t updateInterval)
{
    MB_REQUIRE(ad != nullptr && samplingInterval > 0 && updateInterval > 0, "adress_set_intervals");
    ad->samplingInterval = samplingInterval;
    ad->updateInterval = updateInterval;
    return 0;
}

int mrmd_b200_adress_set_atoms_per_molecule(mrmd_b200_adress* ad, int64_t atomsPerMolecule)
{
    MB_REQUIRE(ad != nullptr && atomsPerMolecule >= 0, "adress_set_atoms_per_molecule");
    ad->uniformAtoms = atomsPerMolecule;
    return 0;
}

// LJ_IdealGas::run, LJ_IdealGas.cpp:227-260
int mrmd_b200_adress_run(mrmd_b200_adress* ad, mrmd_b200_molecules* m, const mrmd_b200_verlet* v, mrmd_b200_atoms* a,
                         double* energy, int64_t* numPairs, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(ad != nullptr && m != nullptr && v != nullptr && a != nullptr, "adress_run");
    MB_REQUIRE(v->half == 1, "adress_run: LJ_IdealGas takes a half Verlet list of molecules");
    MB_REQUIRE(m->numLocal <= v->numParticles || m->numLocal == 0, "adress_run: list has fewer rows than molecules");
    cudaStream_t st = S(stream);
    const bool sampling = (ad->runCounter % ad->samplingInterval) == 0;
    MB_CUDA(cudaMemsetAsync(ad->dResult, 0, 24, st));
    if (m->numLocal > 0)
    {
        // working molecules first (see adressActiveMoleculesKernel); the force grid is sized for the worst case and
        // reads the count on the device
        MB_TRY(ad->activeList.reserve(16 + size_t(m->numLocal) * 4));  // {count, pad[3], list[numLocal]}
        int* activeCount = ad->activeList.as<int>();
        int32_t* activeList = ad->activeList.as<int32_t>() + 4;
        MB_CUDA(cudaMemsetAsync(activeCount, 0, 4, st));
        adressActiveMoleculesKernel<<<gridFor(m->numLocal, 256), 256, 0, st>>>(
            m->v, m->numLocal, v->counts.as<int32_t>(), v->neigh.as<int32_t>(), v->pitch, activeList, activeCount);
        MB_LAUNCHED();
        const bool lanes4 = ad->uniformAtoms == AD_LANES;
        const int blocks = gridFor(m->numLocal * (lanes4 ? AD_LANES : 1), AD_THREADS);
        MB_TRY(ad->partials.reserve(size_t(blocks) * 3 * 8));
        if (lanes4 && sampling)
            adressForceLanes4Kernel<true><<<blocks, AD_THREADS, 0, st>>>(
                m->v, a->v, m->numLocal, v->counts.as<int32_t>(), v->neigh.as<int32_t>(), v->pitch, ad->table, ad->rcSqr,
                ad->numTypes, ad->hist, ad->partials.as<double>(), ad->dResult, ad->dTicket, ad->dErr, activeList, activeCount);
        else if (lanes4)
            adressForceLanes4Kernel<false><<<blocks, AD_THREADS, 0, st>>>(
                m->v, a->v, m->numLocal, v->counts.as<int32_t>(), v->neigh.as<int32_t>(), v->pitch, ad->table, ad->rcSqr,
                ad->numTypes, ad->hist, ad->partials.as<double>(), ad->dResult, ad->dTicket, ad->dErr, activeList, activeCount);
        else if (sampling)
            adressForceKernel<true><<<blocks, AD_THREADS, 0, st>>>(
                m->v, a->v, m->numLocal, v->counts.as<int32_t>(), v->neigh.as<int32_t>(), v->pitch, ad->table, ad->rcSqr,
                ad->numTypes, ad->hist, ad->partials.as<double>(), ad->dResult, ad->dTicket, activeList, activeCount);
        else
            adressForceKernel<false><<<blocks, AD_THREADS, 0, st>>>(
                m->v, a->v, m->numLocal, v->counts.as<int32_t>(), v->neigh.as<int32_t>(), v->pitch, ad->table, ad->rcSqr,
                ad->numTypes, ad->hist, ad->partials.as<double>(), ad->dResult, ad->dTicket, activeList, activeCount);
        MB_LAUNCHED();
    }
    if (ad->runCounter % ad->updateInterval == 0)
    {
        const int64_t n = COMPENSATION_ENERGY_BINS * ad->numTypes;
        if (ad->preUpdateHook != nullptr) MB_TRY(ad->preUpdateHook(ad->hookCtx, ad->hist, 2 * n, st));
        updateMeanCompensationKernel<<<gridFor(n, 128), 128, 0, st>>>(ad->hist, n, 10.0);
        MB_LAUNCHED();
    }
    ad->runCounter += 1;
    if (energy != nullptr || numPairs != nullptr)
    {
        MB_CUDA(cudaMemcpyAsync(ad->hResult, ad->dResult, 24, cudaMemcpyDeviceToHost, st));
        if (ad->uniformAtoms == AD_LANES)
            MB_TRY(adressCheckUniform(ad, st));
        else
            MB_CUDA(cudaStreamSynchronize(st));
        if (energy) *energy = ad->hResult[0];
        if (numPairs) *numPairs = static_cast<int64_t>(ad->hResult[1] + 0.5);
    }
    return 0;
}

// UpdateMolecules::update + LJ_IdealGas::run + ContributeMoleculeForceToAtoms::update for one-atom molecules on a
// tiled periodic list (tiled.cu); run counter, sampling and mean update exactly as LJ_IdealGas.cpp:227-260
int mrmd_b200_adress_run_periodic(mrmd_b200_adress* ad, mrmd_b200_atoms* a, const mrmd_b200_verlet* v,
                                  const mrmd_b200_weight* w, double* energy, int64_t* numPairs, void* stream)
{
    MB_TRY(checkDevice());
    cudaStream_t st = S(stream);
    MB_TRY(adressRunPeriodic(ad, a, v, w, energy != nullptr || numPairs != nullptr, st));  // per-launch values: <ENERGY>
    if (energy != nullptr || numPairs != nullptr)
    {
        MB_CUDA(cudaMemcpyAsync(ad->hResult, ad->dResult, 24, cudaMemcpyDeviceToHost, st));
        MB_CUDA(cudaStreamSynchronize(st));
        if (energy) *energy = ad->hResult[0];
        if (numPairs) *numPairs = static_cast<int64_t>(ad->hResult[1] + 0.5);
    }
    return 0;
}

}  // extern "C"

namespace mrmd_b200
{
int adressCheckUniform(mrmd_b200_adress* ad, cudaStream_t st)
{
    MB_CUDA(cudaMemcpyAsync(ad->hErr, ad->dErr, 4, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    if (*ad->hErr != 0)
    {
        cudaMemsetAsync(ad->dErr, 0, 4, st);
        MB_REQUIRE(false, "adress_run: a molecule does not have the atom count promised by adress_set_atoms_per_molecule");
    }
    return 0;
}

int adressRunPeriodic(mrmd_b200_adress* ad, mrmd_b200_atoms* a, const mrmd_b200_verlet* v, const mrmd_b200_weight* w,
                      bool energy, cudaStream_t st, const int* stop)
{
    MB_REQUIRE(ad != nullptr && a != nullptr && v != nullptr && w != nullptr, "adress_run_periodic");
    MB_REQUIRE(v->tiled, "adress_run_periodic: needs a list from mrmd_b200_verlet_build_periodic");
    MB_REQUIRE(v->numParticles == a->numLocal, "adress_run_periodic: list rows != local atoms");
    const bool sampling = (ad->runCounter % ad->samplingInterval) == 0;
    if (a->numLocal > 0)
        MB_TRY(adressApplyTiled(ad, a, v, w, sampling, energy, st, stop));
    else
        MB_CUDA(cudaMemsetAsync(ad->dResult, 0, 24, st));
    if (ad->runCounter % ad->updateInterval == 0)
    {
        const int64_t n = COMPENSATION_ENERGY_BINS * ad->numTypes;
        if (ad->preUpdateHook != nullptr) MB_TRY(ad->preUpdateHook(ad->hookCtx, ad->hist, 2 * n, st));
        updateMeanCompensationKernel<<<gridFor(n, 128), 128, 0, st>>>(ad->hist, n, 10.0);
        MB_LAUNCHED();
    }
    ad->runCounter += 1;
    return 0;
}

int adressRunPeriodicMolecules(mrmd_b200_adress* ad, const mrmd_b200_molecules* m, mrmd_b200_atoms* a,
                               const mrmd_b200_verlet* v, const mrmd_b200_weight* w, int atomsPerMolecule, bool energy,
                               cudaStream_t st)
{
    MB_REQUIRE(ad != nullptr && m != nullptr && a != nullptr && v != nullptr && w != nullptr, "adress_run_periodic_molecules");
    MB_REQUIRE(v->tiled, "adress_run_periodic_molecules: needs a list from mrmd_b200_verlet_build_periodic_molecules");
    MB_REQUIRE(v->numParticles == m->numLocal && m->numLocal * atomsPerMolecule == a->numLocal,
               "adress_run_periodic_molecules: list rows != local molecules, or atoms != atomsPerMolecule x molecules");
    const bool sampling = (ad->runCounter % ad->samplingInterval) == 0;
    if (m->numLocal > 0)
        MB_TRY(moleculeApplyTiled(ad, m, a, v, w, atomsPerMolecule, sampling, energy, st));
    else
        MB_CUDA(cudaMemsetAsync(ad->dResult, 0, 24, st));
    if (ad->runCounter % ad->updateInterval == 0)
    {
        const int64_t n = COMPENSATION_ENERGY_BINS * ad->numTypes;
        if (ad->preUpdateHook != nullptr) MB_TRY(ad->preUpdateHook(ad->hookCtx, ad->hist, 2 * n, st));
        updateMeanCompensationKernel<<<gridFor(n, 128), 128, 0, st>>>(ad->hist, n, 10.0);
        MB_LAUNCHED();
    }
    ad->runCounter += 1;
    return 0;
}
}  // namespace mrmd_b200

extern "C" {

int mrmd_b200_adress_run_periodic_molecules(mrmd_b200_adress* ad, const mrmd_b200_molecules* m, mrmd_b200_atoms* a,
                                            const mrmd_b200_verlet* v, const mrmd_b200_weight* w, int atomsPerMolecule,
                                            double* energy, int64_t* numPairs, void* stream)
{
    MB_TRY(checkDevice());
    cudaStream_t st = S(stream);
    MB_TRY(adressRunPeriodicMolecules(ad, m, a, v, w, atomsPerMolecule, energy != nullptr || numPairs != nullptr, st));
    if (energy != nullptr || numPairs != nullptr)
    {
        MB_CUDA(cudaMemcpyAsync(ad->hResult, ad->dResult, 24, cudaMemcpyDeviceToHost, st));
        MB_CUDA(cudaStreamSynchronize(st));
        if (energy) *energy = ad->hResult[0];
        if (numPairs) *numPairs = static_cast<int64_t>(ad->hResult[1] + 0.5);
    }
    return 0;
}

int mrmd_b200_adress_read_histogram(const mrmd_b200_adress* ad, int kind, double* dstHost, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(ad != nullptr && dstHost != nullptr && kind >= 0 && kind <= 2, "adress_read_histogram");
    const int64_t n = COMPENSATION_ENERGY_BINS * ad->numTypes;
    const int slot = (kind == 0) ? 2 : (kind == 1 ? 0 : 1);
    MB_CUDA(cudaMemcpyAsync(dstHost, ad->hist + slot * n, size_t(n) * 8, cudaMemcpyDeviceToHost, S(stream)));
    MB_CUDA(cudaStreamSynchronize(S(stream)));
    return 0;
}

}  // extern "C"
