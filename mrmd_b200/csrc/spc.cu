// spc.cu -- SPC water and the Coulomb pair potentials: the other pair-force family that
// walks the half Verlet list of molecules like the AdResS kernel (K14) does.
// Reference: mrmd/action/SPC.hpp:61-362 (class SPC), mrmd/action/Coulomb.hpp:27-46, mrmd/action/CoulombDSF.hpp:42-84,
//            mrmd/util/math.hpp:57-76 (approxErfc), mrmd/util/angle.hpp:28.
//
// SPC_LANES lanes share one molecule alpha and stride over its neighbour row; for three-atom molecules alpha's atoms
// (position + charge) stay in registers, the force on them is reduced over the lanes with shuffles and leaves in one atomic per component,
// the partner's force is collected over alpha's atoms before it is added (9 instead of 27 atomics per molecule pair).
// Cutoff decisions use the uncontracted squared distance (distSqrExact), so they agree with the oracle bit for bit.
#include <algorithm>
#include <cmath>

#include "handles.cuh"

namespace mrmd_b200
{
int buildLJTable(LJTable& table, const double* cappingDistance, const double* rc, const double* sigma,
                 const double* epsilon, int64_t numTypes, int isShifted, double* rcSqrMax);

struct CoulombDev
{
    int kind;  // 0: impl::Coulomb, 1: impl::CoulombDSF
    double alpha, rc, forceShift, energyShift;
};
}  // namespace mrmd_b200

struct mrmd_b200_spc
{
    mrmd_b200::LJTable table{};
    mrmd_b200::CoulombDev coulomb{};
    double rcSqr = 0.0;
    double eqDistanceHO = 0.0, eqDistanceHH = 0.0;
    mrmd_b200_constraints* constraints = nullptr;  // MoleculeConstraints(3, 20) with the three bonds, SPC.hpp:238-250
    mrmd_b200::DevBuf partials;
    double* dResult = nullptr;  // [0..2] LJ energy, Coulomb energy, atom pairs inside the cutoff; [3..5] running sums
    unsigned int* dTicket = nullptr;
    double* hResult = nullptr;  // pinned
};

namespace mrmd_b200
{
constexpr int SPC_THREADS = 128;
constexpr int SPC_LANES = 4;
constexpr double COULOMB_PREFACTOR = 138.935458;  // Coulomb.hpp:32
constexpr double INV_SQRTPI = 0.564189583547756286948079451560772586;

// CoulombDSF::CoulombDSF, CoulombDSF.hpp:77-83
static CoulombDev coulombInit(int kind, double rc, double alpha)
{
    CoulombDev c{kind, alpha, rc, 0.0, 0.0};
    if (kind == 1)
    {
        const double rcSqr = rc * rc;
        const double erfc = std::erfc(alpha * rc);
        const double ex = std::exp(-alpha * alpha * rcSqr);
        c.forceShift = -(erfc / rcSqr + 2.0 * INV_SQRTPI * alpha * ex / rc);
        c.energyShift = erfc / rc;
    }
    return c;
}

// util::approxErfc, math.hpp:57-69
__device__ __forceinline__ double approxErfc(double x, double& expX2)
{
    constexpr double p = 0.3275911;
    constexpr double a1 = 0.254829592;
    constexpr double a2 = -0.284496736;
    constexpr double a3 = 1.421413741;
    constexpr double a4 = -1.453152027;
    constexpr double a5 = 1.061405429;
    const double t = 1.0 / (1.0 + p * x);
    expX2 = exp(-x * x);
    return t * (a1 + t * (a2 + t * (a3 + t * (a4 + t * a5)))) * expX2;
}

// computeForce + computeEnergy of Coulomb.hpp:30-43 / CoulombDSF.hpp:54-75 in one evaluation
template <bool DSF>
__device__ __forceinline__ void coulombForceEnergy(const CoulombDev& c, double distSqr, double q1, double q2, double& ff,
                                                   double& e)
{
    const double prefac = COULOMB_PREFACTOR * q1 * q2;
    const double r = sqrt(distSqr);
    if (!DSF)
    {
        ff = prefac / distSqr;
        e = prefac / r;
        return;
    }
    double expX2;
    const double erfc = approxErfc(c.alpha * r, expX2);
    const double force = prefac * (erfc / r + 2.0 * c.alpha * INV_SQRTPI * expX2 + r * c.forceShift);
    ff = force / distSqr;
    e = prefac * (erfc / r - c.energyShift - c.forceShift * (r - c.rc));
}

__global__ void coulombEvalKernel(CoulombDev c, const double* distSqr, int64_t n, double q1, double q2, double* force,
                                  double* energy)
{
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    double ff, e;
    if (c.kind == 1)
        coulombForceEnergy<true>(c, distSqr[i], q1, q2, ff, e);
    else
        coulombForceEnergy<false>(c, distSqr[i], q1, q2, ff, e);
    force[i] = ff;
    energy[i] = e;
}

__device__ __forceinline__ double lanesSum(double v)
{
#pragma unroll
    for (int o = SPC_LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One molecule pair exactly as SPC.hpp:150-234 walks it (any atom counts, every contribution an atomic)
template <bool DSF>
__device__ __forceinline__ void spcGenericPair(const AtomsView& a, longlong2 ocA, longlong2 ocB, const LJType& lj,
                                            const CoulombDev& coulomb, double rcSqr, double& eLJ, double& eC, double& pairs)
{
    const long long startAlpha = ocA.x, endAlpha = ocA.x + ocA.y;
    const long long startBeta = ocB.x, endBeta = ocB.x + ocB.y;
    {
        const double4 pi = ld4nc(a.pos + startAlpha), pj = ld4nc(a.pos + startBeta);
        const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        const double distSqr = distSqrExact(dx, dy, dz);
        if (distSqr < rcSqr)
        {
            double ff, e;
            ljForceEnergy(lj, distSqr, ff, e);
            eLJ += e;
            atomicAdd(a.force[0] + startBeta, -(dx * ff));
            atomicAdd(a.force[1] + startBeta, -(dy * ff));
            atomicAdd(a.force[2] + startBeta, -(dz * ff));
            atomicAdd(a.force[0] + startAlpha, dx * ff);
            atomicAdd(a.force[1] + startAlpha, dy * ff);
            atomicAdd(a.force[2] + startAlpha, dz * ff);
        }
    }
    for (long long idx = startAlpha; idx < endAlpha; ++idx)
    {
        const double4 pi = ld4nc(a.pos + idx);
        const double q1 = a.charge[idx];
        double fx = 0.0, fy = 0.0, fz = 0.0;
        for (long long jdx = startBeta; jdx < endBeta; ++jdx)
        {
            const double4 pj = ld4nc(a.pos + jdx);
            const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
            const double distSqr = distSqrExact(dx, dy, dz);
            if (distSqr > rcSqr) continue;
            double ff, e;
            coulombForceEnergy<DSF>(coulomb, distSqr, q1, a.charge[jdx], ff, e);
            eC += e;
            pairs += 1.0;
            fx += dx * ff;
            fy += dy * ff;
            fz += dz * ff;
            atomicAdd(a.force[0] + jdx, -(dx * ff));
            atomicAdd(a.force[1] + jdx, -(dy * ff));
            atomicAdd(a.force[2] + jdx, -(dz * ff));
        }
        atomicAdd(a.force[0] + idx, fx);
        atomicAdd(a.force[1] + idx, fy);
        atomicAdd(a.force[2] + idx, fz);
    }
}

// SPC::operator()(CalcInteractions, alpha, sumEnergy), SPC.hpp:143-236.  Pairs of three-atom molecules (water) take
// the register-resident path; any other pair is walked atom range by atom range as the reference does.
template <bool DSF>
__global__ void __launch_bounds__(SPC_THREADS, 4)
    spcForceKernel(MolsView m, AtomsView a, int64_t numLocalMols, const int32_t* __restrict__ counts,
                   const int32_t* __restrict__ neigh, int64_t pitch, LJType lj, CoulombDev coulomb, double rcSqr,
                   double* partials, double* result, unsigned int* ticket)
{
    const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    const int64_t alpha = t / SPC_LANES;
    const int sub = int(t % SPC_LANES);
    double eLJ = 0.0, eC = 0.0, pairs = 0.0;
    // lanes past the last molecule run with an empty row so that the shuffles below see full warps
    const bool valid = alpha < numLocalMols;
    const longlong2 ocA = valid ? m.oc[alpha] : make_longlong2(0, 0);
    const int numNeighbors = valid ? counts[alpha] : 0;
    const int32_t* row = neigh + alpha;
    const bool waterA = ocA.y == 3;
    double4 pA[3];
    double qA[3], fA[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
        pA[i] = waterA ? ld4nc(a.pos + ocA.x + i) : make_double4(0.0, 0.0, 0.0, 0.0);
        qA[i] = waterA ? a.charge[ocA.x + i] : 0.0;
        fA[i][0] = fA[i][1] = fA[i][2] = 0.0;
    }
    for (int n = sub; n < numNeighbors; n += SPC_LANES)
    {
        const int64_t beta = row[int64_t(n) * pitch];
        const longlong2 ocB = m.oc[beta];
        if (!waterA || ocB.y != 3)
        {
            spcGenericPair<DSF>(a, ocA, ocB, lj, coulomb, rcSqr, eLJ, eC, pairs);
            continue;
        }
        const long long startBeta = ocB.x;
        double4 pB[3];
        double qB[3], fB[3][3];
#pragma unroll
        for (int j = 0; j < 3; ++j)
        {
            pB[j] = ld4nc(a.pos + startBeta + j);
            qB[j] = a.charge[startBeta + j];
            fB[j][0] = fB[j][1] = fB[j][2] = 0.0;
        }
        {
            // LJ interaction between oxygen atoms, :165-190
            const double dx = pA[0].x - pB[0].x, dy = pA[0].y - pB[0].y, dz = pA[0].z - pB[0].z;
            const double distSqr = distSqrExact(dx, dy, dz);
            if (distSqr < rcSqr)
            {
                double ff, e;
                ljForceEnergy(lj, distSqr, ff, e);
                eLJ += e;
                fB[0][0] -= dx * ff;
                fB[0][1] -= dy * ff;
                fB[0][2] -= dz * ff;
                fA[0][0] += dx * ff;
                fA[0][1] += dy * ff;
                fA[0][2] += dz * ff;
            }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
#pragma unroll
            for (int j = 0; j < 3; ++j)
            {
                const double dx = pA[i].x - pB[j].x, dy = pA[i].y - pB[j].y, dz = pA[i].z - pB[j].z;
                const double distSqr = distSqrExact(dx, dy, dz);
                if (distSqr > rcSqr) continue;  // :214
                double ff, e;
                coulombForceEnergy<DSF>(coulomb, distSqr, qA[i], qB[j], ff, e);
                eC += e;
                pairs += 1.0;
                fA[i][0] += dx * ff;
                fA[i][1] += dy * ff;
                fA[i][2] += dz * ff;
                fB[j][0] -= dx * ff;
                fB[j][1] -= dy * ff;
                fB[j][2] -= dz * ff;
            }
        }
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int d = 0; d < 3; ++d)
                if (fB[j][d] != 0.0) atomicAdd(a.force[d] + startBeta + j, fB[j][d]);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
            const double f = lanesSum(fA[i][d]);
            if (waterA && sub == 0 && f != 0.0) atomicAdd(a.force[d] + ocA.x + i, f);
        }
    gridReduce3<SPC_THREADS>(eLJ, eC, pairs, partials, result, ticket);
}

// SPC::operator()(BondEnergy, alpha, sumEnergy), SPC.hpp:284-318
__global__ void __launch_bounds__(SPC_THREADS)
    spcBondEnergyKernel(MolsView m, AtomsView a, int64_t numAllMols, double eqDistanceHO, double eqDistanceHH,
                        double* partials, double* result, unsigned int* ticket)
{
    const int64_t alpha = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    double sum = 0.0;
    if (alpha < numAllMols)
    {
        const long long o = m.oc[alpha].x;
        const double4 pO = ld4nc(a.pos + o), pH0 = ld4nc(a.pos + o + 1), pH1 = ld4nc(a.pos + o + 2);
        double d = sqrt(distSqrExact(pO.x - pH0.x, pO.y - pH0.y, pO.z - pH0.z)) - eqDistanceHO;
        sum += d * d;
        d = sqrt(distSqrExact(pO.x - pH1.x, pO.y - pH1.y, pO.z - pH1.z)) - eqDistanceHO;
        sum += d * d;
        d = sqrt(distSqrExact(pH0.x - pH1.x, pH0.y - pH1.y, pH0.z - pH1.z)) - eqDistanceHH;
        sum += d * d;
    }
    gridReduce3<SPC_THREADS>(sum, 0.0, 0.0, partials, result, ticket);
}
}  // namespace mrmd_b200

using namespace mrmd_b200;

extern "C" {

int mrmd_b200_coulomb_eval(int kind, double rc, double alpha, const double* distSqrHost, int64_t n, double q1, double q2,
                           double* forceHost, double* energyHost, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE((kind == 0 || kind == 1) && n >= 0 && (n == 0 || (distSqrHost && forceHost && energyHost)), "coulomb_eval");
    MB_REQUIRE(kind == 0 || rc > 0.0, "coulomb_eval: CoulombDSF needs a positive cutoff");
    if (n == 0) return 0;
    cudaStream_t st = S(stream);
    double* d = nullptr;
    MB_CUDA(cudaMalloc(&d, size_t(n) * 3 * 8));
    MB_CUDA(cudaMemcpyAsync(d, distSqrHost, size_t(n) * 8, cudaMemcpyHostToDevice, st));
    coulombEvalKernel<<<gridFor(n, 128), 128, 0, st>>>(coulombInit(kind, rc, alpha), d, n, q1, q2, d + n, d + 2 * n);
    g_launchCount.fetch_add(1);
    cudaMemcpyAsync(forceHost, d + n, size_t(n) * 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(energyHost, d + 2 * n, size_t(n) * 8, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(d);
    MB_CUDA(e);
    return 0;
}

int mrmd_b200_spc_create(mrmd_b200_spc** out, int coulombKind)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(out != nullptr && (coulombKind == 0 || coulombKind == 1), "spc_create");
    constexpr double sigma = 0.31655578901998815, epsilon = 0.6501695808187486, rc = 1.2, alpha = 2.0;  // SPC.hpp:105-110
    const double cap = 0.7 * sigma;
    auto* spc = new mrmd_b200_spc;
    int rc_ = buildLJTable(spc->table, &cap, &rc, &sigma, &epsilon, 1, 1, &spc->rcSqr);  // SPC.hpp:347
    spc->rcSqr = rc * rc;
    spc->coulomb = coulombInit(coulombKind, rc, alpha);
    spc->eqDistanceHO = 0.1;                                     // :113
    const double angleHOH = 109.47 / 180.0 * M_PI;               // :114, util/angle.hpp:28
    spc->eqDistanceHH = spc->eqDistanceHO * std::sqrt(2.0 - 2.0 * std::cos(angleHOH));  // :115
    if (rc_ == 0 && cudaMalloc(&spc->dResult, 48) != cudaSuccess) rc_ = MRMD_B200_ENOMEM;
    if (rc_ == 0 && cudaMalloc(&spc->dTicket, 4) != cudaSuccess) rc_ = MRMD_B200_ENOMEM;
    if (rc_ == 0 && cudaMallocHost(&spc->hResult, 48) != cudaSuccess) rc_ = MRMD_B200_ENOMEM;
    if (rc_ == 0) rc_ = mrmd_b200_constraints_create(&spc->constraints, 3, 20);
    if (rc_ == 0)
    {
        const int64_t idx[3] = {0, 0, 1}, jdx[3] = {1, 2, 2};  // :353-361
        const double eq[3] = {spc->eqDistanceHO, spc->eqDistanceHO, spc->eqDistanceHH};
        rc_ = mrmd_b200_constraints_set(spc->constraints, idx, jdx, eq, 3);
    }
    if (rc_ != 0)
    {
        mrmd_b200_spc_destroy(spc);
        return rc_;
    }
    cudaMemset(spc->dResult, 0, 48);
    cudaMemset(spc->dTicket, 0, 4);
    *out = spc;
    return 0;
}

int mrmd_b200_spc_destroy(mrmd_b200_spc* spc)
{
    if (spc == nullptr) return 0;
    cudaDeviceSynchronize();
    if (spc->constraints) mrmd_b200_constraints_destroy(spc->constraints);
    if (spc->dResult) cudaFree(spc->dResult);
    if (spc->dTicket) cudaFree(spc->dTicket);
    if (spc->hResult) cudaFreeHost(spc->hResult);
    spc->partials.release();
    delete spc;
    return 0;
}

// SPC::applyForces, SPC.hpp:252-282
int mrmd_b200_spc_apply_forces(mrmd_b200_spc* spc, const mrmd_b200_molecules* m, const mrmd_b200_verlet* v,
                               mrmd_b200_atoms* a, double* energyLJ, double* energyCoulomb, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(spc != nullptr && m != nullptr && v != nullptr && a != nullptr, "spc_apply_forces");
    MB_REQUIRE(v->half == 1 && !v->tiled, "spc_apply_forces: SPC takes a half Verlet list of molecules");
    MB_REQUIRE(m->numLocal <= v->numParticles || m->numLocal == 0, "spc_apply_forces: list has fewer rows than molecules");
    cudaStream_t st = S(stream);
    MB_CUDA(cudaMemsetAsync(spc->dResult, 0, 24, st));
    if (m->numLocal > 0)
    {
        const int blocks = gridFor(m->numLocal * SPC_LANES, SPC_THREADS);
        MB_TRY(spc->partials.reserve(size_t(blocks) * 3 * 8));
        if (spc->coulomb.kind == 1)
            spcForceKernel<true><<<blocks, SPC_THREADS, 0, st>>>(m->v, a->v, m->numLocal, v->counts.as<int32_t>(),
                                                                 v->neigh.as<int32_t>(), v->pitch, spc->table.t[0], spc->coulomb,
                                                                 spc->rcSqr, spc->partials.as<double>(), spc->dResult, spc->dTicket);
        else
            spcForceKernel<false><<<blocks, SPC_THREADS, 0, st>>>(m->v, a->v, m->numLocal, v->counts.as<int32_t>(),
                                                                  v->neigh.as<int32_t>(), v->pitch, spc->table.t[0], spc->coulomb,
                                                                  spc->rcSqr, spc->partials.as<double>(), spc->dResult, spc->dTicket);
        MB_LAUNCHED();
    }
    if (energyLJ != nullptr || energyCoulomb != nullptr)
    {
        MB_CUDA(cudaMemcpyAsync(spc->hResult, spc->dResult, 24, cudaMemcpyDeviceToHost, st));
        MB_CUDA(cudaStreamSynchronize(st));  // Kokkos::fence(), :270
        if (energyLJ) *energyLJ = spc->hResult[0];
        if (energyCoulomb) *energyCoulomb = spc->hResult[1];
    }
    return 0;
}

// SPC::calcBondEnergy, SPC.hpp:320-344
int mrmd_b200_spc_calc_bond_energy(mrmd_b200_spc* spc, const mrmd_b200_molecules* m, const mrmd_b200_atoms* a,
                                   double harmonicPreFactor, double* bondEnergy, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(spc != nullptr && m != nullptr && a != nullptr && bondEnergy != nullptr, "spc_calc_bond_energy");
    cudaStream_t st = S(stream);
    const int64_t nMols = m->numLocal + m->numGhost;
    const int64_t nAtoms = a->numLocal + a->numGhost;
    MB_CUDA(cudaMemsetAsync(spc->dResult, 0, 24, st));
    if (nMols > 0)
    {
        const int blocks = gridFor(nMols, SPC_THREADS);
        MB_TRY(spc->partials.reserve(size_t(blocks) * 3 * 8));
        spcBondEnergyKernel<<<blocks, SPC_THREADS, 0, st>>>(m->v, a->v, nMols, spc->eqDistanceHO, spc->eqDistanceHH,
                                                            spc->partials.as<double>(), spc->dResult, spc->dTicket);
        MB_LAUNCHED();
    }
    MB_CUDA(cudaMemcpyAsync(spc->hResult, spc->dResult, 8, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    *bondEnergy = harmonicPreFactor * spc->hResult[0] / double(nAtoms);
    return 0;
}

// SPC::enforcePositionalConstraints / enforceVelocityConstraints, SPC.hpp:238-250
int mrmd_b200_spc_enforce_positional_constraints(mrmd_b200_spc* spc, const mrmd_b200_molecules* m, mrmd_b200_atoms* a,
                                                 double dt, void* stream)
{
    MB_REQUIRE(spc != nullptr, "spc_enforce_positional_constraints");
    return mrmd_b200_constraints_enforce_positional(spc->constraints, m, a, dt, stream);
}

int mrmd_b200_spc_enforce_velocity_constraints(mrmd_b200_spc* spc, const mrmd_b200_molecules* m, mrmd_b200_atoms* a,
                                               double dt, void* stream)
{
    MB_REQUIRE(spc != nullptr, "spc_enforce_velocity_constraints");
    return mrmd_b200_constraints_enforce_velocity(spc->constraints, m, a, dt, stream);
}

}  // extern "C"
