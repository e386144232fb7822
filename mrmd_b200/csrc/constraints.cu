// constraints.cu -- Berendsen thermostat / barostat and SHAKE / RATTLE molecule constraints (SURVEY.md section 8 row
// (f)2: what the reference's NVT / NPT / Constraints integration tests and tetramer runs put around the hot loop).
// Reference: mrmd/action/BerendsenThermostat.cpp:25-50, BerendsenBarostat.cpp:23-50,
//            mrmd/action/Shake.hpp:35-157 (impl::Shake), :159-251 (MoleculeConstraints), mrmd/data/Bond.hpp:23-28.
// Streaming kernels (24-56 B per atom); SHAKE walks the bonds of a molecule sequentially in one thread exactly like the
// reference's per-molecule lambda, so no atomics are needed (a molecule's atoms belong to it alone).
#include <algorithm>
#include <cmath>
#include <vector>

#include "handles.cuh"

struct mrmd_b200_constraints
{
    int64_t atomsPerMolecule = 0;
    int64_t numIterations = 0;
    int64_t numBonds = 0;
    int64_t maxBondAtoms = 0;       // 1 + the largest atom index a bond names: <= 4 takes the fused SHAKE kernel
    mrmd_b200::DevBuf bondIdx;      // int64 {idx, jdx} per bond
    mrmd_b200::DevBuf bondDist;     // double eqDistance per bond
    mrmd_b200::DevBuf updatedPos;   // double4 per atom (impl::Shake::updatedPos_)
    int* dErr = nullptr;            // set by a kernel that meets a bond outside its molecule
    // set by the step-loop drivers only: every local molecule has exactly maxBondAtoms atoms and owns all local atoms
    bool uniformMolecules = false;
    // the bonds are all pairs (i, j), i < j, of the first maxBondAtoms atoms in lexicographic order (a rigid triangle or
    // tetrahedron): the register-resident kernels apply
    bool allPairs = false;
};

namespace mrmd_b200
{
__global__ void scaleVelocityKernel(AtomsView a, int64_t n, double beta)
{
    const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (idx >= n) return;
    a.vel[0][idx] *= beta;  // BerendsenThermostat.cpp:40-42
    a.vel[1][idx] *= beta;
    a.vel[2][idx] *= beta;
}

// limitAccelerationPerComponent, action/LimitAcceleration.cpp:21-45 (min then max, each on force * invM, times m)
__global__ void limitAccelerationKernel(AtomsView a, int64_t n, double maxAcc)
{
    const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (idx >= n) return;
    const double m = a.mass[idx];
    const double invM = 1.0 / m;
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
        double f = a.force[d][idx];
        f = fmin(f * invM, +maxAcc) * m;
        f = fmax(f * invM, -maxAcc) * m;
        a.force[d][idx] = f;
    }
}

// limitVelocityPerComponent, action/LimitVelocity.cpp:23-43
__global__ void limitVelocityKernel(AtomsView a, int64_t n, double maxVel)
{
    const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (idx >= n) return;
#pragma unroll
    for (int d = 0; d < 3; ++d) a.vel[d][idx] = fmax(fmin(a.vel[d][idx], +maxVel), -maxVel);
}

__global__ void scalePositionKernel(double4* pos, int64_t n, double mu, int sx, int sy, int sz)
{
    const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (idx >= n) return;
    double4 p = ld4(pos + idx);
    if (sx) p.x *= mu;  // BerendsenBarostat.cpp:40-42
    if (sy) p.y *= mu;
    if (sz) p.z *= mu;
    st4(pos + idx, p);
}

// Shake::operator()(UnconstraintUpdate, idx), Shake.hpp:131-137
__global__ void shakeUnconstraintKernel(AtomsView a, int64_t n, double dtv, double dtf, double4* updatedPos)
{
    const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (idx >= n) return;
    const double4 p = ld4(a.pos + idx);
    const double dtfm = dtf / a.mass[idx];
    st4(updatedPos + idx, make_double4(p.x + dtv * a.vel[0][idx] + dtfm * a.force[0][idx],
                                       p.y + dtv * a.vel[1][idx] + dtfm * a.force[1][idx],
                                       p.z + dtv * a.vel[2][idx] + dtfm * a.force[2][idx], 0.0));
}

// MoleculeConstraints::enforcePositionalConstraints lambda (Shake.hpp:179-197) with
// Shake::enforcePositionalConstraint (:84-128) inlined; one thread per molecule
__global__ void shakePositionalKernel(MolsView m, AtomsView a, int64_t numLocalMols, const long long* __restrict__ bondIdx,
                                      const double* __restrict__ bondDist, int64_t numBonds,
                                      const double4* __restrict__ updatedPos, double dtf, int* error)
{
    const int64_t mol = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (mol >= numLocalMols) return;
    const longlong2 oc = m.oc[mol];
    for (int64_t b = 0; b < numBonds; ++b)
    {
        const long long bi = bondIdx[2 * b], bj = bondIdx[2 * b + 1];
        if (bi >= oc.y || bj >= oc.y)  // MRMD_DEVICE_ASSERT_LESS: not enough atoms in molecule to satisfy bond
        {
            *error = 1;
            return;
        }
        const long long idx = oc.x + bi, jdx = oc.x + bj;
        const double eqDistance = bondDist[b];
        const double4 pi = ld4(a.pos + idx), pj = ld4(a.pos + jdx);
        const double dist[3] = {pi.x - pj.x, pi.y - pj.y, pi.z - pj.z};
        const double distSq = dist[0] * dist[0] + dist[1] * dist[1] + dist[2] * dist[2];
        const double4 ui = ld4(updatedPos + idx), uj = ld4(updatedPos + jdx);
        const double upd[3] = {ui.x - uj.x, ui.y - uj.y, ui.z - uj.z};
        const double updSq = upd[0] * upd[0] + upd[1] * upd[1] + upd[2] * upd[2];
        const double invMassI = 1.0 / a.mass[idx], invMassJ = 1.0 / a.mass[jdx];
        const double qa = (invMassI + invMassJ) * (invMassI + invMassJ) * distSq;
        const double qb = 2.0 * (invMassI + invMassJ) * (upd[0] * dist[0] + upd[1] * dist[1] + upd[2] * dist[2]);
        const double qc = updSq - eqDistance * eqDistance;
        double determinant = qb * qb - 4.0 * qa * qc;
        determinant = fmax(0.0, determinant);
        const double root = sqrt(determinant);
        const double lambda1 = (-qb + root) / (2.0 * qa);
        const double lambda2 = (-qb - root) / (2.0 * qa);
        double lambda = (fabs(lambda1) < fabs(lambda2)) ? lambda1 : lambda2;
        lambda /= dtf;
        for (int d = 0; d < 3; ++d)
        {
            a.force[d][idx] += lambda * dist[d];
            a.force[d][jdx] -= lambda * dist[d];
        }
    }
}

// All numConstraintIterations of enforcePositionalConstraints in one pass for bonds among the first NA <= 4 atoms of
// a molecule: the unconstrained update of an iteration only feeds the bonds of the atom's own molecule, so a thread
// keeps its molecule's {pos, pos + dtv vel, force, mass, updatedPos} in shared memory (slot-major, conflict free),
// walks iterations x bonds like the two-kernel sequence does, and writes the forces back once: 112 B per atom of HBM
// traffic instead of numIterations x (112 + gathers).
constexpr int SHAKE_FUSED_THREADS = 64;
constexpr int SHAKE_FUSED_SLOTS = 14;
template <int NA>
__global__ void __launch_bounds__(SHAKE_FUSED_THREADS)
    shakeFusedKernel(MolsView m, AtomsView a, int64_t numLocalMols, const long long* __restrict__ bondIdx,
                     const double* __restrict__ bondDist, int64_t numBonds, int64_t numIterations, double dtv, double dtf,
                     int* error)
{
    __shared__ double sm[SHAKE_FUSED_SLOTS * NA * SHAKE_FUSED_THREADS];
    const int64_t mol = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (mol >= numLocalMols) return;
    const longlong2 oc = m.oc[mol];
    if (oc.y < NA)  // some bond names an atom the molecule does not have
    {
        *error = 1;
        return;
    }
    auto at = [&](int slot, long long atom) -> double& { return sm[(slot * NA + atom) * SHAKE_FUSED_THREADS + threadIdx.x]; };
    enum { POS = 0, BASE = 3, FORCE = 6, MASS = 9, UPD = 10, INVMASS = 13 };
#pragma unroll
    for (int k = 0; k < NA; ++k)
    {
        const double4 p = ld4(a.pos + oc.x + k);
        const double pk[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
            at(POS + d, k) = pk[d];
            at(BASE + d, k) = pk[d] + dtv * a.vel[d][oc.x + k];
            at(FORCE + d, k) = a.force[d][oc.x + k];
        }
        at(MASS, k) = a.mass[oc.x + k];
        at(INVMASS, k) = 1.0 / at(MASS, k);  // the same quotient for every bond of the atom
    }
    for (int64_t it = 0; it < numIterations; ++it)
    {
#pragma unroll
        for (int k = 0; k < NA; ++k)  // Shake::operator()(UnconstraintUpdate, idx), Shake.hpp:131-137
        {
            const double dtfm = dtf * at(INVMASS, k);  // dtf / mass up to an ulp (the divisions bound this kernel)
#pragma unroll
            for (int d = 0; d < 3; ++d) at(UPD + d, k) = at(BASE + d, k) + dtfm * at(FORCE + d, k);
        }
        for (int64_t b = 0; b < numBonds; ++b)  // Shake::enforcePositionalConstraint, :84-128
        {
            const long long i = bondIdx[2 * b], j = bondIdx[2 * b + 1];
            const double eqDistance = bondDist[b];
            const double dist[3] = {at(POS, i) - at(POS, j), at(POS + 1, i) - at(POS + 1, j), at(POS + 2, i) - at(POS + 2, j)};
            const double distSq = dist[0] * dist[0] + dist[1] * dist[1] + dist[2] * dist[2];
            const double upd[3] = {at(UPD, i) - at(UPD, j), at(UPD + 1, i) - at(UPD + 1, j), at(UPD + 2, i) - at(UPD + 2, j)};
            const double updSq = upd[0] * upd[0] + upd[1] * upd[1] + upd[2] * upd[2];
            const double invMassI = at(INVMASS, i), invMassJ = at(INVMASS, j);
            const double qa = (invMassI + invMassJ) * (invMassI + invMassJ) * distSq;
            const double qb = 2.0 * (invMassI + invMassJ) * (upd[0] * dist[0] + upd[1] * dist[1] + upd[2] * dist[2]);
            const double qc = updSq - eqDistance * eqDistance;
            double determinant = qb * qb - 4.0 * qa * qc;
            determinant = fmax(0.0, determinant);
            const double root = sqrt(determinant);
            // one division per bond instead of three: both roots share the denominator, and 1 / dtf joins it (the
            // quotients agree with the reference's to an ulp; FP64 divisions and the root are what this kernel costs)
            const double inv = 1.0 / (2.0 * qa * dtf);
            const double lambda1 = (-qb + root) * inv;
            const double lambda2 = (-qb - root) * inv;
            const double lambda = (fabs(lambda1) < fabs(lambda2)) ? lambda1 : lambda2;
#pragma unroll
            for (int d = 0; d < 3; ++d)
            {
                at(FORCE + d, i) += lambda * dist[d];
                at(FORCE + d, j) -= lambda * dist[d];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NA; ++k)
#pragma unroll
        for (int d = 0; d < 3; ++d) a.force[d][oc.x + k] = at(FORCE + d, k);
}

// NA consecutive doubles of a per-atom plane: one 32-byte access when a four-atom molecule starts on a multiple of four
// atoms (uniform molecules do), instead of four 8-byte accesses a 32-byte stride apart from the neighbouring lanes'
template <int NA>
__device__ __forceinline__ void loadPlane(const double* plane, long long first, double (&v)[NA])
{
    if (NA == 4 && (reinterpret_cast<uintptr_t>(plane + first) & 31) == 0)
    {
        const double4 q = ld4(reinterpret_cast<const double4*>(plane + first));
        v[0] = q.x;
        v[1] = q.y;
        v[2 % NA] = q.z;
        v[3 % NA] = q.w;
    }
    else
    {
#pragma unroll
        for (int k = 0; k < NA; ++k) v[k] = plane[first + k];
    }
}
template <int NA>
__device__ __forceinline__ void storePlane(double* plane, long long first, const double (&v)[NA])
{
    if (NA == 4 && (reinterpret_cast<uintptr_t>(plane + first) & 31) == 0)
        st4(reinterpret_cast<double4*>(plane + first), make_double4(v[0], v[1], v[2 % NA], v[3 % NA]));
    else
    {
#pragma unroll
        for (int k = 0; k < NA; ++k) plane[first + k] = v[k];
    }
}

// shakeFusedKernel for molecules whose bonds are ALL pairs of their NA atoms in lexicographic order (the rigid tetramers of
// BASELINE.json configs[3], three-site water): the atom indices of every bond are compile-time constants, so the whole
// molecule state lives in registers -- no shared memory (the shared-memory traffic of the indexable state, ~600 accesses per
// molecule, bounded the generic kernel) -- and the bonds of an iteration, whose inputs do not depend on each other
// (Shake.hpp:84-128 reads pos and updatedPos, both fixed during the bond loop), overlap.  The forces are accumulated in the
// reference's bond order.
template <int NA>
__global__ void __launch_bounds__(128)
    shakeAllPairsKernel(MolsView m, AtomsView a, int64_t numLocalMols, const double* __restrict__ bondDist,
                        int64_t numIterations, double dtv, double dtf, int* error)
{
    constexpr int NB = NA * (NA - 1) / 2;
    const int64_t mol = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (mol >= numLocalMols) return;
    const longlong2 oc = m.oc[mol];
    if (oc.y < NA)  // some bond names an atom the molecule does not have
    {
        *error = 1;
        return;
    }
    double pos[NA][3], base[NA][3], frc[NA][3], upd[NA][3], invMass[NA], eqSq[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b)
    {
        const double eq = bondDist[b];
        eqSq[b] = eq * eq;
    }
    {
        double vel[3][NA], f[3][NA], mass[NA];
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
            loadPlane<NA>(a.vel[d], oc.x, vel[d]);
            loadPlane<NA>(a.force[d], oc.x, f[d]);
        }
        loadPlane<NA>(a.mass, oc.x, mass);
#pragma unroll
        for (int k = 0; k < NA; ++k)
        {
            const double4 p = ld4(a.pos + oc.x + k);
            pos[k][0] = p.x;
            pos[k][1] = p.y;
            pos[k][2] = p.z;
#pragma unroll
            for (int d = 0; d < 3; ++d)
            {
                base[k][d] = pos[k][d] + dtv * vel[d][k];
                frc[k][d] = f[d][k];
            }
            invMass[k] = 1.0 / mass[k];  // the same quotient for every bond of the atom
        }
    }
    for (int64_t it = 0; it < numIterations; ++it)
    {
#pragma unroll
        for (int k = 0; k < NA; ++k)  // Shake::operator()(UnconstraintUpdate, idx), Shake.hpp:131-137
        {
            const double dtfm = dtf * invMass[k];  // dtf / mass up to an ulp
#pragma unroll
            for (int d = 0; d < 3; ++d) upd[k][d] = base[k][d] + dtfm * frc[k][d];
        }
        int b = 0;
#pragma unroll
        for (int i = 0; i < NA; ++i)
#pragma unroll
            for (int j = i + 1; j < NA; ++j, ++b)  // Shake::enforcePositionalConstraint, :84-128
            {
                const double dist[3] = {pos[i][0] - pos[j][0], pos[i][1] - pos[j][1], pos[i][2] - pos[j][2]};
                const double distSq = dist[0] * dist[0] + dist[1] * dist[1] + dist[2] * dist[2];
                const double du[3] = {upd[i][0] - upd[j][0], upd[i][1] - upd[j][1], upd[i][2] - upd[j][2]};
                const double updSq = du[0] * du[0] + du[1] * du[1] + du[2] * du[2];
                const double im = invMass[i] + invMass[j];
                const double qa = im * im * distSq;
                const double qb = 2.0 * im * (du[0] * dist[0] + du[1] * dist[1] + du[2] * dist[2]);
                const double qc = updSq - eqSq[b];
                const double determinant = fmax(0.0, qb * qb - 4.0 * qa * qc);
                const double root = sqrt(determinant);
                const double inv = 1.0 / (2.0 * qa * dtf);  // one division per bond, see shakeFusedKernel
                const double lambda1 = (-qb + root) * inv;
                const double lambda2 = (-qb - root) * inv;
                const double lambda = (fabs(lambda1) < fabs(lambda2)) ? lambda1 : lambda2;
#pragma unroll
                for (int d = 0; d < 3; ++d)
                {
                    frc[i][d] += lambda * dist[d];
                    frc[j][d] -= lambda * dist[d];
                }
            }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
        double out[NA];
#pragma unroll
        for (int k = 0; k < NA; ++k) out[k] = frc[k][d];
        storePlane<NA>(a.force[d], oc.x, out);
    }
}

// MoleculeConstraints::enforceVelocityConstraints lambda (Shake.hpp:215-231) with
// Shake::enforceVelocityConstraint (:56-82) inlined
__global__ void shakeVelocityKernel(MolsView m, AtomsView a, int64_t numLocalMols, const long long* __restrict__ bondIdx,
                                    int64_t numBonds, int* error)
{
    const int64_t mol = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (mol >= numLocalMols) return;
    const longlong2 oc = m.oc[mol];
    for (int64_t b = 0; b < numBonds; ++b)
    {
        const long long bi = bondIdx[2 * b], bj = bondIdx[2 * b + 1];
        if (bi >= oc.y || bj >= oc.y)
        {
            *error = 1;
            return;
        }
        const long long idx = oc.x + bi, jdx = oc.x + bj;
        const double4 pi = ld4(a.pos + idx), pj = ld4(a.pos + jdx);
        const double dist[3] = {pi.x - pj.x, pi.y - pj.y, pi.z - pj.z};
        const double distSq = dist[0] * dist[0] + dist[1] * dist[1] + dist[2] * dist[2];
        const double invMassI = 1.0 / a.mass[idx], invMassJ = 1.0 / a.mass[jdx];
        const double reducedMass = 1.0 / (invMassI + invMassJ);
        const double relVel[3] = {a.vel[0][idx] - a.vel[0][jdx], a.vel[1][idx] - a.vel[1][jdx], a.vel[2][idx] - a.vel[2][jdx]};
        const double factor = (relVel[0] * dist[0] + relVel[1] * dist[1] + relVel[2] * dist[2]) / distSq * reducedMass;
        for (int d = 0; d < 3; ++d)
        {
            a.vel[d][idx] -= factor * dist[d] * invMassI;
            a.vel[d][jdx] += factor * dist[d] * invMassJ;
        }
    }
}

// enforceVelocityConstraints for bonds among the first NA <= 4 atoms of a molecule: {pos, vel, 1 / mass} of the
// molecule in shared memory, the bonds walked in order, the velocities written back once
// KICK: VelocityVerlet::postForceIntegrate (action/VelocityVerlet.cpp:72-90, the arithmetic of integratePostKernel)
// is applied to the velocities as they are loaded -- for molecules of exactly NA atoms this replaces the separate pass
template <int NA, bool KICK>
__global__ void __launch_bounds__(SHAKE_FUSED_THREADS)
    rattleFusedKernel(MolsView m, AtomsView a, int64_t numLocalMols, const long long* __restrict__ bondIdx, int64_t numBonds,
                      double halfDt, int* error)
{
    __shared__ double sm[7 * NA * SHAKE_FUSED_THREADS];
    const int64_t mol = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (mol >= numLocalMols) return;
    const longlong2 oc = m.oc[mol];
    if (oc.y < NA)
    {
        *error = 1;
        return;
    }
    auto at = [&](int slot, long long atom) -> double& { return sm[(slot * NA + atom) * SHAKE_FUSED_THREADS + threadIdx.x]; };
    enum { POS = 0, VEL = 3, INVMASS = 6 };
#pragma unroll
    for (int k = 0; k < NA; ++k)
    {
        const double4 p = ld4(a.pos + oc.x + k);
        at(POS, k) = p.x;
        at(POS + 1, k) = p.y;
        at(POS + 2, k) = p.z;
        const double mass = a.mass[oc.x + k];
        const double dtfm = halfDt / mass;
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
            double v = a.vel[d][oc.x + k];
            if (KICK) v = __dadd_rn(v, __dmul_rn(dtfm, a.force[d][oc.x + k]));
            at(VEL + d, k) = v;
        }
        at(INVMASS, k) = 1.0 / mass;
    }
    for (int64_t b = 0; b < numBonds; ++b)  // Shake::enforceVelocityConstraint, Shake.hpp:56-82
    {
        const long long i = bondIdx[2 * b], j = bondIdx[2 * b + 1];
        const double dist[3] = {at(POS, i) - at(POS, j), at(POS + 1, i) - at(POS + 1, j), at(POS + 2, i) - at(POS + 2, j)};
        const double distSq = dist[0] * dist[0] + dist[1] * dist[1] + dist[2] * dist[2];
        const double invMassI = at(INVMASS, i), invMassJ = at(INVMASS, j);
        const double relVel[3] = {at(VEL, i) - at(VEL, j), at(VEL + 1, i) - at(VEL + 1, j), at(VEL + 2, i) - at(VEL + 2, j)};
        // (relVel . dist) / distSq * reducedMass with reducedMass = 1 / (invMassI + invMassJ), as one division
        const double factor = (relVel[0] * dist[0] + relVel[1] * dist[1] + relVel[2] * dist[2]) / (distSq * (invMassI + invMassJ));
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
            at(VEL + d, i) -= factor * dist[d] * invMassI;
            at(VEL + d, j) += factor * dist[d] * invMassJ;
        }
    }
#pragma unroll
    for (int k = 0; k < NA; ++k)
#pragma unroll
        for (int d = 0; d < 3; ++d) a.vel[d][oc.x + k] = at(VEL + d, k);
}

// rattleFusedKernel for all-pairs molecules (see shakeAllPairsKernel): state in registers.  The bonds stay sequential
// (every bond reads the velocities the previous one wrote, Shake.hpp:56-82).
template <int NA, bool KICK>
__global__ void __launch_bounds__(128)
    rattleAllPairsKernel(MolsView m, AtomsView a, int64_t numLocalMols, double halfDt, int* error)
{
    const int64_t mol = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (mol >= numLocalMols) return;
    const longlong2 oc = m.oc[mol];
    if (oc.y < NA)
    {
        *error = 1;
        return;
    }
    double pos[NA][3], vel[NA][3], invMass[NA];
    {
        double v0[3][NA], f[3][NA], mass[NA];
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
            loadPlane<NA>(a.vel[d], oc.x, v0[d]);
            if (KICK) loadPlane<NA>(a.force[d], oc.x, f[d]);
        }
        loadPlane<NA>(a.mass, oc.x, mass);
#pragma unroll
        for (int k = 0; k < NA; ++k)
        {
            const double4 p = ld4(a.pos + oc.x + k);
            pos[k][0] = p.x;
            pos[k][1] = p.y;
            pos[k][2] = p.z;
            const double dtfm = halfDt / mass[k];
#pragma unroll
            for (int d = 0; d < 3; ++d)
            {
                double v = v0[d][k];
                if (KICK) v = __dadd_rn(v, __dmul_rn(dtfm, f[d][k]));
                vel[k][d] = v;
            }
            invMass[k] = 1.0 / mass[k];
        }
    }
#pragma unroll
    for (int i = 0; i < NA; ++i)
#pragma unroll
        for (int j = i + 1; j < NA; ++j)  // Shake::enforceVelocityConstraint, Shake.hpp:56-82
        {
            const double dist[3] = {pos[i][0] - pos[j][0], pos[i][1] - pos[j][1], pos[i][2] - pos[j][2]};
            const double distSq = dist[0] * dist[0] + dist[1] * dist[1] + dist[2] * dist[2];
            const double relVel[3] = {vel[i][0] - vel[j][0], vel[i][1] - vel[j][1], vel[i][2] - vel[j][2]};
            const double factor = (relVel[0] * dist[0] + relVel[1] * dist[1] + relVel[2] * dist[2]) / (distSq * (invMass[i] + invMass[j]));
#pragma unroll
            for (int d = 0; d < 3; ++d)
            {
                vel[i][d] -= factor * dist[d] * invMass[i];
                vel[j][d] += factor * dist[d] * invMass[j];
            }
        }
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
        double out[NA];
#pragma unroll
        for (int k = 0; k < NA; ++k) out[k] = vel[k][d];
        storePlane<NA>(a.vel[d], oc.x, out);
    }
}

static int checkBondError(int* dErr, cudaStream_t st)
{
    int h = 0;
    MB_CUDA(cudaMemcpyAsync(&h, dErr, 4, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));  // the reference fences after every kernel too
    MB_REQUIRE(h == 0, "not enough atoms in molecule to satisfy bond");
    return 0;
}

// the launches of enforcePositionalConstraints / enforceVelocityConstraints without the read-back of the bond-range
// flag (step-loop drivers validate the bonds once and must not synchronise every step)
int constraintsEnforcePositional(mrmd_b200_constraints* c, const mrmd_b200_molecules* m, mrmd_b200_atoms* a, double dt,
                                 cudaStream_t st)
{
    MB_CUDA(cudaMemsetAsync(c->dErr, 0, 4, st));
    const int64_t nAll = a->numLocal + a->numGhost;
    if (nAll == 0 || c->numIterations == 0) return 0;
    const double dtv = dt, dtf = 0.5 * dt * dt;  // Shake.hpp:148-149
    if (c->maxBondAtoms >= 2 && c->maxBondAtoms <= 4)
    {
        if (m->numLocal == 0 || c->numBonds == 0) return 0;
        if (c->allPairs && c->maxBondAtoms >= 3)
        {
            const int blocksAll = gridFor(m->numLocal, 128);
            if (c->maxBondAtoms == 3)
                shakeAllPairsKernel<3><<<blocksAll, 128, 0, st>>>(m->v, a->v, m->numLocal, c->bondDist.as<double>(),
                                                                 c->numIterations, dtv, dtf, c->dErr);
            else
                shakeAllPairsKernel<4><<<blocksAll, 128, 0, st>>>(m->v, a->v, m->numLocal, c->bondDist.as<double>(),
                                                                 c->numIterations, dtv, dtf, c->dErr);
            MB_LAUNCHED();
            return 0;
        }
        const int blocks = gridFor(m->numLocal, SHAKE_FUSED_THREADS);
#define MB_SHAKE_FUSED(NA)                                                                                              \
    shakeFusedKernel<NA><<<blocks, SHAKE_FUSED_THREADS, 0, st>>>(m->v, a->v, m->numLocal, c->bondIdx.as<long long>(),   \
                                                                 c->bondDist.as<double>(), c->numBonds, c->numIterations, \
                                                                 dtv, dtf, c->dErr)
        if (c->maxBondAtoms == 2) MB_SHAKE_FUSED(2);
        else if (c->maxBondAtoms == 3) MB_SHAKE_FUSED(3);
        else MB_SHAKE_FUSED(4);
#undef MB_SHAKE_FUSED
        MB_LAUNCHED();
        return 0;
    }
    MB_TRY(c->updatedPos.reserve(size_t(nAll) * 32));  // util::grow(updatedPos_, ...), Shake.hpp:146
    for (int64_t it = 0; it < c->numIterations; ++it)
    {
        shakeUnconstraintKernel<<<gridFor(nAll, 256), 256, 0, st>>>(a->v, nAll, dtv, dtf, c->updatedPos.as<double4>());
        MB_LAUNCHED();
        if (m->numLocal > 0 && c->numBonds > 0)
        {
            shakePositionalKernel<<<gridFor(m->numLocal, 128), 128, 0, st>>>(
                m->v, a->v, m->numLocal, c->bondIdx.as<long long>(), c->bondDist.as<double>(), c->numBonds,
                c->updatedPos.as<double4>(), dtf, c->dErr);
            MB_LAUNCHED();
        }
    }
    return 0;
}

void constraintsSetUniformMolecules(mrmd_b200_constraints* c, bool uniform) { c->uniformMolecules = uniform; }

// postDt > 0: postForceIntegrate(atoms, postDt) first.  It rides along in the fused kernel when every local atom belongs
// to a local molecule of exactly maxBondAtoms atoms (the step-loop drivers' uniform molecules), else it is its own pass.
int constraintsEnforceVelocity(mrmd_b200_constraints* c, const mrmd_b200_molecules* m, mrmd_b200_atoms* a, cudaStream_t st,
                               double postDt)
{
    MB_CUDA(cudaMemsetAsync(c->dErr, 0, 4, st));
    const bool fusedKernel = c->maxBondAtoms >= 2 && c->maxBondAtoms <= 4 && m->numLocal > 0 && c->numBonds > 0;
    const bool kick = postDt > 0.0 && fusedKernel && c->uniformMolecules && a->numLocal == m->numLocal * c->maxBondAtoms;
    if (postDt > 0.0 && !kick) MB_TRY(mrmd_b200_vv_post(a, postDt, st));
    if (m->numLocal == 0 || c->numBonds == 0) return 0;
    if (fusedKernel)
    {
        const int blocks = gridFor(m->numLocal, SHAKE_FUSED_THREADS);
        const double halfDt = 0.5 * postDt;
        if (c->allPairs && c->maxBondAtoms >= 3)
        {
            const int blocksAll = gridFor(m->numLocal, 128);
#define MB_RATTLE_ALL(NA)                                                                                             \
    do                                                                                                                \
    {                                                                                                                 \
        if (kick) rattleAllPairsKernel<NA, true><<<blocksAll, 128, 0, st>>>(m->v, a->v, m->numLocal, halfDt, c->dErr); \
        else rattleAllPairsKernel<NA, false><<<blocksAll, 128, 0, st>>>(m->v, a->v, m->numLocal, 0.0, c->dErr);       \
    } while (0)
            if (c->maxBondAtoms == 3) MB_RATTLE_ALL(3);
            else MB_RATTLE_ALL(4);
#undef MB_RATTLE_ALL
            MB_LAUNCHED();
            return 0;
        }
#define MB_RATTLE_FUSED(NA)                                                                                               \
    do                                                                                                                    \
    {                                                                                                                     \
        if (kick)                                                                                                         \
            rattleFusedKernel<NA, true><<<blocks, SHAKE_FUSED_THREADS, 0, st>>>(                                          \
                m->v, a->v, m->numLocal, c->bondIdx.as<long long>(), c->numBonds, halfDt, c->dErr);                       \
        else                                                                                                              \
            rattleFusedKernel<NA, false><<<blocks, SHAKE_FUSED_THREADS, 0, st>>>(                                         \
                m->v, a->v, m->numLocal, c->bondIdx.as<long long>(), c->numBonds, 0.0, c->dErr);                          \
    } while (0)
        if (c->maxBondAtoms == 2) MB_RATTLE_FUSED(2);
        else if (c->maxBondAtoms == 3) MB_RATTLE_FUSED(3);
        else MB_RATTLE_FUSED(4);
#undef MB_RATTLE_FUSED
        MB_LAUNCHED();
        return 0;
    }
    shakeVelocityKernel<<<gridFor(m->numLocal, 128), 128, 0, st>>>(m->v, a->v, m->numLocal, c->bondIdx.as<long long>(),
                                                                   c->numBonds, c->dErr);
    MB_LAUNCHED();
    return 0;
}
}  // namespace mrmd_b200

using namespace mrmd_b200;

extern "C" {

int mrmd_b200_berendsen_thermostat(mrmd_b200_atoms* a, double currentTemperature, double targetTemperature, double gamma,
                                   void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(a != nullptr, "berendsen_thermostat");
    if (currentTemperature <= 0.0) return 0;  // BerendsenThermostat.cpp:30-33
    MB_REQUIRE(targetTemperature > 0.0, "berendsen_thermostat: target temperature must be positive");
    const double beta = std::sqrt(1.0 + gamma * (targetTemperature / currentTemperature - 1.0));
    if (a->numLocal == 0) return 0;
    scaleVelocityKernel<<<gridFor(a->numLocal, 256), 256, 0, S(stream)>>>(a->v, a->numLocal, beta);
    MB_LAUNCHED();
    return 0;
}

int mrmd_b200_berendsen_barostat(mrmd_b200_atoms* a, double currentPressure, double targetPressure, double gamma,
                                 mrmd_b200_subdomain* s, int stretchX, int stretchY, int stretchZ, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(a != nullptr && s != nullptr, "berendsen_barostat");
    a->posEpoch += 1;
    const double mu = std::cbrt(1.0 + gamma * (currentPressure - targetPressure));
    if (stretchX) mrmd_b200_subdomain_scale_dim(s, mu, 0);
    if (stretchY) mrmd_b200_subdomain_scale_dim(s, mu, 1);
    if (stretchZ) mrmd_b200_subdomain_scale_dim(s, mu, 2);
    if (a->numLocal == 0) return 0;
    scalePositionKernel<<<gridFor(a->numLocal, 256), 256, 0, S(stream)>>>(a->v.pos, a->numLocal, mu, stretchX, stretchY,
                                                                          stretchZ);
    MB_LAUNCHED();
    return 0;
}

int mrmd_b200_limit_acceleration(mrmd_b200_atoms* a, double maxAccelerationPerComponent, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(a != nullptr, "limit_acceleration");
    if (a->numLocal == 0) return 0;
    limitAccelerationKernel<<<gridFor(a->numLocal, 256), 256, 0, S(stream)>>>(a->v, a->numLocal, maxAccelerationPerComponent);
    MB_LAUNCHED();
    return 0;
}

int mrmd_b200_limit_velocity(mrmd_b200_atoms* a, double maxVelocityPerComponent, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(a != nullptr, "limit_velocity");
    if (a->numLocal == 0) return 0;
    limitVelocityKernel<<<gridFor(a->numLocal, 256), 256, 0, S(stream)>>>(a->v, a->numLocal, maxVelocityPerComponent);
    MB_LAUNCHED();
    return 0;
}

int mrmd_b200_constraints_create(mrmd_b200_constraints** out, int64_t atomsPerMolecule, int64_t numConstraintIterations)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(out != nullptr && atomsPerMolecule > 0 && numConstraintIterations >= 0, "constraints_create");
    auto* c = new mrmd_b200_constraints;
    c->atomsPerMolecule = atomsPerMolecule;
    c->numIterations = numConstraintIterations;
    if (cudaMalloc(&c->dErr, 4) != cudaSuccess)
    {
        delete c;
        setLastError("constraints_create: out of device memory");
        return MRMD_B200_ENOMEM;
    }
    *out = c;
    return 0;
}

int mrmd_b200_constraints_destroy(mrmd_b200_constraints* c)
{
    if (c == nullptr) return 0;
    cudaDeviceSynchronize();
    c->bondIdx.release();
    c->bondDist.release();
    c->updatedPos.release();
    if (c->dErr) cudaFree(c->dErr);
    delete c;
    return 0;
}

int mrmd_b200_constraints_set(mrmd_b200_constraints* c, const int64_t* idx, const int64_t* jdx, const double* eqDistance,
                              int64_t numBonds)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(c != nullptr && numBonds >= 0 && (numBonds == 0 || (idx && jdx && eqDistance)), "constraints_set");
    std::vector<long long> pairs(static_cast<size_t>(2 * numBonds));
    c->maxBondAtoms = 0;
    for (int64_t b = 0; b < numBonds; ++b)
    {
        MB_REQUIRE(idx[b] >= 0 && jdx[b] >= 0, "constraints_set: negative atom index");
        pairs[static_cast<size_t>(2 * b)] = idx[b];
        pairs[static_cast<size_t>(2 * b + 1)] = jdx[b];
        c->maxBondAtoms = std::max<int64_t>(c->maxBondAtoms, std::max(idx[b], jdx[b]) + 1);
    }
    MB_TRY(c->bondIdx.reserve(std::max<size_t>(pairs.size() * 8, 16)));
    MB_TRY(c->bondDist.reserve(std::max<size_t>(size_t(numBonds) * 8, 8)));
    if (numBonds > 0)
    {
        MB_CUDA(cudaMemcpy(c->bondIdx.p, pairs.data(), pairs.size() * 8, cudaMemcpyHostToDevice));
        MB_CUDA(cudaMemcpy(c->bondDist.p, eqDistance, size_t(numBonds) * 8, cudaMemcpyHostToDevice));
    }
    c->numBonds = numBonds;
    {
        const int64_t na = c->maxBondAtoms;
        bool all = na >= 2 && na <= 4 && numBonds == na * (na - 1) / 2;
        int64_t b = 0;
        for (int64_t i = 0; all && i < na; ++i)
            for (int64_t j = i + 1; all && j < na; ++j, ++b) all = (idx[b] == i && jdx[b] == j);
        c->allPairs = all;
    }
    return 0;
}

int mrmd_b200_constraints_enforce_positional(mrmd_b200_constraints* c, const mrmd_b200_molecules* m, mrmd_b200_atoms* a,
                                             double dt, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(c != nullptr && m != nullptr && a != nullptr, "constraints_enforce_positional");
    MB_TRY(constraintsEnforcePositional(c, m, a, dt, S(stream)));
    return checkBondError(c->dErr, S(stream));
}

int mrmd_b200_constraints_enforce_velocity(mrmd_b200_constraints* c, const mrmd_b200_molecules* m, mrmd_b200_atoms* a,
                                           double dt, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(c != nullptr && m != nullptr && a != nullptr, "constraints_enforce_velocity");
    (void)dt;  // Shake(atoms, dt) only feeds the positional update (Shake.hpp:139-150)
    MB_TRY(constraintsEnforceVelocity(c, m, a, S(stream), 0.0));
    return checkBondError(c->dErr, S(stream));
}

}  // extern "C"
