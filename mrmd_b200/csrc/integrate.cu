// integrate.cu -- velocity-Verlet and the BAOAB Langevin integrator.
// Reference: mrmd/action/VelocityVerlet.cpp:26-90, mrmd/action/UpdateSteps.hpp:25-79,
//            mrmd/action/VelocityVerletLangevinThermostat.hpp:63-136.
// Streaming kernels, HBM bound: pre reads pos4(32)+vel(24)+force(24)+mass(8), writes pos4(32)+vel(24);
// post reads vel+force+mass, writes vel.  One thread per atom, 256-bit pos4 access, planes coalesced.
#include "handles.cuh"

namespace mrmd_b200
{
// Philox4x32-10 (Salmon et al. SC'11), identical to oracle/mrmd_oracle.cpp:or_philox4x32
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t* out)
{
#pragma unroll
    for (int round = 0; round < 10; ++round)
    {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0);
        const uint32_t lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2);
        const uint32_t lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

// three standard normals for (seed, step, global atom id): 32-bit uniforms, Box-Muller
__device__ __forceinline__ void philoxNormals3(uint64_t seed, uint64_t step, uint64_t idx, double& n0, double& n1,
                                               double& n2)
{
    uint32_t r[4];
    philox4x32_10(uint32_t(idx), uint32_t(idx >> 32), uint32_t(step), uint32_t(step >> 32), uint32_t(seed),
                  uint32_t(seed >> 32), r);
    const double scale = 2.3283064365386963e-10;  // 2^-32
    const double u0 = (double(r[0]) + 0.5) * scale;
    const double u1 = (double(r[1]) + 0.5) * scale;
    const double u2 = (double(r[2]) + 0.5) * scale;
    const double u3 = (double(r[3]) + 0.5) * scale;
    const double ra = sqrt(-2.0 * log(u0));
    const double rb = sqrt(-2.0 * log(u2));
    double s, c;
    sincospi(2.0 * u1, &s, &c);
    n0 = ra * c;
    n1 = ra * s;
    n2 = rb * cospi(2.0 * u3);
}

__device__ __forceinline__ void blockMaxToGlobal(double v, double* dMax)
{
    __shared__ double sMax[32];
    v = warpMax(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sMax[warp] = v;
    __syncthreads();
    if (warp == 0)
    {
        const int nw = (blockDim.x + 31) >> 5;
        v = (lane < nw) ? sMax[lane] : 0.0;
        v = warpMax(v);
        // non-negative doubles order like their bit patterns: integer atomicMax is an exact max
        if (lane == 0 && v > 0.0)
            atomicMax(reinterpret_cast<unsigned long long*>(dMax), static_cast<unsigned long long>(__double_as_longlong(v)));
    }
}

// LANGEVIN = false: VelocityVerlet::preForceIntegrate;  true: ...LangevinThermostat::preForceIntegrate_apply_if
// FUSED_POST: the previous step's postForceIntegrate (same dt, same force) is applied first, as its own rounded
// addition, so the step loops save one pass over vel / force / mass per step
template <bool LANGEVIN, bool FUSED_POST>
__global__ void __launch_bounds__(256)
    integratePreKernel(AtomsView a, int64_t n, double dt, double zeta, double temperature, uint64_t seed,
                       uint64_t step, mrmd_b200_pred pred, double* dMax, const int* __restrict__ stop)
{
    // speculatively queued steps (md.cu / slab.cu): a step behind the one that asked for a rebuild must not run
    if (stop != nullptr && *stop != 0) return;
    const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    double distSqr = 0.0;
    if (idx < n)
    {
        const double dtHalf = 0.5 * dt;
        double4 p = ld4(a.pos + idx);
        const double ox = p.x, oy = p.y, oz = p.z;
        double vx = a.vel[0][idx], vy = a.vel[1][idx], vz = a.vel[2][idx];
        const double m = a.mass[idx];
        const double dtfm = dtHalf / m;  // updateKick, UpdateSteps.hpp:36-39
        // explicit rounding (no FMA contraction): the fused and the separate kick give identical bits
        const double kx = __dmul_rn(dtfm, a.force[0][idx]), ky = __dmul_rn(dtfm, a.force[1][idx]),
                     kz = __dmul_rn(dtfm, a.force[2][idx]);
        if (FUSED_POST)
        {
            vx = __dadd_rn(vx, kx);  // postForceIntegrate of the previous step, VelocityVerlet.cpp:69-90
            vy = __dadd_rn(vy, ky);
            vz = __dadd_rn(vz, kz);
        }
        vx = __dadd_rn(vx, kx);
        vy = __dadd_rn(vy, ky);
        vz = __dadd_rn(vz, kz);
        if (!LANGEVIN)
        {
            p.x += dt * vx;  // updateDrift, UpdateSteps.hpp:51-53
            p.y += dt * vy;
            p.z += dt * vz;
        }
        else
        {
            p.x += dtHalf * vx;
            p.y += dtHalf * vy;
            p.z += dtHalf * vz;
            if (pred1(pred, p.x, p.y, p.z))
            {
                double r0, r1, r2;
                philoxNormals3(seed, step, uint64_t(a.gid[idx]), r0, r1, r2);
                const double dtm = dt / m;  // updateOrnsteinUhlenbeck, UpdateSteps.hpp:68-78
                const double damping = exp(-zeta * dtm);
                const double sigma = sqrt(temperature / m * (1.0 - exp(-2.0 * zeta * dtm)));
                vx *= damping;
                vy *= damping;
                vz *= damping;
                vx += sigma * r0;
                vy += sigma * r1;
                vz += sigma * r2;
            }
            p.x += dtHalf * vx;
            p.y += dtHalf * vy;
            p.z += dtHalf * vz;
        }
        st4(a.pos + idx, p);
        a.vel[0][idx] = vx;
        a.vel[1][idx] = vy;
        a.vel[2][idx] = vz;
        const double dx = ox - p.x, dy = oy - p.y, dz = oz - p.z;
        distSqr = dx * dx + dy * dy + dz * dz;
        // a NaN force or position must not vanish in the maximum: report it as an infinite displacement
        if (!(distSqr == distSqr)) distSqr = __longlong_as_double(0x7ff0000000000000LL);
    }
    blockMaxToGlobal(distSqr, dMax);
}

__global__ void __launch_bounds__(256) integratePostKernel(AtomsView a, int64_t n, double dt)
{
    const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (idx >= n) return;
    const double dtfm = (0.5 * dt) / a.mass[idx];
    a.vel[0][idx] = __dadd_rn(a.vel[0][idx], __dmul_rn(dtfm, a.force[0][idx]));
    a.vel[1][idx] = __dadd_rn(a.vel[1][idx], __dmul_rn(dtfm, a.force[1][idx]));
    a.vel[2][idx] = __dadd_rn(a.vel[2][idx], __dmul_rn(dtfm, a.force[2][idx]));
}

static int fetchMaxDisp(mrmd_b200_atoms* a, double* maxDisplacement, cudaStream_t st)
{
    if (maxDisplacement == nullptr) return 0;
    MB_CUDA(cudaMemcpyAsync(a->hMaxDisp, a->dMaxDisp, 8, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    *maxDisplacement = std::sqrt(*a->hMaxDisp);  // VelocityVerlet.cpp:66
    return 0;
}

// shared by the C ABI and the step-loop drivers (fusedPost: see integratePreKernel)
// the rebuild decision of examples/02:138-143 on the device, for steps queued ahead of the host: accumulates
// sqrt(max |dx|^2) and raises stop = {1, localStep} when the sum reaches the threshold (skin / 2) -- the kernels of the
// steps queued behind it then return at once.  A non-finite displacement stops with stop[2] = 1.
__global__ void displacementDecisionKernel(const double* __restrict__ maxDispSqr, double* accum, double threshold, int* stop,
                                           int localStep)
{
    if (stop[0] != 0) return;
    const double v = *maxDispSqr;
    if (!(v == v) || v > 1.7e308)
    {
        stop[0] = 1;
        stop[1] = localStep;
        stop[2] = 1;
        return;
    }
    const double sum = *accum + sqrt(v);  // the host's  maxDisplacement += sqrt(max)  in the same order
    *accum = sum;
    if (sum >= threshold)
    {
        stop[0] = 1;
        stop[1] = localStep;
    }
}

int displacementDecision(const double* dMaxDispSqr, double* dAccum, double threshold, int* dStop, int localStep, cudaStream_t st)
{
    displacementDecisionKernel<<<1, 1, 0, st>>>(dMaxDispSqr, dAccum, threshold, dStop, localStep);
    MB_LAUNCHED();
    return 0;
}

int integratePre(mrmd_b200_atoms* a, double dt, bool langevin, double zeta, double temperature, uint64_t seed,
                 uint64_t step, const mrmd_b200_pred* pred, bool fusedPost, cudaStream_t st, const int* stop)
{
    // (a queued step that is stopped leaves the scalar cleared: its decision kernel is stopped as well)
    MB_CUDA(cudaMemsetAsync(a->dMaxDisp, 0, 8, st));
    a->posEpoch += 1;
    if (a->numLocal == 0) return 0;
    mrmd_b200_pred p{};
    if (pred != nullptr) p = *pred;
    const int blocks = gridFor(a->numLocal, 256);
#define PRE_LAUNCH(L, F) \
    integratePreKernel<L, F><<<blocks, 256, 0, st>>>(a->v, a->numLocal, dt, zeta, temperature, seed, step, p, a->dMaxDisp, stop)
    if (langevin) { if (fusedPost) PRE_LAUNCH(true, true); else PRE_LAUNCH(true, false); }
    else { if (fusedPost) PRE_LAUNCH(false, true); else PRE_LAUNCH(false, false); }
#undef PRE_LAUNCH
    MB_LAUNCHED();
    return 0;
}
}  // namespace mrmd_b200

using namespace mrmd_b200;

extern "C" {

int mrmd_b200_vv_pre(mrmd_b200_atoms* a, double dt, double* maxDisplacement, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(a != nullptr, "vv_pre");
    cudaStream_t st = S(stream);
    MB_TRY(integratePre(a, dt, false, 0.0, 0.0, 0, 0, nullptr, false, st));
    return fetchMaxDisp(a, maxDisplacement, st);
}

int mrmd_b200_vv_post(mrmd_b200_atoms* a, double dt, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(a != nullptr, "vv_post");
    if (a->numLocal == 0) return 0;
    integratePostKernel<<<gridFor(a->numLocal, 256), 256, 0, S(stream)>>>(a->v, a->numLocal, dt);
    MB_LAUNCHED();
    return 0;
}

int mrmd_b200_langevin_pre(mrmd_b200_atoms* a, double dt, double zeta, double temperature, uint64_t seed,
                           uint64_t step, const mrmd_b200_pred* pred, double* maxDisplacement, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(a != nullptr, "langevin_pre");
    cudaStream_t st = S(stream);
    MB_TRY(integratePre(a, dt, true, zeta, temperature, seed, step, pred, false, st));
    return fetchMaxDisp(a, maxDisplacement, st);
}

}  // extern "C"
