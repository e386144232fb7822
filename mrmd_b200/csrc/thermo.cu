// thermo.cu -- thermodynamic force: axial density histogram, table update, application.
// Reference: mrmd/action/ThermodynamicForce.hpp:32-216, ThermodynamicForce.cpp:25-130,
//            mrmd/analysis/AxialDensityProfile.cpp:21-51, mrmd/data/MultiHistogram.hpp:29-92,
//            MultiHistogram.cpp:48-212, mrmd/util/interpolation.hpp:34-38.
#include <algorithm>
#include <cmath>
#include <vector>

#include "handles.cuh"


namespace mrmd_b200
{
constexpr int TH_THREADS = 256;
constexpr int TH_MAX_SMEM_BINS = 12288;  // 48 KB of 32-bit counters

// analysis::getAxialDensityProfile (AxialDensityProfile.cpp:36-47) accumulated straight into the running
// density profile (ThermodynamicForce.cpp:76).  Counts are privatised per block in shared memory as
// integers and flushed once; integer-valued double adds make the result independent of the order.
template <bool SMEM>
__global__ void __launch_bounds__(TH_THREADS)
    densityHistogramKernel(const double4* pos, int64_t numAtoms, double min, double inverseBinSize, long long numBins,
                           long long numTypes, double* hist)
{
    extern __shared__ unsigned int sCount[];
    const long long total = numBins * numTypes;
    if (SMEM)
    {
        for (long long b = threadIdx.x; b < total; b += TH_THREADS) sCount[b] = 0;
        __syncthreads();
    }
    for (int64_t idx = blockIdx.x * int64_t(TH_THREADS) + threadIdx.x; idx < numAtoms; idx += int64_t(gridDim.x) * TH_THREADS)
    {
        const double4 p = ld4nc(pos + idx);
        const long long bin = histBin(min, inverseBinSize, numBins, p.x);
        if (bin == -1) continue;
        const long long slot = bin * numTypes + typeOf(p);
        if (SMEM) atomicAdd(sCount + slot, 1u);
        else atomicAdd(hist + slot, 1.0);
    }
    if (SMEM)
    {
        __syncthreads();
        for (long long b = threadIdx.x; b < total; b += TH_THREADS)
        {
            const unsigned int c = sCount[b];
            if (c != 0) atomicAdd(hist + b, double(c));
        }
    }
}

// ThermodynamicForce::update_if (ThermodynamicForce.hpp:124-152) in one single-block kernel; the phases are
// the reference's MultiHistogram free functions, separated by block barriers.
__global__ void __launch_bounds__(TH_THREADS)
    thermoUpdateKernel(double* force, double* density, double* smooth, double* grad, const double* forceFactor,
                       long long numBins, long long numTypes, double min, double binSize, double inverseBinSize,
                       double normalizationFactor, double sigma, double range, int enforceSymmetry, int periodic,
                       mrmd_b200_pred pred)
{
    const long long total = numBins * numTypes;
    if (enforceSymmetry)  // MultiHistogram::makeSymmetric, MultiHistogram.cpp:76-90
    {
        const long long maxIdx = numBins - 1;
        for (long long t = threadIdx.x; t < (numBins / 2) * numTypes; t += TH_THREADS)
        {
            const long long i = t / numTypes, j = t % numTypes;
            const double val = 0.5 * (density[i * numTypes + j] + density[(maxIdx - i) * numTypes + j]);
            density[i * numTypes + j] = val;
            density[(maxIdx - i) * numTypes + j] = val;
        }
        __syncthreads();
    }
    for (long long t = threadIdx.x; t < total; t += TH_THREADS) density[t] *= normalizationFactor;  // scale, :48-59
    __syncthreads();
    {
        // smoothen, MultiHistogram.cpp:163-212 (no 1/2 in the exponent, as in the reference)
        const double inverseSigma = 1.0 / sigma;
        const long long delta = static_cast<int>(range * sigma * inverseBinSize);
        for (long long t = threadIdx.x; t < total; t += TH_THREADS)
        {
            const long long b = t / numTypes, h = t % numTypes;
            double normalization = 0.0, acc = 0.0;
            long long jMin = b - delta, jMax = b + delta;
            if (!periodic)
            {
                jMin = (jMin < 0) ? 0 : jMin;
                jMax = (jMax > numBins - 1) ? numBins - 1 : jMax;
            }
            for (long long j = jMin; j <= jMax; ++j)
            {
                long long mapped = j;
                if (periodic)
                {
                    if (mapped < 0) mapped += numBins;
                    if (mapped >= numBins) mapped -= numBins;
                }
                const double u = double(b - j) * binSize * inverseSigma;
                const double eFunc = exp(-(u * u));
                normalization += eFunc;
                acc += density[mapped * numTypes + h] * eFunc;
            }
            smooth[t] = acc / normalization;
        }
    }
    __syncthreads();
    {
        // gradient, MultiHistogram.cpp:113-161, then scale(forceFactor) :61-74 and replace_if_bin_position
        const double inverseSpacing = inverseBinSize;
        const double inverseDoubleSpacing = 0.5 * inverseBinSize;
        for (long long t = threadIdx.x; t < total; t += TH_THREADS)
        {
            const long long i = t / numTypes, j = t % numTypes;
            double g;
            if (i == 0)
                g = periodic ? (smooth[(i + 1) * numTypes + j] - smooth[(numBins - 1) * numTypes + j]) * inverseDoubleSpacing
                             : (smooth[(i + 1) * numTypes + j] - smooth[i * numTypes + j]) * inverseSpacing;
            else if (i == numBins - 1)
                g = periodic ? (smooth[j] - smooth[(i - 1) * numTypes + j]) * inverseDoubleSpacing
                             : (smooth[i * numTypes + j] - smooth[(i - 1) * numTypes + j]) * inverseSpacing;
            else
                g = (smooth[(i + 1) * numTypes + j] - smooth[(i - 1) * numTypes + j]) * inverseDoubleSpacing;
            g *= forceFactor[j];
            const double x = min + (double(i) + 0.5) * binSize;  // getBinPosition, MultiHistogram.hpp:68-74
            if (!pred1(pred, x, x, x)) g = 0.0;
            grad[t] = g;
        }
    }
    __syncthreads();
    for (long long t = threadIdx.x; t < total; t += TH_THREADS)
    {
        force[t] -= grad[t];
        density[t] = 0.0;
    }
}

// apply_if (ThermodynamicForce.hpp:98-122) / applyInterpolated_if (:154-216): x force of LOCAL atoms only
template <bool INTERPOLATED, bool PRED>
__global__ void __launch_bounds__(TH_THREADS)
    thermoApplyKernel(AtomsView a, int64_t numLocal, const double* __restrict__ force, double min, double binSize,
                      double inverseBinSize, long long numBins, long long numTypes, mrmd_b200_pred pred)
{
    const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (idx >= numLocal) return;
    const double4 p = ld4nc(a.pos + idx);
    if (PRED && !pred1(pred, p.x, p.y, p.z)) return;
    const long long bin = histBin(min, inverseBinSize, numBins, p.x);
    if (bin == -1) return;
    const long long type = typeOf(p);
    double add;
    if (!INTERPOLATED)
    {
        add = force[bin * numTypes + type];
    }
    else
    {
        const double binStart = min + double(bin) * binSize;
        const double fracInBin = (p.x - binStart) * inverseBinSize;
        long long left, right;
        double factor;
        if (fracInBin < 0.5)
        {
            left = bin - 1;
            right = bin;
            factor = fracInBin + 0.5;
        }
        else
        {
            left = bin;
            right = bin + 1;
            factor = fracInBin - 0.5;
        }
        if (left >= 0 && right < numBins)
        {
            const double l = force[left * numTypes + type];
            const double r = force[right * numTypes + type];
            add = l + (r - l) * factor;  // util::lerp
        }
        else
            add = force[bin * numTypes + type];
    }
    a.force[0][idx] += add;
}
}  // namespace mrmd_b200

using namespace mrmd_b200;

extern "C" {

int mrmd_b200_thermo_create(mrmd_b200_thermo** out, const double* targetDensity, int64_t numTypes,
                            const mrmd_b200_subdomain* s, double requestedDensityBinWidth, const double* modulation,
                            int enforceSymmetry, int usePeriodicity)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(out && targetDensity && s && modulation, "thermo_create");
    MB_REQUIRE(numTypes > 0, "thermo_create: numTypes must be positive");           // ThermodynamicForce.cpp:49
    MB_REQUIRE(requestedDensityBinWidth > 0.0, "thermo_create: bin width must be positive");
    MB_REQUIRE(s->maxCorner[0] > s->minCorner[0], "thermo_create: empty subdomain");  // MultiHistogram.hpp:44
    auto* t = new mrmd_b200_thermo;
    t->min = s->minCorner[0];
    t->max = s->maxCorner[0];
    t->numBins = static_cast<int64_t>(std::ceil(s->diameter[0] / requestedDensityBinWidth));
    t->numTypes = numTypes;
    t->binSize = (t->max - t->min) / double(t->numBins);
    t->inverseBinSize = 1.0 / t->binSize;
    t->binVolume = s->diameter[1] * s->diameter[2] * t->binSize;
    t->enforceSymmetry = enforceSymmetry;
    t->usePeriodicity = usePeriodicity;
    const size_t bytes = size_t(t->numBins) * numTypes * 8;
    double* slab = nullptr;
    if (cudaMalloc(&slab, 4 * bytes + size_t(numTypes) * 8) != cudaSuccess)
    {
        delete t;
        setLastError("thermo_create: allocation failed");
        return MRMD_B200_ENOMEM;
    }
    cudaMemset(slab, 0, 4 * bytes + size_t(numTypes) * 8);
    t->force = slab;
    t->density = slab + t->numBins * numTypes;
    t->tmpA = slab + 2 * t->numBins * numTypes;
    t->tmpB = slab + 3 * t->numBins * numTypes;
    t->forceFactor = slab + 4 * t->numBins * numTypes;
    std::vector<double> ff(static_cast<size_t>(numTypes));
    for (int64_t i = 0; i < numTypes; ++i) ff[i] = modulation[i] / targetDensity[i];  // :52-55
    cudaMemcpy(t->forceFactor, ff.data(), size_t(numTypes) * 8, cudaMemcpyHostToDevice);
    *out = t;
    return 0;
}

int mrmd_b200_thermo_destroy(mrmd_b200_thermo* t)
{
    if (t == nullptr) return 0;
    cudaDeviceSynchronize();
    if (t->force) cudaFree(t->force);
    delete t;
    return 0;
}

int mrmd_b200_thermo_info(const mrmd_b200_thermo* t, int64_t* numBins, int64_t* numTypes, double* binSize,
                          int64_t* samples)
{
    MB_REQUIRE(t != nullptr, "thermo_info");
    if (numBins) *numBins = t->numBins;
    if (numTypes) *numTypes = t->numTypes;
    if (binSize) *binSize = t->binSize;
    if (samples) *samples = t->samples;
    return 0;
}

int mrmd_b200_thermo_sample(mrmd_b200_thermo* t, const mrmd_b200_atoms* a, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(t != nullptr && a != nullptr, "thermo_sample");
    const int64_t n = a->numLocal;
    if (n > 0)
    {
        const int64_t total = t->numBins * t->numTypes;
        int blocks = std::min<int64_t>(gridFor(n, TH_THREADS), 148 * 8);
        if (total <= TH_MAX_SMEM_BINS)
            densityHistogramKernel<true><<<blocks, TH_THREADS, size_t(total) * 4, S(stream)>>>(
                a->v.pos, n, t->min, t->inverseBinSize, t->numBins, t->numTypes, t->density);
        else
            densityHistogramKernel<false><<<blocks, TH_THREADS, 0, S(stream)>>>(a->v.pos, n, t->min, t->inverseBinSize,
                                                                                t->numBins, t->numTypes, t->density);
        MB_LAUNCHED();
    }
    t->samples += 1;
    return 0;
}

int mrmd_b200_thermo_update(mrmd_b200_thermo* t, double smoothingSigma, double smoothingIntensity,
                            const mrmd_b200_pred* pred, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(t != nullptr, "thermo_update");
    MB_REQUIRE(t->samples > 0, "thermo_update: no density samples (MRMD_HOST_CHECK_GREATER, ThermodynamicForce.hpp:129)");
    mrmd_b200_pred p{};
    if (pred != nullptr) p = *pred;
    const double normalizationFactor = 1.0 / (t->binVolume * double(t->samples));  // :136
    thermoUpdateKernel<<<1, TH_THREADS, 0, S(stream)>>>(t->force, t->density, t->tmpA, t->tmpB, t->forceFactor, t->numBins,
                                                        t->numTypes, t->min, t->binSize, t->inverseBinSize,
                                                        normalizationFactor, smoothingSigma, smoothingIntensity,
                                                        t->enforceSymmetry, t->usePeriodicity, p);
    MB_LAUNCHED();
    t->samples = 0;
    return 0;
}

int mrmd_b200_thermo_apply(const mrmd_b200_thermo* t, mrmd_b200_atoms* a, const mrmd_b200_pred* pred, int interpolated,
                           void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(t != nullptr && a != nullptr, "thermo_apply");
    const int64_t n = a->numLocal;
    if (n == 0) return 0;
    mrmd_b200_pred p{};
    const bool usePred = (pred != nullptr && pred->kind != MRMD_B200_PRED_ALWAYS);
    if (usePred) p = *pred;
    const int blocks = gridFor(n, TH_THREADS);
#define TH_LAUNCH(I, P)                                                                                              \
    thermoApplyKernel<I, P><<<blocks, TH_THREADS, 0, S(stream)>>>(a->v, n, t->force, t->min, t->binSize, t->inverseBinSize, \
                                                                  t->numBins, t->numTypes, p)
    if (interpolated) { if (usePred) TH_LAUNCH(true, true); else TH_LAUNCH(true, false); }
    else { if (usePred) TH_LAUNCH(false, true); else TH_LAUNCH(false, false); }
#undef TH_LAUNCH
    MB_LAUNCHED();
    return 0;
}

int mrmd_b200_thermo_read(const mrmd_b200_thermo* t, int kind, double* dstHost, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(t != nullptr && dstHost != nullptr && (kind == 0 || kind == 1), "thermo_read");
    MB_CUDA(cudaMemcpyAsync(dstHost, kind == 0 ? t->force : t->density, size_t(t->numBins) * t->numTypes * 8,
                            cudaMemcpyDeviceToHost, S(stream)));
    MB_CUDA(cudaStreamSynchronize(S(stream)));
    return 0;
}

int mrmd_b200_thermo_write_force(mrmd_b200_thermo* t, const double* srcHost, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(t != nullptr && srcHost != nullptr, "thermo_write_force");
    MB_CUDA(cudaMemcpyAsync(t->force, srcHost, size_t(t->numBins) * t->numTypes * 8, cudaMemcpyHostToDevice, S(stream)));
    MB_CUDA(cudaStreamSynchronize(S(stream)));
    return 0;
}

int mrmd_b200_thermo_density_ptr(mrmd_b200_thermo* t, double** devicePtr, int64_t* count)
{
    MB_REQUIRE(t != nullptr && devicePtr != nullptr, "thermo_density_ptr");
    *devicePtr = t->density;
    if (count) *count = t->numBins * t->numTypes;
    return 0;
}

int mrmd_b200_thermo_mu(const mrmd_b200_thermo* t, double* muLeftHost, double* muRightHost, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(t != nullptr && muLeftHost != nullptr && muRightHost != nullptr, "thermo_mu");
    std::vector<double> f(static_cast<size_t>(t->numBins * t->numTypes));
    MB_TRY(mrmd_b200_thermo_read(t, 0, f.data(), stream));
    for (int64_t ty = 0; ty < t->numTypes; ++ty)  // ThermodynamicForce.cpp:98-130
    {
        double l = 0.0, r = 0.0;
        for (int64_t i = 0; i < t->numBins / 2; ++i) l += f[i * t->numTypes + ty];
        for (int64_t i = t->numBins / 2; i < t->numBins; ++i) r += f[i * t->numTypes + ty];
        muLeftHost[ty] = l * t->binSize;
        muRightHost[ty] = r * t->binSize;
    }
    return 0;
}

}  // extern "C"
