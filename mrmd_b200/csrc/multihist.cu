// multihist.cu -- data::MultiHistogram as a device-resident type behind the C ABI.
// Reference: mrmd/data/MultiHistogram.hpp:29-194 (the struct, transform, replace_if_bin_position),
//            mrmd/data/MultiHistogram.cpp:27-227 (operators, scale, makeSymmetric, cumulativeMovingAverage, gradient,
//            smoothen, createGrid), mrmd/action/ThermodynamicForce.hpp:57-63 (getForce / getDensityProfile return one),
//            mrmd/analysis/AxialDensityProfile.cpp:21-51 (getAxialDensityProfile returns one).
// data(bin, histogram) is stored row-major (bin slowest) like Kokkos' MultiView on the host.  The tables are tiny
// (numBins x numHistograms doubles): one thread per entry, results identical to the reference's loops entry by entry.
#include <cmath>

#include "handles.cuh"

struct mrmd_b200_hist
{
    double min = 0, max = 0;
    int64_t numBins = 0, numHistograms = 0;
    double binSize = 0, inverseBinSize = 0;
    double* data = nullptr;  // device, numBins x numHistograms
};

namespace mrmd_b200
{
constexpr int MH_THREADS = 128;

struct HistDev
{
    double* data;
    long long numBins, numHistograms;
    double min, binSize, inverseBinSize;
};
static HistDev dev(const mrmd_b200_hist* h) { return {h->data, h->numBins, h->numHistograms, h->min, h->binSize, h->inverseBinSize}; }

// transform(*this, rhs, *this, bin_op::...), MultiHistogram.cpp:27-46
__global__ void histTransformKernel(HistDev a, const double* __restrict__ rhs, int op)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= a.numBins * a.numHistograms) return;
    const double x = a.data[t], y = rhs[t];
    a.data[t] = (op == 0) ? x + y : ((op == 1) ? x - y : ((op == 2) ? x * y : x / y));
}

// scale(real_t) / scale(ScalarView), MultiHistogram.cpp:48-74
__global__ void histScaleKernel(HistDev a, double factor, const double* __restrict__ perHistogram)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= a.numBins * a.numHistograms) return;
    a.data[t] *= (perHistogram != nullptr) ? perHistogram[t % a.numHistograms] : factor;
}

// makeSymmetric, MultiHistogram.cpp:76-90
__global__ void histMakeSymmetricKernel(HistDev a)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (a.numBins / 2) * a.numHistograms) return;
    const long long i = t / a.numHistograms, j = t % a.numHistograms, maxIdx = a.numBins - 1;
    const double val = 0.5 * (a.data[i * a.numHistograms + j] + a.data[(maxIdx - i) * a.numHistograms + j]);
    a.data[i * a.numHistograms + j] = val;
    a.data[(maxIdx - i) * a.numHistograms + j] = val;
}

// cumulativeMovingAverage, MultiHistogram.cpp:92-111
__global__ void histMovingAverageKernel(HistDev average, const double* __restrict__ current, double factor)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= average.numBins * average.numHistograms) return;
    average.data[t] = (factor * average.data[t] + current[t]) / (factor + 1.0);
}

// gradient, MultiHistogram.cpp:113-161
__global__ void histGradientKernel(HistDev in, double* __restrict__ out, int periodic)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= in.numBins * in.numHistograms) return;
    const long long i = t / in.numHistograms, j = t % in.numHistograms, H = in.numHistograms, N = in.numBins;
    const double inverseSpacing = in.inverseBinSize, inverseDoubleSpacing = 0.5 * in.inverseBinSize;
    const double* d = in.data;
    double g;
    if (i == 0)
        g = periodic ? (d[(i + 1) * H + j] - d[(N - 1) * H + j]) * inverseDoubleSpacing : (d[(i + 1) * H + j] - d[i * H + j]) * inverseSpacing;
    else if (i == N - 1)
        g = periodic ? (d[j] - d[(i - 1) * H + j]) * inverseDoubleSpacing : (d[i * H + j] - d[(i - 1) * H + j]) * inverseSpacing;
    else
        g = (d[(i + 1) * H + j] - d[(i - 1) * H + j]) * inverseDoubleSpacing;
    out[t] = g;
}

// smoothen, MultiHistogram.cpp:163-212 (the Gaussian has no 1/2 in the exponent, as in the reference)
__global__ void histSmoothenKernel(HistDev in, double* __restrict__ out, double sigma, double range, int periodic)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= in.numBins * in.numHistograms) return;
    const long long b = t / in.numHistograms, h = t % in.numHistograms;
    const double inverseSigma = 1.0 / sigma;
    const long long delta = static_cast<int>(range * sigma * in.inverseBinSize);
    double normalization = 0.0, acc = 0.0;
    long long jMin = b - delta, jMax = b + delta;
    if (!periodic)
    {
        jMin = (jMin < 0) ? 0 : jMin;
        jMax = (jMax > in.numBins - 1) ? in.numBins - 1 : jMax;
    }
    for (long long j = jMin; j <= jMax; ++j)
    {
        long long mapped = j;
        if (periodic)
        {
            if (mapped < 0) mapped += in.numBins;
            if (mapped >= in.numBins) mapped -= in.numBins;
        }
        const double u = double(b - j) * in.binSize * inverseSigma;
        const double eFunc = exp(-(u * u));
        normalization += eFunc;
        acc += in.data[mapped * in.numHistograms + h] * eFunc;
    }
    out[t] = acc / normalization;
}

// replace_if_bin_position, MultiHistogram.hpp:176-191: the predicate sees the bin position as the coordinate of its axis
__global__ void histReplaceIfKernel(HistDev a, mrmd_b200_pred pred, double newValue)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= a.numBins * a.numHistograms) return;
    const double x = a.min + (double(t / a.numHistograms) + 0.5) * a.binSize;  // getBinPosition, :68-74
    if (pred1(pred, x, x, x)) a.data[t] = newValue;
}

// createGrid, MultiHistogram.cpp:214-227
__global__ void histGridKernel(HistDev a, double* grid)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < a.numBins) grid[i] = a.min + (double(i) + 0.5) * a.binSize;
}

static int histAlloc(mrmd_b200_hist** out, double min, double max, int64_t numBins, int64_t numHistograms)
{
    MB_REQUIRE(out != nullptr, "hist_create");
    MB_REQUIRE(max > min, "hist_create: max must be greater than min");  // MultiHistogram.hpp:44
    MB_REQUIRE(numBins >= 0 && numHistograms >= 0, "hist_create: negative extent");
    auto* h = new mrmd_b200_hist;
    h->min = min;
    h->max = max;
    h->numBins = numBins;
    h->numHistograms = numHistograms;
    h->binSize = (max - min) / double(numBins);
    h->inverseBinSize = 1.0 / h->binSize;
    const size_t bytes = size_t(std::max<int64_t>(numBins * numHistograms, 1)) * 8;
    if (cudaMalloc(&h->data, bytes) != cudaSuccess || cudaMemset(h->data, 0, bytes) != cudaSuccess)
    {
        if (h->data) cudaFree(h->data);
        delete h;
        setLastError("hist_create: out of device memory");
        return MRMD_B200_ENOMEM;
    }
    *out = h;
    return 0;
}
static int sameShape(const mrmd_b200_hist* a, const mrmd_b200_hist* b)
{
    MB_REQUIRE(a != nullptr && b != nullptr, "MultiHistogram: null handle");
    MB_REQUIRE(a->numBins == b->numBins && a->numHistograms == b->numHistograms, "MultiHistogram: extents differ");
    return 0;
}
static int entries(const mrmd_b200_hist* h) { return static_cast<int>(h->numBins * h->numHistograms); }
}  // namespace mrmd_b200

using namespace mrmd_b200;

extern "C" {

int mrmd_b200_hist_create(mrmd_b200_hist** out, double min, double max, int64_t numBins, int64_t numHistograms)
{
    MB_TRY(checkDevice());
    return histAlloc(out, min, max, numBins, numHistograms);
}

int mrmd_b200_hist_clone(mrmd_b200_hist** out, const mrmd_b200_hist* src, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(src != nullptr, "hist_clone");
    MB_TRY(histAlloc(out, src->min, src->max, src->numBins, src->numHistograms));
    if (entries(src) > 0)
        MB_CUDA(cudaMemcpyAsync((*out)->data, src->data, size_t(entries(src)) * 8, cudaMemcpyDeviceToDevice, S(stream)));
    return 0;
}

int mrmd_b200_hist_destroy(mrmd_b200_hist* h)
{
    if (h == nullptr) return 0;
    cudaDeviceSynchronize();
    if (h->data) cudaFree(h->data);
    delete h;
    return 0;
}

int mrmd_b200_hist_info(const mrmd_b200_hist* h, double* min, double* max, int64_t* numBins, int64_t* numHistograms,
                        double* binSize, double* inverseBinSize)
{
    MB_REQUIRE(h != nullptr, "hist_info");
    if (min) *min = h->min;
    if (max) *max = h->max;
    if (numBins) *numBins = h->numBins;
    if (numHistograms) *numHistograms = h->numHistograms;
    if (binSize) *binSize = h->binSize;
    if (inverseBinSize) *inverseBinSize = h->inverseBinSize;
    return 0;
}

void* mrmd_b200_hist_device_data(mrmd_b200_hist* h) { return h ? h->data : nullptr; }

int mrmd_b200_hist_write(mrmd_b200_hist* h, const double* src, int memKind, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(h != nullptr && (src != nullptr || entries(h) == 0), "hist_write");
    if (entries(h) == 0) return 0;
    MB_CUDA(cudaMemcpyAsync(h->data, src, size_t(entries(h)) * 8,
                            memKind == MRMD_B200_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, S(stream)));
    if (memKind == MRMD_B200_MEM_HOST) MB_CUDA(cudaStreamSynchronize(S(stream)));
    return 0;
}

int mrmd_b200_hist_read(const mrmd_b200_hist* h, double* dst, int memKind, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(h != nullptr && (dst != nullptr || entries(h) == 0), "hist_read");
    if (entries(h) == 0) return 0;
    MB_CUDA(cudaMemcpyAsync(dst, h->data, size_t(entries(h)) * 8,
                            memKind == MRMD_B200_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, S(stream)));
    if (memKind == MRMD_B200_MEM_HOST) MB_CUDA(cudaStreamSynchronize(S(stream)));
    return 0;
}

/* data/MultiHistogram.hpp:60-66 */
int64_t mrmd_b200_hist_get_bin(const mrmd_b200_hist* h, double val)
{
    if (h == nullptr) return -1;
    int64_t bin = static_cast<int64_t>(std::floor((val - h->min) * h->inverseBinSize));
    if (bin < 0) bin = -1;
    if (bin >= h->numBins) bin = -1;
    return bin;
}

/* :68-74 */
double mrmd_b200_hist_get_bin_position(const mrmd_b200_hist* h, int64_t binIdx)
{
    return h->min + (double(binIdx) + 0.5) * h->binSize;
}

int mrmd_b200_hist_transform(mrmd_b200_hist* h, const mrmd_b200_hist* rhs, int op, void* stream)
{
    MB_TRY(checkDevice());
    MB_TRY(sameShape(h, rhs));
    MB_REQUIRE(op >= 0 && op <= 3, "hist_transform: op is 0 (+=), 1 (-=), 2 (*=) or 3 (/=)");
    if (entries(h) == 0) return 0;
    histTransformKernel<<<gridFor(entries(h), MH_THREADS), MH_THREADS, 0, S(stream)>>>(dev(h), rhs->data, op);
    MB_LAUNCHED();
    return 0;
}

int mrmd_b200_hist_scale(mrmd_b200_hist* h, double factor, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(h != nullptr, "hist_scale");
    if (entries(h) == 0) return 0;
    histScaleKernel<<<gridFor(entries(h), MH_THREADS), MH_THREADS, 0, S(stream)>>>(dev(h), factor, nullptr);
    MB_LAUNCHED();
    return 0;
}

int mrmd_b200_hist_scale_per_histogram(mrmd_b200_hist* h, const double* factorsHost, int64_t numFactors, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(h != nullptr && factorsHost != nullptr, "hist_scale_per_histogram");
    MB_REQUIRE(numFactors >= h->numHistograms, "hist_scale_per_histogram: fewer factors than histograms");  // MultiHistogram.cpp:63
    if (entries(h) == 0) return 0;
    double* d = nullptr;
    MB_CUDA(cudaMalloc(&d, size_t(h->numHistograms) * 8));
    cudaError_t e = cudaMemcpyAsync(d, factorsHost, size_t(h->numHistograms) * 8, cudaMemcpyHostToDevice, S(stream));
    if (e == cudaSuccess)
    {
        histScaleKernel<<<gridFor(entries(h), MH_THREADS), MH_THREADS, 0, S(stream)>>>(dev(h), 1.0, d);
        g_launchCount.fetch_add(1);
        e = cudaStreamSynchronize(S(stream));
    }
    cudaFree(d);
    MB_CUDA(e);
    return 0;
}

int mrmd_b200_hist_make_symmetric(mrmd_b200_hist* h, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(h != nullptr, "hist_make_symmetric");
    const int64_t n = (h->numBins / 2) * h->numHistograms;
    if (n == 0) return 0;
    histMakeSymmetricKernel<<<gridFor(n, MH_THREADS), MH_THREADS, 0, S(stream)>>>(dev(h));
    MB_LAUNCHED();
    return 0;
}

int mrmd_b200_hist_cumulative_moving_average(mrmd_b200_hist* average, const mrmd_b200_hist* current, double movingAverageFactor,
                                             void* stream)
{
    MB_TRY(checkDevice());
    MB_TRY(sameShape(average, current));
    if (entries(average) == 0) return 0;
    histMovingAverageKernel<<<gridFor(entries(average), MH_THREADS), MH_THREADS, 0, S(stream)>>>(dev(average), current->data,
                                                                                                movingAverageFactor);
    MB_LAUNCHED();
    return 0;
}

int mrmd_b200_hist_gradient(mrmd_b200_hist** out, const mrmd_b200_hist* in, int periodic, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(in != nullptr, "hist_gradient");
    MB_TRY(histAlloc(out, in->min, in->max, in->numBins, in->numHistograms));
    if (entries(in) == 0) return 0;
    MB_REQUIRE(in->numBins >= 2, "hist_gradient: needs at least two bins");
    histGradientKernel<<<gridFor(entries(in), MH_THREADS), MH_THREADS, 0, S(stream)>>>(dev(in), (*out)->data, periodic);
    MB_LAUNCHED();
    return 0;
}

int mrmd_b200_hist_smoothen(mrmd_b200_hist** out, const mrmd_b200_hist* in, double sigma, double range, int periodic,
                            void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(in != nullptr && sigma > 0.0 && range >= 0.0, "hist_smoothen");
    MB_TRY(histAlloc(out, in->min, in->max, in->numBins, in->numHistograms));
    if (entries(in) == 0) return 0;
    histSmoothenKernel<<<gridFor(entries(in), MH_THREADS), MH_THREADS, 0, S(stream)>>>(dev(in), (*out)->data, sigma, range,
                                                                                      periodic);
    MB_LAUNCHED();
    return 0;
}

int mrmd_b200_hist_replace_if_bin_position(mrmd_b200_hist* h, const mrmd_b200_pred* pred, double newValue, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(h != nullptr && pred != nullptr, "hist_replace_if_bin_position");
    if (entries(h) == 0) return 0;
    histReplaceIfKernel<<<gridFor(entries(h), MH_THREADS), MH_THREADS, 0, S(stream)>>>(dev(h), *pred, newValue);
    MB_LAUNCHED();
    return 0;
}

int mrmd_b200_hist_create_grid(const mrmd_b200_hist* h, double* gridHost, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(h != nullptr && (gridHost != nullptr || h->numBins == 0), "hist_create_grid");
    if (h->numBins == 0) return 0;
    double* d = nullptr;
    MB_CUDA(cudaMalloc(&d, size_t(h->numBins) * 8));
    histGridKernel<<<gridFor(h->numBins, MH_THREADS), MH_THREADS, 0, S(stream)>>>(dev(h), d);
    g_launchCount.fetch_add(1);
    cudaError_t e = cudaMemcpyAsync(gridHost, d, size_t(h->numBins) * 8, cudaMemcpyDeviceToHost, S(stream));
    if (e == cudaSuccess) e = cudaStreamSynchronize(S(stream));
    cudaFree(d);
    MB_CUDA(e);
    return 0;
}

/* ThermodynamicForce::getForce() / getDensityProfile() (ThermodynamicForce.hpp:57-63): the table as a MultiHistogram
 * (a copy, like the reference's by-value return of a view-holding struct that callers then modify freely) */
int mrmd_b200_thermo_get_hist(const mrmd_b200_thermo* t, int kind, mrmd_b200_hist** out, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(t != nullptr && (kind == 0 || kind == 1), "thermo_get_hist: kind is 0 (force) or 1 (density profile)");
    MB_TRY(histAlloc(out, t->min, t->max, t->numBins, t->numTypes));
    MB_CUDA(cudaMemcpyAsync((*out)->data, kind == 0 ? t->force : t->density, size_t(t->numBins * t->numTypes) * 8,
                            cudaMemcpyDeviceToDevice, S(stream)));
    return 0;
}

}  // extern "C"
