// analysis.cu -- the diagnostics the reference's drivers print every output interval:
// kinetic energy, system momentum, pressure and mean square displacement as deterministic device reductions.
// Reference: mrmd/analysis/KineticEnergy.hpp:26-47, SystemMomentum.cpp:21-50, Pressure.cpp:23-51,
//            MeanSquareDisplacement.cpp:23-113 (call sites examples/02_LennardJones_NVE.cpp:190-199).
// Streaming reads (24-80 B per atom), HBM bound; each call returns host scalars after one stream sync, like the
// Kokkos::parallel_reduce it replaces.
#include <algorithm>
#include <mutex>

#include "handles.cuh"

struct mrmd_b200_msd
{
    mrmd_b200::DevBuf initialPos;  // double4 per item (MeanSquareDisplacement::initialPosition_)
    int64_t numItems = 0;
};

namespace mrmd_b200
{
constexpr int AN_THREADS = 256;

enum
{
    AN_KINETIC = 0,   // sum m v^2
    AN_MOMENTUM = 1,  // sum v (the reference sums velocities, not m v: SystemMomentum.cpp:30-45)
    AN_PRESSURE = 2,  // sum m v^2 + F . x
    AN_MSD = 3        // sum |dx|^2 with the reference's fold (MeanSquareDisplacement.cpp:70-79)
};

template <int MODE>
__global__ void __launch_bounds__(AN_THREADS)
    diagnosticsKernel(const double4* __restrict__ pos, AtomsView a, int64_t n, const double4* __restrict__ initialPos,
                      double Lx, double Ly, double Lz, double* partials, double* result, unsigned int* ticket)
{
    const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    if (idx < n)
    {
        if (MODE == AN_KINETIC || MODE == AN_PRESSURE)
        {
            const double vx = a.vel[0][idx], vy = a.vel[1][idx], vz = a.vel[2][idx];
            s0 = a.mass[idx] * (vx * vx + vy * vy + vz * vz);
            if (MODE == AN_PRESSURE)
            {
                const double4 p = ld4nc(pos + idx);
                s0 += a.force[0][idx] * p.x + a.force[1][idx] * p.y + a.force[2][idx] * p.z;  // util::dot3(F, x)
            }
        }
        else if (MODE == AN_MOMENTUM)
        {
            s0 = a.vel[0][idx];
            s1 = a.vel[1][idx];
            s2 = a.vel[2][idx];
        }
        else
        {
            const double4 p = ld4nc(pos + idx), q = ld4nc(initialPos + idx);
            double dx = fabs(q.x - p.x), dy = fabs(q.y - p.y), dz = fabs(q.z - p.z);
            if (dx > 0.5 * Lx) dx -= Lx;
            if (dy > 0.5 * Ly) dy -= Ly;
            if (dz > 0.5 * Lz) dz -= Lz;
            s0 = dx * dx + dy * dy + dz * dz;
        }
    }
    gridReduce3<AN_THREADS>(s0, s1, s2, partials, result, ticket);
}

// one scratch per process: the calls below return host values, so they are serialised anyway
struct DiagScratch
{
    std::mutex mutex;
    DevBuf partials;
    double* dResult = nullptr;
    unsigned int* dTicket = nullptr;
    double* hResult = nullptr;
};
static DiagScratch g_diag;

template <int MODE>
static int runDiagnostics(const double4* pos, const AtomsView& a, int64_t n, const double4* initialPos,
                          const double* L, double* out3, cudaStream_t st)
{
    std::lock_guard<std::mutex> lock(g_diag.mutex);
    if (g_diag.dResult == nullptr)
    {
        MB_CUDA(cudaMalloc(&g_diag.dResult, 48));
        MB_CUDA(cudaMalloc(&g_diag.dTicket, 4));
        MB_CUDA(cudaMallocHost(&g_diag.hResult, 48));
        MB_CUDA(cudaMemset(g_diag.dTicket, 0, 4));
    }
    out3[0] = out3[1] = out3[2] = 0.0;
    if (n <= 0) return 0;
    const int blocks = gridFor(n, AN_THREADS);
    MB_TRY(g_diag.partials.reserve(size_t(blocks) * 3 * 8));
    MB_CUDA(cudaMemsetAsync(g_diag.dResult, 0, 48, st));
    diagnosticsKernel<MODE><<<blocks, AN_THREADS, 0, st>>>(pos, a, n, initialPos, L ? L[0] : 0.0, L ? L[1] : 0.0,
                                                          L ? L[2] : 0.0, g_diag.partials.as<double>(), g_diag.dResult,
                                                          g_diag.dTicket);
    MB_LAUNCHED();
    MB_CUDA(cudaMemcpyAsync(g_diag.hResult, g_diag.dResult, 24, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < 3; ++k) out3[k] = g_diag.hResult[k];
    return 0;
}
}  // namespace mrmd_b200

using namespace mrmd_b200;

extern "C" {

int mrmd_b200_kinetic_energy(const mrmd_b200_atoms* a, double* kineticEnergy, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(a != nullptr && kineticEnergy != nullptr, "kinetic_energy");
    double r[3];
    MB_TRY(runDiagnostics<AN_KINETIC>(a->v.pos, a->v, a->numLocal, nullptr, nullptr, r, S(stream)));
    *kineticEnergy = 0.5 * r[0];  // KineticEnergy.hpp:38
    return 0;
}

int mrmd_b200_system_momentum(const mrmd_b200_atoms* a, double* momentum3, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(a != nullptr && momentum3 != nullptr, "system_momentum");
    return runDiagnostics<AN_MOMENTUM>(a->v.pos, a->v, a->numLocal, nullptr, nullptr, momentum3, S(stream));
}

int mrmd_b200_pressure(const mrmd_b200_atoms* a, const mrmd_b200_subdomain* s, double* pressure, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(a != nullptr && s != nullptr && pressure != nullptr, "pressure");
    double r[3];
    // local AND ghost atoms, Pressure.cpp:31
    MB_TRY(runDiagnostics<AN_PRESSURE>(a->v.pos, a->v, a->numLocal + a->numGhost, nullptr, nullptr, r, S(stream)));
    const double volume = s->diameter[0] * s->diameter[1] * s->diameter[2];
    *pressure = r[0] / (3.0 * volume);
    return 0;
}

int mrmd_b200_msd_create(mrmd_b200_msd** out)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(out != nullptr, "msd_create");
    *out = new mrmd_b200_msd;
    return 0;
}

int mrmd_b200_msd_destroy(mrmd_b200_msd* m)
{
    if (m == nullptr) return 0;
    cudaDeviceSynchronize();
    m->initialPos.release();
    delete m;
    return 0;
}

static int msdReset(mrmd_b200_msd* m, const double4* pos, int64_t n, cudaStream_t st)
{
    m->numItems = n;
    MB_TRY(m->initialPos.reserve(size_t(std::max<int64_t>(n, 1)) * 32));
    if (n > 0) MB_CUDA(cudaMemcpyAsync(m->initialPos.p, pos, size_t(n) * 32, cudaMemcpyDeviceToDevice, st));
    return 0;
}

static int msdCalc(const mrmd_b200_msd* m, const double4* pos, int64_t n, const mrmd_b200_subdomain* s, double* out,
                   cudaStream_t st)
{
    MB_REQUIRE(m->numItems == n, "msd_calc: the number of items changed since reset");  // MRMD_HOST_CHECK_EQUAL, :64
    double r[3];
    AtomsView none{};
    MB_TRY(runDiagnostics<AN_MSD>(pos, none, n, m->initialPos.as<double4>(), s->diameter, r, st));
    *out = (n > 0) ? r[0] / double(n) : 0.0;
    return 0;
}

int mrmd_b200_msd_reset_atoms(mrmd_b200_msd* m, const mrmd_b200_atoms* a, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(m != nullptr && a != nullptr, "msd_reset");
    return msdReset(m, a->v.pos, a->numLocal, S(stream));
}

int mrmd_b200_msd_reset_molecules(mrmd_b200_msd* m, const mrmd_b200_molecules* mol, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(m != nullptr && mol != nullptr, "msd_reset");
    return msdReset(m, mol->v.pos, mol->numLocal, S(stream));
}

int mrmd_b200_msd_calc_atoms(const mrmd_b200_msd* m, const mrmd_b200_atoms* a, const mrmd_b200_subdomain* s,
                             double* meanSquareDisplacement, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(m != nullptr && a != nullptr && s != nullptr && meanSquareDisplacement != nullptr, "msd_calc");
    return msdCalc(m, a->v.pos, a->numLocal, s, meanSquareDisplacement, S(stream));
}

int mrmd_b200_msd_calc_molecules(const mrmd_b200_msd* m, const mrmd_b200_molecules* mol, const mrmd_b200_subdomain* s,
                                 double* meanSquareDisplacement, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(m != nullptr && mol != nullptr && s != nullptr && meanSquareDisplacement != nullptr, "msd_calc");
    return msdCalc(m, mol->v.pos, mol->numLocal, s, meanSquareDisplacement, S(stream));
}

}  // extern "C"
