// common.cuh -- shared device helpers and host-side containers of the B200-native MRMD hot path.
//
// HBM layout (ours; the reference's Cabana AoSoA layout is an implementation detail no caller sees
// except through slices):
//   atoms      pos4[cap]   32-byte records {x, y, z, type-as-int64-bits}: one 256-bit load per gather
//              vel, force  three planes each (x[cap], y[cap], z[cap]), streamed coalesced
//              mass, charge, relMass   one plane each
//              gid         int64 plane: global atom id, travels with the record (permute, ghost copy, migration)
//   molecules  pos4[cap]   {X, Y, Z, unused}
//              w4[cap]     {lambda^mod, dlambda/dx, dlambda/dy, dlambda/dz}: everything the AdResS pair
//                          loop needs about the partner molecule in one 256-bit gather
//              oc[cap]     {atomsOffset, numAtoms} as one 16-byte record
//              lambda[cap], force planes
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/mrmd_b200.h"

namespace mrmd_b200
{
// ---------------------------------------------------------------------------------------------
// error handling
void setLastError(const std::string& msg);
int checkDevice();
extern std::atomic<int64_t> g_launchCount;

#define MB_CUDA(expr)                                                                              \
    do                                                                                             \
    {                                                                                              \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
        {                                                                                          \
            ::mrmd_b200::setLastError(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + \
                                      __FILE__ + ":" + std::to_string(__LINE__) + ")");           \
            return static_cast<int>(_e);                                                           \
        }                                                                                          \
    } while (0)

#define MB_TRY(expr)            \
    do                          \
    {                           \
        int _rc = (expr);       \
        if (_rc != 0) return _rc; \
    } while (0)

#define MB_REQUIRE(cond, msg)                                                                  \
    do                                                                                         \
    {                                                                                          \
        if (!(cond))                                                                           \
        {                                                                                      \
            ::mrmd_b200::setLastError(std::string("invalid argument: ") + msg + " [" #cond "]"); \
            return MRMD_B200_EINVAL;                                                           \
        }                                                                                      \
    } while (0)

// counts the launch and checks for a launch error
#define MB_LAUNCHED()                                  \
    do                                                 \
    {                                                  \
        ::mrmd_b200::g_launchCount.fetch_add(1);       \
        MB_CUDA(cudaGetLastError());                   \
    } while (0)

inline cudaStream_t S(void* stream) { return static_cast<cudaStream_t>(stream); }
inline int gridFor(int64_t n, int block) { return static_cast<int>((n + block - 1) / block); }

// grow-only device buffer
struct DevBuf
{
    void* p = nullptr;
    size_t bytes = 0;
    int reserve(size_t need)
    {
        if (need <= bytes) return 0;
        size_t want = need + need / 4 + 256;
        void* q = nullptr;
        MB_CUDA(cudaMalloc(&q, want));
        if (p) cudaFree(p);
        p = q;
        bytes = want;
        return 0;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <class T>
    T* as() const
    {
        return static_cast<T*>(p);
    }
};

// ---------------------------------------------------------------------------------------------
// device views passed by value to kernels
struct AtomsView
{
    double4* pos;  // x, y, z, type bits
    double* vel[3];
    double* force[3];
    double* mass;
    double* charge;
    double* relMass;
    long long* gid;  // global atom id: the Philox counter of the Langevin integrator (initialised to the index)
};

struct MolsView
{
    double4* pos;  // X, Y, Z, unused
    double4* w;    // lambda^mod, grad lambda
    longlong2* oc; // atomsOffset, numAtoms
    double* lambda;
    double* force[3];
};

struct SubdomainDev
{
    double minCorner[3];
    double maxCorner[3];
    double minInner[3];
    double maxInner[3];
    double diameter[3];
};
inline SubdomainDev toDev(const mrmd_b200_subdomain& s)
{
    SubdomainDev d;
    for (int i = 0; i < 3; ++i)
    {
        d.minCorner[i] = s.minCorner[i];
        d.maxCorner[i] = s.maxCorner[i];
        d.minInner[i] = s.minInnerCorner[i];
        d.maxInner[i] = s.maxInnerCorner[i];
        d.diameter[i] = s.diameter[i];
    }
    return d;
}

// Cabana CartesianGrid parameters (see neighbor.cu)
struct GridDev
{
    double min[3];
    double dx[3];
    double rdx[3];
    int n[3];
};

// action/LennardJones.hpp:27-39
struct LJType
{
    double ff1, ff2, ef1, ef2;
    double rcSqr;
    double cappingDistance, cappingDistanceSqr, cappingCoeff;
    double shift;
    double energyAtCappingPoint;
};
constexpr int MAX_LJ_TYPES = 4;  // type pairs kept in kernel parameter space: numTypes <= 4

struct LJTable
{
    LJType t[MAX_LJ_TYPES * MAX_LJ_TYPES];
};

// ---------------------------------------------------------------------------------------------
// host-side containers behind the opaque handles
}  // namespace mrmd_b200

struct mrmd_b200_atoms
{
    int64_t size = 0;
    int64_t capacity = 0;
    int64_t numLocal = 0;
    int64_t numGhost = 0;
    mrmd_b200::AtomsView v{};
    mrmd_b200::AtomsView alt{};  // second set of planes: permute target (ping-pong)
    int64_t altCapacity = 0;
    mrmd_b200::DevBuf staging;
    mrmd_b200::DevBuf sortScratch;
    double* dMaxDisp = nullptr;  // device scalar for the integrators
    double* hMaxDisp = nullptr;  // pinned
    // linked-cell structure left behind by the last cell sort (LinkedCellList + permute): atoms
    // [lcBegin, lcEnd) are ordered by cell of lcGrid; lcCellStart[c] is the absolute index of the first atom
    // of cell c (numCells + 1 entries).  Index ranges stay valid until the next sort / shrinking resize.
    bool lcValid = false;
    mrmd_b200::GridDev lcGrid{};
    int64_t lcBegin = 0, lcEnd = 0;
    int64_t lcNumCells = 0;
    int64_t lcEpoch = 0;
    // every operator that writes local positions bumps posEpoch; a cell sort records it.  The index ranges of the
    // sort stay usable while the atoms keep their order (lcValid), but a NEW tiled list may only be built from them
    // while the positions are the sorted ones (lcPosEpoch == posEpoch): drifted atoms sit in the wrong cells.
    int64_t posEpoch = 0;
    int64_t lcPosEpoch = -1;
    mrmd_b200::DevBuf lcCellStart;  // int32[numCells + 1]
};

struct mrmd_b200_molecules
{
    int64_t size = 0;
    int64_t capacity = 0;
    int64_t numLocal = 0;
    int64_t numGhost = 0;
    mrmd_b200::MolsView v{};
    mrmd_b200::MolsView alt{};
    int64_t altCapacity = 0;
    mrmd_b200::DevBuf staging;
    mrmd_b200::DevBuf sortScratch;
    // Linked-cell structure of the molecules' centres of mass left behind by moleculesCellSortWithAtoms, kept in an
    // atoms-shaped view (pos = the molecules' pos plane, numLocal = local molecules, no other plane) so that the tiled
    // neighbour build runs on molecules unchanged.  Created on first use, owned by the molecules handle.
    mrmd_b200_atoms* lcView = nullptr;
};

struct mrmd_b200_verlet
{
    int half = 1;
    int64_t numParticles = 0;  // rows allocated (== size() of the position slice, as in Cabana)
    int64_t begin = 0, end = 0;
    int64_t width = 0;  // slots per row
    int64_t pitch = 0;  // particles per slot row (>= numParticles, multiple of 32)
    int64_t buildCount = 0;
    mrmd_b200::DevBuf counts;   // int32[pitch]
    mrmd_b200::DevBuf neigh;    // int32[width * pitch]
    // tiled flavour (tiled.cu): neighbours as 16-bit shared-memory slots of the owner's tile
    bool tiled = false;
    mrmd_b200::DevBuf enc;       // uint16[numParticles][width], rows in slot order
    mrmd_b200::DevBuf tileDesc;  // int[tiles][64], see tiled.cu
    mrmd_b200::DevBuf cellLoHi;  // int32[2][extended cells]: first / one-past-last atom of every cell
    int tiledHaloX = 0;          // x-slab decomposition: halo columns at i = -1 and i = nx
    mrmd_b200_subdomain tiledSub{};
    int64_t tiledEpoch = -1;
    int tiledR = 1;  // cells along z spanned by the list radius
    // AdResS step-loop drivers: tiles that lie entirely in the coarse-grained region of this (slab) weighting function
    // get empty rows (no pair of theirs is ever evaluated); never set through the C ABI list builders
    bool tiledCgSkip = false;
    mrmd_b200_weight tiledCgWeight{};
    int tiledCH = 0;
    int tiledSlots = 0;
    int tiledGridN[3] = {0, 0, 0};  // grid the tile geometry (CH, slots) was chosen for: re-used while it stays the same
    // tile sizing: home units per tile aimed at, and the shared-memory bytes per staged slot of the largest consumer
    // (24 for atoms; molecule lists stage all atoms of a molecule per slot in the force kernel)
    int tiledTargetHomes = 110;
    int tiledSlotBytes = 24;
    int tiledSmemBudget = 48 * 1024;  // preferred shared memory of a tile's staged slots (several tiles per SM)
    mrmd_b200::DevBuf tstats;    // int32[4]: max slots per tile, -, overflow flag, tiles with work
    mrmd_b200::DevBuf tileActive;   // uint8[tiles]: 0 for the tiles the build left empty (coarse-grained bulk)
    mrmd_b200::DevBuf activeTiles;  // int32[numActiveTiles], ascending
    int numActiveTiles = 0;
    int* hTstats = nullptr;      // pinned
    mrmd_b200::DevBuf keys[2];  // radix sort ping-pong
    mrmd_b200::DevBuf vals[2];
    mrmd_b200::DevBuf scratch;
    mrmd_b200::DevBuf cellStart;  // int32[numCells + 1]
    mrmd_b200::DevBuf sortedPos;  // double4[n]: x, y, z, original index bits
    mrmd_b200::DevBuf stats;      // int32[4]: max count, ...; int64 total
    int* hStats = nullptr;        // pinned
};

namespace mrmd_b200
{
int atomsEnsureCapacity(mrmd_b200_atoms* a, int64_t capacity, cudaStream_t st);
int atomsEnsureAlt(mrmd_b200_atoms* a, cudaStream_t st);
int molsEnsureCapacity(mrmd_b200_molecules* m, int64_t capacity, cudaStream_t st);
int molsEnsureAlt(mrmd_b200_molecules* m, cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// device helpers
#ifdef __CUDACC__

// 256-bit global loads/stores (LDG.E.ENL2.256 / STG.E.ENL2.256 on sm_100a)
__device__ __forceinline__ double4 ld4(const double4* p)
{
    double4 v;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}
// read-only path; only for data that is not written by the same kernel
__device__ __forceinline__ double4 ld4nc(const double4* p)
{
    double4 v;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st4(double4* p, const double4& v)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ int64_t typeOf(const double4& p) { return __double_as_longlong(p.w); }

// util/IsInSymmetricSlab.hpp:34-50
__device__ __forceinline__ bool slab1(const mrmd_b200_pred& p, double x, double y, double z)
{
    const double c = (p.axis == 0) ? x : ((p.axis == 1) ? y : z);
    const double absDx = fabs(c - p.center);
    return (absDx >= p.slabMin - p.tolerance && absDx <= p.slabMax + p.tolerance);
}
__device__ __forceinline__ bool pred1(const mrmd_b200_pred& p, double x, double y, double z)
{
    switch (p.kind)
    {
        case MRMD_B200_PRED_ALWAYS: return true;
        case MRMD_B200_PRED_NEVER: return false;
        case MRMD_B200_PRED_INTERVAL:
        {
            const double c = (p.axis == 0) ? x : ((p.axis == 1) ? y : z);
            return c > p.slabMin && c < p.slabMax;
        }
        default: return slab1(p, x, y, z);
    }
}
// examples/04_LennardJones_IdealGas_LocalCap.cpp:206-227
__device__ __forceinline__ bool pred2(const mrmd_b200_pred& p, double x1, double y1, double z1, double x2,
                                      double y2, double z2)
{
    switch (p.kind)
    {
        case MRMD_B200_PRED_ALWAYS: return true;
        case MRMD_B200_PRED_NEVER: return false;
        case MRMD_B200_PRED_SLAB_EITHER: return slab1(p, x1, y1, z1) || slab1(p, x2, y2, z2);
        case MRMD_B200_PRED_SLAB_BOTH: return slab1(p, x1, y1, z1) && slab1(p, x2, y2, z2);
        default: return slab1(p, x1, y1, z1);
    }
}

// action/LennardJones.hpp:53-78
__device__ __forceinline__ void ljForceEnergy(const LJType& t, double distSqr, double& ff, double& e)
{
    if (distSqr >= t.cappingDistanceSqr)
    {
        const double frac2 = 1.0 / distSqr;
        const double frac6 = frac2 * frac2 * frac2;
        ff = frac6 * (t.ff1 * frac6 - t.ff2) * frac2;
        e = frac6 * (t.ef1 * frac6 - t.ef2) - t.shift;
        return;
    }
    const double dist = sqrt(distSqr);
    ff = t.cappingCoeff / dist;
    e = t.energyAtCappingPoint - (dist - t.cappingDistance) * t.cappingCoeff - t.shift;
}

// squared distance in the reference's left-to-right order WITHOUT fused multiply-add, so the cutoff
// decisions agree bit for bit with an uncontracted CPU build ("FP contraction")
__device__ __forceinline__ double distSqrExact(double dx, double dy, double dz)
{
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

__device__ __forceinline__ double warpSum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warpMax(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ long long warpSumLL(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Grid-wide sums of per-thread counts (integers or halves, far below 2^52 in total): every partial sum is exact in
// double precision, so the order in which the blocks' fire-and-forget REDs land does not matter -- deterministic
// without the fence / ticket / last-block pass of gridReduce3 (which kept every block of the tiled force kernels
// resident for a device-wide memory fence: 11 % of the stall samples of ljForceTiledKernel).  slotA / slotB (either
// may be nullptr) name the per-launch slots of gridReduce3's result layout; only the RUNNING sums three doubles behind
// them are updated: the step-loop drivers read nothing else between the launches that ask for energies, and a per-launch
// value would need the slot zeroed in front of every launch (one more node in the stream per step).
template <int THREADS>
__device__ __forceinline__ void gridAddExact(double a, double b, double* slotA, double* slotB)
{
    __shared__ double sAdd[2][THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a = warpSum(a);
    if (slotB != nullptr) b = warpSum(b);
    if (lane == 0)
    {
        sAdd[0][warp] = a;
        sAdd[1][warp] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        double sa = 0, sb = 0;
        for (int w = 0; w < THREADS / 32; ++w)
        {
            sa += sAdd[0][w];
            sb += sAdd[1][w];
        }
        if (slotA != nullptr && sa != 0.0) atomicAdd(slotA + 3, sa);
        if (slotB != nullptr && sb != 0.0) atomicAdd(slotB + 3, sb);
    }
}

// Deterministic grid-wide sum of three per-thread scalars: warp shuffle -> block -> per-block partial;
// the last block to arrive (ticket counter) adds the partials in a fixed order and resets the ticket.
// partials holds 3 * gridDim.x doubles; result 6 doubles: [0..2] this launch, [3..5] running sums.
// numBlocks (0: the whole grid): only blocks [0, numBlocks) take part -- the others must have returned before the
// call, and with numBlocks == 0 participants nothing is written (the caller zeroes result[0..2] before the launch).
template <int THREADS>
__device__ __forceinline__ void gridReduce3(double e, double v, double c, double* partials, double* result,
                                            unsigned int* ticket, unsigned int numBlocks = 0)
{
    if (numBlocks == 0) numBlocks = gridDim.x;
    __shared__ double sRed[3][THREADS / 32];
    __shared__ bool sLast;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    e = warpSum(e);
    v = warpSum(v);
    c = warpSum(c);
    if (lane == 0)
    {
        sRed[0][warp] = e;
        sRed[1][warp] = v;
        sRed[2][warp] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        double se = 0, sv = 0, sc = 0;
        for (int w = 0; w < THREADS / 32; ++w)
        {
            se += sRed[0][w];
            sv += sRed[1][w];
            sc += sRed[2][w];
        }
        partials[blockIdx.x] = se;
        partials[gridDim.x + blockIdx.x] = sv;
        partials[2 * gridDim.x + blockIdx.x] = sc;
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        sLast = (t == numBlocks - 1);
    }
    __syncthreads();
    if (!sLast) return;
    __threadfence();
    double acc[3] = {0, 0, 0};
    for (unsigned int b = threadIdx.x; b < numBlocks; b += THREADS)
    {
        acc[0] += __ldcg(partials + b);
        acc[1] += __ldcg(partials + gridDim.x + b);
        acc[2] += __ldcg(partials + 2 * gridDim.x + b);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 3; ++k) acc[k] = warpSum(acc[k]);
    if (lane == 0)
        for (int k = 0; k < 3; ++k) sRed[k][warp] = acc[k];
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int k = 0; k < 3; ++k)
        {
            double s = 0;
            for (int w = 0; w < THREADS / 32; ++w) s += sRed[k][w];
            result[k] = s;
            result[3 + k] += s;  // running sums over launches (pair-interactions/s numerator)
        }
        *ticket = 0;
    }
}

// data/MultiHistogram.hpp:60-66 (multiply by the reciprocal bin size, no FMA)
__device__ __forceinline__ long long histBin(double min, double inverseBinSize, long long numBins, double val)
{
    long long bin = static_cast<long long>(floor(__dmul_rn(__dsub_rn(val, min), inverseBinSize)));
    if (bin < 0) bin = -1;
    if (bin >= numBins) bin = -1;
    return bin;
}

// weighting_function/CheckRegion.hpp:25-38 (REGION_CHECK_EPSILON = 0)
__device__ __forceinline__ bool inAT(double l) { return l >= 1.0; }
__device__ __forceinline__ bool inCG(double l) { return l <= 0.0; }
__device__ __forceinline__ bool inHY(double l) { return !inAT(l) && !inCG(l); }

// Cabana CartesianGrid::locatePoint per dimension: floor((x - min) * rdx), == n -> n-1, then a
// memory-safety clamp (out-of-grid points are undefined in Cabana).  No FMA.
__device__ __forceinline__ int locate1(const GridDev& g, double x, int d)
{
    int c = static_cast<int>(floor(__dmul_rn(__dsub_rn(x, g.min[d]), g.rdx[d])));
    c = (c == g.n[d]) ? c - 1 : c;
    c = max(0, min(c, g.n[d] - 1));
    return c;
}
__device__ __forceinline__ int cardinal(const GridDev& g, int i, int j, int k) { return (i * g.n[1] + j) * g.n[2] + k; }

#endif  // __CUDACC__

// host helpers shared between translation units
GridDev makeGrid(const double* gmin, const double* gmax, const double* delta);
int radixSortPairs(uint32_t* keysIn, uint32_t* valsIn, uint32_t* keysTmp, uint32_t* valsTmp, uint32_t* scratch,
                   int64_t n, int keyBits, uint32_t** keysOut, uint32_t** valsOut, cudaStream_t st);
size_t radixSortScratchBytes(int64_t n);

}  // namespace mrmd_b200
