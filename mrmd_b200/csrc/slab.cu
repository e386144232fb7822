// slab.cu -- one-GPU-per-process x-slab decomposition of the LJ step loop.
//
// The reference has no inter-rank communication (mrmd/communication is single-process periodic self-ghosting,
// MultiResRealAtomsExchange.hpp:26); this is the multi-GPU extension the north star asks for.  Each rank owns
// the slab [xlo, xhi) of the global box.  y / z stay locally periodic (the tiled kernels generate those images
// on the fly), the x pass of GhostExchange becomes an NCCL halo:
//   rebuild   atoms that left the slab migrate to the neighbour rank (full records), the rank sorts its atoms
//             by linked cell, the atoms within rc+skin of each x face are sent as halo atoms (they arrive in
//             (j,k) cell order because the sender's atoms are cell sorted) and the tiled neighbour build treats
//             them as two extra cell columns;
//   step      positions of the halo atoms only (32 B per atom), stored by haloPushKernel straight into the
//             neighbours' IPC-mapped buffers over NVLink and collected by haloPullKernel (ncclSend/ncclRecv as the
//             fallback); no reverse force halo: with the full list every rank computes the complete force of its
//             own atoms;
//   decision  the displacement criterion (examples/02:141-143) uses the maximum over the ranks, gathered through
//             the same peer buffers by maxDisplacementGatherKernel (ncclAllReduce(max) as the fallback), so all
//             ranks rebuild together.
// The periodic wrap in x is applied by the two end ranks when they send across the global boundary.
// NCCL is resolved at run time with dlopen (single-GPU users do not need it); it bootstraps the peer mapping and
// carries the rebuild-time traffic (migration records, counts) and the read-out reductions.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <ctime>
#include <cfloat>
#include <cmath>
#include <vector>

#include "handles.cuh"
#include "hostpipe.cuh"

namespace mrmd_b200
{
struct NcclApi
{
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) getUniqueId = nullptr;
    decltype(&ncclCommInitRank) commInitRank = nullptr;
    decltype(&ncclCommDestroy) commDestroy = nullptr;
    decltype(&ncclSend) send = nullptr;
    decltype(&ncclRecv) recv = nullptr;
    decltype(&ncclGroupStart) groupStart = nullptr;
    decltype(&ncclGroupEnd) groupEnd = nullptr;
    decltype(&ncclAllReduce) allReduce = nullptr;
    decltype(&ncclAllGather) allGather = nullptr;
    decltype(&ncclGetErrorString) getErrorString = nullptr;
};

static NcclApi g_nccl;

static int loadNccl()
{
    if (g_nccl.lib != nullptr) return 0;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (lib == nullptr) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (lib == nullptr)
    {
        setLastError(std::string("cannot load NCCL: ") + dlerror());
        return MRMD_B200_EINVAL;
    }
#define NCCL_SYM(field, name)                                                  \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(lib, name)); \
    if (g_nccl.field == nullptr)                                               \
    {                                                                          \
        setLastError(std::string("NCCL symbol missing: ") + name);             \
        return MRMD_B200_EINVAL;                                               \
    }
    NCCL_SYM(getUniqueId, "ncclGetUniqueId");
    NCCL_SYM(commInitRank, "ncclCommInitRank");
    NCCL_SYM(commDestroy, "ncclCommDestroy");
    NCCL_SYM(send, "ncclSend");
    NCCL_SYM(recv, "ncclRecv");
    NCCL_SYM(groupStart, "ncclGroupStart");
    NCCL_SYM(groupEnd, "ncclGroupEnd");
    NCCL_SYM(allReduce, "ncclAllReduce");
    NCCL_SYM(allGather, "ncclAllGather");
    NCCL_SYM(getErrorString, "ncclGetErrorString");
#undef NCCL_SYM
    g_nccl.lib = lib;
    return 0;
}

#define MB_NCCL(expr)                                                                                       \
    do                                                                                                      \
    {                                                                                                       \
        ncclResult_t _r = (expr);                                                                           \
        if (_r != ncclSuccess)                                                                              \
        {                                                                                                   \
            ::mrmd_b200::setLastError(std::string(#expr) + ": " + g_nccl.getErrorString(_r));               \
            return MRMD_B200_EINVAL;                                                                        \
        }                                                                                                   \
    } while (0)

constexpr int SL_THREADS = 256;
constexpr int SL_RECORD = 14;  // doubles per migrating atom: pos3, type bits, vel3, force3, mass, charge, relMass, id bits

// y / z periodic wrap (PeriodicMapping.cpp:36-51 arithmetic) and the x migration flag: -1 left, +1 right, 0 stays
__global__ void slabWrapFlagKernel(double4* pos, int64_t n, SubdomainDev s, signed char* flag)
{
    const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (idx >= n) return;
    double4 p = ld4(pos + idx);
    double* x = &p.x;
#pragma unroll
    for (int dim = 1; dim < 3; ++dim)
    {
        if (s.maxCorner[dim] <= x[dim])
        {
            x[dim] -= s.diameter[dim];
            x[dim] = fmax(x[dim], s.minCorner[dim]);
        }
        if (x[dim] < s.minCorner[dim])
        {
            x[dim] += s.diameter[dim];
            if (s.maxCorner[dim] <= x[dim]) x[dim] = s.minCorner[dim];
        }
    }
    st4(pos + idx, p);
    flag[idx] = (p.x < s.minCorner[0]) ? -1 : ((p.x >= s.maxCorner[0]) ? 1 : 0);
}

// the same for molecules of apm consecutive atoms, positioned by their centre of mass: the y / z wrap moves the molecule
// with all its atoms by +-L without a clamp (MultiResRealAtomsExchange.cpp:23-73), the x flag follows the centre of mass
__global__ void slabWrapFlagMolKernel(double4* com, double4* pos, int64_t nm, int apm, SubdomainDev s, signed char* flag)
{
    const int64_t mi = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (mi >= nm) return;
    double4 c = ld4(com + mi);
    double* cx = &c.x;
    double shift[3] = {0.0, 0.0, 0.0};
    bool moved = false;
#pragma unroll
    for (int dim = 1; dim < 3; ++dim)
    {
        if (s.maxCorner[dim] <= cx[dim])
        {
            cx[dim] -= s.diameter[dim];
            shift[dim] -= 1.0;
            moved = true;
        }
        if (cx[dim] < s.minCorner[dim])
        {
            cx[dim] += s.diameter[dim];
            shift[dim] += 1.0;
            moved = true;
        }
    }
    flag[mi] = (c.x < s.minCorner[0]) ? -1 : ((c.x >= s.maxCorner[0]) ? 1 : 0);
    if (!moved) return;
    st4(com + mi, c);
    for (int j = 0; j < apm; ++j)
    {
        double4 p = ld4(pos + mi * apm + j);
        double* x = &p.x;
#pragma unroll
        for (int dim = 1; dim < 3; ++dim)
        {
            if (shift[dim] < 0.0) x[dim] -= s.diameter[dim];
            if (shift[dim] > 0.0) x[dim] += s.diameter[dim];
        }
        st4(pos + mi * apm + j, p);
    }
}

__global__ void slabMoleculeInitKernel(MolsView m, int64_t n, int64_t apm)
{
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i < n) m.oc[i] = make_longlong2(i * apm, apm);
}

// stable selection of two subsets of [first, first + n): block counts -> scan -> ranked index lists
template <int MODE>  // 0: migration flags, 1: halo faces by position
__device__ __forceinline__ void selectPredicates(const double4* pos, const signed char* flag, int64_t idx, double lowBound,
                                                 double highBound, bool& lo, bool& hi)
{
    if (MODE == 0)
    {
        const signed char f = flag[idx];
        lo = f < 0;
        hi = f > 0;
    }
    else
    {
        const double x = ld4nc(pos + idx).x;
        lo = x < lowBound;    // GhostExchange.cpp:80 (x < minInnerCorner)
        hi = x >= highBound;  // :89 (x >= maxInnerCorner)
    }
}

__device__ __forceinline__ void blockScan2(long long (&v)[2], long long (&total)[2])
{
    __shared__ long long sWarp[2][SL_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long incl[2];
#pragma unroll
    for (int k = 0; k < 2; ++k)
    {
        long long x = v[k];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        incl[k] = x;
        if (lane == 31) sWarp[k][warp] = x;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 2; ++k)
    {
        long long off = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < SL_THREADS / 32; ++w)
        {
            const long long c = sWarp[k][w];
            if (w < warp) off += c;
            tot += c;
        }
        v[k] = off + incl[k] - v[k];
        total[k] = tot;
    }
    __syncthreads();
}

template <int MODE>
__global__ void __launch_bounds__(SL_THREADS)
    selectCountKernel(const double4* pos, const signed char* flag, int64_t first, int64_t n, double lowBound,
                      double highBound, int64_t* blockCounts)
{
    const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    long long v[2] = {0, 0}, total[2];
    if (j < n)
    {
        bool lo, hi;
        selectPredicates<MODE>(pos, flag, first + j, lowBound, highBound, lo, hi);
        v[0] = lo;
        v[1] = hi;
    }
    blockScan2(v, total);
    if (threadIdx.x == 0)
    {
        blockCounts[2 * blockIdx.x] = total[0];
        blockCounts[2 * blockIdx.x + 1] = total[1];
    }
}

__global__ void __launch_bounds__(SL_THREADS) selectScanKernel(int64_t* blockCounts, int64_t numBlocks, int64_t* totals)
{
    __shared__ long long sCarry[2];
    if (threadIdx.x < 2) sCarry[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t base = 0; base < numBlocks; base += SL_THREADS)
    {
        const int64_t b = base + threadIdx.x;
        long long v[2], total[2];
        v[0] = (b < numBlocks) ? blockCounts[2 * b] : 0;
        v[1] = (b < numBlocks) ? blockCounts[2 * b + 1] : 0;
        blockScan2(v, total);
        if (b < numBlocks)
        {
            blockCounts[2 * b] = v[0] + sCarry[0];
            blockCounts[2 * b + 1] = v[1] + sCarry[1];
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            sCarry[0] += total[0];
            sCarry[1] += total[1];
        }
        __syncthreads();
    }
    if (threadIdx.x < 2) totals[threadIdx.x] = sCarry[threadIdx.x];
}

template <int MODE>
__global__ void __launch_bounds__(SL_THREADS)
    selectIndexKernel(const double4* pos, const signed char* flag, int64_t first, int64_t n, double lowBound,
                      double highBound, const int64_t* blockOffsets, int32_t* lowIdx, int32_t* highIdx)
{
    const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    long long v[2] = {0, 0}, total[2];
    bool lo = false, hi = false;
    if (j < n)
    {
        selectPredicates<MODE>(pos, flag, first + j, lowBound, highBound, lo, hi);
        v[0] = lo;
        v[1] = hi;
    }
    blockScan2(v, total);
    if (lo) lowIdx[blockOffsets[2 * blockIdx.x] + v[0]] = static_cast<int32_t>(first + j);
    if (hi) highIdx[blockOffsets[2 * blockIdx.x + 1] + v[1]] = static_cast<int32_t>(first + j);
}

// full records of the migrating atoms, x shifted by `shift` (periodic wrap at the global ends)
__global__ void packRecordsKernel(AtomsView a, const int32_t* idx, int64_t n, double shift, double* buf)
{
    const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (k >= n) return;
    const int64_t i = idx[k];
    const double4 p = ld4(a.pos + i);
    double* r = buf + k * SL_RECORD;
    r[0] = p.x + shift;
    r[1] = p.y;
    r[2] = p.z;
    r[3] = p.w;
    for (int d = 0; d < 3; ++d)
    {
        r[4 + d] = a.vel[d][i];
        r[7 + d] = a.force[d][i];
    }
    r[10] = a.mass[i];
    r[11] = a.charge[i];
    r[12] = a.relMass[i];
    r[13] = __longlong_as_double(a.gid[i]);
}

__global__ void unpackRecordsKernel(AtomsView a, int64_t first, int64_t n, const double* buf, signed char* flag)
{
    const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (k >= n) return;
    const int64_t i = first + k;
    const double* r = buf + k * SL_RECORD;
    st4(a.pos + i, make_double4(r[0], r[1], r[2], r[3]));
    for (int d = 0; d < 3; ++d)
    {
        a.vel[d][i] = r[4 + d];
        a.force[d][i] = r[7 + d];
    }
    a.mass[i] = r[10];
    a.charge[i] = r[11];
    a.relMass[i] = r[12];
    a.gid[i] = __double_as_longlong(r[13]);
    flag[i] = 0;
}

// positions (+ type bits) of the halo atoms, x shifted by `shift`
__global__ void packPositionsKernel(const double4* pos, const int32_t* idx, int64_t n, double shift, double4* buf)
{
    const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (k >= n) return;
    double4 p = ld4(pos + idx[k]);
    p.x += shift;
    st4(buf + k, p);
}

// ---- per-step position halo over peer memory ------------------------------------------------------------------
// Every rank owns one IPC-shared buffer {flags, region for the left neighbour's atoms, region for the right
// neighbour's}.  haloPushKernel packs this rank's face atoms (image shift applied) with plain stores straight into the
// two neighbours' buffers over NVLink and then publishes a sequence number next to them; haloPullKernel on the
// receiving rank waits for both numbers and copies the records behind its local atoms.  No NCCL call, no staging
// buffer, no host involvement in the per-step exchange.  A region is only rewritten after the rank-wide allreduce of
// the next step, which the reader's stream reaches only after its pull (and force kernel) have finished.
constexpr int SL_MAX_PEERS = 8;    // ranks of one node
// Header of a peer buffer, in 8-byte words (all written by OTHER ranks over NVLink, read by the owner):
//   [0..3]   halo flags {fromLeft, fromRight} for the two parities of the per-step halo (sequence numbers)
//   [4..35]  displacement all-gather: 2 parities x SL_MAX_PEERS x {value, sequence number}
//   [36..39] migration at a rebuild: {count fromLeft, seq, count fromRight, seq}
//   [40..43] halo lists at a rebuild: {count fromLeft, seq, count fromRight, seq}
// behind it (in double4 units): four halo regions (parity x side, p2pCap unit records each), then two migration regions
constexpr int SL_HW_HALO = 0, SL_HW_DECIDE = 4, SL_HW_MIG = 36, SL_HW_HALOCOUNT = 40;
constexpr int SL_P2P_HEADER = 12;  // double4 slots (48 words)
struct PeerBuffers
{
    double4* p[SL_MAX_PEERS];
};

// The rebuild decision needs the maximum displacement over all ranks every step.  Instead of an NCCL all-reduce (tens
// of microseconds of launch and protocol latency for 8 bytes) every rank stores {value, step} into its slot of every
// peer's buffer over NVLink, waits until its own buffer holds this step's value of every rank, and takes the maximum:
// one single-warp kernel per step.  The result also goes to pinned host memory (zero copy) with the step number
// behind it, so the host learns it without a stream synchronisation.  Slots alternate between two sets by step parity:
// a rank can be at most one decision ahead of the slowest one (it needs that rank's value to finish its own).
// Queued steps (slabRunQueued): accum / stop are given, the kernel also evaluates the criterion of examples/02:138-143 on
// the global maximum -- every rank accumulates the same values and stops at the same step -- and returns at once when an
// earlier queued step already stopped.
__global__ void maxDisplacementGatherKernel(const double* myMaxSqr, PeerBuffers peers, int rank, int nranks, double seq,
                                            int parity, double* outDevice, volatile double* outHost, double* accum,
                                            double threshold, int* stop, int localStep)
{
    if (stop != nullptr && stop[0] != 0) return;
    const int t = threadIdx.x;
    const int slotBase = SL_HW_DECIDE + parity * 2 * SL_MAX_PEERS;  // in 8-byte words from the start of the buffer
    if (t < nranks)
    {
        volatile double* slot = reinterpret_cast<double*>(peers.p[t]) + slotBase + 2 * rank;
        slot[0] = *myMaxSqr;
        __threadfence_system();
        slot[1] = seq;
    }
    double m = 0.0;
    if (t < nranks)
    {
        const volatile double* mine = reinterpret_cast<double*>(peers.p[rank]) + slotBase + 2 * t;
        while (mine[1] < seq) {}
        __threadfence_system();
        m = mine[0];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (t == 0)
    {
        *outDevice = m;
        outHost[0] = m;
        __threadfence_system();
        outHost[1] = seq;
        if (accum != nullptr)
        {
            if (!(m == m) || m > 1.7e308)
            {
                stop[0] = 1;
                stop[1] = localStep;
                stop[2] = 1;
            }
            else
            {
                const double sum = *accum + sqrt(m);
                *accum = sum;
                if (sum >= threshold)
                {
                    stop[0] = 1;
                    stop[1] = localStep;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(SL_THREADS)
    haloPushKernel(const double4* __restrict__ pos, const int32_t* __restrict__ idxLow, int64_t nl, double shiftL,
                   double4* dstL, unsigned long long* flagL, const int32_t* __restrict__ idxHigh, int64_t nh, double shiftR,
                   double4* dstR, unsigned long long* flagR, unsigned long long seq, unsigned int* ticket)
{
    const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (k < nl)
    {
        double4 p = ld4(pos + idxLow[k]);
        p.x += shiftL;
        st4(dstL + k, p);
    }
    else if (k < nl + nh)
    {
        double4 p = ld4(pos + idxHigh[k - nl]);
        p.x += shiftR;
        st4(dstR + (k - nl), p);
    }
    // the last block to finish publishes: every block's records are visible system wide before the flags are
    __threadfence_system();
    __shared__ bool sLast;
    __syncthreads();
    if (threadIdx.x == 0) sLast = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (sLast && threadIdx.x == 0)
    {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(flagL) = seq;
        *reinterpret_cast<volatile unsigned long long*>(flagR) = seq;
        *ticket = 0;
    }
}

__global__ void __launch_bounds__(SL_THREADS)
    haloPullKernel(const double4* fromLeft, int64_t nLeft, const double4* fromRight, int64_t nRight,
                   const unsigned long long* flags, unsigned long long seq, double4* dst)
{
    if (threadIdx.x == 0)
    {
        const volatile unsigned long long* f = flags;
        while (f[0] < seq) {}
        while (f[1] < seq) {}
    }
    __syncthreads();
    __threadfence_system();
    const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    // the records were written by another GPU: read them past L1
    if (k < nLeft + nRight)
    {
        const double2* q = reinterpret_cast<const double2*>((k < nLeft) ? fromLeft + k : fromRight + (k - nLeft));
        const double2 lo = __ldcv(q), hi = __ldcv(q + 1);
        st4(dst + k, make_double4(lo.x, lo.y, hi.x, hi.y));
    }
}

// ---- rebuild-time exchange over peer memory ---------------------------------------------------------------------
// Everything a rebuild sends (migrating atoms, the lists of face atoms) travels through the same IPC-mapped peer
// buffers as the per-step halo, with the counts published next to the sequence flags, so that a rebuild costs two host
// round trips (after the migration, to learn the new number of local atoms; after the neighbour build, for the list
// statistics) instead of three NCCL rounds and eight stream synchronisations.  The unit of selection is a block of
// `apm` consecutive atoms positioned by upos[] (the atom itself, or the centre of mass of a molecule).
//
// Two stable selections A and B of units, by one launch each of count / scan / index (blockIdx.y = list):
//   MODE 0  migration: both over [0, n): A = flag < 0 (leaves to the left), B = flag > 0
//   MODE 1  halo faces: A = units of the first cell column with x < lowBound, B = units of the last cell column with
//           x >= highBound; the column ranges are read from the device-side cell prefix array (no host round trip)
//   MODE 2  migration by position over the same two columns of the PREVIOUS sort (bounds = the slab corners): a unit
//           moves by less than the skin between two rebuilds, so a leaver sat in a boundary column
struct SelArgs
{
    const double4* upos;
    const signed char* flag;
    const int32_t* rangeA0;  // MODE 1: first / one-past-last unit of list A and B (device pointers)
    const int32_t* rangeA1;
    const int32_t* rangeB0;
    const int32_t* rangeB1;
    int64_t n;               // MODE 0: units
    double lowBound, highBound;
};

template <int MODE>
__device__ __forceinline__ bool selPredicate(const SelArgs& g, int list, int64_t j, int64_t& unit)
{
    if (MODE == 0)
    {
        unit = j;
        if (j >= g.n) return false;
        const signed char f = g.flag[j];
        return list == 0 ? (f < 0) : (f > 0);
    }
    const int64_t first = list == 0 ? *g.rangeA0 : *g.rangeB0, last = list == 0 ? *g.rangeA1 : *g.rangeB1;
    unit = first + j;
    if (unit >= last) return false;
    const double x = ld4nc(g.upos + unit).x;
    // MODE 1: GhostExchange.cpp:80 / :89 (x < minInnerCorner, x >= maxInnerCorner); MODE 2: the units that left the slab
    // (x < minCorner, x >= maxCorner), looked for in the boundary columns of the previous sort only
    return list == 0 ? (x < g.lowBound) : (x >= g.highBound);
}

__device__ __forceinline__ long long blockScan1(long long v, long long& total)
{
    __shared__ long long sWarp1[SL_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const long long y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) sWarp1[warp] = x;
    __syncthreads();
    long long off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SL_THREADS / 32; ++w)
    {
        const long long c = sWarp1[w];
        if (w < warp) off += c;
        tot += c;
    }
    __syncthreads();
    total = tot;
    return off + x - v;  // exclusive
}

// blockCounts[list * gridDim.x + block]
template <int MODE>
__global__ void __launch_bounds__(SL_THREADS) selCountKernel(SelArgs g, int64_t* blockCounts, int* err)
{
    const int list = blockIdx.y;
    if (MODE != 0 && blockIdx.x == 0 && threadIdx.x == 0)
    {
        // the grid was sized from an estimate of the column population: a fuller column must not go unnoticed
        const int64_t len = list == 0 ? int64_t(*g.rangeA1) - *g.rangeA0 : int64_t(*g.rangeB1) - *g.rangeB0;
        if (len > int64_t(gridDim.x) * blockDim.x) *err = 1;
    }
    const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    int64_t unit;
    long long total;
    blockScan1(selPredicate<MODE>(g, list, j, unit) ? 1 : 0, total);
    if (threadIdx.x == 0) blockCounts[list * int64_t(gridDim.x) + blockIdx.x] = total;
}

// one block per list: exclusive scan of its block counts in place, total -> totals[list]
__global__ void __launch_bounds__(SL_THREADS) selScanKernel(int64_t* blockCounts, int64_t numBlocks, int64_t* totals)
{
    int64_t* bc = blockCounts + blockIdx.x * numBlocks;
    __shared__ long long sCarry;
    if (threadIdx.x == 0) sCarry = 0;
    __syncthreads();
    for (int64_t base = 0; base < numBlocks; base += SL_THREADS)
    {
        const int64_t b = base + threadIdx.x;
        const long long v = (b < numBlocks) ? bc[b] : 0;
        long long total;
        const long long excl = blockScan1(v, total);
        if (b < numBlocks) bc[b] = excl + sCarry;
        __syncthreads();
        if (threadIdx.x == 0) sCarry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = sCarry;
}

// idx[list] receives the selected units in order; a list longer than cap raises err (the count stays the true one)
template <int MODE>
__global__ void __launch_bounds__(SL_THREADS)
    selIndexKernel(SelArgs g, const int64_t* blockOffsets, int32_t* idxA, int32_t* idxB, int64_t cap, int* err)
{
    const int list = blockIdx.y;
    const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    int64_t unit;
    const bool sel = selPredicate<MODE>(g, list, j, unit);
    long long total;
    const long long at = blockOffsets[list * int64_t(gridDim.x) + blockIdx.x] + blockScan1(sel ? 1 : 0, total);
    if (sel)
    {
        if (at < cap) (list == 0 ? idxA : idxB)[at] = static_cast<int32_t>(unit);
        else *err = 1;
    }
}

__device__ __forceinline__ void publishWhenLast(unsigned int* ticket, volatile long long* hdrL, volatile long long* hdrR,
                                                long long countL, long long countR, long long seq)
{
    // the last block to finish publishes: every block's records are visible system wide before the flags are (one
    // cumulative system-scope fence per block behind the block barrier)
    __shared__ bool sLastBlock;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence_system();
        sLastBlock = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (sLastBlock && threadIdx.x == 0)
    {
        __threadfence_system();
        hdrL[0] = countL;
        hdrR[0] = countR;
        __threadfence_system();
        hdrL[1] = seq;
        hdrR[1] = seq;
        *ticket = 0;
    }
}

// full records of the units leaving to the left / right -> the neighbours' migration regions, counts + sequence number
// behind them.  counts: device {toLeft, toRight}; the grid covers 2 * cap units.
__global__ void __launch_bounds__(SL_THREADS)
    migratePushKernel(AtomsView a, int apm, const int32_t* __restrict__ idxA, const int32_t* __restrict__ idxB,
                      const int64_t* __restrict__ counts, int64_t cap, double shiftL, double shiftR, double* dstL, double* dstR,
                      long long* hdrL, long long* hdrR, long long seq, unsigned int* ticket)
{
    const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    const int64_t perList = cap * apm;
    const int list = k < perList ? 0 : 1;
    const int64_t kk = k - list * perList, unit = kk / apm;
    const int64_t cnt = min(counts[list], cap);
    if (kk < perList && unit < cnt)
    {
        const int64_t i = int64_t((list == 0 ? idxA : idxB)[unit]) * apm + kk % apm;
        const double4 p = ld4(a.pos + i);
        double* r = (list == 0 ? dstL : dstR) + kk * SL_RECORD;
        r[0] = p.x + (list == 0 ? shiftL : shiftR);
        r[1] = p.y;
        r[2] = p.z;
        r[3] = p.w;
        for (int d = 0; d < 3; ++d)
        {
            r[4 + d] = a.vel[d][i];
            r[7 + d] = a.force[d][i];
        }
        r[10] = a.mass[i];
        r[11] = a.charge[i];
        r[12] = a.relMass[i];
        r[13] = __longlong_as_double(a.gid[i]);
    }
    publishWhenLast(ticket, hdrL, hdrR, counts[0], counts[1], seq);
}

// waits for both neighbours' records, appends them behind the n resident atoms (right neighbour's first, as the NCCL
// path does), clears the flags of the arrivals and reports {toLeft, toRight, fromLeft, fromRight, seq} to pinned host
// memory.  hdr: own header words SL_HW_MIG...; report: device copy of the counts for the kernels that follow.
__global__ void __launch_bounds__(SL_THREADS)
    migrateUnpackKernel(AtomsView a, int apm, int64_t n, const double* fromLeft, const double* fromRight,
                        const long long* hdr, long long seq, int64_t cap, signed char* unitFlag, const int64_t* sent,
                        int64_t* recvDev, volatile long long* hReport, int* err)
{
    if (threadIdx.x == 0)
    {
        const volatile long long* h = hdr;
        while (h[1] < seq) {}
        while (h[3] < seq) {}
    }
    __syncthreads();
    __threadfence_system();
    const long long cL = __ldcv(hdr), cR = __ldcv(hdr + 2);
    const bool bad = cL > cap || cR > cap || sent[0] > cap || sent[1] > cap;
    const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (!bad && k < (cL + cR) * apm)
    {
        const double* r = (k < cR * apm) ? fromRight + k * SL_RECORD : fromLeft + (k - cR * apm) * SL_RECORD;
        double v[SL_RECORD];
#pragma unroll
        for (int q = 0; q < SL_RECORD; q += 2)
        {
            const double2 w = __ldcv(reinterpret_cast<const double2*>(r + q));  // written by another GPU: past L1
            v[q] = w.x;
            v[q + 1] = w.y;
        }
        const int64_t i = n + k;
        st4(a.pos + i, make_double4(v[0], v[1], v[2], v[3]));
        for (int d = 0; d < 3; ++d)
        {
            a.vel[d][i] = v[4 + d];
            a.force[d][i] = v[7 + d];
        }
        a.mass[i] = v[10];
        a.charge[i] = v[11];
        a.relMass[i] = v[12];
        a.gid[i] = __double_as_longlong(v[13]);
        if (k % apm == 0) unitFlag[i / apm] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        recvDev[0] = cL;
        recvDev[1] = cR;
        if (bad) *err = 1;
        hReport[0] = sent[0];
        hReport[1] = sent[1];
        hReport[2] = cL;
        hReport[3] = cR;
        hReport[5] = bad ? 1 : 0;
        __threadfence_system();
        hReport[4] = seq;
    }
}

// the per-step halo with the counts on the device (rebuild flavour of haloPushKernel): also publishes the counts
__global__ void __launch_bounds__(SL_THREADS)
    haloPushCountedKernel(const double4* __restrict__ pos, const double4* __restrict__ upos, int apm,
                          const int32_t* __restrict__ idxLow, const int32_t* __restrict__ idxHigh,
                          const int64_t* __restrict__ counts, int64_t cap, double shiftL, double shiftR, double4* dstL,
                          double4* dstR, unsigned long long* flagL, unsigned long long* flagR, long long* cntL, long long* cntR,
                          unsigned long long seq, unsigned int* ticket, const int* __restrict__ stop)
{
    if (stop != nullptr && *stop != 0) return;  // a step queued behind the one that asked for a rebuild
    const int rec = apm + (apm > 1 ? 1 : 0);  // double4 per unit: the atoms' positions (+ the centre of mass)
    const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    // list 0 occupies the first min(counts[0], cap) * rec threads, list 1 follows directly (the grid covers 2 * cap units
    // at a rebuild, exactly the two lists otherwise)
    const int64_t n0 = min(counts[0], cap) * rec, n1 = min(counts[1], cap) * rec;
    if (k < n0 + n1)
    {
        const int list = k < n0 ? 0 : 1;
        const int64_t kk = k - list * n0, unit = kk / rec;
        const int part = static_cast<int>(kk % rec);
        const int64_t u = (list == 0 ? idxLow : idxHigh)[unit];
        double4 p = (part < apm) ? ld4(pos + u * apm + part) : ld4(upos + u);
        p.x += (list == 0 ? shiftL : shiftR);
        st4((list == 0 ? dstL : dstR) + kk, p);
    }
    // one system-scope fence per block behind the block barrier (cumulative: it orders the stores of the whole block
    // before the ticket), then the last block to take a ticket publishes
    __shared__ bool sLast;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence_system();
        sLast = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (sLast && threadIdx.x == 0)
    {
        __threadfence_system();
        *reinterpret_cast<volatile long long*>(cntL) = counts[0];
        *reinterpret_cast<volatile long long*>(cntR) = counts[1];
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(flagL) = seq;
        *reinterpret_cast<volatile unsigned long long*>(flagR) = seq;
        *ticket = 0;
    }
}

// rebuild flavour of haloPullKernel: the counts come with the records; reports {sendLow, sendHigh, fromLeft, fromRight}
__global__ void __launch_bounds__(SL_THREADS)
    haloPullCountedKernel(const double4* fromLeft, const double4* fromRight, const unsigned long long* flags,
                          const long long* cnt, unsigned long long seq, int apm, int64_t cap, double4* dstAtoms,
                          double4* dstUnits, const int64_t* sent, int64_t* recvDev, volatile long long* hReport, int* err,
                          const int* __restrict__ stop)
{
    if (stop != nullptr && *stop != 0) return;  // a step queued behind the one that asked for a rebuild
    if (threadIdx.x == 0)
    {
        const volatile unsigned long long* f = flags;
        while (f[0] < seq) {}
        while (f[1] < seq) {}
    }
    __syncthreads();
    __threadfence_system();
    const long long cL = __ldcv(cnt), cR = __ldcv(cnt + 2);
    const bool bad = cL > cap || cR > cap || sent[0] > cap || sent[1] > cap;
    const int rec = apm + (apm > 1 ? 1 : 0);
    const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (!bad && k < (cL + cR) * rec)
    {
        const bool left = k < cL * rec;
        const int64_t kk = left ? k : k - cL * rec;
        const double2* q = reinterpret_cast<const double2*>((left ? fromLeft : fromRight) + kk);
        const double2 lo = __ldcv(q), hi = __ldcv(q + 1);
        const int64_t unit = (left ? 0 : cL) + kk / rec;
        const int part = static_cast<int>(kk % rec);
        if (part < apm) st4(dstAtoms + unit * apm + part, make_double4(lo.x, lo.y, hi.x, hi.y));
        else st4(dstUnits + unit, make_double4(lo.x, lo.y, hi.x, hi.y));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        recvDev[0] = cL;
        recvDev[1] = cR;
        if (bad) *err = 1;
        hReport[14] = *err;  // raised by the selections in front of this kernel (list or column capacity)
        hReport[8] = sent[0];
        hReport[9] = sent[1];
        hReport[10] = cL;
        hReport[11] = cR;
        hReport[13] = bad ? 1 : 0;
        __threadfence_system();
        hReport[12] = static_cast<long long>(seq);
    }
}

// (j, k) cell keys of the received halo units and the prefix arrays of the two halo columns, counts on the device
__global__ void haloKeyCountedKernel(const double4* upos, int64_t first, const int64_t* recv, GridDev g, uint32_t* keys)
{
    const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (k >= recv[0] + recv[1]) return;
    const double4 p = ld4nc(upos + first + k);
    keys[k] = static_cast<uint32_t>(locate1(g, p.y, 1) * g.n[2] + locate1(g, p.z, 2));
}
__global__ void haloCellStartCountedKernel(const uint32_t* keys, const int64_t* recv, int64_t numCells, int64_t first,
                                           int32_t* startLeft, int32_t* startRight)
{
    const int64_t c = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (c > numCells) return;
    for (int side = 0; side < 2; ++side)
    {
        const uint32_t* ks = keys + (side == 0 ? 0 : recv[0]);
        int64_t lo = 0, hi = recv[side];
        while (lo < hi)
        {
            const int64_t mid = (lo + hi) >> 1;
            if (int64_t(ks[mid]) < c) lo = mid + 1;
            else hi = mid;
        }
        (side == 0 ? startLeft : startRight)[c] = static_cast<int32_t>(first + (side == 0 ? 0 : recv[0]) + lo);
    }
}

// (j, k) cell of a halo atom in the receiver's grid (identical y / z grid on every rank)
__global__ void haloKeyKernel(const double4* pos, int64_t first, int64_t n, GridDev g, uint32_t* keys)
{
    const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (k >= n) return;
    const double4 p = ld4nc(pos + first + k);
    keys[k] = static_cast<uint32_t>(locate1(g, p.y, 1) * g.n[2] + locate1(g, p.z, 2));
}

// defined in neighbor.cu
int atomsCellSortDrop(mrmd_b200_atoms* a, int64_t begin, int64_t end, const double* delta, const double* gridMin,
                      const double* gridMax, const signed char* dropFlags, cudaStream_t st);
int cellStartFromKeys(const uint32_t* sortedKeys, int64_t n, int64_t numCells, int32_t* cellStart, int32_t offset,
                      cudaStream_t st);
int atomsCellSortSlab(mrmd_b200_atoms* a, int64_t end, const double* delta, const mrmd_b200_subdomain* sub, cudaStream_t st);
}  // namespace mrmd_b200

struct mrmd_b200_slab
{
    mrmd_b200_md_config cfg{};
    int rank = 0, nranks = 1, left = 0, right = 0;
    double globalMin[3]{}, globalMax[3]{};
    double shiftToLeft = 0.0, shiftToRight = 0.0;  // added to x when sending across the global boundary
    mrmd_b200_subdomain sub{};                     // this rank's slab
    ncclComm_t comm = nullptr;
    mrmd_b200_atoms* atoms = nullptr;  // not owned
    mrmd_b200_verlet* list = nullptr;
    mrmd_b200_lj* lj = nullptr;
    mrmd_b200_adress* adress = nullptr;  // AdResS mode (tiled kernels)
    mrmd_b200_molecules* mols = nullptr;            // apm > 1: molecules of apm consecutive atoms (centres of mass, SHAKE)
    mrmd_b200_constraints* constraints = nullptr;   // apm > 1 with numConstraintIterations > 0
    mrmd_b200_thermo* thermo = nullptr;  // bins over the GLOBAL box, density all-reduced before every update
    double maxDisplacement = DBL_MAX;
    bool postPending = false;
    // MRMD_B200_SLAB_PROFILE=1: phases separated by stream syncs, wall-clock sums printed at destroy (diagnostic only)
    bool profile = false;
    double prof[6] = {0, 0, 0, 0, 0, 0};  // pre, decision, rebuild, halo refresh, force, steps
    double profRebuild[4] = {0, 0, 0, 0};  // migration (to the first host round trip), sort, face lists + halo, list build
    int64_t step = 0, rebuilds = 0, storedPairsNow = 0;
    int64_t haloLeftCount = 0, haloRightCount = 0;   // received
    int64_t sendLeftCount = 0, sendRightCount = 0;   // boundary atoms sent every step
    mrmd_b200::DevBuf flags, blockCounts, idxLow, idxHigh, sendBuf, recvBuf, haloKeys, haloStartLeft, haloStartRight;
    int64_t* dTotals = nullptr;  // [0..1] select totals, [2..5] exchanged counts
    int64_t* hTotals = nullptr;  // pinned, 8 entries
    // per-step halo over peer memory (IPC): own buffer, the neighbours' buffers mapped into this process
    bool p2p = false;
    int apm = 1;             // atoms per unit of selection (1, or the atoms of a molecule)
    int64_t p2pCap = 0;      // units per halo region
    int64_t migCap = 0;      // units per migration region
    int64_t colBound = 0;    // upper bound of the units in one cell column (grid of the face selection)
    long long migSeq = 0;
    // a->posEpoch at the last sort and the kick / drift kernels since: any other change of the positions (an upload
    // through the host-buffer path, a caller's atoms_write) makes the rebuild fall back to the full selection
    int64_t posEpochAtSort = -1, presSinceSort = 0;
    long long* hReport = nullptr;  // pinned, 16 words: [0..5] migration {toL, toR, fromL, fromR, seq, err}, [8..13] halo lists
    int* dErr = nullptr;
    mrmd_b200::DevBuf migIdxA, migIdxB;
    double4* p2pBuf = nullptr;
    mrmd_b200::PeerBuffers peers{};  // every rank's buffer (own entry = p2pBuf)
    double4* peerLeftBuf = nullptr;
    double4* peerRightBuf = nullptr;
    double* hDecide = nullptr;  // pinned, written by maxDisplacementGatherKernel: {max |dx|^2, step}
    double decideSeq = 0.0;
    unsigned long long haloSeq = 0;
    unsigned int* dPushTicket = nullptr;
    double* dScalars = nullptr;  // allreduce scratch
    double* hScalars = nullptr;  // pinned
    std::vector<cudaEvent_t> events;
    mrmd_b200::HostPipe hp;  // host-buffer path (mrmd_b200_slab_run_host)
    // steps queued ahead of the host (slabRunQueued): the displacement criterion is evaluated by the gather kernel
    int* dStop = nullptr;      // {stop, local step that stopped, non-finite displacement}
    double* dAccum = nullptr;  // accumulated displacement (the device copy of maxDisplacement)
    int* hStop = nullptr;      // pinned: dStop, and behind it the accumulated displacement
    int64_t stepsSinceRebuild = 0, lastRebuildInterval = 4;
    // MRMD_B200_SLAB_EVENTS=1: CUDA events at the phase boundaries of every step (no synchronisation added); the mean
    // in-stream time of every phase is printed at destroy (diagnostic only)
    // the per-step halo push runs on its own stream next to the displacement gather (both only need the new positions)
    cudaStream_t sPush = nullptr;
    cudaEvent_t evPreDone = nullptr, evPushDone = nullptr;
    bool evProfile = false;
    std::vector<std::pair<int, cudaEvent_t>> evMarks;
};

namespace mrmd_b200
{
template <int MODE>
static int selectTwo(mrmd_b200_slab* sl, int64_t first, int64_t n, double lowBound, double highBound, int64_t* nLow,
                     int64_t* nHigh, cudaStream_t st)
{
    mrmd_b200_atoms* a = sl->atoms;
    *nLow = *nHigh = 0;
    if (n <= 0) return 0;
    const int blocks = gridFor(n, SL_THREADS);
    MB_TRY(sl->blockCounts.reserve(size_t(blocks) * 16));
    int64_t* bc = sl->blockCounts.as<int64_t>();
    const signed char* flag = sl->flags.as<signed char>();
    selectCountKernel<MODE><<<blocks, SL_THREADS, 0, st>>>(a->v.pos, flag, first, n, lowBound, highBound, bc);
    MB_LAUNCHED();
    selectScanKernel<<<1, SL_THREADS, 0, st>>>(bc, blocks, sl->dTotals);
    MB_LAUNCHED();
    MB_CUDA(cudaMemcpyAsync(sl->hTotals, sl->dTotals, 16, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    *nLow = sl->hTotals[0];
    *nHigh = sl->hTotals[1];
    MB_TRY(sl->idxLow.reserve(size_t(std::max<int64_t>(*nLow, 1)) * 4));
    MB_TRY(sl->idxHigh.reserve(size_t(std::max<int64_t>(*nHigh, 1)) * 4));
    if (*nLow + *nHigh > 0)
    {
        selectIndexKernel<MODE><<<blocks, SL_THREADS, 0, st>>>(a->v.pos, flag, first, n, lowBound, highBound, bc,
                                                               sl->idxLow.as<int32_t>(), sl->idxHigh.as<int32_t>());
        MB_LAUNCHED();
    }
    return 0;
}

// counts to the neighbours: I tell left how many atoms come from its right side and vice versa
static int exchangeCounts(mrmd_b200_slab* sl, int64_t toLeft, int64_t toRight, int64_t* fromLeft, int64_t* fromRight,
                          cudaStream_t st)
{
    sl->hTotals[2] = toLeft;
    sl->hTotals[3] = toRight;
    MB_CUDA(cudaMemcpyAsync(sl->dTotals + 2, sl->hTotals + 2, 16, cudaMemcpyHostToDevice, st));
    MB_NCCL(g_nccl.groupStart());
    MB_NCCL(g_nccl.send(sl->dTotals + 2, 1, ncclInt64, sl->left, sl->comm, st));
    MB_NCCL(g_nccl.send(sl->dTotals + 3, 1, ncclInt64, sl->right, sl->comm, st));
    MB_NCCL(g_nccl.recv(sl->dTotals + 5, 1, ncclInt64, sl->right, sl->comm, st));  // the right rank's "toLeft"
    MB_NCCL(g_nccl.recv(sl->dTotals + 4, 1, ncclInt64, sl->left, sl->comm, st));   // the left rank's "toRight"
    MB_NCCL(g_nccl.groupEnd());
    MB_CUDA(cudaMemcpyAsync(sl->hTotals + 4, sl->dTotals + 4, 16, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    *fromLeft = sl->hTotals[4];
    *fromRight = sl->hTotals[5];
    return 0;
}

static int migrate(mrmd_b200_slab* sl, cudaStream_t st)
{
    mrmd_b200_atoms* a = sl->atoms;
    const int64_t n = a->numLocal;
    a->numGhost = 0;
    a->size = n;
    MB_TRY(sl->flags.reserve(size_t(a->capacity) + 64));
    if (n > 0)
    {
        slabWrapFlagKernel<<<gridFor(n, 256), 256, 0, st>>>(a->v.pos, n, toDev(sl->sub), sl->flags.as<signed char>());
        MB_LAUNCHED();
    }
    int64_t toLeft = 0, toRight = 0, fromLeft = 0, fromRight = 0;
    MB_TRY(selectTwo<0>(sl, 0, n, 0.0, 0.0, &toLeft, &toRight, st));
    MB_TRY(exchangeCounts(sl, toLeft, toRight, &fromLeft, &fromRight, st));
    const int64_t nSend = toLeft + toRight, nRecv = fromLeft + fromRight;
    MB_TRY(sl->sendBuf.reserve(size_t(std::max<int64_t>(nSend, 1)) * SL_RECORD * 8));
    MB_TRY(sl->recvBuf.reserve(size_t(std::max<int64_t>(nRecv, 1)) * SL_RECORD * 8));
    double* sb = sl->sendBuf.as<double>();
    double* rb = sl->recvBuf.as<double>();
    if (toLeft > 0)
    {
        packRecordsKernel<<<gridFor(toLeft, 256), 256, 0, st>>>(a->v, sl->idxLow.as<int32_t>(), toLeft, sl->shiftToLeft, sb);
        MB_LAUNCHED();
    }
    if (toRight > 0)
    {
        packRecordsKernel<<<gridFor(toRight, 256), 256, 0, st>>>(a->v, sl->idxHigh.as<int32_t>(), toRight,
                                                                sl->shiftToRight, sb + toLeft * SL_RECORD);
        MB_LAUNCHED();
    }
    MB_NCCL(g_nccl.groupStart());
    if (toLeft > 0) MB_NCCL(g_nccl.send(sb, size_t(toLeft) * SL_RECORD, ncclDouble, sl->left, sl->comm, st));
    if (toRight > 0)
        MB_NCCL(g_nccl.send(sb + toLeft * SL_RECORD, size_t(toRight) * SL_RECORD, ncclDouble, sl->right, sl->comm, st));
    if (fromRight > 0) MB_NCCL(g_nccl.recv(rb, size_t(fromRight) * SL_RECORD, ncclDouble, sl->right, sl->comm, st));
    if (fromLeft > 0)
        MB_NCCL(g_nccl.recv(rb + fromRight * SL_RECORD, size_t(fromLeft) * SL_RECORD, ncclDouble, sl->left, sl->comm, st));
    MB_NCCL(g_nccl.groupEnd());
    // arrivals are appended, leavers are dropped by the cell sort (they sort behind the last cell)
    MB_TRY(atomsEnsureCapacity(a, n + nRecv, st));
    if (sl->flags.bytes < size_t(n + nRecv) + 64)
    {
        // grow the flag array keeping the flags of the n resident atoms
        mrmd_b200::DevBuf bigger;
        MB_TRY(bigger.reserve(size_t(a->capacity) * 2 + 64));
        if (n > 0) MB_CUDA(cudaMemcpyAsync(bigger.p, sl->flags.p, size_t(n), cudaMemcpyDeviceToDevice, st));
        MB_CUDA(cudaStreamSynchronize(st));
        sl->flags.release();
        sl->flags = bigger;
    }
    a->size = n + nRecv;
    if (nRecv > 0)
    {
        unpackRecordsKernel<<<gridFor(nRecv, 256), 256, 0, st>>>(a->v, n, nRecv, rb, sl->flags.as<signed char>());
        MB_LAUNCHED();
    }
    const double cutoff = sl->cfg.rc + sl->cfg.skin;
    const double delta[3] = {cutoff, cutoff, 0.25 * cutoff};  // as md.cu: fine z order inside a cell column
    MB_TRY(atomsCellSortDrop(a, 0, n + nRecv, delta, sl->sub.minCorner, sl->sub.maxCorner, sl->flags.as<signed char>(), st));
    a->numLocal = n + nRecv - nSend;
    a->size = a->numLocal;
    a->lcEnd = a->numLocal;
    return 0;
}

static int haloExchange(mrmd_b200_slab* sl, cudaStream_t st)
{
    mrmd_b200_atoms* a = sl->atoms;
    const GridDev& g = a->lcGrid;
    const int64_t perX = int64_t(g.n[1]) * g.n[2];
    const int64_t n = a->numLocal;
    // boundary atoms live in the first / last cell column (ghost layer thickness <= cell size)
    MB_CUDA(cudaMemcpyAsync(sl->hTotals + 6, a->lcCellStart.as<int32_t>() + perX, 4, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaMemcpyAsync(reinterpret_cast<int32_t*>(sl->hTotals + 6) + 1,
                            a->lcCellStart.as<int32_t>() + (int64_t(g.n[0]) - 1) * perX, 4, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    const int64_t firstColEnd = reinterpret_cast<int32_t*>(sl->hTotals + 6)[0];
    const int64_t lastColStart = reinterpret_cast<int32_t*>(sl->hTotals + 6)[1];
    int64_t nl = 0, dummy = 0, nh = 0;
    // low face: x < minInner within column 0 ; high face: x >= maxInner within column nx-1
    MB_TRY(selectTwo<1>(sl, 0, firstColEnd, sl->sub.minInnerCorner[0], DBL_MAX, &nl, &dummy, st));
    // keep the low list: copy it aside before the second selection reuses the buffers
    MB_TRY(sl->sendBuf.reserve(size_t(std::max<int64_t>(nl, 1)) * 4 + 64));
    if (nl > 0) MB_CUDA(cudaMemcpyAsync(sl->sendBuf.p, sl->idxLow.p, size_t(nl) * 4, cudaMemcpyDeviceToDevice, st));
    MB_TRY(selectTwo<1>(sl, lastColStart, n - lastColStart, -DBL_MAX, sl->sub.maxInnerCorner[0], &dummy, &nh, st));
    if (nl > 0)
    {
        MB_TRY(sl->idxLow.reserve(size_t(nl) * 4));
        MB_CUDA(cudaMemcpyAsync(sl->idxLow.p, sl->sendBuf.p, size_t(nl) * 4, cudaMemcpyDeviceToDevice, st));
    }
    sl->sendLeftCount = nl;
    sl->sendRightCount = nh;
    int64_t fromLeft = 0, fromRight = 0;
    MB_TRY(exchangeCounts(sl, nl, nh, &fromLeft, &fromRight, st));
    sl->haloLeftCount = fromLeft;
    sl->haloRightCount = fromRight;
    MB_TRY(atomsEnsureCapacity(a, n + fromLeft + fromRight, st));
    a->numGhost = fromLeft + fromRight;
    a->size = n + a->numGhost;
    return 0;
}

// ---- peer-memory data plane (host side) ------------------------------------------------------------------------
static int unitRecord(const mrmd_b200_slab* sl) { return sl->apm + (sl->apm > 1 ? 1 : 0); }  // double4 per halo unit
static int64_t haloRegionOffset(const mrmd_b200_slab* sl, int parity, int side)
{
    return SL_P2P_HEADER + int64_t(parity * 2 + side) * sl->p2pCap * unitRecord(sl);
}
static int64_t migRegionDoubles(const mrmd_b200_slab* sl) { return ((sl->migCap * sl->apm * SL_RECORD + 3) / 4) * 4; }
static int64_t migRegionOffset(const mrmd_b200_slab* sl, int side)
{
    return haloRegionOffset(sl, 2, 0) + side * (migRegionDoubles(sl) / 4);
}
static long long* headerWords(double4* buf) { return reinterpret_cast<long long*>(buf); }
// the position array the units are selected by, and its halo destination behind the local units
static double4* unitPositions(mrmd_b200_slab* sl);

// my face units -> the neighbours' halo regions of this sequence number's parity.  countsDev != nullptr (rebuild): the
// counts were just produced on the device; otherwise the lists of the last rebuild are re-sent.
static int haloPush(mrmd_b200_slab* sl, bool rebuild, cudaStream_t st, const int* stop = nullptr)
{
    mrmd_b200_atoms* a = sl->atoms;
    const unsigned long long seq = ++sl->haloSeq;
    const int parity = static_cast<int>(seq & 1);
    const int rec = unitRecord(sl);
    if (!rebuild)
        MB_REQUIRE(sl->sendLeftCount <= sl->p2pCap && sl->sendRightCount <= sl->p2pCap, "slab: the position halo exceeds the peer buffer");
    // my low face goes to the left neighbour's "from right" region, my high face to the right neighbour's "from left"
    double4* dstL = sl->peerLeftBuf + haloRegionOffset(sl, parity, 1);
    double4* dstR = sl->peerRightBuf + haloRegionOffset(sl, parity, 0);
    long long* hL = headerWords(sl->peerLeftBuf);
    long long* hR = headerWords(sl->peerRightBuf);
    const int64_t units = rebuild ? 2 * sl->p2pCap : sl->sendLeftCount + sl->sendRightCount;
    haloPushCountedKernel<<<std::max(1, gridFor(units * rec, SL_THREADS)), SL_THREADS, 0, st>>>(
        a->v.pos, unitPositions(sl), sl->apm, sl->idxLow.as<int32_t>(), sl->idxHigh.as<int32_t>(), sl->dTotals + 2, sl->p2pCap,
        sl->shiftToLeft, sl->shiftToRight, dstL, dstR, reinterpret_cast<unsigned long long*>(hL + SL_HW_HALO + parity * 2 + 1),
        reinterpret_cast<unsigned long long*>(hR + SL_HW_HALO + parity * 2 + 0), hL + SL_HW_HALOCOUNT + parity * 4 + 2,
        hR + SL_HW_HALOCOUNT + parity * 4 + 0, seq, sl->dPushTicket, stop);
    MB_LAUNCHED();
    return 0;
}

// the neighbours' face units of the current sequence number -> behind my local units
static int haloPull(mrmd_b200_slab* sl, bool rebuild, cudaStream_t st, const int* stop = nullptr)
{
    mrmd_b200_atoms* a = sl->atoms;
    const unsigned long long seq = sl->haloSeq;
    const int parity = static_cast<int>(seq & 1);
    const int rec = unitRecord(sl);
    const int64_t units = rebuild ? 2 * sl->p2pCap : sl->haloLeftCount + sl->haloRightCount;
    const int64_t nUnits = a->numLocal / sl->apm;
    long long* h = headerWords(sl->p2pBuf);
    haloPullCountedKernel<<<std::max(1, gridFor(units * rec, SL_THREADS)), SL_THREADS, 0, st>>>(
        sl->p2pBuf + haloRegionOffset(sl, parity, 0), sl->p2pBuf + haloRegionOffset(sl, parity, 1),
        reinterpret_cast<const unsigned long long*>(h + SL_HW_HALO + parity * 2), h + SL_HW_HALOCOUNT + parity * 4, seq, sl->apm,
        sl->p2pCap, a->v.pos + a->numLocal, unitPositions(sl) + nUnits, sl->dTotals + 2, sl->dTotals + 6, sl->hReport, sl->dErr,
        stop);
    MB_LAUNCHED();
    return 0;
}

static int pollReport(mrmd_b200_slab* sl, int word, long long seq, cudaStream_t st, const char* what)
{
    volatile long long* h = sl->hReport;
    for (unsigned spin = 0; h[word] != seq; ++spin)
    {
        if ((spin & 0xfff) == 0xfff)
        {
            const cudaError_t q = cudaStreamQuery(st);
            if (q != cudaSuccess && q != cudaErrorNotReady) MB_CUDA(q);
            if (q == cudaSuccess && h[word] != seq)
            {
                setLastError(std::string("slab: ") + what + " finished without a report");
                return MRMD_B200_EINVAL;
            }
        }
    }
    return 0;
}

// positions of the boundary atoms -> neighbours' halo slots [n, n + haloLeft) and [n + haloLeft, ...) (NCCL path)
static int haloRefresh(mrmd_b200_slab* sl, cudaStream_t st)
{
    mrmd_b200_atoms* a = sl->atoms;
    const int64_t n = a->numLocal, nl = sl->sendLeftCount, nh = sl->sendRightCount;
    if (sl->p2p)
    {
        MB_TRY(haloPush(sl, false, st));
        return haloPull(sl, false, st);
    }
    MB_TRY(sl->sendBuf.reserve(size_t(std::max<int64_t>(nl + nh, 1)) * 32));
    double4* sb = sl->sendBuf.as<double4>();
    if (nl > 0)
    {
        packPositionsKernel<<<gridFor(nl, 256), 256, 0, st>>>(a->v.pos, sl->idxLow.as<int32_t>(), nl, sl->shiftToLeft, sb);
        MB_LAUNCHED();
    }
    if (nh > 0)
    {
        packPositionsKernel<<<gridFor(nh, 256), 256, 0, st>>>(a->v.pos, sl->idxHigh.as<int32_t>(), nh, sl->shiftToRight,
                                                              sb + nl);
        MB_LAUNCHED();
    }
    MB_NCCL(g_nccl.groupStart());
    if (nl > 0) MB_NCCL(g_nccl.send(sb, size_t(nl) * 4, ncclDouble, sl->left, sl->comm, st));
    if (nh > 0) MB_NCCL(g_nccl.send(sb + nl, size_t(nh) * 4, ncclDouble, sl->right, sl->comm, st));
    if (sl->haloRightCount > 0)
        MB_NCCL(g_nccl.recv(a->v.pos + n + sl->haloLeftCount, size_t(sl->haloRightCount) * 4, ncclDouble, sl->right,
                            sl->comm, st));
    if (sl->haloLeftCount > 0)
        MB_NCCL(g_nccl.recv(a->v.pos + n, size_t(sl->haloLeftCount) * 4, ncclDouble, sl->left, sl->comm, st));
    MB_NCCL(g_nccl.groupEnd());
    return 0;
}

static const char* const EV_NAMES[] = {"(step start)", "kick/drift", "halo push", "displacement gather", "halo pull", "force",
                                       "post", "rb: wrap+select+migrate", "rb: host round trip 1", "rb: sort", "rb: face lists",
                                       "rb: halo push+pull", "rb: halo cells", "rb: list build (+round trip 2)"};
static void evMark(mrmd_b200_slab* sl, int tag, cudaStream_t st)
{
    if (!sl->evProfile || sl->evMarks.size() > 40000) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    sl->evMarks.emplace_back(tag, e);
}
static void evReport(mrmd_b200_slab* sl)
{
    if (!sl->evProfile || sl->evMarks.size() < 2) return;
    cudaDeviceSynchronize();
    double sum[16] = {0};
    long cnt[16] = {0};
    long steps = 0;
    // the first quarter is warm-up (lattice melting, first-touch allocations)
    const size_t first = sl->evMarks.size() / 4;
    for (size_t k = first + 1; k < sl->evMarks.size(); ++k)
    {
        float ms = 0.f;
        const int tag = sl->evMarks[k].first;
        if (tag == 0)
        {
            ++steps;
            // the gap between the end of a step and the start of the next one (host work between slabStep calls)
            if (cudaEventElapsedTime(&ms, sl->evMarks[k - 1].second, sl->evMarks[k].second) == cudaSuccess) sum[0] += ms * 1e3;
            continue;
        }
        if (cudaEventElapsedTime(&ms, sl->evMarks[k - 1].second, sl->evMarks[k].second) != cudaSuccess) continue;
        sum[tag] += ms * 1e3;
        cnt[tag] += 1;
    }
    std::fprintf(stderr, "[mrmd_b200 slab rank %d] in-stream us per step over %ld steps:", sl->rank, steps);
    double total = 0.0;
    for (int t = 0; t < 14; ++t)
        if (sum[t] > 0.0)
        {
            std::fprintf(stderr, " %s %.1f (x%ld)%s", EV_NAMES[t], sum[t] / std::max(steps, 1L), cnt[t], t < 13 ? "," : "");
            total += sum[t];
        }
    std::fprintf(stderr, " | total %.1f\n", total / std::max(steps, 1L));
    for (auto& m : sl->evMarks) cudaEventDestroy(m.second);
    sl->evMarks.clear();
}

static double profMark(mrmd_b200_slab* sl, cudaStream_t st, double& last)
{
    if (!sl->profile) return 0.0;
    cudaStreamSynchronize(st);
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    const double now = ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
    const double d = now - last;
    last = now;
    return d;
}

static double4* unitPositions(mrmd_b200_slab* sl) { return sl->apm > 1 ? sl->mols->v.pos : sl->atoms->v.pos; }

// apm > 1: room for `units` molecules with atomsOffset = apm * m, numAtoms = apm for every slot (arrivals and halo
// molecules use the slots behind the local ones)
static int slabEnsureMolecules(mrmd_b200_slab* sl, int64_t units, cudaStream_t st)
{
    if (sl->apm <= 1) return 0;
    mrmd_b200_molecules* m = sl->mols;
    if (units > m->capacity)
    {
        MB_TRY(molsEnsureCapacity(m, units, st));
        slabMoleculeInitKernel<<<gridFor(m->capacity, 256), 256, 0, st>>>(m->v, m->capacity, sl->apm);
        MB_LAUNCHED();
    }
    return 0;
}

// The rebuild over peer memory: two host round trips (the migration counts, the list statistics).
static int slabRebuildP2P(mrmd_b200_slab* sl, cudaStream_t st)
{
    mrmd_b200_atoms* a = sl->atoms;
    const int apm = sl->apm;
    const int64_t n = a->numLocal, nUnits = n / apm;
    a->numGhost = 0;
    a->size = n;
    // room for the arrivals of the migration and, after it, for the halo units
    MB_TRY(atomsEnsureCapacity(a, n + 2 * std::max(sl->migCap, sl->p2pCap) * apm, st));
    MB_TRY(slabEnsureMolecules(sl, nUnits + 2 * std::max(sl->migCap, sl->p2pCap), st));
    MB_TRY(sl->flags.reserve(size_t(a->capacity) + 64));
    const int blocksAll = std::max(1, gridFor(nUnits, SL_THREADS));
    MB_TRY(sl->blockCounts.reserve(size_t(std::max<int64_t>(blocksAll, gridFor(sl->colBound, SL_THREADS))) * 16 + 64));
    int64_t* bc = sl->blockCounts.as<int64_t>();
    signed char* flag = sl->flags.as<signed char>();
    MB_CUDA(cudaMemsetAsync(sl->dErr, 0, 4, st));
    // ---- 1. wrap y / z, flag and list the leavers, push their records, take in the arrivals
    // Atoms that leave in x sat in the first / last cell column of the previous sort (they moved by less than the skin
    // since): the leavers are selected there by position, the y / z wrap and the drop of the leavers ride in the sort's
    // key kernel -- no pass over all atoms before the migration.  Needs the previous sort's cell ranges and positions
    // that only the step loop changed.
    const bool columnsOnly = apm == 1 && sl->rebuilds > 0 && a->lcValid && a->lcEnd == n && a->posEpoch == sl->posEpochAtSort + sl->presSinceSort;
    if (columnsOnly)
    {
        const GridDev& g0 = a->lcGrid;
        const int64_t perX0 = int64_t(g0.n[1]) * g0.n[2];
        const int32_t* cs = a->lcCellStart.as<int32_t>();
        SelArgs leave{};
        leave.upos = a->v.pos;
        leave.rangeA0 = cs;
        leave.rangeA1 = cs + perX0;
        leave.rangeB0 = cs + (int64_t(g0.n[0]) - 1) * perX0;
        leave.rangeB1 = cs + a->lcNumCells;
        leave.lowBound = sl->sub.minCorner[0];
        leave.highBound = sl->sub.maxCorner[0];
        const int blocksCol = std::max(1, gridFor(sl->colBound, SL_THREADS));
        selCountKernel<2><<<dim3(blocksCol, 2), SL_THREADS, 0, st>>>(leave, bc, sl->dErr);
        MB_LAUNCHED();
        selScanKernel<<<2, SL_THREADS, 0, st>>>(bc, blocksCol, sl->dTotals);
        MB_LAUNCHED();
        selIndexKernel<2><<<dim3(blocksCol, 2), SL_THREADS, 0, st>>>(leave, bc, sl->migIdxA.as<int32_t>(), sl->migIdxB.as<int32_t>(),
                                                                    sl->migCap, sl->dErr);
        MB_LAUNCHED();
    }
    else if (apm > 1)
    {
        mrmd_b200_molecules* m = sl->mols;
        m->numLocal = nUnits;
        m->numGhost = 0;
        m->size = nUnits;
        MB_TRY(mrmd_b200_molecules_update(m, a, &sl->cfg.weight, st));  // the current centres of mass
        if (nUnits > 0)
        {
            slabWrapFlagMolKernel<<<gridFor(nUnits, 256), 256, 0, st>>>(m->v.pos, a->v.pos, nUnits, apm, toDev(sl->sub), flag);
            MB_LAUNCHED();
        }
        a->posEpoch += 1;
    }
    else if (n > 0)
    {
        slabWrapFlagKernel<<<gridFor(n, 256), 256, 0, st>>>(a->v.pos, n, toDev(sl->sub), flag);
        MB_LAUNCHED();
    }
    if (!columnsOnly)
    {
        SelArgs mig{};
        mig.upos = unitPositions(sl);
        mig.flag = flag;
        mig.n = nUnits;
        selCountKernel<0><<<dim3(blocksAll, 2), SL_THREADS, 0, st>>>(mig, bc, sl->dErr);
        MB_LAUNCHED();
        selScanKernel<<<2, SL_THREADS, 0, st>>>(bc, blocksAll, sl->dTotals);
        MB_LAUNCHED();
        selIndexKernel<0><<<dim3(blocksAll, 2), SL_THREADS, 0, st>>>(mig, bc, sl->migIdxA.as<int32_t>(), sl->migIdxB.as<int32_t>(),
                                                                    sl->migCap, sl->dErr);
        MB_LAUNCHED();
    }
    const long long mseq = ++sl->migSeq;
    {
        double* dstL = reinterpret_cast<double*>(sl->peerLeftBuf + migRegionOffset(sl, 1));   // the left rank's "from right"
        double* dstR = reinterpret_cast<double*>(sl->peerRightBuf + migRegionOffset(sl, 0));  // the right rank's "from left"
        migratePushKernel<<<std::max(1, gridFor(2 * sl->migCap * apm, SL_THREADS)), SL_THREADS, 0, st>>>(
            a->v, apm, sl->migIdxA.as<int32_t>(), sl->migIdxB.as<int32_t>(), sl->dTotals, sl->migCap, sl->shiftToLeft,
            sl->shiftToRight, dstL, dstR, headerWords(sl->peerLeftBuf) + SL_HW_MIG + 2, headerWords(sl->peerRightBuf) + SL_HW_MIG,
            mseq, sl->dPushTicket);
        MB_LAUNCHED();
        migrateUnpackKernel<<<std::max(1, gridFor(2 * sl->migCap * apm, SL_THREADS)), SL_THREADS, 0, st>>>(
            a->v, apm, n, reinterpret_cast<const double*>(sl->p2pBuf + migRegionOffset(sl, 0)),
            reinterpret_cast<const double*>(sl->p2pBuf + migRegionOffset(sl, 1)), headerWords(sl->p2pBuf) + SL_HW_MIG, mseq,
            sl->migCap, flag, sl->dTotals, sl->dTotals + 4, sl->hReport, sl->dErr);
        MB_LAUNCHED();
    }
    double last = 0.0;
    profMark(sl, st, last);
    evMark(sl, 7, st);
    MB_TRY(pollReport(sl, 4, mseq, st, "migration"));  // host round trip 1: the new number of local atoms
    sl->profRebuild[0] += profMark(sl, st, last);
    MB_REQUIRE(sl->hReport[5] == 0, "slab: more atoms migrate in one rebuild than the peer buffer holds");
    const int64_t nSend = (sl->hReport[0] + sl->hReport[1]) * apm, nRecv = (sl->hReport[2] + sl->hReport[3]) * apm;
    evMark(sl, 8, st);
    // ---- 2. sort by linked cell; the leavers sort behind the last cell and are dropped
    a->size = n + nRecv;
    const double cutoff = sl->cfg.rc + sl->cfg.skin;
    const double delta[3] = {cutoff, cutoff, 0.25 * cutoff};  // as md.cu: fine z order inside a cell column
    const mrmd_b200_atoms* lc = a;  // the linked-cell structure of the units
    if (apm > 1)
    {
        // centres of mass of the wrapped residents and of the arrivals (UpdateMolecules::update, as the reference
        // recomputes them before its list build), then LinkedCellList + permute on them, the atoms moving in blocks
        mrmd_b200_molecules* m = sl->mols;
        const int64_t units = (n + nRecv) / apm;
        m->numLocal = units;
        m->size = units;
        MB_TRY(mrmd_b200_molecules_update(m, a, &sl->cfg.weight, st));
        MB_TRY(moleculesCellSortWithAtoms(m, a, units, apm, delta, sl->sub.minCorner, sl->sub.maxCorner, flag, st));
        a->numLocal = n + nRecv - nSend;
        m->numLocal = a->numLocal / apm;
        m->size = m->numLocal;
        m->lcView->numLocal = m->numLocal;
        m->lcView->size = m->numLocal;
        m->lcView->lcEnd = m->numLocal;
        lc = m->lcView;
    }
    else
    {
        if (columnsOnly) MB_TRY(atomsCellSortSlab(a, n + nRecv, delta, &sl->sub, st));
        else MB_TRY(atomsCellSortDrop(a, 0, n + nRecv, delta, sl->sub.minCorner, sl->sub.maxCorner, flag, st));
        a->numLocal = n + nRecv - nSend;
        a->lcEnd = a->numLocal;
    }
    a->size = a->numLocal;
    sl->posEpochAtSort = a->posEpoch;
    sl->presSinceSort = 0;
    MB_TRY(atomsEnsureCapacity(a, a->numLocal + 2 * sl->p2pCap * apm, st));
    MB_TRY(slabEnsureMolecules(sl, a->numLocal / apm + 2 * sl->p2pCap, st));
    if (apm > 1) sl->mols->lcView->v.pos = sl->mols->v.pos;
    sl->profRebuild[1] += profMark(sl, st, last);
    evMark(sl, 9, st);
    // ---- 3. face lists from the first / last cell column (ranges read on the device), push, pull
    const GridDev& g = lc->lcGrid;
    const int64_t perX = int64_t(g.n[1]) * g.n[2];
    {
        const int32_t* cs = lc->lcCellStart.as<int32_t>();
        SelArgs face{};
        face.upos = unitPositions(sl);
        face.rangeA0 = cs;
        face.rangeA1 = cs + perX;
        face.rangeB0 = cs + (int64_t(g.n[0]) - 1) * perX;
        face.rangeB1 = cs + lc->lcNumCells;
        face.lowBound = sl->sub.minInnerCorner[0];   // low face: x < minInner within column 0
        face.highBound = sl->sub.maxInnerCorner[0];  // high face: x >= maxInner within column nx-1
        const int blocksCol = std::max(1, gridFor(sl->colBound, SL_THREADS));
        selCountKernel<1><<<dim3(blocksCol, 2), SL_THREADS, 0, st>>>(face, bc, sl->dErr);
        MB_LAUNCHED();
        selScanKernel<<<2, SL_THREADS, 0, st>>>(bc, blocksCol, sl->dTotals + 2);
        MB_LAUNCHED();
        selIndexKernel<1><<<dim3(blocksCol, 2), SL_THREADS, 0, st>>>(face, bc, sl->idxLow.as<int32_t>(), sl->idxHigh.as<int32_t>(),
                                                                    sl->p2pCap, sl->dErr);
        MB_LAUNCHED();
    }
    evMark(sl, 10, st);
    MB_TRY(haloPush(sl, true, st));
    MB_TRY(haloPull(sl, true, st));
    sl->profRebuild[2] += profMark(sl, st, last);
    evMark(sl, 11, st);
    // ---- 4. linked cells of the received halo units (they arrive in (j, k) order), tiled neighbour build
    MB_TRY(sl->haloKeys.reserve(size_t(2 * sl->p2pCap) * 4));
    MB_TRY(sl->haloStartLeft.reserve(size_t(perX + 1) * 4));
    MB_TRY(sl->haloStartRight.reserve(size_t(perX + 1) * 4));
    haloKeyCountedKernel<<<gridFor(2 * sl->p2pCap, 256), 256, 0, st>>>(unitPositions(sl), a->numLocal / apm, sl->dTotals + 6, g,
                                                                      sl->haloKeys.as<uint32_t>());
    MB_LAUNCHED();
    haloCellStartCountedKernel<<<gridFor(perX + 1, 256), 256, 0, st>>>(sl->haloKeys.as<uint32_t>(), sl->dTotals + 6, perX,
                                                                      a->numLocal / apm, sl->haloStartLeft.as<int32_t>(),
                                                                      sl->haloStartRight.as<int32_t>());
    MB_LAUNCHED();
    evMark(sl, 12, st);
    // host round trip 2 (list statistics)
    if (apm > 1)
        MB_TRY(verletBuildTiledMolecules(sl->list, sl->mols, &sl->sub, cutoff, 1.0, sl->cfg.maxNeighbors, apm,
                                         sl->haloStartLeft.as<int32_t>(), sl->haloStartRight.as<int32_t>(), st));
    else
        MB_TRY(verletBuildTiled(sl->list, a, &sl->sub, cutoff, 1.0, sl->cfg.maxNeighbors, sl->haloStartLeft.as<int32_t>(),
                                sl->haloStartRight.as<int32_t>(), st));
    sl->profRebuild[3] += profMark(sl, st, last);
    evMark(sl, 13, st);
    MB_TRY(pollReport(sl, 12, static_cast<long long>(sl->haloSeq), st, "halo exchange"));
    MB_REQUIRE(sl->hReport[13] == 0, "slab: the position halo exceeds the peer buffer (density more than tripled since slab_create)");
    MB_REQUIRE(sl->hReport[14] == 0, "slab: a selection list exceeded its capacity during the rebuild");
    sl->sendLeftCount = sl->hReport[8];
    sl->sendRightCount = sl->hReport[9];
    sl->haloLeftCount = sl->hReport[10];
    sl->haloRightCount = sl->hReport[11];
    a->numGhost = (sl->haloLeftCount + sl->haloRightCount) * apm;
    a->size = a->numLocal + a->numGhost;
    int64_t total = 0;
    MB_TRY(mrmd_b200_verlet_info(sl->list, nullptr, nullptr, &total, nullptr));
    sl->storedPairsNow = total;
    sl->rebuilds += 1;
    return 0;
}

static int slabRebuild(mrmd_b200_slab* sl, cudaStream_t st)
{
    mrmd_b200_atoms* a = sl->atoms;
    if (sl->p2p) return slabRebuildP2P(sl, st);
    MB_REQUIRE(sl->apm == 1, "slab: multi-atom molecules need the peer-memory data plane (CUDA IPC between the ranks)");
    MB_TRY(migrate(sl, st));
    MB_TRY(haloExchange(sl, st));
    MB_TRY(haloRefresh(sl, st));
    // linked cells of the received halo atoms: they arrive in (j, k) order
    const GridDev& g = a->lcGrid;
    const int64_t perX = int64_t(g.n[1]) * g.n[2];
    const int64_t n = a->numLocal;
    MB_TRY(sl->haloKeys.reserve(size_t(std::max<int64_t>(a->numGhost, 1)) * 4));
    MB_TRY(sl->haloStartLeft.reserve(size_t(perX + 1) * 4));
    MB_TRY(sl->haloStartRight.reserve(size_t(perX + 1) * 4));
    uint32_t* keys = sl->haloKeys.as<uint32_t>();
    if (a->numGhost > 0)
    {
        haloKeyKernel<<<gridFor(a->numGhost, 256), 256, 0, st>>>(a->v.pos, n, a->numGhost, g, keys);
        MB_LAUNCHED();
    }
    MB_TRY(cellStartFromKeys(keys, sl->haloLeftCount, perX, sl->haloStartLeft.as<int32_t>(), static_cast<int32_t>(n), st));
    MB_TRY(cellStartFromKeys(keys + sl->haloLeftCount, sl->haloRightCount, perX, sl->haloStartRight.as<int32_t>(),
                             static_cast<int32_t>(n + sl->haloLeftCount), st));
    const double cutoff = sl->cfg.rc + sl->cfg.skin;
    MB_TRY(verletBuildTiled(sl->list, a, &sl->sub, cutoff, 1.0, sl->cfg.maxNeighbors, sl->haloStartLeft.as<int32_t>(),
                            sl->haloStartRight.as<int32_t>(), st));
    int64_t total = 0;
    MB_TRY(mrmd_b200_verlet_info(sl->list, nullptr, nullptr, &total, nullptr));
    sl->storedPairsNow = total;
    sl->rebuilds += 1;
    return 0;
}

// Peer-memory halo: allocate the IPC buffer, agree on its capacity, swap handles with the two neighbours through the
// communicator and map their buffers.  Any failure (no peer access, IPC unavailable) leaves the NCCL halo in place.
static int setupPeerHalo(mrmd_b200_slab* sl, double cutoff, double width)
{
    // face atoms ~ numLocal * cutoff / width; three times that, the same on every rank
    const double estimate = 3.0 * double(sl->atoms->numLocal / sl->apm) * cutoff / std::max(width, cutoff) + 4096.0;
    sl->hScalars[0] = estimate;
    MB_CUDA(cudaMemcpy(sl->dScalars, sl->hScalars, 8, cudaMemcpyHostToDevice));
    MB_NCCL(g_nccl.allReduce(sl->dScalars, sl->dScalars, 1, ncclDouble, ncclMax, sl->comm, nullptr));
    MB_CUDA(cudaMemcpy(sl->hScalars, sl->dScalars, 8, cudaMemcpyDeviceToHost));
    const int64_t cap = (static_cast<int64_t>(sl->hScalars[0]) + 63) & ~int64_t(63);  // units per halo region
    sl->p2pCap = cap;
    // migrating units per rebuild and face: those within the displacement bound of a face, a small fraction of the face
    // layer; sized like it anyway (the same on every rank)
    sl->migCap = cap / 4 + 4096;
    // units of one cell column (the face selection runs over the first and last column): three times the mean
    sl->colBound = cap + 4096;
    const size_t bytes = size_t(migRegionOffset(sl, 2)) * 32;
    double4* buf = nullptr;
    cudaIpcMemHandle_t mine;
    bool ok = cudaMalloc(&buf, bytes) == cudaSuccess && cudaMemset(buf, 0, bytes) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine, buf) == cudaSuccess;
    // handles travel as bytes through the communicator (every rank takes part, also when its own setup failed)
    const size_t hs = sizeof(cudaIpcMemHandle_t);
    const int R = sl->nranks;
    if (R > SL_MAX_PEERS) ok = false;
    unsigned char* dH = nullptr;
    MB_CUDA(cudaMalloc(&dH, size_t(R + 1) * hs));
    unsigned char okByte = ok ? 1 : 0;
    if (ok) MB_CUDA(cudaMemcpy(dH, &mine, hs, cudaMemcpyHostToDevice));
    MB_NCCL(g_nccl.allGather(dH, dH + hs, hs, ncclUint8, sl->comm, nullptr));
    MB_CUDA(cudaDeviceSynchronize());
    std::vector<cudaIpcMemHandle_t> all(static_cast<size_t>(R));
    MB_CUDA(cudaMemcpy(all.data(), dH + hs, size_t(R) * hs, cudaMemcpyDeviceToHost));
    for (int r = 0; r < R && r < SL_MAX_PEERS; ++r) sl->peers.p[r] = nullptr;
    for (int r = 0; r < R && ok; ++r)
    {
        if (r == sl->rank)
        {
            sl->peers.p[r] = buf;
            continue;
        }
        void* q = nullptr;
        ok = cudaIpcOpenMemHandle(&q, all[static_cast<size_t>(r)], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
        sl->peers.p[r] = static_cast<double4*>(q);
    }
    void* pl = ok ? sl->peers.p[sl->left] : nullptr;
    void* pr = ok ? sl->peers.p[sl->right] : nullptr;
    okByte = ok ? 1 : 0;
    cudaGetLastError();  // a failed IPC call must not poison later launches
    // everybody or nobody: the path is a property of the run, not of a rank
    sl->hScalars[0] = okByte ? 1.0 : 0.0;
    MB_CUDA(cudaMemcpy(sl->dScalars, sl->hScalars, 8, cudaMemcpyHostToDevice));
    MB_NCCL(g_nccl.allReduce(sl->dScalars, sl->dScalars, 1, ncclDouble, ncclMin, sl->comm, nullptr));
    MB_CUDA(cudaMemcpy(sl->hScalars, sl->dScalars, 8, cudaMemcpyDeviceToHost));
    cudaFree(dH);
    sl->p2pBuf = buf;
    sl->peerLeftBuf = static_cast<double4*>(pl);
    sl->peerRightBuf = static_cast<double4*>(pr);
    sl->p2p = sl->hScalars[0] > 0.5;
    if (sl->p2p)
    {
        MB_CUDA(cudaMallocHost(&sl->hReport, 16 * 8));
        std::memset(sl->hReport, 0, 16 * 8);
        MB_CUDA(cudaMalloc(&sl->dErr, 4));
        MB_CUDA(cudaMemset(sl->dErr, 0, 4));
        MB_TRY(sl->migIdxA.reserve(size_t(sl->migCap) * 4));
        MB_TRY(sl->migIdxB.reserve(size_t(sl->migCap) * 4));
        MB_TRY(sl->idxLow.reserve(size_t(sl->p2pCap) * 4));
        MB_TRY(sl->idxHigh.reserve(size_t(sl->p2pCap) * 4));
        MB_CUDA(cudaMalloc(&sl->dPushTicket, 4));
        MB_CUDA(cudaMemset(sl->dPushTicket, 0, 4));
        MB_CUDA(cudaMallocHost(&sl->hDecide, 16));
        sl->hDecide[0] = sl->hDecide[1] = 0.0;
    }
    return 0;
}

// mrmd_b200_adress::preUpdateHook: the compensation-energy samples of all slabs enter the mean
static int sumOverRanks(void* ctx, double* sums, int64_t count, cudaStream_t st)
{
    auto* sl = static_cast<mrmd_b200_slab*>(ctx);
    MB_NCCL(g_nccl.allReduce(sums, sums, size_t(count), ncclDouble, ncclSum, sl->comm, st));
    return 0;
}

static int slabAfterDecision(mrmd_b200_slab* sl, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1, bool wantEnergy,
                             bool deferPost, cudaEvent_t evPosReady, bool rebuildNow, bool pushed, const int* stop, double last);

static int slabStep(mrmd_b200_slab* sl, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1, bool wantEnergy,
                    bool deferPost = true, cudaEvent_t evPosReady = nullptr)
{
    const mrmd_b200_md_config& c = sl->cfg;
    mrmd_b200_atoms* a = sl->atoms;
    double last = 0.0;
    profMark(sl, st, last);
    evMark(sl, 0, st);
    if (sl->constraints != nullptr)  // tests/Constraints/Constraints.cpp:53-54
        MB_TRY(constraintsEnforcePositional(sl->constraints, sl->mols, a, c.dt, st));
    // the previous step's postForceIntegrate rides in front of this kick (flushed when a run returns)
    MB_TRY(integratePre(a, c.dt, c.integrator == 1, c.zeta, c.temperature, c.seed,
                        uint64_t(sl->step), nullptr, sl->postPending, st));
    sl->postPending = false;
    sl->presSinceSort += 1;
    sl->prof[0] += profMark(sl, st, last);
    // the positions are final: the face atoms leave for the neighbours right away, so that the NVLink transfer overlaps
    // the rebuild decision (double-buffered regions; a push that a rebuild supersedes is simply never pulled).  Before
    // the first rebuild there are no lists yet.
    const bool earlyPush = sl->p2p && sl->rebuilds > 0;
    if (sl->apm > 1 && sl->rebuilds > 0)  // the centres of mass travel with the face molecules
        MB_TRY(mrmd_b200_molecules_update(sl->mols, a, &c.weight, st));
    evMark(sl, 1, st);
    if (earlyPush)
    {
        if (sl->sPush == nullptr)
        {
            int prioLow = 0, prioHigh = 0;
            MB_CUDA(cudaDeviceGetStreamPriorityRange(&prioLow, &prioHigh));
            MB_CUDA(cudaStreamCreateWithPriority(&sl->sPush, cudaStreamNonBlocking, prioHigh));
            MB_CUDA(cudaEventCreateWithFlags(&sl->evPreDone, cudaEventDisableTiming));
            MB_CUDA(cudaEventCreateWithFlags(&sl->evPushDone, cudaEventDisableTiming));
        }
        MB_CUDA(cudaEventRecord(sl->evPreDone, st));
        MB_CUDA(cudaStreamWaitEvent(sl->sPush, sl->evPreDone, 0));
        MB_TRY(haloPush(sl, false, sl->sPush));
        MB_CUDA(cudaEventRecord(sl->evPushDone, sl->sPush));
    }
    evMark(sl, 2, st);
    // the rebuild decision is collective: global maximum of the squared displacement
    if (sl->p2p)
    {
        sl->decideSeq += 1.0;
        const int parity = static_cast<int>(static_cast<long long>(sl->decideSeq) & 1);
        maxDisplacementGatherKernel<<<1, 32, 0, st>>>(a->dMaxDisp, sl->peers, sl->rank, sl->nranks, sl->decideSeq, parity,
                                                      a->dMaxDisp, sl->hDecide, nullptr, 0.0, nullptr, 0);
        MB_LAUNCHED();
        evMark(sl, 3, st);
        // the kernel writes {max, step} into pinned memory: poll it instead of synchronising the stream
        volatile double* h = sl->hDecide;
        for (unsigned spin = 0; h[1] != sl->decideSeq; ++spin)
        {
            if ((spin & 0xfff) == 0xfff)
            {
                const cudaError_t q = cudaStreamQuery(st);
                if (q != cudaSuccess && q != cudaErrorNotReady) MB_CUDA(q);
                if (q == cudaSuccess && h[1] != sl->decideSeq) MB_REQUIRE(false, "slab: displacement gather finished without a result");
            }
        }
        *a->hMaxDisp = h[0];
    }
    else
    {
        MB_NCCL(g_nccl.allReduce(a->dMaxDisp, a->dMaxDisp, 1, ncclDouble, ncclMax, sl->comm, st));
        MB_CUDA(cudaMemcpyAsync(a->hMaxDisp, a->dMaxDisp, 8, cudaMemcpyDeviceToHost, st));
        MB_CUDA(cudaStreamSynchronize(st));
    }
    // whatever follows (halo pull, or a rebuild that re-orders the atoms and pushes again) is ordered behind the push
    if (earlyPush) MB_CUDA(cudaStreamWaitEvent(st, sl->evPushDone, 0));
    MB_REQUIRE(std::isfinite(*a->hMaxDisp), "slab_run: non-finite position, velocity or force (the system blew up)");
    sl->maxDisplacement += std::sqrt(*a->hMaxDisp);
    sl->prof[1] += profMark(sl, st, last);
    return slabAfterDecision(sl, st, e0, e1, wantEnergy, deferPost, evPosReady, sl->maxDisplacement >= c.skin * 0.5, earlyPush,
                             nullptr, last);
}

// the step behind the rebuild decision: rebuild or halo pull, force, postForceIntegrate.  stop (device, optional): the step
// is queued ahead of the host, its kernels return at once when an earlier queued step asked for a rebuild.
static int slabAfterDecision(mrmd_b200_slab* sl, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1, bool wantEnergy,
                             bool deferPost, cudaEvent_t evPosReady, bool rebuildNow, bool pushed, const int* stop, double last)
{
    const mrmd_b200_md_config& c = sl->cfg;
    mrmd_b200_atoms* a = sl->atoms;
    if (rebuildNow)
    {
        sl->maxDisplacement = 0.0;
        if (sl->stepsSinceRebuild > 0) sl->lastRebuildInterval = sl->stepsSinceRebuild;
        sl->stepsSinceRebuild = 0;
        MB_TRY(slabRebuild(sl, st));
        sl->prof[2] += profMark(sl, st, last);
    }
    else
    {
        if (pushed) MB_TRY(haloPull(sl, false, st, stop));
        else MB_TRY(haloRefresh(sl, st));
        sl->prof[3] += profMark(sl, st, last);
        evMark(sl, 4, st);
    }
    if (evPosReady != nullptr) MB_CUDA(cudaEventRecord(evPosReady, st));  // positions and atom order are final
    if (sl->apm > 1)
    {
        // AdResS on molecules of apm atoms: LJ_IdealGas + ContributeMoleculeForceToAtoms as one kernel over the tiled
        // list of the centres of mass (local + halo molecules), SHAKE / RATTLE stay rank-local (whole molecules migrate)
        for (int d = 0; d < 3; ++d) MB_CUDA(cudaMemsetAsync(a->v.force[d], 0, size_t(a->numLocal) * 8, st));
        if (e0) MB_CUDA(cudaEventRecord(e0, st));
        MB_TRY(adressRunPeriodicMolecules(sl->adress, sl->mols, a, sl->list, &c.weight, sl->apm, wantEnergy, st));
        if (e1) MB_CUDA(cudaEventRecord(e1, st));
    }
    else if (c.adress)
    {
        // SURVEY.md section 3.5 on a slab: thermodynamic force (density histogram all-reduced over the ranks
        // before an update, table replicated), then the tiled AdResS kernel over local + halo atoms
        for (int d = 0; d < 3; ++d) MB_CUDA(cudaMemsetAsync(a->v.force[d], 0, size_t(a->numLocal) * 8, st));
        if (sl->thermo != nullptr)
        {
            mrmd_b200_thermo* t = sl->thermo;
            if (c.thermoSampleInterval > 0 && sl->step % c.thermoSampleInterval == 0)
                MB_TRY(mrmd_b200_thermo_sample(t, a, st));
            if (c.thermoUpdateInterval > 0 && sl->step > 0 && sl->step % c.thermoUpdateInterval == 0 && t->samples > 0)
            {
                MB_NCCL(g_nccl.allReduce(t->density, t->density, size_t(t->numBins * t->numTypes), ncclDouble, ncclSum,
                                         sl->comm, st));
                MB_TRY(mrmd_b200_thermo_update(t, c.thermoSmoothingSigma, c.thermoSmoothingIntensity, nullptr, st));
            }
            MB_TRY(mrmd_b200_thermo_apply(t, a, nullptr, 0, st));
        }
        if (e0) MB_CUDA(cudaEventRecord(e0, st));
        MB_TRY(adressRunPeriodic(sl->adress, a, sl->list, &c.weight, wantEnergy, st, stop));
        if (e1) MB_CUDA(cudaEventRecord(e1, st));
    }
    else
    {
        if (e0) MB_CUDA(cudaEventRecord(e0, st));
        MB_TRY(ljApplyTiled(sl->lj, a, sl->list, false, wantEnergy, st, stop));
        if (e1) MB_CUDA(cudaEventRecord(e1, st));
    }
    sl->prof[4] += profMark(sl, st, last);
    sl->prof[5] += 1.0;
    evMark(sl, 5, st);
    if (sl->constraints != nullptr)
        MB_TRY(constraintsEnforceVelocity(sl->constraints, sl->mols, a, st, c.dt));  // RATTLE with the kick riding along
    else if (deferPost) sl->postPending = true;  // rides in front of the next step's kick
    else MB_TRY(mrmd_b200_vv_post(a, c.dt, st));
    sl->step += 1;
    sl->stepsSinceRebuild += 1;
    return 0;
}

// Steps queued ahead of the host (as md.cu:runQueued, collective): up to `count` steps are enqueued back to back -- kick /
// drift, halo push, displacement gather WITH the rebuild criterion evaluated on the device, halo pull, force -- and the
// step whose accumulated global displacement reaches skin / 2 raises a flag on every rank (all ranks see the same
// maxima), after which the kernels queued behind it return at once.  One host synchronisation per chunk instead of one
// host poll per step: the ranks no longer wait for each other's hosts.
static bool slabCanQueue(const mrmd_b200_slab* sl, int64_t ahead)
{
    const mrmd_b200_md_config& c = sl->cfg;
    if (!sl->p2p || sl->apm > 1 || sl->rebuilds == 0 || sl->profile || sl->evProfile) return false;
    // measured on 2 GPUs (1M atoms each): 0.489 ms / step queued against 0.482 with the per-step host poll -- the host
    // round trip is not what the slab step waits for, and the kernels queued behind a stop cost a little: opt-in
    if (std::getenv("MRMD_B200_SLAB_QUEUED_STEPS") == nullptr) return false;
    const int64_t step = sl->step + ahead;
    if (c.adress)
    {
        if (sl->thermo != nullptr)
        {
            if (c.thermoSampleInterval > 0 && step % c.thermoSampleInterval == 0) return false;
            if (c.thermoUpdateInterval > 0 && step > 0 && step % c.thermoUpdateInterval == 0) return false;
        }
        const int64_t run = sl->adress->runCounter + ahead;
        if (run % sl->adress->samplingInterval == 0 || run % sl->adress->updateInterval == 0) return false;
    }
    return true;
}

static int slabRunQueued(mrmd_b200_slab* sl, int64_t count, cudaStream_t st, cudaEvent_t* events, bool energyOnLast,
                         bool energyAlways, int64_t* done, bool* stopped)
{
    const mrmd_b200_md_config& c = sl->cfg;
    mrmd_b200_atoms* a = sl->atoms;
    if (sl->dStop == nullptr)
    {
        MB_CUDA(cudaMalloc(&sl->dStop, 16));
        MB_CUDA(cudaMalloc(&sl->dAccum, 8));
        MB_CUDA(cudaMallocHost(&sl->hStop, 32));
    }
    double* hAccum = reinterpret_cast<double*>(sl->hStop + 4);
    *hAccum = sl->maxDisplacement;
    MB_CUDA(cudaMemsetAsync(sl->dStop, 0, 16, st));
    MB_CUDA(cudaMemcpyAsync(sl->dAccum, hAccum, 8, cudaMemcpyHostToDevice, st));
    const int64_t step0 = sl->step, since0 = sl->stepsSinceRebuild, run0 = c.adress ? sl->adress->runCounter : 0;
    const unsigned long long haloSeq0 = sl->haloSeq;
    const double decideSeq0 = sl->decideSeq;
    for (int64_t k = 0; k < count; ++k)
    {
        MB_TRY(integratePre(a, c.dt, c.integrator == 1, c.zeta, c.temperature, c.seed, uint64_t(sl->step), nullptr,
                            sl->postPending, st, sl->dStop));
        sl->postPending = false;
        sl->presSinceSort += 1;
        MB_TRY(haloPush(sl, false, st, sl->dStop));
        sl->decideSeq += 1.0;
        const int parity = static_cast<int>(static_cast<long long>(sl->decideSeq) & 1);
        maxDisplacementGatherKernel<<<1, 32, 0, st>>>(a->dMaxDisp, sl->peers, sl->rank, sl->nranks, sl->decideSeq, parity,
                                                      a->dMaxDisp, sl->hDecide, sl->dAccum, c.skin * 0.5, sl->dStop,
                                                      static_cast<int>(k));
        MB_LAUNCHED();
        const bool wantEnergy = energyAlways || (energyOnLast && k == count - 1);
        MB_TRY(slabAfterDecision(sl, st, events ? events[2 * k] : nullptr, events ? events[2 * k + 1] : nullptr, wantEnergy, true,
                                 nullptr, false, true, sl->dStop, 0.0));
    }
    MB_CUDA(cudaMemcpyAsync(sl->hStop, sl->dStop, 16, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaMemcpyAsync(hAccum, sl->dAccum, 8, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    MB_REQUIRE(sl->hStop[2] == 0, "slab_run: non-finite position, velocity or force (the system blew up)");
    *stopped = sl->hStop[0] != 0;
    *done = *stopped ? sl->hStop[1] : count;
    sl->maxDisplacement = *hAccum;
    // the host-side counters follow the steps that really ran: a stopped step ran its kick / drift, push and gather
    const int64_t ran = *done + (*stopped ? 1 : 0);
    sl->haloSeq = haloSeq0 + static_cast<unsigned long long>(ran);
    sl->decideSeq = decideSeq0 + double(ran);
    sl->step = step0 + *done;
    sl->stepsSinceRebuild = since0 + *done;
    if (c.adress) sl->adress->runCounter = run0 + *done;
    sl->postPending = !*stopped;
    return 0;
}
}  // namespace mrmd_b200

using namespace mrmd_b200;

extern "C" {

int mrmd_b200_nccl_unique_id(void* out128)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(out128 != nullptr, "nccl_unique_id");
    MB_TRY(loadNccl());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    MB_NCCL(g_nccl.getUniqueId(&id));
    std::memcpy(out128, &id, 128);
    return 0;
}

int mrmd_b200_slab_create(mrmd_b200_slab** out, const mrmd_b200_md_config* cfg, const double* globalMin,
                          const double* globalMax, int rank, int nranks, const void* uniqueId128, mrmd_b200_atoms* atoms,
                          void* stream)
{
    return mrmd_b200_slab_create_cuts(out, cfg, globalMin, globalMax, nullptr, rank, nranks, uniqueId128, atoms, stream);
}

int mrmd_b200_slab_create_cuts(mrmd_b200_slab** out, const mrmd_b200_md_config* cfg, const double* globalMin,
                               const double* globalMax, const double* cuts, int rank, int nranks,
                               const void* uniqueId128, mrmd_b200_atoms* atoms, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(out && cfg && globalMin && globalMax && uniqueId128 && atoms, "slab_create");
    MB_REQUIRE(nranks >= 2 && rank >= 0 && rank < nranks, "slab_create: needs at least two ranks");
    const int64_t apmCfg = std::max<int64_t>(cfg->atomsPerMolecule, 1);
    MB_REQUIRE(apmCfg == 1 || (apmCfg == 4 && cfg->adress && !cfg->useThermoForce),
               "slab_create: molecules of one or four atoms (four: AdResS without thermodynamic force)");
    MB_REQUIRE(atoms->numLocal % apmCfg == 0, "slab_create: local atoms are not a multiple of atomsPerMolecule");
    MB_REQUIRE(cfg->numConstraintIterations >= 0 && (cfg->numConstraintIterations == 0 || (apmCfg > 1 && cfg->bondLength > 0.0)),
               "slab_create: constraints need multi-atom molecules and a positive bond length");
    if (cfg->adress)
    {
        // the tiled AdResS kernel evaluates lambda at image positions: across the global periodic x faces (and y, z
        // for a spherical region, checked per launch) the AT + HY region must stay a list radius away
        const mrmd_b200_weight& w = cfg->weight;
        const double reach = (w.kind == MRMD_B200_WEIGHT_SLAB) ? 0.5 * w.atRegion + w.hyRegion : w.atRegion + w.hyRegion;
        const double cut = cfg->rc + cfg->skin;
        MB_REQUIRE(w.center[0] - reach >= globalMin[0] + cut && w.center[0] + reach <= globalMax[0] - cut,
                   "slab_create: the AT/HY region reaches the periodic x boundary of the global box");
    }
    MB_TRY(loadNccl());
    auto* sl = new mrmd_b200_slab;
    sl->cfg = *cfg;
    sl->rank = rank;
    sl->nranks = nranks;
    sl->left = (rank - 1 + nranks) % nranks;
    sl->right = (rank + 1) % nranks;
    const double cutoff = cfg->rc + cfg->skin;
    const double lx = globalMax[0] - globalMin[0];
    // equal-width slabs, or the caller's cuts (cost-balanced slabs: narrow over the AT / HY region, wide over CG)
    if (cuts != nullptr)
    {
        MB_REQUIRE(cuts[0] == globalMin[0] && cuts[nranks] == globalMax[0], "slab_create: cuts must span the global box");
        for (int r = 0; r < nranks; ++r) MB_REQUIRE(cuts[r + 1] - cuts[r] >= cutoff, "slab_create: slab narrower than rc + skin");
    }
    const double equalWidth = lx / nranks;
    const double lo = cuts ? cuts[rank] : globalMin[0] + rank * equalWidth;
    const double hi = cuts ? cuts[rank + 1] : ((rank == nranks - 1) ? globalMax[0] : globalMin[0] + (rank + 1) * equalWidth);
    const double width = hi - lo;
    double mn[3] = {lo, globalMin[1], globalMin[2]};
    double mx[3] = {hi, globalMax[1], globalMax[2]};
    const double th[3] = {cutoff, cutoff, cutoff};
    mrmd_b200_subdomain_init(&sl->sub, mn, mx, th);
    for (int d = 0; d < 3; ++d)
    {
        sl->globalMin[d] = globalMin[d];
        sl->globalMax[d] = globalMax[d];
    }
    sl->shiftToLeft = (rank == 0) ? lx : 0.0;             // crossing the low end of the box: x + Lx
    sl->shiftToRight = (rank == nranks - 1) ? -lx : 0.0;  // crossing the high end: x - Lx
    sl->atoms = atoms;
    sl->apm = static_cast<int>(apmCfg);
    sl->profile = std::getenv("MRMD_B200_SLAB_PROFILE") != nullptr;
    sl->evProfile = std::getenv("MRMD_B200_SLAB_EVENTS") != nullptr;
    int rc = 0;
    if (width < cutoff)
    {
        setLastError("slab_create: slab narrower than rc + skin");
        rc = MRMD_B200_EINVAL;
    }
    ncclUniqueId id;
    std::memcpy(&id, uniqueId128, 128);
    if (rc == 0 && g_nccl.commInitRank(&sl->comm, nranks, id, rank) != ncclSuccess)
    {
        setLastError("slab_create: ncclCommInitRank failed");
        rc = MRMD_B200_EINVAL;
    }
    if (rc == 0) rc = mrmd_b200_verlet_create(&sl->list, 0);
    if (rc == 0 && cfg->adress)
    {
        // tiles in the coarse-grained region hold ideal-gas pairs only: the tiled build leaves their rows empty
        sl->list->tiledCgSkip = true;
        sl->list->tiledCgWeight = cfg->weight;
    }
    if (rc == 0) rc = mrmd_b200_lj_create(&sl->lj, &cfg->cappingDistance, &cfg->rc, &cfg->sigma, &cfg->epsilon, 1, 0);
    if (rc == 0 && cfg->adress)
    {
        rc = mrmd_b200_adress_create(&sl->adress, &cfg->cappingDistance, &cfg->rc, &cfg->sigma, &cfg->epsilon, 1,
                                     cfg->doShift);
        if (rc == 0)
        {
            sl->adress->preUpdateHook = sumOverRanks;
            sl->adress->hookCtx = sl;
        }
        if (rc == 0 && cfg->useThermoForce)
        {
            mrmd_b200_subdomain global;
            mrmd_b200_subdomain_init(&global, globalMin, globalMax, th);
            rc = mrmd_b200_thermo_create(&sl->thermo, &cfg->thermoTargetDensity, 1, &global, cfg->thermoBinWidth,
                                         &cfg->thermoModulation, 0, 0);
        }
    }
    if (rc == 0 && sl->apm > 1)
    {
        const int64_t numMols = atoms->numLocal / sl->apm;
        rc = mrmd_b200_molecules_create(&sl->mols, std::max<int64_t>(numMols, 1));
        if (rc == 0)
        {
            sl->mols->size = numMols;
            sl->mols->numLocal = numMols;
            slabMoleculeInitKernel<<<gridFor(sl->mols->capacity, 256), 256>>>(sl->mols->v, sl->mols->capacity, sl->apm);
            g_launchCount.fetch_add(1);
            if (cudaDeviceSynchronize() != cudaSuccess) rc = MRMD_B200_EINVAL;
        }
        if (rc == 0 && cfg->numConstraintIterations > 0)
        {
            // MoleculeConstraints(apm, iterations) with a bond between every two atoms of a molecule (as md.cu)
            rc = mrmd_b200_constraints_create(&sl->constraints, sl->apm, cfg->numConstraintIterations);
            std::vector<int64_t> bi, bj;
            std::vector<double> eq;
            for (int64_t i = 0; i < sl->apm; ++i)
                for (int64_t j = i + 1; j < sl->apm; ++j)
                {
                    bi.push_back(i);
                    bj.push_back(j);
                    eq.push_back(cfg->bondLength);
                }
            if (rc == 0) rc = mrmd_b200_constraints_set(sl->constraints, bi.data(), bj.data(), eq.data(), int64_t(eq.size()));
            if (rc == 0) constraintsSetUniformMolecules(sl->constraints, true);
        }
        if (rc == 0 && std::getenv("MRMD_B200_SLAB_NO_P2P") != nullptr)
        {
            setLastError("slab_create: multi-atom molecules need the peer-memory data plane");
            rc = MRMD_B200_EINVAL;
        }
    }
    if (rc == 0 && cudaMalloc(&sl->dTotals, 64) != cudaSuccess) rc = MRMD_B200_ENOMEM;
    if (rc == 0 && cudaMallocHost(&sl->hTotals, 64) != cudaSuccess) rc = MRMD_B200_ENOMEM;
    if (rc == 0 && cudaMalloc(&sl->dScalars, 64) != cudaSuccess) rc = MRMD_B200_ENOMEM;
    if (rc == 0 && cudaMallocHost(&sl->hScalars, 64) != cudaSuccess) rc = MRMD_B200_ENOMEM;
    if (rc == 0 && std::getenv("MRMD_B200_SLAB_NO_P2P") == nullptr) rc = setupPeerHalo(sl, cutoff, width);
    if (rc != 0)
    {
        mrmd_b200_slab_destroy(sl);
        return rc;
    }
    (void)stream;
    *out = sl;
    return 0;
}

int mrmd_b200_slab_destroy(mrmd_b200_slab* sl)
{
    if (sl == nullptr) return 0;
    cudaDeviceSynchronize();
    evReport(sl);
    if (sl->profile && sl->prof[5] > 0)
        std::fprintf(stderr,
                     "[mrmd_b200 slab rank %d] us/step over %.0f steps (%lld rebuilds): pre %.1f, decision %.1f, rebuild %.1f, "
                     "halo refresh %.1f, force %.1f\n",
                     sl->rank, sl->prof[5], static_cast<long long>(sl->rebuilds), sl->prof[0] / sl->prof[5],
                     sl->prof[1] / sl->prof[5], sl->prof[2] / sl->prof[5], sl->prof[3] / sl->prof[5], sl->prof[4] / sl->prof[5]);
    if (sl->profile && sl->rebuilds > 0)
        std::fprintf(stderr, "[mrmd_b200 slab rank %d] us per rebuild: migration %.1f, sort %.1f, face lists + halo %.1f, list build %.1f\n",
                     sl->rank, sl->profRebuild[0] / sl->rebuilds, sl->profRebuild[1] / sl->rebuilds,
                     sl->profRebuild[2] / sl->rebuilds, sl->profRebuild[3] / sl->rebuilds);
    for (auto e : sl->events) cudaEventDestroy(e);
    sl->hp.destroy();
    if (sl->sPush != nullptr) cudaStreamDestroy(sl->sPush);
    if (sl->evPreDone != nullptr) cudaEventDestroy(sl->evPreDone);
    if (sl->evPushDone != nullptr) cudaEventDestroy(sl->evPushDone);
    if (sl->comm != nullptr) g_nccl.commDestroy(sl->comm);
    mrmd_b200_verlet_destroy(sl->list);
    for (int r = 0; r < SL_MAX_PEERS && r < sl->nranks; ++r)
        if (sl->peers.p[r] != nullptr && r != sl->rank) cudaIpcCloseMemHandle(sl->peers.p[r]);
    if (sl->hDecide != nullptr) cudaFreeHost(sl->hDecide);
    if (sl->dStop != nullptr) cudaFree(sl->dStop);
    if (sl->dAccum != nullptr) cudaFree(sl->dAccum);
    if (sl->hStop != nullptr) cudaFreeHost(sl->hStop);
    if (sl->hReport != nullptr) cudaFreeHost(sl->hReport);
    if (sl->dErr != nullptr) cudaFree(sl->dErr);
    sl->migIdxA.release();
    sl->migIdxB.release();
    if (sl->p2pBuf != nullptr) cudaFree(sl->p2pBuf);
    if (sl->dPushTicket != nullptr) cudaFree(sl->dPushTicket);
    mrmd_b200_lj_destroy(sl->lj);
    mrmd_b200_adress_destroy(sl->adress);
    mrmd_b200_thermo_destroy(sl->thermo);
    mrmd_b200_constraints_destroy(sl->constraints);
    mrmd_b200_molecules_destroy(sl->mols);
    if (sl->dTotals) cudaFree(sl->dTotals);
    if (sl->hTotals) cudaFreeHost(sl->hTotals);
    if (sl->dScalars) cudaFree(sl->dScalars);
    if (sl->hScalars) cudaFreeHost(sl->hScalars);
    for (mrmd_b200::DevBuf* b : {&sl->flags, &sl->blockCounts, &sl->idxLow, &sl->idxHigh, &sl->sendBuf, &sl->recvBuf,
                                 &sl->haloKeys, &sl->haloStartLeft, &sl->haloStartRight})
        b->release();
    delete sl;
    return 0;
}

// every rank: nsteps collective steps through this rank's HOST buffers (hostpipe.cuh).  The resident atoms change at
// a rebuild (migration, re-sort): the buffers hold numLocal rows in the rank's current atom order, scalarsHost[3]
// reports the count after each step.
int mrmd_b200_slab_run_host(mrmd_b200_slab* sl, int64_t nsteps, double* posHost, double* velHost, double* scalarsHost,
                            mrmd_b200_md_stats* stats, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(sl != nullptr && nsteps >= 0 && posHost != nullptr && velHost != nullptr, "slab_run_host");
    cudaStream_t st = S(stream);
    if (sl->postPending)
    {
        MB_TRY(mrmd_b200_vv_post(sl->atoms, sl->cfg.dt, st));
        sl->postPending = false;
    }
    const int64_t rebuilds0 = sl->rebuilds;
    const double* dRes = sl->cfg.adress ? sl->adress->dResult : sl->lj->dResult;
    int64_t storedSum = 0;
    MB_TRY(hostPipeRun(sl->hp, sl->atoms, nsteps, posHost, velHost, scalarsHost, dRes, &sl->maxDisplacement, true, st,
                       [&](cudaEvent_t evPosReady) -> int
                       {
                           MB_TRY(slabStep(sl, st, nullptr, nullptr, true, false, evPosReady));
                           storedSum += sl->storedPairsNow;
                           return 0;
                       }));
    MB_CUDA(cudaStreamSynchronize(st));
    if (stats != nullptr)
    {
        *stats = mrmd_b200_md_stats{};
        stats->steps = nsteps;
        stats->rebuilds = sl->rebuilds - rebuilds0;
        stats->storedPairs = storedSum;
        stats->numLocal = sl->atoms->numLocal;
        stats->numGhost = sl->atoms->numGhost;
        stats->maxDisplacement = sl->maxDisplacement;
    }
    return 0;
}

int mrmd_b200_slab_set_energy_every_step(mrmd_b200_slab* sl, int enabled)
{
    MB_REQUIRE(sl != nullptr, "slab_set_energy_every_step");
    sl->cfg.energyEveryStep = enabled ? 1 : 0;
    return 0;
}

int mrmd_b200_slab_run(mrmd_b200_slab* sl, int64_t nsteps, int timeForceKernel, mrmd_b200_md_stats* stats, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(sl != nullptr && nsteps >= 0, "slab_run");
    cudaStream_t st = S(stream);
    const int64_t rebuilds0 = sl->rebuilds;
    const bool adress = sl->cfg.adress != 0;
    double* dRes = adress ? sl->adress->dResult : sl->lj->dResult;
    double* hRes = adress ? sl->adress->hResult : sl->lj->hResult;
    MB_CUDA(cudaMemcpyAsync(hRes, dRes, 48, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    const double pairs0 = adress ? hRes[4] : hRes[5];
    const double active0 = adress ? hRes[5] : 0.0;
    const int nTimed = timeForceKernel ? static_cast<int>(std::min<int64_t>(nsteps, 1 << 16)) : 0;
    while (static_cast<int>(sl->events.size()) < 2 * nTimed)
    {
        cudaEvent_t e;
        MB_CUDA(cudaEventCreate(&e));
        sl->events.push_back(e);
    }
    int64_t storedSum = 0;
    const bool energyAlways = sl->cfg.energyEveryStep != 0;
    for (int64_t i = 0; i < nsteps;)
    {
        // as many steps as the last rebuild interval suggests are queued ahead of the host (see slabRunQueued); every
        // rank takes the same decisions from the same counters
        int64_t count = 0;
        const int64_t want = std::max<int64_t>(1, std::min<int64_t>(16, sl->lastRebuildInterval - sl->stepsSinceRebuild + 1));
        while (count < want && i + count < nsteps && slabCanQueue(sl, count)) ++count;
        if (count >= 1)
        {
            int64_t done = 0;
            bool stopped = false;
            MB_TRY(slabRunQueued(sl, count, st, (i + count <= nTimed) ? sl->events.data() + 2 * i : nullptr, i + count == nsteps,
                                 energyAlways, &done, &stopped));
            storedSum += sl->storedPairsNow * done;
            i += done;
            if (stopped)
            {
                // step i ran kick / drift, push and gather on the device and reached skin / 2: rebuild, force, kick
                MB_TRY(slabAfterDecision(sl, st, i < nTimed ? sl->events[2 * i] : nullptr, i < nTimed ? sl->events[2 * i + 1] : nullptr,
                                         i == nsteps - 1 || energyAlways, true, nullptr, true, true, nullptr, 0.0));
                storedSum += sl->storedPairsNow;
                i += 1;
            }
            continue;
        }
        MB_TRY(slabStep(sl, st, i < nTimed ? sl->events[2 * i] : nullptr, i < nTimed ? sl->events[2 * i + 1] : nullptr,
                        i == nsteps - 1 || energyAlways));
        storedSum += sl->storedPairsNow;
        ++i;
    }
    if (sl->postPending)
    {
        MB_TRY(mrmd_b200_vv_post(sl->atoms, sl->cfg.dt, st));
        sl->postPending = false;
    }
    MB_CUDA(cudaMemcpyAsync(hRes, dRes, 48, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    if (stats != nullptr)
    {
        // energy and virial of the last step and the pair counters summed over the ranks (a pair across a slab
        // face counts one half on either side, so only the global sums are whole numbers)
        sl->hScalars[0] = hRes[0];
        sl->hScalars[1] = adress ? 0.0 : hRes[1];
        sl->hScalars[2] = (adress ? hRes[4] : hRes[5]) - pairs0;
        sl->hScalars[3] = adress ? hRes[5] - active0 : 0.0;
        MB_CUDA(cudaMemcpyAsync(sl->dScalars, sl->hScalars, 32, cudaMemcpyHostToDevice, st));
        MB_NCCL(g_nccl.allReduce(sl->dScalars, sl->dScalars, 4, ncclDouble, ncclSum, sl->comm, st));
        MB_CUDA(cudaMemcpyAsync(sl->hScalars, sl->dScalars, 32, cudaMemcpyDeviceToHost, st));
        MB_CUDA(cudaStreamSynchronize(st));
        stats->steps = nsteps;
        stats->rebuilds = sl->rebuilds - rebuilds0;
        stats->storedPairs = storedSum;
        stats->numLocal = sl->atoms->numLocal;
        stats->numGhost = sl->atoms->numGhost;
        stats->energy = sl->hScalars[0];
        stats->virial = sl->hScalars[1];
        stats->pairInteractions = static_cast<int64_t>(sl->hScalars[2] + 0.5);
        stats->activePairs = adress ? static_cast<int64_t>(hRes[5] - active0 + 0.5) : 0;  // this rank's rows
        stats->maxDisplacement = sl->maxDisplacement;
        double ms = 0.0;
        for (int i = 0; i < nTimed; ++i)
        {
            float t = 0.f;
            MB_CUDA(cudaEventElapsedTime(&t, sl->events[2 * i], sl->events[2 * i + 1]));
            ms += t;
        }
        stats->forceKernelMs = ms;
    }
    return 0;
}

}  // extern "C"
