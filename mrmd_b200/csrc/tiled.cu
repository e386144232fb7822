// tiled.cu -- shared-memory tiled neighbour build and Lennard-Jones force for cell-sorted, periodic systems:
// the B200 fast path behind Cabana::VerletList::build + LennardJones::apply.
//
// Why: a thread-per-atom walk over a Verlet list issues one divergent 32-byte gather per pair.  ncu
// (profiles/r01_lj_generic_*.txt) shows the L1 address stage saturating at ~2.3 cycles per lane-gather (and
// ~1.3 cycles per lane for the fp64 RED scatter of the half list): the list kernels are LSU-wavefront bound,
// far from HBM.  Here a block owns a tile = one (x,y) cell column x CH cells along z of the linked-cell grid
// left behind by LinkedCellList + permute.  Atoms are stored in cell order (z fastest), so the 3x3 neighbour
// columns are nine contiguous index ranges: they are copied once, coalesced, into shared memory as 24-byte
// {x, y, z} records (periodic images are produced on the fly by adding +-L exactly as GhostExchange /
// UpdateGhostAtoms do, so the staged coordinates are bit-identical to the ghost atoms' coordinates) and every
// neighbour of a home atom becomes a 16-bit shared-memory slot.
// Force kernels: TL_GROUP lanes share one home atom (measured on B200 for 8 / 4 / 2 / 1 lanes in round 1: 225 / 203 / 200 /
// 308 us per 1M atoms; fewer lanes amortise the per-home work over more homes per warp until shared-memory conflicts and
// row-length imbalance take over): each lane owns every TL_GROUP-th list entry, stored contiguously (tiledRowIndex) so
// that its entries arrive with 16-byte loads a pass ahead; the force components are combined with warp shuffles.
// Build kernel: one lane per home atom, groups of TL_BUILD_GROUP consecutive homes sweep the same candidates in lock step
// (see verletBuildTiledKernel).
// No atomics (full list: an atom accumulates only its own force), no ghost refresh / fold-back in the step,
// 2 bytes of list traffic per pair.
// Tried and dropped (measured slower): dealing the staged slots round-robin over all threads instead of one piece
// per warp; persistent blocks that prefetch the next tile's raw records with cp.async while computing (the extra
// 32 B/slot of shared memory costs a resident block: 230 vs 194 us); tiles of 80 or 145 instead of 110 home atoms;
// a single-precision candidate filter in the build (12-byte records, double-precision re-check inside a margin around
// r^2): pair sets stayed bit exact, no gain; several pairs per lane evaluated phase by phase in the force kernel
// (registers cost a resident block).  Round-2 measurements: profiles/r02_build_experiments.md.
//
// Pair set: identical to the reference's list over local + ghost atoms (same criterion, same uncontracted
// distance arithmetic, Cabana's stencil pruning re-checked on accepted pairs): tests decode the slots back to
// (local partner, image shift) and compare with the oracle's list mapped through correspondingRealAtom.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "handles.cuh"
#include "weight.cuh"

namespace mrmd_b200
{
// measurement knobs (profiles/variants.py builds side-by-side libraries with -D overrides); the defaults are the product
#ifndef MRMD_TL_THREADS_BUILD
#define MRMD_TL_THREADS_BUILD 128
#endif
#ifndef MRMD_TL_THREADS_FORCE
#define MRMD_TL_THREADS_FORCE 128
#endif
constexpr int TL_THREADS_BUILD = MRMD_TL_THREADS_BUILD;  // neighbour build / decode
constexpr int TL_THREADS_FORCE = MRMD_TL_THREADS_FORCE;  // force kernels (measured: 256 -> 194 us, 128 -> 184 us)
constexpr int TL_GROUP = 2;                       // lanes per home atom
#ifndef MRMD_TL_BUILD_BATCH
#define MRMD_TL_BUILD_BATCH 4
#endif
#ifndef MRMD_TL_STAGE_UNROLL
#define MRMD_TL_STAGE_UNROLL 4
#endif
constexpr int TL_STAGE_UNROLL = MRMD_TL_STAGE_UNROLL;  // loads a lane keeps in flight while staging a tile
constexpr int TL_BUILD_BATCH = MRMD_TL_BUILD_BATCH;  // candidates a lane of the neighbour build tests per step
#ifndef MRMD_TL_BUILD_GROUP
#define MRMD_TL_BUILD_GROUP 4
#endif
constexpr int TL_BUILD_GROUP = MRMD_TL_BUILD_GROUP;  // consecutive homes (lanes) that sweep the same candidates
constexpr int TL_PIECES = 27;                     // 9 columns x {low z-wrap, main, high z-wrap}
constexpr int TL_MAX_CH = 64;                     // home cells per tile along z
constexpr int TL_MAX_R = 4;                       // the list radius spans at most this many cells along z
constexpr int TL_CELLS = TL_MAX_CH + 2 * TL_MAX_R + 1;
constexpr int TL_DESC_INTS = 64;                  // per-tile descriptor in global memory

struct TileParams
{
    GridDev g;  // linked-cell grid of the sorted local atoms: cells >= radius along x and y, any size along z
    int R;      // cells along z that the list radius spans: ceil(radius / g.dx[2]) (1 for cubic cells)
    int CH;
    int numChunks;
    int periodic[3];  // ghost layer thickness > 0 on that axis
    double L[3];      // subdomain.diameter
    double minInner[3], maxInner[3];
    int cap;    // shared-memory slots per tile
    int haloX;  // x-slab decomposition: columns -1 and nx hold the halo atoms received from the neighbour ranks
};

// cell ranges are looked up in cellLo / cellHi (absolute atom indices) indexed over the grid extended by the
// two halo columns when haloX is set
__device__ __forceinline__ int extCell(const TileParams& tp, int i, int j, int k)
{
    return ((i + tp.haloX) * tp.g.n[1] + j) * tp.g.n[2] + k;
}

// descriptor layout (ints): [0..26] pieceStart, [27..53] pieceLen, [54] homeStart, [55] homeCount,
// [56] slot of the first home atom, [57] total slots
struct TileDesc
{
    int pieceStart[TL_PIECES];
    int pieceLen[TL_PIECES];
    int pieceSlot[TL_PIECES + 1];
    int homeStart, homeCount, selfSlot0, totalSlots;
};

// Row layout of the tiled list (width is a multiple of 64): the TL_GROUP lanes that share a row owner take the entries
// n = lane, lane + TL_GROUP, ...; entry n sits at (n % TL_GROUP) * (width / TL_GROUP) + n / TL_GROUP, so each lane's
// entries are contiguous and every eight of them are one aligned 16-byte word.
__host__ __device__ __forceinline__ int tiledRowIndex(int n, int width)
{
    return (n % TL_GROUP) * (width / TL_GROUP) + (n / TL_GROUP);
}

// image shift of piece p = (column r = p / 3, z part w = p % 3) of the tile at (ci, cj)
__device__ __forceinline__ void pieceShift(const TileParams& tp, int ci, int cj, int p, int& sx, int& sy, int& sz)
{
    const int r = p / 3, w = p % 3;
    const int ii = ci + r / 3 - 1, jj = cj + r % 3 - 1;
    sx = tp.haloX ? 0 : ((ii < 0) ? -1 : ((ii >= tp.g.n[0]) ? 1 : 0));  // halo atoms arrive already shifted
    sy = (jj < 0) ? -1 : ((jj >= tp.g.n[1]) ? 1 : 0);
    sz = (w == 0) ? -1 : ((w == 2) ? 1 : 0);
}

// one block per tile: writes the descriptor (once per rebuild)
__global__ void __launch_bounds__(32) tileDescKernel(TileParams tp, const int32_t* __restrict__ cellLo,
                                                     const int32_t* __restrict__ cellHi, int* desc, int* maxSlots)
{
    const int tile = blockIdx.x;
    const int col = tile / tp.numChunks, chunk = tile % tp.numChunks;
    const int ci = col / tp.g.n[1], cj = col % tp.g.n[1];
    const int nz = tp.g.n[2];
    const int k0 = chunk * tp.CH;
    const int k1 = min(k0 + tp.CH, nz) - 1;
    const int t = threadIdx.x;
    int start = 0, len = 0;
    if (t < TL_PIECES)
    {
        const int r = t / 3, w = t % 3;
        int sx, sy, sz;
        pieceShift(tp, ci, cj, t, sx, sy, sz);
        int ii = ci + r / 3 - 1 - sx * tp.g.n[0], jj = cj + r % 3 - 1 - sy * tp.g.n[1];
        bool exists = !((sx != 0 && !tp.periodic[0]) || (sy != 0 && !tp.periodic[1]));
        int klo, khi;
        const int R = tp.R;
        if (w == 1) { klo = max(k0 - R, 0); khi = min(k1 + R, nz - 1); }
        else if (w == 0) { klo = nz + (k0 - R); khi = nz - 1; exists = exists && (k0 - R < 0) && tp.periodic[2]; }
        else { klo = 0; khi = k1 + R - nz; exists = exists && (k1 + R > nz - 1) && tp.periodic[2]; }
        if (exists)
        {
            start = cellLo[extCell(tp, ii, jj, klo)];
            len = cellHi[extCell(tp, ii, jj, khi)] - start;
        }
        desc[tile * TL_DESC_INTS + t] = start;
        desc[tile * TL_DESC_INTS + TL_PIECES + t] = len;
    }
    // total slots and the slot of the first home atom (pieces before the centre main piece + offset inside it)
    int before = (t < 13) ? len : 0, total = (t < TL_PIECES) ? len : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        before += __shfl_xor_sync(0xffffffffu, before, o);
        total += __shfl_xor_sync(0xffffffffu, total, o);
    }
    if (t == 0)
    {
        const int homeStart = cellLo[extCell(tp, ci, cj, k0)];
        const int homeCount = cellHi[extCell(tp, ci, cj, k1)] - homeStart;
        const int centreStart = cellLo[extCell(tp, ci, cj, max(k0 - tp.R, 0))];
        desc[tile * TL_DESC_INTS + 54] = homeStart;
        desc[tile * TL_DESC_INTS + 55] = homeCount;
        desc[tile * TL_DESC_INTS + 56] = before + (homeStart - centreStart);
        desc[tile * TL_DESC_INTS + 57] = total;
        if (maxSlots != nullptr) atomicMax(maxSlots, total);
    }
}

__device__ __forceinline__ void loadTileDesc(const int* __restrict__ desc, TileDesc& td, int tile)
{
    const int t = threadIdx.x;
    const int* d = desc + size_t(tile) * TL_DESC_INTS;
    if (t < TL_PIECES)
    {
        td.pieceStart[t] = d[t];
        const int len = d[TL_PIECES + t];
        td.pieceLen[t] = len;
        int x = len;  // inclusive scan over the 27 pieces (warp 0)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int y = __shfl_up_sync(0x07ffffffu, x, o);
            if (t >= o) x += y;
        }
        td.pieceSlot[t + 1] = x;
        if (t == 0) td.pieceSlot[0] = 0;
    }
    else if (t == 32)
    {
        td.homeStart = d[54];
        td.homeCount = d[55];
        td.selfSlot0 = d[56];
        td.totalSlots = d[57];
    }
    __syncthreads();
}

// copies the tile's atoms into shared memory (SoA), image shifts applied with one addition per shifted axis.
// BUILD: also records the source index, or -1 when the atom does not qualify as a ghost image for the piece's
// shift (x < minInner for +L, x >= maxInner for -L: GhostExchange.cpp:80,89).
template <bool BUILD, bool TYPES>
__device__ __forceinline__ void stageTile(const TileParams& tp, const TileDesc& td, const double4* __restrict__ pos,
                                          double* sx_, double* sy_, double* sz_, int* sIdx, unsigned char* sType, int tile)
{
    const int col = tile / tp.numChunks;
    const int ci = col / tp.g.n[1], cj = col % tp.g.n[1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int p = warp; p < TL_PIECES; p += blockDim.x / 32)
    {
        const int len = td.pieceLen[p];
        if (len == 0) continue;
        int ix, iy, iz;
        pieceShift(tp, ci, cj, p, ix, iy, iz);
        // x + (+-L) is the ghost layer's single addition; x + 0.0 leaves x bit-identical
        const double shx = double(ix) * tp.L[0], shy = double(iy) * tp.L[1], shz = double(iz) * tp.L[2];
        const int start = td.pieceStart[p], slot0 = td.pieceSlot[p];
        // TL_STAGE_UNROLL loads in flight per lane: the pieces of a warp are walked one after the other, and one load per
        // trip left the staging phase waiting out a full memory latency per 32 atoms
        for (int k0 = lane; k0 < len; k0 += 32 * TL_STAGE_UNROLL)
        {
            double4 raw[TL_STAGE_UNROLL];
#pragma unroll
            for (int u = 0; u < TL_STAGE_UNROLL; ++u)
                if (k0 + 32 * u < len) raw[u] = ld4nc(pos + start + k0 + 32 * u);
#pragma unroll
            for (int u = 0; u < TL_STAGE_UNROLL; ++u)
            {
                const int k = k0 + 32 * u;
                if (k >= len) break;
                double qx = raw[u].x + shx;
                if (BUILD)
                {
                    // an atom that the reference would not have turned into this ghost image is parked at
                    // x = +inf: it fails every distance test without a separate flag
                    bool ok = true;
                    if (ix > 0) ok = ok && (raw[u].x < tp.minInner[0]);
                    if (ix < 0) ok = ok && (raw[u].x >= tp.maxInner[0]);
                    if (iy > 0) ok = ok && (raw[u].y < tp.minInner[1]);
                    if (iy < 0) ok = ok && (raw[u].y >= tp.maxInner[1]);
                    if (iz > 0) ok = ok && (raw[u].z < tp.minInner[2]);
                    if (iz < 0) ok = ok && (raw[u].z >= tp.maxInner[2]);
                    if (!ok) qx = __longlong_as_double(0x7ff0000000000000LL);
                }
                sx_[3 * (slot0 + k)] = qx;
                sy_[3 * (slot0 + k)] = raw[u].y + shy;
                sz_[3 * (slot0 + k)] = raw[u].z + shz;
                if (TYPES) sType[slot0 + k] = static_cast<unsigned char>(typeOf(raw[u]));
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ double minDist1T(const GridDev& g, double x, int c, int d)
{
    const double xc = __dadd_rn(g.min[d], __dmul_rn(double(c) + 0.5, g.dx[d]));
    const double rr = __dsub_rn(fabs(__dsub_rn(x, xc)), __dmul_rn(0.5, g.dx[d]));
    return (rr > 0.0) ? rr : 0.0;
}

// Cabana's stencil pruning for an accepted pair: the stencil cell that holds n must be within r of p
__device__ __forceinline__ bool cabanaCellReachable(const GridDev& cg, double px, double py, double pz, double qx,
                                                    double qy, double qz, double rsqr)
{
    const int a = locate1(cg, qx, 0), b = locate1(cg, qy, 1), c = locate1(cg, qz, 2);
    const double rx = minDist1T(cg, px, a, 0), ry = minDist1T(cg, py, b, 1), rz = minDist1T(cg, pz, c, 2);
    return __dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz)) <= rsqr;
}

// AdResS with a slab weighting function: true when the three staged columns of the tile (padded by one more cell on
// each side for the drift until the next sort) lie beyond the hybrid region, i.e. every pair of the tile is an
// ideal-gas pair (LJ_IdealGas.cpp:102-107).  The force kernel skips such tiles and, in the step-loop drivers, so does
// the neighbour build (their rows stay empty); both use this one criterion.
__device__ __forceinline__ bool tileAllCoarseGrained(const TileParams& tp, const mrmd_b200_weight& w, int tile)
{
    const int col = tile / tp.numChunks;
    const int ci = col / tp.g.n[1];
    const double lo = tp.g.min[0] + double(ci - 2) * tp.g.dx[0];
    const double hi = tp.g.min[0] + double(ci + 3) * tp.g.dx[0];
    if (w.kind == MRMD_B200_WEIGHT_SLAB)
    {
        const double reach = 0.5 * w.atRegion + w.hyRegion;
        return (lo - w.center[0] > reach) || (w.center[0] - hi > reach);
    }
    // spherical region: distance from the centre to the box of everything the tile stages (in the tile's own image
    // frame), padded by one more cell on every side
    const int cj = col % tp.g.n[1], chunk = tile % tp.numChunks;
    const int k0 = chunk * tp.CH, k1 = min(k0 + tp.CH, tp.g.n[2]) - 1;
    const double ylo = tp.g.min[1] + double(cj - 2) * tp.g.dx[1], yhi = tp.g.min[1] + double(cj + 3) * tp.g.dx[1];
    const double zlo = tp.g.min[2] + double(k0 - tp.R - 1) * tp.g.dx[2], zhi = tp.g.min[2] + double(k1 + tp.R + 2) * tp.g.dx[2];
    const double dx = fmax(fmax(lo - w.center[0], w.center[0] - hi), 0.0);
    const double dy = fmax(fmax(ylo - w.center[1], w.center[1] - yhi), 0.0);
    const double dz = fmax(fmax(zlo - w.center[2], w.center[2] - zhi), 0.0);
    const double reach = w.atRegion + w.hyRegion;
    return dx * dx + dy * dy + dz * dz > reach * reach;
}

// Measured in round 2 and dropped (profiles/r02_build_experiments.md): (a) walking the nine ranges of a home as ONE flat
// candidate sequence, so that a warp's trip count is max_h(sum_r c[h][r]) instead of sum_r max_h(c[h][r]) (548 instead of
// 787 warp iterations per 110-atom tile): some lane crosses a column boundary in almost every iteration and the divergent
// advance costs more than the saved iterations (838 vs 461 us per 1M atoms); (b) taking the homes of a tile up in order
// of candidate count / row length (whole tile or windows of 32 / 64 homes): the homes of a warp are then no longer
// neighbours in z, their shared-memory reads scatter over the staged set, bank conflicts eat the balance (build +5 %,
// force +0.5 %); (c) tiles sized for a 55 KB budget including the per-home tables of (a): half the homes per tile, twice
// the staging (build 694 us, force 213 us).
__device__ __forceinline__ void stSharedU16(unsigned addr, uint16_t v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}

// Neighbour build on tiles, one lane per home atom.  The homes of a warp are 32 consecutive atoms of the column (z
// order); in each of the nine columns a lane sweeps the slot range that the cutoff sphere around its home cuts out of
// the column, all lanes in lock step, so the lanes of a warp read staged records a few slots apart (mostly the same
// shared-memory lines) and the loop body is the bare criterion: three loads, the uncontracted distance, one predicated
// 2-byte store into the lane's row buffer.  Rows fill in scan order (column by column, slots ascending): deterministic,
// no atomics, no ballots.  The trip count of a column is the longest range of the warp (about 25 slots against a mean of
// 15).  The round-1 kernel gave two lanes to a home and compacted through a warp ballot per step: 52 issued instructions
// per step against 17 here (profiles/r02_build_ncu.txt: 379 M warp instructions per 1 M atoms, issue bound).
template <bool HALF>
__global__ void __launch_bounds__(TL_THREADS_BUILD, 4)
    verletBuildTiledKernel(TileParams tp, GridDev cabanaGrid, const double4* __restrict__ pos,
                           const int32_t* __restrict__ cellLo, const int* __restrict__ desc, double rsqr, int width,
                           int32_t* __restrict__ counts, uint16_t* __restrict__ enc, int32_t* stats, int32_t* tstats,
                           unsigned char* __restrict__ tileActive, int cgSkip, mrmd_b200_weight cgWeight)
{
    extern __shared__ double sTile[];
    __shared__ TileDesc td;
    __shared__ int cellSlot[9][TL_CELLS];  // slot where virtual cell v (= k0 - 1 + v) of column r starts
    {
        const int* d = desc + size_t(blockIdx.x) * TL_DESC_INTS;
        const int homeStart = d[54], homeCount = d[55];
        // the tile geometry is re-used from the previous rebuild without a host round trip: a tile that outgrew the
        // staged capacity raises a flag (the host rebuilds with measured tiles) and leaves its rows empty
        const bool overflow = d[57] > tp.cap;
        if (threadIdx.x == 0)
        {
            atomicMax(tstats, d[57]);  // sizes the next rebuild's tiles
            if (overflow) tstats[2] = 1;
        }
        const bool skipTile = overflow || (cgSkip && tileAllCoarseGrained(tp, cgWeight, blockIdx.x));
        if (threadIdx.x == 0) tileActive[blockIdx.x] = skipTile ? 0 : 1;
        if (skipTile)
        {
            // AdResS step loops: no pair of a coarse-grained tile is ever evaluated, its rows stay empty
            for (int h = threadIdx.x; h < homeCount; h += blockDim.x) counts[homeStart + h] = 0;
            return;
        }
    }
    loadTileDesc(desc, td, blockIdx.x);
    double* sx_ = sTile;
    double* sy_ = sx_ + 1;  // interleaved {x, y, z} records: one address per slot, conflict-free for consecutive slots
    double* sz_ = sx_ + 2;
    const int tile = blockIdx.x;
    const int col = tile / tp.numChunks, chunk = tile % tp.numChunks;
    const int ci = col / tp.g.n[1], cj = col % tp.g.n[1];
    const int nz = tp.g.n[2];
    const int k0 = chunk * tp.CH;
    const int k1 = min(k0 + tp.CH, nz) - 1;
    const int nk = k1 - k0 + 1;
    const int R = tp.R;
    const int nv = nk + 2 * R + 1;  // virtual cells k0-R .. k1+R plus the end sentinel
    // a warp takes every (blockDim.x / 32)-th column, its lanes the virtual cells; the loads of a thread are independent
    // and issued together (no division, one memory latency)
    {
        constexpr int V_TRIPS = (TL_CELLS + 31) / 32;
        const int warp = threadIdx.x >> 5, lane32 = threadIdx.x & 31, warps = blockDim.x >> 5;
        for (int r0 = warp; r0 < 9; r0 += 3 * warps)
        {
            int slotOf[3][V_TRIPS];
#pragma unroll
            for (int c = 0; c < 3; ++c)
            {
                const int r = r0 + c * warps;
                int ii = ci + r / 3 - 1, jj = cj + r % 3 - 1;
                if (!tp.haloX)
                {
                    if (ii < 0) ii += tp.g.n[0]; else if (ii >= tp.g.n[0]) ii -= tp.g.n[0];
                }
                if (jj < 0) jj += tp.g.n[1]; else if (jj >= tp.g.n[1]) jj -= tp.g.n[1];
#pragma unroll
                for (int t = 0; t < V_TRIPS; ++t)
                {
                    const int v = lane32 + 32 * t;
                    const int kv = k0 - R + v;
                    // virtual cells below 0 / beyond nz-1 live in the low / high z-wrap piece (cells nz+kv / kv-nz of the column)
                    const int piece = r * 3 + ((kv < 0) ? 0 : ((kv >= nz) ? 2 : 1));
                    int slot = 0;
                    if (r < 9 && v < nv)
                    {
                        if (v == nv - 1) slot = td.pieceSlot[r * 3 + 3];
                        else if (td.pieceLen[piece] == 0) slot = td.pieceSlot[piece];
                        else
                        {
                            const int kk = (kv < 0) ? kv + nz : ((kv >= nz) ? kv - nz : kv);
                            slot = td.pieceSlot[piece] + (cellLo[extCell(tp, ii, jj, kk)] - td.pieceStart[piece]);
                        }
                    }
                    slotOf[c][t] = slot;
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int t = 0; t < V_TRIPS; ++t)
                {
                    const int r = r0 + c * warps, v = lane32 + 32 * t;
                    if (r < 9 && v < nv) cellSlot[r][v] = slotOf[c][t];
                }
        }
    }
    stageTile<true, false>(tp, td, pos, sx_, sy_, sz_, nullptr, nullptr, tile);

    const int lane = threadIdx.x & 31;
    int mx = 0;
    long long total = 0;
    const double rsqrSafe = rsqr * (1.0 - 1e-9);
    // Accepted slots are collected in the lane's own shared-memory row, in the layout of the list (tiledRowIndex), and
    // leave as coalesced 4-byte words.  The row pitch is an odd number of words: lanes with equal counts hit different
    // banks, and the copy-out below reads consecutive words.
    const int rowWords = (width >> 1) + 1;
    uint32_t* const sRows = reinterpret_cast<uint32_t*>(sx_ + 3 * (tp.cap + TL_BUILD_BATCH));
    uint16_t* sRow = reinterpret_cast<uint16_t*>(sRows + threadIdx.x * rowWords);
    const unsigned rowBegin = static_cast<unsigned>(__cvta_generic_to_shared(sRow)), rowEnd = rowBegin + 2u * unsigned(width);
    const int halfWidth = width / TL_GROUP;
    const int safeHi = __double2hiint(rsqrSafe);  // d2 > 0: the high words order like the values
    const int wordsPerLane = halfWidth >> 1;  // 4-byte words of the entries of one of the TL_GROUP list lanes
    const int eFirst = lane / wordsPerLane + TL_GROUP * ((lane % wordsPerLane) << 1);  // first entry of row word `lane`
    const float hzMax = sqrtf(static_cast<float>(rsqr)) * 1.001f * static_cast<float>(tp.g.rdx[2]) + 1e-3f;  // list radius in cells
    const int vBase = k0 - R;  // cell index of virtual cell 0
    const double zBase = tp.g.min[2] + double(vBase) * tp.g.dx[2];
    const bool clampZ = !tp.periodic[2];  // atoms beyond a non-periodic face are binned into the boundary cells
    for (int hBase = 0; hBase < td.homeCount; hBase += blockDim.x)
    {
        const int h = hBase + threadIdx.x;
        const bool active = h < td.homeCount;
        const int i = td.homeStart + h;
        const int selfSlot = active ? td.selfSlot0 + h : -1;
        const int safeSlot = active ? selfSlot : 0;
        // (a lane without a home sits at x = +inf: it accepts nothing)
        const double px = active ? sx_[3 * safeSlot] : __longlong_as_double(0x7ff0000000000000LL);
        const double py = sy_[3 * safeSlot], pz = sz_[3 * safeSlot];
        unsigned rowAt = rowBegin;  // shared-memory byte address of the next entry of the lane's row
        unsigned borderAcc = 0xffffffffu;
        // The TL_BUILD_GROUP lanes of a group (consecutive homes of the column, i.e. neighbours in z) sweep the same
        // candidates in lock step, so that a load instruction of the warp touches 32 / TL_BUILD_GROUP records instead of
        // 32 (shared-memory bandwidth, 128 B per clock and SM, is what bounds a sweep over per-lane ranges: 24 B x 32
        // lanes per step).  In every column the group scans the cells within the list radius of its lowest and highest
        // home along z.  Atoms are in cell order along z, so that is one slot range per column.  (Cutting the range of a
        // column down by the homes' distance to it was measured and dropped: the nine per-lane range computations and
        // group reductions it takes were 20 % of the kernel's stall samples, more than the shorter sweeps gave back.)  This
        // is bookkeeping, not the list criterion: single precision on coordinates relative to the tile (float error
        // ~1e-6), widened by 0.1 % of r and 1e-3 of a cell, far more than that error and than the ulp by which an atom
        // may sit outside its cell.
        const float zCells = static_cast<float>((pz - zBase) * tp.g.rdx[2]);  // z in cells, relative to virtual cell 0
        float zLo = active ? zCells : 3.0e38f, zHi = active ? zCells : -3.0e38f;
#pragma unroll
        for (int o = 1; o < TL_BUILD_GROUP; o <<= 1)
        {
            zLo = fminf(zLo, __shfl_xor_sync(0xffffffffu, zLo, o));
            zHi = fmaxf(zHi, __shfl_xor_sync(0xffffffffu, zHi, o));
        }
        int vA = 0, vB = -1;
        if (zHi >= zLo)  // a group with a home
        {
            int kA = static_cast<int>(floorf(zLo - hzMax)) + vBase;
            int kB = static_cast<int>(floorf(zHi + hzMax)) + vBase;
            if (clampZ)
            {
                kA = max(0, min(kA, nz - 1));
                kB = max(0, min(kB, nz - 1));
            }
            vA = max(kA - vBase, 0);
            vB = min(kB - vBase, nv - 2);
        }
        int rs0[9], rs1[9], rIters[9];
#pragma unroll
        for (int r = 0; r < 9; ++r)
        {
            int s0 = 0, s1 = 0;
            if (vB >= vA)
            {
                s0 = cellSlot[r][vA];
                s1 = cellSlot[r][vB + 1];
            }
            rs0[r] = s0;
            rs1[r] = s1;
            rIters[r] = __reduce_max_sync(0xffffffffu, s1 - s0);
        }
#pragma unroll
        for (int r = 0; r < 9; ++r)
        {
            const int s0 = rs0[r], s1 = rs1[r], iters = rIters[r];
            // Candidates are taken TL_BUILD_BATCH at a time: all loads and distances first, then the (predicated)
            // stores, so that the dependent chains interleave (the compiler does not move shared-memory loads across the
            // stores).  Steps past the end of a group's range read the records behind it (the staged set ends with
            // TL_BUILD_BATCH spare records) and are rejected by their position in the range.
            const int len = s1 - s0;
            const int selfAt = (r == 4) ? selfSlot - s0 : -1;  // only the centre column holds the home atom itself
            const double* q = sx_ + 3 * s0;
            for (int it = 0; it < iters; it += TL_BUILD_BATCH, q += 3 * TL_BUILD_BATCH)
            {
                double d2[TL_BUILD_BATCH];
                bool ok[TL_BUILD_BATCH];
#pragma unroll
                for (int u = 0; u < TL_BUILD_BATCH; ++u)
                {
                    const double qx = q[3 * u], qy = q[3 * u + 1], qz = q[3 * u + 2];
                    d2[u] = distSqrExact(px - qx, py - qy, pz - qz);
                    ok[u] = (it + u < len) && (d2[u] <= rsqr);
                    if (r == 4) ok[u] = ok[u] && (it + u != selfAt);
                    if (HALF) ok[u] = ok[u] && ((qx > px) || ((qx == px) && ((qy > py) || ((qy == py) && (qz > pz)))));
                }
#pragma unroll
                for (int u = 0; u < TL_BUILD_BATCH; ++u)
                {
                    // smallest high word of d2 above that of (1 - 1e-9) r^2 (wrapping below it): integer pipe, one
                    // instruction; candidates that are not accepted may raise the flag too, the filter below is exact
                    borderAcc = min(borderAcc, unsigned(__double2hiint(d2[u]) - safeHi));
                    if (ok[u] && rowAt < rowEnd) stSharedU16(rowAt, static_cast<uint16_t>(s0 + it + u));
                    rowAt += ok[u] ? 2u : 0u;
                }
            }
        }
        const unsigned count = (rowAt - rowBegin) >> 1;
        const bool border = borderAcc <= unsigned(__double2hiint(rsqr) - safeHi);
        static_assert(TL_GROUP == 2, "verletBuildTiledKernel writes the row layout of two list lanes per home");
        // Cabana prunes stencil cells by their distance to p; the cell holding n can only fail that test when d2 is
        // within rounding of r^2 (the cell contains n): rows that accepted a partner with d2 > (1 - 1e-9) r^2 are
        // filtered once more, in place (rare)
        unsigned dropped = 0;
        if (border)
        {
            const unsigned stored = min(count, unsigned(width));
            unsigned kept = 0;
            for (unsigned n = 0; n < stored; ++n)
            {
                const uint16_t s = sRow[n];
                const double* q = sx_ + 3 * s;
                const double qx = q[0], qy = q[1], qz = q[2];
                const double d2 = distSqrExact(px - qx, py - qy, pz - qz);
                if (d2 > rsqrSafe && !cabanaCellReachable(cabanaGrid, px, py, pz, qx, qy, qz, rsqr)) continue;
                sRow[kept] = s;
                kept += 1;
            }
            // rows longer than the buffer are rebuilt with a wider list anyway (the host grows width to the longest count)
            dropped = stored - kept;
        }
        const int finalCount = int(count - dropped);
        if (active)
        {
            counts[i] = finalCount;
            mx = max(mx, finalCount);
            total += finalCount;
        }
        // The rows of the warp's 32 homes leave one after the other, a 4-byte word per lane, in the layout of the list
        // (tiledRowIndex: the entries n, n + TL_GROUP, ... of one list lane are contiguous); words without a stored
        // entry are skipped (entries past count are never read).  The stored count travels in the padding word of the row.
        sRows[threadIdx.x * rowWords + (width >> 1)] = active ? unsigned(min(finalCount, width)) : 0u;
        __syncwarp();
        {
            const int warpFirst = threadIdx.x - lane;
            const uint32_t* cntp = sRows + warpFirst * rowWords + (width >> 1);
            const uint16_t* sp = reinterpret_cast<const uint16_t*>(sRows + warpFirst * rowWords) + eFirst;
            uint32_t* dp = reinterpret_cast<uint32_t*>(enc + size_t(td.homeStart + hBase + warpFirst) * width) + lane;
#pragma unroll 4
            for (int rr = 0; rr < 32; ++rr, cntp += rowWords, sp += 2 * rowWords, dp += (width >> 1))
            {
                const int cnt = int(*cntp);  // same word for all lanes
                if (eFirst < cnt) *dp = uint32_t(sp[0]) | (uint32_t(sp[TL_GROUP]) << 16);
                if (width > 64)  // rows wider than 64 entries: the words 32, 33, ... of the row
                    for (int c = lane + 32; c < (width >> 1); c += 32)
                    {
                        // word c holds the entries e and e + TL_GROUP of list lane c / wordsPerLane
                        const int e = c / wordsPerLane + TL_GROUP * ((c % wordsPerLane) << 1);
                        if (e < cnt) dp[c - lane] = uint32_t(sp[e - eFirst]) | (uint32_t(sp[e - eFirst + TL_GROUP]) << 16);
                    }
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        total += __shfl_xor_sync(0xffffffffu, total, o);
    }
    if (lane == 0 && total > 0)
    {
        atomicMax(stats, mx);
        atomicAdd(reinterpret_cast<unsigned long long*>(stats + 2), static_cast<unsigned long long>(total));
    }
}

// ordered list of the tiles with work (one block): the AdResS force kernels launch over it instead of over all tiles
__global__ void __launch_bounds__(1024) compactActiveTilesKernel(const unsigned char* __restrict__ tileActive, int tiles,
                                                                 int32_t* __restrict__ activeTiles, int32_t* count)
{
    __shared__ int sWarp[32];
    __shared__ int sBase;
    if (threadIdx.x == 0) sBase = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < tiles; base += blockDim.x)
    {
        const int t = base + threadIdx.x;
        const bool on = t < tiles && tileActive[t] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, on);
        if (lane == 0) sWarp[warp] = __popc(m);
        __syncthreads();
        int off = sBase;
        for (int w = 0; w < warp; ++w) off += sWarp[w];
        if (on) activeTiles[off + __popc(m & ((1u << lane) - 1u))] = t;
        __syncthreads();
        if (threadIdx.x == 0)
        {
            int tot = 0;
            for (int w = 0; w < int(blockDim.x >> 5); ++w) tot += sWarp[w];
            sBase += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = sBase;
}

// reciprocal from the hardware seed (20 mantissa bits, error e <= 2^-20) with one cubically convergent step
// r (1 + e + e^2): relative error e^3 < 1e-18, three DFMA, no slow-path branch (the force is compared to 1e-10 relative;
// the cutoff decisions never use it)
__device__ __forceinline__ double fastRcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#ifdef MRMD_RCP_NEWTON2  // measurement knob: the round-1 variant, two Newton steps (four DFMA)
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
#else
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
#endif
}

// Sums four per-lane values over the TL_GROUP lanes of a group.  Every step halves the number of values a lane carries
// instead of running a butterfly per value (8 lanes: four 64-bit shuffles instead of twelve).  Afterwards a lane holds
// TL_VPL of the four totals: out[k] of lane gl is value groupSumValue(gl, k) (-1: a duplicate another lane stores).
constexpr int TL_VPL = (TL_GROUP >= 4) ? 1 : 4 / TL_GROUP;
__device__ __forceinline__ void groupSum4(double v0, double v1, double v2, double v3, int gl, double (&out)[TL_VPL])
{
    static_assert(TL_GROUP == 8 || TL_GROUP == 4 || TL_GROUP == 2 || TL_GROUP == 1, "groupSum4: groups of 8, 4, 2 or 1 lanes");
    if constexpr (TL_GROUP >= 4)
    {
        constexpr int HI = TL_GROUP / 2, LO = TL_GROUP / 4;  // the two lane bits that select the value a lane ends up with
        const bool bh = (gl & HI) != 0, bl = (gl & LO) != 0;
        const double r0 = __shfl_xor_sync(0xffffffffu, bh ? v0 : v2, HI);
        const double r1 = __shfl_xor_sync(0xffffffffu, bh ? v1 : v3, HI);
        const double u0 = (bh ? v2 : v0) + r0;
        const double u1 = (bh ? v3 : v1) + r1;
        const double r2 = __shfl_xor_sync(0xffffffffu, bl ? u0 : u1, LO);
        double w = (bl ? u1 : u0) + r2;
        if (TL_GROUP == 8) w += __shfl_xor_sync(0xffffffffu, w, 1);
        out[0] = w;
    }
    else if constexpr (TL_GROUP == 1)
    {
        out[0] = v0;
        out[1 % TL_VPL] = v1;
        out[2 % TL_VPL] = v2;
        out[3 % TL_VPL] = v3;
    }
    else
    {
        const bool b = (gl & 1) != 0;  // lane 0 ends up with (v0, v1), lane 1 with (v2, v3)
        const double r0 = __shfl_xor_sync(0xffffffffu, b ? v0 : v2, 1);
        const double r1 = __shfl_xor_sync(0xffffffffu, b ? v1 : v3, 1);
        out[0] = (b ? v2 : v0) + r0;
        out[TL_VPL - 1] = (b ? v3 : v1) + r1;
    }
}
__device__ __forceinline__ int groupSumValue(int gl, int k)
{
    if (TL_GROUP == 8) return ((gl & 1) == 0) ? (gl >> 1) : -1;
    return gl * TL_VPL + k;
}
// lane and slot that hold value v
__device__ __forceinline__ int groupSumLane(int v) { return (TL_GROUP == 8) ? 2 * v : v / TL_VPL; }
__device__ __forceinline__ int groupSumSlot(int v) { return v % TL_VPL; }

// 16-byte words (eight entries each) of a lane's half of a list row that the LJ kernel loads a pass ahead
#ifndef MRMD_LJT_WORDS
#define MRMD_LJT_WORDS 1
#endif
constexpr int LJT_WORDS = (MRMD_LJT_WORDS * 8 * TL_GROUP <= 64) ? MRMD_LJT_WORDS : 64 / (8 * TL_GROUP);
// ... and further words, also loaded a pass ahead, that a rolled loop walks (rows are at least 64 entries wide)
#ifndef MRMD_LJT_MORE_WORDS
#define MRMD_LJT_MORE_WORDS 3
#endif
constexpr int LJT_MORE = ((MRMD_LJT_WORDS + MRMD_LJT_MORE_WORDS) * 8 * TL_GROUP <= 64) ? MRMD_LJT_MORE_WORDS : 64 / (8 * TL_GROUP) - LJT_WORDS;

// row length and list words of the home that lane (group, gl) works on in the pass starting at home hBase
__device__ __forceinline__ void prefetchListRows(const TileDesc& td, const int32_t* __restrict__ counts,
                                                 const uint16_t* __restrict__ enc, int width, int hBase, int group, int gl,
                                                 int& count, uint4 (&words)[LJT_WORDS])
{
    const int h = hBase + group;
    const bool active = h < td.homeCount;
    const size_t i = active ? size_t(td.homeStart + h) : 0;  // lanes without a home read row 0 (always there)
    count = active ? counts[i] : 0;
    // lane gl owns the entries gl, gl + TL_GROUP, ... of the row; the row layout keeps them contiguous
    // (tiledRowIndex), so they arrive as 16-byte words
    const uint4* row = reinterpret_cast<const uint4*>(enc + i * width + gl * (width / TL_GROUP));
#pragma unroll
    for (int k = 0; k < LJT_WORDS; ++k) words[k] = row[k];
}

// one pair of LennardJones::apply_if's inner loop (LennardJones.hpp:176-196) against a staged partner.
// CAPPED_BRANCH = false: the force law without its capped branch (LennardJones.hpp:56-67), straight-line code; a pair
// closer than the capping distance raises cappedSeen and the caller evaluates the row again with the branch.
template <bool SINGLE_TYPE, bool ENERGY, bool CAPPED_BRANCH>
__device__ __forceinline__ void ljPair(const double* sx_, const double* sy_, const double* sz_,
                                       const unsigned char* sType, int slot, bool valid, double xi, double yi, double zi,
                                       int typeI, const LJType& t0, const LJTable& table, int64_t numTypesQuirk,
                                       double rcSqr, double& fx, double& fy, double& fz, double& energy, double& virial,
                                       int& pairs, int& minHi)
{
    // predicated, not branched: consecutive list entries are independent, and without a branch around every pair
    // the compiler interleaves their dependent FP64 chains (an invalid entry reads slot 0 and contributes zero)
    const int s = valid ? slot : 0;
    const double dx = xi - sx_[3 * s];
    const double dy = yi - sy_[3 * s];
    const double dz = zi - sz_[3 * s];
    const double distSqr = distSqrExact(dx, dy, dz);
    const bool in = valid && (distSqr <= rcSqr);  // LennardJones.hpp:182 skips distSqr > rcSqr
    const LJType& t = SINGLE_TYPE ? t0 : table.t[typeI * numTypesQuirk + sType[s]];
    const double d2 = in ? distSqr : rcSqr;  // keeps the arithmetic of a skipped pair finite
    double ff, e;
    if (!CAPPED_BRANCH || d2 >= t.cappingDistanceSqr)  // LennardJones.hpp:56-67
    {
        const double frac2 = fastRcp(d2);
        const double frac6 = frac2 * frac2 * frac2;
        ff = frac6 * (t.ff1 * frac6 - t.ff2) * frac2;
        e = frac6 * (t.ef1 * frac6 - t.ef2) - t.shift;
        // (the smallest high word of d2 tells the caller whether a pair came closer than the capping distance)
        if (!CAPPED_BRANCH) minHi = min(minHi, __double2hiint(d2));
    }
    else
        ljForceEnergy(t, d2, ff, e);
    ff = in ? ff : 0.0;
    if (ENERGY)
    {
        energy += in ? e : 0.0;
        virial -= 0.5 * ff * d2;
    }
    pairs += in ? 1 : 0;  // integer pipe: the FP64 pipe is the busy one
    fx += dx * ff;
    fy += dy * ff;
    fz += dz * ff;
}

// LennardJones::apply over the tiled list (full list: row owners only).  ACCUMULATE = false stores the force
// (the driver then skips the force reset); true adds like the reference.  ENERGY = false skips the energy /
// virial accumulation (the driver asks for them on the last step of a run only).
#ifndef MRMD_LJT_MINBLOCKS
#define MRMD_LJT_MINBLOCKS 6
#endif
template <bool SINGLE_TYPE, bool ACCUMULATE, bool ENERGY>
__global__ void __launch_bounds__(TL_THREADS_FORCE, MRMD_LJT_MINBLOCKS)
    ljForceTiledKernel(TileParams tp, AtomsView a, const int* __restrict__ desc, const int32_t* __restrict__ counts,
                       const uint16_t* __restrict__ enc, int width, LJTable table, double rcSqr, int64_t numTypesQuirk,
                       double* partials, double* result, unsigned int* ticket, const int* __restrict__ stop)
{
    if (stop != nullptr && *stop != 0) return;  // a step queued behind the one that asked for a rebuild
    extern __shared__ double sTile[];
    __shared__ TileDesc td;
    loadTileDesc(desc, td, blockIdx.x);
    double* sx_ = sTile;
    double* sy_ = sx_ + 1;  // interleaved {x, y, z} records: one address per slot, conflict-free for consecutive slots
    double* sz_ = sx_ + 2;
    unsigned char* sType = reinterpret_cast<unsigned char*>(sx_ + 3 * tp.cap);

    // warp-uniform control flow, see verletBuildTiledKernel
    const int group = threadIdx.x / TL_GROUP, gl = threadIdx.x % TL_GROUP;
    double energy = 0.0, virial = 0.0;
    int pairs = 0;
    const LJType t0 = table.t[0];
    // high word of the largest capping distance (squared) of the table: a row with a smaller d2 takes the slow path
    int capHi = __double2hiint(t0.cappingDistanceSqr);
    if (!SINGLE_TYPE)
        for (int k = 1; k < MAX_LJ_TYPES * MAX_LJ_TYPES; ++k) capHi = max(capHi, __double2hiint(table.t[k].cappingDistanceSqr));
    // The row length and the list words of a pass are loaded one pass ahead (those of the first pass while the tile is
    // staged): a dependent global load at the head of every pass and one per entry behind the prefetched words were
    // 11 % of the kernel's stall samples (profiles/r02_lj_force_ncu.txt).
    const int homesPerPass = blockDim.x / TL_GROUP;
    int countNext = 0;
    uint4 wordsNext[LJT_WORDS];
    prefetchListRows(td, counts, enc, width, 0, group, gl, countNext, wordsNext);
    stageTile<false, !SINGLE_TYPE>(tp, td, a.pos, sx_, sy_, sz_, nullptr, sType, blockIdx.x);
    for (int hBase = 0; hBase < td.homeCount; hBase += homesPerPass)
    {
        const int h = hBase + group;
        const bool active = h < td.homeCount;
        const int i = td.homeStart + h;
        const int selfSlot = active ? td.selfSlot0 + h : 0;
        const double xi = sx_[3 * selfSlot], yi = sy_[3 * selfSlot], zi = sz_[3 * selfSlot];
        const int typeI = SINGLE_TYPE ? 0 : sType[selfSlot];
        double fx = 0.0, fy = 0.0, fz = 0.0, ePass = 0.0, vPass = 0.0;
        int pPass = 0;
        int minHi = 0x7fffffff;
        const int numNeighbors = min(countNext, width);
        unsigned words[4 * (LJT_WORDS + LJT_MORE)];
#pragma unroll
        for (int k = 0; k < LJT_WORDS; ++k)
        {
            words[4 * k] = wordsNext[k].x;
            words[4 * k + 1] = wordsNext[k].y;
            words[4 * k + 2] = wordsNext[k].z;
            words[4 * k + 3] = wordsNext[k].w;
        }
        // the words behind the unrolled steps are loaded now and used after them
        {
            const uint4* row = reinterpret_cast<const uint4*>(enc + size_t(active ? i : 0) * width + gl * (width / TL_GROUP));
#pragma unroll
            for (int k = LJT_WORDS; k < LJT_WORDS + LJT_MORE; ++k)
            {
                const uint4 w = row[k];
                words[4 * k] = w.x;
                words[4 * k + 1] = w.y;
                words[4 * k + 2] = w.z;
                words[4 * k + 3] = w.w;
            }
        }
        if (hBase + homesPerPass < td.homeCount)  // block uniform
            prefetchListRows(td, counts, enc, width, hBase + homesPerPass, group, gl, countNext, wordsNext);
        const int mine = (numNeighbors - gl + TL_GROUP - 1) / TL_GROUP;  // entries of this lane
        const int iters = __reduce_max_sync(0xffffffffu, (numNeighbors + TL_GROUP - 1) / TL_GROUP);
        // the first words without a test per step (a row shorter than them is padded with predicated-off steps)
        if (iters > 0)  // warp uniform
        {
#pragma unroll
            for (int it = 0; it < 8 * LJT_WORDS; ++it)
            {
                const int slot = (words[it >> 1] >> (16 * (it & 1))) & 0xffffu;
                ljPair<SINGLE_TYPE, ENERGY, false>(sx_, sy_, sz_, sType, slot, it < mine, xi, yi, zi, typeI, t0, table,
                                                   numTypesQuirk, rcSqr, fx, fy, fz, ePass, vPass, pPass, minHi);
            }
        }
        // the entries behind the unrolled steps: a rolled loop that takes two entries per trip from the lowest of the
        // remaining prefetched words and moves the others down (registers cannot be indexed)
        if (LJT_MORE > 0)
        {
            const int more = min(iters, 8 * (LJT_WORDS + LJT_MORE));
#pragma unroll 1
            for (int it = 8 * LJT_WORDS; it < more; it += 2)
            {
                const unsigned w = words[4 * LJT_WORDS];
#pragma unroll
                for (int k = 4 * LJT_WORDS; k + 1 < 4 * (LJT_WORDS + LJT_MORE); ++k) words[k] = words[k + 1];
                ljPair<SINGLE_TYPE, ENERGY, false>(sx_, sy_, sz_, sType, w & 0xffffu, it < mine, xi, yi, zi, typeI, t0, table,
                                            numTypesQuirk, rcSqr, fx, fy, fz, ePass, vPass, pPass, minHi);
#ifdef MRMD_LJT_TAIL_BRANCH
                if (it + 1 < iters)
#endif
                    ljPair<SINGLE_TYPE, ENERGY, false>(sx_, sy_, sz_, sType, w >> 16, it + 1 < mine, xi, yi, zi, typeI, t0, table,
                                                numTypesQuirk, rcSqr, fx, fy, fz, ePass, vPass, pPass, minHi);
            }
        }
        if (iters > 8 * (LJT_WORDS + LJT_MORE))  // rows longer than the prefetched words (lists wider than 64 entries)
        {
            const uint16_t* mineRow = enc + size_t(active ? i : 0) * width + gl * (width / TL_GROUP);
            for (int it = 8 * (LJT_WORDS + LJT_MORE); it < iters; ++it)
            {
                if (it < mine)
                    ljPair<SINGLE_TYPE, ENERGY, false>(sx_, sy_, sz_, sType, mineRow[it], true, xi, yi, zi, typeI, t0, table,
                                                numTypesQuirk, rcSqr, fx, fy, fz, ePass, vPass, pPass, minHi);
            }
        }
        // a pair inside the capping distance (a lattice that melts, an overlap): the warp evaluates its rows again with
        // the capped branch of the force law
        if (__any_sync(0xffffffffu, minHi <= capHi))
        {
            fx = fy = fz = ePass = vPass = 0.0;
            pPass = 0;
            const uint16_t* mineRow = enc + size_t(active ? i : 0) * width + gl * (width / TL_GROUP);
            for (int it = 0; it < mine; ++it)
                ljPair<SINGLE_TYPE, ENERGY, true>(sx_, sy_, sz_, sType, mineRow[it], true, xi, yi, zi, typeI, t0, table,
                                                  numTypesQuirk, rcSqr, fx, fy, fz, ePass, vPass, pPass, minHi);
        }
        energy += ePass;
        virial += vPass;
        pairs += pPass;
        // three lanes of the group end up with the x / y / z total and store it
        double f[TL_VPL];
        groupSum4(fx, fy, fz, 0.0, gl, f);
#pragma unroll
        for (int k = 0; k < TL_VPL; ++k)
        {
            const int comp = groupSumValue(gl, k);
            if (active && comp >= 0 && comp < 3)
            {
                double* plane = (comp == 0) ? a.force[0] : ((comp == 1) ? a.force[1] : a.force[2]);
                if (ACCUMULATE) plane[i] += f[k];
                else plane[i] = f[k];
            }
        }
    }
    // every pair is visited from both sides
    if (ENERGY) gridReduce3<TL_THREADS_FORCE>(0.5 * energy, 0.5 * virial, 0.5 * double(pairs), partials, result, ticket);
    else gridAddExact<TL_THREADS_FORCE>(0.5 * double(pairs), 0.0, result + 2, nullptr);
}

// ---- AdResS on tiles ----------------------------------------------------------------------------------------
// UpdateMolecules::update + LJ_IdealGas::run + ContributeMoleculeForceToAtoms::update for one-atom molecules
// (data::createMoleculeForEachAtom, relativeMass 1: molecule i is atom i and its centre of mass is the atom's
// position, bit for bit), as one kernel over the tiled full list.  Per stored half pair the reference adds
// +-d*ff*w to both atoms, -V_ij*grad(lambda) to both molecules and V_ij to both compensation bins
// (LJ_IdealGas.cpp:102-225); seen from both ends of a full list each row owner collects exactly its own share,
// so there are no atomics on forces and the molecule force goes straight into the atom force.
constexpr int TL_COMPENSATION_BINS = 200;  // LJ_IdealGas.hpp:54
constexpr int TL_SMEM_PER_SLOT_ADRESS = 33;  // x, y, z, lambda^mod + type byte

// stages {x, y, z, lambda^mod}; the weight is evaluated at the image position, as UpdateMolecules does for the
// reference's ghost molecules
template <bool TYPES>
__device__ __forceinline__ void stageTileAdress(const TileParams& tp, const TileDesc& td, const double4* __restrict__ pos,
                                                const mrmd_b200_weight& w, double* rec, unsigned char* sType, int tile)
{
    const int col = tile / tp.numChunks;
    const int ci = col / tp.g.n[1], cj = col % tp.g.n[1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int p = warp; p < TL_PIECES; p += blockDim.x / 32)
    {
        const int len = td.pieceLen[p];
        if (len == 0) continue;
        int ix, iy, iz;
        pieceShift(tp, ci, cj, p, ix, iy, iz);
        const double shx = double(ix) * tp.L[0], shy = double(iy) * tp.L[1], shz = double(iz) * tp.L[2];
        const int start = td.pieceStart[p], slot0 = td.pieceSlot[p];
        // TL_STAGE_UNROLL loads in flight per lane, as in stageTile
        for (int k0 = lane; k0 < len; k0 += 32 * TL_STAGE_UNROLL)
        {
            double4 raw[TL_STAGE_UNROLL];
#pragma unroll
            for (int u = 0; u < TL_STAGE_UNROLL; ++u)
                if (k0 + 32 * u < len) raw[u] = ld4nc(pos + start + k0 + 32 * u);
#pragma unroll
            for (int u = 0; u < TL_STAGE_UNROLL; ++u)
            {
                const int k = k0 + 32 * u;
                if (k >= len) break;
                const double qx = raw[u].x + shx, qy = raw[u].y + shy, qz = raw[u].z + shz;
                double* r = rec + 4 * (slot0 + k);
                r[0] = qx;
                r[1] = qy;
                r[2] = qz;
                r[3] = weightModLambda(w, qx, qy, qz);
                if (TYPES) sType[slot0 + k] = static_cast<unsigned char>(typeOf(raw[u]));
            }
        }
    }
    __syncthreads();
}

// one molecule pair of LJ_IdealGas::operator() (LJ_IdealGas.cpp:96-205) seen from the row owner alpha
template <bool SINGLE_TYPE, bool ENERGY>
__device__ __forceinline__ void adressPair(const double* rec, const unsigned char* sType, int slot, double xi, double yi,
                                           double zi, int typeI, double modA, bool cgA, bool hyA, const LJType& t0,
                                           const LJTable& table, int64_t numTypes, double rcSqr, double& fx, double& fy,
                                           double& fz, double& energy, double& vsum, double& pairs, double& activePairs)
{
    const double* q = rec + 4 * slot;
    const double modB = q[3];
    if (cgA && inCG(modB)) return;  // ideal gas, :102-107
    activePairs += 1.0;
    const double dx = xi - q[0];
    const double dy = yi - q[1];
    const double dz = zi - q[2];
    const double distSqr = distSqrExact(dx, dy, dz);
    if (distSqr > rcSqr) return;  // :137
    const LJType& t = SINGLE_TYPE ? t0 : table.t[typeI * numTypes + sType[slot]];
    double ff, e = 0.0;
    if (distSqr >= t.cappingDistanceSqr)
    {
        const double frac2 = fastRcp(distSqr);
        const double frac6 = frac2 * frac2 * frac2;
        ff = frac6 * (t.ff1 * frac6 - t.ff2) * frac2;
        // the pair energy is only needed for the drift terms of a hybrid row owner and for the returned total
        if (ENERGY || hyA) e = frac6 * (t.ef1 * frac6 - t.ef2) - t.shift;
    }
    else
        ljForceEnergy(t, distSqr, ff, e);
    const double weighting = 0.5 * (modA + modB);
    const double ffactor = ff * weighting;
    fx += dx * ffactor;
    fy += dy * ffactor;
    fz += dz * ffactor;
    if (ENERGY) energy += e * weighting;
    pairs += 1.0;
    if (hyA) vsum += 0.5 * e;  // V_ij of the drift force and of the compensation sampling, :160-200
}

// adressPair as straight-line code for the pair loop: ideal-gas pairs, pairs beyond the cutoff and entries past the end
// of a lane's row are predicated off (they read a valid record and contribute zero), the capped branch of the force law
// is left to the caller (minHi: smallest high word of the d2 seen, see ljForceTiledKernel).  Three branches per pair in
// adressPair kept the dependent FP64 chains of consecutive pairs from overlapping.
template <bool SINGLE_TYPE, bool ENERGY>
__device__ __forceinline__ void adressPairFlat(const double* rec, const unsigned char* sType, int slot, bool valid, double xi,
                                               double yi, double zi, int typeI, double modA, bool cgA, bool hyA,
                                               const LJType& t0, const LJTable& table, int64_t numTypes, double rcSqr,
                                               double& fx, double& fy, double& fz, double& energy, double& vsum, int& pairs,
                                               int& activePairs, int& minHi)
{
    const int s = valid ? slot : 0;
    const double* q = rec + 4 * s;
    const double modB = q[3];
    const bool act = valid && !(cgA && inCG(modB));  // ideal gas, :102-107
    activePairs += act ? 1 : 0;
    const double dx = xi - q[0];
    const double dy = yi - q[1];
    const double dz = zi - q[2];
    const double distSqr = distSqrExact(dx, dy, dz);
    const bool in = act && (distSqr <= rcSqr);  // :137
    const double d2 = in ? distSqr : rcSqr;
    const LJType& t = SINGLE_TYPE ? t0 : table.t[typeI * numTypes + sType[s]];
    const double frac2 = fastRcp(d2);
    const double frac6 = frac2 * frac2 * frac2;
    const double ff = frac6 * (t.ff1 * frac6 - t.ff2) * frac2;
    const double e = frac6 * (t.ef1 * frac6 - t.ef2) - t.shift;
    minHi = min(minHi, __double2hiint(d2));
    const double weighting = 0.5 * (modA + modB);
    const double ffactor = in ? ff * weighting : 0.0;
    fx += dx * ffactor;
    fy += dy * ffactor;
    fz += dz * ffactor;
    if (ENERGY) energy += in ? e * weighting : 0.0;
    pairs += in ? 1 : 0;
    vsum += (in && hyA) ? 0.5 * e : 0.0;  // V_ij of the drift force and of the compensation sampling, :160-200
}

#ifndef MRMD_ADT_PREFETCH_FORCE
#define MRMD_ADT_PREFETCH_FORCE 1
#endif
#ifndef MRMD_ADT_MORE_WORDS
#define MRMD_ADT_MORE_WORDS 3
#endif
constexpr int ADT_MORE = (MRMD_ADT_MORE_WORDS < LJT_MORE) ? MRMD_ADT_MORE_WORDS : LJT_MORE;  // rolled-loop words of the AdResS kernel
template <bool SINGLE_TYPE, bool SAMPLING, bool ENERGY>
__global__ void __launch_bounds__(TL_THREADS_FORCE, 6)
    adressForceTiledKernel(TileParams tp, AtomsView a, const int* __restrict__ desc, const int32_t* __restrict__ counts,
                           const uint16_t* __restrict__ enc, int width, LJTable table, double rcSqr, int64_t numTypes,
                           mrmd_b200_weight w, double* hist, const int32_t* __restrict__ activeTiles, double* partials,
                           double* result, unsigned int* ticket, const int* __restrict__ stop)
{
    if (stop != nullptr && *stop != 0) return;  // a step queued behind the one that asked for a rebuild
    extern __shared__ double sTile[];
    __shared__ TileDesc td;
    double energy = 0.0, pairs = 0.0, activePairs = 0.0;
    // activeTiles (step-loop drivers): the tiles the neighbour build found outside the coarse-grained bulk
    const int tile = (activeTiles != nullptr) ? activeTiles[blockIdx.x] : blockIdx.x;
    // slab weighting: when the three staged columns (padded by one more cell on each side for the drift since the
    // last sort) lie beyond the hybrid region, every pair of the tile is an ideal-gas pair
    const bool skip = tileAllCoarseGrained(tp, w, tile);
    if (!skip)
    {
        loadTileDesc(desc, td, tile);
        double* rec = sTile;
        unsigned char* sType = reinterpret_cast<unsigned char*>(rec + 4 * tp.cap);
        const int group = threadIdx.x / TL_GROUP, gl = threadIdx.x % TL_GROUP;
        // row length and first list words a pass ahead, as in ljForceTiledKernel
        const int homesPerPass = blockDim.x / TL_GROUP;
        int countNext = 0;
        uint4 wordsNext[LJT_WORDS];
        prefetchListRows(td, counts, enc, width, 0, group, gl, countNext, wordsNext);
        stageTileAdress<!SINGLE_TYPE>(tp, td, a.pos, w, rec, sType, tile);

        const LJType t0 = table.t[0];
        const int64_t T = numTypes;
        int capHi = __double2hiint(t0.cappingDistanceSqr);  // see ljForceTiledKernel
        if (!SINGLE_TYPE)
            for (int k = 1; k < MAX_LJ_TYPES * MAX_LJ_TYPES; ++k) capHi = max(capHi, __double2hiint(table.t[k].cappingDistanceSqr));
        const double inverseBinSize = 1.0 / ((1.0 - 0.0) / double(TL_COMPENSATION_BINS));
        for (int hBase = 0; hBase < td.homeCount; hBase += homesPerPass)
        {
            const int h = hBase + group;
            const bool active = h < td.homeCount;
            const int i = td.homeStart + h;
            const int selfSlot = active ? td.selfSlot0 + h : 0;
            const double xi = rec[4 * selfSlot], yi = rec[4 * selfSlot + 1], zi = rec[4 * selfSlot + 2];
            const int typeI = SINGLE_TYPE ? 0 : sType[selfSlot];
            double lambda, modA, gx, gy, gz;
            weightEval(w, xi, yi, zi, lambda, modA, gx, gy, gz);
            const bool hyA = inHY(modA), cgA = inCG(modA);
            double fx = 0.0, fy = 0.0, fz = 0.0, vsum = 0.0, ePass = 0.0;
            int pPass = 0, aPass = 0, minHi = 0x7fffffff;
            // the force the row owner adds to (zero, or what the thermodynamic force left) is fetched now, not at the store
            double fOld[TL_VPL];
#pragma unroll
            for (int k = 0; k < TL_VPL; ++k)
            {
                const int comp = groupSumValue(gl, k);
                fOld[k] = 0.0;
#if MRMD_ADT_PREFETCH_FORCE
                if (active && comp >= 0 && comp < 3)
                    fOld[k] = ((comp == 0) ? a.force[0] : ((comp == 1) ? a.force[1] : a.force[2]))[i];
#endif
            }
            const int numNeighbors = min(countNext, width);
            unsigned words[4 * (LJT_WORDS + ADT_MORE)];
#pragma unroll
            for (int k = 0; k < LJT_WORDS; ++k)
            {
                words[4 * k] = wordsNext[k].x;
                words[4 * k + 1] = wordsNext[k].y;
                words[4 * k + 2] = wordsNext[k].z;
                words[4 * k + 3] = wordsNext[k].w;
            }
            const uint16_t* mineRow = enc + size_t(active ? i : 0) * width + gl * (width / TL_GROUP);
            // the words behind the unrolled steps are loaded now and used after them
#pragma unroll
            for (int k = LJT_WORDS; k < LJT_WORDS + ADT_MORE; ++k)
            {
                const uint4 wk = reinterpret_cast<const uint4*>(mineRow)[k];
                words[4 * k] = wk.x;
                words[4 * k + 1] = wk.y;
                words[4 * k + 2] = wk.z;
                words[4 * k + 3] = wk.w;
            }
            if (hBase + homesPerPass < td.homeCount)  // block uniform
                prefetchListRows(td, counts, enc, width, hBase + homesPerPass, group, gl, countNext, wordsNext);
            const int mine = (numNeighbors - gl + TL_GROUP - 1) / TL_GROUP;
            const int iters = __reduce_max_sync(0xffffffffu, (numNeighbors + TL_GROUP - 1) / TL_GROUP);
            if (iters > 0)  // warp uniform
            {
#pragma unroll
                for (int it = 0; it < 8 * LJT_WORDS; ++it)
                {
                    const int slot = (words[it >> 1] >> (16 * (it & 1))) & 0xffffu;
                    adressPairFlat<SINGLE_TYPE, ENERGY>(rec, sType, slot, it < mine, xi, yi, zi, typeI, modA, cgA, hyA, t0, table, T,
                                                        rcSqr, fx, fy, fz, ePass, vsum, pPass, aPass, minHi);
                }
            }
            if (ADT_MORE > 0)
            {
                const int more = min(iters, 8 * (LJT_WORDS + ADT_MORE));
#pragma unroll 1
                for (int it = 8 * LJT_WORDS; it < more; it += 2)
                {
                    const unsigned wq = words[4 * LJT_WORDS];
#pragma unroll
                    for (int k = 4 * LJT_WORDS; k + 1 < 4 * (LJT_WORDS + ADT_MORE); ++k) words[k] = words[k + 1];
                    adressPairFlat<SINGLE_TYPE, ENERGY>(rec, sType, wq & 0xffffu, it < mine, xi, yi, zi, typeI, modA, cgA, hyA, t0,
                                                        table, T, rcSqr, fx, fy, fz, ePass, vsum, pPass, aPass, minHi);
                    adressPairFlat<SINGLE_TYPE, ENERGY>(rec, sType, wq >> 16, it + 1 < mine, xi, yi, zi, typeI, modA, cgA, hyA, t0,
                                                        table, T, rcSqr, fx, fy, fz, ePass, vsum, pPass, aPass, minHi);
                }
            }
            for (int it = 8 * (LJT_WORDS + ADT_MORE); it < iters; ++it)  // the entries behind the prefetched words
            {
                const bool valid = it < mine;
                adressPairFlat<SINGLE_TYPE, ENERGY>(rec, sType, valid ? mineRow[it] : 0, valid, xi, yi, zi, typeI, modA, cgA, hyA,
                                                    t0, table, T, rcSqr, fx, fy, fz, ePass, vsum, pPass, aPass, minHi);
            }
            // a pair inside the capping distance: the warp evaluates its rows again with the reference's branches
            {
                double eP = ePass, pP = double(pPass), aP = double(aPass);
                if (__any_sync(0xffffffffu, minHi <= capHi))
                {
                    fx = fy = fz = vsum = eP = pP = aP = 0.0;
                    for (int it = 0; it < mine; ++it)
                        adressPair<SINGLE_TYPE, ENERGY>(rec, sType, mineRow[it], xi, yi, zi, typeI, modA, cgA, hyA, t0, table, T,
                                                        rcSqr, fx, fy, fz, eP, vsum, pP, aP);
                }
                energy += eP;
                pairs += pP;
                activePairs += aP;
            }
            // three lanes (slots) of the group end up with the x / y / z total, a fourth with sum(V_ij), which the storing
            // lanes fetch from it
            double f[TL_VPL];
            groupSum4(fx, fy, fz, vsum, gl, f);
            vsum = __shfl_sync(0xffffffffu, f[groupSumSlot(3)], (threadIdx.x & 31) - gl + groupSumLane(3));
            // molecule force: drift force -sum(V_ij) grad(lambda) (:163-169) plus the drift compensation
            // mean[bin] grad(lambda) (:209-222); with one atom of relativeMass 1 per molecule
            // ContributeMoleculeForceToAtoms adds it to the atom unchanged
            double scale = 0.0;
            if (active && hyA)
            {
                scale = -vsum;
                const long long bin = histBin(0.0, inverseBinSize, TL_COMPENSATION_BINS, lambda);
                if (bin != -1)
                {
                    scale += hist[2 * TL_COMPENSATION_BINS * T + bin * T + typeI];
                    if (SAMPLING && gl == 0)
                    {
                        atomicAdd(hist + bin * T + typeI, vsum);
                        atomicAdd(hist + TL_COMPENSATION_BINS * T + bin * T + typeI, 1.0);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < TL_VPL; ++k)
            {
                const int comp = groupSumValue(gl, k);
                if (active && comp >= 0 && comp < 3)
                {
                    double out = f[k];
                    if (hyA) out += scale * ((comp == 0) ? gx : ((comp == 1) ? gy : gz));
                    if (out != 0.0)
                    {
                        double* plane = (comp == 0) ? a.force[0] : ((comp == 1) ? a.force[1] : a.force[2]);
#if MRMD_ADT_PREFETCH_FORCE
                        plane[i] = fOld[k] + out;
#else
                        plane[i] += out;
#endif
                    }
                }
            }
        }
    }
    // every pair is visited from both sides
    if (ENERGY) gridReduce3<TL_THREADS_FORCE>(0.5 * energy, 0.5 * pairs, 0.5 * activePairs, partials, result, ticket);
    else gridAddExact<TL_THREADS_FORCE>(0.5 * pairs, 0.5 * activePairs, result + 1, result + 2);
}

// ---- AdResS on tiles for molecules of NA atoms (the tetramers of BASELINE.json configs[3]) --------------------------
// The tiled list is built on the molecules' centres of mass (mrmd_b200_verlet_build_periodic_molecules: the same builder
// on the cell-sorted centres of mass); a staged slot is a molecule: its NA atom positions and lambda^mod at its centre of
// mass (images: positions + the image shift, the weight at the shifted centre of mass).  NA lanes share one home
// molecule alpha, lane i owns atom i of alpha: per partner molecule it evaluates its atom against the partner's NA atoms
// (LJ_IdealGas.cpp:96-205 seen from the row owner), sum(V_ij) is combined over the lanes for the drift force
// -sum(V_ij) grad(lambda_alpha) and the drift compensation (:209-222), and ContributeMoleculeForceToAtoms hands the
// molecule force to the atoms by relative mass.  Full list: no atomics on forces, no ghost molecules, no reverse halo.
#ifndef MRMD_TL_THREADS_MOL
#define MRMD_TL_THREADS_MOL 128
#endif
constexpr int TL_THREADS_MOL = MRMD_TL_THREADS_MOL;
template <int NA>
constexpr int molSlotDoubles() { return 3 * NA + 1; }
template <int NA>
constexpr int molSlotBytes() { return molSlotDoubles<NA>() * 8 + NA; }  // + one type byte per atom

template <int NA, bool SINGLE_TYPE, bool SAMPLING, bool ENERGY>
__global__ void __launch_bounds__(TL_THREADS_MOL)
    moleculeForceTiledKernel(TileParams tp, AtomsView a, const double4* __restrict__ com, const int* __restrict__ desc,
                             const int32_t* __restrict__ counts, const uint16_t* __restrict__ enc, int width, LJTable table,
                             double rcSqr, int64_t numTypes, mrmd_b200_weight w, double* hist, int maxHomes,
                             const int32_t* __restrict__ activeTiles, double* partials, double* result, unsigned int* ticket)
{
    static_assert(NA == 2 || NA == 4 || NA == 8, "lanes per molecule: a power of two");
    constexpr int REC = molSlotDoubles<NA>();
    extern __shared__ double sTile[];
    __shared__ TileDesc td;
    double energy = 0.0;
    int pairs = 0, activePairs = 0;
    const int tile = (activeTiles != nullptr) ? activeTiles[blockIdx.x] : blockIdx.x;
    if (!tileAllCoarseGrained(tp, w, tile))
    {
        loadTileDesc(desc, td, tile);
        double* rec = sTile;
        // behind the slots: the list rows of the tile's home molecules (copied with 16-byte loads, read as broadcasts by
        // the lanes of a molecule), then the type bytes
        uint16_t* sRows = reinterpret_cast<uint16_t*>(rec + size_t(REC) * tp.cap);
        unsigned char* sType = reinterpret_cast<unsigned char*>(sRows + size_t(maxHomes) * width);
        {
            const int wordsPerRow = width / 8;
            const uint4* src = reinterpret_cast<const uint4*>(enc + size_t(td.homeStart) * width);
            uint4* dst = reinterpret_cast<uint4*>(sRows);
            const int homesHere = min(td.homeCount, maxHomes);
            // (the loads first: several in flight per thread)
            for (int e0 = threadIdx.x; e0 < homesHere * wordsPerRow; e0 += TL_STAGE_UNROLL * blockDim.x)
            {
                uint4 v[TL_STAGE_UNROLL];
#pragma unroll
                for (int u = 0; u < TL_STAGE_UNROLL; ++u)
                    if (e0 + u * blockDim.x < homesHere * wordsPerRow) v[u] = src[e0 + u * blockDim.x];
#pragma unroll
                for (int u = 0; u < TL_STAGE_UNROLL; ++u)
                    if (e0 + u * blockDim.x < homesHere * wordsPerRow) dst[e0 + u * blockDim.x] = v[u];
            }
        }
        {
            // staging: one piece per warp, a lane per (molecule, atom)
            const int col = tile / tp.numChunks;
            const int ci = col / tp.g.n[1], cj = col % tp.g.n[1];
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            for (int p = warp; p < TL_PIECES; p += blockDim.x / 32)
            {
                const int len = td.pieceLen[p];
                if (len == 0) continue;
                int ix, iy, iz;
                pieceShift(tp, ci, cj, p, ix, iy, iz);
                const double shx = double(ix) * tp.L[0], shy = double(iy) * tp.L[1], shz = double(iz) * tp.L[2];
                const int start = td.pieceStart[p], slot0 = td.pieceSlot[p];
                // TL_STAGE_UNROLL loads in flight per lane (as in stageTile): the staging phase of this kernel was its
                // largest stall (long scoreboard + the barrier behind it, profiles/r02_molecule_force_ncu.txt)
                for (int e0 = lane; e0 < len * NA; e0 += 32 * TL_STAGE_UNROLL)
                {
                    double4 raw[TL_STAGE_UNROLL], cm[TL_STAGE_UNROLL];
#pragma unroll
                    for (int u = 0; u < TL_STAGE_UNROLL; ++u)
                    {
                        const int e = e0 + 32 * u;
                        if (e < len * NA)
                        {
                            raw[u] = ld4nc(a.pos + size_t(start) * NA + e);
                            if (e % NA == 0) cm[u] = ld4nc(com + start + e / NA);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < TL_STAGE_UNROLL; ++u)
                    {
                        const int e = e0 + 32 * u;
                        if (e >= len * NA) break;
                        const int k = e / NA, j = e % NA;
                        double* r = rec + size_t(REC) * (slot0 + k);
                        r[3 * j] = raw[u].x + shx;
                        r[3 * j + 1] = raw[u].y + shy;
                        r[3 * j + 2] = raw[u].z + shz;
                        if (!SINGLE_TYPE) sType[(slot0 + k) * NA + j] = static_cast<unsigned char>(typeOf(raw[u]));
                        if (j == 0) r[3 * NA] = weightModLambda(w, cm[u].x + shx, cm[u].y + shy, cm[u].z + shz);
                    }
                }
            }
            __syncthreads();
        }
        const int group = threadIdx.x / NA, li = threadIdx.x % NA;
        const int64_t T = numTypes;
        const LJType t0 = table.t[0];
        const double inverseBinSize = 1.0 / ((1.0 - 0.0) / double(TL_COMPENSATION_BINS));
        for (int hBase = 0; hBase < td.homeCount; hBase += blockDim.x / NA)
        {
            const int h = hBase + group;
            const bool active = h < td.homeCount;
            const int u = td.homeStart + (active ? h : 0);  // molecule
            const int selfSlot = td.selfSlot0 + (active ? h : 0);
            const double* me = rec + size_t(REC) * selfSlot + 3 * li;
            const double xi = me[0], yi = me[1], zi = me[2];
            const int typeI = SINGLE_TYPE ? 0 : sType[selfSlot * NA + li];
            const double4 c = ld4nc(com + u);
            double lambda, modA, gx, gy, gz;
            weightEval(w, c.x, c.y, c.z, lambda, modA, gx, gy, gz);
            const bool hyA = inHY(modA), cgA = inCG(modA);
            double fx = 0.0, fy = 0.0, fz = 0.0, vsum = 0.0;
            const int numNeighbors = active ? min(counts[u], width) : 0;
            // rows of homes beyond the staged ones (a tile with more homes than any tile of the build) come from global
            const uint16_t* row = (h < maxHomes) ? sRows + size_t(active ? h : 0) * width : enc + size_t(u) * width;
            const int iters = __reduce_max_sync(0xffffffffu, numNeighbors);
            for (int n = 0; n < iters; ++n)
            {
                if (n >= numNeighbors) continue;
                const int slot = row[tiledRowIndex(n, width)];
                const double* q = rec + size_t(REC) * slot;
                const double modB = q[3 * NA];
                if (cgA && inCG(modB)) continue;  // ideal gas, LJ_IdealGas.cpp:102-107
                if (li == 0) activePairs += 1;
                const double weighting = 0.5 * (modA + modB);
                // the NA partner atoms are evaluated predicated, not branched: their FP64 chains interleave
#pragma unroll
                for (int j = 0; j < NA; ++j)
                {
                    const double dx = xi - q[3 * j];
                    const double dy = yi - q[3 * j + 1];
                    const double dz = zi - q[3 * j + 2];
                    const double distSqr = distSqrExact(dx, dy, dz);
                    const bool in = distSqr <= rcSqr;  // :137 skips distSqr > rcSqr
                    const LJType& t = SINGLE_TYPE ? t0 : table.t[typeI * T + sType[slot * NA + j]];
                    const double d2 = in ? distSqr : rcSqr;  // keeps the arithmetic of a skipped pair finite
                    double ff, e;
                    if (d2 >= t.cappingDistanceSqr)
                    {
                        const double frac2 = fastRcp(d2);
                        const double frac6 = frac2 * frac2 * frac2;
                        ff = frac6 * (t.ff1 * frac6 - t.ff2) * frac2;
                        e = frac6 * (t.ef1 * frac6 - t.ef2) - t.shift;
                    }
                    else
                        ljForceEnergy(t, d2, ff, e);
                    const double ffactor = in ? ff * weighting : 0.0;
                    fx += dx * ffactor;
                    fy += dy * ffactor;
                    fz += dz * ffactor;
                    pairs += in ? 1 : 0;
                    if (ENERGY) energy += in ? e * weighting : 0.0;
                    if (hyA) vsum += in ? 0.5 * e : 0.0;  // V_ij of the drift force and of the compensation sampling, :160-200
                }
            }
            // drift force -sum(V_ij) grad(lambda) (:163-169) plus the drift compensation mean[bin] grad(lambda) of the
            // molecule's FIRST atom type (:209-222); the molecule force goes to the atoms by relative mass
            // (ContributeMoleculeForceToAtoms.cpp:34-45)
            double vAll = vsum;
#pragma unroll
            for (int o = NA / 2; o > 0; o >>= 1) vAll += __shfl_xor_sync(0xffffffffu, vAll, o);
            if (active)
            {
                const int64_t ai = int64_t(u) * NA + li;
                if (hyA)
                {
                    double scale = -vAll;
                    const long long bin = histBin(0.0, inverseBinSize, TL_COMPENSATION_BINS, lambda);
                    if (bin != -1)
                    {
                        scale += hist[2 * TL_COMPENSATION_BINS * T + bin * T + (SINGLE_TYPE ? 0 : sType[selfSlot * NA])];
                        if (SAMPLING)
                        {
                            atomicAdd(hist + bin * T + typeI, vsum);
                            atomicAdd(hist + TL_COMPENSATION_BINS * T + bin * T + typeI, 1.0);
                        }
                    }
                    const double rm = a.relMass[ai];
                    fx += rm * (scale * gx);
                    fy += rm * (scale * gy);
                    fz += rm * (scale * gz);
                }
                if (fx != 0.0 || fy != 0.0 || fz != 0.0)
                {
                    a.force[0][ai] += fx;
                    a.force[1][ai] += fy;
                    a.force[2][ai] += fz;
                }
            }
        }
    }
    // every pair is visited from both sides
    if (ENERGY) gridReduce3<TL_THREADS_MOL>(0.5 * energy, 0.5 * double(pairs), 0.5 * double(activePairs), partials, result, ticket);
    else gridAddExact<TL_THREADS_MOL>(0.5 * double(pairs), 0.5 * double(activePairs), result + 1, result + 2);
}

// decode the 16-bit slots back to (local partner index, image shift code) in Cabana's row-major layout
__global__ void __launch_bounds__(TL_THREADS_BUILD)
    decodeTiledKernel(TileParams tp, const int* __restrict__ desc, const int32_t* __restrict__ counts,
                      const uint16_t* __restrict__ enc, int width, int32_t* partner, int32_t* shiftCode)
{
    __shared__ TileDesc td;
    loadTileDesc(desc, td, blockIdx.x);
    const int col = blockIdx.x / tp.numChunks;
    const int ci = col / tp.g.n[1], cj = col % tp.g.n[1];
    for (int h = threadIdx.x; h < td.homeCount; h += blockDim.x)
    {
        const int i = td.homeStart + h;
        const int cnt = min(counts[i], width);
        for (int n = 0; n < width; ++n)
        {
            int j = -1, code = -1;
            if (n < cnt)
            {
                const int slot = enc[size_t(i) * width + tiledRowIndex(n, width)];
                int p = 0;
                while (p + 1 < TL_PIECES && td.pieceSlot[p + 1] <= slot) ++p;
                j = td.pieceStart[p] + (slot - td.pieceSlot[p]);
                int sx, sy, sz;
                pieceShift(tp, ci, cj, p, sx, sy, sz);
                code = (sx + 1) + 3 * (sy + 1) + 9 * (sz + 1);
            }
            partner[size_t(i) * width + n] = j;
            shiftCode[size_t(i) * width + n] = code;
        }
    }
}

// cellLo / cellHi over the extended grid: local cells from the linked-cell prefix array, halo columns (x-slab
// decomposition) from the per-column prefix arrays of the received halo atoms
__global__ void extCellRangesKernel(const int32_t* __restrict__ localStart, int64_t numLocalCells,
                                    const int32_t* __restrict__ haloLeft, const int32_t* __restrict__ haloRight,
                                    int64_t cellsPerColumnX, int32_t* cellLo, int32_t* cellHi)
{
    const int64_t c = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    const int64_t haloCells = (haloLeft != nullptr) ? cellsPerColumnX : 0;
    if (c >= numLocalCells + 2 * haloCells) return;
    int32_t lo, hi;
    if (c < haloCells) { lo = haloLeft[c]; hi = haloLeft[c + 1]; }
    else if (c < haloCells + numLocalCells) { lo = localStart[c - haloCells]; hi = localStart[c - haloCells + 1]; }
    else { lo = haloRight[c - haloCells - numLocalCells]; hi = haloRight[c - haloCells - numLocalCells + 1]; }
    cellLo[c] = lo;
    cellHi[c] = hi;
}

static int makeTileParams(const mrmd_b200_atoms* a, const mrmd_b200_subdomain* s, int CH, int cap, int haloX, int R,
                          TileParams& tp)
{
    tp.g = a->lcGrid;
    tp.R = R;
    tp.haloX = haloX;
    tp.CH = CH;
    tp.numChunks = (tp.g.n[2] + CH - 1) / CH;
    tp.cap = cap;
    for (int d = 0; d < 3; ++d)
    {
        tp.periodic[d] = s->ghostLayerThickness[d] > 0.0 ? 1 : 0;
        tp.L[d] = s->diameter[d];
        tp.minInner[d] = s->minInnerCorner[d];
        tp.maxInner[d] = s->maxInnerCorner[d];
    }
    return 0;
}

constexpr int TL_SMEM_PER_SLOT_BUILD = 24;  // x, y, z (the builder adds TL_BUILD_BATCH spare records)
constexpr int TL_SMEM_PER_SLOT_FORCE = 25;  // x, y, z + type byte
constexpr int TL_SMEM_BUDGET = 48 * 1024;   // preferred budget of a tile's staged positions (several tiles per SM)
constexpr int TL_SMEM_MAX = 200 * 1024;

int tiledConfigure()
{
    static bool done = false;
    if (done) return 0;
    // the tiles live in shared memory and barely use L1: ask for the largest carveout so that residency is set by
    // registers, not by the default shared-memory split
#define TL_SET(K)                                                                                    \
    MB_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, TL_SMEM_MAX));     \
    MB_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared))
    TL_SET((verletBuildTiledKernel<true>));
    TL_SET((verletBuildTiledKernel<false>));
    TL_SET((ljForceTiledKernel<true, true, true>));
    TL_SET((ljForceTiledKernel<true, true, false>));
    TL_SET((ljForceTiledKernel<true, false, true>));
    TL_SET((ljForceTiledKernel<true, false, false>));
    TL_SET((ljForceTiledKernel<false, true, true>));
    TL_SET((ljForceTiledKernel<false, true, false>));
    TL_SET((ljForceTiledKernel<false, false, true>));
    TL_SET((ljForceTiledKernel<false, false, false>));
    TL_SET((adressForceTiledKernel<true, true, true>));
    TL_SET((adressForceTiledKernel<true, true, false>));
    TL_SET((adressForceTiledKernel<true, false, true>));
    TL_SET((adressForceTiledKernel<true, false, false>));
    TL_SET((adressForceTiledKernel<false, true, true>));
    TL_SET((adressForceTiledKernel<false, true, false>));
    TL_SET((adressForceTiledKernel<false, false, true>));
    TL_SET((adressForceTiledKernel<false, false, false>));
    TL_SET((moleculeForceTiledKernel<4, true, true, true>));
    TL_SET((moleculeForceTiledKernel<4, true, true, false>));
    TL_SET((moleculeForceTiledKernel<4, true, false, true>));
    TL_SET((moleculeForceTiledKernel<4, true, false, false>));
    TL_SET((moleculeForceTiledKernel<4, false, true, true>));
    TL_SET((moleculeForceTiledKernel<4, false, true, false>));
    TL_SET((moleculeForceTiledKernel<4, false, false, true>));
    TL_SET((moleculeForceTiledKernel<4, false, false, false>));
#undef TL_SET
    done = true;
    return 0;
}

int ljApplyTiled(mrmd_b200_lj* lj, mrmd_b200_atoms* a, const mrmd_b200_verlet* v, bool accumulate, bool energy,
                 cudaStream_t st, const int* stop)
{
    MB_REQUIRE(a->lcValid && a->lcEpoch == v->tiledEpoch, "lj_apply: the atoms were re-sorted after this tiled list was built");
    MB_REQUIRE(!v->half, "lj_apply: tiled lists are full lists");
    MB_TRY(tiledConfigure());
    TileParams tp;
    MB_TRY(makeTileParams(a, &v->tiledSub, v->tiledCH, v->tiledSlots, v->tiledHaloX, v->tiledR, tp));
    const int tiles = tp.g.n[0] * tp.g.n[1] * tp.numChunks;
    MB_TRY(lj->partials.reserve(size_t(tiles) * 3 * 8));
    if (energy) MB_CUDA(cudaMemsetAsync(lj->dResult, 0, 24, st));  // without energy only the running sums move (gridAddExact)
    const size_t smem = size_t(v->tiledSlots) * TL_SMEM_PER_SLOT_FORCE + 16;
    const bool single = (lj->numTypes == 1);
#define LJT_LAUNCH(S1, ACC, EN)                                                                                       \
    ljForceTiledKernel<S1, ACC, EN><<<tiles, TL_THREADS_FORCE, smem, st>>>(                                           \
        tp, a->v, v->tileDesc.as<int>(), v->counts.as<int32_t>(), v->enc.as<uint16_t>(), static_cast<int>(v->width), \
        lj->table, lj->rcSqr, lj->numTypesQuirk, lj->partials.as<double>(), lj->dResult, lj->dTicket, stop)
    if (single)
    {
        if (accumulate) { if (energy) LJT_LAUNCH(true, true, true); else LJT_LAUNCH(true, true, false); }
        else { if (energy) LJT_LAUNCH(true, false, true); else LJT_LAUNCH(true, false, false); }
    }
    else
    {
        if (accumulate) { if (energy) LJT_LAUNCH(false, true, true); else LJT_LAUNCH(false, true, false); }
        else { if (energy) LJT_LAUNCH(false, false, true); else LJT_LAUNCH(false, false, false); }
    }
#undef LJT_LAUNCH
    MB_LAUNCHED();
    return 0;
}

// the force kernel of mrmd_b200_adress_run_periodic (adress.cu owns the run counter and the histogram update)
int adressApplyTiled(mrmd_b200_adress* ad, mrmd_b200_atoms* a, const mrmd_b200_verlet* v, const mrmd_b200_weight* w,
                     bool sampling, bool energy, cudaStream_t st, const int* stop)
{
    MB_REQUIRE(a->lcValid && a->lcEpoch == v->tiledEpoch,
               "adress_run_periodic: the atoms were re-sorted after this tiled list was built");
    MB_REQUIRE(!v->half, "adress_run_periodic: tiled lists are full lists");
    MB_REQUIRE(ad->numTypes <= 255, "adress_run_periodic: more than 255 atom types");
    // the row owner evaluates lambda at its own position and at the partner's IMAGE position; that equals the
    // reference's half-list evaluation only if images near a periodic boundary are coarse grained like their
    // originals: the AT + HY region has to stay a list radius away from the periodic faces it depends on
    {
        const mrmd_b200_subdomain& s = v->tiledSub;
        const int axes = (w->kind == MRMD_B200_WEIGHT_SLAB) ? 1 : 3;
        const double reach = (w->kind == MRMD_B200_WEIGHT_SLAB) ? 0.5 * w->atRegion + w->hyRegion : w->atRegion + w->hyRegion;
        for (int d = 0; d < axes; ++d)
        {
            if (s.ghostLayerThickness[d] == 0.0 || (d == 0 && v->tiledHaloX)) continue;
            MB_REQUIRE(w->center[d] - reach >= s.minCorner[d] + s.ghostLayerThickness[d] &&
                           w->center[d] + reach <= s.maxCorner[d] - s.ghostLayerThickness[d],
                       "adress_run_periodic: the AT/HY region reaches a periodic boundary (use the generic path)");
        }
    }
    MB_TRY(tiledConfigure());
    TileParams tp;
    MB_TRY(makeTileParams(a, &v->tiledSub, v->tiledCH, v->tiledSlots, v->tiledHaloX, v->tiledR, tp));
    // the step-loop drivers' lists know their tiles outside the coarse-grained bulk: launch over those only
    const int tiles = v->tiledCgSkip ? v->numActiveTiles : tp.g.n[0] * tp.g.n[1] * tp.numChunks;
    const int32_t* activeTiles = v->tiledCgSkip ? v->activeTiles.as<int32_t>() : nullptr;
    MB_TRY(ad->partials.reserve(size_t(std::max(tiles, 1)) * 3 * 8));
    if (energy) MB_CUDA(cudaMemsetAsync(ad->dResult, 0, 24, st));  // without energy only the running sums move
    if (tiles == 0) return 0;
    const size_t smem = size_t(v->tiledSlots) * TL_SMEM_PER_SLOT_ADRESS + 16;
    MB_REQUIRE(smem <= size_t(TL_SMEM_MAX), "adress_run_periodic: a tile exceeds shared memory");
    const bool single = (ad->numTypes == 1);
#define ADT_LAUNCH(S1, SAMP, EN)                                                                                      \
    adressForceTiledKernel<S1, SAMP, EN><<<tiles, TL_THREADS_FORCE, smem, st>>>(                                      \
        tp, a->v, v->tileDesc.as<int>(), v->counts.as<int32_t>(), v->enc.as<uint16_t>(), static_cast<int>(v->width), \
        ad->table, ad->rcSqr, ad->numTypes, *w, ad->hist, activeTiles, ad->partials.as<double>(), ad->dResult, ad->dTicket, \
        stop)
    if (single)
    {
        if (sampling) { if (energy) ADT_LAUNCH(true, true, true); else ADT_LAUNCH(true, true, false); }
        else { if (energy) ADT_LAUNCH(true, false, true); else ADT_LAUNCH(true, false, false); }
    }
    else
    {
        if (sampling) { if (energy) ADT_LAUNCH(false, true, true); else ADT_LAUNCH(false, true, false); }
        else { if (energy) ADT_LAUNCH(false, false, true); else ADT_LAUNCH(false, false, false); }
    }
#undef ADT_LAUNCH
    MB_LAUNCHED();
    return 0;
}

// LJ_IdealGas for molecules of atomsPerMolecule consecutive atoms over a tiled list of their centres of mass
// (mrmd_b200_adress_run_periodic_molecules; adress.cu owns the run counter and the histogram update)
int moleculeApplyTiled(mrmd_b200_adress* ad, const mrmd_b200_molecules* m, mrmd_b200_atoms* a, const mrmd_b200_verlet* v,
                       const mrmd_b200_weight* w, int atomsPerMolecule, bool sampling, bool energy, cudaStream_t st)
{
    MB_REQUIRE(atomsPerMolecule == 4, "adress_run_periodic_molecules: four atoms per molecule");
    const mrmd_b200_atoms* lv = m->lcView;
    MB_REQUIRE(lv != nullptr && lv->lcValid && lv->lcEpoch == v->tiledEpoch,
               "adress_run_periodic_molecules: the molecules were re-sorted after this tiled list was built");
    MB_REQUIRE(!v->half, "adress_run_periodic_molecules: tiled lists are full lists");
    MB_REQUIRE(ad->numTypes <= 255, "adress_run_periodic_molecules: more than 255 atom types");
    {
        // as adressApplyTiled: images near a periodic boundary must be coarse grained like their originals
        const mrmd_b200_subdomain& s = v->tiledSub;
        const int axes = (w->kind == MRMD_B200_WEIGHT_SLAB) ? 1 : 3;
        const double reach = (w->kind == MRMD_B200_WEIGHT_SLAB) ? 0.5 * w->atRegion + w->hyRegion : w->atRegion + w->hyRegion;
        for (int d = 0; d < axes; ++d)
        {
            if (s.ghostLayerThickness[d] == 0.0 || (d == 0 && v->tiledHaloX)) continue;
            MB_REQUIRE(w->center[d] - reach >= s.minCorner[d] + s.ghostLayerThickness[d] &&
                           w->center[d] + reach <= s.maxCorner[d] - s.ghostLayerThickness[d],
                       "adress_run_periodic_molecules: the AT/HY region reaches a periodic boundary (use the generic path)");
        }
    }
    MB_TRY(tiledConfigure());
    TileParams tp;
    MB_TRY(makeTileParams(lv, &v->tiledSub, v->tiledCH, v->tiledSlots, v->tiledHaloX, v->tiledR, tp));
    const int tiles = v->tiledCgSkip ? v->numActiveTiles : tp.g.n[0] * tp.g.n[1] * tp.numChunks;
    const int32_t* activeTiles = v->tiledCgSkip ? v->activeTiles.as<int32_t>() : nullptr;
    MB_TRY(ad->partials.reserve(size_t(std::max(tiles, 1)) * 3 * 8));
    if (energy) MB_CUDA(cudaMemsetAsync(ad->dResult, 0, 24, st));  // without energy only the running sums move
    if (tiles == 0) return 0;
    // homes whose list rows are staged: the tallest tile holds about twice the aimed-at number (the rest reads global)
    const int maxHomes = std::min(2 * v->tiledTargetHomes + 16, 64);
    const size_t smem = size_t(v->tiledSlots) * molSlotBytes<4>() + size_t(maxHomes) * size_t(v->width) * 2 + 16;
    MB_REQUIRE(smem <= size_t(TL_SMEM_MAX), "adress_run_periodic_molecules: a tile exceeds shared memory");
    const bool single = (ad->numTypes == 1);
#define MOL_LAUNCH(S1, SAMP, EN)                                                                                         \
    moleculeForceTiledKernel<4, S1, SAMP, EN><<<tiles, TL_THREADS_MOL, smem, st>>>(                                      \
        tp, a->v, m->v.pos, v->tileDesc.as<int>(), v->counts.as<int32_t>(), v->enc.as<uint16_t>(), static_cast<int>(v->width), \
        ad->table, ad->rcSqr, ad->numTypes, *w, ad->hist, maxHomes, activeTiles, ad->partials.as<double>(), ad->dResult,    \
        ad->dTicket)
    if (single)
    {
        if (sampling) { if (energy) MOL_LAUNCH(true, true, true); else MOL_LAUNCH(true, true, false); }
        else { if (energy) MOL_LAUNCH(true, false, true); else MOL_LAUNCH(true, false, false); }
    }
    else
    {
        if (sampling) { if (energy) MOL_LAUNCH(false, true, true); else MOL_LAUNCH(false, true, false); }
        else { if (energy) MOL_LAUNCH(false, false, true); else MOL_LAUNCH(false, false, false); }
    }
#undef MOL_LAUNCH
    MB_LAUNCHED();
    return 0;
}
}  // namespace mrmd_b200

namespace mrmd_b200
{
// haloLeft / haloRight (device, ny * nz + 1 absolute prefix indices each, or nullptr): x-slab decomposition, the
// atoms received from the left / right neighbour rank sit behind the local atoms in (j, k) cell order
int verletBuildTiled(mrmd_b200_verlet* v, const mrmd_b200_atoms* a, const mrmd_b200_subdomain* s, double radius,
                     double cellRatio, int64_t maxNeigh, const int32_t* haloLeft, const int32_t* haloRight,
                     cudaStream_t st)
{
    const int haloX = (haloLeft != nullptr && haloRight != nullptr) ? 1 : 0;
    MB_REQUIRE(v != nullptr && a != nullptr && s != nullptr, "verlet_build_periodic");
    MB_REQUIRE(radius > 0.0 && cellRatio > 0.0 && maxNeigh > 0, "verlet_build_periodic: bad radius / ratio / width");
    MB_REQUIRE(a->lcValid && a->lcBegin == 0 && a->lcEnd == a->numLocal,
               "verlet_build_periodic: sort the local atoms first (LinkedCellList + permute over [0, numLocalAtoms))");
    MB_REQUIRE(a->lcPosEpoch == a->posEpoch,
               "verlet_build_periodic: positions changed since the last LinkedCellList + permute (the atoms no longer sit in "
               "the cells of the sort): sort again before building");
    const GridDev& g = a->lcGrid;
    // cells are at least one radius wide along x and y (3 x 3 columns around a tile); along z they may be finer: a
    // LinkedCellList with gridDelta = (r, r, r / 4) keeps the atoms of a column in finer z order and lets the
    // builder scan only the z interval the cutoff sphere cuts out of each column
    const int R = static_cast<int>(std::ceil(radius / g.dx[2] - 1e-12));
    MB_REQUIRE(R >= 1 && R <= TL_MAX_R, "verlet_build_periodic: linked cells along z finer than radius / 4");
    for (int d = 0; d < 3; ++d)
    {
        const int reach = (d == 2) ? R : 1;
        MB_REQUIRE(g.min[d] == s->minCorner[d] && std::fabs(g.dx[d] * g.n[d] - s->diameter[d]) <= 1e-9 * s->diameter[d],
                   "verlet_build_periodic: the linked-cell grid must span the subdomain");
        MB_REQUIRE(g.dx[d] * reach >= radius, "verlet_build_periodic: linked cells smaller than the list radius");
        MB_REQUIRE(g.n[d] >= 2 * reach + 1 || s->ghostLayerThickness[d] == 0.0 || (d == 0 && haloX),
                   "verlet_build_periodic: a periodic axis is shorter than three list radii");
        MB_REQUIRE(s->ghostLayerThickness[d] == 0.0 || s->ghostLayerThickness[d] <= g.dx[d] * reach,
                   "verlet_build_periodic: ghost layer thicker than the cells the list radius spans");
    }
    v->tiledR = R;
    MB_TRY(tiledConfigure());
    const int64_t n = a->numLocal;
    if (v->hStats == nullptr) MB_CUDA(cudaMallocHost(&v->hStats, 16));
    if (v->hTstats == nullptr) MB_CUDA(cudaMallocHost(&v->hTstats, 16));
    MB_TRY(v->stats.reserve(16));
    MB_TRY(v->tstats.reserve(16));
    v->numParticles = n;
    v->begin = 0;
    v->end = n;
    v->pitch = n;
    MB_TRY(v->counts.reserve(size_t(std::max<int64_t>(n, 1)) * 4));
    v->buildCount += 1;
    v->tiled = true;
    v->tiledSub = *s;
    v->tiledEpoch = a->lcEpoch;
    {
        const int64_t perX = int64_t(g.n[1]) * g.n[2];
        const int64_t extCells = a->lcNumCells + (haloX ? 2 * perX : 0);
        MB_TRY(v->cellLoHi.reserve(size_t(extCells) * 8));
        extCellRangesKernel<<<gridFor(extCells, 256), 256, 0, st>>>(a->lcCellStart.as<int32_t>(), a->lcNumCells,
                                                                   haloX ? haloLeft : nullptr, haloX ? haloRight : nullptr,
                                                                   perX, v->cellLoHi.as<int32_t>(),
                                                                   v->cellLoHi.as<int32_t>() + extCells);
        MB_LAUNCHED();
    }
    const int32_t* cellLo = v->cellLoHi.as<int32_t>();
    const int32_t* cellHi = cellLo + (a->lcNumCells + (haloX ? 2 * int64_t(g.n[1]) * g.n[2] : 0));
    // Cabana's grid for the stencil-pruning check (grid_min/max = ghost corners, delta = radius * ratio)
    const double gs = radius * cellRatio;
    const double delta[3] = {gs, gs, gs};
    const GridDev cabanaGrid = makeGrid(s->minGhostCorner, s->maxGhostCorner, delta);
    const double rsqr = radius * radius;
    int64_t width = (std::max<int64_t>(maxNeigh, 1) + 63) & ~int64_t(63);  // tiledRowIndex: eight lanes x 16-byte words
    if (v->width > width && v->enc.bytes >= size_t(v->width) * std::max<int64_t>(n, 1) * 2) width = v->width;
    MB_REQUIRE(width <= 1024, "verlet_build_periodic: more than 1024 neighbours per atom");
    // shared memory of the builder: positions + one row buffer per lane (width x 2 bytes + one padding word each)
    auto buildSmem = [&](int slots, int64_t w)
    { return size_t(slots + TL_BUILD_BATCH) * TL_SMEM_PER_SLOT_BUILD + 16 + size_t(TL_THREADS_BUILD) * size_t(w / 2 + 1) * 4; };

    // Tile geometry (CH cells per tile, staged-slot capacity).  The first build measures the tiles (one host round
    // trip); later builds on the same grid re-use CH with the largest tile of the previous build plus head room as the
    // capacity and verify after the fact (tstats[2], read back together with the list statistics): no extra
    // synchronisation per rebuild.
    const bool sameGrid = v->tiledCH > 0 && v->tiledHaloX == haloX && v->tiledGridN[0] == g.n[0] &&
                          v->tiledGridN[1] == g.n[1] && v->tiledGridN[2] == g.n[2] && v->hTstats[0] > 0;
    for (int attempt = 0; attempt < 4; ++attempt)
    {
        TileParams tp;
        int tiles = 0, CH = 0;
        if (sameGrid && attempt == 0)
        {
            CH = v->tiledCH;
            MB_TRY(makeTileParams(a, s, CH, 0, haloX, R, tp));
            tiles = g.n[0] * g.n[1] * tp.numChunks;
            // 3 % head room over the last build's largest tile (atoms move by less than the skin between rebuilds)
            v->tiledSlots = (std::min(v->hTstats[0] + v->hTstats[0] / 32 + 8, 65534) + 1) & ~1;
            MB_TRY(v->tileDesc.reserve(size_t(tiles) * TL_DESC_INTS * 4));
            tileDescKernel<<<tiles, 32, 0, st>>>(tp, cellLo, cellHi, v->tileDesc.as<int>(), nullptr);
            MB_LAUNCHED();
        }
        else
        {
            // choose CH: ~110 home atoms per tile, at least ~2 tiles per SM, staged slots within the smem budget
            const double perCell = double(n) / double(std::max<int64_t>(a->lcNumCells, 1));
            CH = std::max(1, std::min({TL_MAX_CH, g.n[2],
                                       static_cast<int>(std::ceil(double(v->tiledTargetHomes) / std::max(perCell, 0.1)))}));
            while (CH > 1 && int64_t(g.n[0]) * g.n[1] * ((g.n[2] + CH - 1) / CH) < 2 * 148) CH = (CH + 1) / 2;
            for (;;)
            {
                MB_TRY(makeTileParams(a, s, CH, 0, haloX, R, tp));
                tiles = g.n[0] * g.n[1] * tp.numChunks;
                MB_TRY(v->tileDesc.reserve(size_t(tiles) * TL_DESC_INTS * 4));
                MB_CUDA(cudaMemsetAsync(v->tstats.p, 0, 16, st));
                tileDescKernel<<<tiles, 32, 0, st>>>(tp, cellLo, cellHi, v->tileDesc.as<int>(), v->tstats.as<int>());
                MB_LAUNCHED();
                MB_CUDA(cudaMemcpyAsync(v->hTstats, v->tstats.p, 16, cudaMemcpyDeviceToHost, st));
                MB_CUDA(cudaStreamSynchronize(st));
                const int slots = v->hTstats[0];
                if (slots * std::max(TL_SMEM_PER_SLOT_BUILD, v->tiledSlotBytes) <= v->tiledSmemBudget || CH == 1)
                {
                    MB_REQUIRE(buildSmem(slots, width) + 64 <= size_t(TL_SMEM_MAX) && slots < 65535,
                               "verlet_build_periodic: a tile exceeds shared memory");
                    v->tiledSlots = (std::max(slots, 1) + 1) & ~1;  // even: keeps the arrays behind the positions 16-byte aligned
                    break;
                }
                CH = (CH + 1) / 2;
            }
        }
        v->tiledCH = CH;
        v->tiledHaloX = haloX;
        for (int d = 0; d < 3; ++d) v->tiledGridN[d] = g.n[d];
        tp.cap = v->tiledSlots;
        {
            // rows are loaded in whole 16-byte words, also behind the stored entries (never used, but they should not be
            // uninitialised memory either): a fresh allocation is cleared once
            const void* before = v->enc.p;
            MB_TRY(v->enc.reserve(size_t(width) * std::max<int64_t>(n, 1) * 2));
            if (v->enc.p != before) MB_CUDA(cudaMemsetAsync(v->enc.p, 0, v->enc.bytes, st));
        }
        MB_TRY(v->tileActive.reserve(size_t(tiles)));
        MB_TRY(v->activeTiles.reserve(size_t(tiles) * 4));
        v->width = width;
        MB_CUDA(cudaMemsetAsync(v->stats.p, 0, 16, st));
        MB_CUDA(cudaMemsetAsync(v->tstats.p, 0, 16, st));
        const size_t smem = buildSmem(v->tiledSlots, width);
        MB_REQUIRE(smem <= size_t(TL_SMEM_MAX), "verlet_build_periodic: a tile exceeds shared memory");
#define VBT_LAUNCH(H)                                                                                                  \
    verletBuildTiledKernel<H><<<tiles, TL_THREADS_BUILD, smem, st>>>(                                                  \
        tp, cabanaGrid, a->v.pos, cellLo, v->tileDesc.as<int>(), rsqr, static_cast<int>(width), v->counts.as<int32_t>(), \
        v->enc.as<uint16_t>(), v->stats.as<int32_t>(), v->tstats.as<int32_t>(), v->tileActive.as<unsigned char>(),     \
        v->tiledCgSkip ? 1 : 0, v->tiledCgWeight)
        if (v->half) VBT_LAUNCH(true);
        else VBT_LAUNCH(false);
#undef VBT_LAUNCH
        MB_LAUNCHED();
        if (v->tiledCgSkip)  // the list of tiles with work is only used by the AdResS step loops
        {
            compactActiveTilesKernel<<<1, 1024, 0, st>>>(v->tileActive.as<unsigned char>(), tiles, v->activeTiles.as<int32_t>(),
                                                         v->tstats.as<int32_t>() + 3);
            MB_LAUNCHED();
        }
        MB_CUDA(cudaMemcpyAsync(v->hStats, v->stats.p, 16, cudaMemcpyDeviceToHost, st));
        MB_CUDA(cudaMemcpyAsync(v->hTstats, v->tstats.p, 16, cudaMemcpyDeviceToHost, st));
        MB_CUDA(cudaStreamSynchronize(st));
        if (v->hTstats[2] != 0)
        {
            v->tiledCH = 0;  // a tile outgrew the re-used geometry: measure again
            v->hTstats[0] = 0;
            continue;
        }
        v->numActiveTiles = v->hTstats[3];
        if (v->hStats[0] <= width) return 0;
        width = (int64_t(v->hStats[0]) + 63) & ~int64_t(63);
        MB_REQUIRE(width <= 1024, "verlet_build_periodic: more than 1024 neighbours per atom");
    }
    setLastError("verlet_build_periodic: neighbour table overflow after refill");
    return MRMD_B200_ECAPACITY;
}
}  // namespace mrmd_b200

namespace mrmd_b200
{
int verletBuildTiledMolecules(mrmd_b200_verlet* v, mrmd_b200_molecules* m, const mrmd_b200_subdomain* s, double radius,
                              double cellRatio, int64_t maxNeigh, int atomsPerMolecule, const int32_t* haloLeft,
                              const int32_t* haloRight, cudaStream_t st)
{
    MB_REQUIRE(v != nullptr && m != nullptr && m->lcView != nullptr,
               "verlet_build_periodic_molecules: sort the molecules first (mrmd_b200_molecules_cell_sort_with_atoms)");
    MB_REQUIRE(atomsPerMolecule == 4, "verlet_build_periodic_molecules: four atoms per molecule");
    m->lcView->v.pos = m->v.pos;
    // a staged slot of the force kernel holds all atoms of a molecule: fewer homes per tile than for atoms
    v->tiledTargetHomes = 56;  // measured 16 / 28 / 56: 0.44 / 0.35 / 0.32 ms per 4M atoms (TL_MAX_CH caps the tile height)
    v->tiledSlotBytes = molSlotBytes<4>();
    v->tiledSmemBudget = TL_SMEM_BUDGET;
    // measurement knobs: homes per tile aimed at, shared-memory budget of a tile in KB
    if (const char* e = std::getenv("MRMD_B200_MOL_HOMES")) v->tiledTargetHomes = std::max(4, std::atoi(e));
    if (const char* e = std::getenv("MRMD_B200_MOL_SMEM_KB")) v->tiledSmemBudget = std::min(TL_SMEM_MAX, std::max(16, std::atoi(e)) * 1024);
    return verletBuildTiled(v, m->lcView, s, radius, cellRatio, maxNeigh, haloLeft, haloRight, st);
}
}  // namespace mrmd_b200

using namespace mrmd_b200;

extern "C" {

int mrmd_b200_verlet_build_periodic_molecules(mrmd_b200_verlet* v, mrmd_b200_molecules* m, const mrmd_b200_subdomain* s,
                                              double radius, double cellRatio, int64_t maxNeigh, int atomsPerMolecule,
                                              void* stream)
{
    MB_TRY(checkDevice());
    return verletBuildTiledMolecules(v, m, s, radius, cellRatio, maxNeigh, atomsPerMolecule, nullptr, nullptr, S(stream));
}

int mrmd_b200_verlet_read_periodic_molecules(const mrmd_b200_verlet* v, const mrmd_b200_molecules* m, int32_t* countsHost,
                                             int32_t* partnerHost, int32_t* shiftCodeHost, void* stream)
{
    MB_REQUIRE(m != nullptr && m->lcView != nullptr, "verlet_read_periodic_molecules: not a molecule list");
    return mrmd_b200_verlet_read_periodic(v, m->lcView, countsHost, partnerHost, shiftCodeHost, stream);
}

int mrmd_b200_verlet_build_periodic(mrmd_b200_verlet* v, const mrmd_b200_atoms* a, const mrmd_b200_subdomain* s,
                                    double radius, double cellRatio, int64_t maxNeigh, void* stream)
{
    MB_TRY(checkDevice());
    return verletBuildTiled(v, a, s, radius, cellRatio, maxNeigh, nullptr, nullptr, S(stream));
}

int mrmd_b200_verlet_read_periodic(const mrmd_b200_verlet* v, const mrmd_b200_atoms* a, int32_t* countsHost,
                                   int32_t* partnerHost, int32_t* shiftCodeHost, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(v != nullptr && a != nullptr && v->tiled, "verlet_read_periodic: not a tiled list");
    MB_REQUIRE(a->lcValid && a->lcEpoch == v->tiledEpoch, "verlet_read_periodic: atoms were re-sorted since the build");
    cudaStream_t st = S(stream);
    const int64_t n = v->numParticles;
    if (n == 0) return 0;
    TileParams tp;
    MB_TRY(makeTileParams(a, &v->tiledSub, v->tiledCH, v->tiledSlots, v->tiledHaloX, v->tiledR, tp));
    const int tiles = tp.g.n[0] * tp.g.n[1] * tp.numChunks;
    int32_t* d = nullptr;
    MB_CUDA(cudaMalloc(&d, size_t(n) * v->width * 8));
    decodeTiledKernel<<<tiles, TL_THREADS_BUILD, 0, st>>>(tp, v->tileDesc.as<int>(), v->counts.as<int32_t>(),
                                                    v->enc.as<uint16_t>(), static_cast<int>(v->width), d,
                                                    d + n * v->width);
    g_launchCount.fetch_add(1);
    if (countsHost) cudaMemcpyAsync(countsHost, v->counts.p, size_t(n) * 4, cudaMemcpyDeviceToHost, st);
    if (partnerHost) cudaMemcpyAsync(partnerHost, d, size_t(n) * v->width * 4, cudaMemcpyDeviceToHost, st);
    if (shiftCodeHost) cudaMemcpyAsync(shiftCodeHost, d + n * v->width, size_t(n) * v->width * 4, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(d);
    MB_CUDA(e);
    return 0;
}

}  // extern "C"
