// neighbor.cu -- cell binning, atomic-free radix sort, spatial permutation and Verlet-list build
// (SURVEY.md K10, K11).  Replaces Cabana::LinkedCellList + Cabana::permute and Cabana::VerletList::build
// (Cabana 0.7, 7914d28: not in the reference tree; call sites tests/NVT/NVT.cpp:136-154,
// examples/02_LennardJones_NVE.cpp:156-163, tests/LennardJones/LennardJones.cpp:112-118).
//
// Semantics kept bit for bit with oracle/mrmd_oracle.cpp (Grid, or_verlet_build):
//   grid     n_d = floor((max_d-min_d) * (1/delta)), dx_d = (max_d-min_d)/n_d, rdx_d = 1/dx_d
//   cell     i_d = floor((x_d-min_d)*rdx_d), i_d == n_d -> n_d-1; cardinal = (i*ny + j)*nz + k
//   pair     p in [begin,end), n any particle, valid(p,n), dx*dx+dy*dy+dz*dz <= r*r (no FMA), and the
//            stencil cell of n passes minDistanceToPoint(x_p, cell) <= r*r
//   valid    full: p != n;  half: p != n and x_n lexicographically greater than x_p
// The neighbour table is stored slot-major (neigh[slot * pitch + particle]) so that a thread-per-particle
// force kernel reads it coalesced; mrmd_b200_verlet_read exports Cabana's row-major VerletLayout2D.
#include <algorithm>
#include <cmath>

#include "common.cuh"


namespace mrmd_b200
{
GridDev makeGrid(const double* gmin, const double* gmax, const double* delta)
{
    GridDev g;
    for (int d = 0; d < 3; ++d)
    {
        g.min[d] = gmin[d];
        int n = static_cast<int>(std::floor((gmax[d] - gmin[d]) * (1.0 / delta[d])));
        if (n < 1) n = 1;
        g.n[d] = n;
        g.dx[d] = (gmax[d] - gmin[d]) / n;
        g.rdx[d] = 1.0 / g.dx[d];
    }
    return g;
}

// ---------------------------------------------------------------------------------------------
// stable LSD radix sort on (uint32 key, uint32 value) pairs, 8-bit digits, no atomics: ranks come
// from warp match/ballot, block and grid offsets from prefix sums.
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_BINS = 256;

__device__ __forceinline__ uint32_t digitOf(uint32_t key, int shift) { return (key >> shift) & 0xFFu; }

// per-warp digit counts of this block's tile -> sWarp[warp][digit]; returns each item's rank among
// equal digits seen so far by its warp (stable: warps own contiguous sub-tiles, rounds go upward)
__device__ __forceinline__ void rankTile(const uint32_t* keys, int64_t n, int shift, int64_t tileBase,
                                         uint32_t (*sWarp)[RS_BINS], uint32_t* digits, uint32_t* ranks)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b = threadIdx.x; b < RS_WARPS * RS_BINS; b += RS_THREADS) (&sWarp[0][0])[b] = 0;
    __syncthreads();
    const int64_t warpBase = tileBase + int64_t(warp) * 32 * RS_ITEMS;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r)
    {
        const int64_t i = warpBase + r * 32 + lane;
        const bool valid = i < n;
        const uint32_t d = valid ? digitOf(keys[i], shift) : 0xFFFFFFFFu;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t before = __popc(peers & ((1u << lane) - 1u));
        uint32_t base = 0;
        if (valid) base = sWarp[warp][d];
        __syncwarp();
        if (valid && before == 0) sWarp[warp][d] = base + __popc(peers);  // one writer per digit
        __syncwarp();
        digits[r] = d;
        ranks[r] = base + before;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(RS_THREADS) rsHistogramKernel(const uint32_t* keys, int64_t n, int shift,
                                                                uint32_t* blockHist, int numBlocks)
{
    __shared__ uint32_t sWarp[RS_WARPS][RS_BINS];
    uint32_t digits[RS_ITEMS], ranks[RS_ITEMS];
    rankTile(keys, n, shift, int64_t(blockIdx.x) * RS_TILE, sWarp, digits, ranks);
    uint32_t sum = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) sum += sWarp[w][threadIdx.x];
    blockHist[size_t(threadIdx.x) * numBlocks + blockIdx.x] = sum;  // digit-major
}

// exclusive scan of every digit row of blockHist (digit-major: row d holds the numBlocks per-block counts
// of digit d); block d scans row d and leaves the row total in digitTotals[d]
__global__ void __launch_bounds__(RS_THREADS) rsScanRowsKernel(uint32_t* blockHist, int numBlocks, uint32_t* digitTotals)
{
    __shared__ uint32_t sWarpSum[RS_WARPS];
    __shared__ uint32_t sCarry;
    uint32_t* row = blockHist + size_t(blockIdx.x) * numBlocks;
    if (threadIdx.x == 0) sCarry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < numBlocks; base += RS_THREADS)
    {
        const int i = base + threadIdx.x;
        const uint32_t v = (i < numBlocks) ? row[i] : 0;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) sWarpSum[warp] = x;
        __syncthreads();
        uint32_t warpOffset = 0, total = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w)
        {
            const uint32_t c = sWarpSum[w];
            if (w < warp) warpOffset += c;
            total += c;
        }
        const uint32_t carry = sCarry;
        if (i < numBlocks) row[i] = carry + warpOffset + x - v;
        __syncthreads();
        if (threadIdx.x == 0) sCarry = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) digitTotals[blockIdx.x] = sCarry;
}

__global__ void __launch_bounds__(RS_THREADS)
    rsScatterKernel(const uint32_t* keysIn, const uint32_t* valsIn, uint32_t* keysOut, uint32_t* valsOut, int64_t n,
                    int shift, const uint32_t* scannedHist, const uint32_t* digitTotals, int numBlocks)
{
    __shared__ uint32_t sWarp[RS_WARPS][RS_BINS];
    __shared__ uint32_t sDigitWarp[RS_WARPS];
    uint32_t digits[RS_ITEMS], ranks[RS_ITEMS];
    const int64_t tileBase = int64_t(blockIdx.x) * RS_TILE;
    // exclusive scan of the 256 digit totals (thread t <-> digit t)
    uint32_t digitBase;
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const uint32_t v = digitTotals[threadIdx.x];
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) sDigitWarp[warp] = x;
        __syncthreads();
        uint32_t off = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w)
            if (w < warp) off += sDigitWarp[w];
        digitBase = off + x - v;
    }
    rankTile(keysIn, n, shift, tileBase, sWarp, digits, ranks);
    {
        // per digit: exclusive prefix over warps plus the global offset of (digit, block)
        uint32_t run = digitBase + scannedHist[size_t(threadIdx.x) * numBlocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w)
        {
            const uint32_t c = sWarp[w][threadIdx.x];
            sWarp[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warpBase = tileBase + int64_t(warp) * 32 * RS_ITEMS;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r)
    {
        const int64_t i = warpBase + r * 32 + lane;
        if (i < n)
        {
            const uint32_t dst = sWarp[warp][digits[r]] + ranks[r];
            keysOut[dst] = keysIn[i];
            valsOut[dst] = valsIn[i];
        }
    }
}

size_t radixSortScratchBytes(int64_t n)
{
    const int64_t numBlocks = (n + RS_TILE - 1) / RS_TILE;
    return size_t(RS_BINS) * size_t(std::max<int64_t>(numBlocks, 1) + 1) * 4;
}

int radixSortPairs(uint32_t* keysIn, uint32_t* valsIn, uint32_t* keysTmp, uint32_t* valsTmp, uint32_t* scratch,
                   int64_t n, int keyBits, uint32_t** keysOut, uint32_t** valsOut, cudaStream_t st)
{
    uint32_t *kIn = keysIn, *vIn = valsIn, *kOut = keysTmp, *vOut = valsTmp;
    if (n > 0)
    {
        const int numBlocks = static_cast<int>((n + RS_TILE - 1) / RS_TILE);
        const int passes = std::max(1, (keyBits + 7) / 8);
        for (int p = 0; p < passes; ++p)
        {
            rsHistogramKernel<<<numBlocks, RS_THREADS, 0, st>>>(kIn, n, 8 * p, scratch, numBlocks);
            MB_LAUNCHED();
            uint32_t* digitTotals = scratch + size_t(RS_BINS) * numBlocks;
            rsScanRowsKernel<<<RS_BINS, RS_THREADS, 0, st>>>(scratch, numBlocks, digitTotals);
            MB_LAUNCHED();
            rsScatterKernel<<<numBlocks, RS_THREADS, 0, st>>>(kIn, vIn, kOut, vOut, n, 8 * p, scratch, digitTotals,
                                                              numBlocks);
            MB_LAUNCHED();
            std::swap(kIn, kOut);
            std::swap(vIn, vOut);
        }
    }
    *keysOut = kIn;
    *valsOut = vIn;
    return 0;
}

static int bitsFor(int64_t numCells)
{
    int bits = 1;
    while ((int64_t(1) << bits) < numCells) ++bits;
    return bits;
}

// ---------------------------------------------------------------------------------------------
// dropFlags (optional): atoms flagged non-zero get the key numCells and sort behind the last cell (atoms that
// migrated to another rank, slab.cu)
__global__ void cellKeyKernel(const double4* pos, int64_t first, int64_t count, GridDev g, uint32_t* keys,
                              uint32_t* vals, int32_t* cellIdOut, const signed char* dropFlags = nullptr,
                              uint32_t dropKey = 0)
{
    const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (j >= count) return;
    const double4 p = ld4nc(pos + first + j);
    int c = cardinal(g, locate1(g, p.x, 0), locate1(g, p.y, 1), locate1(g, p.z, 2));
    if (dropFlags != nullptr && dropFlags[first + j] != 0) c = static_cast<int>(dropKey);
    keys[j] = static_cast<uint32_t>(c);
    vals[j] = static_cast<uint32_t>(first + j);
    if (cellIdOut != nullptr) cellIdOut[first + j] = c;
}

// x-slab decomposition (slab.cu): the key kernel of the rebuild also applies the y / z periodic wrap
// (PeriodicMapping.cpp:36-51 arithmetic, written back) and gives the atoms that left the slab in x the key numCells, so
// that they sort behind the last cell and are dropped -- one pass over the positions instead of three
__global__ void cellKeySlabKernel(double4* pos, int64_t count, GridDev g, SubdomainDev s, uint32_t* keys, uint32_t* vals,
                                  uint32_t dropKey)
{
    const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (j >= count) return;
    double4 p = ld4(pos + j);
    double* x = &p.x;
    bool moved = false;
#pragma unroll
    for (int dim = 1; dim < 3; ++dim)
    {
        if (s.maxCorner[dim] <= x[dim])
        {
            x[dim] -= s.diameter[dim];
            x[dim] = fmax(x[dim], s.minCorner[dim]);
            moved = true;
        }
        if (x[dim] < s.minCorner[dim])
        {
            x[dim] += s.diameter[dim];
            if (s.maxCorner[dim] <= x[dim]) x[dim] = s.minCorner[dim];
            moved = true;
        }
    }
    if (moved) st4(pos + j, p);
    const bool gone = (p.x < s.minCorner[0]) || (p.x >= s.maxCorner[0]);
    keys[j] = gone ? dropKey : static_cast<uint32_t>(cardinal(g, locate1(g, p.x, 0), locate1(g, p.y, 1), locate1(g, p.z, 2)));
    vals[j] = static_cast<uint32_t>(j);
}

// slot i of the new arrays receives record src(i): perm inside [begin,end), identity outside
__global__ void permuteAtomsKernel(AtomsView dst, AtomsView src, const uint32_t* perm, int64_t begin, int64_t end,
                                   int64_t size)
{
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= size) return;
    const int64_t s = (i >= begin && i < end) ? int64_t(perm[i - begin]) : i;
    st4(dst.pos + i, ld4nc(src.pos + s));
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
        dst.vel[d][i] = src.vel[d][s];
        dst.force[d][i] = src.force[d][s];
    }
    dst.mass[i] = src.mass[s];
    dst.charge[i] = src.charge[s];
    dst.relMass[i] = src.relMass[s];
    dst.gid[i] = src.gid[s];
}

// molecules of apm consecutive atoms: atom block k of the new arrays receives block perm[k] (all members)
__global__ void permuteAtomBlocksKernel(AtomsView dst, AtomsView src, const uint32_t* perm, int64_t numBlocks, int apm)
{
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= numBlocks * apm) return;
    const int64_t s = int64_t(perm[i / apm]) * apm + i % apm;
    st4(dst.pos + i, ld4nc(src.pos + s));
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
        dst.vel[d][i] = src.vel[d][s];
        dst.force[d][i] = src.force[d][s];
    }
    dst.mass[i] = src.mass[s];
    dst.charge[i] = src.charge[s];
    dst.relMass[i] = src.relMass[s];
    dst.gid[i] = src.gid[s];
}
__global__ void permutePos4Kernel(double4* dst, const double4* src, const uint32_t* perm, int64_t n)
{
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i < n) st4(dst + i, ld4nc(src + perm[i]));
}

__global__ void permuteMolsKernel(MolsView dst, MolsView src, const uint32_t* perm, int64_t begin, int64_t end,
                                  int64_t size)
{
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= size) return;
    const int64_t s = (i >= begin && i < end) ? int64_t(perm[i - begin]) : i;
    st4(dst.pos + i, ld4nc(src.pos + s));
    st4(dst.w + i, ld4nc(src.w + s));
    dst.oc[i] = src.oc[s];
    dst.lambda[i] = src.lambda[s];
#pragma unroll
    for (int d = 0; d < 3; ++d) dst.force[d][i] = src.force[d][s];
}

// cellStart[c] = offset + first sorted slot whose key is >= c, c in [0, numCells].  One thread per sorted slot k (and one
// behind the last): it owns the cells c with key[k-1] < c <= key[k] -- every cell is written exactly once, a run of empty
// cells by the slot behind it; two coalesced loads per thread instead of a binary search per cell (17 -> 4 us per 1M atoms
// and 442k cells).
__global__ void cellStartKernel(const uint32_t* __restrict__ sortedKeys, int64_t n, int64_t numCells, int32_t* cellStart,
                                int32_t offset)
{
    const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (n == 0)  // no slots: the one block fills the table
    {
        for (int64_t c = threadIdx.x; c <= numCells; c += blockDim.x) cellStart[c] = offset;
        return;
    }
    if (k > n) return;
    const int64_t prev = (k == 0) ? -1 : int64_t(sortedKeys[k - 1]);
    const int64_t cur = (k == n) ? numCells : min(int64_t(sortedKeys[k]), numCells);
    for (int64_t c = prev + 1; c <= cur; ++c) cellStart[c] = static_cast<int32_t>(k) + offset;
}

__global__ void gatherSortedPosKernel(const double4* pos, const uint32_t* sortedIdx, int64_t n, double4* sortedPos)
{
    const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (k >= n) return;
    const uint32_t idx = sortedIdx[k];
    double4 p = ld4nc(pos + idx);
    p.w = __longlong_as_double(static_cast<long long>(idx));
    st4(sortedPos + k, p);
}

// Cabana CartesianGrid::minDistanceToPoint, no FMA (oracle Grid::minDistanceToPoint)
__device__ __forceinline__ double minDist1(const GridDev& g, double x, int c, int d)
{
    const double xc = __dadd_rn(g.min[d], __dmul_rn(double(c) + 0.5, g.dx[d]));
    const double rr = __dsub_rn(fabs(__dsub_rn(x, xc)), __dmul_rn(0.5, g.dx[d]));
    return (rr > 0.0) ? rr : 0.0;
}

// One thread per particle in cell order.  Threads of a warp sit in the same or adjacent cells, so the
// candidate loops are nearly uniform and the 256-bit candidate loads are L1 broadcast hits.  Rows are
// written slot-major; with cell-sorted particles the writes of a warp are contiguous.
template <bool HALF>
__global__ void __launch_bounds__(128)
    verletBuildKernel(const double4* __restrict__ sortedPos, const int32_t* __restrict__ cellStart, int64_t n,
                      GridDev g, int cellRange, double rsqr, int64_t begin, int64_t end, int64_t width, int64_t pitch,
                      int32_t* __restrict__ counts, int32_t* __restrict__ neigh, int32_t* stats)
{
    const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    int count = 0;
    if (k < n)
    {
        const double4 p = ld4nc(sortedPos + k);
        const int64_t pid = __double_as_longlong(p.w);
        if (pid >= begin && pid < end)
        {
            const int ci = locate1(g, p.x, 0), cj = locate1(g, p.y, 1), ck = locate1(g, p.z, 2);
            // a cell left of the home cell only holds x_n < x_p: never a valid half-list partner
            const int imin = HALF ? ci : max(0, ci - cellRange);
            const int imax = min(g.n[0], ci + cellRange + 1);
            const int jmin = max(0, cj - cellRange), jmax = min(g.n[1], cj + cellRange + 1);
            const int kmin = max(0, ck - cellRange), kmax = min(g.n[2], ck + cellRange + 1);
            for (int ii = imin; ii < imax; ++ii)
            {
                const double rx = minDist1(g, p.x, ii, 0);
                const double rx2 = __dmul_rn(rx, rx);
                for (int jj = jmin; jj < jmax; ++jj)
                {
                    const double ry = minDist1(g, p.y, jj, 1);
                    const double rxy2 = __dadd_rn(rx2, __dmul_rn(ry, ry));
                    for (int kk = kmin; kk < kmax; ++kk)
                    {
                        const double rz = minDist1(g, p.z, kk, 2);
                        if (!(__dadd_rn(rxy2, __dmul_rn(rz, rz)) <= rsqr)) continue;
                        const int c = cardinal(g, ii, jj, kk);
                        const int s0 = cellStart[c], s1 = cellStart[c + 1];
                        for (int s = s0; s < s1; ++s)
                        {
                            const double4 q = ld4nc(sortedPos + s);
                            bool valid = (s != k);
                            if (HALF)
                                valid = valid && ((q.x > p.x) || ((q.x == p.x) && ((q.y > p.y) || ((q.y == p.y) && (q.z > p.z)))));
                            if (!valid) continue;
                            const double d2 = distSqrExact(p.x - q.x, p.y - q.y, p.z - q.z);
                            if (d2 <= rsqr)
                            {
                                if (count < width) neigh[int64_t(count) * pitch + pid] = static_cast<int32_t>(__double_as_longlong(q.w));
                                ++count;
                            }
                        }
                    }
                }
            }
            counts[pid] = count;
        }
    }
    // max row length, sum of row lengths
    int mx = count;
    long long total = count;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        total += __shfl_xor_sync(0xffffffffu, total, o);
    }
    if ((threadIdx.x & 31) == 0 && total > 0)
    {
        atomicMax(stats, mx);
        atomicAdd(reinterpret_cast<unsigned long long*>(stats + 2), static_cast<unsigned long long>(total));
    }
}

__global__ void transposeListKernel(const int32_t* counts, const int32_t* neigh, int64_t numParticles, int64_t width,
                                    int64_t pitch, int32_t* rowMajor)
{
    const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (t >= numParticles * width) return;
    const int64_t i = t / width, s = t % width;
    rowMajor[t] = (s < counts[i]) ? neigh[s * pitch + i] : -1;
}

static int cellSortPrepare(DevBuf& scratch, int64_t count, uint32_t** keys0, uint32_t** vals0, uint32_t** keys1,
                           uint32_t** vals1, uint32_t** hist)
{
    const size_t arr = (size_t(count) * 4 + 255) & ~size_t(255);
    MB_TRY(scratch.reserve(4 * arr + radixSortScratchBytes(count)));
    char* p = scratch.as<char>();
    *keys0 = reinterpret_cast<uint32_t*>(p);
    *vals0 = reinterpret_cast<uint32_t*>(p + arr);
    *keys1 = reinterpret_cast<uint32_t*>(p + 2 * arr);
    *vals1 = reinterpret_cast<uint32_t*>(p + 3 * arr);
    *hist = reinterpret_cast<uint32_t*>(p + 4 * arr);
    return 0;
}

static int verletBuild(mrmd_b200_verlet* v, const double4* pos, int64_t nAll, int64_t begin, int64_t end, double radius,
                       double cellRatio, const double* gridMin, const double* gridMax, int64_t maxNeigh, cudaStream_t st)
{
    MB_REQUIRE(v != nullptr && gridMin != nullptr && gridMax != nullptr, "verlet_build");
    MB_REQUIRE(begin >= 0 && begin <= end && end <= nAll, "verlet_build: [begin,end) outside the particle range");
    MB_REQUIRE(radius > 0.0 && cellRatio > 0.0, "verlet_build: radius and cell ratio must be positive");
    MB_REQUIRE(nAll < (int64_t(1) << 31), "verlet_build: more than 2^31 particles");
    const double gridSize = cellRatio * radius;
    const double delta[3] = {gridSize, gridSize, gridSize};
    const GridDev g = makeGrid(gridMin, gridMax, delta);
    const int64_t numCells = int64_t(g.n[0]) * g.n[1] * g.n[2];
    MB_REQUIRE(numCells < (int64_t(1) << 31), "verlet_build: too many cells");
    const int cellRange = static_cast<int>(std::ceil(1.0 / cellRatio));
    const double rsqr = radius * radius;

    v->tiled = false;
    v->numParticles = nAll;
    v->begin = begin;
    v->end = end;
    v->pitch = (nAll + 31) & ~int64_t(31);
    if (v->pitch == 0) v->pitch = 32;
    if (v->hStats == nullptr) MB_CUDA(cudaMallocHost(&v->hStats, 16));
    MB_TRY(v->stats.reserve(16));
    MB_TRY(v->counts.reserve(size_t(v->pitch) * 4));
    MB_CUDA(cudaMemsetAsync(v->counts.p, 0, size_t(v->pitch) * 4, st));
    v->buildCount += 1;
    if (nAll == 0)
    {
        v->width = std::max<int64_t>(maxNeigh, 1);
        return 0;
    }
    // bin all particles: keys = cell index, values = particle index
    for (int b = 0; b < 2; ++b)
    {
        MB_TRY(v->keys[b].reserve(size_t(nAll) * 4));
        MB_TRY(v->vals[b].reserve(size_t(nAll) * 4));
    }
    MB_TRY(v->scratch.reserve(radixSortScratchBytes(nAll)));
    MB_TRY(v->cellStart.reserve(size_t(numCells + 1) * 4));
    MB_TRY(v->sortedPos.reserve(size_t(nAll) * 32));
    cellKeyKernel<<<gridFor(nAll, 256), 256, 0, st>>>(pos, 0, nAll, g, v->keys[0].as<uint32_t>(), v->vals[0].as<uint32_t>(),
                                                      nullptr);
    MB_LAUNCHED();
    uint32_t *sortedKeys, *sortedIdx;
    MB_TRY(radixSortPairs(v->keys[0].as<uint32_t>(), v->vals[0].as<uint32_t>(), v->keys[1].as<uint32_t>(),
                          v->vals[1].as<uint32_t>(), v->scratch.as<uint32_t>(), nAll, bitsFor(numCells), &sortedKeys,
                          &sortedIdx, st));
    cellStartKernel<<<gridFor(nAll + 1, 256), 256, 0, st>>>(sortedKeys, nAll, numCells, v->cellStart.as<int32_t>(), 0);
    MB_LAUNCHED();
    gatherSortedPosKernel<<<gridFor(nAll, 256), 256, 0, st>>>(pos, sortedIdx, nAll, v->sortedPos.as<double4>());
    MB_LAUNCHED();

    int64_t width = std::max<int64_t>(maxNeigh, 1);
    if (v->width > width && v->neigh.bytes >= size_t(v->width) * v->pitch * 4) width = v->width;  // keep a widened table
    for (int attempt = 0; attempt < 3; ++attempt)
    {
        MB_TRY(v->neigh.reserve(size_t(width) * v->pitch * 4));
        v->width = width;
        MB_CUDA(cudaMemsetAsync(v->stats.p, 0, 16, st));
        if (v->half)
            verletBuildKernel<true><<<gridFor(nAll, 128), 128, 0, st>>>(
                v->sortedPos.as<double4>(), v->cellStart.as<int32_t>(), nAll, g, cellRange, rsqr, begin, end, width,
                v->pitch, v->counts.as<int32_t>(), v->neigh.as<int32_t>(), v->stats.as<int32_t>());
        else
            verletBuildKernel<false><<<gridFor(nAll, 128), 128, 0, st>>>(
                v->sortedPos.as<double4>(), v->cellStart.as<int32_t>(), nAll, g, cellRange, rsqr, begin, end, width,
                v->pitch, v->counts.as<int32_t>(), v->neigh.as<int32_t>(), v->stats.as<int32_t>());
        MB_LAUNCHED();
        MB_CUDA(cudaMemcpyAsync(v->hStats, v->stats.p, 16, cudaMemcpyDeviceToHost, st));
        MB_CUDA(cudaStreamSynchronize(st));
        if (v->hStats[0] <= width) return 0;
        width = v->hStats[0];  // Cabana: reallocate to the maximum count and refill
    }
    setLastError("verlet_build: neighbour table overflow after refill");
    return MRMD_B200_ECAPACITY;
}
}  // namespace mrmd_b200

using namespace mrmd_b200;

extern "C" {

}  // extern "C"

namespace mrmd_b200
{
static int atomsCellSortImpl(mrmd_b200_atoms* a, int64_t begin, int64_t end, const double* delta, const double* gridMin,
                             const double* gridMax, int32_t* cellIdOut, const signed char* dropFlags, cudaStream_t st,
                             const SubdomainDev* slabWrap = nullptr);

// the slab rebuild's sort over [0, end): y / z wrap, drop of the atoms outside [min_x, max_x), cell sort (cellKeySlabKernel)
int atomsCellSortSlab(mrmd_b200_atoms* a, int64_t end, const double* delta, const mrmd_b200_subdomain* sub, cudaStream_t st)
{
    const SubdomainDev sd = toDev(*sub);
    return atomsCellSortImpl(a, 0, end, delta, sub->minCorner, sub->maxCorner, nullptr, nullptr, st, &sd);
}

// cell sort that also removes the atoms flagged in dropFlags: they end up behind lcCellStart[numCells]
int atomsCellSortDrop(mrmd_b200_atoms* a, int64_t begin, int64_t end, const double* delta, const double* gridMin,
                      const double* gridMax, const signed char* dropFlags, cudaStream_t st)
{
    return atomsCellSortImpl(a, begin, end, delta, gridMin, gridMax, nullptr, dropFlags, st);
}

// prefix array of n sorted keys: cellStart[c] = offset + first position with key >= c, c in [0, numCells]
int cellStartFromKeys(const uint32_t* sortedKeys, int64_t n, int64_t numCells, int32_t* cellStart, int32_t offset,
                      cudaStream_t st)
{
    cellStartKernel<<<gridFor(n + 1, 256), 256, 0, st>>>(sortedKeys, n, numCells, cellStart, offset);
    MB_LAUNCHED();
    return 0;
}

// LinkedCellList over the centres of mass of the molecules [0, count) + permute for molecules of apm consecutive atoms
// (atomsOffset = apm * m for every molecule, before and after): the atoms move in blocks, the centres of mass are
// permuted along, and the linked-cell structure is left in m->lcView for the tiled neighbour build on molecules.
// dropFlags (optional, per molecule): flagged molecules sort behind the last cell (they migrated to another rank).
int moleculesCellSortWithAtoms(mrmd_b200_molecules* m, mrmd_b200_atoms* a, int64_t count, int apm, const double* delta,
                               const double* gridMin, const double* gridMax, const signed char* dropFlags, cudaStream_t st)
{
    MB_REQUIRE(m != nullptr && a != nullptr && apm >= 1 && count >= 0, "molecules_cell_sort_with_atoms");
    MB_REQUIRE(count <= m->capacity && count * apm <= a->capacity, "molecules_cell_sort_with_atoms: range outside the containers");
    if (m->lcView == nullptr) m->lcView = new mrmd_b200_atoms;
    mrmd_b200_atoms* lv = m->lcView;
    const GridDev g = makeGrid(gridMin, gridMax, delta);
    const int64_t numCells = int64_t(g.n[0]) * g.n[1] * g.n[2];
    MB_REQUIRE(numCells < (int64_t(1) << 31) - 1, "molecules_cell_sort_with_atoms: too many cells");
    MB_TRY(lv->lcCellStart.reserve(size_t(numCells + 1) * 4));
    if (count > 0)
    {
        uint32_t *k0, *v0, *k1, *v1, *hist;
        MB_TRY(cellSortPrepare(m->sortScratch, count, &k0, &v0, &k1, &v1, &hist));
        MB_TRY(atomsEnsureAlt(a, st));
        MB_TRY(molsEnsureAlt(m, st));
        cellKeyKernel<<<gridFor(count, 256), 256, 0, st>>>(m->v.pos, 0, count, g, k0, v0, nullptr, dropFlags,
                                                           static_cast<uint32_t>(numCells));
        MB_LAUNCHED();
        uint32_t *sortedKeys, *perm;
        MB_TRY(radixSortPairs(k0, v0, k1, v1, hist, count, bitsFor(numCells + 1), &sortedKeys, &perm, st));
        permuteAtomBlocksKernel<<<gridFor(count * apm, 256), 256, 0, st>>>(a->alt, a->v, perm, count, apm);
        MB_LAUNCHED();
        std::swap(a->v, a->alt);
        // only the centres of mass travel with the molecules: offsets stay apm * m, lambda / force are per-step values
        permutePos4Kernel<<<gridFor(count, 256), 256, 0, st>>>(m->alt.pos, m->v.pos, perm, count);
        MB_LAUNCHED();
        MB_CUDA(cudaMemcpyAsync(m->v.pos, m->alt.pos, size_t(count) * 32, cudaMemcpyDeviceToDevice, st));
        cellStartKernel<<<gridFor(count + 1, 256), 256, 0, st>>>(sortedKeys, count, numCells, lv->lcCellStart.as<int32_t>(), 0);
        MB_LAUNCHED();
    }
    else
        MB_CUDA(cudaMemsetAsync(lv->lcCellStart.p, 0, size_t(numCells + 1) * 4, st));
    a->posEpoch += 1;
    a->lcValid = false;  // the atoms are in molecule order, not in an atom cell order
    lv->v.pos = m->v.pos;
    lv->lcValid = true;
    lv->lcGrid = g;
    lv->lcBegin = 0;
    lv->lcEnd = count;
    lv->numLocal = count;
    lv->size = count;
    lv->lcNumCells = numCells;
    lv->lcEpoch += 1;
    lv->lcPosEpoch = lv->posEpoch;
    return 0;
}

static int atomsCellSortImpl(mrmd_b200_atoms* a, int64_t begin, int64_t end, const double* delta, const double* gridMin,
                             const double* gridMax, int32_t* cellIdOut, const signed char* dropFlags, cudaStream_t st,
                             const SubdomainDev* slabWrap)
{
    MB_REQUIRE(a != nullptr && delta != nullptr && gridMin != nullptr && gridMax != nullptr, "atoms_cell_sort");
    MB_REQUIRE(begin >= 0 && begin <= end && end <= a->size, "atoms_cell_sort: range outside the container");
    const int64_t count = end - begin;
    if (count == 0) return 0;
    const GridDev g = makeGrid(gridMin, gridMax, delta);
    const int64_t numCells = int64_t(g.n[0]) * g.n[1] * g.n[2];
    MB_REQUIRE(numCells < (int64_t(1) << 31) - 1, "atoms_cell_sort: too many cells");
    uint32_t *k0, *v0, *k1, *v1, *hist;
    MB_TRY(cellSortPrepare(a->sortScratch, count, &k0, &v0, &k1, &v1, &hist));
    MB_TRY(atomsEnsureAlt(a, st));
    if (slabWrap != nullptr)
        cellKeySlabKernel<<<gridFor(count, 256), 256, 0, st>>>(a->v.pos, count, g, *slabWrap, k0, v0, static_cast<uint32_t>(numCells));
    else
        cellKeyKernel<<<gridFor(count, 256), 256, 0, st>>>(a->v.pos, begin, count, g, k0, v0, cellIdOut, dropFlags,
                                                           static_cast<uint32_t>(numCells));
    MB_LAUNCHED();
    uint32_t *sortedKeys, *perm;
    MB_TRY(radixSortPairs(k0, v0, k1, v1, hist, count, bitsFor(numCells + 1), &sortedKeys, &perm, st));
    permuteAtomsKernel<<<gridFor(a->size, 256), 256, 0, st>>>(a->alt, a->v, perm, begin, end, a->size);
    MB_LAUNCHED();
    std::swap(a->v, a->alt);
    // keep the linked-cell structure for the tiled (shared-memory staged) neighbour and force kernels
    MB_TRY(a->lcCellStart.reserve(size_t(numCells + 1) * 4));
    cellStartKernel<<<gridFor(count + 1, 256), 256, 0, st>>>(sortedKeys, count, numCells, a->lcCellStart.as<int32_t>(),
                                                                static_cast<int32_t>(begin));
    MB_LAUNCHED();
    a->lcValid = true;
    a->lcGrid = g;
    a->lcBegin = begin;
    a->lcEnd = end;
    a->lcNumCells = numCells;
    a->lcEpoch += 1;
    a->lcPosEpoch = a->posEpoch;
    return 0;
}
}  // namespace mrmd_b200

extern "C" {

int mrmd_b200_atoms_cell_sort(mrmd_b200_atoms* a, int64_t begin, int64_t end, const double* delta,
                              const double* gridMin, const double* gridMax, int32_t* cellIdOut, void* stream)
{
    MB_TRY(checkDevice());
    return atomsCellSortImpl(a, begin, end, delta, gridMin, gridMax, cellIdOut, nullptr, S(stream));
}

int mrmd_b200_molecules_cell_sort_with_atoms(mrmd_b200_molecules* m, mrmd_b200_atoms* a, int atomsPerMolecule,
                                             const double* delta, const double* gridMin, const double* gridMax, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(m != nullptr && a != nullptr && delta != nullptr && gridMin != nullptr && gridMax != nullptr,
               "molecules_cell_sort_with_atoms");
    MB_REQUIRE(atomsPerMolecule >= 1 && m->numLocal * atomsPerMolecule == a->numLocal,
               "molecules_cell_sort_with_atoms: local atoms != atomsPerMolecule x local molecules");
    return moleculesCellSortWithAtoms(m, a, m->numLocal, atomsPerMolecule, delta, gridMin, gridMax, nullptr, S(stream));
}

int mrmd_b200_molecules_cell_sort(mrmd_b200_molecules* m, int64_t begin, int64_t end, const double* delta,
                                  const double* gridMin, const double* gridMax, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(m != nullptr && delta != nullptr && gridMin != nullptr && gridMax != nullptr, "molecules_cell_sort");
    MB_REQUIRE(begin >= 0 && begin <= end && end <= m->size, "molecules_cell_sort: range outside the container");
    const int64_t count = end - begin;
    if (count == 0) return 0;
    cudaStream_t st = S(stream);
    const GridDev g = makeGrid(gridMin, gridMax, delta);
    const int64_t numCells = int64_t(g.n[0]) * g.n[1] * g.n[2];
    uint32_t *k0, *v0, *k1, *v1, *hist;
    MB_TRY(cellSortPrepare(m->sortScratch, count, &k0, &v0, &k1, &v1, &hist));
    MB_TRY(molsEnsureAlt(m, st));
    cellKeyKernel<<<gridFor(count, 256), 256, 0, st>>>(m->v.pos, begin, count, g, k0, v0, nullptr);
    MB_LAUNCHED();
    uint32_t *sortedKeys, *perm;
    MB_TRY(radixSortPairs(k0, v0, k1, v1, hist, count, bitsFor(numCells), &sortedKeys, &perm, st));
    permuteMolsKernel<<<gridFor(m->size, 256), 256, 0, st>>>(m->alt, m->v, perm, begin, end, m->size);
    MB_LAUNCHED();
    std::swap(m->v, m->alt);
    return 0;
}

int mrmd_b200_verlet_create(mrmd_b200_verlet** out, int half)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(out != nullptr, "verlet_create");
    auto* v = new mrmd_b200_verlet;
    v->half = half ? 1 : 0;
    *out = v;
    return 0;
}

int mrmd_b200_verlet_destroy(mrmd_b200_verlet* v)
{
    if (v == nullptr) return 0;
    cudaDeviceSynchronize();
    v->counts.release();
    v->neigh.release();
    v->enc.release();
    v->tileDesc.release();
    v->cellLoHi.release();
    for (int b = 0; b < 2; ++b)
    {
        v->keys[b].release();
        v->vals[b].release();
    }
    v->scratch.release();
    v->cellStart.release();
    v->sortedPos.release();
    v->stats.release();
    v->tstats.release();
    v->tileActive.release();
    v->activeTiles.release();
    if (v->hStats) cudaFreeHost(v->hStats);
    if (v->hTstats) cudaFreeHost(v->hTstats);
    delete v;
    return 0;
}

int mrmd_b200_verlet_build_atoms(mrmd_b200_verlet* v, const mrmd_b200_atoms* a, int64_t begin, int64_t end,
                                 double radius, double cellRatio, const double* gridMin, const double* gridMax,
                                 int64_t maxNeigh, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(a != nullptr, "verlet_build_atoms");
    return verletBuild(v, a->v.pos, a->size, begin, end, radius, cellRatio, gridMin, gridMax, maxNeigh, S(stream));
}

int mrmd_b200_verlet_build_molecules(mrmd_b200_verlet* v, const mrmd_b200_molecules* m, int64_t begin, int64_t end,
                                     double radius, double cellRatio, const double* gridMin, const double* gridMax,
                                     int64_t maxNeigh, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(m != nullptr, "verlet_build_molecules");
    return verletBuild(v, m->v.pos, m->size, begin, end, radius, cellRatio, gridMin, gridMax, maxNeigh, S(stream));
}

int mrmd_b200_verlet_info(const mrmd_b200_verlet* v, int64_t* numParticles, int64_t* width, int64_t* totalPairs,
                          int* half)
{
    MB_REQUIRE(v != nullptr, "verlet_info");
    if (numParticles) *numParticles = v->numParticles;
    if (width) *width = v->width;
    if (totalPairs)
    {
        int64_t total = 0;
        if (v->hStats != nullptr) std::memcpy(&total, v->hStats + 2, 8);
        *totalPairs = total;
    }
    if (half) *half = v->half;
    return 0;
}

int mrmd_b200_verlet_read(const mrmd_b200_verlet* v, int32_t* counts, int32_t* neighbors, int memKind, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(v != nullptr, "verlet_read");
    cudaStream_t st = S(stream);
    const int64_t n = v->numParticles;
    if (n == 0) return 0;
    const cudaMemcpyKind kind = (memKind == MRMD_B200_MEM_HOST) ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (counts != nullptr) MB_CUDA(cudaMemcpyAsync(counts, v->counts.p, size_t(n) * 4, kind, st));
    if (neighbors != nullptr)
    {
        auto* self = const_cast<mrmd_b200_verlet*>(v);
        int32_t* rowMajor = neighbors;
        if (memKind == MRMD_B200_MEM_HOST)
        {
            // reuse the radix-sort ping-pong buffer as staging only if large enough, else a fresh buffer
            MB_TRY(self->sortedPos.reserve(std::max(size_t(n) * 32, size_t(n) * v->width * 4)));
            rowMajor = self->sortedPos.as<int32_t>();
        }
        transposeListKernel<<<gridFor(n * v->width, 256), 256, 0, st>>>(v->counts.as<int32_t>(), v->neigh.as<int32_t>(),
                                                                       n, v->width, v->pitch, rowMajor);
        MB_LAUNCHED();
        if (memKind == MRMD_B200_MEM_HOST)
            MB_CUDA(cudaMemcpyAsync(neighbors, rowMajor, size_t(n) * v->width * 4, cudaMemcpyDeviceToHost, st));
    }
    MB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
