// ghost.cu -- periodic self-ghosting.
// Reference: mrmd/communication/PeriodicMapping.cpp:30-58, GhostExchange.cpp:59-194, UpdateGhostAtoms.cpp:31-68,
//            AccumulateForce.cpp:25-47, MultiResRealAtomsExchange.cpp:23-73,
//            MultiResPeriodicGhostExchange.cpp:60-285.
//
// Ghost creation keeps the reference's deterministic order (the order of its parallel_scan): per axis the
// atoms below minInnerCorner are appended first (shifted +L), then those at or above maxInnerCorner
// (shifted -L).  Selection and copy are fused: block counts -> one-block scan -> ranked copy, no atomics.
#include <algorithm>

#include "handles.cuh"


namespace mrmd_b200
{
constexpr int GH_THREADS = 256;

static int corrEnsure(mrmd_b200_ghost* g, int64_t capacity, int64_t keep, cudaStream_t st)
{
    if (capacity <= g->corrCapacity) return 0;
    const int64_t cap = capacity + capacity / 8 + 32;
    int64_t* q = nullptr;
    MB_CUDA(cudaMalloc(&q, size_t(cap) * 8));
    MB_CUDA(cudaMemsetAsync(q, 0xFF, size_t(cap) * 8, st));  // -1
    if (g->corr != nullptr)
    {
        if (keep > 0) MB_CUDA(cudaMemcpyAsync(q, g->corr, size_t(std::min(keep, g->corrCapacity)) * 8, cudaMemcpyDeviceToDevice, st));
        MB_CUDA(cudaStreamSynchronize(st));
        cudaFree(g->corr);
    }
    g->corr = q;
    g->corrCapacity = cap;
    return 0;
}

// communication/PeriodicMapping.cpp:30-58
__global__ void mapIntoDomainKernel(double4* pos, int64_t n, SubdomainDev s)
{
    const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (idx >= n) return;
    double4 p = ld4(pos + idx);
    double* x = &p.x;
#pragma unroll
    for (int dim = 0; dim < 3; ++dim)
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
}

// exclusive block scan of up to 4 counters held per thread; total[] valid in all threads afterwards
template <int K>
__device__ __forceinline__ void blockExclusiveScan(long long (&v)[K], long long (&total)[K])
{
    __shared__ long long sWarp[K][GH_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long incl[K];
#pragma unroll
    for (int k = 0; k < K; ++k)
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
    for (int k = 0; k < K; ++k)
    {
        long long off = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < GH_THREADS / 32; ++w)
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

// --- atom granular ghost creation --------------------------------------------------------------
__global__ void __launch_bounds__(GH_THREADS)
    ghostCountKernel(const double4* pos, int64_t n, int axis, double minInner, double maxInner, int64_t* blockCounts)
{
    const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    long long v[2] = {0, 0}, total[2];
    if (idx < n)
    {
        const double4 p = ld4nc(pos + idx);
        const double x = (axis == 0) ? p.x : ((axis == 1) ? p.y : p.z);
        v[0] = (x < minInner) ? 1 : 0;    // GhostExchange.cpp:80
        v[1] = (x >= maxInner) ? 1 : 0;   // :89
    }
    blockExclusiveScan<2>(v, total);
    if (threadIdx.x == 0)
    {
        blockCounts[2 * blockIdx.x + 0] = total[0];
        blockCounts[2 * blockIdx.x + 1] = total[1];
    }
}

// single block: exclusive scan over the per-block counters (K interleaved counters), totals to dTotals
template <int K>
__global__ void __launch_bounds__(GH_THREADS) ghostScanBlocksKernel(int64_t* blockCounts, int64_t numBlocks, int64_t* dTotals)
{
    __shared__ long long sCarry[K];
    if (threadIdx.x < K) sCarry[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t base = 0; base < numBlocks; base += GH_THREADS)
    {
        const int64_t b = base + threadIdx.x;
        long long v[K], total[K];
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = (b < numBlocks) ? blockCounts[K * b + k] : 0;
        blockExclusiveScan<K>(v, total);
        if (b < numBlocks)
        {
#pragma unroll
            for (int k = 0; k < K; ++k) blockCounts[K * b + k] = v[k] + sCarry[k];
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
#pragma unroll
            for (int k = 0; k < K; ++k) sCarry[k] += total[k];
        }
        __syncthreads();
    }
    if (threadIdx.x < K) dTotals[threadIdx.x] = sCarry[threadIdx.x];
}

__device__ __forceinline__ void copyAtomRecord(const AtomsView& a, int64_t dst, int64_t src, int axis, double shift)
{
    double4 p = ld4(a.pos + src);
    if (axis == 0) p.x += shift;
    else if (axis == 1) p.y += shift;
    else p.z += shift;
    st4(a.pos + dst, p);
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
        a.vel[d][dst] = a.vel[d][src];
        a.force[d][dst] = a.force[d][src];
    }
    a.mass[dst] = a.mass[src];
    a.charge[dst] = a.charge[src];
    a.relMass[dst] = a.relMass[src];
    a.gid[dst] = a.gid[src];
}

__device__ __forceinline__ int64_t rootOf(const int64_t* corr, int64_t realIdx)
{
    while (corr[realIdx] != -1) realIdx = corr[realIdx];  // GhostExchange.cpp:133-138
    return realIdx;
}

__global__ void __launch_bounds__(GH_THREADS)
    ghostCopyKernel(AtomsView a, int64_t n, int axis, double minInner, double maxInner, double diameter,
                    const int64_t* blockOffsets, const int64_t* dTotals, int64_t* corr)
{
    const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    long long v[2] = {0, 0}, total[2];
    bool low = false, high = false;
    if (idx < n)
    {
        const double4 p = ld4(a.pos + idx);
        const double x = (axis == 0) ? p.x : ((axis == 1) ? p.y : p.z);
        low = x < minInner;
        high = x >= maxInner;
        v[0] = low ? 1 : 0;
        v[1] = high ? 1 : 0;
    }
    blockExclusiveScan<2>(v, total);
    if (low)
    {
        const int64_t g = n + blockOffsets[2 * blockIdx.x + 0] + v[0];  // GhostExchange.cpp:123-140
        copyAtomRecord(a, g, idx, axis, +diameter);
        corr[g] = rootOf(corr, idx);
    }
    if (high)
    {
        const int64_t g = n + dTotals[0] + blockOffsets[2 * blockIdx.x + 1] + v[1];  // :142-160
        copyAtomRecord(a, g, idx, axis, -diameter);
        corr[g] = rootOf(corr, idx);
    }
}

// communication/UpdateGhostAtoms.cpp:31-68
__global__ void ghostUpdateKernel(double4* pos, int64_t numLocal, int64_t numGhost, const int64_t* corr, SubdomainDev s)
{
    const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (j >= numGhost) return;
    const int64_t idx = numLocal + j;
    const int64_t realIdx = corr[idx];
    double4 g = ld4(pos + idx);
    const double4 r = ld4(pos + realIdx);
    const double dx[3] = {g.x - r.x, g.y - r.y, g.z - r.z};
    double out[3] = {r.x, r.y, r.z};
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
        const double delta = 0.1 * s.diameter[d];
        if (dx[d] > +delta) out[d] += s.diameter[d];
        if (dx[d] < -delta) out[d] -= s.diameter[d];
    }
    g.x = out[0];
    g.y = out[1];
    g.z = out[2];
    st4(pos + idx, g);
}

// communication/AccumulateForce.cpp:25-47
__global__ void ghostFoldKernel(AtomsView a, int64_t numLocal, int64_t numGhost, const int64_t* corr)
{
    const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (j >= numGhost) return;
    const int64_t idx = numLocal + j;
    const int64_t realIdx = corr[idx];
    if (realIdx == -1) return;
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
        const double f = a.force[d][idx];
        if (f != 0.0) atomicAdd(a.force[d] + realIdx, f);
        a.force[d][idx] = 0.0;
    }
}

__global__ void fillInt64Kernel(int64_t* p, int64_t n, int64_t value)
{
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i < n) p[i] = value;
}

// --- molecule granular variants ------------------------------------------------------------------
// communication/MultiResRealAtomsExchange.cpp:23-73
__global__ void mrMapIntoDomainKernel(MolsView m, AtomsView a, int64_t numLocalMols, SubdomainDev s)
{
    const int64_t mi = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (mi >= numLocalMols) return;
    double4 mp = ld4(m.pos + mi);
    double* mx = &mp.x;
    const longlong2 oc = m.oc[mi];
    double shift[3] = {0.0, 0.0, 0.0};
    bool moved = false;
#pragma unroll
    for (int dim = 0; dim < 3; ++dim)
    {
        if (s.maxCorner[dim] <= mx[dim])
        {
            mx[dim] -= s.diameter[dim];
            shift[dim] -= 1.0;
            moved = true;
        }
        if (mx[dim] < s.minCorner[dim])
        {
            mx[dim] += s.diameter[dim];
            shift[dim] += 1.0;
            moved = true;
        }
    }
    if (!moved) return;
    st4(m.pos + mi, mp);
    for (long long ai = oc.x; ai < oc.x + oc.y; ++ai)
    {
        double4 p = ld4(a.pos + ai);
        double* x = &p.x;
#pragma unroll
        for (int dim = 0; dim < 3; ++dim)
        {
            // the reference applies -L then possibly +L as two separate additions
            if (shift[dim] < 0.0) x[dim] -= s.diameter[dim];
            if (shift[dim] > 0.0) x[dim] += s.diameter[dim];
        }
        st4(a.pos + ai, p);
    }
}

// counters (reference naming, MultiResPeriodicGhostExchange.cpp:23-43): 0 positiveMolecules (>= maxInner),
// 1 positiveAtoms, 2 negativeMolecules (< minInner), 3 negativeAtoms
__global__ void __launch_bounds__(GH_THREADS)
    mrGhostCountKernel(MolsView m, int64_t n, int axis, double minInner, double maxInner, int64_t* blockCounts)
{
    const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    long long v[4] = {0, 0, 0, 0}, total[4];
    if (idx < n)
    {
        const double4 p = ld4nc(m.pos + idx);
        const double x = (axis == 0) ? p.x : ((axis == 1) ? p.y : p.z);
        const long long na = m.oc[idx].y;
        if (x >= maxInner)
        {
            v[0] = 1;
            v[1] = na;
        }
        if (x < minInner)
        {
            v[2] = 1;
            v[3] = na;
        }
    }
    blockExclusiveScan<4>(v, total);
    if (threadIdx.x == 0)
        for (int k = 0; k < 4; ++k) blockCounts[4 * blockIdx.x + k] = total[k];
}

__device__ __forceinline__ void copyMolecule(const MolsView& m, const AtomsView& a, int64_t gm, int64_t src, int64_t ga,
                                             int axis, double shift, int64_t* corr)
{
    double4 mp = ld4(m.pos + src);
    if (axis == 0) mp.x += shift;
    else if (axis == 1) mp.y += shift;
    else mp.z += shift;
    st4(m.pos + gm, mp);
    st4(m.w + gm, ld4(m.w + src));
    m.lambda[gm] = m.lambda[src];
#pragma unroll
    for (int d = 0; d < 3; ++d) m.force[d][gm] = m.force[d][src];
    const longlong2 oc = m.oc[src];
    m.oc[gm] = make_longlong2(ga, oc.y);
    for (long long k = 0; k < oc.y; ++k)
    {
        copyAtomRecord(a, ga + k, oc.x + k, axis, shift);
        corr[ga + k] = rootOf(corr, oc.x + k);
    }
}

__global__ void __launch_bounds__(GH_THREADS)
    mrGhostCopyKernel(MolsView m, AtomsView a, int64_t nMols, int64_t nAtoms, int axis, double minInner, double maxInner,
                      double diameter, const int64_t* blockOffsets, const int64_t* dTotals, int64_t* corr)
{
    const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    long long v[4] = {0, 0, 0, 0}, total[4];
    bool pos = false, neg = false;
    if (idx < nMols)
    {
        const double4 p = ld4(m.pos + idx);
        const double x = (axis == 0) ? p.x : ((axis == 1) ? p.y : p.z);
        const long long na = m.oc[idx].y;
        pos = x >= maxInner;
        neg = x < minInner;
        if (pos)
        {
            v[0] = 1;
            v[1] = na;
        }
        if (neg)
        {
            v[2] = 1;
            v[3] = na;
        }
    }
    blockExclusiveScan<4>(v, total);
    const int64_t* off = blockOffsets + 4 * blockIdx.x;
    if (pos)  // appended first, shifted by -L (:150-190)
        copyMolecule(m, a, nMols + off[0] + v[0], idx, nAtoms + off[1] + v[1], axis, -diameter, corr);
    if (neg)  // then the ones below minInner, shifted by +L (:192-234)
        copyMolecule(m, a, nMols + dTotals[0] + off[2] + v[2], idx, nAtoms + dTotals[1] + off[3] + v[3], axis, +diameter,
                     corr);
}

static int ghostAxis(mrmd_b200_ghost* g, mrmd_b200_atoms* a, const mrmd_b200_subdomain* s, int axis, cudaStream_t st)
{
    const int64_t n = a->numLocal + a->numGhost;
    if (n == 0) return 0;
    const int numBlocks = gridFor(n, GH_THREADS);
    MB_TRY(g->blockCounts.reserve(size_t(numBlocks) * 2 * 8));
    int64_t* bc = g->blockCounts.as<int64_t>();
    ghostCountKernel<<<numBlocks, GH_THREADS, 0, st>>>(a->v.pos, n, axis, s->minInnerCorner[axis], s->maxInnerCorner[axis], bc);
    MB_LAUNCHED();
    ghostScanBlocksKernel<2><<<1, GH_THREADS, 0, st>>>(bc, numBlocks, g->dTotals);
    MB_LAUNCHED();
    MB_CUDA(cudaMemcpyAsync(g->hTotals, g->dTotals, 16, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));  // the reference reads the two counts on the host too (:107-108)
    const int64_t n0 = g->hTotals[0], n1 = g->hTotals[1];
    // atoms.resize(numLocal + numGhost + n0 + n1), GhostExchange.cpp:111-117
    if (a->size < n) a->size = n;
    MB_TRY(atomsEnsureCapacity(a, n + n0 + n1, st));
    a->size = n + n0 + n1;
    MB_TRY(corrEnsure(g, a->capacity, n, st));
    if (n0 + n1 > 0)
    {
        ghostCopyKernel<<<numBlocks, GH_THREADS, 0, st>>>(a->v, n, axis, s->minInnerCorner[axis], s->maxInnerCorner[axis],
                                                          s->diameter[axis], bc, g->dTotals, g->corr);
        MB_LAUNCHED();
    }
    a->numGhost += n0 + n1;
    return 0;
}

static int ghostReset(mrmd_b200_ghost* g, int64_t capacity, cudaStream_t st)
{
    MB_TRY(corrEnsure(g, capacity, 0, st));
    MB_CUDA(cudaMemsetAsync(g->corr, 0xFF, size_t(g->corrCapacity) * 8, st));
    return 0;
}

static int mrGhostAxis(mrmd_b200_ghost* g, mrmd_b200_molecules* m, mrmd_b200_atoms* a, const mrmd_b200_subdomain* s,
                       int axis, cudaStream_t st)
{
    const int64_t nm = m->numLocal + m->numGhost;
    const int64_t na = a->numLocal + a->numGhost;
    if (nm == 0) return 0;
    const int numBlocks = gridFor(nm, GH_THREADS);
    MB_TRY(g->blockCounts.reserve(size_t(numBlocks) * 4 * 8));
    int64_t* bc = g->blockCounts.as<int64_t>();
    mrGhostCountKernel<<<numBlocks, GH_THREADS, 0, st>>>(m->v, nm, axis, s->minInnerCorner[axis], s->maxInnerCorner[axis], bc);
    MB_LAUNCHED();
    ghostScanBlocksKernel<4><<<1, GH_THREADS, 0, st>>>(bc, numBlocks, g->dTotals);
    MB_LAUNCHED();
    MB_CUDA(cudaMemcpyAsync(g->hTotals, g->dTotals, 32, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    const int64_t posM = g->hTotals[0], posA = g->hTotals[1], negM = g->hTotals[2], negA = g->hTotals[3];
    if (m->size < nm) m->size = nm;
    if (a->size < na) a->size = na;
    MB_TRY(molsEnsureCapacity(m, nm + posM + negM, st));
    m->size = nm + posM + negM;
    MB_TRY(atomsEnsureCapacity(a, na + posA + negA, st));
    a->size = na + posA + negA;
    MB_TRY(corrEnsure(g, a->capacity, na, st));
    if (posM + negM > 0)
    {
        mrGhostCopyKernel<<<numBlocks, GH_THREADS, 0, st>>>(m->v, a->v, nm, na, axis, s->minInnerCorner[axis],
                                                            s->maxInnerCorner[axis], s->diameter[axis], bc, g->dTotals,
                                                            g->corr);
        MB_LAUNCHED();
    }
    m->numGhost += posM + negM;
    a->numGhost += posA + negA;
    return 0;
}
}  // namespace mrmd_b200

using namespace mrmd_b200;

extern "C" {

int mrmd_b200_ghost_create(mrmd_b200_ghost** out)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(out != nullptr, "ghost_create");
    auto* g = new mrmd_b200_ghost;
    if (cudaMalloc(&g->dTotals, 32) != cudaSuccess || cudaMallocHost(&g->hTotals, 32) != cudaSuccess)
    {
        delete g;
        setLastError("ghost_create: allocation failed");
        return MRMD_B200_ENOMEM;
    }
    *out = g;
    return 0;
}

int mrmd_b200_ghost_destroy(mrmd_b200_ghost* g)
{
    if (g == nullptr) return 0;
    cudaDeviceSynchronize();
    if (g->corr) cudaFree(g->corr);
    if (g->dTotals) cudaFree(g->dTotals);
    if (g->hTotals) cudaFreeHost(g->hTotals);
    g->blockCounts.release();
    delete g;
    return 0;
}

int mrmd_b200_ghost_map_into_domain(mrmd_b200_atoms* a, const mrmd_b200_subdomain* s, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(a != nullptr && s != nullptr, "ghost_map_into_domain");
    a->posEpoch += 1;
    if (a->numLocal == 0) return 0;
    mapIntoDomainKernel<<<gridFor(a->numLocal, 256), 256, 0, S(stream)>>>(a->v.pos, a->numLocal, toDev(*s));
    MB_LAUNCHED();
    return 0;
}

int mrmd_b200_ghost_reset(mrmd_b200_ghost* g, mrmd_b200_atoms* a, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(g != nullptr && a != nullptr, "ghost_reset");
    return ghostReset(g, std::max(a->capacity, a->numLocal), S(stream));
}

int mrmd_b200_ghost_create_atoms(mrmd_b200_ghost* g, mrmd_b200_atoms* a, const mrmd_b200_subdomain* s, int axis,
                                 void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(g != nullptr && a != nullptr && s != nullptr && axis < 3, "ghost_create_atoms");
    cudaStream_t st = S(stream);
    if (axis >= 0)
    {
        MB_TRY(corrEnsure(g, a->capacity, a->numLocal + a->numGhost, st));
        return ghostAxis(g, a, s, axis, st);
    }
    // createGhostAtomsXYZ, GhostExchange.cpp:171-181
    MB_TRY(ghostReset(g, a->capacity, st));
    a->numGhost = 0;
    for (int ax = 0; ax < 3; ++ax) MB_TRY(ghostAxis(g, a, s, ax, st));
    return 0;
}

int mrmd_b200_ghost_update(mrmd_b200_ghost* g, mrmd_b200_atoms* a, const mrmd_b200_subdomain* s, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(g != nullptr && a != nullptr && s != nullptr, "ghost_update");
    if (a->numGhost == 0) return 0;
    MB_REQUIRE(g->corrCapacity >= a->numLocal + a->numGhost, "ghost_update: no correspondingRealAtom for these atoms");
    ghostUpdateKernel<<<gridFor(a->numGhost, 256), 256, 0, S(stream)>>>(a->v.pos, a->numLocal, a->numGhost, g->corr, toDev(*s));
    MB_LAUNCHED();
    return 0;
}

int mrmd_b200_ghost_contribute_back(mrmd_b200_ghost* g, mrmd_b200_atoms* a, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(g != nullptr && a != nullptr, "ghost_contribute_back");
    if (a->numGhost == 0) return 0;
    MB_REQUIRE(g->corrCapacity >= a->numLocal + a->numGhost, "ghost_contribute_back: no correspondingRealAtom for these atoms");
    ghostFoldKernel<<<gridFor(a->numGhost, 256), 256, 0, S(stream)>>>(a->v, a->numLocal, a->numGhost, g->corr);
    MB_LAUNCHED();
    return 0;
}

int mrmd_b200_ghost_read_corresponding(const mrmd_b200_ghost* g, int64_t* dst, int64_t first, int64_t count, int memKind,
                                       void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(g != nullptr && dst != nullptr && first >= 0 && count >= 0 && first + count <= g->corrCapacity,
               "ghost_read_corresponding");
    if (count == 0) return 0;
    MB_CUDA(cudaMemcpyAsync(dst, g->corr + first, size_t(count) * 8,
                            memKind == MRMD_B200_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, S(stream)));
    MB_CUDA(cudaStreamSynchronize(S(stream)));
    return 0;
}

int mrmd_b200_ghost_write_corresponding(mrmd_b200_ghost* g, const int64_t* src, int64_t first, int64_t count, int memKind,
                                        void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(g != nullptr && src != nullptr && first >= 0 && count >= 0, "ghost_write_corresponding");
    MB_TRY(corrEnsure(g, first + count, g->corrCapacity, S(stream)));
    if (count == 0) return 0;
    MB_CUDA(cudaMemcpyAsync(g->corr + first, src, size_t(count) * 8,
                            memKind == MRMD_B200_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, S(stream)));
    MB_CUDA(cudaStreamSynchronize(S(stream)));
    return 0;
}

int mrmd_b200_ghost_mr_map_into_domain(mrmd_b200_molecules* m, mrmd_b200_atoms* a, const mrmd_b200_subdomain* s,
                                       void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(m != nullptr && a != nullptr && s != nullptr, "ghost_mr_map_into_domain");
    a->posEpoch += 1;
    if (m->numLocal == 0) return 0;
    mrMapIntoDomainKernel<<<gridFor(m->numLocal, 256), 256, 0, S(stream)>>>(m->v, a->v, m->numLocal, toDev(*s));
    MB_LAUNCHED();
    return 0;
}

int mrmd_b200_ghost_mr_create_atoms(mrmd_b200_ghost* g, mrmd_b200_molecules* m, mrmd_b200_atoms* a,
                                    const mrmd_b200_subdomain* s, int axis, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(g != nullptr && m != nullptr && a != nullptr && s != nullptr && axis < 3, "ghost_mr_create_atoms");
    cudaStream_t st = S(stream);
    if (axis >= 0)
    {
        MB_TRY(corrEnsure(g, a->capacity, a->numLocal + a->numGhost, st));
        return mrGhostAxis(g, m, a, s, axis, st);
    }
    // createGhostAtomsXYZ, MultiResPeriodicGhostExchange.cpp:250-265
    MB_TRY(ghostReset(g, a->capacity, st));
    m->numGhost = 0;
    a->numGhost = 0;
    for (int ax = 0; ax < 3; ++ax) MB_TRY(mrGhostAxis(g, m, a, s, ax, st));
    return 0;
}

}  // extern "C"
