/*
 * mrmd_oracle.cpp -- CPU parity oracle (TEST INFRASTRUCTURE, see mrmd_oracle.h).
 *
 * Restates, function by function, the arithmetic of the MRMD hot path.  All
 * citations are relative to the reference tree (XzzX/mrmd).  Build with
 * -ffp-contract=off: the neighbour cutoff test, the cell index and the
 * histogram bin must not be contracted into FMAs.
 *
 * Parity: MRMD arithmetic pinned by the reference goldens (tests/golden);
 * Cabana grid arithmetic / list cutoff inclusivity / in-cell order and the
 * Langevin random stream are "parity unpinned" (not in the reference tree).
 */
#include "mrmd_oracle.h"

#include <omp.h>

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numbers>
#include <vector>

namespace
{
constexpr double PI = 3.14159265358979323846;  // mrmd/constants.hpp

inline const double* P(const double* pos, int64_t stride, int64_t i) { return pos + i * stride; }

// ---------------------------------------------------------------------------
// predicates: util/IsInSymmetricSlab.hpp:34-50 and the lambdas of
// examples/04_LennardJones_IdealGas_LocalCap.cpp:206-227
inline bool slab1(const or_pred_t& p, double x, double y, double z)
{
    const double c[3] = {x, y, z};
    const double dx = c[p.axis] - p.center;
    const double absDx = std::abs(dx);
    return (absDx >= p.slabMin - p.tolerance && absDx <= p.slabMax + p.tolerance);
}
inline bool pred1(const or_pred_t* p, double x, double y, double z)
{
    if (p == nullptr) return true;
    switch (p->kind)
    {
        case OR_PRED_ALWAYS: return true;
        case OR_PRED_NEVER: return false;
        case OR_PRED_INTERVAL:
        {
            const double c[3] = {x, y, z};
            return c[p->axis] > p->slabMin && c[p->axis] < p->slabMax;
        }
        default: return slab1(*p, x, y, z);
    }
}
inline bool pred2(const or_pred_t* p, const double* a, const double* b)
{
    if (p == nullptr) return true;
    switch (p->kind)
    {
        case OR_PRED_ALWAYS: return true;
        case OR_PRED_NEVER: return false;
        case OR_PRED_SLAB_EITHER: return slab1(*p, a[0], a[1], a[2]) || slab1(*p, b[0], b[1], b[2]);
        case OR_PRED_SLAB_BOTH: return slab1(*p, a[0], a[1], a[2]) && slab1(*p, b[0], b[1], b[2]);
        default: return slab1(*p, a[0], a[1], a[2]);
    }
}

// action/LennardJones.hpp:53-78
inline void ljForceEnergy(const or_lj_type_t& t, double distSqr, double& ff, double& e)
{
    if (distSqr >= t.cappingDistanceSqr)
    {
        const double frac2 = 1.0 / distSqr;
        const double frac6 = frac2 * frac2 * frac2;
        ff = frac6 * (t.ff1 * frac6 - t.ff2) * frac2;
        e = frac6 * (t.ef1 * frac6 - t.ef2) - t.shift;
        return;
    }
    const double dist = std::sqrt(distSqr);
    ff = t.cappingCoeff / dist;
    e = t.energyAtCappingPoint - (dist - t.cappingDistance) * t.cappingCoeff - t.shift;
}

// data/MultiHistogram.hpp:60-66
inline int64_t histBin(double min, double inverseBinSize, int64_t numBins, double val)
{
    int64_t bin = static_cast<int64_t>(std::floor((val - min) * inverseBinSize));
    if (bin < 0) bin = -1;
    if (bin >= numBins) bin = -1;
    return bin;
}

// util/math.hpp:31-46
inline double powInt(double x, int64_t n)
{
    double ww = x;
    double yy = 1.0;
    for (int64_t nn = (n > 0) ? n : -n; nn != 0; nn >>= 1)
    {
        if ((nn & 1) == 1) yy *= ww;
        ww *= ww;
    }
    return (n > 0) ? yy : 1.0 / yy;
}

// weighting_function/CheckRegion.hpp:25-38 (REGION_CHECK_EPSILON = 0)
inline bool inAT(double l) { return l >= 1.0; }
inline bool inCG(double l) { return l <= 0.0; }
inline bool inHY(double l) { return !inAT(l) && !inCG(l); }

// Cabana 0.7 CartesianGrid (published algorithm; not in the reference tree)
struct Grid
{
    double min[3], max[3], dx[3], rdx[3];
    int n[3];
    static int cellsBetween(double mx, double mn, double rdelta) { return static_cast<int>(std::floor((mx - mn) * rdelta)); }
    Grid(const double* gmin, const double* gmax, const double* delta)
    {
        for (int d = 0; d < 3; ++d)
        {
            min[d] = gmin[d];
            max[d] = gmax[d];
            n[d] = cellsBetween(gmax[d], gmin[d], 1.0 / delta[d]);
            if (n[d] < 1) n[d] = 1;
            dx[d] = (gmax[d] - gmin[d]) / n[d];
            rdx[d] = 1.0 / dx[d];
        }
    }
    int locate1(double x, int d) const
    {
        int c = cellsBetween(x, min[d], rdx[d]);
        c = (c == n[d]) ? c - 1 : c;
        // memory-safety clamp for out-of-grid points (undefined in Cabana)
        if (c < 0) c = 0;
        if (c > n[d] - 1) c = n[d] - 1;
        return c;
    }
    int cardinal(int i, int j, int k) const { return (i * n[1] + j) * n[2] + k; }
    int cellOf(const double* x) const { return cardinal(locate1(x[0], 0), locate1(x[1], 1), locate1(x[2], 2)); }
    int64_t numCells() const { return int64_t(n[0]) * n[1] * n[2]; }
    double minDistanceToPoint(const double* x, int i, int j, int k) const
    {
        const int c[3] = {i, j, k};
        double r[3];
        for (int d = 0; d < 3; ++d)
        {
            const double xc = min[d] + (c[d] + 0.5) * dx[d];
            const double rr = std::fabs(x[d] - xc) - 0.5 * dx[d];
            r[d] = (rr > 0.0) ? rr : 0.0;
        }
        return r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    }
};

void stableCellSort(const int32_t* cellId, int64_t begin, int64_t end, int64_t numCells, int64_t* perm,
                    int64_t* offsets)
{
    std::fill(offsets, offsets + numCells + 1, int64_t(0));
    for (int64_t i = begin; i < end; ++i) offsets[cellId[i] + 1] += 1;
    for (int64_t c = 0; c < numCells; ++c) offsets[c + 1] += offsets[c];
    std::vector<int64_t> cursor(offsets, offsets + numCells);
    for (int64_t i = begin; i < end; ++i) perm[cursor[cellId[i]]++] = i;
}

inline void atomicAdd(double& target, double v)
{
#pragma omp atomic
    target += v;
}
}  // namespace

extern "C" {

void or_set_threads(int n) { omp_set_num_threads(n); }
int or_get_max_threads(void) { return omp_get_max_threads(); }

// data/Subdomain.hpp:41-62
void or_subdomain_init(or_subdomain_t* s, const double* minCorner, const double* maxCorner, const double* thickness)
{
    for (int d = 0; d < 3; ++d)
    {
        s->minCorner[d] = minCorner[d];
        s->maxCorner[d] = maxCorner[d];
        s->ghostLayerThickness[d] = thickness[d];
        s->minGhostCorner[d] = s->minCorner[d] - s->ghostLayerThickness[d];
        s->maxGhostCorner[d] = s->maxCorner[d] + s->ghostLayerThickness[d];
        s->minInnerCorner[d] = s->minCorner[d] + s->ghostLayerThickness[d];
        s->maxInnerCorner[d] = s->maxCorner[d] - s->ghostLayerThickness[d];
        s->diameter[d] = s->maxCorner[d] - s->minCorner[d];
        s->diameterWithGhostLayer[d] = s->maxCorner[d] - s->minCorner[d] + 2.0 * s->ghostLayerThickness[d];
    }
}

// data/Subdomain.cpp:24-32
void or_subdomain_scale_dim(or_subdomain_t* s, double factor, int axis)
{
    double mn[3], mx[3], th[3];
    for (int d = 0; d < 3; ++d)
    {
        mn[d] = s->minCorner[d];
        mx[d] = s->maxCorner[d];
        th[d] = s->ghostLayerThickness[d];
    }
    mn[axis] *= factor;
    mx[axis] *= factor;
    or_subdomain_init(s, mn, mx, th);
}

// ---------------------------------------------------------------------------
// action/LennardJones.cpp:59-116 (host precompute + the 1-thread-per-type init kernel)
void or_lj_init(or_lj_type_t* out, const double* cappingDistance, const double* rc, const double* sigma,
                const double* epsilon, int64_t numTypes, int isShifted)
{
    for (int64_t t = 0; t < numTypes * numTypes; ++t)
    {
        or_lj_type_t v;
        std::memset(&v, 0, sizeof(v));
        const double sig2 = sigma[t] * sigma[t];
        const double sig6 = sig2 * sig2 * sig2;
        v.ff1 = 48.0 * epsilon[t] * sig6 * sig6;
        v.ff2 = 24.0 * epsilon[t] * sig6;
        v.ef1 = 4.0 * epsilon[t] * sig6 * sig6;
        v.ef2 = 4.0 * epsilon[t] * sig6;
        v.rcSqr = rc[t] * rc[t];
        v.cappingDistance = cappingDistance[t];
        // operator()(typeIdx), LennardJones.cpp:62-79
        const double capDist = v.cappingDistance;
        v.cappingDistance = 0.0;
        v.cappingDistanceSqr = 0.0;
        double ff, e;
        ljForceEnergy(v, capDist * capDist, ff, e);
        v.cappingCoeff = ff * capDist;
        v.energyAtCappingPoint = e;
        v.cappingDistance = capDist;
        v.cappingDistanceSqr = capDist * capDist;
        if (isShifted)
        {
            ljForceEnergy(v, v.rcSqr, ff, e);
            v.shift = e;
        }
        out[t] = v;
    }
}

void or_lj_force_energy(const or_lj_type_t* table, int64_t typeIdx, double distSqr, double* forceFactor, double* energy)
{
    ljForceEnergy(table[typeIdx], distSqr, *forceFactor, *energy);
}

// action/LennardJones.hpp:135-206
int64_t or_lj_apply(or_atom_t* atoms, int64_t numLocal, const int32_t* counts, const int32_t* neigh, int64_t width,
                    const or_lj_type_t* table, double rcSqr, int64_t numTypesQuirk, const or_pred_t* pred,
                    double* energyVirial)
{
    double energy = 0.0;
    double virial = 0.0;
    int64_t pairs = 0;
#pragma omp parallel for schedule(static) reduction(+ : energy, virial, pairs)
    for (int64_t idx = 0; idx < numLocal; ++idx)
    {
        const double posTmp[3] = {atoms[idx].pos[0], atoms[idx].pos[1], atoms[idx].pos[2]};
        double forceTmp[3] = {0.0, 0.0, 0.0};
        const int64_t numNeighbors = counts[idx];
        for (int64_t n = 0; n < numNeighbors; ++n)
        {
            const int64_t jdx = neigh[idx * width + n];
            if (!pred2(pred, posTmp, atoms[jdx].pos)) continue;
            const double dx = posTmp[0] - atoms[jdx].pos[0];
            const double dy = posTmp[1] - atoms[jdx].pos[1];
            const double dz = posTmp[2] - atoms[jdx].pos[2];
            const double distSqr = dx * dx + dy * dy + dz * dz;
            if (distSqr > rcSqr) continue;
            const int64_t typeIdx = atoms[idx].type * numTypesQuirk + atoms[jdx].type;
            double ff, e;
            ljForceEnergy(table[typeIdx], distSqr, ff, e);
            energy += e;
            virial -= 0.5 * ff * distSqr;
            pairs += 1;
            forceTmp[0] += dx * ff;
            forceTmp[1] += dy * ff;
            forceTmp[2] += dz * ff;
            atomicAdd(atoms[jdx].force[0], -(dx * ff));
            atomicAdd(atoms[jdx].force[1], -(dy * ff));
            atomicAdd(atoms[jdx].force[2], -(dz * ff));
        }
        atomicAdd(atoms[idx].force[0], forceTmp[0]);
        atomicAdd(atoms[idx].force[1], forceTmp[1]);
        atomicAdd(atoms[idx].force[2], forceTmp[2]);
    }
    energyVirial[0] = energy;
    energyVirial[1] = virial;
    return pairs;
}

// ---------------------------------------------------------------------------
// Cabana::LinkedCellList binning as used in tests/NVT/NVT.cpp:136-144
int64_t or_cell_ids(const double* pos, int64_t stride, int64_t begin, int64_t end, const double* delta,
                    const double* gmin, const double* gmax, int32_t* cellId, int32_t* dims)
{
    const Grid grid(gmin, gmax, delta);
#pragma omp parallel for schedule(static)
    for (int64_t i = begin; i < end; ++i) cellId[i] = grid.cellOf(P(pos, stride, i));
    if (dims != nullptr)
        for (int d = 0; d < 3; ++d) dims[d] = grid.n[d];
    return grid.numCells();
}

void or_cell_perm(const int32_t* cellId, int64_t begin, int64_t end, int64_t numCells, int64_t* perm,
                  int64_t* cellOffsets)
{
    stableCellSort(cellId, begin, end, numCells, perm, cellOffsets);
}

// Cabana::permute semantics: slot begin+k receives the record perm[k]
void or_permute_atoms(or_atom_t* atoms, int64_t begin, int64_t end, const int64_t* perm)
{
    std::vector<or_atom_t> tmp(static_cast<size_t>(end - begin));
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < end - begin; ++k) tmp[k] = atoms[perm[k]];
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < end - begin; ++k) atoms[begin + k] = tmp[k];
}

void or_permute_molecules(or_molecule_t* mols, int64_t begin, int64_t end, const int64_t* perm)
{
    std::vector<or_molecule_t> tmp(static_cast<size_t>(end - begin));
    for (int64_t k = 0; k < end - begin; ++k) tmp[k] = mols[perm[k]];
    for (int64_t k = 0; k < end - begin; ++k) mols[begin + k] = tmp[k];
}

// Cabana::VerletList<.., Half/FullNeighborTag, VerletLayout2D, TeamOpTag>::build
// (call sites: examples/02_LennardJones_NVE.cpp:156-163, tests/LennardJones/LennardJones.cpp:112-118)
int64_t or_verlet_build(const double* pos, int64_t stride, int64_t nAll, int64_t begin, int64_t end, double radius,
                        double ratio, const double* gmin, const double* gmax, int half, int64_t width,
                        int32_t* counts, int32_t* neigh)
{
    const double gridSize = ratio * radius;
    const double delta[3] = {gridSize, gridSize, gridSize};
    const Grid grid(gmin, gmax, delta);
    const int cellRange = static_cast<int>(std::ceil(1.0 / ratio));
    const double rsqr = radius * radius;
    const int64_t numCells = grid.numCells();

    // bin ALL particles (local + ghost), candidates for neighbours
    std::vector<int32_t> cellId(static_cast<size_t>(nAll));
    std::vector<int32_t> ci(static_cast<size_t>(nAll) * 3);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nAll; ++i)
    {
        const double* x = P(pos, stride, i);
        for (int d = 0; d < 3; ++d) ci[i * 3 + d] = grid.locate1(x[d], d);
        cellId[i] = grid.cardinal(ci[i * 3], ci[i * 3 + 1], ci[i * 3 + 2]);
    }
    std::vector<int64_t> perm(static_cast<size_t>(nAll));
    std::vector<int64_t> offsets(static_cast<size_t>(numCells) + 1);
    stableCellSort(cellId.data(), 0, nAll, numCells, perm.data(), offsets.data());

    std::fill(counts, counts + nAll, 0);
    int64_t maxCount = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(max : maxCount)
    for (int64_t pid = begin; pid < end; ++pid)
    {
        const double* xp = P(pos, stride, pid);
        const int i = ci[pid * 3], j = ci[pid * 3 + 1], k = ci[pid * 3 + 2];
        const int imin = std::max(0, i - cellRange), imax = std::min(grid.n[0], i + cellRange + 1);
        const int jmin = std::max(0, j - cellRange), jmax = std::min(grid.n[1], j + cellRange + 1);
        const int kmin = std::max(0, k - cellRange), kmax = std::min(grid.n[2], k + cellRange + 1);
        int64_t count = 0;
        for (int ii = imin; ii < imax; ++ii)
            for (int jj = jmin; jj < jmax; ++jj)
                for (int kk = kmin; kk < kmax; ++kk)
                {
                    if (!(grid.minDistanceToPoint(xp, ii, jj, kk) <= rsqr)) continue;
                    const int c = grid.cardinal(ii, jj, kk);
                    for (int64_t s = offsets[c]; s < offsets[c + 1]; ++s)
                    {
                        const int64_t nid = perm[s];
                        const double* xn = P(pos, stride, nid);
                        bool valid;
                        if (half)
                            valid = (pid != nid) &&
                                    ((xn[0] > xp[0]) ||
                                     ((xn[0] == xp[0]) && ((xn[1] > xp[1]) || ((xn[1] == xp[1]) && (xn[2] > xp[2])))));
                        else
                            valid = (pid != nid);
                        if (!valid) continue;
                        const double dx = xp[0] - xn[0];
                        const double dy = xp[1] - xn[1];
                        const double dz = xp[2] - xn[2];
                        const double distSqr = dx * dx + dy * dy + dz * dz;
                        if (distSqr <= rsqr)
                        {
                            if (count < width) neigh[pid * width + count] = static_cast<int32_t>(nid);
                            count += 1;
                        }
                    }
                }
        counts[pid] = static_cast<int32_t>(count);
        maxCount = std::max(maxCount, count);
    }
    return maxCount;
}

// tests/LennardJones/LennardJones.cpp:40-70
int64_t or_count_within_cutoff(const double* pos, int64_t stride, int64_t numLocal, int64_t numAll, double cutoff,
                               const double* box, int periodic)
{
    const double rcSqr = cutoff * cutoff;
    int64_t count = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : count)
    for (int64_t idx = 0; idx < numLocal; ++idx)
    {
        const double* a = P(pos, stride, idx);
        for (int64_t jdx = idx + 1; jdx < numAll; ++jdx)
        {
            const double* b = P(pos, stride, jdx);
            double dx = std::abs(a[0] - b[0]);
            if (periodic && (dx > box[0] * 0.5)) dx -= box[0];
            double dy = std::abs(a[1] - b[1]);
            if (periodic && (dy > box[1] * 0.5)) dy -= box[1];
            double dz = std::abs(a[2] - b[2]);
            if (periodic && (dz > box[2] * 0.5)) dz -= box[2];
            const double distSqr = dx * dx + dy * dy + dz * dz;
            if (distSqr < rcSqr) ++count;
        }
    }
    return count;
}

// ---------------------------------------------------------------------------
// communication/PeriodicMapping.cpp:30-58
void or_periodic_map(or_atom_t* atoms, int64_t numLocal, const or_subdomain_t* s)
{
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < numLocal; ++idx)
    {
        for (int dim = 0; dim < 3; ++dim)
        {
            double& x = atoms[idx].pos[dim];
            if (s->maxCorner[dim] <= x)
            {
                x -= s->diameter[dim];
                x = std::max(x, s->minCorner[dim]);
            }
            if (x < s->minCorner[dim])
            {
                x += s->diameter[dim];
                if (s->maxCorner[dim] <= x) x = s->minCorner[dim];
            }
        }
    }
}

// communication/GhostExchange.cpp:59-169
int64_t or_ghost_create_axis(or_atom_t* atoms, int64_t numLocal, int64_t numGhost, int64_t capacity,
                             const or_subdomain_t* s, int axis, int64_t* corr)
{
    std::vector<int64_t> low, high;
    const int64_t n = numLocal + numGhost;
    for (int64_t idx = 0; idx < n; ++idx)  // the ordered parallel_scan, :76-106
    {
        if (atoms[idx].pos[axis] < s->minInnerCorner[axis]) low.push_back(idx);
        if (atoms[idx].pos[axis] >= s->maxInnerCorner[axis]) high.push_back(idx);
    }
    const int64_t n0 = static_cast<int64_t>(low.size());
    const int64_t n1 = static_cast<int64_t>(high.size());
    if (n + n0 + n1 > capacity) return -1;
    auto root = [&](int64_t realIdx) {
        while (corr[realIdx] != -1) realIdx = corr[realIdx];
        return realIdx;
    };
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n0; ++k)
    {
        const int64_t g = n + k;
        atoms[g] = atoms[low[k]];
        atoms[g].pos[axis] += s->diameter[axis];
        corr[g] = root(low[k]);
    }
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n1; ++k)
    {
        const int64_t g = n + n0 + k;
        atoms[g] = atoms[high[k]];
        atoms[g].pos[axis] -= s->diameter[axis];
        corr[g] = root(high[k]);
    }
    return numGhost + n0 + n1;
}

// communication/GhostExchange.cpp:171-187
int64_t or_ghost_create_xyz(or_atom_t* atoms, int64_t numLocal, int64_t capacity, const or_subdomain_t* s,
                            int64_t* corr)
{
    for (int64_t i = 0; i < capacity; ++i) corr[i] = -1;
    int64_t numGhost = 0;
    for (int axis = 0; axis < 3; ++axis)
    {
        numGhost = or_ghost_create_axis(atoms, numLocal, numGhost, capacity, s, axis, corr);
        if (numGhost < 0) return -1;
    }
    return numGhost;
}

// communication/UpdateGhostAtoms.cpp:31-68
void or_ghost_update_pos(or_atom_t* atoms, int64_t numLocal, int64_t numGhost, const int64_t* corr,
                         const or_subdomain_t* s)
{
#pragma omp parallel for schedule(static)
    for (int64_t idx = numLocal; idx < numLocal + numGhost; ++idx)
    {
        const int64_t realIdx = corr[idx];
        double dx[3];
        for (int d = 0; d < 3; ++d) dx[d] = atoms[idx].pos[d] - atoms[realIdx].pos[d];
        for (int d = 0; d < 3; ++d) atoms[idx].pos[d] = atoms[realIdx].pos[d];
        for (int d = 0; d < 3; ++d)
        {
            const double delta = 0.1 * s->diameter[d];
            if (dx[d] > +delta) atoms[idx].pos[d] += s->diameter[d];
        }
        for (int d = 0; d < 3; ++d)
        {
            const double delta = 0.1 * s->diameter[d];
            if (dx[d] < -delta) atoms[idx].pos[d] -= s->diameter[d];
        }
    }
}

// Cabana::deep_copy(force, 0) of the drivers (examples/02_LennardJones_NVE.cpp:174-175)
void or_zero_force(or_atom_t* atoms, int64_t n)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) atoms[i].force[0] = atoms[i].force[1] = atoms[i].force[2] = 0.0;
}

// communication/AccumulateForce.cpp:25-47
void or_ghost_fold_force(or_atom_t* atoms, int64_t numLocal, int64_t numGhost, const int64_t* corr)
{
#pragma omp parallel for schedule(static)
    for (int64_t idx = numLocal; idx < numLocal + numGhost; ++idx)
    {
        if (corr[idx] == -1) continue;
        const int64_t realIdx = corr[idx];
        for (int d = 0; d < 3; ++d)
        {
            atomicAdd(atoms[realIdx].force[d], atoms[idx].force[d]);
            atoms[idx].force[d] = 0.0;
        }
    }
}

// communication/MultiResRealAtomsExchange.cpp:23-73
void or_mr_periodic_map(or_molecule_t* mols, int64_t numLocalMols, or_atom_t* atoms, const or_subdomain_t* s)
{
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < numLocalMols; ++m)
    {
        for (int dim = 0; dim < 3; ++dim)
        {
            double& mx = mols[m].pos[dim];
            const int64_t a0 = mols[m].atomsOffset;
            const int64_t a1 = a0 + mols[m].numAtoms;
            if (s->maxCorner[dim] <= mx)
            {
                mx -= s->diameter[dim];
                for (int64_t a = a0; a < a1; ++a) atoms[a].pos[dim] -= s->diameter[dim];
            }
            if (mx < s->minCorner[dim])
            {
                mx += s->diameter[dim];
                for (int64_t a = a0; a < a1; ++a) atoms[a].pos[dim] += s->diameter[dim];
            }
        }
    }
}

// communication/MultiResPeriodicGhostExchange.cpp:60-248
int or_mr_ghost_create_axis(or_molecule_t* mols, int64_t numLocalMols, int64_t numGhostMols, int64_t molCapacity,
                            or_atom_t* atoms, int64_t numLocalAtoms, int64_t numGhostAtoms, int64_t atomCapacity,
                            const or_subdomain_t* s, int axis, int64_t* corrAtoms, int64_t* out)
{
    struct Sel
    {
        int64_t mol;
        int64_t atomPrefix;
    };
    std::vector<Sel> negative, positive;  // naming as in the reference: "negative" = below minInner
    int64_t negAtoms = 0, posAtoms = 0;
    const int64_t nm = numLocalMols + numGhostMols;
    for (int64_t idx = 0; idx < nm; ++idx)
    {
        if (mols[idx].pos[axis] < s->minInnerCorner[axis])
        {
            negative.push_back({idx, negAtoms});
            negAtoms += mols[idx].numAtoms;
        }
        if (mols[idx].pos[axis] >= s->maxInnerCorner[axis])
        {
            positive.push_back({idx, posAtoms});
            posAtoms += mols[idx].numAtoms;
        }
    }
    const int64_t na = numLocalAtoms + numGhostAtoms;
    const int64_t nPos = static_cast<int64_t>(positive.size());
    const int64_t nNeg = static_cast<int64_t>(negative.size());
    if (nm + nPos + nNeg > molCapacity || na + posAtoms + negAtoms > atomCapacity) return -1;
    auto root = [&](int64_t realIdx) {
        while (corrAtoms[realIdx] != -1) realIdx = corrAtoms[realIdx];
        return realIdx;
    };
    // "positive" molecules (>= maxInner) first, shifted by -L (:150-190)
    for (int64_t k = 0; k < nPos; ++k)
    {
        const int64_t src = positive[k].mol;
        const int64_t a0 = mols[src].atomsOffset;
        const int64_t size = mols[src].numAtoms;
        const int64_t gm = nm + k;
        const int64_t ga = na + positive[k].atomPrefix;
        mols[gm] = mols[src];
        mols[gm].pos[axis] -= s->diameter[axis];
        mols[gm].atomsOffset = ga;
        mols[gm].numAtoms = size;
        for (int64_t a = 0; a < size; ++a)
        {
            atoms[ga + a] = atoms[a0 + a];
            atoms[ga + a].pos[axis] -= s->diameter[axis];
            corrAtoms[ga + a] = root(a0 + a);
        }
    }
    // then the "negative" ones (< minInner), shifted by +L (:192-234)
    for (int64_t k = 0; k < nNeg; ++k)
    {
        const int64_t src = negative[k].mol;
        const int64_t a0 = mols[src].atomsOffset;
        const int64_t size = mols[src].numAtoms;
        const int64_t gm = nm + nPos + k;
        const int64_t ga = na + posAtoms + negative[k].atomPrefix;
        mols[gm] = mols[src];
        mols[gm].pos[axis] += s->diameter[axis];
        mols[gm].atomsOffset = ga;
        mols[gm].numAtoms = size;
        for (int64_t a = 0; a < size; ++a)
        {
            atoms[ga + a] = atoms[a0 + a];
            atoms[ga + a].pos[axis] += s->diameter[axis];
            corrAtoms[ga + a] = root(a0 + a);
        }
    }
    out[0] = numGhostMols + nPos + nNeg;
    out[1] = numGhostAtoms + posAtoms + negAtoms;
    return 0;
}

// communication/MultiResPeriodicGhostExchange.cpp:250-265
int or_mr_ghost_create_xyz(or_molecule_t* mols, int64_t numLocalMols, int64_t molCapacity, or_atom_t* atoms,
                           int64_t numLocalAtoms, int64_t atomCapacity, const or_subdomain_t* s, int64_t* corrAtoms,
                           int64_t* out)
{
    for (int64_t i = 0; i < atomCapacity; ++i) corrAtoms[i] = -1;
    out[0] = 0;
    out[1] = 0;
    for (int axis = 0; axis < 3; ++axis)
    {
        const int rc = or_mr_ghost_create_axis(mols, numLocalMols, out[0], molCapacity, atoms, numLocalAtoms, out[1],
                                               atomCapacity, s, axis, corrAtoms, out);
        if (rc != 0) return rc;
    }
    return 0;
}

// ---------------------------------------------------------------------------
// action/VelocityVerlet.cpp:26-67 with UpdateSteps.hpp:25-54
double or_vv_pre(or_atom_t* atoms, int64_t numLocal, double dt)
{
    const double dtHalf = 0.5 * dt;
    const double dtFull = dt;
    double maxDistSqr = 0.0;
#pragma omp parallel for schedule(static) reduction(max : maxDistSqr)
    for (int64_t idx = 0; idx < numLocal; ++idx)
    {
        or_atom_t& a = atoms[idx];
        double dx[3] = {a.pos[0], a.pos[1], a.pos[2]};
        const double dtfm = dtHalf / a.mass;
        for (int d = 0; d < 3; ++d) a.vel[d] += dtfm * a.force[d];
        for (int d = 0; d < 3; ++d) a.pos[d] += dtFull * a.vel[d];
        for (int d = 0; d < 3; ++d) dx[d] -= a.pos[d];
        const double distSqr = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
        maxDistSqr = std::max(distSqr, maxDistSqr);
    }
    return std::sqrt(maxDistSqr);
}

// action/VelocityVerlet.cpp:69-90
void or_vv_post(or_atom_t* atoms, int64_t numLocal, double dt)
{
    const double dtHalf = 0.5 * dt;
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < numLocal; ++idx)
    {
        or_atom_t& a = atoms[idx];
        const double dtfm = dtHalf / a.mass;
        for (int d = 0; d < 3; ++d) a.vel[d] += dtfm * a.force[d];
    }
}

// Philox4x32-10 (Salmon et al., SC'11).  Replaces Kokkos::Random_XorShift1024_Pool
// (VelocityVerletLangevinThermostat.hpp:32) whose stream is scheduling dependent.
void or_philox4x32(const uint32_t* ctrIn, const uint32_t* keyIn, uint32_t* out)
{
    uint32_t c[4] = {ctrIn[0], ctrIn[1], ctrIn[2], ctrIn[3]};
    uint32_t k[2] = {keyIn[0], keyIn[1]};
    for (int round = 0; round < 10; ++round)
    {
        const uint64_t p0 = uint64_t(0xD2511F53u) * c[0];
        const uint64_t p1 = uint64_t(0xCD9E8D57u) * c[2];
        const uint32_t n0 = uint32_t(p1 >> 32) ^ c[1] ^ k[0];
        const uint32_t n1 = uint32_t(p1);
        const uint32_t n2 = uint32_t(p0 >> 32) ^ c[3] ^ k[1];
        const uint32_t n3 = uint32_t(p0);
        c[0] = n0;
        c[1] = n1;
        c[2] = n2;
        c[3] = n3;
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    for (int i = 0; i < 4; ++i) out[i] = c[i];
}

void or_philox_normals(uint64_t seed, uint64_t step, uint64_t idx, double* out4)
{
    const uint32_t ctr[4] = {uint32_t(idx), uint32_t(idx >> 32), uint32_t(step), uint32_t(step >> 32)};
    const uint32_t key[2] = {uint32_t(seed), uint32_t(seed >> 32)};
    uint32_t r[4];
    or_philox4x32(ctr, key, r);
    const double scale = 2.3283064365386963e-10;  // 2^-32
    const double u0 = (double(r[0]) + 0.5) * scale;
    const double u1 = (double(r[1]) + 0.5) * scale;
    const double u2 = (double(r[2]) + 0.5) * scale;
    const double u3 = (double(r[3]) + 0.5) * scale;
    const double ra = std::sqrt(-2.0 * std::log(u0));
    const double rb = std::sqrt(-2.0 * std::log(u2));
    out4[0] = ra * std::cos(2.0 * PI * u1);
    out4[1] = ra * std::sin(2.0 * PI * u1);
    out4[2] = rb * std::cos(2.0 * PI * u3);
    out4[3] = rb * std::sin(2.0 * PI * u3);
}

// action/VelocityVerletLangevinThermostat.hpp:63-136 with UpdateSteps.hpp:56-79
double or_langevin_pre(or_atom_t* atoms, int64_t numLocal, double dt, double zeta, double temperature, uint64_t seed,
                       uint64_t step, const or_pred_t* pred)
{
    return or_langevin_pre_ids(atoms, numLocal, dt, zeta, temperature, seed, step, pred, nullptr);
}

// ids (optional): the global atom id that keys the noise instead of the index, so that a trajectory does not depend on
// the atom order (spatial sort) or on how many ranks share the system
double or_langevin_pre_ids(or_atom_t* atoms, int64_t numLocal, double dt, double zeta, double temperature, uint64_t seed,
                           uint64_t step, const or_pred_t* pred, const int64_t* ids)
{
    const double dtHalf = 0.5 * dt;
    const double dtFull = dt;
    double maxDistSqr = 0.0;
#pragma omp parallel for schedule(static) reduction(max : maxDistSqr)
    for (int64_t idx = 0; idx < numLocal; ++idx)
    {
        or_atom_t& a = atoms[idx];
        double dx[3] = {a.pos[0], a.pos[1], a.pos[2]};
        const double dtfm = dtHalf / a.mass;
        for (int d = 0; d < 3; ++d) a.vel[d] += dtfm * a.force[d];
        for (int d = 0; d < 3; ++d) a.pos[d] += dtHalf * a.vel[d];
        if (pred1(pred, a.pos[0], a.pos[1], a.pos[2]))
        {
            double rnd[4];
            or_philox_normals(seed, step, uint64_t(ids != nullptr ? ids[idx] : idx), rnd);
            const double dtm = dtFull / a.mass;
            const double damping = std::exp(-zeta * dtm);
            const double sigma = std::sqrt(temperature / a.mass * (1.0 - std::exp(-2.0 * zeta * dtm)));
            for (int d = 0; d < 3; ++d) a.vel[d] *= damping;
            for (int d = 0; d < 3; ++d) a.vel[d] += sigma * rnd[d];
        }
        for (int d = 0; d < 3; ++d) a.pos[d] += dtHalf * a.vel[d];
        for (int d = 0; d < 3; ++d) dx[d] -= a.pos[d];
        const double distSqr = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
        maxDistSqr = std::max(distSqr, maxDistSqr);
    }
    return std::sqrt(maxDistSqr);
}

// ---------------------------------------------------------------------------
// weighting_function/Slab.hpp:161-187 and Spherical.hpp:37-75 (the latter through the
// adapter modulatedLambda := lambda)
void or_weight_eval(const or_weight_t* w, double x, double y, double z, double* lambda, double* modLambda,
                    double* grad)
{
    if (w->kind == OR_WEIGHT_SLAB)
    {
        const double atHalf = 0.5 * w->atRegion;
        const int64_t exponent = 2 * w->exponent;
        const double dx = x - w->center[0];
        const double absDx = std::abs(dx);
        if (absDx < atHalf || (w->abrupt && !(absDx > atHalf + w->hyRegion)))
        {
            *lambda = 1.0;
            *modLambda = 1.0;
            grad[0] = grad[1] = grad[2] = 0.0;
        }
        else if (absDx > atHalf + w->hyRegion)
        {
            *lambda = 0.0;
            *modLambda = 0.0;
            grad[0] = grad[1] = grad[2] = 0.0;
        }
        else
        {
            const double arg = PI / (2.0 * w->hyRegion) * (absDx - atHalf);
            const double base = std::cos(arg);
            *lambda = base * base;
            *modLambda = powInt(base, exponent);
            const double factor =
                -PI / (2.0 * w->hyRegion) * double(exponent) * std::sin(arg) * powInt(base, exponent - 1) / absDx;
            grad[0] = factor * dx;
            grad[1] = 0.0;
            grad[2] = 0.0;
        }
        return;
    }
    const double atRadiusSqr = w->atRegion * w->atRegion;
    const double cgRadiusSqr = (w->atRegion + w->hyRegion) * (w->atRegion + w->hyRegion);
    const double dx[3] = {x - w->center[0], y - w->center[1], z - w->center[2]};
    const double dxSqr = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
    if (dxSqr < atRadiusSqr)
    {
        *lambda = 1.0;
        *modLambda = 1.0;
        grad[0] = grad[1] = grad[2] = 0.0;
        return;
    }
    if (dxSqr > cgRadiusSqr)
    {
        *lambda = 0.0;
        *modLambda = 0.0;
        grad[0] = grad[1] = grad[2] = 0.0;
        return;
    }
    const double r = std::sqrt(dxSqr);
    const double arg = PI / (2.0 * w->hyRegion) * (r - w->atRegion);
    const double base = std::cos(arg);
    *lambda = powInt(base, w->exponent);
    *modLambda = *lambda;
    const double factor =
        -PI / (2.0 * w->hyRegion) * double(w->exponent) * std::sin(arg) * powInt(base, w->exponent - 1) / r;
    grad[0] = factor * dx[0];
    grad[1] = factor * dx[1];
    grad[2] = factor * dx[2];
}

// action/UpdateMolecules.hpp:24-70
void or_update_molecules(or_molecule_t* mols, int64_t numAllMols, const or_atom_t* atoms, const or_weight_t* w)
{
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < numAllMols; ++m)
    {
        const int64_t a0 = mols[m].atomsOffset;
        const int64_t a1 = a0 + mols[m].numAtoms;
        mols[m].pos[0] = mols[m].pos[1] = mols[m].pos[2] = 0.0;
        for (int64_t a = a0; a < a1; ++a)
            for (int d = 0; d < 3; ++d) mols[m].pos[d] += atoms[a].pos[d] * atoms[a].relMass;
        or_weight_eval(w, mols[m].pos[0], mols[m].pos[1], mols[m].pos[2], &mols[m].lambda, &mols[m].modLambda,
                       mols[m].gradLambda);
    }
}

// action/ContributeMoleculeForceToAtoms.cpp:23-48
void or_contribute_molecule_force(const or_molecule_t* mols, int64_t numAllMols, or_atom_t* atoms)
{
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < numAllMols; ++m)
    {
        const int64_t a0 = mols[m].atomsOffset;
        const int64_t a1 = a0 + mols[m].numAtoms;
        for (int64_t a = a0; a < a1; ++a)
            for (int d = 0; d < 3; ++d) atoms[a].force[d] += atoms[a].relMass * mols[m].force[d];
    }
}

// action/LJ_IdealGas.cpp:262-292
or_adress_t* or_adress_create(const double* cappingDistance, const double* rc, const double* sigma,
                              const double* epsilon, int64_t numTypes, int doShift)
{
    auto* a = new or_adress_t;
    a->numTypes = numTypes;
    a->numBins = 200;
    a->runCounter = 0;
    a->samplingInterval = 200;
    a->updateInterval = 20000;
    a->table = new or_lj_type_t[numTypes * numTypes];
    or_lj_init(a->table, cappingDistance, rc, sigma, epsilon, numTypes, doShift);
    const double maxRC = *std::max_element(rc, rc + numTypes * numTypes);
    a->rcSqr = maxRC * maxRC;
    a->compensationEnergy = new double[a->numBins * numTypes]();
    a->compensationEnergyCounter = new double[a->numBins * numTypes]();
    a->meanCompensationEnergy = new double[a->numBins * numTypes]();
    return a;
}

void or_adress_destroy(or_adress_t* a)
{
    delete[] a->table;
    delete[] a->compensationEnergy;
    delete[] a->compensationEnergyCounter;
    delete[] a->meanCompensationEnergy;
    delete a;
}

// action/LJ_IdealGas.cpp:52-260.  The reference's non-atomic counter update (:207) is a
// race under OpenMP; the oracle accumulates it atomically (== the serial result).
double or_adress_run(or_adress_t* A, or_molecule_t* mols, int64_t numLocalMols, const int32_t* counts,
                     const int32_t* neigh, int64_t width, or_atom_t* atoms, int64_t* numActive)
{
    const bool sampling = (A->runCounter % A->samplingInterval) == 0;
    const int64_t T = A->numTypes;
    const double binMin = 0.0, binMax = 1.0;
    const double binSize = (binMax - binMin) / double(A->numBins);
    const double inverseBinSize = 1.0 / binSize;
    double energy = 0.0;
    int64_t active = 0;
#pragma omp parallel for schedule(dynamic, 128) reduction(+ : energy, active)
    for (int64_t alpha = 0; alpha < numLocalMols; ++alpha)
    {
        double forceTmpAlpha[3] = {0.0, 0.0, 0.0};
        const double modLambdaAlpha = mols[alpha].modLambda;
        int64_t binAlpha = -1;
        if (inHY(modLambdaAlpha)) binAlpha = histBin(binMin, inverseBinSize, A->numBins, mols[alpha].lambda);
        const double gradAlpha[3] = {mols[alpha].gradLambda[0], mols[alpha].gradLambda[1], mols[alpha].gradLambda[2]};
        const int64_t startAlpha = mols[alpha].atomsOffset;
        const int64_t endAlpha = startAlpha + mols[alpha].numAtoms;
        const int64_t numNeighbors = counts[alpha];
        for (int64_t n = 0; n < numNeighbors; ++n)
        {
            const int64_t beta = neigh[alpha * width + n];
            double forceTmpBeta[3] = {0.0, 0.0, 0.0};
            const double modLambdaBeta = mols[beta].modLambda;
            const double gradBeta[3] = {mols[beta].gradLambda[0], mols[beta].gradLambda[1], mols[beta].gradLambda[2]};
            const double weighting = 0.5 * (modLambdaAlpha + modLambdaBeta);
            if (inCG(modLambdaAlpha) && inCG(modLambdaBeta)) continue;
            const int64_t startBeta = mols[beta].atomsOffset;
            const int64_t endBeta = startBeta + mols[beta].numAtoms;
            for (int64_t idx = startAlpha; idx < endAlpha; ++idx)
            {
                const double posTmp[3] = {atoms[idx].pos[0], atoms[idx].pos[1], atoms[idx].pos[2]};
                double forceTmpIdx[3] = {0.0, 0.0, 0.0};
                for (int64_t jdx = startBeta; jdx < endBeta; ++jdx)
                {
                    const double dx = posTmp[0] - atoms[jdx].pos[0];
                    const double dy = posTmp[1] - atoms[jdx].pos[1];
                    const double dz = posTmp[2] - atoms[jdx].pos[2];
                    const double distSqr = dx * dx + dy * dy + dz * dz;
                    if (distSqr > A->rcSqr) continue;
                    const int64_t typeIdx = atoms[idx].type * T + atoms[jdx].type;
                    double ff, e;
                    ljForceEnergy(A->table[typeIdx], distSqr, ff, e);
                    const double ffactor = ff * weighting;
                    active += 1;
                    forceTmpIdx[0] += dx * ffactor;
                    forceTmpIdx[1] += dy * ffactor;
                    forceTmpIdx[2] += dz * ffactor;
                    atomicAdd(atoms[jdx].force[0], -(dx * ffactor));
                    atomicAdd(atoms[jdx].force[1], -(dy * ffactor));
                    atomicAdd(atoms[jdx].force[2], -(dz * ffactor));
                    energy += e * weighting;
                    const double Vij = 0.5 * e;
                    if (inHY(modLambdaAlpha) || inHY(modLambdaBeta))
                    {
                        for (int d = 0; d < 3; ++d) forceTmpAlpha[d] += -Vij * gradAlpha[d];
                        for (int d = 0; d < 3; ++d) forceTmpBeta[d] += -Vij * gradBeta[d];
                        if (sampling)
                        {
                            int64_t binBeta = -1;
                            if (inHY(modLambdaBeta))
                                binBeta = histBin(binMin, inverseBinSize, A->numBins, mols[beta].lambda);
                            if (inHY(modLambdaAlpha) && (binAlpha != -1))
                                atomicAdd(A->compensationEnergy[binAlpha * T + atoms[idx].type], Vij);
                            if (inHY(modLambdaBeta) && (binBeta != -1))
                                atomicAdd(A->compensationEnergy[binBeta * T + atoms[jdx].type], Vij);
                        }
                    }
                }
                for (int d = 0; d < 3; ++d) atomicAdd(atoms[idx].force[d], forceTmpIdx[d]);
            }
            for (int d = 0; d < 3; ++d) atomicAdd(mols[beta].force[d], forceTmpBeta[d]);
        }
        if (sampling)
        {
            for (int64_t atomIdx = startAlpha; atomIdx < endAlpha; ++atomIdx)
                if (inHY(modLambdaAlpha) && (binAlpha != -1))
                    atomicAdd(A->compensationEnergyCounter[binAlpha * T + atoms[atomIdx].type], 1.0);
        }
        if (inHY(modLambdaAlpha) && (binAlpha != -1))
        {
            const double mean = A->meanCompensationEnergy[binAlpha * T + atoms[startAlpha].type];
            for (int d = 0; d < 3; ++d) forceTmpAlpha[d] += mean * gradAlpha[d];
        }
        for (int d = 0; d < 3; ++d) atomicAdd(mols[alpha].force[d], forceTmpAlpha[d]);
    }
    // updateMeanCompensationEnergy, LJ_IdealGas.cpp:21-50 (runningAverageFactor = 10)
    if (A->runCounter % A->updateInterval == 0)
    {
        const double f = 10.0;
        for (int64_t i = 0; i < A->numBins * T; ++i)
        {
            if (A->compensationEnergyCounter[i] < 0.5) continue;
            const double e = A->compensationEnergy[i] / A->compensationEnergyCounter[i];
            A->meanCompensationEnergy[i] = (f * A->meanCompensationEnergy[i] + e) / (f + 1.0);
            A->compensationEnergy[i] = 0.0;
            A->compensationEnergyCounter[i] = 0.0;
        }
    }
    A->runCounter += 1;
    if (numActive != nullptr) *numActive = active;
    return energy;
}

// ---------------------------------------------------------------------------
int64_t or_hist_get_bin(double min, double max, int64_t numBins, double val)
{
    const double binSize = (max - min) / double(numBins);
    return histBin(min, 1.0 / binSize, numBins, val);
}

// data/MultiHistogram.cpp:48-59
void or_hist_scale(double* data, int64_t numBins, int64_t numHist, double factor)
{
    for (int64_t i = 0; i < numBins * numHist; ++i) data[i] *= factor;
}
// data/MultiHistogram.cpp:61-74
void or_hist_scale_per_hist(double* data, int64_t numBins, int64_t numHist, const double* factors)
{
    for (int64_t i = 0; i < numBins; ++i)
        for (int64_t j = 0; j < numHist; ++j) data[i * numHist + j] *= factors[j];
}
// data/MultiHistogram.cpp:76-90
void or_hist_make_symmetric(double* data, int64_t numBins, int64_t numHist)
{
    const int64_t maxIdx = numBins - 1;
    for (int64_t i = 0; i < numBins / 2; ++i)
        for (int64_t j = 0; j < numHist; ++j)
        {
            const double val = 0.5 * (data[i * numHist + j] + data[(maxIdx - i) * numHist + j]);
            data[i * numHist + j] = val;
            data[(maxIdx - i) * numHist + j] = val;
        }
}
// data/MultiHistogram.cpp:113-161
void or_hist_gradient(const double* in, double* out, double min, double max, int64_t numBins, int64_t numHist,
                      int periodic)
{
    const double binSize = (max - min) / double(numBins);
    const double inverseSpacing = 1.0 / binSize;
    const double inverseDoubleSpacing = 0.5 * inverseSpacing;
    auto at = [&](int64_t i, int64_t j) { return in[i * numHist + j]; };
    for (int64_t i = 0; i < numBins; ++i)
        for (int64_t j = 0; j < numHist; ++j)
        {
            double g;
            if (i == 0)
                g = periodic ? (at(i + 1, j) - at(numBins - 1, j)) * inverseDoubleSpacing
                             : (at(i + 1, j) - at(i, j)) * inverseSpacing;
            else if (i == numBins - 1)
                g = periodic ? (at(0, j) - at(i - 1, j)) * inverseDoubleSpacing
                             : (at(i, j) - at(i - 1, j)) * inverseSpacing;
            else
                g = (at(i + 1, j) - at(i - 1, j)) * inverseDoubleSpacing;
            out[i * numHist + j] = g;
        }
}
// data/MultiHistogram.cpp:163-212
void or_hist_smoothen(const double* in, double* out, double min, double max, int64_t numBins, int64_t numHist,
                      double sigma, double range, int periodic)
{
    const double binSize = (max - min) / double(numBins);
    const double inverseBinSize = 1.0 / binSize;
    const double inverseSigma = 1.0 / sigma;
    const int64_t delta = static_cast<int>(range * sigma * inverseBinSize);
    for (int64_t b = 0; b < numBins; ++b)
        for (int64_t h = 0; h < numHist; ++h)
        {
            double normalization = 0.0;
            double acc = 0.0;
            int64_t jMin = b - delta;
            int64_t jMax = b + delta;
            if (!periodic)
            {
                jMin = std::max<int64_t>(0, jMin);
                jMax = std::min<int64_t>(numBins - 1, jMax);
            }
            for (int64_t j = jMin; j <= jMax; ++j)
            {
                int64_t mapped = j;
                if (periodic)
                {
                    if (mapped < 0) mapped += numBins;
                    if (mapped >= numBins) mapped -= numBins;
                }
                const double t = double(b - j) * binSize * inverseSigma;
                const double eFunc = std::exp(-(t * t));
                normalization += eFunc;
                acc += in[mapped * numHist + h] * eFunc;
            }
            out[b * numHist + h] = acc / normalization;
        }
}

// action/BerendsenThermostat.cpp:25-50
void or_berendsen_thermostat(or_atom_t* atoms, int64_t numLocal, double currentTemperature, double targetTemperature,
                             double gamma)
{
    if (currentTemperature <= 0.0) return;
    const double beta = std::sqrt(1.0 + gamma * (targetTemperature / currentTemperature - 1.0));
    for (int64_t idx = 0; idx < numLocal; ++idx)
        for (int d = 0; d < 3; ++d) atoms[idx].vel[d] *= beta;
}

// action/LimitAcceleration.cpp:21-45, action/LimitVelocity.cpp:23-43
void or_limit_acceleration(or_atom_t* atoms, int64_t numLocal, double maxAcc)
{
    for (int64_t i = 0; i < numLocal; ++i)
    {
        const double m = atoms[i].mass, invM = 1.0 / m;
        for (int d = 0; d < 3; ++d)
        {
            atoms[i].force[d] = std::min(atoms[i].force[d] * invM, +maxAcc) * m;
            atoms[i].force[d] = std::max(atoms[i].force[d] * invM, -maxAcc) * m;
        }
    }
}

void or_limit_velocity(or_atom_t* atoms, int64_t numLocal, double maxVel)
{
    for (int64_t i = 0; i < numLocal; ++i)
        for (int d = 0; d < 3; ++d)
        {
            atoms[i].vel[d] = std::min(atoms[i].vel[d], +maxVel);
            atoms[i].vel[d] = std::max(atoms[i].vel[d], -maxVel);
        }
}

// action/BerendsenBarostat.cpp:23-50
void or_berendsen_barostat(or_atom_t* atoms, int64_t numLocal, double currentPressure, double targetPressure, double gamma,
                           or_subdomain_t* s, int stretchX, int stretchY, int stretchZ)
{
    const double mu = std::cbrt(1.0 + gamma * (currentPressure - targetPressure));
    const int stretch[3] = {stretchX, stretchY, stretchZ};
    for (int d = 0; d < 3; ++d)
        if (stretch[d]) or_subdomain_scale_dim(s, mu, d);
    for (int64_t idx = 0; idx < numLocal; ++idx)
        for (int d = 0; d < 3; ++d)
            if (stretch[d]) atoms[idx].pos[d] *= mu;
}

// action/Shake.hpp:167-201 (MoleculeConstraints::enforcePositionalConstraints) with impl::Shake :84-137;
// bonds: numBonds x {idx, jdx} relative to the molecule's first atom, eqDistance[numBonds]
int or_shake_positional(const or_molecule_t* mols, int64_t numLocalMols, or_atom_t* atoms, int64_t numAllAtoms,
                        const int64_t* bondIdx, const double* eqDistance, int64_t numBonds, int64_t numIterations,
                        double dt)
{
    const double dtv = dt, dtf = 0.5 * dt * dt;
    std::vector<double> updated(static_cast<size_t>(3 * numAllAtoms));
    for (int64_t it = 0; it < numIterations; ++it)
    {
        for (int64_t idx = 0; idx < numAllAtoms; ++idx)
        {
            const double dtfm = dtf / atoms[idx].mass;
            for (int d = 0; d < 3; ++d)
                updated[3 * idx + d] = atoms[idx].pos[d] + dtv * atoms[idx].vel[d] + dtfm * atoms[idx].force[d];
        }
        for (int64_t mol = 0; mol < numLocalMols; ++mol)
        {
            const int64_t start = mols[mol].atomsOffset, count = mols[mol].numAtoms;
            for (int64_t b = 0; b < numBonds; ++b)
            {
                if (bondIdx[2 * b] >= count || bondIdx[2 * b + 1] >= count) return -1;
                const int64_t idx = start + bondIdx[2 * b], jdx = start + bondIdx[2 * b + 1];
                double dist[3], upd[3];
                for (int d = 0; d < 3; ++d)
                {
                    dist[d] = atoms[idx].pos[d] - atoms[jdx].pos[d];
                    upd[d] = updated[3 * idx + d] - updated[3 * jdx + d];
                }
                const double distSq = dist[0] * dist[0] + dist[1] * dist[1] + dist[2] * dist[2];
                const double updSq = upd[0] * upd[0] + upd[1] * upd[1] + upd[2] * upd[2];
                const double invMassI = 1.0 / atoms[idx].mass, invMassJ = 1.0 / atoms[jdx].mass;
                const double a = (invMassI + invMassJ) * (invMassI + invMassJ) * distSq;
                const double bq = 2.0 * (invMassI + invMassJ) * (upd[0] * dist[0] + upd[1] * dist[1] + upd[2] * dist[2]);
                const double c = updSq - eqDistance[b] * eqDistance[b];
                double determinant = bq * bq - 4.0 * a * c;
                determinant = std::max(0.0, determinant);
                const double lambda1 = (-bq + std::sqrt(determinant)) / (2.0 * a);
                const double lambda2 = (-bq - std::sqrt(determinant)) / (2.0 * a);
                double lambda = std::abs(lambda1) < std::abs(lambda2) ? lambda1 : lambda2;
                lambda /= dtf;
                for (int d = 0; d < 3; ++d)
                {
                    atoms[idx].force[d] += lambda * dist[d];
                    atoms[jdx].force[d] -= lambda * dist[d];
                }
            }
        }
    }
    return 0;
}

// action/Shake.hpp:203-233 (enforceVelocityConstraints) with impl::Shake :56-82
int or_shake_velocity(const or_molecule_t* mols, int64_t numLocalMols, or_atom_t* atoms, const int64_t* bondIdx,
                      int64_t numBonds)
{
    for (int64_t mol = 0; mol < numLocalMols; ++mol)
    {
        const int64_t start = mols[mol].atomsOffset, count = mols[mol].numAtoms;
        for (int64_t b = 0; b < numBonds; ++b)
        {
            if (bondIdx[2 * b] >= count || bondIdx[2 * b + 1] >= count) return -1;
            const int64_t idx = start + bondIdx[2 * b], jdx = start + bondIdx[2 * b + 1];
            double dist[3], relVel[3];
            for (int d = 0; d < 3; ++d)
            {
                dist[d] = atoms[idx].pos[d] - atoms[jdx].pos[d];
                relVel[d] = atoms[idx].vel[d] - atoms[jdx].vel[d];
            }
            const double distSq = dist[0] * dist[0] + dist[1] * dist[1] + dist[2] * dist[2];
            const double invMassI = 1.0 / atoms[idx].mass, invMassJ = 1.0 / atoms[jdx].mass;
            const double reducedMass = 1.0 / (invMassI + invMassJ);
            const double factor = (relVel[0] * dist[0] + relVel[1] * dist[1] + relVel[2] * dist[2]) / distSq * reducedMass;
            for (int d = 0; d < 3; ++d)
            {
                atoms[idx].vel[d] -= factor * dist[d] * invMassI;
                atoms[jdx].vel[d] += factor * dist[d] * invMassJ;
            }
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------
// util/math.hpp:57-76 (Abramowitz & Stegun 7.1.27; returns exp(-x^2) through expX2)
static inline double approxErfc(double x, double& expX2)
{
    constexpr double p = 0.3275911;
    constexpr double a1 = 0.254829592;
    constexpr double a2 = -0.284496736;
    constexpr double a3 = 1.421413741;
    constexpr double a4 = -1.453152027;
    constexpr double a5 = 1.061405429;
    const double t = 1.0 / (1.0 + p * x);
    expX2 = std::exp(-x * x);
    return t * (a1 + t * (a2 + t * (a3 + t * (a4 + t * a5)))) * expX2;
}

double or_approx_erfc(double x)
{
    double tmp;
    return approxErfc(x, tmp);
}

// action/Coulomb.hpp:27-46 (kind 0) and action/CoulombDSF.hpp:42-84 (kind 1)
struct CoulombParams
{
    int kind;
    double alpha, rc, forceShift, energyShift;
};
static CoulombParams coulombInit(int kind, double rc, double alpha)
{
    CoulombParams c{kind, alpha, rc, 0.0, 0.0};
    if (kind == 1)
    {
        const double rcSqr = rc * rc;
        const double erfc = std::erfc(alpha * rc);
        const double ex = std::exp(-alpha * alpha * rcSqr);
        c.forceShift = -(erfc / rcSqr + 2.0 * std::numbers::inv_sqrtpi_v<double> * alpha * ex / rc);
        c.energyShift = erfc / rc;
    }
    return c;
}
static inline double coulombForce(const CoulombParams& c, double distSqr, double q1, double q2)
{
    const double prefac = 138.935458 * q1 * q2;
    if (c.kind == 0) return prefac / distSqr;
    const double r = std::sqrt(distSqr);
    double expX2;
    const double erfc = approxErfc(c.alpha * r, expX2);
    const double force = prefac * (erfc / r + 2.0 * c.alpha * std::numbers::inv_sqrtpi_v<double> * expX2 + r * c.forceShift);
    return force / distSqr;
}
static inline double coulombEnergy(const CoulombParams& c, double distSqr, double q1, double q2)
{
    const double r = std::sqrt(distSqr);
    const double prefac = 138.935458 * q1 * q2;
    if (c.kind == 0) return prefac / r;
    double tmp;
    const double erfc = approxErfc(c.alpha * r, tmp);
    return prefac * (erfc / r - c.energyShift - c.forceShift * (r - c.rc));
}

void or_coulomb_eval(int kind, double rc, double alpha, const double* distSqr, int64_t n, double q1, double q2,
                     double* force, double* energy)
{
    const CoulombParams c = coulombInit(kind, rc, alpha);
    for (int64_t i = 0; i < n; ++i)
    {
        force[i] = coulombForce(c, distSqr[i], q1, q2);
        energy[i] = coulombEnergy(c, distSqr[i], q1, q2);
    }
}

// action/SPC.hpp:143-236 (operator()(CalcInteractions) + applyForces): capped, shifted O-O Lennard-Jones between the
// first atoms of two molecules (strict <) and Coulomb between all atom pairs (skipped only for >); energies[0] = LJ,
// energies[1] = Coulomb.  coulombKind 0 is the reference's impl::Coulomb member; 1 swaps in CoulombDSF(rc, alpha).
void or_spc_apply_forces(const or_molecule_t* mols, int64_t numLocalMols, const int32_t* counts, const int32_t* neigh,
                         int64_t width, or_atom_t* atoms, int coulombKind, double* energies)
{
    constexpr double sigma = 0.31655578901998815, epsilon = 0.6501695808187486, rc = 1.2, alpha = 2.0;  // :105-110
    const double cap = 0.7 * sigma;
    or_lj_type_t lj;
    or_lj_init(&lj, &cap, &rc, &sigma, &epsilon, 1, 1);  // SPC(): LJ_({0.7 sigma}, {rc}, {sigma}, {epsilon}, 1, true)
    const CoulombParams cp = coulombInit(coulombKind, rc, alpha);
    const double rcSqr = rc * rc;
    double eLJ = 0.0, eC = 0.0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : eLJ, eC)
    for (int64_t a = 0; a < numLocalMols; ++a)
    {
        const int64_t numNeighbors = counts[a];
        for (int64_t n = 0; n < numNeighbors; ++n)
        {
            const int64_t b = neigh[a * width + n];
            const int64_t startA = mols[a].atomsOffset, endA = startA + mols[a].numAtoms;
            const int64_t startB = mols[b].atomsOffset, endB = startB + mols[b].numAtoms;
            {
                const double dx[3] = {atoms[startA].pos[0] - atoms[startB].pos[0], atoms[startA].pos[1] - atoms[startB].pos[1],
                                      atoms[startA].pos[2] - atoms[startB].pos[2]};
                const double distSqr = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
                if (distSqr < rcSqr)
                {
                    double ff, e;
                    ljForceEnergy(lj, distSqr, ff, e);
                    eLJ += e;
                    for (int d = 0; d < 3; ++d) atomicAdd(atoms[startB].force[d], -(dx[d] * ff));
                    for (int d = 0; d < 3; ++d) atomicAdd(atoms[startA].force[d], dx[d] * ff);
                }
            }
            for (int64_t idx = startA; idx < endA; ++idx)
            {
                const double q1 = atoms[idx].charge;
                double forceTmpIdx[3] = {0.0, 0.0, 0.0};
                for (int64_t jdx = startB; jdx < endB; ++jdx)
                {
                    const double q2 = atoms[jdx].charge;
                    const double dx[3] = {atoms[idx].pos[0] - atoms[jdx].pos[0], atoms[idx].pos[1] - atoms[jdx].pos[1],
                                          atoms[idx].pos[2] - atoms[jdx].pos[2]};
                    const double distSqr = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
                    if (distSqr > rcSqr) continue;
                    const double ffactor = coulombForce(cp, distSqr, q1, q2);
                    eC += coulombEnergy(cp, distSqr, q1, q2);
                    for (int d = 0; d < 3; ++d) forceTmpIdx[d] += dx[d] * ffactor;
                    for (int d = 0; d < 3; ++d) atomicAdd(atoms[jdx].force[d], -(dx[d] * ffactor));
                }
                for (int d = 0; d < 3; ++d) atomicAdd(atoms[idx].force[d], forceTmpIdx[d]);
            }
        }
    }
    energies[0] = eLJ;
    energies[1] = eC;
}

// action/SPC.hpp:284-344 (BondEnergy + calcBondEnergy): O, H0, H1 are the first three atoms of every molecule
double or_spc_bond_energy(const or_molecule_t* mols, int64_t numAllMols, const or_atom_t* atoms, int64_t numAllAtoms,
                          double harmonicPreFactor)
{
    constexpr double eqDistanceHO = 0.1;
    const double angleHOH = 109.47 / 180.0 * M_PI;  // util/angle.hpp:28
    const double eqDistanceHH = eqDistanceHO * std::sqrt(2.0 - 2.0 * std::cos(angleHOH));
    double energy = 0.0;
    auto dist = [&](int64_t i, int64_t j)
    {
        const double dx[3] = {atoms[i].pos[0] - atoms[j].pos[0], atoms[i].pos[1] - atoms[j].pos[1], atoms[i].pos[2] - atoms[j].pos[2]};
        return std::sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]);
    };
    for (int64_t m = 0; m < numAllMols; ++m)
    {
        const int64_t o = mols[m].atomsOffset, h0 = o + 1, h1 = o + 2;
        double d = dist(o, h0) - eqDistanceHO;
        energy += d * d;
        d = dist(o, h1) - eqDistanceHO;
        energy += d * d;
        d = dist(h0, h1) - eqDistanceHH;
        energy += d * d;
    }
    return harmonicPreFactor * energy / double(numAllAtoms);
}

// analysis/KineticEnergy.hpp:26-39
double or_kinetic_energy(const or_atom_t* atoms, int64_t numLocal)
{
    double velSqr = 0.0;
    for (int64_t idx = 0; idx < numLocal; ++idx)
    {
        const double* v = atoms[idx].vel;
        velSqr += atoms[idx].mass * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    }
    return 0.5 * velSqr;
}

// analysis/SystemMomentum.cpp:21-50 (sums velocities, one reduction per component)
void or_system_momentum(const or_atom_t* atoms, int64_t numLocal, double* out3)
{
    for (int dim = 0; dim < 3; ++dim)
    {
        double sum = 0.0;
        for (int64_t idx = 0; idx < numLocal; ++idx) sum += atoms[idx].vel[dim];
        out3[dim] = sum;
    }
}

// analysis/Pressure.cpp:23-51 (local and ghost atoms)
double or_pressure(const or_atom_t* atoms, int64_t numAll, const or_subdomain_t* s)
{
    double pressure = 0.0;
    for (int64_t idx = 0; idx < numAll; ++idx)
    {
        const or_atom_t& a = atoms[idx];
        pressure += a.mass * (a.vel[0] * a.vel[0] + a.vel[1] * a.vel[1] + a.vel[2] * a.vel[2]);
        pressure += a.force[0] * a.pos[0] + a.force[1] * a.pos[1] + a.force[2] * a.pos[2];
    }
    const double volume = s->diameter[0] * s->diameter[1] * s->diameter[2];
    return pressure / (3.0 * volume);
}

// analysis/MeanSquareDisplacement.cpp:59-84; initialPos: numItems x 3 doubles saved by reset (:23-40)
double or_msd(const or_atom_t* atoms, const double* initialPos, int64_t numItems, const or_subdomain_t* s)
{
    double sq = 0.0;
    for (int64_t idx = 0; idx < numItems; ++idx)
    {
        double dx[3];
        for (int d = 0; d < 3; ++d)
        {
            dx[d] = std::abs(initialPos[3 * idx + d] - atoms[idx].pos[d]);
            if (dx[d] > 0.5 * s->diameter[d]) dx[d] -= s->diameter[d];
        }
        sq += dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
    }
    return sq / double(numItems);
}

// analysis/AxialDensityProfile.cpp:21-51
void or_density_profile(const or_atom_t* atoms, int64_t numAtoms, int64_t numTypes, double min, double max,
                        int64_t numBins, int axis, double* hist)
{
    const double binSize = (max - min) / double(numBins);
    const double inverseBinSize = 1.0 / binSize;
    for (int64_t i = 0; i < numBins * numTypes; ++i) hist[i] = 0.0;
    for (int64_t idx = 0; idx < numAtoms; ++idx)
    {
        const int64_t bin = histBin(min, inverseBinSize, numBins, atoms[idx].pos[axis]);
        if (bin == -1) continue;
        hist[bin * numTypes + atoms[idx].type] += 1.0;
    }
}

// action/ThermodynamicForce.cpp:25-57
or_thermo_t* or_thermo_create(const double* targetDensity, int64_t numTypes, const or_subdomain_t* s,
                              double requestedBinWidth, const double* modulation, int enforceSymmetry,
                              int usePeriodicity)
{
    auto* t = new or_thermo_t;
    t->min = s->minCorner[0];
    t->max = s->maxCorner[0];
    t->numBins = static_cast<int64_t>(std::ceil(s->diameter[0] / requestedBinWidth));
    t->numTypes = numTypes;
    t->binSize = (t->max - t->min) / double(t->numBins);
    t->inverseBinSize = 1.0 / t->binSize;
    t->binVolume = s->diameter[1] * s->diameter[2] * t->binSize;
    t->samples = 0;
    t->enforceSymmetry = enforceSymmetry;
    t->usePeriodicity = usePeriodicity;
    t->force = new double[t->numBins * numTypes]();
    t->density = new double[t->numBins * numTypes]();
    t->forceFactor = new double[numTypes];
    for (int64_t i = 0; i < numTypes; ++i) t->forceFactor[i] = modulation[i] / targetDensity[i];
    return t;
}

void or_thermo_destroy(or_thermo_t* t)
{
    delete[] t->force;
    delete[] t->density;
    delete[] t->forceFactor;
    delete t;
}

// action/ThermodynamicForce.cpp:74-86
void or_thermo_sample(or_thermo_t* t, const or_atom_t* atoms, int64_t numLocal)
{
    std::vector<double> h(static_cast<size_t>(t->numBins * t->numTypes));
    or_density_profile(atoms, numLocal, t->numTypes, t->min, t->max, t->numBins, 0, h.data());
    for (int64_t i = 0; i < t->numBins * t->numTypes; ++i) t->density[i] += h[i];
    t->samples += 1;
}

// action/ThermodynamicForce.hpp:124-152
void or_thermo_update(or_thermo_t* t, double smoothingSigma, double smoothingIntensity, const or_pred_t* pred)
{
    const int64_t nb = t->numBins, nt = t->numTypes;
    if (t->enforceSymmetry) or_hist_make_symmetric(t->density, nb, nt);
    const double normalizationFactor = 1.0 / (t->binVolume * double(t->samples));
    or_hist_scale(t->density, nb, nt, normalizationFactor);
    std::vector<double> smooth(static_cast<size_t>(nb * nt)), grad(static_cast<size_t>(nb * nt));
    or_hist_smoothen(t->density, smooth.data(), t->min, t->max, nb, nt, smoothingSigma, smoothingIntensity,
                     t->usePeriodicity);
    or_hist_gradient(smooth.data(), grad.data(), t->min, t->max, nb, nt, t->usePeriodicity);
    or_hist_scale_per_hist(grad.data(), nb, nt, t->forceFactor);
    for (int64_t b = 0; b < nb; ++b)
    {
        const double x = t->min + (double(b) + 0.5) * t->binSize;  // getBinPosition
        if (!pred1(pred, x, x, x))
            for (int64_t h = 0; h < nt; ++h) grad[b * nt + h] = 0.0;
    }
    for (int64_t i = 0; i < nb * nt; ++i) t->force[i] -= grad[i];
    for (int64_t i = 0; i < nb * nt; ++i) t->density[i] = 0.0;
    t->samples = 0;
}

// action/ThermodynamicForce.hpp:98-122 (apply_if) and :154-216 (applyInterpolated_if)
void or_thermo_apply(const or_thermo_t* t, or_atom_t* atoms, int64_t numLocal, const or_pred_t* pred,
                     int interpolated)
{
    const int64_t nt = t->numTypes;
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < numLocal; ++idx)
    {
        const double xPos = atoms[idx].pos[0];
        if (!pred1(pred, atoms[idx].pos[0], atoms[idx].pos[1], atoms[idx].pos[2])) continue;
        const int64_t bin = histBin(t->min, t->inverseBinSize, t->numBins, xPos);
        if (bin == -1) continue;
        const int64_t type = atoms[idx].type;
        if (!interpolated)
        {
            atoms[idx].force[0] += t->force[bin * nt + type];
            continue;
        }
        const double binStart = t->min + double(bin) * t->binSize;
        const double fracInBin = (xPos - binStart) * t->inverseBinSize;
        int64_t left, right;
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
        if (left >= 0 && right < t->numBins)
        {
            const double l = t->force[left * nt + type];
            const double r = t->force[right * nt + type];
            atoms[idx].force[0] += l + (r - l) * factor;  // util::lerp, util/interpolation.hpp:34-38
        }
        else
        {
            atoms[idx].force[0] += t->force[bin * nt + type];
        }
    }
}

// action/ThermodynamicForce.cpp:98-130
void or_thermo_mu(const or_thermo_t* t, double* muLeft, double* muRight)
{
    const int64_t nb = t->numBins, nt = t->numTypes;
    for (int64_t ty = 0; ty < nt; ++ty)
    {
        muLeft[ty] = 0.0;
        muRight[ty] = 0.0;
        for (int64_t i = 0; i < nb / 2; ++i) muLeft[ty] += t->force[i * nt + ty];
        muLeft[ty] *= t->binSize;
        for (int64_t i = nb / 2; i < nb; ++i) muRight[ty] += t->force[i * nt + ty];
        muRight[ty] *= t->binSize;
    }
}

}  // extern "C"
