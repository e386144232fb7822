// containers.cu -- device containers behind data::Atoms / data::Molecules and the slice transfer kernels.
// Reference: mrmd/data/Atoms.hpp:33-177, mrmd/data/Molecules.hpp:27-178, mrmd/data/MoleculesFromAtoms.cpp:19-39.
#include <mutex>

#include "common.cuh"

namespace mrmd_b200
{
std::atomic<int64_t> g_launchCount{0};
static std::mutex g_errMutex;
static std::string g_lastError;

void setLastError(const std::string& msg)
{
    std::lock_guard<std::mutex> lock(g_errMutex);
    g_lastError = msg;
}

int checkDevice()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
    {
        setLastError("no CUDA device available: the mrmd_b200 hot path has no CPU fallback");
        return MRMD_B200_ENODEVICE;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// slab carving: one allocation per view, planes 256-byte aligned
static size_t alignUp(size_t v) { return (v + 255) & ~size_t(255); }

static size_t atomsSlabBytes(int64_t cap) { return alignUp(size_t(cap) * 32) + 10 * alignUp(size_t(cap) * 8); }
static void carveAtoms(void* slab, int64_t cap, AtomsView& v)
{
    char* p = static_cast<char*>(slab);
    v.pos = reinterpret_cast<double4*>(p);
    p += alignUp(size_t(cap) * 32);
    double** planes[9] = {&v.vel[0], &v.vel[1], &v.vel[2], &v.force[0], &v.force[1], &v.force[2], &v.mass, &v.charge,
                          &v.relMass};
    for (auto* pl : planes)
    {
        *pl = reinterpret_cast<double*>(p);
        p += alignUp(size_t(cap) * 8);
    }
    v.gid = reinterpret_cast<long long*>(p);
}
static size_t molsSlabBytes(int64_t cap)
{
    return 2 * alignUp(size_t(cap) * 32) + alignUp(size_t(cap) * 16) + 4 * alignUp(size_t(cap) * 8);
}
static void carveMols(void* slab, int64_t cap, MolsView& v)
{
    char* p = static_cast<char*>(slab);
    v.pos = reinterpret_cast<double4*>(p);
    p += alignUp(size_t(cap) * 32);
    v.w = reinterpret_cast<double4*>(p);
    p += alignUp(size_t(cap) * 32);
    v.oc = reinterpret_cast<longlong2*>(p);
    p += alignUp(size_t(cap) * 16);
    double** planes[4] = {&v.lambda, &v.force[0], &v.force[1], &v.force[2]};
    for (auto* pl : planes)
    {
        *pl = reinterpret_cast<double*>(p);
        p += alignUp(size_t(cap) * 8);
    }
}

static int copyAtomsView(const AtomsView& dst, const AtomsView& src, int64_t n, cudaStream_t st)
{
    if (n <= 0) return 0;
    MB_CUDA(cudaMemcpyAsync(dst.pos, src.pos, size_t(n) * 32, cudaMemcpyDeviceToDevice, st));
    const double* s[9] = {src.vel[0], src.vel[1], src.vel[2], src.force[0], src.force[1], src.force[2], src.mass,
                          src.charge, src.relMass};
    double* d[9] = {dst.vel[0], dst.vel[1], dst.vel[2], dst.force[0], dst.force[1], dst.force[2], dst.mass,
                    dst.charge, dst.relMass};
    for (int i = 0; i < 9; ++i) MB_CUDA(cudaMemcpyAsync(d[i], s[i], size_t(n) * 8, cudaMemcpyDeviceToDevice, st));
    MB_CUDA(cudaMemcpyAsync(dst.gid, src.gid, size_t(n) * 8, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// a fresh container numbers its atoms 0, 1, ...: the default global id is the index
__global__ void iotaIdKernel(long long* gid, int64_t first, int64_t n)
{
    const int64_t i = first + blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i < n) gid[i] = i;
}
static int copyMolsView(const MolsView& dst, const MolsView& src, int64_t n, cudaStream_t st)
{
    if (n <= 0) return 0;
    MB_CUDA(cudaMemcpyAsync(dst.pos, src.pos, size_t(n) * 32, cudaMemcpyDeviceToDevice, st));
    MB_CUDA(cudaMemcpyAsync(dst.w, src.w, size_t(n) * 32, cudaMemcpyDeviceToDevice, st));
    MB_CUDA(cudaMemcpyAsync(dst.oc, src.oc, size_t(n) * 16, cudaMemcpyDeviceToDevice, st));
    MB_CUDA(cudaMemcpyAsync(dst.lambda, src.lambda, size_t(n) * 8, cudaMemcpyDeviceToDevice, st));
    for (int d = 0; d < 3; ++d)
        MB_CUDA(cudaMemcpyAsync(dst.force[d], src.force[d], size_t(n) * 8, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int atomsEnsureCapacity(mrmd_b200_atoms* a, int64_t capacity, cudaStream_t st)
{
    if (capacity <= a->capacity) return 0;
    int64_t cap = capacity + capacity / 8 + 32;
    cap = (cap + 31) & ~int64_t(31);
    void* slab = nullptr;
    MB_CUDA(cudaMalloc(&slab, atomsSlabBytes(cap)));
    MB_CUDA(cudaMemsetAsync(slab, 0, atomsSlabBytes(cap), st));
    AtomsView nv;
    carveAtoms(slab, cap, nv);
    MB_TRY(copyAtomsView(nv, a->v, a->size, st));
    if (cap > a->size)
    {
        iotaIdKernel<<<gridFor(cap - a->size, 256), 256, 0, st>>>(nv.gid, a->size, cap);
        MB_LAUNCHED();
    }
    if (a->v.pos != nullptr)
    {
        MB_CUDA(cudaStreamSynchronize(st));
        cudaFree(a->v.pos);
    }
    a->v = nv;
    a->capacity = cap;
    return 0;
}

int atomsEnsureAlt(mrmd_b200_atoms* a, cudaStream_t st)
{
    if (a->altCapacity == a->capacity && a->alt.pos != nullptr) return 0;
    if (a->alt.pos != nullptr)
    {
        MB_CUDA(cudaStreamSynchronize(st));
        cudaFree(a->alt.pos);
        a->alt.pos = nullptr;
    }
    void* slab = nullptr;
    MB_CUDA(cudaMalloc(&slab, atomsSlabBytes(a->capacity)));
    MB_CUDA(cudaMemsetAsync(slab, 0, atomsSlabBytes(a->capacity), st));
    carveAtoms(slab, a->capacity, a->alt);
    iotaIdKernel<<<gridFor(a->capacity, 256), 256, 0, st>>>(a->alt.gid, 0, a->capacity);
    MB_LAUNCHED();
    a->altCapacity = a->capacity;
    return 0;
}

int molsEnsureCapacity(mrmd_b200_molecules* m, int64_t capacity, cudaStream_t st)
{
    if (capacity <= m->capacity) return 0;
    int64_t cap = capacity + capacity / 8 + 32;
    cap = (cap + 31) & ~int64_t(31);
    void* slab = nullptr;
    MB_CUDA(cudaMalloc(&slab, molsSlabBytes(cap)));
    MB_CUDA(cudaMemsetAsync(slab, 0, molsSlabBytes(cap), st));
    MolsView nv;
    carveMols(slab, cap, nv);
    MB_TRY(copyMolsView(nv, m->v, m->size, st));
    if (m->v.pos != nullptr)
    {
        MB_CUDA(cudaStreamSynchronize(st));
        cudaFree(m->v.pos);
    }
    m->v = nv;
    m->capacity = cap;
    return 0;
}

int molsEnsureAlt(mrmd_b200_molecules* m, cudaStream_t st)
{
    if (m->altCapacity == m->capacity && m->alt.pos != nullptr) return 0;
    if (m->alt.pos != nullptr)
    {
        MB_CUDA(cudaStreamSynchronize(st));
        cudaFree(m->alt.pos);
        m->alt.pos = nullptr;
    }
    void* slab = nullptr;
    MB_CUDA(cudaMalloc(&slab, molsSlabBytes(m->capacity)));
    MB_CUDA(cudaMemsetAsync(slab, 0, molsSlabBytes(m->capacity), st));
    carveMols(slab, m->capacity, m->alt);
    m->altCapacity = m->capacity;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// slice transfer kernels.  buffer element (j, d): buf[(j / vlen) * stride + d * vlen + (j % vlen)]
__device__ __forceinline__ int64_t sliceIndex(int64_t j, int d, int64_t stride, int64_t vlen)
{
    return (j / vlen) * stride + int64_t(d) * vlen + (j % vlen);
}

// component accessors into the internal layout; comp 3 of pos4 carries the type bits
template <bool WRITE>
__global__ void atomFieldKernel(AtomsView v, int field, double* buf, int64_t first, int64_t count, int64_t stride,
                                int64_t vlen)
{
    const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (j >= count) return;
    const int64_t i = first + j;
    switch (field)
    {
        case MRMD_B200_ATOM_POS:
        {
            double* p = reinterpret_cast<double*>(v.pos + i);
            for (int d = 0; d < 3; ++d)
            {
                if (WRITE) p[d] = buf[sliceIndex(j, d, stride, vlen)];
                else buf[sliceIndex(j, d, stride, vlen)] = p[d];
            }
            break;
        }
        case MRMD_B200_ATOM_TYPE:
        {
            double* p = reinterpret_cast<double*>(v.pos + i) + 3;  // raw 64-bit copy of the int64
            if (WRITE) *p = buf[sliceIndex(j, 0, stride, vlen)];
            else buf[sliceIndex(j, 0, stride, vlen)] = *p;
            break;
        }
        case MRMD_B200_ATOM_VEL:
        case MRMD_B200_ATOM_FORCE:
        {
            double* const* pl = (field == MRMD_B200_ATOM_VEL) ? v.vel : v.force;
            for (int d = 0; d < 3; ++d)
            {
                if (WRITE) pl[d][i] = buf[sliceIndex(j, d, stride, vlen)];
                else buf[sliceIndex(j, d, stride, vlen)] = pl[d][i];
            }
            break;
        }
        case MRMD_B200_ATOM_ID:
        {
            double* p = reinterpret_cast<double*>(v.gid + i);  // raw 64-bit copy of the int64
            if (WRITE) *p = buf[sliceIndex(j, 0, stride, vlen)];
            else buf[sliceIndex(j, 0, stride, vlen)] = *p;
            break;
        }
        default:
        {
            double* pl = (field == MRMD_B200_ATOM_MASS) ? v.mass : ((field == MRMD_B200_ATOM_CHARGE) ? v.charge : v.relMass);
            if (WRITE) pl[i] = buf[sliceIndex(j, 0, stride, vlen)];
            else buf[sliceIndex(j, 0, stride, vlen)] = pl[i];
        }
    }
}

template <bool WRITE>
__global__ void molFieldKernel(MolsView v, int field, double* buf, int64_t first, int64_t count, int64_t stride,
                               int64_t vlen)
{
    const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (j >= count) return;
    const int64_t i = first + j;
    switch (field)
    {
        case MRMD_B200_MOL_POS:
        {
            double* p = reinterpret_cast<double*>(v.pos + i);
            for (int d = 0; d < 3; ++d)
            {
                if (WRITE) p[d] = buf[sliceIndex(j, d, stride, vlen)];
                else buf[sliceIndex(j, d, stride, vlen)] = p[d];
            }
            break;
        }
        case MRMD_B200_MOL_FORCE:
            for (int d = 0; d < 3; ++d)
            {
                if (WRITE) v.force[d][i] = buf[sliceIndex(j, d, stride, vlen)];
                else buf[sliceIndex(j, d, stride, vlen)] = v.force[d][i];
            }
            break;
        case MRMD_B200_MOL_LAMBDA:
            if (WRITE) v.lambda[i] = buf[sliceIndex(j, 0, stride, vlen)];
            else buf[sliceIndex(j, 0, stride, vlen)] = v.lambda[i];
            break;
        case MRMD_B200_MOL_MODULATED_LAMBDA:
        {
            double* p = reinterpret_cast<double*>(v.w + i);
            if (WRITE) p[0] = buf[sliceIndex(j, 0, stride, vlen)];
            else buf[sliceIndex(j, 0, stride, vlen)] = p[0];
            break;
        }
        case MRMD_B200_MOL_GRAD_LAMBDA:
        {
            double* p = reinterpret_cast<double*>(v.w + i) + 1;
            for (int d = 0; d < 3; ++d)
            {
                if (WRITE) p[d] = buf[sliceIndex(j, d, stride, vlen)];
                else buf[sliceIndex(j, d, stride, vlen)] = p[d];
            }
            break;
        }
        default:
        {
            double* p = reinterpret_cast<double*>(v.oc + i) + ((field == MRMD_B200_MOL_ATOMS_OFFSET) ? 0 : 1);
            if (WRITE) *p = buf[sliceIndex(j, 0, stride, vlen)];
            else buf[sliceIndex(j, 0, stride, vlen)] = *p;
        }
    }
}

__global__ void fillPlaneKernel(double* p, int64_t n, double value)
{
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i < n) p[i] = value;
}
__global__ void fillPos4Kernel(double4* p, int64_t n, double value, int first, int last)
{
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    double* q = reinterpret_cast<double*>(p + i);
    for (int d = first; d <= last; ++d) q[d] = value;
}
__global__ void fillOcKernel(longlong2* p, int64_t n, long long value, int which)
{
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    if (which == 0) p[i].x = value;
    else p[i].y = value;
}
__global__ void moleculePerAtomKernel(MolsView m, int64_t n)
{
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    m.oc[i] = make_longlong2(i, 1);
}

static int ncompOf(bool atoms, int field)
{
    if (atoms) return (field <= MRMD_B200_ATOM_FORCE) ? 3 : 1;
    return (field == MRMD_B200_MOL_POS || field == MRMD_B200_MOL_FORCE || field == MRMD_B200_MOL_GRAD_LAMBDA) ? 3 : 1;
}
static size_t spanElems(int64_t count, int ncomp, int64_t stride, int64_t vlen)
{
    if (count <= 0) return 0;
    const int64_t j = count - 1;
    // largest index touched: last SoA block fully (conservative) or exact for vlen == 1
    if (vlen == 1) return size_t(j * stride + ncomp);
    return size_t((j / vlen) * stride + int64_t(ncomp) * vlen);
}

template <class Handle, class Launch>
static int transfer(Handle* h, bool isAtoms, int field, void* buf, int64_t first, int64_t count, int64_t stride,
                    int64_t vlen, int memKind, bool write, cudaStream_t st, Launch launch)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(h != nullptr, "null handle");
    MB_REQUIRE(field >= 0 && field <= (isAtoms ? 7 : 6), "unknown field");
    MB_REQUIRE(first >= 0 && count >= 0 && first + count <= h->size, "range outside the container");
    MB_REQUIRE(vlen >= 1, "vector length must be >= 1");
    if (count == 0) return 0;
    MB_REQUIRE(buf != nullptr, "null buffer");
    const int ncomp = ncompOf(isAtoms, field);
    MB_REQUIRE(stride >= int64_t(ncomp) * vlen, "stride smaller than ncomp * vlen");
    double* dbuf = static_cast<double*>(buf);
    const size_t span = spanElems(count, ncomp, stride, vlen) * 8;
    if (memKind == MRMD_B200_MEM_HOST)
    {
        MB_TRY(h->staging.reserve(span));
        dbuf = h->staging.template as<double>();
        if (write) MB_CUDA(cudaMemcpyAsync(dbuf, buf, span, cudaMemcpyHostToDevice, st));
        else if (!(vlen == 1 && stride == ncomp))
            MB_CUDA(cudaMemcpyAsync(dbuf, buf, span, cudaMemcpyHostToDevice, st));  // keep the gaps intact
    }
    launch(dbuf);
    MB_LAUNCHED();
    if (memKind == MRMD_B200_MEM_HOST && !write)
    {
        MB_CUDA(cudaMemcpyAsync(buf, dbuf, span, cudaMemcpyDeviceToHost, st));
        MB_CUDA(cudaStreamSynchronize(st));
    }
    return 0;
}

// dense (n x ncomp) device buffer <-> the internal planes of one field; used by the pipelined host-buffer path of
// md.cu, which owns its staging buffers and streams
int atomsFieldToDense(const mrmd_b200_atoms* a, int field, double* devBuf, int64_t n, cudaStream_t st)
{
    if (n <= 0) return 0;
    atomFieldKernel<false><<<gridFor(n, 256), 256, 0, st>>>(a->v, field, devBuf, 0, n, ncompOf(true, field), 1);
    MB_LAUNCHED();
    return 0;
}
int atomsFieldFromDense(mrmd_b200_atoms* a, int field, const double* devBuf, int64_t n, cudaStream_t st)
{
    if (field == MRMD_B200_ATOM_POS) a->posEpoch += 1;
    if (n <= 0) return 0;
    atomFieldKernel<true><<<gridFor(n, 256), 256, 0, st>>>(a->v, field, const_cast<double*>(devBuf), 0, n,
                                                          ncompOf(true, field), 1);
    MB_LAUNCHED();
    return 0;
}
}  // namespace mrmd_b200

using namespace mrmd_b200;

extern "C" {

const char* mrmd_b200_last_error(void)
{
    static thread_local std::string copy;
    {
        std::lock_guard<std::mutex> lock(g_errMutex);
        copy = g_lastError;
    }
    return copy.c_str();
}

int mrmd_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int mrmd_b200_set_device(int device)
{
    MB_TRY(checkDevice());
    MB_CUDA(cudaSetDevice(device));
    return 0;
}

int mrmd_b200_sync(void* stream)
{
    MB_TRY(checkDevice());
    MB_CUDA(cudaStreamSynchronize(S(stream)));
    return 0;
}

int64_t mrmd_b200_launch_count(void) { return g_launchCount.load(); }

// data/Subdomain.hpp:41-62
void mrmd_b200_subdomain_init(mrmd_b200_subdomain* s, const double* minCorner, const double* maxCorner,
                              const double* t)
{
    for (int d = 0; d < 3; ++d)
    {
        s->minCorner[d] = minCorner[d];
        s->maxCorner[d] = maxCorner[d];
        s->ghostLayerThickness[d] = t[d];
        s->minGhostCorner[d] = s->minCorner[d] - s->ghostLayerThickness[d];
        s->maxGhostCorner[d] = s->maxCorner[d] + s->ghostLayerThickness[d];
        s->minInnerCorner[d] = s->minCorner[d] + s->ghostLayerThickness[d];
        s->maxInnerCorner[d] = s->maxCorner[d] - s->ghostLayerThickness[d];
        s->diameter[d] = s->maxCorner[d] - s->minCorner[d];
        s->diameterWithGhostLayer[d] = s->maxCorner[d] - s->minCorner[d] + 2.0 * s->ghostLayerThickness[d];
    }
}

void mrmd_b200_subdomain_scale_dim(mrmd_b200_subdomain* s, double factor, int axis)
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
    mrmd_b200_subdomain_init(s, mn, mx, th);
}

// --- atoms -------------------------------------------------------------------------------------
int mrmd_b200_atoms_create(mrmd_b200_atoms** out, int64_t size)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(out != nullptr && size >= 0, "atoms_create");
    auto* a = new mrmd_b200_atoms;
    int rc = atomsEnsureCapacity(a, size > 0 ? size : 1, nullptr);
    if (rc == 0 && cudaMalloc(&a->dMaxDisp, 8) != cudaSuccess) rc = MRMD_B200_ENOMEM;
    if (rc == 0 && cudaMallocHost(&a->hMaxDisp, 8) != cudaSuccess) rc = MRMD_B200_ENOMEM;
    if (rc != 0)
    {
        delete a;
        return rc;
    }
    a->size = size;
    MB_CUDA(cudaStreamSynchronize(nullptr));
    *out = a;
    return 0;
}

int mrmd_b200_atoms_destroy(mrmd_b200_atoms* a)
{
    if (a == nullptr) return 0;
    cudaDeviceSynchronize();
    if (a->v.pos) cudaFree(a->v.pos);
    if (a->alt.pos) cudaFree(a->alt.pos);
    if (a->dMaxDisp) cudaFree(a->dMaxDisp);
    if (a->hMaxDisp) cudaFreeHost(a->hMaxDisp);
    a->staging.release();
    a->sortScratch.release();
    a->lcCellStart.release();
    delete a;
    return 0;
}

int mrmd_b200_atoms_reserve(mrmd_b200_atoms* a, int64_t capacity, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(a != nullptr && capacity >= 0, "atoms_reserve");
    return atomsEnsureCapacity(a, capacity, S(stream));
}

int mrmd_b200_atoms_resize(mrmd_b200_atoms* a, int64_t size, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(a != nullptr && size >= 0, "atoms_resize");
    MB_TRY(atomsEnsureCapacity(a, size, S(stream)));
    a->size = size;
    if (size < a->lcEnd) a->lcValid = false;
    return 0;
}

int64_t mrmd_b200_atoms_size(const mrmd_b200_atoms* a) { return a ? a->size : 0; }

int mrmd_b200_atoms_set_counts(mrmd_b200_atoms* a, int64_t numLocal, int64_t numGhost)
{
    MB_REQUIRE(a != nullptr && numLocal >= 0 && numGhost >= 0, "atoms_set_counts");
    a->numLocal = numLocal;
    a->numGhost = numGhost;
    return 0;
}

int mrmd_b200_atoms_get_counts(const mrmd_b200_atoms* a, int64_t* numLocal, int64_t* numGhost)
{
    MB_REQUIRE(a != nullptr, "atoms_get_counts");
    if (numLocal) *numLocal = a->numLocal;
    if (numGhost) *numGhost = a->numGhost;
    return 0;
}

int mrmd_b200_atoms_write(mrmd_b200_atoms* a, int field, const void* src, int64_t first, int64_t count,
                          int64_t stride, int64_t vlen, int memKind, void* stream)
{
    cudaStream_t st = S(stream);
    if (a != nullptr && field == MRMD_B200_ATOM_POS) a->posEpoch += 1;
    return transfer(a, true, field, const_cast<void*>(src), first, count, stride, vlen, memKind, true, st,
                    [&](double* dbuf) {
                        atomFieldKernel<true><<<gridFor(count, 256), 256, 0, st>>>(a->v, field, dbuf, first, count,
                                                                                   stride, vlen);
                    });
}

int mrmd_b200_atoms_read(const mrmd_b200_atoms* a, int field, void* dst, int64_t first, int64_t count,
                         int64_t stride, int64_t vlen, int memKind, void* stream)
{
    cudaStream_t st = S(stream);
    auto* h = const_cast<mrmd_b200_atoms*>(a);
    return transfer(h, true, field, dst, first, count, stride, vlen, memKind, false, st, [&](double* dbuf) {
        atomFieldKernel<false><<<gridFor(count, 256), 256, 0, st>>>(h->v, field, dbuf, first, count, stride, vlen);
    });
}

int mrmd_b200_atoms_fill(mrmd_b200_atoms* a, int field, double value, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(a != nullptr && field >= 0 && field <= 6, "atoms_fill");
    const int64_t n = a->size;
    if (n == 0) return 0;
    cudaStream_t st = S(stream);
    const int g = gridFor(n, 256);
    if (value == 0.0 && field != MRMD_B200_ATOM_POS && field != MRMD_B200_ATOM_TYPE)
    {
        double* planes[3] = {nullptr, nullptr, nullptr};
        int np = 1;
        if (field == MRMD_B200_ATOM_VEL || field == MRMD_B200_ATOM_FORCE)
        {
            np = 3;
            for (int d = 0; d < 3; ++d) planes[d] = (field == MRMD_B200_ATOM_VEL) ? a->v.vel[d] : a->v.force[d];
        }
        else
            planes[0] = (field == MRMD_B200_ATOM_MASS) ? a->v.mass : ((field == MRMD_B200_ATOM_CHARGE) ? a->v.charge : a->v.relMass);
        for (int d = 0; d < np; ++d) MB_CUDA(cudaMemsetAsync(planes[d], 0, size_t(n) * 8, st));
        return 0;
    }
    switch (field)
    {
        case MRMD_B200_ATOM_POS:
            a->posEpoch += 1;
            fillPos4Kernel<<<g, 256, 0, st>>>(a->v.pos, n, value, 0, 2);
            MB_LAUNCHED();
            break;
        case MRMD_B200_ATOM_TYPE:
        {
            const long long t = static_cast<long long>(value);
            double bits;
            std::memcpy(&bits, &t, 8);
            fillPos4Kernel<<<g, 256, 0, st>>>(a->v.pos, n, bits, 3, 3);
            MB_LAUNCHED();
            break;
        }
        case MRMD_B200_ATOM_VEL:
        case MRMD_B200_ATOM_FORCE:
            for (int d = 0; d < 3; ++d)
            {
                fillPlaneKernel<<<g, 256, 0, st>>>((field == MRMD_B200_ATOM_VEL) ? a->v.vel[d] : a->v.force[d], n, value);
                MB_LAUNCHED();
            }
            break;
        default:
            fillPlaneKernel<<<g, 256, 0, st>>>(
                (field == MRMD_B200_ATOM_MASS) ? a->v.mass : ((field == MRMD_B200_ATOM_CHARGE) ? a->v.charge : a->v.relMass), n, value);
            MB_LAUNCHED();
    }
    return 0;
}

int mrmd_b200_atoms_copy(mrmd_b200_atoms* dst, const mrmd_b200_atoms* src, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(dst != nullptr && src != nullptr, "atoms_copy");
    MB_TRY(atomsEnsureCapacity(dst, src->size, S(stream)));
    dst->size = src->size;
    dst->numLocal = src->numLocal;
    dst->numGhost = src->numGhost;
    dst->posEpoch += 1;
    dst->lcValid = false;
    return copyAtomsView(dst->v, src->v, src->size, S(stream));
}

// --- molecules ---------------------------------------------------------------------------------
int mrmd_b200_molecules_create(mrmd_b200_molecules** out, int64_t size)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(out != nullptr && size >= 0, "molecules_create");
    auto* m = new mrmd_b200_molecules;
    int rc = molsEnsureCapacity(m, size > 0 ? size : 1, nullptr);
    if (rc != 0)
    {
        delete m;
        return rc;
    }
    m->size = size;
    MB_CUDA(cudaStreamSynchronize(nullptr));
    *out = m;
    return 0;
}

int mrmd_b200_molecules_destroy(mrmd_b200_molecules* m)
{
    if (m == nullptr) return 0;
    cudaDeviceSynchronize();
    if (m->v.pos) cudaFree(m->v.pos);
    if (m->alt.pos) cudaFree(m->alt.pos);
    m->staging.release();
    m->sortScratch.release();
    if (m->lcView != nullptr)
    {
        m->lcView->lcCellStart.release();
        delete m->lcView;
    }
    delete m;
    return 0;
}

int mrmd_b200_molecules_resize(mrmd_b200_molecules* m, int64_t size, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(m != nullptr && size >= 0, "molecules_resize");
    MB_TRY(molsEnsureCapacity(m, size, S(stream)));
    m->size = size;
    return 0;
}

int64_t mrmd_b200_molecules_size(const mrmd_b200_molecules* m) { return m ? m->size : 0; }

int mrmd_b200_molecules_set_counts(mrmd_b200_molecules* m, int64_t numLocal, int64_t numGhost)
{
    MB_REQUIRE(m != nullptr && numLocal >= 0 && numGhost >= 0, "molecules_set_counts");
    m->numLocal = numLocal;
    m->numGhost = numGhost;
    return 0;
}

int mrmd_b200_molecules_get_counts(const mrmd_b200_molecules* m, int64_t* numLocal, int64_t* numGhost)
{
    MB_REQUIRE(m != nullptr, "molecules_get_counts");
    if (numLocal) *numLocal = m->numLocal;
    if (numGhost) *numGhost = m->numGhost;
    return 0;
}

int mrmd_b200_molecules_write(mrmd_b200_molecules* m, int field, const void* src, int64_t first, int64_t count,
                              int64_t stride, int64_t vlen, int memKind, void* stream)
{
    cudaStream_t st = S(stream);
    return transfer(m, false, field, const_cast<void*>(src), first, count, stride, vlen, memKind, true, st,
                    [&](double* dbuf) {
                        molFieldKernel<true><<<gridFor(count, 256), 256, 0, st>>>(m->v, field, dbuf, first, count,
                                                                                  stride, vlen);
                    });
}

int mrmd_b200_molecules_read(const mrmd_b200_molecules* m, int field, void* dst, int64_t first, int64_t count,
                             int64_t stride, int64_t vlen, int memKind, void* stream)
{
    cudaStream_t st = S(stream);
    auto* h = const_cast<mrmd_b200_molecules*>(m);
    return transfer(h, false, field, dst, first, count, stride, vlen, memKind, false, st, [&](double* dbuf) {
        molFieldKernel<false><<<gridFor(count, 256), 256, 0, st>>>(h->v, field, dbuf, first, count, stride, vlen);
    });
}

int mrmd_b200_molecules_fill(mrmd_b200_molecules* m, int field, double value, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(m != nullptr && field >= 0 && field <= 6, "molecules_fill");
    const int64_t n = m->size;
    if (n == 0) return 0;
    cudaStream_t st = S(stream);
    const int g = gridFor(n, 256);
    switch (field)
    {
        case MRMD_B200_MOL_POS: fillPos4Kernel<<<g, 256, 0, st>>>(m->v.pos, n, value, 0, 2); break;
        case MRMD_B200_MOL_MODULATED_LAMBDA: fillPos4Kernel<<<g, 256, 0, st>>>(m->v.w, n, value, 0, 0); break;
        case MRMD_B200_MOL_GRAD_LAMBDA: fillPos4Kernel<<<g, 256, 0, st>>>(m->v.w, n, value, 1, 3); break;
        case MRMD_B200_MOL_LAMBDA: fillPlaneKernel<<<g, 256, 0, st>>>(m->v.lambda, n, value); break;
        case MRMD_B200_MOL_FORCE:
            for (int d = 0; d < 2; ++d)
            {
                fillPlaneKernel<<<g, 256, 0, st>>>(m->v.force[d], n, value);
                MB_LAUNCHED();
            }
            fillPlaneKernel<<<g, 256, 0, st>>>(m->v.force[2], n, value);
            break;
        case MRMD_B200_MOL_ATOMS_OFFSET: fillOcKernel<<<g, 256, 0, st>>>(m->v.oc, n, (long long)value, 0); break;
        default: fillOcKernel<<<g, 256, 0, st>>>(m->v.oc, n, (long long)value, 1);
    }
    MB_LAUNCHED();
    return 0;
}

// data/MoleculesFromAtoms.cpp:19-39
int mrmd_b200_molecules_for_each_atom(mrmd_b200_molecules** out, const mrmd_b200_atoms* a, void* stream)
{
    MB_TRY(checkDevice());
    MB_REQUIRE(out != nullptr && a != nullptr, "molecules_for_each_atom");
    const int64_t size = a->numLocal + a->numGhost;
    mrmd_b200_molecules* m = nullptr;
    MB_TRY(mrmd_b200_molecules_create(&m, 2 * size));
    if (size > 0)
    {
        moleculePerAtomKernel<<<gridFor(size, 256), 256, 0, S(stream)>>>(m->v, size);
        MB_LAUNCHED();
    }
    m->numLocal = a->numLocal;
    m->numGhost = a->numGhost;
    *out = m;
    return 0;
}

}  // extern "C"
