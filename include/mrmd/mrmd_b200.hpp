// mrmd_b200.hpp -- C++20 host-side mirror of the reference's public API for the hot path.
//
// Same namespaces, class names and member signatures as XzzX/mrmd (data::Atoms, data::Molecules,
// data::Subdomain, HalfVerletList::build, action::LennardJones::apply/apply_if, action::LJ_IdealGas::run,
// action::UpdateMolecules, action::ThermodynamicForce, action::VelocityVerlet,
// action::VelocityVerletLangevinThermostat, communication::GhostLayer / MultiResGhostLayer,
// weighting_function::Slab / Spherical, util::IsInSymmetricSlab), each member forwarding to the C ABI of
// include/mrmd_b200.h -- no Kokkos, no Cabana.  The per-file headers next to this one
// (action/LennardJones.hpp, data/Atoms.hpp, ...) only include this file, so a driver written against the
// reference keeps its #include lines.
//
// Differences a port of a reference driver has to know (all forced by the C boundary):
//   * device lambdas cannot cross a C ABI: predicates are parametric objects (util::IsInSymmetricSlab and
//     either()/both() combinations of it) instead of KOKKOS_LAMBDAs;
//   * slices (getPos() ...) are handles for the operators; element access from the host goes through
//     data::HostAtoms / HostMolecules + deep_copy, as in the reference's unit tests;
//   * errors follow the reference: message on stderr + abort (MRMD_HOST_CHECK behaviour, assert/verbose.hpp).
#pragma once

#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "../mrmd_b200.h"

namespace mrmd
{
using idx_t = int64_t;   // datatypes.hpp:91
using real_t = double;   // datatypes.hpp:99
using Point3D = std::array<real_t, 3>;
using Vector3D = std::array<real_t, 3>;
constexpr idx_t DIMENSIONS = 3;
constexpr real_t pi = 3.14159265358979323846;

constexpr real_t operator""_r(long double v) { return static_cast<real_t>(v); }
constexpr real_t operator""_r(unsigned long long v) { return static_cast<real_t>(v); }
template <typename T>
constexpr real_t real_c(T t) { return static_cast<real_t>(t); }
template <typename T>
constexpr idx_t idx_c(const T& v) { return static_cast<idx_t>(v); }

enum class AXIS : idx_t { X = 0, Y = 1, Z = 2 };  // datatypes.hpp:140-145
template <class Enum>
constexpr std::underlying_type_t<Enum> to_underlying(Enum e) noexcept { return static_cast<std::underlying_type_t<Enum>>(e); }

namespace detail
{
[[noreturn]] inline void fail(int rc, const char* what)
{
    std::fprintf(stderr, "mrmd_b200: %s failed (code %d): %s\n", what, rc, mrmd_b200_last_error());
    std::abort();
}
inline void check(int rc, const char* what)
{
    if (rc != 0) fail(rc, what);
}
}  // namespace detail

inline void* defaultStream = nullptr;  ///< cudaStream_t used by every wrapper call (NULL stream by default)
inline void fence() { detail::check(mrmd_b200_sync(defaultStream), "sync"); }

// ------------------------------------------------------------------------------------------------------
namespace util
{
/// util::ExponentialMovingAverage (util/ExponentialMovingAverage.hpp:23-60): host-side running average of the drivers'
/// statistics (tests/NVT/NVT.cpp:131-132)
class ExponentialMovingAverage
{
public:
    explicit ExponentialMovingAverage(const real_t& weightingFactor) : alpha_(weightingFactor), beta_(1_r - weightingFactor) {}
    operator real_t() const { return val_; }
    real_t toReal() const { return val_; }
    void append(const real_t& val)
    {
        val_ = isFirstVal_ ? val : val * alpha_ + val_ * beta_;
        isFirstVal_ = false;
    }

private:
    real_t alpha_, beta_;
    real_t val_ = 0_r;
    bool isFirstVal_ = true;
};
inline ExponentialMovingAverage& operator<<(ExponentialMovingAverage& lhs, const real_t& rhs)
{
    lhs.append(rhs);
    return lhs;
}

/// util::IsInSymmetricSlab (util/IsInSymmetricSlab.hpp:24-64) as a parametric predicate
class IsInSymmetricSlab
{
public:
    IsInSymmetricSlab(const Point3D& center, real_t slabMin, real_t slabMax, AXIS axis = AXIS::X, real_t tolerance = 0_r)
    {
        desc_ = {MRMD_B200_PRED_SLAB, static_cast<int32_t>(to_underlying(axis)), center[to_underlying(axis)], slabMin,
                 slabMax, tolerance};
    }
    bool operator()(real_t x, real_t y, real_t z) const
    {
        const real_t c[3] = {x, y, z};
        const real_t absDx = std::abs(c[desc_.axis] - desc_.center);
        return absDx >= desc_.slabMin - desc_.tolerance && absDx <= desc_.slabMax + desc_.tolerance;
    }
    bool operator()(real_t coord) const
    {
        const real_t absDx = std::abs(coord - desc_.center);
        return absDx >= desc_.slabMin - desc_.tolerance && absDx <= desc_.slabMax + desc_.tolerance;
    }
    /// pred(p1) || pred(p2)  (examples/04_LennardJones_IdealGas_LocalCap.cpp:209-216)
    mrmd_b200_pred either() const { auto d = desc_; d.kind = MRMD_B200_PRED_SLAB_EITHER; return d; }
    /// pred(p1) && pred(p2)  (examples/04:220-227)
    mrmd_b200_pred both() const { auto d = desc_; d.kind = MRMD_B200_PRED_SLAB_BOTH; return d; }
    const mrmd_b200_pred& desc() const { return desc_; }

private:
    mrmd_b200_pred desc_{};
};
}  // namespace util

// ------------------------------------------------------------------------------------------------------
namespace data
{
/// data::Subdomain (data/Subdomain.hpp:37-110)
struct Subdomain
{
    Subdomain() = default;
    Subdomain(const Point3D& minCornerArg, const Point3D& maxCornerArg, const Vector3D& ghostLayerThicknessArg)
    {
        mrmd_b200_subdomain s;
        mrmd_b200_subdomain_init(&s, minCornerArg.data(), maxCornerArg.data(), ghostLayerThicknessArg.data());
        from(s);
    }
    Subdomain(const Point3D& mn, const Point3D& mx, const real_t& t) : Subdomain(mn, mx, Vector3D{t, t, t}) {}
    void scaleDim(const real_t& f, const AXIS& axis)
    {
        auto s = c();
        mrmd_b200_subdomain_scale_dim(&s, f, static_cast<int>(to_underlying(axis)));
        from(s);
    }
    void scale(const real_t& f)
    {
        scaleDim(f, AXIS::X);
        scaleDim(f, AXIS::Y);
        scaleDim(f, AXIS::Z);
    }
    real_t getVolume() const { return diameter[0] * diameter[1] * diameter[2]; }
    Point3D getCenter() const
    {
        return {(minCorner[0] + maxCorner[0]) * 0.5_r, (minCorner[1] + maxCorner[1]) * 0.5_r, (minCorner[2] + maxCorner[2]) * 0.5_r};
    }
    mrmd_b200_subdomain c() const
    {
        mrmd_b200_subdomain s;
        for (int d = 0; d < 3; ++d)
        {
            s.minCorner[d] = minCorner[d];
            s.maxCorner[d] = maxCorner[d];
            s.ghostLayerThickness[d] = ghostLayerThickness[d];
            s.minGhostCorner[d] = minGhostCorner[d];
            s.maxGhostCorner[d] = maxGhostCorner[d];
            s.minInnerCorner[d] = minInnerCorner[d];
            s.maxInnerCorner[d] = maxInnerCorner[d];
            s.diameter[d] = diameter[d];
            s.diameterWithGhostLayer[d] = diameterWithGhostLayer[d];
        }
        return s;
    }
    Point3D minCorner{}, maxCorner{}, ghostLayerThickness{}, minGhostCorner{}, maxGhostCorner{}, minInnerCorner{},
        maxInnerCorner{};
    Vector3D diameter{}, diameterWithGhostLayer{};

private:
    void from(const mrmd_b200_subdomain& s)
    {
        for (int d = 0; d < 3; ++d)
        {
            minCorner[d] = s.minCorner[d];
            maxCorner[d] = s.maxCorner[d];
            ghostLayerThickness[d] = s.ghostLayerThickness[d];
            minGhostCorner[d] = s.minGhostCorner[d];
            maxGhostCorner[d] = s.maxGhostCorner[d];
            minInnerCorner[d] = s.minInnerCorner[d];
            maxInnerCorner[d] = s.maxInnerCorner[d];
            diameter[d] = s.diameter[d];
            diameterWithGhostLayer[d] = s.diameterWithGhostLayer[d];
        }
    }
};

class Atoms;
class Molecules;

/// position slice handle: what drivers pass to VerletList::build / LinkedCellList
struct AtomsPosSlice { const Atoms* owner; };
struct MoleculesPosSlice { const Molecules* owner; };
/// force slice handle: target of Cabana::deep_copy(force, value)
struct AtomsForceSlice { const Atoms* owner; };

/// host mirror with element access, as data::HostAtoms in the reference's unit tests
template <int NCOMP, class T = real_t>
struct HostSlice
{
    std::vector<T>* v;
    T& operator()(idx_t i, int d = 0) const { return (*v)[static_cast<size_t>(i) * NCOMP + d]; }
};

class HostAtoms
{
public:
    explicit HostAtoms(idx_t n = 0) { resize(static_cast<size_t>(n)); }
    void resize(size_t n)
    {
        pos.resize(3 * n); vel.resize(3 * n); force.resize(3 * n);
        type.resize(n); mass.resize(n); charge.resize(n); relativeMass.resize(n);
    }
    size_t size() const { return type.size(); }
    HostSlice<3> getPos() { return {&pos}; }
    HostSlice<3> getVel() { return {&vel}; }
    HostSlice<3> getForce() { return {&force}; }
    HostSlice<1, idx_t> getType() { return {&type}; }
    HostSlice<1> getMass() { return {&mass}; }
    HostSlice<1> getCharge() { return {&charge}; }
    HostSlice<1> getRelativeMass() { return {&relativeMass}; }
    idx_t numLocalAtoms = 0, numGhostAtoms = 0;
    std::vector<real_t> pos, vel, force, mass, charge, relativeMass;
    std::vector<idx_t> type;
};

/// data::Atoms (data/Atoms.hpp:33-147): value-semantic handle to device data (copies share the data, like Kokkos views)
class Atoms
{
public:
    explicit Atoms(const idx_t numAtoms)
    {
        mrmd_b200_atoms* h = nullptr;
        detail::check(mrmd_b200_atoms_create(&h, numAtoms), "Atoms");
        h_.reset(h, [](mrmd_b200_atoms* p) { mrmd_b200_atoms_destroy(p); });
    }
    AtomsPosSlice getPos() const { return {this}; }
    AtomsForceSlice getForce() const { return {this}; }
    void setForce(const real_t& val) const { detail::check(mrmd_b200_atoms_fill(h_.get(), MRMD_B200_ATOM_FORCE, val, defaultStream), "setForce"); }
    auto size() const { return static_cast<size_t>(mrmd_b200_atoms_size(h_.get())); }
    void resize(size_t size) { detail::check(mrmd_b200_atoms_resize(h_.get(), idx_c(size), defaultStream), "resize"); }
    void removeGhostAtoms()
    {
        numGhostAtoms = 0;
        resize(static_cast<size_t>(numLocalAtoms));
    }
    mrmd_b200_atoms* handle() const { return h_.get(); }
    /// push the public counters into the device container / pull them back (called by every operator)
    void push() const { detail::check(mrmd_b200_atoms_set_counts(h_.get(), numLocalAtoms, numGhostAtoms), "counts"); }
    void pull() { detail::check(mrmd_b200_atoms_get_counts(h_.get(), &numLocalAtoms, &numGhostAtoms), "counts"); }

    idx_t numLocalAtoms = 0;
    idx_t numGhostAtoms = 0;

private:
    std::shared_ptr<mrmd_b200_atoms> h_;
};
using DeviceAtoms = Atoms;

inline void deep_copy(Atoms& dst, const HostAtoms& src)
{
    dst.numLocalAtoms = src.numLocalAtoms;
    dst.numGhostAtoms = src.numGhostAtoms;
    dst.resize(src.size());
    dst.push();
    const idx_t n = idx_c(src.size());
    auto* h = dst.handle();
    auto w = [&](int field, const void* p, int ncomp) { detail::check(mrmd_b200_atoms_write(h, field, p, 0, n, ncomp, 1, MRMD_B200_MEM_HOST, defaultStream), "deep_copy"); };
    w(MRMD_B200_ATOM_POS, src.pos.data(), 3); w(MRMD_B200_ATOM_VEL, src.vel.data(), 3); w(MRMD_B200_ATOM_FORCE, src.force.data(), 3);
    w(MRMD_B200_ATOM_TYPE, src.type.data(), 1); w(MRMD_B200_ATOM_MASS, src.mass.data(), 1);
    w(MRMD_B200_ATOM_CHARGE, src.charge.data(), 1); w(MRMD_B200_ATOM_RELATIVE_MASS, src.relativeMass.data(), 1);
    fence();
}
inline void deep_copy(HostAtoms& dst, const Atoms& src)
{
    dst.numLocalAtoms = src.numLocalAtoms;
    dst.numGhostAtoms = src.numGhostAtoms;
    dst.resize(src.size());
    const idx_t n = idx_c(src.size());
    auto* h = src.handle();
    auto r = [&](int field, void* p, int ncomp) { detail::check(mrmd_b200_atoms_read(h, field, p, 0, n, ncomp, 1, MRMD_B200_MEM_HOST, defaultStream), "deep_copy"); };
    r(MRMD_B200_ATOM_POS, dst.pos.data(), 3); r(MRMD_B200_ATOM_VEL, dst.vel.data(), 3); r(MRMD_B200_ATOM_FORCE, dst.force.data(), 3);
    r(MRMD_B200_ATOM_TYPE, dst.type.data(), 1); r(MRMD_B200_ATOM_MASS, dst.mass.data(), 1);
    r(MRMD_B200_ATOM_CHARGE, dst.charge.data(), 1); r(MRMD_B200_ATOM_RELATIVE_MASS, dst.relativeMass.data(), 1);
}

class HostMolecules
{
public:
    explicit HostMolecules(idx_t n = 0) { resize(static_cast<size_t>(n)); }
    void resize(size_t n)
    {
        pos.resize(3 * n); force.resize(3 * n); gradLambda.resize(3 * n);
        lambda.resize(n); modulatedLambda.resize(n); atomsOffset.resize(n); numAtoms.resize(n);
    }
    size_t size() const { return lambda.size(); }
    HostSlice<3> getPos() { return {&pos}; }
    HostSlice<3> getForce() { return {&force}; }
    HostSlice<1> getLambda() { return {&lambda}; }
    HostSlice<1> getModulatedLambda() { return {&modulatedLambda}; }
    HostSlice<3> getGradLambda() { return {&gradLambda}; }
    HostSlice<1, idx_t> getAtomsOffset() { return {&atomsOffset}; }
    HostSlice<1, idx_t> getNumAtoms() { return {&numAtoms}; }
    idx_t numLocalMolecules = 0, numGhostMolecules = 0;
    std::vector<real_t> pos, force, lambda, modulatedLambda, gradLambda;
    std::vector<idx_t> atomsOffset, numAtoms;
};

/// data::Molecules (data/Molecules.hpp:27-148)
class Molecules
{
public:
    explicit Molecules(const idx_t numMolecules)
    {
        mrmd_b200_molecules* h = nullptr;
        detail::check(mrmd_b200_molecules_create(&h, numMolecules), "Molecules");
        h_.reset(h, [](mrmd_b200_molecules* p) { mrmd_b200_molecules_destroy(p); });
    }
    explicit Molecules(mrmd_b200_molecules* adopt) { h_.reset(adopt, [](mrmd_b200_molecules* p) { mrmd_b200_molecules_destroy(p); }); }
    MoleculesPosSlice getPos() const { return {this}; }
    void setForce(const real_t& val) const { detail::check(mrmd_b200_molecules_fill(h_.get(), MRMD_B200_MOL_FORCE, val, defaultStream), "setForce"); }
    idx_t size() const { return mrmd_b200_molecules_size(h_.get()); }
    void resize(size_t size) { detail::check(mrmd_b200_molecules_resize(h_.get(), idx_c(size), defaultStream), "resize"); }
    mrmd_b200_molecules* handle() const { return h_.get(); }
    void push() const { detail::check(mrmd_b200_molecules_set_counts(h_.get(), numLocalMolecules, numGhostMolecules), "counts"); }
    void pull() { detail::check(mrmd_b200_molecules_get_counts(h_.get(), &numLocalMolecules, &numGhostMolecules), "counts"); }
    idx_t numLocalMolecules = 0;
    idx_t numGhostMolecules = 0;

private:
    std::shared_ptr<mrmd_b200_molecules> h_;
};

inline void deep_copy(Molecules& dst, const HostMolecules& src)
{
    dst.numLocalMolecules = src.numLocalMolecules;
    dst.numGhostMolecules = src.numGhostMolecules;
    dst.resize(src.size());
    dst.push();
    const idx_t n = idx_c(src.size());
    auto* h = dst.handle();
    auto w = [&](int field, const void* p, int ncomp) { detail::check(mrmd_b200_molecules_write(h, field, p, 0, n, ncomp, 1, MRMD_B200_MEM_HOST, defaultStream), "deep_copy"); };
    w(MRMD_B200_MOL_POS, src.pos.data(), 3); w(MRMD_B200_MOL_FORCE, src.force.data(), 3); w(MRMD_B200_MOL_LAMBDA, src.lambda.data(), 1);
    w(MRMD_B200_MOL_MODULATED_LAMBDA, src.modulatedLambda.data(), 1); w(MRMD_B200_MOL_GRAD_LAMBDA, src.gradLambda.data(), 3);
    w(MRMD_B200_MOL_ATOMS_OFFSET, src.atomsOffset.data(), 1); w(MRMD_B200_MOL_NUM_ATOMS, src.numAtoms.data(), 1);
    fence();
}
inline void deep_copy(HostMolecules& dst, const Molecules& src)
{
    dst.numLocalMolecules = src.numLocalMolecules;
    dst.numGhostMolecules = src.numGhostMolecules;
    dst.resize(static_cast<size_t>(src.size()));
    const idx_t n = src.size();
    auto* h = src.handle();
    auto r = [&](int field, void* p, int ncomp) { detail::check(mrmd_b200_molecules_read(h, field, p, 0, n, ncomp, 1, MRMD_B200_MEM_HOST, defaultStream), "deep_copy"); };
    r(MRMD_B200_MOL_POS, dst.pos.data(), 3); r(MRMD_B200_MOL_FORCE, dst.force.data(), 3); r(MRMD_B200_MOL_LAMBDA, dst.lambda.data(), 1);
    r(MRMD_B200_MOL_MODULATED_LAMBDA, dst.modulatedLambda.data(), 1); r(MRMD_B200_MOL_GRAD_LAMBDA, dst.gradLambda.data(), 3);
    r(MRMD_B200_MOL_ATOMS_OFFSET, dst.atomsOffset.data(), 1); r(MRMD_B200_MOL_NUM_ATOMS, dst.numAtoms.data(), 1);
}

/// data::createMoleculeForEachAtom (data/MoleculesFromAtoms.cpp:19-39)
inline Molecules createMoleculeForEachAtom(Atoms& atoms)
{
    atoms.push();
    mrmd_b200_molecules* h = nullptr;
    detail::check(mrmd_b200_molecules_for_each_atom(&h, atoms.handle(), defaultStream), "createMoleculeForEachAtom");
    Molecules m(h);
    m.pull();
    return m;
}

/// data::MultiHistogram (data/MultiHistogram.hpp:29-92): numBins x numHistograms values over [min, max) on the device.
/// `data` stands in for the Kokkos MultiView: data(bin, histogram) reads one entry, toHost() / fromHost() are
/// create_mirror_view_and_copy / deep_copy with a dense row-major host vector.
struct MultiHistogram
{
    struct View
    {
        std::shared_ptr<mrmd_b200_hist> h;
        idx_t numBins = 0, numHistograms = 0;
        std::vector<real_t> toHost() const
        {
            std::vector<real_t> out(static_cast<size_t>(numBins * numHistograms));
            detail::check(mrmd_b200_hist_read(h.get(), out.data(), MRMD_B200_MEM_HOST, defaultStream), "MultiHistogram::data");
            return out;
        }
        void fromHost(const std::vector<real_t>& values) const
        {
            if (idx_c(values.size()) != numBins * numHistograms) detail::fail(MRMD_B200_EINVAL, "MultiHistogram::data: extents differ");
            detail::check(mrmd_b200_hist_write(h.get(), values.data(), MRMD_B200_MEM_HOST, defaultStream), "MultiHistogram::data");
        }
        real_t operator()(idx_t bin, idx_t histogram) const { return toHost()[static_cast<size_t>(bin * numHistograms + histogram)]; }
        idx_t extent(int dim) const { return dim == 0 ? numBins : numHistograms; }
    };

    MultiHistogram(const std::string& label, const real_t minArg, const real_t maxArg, idx_t numBinsArg, idx_t numHistogramsArg)
        : min(minArg), max(maxArg), numBins(numBinsArg), numHistograms(numHistogramsArg),
          binSize((maxArg - minArg) / real_c(numBinsArg)), inverseBinSize(1_r / ((maxArg - minArg) / real_c(numBinsArg))), label_(label)
    {
        mrmd_b200_hist* h = nullptr;
        detail::check(mrmd_b200_hist_create(&h, minArg, maxArg, numBinsArg, numHistogramsArg), "MultiHistogram");
        adopt(h);
    }
    MultiHistogram(const std::string& label, const MultiHistogram& histogram)
        : min(histogram.min), max(histogram.max), numBins(histogram.numBins), numHistograms(histogram.numHistograms),
          binSize(histogram.binSize), inverseBinSize(histogram.inverseBinSize), label_(label)
    {
        mrmd_b200_hist* h = nullptr;
        detail::check(mrmd_b200_hist_clone(&h, histogram.handle(), defaultStream), "MultiHistogram");
        adopt(h);
    }
    /// takes over a handle produced by the C ABI (gradient, smoothen, ThermodynamicForce::getForce)
    MultiHistogram(const std::string& label, mrmd_b200_hist* h) : min(info(h, 0)), max(info(h, 1)), numBins(extent(h, 0)),
          numHistograms(extent(h, 1)), binSize(info(h, 2)), inverseBinSize(info(h, 3)), label_(label)
    {
        adopt(h);
    }

    idx_t getBin(const real_t& val) const { return mrmd_b200_hist_get_bin(handle(), val); }
    real_t getBinPosition(idx_t binIdx) const { return mrmd_b200_hist_get_bin_position(handle(), binIdx); }

    const real_t min;
    const real_t max;
    const idx_t numBins;
    const idx_t numHistograms;
    const real_t binSize;
    const real_t inverseBinSize;
    View data;

    MultiHistogram& operator+=(const MultiHistogram& rhs) { return transform(rhs, 0); }
    MultiHistogram& operator-=(const MultiHistogram& rhs) { return transform(rhs, 1); }
    MultiHistogram& operator*=(const MultiHistogram& rhs) { return transform(rhs, 2); }
    MultiHistogram& operator/=(const MultiHistogram& rhs) { return transform(rhs, 3); }
    void scale(const real_t& scalingFactor) { detail::check(mrmd_b200_hist_scale(handle(), scalingFactor, defaultStream), "MultiHistogram::scale"); }
    void scale(const std::vector<real_t>& scalingFactor)
    {
        detail::check(mrmd_b200_hist_scale_per_histogram(handle(), scalingFactor.data(), idx_c(scalingFactor.size()), defaultStream), "MultiHistogram::scale");
    }
    void makeSymmetric() { detail::check(mrmd_b200_hist_make_symmetric(handle(), defaultStream), "MultiHistogram::makeSymmetric"); }
    mrmd_b200_hist* handle() const { return data.h.get(); }
    const std::string& label() const { return label_; }

private:
    void adopt(mrmd_b200_hist* h)
    {
        data.h.reset(h, [](mrmd_b200_hist* p) { mrmd_b200_hist_destroy(p); });
        data.numBins = numBins;
        data.numHistograms = numHistograms;
    }
    MultiHistogram& transform(const MultiHistogram& rhs, int op)
    {
        detail::check(mrmd_b200_hist_transform(handle(), rhs.handle(), op, defaultStream), "MultiHistogram::transform");
        return *this;
    }
    static real_t info(const mrmd_b200_hist* h, int which)
    {
        double v[4] = {0, 0, 0, 0};
        detail::check(mrmd_b200_hist_info(h, &v[0], &v[1], nullptr, nullptr, &v[2], &v[3]), "MultiHistogram");
        return v[which];
    }
    static idx_t extent(const mrmd_b200_hist* h, int which)
    {
        int64_t n[2] = {0, 0};
        detail::check(mrmd_b200_hist_info(h, nullptr, nullptr, &n[0], &n[1], nullptr, nullptr), "MultiHistogram");
        return n[which];
    }
    std::string label_;
};

/// data/MultiHistogram.cpp:92-111
inline void cumulativeMovingAverage(MultiHistogram& average, const MultiHistogram& current, const real_t movingAverageFactor = 10_r)
{
    detail::check(mrmd_b200_hist_cumulative_moving_average(average.handle(), current.handle(), movingAverageFactor, defaultStream), "cumulativeMovingAverage");
}
/// :113-161
inline MultiHistogram gradient(const MultiHistogram& input, const bool periodic = false)
{
    mrmd_b200_hist* h = nullptr;
    detail::check(mrmd_b200_hist_gradient(&h, input.handle(), periodic, defaultStream), "gradient");
    return MultiHistogram("gradient", h);
}
/// :163-212
inline MultiHistogram smoothen(MultiHistogram& input, const real_t& sigma, const real_t& range, const bool periodic = false)
{
    mrmd_b200_hist* h = nullptr;
    detail::check(mrmd_b200_hist_smoothen(&h, input.handle(), sigma, range, periodic, defaultStream), "smoothen");
    return MultiHistogram("smooth-input", h);
}
/// :214-227 (the ScalarView of bin positions, on the host)
inline std::vector<real_t> createGrid(const MultiHistogram& input)
{
    std::vector<real_t> grid(static_cast<size_t>(input.numBins));
    detail::check(mrmd_b200_hist_create_grid(input.handle(), grid.data(), defaultStream), "createGrid");
    return grid;
}
/// data/MultiHistogram.hpp:176-191 with a parametric one-coordinate predicate (util::IsInSymmetricSlab, or any
/// mrmd_b200_pred such as MRMD_B200_PRED_INTERVAL) in place of the device lambda
inline void replace_if_bin_position(MultiHistogram& hist, const mrmd_b200_pred& pred, real_t newValue)
{
    detail::check(mrmd_b200_hist_replace_if_bin_position(hist.handle(), &pred, newValue, defaultStream), "replace_if_bin_position");
}
inline void replace_if_bin_position(MultiHistogram& hist, const util::IsInSymmetricSlab& pred, real_t newValue)
{
    replace_if_bin_position(hist, pred.desc(), newValue);
}
}  // namespace data

// ------------------------------------------------------------------------------------------------------
/// Cabana::LinkedCellList as used in tests/NVT/NVT.cpp:136-144: binning + permutation run in atoms.permute
class LinkedCellList
{
public:
    LinkedCellList(const data::AtomsPosSlice& pos, idx_t begin, idx_t end, const real_t gridDelta[3], const real_t gridMin[3],
                   const real_t gridMax[3])
        : owner_(pos.owner), begin_(begin), end_(end)
    {
        for (int d = 0; d < 3; ++d)
        {
            delta_[d] = gridDelta[d];
            min_[d] = gridMin[d];
            max_[d] = gridMax[d];
        }
    }
    /// atoms.permute(linkedCellList) (data/Atoms.hpp:95)
    void permute(data::Atoms& atoms) const
    {
        atoms.push();
        detail::check(mrmd_b200_atoms_cell_sort(atoms.handle(), begin_, end_, delta_, min_, max_, nullptr, defaultStream), "permute");
    }

private:
    const data::Atoms* owner_;
    idx_t begin_, end_;
    real_t delta_[3], min_[3], max_[3];
};

namespace detail
{
template <bool HALF>
class VerletListT
{
public:
    VerletListT()
    {
        mrmd_b200_verlet* h = nullptr;
        check(mrmd_b200_verlet_create(&h, HALF ? 1 : 0), "VerletList");
        h_.reset(h, [](mrmd_b200_verlet* p) { mrmd_b200_verlet_destroy(p); });
    }
    VerletListT(const data::AtomsPosSlice& pos, idx_t begin, idx_t end, real_t radius, real_t cellRatio, const real_t gridMin[3],
                const real_t gridMax[3], idx_t maxNeigh = 64)
        : VerletListT()
    {
        build(pos, begin, end, radius, cellRatio, gridMin, gridMax, maxNeigh);
    }
    /// VerletList::build(pos, begin, end, radius, cellRatio, gridMin, gridMax, maxNeigh) (examples/02:156-163)
    void build(const data::AtomsPosSlice& pos, idx_t begin, idx_t end, real_t radius, real_t cellRatio, const real_t gridMin[3],
               const real_t gridMax[3], idx_t maxNeigh = 64)
    {
        pos.owner->push();
        hostValid_ = false;
        check(mrmd_b200_verlet_build_atoms(h_.get(), pos.owner->handle(), begin, end, radius, cellRatio, gridMin, gridMax,
                                           maxNeigh, defaultStream), "VerletList::build");
    }
    void build(const data::MoleculesPosSlice& pos, idx_t begin, idx_t end, real_t radius, real_t cellRatio, const real_t gridMin[3],
               const real_t gridMax[3], idx_t maxNeigh = 64)
    {
        pos.owner->push();
        hostValid_ = false;
        check(mrmd_b200_verlet_build_molecules(h_.get(), pos.owner->handle(), begin, end, radius, cellRatio, gridMin, gridMax,
                                               maxNeigh, defaultStream), "VerletList::build");
    }
    /// B200 fast path (extension): createGhostAtoms + build in one tiled pass, see mrmd_b200_verlet_build_periodic
    void buildPeriodic(const data::Atoms& atoms, const data::Subdomain& subdomain, real_t radius, real_t cellRatio = 1_r, idx_t maxNeigh = 64)
    {
        atoms.push();
        hostValid_ = false;
        const auto s = subdomain.c();
        check(mrmd_b200_verlet_build_periodic(h_.get(), atoms.handle(), &s, radius, cellRatio, maxNeigh, defaultStream), "buildPeriodic");
    }
    /// sum of list._data.counts (tests/LennardJones/LennardJones.cpp:72-80)
    size_t totalPairs() const
    {
        int64_t t = 0;
        check(mrmd_b200_verlet_info(h_.get(), nullptr, nullptr, &t, nullptr), "verlet_info");
        return static_cast<size_t>(t);
    }
    /// host copy of the Cabana VerletLayout2D table: counts[numParticles], neighbors[numParticles][width]
    void toHost(std::vector<int32_t>& counts, std::vector<int32_t>& neighbors, idx_t& width) const
    {
        int64_t n = 0, w = 0;
        check(mrmd_b200_verlet_info(h_.get(), &n, &w, nullptr, nullptr), "verlet_info");
        counts.assign(static_cast<size_t>(n), 0);
        neighbors.assign(static_cast<size_t>(n * w), -1);
        width = w;
        check(mrmd_b200_verlet_read(h_.get(), counts.data(), neighbors.data(), MRMD_B200_MEM_HOST, defaultStream), "verlet_read");
    }
    mrmd_b200_verlet* handle() const { return h_.get(); }
    /// host-side element access behind Cabana::NeighborList<...>::numNeighbor / getNeighbor (a host copy of the table
    /// is fetched once per build; lists from buildPeriodic are read through toHostPeriodic instead)
    idx_t numNeighbor(idx_t particle) const
    {
        fetch();
        return hostCounts_[static_cast<size_t>(particle)];
    }
    idx_t getNeighbor(idx_t particle, idx_t n) const
    {
        fetch();
        return hostNeighbors_[static_cast<size_t>(particle * hostWidth_ + n)];
    }

private:
    void fetch() const
    {
        if (hostValid_) return;
        toHost(hostCounts_, hostNeighbors_, hostWidth_);
        hostValid_ = true;
    }
    std::shared_ptr<mrmd_b200_verlet> h_;
    mutable bool hostValid_ = false;
    mutable std::vector<int32_t> hostCounts_, hostNeighbors_;
    mutable idx_t hostWidth_ = 0;
};
/// Cabana::NeighborList<VerletList>: the static accessors the reference's kernels use (datatypes.hpp:192-193)
template <bool HALF>
struct NeighborListT
{
    static idx_t numNeighbor(const VerletListT<HALF>& list, idx_t particle) { return list.numNeighbor(particle); }
    static idx_t getNeighbor(const VerletListT<HALF>& list, idx_t particle, idx_t n) { return list.getNeighbor(particle, n); }
};
}  // namespace detail
using HalfVerletList = detail::VerletListT<true>;   // datatypes.hpp:184-187
using FullVerletList = detail::VerletListT<false>;  // datatypes.hpp:188-191
using HalfNeighborList = detail::NeighborListT<true>;
using FullNeighborList = detail::NeighborListT<false>;

// ------------------------------------------------------------------------------------------------------
namespace action
{
/// action::LennardJones (action/LennardJones.hpp:98-133)
class LennardJones
{
public:
    LennardJones(const real_t rc, const real_t& sigma, const real_t& epsilon, const real_t& cappingDistance = 0_r)
        : LennardJones({cappingDistance}, {rc}, {sigma}, {epsilon}, 1, false)
    {
    }
    LennardJones(const std::vector<real_t>& cappingDistance, const std::vector<real_t>& rc, const std::vector<real_t>& sigma,
                 const std::vector<real_t>& epsilon, const idx_t& numTypes, const bool isShifted)
    {
        mrmd_b200_lj* h = nullptr;
        detail::check(mrmd_b200_lj_create(&h, cappingDistance.data(), rc.data(), sigma.data(), epsilon.data(), numTypes, isShifted), "LennardJones");
        h_.reset(h, [](mrmd_b200_lj* p) { mrmd_b200_lj_destroy(p); });
    }
    template <class List>
    void apply(data::Atoms& atoms, List& verletList)
    {
        atoms.push();
        detail::check(mrmd_b200_lj_apply(h_.get(), atoms.handle(), verletList.handle(), nullptr, defaultStream), "LennardJones::apply");
    }
    /// apply_if with a parametric two-position predicate (IsInSymmetricSlab::either() / both())
    template <class List>
    void apply_if(const data::Atoms& atoms, const List& verletList, const mrmd_b200_pred& pred)
    {
        atoms.push();
        detail::check(mrmd_b200_lj_apply(h_.get(), atoms.handle(), verletList.handle(), &pred, defaultStream), "LennardJones::apply_if");
    }
    real_t getEnergy() const { real_t e = 0; detail::check(mrmd_b200_lj_get(h_.get(), &e, nullptr, nullptr, defaultStream), "getEnergy"); return e; }
    real_t getVirial() const { real_t v = 0; detail::check(mrmd_b200_lj_get(h_.get(), nullptr, &v, nullptr, defaultStream), "getVirial"); return v; }
    idx_t getNumPairs() const { int64_t p = 0; detail::check(mrmd_b200_lj_get(h_.get(), nullptr, nullptr, &p, defaultStream), "getNumPairs"); return p; }

private:
    std::shared_ptr<mrmd_b200_lj> h_;
};

/// action::VelocityVerlet (action/VelocityVerlet.hpp)
namespace VelocityVerlet
{
inline real_t preForceIntegrate(data::Atoms& atoms, const real_t dt)
{
    atoms.push();
    real_t d = 0;
    detail::check(mrmd_b200_vv_pre(atoms.handle(), dt, &d, defaultStream), "VelocityVerlet::preForceIntegrate");
    return d;
}
inline void postForceIntegrate(data::Atoms& atoms, const real_t dt)
{
    atoms.push();
    detail::check(mrmd_b200_vv_post(atoms.handle(), dt, defaultStream), "VelocityVerlet::postForceIntegrate");
}
}  // namespace VelocityVerlet

/// action::VelocityVerletLangevinThermostat (action/VelocityVerletLangevinThermostat.hpp:29-61)
class VelocityVerletLangevinThermostat
{
public:
    VelocityVerletLangevinThermostat(const real_t& zeta, const real_t& temperature) { set(zeta, temperature); }
    void set(const real_t& zeta, const real_t& temperature)
    {
        zeta_ = zeta;
        temperature_ = temperature;
    }
    real_t preForceIntegrate(data::Atoms& atoms, const real_t dt) { return run(atoms, dt, nullptr); }
    real_t preForceIntegrate_apply_if(data::Atoms& atoms, const real_t dt, const util::IsInSymmetricSlab& pred) { return run(atoms, dt, &pred.desc()); }
    void postForceIntegrate(data::Atoms& atoms, const real_t dt) { VelocityVerlet::postForceIntegrate(atoms, dt); }

private:
    real_t run(data::Atoms& atoms, real_t dt, const mrmd_b200_pred* pred)
    {
        atoms.push();
        real_t d = 0;
        detail::check(mrmd_b200_langevin_pre(atoms.handle(), dt, zeta_, temperature_, seed_, step_++, pred, &d, defaultStream),
                      "VelocityVerletLangevinThermostat::preForceIntegrate");
        return d;
    }
    uint64_t seed_ = 1234;  // pool seed of the reference (:32); keys the Philox stream
    uint64_t step_ = 0;
    real_t zeta_ = 0, temperature_ = 0;
};
}  // namespace action

// ------------------------------------------------------------------------------------------------------
namespace communication
{
/// communication::GhostLayer (communication/GhostLayer.hpp:28-56)
class GhostLayer
{
public:
    GhostLayer()
    {
        mrmd_b200_ghost* h = nullptr;
        detail::check(mrmd_b200_ghost_create(&h), "GhostLayer");
        h_.reset(h, [](mrmd_b200_ghost* p) { mrmd_b200_ghost_destroy(p); });
    }
    void exchangeRealAtoms(data::Atoms& atoms, const data::Subdomain& subdomain)
    {
        atoms.push();
        const auto s = subdomain.c();
        detail::check(mrmd_b200_ghost_map_into_domain(atoms.handle(), &s, defaultStream), "exchangeRealAtoms");
    }
    void createGhostAtoms(data::Atoms& atoms, const data::Subdomain& subdomain)
    {
        atoms.push();
        const auto s = subdomain.c();
        detail::check(mrmd_b200_ghost_create_atoms(h_.get(), atoms.handle(), &s, -1, defaultStream), "createGhostAtoms");
        atoms.pull();
    }
    void updateGhostAtoms(data::Atoms& atoms, const data::Subdomain& subdomain)
    {
        atoms.push();
        const auto s = subdomain.c();
        detail::check(mrmd_b200_ghost_update(h_.get(), atoms.handle(), &s, defaultStream), "updateGhostAtoms");
    }
    void contributeBackGhostToReal(data::Atoms& atoms)
    {
        atoms.push();
        detail::check(mrmd_b200_ghost_contribute_back(h_.get(), atoms.handle(), defaultStream), "contributeBackGhostToReal");
    }

protected:
    std::shared_ptr<mrmd_b200_ghost> h_;
};

/// communication::MultiResGhostLayer (communication/MultiResGhostLayer.hpp:29-61)
class MultiResGhostLayer : public GhostLayer
{
public:
    using GhostLayer::contributeBackGhostToReal;
    using GhostLayer::updateGhostAtoms;
    void exchangeRealAtoms(data::Molecules& molecules, data::Atoms& atoms, const data::Subdomain& subdomain)
    {
        molecules.push();
        atoms.push();
        const auto s = subdomain.c();
        detail::check(mrmd_b200_ghost_mr_map_into_domain(molecules.handle(), atoms.handle(), &s, defaultStream), "exchangeRealAtoms");
    }
    void createGhostAtoms(data::Molecules& molecules, data::Atoms& atoms, const data::Subdomain& subdomain)
    {
        molecules.push();
        atoms.push();
        const auto s = subdomain.c();
        detail::check(mrmd_b200_ghost_mr_create_atoms(h_.get(), molecules.handle(), atoms.handle(), &s, -1, defaultStream), "createGhostAtoms");
        molecules.pull();
        atoms.pull();
    }
};
}  // namespace communication

// ------------------------------------------------------------------------------------------------------
namespace weighting_function
{
/// weighting_function::Slab (weighting_function/Slab.hpp:27-201)
class Slab
{
public:
    enum class InterfaceType { SMOOTH, ABRUPT };
    Slab(const Point3D& center, const real_t atomisticRegionDiameter, const real_t hybridRegionDiameter, const idx_t nu,
         const InterfaceType interfaceType = InterfaceType::SMOOTH)
    {
        w_.kind = MRMD_B200_WEIGHT_SLAB;
        w_.abrupt = interfaceType == InterfaceType::ABRUPT;
        for (int d = 0; d < 3; ++d) w_.center[d] = center[d];
        w_.atRegion = atomisticRegionDiameter;
        w_.hyRegion = hybridRegionDiameter;
        w_.exponent = nu;
    }
    const mrmd_b200_weight& desc() const { return w_; }

private:
    mrmd_b200_weight w_{};
};
/// weighting_function::Spherical (weighting_function/Spherical.hpp:25-99); modulatedLambda := lambda
class Spherical
{
public:
    Spherical(const Point3D& center, const real_t atomisticRadius, const real_t hybridRegionDiameter, const int exponent)
    {
        w_.kind = MRMD_B200_WEIGHT_SPHERICAL;
        for (int d = 0; d < 3; ++d) w_.center[d] = center[d];
        w_.atRegion = atomisticRadius;
        w_.hyRegion = hybridRegionDiameter;
        w_.exponent = exponent;
    }
    const mrmd_b200_weight& desc() const { return w_; }

private:
    mrmd_b200_weight w_{};
};
inline bool isInATRegion(const real_t& lambda) { return lambda >= 1_r; }  // CheckRegion.hpp:27-38
inline bool isInCGRegion(const real_t& lambda) { return lambda <= 0_r; }
inline bool isInHYRegion(const real_t& lambda) { return !isInATRegion(lambda) && !isInCGRegion(lambda); }
}  // namespace weighting_function

namespace action
{
/// action::UpdateMolecules::update (action/UpdateMolecules.hpp:24-70)
namespace UpdateMolecules
{
template <typename WEIGHTING_FUNCTION>
void update(const data::Molecules& molecules, const data::Atoms& atoms, const WEIGHTING_FUNCTION& weight)
{
    molecules.push();
    atoms.push();
    detail::check(mrmd_b200_molecules_update(molecules.handle(), atoms.handle(), &weight.desc(), defaultStream), "UpdateMolecules::update");
}
}  // namespace UpdateMolecules

/// action::ContributeMoleculeForceToAtoms::update (action/ContributeMoleculeForceToAtoms.cpp:23-48)
namespace ContributeMoleculeForceToAtoms
{
inline void update(const data::Molecules& molecules, const data::Atoms& atoms)
{
    molecules.push();
    atoms.push();
    detail::check(mrmd_b200_molecules_contribute_force(molecules.handle(), atoms.handle(), defaultStream), "ContributeMoleculeForceToAtoms::update");
}
}  // namespace ContributeMoleculeForceToAtoms

/// action::LJ_IdealGas (action/LJ_IdealGas.hpp:34-104)
class LJ_IdealGas
{
public:
    LJ_IdealGas(const real_t& cappingDistance, const real_t& rc, const real_t& sigma, const real_t& epsilon, const bool doShift)
        : LJ_IdealGas({cappingDistance}, {rc}, {sigma}, {epsilon}, 1, doShift)
    {
    }
    LJ_IdealGas(const std::vector<real_t>& cappingDistance, const std::vector<real_t>& rc, const std::vector<real_t>& sigma,
                const std::vector<real_t>& epsilon, const idx_t numTypes, const bool doShift)
        : numTypes_(numTypes)
    {
        mrmd_b200_adress* h = nullptr;
        detail::check(mrmd_b200_adress_create(&h, cappingDistance.data(), rc.data(), sigma.data(), epsilon.data(), numTypes, doShift), "LJ_IdealGas");
        h_.reset(h, [](mrmd_b200_adress* p) { mrmd_b200_adress_destroy(p); });
    }
    /// extension: promise that every molecule has this many atoms; 4 selects the four-lanes-per-molecule kernel
    void setAtomsPerMolecule(const idx_t& atomsPerMolecule) { detail::check(mrmd_b200_adress_set_atoms_per_molecule(h_.get(), atomsPerMolecule), "setAtomsPerMolecule"); }
    void setCompensationEnergySamplingInterval(const idx_t& interval) { sampling_ = interval; detail::check(mrmd_b200_adress_set_intervals(h_.get(), sampling_, update_), "interval"); }
    void setCompensationEnergyUpdateInterval(const idx_t& interval) { update_ = interval; detail::check(mrmd_b200_adress_set_intervals(h_.get(), sampling_, update_), "interval"); }
    /// 200 x numTypes doubles, row-major (getMeanCompensationEnergy().data on the host)
    std::vector<real_t> getMeanCompensationEnergy() const
    {
        std::vector<real_t> out(static_cast<size_t>(200 * numTypes_));
        detail::check(mrmd_b200_adress_read_histogram(h_.get(), 0, out.data(), defaultStream), "getMeanCompensationEnergy");
        return out;
    }
    real_t run(data::Molecules& molecules, HalfVerletList& verletList, data::Atoms& atoms)
    {
        molecules.push();
        atoms.push();
        real_t e = 0;
        detail::check(mrmd_b200_adress_run(h_.get(), molecules.handle(), verletList.handle(), atoms.handle(), &e, nullptr, defaultStream), "LJ_IdealGas::run");
        return e;
    }

private:
    std::shared_ptr<mrmd_b200_adress> h_;
    idx_t numTypes_ = 1, sampling_ = 200, update_ = 20000;
};

/// action::ThermodynamicForce (action/ThermodynamicForce.hpp:32-96)
class ThermodynamicForce
{
public:
    ThermodynamicForce(const std::vector<real_t>& targetDensity, const data::Subdomain& subdomain, const real_t& requestedDensityBinWidth,
                       const std::vector<real_t>& thermodynamicForceModulation, const bool enforceSymmetry = false,
                       const bool usePeriodicity = false)
    {
        if (targetDensity.size() != thermodynamicForceModulation.size()) detail::fail(MRMD_B200_EINVAL, "ThermodynamicForce: size mismatch");
        const auto s = subdomain.c();
        mrmd_b200_thermo* h = nullptr;
        detail::check(mrmd_b200_thermo_create(&h, targetDensity.data(), idx_c(targetDensity.size()), &s, requestedDensityBinWidth,
                                              thermodynamicForceModulation.data(), enforceSymmetry, usePeriodicity), "ThermodynamicForce");
        h_.reset(h, [](mrmd_b200_thermo* p) { mrmd_b200_thermo_destroy(p); });
        gridMin_ = subdomain.minCorner[0];
    }
    ThermodynamicForce(const real_t targetDensity, const data::Subdomain& subdomain, const real_t& requestedDensityBinWidth,
                       const real_t thermodynamicForceModulation, const bool enforceSymmetry = false, const bool usePeriodicity = false)
        : ThermodynamicForce(std::vector<real_t>{targetDensity}, subdomain, requestedDensityBinWidth, {thermodynamicForceModulation},
                             enforceSymmetry, usePeriodicity)
    {
    }
    void sample(data::Atoms& atoms) { atoms.push(); detail::check(mrmd_b200_thermo_sample(h_.get(), atoms.handle(), defaultStream), "sample"); }
    void update(const real_t& sigma, const real_t& intensity) { detail::check(mrmd_b200_thermo_update(h_.get(), sigma, intensity, nullptr, defaultStream), "update"); }
    void update_if(const real_t& sigma, const real_t& intensity, const util::IsInSymmetricSlab& pred) { detail::check(mrmd_b200_thermo_update(h_.get(), sigma, intensity, &pred.desc(), defaultStream), "update_if"); }
    void apply(const data::Atoms& atoms) const { atoms.push(); detail::check(mrmd_b200_thermo_apply(h_.get(), atoms.handle(), nullptr, 0, defaultStream), "apply"); }
    void apply_if(const data::Atoms& atoms, const util::IsInSymmetricSlab& pred) const { atoms.push(); detail::check(mrmd_b200_thermo_apply(h_.get(), atoms.handle(), &pred.desc(), 0, defaultStream), "apply_if"); }
    void applyInterpolated_if(const data::Atoms& atoms, const util::IsInSymmetricSlab& pred) const { atoms.push(); detail::check(mrmd_b200_thermo_apply(h_.get(), atoms.handle(), &pred.desc(), 1, defaultStream), "applyInterpolated_if"); }
    idx_t getNumberOfDensityProfileSamples() const { int64_t s = 0; detail::check(mrmd_b200_thermo_info(h_.get(), nullptr, nullptr, nullptr, &s), "info"); return s; }
    idx_t numBins() const { int64_t n = 0; detail::check(mrmd_b200_thermo_info(h_.get(), &n, nullptr, nullptr, nullptr), "info"); return n; }
    idx_t numTypes() const { int64_t n = 0; detail::check(mrmd_b200_thermo_info(h_.get(), nullptr, &n, nullptr, nullptr), "info"); return n; }
    real_t binSize() const { real_t b = 0; detail::check(mrmd_b200_thermo_info(h_.get(), nullptr, nullptr, &b, nullptr), "info"); return b; }
    /// data::createGrid(getForce()) (data/MultiHistogram.hpp): the bin centres min + (i + 1/2) binSize
    std::vector<real_t> createGrid() const
    {
        std::vector<real_t> grid(static_cast<size_t>(numBins()));
        const real_t b = binSize();
        for (size_t i = 0; i < grid.size(); ++i) grid[i] = gridMin_ + (real_c(i) + 0.5_r) * b;
        return grid;
    }
    /// ThermodynamicForce.hpp:57-63: the table / the running density profile as a data::MultiHistogram (a copy)
    data::MultiHistogram getForce() const { return hist(0, "thermodynamic-force"); }
    data::MultiHistogram getDensityProfile() const { return hist(1, "density-profile"); }
    /// getForce(typeId) / getDensityProfile(typeId) (:58, :64): one histogram, on the host
    std::vector<real_t> getForce(const idx_t& typeId) const { return column(read(0), typeId); }
    std::vector<real_t> getDensityProfile(const idx_t& typeId) const { return column(read(1), typeId); }
    void setForce(const std::vector<real_t>& forces) const { detail::check(mrmd_b200_thermo_write_force(h_.get(), forces.data(), defaultStream), "setForce"); }
    std::vector<real_t> getMuLeft() const { return mu().first; }
    std::vector<real_t> getMuRight() const { return mu().second; }

private:
    data::MultiHistogram hist(int kind, const char* label) const
    {
        mrmd_b200_hist* h = nullptr;
        detail::check(mrmd_b200_thermo_get_hist(h_.get(), kind, &h, defaultStream), label);
        return data::MultiHistogram(label, h);
    }
    std::vector<real_t> column(const std::vector<real_t>& all, idx_t typeId) const
    {
        const idx_t nb = numBins(), nt = numTypes();
        std::vector<real_t> out(static_cast<size_t>(nb));
        for (idx_t i = 0; i < nb; ++i) out[static_cast<size_t>(i)] = all[static_cast<size_t>(i * nt + typeId)];
        return out;
    }
    std::vector<real_t> read(int kind) const
    {
        int64_t nb = 0, nt = 0;
        detail::check(mrmd_b200_thermo_info(h_.get(), &nb, &nt, nullptr, nullptr), "info");
        std::vector<real_t> out(static_cast<size_t>(nb * nt));
        detail::check(mrmd_b200_thermo_read(h_.get(), kind, out.data(), defaultStream), "read");
        return out;
    }
    std::pair<std::vector<real_t>, std::vector<real_t>> mu() const
    {
        int64_t nt = 0;
        detail::check(mrmd_b200_thermo_info(h_.get(), nullptr, &nt, nullptr, nullptr), "info");
        std::vector<real_t> l(static_cast<size_t>(nt)), r(static_cast<size_t>(nt));
        detail::check(mrmd_b200_thermo_mu(h_.get(), l.data(), r.data(), defaultStream), "mu");
        return {l, r};
    }
    std::shared_ptr<mrmd_b200_thermo> h_;
    real_t gridMin_ = 0;
};
}  // namespace action

namespace data
{
/// data::Bond (data/Bond.hpp:23-28)
struct Bond
{
    idx_t idx;          ///< relative index of first atom
    idx_t jdx;          ///< relative index of second atom
    real_t eqDistance;  ///< equilibrium distance of the bond
};
using BondView = std::vector<Bond>;  ///< host side stand-in for Kokkos::View<Bond*> (and its host mirror)
}  // namespace data

namespace action
{
/// action::limitAccelerationPerComponent (action/LimitAcceleration.cpp:21-45)
inline void limitAccelerationPerComponent(data::Atoms& atoms, const real_t& maxAccelerationPerComponent)
{
    atoms.push();
    detail::check(mrmd_b200_limit_acceleration(atoms.handle(), maxAccelerationPerComponent, defaultStream), "limitAccelerationPerComponent");
}
/// action::limitVelocityPerComponent (action/LimitVelocity.cpp:23-43)
inline void limitVelocityPerComponent(data::Atoms& atoms, const real_t& maxVelocityPerComponent)
{
    atoms.push();
    detail::check(mrmd_b200_limit_velocity(atoms.handle(), maxVelocityPerComponent, defaultStream), "limitVelocityPerComponent");
}

/// action::BerendsenThermostat::apply (action/BerendsenThermostat.cpp:25-50)
namespace BerendsenThermostat
{
inline void apply(data::Atoms& atoms, const real_t& currentTemperature, const real_t& targetTemperature, const real_t& gamma)
{
    atoms.push();
    detail::check(mrmd_b200_berendsen_thermostat(atoms.handle(), currentTemperature, targetTemperature, gamma, defaultStream),
                  "BerendsenThermostat::apply");
}
}  // namespace BerendsenThermostat

/// action::BerendsenBarostat::apply (action/BerendsenBarostat.cpp:23-50)
namespace BerendsenBarostat
{
inline void apply(data::Atoms& atoms, const real_t& currentPressure, const real_t& targetPressure, const real_t& gamma,
                  data::Subdomain& subdomain, bool stretchX = true, bool stretchY = true, bool stretchZ = true)
{
    atoms.push();
    // the device scales the positions with the same mu the host applies to the subdomain (Subdomain::scaleDim)
    auto s = subdomain.c();
    detail::check(mrmd_b200_berendsen_barostat(atoms.handle(), currentPressure, targetPressure, gamma, &s, stretchX, stretchY,
                                               stretchZ, defaultStream),
                  "BerendsenBarostat::apply");
    const real_t mu = std::cbrt(1_r + gamma * (currentPressure - targetPressure));
    if (stretchX) subdomain.scaleDim(mu, AXIS::X);
    if (stretchY) subdomain.scaleDim(mu, AXIS::Y);
    if (stretchZ) subdomain.scaleDim(mu, AXIS::Z);
}
}  // namespace BerendsenBarostat

/// action::MoleculeConstraints (action/Shake.hpp:159-251): SHAKE / RATTLE over the bonds of every local molecule
class MoleculeConstraints
{
public:
    MoleculeConstraints(idx_t atomsPerMolecule, idx_t numConstraintIterations)
    {
        mrmd_b200_constraints* h = nullptr;
        detail::check(mrmd_b200_constraints_create(&h, atomsPerMolecule, numConstraintIterations), "MoleculeConstraints");
        h_.reset(h, [](mrmd_b200_constraints* p) { mrmd_b200_constraints_destroy(p); });
    }
    void setConstraints(const data::BondView& bonds)
    {
        std::vector<int64_t> idx, jdx;
        std::vector<real_t> eq;
        for (const auto& b : bonds)
        {
            idx.push_back(b.idx);
            jdx.push_back(b.jdx);
            eq.push_back(b.eqDistance);
        }
        detail::check(mrmd_b200_constraints_set(h_.get(), idx.data(), jdx.data(), eq.data(), idx_c(bonds.size())), "setConstraints");
    }
    void enforcePositionalConstraints(data::Molecules& molecules, data::Atoms& atoms, const real_t dt)
    {
        molecules.push();
        atoms.push();
        detail::check(mrmd_b200_constraints_enforce_positional(h_.get(), molecules.handle(), atoms.handle(), dt, defaultStream),
                      "enforcePositionalConstraints");
    }
    void enforceVelocityConstraints(data::Molecules& molecules, data::Atoms& atoms, const real_t dt)
    {
        molecules.push();
        atoms.push();
        detail::check(mrmd_b200_constraints_enforce_velocity(h_.get(), molecules.handle(), atoms.handle(), dt, defaultStream),
                      "enforceVelocityConstraints");
    }

private:
    std::shared_ptr<mrmd_b200_constraints> h_;
};

namespace impl
{
/// action::impl::Coulomb (action/Coulomb.hpp:27-46), evaluated on the device
class Coulomb
{
public:
    real_t computeForce(const real_t& distSqr, const real_t q1, const real_t q2) const { return eval(distSqr, q1, q2, true); }
    real_t computeEnergy(const real_t& distSqr, const real_t q1, const real_t q2) const { return eval(distSqr, q1, q2, false); }

protected:
    real_t eval(const real_t distSqr, const real_t q1, const real_t q2, const bool force) const
    {
        real_t f = 0_r, e = 0_r;
        detail::check(mrmd_b200_coulomb_eval(kind_, rc_, alpha_, &distSqr, 1, q1, q2, &f, &e, defaultStream), "Coulomb");
        return force ? f : e;
    }
    int kind_ = MRMD_B200_COULOMB_PLAIN;
    real_t rc_ = 0_r, alpha_ = 0_r;
};

/// action::impl::CoulombDSF (action/CoulombDSF.hpp:42-84)
class CoulombDSF : public Coulomb
{
public:
    CoulombDSF(const real_t& rc, const real_t& alpha)
    {
        kind_ = MRMD_B200_COULOMB_DSF;
        rc_ = rc;
        alpha_ = alpha;
    }
};
}  // namespace impl

/// action::SPC (action/SPC.hpp:61-362)
class SPC
{
public:
    real_t sumEnergyLJ_ = 0_r;
    real_t sumEnergyCoulomb_ = 0_r;
    auto getEnergyLJ() const { return sumEnergyLJ_; }
    auto getEnergyCoulomb() const { return sumEnergyCoulomb_; }

    static constexpr real_t massO = 15.999_r;
    static constexpr real_t chargeO = -0.82_r;
    static constexpr real_t massH = 1.008_r;
    static constexpr real_t chargeH = +0.41_r;
    static constexpr real_t sigma = 0.31655578901998815_r;
    static constexpr real_t epsilon = 0.6501695808187486_r;
    static constexpr real_t rc = 1.2_r;
    static constexpr real_t alpha = 2.0_r;
    static constexpr real_t eqDistanceHO = 0.1_r;
    static constexpr real_t angleHOH = 109.47_r / 180_r * 3.14159265358979323846;  // util::degToRad(109.47_r)
    const real_t eqDistanceHH = eqDistanceHO * std::sqrt(2_r - 2_r * std::cos(angleHOH));

    /// coulombKind MRMD_B200_COULOMB_PLAIN is the reference's member; MRMD_B200_COULOMB_DSF is an extension
    explicit SPC(int coulombKind = MRMD_B200_COULOMB_PLAIN)
    {
        mrmd_b200_spc* h = nullptr;
        detail::check(mrmd_b200_spc_create(&h, coulombKind), "SPC");
        h_.reset(h, [](mrmd_b200_spc* p) { mrmd_b200_spc_destroy(p); });
    }
    void applyForces(data::Molecules& molecules, HalfVerletList& verletList, data::Atoms& atoms)
    {
        molecules.push();
        atoms.push();
        detail::check(mrmd_b200_spc_apply_forces(h_.get(), molecules.handle(), verletList.handle(), atoms.handle(), &sumEnergyLJ_,
                                                 &sumEnergyCoulomb_, defaultStream), "SPC::applyForces");
    }
    real_t calcBondEnergy(data::Molecules& molecules, data::Atoms& atoms, const real_t& harmonicPreFactor)
    {
        molecules.push();
        atoms.push();
        real_t e = 0_r;
        detail::check(mrmd_b200_spc_calc_bond_energy(h_.get(), molecules.handle(), atoms.handle(), harmonicPreFactor, &e, defaultStream),
                      "SPC::calcBondEnergy");
        return e;
    }
    void enforcePositionalConstraints(data::Molecules& molecules, data::Atoms& atoms, real_t dt)
    {
        molecules.push();
        atoms.push();
        detail::check(mrmd_b200_spc_enforce_positional_constraints(h_.get(), molecules.handle(), atoms.handle(), dt, defaultStream),
                      "SPC::enforcePositionalConstraints");
    }
    void enforceVelocityConstraints(data::Molecules& molecules, data::Atoms& atoms, real_t dt)
    {
        molecules.push();
        atoms.push();
        detail::check(mrmd_b200_spc_enforce_velocity_constraints(h_.get(), molecules.handle(), atoms.handle(), dt, defaultStream),
                      "SPC::enforceVelocityConstraints");
    }

private:
    std::shared_ptr<mrmd_b200_spc> h_;
};
}  // namespace action

// ------------------------------------------------------------------------------------------------------
namespace analysis
{
/// analysis::getKineticEnergy / getMeanKineticEnergy (analysis/KineticEnergy.hpp:26-47)
inline real_t getKineticEnergy(data::Atoms& atoms)
{
    atoms.push();
    real_t e = 0;
    detail::check(mrmd_b200_kinetic_energy(atoms.handle(), &e, defaultStream), "getKineticEnergy");
    return e;
}
inline real_t getMeanKineticEnergy(data::Atoms& atoms) { return getKineticEnergy(atoms) / real_c(atoms.numLocalAtoms); }
/// analysis::getSystemMomentum (analysis/SystemMomentum.cpp:21-50)
inline Vector3D getSystemMomentum(data::Atoms& atoms)
{
    atoms.push();
    Vector3D p{};
    detail::check(mrmd_b200_system_momentum(atoms.handle(), p.data(), defaultStream), "getSystemMomentum");
    return p;
}
/// analysis::getPressure (analysis/Pressure.cpp:23-51)
inline real_t getPressure(data::Atoms& atoms, const data::Subdomain& subdomain)
{
    atoms.push();
    const auto s = subdomain.c();
    real_t p = 0;
    detail::check(mrmd_b200_pressure(atoms.handle(), &s, &p, defaultStream), "getPressure");
    return p;
}
/// analysis::MeanSquareDisplacement (analysis/MeanSquareDisplacement.hpp:24-49)
class MeanSquareDisplacement
{
public:
    MeanSquareDisplacement()
    {
        mrmd_b200_msd* h = nullptr;
        detail::check(mrmd_b200_msd_create(&h), "MeanSquareDisplacement");
        h_.reset(h, [](mrmd_b200_msd* p) { mrmd_b200_msd_destroy(p); });
    }
    void reset(data::Atoms& atoms)
    {
        atoms.push();
        detail::check(mrmd_b200_msd_reset_atoms(h_.get(), atoms.handle(), defaultStream), "MeanSquareDisplacement::reset");
    }
    void reset(data::Molecules& molecules)
    {
        molecules.push();
        detail::check(mrmd_b200_msd_reset_molecules(h_.get(), molecules.handle(), defaultStream), "MeanSquareDisplacement::reset");
    }
    real_t calc(data::Atoms& atoms, const data::Subdomain& subdomain)
    {
        atoms.push();
        const auto s = subdomain.c();
        real_t v = 0;
        detail::check(mrmd_b200_msd_calc_atoms(h_.get(), atoms.handle(), &s, &v, defaultStream), "MeanSquareDisplacement::calc");
        return v;
    }
    real_t calc(data::Molecules& molecules, const data::Subdomain& subdomain)
    {
        molecules.push();
        const auto s = subdomain.c();
        real_t v = 0;
        detail::check(mrmd_b200_msd_calc_molecules(h_.get(), molecules.handle(), &s, &v, defaultStream), "MeanSquareDisplacement::calc");
        return v;
    }

private:
    std::shared_ptr<mrmd_b200_msd> h_;
};
}  // namespace analysis
}  // namespace mrmd

/// stand-ins for the third-party calls the reference's drivers make themselves
namespace Kokkos
{
/// Kokkos::ScopeGuard scope_guard(argc, argv) (examples/02_LennardJones_NVE.cpp:242): nothing to initialise here
struct ScopeGuard
{
    ScopeGuard() = default;
    ScopeGuard(int&, char**) {}
    ScopeGuard(const ScopeGuard&) = delete;
    ScopeGuard& operator=(const ScopeGuard&) = delete;
};
/// Kokkos::fence(): every wrapper call is asynchronous on mrmd::defaultStream
inline void fence() { mrmd::fence(); }
/// Kokkos::Timer (examples/02:117): wall clock since construction / reset
class Timer
{
public:
    Timer() : start_(std::chrono::steady_clock::now()) {}
    double seconds() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - start_).count(); }
    void reset() { start_ = std::chrono::steady_clock::now(); }

private:
    std::chrono::steady_clock::time_point start_;
};
}  // namespace Kokkos

namespace Cabana
{
/// Cabana::deep_copy(force, value) (examples/02_LennardJones_NVE.cpp:174-175)
inline void deep_copy(const mrmd::data::AtomsForceSlice& force, mrmd::real_t value) { force.owner->setForce(value); }
}  // namespace Cabana
