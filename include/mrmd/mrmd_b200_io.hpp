// mrmd_b200_io.hpp -- the on-disk formats at the edges of the hot loop, host side only:
// GRO restore / dump, whitespace "x y z" restore, thermodynamic-force profile dump / restore.  Same free functions and
// file formats as mrmd::io, so states and force profiles written by the reference can be continued on the B200 path
// and vice versa.  Reference: mrmd/io/RestoreGRO.cpp:24-147, DumpGRO.cpp:26-89, RestoreTXT.cpp:24-74,
// DumpThermoForce.cpp:24-65, DumpProfile.cpp:22-49, RestoreThermoForce.cpp:24-98.  (H5MD needs HDF5: out of scope.)
#pragma once

#include <cmath>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "mrmd_b200.hpp"

namespace mrmd::io
{
/// io::restoreGRO (io/RestoreGRO.cpp:24-147): "%5d%5s%5s%5d%8lf%8lf%8lf[%8lf%8lf%8lf]" atom lines, last line = box;
/// mass 1, relativeMass 1, type 0, force 0; subdomain = [minCorner, minCorner + box] with the given ghost layer
inline void restoreGRO(const std::string& filename, data::Subdomain& subdomain, data::Atoms& atoms, bool containsGhostAtoms = false)
{
    std::ifstream fin(filename);
    if (!fin.is_open())
    {
        std::cerr << "Could not open file: " << filename << std::endl;
        std::exit(EXIT_FAILURE);
    }
    if (containsGhostAtoms)
    {
        std::cout << "Reading a file with ghost atoms is not supported." << std::endl;
        std::exit(EXIT_FAILURE);
    }
    char buf[1024];
    fin.getline(buf, 1024);  // comment line
    fin.getline(buf, 1024);
    int numAtomsInt = 0;
    std::sscanf(buf, "%d", &numAtomsInt);
    const idx_t numAtoms = idx_c(numAtomsInt);
    if (numAtoms <= 0)
    {
        std::cerr << "Invalid GRO atom count in file " << filename << std::endl;
        std::exit(EXIT_FAILURE);
    }
    data::HostAtoms h_atoms(numAtoms);
    auto h_pos = h_atoms.getPos();
    auto h_vel = h_atoms.getVel();
    bool warningNoVelocities = false;
    idx_t idx = 0;
    while (idx < numAtoms && !fin.eof())
    {
        int tmpInt;
        char tmpChar[6];
        double f[6];
        fin.getline(buf, 1024);
        const int parsed = std::sscanf(buf, "%5d%5s%5s%5d%8lf%8lf%8lf%8lf%8lf%8lf", &tmpInt, tmpChar, tmpChar, &tmpInt, &f[0], &f[1],
                                       &f[2], &f[3], &f[4], &f[5]);
        if (parsed != 10 && parsed != 7)
        {
            std::cerr << "Invalid GRO atom line in file " << filename << ": expected 7 fields for coordinates (or 10 fields including "
                      << "velocities), but parsed " << parsed << " fields from line: \"" << buf << "\"" << std::endl;
            std::exit(EXIT_FAILURE);
        }
        if (parsed == 7) warningNoVelocities = true;
        for (int d = 0; d < 3; ++d)
        {
            h_pos(idx, d) = real_c(f[d]);
            h_vel(idx, d) = (parsed == 10) ? real_c(f[3 + d]) : 0_r;
        }
        h_atoms.getMass()(idx) = 1_r;
        h_atoms.getRelativeMass()(idx) = 1_r;
        ++idx;
    }
    Vector3D diameter{};
    fin >> diameter[0] >> diameter[1] >> diameter[2];
    if (warningNoVelocities)
        std::cout << "Warning: Some lines in file " << filename
                  << " do not contain velocities. Respective velocities have been set to zero, but this may lead to unexpected behavior."
                  << std::endl;
    subdomain = data::Subdomain(subdomain.minCorner,
                                {subdomain.minCorner[0] + diameter[0], subdomain.minCorner[1] + diameter[1],
                                 subdomain.minCorner[2] + diameter[2]},
                                subdomain.ghostLayerThickness);
    if (idx != numAtoms)
    {
        std::cerr << "GRO file " << filename << " ends after " << idx << " of " << numAtoms << " atoms" << std::endl;
        std::exit(EXIT_FAILURE);
    }
    h_atoms.numLocalAtoms = idx;
    h_atoms.numGhostAtoms = 0;
    data::deep_copy(atoms, h_atoms);
}

/// io::dumpGRO (io/DumpGRO.cpp:26-89)
inline void dumpGRO(const std::string& filename, data::Atoms& atoms, const data::Subdomain& subdomain, const real_t& timestamp,
                    const std::string& title, const std::string& resName, const std::vector<std::string>& typeNames,
                    bool dumpGhosts = true, bool dumpVelocities = false)
{
    data::HostAtoms h_atoms(0);
    data::deep_copy(h_atoms, atoms);
    auto pos = h_atoms.getPos();
    auto vel = h_atoms.getVel();
    auto type = h_atoms.getType();
    std::ofstream fout(filename);
    if (!fout.is_open())
    {
        std::cerr << "Could not open file: " << filename << std::endl;
        std::exit(EXIT_FAILURE);
    }
    const idx_t lastAtomIdx = atoms.numLocalAtoms + (dumpGhosts ? atoms.numGhostAtoms : 0);
    fout << title << ", t=" << timestamp << std::endl;
    fout << lastAtomIdx << std::endl;
    for (idx_t idx = 0; idx < lastAtomIdx; ++idx)
    {
        const auto& typeName = typeNames[static_cast<size_t>(type(idx))];
        char buf[1024];
        if (!dumpVelocities)
            std::snprintf(buf, sizeof(buf), "%5d%-5s%5s%5d%8.3f%8.3f%8.3f", static_cast<int>(idx + 1), resName.c_str(),
                          typeName.c_str(), static_cast<int>(idx + 1), pos(idx, 0), pos(idx, 1), pos(idx, 2));
        else
            std::snprintf(buf, sizeof(buf), "%5d%-5s%5s%5d%8.3f%8.3f%8.3f%8.4f%8.4f%8.4f", static_cast<int>(idx + 1), resName.c_str(),
                          typeName.c_str(), static_cast<int>(idx + 1), pos(idx, 0), pos(idx, 1), pos(idx, 2), vel(idx, 0),
                          vel(idx, 1), vel(idx, 2));
        fout << std::string(buf) << std::endl;
    }
    fout << "    " << subdomain.diameter[0] << " " << subdomain.diameter[1] << " " << subdomain.diameter[2] << std::endl;
}

/// io::dumpCSV (io/DumpCSV.cpp:26-52): "idx, mol, type, ghost, pos_x, ..., vel_z" with mol = idx / 3 as the reference writes it
inline void dumpCSV(const std::string& filename, data::Atoms& atoms, bool dumpGhosts = true)
{
    data::HostAtoms at(0);
    data::deep_copy(at, atoms);
    auto pos = at.getPos();
    auto vel = at.getVel();
    auto type = at.getType();
    std::ofstream fout(filename);
    if (!fout.is_open())
    {
        std::cerr << "Could not open file: " << filename << std::endl;
        std::exit(EXIT_FAILURE);
    }
    fout << "idx, mol, type, ghost, pos_x, pos_y, pos_z, vel_x, vel_y, vel_z" << std::endl;
    const idx_t lastAtomIdx = atoms.numLocalAtoms + (dumpGhosts ? atoms.numGhostAtoms : 0);
    for (idx_t idx = 0; idx < lastAtomIdx; ++idx)
    {
        fout << idx << ", " << idx / 3 << ", " << type(idx) << ", " << ((idx < atoms.numLocalAtoms) ? 0 : 1) << ", " << pos(idx, 0)
             << ", " << pos(idx, 1) << ", " << pos(idx, 2) << ", " << vel(idx, 0) << ", " << vel(idx, 1) << ", " << vel(idx, 2)
             << std::endl;
    }
}

/// io::restoreAtoms (io/RestoreTXT.cpp:24-74): whitespace separated "x y z" triples, mass 1, type 0
inline data::Atoms restoreAtoms(const std::string& filename)
{
    std::ifstream fin(filename);
    if (!fin.is_open())
    {
        std::cerr << "Could not open file: " << filename << std::endl;
        std::exit(EXIT_FAILURE);
    }
    std::vector<real_t> xyz;
    while (!fin.eof())
    {
        double x, y, z;
        fin >> x >> y >> z;
        if (fin.eof() || fin.fail()) break;
        if (std::isnan(x) || std::isnan(y) || std::isnan(z))
        {
            std::cout << "invalid position: " << x << " " << y << " " << z << std::endl;
            std::exit(EXIT_FAILURE);
        }
        xyz.insert(xyz.end(), {x, y, z});
    }
    const idx_t n = idx_c(xyz.size() / 3);
    data::HostAtoms h_atoms(n);
    h_atoms.pos = xyz;
    for (idx_t i = 0; i < n; ++i) h_atoms.getMass()(i) = 1_r;
    h_atoms.numLocalAtoms = n;
    data::Atoms atoms(n);
    data::deep_copy(atoms, h_atoms);
    return atoms;
}

namespace detail
{
/// DumpProfile::dumpScalarView (io/DumpProfile.cpp:26-35): one line, single blanks, default ostream precision
inline void dumpLine(std::ofstream& f, const std::vector<real_t>& values, real_t normalizationFactor = 1_r)
{
    for (size_t idx = 0; idx < values.size(); ++idx) f << values[idx] * normalizationFactor << ((idx + 1 < values.size()) ? " " : "");
    f << std::endl;
}
inline std::vector<real_t> forceOfType(const action::ThermodynamicForce& tf, idx_t typeId)
{
    return tf.getForce(typeId);
}
}  // namespace detail

/// io::dumpThermoForce for one type (io/DumpThermoForce.cpp:24-39): grid line, force line
inline void dumpThermoForce(const std::string& filename, const action::ThermodynamicForce& thermodynamicForce, const idx_t& typeId)
{
    std::ofstream f(filename);
    detail::dumpLine(f, thermodynamicForce.createGrid());
    detail::dumpLine(f, detail::forceOfType(thermodynamicForce, typeId));
}
/// io::dumpThermoForce for all types (:41-63): grid line, one force line per type
inline void dumpThermoForce(const std::string& filename, const action::ThermodynamicForce& thermodynamicForce)
{
    std::ofstream f(filename);
    detail::dumpLine(f, thermodynamicForce.createGrid());
    for (idx_t typeId = 0; typeId < thermodynamicForce.numTypes(); ++typeId)
        detail::dumpLine(f, detail::forceOfType(thermodynamicForce, typeId));
}

/// io::restoreThermoForce (io/RestoreThermoForce.cpp:24-98): bin width = difference of the first two grid values
inline action::ThermodynamicForce restoreThermoForce(const std::string& filename, const data::Subdomain& subdomain,
                                                     const std::vector<real_t>& targetDensities = {1_r},
                                                     const std::vector<real_t>& thermodynamicForceModulations = {1_r},
                                                     const bool enforceSymmetry = false, const bool usePeriodicity = false,
                                                     const idx_t maxNumForces = 10)
{
    std::ifstream file(filename);
    std::string line, word;
    std::getline(file, line);
    std::vector<real_t> grid;
    {
        std::stringstream s(line);
        while (s >> word) grid.push_back(std::stod(word));
    }
    if (grid.size() < 2) mrmd::detail::fail(MRMD_B200_EINVAL, "restoreThermoForce: fewer than two grid points");
    const real_t binWidth = grid[1] - grid[0];
    std::vector<std::vector<real_t>> rows;
    while (std::getline(file, line))
    {
        std::stringstream s(line);
        std::vector<real_t> row;
        while (s >> word) row.push_back(std::stod(word));
        rows.push_back(row);
        if (idx_c(rows.size()) > maxNumForces) mrmd::detail::fail(MRMD_B200_EINVAL, "restoreThermoForce: more force lines than maxNumForces");
    }
    action::ThermodynamicForce thermodynamicForce(targetDensities, subdomain, binWidth, thermodynamicForceModulations, enforceSymmetry,
                                                  usePeriodicity);
    const idx_t nb = thermodynamicForce.numBins(), nt = thermodynamicForce.numTypes();
    if (idx_c(rows.size()) != nt || idx_c(grid.size()) != nb)
        mrmd::detail::fail(MRMD_B200_EINVAL, "restoreThermoForce: the file does not match the force table (bins x types)");
    std::vector<real_t> forces(static_cast<size_t>(nb * nt));
    for (idx_t t = 0; t < nt; ++t)
        for (idx_t i = 0; i < nb; ++i) forces[static_cast<size_t>(i * nt + t)] = rows[static_cast<size_t>(t)][static_cast<size_t>(i)];
    thermodynamicForce.setForce(forces);
    return thermodynamicForce;
}
}  // namespace mrmd::io
