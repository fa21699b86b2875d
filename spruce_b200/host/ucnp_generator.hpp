// ucnp_generator.hpp -- the UCNP problem generator of the reference, host-only (no device library): `.settings` sweeps -> one directory per set of
// conditions with plasma.settings, ucnp.config and init.state, byte for byte what the reference's own generator writes (the files are what
// `run` reads on the other side of the path).
//
// Mirrors   Settings / UCNP                 source/user-interface/settings.{hpp,cpp}, source/ucnp/ucnp_settings.{hpp,cpp}, source/ucnp/ucnputils.cpp
//           ConfigHandler                   source/user-interface/ConfigHandler.{hpp,cpp}
//           StateHandler                    source/user-interface/StateHandler.{hpp,cpp}
//           AntiHelmholtz / CurrentLoop     source/ucnp/antihelmholtz.cpp, source/ucnp/currentloop.cpp   (long double, std::comp_ellint_1/2)
//           Grid::MeshGrid / Gaussian2D / Exp2D, PlasmaDomain::convertCellSizesToCellPositions      source/mhd/grid.cpp:374-408, plasmadomain.cpp:169-192
//
// Parity: tests/test_host_gengrids.py compares every file with the output of the reference's own generator (execs/gengrids.cpp, compiled from the reference's
// sources by the test infrastructure) on sweeps that take every branch.  To be byte-identical the arithmetic keeps the reference's operation order and library calls
// (std::pow where it calls pow, long double where it uses long double), and two quirks are kept on purpose:
//   * the derived plasma characteristics (w_pi, tau, ...) are evaluated ONCE, on the first set of conditions, and appended to every set (ucnp_settings.cpp:78-89
//     reads them through getval(), i.e. through the array chosen before the loop) -- so a `duration = tau = 0.3` is the same in every set;
//   * a `%` comment also removes the character in front of it (settings.cpp:277).
#pragma once
#include <algorithm>
#include <cmath>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "grid.hpp"
#include "utils.hpp"

namespace fs = std::filesystem;

namespace ucnpgen {

// source/constants.hpp:8-19
constexpr double kBoltzmann = 1.3807e-16, kElectronMass = 9.1094e-28, kPi = 3.14159265358979323846, kCharge = 4.80320425e-10;

using StrVec = std::vector<std::string>;
using StrMat = std::vector<StrVec>;

// what an ostream prints for a value at the given precision (Settings::num2str, settings.hpp:51-57)
template <typename T> std::string num2str(T v, int prec = 4)
{
    std::ostringstream ss;
    ss.precision(prec);
    ss << v;
    return ss.str();
}

// ---------------------------------------------------------------------------------------------------------------------------------------------------
// Settings + UCNP: rows `name = unit = v1, v2, ...`; the sets of conditions are the Cartesian product of the value lists (later rows vary fastest), times
// `runs` copies when a `runs` row is present.  A unit is `cgs`, `opt` (a word, not a number) or the name of another cgs row (the value is a multiple of it).
// ---------------------------------------------------------------------------------------------------------------------------------------------------
class Settings {
public:
    explicit Settings(const fs::path &path, bool ucnp = true)
    {
        SPRUCE_REQUIRE(path.extension().string() == ".settings", "Path must point to a file with <.settings> extension.");
        for (const StrVec &row : read_rows(path)) {
            SPRUCE_REQUIRE(row.size() >= 3, "a settings row is <name = unit = value[, value...]>");
            m_names.push_back(row[0]);
            m_units.push_back(row[1]);
            m_vals.emplace_back(row.begin() + 2, row.end());
        }
        SPRUCE_REQUIRE(m_names.size() >= 2, "a .settings file holds at least two rows");
        m_runs_found = is_name("runs");
        if (m_runs_found) {                                              // settings.cpp:43-51
            const StrVec &rv = m_vals[name2ind("runs")];
            SPRUCE_REQUIRE(rv.size() == 1, "Only one value can be given for <runs>");
            m_runs = std::stoi(rv[0]);
            SPRUCE_REQUIRE(m_runs > 0, "<runs> must be an integer greater than zero");
        }
        for (const std::string &a : m_vals[0]) for (const std::string &b : m_vals[1]) m_unique.push_back({a, b});
        for (size_t k = 2; k < m_vals.size(); k++) {                    // settings.cpp:53-54: one more column, the new row's values varying fastest
            StrMat next;
            next.reserve(m_unique.size() * m_vals[k].size());
            for (const StrVec &row : m_unique) for (const std::string &v : m_vals[k]) { next.push_back(row); next.back().push_back(v); }
            m_unique.swap(next);
        }
        if (m_runs_found) {                                              // process_runs, settings.cpp:81-96
            const size_t loc = name2ind("runs");
            StrMat next;
            for (const StrVec &row : m_unique) for (int r = 1; r <= m_runs; r++) { next.push_back(row); next.back()[loc] = num2str(r); }
            m_unique.swap(next);
        }
        m_possible_units = {"cgs", "opt"};
        for (size_t i = 0; i < m_names.size(); i++) if (m_units[i] == "cgs") m_possible_units.push_back(m_names[i]);
        if (ucnp) for (const char *c : kCharacteristics) m_possible_units.push_back(c);
        for (size_t i = 0; i < m_names.size(); i++)
            if (std::find(m_possible_units.begin(), m_possible_units.end(), m_units[i]) == m_possible_units.end()) spruce_die("The units for variable <" + m_names[i] + "> are not valid.");
        choose_array(0);
        if (ucnp) append_characteristics();
        choose_array(0);
    }

    void choose_array(int ind)
    {
        SPRUCE_REQUIRE(ind >= 0 && ind < (int)m_unique.size(), "<array> is out of bounds for <m_unique>");
        m_array = ind;
        m_cur = m_unique[ind];
    }
    int array_size() const { return (int)m_unique.size(); }
    const StrVec &names() const { return m_names; }
    bool is_name(const std::string &s) const { return std::find(m_names.begin(), m_names.end(), s) != m_names.end(); }
    size_t name2ind(const std::string &s) const
    {
        const auto it = std::find(m_names.begin(), m_names.end(), s);
        if (it == m_names.end()) spruce_die("Variable name <" + s + "> not found within <m_names>.");
        return (size_t)(it - m_names.begin());
    }
    // numeric value in cgs: a row whose unit is another row's name is a multiple of that row (settings.cpp:203-220)
    double getval(const std::string &name) const
    {
        const size_t loc = name2ind(name);
        double v = std::stod(m_cur[loc]);
        if (m_units[loc] != "cgs") {
            const size_t base = name2ind(m_units[loc]);
            SPRUCE_REQUIRE(m_units[base] == "cgs", "Variable units can only be expressed in terms of another variable that is expressed in <m_unit_str> units.");
            v *= std::stod(m_cur[base]);
        }
        return v;
    }
    std::string getopt(const std::string &name) const
    {
        const size_t loc = name2ind(name);
        SPRUCE_REQUIRE(m_units[loc] == "opt", "Requested variable is not of type <opt>");
        return m_cur[loc];
    }
    std::string getvar(const std::string &name) const { return m_units[name2ind(name)] == "opt" ? getopt(name) : num2str(getval(name)); }
    // set_<n>[/run_<r>] (settings.cpp:222-238)
    fs::path set_path(int offset) const
    {
        if (!m_runs_found) return fs::path("set_" + num2str(m_array + offset));
        return fs::path("set_" + num2str(m_array / m_runs + offset)) / ("run_" + getvar("runs"));
    }
    // <name>.settings of the chosen set: one `name = unit = value` per row, no newline after the last (settings.cpp:241-251)
    void write_array_params(const fs::path &dir, const std::string &name) const
    {
        fs::create_directories(dir);
        std::ofstream out(dir / (name + ".settings"));
        for (size_t i = 0; i < m_names.size(); i++) out << m_names[i] << " = " << m_units[i] << " = " << m_cur[i] << (i + 1 < m_names.size() ? "\n" : "");
    }

private:
    static constexpr const char *kCharacteristics[9] = {"w_pi", "w_pe", "l_deb", "sig", "tau", "tau_x", "tau_y", "a", "w_pe_inv"};   // ucnp_settings.hpp:19
    StrVec m_names, m_units, m_possible_units, m_cur;
    StrMat m_vals, m_unique;
    bool m_runs_found = false;
    int m_runs = -1, m_array = -1;

    // settings.cpp:258-325: `=` separates like `,`, blanks are dropped, `%` starts a comment, empty lines are skipped
    static StrMat read_rows(const fs::path &path)
    {
        SPRUCE_REQUIRE(fs::exists(path) && !fs::is_empty(path), "File must exist and not be empty.");
        std::ifstream in(path);
        StrMat rows;
        while (in.good()) {
            std::string line;
            std::getline(in, line);
            const size_t c = line.find('%');
            if (c == 0) line.clear();
            else if (c != std::string::npos) line = line.substr(0, c - 1);
            std::replace(line.begin(), line.end(), '=', ',');
            line.erase(std::remove(line.begin(), line.end(), ' '), line.end());
            if (line.empty()) continue;
            StrVec cells;
            std::istringstream ss(line);
            while (ss.good()) { std::string cell; std::getline(ss, cell, ','); cells.push_back(cell); }
            rows.push_back(cells);
        }
        return rows;
    }
    // ucnputils.cpp (phys::) through get_characteristic, ucnp_settings.cpp:58-75
    double characteristic(int which) const
    {
        auto plasma_freq = [](double n, double m) { return std::sqrt(4 * kPi * n * std::pow(kCharge, 2.) / m); };
        auto tau_exp = [](double sig, double m, double T) { return std::sqrt(m * std::pow(sig, 2.) / (kBoltzmann * T)); };
        switch (which) {
        case 0: return plasma_freq(getval("n"), getval("m_i"));
        case 1: return plasma_freq(getval("n"), kElectronMass);
        case 2: return std::sqrt(kBoltzmann * getval("Te") / (4 * kPi * getval("n") * kCharge * kCharge));
        case 3: return std::pow(getval("sig_x") * getval("sig_y"), 1. / 2.);
        case 4: return tau_exp(characteristic(3), getval("m_i"), getval("Te") + getval("Ti"));
        case 5: return tau_exp(getval("sig_x"), getval("m_i"), getval("Te") + getval("Ti"));
        case 6: return tau_exp(getval("sig_y"), getval("m_i"), getval("Te") + getval("Ti"));
        case 7: return std::pow(3. / (4. * kPi * getval("n")), 1. / 3.);
        default: return 1. / plasma_freq(getval("n"), kElectronMass);
        }
    }
    void append_characteristics()
    {
        for (const char *dep : {"n", "m_i", "Te", "Ti", "sig_x", "sig_y"})            // check_for_dependencies, ucnp_settings.cpp:43-55
            if (!is_name(dep)) spruce_die(std::string("Dependency <") + dep + "> is not a possible unit.");
        for (int c = 0; c < 9; c++) {
            m_names.push_back(kCharacteristics[c]);
            m_units.push_back("cgs");
            const std::string v = num2str(characteristic(c));                          // of the set chosen before this loop: the same string for every set
            for (StrVec &row : m_unique) row.push_back(v);
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------------------------------------------
// ConfigHandler: the template .config with the lines of every settings row that is also a config key rewritten
// ---------------------------------------------------------------------------------------------------------------------------------------------------
class ConfigHandler {
public:
    explicit ConfigHandler(const fs::path &path)
    {
        SPRUCE_REQUIRE(path.extension().string() == ".config", "the config template must have extension .config");
        SPRUCE_REQUIRE(fs::exists(path) && !fs::is_empty(path), "File must exist and not be empty.");
        std::ifstream in(path);
        while (in.good()) { std::string line; std::getline(in, line); m_lines.push_back(line); }      // (a final newline leaves one empty entry, as in the reference)
        for (const std::string &line : m_lines) {
            std::string lhs, rhs;
            parse(line, lhs, rhs);
            if (rhs == "true" && std::find(kSets.begin(), kSets.end(), lhs) != kSets.end()) m_eqs_name = lhs;
        }
        SPRUCE_REQUIRE(!m_eqs_name.empty(), "Active equation set not found in .config file.");
        // PlasmaDomain::m_config_names (plasmadomain.hpp:46-51), then the config_names() of the four modules ConfigHandler.hpp:32 lists, then the equation set's
        m_keys = {"x_bound_1", "x_bound_2", "y_bound_1", "y_bound_2", "epsilon", "density_min", "temp_min", "thermal_energy_min", "max_iterations", "iter_output_interval",
                  "time_output_interval", "output_flags", "xdim", "ydim", "open_boundary_strength", "std_out_interval", "write_interval", "open_boundary_decay_base", "x_origin",
                  "y_origin", "time_integrator", "duration", "sg_opt", "write_precision", "multispecies_mode",
                  "timescale", "lengthscale", "strength", "gt_species", "gt_strength", "gt_use_diffusion", "gt_use_global_temp",
                  "visc_output_to_file", "visc_strength", "visc_vars_diff", "visc_vars_update"};
        if (m_eqs_name == "ideal_2F") for (const char *k : {"use_sub_cycling", "epsilon_courant", "smooth_fields", "viscosity"}) m_keys.push_back(k);      // ideal2F.hpp:21-22
    }
    const std::string &eqs_set_name() const { return m_eqs_name; }
    bool is_config(const std::string &name) const { return std::find(m_keys.begin(), m_keys.end(), name) != m_keys.end(); }
    void update_config(const std::string &name, const std::string &val)
    {
        SPRUCE_REQUIRE(is_config(name), "not a config name");
        bool found = false;
        for (std::string &line : m_lines) {
            std::string lhs, rhs;
            parse(line, lhs, rhs);
            if (lhs == name) { line = lhs + " = " + val; found = true; }
        }
        if (!found) m_lines.push_back(name + " = " + val);
    }
    void write_config_file(const fs::path &dir) const
    {
        fs::create_directories(dir);
        std::ofstream out(dir / "ucnp.config");
        for (const std::string &line : m_lines) out << line << "\n";
    }

private:
    static inline const StrVec kSets = {"ideal_mhd", "ideal_mhd_cons", "ideal_mhd_2E", "ideal_2F"};       // equationset.hpp:22
    StrVec m_lines, m_keys;
    std::string m_eqs_name;
    static void parse(const std::string &line, std::string &lhs, std::string &rhs)      // ConfigHandler.cpp:69-76: `lhs = rhs # comment`, blanks dropped
    {
        std::istringstream ss(line);
        lhs.clear(); rhs.clear();
        std::getline(ss, lhs, '=');
        std::getline(ss, rhs, '#');
        lhs.erase(std::remove(lhs.begin(), lhs.end(), ' '), lhs.end());
        rhs.erase(std::remove(rhs.begin(), rhs.end(), ' '), rhs.end());
    }
};

// ---------------------------------------------------------------------------------------------------------------------------------------------------
// The quadrupole field of an anti-Helmholtz pair (two coaxial current loops with opposite currents), scaled to a given gradient on the axis.
// Field of one loop: Simpson et al., "Simple analytic expressions for the magnetic field of a circular current loop" (the formulas the reference cites),
// in long double with the library's complete elliptic integrals; operation order as in currentloop.cpp:25-64 so that the rounded doubles agree.
// ---------------------------------------------------------------------------------------------------------------------------------------------------
struct CurrentLoop {
    long double radius = 1, scale = 1, offset = 0;      // loop radius; current / pi; position of the loop's plane along the symmetry axis
    int axis = 0, perm[3] = {1, 2, 0};                 // symmetry axis; the input components that play x, y, z of the loop's own frame
    CurrentLoop() = default;
    CurrentLoop(long double a, long double current, int ax, long double pos) : radius(a), scale(current / kPi), offset(pos), axis(ax)
    {
        const int p[3][3] = {{1, 2, 0}, {2, 0, 1}, {0, 1, 2}};
        for (int k = 0; k < 3; k++) perm[k] = p[ax][k];
    }
    void field(const long double *pos_in, long double *B) const
    {
        long double pos[3] = {pos_in[0], pos_in[1], pos_in[2]};
        pos[axis] -= offset;
        const long double x = pos[perm[0]], y = pos[perm[1]], z = pos[perm[2]], a = radius;
        const long double rho = std::sqrt(x * x + y * y);
        const long double r = std::sqrt(x * x + y * y + z * z);
        const long double alpha = std::sqrt(a * a + r * r - 2. * a * rho);
        const long double beta = std::sqrt(a * a + r * r + 2. * a * rho);
        const long double k = std::sqrt(1. - std::pow(alpha, 2.) / std::pow(beta, 2.));
        const long double K = std::comp_ellint_1(k), E = std::comp_ellint_2(k);
        const long double fac = scale / (2. * alpha * alpha * beta);
        B[perm[0]] = fac * x * z / (rho * rho) * ((a * a + r * r) * E - alpha * alpha * K);
        B[perm[1]] = fac * y * z / (rho * rho) * ((a * a + r * r) * E - alpha * alpha * K);
        B[perm[2]] = fac * ((a * a - r * r) * E + alpha * alpha * K);
        if (rho / a < 1e-5) {                           // on the axis the expressions above are 0/0: near-axis form
            B[perm[0]] = 3 * kPi * a * a * x * z / (4 * (a * a + z * z));
            B[perm[1]] = 3 * kPi * a * a * y * z / (4 * (a * a + z * z));
        }
    }
};
struct AntiHelmholtz {
    CurrentLoop right, left;
    long double conv = 1.;
    AntiHelmholtz(long double a, long double sep, long double gradient, int ax) : right(a, -1, ax, sep / 2.), left(a, 1, ax, -sep / 2.)
    {
        long double unit[3] = {0, 0, 0}, B[3];
        unit[ax] = 1;
        field(unit, B);                                  // with conv = 1: the raw field one unit from the centre, on the axis
        conv = std::abs(gradient / B[ax]);
    }
    void field(const long double *pos, long double *B) const
    {
        long double b1[3], b2[3];
        right.field(pos, b1);
        left.field(pos, b2);
        for (int k = 0; k < 3; k++) B[k] = (b1[k] + b2[k]) * conv;
    }
};

// ---------------------------------------------------------------------------------------------------------------------------------------------------
// StateHandler: the grids of init.state for the chosen set of conditions
// ---------------------------------------------------------------------------------------------------------------------------------------------------
class StateHandler {
public:
    explicit StateHandler(const std::string &eqs_name) : m_eqs(eqs_name)
    {
        m_names = {"d_x", "d_y", "pos_x", "pos_y", "be_x", "be_y", "be_z"};                                                        // PlasmaDomain::m_gridnames
        StrVec sv;                                                                                                                  // state_variables() of the set
        if (eqs_name == "ideal_mhd" || eqs_name == "ideal_mhd_cons") sv = {"rho", "temp", "mom_x", "mom_y", "mom_z", "bi_x", "bi_y", "bi_z", "grav_x", "grav_y"};
        else if (eqs_name == "ideal_mhd_2E") sv = {"rho", "i_temp", "e_temp", "mom_x", "mom_y", "bi_x", "bi_y", "grav_x", "grav_y"};
        else if (eqs_name == "ideal_2F") sv = {"i_rho", "e_rho", "i_mom_x", "i_mom_y", "e_mom_x", "e_mom_y", "i_temp", "e_temp", "bi_x", "bi_y", "bi_z", "E_x", "E_y", "E_z", "grav_x", "grav_y"};
        else spruce_die("Equation set not recognized.");
        m_names.insert(m_names.end(), sv.begin(), sv.end());
        m_grids.resize(m_names.size());
        m_set.assign(m_names.size(), false);
    }

    void setup(const Settings &s)
    {
        m_xdim = s.getval("Nx"); m_ydim = s.getval("Ny"); m_ion_mass = s.getval("m_i"); m_gamma = s.getval("adiabatic_index");
        const int nx = (int)m_xdim, ny = (int)m_ydim;
        std::vector<double> dxv, dyv;
        const std::string grid_opt = s.getopt("grid_opt");
        if (grid_opt == "uniform") {
            dxv.assign(nx, 2 * s.getval("x_lim") / m_xdim);
            dyv.assign(ny, 2 * s.getval("y_lim") / m_ydim);
        } else if (grid_opt == "non-uniform") {
            dxv = non_uniform_spacing(nx, s.getval("x_lim"), s.getval("grid_growth"), s.getval("grid_spread"));
            dyv = non_uniform_spacing(ny, s.getval("y_lim"), s.getval("grid_growth"), s.getval("grid_spread"));
        } else spruce_die("<grid_opt> must be <uniform> or <non-uniform>.");
        Grid d_x(nx, ny), d_y(nx, ny);
        for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) { d_x(i, j) = dxv[i]; d_y(i, j) = dyv[j]; }
        const Grid pos_x = centred_positions(d_x, 0), pos_y = centred_positions(d_y, 1);
        set("d_x", d_x); set("d_y", d_y); set("pos_x", pos_x); set("pos_y", pos_y);

        const AntiHelmholtz quad(30., 120., s.getval("dBdx"), 0);                     // StateHandler.cpp:133: radius 30, separation 120, symmetry axis x
        Grid B[3] = {Grid(nx, ny), Grid(nx, ny), Grid(nx, ny)};
        for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) {
            const long double p[3] = {pos_x(i, j), pos_y(i, j), 0.};
            long double b[3];
            quad.field(p, b);
            for (int c = 0; c < 3; c++) B[c](i, j) = (double)b[c];
        }
        set("be_x", B[0]); set("be_y", B[1]); set("be_z", B[2]);

        const Grid n = density(s, pos_x, pos_y);
        const Grid zeros(nx, ny, 0.0);
        auto scaled = [&](double f) { Grid g(nx, ny); for (int k = 0; k < g.size(); k++) g.data()[k] = n.data()[k] * f; return g; };
        if (m_eqs == "ideal_2F") {
            set("i_rho", scaled(m_ion_mass)); set("e_rho", scaled(kElectronMass));
            set("i_temp", Grid(nx, ny, s.getval("Ti"))); set("e_temp", Grid(nx, ny, s.getval("Te")));
        } else {
            set("rho", scaled(m_ion_mass));
            if (m_eqs == "ideal_mhd_2E") { set("e_temp", Grid(nx, ny, s.getval("Te"))); set("i_temp", Grid(nx, ny, s.getval("Ti"))); }
            else set("temp", Grid(nx, ny, s.getval("Te")));
        }
        for (size_t g = 0; g < m_names.size(); g++) if (!m_set[g]) { m_grids[g] = zeros; m_set[g] = true; }      // momenta, induced field, E, gravity
    }

    // StateHandler.cpp:82-100: the header scalars at stream precision 6, the grids at 16 significant digits
    void write_state_file(const fs::path &dir) const
    {
        fs::create_directories(dir);
        std::ofstream out(dir / "init.state");
        out << "xdim,ydim\n" << m_xdim << "," << m_ydim << "\nion_mass\n" << m_ion_mass << "\nadiabatic_index\n" << m_gamma << "\nt=" << 0. << "\n";
        for (size_t g = 0; g < m_names.size(); g++) out << m_names[g] << "\n" << m_grids[g].format(',', '\n', -1);
    }

private:
    std::string m_eqs;
    StrVec m_names;
    std::vector<Grid> m_grids;
    std::vector<bool> m_set;
    double m_xdim = 0, m_ydim = 0, m_ion_mass = 0, m_gamma = 0;

    void set(const std::string &name, const Grid &g)
    {
        const auto it = std::find(m_names.begin(), m_names.end(), name);
        if (it == m_names.end()) spruce_die("<" + name + "> is not a valid grid name.");
        m_grids[it - m_names.begin()] = g;
        m_set[it - m_names.begin()] = true;
    }
    // StateHandler.cpp:265-297: cell sizes grow geometrically away from a central cell, dr(i) = 1 + A (dr(i-1) - B), normalised so that the outermost centre sits at r_lim
    static std::vector<double> non_uniform_spacing(int N, double r_lim, double A, double B)
    {
        SPRUCE_REQUIRE(A > 1, "The growth factor must be greater than one.");
        SPRUCE_REQUIRE(B < 1 && B > 0, "The spread factor must be positive and less than one.");
        SPRUCE_REQUIRE(N % 2 != 0, "Number of grids must be odd for this function because it assumes a central grid cell exists.");
        std::vector<double> dr(N, 0.0), r(N, 0.0);
        dr[N / 2] = 1.;
        for (int i = N / 2 + 1; i < N; i++) {
            dr[i] = 1. + A * (dr[i - 1] - B);
            r[i] = r[i - 1] + (dr[i - 1] + dr[i]) / 2.;
        }
        const double norm = *std::max_element(r.begin(), r.end());
        for (int i = 0; i < N; i++) { r[i] /= norm; dr[i] /= norm; }
        for (int i = 0; i < N; i++) { r[i] *= r_lim; dr[i] *= r_lim; }
        for (int i = 0; i < N / 2; i++) dr[i] = dr[N - 1 - i];
        return dr;
    }
    // plasmadomain.cpp:169-192 with origin "center": running sum of half sizes along the axis, then shifted by half the extent
    static Grid centred_positions(const Grid &d, int axis)
    {
        const int nx = d.rows(), ny = d.cols();
        Grid pos(nx, ny);
        for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) {
            const int ip = i - (1 - axis), jp = j - axis;
            if ((axis == 0 && i == 0) || (axis == 1 && j == 0)) pos(i, j) = 0.5 * d(i, j);
            else pos(i, j) = pos(ip, jp) + 0.5 * d(ip, jp) + 0.5 * d(i, j);
        }
        const double shift = 0.5 * (pos(nx - 1, ny - 1) + 0.5 * d(nx - 1, ny - 1));
        for (double &p : pos.data()) p -= shift;
        return pos;
    }
    // grid.cpp:386-408
    static Grid gaussian2d(const Grid &x, const Grid &y, double amp, double floor, double sx, double sy)
    {
        Grid g(x.rows(), x.cols());
        for (int i = 0; i < g.rows(); i++) for (int j = 0; j < g.cols(); j++)
            g(i, j) = floor + amp * std::exp(-0.5 * std::pow((x(i, j) - 0.) / std::abs(sx), 2)) * std::exp(-0.5 * std::pow((y(i, j) - 0.) / std::abs(sy), 2));
        return g;
    }
    static Grid exp2d(const Grid &x, const Grid &y, double amp, double floor, double sx, double sy)
    {
        const double eta = sy / sx;
        Grid g(x.rows(), x.cols());
        for (int i = 0; i < g.rows(); i++) for (int j = 0; j < g.cols(); j++)
            g(i, j) = floor + amp * std::exp(-std::sqrt(std::pow(x(i, j) - 0., 2.) + std::pow(y(i, j) - 0., 2.) / eta) / sx);
        return g;
    }
    // setup_density, StateHandler.cpp:207-263: base profile, optional ion hole, optional localised wave train ("n_shock"), optional ion-acoustic modulation ("n_iaw")
    static Grid density(const Settings &s, const Grid &x, const Grid &y)
    {
        const double n_max = s.getval("n"), n_min = s.getval("n_min");
        const std::string dist = s.getopt("n_dist");
        Grid n(x.rows(), x.cols(), 0.0);
        if (dist == "gaussian") n = gaussian2d(x, y, n_max, n_min, s.getval("sig_x"), s.getval("sig_y"));
        else if (dist == "exponential") n = exp2d(x, y, n_max, n_min, s.getval("sig_x"), s.getval("sig_y"));
        else if (dist == "uniform") n = Grid((size_t)s.getval("Nx"), (size_t)s.getval("Ny"), n_max);
        else spruce_die("Density distribution options are: <gaussian>, <exponential>, or <uniform>.");
        const int cells = n.size();
        if (s.getopt("n_hole") == "true") {
            const Grid hole = gaussian2d(x, y, s.getval("n_hole_amp"), 0, s.getval("n_hole_size"), s.getval("sig_y"));
            for (int k = 0; k < cells; k++) n.data()[k] -= hole.data()[k];
        }
        if (s.getopt("n_shock") == "true") {
            const double amp = s.getval("n_shock_amp"), lam = s.getval("n_shock_lam"), sg = s.getval("n_shock_sig"), sx = s.getval("sig_x"), sy = s.getval("sig_y");
            const double kx = 2 * kPi / lam;
            const Grid env = gaussian2d(x, y, 1, 0, sg, sy / sx * sg);
            for (int k = 0; k < cells; k++) n.data()[k] += (std::cos(x.data()[k] * kx) * amp) * env.data()[k];
            std::cout << *std::max_element(n.data().begin(), n.data().end()) << std::endl;      // the reference prints the peak density here (:239)
        }
        if (s.getopt("n_iaw") == "true") {
            const double amp = s.getval("n_iaw_amp"), phase = s.getval("n_iaw_phase") * kPi / 180, kx = 2 * kPi / s.getval("n_iaw_sig");
            for (int k = 0; k < cells; k++) n.data()[k] += (n.data()[k] * amp) * std::sin(x.data()[k] * kx - phase);
        }
        return n;
    }
};

}  // namespace ucnpgen
