// plasmadomain.cpp -- construction, file I/O and the run loop of the B200-backed PlasmaDomain.
// Reference behaviour followed: constructor sequence plasmadomain.cpp:13-75, iteration bounds :138-161, readers
// fileio.cpp:14-122, writers fileio.cpp:125-255, run loop evolution.cpp:8-57.
#include "plasmadomain.hpp"
#include "utils.hpp"
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

void PlasmaDomain::check(int rc)
{
    if (rc != SPRUCE_OK) spruce_die(std::string("spruce_b200: ") + spruce_last_error());
}

PlasmaDomain::PlasmaDomain(const fs::path &out_path, const fs::path &config_path, const fs::path &state_file, bool continue_mode, bool overwrite_init)
    : m_overwrite_init(overwrite_init), m_continue_mode(continue_mode), m_out_directory(out_path), m_module_handler(*this)
{
    for (auto &g : m_grids) g = Grid::Zero(1, 1);
    // the .config travels with the run: copy it into the output directory unless it already lives there
    const fs::path new_config_path = m_out_directory / config_path.filename();
    if (rank() != 0) {
        // files are rank 0's business
    } else if (!fs::exists(new_config_path) || !fs::equivalent(config_path, new_config_path)) {
        std::cout << "Copying " << config_path.string() << " into " << m_out_directory.string() << "...\n";
        if (fs::exists(new_config_path)) fs::remove(new_config_path);
        fs::copy(config_path, new_config_path, fs::copy_options::overwrite_existing);
    } else {
        std::cout << config_path.string() << " already located in output directory.\n";
    }
    std::cout << "Reading config file...\n";
    readConfigFile(config_path);
    SPRUCE_REQUIRE(m_eqs.get() != nullptr, "an equation set must be activated in the .config file");
    std::cout << "Reading state file...\n";
    readStateFile(state_file, continue_mode);
    std::cout << "Validating input data...\n";
    for (const Grid &g : m_grids) SPRUCE_REQUIRE(g.size() != 1, "All internal grid quantities for PlasmaDomain must be initialized");
    computeIterationBounds();
    SlabComm::partition((int)m_xdim, nRanks(), rank(), m_row0, m_nx_local);
    SPRUCE_REQUIRE(m_nx_local >= 4, "every slab needs at least 4 rows: fewer ranks for this grid");
    SPRUCE_REQUIRE(nRanks() == 1 || !m_module_handler.hasHostModules(), "host-resident modules (sg_filtering, tracer_particles, coulomb_explosion, global_temperature) work on whole planes: one rank only");
    m_eqs->setupEquationSet();
    if (m_multispecies_mode) check(spruce_multispecies_mode(m_dev, 1));       // before the modules: they hand their ms_electron_heating_fraction over in setupModule
    m_module_handler.setupModules();
    if (!continue_mode && m_overwrite_init) {
        std::cout << "Writing out init.state...\n";
        writeStateFile("init");
    }
    if (!continue_mode && m_write_interval > 0) {
        outputPreamble();
        storeGrids();
        writeToOutFile();
    }
}

PlasmaDomain::~PlasmaDomain()
{
    if (m_dev) spruce_domain_destroy(m_dev);
}

// plasmadomain.cpp:138-161
void PlasmaDomain::computeIterationBounds()
{
    SPRUCE_REQUIRE(m_xdim > 4 && m_ydim > 4, "Grid too small for ghost zones");
    m_xl = x_bound_1 == BoundaryCondition::Periodic ? 0 : 2;
    m_xu = x_bound_2 == BoundaryCondition::Periodic ? (int)m_xdim - 1 : (int)m_xdim - 3;
    m_yl = y_bound_1 == BoundaryCondition::Periodic ? 0 : 2;
    m_yu = y_bound_2 == BoundaryCondition::Periodic ? (int)m_ydim - 1 : (int)m_ydim - 3;
    m_ghost_zone_mask = Grid(m_xdim, m_ydim, 0.0);
    for (int i = m_xl; i <= m_xu; i++) for (int j = m_yl; j <= m_yu; j++) m_ghost_zone_mask(i, j) = 1.0;
}

void PlasmaDomain::createDevice()
{
    if (m_dev) return;
    spruce_config c;
    std::memset(&c, 0, sizeof(c));
    c.abi_version = SPRUCE_ABI_VERSION;
    c.equation_set = m_eqs->device_id();
    c.xdim = (int)m_xdim; c.ydim = (int)m_ydim;
    c.x_bound_1 = (int)x_bound_1; c.x_bound_2 = (int)x_bound_2; c.y_bound_1 = (int)y_bound_1; c.y_bound_2 = (int)y_bound_2;   // same enum order as SPRUCE_BC_*
    c.time_integrator = (int)m_time_integrator;
    c.device = nRanks() > 1 ? rank() : -1;                            // one rank per GPU, in device order
    c.row0 = m_row0; c.nx_local = m_nx_local; c.rank = rank(); c.n_ranks = nRanks();
    c.ion_mass = m_ion_mass; c.adiabatic_index = m_adiabatic_index; c.epsilon = epsilon;
    c.density_min = density_min; c.temp_min = temp_min; c.thermal_energy_min = thermal_energy_min;
    c.open_boundary_strength = open_boundary_strength; c.open_boundary_decay_base = open_boundary_decay_base; c.time = m_time;
    check(spruce_domain_create(&c, &m_dev));
    // the reference keeps d_x, d_y as planes but requires d_x = d_x(i), d_y = d_y(j) (README.md:41)
    std::vector<double> dx(m_xdim), dy(m_ydim);
    for (size_t i = 0; i < m_xdim; i++) dx[i] = m_grids[d_x](i, 0);
    for (size_t j = 0; j < m_ydim; j++) dy[j] = m_grids[d_y](0, j);
    for (size_t i = 0; i < m_xdim; i++) for (size_t j = 0; j < m_ydim; j++)
        SPRUCE_REQUIRE(m_grids[d_x](i, j) == dx[i] && m_grids[d_y](i, j) == dy[j], "d_x must vary with i only and d_y with j only (rectilinear grid)");
    check(spruce_set_cell_sizes(m_dev, dx.data(), dx.size(), dy.data(), dy.size()));
    for (int g : {be_x, be_y, be_z}) check(spruce_grid_upload(m_dev, m_gridnames[g].c_str(), slab(m_grids[g]), slabCount()));
}

// fileio.cpp:14-80
void PlasmaDomain::readStateFile(const fs::path &state_file, bool continue_mode)
{
    std::ifstream in(state_file.string());
    SPRUCE_REQUIRE(in.good(), "cannot open state file " + state_file.string());
    std::string line, element;
    std::getline(in, line);
    while (in.good() && (line.empty() || line[0] == '#')) {
        if (!line.empty()) m_comment_lines.push_back(line);
        std::getline(in, line);
    }
    clearWhitespace(line);
    SPRUCE_REQUIRE(line == "xdim,ydim", "state file must start with xdim,ydim");
    getCleanedLine(in, line);
    { std::istringstream ss(line); std::getline(ss, element, ','); m_xdim = std::stoi(element); std::getline(ss, element, ','); m_ydim = std::stoi(element); }
    getCleanedLine(in, line); SPRUCE_REQUIRE(line == "ion_mass", "expected ion_mass");
    getCleanedLine(in, line); m_ion_mass = std::stod(line);
    getCleanedLine(in, line); SPRUCE_REQUIRE(line == "adiabatic_index", "expected adiabatic_index");
    getCleanedLine(in, line); m_adiabatic_index = std::stod(line);
    getCleanedLine(in, line);
    { std::istringstream ss(line); std::getline(ss, element, '='); SPRUCE_REQUIRE(element == "t", "expected t=..."); std::getline(ss, element); }
    m_time = continue_mode ? std::stod(element) : 0.0;      // fileio.cpp:50-51
    while (getCleanedLine(in, line)) {
        if (line.empty()) continue;
        const std::string var_name = line;
        Grid g(m_xdim, m_ydim);
        std::vector<std::string> rows(m_xdim);
        for (size_t i = 0; i < m_xdim; i++) getCleanedLine(in, rows[i]);
        std::vector<size_t> counts(m_xdim, 0);
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)m_xdim; i++)                       // rows parse independently (std::from_chars)
            counts[(size_t)i] = parseDelimitedRow(rows[(size_t)i].data(), rows[(size_t)i].data() + rows[(size_t)i].size(), g.ptr() + (size_t)i * m_ydim, m_ydim);
        for (size_t i = 0; i < m_xdim; i++) {
            SPRUCE_REQUIRE(counts[i] != (size_t)-1, "Encountered non-numerical row in .state file sooner than expected");
            SPRUCE_REQUIRE(counts[i] <= m_ydim, "Row in .state file is too long (greater than ydim)");
        }
        auto it = std::find(m_gridnames.begin(), m_gridnames.end(), var_name);
        if (it != m_gridnames.end()) m_grids[it - m_gridnames.begin()] = g;
        else m_eqs->hostGrid(m_eqs->name2index(var_name)) = g;
    }
}

// fileio.cpp:84-122
void PlasmaDomain::readConfigFile(const fs::path &config_file)
{
    std::ifstream in(config_file.string());
    SPRUCE_REQUIRE(in.good(), "cannot open config file " + config_file.string());
    std::string line, lhs, rhs;
    while (std::getline(in, line)) {
        clearWhitespace(line);
        if (line.empty() || line[0] == '#') continue;
        splitAssignment(line, lhs, rhs);
        if (EquationSet::isEquationSetName(lhs)) { EquationSet::instantiateWithConfig(m_eqs, *this, in, lhs, rhs == "true"); continue; }
        if (m_module_handler.isModuleName(lhs)) {
            SPRUCE_REQUIRE(m_eqs.get() != nullptr, "<m_eqs> must be instantiated before instantiating modules");
            m_module_handler.instantiateModule(lhs, in, rhs == "true");
            continue;
        }
        const std::vector<std::string> rhs_vec = splitString(rhs, ',');
        auto it = std::find(m_config_names.begin(), m_config_names.end(), lhs);
        if (it == m_config_names.end()) spruce_die(lhs + " is not a valid config name.");
        const int index = (int)(it - m_config_names.begin());
        if (rhs_vec.size() == 1) handleSingleConfig(index, rhs);
        else if (rhs_vec.size() > 1) {
            SPRUCE_REQUIRE(m_config_names[index] == "output_flags", "Only output_flags support list specifier");
            SPRUCE_REQUIRE(m_eqs.get() != nullptr, "equation_set must be defined in config file before output_flags");
            for (const std::string &v : rhs_vec) m_eqs->setOutputFlag(v, true);
        }
    }
}

// fileio.cpp:297-325
void PlasmaDomain::handleSingleConfig(int i, const std::string &rhs)
{
    const std::string &k = m_config_names[i];
    if (k == "x_bound_1") x_bound_1 = stringToBoundaryCondition(rhs);
    else if (k == "x_bound_2") x_bound_2 = stringToBoundaryCondition(rhs);
    else if (k == "y_bound_1") y_bound_1 = stringToBoundaryCondition(rhs);
    else if (k == "y_bound_2") y_bound_2 = stringToBoundaryCondition(rhs);
    else if (k == "epsilon") epsilon = std::stod(rhs);
    else if (k == "density_min") density_min = std::stod(rhs);
    else if (k == "temp_min") temp_min = std::stod(rhs);
    else if (k == "thermal_energy_min") thermal_energy_min = std::stod(rhs);
    else if (k == "max_iterations") max_iterations = std::stoi(rhs);
    else if (k == "iter_output_interval") m_iter_output_interval = std::stoi(rhs);
    else if (k == "time_output_interval") m_time_output_interval = std::stod(rhs);
    else if (k == "output_flags") { SPRUCE_REQUIRE(m_eqs.get() != nullptr, "equation_set must be defined in config file before output_flags"); m_eqs->setOutputFlag(rhs, true); }
    else if (k == "xdim") m_xdim = std::stoi(rhs);
    else if (k == "ydim") m_ydim = std::stoi(rhs);
    else if (k == "open_boundary_strength") open_boundary_strength = std::stod(rhs);
    else if (k == "write_interval") m_write_interval = std::stoi(rhs);
    else if (k == "std_out_interval") m_std_out_interval = std::stoi(rhs);
    else if (k == "open_boundary_decay_base") open_boundary_decay_base = std::stod(rhs);
    else if (k == "time_integrator") m_time_integrator = stringToTimeIntegrator(rhs);
    else if (k == "duration") m_duration = std::stod(rhs);
    else if (k == "write_precision") m_write_precision = std::stoi(rhs);
    else if (k == "multispecies_mode") m_multispecies_mode = (rhs == "true");      // cumulative_electron / ion / joule_heating planes in mhd.out (fileio.cpp:164-183), kept on the device
    else if (k == "sg_opt") m_sg_opt = rhs;                 // read by the sg_filtering module only (plasmadomain.cpp:51), which is not ported and refuses itself
    else if (k == "x_origin" || k == "y_origin") {}         // accepted and unused, as in the reference (fileio.cpp: parsed, never read on the run path)
}

PlasmaDomain::BoundaryCondition PlasmaDomain::stringToBoundaryCondition(const std::string &str) const
{
    auto it = std::find(m_boundary_condition_names.begin(), m_boundary_condition_names.end(), str);
    SPRUCE_REQUIRE(it != m_boundary_condition_names.end(), "unknown boundary condition <" + str + ">");
    return static_cast<BoundaryCondition>(it - m_boundary_condition_names.begin());
}
PlasmaDomain::TimeIntegrator PlasmaDomain::stringToTimeIntegrator(const std::string &str) const
{
    auto it = std::find(m_time_integrator_names.begin(), m_time_integrator_names.end(), str);
    SPRUCE_REQUIRE(it != m_time_integrator_names.end(), "unknown time integrator <" + str + ">");
    return static_cast<TimeIntegrator>(it - m_time_integrator_names.begin());
}

// plasmadomain.hpp:248-256: precision(-1) leaves the stream at its default of 6 significant digits
std::string PlasmaDomain::num2str(double num)
{
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%.6g", num);
    return buf;
}

// fileio.cpp:125-139
void PlasmaDomain::outputPreamble()
{
    if (rank() != 0) return;
    std::ofstream out(m_out_directory / m_out_filename);
    for (const std::string &c : m_comment_lines) out << c << std::endl;
    out << "xdim,ydim" << std::endl << m_xdim << "," << m_ydim << std::endl;
    for (int v : {pos_x, pos_y, be_x, be_y, be_z}) out << m_gridnames[v] << std::endl << m_grids[v].format(',', '\n');
}

// fileio.cpp:144-200
void PlasmaDomain::storeGrids()
{
    // on slabs every rank takes part in the gathers (grid(i), the modules' planes); only rank 0 formats and keeps the text
    const bool writer = rank() == 0;
    if (writer) m_data_to_write.push_back("t=" + num2str(m_time) + '\n');
    for (int i = 0; i < m_eqs->num_variables(); i++) {
        if (!m_eqs->getOutputFlag(i)) continue;
        const Grid &g = m_eqs->grid(i);
        if (!writer) continue;
        m_data_to_write.push_back(m_eqs->index2name(i) + '\n');
        m_data_to_write.push_back(g.format(',', '\n', m_write_precision));
    }
    if (m_multispecies_mode) {                                                  // fileio.cpp:164-183: between the equation set's variables and the modules' planes
        for (const char *name : {"cumulative_electron_heating", "cumulative_ion_heating", "cumulative_joule_heating"}) {
            Grid g(m_xdim, m_ydim);
            check(spruce_module_output(m_dev, name, slab(g), slabCount()));
            gatherRows(g);
            if (writer) { m_data_to_write.push_back(std::string(name) + '\n'); m_data_to_write.push_back(g.format(',', '\n', m_write_precision)); }
        }
    }
    std::vector<std::string> names; std::vector<Grid> grids;
    m_module_handler.getFileOutputData(names, grids);
    for (size_t i = 0; writer && i < names.size(); i++) { m_data_to_write.push_back(names[i] + '\n'); m_data_to_write.push_back(grids[i].format(',', '\n', m_write_precision)); }
    m_store_counter++;
}

// fileio.cpp:205-215
void PlasmaDomain::writeToOutFile()
{
    m_store_counter = 0;
    if (rank() != 0) return;
    std::ofstream out(m_out_directory / m_out_filename, std::ofstream::app);
    for (const std::string &s : m_data_to_write) out << s;
    m_data_to_write.clear();
    m_store_counter = 0;
}

// fileio.cpp:221-255
void PlasmaDomain::writeStateFile(const std::string &stem, int precision)
{
    if (nRanks() > 1) {                                               // gather first (collective), then rank 0 writes from its staging copies
        for (int i : m_eqs->state_variables()) m_eqs->grid(i);
        if (rank() != 0) return;
    }
    const bool gathered = nRanks() > 1;
    const fs::path filename = stem == "mhd" ? "mhd" + std::to_string(m_state_identifier) + ".state" : stem + ".state";
    std::ofstream f(m_out_directory / filename);
    for (const std::string &c : m_comment_lines) f << c << std::endl;
    f << "xdim,ydim\n" << m_xdim << "," << m_ydim << std::endl;
    f << "ion_mass\n" << m_ion_mass << std::endl;
    f << "adiabatic_index\n" << m_adiabatic_index << std::endl;
    f << "t=" << m_time << std::endl;
    for (size_t i = 0; i < m_gridnames.size(); i++) f << m_gridnames[i] << std::endl << m_grids[i].format(',', '\n', precision);
    for (int i : m_eqs->state_variables()) f << m_eqs->index2name(i) << std::endl << (gathered ? m_eqs->hostGrid(i) : m_eqs->grid(i)).format(',', '\n', precision);
}

void PlasmaDomain::updateStateIdentifier() { m_state_identifier = m_state_identifier == 1 ? 2 : 1; }

// fileio.cpp:283-295
void PlasmaDomain::printUpdate(int iter, double time, double dt) const
{
    std::cout << "Iter: " << iter;
    if (max_iterations > 0) std::cout << "/" << max_iterations;
    std::cout << "|t: " << time;
    if (m_max_time > 0.0) std::cout << "/" << m_max_time;
    std::cout << "|dt: " << dt;
    for (const std::string &m : m_module_handler.getCommandLineMessages()) std::cout << "|" << m;
    std::cout << std::endl;
}

// evolution.cpp:8-57.  advanceTime() calls are batched on the device up to the next iteration at which the host has to
// look (output, stdout line with module messages, time-based output, wall-clock limit).
void PlasmaDomain::run(double time_duration, double cluster_time)
{
    if (time_duration > 0.0) { m_duration = time_duration; m_max_time = m_time + m_duration; }
    else { SPRUCE_REQUIRE(m_duration >= 0.0, "Duration must be specified on command line, or in state file"); m_max_time = m_duration; }
    std::cout << "Begining simulation...\n";
    const bool per_step_messages = m_std_out_interval > 0 && !m_module_handler.getCommandLineMessages().empty();
    // host-resident modules (sg_filtering): one step per device call, their hooks around it (evolution.cpp:64-66, 74).  Exact for post-iterate hooks;
    // a host pre / iterate hook that edits the state makes the device recompute the step size, where the reference keeps the one it took before.
    const bool host_hooks = m_module_handler.hasHostModules();
    std::vector<double> dts;
    while (m_time < m_max_time && (max_iterations < 0 || m_iter < max_iterations)) {
        int n = max_iterations > 0 ? max_iterations - m_iter : 1024;
        if (m_iter_output_interval > 0) n = std::min(n, m_iter_output_interval - (m_iter % m_iter_output_interval));
        if (m_time_output_interval > 0.0 || per_step_messages || host_hooks) n = 1;
        if (cluster_time > 0) n = std::min(n, 16);
        dts.assign(n, 0.0);
        int done = 0;
        if (host_hooks) { const double dt = m_eqs->nextStepSize(); m_module_handler.preIterateModules(dt); m_module_handler.iterateModules(dt); }
        check(spruce_advance(m_dev, n, m_max_time, dts.data(), &done));
        SPRUCE_REQUIRE(done > 0, "the device made no progress");
        if (host_hooks) m_module_handler.postIterateModules(dts[0]);                    // m_iter still names the step just integrated (evolution.cpp:74 before :81)
        bool store_time = false;
        for (int s = 0; s < done; s++) {
            const int old_time_iter = (int)(m_time / m_time_output_interval);
            if (m_std_out_interval > 0 && m_iter % m_std_out_interval == 0) printUpdate(m_iter, m_time, dts[s]);
            m_time += dts[s];                                                           // evolution.cpp:80-81
            m_iter++;
            store_time = m_time_output_interval > 0.0 && (int)(m_time / m_time_output_interval) > old_time_iter;
        }
        const bool store_iter = m_iter_output_interval > 0 && m_iter % m_iter_output_interval == 0;
        if (store_iter || store_time) {
            storeGrids();
            if (m_multispecies_mode) check(spruce_multispecies_reset(m_dev));          // cumulative quantities restart between outputs (evolution.cpp:36-41)
        }
        if (m_write_interval > 0 && m_store_counter > 0 && m_store_counter % m_write_interval == 0) {
            writeToOutFile();
            updateStateIdentifier();
            writeStateFile("end");
        }
        if (cluster_time > 0 && SlabComm::instance().broadcast0(elapsed() > cluster_time ? 1 : 0)) break;      // rank 0's clock decides for every rank
    }
    writeToOutFile();
    writeStateFile("end");
    if (nRanks() > 1) {
        SlabComm::instance().barrier();
        SlabComm::instance().markFinished();
        if (rank() != 0) { spruce_domain_destroy(m_dev); m_dev = nullptr; std::_Exit(0); }       // rank 0 alone reports the end of the run
    }
    if (m_time >= m_max_time || (max_iterations > 0 && m_iter >= max_iterations)) {
        // the reference ends a completed run with assert(false) so that wrapper scripts stop (evolution.cpp:54-56): same exit status 134
        spruce_die("Simulation successfully reached max simulation time or iterations. Printing this error to end recursive scripts.");
    }
}
