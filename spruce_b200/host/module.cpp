// module.cpp -- Module base, ModuleHandler and the three solar modules that run on the device.
// Config-block parsing follows module.cpp:11-35 / modulehandler.cpp:77-112 of the reference; the parsed values are
// handed to libspruce_b200.so in setupModule(), in config order (= execution order).
#include "module.hpp"
#include "plasmadomain.hpp"
#include "utils.hpp"
#include <cmath>
#include <iostream>

void Module::configureModule(std::ifstream &in)
{
    std::vector<std::string> lhs_all, rhs_all;
    std::string line, lhs, rhs;
    std::getline(in, line);
    SPRUCE_REQUIRE(!line.empty() && line[0] == '{', "All Module activation configs must be immediately followed by curly brackets (on their own lines) to enclose Module configs");
    while (std::getline(in, line)) {
        if (!line.empty() && line[0] == '}') break;
        clearWhitespace(line);
        if (line.empty() || line[0] == '#') continue;
        splitAssignment(line, lhs, rhs);
        lhs_all.push_back(lhs); rhs_all.push_back(rhs);
    }
    parseModuleConfigs(lhs_all, rhs_all);
}

void ModuleHandler::setupModules() { for (auto &m : m_modules) m->setupModule(); }
bool ModuleHandler::isModuleName(const std::string &name) const { return std::find(m_module_names.begin(), m_module_names.end(), name) != m_module_names.end(); }

void ModuleHandler::instantiateModule(const std::string &name, std::ifstream &in, bool active)
{
    if (!active) {   // fast-forward over the block
        std::string line;
        std::getline(in, line); clearWhitespace(line);
        SPRUCE_REQUIRE(!line.empty() && line[0] == '{', "All Modules activation/deactivation configs must be immediately followed by curly brackets");
        do { if (!std::getline(in, line)) break; clearWhitespace(line); } while (line.empty() || line[0] != '}');
        return;
    }
    SPRUCE_REQUIRE(dynamic_cast<IdealMHD *>(m_pd.m_eqs.get()) != nullptr, "Module designed for IdealMHD EquationSet (ensure that equation_set is set before modules in the config)");
    if (name == "thermal_conduction") m_modules.emplace_back(new ThermalConduction(m_pd));
    else if (name == "radiative_losses") m_modules.emplace_back(new RadiativeLosses(m_pd));
    else if (name == "ambient_heating") m_modules.emplace_back(new AmbientHeating(m_pd));
    else spruce_die("Module <" + name + "> is not ported to the B200 path yet (thermal_conduction, radiative_losses, ambient_heating are).");
    m_modules.back()->configureModule(in);
}

std::vector<std::string> ModuleHandler::getCommandLineMessages() const
{
    std::vector<std::string> out;
    for (auto &m : m_modules) { const std::string s = m->commandLineMessage(); if (!s.empty()) out.push_back(s); }
    return out;
}
void ModuleHandler::getFileOutputData(std::vector<std::string> &names, std::vector<Grid> &grids) const { for (auto &m : m_modules) m->fileOutput(names, grids); }

static int integrator_id(std::string s, const char *who)
{
    if (s.empty()) s = "euler";
    if (s == "euler") return SPRUCE_TI_EULER;
    if (s == "rk2") return SPRUCE_TI_RK2;
    if (s == "rk4") return SPRUCE_TI_RK4;
    spruce_die(std::string("Invalid time integrator given for ") + who + " module");
}
static void no_file_output(bool flag, const char *who)
{
    if (flag) std::cerr << who << ": output_to_file is not provided by the B200 path yet; the run proceeds without the module's diagnostic planes.\n";
}

// thermalconduction.cpp:16-30
void ThermalConduction::parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs)
{
    for (size_t i = 0; i < lhs.size(); i++) {
        const std::string &k = lhs[i], &v = rhs[i];
        if (k == "flux_saturation") flux_saturation = (v == "true");
        else if (k == "epsilon") epsilon = std::stod(v);
        else if (k == "dt_subcycle_min") dt_subcycle_min = std::stod(v);
        else if (k == "output_to_file") output_to_file = (v == "true");
        else if (k == "time_integrator") time_integrator = v;
        else if (k == "inactive_mode") inactive_mode = (v == "true");
        else if (k == "weakening_factor") weakening_factor = std::stod(v);
        else if (k == "ms_electron_heating_fraction") { }
        else std::cerr << k << " config not recognized for Thermal Conduction Module.\n";
    }
}
void ThermalConduction::setupModule()
{
    SPRUCE_REQUIRE(!inactive_mode, "thermal_conduction inactive_mode is a diagnostic of the CPU build");
    no_file_output(output_to_file, "thermal_conduction");
    PlasmaDomain::check(spruce_module_thermal_conduction(m_pd.device(), flux_saturation, integrator_id(time_integrator, "Thermal Conduction"), epsilon, dt_subcycle_min, weakening_factor));
}
std::string ThermalConduction::commandLineMessage() const
{
    int n = 0;
    spruce_module_subcycles(m_pd.device(), "thermal_conduction", &n);
    return "Thermal Subcycles: " + std::to_string(n);
}

// radiativelosses.cpp:17-31
void RadiativeLosses::parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs)
{
    for (size_t i = 0; i < lhs.size(); i++) {
        const std::string &k = lhs[i], &v = rhs[i];
        if (k == "cutoff_ramp") cutoff_ramp = std::stod(v);
        else if (k == "cutoff_temp") cutoff_temp = std::stod(v);
        else if (k == "epsilon") epsilon = std::stod(v);
        else if (k == "output_to_file") output_to_file = (v == "true");
        else if (k == "time_integrator") time_integrator = v;
        else if (k == "inactive_mode") inactive_mode = (v == "true");
        else if (k == "prevent_subcycling") prevent_subcycling = (v == "true");
        else if (k == "ms_electron_heating_fraction") { }
        else std::cerr << k << " config not recognized.\n";
    }
}
void RadiativeLosses::setupModule()
{
    SPRUCE_REQUIRE(!inactive_mode, "radiative_losses inactive_mode is a diagnostic of the CPU build");
    no_file_output(output_to_file, "radiative_losses");
    PlasmaDomain::check(spruce_module_radiative_losses(m_pd.device(), integrator_id(time_integrator, "Radiative Losses"), cutoff_ramp, cutoff_temp, epsilon, prevent_subcycling));
}
std::string RadiativeLosses::commandLineMessage() const
{
    int n = 0;
    spruce_module_subcycles(m_pd.device(), "radiative_losses", &n);
    return "Radiative Subcycles: " + std::to_string(n);
}

// ambientheating.cpp:11-40: the static heating plane is built on the host with the host libm, once, as in the reference
void AmbientHeating::parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs)
{
    for (size_t i = 0; i < lhs.size(); i++) {
        const std::string &k = lhs[i], &v = rhs[i];
        if (k == "heating_rate") heating_rate = std::stod(v);
        else if (k == "exp_mode") exp_mode = (v == "true");
        else if (k == "exp_base_heating_rate") exp_base_heating_rate = std::stod(v);
        else if (k == "exp_scale_height") exp_scale_height = std::stod(v);
        else if (k == "split_exp_mode") split_exp_mode = (v == "true");
        else if (k == "split_exp_scale_height") split_exp_scale_height = std::stod(v);
        else if (k == "split_exp_start_height") split_exp_start_height = std::stod(v);
        else if (k == "ms_electron_heating_fraction") { }
        else std::cerr << k << " config not recognized.\n";
    }
}
void AmbientHeating::setupModule()
{
    const size_t nx = m_pd.xdim(), ny = m_pd.ydim();
    const Grid &mask = m_pd.ghostZoneMask(), &pos_y = m_pd.m_grids[PlasmaDomain::pos_y];
    Grid heating(nx, ny);
    for (size_t i = 0; i < nx; i++) for (size_t j = 0; j < ny; j++) {
        double h;
        if (!exp_mode) h = mask(i, j) * heating_rate;
        else {
            h = (mask(i, j) * exp_base_heating_rate) * std::exp((pos_y(i, j) * -1.0) / exp_scale_height);
            if (split_exp_mode) {
                const double sb = exp_base_heating_rate * std::exp((exp_scale_height - split_exp_scale_height) * split_exp_start_height / (exp_scale_height * split_exp_scale_height));
                const double h2 = (mask(i, j) * sb) * std::exp((pos_y(i, j) * -1.0) / split_exp_scale_height);
                h = (h < h2) ? h2 : h;
            }
        }
        heating(i, j) = h;
    }
    PlasmaDomain::check(spruce_module_ambient_heating(m_pd.device(), heating.ptr(), heating.size()));
}
