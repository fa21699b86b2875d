// module.cpp -- Module base, ModuleHandler and the three solar modules that run on the device.
// Config-block parsing follows module.cpp:11-35 / modulehandler.cpp:77-112 of the reference; the parsed values are
// handed to libspruce_b200.so in setupModule(), in config order (= execution order).
#include "module.hpp"
#include "sgfilter.hpp"
#include "tracer.hpp"
#include "ucnp_modules.hpp"
#include "viscosity_profile.hpp"
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
    if (name == "eic_thermalization") {
        m_modules.emplace_back(new EICThermalization(m_pd));
        m_modules.back()->configureModule(in);
        return;
    }
    if (name == "coulomb_explosion" || name == "global_temperature") {                  // the UCNP modules take any equation set that has the grids they name
        if (name == "coulomb_explosion") m_modules.emplace_back(new CoulombExplosion(m_pd));
        else m_modules.emplace_back(new GlobalTemperature(m_pd));
        m_modules.back()->configureModule(in);
        return;
    }
    if (name == "artificial_viscosity") {                                               // generic over the equation set in the reference (viscosity.cpp:26-35); the device library
        m_modules.emplace_back(new Viscosity(m_pd));                                    // takes it on ideal_mhd and ideal_mhd_2E and refuses ideal_2F with a message
        m_modules.back()->configureModule(in);
        return;
    }
    SPRUCE_REQUIRE(dynamic_cast<IdealMHD *>(m_pd.m_eqs.get()) != nullptr, "Module designed for IdealMHD EquationSet (ensure that equation_set is set before modules in the config)");
    if (name == "thermal_conduction") m_modules.emplace_back(new ThermalConduction(m_pd));
    else if (name == "radiative_losses") m_modules.emplace_back(new RadiativeLosses(m_pd));
    else if (name == "ambient_heating") m_modules.emplace_back(new AmbientHeating(m_pd));
    else if (name == "physical_viscosity") m_modules.emplace_back(new PhysicalViscosity(m_pd));
    else if (name == "ambient_heating_sink") m_modules.emplace_back(new AmbientHeatingSink(m_pd));
    else if (name == "localized_heating") m_modules.emplace_back(new GaussianSource(m_pd, GaussianSource::Heating));
    else if (name == "mass_injection") m_modules.emplace_back(new GaussianSource(m_pd, GaussianSource::Mass));
    else if (name == "momentum_injection") m_modules.emplace_back(new GaussianSource(m_pd, GaussianSource::Momentum));
    else if (name == "div_cleaning") m_modules.emplace_back(new DivCleaning(m_pd));
    else if (name == "field_heating") m_modules.emplace_back(new FieldHeating(m_pd));
    else if (name == "boundary_outflow") m_modules.emplace_back(new BoundaryOutflow(m_pd));
    else if (name == "anomalous_resistivity") m_modules.emplace_back(new AnomalousResistivity(m_pd));
    else if (name == "sg_filtering") m_modules.emplace_back(new SGFilter(m_pd));
    else if (name == "tracer_particles") m_modules.emplace_back(new TracerParticles(m_pd));
    else spruce_die("Module <" + name + "> is not a module of the B200 path (thermal_conduction, radiative_losses, ambient_heating, artificial_viscosity, physical_viscosity, "
                    "eic_thermalization, ambient_heating_sink, localized_heating, mass_injection, momentum_injection, div_cleaning, field_heating, boundary_outflow, anomalous_resistivity are; "
                    "sg_filtering, tracer_particles, coulomb_explosion and global_temperature run on the host).");
    m_modules.back()->configureModule(in);
}

std::vector<std::string> ModuleHandler::getCommandLineMessages() const
{
    std::vector<std::string> out;
    for (auto &m : m_modules) { const std::string s = m->commandLineMessage(); if (!s.empty()) out.push_back(s); }
    return out;
}
bool ModuleHandler::hasHostModules() const { for (auto &m : m_modules) if (!m->device_resident()) return true; return false; }
void ModuleHandler::preIterateModules(double dt) { for (auto &m : m_modules) if (!m->device_resident()) m->preIterateModule(dt); }
void ModuleHandler::iterateModules(double dt) { for (auto &m : m_modules) if (!m->device_resident()) m->iterateModule(dt); }
void ModuleHandler::postIterateModules(double dt) { for (auto &m : m_modules) if (!m->device_resident()) m->postIterateModule(dt); }
void ModuleHandler::getFileOutputData(std::vector<std::string> &names, std::vector<Grid> &grids) const { for (auto &m : m_modules) m->fileOutput(names, grids); }

static int integrator_id(std::string s, const char *who)
{
    if (s.empty()) s = "euler";
    if (s == "euler") return SPRUCE_TI_EULER;
    if (s == "rk2") return SPRUCE_TI_RK2;
    if (s == "rk4") return SPRUCE_TI_RK4;
    spruce_die(std::string("Invalid time integrator given for ") + who + " module");
}

// ms_electron_heating_fraction (multispecies_mode): checked like the reference's asserts (e.g. thermalconduction.cpp:40), handed over when the config sets it
// (the library holds the reference's defaults)
static void send_ms_fraction(PlasmaDomain &pd, const char *module, const char *label, double f, bool given)
{
    if (!(f >= 0.0 && f <= 1.0)) spruce_die(std::string(label) + " MS electron heating fraction must be between 0 and 1");
    if (given) PlasmaDomain::check(spruce_module_ms_fraction(pd.device(), module, f));
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
        else if (k == "ms_electron_heating_fraction") { ms_electron_heating_fraction = std::stod(v); ms_given = true; }
        else std::cerr << k << " config not recognized for Thermal Conduction Module.\n";
    }
}
void ThermalConduction::setupModule()
{
    PlasmaDomain::check(spruce_module_thermal_conduction(m_pd.device(), flux_saturation, integrator_id(time_integrator, "Thermal Conduction"), epsilon, dt_subcycle_min, weakening_factor));
    if (output_to_file) PlasmaDomain::check(spruce_module_output_to_file(m_pd.device(), "thermal_conduction", 1));
    if (inactive_mode) PlasmaDomain::check(spruce_module_inactive_mode(m_pd.device(), "thermal_conduction", 1));       // evaluated, not applied (:109)
    send_ms_fraction(m_pd, "thermal_conduction", "Thermal Conduction", ms_electron_heating_fraction, ms_given);
}
// the device keeps the two diagnostic planes of the last step; zero planes before the first one, like the reference's
static void append_device_plane(PlasmaDomain &pd, const char *name, std::vector<std::string> &names, std::vector<Grid> &grids)
{
    Grid g(pd.xdim(), pd.ydim());
    PlasmaDomain::check(spruce_module_output(pd.device(), name, pd.slab(g), pd.slabCount()));
    pd.gatherRows(g);
    names.push_back(name);
    grids.push_back(g);
}
void ThermalConduction::fileOutput(std::vector<std::string> &names, std::vector<Grid> &grids)
{
    if (!output_to_file) return;
    append_device_plane(m_pd, "thermal_conduction", names, grids);
    append_device_plane(m_pd, "flux_saturation", names, grids);
}
std::string ThermalConduction::commandLineMessage() const
{
    int n = 0;
    spruce_module_subcycles(m_pd.device(), "thermal_conduction", &n);
    return "Thermal Subcycles: " + std::to_string(n) + (inactive_mode ? " (Not Applied)" : "");      // thermalconduction.cpp:244
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
        else if (k == "ms_electron_heating_fraction") { ms_electron_heating_fraction = std::stod(v); ms_given = true; }
        else std::cerr << k << " config not recognized.\n";
    }
}
void RadiativeLosses::setupModule()
{
    PlasmaDomain::check(spruce_module_radiative_losses(m_pd.device(), integrator_id(time_integrator, "Radiative Losses"), cutoff_ramp, cutoff_temp, epsilon, prevent_subcycling));
    if (output_to_file) PlasmaDomain::check(spruce_module_output_to_file(m_pd.device(), "radiative_losses", 1));
    if (inactive_mode) PlasmaDomain::check(spruce_module_inactive_mode(m_pd.device(), "radiative_losses", 1));         // evaluated, not applied (:98)
    send_ms_fraction(m_pd, "radiative_losses", "Rad. Losses", ms_electron_heating_fraction, ms_given);
}
void RadiativeLosses::fileOutput(std::vector<std::string> &names, std::vector<Grid> &grids)
{
    if (output_to_file) append_device_plane(m_pd, "rad", names, grids);
}
std::string RadiativeLosses::commandLineMessage() const
{
    int n = 0;
    spruce_module_subcycles(m_pd.device(), "radiative_losses", &n);
    return "Radiative Subcycles: " + std::to_string(n) + (inactive_mode ? " (Not Applied)" : "");    // radiativelosses.cpp:171
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
        else if (k == "ms_electron_heating_fraction") { ms_electron_heating_fraction = std::stod(v); ms_given = true; }
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
    PlasmaDomain::check(spruce_module_ambient_heating(m_pd.device(), m_pd.slab(heating), m_pd.slabCount()));
    send_ms_fraction(m_pd, "ambient_heating", "Ambient Heating", ms_electron_heating_fraction, ms_given);
}

// ambientheatingsink.cpp:12-25
void AmbientHeatingSink::parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs)
{
    for (size_t i = 0; i < lhs.size(); i++) {
        const std::string &k = lhs[i], &v = rhs[i];
        if (k == "heating_rate") heating_rate = std::stod(v);
        else if (k == "exp_mode") exp_mode = (v == "true");
        else if (k == "exp_base_heating_rate") exp_base_heating_rate = std::stod(v);
        else if (k == "exp_scale_height") exp_scale_height = std::stod(v);
        else if (k == "center_x") center_x = std::stod(v);
        else if (k == "half_width") half_width = std::stod(v);
        else if (k == "ms_electron_heating_fraction") { ms_electron_heating_fraction = std::stod(v); ms_given = true; }
        else std::cerr << k << " config not recognized.\n";
    }
}
// ambientheatingsink.cpp:27-33: the reduction plane (host libm), applied on the device after every step (:35-37)
void AmbientHeatingSink::setupModule()
{
    SPRUCE_REQUIRE(ms_electron_heating_fraction >= 0.0 && ms_electron_heating_fraction <= 1.0, "Ambient Heating Sink MS electron heating fraction must be between 0 and 1");
    const size_t nx = m_pd.xdim(), ny = m_pd.ydim();
    const Grid &mask = m_pd.ghostZoneMask(), &pos_x = m_pd.m_grids[PlasmaDomain::pos_x], &pos_y = m_pd.m_grids[PlasmaDomain::pos_y];
    Grid reduction(nx, ny);
    for (size_t i = 0; i < nx; i++) for (size_t j = 0; j < ny; j++) {
        if (exp_mode) {
            const double q = (pos_x(i, j) - center_x) / half_width;
            const double para = 1.0 - q * q;
            reduction(i, j) = ((mask(i, j) * exp_base_heating_rate) * std::exp((-1.0 * pos_y(i, j)) / exp_scale_height)) * ((para < 0.0) ? 0.0 : para);
        } else reduction(i, j) = mask(i, j) * heating_rate;
    }
    PlasmaDomain::check(spruce_module_ambient_heating_sink(m_pd.device(), m_pd.slab(reduction), m_pd.slabCount()));
    send_ms_fraction(m_pd, "ambient_heating_sink", "Ambient Heating Sink", ms_electron_heating_fraction, ms_given);
}

// localizedheating.cpp:14-29, massinjection.cpp:14-26, momentuminjection.cpp:16-33
void GaussianSource::parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs)
{
    const char *peak_key = m_kind == Heating ? "max_heating_rate" : m_kind == Mass ? "max_injection_rate" : "max_accel";
    for (size_t i = 0; i < lhs.size(); i++) {
        const std::string &k = lhs[i], &v = rhs[i];
        if (k == "start_time") start_time = std::stod(v);
        else if (k == "duration") duration = std::stod(v);
        else if (k == peak_key) peak = std::stod(v);
        else if (k == "stddev_x") stddev_x = std::stod(v);
        else if (k == "stddev_y") stddev_y = std::stod(v);
        else if (k == "center_x") center_x = std::stod(v);
        else if (k == "center_y") center_y = std::stod(v);
        else if (m_kind == Heating && k == "ramp_time") ramp_time = std::stod(v);
        else if (m_kind == Heating && k == "ms_electron_heating_fraction") { ms_electron_heating_fraction = std::stod(v); ms_given = true; }
        else if (m_kind == Momentum && k == "dir_x") dir_x = std::stod(v);
        else if (m_kind == Momentum && k == "dir_y") dir_y = std::stod(v);
        else if (m_kind == Momentum && k == "template_angle") template_angle = std::stod(v);
        else if (m_kind == Momentum && k == "oscillatory") oscillatory = (v == "true");
        else if (m_kind == Momentum && k == "oscillation_period") oscillation_period = std::stod(v);
        else std::cerr << k << " config not recognized.\n";
    }
}
void GaussianSource::setupModule()
{
    spruce_domain *dev = m_pd.device();
    if (m_kind == Heating) {
        SPRUCE_REQUIRE(ms_electron_heating_fraction >= 0.0 && ms_electron_heating_fraction <= 1.0, "Localized Heating MS electron heating fraction must be between 0 and 1");
        PlasmaDomain::check(spruce_module_localized_heating(dev, start_time, duration, peak, stddev_x, stddev_y, center_x, center_y, ramp_time));
        send_ms_fraction(m_pd, "localized_heating", "Localized Heating", ms_electron_heating_fraction, ms_given);
    } else if (m_kind == Mass) {
        PlasmaDomain::check(spruce_module_mass_injection(dev, start_time, duration, peak, stddev_x, stddev_y, center_x, center_y));
    } else {
        SPRUCE_REQUIRE(!(dir_x == 0.0 && dir_y == 0.0), "Momentum Injection module must be given a nonzero acceleration direction");
        PlasmaDomain::check(spruce_module_momentum_injection(dev, start_time, duration, peak, stddev_x, stddev_y, center_x, center_y, dir_x, dir_y, template_angle,
                                                             oscillatory ? 1 : 0, oscillation_period));
    }
}
// localizedheating.cpp:69-78, massinjection.cpp:55-64, momentuminjection.cpp:78-87
std::string GaussianSource::commandLineMessage() const
{
    std::ostringstream oss;
    oss.precision(4);
    oss << center_x << "," << center_y;
    std::string result = std::string(m_kind == Heating ? "Heating at " : m_kind == Mass ? "Mass injection at " : "Momentum injection at ") + oss.str();
    const double t = m_pd.time();
    result += (t < start_time || t > start_time + duration) ? " Off" : " On";
    return result;
}

// divcleaning.cpp:11-19
void DivCleaning::parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs)
{
    for (size_t i = 0; i < lhs.size(); i++) {
        const std::string &k = lhs[i], &v = rhs[i];
        if (k == "epsilon") epsilon = std::stod(v);
        if (k == "time_scale") time_scale = std::stod(v);
        else std::cerr << k << " config not recognized.\n";            // the reference prints this for "epsilon" too (an if where an else-if was meant)
    }
}
void DivCleaning::setupModule() { PlasmaDomain::check(spruce_module_div_cleaning(m_pd.device(), epsilon, time_scale)); }

// fieldheating.cpp:14-28
void FieldHeating::parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs)
{
    for (size_t i = 0; i < lhs.size(); i++) {
        const std::string &k = lhs[i], &v = rhs[i];
        if (k == "output_to_file") output_to_file = (v == "true");
        else if (k == "coeff") coeff = std::stod(v);
        else if (k == "current_pow") current_pow = std::stod(v);
        else if (k == "b_pow") b_pow = std::stod(v);
        else if (k == "n_pow") n_pow = std::stod(v);
        else if (k == "roc_pow") roc_pow = std::stod(v);
        else if (k == "inactive_mode") inactive_mode = (v == "true");
        else std::cerr << k << " config not recognized.\n";
    }
}
void FieldHeating::setupModule()
{
    PlasmaDomain::check(spruce_module_field_heating(m_pd.device(), coeff, current_pow, b_pow, n_pow, roc_pow, inactive_mode ? 1 : 0));
}
// fieldheating.cpp:73-80
void FieldHeating::fileOutput(std::vector<std::string> &names, std::vector<Grid> &grids)
{
    if (output_to_file) append_device_plane(m_pd, "field_heating", names, grids);
}
// fieldheating.cpp:64-71
std::string FieldHeating::commandLineMessage() const
{
    std::string message = "Field Heating";
    message += (coeff == 0.0) ? " Zero" : " On";
    if (inactive_mode) message += " (Not Applied)";
    return message;
}

// boundaryoutflow.cpp:16-31
void BoundaryOutflow::parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs)
{
    for (size_t i = 0; i < lhs.size(); i++) {
        const std::string &k = lhs[i], &v = rhs[i];
        if (k == "max_accel") max_accel = std::stod(v);
        else if (k == "falloff_length") falloff_length = std::stod(v);
        else if (k == "boundary") boundary = v;
        else if (k == "falloff_shape") falloff_shape = v;
        else if (k == "feather_length") feather_length = std::stod(v);
        else if (k == "field_aligned_mode") field_aligned_mode = (v == "true");
        else if (k == "dynamic_mode") dynamic_mode = (v == "true");
        else if (k == "dynamic_time") dynamic_time = std::stod(v);
        else if (k == "dynamic_target_speed") dynamic_target_speed = std::stod(v);
        else std::cerr << k << " config not recognized.\n";
    }
}
// boundaryoutflow.cpp:33-37: the template and the outflow window are built by the device library from pos_x / pos_y
void BoundaryOutflow::setupModule()
{
    SPRUCE_REQUIRE(falloff_shape == "exp" || falloff_shape == "gaussian" || falloff_shape == "flat", "BoundaryOutflow shape must be exp or gaussian or flat");
    const int b = boundary == "x_bound_1" ? 0 : boundary == "x_bound_2" ? 1 : boundary == "y_bound_1" ? 2 : boundary == "y_bound_2" ? 3 : -1;
    SPRUCE_REQUIRE(b >= 0, "BoundaryOutflow boundary config must be {x,y}_bound_{1,2}");
    const int sh = falloff_shape == "exp" ? 0 : falloff_shape == "gaussian" ? 1 : 2;
    const Grid &x = m_pd.m_grids[PlasmaDomain::pos_x], &y = m_pd.m_grids[PlasmaDomain::pos_y];
    PlasmaDomain::check(spruce_module_boundary_outflow(m_pd.device(), x.ptr(), y.ptr(), x.size(),   // the whole domain on every rank: the template is built from global positions
                                                       max_accel, falloff_length, b, sh, feather_length,
                                                       field_aligned_mode ? 1 : 0, dynamic_mode ? 1 : 0, dynamic_time, dynamic_target_speed));
}
// boundaryoutflow.cpp:65-74
std::string BoundaryOutflow::commandLineMessage() const
{
    double mean = 0.0, accel = 0.0;
    PlasmaDomain::check(spruce_module_boundary_outflow_state(m_pd.device(), &mean, &accel));
    return boundary + " boundary outflow enforced (max " + std::to_string(mean) + " cm/s outflow) (accel. " + std::to_string(accel) + " cm/s^2)";
}

// anomalousresistivity.cpp:46-68
void AnomalousResistivity::parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs)
{
    for (size_t i = 0; i < lhs.size(); i++) {
        const std::string &k = lhs[i], &v = rhs[i];
        if (k == "time_scale") time_scale = std::stod(v);
        else if (k == "output_to_file") output_to_file = (v == "true");
        else if (k == "frobenius_metric_coeff") frobenius_metric_coeff = std::stod(v);
        else if (k == "smoothing_sigma") smoothing_sigma = std::stod(v);
        else if (k == "safety_factor") safety_factor = std::stod(v);
        else if (k == "metric_smoothing") metric_smoothing = (v == "true");
        else if (k == "time_integrator") time_integrator = v;
        else if (k == "template_mode") template_mode = v;
        else if (k == "flood_fill_max_radius") flood_fill_max_radius = std::stod(v);
        else if (k == "flood_fill_argmin_radius") flood_fill_argmin_radius = std::stod(v);
        else if (k == "flood_fill_min_current") flood_fill_min_current = std::stod(v);
        else if (k == "flood_fill_current_ramp_length") flood_fill_current_ramp_length = std::stod(v);
        else if (k == "resistivity_model") resistivity_model = v;
        else if (k == "gradient_correction") gradient_correction = (v == "true");
        else if (k == "resistivity_model_params") { resistivity_model_params.clear(); for (const std::string &el : splitString(v, ',')) resistivity_model_params.push_back(std::stod(el)); }
        else if (k == "flood_fill_threshold") flood_fill_threshold = std::stod(v);
        else std::cerr << k << " config not recognized.\n";
    }
}
// anomalousresistivity.cpp:18-44: the null point search, the smoothing kernel and the first template are the device library's
void AnomalousResistivity::setupModule()
{
    SPRUCE_REQUIRE(resistivity_model != "time_scale" || time_scale > 0.0, "Anomalous Resistivity time_scale must be specified, and positive, if run in time_scale mode");
    SPRUCE_REQUIRE(!metric_smoothing || smoothing_sigma > 0.0, "smoothing_sigma must be a positive number");
    SPRUCE_REQUIRE(time_integrator.empty() || time_integrator == "euler" || time_integrator == "rk2" || time_integrator == "rk4", "Invalid time integrator for anomalous resistivity module");
    SPRUCE_REQUIRE(template_mode == "flood_fill" || template_mode == "frobenius", "Invalid template_mode for anomalous resistivity module");
    const int model = resistivity_model == "time_scale" ? 0 : resistivity_model == "syntelis_19" ? 1 : resistivity_model == "ys_94" ? 2 : -1;
    SPRUCE_REQUIRE(model >= 0, "Invalid resistivity_model for anomalous resistivity module");
    SPRUCE_REQUIRE(model == 0 || resistivity_model_params.size() == 3, "resistivity_model_params must hold three values for this resistivity_model");
    double p[SPRUCE_AR_N_PARAMS] = {time_scale, frobenius_metric_coeff, smoothing_sigma, safety_factor, metric_smoothing ? 1.0 : 0.0,
                                    time_integrator == "rk2" ? 1.0 : time_integrator == "rk4" ? 2.0 : 0.0, template_mode == "flood_fill" ? 1.0 : 0.0, flood_fill_max_radius,
                                    flood_fill_argmin_radius, flood_fill_min_current, flood_fill_current_ramp_length, flood_fill_threshold, (double)model,
                                    gradient_correction ? 1.0 : 0.0, 0.0, 0.0, 0.0};
    for (size_t k = 0; k < resistivity_model_params.size() && k < 3; k++) p[14 + k] = resistivity_model_params[k];
    const Grid &x = m_pd.m_grids[PlasmaDomain::pos_x], &y = m_pd.m_grids[PlasmaDomain::pos_y];
    PlasmaDomain::check(spruce_module_anomalous_resistivity(m_pd.device(), x.ptr(), y.ptr(), x.size(), p, SPRUCE_AR_N_PARAMS));
    if (output_to_file) PlasmaDomain::check(spruce_module_output_to_file(m_pd.device(), "anomalous_resistivity", 1));
}
// anomalousresistivity.cpp:315-318
std::string AnomalousResistivity::commandLineMessage() const
{
    int n = 0;
    spruce_module_subcycles(m_pd.device(), "anomalous_resistivity", &n);
    return "Anomalous Resistivity Subcycles: " + std::to_string(n);
}
// anomalousresistivity.cpp:320-329
void AnomalousResistivity::fileOutput(std::vector<std::string> &names, std::vector<Grid> &grids)
{
    if (!output_to_file) return;
    append_device_plane(m_pd, "anomalous_diffusivity", names, grids);
    append_device_plane(m_pd, "anomalous_template", names, grids);
    append_device_plane(m_pd, "joule_heating", names, grids);
}

// viscosity.cpp:6-24
void Viscosity::parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs)
{
    for (size_t i = 0; i < lhs.size(); i++) {
        const std::string &k = lhs[i], &v = rhs[i];
        if (k == "visc_output_visc") m_output_visc = (v == "true");                 // (the reference leaves these four uninitialised when the config is silent,
        else if (k == "visc_output_lap") m_output_lap = (v == "true");              //  viscosity.hpp:60-63: here they default to false)
        else if (k == "visc_output_strength") m_output_strength = (v == "true");
        else if (k == "visc_output_timescale") m_output_timescale = (v == "true");
        else if (k == "hv_time_integrator") m_hv_time_integrator = v;
        else if (k == "gradient_correction") m_gradient_correction = (v == "true");
        else if (k == "hv_epsilon") m_hv_epsilon = std::stod(v);
        else if (k == "visc_opt") m_inp_visc_opt = v;
        else if (k == "boundary_falloff_shape") m_boundary_falloff_shape = v;
        else if (k == "visc_strength") m_inp_strength = v;
        else if (k == "visc_vars_to_diff") m_inp_vars_to_diff = v;
        else if (k == "visc_vars_to_evol") m_inp_vars_to_evol = v;
        else if (k == "visc_length") m_inp_length = v;
        else if (k == "visc_species") m_inp_species = v;
        else std::cerr << k << " config not recognized.\n";
    }
}
// viscosity.cpp:278-325, gaussian and exp shapes (the elliptical shapes are not ported); host libm, static profile
Grid Viscosity::getBoundaryViscosity(double strength, double length) const
{
    Grid result;
    const std::string shape = m_boundary_falloff_shape.empty() ? "gaussian" : m_boundary_falloff_shape;          // viscosity.cpp:90-93 defaults to gaussian
    if (!boundaryViscosityProfile(m_pd.m_grids[PlasmaDomain::pos_x], m_pd.m_grids[PlasmaDomain::pos_y], strength, length, shape, result))      // viscosity_profile.hpp
        spruce_die("boundary_falloff_shape <" + shape + "> in Viscosity module not recognized");
    return result;
}
// viscosity.cpp:37-110
void Viscosity::setupModule()
{
    for (std::string *s : {&m_inp_visc_opt, &m_inp_strength, &m_inp_vars_to_diff, &m_inp_vars_to_evol, &m_inp_length, &m_inp_species}) clearWhitespace(*s);
    const std::vector<std::string> opt = splitString(m_inp_visc_opt, ','), str = splitString(m_inp_strength, ','), diff = splitString(m_inp_vars_to_diff, ','),
                                   evol = splitString(m_inp_vars_to_evol, ','), len = splitString(m_inp_length, ','), spec = splitString(m_inp_species, ',');
    const size_t n = opt.size();
    SPRUCE_REQUIRE(n > 0 && str.size() == n && diff.size() == n && evol.size() == n && len.size() == n && spec.size() == n, "every viscosity list must have one entry per term");
    if (m_boundary_falloff_shape.empty()) m_boundary_falloff_shape = "gaussian";
    SPRUCE_REQUIRE(m_boundary_falloff_shape == "gaussian" || m_boundary_falloff_shape == "exp" || m_boundary_falloff_shape == "exp_elliptical" || m_boundary_falloff_shape == "gaussian_elliptical",
                   "Invalid boundary falloff shape given for Viscosity module");
    PlasmaDomain::check(spruce_module_viscosity(m_pd.device(), integrator_id(m_hv_time_integrator, "Viscosity"), m_hv_epsilon, m_gradient_correction));
    for (size_t i = 0; i < n; i++) {
        SPRUCE_REQUIRE(std::stod(len[i]) >= 0, "Length constants must greater than or equal to zero.");
        const double strength = std::stod(str[i]);
        if (opt[i] == "boundary" || opt[i] == "boundary_global") {
            const Grid prof = getBoundaryViscosity(strength, std::stod(len[i]));
            PlasmaDomain::check(spruce_module_viscosity_term(m_pd.device(), opt[i].c_str(), strength, diff[i].c_str(), evol[i].c_str(), spec[i].c_str(), m_pd.slab(prof), m_pd.slabCount()));
        } else {
            PlasmaDomain::check(spruce_module_viscosity_term(m_pd.device(), opt[i].c_str(), strength, diff[i].c_str(), evol[i].c_str(), spec[i].c_str(), nullptr, 0));
        }
    }
    m_vars_to_evol = evol;
    if (m_output_visc || m_output_lap || m_output_strength || m_output_timescale) PlasmaDomain::check(spruce_module_output_to_file(m_pd.device(), "artificial_viscosity", 1));
}
// viscosity.cpp:351-376: per flag, one plane per term in config order -- <evolved>_dqdt, _lap, _str, _dt (:103-107) --, each what the term's last evaluation on the
// device left (zero planes before the first one)
void Viscosity::fileOutput(std::vector<std::string> &names, std::vector<Grid> &grids)
{
    const struct { bool on; const char *plane, *suffix; } kinds[4] = {{m_output_visc, "visc_dqdt:", "_dqdt"}, {m_output_lap, "visc_lap:", "_lap"},
                                                                      {m_output_strength, "visc_str:", "_str"}, {m_output_timescale, "visc_dt:", "_dt"}};
    for (const auto &k : kinds) {
        if (!k.on) continue;
        for (size_t i = 0; i < m_vars_to_evol.size(); i++) {
            append_device_plane(m_pd, (k.plane + std::to_string(i)).c_str(), names, grids);
            names.back() = m_vars_to_evol[i] + k.suffix;
        }
    }
}

// eic_thermalization.cpp:12-25: every grid the module reads must exist in the equation set
void EICThermalization::setupModule()
{
    for (const char *name : {"n", "e_temp", "e_thermal_energy", "i_thermal_energy"})
        if (!m_pd.m_eqs->is_var(name)) spruce_die(std::string("Grid <") + name + "> was not found within the EquationSet.");
    PlasmaDomain::check(spruce_module_eic_thermalization(m_pd.device()));
}

// physicalviscosity.cpp:17-35
void PhysicalViscosity::parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs)
{
    for (size_t i = 0; i < lhs.size(); i++) {
        const std::string &k = lhs[i], &v = rhs[i];
        if (k == "output_to_file") output_to_file = (v == "true");
        else if (k == "coeff") coeff = std::stod(v);
        else if (k == "ramp_length") ramp_length = std::stod(v);
        else if (k == "buffer_length") buffer_length = std::stod(v);
        else if (k == "epsilon") epsilon = std::stod(v);
        else if (k == "heating_on") heating_on = (v == "true");
        else if (k == "force_on") force_on = (v == "true");
        else if (k == "inactive_mode") inactive_mode = (v == "true");
        else if (k == "gradient_correction") gradient_correction = (v == "true");
        else if (k == "time_integrator") time_integrator = v;
        else if (k == "ms_electron_heating_fraction") { ms_electron_heating_fraction = std::stod(v); ms_given = true; }
        else std::cerr << k << " config not recognized.\n";
    }
}
// physicalviscosity.cpp:247-267: elliptical Gaussian ramp of the coefficient towards the domain edge (buffer_length is parsed but unused)
Grid PhysicalViscosity::constructCoefficientGrid(double strength, double ramp, double) const
{
    const size_t nx = m_pd.xdim(), ny = m_pd.ydim();
    Grid result = Grid::Ones(nx, ny);
    if (ramp == 0.0) { for (double &q : result.data()) q = strength * q; return result; }
    const Grid &x = m_pd.m_grids[PlasmaDomain::pos_x], &y = m_pd.m_grids[PlasmaDomain::pos_y];
    double x_min = x(0, 0), x_max = x(0, 0), y_min = y(0, 0), y_max = y(0, 0);
    for (size_t i = 0; i < nx; i++) for (size_t j = 0; j < ny; j++) {
        x_min = std::min(x_min, x(i, j)); x_max = std::max(x_max, x(i, j)); y_min = std::min(y_min, y(i, j)); y_max = std::max(y_max, y(i, j));
    }
    const double xc = 0.5 * (x_min + x_max), yc = 0.5 * (y_min + y_max);
    const double px = std::pow(x_max - xc, 2.0), py = std::pow(y_max - yc, 2.0);
    const double s_length = ramp / std::min(x_max - xc, y_max - yc);
    for (size_t i = 0; i < nx; i++) for (size_t j = 0; j < ny; j++) {
        const double ex = x(i, j) - xc, ey = y(i, j) - yc;
        const double s = (ex * ex) / px + (ey * ey) / py;
        const double a = std::max((s + 2.0 * s_length) - 1.0, 0.0) / s_length;
        result(i, j) = strength * ((1.0 / 0.99) * std::max(std::exp((a * a) * -2.3) - 0.01, 0.0));
    }
    return result;
}
// physicalviscosity.cpp:37-45
void PhysicalViscosity::setupModule()
{
    SPRUCE_REQUIRE(time_integrator.empty() || time_integrator == "euler" || time_integrator == "rk2", "Invalid time integrator given for Physical Viscosity module");
    SPRUCE_REQUIRE(ms_electron_heating_fraction >= 0.0 && ms_electron_heating_fraction <= 1.0, "Physical Viscosity MS electron heating fraction must be between 0 and 1");
    const Grid cg = constructCoefficientGrid(coeff, ramp_length, buffer_length);
    PlasmaDomain::check(spruce_module_physical_viscosity(m_pd.device(), coeff, m_pd.slab(cg), m_pd.slabCount(), epsilon, heating_on, force_on, gradient_correction,
                                                         integrator_id(time_integrator, "Physical Viscosity"), inactive_mode));
    if (output_to_file) PlasmaDomain::check(spruce_module_output_to_file(m_pd.device(), "physical_viscosity", 1));
    send_ms_fraction(m_pd, "physical_viscosity", "Physical Viscosity", ms_electron_heating_fraction, ms_given);
}
// physicalviscosity.cpp:292-308: the sub-cycle averages of the last step, kept on the device (zero planes before the first step)
void PhysicalViscosity::fileOutput(std::vector<std::string> &names, std::vector<Grid> &grids)
{
    if (!output_to_file) return;
    if (heating_on) append_device_plane(m_pd, "viscous_heating", names, grids);
    if (force_on) for (const char *nm : {"viscous_force_x", "viscous_force_y", "viscous_force_z"}) append_device_plane(m_pd, nm, names, grids);
}
// physicalviscosity.cpp:269-287
std::string PhysicalViscosity::commandLineMessage() const
{
    std::string message;
    if (heating_on) { message += "Viscous Heating"; message += (coeff == 0.0) ? " Zero" : " On"; if (force_on) message += ", "; }
    if (force_on) { message += "Viscous Force"; message += (coeff == 0.0) ? " Zero" : " On"; }
    int n = 1;
    PlasmaDomain::check(spruce_module_subcycles(m_pd.device(), "physical_viscosity", &n));
    if (force_on || heating_on) message += ", " + std::to_string(n) + " Subcycle(s)";
    if (inactive_mode) message += " (Not Applied)";
    return message;
}

// ---- sg_filtering (source/modules/sgfilter.cpp), host-resident
void SGFilter::parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs)
{
    for (size_t i = 0; i < lhs.size(); i++) {
        if (lhs[i] == "filter_interval") filter_interval = (int)std::stod(rhs[i]);       // sgfilter.cpp:14: stod into an int member
        else std::cerr << lhs[i] << " config not recognized.\n";
    }
}
void SGFilter::postIterateModule(double)
{
    // m_iter is the index of the step that has just been integrated: advanceTime increments it after the post-iterate hooks (evolution.cpp:74-81)
    if (filter_interval > 0 && m_pd.iter() != 0 && m_pd.iter() % filter_interval == 0) applyFilter();
}
void SGFilter::applyFilter()
{
    const bool xp = m_pd.xPeriodic(), yp = m_pd.yPeriodic();
    for (const char *name : {"rho", "thermal_energy"}) {                                  // sgfilter.cpp:38
        Grid &g = m_pd.eqs()->grid(name);                                                 // staged from the device
        singleVarSavitzkyGolay(g, m_pd.xl(), m_pd.xu(), m_pd.yl(), m_pd.yu(), xp, yp);
        m_pd.eqs()->pushGrid(name);
    }
    m_pd.eqs()->propagateChanges();                                                       // sgfilter.cpp:42
}
void SGFilter::singleVarSavitzkyGolay(Grid &grid, int xl, int xu, int yl, int yu, bool x_periodic, bool y_periodic)
{
    (void)x_periodic;                                         // the reference wraps the row indices (sgfilter.cpp:62-67) and then never reads them
    // every tap reads grid(j, j) with j a COLUMN index: the reference's accessor asserts (aborts) as soon as a column index is not a valid row index
    SPRUCE_REQUIRE(grid.cols() <= grid.rows(), "sg_filtering: the reference reads grid(j, j) for column indices j (sgfilter.cpp:75) and aborts when ydim > xdim");
    sgFilterPlane(grid, xl, xu, yl, yu, y_periodic);          // sgfilter.hpp
}

// ---- tracer_particles (source/modules/solar/tracerparticles.cpp), host-resident
void TracerParticles::parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs)
{
    for (size_t i = 0; i < lhs.size(); i++) {
        if (lhs[i] == "init_file") m_init_filename = fs::path(rhs[i]);
        else std::cerr << lhs[i] << " config not recognized.\n";
    }
}
void TracerParticles::setupModule()
{
    const Grid &px = m_pd.m_grids[PlasmaDomain::pos_x], &py = m_pd.m_grids[PlasmaDomain::pos_y];
    for (size_t i = 0; i < m_pd.xdim(); i++) x_vec.push_back(px(i, 0));
    for (size_t j = 0; j < m_pd.ydim(); j++) y_vec.push_back(py(0, j));
    if (m_pd.m_continue_mode) {                                  // :20-23
        m_init_filename = m_pd.m_out_directory / fs::path("end.tpstate");
        SPRUCE_REQUIRE(fs::is_regular_file(m_init_filename), "Continue directory must contain end.tpstate file for tracer particles");
    } else {                                                     // :24-35: the particle file travels with the run, an old output is discarded
        const fs::path new_init_path = m_pd.m_out_directory / "init.tpstate";
        SPRUCE_REQUIRE(fs::is_regular_file(m_init_filename), "Tracer particle initialization file was not found");
        if (!fs::exists(new_init_path) || !fs::equivalent(m_init_filename, new_init_path)) {
            if (fs::exists(new_init_path)) fs::remove(new_init_path);
            fs::copy(m_init_filename, new_init_path, fs::copy_options::overwrite_existing);
        }
        const fs::path old_out_path = m_pd.m_out_directory / "particles.tpout";
        if (fs::exists(old_out_path)) fs::remove(old_out_path);
    }
    readTPStateFile(m_init_filename);                            // (the reference's particle-count assertion runs before the file is read, i.e. never fires: :17)
    if (!m_pd.m_continue_mode) writeToTPOutFile(0.0);
}
void TracerParticles::iterateModule(double dt)
{
    const Grid v_x = m_pd.eqs()->grid("v_x"), v_y = m_pd.eqs()->grid("v_y");          // staged from the device
    const bool xper = m_pd.x_bound_1 == PlasmaDomain::BoundaryCondition::Periodic, yper = m_pd.y_bound_1 == PlasmaDomain::BoundaryCondition::Periodic;
    for (int i = (int)m_particles.size() - 1; i >= 0; i--) {
        std::vector<double> &p = m_particles[i];
        const double v_x_p = bilinearInterpolate(p, v_x, x_vec, y_vec), v_y_p = bilinearInterpolate(p, v_y, x_vec, y_vec);
        const std::vector<double> half = {p[0] + 0.5 * dt * v_x_p, p[1] + 0.5 * dt * v_y_p};
        const double v_x_h = bilinearInterpolate(half, v_x, x_vec, y_vec), v_y_h = bilinearInterpolate(half, v_y, x_vec, y_vec);
        p[0] += dt * v_x_h; p[1] += dt * v_y_h;
        if (p[0] < x_vec[0] || p[0] > x_vec.back()) {
            if (xper) { const double width = x_vec.back() - x_vec[0]; p[0] = std::fmod((p[0] - x_vec[0] + width), width) + x_vec[0]; }
            else {
                m_particles.erase(m_particles.begin() + i); m_labels.erase(m_labels.begin() + i);
                continue;                                        // the reference goes on to test p[1] through its reference to the erased element (:76): not reproduced
            }
        }
        if (p[1] < y_vec[0] || p[1] > y_vec.back()) {
            if (yper) { const double height = y_vec.back() - y_vec[0]; p[1] = std::fmod((p[1] - y_vec[0] + height), height) + y_vec[0]; }
            else { m_particles.erase(m_particles.begin() + i); m_labels.erase(m_labels.begin() + i); }
        }
    }
    const int old_time_iter = (int)(m_pd.m_time / m_pd.m_time_output_interval), new_time_iter = (int)((m_pd.m_time + dt) / m_pd.m_time_output_interval);
    const bool store_1 = m_pd.m_iter_output_interval > 0 && (m_pd.m_iter + 1) % m_pd.m_iter_output_interval == 0;
    const bool store_2 = m_pd.m_time_output_interval > 0.0 && new_time_iter > old_time_iter;
    if (store_1 || store_2) writeToTPOutFile(dt);
    writeTPStateFile();
}
void TracerParticles::readTPStateFile(const fs::path &init_path)
{
    SPRUCE_REQUIRE(fs::is_regular_file(init_path), "Tracer particle initialization file must exist");
    std::ifstream in(init_path.string());
    std::string line;
    while (std::getline(in, line)) {
        clearWhitespace(line);
        if (line.empty() || line[0] == '#') continue;
        const std::vector<std::string> parts = splitString(line, '#');
        m_labels.push_back(parts.size() == 1 ? "" : parts[1]);
        const std::vector<std::string> xy = splitString(parts[0], ',');
        m_particles.push_back({std::stod(xy[0]), std::stod(xy[1])});
    }
}
void TracerParticles::writeTPStateFile()
{
    std::ofstream out((m_pd.m_out_directory / m_end_filename).string());
    out.precision(std::numeric_limits<double>::digits10 + 1);
    for (size_t i = 0; i < m_particles.size(); i++) out << m_particles[i][0] << "," << m_particles[i][1] << "#" << m_labels[i] << std::endl;
}
void TracerParticles::writeToTPOutFile(double dt)
{
    std::ofstream out((m_pd.m_out_directory / m_out_filename).string(), std::ofstream::app);
    out.precision(std::numeric_limits<double>::digits10 + 1);
    out << "t=" << m_pd.m_time + dt << std::endl;
    for (size_t i = 0; i < m_particles.size(); i++) out << m_particles[i][0] << "," << m_particles[i][1] << "#" << m_labels[i] << std::endl;
}

// ---- coulomb_explosion (source/modules/ucnp/coulomb_explosion.cpp), host-resident
void CoulombExplosion::parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs)
{
    for (size_t i = 0; i < lhs.size(); i++) {
        if (lhs[i] == "timescale") m_timescale = std::stod(rhs[i]);
        else if (lhs[i] == "lengthscale") m_lengthscale = std::stod(rhs[i]);
        else if (lhs[i] == "strength") m_strength = std::stod(rhs[i]);
        else if (lhs[i] == "output_to_file") output_to_file = (rhs[i] == "true");
        else std::cerr << lhs[i] << " config not recognized.\n";
    }
}
void CoulombExplosion::setupModule()
{
    m_vars.assign(num_vars, Grid::Zero(m_pd.xdim(), m_pd.ydim()));
    for (const char *name : {"press", "n", "mom_x", "mom_y"})
        if (!m_pd.eqs()->is_var(name)) spruce_die(std::string("Grid <") + name + "> was not found within the EquationSet.");
}
void CoulombExplosion::postIterateModule(double dt)
{
    if (!(m_pd.time() < 3 * m_timescale)) return;                                         // m_time still names the start of the step (evolution.cpp:74 before :80)
    const Grid n = m_pd.eqs()->grid("n");                                                 // staged from the device
    if (output_to_file) {                                                                 // the reference differentiates the pressure every step (:65-66); only fileOutput reads it
        const Grid press = m_pd.eqs()->grid("press");
        m_vars[dP_x] = m_pd.derivative1D(press, 0);
        m_vars[dP_y] = m_pd.derivative1D(press, 1);
    }
    const std::string why = ucnp::coulombExplosionForce(m_pd.m_grids[PlasmaDomain::pos_x], m_pd.m_grids[PlasmaDomain::pos_y], n, m_pd.time(), m_timescale, m_lengthscale, m_strength,
                                                        m_vars[F_x], m_vars[F_y]);
    if (!why.empty()) spruce_die(why);
    const Vars force[2] = {F_x, F_y};
    const char *mom[2] = {"mom_x", "mom_y"};
    for (int k = 0; k < 2; k++) {                                                         // :83-84
        Grid &m = m_pd.eqs()->grid(mom[k]);
        const double *f = m_vars[force[k]].ptr();
        for (size_t c = 0; c < (size_t)m.size(); c++) m.ptr()[c] += f[c] * dt;
        m_pd.eqs()->pushGrid(mom[k]);
    }
    m_pd.eqs()->propagateChanges();
}
void CoulombExplosion::fileOutput(std::vector<std::string> &var_names, std::vector<Grid> &var_grids)
{
    if (!output_to_file) return;
    for (int i = 0; i < num_vars; i++) { var_names.push_back(m_var_names[i]); var_grids.push_back(m_vars[i]); }
}

// ---- global_temperature (source/modules/ucnp/global_temperature.cpp), host-resident
void GlobalTemperature::parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs)
{
    for (size_t i = 0; i < lhs.size(); i++) {
        if (lhs[i] == "gt_species") m_species = splitString(rhs[i], ',');
        else if (lhs[i] == "gt_strength") m_strength = std::stod(rhs[i]);
        else if (lhs[i] == "gt_use_diffusion") m_use_diffusion = rhs[i] == "true";
        else if (lhs[i] == "gt_use_global_temp") m_use_global_temp = rhs[i] == "true";
        else std::cerr << lhs[i] << " config not recognized.\n";
    }
}
void GlobalTemperature::setupModule()
{
    SPRUCE_REQUIRE(!m_use_global_temp, "global_temperature: gt_use_global_temp = true is not provided by the B200 path (a domain integral inside every propagateChanges, "
                                       "global_temperature.cpp:45-64; the device fuses propagateChanges into its stage kernels); gt_use_diffusion is");
    const std::vector<std::string> species = m_pd.eqs()->species();
    m_species_ind.assign(m_species.size(), -1);
    for (size_t i = 0; i < m_species.size(); i++) {
        for (size_t j = 0; j < species.size(); j++) if (m_species[i] == species[j]) m_species_ind[i] = (int)j;
        if (m_species_ind[i] < 0) {
            std::cerr << "Species <" << m_species[i] << "> does not correspond to a species within the active equation set." << std::endl;
            spruce_die("<gt_species> was specified incorrectly in the .config file.");
        }
        SPRUCE_REQUIRE(m_species[i] == "e" || m_species[i] == "i", "Species name must be <e> or <i>.");
    }
    m_dr = ucnp::diffusionLengthSquared(m_pd.m_grids[PlasmaDomain::d_x], m_pd.m_grids[PlasmaDomain::d_y], m_pd.ghostZoneMask());
}
void GlobalTemperature::postIterateModule(double dt)
{
    if (!m_use_diffusion) return;
    EquationSet *eqs = m_pd.eqs();
    for (size_t i = 0; i < m_species.size(); i++) {
        Grid temp = eqs->grid(eqs->temperatures()[m_species_ind[i]]);                     // staged from the device
        const Grid n = eqs->grid(eqs->number_densities()[m_species_ind[i]]);
        ucnp::diffuseTemperature(temp, m_dr, dt, m_pd.epsilon, m_strength, [&](const Grid &q) { return m_pd.laplacian(q); });
        const int e_index = eqs->thermal_energies()[m_species_ind[i]];
        Grid &e = eqs->grid(e_index);
        for (size_t c = 0; c < (size_t)e.size(); c++) e.ptr()[c] = n.ptr()[c] * ucnp::kBoltzmann * temp.ptr()[c] / (m_pd.adiabaticIndex() - 1.);      // :90
        eqs->pushGrid(eqs->index2name(e_index));
        eqs->propagateChanges();
    }
}
