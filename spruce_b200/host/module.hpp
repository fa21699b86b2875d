// module.hpp -- host mirror of the reference's Module plug-in interface and ModuleHandler
// (source/modules/module.hpp:15-58, modulehandler.hpp:17-39).  A ported module parses its own config block exactly as
// in the reference and registers itself on the device in setupModule(); its per-step hooks then run inside
// spruce_advance in config order.  The hook virtuals stay for host-side modules that are not ported.
#pragma once
#include "grid.hpp"
#include <filesystem>
#include <fstream>
namespace fs = std::filesystem;
#include <memory>
#include <string>
#include <vector>

class PlasmaDomain;

class Module {
public:
    explicit Module(PlasmaDomain &pd) : m_pd(pd) {}
    virtual ~Module() {}
    void configureModule(std::ifstream &in);
    virtual void setupModule() {}
    virtual void iterateModule(double) {}
    virtual void preIterateModule(double) {}
    virtual void postIterateModule(double) {}
    virtual void computeTimeDerivativesModule(const std::vector<Grid> &, std::vector<Grid> &) {}
    virtual void preRecomputeDerivedModule(std::vector<Grid> &) const {}
    virtual std::string commandLineMessage() const { return ""; }
    virtual void fileOutput(std::vector<std::string> &, std::vector<Grid> &) {}
    virtual std::vector<std::string> config_names() const { return {}; }
    bool ms_given = false;               // the config block sets ms_electron_heating_fraction (multispecies_mode)
    virtual bool device_resident() const { return false; }   // true: hooks run on the GPU inside spruce_advance

protected:
    PlasmaDomain &m_pd;
    virtual void parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs) = 0;
};

class ModuleHandler {
public:
    explicit ModuleHandler(PlasmaDomain &pd) : m_pd(pd) {}
    void setupModules();
    void instantiateModule(const std::string &name, std::ifstream &in, bool active = true);
    bool isModuleName(const std::string &name) const;
    std::vector<std::string> getCommandLineMessages() const;
    void getFileOutputData(std::vector<std::string> &names, std::vector<Grid> &grids) const;
    bool empty() const { return m_modules.empty(); }
    // host-resident modules (device_resident() == false) keep their hooks on the host: PlasmaDomain::run then advances one step per
    // spruce_advance call and calls them around it, in config order (modulehandler.cpp:42-61)
    bool hasHostModules() const;
    void preIterateModules(double dt);
    void iterateModules(double dt);
    void postIterateModules(double dt);

private:
    PlasmaDomain &m_pd;
    std::vector<std::unique_ptr<Module>> m_modules;
    static inline std::vector<std::string> m_module_names = {   // modulehandler.hpp:34-38
        "radiative_losses", "thermal_conduction", "ambient_heating", "anomalous_resistivity", "momentum_injection", "localized_heating",
        "field_heating", "tracer_particles", "sg_filtering", "coulomb_explosion", "eic_thermalization", "artificial_viscosity", "global_temperature",
        "mass_injection", "ambient_heating_sink", "physical_viscosity", "div_cleaning", "boundary_outflow"};
};

// source/modules/viscosity.hpp ("artificial_viscosity")
class Viscosity : public Module {
public:
    explicit Viscosity(PlasmaDomain &pd) : Module(pd) {}
    void setupModule() override;
    void fileOutput(std::vector<std::string> &names, std::vector<Grid> &grids) override;      // viscosity.cpp:351-376
    bool device_resident() const override { return true; }
private:
    std::string m_inp_visc_opt, m_inp_strength, m_inp_vars_to_diff, m_inp_vars_to_evol, m_inp_length, m_inp_species;
    std::string m_hv_time_integrator, m_boundary_falloff_shape;
    std::vector<std::string> m_vars_to_evol;
    bool m_gradient_correction = false, m_output_visc = false, m_output_lap = false, m_output_strength = false, m_output_timescale = false;
    double m_hv_epsilon = 1.0;
    Grid getBoundaryViscosity(double strength, double length) const;
    void parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs) override;
};
// source/modules/solar/thermalconduction.hpp
class ThermalConduction : public Module {
public:
    explicit ThermalConduction(PlasmaDomain &pd) : Module(pd) {}
    void setupModule() override;
    std::string commandLineMessage() const override;
    void fileOutput(std::vector<std::string> &names, std::vector<Grid> &grids) override;      // thermalconduction.cpp:226-237
    bool device_resident() const override { return true; }
private:
    bool flux_saturation = false, output_to_file = false, inactive_mode = false;
    double epsilon = 0.0, dt_subcycle_min = 0.0, weakening_factor = 1.0, ms_electron_heating_fraction = 1.0;      // thermalconduction.hpp:34
    std::string time_integrator;
    void parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs) override;
};
// source/modules/solar/radiativelosses.hpp
class RadiativeLosses : public Module {
public:
    explicit RadiativeLosses(PlasmaDomain &pd) : Module(pd) {}
    void setupModule() override;
    std::string commandLineMessage() const override;
    void fileOutput(std::vector<std::string> &names, std::vector<Grid> &grids) override;      // radiativelosses.cpp:172-179
    bool device_resident() const override { return true; }
private:
    double cutoff_ramp = 0.0, cutoff_temp = 0.0, epsilon = 0.0, ms_electron_heating_fraction = 1.0;      // radiativelosses.hpp:31
    bool output_to_file = false, inactive_mode = false, prevent_subcycling = false;
    std::string time_integrator;
    void parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs) override;
};
// source/modules/solar/ambientheating.hpp
class AmbientHeating : public Module {
public:
    explicit AmbientHeating(PlasmaDomain &pd) : Module(pd) {}
    void setupModule() override;
    std::string commandLineMessage() const override { return exp_mode ? "Ambient Heating On (Exp. Mode)" : "Ambient Heating On"; }
    bool device_resident() const override { return true; }
private:
    double heating_rate = 0.0, exp_base_heating_rate = 0.0, exp_scale_height = 1.0, split_exp_scale_height = 1.0, split_exp_start_height = 0.0;
    double ms_electron_heating_fraction = 0.5;      // ambientheating.hpp:28
    bool exp_mode = false, split_exp_mode = false;
    void parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs) override;
};
// ---- pointwise solar source terms (post-iterate hooks); the device library builds the Gaussian templates from these keys
// source/modules/solar/ambientheatingsink.hpp
class AmbientHeatingSink : public Module {
public:
    explicit AmbientHeatingSink(PlasmaDomain &pd) : Module(pd) {}
    void setupModule() override;
    std::string commandLineMessage() const override { return exp_mode ? "Ambient Heating On (Exp. Mode)" : "Ambient Heating On"; }   // ambientheatingsink.cpp:44-47
    bool device_resident() const override { return true; }
private:
    double heating_rate = 0.0, exp_base_heating_rate = 0.0, exp_scale_height = 1.0, center_x = 0.0, half_width = 1.0, ms_electron_heating_fraction = 0.5;
    bool exp_mode = false;
    void parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs) override;
};
// source/modules/solar/localizedheating.hpp, massinjection.hpp, momentuminjection.hpp share their window / Gaussian keys
class GaussianSource : public Module {
public:
    enum Kind { Heating, Mass, Momentum };
    GaussianSource(PlasmaDomain &pd, Kind kind) : Module(pd), m_kind(kind) {}
    void setupModule() override;
    std::string commandLineMessage() const override;
    bool device_resident() const override { return true; }
private:
    Kind m_kind;
    double start_time = 0.0, duration = 0.0, peak = 0.0, stddev_x = 1.0, stddev_y = 1.0, center_x = 0.0, center_y = 0.0, ramp_time = 0.0;
    double dir_x = 0.0, dir_y = 0.0, template_angle = 0.0, oscillation_period = 1.0, ms_electron_heating_fraction = 0.0;
    bool oscillatory = false;
    void parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs) override;
};
// source/modules/solar/divcleaning.hpp
class DivCleaning : public Module {
public:
    explicit DivCleaning(PlasmaDomain &pd) : Module(pd) {}
    void setupModule() override;
    std::string commandLineMessage() const override { return "Div. Cleaning On"; }
    bool device_resident() const override { return true; }
private:
    double epsilon = 0.1, time_scale = 1.0;
    void parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs) override;
};
// source/modules/solar/fieldheating.hpp
class FieldHeating : public Module {
public:
    explicit FieldHeating(PlasmaDomain &pd) : Module(pd) {}
    void setupModule() override;
    std::string commandLineMessage() const override;
    void fileOutput(std::vector<std::string> &var_names, std::vector<Grid> &var_grids) override;
    bool device_resident() const override { return true; }
private:
    double coeff = 0.0, current_pow = 0.0, b_pow = 0.0, n_pow = 0.0, roc_pow = 0.0;
    bool inactive_mode = false, output_to_file = false;
    void parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs) override;
};
// source/modules/solar/boundaryoutflow.hpp
class BoundaryOutflow : public Module {
public:
    explicit BoundaryOutflow(PlasmaDomain &pd) : Module(pd) {}
    void setupModule() override;
    std::string commandLineMessage() const override;
    bool device_resident() const override { return true; }
private:
    double max_accel = 0.0, falloff_length = 1.0, feather_length = 0.0, dynamic_time = 1.0, dynamic_target_speed = 0.0;
    std::string boundary = "y_bound_2", falloff_shape = "exp";
    bool field_aligned_mode = false, dynamic_mode = false;
    void parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs) override;
};
// source/modules/solar/anomalousresistivity.hpp ("anomalous_resistivity")
class AnomalousResistivity : public Module {
public:
    explicit AnomalousResistivity(PlasmaDomain &pd) : Module(pd) {}
    void setupModule() override;
    std::string commandLineMessage() const override;
    void fileOutput(std::vector<std::string> &var_names, std::vector<Grid> &var_grids) override;
    bool device_resident() const override { return true; }
private:
    // defaults: anomalousresistivity.hpp:16-39
    double time_scale = 1.0, frobenius_metric_coeff = 1.0e50, smoothing_sigma = 3.0, safety_factor = 1.0, flood_fill_max_radius = -1.0, flood_fill_argmin_radius = 5.0e9,
           flood_fill_min_current = -1.0, flood_fill_current_ramp_length = 1.0e-5, flood_fill_threshold = 1.0;
    bool metric_smoothing = true, gradient_correction = false, output_to_file = false;
    std::string time_integrator = "", template_mode = "flood_fill", resistivity_model = "time_scale";
    std::vector<double> resistivity_model_params;
    void parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs) override;
};
// source/modules/ucnp/eic_thermalization.hpp -- electron-ion collisional energy exchange (two-fluid equation set only)
class EICThermalization : public Module {
public:
    explicit EICThermalization(PlasmaDomain &pd) : Module(pd) {}
    void setupModule() override;
    bool device_resident() const override { return true; }
private:
    void parseModuleConfigs(std::vector<std::string>, std::vector<std::string>) override {}   // eic_thermalization.cpp:7-10: no keys
};
// source/modules/solar/physicalviscosity.hpp ("physical_viscosity")
class PhysicalViscosity : public Module {
public:
    explicit PhysicalViscosity(PlasmaDomain &pd) : Module(pd) {}
    void setupModule() override;
    std::string commandLineMessage() const override;
    void fileOutput(std::vector<std::string> &names, std::vector<Grid> &grids) override;
    bool device_resident() const override { return true; }
private:
    double coeff = 0.0, ramp_length = 0.0, buffer_length = 0.0, epsilon = 1.0, ms_electron_heating_fraction = 0.0;
    bool heating_on = true, force_on = true, output_to_file = false, inactive_mode = false, gradient_correction = false;
    std::string time_integrator;
    Grid constructCoefficientGrid(double strength, double ramp_length, double buffer_length) const;
    void parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs) override;
};
// source/modules/sgfilter.hpp ("sg_filtering"): a HOST-resident module -- its post-iterate hook works on host Grids staged through the C ABI
// (download rho / thermal_energy, filter, upload, propagateChanges), the pattern every un-ported module can use
class SGFilter : public Module {
public:
    explicit SGFilter(PlasmaDomain &pd) : Module(pd) {}
    void postIterateModule(double dt) override;                                        // sgfilter.cpp:19-23
    std::string commandLineMessage() const override { return "SG Filtering On"; }      // sgfilter.cpp:25-28
    // singleVarSavitzkyGolay (sgfilter.cpp:46-82), the reference's index quirk included: the 5 x 5 window reads grid(j[v], j[v])
    static void singleVarSavitzkyGolay(Grid &grid, int xl, int xu, int yl, int yu, bool x_periodic, bool y_periodic);
private:
    int filter_interval = 0;
    void applyFilter();
    void parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs) override;
};
// source/modules/solar/tracerparticles.hpp ("tracer_particles"): HOST-resident -- Lagrangian point particles advected with the flow by a midpoint step
// per iteration; reads v_x / v_y staged from the device, never edits the state; init.tpstate in, particles.tpout / end.tpstate out
class TracerParticles : public Module {
public:
    explicit TracerParticles(PlasmaDomain &pd) : Module(pd) {}
    void setupModule() override;                                                        // tracerparticles.cpp:16-41
    void iterateModule(double dt) override;                                             // tracerparticles.cpp:54-96
    std::string commandLineMessage() const override { return "Tracer Particles On"; }
private:
    fs::path m_out_filename{"particles.tpout"}, m_init_filename{"init.tpstate"}, m_end_filename{"end.tpstate"};
    std::vector<double> x_vec, y_vec;
    std::vector<std::vector<double>> m_particles;
    std::vector<std::string> m_labels;
    void parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs) override;
    void readTPStateFile(const fs::path &init_path);
    void writeTPStateFile();
    void writeToTPOutFile(double dt);
};
// source/modules/ucnp/coulomb_explosion.hpp ("coulomb_explosion"): HOST-resident -- once per step, a radial histogram of n, the enclosed charge of a
// decaying non-neutrality and the resulting force on mom_x / mom_y (host/ucnp_modules.hpp); planes staged through the C ABI, then propagateChanges
class CoulombExplosion : public Module {
public:
    explicit CoulombExplosion(PlasmaDomain &pd) : Module(pd) {}
    void setupModule() override;                                                        // coulomb_explosion.cpp:18-32
    void postIterateModule(double dt) override;                                         // coulomb_explosion.cpp:49-87
    std::string commandLineMessage() const override { return "Coulomb Explosion On"; }
    void fileOutput(std::vector<std::string> &var_names, std::vector<Grid> &var_grids) override;
    std::vector<std::string> config_names() const override { return {"timescale", "lengthscale", "strength"}; }
private:
    double m_timescale = 0.0, m_lengthscale = 0.0, m_strength = 0.0;
    bool output_to_file = false;
    enum Vars { F_x, F_y, dP_x, dP_y, num_vars };
    std::vector<std::string> m_var_names{"F_x", "F_y", "dP_x", "dP_y"};
    std::vector<Grid> m_vars;
    void parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs) override;
};
// source/modules/ucnp/global_temperature.hpp ("global_temperature"): HOST-resident -- gt_use_diffusion: sub-cycled midpoint diffusion of each listed species'
// temperature after the step (laplacian through the device operator), thermal energy rebuilt from it.  gt_use_global_temp (a domain integral inside EVERY
// propagateChanges, global_temperature.cpp:45-64) is refused: the device fuses propagateChanges into its stage kernels.
class GlobalTemperature : public Module {
public:
    explicit GlobalTemperature(PlasmaDomain &pd) : Module(pd) {}
    void setupModule() override;                                                        // global_temperature.cpp:17-43
    void postIterateModule(double dt) override;                                         // global_temperature.cpp:66-94
    std::vector<std::string> config_names() const override { return {"gt_species", "gt_strength", "gt_use_diffusion", "gt_use_global_temp"}; }
private:
    bool m_use_diffusion = false, m_use_global_temp = false;
    std::vector<std::string> m_species;
    std::vector<int> m_species_ind;
    double m_strength = 0.0;
    Grid m_dr;
    void parseModuleConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs) override;
};
