// equationset.hpp -- host mirror of the reference's EquationSet plug-in interface (source/equationsets/equationset.hpp:18-171)
// for the B200 path: the variable registry, name/index maps and output flags stay here; every arithmetic member
// (computeTimeDerivatives, applyTimeDerivatives, propagateChanges, getDT) is a call into libspruce_b200.so.
#pragma once
#include "grid.hpp"
#include <fstream>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

class PlasmaDomain;

class EquationSet {
public:
    static const inline std::vector<std::string> m_sets{"ideal_mhd", "ideal_mhd_cons", "ideal_mhd_2E", "ideal_2F"};   // equationset.hpp:22
    static bool isEquationSetName(const std::string &name);
    static std::unique_ptr<EquationSet> instantiateDefault(PlasmaDomain &pd, const std::string &name);
    static void instantiateWithConfig(std::unique_ptr<EquationSet> &eqs, PlasmaDomain &pd, std::ifstream &in, const std::string &name, bool active);

    EquationSet(PlasmaDomain &pd, std::vector<std::string> var_names);
    virtual ~EquationSet() {}
    void configureEquationSet(std::ifstream &in);
    void setupEquationSet();                       // uploads the state variables, runs populateVariablesFromState on the device

    virtual int device_id() const = 0;            // SPRUCE_EQS_*
    virtual void configureDevice() {}             // hand the parsed equation-set block to the device (after createDevice)
    virtual std::vector<int> state_variables() const = 0;
    virtual std::vector<int> evolved_variables() const = 0;
    virtual std::vector<std::string> species() const = 0;
    virtual std::vector<int> densities() const = 0;
    virtual std::vector<int> number_densities() const = 0;
    virtual std::vector<std::vector<int>> momenta() const = 0;
    virtual std::vector<std::vector<int>> velocities() const = 0;
    virtual std::vector<int> thermal_energies() const = 0;
    virtual std::vector<int> pressures() const = 0;
    virtual std::vector<int> temperatures() const = 0;
    virtual std::vector<int> fields() const = 0;
    virtual std::vector<int> timescale() const = 0;

    Grid &grid(int index);                         // refreshed from the device (derived variables are materialised on demand)
    Grid &grid(const std::string &name);
    Grid &hostGrid(int index) { return m_grids[index]; }   // host staging copy as read from the .state file
    void pushGrid(const std::string &name);        // host staging copy -> device (a host-side module edited an evolved plane)
    void setOutputFlag(int index, bool flag) { m_output_flags[index] = flag; }
    void setOutputFlag(const std::string &name, bool flag) { m_output_flags[name2index(name)] = flag; }
    bool getOutputFlag(int index) const { return m_output_flags[index]; }

    std::vector<std::string> allNames() const { return m_var_names; }
    std::string index2name(int index) const { return m_var_names[index]; }
    int name2index(const std::string &name) const;
    int name2evolvedindex(const std::string &name) const;
    bool is_var(const std::string &name) const { return m_var_indices.count(name) != 0; }
    int num_variables() const { return (int)m_var_names.size(); }
    int num_species() const { return (int)species().size(); }
    bool allStateGridsInitialized() const;

    std::vector<Grid> computeTimeDerivatives();    // one RHS evaluation on the primary state (equationset.cpp:204-210)
    void propagateChanges();                       // equationset.cpp:212-220 on the device
    double nextStepSize();                         // epsilon * getDT().min(...)  (evolution.cpp:62)
    void connectSlabs();                           // run -g N: zero-plane masks, CUDA IPC handles (slabcomm.hpp)

protected:
    PlasmaDomain &m_pd;
    std::vector<Grid> m_grids;
    const std::vector<std::string> m_var_names;
    std::unordered_map<std::string, int> m_var_indices;
    std::vector<bool> m_output_flags;
    virtual void parseEquationSetConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs) = 0;
};

// source/equationsets/idealmhd.hpp:14-67
class IdealMHD : public EquationSet {
public:
    explicit IdealMHD(PlasmaDomain &pd);
    enum Vars { rho, temp, mom_x, mom_y, mom_z, bi_x, bi_y, bi_z, grav_x, grav_y, n, press, thermal_energy, v_x, v_y, v_z, kinetic_energy,
                b_x, b_y, b_z, b_mag, b_hat_x, b_hat_y, b_hat_z, dt };
    static std::vector<std::string> def_var_names()
    {
        return {"rho", "temp", "mom_x", "mom_y", "mom_z", "bi_x", "bi_y", "bi_z", "grav_x", "grav_y", "n", "press", "thermal_energy", "v_x", "v_y", "v_z",
                "kinetic_energy", "b_x", "b_y", "b_z", "b_mag", "b_hat_x", "b_hat_y", "b_hat_z", "dt"};
    }
    int device_id() const override;
    std::vector<int> state_variables() const override { return {rho, temp, mom_x, mom_y, mom_z, bi_x, bi_y, bi_z, grav_x, grav_y}; }
    std::vector<int> evolved_variables() const override { return {rho, mom_x, mom_y, mom_z, thermal_energy, bi_x, bi_y, bi_z}; }
    std::vector<std::string> species() const override { return {"i"}; }
    std::vector<int> densities() const override { return {rho}; }
    std::vector<int> number_densities() const override { return {n}; }
    std::vector<std::vector<int>> momenta() const override { return {{mom_x, mom_y, mom_z}}; }
    std::vector<std::vector<int>> velocities() const override { return {{v_x, v_y, v_z}}; }
    std::vector<int> thermal_energies() const override { return {thermal_energy}; }
    std::vector<int> pressures() const override { return {press}; }
    std::vector<int> temperatures() const override { return {temp}; }
    std::vector<int> fields() const override { return {bi_x, bi_y, bi_z}; }
    std::vector<int> timescale() const override { return {dt}; }

    void configureDevice() override;
private:
    double m_global_viscosity = 0.0;              // idealmhd.hpp:48; read by the open_moc boundary only (idealmhd.cpp:90)
    bool m_moc_b_limiting = false, m_moc_mom_limiting = false;          // idealmhd.hpp:59-64
    double m_moc_b_lim[2] = {0.1, 10.0}, m_moc_mom_lim[2] = {0.1, 10.0};
    void parseEquationSetConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs) override;
};

// source/equationsets/idealmhd2E.hpp:12-58 -- one fluid with separate ion / electron thermal energies
class IdealMHD2E : public EquationSet {
public:
    explicit IdealMHD2E(PlasmaDomain &pd);
    enum Vars { rho, i_temp, e_temp, mom_x, mom_y, bi_x, bi_y, grav_x, grav_y, n, i_press, e_press, press, i_thermal_energy, e_thermal_energy, v_x, v_y, kinetic_energy,
                b_x, b_y, b_mag, b_hat_x, b_hat_y, dt };
    static std::vector<std::string> def_var_names()
    {
        return {"rho", "i_temp", "e_temp", "mom_x", "mom_y", "bi_x", "bi_y", "grav_x", "grav_y", "n", "i_press", "e_press", "press", "i_thermal_energy", "e_thermal_energy",
                "v_x", "v_y", "kinetic_energy", "b_x", "b_y", "b_mag", "b_hat_x", "b_hat_y", "dt"};
    }
    int device_id() const override;
    std::vector<int> state_variables() const override { return {rho, i_temp, e_temp, mom_x, mom_y, bi_x, bi_y, grav_x, grav_y}; }
    std::vector<int> evolved_variables() const override { return {rho, mom_x, mom_y, i_thermal_energy, e_thermal_energy, bi_x, bi_y}; }
    std::vector<std::string> species() const override { return {"i", "e"}; }
    std::vector<int> densities() const override { return {rho, rho}; }
    std::vector<int> number_densities() const override { return {n, n}; }
    std::vector<std::vector<int>> momenta() const override { return {{mom_x, mom_y}, {mom_x, mom_y}}; }
    std::vector<std::vector<int>> velocities() const override { return {{v_x, v_y}, {v_x, v_y}}; }
    std::vector<int> thermal_energies() const override { return {i_thermal_energy, e_thermal_energy}; }
    std::vector<int> pressures() const override { return {i_press, e_press}; }
    std::vector<int> temperatures() const override { return {i_temp, e_temp}; }
    std::vector<int> fields() const override { return {bi_x, bi_y}; }
    std::vector<int> timescale() const override { return {dt, dt}; }
private:
    void parseEquationSetConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs) override;
};

// source/equationsets/ideal2F.hpp:16-77 -- ion + electron fluids with Maxwell's equations (the non-sub-cycled update)
class Ideal2F : public EquationSet {
public:
    explicit Ideal2F(PlasmaDomain &pd);
    enum Vars { i_rho, e_rho, i_mom_x, i_mom_y, e_mom_x, e_mom_y, i_temp, e_temp, bi_x, bi_y, bi_z, E_x, E_y, E_z, grav_x, grav_y,
                i_n, e_n, i_v_x, i_v_y, e_v_x, e_v_y, j_x, j_y, i_press, e_press, press, i_thermal_energy, e_thermal_energy,
                rho, rho_c, n, dn, dt, dt_i, b_x, b_y, b_z, b_mag, b_mag_xy, b_hat_x, b_hat_y, curlE_z, divE, divB, i_dPdx, e_dPdx };
    static std::vector<std::string> def_var_names()
    {
        return {"i_rho", "e_rho", "i_mom_x", "i_mom_y", "e_mom_x", "e_mom_y", "i_temp", "e_temp", "bi_x", "bi_y", "bi_z", "E_x", "E_y", "E_z", "grav_x", "grav_y",
                "i_n", "e_n", "i_v_x", "i_v_y", "e_v_x", "e_v_y", "j_x", "j_y", "i_press", "e_press", "press", "i_thermal_energy", "e_thermal_energy",
                "rho", "rho_c", "n", "dn", "dt", "dt_i", "b_x", "b_y", "b_z", "b_mag", "b_mag_xy", "b_hat_x", "b_hat_y", "curlE_z", "divE", "divB", "i_dPdx", "e_dPdx"};
    }
    int device_id() const override;
    void configureDevice() override;
    std::vector<int> state_variables() const override { return {i_rho, e_rho, i_mom_x, i_mom_y, e_mom_x, e_mom_y, i_temp, e_temp, bi_x, bi_y, bi_z, E_x, E_y, E_z, grav_x, grav_y}; }
    std::vector<int> evolved_variables() const override { return {i_rho, e_rho, i_mom_x, i_mom_y, e_mom_x, e_mom_y, i_thermal_energy, e_thermal_energy, E_x, E_y, E_z, bi_x, bi_y, bi_z}; }
    std::vector<std::string> species() const override { return {"i", "e"}; }
    std::vector<int> densities() const override { return {i_rho, e_rho}; }
    std::vector<int> number_densities() const override { return {i_n, e_n}; }
    std::vector<std::vector<int>> momenta() const override { return {{i_mom_x, i_mom_y}, {e_mom_x, e_mom_y}}; }
    std::vector<std::vector<int>> velocities() const override { return {{i_v_x, i_v_y}, {e_v_x, e_v_y}}; }
    std::vector<int> thermal_energies() const override { return {i_thermal_energy, e_thermal_energy}; }
    std::vector<int> pressures() const override { return {i_press, e_press}; }
    std::vector<int> temperatures() const override { return {i_temp, e_temp}; }
    std::vector<int> fields() const override { return {E_x, E_y, E_z}; }
    std::vector<int> timescale() const override { return {dt_i, dt}; }

private:
    bool m_use_sub_cycling = true, m_remove_curl_terms = false;       // ideal2F.hpp:62-65
    void parseEquationSetConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs) override;
};
