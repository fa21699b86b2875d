// plasmadomain.hpp -- host mirror of the reference's PlasmaDomain driver (source/mhd/plasmadomain.hpp:18-257) on the B200 path.
// Same public surface (constructor, run, readStateFile, readConfigFile), same .state/.config inputs, same mhd.out /
// end.state outputs.  The grids of the equation set live in the device arena behind `m_dev`; the time loop is
// spruce_advance() batched between output iterations.
#pragma once
#include <chrono>
#include <filesystem>
namespace fs = std::filesystem;
#include "../../include/spruce_b200.h"
#include "equationset.hpp"
#include "grid.hpp"
#include "module.hpp"
#include "slabcomm.hpp"
#include "utils.hpp"
#include <string>
#include <vector>

class PlasmaDomain {
public:
    enum class BoundaryCondition { Periodic, Open, Fixed, Reflect, OpenMoC, OpenUCNP };      // plasmadomain.hpp:22
    static inline std::vector<std::string> m_boundary_condition_names = {"periodic", "open", "fixed", "reflect", "open_moc", "open_ucnp"};
    enum class TimeIntegrator { Euler, RK2, RK4 };                                            // plasmadomain.hpp:29
    static inline std::vector<std::string> m_time_integrator_names = {"euler", "rk2", "rk4"};
    enum Grids { d_x, d_y, pos_x, pos_y, be_x, be_y, be_z };                                  // plasmadomain.hpp:35
    static const inline std::vector<std::string> m_gridnames = {"d_x", "d_y", "pos_x", "pos_y", "be_x", "be_y", "be_z"};
    std::vector<Grid> m_grids{m_gridnames.size()};
    static inline std::vector<std::string> m_config_names = {                                 // plasmadomain.hpp:46-51
        "x_bound_1", "x_bound_2", "y_bound_1", "y_bound_2", "epsilon", "density_min", "temp_min", "thermal_energy_min", "max_iterations",
        "iter_output_interval", "time_output_interval", "output_flags", "xdim", "ydim", "open_boundary_strength", "std_out_interval", "write_interval",
        "open_boundary_decay_base", "x_origin", "y_origin", "time_integrator", "duration", "sg_opt", "write_precision", "multispecies_mode"};

    PlasmaDomain(const fs::path &out_path, const fs::path &config_path, const fs::path &state_file, bool continue_mode, bool overwrite_init);
    ~PlasmaDomain();
    void readStateFile(const fs::path &state_file, bool continue_mode = true);
    void readConfigFile(const fs::path &config_file);
    void run(double time_duration, double cluster_time);

    // ---- what the reference exposes to EquationSets / Modules through friend declarations (plasmadomain.hpp:65-91)
    spruce_domain *device() const { return m_dev; }
    size_t xdim() const { return m_xdim; }
    double time() const { return m_time; }                                                     // m_time, plasmadomain.hpp:75
    int iter() const { return m_iter; }                                                        // m_iter
    bool xPeriodic() const { return x_bound_1 == BoundaryCondition::Periodic && x_bound_2 == BoundaryCondition::Periodic; }
    bool yPeriodic() const { return y_bound_1 == BoundaryCondition::Periodic && y_bound_2 == BoundaryCondition::Periodic; }
    size_t ydim() const { return m_ydim; }
    int xl() const { return m_xl; }
    int xu() const { return m_xu; }
    int yl() const { return m_yl; }
    int yu() const { return m_yu; }
    const Grid &ghostZoneMask() const { return m_ghost_zone_mask; }
    EquationSet *eqs() const { return m_eqs.get(); }
    double ionMass() const { return m_ion_mass; }
    double adiabaticIndex() const { return m_adiabatic_index; }
    // ---- differential operators for host-side modules that are not ported (plasmadomain.hpp:194-242): a host Grid in, a host Grid out, the
    //      evaluation on the device through spruce_operator / spruce_operator2 (bit-identical to derivs.cpp over the iteration bounds)
    Grid derivative1D(const Grid &quantity, int index) const { return deviceOperator("derivative1D", index, quantity, nullptr); }                 // derivs.cpp:223
    Grid secondDerivative1D(const Grid &quantity, int index) const { return deviceOperator("secondDerivative1D", index, quantity, nullptr); }     // derivs.cpp:417
    Grid laplacian(const Grid &quantity) const { return deviceOperator("laplacian", 0, quantity, nullptr); }                                      // derivs.cpp:458
    Grid transportDerivative1D(const Grid &quantity, const Grid &vel, int index) const { return deviceOperator("transportDerivative1D", index, quantity, &vel); }   // derivs.cpp:122
    Grid divergence2D(const Grid &a_x, const Grid &a_y) const { return deviceOperator2("divergence2D", a_x, a_y, nullptr); }                       // derivs.cpp:407
    Grid divergence2D(const std::vector<Grid> &a) const { return divergence2D(a.at(0), a.at(1)); }
    Grid curl2D(const Grid &x, const Grid &y) const { return deviceOperator2("curl2D", x, y, nullptr); }                                          // derivs.cpp:472
    Grid transportDivergence2D(const Grid &quantity, const std::vector<Grid> &vel) const { return deviceOperator2("transportDivergence2D", quantity, vel.at(0), &vel.at(1)); }   // derivs.cpp:216
    std::vector<Grid> curlZ(const Grid &z) const                                                                                               // derivs.cpp:465
    {
        Grid ry = derivative1D(z, 0);
        for (double &v : ry.data()) v = -v;
        return {derivative1D(z, 1), ry};
    }
    // ---- slab decomposition (`run -g N`, slabcomm.hpp): this rank owns rows [row0, row0 + nxLocal) of every plane; host Grids stay global
    int rank() const { return SlabComm::instance().rank(); }
    int nRanks() const { return SlabComm::instance().nRanks(); }
    int row0() const { return m_row0; }
    int nxLocal() const { return m_nx_local; }
    size_t slabCount() const { return (size_t)m_nx_local * m_ydim; }
    const double *slab(const Grid &g) const { return g.ptr() + (size_t)m_row0 * m_ydim; }
    double *slab(Grid &g) const { return g.ptr() + (size_t)m_row0 * m_ydim; }
    void gatherRows(Grid &g) const { SlabComm::instance().allGatherRows(g.ptr(), m_ydim, m_row0, m_nx_local, (size_t)g.size()); }   // every rank ends up with the whole plane
    void createDevice();                 // called by EquationSet::setupEquationSet once config + state are known
    static void check(int rc);           // non-zero status -> message on stderr + abort (the reference's assert style)

private:
    using clock_type = std::chrono::steady_clock;
    std::chrono::time_point<clock_type> m_start{clock_type::now()};
    double elapsed() const { return std::chrono::duration<double>(clock_type::now() - m_start).count(); }

    TimeIntegrator m_time_integrator{TimeIntegrator::Euler};
    double m_time{0.0}, m_duration{-1.0}, m_max_time{-1.0};
    int m_iter{0}, max_iterations{100};
    bool m_overwrite_init{false}, m_continue_mode{false};
    fs::path m_out_directory, m_out_filename{"mhd.out"};
    std::vector<std::string> m_comment_lines;
    int m_iter_output_interval{1}, m_std_out_interval{1}, m_write_precision{4}, m_write_interval{1}, m_store_counter{0}, m_state_identifier{1};
    double m_time_output_interval{-1.0};
    std::vector<std::string> m_data_to_write;
    int m_xl{0}, m_xu{0}, m_yl{0}, m_yu{0};
    Grid m_ghost_zone_mask;
    BoundaryCondition x_bound_1{BoundaryCondition::Periodic}, x_bound_2{BoundaryCondition::Periodic}, y_bound_1{BoundaryCondition::Periodic}, y_bound_2{BoundaryCondition::Periodic};
    double open_boundary_strength{0.0}, open_boundary_decay_base{1.0};
    size_t m_xdim{0}, m_ydim{0};
    int m_row0{0}, m_nx_local{0};
    bool m_multispecies_mode{false};
    std::string m_sg_opt;
    double m_ion_mass{0.0}, m_adiabatic_index{0.0};
    double epsilon{0.0}, density_min{0.0}, temp_min{0.0}, thermal_energy_min{0.0};
    ModuleHandler m_module_handler;
    std::unique_ptr<EquationSet> m_eqs;
    spruce_domain *m_dev{nullptr};

    Grid deviceOperator(const char *op, int index, const Grid &q, const Grid *vel) const
    {
        SPRUCE_REQUIRE(nRanks() == 1, "the operator members work on whole planes: one rank only (spruce_b200.h)");
        Grid out(q.rows(), q.cols());
        check(spruce_operator(m_dev, op, index, q.ptr(), vel ? vel->ptr() : nullptr, out.ptr(), (size_t)q.size()));
        return out;
    }
    Grid deviceOperator2(const char *op, const Grid &a, const Grid &b, const Grid *c) const
    {
        SPRUCE_REQUIRE(nRanks() == 1, "the operator members work on whole planes: one rank only (spruce_b200.h)");
        Grid out(a.rows(), a.cols());
        check(spruce_operator2(m_dev, op, a.ptr(), b.ptr(), c ? c->ptr() : nullptr, out.ptr(), (size_t)a.size()));
        return out;
    }
    void computeIterationBounds();
    void outputPreamble();
    void storeGrids();
    void writeToOutFile();
    void writeStateFile(const std::string &filename_stem = "mhd", int precision = -1);
    void updateStateIdentifier();
    void printUpdate(int iter, double time, double dt) const;
    void handleSingleConfig(int setting_index, const std::string &rhs);
    BoundaryCondition stringToBoundaryCondition(const std::string &str) const;
    TimeIntegrator stringToTimeIntegrator(const std::string &str) const;
    static std::string num2str(double num);
    friend class ModuleHandler;
    friend class EICThermalization;
    friend class PhysicalViscosity;
    friend class TracerParticles;
    friend class CoulombExplosion;
    friend class GlobalTemperature;
};
