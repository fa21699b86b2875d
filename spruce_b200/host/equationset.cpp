// equationset.cpp -- EquationSet / IdealMHD on the B200 path: registry + config parsing on the host, arithmetic on the device.
#include "equationset.hpp"
#include "plasmadomain.hpp"
#include "utils.hpp"
#include <algorithm>
#include <iostream>

EquationSet::EquationSet(PlasmaDomain &pd, std::vector<std::string> var_names) : m_pd(pd), m_var_names(std::move(var_names))
{
    m_grids.assign(m_var_names.size(), Grid::Zero(1, 1));
    m_output_flags.assign(m_var_names.size(), false);
    for (size_t i = 0; i < m_var_names.size(); i++) m_var_indices[m_var_names[i]] = (int)i;
}

bool EquationSet::isEquationSetName(const std::string &name) { return std::find(m_sets.begin(), m_sets.end(), name) != m_sets.end(); }

std::unique_ptr<EquationSet> EquationSet::instantiateDefault(PlasmaDomain &pd, const std::string &name)
{
    if (name == "ideal_mhd") return std::unique_ptr<EquationSet>(new IdealMHD(pd));
    if (name == "ideal_2F") return std::unique_ptr<EquationSet>(new Ideal2F(pd));
    if (name == "ideal_mhd_2E") return std::unique_ptr<EquationSet>(new IdealMHD2E(pd));
    if (isEquationSetName(name)) spruce_die("Equation set <" + name + "> is not ported to the B200 path yet (ideal_mhd, ideal_mhd_2E and ideal_2F are).");
    spruce_die("Equation set name <" + name + "> not recognized.");
}

// equationset.cpp:38-56: an inactive set's block is skipped, an active one is instantiated and configured from its block
void EquationSet::instantiateWithConfig(std::unique_ptr<EquationSet> &eqs, PlasmaDomain &pd, std::ifstream &in, const std::string &name, bool active)
{
    if (active) {
        eqs = instantiateDefault(pd, name);
        eqs->configureEquationSet(in);
        return;
    }
    std::string line;
    std::getline(in, line); clearWhitespace(line);
    SPRUCE_REQUIRE(!line.empty() && line[0] == '{', "All Equation set activation/deactivation configs must be immediately followed by curly brackets");
    do { if (!std::getline(in, line)) break; clearWhitespace(line); } while (line.empty() || line[0] != '}');
}

// equationset.cpp:64-85 (blank and comment lines inside the block are skipped instead of looping forever)
void EquationSet::configureEquationSet(std::ifstream &in)
{
    std::vector<std::string> lhs_all, rhs_all;
    std::string line, lhs, rhs;
    std::getline(in, line);
    SPRUCE_REQUIRE(!line.empty() && line[0] == '{', "All equation set activation configs must be immediately followed by curly brackets (on their own lines)");
    while (std::getline(in, line)) {
        if (!line.empty() && line[0] == '}') break;
        clearWhitespace(line);
        if (line.empty() || line[0] == '#') continue;
        splitAssignment(line, lhs, rhs);
        lhs_all.push_back(lhs); rhs_all.push_back(rhs);
    }
    parseEquationSetConfigs(lhs_all, rhs_all);
}

bool EquationSet::allStateGridsInitialized() const
{
    for (int i : state_variables()) if (m_grids[i].size() == 1) return false;
    return true;
}

// equationset.cpp:87-104: populateVariablesFromState runs on the device
void EquationSet::setupEquationSet()
{
    SPRUCE_REQUIRE(allStateGridsInitialized(), "All variables specified as state variables for the current EquationSet must be specified in the .state file");
    m_pd.createDevice();
    configureDevice();
    for (int v : state_variables()) PlasmaDomain::check(spruce_grid_upload(m_pd.device(), index2name(v).c_str(), m_pd.slab(m_grids[v]), m_pd.slabCount()));
    if (m_pd.nRanks() > 1) connectSlabs();
    PlasmaDomain::check(spruce_eqs_setup(m_pd.device()));
    if (m_pd.nRanks() > 1) {
        SlabComm::instance().barrier();                     // every rank has mapped its neighbours and finished its local setup
        PlasmaDomain::check(spruce_mgpu_initial_exchange(m_pd.device()));
    }
    name2index("dt");
}

// The host side of the peer-store halo transport (spruce_b200.h "multi-GPU"; spruce_b200/multigpu.py does the same over torch.distributed): the zero-plane
// masks are OR-ed over the ranks, every rank exports one 64-byte CUDA IPC handle and maps all of them.  From then on spruce_advance exchanges halos and the
// dt minimum between the GPUs itself.
void EquationSet::connectSlabs()
{
    SlabComm &comm = SlabComm::instance();
    int local_mask = 0;
    PlasmaDomain::check(spruce_plane_activity(m_pd.device(), &local_mask, -1));
    PlasmaDomain::check(spruce_plane_activity(m_pd.device(), nullptr, comm.allOr(local_mask)));
    unsigned char mine[64] = {0}, all[SlabShared::kMaxRanks * 64];
    int ok = spruce_mgpu_ipc_export(m_pd.device(), mine) == SPRUCE_OK;
    comm.allGather64(mine, all);
    ok = ok && spruce_mgpu_ipc_connect(m_pd.device(), all, comm.nRanks()) == SPRUCE_OK;
    const std::string why = ok ? "" : spruce_last_error();
    if (!comm.allMin(ok)) spruce_die("run -g: the GPUs of this node cannot map each other's memory (CUDA IPC / peer access)" + (why.empty() ? std::string(" -- on another rank") : ": " + why));
}

int EquationSet::name2index(const std::string &name) const
{
    auto it = m_var_indices.find(name);
    if (it == m_var_indices.end()) spruce_die("Variable name <" + name + "> not recognized");
    return it->second;
}
int EquationSet::name2evolvedindex(const std::string &name) const
{
    const std::vector<int> ev = evolved_variables();
    for (size_t i = 0; i < ev.size(); i++) if (index2name(ev[i]) == name) return (int)i;
    spruce_die("<" + name + "> does not correspond to an evolved variable.");
}

Grid &EquationSet::grid(int index)
{
    SPRUCE_REQUIRE(index >= 0 && index < num_variables(), "Grid index must be within range of m_grids");
    Grid &g = m_grids[index];
    if (g.rows() != (int)m_pd.xdim() || g.cols() != (int)m_pd.ydim()) g = Grid(m_pd.xdim(), m_pd.ydim());
    PlasmaDomain::check(spruce_grid_download(m_pd.device(), m_var_names[index].c_str(), m_pd.slab(g), m_pd.slabCount()));
    m_pd.gatherRows(g);                                     // slabs: a collective call -- every rank ends up with the whole plane
    return g;
}
Grid &EquationSet::grid(const std::string &name) { return grid(name2index(name)); }

void EquationSet::pushGrid(const std::string &name)
{
    Grid &g = m_grids[name2index(name)];
    PlasmaDomain::check(spruce_grid_upload(m_pd.device(), name.c_str(), m_pd.slab(g), m_pd.slabCount()));
}

std::vector<Grid> EquationSet::computeTimeDerivatives()
{
    SPRUCE_REQUIRE(m_pd.nRanks() == 1, "computeTimeDerivatives() returns whole planes: one rank only");
    const size_t np = m_pd.xdim() * m_pd.ydim(), ne = evolved_variables().size();
    std::vector<double> k(ne * np);
    PlasmaDomain::check(spruce_eqs_time_derivatives(m_pd.device(), k.data(), k.size()));
    std::vector<Grid> out;
    for (size_t v = 0; v < ne; v++) out.emplace_back(m_pd.xdim(), m_pd.ydim(), std::vector<double>(k.begin() + v * np, k.begin() + (v + 1) * np));
    return out;
}
void EquationSet::propagateChanges() { PlasmaDomain::check(spruce_eqs_propagate_changes(m_pd.device())); }
double EquationSet::nextStepSize()
{
    double s = 0.0;
    PlasmaDomain::check(spruce_next_step_size(m_pd.device(), &s));
    return s;
}

IdealMHD::IdealMHD(PlasmaDomain &pd) : EquationSet(pd, def_var_names()) {}
int IdealMHD::device_id() const { return SPRUCE_EQS_IDEAL_MHD; }

// idealmhd.cpp:12-40: same keys, same bounds checks
void IdealMHD::parseEquationSetConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs)
{
    for (size_t i = 0; i < lhs.size(); i++) {
        const std::string &k = lhs[i];
        if (k == "global_viscosity") { m_global_viscosity = std::stod(rhs[i]); continue; }
        if (k == "viscosity_opt") continue;
        if (k == "moc_b_limiting") { m_moc_b_limiting = (rhs[i] == "true"); continue; }
        if (k == "moc_mom_limiting") { m_moc_mom_limiting = (rhs[i] == "true"); continue; }
        if (k == "moc_b_lower_lim") { m_moc_b_lim[0] = std::stod(rhs[i]); SPRUCE_REQUIRE(m_moc_b_lim[0] <= 1.0, "MoC B field lower threshold should be <=1.0"); continue; }
        if (k == "moc_b_upper_lim") { m_moc_b_lim[1] = std::stod(rhs[i]); SPRUCE_REQUIRE(m_moc_b_lim[1] >= 1.0, "MoC B field upper threshold should be >=1.0"); continue; }
        if (k == "moc_mom_lower_lim") { m_moc_mom_lim[0] = std::stod(rhs[i]); SPRUCE_REQUIRE(m_moc_mom_lim[0] <= 1.0, "MoC momentum threshold should be <=1.0"); continue; }
        if (k == "moc_mom_upper_lim") { m_moc_mom_lim[1] = std::stod(rhs[i]); SPRUCE_REQUIRE(m_moc_mom_lim[1] >= 1.0, "MoC momentum threshold should be >=1.0"); continue; }
        spruce_die(k + " is not recognized for this equation set.");
    }
}

void IdealMHD::configureDevice()
{
    PlasmaDomain::check(spruce_eqs_ideal_mhd_options(m_pd.device(), m_global_viscosity));
    if (m_moc_b_limiting || m_moc_mom_limiting)
        PlasmaDomain::check(spruce_eqs_ideal_mhd_moc_limiting(m_pd.device(), m_moc_b_limiting ? 1 : 0, m_moc_b_lim[0], m_moc_b_lim[1], m_moc_mom_limiting ? 1 : 0,
                                                              m_moc_mom_lim[0], m_moc_mom_lim[1]));
}

IdealMHD2E::IdealMHD2E(PlasmaDomain &pd) : EquationSet(pd, def_var_names()) {}
int IdealMHD2E::device_id() const { return SPRUCE_EQS_IDEAL_MHD_2E; }
// idealmhd2E.cpp:10-20: both keys are parsed and never read by this equation set
void IdealMHD2E::parseEquationSetConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs)
{
    for (size_t i = 0; i < lhs.size(); i++) {
        if (lhs[i] == "global_viscosity") std::stod(rhs[i]);
        else if (lhs[i] == "viscosity_opt") { }
        else spruce_die(lhs[i] + " is not recognized for this equation set.");
    }
}

Ideal2F::Ideal2F(PlasmaDomain &pd) : EquationSet(pd, def_var_names()) {}
int Ideal2F::device_id() const { return SPRUCE_EQS_IDEAL_2F; }

// ideal2F.cpp:10-28
void Ideal2F::parseEquationSetConfigs(std::vector<std::string> lhs, std::vector<std::string> rhs)
{
    for (size_t i = 0; i < lhs.size(); i++) {
        const std::string &k = lhs[i];
        if (k == "use_sub_cycling") m_use_sub_cycling = (rhs[i] == "true");
        else if (k == "epsilon_courant") std::stod(rhs[i]);          // only read by the sub-cycled Maxwell update
        else if (k == "verbose_2F") { }
        else if (k == "remove_curl_terms") m_remove_curl_terms = (rhs[i] == "true");
        else spruce_die(k + " is not recognized for this equation set.");
    }
}
void Ideal2F::configureDevice()
{
    if (m_remove_curl_terms) m_use_sub_cycling = false;              // setupEquationSetDerived, ideal2F.cpp:24-28
    PlasmaDomain::check(spruce_eqs_ideal2f_options(m_pd.device(), m_use_sub_cycling ? 1 : 0, m_remove_curl_terms ? 1 : 0));
}
