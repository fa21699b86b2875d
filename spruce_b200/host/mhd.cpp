// mhd.cpp -- mhdSolve (source/mhd/mhd.cpp:6-31): build a PlasmaDomain and run it.
#include "mhd.hpp"
#include "plasmadomain.hpp"
#include "utils.hpp"

void mhdSolve(const fs::path &prev_run_directory, double time_duration, double cluster_time)
{
    fs::path state_filename, config_filename;
    for (auto const &e : fs::directory_iterator{prev_run_directory}) {
        if (e.path().extension().string() == ".config") {
            SPRUCE_REQUIRE(config_filename.empty(), "There must be only one .config file in the specified directory");
            config_filename = e.path();
        } else if (e.path().filename() == "end.state") state_filename = e.path();
    }
    SPRUCE_REQUIRE(!config_filename.empty() && !state_filename.empty(), "There must be a .state and a .config file in the specified directory");
    PlasmaDomain simulation(prev_run_directory, config_filename, state_filename, true, false);
    simulation.run(time_duration, cluster_time);
}

void mhdSolve(const fs::path &state_filename, const fs::path &config_filename, const fs::path &output_pathname, double time_duration, bool overwrite_init, double cluster_time)
{
    PlasmaDomain simulation(output_pathname, config_filename, state_filename, false, overwrite_init);
    simulation.run(time_duration, cluster_time);
}
