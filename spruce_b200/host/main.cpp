// main.cpp -- command line of the B200-backed solver; same flags as the reference's main.cpp:16-91:
//   run -m {input,continue} -o <dir> [-s state] [-c config] [-d duration] [-r walltime]
#include "mhd.hpp"
#include "utils.hpp"
#include <iostream>

int main(int argc, char *argv[])
{
    const std::string run_mode = getCommandLineArg(argc, argv, "-m", "--mode");
    const std::string dur = getCommandLineArg(argc, argv, "-d", "--duration");
    const double time_duration = dur.empty() ? -1.0 : std::stod(dur);
    const std::string rt = getCommandLineArg(argc, argv, "-r", "--runtime");
    const double cluster_time = rt.empty() ? -1.0 : std::stod(rt);
    const fs::path out_path(getCommandLineArg(argc, argv, "-o", "--output"));
    SPRUCE_REQUIRE(!out_path.empty(), "output directory must be specified");

    if (run_mode == "continue") {
        SPRUCE_REQUIRE(!dur.empty(), "In Continue Mode, duration of simulation must be specified on command line");
        SPRUCE_REQUIRE(fs::is_directory(out_path), "Given output directory of previous run must be existing directory");
        std::cout << "Running in Continue Mode for " << time_duration << " s...\n";
        mhdSolve(out_path, time_duration, cluster_time);
        return 0;
    }
    if (run_mode != "input") { std::cerr << "Mode '" << run_mode << "' not recognized\n"; return 1; }
    fs::path config_path(getCommandLineArg(argc, argv, "-c", "--config")), grid_path(getCommandLineArg(argc, argv, "-s", "--state"));
    const bool seek_config = config_path.empty(), seek_grids = grid_path.empty();
    if (seek_config || seek_grids) {
        for (auto const &e : fs::directory_iterator{out_path}) {
            if (seek_config && e.path().extension().string() == ".config") {
                SPRUCE_REQUIRE(config_path.empty(), "There must be only one .config file in the specified directory");
                config_path = e.path();
                std::cout << "Found configuration file " << config_path << std::endl;
            } else if (seek_grids && e.path().filename() == "init.state") {
                grid_path = e.path();
                std::cout << "Found initializing state file " << grid_path << std::endl;
            }
        }
    }
    SPRUCE_REQUIRE(!config_path.empty(), "Config file not specified and not found in output directory");
    SPRUCE_REQUIRE(!grid_path.empty(), "Initializing state file not specified and not found in output directory");
    fs::create_directories(out_path);
    if (grid_path.extension().string() != ".state") { std::cerr << "Grids must be specified in .state file.\n"; return 1; }
    SPRUCE_REQUIRE(fs::is_regular_file(grid_path), "Given state file must exist and be a file");
    if (dur.empty()) std::cout << "Running in Input Mode from the state file " << grid_path.string() << " for duration specified in .config file." << std::endl;
    else std::cout << "Running in Input Mode from the state file " << grid_path.string() << " for " << time_duration << " s...\n";
    mhdSolve(grid_path, config_path, out_path, time_duration, !seek_grids, cluster_time);
    return 0;
}
