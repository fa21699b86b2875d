// main.cpp -- command line of the B200-backed solver; same flags as the reference's main.cpp:16-91:
//   run -m {input,continue} -o <dir> [-s state] [-c config] [-d duration] [-r walltime]
// plus one of its own:  -g N / --gpus N (or SPRUCE_GPUS=N): slab-decompose the grid along x over N GPUs of this node, one forked rank per GPU (slabcomm.hpp);
// the files written are the ones a single rank writes.
#include "mhd.hpp"
#include "slabcomm.hpp"
#include "utils.hpp"
#include <fstream>
#include <iostream>

// xdim * ydim from the head of a .state file (fileio.cpp:27-33): the size of the shared plane through which the ranks gather output variables
static size_t peekPlaneSize(const fs::path &state_file)
{
    std::ifstream in(state_file.string());
    SPRUCE_REQUIRE(in.good(), "cannot open state file " + state_file.string());
    std::string line;
    while (getCleanedLine(in, line)) {
        if (line.empty() || line[0] == '#') continue;
        SPRUCE_REQUIRE(line == "xdim,ydim", "state file must start with xdim,ydim");
        getCleanedLine(in, line);
        const std::vector<std::string> dims = splitString(line, ',');
        SPRUCE_REQUIRE(dims.size() == 2, "xdim,ydim needs two numbers");
        return (size_t)std::stoul(dims[0]) * (size_t)std::stoul(dims[1]);
    }
    spruce_die("state file " + state_file.string() + " is empty");
}
// Runs `solve` on every rank.  One rank: a plain call.  N ranks: fork before anything touches CUDA, share the exchange buffers, report rank 0's status.
template <class Solve> static int runRanks(int n_gpus, const fs::path &state_file, Solve solve)
{
    if (n_gpus <= 1) { solve(); return 0; }
    SlabComm &comm = SlabComm::instance();
    SPRUCE_REQUIRE(comm.create(n_gpus, peekPlaneSize(state_file)), "-g: between 2 and 16 ranks, and enough shared memory for one plane");
    spruce_die_hook() = [] { SlabComm::instance().markFailed(); };
    return comm.launch([&](int) { solve(); });
}

int main(int argc, char *argv[])
{
    const std::string run_mode = getCommandLineArg(argc, argv, "-m", "--mode");
    const std::string dur = getCommandLineArg(argc, argv, "-d", "--duration");
    const double time_duration = dur.empty() ? -1.0 : std::stod(dur);
    const std::string rt = getCommandLineArg(argc, argv, "-r", "--runtime");
    const double cluster_time = rt.empty() ? -1.0 : std::stod(rt);
    const fs::path out_path(getCommandLineArg(argc, argv, "-o", "--output"));
    SPRUCE_REQUIRE(!out_path.empty(), "output directory must be specified");
    std::string gpus = getCommandLineArg(argc, argv, "-g", "--gpus");
    if (gpus.empty() && std::getenv("SPRUCE_GPUS")) gpus = std::getenv("SPRUCE_GPUS");
    const int n_gpus = gpus.empty() ? 1 : std::stoi(gpus);

    if (run_mode == "continue") {
        SPRUCE_REQUIRE(!dur.empty(), "In Continue Mode, duration of simulation must be specified on command line");
        SPRUCE_REQUIRE(fs::is_directory(out_path), "Given output directory of previous run must be existing directory");
        std::cout << "Running in Continue Mode for " << time_duration << " s...\n";
        return runRanks(n_gpus, out_path / "end.state", [&] { mhdSolve(out_path, time_duration, cluster_time); });
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
    return runRanks(n_gpus, grid_path, [&] { mhdSolve(grid_path, config_path, out_path, time_duration, !seek_grids, cluster_time); });
}
