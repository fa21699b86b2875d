// mhd.hpp -- driver entry points with the reference's signatures (source/mhd/mhd.hpp:13-17).
#pragma once
#include <filesystem>
namespace fs = std::filesystem;

// resume a finished run from its output directory (end.state + the config stored there)
void mhdSolve(const fs::path &prev_run_directory, double time_duration, double cluster_time);
// start from a .state file and a .config file
void mhdSolve(const fs::path &state_filename, const fs::path &config_filename, const fs::path &output_pathname, double time_duration, bool overwrite_init, double cluster_time);
