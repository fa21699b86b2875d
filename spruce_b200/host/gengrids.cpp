// gengrids.cpp -- the drop-in for the reference's UCNP problem generator (execs/gengrids.cpp): same command line, same files.
//
//   gengrids -p <output dir> -s <sweep>.settings -c <template>.config [-o 0|1] [-a <set number offset>]
//
// For every set of conditions of the sweep: <output dir>/set_<k>[/run_<r>]/{plasma.settings, ucnp.config, init.state} -- the inputs of `run`.
// Host-only: nothing here touches the device library (ucnp_generator.hpp).
#include "ucnp_generator.hpp"

int main(int argc, char *argv[])
{
    using namespace ucnpgen;
    const fs::path out_dir = getCommandLineArg(argc, argv, "-p", "--path");
    const fs::path settings_path = getCommandLineArg(argc, argv, "-s", "--settings");
    const fs::path config_path = getCommandLineArg(argc, argv, "-c", "--config");
    const std::string array_str = getCommandLineArg(argc, argv, "-a", "--array"), overwrite_str = getCommandLineArg(argc, argv, "-o", "--overwrite");
    SPRUCE_REQUIRE(settings_path.extension().string() == ".settings" && fs::exists(settings_path), "Error: settings file must exist and have extension .settings");
    SPRUCE_REQUIRE(config_path.extension().string() == ".config" && fs::exists(config_path), "Error: config file must exist and have extension .config");
    const int overwrite = overwrite_str.empty() ? 0 : std::stoi(overwrite_str), offset = array_str.empty() ? 0 : std::stoi(array_str);
    SPRUCE_REQUIRE(overwrite == 0 || overwrite == 1, "Error: the overwrite flag is logical and must be 0 or 1");
    fs::create_directories(out_dir);
    Settings sweep(settings_path);
    std::cout << "Valid task array: [0 " << num2str(sweep.array_size() - 1) << "]" << std::endl;
    for (int k = 0; k < sweep.array_size(); k++) {
        sweep.choose_array(k);
        const fs::path set_dir = out_dir / sweep.set_path(offset);
        if (overwrite == 0) SPRUCE_REQUIRE(!fs::exists(set_dir), "Error: folder already exists and overwrite_flag=0");
        sweep.write_array_params(set_dir, "plasma");
        ConfigHandler config(config_path);
        for (const std::string &name : sweep.names())
            if (config.is_config(name)) config.update_config(name, sweep.getvar(name));      // a settings row named like a config key overrides the template's line
        config.write_config_file(set_dir);
        StateHandler state(config.eqs_set_name());
        state.setup(sweep);
        state.write_state_file(set_dir);
    }
    return 0;
}
