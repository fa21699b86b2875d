// host_scan_check.cpp -- TEST INFRASTRUCTURE.  Compiles spruce_b200/csrc/host_scan.hpp (the zero-plane test spruce_grid_upload runs while a copy is in flight) for tests/test_host_scan.py.
#include "../../spruce_b200/csrc/host_scan.hpp"
extern "C" int plane_nonzero(const double *p, size_t n, unsigned threads) { return spruce::host_plane_nonzero(p, n, threads) ? 1 : 0; }
