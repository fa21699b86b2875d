// text_io_bench.cpp -- host-side measurement for SURVEY 8f-1 (.state / mhd.out text I/O): one plane written and read back with
//   (a) the reference's way: std::ostringstream << std::setprecision(p) << double per element, std::getline + std::stod per element (grid.cpp:412-427, fileio.cpp:53-78)
//   (b) the host shell's way: Grid::format (rows in parallel, std::to_chars) and parseDelimitedRow (rows in parallel, std::from_chars)
// and checks that (b)'s text is byte-identical to (a)'s.   g++ -O2 -fopenmp -std=c++17 -I spruce_b200/host scripts/text_io_bench.cpp -o /tmp/text_io_bench && /tmp/text_io_bench 2048
#include "grid.hpp"
#include "utils.hpp"
#include <chrono>
#include <cmath>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <random>
#include <sstream>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv)
{
    const size_t n = argc > 1 ? (size_t)std::atoi(argv[1]) : 2048;
    const int prec = argc > 2 ? std::atoi(argv[2]) : 4;
    Grid g(n, n);
    std::mt19937_64 rng(7);
    for (size_t i = 0; i < n; i++) for (size_t j = 0; j < n; j++) g(i, j) = std::ldexp(std::uniform_real_distribution<double>(-1.0, 1.0)(rng), (int)(rng() % 80) - 40);
    double t0 = now();
    std::ostringstream ss;
    ss.precision(prec == -1 ? std::numeric_limits<double>::digits10 + 1 : prec);                 // grid.cpp:415-416
    for (size_t i = 0; i < n; i++) for (size_t j = 0; j < n; j++) { ss << g(i, j); ss << (j + 1 < n ? ',' : '\n'); }
    const std::string ref = ss.str();
    const double t_ref_w = now() - t0;
    t0 = now();
    const std::string ours = g.format(',', '\n', prec);
    const double t_our_w = now() - t0;
    if (ours != ref) { std::cout << "TEXT DIFFERS\n"; return 1; }
    // read back
    t0 = now();
    Grid a(n, n);
    { std::istringstream in(ref); std::string line, el; for (size_t i = 0; i < n; i++) { std::getline(in, line); std::istringstream ls(line); for (size_t j = 0; j < n; j++) { std::getline(ls, el, ','); a(i, j) = std::stod(el); } } }
    const double t_ref_r = now() - t0;
    t0 = now();
    Grid b(n, n);
    {
        std::vector<std::pair<const char *, const char *>> rows; rows.reserve(n);
        const char *p = ours.data(), *end = p + ours.size();
        while (p < end) { const char *q = (const char *)memchr(p, '\n', (size_t)(end - p)); if (!q) q = end; rows.emplace_back(p, q); p = q + 1; }
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)n; i++) parseDelimitedRow(rows[(size_t)i].first, rows[(size_t)i].second, b.ptr() + (size_t)i * n, n);
    }
    const double t_our_r = now() - t0;
    for (size_t k = 0; k < n * n; k++) if (a.ptr()[k] != b.ptr()[k]) { std::cout << "VALUES DIFFER\n"; return 1; }
    const double mb = (double)ref.size() / 1.0e6;
    std::cout << "{\"plane\": \"" << n << "x" << n << "\", \"precision\": " << prec << ", \"text_MB\": " << mb
              << ", \"write_ref_s\": " << t_ref_w << ", \"write_ours_s\": " << t_our_w << ", \"read_ref_s\": " << t_ref_r << ", \"read_ours_s\": " << t_our_r
              << ", \"write_speedup\": " << t_ref_w / t_our_w << ", \"read_speedup\": " << t_ref_r / t_our_r << ", \"identical_text\": true}\n";
    return 0;
}
