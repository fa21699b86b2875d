"""Host-shell text I/O (spruce_b200/host/grid.hpp, utils.hpp): Grid::format must emit exactly the characters printf("%.{p}g")
emits (what the reference's ostringstream << setprecision(p) << double writes, fileio.cpp / grid.cpp:412-427), and
parseDelimitedRow must read rows back to the same doubles strtod gives.  Compiled here with g++ (no CUDA needed)."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
SRC = r'''
#include "grid.hpp"
#include "utils.hpp"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
int main()
{
    std::mt19937_64 rng(12345);
    const size_t R = 37, Cc = 211;
    Grid g(R, Cc);
    for (size_t i = 0; i < R; i++) for (size_t j = 0; j < Cc; j++) {
        const double m = std::uniform_real_distribution<double>(-1.0, 1.0)(rng);
        const int e = (int)(rng() % 600) - 300;
        double v = std::ldexp(m, e);
        if ((i * Cc + j) % 97 == 0) v = 0.0;
        if ((i * Cc + j) % 101 == 0) v = (double)((long long)(rng() % 2000000) - 1000000);
        if ((i * Cc + j) % 103 == 0) v = 1.0e-320;                 // denormal
        g(i, j) = v;
    }
    g(0, 0) = -0.0; g(0, 1) = 1e22; g(0, 2) = 123456789012345678.0; g(0, 3) = 0.1; g(0, 4) = 100000.0; g(0, 5) = 1e-5; g(0, 6) = 9.9999999e-5;
    for (int prec : {4, 6, 16, 17, -1}) {
        const int p = prec == -1 ? 16 : prec;
        std::string ref;
        char buf[64];
        for (size_t i = 0; i < R; i++) for (size_t j = 0; j < Cc; j++) {
            std::snprintf(buf, sizeof(buf), "%.*g", p, g(i, j));
            ref += buf;
            ref.push_back(j + 1 < Cc ? ',' : (i + 1 < R ? ';' : '\n'));
        }
        const std::string got = g.format(',', ';', prec, '\n');
        if (got != ref) { std::printf("FORMAT MISMATCH at precision %d\n", prec); return 1; }
        // parse the rows back
        size_t pos = 0;
        for (size_t i = 0; i < R; i++) {
            size_t e = got.find(i + 1 < R ? ';' : '\n', pos);
            std::vector<double> row(Cc);
            const size_t n = parseDelimitedRow(got.data() + pos, got.data() + e, row.data(), Cc);
            if (n != Cc) { std::printf("PARSE COUNT %zu at row %zu\n", n, i); return 1; }
            const char *q = got.data() + pos;
            for (size_t j = 0; j < Cc; j++) {
                char *end = nullptr;
                const double v = std::strtod(q, &end);
                if (std::memcmp(&v, &row[j], sizeof(double)) != 0) { std::printf("PARSE MISMATCH %zu %zu\n", i, j); return 1; }
                q = end + 1;
            }
            pos = e + 1;
        }
    }
    double tmp[4];
    if (parseDelimitedRow("1.5,abc", (const char *)"1.5,abc" + 7, tmp, 4) != (size_t)-1) { std::printf("non-numeric field accepted\n"); return 1; }
    const char *five = "1,2,3,4,5";
    if (parseDelimitedRow(five, five + 9, tmp, 4) <= 4) { std::printf("over-long row accepted\n"); return 1; }
    std::printf("OK\n");
    return 0;
}
'''


def test_grid_format_and_row_parser_match_printf_and_strtod(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-fopenmp", "-I", str(ROOT / "spruce_b200" / "host"), str(src), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], stdout=subprocess.PIPE, timeout=120)
    assert r.returncode == 0 and b"OK" in r.stdout, r.stdout.decode()
