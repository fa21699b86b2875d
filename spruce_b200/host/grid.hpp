// grid.hpp -- host-side staging type with the reference Grid's shape and text format (source/mhd/grid.hpp:15-165).
// On the B200 path a Grid never carries arithmetic: the field store is the device arena behind include/spruce_b200.h.
// What remains on the host is what file I/O and module set-up need: element access in the reference layout
// (row-major i*cols + j, source/mhd/grid.cpp:516-526) and the character-delimited text form (grid.cpp:412-427).
#pragma once
#include <cassert>
#include <charconv>
#include <cstdio>
#include <limits>
#include <string>
#include <vector>

class Grid {
public:
    Grid() : m_rows(1), m_cols(1), m_data(1, 0.0) {}
    Grid(size_t rows, size_t cols, double val = 0.0) : m_rows(rows), m_cols(cols), m_data(rows * cols, val) {}
    Grid(size_t rows, size_t cols, std::vector<double> data) : m_rows(rows), m_cols(cols), m_data(std::move(data)) { assert(m_data.size() == rows * cols); }
    static Grid Zero(size_t rows, size_t cols) { return Grid(rows, cols, 0.0); }
    static Grid Ones(size_t rows, size_t cols) { return Grid(rows, cols, 1.0); }

    double &operator()(size_t i, size_t j) { assert(i < m_rows && j < m_cols); return m_data[i * m_cols + j]; }
    double operator()(size_t i, size_t j) const { assert(i < m_rows && j < m_cols); return m_data[i * m_cols + j]; }
    int rows() const { return (int)m_rows; }
    int cols() const { return (int)m_cols; }
    int size() const { return (int)m_data.size(); }
    const std::vector<double> &data() const { return m_data; }
    std::vector<double> &data() { return m_data; }
    const double *ptr() const { return m_data.data(); }
    double *ptr() { return m_data.data(); }

    // Same text as the reference's ostringstream << double with `precision` significant digits (default float format ==
    // printf %.{p}g); precision -1 means digits10 + 1 = 16.  Elements separated by element_delim, rows by row_delim, the
    // last row delimiter replaced by end_delim.
    std::string format(char element_delim = ',', char row_delim = '\n', int precision = 4, char end_delim = '\n') const
    {
        const int p = precision == -1 ? std::numeric_limits<double>::digits10 + 1 : (precision <= 0 ? 6 : precision);
        // rows are formatted independently (OpenMP when the host shell is built with it) and concatenated in order;
        // std::to_chars(general, p) is specified to produce what printf("%.{p}g") produces, at a fraction of the cost
        std::vector<std::string> rows(m_rows);
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)m_rows; i++) {
            std::string &out = rows[(size_t)i];
            out.reserve(m_cols * (size_t)(p + 8));
            char buf[64];
            for (size_t j = 0; j < m_cols; j++) {
                const auto r = std::to_chars(buf, buf + sizeof(buf), m_data[(size_t)i * m_cols + j], std::chars_format::general, p);
                out.append(buf, (size_t)(r.ptr - buf));
                out.push_back(j + 1 < m_cols ? element_delim : ((size_t)i + 1 < m_rows ? row_delim : end_delim));
            }
        }
        size_t total = 0;
        for (const std::string &r : rows) total += r.size();
        std::string all;
        all.reserve(total);
        for (const std::string &r : rows) all += r;
        return all;
    }

private:
    size_t m_rows, m_cols;
    std::vector<double> m_data;
};
