// utils.hpp -- string helpers with the reference's semantics (source/mhd/utils.cpp:8-53).
#pragma once
#include <algorithm>
#include <charconv>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

// the reference aborts through assert(); keep the message + SIGABRT behaviour without depending on NDEBUG
// (on N GPUs a dying rank first tells its peers, so that they leave their barriers: slabcomm.hpp)
inline void (*&spruce_die_hook())() { static void (*hook)() = nullptr; return hook; }
[[noreturn]] inline void spruce_die(const std::string &msg)
{
    std::cerr << msg << std::endl;
    if (spruce_die_hook()) spruce_die_hook()();
    std::abort();
}
#define SPRUCE_REQUIRE(cond, msg) do { if (!(cond)) spruce_die(std::string("Assertion `") + #cond + "' failed: " + (msg)); } while (0)

inline void clearWhitespace(std::string &s)
{
    s.erase(std::remove_if(s.begin(), s.end(), [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\b' || c == '\r' || c == '\f' || c == '\a'; }), s.end());
}
inline std::istream &getCleanedLine(std::istream &is, std::string &s, char delim = '\n')
{
    std::getline(is, s, delim);
    clearWhitespace(s);
    return is;
}
inline std::vector<std::string> splitString(const std::string &s, char delim)
{
    std::vector<std::string> out;
    std::istringstream ss(s);
    std::string el;
    while (std::getline(ss, el, delim)) if (!el.empty()) out.push_back(el);
    return out;
}
// flag/value pairs; a flag may appear once (utils.cpp:39-53)
inline std::string getCommandLineArg(int argc, char *argv[], const std::string &short_flag, const std::string &long_flag)
{
    std::string result;
    for (int i = 1; i < argc; i += 2) {
        SPRUCE_REQUIRE(argv[i][0] == '-' && i + 1 < argc, "Arguments must be given as flag followed by non-flag");
        if (short_flag == argv[i] || long_flag == argv[i]) {
            SPRUCE_REQUIRE(result.empty(), "Command line flags cannot be used more than once");
            result = argv[i + 1];
        }
    }
    return result;
}
// key = value # comment  -> (lhs, rhs) of an already whitespace-stripped line
inline void splitAssignment(const std::string &line, std::string &lhs, std::string &rhs)
{
    std::istringstream ss(line);
    lhs.clear(); rhs.clear();
    std::getline(ss, lhs, '=');
    std::getline(ss, rhs, '#');
}

// One delimiter-separated row of numbers (already stripped of whitespace) -> out[0..max_n).  Returns the number of values, or
// (size_t)-1 when a field is not a number.  std::from_chars accepts exactly what the reference's writers emit (no leading '+').
inline size_t parseDelimitedRow(const char *p, const char *end, double *out, size_t max_n, char delim = ',')
{
    size_t n = 0;
    while (p < end) {
        double v = 0.0;
        const auto r = std::from_chars(p, end, v);
        if (r.ec == std::errc::result_out_of_range) {              // overflow / underflow: strtod's value (inf / denormal / 0), as the reference's stod-free reader gets
            char *e2 = nullptr;
            v = std::strtod(std::string(p, r.ptr).c_str(), &e2);
        } else if (r.ec != std::errc() || r.ptr == p) {
            return (size_t)-1;
        }
        if (n >= max_n) return max_n + 1;                          // row too long
        out[n++] = v;
        p = (r.ptr < end && *r.ptr == delim) ? r.ptr + 1 : r.ptr;
        if (r.ptr < end && *r.ptr != delim) return (size_t)-1;
    }
    return n;
}
