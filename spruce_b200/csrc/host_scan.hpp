// host_scan.hpp -- is a host plane identically +-0?  (plain C++; capi.cu uses it, tests/test_host_scan.py compiles it with g++.)
// The library tracks which uploaded planes are exactly zero (mom_z, bi_z, be_*, grav_*): their transports are exactly zero in the reference too, so the stage kernel's
// 2-D instance may skip them.  The test is an OR over the bit patterns with the sign bit shifted out.  A 4096^2 plane is 134 MB: one core needs ~12 ms for it, seven
// tracked planes ~85 ms -- more than the PCIe copies of the whole job -- so large planes are scanned by several threads, each stopping at its first non-zero block, and
// spruce_grid_upload runs the scan WHILE the asynchronous copy of the same plane is in flight.
#pragma once
#include <atomic>
#include <cstddef>
#include <cstring>
#include <thread>
#include <vector>

namespace spruce {

inline bool host_range_nonzero(const double *p, size_t n, const std::atomic<bool> *stop)
{
    for (size_t k0 = 0; k0 < n; k0 += 4096) {                 // blocks: vectorisable OR, early exit per block
        if (stop && stop->load(std::memory_order_relaxed)) return false;
        const size_t k1 = k0 + 4096 < n ? k0 + 4096 : n;
        unsigned long long acc = 0ULL;
        for (size_t k = k0; k < k1; k++) { unsigned long long b; std::memcpy(&b, p + k, sizeof(b)); acc |= b; }
        if ((acc << 1) != 0ULL) return true;
    }
    return false;
}

// true when any value of the plane is not +-0 (NaNs and denormals count as non-zero)
inline bool host_plane_nonzero(const double *host, size_t count, unsigned max_threads = 8)
{
    const size_t kParallelFrom = (size_t)1 << 20;             // 8 MB: below this one thread is faster than starting more
    unsigned nt = std::thread::hardware_concurrency();
    if (nt > max_threads) nt = max_threads;
    if (count < kParallelFrom || nt < 2) return host_range_nonzero(host, count, nullptr);
    std::atomic<bool> found(false);
    std::vector<std::thread> th;
    const size_t chunk = (count + nt - 1) / nt;
    auto work = [&](size_t a, size_t b) { if (host_range_nonzero(host + a, b - a, &found)) found.store(true, std::memory_order_relaxed); };
    try {
        for (unsigned t = 1; t < nt; t++) {
            const size_t a = (size_t)t * chunk, b = a + chunk < count ? a + chunk : count;
            if (a < b) th.emplace_back(work, a, b);
        }
    } catch (...) {                                           // no more threads to be had: the caller's thread does the rest below
        for (auto &t : th) t.join();
        return found.load() || host_range_nonzero(host, count, nullptr);
    }
    work(0, chunk < count ? chunk : count);
    for (auto &t : th) t.join();
    return found.load();
}

}  // namespace spruce
