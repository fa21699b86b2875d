// slabcomm.hpp -- the drop-in binary on N GPUs of one node: `run -g N` forks one rank per GPU BEFORE anything touches CUDA; the ranks share one anonymous
// memory mapping for everything the host side of a slab decomposition has to exchange:
//   * the 64-byte CUDA IPC handles of the peer-store halo transport (spruce_mgpu_ipc_export / _connect) and the zero-plane masks (spruce_plane_activity),
//   * one full plane of doubles through which the slabs of an output variable are gathered: every rank writes its rows, every rank reads the whole plane
//     (text I/O stays on rank 0: .state in, mhd.out / end.state out -- fileio.cpp:14-80,144-255 are single-writer by nature),
//   * a sense-reversing barrier and a failure flag, so that a rank that dies (spruce_die -> abort) takes the others down instead of leaving them spinning.
// The time loop itself needs none of this: spruce_advance on a slab exchanges halos and the dt minimum between the GPUs over NVLink, without the host.
// Row ranges per rank are spruce_b200/multigpu.py's partition(): as even as possible, the first xdim % N ranks one row longer.
#pragma once
#include <atomic>
#include <csignal>
#include <new>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

struct SlabShared {
    static constexpr int kMaxRanks = 16;
    std::atomic<int> arrived, generation, failed;
    std::atomic<int> finished[kMaxRanks];
    int n_ranks;
    int masks[kMaxRanks];
    int flags[kMaxRanks];
    unsigned char ipc[kMaxRanks][64];
    size_t plane_doubles;
    // followed by plane_doubles doubles
    double *plane() { return reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(this) + ((sizeof(SlabShared) + 63) / 64) * 64); }
};

class SlabComm {
public:
    // one rank, no sharing: what a plain `run` uses
    SlabComm() {}
    static SlabComm &instance() { static SlabComm c; return c; }

    int rank() const { return m_rank; }
    int nRanks() const { return m_n; }
    bool active() const { return m_n > 1; }

    // (row0, rows) of a rank: multigpu.py partition()
    static void partition(int xdim, int n_ranks, int rank, int &row0, int &rows)
    {
        const int base = xdim / n_ranks, rem = xdim % n_ranks;
        row0 = rank * base + (rank < rem ? rank : rem);
        rows = base + (rank < rem ? 1 : 0);
    }

    // parent, before fork
    bool create(int n_ranks, size_t plane_doubles)
    {
        if (n_ranks < 2 || n_ranks > SlabShared::kMaxRanks) return false;
        m_bytes = ((sizeof(SlabShared) + 63) / 64) * 64 + plane_doubles * sizeof(double);
        void *p = mmap(nullptr, m_bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
        if (p == MAP_FAILED) return false;
        m_sh = new (p) SlabShared();
        m_sh->arrived = 0; m_sh->generation = 0; m_sh->failed = 0;
        for (auto &f : m_sh->finished) f = 0;
        m_sh->n_ranks = n_ranks; m_sh->plane_doubles = plane_doubles;
        m_n = n_ranks;
        return true;
    }
    void becomeRank(int rank) { m_rank = rank; }

    // every rank: wait for all; a failed peer ends this rank quietly (the parent reports)
    void barrier()
    {
        if (!active()) return;
        const int gen = m_sh->generation.load(std::memory_order_acquire);
        if (m_sh->arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == m_n) {
            m_sh->arrived.store(0, std::memory_order_relaxed);
            m_sh->generation.fetch_add(1, std::memory_order_release);
            return;
        }
        unsigned spins = 0;
        while (m_sh->generation.load(std::memory_order_acquire) == gen) {
            if (m_sh->failed.load(std::memory_order_relaxed)) std::_Exit(1);
            if (++spins > 2000) usleep(50);
        }
    }
    void markFailed() { if (active()) m_sh->failed.store(1); }
    void markFinished() { if (active()) m_sh->finished[m_rank].store(1); }
    bool finished(int rank) const { return m_sh->finished[rank].load() != 0; }

    // bitwise OR of one int per rank (zero-plane masks), minimum of one int per rank (did every rank succeed?)
    int allOr(int mine) { return allReduce(mine, true); }
    int allMin(int mine) { return allReduce(mine, false); }
    // rank 0's value everywhere (wall-clock decisions must not differ between ranks)
    int broadcast0(int mine)
    {
        if (!active()) return mine;
        if (m_rank == 0) m_sh->flags[0] = mine;
        barrier();
        const int v = m_sh->flags[0];
        barrier();
        return v;
    }
    // all ranks' 64-byte handles in rank order
    void allGather64(const void *mine, void *all)
    {
        std::memcpy(m_sh->ipc[m_rank], mine, 64);
        barrier();
        std::memcpy(all, m_sh->ipc, (size_t)64 * m_n);
        barrier();
    }
    // `plane` holds this rank's rows [row0, row0 + rows) of an xdim x ydim plane; afterwards it holds every rank's rows
    void allGatherRows(double *plane, size_t ydim, int row0, int rows, size_t total_doubles)
    {
        if (!active()) return;
        if (total_doubles > m_sh->plane_doubles) { std::fprintf(stderr, "slab gather: plane of %zu doubles exceeds the shared buffer (%zu)\n", total_doubles, m_sh->plane_doubles); markFailed(); std::abort(); }
        std::memcpy(m_sh->plane() + (size_t)row0 * ydim, plane + (size_t)row0 * ydim, sizeof(double) * (size_t)rows * ydim);
        barrier();
        std::memcpy(plane, m_sh->plane(), sizeof(double) * total_doubles);
        barrier();
    }

    // Parent: fork n ranks, each runs body(rank) and must not return normally from a completed run (the reference ends a run with abort(), evolution.cpp:54-56;
    // ranks > 0 leave through _Exit).  Returns the exit status for main(): rank 0's, or 1 when a rank died before finishing.
    template <class Body> int launch(Body body)
    {
        pid_t pids[SlabShared::kMaxRanks];
        std::fflush(stdout); std::fflush(stderr);
        for (int r = 0; r < m_n; r++) {
            pids[r] = fork();
            if (pids[r] < 0) { std::perror("fork"); m_sh->failed.store(1); return 1; }
            if (pids[r] == 0) {
                becomeRank(r);
                if (r > 0 && !std::freopen("/dev/null", "w", stdout)) std::_Exit(1);         // one stdout: rank 0's
                body(r);
                std::fflush(stdout);
                std::_Exit(0);
            }
        }
        int status0 = 0, left = m_n, status[SlabShared::kMaxRanks] = {0};
        bool died = false, unclean[SlabShared::kMaxRanks] = {false};
        while (left > 0) {
            int st = 0;
            const pid_t p = wait(&st);
            if (p < 0) break;
            int r = 0;
            while (r < m_n && pids[r] != p) r++;
            if (r == m_n) continue;
            left--;
            status[r] = st;
            if (r == 0) status0 = st;
            unclean[r] = !(finished(r) || (WIFEXITED(st) && WEXITSTATUS(st) == 0));
            if (unclean[r] && !died) { died = true; m_sh->failed.store(1); }                     // a rank died mid-run: release the others from their barriers
        }
        if (died) {                                                                              // name the rank that failed, not the ones that followed it out (they leave with status 1)
            int culprit = -1;
            for (int r = 0; r < m_n && culprit < 0; r++) if (unclean[r] && WIFSIGNALED(status[r])) culprit = r;
            for (int r = 0; r < m_n && culprit < 0; r++) if (unclean[r]) culprit = r;
            std::fprintf(stderr, "run: rank %d ended early (status 0x%x); the other ranks were stopped\n", culprit, status[culprit]);
        }
        if (died) return 1;
        if (WIFSIGNALED(status0)) { std::fflush(stderr); signal(WTERMSIG(status0), SIG_DFL); raise(WTERMSIG(status0)); }      // abort-on-success, like the reference
        return WIFEXITED(status0) ? WEXITSTATUS(status0) : 1;
    }

private:
    int m_rank = 0, m_n = 1;
    size_t m_bytes = 0;
    SlabShared *m_sh = nullptr;
    int allReduce(int mine, bool is_or)
    {
        if (!active()) return mine;
        m_sh->masks[m_rank] = mine;
        barrier();
        int v = m_sh->masks[0];
        for (int r = 1; r < m_n; r++) v = is_or ? (v | m_sh->masks[r]) : (m_sh->masks[r] < v ? m_sh->masks[r] : v);
        barrier();
        return v;
    }
};
