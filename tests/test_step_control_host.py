"""The step-control kernels of mhd_kernels.cuh executed on the HOST (their source cut out of the header, CUDA keywords replaced by stand-ins, one
thread): the fused form of plain runs (k_step_open, k_step_mid, k_step_close) must leave exactly the state the separate one-thread kernels
(k_step_begin, k_dtmin_reset, k_dt_validate, k_step_end) leave -- step sizes, time, iteration count, the running minimum and its skip window, the
`done` flag at max_time -- over random sequences of per-step minima, including steps whose minimum leaves the skip window (fallback evaluation)
and advance calls that stop at max_time.  Single rank (the all-gather branches are compiled, not entered)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
HDR = ROOT / "spruce_b200" / "csrc" / "mhd_kernels.cuh"
BUILD = ROOT / "tests" / "hostcheck" / "_build"
LIB = BUILD / "libstep_ctl_check.so"

PRELUDE = r'''
#include <cstring>
#include <cstdint>
#define __global__
#define __device__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)
struct Dim3 { unsigned x, y, z; };
static Dim3 threadIdx = {0, 0, 0};
static inline void __syncthreads() {}
static inline void __threadfence_system() {}
static inline void __nanosleep(unsigned) {}
static inline long long __double_as_longlong(double x) { long long b; std::memcpy(&b, &x, 8); return b; }
static inline double __longlong_as_double(long long b) { double x; std::memcpy(&x, &b, 8); return x; }
static inline int atomicExch(int *p, int v) { int o = *p; *p = v; return o; }
static inline unsigned long long ld_acquire_sys(const unsigned long long *p) { return *p; }
static inline void st_release_sys(unsigned long long *p, unsigned long long v) { *p = v; }
static inline unsigned long long global_timer_ns() { return 0; }
constexpr int MAX_RANKS = 16;
'''

DRIVER = r'''
static void stage(StepCtl *c, double pruned) {            // what the primary stage kernel does to the running minimum (atomicMin), unless the run is done
    if (c->done) return;
    unsigned long long b = (unsigned long long)__double_as_longlong(pruned);
    if (b < c->dtmin_bits) c->dtmin_bits = b;
}
static int n_full = 0;
static void full(StepCtl *c, double truth) { if (c->done || !c->need_full) return; n_full++; stage(c, truth); }       // k_dt_full
// truth[k]: the exact minimum after step k; junk[k]: what a pruned evaluation reports when the exact minimum lies above the window
extern "C" int fallback_count() { return n_full; }
extern "C" void run_steps(int fused, StepCtl *c, int n, const double *truth, const double *junk, double *hist, double max_time)
{
    c->max_time = max_time;
    DtGatherArgs G{}; G.ctl = c; G.world = 1;
    for (int k = 0; k < n; k++) {
        if (fused) { if (k == 0) k_step_open(c, hist, 0, 1); }
        else { k_step_begin(c, hist, k); k_dtmin_reset(c, 1); }
        const double thr = __longlong_as_double((long long)c->thr_bits);
        const bool inside = c->inv_thr == 0.0 || truth[k] <= thr;            // cells at or below the window's top are always evaluated
        stage(c, inside ? truth[k] : junk[k]);
        if (fused) { k_step_mid(G); full(c, truth[k]); k_step_close(G, hist, k + 1 < n ? k + 1 : -1, 1); }
        else { k_dt_validate(c); full(c, truth[k]); k_step_end(c); }
    }
}
extern "C" int ctl_size() { return (int)sizeof(StepCtl); }
'''


def cut(text, start, end):
    i = text.index(start)
    return text[i:text.index(end, i)]


@pytest.fixture(scope="module")
def lib():
    BUILD.mkdir(exist_ok=True)
    h = HDR.read_text()
    flags = cut(h, "struct PeerFlags {", "__device__ __forceinline__ unsigned long long ld_acquire_sys")
    wait = cut(h, "// spin until *flag >= seq", "struct PushArgs {")
    ctl = cut(h, "struct StepCtl {", "// the rare fallback of the skip test")
    gather = cut(h, "struct DtGatherArgs {", "// all-gather + reduce of the four module reduction words")
    src = BUILD / "step_ctl_check.cpp"
    text = PRELUDE + flags + wait + ctl + gather + DRIVER
    if not LIB.exists() or not src.exists() or src.read_text() != text:
        src.write_text(text)
        subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-o", str(LIB), str(src)], check=True)
    return C.CDLL(str(LIB))


class StepCtl(C.Structure):          # mhd_kernels.cuh: struct StepCtl
    _fields_ = [("step", C.c_double), ("time", C.c_double), ("max_time", C.c_double), ("epsilon", C.c_double), ("iter", C.c_longlong), ("done", C.c_int), ("need_full", C.c_int),
                ("dtmin_bits", C.c_ulonglong), ("inv_thr", C.c_double), ("thr_bits", C.c_ulonglong), ("prune_factor", C.c_double), ("full_seq", C.c_ulonglong)]


def bits(x):
    return int(np.float64(x).view(np.uint64))


def fresh(dt0, prune):
    c = StepCtl()
    c.epsilon, c.prune_factor, c.dtmin_bits, c.thr_bits = 0.2, prune, bits(dt0), bits(np.inf)
    return c


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("prune", [1.05, 1.0000001, 0.0])
def test_fused_step_control_equals_the_separate_kernels(lib, seed, prune):
    assert lib.ctl_size() == C.sizeof(StepCtl)
    rng = np.random.default_rng(seed)
    a, b = fresh(0.031, prune), fresh(0.031, prune)
    full0, early = lib.fallback_count(), 0
    for call in range(5):                                    # chained advance calls, as spruce_advance issues them
        n = int(rng.integers(1, 9))
        # a slowly drifting minimum with occasional jumps above the 5 % window (and below it)
        truth = 0.03 * np.exp(np.cumsum(rng.normal(0.0, 0.04, n))) * (1.0 + 0.3 * (rng.random(n) < 0.2))
        junk = truth * (1.0 + rng.random(n))                 # a pruned evaluation can only report something at or above the exact minimum
        max_time = -1.0 if call % 2 == 0 else a.time + 0.2 * float(np.sum(truth[: max(1, n // 2)]))      # every other call stops early
        ha, hb = np.zeros(n + 1), np.zeros(n + 1)
        vp = lambda x: x.ctypes.data_as(C.c_void_p)
        it0 = int(a.iter)
        lib.run_steps(0, C.byref(a), n, vp(truth), vp(junk), vp(ha), C.c_double(max_time))
        lib.run_steps(1, C.byref(b), n, vp(truth), vp(junk), vp(hb), C.c_double(max_time))
        early += int(a.iter) - it0 < n
        for f, _ in StepCtl._fields_:
            if f in ("need_full", "full_seq"):
                continue                                     # scratch between the kernels of a step; reset by the next begin
            assert getattr(a, f) == getattr(b, f), (call, f, getattr(a, f), getattr(b, f))
        assert np.array_equal(ha, hb), (call, ha, hb)
        assert a.iter > 0
        a.done = b.done = 0                                  # spruce_advance clears the flag after reading it
    assert early >= 1, "no advance call stopped at max_time: the test would not see the done logic"
    if prune > 1.0:
        assert lib.fallback_count() > full0, "the skip window was never missed: the test would not see the fallback logic"
