// chunk_plan.hpp -- how the rows of a slab are split over the CTA rows of one k_mhd_stage_xy launch (host-side, plain C++: capi.cu uses it,
// tests/test_chunk_plan_host.py compiles it with g++ and checks that every row of every slab shape is covered exactly once).
#pragma once

namespace spruce {

constexpr int PLAN_MAX_ROWS = 56;      // = XY_CHUNK (mhd_stage_xy.cuh): rows per CTA, sizes the x tables in shared memory
constexpr int PLAN_EDGE_DELTA = 8;     // = XY_EDGE_DELTA: the edge CTA rows of an overlapped slab launch are this much shorter
constexpr int PLAN_HALO = 2;

struct ChunkPlan { int rows; int edge; int n_interior; };

// A CTA marches over `rows` rows (plus 4 warm-up rows and a prologue), so few, long chunks are cheap, but the grid should fill whole waves of
// resident CTAs (`cap` = 148 SMs x 5 for the 2-D instance, x 4 otherwise).  The plan takes the smallest number of waves PLAN_MAX_ROWS allows and
// spreads the rows evenly over the CTA rows that fit in them.  split: the first and the last `edge` rows form their own launch on the communication
// stream (they start first and are PLAN_EDGE_DELTA rows shorter, so that their rows travel while the interior chunks still run); `virt` = the rows a
// uniform split of all CTA rows would have to cover.
inline ChunkPlan plan_chunk_rows(int nx, int strips, int cap, bool split, int override_rows)
{
    ChunkPlan p{};
    const int delta = split ? PLAN_EDGE_DELTA : 0, virt = nx + 2 * delta;
    if (override_rows > 0) p.rows = override_rows < PLAN_MAX_ROWS ? override_rows : PLAN_MAX_ROWS;      // SPRUCE_CHUNK_ROWS: tuning sweeps
    else {
        const long long min_ctas = (long long)strips * ((virt + PLAN_MAX_ROWS - 1) / PLAN_MAX_ROWS);
        const long long waves = (min_ctas + cap - 1) / cap;
        long long cta_rows = waves * cap / strips;                            // CTA rows that fit in those waves
        if (cta_rows < 1) cta_rows = 1;
        p.rows = (int)((virt + cta_rows - 1) / cta_rows);
        if (p.rows > PLAN_MAX_ROWS) p.rows = PLAN_MAX_ROWS;
        if (p.rows < 2 * PLAN_EDGE_DELTA) p.rows = 2 * PLAN_EDGE_DELTA;
    }
    p.edge = split ? p.rows - delta : 0;
    if (split && p.edge < 2 * PLAN_HALO) p.edge = 2 * PLAN_HALO;             // (an override below 12 rows) the rows a neighbour needs come from the edge launch
    const int body = nx - 2 * p.edge;
    p.n_interior = body > 0 ? (body + p.rows - 1) / p.rows : 0;
    return p;
}

// Launch arguments of one part of a stage: 0 = every row in one launch, 1 = the edge launch (CTA row 0: the first `edge` rows, CTA row 1: the last),
// 2 = the interior launch.  The kernel maps CTA row b to [row_begin + b*chunk_rows, min(.. + chunk_rows, row_end)), or to [edge2_begin, edge2_end)
// for b == 1 when edge2_begin >= 0.
struct StageRows { int chunk_rows, row_begin, row_end, edge2_begin, edge2_end, grid_y; };
inline StageRows stage_rows(const ChunkPlan &cp, int nx, int part)
{
    StageRows s{cp.rows, cp.edge, nx - cp.edge, -1, -1, cp.n_interior};
    if (part == 1) s = StageRows{cp.edge, 0, cp.edge, nx - cp.edge, nx, 2};
    return s;
}

}  // namespace spruce
