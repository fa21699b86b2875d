// chunk_plan_check.cpp -- TEST INFRASTRUCTURE.  Compiles spruce_b200/csrc/chunk_plan.hpp for the host and walks the launches capi.cu forms from it with the
// row mapping of k_mhd_stage_xy (mhd_stage_xy.cuh: r0 / r1 from blockIdx.y): every row of the slab must be covered exactly once, no CTA may get more rows
// than the shared-memory x tables hold, and on an overlapped slab the rows a neighbour needs must come from the edge launch.
#include "../../spruce_b200/csrc/chunk_plan.hpp"
#include <vector>
using namespace spruce;

// the kernel's mapping of a CTA row to its rows (mhd_stage_xy.cuh)
static void cta_rows(const StageRows &A, int by, int *r0, int *r1)
{
    const bool edge2 = A.edge2_begin >= 0 && by == 1;
    *r0 = edge2 ? A.edge2_begin : A.row_begin + by * A.chunk_rows;
    const int end = *r0 + A.chunk_rows < A.row_end ? *r0 + A.chunk_rows : A.row_end;
    *r1 = edge2 ? A.edge2_end : end;
}

// returns 0 when the plan is sound, otherwise a code naming what failed
extern "C" int chunk_plan_check(int nx, int strips, int cap, int split, int override_rows, int *rows_out, int *ctas_out)
{
    const ChunkPlan cp = plan_chunk_rows(nx, strips, cap, split != 0, override_rows);
    *rows_out = cp.rows;
    std::vector<int> hits(nx, 0), from_edge(nx, 0);
    int ctas = 0;
    const int parts[2] = {split ? 1 : 0, split ? 2 : -1};
    for (int part : parts) {
        if (part < 0) continue;
        const StageRows A = stage_rows(cp, nx, part);
        if (A.chunk_rows > PLAN_MAX_ROWS) return 1;
        for (int by = 0; by < A.grid_y; by++) {
            int r0, r1;
            cta_rows(A, by, &r0, &r1);
            if (r1 - r0 > PLAN_MAX_ROWS) return 2;
            if (r1 <= r0) return 3;                                  // an empty CTA row: wasted launch slot (and a zero-row prologue)
            for (int r = r0; r < r1; r++) { if (r < 0 || r >= nx) return 4; hits[r]++; if (part == 1) from_edge[r] = 1; }
            ctas += strips;
        }
    }
    for (int r = 0; r < nx; r++) if (hits[r] != 1) return 5;
    if (split) for (int h = 0; h < PLAN_HALO; h++) if (!from_edge[h] || !from_edge[nx - 1 - h]) return 6;
    *ctas_out = ctas;
    return 0;
}
