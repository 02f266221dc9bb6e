// TEST INFRASTRUCTURE ONLY -- host emulation of the tile functions in pysp_b200/csrc/*.cuh.
//
// The CUDA tile functions are written as barrier-separated phases whose work items are independent, so
// the very same source compiles for the host (PYSP_HOST_EMU): a phase runs its items serially, a barrier
// is a no-op.  This lets tests/test_tile_logic.py diff the exact kernel logic (tiling, halos, the six
// border rules, band seams, flips) against the oracle in a container without a GPU.  It is NOT a CPU
// fallback: it is never built into libpysp_b200.so, never imported by the pysp_b200 package, and its
// entry point takes HOST pointers.  Build: tests/host_emu/build.sh (g++ -ffp-contract=off).
#define PYSP_HOST_EMU 1
#include <stdlib.h>
#include <vector>

struct uint2 { unsigned int x, y; };

#include "../../pysp_b200/csrc/ahd_select.cuh"
#include "../../pysp_b200/csrc/median_stage.cuh"
#include "../../pysp_b200/csrc/develop_plan.h"

using namespace pysp;

static char g_err[512];

extern "C" const char* emu_last_error(void) { return g_err; }

extern "C" int emu_develop(const pysp_develop_args* a, int tw1, int th1) {
    // tile sizes are compile-time in the kernels; the emulation instantiates the product's (60x28) and a
    // small one (12x8) that puts many tile seams and partial tiles into small test frames
    DevelopPlan plan;
    int rc = plan_develop(a, tw1, th1, tw1, th1, &plan, g_err, sizeof(g_err));
    if (rc) return rc;
    std::vector<float> smem(64 * 1024);
    auto run_select = [&](auto tile_fn_edge, auto tile_fn_int, int TW, int TH) {
        const SelectParams& p = plan.select;
        for (int t = 0; t < plan.select_tiles; ++t) {
            int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
            int x0 = tx * TW, y0 = p.y_begin + ty * TH;
            bool edge = x0 < 6 || y0 < 6 || x0 + TW + 6 > p.g.W || y0 + TH + 6 > p.g.H || y0 + TH > p.y_end;
            for (size_t i = 0; i < smem.size(); ++i) smem[i] = __builtin_nanf("");   // poison
            if (edge) tile_fn_edge(p, smem.data(), tx, ty); else tile_fn_int(p, smem.data(), tx, ty);
        }
    };
    auto run_median = [&](auto tile_fn_edge, auto tile_fn_int, int TW, int TH) {
        for (int s = 0; s < plan.n_stages; ++s) {
            const MedianParams& p = plan.median[s];
            for (int t = 0; t < plan.median_tiles[s]; ++t) {
                int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
                int x0 = tx * TW, y0 = p.y_begin + ty * TH;
                bool edge = x0 < 4 || y0 < 4 || x0 + TW + 4 > p.g.W || y0 + TH + 4 > p.g.H || y0 + TH > p.y_end;
                for (size_t i = 0; i < smem.size(); ++i) smem[i] = __builtin_nanf("");
                if (edge) tile_fn_edge(p, smem.data(), tx, ty); else tile_fn_int(p, smem.data(), tx, ty);
            }
        }
    };
    if (tw1 == 60 && th1 == 28) {
        run_select([](const SelectParams& p, float* s, int x, int y) { select_tile<60, 28, true>(p, s, x, y); },
                   [](const SelectParams& p, float* s, int x, int y) { select_tile<60, 28, false>(p, s, x, y); }, 60, 28);
        run_median([](const MedianParams& p, float* s, int x, int y) { median_tile<60, 28, true>(p, s, x, y); },
                   [](const MedianParams& p, float* s, int x, int y) { median_tile<60, 28, false>(p, s, x, y); }, 60, 28);
    } else if (tw1 == 12 && th1 == 8) {
        run_select([](const SelectParams& p, float* s, int x, int y) { select_tile<12, 8, true>(p, s, x, y); },
                   [](const SelectParams& p, float* s, int x, int y) { select_tile<12, 8, false>(p, s, x, y); }, 12, 8);
        run_median([](const MedianParams& p, float* s, int x, int y) { median_tile<12, 8, true>(p, s, x, y); },
                   [](const MedianParams& p, float* s, int x, int y) { median_tile<12, 8, false>(p, s, x, y); }, 12, 8);
    } else {
        snprintf(g_err, sizeof(g_err), "emu_develop: tile %dx%d not instantiated", tw1, th1);
        return PYSP_ERR_INVALID;
    }
    return PYSP_OK;
}
