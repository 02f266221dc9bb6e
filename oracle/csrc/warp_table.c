/* ORACLE (test infrastructure only; never linked into or loaded by pysp_b200/).
 *
 * Plain-C restatement of the DNG WarpRectilinear coordinate table of the reference
 * (dng_warp_corr/dng_warp_rectilinear_coords.pyx:18-40 compute_table, :44-65 offset_table, :67-95 the two entry points),
 * with the arithmetic types of the C that Cython generates from it (checked against oracle/_ref/*.c):
 *   - `x ** 2`, `r ** 4`, `r ** 6` on a C float are powf(x, 2.0f) ... (libm, not correctly rounded in general);
 *   - `sqrt` is libc's double sqrt on the float sum, rounded back to float;
 *   - the literals in `2 * dx * dy` and `2 * dx ** 2` are C doubles, so the two tangential terms are evaluated in double
 *     and rounded once when they are stored into the float dxt / dyt;
 *   - the entry points compute the optical centre and the normalisation radius in float; `np.sqrt` of the Python float
 *     (a double) is rounded when it is assigned to `cdef float m`.
 * Built with -O2 -ffp-contract=off (no fused multiply-add), like the reference extension in oracle/build_ref.py.
 */
#include <math.h>
#include <stddef.h>

void oracle_warp_scalars(unsigned width, unsigned height, float cnx, float cny, float* cx, float* cy, float* m) {
    *cx = (width - 1) * cnx;                                      /* pyx:73  unsigned * float -> float */
    *cy = (height - 1) * cny;
    float a = fabsf((width - 1) - *cx), b = fabsf(-*cx);          /* pyx:75  max(abs(-cx), abs(width - 1 - cx)) */
    float mdx = a > b ? a : b;
    a = fabsf((height - 1) - *cy); b = fabsf(-*cy);
    float mdy = a > b ? a : b;
    *m = (float)sqrt((double)(powf(mdx, 2.0f) + powf(mdy, 2.0f)));    /* pyx:77  np.sqrt(float32 sum as a Python float) */
}

static void one(float sx, float sy, const float k[6], float m, float cx, float cy, float scale, float* ox, float* oy) {
    const float kr0 = k[0], kr1 = k[1], kr2 = k[2], kr3 = k[3], kt0 = k[4], kt1 = k[5];
    float dx = (sx - cx) / m, dy = (sy - cy) / m;
    float r = (float)sqrt((double)(powf(dx, 2.0f) + powf(dy, 2.0f)));
    float f = ((kr0 + (kr1 * powf(r, 2.0f))) + (kr2 * powf(r, 4.0f))) + (kr3 * powf(r, 6.0f));
    float dxr = f * dx, dyr = f * dy;
    float dxt = (float)((kt0 * ((2.0 * dx) * dy)) + (kt1 * (powf(r, 2.0f) + (2.0 * powf(dx, 2.0f)))));
    float dyt = (float)((kt1 * ((2.0 * dx) * dy)) + (kt0 * (powf(r, 2.0f) + (2.0 * powf(dy, 2.0f)))));
    float xp = cx + (m * (dxr + dxt));
    float yp = cy + (m * (dyr + dyt));
    *ox = sx + ((xp - sx) * scale);
    *oy = sy + ((yp - sy) * scale);
}

/* table[H][W][2]; seed NULL: compute_remapping_table, else compute_offset_remapping_table (seed[H][W][2]) */
void oracle_warp_table(float* table, const float* seed, unsigned width, unsigned height, const float k[6], float cnx, float cny,
                       float scale) {
    float cx, cy, m;
    oracle_warp_scalars(width, height, cnx, cny, &cx, &cy, &m);
    for (unsigned y = 0; y < height; ++y)
        for (unsigned x = 0; x < width; ++x) {
            size_t o = ((size_t)y * width + x) * 2;
            float sx = seed ? seed[o] : (float)(int)x, sy = seed ? seed[o + 1] : (float)(int)y;
            one(sx, sy, k, m, cx, cy, scale, &table[o], &table[o + 1]);
        }
}
