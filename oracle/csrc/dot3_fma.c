/* ORACLE (test infrastructure only; never linked into or loaded by pysp_b200/).
 *
 * f32( M . c ) with the float64 accumulation order of the reference's `np.dot(rgb, color_mat.T)`
 * (colorize/transform.py:52-53).  NumPy hands that product to OpenBLAS dgemm, whose x86-64 kernels accumulate with
 * fused multiply-adds over k = 0, 1, 2:   acc = m0*c0 ; acc = fma(m1, c1, acc) ; acc = fma(m2, c2, acc).
 * Pinned by tests/golden/dot_fma_pins.npz: inputs on which the fused and the unfused float64 sums round to
 * different float32 values, with the outputs of the unmodified reference (tests/golden/make_fma_pins.py).
 * fma() is the correctly rounded C99 function; built with -ffp-contract=off so nothing else is contracted.
 */
#include <math.h>
#include <stddef.h>

void oracle_dot3_fma(const float* rgb, size_t n, const double* m, float* out) {
    for (size_t i = 0; i < n; ++i) {
        const double c0 = rgb[3 * i], c1 = rgb[3 * i + 1], c2 = rgb[3 * i + 2];
        for (int k = 0; k < 3; ++k) {
            double acc = m[3 * k] * c0;
            acc = fma(m[3 * k + 1], c1, acc);
            acc = fma(m[3 * k + 2], c2, acc);
            out[3 * i + k] = (float)acc;
        }
    }
}
