/* direct_conv.c -- CPU oracle, TEST INFRASTRUCTURE ONLY.
 * Plain-C restatement of what CartesianGrids' CircularConvolution computes for
 * ImmersedLayers' inverse_laplacian! (src/grid_operators.jl:153-179; SURVEY.md
 * A.5): out[i,j] = sum_{k,l} G(|i-k|,|j-l|) w[k,l] over the field's own index
 * box, accumulated in long double, O(P^2).  Used to check the FFT-based oracle
 * (and through it the CUDA path) on small boxes.  Column-major, x fastest. */
#include <stdlib.h>

void ilm_oracle_direct_conv(const double* G, int ldg, const double* w, int mx, int my, double* out) {
    for (int j = 0; j < my; ++j)
        for (int i = 0; i < mx; ++i) {
            long double s = 0.0L;
            for (int l = 0; l < my; ++l) {
                const double* gcol = G + (size_t)abs(j - l) * ldg;
                const double* wcol = w + (size_t)l * mx;
                for (int k = 0; k < mx; ++k) s += (long double)gcol[abs(i - k)] * (long double)wcol[k];
            }
            out[(size_t)j * mx + i] = (double)s;
        }
}

/* ldiv!: out = (G*w - c0*sum(w)) / factor */
void ilm_oracle_inverse_laplacian(const double* G, int ldg, const double* w, int mx, int my, double c0,
                                  double factor, double* out) {
    long double sum = 0.0L;
    for (long q = 0; q < (long)mx * my; ++q) sum += w[q];
    ilm_oracle_direct_conv(G, ldg, w, mx, my, out);
    for (long q = 0; q < (long)mx * my; ++q) out[q] = (double)(((long double)out[q] - (long double)c0 * sum) / factor);
}
