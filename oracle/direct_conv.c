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

/* Table form of a Schur complement built by probing (src/matrix_operators.jl:9-30 with the column loop
 * written out): A[k,c] = coef * sum_p sum_q wE[k,p] (K(|ip-iq|,|jp-jq|) - c0) wR[c,q], p over the window
 * of point k, q over the window of point c (regularize -> convolution -> interpolate of
 * a unit vector is a pure table look-up).  pi/pj: N*W2 window indices (negative = masked entry); entries
 * of K beyond the table (nk x nk, leading dimension ldk) count as zero (truncated integrating-factor
 * tables).  Accumulated in long double; columns [c_lo, c_hi), A is N x (c_hi-c_lo) column-major. */
void ilm_oracle_table_schur(const double* K, int ldk, int nk, int N, int W2, const long* pi, const long* pj,
                            const double* wE, const double* wR, int c_lo, int c_hi, double c0, double coef, double* A) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = c_lo; c < c_hi; ++c) {
        const long* qi = pi + (size_t)c * W2;
        const long* qj = pj + (size_t)c * W2;
        const double* r = wR + (size_t)c * W2;
        for (int k = 0; k < N; ++k) {
            const long* ki = pi + (size_t)k * W2;
            const long* kj = pj + (size_t)k * W2;
            const double* e = wE + (size_t)k * W2;
            long double s = 0.0L;
            for (int p = 0; p < W2; ++p) {
                if (ki[p] < 0 || e[p] == 0.0) continue;
                long double sp = 0.0L;
                for (int q = 0; q < W2; ++q) {
                    if (qi[q] < 0 || r[q] == 0.0) continue;
                    long di = labs(ki[p] - qi[q]), dj = labs(kj[p] - qj[q]);
                    double g = (di < nk && dj < nk) ? K[(size_t)dj * ldk + di] : 0.0;
                    sp += ((long double)g - (long double)c0) * (long double)r[q];
                }
                s += (long double)e[p] * sp;
            }
            A[(size_t)(c - c_lo) * N + k] = (double)((long double)coef * s);
        }
    }
}
