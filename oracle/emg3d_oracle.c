/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the emg3d multigrid hot path.
 * See kernels.inc for what is restated and from where.  Built by oracle/Makefile
 * into oracle/libemg3d_oracle.so and loaded through ctypes by oracle/__init__.py.
 * Parity status: pinned against outputs of the reference itself (imported in the
 * build container) and the reference's own regression data, see
 * tests/golden/make_golden.py and tests/test_oracle_golden.py.                  */
#include <complex.h>
#include <stdlib.h>
#include <string.h>

#define T double complex
#define NAME(x) x##_c
#include "kernels.inc"
#undef T
#undef NAME

#define T double
#define NAME(x) x##_r
#include "kernels.inc"
#undef T
#undef NAME

/* 1-D restriction weights (core.py:2004-2076; Muld06 Eq. 9).  Dual-cell widths
 * d with half cells at both ends; left/right weights from the distance between
 * fine and coarse cell centres.                                                */
void orc_restrict_weights(const double *nodes, const double *centers, const double *h,
                          const double *cnodes, const double *ccenters,
                          const double *ch, int ncn, double *wl, double *w0, double *wr)
{
    /* ncn = number of coarse nodes; fine cells = 2 (ncn - 1) */
    int nf = 2 * (ncn - 1), nc = ncn - 1;
    for (int i = 0; i < ncn; ++i) {
        double dl = i == 0 ? h[0] / 2 : (h[2 * i - 2] + h[2 * i - 1]) / 2.0;
        double dr = i == ncn - 1 ? h[nf - 1] / 2 : (h[2 * i] + h[2 * i + 1]) / 2.0;
        double num_l = i == 0 ? (nodes[0] - h[0] / 2) - (cnodes[0] - ch[0] / 2)
                              : centers[2 * i - 1] - ccenters[i - 1];
        double num_r = i == ncn - 1
                           ? (cnodes[ncn - 1] + ch[nc - 1] / 2) - (nodes[nf] + h[nf - 1] / 2)
                           : ccenters[i] - centers[2 * i];
        wl[i] = (1.0 / dl) * num_l;
        w0[i] = 1.0;
        wr[i] = (1.0 / dr) * num_r;
    }
}
