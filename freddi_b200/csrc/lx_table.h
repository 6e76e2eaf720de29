// The band integral of the Planck function as a function of ONE variable — format of the table and its evaluation
// (host + device).  The table itself is built on the host by lx_table.cpp.
//
// planck_nu1_nu2 (cpp/src/spectrum.cpp:26-37) integrates B_nu(T) over [nu1, nu2] with odeint's controlled Cash-Karp stepper
// (eps_abs = 0, eps_rel = 1e-4, first step 0.01 (nu2 - nu1)).  With x = (h/kT) nu the integrand is (kT/h)^3 x^3 / (e^x - 1), and
// the controller's error ratio |x_err| / (eps_rel (|x| + dt |dxdt|)) does not see a common factor of the state: the whole
// walk — every step size, every accept / reject / grow decision, and so the result with its ~1e-6 quadrature error — is a
// function of x1 = h nu1 / kT alone once nu2 / nu1 is fixed (which it is per model: emax / emin):
//
//     integral = nu1^4 J(x1),      J(x1) = the walk's result for nu1 = 1, nu2 = r, h/kT = x1.
//
// J is piecewise analytic in x1: it jumps (by ~1e-6 relative) where a controller decision flips, i.e. where one step's error
// ratio crosses 1 (accept / reject) or 0.5 (grow or not), has a kink where it crosses 5^-5 or 91.125 (the clamps of the step
// factor), and where the last step's clipping changes.  lx_table.cpp locates these breakpoints down to slivers of 2e-9 relative
// (an error ratio carries ~1e-10 of rounding noise: closer to a threshold the decisions alternate from one double to the next, in
// the reference as in any restatement) and fits q(x1) = J(x1) e^{x1} on every piece in between (split further where needed) by a
// degree-11 polynomial to <= 2e-12 relative — six orders below the walk's own quadrature error, two below the 1e-10 parity
// tolerance.  A ring then costs one table look-up, one exponential and 11 FMAs instead of ~5.5 trial steps of ~100 FP64
// instructions.  Rings in a sliver (about one in 1e8) or in a sub-piece that could not be fitted (flagged trials = -1), rings
// outside the table's range (x1 < 2^-5 or >= 2^8), models whose band ratio is not the table's, and ensembles created with
// exact_lx_walk walk as before.  The look-up reproduces the walk of THIS library (planck_walk, disc_engine.h), whose decisions
// agree with odeint's except within rounding of a threshold — the set no restatement can decide.
#ifndef FB200_LX_TABLE_H
#define FB200_LX_TABLE_H

#include <stdint.h>
#include <string.h>

#include "fb200_math.h"

namespace fb200 {

// Blob layout, in doubles (the blob is 128-byte aligned):
//   [0] r = nu2 / nu1   [1] number of sub-pieces   [2] eps_rel   [3] largest fit error   [4] breakpoints   [5] walked sub-pieces
//   [6] their ln-width   [7] reserved
//   [FB_LX_O_CELL ...)  one int2 per cell: (first, last) sub-piece overlapping the cell.  A cell is 1/64 of an octave of x1:
//                       its index is the top 18 bits of the double (sign, exponent, 6 mantissa bits) minus that of 2^-5.
//   [FB_LX_O_EDGE ...)  edges[k] = the first double of sub-piece k (edges[n_sub] = 2^8)
//   [FB_LX_O_REC ...)   16 doubles per sub-piece: mid, 1 / half-width, a_0 .. a_11 (q = sum a_j u^j, u = (x1 - mid) / half-width),
//                       the walk's try_step count on this piece (-1: not fitted, the kernels walk), unused
#define FB_LX_DEG 11
#define FB_LX_CELLS (13 * 64)
#define FB_LX_MAX_SUB 2047
#define FB_LX_O_CELL 8
#define FB_LX_O_EDGE 848
#define FB_LX_O_REC (FB_LX_O_EDGE + FB_LX_MAX_SUB + 1)
#define FB_LX_REC 16
#define FB_LX_X_LO 0.03125
#define FB_LX_X_HI 256.0
#define FB_LX_CELL0 (0x3fa00000 >> 14) /* top bits of 2^-5 */

static_assert((FB_LX_O_EDGE * 8) % 128 == 0 && (FB_LX_O_REC * 8) % 128 == 0, "table sections are 128-byte aligned");

inline size_t lx_table_doubles(int n_sub) { return size_t(FB_LX_O_REC) + size_t(FB_LX_REC) * size_t(n_sub); }

FB_HD int lx_cell_of(double x1) {
#ifdef __CUDA_ARCH__
    return (__double2hiint(x1) >> 14) - FB_LX_CELL0;
#else
    uint64_t b;
    memcpy(&b, &x1, 8);
    return int(uint32_t(b >> 32) >> 14) - FB_LX_CELL0;
#endif
}

// q(x1) = J(x1) e^{x1} for FB_LX_X_LO <= x1 < FB_LX_X_HI; *trials = the try_step calls the walk makes there.
FB_HD double lx_table_q(const double* __restrict__ tb, double x1, int* trials) {
    const int c = lx_cell_of(x1);
#ifdef __CUDA_ARCH__
    const int2 ck = __ldg(reinterpret_cast<const int2*>(tb + FB_LX_O_CELL) + c);
    int k = ck.x;
    const int k1 = ck.y;
    while (k < k1 && x1 >= __ldg(tb + FB_LX_O_EDGE + k + 1)) ++k;
    const double2* rec = reinterpret_cast<const double2*>(tb + FB_LX_O_REC + FB_LX_REC * k);
    const double2 h0 = __ldg(rec), h1 = __ldg(rec + 1), h2 = __ldg(rec + 2), h3 = __ldg(rec + 3), h4 = __ldg(rec + 4), h5 = __ldg(rec + 5),
                  h6 = __ldg(rec + 6), h7 = __ldg(rec + 7);
    const double u = (x1 - h0.x) * h0.y;
    double q = fma(h6.y, u, h6.x);
    q = fma(q, u, h5.y); q = fma(q, u, h5.x);
    q = fma(q, u, h4.y); q = fma(q, u, h4.x);
    q = fma(q, u, h3.y); q = fma(q, u, h3.x);
    q = fma(q, u, h2.y); q = fma(q, u, h2.x);
    q = fma(q, u, h1.y); q = fma(q, u, h1.x);
    *trials = int(h7.x);
    return q;
#else
    const int* ck = reinterpret_cast<const int*>(tb + FB_LX_O_CELL) + 2 * c;
    int k = ck[0];
    const int k1 = ck[1];
    while (k < k1 && x1 >= tb[FB_LX_O_EDGE + k + 1]) ++k;
    const double* rec = tb + FB_LX_O_REC + FB_LX_REC * k;
    const double u = (x1 - rec[0]) * rec[1];
    double q = rec[2 + FB_LX_DEG];
    for (int j = FB_LX_DEG - 1; j >= 0; --j) q = fma(q, u, rec[2 + j]);
    *trials = int(rec[14]);
    return q;
#endif
}

// Host side (lx_table.cpp): the table for a band ratio, built on first use and kept for the life of the process (nullptr when
// the build fails its own checks; the kernels then walk).  `doubles` = length of the blob.
const double* lx_table_get(double r, size_t* doubles);

}  // namespace fb200
#endif
