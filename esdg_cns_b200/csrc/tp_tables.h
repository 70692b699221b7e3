// 1-D generators of the tensor-product operator tables (Gauss-Legendre nodes, Lagrange differentiation / end-point
// interpolation) and the comparison used to prove that a caller's operators are Kronecker products of them.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "esdg_ctx.h"

namespace esdg_host {

// Gauss-Legendre nodes/weights on [-1,1], ascending (Newton on P_n).
inline void gauss_legendre(int n, double* x, double* w) {
    for (int i = 0; i < n; ++i) {
        double z = std::cos(M_PI * (i + 0.75) / (n + 0.5));
        double pp = 1.0;
        for (int it = 0; it < 100; ++it) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 0; j < n; ++j) {
                const double p3 = p2;
                p2 = p1;
                p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0);
            }
            pp = n * (z * p1 - p2) / (z * z - 1.0);
            const double dz = p1 / pp;
            z -= dz;
            if (std::fabs(dz) < 1e-16) break;
        }
        x[n - 1 - i] = z;
        w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
    }
}

inline void lagrange_tables(int n, const double* x, double* D, double* lm, double* lp) {
    std::vector<double> c(n, 1.0);
    for (int j = 0; j < n; ++j)
        for (int m = 0; m < n; ++m)
            if (m != j) c[j] *= (x[j] - x[m]);
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int j = 0; j < n; ++j)
            if (i != j) {
                D[i * n + j] = c[i] / (c[j] * (x[i] - x[j]));
                s += D[i * n + j];
            }
        D[i * n + i] = -s;
    }
    for (int j = 0; j < n; ++j) {
        double a = 1.0, b = 1.0;
        for (int m = 0; m < n; ++m)
            if (m != j) {
                a *= (-1.0 - x[m]) / (x[j] - x[m]);
                b *= (1.0 - x[m]) / (x[j] - x[m]);
            }
        lm[j] = a;
        lp[j] = b;
    }
}

// max |A - B| / max |B| for a provided column-major matrix against a canonical generator
template <class F>
double rel_mismatch(const HostMat& A, F canon) {
    double err = 0.0, sc = 0.0;
    for (int64_t j = 0; j < A.cols; ++j)
        for (int64_t i = 0; i < A.rows; ++i) {
            const double c = canon((int)i, (int)j);
            err = std::max(err, std::fabs(A.at(i, j) - c));
            sc = std::max(sc, std::fabs(c));
        }
    return err / std::max(sc, 1e-300);
}

}  // namespace esdg_host
