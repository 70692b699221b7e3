// W independent FP64 lanes per thread, written so that every source statement advances all W
// dependency chains together.  The hex kernels run 2-3 warps per scheduler (register- and
// shared-memory-bound occupancy) and an FP64 instruction can issue only every other cycle, so the
// pipe is filled from instruction-level parallelism: two independent two-point fluxes / face
// points / volume nodes per thread, interleaved statement by statement.
#pragma once
#include <cuda_runtime.h>

namespace esdg {

#define VFOR _Pragma("unroll") for (int q_ = 0; q_ < W; ++q_)

template <int W>
struct VD {
    double v[W];
    __device__ __forceinline__ double& operator[](int i) { return v[i]; }
    __device__ __forceinline__ const double& operator[](int i) const { return v[i]; }
};
template <int W>
struct VB {
    bool v[W];
};

template <int W> __device__ __forceinline__ VD<W> vset(double a) { VD<W> r; VFOR r.v[q_] = a; return r; }
template <int W> __device__ __forceinline__ VD<W> operator+(const VD<W>& a, const VD<W>& b) { VD<W> r; VFOR r.v[q_] = a.v[q_] + b.v[q_]; return r; }
template <int W> __device__ __forceinline__ VD<W> operator-(const VD<W>& a, const VD<W>& b) { VD<W> r; VFOR r.v[q_] = a.v[q_] - b.v[q_]; return r; }
template <int W> __device__ __forceinline__ VD<W> operator*(const VD<W>& a, const VD<W>& b) { VD<W> r; VFOR r.v[q_] = a.v[q_] * b.v[q_]; return r; }
template <int W> __device__ __forceinline__ VD<W> operator-(const VD<W>& a) { VD<W> r; VFOR r.v[q_] = -a.v[q_]; return r; }
template <int W> __device__ __forceinline__ VD<W> operator*(double a, const VD<W>& b) { VD<W> r; VFOR r.v[q_] = a * b.v[q_]; return r; }
template <int W> __device__ __forceinline__ VD<W> operator*(const VD<W>& a, double b) { VD<W> r; VFOR r.v[q_] = a.v[q_] * b; return r; }
template <int W> __device__ __forceinline__ VD<W> operator+(const VD<W>& a, double b) { VD<W> r; VFOR r.v[q_] = a.v[q_] + b; return r; }
template <int W> __device__ __forceinline__ VD<W> operator-(const VD<W>& a, double b) { VD<W> r; VFOR r.v[q_] = a.v[q_] - b; return r; }
template <int W> __device__ __forceinline__ VD<W> operator-(double a, const VD<W>& b) { VD<W> r; VFOR r.v[q_] = a - b.v[q_]; return r; }
template <int W> __device__ __forceinline__ VD<W> vfma(const VD<W>& a, const VD<W>& b, const VD<W>& c) { VD<W> r; VFOR r.v[q_] = fma(a.v[q_], b.v[q_], c.v[q_]); return r; }
template <int W> __device__ __forceinline__ VD<W> vfma(double a, const VD<W>& b, const VD<W>& c) { VD<W> r; VFOR r.v[q_] = fma(a, b.v[q_], c.v[q_]); return r; }
template <int W> __device__ __forceinline__ VD<W> vfma(const VD<W>& a, double b, const VD<W>& c) { VD<W> r; VFOR r.v[q_] = fma(a.v[q_], b, c.v[q_]); return r; }
template <int W> __device__ __forceinline__ VD<W> vfma(const VD<W>& a, const VD<W>& b, double c) { VD<W> r; VFOR r.v[q_] = fma(a.v[q_], b.v[q_], c); return r; }
template <int W> __device__ __forceinline__ VD<W> vfma(const VD<W>& a, double b, double c) { VD<W> r; VFOR r.v[q_] = fma(a.v[q_], b, c); return r; }
template <int W> __device__ __forceinline__ VD<W> vfma(double a, const VD<W>& b, double c) { VD<W> r; VFOR r.v[q_] = fma(a, b.v[q_], c); return r; }
template <int W> __device__ __forceinline__ VD<W> vabs(const VD<W>& a) { VD<W> r; VFOR r.v[q_] = fabs(a.v[q_]); return r; }
template <int W> __device__ __forceinline__ VD<W> vmax(const VD<W>& a, const VD<W>& b) { VD<W> r; VFOR r.v[q_] = fmax(a.v[q_], b.v[q_]); return r; }
template <int W> __device__ __forceinline__ VB<W> vlt(const VD<W>& a, const VD<W>& b) { VB<W> r; VFOR r.v[q_] = a.v[q_] < b.v[q_]; return r; }
template <int W> __device__ __forceinline__ VD<W> vsel(const VB<W>& c, const VD<W>& a, const VD<W>& b) { VD<W> r; VFOR r.v[q_] = c.v[q_] ? a.v[q_] : b.v[q_]; return r; }

// 1/x: MUFU seed (>= 20 bits: rcp.approx.ftz.f64 reads the high word only) + ONE third-order step
// y (1 + e + e^2), e = 1 - x y: relative error e^3 < 2^-60, three dependent FMAs instead of the four of two
// Newton steps (<= 1 ulp for normal x; tests/test_gpu_parity.py::test_kernel_log_exp_accuracy covers it through log/exp).
template <int W>
__device__ __forceinline__ VD<W> vrcp(const VD<W>& x) {
    VD<W> y, e;
    VFOR asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y.v[q_]) : "d"(x.v[q_]));
    VFOR e.v[q_] = fma(-x.v[q_], y.v[q_], 1.0);
    VFOR e.v[q_] = fma(e.v[q_], e.v[q_], e.v[q_]);
    VFOR y.v[q_] = fma(y.v[q_], e.v[q_], y.v[q_]);
    return y;
}

// sqrt(x), x >= 0, for the wave speed: MUFU rsqrt seed + one third-order step on y ~ 1/sqrt(x),
// y (1 + e/2 + 3 e^2/8), e = 1 - x y^2; result x*y (relative error ~2e-16, not correctly rounded).
// x is clamped to 1e-300 so that x == 0 gives 1e-150 instead of 0 * inf (the wave speed adds it to O(1) sound speeds).
template <int W>
__device__ __forceinline__ VD<W> vsqrt_fast(const VD<W>& x0) {
    VD<W> x, y, e, q;
    VFOR x.v[q_] = fmax(x0.v[q_], 1e-300);
    VFOR asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y.v[q_]) : "d"(x.v[q_]));
    VFOR e.v[q_] = fma(-x.v[q_] * y.v[q_], y.v[q_], 1.0);
    VFOR q.v[q_] = e.v[q_] * fma(.375, e.v[q_], .5);
    VFOR y.v[q_] = fma(y.v[q_], q.v[q_], y.v[q_]);
    VFOR y.v[q_] = x.v[q_] * y.v[q_];
    return y;
}

// sqrt(x), x >= 0: MUFU rsqrt seed, two Newton steps on y ~ 1/sqrt(x), one correction of s = x*y
// (<= 1 ulp; the IEEE sqrt() is ~2.5x the instructions).  x == 0 -> 0.
template <int W>
__device__ __forceinline__ VD<W> vsqrt(const VD<W>& x) {
    VD<W> y, e, s, h;
    VFOR asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y.v[q_]) : "d"(x.v[q_]));
    VFOR e.v[q_] = fma(-x.v[q_] * y.v[q_], y.v[q_], 1.0);
    VFOR y.v[q_] = fma(.5 * y.v[q_], e.v[q_], y.v[q_]);
    VFOR e.v[q_] = fma(-x.v[q_] * y.v[q_], y.v[q_], 1.0);
    VFOR y.v[q_] = fma(.5 * y.v[q_], e.v[q_], y.v[q_]);
    VFOR s.v[q_] = x.v[q_] * y.v[q_];
    VFOR h.v[q_] = fma(-s.v[q_], s.v[q_], x.v[q_]);
    VFOR s.v[q_] = fma(.5 * y.v[q_], h.v[q_], s.v[q_]);
    VFOR s.v[q_] = x.v[q_] > 0.0 ? s.v[q_] : 0.0;
    return s;
}

}  // namespace esdg
