// Prelude for generated model headers (see triflow_b200/codegen.py).
//
// The generated F / J bodies are written against the TF_* operation macros so
// that the same text compiles (a) with nvcc for sm_100a, where every operation
// is an IEEE round-to-nearest intrinsic that ptxas may not contract into an FMA,
// and (b) with g++ -ffp-contract=off for the CPU bit-parity test of the
// translator (tests/test_codegen_cpu.py).  Operation order is the printed order
// of the reference's lambdify source (reference triflow/core/compilers.py:207-219).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define TF_HD __host__ __device__
#define TF_D __device__
#define TF_INLINE __forceinline__
#define TF_RESTRICT __restrict__
#else
#define TF_HD
#define TF_D
#define TF_INLINE inline __attribute__((always_inline))
#define TF_RESTRICT __restrict__
#endif

#ifdef __CUDA_ARCH__
#define TF_ADD(a, b) __dadd_rn((a), (b))
#define TF_SUB(a, b) __dsub_rn((a), (b))
#define TF_MUL(a, b) __dmul_rn((a), (b))
#define TF_DIV(a, b) __ddiv_rn((a), (b))
#define TF_SQRT(a) __dsqrt_rn((a))
#define TF_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define TF_ADD(a, b) ((a) + (b))
#define TF_SUB(a, b) ((a) - (b))
#define TF_MUL(a, b) ((a) * (b))
#define TF_DIV(a, b) ((a) / (b))
#define TF_SQRT(a) sqrt((a))
#define TF_FMA(a, b, c) fma((a), (b), (c))
#endif
#define TF_POW(a, b) pow((a), (b))
// x**n for a small literal integer n, evaluated in double-double so that the result is
// (almost always) the correctly rounded power, like glibc's pow which NumPy calls for
// array**n (n not in {-1, 0, 0.5, 1, 2}).  CUDA's pow is only accurate to 2 ulp, and the
// expanded stencils amplify one ulp of a large term by many orders of magnitude.
TF_HD TF_INLINE double tf_powi(double x, int n) {
  const int m = n < 0 ? -n : n;
  double hi = x, lo = 0.0;                       // running power as hi + lo
  for (int k = 1; k < m; ++k) {
    const double p = hi * x;
    const double e = TF_FMA(hi, x, -p) + lo * x; // exact product error + carried low part
    const double s = p + e;
    lo = e - (s - p);
    hi = s;
  }
  if (n >= 0) return hi + lo;
  const double q0 = 1.0 / hi;
  const double r = TF_FMA(-hi, q0, 1.0) - lo * q0;   // 1 - (hi + lo) * q0
  return TF_FMA(q0, r, q0);
}
#define TF_NAN (__builtin_nan(""))
#define TF_INF (__builtin_inf())

// numpy.maximum / numpy.minimum semantics (NaN propagates from the first operand,
// ties return the first operand)
TF_HD TF_INLINE double TF_MAX(double a, double b) { return (a >= b || a != a) ? a : b; }
TF_HD TF_INLINE double TF_MIN(double a, double b) { return (a <= b || a != a) ? a : b; }
TF_HD TF_INLINE double TF_SIGN(double a) { return a > 0 ? 1.0 : (a < 0 ? -1.0 : a); }

// Division by a uniform constant cst[j].  Exact mode: a true IEEE division.
// Fast mode (template argument TF_FD = true): q0 = a*rc; r = fma(-q0, c, a); q = fma(r, rc, q0)
// with rc = RN(1/c) from the host table (entry TF_NCONST + j).  The residual r is
// exact, so q is the correctly rounded quotient except when a/c lies within
// ~2^-105 (relative) of a rounding boundary, where it may be off by one ulp.
TF_HD TF_INLINE double tf_fdiv(double a, double c, double rc) {
  const double q0 = TF_MUL(a, rc);
  const double r = TF_FMA(-q0, c, a);
  return TF_FMA(r, rc, q0);
}
// TF_FD is the bool template parameter of the generated tf_model_F / tf_model_J
#define TF_DIVC(a, j) (TF_FD ? tf_fdiv((a), cst[(j)], cst[TF_NCONST + (j)]) : TF_DIV((a), cst[(j)]))

// One node's inputs: stencil window of every field (dependent variables first,
// then helper functions), per-node parameter values, and x.
struct TfNodeIn {
  double w[TF_NFIELD][TF_WW];
  double np[TF_NNODEPAR > 0 ? TF_NNODEPAR : 1];
  double x;
};
