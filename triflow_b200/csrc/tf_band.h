// Chunk-level mathematics of the banded LU solver (host + device).
//
// Replaces scipy.sparse.linalg.factorized / spsolve (SuperLU) at reference
// triflow/core/schemes.py:149,157,557.  A system of n unknowns with scalar
// half-bandwidth BETA is cut into chunks of C >= BETA consecutive unknowns, one
// chunk per thread.  Everything that couples chunks is expressed as an
// *associative* operator so that the chain over chunks can be evaluated with a
// parallel scan (warp shuffles -> shared memory -> decoupled look-back):
//
//  * LU factorisation (no pivoting).  Eliminating all rows before a chunk
//    changes only the chunk's leading BETA x BETA block, A' = A - X.  The map
//    X_in -> X_out of one chunk is a linear-fractional (Redheffer) map
//        X_out = P + Q X (I - R X)^-1 S ,
//    closed under composition (StarMap::combine).  run1 computes (P,Q,R,S) of a
//    chunk from its rows; after the scan, run2 redoes the elimination with the
//    true X_in and emits the L multipliers and U rows.
//  * forward / backward substitution are affine recurrences on the last BETA
//    solution values: s_out = Phi s_in + c  (AffMap::combine).
//
// All loops have compile-time bounds so that the small matrices live in
// registers when compiled for the device.
#pragma once

#ifndef TF_HD
#ifdef __CUDACC__
#define TF_HD __host__ __device__
#define TF_INLINE __forceinline__
#else
#define TF_HD
#define TF_INLINE inline __attribute__((always_inline))
#endif
#endif

#ifdef __CUDACC__
#define TF_UNROLL _Pragma("unroll")
#else
#define TF_UNROLL
#endif

// largest multiplier |l| = |a / pivot| accepted before the factorisation is reported as failed
#ifndef TF_GROWTH_LIMIT
#define TF_GROWTH_LIMIT 1e10
#endif

// reciprocal of a pivot: correctly rounded, without the IEEE-division slow path
#ifdef __CUDA_ARCH__
#define TF_RCP(x) __drcp_rn(x)
#else
#define TF_RCP(x) (1.0 / (x))
#endif

namespace tfb {

template <int A, int B> struct Min { static constexpr int v = A < B ? A : B; };
template <int A, int B> struct Max { static constexpr int v = A > B ? A : B; };

// ----------------------------------------------------------- small dense algebra
// C = A * B (n x n, row-major flat)
template <int N>
TF_HD TF_INLINE void mm(const double* A, const double* B, double* C) {
  TF_UNROLL for (int i = 0; i < N; ++i)
    TF_UNROLL for (int j = 0; j < N; ++j) {
      double s = 0.0;
      TF_UNROLL for (int k = 0; k < N; ++k) s += A[i * N + k] * B[k * N + j];
      C[i * N + j] = s;
    }
}
// C += A * B
template <int N>
TF_HD TF_INLINE void mma(const double* A, const double* B, double* C) {
  TF_UNROLL for (int i = 0; i < N; ++i)
    TF_UNROLL for (int j = 0; j < N; ++j) {
      double s = C[i * N + j];
      TF_UNROLL for (int k = 0; k < N; ++k) s += A[i * N + k] * B[k * N + j];
      C[i * N + j] = s;
    }
}

// Solve M X = RHS for NR right-hand sides, in place (RHS <- X); M is destroyed.
// Gaussian elimination with partial pivoting done by predicated row exchanges so
// that every index is a compile-time constant.
template <int N, int NR>
TF_HD TF_INLINE void solve_inplace(double* M, double* RHS) {
  TF_UNROLL for (int k = 0; k < N; ++k) {
    TF_UNROLL for (int r = k + 1; r < N; ++r) {
      const bool sw = fabs(M[r * N + k]) > fabs(M[k * N + k]);
      TF_UNROLL for (int c = k; c < N; ++c) {
        const double a = M[k * N + c], b = M[r * N + c];
        M[k * N + c] = sw ? b : a;
        M[r * N + c] = sw ? a : b;
      }
      TF_UNROLL for (int c = 0; c < NR; ++c) {
        const double a = RHS[k * NR + c], b = RHS[r * NR + c];
        RHS[k * NR + c] = sw ? b : a;
        RHS[r * NR + c] = sw ? a : b;
      }
    }
    const double inv = TF_RCP(M[k * N + k]);
    TF_UNROLL for (int c = k + 1; c < N; ++c) M[k * N + c] *= inv;
    TF_UNROLL for (int c = 0; c < NR; ++c) RHS[k * NR + c] *= inv;
    TF_UNROLL for (int r = 0; r < N; ++r) {
      if (r == k) continue;
      const double f = M[r * N + k];
      TF_UNROLL for (int c = k + 1; c < N; ++c) M[r * N + c] -= f * M[k * N + c];
      TF_UNROLL for (int c = 0; c < NR; ++c) RHS[r * NR + c] -= f * RHS[k * NR + c];
    }
  }
}

// ------------------------------------------------------------------ StarMap
// X_out = P + Q X (I - R X)^-1 S ;  d = [P | Q | R | S], each B x B row-major.
template <int B>
struct StarMap {
  static constexpr int K = 4 * B * B;
  double d[K];
  TF_HD TF_INLINE double* P() { return d; }
  TF_HD TF_INLINE double* Q() { return d + B * B; }
  TF_HD TF_INLINE double* R() { return d + 2 * B * B; }
  TF_HD TF_INLINE double* S() { return d + 3 * B * B; }
  TF_HD TF_INLINE const double* P() const { return d; }
  TF_HD TF_INLINE const double* Q() const { return d + B * B; }
  TF_HD TF_INLINE const double* R() const { return d + 2 * B * B; }
  TF_HD TF_INLINE const double* S() const { return d + 3 * B * B; }

  TF_HD static TF_INLINE StarMap identity() {
    StarMap m;
    TF_UNROLL for (int i = 0; i < K; ++i) m.d[i] = 0.0;
    TF_UNROLL for (int i = 0; i < B; ++i) { m.Q()[i * B + i] = 1.0; m.S()[i * B + i] = 1.0; }
    return m;
  }
  // `a` is applied first (earlier chunk), `b` second.
  TF_HD static TF_INLINE StarMap combine(const StarMap& a, const StarMap& b) {
    // K = (I - Pa Rb)^-1 ; KPQ = K [Pa | Qa]
    double Mx[B * B], KPQ[B * 2 * B];
    TF_UNROLL for (int i = 0; i < B; ++i)
      TF_UNROLL for (int j = 0; j < B; ++j) {
        double s = (i == j) ? 1.0 : 0.0;
        TF_UNROLL for (int k = 0; k < B; ++k) s -= a.P()[i * B + k] * b.R()[k * B + j];
        Mx[i * B + j] = s;
        KPQ[i * 2 * B + j] = a.P()[i * B + j];
        KPQ[i * 2 * B + B + j] = a.Q()[i * B + j];
      }
    solve_inplace<B, 2 * B>(Mx, KPQ);
    double KP[B * B], KQ[B * B];
    TF_UNROLL for (int i = 0; i < B; ++i)
      TF_UNROLL for (int j = 0; j < B; ++j) {
        KP[i * B + j] = KPQ[i * 2 * B + j];
        KQ[i * B + j] = KPQ[i * 2 * B + B + j];
      }
    StarMap o;
    double T1[B * B], T2[B * B], T3[B * B];
    mm<B>(KP, b.S(), T1);                       // T1 = K Pa Sb
    TF_UNROLL for (int i = 0; i < B * B; ++i) o.P()[i] = b.P()[i];
    mma<B>(b.Q(), T1, o.P());                   // P = Pb + Qb K Pa Sb
    mm<B>(b.Q(), KQ, o.Q());                    // Q = Qb K Qa
    mm<B>(b.R(), KQ, T2);                       // T2 = Rb K Qa
    TF_UNROLL for (int i = 0; i < B * B; ++i) o.R()[i] = a.R()[i];
    mma<B>(a.S(), T2, o.R());                   // R = Ra + Sa Rb K Qa
    TF_UNROLL for (int i = 0; i < B * B; ++i) T3[i] = b.S()[i];
    mma<B>(b.R(), T1, T3);                      // T3 = Sb + Rb K Pa Sb
    mm<B>(a.S(), T3, o.S());                    // S = Sa (I - Rb Pa)^-1 Sb
    return o;
  }
};

// ------------------------------------------------------------------- AffMap
// s_out = Phi s_in + c ; d = [Phi (B x B) | c (B)]
template <int B>
struct AffMap {
  static constexpr int K = B * B + B;
  double d[K];
  TF_HD TF_INLINE double* Phi() { return d; }
  TF_HD TF_INLINE double* c() { return d + B * B; }
  TF_HD TF_INLINE const double* Phi() const { return d; }
  TF_HD TF_INLINE const double* c() const { return d + B * B; }
  TF_HD static TF_INLINE AffMap identity() {
    AffMap m;
    TF_UNROLL for (int i = 0; i < K; ++i) m.d[i] = 0.0;
    TF_UNROLL for (int i = 0; i < B; ++i) m.Phi()[i * B + i] = 1.0;
    return m;
  }
  TF_HD static TF_INLINE AffMap combine(const AffMap& a, const AffMap& b) {
    AffMap o;
    mm<B>(b.Phi(), a.Phi(), o.Phi());
    TF_UNROLL for (int i = 0; i < B; ++i) {
      double s = b.c()[i];
      TF_UNROLL for (int k = 0; k < B; ++k) s += b.Phi()[i * B + k] * a.c()[k];
      o.c()[i] = s;
    }
    return o;
  }
};

// ---------------------------------------------------------------- LU of a chunk
// Row storage: A[r][BETA + d] = A(r, r + d), d in [-BETA, BETA], for the chunk's
// own rows r = 0..C-1 followed by the next chunk's first BETA rows (r = C..C+BETA-1,
// only their entries in columns < C are used).
template <int BETA, int C>
struct ChunkLU {
  static constexpr int W = 2 * BETA + 1;
  static constexpr int RT = C + BETA;

  // (P,Q,R,S) of the chunk.
  TF_HD static TF_INLINE void run1(const double (&A)[RT][W], StarMap<BETA>& out, int& bad) {
    double T[C][W];
    TF_UNROLL for (int r = 0; r < C; ++r)
      TF_UNROLL for (int j = 0; j < W; ++j) T[r][j] = A[r][j];
    // RHS columns: [E_hat (BETA) | Cn (BETA)], Y = T^-1 RHS
    double Y[C][2 * BETA];
    TF_UNROLL for (int r = 0; r < C; ++r)
      TF_UNROLL for (int b = 0; b < BETA; ++b) {
        Y[r][b] = (r == b) ? 1.0 : 0.0;
        // Cn(r, b) = A(r, C + b): band offset d = C + b - r
        Y[r][BETA + b] = (C + b - r <= BETA) ? A[r][BETA + C + b - r] : 0.0;
      }
    // banded LU without pivoting + forward substitution
    TF_UNROLL for (int k = 0; k < C; ++k) {
      const double piv = T[k][BETA];
      if (!(piv != 0.0) || !(fabs(piv) < 1e300)) bad |= 1;
      const double inv = TF_RCP(piv);
      TF_UNROLL for (int r = k + 1; r < C && r <= k + BETA; ++r) {
        const double l = T[r][BETA + k - r] * inv;
        TF_UNROLL for (int c = k + 1; c < C && c <= k + BETA; ++c)
          T[r][BETA + c - r] -= l * T[k][BETA + c - k];
        TF_UNROLL for (int b = 0; b < 2 * BETA; ++b) Y[r][b] -= l * Y[k][b];
      }
    }
    // back substitution
    TF_UNROLL for (int k = C - 1; k >= 0; --k) {
      const double inv = TF_RCP(T[k][BETA]);
      TF_UNROLL for (int b = 0; b < 2 * BETA; ++b) {
        double s = Y[k][b];
        TF_UNROLL for (int c = k + 1; c < C && c <= k + BETA; ++c)
          s -= T[k][BETA + c - k] * Y[c][b];
        Y[k][b] = s * inv;
      }
    }
    // R0 = E^T Y_E, S0 = E^T Y_C ; Q0 = Rn Y_E, P0 = Rn Y_C
    TF_UNROLL for (int a = 0; a < BETA; ++a)
      TF_UNROLL for (int b = 0; b < BETA; ++b) {
        out.R()[a * BETA + b] = Y[a][b];
        out.S()[a * BETA + b] = Y[a][BETA + b];
        double q = 0.0, p = 0.0;
        // Rn(a, c) = A(C + a, c): band offset d = c - C - a >= -BETA
        TF_UNROLL for (int c = C - BETA; c < C; ++c) {
          if (c >= 0 && c - C - a >= -BETA) {
            const double rn = A[C + a][BETA + c - C - a];
            q += rn * Y[c][b];
            p += rn * Y[c][BETA + b];
          }
        }
        out.Q()[a * BETA + b] = q;
        out.P()[a * BETA + b] = p;
      }
  }

  // Elimination with the true incoming update X (row-major BETA x BETA).
  // Outputs: Uf[r] = {1/pivot, u(r,r+1..r+BETA)}; Lown[r][q-1] = l(r, r-q) for own
  // pivots (q <= r); Lnext[a][q-1] = l(C+a, C+a-q) for own pivots (q > a).
  TF_HD static TF_INLINE void run2(const double (&A)[RT][W], const double* X,
                                   double (&Uf)[C][BETA + 1], double (&Lown)[C][BETA],
                                   double (&Lnext)[BETA][BETA], int& bad) {
    double T[RT][W];
    TF_UNROLL for (int r = 0; r < RT; ++r)
      TF_UNROLL for (int j = 0; j < W; ++j) T[r][j] = A[r][j];
    TF_UNROLL for (int a = 0; a < BETA; ++a)
      TF_UNROLL for (int b = 0; b < BETA; ++b) T[a][BETA + b - a] -= X[a * BETA + b];
    TF_UNROLL for (int r = 0; r < C; ++r)
      TF_UNROLL for (int q = 0; q < BETA; ++q) Lown[r][q] = 0.0;
    TF_UNROLL for (int a = 0; a < BETA; ++a)
      TF_UNROLL for (int q = 0; q < BETA; ++q) Lnext[a][q] = 0.0;
    TF_UNROLL for (int k = 0; k < C; ++k) {
      const double piv = T[k][BETA];
      if (!(piv != 0.0) || !(fabs(piv) < 1e300)) bad |= 1;
      // no pivoting here (SuperLU row-pivots): a pivot far smaller than the entries it
      // eliminates means multipliers > 1e10 and a useless factor -- say so (status bit 3)
      {
        double mx = 0.0;
        TF_UNROLL for (int r = k + 1; r < RT && r <= k + BETA; ++r) mx = fmax(mx, fabs(T[r][BETA + k - r]));
        if (fabs(piv) * TF_GROWTH_LIMIT < mx) bad |= 8;
      }
      const double inv = TF_RCP(piv);
      Uf[k][0] = inv;
      TF_UNROLL for (int c = 1; c <= BETA; ++c) Uf[k][c] = T[k][BETA + c];
      TF_UNROLL for (int r = k + 1; r < RT && r <= k + BETA; ++r) {
        const double l = T[r][BETA + k - r] * inv;
        if (r < C) Lown[r][r - k - 1] = l; else Lnext[r - C][r - k - 1] = l;
        TF_UNROLL for (int c = k + 1; c <= k + BETA; ++c)
          if (c - r >= -BETA && c - r <= BETA) T[r][BETA + c - r] -= l * T[k][BETA + c - k];
      }
    }
  }

  // Same elimination, additionally returning the update X_out left on the next
  // block.  The BETA extra rows must hold only their entries in columns < C (zeros
  // elsewhere), so that after the elimination they hold -X_out.
  TF_HD static TF_INLINE void run2x(const double (&A)[RT][W], const double* X,
                                    double (&Uf)[C][BETA + 1], double (&Lown)[C][BETA],
                                    double (&Lnext)[BETA][BETA], double* Xout, int& bad) {
    double T[RT][W];
    TF_UNROLL for (int r = 0; r < RT; ++r)
      TF_UNROLL for (int j = 0; j < W; ++j) T[r][j] = A[r][j];
    TF_UNROLL for (int a = 0; a < BETA; ++a)
      TF_UNROLL for (int b = 0; b < BETA; ++b) T[a][BETA + b - a] -= X[a * BETA + b];
    TF_UNROLL for (int r = 0; r < C; ++r)
      TF_UNROLL for (int q = 0; q < BETA; ++q) Lown[r][q] = 0.0;
    TF_UNROLL for (int a = 0; a < BETA; ++a)
      TF_UNROLL for (int q = 0; q < BETA; ++q) Lnext[a][q] = 0.0;
    TF_UNROLL for (int k = 0; k < C; ++k) {
      const double piv = T[k][BETA];
      if (!(piv != 0.0) || !(fabs(piv) < 1e300)) bad |= 1;
      // no pivoting here (SuperLU row-pivots): a pivot far smaller than the entries it
      // eliminates means multipliers > 1e10 and a useless factor -- say so (status bit 3)
      {
        double mx = 0.0;
        TF_UNROLL for (int r = k + 1; r < RT && r <= k + BETA; ++r) mx = fmax(mx, fabs(T[r][BETA + k - r]));
        if (fabs(piv) * TF_GROWTH_LIMIT < mx) bad |= 8;
      }
      const double inv = TF_RCP(piv);
      Uf[k][0] = inv;
      TF_UNROLL for (int c = 1; c <= BETA; ++c) Uf[k][c] = T[k][BETA + c];
      TF_UNROLL for (int r = k + 1; r < RT && r <= k + BETA; ++r) {
        const double l = T[r][BETA + k - r] * inv;
        if (r < C) Lown[r][r - k - 1] = l; else Lnext[r - C][r - k - 1] = l;
        TF_UNROLL for (int c = k + 1; c <= k + BETA; ++c)
          if (c - r >= -BETA && c - r <= BETA) T[r][BETA + c - r] -= l * T[k][BETA + c - k];
      }
    }
    TF_UNROLL for (int a = 0; a < BETA; ++a)
      TF_UNROLL for (int b = 0; b < BETA; ++b) Xout[a * BETA + b] = -T[C + a][BETA + b - a];
  }
};

// ------------------------------------------------------ substitution recurrences
// Forward:  y_r = f_r - sum_{q=1..BETA} L[r][q-1] y_{r-q};  state s[t] = y_{-1-t}.
template <int BETA, int C>
TF_HD TF_INLINE void fwd_chunk(const double (&L)[C][BETA], const double* f,
                               const double* s_in, double* y) {
  TF_UNROLL for (int r = 0; r < C; ++r) {
    double s = f[r];
    TF_UNROLL for (int q = 1; q <= BETA; ++q) {
      const double prev = (r - q >= 0) ? y[(r - q >= 0) ? r - q : 0] : s_in[(q - r - 1 >= 0) ? q - r - 1 : 0];
      s -= L[r][q - 1] * prev;
    }
    y[r] = s;
  }
}
// Affine map of the chunk for the forward recurrence (given y0 = run with s_in=0).
template <int BETA, int C>
TF_HD TF_INLINE void fwd_map(const double (&L)[C][BETA], const double* y0, AffMap<BETA>& m) {
  double zero[C];
  TF_UNROLL for (int r = 0; r < C; ++r) zero[r] = 0.0;
  TF_UNROLL for (int t = 0; t < BETA; ++t) {
    double e[BETA], h[C];
    TF_UNROLL for (int i = 0; i < BETA; ++i) e[i] = (i == t) ? 1.0 : 0.0;
    fwd_chunk<BETA, C>(L, zero, e, h);
    TF_UNROLL for (int i = 0; i < BETA; ++i) m.Phi()[i * BETA + t] = h[C - 1 - i];
  }
  TF_UNROLL for (int i = 0; i < BETA; ++i) m.c()[i] = y0[C - 1 - i];
}

// Backward: x_r = (y_r - sum_{q=1..BETA} Uf[r][q] x_{r+q}) * Uf[r][0]; s[t] = x_{C+t}.
template <int BETA, int C>
TF_HD TF_INLINE void bwd_chunk(const double (&Uf)[C][BETA + 1], const double* y,
                               const double* s_in, double* x) {
  TF_UNROLL for (int r = C - 1; r >= 0; --r) {
    double s = y[r];
    TF_UNROLL for (int q = 1; q <= BETA; ++q) {
      const double nxt = (r + q < C) ? x[(r + q < C) ? r + q : 0] : s_in[(r + q - C >= 0 && r + q < C + BETA) ? r + q - C : 0];
      s -= Uf[r][q] * nxt;
    }
    x[r] = s * Uf[r][0];
  }
}
template <int BETA, int C>
TF_HD TF_INLINE void bwd_map(const double (&Uf)[C][BETA + 1], const double* x0, AffMap<BETA>& m) {
  double zero[C];
  TF_UNROLL for (int r = 0; r < C; ++r) zero[r] = 0.0;
  TF_UNROLL for (int t = 0; t < BETA; ++t) {
    double e[BETA], h[C];
    TF_UNROLL for (int i = 0; i < BETA; ++i) e[i] = (i == t) ? 1.0 : 0.0;
    bwd_chunk<BETA, C>(Uf, zero, e, h);
    TF_UNROLL for (int i = 0; i < BETA; ++i) m.Phi()[i * BETA + t] = h[i];
  }
  TF_UNROLL for (int i = 0; i < BETA; ++i) m.c()[i] = x0[i];
}

}  // namespace tfb
