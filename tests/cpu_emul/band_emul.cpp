// CPU emulation of the chunked scan solver using the same chunk math as the
// kernels (triflow_b200/csrc/tf_band.h).  Test infrastructure only.
// Solves a random banded system chunk by chunk: run1 -> sequential "scan" with
// StarMap::combine -> run2 -> forward / backward affine scans, and returns the
// solution so the python side can compare with a dense solve.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "tf_band.h"

using namespace tfb;

template <int BETA, int C>
int solve_banded(int nchunks, const double* Aband /* n x (2BETA+1) */, const double* f, double* x,
                 double* Lout, double* Uout) {
  const int n = nchunks * C;
  constexpr int W = 2 * BETA + 1;
  auto Aat = [&](int r, int d) -> double {  // A(r, r+d)
    if (r < 0 || r >= n || r + d < 0 || r + d >= n) return 0.0;
    return Aband[r * W + BETA + d];
  };
  std::vector<StarMap<BETA>> maps(nchunks), excl(nchunks);
  int bad = 0;
  typedef double Rows[C + BETA][W];
  std::vector<double> rowsbuf((size_t)nchunks * (C + BETA) * W);
  for (int c = 0; c < nchunks; ++c) {
    Rows& A = *(Rows*)&rowsbuf[(size_t)c * (C + BETA) * W];
    for (int r = 0; r < C + BETA; ++r)
      for (int j = 0; j < W; ++j) A[r][j] = Aat(c * C + r, j - BETA);
    ChunkLU<BETA, C>::run1(A, maps[c], bad);
  }
  // exclusive scan
  StarMap<BETA> acc = StarMap<BETA>::identity();
  for (int c = 0; c < nchunks; ++c) {
    excl[c] = acc;
    acc = StarMap<BETA>::combine(acc, maps[c]);
  }
  std::vector<double> L((size_t)n * BETA, 0.0), U((size_t)n * (BETA + 1), 0.0);
  for (int c = 0; c < nchunks; ++c) {
    Rows& A = *(Rows*)&rowsbuf[(size_t)c * (C + BETA) * W];
    double Uf[C][BETA + 1], Lown[C][BETA], Lnext[BETA][BETA];
    ChunkLU<BETA, C>::run2(A, excl[c].P(), Uf, Lown, Lnext, bad);
    for (int r = 0; r < C; ++r) {
      for (int q = 0; q <= BETA; ++q) U[(size_t)(c * C + r) * (BETA + 1) + q] = Uf[r][q];
      for (int q = 1; q <= BETA; ++q)
        if (q <= r) L[(size_t)(c * C + r) * BETA + q - 1] = Lown[r][q - 1];
    }
    if (c + 1 < nchunks)
      for (int a = 0; a < BETA; ++a)
        for (int q = 1; q <= BETA; ++q)
          if (q > a) L[(size_t)((c + 1) * C + a) * BETA + q - 1] = Lnext[a][q - 1];
  }
  if (Lout) memcpy(Lout, L.data(), L.size() * sizeof(double));
  if (Uout) memcpy(Uout, U.data(), U.size() * sizeof(double));
  // forward
  std::vector<double> y(n);
  {
    std::vector<AffMap<BETA>> am(nchunks);
    for (int c = 0; c < nchunks; ++c) {
      double (*Lc)[BETA] = (double (*)[BETA]) & L[(size_t)c * C * BETA];
      double s0[BETA] = {0}, y0[C];
      fwd_chunk<BETA, C>(*(double (*)[C][BETA])Lc, f + c * C, s0, y0);
      fwd_map<BETA, C>(*(double (*)[C][BETA])Lc, y0, am[c]);
    }
    AffMap<BETA> a = AffMap<BETA>::identity();
    for (int c = 0; c < nchunks; ++c) {
      double (*Lc)[BETA] = (double (*)[BETA]) & L[(size_t)c * C * BETA];
      fwd_chunk<BETA, C>(*(double (*)[C][BETA])Lc, f + c * C, a.c(), &y[c * C]);
      a = AffMap<BETA>::combine(a, am[c]);
    }
  }
  // backward (scan order reversed)
  {
    std::vector<AffMap<BETA>> am(nchunks);
    for (int c = 0; c < nchunks; ++c) {
      double (*Uc)[BETA + 1] = (double (*)[BETA + 1]) & U[(size_t)c * C * (BETA + 1)];
      double s0[BETA] = {0}, x0[C];
      bwd_chunk<BETA, C>(*(double (*)[C][BETA + 1])Uc, &y[c * C], s0, x0);
      bwd_map<BETA, C>(*(double (*)[C][BETA + 1])Uc, x0, am[c]);
    }
    AffMap<BETA> a = AffMap<BETA>::identity();
    for (int c = nchunks - 1; c >= 0; --c) {
      double (*Uc)[BETA + 1] = (double (*)[BETA + 1]) & U[(size_t)c * C * (BETA + 1)];
      bwd_chunk<BETA, C>(*(double (*)[C][BETA + 1])Uc, &y[c * C], a.c(), &x[c * C]);
      a = AffMap<BETA>::combine(a, am[c]);
    }
  }
  return bad;
}

extern "C" int band_solve(int beta, int C, int nchunks, const double* A, const double* f, double* x,
                          double* L, double* U) {
#define CASE(B, CC) if (beta == B && C == CC) return solve_banded<B, CC>(nchunks, A, f, x, L, U);
  CASE(1, 4) CASE(1, 8) CASE(2, 8) CASE(2, 4) CASE(3, 8) CASE(5, 8) CASE(5, 6) CASE(2, 2) CASE(4, 4)
  return -1;
}
