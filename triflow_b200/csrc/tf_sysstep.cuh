// System-resident implicit step: one CTA owns one whole system for the whole step.
//
// For ensembles of small systems (one system fits one CTA: N <= 32 * MAXW * M nodes,
// config 5 of BASELINE.json) the per-kernel pipeline of tf_kernels.cuh moves
// Q = 8[(1+s)B + (1+4s+s(s-1)/2)v] bytes per node and step through HBM (factor written
// once, read once per stage, every stage vector written and read back).  None of this
// has to leave the SM: here the state U and the stage vectors k_j live in shared memory,
// the banded factor and the forward-substitution result live in the registers of the
// thread that owns the rows, and a step reads U once and writes U+ once (16 bytes per
// node instead of 224 for ROS3PRw).  The algorithm is the one of the per-kernel path
// (same chunking, same two-pass chunk scans) with one simplification: a non-periodic system
// is purely banded, so the border block the pipeline shares with periodic systems (last P
// nodes ordered last, Schur complement, x_b) does not exist here.  The two paths agree to
// rounding; tests compare them to 1e-12.
//
// Replaces, in one launch (reference file:line): compute_J_numpy + I - gamma*dt*J +
// factorized (compilers.py:292-332, schemes.py:146-149), the s stages of
// ROW_general._fixed_step (schemes.py:150-163), update and error norm (:164-174), and the
// Theta step (:548-559) as the one-stage case.
//
// Scope: scalar models with a tridiagonal Jacobian (V == 1, P == 1), non-periodic, as many
// stages as fit shared memory ((s + 2) vectors of the system: s <= 4 at N = 4096, every
// tableau (s <= 6) up to N = 2048).  Everything else keeps the per-kernel path.
#pragma once

#if (TF_NVAR == 1) && (TF_P == 1)
#define TF_HAS_SYSSTEP 1


#ifdef TF_TRACE
#define SYS_CLK(ph) do { if (threadIdx.x == 0) { const long long t_ = clock64(); tf_trace[(blockIdx.x & 1023) * 32 + (ph)] += (unsigned long long)(t_ - sh.t0); sh.t0 = t_; } } while (0)
#else
#define SYS_CLK(ph) do { } while (0)
#endif

namespace tfk {

// Exclusive prefix of every thread's element over the CTA (one tile, no look-back).
// REV: the order runs from the last thread to the first (backward substitution); the
// combine tree is the mirror image of the forward one, i.e. the tree the per-kernel
// backward sweep builds with its reversed thread -> chunk assignment.
template <class Mon, bool REV>
__device__ __forceinline__ Mon cta_scan(const Mon& mine, double* smem, const int area) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int ml = REV ? 31 - lane : lane;               // position in scan order
  const int mw = REV ? nwarps - 1 - warp : warp;
  Mon incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const Mon o = REV ? shfl_down(incl, d) : shfl_up(incl, d);
    const Mon c = Mon::combine(o, incl);
    incl = select(ml >= d, c, incl);
  }
  Mon excl = REV ? shfl_down(incl, 1) : shfl_up(incl, 1);
  excl = select(ml == 0, Mon::identity(), excl);
  // two scratch areas used in turn (factor 0, forward 1, backward 0, forward 1, ...): the
  // barrier of the scan in between orders the reads of one use against the writes of the
  // next one, so a scan costs one barrier
  smem += area * (MAXW * KMAX);
  if (ml == 31) {
#pragma unroll
    for (int k = 0; k < Mon::K; ++k) smem[mw * Mon::K + k] = incl.d[k];
  }
  __syncthreads();
  Mon w = Mon::identity();
  if (lane < nwarps) {
#pragma unroll
    for (int k = 0; k < Mon::K; ++k) w.d[k] = smem[lane * Mon::K + k];
  }
  // warp totals: lanes >= nwarps hold the identity, levels d >= nwarps change nothing below
  Mon wi = w;
#pragma unroll
  for (int d = 1; d < MAXW; d <<= 1) {
    const Mon o = shfl_up(wi, d);
    const Mon c = Mon::combine(o, wi);
    wi = select(lane >= d, c, wi);
  }
  Mon we = shfl_idx(wi, mw > 0 ? mw - 1 : 0);
  we = select(mw == 0, Mon::identity(), we);
  // (the pipeline's tile_scan composes an identity carry in front: exact, hence omitted)
  return Mon::combine(we, excl);
}

constexpr int SYS_RPC = C + BETA;            // rows a chunk evaluates: its own + BETA of the next
// chunk 0 + the chunks from the first one that reaches the last P nodes to the end of the
// padding: at most 32 (a whole warp-block of padding) + ceil((2P + M + EX) / M) + 1 of them
constexpr int SYS_EDGE_CHUNKS = 34 + (2 * P + M + EX + M - 1) / M;

struct SysShared {
  double scan[2 * MAXW * KMAX];
  double edge[SYS_EDGE_CHUNKS * SYS_RPC * WB];  // band rows of the chunks that touch the domain ends
  double cst[NC2];
  double lnext[NT * BETA * BETA];      // L multipliers a chunk leaves on the next chunk's rows
  double err[MAXW];
  unsigned long long bar[2];
  long long t0;                        // phase clock of the TF_TRACE build
};

// Band row of A = I - a*J for node i at the ends of the domain, for ONE node and V == 1.
// A non-periodic system is purely banded: a stencil neighbour outside the domain is the end
// node itself (edge replication, compilers.py:261-264,311-319), so its Jacobian entry is
// added onto the column of the end node -- inside the band.  The per-kernel pipeline orders
// the last P nodes last and eliminates them through a border block because it shares that
// code with periodic systems; here no border block, no Schur complement and no x_b exist.
// Entries of one row are summed in kk order, like the reference's COO -> CSC duplicate sum.
// Every such row gets its own thread in the pre-pass below instead of being walked by the
// two threads that own the end chunks while 15 warps wait.
__device__ __noinline__ void sys_edge_row(int i, const Geom& g, const double* sU, const Buf& lb,
                                          int sys, double a, const double* cst,
                                          double* out /* [WB] */) {
  static_assert(V == 1, "rows == nodes");
  double rows[WB];
#pragma unroll
  for (int d = 0; d < WB; ++d) rows[d] = 0.0;
  const int npad = g.nblk * 32 * M;
  if (i >= g.N && i < npad) rows[BETA] = 1.0;   // padding: identity (beyond npad: no row)
  if (i < g.N) {
    double jv[NNZ];
    TfNodeIn in;
#pragma unroll
    for (int o = 0; o < TF_WW; ++o) {
      const int j = map_node(i - P + o, g);
      in.w[0][o] = sU[(int)vidx(j, 0)];
#pragma unroll
      for (int h = 0; h < NH; ++h)
        in.w[V + h][o] = lb.H[(sys * (long long)NH + h) * hstride(g) + nidx(j)];
    }
#if TF_NNODEPAR > 0
#pragma unroll
    for (int q = 0; q < TF_NNODEPAR; ++q)
      in.np[q] = lb.NP[(sys * (long long)TF_NNODEPAR + q) * hstride(g) + nidx(i)];
#endif
#if TF_USES_X
    in.x = lb.X[nidx(i)];
#else
    in.x = 0.0;
#endif
    tf_model_J<FD>(cst, in, jv);
#pragma unroll
    for (int kk = 0; kk < NNZ; ++kk) {
      const int d = map_node(i + tf_j_off(kk), g) - i;       // clamped neighbour, |d| <= P
#pragma unroll
      for (int dd = -BETA; dd <= BETA; ++dd) rows[BETA + dd] += (d == dd) ? jv[kk] : 0.0;
    }
#pragma unroll
    for (int d = 0; d < WB; ++d) {
      const double sv = __dmul_rn(a, rows[d]);
      rows[d] = (d == BETA) ? __dsub_rn(1.0, sv) : -sv;
    }
  }
#pragma unroll
  for (int d = 0; d < WB; ++d) out[d] = rows[d];
}

// first chunk (>= 1) that is not "all regular"; chunk 0 never is
__device__ __forceinline__ int sys_first_tail_chunk(const Geom& g) {
  constexpr int NODES = M + EX;
  const int q = g.N - P - NODES;
  const int c = q < 0 ? 1 : q / M + 1;
  return c < 1 ? 1 : c;
}
__device__ __forceinline__ int sys_edge_slot(int chunk, int cfirst) {
  return chunk == 0 ? 0 : 1 + (chunk - cfirst);
}

// ---- factorisation of the thread's rows into registers (factor_body_stream, V == 1)
__device__ __forceinline__ void sys_factor(const Geom& g, const Buf& lb, int sys, double a,
                                           const double* cst, SysShared& sh,
                                           double (&Lr)[C][BETA], double (&Ur)[C][BETA + 1],
                                           int& bad) {
  constexpr int NODES = M + EX;
  constexpr int NSB = C / BETA;
  const int chunk = threadIdx.x;
  const int i0 = chunk * M;
  double win[NF][NODES + 2 * P];
  Star mine = Star::identity();
  const bool allreg = i0 >= P && i0 + NODES <= g.N - P;   // no stencil leaves the domain
  // pre-pass: the rows of the end chunks, one row per thread
  const int cfirst = sys_first_tail_chunk(g);
  const int nchunks = (int)blockDim.x;
  const int nedge = 1 + (nchunks - cfirst);                   // chunk 0 + tail chunks
  if (nedge > SYS_EDGE_CHUNKS) asm volatile("trap;");         // cannot happen (bound above)
  // rows are dealt to the warps first (row k -> warp k mod nwarps): the row kinds (regular,
  // top, bottom, border, padding) take different branches, which would serialise inside a warp
  // (warp 0 is left out when there are others: its thread 0 is issuing the next system's TMA)
  const int nwarps_a = (int)blockDim.x >> 5;
  const int nwarps_e = nwarps_a > 1 ? nwarps_a - 1 : 1, w_e = (int)(threadIdx.x >> 5) - (nwarps_a > 1 ? 1 : 0);
  for (int idx = (w_e >= 0 ? (int)(threadIdx.x & 31) * nwarps_e + w_e : nedge * SYS_RPC);
       idx < nedge * SYS_RPC; idx += 32 * nwarps_e) {
    const int slot = idx / SYS_RPC, m = idx - slot * SYS_RPC;
    const int c = slot == 0 ? 0 : cfirst + slot - 1;
    sys_edge_row(c * M + m, g, lb.U + sys * vstride(g), lb, sys, a, cst, sh.edge + idx * WB);
  }
  __syncthreads();
  SYS_CLK(5);                                            // edge-row pre-pass
  const double* erow = sh.edge + sys_edge_slot(chunk, cfirst) * SYS_RPC * WB;
  // the hot code only knows the regular rule; the generic one lives in sys_edge_row (keeping
  // both in this function cost every thread 20 % through register allocation)
  auto get_row = [&](double (&row)[WB], int m) {
    if (!allreg) {
#pragma unroll
      for (int d = 0; d < WB; ++d) row[d] = erow[m * WB + d];
    } else {
      node_row<NODES>(row, win, m, i0, g, lb, sys, a, cst, true);
    }
  };
  if (allreg) load_windows<NODES, 0>(win, i0, g, lb, sys, nullptr);
  {
    double cur[BETA][WB], nxt[BETA][WB];
#pragma unroll
    for (int r = 0; r < BETA; ++r) get_row(cur[r], r);
#pragma unroll
    for (int k = 0; k < NSB; ++k) {
#pragma unroll
      for (int r = 0; r < BETA; ++r)
        get_row(nxt[r], (k + 1) * BETA + r);
      double Dh[BETA * BETA], Z[BETA * 2 * BETA], Rr[BETA * BETA];
#pragma unroll
      for (int r = 0; r < BETA; ++r)
#pragma unroll
        for (int c = 0; c < BETA; ++c) {
          Dh[r * BETA + c] = cur[r][BETA + c - r] - mine.P()[r * BETA + c];
          Z[r * 2 * BETA + c] = (c <= r) ? cur[r][BETA + BETA + c - r] : 0.0;
          Z[r * 2 * BETA + BETA + c] = mine.Q()[r * BETA + c];
          Rr[r * BETA + c] = (c >= r) ? nxt[r][BETA + c - BETA - r] : 0.0;
        }
      tfb::solve_inplace<BETA, 2 * BETA>(Dh, Z);
      double Z1[BETA * BETA], Z2[BETA * BETA];
#pragma unroll
      for (int r = 0; r < BETA; ++r)
#pragma unroll
        for (int c = 0; c < BETA; ++c) {
          Z1[r * BETA + c] = Z[r * 2 * BETA + c];
          Z2[r * BETA + c] = Z[r * 2 * BETA + BETA + c];
        }
      Star nx;
      tfb::mm<BETA>(Rr, Z1, nx.P());
      tfb::mm<BETA>(Rr, Z2, nx.Q());
#pragma unroll
      for (int q = 0; q < BETA * BETA; ++q) nx.R()[q] = mine.R()[q];
      tfb::mma<BETA>(mine.S(), Z2, nx.R());
      tfb::mm<BETA>(mine.S(), Z1, nx.S());
      mine = nx;
#pragma unroll
      for (int r = 0; r < BETA; ++r)
#pragma unroll
        for (int d = 0; d < WB; ++d) cur[r][d] = nxt[r][d];
    }
#pragma unroll
    for (int q = 0; q < Star::K; ++q)
      if (!(fabs(mine.d[q]) < 1e300)) bad |= 1;
  }
  SYS_CLK(6);                                            // factor pass 1
  const Star pre = cta_scan<Star, false>(mine, sh.scan, 0);
  SYS_CLK(7);                                            // factor scan
  double X[BETA * BETA];
#pragma unroll
  for (int q = 0; q < BETA * BETA; ++q) X[q] = pre.P()[q];
  double Lprev[BETA][BETA];
#pragma unroll
  for (int r = 0; r < BETA; ++r)
#pragma unroll
    for (int q = 0; q < BETA; ++q) Lprev[r][q] = 0.0;
  double cur[BETA][WB], nxt[BETA][WB];
#pragma unroll
  for (int r = 0; r < BETA; ++r) get_row(cur[r], r);
#pragma unroll
  for (int k = 0; k < NSB; ++k) {
#pragma unroll
    for (int r = 0; r < BETA; ++r)
      get_row(nxt[r], (k + 1) * BETA + r);
    double A2[2 * BETA][WB];
#pragma unroll
    for (int r = 0; r < BETA; ++r)
#pragma unroll
      for (int d = 0; d < WB; ++d) {
        A2[r][d] = cur[r][d];
        A2[BETA + r][d] = (BETA + r + d - BETA < BETA) ? nxt[r][d] : 0.0;
      }
    double Uf[BETA][BETA + 1], Lown[BETA][BETA], Lnext[BETA][BETA], Xo[BETA * BETA];
    tfb::ChunkLU<BETA, BETA>::run2x(A2, X, Uf, Lown, Lnext, Xo, bad);
#pragma unroll
    for (int r = 0; r < BETA; ++r) {
      const int row = k * BETA + r;
#pragma unroll
      for (int q = 0; q <= BETA; ++q) Ur[row][q] = Uf[r][q];
#pragma unroll
      for (int q = 1; q <= BETA; ++q) {
        if (q <= r) Lr[row][q - 1] = Lown[r][q - 1];
        else if (k > 0) Lr[row][q - 1] = Lprev[r][q - 1];
        else Lr[row][q - 1] = 0.0;                // filled from the previous chunk below
      }
    }
#pragma unroll
    for (int r = 0; r < BETA; ++r)
#pragma unroll
      for (int q = 0; q < BETA; ++q) Lprev[r][q] = Lnext[r][q];
#pragma unroll
    for (int q = 0; q < BETA * BETA; ++q) X[q] = Xo[q];
#pragma unroll
    for (int r = 0; r < BETA; ++r)
#pragma unroll
      for (int d = 0; d < WB; ++d) cur[r][d] = nxt[r][d];
  }
  // multipliers of the next chunk's first BETA rows with respect to this chunk's pivots
#pragma unroll
  for (int r = 0; r < BETA; ++r)
#pragma unroll
    for (int q = 0; q < BETA; ++q) sh.lnext[(threadIdx.x * BETA + r) * BETA + q] = Lprev[r][q];
  __syncthreads();
  if (chunk > 0) {
#pragma unroll
    for (int r = 0; r < BETA; ++r)
#pragma unroll
      for (int q = r + 1; q <= BETA; ++q)
        Lr[r][q - 1] = sh.lnext[((threadIdx.x - 1) * BETA + r) * BETA + q - 1];
  }
}

// ---- one Rosenbrock stage: forward substitution of dt*F(U_i) + sum cfac_j k_j, border
//      solve, backward substitution; k_i -> shared memory, or (last stage) U+ -> HBM.
template <int I, bool LAST>
__device__ __forceinline__ void sys_stage(const Geom& g, const Buf& b, const Buf& lb, int sys,
                                          const TfStepDesc& sd, double dt, const double* cst,
                                          SysShared& sh, const double* sU, double* sK, double* sS,
                                          const double (&Lr)[C][BETA], const double (&Ur)[C][BETA + 1],
                                          double& emax) {
  const int T = blockDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = T >> 5;
  const int chunk = threadIdx.x;
  const int i0 = chunk * M;
  const int sb = (warp * C) * 32 + lane;               // own chunk inside the system
  const long long vs = vstride(g);
  Stage st;
  st.nprev = I;
  st.istage = I;
#pragma unroll
  for (int q = 0; q < MAXS; ++q) {
    st.alpha[q] = q < I ? sd.alpha[I][q] : 0.0;
    st.cfac[q] = q < I ? sd.cfac[I][q] : 0.0;
  }
  double y[C];
  // ---------------------------------------------------------------- forward
  {
    double own[C];
#pragma unroll
    for (int r = 0; r < C; ++r) own[r] = stage_value<I>(sU, lb, sys * vs, sb + (long long)r * 32, &st);
    const double* sH = sU;                              // where neighbours find the stage state
    if (I > 0) {
#pragma unroll
      for (int r = 0; r < C; ++r) sS[sb + r * 32] = own[r];
      sH = sS;
    }
    double win[NF][M + 2 * P];
    // Stencil windows.  Without helper fields every chunk takes its halo from the neighbouring
    // threads' values in shared memory; the chunks at the ends of the (non-periodic) domain
    // then replace what lies outside by the end values (edge replication, map_node).  The
    // generic window loader stays out of this function: it cost the end threads ~2000 cycles
    // per stage while everybody else waited at the scan barrier.
    constexpr bool FASTWIN = (NH == 0) && (P <= M);
    bool interior = true;
    if (!FASTWIN) {
      interior = i0 >= P && i0 + M + P <= g.N && P <= M && NH == 0 && chunk > 0 && chunk < T - 1;
      if (!interior) {
        Buf eb = lb;                                    // edge chunks: clamped window
        eb.U = const_cast<double*>(sU) - sys * vs;
        load_windows<M, I>(win, i0, g, eb, sys, &st);
      }
    }
    if (I > 0) __syncthreads();
    if (interior) {
#pragma unroll
      for (int w = 0; w < M + 2 * P; ++w) {
        const int rel = w - P;
        const int dc = rel < 0 ? -1 : (rel >= M ? 1 : 0);
        const int m = rel - dc * M;
        const int t2 = (int)threadIdx.x + dc;
        const int nb = ((t2 >> 5) * C) * 32 + (t2 & 31);
#pragma unroll
        for (int e = 0; e < V; ++e)
          win[e][w] = (dc == 0) ? own[(m * V + e) < C ? (m * V + e) : 0]
                                : ((t2 >= 0 && t2 < T) ? sH[nb + (m * V + e) * 32] : 0.0);
      }
      if (FASTWIN && (i0 < P || i0 + M + P > g.N)) {      // domain ends and padding
#pragma unroll
        for (int w = 0; w < M + 2 * P; ++w) {
          const int j = i0 - P + w;
          if (j < 0 || j >= g.N) {
            const int jc = j < 0 ? 0 : g.N - 1;
#pragma unroll
            for (int e = 0; e < V; ++e) win[e][w] = sH[(int)vidx(jc, e)];
          }
        }
      }
    }
    RecState rs;
    rs.init();
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const int i = i0 + m;
      double fe[V];
#pragma unroll
      for (int e = 0; e < V; ++e) fe[e] = 0.0;
      {                                   // padding nodes too (finite fill, rhs zeroed below)
        TfNodeIn in;
        node_inputs<M>(in, win, m, i, g, b, sys);
        tf_model_F_solver<FD>(cst, in, fe);
      }
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const int r = m * V + e;
        double rhs = __dmul_rn(dt, fe[e]);
#pragma unroll
        for (int q = 0; q < I; ++q) rhs = __fma_rn(st.cfac[q], sK[q * (T * C) + sb + r * 32], rhs);
        rhs = (i < g.N) ? rhs : 0.0;
        y[r] = rhs;
        double coef[BETA];
#pragma unroll
        for (int q = 0; q < BETA; ++q) coef[q] = Lr[r][q];
        rs.step(coef, rhs, 1.0);
      }
    }
    Aff mine;
    rs.to_map(mine);
    SYS_CLK(8);                                          // stage: state, halo, F, first pass
    const Aff pre = cta_scan<Aff, false>(mine, sh.scan, 1);
    SYS_CLK(9);                                          // forward scan
    double sv[BETA];
#pragma unroll
    for (int t = 0; t < BETA; ++t) sv[t] = pre.c()[t];
#pragma unroll
    for (int r = 0; r < C; ++r) {
      double v = y[r];
#pragma unroll
      for (int q = 0; q < BETA; ++q) v -= Lr[r][q] * sv[q];
#pragma unroll
      for (int q = BETA - 1; q > 0; --q) sv[q] = sv[q - 1];
      sv[0] = v;
      y[r] = v;
    }
  }
  // --------------------------------------------------------------- backward
  auto load_y = [&](int r) -> double { return y[r]; };
  Aff mine;
  {
    RecState rs;
    rs.init();
#pragma unroll
    for (int r = C - 1; r >= 0; --r) {
      double coef[BETA];
#pragma unroll
      for (int q = 0; q < BETA; ++q) coef[q] = Ur[r][q + 1];
      rs.step(coef, load_y(r), Ur[r][0]);
    }
    rs.to_map(mine);
  }
  SYS_CLK(10);                                           // fwd pass 2, border, bwd pass 1
  const Aff pre = cta_scan<Aff, true>(mine, sh.scan, 0);
  SYS_CLK(11);                                           // backward scan
  double sv[BETA];
#pragma unroll
  for (int t = 0; t < BETA; ++t) sv[t] = pre.c()[t];
#pragma unroll
  for (int r = C - 1; r >= 0; --r) {
    double v = load_y(r);
#pragma unroll
    for (int q = 0; q < BETA; ++q) v -= Ur[r][q + 1] * sv[q];
    v *= Ur[r][0];
#pragma unroll
    for (int q = BETA - 1; q > 0; --q) sv[q] = sv[q - 1];
    sv[0] = v;
    double k = v;
    double kprev[I > 0 ? I : 1];
#pragma unroll
    for (int q = 0; q < I; ++q) {
      kprev[q] = sK[q * (T * C) + sb + r * 32];
      k = __fma_rn(-st.cfac[q], kprev[q], k);
    }
    if (!LAST) {
      sK[I * (T * C) + sb + r * 32] = k;
    } else {
      double acc = 0.0, accp = 0.0;
#pragma unroll
      for (int q = 0; q <= I; ++q) {
        const double kq = (q < I) ? kprev[q < I ? q : 0] : k;
        const double t = __dmul_rn(sd.b[q], kq);
        acc = (q == 0) ? t : __dadd_rn(acc, t);
        const double tp = __dmul_rn(sd.bp[q], kq);
        accp = (q == 0) ? tp : __dadd_rn(accp, tp);
      }
      const double un = __dadd_rn(sU[sb + r * 32], acc);
      b.Un[sys * vs + sb + (long long)r * 32] = un;
      if (sd.has_pred) {
        const double e = fabs(__dsub_rn(un, __dadd_rn(un, accp)));
        emax = (e > emax || e != e) ? e : emax;
      }
    }
  }
  (void)nwarps;
}


// Run-time stage index for the long tableaux (s > 3: ROS3PRL, RODASPR).
// IMAX bounds the number of earlier stage vectors (compile time, loops unrolled and
// predicated on the run-time stage index I <= IMAX): one instance serves every stage,
// which keeps the kernel inside the instruction cache.
template <int IMAX>
__device__ __forceinline__ void sys_stage_rt(const int I, const bool LAST, const Geom& g, const Buf& b,
                                          const Buf& lb, int sys, const TfStepDesc& sd, double dt,
                                          const double* cst, SysShared& sh, const double* sU,
                                          double* sK, double* sS, const double (&Lr)[C][BETA],
                                          const double (&Ur)[C][BETA + 1], double& emax) {
  const int T = blockDim.x;
  const int TC = T * C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunk = threadIdx.x;
  const int i0 = chunk * M;
  const int sb = (warp * C) * 32 + lane;               // own chunk inside the system
  const long long vs = vstride(g);
  double alpha[IMAX > 0 ? IMAX : 1], cfac[IMAX > 0 ? IMAX : 1];
#pragma unroll
  for (int q = 0; q < IMAX; ++q) {
    alpha[q] = q < I ? sd.alpha[I][q] : 0.0;
    cfac[q] = q < I ? sd.cfac[I][q] : 0.0;
  }
  double y[C];
  // ---------------------------------------------------------------- forward
  {
    // stage state U + ((alpha_0 k_0 + alpha_1 k_1) + ...), the reference's summation order
    double own[C];
#pragma unroll
    for (int r = 0; r < C; ++r) {
      double u = sU[sb + r * 32];
      if (IMAX > 0 && I > 0) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < IMAX; ++q)
          if (q < I) {
            const double term = __dmul_rn(alpha[q], sK[q * TC + sb + r * 32]);
            acc = (q == 0) ? term : __dadd_rn(acc, term);
          }
        u = __dadd_rn(u, acc);
      }
      own[r] = u;
    }
    const double* sH = sU;                              // where neighbours find the stage state
    if (I > 0) {
#pragma unroll
      for (int r = 0; r < C; ++r) sS[sb + r * 32] = own[r];
      sH = sS;
    }
    double win[NF][M + 2 * P];
    const bool interior = i0 >= P && i0 + M + P <= g.N && P <= M && NH == 0 && chunk > 0 &&
                          chunk < T - 1;
    if (!interior) {                                    // edge chunks: clamped window (rare path)
      Stage st;
      st.nprev = I;
#pragma unroll
      for (int q = 0; q < MAXS; ++q) st.alpha[q] = (q < IMAX && q < I) ? sd.alpha[I][q < IMAX ? q : 0] : 0.0;
      load_windows<M, -1>(win, i0, g, lb, sys, &st);
    }
    if (I > 0) __syncthreads();
    if (interior) {
#pragma unroll
      for (int w = 0; w < M + 2 * P; ++w) {
        const int rel = w - P;
        const int dc = rel < 0 ? -1 : (rel >= M ? 1 : 0);
        const int m = rel - dc * M;
        const int t2 = (int)threadIdx.x + dc;
        const int nb = ((t2 >> 5) * C) * 32 + (t2 & 31);
#pragma unroll
        for (int e = 0; e < V; ++e)
          win[e][w] = (dc == 0) ? own[(m * V + e) < C ? (m * V + e) : 0] : sH[nb + (m * V + e) * 32];
      }
    }
    RecState rs;
    rs.init();
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const int i = i0 + m;
      double fe[V];
#pragma unroll
      for (int e = 0; e < V; ++e) fe[e] = 0.0;
      {                                   // padding nodes too (finite fill, rhs zeroed below)
        TfNodeIn in;
        node_inputs<M>(in, win, m, i, g, b, sys);
        tf_model_F_solver<FD>(cst, in, fe);
      }
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const int r = m * V + e;
        double rhs = __dmul_rn(dt, fe[e]);
#pragma unroll
        for (int q = 0; q < IMAX; ++q)
          if (q < I) rhs = __fma_rn(cfac[q], sK[q * TC + sb + r * 32], rhs);
        rhs = (i < g.N) ? rhs : 0.0;
        y[r] = rhs;
        double coef[BETA];
#pragma unroll
        for (int q = 0; q < BETA; ++q) coef[q] = Lr[r][q];
        rs.step(coef, rhs, 1.0);
      }
    }
    Aff mine;
    rs.to_map(mine);
    const Aff pre = cta_scan<Aff, false>(mine, sh.scan, 1);
    double sv[BETA];
#pragma unroll
    for (int t = 0; t < BETA; ++t) sv[t] = pre.c()[t];
#pragma unroll
    for (int r = 0; r < C; ++r) {
      double v = y[r];
#pragma unroll
      for (int q = 0; q < BETA; ++q) v -= Lr[r][q] * sv[q];
#pragma unroll
      for (int q = BETA - 1; q > 0; --q) sv[q] = sv[q - 1];
      sv[0] = v;
      y[r] = v;
    }
  }
  // --------------------------------------------------------------- backward
  Aff mine;
  {
    RecState rs;
    rs.init();
#pragma unroll
    for (int r = C - 1; r >= 0; --r) {
      double coef[BETA];
#pragma unroll
      for (int q = 0; q < BETA; ++q) coef[q] = Ur[r][q + 1];
      rs.step(coef, y[r], Ur[r][0]);
    }
    rs.to_map(mine);
  }
  const Aff pre = cta_scan<Aff, true>(mine, sh.scan, 0);
  double sv[BETA];
#pragma unroll
  for (int t = 0; t < BETA; ++t) sv[t] = pre.c()[t];
#pragma unroll
  for (int r = C - 1; r >= 0; --r) {
    double v = y[r];
#pragma unroll
    for (int q = 0; q < BETA; ++q) v -= Ur[r][q + 1] * sv[q];
    v *= Ur[r][0];
#pragma unroll
    for (int q = BETA - 1; q > 0; --q) sv[q] = sv[q - 1];
    sv[0] = v;
    double k = v;
    double kprev[IMAX > 0 ? IMAX : 1];
#pragma unroll
    for (int q = 0; q < IMAX; ++q)
      if (q < I) {
        kprev[q] = sK[q * TC + sb + r * 32];
        k = __fma_rn(-cfac[q], kprev[q], k);
      }
    if (!LAST) {
      sK[I * TC + sb + r * 32] = k;
    } else {
      // U + ((b0 k0 + b1 k1) + ...) in the reference's summation order (schemes.py:164-170)
      double acc = 0.0, accp = 0.0;
#pragma unroll
      for (int q = 0; q <= IMAX; ++q)
        if (q <= I) {
          const double kq = (q < I) ? kprev[q < IMAX ? q : 0] : k;
          const double t = __dmul_rn(sd.b[q], kq);
          acc = (q == 0) ? t : __dadd_rn(acc, t);
          const double tp = __dmul_rn(sd.bp[q], kq);
          accp = (q == 0) ? tp : __dadd_rn(accp, tp);
        }
      const double un = __dadd_rn(sU[sb + r * 32], acc);
      b.Un[sys * vs + sb + (long long)r * 32] = un;
      if (sd.has_pred) {
        const double e = fabs(__dsub_rn(un, __dadd_rn(un, accp)));
        emax = (e > emax || e != e) ? e : emax;
      }
    }
  }
}

// all stages of one step
template <int IMAX>
__device__ __forceinline__ void sys_stages(const Geom& g, const Buf& b, const Buf& lb, int sys,
                                           const TfStepDesc& sd, double dt, const double* cst,
                                           SysShared& sh, const double* sU, double* sK, double* sS,
                                           const double (&Lr)[C][BETA],
                                           const double (&Ur)[C][BETA + 1], double& emax) {
#pragma unroll 1
  for (int I = 0; I < sd.s; ++I) {
    if (I > 0) __syncthreads();                         // k_{I-1} of the neighbours is in place
    sys_stage_rt<IMAX>(I, I == sd.s - 1, g, b, lb, sys, sd, dt, cst, sh, sU, sK, sS, Lr, Ur, emax);
  }
}

}  // namespace tfk

// One CTA per system, persistent over the systems of the batch (grid = resident CTAs).
// The next system's U is fetched by a bulk asynchronous copy (TMA) while the current
// one is stepped.
template <bool LONG_TABLEAU>
__device__ __forceinline__ void sysstep_body(const tfk::Geom& g, const tfk::Buf& b,
                                             const TfStepDesc& sd) {
  using namespace tfk;
  extern __shared__ __align__(128) double dsm_sys[];
  double* dsm = dsm_sys;
  __shared__ SysShared sh;
  const int T = blockDim.x;
  const int TC = T * C;                                  // doubles per vector of one system
  double* sUb[2] = {dsm, dsm + TC};
  const int nk = sd.s > 1 ? sd.s - 1 : 0;                // stage vectors kept in shared memory
  double* sK = dsm + 2 * TC;                             // [nk][TC]
  double* sS = dsm + (2 + nk) * TC;
  const long long vs = vstride(g);
  const unsigned bytes = (unsigned)(TC * sizeof(double));
  constexpr int CST_PT = (NC2 + 31) / 32;                // constants per thread (T >= 32)
  if (threadIdx.x == 0) { mbar_init(&sh.bar[0], 1); mbar_init(&sh.bar[1], 1); }
  __syncthreads();
  unsigned phase[2] = {0u, 0u};
  int it = 0;
#ifdef TF_TRACE
  if (threadIdx.x == 0) sh.t0 = clock64();
#endif
  auto is_active = [&](int s) { return b.active == nullptr || b.active[s] != 0; };
  // first system of this CTA
  int sys = blockIdx.x;
  while (sys < g.batch && !is_active(sys)) sys += gridDim.x;
  if (sys < g.batch && threadIdx.x == 0) {
    mbar_expect_tx(&sh.bar[0], bytes);
    bulk_g2s(sUb[0], b.U + sys * vs, bytes, &sh.bar[0]);
  }
  while (sys < g.batch) {
    const int cur = it & 1;
    int nxt = sys + gridDim.x;
    while (nxt < g.batch && !is_active(nxt)) nxt += gridDim.x;
    SYS_CLK(13);
    const double a = (b.asys != nullptr) ? b.asys[sys] : sd.a;
    const double dt = (b.dtsys != nullptr) ? b.dtsys[sys] : sd.dt;
    if (it == 0) {
      for (int k = threadIdx.x; k < NC2; k += T) sh.cst[k] = b.cst[(long long)sys * NC2 + k];
    }
    // The hand-over work between two systems is dealt to different warps so that it overlaps
    // (thread 0 alone took ~2000 cycles for it): warp 0 issues the TMA of the next U, warp
    // `w_err` writes the error estimate, warp `w_cst` fetches the next system's constants
    // (loaded now, stored behind the barrier that ends this system).
    const int nw_ = T >> 5, warp_ = (int)threadIdx.x >> 5, lane_ = (int)threadIdx.x & 31;
    const int w_err = nw_ > 1 ? 1 : 0, w_cst = nw_ > 2 ? 2 : 0;
    double cst_next[CST_PT];
#pragma unroll
    for (int k = 0; k < CST_PT; ++k) {
      const int idx = lane_ + k * 32;
      cst_next[k] = 0.0;                              // predicated load: nothing waits for it here
      if (warp_ == w_cst && nxt < g.batch && idx < NC2)
        cst_next[k] = __ldg(b.cst + (long long)nxt * NC2 + idx);
    }
    const double* sU = sUb[cur];
    Buf lb = b;                                          // U and k_j of this system: shared memory
    lb.U = const_cast<double*>(sU) - sys * vs;
#pragma unroll
    for (int q = 0; q < MAXS; ++q) lb.K[q] = (q < nk) ? sK + q * TC - sys * vs : nullptr;
    SYS_CLK(14);
    mbar_wait(&sh.bar[cur], phase[cur]);
    phase[cur] ^= 1u;
    SYS_CLK(15);
    __syncthreads();
    // next system's U: issued here (~900 cycles for the issuing thread) because warp 0 has
    // no row in the end-row pre-pass that follows; buffer cur^1 was released by the barrier
    // that ended the last system
    if (nxt < g.batch && threadIdx.x == 0) {
      mbar_expect_tx(&sh.bar[cur ^ 1], bytes);
      bulk_g2s(sUb[cur ^ 1], b.U + nxt * vs, bytes, &sh.bar[cur ^ 1]);
    }
    double Lr[C][BETA], Ur[C][BETA + 1];
    int bad = 0;
    SYS_CLK(0);                                          // wait for U (TMA) + constants
    sys_factor(g, lb, sys, a, sh.cst, sh, Lr, Ur, bad);
    SYS_CLK(1);
    if (bad) atomicOr(b.status + sys, bad);
    SYS_CLK(2);
    double emax = 0.0;
    if constexpr (LONG_TABLEAU) {
      sys_stages<MAXS - 1>(g, b, lb, sys, sd, dt, sh.cst, sh, sU, sK, sS, Lr, Ur, emax);
    } else {                                   // the common tableaux: stage index at compile time
      switch (sd.s) {
        case 1:
          sys_stage<0, true>(g, b, lb, sys, sd, dt, sh.cst, sh, sU, sK, sS, Lr, Ur, emax);
          break;
        case 2:
          sys_stage<0, false>(g, b, lb, sys, sd, dt, sh.cst, sh, sU, sK, sS, Lr, Ur, emax);
          __syncthreads();
          sys_stage<1, true>(g, b, lb, sys, sd, dt, sh.cst, sh, sU, sK, sS, Lr, Ur, emax);
          break;
        default:
          sys_stage<0, false>(g, b, lb, sys, sd, dt, sh.cst, sh, sU, sK, sS, Lr, Ur, emax);
          __syncthreads();
          sys_stage<1, false>(g, b, lb, sys, sd, dt, sh.cst, sh, sU, sK, sS, Lr, Ur, emax);
          __syncthreads();
          sys_stage<2, true>(g, b, lb, sys, sd, dt, sh.cst, sh, sU, sK, sS, Lr, Ur, emax);
          break;
      }
    }
    SYS_CLK(3);                                          // all stages
    // error estimate of the system
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (sd.has_pred) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, emax, d);
        emax = (o > emax || o != o) ? o : emax;
      }
      if (lane == 0) sh.err[warp] = emax;
    }
    __syncthreads();                                     // releases sU[cur], sK, sS, sh.*
#pragma unroll
    for (int k = 0; k < CST_PT; ++k) {
      const int idx = lane + k * 32;
      if (warp == w_cst && nxt < g.batch && idx < NC2) sh.cst[idx] = cst_next[k];
    }
    SYS_CLK(4);
    if (warp == w_err) {                                 // max over the warps (NaN sticks)
      double e = 0.0;
      if (sd.has_pred) {
        e = (lane < (T >> 5)) ? sh.err[lane] : 0.0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          const double o = __shfl_xor_sync(0xffffffffu, e, d);
          e = (o > e || o != o) ? o : e;
        }
      }
      if (lane == 0) b.err[sys] = e;
    }
    SYS_CLK(12);
    sys = nxt;
    ++it;
  }
}

// s <= 3 (ROS2, ROS3PRw, Theta): stage index at compile time
extern "C" __global__ void __launch_bounds__(tfk::NT, 1) tf_k_sysstep(tfk::Geom g, tfk::Buf b,
                                                                      TfStepDesc sd) {
  sysstep_body<false>(g, b, sd);
}
// s > 3 (ROS3PRL, RODASPR): one stage body with a run-time stage index
extern "C" __global__ void __launch_bounds__(tfk::NT, 1) tf_k_sysstep_rt(tfk::Geom g, tfk::Buf b,
                                                                         TfStepDesc sd) {
  sysstep_body<true>(g, b, sd);
}

#endif  // V == 1 && P == 1
