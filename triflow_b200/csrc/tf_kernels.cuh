// Model-specialised kernels of the implicit method-of-lines hot path (sm_100a).
//
// This file is compiled once per model: the generated model header (TF_NVAR,
// TF_P, ..., tf_model_F, tf_model_J; see triflow_b200/codegen.py) is included
// first, then this file, with -DTF_M=<nodes per thread> -DTF_WARPS=<warps/CTA>.
//
// Replaces, on the device (reference file:line):
//   compute_F_numpy / compute_J_numpy + ghost-cell padding  compilers.py:227-332
//   A = I - gamma*dt*J, factorized(A), luf(b)               schemes.py:148-163
//   stage combination, update, error norm                   schemes.py:153-174
//   Theta step                                              schemes.py:548-559
//
// Data layout in HBM ("lane-transposed chunks").  A system has N nodes, V = TF_NVAR
// unknowns per node, interleaved (u = i*V + e, the reference's uflat order).  Each
// thread owns a chunk of M consecutive nodes (C = M*V unknowns); a warp owns 32
// consecutive chunks (a "warp-block").  Element j of lane l of block b lives at
//     (b*C + j)*32 + l
// so that every per-thread sequential access is a fully coalesced 256-byte warp
// access with no shared-memory staging, and stencil halos are a neighbouring
// lane's elements (same cache lines).  Factors L (BETA per unknown), U (BETA+1 per
// unknown, pivot stored inverted) and all stage vectors use the same layout.
#pragma once
#include <stdint.h>
#include "tf_band.h"
#include "tf_params.h"

#ifndef TF_M
#define TF_M 8
#endif
#ifndef TF_WARPS
#define TF_WARPS 8
#endif

namespace tfk {
typedef TfGeom Geom; typedef TfBuf Buf; typedef TfStage Stage; typedef TfUpdate Update;

constexpr int V = TF_NVAR;
constexpr int P = TF_P;
constexpr int BETA = (P * V + V - 1) > 0 ? (P * V + V - 1) : 1;
constexpr int NB = (P * V) > 0 ? (P * V) : 1;     // border unknowns (last P nodes)
constexpr int M = TF_M;
constexpr int C = M * V;
constexpr int WARPS = TF_WARPS;
constexpr int NT = 32 * WARPS;
constexpr int WB = 2 * BETA + 1;
constexpr int EX = (BETA + V - 1) / V;             // extra nodes needed from the next chunk
constexpr int NF = TF_NFIELD;
constexpr int NH = TF_NHELP;
constexpr int NNZ = TF_NNZ;
constexpr int NC2 = 2 * (TF_NCONST > 0 ? TF_NCONST : 1);
constexpr int MAXS = TF_MAXS;
static_assert(C >= BETA, "chunk must hold at least BETA unknowns");
static_assert(P >= 1, "models without spatial stencil are not supported");
static_assert((M & (M - 1)) == 0, "TF_M must be a power of two");

typedef tfb::StarMap<BETA> Star;
typedef tfb::AffMap<BETA> Aff;
constexpr int KMAX = Star::K > Aff::K ? Star::K : Aff::K;

// ----------------------------------------------------------------- indexing
__device__ __forceinline__ long long vstride(const Geom& g) { return (long long)g.nblk * C * 32; }
__device__ __forceinline__ long long hstride(const Geom& g) { return (long long)g.nblk * M * 32; }
// unknown e of node i
__device__ __forceinline__ long long vidx(int i, int e) {
  const int chunk = i / M, m = i % M;
  return ((long long)(chunk >> 5) * C + m * V + e) * 32 + (chunk & 31);
}
// node-plane index (helpers, x, per-node parameters)
__device__ __forceinline__ long long nidx(int i) {
  const int chunk = i / M, m = i % M;
  return ((long long)(chunk >> 5) * M + m) * 32 + (chunk & 31);
}
// unknown row r (global unknown index) -> (chunk-layout base index)
__device__ __forceinline__ long long ridx(int r) { return vidx(r / V, r % V); }

// stencil neighbour: wrap (periodic) or clamp (edge replication)
__device__ __forceinline__ int map_node(int j, const Geom& g) {
  if (g.periodic) {
    if (j < 0) j += g.N;
    else if (j >= g.N) j -= g.N;
  }
  return j < 0 ? 0 : (j >= g.N ? g.N - 1 : j);
}

// ----------------------------------------------------------- scan machinery
template <class Mon>
__device__ __forceinline__ Mon shfl_up(const Mon& v, int d) {
  Mon o;
#pragma unroll
  for (int k = 0; k < Mon::K; ++k) o.d[k] = __shfl_up_sync(0xffffffffu, v.d[k], d);
  return o;
}
template <class Mon>
__device__ __forceinline__ Mon shfl_down(const Mon& v, int d) {
  Mon o;
#pragma unroll
  for (int k = 0; k < Mon::K; ++k) o.d[k] = __shfl_down_sync(0xffffffffu, v.d[k], d);
  return o;
}
template <class Mon>
__device__ __forceinline__ Mon shfl_idx(const Mon& v, int src) {
  Mon o;
#pragma unroll
  for (int k = 0; k < Mon::K; ++k) o.d[k] = __shfl_sync(0xffffffffu, v.d[k], src);
  return o;
}
template <class Mon>
__device__ __forceinline__ Mon select(bool c, const Mon& a, const Mon& b) {
  Mon o;
#pragma unroll
  for (int k = 0; k < Mon::K; ++k) o.d[k] = c ? a.d[k] : b.d[k];
  return o;
}
// inclusive scan over the lanes of a warp, lower lane = earlier
template <class Mon>
__device__ __forceinline__ Mon warp_scan(Mon v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const Mon o = shfl_up(v, d);
    const Mon c = Mon::combine(o, v);
    v = select(lane >= d, c, v);
  }
  return v;
}
// ordered reduction, HIGHER lane = earlier; result valid in lane 0
template <class Mon>
__device__ __forceinline__ Mon warp_reduce_rev(Mon v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const Mon o = shfl_down(v, d);
    const Mon c = Mon::combine(o, v);
    v = select(lane + d < 32, c, v);
  }
  return v;
}

__device__ __forceinline__ int ld_flag(const int* p) {
  return *((const volatile int*)p);
}

// Decoupled look-back over the tiles of one system.  Called by all 32 lanes of
// warp 0 with the tile aggregate; returns the exclusive prefix of the tile.
template <class Mon>
__device__ Mon lookback(const Mon& aggregate, const Buf& b, long long gbase, int tile, int lane) {
  int* flags = b.flags + 1 + gbase;
  double* agg = b.lbagg + (gbase + tile) * KMAX;
  double* inc = b.lbinc + (gbase + tile) * KMAX;
  if (tile == 0) {
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < Mon::K; ++k) inc[k] = aggregate.d[k];
      __threadfence();
      *((volatile int*)(flags + tile)) = 2;
    }
    return Mon::identity();
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < Mon::K; ++k) agg[k] = aggregate.d[k];
    __threadfence();
    *((volatile int*)(flags + tile)) = 1;
  }
  Mon prefix = Mon::identity();
  int look = tile - 1;
  while (true) {
    const int t = look - lane;            // lane 0 = nearest predecessor
    int f = 2;
    if (t >= 0) {
      do { f = ld_flag(flags + t); } while (f == 0);
    }
    __threadfence();
    const unsigned m2 = __ballot_sync(0xffffffffu, f == 2);
    const int kstop = __ffs(m2) - 1;      // nearest tile with an inclusive prefix
    Mon e = Mon::identity();
    if (t >= 0 && (kstop < 0 || lane <= kstop)) {
      const double* src = ((f == 2) ? b.lbinc : b.lbagg) + (gbase + t) * KMAX;
#pragma unroll
      for (int k = 0; k < Mon::K; ++k) e.d[k] = __ldcg(src + k);
    }
    Mon w = warp_reduce_rev(e, lane);
    w = shfl_idx(w, 0);
    prefix = Mon::combine(w, prefix);
    if (kstop >= 0) break;
    look -= 32;
  }
  if (lane == 0) {
    const Mon incl = Mon::combine(prefix, aggregate);
#pragma unroll
    for (int k = 0; k < Mon::K; ++k) inc[k] = incl.d[k];
    __threadfence();
    *((volatile int*)(flags + tile)) = 2;
  }
  return prefix;
}

// Exclusive prefix of every thread's element over (lane, warp, tile) order.
// `carry` (optional, smem-free) is an extra prefix applied before tile 0 handling
// when look-back is disabled (used by the border kernel's sequential tile loop).
template <class Mon>
__device__ Mon tile_scan(const Mon& mine, double* smem /* WARPS*K doubles */, bool use_lookback,
                         const Buf& b, long long gbase, int tile, const Mon& carry, Mon* tile_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const Mon incl = warp_scan(mine, lane);
  Mon excl = shfl_up(incl, 1);
  excl = select(lane == 0, Mon::identity(), excl);
  __syncthreads();                         // smem reuse across successive scans
  if (lane == 31) {
#pragma unroll
    for (int k = 0; k < Mon::K; ++k) smem[warp * Mon::K + k] = incl.d[k];
  }
  __syncthreads();
  if (warp == 0) {
    Mon w = Mon::identity();
    if (lane < WARPS) {
#pragma unroll
      for (int k = 0; k < Mon::K; ++k) w.d[k] = smem[lane * Mon::K + k];
    }
    const Mon wi = warp_scan(w, lane);
    Mon we = shfl_up(wi, 1);
    we = select(lane == 0, Mon::identity(), we);
    const Mon total = shfl_idx(wi, WARPS - 1);
    Mon tp = carry;
    if (use_lookback) tp = lookback(total, b, gbase, tile, lane);
    const Mon wp = Mon::combine(tp, we);
    __syncwarp();
    if (lane < WARPS) {
#pragma unroll
      for (int k = 0; k < Mon::K; ++k) smem[lane * Mon::K + k] = wp.d[k];
    }
    if (tile_total != nullptr && lane == 0) {
      // inclusive total of the tile including carry, parked after the warp slots
      const Mon tt = Mon::combine(tp, total);
#pragma unroll
      for (int k = 0; k < Mon::K; ++k) smem[WARPS * Mon::K + k] = tt.d[k];
    }
  }
  __syncthreads();
  Mon wp;
#pragma unroll
  for (int k = 0; k < Mon::K; ++k) wp.d[k] = smem[warp * Mon::K + k];
  if (tile_total != nullptr) {
#pragma unroll
    for (int k = 0; k < Mon::K; ++k) tile_total->d[k] = smem[WARPS * Mon::K + k];
  }
  return Mon::combine(wp, excl);
}

// Tile / system assignment.  One tile per system: blockIdx.  Otherwise a ticket,
// so that every tile a CTA may wait for is already running (look-back progress).
__device__ __forceinline__ void resolve_tile(const Geom& g, const Buf& b, int& sys, int& tile) {
  if (g.tiles == 1) {
    sys = blockIdx.x;
    tile = 0;
    return;
  }
  __shared__ int s_ticket;
  if (threadIdx.x == 0) s_ticket = atomicAdd(b.flags, 1);
  __syncthreads();
  const int t = s_ticket;
  sys = t / g.tiles;
  tile = t - sys * g.tiles;
}

// ------------------------------------------------------------ stencil windows
// Values of every field at nodes i0-P .. i0+NODES-1+P for the stage state
// U + sum_j alpha_j K_j (dependent variables) and the helper planes.
template <int NODES>
__device__ __forceinline__ void load_windows(double (&win)[NF][NODES + 2 * P], int i0, const Geom& g,
                                             const Buf& b, int sys, const Stage* st) {
  const double* U = b.U + sys * vstride(g);
#pragma unroll
  for (int w = 0; w < NODES + 2 * P; ++w) {
    const int j = map_node(i0 - P + w, g);
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const long long a = vidx(j, e);
      double u = U[a];
      if (st != nullptr && st->nprev > 0) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < MAXS; ++q)
          if (q < st->nprev) {
            const double term = __dmul_rn(st->alpha[q], b.K[q][sys * vstride(g) + a]);
            acc = (q == 0) ? term : __dadd_rn(acc, term);
          }
        u = __dadd_rn(u, acc);
      }
      win[e][w] = u;
    }
#pragma unroll
    for (int h = 0; h < NH; ++h)
      win[V + h][w] = b.H[(sys * (long long)NH + h) * hstride(g) + nidx(j)];
  }
}

template <int NODES>
__device__ __forceinline__ void node_inputs(TfNodeIn& in, const double (&win)[NF][NODES + 2 * P], int m,
                                            int i, const Geom& g, const Buf& b, int sys) {
#pragma unroll
  for (int f = 0; f < NF; ++f)
#pragma unroll
    for (int o = 0; o < TF_WW; ++o) in.w[f][o] = win[f][m + o];
#if TF_NNODEPAR > 0
#pragma unroll
  for (int q = 0; q < TF_NNODEPAR; ++q)
    in.np[q] = b.NP[(sys * (long long)TF_NNODEPAR + q) * hstride(g) + nidx(i)];
#endif
#if TF_USES_X
  in.x = b.X[nidx(i)];
#else
  in.x = 0.0;
#endif
}

// -------------------------------------------------------------- J -> A rows
// Rows of A = I - a*J for a node that touches the domain ends, the border (last
// P nodes) or the padding.  Dynamic indexing on purpose: rare path.
__device__ __noinline__ void assemble_special(int i, const Geom& g, const double* jv, double a,
                                              double* rows /* [V][WB] */, double* btab) {
  for (int k = 0; k < V * WB; ++k) rows[k] = 0.0;
  const int nint = g.N - P;                 // interior nodes
  if (i >= nint) {                          // border or padding: identity row in the band
    for (int e = 0; e < V; ++e) rows[e * WB + BETA] = 1.0;
    if (i >= g.N) return;
    double* Ft = btab + 2 * NB * NB;
    double* Fb = btab + 3 * NB * NB;
    double* Ab = btab + 4 * NB * NB;
    for (int e = 0; e < V; ++e) {
      const int r = (i - nint) * V + e;
      for (int c = 0; c < NB; ++c) { Ft[r * NB + c] = 0.0; Fb[r * NB + c] = 0.0; Ab[r * NB + c] = 0.0; }
    }
    for (int kk = 0; kk < NNZ; ++kk) {
      const int e = tf_j_eq(kk), var = tf_j_var(kk), off = tf_j_off(kk);
      const int r = (i - nint) * V + e;
      const int j = map_node(i + off, g);
      if (j >= nint) Ab[r * NB + (j - nint) * V + var] += jv[kk];
      else if (j < P) Ft[r * NB + j * V + var] += jv[kk];
      else Fb[r * NB + (j - (g.N - 2 * P)) * V + var] += jv[kk];
    }
    return;
  }
  double* Et = btab;
  double* Eb = btab + NB * NB;
  const bool top = i < P, bot = i >= g.N - 2 * P;
  if (top) for (int e = 0; e < V; ++e) for (int c = 0; c < NB; ++c) Et[(i * V + e) * NB + c] = 0.0;
  if (bot) for (int e = 0; e < V; ++e) for (int c = 0; c < NB; ++c)
    Eb[((i - (g.N - 2 * P)) * V + e) * NB + c] = 0.0;
  for (int kk = 0; kk < NNZ; ++kk) {
    const int e = tf_j_eq(kk), var = tf_j_var(kk), off = tf_j_off(kk);
    const int j = map_node(i + off, g);
    if (j < nint) {
      rows[e * WB + BETA + (j - i) * V + var - e] += jv[kk];
    } else {
      const int c = (j - nint) * V + var;
      if (top) Et[(i * V + e) * NB + c] += jv[kk];
      else Eb[((i - (g.N - 2 * P)) * V + e) * NB + c] += jv[kk];
    }
  }
  for (int e = 0; e < V; ++e)
    for (int d = 0; d < WB; ++d) {
      const double s = __dmul_rn(a, rows[e * WB + d]);
      rows[e * WB + d] = (d == BETA) ? __dsub_rn(1.0, s) : -s;
    }
}

// Band rows of the thread's chunk (+ the next chunk's first BETA rows).
__device__ __forceinline__ void assemble_rows(double (&A)[C + BETA][WB], int i0, const Geom& g,
                                              const Buf& b, int sys, double a, const double* cst) {
  constexpr int NODES = M + EX;
  double win[NF][NODES + 2 * P];
  load_windows<NODES>(win, i0, g, b, sys, nullptr);
  const int npad = g.nblk * 32 * M;
#pragma unroll
  for (int m = 0; m < NODES; ++m) {
    const int i = i0 + m;
    double rows[V][WB];
#pragma unroll
    for (int e = 0; e < V; ++e)
#pragma unroll
      for (int d = 0; d < WB; ++d) rows[e][d] = 0.0;
    if (i < npad) {
      double jv[NNZ];
      if (i < g.N) {
        TfNodeIn in;
        node_inputs<NODES>(in, win, m, i, g, b, sys);
        tf_model_J(cst, in, jv);
      }
      if (i >= P && i < g.N - 2 * P) {
#pragma unroll
        for (int e = 0; e < V; ++e) rows[e][BETA] = 1.0;
#pragma unroll
        for (int kk = 0; kk < NNZ; ++kk) {
          const int e = tf_j_eq(kk), d = tf_j_off(kk) * V + tf_j_var(kk) - tf_j_eq(kk);
          const double s = __dmul_rn(a, jv[kk]);
          rows[e][BETA + d] = (d == 0) ? __dsub_rn(1.0, s) : -s;
        }
      } else {
        double tmp[V * WB];
        assemble_special(i, g, jv, a, tmp, b.btab + (long long)sys * 5 * NB * NB);
#pragma unroll
        for (int e = 0; e < V; ++e)
#pragma unroll
          for (int d = 0; d < WB; ++d) rows[e][d] = tmp[e * WB + d];
      }
    }
#pragma unroll
    for (int e = 0; e < V; ++e)
      if (m * V + e < C + BETA) {
#pragma unroll
        for (int d = 0; d < WB; ++d) A[(m * V + e < C + BETA) ? m * V + e : 0][d] = rows[e][d];
      }
  }
}

// ------------------------------------------------------------------ kernels
}  // namespace tfk

using namespace tfk;

// natural (sys, node, comp) -> chunk layout; padding nodes get `fill`
extern "C" __global__ void tf_k_pack(Geom g, const double* __restrict__ src, double* __restrict__ dst,
                                     int ncomp, int nsys, double fill) {
  const long long per = (long long)g.nblk * 32 * M * ncomp;
  const long long total = per * nsys;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int sys = (int)(t / per);
    const long long o = t - sys * per;          // chunk-layout offset
    const int lane = (int)(o & 31);
    const long long q = o >> 5;
    const int cc = M * ncomp;
    const int blk = (int)(q / cc), j = (int)(q % cc);
    const int i = (blk * 32 + lane) * M + j / ncomp, e = j % ncomp;
    dst[t] = (i < g.N) ? src[((long long)sys * g.N + i) * ncomp + e] : fill;
  }
}

extern "C" __global__ void tf_k_unpack(Geom g, const double* __restrict__ src, double* __restrict__ dst,
                                       int ncomp, int nsys) {
  const long long per = (long long)g.N * ncomp;
  const long long total = per * nsys;
  const long long sper = (long long)g.nblk * 32 * M * ncomp;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int sys = (int)(t / per);
    const long long o = t - sys * per;
    const int i = (int)(o / ncomp), e = (int)(o % ncomp);
    const int chunk = i / M, m = i % M;
    dst[t] = src[sys * sper + ((long long)(chunk >> 5) * M * ncomp + m * ncomp + e) * 32 + (chunk & 31)];
  }
}

// F(U) in natural layout (compatibility path of model.F), one thread per chunk
extern "C" __global__ void __launch_bounds__(NT) tf_k_eval_F(Geom g, Buf b, double* __restrict__ out) {
  const int chunks = g.nblk * 32;
  const long long t = blockIdx.x * (long long)NT + threadIdx.x;
  if (t >= (long long)chunks * g.batch) return;
  const int sys = (int)(t / chunks), chunk = (int)(t % chunks);
  const int i0 = chunk * M;
  if (i0 >= g.N) return;
  const double* cst = b.cst + (long long)sys * NC2;
  double win[NF][M + 2 * P];
  load_windows<M>(win, i0, g, b, sys, nullptr);
#pragma unroll
  for (int m = 0; m < M; ++m) {
    const int i = i0 + m;
    if (i < g.N) {
      TfNodeIn in;
      node_inputs<M>(in, win, m, i, g, b, sys);
      double f[V];
      tf_model_F(cst, in, f);
#pragma unroll
      for (int e = 0; e < V; ++e) out[((long long)sys * g.N + i) * V + e] = f[e];
    }
  }
}

// nonzero Jacobian values per node, natural layout [sys][node][nnz]
extern "C" __global__ void __launch_bounds__(NT) tf_k_eval_J(Geom g, Buf b, double* __restrict__ out) {
  const int chunks = g.nblk * 32;
  const long long t = blockIdx.x * (long long)NT + threadIdx.x;
  if (t >= (long long)chunks * g.batch) return;
  const int sys = (int)(t / chunks), chunk = (int)(t % chunks);
  const int i0 = chunk * M;
  if (i0 >= g.N) return;
  const double* cst = b.cst + (long long)sys * NC2;
  double win[NF][M + 2 * P];
  load_windows<M>(win, i0, g, b, sys, nullptr);
#pragma unroll
  for (int m = 0; m < M; ++m) {
    const int i = i0 + m;
    if (i < g.N) {
      TfNodeIn in;
      node_inputs<M>(in, win, m, i, g, b, sys);
      double jv[NNZ];
      tf_model_J(cst, in, jv);
#pragma unroll
      for (int k = 0; k < NNZ; ++k) out[((long long)sys * g.N + i) * NNZ + k] = jv[k];
    }
  }
}

// ---- factor: A = I - a*J(U) -> banded LU (chunk scan with linear-fractional maps)
extern "C" __global__ void __launch_bounds__(NT) tf_k_factor(Geom g, Buf b, double a) {
  __shared__ double smem[(WARPS + 1) * KMAX];
  int sys, tile;
  resolve_tile(g, b, sys, tile);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int blk = tile * WARPS + warp;
  const bool active = blk < g.nblk;
  const int chunk = blk * 32 + lane;
  const double* cst = b.cst + (long long)sys * NC2;
  double A[C + BETA][WB];
  Star mine = Star::identity();
  int bad = 0;
  if (active) {
    assemble_rows(A, chunk * M, g, b, sys, a, cst);
    tfb::ChunkLU<BETA, C>::run1(A, mine, bad);
  }
  const Star pre = tile_scan<Star>(mine, smem, g.tiles > 1, b, (long long)sys * g.tiles, tile,
                                   Star::identity(), nullptr);
  if (active) {
    double Uf[C][BETA + 1], Lown[C][BETA], Lnext[BETA][BETA];
    tfb::ChunkLU<BETA, C>::run2(A, pre.P(), Uf, Lown, Lnext, bad);
    double* Lg = b.Lf + sys * vstride(g) * BETA;
    double* Ug = b.Uf + sys * vstride(g) * (BETA + 1);
#pragma unroll
    for (int r = 0; r < C; ++r) {
#pragma unroll
      for (int q = 0; q <= BETA; ++q) Ug[((long long)blk * C + r) * 32 * (BETA + 1) + q * 32 + lane] = Uf[r][q];
#pragma unroll
      for (int q = 1; q <= BETA; ++q)
        if (q <= r) Lg[((long long)blk * C + r) * 32 * BETA + (q - 1) * 32 + lane] = Lown[r][q - 1];
        else if (chunk == 0) Lg[((long long)blk * C + r) * 32 * BETA + (q - 1) * 32 + lane] = 0.0;
    }
    const int nchunk = chunk + 1;
    if (nchunk < g.nblk * 32) {
      const int nb = nchunk >> 5, nl = nchunk & 31;
#pragma unroll
      for (int r = 0; r < BETA; ++r)
#pragma unroll
        for (int q = r + 1; q <= BETA; ++q)
          Lg[((long long)nb * C + r) * 32 * BETA + (q - 1) * 32 + nl] = Lnext[r][q - 1];
    }
    if (bad) atomicOr(b.status + sys, 1);
  }
}

// ---- forward substitution of one stage: rhs = dt*F(U_i) + sum cfac_j k_j ; L y = rhs
extern "C" __global__ void __launch_bounds__(NT) tf_k_fwd(Geom g, Buf b, Stage st) {
  __shared__ double smem[(WARPS + 1) * KMAX];
  int sys, tile;
  resolve_tile(g, b, sys, tile);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int blk = tile * WARPS + warp;
  const bool active = blk < g.nblk;
  const int chunk = blk * 32 + lane;
  const int i0 = chunk * M;
  const double* cst = b.cst + (long long)sys * NC2;
  const long long vs = vstride(g);
  double L[C][BETA], f[C], y[C];
  Aff mine = Aff::identity();
  if (active) {
    double win[NF][M + 2 * P];
    load_windows<M>(win, i0, g, b, sys, &st);
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const int i = i0 + m;
      double fe[V];
#pragma unroll
      for (int e = 0; e < V; ++e) fe[e] = 0.0;
      if (i < g.N) {
        TfNodeIn in;
        node_inputs<M>(in, win, m, i, g, b, sys);
        tf_model_F(cst, in, fe);
      }
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const int r = m * V + e;
        double rhs = st.dt * fe[e];
        const long long a = ((long long)blk * C + r) * 32 + lane;
#pragma unroll
        for (int q = 0; q < MAXS; ++q)
          if (q < st.nprev) rhs += st.cfac[q] * b.K[q][sys * vs + a];
        f[r] = (i < g.N) ? rhs : 0.0;
      }
    }
    const double* Lg = b.Lf + sys * vs * BETA;
#pragma unroll
    for (int r = 0; r < C; ++r)
#pragma unroll
      for (int q = 0; q < BETA; ++q) L[r][q] = Lg[((long long)blk * C + r) * 32 * BETA + q * 32 + lane];
    double s0[BETA];
#pragma unroll
    for (int t = 0; t < BETA; ++t) s0[t] = 0.0;
    tfb::fwd_chunk<BETA, C>(L, f, s0, y);
    tfb::fwd_map<BETA, C>(L, y, mine);
  }
  const Aff pre = tile_scan<Aff>(mine, smem, g.tiles > 1, b, (long long)sys * g.tiles, tile,
                                 Aff::identity(), nullptr);
  if (active) {
    tfb::fwd_chunk<BETA, C>(L, f, pre.c(), y);
    double* Y = b.Y + sys * vs;
#pragma unroll
    for (int r = 0; r < C; ++r) Y[((long long)blk * C + r) * 32 + lane] = y[r];
  }
}

// ---- border fill.  The last P nodes ("border", NB unknowns) are ordered last:
//   [ A^  E ] = [ L^   0 ] [ U^  W ]      W = L^-1 E   (fill column, NB per row)
//   [ F^T Ab]   [ G^T  I ] [ 0   S ]      G^T = F^T U^-1 (fill row),  S = Ab - G^T W
// E / F^T hold the periodic corner entries (top rows) and the natural coupling of
// the last interior rows (bottom rows), so periodic and non-periodic systems share
// one code path and the periodic corners cost no Woodbury pass.  W and G decay
// away from the top; the CTA walks tiles from the top until the carried state is
// exactly zero, then does the bottom rows.  One CTA per system.
__device__ __forceinline__ long long fidx(int R, int q, int width) {
  const int chunk = R / C, j = R % C;
  return (((long long)(chunk >> 5) * C + j) * width + q) * 32 + (chunk & 31);
}

extern "C" __global__ void __launch_bounds__(NT) tf_k_border_fill(Geom g, Buf b, double a) {
  __shared__ double smem[(WARPS + 1) * KMAX];
  __shared__ double s_red[WARPS][NB * NB];
  __shared__ int s_alive;
  const int sys = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long vs = vstride(g);
  const double* bt = b.btab + (long long)sys * 5 * NB * NB;
  const double* Lg = b.Lf + sys * vs * BETA;
  const double* Ug = b.Uf + sys * vs * (BETA + 1);
  double* Wg = b.Wb + sys * vs * NB;
  double* Gg = b.Gb + sys * vs * NB;
  const int ntile = (g.nblk + WARPS - 1) / WARPS;
  const int tile_rows = WARPS * 32 * C;
  const int bot0 = g.nhat - NB;
  const int tail0 = bot0 / tile_rows;
  double cw[NB][BETA], cg[NB][BETA];            // carried recurrence states
#pragma unroll
  for (int c = 0; c < NB; ++c)
#pragma unroll
    for (int t = 0; t < BETA; ++t) { cw[c][t] = 0.0; cg[c][t] = 0.0; }
  int lead_rows = -1;
  int tile = 0;
  while (tile < ntile) {
    const int blk = tile * WARPS + warp;
    const bool active = blk < g.nblk;
    const int chunk = blk * 32 + lane;
    const int r0 = chunk * C;
    double L[C][BETA], L2[C][BETA], inv[C];
    if (active) {
#pragma unroll
      for (int r = 0; r < C; ++r) {
#pragma unroll
        for (int q = 0; q < BETA; ++q) L[r][q] = Lg[((long long)blk * C + r) * 32 * BETA + q * 32 + lane];
        inv[r] = Ug[((long long)blk * C + r) * 32 * (BETA + 1) + lane];
#pragma unroll
        for (int q = 1; q <= BETA; ++q) {
          const int R = r0 + r - q;
          L2[r][q - 1] = (R >= 0) ? Ug[fidx(R, q, BETA + 1)] * Ug[fidx(R, 0, BETA + 1)] : 0.0;
        }
      }
    }
#pragma unroll 1
    for (int pass = 0; pass < 2 * NB; ++pass) {
      const bool isW = pass < NB;
      const int c = isW ? pass : pass - NB;
      double f[C], y[C];
      Aff mine = Aff::identity();
      if (active) {
#pragma unroll
        for (int r = 0; r < C; ++r) {
          const int gr = r0 + r;
          double v = 0.0;
          if (gr < NB) v = isW ? bt[0 * NB * NB + gr * NB + c] : bt[2 * NB * NB + c * NB + gr];
          else if (gr >= bot0 && gr < g.nhat)
            v = isW ? bt[1 * NB * NB + (gr - bot0) * NB + c] : bt[3 * NB * NB + c * NB + (gr - bot0)];
          f[r] = -(a * v);
        }
        double s0[BETA];
#pragma unroll
        for (int t = 0; t < BETA; ++t) s0[t] = 0.0;
        if (isW) { tfb::fwd_chunk<BETA, C>(L, f, s0, y); tfb::fwd_map<BETA, C>(L, y, mine); }
        else { tfb::fwd_chunk<BETA, C>(L2, f, s0, y); tfb::fwd_map<BETA, C>(L2, y, mine); }
      }
      Aff carry;
#pragma unroll
      for (int k = 0; k < BETA * BETA; ++k) carry.d[k] = 0.0;
#pragma unroll
      for (int t = 0; t < BETA; ++t) {
        double v = 0.0;
#pragma unroll
        for (int cc = 0; cc < NB; ++cc) if (cc == c) v = isW ? cw[cc][t] : cg[cc][t];
        carry.c()[t] = v;
      }
      Aff total;
      const Aff pre = tile_scan<Aff>(mine, smem, false, b, 0, 0, carry, &total);
#pragma unroll
      for (int t = 0; t < BETA; ++t)
#pragma unroll
        for (int cc = 0; cc < NB; ++cc) if (cc == c) { if (isW) cw[cc][t] = total.c()[t]; else cg[cc][t] = total.c()[t]; }
      if (active) {
        if (isW) tfb::fwd_chunk<BETA, C>(L, f, pre.c(), y);
        else tfb::fwd_chunk<BETA, C>(L2, f, pre.c(), y);
#pragma unroll
        for (int r = 0; r < C; ++r) {
          const long long o = ((long long)blk * C + r) * 32 * NB + c * 32 + lane;
          if (isW) Wg[o] = y[r]; else Gg[o] = y[r] * inv[r];
        }
      }
    }
    // continue while any carried state is non-zero
    bool alive = false;
#pragma unroll
    for (int c = 0; c < NB; ++c)
#pragma unroll
      for (int t = 0; t < BETA; ++t) alive = alive || (cw[c][t] != 0.0) || (cg[c][t] != 0.0);
    __syncthreads();
    if (threadIdx.x == 0) s_alive = alive ? 1 : 0;
    __syncthreads();
    alive = s_alive != 0;
    int rows_done = (tile + 1) * tile_rows;
    if (rows_done > g.nhat) rows_done = g.nhat;
    if (lead_rows < 0) {
      if (!alive || tile + 1 >= ntile) { lead_rows = rows_done; }
    }
    if (lead_rows >= 0 && tile + 1 < tail0 && !alive) tile = tail0; else tile = tile + 1;
  }
  if (lead_rows < 0) lead_rows = g.nhat;
  __syncthreads();
  // S = (I - a Ab) - sum_r G[r][.]^T W[r][.] over rows where both may be non-zero
  double part[NB][NB];
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) part[i][j] = 0.0;
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const int lo = pass == 0 ? 0 : (lead_rows > bot0 ? lead_rows : bot0);
    const int hi = pass == 0 ? lead_rows : g.nhat;
    for (int r = lo + threadIdx.x; r < hi; r += NT) {
      double gv[NB], wv[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) { gv[c] = Gg[fidx(r, c, NB)]; wv[c] = Wg[fidx(r, c, NB)]; }
#pragma unroll
      for (int i = 0; i < NB; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) part[i][j] += gv[i] * wv[j];
    }
  }
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      double v = part[i][j];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      if (lane == 0) s_red[warp][i * NB + j] = v;
    }
  __syncthreads();
  if (threadIdx.x == 0) {
    double S[NB * NB], I[NB * NB];
    for (int i = 0; i < NB; ++i)
      for (int j = 0; j < NB; ++j) {
        double v = 0.0;
        for (int w = 0; w < WARPS; ++w) v += s_red[w][i * NB + j];
        S[i * NB + j] = ((i == j) ? 1.0 : 0.0) - a * bt[4 * NB * NB + i * NB + j] - v;
        I[i * NB + j] = (i == j) ? 1.0 : 0.0;
      }
    tfb::solve_inplace<NB, NB>(S, I);
    bool bad = false;
    for (int k = 0; k < NB * NB; ++k) {
      b.Sinv[(long long)sys * NB * NB + k] = I[k];
      if (!(fabs(I[k]) < 1e300)) bad = true;
    }
    if (bad) atomicOr(b.status + sys, 2);
    b.lead[sys * 2 + 0] = lead_rows;
    b.lead[sys * 2 + 1] = lead_rows;
  }
}

// ---- border solve: x_b = S^-1 (y_b - G^T y)   (one warp per system)
extern "C" __global__ void tf_k_border_solve(Geom g, Buf b) {
  const int sys = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (sys >= g.batch) return;
  const long long vs = vstride(g);
  const double* Y = b.Y + sys * vs;
  const double* G = b.Gb + sys * vs * NB;
  const int glead = b.lead[sys * 2 + 1];
  double acc[NB];
#pragma unroll
  for (int c = 0; c < NB; ++c) acc[c] = 0.0;
  const int bot0 = g.nhat - NB;
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const int lo = pass == 0 ? 0 : (glead > bot0 ? glead : bot0);
    const int hi = pass == 0 ? (glead < g.nhat ? glead : g.nhat) : g.nhat;
    for (int r = lo + lane; r < hi; r += 32) {
      const long long a = ridx(r);
      const double yr = Y[a];
#pragma unroll
      for (int c = 0; c < NB; ++c) acc[c] += G[(a >> 5) * 32 * NB + c * 32 + (a & 31)] * yr;
    }
  }
#pragma unroll
  for (int c = 0; c < NB; ++c)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], d);
  if (lane == 0) {
    double yb[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) yb[c] = Y[ridx(g.nhat + c)] - acc[c];
    const double* Si = b.Sinv + (long long)sys * NB * NB;
#pragma unroll
    for (int r = 0; r < NB; ++r) {
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < NB; ++c) s += Si[r * NB + c] * yb[c];
      b.xb[(long long)sys * NB + r] = s;
    }
  }
}

// ---- backward substitution: U x = y - W x_b ; k_i = x - sum cfac_j k_j
extern "C" __global__ void __launch_bounds__(NT) tf_k_bwd(Geom g, Buf b, Stage st) {
  __shared__ double smem[(WARPS + 1) * KMAX];
  int sys, tile;
  resolve_tile(g, b, sys, tile);
  const int lane_l = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int blk_l = tile * WARPS + warp;           // logical (reversed) block
  const bool active = blk_l < g.nblk;
  const int blk = g.nblk - 1 - blk_l;
  const int lane = 31 - lane_l;
  const int chunk = blk * 32 + lane;
  const long long vs = vstride(g);
  double Uf[C][BETA + 1], y[C], x[C];
  Aff mine = Aff::identity();
  if (active) {
    const double* Ug = b.Uf + sys * vs * (BETA + 1);
    const double* Y = b.Y + sys * vs;
#pragma unroll
    for (int r = 0; r < C; ++r) {
#pragma unroll
      for (int q = 0; q <= BETA; ++q) Uf[r][q] = Ug[((long long)blk * C + r) * 32 * (BETA + 1) + q * 32 + lane];
      y[r] = Y[((long long)blk * C + r) * 32 + lane];
    }
    // border coupling
    const int r0 = chunk * C;
    const int wlead = b.lead[sys * 2 + 0];
    if (r0 < wlead || (r0 + C > g.nhat - NB && r0 < g.nhat + NB)) {
      const double* W = b.Wb + sys * vs * NB;
      double xb[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) xb[c] = b.xb[(long long)sys * NB + c];
#pragma unroll
      for (int r = 0; r < C; ++r) {
        const int gr = r0 + r;
        if (gr < wlead || (gr >= g.nhat - NB && gr < g.nhat)) {
          double s = y[r];
#pragma unroll
          for (int c = 0; c < NB; ++c) s -= W[((long long)blk * C + r) * 32 * NB + c * 32 + lane] * xb[c];
          y[r] = s;
        } else if (gr >= g.nhat && gr < g.nhat + NB) {
#pragma unroll
          for (int c = 0; c < NB; ++c) if (gr - g.nhat == c) y[r] = xb[c];
        }
      }
    }
    double s0[BETA];
#pragma unroll
    for (int t = 0; t < BETA; ++t) s0[t] = 0.0;
    tfb::bwd_chunk<BETA, C>(Uf, y, s0, x);
    tfb::bwd_map<BETA, C>(Uf, x, mine);
  }
  const Aff pre = tile_scan<Aff>(mine, smem, g.tiles > 1, b, (long long)sys * g.tiles, tile,
                                 Aff::identity(), nullptr);
  if (active) {
    tfb::bwd_chunk<BETA, C>(Uf, y, pre.c(), x);
    double* Kout = b.K[st.istage] + sys * vs;
#pragma unroll
    for (int r = 0; r < C; ++r) {
      const long long a = ((long long)blk * C + r) * 32 + lane;
      double k = x[r];
#pragma unroll
      for (int q = 0; q < MAXS; ++q)
        if (q < st.nprev) k -= st.cfac[q] * b.K[q][sys * vs + a];
      Kout[a] = k;
    }
  }
}

// ---- U_new = U + sum b_i k_i ; err = || U_new - (U_new + sum bp_i k_i) ||_inf ; Dirichlet hook
extern "C" __global__ void __launch_bounds__(256) tf_k_update(Geom g, Buf b, Update up) {
  const long long vs = vstride(g);
  const int sys = blockIdx.y;
  const double* U = b.U + sys * vs;
  double* Un = b.Un + sys * vs;
  double emax = 0.0;
  for (long long a = blockIdx.x * (long long)blockDim.x + threadIdx.x; a < vs;
       a += (long long)gridDim.x * blockDim.x) {
    double acc = 0.0, accp = 0.0;
#pragma unroll
    for (int q = 0; q < MAXS; ++q)
      if (q < up.s) {
        const double k = b.K[q][sys * vs + a];
        const double t = __dmul_rn(up.b[q], k);
        acc = (q == 0) ? t : __dadd_rn(acc, t);
        if (up.has_pred) {
          const double tp = __dmul_rn(up.bp[q], k);
          accp = (q == 0) ? tp : __dadd_rn(accp, tp);
        }
      }
    double un = __dadd_rn(U[a], acc);
    if (up.has_pred) {
      const double e = fabs(__dsub_rn(un, __dadd_rn(un, accp)));
      // padding unknowns carry zeros: no contribution
      emax = (e > emax || e != e) ? e : emax;
    }
    Un[a] = un;
  }
  if (up.has_pred) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const double o = __shfl_xor_sync(0xffffffffu, emax, d);
      emax = (o > emax || o != o) ? o : emax;
    }
    if ((threadIdx.x & 31) == 0)
      atomicMax((unsigned long long*)(b.err + sys), (unsigned long long)__double_as_longlong(emax));
  }
}

// Dirichlet-style hook: U[var][0] = left, U[var][N-1] = right  (README.md:126-129)
extern "C" __global__ void tf_k_dirichlet(Geom g, double* __restrict__ U, const double* __restrict__ dir,
                                          int mask) {
  const int sys = blockIdx.x * blockDim.x + threadIdx.x;
  if (sys >= g.batch) return;
  double* u = U + sys * vstride(g);
#pragma unroll
  for (int e = 0; e < V; ++e) {
    if (mask & (1 << (2 * e))) u[vidx(0, e)] = dir[2 * e];
    if (mask & (1 << (2 * e + 1))) u[vidx(g.N - 1, e)] = dir[2 * e + 1];
  }
}
