// Model-specialised kernels of the implicit method-of-lines hot path (sm_100a).
//
// This file is compiled once per model: the generated model header (TF_NVAR,
// TF_P, ..., tf_model_F, tf_model_J; see triflow_b200/codegen.py) is included
// first, then this file, with -DTF_M=<nodes per thread> (warps per CTA are a launch-time choice).
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
#ifndef TF_FAST_DIV
#define TF_FAST_DIV 0
#endif
#ifndef TF_LB_FIRST
#define TF_LB_FIRST 4 /* look-back: predecessors awaited before the first evaluation */
#endif
#define TF_MAXW 16 /* max warps per CTA (runtime: blockDim.x / 32) */
#ifndef TF_MINB
/* min CTAs of 512 threads per SM for the sweep kernels: register cap 64 for narrow bands,
   128 for wide bands (their recurrence state alone is 60 doubles) */
#define TF_MINB (((TF_P * TF_NVAR + TF_NVAR - 1) <= 2) ? 2 : 1)
#endif

// Optional per-CTA phase time stamps (build the cubin with -DTF_TRACE; tools/trace_tiles.py).
// Slot = flag epoch of the chained launch (mod 16), 8 stamps (globaltimer ns) per CTA.
#ifdef TF_TRACE
__device__ unsigned long long tf_trace[16 * 4096 * 8];
__device__ unsigned long long tf_trace2[16 * 4096 * 8];   // look-back rounds: [0..3] start ns, [4..7] ready run
__device__ __forceinline__ void tf_stamp(int slot, int cta, int ph) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  tf_trace[(((slot) & 15) * 4096 + ((cta) & 4095)) * 8 + ph] = t;
}
#define TF_STAMP(slot, cta, ph) do { if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) == 0) tf_stamp(slot, cta, ph); } while (0)
#define TF_STAMP_LANE0(slot, cta, ph) do { if ((threadIdx.x & 31) == 0) tf_stamp(slot, cta, ph); } while (0)
#else
#define TF_STAMP(slot, cta, ph) do { } while (0)
#define TF_STAMP_LANE0(slot, cta, ph) do { } while (0)
#endif

namespace tfk {
typedef TfGeom Geom; typedef TfBuf Buf; typedef TfStage Stage;

constexpr int V = TF_NVAR;
constexpr int P = TF_P;
constexpr int BETA = (P * V + V - 1) > 0 ? (P * V + V - 1) : 1;
constexpr int NB = (P * V) > 0 ? (P * V) : 1;     // border unknowns (last P nodes)
constexpr int M = TF_M;
constexpr int C = M * V;
constexpr int MAXW = TF_MAXW;
constexpr int NT = 32 * MAXW;         // launch bound of the tile kernels
constexpr bool FD = TF_FAST_DIV != 0;  // fast division by uniform constants in the solver path
// the factor kernel keeps whole band rows in registers: fewer threads per CTA for wide bands
#ifndef TF_FACTOR_NT
#define TF_FACTOR_NT (((TF_P * TF_NVAR + TF_NVAR - 1) <= 1) ? 512 : 256)
#endif
#ifndef TF_FACTOR_MINB
#define TF_FACTOR_MINB 2
#endif
constexpr int NT_FACTOR = TF_FACTOR_NT;
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
// out-of-line combine for wide maps (one copy of the code, one stack frame)
template <class Mon>
__device__ __noinline__ Mon combine_call(const Mon& a, const Mon& b) {
  return Mon::combine(a, b);
}

// inclusive scan over the lanes of a warp, lower lane = earlier
template <class Mon>
__device__ __forceinline__ Mon warp_scan(Mon v, int lane) {
  if (Mon::K > 1000) {                     // (rolled variant for very wide maps: measured slower at K = 30, 100)
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
      const Mon o = shfl_up(v, d);
      const Mon c = combine_call(o, v);
      v = select(lane >= d, c, v);
    }
    return v;
  }
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
  if (Mon::K > 1000) {
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
      const Mon o = shfl_down(v, d);
      const Mon c = combine_call(o, v);
      v = select(lane + d < 32, c, v);
    }
    return v;
  }
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const Mon o = shfl_down(v, d);
    const Mon c = Mon::combine(o, v);
    v = select(lane + d < 32, c, v);
  }
  return v;
}

// Flags are published with a release store and polled with RELAXED loads followed by one
// acquire fence once the poll succeeded.  (ld.acquire.gpu compiles to LDG.STRONG.GPU +
// CCTL.IVALL: every poll of a spin loop would invalidate the whole L1 of the SM, under the
// feet of the co-resident CTAs.)  The payload itself is read with L1-bypassing loads.
__device__ __forceinline__ int ld_flag(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void acquire_fence() {
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
__device__ __forceinline__ void st_flag(int* p, int v) {
  asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// "absorbing" maps: composing anything earlier in front of them changes nothing
// that is used downstream (exact, not approximate: the propagator is exactly 0).
__device__ __forceinline__ bool absorbing(const Aff& m) {
  bool z = true;
#pragma unroll
  for (int k = 0; k < BETA * BETA; ++k) z = z && (m.d[k] == 0.0);
  return z;
}
__device__ __forceinline__ bool absorbing(const Star& m) {
  bool zq = true, zs = true;
#pragma unroll
  for (int k = 0; k < BETA * BETA; ++k) {
    zq = zq && (m.Q()[k] == 0.0);
    zs = zs && (m.S()[k] == 0.0);
  }
  return zq || zs;
}

// Look-back records: every value travels with its own tag in one 16-byte word (single
// transaction, as in CUB's packed tile descriptors), so neither the publisher nor the
// reader needs a memory fence: a record is valid once all its K words carry the tag.
// (A release store costs a MEMBAR.ALL.GPU, an acquire load a CCTL.IVALL; measured with
// tools/trace_tiles.py these fences made the look-back the longest phase of every sweep.)
struct __align__(16) LbWord { double v; long long tag; };
// Published with a 128-bit atomic exchange: atomics are performed at the L2, whereas plain
// stores were observed (tools/trace_tiles.py) to become visible to the polling tiles up to
// ~10 us late while their SM was busy.
__device__ __forceinline__ void st_word(LbWord* p, double v, long long tag) {
  long long o0, o1;
  asm volatile(
      "{\n.reg .b128 q, r;\nmov.b128 q, {%3, %4};\n"
      "atom.relaxed.gpu.global.exch.b128 r, [%2], q;\nmov.b128 {%0, %1}, r;\n}"
      : "=l"(o0), "=l"(o1)
      : "l"(p), "l"(__double_as_longlong(v)), "l"(tag)
      : "memory");
  (void)o0; (void)o1;
}
__device__ __forceinline__ void ld_word(const LbWord* p, double& v, long long& t) {
  long long a;
  asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(t) : "l"(p) : "memory");
  v = __longlong_as_double(a);
}
template <class Mon>
__device__ __forceinline__ void lb_publish(LbWord* dst, const Mon& m, long long tag, int lane) {
#pragma unroll
  for (int k0 = 0; k0 < Mon::K; k0 += 32) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 32; ++k)
      if (k0 + k < Mon::K && lane == k) v = m.d[(k0 + k < Mon::K) ? k0 + k : 0];
    if (k0 + lane < Mon::K) st_word(dst + k0 + lane, v, tag);
  }
}
// Both records of a tile (inclusive prefix and aggregate) in ONE memory round trip: all 2K
// loads are issued before the first tag is looked at (a compare right behind each load
// serialises them: K round trips of ~1 us each per record, seen in the tile trace).
template <class Mon>
__device__ __forceinline__ void lb_read2(const LbWord* pi, const LbWord* pa, Mon& mi, Mon& ma,
                                         long long FI, long long FA, bool& isI, bool& isA) {
  long long ti[Mon::K], ta[Mon::K];
#pragma unroll
  for (int k = 0; k < Mon::K; ++k) ld_word(pi + k, mi.d[k], ti[k]);
#pragma unroll
  for (int k = 0; k < Mon::K; ++k) ld_word(pa + k, ma.d[k], ta[k]);
  isI = true;
  isA = true;
#pragma unroll
  for (int k = 0; k < Mon::K; ++k) {
    isI = isI && (ti[k] == FI);
    isA = isA && (ta[k] == FA);
  }
}

// Decoupled look-back over the tiles of one system.  Called by all 32 lanes of
// warp 0 with the tile aggregate (the same value in every lane); returns the exclusive
// prefix of the tile.  Records carry the launch epoch (no reset between launches):
// tag = epoch*4 + {1: aggregate, 2: inclusive prefix}.  The walk stops at the nearest
// inclusive prefix, at the start of the system, or as soon as the accumulated
// aggregate is absorbing (its propagator underflowed to exactly zero), which for
// well-conditioned systems happens after a few tiles.
template <class Mon>
__device__ Mon lookback(const Mon& aggregate, const Buf& b, long long gbase, int tile, int lane,
                        int epoch) {
  LbWord* agg = (LbWord*)b.lbagg + (gbase + tile) * KMAX;
  LbWord* inc = (LbWord*)b.lbinc + (gbase + tile) * KMAX;
  const long long FA = epoch * 4LL + 1, FI = epoch * 4LL + 2;
  if (tile == 0) {
    lb_publish(inc, aggregate, FI, lane);
    return Mon::identity();
  }
  lb_publish(agg, aggregate, FA, lane);
  Mon prefix = Mon::identity();
  int look = tile - 1;
  bool finished = false;
#ifdef TF_TRACE
  int tr_rounds = 0, tr_depth = 0;
#endif
  while (!finished) {
    const int t = look - lane;            // lane 0 = nearest predecessor
    // Wait for the TF_LB_FIRST nearest predecessors (a few tiles usually make the
    // aggregate absorbing) and evaluate the contiguous run that is ready; if that is not
    // conclusive, wait for the whole window of 32 and evaluate again.  Every round costs
    // two dependent memory round trips that queue behind the bulk loads of the sweep.
    int need = TF_LB_FIRST;               // lanes below `need` wait for their tile
    Mon w;
    // Warp-uniform control flow: the whole warp repeats the round until the lanes below `need`
    // have their record, every decision is a vote.  (A per-lane polling loop in front of the
    // shuffles made the compiler emit them for a divergent warp: WARPSYNC.COLLECTIVE sequences
    // in which warp 0 spent most of the look-back, ncu source view.)
    Mon e = Mon::identity();
    bool isI = true, ready = (t < 0), first = true;
    while (true) {
      if (t >= 0 && (first || (lane < need && !ready))) {   // both records in one round trip
        Mon ea;
        bool isA;
        const LbWord* pi = (const LbWord*)b.lbinc + (gbase + t) * KMAX;
        const LbWord* pa = (const LbWord*)b.lbagg + (gbase + t) * KMAX;
        lb_read2(pi, pa, e, ea, FI, FA, isI, isA);
        if (!isI) e = ea;
        ready = isI || isA;
      }
      first = false;
      if (!__all_sync(0xffffffffu, ready || lane >= need)) continue;
      const unsigned mr = __ballot_sync(0xffffffffu, ready);
      const unsigned m2 = __ballot_sync(0xffffffffu, ready && isI);
      const int kr = (~mr == 0u) ? 32 : (__ffs(~mr) - 1);     // contiguous ready lanes
      const int kstop = __ffs(m2) - 1;                         // nearest inclusive (or start)
      const bool hit = (kstop >= 0 && kstop < kr);
      const int last = hit ? kstop : kr - 1;
      const Mon el = select(t >= 0 && lane <= last, e, Mon::identity());
      w = warp_reduce_rev(el, lane);
      w = shfl_idx(w, 0);
#ifdef TF_TRACE
      if (lane == 0 && tr_rounds < 4) {
        unsigned long long tn;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tn));
        tf_trace2[(((epoch) & 15) * 4096 + ((tile) & 4095)) * 8 + tr_rounds] = tn;
        tf_trace2[(((epoch) & 15) * 4096 + ((tile) & 4095)) * 8 + 4 + tr_rounds] = kr * 100 + (hit ? 50 : 0) + (absorbing(w) ? 1 : 0);
      }
      tr_rounds += 1;
      if (hit || absorbing(w) || kr == 32) tr_depth += last + 1;
#endif
      if (__all_sync(0xffffffffu, hit || absorbing(w))) { finished = true; break; }
      if (kr == 32) break;                // full window of aggregates: go further back
      need = 32;                          // not conclusive: wait for the whole window
    }
    prefix = Mon::combine(w, prefix);
    look -= 32;
  }
#ifdef TF_TRACE
  if (lane == 0) tf_trace[(((epoch) & 15) * 4096 + ((tile) & 4095)) * 8 + 2] = tr_rounds * 1000 + tr_depth;
#endif
  lb_publish(inc, Mon::combine(prefix, aggregate), FI, lane);
  return prefix;
}

// Exclusive prefix of every thread's element over (lane, warp, tile) order.
// `carry` is the prefix of the tile when look-back is disabled (single tile, or the
// border kernel's sequential tile loop).  `reuse`: the shared scratch was used by an
// earlier scan of the same kernel (needs a barrier before it is overwritten).
//
// Single tile without a tile total: after ONE barrier every warp scans the warp totals
// itself (redundantly), so no warp waits for warp 0.  With look-back, warp 0 resolves
// the tile prefix and the others wait for it.
template <class Mon>
__device__ Mon tile_scan(const Mon& mine, double* smem /* (MAXW+1)*K doubles */, bool use_lookback,
                         const Buf& b, long long gbase, int tile, int epoch, const Mon& carry,
                         Mon* tile_total, bool reuse = false) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const Mon incl = warp_scan(mine, lane);
  Mon excl = shfl_up(incl, 1);
  excl = select(lane == 0, Mon::identity(), excl);
  if (reuse) __syncthreads();
  if (lane == 31) {
#pragma unroll
    for (int k = 0; k < Mon::K; ++k) smem[warp * Mon::K + k] = incl.d[k];
  }
  __syncthreads();
  if (!use_lookback && tile_total == nullptr) {
    Mon w = Mon::identity();
    if (lane < nwarps) {
#pragma unroll
      for (int k = 0; k < Mon::K; ++k) w.d[k] = smem[lane * Mon::K + k];
    }
    const Mon wi = warp_scan(w, lane);
    Mon we = shfl_idx(wi, warp > 0 ? warp - 1 : 0);
    we = select(warp == 0, Mon::identity(), we);
    return Mon::combine(Mon::combine(carry, we), excl);
  }
  if (warp == 0) {
    Mon w = Mon::identity();
    if (lane < nwarps) {
#pragma unroll
      for (int k = 0; k < Mon::K; ++k) w.d[k] = smem[lane * Mon::K + k];
    }
    const Mon wi = warp_scan(w, lane);
    Mon we = shfl_up(wi, 1);
    we = select(lane == 0, Mon::identity(), we);
    const Mon total = shfl_idx(wi, nwarps - 1);
    Mon tp = carry;
    TF_STAMP_LANE0(epoch, tile, 4);
    if (use_lookback) tp = lookback(total, b, gbase, tile, lane, epoch);
    TF_STAMP_LANE0(epoch, tile, 5);
    const Mon wp = Mon::combine(tp, we);
    __syncwarp();
    if (lane < nwarps) {
#pragma unroll
      for (int k = 0; k < Mon::K; ++k) smem[lane * Mon::K + k] = wp.d[k];
    }
    if (tile_total != nullptr && lane == 0) {
      const Mon tt = Mon::combine(tp, total);
#pragma unroll
      for (int k = 0; k < Mon::K; ++k) smem[MAXW * Mon::K + k] = tt.d[k];
    }
  }
  __syncthreads();
  Mon wp;
#pragma unroll
  for (int k = 0; k < Mon::K; ++k) wp.d[k] = smem[warp * Mon::K + k];
  if (tile_total != nullptr) {
#pragma unroll
    for (int k = 0; k < Mon::K; ++k) tile_total->d[k] = smem[MAXW * Mon::K + k];
  }
  return Mon::combine(wp, excl);
}

// Tile / system assignment.  One tile per system: blockIdx.  Otherwise a ticket
// (monotone counter), so that every tile a CTA may wait for is already running
// (look-back forward progress).  The flag epoch and the ticket base of the launch live
// in device memory (b.ctl) and are advanced by the last CTA of every chained kernel
// (finish_chain), so kernel parameters do not change from step to step and a whole
// step can be replayed from a CUDA graph.
__device__ __forceinline__ void resolve_tile(const Geom& g, const Buf& b, int& sys, int& tile,
                                             int& epoch) {
  if (g.tiles == 1) {
    sys = blockIdx.x;
    tile = 0;
    epoch = 0;
    return;
  }
  __shared__ unsigned s_ticket;
  __shared__ int s_epoch;
  if (threadIdx.x == 0) {
    s_epoch = ld_flag(b.ctl + 0);
    const unsigned base = (unsigned)ld_flag(b.ctl + 1);
    s_ticket = atomicAdd((unsigned*)b.flags, 1u) - base;
  }
  __syncthreads();
  const int t = (int)s_ticket;
  epoch = s_epoch;
  sys = t / g.tiles;
  tile = t - sys * g.tiles;
}

// Last CTA of a chained launch: new epoch and ticket base for the next launch.
__device__ __forceinline__ void finish_chain(const Geom& g, const Buf& b, int epoch) {
  if (g.tiles == 1) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x;
    const unsigned d = atomicAdd((unsigned*)(b.ctl + 2), 1u);
    if (d == total - 1) {
      b.ctl[2] = 0;
      b.ctl[0] = (epoch >= (1 << 28)) ? 1 : epoch + 1;
      *((unsigned*)(b.ctl + 1)) += total;
    }
  }
}

// ------------------------------------------------------------ stencil windows
// Values of every field at nodes i0-P .. i0+NODES-1+P for the stage state
// U + sum_j alpha_j K_j (dependent variables) and the helper planes.
// stage state of one unknown: U + ((alpha_0 k_0 + alpha_1 k_1) + ...) in the reference's
// summation order (schemes.py:153-155).  NPREV < 0: number of terms known at run time.
template <int NPREV>
__device__ __forceinline__ double stage_value(const double* __restrict__ U, const Buf& b,
                                              long long sysoff, long long a, const Stage* st) {
  double u = U[a];
  constexpr int NQ = NPREV < 0 ? MAXS : NPREV;
  if (NQ > 0 && st != nullptr) {
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < NQ; ++q)
      if (NPREV >= 0 || q < st->nprev) {
        const double term = __dmul_rn(st->alpha[q], b.K[q][sysoff + a]);
        acc = (q == 0) ? term : __dadd_rn(acc, term);
      }
    if (NPREV >= 0 || st->nprev > 0) u = __dadd_rn(u, acc);
  }
  return u;
}

template <int NODES, int NPREV>
__device__ __forceinline__ void load_windows(double (&win)[NF][NODES + 2 * P], int i0, const Geom& g,
                                             const Buf& b, int sys, const Stage* st) {
  const long long so = sys * vstride(g);
  const double* U = b.U + so;
  const int chunk = i0 / M;
  if (i0 >= P && i0 + NODES + P <= g.N && NODES + P <= 2 * M && P <= M) {
    // interior: every window node is a real node of this, the previous or the next
    // chunk(s): addresses are lane-neighbours of the own chunk, offsets are constants
#pragma unroll
    for (int w = 0; w < NODES + 2 * P; ++w) {
      const int rel = w - P;                               // node offset from i0
      const int dc = rel < 0 ? -1 : (rel >= M ? 1 : 0);    // neighbour chunk (compile time)
      const int m = rel - dc * M;
      const int ch = chunk + dc;
      const long long cb = ((long long)(ch >> 5) * C) * 32 + (ch & 31);
      const long long hb = ((long long)(ch >> 5) * M) * 32 + (ch & 31);
#pragma unroll
      for (int e = 0; e < V; ++e)
        win[e][w] = stage_value<NPREV>(U, b, so, cb + (long long)(m * V + e) * 32, st);
#pragma unroll
      for (int h = 0; h < NH; ++h)
        win[V + h][w] = b.H[(sys * (long long)NH + h) * hstride(g) + hb + (long long)m * 32];
    }
    return;
  }
#pragma unroll
  for (int w = 0; w < NODES + 2 * P; ++w) {
    const int j = map_node(i0 - P + w, g);
#pragma unroll
    for (int e = 0; e < V; ++e) win[e][w] = stage_value<NPREV>(U, b, so, vidx(j, e), st);
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

// F as the SOLVER kernels evaluate it.  For a model that is homogeneous linear in the stencil
// values with uniform coefficients (TF_F_LINEAR, decided symbolically by the code generator:
// F == sum_k J_k u_k identically; advection-diffusion, heat) it is that sum -- NNZ fused
// multiply-adds with Jacobian constants that are table look-ups -- instead of the reference's
// expanded expression (advection-diffusion: 3 instead of 32 fp64 instructions per node and
// stage).  Same value up to rounding (the terms that cancel are the same); `model.F`
// (tf_k_eval_F) and the exact-division build keep the reference's operation order.
#ifndef TF_F_LINEAR
#define TF_F_LINEAR 0
#endif
#ifndef TF_F_SPLIT
#define TF_F_SPLIT 0
#endif
template <bool FDIV>
__device__ __forceinline__ void tf_model_F_solver(const double* __restrict__ cst, const TfNodeIn& in,
                                                  double (&out)[V]) {
#if TF_F_LINEAR && TF_FAST_DIV
  double jv[NNZ];
  tf_model_J<FDIV>(cst, in, jv);
  bool first[V];
#pragma unroll
  for (int e = 0; e < V; ++e) { out[e] = 0.0; first[e] = true; }
#pragma unroll
  for (int kk = 0; kk < NNZ; ++kk) {
    const int e = tf_j_eq(kk);
    const double u = in.w[tf_j_var(kk)][P + tf_j_off(kk)];
    out[e] = first[e] ? __dmul_rn(jv[kk], u) : __fma_rn(jv[kk], u, out[e]);
    first[e] = false;
  }
#elif TF_F_SPLIT && TF_FAST_DIV && !defined(TF_NO_F_SPLIT)
  // nonlinear polynomial-like models: the generator's monomial-collected form (one host-evaluated
  // coefficient per monomial of the stencil values, no division on the device)
  tf_model_Fs<FDIV>(cst, in, out);
#else
  tf_model_F<FDIV>(cst, in, out);
#endif
}

// -------------------------------------------------------------- J -> A rows
// Rows of A = I - a*J for a node that touches the domain ends, the border (last
// P nodes) or the padding.  Dynamic indexing on purpose: rare path.
__device__ __noinline__ void assemble_special(int i, const Geom& g, const double* jv, double a,
                                              double* rows /* [V][WB] */, double* btab) {
  // btab == nullptr: the node is evaluated a second time as one of the extra rows of the
  // previous chunk; only its owner writes the border tables (two writers would race on the
  // zero-then-accumulate below when they run in different warps)
  for (int k = 0; k < V * WB; ++k) rows[k] = 0.0;
  const int nint = g.N - P;                 // interior nodes
  // The border tables live in global memory: every entry is accumulated in a local row and
  // stored once (no read-modify-write round trips on the critical path of the end tiles).
  if (i >= nint) {                          // border or padding: identity row in the band
    for (int e = 0; e < V; ++e) rows[e * WB + BETA] = 1.0;
    if (i >= g.N || btab == nullptr) return;
    double* Ft = btab + 2 * NB * NB;
    double* Fb = btab + 3 * NB * NB;
    double* Ab = btab + 4 * NB * NB;
    for (int e = 0; e < V; ++e) {
      double ft[NB], fb[NB], ab[NB];
      for (int c = 0; c < NB; ++c) { ft[c] = 0.0; fb[c] = 0.0; ab[c] = 0.0; }
      for (int kk = 0; kk < NNZ; ++kk) {
        if (tf_j_eq(kk) != e) continue;
        const int var = tf_j_var(kk), off = tf_j_off(kk);
        const int j = map_node(i + off, g);
        if (j >= nint) ab[(j - nint) * V + var] += jv[kk];
        else if (j < P) ft[j * V + var] += jv[kk];
        else fb[(j - (g.N - 2 * P)) * V + var] += jv[kk];
      }
      const int r = (i - nint) * V + e;
      for (int c = 0; c < NB; ++c) { Ft[r * NB + c] = ft[c]; Fb[r * NB + c] = fb[c]; Ab[r * NB + c] = ab[c]; }
    }
    return;
  }
  const bool top = i < P, bot = i >= g.N - 2 * P;
  const bool wr = btab != nullptr;
  double eacc[V][NB];
  for (int e = 0; e < V; ++e) for (int c = 0; c < NB; ++c) eacc[e][c] = 0.0;
  for (int kk = 0; kk < NNZ; ++kk) {
    const int e = tf_j_eq(kk), var = tf_j_var(kk), off = tf_j_off(kk);
    const int j = map_node(i + off, g);
    if (j < nint) {
      rows[e * WB + BETA + (j - i) * V + var - e] += jv[kk];
    } else {
      eacc[e][(j - nint) * V + var] += jv[kk];
    }
  }
  if (wr && (top || bot)) {
    double* Et = btab;
    double* Eb = btab + NB * NB;
    for (int e = 0; e < V; ++e)
      for (int c = 0; c < NB; ++c) {
        if (top) Et[(i * V + e) * NB + c] = eacc[e][c];
        if (bot) Eb[((i - (g.N - 2 * P)) * V + e) * NB + c] = eacc[e][c];
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
  load_windows<NODES, 0>(win, i0, g, b, sys, nullptr);
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
        tf_model_J<FD>(cst, in, jv);
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
        assemble_special(i, g, jv, a, tmp, m < M ? b.btab + (long long)sys * 5 * NB * NB : nullptr);
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
extern "C" __global__ void __launch_bounds__(256) tf_k_eval_F(Geom g, Buf b, double* __restrict__ out) {
  const int chunks = g.nblk * 32;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)chunks * g.batch) return;
  const int sys = (int)(t / chunks), chunk = (int)(t % chunks);
  const int i0 = chunk * M;
  if (i0 >= g.N) return;
  const double* cst = b.cst + (long long)sys * NC2;
  double win[NF][M + 2 * P];
  load_windows<M, 0>(win, i0, g, b, sys, nullptr);
#pragma unroll
  for (int m = 0; m < M; ++m) {
    const int i = i0 + m;
    if (i < g.N) {
      TfNodeIn in;
      node_inputs<M>(in, win, m, i, g, b, sys);
      double f[V];
      tf_model_F<false>(cst, in, f);
#pragma unroll
      for (int e = 0; e < V; ++e) out[((long long)sys * g.N + i) * V + e] = f[e];
    }
  }
}

// nonzero Jacobian values per node, natural layout [sys][node][nnz]
extern "C" __global__ void __launch_bounds__(256) tf_k_eval_J(Geom g, Buf b, double* __restrict__ out) {
  const int chunks = g.nblk * 32;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)chunks * g.batch) return;
  const int sys = (int)(t / chunks), chunk = (int)(t % chunks);
  const int i0 = chunk * M;
  if (i0 >= g.N) return;
  const double* cst = b.cst + (long long)sys * NC2;
  double win[NF][M + 2 * P];
  load_windows<M, 0>(win, i0, g, b, sys, nullptr);
#pragma unroll
  for (int m = 0; m < M; ++m) {
    const int i = i0 + m;
    if (i < g.N) {
      TfNodeIn in;
      node_inputs<M>(in, win, m, i, g, b, sys);
      double jv[NNZ];
      tf_model_J<false>(cst, in, jv);
#pragma unroll
      for (int k = 0; k < NNZ; ++k) out[((long long)sys * g.N + i) * NNZ + k] = jv[k];
    }
  }
}

// ---- factor: A = I - a*J(U) -> banded LU (chunk scan with linear-fractional maps)
__device__ __forceinline__ void factor_body_rows(const Geom& g, const Buf& b, double a_uniform) {
  __shared__ double smem[(MAXW + 1) * KMAX];
  int sys, tile, epoch;
  resolve_tile(g, b, sys, tile, epoch);
  if (b.active != nullptr && !b.active[sys]) return;       // finished ensemble member
  const double a = (b.asys != nullptr) ? b.asys[sys] : a_uniform;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int blk = tile * nwarps + warp;
  const bool active = blk < g.nblk;
  const int chunk = blk * 32 + lane;
  __shared__ double s_cst[NC2];
  for (int k = threadIdx.x; k < NC2; k += blockDim.x) s_cst[k] = b.cst[(long long)sys * NC2 + k];
  __syncthreads();
  const double* cst = s_cst;
  if (tile == 0 && threadIdx.x == 0) b.err[sys] = 0.0;     // consumed by the last bwd
  double A[C + BETA][WB];
  Star mine = Star::identity();
  int bad = 0;
  if (active) {
    assemble_rows(A, chunk * M, g, b, sys, a, cst);
    tfb::ChunkLU<BETA, C>::run1(A, mine, bad);
  }
  const Star pre = tile_scan<Star>(mine, smem, g.tiles > 1, b, (long long)sys * g.tiles, tile,
                                   epoch, Star::identity(), nullptr);
  if (active) {
    double Uf[C][BETA + 1], Lown[C][BETA], Lnext[BETA][BETA];
    tfb::ChunkLU<BETA, C>::run2(A, pre.P(), Uf, Lown, Lnext, bad);
    double* Lg = b.Lf + sys * vstride(g) * BETA;
    double* Ug = b.Uf + sys * vstride(g) * (BETA + 1);
#pragma unroll
    for (int r = 0; r < C; ++r) {
#pragma unroll
      for (int q = 0; q <= BETA; ++q)
        Ug[((long long)blk * C + r) * 32 * (BETA + 1) + q * 32 + lane] = Uf[r][q];
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
    if (bad) atomicOr(b.status + sys, bad);
  }
  finish_chain(g, b, epoch);
}

// ---- streaming factorisation for scalar models (V == 1: rows == nodes, BETA == P).
// The chunk is walked in sub-blocks of BETA rows.  Pass 1 carries the chunk's
// linear-fractional map (P,Q,R,S) through the sub-blocks -- block elimination with a
// symbolic incoming update -- keeping O(BETA^2) state; pass 2 re-evaluates the rows
// (J is cheap) and does the scalar elimination with the true incoming update,
// writing L and U as it goes.  No row storage, hence few registers.
template <int NODES>
__device__ __forceinline__ void node_row(double (&row)[WB], const double (&win)[NF][NODES + 2 * P],
                                         int m, int i0, const Geom& g, const Buf& b, int sys, double a,
                                         const double* cst, bool all_regular) {
  const int i = i0 + m;
  if (all_regular) {                              // whole chunk away from the domain ends
    TfNodeIn in;
    node_inputs<NODES>(in, win, m, i, g, b, sys);
    double jv[NNZ];
    tf_model_J<FD>(cst, in, jv);
#pragma unroll
    for (int d = 0; d < WB; ++d) row[d] = 0.0;
    row[BETA] = 1.0;
#pragma unroll
    for (int kk = 0; kk < NNZ; ++kk) {
      const int d = tf_j_off(kk);
      const double s = __dmul_rn(a, jv[kk]);
      row[BETA + d] = (d == 0) ? __dsub_rn(1.0, s) : -s;
    }
    return;
  }
  const int npad = g.nblk * 32 * M;
#pragma unroll
  for (int d = 0; d < WB; ++d) row[d] = 0.0;
  if (i >= npad) return;                          // beyond the system: no coupling
  double jv[NNZ];
#pragma unroll
  for (int kk = 0; kk < NNZ; ++kk) jv[kk] = 0.0;
  if (i < g.N) {
    TfNodeIn in;
    node_inputs<NODES>(in, win, m, i, g, b, sys);
    tf_model_J<FD>(cst, in, jv);
  }
  if (i >= P && i < g.N - 2 * P) {
    row[BETA] = 1.0;
#pragma unroll
    for (int kk = 0; kk < NNZ; ++kk) {
      const int d = tf_j_off(kk);
      const double s = __dmul_rn(a, jv[kk]);
      row[BETA + d] = (d == 0) ? __dsub_rn(1.0, s) : -s;
    }
  } else {
    double jvl[NNZ], tmp[WB];                     // the rare path works on private copies
#pragma unroll
    for (int kk = 0; kk < NNZ; ++kk) jvl[kk] = jv[kk];
    assemble_special(i, g, jvl, a, tmp, m < M ? b.btab + (long long)sys * 5 * NB * NB : nullptr);
#pragma unroll
    for (int d = 0; d < WB; ++d) row[d] = tmp[d];
  }
}

__device__ __forceinline__ void factor_body_stream(const Geom& g, const Buf& b, double a_uniform) {
  static_assert(V == 1 || true, "");
  constexpr int NODES = M + EX;
  constexpr int NSB = C / BETA;                   // sub-blocks per chunk
  __shared__ double smem[(MAXW + 1) * KMAX];
  __shared__ double s_cst[NC2];
  int sys, tile, epoch;
  resolve_tile(g, b, sys, tile, epoch);
  if (b.active != nullptr && !b.active[sys]) return;       // finished ensemble member
  const double a = (b.asys != nullptr) ? b.asys[sys] : a_uniform;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int blk = tile * nwarps + warp;
  const bool active = blk < g.nblk;
  const int chunk = blk * 32 + lane;
  const int i0 = chunk * M;
  for (int k = threadIdx.x; k < NC2; k += blockDim.x) s_cst[k] = b.cst[(long long)sys * NC2 + k];
  __syncthreads();
  const double* cst = s_cst;
  if (tile == 0 && threadIdx.x == 0) b.err[sys] = 0.0;     // consumed by the last bwd
  double win[NF][NODES + 2 * P];
  Star mine = Star::identity();
  int bad = 0;
  // warp-uniform: in the warp that holds rows at the ends of the domain every lane takes the
  // generic row path (nearly as fast for regular rows) -- with a per-lane choice that warp ran
  // the whole row assembly twice, once per side of the branch (the first and the last tile are
  // on the critical path of every scan)
  const bool allreg = __all_sync(0xffffffffu, i0 >= P && i0 + NODES <= g.N - 2 * P);
  if (active) {
    load_windows<NODES, 0>(win, i0, g, b, sys, nullptr);
    double cur[BETA][WB], nxt[BETA][WB];
#pragma unroll
    for (int r = 0; r < BETA; ++r) node_row<NODES>(cur[r], win, r, i0, g, b, sys, a, cst, allreg);
#pragma unroll
    for (int k = 0; k < NSB; ++k) {
#pragma unroll
      for (int r = 0; r < BETA; ++r) node_row<NODES>(nxt[r], win, (k + 1) * BETA + r, i0, g, b, sys, a, cst, allreg);
      // Dh = D_k - P ; solve Dh [Z1 | Z2] = [C_k | Q]
      double Dh[BETA * BETA], Z[BETA * 2 * BETA], Rr[BETA * BETA];
#pragma unroll
      for (int r = 0; r < BETA; ++r)
#pragma unroll
        for (int c = 0; c < BETA; ++c) {
          Dh[r * BETA + c] = cur[r][BETA + c - r] - mine.P()[r * BETA + c];
          Z[r * 2 * BETA + c] = (c <= r) ? cur[r][BETA + BETA + c - r] : 0.0;      // C_k
          Z[r * 2 * BETA + BETA + c] = mine.Q()[r * BETA + c];
          Rr[r * BETA + c] = (c >= r) ? nxt[r][BETA + c - BETA - r] : 0.0;          // R_k
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
  const Star pre = tile_scan<Star>(mine, smem, g.tiles > 1, b, (long long)sys * g.tiles, tile,
                                   epoch, Star::identity(), nullptr);
  if (active) {
    double X[BETA * BETA];
#pragma unroll
    for (int q = 0; q < BETA * BETA; ++q) X[q] = pre.P()[q];
    double* Lg = b.Lf + sys * vstride(g) * BETA;
    double* Ug = b.Uf + sys * vstride(g) * (BETA + 1);
    double Lprev[BETA][BETA];
#pragma unroll
    for (int r = 0; r < BETA; ++r)
#pragma unroll
      for (int q = 0; q < BETA; ++q) Lprev[r][q] = 0.0;
    double cur[BETA][WB], nxt[BETA][WB];
#pragma unroll
    for (int r = 0; r < BETA; ++r) node_row<NODES>(cur[r], win, r, i0, g, b, sys, a, cst, allreg);
#pragma unroll
    for (int k = 0; k < NSB; ++k) {
#pragma unroll
      for (int r = 0; r < BETA; ++r) node_row<NODES>(nxt[r], win, (k + 1) * BETA + r, i0, g, b, sys, a, cst, allreg);
      // window: rows of block k, then the sub-diagonal part of block k+1's rows
      double A2[2 * BETA][WB];
#pragma unroll
      for (int r = 0; r < BETA; ++r)
#pragma unroll
        for (int d = 0; d < WB; ++d) {
          A2[r][d] = cur[r][d];
          // row BETA + r, column offset d - BETA: keep columns < BETA (block k) only
          A2[BETA + r][d] = (BETA + r + d - BETA < BETA) ? nxt[r][d] : 0.0;
        }
      double Uf[BETA][BETA + 1], Lown[BETA][BETA], Lnext[BETA][BETA], Xo[BETA * BETA];
      tfb::ChunkLU<BETA, BETA>::run2x(A2, X, Uf, Lown, Lnext, Xo, bad);
#pragma unroll
      for (int r = 0; r < BETA; ++r) {
        const int row = k * BETA + r;
#pragma unroll
        for (int q = 0; q <= BETA; ++q)
          Ug[((long long)blk * C + row) * 32 * (BETA + 1) + q * 32 + lane] = Uf[r][q];
#pragma unroll
        for (int q = 1; q <= BETA; ++q) {
          const long long o = ((long long)blk * C + row) * 32 * BETA + (q - 1) * 32 + lane;
          if (q <= r) Lg[o] = Lown[r][q - 1];
          else if (k > 0) Lg[o] = Lprev[r][q - 1];
          else if (chunk == 0) Lg[o] = 0.0;
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
    const int nchunk = chunk + 1;
    if (nchunk < g.nblk * 32) {
      const int nb = nchunk >> 5, nl = nchunk & 31;
#pragma unroll
      for (int r = 0; r < BETA; ++r)
#pragma unroll
        for (int q = r + 1; q <= BETA; ++q)
          Lg[((long long)nb * C + r) * 32 * BETA + (q - 1) * 32 + nl] = Lprev[r][q - 1];
    }
    if (bad) atomicOr(b.status + sys, bad);
  }
  finish_chain(g, b, epoch);
}

#ifndef TF_FACTOR_STREAM
#define TF_FACTOR_STREAM (TF_NVAR == 1)
#endif
extern "C" __global__ void __launch_bounds__(NT_FACTOR, TF_FACTOR_MINB) tf_k_factor(Geom g, Buf b, double a) {
#if TF_FACTOR_STREAM
  if constexpr (V == 1 && C % BETA == 0) factor_body_stream(g, b, a);
  else factor_body_rows(g, b, a);
#else
  factor_body_rows(g, b, a);
#endif
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

extern "C" __global__ void __launch_bounds__(NT) tf_k_border_fill(Geom g, Buf b, double a_uniform) {
  __shared__ double smem[(MAXW + 1) * KMAX];
  __shared__ double s_red[MAXW][NB * NB];
  __shared__ int s_alive;
  const int sys = blockIdx.x;
  if (b.active != nullptr && !b.active[sys]) return;
  const double a = (b.asys != nullptr) ? b.asys[sys] : a_uniform;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const long long vs = vstride(g);
  const double* bt = b.btab + (long long)sys * 5 * NB * NB;
  const double* Lg = b.Lf + sys * vs * BETA;
  const double* Ug = b.Uf + sys * vs * (BETA + 1);
  double* Wg = b.Wb + sys * vs * NB;
  double* Gg = b.Gb + sys * vs * NB;
  const int ntile = (g.nblk + nwarps - 1) / nwarps;
  const int tile_rows = nwarps * 32 * C;
  const int bot0 = g.nhat - NB;
  // ---- top part: periodic corner entries (E_top, F_top).  Skipped when they vanish
  //      (non-periodic systems), otherwise walk tiles until the carried state is 0.
  bool any_top = false;
  for (int k = 0; k < NB * NB; ++k)
    any_top = any_top || (bt[0 * NB * NB + k] != 0.0) || (bt[2 * NB * NB + k] != 0.0);
  int lead_rows = 0;
  if (any_top) {
    double cw[NB][BETA], cg[NB][BETA];            // carried recurrence states
#pragma unroll
    for (int c = 0; c < NB; ++c)
#pragma unroll
      for (int t = 0; t < BETA; ++t) { cw[c][t] = 0.0; cg[c][t] = 0.0; }
    for (int tile = 0; tile < ntile; ++tile) {
      const int blk = tile * nwarps + warp;
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
        const Aff pre = tile_scan<Aff>(mine, smem, false, b, 0, 0, 0, carry, &total, true);
#pragma unroll
        for (int t = 0; t < BETA; ++t)
#pragma unroll
          for (int cc = 0; cc < NB; ++cc)
            if (cc == c) { if (isW) cw[cc][t] = total.c()[t]; else cg[cc][t] = total.c()[t]; }
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
      bool alive = false;
#pragma unroll
      for (int c = 0; c < NB; ++c)
#pragma unroll
        for (int t = 0; t < BETA; ++t) alive = alive || (cw[c][t] != 0.0) || (cg[c][t] != 0.0);
      __syncthreads();
      if (threadIdx.x == 0) s_alive = alive ? 1 : 0;
      __syncthreads();
      lead_rows = (tile + 1) * tile_rows;
      if (lead_rows > g.nhat) lead_rows = g.nhat;
      if (s_alive == 0) break;
    }
  }
  __syncthreads();
  // ---- bottom part: the natural coupling of the last NB interior rows to the border
  //      (E_bot, F_bot).  L^-1 only propagates downwards, so it is a local NB-row solve,
  //      superposed on what the top part left in those rows.
  if (threadIdx.x < 2 * NB) {
    const bool isW = threadIdx.x < NB;
    const int c = isW ? threadIdx.x : threadIdx.x - NB;
    double loc[NB];
    for (int j = 0; j < NB; ++j) {
      const int gr = bot0 + j;
      double v = -(a * (isW ? bt[1 * NB * NB + j * NB + c] : bt[3 * NB * NB + c * NB + j]));
      for (int q = 1; q <= BETA; ++q) {
        const int jj = j - q;
        if (jj < 0) break;
        const double coef = isW ? Lg[fidx(gr, q - 1, BETA)]
                                : Ug[fidx(gr - q, q, BETA + 1)] * Ug[fidx(gr - q, 0, BETA + 1)];
        v -= coef * loc[jj];
      }
      loc[j] = v;
      const double out = isW ? v : v * Ug[fidx(gr, 0, BETA + 1)];
      double* dst = (isW ? Wg : Gg) + fidx(gr, c, NB);
      *dst = (gr < lead_rows ? *dst : 0.0) + out;
    }
  }
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
    for (int r = lo + threadIdx.x; r < hi; r += blockDim.x) {
      double gv[NB], wv[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) { gv[c] = Gg[fidx(r, c, NB)]; wv[c] = Wg[fidx(r, c, NB)]; }
#pragma unroll
      for (int i = 0; i < NB; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) part[i][j] = __fma_rn(gv[i], wv[j], part[i][j]);
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
        for (int w = 0; w < nwarps; ++w) v += s_red[w][i * NB + j];
        S[i * NB + j] = __dsub_rn(__fma_rn(-a, bt[4 * NB * NB + i * NB + j], (i == j) ? 1.0 : 0.0), v);
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

// rows of the border fill that may be non-zero: the leading `lead` rows and the last
// NB interior rows
__device__ __forceinline__ bool fill_row(int gr, int lead, const Geom& g) {
  return gr < lead || (gr >= g.nhat - NB && gr < g.nhat);
}

// ---- forward substitution of one stage: rhs = dt*F(U_i) + sum cfac_j k_j ; L y = rhs
//      + per-tile partial sums of G^T y for the border solve
// ---- bulk asynchronous copies (TMA, cp.async.bulk) of a CTA's tile into shared memory.
// A tile is one contiguous byte range in the lane-transposed layout, so the whole
// tile is requested by one thread with one instruction per array and lands while
// the other threads do index work; completion is signalled on an mbarrier.
__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes,
                                         unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TF_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TF_DONE;\n"
      "bra TF_WAIT;\n"
      "TF_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}

// Streaming recurrences: only the last BETA values of the solution and of the BETA
// homogeneous solutions are kept, so the register footprint is O(BETA^2) and the
// factor rows are consumed as they arrive; the second pass re-reads them (L1/L2 hits).
struct RecState {
  double s[BETA];          // s[t] = value at distance t+1 behind the current row
  double h[BETA][BETA];    // h[j][t]: same for the homogeneous solution started from e_j
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int t = 0; t < BETA; ++t) {
      s[t] = 0.0;
#pragma unroll
      for (int j = 0; j < BETA; ++j) h[j][t] = (j == t) ? 1.0 : 0.0;
    }
  }
  // coef[q-1] multiplies the value q rows behind; scale multiplies the result
  __device__ __forceinline__ void step(const double (&coef)[BETA], double rhs, double scale) {
    double v = rhs;
#pragma unroll
    for (int q = 0; q < BETA; ++q) v -= coef[q] * s[q];
    v *= scale;
#pragma unroll
    for (int q = BETA - 1; q > 0; --q) s[q] = s[q - 1];
    s[0] = v;
#pragma unroll
    for (int j = 0; j < BETA; ++j) {
      double w = 0.0;
#pragma unroll
      for (int q = 0; q < BETA; ++q) w -= coef[q] * h[j][q];
      w *= scale;
#pragma unroll
      for (int q = BETA - 1; q > 0; --q) h[j][q] = h[j][q - 1];
      h[j][0] = w;
    }
  }
  __device__ __forceinline__ void to_map(Aff& m) const {
#pragma unroll
    for (int i = 0; i < BETA; ++i) {
      m.c()[i] = s[i];
#pragma unroll
      for (int j = 0; j < BETA; ++j) m.Phi()[i * BETA + j] = h[j][i];
    }
  }
};

template <int NPREV>
__device__ __forceinline__ void fwd_body(const Geom& g, const Buf& b, const Stage& st) {
  __shared__ double smem[(MAXW + 1) * KMAX];
  __shared__ double s_part[MAXW][NB];
  __shared__ __align__(8) unsigned long long s_bar;
  extern __shared__ __align__(128) double dsm[];
  int sys, tile, epoch;
  resolve_tile(g, b, sys, tile, epoch);
  if (b.active != nullptr && !b.active[sys]) return;       // finished ensemble member
  TF_STAMP(epoch, tile, 0);
  const double dt = (b.dtsys != nullptr) ? b.dtsys[sys] : st.dt;
  const int T = blockDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = T >> 5;
  const int glead = st.use_partials ? b.lead[sys * 2 + 1] : 0;    // needed late: load early
  double* sL = dsm;                        // [T*C*BETA] L rows of the tile (TMA)
  double* sS = dsm + (size_t)T * C * BETA; // [T*C]      stage state of the tile (halo sharing),
  double* sF = sS;                         //            reused as right-hand side stash
  const int blk0 = tile * nwarps;
  const int nact = (g.nblk - blk0) < nwarps ? (g.nblk - blk0) : nwarps;   // active warp-blocks
  const int blk = blk0 + warp;
  const bool active = warp < nact;
  const int chunk = blk * 32 + lane;
  const int i0 = chunk * M;
  __shared__ double s_cst[NC2];                                   // uniform constants of the system
  const double* cst = s_cst;
  const long long vs = vstride(g);
  const long long cb = ((long long)blk * C) * 32 + lane;          // own chunk, element 0
  const int sb = (warp * C) * 32 + lane;                          // same, inside the tile
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  for (int k = threadIdx.x; k < NC2; k += blockDim.x) s_cst[k] = b.cst[(long long)sys * NC2 + k];
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned bytes = (unsigned)(nact * C * 32 * BETA * sizeof(double));
    mbar_expect_tx(&s_bar, bytes);
    bulk_g2s(sL, b.Lf + (sys * vs + (long long)blk0 * C * 32) * BETA, bytes, &s_bar);
  }
  // own stage state -> registers -> shared (neighbours read their halo from there);
  // chunks at a tile / domain edge fetch their whole window from global memory in the
  // same round trip
  double own[C];
  double win[NF][M + 2 * P];
  const bool interior = i0 >= P && i0 + M + P <= g.N && P <= M && NH == 0 &&
                        chunk > blk0 * 32 && chunk < (blk0 + nact) * 32 - 1;
  if (active) {
    const double* U = b.U + sys * vs;
#pragma unroll
    for (int r = 0; r < C; ++r) own[r] = stage_value<NPREV>(U, b, sys * vs, cb + (long long)r * 32, &st);
    if (!interior) load_windows<M, NPREV>(win, i0, g, b, sys, &st);
#pragma unroll
    for (int r = 0; r < C; ++r) sS[sb + r * 32] = own[r];
  }
  __syncthreads();
  Aff mine = Aff::identity();
  if (active) {
    if (interior) {
#pragma unroll
      for (int w = 0; w < M + 2 * P; ++w) {
        const int rel = w - P;
        const int dc = rel < 0 ? -1 : (rel >= M ? 1 : 0);
        const int m = rel - dc * M;
        // neighbour chunk inside the tile: lane +- 1, possibly in the adjacent warp-block
        const int t2 = (int)threadIdx.x + dc;
        const int nb = ((t2 >> 5) * C) * 32 + (t2 & 31);
#pragma unroll
        for (int e = 0; e < V; ++e)
          win[e][w] = (dc == 0) ? own[(m * V + e) < C ? (m * V + e) : 0] : sS[nb + (m * V + e) * 32];
      }
    }
  }
  __syncthreads();                         // every halo has been read: sS becomes sF
  TF_STAMP(epoch, tile, 1);
  if (active) {
    mbar_wait(&s_bar, 0);
    RecState rs;
    rs.init();
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const int i = i0 + m;
      double fe[V];
#pragma unroll
      for (int e = 0; e < V; ++e) fe[e] = 0.0;
      if (i < g.N) {
        TfNodeIn in;
        node_inputs<M>(in, win, m, i, g, b, sys);
        tf_model_F_solver<FD>(cst, in, fe);
      }
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const int r = m * V + e;
        // explicit roundings: the system-resident kernel (tf_sysstep.cuh) must contract alike
        double rhs = __dmul_rn(dt, fe[e]);
#pragma unroll
        for (int q = 0; q < (NPREV < 0 ? MAXS : NPREV); ++q)
          if (NPREV >= 0 || q < st.nprev)
            rhs = __fma_rn(st.cfac[q], b.K[q][sys * vs + cb + (long long)r * 32], rhs);
        rhs = (i < g.N) ? rhs : 0.0;
        sF[sb + r * 32] = rhs;
        double coef[BETA];
#pragma unroll
        for (int q = 0; q < BETA; ++q) coef[q] = sL[((warp * C + r) * BETA + q) * 32 + lane];
        rs.step(coef, rhs, 1.0);
      }
    }
    rs.to_map(mine);
  }
  TF_STAMP(epoch, tile, 3);
  const Aff pre = tile_scan<Aff>(mine, smem, g.tiles > 1, b, (long long)sys * g.tiles, tile,
                                 epoch, Aff::identity(), nullptr);
  TF_STAMP(epoch, tile, 6);
  // second pass with the true incoming state; when the border fill of this step is
  // already complete (every stage but the first) the tile also reduces its share of G^T y
  const int tile_rows = nwarps * 32 * C;
  const int t0 = tile * tile_rows;
  // (the last NB interior rows are added by the backward sweep itself: non-periodic
  //  systems then skip this reduction altogether)
  const bool gtile = st.use_partials && t0 < glead;
  double acc[NB];
#pragma unroll
  for (int c = 0; c < NB; ++c) acc[c] = 0.0;
  if (active) {
    double sv[BETA];
#pragma unroll
    for (int t = 0; t < BETA; ++t) sv[t] = pre.c()[t];
    double* Y = b.Y + sys * vs + cb;
    const double* G = b.Gb + sys * vs * NB + cb * NB - (long long)lane * (NB - 1);
    const int r0 = chunk * C;
#pragma unroll
    for (int r = 0; r < C; ++r) {
      double v = sF[sb + r * 32];
#pragma unroll
      for (int q = 0; q < BETA; ++q) v -= sL[((warp * C + r) * BETA + q) * 32 + lane] * sv[q];
#pragma unroll
      for (int q = BETA - 1; q > 0; --q) sv[q] = sv[q - 1];
      sv[0] = v;
      Y[(long long)r * 32] = v;
      if (gtile && r0 + r < glead && r0 + r < g.nhat - NB) {
#pragma unroll
        for (int c = 0; c < NB; ++c) acc[c] += G[((long long)r * NB + c) * 32] * v;
      }
    }
  }
  if (gtile) {
#pragma unroll
    for (int c = 0; c < NB; ++c) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], d);
      if (lane == 0) s_part[warp][c] = acc[c];
    }
    __syncthreads();
    if (threadIdx.x < NB) {
      double v = 0.0;
      for (int w = 0; w < nwarps; ++w) v += s_part[w][threadIdx.x];
      b.gpart[((long long)sys * g.tiles + tile) * NB + threadIdx.x] = v;
    }
  }
  TF_STAMP(epoch, tile, 7);
  finish_chain(g, b, epoch);
}

#define TF_FWD_KERNEL(name, NP)                                                          \
  extern "C" __global__ void __launch_bounds__(NT, TF_MINB) name(Geom g, Buf b, Stage st) { \
    fwd_body<NP>(g, b, st);                                                               \
  }
TF_FWD_KERNEL(tf_k_fwd0, 0)
TF_FWD_KERNEL(tf_k_fwd1, 1)
TF_FWD_KERNEL(tf_k_fwd2, 2)
TF_FWD_KERNEL(tf_k_fwdg, -1)

// x_b from the per-tile partial sums the forward sweep left (fixed summation order);
// cheap enough that every chunk that needs x_b recomputes it: no barrier, no flag.
__device__ __forceinline__ void border_solution_partials(double (&xb)[NB], const Geom& g, const Buf& b,
                                                         int sys, int fwd_tiles, int fwd_tile_rows) {
  const double* Y = b.Y + sys * vstride(g);
  const int glead = b.lead[sys * 2 + 1];
  double acc[NB];
#pragma unroll
  for (int c = 0; c < NB; ++c) acc[c] = 0.0;
  const int nlead = (glead + fwd_tile_rows - 1) / fwd_tile_rows;
  for (int t = 0; t < nlead && t < fwd_tiles; ++t) {
#pragma unroll
    for (int c = 0; c < NB; ++c) acc[c] += __ldcg(b.gpart + ((long long)sys * fwd_tiles + t) * NB + c);
  }
  {
    const double* G = b.Gb + sys * vstride(g) * NB;
    for (int r = g.nhat - NB; r < g.nhat; ++r) {       // natural coupling of the last rows
      const double yr = Y[ridx(r)];
#pragma unroll
      for (int c = 0; c < NB; ++c) acc[c] = __fma_rn(G[fidx(r, c, NB)], yr, acc[c]);
    }
  }
  double yb[NB];
#pragma unroll
  for (int c = 0; c < NB; ++c) yb[c] = __dsub_rn(Y[ridx(g.nhat + c)], acc[c]);
  const double* Si = b.Sinv + (long long)sys * NB * NB;
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < NB; ++c) s = __fma_rn(Si[r * NB + c], yb[c], s);
    xb[r] = s;
  }
}

// x_b = S^-1 (y_b - G^T y).  G is non-zero only on the leading `lead` rows and the
// last NB interior rows.  The first tile of the backward sweep (it owns the border)
// reduces G^T y with the whole CTA in a fixed order and publishes x_b; the few other
// chunks that need it (top tiles, processed last) read it behind an epoch flag.
__device__ __forceinline__ void border_solution_cta(double (&xb)[NB], const Geom& g, const Buf& b,
                                                    const Stage& st, int sys, int epoch,
                                                    double (*s_red)[NB], double* s_xb) {
  const long long vs = vstride(g);
  const double* Y = b.Y + sys * vs;
  const double* G = b.Gb + sys * vs * NB;
  const int glead = b.lead[sys * 2 + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  double acc[NB];
#pragma unroll
  for (int c = 0; c < NB; ++c) acc[c] = 0.0;
  const int bot0 = g.nhat - NB;
  if (st.use_partials) {
    // the forward sweep left per-tile partial sums: add them in tile order
    if (threadIdx.x == 0) {
      const int ft = st.fwd_tiles, fr = st.fwd_tile_rows;
      const int nlead = (glead + fr - 1) / fr;
      for (int t = 0; t < nlead && t < ft; ++t) {
#pragma unroll
        for (int c = 0; c < NB; ++c) acc[c] += __ldcg(b.gpart + ((long long)sys * ft + t) * NB + c);
      }
      for (int r = bot0; r < g.nhat; ++r) {
        const double yr = Y[ridx(r)];
#pragma unroll
        for (int c = 0; c < NB; ++c) acc[c] += G[fidx(r, c, NB)] * yr;
      }
#pragma unroll
      for (int c = 0; c < NB; ++c) s_red[0][c] = acc[c];
    }
  } else
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const int lo = pass == 0 ? 0 : (glead > bot0 ? glead : bot0);
    const int hi = pass == 0 ? (glead < g.nhat ? glead : g.nhat) : g.nhat;
    for (int r = lo + (int)threadIdx.x; r < hi; r += (int)blockDim.x) {
      const double yr = Y[ridx(r)];
#pragma unroll
      for (int c = 0; c < NB; ++c) acc[c] += G[fidx(r, c, NB)] * yr;
    }
  }
  if (!st.use_partials) {
#pragma unroll
    for (int c = 0; c < NB; ++c) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], d);
      if (lane == 0) s_red[warp][c] = acc[c];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double yb[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      double v = 0.0;
      for (int w = 0; w < (st.use_partials ? 1 : nwarps); ++w) v += s_red[w][c];
      yb[c] = Y[ridx(g.nhat + c)] - v;
    }
    const double* Si = b.Sinv + (long long)sys * NB * NB;
#pragma unroll
    for (int r = 0; r < NB; ++r) {
      double sum = 0.0;
#pragma unroll
      for (int c = 0; c < NB; ++c) sum += Si[r * NB + c] * yb[c];
      s_xb[r] = sum;
      b.xb[(long long)sys * (NB + 1) + r] = sum;
    }
    st_flag((int*)(b.xb + (long long)sys * (NB + 1) + NB), epoch);
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < NB; ++c) xb[c] = s_xb[c];
}

// ---- backward substitution: U x = y - W x_b ; k_i = x - sum cfac_j k_j ;
//      last stage: U_new = U + sum b_i k_i and the embedded error estimate
template <int NPREV, int LAST>
__device__ __forceinline__ void bwd_body(const Geom& g, const Buf& b, const Stage& st) {
  __shared__ double smem[(MAXW + 1) * KMAX];
  __shared__ double s_err[MAXW];
  __shared__ double s_bred[MAXW][NB];
  __shared__ double s_xb[NB];
  __shared__ __align__(8) unsigned long long s_bar;
  extern __shared__ __align__(128) double dsm[];
  int sys, tile, epoch;
  resolve_tile(g, b, sys, tile, epoch);
  if (b.active != nullptr && !b.active[sys]) return;       // finished ensemble member
  const int T = blockDim.x;
  const int lane_l = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = T >> 5;
  double* sU = dsm;                               // [T*C*(BETA+1)] U rows of the tile (TMA)
  // logical (reversed) blocks [tile*nwarps, +nact) are the actual blocks [bmin, bmin+nact)
  const int bl0 = tile * nwarps;
  const int nact = (g.nblk - bl0) < nwarps ? (g.nblk - bl0) : nwarps;
  const int bmin = g.nblk - bl0 - nact;
  const bool active = warp < nact;
  const int blk = g.nblk - 1 - (bl0 + warp);
  const int lane = 31 - lane_l;
  const int chunk = blk * 32 + lane;
  const long long vs = vstride(g);
  const long long cb = ((long long)blk * C) * 32 + lane;
  const int sb = ((blk - bmin) * C) * 32 + lane;            // own chunk inside the tile
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned by = (unsigned)(nact * C * 32 * sizeof(double));
    mbar_expect_tx(&s_bar, by * (BETA + 1));
    bulk_g2s(sU, b.Uf + (sys * vs + (long long)bmin * C * 32) * (BETA + 1), by * (BETA + 1), &s_bar);
  }
  double yreg[C];
  if (active) {
    const double* Yg = b.Y + sys * vs + cb;
#pragma unroll
    for (int r = 0; r < C; ++r) yreg[r] = Yg[(long long)r * 32];
    // the second pass reads k_j (and U for the update): start those lines moving now
#pragma unroll
    for (int r = 0; r < C; ++r) {
#pragma unroll
      for (int q = 0; q < (NPREV < 0 ? MAXS : NPREV); ++q)
        if (NPREV >= 0 || q < st.nprev)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(b.K[q] + sys * vs + cb + (long long)r * 32));
      if (LAST != 0)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(b.U + sys * vs + cb + (long long)r * 32));
    }
  }
  const int r0 = chunk * C;
  const int wlead = b.lead[sys * 2 + 0];
  const bool flagged = active && (r0 < wlead || (r0 + C > g.nhat - NB && r0 < g.nhat + NB));
  const double* W = b.Wb + sys * vs * NB + cb * NB - (long long)lane * (NB - 1);
  double xb[NB];
#pragma unroll
  for (int c = 0; c < NB; ++c) xb[c] = 0.0;
  // border coupling of this chunk: y <- y - W x_b on fill rows, border rows <- x_b
  // (applied on the fly in both passes)
  if (st.use_partials) {
    if (flagged) border_solution_partials(xb, g, b, sys, st.fwd_tiles, st.fwd_tile_rows);
  } else if (tile == 0) {
    border_solution_cta(xb, g, b, st, sys, epoch, s_bred, s_xb);
  } else if (flagged) {
    const int* fl = (const int*)(b.xb + (long long)sys * (NB + 1) + NB);
    while (ld_flag(fl) != epoch) {}
    acquire_fence();
#pragma unroll
    for (int c = 0; c < NB; ++c) xb[c] = __ldcg(b.xb + (long long)sys * (NB + 1) + c);
  }
  auto load_y = [&](int r) -> double {
    double yv = yreg[r];
    if (flagged) {
      const int gr = r0 + r;
      if (fill_row(gr, wlead, g)) {
#pragma unroll
        for (int c = 0; c < NB; ++c) yv = __fma_rn(-W[((long long)r * NB + c) * 32], xb[c], yv);
      } else if (gr >= g.nhat && gr < g.nhat + NB) {
#pragma unroll
        for (int c = 0; c < NB; ++c) if (gr - g.nhat == c) yv = xb[c];
      }
    }
    return yv;
  };
  const bool last = LAST < 0 ? (st.is_last != 0) : (LAST != 0);
  constexpr int NQ = NPREV < 0 ? MAXS : NPREV;
  Aff mine = Aff::identity();
  mbar_wait(&s_bar, 0);
  if (active) {
    RecState rs;
    rs.init();
#pragma unroll
    for (int r = C - 1; r >= 0; --r) {
      double coef[BETA];
#pragma unroll
      for (int q = 0; q < BETA; ++q) coef[q] = sU[sb * (BETA + 1) - lane * BETA + (r * (BETA + 1) + q + 1) * 32];
      rs.step(coef, load_y(r), sU[sb * (BETA + 1) - lane * BETA + (r * (BETA + 1)) * 32]);
    }
    rs.to_map(mine);
  }
  const Aff pre = tile_scan<Aff>(mine, smem, g.tiles > 1, b, (long long)sys * g.tiles, tile,
                                 epoch, Aff::identity(), nullptr);
  double emax = 0.0;
  if (active) {
    double sv[BETA];
#pragma unroll
    for (int t = 0; t < BETA; ++t) sv[t] = pre.c()[t];
    double* Kout = b.K[st.istage] + sys * vs + cb;
#pragma unroll
    for (int r = C - 1; r >= 0; --r) {
      double v = load_y(r);
#pragma unroll
      for (int q = 0; q < BETA; ++q) v -= sU[sb * (BETA + 1) - lane * BETA + (r * (BETA + 1) + q + 1) * 32] * sv[q];
      v *= sU[sb * (BETA + 1) - lane * BETA + (r * (BETA + 1)) * 32];
#pragma unroll
      for (int q = BETA - 1; q > 0; --q) sv[q] = sv[q - 1];
      sv[0] = v;
      const long long a = cb + (long long)r * 32;
      double k = v;
      double kprev[NQ > 0 ? NQ : 1];
#pragma unroll
      for (int q = 0; q < NQ; ++q)
        if (NPREV >= 0 || q < st.nprev) { kprev[q] = b.K[q][sys * vs + a]; k = __fma_rn(-st.cfac[q], kprev[q], k); }
      if (!last) {
        Kout[(long long)r * 32] = k;
      } else {
        // U + ((b0 k0 + b1 k1) + ...) in the reference's summation order (schemes.py:164-170)
        double acc = 0.0, accp = 0.0;
#pragma unroll
        for (int q = 0; q <= NQ && q < MAXS; ++q)
          if (NPREV >= 0 || q <= st.nprev) {
            const double kq = (q < (NPREV < 0 ? st.nprev : NPREV)) ? kprev[q < NQ ? q : 0] : k;
            const double t = __dmul_rn(st.b[q], kq);
            acc = (q == 0) ? t : __dadd_rn(acc, t);
            const double tp = __dmul_rn(st.bp[q], kq);
            accp = (q == 0) ? tp : __dadd_rn(accp, tp);
          }
        const double un = __dadd_rn(b.U[sys * vs + a], acc);
        b.Un[sys * vs + a] = un;
        if (st.has_pred) {
          const double e = fabs(__dsub_rn(un, __dadd_rn(un, accp)));
          emax = (e > emax || e != e) ? e : emax;
        }
      }
    }
  }
  if (last && st.has_pred) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const double o = __shfl_xor_sync(0xffffffffu, emax, d);
      emax = (o > emax || o != o) ? o : emax;
    }
    if (lane_l == 0) s_err[warp] = emax;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < nwarps; ++w) emax = (s_err[w] > emax || s_err[w] != s_err[w]) ? s_err[w] : emax;
      atomicMax((unsigned long long*)(b.err + sys), (unsigned long long)__double_as_longlong(emax));
    }
  }
  finish_chain(g, b, epoch);
}

#define TF_BWD_KERNEL(name, NP, LS)                                                      \
  extern "C" __global__ void __launch_bounds__(NT, TF_MINB) name(Geom g, Buf b, Stage st) { \
    bwd_body<NP, LS>(g, b, st);                                                           \
  }
TF_BWD_KERNEL(tf_k_bwd0n, 0, 0)
TF_BWD_KERNEL(tf_k_bwd0l, 0, 1)
TF_BWD_KERNEL(tf_k_bwd1n, 1, 0)
TF_BWD_KERNEL(tf_k_bwd1l, 1, 1)
TF_BWD_KERNEL(tf_k_bwd2l, 2, 1)
TF_BWD_KERNEL(tf_k_bwdg, -1, -1)

#include "tf_sysstep.cuh"
#define TF_GS_MULTI 0
#include "tf_gridstep.cuh"      // tf_k_gridstep: one GPU
#undef TF_GS_MULTI
#define TF_GS_MULTI 1
#include "tf_gridstep.cuh"      // tf_k_gridstep_mr, tf_k_gs_seed: one grid over several GPUs
#undef TF_GS_MULTI

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

// ---- per-member step-size control of an ensemble: ROW_general._variable_step
// (reference core/schemes.py:176-238) run independently for every system.
struct TfCtl {            // one per system
  double t, next, dt_int, dt_try;
  int phase;              // 0: attempting, 1: final exact step to the target, 2: done
  int iters, nfs, fail;   // fail: 3 = max_iter, 4 = dt_min (TF_E* codes)
};

extern "C" __global__ void tf_k_ctl_init(int batch, TfCtl* c, double* dtsys, double* asys, int* act,
                                         const double* internal_dt, double t0, double dt_out,
                                         double gamma, const double* t0v, const double* dtv,
                                         const int* maskv) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= batch) return;
  // (t0v / dtv / maskv: per-member start time, span and participation, or null)
  const double ts = t0v ? t0v[s] : t0, span = dtv ? dtv[s] : dt_out;
  const int on = maskv ? maskv[s] : 1;
  double d = internal_dt[s] < 0.0 ? 1e-6 : internal_dt[s];
  d = d < span ? d : span;
  c[s].t = ts;
  c[s].next = ts + span;
  c[s].dt_int = on ? d : internal_dt[s];
  c[s].dt_try = d;
  c[s].phase = on ? 0 : 2;
  c[s].iters = 0;
  c[s].nfs = 0;
  c[s].fail = 0;
  dtsys[s] = d;
  asys[s] = gamma * d;
  act[s] = on;
}

// after one attempt of every active member: accept / reject, next dt, commit flags
extern "C" __global__ void tf_k_ctl_update(int batch, TfCtl* c, double* dtsys, double* asys, int* act,
                                           int* commit, const double* err, int* n_active,
                                           double tol, double safety, int max_iter, double dt_min,
                                           double gamma, int* status) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= batch) return;
  commit[s] = 0;
  if (!act[s]) return;
  TfCtl k = c[s];
  // a bad pivot only counts when the attempt it belongs to is taken (a rejected attempt with
  // too large a step is retried with a smaller one, like the reference's loop)
  const int bad = status[s];
  status[s] = 0;
  const double e = err[s];
  k.nfs += 1;
  if (k.phase == 1) {                       // the exact step to the target is always taken
    commit[s] = 1;
    k.t = k.next;
    k.phase = 2;
    act[s] = 0;
  } else {
    const double new_t = k.t + k.dt_try;
    k.dt_int = safety * k.dt_try * sqrt(tol / e);
    if (e > tol) {                          // rejected: same state, smaller step
      k.dt_try = k.dt_int;
    } else if (new_t >= k.next) {           // passes the output time: redo with the exact dt
      k.phase = 1;
      k.dt_try = k.next - k.t;
    } else {                                // accepted internal step
      commit[s] = 1;
      k.t = new_t;
      k.iters += 1;
      k.dt_try = k.dt_int;
      if (max_iter > 0 && k.iters > max_iter) { k.fail = 3; k.phase = 2; act[s] = 0; }
      else if (dt_min > 0.0 && k.dt_int < dt_min) { k.fail = 4; k.phase = 2; act[s] = 0; }
    }
  }
  if (commit[s] && bad && !k.fail) { k.fail = 5; k.phase = 2; act[s] = 0; }   // TF_ESINGULAR
  dtsys[s] = k.dt_try;
  asys[s] = gamma * k.dt_try;
  c[s] = k;
  if (act[s]) atomicAdd(n_active, 1);
}

// U <- U_new for the members whose attempt was accepted
extern "C" __global__ void tf_k_commit(Geom g, double* __restrict__ U, const double* __restrict__ Un,
                                       const int* __restrict__ commit) {
  const int sys = blockIdx.x;                 // (batch in grid.x: no 65535 limit)
  if (!commit[sys]) return;
  const long long vs = vstride(g);
  for (long long a = blockIdx.y * (long long)blockDim.x + threadIdx.x; a < vs;
       a += (long long)gridDim.y * blockDim.x)
    U[sys * vs + a] = Un[sys * vs + a];
}

// ---- outer Richardson controller of an ensemble: schemes.time_stepping (reference
// core/schemes.py:33-66, the wrapper Simulation puts around EVERY scheme,
// simulation.py:190-197), run independently for every member.  One attempt = one coarse
// scheme call over m*dt_ against ten fine calls over dt_; the host only sequences the calls.
struct TfRich {           // one per system
  double t, target, idt, base, trial;
  int phase;              // 0: attempting, 1: last call up to the target, 2: done
  int calls, nfs, fail;
};
__device__ __forceinline__ void rich_decide(TfRich& r, int m) {
  if (r.t + r.idt <= r.target) {            // while t + internal_dt <= next_step: one_step(internal_dt / m)
    r.phase = 0;
    r.base = r.idt / m;
    r.trial = r.base;
  } else if (r.t < r.target) {
    r.phase = 1;
  } else {
    r.phase = 2;
  }
}
extern "C" __global__ void tf_k_rich_init(int batch, TfRich* R, const double* idt_in, double t0,
                                          double dt_out, int m) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= batch) return;
  TfRich r;
  r.t = t0;
  r.target = t0 + dt_out;
  r.idt = (idt_in[s] > 0.0) ? idt_in[s] : dt_out;      // "internal_dt if internal_dt else dt"
  r.base = r.trial = 0.0;
  r.calls = r.nfs = r.fail = 0;
  rich_decide(r, m);
  R[s] = r;
}
// arguments of the next scheme call of every member; which: 0 coarse (or the last call), 1 fine
extern "C" __global__ void tf_k_rich_call(int batch, const TfRich* R, int which, double* dtcall,
                                          double* acall, double* tcall, int* mask, int* n_on,
                                          int m, double gamma) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= batch) return;
  const TfRich r = R[s];
  int on = 0;
  double d = 0.0;
  if (which == 0) {
    if (r.phase == 0) { on = 1; d = m * r.trial; }
    else if (r.phase == 1) { on = 1; d = r.target - r.t; }
  } else if (r.phase == 0) {
    on = 1;
    d = r.trial;
  }
  dtcall[s] = d;
  acall[s] = gamma * d;
  tcall[s] = r.t;
  mask[s] = on;
  if (on) atomicAdd(n_on, 1);
}
// dst <- src for the members in the given phase
extern "C" __global__ void tf_k_rich_copy(Geom g, const TfRich* R, int phase, double* __restrict__ dst,
                                          const double* __restrict__ src) {
  const int sys = blockIdx.x;
  if (R[sys].phase != phase) return;
  const long long vs = vstride(g);
  for (long long a = blockIdx.y * (long long)blockDim.x + threadIdx.x; a < vs;
       a += (long long)gridDim.y * blockDim.x)
    dst[sys * vs + a] = src[sys * vs + a];
}
// bookkeeping after a scheme call (the call returns t + dt)
extern "C" __global__ void tf_k_rich_after(int batch, TfRich* R, int which, const int* mask,
                                           const TfCtl* c) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= batch || !mask[s]) return;
  TfRich r = R[s];
  r.calls += 1;
  if (c != nullptr) {                       // inner scheme with its own controller
    r.nfs += c[s].nfs;
    if (c[s].fail && !r.fail) r.fail = c[s].fail;
  } else {
    r.nfs += 1;
  }
  if (which == 1) r.t = r.t + r.trial;
  else if (r.phase == 1) { r.t = r.target; r.phase = 2; }
  R[s] = r;
}
// err = max over variables of ||coarse - fine||_2 / (m^2 - 1); then the step-size decision
extern "C" __global__ void __launch_bounds__(256) tf_k_rich_update(Geom g, TfRich* R, const double* __restrict__ U,
                                                                   const double* __restrict__ Uc, double tol,
                                                                   int m, double reject) {
  const int sys = blockIdx.x;
  __shared__ double s_red[8][V];
  __shared__ TfRich s_r;
  if (threadIdx.x == 0) s_r = R[sys];
  __syncthreads();
  if (s_r.phase != 0) return;
  const long long vs = vstride(g);
  double acc[V];
#pragma unroll
  for (int e = 0; e < V; ++e) acc[e] = 0.0;
  for (int i = threadIdx.x; i < g.N; i += blockDim.x) {
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const double d = Uc[sys * vs + vidx(i, e)] - U[sys * vs + vidx(i, e)];
      acc[e] = __fma_rn(d, d, acc[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < V; ++e) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], d);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][e] = acc[e];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double err = 0.0;
    for (int e = 0; e < V; ++e) {
      double v = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += s_red[w][e];
      const double en = sqrt(v) / (double)(m * m - 1);
      err = (en > err || en != en) ? en : err;
    }
    TfRich r = s_r;
    const double nd = sqrt(r.base * r.base * tol / err);
    if (nd < r.base / reject) {
      r.trial = nd;                          // rejected: again, from the state reached
    } else {
      r.idt = nd;
      rich_decide(r, m);
    }
    if (r.fail) r.phase = 2;
    R[sys] = r;
  }
}
extern "C" __global__ void tf_k_rich_read(int batch, const TfRich* R, double* idt, int* calls, int* nfs,
                                          int* fail, int* n_on) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= batch) return;
  idt[s] = R[s].idt;
  calls[s] = R[s].calls;
  nfs[s] = R[s].nfs;
  fail[s] = R[s].fail;
  if (R[s].phase != 2) atomicAdd(n_on, 1);
}

extern "C" __global__ void tf_k_ctl_read(int batch, const TfCtl* c, double* internal_dt, int* nfs,
                                         int* fail, double* t) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= batch) return;
  internal_dt[s] = c[s].dt_int;
  nfs[s] = c[s].nfs;
  fail[s] = c[s].fail;
  t[s] = c[s].t;
}
