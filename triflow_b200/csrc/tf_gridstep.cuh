// Grid-resident implicit step: ONE cooperative launch does a whole Rosenbrock / Theta step of a
// single long grid.  Every CTA owns one tile of the grid for the whole step.
//
// The per-kernel pipeline of tf_kernels.cuh runs factor -> border fill -> s x (fwd, bwd) as
// 2 + 2s dependent launches; inside each launch all tiles load together, compute together and
// look back together, and the factor, y and every stage vector make a round trip through
// memory between launches (KS N = 2^20: 8 launches of 20-76 us for 46 us worth of HBM traffic).
// Here the tiles are resident (cooperative launch, one tile per CTA, one CTA per SM):
//
//  * the L part of the factor never leaves shared memory, the forward-substitution result y
//    never leaves the registers (the backward sweep of a stage runs in the same threads with
//    the mirrored scan), stage-state halos are a neighbouring thread's shared-memory element;
//  * tiles are coupled ONLY through 16-byte (value, tag) words (tf_kernels.cuh, LbWord):
//    look-back records of the chunk scans, the P stage-state values at each tile edge, the
//    L multipliers that cross a tile boundary, the periodic-border quantities (partial sums
//    of G^T y and G^T W, x_b).  No fence, no grid-wide barrier, no kernel boundary: a tile
//    starts the next phase as soon as its own neighbours are done;
//  * the border fill of periodic systems (W = L^-1 E, G^T = F^T U^-1, tf_k_border_fill of the
//    pipeline: a serial walk over the leading tiles) costs no phase of its own: both are
//    forward recurrences, so their states ride on the scan of the first forward sweep (one fat
//    affine map, one look-back), and every thread learns from its incoming state whether its
//    rows of W and G are non-zero at all;
//  * every thread owns G = 2 adjacent chunks (2M nodes) which sit in adjacent lanes of the
//    lane-transposed layout, so the per-node cost of the scans halves and each thread's two
//    recurrences interleave.
//
// Same algorithm as the pipeline (chunk scans with linear-fractional / affine maps, border
// block for the periodic corners), agreeing with it to rounding; tests compare both paths.
// Every wait is bounded: on time-out the status bit 4 is set and the launch still ends.
//
// Replaces, in one launch (reference file:line): compute_J_numpy + I - gamma*dt*J + factorized
// (compilers.py:292-332, schemes.py:146-149), the s stages of ROW_general._fixed_step
// (schemes.py:150-163), update and error norm (:164-174); Theta (:548-559) is the 1-stage case.
//
// Several GPUs, one grid: the same kernel, one slab of whole tiles per GPU.  A tile's
// records live in the record area of the GPU that owns the tile; the few tiles next to a slab
// boundary (and the border-block protocol) read their neighbours' words straight out of the
// peer GPU's memory over NVLink (IPC-mapped record areas, system-scope loads) -- the halo
// exchange and the reduced-system coupling are fused into the step kernel, there is no
// collective call and no kernel boundary between ranks.  The stage state at slab edges of the
// NEXT step travels the same way (words tagged with the next epoch).
//
// Scope: scalar models (V == 1) without helper fields, s <= 3, one system (batch == 1),
// tiles <= resident CTAs (per GPU).  Everything else keeps the per-kernel pipeline.
//
// This file is included twice (tf_kernels.cuh): TF_GS_MULTI 0 -> tf_k_gridstep (one GPU: nothing
// of the several-GPU form is compiled in, its branches cost the single-GPU kernel 17-25 %),
// TF_GS_MULTI 1 -> tf_k_gridstep_mr + tf_k_gs_seed (slab states).

#if (TF_NVAR == 1) && (TF_NHELP == 0)
#define TF_HAS_GRIDSTEP 1

#if TF_GS_MULTI
#define GS_NS gs_multi
#define TF_GS_KNAME tf_k_gridstep_mr
#else
#define GS_NS gs_single
#define TF_GS_KNAME tf_k_gridstep
#endif

#ifndef TF_GS_ONCE
#ifndef TF_GS_NT
#define TF_GS_NT 448      /* max threads per CTA: 14 warps (4 per scheduler at most) x 128 registers */
#endif
#ifndef TF_GS_SPIN
#define TF_GS_SPIN (1 << 19)
#endif

// Optional per-tile phase time stamps (cubin built with -DTF_GS_TRACE; tools/gs_trace.py):
// globaltimer ns of thread 0 at the phase boundaries of the last step.
#ifdef TF_GS_TRACE
__device__ unsigned long long tf_gs_trace[512 * 32];
#define GS_STAMP(ph) do { if (threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); tf_gs_trace[(cx.tile & 511) * 32 + (ph)] = t_; } } while (0)
// (latest thread of the tile instead of thread 0)
#define GS_STAMP_MAX(ph) do { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); atomicMax(&tf_gs_trace[(cx.tile & 511) * 32 + (ph)], t_); } while (0)
#else
#define GS_STAMP(ph) do { } while (0)
#define GS_STAMP_MAX(ph) do { } while (0)
#endif

// where every tile of an aborted launch was waiting (wait site, 0: not waiting); read by
// tools/slab_check.py through tf_model_read_symbol
__device__ int tf_gs_stuck[1024];
__device__ unsigned long long tf_gs_when[2];     // globaltimer: launch start (tile 0), first time-out
__device__ __forceinline__ void gs_note_stuck(int ltile, int site) {
  if (tf_gs_stuck[ltile & 1023] == 0) tf_gs_stuck[ltile & 1023] = site;
}

namespace tfk {

#ifndef TF_GS_G
#define TF_GS_G 2
#endif
constexpr int GS_G = TF_GS_G;                   // chunks per thread (1 or 2)
constexpr int GS_NT = TF_GS_NT;
constexpr int GS_MAXW = GS_NT / 32;
constexpr int GS_NPH = 1 + 2 * 3;               // look-back phases: factor, (fwd, bwd) x 3 stages
constexpr int GS_STAGES = 3;
static_assert(32 % GS_G == 0, "a thread's chunks lie in one warp-block");
static_assert(C % BETA == 0, "streaming factorisation walks sub-blocks of BETA rows");
// words per halo slot: the stage state of P nodes; at stage 0 the factorisation's windows reach
// P + EX nodes into the next tile (rows of the next chunk), read as words across GPUs
constexpr int GS_HW = P + EX;
static_assert(GS_HW <= M, "the first chunk of a tile publishes its left halo");
}  // namespace tfk
#endif  // TF_GS_ONCE

namespace tfk {
namespace GS_NS {
constexpr bool GS_MULTI = TF_GS_MULTI != 0;

// ---- map of the first forward sweep of a periodic system.  The border fill (W = L^-1 E,
//      G^T = F^T U^-1) consists of forward recurrences whose right-hand sides are non-zero in the
//      first NB rows of the system only, so the state entering any later chunk is
//          (propagators of all chunks in between) x (state leaving the system's first chunk).
//      The scan therefore carries, beside the substitution itself (PhiL, cy), only the prefix
//      product of the propagator of the G recurrence (PhiG; forward substitution with U^T in
//      "push" form: the state is the sum already pushed onto the next BETA rows).  W shares
//      PhiL.  The first chunk enters with identity propagators: an exclusive prefix never uses
//      the first element's propagator for its c-part, and the products then start behind it.
struct AffB {
  static constexpr int K = 2 * BETA * BETA + BETA;
  double d[K];
  __device__ __forceinline__ double* PhiL() { return d; }
  __device__ __forceinline__ double* cy() { return d + BETA * BETA; }
  __device__ __forceinline__ double* PhiG() { return d + BETA * BETA + BETA; }
  __device__ __forceinline__ const double* PhiL() const { return d; }
  __device__ __forceinline__ const double* cy() const { return d + BETA * BETA; }
  __device__ __forceinline__ const double* PhiG() const { return d + BETA * BETA + BETA; }
  __device__ static __forceinline__ AffB identity() {
    AffB m;
#pragma unroll
    for (int i = 0; i < K; ++i) m.d[i] = 0.0;
#pragma unroll
    for (int i = 0; i < BETA; ++i) { m.PhiL()[i * BETA + i] = 1.0; m.PhiG()[i * BETA + i] = 1.0; }
    return m;
  }
  // `a` first (earlier rows), `b` second
  __device__ static __forceinline__ AffB combine(const AffB& a, const AffB& b) {
    AffB o;
    tfb::mm<BETA>(b.PhiL(), a.PhiL(), o.PhiL());
    tfb::mm<BETA>(b.PhiG(), a.PhiG(), o.PhiG());
#pragma unroll
    for (int i = 0; i < BETA; ++i) {
      double sy = b.cy()[i];
#pragma unroll
      for (int k = 0; k < BETA; ++k) sy += b.PhiL()[i * BETA + k] * a.cy()[k];
      o.cy()[i] = sy;
    }
    return o;
  }
};
using tfk::absorbing;            // (the overloads for the pipeline's maps)
__device__ __forceinline__ bool absorbing(const AffB& m) {
  bool z = true;
#pragma unroll
  for (int k = 0; k < BETA * BETA; ++k) z = z && (m.PhiL()[k] == 0.0) && (m.PhiG()[k] == 0.0);
  return z;
}
constexpr int GS_KMAX = KMAX > AffB::K ? KMAX : AffB::K;

// ---- several GPUs: kernel argument describing the slab decomposition (tf_params.h: TfGsMulti)
// ---- record area (device memory, 16-byte words).  Every rank owns `tiles` (= tiles per rank)
//      consecutive tiles and keeps THEIR records in its own area `bases[rank]`; accessors take
//      the global tile index.  The single words live with the rank that writes them: w0 with
//      rank 0 (first chunk), x_b and F_top with the last rank (border rows).
struct GsRec {
  LbWord* bases[TF_GS_MAXRANKS];
  LbWord* base;          // own area
  int tiles;             // tiles per rank
  int nranks;
  int first;             // global index of this rank's first tile
  __device__ __forceinline__ LbWord* area(int& tile) const {     // owner's area; tile -> local index
    if (!GS_MULTI || nranks == 1) return base;
    const int lt = tile - first;
    if ((unsigned)lt < (unsigned)tiles) { tile = lt; return base; }   // (the common case: own tile)
    const int r = tile / tiles;
    tile -= r * tiles;
    return bases[r];
  }
  __device__ __forceinline__ LbWord* lb(int ph, int tile, int which) const {   // which: 0 agg, 1 inc
    LbWord* base = area(tile);
    return base + (((long long)ph * tiles + tile) * 2 + which) * GS_KMAX;
  }
  __device__ __forceinline__ long long o_halo() const { return (long long)GS_NPH * tiles * 2 * GS_KMAX; }
  __device__ __forceinline__ LbWord* halo(int stage, int tile, int slot) const {
    LbWord* base = area(tile);
    return base + o_halo() + (((long long)stage * tiles + tile) * 3 + slot) * GS_HW;
  }
  __device__ __forceinline__ long long o_lnext() const { return o_halo() + (long long)GS_STAGES * tiles * 3 * GS_HW; }
  __device__ __forceinline__ LbWord* lnext(int tile) const {
    LbWord* base = area(tile);
    return base + o_lnext() + (long long)tile * BETA * BETA;
  }
  __device__ __forceinline__ long long o_alive() const { return o_lnext() + (long long)tiles * BETA * BETA; }
  __device__ __forceinline__ LbWord* alive(int tile) const {
    LbWord* base = area(tile);
    return base + o_alive() + tile;
  }
  __device__ __forceinline__ long long o_gpart() const { return o_alive() + tiles; }
  __device__ __forceinline__ LbWord* gpart(int stage, int tile) const {
    LbWord* base = area(tile);
    return base + o_gpart() + ((long long)stage * tiles + tile) * NB;
  }
  __device__ __forceinline__ long long o_spart() const { return o_gpart() + (long long)GS_STAGES * tiles * NB; }
  __device__ __forceinline__ LbWord* spart(int tile) const {
    LbWord* base = area(tile);
    return base + o_spart() + (long long)tile * NB * NB;
  }
  __device__ __forceinline__ long long o_misc() const { return o_spart() + (long long)tiles * NB * NB; }
  __device__ __forceinline__ LbWord* last_area() const { return (!GS_MULTI || nranks == 1) ? base : bases[nranks - 1]; }
  __device__ __forceinline__ LbWord* first_area() const { return (!GS_MULTI || nranks == 1) ? base : bases[0]; }
  __device__ __forceinline__ LbWord* xb(int stage) const { return last_area() + o_misc() + stage * NB; }
  __device__ __forceinline__ LbWord* ftop() const { return last_area() + o_misc() + GS_STAGES * NB; }
  __device__ __forceinline__ LbWord* w0() const { return first_area() + o_misc() + GS_STAGES * NB + NB * NB; }   // [2][NB][BETA]
  __device__ __forceinline__ long long o_err() const { return o_misc() + GS_STAGES * NB + NB * NB + 2 * NB * BETA; }
  __device__ __forceinline__ double* errt() const { return (double*)(base + o_err()); }   // [tiles] doubles
};
// (the host side mirrors the size of this area in tf_host.cu: gs_words)

struct GsShared {
  double scan[(GS_MAXW + 2) * GS_KMAX];   // warp totals, [GS_MAXW]: tile prefix
  double cst[NC2];
  double halo[2][P];                   // stage state left / right of the tile
  double red[GS_MAXW][NB * NB > NB ? NB * NB : NB];
  double gather[32 * NB * NB];         // words of other tiles collected by the last tile
  double yb[2 * NB];                   // y of the last NB interior rows and of the border rows
  double xb[NB];
  double sinv[NB * NB];
  double ftop[NB * NB];                // periodic corner block F_top (lives with the border rows)
  double err[GS_MAXW];
  double gbot[NB * NB];                // last tile: G of the last NB interior rows [row][col]
  double sbot[NB * NB];                // last tile: G^T W over those rows
  int nalive;                          // last tile: leading tiles whose rows of W / G are non-zero
  int epoch;
  volatile int abort;
};

struct GsCtx {
  GsRec rec;
  GsShared* sh;
  int* status;
  long long tag;
  int tile, tiles, T;     // GLOBAL tile index / tile count
  int ltile;              // tile index on this rank
  int rank, nranks;
  int node_off;           // global index of this rank's first node (== first row: V == 1)
  int nodes_local;        // node slots of this rank's slab (tiles per rank x tile nodes)
};

// Across GPUs the words are read over NVLink: system scope (peer data is not cached in L2,
// and a gpu-scope load may be served from L1)
__device__ __forceinline__ void ld_word_sys(const LbWord* p, double& v, long long& t) {
  long long a;
  asm volatile("ld.relaxed.sys.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(t) : "l"(p) : "memory");
  v = __longlong_as_double(a);
}
__device__ __forceinline__ void st_word_sys(LbWord* p, double v, long long tag) {
  long long o0, o1;
  asm volatile(
      "{\n.reg .b128 q, r;\nmov.b128 q, {%3, %4};\n"
      "atom.relaxed.sys.global.exch.b128 r, [%2], q;\nmov.b128 {%0, %1}, r;\n}"
      : "=l"(o0), "=l"(o1)
      : "l"(p), "l"(__double_as_longlong(v)), "l"(tag)
      : "memory");
}

// ---- bounded waits on tagged words
// (the first tile that gives up records where: status = 4 | site << 8 | tile << 16)
__device__ __forceinline__ bool gs_aborted(GsCtx& cx, int spins, int site) {
  if (cx.sh->abort) { gs_note_stuck(cx.ltile, site); return true; }
  if ((spins & 1023) == 1023 && (ld_flag(cx.status) & 4)) {
    cx.sh->abort = 1;
    gs_note_stuck(cx.ltile, site);
    return true;
  }
  if (spins > TF_GS_SPIN) {
    gs_note_stuck(cx.ltile, site | 0x100);
    if (!(ld_flag(cx.status) & 4)) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tf_gs_when[1]));
    const int code = 4 | (site << 8) | (cx.tile << 16);
    int old = ld_flag(cx.status);
    while (!(old & 4)) {
      const int seen = atomicCAS(cx.status, old, old | code);
      if (seen == old) break;
      old = seen;
    }
    cx.sh->abort = 1;
    return true;
  }
  return false;
}
__device__ __noinline__ double gs_wait(GsCtx& cx, const LbWord* p, int site) {
  double v = 0.0;
  long long t;
  int spins = 0;
  while (true) {
    if (GS_MULTI && cx.nranks > 1) ld_word_sys(p, v, t); else ld_word(p, v, t);
    if (t == cx.tag) return v;
    if (gs_aborted(cx, ++spins, site)) return 0.0;
  }
}
// n consecutive words in ONE round trip per poll (all loads issued before any tag is compared)
template <int N_>
__device__ __noinline__ void gs_wait_many(GsCtx& cx, const LbWord* p, double (&out)[N_], int site) {
  int spins = 0;
  while (true) {
    long long t[N_];
#pragma unroll
    for (int k = 0; k < N_; ++k) {
      if (GS_MULTI && cx.nranks > 1) ld_word_sys(p + k, out[k], t[k]); else ld_word(p + k, out[k], t[k]);
    }
    bool ok = true;
#pragma unroll
    for (int k = 0; k < N_; ++k) ok = ok && (t[k] == cx.tag);
    if (ok) return;
    if (gs_aborted(cx, ++spins, site)) {
#pragma unroll
      for (int k = 0; k < N_; ++k) out[k] = 0.0;
      return;
    }
  }
}
__device__ __forceinline__ void gs_post_tag(GsCtx& cx, LbWord* p, double v, long long tag) {
  if (GS_MULTI && cx.nranks > 1) st_word_sys(p, v, tag); else st_word(p, v, tag);
}
__device__ __forceinline__ void gs_post(GsCtx& cx, LbWord* p, double v) { gs_post_tag(cx, p, v, cx.tag); }
// (look-back records: same, whole maps)
template <class Mon>
__device__ __forceinline__ void gs_publish(GsCtx& cx, LbWord* dst, const Mon& m, long long tag, int lane) {
  if (!GS_MULTI || cx.nranks == 1) { lb_publish(dst, m, tag, lane); return; }
  if constexpr (GS_MULTI) {
#pragma unroll
    for (int k0 = 0; k0 < Mon::K; k0 += 32) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k0 + k < Mon::K && lane == k) v = m.d[(k0 + k < Mon::K) ? k0 + k : 0];
      if (k0 + lane < Mon::K) st_word_sys(dst + k0 + lane, v, tag);
    }
  }
}
template <class Mon>
__device__ __forceinline__ void gs_read2(GsCtx& cx, const LbWord* pi, const LbWord* pa, Mon& mi, Mon& ma,
                                         long long FI, long long FA, bool& isI, bool& isA) {
  if (!GS_MULTI || cx.nranks == 1) { lb_read2(pi, pa, mi, ma, FI, FA, isI, isA); return; }
  if constexpr (GS_MULTI) {
    long long ti[Mon::K], ta[Mon::K];
#pragma unroll
    for (int k = 0; k < Mon::K; ++k) ld_word_sys(pi + k, mi.d[k], ti[k]);
#pragma unroll
    for (int k = 0; k < Mon::K; ++k) ld_word_sys(pa + k, ma.d[k], ta[k]);
    isI = true;
    isA = true;
#pragma unroll
    for (int k = 0; k < Mon::K; ++k) {
      isI = isI && (ti[k] == FI);
      isA = isA && (ta[k] == FA);
    }
  }
}

// ---- decoupled look-back (see lookback() in tf_kernels.cuh; bounded, own record arrays)
// (lt: position of the tile in the order of the scan; records are stored under the REAL tile
//  index, i.e. in the area of the GPU that owns and writes them)
template <class Mon>
__device__ __noinline__ Mon gs_lookback(const Mon& aggregate, GsCtx& cx, int ph, int lt, int lane,
                                        bool rev) {
  const long long FA = cx.tag * 4 + 1, FI = cx.tag * 4 + 2;
  auto real = [&](int l) { return rev ? cx.tiles - 1 - l : l; };
  // logical tile lt of phase ph lives in slot lt (the backward phases number tiles from the end)
  if (lt == 0) {
    gs_publish(cx, cx.rec.lb(ph, real(0), 1), aggregate, FI, lane);
    return Mon::identity();
  }
  gs_publish(cx, cx.rec.lb(ph, real(lt), 0), aggregate, FA, lane);
  Mon prefix = Mon::identity();
  int look = lt - 1;
  bool finished = false;
  int spins = 0;
  while (!finished) {
    const int t = look - lane;
    int need = TF_LB_FIRST;
    Mon w;
    Mon e = Mon::identity();
    bool isI = true, ready = (t < 0), first = true;
    while (true) {
      if (t >= 0 && (first || (lane < need && !ready))) {
        Mon ea;
        bool isA;
        gs_read2(cx, cx.rec.lb(ph, real(t), 1), cx.rec.lb(ph, real(t), 0), e, ea, FI, FA, isI, isA);
        if (!isI) e = ea;
        ready = isI || isA;
      }
      first = false;
      if (!__all_sync(0xffffffffu, ready || lane >= need)) {
        const bool ab = gs_aborted(cx, ++spins, 16 + ph);
        if (__any_sync(0xffffffffu, ab)) return Mon::identity();
        continue;
      }
      const unsigned mr = __ballot_sync(0xffffffffu, ready);
      const unsigned m2 = __ballot_sync(0xffffffffu, ready && isI);
      const int kr = (~mr == 0u) ? 32 : (__ffs(~mr) - 1);
      const int kstop = __ffs(m2) - 1;
      const bool hit = (kstop >= 0 && kstop < kr);
      const int last = hit ? kstop : kr - 1;
      w = select(t >= 0 && lane <= last, e, Mon::identity());
#pragma unroll 1
      for (int d = 1; d < 32; d <<= 1) {                    // ordered reduction, lane 0 = nearest
        const Mon o = shfl_down(w, d);
        const Mon c = Mon::combine(o, w);
        w = select(lane + d < 32, c, w);
      }
      w = shfl_idx(w, 0);
      if (__all_sync(0xffffffffu, hit || absorbing(w))) { finished = true; break; }
      if (kr == 32) break;
      need = 32;
    }
    prefix = Mon::combine(w, prefix);
    look -= 32;
  }
  gs_publish(cx, cx.rec.lb(ph, real(lt), 1), Mon::combine(prefix, aggregate), FI, lane);
  return prefix;
}

// ---- exclusive prefix of every thread's element over (thread, tile) order, or the mirrored
//      order (REV: backward substitution).  Warp shuffles -> shared memory -> look-back.
#ifndef TF_GS_SCAN_ATTR
#define TF_GS_SCAN_ATTR __forceinline__
#endif
template <class Mon, bool REV>
__device__ TF_GS_SCAN_ATTR Mon gs_scan(const Mon& mine, GsCtx& cx, int ph) {
  double* smem = cx.sh->scan;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = cx.T >> 5;
  const int ml = REV ? 31 - lane : lane;
  const int mw = REV ? nwarps - 1 - warp : warp;
  Mon incl = mine;
  // (level loops are rolled: the kernel is one long straight line, code size matters)
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    const Mon o = REV ? shfl_down(incl, d) : shfl_up(incl, d);
    const Mon c = Mon::combine(o, incl);
    incl = select(ml >= d, c, incl);
  }
  Mon excl = REV ? shfl_down(incl, 1) : shfl_up(incl, 1);
  excl = select(ml == 0, Mon::identity(), excl);
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
  Mon wi = w;
#pragma unroll 1
  for (int d = 1; d < GS_MAXW; d <<= 1) {
    const Mon o = shfl_up(wi, d);
    const Mon c = Mon::combine(o, wi);
    wi = select(lane >= d, c, wi);
  }
  Mon we = shfl_idx(wi, mw > 0 ? mw - 1 : 0);
  we = select(mw == 0, Mon::identity(), we);
  const Mon local = Mon::combine(we, excl);
  if (cx.tiles == 1) return local;
  if (warp == 0) {
    const Mon total = shfl_idx(wi, nwarps - 1);
    const int lt = REV ? cx.tiles - 1 - cx.tile : cx.tile;
    const Mon tp = gs_lookback(total, cx, ph, lt, lane, REV);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < Mon::K; ++k) smem[GS_MAXW * GS_KMAX + k] = tp.d[k];
    }
  }
  __syncthreads();
  Mon tp;
#pragma unroll
  for (int k = 0; k < Mon::K; ++k) tp.d[k] = smem[GS_MAXW * GS_KMAX + k];
  return Mon::combine(tp, local);
}

// ---- geometry of a thread's chunks
struct GsThread {
  int t, T;            // thread in tile, threads per tile
  int chunk0;          // first chunk, GLOBAL index (all row / node logic)
  int blk, cl;         // warp-block (on this rank) and lane of the first chunk in the layout
  bool active;
  __device__ __forceinline__ long long cb(int h) const { return ((long long)blk * C) * 32 + cl + h; }
};
// shared-memory layouts (thread-minor: conflict-free)
__device__ __forceinline__ int gs_sl(int r, int q, int h, int t, int T) { return ((r * BETA + q) * GS_G + h) * T + t; }
__device__ __forceinline__ int gs_ss(int r, int h, int t, int T) { return (r * GS_G + h) * T + t; }

// stage state of node j (any tile) from global memory: U + ((alpha_0 k_0 + ...)).  Valid for
// nodes of the own tile (same-CTA writes, ordered by a barrier) and, at stage 0, for any node.
// (IT >= 0: stage index at compile time; IT < 0: run-time index irt, at most GS_STAGES - 1 terms)
template <int IT>
__device__ __forceinline__ double gs_state_global(const GsCtx& cx, const Buf& b, int j, const TfStepDesc& sd,
                                                  int irt = 0) {
  constexpr int QMAX = IT < 0 ? GS_STAGES - 1 : IT;
  const int I = IT < 0 ? irt : IT;
  const long long a = vidx(GS_MULTI ? j - cx.node_off : j, 0);
  double u = b.U[a];
  if (QMAX > 0) {
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < QMAX; ++q)
      if (q < I) {
        const double term = __dmul_rn(sd.alpha[I][q], b.K[q][a]);
        acc = (q == 0) ? term : __dadd_rn(acc, term);
      }
    if (I > 0) u = __dadd_rn(u, acc);
  }
  return u;
}
// stage state of node jm (already wrapped / clamped into [0, N)) wherever it lives
template <int IT>
__device__ __noinline__ double gs_node_value(GsCtx& cx, const Geom& g, const Buf& b, int jm,
                                             const TfStepDesc& sd, int irt = 0) {
  const int I = IT < 0 ? irt : IT;
  const int TN = cx.T * GS_G * M;
  const int ot = jm / TN;
  const bool mine = !GS_MULTI || (jm >= cx.node_off && jm < cx.node_off + cx.nodes_local);
  if (ot == cx.tile || (I == 0 && mine)) return gs_state_global<IT>(cx, b, jm, sd, irt);
  const int loc = jm - ot * TN;
  if (loc < (I == 0 ? GS_HW : P)) return gs_wait(cx, cx.rec.halo(I, ot, 0) + loc, 10);
  if (loc >= TN - P) return gs_wait(cx, cx.rec.halo(I, ot, 1) + (loc - (TN - P)), 11);
  return gs_wait(cx, cx.rec.halo(I, ot, 2) + (jm - (g.N - P)), 12);
}

// U at nodes i0-P .. i0+NODES-1+P (global indices) for the Jacobian rows of a chunk: straight
// loads inside this rank's slab, the generic path (wrap / clamp, another rank's words) at the
// ends of the domain and of the slab
template <int NODES>
__device__ __forceinline__ void gs_load_windows(double (&win)[NF][NODES + 2 * P], int i0, GsCtx& cx,
                                                const Geom& g, const Buf& b, const TfStepDesc& sd) {
  const int lo = i0 - P, hi = i0 + NODES + P;
  if (!GS_MULTI) {                 // one GPU: the pipeline's loader (neighbour-lane addresses)
    load_windows<NODES, 0>(win, i0, g, b, 0, nullptr);
    return;
  }
  if (lo >= cx.node_off && hi <= cx.node_off + cx.nodes_local && lo >= 0 && hi <= g.N) {
    // inside the slab, away from the ends of the domain: the pipeline's loader on slab-local
    // indices (neighbour-lane addresses with constant offsets; a slab starts at a warp-block)
    Geom gl = g;
    gl.N = cx.nodes_local;
    gl.periodic = 0;
    load_windows<NODES, 0>(win, i0 - cx.node_off, gl, b, 0, nullptr);
    return;
  }
  // ends of the domain / of the slab: own nodes from memory, another rank's as words -- all
  // words of the window polled together (one NVLink round trip per poll, not one per node)
  const LbWord* wp[NODES + 2 * P];
  const int TN = cx.T * GS_G * M;
#pragma unroll 1
  for (int w = 0; w < NODES + 2 * P; ++w) {
    const int j = lo + w;
    wp[w] = nullptr;
    win[0][w] = 0.0;
    // (beyond the padding of the last tile nothing is read: rows there are identity rows)
    if (j >= g.N + P) continue;
    const int jm = map_node(j, g);
    if (jm >= cx.node_off && jm < cx.node_off + cx.nodes_local) {
      win[0][w] = b.U[vidx(jm - cx.node_off, 0)];
      continue;
    }
    const int ot = jm / TN, loc = jm - ot * TN;
    wp[w] = (loc < GS_HW) ? cx.rec.halo(0, ot, 0) + loc
          : (loc >= TN - P) ? cx.rec.halo(0, ot, 1) + (loc - (TN - P))
                            : cx.rec.halo(0, ot, 2) + (jm - (g.N - P));
  }
  int spins = 0;
  while (true) {
    bool ok = true;
#pragma unroll 1
    for (int w = 0; w < NODES + 2 * P; ++w)
      if (wp[w] != nullptr) {
        long long t;
        double v;
        ld_word_sys(wp[w], v, t);
        if (t == cx.tag) { win[0][w] = v; wp[w] = nullptr; }
        else ok = false;
      }
    if (ok) break;
    if (gs_aborted(cx, ++spins, 14)) break;
  }
}

// ------------------------------------------------------------------ factorisation
// pass 1 of one chunk: its linear-fractional map (factor_body_stream of tf_kernels.cuh)
__device__ __noinline__ void gs_factor_pass1(const Geom& g, const Buf& b, GsCtx& cx, const TfStepDesc& sd,
                                             int chunk, double a, const double* cst, Star& out, int& bad) {
  constexpr int NODES = M + EX;
  constexpr int NSB = C / BETA;
  const int i0 = chunk * M;
  // (uniform over the lanes that are here together: see factor_body_stream)
  const unsigned am = __activemask();
  const bool allreg = __all_sync(am, i0 >= P && i0 + NODES <= g.N - 2 * P);
  double win[NF][NODES + 2 * P];
  gs_load_windows<NODES>(win, i0, cx, g, b, sd);
  Star mine = Star::identity();
  double cur[BETA][WB], nxt[BETA][WB];
#pragma unroll
  for (int r = 0; r < BETA; ++r) node_row<NODES>(cur[r], win, r, i0, g, b, 0, a, cst, allreg);
  __syncwarp(am);      // the lane that walked special rows rejoins (else the warp stays split)
#pragma unroll
  for (int k = 0; k < NSB; ++k) {
#pragma unroll
    for (int r = 0; r < BETA; ++r) node_row<NODES>(nxt[r], win, (k + 1) * BETA + r, i0, g, b, 0, a, cst, allreg);
    __syncwarp(am);
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
  out = mine;
}

// pass 2 of one chunk: elimination with the true incoming update X (in/out); U rows -> global,
// L multipliers -> shared memory (own rows) / Lout (what is left on the next chunk's first rows)
__device__ __noinline__ void gs_factor_pass2(const Geom& g, const Buf& b, GsCtx& cx, const TfStepDesc& sd,
                                             const GsThread& th, int h,
                                             double a, const double* cst, double* sL, double* X,
                                             double (&Lout)[BETA][BETA], double* phiG, int& bad) {
  constexpr int NODES = M + EX;
  constexpr int NSB = C / BETA;
  const int chunk = th.chunk0 + h;
  const int i0 = chunk * M;
  // (uniform over the lanes that are here together: see factor_body_stream)
  const unsigned am = __activemask();
  const bool allreg = __all_sync(am, i0 >= P && i0 + NODES <= g.N - 2 * P);
  double win[NF][NODES + 2 * P];
  gs_load_windows<NODES>(win, i0, cx, g, b, sd);
  double* Ug = b.Uf + th.cb(h) * (BETA + 1) - (long long)(th.cl + h) * BETA;   // row r, entry q: [(r*(BETA+1)+q)*32]
  double Lprev[BETA][BETA];
#pragma unroll
  for (int r = 0; r < BETA; ++r)
#pragma unroll
    for (int q = 0; q < BETA; ++q) Lprev[r][q] = 0.0;
  // homogeneous solutions of the push recurrence of G^T = F^T U^-1 (see AffB): hG[j][t] is
  // what solution j has pushed onto the row t + 1 ahead
  double hG[BETA][BETA];
#pragma unroll
  for (int j = 0; j < BETA; ++j)
#pragma unroll
    for (int t = 0; t < BETA; ++t) hG[j][t] = (j == t) ? 1.0 : 0.0;
  double cur[BETA][WB], nxt[BETA][WB];
#pragma unroll
  for (int r = 0; r < BETA; ++r) node_row<NODES>(cur[r], win, r, i0, g, b, 0, a, cst, allreg);
  __syncwarp(am);      // the lane that walked special rows rejoins (else the warp stays split)
#pragma unroll
  for (int k = 0; k < NSB; ++k) {
#pragma unroll
    for (int r = 0; r < BETA; ++r) node_row<NODES>(nxt[r], win, (k + 1) * BETA + r, i0, g, b, 0, a, cst, allreg);
    __syncwarp(am);
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
      for (int q = 0; q <= BETA; ++q) Ug[(long long)(row * (BETA + 1) + q) * 32] = Uf[r][q];
#pragma unroll
      for (int j = 0; j < BETA; ++j) {
        const double gv = -(hG[j][0] * Uf[r][0]);
#pragma unroll
        for (int t = 0; t < BETA; ++t) hG[j][t] = ((t + 1 < BETA) ? hG[j][t + 1 < BETA ? t + 1 : 0] : 0.0) + Uf[r][t + 1] * gv;
      }
#pragma unroll
      for (int q = 1; q <= BETA; ++q) {
        if (q <= r) sL[gs_sl(row, q - 1, h, th.t, th.T)] = Lown[r][q - 1];
        else if (k > 0) sL[gs_sl(row, q - 1, h, th.t, th.T)] = Lprev[r][q - 1];
        // k == 0, q > r: multipliers with respect to the previous chunk's pivots (filled by it)
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
#pragma unroll
  for (int r = 0; r < BETA; ++r)
#pragma unroll
    for (int q = 0; q < BETA; ++q) Lout[r][q] = Lprev[r][q];
#pragma unroll
  for (int i = 0; i < BETA; ++i)
#pragma unroll
    for (int j = 0; j < BETA; ++j) phiG[i * BETA + j] = hG[j][i];
}

__device__ __forceinline__ void gs_factor(const Geom& g, const Buf& b, GsCtx& cx, const GsThread& th,
                                          const TfStepDesc& sd, double a, double* sL,
                                          double (&phiG)[GS_G][BETA * BETA]) {
  GsShared& sh = *cx.sh;
  int bad = 0;
  Star mine = Star::identity();
  if (th.active) {
    gs_factor_pass1(g, b, cx, sd, th.chunk0, a, sh.cst, mine, bad);
#pragma unroll 1
    for (int h = 1; h < GS_G; ++h) {
      Star m1;
      gs_factor_pass1(g, b, cx, sd, th.chunk0 + h, a, sh.cst, m1, bad);
      mine = Star::combine(mine, m1);
    }
  }
  GS_STAMP(1);
  GS_STAMP_MAX(22);
  const Star pre = gs_scan<Star, false>(mine, cx, 0);
  GS_STAMP(2);
#pragma unroll
  for (int h = 0; h < GS_G; ++h)
#pragma unroll
    for (int q = 0; q < BETA * BETA; ++q) phiG[h][q] = (q / BETA == q % BETA) ? 1.0 : 0.0;
  if (th.active) {
    double X[BETA * BETA];
#pragma unroll
    for (int q = 0; q < BETA * BETA; ++q) X[q] = pre.P()[q];
    double Lout[BETA][BETA];
#pragma unroll 1
    for (int h = 0; h < GS_G; ++h) {
      gs_factor_pass2(g, b, cx, sd, th, h, a, sh.cst, sL, X, Lout, phiG[h], bad);
      // multipliers left on the next chunk's first BETA rows (entries q > r)
      const bool cross = (h == GS_G - 1) && (th.t == th.T - 1);
      if (!cross) {
        const int h2 = (h + 1 < GS_G) ? h + 1 : 0;
        const int t2 = (h + 1 < GS_G) ? th.t : th.t + 1;
#pragma unroll
        for (int r = 0; r < BETA; ++r)
#pragma unroll
          for (int q = r + 1; q <= BETA; ++q) sL[gs_sl(r, q - 1, h2, t2, th.T)] = Lout[r][q - 1];
      } else if (cx.tile + 1 < cx.tiles) {
#pragma unroll
        for (int r = 0; r < BETA; ++r)
#pragma unroll
          for (int q = r + 1; q <= BETA; ++q) gs_post(cx, cx.rec.lnext(cx.tile) + r * BETA + q - 1, Lout[r][q - 1]);
      }
    }
    if (bad) atomicOr(cx.status, bad);
  }
  // first rows of the tile: multipliers with respect to the previous tile's pivots
  if (th.t < BETA * BETA) {                 // (one word per thread: a single round trip)
    const int r = th.t / BETA, q = th.t % BETA + 1;
    if (q > r)
      sL[gs_sl(r, q - 1, 0, 0, th.T)] =
          (cx.tile == 0) ? 0.0 : gs_wait(cx, cx.rec.lnext(cx.tile - 1) + r * BETA + q - 1, 1);
  }
  __syncthreads();
  // the periodic corner block F_top was assembled with the border rows (last tile); the first
  // chunk of the system needs it for its rows of G
  if (g.periodic && threadIdx.x < NB * NB) {
    const double v = b.btab[2 * NB * NB + threadIdx.x];
    if (cx.tiles == 1) sh.ftop[threadIdx.x] = v;
    else if (cx.tile == cx.tiles - 1) gs_post(cx, cx.rec.ftop() + threadIdx.x, v);   // (own area)
  }
  GS_STAMP(3);
}

// address of (global row gr, entry q) in an array of this rank's slab
__device__ __forceinline__ long long gs_fidx(const GsCtx& cx, int gr, int q, int width) {
  return fidx(GS_MULTI ? gr - cx.node_off : gr, q, width);
}

// ------------------------------------------------------------------ border block
// L multiplier (row gr, entry q) of the own tile from shared memory
__device__ __forceinline__ double gs_sl_at(const double* sL, int gr, int q, int row0, int T) {
  const int loc = gr - row0;
  const int ch = loc / C, r = loc - ch * C;
  return sL[gs_sl(r, q, ch % GS_G, ch / GS_G, T)];
}

// right-hand sides of the fill recurrences: -a * E_top (rows of W) / -a * F_top (rows of G);
// non-zero only in the first NB rows of the system
__device__ __forceinline__ double gs_fw(const double* bt, double a, int gr, int c) {
  return (gr < NB) ? -(a * bt[gr * NB + c]) : 0.0;
}
__device__ __forceinline__ double gs_fg(const double* ftop, double a, int gr, int c) {
  return (gr < NB) ? -(a * ftop[c * NB + gr]) : 0.0;
}

// partial sum over this tile's rows of G[r][i] * W[r][j] -> sh.red[0][i*NB+j]; a thread
// counts the rows it knows to be non-zero (its own if it is alive, and the bottom rows)
__device__ __forceinline__ void gs_gw_partial(const Geom& g, const Buf& b, GsCtx& cx, const GsThread& th,
                                              bool th_alive, bool with_bottom) {
  GsShared& sh = *cx.sh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = cx.T >> 5;
  const int bot0 = g.nhat - NB;
  double part[NB][NB];
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) part[i][j] = 0.0;
  if (th.active) {
#pragma unroll 1
    for (int h = 0; h < GS_G; ++h) {
      const int r0 = (th.chunk0 + h) * C;
      if (th_alive || (with_bottom && r0 + C > bot0 && r0 < g.nhat)) {
#pragma unroll
        for (int r = 0; r < C; ++r) {
          const int gr = r0 + r;
          if (gr < g.nhat && (th_alive || (with_bottom && gr >= bot0))) {
            double gv[NB], wv[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) { gv[c] = b.Gb[gs_fidx(cx, gr, c, NB)]; wv[c] = b.Wb[gs_fidx(cx, gr, c, NB)]; }
#pragma unroll
            for (int i = 0; i < NB; ++i)
#pragma unroll
              for (int j = 0; j < NB; ++j) part[i][j] = __fma_rn(gv[i], wv[j], part[i][j]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      double v = part[i][j];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      if (lane == 0) sh.red[warp][i * NB + j] = v;
    }
  __syncthreads();
  if (threadIdx.x < NB * NB) {
    double v = 0.0;
    for (int w = 0; w < nwarps; ++w) v += sh.red[w][threadIdx.x];
    sh.red[0][threadIdx.x] = v;          // (column x is read and written by thread x only)
  }
  __syncthreads();
}

// Sum over the alive tiles (0 .. sh.nalive-1) of `per` words each, in tile order, added to
// sh.red[0][0 .. per).  All threads of the (last) tile take part; one memory round trip per
// 32 tiles.
template <int KIND>   // 0: spart words, 1: gpart words of stage `which`
__device__ __noinline__ void gs_gather_sum(GsCtx& cx, int which, int per, int site) {
  GsShared& sh = *cx.sh;
  for (int base = 0; base < sh.nalive; base += 32) {
    const int n = sh.nalive - base < 32 ? sh.nalive - base : 32;
    for (int idx = threadIdx.x; idx < n * per; idx += cx.T)
      sh.gather[idx] = gs_wait(cx, (KIND == 0 ? cx.rec.spart(base + idx / per)
                                              : cx.rec.gpart(which, base + idx / per)) + idx % per, site);
    __syncthreads();
    if ((int)threadIdx.x < per) {
      double v = sh.red[0][threadIdx.x];
      for (int t = 0; t < n; ++t) v += sh.gather[t * per + threadIdx.x];
      sh.red[0][threadIdx.x] = v;
    }
    __syncthreads();
  }
}

// Last tile, right after the factorisation: the bottom rows of W and G (natural coupling of the
// last NB interior rows to the border block; a local NB-row solve) and their share of G^T W.
// The fill of a periodic system is added onto these rows later, if it reaches them at all.
__device__ __noinline__ void gs_border_bottom(const Geom& g, const Buf& b, GsCtx& cx, double a,
                                              const double* sL) {
  GsShared& sh = *cx.sh;
  const int TR = cx.T * GS_G * C;
  const int row0 = cx.tile * TR;
  const int bot0 = g.nhat - NB;
  const double* bt = b.btab;
  __shared__ double s_wbot[NB * NB];
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
        const double coef = isW ? gs_sl_at(sL, gr, q - 1, row0, cx.T)
                                : b.Uf[gs_fidx(cx, gr - q, q, BETA + 1)] * b.Uf[gs_fidx(cx, gr - q, 0, BETA + 1)];
        v -= coef * loc[jj];
      }
      loc[j] = v;
      const double out = isW ? v : v * b.Uf[gs_fidx(cx, gr, 0, BETA + 1)];
      ((isW ? b.Wb : b.Gb) + gs_fidx(cx, gr, c, NB))[0] = out;
      (isW ? s_wbot : sh.gbot)[j * NB + c] = out;
    }
  }
  __syncthreads();
  if (threadIdx.x < NB * NB) {
    const int i = threadIdx.x / NB, j = threadIdx.x % NB;
    double v = 0.0;
    for (int r = 0; r < NB; ++r) v = __fma_rn(sh.gbot[r * NB + i], s_wbot[r * NB + j], v);
    sh.sbot[threadIdx.x] = v;
  }
  __syncthreads();
}

// Last tile, after the first forward scan: how many leading tiles have non-zero rows of W / G,
// S = (I - a Ab) - G^T W and its inverse.  own_alive: the fill reached this tile too.
__device__ __noinline__ void gs_border_finish(const Geom& g, const Buf& b, GsCtx& cx, const GsThread& th,
                                              double a, bool th_alive, bool own_alive) {
  GsShared& sh = *cx.sh;
  const double* bt = b.btab;
  // the alive tiles are a prefix (a state that has become exactly zero stays zero)
  if (threadIdx.x < 32) {
    int na = 0;
    if (g.periodic) {
      for (int base = 0; base < cx.tiles - 1; base += 32) {
        const int t = base + (int)threadIdx.x;
        const bool al = (t < cx.tiles - 1) && gs_wait(cx, cx.rec.alive(t), 6) != 0.0;
        const unsigned m = __ballot_sync(0xffffffffu, al);
        const int lead = (~m == 0u) ? 32 : (__ffs(~m) - 1);
        na += lead;
        if (lead < 32) break;
      }
    }
    if (threadIdx.x == 0) sh.nalive = na;
  }
  if (own_alive) {
    gs_gw_partial(g, b, cx, th, th_alive, true);             // own rows incl. the bottom ones
  } else {
    if (threadIdx.x < NB * NB) sh.red[0][threadIdx.x] = sh.sbot[threadIdx.x];
    __syncthreads();
  }
  gs_gather_sum<0>(cx, 0, NB * NB, 7);
  if (threadIdx.x == 0) {
    double S[NB * NB], I[NB * NB];
    for (int i = 0; i < NB; ++i)
      for (int j = 0; j < NB; ++j) {
        S[i * NB + j] = __dsub_rn(__fma_rn(-a, bt[4 * NB * NB + i * NB + j], (i == j) ? 1.0 : 0.0),
                                  sh.red[0][i * NB + j]);
        I[i * NB + j] = (i == j) ? 1.0 : 0.0;
      }
    tfb::solve_inplace<NB, NB>(S, I);
    bool bad = false;
    for (int k = 0; k < NB * NB; ++k) {
      sh.sinv[k] = I[k];
      b.Sinv[k] = I[k];
      if (!(fabs(I[k]) < 1e300)) bad = true;
    }
    if (bad) atomicOr(cx.status, 2);
  }
  __syncthreads();
}

// ------------------------------------------------------------------ one stage
// th_alive / tile_alive: this thread's / this tile's rows of W and G may be non-zero (set by
// stage 0, used by every stage)
// IT >= 0 / LT: stage index and "last stage" at compile time.  IT < 0 (-DTF_GS_RTSTAGE): ONE
// body for the stages behind the first, index irt and lastrt at run time -- a third less code
// per step for a kernel whose instruction fetches stall it.
template <int IT, bool LT>
__device__ __forceinline__ void gs_stage(const Geom& g, const Buf& b, GsCtx& cx, const GsThread& th,
                                         const TfStepDesc& sd, double a, const double* sL, double* sS,
                                         const double (&phiG)[GS_G][BETA * BETA], bool& th_alive,
                                         bool& tile_alive, double& emax, int irt = 0, bool lastrt = false) {
  constexpr int QMAX = IT < 0 ? GS_STAGES - 1 : IT;      // bound of the loops over previous stages
  const int I = IT < 0 ? irt : IT;
  const bool LAST = IT < 0 ? lastrt : LT;
  GsShared& sh = *cx.sh;
  const int T = cx.T;
  const int TN = T * GS_G * M;
  const bool last_tile = cx.tile == cx.tiles - 1;
  const double* cst = sh.cst;
  const int bot0 = g.nhat - NB;
  double y[GS_G][C];
  // ---------------------------------------------------------------- forward
  {
    double own[GS_G][C];
#pragma unroll
    for (int h = 0; h < GS_G; ++h)
#pragma unroll
      for (int r = 0; r < C; ++r) own[h][r] = 0.0;
    if (th.active) {
#pragma unroll
      for (int h = 0; h < GS_G; ++h)
#pragma unroll
        for (int r = 0; r < C; ++r) {
          const long long aidx = th.cb(h) + (long long)r * 32;
          double u = b.U[aidx];
          if (QMAX > 0) {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < QMAX; ++q)
              if (q < I) {
                const double term = __dmul_rn(sd.alpha[I][q], b.K[q][aidx]);
                acc = (q == 0) ? term : __dadd_rn(acc, term);
              }
            if (I > 0) u = __dadd_rn(u, acc);
          }
          own[h][r] = u;
        }
    }
    // ghost cells inside the padding: the slots of nodes N .. N+P-1 carry the state of the
    // wrapped / clamped neighbour (compilers.py:257-264), so the windows below need no
    // special case at the end of the domain.  (The slots may lie in a warp-block beyond the
    // system's last one: they only exist in shared memory.)
    bool ghost = false;
    if (last_tile) {
#pragma unroll
      for (int h = 0; h < GS_G; ++h) {
        const int i0 = (th.chunk0 + h) * M;
        if (i0 + M > g.N && i0 < g.N + P) {
          ghost = true;
#pragma unroll
          for (int m = 0; m < M; ++m) {
            const int j = i0 + m;
            if (j >= g.N && j < g.N + P) own[h][m] = gs_node_value<IT>(cx, g, b, map_node(j, g), sd, irt);
          }
        }
      }
    }
    if (th.active || ghost) {
#pragma unroll
      for (int h = 0; h < GS_G; ++h)
#pragma unroll
        for (int r = 0; r < C; ++r) sS[gs_ss(r, h, th.t, T)] = own[h][r];
    }
    // stage state at the P nodes left and right of the tile
    if (threadIdx.x < 2 * P) {
      const int side = threadIdx.x / P, k = threadIdx.x - side * P;
      const int j = side == 0 ? cx.tile * TN - P + k : (cx.tile + 1) * TN + k;
      // (right of a tile that ends in padding: no real node reads it, nobody publishes it)
      sh.halo[side][k] = (j < g.N + P) ? gs_node_value<IT>(cx, g, b, map_node(j, g), sd, irt) : 0.0;
    }
    // the first chunk of a periodic system starts the recurrences of W and G: F_top
    if (IT == 0 && g.periodic && cx.tiles > 1 && cx.tile == 0 && threadIdx.x >= 2 * P &&
        threadIdx.x < 2 * P + NB * NB)
      sh.ftop[threadIdx.x - 2 * P] = gs_wait(cx, cx.rec.ftop() + (threadIdx.x - 2 * P), 5);
    __syncthreads();
    double rhs_all[GS_G][C];
    Aff mh[GS_G];
#pragma unroll
    for (int h = 0; h < GS_G; ++h) mh[h] = Aff::identity();
    if (th.active) {
#pragma unroll
      for (int h = 0; h < GS_G; ++h) {
        const int i0 = (th.chunk0 + h) * M;
        double win[NF][M + 2 * P];
#pragma unroll
        for (int w = 0; w < M + 2 * P; ++w) {
          const int rel = w - P;
          double v;
          if (rel < 0) {
            if (h > 0) v = own[h > 0 ? h - 1 : 0][M + rel];
            else v = (th.t > 0) ? sS[gs_ss(M + rel, GS_G - 1, th.t - 1, T)] : sh.halo[0][P + rel];
          } else if (rel >= M) {
            if (h < GS_G - 1) v = own[h < GS_G - 1 ? h + 1 : 0][rel - M];
            else v = (th.t < T - 1) ? sS[gs_ss(rel - M, 0, th.t + 1, T)] : sh.halo[1][rel - M];
          } else {
            v = own[h][rel];
          }
          win[0][w] = v;
        }
        RecState rs;
        rs.init();
#pragma unroll
        for (int m = 0; m < M; ++m) {
          const int i = i0 + m;
          double fe[V];
          fe[0] = 0.0;
          {
            TfNodeIn in;
            node_inputs<M>(in, win, m, i, g, b, 0);
            tf_model_F_solver<FD>(cst, in, fe);
          }
          double rhs = __dmul_rn(sd.dt, fe[0]);
#pragma unroll
          for (int q = 0; q < QMAX; ++q)
            if (q < I) rhs = __fma_rn(sd.cfac[I][q], b.K[q][th.cb(h) + (long long)m * 32], rhs);
          rhs = (i < g.N) ? rhs : 0.0;
          rhs_all[h][m] = rhs;
          double coef[BETA];
#pragma unroll
          for (int q = 0; q < BETA; ++q) coef[q] = sL[gs_sl(m, q, h, th.t, T)];
          rs.step(coef, rhs, 1.0);
        }
        rs.to_map(mh[h]);
      }
    }
    GS_STAMP(4 + 6 * I);
    double sv[BETA];
    double ws[NB][BETA], pg[NB][BETA];       // incoming states of the fill recurrences (stage 0)
#pragma unroll
    for (int c = 0; c < NB; ++c)
#pragma unroll
      for (int t = 0; t < BETA; ++t) { ws[c][t] = 0.0; pg[c][t] = 0.0; }
    const bool first_chunk = (cx.tile == 0 && th.t == 0);
    if (IT == 0 && g.periodic) {
      AffB mb[GS_G];
#pragma unroll
      for (int h = 0; h < GS_G; ++h) {
        mb[h] = AffB::identity();
#pragma unroll
        for (int k = 0; k < BETA * BETA; ++k) { mb[h].PhiL()[k] = mh[h].Phi()[k]; mb[h].PhiG()[k] = phiG[h][k]; }
#pragma unroll
        for (int k = 0; k < BETA; ++k) mb[h].cy()[k] = mh[h].c()[k];
      }
      double w0[2][NB][BETA];               // states leaving the system's first chunk
      if (first_chunk && th.active) {
        const double* Ug = b.Uf + th.cb(0) * (BETA + 1) - (long long)(th.cl) * BETA;
#pragma unroll
        for (int c = 0; c < NB; ++c) {
          double sW[BETA], p[BETA];
#pragma unroll
          for (int t = 0; t < BETA; ++t) { sW[t] = 0.0; p[t] = 0.0; }
#pragma unroll
          for (int r = 0; r < C; ++r) {
            double v = gs_fw(b.btab, a, r, c);
#pragma unroll
            for (int q = 0; q < BETA; ++q) v -= sL[gs_sl(r, q, 0, 0, T)] * sW[q];
#pragma unroll
            for (int q = BETA - 1; q > 0; --q) sW[q] = sW[q - 1];
            sW[0] = v;
            const double gv = (gs_fg(sh.ftop, a, r, c) - p[0]) * Ug[(long long)(r * (BETA + 1)) * 32];
#pragma unroll
            for (int t = 0; t < BETA; ++t)
              p[t] = ((t + 1 < BETA) ? p[t + 1 < BETA ? t + 1 : 0] : 0.0) + Ug[(long long)(r * (BETA + 1) + t + 1) * 32] * gv;
          }
#pragma unroll
          for (int t = 0; t < BETA; ++t) {
            w0[0][c][t] = sW[t];
            w0[1][c][t] = p[t];
            gs_post(cx, cx.rec.w0() + c * BETA + t, sW[t]);
            gs_post(cx, cx.rec.w0() + (NB + c) * BETA + t, p[t]);
          }
        }
#pragma unroll
        for (int k = 0; k < BETA * BETA; ++k) {      // the products start behind the first chunk
          mb[0].PhiL()[k] = (k / BETA == k % BETA) ? 1.0 : 0.0;
          mb[0].PhiG()[k] = (k / BETA == k % BETA) ? 1.0 : 0.0;
        }
      }
      AffB mineB = mb[0];
#pragma unroll
      for (int h = 1; h < GS_G; ++h) mineB = AffB::combine(mineB, mb[h]);
      const AffB pre = gs_scan<AffB, false>(mineB, cx, 1);
#pragma unroll
      for (int t = 0; t < BETA; ++t) sv[t] = pre.cy()[t];
      if (first_chunk) {
        if (th.active) {
          // (chunk 0 itself starts from zero; its rows are redone below with the rhs)
        }
      } else if (th.active) {
        bool nz = false;
#pragma unroll
        for (int k = 0; k < BETA * BETA; ++k) nz = nz || (pre.PhiL()[k] != 0.0) || (pre.PhiG()[k] != 0.0);
        if (nz) {
          double w0all[2 * NB * BETA];            // [W | G][NB][BETA]: one round trip
          gs_wait_many<2 * NB * BETA>(cx, cx.rec.w0(), w0all, 13);
#pragma unroll
          for (int c = 0; c < NB; ++c) {
            double vW[BETA], vG[BETA];
#pragma unroll
            for (int t = 0; t < BETA; ++t) {
              vW[t] = w0all[c * BETA + t];
              vG[t] = w0all[(NB + c) * BETA + t];
            }
#pragma unroll
            for (int i = 0; i < BETA; ++i) {
              double sw = 0.0, sg = 0.0;
#pragma unroll
              for (int k = 0; k < BETA; ++k) {
                sw += pre.PhiL()[i * BETA + k] * vW[k];
                sg += pre.PhiG()[i * BETA + k] * vG[k];
              }
              ws[c][i] = sw;
              pg[c][i] = sg;
            }
          }
        }
      }
    } else {
      Aff mine = mh[0];
#pragma unroll
      for (int h = 1; h < GS_G; ++h) mine = Aff::combine(mine, mh[h]);
      const Aff pre = gs_scan<Aff, false>(mine, cx, 1 + 2 * I);
#pragma unroll
      for (int t = 0; t < BETA; ++t) sv[t] = pre.c()[t];
    }
    GS_STAMP(5 + 6 * I);
    if (th.active) {
#pragma unroll
      for (int h = 0; h < GS_G; ++h)
#pragma unroll
        for (int r = 0; r < C; ++r) {
          double v = rhs_all[h][r];
#pragma unroll
          for (int q = 0; q < BETA; ++q) v -= sL[gs_sl(r, q, h, th.t, T)] * sv[q];
#pragma unroll
          for (int q = BETA - 1; q > 0; --q) sv[q] = sv[q - 1];
          sv[0] = v;
          y[h][r] = v;
        }
    }
    // ------------------------------------------------------------ border fill (once per step)
    if (IT == 0) {
      th_alive = false;
      if (g.periodic && th.active) {
        th_alive = first_chunk;
#pragma unroll
        for (int c = 0; c < NB; ++c)
#pragma unroll
          for (int t = 0; t < BETA; ++t) th_alive = th_alive || (ws[c][t] != 0.0) || (pg[c][t] != 0.0);
        if (th_alive) {
          // this thread's rows of W = L^-1 E and G^T = F^T U^-1 with the true incoming states
#pragma unroll 1
          for (int h = 0; h < GS_G; ++h) {
            const int r0 = (th.chunk0 + h) * C;
            const double* Ug = b.Uf + th.cb(h) * (BETA + 1) - (long long)(th.cl + h) * BETA;
#pragma unroll
            for (int r = 0; r < C; ++r) {
              const int gr = r0 + r;
              const double inv = Ug[(long long)(r * (BETA + 1)) * 32];
              double uq[BETA];
#pragma unroll
              for (int t = 0; t < BETA; ++t) uq[t] = Ug[(long long)(r * (BETA + 1) + t + 1) * 32];
              const long long o = (th.cb(h) + (long long)r * 32) * NB - (long long)(th.cl + h) * (NB - 1);
#pragma unroll
              for (int c = 0; c < NB; ++c) {
                double v = gs_fw(b.btab, a, gr, c);
#pragma unroll
                for (int q = 0; q < BETA; ++q) v -= sL[gs_sl(r, q, h, th.t, T)] * ws[c][q];
#pragma unroll
                for (int q = BETA - 1; q > 0; --q) ws[c][q] = ws[c][q - 1];
                ws[c][0] = v;
                const bool onb = gr >= bot0 && gr < g.nhat;      // bottom rows: already hold their local part
                b.Wb[o + (long long)c * 32] = onb ? b.Wb[o + (long long)c * 32] + v : v;
                const double gv = (gs_fg(sh.ftop, a, gr, c) - pg[c][0]) * inv;
#pragma unroll
                for (int t = 0; t < BETA; ++t)
                  pg[c][t] = ((t + 1 < BETA) ? pg[c][t + 1 < BETA ? t + 1 : 0] : 0.0) + uq[t] * gv;
                b.Gb[o + (long long)c * 32] = onb ? b.Gb[o + (long long)c * 32] + gv : gv;
              }
            }
          }
        }
      }
      tile_alive = __syncthreads_or(th_alive ? 1 : 0) != 0;
      if (last_tile) {
        gs_border_finish(g, b, cx, th, a, th_alive, tile_alive);
      } else {
        if (threadIdx.x == 0) gs_post(cx, cx.rec.alive(cx.tile), tile_alive ? 1.0 : 0.0);
        if (tile_alive) {
          gs_gw_partial(g, b, cx, th, th_alive, false);
          if (threadIdx.x < NB * NB) gs_post(cx, cx.rec.spart(cx.tile) + threadIdx.x, sh.red[0][threadIdx.x]);
        }
      }
    }
  }
  // ---------------------------------------------------------------- border solve
  {
    // partial sums of G^T y over the rows of the fill (border_solution_* of the pipeline)
    if (tile_alive) {
      double acc[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) acc[c] = 0.0;
      if (th_alive) {
#pragma unroll
        for (int h = 0; h < GS_G; ++h) {
          const int r0 = (th.chunk0 + h) * C;
#pragma unroll
          for (int r = 0; r < C; ++r)
            if (r0 + r < bot0) {
#pragma unroll
              for (int c = 0; c < NB; ++c) acc[c] += b.Gb[gs_fidx(cx, r0 + r, c, NB)] * y[h][r];
            }
        }
      }
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
      for (int c = 0; c < NB; ++c) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], d);
        if (lane == 0) sh.red[warp][c] = acc[c];
      }
      __syncthreads();
      if (threadIdx.x < NB) {
        double v = 0.0;
        for (int w = 0; w < (T >> 5); ++w) v += sh.red[w][threadIdx.x];
        if (last_tile) sh.red[0][threadIdx.x] = v;
        else gs_post(cx, cx.rec.gpart(I, cx.tile) + threadIdx.x, v);
      }
      __syncthreads();
    } else if (last_tile) {
      if (threadIdx.x < NB) sh.red[0][threadIdx.x] = 0.0;
      __syncthreads();
    }
    // x_b = S^-1 (y_b - G^T y): the last tile owns the bottom rows and the border rows
    if (last_tile) {
      if (th.active) {
#pragma unroll
        for (int h = 0; h < GS_G; ++h) {
          const int r0 = (th.chunk0 + h) * C;
          if (r0 + C > bot0 && r0 < g.nhat + NB) {
#pragma unroll
            for (int r = 0; r < C; ++r) {
              const int k = r0 + r - bot0;
              if (k >= 0 && k < 2 * NB) sh.yb[k] = y[h][r];
            }
          }
        }
      }
      gs_gather_sum<1>(cx, I, NB, 8);             // (starts with a barrier-free poll,
      __syncthreads();                                        //  ends with one)
      if (threadIdx.x == 0) {
        double acc[NB];
        for (int c = 0; c < NB; ++c) acc[c] = sh.red[0][c];
        for (int j = 0; j < NB; ++j) {                        // natural coupling of the last rows
          const int gr = bot0 + j;
          for (int c = 0; c < NB; ++c)
            acc[c] = __fma_rn(tile_alive ? b.Gb[gs_fidx(cx, gr, c, NB)] : sh.gbot[j * NB + c], sh.yb[j], acc[c]);
        }
        double ybv[NB];
        for (int c = 0; c < NB; ++c) ybv[c] = __dsub_rn(sh.yb[NB + c], acc[c]);
        for (int r = 0; r < NB; ++r) {
          double s = 0.0;
          for (int c = 0; c < NB; ++c) s = __fma_rn(sh.sinv[r * NB + c], ybv[c], s);
          sh.xb[r] = s;
          if (sh.nalive > 0) gs_post(cx, cx.rec.xb(I) + r, s);
        }
      }
      __syncthreads();
    } else if (tile_alive) {
      if (threadIdx.x < NB) sh.xb[threadIdx.x] = gs_wait(cx, cx.rec.xb(I) + threadIdx.x, 9);
      __syncthreads();
    }
  }
  GS_STAMP(6 + 6 * I);
  // --------------------------------------------------------------- backward
  // y <- y - W x_b on the rows of the fill, border rows <- x_b
  if (th.active && (th_alive || last_tile)) {
    double xb[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) xb[c] = sh.xb[c];
#pragma unroll
    for (int h = 0; h < GS_G; ++h) {
      const int r0 = (th.chunk0 + h) * C;
      if (th_alive || (r0 + C > bot0 && r0 < g.nhat + NB)) {
#pragma unroll
        for (int r = 0; r < C; ++r) {
          const int gr = r0 + r;
          if (gr < g.nhat && (th_alive || gr >= bot0)) {
#pragma unroll
            for (int c = 0; c < NB; ++c) y[h][r] = __fma_rn(-b.Wb[gs_fidx(cx, gr, c, NB)], xb[c], y[h][r]);
          } else if (gr >= g.nhat && gr < g.nhat + NB) {
#pragma unroll
            for (int c = 0; c < NB; ++c) if (gr - g.nhat == c) y[h][r] = xb[c];
          }
        }
      }
    }
  }
  Aff mine = Aff::identity();
  if (th.active) {
    Aff mh[GS_G];
#pragma unroll
    for (int h = GS_G - 1; h >= 0; --h) {
      const double* Ug = b.Uf + th.cb(h) * (BETA + 1) - (long long)(th.cl + h) * BETA;
      RecState rs;
      rs.init();
#pragma unroll
      for (int r = C - 1; r >= 0; --r) {
        double coef[BETA];
#pragma unroll
        for (int q = 0; q < BETA; ++q) coef[q] = Ug[(long long)(r * (BETA + 1) + q + 1) * 32];
        rs.step(coef, y[h][r], Ug[(long long)(r * (BETA + 1)) * 32]);
      }
      rs.to_map(mh[h]);
    }
    mine = mh[GS_G - 1];                                 // mirrored order: the later chunk first
#pragma unroll
    for (int h = GS_G - 2; h >= 0; --h) mine = Aff::combine(mine, mh[h]);
  }
  GS_STAMP(7 + 6 * I);
  const Aff pre = gs_scan<Aff, true>(mine, cx, 2 + 2 * I);
  GS_STAMP(8 + 6 * I);
  if (th.active) {
    double sv[BETA];
#pragma unroll
    for (int t = 0; t < BETA; ++t) sv[t] = pre.c()[t];
#pragma unroll
    for (int h = GS_G - 1; h >= 0; --h) {
      const double* Ug = b.Uf + th.cb(h) * (BETA + 1) - (long long)(th.cl + h) * BETA;
#pragma unroll
      for (int r = C - 1; r >= 0; --r) {
        double v = y[h][r];
#pragma unroll
        for (int q = 0; q < BETA; ++q) v -= Ug[(long long)(r * (BETA + 1) + q + 1) * 32] * sv[q];
        v *= Ug[(long long)(r * (BETA + 1)) * 32];
#pragma unroll
        for (int q = BETA - 1; q > 0; --q) sv[q] = sv[q - 1];
        sv[0] = v;
        const long long aidx = th.cb(h) + (long long)r * 32;
        double k = v;
        double kprev[QMAX > 0 ? QMAX : 1];
#pragma unroll
        for (int q = 0; q < QMAX; ++q)
          if (q < I) {
            kprev[q] = b.K[q][aidx];
            k = __fma_rn(-sd.cfac[I][q], kprev[q], k);
          }
        if (!LAST) {
          b.K[I][aidx] = k;
        } else {
          double acc = 0.0, accp = 0.0;
#pragma unroll
          for (int q = 0; q <= QMAX; ++q)
            if (q <= I) {
              const double kq = (q < I) ? kprev[q < QMAX ? q : 0] : k;
              const double t = __dmul_rn(sd.b[q], kq);
              acc = (q == 0) ? t : __dadd_rn(acc, t);
              const double tp = __dmul_rn(sd.bp[q], kq);
              accp = (q == 0) ? tp : __dadd_rn(accp, tp);
            }
          const double un = __dadd_rn(b.U[aidx], acc);
          b.Un[aidx] = un;
          if (sd.has_pred) {
            const double e = fabs(__dsub_rn(un, __dadd_rn(un, accp)));
            emax = (e > emax || e != e) ? e : emax;
          }
        }
      }
    }
  }
  if (!LAST && cx.tiles > 1 && th.active) {
    // stage state of the next stage at the tile edges (and at the last P real nodes) for the
    // neighbouring tiles; out of the hot loop: only the threads at the edges take this path
#pragma unroll
    for (int h = 0; h < GS_G; ++h) {
      const int i0 = (th.chunk0 + h) * M;
      const bool first = th.t == 0 && h == 0, lastc = th.t == T - 1 && h == GS_G - 1;
      const bool tail = last_tile && i0 + M > g.N - P && i0 < g.N;
      if (first || lastc || tail) {
#pragma unroll 1
        for (int r = 0; r < C; ++r) {
          const int i = i0 + r;
          const int loc = i - cx.tile * TN;
          const bool e0 = loc < P, e1 = loc >= TN - P, e2 = last_tile && i >= g.N - P && i < g.N;
          if (e0 || e1 || e2) {
            const long long aidx = th.cb(h) + (long long)r * 32;
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q <= QMAX; ++q)
              if (q <= I) {
                const double term = __dmul_rn(sd.alpha[I + 1][q], b.K[q][aidx]);
                acc = (q == 0) ? term : __dadd_rn(acc, term);
              }
            const double nv = __dadd_rn(b.U[aidx], acc);
            if (e0) gs_post(cx, cx.rec.halo(I + 1, cx.tile, 0) + loc, nv);
            if (e1) gs_post(cx, cx.rec.halo(I + 1, cx.tile, 1) + (loc - (TN - P)), nv);
            if (e2) gs_post(cx, cx.rec.halo(I + 1, cx.tile, 2) + (i - (g.N - P)), nv);
          }
        }
      }
    }
  }
  if (GS_MULTI && LAST && cx.nranks > 1 && th.active) {
    // the state of the NEXT step at the tile edges: another rank's first phases read it as
    // words (there is no kernel boundary between ranks); tagged with the next epoch
#pragma unroll
    for (int h = 0; h < GS_G; ++h) {
      const int i0 = (th.chunk0 + h) * M;
      const bool first = th.t == 0 && h == 0, lastc = th.t == T - 1 && h == GS_G - 1;
      const bool tail = last_tile && i0 + M > g.N - P && i0 < g.N;
      if (first || lastc || tail) {
#pragma unroll 1
        for (int r = 0; r < C; ++r) {
          const int i = i0 + r;
          const int loc = i - cx.tile * TN;
          const bool e0 = loc < GS_HW, e1 = loc >= TN - P, e2 = last_tile && i >= g.N - P && i < g.N;
          if (e0 || e1 || e2) {
            const double nv = b.Un[th.cb(h) + (long long)r * 32];
            const long long nt = (cx.tag >= (1 << 28)) ? 1 : cx.tag + 1;
            if (e0) gs_post_tag(cx, cx.rec.halo(0, cx.tile, 0) + loc, nv, nt);
            if (e1) gs_post_tag(cx, cx.rec.halo(0, cx.tile, 1) + (loc - (TN - P)), nv, nt);
            if (e2) gs_post_tag(cx, cx.rec.halo(0, cx.tile, 2) + (i - (g.N - P)), nv, nt);
          }
        }
      }
    }
  }
  if (!LAST) __syncthreads();       // k_I of the tile is in place before the next stage reads it
  GS_STAMP(9 + 6 * I);
}

}  // namespace GS_NS
}  // namespace tfk

#ifndef TF_GS_ONCE
#ifndef TF_GS_MINB
#define TF_GS_MINB 1
#endif
// (read by the host: chunks per thread, max threads per CTA)
extern "C" __device__ int tf_gs_cfg[2] = {TF_GS_G, TF_GS_NT};
#endif
// `g` describes the WHOLE grid (N, nhat, nblk of all ranks together); mr this rank's slab
extern "C" __global__ void __launch_bounds__(tfk::GS_NT, TF_GS_MINB) TF_GS_KNAME(tfk::Geom g, tfk::Buf b,
                                                                        TfStepDesc sd, TfGsMulti mr) {
  using namespace tfk;
  using namespace tfk::GS_NS;
  extern __shared__ __align__(128) double dsm_gs[];
  __shared__ GsShared sh;
  const int T = blockDim.x;
  double* sL = dsm_gs;                                   // [C][BETA][G][T]
  double* sS = dsm_gs + (size_t)T * GS_G * C * BETA;     // [C][G][T]
  GsCtx cx;
  cx.rec.base = (LbWord*)b.gs;
  cx.rec.tiles = (int)gridDim.x;
  cx.rec.nranks = mr.nranks;
  cx.rec.first = (GS_MULTI ? mr.rank : 0) * (int)gridDim.x;
#pragma unroll
  for (int r = 0; r < TF_GS_MAXRANKS; ++r) cx.rec.bases[r] = (LbWord*)mr.bases[r];
  cx.sh = &sh;
  cx.status = b.status;
  cx.rank = GS_MULTI ? mr.rank : 0;
  cx.nranks = GS_MULTI ? mr.nranks : 1;
  cx.ltile = (int)blockIdx.x;
  cx.tile = cx.rank * (int)gridDim.x + (int)blockIdx.x;
  cx.tiles = GS_MULTI ? mr.tiles_total : (int)gridDim.x;
  cx.T = T;
  cx.nodes_local = (int)gridDim.x * T * GS_G * M;
  cx.node_off = cx.rank * cx.nodes_local;
  if (threadIdx.x == 0) {
    sh.epoch = ld_flag(b.ctl + 0);
    sh.abort = 0;
    sh.nalive = 0;
    tf_gs_stuck[blockIdx.x & 1023] = 0;
    if (blockIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tf_gs_when[0]));
  }
  for (int k = threadIdx.x; k < NC2; k += T) sh.cst[k] = b.cst[k];
  __syncthreads();
  cx.tag = sh.epoch;
  GsThread th;
  th.t = (int)threadIdx.x;
  th.T = T;
  th.chunk0 = (cx.tile * T + th.t) * GS_G;
  {
    const int lchunk0 = (cx.ltile * T + th.t) * GS_G;
    th.blk = lchunk0 >> 5;
    th.cl = lchunk0 & 31;
  }
  th.active = th.blk < mr.nblk_local;
  const double a = sd.a;
  GS_STAMP(0);
  double emax = 0.0;
  // (tiles behind the last live one -- several GPUs, the grid does not fill the last slab --
  //  hold padding only: nothing to do but the bookkeeping at the end)
  if (cx.tile < cx.tiles) {
  double phiG[GS_G][BETA * BETA];
  gs_factor(g, b, cx, th, sd, a, sL, phiG);
  if (cx.tile == cx.tiles - 1) gs_border_bottom(g, b, cx, a, sL);
  bool th_alive = false, tile_alive = false;
#ifdef TF_GS_RTSTAGE
  if (sd.s == 1) {
    gs_stage<0, true>(g, b, cx, th, sd, a, sL, sS, phiG, th_alive, tile_alive, emax);
  } else {
    gs_stage<0, false>(g, b, cx, th, sd, a, sL, sS, phiG, th_alive, tile_alive, emax);
#pragma unroll 1
    for (int i = 1; i < sd.s; ++i)
      gs_stage<-1, false>(g, b, cx, th, sd, a, sL, sS, phiG, th_alive, tile_alive, emax, i, i == sd.s - 1);
  }
#else
  switch (sd.s) {
    case 1:
      gs_stage<0, true>(g, b, cx, th, sd, a, sL, sS, phiG, th_alive, tile_alive, emax);
      break;
    case 2:
      gs_stage<0, false>(g, b, cx, th, sd, a, sL, sS, phiG, th_alive, tile_alive, emax);
      gs_stage<1, true>(g, b, cx, th, sd, a, sL, sS, phiG, th_alive, tile_alive, emax);
      break;
    default:
      gs_stage<0, false>(g, b, cx, th, sd, a, sL, sS, phiG, th_alive, tile_alive, emax);
      gs_stage<1, false>(g, b, cx, th, sd, a, sL, sS, phiG, th_alive, tile_alive, emax);
      gs_stage<2, true>(g, b, cx, th, sd, a, sL, sS, phiG, th_alive, tile_alive, emax);
      break;
  }
#endif
  }
  // error estimate: per-tile maximum, reduced by the last CTA to finish
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const double o = __shfl_xor_sync(0xffffffffu, emax, d);
    emax = (o > emax || o != o) ? o : emax;
  }
  if (lane == 0) sh.err[warp] = emax;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (T >> 5); ++w) emax = (sh.err[w] > emax || sh.err[w] != sh.err[w]) ? sh.err[w] : emax;
    double* errt = cx.rec.errt();
    errt[cx.ltile] = emax;
    __threadfence();
    const unsigned d = atomicAdd((unsigned*)(b.ctl + 2), 1u);
    if (d == gridDim.x - 1) {
      __threadfence();
      double e = 0.0;
      for (int t = 0; t < (int)gridDim.x; ++t) {
        const double v = __ldcg(errt + t);
        e = (v > e || v != v) ? v : e;
      }
      b.err[0] = e;
      b.ctl[2] = 0;
      b.ctl[0] = (sh.epoch >= (1 << 28)) ? 1 : sh.epoch + 1;
    }
  }
}

#if TF_GS_MULTI
// Several GPUs, after an upload: the stage-0 words of this rank's tile edges from U (every later
// step gets them from the last stage of the step before)
extern "C" __global__ void tf_k_gs_seed(tfk::Geom g, tfk::Buf b, TfGsMulti mr, int tiles_local, int T) {
  using namespace tfk;
  using namespace tfk::GS_NS;
  GsRec rec;
  rec.base = (LbWord*)b.gs;
  rec.tiles = tiles_local;
  rec.nranks = 1;                                 // own area, local tile indices
  rec.first = 0;
  const long long tag = ld_flag(b.ctl + 0);
  const int TN = T * GS_G * M;
  const int node_off = mr.rank * tiles_local * TN;
  const int k = threadIdx.x;
  if (k >= GS_HW) return;
  for (int lt = blockIdx.x; lt < tiles_local; lt += gridDim.x) {
    const int first = node_off + lt * TN;
    if (first + k < g.N) st_word_sys(rec.halo(0, lt, 0) + k, b.U[vidx(lt * TN + k, 0)], tag);
    if (k >= P) continue;
    if (first + TN - P + k < g.N)
      st_word_sys(rec.halo(0, lt, 1) + k, b.U[vidx(lt * TN + TN - P + k, 0)], tag);
    const int last = g.N - P + k;                 // the last P real nodes of the whole grid
    if (last >= first && last < first + TN) st_word_sys(rec.halo(0, lt, 2) + k, b.U[vidx(last - node_off, 0)], tag);
  }
}
#endif

#undef GS_NS
#undef TF_GS_KNAME
#ifndef TF_GS_ONCE
#define TF_GS_ONCE 1
#endif
#endif  // V == 1 && no helper fields
