/* triflow_b200 -- C ABI of the B200-native implicit method-of-lines hot path.
 *
 * The reference (celliern/triflow) is pure Python and has no FFI; these entry
 * points are what its plugin points would bind for this path.  Each function
 * cites the reference interface it replaces (paths relative to the reference
 * repository).  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions: every call returns 0 on success or a TF_E* code; the message of
 * the last failure on the calling thread is tf_last_error().  Handles are
 * opaque.  Host buffers are caller-owned plain double arrays; device buffers are
 * library-owned.  One CUDA stream per context; calls on one context are not
 * thread-safe, distinct contexts are independent.
 */
#ifndef TRIFLOW_B200_H
#define TRIFLOW_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TF_OK 0
#define TF_EINVAL 1   /* bad argument                                            */
#define TF_ECUDA 2    /* CUDA runtime / driver error (see tf_last_error)         */
#define TF_EMAXITER 3 /* adaptive step: internal iterations above max_iter       */
#define TF_EDTMIN 4   /* adaptive step: internal time step below dt_min          */
#define TF_ESINGULAR 5 /* zero / non-finite pivot in the banded factorisation    */

typedef struct tf_ctx_s* tf_ctx_t;
typedef struct tf_model_s* tf_model_t;
typedef struct tf_state_s* tf_state_t;
typedef struct tf_scheme_s* tf_scheme_t;
typedef struct tf_ring_s* tf_ring_t;

/* Static description of a lowered model (produced by triflow_b200/codegen.py from
 * the attributes a compiler plugin receives: core/model.py:244-291). */
typedef struct {
  int nvar;        /* model._nvar                                   */
  int nhelp;       /* len(model._help_funcs)                        */
  int half_width;  /* -model._bounds[0]                             */
  int nnz;         /* len(model._J_sparse_array)                    */
  int n_const;     /* host-evaluated uniform sub-expressions        */
  int n_nodepar;   /* parameters passed as per-node arrays          */
  int uses_x;      /* expressions reference the coordinate x        */
  int chunk_nodes; /* nodes per thread the cubin was built for      */
  int warps;       /* unused (warps per CTA are chosen at launch)   */
} tf_model_desc;

const char* tf_last_error(void);

/* context = device + stream */
int tf_ctx_create(int device, tf_ctx_t* out);
int tf_ctx_destroy(tf_ctx_t ctx);
int tf_ctx_sync(tf_ctx_t ctx);

/* Asynchronous host-buffer mode (off by default).  When enabled, tf_state_upload,
 * tf_state_download and tf_scheme_step enqueue their work on the context's stream and
 * return without waiting: host buffers must be pinned (tf_host_alloc) and must not be
 * touched until tf_ctx_sync; failures of the factorisation are then reported by
 * tf_state_status instead of the return code.  Several contexts on one device overlap
 * the copies of one group of systems with the stepping of another (the reference copies
 * nothing: its fields live in host memory, core/fields.py:146-183). */
int tf_ctx_set_async(tf_ctx_t ctx, int enable);

/* pinned host memory for upload / download buffers */
int tf_host_alloc(size_t nbytes, void** out);
/* write-combined variant for buffers the host only WRITES (upload sources): the device reads
 * them over PCIe without snooping the CPU caches; reading them from the CPU is very slow */
int tf_host_alloc_wc(size_t nbytes, void** out);
int tf_host_free(void* p);

/* Replaces numpy_compiler(model) -> (compute_F, compute_J)
 * (core/compilers.py:181-224): loads the sm_100a cubin generated for the model. */
int tf_model_load(tf_ctx_t ctx, const void* cubin, size_t nbytes, const tf_model_desc* desc,
                  tf_model_t* out);
int tf_model_unload(tf_model_t model);
/* Introspection: copy a __device__ symbol of the model's cubin to the host (kernels built
 * with -DTF_TRACE keep per-CTA phase time stamps in `tf_trace`; tools/trace_tiles.py). */
int tf_model_read_symbol(tf_model_t model, const char* name, void* dst, size_t nbytes);

/* Device-resident fields of `batch` independent systems of n_nodes nodes.
 * Replaces the per-call padding / view construction of init_computation_numpy
 * (core/compilers.py:227-278). */
int tf_state_create(tf_ctx_t ctx, tf_model_t model, int n_nodes, int batch, int periodic,
                    tf_state_t* out);
int tf_state_destroy(tf_state_t st);

/* Upload host data; any pointer may be NULL (left unchanged).
 *   x        [n_nodes]                       fields['x']
 *   u        [batch][n_nodes*nvar]           BaseFields.uflat layout (core/fields.py:146-159)
 *   helpers  [batch][nhelp][n_nodes]
 *   nodepars [batch][n_nodepar][n_nodes]     array-valued parameters (core/routines.py:40)
 *   consts   [batch][2*max(1,n_const)]       Lowered.uniform_table(...)                  */
int tf_state_upload(tf_state_t st, const double* x, const double* u, const double* helpers,
                    const double* nodepars, const double* consts);
/* u [batch][n_nodes*nvar]: inverse of upload, BaseFields.fill order (core/fields.py:173-183) */
int tf_state_download(tf_state_t st, double* u);
/* same layout written to DEVICE memory (dev_u: batch*n_nodes*nvar doubles on the state's GPU),
 * e.g. the send buffer of the final gather of a sharded ensemble: no host round trip */
int tf_state_download_device(tf_state_t st, double* dev_u);

/* F_Routine.__call__ -> compute_F_numpy (core/routines.py:37-45, compilers.py:281-289):
 * out [batch][n_nodes*nvar]. */
int tf_eval_F(tf_state_t st, double* out);
/* J_Routine.__call__ -> compute_J_numpy (core/routines.py:82-91, compilers.py:292-301):
 * out [batch][n_nodes][nnz], the per-node values of model._J_sparse_array; the
 * caller places them in CSC with the reference's index rule (compilers.py:303-331). */
int tf_eval_J(tf_state_t st, double* out);

/* Rosenbrock-Wanner tableau (ROW_general.__init__, core/schemes.py:81-99).  Theta(theta)
 * (core/schemes.py:518-559) is the one-stage tableau gamma=[[theta]], b=[1].
 * alpha, gamma: [s][s] row-major; b: [s]; b_pred: [s] or NULL. */
int tf_scheme_create(tf_ctx_t ctx, int s, const double* alpha, const double* gamma,
                     const double* b, const double* b_pred, tf_scheme_t* out);
int tf_scheme_destroy(tf_scheme_t sc);

/* Declarative Dirichlet hook: U[var][0] = left, U[var][-1] = right, applied where the
 * reference calls hook(t, fields, pars) (core/schemes.py:139,145,224,549,558). */
int tf_hook_set_dirichlet(tf_state_t st, int var, int has_left, double left, int has_right,
                          double right);
int tf_hook_clear(tf_state_t st);

/* n_steps fixed steps of ROW_general._fixed_step + post-hook (core/schemes.py:137-174)
 * on every system.  err_out [batch] (or NULL): ||U_new - U_pred||_inf of the last step
 * (NaN when the tableau has no b_pred). */
int tf_scheme_step(tf_state_t st, tf_scheme_t sc, double dt, int n_steps, double* err_out);

/* One call of ROW_general._variable_step (core/schemes.py:176-238) on a single system
 * (batch == 1): embedded-error step-size control up to t + dt.
 *   internal_dt  in/out, < 0 means "None" (first call starts at 1e-6)
 *   max_iter / dt_min  <= 0 means "None"
 *   n_fixed_steps out: number of _fixed_step evaluations performed            */
int tf_scheme_advance(tf_state_t st, tf_scheme_t sc, double t, double dt, double tol,
                      double safety_factor, int max_iter, double dt_min, int recompute_target,
                      double* internal_dt, int* n_fixed_steps, double* last_err);

/* ROW_general._variable_step (core/schemes.py:176-238) run independently for every member
 * of an ensemble (each system keeps its own internal step size; systems must fit one CTA
 * tile, i.e. n_nodes <= 16*32*chunk_nodes).  All arrays have `batch` entries.
 *   internal_dt  in/out, < 0 means "None";  n_fixed_steps, fail (0 or TF_EMAXITER /
 *   TF_EDTMIN per member) out.  Returns TF_EMAXITER / TF_EDTMIN if any member failed. */
int tf_ensemble_advance(tf_state_t st, tf_scheme_t sc, double t, double dt, double tol,
                        double safety_factor, int max_iter, double dt_min, double* internal_dt,
                        int* n_fixed_steps, int* fail);

/* Replaces schemes.time_stepping (core/schemes.py:33-66), the Richardson controller that
 * Simulation wraps around EVERY scheme by default (core/simulation.py:190-197), for every
 * member of an ensemble, on the device: one coarse scheme call over m*dt_ against ten fine
 * calls over dt_, err = max_var ||coarse - fine||_2 / (m^2 - 1), dt_ <- sqrt(dt^2 tol / err),
 * rejected (and repeated from the state reached, as the reference does) while
 * dt_ < dt / reject_factor.  inner_adaptive != 0: the wrapped scheme runs its own
 * embedded-error controller in every call (the reference's default double wrapping of
 * ROS3PRw / ROS3PRL / RODASPR; in_* are that controller's arguments, inner_dt its per-member
 * state); 0: every call is one fixed step (ROS2, Theta, time_stepping=False).
 *   outer_dt  in/out per member, <= 0 means "None";  n_calls: scheme calls, n_fixed_steps:
 *   fixed steps, fail: 0 or TF_E* per member (out, may be null). */
int tf_ensemble_richardson(tf_state_t st, tf_scheme_t sc, int inner_adaptive, double t, double dt,
                           double tol, int m, double reject_factor, double in_tol,
                           double in_safety_factor, int in_max_iter, double in_dt_min,
                           double* outer_dt, double* inner_dt, int* n_calls, int* n_fixed_steps,
                           int* fail);

/* Optional: keep the factorisation across steps while gamma*dt is unchanged.  Only valid
 * for models whose Jacobian does not depend on the state (linear models with uniform
 * parameters; the reference rebuilds and refactorises every step regardless,
 * core/schemes.py:146-149).  Off by default; uploading new constants invalidates it. */
int tf_state_set_factor_reuse(tf_state_t st, int enable);

/* Which kernels run ROW_general._fixed_step / Theta.__call__ (core/schemes.py:142-174,548-559);
 * every choice runs the same algorithm and agrees with the others to rounding.
 *   1 (default) automatic:
 *       - system-resident kernel when a whole system fits one CTA (tridiagonal scalar model,
 *         non-periodic, N <= 4096, (stages + 2) vectors within 200 KB of shared memory): one
 *         launch per step, U and the stage vectors in shared memory, the factor in registers;
 *       - grid-resident kernel for one long grid where it pays (scalar model without helper
 *         fields, stages <= 3, pentadiagonal or wider, N >= 786432): one cooperative launch
 *         per step, one resident tile per CTA, tiles coupled through tagged 16-byte words;
 *       - else the per-kernel pipeline (factor, border fill, 2 sweeps per stage).
 *   0 the per-kernel pipeline only
 *   2 the run-time-stage variant of the system-resident kernel for every tableau
 *   3 the grid-resident kernel wherever it applies (any N) */
int tf_state_set_fusion(tf_state_t st, int enable);

/* status bits per system: bit0 bad pivot, bit1 singular border block, bit2 a tile of the
 * grid-resident step gave up waiting for a neighbour (bits 8.. say where) */
int tf_state_status(tf_state_t st, int* status);
/* number of kernels launched on this state's context since creation */
long long tf_ctx_launch_count(tf_ctx_t ctx);
/* CUDA-event timing on the context's stream: start/stop bracket, elapsed ms */
int tf_ctx_timer_start(tf_ctx_t ctx);
int tf_ctx_timer_stop(tf_ctx_t ctx, float* ms);
/* Output ring: the device-resident output path behind Simulation.stream / the container's
 * buffered writes (core/simulation.py:244-253, plugins/container.py:99-137).  tf_ring_push
 * snapshots the current state (uflat layout, core/fields.py:146-159) into the next of `slots`
 * pinned host buffers with an asynchronous copy on the ring's own stream and returns at once:
 * stepping continues while the copy runs.  tf_ring_pop hands out the oldest snapshot (block:
 * wait for its copy; else *data stays NULL if it is not there yet); tf_ring_release frees its
 * slot.  push fails with TF_EINVAL when all slots are in flight.  One producer thread (the one
 * stepping the state) and one consumer thread may use a ring concurrently. */
int tf_ring_create(tf_state_t st, int slots, tf_ring_t* out);
int tf_ring_destroy(tf_ring_t ring);
int tf_ring_push(tf_ring_t ring, double t);
int tf_ring_pop(tf_ring_t ring, int block, const double** data, double* t);
int tf_ring_release(tf_ring_t ring);

/* ---- One grid over several GPUs.  The reference has no analogue: it grows N only
 * through sparse storage on one host (source_doc/source/user_guide.rst:183-187); the solve that
 * is split here is `factorized(A)` / `luf(b)` of core/schemes.py:149,157 and the stencil
 * evaluation of core/compilers.py:281-332 at the slab edges.
 * Every rank (one GPU each) owns a slab of consecutive nodes, a whole number of tiles of the
 * grid-resident step kernel; rank r of `nranks` creates its slab of the N-node grid with
 * tf_state_create_slab, uploads / downloads ITS nodes only (tf_state_upload / _download with
 * n_local values, node_off .. node_off + n_local - 1 of the grid), publishes its record area
 * (tf_state_slab_export: a 64-byte CUDA IPC handle) and maps everybody else's
 * (tf_state_slab_attach: handles of all ranks in rank order).  tf_scheme_step then runs the
 * SAME step kernel on every GPU at once: halo values, scan records and the periodic border
 * block cross the slab boundaries as tagged 16-byte words read straight from the peer GPU's
 * memory over NVLink -- no collective call and no kernel boundary inside a step.  All ranks
 * must call tf_scheme_step with the same arguments; the error estimate returned is the rank's
 * own (reduce with max over ranks).  Scalar models with uniform parameters, tableaux of at
 * most 3 stages, fixed steps; tf_hook_set_dirichlet on non-periodic grids (every rank sets the
 * same hook, the ranks that own the ends apply it).  tf_state_slab_attach_local is the single-process
 * form: the states of all ranks (rank order) live in the calling process. */
int tf_state_create_slab(tf_ctx_t ctx, tf_model_t m, int N, int periodic, int rank, int nranks,
                         tf_state_t* out);
int tf_state_slab_info(tf_state_t st, int* node_off, int* n_local, int* tiles_local, int* tiles_total);
int tf_state_slab_export(tf_state_t st, void* handle64);
int tf_state_slab_attach(tf_state_t st, const void* handles);
int tf_state_slab_attach_local(tf_state_t st, const tf_state_t* all);

/* measured fp64 throughput of the device (DFMA per second, 8 independent chains per thread):
 * the denominator for kernels that are bound by the fp64 pipe rather than by HBM */
int tf_ctx_fp64_peak(tf_ctx_t ctx, double* dfma_per_s);
/* per-kernel-family accumulated device time (ms) since the last reset; names out */
int tf_ctx_profile(tf_ctx_t ctx, int enable);
int tf_ctx_profile_read(tf_ctx_t ctx, int family, float* ms, long long* launches);

#ifdef __cplusplus
}
#endif
#endif
