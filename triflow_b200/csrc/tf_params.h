// Kernel parameter blocks shared by the model-specialised kernels
// (tf_kernels.cuh, compiled per model to a cubin) and the host driver
// (tf_host.cu, compiled once into libtriflow_b200.so).  Plain C structs.
#pragma once

#define TF_MAXS 6 /* max Rosenbrock stages (RODASPR) */

struct TfGeom {
  int N;        /* real nodes per system */
  int nblk;     /* warp-blocks (32 chunks of M nodes) per system */
  int tiles;    /* CTAs per system */
  int batch;    /* systems */
  int periodic;
  int nhat;     /* interior unknowns (N - P) * V */
};

struct TfBuf {
  double* U;            /* state                     [batch][nblk*C*32] */
  double* Un;           /* next state (ping-pong) */
  double* K[TF_MAXS];   /* stage vectors */
  double* Y;            /* forward-substitution result */
  double* H;            /* helper fields             [batch][NH][nblk*M*32] */
  double* NP;           /* per-node parameters       [batch][NNODEPAR][nblk*M*32] */
  double* X;            /* node coordinates          [nblk*M*32] */
  double* cst;          /* uniform constants         [batch][2*max(1,NCONST)] */
  double* Lf;           /* L multipliers             [batch][nblk*C*32*BETA] */
  double* Uf;           /* U rows, pivot inverted    [batch][nblk*C*32*(BETA+1)] */
  double* Wb;           /* border fill column L^-1 E [batch][nblk*C*32*NB] */
  double* Gb;           /* border fill row F^T U^-1  [batch][nblk*C*32*NB] */
  double* btab;         /* raw J of border couplings [batch][5][NB][NB] */
  double* Sinv;         /* inverse border Schur complement [batch][NB][NB] */
  double* xb;           /* border solution + epoch flag [batch][NB+1] */
  int* lead;            /* [batch][2] leading rows of W / G that may be non-zero */
  int* status;          /* [batch] bit0: bad pivot, bit1: singular border block */
  double* err;          /* [batch] embedded error estimate of the last step */
  int* flags;           /* [0] ticket counter of the chained launches */
  int* ctl;             /* device-side launch bookkeeping: [0] epoch, [1] ticket base, [2] CTAs done */
  /* per-member adaptive stepping of ensembles (NULL: one dt for all systems) */
  const double* dtsys;  /* [batch] time step of this attempt */
  const double* asys;   /* [batch] gamma_ii * dt of this attempt */
  const int* active;    /* [batch] 0: member finished, its CTAs return immediately */
  double* lbagg;        /* [batch*tiles][KMAX] tile aggregates, 16-byte (value, tag) words */
  double* lbinc;        /* [batch*tiles][KMAX] inclusive prefixes, same format */
  double* gpart;        /* [batch*fwd_tiles][NB] per-tile partial G^T y of the last fwd */
  void* gs;             /* record area of the grid-resident step (tf_gridstep.cuh), 16-byte words */
};

struct TfStage {
  int nprev;               /* number of previous stage vectors used */
  int istage;              /* index of the stage vector written */
  double alpha[TF_MAXS];   /* U_i = U + sum alpha_j k_j */
  double cfac[TF_MAXS];    /* gamma_ij / gamma_ii */
  double dt;
  int use_partials;        /* border fill known when the fwd sweep runs: fwd leaves partial G^T y */
  int fwd_tiles;           /* tiling of the fwd launch (indexing of gpart) */
  int fwd_tile_rows;
  int is_last;             /* bwd of the last stage: write U_new and the error estimate */
  int has_pred;
  double b[TF_MAXS];
  double bp[TF_MAXS];
};


/* One grid over several GPUs (tf_gridstep.cuh): every rank owns a slab of whole tiles. */
#define TF_GS_MAXRANKS 8
struct TfGsMulti {
  void* bases[TF_GS_MAXRANKS];   /* record areas of all ranks (own + IPC-mapped peers) */
  int rank, nranks;
  int nblk_local;                /* warp-blocks that hold data on this rank */
  int tiles_total;               /* live tiles of the whole grid (<= nranks x tiles per rank; the
                                    tiles behind them hold padding only and do nothing) */
};

/* Whole-step descriptor of the system-resident kernel (tf_sysstep.cuh): every stage of
   one ROW_general._fixed_step / Theta step (reference core/schemes.py:142-174,548-559). */
struct TfStepDesc {
  int s;                            /* stages */
  int has_pred;
  double dt;
  double a;                         /* gamma_ii * dt */
  double alpha[TF_MAXS][TF_MAXS];   /* U_i = U + sum_{j<i} alpha[i][j] k_j */
  double cfac[TF_MAXS][TF_MAXS];    /* gamma_ij / gamma_ii */
  double b[TF_MAXS];
  double bp[TF_MAXS];
};
