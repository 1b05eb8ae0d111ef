/*
 * hydrob200.h — C-ABI of libhydrob200.so, the B200 (sm_100a) forward-run engine for
 * HydroModels.jl bucket models.
 *
 * This is the drop-in boundary of SURVEY.md §8(b).  The reference (pure Julia) has no FFI for
 * this path; these entry points are what a Julia `ccall` binding for
 *     (model::HydroModel)(input, params, config; initstates)      /root/reference/src/model.jl:241-309
 * and the component functors it dispatches to
 *     HydroBucket   /root/reference/src/bucket.jl:169-256   (hydrosolve: src/solve.jl:31-55)
 *     HydroFlux     /root/reference/src/flux.jl:157-203
 *     UnitHydrograph/root/reference/src/uh.jl:187-274
 *     HydroRoute    /root/reference/src/route.jl:336-404
 * binds (INTEGRATION.md shows the stub).  Plain C: pointers and sizes only, every function
 * returns an int status (0 = ok), no exceptions cross the boundary, `hb_last_error()` returns
 * a thread-local message the host turns into the reference's exceptions.
 *
 * Division of labour: the host language lowers the model's flux/state expressions to a CUDA C
 * *body* (sections PROLOGUE / STEP / EPILOGUE); the library splices
 * the body into its hand-written kernel templates, compiles them with NVRTC for sm_100a,
 * caches the cubin and launches it.  There is no CPU fallback: without a CUDA device every
 * compute entry point fails with HB_ERR_NO_DEVICE.
 *
 * Array layouts are the reference's Julia column-major layouts, i.e. in C order:
 *     input   [T][N][V]      (Julia  input[V, N, T];   src/bucket.jl:207-214)
 *     output  [T][N][R]      (Julia  out[rows, N, T];  src/bucket.jl:243, src/model.jl:308)
 *     states  [S][N]         (flat, state-major / node-fastest; src/model.jl:116-119)
 *     params  flat values, one (offset, mode) per parameter name (src/utils.jl:174-201)
 * All floating point is IEEE double (the reference's promote_type, src/solve.jl:42).
 */
#ifndef HYDROB200_H
#define HYDROB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_ABI_VERSION 1

/* status codes */
enum {
    HB_OK = 0,
    HB_ERR_INVALID = 1,    /* bad argument / descriptor (DimensionMismatch, ArgumentError)   */
    HB_ERR_NO_DEVICE = 2,  /* no CUDA device or driver: there is no CPU fallback             */
    HB_ERR_COMPILE = 3,    /* NVRTC rejected the generated kernel (log in hb_last_error)     */
    HB_ERR_CUDA = 4,       /* CUDA runtime / driver error                                    */
    HB_ERR_NOMEM = 5
};

/* what the stepper writes */
enum {
    HB_OUT_ROWS = 0,       /* out[T][N][R]: the rows the body emits (all rows = reference result) */
    HB_OUT_OBJECTIVE = 1   /* objective[N]: fused metric of the body's simulated row vs target   */
};

/* metrics of ext/HydroModelsOptimizationExt/HydroModelsOptimizationExt.jl:34-68 (value the
 * optimiser minimises: -KGE, -NSE, -LogKGE, MSE) */
enum { HB_METRIC_MSE = 0, HB_METRIC_NSE = 1, HB_METRIC_KGE = 2, HB_METRIC_LOGKGE = 3,
       HB_METRIC_SSE = 4 /* sum of squared errors: the loss of the gradient example, README.md:161-187 */ };

/* parameter addressing modes: value of parameter j for node n is
 *   mode 0: params[offset_j]                 scalar (2D single-node call)
 *   mode 1: params[offset_j + n]             one value per node (htypes == 1:N)
 *   mode 2+k: params[offset_j + htypes[k][n]] gather through 0-based htypes set k           */
enum { HB_PARAM_SCALAR = 0, HB_PARAM_PER_NODE = 1, HB_PARAM_GATHER = 2 };

#define HB_MAX_PARAMS 96
#define HB_MAX_UH 8
#define HB_MAX_TANGENTS 16

typedef struct hb_model_s *hb_model_t;

/* Static description of one fused per-node stepper kernel (a HydroModel made of buckets,
 * fluxes, neural fluxes and unit hydrographs).  Everything here is compile-time for the
 * generated kernel. */
typedef struct {
    int32_t struct_size;         /* sizeof(hb_stepper_spec) */
    int32_t n_inputs;            /* V: forcing rows read per node-step                        */
    int32_t n_rows;              /* R: rows written per node-step (HB_OUT_ROWS)               */
    int32_t n_states;            /* S: state variables over all buckets                       */
    int32_t n_params;            /* parameter names used by the body                          */
    int32_t n_uh;                /* unit hydrographs fused into the stepper                   */
    int32_t uh_kmax[HB_MAX_UH];  /* ordinates per UH (weights are zero-padded to this length) */
    int32_t nn_words;            /* doubles of MLP parameters staged in shared memory         */
    int32_t output_mode;         /* HB_OUT_ROWS | HB_OUT_OBJECTIVE                            */
    int32_t metric;              /* HB_METRIC_* (objective mode)                              */
    int32_t shared_forcing;      /* 1: input is [T][1][V], shared by all nodes (ensembles)    */
    int32_t block_threads;       /* 0 = default (128)                                         */
    int32_t stages;              /* forcing pipeline depth in time steps, 0 = default (4)     */
    int32_t max_registers;       /* 0 = default (128); register budget per thread -> __launch_bounds__ min blocks */
    int32_t precision;           /* 0 = Float64 (reference arithmetic), 1 = opt-in Float32: `input` and `out` are float
                                    arrays and the body computes in float; everything else stays double */
    int32_t nn_tensor;           /* 1: the body contracts its dense layers on the tensor cores.  Float64: FP64 MMA (DMMA)
                                    fragments, MLP parameters staged in shared memory + 1 KB scratch per warp.  Float32: the
                                    tcgen05 path, see nn_tc_bytes */
    int32_t nodes_per_warp;      /* 0 / 32: one node per lane.  16: lanes l and l+16 carry the same node, so the warp-cooperative
                                    MLPs spread 16 basins over 32 lanes and a small, MLP-heavy run gets twice the warps */
    int32_t n_tangents;          /* D > 0: gradient run (HB_OUT_OBJECTIVE, Float64, no MLPs): the body was lowered with D
                                    parameter directions and the kernel also returns d objective / d direction per node —
                                    the forward-mode counterpart of the reference's Zygote gradient through its
                                    ImmutableSolver (README.md:161-187) */
    int32_t n_aux_inputs;        /* V2 > 0: a second per-node input stream aux_input[T][N][V2] (hb_run_device only): rows another
                                    kernel produced for this model, which the body copies into its records (hb_aux[k]) — the
                                    route state / outflow between the two stepper passes of the generic routed path */
    int32_t nn_tc_bytes;         /* precision = 1 and nn_tensor = 1: the body's tensor-shaped dense layers run on tcgen05.mma
                                    (kind::tf32, 3xTF32 operand split, accumulator in TMEM; templates/hb_tc.cuh) and need this
                                    many bytes of shared memory for their A / B tiles.  128 threads per block. */
} hb_stepper_spec;

/* One forward run.  Pointers are HOST pointers for hb_run and DEVICE pointers for
 * hb_run_device.  Optional pointers may be NULL. */
typedef struct {
    int32_t struct_size;         /* sizeof(hb_run_desc) */
    int32_t n_htype_sets;
    int64_t n_nodes;             /* N */
    int64_t n_steps;             /* T */
    const void *input;           /* [T][N][V]  ([T][1][V] when shared_forcing); double, or float in Float32 mode */
    const double *params;        /* flat parameter values                                     */
    int64_t n_param_values;      /* length of params                                          */
    const int32_t *param_offset; /* [n_params]  (HOST pointer in both entry points)           */
    const int32_t *param_mode;   /* [n_params]  (HOST pointer in both entry points)           */
    const int32_t *htypes;       /* [n_htype_sets][N], 0-based, or NULL                       */
    const double *initstates;    /* [S][N] or NULL (= zeros, src/bucket.jl:180,222)           */
    const double *uh_weights;    /* [sum_u kmax_u][N]: normalised ordinates, zero-padded      */
    const double *uh_hist_in;    /* [sum_u (kmax_u-1)][N] past UH inputs (chunked runs) / NULL */
    const double *nn_params;     /* [nn_words] MLP parameters (shared by all nodes)           */
    const double *target;        /* objective mode: observed series [T] or [T][N]             */
    int32_t target_per_node;     /* 0: target[T] shared, 1: target[T][N]                      */
    int32_t warm_up;             /* objective uses steps warm_up..T (1-based, as the reference) */
    double min_value;            /* state clamp (src/solve.jl:40,50); HydroConfig requires > 0 (src/config.jl:46), a NamedTuple config does not */
    void *out;                   /* HB_OUT_ROWS: [T][N][R]; double, or float in Float32 mode  */
    double *objective;           /* HB_OUT_OBJECTIVE: [N]                                     */
    double *final_states;        /* [S][N] or NULL: states after step T (resume / BMI)        */
    double *uh_hist_out;         /* [sum_u (kmax_u-1)][N] or NULL                             */
    float *elapsed_ms;           /* optional HOST pointer: device time of the kernel(s)       */
    double *gradient;            /* gradient runs (spec.n_tangents = D): [D][N], d objective[n] / d direction k */
    const void *aux_input;       /* spec.n_aux_inputs = V2 > 0: DEVICE [T][N][V2] (hb_run_device); NULL = the body reads zeros */
    int32_t n_compact_rows;      /* > 0 (hb_run_device, <= 4): store only rows compact_rows[] of every step, `out` is then */
    int32_t compact_rows[4];     /*   [T][N][n_compact_rows].  Same kernel, same arithmetic: the selected rows are bit-identical
                                      to those of a full run (first pass of the generic routed path) */
} hb_run_desc;

typedef struct {
    int32_t registers_per_thread;
    int32_t static_smem_bytes;
    int32_t dynamic_smem_bytes;
    int32_t local_bytes_per_thread; /* spills */
    int32_t block_threads;
    int32_t max_blocks_per_sm;
    int32_t from_cache;             /* 1 if the cubin came from the on-disk cache            */
    int32_t cubin_bytes;
} hb_kernel_info;

/* ---- library / device ------------------------------------------------------------------ */
int hb_version(void);                         /* HB_ABI_VERSION */
const char *hb_last_error(void);              /* thread-local, never NULL */
int hb_device_count(int *count);              /* HB_OK with *count == 0 when no driver */
int hb_init(int device);                      /* select device for this process (one GPU per process) */
int hb_set_cache_dir(const char *path);       /* cubin cache directory (default: $HB_CACHE_DIR or <lib dir>/cache) */

/* ---- compile ---------------------------------------------------------------------------- */
/* Splice `body` into the stepper template, NVRTC-compile for sm_100a (cached by source hash).
 * Works without a device (compile only); the module is loaded lazily on first run. */
int hb_compile_stepper(const hb_stepper_spec *spec, const char *body, hb_model_t *out);
/* The full CUDA source after splicing (for inspection, nvcc -Xptxas -v, ncu source pages). */
int hb_model_source(hb_model_t m, const char **src, int64_t *len);
int hb_model_cubin(hb_model_t m, const void **cubin, int64_t *len);
int hb_model_info(hb_model_t m, hb_kernel_info *info);   /* needs a device */
int hb_free_model(hb_model_t m);

/* ---- run --------------------------------------------------------------------------------- */
/* Host buffers: stages H2D, launches, stages D2H, returns after the stream is idle.  The run is cut into time
 * chunks that stream through a 3-slot device ring (H2D | kernel | D2H overlap; the forcing and the result never
 * have to fit in HBM together).  Pinned host buffers (hb_malloc_host / cudaHostRegister) are DMA'd directly;
 * pageable ones (a plain Julia Array) go through a pinned staging ring filled / drained by library host threads. */
int hb_run(hb_model_t m, const hb_run_desc *desc);
/* Device buffers, asynchronous on `stream` (a cudaStream_t, NULL = default stream).
 * Threading contract: a model handle is SINGLE-STREAM — the MLP parameter bank and the route / grid workspaces belong to
 * the handle, so launches of one handle must be ordered on one stream (or serialised by the caller with events); the
 * library is re-entrant across handles.  hb_init selects ONE device per process (one process per GPU). */
int hb_run_device(hb_model_t m, const hb_run_desc *desc, void *stream);

/* ---- grid / graph routing (HydroRoute, /root/reference/src/route.jl:336-404) ------------------- */
/* One route component with one state and one outflow, stepping in place on the model's records:
 * rows `in_row[]` of rec[T][N][R] are its inputs (written earlier by the per-node stepper), rows
 * `s_row` / `o_row` receive its state and outflow.  Topology = CSR list of UPSTREAM nodes
 * (0-based): the D8 push of build_aggr_func(flwdir, positions) or adjacency' * q, as a gather. */
typedef struct {
    int32_t struct_size;        /* sizeof(hb_route_spec) */
    int32_t n_inputs;           /* route inputs read from the records */
    int32_t n_params;
    int32_t block_threads;      /* 0 = default (256) */
} hb_route_spec;

typedef struct {
    int32_t struct_size;        /* sizeof(hb_route_desc) */
    int32_t n_htype_sets;
    int64_t n_nodes;            /* N */
    int64_t n_steps;            /* T */
    double *records;            /* DEVICE [T][N][R] */
    int32_t n_rows;             /* R */
    int32_t s_row, o_row;       /* record rows of the route state / outflow */
    const int32_t *in_row;      /* HOST [n_inputs] record rows of the route inputs */
    const double *params;       /* DEVICE flat parameter values */
    int64_t n_param_values;
    const int32_t *param_offset;/* HOST [n_params] */
    const int32_t *param_mode;  /* HOST [n_params] */
    const int32_t *htypes;      /* DEVICE [n_htype_sets][N] 0-based or NULL */
    const double *initstates;   /* DEVICE [N] or NULL (= zeros) */
    const int32_t *up_ptr;      /* DEVICE [N+1] */
    const int32_t *up_idx;      /* DEVICE [up_ptr[N]] */
    double min_value;
    double *final_states;       /* DEVICE [N] or NULL */
    float *elapsed_ms;          /* optional HOST pointer */
    /* planes mode (both non-NULL; `records` may then be NULL): the route inputs come from a compact DEVICE array
     * route_inputs[T][N][n_inputs] and the state / outflow go to route_out[T][N][2] — every access coalesced.  The generic
     * routed path is then three passes: stepper (rows = route inputs) -> this -> stepper (all rows, route_out as its
     * auxiliary input), each record written once. */
    const double *route_inputs;
    double *route_out;
} hb_route_desc;

int hb_compile_route(const hb_route_spec *spec, const char *body, hb_model_t *out);
/* T+1 launches of hb_route_step on `stream` (device pointers; returns after enqueueing unless
 * elapsed_ms is requested). */
int hb_run_route_device(hb_model_t m, const hb_route_desc *desc, void *stream);
/* ONE routing step (t = -1 is the prologue that only produces the outflows of step 0) on caller-owned
 * ping-pong buffers: s_prev/s_next [n_nodes], q_cur/q_next [n_nodes + n_ghost].  Multi-GPU runs own
 * a slice of the nodes; `up_idx` entries >= n_nodes address ghost slots of q_cur that the host fills
 * with the neighbouring ranks' boundary outflows between steps (NCCL send/recv, see parallel.py). */
int hb_route_step_device(hb_model_t m, const hb_route_desc *desc, int64_t t, const double *s_prev, double *s_next,
                         const double *q_cur, double *q_next, void *stream);

/* ---- reverse-mode gradient: the adjoint of the stepper --------------------------------------------------------------
 * What the reference computes with Zygote through its ImmutableSolver (/root/reference/README.md:161-187, src/solve.jl:62-79;
 * hybrid training, docs/src/tutorials/neuralnetwork_embeding.md): d loss / d (parameters, initial states, neural-network
 * weights) of   loss_n = scale[n] * sum_{t >= warm_up} (sim_n(t) - target(t))^2   for per-node models of buckets, fluxes and
 * neural fluxes.  The host lowers the SAME step expressions to a reverse-mode body (PROLOGUE / STEP / EPILOGUE); the kernel
 * walks the stored state trajectory backwards (templates/adjoint.cu).  All pointers DEVICE unless noted. */
typedef struct {
    int32_t struct_size;         /* sizeof(hb_adjoint_spec) */
    int32_t n_inputs, n_states, n_params;
    int32_t nn_words;            /* MLP parameters (all chains) */
    int32_t block_threads;       /* 0 = 128 */
    int32_t acc_doubles_per_warp;/* shared-memory accumulator fragments of the dense layers' weight gradients (from the lowering) */
} hb_adjoint_spec;
typedef struct {
    int32_t struct_size;         /* sizeof(hb_adjoint_desc) */
    int32_t n_htype_sets;
    int64_t n_nodes, n_steps;
    const double *input;         /* [T][N][V] */
    const double *trajectory;    /* [T][N][S]: post-update states of every step (forward run with rows = the states) */
    const double *params;
    int64_t n_param_values;
    const int32_t *param_offset; /* HOST [n_params] */
    const int32_t *param_mode;   /* HOST [n_params] */
    const int32_t *htypes;
    const double *initstates;    /* [S][N] or NULL */
    const double *nn_params;     /* [nn_words] */
    const double *target;        /* [T] or [T][N]; NULL: loss_n = scale[n] * sum_{t >= warm_up} sim_n(t) — the linear functional the
                                    reference's gradient tests differentiate (test/gradient/run_hybrid_gradient.jl:39-44) */
    int32_t target_per_node, warm_up;
    double min_value;
    const double *scale;         /* [N]: 1 (SSE), 1/count (MSE), 1/sum (obs - mean obs)^2 (NSE) */
    double *grad_params;         /* [n_params][N]: d loss_n / d (parameter slot j of node n) */
    double *grad_init;           /* [S][N] */
    double *grad_nn;             /* [nn_words]: summed over all nodes (zeroed by the call) */
    double *loss;                /* [N] */
    float *elapsed_ms;           /* optional HOST pointer */
} hb_adjoint_desc;
int hb_compile_adjoint(const hb_adjoint_spec *spec, const char *body, hb_model_t *out);
int hb_run_adjoint_device(hb_model_t m, const hb_adjoint_desc *desc, void *stream);

/* ---- fused dense-grid model: per-cell buckets + D8 HydroRoute in one pass ----------------------------
 * Replaces, for `[buckets..., HydroRoute]` models whose node list is the complete row-major grid, the two-pass
 * generic path above (/root/reference/src/model.jl:273-309 sequencing + src/route.jl:336-404).  Two kernel
 * variants: HB_GRID_WAVE (templates/gridwave.cu, default: warps exchange outflows through L2 with
 * release/acquire step counters, no redundant work) and HB_GRID_TILE (templates/grid.cu: tile + halo). */
#define HB_GRID_TILE 0
#define HB_GRID_WAVE 1
typedef struct {
    int32_t struct_size;        /* sizeof(hb_grid_spec) */
    int32_t n_inputs;           /* V */
    int32_t n_rows;             /* R: all model rows (bucket rows + the route's state and outflow) */
    int32_t n_states;           /* S: bucket states (the route state is separate) */
    int32_t n_params;           /* all parameter slots: per-cell part first, then the route's */
    int32_t n_params_stepper;   /* slots of the per-cell part */
    int32_t n_route_inputs;
    int32_t s_row, o_row;       /* record rows of the route state / outflow */
    int32_t steps_per_chunk;    /* T_c: steps per launch.  TILE: = halo depth (1..8).  WAVE: 1..4096, clamped at run
                                 * time so that the dependency cone fits in the co-resident blocks; 0 = chosen from the
                                 * grid width (a quarter of the co-resident rows, 4..32) */
    int32_t block_w, block_h;   /* TILE: region (tile + halo) in cells; block_w % 32 == 0, block_w*block_h <= 1024.
                                 * WAVE: block_w = threads per block (0 = 128), block_h ignored */
    int32_t variant;            /* HB_GRID_TILE | HB_GRID_WAVE */
    int32_t stages;             /* WAVE: forcing pipeline depth in time steps (0 = 3) */
    int32_t max_registers;      /* WAVE: register budget per thread (0 = compiler's choice at 4 blocks/SM) */
} hb_grid_spec;

typedef struct {
    int32_t struct_size;        /* sizeof(hb_grid_desc) */
    int32_t n_htype_sets;
    int32_t rows, cols;         /* grid shape; node n = r * cols + c (0-based), N = rows * cols */
    int64_t n_steps;            /* T */
    const double *input;        /* DEVICE [T][N][V] */
    double *out;                /* DEVICE [T][N][R] */
    const double *params;       /* DEVICE flat values */
    int64_t n_param_values;
    const int32_t *param_offset;/* HOST [n_params] */
    const int32_t *param_mode;  /* HOST [n_params] */
    const int32_t *htypes;      /* DEVICE [n_htype_sets][N] 0-based or NULL */
    const double *init_bucket;  /* DEVICE [S][N] or NULL */
    const double *init_route;   /* DEVICE [N] or NULL */
    const uint8_t *flwdir;      /* DEVICE [rows][cols] D8 codes */
    double min_value;
    double *final_bucket;       /* DEVICE [S][N] or NULL */
    double *final_route;        /* DEVICE [N] or NULL */
    float *elapsed_ms;          /* optional HOST pointer */
    int32_t out_row0, out_rows; /* WAVE: records are written only for the cells of rows [out_row0, out_row0 + out_rows) x columns   */
    int32_t out_col0, out_cols; /*   [out_col0, out_col0 + out_cols), as out[T][out_rows][out_cols][R] (0 rows / cols = all).  A strip or */
                                /*   panel of a multi-GPU run computes ghost rows / columns (depth = steps between state exchanges)     */
                                /*   whose records belong to the neighbouring rank.                                                      */
    /* WAVE, pitched addressing (all 0 = compact arrays): run a sub-grid (a column panel of a wider grid) in place on the wider
     * grid's arrays.  Forcing of local cell (r, c) at step t is input[(t * in_t_stride + r * in_pitch + in_col0 + c) * V]; the
     * record of an owned cell goes to out[(t * out_t_stride + (r - out_row0) * out_pitch + out_base + (c - out_col0)) * R]. */
    int64_t in_t_stride, out_t_stride;
    int32_t in_pitch, in_col0, out_pitch, out_base;
} hb_grid_desc;

int hb_compile_grid(const hb_grid_spec *spec, const char *body, hb_model_t *out);
int hb_run_grid_device(hb_model_t m, const hb_grid_desc *desc, void *stream);   /* ceil(T / T_c) launches */
/* WAVE variant: synchronises `stream` and reports whether any warp of the runs enqueued so far gave up waiting
 * for a neighbour (bounded poll loops; would indicate a forward-progress bug or a foreign kernel hogging the
 * SMs).  Returns HB_ERR_CUDA with a message if so; `steps_per_launch` (optional) receives the clamped T_c. */
int hb_grid_status(hb_model_t m, void *stream, int32_t *steps_per_launch);

/* ---- IRF river routing: SparseRouteConvolution (/root/reference/src/route.jl:593-610) -------------------
 * out[t][d] = sum over the kernel rows of destination d (CSR `row_ptr`, original row order) of the causal
 * convolution of in[.][src[row]] with weights[.][row].  DEVICE pointers; in [T][n_in], out [T][n_out],
 * weights [H][nnz] with rows in CSR order. */
int hb_sparse_route_conv_device(int n_out, int n_in, int64_t n_steps, int horizon, int64_t nnz, const int32_t *row_ptr,
                                const int32_t *src, const double *weights, const double *input, double *out, void *stream);
/* Same, with `dest_order` (DEVICE, a permutation of 0..n_out-1 or NULL): the sequence in which the kernel's blocks take the
 * destinations, 32 per block.  A block runs for as many rows as its longest destination, so a network whose destinations
 * have very different row counts (a few outlets with hundreds of upstream sources) wants them sorted by row count; the
 * result does not depend on the order. */
int hb_sparse_route_conv_ordered_device(int n_out, int n_in, int64_t n_steps, int horizon, int64_t nnz, const int32_t *row_ptr,
                                        const int32_t *src, const double *weights, const double *input,
                                        const int32_t *dest_order, double *out, void *stream);

/* River networks are skewed: the outlet's kernel has one row per node of the basin, and a convolution block runs for as
 * many rows as its longest destination.  The host therefore splits long destinations into segments of rows, convolves
 * every segment as a "virtual destination" (hb_sparse_route_conv_ordered_device with n_out = n_virtual into
 * part[T][n_virtual]) and this kernel adds a destination's segments in order: out[t][d] = sum part[t][seg_ptr[d] ..
 * seg_ptr[d+1]).  Deterministic; differs from the reference's single running sum by rounding only (<= 1e-13 relative). */
int hb_irf_reduce_segments_device(int32_t n_out, int64_t n_steps, int32_t n_virtual, const int32_t *seg_ptr, const double *part,
                                  double *out, void *stream);
/* aggregate_route_kernel (/root/reference/src/route.jl:523-577) on the device: composition of the node kernels down the
 * network, level by level (level_row_ptr is a HOST array of n_levels + 1 row offsets; everything else DEVICE).  Row r
 * belongs to node row_dest[r]; its contributions contrib_ptr[r] .. contrib_ptr[r+1] name finished rows of upstream
 * neighbours (contrib_row) with their edge weights, in the reference's accumulation order; a row without contributions is
 * the node's own kernel.  weights is [n_rows][horizon].  Bit-identical to the reference's arithmetic. */
int hb_irf_compose_device(int32_t n_levels, const int64_t *level_row_ptr, const int32_t *row_dest, const int64_t *contrib_ptr,
                          const int64_t *contrib_row, const double *contrib_ew, const double *node_kernels, double *weights,
                          int32_t horizon, void *stream);
/* out row i = w row perm[i] (perm NULL = identity); transpose = 1 writes out[horizon][nnz], the layout
 * hb_sparse_route_conv_device takes, else out[nnz][horizon] (SparseRouteKernel.weights). */
int hb_irf_gather_rows_device(const double *w, const int64_t *perm, int64_t nnz, int32_t horizon, int32_t transpose, double *out,
                              void *stream);

/* ---- raster ingest (/root/reference/ext/HydroModelsRastersExt) -------------------------------------------------
 * compute_d8_flow_direction (HydroModelsRastersExt.jl:32-109): steepest-descent D8 codes (1 E, 2 SE, 4 S, 8 SW, 16 W,
 * 32 NW, 64 N, 128 NE; 0 for NaN cells) from a DEM.  dem[r * row_stride + c * col_stride] (element strides: a Julia
 * matrix is row_stride 1, col_stride rows); flow_dir is [rows][cols] row-major, the layout hb_grid_desc.flwdir takes. */
int hb_d8_flow_direction_device(const double *dem, int32_t rows, int32_t cols, int64_t row_stride, int64_t col_stride,
                                double cellsize_x, double cellsize_y, uint8_t *flow_dir, void *stream);
/* load_netcdf_data + input assembly (netcdf_loader.jl:29-62, :119-136): gather `n_vars` rasters (DEVICE array of DEVICE
 * pointers; element (t, r, c) at t * t_stride + r * row_stride + c * col_stride) at the node positions (0-based) into
 * the stepper's forcing layout out[n_steps][n_nodes][n_vars].  n_steps <= 65535 per call. */
int hb_pack_forcing_device(const double *const *vars_device_array, int32_t n_vars, int64_t n_steps, int64_t t_stride,
                           int64_t row_stride, int64_t col_stride, const int32_t *node_row, const int32_t *node_col,
                           int64_t n_nodes, double *out, void *stream);
/* Same for the dense case — the nodes are all rows x cols cells in row-major order (n = r * cols + c), n_vars <= 5: no
 * position arrays, 32 x 32 tiles through shared memory so reads follow the raster's fastest axis and writes are whole runs
 * of out[t][n][v] whatever the raster's memory order. */
int hb_pack_forcing_dense_device(const double *const *vars_device_array, int n_vars, int64_t n_steps, int64_t t_stride,
                                 int64_t row_stride, int64_t col_stride, int rows, int cols, double *out, void *stream);

/* The same gather from HOST rasters, streamed (netcdf_loader.jl:29-62, :119-136 + the upload): `vars_host[v]` is raster v in
 * host memory (pinned or ordinary pageable memory — a Julia Array, a numpy array, an mmap'd file), time axis slowest
 * (t_stride == rows * cols; element (t, r, c) at t * t_stride + r * row_stride + c * col_stride).  Time chunks go through two
 * device scratch slots (and two pinned staging slots when the source is pageable, filled by the library's host threads), so the
 * copy of chunk c + 1 runs under the pack kernel of chunk c and the rasters never sit on the device as a whole.
 * node_row / node_col: DEVICE arrays of 0-based positions, or both NULL for the dense case (all cells row-major, n_vars <= 5).
 * `out`: DEVICE [n_steps][n_nodes][n_vars].  Synchronous: returns when `out` is complete on `stream`. */
int hb_ingest_forcing(const double *const *vars_host, int32_t n_vars, int64_t n_steps, int64_t t_stride, int64_t row_stride,
                      int64_t col_stride, int32_t rows, int32_t cols, const int32_t *node_row, const int32_t *node_col,
                      int64_t n_nodes, double *out, void *stream);

/* ---- on-device calibration: differential evolution around the ensemble objective kernel -----------------------
 * Replaces the host loop `solve(OptimizationProblem(component, input, target; ...), BBO_adaptive_de_rand_1_bin...)`
 * (/root/reference/ext/HydroModelsOptimizationExt/HydroModelsOptimizationExt.jl:72-199, docs/src/tutorials/
 * optimimal_parameters.md:189-195) for population methods: candidates live where the stepper reads its parameters
 * (dimension d of member i at base[dim_offset[d] + i]); one generation = hb_de_trial_device, hb_run_device in
 * objective mode on the trial buffer, hb_de_select_device — all DEVICE pointers, asynchronous on `stream`.
 * DE/rand/1/bin, clamped to [lower, upper]; counter-based SplitMix64 draws keyed by (seed, generation, member). */
int hb_de_trial_device(const double *pop, double *trial, const int64_t *dim_offset, int32_t n_dims, int32_t n_members,
                       const double *lower, const double *upper, double F, double CR, uint64_t seed, uint32_t generation,
                       void *stream);
/* trial replaces member where trial_fitness <= fitness, or where the member's own fitness is NaN and the trial's is not
 * (a NaN trial never wins); `history` (optional, DEVICE [2 * slots]) receives the best objective of the population and
 * its index at slot `history_slot`.  `pop`/`trial` and `fitness`/`trial_fitness` must not alias. */
int hb_de_select_device(double *pop, const double *trial, double *fitness, const double *trial_fitness,
                        const int64_t *dim_offset, int32_t n_dims, int32_t n_members, double *history, int32_t history_slot,
                        void *stream);
/* Best (lowest, NaN ignored) objective of `fitness[n_members]` and its first index -> history[2*slot], history[2*slot+1]
 * (generation 0 bookkeeping: no selection). */
int hb_de_best_device(const double *fitness, int32_t n_members, double *history, int32_t history_slot, void *stream);

/* The same DE with the population SHARDED over the ranks of a communicator (one process per GPU): global member g lives on rank
 * g / n_local at local index g % n_local; every rank holds the whole population as pop_all[world][n_dims][n_local] (the ranks'
 * blocks, refreshed by ONE hb_comm_all_gather per generation — the only communication) because the parents come from all
 * members; trials are written where the local stepper reads its parameters, objectives and selection stay local
 * (`pop_block` = pop_all + rank * n_dims * n_local; `history`: this rank's best).  Draws are keyed by the global member index,
 * so the search follows the single-GPU trajectory bit for bit.  n_total % n_local == 0, member0 = rank * n_local. */
int hb_de_trial_sharded_device(const double *pop_all, double *trial, const int64_t *dim_offset, int32_t n_dims, int32_t n_local,
                               int32_t n_total, int32_t member0, const double *lower, const double *upper, double F, double CR,
                               uint64_t seed, uint32_t generation, void *stream);
int hb_de_select_sharded_device(double *pop_block, const double *trial, double *fitness, const double *trial_fitness,
                                const int64_t *dim_offset, int32_t n_dims, int32_t n_local, double *history, int32_t history_slot,
                                void *stream);

/* Adaptive DE — BlackBoxOptim.jl's `adaptive_de_rand_1_bin_radiuslimited`, the optimiser the reference's examples pass to
 * `solve` (ext/HydroModelsOptimizationExt/examples.jl:28, docs/src/tutorials/optimimal_parameters.md:189-195), generation-
 * synchronous: per-member F and CR (`f`, `cr`: DEVICE [n_members], drawn by hb_ade_init_device; a member whose trial did not
 * strictly improve it redraws both from the bimodal Cauchy mixtures F ~ {C(0.65,0.1) | C(1,0.1)}, CR ~ {C(0.1,0.1) |
 * C(0.95,0.1)}, above 1 -> 1, below 0 -> redrawn), parents from a window of `radius` (>= 4, BlackBoxOptim: 8) consecutive members
 * around the target, out-of-bounds components placed at random between the target's value and the bound.  Same buffers,
 * aliasing rule and history as the plain DE entry points; `generation` >= 1 keys the draws. */
int hb_ade_init_device(double *f, double *cr, int32_t n_members, uint64_t seed, void *stream);
int hb_ade_trial_device(const double *pop, double *trial, const int64_t *dim_offset, int32_t n_dims, int32_t n_members,
                        const double *lower, const double *upper, const double *f, const double *cr, int32_t radius, uint64_t seed,
                        uint32_t generation, void *stream);
int hb_ade_select_device(double *pop, const double *trial, double *fitness, const double *trial_fitness,
                         const int64_t *dim_offset, int32_t n_dims, int32_t n_members, double *f, double *cr, uint64_t seed,
                         uint32_t generation, double *history, int32_t history_slot, void *stream);

/* ---- memory helpers (so a host without a CUDA binding can own device / pinned buffers) ---- */
int hb_malloc_device(void **ptr, int64_t bytes);
int hb_free_device(void *ptr);
int hb_malloc_host(void **ptr, int64_t bytes);   /* pinned */
int hb_free_host(void *ptr);
int hb_memcpy_h2d(void *dst, const void *src, int64_t bytes, void *stream);
int hb_memcpy_d2h(void *dst, const void *src, int64_t bytes, void *stream);
int hb_memset_device(void *dst, int value, int64_t bytes, void *stream);
int hb_stream_create(void **stream);              /* a non-blocking cudaStream_t for the *_device entry points */
int hb_stream_destroy(void *stream);
int hb_stream_synchronize(void *stream);
int hb_convert_f64_f32_device(const double *src, float *dst, int64_t n, void *stream);   /* DEVICE arrays */


/* ---- multi-GPU: one process per GPU (hb_init picks the device), a communicator inside the library -------------------
 * The reference is single-process and single-threaded (src/config.jl:36 `parallel` is never read); SURVEY.md §8(e) shards the
 * node axis.  Lumped / ensemble runs need no exchange (every rank calls hb_run on its slice; hb_comm_all_gather collects
 * objective vectors).  A dense D8 grid (src/route.jl:336-404) is cut into row strips or column panels with `halo` ghost
 * rows / columns; every `halo` steps the ghost cells' bucket and river STATES are refreshed from the neighbouring ranks with
 * hb_halo_exchange_device.  NCCL is resolved at run time from libnccl.so.2 ($HB_NCCL_LIB overrides); the 128-byte id of
 * rank 0 travels to the other ranks by whatever the host has (MPI, a file, torch.distributed). */
typedef struct hb_comm_s *hb_comm_t;
int hb_comm_unique_id(void *id, int32_t id_bytes);                 /* rank 0: ncclGetUniqueId into id[128] */
int hb_comm_init(const void *id, int32_t rank, int32_t world, hb_comm_t *out);      /* collective (ncclCommInitRank) */
int hb_comm_rank(hb_comm_t c, int32_t *rank, int32_t *world);
int hb_comm_free(hb_comm_t c);
int hb_comm_all_gather(hb_comm_t c, const void *send, void *recv, int64_t bytes_per_rank, void *stream);   /* DEVICE buffers */
/* Collective symmetric allocation: every rank allocates `bytes` (zeroed) and maps the other ranks' copies through CUDA IPC
 * (NVLink peer access); hb_comm_peer_ptr returns rank `peer`'s copy as an address valid in THIS process. */
int hb_comm_alloc(hb_comm_t c, int64_t bytes, void **local_ptr, int32_t *segment);
int hb_comm_peer_ptr(hb_comm_t c, int32_t segment, int32_t peer, void **ptr);

/* One neighbour of a sub-grid.  All planes are [rows][pitch] double arrays (bucket state s of cell (r, c) at
 * planes[s][r * pitch + c]; the river state is one more plane).  The block I send and the block I receive have the same shape. */
typedef struct {
    int32_t peer;               /* rank on the other side */
    int32_t s_row0, s_col0;     /* first cell of the block of MY owned cells the peer needs */
    int32_t r_row0, r_col0;     /* first cell of MY ghost block the peer fills */
    int32_t nrows, ncols;       /* block shape */
    int32_t peer_slot;          /* index of the matching edge in the PEER's edge list (its staging slot; peer-store mode) */
} hb_halo_edge;
#define HB_HALO_NCCL 0          /* pack kernel, one grouped ncclSend/ncclRecv per neighbour on `stream`, unpack kernel */
#define HB_HALO_PEER 1          /* pack kernel stores into the neighbours' staging over NVLink + epoch flag; unpack kernel waits for it */
/* Collective, once per communicator before the first HB_HALO_PEER exchange: reserve the symmetric staging
 * (`max_block_bytes` >= n_planes * nrows * ncols * 8 of the largest block). */
int hb_halo_enable_peer(hb_comm_t c, int64_t max_block_bytes);
/* Refresh the ghost blocks of `planes` (HOST array of n_planes DEVICE pointers) from the neighbours, asynchronous on `stream`.
 * Every rank calls it the same number of times, in the same order, on one stream; at most 8 planes and 8 edges. */
int hb_halo_exchange_device(hb_comm_t c, double *const *planes, int32_t n_planes, int32_t pitch, const hb_halo_edge *edges,
                            int32_t n_edges, int32_t mode, void *stream);
int hb_halo_status(hb_comm_t c, void *stream);    /* synchronises; error if a peer-mode wait expired */

/* ---- measurement probes (bench.py) --------------------------------------------------------- */
int hb_probe_fp64_tflops(double *tflops);                       /* measured FP64 FMA peak (2 flop/FMA) */
int hb_probe_dmma_tflops(double *tflops);                       /* measured FP64 tensor (DMMA m8n8k4) peak */
int hb_fill_device(double *p, int64_t n, double v, void *stream); /* L2 flush / buffer fill           */
/* Test aid for the Float32 tensor path (tcgen05.mma kind::tf32, 3xTF32 split, accumulator in TMEM; templates/hb_tc.cuh):
 * D[128][n] = A[128][k] * W^T with W[n_idx + n * k_idx] (the Lux Dense weight layout), DEVICE pointers,
 * (k, n) in {(16, 16), (24, 32), (64, 64)}; `reps` rounds through the same TMEM columns. */
int hb_probe_tc_layer(int k, int n, const float *A, const double *W, float *D, int reps, void *stream);
/* Test aid: `n_blocks` blocks holding `smem_bytes` of shared memory each spin for `ms` milliseconds on `stream`
 * (asynchronous) — a co-tenant kernel for the wave grid kernel's forward-progress tests. */
int hb_probe_occupy_sms(int n_blocks, int smem_bytes, double ms, void *stream);
/* Host <-> device link (the roofline of hb_run's host-buffer path): `reps` cudaMemcpyAsync of `bytes` between pinned
 * host memory and the device; gbs[4] = H2D alone, D2H alone, H2D while D2H runs, D2H while H2D runs (GB/s). */
int hb_probe_pcie_gbs(int64_t bytes, int reps, double *gbs);
/* Host threads hb_run uses to move chunks of PAGEABLE caller arrays through its pinned staging ring
 * ($HB_HOST_THREADS, default min(16, cores of the process's affinity mask)). */
int hb_host_threads(int *n);
/* The copy those threads perform (parallel, streaming stores): host to host, no device needed. */
int hb_host_copy(void *dst, const void *src, int64_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* HYDROB200_H */
