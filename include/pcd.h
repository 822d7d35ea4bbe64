/* pcd.h -- C ABI of the B200-native caustic-design hot path (libpcd_b200.so).
 *
 * The reference (dylanmsu/poisson_caustic_design) has no FFI: its boundary is the C++ API
 *   void poisson_solver(vector<vector<double>>& D, vector<vector<double>>& phi, int width, int height,
 *                       int max_iterations, double convergence_threshold, int max_threads)   src/solver.h:8
 *   class Caustic_design { ... }                                                             src/caustic_design.h:7-66
 * and the CLI in main.cpp:137-272.  The host C++ shim under poisson_caustic_design_b200/host/ keeps those
 * signatures and calls only the functions below.  Everything here is plain pointers and sizes; all
 * arithmetic is fp64; grids are flat row-major [y*W + x]; points are SoA.
 *
 * Conventions: every call returns a pcd_status (0 = ok); no exceptions and no stdout cross this
 * boundary; pcd_last_error() gives the message of the last failure on the calling thread.  A
 * context / solver is not thread-safe; distinct ones are independent.  All pointers are HOST pointers
 * unless the name says `_dev`.
 */
#ifndef PCD_H
#define PCD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCD_ABI_VERSION 2   /* 2: per-slab SM reserve, persistent peer run, pcd_multi_*, tolerances, DCT backend (round 2) */

typedef enum pcd_status {
    PCD_OK = 0,
    PCD_ERR_INVALID = 1,      /* bad argument / size */
    PCD_ERR_CUDA = 2,         /* a CUDA runtime call failed (message has the CUDA error string) */
    PCD_ERR_NO_DEVICE = 3,    /* no usable CUDA device: the library never falls back to a CPU path */
    PCD_ERR_RASTER_MISS = 4,  /* a raster / inverse-map sample hit no triangle; the reference exit(0)s here, src/mesh.cpp:276-281 */
    PCD_ERR_STATE = 5,        /* call order (e.g. iteration before pcd_initialize_solvers) */
    PCD_ERR_UNSUPPORTED = 6
} pcd_status;

/* Public data members of Caustic_design / Mesh (src/caustic_design.h:16-30, src/mesh.h:51-52), flattened.
 * Grid fields have res_x*res_y elements, vertex fields mesh_res_x*mesh_res_y. */
typedef enum pcd_field {
    PCD_FIELD_PHI = 0,              /* grid  */
    PCD_FIELD_H = 1,                /* grid  */
    PCD_FIELD_RASTER = 2,           /* grid, mean-removed D as handed to the solver */
    PCD_FIELD_PIXELS = 3,           /* grid  */
    PCD_FIELD_DIVERGENCE = 4,       /* grid, mean-removed */
    PCD_FIELD_NORM_X = 5,           /* grid  */
    PCD_FIELD_NORM_Y = 6,           /* grid  */
    PCD_FIELD_GRADIENT_X = 7,       /* grid, computed on demand from phi (never materialised on the hot path) */
    PCD_FIELD_GRADIENT_Y = 8,       /* grid, on demand */
    PCD_FIELD_ERRORS = 9,           /* vertex */
    PCD_FIELD_TARGET_AREAS = 10,    /* vertex */
    PCD_FIELD_VERTEX_GRADIENT_X = 11,
    PCD_FIELD_VERTEX_GRADIENT_Y = 12,
    PCD_FIELD_NORMALS_X = 13,
    PCD_FIELD_NORMALS_Y = 14,
    PCD_FIELD_TARGET_X = 15,        /* mesh->target_points[i][0] */
    PCD_FIELD_TARGET_Y = 16,
    PCD_FIELD_TARGET_Z = 17,
    PCD_FIELD_SOURCE_X = 18,        /* mesh->source_points[i][0] */
    PCD_FIELD_SOURCE_Y = 19,
    PCD_FIELD_SOURCE_Z = 20,
    PCD_FIELD_COUNT = 21
} pcd_field;

/* Which SOR kernel family runs a solve. */
typedef enum pcd_solver_path {
    PCD_SOLVER_AUTO = 0,       /* resident when the grid fits on chip, tiled otherwise */
    PCD_SOLVER_STREAMING = 1,  /* one launch per colour, phi/D streamed from HBM/L2 (also the NaN-hole path) */
    PCD_SOLVER_RESIDENT = 2,   /* one persistent cooperative kernel, phi/D live in registers/shared memory */
    PCD_SOLVER_TILED = 3,      /* temporal blocking: several sweeps per pass over shared-memory tiles, ping-pong fields */
    PCD_SOLVER_DCT = 4         /* OPT-IN direct backend: the converged field of the same discrete operator by
                                  a DCT-II diagonalisation (four fp64 GEMMs); ignores max_iterations / tol / warm start, so it
                                  is not the parity path (the reference stops at max|delta| < tol); never chosen by AUTO;
                                  D with NaN holes falls back to the masked sweeps */
} pcd_solver_path;

typedef struct pcd_config {
    int mesh_res_x, mesh_res_y;   /* Caustic_design::set_mesh_resolution    src/caustic_design.cpp:28-31 */
    int res_x, res_y;             /* Caustic_design::set_domain_resolution  :33-36 */
    double width, height;         /* Caustic_design::set_mesh_size          :38-41 */
    double focal_l;               /* set_lens_focal_length                  :43-45 */
    double thickness;             /* set_lens_thickness                     :47-49 */
    int device;                   /* CUDA device ordinal */
    int solver_path;              /* pcd_solver_path */
} pcd_config;

typedef struct pcd_solve_info {
    int sweeps;             /* full red+black sweeps executed */
    int converged_at;       /* sweeps up to and including the first with max|delta| < tol; 0 = cap hit */
    double last_max_update; /* max|delta| of the sweep that satisfied the test (or of the last sweep) */
    double device_ms;       /* CUDA-event time of the whole solve on its stream */
    double kernel_ms;       /* CUDA-event time of the SOR kernel launches alone (roofline numerator's clock) */
    int launches;           /* kernel launches issued by the solve */
    int path;               /* pcd_solver_path actually used */
} pcd_solve_info;

typedef struct pcd_ctx pcd_ctx;
typedef struct pcd_solver pcd_solver;

/* ---- library ------------------------------------------------------------------------------- */
int pcd_abi_version(void);
const char *pcd_last_error(void);
int pcd_device_count(int *count);
/* total kernel launches issued by this library in this process (bench.py's gpu_launches) */
long long pcd_launch_count(void);

/* ---- class Caustic_design (src/caustic_design.h:7-66) ---------------------------------------- */
int pcd_create(const pcd_config *cfg, pcd_ctx **out);                     /* ctor + setters :5-53 */
void pcd_destroy(pcd_ctx *ctx);
int pcd_initialize_solvers(pcd_ctx *ctx, const double *image);            /* :334-364  image[res_y*res_x] */
int pcd_perform_transport_iteration(pcd_ctx *ctx, double *step_out);      /* :190-266 */
int pcd_perform_height_map_iteration(pcd_ctx *ctx, int itr, double *update_sum_out /* may be NULL */); /* :269-332 */
/* Runs main.cpp:243-256 on the device (<= max_iters iterations, stop when step < conv_tres) with a
 * single host read of the step per iteration; steps_out (may be NULL) receives each step. */
int pcd_run_transport(pcd_ctx *ctx, int max_iters, double conv_tres, int *iters_out, double *steps_out);
int pcd_field_size(const pcd_ctx *ctx, int field, long *n_out);
int pcd_get_field(pcd_ctx *ctx, int field, double *dst);
int pcd_set_field(pcd_ctx *ctx, int field, const double *src);            /* members are public in the reference */
int pcd_inverted_transport_map(pcd_ctx *ctx, double *out_x, double *out_y); /* Mesh::calculate_inverted_transport_map src/mesh.cpp:348-409 */
int pcd_last_solve_info(const pcd_ctx *ctx, pcd_solve_info *info);
/* resident kernel of the context's last built-in solve (see pcd_solver_resident_exchange): 0 none, 1 per phase, 2 per sweep */
int pcd_resident_exchange(const pcd_ctx *ctx);
/* totals over every solve since the last reset: sweeps, launches, kernel_ms, device_ms are summed */
int pcd_solve_totals(pcd_ctx *ctx, pcd_solve_info *totals, int reset);
/* device timing on the context's own stream (torch.cuda.Event only sees torch's stream): slots 0..7 */
int pcd_event_record(pcd_ctx *ctx, int slot);
int pcd_event_elapsed_ms(pcd_ctx *ctx, int slot_start, int slot_stop, double *ms_out);
/* evicts L2 by writing a 256 MiB scratch buffer on the context's stream (bench hygiene between steps) */
int pcd_flush_l2(pcd_ctx *ctx);
/* stage entry points (per-stage parity tests drive these with the oracle's inputs) */
/* Poisson solves of a context can be handed to an external solver (the multi-GPU slab driver, slab.py): the hook
 * receives DEVICE pointers of the context's width x height arrays (D read-only, phi in/out; the reference call it
 * replaces is poisson_solver, src/solver.h:8, as issued at src/caustic_design.cpp:222,311).  It runs on the calling
 * thread after the context's stream has been drained and must leave phi complete on the device when it returns;
 * a non-zero return value becomes the status of the pcd_* call that needed the solve.  NULL restores the built-in
 * solver. */
typedef int (*pcd_solve_hook)(void *user, const double *D_dev, double *phi_dev, int width, int height, int max_iterations,
                              double tol, pcd_solve_info *info);
int pcd_set_solve_hook(pcd_ctx *ctx, pcd_solve_hook hook, void *user);
/* Stopping thresholds (max|delta| of a sweep) of the two Poisson solves; <= 0 keeps the current value.
 * Defaults: transport 1e-7 = the reference's (src/caustic_design.cpp:222); height 1e-9, TIGHTER than the
 * reference's 1e-8 (:311): at 1e-8 the reference's lexicographic sweeps and red-black sweeps stop on opposite
 * sides of the converged discrete field (C1: 6.7e-5 and 7.3e-5 of the height range away from it, 1.4e-4 from
 * each other); at 1e-9 the red-black result is 7e-6 from the converged field, i.e. the remaining distance to the
 * reference is the reference's own truncation error (tests/test_oracle_golden.py::test_height_truncation_evidence).
 * pcd_set_tolerances(ctx, 0, 1e-8) restores the reference's threshold. */
int pcd_set_tolerances(pcd_ctx *ctx, double transport_tol, double height_tol);
int pcd_stage_errors(pcd_ctx *ctx);                                       /* caustic_design.cpp:194-209 */
int pcd_stage_raster(pcd_ctx *ctx);                                       /* :212-213, no mean removal */
int pcd_stage_subtract_average(pcd_ctx *ctx);                             /* :221 */
int pcd_stage_solve_transport(pcd_ctx *ctx);                              /* :222 */
int pcd_stage_step(pcd_ctx *ctx, double *step_out);                       /* :225-265 */

/* ---- poisson_solver (src/solver.h:8) ----------------------------------------------------------- */
/* One-shot drop-in: D and phi are host arrays [height*width]; phi is in/out (warm start).  max_threads of
 * the reference has no meaning here. */
int pcd_poisson_solver(const double *D, double *phi, int width, int height, int max_iterations,
                       double convergence_threshold, int device, pcd_solve_info *info /* may be NULL */);

/* Device-resident solver object (bench / repeated solves / explicit path selection). */
int pcd_solver_create(int width, int height, int device, int solver_path, pcd_solver **out);
void pcd_solver_destroy(pcd_solver *s);
int pcd_solver_upload(pcd_solver *s, const double *D /* or NULL */, const double *phi /* or NULL */);
int pcd_solver_download(pcd_solver *s, double *phi);
/* the same with DEVICE arrays on the solver's device (device-to-device copies on the solver's stream; either
 * argument of load may be NULL) */
int pcd_solver_load_device(pcd_solver *s, const double *D_dev, const double *phi_dev);
int pcd_solver_store_device(pcd_solver *s, double *phi_dev);
/* check_lag: convergence of sweep s is acted on after sweep s+check_lag (resident path) or at the next
 * multiple of check_lag (streaming path); <= 0 selects the default, values above 4094 are clamped to 4094.  The field is bit-identical to
 * the red-black oracle run for `sweeps` sweeps. */
int pcd_solver_set_check_lag(pcd_solver *s, int check_lag);
int pcd_solver_run(pcd_solver *s, int max_iterations, double convergence_threshold, pcd_solve_info *info);
int pcd_solver_path_used(const pcd_solver *s);
/* Which resident kernel the last pcd_solver_run used: 0 = none (another path), 1 = one neighbour exchange per colour
 * phase, 2 = one per sweep (deep halos: even width, no NaN holes, more than two rows per CTA).  Same results. */
int pcd_solver_resident_exchange(const pcd_solver *s);
/* What PCD_SOLVER_AUTO would choose for a width x height grid on a device with sm_count SMs -- pure host logic, no CUDA
 * call (testable without a GPU).  *path: PCD_SOLVER_RESIDENT or PCD_SOLVER_TILED; for the resident path also the rows per
 * CTA (3..9; 1..7 for the exchange-per-phase kernel), the number of CTAs, whether the solve runs on the transposed grid
 * (more than 1024 columns) and whether only the deep-halo kernel applies (8-9 rows per CTA or transposed: NaN holes then
 * go to the large-grid paths).  Any output pointer may be NULL. */
int pcd_solver_plan(int width, int height, int sm_count, int *path, int *rows_per_cta, int *ctas, int *transposed,
                    int *deep_only);

/* ---- row-slab solver for multi-GPU runs (no counterpart in the reference) -------------------
 * One process per GPU owns global rows [row0, row0+rows) of a width x height grid plus GH =
 * pcd_slab_ghost_rows() ghost rows above and below (local row r <-> global row row0-GH+r).  The host layer
 * (poisson_caustic_design_b200/slab.py, torch.distributed) refreshes the ghost rows from the neighbouring ranks --
 * GH rows after every wavefront pass (pcd_slab_pass: up to pcd_slab_sweeps_per_pass() sweeps, the field
 * ping-pongs between two buffers) or one row after every colour phase (pcd_slab_sweep_colour, the NaN-hole path)
 * -- and all-reduces the per-sweep max.  `cuda_stream` is the caller's stream (torch's current stream) so
 * kernels and collectives order naturally. */
typedef struct pcd_slab pcd_slab;
int pcd_slab_ghost_rows(void);
int pcd_slab_sweeps_per_pass(void);
int pcd_slab_create(int width, int height, int row0, int rows, int device, void *cuda_stream, pcd_slab **out);
void pcd_slab_destroy(pcd_slab *s);
int pcd_slab_set_sm_reserve(pcd_slab *s, int n);   /* SMs this slab's pass kernels leave free for the collective's kernels (default 0) */
int pcd_slab_device_ptrs(pcd_slab *s, void **phi0_dev, void **phi1_dev, void **sweep_max_dev);
int pcd_slab_current(const pcd_slab *s);   /* which of the two phi buffers holds the field */
int pcd_slab_error_word(pcd_slab *s, void **err_dev);   /* device address of the int error word (see pcd_slab_peer_status) */
int pcd_slab_has_nan(const pcd_slab *s);   /* D (owned rows and their neighbours) contains NaN */
int pcd_slab_upload(pcd_slab *s, const double *D_rows_with_ghosts, const double *phi_rows_with_ghosts);
int pcd_slab_download(pcd_slab *s, double *phi_owned_rows);
int pcd_slab_sweep_colour(pcd_slab *s, int colour, int slot);
int pcd_slab_pass(pcd_slab *s, int nsweeps, int slot);
/* the same pass over owned rows [row_begin, row_begin+row_count) only, on cuda_stream (NULL = the slab's), without
 * switching buffers; pcd_slab_flip switches them once every part has been issued (exchange/compute overlap) */
int pcd_slab_pass_part(pcd_slab *s, int nsweeps, int slot, int row_begin, int row_count, void *cuda_stream);
int pcd_slab_flip(pcd_slab *s);
int pcd_slab_clear_max(pcd_slab *s, int n_slots);
int pcd_slab_clear_max_range(pcd_slab *s, int first_slot, int n_slots);
/* D and phi of the slab (ghost rows included) from / owned rows of phi back to full width x height DEVICE arrays
 * on the slab's device (device-to-device, on the slab's stream; load waits for it) */
int pcd_slab_load_device(pcd_slab *s, const double *D_full_dev, const double *phi_full_dev);
int pcd_slab_store_device(pcd_slab *s, double *phi_full_dev);
/* Ghost-row exchange FUSED into the pass (no host-side collective on the data path): the kernel stores the
 * pcd_slab_ghost_rows() rows next to a slab edge straight into the neighbour's field (peer memory over NVLink)
 * and raises a sequence flag there; the neighbour's next pass waits for that flag on the device.  Neighbours are
 * attached once: across processes through CUDA IPC handles (export -> send with torch.distributed -> connect),
 * inside one process directly.  side 0 = the slab owning the rows above, 1 = the rows below.  All ranks issue
 * the same sequence of pcd_slab_peer_run calls; pcd_slab_peer_status waits for the stream and reports whether a
 * pass ran into the 3 s limit waiting for a neighbour. */
int pcd_slab_peer_handle_bytes(void);
int pcd_slab_peer_export(pcd_slab *s, unsigned char *handles);
int pcd_slab_peer_connect_ipc(pcd_slab *s, int side, const unsigned char *handles, int peer_row0, int peer_rows);
int pcd_slab_peer_connect_local(pcd_slab *s, int side, pcd_slab *peer);
int pcd_slab_peer_run(pcd_slab *s, int nsweeps, int slot);
int pcd_slab_peer_status(pcd_slab *s, int *timed_out);
/* asynchronous variant for the convergence block: copies the slab's error word (1 = a pass ran into the limit) into
 * the DEVICE double *dst_dev on the slab's stream, so the host layer can fold it into the all-reduce of the maxima
 * and every rank learns about a stalled neighbour in the same block.  The word is cleared by pcd_slab_upload /
 * pcd_slab_load_device (a new solve starts clean). */
int pcd_slab_peer_error_to(pcd_slab *s, double *dst_dev);

/* ---- one Poisson problem on several GPUs of ONE process (the C++ host's --gpus N; torch-free) ------------------
 * The grid is cut into one row slab per entry of `devices` (the same device may appear several times: G slabs on one
 * GPU, the emulation used by single-GPU tests); neighbouring slabs exchange ghost rows inside the persistent pass
 * kernel through peer access.  pcd_multi_solve has the contract of pcd_solver_run with DEVICE arrays on devices[0];
 * pcd_multi_attach installs it as the Poisson solver of a context created on devices[0] (pcd_set_solve_hook), so the
 * reference call it stands behind is poisson_solver as issued at src/caustic_design.cpp:222,311.  Bit-identical to
 * the single-GPU large-grid solver.  NaN holes, slabs thinner than 2 * pcd_slab_ghost_rows() rows, or devices
 * without peer access to each other run on devices[0] alone (same result).  Destroy (or detach) the solver before the context it is attached to. */
typedef struct pcd_multi pcd_multi;
int pcd_multi_create(int width, int height, const int *devices, int n_devices, pcd_multi **out);
void pcd_multi_destroy(pcd_multi *m);
int pcd_multi_device_count(const pcd_multi *m);
int pcd_multi_set_check_every(pcd_multi *m, int sweeps);   /* sweeps between convergence checks (default 64) */
int pcd_multi_solve(pcd_multi *m, const double *D_dev, double *phi_dev, int max_iterations, double convergence_threshold,
                    pcd_solve_info *info /* may be NULL */);
int pcd_multi_attach(pcd_multi *m, pcd_ctx *ctx /* NULL detaches */);

#ifdef __cplusplus
}
#endif
#endif /* PCD_H */
