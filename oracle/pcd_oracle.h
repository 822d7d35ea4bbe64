/* TEST INFRASTRUCTURE ONLY -- plain-C restatement ("port") of the reference hot path.
 *
 * Every function cites the file:line of /root/reference it restates.  Pinned bit-for-bit
 * against the reference itself (oracle/_ref, built from the unmodified sources) by
 * tests/test_oracle_pin.py in the build container and against the committed golden
 * vectors (tests/golden/, generated from oracle/_ref) everywhere else.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  It is never part of the product path.
 */
#ifndef PCD_ORACLE_H
#define PCD_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- src/solver.cpp:70-147 ------------------------------------------------------------ */

/* Lexicographic in-place Gauss-Seidel SOR, exactly the reference with max_threads=1
 * (one tile covering the grid).  Returns the number of sweeps executed; *last_max_update
 * receives max|delta| of the last sweep. */
int pcdo_poisson_lex(const double *D, double *phi, int W, int H, int max_iterations,
                     double tol, double *last_max_update);

/* Red-black ordering of the same update ((x+y) even first, then odd), same omega, same
 * stopping rule (max over both colours of one sweep < tol).  After the converged sweep,
 * `extra_sweeps` further sweeps are executed (the CUDA solver tests convergence with a fixed
 * lag).  Returns sweeps executed; *converged_at = number of sweeps up to and
 * including the first one with max|delta| < tol (0 if the cap was hit first). */
int pcdo_poisson_rb(const double *D, double *phi, int W, int H, int max_iterations, double tol,
                    int extra_sweeps, int *converged_at, double *last_max_update);

/* ---- src/utils.cpp --------------------------------------------------------------------- */
void pcdo_subtract_average(double *raster, long n);                                  /* :60-86  */
void pcdo_gradient(const double *g, int W, int H, double *gx, double *gy);           /* :3-20   */
void pcdo_divergence(const double *nx, const double *ny, int W, int H, double *out); /* :22-39  */
void pcdo_scale_matrix_proportional(const double *m, long n, double lo, double hi, double *out); /* :88-129 */
double pcdo_bilinear(const double *img, int W, int H, double x, double y);  /* src/caustic_design.cpp:156-188 */

/* ---- class Caustic_design (src/caustic_design.h:7-66) ----------------------------------- */
typedef struct pcdo_design pcdo_design;

/* solver_mode: 0 = lexicographic (the reference), 1 = red-black */
pcdo_design *pcdo_create(int mesh_nx, int mesh_ny, int res_x, int res_y, double width, double height,
                         double focal_l, double thickness, int solver_mode);
void pcdo_destroy(pcdo_design *d);
void pcdo_initialize_solvers(pcdo_design *d, const double *image);      /* src/caustic_design.cpp:334-364 */
/* returns the step size, NaN with *miss=1 when a raster sample hits no triangle (the
 * reference exit(0)s there, src/mesh.cpp:276-281) */
double pcdo_perform_transport_iteration(pcdo_design *d, int *miss);     /* :190-266 */
int pcdo_perform_height_map_iteration(pcdo_design *d, int itr);         /* :269-332, nonzero on miss */
int pcdo_last_sweeps(const pcdo_design *d);
/* field ids = include/pcd.h enum pcd_field; returns element count (dst may be NULL) */
long pcdo_get_field(const pcdo_design *d, int field, double *dst);
int pcdo_set_field(pcdo_design *d, int field, const double *src);
/* src/mesh.cpp:348-409; returns the number of points produced (V unless a query missed) */
long pcdo_inverted_transport_map(pcdo_design *d, double *out_x, double *out_y);
/* stage entry points used by per-stage tests (operate on the design's current state) */
void pcdo_stage_errors(pcdo_design *d);                  /* caustic_design.cpp:194-209 */
int pcdo_stage_raster(pcdo_design *d);                   /* :212-213 (no mean removal) */
double pcdo_stage_step(pcdo_design *d);                  /* :225-265 from d->phi */

#ifdef __cplusplus
}
#endif
#endif
