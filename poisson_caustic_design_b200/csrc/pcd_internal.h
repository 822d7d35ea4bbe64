// Internal declarations shared by the .cu translation units of libpcd_b200.so.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>

#include "pcd.h"

namespace pcd {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<long long> g_launches;

#define PCD_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            ::pcd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                             __LINE__);                                                         \
            return PCD_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

#define PCD_TRY(expr)                \
    do {                             \
        int _s = (expr);             \
        if (_s != PCD_OK) return _s; \
    } while (0)

// counts the launch and checks the launch error
#define PCD_LAUNCHED()                                                                      \
    do {                                                                                    \
        ::pcd::g_launches.fetch_add(1, std::memory_order_relaxed);                          \
        cudaError_t _e = cudaGetLastError();                                                \
        if (_e != cudaSuccess) {                                                            \
            ::pcd::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),    \
                             __FILE__, __LINE__);                                           \
            return PCD_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

int select_device(int device);

// ---- solver -----------------------------------------------------------------------------------
// omega = 2/(1+pi/W) with the reference's truncated pi, src/solver.cpp:71
inline double sor_omega(int W) { return 2.0 / (1.0 + 3.14159265 / W); }

}  // namespace pcd

struct pcd_solver {
    int W = 0, H = 0, device = 0;
    int path_req = PCD_SOLVER_AUTO, path_used = PCD_SOLVER_STREAMING;
    int check_lag = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    double *D = nullptr, *phi = nullptr;  // owned device arrays (standalone use)
    bool own_fields = false;
    unsigned long long *sweep_max = nullptr;  // device: per-sweep max|delta| bit patterns (ring)
    unsigned long long *h_sweep_max = nullptr;  // pinned host mirror
    int ring = 0;
    unsigned char *mask = nullptr;  // device: 4-bit neighbour masks, only when D has NaN holes
    int *d_flags = nullptr;         // device scratch: [0] = D has NaN
    int *h_flags = nullptr;         // pinned
    // resident path
    void *res_state = nullptr;      // device: control block of the persistent kernel
    void *h_res_state = nullptr;    // pinned mirror
    double *halo = nullptr;         // device: boundary-row exchange buffers
    double *phi_alt = nullptr;      // device: ping-pong partner of phi on the tiled path
    unsigned *wave_ctl = nullptr;   // device: per-CTA sequence words + error word of the persistent wavefront kernel
    unsigned wave_seq = 0;          // sequence number of the last pass it executed
    double *d_split = nullptr;      // parity-split copy of D for the TMA staging of the wavefront kernel (sor_tiled.cu)
    alignas(64) unsigned char dmap[128] = {};   // its tensor map (CUtensorMap)
    int dmap_state = 0;             // 0 undecided, 1 in use, -1 not available (cp.async staging)
    unsigned long long *wave_trace = nullptr;   // PCD_WAVE_TRACE diagnostics (timestamps of the last launch)
    int wave_trace_npass = 0;
    cudaStream_t aux_stream = nullptr;                         // copies a block's maxima out while the next block runs
    cudaEvent_t ev_blk[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    void *dct_state = nullptr;      // opt-in direct backend (dct_solver.cu): DCT matrices, eigenvalues, temporaries
    int res_ctas = 0, res_threads = 0, res_rows_per_cta = 0, res_n_big = 0;
    bool res_tr = false;              // resident solve on the TRANSPOSED grid (W > 1024 columns but H <= 1024 rows)
    double *tr_D = nullptr, *tr_phi = nullptr;   // its transposed copies of D and phi (H x W doubles each, allocated on first use)
    int res_exchange = 0;   // resident kernel of the last run: 0 none, 1 exchange per colour phase, 2 per sweep (deep halos)
    int res_pairs = 0;   // CTA-pair (cluster) launch of the resident kernel: 0 undecided, 1 in use, -1 not available
    size_t res_smem = 0;
    int sm_count = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr;
};

namespace pcd {
// run_resident's way of saying "not on chip after all" (never leaves solver_run): strips of 8-9 rows exist only in the
// deep-halo kernel, which does not take NaN holes
constexpr int PCD_RES_FALLBACK = 1000;
int resident_plan(pcd_solver *s);   // 1 when the grid fits the on-chip path (fills s->res_*); host logic only (sor_resident.cu)
int solver_init(pcd_solver *s, int W, int H, int device, int path, cudaStream_t stream);
void solver_free(pcd_solver *s);
// D_dev / phi_dev: device arrays of W*H doubles; phi in/out
int solver_run(pcd_solver *s, const double *D_dev, double *phi_dev, int max_iterations, double tol,
               pcd_solve_info *info);
}  // namespace pcd

struct pcd_ctx {
    pcd_config cfg{};
    int V = 0, T = 0;
    long N = 0;
    cudaStream_t stream = nullptr;
    bool initialized = false;
    // mesh (SoA)
    double *tx = nullptr, *ty = nullptr, *tz = nullptr, *sx = nullptr, *sy = nullptr, *sz = nullptr;
    // fields
    double *pixels = nullptr, *target_areas = nullptr, *errors = nullptr, *raster = nullptr, *phi = nullptr,
           *h = nullptr, *vgx = nullptr, *vgy = nullptr, *normals_x = nullptr, *normals_y = nullptr,
           *norm_x = nullptr, *norm_y = nullptr, *divergence = nullptr, *inv_x = nullptr, *inv_y = nullptr,
           *hv = nullptr;
    // sample lattices: domain nodes (src/mesh.cpp:241-253) and inverse-map queries (:356-360)
    double *xs = nullptr, *ys = nullptr, *qxs = nullptr, *qys = nullptr;
    int *owner = nullptr;      // N: owning triangle of each domain sample in the deformed (target) mesh
    int *owner_src = nullptr;  // N: same for the regular (source) mesh, computed once
    bool owner_src_valid = false;
    int *owner_v = nullptr;    // V: owning triangle of each inverse-map query
    // reductions
    double *partials = nullptr;  // device scratch for deterministic two-stage sums
    int n_partials = 0;
    double *d_scalars = nullptr;  // device: small result slots
    double *h_scalars = nullptr;  // pinned mirror
    unsigned long long *d_bits = nullptr;  // device: atomicMax/Min slots
    int *d_flags = nullptr, *h_flags = nullptr;
    pcd_solver solver;
    pcd_solve_info last{};
    pcd_solve_info totals{};
    cudaEvent_t events[8] = {};
    void *l2_scratch = nullptr;
    double transport_tol = 0.0000001;  // src/caustic_design.cpp:222
    double height_tol = 1e-9;          // the reference uses 1e-8 (:311); see pcd_set_tolerances in pcd.h
    pcd_solve_hook solve_hook = nullptr;  // external Poisson solver (pcd_set_solve_hook)
    void *solve_hook_user = nullptr;
};

namespace pcd {
// solve on the context's grid and fold the result into ctx->last / ctx->totals
inline int ctx_solve(pcd_ctx *c, const double *D_dev, double *x_dev, double tol) {
    int rc;
    if (c->solve_hook) {  // external (multi-GPU) solver: sees finished inputs, leaves a finished phi
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { set_error("stream synchronize before the solve hook: %s", cudaGetErrorString(e)); return PCD_ERR_CUDA; }
        c->last = pcd_solve_info{};
        rc = c->solve_hook(c->solve_hook_user, D_dev, x_dev, c->cfg.res_x, c->cfg.res_y, 100000, tol, &c->last);
        if (rc != PCD_OK) set_error("the solve hook failed with status %d", rc);
    } else {
        rc = solver_run(&c->solver, D_dev, x_dev, 100000, tol, &c->last);  // cap 100000: src/caustic_design.cpp:222,311
    }
    c->totals.sweeps += c->last.sweeps;
    c->totals.launches += c->last.launches;
    c->totals.kernel_ms += c->last.kernel_ms;
    c->totals.device_ms += c->last.device_ms;
    c->totals.path = c->last.path;
    return rc;
}
// init_kernels.cu
int k_init(pcd_ctx *c, const double *image_host);
// transport_kernels.cu
int k_errors(pcd_ctx *c);
int k_raster_target(pcd_ctx *c);            // errors -> raster (nodal samples of the deformed mesh)
int k_subtract_average(pcd_ctx *c, double *grid);
int k_step(pcd_ctx *c, double *step_out_host);  // gradient+bilinear+step_grid+max displacement
int k_gradient(pcd_ctx *c, const double *grid, double *gx, double *gy);
// height_kernels.cu
int k_inverse_map(pcd_ctx *c);              // -> inv_x, inv_y
int k_height_iteration(pcd_ctx *c, double *update_sum_host);
// shared helpers
int check_miss(pcd_ctx *c, const char *what);
}  // namespace pcd
