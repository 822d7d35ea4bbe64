// C ABI of libpcd_b200.so (include/pcd.h).  Thin: argument checks, device memory, stage order.
#include <cstdarg>
#include <cstring>
#include <new>

#include "common.cuh"

namespace pcd {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int select_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_error("no usable CUDA device (%s); this library has no CPU path", e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
        cudaGetLastError();
        return PCD_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) {
        set_error("device %d out of range (%d devices)", device, n);
        return PCD_ERR_INVALID;
    }
    PCD_CUDA(cudaSetDevice(device));
    return PCD_OK;
}

template <typename T>
static int dmalloc(T **p, size_t n) {
    PCD_CUDA(cudaMalloc((void **)p, sizeof(T) * (n ? n : 1)));
    PCD_CUDA(cudaMemset(*p, 0, sizeof(T) * (n ? n : 1)));
    return PCD_OK;
}

static double *field_ptr(const pcd_ctx *c, int field, long *n) {
    const long N = c->N, V = c->V;
    switch (field) {
        case PCD_FIELD_PHI: *n = N; return c->phi;
        case PCD_FIELD_H: *n = N; return c->h;
        case PCD_FIELD_RASTER: *n = N; return c->raster;
        case PCD_FIELD_PIXELS: *n = N; return c->pixels;
        case PCD_FIELD_DIVERGENCE: *n = N; return c->divergence;
        case PCD_FIELD_NORM_X: *n = N; return c->norm_x;
        case PCD_FIELD_NORM_Y: *n = N; return c->norm_y;
        case PCD_FIELD_GRADIENT_X: *n = N; return nullptr;  // on demand
        case PCD_FIELD_GRADIENT_Y: *n = N; return nullptr;
        case PCD_FIELD_ERRORS: *n = V; return c->errors;
        case PCD_FIELD_TARGET_AREAS: *n = V; return c->target_areas;
        case PCD_FIELD_VERTEX_GRADIENT_X: *n = V; return c->vgx;
        case PCD_FIELD_VERTEX_GRADIENT_Y: *n = V; return c->vgy;
        case PCD_FIELD_NORMALS_X: *n = V; return c->normals_x;
        case PCD_FIELD_NORMALS_Y: *n = V; return c->normals_y;
        case PCD_FIELD_TARGET_X: *n = V; return c->tx;
        case PCD_FIELD_TARGET_Y: *n = V; return c->ty;
        case PCD_FIELD_TARGET_Z: *n = V; return c->tz;
        case PCD_FIELD_SOURCE_X: *n = V; return c->sx;
        case PCD_FIELD_SOURCE_Y: *n = V; return c->sy;
        case PCD_FIELD_SOURCE_Z: *n = V; return c->sz;
        default: *n = -1; return nullptr;
    }
}

static void ctx_free(pcd_ctx *c) {
    if (!c) return;
    double **ds[] = {&c->tx, &c->ty, &c->tz, &c->sx, &c->sy, &c->sz, &c->pixels, &c->target_areas, &c->errors, &c->raster,
                     &c->phi, &c->h, &c->vgx, &c->vgy, &c->normals_x, &c->normals_y, &c->norm_x, &c->norm_y, &c->divergence,
                     &c->inv_x, &c->inv_y, &c->hv, &c->xs, &c->ys, &c->qxs, &c->qys, &c->partials, &c->d_scalars};
    for (double **p : ds) cudaFree(*p);
    cudaFree(c->owner); cudaFree(c->owner_src); cudaFree(c->owner_v); cudaFree(c->d_bits); cudaFree(c->d_flags);
    cudaFreeHost(c->h_scalars); cudaFreeHost(c->h_flags);
    solver_free(&c->solver);
    for (cudaEvent_t e : c->events) if (e) cudaEventDestroy(e);
    cudaFree(c->l2_scratch);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

static int ctx_alloc(pcd_ctx *c) {
    const size_t N = (size_t)c->N, V = (size_t)c->V;
    PCD_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    double **vs[] = {&c->tx, &c->ty, &c->tz, &c->sx, &c->sy, &c->sz, &c->target_areas, &c->errors, &c->vgx, &c->vgy,
                     &c->normals_x, &c->normals_y, &c->inv_x, &c->inv_y, &c->hv};
    for (double **p : vs) PCD_TRY(dmalloc(p, V));
    double **gs[] = {&c->pixels, &c->raster, &c->phi, &c->h, &c->norm_x, &c->norm_y, &c->divergence};
    for (double **p : gs) PCD_TRY(dmalloc(p, N));
    PCD_TRY(dmalloc(&c->xs, (size_t)c->cfg.res_x));
    PCD_TRY(dmalloc(&c->ys, (size_t)c->cfg.res_y));
    PCD_TRY(dmalloc(&c->qxs, (size_t)c->cfg.mesh_res_x));
    PCD_TRY(dmalloc(&c->qys, (size_t)c->cfg.mesh_res_y));
    PCD_TRY(dmalloc(&c->owner, N));
    PCD_TRY(dmalloc(&c->owner_src, N));
    PCD_TRY(dmalloc(&c->owner_v, V));
    c->n_partials = 2 * RED_BLOCKS;
    PCD_TRY(dmalloc(&c->partials, (size_t)c->n_partials));
    PCD_TRY(dmalloc(&c->d_scalars, 8));
    PCD_TRY(dmalloc(&c->d_bits, 8));
    PCD_TRY(dmalloc(&c->d_flags, 4));
    PCD_CUDA(cudaMallocHost(&c->h_scalars, sizeof(double) * 8));
    PCD_CUDA(cudaMallocHost(&c->h_flags, sizeof(int) * 4));
    for (cudaEvent_t &e : c->events) PCD_CUDA(cudaEventCreate(&e));
    PCD_TRY(solver_init(&c->solver, c->cfg.res_x, c->cfg.res_y, c->cfg.device, c->cfg.solver_path, c->stream));
    // the zero-fills above ran on the legacy default stream, which non-blocking streams do not wait for
    PCD_CUDA(cudaDeviceSynchronize());
    return PCD_OK;
}

}  // namespace pcd

using namespace pcd;

extern "C" {

int pcd_abi_version(void) { return PCD_ABI_VERSION; }
const char *pcd_last_error(void) { return g_err; }
long long pcd_launch_count(void) { return g_launches.load(); }

int pcd_device_count(int *count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    if (count) *count = n;
    return PCD_OK;
}

int pcd_create(const pcd_config *cfg, pcd_ctx **out) {
    if (!cfg || !out) { set_error("pcd_create: null argument"); return PCD_ERR_INVALID; }
    *out = nullptr;
    if (cfg->mesh_res_x < 2 || cfg->mesh_res_y < 2 || cfg->res_x < 2 || cfg->res_y < 2 || !(cfg->width > 0) || !(cfg->height > 0)) {
        set_error("pcd_create: mesh %dx%d / domain %dx%d / size %gx%g is not a valid setup", cfg->mesh_res_x, cfg->mesh_res_y,
                  cfg->res_x, cfg->res_y, cfg->width, cfg->height);
        return PCD_ERR_INVALID;
    }
    PCD_TRY(select_device(cfg->device));
    pcd_ctx *c = new (std::nothrow) pcd_ctx();
    if (!c) { set_error("out of host memory"); return PCD_ERR_INVALID; }
    c->cfg = *cfg;
    c->V = cfg->mesh_res_x * cfg->mesh_res_y;
    c->T = 2 * (cfg->mesh_res_x - 1) * (cfg->mesh_res_y - 1);
    c->N = (long)cfg->res_x * cfg->res_y;
    int rc = ctx_alloc(c);
    if (rc != PCD_OK) { ctx_free(c); return rc; }
    *out = c;
    return PCD_OK;
}

void pcd_destroy(pcd_ctx *ctx) {
    if (ctx) {
        cudaSetDevice(ctx->cfg.device);
        ctx_free(ctx);
    }
}

#define NEED_CTX(c)                                                             \
    do {                                                                        \
        if (!(c)) { set_error("null context"); return PCD_ERR_INVALID; }        \
        PCD_TRY(select_device((c)->cfg.device));                                \
    } while (0)
#define NEED_INIT(c)                                                                                          \
    do {                                                                                                      \
        NEED_CTX(c);                                                                                          \
        if (!(c)->initialized) { set_error("pcd_initialize_solvers has not been called"); return PCD_ERR_STATE; } \
    } while (0)

int pcd_initialize_solvers(pcd_ctx *ctx, const double *image) {
    NEED_CTX(ctx);
    if (!image) { set_error("null image"); return PCD_ERR_INVALID; }
    PCD_TRY(k_init(ctx, image));
    ctx->initialized = true;
    return PCD_OK;
}

int pcd_set_solve_hook(pcd_ctx *ctx, pcd_solve_hook hook, void *user) {
    NEED_CTX(ctx);
    ctx->solve_hook = hook;
    ctx->solve_hook_user = hook ? user : nullptr;
    return PCD_OK;
}

int pcd_set_tolerances(pcd_ctx *ctx, double transport_tol, double height_tol) {
    NEED_CTX(ctx);
    if (transport_tol > 0.0) ctx->transport_tol = transport_tol;
    if (height_tol > 0.0) ctx->height_tol = height_tol;
    return PCD_OK;
}

int pcd_stage_errors(pcd_ctx *ctx) { NEED_INIT(ctx); return k_errors(ctx); }
int pcd_stage_raster(pcd_ctx *ctx) {
    NEED_INIT(ctx);
    PCD_TRY(k_raster_target(ctx));
    return check_miss(ctx, "target raster");
}
int pcd_stage_subtract_average(pcd_ctx *ctx) { NEED_INIT(ctx); return k_subtract_average(ctx, ctx->raster); }
int pcd_stage_solve_transport(pcd_ctx *ctx) {
    NEED_INIT(ctx);
    return ctx_solve(ctx, ctx->raster, ctx->phi, ctx->transport_tol);  // caustic_design.cpp:222
}
int pcd_stage_step(pcd_ctx *ctx, double *step_out) {
    NEED_INIT(ctx);
    double s = 0.0;
    PCD_TRY(k_step(ctx, &s));
    if (step_out) *step_out = s;
    return PCD_OK;
}

int pcd_perform_transport_iteration(pcd_ctx *ctx, double *step_out) {
    NEED_INIT(ctx);
    PCD_TRY(k_errors(ctx));
    PCD_TRY(k_raster_target(ctx));
    PCD_TRY(check_miss(ctx, "target raster"));
    PCD_TRY(k_subtract_average(ctx, ctx->raster));
    PCD_TRY(ctx_solve(ctx, ctx->raster, ctx->phi, ctx->transport_tol));
    double s = 0.0;
    PCD_TRY(k_step(ctx, &s));
    if (step_out) *step_out = s;
    return PCD_OK;
}

int pcd_run_transport(pcd_ctx *ctx, int max_iters, double conv_tres, int *iters_out, double *steps_out) {
    NEED_INIT(ctx);
    int it = 0;
    for (; it < max_iters;) {  // main.cpp:243-256
        double step = 0.0;
        PCD_TRY(pcd_perform_transport_iteration(ctx, &step));
        if (steps_out) steps_out[it] = step;
        ++it;
        if (step < conv_tres) break;
    }
    if (iters_out) *iters_out = it;
    return PCD_OK;
}

int pcd_perform_height_map_iteration(pcd_ctx *ctx, int itr, double *update_sum_out) {
    NEED_INIT(ctx);
    (void)itr;
    return k_height_iteration(ctx, update_sum_out);
}

int pcd_field_size(const pcd_ctx *ctx, int field, long *n_out) {
    if (!ctx || !n_out) { set_error("null argument"); return PCD_ERR_INVALID; }
    long n;
    field_ptr(ctx, field, &n);
    if (n < 0) { set_error("unknown field %d", field); return PCD_ERR_INVALID; }
    *n_out = n;
    return PCD_OK;
}

int pcd_get_field(pcd_ctx *ctx, int field, double *dst) {
    NEED_CTX(ctx);
    if (!dst) { set_error("null destination"); return PCD_ERR_INVALID; }
    long n;
    double *p = field_ptr(ctx, field, &n);
    if (n < 0) { set_error("unknown field %d", field); return PCD_ERR_INVALID; }
    if (field == PCD_FIELD_GRADIENT_X || field == PCD_FIELD_GRADIENT_Y) {
        // calculate_gradient(phi), src/utils.cpp:3-20; norm_x/norm_y are free before the height stage
        // but must survive after it, so a temporary pair is used
        double *gx = nullptr, *gy = nullptr;
        PCD_CUDA(cudaMalloc(&gx, sizeof(double) * n));
        PCD_CUDA(cudaMalloc(&gy, sizeof(double) * n));
        int rc = k_gradient(ctx, ctx->phi, gx, gy);
        if (rc == PCD_OK) {
            cudaError_t e = cudaMemcpyAsync(dst, field == PCD_FIELD_GRADIENT_X ? gx : gy, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { set_error("gradient download failed: %s", cudaGetErrorString(e)); rc = PCD_ERR_CUDA; }
        }
        cudaFree(gx); cudaFree(gy);
        return rc;
    }
    PCD_CUDA(cudaMemcpyAsync(dst, p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    PCD_CUDA(cudaStreamSynchronize(ctx->stream));
    return PCD_OK;
}

int pcd_set_field(pcd_ctx *ctx, int field, const double *src) {
    NEED_CTX(ctx);
    if (!src) { set_error("null source"); return PCD_ERR_INVALID; }
    long n;
    double *p = field_ptr(ctx, field, &n);
    if (n < 0 || !p) { set_error("field %d is not writable", field); return PCD_ERR_INVALID; }
    PCD_CUDA(cudaMemcpyAsync(p, src, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    PCD_CUDA(cudaStreamSynchronize(ctx->stream));
    // the owner map of the source mesh is cached across height iterations (the reference never moves it): a caller
    // that does move it gets a fresh map
    if (field == PCD_FIELD_SOURCE_X || field == PCD_FIELD_SOURCE_Y) ctx->owner_src_valid = false;
    return PCD_OK;
}

int pcd_inverted_transport_map(pcd_ctx *ctx, double *out_x, double *out_y) {
    NEED_INIT(ctx);
    if (!out_x || !out_y) { set_error("null destination"); return PCD_ERR_INVALID; }
    PCD_TRY(k_inverse_map(ctx));
    PCD_TRY(check_miss(ctx, "inverse transport map"));
    PCD_CUDA(cudaMemcpyAsync(out_x, ctx->inv_x, sizeof(double) * ctx->V, cudaMemcpyDeviceToHost, ctx->stream));
    PCD_CUDA(cudaMemcpyAsync(out_y, ctx->inv_y, sizeof(double) * ctx->V, cudaMemcpyDeviceToHost, ctx->stream));
    PCD_CUDA(cudaStreamSynchronize(ctx->stream));
    return PCD_OK;
}

int pcd_last_solve_info(const pcd_ctx *ctx, pcd_solve_info *info) {
    if (!ctx || !info) { set_error("null argument"); return PCD_ERR_INVALID; }
    *info = ctx->last;
    return PCD_OK;
}

int pcd_resident_exchange(const pcd_ctx *ctx) { return ctx ? ctx->solver.res_exchange : -1; }

int pcd_solve_totals(pcd_ctx *ctx, pcd_solve_info *totals, int reset) {
    if (!ctx) { set_error("null argument"); return PCD_ERR_INVALID; }
    if (totals) *totals = ctx->totals;
    if (reset) ctx->totals = pcd_solve_info{};
    return PCD_OK;
}

int pcd_event_record(pcd_ctx *ctx, int slot) {
    NEED_CTX(ctx);
    if (slot < 0 || slot >= 8) { set_error("event slot %d out of range", slot); return PCD_ERR_INVALID; }
    PCD_CUDA(cudaEventRecord(ctx->events[slot], ctx->stream));
    return PCD_OK;
}

int pcd_event_elapsed_ms(pcd_ctx *ctx, int slot_start, int slot_stop, double *ms_out) {
    NEED_CTX(ctx);
    if (slot_start < 0 || slot_start >= 8 || slot_stop < 0 || slot_stop >= 8 || !ms_out) { set_error("bad event slots"); return PCD_ERR_INVALID; }
    PCD_CUDA(cudaEventSynchronize(ctx->events[slot_stop]));
    float ms = 0.f;
    PCD_CUDA(cudaEventElapsedTime(&ms, ctx->events[slot_start], ctx->events[slot_stop]));
    *ms_out = ms;
    return PCD_OK;
}

int pcd_flush_l2(pcd_ctx *ctx) {
    NEED_CTX(ctx);
    const size_t bytes = 256u << 20;
    if (!ctx->l2_scratch) PCD_CUDA(cudaMalloc(&ctx->l2_scratch, bytes));
    PCD_CUDA(cudaMemsetAsync(ctx->l2_scratch, 0x5a, bytes, ctx->stream));
    return PCD_OK;
}

// ---- poisson_solver ------------------------------------------------------------------------------
int pcd_solver_create(int width, int height, int device, int solver_path, pcd_solver **out) {
    if (!out) { set_error("null argument"); return PCD_ERR_INVALID; }
    *out = nullptr;
    PCD_TRY(select_device(device));
    pcd_solver *s = new (std::nothrow) pcd_solver();
    if (!s) { set_error("out of host memory"); return PCD_ERR_INVALID; }
    int rc = solver_init(s, width, height, device, solver_path, nullptr);
    if (rc == PCD_OK) {
        const size_t n = (size_t)width * height;
        cudaError_t e = cudaMalloc(&s->D, sizeof(double) * n);
        if (e == cudaSuccess) e = cudaMalloc(&s->phi, sizeof(double) * n);
        if (e == cudaSuccess) e = cudaMemset(s->D, 0, sizeof(double) * n);
        if (e == cudaSuccess) e = cudaMemset(s->phi, 0, sizeof(double) * n);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();  // zero-fills ran on the legacy default stream
        s->own_fields = true;
        if (e != cudaSuccess) { set_error("solver allocation failed: %s", cudaGetErrorString(e)); rc = PCD_ERR_CUDA; }
    }
    if (rc != PCD_OK) { solver_free(s); delete s; return rc; }
    *out = s;
    return PCD_OK;
}

void pcd_solver_destroy(pcd_solver *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    solver_free(s);
    delete s;
}

int pcd_solver_upload(pcd_solver *s, const double *D, const double *phi) {
    if (!s) { set_error("null solver"); return PCD_ERR_INVALID; }
    PCD_TRY(select_device(s->device));
    const size_t bytes = sizeof(double) * (size_t)s->W * s->H;
    if (D) PCD_CUDA(cudaMemcpyAsync(s->D, D, bytes, cudaMemcpyHostToDevice, s->stream));
    if (phi) PCD_CUDA(cudaMemcpyAsync(s->phi, phi, bytes, cudaMemcpyHostToDevice, s->stream));
    PCD_CUDA(cudaStreamSynchronize(s->stream));
    return PCD_OK;
}

int pcd_solver_download(pcd_solver *s, double *phi) {
    if (!s || !phi) { set_error("null argument"); return PCD_ERR_INVALID; }
    PCD_TRY(select_device(s->device));
    PCD_CUDA(cudaMemcpyAsync(phi, s->phi, sizeof(double) * (size_t)s->W * s->H, cudaMemcpyDeviceToHost, s->stream));
    PCD_CUDA(cudaStreamSynchronize(s->stream));
    return PCD_OK;
}

int pcd_solver_load_device(pcd_solver *s, const double *D_dev, const double *phi_dev) {
    if (!s) { set_error("null solver"); return PCD_ERR_INVALID; }
    PCD_TRY(select_device(s->device));
    const size_t bytes = sizeof(double) * (size_t)s->W * s->H;
    if (D_dev) PCD_CUDA(cudaMemcpyAsync(s->D, D_dev, bytes, cudaMemcpyDeviceToDevice, s->stream));
    if (phi_dev) PCD_CUDA(cudaMemcpyAsync(s->phi, phi_dev, bytes, cudaMemcpyDeviceToDevice, s->stream));
    PCD_CUDA(cudaStreamSynchronize(s->stream));
    return PCD_OK;
}

int pcd_solver_store_device(pcd_solver *s, double *phi_dev) {
    if (!s || !phi_dev) { set_error("null argument"); return PCD_ERR_INVALID; }
    PCD_TRY(select_device(s->device));
    PCD_CUDA(cudaMemcpyAsync(phi_dev, s->phi, sizeof(double) * (size_t)s->W * s->H, cudaMemcpyDeviceToDevice, s->stream));
    PCD_CUDA(cudaStreamSynchronize(s->stream));
    return PCD_OK;
}

int pcd_solver_set_check_lag(pcd_solver *s, int check_lag) {
    if (!s) { set_error("null solver"); return PCD_ERR_INVALID; }
    // the pinned mirror of the per-sweep maxima holds 4096 entries and the resident path reads lag + 2 of them
    s->check_lag = check_lag < 0 ? 0 : (check_lag > 4094 ? 4094 : check_lag);
    return PCD_OK;
}

int pcd_solver_run(pcd_solver *s, int max_iterations, double convergence_threshold, pcd_solve_info *info) {
    if (!s) { set_error("null solver"); return PCD_ERR_INVALID; }
    return solver_run(s, s->D, s->phi, max_iterations, convergence_threshold, info);
}

int pcd_solver_path_used(const pcd_solver *s) { return s ? s->path_used : -1; }

int pcd_solver_resident_exchange(const pcd_solver *s) { return s ? s->res_exchange : -1; }

int pcd_solver_plan(int width, int height, int sm_count, int *path, int *rows_per_cta, int *ctas, int *transposed,
                    int *deep_only) {
    if (width < 1 || height < 1 || height > 65535 || sm_count < 1) { set_error("pcd_solver_plan: invalid argument"); return PCD_ERR_INVALID; }
    pcd_solver s;
    s.W = width; s.H = height; s.sm_count = sm_count;
    const int fits = pcd::resident_plan(&s);
    if (path) *path = fits ? PCD_SOLVER_RESIDENT : PCD_SOLVER_TILED;
    if (rows_per_cta) *rows_per_cta = fits ? s.res_rows_per_cta : 0;
    if (ctas) *ctas = fits ? s.res_ctas : 0;
    if (transposed) *transposed = fits && s.res_tr ? 1 : 0;
    if (deep_only) *deep_only = fits && (s.res_tr || s.res_rows_per_cta > 7) ? 1 : 0;
    return PCD_OK;
}

int pcd_poisson_solver(const double *D, double *phi, int width, int height, int max_iterations,
                       double convergence_threshold, int device, pcd_solve_info *info) {
    if (!D || !phi) { set_error("null field"); return PCD_ERR_INVALID; }
    pcd_solver *s = nullptr;
    PCD_TRY(pcd_solver_create(width, height, device, PCD_SOLVER_AUTO, &s));
    int rc = pcd_solver_upload(s, D, phi);
    if (rc == PCD_OK) rc = pcd_solver_run(s, max_iterations, convergence_threshold, info);
    if (rc == PCD_OK) rc = pcd_solver_download(s, phi);
    pcd_solver_destroy(s);
    return rc;
}

}  // extern "C"
