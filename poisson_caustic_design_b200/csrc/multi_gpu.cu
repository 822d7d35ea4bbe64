// One Poisson problem on several GPUs of THIS process (SURVEY 8e from the C++ host: `caustic_design --gpus N`), no
// torch, no NCCL: the grid is cut into row slabs (sor_slab.cu), one per entry of `devices`; neighbouring slabs are
// attached to each other through peer access and exchange their ghost rows INSIDE the persistent pass kernel (plain
// stores into the neighbour's HBM over NVLink + per-strip sequence flags, sor_tiled.cu).  The host thread only
// launches one kernel per slab and block of `check_every` sweeps and reduces the per-sweep maxima of the slabs (exact:
// max of maxima), so the result is bit-identical to the single-GPU large-grid solver, which tests the stopping rule on
// the same schedule.  The same device may appear several times in `devices` (G slabs on one GPU: the emulation the
// single-GPU tests use); its slabs then share a stream and are advanced pass by pass.
//
// Replaces nothing in the reference (its only parallelism is the thread tiling of src/solver.cpp:73-83); it sits
// behind the same call, poisson_solver(D, phi, ...) as issued at src/caustic_design.cpp:222,311, through
// pcd_set_solve_hook.
#include <cstring>
#include <vector>

#include "sor_common.cuh"

struct pcd_multi {
    int W = 0, H = 0, n = 0;
    int check_every = 64;
    std::vector<int> dev;
    std::vector<pcd_slab *> slabs;
    std::vector<cudaStream_t> streams;          // one per slab; slabs on the same device share one
    std::vector<bool> owns_stream;
    std::vector<unsigned long long *> h_max;    // pinned, 4096 per slab: two blocks in flight (low / high half)
    std::vector<unsigned long long *> d_max;    // the slabs' per-sweep maxima (device)
    std::vector<int *> d_err;                   // the slabs' error words (device)
    std::vector<int *> h_err;                   // pinned, 2 per slab
    std::vector<cudaStream_t> aux;              // per slab: copies a block's maxima out while the next block runs
    std::vector<cudaEvent_t> ev_blk, ev_copied; // 2 per slab
    bool shared_devices = false;
    bool peers = false;                          // every slab thick enough for the fused exchange
    pcd_solver fallback;                         // single-GPU solver on devices[0]: NaN holes, slabs too thin
    bool fallback_ready = false;
    pcd_ctx *attached = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;      // device time of a solve, on streams[0]
};

using namespace pcd;

static int multi_hook(void *user, const double *D_dev, double *phi_dev, int width, int height, int max_iterations, double tol,
                      pcd_solve_info *info) {
    pcd_multi *m = static_cast<pcd_multi *>(user);
    if (width != m->W || height != m->H) { set_error("pcd_multi: attached to a %dx%d context, built for %dx%d", width, height, m->W, m->H); return PCD_ERR_INVALID; }
    return pcd_multi_solve(m, D_dev, phi_dev, max_iterations, tol, info);
}

extern "C" {

int pcd_multi_create(int width, int height, const int *devices, int n_devices, pcd_multi **out) {
    if (!out || !devices || n_devices < 1 || width < 1 || height < n_devices) {
        set_error("pcd_multi_create: bad arguments (%d devices for a %dx%d grid)", n_devices, width, height);
        return PCD_ERR_INVALID;
    }
    *out = nullptr;
    pcd_multi *m = new pcd_multi();
    m->W = width; m->H = height; m->n = n_devices;
    m->dev.assign(devices, devices + n_devices);
    const int GH = pcd_slab_ghost_rows();
    int rc = PCD_OK;
    m->peers = n_devices > 1;
    for (int g = 0; g < n_devices && rc == PCD_OK; ++g) {
        rc = select_device(devices[g]);
        if (rc != PCD_OK) break;
        cudaStream_t st = nullptr;
        bool own = true;
        for (int q = 0; q < g; ++q)
            if (devices[q] == devices[g]) { st = m->streams[q]; own = false; m->shared_devices = true; break; }
        if (own && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { set_error("pcd_multi_create: stream creation failed"); rc = PCD_ERR_CUDA; break; }
        m->streams.push_back(st);
        m->owns_stream.push_back(own);
        const int row0 = (int)((long long)g * height / n_devices), row1 = (int)((long long)(g + 1) * height / n_devices);
        if (row1 - row0 < 2 * GH) m->peers = false;
        pcd_slab *s = nullptr;
        rc = pcd_slab_create(width, height, row0, row1 - row0, devices[g], st, &s);
        if (rc != PCD_OK) break;
        m->slabs.push_back(s);
        void *mx = nullptr;
        pcd_slab_device_ptrs(s, nullptr, nullptr, &mx);
        m->d_max.push_back(static_cast<unsigned long long *>(mx));
        void *ew = nullptr;
        pcd_slab_error_word(s, &ew);
        m->d_err.push_back(static_cast<int *>(ew));
        unsigned long long *h = nullptr;
        int *he = nullptr;
        cudaStream_t ax = nullptr;
        if (cudaMallocHost(&h, sizeof(unsigned long long) * 4096) != cudaSuccess || cudaMallocHost(&he, sizeof(int) * 2) != cudaSuccess ||
            cudaStreamCreateWithFlags(&ax, cudaStreamNonBlocking) != cudaSuccess) {
            set_error("pcd_multi_create: pinned allocation / stream creation failed");
            rc = PCD_ERR_CUDA;
            break;
        }
        he[0] = he[1] = 0;
        m->h_max.push_back(h);
        m->h_err.push_back(he);
        m->aux.push_back(ax);
        for (int i = 0; i < 4; ++i) {
            cudaEvent_t ev = nullptr;
            if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { set_error("pcd_multi_create: event creation failed"); rc = PCD_ERR_CUDA; break; }
            (i < 2 ? m->ev_blk : m->ev_copied).push_back(ev);
        }
    }
    if (rc == PCD_OK && m->peers)
        for (int g = 0; g + 1 < n_devices && rc == PCD_OK && m->peers; ++g) {
            rc = pcd_slab_peer_connect_local(m->slabs[g], 1, m->slabs[g + 1]);
            if (rc == PCD_OK) rc = pcd_slab_peer_connect_local(m->slabs[g + 1], 0, m->slabs[g]);
            if (rc == PCD_ERR_UNSUPPORTED) {   // no peer access between two of the devices: every solve runs on devices[0] alone
                m->peers = false;
                rc = PCD_OK;
            }
        }
    if (rc == PCD_OK) {
        // the full fields live on devices[0]: let every other device read / write them directly where it can (the
        // copies fall back to staging through the host where it cannot)
        for (int g = 1; g < n_devices; ++g) {
            if (devices[g] == devices[0]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[g], devices[0]) == cudaSuccess && can) {
                cudaSetDevice(devices[g]);
                if (cudaDeviceEnablePeerAccess(devices[0], 0) != cudaSuccess) cudaGetLastError();
            }
        }
        cudaGetLastError();
    }
    if (rc != PCD_OK) { pcd_multi_destroy(m); return rc; }
    *out = m;
    return PCD_OK;
}

void pcd_multi_destroy(pcd_multi *m) {
    if (!m) return;
    if (m->attached) pcd_set_solve_hook(m->attached, nullptr, nullptr);
    for (pcd_slab *s : m->slabs) pcd_slab_destroy(s);
    for (size_t g = 0; g < m->streams.size(); ++g)
        if (m->owns_stream[g] && m->streams[g]) { cudaSetDevice(m->dev[g]); cudaStreamDestroy(m->streams[g]); }
    for (unsigned long long *h : m->h_max) cudaFreeHost(h);
    for (int *h : m->h_err) cudaFreeHost(h);
    for (size_t g = 0; g < m->aux.size(); ++g) { cudaSetDevice(m->dev[g]); cudaStreamDestroy(m->aux[g]); }
    for (cudaEvent_t e : m->ev_blk) cudaEventDestroy(e);
    for (cudaEvent_t e : m->ev_copied) cudaEventDestroy(e);
    if (m->fallback_ready) { cudaSetDevice(m->dev[0]); solver_free(&m->fallback); }
    if (m->e0) { cudaSetDevice(m->dev[0]); cudaEventDestroy(m->e0); cudaEventDestroy(m->e1); }
    delete m;
}

int pcd_multi_set_check_every(pcd_multi *m, int sweeps) {
    if (!m || sweeps < 1) { set_error("pcd_multi_set_check_every: bad arguments"); return PCD_ERR_INVALID; }
    const int TS = pcd_slab_sweeps_per_pass();
    m->check_every = sweeps > 2048 ? 2048 : (sweeps + TS - 1) / TS * TS;
    return PCD_OK;
}

int pcd_multi_device_count(const pcd_multi *m) { return m ? m->n : 0; }

// D_dev / phi_dev: width x height arrays on devices[0]; phi in/out (warm start).  Same contract as pcd_solver_run.
int pcd_multi_solve(pcd_multi *m, const double *D_dev, double *phi_dev, int max_iterations, double tol, pcd_solve_info *info) {
    if (!m || !D_dev || !phi_dev) { set_error("pcd_multi_solve: null argument"); return PCD_ERR_INVALID; }
    pcd_solve_info local{};
    if (!info) info = &local;
    *info = pcd_solve_info{};
    info->path = PCD_SOLVER_TILED;
    if (max_iterations <= 0) return PCD_OK;
    const int n = m->n;
    bool nan = false;
    if (m->peers) {
        for (int g = 0; g < n; ++g) PCD_TRY(pcd_slab_load_device(m->slabs[g], D_dev, phi_dev));   // waits for its stream
        for (int g = 0; g < n; ++g) nan = nan || pcd_slab_has_nan(m->slabs[g]) != 0;
    }
    if (!m->peers || nan) {
        // one slab would be thinner than two ghost depths, or D has NaN holes (their neighbour rule needs the masked
        // per-colour kernels): the whole problem runs on devices[0]
        PCD_TRY(select_device(m->dev[0]));
        if (!m->fallback_ready) {
            PCD_TRY(solver_init(&m->fallback, m->W, m->H, m->dev[0], PCD_SOLVER_AUTO, nullptr));
            m->fallback_ready = true;
        }
        return solver_run(&m->fallback, D_dev, phi_dev, max_iterations, tol, info);
    }
    const int TS = pcd_slab_sweeps_per_pass();
    PCD_TRY(select_device(m->dev[0]));
    if (!m->e0) {
        PCD_CUDA(cudaEventCreate(&m->e0));
        PCD_CUDA(cudaEventCreate(&m->e1));
    }
    PCD_CUDA(cudaEventRecord(m->e0, m->streams[0]));
    // Blocks of check_every sweeps; the stopping rule is evaluated ONE BLOCK LATE (block b+1 is queued before the maxima
    // of block b are read; they leave on the slabs' side streams), so no GPU drains between blocks.  Same schedule as
    // the single-GPU large-grid solver (run_tiled) => same bits.
    constexpr int HALF = 2048;
    int launched = 0, done = 0, conv = 0, blk = 0, head = 0, npend = 0;
    int pend_k[2] = {0, 0}, pend_first[2] = {0, 0};
    double last = 0.0;
    for (;;) {
        if (launched < max_iterations && !conv && npend < 2) {
            const int k = max_iterations - launched < m->check_every ? max_iterations - launched : m->check_every;
            const int b = blk & 1, off = b * HALF;
            for (int g = 0; g < n; ++g) PCD_TRY(pcd_slab_clear_max_range(m->slabs[g], off, k));
            if (m->shared_devices) {     // slabs of one device run one after the other: advance all slabs pass by pass
                for (int j = 0; j < k; j += TS)
                    for (int g = 0; g < n; ++g) PCD_TRY(pcd_slab_peer_run(m->slabs[g], k - j < TS ? k - j : TS, off + j));
            } else {                     // one persistent launch per slab for the whole block
                for (int g = 0; g < n; ++g) PCD_TRY(pcd_slab_peer_run(m->slabs[g], k, off));
            }
            info->launches += m->shared_devices ? n * ((k + TS - 1) / TS) : n * (1 + (k % TS ? 1 : 0));
            for (int g = 0; g < n; ++g) {
                PCD_TRY(select_device(m->dev[g]));
                PCD_CUDA(cudaEventRecord(m->ev_blk[2 * g + b], m->streams[g]));
                PCD_CUDA(cudaStreamWaitEvent(m->aux[g], m->ev_blk[2 * g + b], 0));
                PCD_CUDA(cudaMemcpyAsync(m->h_max[g] + off, m->d_max[g] + off, sizeof(unsigned long long) * k, cudaMemcpyDeviceToHost, m->aux[g]));
                PCD_CUDA(cudaMemcpyAsync(m->h_err[g] + b, m->d_err[g], sizeof(int), cudaMemcpyDeviceToHost, m->aux[g]));
                PCD_CUDA(cudaEventRecord(m->ev_copied[2 * g + b], m->aux[g]));
            }
            pend_k[b] = k; pend_first[b] = launched;
            launched += k; ++blk; ++npend;
            continue;
        }
        if (!npend) break;
        const int b = head & 1, off = b * HALF, k = pend_k[b];
        for (int g = 0; g < n; ++g) {
            PCD_TRY(select_device(m->dev[g]));
            PCD_CUDA(cudaEventSynchronize(m->ev_copied[2 * g + b]));
            if (m->h_err[g][b]) { set_error("pcd_multi_solve: slab %d gave up waiting for a neighbour's ghost rows", g); return PCD_ERR_CUDA; }
        }
        for (int j = 0; j < k && !conv; ++j) {
            double mx = 0.0;
            for (int g = 0; g < n; ++g) {
                double v;
                memcpy(&v, &m->h_max[g][off + j], sizeof(double));
                if (v > mx) mx = v;
            }
            if (mx < tol) conv = pend_first[b] + j + 1;
            if (conv || j == k - 1) last = mx;
        }
        done = pend_first[b] + k;
        ++head; --npend;
    }
    for (int g = 0; g < n; ++g) PCD_TRY(pcd_slab_store_device(m->slabs[g], phi_dev));
    for (int g = 0; g < n; ++g) {
        PCD_TRY(select_device(m->dev[g]));
        PCD_CUDA(cudaStreamSynchronize(m->streams[g]));
    }
    PCD_TRY(select_device(m->dev[0]));
    PCD_CUDA(cudaEventRecord(m->e1, m->streams[0]));
    PCD_CUDA(cudaEventSynchronize(m->e1));
    float ms = 0.f;
    PCD_CUDA(cudaEventElapsedTime(&ms, m->e0, m->e1));
    info->sweeps = done;
    info->converged_at = conv;
    info->last_max_update = last;
    info->device_ms = info->kernel_ms = ms;
    return PCD_OK;
}

// Installs the multi-GPU solver as the Poisson solver of `ctx` (which must live on devices[0] and have the same grid);
// NULL ctx detaches.  The context's transport and height iterations are unchanged.
int pcd_multi_attach(pcd_multi *m, pcd_ctx *ctx) {
    if (!m) { set_error("pcd_multi_attach: null solver"); return PCD_ERR_INVALID; }
    if (m->attached) { pcd_set_solve_hook(m->attached, nullptr, nullptr); m->attached = nullptr; }
    if (!ctx) return PCD_OK;
    if (ctx->cfg.device != m->dev[0] || ctx->cfg.res_x != m->W || ctx->cfg.res_y != m->H) {
        set_error("pcd_multi_attach: the context (device %d, %dx%d) does not match the solver (device %d, %dx%d)", ctx->cfg.device,
                  ctx->cfg.res_x, ctx->cfg.res_y, m->dev[0], m->W, m->H);
        return PCD_ERR_INVALID;
    }
    PCD_TRY(pcd_set_solve_hook(ctx, multi_hook, m));
    m->attached = ctx;
    return PCD_OK;
}

}  // extern "C"
