// K-SOR: red-black SOR for the 5-point Poisson problem, fp64, sm_100a.
//
// Restates src/solver.cpp:12-61 (patial_relax) + :70-147 (poisson_solver) of the reference with a
// red-black ordering ((x+y) even first, then odd): same omega (:71), same per-cell update (:47),
// same neighbour rule (:29-44: a neighbour counts when it is inside the grid and its D is not NaN),
// same stopping rule (:50-53,:142: max|delta| of a full sweep < tol), phi warm-started in place.
// The expression order of the update is kept literally and the library is compiled with
// -fmad=false, so the field after n sweeps is bit-identical to the red-black restatement
// kept under oracle/ (test infrastructure; nothing here includes or links it).
//
// Three kernel families behind one entry point (solver_run), all producing the same bits:
//   * streaming (this file): one launch per colour, phi/D streamed through L2/HBM; any grid size; with per-cell
//                 neighbour masks it is also the path for D with NaN holes.
//   * resident  (sor_resident.cu): ONE persistent kernel per solve for grids that fit on chip (one side <= 1024, the
//                 other <= 1332; wide grids run transposed).  Each CTA (one per SM, launched as clusters of two) owns a slab
//                 of rows; phi of a thread's column pair lives in registers for the whole solve (D in registers or shared
//                 memory), only the left/right neighbour goes through shared memory; slab boundary rows travel as 16-byte
//                 flag-in-data messages through L2 (DSMEM inside a pair), once per sweep (deep halos) or once per colour
//                 phase, polled by the consuming thread; no grid-wide and no CTA-wide barrier in the sweep loop; per-sweep
//                 verdicts are published with one atomic per CTA and acted on with a fixed lag.
//   * wavefront (sor_tiled.cu): temporal blocking for larger grids and multi-GPU slabs, one persistent launch per block of
//                 sweeps between two convergence tests.
// plus the opt-in direct backend (dct_solver.cu).
#include <cstdlib>
#include <cstring>

#include "sor_common.cuh"

namespace pcd {

// ------------------------------------------------------------------------------------------------
// neighbour masks (bit0 left, bit1 up, bit2 right, bit3 down), only needed when D has NaN holes
// ------------------------------------------------------------------------------------------------
__global__ void detect_nan_kernel(const double *__restrict__ D, long n, int *flag) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long stride = (long)gridDim.x * blockDim.x;
    int found = 0;
    for (; i < n; i += stride) found |= isnan(D[i]);
    if (__any_sync(0xffffffffu, found) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

__global__ void build_mask_kernel(const double *__restrict__ D, unsigned char *__restrict__ mask, int W, int H) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    size_t i = (size_t)y * W + x;
    unsigned m = 0;
    if (x != 0 && !isnan(D[i - 1])) m |= 1;
    if (y != 0 && !isnan(D[i - W])) m |= 2;
    if (x != W - 1 && !isnan(D[i + 1])) m |= 4;
    if (y != H - 1 && !isnan(D[i + W])) m |= 8;
    mask[i] = (unsigned char)m;
}

// ------------------------------------------------------------------------------------------------
// streaming path
// ------------------------------------------------------------------------------------------------
template <bool MASKED>
__global__ void __launch_bounds__(256)
sor_colour_kernel(double *__restrict__ phi, const double *__restrict__ D, const unsigned char *__restrict__ mask,
                  int W, int H, int colour, SorW w, unsigned long long *__restrict__ slot) {
    const int y = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int x = 2 * k + ((y + colour) & 1);
    double a = 0.0;
    if (x < W) {
        const size_t i = (size_t)y * W + x;
        unsigned m;
        if (MASKED) m = mask[i];
        else m = (x != 0 ? 1u : 0u) | (y != 0 ? 2u : 0u) | (x != W - 1 ? 4u : 0u) | (y != H - 1 ? 8u : 0u);
        const double val = phi[i];
        double s = 0.0;
        if (m & 1) s += phi[i - 1];
        if (m & 2) s += phi[i - W];
        if (m & 4) s += phi[i + 1];
        if (m & 8) s += phi[i + W];
        const int cnt = __popc(m);
        const double delta = wsel(w, cnt) * (s - (double)cnt * val - D[i]);
        const double ad = fabs(delta);
        a = ad > 0.0 ? ad : 0.0;  // NaN never enters the max (src/solver.cpp:50-53)
        phi[i] = val + delta;
    }
    a = warp_max(a);
    __shared__ double wm[8];
    if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x < 32) {
        a = threadIdx.x < (blockDim.x >> 5) ? wm[threadIdx.x] : 0.0;
        a = warp_max(a);
        if (threadIdx.x == 0 && a > 0.0) atomicMax(slot, (unsigned long long)__double_as_longlong(a));
    }
}

static int run_streaming(pcd_solver *s, const double *D, double *phi, int max_it, double tol, bool masked,
                         pcd_solve_info *info) {
    const int W = s->W, H = s->H;
    const SorW w = make_w(W);
    const int K = (W + 1) / 2;
    dim3 block(256), grid((K + 255) / 256, H);
    int chunk = s->check_lag > 0 ? s->check_lag : 64;
    if (chunk > 4096) chunk = 4096;  // size of the pinned mirror
    int done = 0, conv = 0;
    double last = 0.0;
    while (done < max_it && !conv) {
        const int k = max_it - done < chunk ? max_it - done : chunk;
        PCD_CUDA(cudaMemsetAsync(s->sweep_max, 0, sizeof(unsigned long long) * k, s->stream));
        PCD_CUDA(cudaEventRecord(s->evk0, s->stream));
        for (int j = 0; j < k; ++j)
            for (int colour = 0; colour < 2; ++colour) {
                if (masked)
                    sor_colour_kernel<true><<<grid, block, 0, s->stream>>>(phi, D, s->mask, W, H, colour, w, s->sweep_max + j);
                else
                    sor_colour_kernel<false><<<grid, block, 0, s->stream>>>(phi, D, nullptr, W, H, colour, w, s->sweep_max + j);
                PCD_LAUNCHED();
                info->launches++;
            }
        PCD_CUDA(cudaEventRecord(s->evk1, s->stream));
        PCD_CUDA(cudaMemcpyAsync(s->h_sweep_max, s->sweep_max, sizeof(unsigned long long) * k, cudaMemcpyDeviceToHost, s->stream));
        PCD_CUDA(cudaStreamSynchronize(s->stream));
        {
            float kms = 0.f;
            PCD_CUDA(cudaEventElapsedTime(&kms, s->evk0, s->evk1));
            info->kernel_ms += kms;
        }
        for (int j = 0; j < k; ++j) {
            double m;
            memcpy(&m, &s->h_sweep_max[j], sizeof(double));
            if (!conv && m < tol) {
                conv = done + j + 1;
                last = m;
            }
            if (!conv && j == k - 1) last = m;
        }
        done += k;
    }
    info->sweeps = done;
    info->converged_at = conv;
    info->last_max_update = last;
    return PCD_OK;
}

// Tiled path (sor_tiled.cu): TS sweeps per pass, ping-pong between phi and a second buffer.  A block of `chunk` sweeps
// between two convergence tests is ONE persistent launch (the CTAs synchronise with their neighbours between passes, no
// kernel boundary); PCD_WAVE_LAUNCH_PER_PASS=1 restores one launch per pass (diagnostics).
// The stopping rule is evaluated ONE BLOCK LATE: block b+1 is queued before the maxima of block b are read (they are
// copied out on a side stream), so the device never drains between blocks.  A solve that meets the rule in block b
// therefore also executes block b+1 -- the same schedule as the multi-GPU slab solvers (slab.py, multi_gpu.cu), which
// is what makes an N-GPU solve bit-identical to this one.
static int run_tiled(pcd_solver *s, const double *D, double *phi, int max_it, double tol, pcd_solve_info *info) {
    const int W = s->W, H = s->H;
    const int TS = tiled_sweeps_per_pass();
    constexpr int HALF = 2048;   // the slot ring / pinned mirror hold two blocks in flight
    if (!s->phi_alt) PCD_CUDA(cudaMalloc(&s->phi_alt, sizeof(double) * (size_t)W * H));
    static const bool per_pass = getenv("PCD_WAVE_LAUNCH_PER_PASS") != nullptr;
    if (!s->wave_ctl) {
        PCD_CUDA(cudaMalloc(&s->wave_ctl, sizeof(unsigned) * (WAVE_MAX_CTAS + 16)));
        PCD_CUDA(cudaMemsetAsync(s->wave_ctl, 0, sizeof(unsigned) * (WAVE_MAX_CTAS + 16), s->stream));
        PCD_CUDA(cudaStreamCreateWithFlags(&s->aux_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            PCD_CUDA(cudaEventCreateWithFlags(&s->ev_blk[i], cudaEventDisableTiming));
            PCD_CUDA(cudaEventCreateWithFlags(&s->ev_copied[i], cudaEventDisableTiming));
        }
    }
    // D staged by TMA from a parity-split copy (one bulk tensor copy per row and CTA instead of two cp.async per thread)
    if (!per_pass && s->dmap_state == 0) {
        s->dmap_state = -1;
        const size_t n = (size_t)tiled_dsplit_pitch(W) * 2 * H;
        if (cudaMalloc(&s->d_split, n * sizeof(double)) == cudaSuccess && tiled_dmap_encode(s->dmap, s->d_split, W, H) == PCD_OK) s->dmap_state = 1;
        else cudaGetLastError();
    }
    const void *dmap = (!per_pass && s->dmap_state == 1) ? s->dmap : nullptr;
    if (dmap) PCD_TRY(tiled_dsplit(D, s->d_split, W, H, s->stream));
    int chunk = s->check_lag > 0 ? s->check_lag : 64;
    chunk = ((chunk + TS - 1) / TS) * TS;  // whole passes
    if (chunk > HALF) chunk = HALF;
    double *buf[2] = {phi, s->phi_alt};
    int cur = 0;
    int launched = 0, done = 0, conv = 0, blk = 0, head = 0, npend = 0;
    int pend_k[2] = {0, 0}, pend_first[2] = {0, 0};
    double last = 0.0;
    PCD_CUDA(cudaEventRecord(s->evk0, s->stream));
    for (;;) {
        if (launched < max_it && !conv && npend < 2) {
            const int k = max_it - launched < chunk ? max_it - launched : chunk;
            const int b = blk & 1, off = b * HALF;
            PCD_CUDA(cudaMemsetAsync(s->sweep_max + off, 0, sizeof(unsigned long long) * k, s->stream));
            for (int j = 0; j < k;) {
                if (per_pass) {
                    const int ns = k - j < TS ? k - j : TS;
                    PCD_TRY(tiled_pass(buf[cur], buf[cur ^ 1], D, W, H, 0, H, 0, ns, s->sweep_max + off + j, s->sm_count, 0, s->stream));
                    cur ^= 1;
                    j += ns;
                } else {
                    const int spp = k - j >= TS ? TS : 1;
                    const int npass = k - j >= TS ? (k - j) / TS : k - j;
                    WavePeer pr;
                    pr.buf[0] = buf[0]; pr.buf[1] = buf[1];
                    pr.cur = cur;
                    pr.npass = npass;
                    pr.done = s->wave_ctl;
                    pr.err = reinterpret_cast<int *>(s->wave_ctl + WAVE_MAX_CTAS);
                    pr.seq0 = s->wave_seq;
                    static const char *trace_prefix = getenv("PCD_WAVE_TRACE");   // diagnostics: per-CTA, per-pass timestamps
                    if (trace_prefix && npass <= 64) {
                        if (!s->wave_trace) PCD_CUDA(cudaMalloc(&s->wave_trace, sizeof(unsigned long long) * 4 * WAVE_MAX_CTAS * 64));
                        PCD_CUDA(cudaMemsetAsync(s->wave_trace, 0, sizeof(unsigned long long) * 4 * WAVE_MAX_CTAS * 64, s->stream));
                        pr.trace = s->wave_trace;
                        s->wave_trace_npass = npass;
                    }
                    PCD_TRY(tiled_run_peer(D, W, H, 0, H, 0, spp, s->sweep_max + off + j, pr, dmap, s->sm_count, 0, s->stream));
                    s->wave_seq += (unsigned)npass;
                    cur ^= (npass & 1);
                    j += npass * spp;
                }
                info->launches++;
            }
            PCD_CUDA(cudaEventRecord(s->ev_blk[b], s->stream));
            PCD_CUDA(cudaStreamWaitEvent(s->aux_stream, s->ev_blk[b], 0));
            PCD_CUDA(cudaMemcpyAsync(s->h_sweep_max + off, s->sweep_max + off, sizeof(unsigned long long) * k, cudaMemcpyDeviceToHost, s->aux_stream));
            PCD_CUDA(cudaMemcpyAsync(s->h_flags + 2 + b, s->wave_ctl + WAVE_MAX_CTAS, sizeof(int), cudaMemcpyDeviceToHost, s->aux_stream));
            PCD_CUDA(cudaEventRecord(s->ev_copied[b], s->aux_stream));
            pend_k[b] = k; pend_first[b] = launched;
            launched += k; ++blk; ++npend;
            continue;
        }
        if (!npend) break;
        const int b = head & 1, off = b * HALF, k = pend_k[b];
        PCD_CUDA(cudaEventSynchronize(s->ev_copied[b]));
        if (s->h_flags[2 + b]) {
            set_error("wavefront K-SOR kernel: a CTA gave up waiting for a neighbouring CTA");
            return PCD_ERR_CUDA;
        }
        for (int j = 0; j < k && !conv; ++j) {
            double m;
            memcpy(&m, &s->h_sweep_max[off + j], sizeof(double));
            if (m < tol) conv = pend_first[b] + j + 1;
            if (conv || j == k - 1) last = m;
        }
        done = pend_first[b] + k;
        ++head; --npend;
    }
    PCD_CUDA(cudaEventRecord(s->evk1, s->stream));
    if (cur != 0) PCD_CUDA(cudaMemcpyAsync(phi, buf[cur], sizeof(double) * (size_t)W * H, cudaMemcpyDeviceToDevice, s->stream));
    PCD_CUDA(cudaStreamSynchronize(s->stream));
    {
        float kms = 0.f;
        PCD_CUDA(cudaEventElapsedTime(&kms, s->evk0, s->evk1));
        info->kernel_ms += kms;
    }
    info->sweeps = done;
    info->converged_at = conv;
    info->last_max_update = last;
    return PCD_OK;
}

// ------------------------------------------------------------------------------------------------
// solver object
// ------------------------------------------------------------------------------------------------
int solver_init(pcd_solver *s, int W, int H, int device, int path, cudaStream_t stream) {
    if (W < 1 || H < 1 || H > 65535) {
        set_error("poisson solver: unsupported grid %dx%d", W, H);
        return PCD_ERR_INVALID;
    }
    s->W = W; s->H = H; s->device = device; s->path_req = path;
    PCD_TRY(select_device(device));
    cudaDeviceProp prop;
    PCD_CUDA(cudaGetDeviceProperties(&prop, device));
    s->sm_count = prop.multiProcessorCount;
    if (stream) {
        s->stream = stream;
        s->own_stream = false;
    } else {
        PCD_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        s->own_stream = true;
    }
    s->ring = RES_MAX_SWEEPS_PER_LAUNCH;
    PCD_CUDA(cudaMalloc(&s->sweep_max, 2 * sizeof(unsigned long long) * (size_t)s->ring));  // per-sweep max + arrival slot
    PCD_CUDA(cudaMallocHost(&s->h_sweep_max, sizeof(unsigned long long) * 4096));
    PCD_CUDA(cudaMalloc(&s->d_flags, sizeof(int) * 4));
    PCD_CUDA(cudaMallocHost(&s->h_flags, sizeof(int) * 8));
    memset(s->h_flags, 0, sizeof(int) * 8);
    PCD_CUDA(cudaMalloc(&s->res_state, sizeof(ResState)));
    PCD_CUDA(cudaMallocHost(&s->h_res_state, sizeof(ResState)));
    PCD_CUDA(cudaEventCreate(&s->ev0));
    PCD_CUDA(cudaEventCreate(&s->ev1));
    PCD_CUDA(cudaEventCreate(&s->evk0));
    PCD_CUDA(cudaEventCreate(&s->evk1));
    // AUTO: resident when the grid fits on chip, else tiled (temporal blocking); STREAMING = plain colour launches
    s->path_used = (path == PCD_SOLVER_STREAMING) ? PCD_SOLVER_STREAMING : (path == PCD_SOLVER_DCT ? PCD_SOLVER_DCT : PCD_SOLVER_TILED);
    if (path == PCD_SOLVER_AUTO || path == PCD_SOLVER_RESIDENT) {
        int coop = 0;
        PCD_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
        if (coop && resident_plan(s)) {
            s->path_used = PCD_SOLVER_RESIDENT;
            PCD_CUDA(cudaMalloc(&s->halo, (size_t)s->res_ctas * 2 * resident_slots() * sizeof(uint4)));
        } else if (path == PCD_SOLVER_RESIDENT) {
            set_error("poisson solver: grid %dx%d does not fit the resident path on this device", W, H);
            return PCD_ERR_UNSUPPORTED;
        }
    }
    return PCD_OK;
}

void solver_free(pcd_solver *s) {
    if (!s) return;
    if (s->wave_trace) {   // PCD_WAVE_TRACE=<prefix>: <prefix>_solver<W>x<H>.bin, same format as the slabs' dumps (tools/wave_trace.py)
        const char *prefix = getenv("PCD_WAVE_TRACE");
        const size_t n = (size_t)4 * WAVE_MAX_CTAS * (s->wave_trace_npass > 0 ? s->wave_trace_npass : 1);
        unsigned long long *h = (unsigned long long *)malloc(n * sizeof(unsigned long long));
        if (prefix && h && cudaMemcpy(h, s->wave_trace, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
            char path[1200];
            snprintf(path, sizeof(path), "%s_solver%dx%d.bin", prefix, s->W, s->H);
            if (FILE *f = fopen(path, "wb")) {
                const int hdr[2] = {s->wave_trace_npass, WAVE_MAX_CTAS};
                fwrite(hdr, sizeof(int), 2, f);
                fwrite(h, sizeof(unsigned long long), n, f);
                fclose(f);
            }
        }
        free(h);
        cudaFree(s->wave_trace);
    }
    dct_free(s);
    cudaFree(s->sweep_max); cudaFreeHost(s->h_sweep_max);
    cudaFree(s->mask); cudaFree(s->d_flags); cudaFreeHost(s->h_flags);
    cudaFree(s->res_state); cudaFreeHost(s->h_res_state); cudaFree(s->halo); cudaFree(s->phi_alt); cudaFree(s->wave_ctl);
    cudaFree(s->d_split);
    cudaFree(s->tr_D); cudaFree(s->tr_phi);
    if (s->aux_stream) cudaStreamDestroy(s->aux_stream);
    for (int i = 0; i < 2; ++i) {
        if (s->ev_blk[i]) cudaEventDestroy(s->ev_blk[i]);
        if (s->ev_copied[i]) cudaEventDestroy(s->ev_copied[i]);
    }
    if (s->own_fields) { cudaFree(s->D); cudaFree(s->phi); }
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->evk0) cudaEventDestroy(s->evk0);
    if (s->evk1) cudaEventDestroy(s->evk1);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    *s = pcd_solver();
}

int solver_run(pcd_solver *s, const double *D, double *phi, int max_iterations, double tol, pcd_solve_info *info) {
    pcd_solve_info local{};
    if (!info) info = &local;
    *info = pcd_solve_info{};
    info->path = s->path_used;
    s->res_exchange = 0;
    PCD_TRY(select_device(s->device));
    if (max_iterations <= 0) return PCD_OK;  // the reference's loop body never runs (src/solver.cpp:92)
    PCD_CUDA(cudaEventRecord(s->ev0, s->stream));
    int rc = PCD_OK;
    int path = s->path_used;
    if (path == PCD_SOLVER_RESIDENT) {
        rc = run_resident(s, D, phi, max_iterations, tol, info);
        if (rc == PCD_RES_FALLBACK) {   // strips of 8-9 rows only exist in the deep-halo kernel, and D has NaN holes: large-grid paths
            *info = pcd_solve_info{};
            path = PCD_SOLVER_TILED;
            info->path = path;
        }
    }
    if (path != PCD_SOLVER_RESIDENT) {
        // NaN holes change the neighbour rule (src/solver.cpp:29-44); only then is a mask array built
        const long n = (long)s->W * s->H;
        PCD_CUDA(cudaMemsetAsync(s->d_flags, 0, sizeof(int), s->stream));
        detect_nan_kernel<<<296, 256, 0, s->stream>>>(D, n, s->d_flags);
        PCD_LAUNCHED();
        info->launches++;
        PCD_CUDA(cudaMemcpyAsync(s->h_flags, s->d_flags, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        PCD_CUDA(cudaStreamSynchronize(s->stream));
        const bool masked = s->h_flags[0] != 0;
        if (masked) {
            if (!s->mask) PCD_CUDA(cudaMalloc(&s->mask, (size_t)n));
            build_mask_kernel<<<dim3((s->W + 255) / 256, s->H), 256, 0, s->stream>>>(D, s->mask, s->W, s->H);
            PCD_LAUNCHED();
            info->launches++;
        }
        // the tiled path derives neighbour counts from coordinates, the direct backend needs the plain operator: NaN
        // holes go through the masked colour kernels
        if (path == PCD_SOLVER_TILED && !masked) {
            rc = run_tiled(s, D, phi, max_iterations, tol, info);
        } else if (path == PCD_SOLVER_DCT && !masked) {
            rc = run_dct(s, D, phi, info);
        } else {
            info->path = PCD_SOLVER_STREAMING;
            rc = run_streaming(s, D, phi, max_iterations, tol, masked, info);
        }
    }
    PCD_TRY(rc);
    PCD_CUDA(cudaEventRecord(s->ev1, s->stream));
    PCD_CUDA(cudaEventSynchronize(s->ev1));
    float ms = 0.f;
    PCD_CUDA(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    info->device_ms = ms;
    return PCD_OK;
}

}  // namespace pcd
