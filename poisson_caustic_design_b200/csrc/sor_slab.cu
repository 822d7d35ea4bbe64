// K-SOR on a ROW SLAB of a larger grid (multi-GPU): the process owns global rows
// [row0, row0+rows) of a W x H grid plus GH ghost rows above and below (local row r <-> global row
// row0 - GH + r).  Two ways to advance:
//   * pcd_slab_pass        : one wavefront pass (sor_tiled.cu) = up to TS full sweeps, field ping-pongs between
//                            two buffers; afterwards the host layer refreshes the GH ghost rows from the
//                            neighbouring ranks (one exchange per TS sweeps, GH = 2*TS+1 rows each way);
//   * pcd_slab_sweep_colour: one colour phase in place (masked kernel; the NaN-hole path), one ghost row
//                            exchanged per phase.
// The host layer (slab.py, torch.distributed: NCCL over NVLink, gloo in CPU tests) does the exchanges and
// all-reduces the per-sweep max.  Per-cell arithmetic is the single-GPU kernels', so a G-slab solve is
// bit-identical to the single-GPU solve (red-black updates of one colour are order-independent, max is exact).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "sor_common.cuh"

// Control words behind phi[0] (one CUDA-IPC handle exports the field and them): unsigned ctl[]
//   [0, 512)     per-strip flags raised by the UPPER neighbour (sequence number of its last pass whose edge rows are here)
//   [512, 1024)  the same from the LOWER neighbour
//   [1024]       error word: a pass ran into its time limit waiting for a neighbour
constexpr int CTL_FROM_UP = 0, CTL_FROM_DN = pcd::WAVE_MAX_STRIPS, CTL_ERR = 2 * pcd::WAVE_MAX_STRIPS;
constexpr size_t PCD_SLAB_CTL_BYTES = (2 * pcd::WAVE_MAX_STRIPS + 64) * sizeof(unsigned);

struct pcd_slab {
    int W = 0, H = 0, row0 = 0, rows = 0, device = 0, GH = 0;
    int sm_count = 0, sm_reserve = 0;   // SMs of this slab's device / SMs its passes leave free for an overlapped exchange
    cudaStream_t stream = nullptr;   // caller's stream (e.g. torch's current stream), never owned
    double *phi[2] = {nullptr, nullptr};  // (rows + 2*GH) x W each; phi[cur] holds the field
    double *D = nullptr;
    int cur = 0;
    unsigned char *mask = nullptr;                  // neighbour masks (colour path)
    unsigned long long *sweep_max = nullptr;        // [ring]
    int *d_flag = nullptr;
    int has_nan = 0;
    int ring = 0;
    long long launches = 0;
    // fused peer exchange (pcd_slab_peer_*): control words live behind phi[0] so that one IPC handle covers them
    unsigned *ctl = nullptr;              // see CTL_* above
    unsigned *done = nullptr;             // per-CTA sequence words of the persistent pass kernel [WAVE_MAX_CTAS]
    int passes_per_launch = 0;            // 0 = a whole block of sweeps per launch; 1 when a neighbour shares this device
    double *d_split = nullptr;            // parity-split copy of D (TMA staging of the pass kernel), refreshed when D changes
    alignas(64) unsigned char dmap[128] = {};
    bool dmap_ok = false;
    unsigned long long *trace = nullptr;  // PCD_WAVE_TRACE=<prefix>: timestamps of the last launch, dumped at destroy
    int trace_npass = 0;
    double *peer_phi[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [side][buffer], side 0 = up, 1 = down
    unsigned *peer_ctl[2] = {nullptr, nullptr};
    int peer_row0[2] = {0, 0};
    bool peer_ipc[2] = {false, false};
    unsigned seq = 0;
};

namespace pcd {

__global__ void slab_mask_kernel(const double *__restrict__ D, unsigned char *__restrict__ mask, int W, int H, int row0,
                                 int rows, int GH, int *__restrict__ nan_flag) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y + GH;  // owned local rows GH..GH+rows-1
    if (x >= W) return;
    const int gy = row0 + r - GH;
    const size_t i = (size_t)r * W + x;
    unsigned m = 0;
    bool any_nan = isnan(D[i]);
    if (x != 0) { if (!isnan(D[i - 1])) m |= 1; else any_nan = true; }
    if (gy != 0) { if (!isnan(D[i - W])) m |= 2; else any_nan = true; }
    if (x != W - 1) { if (!isnan(D[i + 1])) m |= 4; else any_nan = true; }
    if (gy != H - 1) { if (!isnan(D[i + W])) m |= 8; else any_nan = true; }
    mask[i] = (unsigned char)m;
    if (any_nan) atomicOr(nan_flag, 1);
}

__global__ void __launch_bounds__(256)
sor_slab_colour_kernel(double *__restrict__ phi, const double *__restrict__ D, const unsigned char *__restrict__ mask,
                       int W, int row0, int GH, int colour, SorW w, unsigned long long *__restrict__ slot) {
    const int r = blockIdx.y + GH;
    const int gy = row0 + r - GH;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int x = 2 * k + ((gy + colour) & 1);
    double a = 0.0;
    if (x < W) {
        const size_t i = (size_t)r * W + x;
        const unsigned m = mask[i];
        const double val = phi[i];
        double s = 0.0;
        if (m & 1) s += phi[i - 1];
        if (m & 2) s += phi[i - W];
        if (m & 4) s += phi[i + 1];
        if (m & 8) s += phi[i + W];
        const int cnt = __popc(m);
        const double delta = wsel(w, cnt) * (s - (double)cnt * val - D[i]);
        const double ad = fabs(delta);
        a = ad > 0.0 ? ad : 0.0;
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

// the slab's error word as a double next to the per-sweep maxima, so it rides in the same all-reduce (MAX)
__global__ void slab_error_to_kernel(const unsigned *err, double *dst) { *dst = *err ? 1.0 : 0.0; }

}  // namespace pcd

using namespace pcd;

extern "C" {

int pcd_slab_ghost_rows(void) { return 2 * tiled_sweeps_per_pass() + 1; }
int pcd_slab_sweeps_per_pass(void) { return tiled_sweeps_per_pass(); }
int pcd_slab_set_sm_reserve(pcd_slab *s, int n) {
    if (!s) { set_error("null slab"); return PCD_ERR_INVALID; }
    s->sm_reserve = n < 0 ? 0 : n;
    return PCD_OK;
}

int pcd_slab_create(int width, int height, int row0, int rows, int device, void *cuda_stream, pcd_slab **out) {
    if (!out) { set_error("null argument"); return PCD_ERR_INVALID; }
    *out = nullptr;
    if (width < 1 || height < 1 || rows < 1 || row0 < 0 || row0 + rows > height || rows > 65535) {
        set_error("pcd_slab_create: rows [%d,%d) of a %dx%d grid is not a valid slab", row0, row0 + rows, width, height);
        return PCD_ERR_INVALID;
    }
    PCD_TRY(select_device(device));
    pcd_slab *s = new pcd_slab();
    s->W = width; s->H = height; s->row0 = row0; s->rows = rows; s->device = device;
    s->GH = pcd_slab_ghost_rows();
    cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, device);
    s->stream = (cudaStream_t)cuda_stream;
    s->ring = 4096;
    const size_t n = (size_t)(rows + 2 * s->GH) * width;
    cudaError_t e = cudaMalloc(&s->phi[0], n * sizeof(double) + PCD_SLAB_CTL_BYTES);
    if (e == cudaSuccess) e = cudaMalloc(&s->phi[1], n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&s->D, n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&s->mask, n);
    if (e == cudaSuccess) e = cudaMalloc(&s->sweep_max, sizeof(unsigned long long) * s->ring);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_flag, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&s->done, sizeof(unsigned) * WAVE_MAX_CTAS);
    if (e == cudaSuccess) e = cudaMemset(s->done, 0, sizeof(unsigned) * WAVE_MAX_CTAS);
    if (e == cudaSuccess) e = cudaMemset(s->phi[0], 0, n * sizeof(double) + PCD_SLAB_CTL_BYTES);
    if (e == cudaSuccess) e = cudaMemset(s->phi[1], 0, n * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(s->D, 0, n * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(s->mask, 0, n);
    if (e == cudaSuccess) e = cudaMemset(s->sweep_max, 0, sizeof(unsigned long long) * s->ring);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        set_error("slab allocation failed: %s", cudaGetErrorString(e));
        cudaFree(s->phi[0]); cudaFree(s->phi[1]); cudaFree(s->D); cudaFree(s->mask); cudaFree(s->sweep_max); cudaFree(s->d_flag);
        cudaFree(s->done);
        delete s;
        return PCD_ERR_CUDA;
    }
    s->ctl = reinterpret_cast<unsigned *>(s->phi[0] + n);
    {   // optional: D staged by TMA from a parity-split copy
        const size_t ns = (size_t)tiled_dsplit_pitch(width) * 2 * (rows + 2 * s->GH);
        if (cudaMalloc(&s->d_split, ns * sizeof(double)) == cudaSuccess && cudaMemset(s->d_split, 0, ns * sizeof(double)) == cudaSuccess &&
            tiled_dmap_encode(s->dmap, s->d_split, width, rows + 2 * s->GH) == PCD_OK)
            s->dmap_ok = true;
        else
            cudaGetLastError();
    }
    *out = s;
    return PCD_OK;
}

void pcd_slab_destroy(pcd_slab *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->trace) {   // PCD_WAVE_TRACE=<prefix>: <prefix>_row<row0>.bin = int32 header [npass, max_ctas] + u64 [npass][max_ctas][4]
        const char *prefix = getenv("PCD_WAVE_TRACE");
        const size_t n = (size_t)4 * WAVE_MAX_CTAS * (s->trace_npass > 0 ? s->trace_npass : 1);
        unsigned long long *h = (unsigned long long *)malloc(n * sizeof(unsigned long long));
        if (prefix && h && cudaMemcpy(h, s->trace, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
            char path[1200];
            snprintf(path, sizeof(path), "%s_row%d.bin", prefix, s->row0);
            if (FILE *f = fopen(path, "wb")) {
                const int hdr[2] = {s->trace_npass, WAVE_MAX_CTAS};
                fwrite(hdr, sizeof(int), 2, f);
                fwrite(h, sizeof(unsigned long long), n, f);
                fclose(f);
            }
        }
        free(h);
        cudaFree(s->trace);
    }
    for (int side = 0; side < 2; ++side)
        if (s->peer_ipc[side]) {
            cudaIpcCloseMemHandle(s->peer_phi[side][0]);
            cudaIpcCloseMemHandle(s->peer_phi[side][1]);
        }
    cudaFree(s->phi[0]); cudaFree(s->phi[1]); cudaFree(s->D); cudaFree(s->mask); cudaFree(s->sweep_max); cudaFree(s->d_flag);
    cudaFree(s->done);
    cudaFree(s->d_split);
    delete s;
}

int pcd_slab_device_ptrs(pcd_slab *s, void **phi0_dev, void **phi1_dev, void **sweep_max_dev) {
    if (!s) { set_error("null slab"); return PCD_ERR_INVALID; }
    if (phi0_dev) *phi0_dev = s->phi[0];
    if (phi1_dev) *phi1_dev = s->phi[1];
    if (sweep_max_dev) *sweep_max_dev = s->sweep_max;
    return PCD_OK;
}

// device address of the slab's error word (int; 1 = a pass ran into its time limit waiting for a neighbour)
int pcd_slab_error_word(pcd_slab *s, void **err_dev) {
    if (!s || !err_dev) { set_error("null argument"); return PCD_ERR_INVALID; }
    *err_dev = s->ctl + CTL_ERR;
    return PCD_OK;
}

int pcd_slab_current(const pcd_slab *s) { return s ? s->cur : -1; }
int pcd_slab_has_nan(const pcd_slab *s) { return s ? s->has_nan : -1; }

// D_rows / phi_rows: host arrays of (rows + 2*GH) x W doubles INCLUDING the ghost rows (ghost rows outside the
// grid are ignored); either may be NULL.
int pcd_slab_upload(pcd_slab *s, const double *D_rows, const double *phi_rows) {
    if (!s) { set_error("null slab"); return PCD_ERR_INVALID; }
    PCD_TRY(select_device(s->device));
    const size_t bytes = (size_t)(s->rows + 2 * s->GH) * s->W * sizeof(double);
    if (D_rows) {
        PCD_CUDA(cudaMemcpyAsync(s->D, D_rows, bytes, cudaMemcpyHostToDevice, s->stream));
        PCD_CUDA(cudaMemsetAsync(s->d_flag, 0, sizeof(int), s->stream));
        slab_mask_kernel<<<dim3((s->W + 255) / 256, s->rows), 256, 0, s->stream>>>(s->D, s->mask, s->W, s->H, s->row0, s->rows,
                                                                                 s->GH, s->d_flag);
        PCD_LAUNCHED();
        s->launches++;
        PCD_CUDA(cudaMemcpyAsync(&s->has_nan, s->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        if (s->dmap_ok) PCD_TRY(tiled_dsplit(s->D, s->d_split, s->W, s->rows + 2 * s->GH, s->stream));
    }
    if (phi_rows) {
        PCD_CUDA(cudaMemcpyAsync(s->phi[0], phi_rows, bytes, cudaMemcpyHostToDevice, s->stream));
        s->cur = 0;
    }
    PCD_CUDA(cudaMemsetAsync(s->ctl + CTL_ERR, 0, sizeof(unsigned), s->stream));   // a new solve starts with a clean error word
    PCD_CUDA(cudaStreamSynchronize(s->stream));
    return PCD_OK;
}

// the same from full W x H device arrays (multi-GPU solve hook: every rank holds the whole right-hand side)
int pcd_slab_load_device(pcd_slab *s, const double *D_full, const double *phi_full) {
    if (!s || !D_full || !phi_full) { set_error("null argument"); return PCD_ERR_INVALID; }
    PCD_TRY(select_device(s->device));
    const size_t W = (size_t)s->W;
    const size_t all = (size_t)(s->rows + 2 * s->GH) * W * sizeof(double);
    const int lo = s->row0 - s->GH > 0 ? s->row0 - s->GH : 0;
    const int hi = s->row0 + s->rows + s->GH < s->H ? s->row0 + s->rows + s->GH : s->H;
    const size_t off = (size_t)(lo - (s->row0 - s->GH)) * W, cnt = (size_t)(hi - lo) * W * sizeof(double);
    PCD_CUDA(cudaMemsetAsync(s->D, 0, all, s->stream));       // ghost rows outside the grid
    PCD_CUDA(cudaMemsetAsync(s->phi[0], 0, all, s->stream));
    PCD_CUDA(cudaMemcpyAsync(s->D + off, D_full + (size_t)lo * W, cnt, cudaMemcpyDeviceToDevice, s->stream));
    PCD_CUDA(cudaMemcpyAsync(s->phi[0] + off, phi_full + (size_t)lo * W, cnt, cudaMemcpyDeviceToDevice, s->stream));
    PCD_CUDA(cudaMemsetAsync(s->ctl + CTL_ERR, 0, sizeof(unsigned), s->stream));   // a new solve starts with a clean error word
    s->cur = 0;
    PCD_CUDA(cudaMemsetAsync(s->d_flag, 0, sizeof(int), s->stream));
    slab_mask_kernel<<<dim3((s->W + 255) / 256, s->rows), 256, 0, s->stream>>>(s->D, s->mask, s->W, s->H, s->row0, s->rows, s->GH,
                                                                             s->d_flag);
    PCD_LAUNCHED();
    s->launches++;
    PCD_CUDA(cudaMemcpyAsync(&s->has_nan, s->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    if (s->dmap_ok) PCD_TRY(tiled_dsplit(s->D, s->d_split, s->W, s->rows + 2 * s->GH, s->stream));
    PCD_CUDA(cudaStreamSynchronize(s->stream));
    return PCD_OK;
}

int pcd_slab_store_device(pcd_slab *s, double *phi_full) {
    if (!s || !phi_full) { set_error("null argument"); return PCD_ERR_INVALID; }
    PCD_TRY(select_device(s->device));
    PCD_CUDA(cudaMemcpyAsync(phi_full + (size_t)s->row0 * s->W, s->phi[s->cur] + (size_t)s->GH * s->W,
                             (size_t)s->rows * s->W * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    return PCD_OK;
}

int pcd_slab_download(pcd_slab *s, double *phi_owned_rows) {
    if (!s || !phi_owned_rows) { set_error("null argument"); return PCD_ERR_INVALID; }
    PCD_TRY(select_device(s->device));
    PCD_CUDA(cudaMemcpyAsync(phi_owned_rows, s->phi[s->cur] + (size_t)s->GH * s->W, (size_t)s->rows * s->W * sizeof(double),
                             cudaMemcpyDeviceToHost, s->stream));
    PCD_CUDA(cudaStreamSynchronize(s->stream));
    return PCD_OK;
}

// one colour phase of sweep slot `slot` (0 <= slot < 4096) in place on the current buffer; asynchronous
int pcd_slab_sweep_colour(pcd_slab *s, int colour, int slot) {
    if (!s || slot < 0 || slot >= s->ring) { set_error("bad slab / slot"); return PCD_ERR_INVALID; }
    const int K = (s->W + 1) / 2;
    sor_slab_colour_kernel<<<dim3((K + 255) / 256, s->rows), 256, 0, s->stream>>>(s->phi[s->cur], s->D, s->mask, s->W, s->row0,
                                                                                   s->GH, colour, make_w(s->W), s->sweep_max + slot);
    PCD_LAUNCHED();
    s->launches++;
    return PCD_OK;
}

// one wavefront pass: nsweeps (<= pcd_slab_sweeps_per_pass()) full sweeps, maxima into slots [slot, slot+nsweeps);
// the field moves to the other buffer.  Requires valid ghost rows and no NaN in D.
int pcd_slab_pass(pcd_slab *s, int nsweeps, int slot) {
    if (!s || slot < 0 || slot + nsweeps > s->ring || nsweeps < 1 || nsweeps > tiled_sweeps_per_pass()) {
        set_error("bad slab / slot / sweep count");
        return PCD_ERR_INVALID;
    }
    if (s->has_nan) { set_error("pcd_slab_pass: D has NaN holes, use the colour path"); return PCD_ERR_UNSUPPORTED; }
    PCD_TRY(tiled_pass(s->phi[s->cur], s->phi[s->cur ^ 1], s->D, s->W, s->H, s->row0, s->rows, s->row0 - s->GH, nsweeps,
                       s->sweep_max + slot, s->sm_count, s->sm_reserve, s->stream));
    s->cur ^= 1;
    s->launches++;
    return PCD_OK;
}

// The same pass restricted to owned rows [row_begin, row_begin+row_count) (global indices), on `cuda_stream` (NULL =
// the slab's stream), WITHOUT switching buffers: the host layer runs the bands next to the slab edges first,
// starts the ghost-row exchange on a side stream, runs the interior, then calls pcd_slab_flip.
int pcd_slab_pass_part(pcd_slab *s, int nsweeps, int slot, int row_begin, int row_count, void *cuda_stream) {
    if (!s || slot < 0 || slot + nsweeps > s->ring || nsweeps < 1 || nsweeps > tiled_sweeps_per_pass() ||
        row_begin < s->row0 || row_count < 1 || row_begin + row_count > s->row0 + s->rows) {
        set_error("bad slab / slot / sweep count / row range");
        return PCD_ERR_INVALID;
    }
    if (s->has_nan) { set_error("pcd_slab_pass_part: D has NaN holes, use the colour path"); return PCD_ERR_UNSUPPORTED; }
    PCD_TRY(tiled_pass(s->phi[s->cur], s->phi[s->cur ^ 1], s->D, s->W, s->H, row_begin, row_count, s->row0 - s->GH, nsweeps,
                       s->sweep_max + slot, s->sm_count, s->sm_reserve, cuda_stream ? (cudaStream_t)cuda_stream : s->stream));
    s->launches++;
    return PCD_OK;
}

int pcd_slab_flip(pcd_slab *s) {
    if (!s) { set_error("null slab"); return PCD_ERR_INVALID; }
    s->cur ^= 1;
    return PCD_OK;
}

// ---- ghost-row exchange fused into the pass (peer memory over NVLink; see WavePeer in sor_common.cuh) ----

int pcd_slab_peer_handle_bytes(void) { return 2 * (int)sizeof(cudaIpcMemHandle_t); }

// IPC handles of the two field buffers (the control words sit behind buffer 0) for the neighbouring processes
int pcd_slab_peer_export(pcd_slab *s, unsigned char *handles) {
    if (!s || !handles) { set_error("null argument"); return PCD_ERR_INVALID; }
    PCD_TRY(select_device(s->device));
    cudaIpcMemHandle_t h[2];
    PCD_CUDA(cudaIpcGetMemHandle(&h[0], s->phi[0]));
    PCD_CUDA(cudaIpcGetMemHandle(&h[1], s->phi[1]));
    memcpy(handles, h, sizeof(h));
    return PCD_OK;
}

static int peer_attach(pcd_slab *s, int side, double *phi0, double *phi1, int peer_row0, int peer_rows, bool ipc) {
    const int expect = side == 0 ? s->row0 - peer_rows : s->row0 + s->rows;
    if (peer_row0 != expect || peer_rows < 2 * s->GH || s->rows < 2 * s->GH) {
        set_error("pcd_slab_peer_connect: rows [%d,%d) are not the %s neighbour of [%d,%d), or a slab is thinner than %d rows",
                  peer_row0, peer_row0 + peer_rows, side == 0 ? "upper" : "lower", s->row0, s->row0 + s->rows, 2 * s->GH);
        return PCD_ERR_INVALID;
    }
    if (tiled_strips(s->W) > WAVE_MAX_STRIPS) {
        set_error("pcd_slab_peer_connect: a %d-wide grid has more than %d strips", s->W, WAVE_MAX_STRIPS);
        return PCD_ERR_UNSUPPORTED;
    }
    s->peer_phi[side][0] = phi0;
    s->peer_phi[side][1] = phi1;
    s->peer_ctl[side] = reinterpret_cast<unsigned *>(phi0 + (size_t)(peer_rows + 2 * s->GH) * s->W);
    s->peer_row0[side] = peer_row0;
    s->peer_ipc[side] = ipc;
    return PCD_OK;
}

// side 0: the slab that owns the rows above, side 1: the rows below.  `handles` from pcd_slab_peer_export in the
// neighbour's process.
int pcd_slab_peer_connect_ipc(pcd_slab *s, int side, const unsigned char *handles, int peer_row0, int peer_rows) {
    if (!s || !handles || side < 0 || side > 1 || s->peer_phi[side][0]) { set_error("bad slab / side / handles"); return PCD_ERR_INVALID; }
    PCD_TRY(select_device(s->device));
    cudaIpcMemHandle_t h[2];
    memcpy(h, handles, sizeof(h));
    void *p0 = nullptr, *p1 = nullptr;
    PCD_CUDA(cudaIpcOpenMemHandle(&p0, h[0], cudaIpcMemLazyEnablePeerAccess));
    cudaError_t e = cudaIpcOpenMemHandle(&p1, h[1], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaIpcCloseMemHandle(p0);
        set_error("cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
        return PCD_ERR_CUDA;
    }
    const int rc = peer_attach(s, side, (double *)p0, (double *)p1, peer_row0, peer_rows, true);
    if (rc != PCD_OK) { cudaIpcCloseMemHandle(p0); cudaIpcCloseMemHandle(p1); }
    return rc;
}

// The neighbour lives in this process: on the same device (several slabs per GPU: single-GPU tests of the fused path;
// the slabs' kernels then run one after the other, so every launch is limited to ONE pass) or on another device of
// this process (the torch-free multi-GPU host, host/multi_gpu.cpp: peer access is enabled here).
int pcd_slab_peer_connect_local(pcd_slab *s, int side, pcd_slab *peer) {
    if (!s || !peer || side < 0 || side > 1 || peer->W != s->W) { set_error("bad slab / side / peer"); return PCD_ERR_INVALID; }
    if (peer->device != s->device) {
        PCD_TRY(select_device(s->device));
        int can = 0;
        PCD_CUDA(cudaDeviceCanAccessPeer(&can, s->device, peer->device));
        if (!can) { set_error("device %d cannot access the memory of device %d", s->device, peer->device); return PCD_ERR_UNSUPPORTED; }
        cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
        if (e != cudaSuccess) { set_error("cudaDeviceEnablePeerAccess(%d): %s", peer->device, cudaGetErrorString(e)); return PCD_ERR_CUDA; }
    } else {
        s->passes_per_launch = 1;
    }
    return peer_attach(s, side, peer->phi[0], peer->phi[1], peer->row0, peer->rows, false);
}

// nsweeps sweeps as ceil(nsweeps / TS) fused passes, maxima into slots [slot, slot+nsweeps).  Asynchronous: ONE persistent
// launch on the slab's stream for all full passes (plus one for a trailing single sweep).  Every rank must issue the
// same sequence of peer runs.
int pcd_slab_peer_run(pcd_slab *s, int nsweeps, int slot) {
    if (!s || slot < 0 || nsweeps < 1 || slot + nsweeps > s->ring) { set_error("bad slab / slot / sweep count"); return PCD_ERR_INVALID; }
    if (s->has_nan) { set_error("pcd_slab_peer_run: D has NaN holes, use the colour path"); return PCD_ERR_UNSUPPORTED; }
    PCD_TRY(select_device(s->device));
    const int TS = tiled_sweeps_per_pass();
    int j = 0;
    while (j < nsweeps) {
        const int left = nsweeps - j;
        const int spp = left >= TS ? TS : 1;                       // sweeps per pass of this launch
        int npass = left >= TS ? left / TS : left;
        if (s->passes_per_launch > 0 && npass > s->passes_per_launch) npass = s->passes_per_launch;
        WavePeer pr;
        pr.gh = s->GH;
        pr.buf[0] = s->phi[0]; pr.buf[1] = s->phi[1];
        pr.cur = s->cur;
        pr.npass = npass;
        pr.done = s->done;
        pr.err = reinterpret_cast<int *>(s->ctl + CTL_ERR);
        pr.seq0 = s->seq;
        if (s->peer_phi[0][0]) {
            pr.up_buf[0] = s->peer_phi[0][0]; pr.up_buf[1] = s->peer_phi[0][1]; pr.up_grow0 = s->peer_row0[0] - s->GH;
            pr.wait_up = s->ctl + CTL_FROM_UP; pr.sig_up = s->peer_ctl[0] + CTL_FROM_DN;   // I am its lower neighbour
        }
        if (s->peer_phi[1][0]) {
            pr.dn_buf[0] = s->peer_phi[1][0]; pr.dn_buf[1] = s->peer_phi[1][1]; pr.dn_grow0 = s->peer_row0[1] - s->GH;
            pr.wait_dn = s->ctl + CTL_FROM_DN; pr.sig_dn = s->peer_ctl[1] + CTL_FROM_UP;   // I am its upper neighbour
        }
        static const char *trace_prefix = getenv("PCD_WAVE_TRACE");   // diagnostics: per-CTA, per-pass timestamps
        if (trace_prefix && npass <= 64) {
            if (!s->trace) PCD_CUDA(cudaMalloc(&s->trace, sizeof(unsigned long long) * 4 * WAVE_MAX_CTAS * 64));
            PCD_CUDA(cudaMemsetAsync(s->trace, 0, sizeof(unsigned long long) * 4 * WAVE_MAX_CTAS * 64, s->stream));
            pr.trace = s->trace;
            s->trace_npass = npass;
        }
        PCD_TRY(tiled_run_peer(s->D, s->W, s->H, s->row0, s->rows, s->row0 - s->GH, spp, s->sweep_max + slot + j, pr,
                               s->dmap_ok ? s->dmap : nullptr, s->sm_count, s->sm_reserve, s->stream));
        s->seq += (unsigned)npass;
        s->cur ^= (npass & 1);
        s->launches++;
        j += npass * spp;
    }
    return PCD_OK;
}

// waits for the slab's stream; *timed_out = 1 when a pass gave up waiting for a neighbour (results are then invalid)
int pcd_slab_peer_status(pcd_slab *s, int *timed_out) {
    if (!s || !timed_out) { set_error("null argument"); return PCD_ERR_INVALID; }
    PCD_TRY(select_device(s->device));
    PCD_CUDA(cudaMemcpyAsync(timed_out, s->ctl + CTL_ERR, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    PCD_CUDA(cudaStreamSynchronize(s->stream));
    return PCD_OK;
}

int pcd_slab_peer_error_to(pcd_slab *s, double *dst_dev) {
    if (!s || !dst_dev) { set_error("null argument"); return PCD_ERR_INVALID; }
    slab_error_to_kernel<<<1, 1, 0, s->stream>>>(s->ctl + CTL_ERR, dst_dev);
    PCD_LAUNCHED();
    return PCD_OK;
}

int pcd_slab_clear_max_range(pcd_slab *s, int first, int n_slots) {
    if (!s || first < 0 || n_slots < 0 || first + n_slots > s->ring) { set_error("bad slab / slot range"); return PCD_ERR_INVALID; }
    PCD_CUDA(cudaMemsetAsync(s->sweep_max + first, 0, sizeof(unsigned long long) * n_slots, s->stream));
    return PCD_OK;
}

int pcd_slab_clear_max(pcd_slab *s, int n_slots) {
    if (!s || n_slots < 0 || n_slots > s->ring) { set_error("bad slab / slot count"); return PCD_ERR_INVALID; }
    PCD_CUDA(cudaMemsetAsync(s->sweep_max, 0, sizeof(unsigned long long) * n_slots, s->stream));
    return PCD_OK;
}

}  // extern "C"
