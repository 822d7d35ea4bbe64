// K-SOR on a ROW SLAB of a larger grid (multi-GPU): the process owns global rows
// [row0, row0+rows) of a W x H grid plus one ghost row above and below (local row r <-> global row
// row0 + r - 1).  One launch updates the cells of one colour of the owned rows; the host layer exchanges
// the two boundary rows with the neighbouring ranks between colour phases (torch.distributed / NCCL over
// NVLink) and all-reduces the per-sweep max.  The per-cell arithmetic is sor_colour_kernel's, so a
// G-slab solve is bit-identical to the single-GPU solve (red-black updates of one colour are
// order-independent and max is exact).
#include "sor_common.cuh"

struct pcd_slab {
    int W = 0, H = 0, row0 = 0, rows = 0, device = 0;
    cudaStream_t stream = nullptr;   // caller's stream (e.g. torch's current stream), never owned
    double *phi = nullptr, *D = nullptr;            // (rows+2) x W each
    unsigned char *mask = nullptr;                  // (rows+2) x W neighbour masks (always built: D never changes)
    unsigned long long *sweep_max = nullptr;        // [ring]
    int ring = 0;
    long long launches = 0;
};

namespace pcd {

__global__ void slab_mask_kernel(const double *__restrict__ D, unsigned char *__restrict__ mask, int W, int H, int row0, int rows) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y + 1;  // owned local rows 1..rows
    if (x >= W) return;
    const int gy = row0 + r - 1;
    const size_t i = (size_t)r * W + x;
    unsigned m = 0;
    if (x != 0 && !isnan(D[i - 1])) m |= 1;
    if (gy != 0 && !isnan(D[i - W])) m |= 2;
    if (x != W - 1 && !isnan(D[i + 1])) m |= 4;
    if (gy != H - 1 && !isnan(D[i + W])) m |= 8;
    mask[i] = (unsigned char)m;
}

__global__ void __launch_bounds__(256)
sor_slab_colour_kernel(double *__restrict__ phi, const double *__restrict__ D, const unsigned char *__restrict__ mask,
                       int W, int row0, int colour, SorW w, unsigned long long *__restrict__ slot) {
    const int r = blockIdx.y + 1;
    const int gy = row0 + r - 1;
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

}  // namespace pcd

using namespace pcd;

extern "C" {

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
    s->stream = (cudaStream_t)cuda_stream;
    s->ring = 4096;
    const size_t n = (size_t)(rows + 2) * width;
    cudaError_t e = cudaMalloc(&s->phi, n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&s->D, n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&s->mask, n);
    if (e == cudaSuccess) e = cudaMalloc(&s->sweep_max, sizeof(unsigned long long) * s->ring);
    if (e == cudaSuccess) e = cudaMemset(s->phi, 0, n * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(s->D, 0, n * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(s->sweep_max, 0, sizeof(unsigned long long) * s->ring);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        set_error("slab allocation failed: %s", cudaGetErrorString(e));
        cudaFree(s->phi); cudaFree(s->D); cudaFree(s->mask); cudaFree(s->sweep_max);
        delete s;
        return PCD_ERR_CUDA;
    }
    *out = s;
    return PCD_OK;
}

void pcd_slab_destroy(pcd_slab *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    cudaFree(s->phi); cudaFree(s->D); cudaFree(s->mask); cudaFree(s->sweep_max);
    delete s;
}

int pcd_slab_device_ptrs(pcd_slab *s, void **phi_dev, void **D_dev, void **sweep_max_dev) {
    if (!s) { set_error("null slab"); return PCD_ERR_INVALID; }
    if (phi_dev) *phi_dev = s->phi;
    if (D_dev) *D_dev = s->D;
    if (sweep_max_dev) *sweep_max_dev = s->sweep_max;
    return PCD_OK;
}

// D_rows / phi_rows: host arrays of (rows+2) x W doubles INCLUDING the two ghost rows (ghost rows outside the
// grid are ignored); either may be NULL.
int pcd_slab_upload(pcd_slab *s, const double *D_rows, const double *phi_rows) {
    if (!s) { set_error("null slab"); return PCD_ERR_INVALID; }
    PCD_TRY(select_device(s->device));
    const size_t bytes = (size_t)(s->rows + 2) * s->W * sizeof(double);
    if (D_rows) {
        PCD_CUDA(cudaMemcpyAsync(s->D, D_rows, bytes, cudaMemcpyHostToDevice, s->stream));
        slab_mask_kernel<<<dim3((s->W + 255) / 256, s->rows), 256, 0, s->stream>>>(s->D, s->mask, s->W, s->H, s->row0, s->rows);
        PCD_LAUNCHED();
        s->launches++;
    }
    if (phi_rows) PCD_CUDA(cudaMemcpyAsync(s->phi, phi_rows, bytes, cudaMemcpyHostToDevice, s->stream));
    PCD_CUDA(cudaStreamSynchronize(s->stream));
    return PCD_OK;
}

int pcd_slab_download(pcd_slab *s, double *phi_owned_rows) {
    if (!s || !phi_owned_rows) { set_error("null argument"); return PCD_ERR_INVALID; }
    PCD_TRY(select_device(s->device));
    PCD_CUDA(cudaMemcpyAsync(phi_owned_rows, s->phi + s->W, (size_t)s->rows * s->W * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    PCD_CUDA(cudaStreamSynchronize(s->stream));
    return PCD_OK;
}

// one colour phase of sweep slot `slot` (0 <= slot < 4096); asynchronous on the slab's stream
int pcd_slab_sweep_colour(pcd_slab *s, int colour, int slot) {
    if (!s || slot < 0 || slot >= s->ring) { set_error("bad slab / slot"); return PCD_ERR_INVALID; }
    const int K = (s->W + 1) / 2;
    sor_slab_colour_kernel<<<dim3((K + 255) / 256, s->rows), 256, 0, s->stream>>>(s->phi, s->D, s->mask, s->W, s->row0, colour,
                                                                                   make_w(s->W), s->sweep_max + slot);
    PCD_LAUNCHED();
    s->launches++;
    return PCD_OK;
}

int pcd_slab_clear_max(pcd_slab *s, int n_slots) {
    if (!s || n_slots < 0 || n_slots > s->ring) { set_error("bad slab / slot count"); return PCD_ERR_INVALID; }
    PCD_CUDA(cudaMemsetAsync(s->sweep_max, 0, sizeof(unsigned long long) * n_slots, s->stream));
    return PCD_OK;
}

}  // extern "C"
