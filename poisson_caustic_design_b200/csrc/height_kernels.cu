// Normal-integration stage (Caustic_design::perform_height_map_iteration, src/caustic_design.cpp:269-332):
// K-INV (inverse transport map) -> K-NORM (refraction normals) -> K-RSRC (nodal rasters of the regular
// mesh) -> K-DIV (+mean removal) -> K-SOR(tol 1e-8, warm-started h) -> K-HGT (vertex heights).
#include "common.cuh"

namespace pcd {

int raster_source2(pcd_ctx *c, const double *v0, const double *v1, double *o0, double *o1);

// Mesh::calculate_refractive_normals_uniform, src/mesh.cpp:677-722 (normalize: src/utils.cpp:370-384)
__global__ void normals_kernel(const double *__restrict__ inv_x, const double *__restrict__ inv_y,
                               const double *__restrict__ sx, const double *__restrict__ sy,
                               const double *__restrict__ sz, int V, double focal_len, double refractive_index,
                               double *__restrict__ nx_out, double *__restrict__ ny_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    double t0 = inv_x[i] - sx[i], t1 = inv_y[i] - sy[i], t2 = 0 - sz[i] + focal_len;
    double squared_len = 0;
    squared_len += t0 * t0;
    squared_len += t1 * t1;
    squared_len += t2 * t2;
    const double len = sqrt(squared_len);
    t0 = t0 / len; t1 = t1 / len; t2 = t2 / len;
    const double x_normal = t0 - 0.0 * refractive_index;  // incident = (0,0,1)
    const double y_normal = t1 - 0.0 * refractive_index;
    const double z_normal = t2 - 1.0 * refractive_index;
    nx_out[i] = x_normal / z_normal;
    ny_out[i] = y_normal / z_normal;
}

// calculate_divergence, src/utils.cpp:22-39: central differences, 0 on the outermost ring
__global__ void divergence_kernel(const double *__restrict__ Nx, const double *__restrict__ Ny, int W, int H,
                                  double *__restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t i = (size_t)y * W + x;
    if (x == 0 || x == W - 1 || y == 0 || y == H - 1) {
        out[i] = 0.0;
    } else {
        const double dxx = (Nx[i + 1] - Nx[i - 1]) / 2.0;
        const double dyy = (Ny[i + W] - Ny[i - W]) / 2.0;
        out[i] = dxx + dyy;
    }
}

// bilinear sample of h at the source vertices (src/caustic_design.cpp:156-188,323-329) + running min
__global__ void __launch_bounds__(128)
vertex_height_kernel(const double *__restrict__ h, const double *__restrict__ sx, const double *__restrict__ sy, int V,
                     int W, int H, double width, double height, double *__restrict__ hv,
                     unsigned long long *__restrict__ min_key) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long key = ~0ull;
    if (i < V) {
        const double x = (sx[i] / width) * (W)-0.5, y = (sy[i] / height) * (H)-0.5;
        const int x0 = min(max((int)floor(x), 0), W - 1), x1 = min(max((int)ceil(x), 0), W - 1);
        const int y0 = min(max((int)floor(y), 0), H - 1), y1 = min(max((int)ceil(y), 0), H - 1);
        const double fx1 = x - x0, fx0 = 1.0 - fx1, fy1 = y - y0, fy0 = 1.0 - fy1;
        const double top = fx0 * h[(size_t)y0 * W + x0] + fx1 * h[(size_t)y0 * W + x1];
        const double bottom = fx0 * h[(size_t)y1 * W + x0] + fx1 * h[(size_t)y1 * W + x1];
        const double v = fy0 * top + fy1 * bottom;
        hv[i] = v;
        if (!isnan(v)) key = ordered_key(v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o);
        key = k2 < key ? k2 : key;
    }
    if ((threadIdx.x & 31) == 0) atomicMin(min_key, key);
}

// Mesh::set_source_heights, src/mesh.cpp:724-742: shift by min(0, min h_v), z <- h_v, sum of squared updates
__global__ void set_heights_kernel(const double *__restrict__ hv, double *__restrict__ sz, int V,
                                   const unsigned long long *__restrict__ min_key, double *__restrict__ upd) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const double mn = ordered_unkey(min_key[0]);
    const double max_h = mn < 0.0 ? mn : 0.0;  // the reference's (misnamed) max_h, :726-733
    const double nh = hv[i] - max_h;
    const double d = nh - sz[i];
    upd[i] = d * d;
    sz[i] = nh;
}

int k_height_iteration(pcd_ctx *c, double *update_sum_host) {
    const int W = c->cfg.res_x, H = c->cfg.res_y, V = c->V;
    cudaStream_t st = c->stream;
    PCD_TRY(k_inverse_map(c));
    PCD_TRY(check_miss(c, "inverse transport map"));
    const double focal_len = W / c->cfg.width * c->cfg.focal_l;  // src/caustic_design.cpp:271
    normals_kernel<<<(V + 255) / 256, 256, 0, st>>>(c->inv_x, c->inv_y, c->sx, c->sy, c->sz, V, focal_len, 1.49,
                                                    c->normals_x, c->normals_y);
    PCD_LAUNCHED();
    PCD_CUDA(cudaMemsetAsync(c->d_flags, 0, sizeof(int), st));
    PCD_TRY(raster_source2(c, c->normals_x, c->normals_y, c->norm_x, c->norm_y));
    PCD_TRY(check_miss(c, "source raster"));
    divergence_kernel<<<dim3((W + 127) / 128, H), 128, 0, st>>>(c->norm_x, c->norm_y, W, H, c->divergence);
    PCD_LAUNCHED();
    PCD_TRY(k_subtract_average(c, c->divergence));
    PCD_TRY(ctx_solve(c, c->divergence, c->h, c->height_tol));  // :311 (threshold: pcd_set_tolerances)
    const unsigned long long init_key = ~0ull;
    PCD_CUDA(cudaMemcpyAsync(c->d_bits + 2, &init_key, sizeof(init_key), cudaMemcpyHostToDevice, st));
    vertex_height_kernel<<<(V + 127) / 128, 128, 0, st>>>(c->h, c->sx, c->sy, V, W, H, c->cfg.width, c->cfg.height, c->hv,
                                                          c->d_bits + 2);
    PCD_LAUNCHED();
    set_heights_kernel<<<(V + 255) / 256, 256, 0, st>>>(c->hv, c->sz, V, c->d_bits + 2, c->inv_x /* scratch */);
    PCD_LAUNCHED();
    PCD_TRY(reduce_sum(c->inv_x, V, c->partials, c->d_scalars + 2, st));
    PCD_CUDA(cudaMemcpyAsync(c->h_scalars, c->d_scalars + 2, sizeof(double), cudaMemcpyDeviceToHost, st));
    PCD_CUDA(cudaStreamSynchronize(st));
    if (update_sum_host) *update_sum_host = c->h_scalars[0];
    return PCD_OK;
}

}  // namespace pcd
