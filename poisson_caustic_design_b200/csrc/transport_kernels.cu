// Stage kernels of one optimal-transport iteration (src/caustic_design.cpp:190-266), everything
// device-resident: K-AREA/K-ERR, K-RAST (scatter rasteriser replacing the BVH), K-MEAN, K-STEP.
#include "common.cuh"

namespace pcd {

// ------------------------------------------------------------------------------------------------
// reproducible reductions
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RED_THREADS) sum_partials_kernel(const double *__restrict__ in, long n,
                                                                    double *__restrict__ partials) {
    __shared__ double scratch[32];
    double acc = 0.0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) acc += in[i];
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(RED_THREADS) sum_final_kernel(const double *__restrict__ partials, int n,
                                                                 double *__restrict__ out) {
    __shared__ double scratch[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) out[0] = acc;
}

int reduce_sum(const double *in, long n, double *partials, double *out, cudaStream_t st) {
    sum_partials_kernel<<<RED_BLOCKS, RED_THREADS, 0, st>>>(in, n, partials);
    PCD_LAUNCHED();
    sum_final_kernel<<<1, RED_THREADS, 0, st>>>(partials, RED_BLOCKS, out);
    PCD_LAUNCHED();
    return PCD_OK;
}

// subtractAverage (src/utils.cpp:60-86): mean over the non-NaN entries
__global__ void __launch_bounds__(RED_THREADS) nansum_partials_kernel(const double *__restrict__ in, long n,
                                                                       double *__restrict__ partials) {
    __shared__ double scratch[32];
    double acc = 0.0, cnt = 0.0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const double v = in[i];
        if (!isnan(v)) { acc += v; cnt += 1.0; }
    }
    acc = block_sum(acc, scratch);
    cnt = block_sum(cnt, scratch);
    if (threadIdx.x == 0) { partials[blockIdx.x] = acc; partials[gridDim.x + blockIdx.x] = cnt; }
}

__global__ void __launch_bounds__(RED_THREADS) mean_final_kernel(const double *__restrict__ partials, int n,
                                                                  double *__restrict__ out) {
    __shared__ double scratch[32];
    double acc = 0.0, cnt = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { acc += partials[i]; cnt += partials[n + i]; }
    acc = block_sum(acc, scratch);
    cnt = block_sum(cnt, scratch);
    if (threadIdx.x == 0) out[0] = acc / cnt;
}

__global__ void subtract_scalar_kernel(double *__restrict__ x, long n, const double *__restrict__ avg) {
    const double a = avg[0];
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const double v = x[i];
        if (!isnan(v)) x[i] = v - a;
    }
}

int k_subtract_average(pcd_ctx *c, double *grid) {
    nansum_partials_kernel<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(grid, c->N, c->partials);
    PCD_LAUNCHED();
    mean_final_kernel<<<1, RED_THREADS, 0, c->stream>>>(c->partials, RED_BLOCKS, c->d_scalars + 0);
    PCD_LAUNCHED();
    subtract_scalar_kernel<<<RED_BLOCKS * 2, 256, 0, c->stream>>>(grid, c->N, c->d_scalars + 0);
    PCD_LAUNCHED();
    return PCD_OK;
}

// ------------------------------------------------------------------------------------------------
// K-AREA + K-ERR: D_v = (target_v - area_v) / area_v over the median-dual cell of vertex v
//   src/mesh.cpp:174-231 (quads), src/polygon_utils.cpp:194-236,401-414 (signed shoelace),
//   src/caustic_design.cpp:199-209.  Nothing is materialised: each quad is rebuilt in registers.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double quad_area(double vx, double vy, double jx, double jy, double kx, double ky) {
    // [v, (v+j)/2, (v+j+k)/3, (v+k)/2], src/mesh.cpp:174-199
    const double x0 = vx, y0 = vy;
    const double x1 = (vx + jx) / 2.0, y1 = (vy + jy) / 2.0;
    const double x2 = (vx + jx + kx) / 3.0, y2 = (vy + jy + ky) / 3.0;
    const double x3 = (vx + kx) / 2.0, y3 = (vy + ky) / 2.0;
    double area = 0.0;
    area += (x0 * y1) - (x1 * y0);
    area += (x1 * y2) - (x2 * y1);
    area += (x2 * y3) - (x3 * y2);
    area += (x3 * y0) - (x0 * y3);
    return 0.5 * area;
}

__global__ void errors_kernel(const double *__restrict__ tx, const double *__restrict__ ty,
                              const double *__restrict__ target_areas, double *__restrict__ errors, int nx, int ny) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nx * ny) return;
    const int i = v / nx, j = v - i * nx;
    const double vx = tx[v], vy = ty[v];
    const bool up = i > 0, dn = i < ny - 1, lf = j > 0, rt = j < nx - 1;
    double area = 0.0;
    // adjacent triangles in ascending index; (v, j, k) = v and the next two vertices in triangle order
    if (up && lf) area += quad_area(vx, vy, tx[v - 1], ty[v - 1], tx[v - nx], ty[v - nx]);
    if (up && rt) {
        area += quad_area(vx, vy, tx[v - nx], ty[v - nx], tx[v - nx + 1], ty[v - nx + 1]);
        area += quad_area(vx, vy, tx[v - nx + 1], ty[v - nx + 1], tx[v + 1], ty[v + 1]);
    }
    if (dn && lf) {
        area += quad_area(vx, vy, tx[v + nx - 1], ty[v + nx - 1], tx[v - 1], ty[v - 1]);
        area += quad_area(vx, vy, tx[v + nx], ty[v + nx], tx[v + nx - 1], ty[v + nx - 1]);
    }
    if (dn && rt) area += quad_area(vx, vy, tx[v + 1], ty[v + 1], tx[v + nx], ty[v + nx]);
    errors[v] = (target_areas[v] - area) / area;
}

int k_errors(pcd_ctx *c) {
    errors_kernel<<<(c->V + 127) / 128, 128, 0, c->stream>>>(c->tx, c->ty, c->target_areas, c->errors,
                                                              c->cfg.mesh_res_x, c->cfg.mesh_res_y);
    PCD_LAUNCHED();
    return PCD_OK;
}

// ------------------------------------------------------------------------------------------------
// K-RAST: point location of a sample lattice in a (possibly folded) triangle mesh.
//   Reference: BVH build + query per sample (src/bvh.cpp), first hit in BVH order wins.
//   Here: pass 1 scatters every triangle over the lattice points of its bounding box and keeps the
//   LOWEST triangle index that passes the reference's inside test (atomicMin); pass 2 gathers.
//   Where triangles do not overlap (no folds) both rules select the same piecewise-linear value.
// ------------------------------------------------------------------------------------------------
__global__ void fill_int_kernel(int *p, long n, int v) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void raster_owner_kernel(const double *__restrict__ px, const double *__restrict__ py, int nx, int T,
                                    const double *__restrict__ xs, const double *__restrict__ ys, int SW, int SH,
                                    int *__restrict__ owner) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    int a, b, c;
    tri_vertices(t, nx, a, b, c);
    const double ax = px[a], ay = py[a], bx = px[b], by = py[b], cx = px[c], cy = py[c];
    const double minx = fmin(ax, fmin(bx, cx)), maxx = fmax(ax, fmax(bx, cx));
    const double miny = fmin(ay, fmin(by, cy)), maxy = fmax(ay, fmax(by, cy));
    const double x0 = xs[0], y0 = ys[0];
    const double dx = SW > 1 ? (xs[SW - 1] - x0) / (SW - 1) : 1.0, dy = SH > 1 ? (ys[SH - 1] - y0) / (SH - 1) : 1.0;
    // conservative index window (+-1 lattice step), the exact test below decides
    int jlo = (int)floor((minx - x0) / dx) - 1, jhi = (int)ceil((maxx - x0) / dx) + 1;
    int ilo = (int)floor((miny - y0) / dy) - 1, ihi = (int)ceil((maxy - y0) / dy) + 1;
    jlo = max(jlo, 0); ilo = max(ilo, 0); jhi = min(jhi, SW - 1); ihi = min(ihi, SH - 1);
    for (int i = ilo; i <= ihi; ++i) {
        const double y = ys[i];
        for (int j = jlo; j <= jhi; ++j) {
            double u, v, w;
            barycentric(ax, ay, bx, by, cx, cy, xs[j], y, u, v, w);
            if (bary_inside(u, v)) atomicMin(owner + (size_t)i * SW + j, t);
        }
    }
}

template <int NV>
__global__ void raster_gather_kernel(const double *__restrict__ px, const double *__restrict__ py, int nx,
                                     const double *__restrict__ xs, const double *__restrict__ ys, int SW, int SH,
                                     const int *__restrict__ owner, const double *__restrict__ val0,
                                     const double *__restrict__ val1, double *__restrict__ out0,
                                     double *__restrict__ out1, int *__restrict__ miss) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= SW) return;
    const size_t s = (size_t)i * SW + j;
    const int t = owner[s];
    if (t == OWNER_NONE) {
        out0[s] = NAN;
        if (NV > 1) out1[s] = NAN;
        atomicOr(miss, 1);
        return;
    }
    int a, b, c;
    tri_vertices(t, nx, a, b, c);
    double u, v, w;
    barycentric(px[a], py[a], px[b], py[b], px[c], py[c], xs[j], ys[i], u, v, w);
    out0[s] = val0[a] * u + val0[b] * v + val0[c] * w;  // src/mesh.cpp:266-269
    if (NV > 1) out1[s] = val1[a] * u + val1[b] * v + val1[c] * w;
}

// owner map of a lattice in mesh (px,py)
static int locate(pcd_ctx *c, const double *px, const double *py, const double *xs, const double *ys, int SW, int SH,
                  int *owner) {
    fill_int_kernel<<<RED_BLOCKS, 256, 0, c->stream>>>(owner, (long)SW * SH, OWNER_NONE);
    PCD_LAUNCHED();
    raster_owner_kernel<<<(c->T + 127) / 128, 128, 0, c->stream>>>(px, py, c->cfg.mesh_res_x, c->T, xs, ys, SW, SH, owner);
    PCD_LAUNCHED();
    return PCD_OK;
}

int check_miss(pcd_ctx *c, const char *what) {
    PCD_CUDA(cudaMemcpyAsync(c->h_flags, c->d_flags, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PCD_CUDA(cudaStreamSynchronize(c->stream));
    if (c->h_flags[0]) {
        set_error("interpolation miss in %s: a sample point lies in no triangle of the mesh", what);
        return PCD_ERR_RASTER_MISS;
    }
    return PCD_OK;
}

int k_raster_target(pcd_ctx *c) {
    const int W = c->cfg.res_x, H = c->cfg.res_y;
    PCD_CUDA(cudaMemsetAsync(c->d_flags, 0, sizeof(int), c->stream));
    PCD_TRY(locate(c, c->tx, c->ty, c->xs, c->ys, W, H, c->owner));
    raster_gather_kernel<1><<<dim3((W + 127) / 128, H), 128, 0, c->stream>>>(
        c->tx, c->ty, c->cfg.mesh_res_x, c->xs, c->ys, W, H, c->owner, c->errors, nullptr, c->raster, nullptr, c->d_flags);
    PCD_LAUNCHED();
    return PCD_OK;
}

// used by the height stage: two per-vertex fields of the REGULAR source mesh -> nodal rasters
int raster_source2(pcd_ctx *c, const double *v0, const double *v1, double *o0, double *o1) {
    const int W = c->cfg.res_x, H = c->cfg.res_y;
    if (!c->owner_src_valid) {  // source xy never change (src/mesh.cpp:20-22): locate once
        PCD_TRY(locate(c, c->sx, c->sy, c->xs, c->ys, W, H, c->owner_src));
        c->owner_src_valid = true;
    }
    raster_gather_kernel<2><<<dim3((W + 127) / 128, H), 128, 0, c->stream>>>(
        c->sx, c->sy, c->cfg.mesh_res_x, c->xs, c->ys, W, H, c->owner_src, v0, v1, o0, o1, c->d_flags);
    PCD_LAUNCHED();
    return PCD_OK;
}

// inverse transport map (src/mesh.cpp:348-409): locate the regular lattice in the deformed mesh and
// blend the regular positions; border vertices are snapped to the border (:381-402)
__global__ void inverse_map_kernel(const double *__restrict__ tx, const double *__restrict__ ty,
                                   const double *__restrict__ sx, const double *__restrict__ sy, int nx, int ny,
                                   const double *__restrict__ qxs, const double *__restrict__ qys,
                                   const int *__restrict__ owner, double width, double height,
                                   double *__restrict__ inv_x, double *__restrict__ inv_y, int *__restrict__ miss) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= nx) return;
    const int s = i * nx + j;
    const int t = owner[s];
    if (t == OWNER_NONE) {
        inv_x[s] = NAN; inv_y[s] = NAN;
        atomicOr(miss, 1);
        return;
    }
    int a, b, c;
    tri_vertices(t, nx, a, b, c);
    double u, v, w;
    barycentric(tx[a], ty[a], tx[b], ty[b], tx[c], ty[c], qxs[j], qys[i], u, v, w);
    double ix = sx[a] * u + sx[b] * v + sx[c] * w;
    double iy = sy[a] * u + sy[b] * v + sy[c] * w;
    if (j == 0) ix = 0; else if (j == nx - 1) ix = width;
    if (i == 0) iy = 0; else if (i == ny - 1) iy = height;
    inv_x[s] = ix; inv_y[s] = iy;
}

int k_inverse_map(pcd_ctx *c) {
    const int nx = c->cfg.mesh_res_x, ny = c->cfg.mesh_res_y;
    PCD_CUDA(cudaMemsetAsync(c->d_flags, 0, sizeof(int), c->stream));
    PCD_TRY(locate(c, c->tx, c->ty, c->qxs, c->qys, nx, ny, c->owner_v));
    inverse_map_kernel<<<dim3((nx + 127) / 128, ny), 128, 0, c->stream>>>(c->tx, c->ty, c->sx, c->sy, nx, ny, c->qxs, c->qys,
                                                                         c->owner_v, c->cfg.width, c->cfg.height,
                                                                         c->inv_x, c->inv_y, c->d_flags);
    PCD_LAUNCHED();
    return PCD_OK;
}

// ------------------------------------------------------------------------------------------------
// K-STEP: gradient (src/utils.cpp:3-20) evaluated on the fly at the 4 bilinear taps
//   (src/caustic_design.cpp:156-188,228-240), constrained vertex step (src/mesh.cpp:485-532),
//   max displacement (src/caustic_design.cpp:253-265).  One thread per vertex.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double grad_x_at(const double *__restrict__ g, int W, int x, int y) {
    return (g[(size_t)y * W + min(x + 1, W - 1)] - g[(size_t)y * W + max(x - 1, 0)]) / 2.0;
}
__device__ __forceinline__ double grad_y_at(const double *__restrict__ g, int W, int H, int x, int y) {
    return (g[(size_t)min(y + 1, H - 1) * W + x] - g[(size_t)max(y - 1, 0) * W + x]) / 2.0;
}

struct BilinearTaps {
    int x0, x1, y0, y1;
    double fx0, fx1, fy0, fy1;
};

__device__ __forceinline__ BilinearTaps bilinear_taps(double x, double y, int W, int H) {
    BilinearTaps t;
    t.x0 = min(max((int)floor(x), 0), W - 1);
    t.x1 = min(max((int)ceil(x), 0), W - 1);
    t.y0 = min(max((int)floor(y), 0), H - 1);
    t.y1 = min(max((int)ceil(y), 0), H - 1);
    t.fx1 = x - t.x0; t.fx0 = 1.0 - t.fx1;
    t.fy1 = y - t.y0; t.fy0 = 1.0 - t.fy1;
    return t;
}

__global__ void __launch_bounds__(128)
step_kernel(double *__restrict__ tx, double *__restrict__ ty, const double *__restrict__ phi, int nx, int ny, int W, int H,
            double width, double height, double min_t, double step_size, double *__restrict__ vgx,
            double *__restrict__ vgy, unsigned long long *__restrict__ max_bits) {
    __shared__ double scratch[32];
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    double dist = 0.0;
    if (v < nx * ny) {
        const double ox = tx[v], oy = ty[v];
        const BilinearTaps t = bilinear_taps((ox / width) * (W)-0.5, (oy / height) * (H)-0.5, W, H);
        double top = t.fx0 * grad_x_at(phi, W, t.x0, t.y0) + t.fx1 * grad_x_at(phi, W, t.x1, t.y0);
        double bot = t.fx0 * grad_x_at(phi, W, t.x0, t.y1) + t.fx1 * grad_x_at(phi, W, t.x1, t.y1);
        const double gx = t.fy0 * top + t.fy1 * bot;
        top = t.fx0 * grad_y_at(phi, W, H, t.x0, t.y0) + t.fx1 * grad_y_at(phi, W, H, t.x1, t.y0);
        bot = t.fx0 * grad_y_at(phi, W, H, t.x0, t.y1) + t.fx1 * grad_y_at(phi, W, H, t.x1, t.y1);
        const double gy = t.fy0 * top + t.fy1 * bot;
        vgx[v] = gx; vgy[v] = gy;
        const int i = v / nx, j = v - i * nx;
        const double vx = (j == 0 || j == nx - 1) ? 0.0 : gx;  // src/mesh.cpp:493-511
        const double vy = (i == 0 || i == ny - 1) ? 0.0 : gy;
        const double nxp = ox + vx * min_t * step_size, nyp = oy + vy * min_t * step_size;
        tx[v] = nxp; ty[v] = nyp;
        const double ddx = ox - nxp, ddy = oy - nyp, ddz = 0.0;
        dist = sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
        if (!(dist > 0.0)) dist = 0.0;
    }
    dist = block_max(dist, scratch);
    if (threadIdx.x == 0 && dist > 0.0) atomicMax(max_bits, (unsigned long long)__double_as_longlong(dist));
}

int k_step(pcd_ctx *c, double *step_out_host) {
    const int nx = c->cfg.mesh_res_x, ny = c->cfg.mesh_res_y;
    PCD_CUDA(cudaMemsetAsync(c->d_bits, 0, sizeof(unsigned long long), c->stream));
    const double min_t = c->cfg.width / nx;     // src/mesh.cpp:522
    const double step_size = (double)0.05f;     // src/caustic_design.cpp:250 passes the float literal 0.05f
    step_kernel<<<(c->V + 127) / 128, 128, 0, c->stream>>>(c->tx, c->ty, c->phi, nx, ny, c->cfg.res_x, c->cfg.res_y,
                                                            c->cfg.width, c->cfg.height, min_t, step_size, c->vgx, c->vgy,
                                                            c->d_bits);
    PCD_LAUNCHED();
    unsigned long long bits = 0;
    PCD_CUDA(cudaMemcpyAsync(c->h_scalars, c->d_bits, sizeof(bits), cudaMemcpyDeviceToHost, c->stream));
    PCD_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(&bits, c->h_scalars, sizeof(bits));
    double m;
    memcpy(&m, &bits, sizeof(m));
    *step_out_host = m / c->cfg.width;          // src/caustic_design.cpp:265
    return PCD_OK;
}

__global__ void gradient_kernel(const double *__restrict__ g, int W, int H, double *__restrict__ gx,
                                double *__restrict__ gy) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    gx[(size_t)y * W + x] = grad_x_at(g, W, x, y);
    gy[(size_t)y * W + x] = grad_y_at(g, W, H, x, y);
}

int k_gradient(pcd_ctx *c, const double *grid, double *gx, double *gy) {
    const int W = c->cfg.res_x, H = c->cfg.res_y;
    gradient_kernel<<<dim3((W + 127) / 128, H), 128, 0, c->stream>>>(grid, W, H, gx, gy);
    PCD_LAUNCHED();
    return PCD_OK;
}

}  // namespace pcd
