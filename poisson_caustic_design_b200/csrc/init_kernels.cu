// One-time setup on the device (Caustic_design::initialize_solvers, src/caustic_design.cpp:334-364):
// min-max normalisation of the image, structured mesh, sample lattices, and K-TAREA -- the
// per-vertex target areas obtained by clipping each dual-cell quad against the pixel squares
// (src/polygon_utils.cpp:311-389) with an in-register Sutherland-Hodgman clipper (:14-136).
#include "common.cuh"

namespace pcd {

// ---- scale_matrix_proportional (src/utils.cpp:88-129) ----------------------------------------------
__global__ void minmax_kernel(const double *__restrict__ in, long n, unsigned long long *__restrict__ keys) {
    unsigned long long lo = ~0ull, hi = 0ull;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const double v = in[i];
        if (!isnan(v)) {
            const unsigned long long k = ordered_key(v);
            lo = k < lo ? k : lo;
            hi = k > hi ? k : hi;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(keys + 0, lo);
        atomicMax(keys + 1, hi);
    }
}

__global__ void scale_kernel(const double *__restrict__ in, long n, const unsigned long long *__restrict__ keys,
                             double lo, double hi, double *__restrict__ out) {
    const double mn = ordered_unkey(keys[0]), mx = ordered_unkey(keys[1]);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const double v = in[i];
        out[i] = isnan(v) ? 0.0 : lo + (hi - lo) * (v - mn) / (mx - mn);
    }
}

// ---- Mesh::generate_structured_mesh (src/mesh.cpp:45-64) + lattices ----------------------------------
__global__ void mesh_kernel(double *tx, double *ty, double *tz, double *sx, double *sy, double *sz, int nx, int ny,
                            double width, double height) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nx * ny) return;
    const int i = v / nx, j = v - i * nx;
    const double x = (double)j * width / (nx - 1), y = (double)i * height / (ny - 1);
    tx[v] = x; ty[v] = y; tz[v] = 0.0;
    sx[v] = x; sy[v] = y; sz[v] = 0.0;
}

__global__ void lattice_kernel(double *xs, double *ys, int W, int H, double *qxs, double *qys, int nx, int ny,
                               double width, double height) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double epsilon = 1e-8;
    // nodal raster samples, src/mesh.cpp:241-253
    if (i < W) xs[i] = (double)i * (width - epsilon) / (W - 1) + 0.5 * epsilon;
    if (i < H) ys[i] = (double)i * (height - epsilon) / (H - 1) + 0.5 * epsilon;
    // inverse-map queries at the (regular) source vertices, src/mesh.cpp:356-360
    if (i < nx) qxs[i] = epsilon + ((double)i * width / (nx - 1)) * ((width - 2 * epsilon) / width);
    if (i < ny) qys[i] = epsilon + ((double)i * height / (ny - 1)) * ((height - 2 * epsilon) / height);
}

// ---- Sutherland-Hodgman, restated from src/polygon_utils.cpp:14-136 ---------------------------------
struct V2 { double x, y; };
constexpr int CLIP_MAX = 12;  // a convex quad clipped by 4 half-planes has at most 8 vertices
struct Poly { int len; V2 v[CLIP_MAX]; };

__device__ __forceinline__ double cross2(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }

__device__ __forceinline__ int left_of(V2 a, V2 b, V2 c) {  // :33-41
    const V2 t1 = {b.x - a.x, b.y - a.y}, t2 = {c.x - b.x, c.y - b.y};
    const double x = cross2(t1, t2);
    return x < 0 ? -1 : x > 0;
}

__device__ __forceinline__ int line_sect(V2 x0, V2 x1, V2 y0, V2 y1, V2 &res) {  // :43-60
    const V2 dx = {x1.x - x0.x, x1.y - x0.y}, dy = {y1.x - y0.x, y1.y - y0.y}, d = {x0.x - y0.x, x0.y - y0.y};
    double dyx = cross2(dy, dx);
    if (!dyx) return 0;
    dyx = cross2(d, dx) / dyx;
    if (dyx <= 0 || dyx >= 1) return 0;
    res.x = y0.x + dyx * dy.x;
    res.y = y0.y + dyx * dy.y;
    return 1;
}

__device__ __forceinline__ void poly_append(Poly &p, V2 v) {
    if (p.len < CLIP_MAX) p.v[p.len++] = v;
}

__device__ void poly_edge_clip(const Poly &sub, V2 x0, V2 x1, int left, Poly &res) {  // :94-116
    V2 tmp;
    V2 v0 = sub.v[sub.len - 1], v1;
    res.len = 0;
    int side0 = left_of(x0, x1, v0), side1;
    if (side0 != -left) poly_append(res, v0);
    for (int i = 0; i < sub.len; i++) {
        v1 = sub.v[i];
        side1 = left_of(x0, x1, v1);
        if (side0 + side1 == 0 && side0)
            if (line_sect(x0, x1, v0, v1, tmp)) poly_append(res, tmp);
        if (i == sub.len - 1) break;
        if (side1 != -left) poly_append(res, v1);
        v0 = v1;
        side0 = side1;
    }
}

__device__ double clip_area(const Poly &sub, const V2 *clip) {  // poly_clip :118-136 + calculate_polygon_area :172-192
    Poly a, b;
    Poly *p1 = &a, *p2 = &b, *tmp;
    a.len = 0; b.len = 0;
    const int dir = left_of(clip[0], clip[1], clip[2]);
    poly_edge_clip(sub, clip[3], clip[0], dir, *p2);
    for (int i = 0; i < 3; i++) {
        tmp = p2; p2 = p1; p1 = tmp;
        if (p1->len == 0) { p2->len = 0; break; }
        poly_edge_clip(*p1, clip[i], clip[i + 1], dir, *p2);
    }
    const int n = p2->len;
    if (n < 3) return 0.0;
    double area = 0.0;
    for (int i = 0; i < n; i++) {
        const int j = (i + 1) % n;
        area += (p2->v[i].x * p2->v[j].y) - (p2->v[j].x * p2->v[i].y);
    }
    return 0.5 * area;
}

// integrate_cell_intensities, src/polygon_utils.cpp:311-366 (pixel side = width/image_w on BOTH axes, :327)
__device__ double integrate_quad(const double *__restrict__ image, const Poly &quad, int image_w, int image_h,
                                 double width) {
    double xmin = quad.v[0].x, xmax = xmin, ymin = quad.v[0].y, ymax = ymin;
    for (int i = 1; i < 4; ++i) {
        xmin = fmin(xmin, quad.v[i].x); xmax = fmax(xmax, quad.v[i].x);
        ymin = fmin(ymin, quad.v[i].y); ymax = fmax(ymax, quad.v[i].y);
    }
    double intensity = 0.0;
    const double px = width / ((double)image_w);
    const int y_begin = (int)fmax(floor(ymin / px), 0.0), x_begin = (int)fmax(floor(xmin / px), 0.0);
    const double y_end = fmin(ceil(ymax / px), (double)image_h), x_end = fmin(ceil(xmax / px), (double)image_w);
    for (int y = y_begin; y < y_end; ++y)
        for (int x = x_begin; x < x_end; x++) {
            double cx = (double)x + 0.5, cy = (double)y + 0.5;
            cx *= px; cy *= px;
            V2 sq[4];
            sq[0].x = cx - px / 2.0; sq[0].y = cy - px / 2.0;
            sq[1].x = cx - px / 2.0; sq[1].y = cy + px / 2.0;
            sq[2].x = cx + px / 2.0; sq[2].y = cy + px / 2.0;
            sq[3].x = cx + px / 2.0; sq[3].y = cy - px / 2.0;
            intensity += clip_area(quad, sq) * image[(size_t)y * image_w + x];
        }
    return intensity;
}

__device__ __forceinline__ void make_quad(Poly &q, double vx, double vy, double jx, double jy, double kx, double ky) {
    q.len = 4;
    q.v[0].x = vx;                     q.v[0].y = vy;
    q.v[1].x = (vx + jx) / 2.0;        q.v[1].y = (vy + jy) / 2.0;
    q.v[2].x = (vx + jx + kx) / 3.0;   q.v[2].y = (vy + jy + ky) / 3.0;
    q.v[3].x = (vx + kx) / 2.0;        q.v[3].y = (vy + ky) / 2.0;
}

// get_target_partitioned_areas, src/polygon_utils.cpp:368-389 (before normalisation).
// One thread per (vertex, adjacent triangle) pair: slot = 6*v + a, a = position in ascending triangle order.
__global__ void __launch_bounds__(128)
target_quads_kernel(const double *__restrict__ tx, const double *__restrict__ ty, const double *__restrict__ pixels,
                    int nx, int ny, int W, int H, double width, double *__restrict__ quad_int) {
    const long id = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= 6L * nx * ny) return;
    const int v = (int)(id / 6), a = (int)(id - 6L * v);
    const int i = v / nx, j = v - i * nx;
    const bool up = i > 0, dn = i < ny - 1, lf = j > 0, rt = j < nx - 1;
    int vj = -1, vk = -1;
    switch (a) {
        case 0: if (up && lf) { vj = v - 1;       vk = v - nx; } break;
        case 1: if (up && rt) { vj = v - nx;      vk = v - nx + 1; } break;
        case 2: if (up && rt) { vj = v - nx + 1;  vk = v + 1; } break;
        case 3: if (dn && lf) { vj = v + nx - 1;  vk = v - 1; } break;
        case 4: if (dn && lf) { vj = v + nx;      vk = v + nx - 1; } break;
        default: if (dn && rt) { vj = v + 1;      vk = v + nx; } break;
    }
    double r = 0.0;
    if (vj >= 0) {
        Poly q;
        make_quad(q, tx[v], ty[v], tx[vj], ty[vj], tx[vk], ty[vk]);
        r = integrate_quad(pixels, q, W, H, width);
    }
    quad_int[id] = r;
}

__global__ void target_cells_kernel(const double *__restrict__ quad_int, int V, double *__restrict__ target_areas) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    double total = 0.0;
#pragma unroll
    for (int a = 0; a < 6; ++a) total += quad_int[6L * v + a];
    target_areas[v] = total;
}

__global__ void scale_by_sum_kernel(double *__restrict__ x, int n, const double *__restrict__ sum, double numer) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double scaling = numer / sum[0];  // src/polygon_utils.cpp:382
    x[i] *= scaling;
}

int k_init(pcd_ctx *c, const double *image_host) {
    const int nx = c->cfg.mesh_res_x, ny = c->cfg.mesh_res_y, W = c->cfg.res_x, H = c->cfg.res_y;
    cudaStream_t st = c->stream;
    // image -> raster scratch -> pixels (min-max scaled to [0,1], :335)
    PCD_CUDA(cudaMemcpyAsync(c->raster, image_host, sizeof(double) * c->N, cudaMemcpyHostToDevice, st));
    const unsigned long long init_keys[2] = {~0ull, 0ull};
    PCD_CUDA(cudaMemcpyAsync(c->d_bits, init_keys, sizeof(init_keys), cudaMemcpyHostToDevice, st));
    minmax_kernel<<<RED_BLOCKS, 256, 0, st>>>(c->raster, c->N, c->d_bits);
    PCD_LAUNCHED();
    scale_kernel<<<RED_BLOCKS * 2, 256, 0, st>>>(c->raster, c->N, c->d_bits, 0.0, (double)1.0f, c->pixels);
    PCD_LAUNCHED();
    mesh_kernel<<<(c->V + 255) / 256, 256, 0, st>>>(c->tx, c->ty, c->tz, c->sx, c->sy, c->sz, nx, ny, c->cfg.width, c->cfg.height);
    PCD_LAUNCHED();
    int m = W > H ? W : H;
    m = m > nx ? m : nx;
    m = m > ny ? m : ny;
    lattice_kernel<<<(m + 255) / 256, 256, 0, st>>>(c->xs, c->ys, W, H, c->qxs, c->qys, nx, ny, c->cfg.width, c->cfg.height);
    PCD_LAUNCHED();
    // K-TAREA; the 6V per-quad integrals go through norm_x (N >= 6V is not guaranteed -> own scratch)
    double *quad_int = nullptr;
    PCD_CUDA(cudaMalloc(&quad_int, sizeof(double) * 6 * (size_t)c->V));
    target_quads_kernel<<<(unsigned)((6L * c->V + 127) / 128), 128, 0, st>>>(c->tx, c->ty, c->pixels, nx, ny, W, H, c->cfg.width, quad_int);
    PCD_LAUNCHED();
    target_cells_kernel<<<(c->V + 255) / 256, 256, 0, st>>>(quad_int, c->V, c->target_areas);
    PCD_LAUNCHED();
    PCD_TRY(reduce_sum(c->target_areas, c->V, c->partials, c->d_scalars + 1, st));
    scale_by_sum_kernel<<<(c->V + 255) / 256, 256, 0, st>>>(c->target_areas, c->V, c->d_scalars + 1, c->cfg.width * c->cfg.height);
    PCD_LAUNCHED();
    // phi, h <- 0 (:354-363)
    PCD_CUDA(cudaMemsetAsync(c->phi, 0, sizeof(double) * c->N, st));
    PCD_CUDA(cudaMemsetAsync(c->h, 0, sizeof(double) * c->N, st));
    PCD_CUDA(cudaStreamSynchronize(st));
    PCD_CUDA(cudaFree(quad_int));
    c->owner_src_valid = false;
    return PCD_OK;
}

}  // namespace pcd
