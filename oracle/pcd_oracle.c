/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference hot path (see pcd_oracle.h).
 *
 * Style: flat row-major arrays instead of vector<vector<double>>, SoA points instead of
 * vector<point_t>; every floating-point expression keeps the reference's operand order so that
 * an -O2 build without FMA contraction reproduces the reference bit for bit.
 */
#include "pcd_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ===================================================================================== */
/* Poisson solver                                                                          */
/* ===================================================================================== */

/* src/solver.cpp:12-61 (patial_relax), one cell. */
static inline double relax_cell(const double *D, double *phi, int W, int H, int x, int y, double omega,
                                double *max_update) {
    size_t i = (size_t)y * W + x;
    double val = phi[i];
    double neighbor_sum = 0.0, neighbor_cnt = 0.0;
    if (x != 0 && !isnan(D[i - 1])) { neighbor_cnt += 1.0; neighbor_sum += phi[i - 1]; }          /* :29-32 */
    if (y != 0 && !isnan(D[i - W])) { neighbor_cnt += 1.0; neighbor_sum += phi[i - W]; }          /* :33-36 */
    if (x != W - 1 && !isnan(D[i + 1])) { neighbor_cnt += 1.0; neighbor_sum += phi[i + 1]; }      /* :37-40 */
    if (y != H - 1 && !isnan(D[i + W])) { neighbor_cnt += 1.0; neighbor_sum += phi[i + W]; }      /* :41-44 */
    double delta = omega / neighbor_cnt * (neighbor_sum - neighbor_cnt * val - D[i]);             /* :47 */
    double abs_delta = fabs(delta);
    if (abs_delta > *max_update) *max_update = abs_delta;                                         /* :50-53 */
    phi[i] += delta;                                                                              /* :56 */
    return delta;
}

int pcdo_poisson_lex(const double *D, double *phi, int W, int H, int max_iterations, double tol,
                     double *last_max_update) {
    double omega = 2.0 / (1.0 + 3.14159265 / W);                                                  /* :71 */
    int sweeps = 0;
    double max_update = 0.0;
    for (int it = 0; it < max_iterations; ++it) {                                                 /* :92 */
        max_update = 0.0;
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) relax_cell(D, phi, W, H, x, y, omega, &max_update);
        ++sweeps;
        if (max_update < tol) break;                                                              /* :142-146 */
    }
    if (last_max_update) *last_max_update = max_update;
    return sweeps;
}

int pcdo_poisson_rb(const double *D, double *phi, int W, int H, int max_iterations, double tol,
                    int extra_sweeps, int *converged_at, double *last_max_update) {
    double omega = 2.0 / (1.0 + 3.14159265 / W);
    int sweeps = 0, conv = 0, remaining = -1;
    double max_update = 0.0;
    for (int it = 0; it < max_iterations; ++it) {
        max_update = 0.0;
        for (int colour = 0; colour < 2; ++colour)
            for (int y = 0; y < H; ++y)
                for (int x = (y + colour) & 1; x < W; x += 2) relax_cell(D, phi, W, H, x, y, omega, &max_update);
        ++sweeps;
        if (remaining < 0) {
            if (max_update < tol) {
                conv = sweeps;
                remaining = extra_sweeps;
            }
        } else {
            --remaining;
        }
        if (remaining == 0) break;
    }
    if (converged_at) *converged_at = conv;
    if (last_max_update) *last_max_update = max_update;
    return sweeps;
}

/* ===================================================================================== */
/* src/utils.cpp numeric helpers                                                           */
/* ===================================================================================== */

void pcdo_subtract_average(double *raster, long n) {                                    /* :60-86 */
    double sum = 0.0;
    int count = 0;
    for (long i = 0; i < n; ++i)
        if (!isnan(raster[i])) { sum += raster[i]; count++; }
    double average = sum / count;
    for (long i = 0; i < n; ++i)
        if (!isnan(raster[i])) raster[i] = raster[i] - average;
}

void pcdo_gradient(const double *g, int W, int H, double *gx, double *gy) {             /* :3-20 */
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int xp = x + 1 < W - 1 ? x + 1 : W - 1, xm = x - 1 > 0 ? x - 1 : 0;
            int yp = y + 1 < H - 1 ? y + 1 : H - 1, ym = y - 1 > 0 ? y - 1 : 0;
            gx[(size_t)y * W + x] = (g[(size_t)y * W + xp] - g[(size_t)y * W + xm]) / 2.0;
            gy[(size_t)y * W + x] = (g[(size_t)yp * W + x] - g[(size_t)ym * W + x]) / 2.0;
        }
}

void pcdo_divergence(const double *nx, const double *ny, int W, int H, double *out) {   /* :22-39 */
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            size_t i = (size_t)y * W + x;
            if (x == 0 || x == W - 1 || y == 0 || y == H - 1) {
                out[i] = 0.0;
            } else {
                double dxx = (nx[i + 1] - nx[i - 1]) / 2.0;
                double dyy = (ny[i + W] - ny[i - W]) / 2.0;
                out[i] = dxx + dyy;
            }
        }
}

void pcdo_scale_matrix_proportional(const double *m, long n, double lo, double hi, double *out) { /* :88-129 */
    double mn = m[0], mx = m[0];
    for (long i = 0; i < n; ++i)
        if (!isnan(m[i])) {
            if (m[i] < mn) mn = m[i];
            if (m[i] > mx) mx = m[i];
        }
    for (long i = 0; i < n; ++i)
        out[i] = isnan(m[i]) ? 0.0 : lo + (hi - lo) * (m[i] - mn) / (mx - mn);
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

double pcdo_bilinear(const double *img, int W, int H, double x, double y) {  /* caustic_design.cpp:156-188 */
    int x0 = (int)floor(x), y0 = (int)floor(y), x1 = (int)ceil(x), y1 = (int)ceil(y);
    x0 = clampi(x0, 0, W - 1); x1 = clampi(x1, 0, W - 1);
    y0 = clampi(y0, 0, H - 1); y1 = clampi(y1, 0, H - 1);
    double fx1 = x - x0, fx0 = 1.0 - fx1;
    double fy1 = y - y0, fy0 = 1.0 - fy1;
    double top = fx0 * img[(size_t)y0 * W + x0] + fx1 * img[(size_t)y0 * W + x1];
    double bottom = fx0 * img[(size_t)y1 * W + x0] + fx1 * img[(size_t)y1 * W + x1];
    return fy0 * top + fy1 * bottom;
}

/* ===================================================================================== */
/* Design state                                                                            */
/* ===================================================================================== */

typedef struct { double min_x, min_y, max_x, max_y; int first_child, first_face, nb_faces, is_leaf; } bvh_node;

typedef struct {
    bvh_node *nodes; int n_nodes, cap_nodes;
    double *cx, *cy;      /* centroids, permuted by split() */
    int *ids;             /* sorted_triangle_ids */
} bvh_t;

struct pcdo_design {
    int nx, ny, W, H, V, T;
    double width, height, focal_l, thickness;
    int solver_mode, last_sweeps;
    int *tri;                              /* T x 3, src/mesh.cpp:56-63 */
    int *adj; int *adj_n;                  /* per vertex: adjacent triangles in the reference's iteration order */
    double *tx, *ty, *tz, *sx, *sy, *sz;   /* target_points / source_points */
    double *pixels, *target_areas, *errors, *raster, *phi, *h, *gx, *gy, *vgx, *vgy;
    double *normals_x, *normals_y, *norm_x, *norm_y, *divergence;
    bvh_t bvh;
};

static double *dalloc(size_t n) { return (double *)calloc(n ? n : 1, sizeof(double)); }

/* src/mesh.cpp:106-147.  vertex_to_triangles[v] lists triangles in ascending index; the
 * reference then copies them through a std::unordered_set<int>, whose iteration order in
 * libstdc++ (13 buckets for <= 6 keys, identity hash, new key goes to the front of its bucket's
 * run, a new bucket's run goes to the front of the list) decides the summation order of the
 * cell's quads.  Reproduced here so that `errors` match bit for bit. */
static int unordered_set_order(const int *ins, int n, int *out) {
    int list[8], len = 0;
    for (int k = 0; k < n; ++k) {
        int key = ins[k], b = key % 13, pos = -1;
        for (int i = 0; i < len; ++i)
            if (list[i] % 13 == b) { pos = i; break; }
        if (pos < 0) pos = 0;
        for (int i = len; i > pos; --i) list[i] = list[i - 1];
        list[pos] = key;
        ++len;
    }
    for (int i = 0; i < len; ++i) out[i] = list[i];
    return len;
}

pcdo_design *pcdo_create(int mesh_nx, int mesh_ny, int res_x, int res_y, double width, double height,
                         double focal_l, double thickness, int solver_mode) {
    pcdo_design *d = (pcdo_design *)calloc(1, sizeof(pcdo_design));
    d->nx = mesh_nx; d->ny = mesh_ny; d->W = res_x; d->H = res_y;
    d->width = width; d->height = height; d->focal_l = focal_l; d->thickness = thickness;
    d->solver_mode = solver_mode;
    d->V = mesh_nx * mesh_ny;
    d->T = 2 * (mesh_nx - 1) * (mesh_ny - 1);
    return d;
}

void pcdo_destroy(pcdo_design *d) {
    if (!d) return;
    free(d->tri); free(d->adj); free(d->adj_n);
    free(d->tx); free(d->ty); free(d->tz); free(d->sx); free(d->sy); free(d->sz);
    free(d->pixels); free(d->target_areas); free(d->errors); free(d->raster); free(d->phi); free(d->h);
    free(d->gx); free(d->gy); free(d->vgx); free(d->vgy);
    free(d->normals_x); free(d->normals_y); free(d->norm_x); free(d->norm_y); free(d->divergence);
    free(d->bvh.nodes); free(d->bvh.cx); free(d->bvh.cy); free(d->bvh.ids);
    free(d);
}

/* ---- dual-cell quads ------------------------------------------------------------------ */

/* src/mesh.cpp:174-199 (get_triangle_quad) with edge_centroid/triangle_centroid :154-172 */
static void triangle_quad(const pcdo_design *d, const double *px, const double *py, int v, int t, double qx[4],
                          double qy[4]) {
    const int *tr = d->tri + 3 * t;
    int j = -1, k = -1;
    for (int i = 0; i < 3; ++i)
        if (tr[i] == v) { j = tr[(i + 1) % 3]; k = tr[(i + 2) % 3]; break; }
    qx[0] = px[v];                               qy[0] = py[v];
    qx[1] = (px[v] + px[j]) / 2.0;               qy[1] = (py[v] + py[j]) / 2.0;
    qx[2] = (px[v] + px[j] + px[k]) / 3.0;       qy[2] = (py[v] + py[j] + py[k]) / 3.0;
    qx[3] = (px[v] + px[k]) / 2.0;               qy[3] = (py[v] + py[k]) / 2.0;
}

/* src/polygon_utils.cpp:194-214 / :172-192 (signed shoelace) */
static double shoelace(const double *x, const double *y, int n) {
    if (n < 3) return 0.0;
    double area = 0.0;
    for (int i = 0; i < n; ++i) {
        int j = (i + 1) % n;
        area += (x[i] * y[j]) - (x[j] * y[i]);
    }
    return 0.5 * area;
}

/* ---- Sutherland-Hodgman (src/polygon_utils.cpp:14-136) -------------------------------- */
typedef struct { double x, y; } vec_t;
#define POLY_MAX 32
typedef struct { int len; vec_t v[POLY_MAX]; } poly_t;

static double cross2(const vec_t *a, const vec_t *b) { return a->x * b->y - a->y * b->x; }

static int left_of(const vec_t *a, const vec_t *b, const vec_t *c) {                  /* :33-41 */
    vec_t t1 = {b->x - a->x, b->y - a->y}, t2 = {c->x - b->x, c->y - b->y};
    double x = cross2(&t1, &t2);
    return x < 0 ? -1 : x > 0;
}

static int line_sect(const vec_t *x0, const vec_t *x1, const vec_t *y0, const vec_t *y1, vec_t *res) { /* :43-60 */
    vec_t dx = {x1->x - x0->x, x1->y - x0->y}, dy = {y1->x - y0->x, y1->y - y0->y};
    vec_t dd = {x0->x - y0->x, x0->y - y0->y};
    double dyx = cross2(&dy, &dx);
    if (!dyx) return 0;
    dyx = cross2(&dd, &dx) / dyx;
    if (dyx <= 0 || dyx >= 1) return 0;
    res->x = y0->x + dyx * dy.x;
    res->y = y0->y + dyx * dy.y;
    return 1;
}

static void poly_append(poly_t *p, const vec_t *v) { if (p->len < POLY_MAX) p->v[p->len++] = *v; }

static void poly_edge_clip(const poly_t *sub, const vec_t *x0, const vec_t *x1, int left, poly_t *res) { /* :94-116 */
    vec_t tmp;
    const vec_t *v0 = sub->v + sub->len - 1, *v1;
    res->len = 0;
    int side0 = left_of(x0, x1, v0), side1;
    if (side0 != -left) poly_append(res, v0);
    for (int i = 0; i < sub->len; i++) {
        v1 = sub->v + i;
        side1 = left_of(x0, x1, v1);
        if (side0 + side1 == 0 && side0)
            if (line_sect(x0, x1, v0, v1, &tmp)) poly_append(res, &tmp);
        if (i == sub->len - 1) break;
        if (side1 != -left) poly_append(res, v1);
        v0 = v1;
        side0 = side1;
    }
}

static void poly_clip(const poly_t *sub, const poly_t *clip, poly_t *out) {          /* :118-136 */
    poly_t a, b, *p1 = &a, *p2 = &b, *tmp;
    p1->len = 0; p2->len = 0;
    int dir = left_of(clip->v, clip->v + 1, clip->v + 2);                           /* poly_winding :89-92 */
    poly_edge_clip(sub, clip->v + clip->len - 1, clip->v, dir, p2);
    for (int i = 0; i < clip->len - 1; i++) {
        tmp = p2; p2 = p1; p1 = tmp;
        if (p1->len == 0) { p2->len = 0; break; }
        poly_edge_clip(p1, clip->v + i, clip->v + i + 1, dir, p2);
    }
    *out = *p2;
}

static double poly_area(const poly_t *p) {                                          /* :172-192 */
    if (p->len < 3) return 0.0;
    double area = 0.0;
    for (int i = 0; i < p->len; i++) {
        int j = (i + 1) % p->len;
        area += (p->v[i].x * p->v[j].y) - (p->v[j].x * p->v[i].y);
    }
    return 0.5 * area;
}

/* src/polygon_utils.cpp:311-366 */
static double integrate_cell_intensities(const double *image, const double *qx, const double *qy, int image_w,
                                         int image_h, double width) {
    poly_t polygon; polygon.len = 0;
    for (int i = 0; i < 4; ++i) { vec_t p = {qx[i], qy[i]}; poly_append(&polygon, &p); }
    double xmin = qx[0], xmax = qx[0], ymin = qy[0], ymax = qy[0];                   /* :138-170 */
    for (int i = 1; i < 4; ++i) {
        if (qx[i] < xmin) xmin = qx[i]; else if (qx[i] > xmax) xmax = qx[i];
        if (qy[i] < ymin) ymin = qy[i]; else if (qy[i] > ymax) ymax = qy[i];
    }
    double intensity = 0.0;
    double px = width / ((double)image_w);                                          /* :327 */
    int y_begin = (int)fmax(floor(ymin / px), 0.0f), x_begin = (int)fmax(floor(xmin / px), 0.0f);
    double y_end = fmin(ceil(ymax / px), image_h), x_end = fmin(ceil(xmax / px), image_w);
    for (int y = y_begin; y < y_end; ++y) {
        for (int x = x_begin; x < x_end; x++) {
            double cx = (double)x + 0.5, cy = (double)y + 0.5;
            cx *= px; cy *= px;
            poly_t pixel; pixel.len = 4;
            pixel.v[0].x = cx - px / 2.0f; pixel.v[0].y = cy - px / 2.0f;
            pixel.v[1].x = cx - px / 2.0f; pixel.v[1].y = cy + px / 2.0f;
            pixel.v[2].x = cx + px / 2.0f; pixel.v[2].y = cy + px / 2.0f;
            pixel.v[3].x = cx + px / 2.0f; pixel.v[3].y = cy - px / 2.0f;
            poly_t result;
            poly_clip(&polygon, &pixel, &result);
            intensity += poly_area(&result) * image[(size_t)y * image_w + x];
        }
    }
    return intensity;
}

/* ---- BVH (src/bvh.cpp) ---------------------------------------------------------------- */

static int bvh_split(bvh_t *b, int start, int end, int dim, float split_value) {     /* :50-76 */
    double *c = dim == 0 ? b->cx : b->cy;
    int left = start, right = end - 1;
    while (left < right) {
        while (left < end && c[left] < split_value) left += 1;
        while (right >= start && c[right] >= split_value) right -= 1;
        if (left >= right) break;
        double t;
        t = b->cx[left]; b->cx[left] = b->cx[right]; b->cx[right] = t;
        t = b->cy[left]; b->cy[left] = b->cy[right]; b->cy[right] = t;
        int ti = b->ids[left]; b->ids[left] = b->ids[right]; b->ids[right] = ti;
        ++left; --right;
    }
    /* the reference reads centroids[left] here even when left == end; both outcomes of that
     * comparison return `end`, so this guard is value-equivalent wherever the reference is defined */
    if (left >= end) return end;
    return c[left] <= split_value ? (end < left + 1 ? end : left + 1) : left;
}

static void bvh_build_node(bvh_t *b, const int *tri, const double *px, const double *py, int node, int start,
                           int end, int level, int target_cell_size, int max_depth) {  /* :78-160 */
    double min_x = INFINITY, min_y = INFINITY, max_x = -INFINITY, max_y = -INFINITY;
    for (int i = start; i < end; i++)
        for (int j = 0; j < 3; j++) {
            int p = tri[3 * b->ids[i] + j];
            min_x = fmin(min_x, px[p]); min_y = fmin(min_y, py[p]);
            max_x = fmax(max_x, px[p]); max_y = fmax(max_y, py[p]);
        }
    double epsilon = DBL_EPSILON;
    min_x -= 0.5 * epsilon; min_y -= 0.5 * epsilon; max_x += 0.5 * epsilon; max_y += 0.5 * epsilon;
    b->nodes[node].min_x = min_x; b->nodes[node].min_y = min_y;
    b->nodes[node].max_x = max_x; b->nodes[node].max_y = max_y;
    if (end - start <= target_cell_size || level >= max_depth) {
        b->nodes[node].is_leaf = 1; b->nodes[node].first_face = start;
        b->nodes[node].nb_faces = end - start > 0 ? end - start : 0;
        return;
    }
    b->nodes[node].is_leaf = 0;
    int dim = 0;
    if ((max_x - min_x) < (max_y - min_y)) dim = 1;
    double split_value = dim == 0 ? 0.5f * (max_x + min_x) : 0.5f * (max_y + min_y);
    int mid = bvh_split(b, start, end, dim, (float)split_value);                     /* float parameter, bvh.h:36 */
    if (mid == start || mid == end) {
        b->nodes[node].is_leaf = 1; b->nodes[node].first_face = start;
        b->nodes[node].nb_faces = end - start > 0 ? end - start : 0;
        return;
    }
    int child = b->n_nodes;
    b->nodes[node].first_child = child;
    if (b->n_nodes + 2 > b->cap_nodes) {
        b->cap_nodes = 2 * b->cap_nodes + 2;
        b->nodes = (bvh_node *)realloc(b->nodes, sizeof(bvh_node) * (size_t)b->cap_nodes);
    }
    memset(b->nodes + b->n_nodes, 0, 2 * sizeof(bvh_node));
    b->n_nodes += 2;
    bvh_build_node(b, tri, px, py, child, start, mid, level + 1, target_cell_size, max_depth);
    bvh_build_node(b, tri, px, py, child + 1, mid, end, level + 1, target_cell_size, max_depth);
}

static void bvh_build(pcdo_design *d, const double *px, const double *py, int target_cell_size, int max_depth) { /* :20-48 */
    bvh_t *b = &d->bvh;
    int T = d->T;
    if (!b->cx) {
        b->cx = dalloc(T); b->cy = dalloc(T); b->ids = (int *)malloc(sizeof(int) * (size_t)(T ? T : 1));
        b->cap_nodes = 2 * T + 2;
        b->nodes = (bvh_node *)malloc(sizeof(bvh_node) * (size_t)b->cap_nodes);
    }
    for (int i = 0; i < T; i++) {
        /* calculate_polygon_centroid, src/polygon_utils.cpp:238-263 */
        double vx[3], vy[3];
        for (int j = 0; j < 3; ++j) { vx[j] = px[d->tri[3 * i + j]]; vy[j] = py[d->tri[3 * i + j]]; }
        double c0 = 0.0, c1 = 0.0, signed_area = 0;
        for (int k = 0; k < 3; k++) {
            double x0 = vx[k], y0 = vy[k], x1 = vx[(k + 1) % 3], y1 = vy[(k + 1) % 3];
            double area = (x0 * y1) - (x1 * y0);
            signed_area += area;
            c0 += (x0 + x1) * area;
            c1 += (y0 + y1) * area;
        }
        signed_area *= 0.5;
        c0 /= 6 * signed_area;
        c1 /= 6 * signed_area;
        b->cx[i] = c0; b->cy[i] = c1; b->ids[i] = i;
    }
    memset(b->nodes, 0, sizeof(bvh_node));
    b->n_nodes = 1;
    bvh_build_node(b, d->tri, px, py, 0, 0, T, 0, target_cell_size, max_depth);
}

typedef struct { int face_id; double bc[3]; } hit_t;

/* src/bvh.cpp:162-190 */
static void barycentric(double t0x, double t0y, double t1x, double t1y, double t2x, double t2y, double px, double py,
                        double out[3]) {
    double v0x = t2x - t0x, v0y = t2y - t0y, v1x = t1x - t0x, v1y = t1y - t0y, v2x = px - t0x, v2y = py - t0y;
    double dot00 = v0x * v0x + v0y * v0y, dot01 = v0x * v1x + v0y * v1y, dot02 = v0x * v2x + v0y * v2y;
    double dot11 = v1x * v1x + v1y * v1y, dot12 = v1x * v2x + v1y * v2y;
    double denom = dot00 * dot11 - dot01 * dot01;
    double inv_denom = 1 / denom;
    double u = (dot11 * dot02 - dot01 * dot12) * inv_denom;
    double v = (dot00 * dot12 - dot01 * dot02) * inv_denom;
    out[0] = u; out[1] = v; out[2] = 1.0f - u - v;
}

static int inside_bbox(const bvh_node *n, double x, double y) {
    return n->min_x <= x && x <= n->max_x && n->min_y <= y && y <= n->max_y;
}

static void bvh_intersect(const bvh_t *b, const int *tri, const double *px, const double *py, int node, double x,
                          double y, hit_t *hit, int *found) {                       /* :196-255 */
    const bvh_node *n = b->nodes + node;
    if (n->is_leaf) {
        if (inside_bbox(n, x, y)) {
            for (int i = n->first_face; i < n->first_face + n->nb_faces; ++i) {
                const int *t = tri + 3 * b->ids[i];
                double eps = 1e-12, bc[3];
                barycentric(px[t[2]], py[t[2]], px[t[1]], py[t[1]], px[t[0]], py[t[0]], x, y, bc);
                if ((bc[0] >= -eps && bc[1] >= -eps) && ((bc[0] + bc[1]) <= 1.0f + eps)) {
                    hit->bc[0] = bc[0]; hit->bc[1] = bc[1]; hit->bc[2] = bc[2];
                    hit->face_id = b->ids[i];
                    *found = 1;
                    return;
                } else {
                    *found = 0;
                }
            }
        }
    } else {
        int c1 = n->first_child, c2 = n->first_child + 1;
        if (inside_bbox(b->nodes + c1, x, y)) {
            bvh_intersect(b, tri, px, py, c1, x, y, hit, found);
            if (*found) return;
        }
        if (inside_bbox(b->nodes + c2, x, y)) {
            bvh_intersect(b, tri, px, py, c2, x, y, hit, found);
            if (*found) return;
        }
    }
}

static int bvh_query(const pcdo_design *d, const double *px, const double *py, double x, double y, hit_t *hit) { /* :257-267 */
    int found = 0;
    if (inside_bbox(d->bvh.nodes, x, y)) bvh_intersect(&d->bvh, d->tri, px, py, 0, x, y, hit, &found);
    return found;
}

/* src/mesh.cpp:234-288 (target) and :291-345 (source): nodal raster of per-vertex values */
static int interpolate_raster(pcdo_design *d, const double *px, const double *py, const double *values, double *out) {
    bvh_build(d, px, py, 5, 30);
    double epsilon = 1e-8;
    int W = d->W, H = d->H, miss = 0;
    for (int i = 0; i < H; ++i) {
        double y = (double)i * (d->height - epsilon) / (H - 1) + 0.5 * epsilon;
        for (int j = 0; j < W; ++j) {
            double x = (double)j * (d->width - epsilon) / (W - 1) + 0.5 * epsilon;
            hit_t hit;
            if (bvh_query(d, px, py, x, y, &hit)) {
                const int *t = d->tri + 3 * hit.face_id;
                out[(size_t)i * W + j] = values[t[0]] * hit.bc[0] + values[t[1]] * hit.bc[1] + values[t[2]] * hit.bc[2];
            } else {
                out[(size_t)i * W + j] = NAN;
                miss = 1;
            }
        }
    }
    return miss;
}

/* ===================================================================================== */
/* Pipeline                                                                                */
/* ===================================================================================== */

void pcdo_initialize_solvers(pcdo_design *d, const double *image) {      /* caustic_design.cpp:334-364 */
    int nx = d->nx, ny = d->ny, V = d->V, T = d->T;
    size_t N = (size_t)d->W * d->H;
    d->pixels = dalloc(N);
    pcdo_scale_matrix_proportional(image, (long)N, 0, 1.0f, d->pixels);
    /* Mesh::generate_structured_mesh, src/mesh.cpp:45-64 */
    d->tx = dalloc(V); d->ty = dalloc(V); d->tz = dalloc(V);
    d->sx = dalloc(V); d->sy = dalloc(V); d->sz = dalloc(V);
    for (int i = 0; i < ny; ++i)
        for (int j = 0; j < nx; ++j) {
            d->tx[i * nx + j] = (double)j * d->width / (nx - 1);
            d->ty[i * nx + j] = (double)i * d->height / (ny - 1);
        }
    memcpy(d->sx, d->tx, sizeof(double) * V); memcpy(d->sy, d->ty, sizeof(double) * V);
    d->tri = (int *)malloc(sizeof(int) * 3 * (size_t)(T ? T : 1));
    int t = 0;
    for (int i = 0; i < ny - 1; ++i)
        for (int j = 0; j < nx - 1; ++j) {
            int idx = i * nx + j;
            d->tri[3 * t] = idx; d->tri[3 * t + 1] = idx + 1; d->tri[3 * t + 2] = idx + nx; ++t;
            d->tri[3 * t] = idx + nx; d->tri[3 * t + 1] = idx + 1; d->tri[3 * t + 2] = idx + nx + 1; ++t;
        }
    /* build_vertex_to_triangles :106-118 + find_adjacent_elements :121-147 */
    d->adj = (int *)malloc(sizeof(int) * 6 * (size_t)V);
    d->adj_n = (int *)calloc((size_t)V, sizeof(int));
    int *tmp = (int *)malloc(sizeof(int) * 6 * (size_t)V);
    for (int k = 0; k < T; ++k)
        for (int c = 0; c < 3; ++c) { int v = d->tri[3 * k + c]; tmp[6 * v + d->adj_n[v]++] = k; }
    for (int v = 0; v < V; ++v) d->adj_n[v] = unordered_set_order(tmp + 6 * v, d->adj_n[v], d->adj + 6 * v);
    free(tmp);
    /* get_target_partitioned_areas, src/polygon_utils.cpp:368-389 */
    d->target_areas = dalloc(V);
    double sum_target_area = 0.0f;
    for (int v = 0; v < V; ++v) {
        double total = 0.0f;
        for (int a = 0; a < d->adj_n[v]; ++a) {
            double qx[4], qy[4];
            triangle_quad(d, d->tx, d->ty, v, d->adj[6 * v + a], qx, qy);
            total += integrate_cell_intensities(d->pixels, qx, qy, d->W, d->H, d->width);
        }
        d->target_areas[v] = total;
        sum_target_area += total;
    }
    double scaling = (d->width * d->height) / sum_target_area;
    for (int v = 0; v < V; ++v) d->target_areas[v] *= scaling;
    d->phi = dalloc(N); d->h = dalloc(N);
    d->errors = dalloc(V); d->raster = dalloc(N); d->gx = dalloc(N); d->gy = dalloc(N);
    d->vgx = dalloc(V); d->vgy = dalloc(V); d->normals_x = dalloc(V); d->normals_y = dalloc(V);
    d->norm_x = dalloc(N); d->norm_y = dalloc(N); d->divergence = dalloc(N);
}

static int solve(pcdo_design *d, const double *rhs, double *x, double tol) {
    int sweeps;
    if (d->solver_mode == 0) sweeps = pcdo_poisson_lex(rhs, x, d->W, d->H, 100000, tol, NULL);
    else sweeps = pcdo_poisson_rb(rhs, x, d->W, d->H, 100000, tol, 0, NULL, NULL);
    d->last_sweeps = sweeps;
    return sweeps;
}

void pcdo_stage_errors(pcdo_design *d) {                                  /* caustic_design.cpp:194-209 */
    for (int v = 0; v < d->V; ++v) {
        double source_area = 0.0f, cell_area = 0.0f;                      /* polygon_utils.cpp:401-414, :216-236 */
        for (int a = 0; a < d->adj_n[v]; ++a) {
            double qx[4], qy[4];
            triangle_quad(d, d->tx, d->ty, v, d->adj[6 * v + a], qx, qy);
            double area = shoelace(qx, qy, 4);
            source_area += area;
            cell_area += area;
        }
        d->errors[v] = (d->target_areas[v] - source_area) / cell_area;
    }
}

int pcdo_stage_raster(pcdo_design *d) { return interpolate_raster(d, d->tx, d->ty, d->errors, d->raster); }

double pcdo_stage_step(pcdo_design *d) {                                  /* caustic_design.cpp:225-265 */
    int V = d->V, W = d->W, H = d->H, nx = d->nx, ny = d->ny;
    pcdo_gradient(d->phi, W, H, d->gx, d->gy);
    for (int i = 0; i < V; ++i) {
        double sxp = (d->tx[i] / d->width) * (W) - 0.5, syp = (d->ty[i] / d->height) * (H) - 0.5;
        d->vgx[i] = pcdo_bilinear(d->gx, W, H, sxp, syp);
        d->vgy[i] = pcdo_bilinear(d->gy, W, H, sxp, syp);
    }
    /* Mesh::step_grid, src/mesh.cpp:485-532; step_size = (double)0.05f, caustic_design.cpp:250 */
    double step_size = 0.05f, min_t = (d->width / nx), min_step = 0.0f;
    for (int i = 0; i < V; ++i) {
        int y = i / nx, x = i % nx;
        double vx = d->vgx[i], vy = d->vgy[i];
        int bx = (x == 0 || x == nx - 1), by = (y == 0 || y == ny - 1);
        if (bx) vx = 0;
        if (by) vy = 0;
        double ox = d->tx[i], oy = d->ty[i];
        d->tx[i] += vx * min_t * step_size;
        d->ty[i] += vy * min_t * step_size;
        double dx = ox - d->tx[i], dy = oy - d->ty[i], dz = 0.0;
        double dist = sqrt(dx * dx + dy * dy + dz * dz);
        if (min_step < dist) min_step = dist;
    }
    return min_step / d->width;
}

double pcdo_perform_transport_iteration(pcdo_design *d, int *miss) {     /* caustic_design.cpp:190-266 */
    pcdo_stage_errors(d);
    int m = pcdo_stage_raster(d);
    if (miss) *miss = m;
    if (m) return NAN;
    pcdo_subtract_average(d->raster, (long)d->W * d->H);
    solve(d, d->raster, d->phi, 0.0000001);
    return pcdo_stage_step(d);
}

long pcdo_inverted_transport_map(pcdo_design *d, double *out_x, double *out_y) {   /* src/mesh.cpp:348-409 */
    bvh_build(d, d->tx, d->ty, 5, 30);
    double epsilon = 1e-8, width = d->width, height = d->height;
    int nx = d->nx, ny = d->ny;
    long n = 0;
    for (int i = 0; i < d->V; ++i) {
        double qx = epsilon + d->sx[i] * ((width - 2 * epsilon) / width);
        double qy = epsilon + d->sy[i] * ((height - 2 * epsilon) / height);
        hit_t hit;
        if (!bvh_query(d, d->tx, d->ty, qx, qy, &hit)) continue;
        const int *t = d->tri + 3 * hit.face_id;
        double ix = d->sx[t[0]] * hit.bc[0] + d->sx[t[1]] * hit.bc[1] + d->sx[t[2]] * hit.bc[2];
        double iy = d->sy[t[0]] * hit.bc[0] + d->sy[t[1]] * hit.bc[1] + d->sy[t[2]] * hit.bc[2];
        int y = i / nx, x = i % nx;
        if (x == 0) ix = 0; else if (x == nx - 1) ix = width;
        if (y == 0) iy = 0; else if (y == ny - 1) iy = height;
        out_x[n] = ix; out_y[n] = iy; ++n;
    }
    return n;
}

int pcdo_perform_height_map_iteration(pcdo_design *d, int itr) {          /* caustic_design.cpp:269-332 */
    (void)itr;
    int V = d->V, W = d->W, H = d->H;
    size_t N = (size_t)W * H;
    /* Mesh::calculate_refractive_normals_uniform, src/mesh.cpp:677-722 */
    double focal_len = W / d->width * d->focal_l, refractive_index = 1.49;
    double *ivx = dalloc(V), *ivy = dalloc(V);
    long n = pcdo_inverted_transport_map(d, ivx, ivy);
    if (n != V) { free(ivx); free(ivy); return 1; }
    for (int i = 0; i < V; ++i) {
        double t[3] = {ivx[i] - d->sx[i], ivy[i] - d->sy[i], 0 - d->sz[i] + focal_len};
        double squared_len = 0;
        for (int k = 0; k < 3; ++k) squared_len += t[k] * t[k];                /* normalize, utils.cpp:370-384 */
        double len = sqrt(squared_len);
        for (int k = 0; k < 3; ++k) t[k] = t[k] / len;
        double inc[3] = {0.0f, 0.0f, 1.0f};
        double x_normal = t[0] - inc[0] * refractive_index;
        double y_normal = t[1] - inc[1] * refractive_index;
        double z_normal = t[2] - inc[2] * refractive_index;
        d->normals_x[i] = x_normal / z_normal;
        d->normals_y[i] = y_normal / z_normal;
    }
    free(ivx); free(ivy);
    if (interpolate_raster(d, d->sx, d->sy, d->normals_x, d->norm_x)) return 1;
    if (interpolate_raster(d, d->sx, d->sy, d->normals_y, d->norm_y)) return 1;
    pcdo_divergence(d->norm_x, d->norm_y, W, H, d->divergence);
    pcdo_subtract_average(d->divergence, (long)N);
    solve(d, d->divergence, d->h, 0.00000001);
    /* caustic_design.cpp:323-330 + Mesh::set_source_heights, src/mesh.cpp:724-742 */
    double *hv = dalloc(V);
    for (int i = 0; i < V; ++i)
        hv[i] = pcdo_bilinear(d->h, W, H, (d->sx[i] / d->width) * (W) - 0.5, (d->sy[i] / d->height) * (H) - 0.5);
    double max_h = 0;
    for (int i = 0; i < V; ++i)
        if (max_h > hv[i]) max_h = hv[i];
    for (int i = 0; i < V; ++i) d->sz[i] = hv[i] - max_h;
    free(hv);
    return 0;
}

int pcdo_last_sweeps(const pcdo_design *d) { return d->last_sweeps; }

static double *field_ptr(const pcdo_design *d, int field, long *n) {
    long N = (long)d->W * d->H, V = d->V;
    switch (field) {
        case 0: *n = N; return d->phi;
        case 1: *n = N; return d->h;
        case 2: *n = N; return d->raster;
        case 3: *n = N; return d->pixels;
        case 4: *n = N; return d->divergence;
        case 5: *n = N; return d->norm_x;
        case 6: *n = N; return d->norm_y;
        case 7: *n = N; return d->gx;
        case 8: *n = N; return d->gy;
        case 9: *n = V; return d->errors;
        case 10: *n = V; return d->target_areas;
        case 11: *n = V; return d->vgx;
        case 12: *n = V; return d->vgy;
        case 13: *n = V; return d->normals_x;
        case 14: *n = V; return d->normals_y;
        case 15: *n = V; return d->tx;
        case 16: *n = V; return d->ty;
        case 17: *n = V; return d->tz;
        case 18: *n = V; return d->sx;
        case 19: *n = V; return d->sy;
        case 20: *n = V; return d->sz;
        default: *n = -1; return NULL;
    }
}

long pcdo_get_field(const pcdo_design *d, int field, double *dst) {
    long n;
    double *p = field_ptr(d, field, &n);
    if (n > 0 && dst && p) memcpy(dst, p, sizeof(double) * (size_t)n);
    return n;
}

int pcdo_set_field(pcdo_design *d, int field, const double *src) {
    long n;
    double *p = field_ptr(d, field, &n);
    if (n <= 0 || !p) return -1;
    memcpy(p, src, sizeof(double) * (size_t)n);
    return 0;
}
