// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// C-ABI harness over the *unmodified* reference sources.  oracle/Makefile compiles
// /root/reference/src/{caustic_design,polygon_utils,solver,utils,mesh,bvh}.cpp where they
// lie and links them with this file into oracle/_ref/libpcd_ref.so.  Nothing of the
// reference is copied: this file only calls the reference's public C++ interface
// (src/caustic_design.h:7-66, src/solver.h:8, src/utils.h:13-37, src/mesh.h:47-104) and
// flattens its vector<vector<double>> members into caller-provided double arrays so that
// Python (ctypes) can pin the restatement in oracle/pcd_oracle.c, generate the golden
// vectors under tests/golden/, and time the reference's own CPU solver (bench.py
// --impl reference / cpu_baseline).
//
// Setup mirrors main.cpp:224-237 (setters, then initialize_solvers) and the loops of
// main.cpp:243-262 are driven from Python one call at a time.

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>
#include <fcntl.h>

#include "caustic_design.h"   // -I/root/reference/src

namespace {

typedef std::vector<std::vector<double>> grid_t;

grid_t to_grid(const double *src, int w, int h) {
    grid_t g(h, std::vector<double>(w));
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) g[y][x] = src[(size_t)y * w + x];
    return g;
}

void from_grid(const grid_t &g, double *dst) {
    size_t k = 0;
    for (const auto &row : g)
        for (double v : row) dst[k++] = v;
}

// The reference prints progress with printf/std::cout from inside the hot loop; silence it
// by pointing fd 1 at /dev/null for the duration of a call when asked to.
struct Quiet {
    int saved = -1;
    explicit Quiet(bool on) {
        if (!on) return;
        fflush(stdout);
        std::cout.flush();
        saved = dup(1);
        int nul = open("/dev/null", O_WRONLY);
        dup2(nul, 1);
        close(nul);
    }
    ~Quiet() {
        if (saved < 0) return;
        fflush(stdout);
        std::cout.flush();
        dup2(saved, 1);
        close(saved);
    }
};

bool g_quiet = true;

}  // namespace

extern "C" {

void ref_set_quiet(int on) { g_quiet = on != 0; }

// ---- src/solver.h:8 ------------------------------------------------------------------
void ref_poisson_solver(const double *D, double *phi, int width, int height, int max_iterations,
                        double convergence_threshold, int max_threads) {
    Quiet q(g_quiet);
    grid_t in = to_grid(D, width, height);
    grid_t out = to_grid(phi, width, height);
    poisson_solver(in, out, width, height, max_iterations, convergence_threshold, max_threads);
    from_grid(out, phi);
}

// Same call but conversion to/from vector<vector> kept outside the timed region:
// returns seconds spent inside the reference's poisson_solver only.
double ref_poisson_solver_timed(const double *D, double *phi, int width, int height, int max_iterations,
                                double convergence_threshold, int max_threads) {
    Quiet q(g_quiet);
    grid_t in = to_grid(D, width, height);
    grid_t out = to_grid(phi, width, height);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    poisson_solver(in, out, width, height, max_iterations, convergence_threshold, max_threads);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    from_grid(out, phi);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

// ---- src/utils.h ---------------------------------------------------------------------
void ref_subtract_average(double *raster, int w, int h) {
    grid_t g = to_grid(raster, w, h);
    subtractAverage(g);
    from_grid(g, raster);
}

void ref_calculate_gradient(const double *grid, int w, int h, double *gx, double *gy) {
    grid_t g = to_grid(grid, w, h);
    auto grad = calculate_gradient(g);
    from_grid(grad[0], gx);
    from_grid(grad[1], gy);
}

void ref_calculate_divergence(const double *nx_, const double *ny_, int w, int h, double *out) {
    grid_t a = to_grid(nx_, w, h), b = to_grid(ny_, w, h);
    grid_t d = calculate_divergence(a, b, w, h);
    from_grid(d, out);
}

int ref_scale_matrix_proportional(const double *img, int w, int h, double lo, double hi, double *out) {
    try {
        grid_t g = to_grid(img, w, h);
        grid_t s = scale_matrix_proportional(g, lo, hi);
        from_grid(s, out);
        return 0;
    } catch (const std::exception &) {
        return 1;
    }
}

// ---- class Caustic_design (src/caustic_design.h:7-66) ----------------------------------
void *ref_cd_create(int mesh_res_x, int mesh_res_y, int res_x, int res_y, double width, double height,
                    double focal_l, double thickness, int nthreads) {
    Caustic_design *cd = new Caustic_design();
    cd->set_mesh_resolution(mesh_res_x, mesh_res_y);     // main.cpp:226
    cd->set_domain_resolution(res_x, res_y);             // main.cpp:227
    cd->set_mesh_size(width, height);                    // main.cpp:231
    cd->set_lens_focal_length(focal_l);                  // main.cpp:233
    cd->set_lens_thickness(thickness);                   // main.cpp:234
    cd->set_solver_max_threads(nthreads);                // main.cpp:235
    return cd;
}

void ref_cd_destroy(void *p) {
    Caustic_design *cd = (Caustic_design *)p;
    delete cd->mesh;  // the reference leaks it (src/caustic_design.cpp:16-18); the harness does not
    delete cd;
}

void ref_cd_initialize_solvers(void *p, const double *image) {
    Caustic_design *cd = (Caustic_design *)p;
    Quiet q(g_quiet);
    cd->initialize_solvers(to_grid(image, cd->resolution_x, cd->resolution_y));  // main.cpp:237
}

double ref_cd_perform_transport_iteration(void *p) {
    Quiet q(g_quiet);
    return ((Caustic_design *)p)->perform_transport_iteration();  // main.cpp:246
}

void ref_cd_perform_height_map_iteration(void *p, int itr) {
    Quiet q(g_quiet);
    ((Caustic_design *)p)->perform_height_map_iteration(itr);  // main.cpp:261
}

void ref_cd_save_solid_obj_source(void *p, const char *filename) {
    ((Caustic_design *)p)->save_solid_obj_source(filename);  // main.cpp:266
}

void ref_cd_export_parameterization_svg(void *p, const char *filename, double line_width) {
    ((Caustic_design *)p)->export_paramererization_to_svg(filename, line_width);  // main.cpp:240
}

void ref_cd_export_inverted_svg(void *p, const char *filename, double stroke_width) {
    Quiet q(g_quiet);
    ((Caustic_design *)p)->export_inverted_transport_map(filename, stroke_width);  // main.cpp:250
}

// Field ids shared with include/pcd.h (pcd_field).  Returns the number of doubles written
// (or that would be written when dst == NULL), -1 for an unknown id.
long ref_cd_get_field(void *p, int field, double *dst) {
    Caustic_design *cd = (Caustic_design *)p;
    auto grid = [&](const grid_t &g) -> long {
        long n = 0;
        for (const auto &r : g) n += (long)r.size();
        if (dst) from_grid(g, dst);
        return n;
    };
    auto vec = [&](const std::vector<double> &v) -> long {
        if (dst) std::memcpy(dst, v.data(), v.size() * sizeof(double));
        return (long)v.size();
    };
    auto pts = [&](const std::vector<point_t> &v, int c) -> long {
        if (dst)
            for (size_t i = 0; i < v.size(); ++i) dst[i] = v[i][c];
        return (long)v.size();
    };
    switch (field) {
        case 0: return grid(cd->phi);
        case 1: return grid(cd->h);
        case 2: return grid(cd->raster);
        case 3: return grid(cd->pixels);
        case 4: return grid(cd->divergence);
        case 5: return grid(cd->norm_x);
        case 6: return grid(cd->norm_y);
        case 7: return cd->gradient.size() == 2 ? grid(cd->gradient[0]) : 0;
        case 8: return cd->gradient.size() == 2 ? grid(cd->gradient[1]) : 0;
        case 9: return vec(cd->errors);
        case 10: return vec(cd->target_areas);
        case 11: return cd->vertex_gradient.size() == 2 ? vec(cd->vertex_gradient[0]) : 0;
        case 12: return cd->vertex_gradient.size() == 2 ? vec(cd->vertex_gradient[1]) : 0;
        case 13: return cd->normals.size() >= 2 ? vec(cd->normals[0]) : 0;
        case 14: return cd->normals.size() >= 2 ? vec(cd->normals[1]) : 0;
        case 15: return pts(cd->mesh->target_points, 0);
        case 16: return pts(cd->mesh->target_points, 1);
        case 17: return pts(cd->mesh->target_points, 2);
        case 18: return pts(cd->mesh->source_points, 0);
        case 19: return pts(cd->mesh->source_points, 1);
        case 20: return pts(cd->mesh->source_points, 2);
        default: return -1;
    }
}

// Writable state (all of it is public in the reference): lets a test start a reference
// stage from a chosen state.  Returns 0 on success.
int ref_cd_set_field(void *p, int field, const double *src) {
    Caustic_design *cd = (Caustic_design *)p;
    auto grid = [&](grid_t &g) {
        size_t k = 0;
        for (auto &r : g)
            for (double &v : r) v = src[k++];
        return 0;
    };
    auto pts = [&](std::vector<point_t> &v, int c) {
        for (size_t i = 0; i < v.size(); ++i) v[i][c] = src[i];
        return 0;
    };
    switch (field) {
        case 0: return grid(cd->phi);
        case 1: return grid(cd->h);
        case 10: std::memcpy(cd->target_areas.data(), src, cd->target_areas.size() * sizeof(double)); return 0;
        case 15: return pts(cd->mesh->target_points, 0);
        case 16: return pts(cd->mesh->target_points, 1);
        case 17: return pts(cd->mesh->target_points, 2);
        case 18: return pts(cd->mesh->source_points, 0);
        case 19: return pts(cd->mesh->source_points, 1);
        case 20: return pts(cd->mesh->source_points, 2);
        default: return -1;
    }
}

// Inverse transport map (src/mesh.cpp:348-409), x then y, V entries each.  Returns the
// number of points the reference produced (a miss silently shortens the array).
long ref_cd_inverted_transport_map(void *p, double *out_x, double *out_y) {
    Caustic_design *cd = (Caustic_design *)p;
    std::vector<point_t> inv = cd->mesh->calculate_inverted_transport_map();
    for (size_t i = 0; i < inv.size(); ++i) {
        out_x[i] = inv[i][0];
        out_y[i] = inv[i][1];
    }
    return (long)inv.size();
}

}  // extern "C"
