// class Caustic_design on top of the C ABI (include/pcd.h).  Mirrors src/caustic_design.cpp call for call;
// the arithmetic lives in libpcd_b200.so.
#include "caustic_design.h"

#include <cstdio>
#include <iostream>
#include <stdexcept>

#include "host_internal.h"
#include "pcd.h"

namespace {

void to_grid(const std::vector<double> &flat, int w, int h, std::vector<std::vector<double>> &g) {
    g.assign(h, std::vector<double>(w));
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) g[y][x] = flat[(size_t)y * w + x];
}

void soa(const std::vector<point_t> &pts, std::vector<double> &x, std::vector<double> &y, std::vector<double> &z) {
    const size_t n = pts.size();
    x.resize(n); y.resize(n); z.resize(n);
    for (size_t i = 0; i < n; ++i) { x[i] = pts[i][0]; y[i] = pts[i][1]; z[i] = pts[i][2]; }
}

}  // namespace

// ---- Mesh exporters ---------------------------------------------------------------------------------
void Mesh::export_paramererization_to_svg(std::string filename, double stroke_width) {
    std::vector<double> x, y, z;
    soa(target_points, x, y, z);
    pcdh::export_grid_svg(x.data(), y.data(), res_x, res_y, width, height, filename, stroke_width);
}

void Mesh::save_solid_obj_source(double thickness, const std::string &filename) {
    std::vector<double> x, y, z;
    soa(source_points, x, y, z);
    pcdh::save_solid_obj(x.data(), y.data(), z.data(), x.data(), y.data(), res_x, res_y, width, height, thickness, filename);
}

void Mesh::save_solid_obj_target(double thickness, const std::string &filename) {
    std::vector<double> fx, fy, fz, bx, by, bz;
    soa(target_points, fx, fy, fz);
    soa(source_points, bx, by, bz);
    pcdh::save_solid_obj(fx.data(), fy.data(), fz.data(), bx.data(), by.data(), res_x, res_y, width, height, thickness, filename);
}

// ---- Caustic_design ------------------------------------------------------------------------------------
Caustic_design::Caustic_design(/* args */) {
    this->mesh = nullptr;
    this->mesh_res_x = 0;
    this->mesh_res_y = 0;
    this->resolution_x = 0;
    this->resolution_y = 0;
    this->width = 0.0f;
    this->height = 0.0f;
    this->focal_l = 0.0f;
    this->thickness = 0.0f;
    this->nthreads = 0;
    this->ctx = nullptr;
    this->multi = nullptr;
    this->field_sync = SYNC_ALL;
    this->device = pcd_get_default_device();
    this->solver_path = PCD_SOLVER_AUTO;
}

Caustic_design::~Caustic_design() {
    if (multi) pcd_multi_destroy(multi);   // before the context it is attached to
    if (ctx) pcd_destroy(ctx);
    delete mesh;  // the reference leaks its Mesh (src/caustic_design.cpp:16-18)
}

void Caustic_design::check(int rc, const char *what) {
    if (rc == PCD_OK) return;
    if (rc == PCD_ERR_RASTER_MISS) {
        // the reference prints and exit(0)s from inside the rasteriser (src/mesh.cpp:276-281)
        printf("interpolation miss!\r\n");
    }
    throw std::runtime_error(std::string(what) + ": " + pcd_last_error());
}

void Caustic_design::set_mesh_resolution(int width, int height) { this->mesh_res_x = width; this->mesh_res_y = height; }
void Caustic_design::set_domain_resolution(int width, int height) { this->resolution_x = width; this->resolution_y = height; }
void Caustic_design::set_mesh_size(double width, double height) { this->width = width; this->height = height; }
void Caustic_design::set_lens_focal_length(double focal_length) { this->focal_l = focal_length; }
void Caustic_design::set_lens_thickness(double thickness) { this->thickness = thickness; }
void Caustic_design::set_solver_max_threads(int n_threads) { this->nthreads = n_threads; }

void Caustic_design::save_solid_obj_target(const std::string &filename) { this->mesh->save_solid_obj_target(thickness, filename); }
void Caustic_design::save_solid_obj_source(const std::string &filename) { this->mesh->save_solid_obj_source(thickness, filename); }

void Caustic_design::export_paramererization_to_svg(const std::string &filename, double line_width) {
    mesh->export_paramererization_to_svg(filename, line_width);
}

void Caustic_design::export_inverted_transport_map(std::string filename, double stroke_width) {
    const size_t V = (size_t)mesh_res_x * mesh_res_y;
    std::vector<double> x(V), y(V);
    check(pcd_inverted_transport_map(ctx, x.data(), y.data()), "export_inverted_transport_map");
    pcdh::export_grid_svg(x.data(), y.data(), mesh_res_x, mesh_res_y, width, height, filename, stroke_width);
}

// level 0: per-vertex members; level 1: + grid members
void Caustic_design::pull(int level) {
    const size_t V = (size_t)mesh_res_x * mesh_res_y, N = (size_t)resolution_x * resolution_y;
    std::vector<double> a(V), b(V), c(V);
    auto get = [&](int f, std::vector<double> &dst) { check(pcd_get_field(ctx, f, dst.data()), "field download"); };
    get(PCD_FIELD_TARGET_X, a); get(PCD_FIELD_TARGET_Y, b); get(PCD_FIELD_TARGET_Z, c);
    for (size_t i = 0; i < V; ++i) { mesh->target_points[i][0] = a[i]; mesh->target_points[i][1] = b[i]; mesh->target_points[i][2] = c[i]; }
    get(PCD_FIELD_SOURCE_X, a); get(PCD_FIELD_SOURCE_Y, b); get(PCD_FIELD_SOURCE_Z, c);
    for (size_t i = 0; i < V; ++i) { mesh->source_points[i][0] = a[i]; mesh->source_points[i][1] = b[i]; mesh->source_points[i][2] = c[i]; }
    errors.resize(V); get(PCD_FIELD_ERRORS, errors);
    target_areas.resize(V); get(PCD_FIELD_TARGET_AREAS, target_areas);
    vertex_gradient.assign(2, std::vector<double>(V));
    get(PCD_FIELD_VERTEX_GRADIENT_X, vertex_gradient[0]); get(PCD_FIELD_VERTEX_GRADIENT_Y, vertex_gradient[1]);
    normals.assign(3, std::vector<double>(V, 1.0));  // z component is n_z/n_z = 1 (src/mesh.cpp:718)
    get(PCD_FIELD_NORMALS_X, normals[0]); get(PCD_FIELD_NORMALS_Y, normals[1]);
    if (level < 1) return;
    std::vector<double> g(N);
    auto grid = [&](int f, std::vector<std::vector<double>> &dst) { get(f, g); to_grid(g, resolution_x, resolution_y, dst); };
    grid(PCD_FIELD_PHI, phi); grid(PCD_FIELD_H, h); grid(PCD_FIELD_RASTER, raster); grid(PCD_FIELD_PIXELS, pixels);
    grid(PCD_FIELD_DIVERGENCE, divergence); grid(PCD_FIELD_NORM_X, norm_x); grid(PCD_FIELD_NORM_Y, norm_y);
    gradient.assign(2, std::vector<std::vector<double>>());
    grid(PCD_FIELD_GRADIENT_X, gradient[0]); grid(PCD_FIELD_GRADIENT_Y, gradient[1]);
}

void Caustic_design::sync_fields() { pull(1); }

// Everything of the public state the reference's next call would read: both point sets of the mesh (all three
// coordinates) and the warm starts phi and h.  SYNC_ALL calls it before every iteration, so that a caller who edits the
// public members between calls sees the reference's behaviour; the other sync modes leave the device state alone.
void Caustic_design::push_mesh() {
    std::vector<double> x, y, z;
    soa(mesh->target_points, x, y, z);
    check(pcd_set_field(ctx, PCD_FIELD_TARGET_X, x.data()), "mesh upload");
    check(pcd_set_field(ctx, PCD_FIELD_TARGET_Y, y.data()), "mesh upload");
    check(pcd_set_field(ctx, PCD_FIELD_TARGET_Z, z.data()), "mesh upload");
    soa(mesh->source_points, x, y, z);
    if (x != pushed_sx || y != pushed_sy) {   // the device caches a point-location map of the source mesh: only invalidate it on a real edit
        check(pcd_set_field(ctx, PCD_FIELD_SOURCE_X, x.data()), "mesh upload");
        check(pcd_set_field(ctx, PCD_FIELD_SOURCE_Y, y.data()), "mesh upload");
        pushed_sx = x; pushed_sy = y;
    }
    check(pcd_set_field(ctx, PCD_FIELD_SOURCE_Z, z.data()), "mesh upload");
    auto grid_up = [&](const std::vector<std::vector<double>> &g, int field) {
        if ((int)g.size() != resolution_y || g.empty() || (int)g[0].size() != resolution_x) return;   // not mirrored yet
        std::vector<double> flat((size_t)resolution_x * resolution_y);
        for (int yy = 0; yy < resolution_y; ++yy)
            for (int xx = 0; xx < resolution_x; ++xx) flat[(size_t)yy * resolution_x + xx] = g[yy][xx];
        check(pcd_set_field(ctx, field, flat.data()), "field upload");
    };
    grid_up(phi, PCD_FIELD_PHI);
    grid_up(h, PCD_FIELD_H);
}

int Caustic_design::last_solver_sweeps() const {
    pcd_solve_info info{};
    if (!ctx || pcd_last_solve_info(ctx, &info) != PCD_OK) return 0;
    return info.sweeps;
}

double Caustic_design::perform_transport_iteration() {
    if (!ctx) throw std::runtime_error("perform_transport_iteration: initialize_solvers has not been called");
    if (field_sync == SYNC_ALL) push_mesh();  // the mesh is a public member: honour edits made by the caller
    double step = 0.0;
    check(pcd_perform_transport_iteration(ctx, &step), "perform_transport_iteration");
    pcd_solve_info info{};
    pcd_last_solve_info(ctx, &info);
    printf("\33[2K\r");
    printf("\tPoisson solver max_update: %.2e, convergence at %.2e\r", info.last_max_update, 0.0000001);
    if (info.converged_at > 0) printf("\r\n");
    if (field_sync == SYNC_ALL) pull(1);
    else if (field_sync == SYNC_VERTEX) pull(0);
    return step;
}

void Caustic_design::perform_height_map_iteration(int itr) {
    if (!ctx) throw std::runtime_error("perform_height_map_iteration: initialize_solvers has not been called");
    if (field_sync == SYNC_ALL) push_mesh();
    double max_update = 0.0;
    check(pcd_perform_height_map_iteration(ctx, itr, &max_update), "perform_height_map_iteration");
    pcd_solve_info info{};
    pcd_last_solve_info(ctx, &info);
    printf("\33[2K\r");
    printf("\tPoisson solver max_update: %.2e, convergence at %.2e\r", info.last_max_update, 0.00000001);
    if (info.converged_at > 0) printf("\r\n");
    printf("height max update %.5e\r\n", max_update);  // src/caustic_design.cpp:331
    if (field_sync == SYNC_ALL) pull(1);
    else if (field_sync == SYNC_VERTEX) pull(0);
}

void Caustic_design::initialize_solvers(std::vector<std::vector<double>> image) {
    // scale_matrix_proportional's input checks (src/utils.cpp:91-101)
    if (image.empty()) throw std::invalid_argument("Input matrix is empty.");
    for (const auto &row : image)
        if (row.size() != image[0].size()) throw std::invalid_argument("Input matrix has inconsistent row sizes.");
    if ((int)image.size() != resolution_y || (int)image[0].size() != resolution_x)
        throw std::invalid_argument("image size does not match set_domain_resolution");
    if (multi) { pcd_multi_destroy(multi); multi = nullptr; }
    if (ctx) { pcd_destroy(ctx); ctx = nullptr; }
    pcd_config cfg{};
    cfg.mesh_res_x = mesh_res_x; cfg.mesh_res_y = mesh_res_y;
    cfg.res_x = resolution_x; cfg.res_y = resolution_y;
    cfg.width = width; cfg.height = height;
    cfg.focal_l = focal_l; cfg.thickness = thickness;
    cfg.device = device; cfg.solver_path = solver_path;
    check(pcd_create(&cfg, &ctx), "initialize_solvers");
    if (devices.size() > 1) {
        check(pcd_multi_create(resolution_x, resolution_y, devices.data(), (int)devices.size(), &multi), "initialize_solvers (multi-GPU)");
        check(pcd_multi_attach(multi, ctx), "initialize_solvers (multi-GPU)");
    }
    std::vector<double> flat((size_t)resolution_x * resolution_y);
    for (int y = 0; y < resolution_y; ++y)
        for (int x = 0; x < resolution_x; ++x) flat[(size_t)y * resolution_x + x] = image[y][x];
    delete mesh;
    mesh = new Mesh(width, height, mesh_res_x, mesh_res_y);
    printf("%i, %i, %f, %f\r\n", mesh_res_x, mesh_res_y, width, height);  // src/mesh.cpp:46
    const size_t V = (size_t)mesh_res_x * mesh_res_y;
    mesh->target_points.assign(V, point_t(3, 0.0));
    mesh->source_points.assign(V, point_t(3, 0.0));
    check(pcd_initialize_solvers(ctx, flat.data()), "initialize_solvers");
    std::cout << "built mesh" << std::endl;  // src/caustic_design.cpp:341
    pull(field_sync == SYNC_ALL ? 1 : 0);
    {
        std::vector<double> z;
        soa(mesh->source_points, pushed_sx, pushed_sy, z);
    }
    std::cout << target_areas.size() << std::endl;  // :350
}
