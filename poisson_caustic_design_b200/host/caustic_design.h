// Drop-in for the reference's src/caustic_design.h + the part of src/mesh.h its callers touch: same class
// name, same public data members, same method signatures (including the reference's spelling
// `export_paramererization_to_svg`).  All per-iteration work runs device-resident behind the C ABI
// (include/pcd.h); the public members are host mirrors refreshed after each call according to
// `set_field_sync` (default SYNC_ALL = every member the reference would have updated).
#ifndef _CAUSTIC_DESIGN_H
#define _CAUSTIC_DESIGN_H

#include <string>
#include <vector>

#include "solver.h"

typedef std::vector<double> point_t;
typedef std::vector<point_t> polygon_t;

struct pcd_ctx;
struct pcd_multi;

// src/mesh.h:33-104, reduced to what Caustic_design's callers use: the two point sets and the exporters.
class Mesh {
   public:
    Mesh(double width, double height, int res_x, int res_y) : width(width), height(height), res_x(res_x), res_y(res_y) {}
    std::vector<point_t> source_points;
    std::vector<point_t> target_points;
    double width, height;
    int res_x, res_y;
    void export_paramererization_to_svg(std::string filename, double stroke_width);  // src/mesh.cpp:608-610
    void save_solid_obj_target(double thickness, const std::string &filename);       // :757-759
    void save_solid_obj_source(double thickness, const std::string &filename);       // :753-755
};

class Caustic_design {
   public:
    enum FieldSync {
        SYNC_ALL = 0,     // refresh every public member after each call (faithful drop-in, default)
        SYNC_VERTEX = 1,  // only per-vertex members (errors, vertex_gradient, normals, mesh points)
        SYNC_NONE = 2     // nothing; call sync_fields() when the members are needed
    };

    Caustic_design(/* args */);
    ~Caustic_design();

    Mesh *mesh;
    std::vector<std::vector<double>> phi;
    std::vector<double> errors;
    std::vector<std::vector<polygon_t>> target_cells;        // never materialised on the device path: stays empty
    std::vector<std::vector<polygon_t>> first_target_cells;  // idem
    std::vector<double> target_areas;
    std::vector<std::vector<double>> pixels;
    std::vector<std::vector<double>> raster;
    std::vector<std::vector<std::vector<double>>> gradient;
    std::vector<std::vector<double>> h;
    std::vector<std::vector<double>> divergence;
    std::vector<std::vector<double>> norm_x;
    std::vector<std::vector<double>> norm_y;
    std::vector<std::vector<double>> vertex_gradient;
    std::vector<std::vector<double>> normals;

    int mesh_res_x;
    int mesh_res_y;

    int resolution_x;
    int resolution_y;

    double width;
    double height;

    double focal_l;
    double thickness;
    int nthreads;

    double perform_transport_iteration();

    void perform_height_map_iteration(int itr);

    void initialize_solvers(std::vector<std::vector<double>> image);

    void set_mesh_resolution(int width, int heigth);
    void set_domain_resolution(int width, int heigth);
    void set_mesh_size(double width, double heigth);
    void set_lens_focal_length(double focal_length);
    void set_lens_thickness(double thickness);
    void set_solver_max_threads(int n_threads);

    void save_solid_obj_target(const std::string &filename);
    void save_solid_obj_source(const std::string &filename);

    void export_paramererization_to_svg(const std::string &filename, double line_width);

    void export_inverted_transport_map(std::string filename, double stroke_width);

    // ---- additions (none changes the behaviour of the members above) -------------------------------
    void set_field_sync(FieldSync mode) { field_sync = mode; }
    void set_device(int device) { this->device = device; }
    void set_solver_path(int path) { solver_path = path; }
    // Poisson solves spread over several GPUs of this process as row slabs; devices[0] also runs the other
    // stages.  Bit-identical to one GPU.  Call before initialize_solvers; an empty list = single GPU.
    void set_devices(const std::vector<int> &devices) { this->devices = devices; if (!devices.empty()) device = devices[0]; }
    void sync_fields();               // pull every public member from the device now
    void push_mesh();                 // upload what the reference's next call would read from the public members: both
                                      // point sets of the mesh (x, y, z) and the warm starts phi and h (SYNC_ALL does it itself)
    int last_solver_sweeps() const;   // sweeps of the most recent Poisson solve

   private:
    pcd_ctx *ctx;
    pcd_multi *multi;
    std::vector<int> devices;
    std::vector<double> pushed_sx, pushed_sy;   // source x / y as last uploaded (the device caches a map of the source mesh)
    FieldSync field_sync;
    int device;
    int solver_path;
    void pull(int level);
    void check(int rc, const char *what);
};

#endif  // _CAUSTIC_DESIGN_H
