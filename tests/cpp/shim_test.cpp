// Exercises the C++ drop-in API exactly the way reference callers use it (src/solver.h:8,
// src/caustic_design.h:7-66): reads a raw problem from argv[1], writes results to argv[2].
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <vector>

#include "caustic_design.h"
#include "solver.h"

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    std::ifstream in(argv[1], std::ios::binary);
    int W, H, nx, ny;
    in.read((char *)&W, 4); in.read((char *)&H, 4); in.read((char *)&nx, 4); in.read((char *)&ny, 4);
    std::vector<std::vector<double>> D(H, std::vector<double>(W)), phi(H, std::vector<double>(W, 0.0)), img(H, std::vector<double>(W));
    for (auto &r : D) in.read((char *)r.data(), 8 * W);
    for (auto &r : img) in.read((char *)r.data(), 8 * W);
    std::ofstream out(argv[2], std::ios::binary);
    // 1. free function
    poisson_solver(D, phi, W, H, 100000, 1e-7, 4);
    for (auto &r : phi) out.write((char *)r.data(), 8 * W);
    // 2. class, public members read back like main.cpp does
    Caustic_design cd;
    cd.set_mesh_resolution(nx, ny);
    cd.set_domain_resolution(W, H);
    cd.set_mesh_size(0.5, 0.5 * ny / nx);
    cd.set_lens_focal_length(1.5);
    cd.set_lens_thickness(0.1);
    cd.set_solver_max_threads(1);
    cd.initialize_solvers(img);
    double steps[2];
    for (int i = 0; i < 2; ++i) steps[i] = cd.perform_transport_iteration();
    out.write((char *)steps, 16);
    for (auto &p : cd.mesh->target_points) out.write((char *)p.data(), 24);
    for (auto &r : cd.phi) out.write((char *)r.data(), 8 * W);
    out.write((char *)cd.errors.data(), 8 * cd.errors.size());
    // the public members belong to the caller (src/caustic_design.h:16-30): put the mesh back on the source lattice and
    // zero the warm start -- the next iteration must then repeat iteration 0
    cd.mesh->target_points = cd.mesh->source_points;
    for (auto &r : cd.phi) std::fill(r.begin(), r.end(), 0.0);
    const double again = cd.perform_transport_iteration();
    out.write((char *)&again, 8);
    cd.perform_height_map_iteration(0);
    for (auto &p : cd.mesh->source_points) out.write((char *)p.data(), 24);
    for (auto &r : cd.h) out.write((char *)r.data(), 8 * W);
    // error behaviour: wrong image size -> std::invalid_argument (src/utils.cpp:91-101 family)
    int caught = 0;
    try { Caustic_design bad; bad.set_domain_resolution(8, 8); bad.set_mesh_resolution(4, 4); bad.set_mesh_size(1, 1); bad.initialize_solvers({}); }
    catch (const std::invalid_argument &) { caught = 1; }
    out.write((char *)&caught, 4);
    return 0;
}
