#include "solver.h"

#include <cstdio>
#include <stdexcept>

#include "pcd.h"

double solver_progress = 0.0f;

static int g_device = 0;
void pcd_set_default_device(int device) { g_device = device; }
int pcd_get_default_device() { return g_device; }

void poisson_solver(std::vector<std::vector<double>> &D, std::vector<std::vector<double>> &phi, int width, int height,
                    int max_iterations, double convergence_threshold, int max_threads) {
    (void)max_threads;
    const size_t n = (size_t)width * height;
    std::vector<double> d(n), p(n);
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            d[(size_t)y * width + x] = D[y][x];
            p[(size_t)y * width + x] = phi[y][x];
        }
    pcd_solve_info info{};
    const int rc = pcd_poisson_solver(d.data(), p.data(), width, height, max_iterations, convergence_threshold, g_device, &info);
    if (rc != PCD_OK) throw std::runtime_error(std::string("poisson_solver: ") + pcd_last_error());
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) phi[y][x] = p[(size_t)y * width + x];
    // the reference's progress line (src/solver.cpp:136-145), once, with the value that ended the loop
    printf("\33[2K\r");
    printf("\tPoisson solver max_update: %.2e, convergence at %.2e\r", info.last_max_update, convergence_threshold);
    if (info.converged_at > 0) printf("\r\n");
}
