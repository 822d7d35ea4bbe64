// Drop-in for the reference's src/solver.h: same free function, same argument meaning.
//   D    read-only right-hand side [height][width]; NaN marks a hole (src/solver.cpp:29-44)
//   phi  in/out, warm start (src/solver.cpp:85-90), sized [height][width] by the caller
// Runs on the GPU through the C ABI (include/pcd.h: pcd_poisson_solver); `max_threads` is accepted for
// source compatibility and ignored.  Throws std::runtime_error when no CUDA device is usable -- there
// is no CPU fallback.  Like the reference it prints the last max_update line to stdout.
#ifndef SOLVER_H
#define SOLVER_H

#include <cmath>
#include <string>
#include <thread>
#include <vector>

void poisson_solver(std::vector<std::vector<double>> &D, std::vector<std::vector<double>> &phi, int width, int height,
                    int max_iterations, double convergence_threshold, int max_threads);

extern double solver_progress;  // src/solver.h:10 (never updated by the reference either)

// device used by poisson_solver() and new Caustic_design objects (default 0)
void pcd_set_default_device(int device);
int pcd_get_default_device();

#endif  // SOLVER_H
