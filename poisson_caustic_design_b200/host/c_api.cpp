// extern "C" view of the host-only pieces (include/pcd_host.h).
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "host_internal.h"
#include "pcd_host.h"

namespace pcdh {
extern std::string g_host_err;
}

extern "C" {

const char *pcd_host_last_error(void) { return pcdh::g_host_err.c_str(); }

int pcd_host_load_png_gray(const char *path, int *width, int *height, double **gray) {
    std::vector<std::vector<double>> grid;
    if (!pcdh::image_to_grid(path, grid, pcdh::g_host_err)) return 1;
    const int h = (int)grid.size(), w = (int)grid[0].size();
    double *out = (double *)malloc(sizeof(double) * (size_t)w * h);
    for (int y = 0; y < h; ++y) memcpy(out + (size_t)y * w, grid[y].data(), sizeof(double) * w);
    *width = w; *height = h; *gray = out;
    return 0;
}

void pcd_host_free(void *p) { free(p); }

void pcd_host_resize_nearest(const double *src, int old_w, int old_h, double *dst, int new_w, int new_h) {
    for (int y = 0; y < new_h; ++y)
        for (int x = 0; x < new_w; ++x) {
            const int sx = x * old_w / new_w, sy = y * old_h / new_h;  // main.cpp:21-22
            dst[(size_t)y * new_w + x] = src[(size_t)sy * old_w + sx];
        }
}

int pcd_host_save_solid_obj(const double *fx, const double *fy, const double *fz, const double *bx, const double *by,
                            int res_x, int res_y, double width, double height, double thickness, const char *path) {
    return pcdh::save_solid_obj(fx, fy, fz, bx, by, res_x, res_y, width, height, thickness, path) ? 0 : 1;
}

int pcd_host_save_heightmap_json(const double *h, int res_x, int res_y, const char *path) {
    if (!pcdh::save_heightmap_json(h, res_x, res_y, path)) { pcdh::g_host_err = "Failed to open output file."; return 1; }
    return 0;
}

int pcd_host_export_grid_svg(const double *px, const double *py, int res_x, int res_y, double width, double height,
                             const char *path, double stroke_width) {
    return pcdh::export_grid_svg(px, py, res_x, res_y, width, height, path, stroke_width) ? 0 : 1;
}

}  // extern "C"
