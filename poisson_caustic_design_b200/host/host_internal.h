// Internal declarations of libpcd_host.so.
#pragma once

#include <string>
#include <vector>

namespace pcdh {

bool load_png_rgb(const std::string &path, int &width, int &height, std::vector<unsigned char> &rgb, std::string &err);
bool image_to_grid(const std::string &path, std::vector<std::vector<double>> &image_grid, std::string &err);
void resize_image(const std::vector<std::vector<double>> &input_image, std::vector<std::vector<double>> &output_image,
                  int new_width, int new_height);

bool save_solid_obj(const double *fx, const double *fy, const double *fz, const double *bx, const double *by, int res_x,
                    int res_y, double width, double height, double thickness, const std::string &filename);
bool save_heightmap_json(const double *h, int res_x, int res_y, const std::string &filename);
bool export_grid_svg(const double *px, const double *py, int res_x, int res_y, double width, double height,
                     const std::string &filename, double stroke_width);

}  // namespace pcdh
