// CLI drop-in for the reference's main.cpp:137-272: same flags, defaults, stdout lines and output files
// (<output>output.obj, <output>heightmap.json, optional <progress_out>parameterization_<k>.svg / inverted.svg);
// the transport and height loops run on the GPU through class Caustic_design (host/caustic_design.h).
// Extra flags (--device N, --gpus N, --solver_path auto|streaming|resident|tiled|dct, --quiet) do not change any default.
#include <cmath>
#include <cstdio>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "caustic_design.h"
#include "host_internal.h"
#include "pcd_host.h"

static void usage() {
    std::cout << "  caustic_design {OPTIONS}\n\n  OPTIONS:\n\n"
                 "      -h, --help                        Display this help menu\n"
                 "      --input_png=[image]               Image input\n"
                 "      --progress_out=[progress]         Transport progress output in SVG\n"
                 "      --output=[output]                 3D OBJ file output\n"
                 "      --res_w=[resolution]              Mesh resolution\n"
                 "      --mesh_width=[width]              Lens width\n"
                 "      --focal_l=[focal_length]          Focal length\n"
                 "      --thickness=[thickness]           Lens Thickness\n"
                 "      --threads=[max_threads]           Number of CPU threads to use (ignored: GPU path)\n"
                 "      --conv_tres=[convergence]         Contrast convergence treshold\n"
                 "      --device=[n]                      CUDA device ordinal (default 0)\n"
                 "      --gpus=[n]                        Poisson solves as row slabs on n GPUs (devices device..device+n-1)\n"
                 "      --solver_path=[auto|streaming|resident|tiled|dct]\n";
}

int main(int argc, char const *argv[]) {
    pcd_cli_options opt;
    if (pcd_host_parse_cli(argc, argv, &opt) != 0) {
        std::cerr << pcd_host_last_error() << std::endl;
        usage();
        return 1;
    }
    if (opt.help) {
        usage();
        return 0;
    }
    const std::string image_filename = opt.input_png, progress_path = opt.progress_out, output_path = opt.output;
    const bool output_progress = opt.has_progress_out != 0;
    const int mesh_resolution = opt.res_w;
    const double lens_width = opt.mesh_width, lens_focal_l = opt.focal_l, lens_thickness = opt.thickness, convergence = opt.conv_tres;

    // Load image to grid (main.cpp:216-222)
    std::vector<std::vector<double>> pixels;
    std::string err;
    if (!pcdh::image_to_grid(image_filename, pixels, err)) throw std::runtime_error(err);
    double aspect_ratio = (double)pixels[0].size() / (double)pixels.size();

    std::vector<std::vector<double>> resized_pixels;
    pcdh::resize_image(pixels, resized_pixels, 4 * mesh_resolution, 4 * mesh_resolution / aspect_ratio);

    pcd_set_default_device(opt.device);
    Caustic_design caustic_design;
    caustic_design.set_solver_path(opt.solver_path);
    if (opt.gpus > 1) {   // one process, several GPUs: row slabs with the ghost-row exchange inside the pass kernel
        std::vector<int> devices;
        for (int g = 0; g < opt.gpus; ++g) devices.push_back(opt.device + g);
        caustic_design.set_devices(devices);
    }
    // the CLI only needs the mesh between iterations (SVG progress) and h at the end
    caustic_design.set_field_sync(Caustic_design::SYNC_VERTEX);

    caustic_design.set_mesh_resolution(mesh_resolution, mesh_resolution / aspect_ratio);
    caustic_design.set_domain_resolution(4 * mesh_resolution, 4 * mesh_resolution / aspect_ratio);

    double mesh_height = floor((mesh_resolution) / aspect_ratio) * (lens_width / (mesh_resolution));

    caustic_design.set_mesh_size(lens_width, mesh_height);

    caustic_design.set_lens_focal_length(lens_focal_l);
    caustic_design.set_lens_thickness(lens_thickness);
    caustic_design.set_solver_max_threads(opt.threads);

    caustic_design.initialize_solvers(resized_pixels);

    if (output_progress) {
        caustic_design.export_paramererization_to_svg(progress_path + "parameterization_0.svg", 0.5f);
    }

    for (int itr = 0; itr < 50; itr++) {
        printf("starting iteration %i\r\n", itr);

        double step_size = caustic_design.perform_transport_iteration();

        if (output_progress) {
            caustic_design.export_paramererization_to_svg(progress_path + "parameterization_" + std::to_string(itr + 1) + ".svg", 1.0f);
            caustic_design.export_inverted_transport_map(progress_path + "inverted.svg", 1.0f);
        }

        printf("\tTransport step size = %f, convergence at %f\r\n", step_size, convergence);

        if (step_size < convergence) break;
    }

    printf("\033[0;32mTransport map solver done! Starting height solver.\033[0m\r\n");

    for (int itr = 0; itr < 3; itr++) {
        caustic_design.perform_height_map_iteration(itr);
    }

    printf("Height solver done! Exporting as solidified obj\r\n");

    caustic_design.save_solid_obj_source(output_path + "output.obj");

    // Save the heightmap to JSON
    caustic_design.sync_fields();
    const int W = caustic_design.resolution_x, H = caustic_design.resolution_y;
    std::vector<double> flat((size_t)W * H);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) flat[(size_t)y * W + x] = caustic_design.h[y][x];
    if (!pcdh::save_heightmap_json(flat.data(), W, H, output_path + "heightmap.json"))
        throw std::runtime_error("Failed to open output file.");

    return 0;
}
