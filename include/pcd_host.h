/* pcd_host.h -- C entry points of libpcd_host.so: the host-only pieces around the hot path
 * (SURVEY 8f "next" rows f-1..f-3), exported so that tests can drive them without a GPU.
 *
 *   PNG ingest + grayscale + nearest resize   main.cpp:13-109,216-222   (zlib only; the reference uses libpng)
 *   solid OBJ / heightmap JSON / grid SVG      src/utils.cpp:176-307, main.cpp:111-135
 *   CLI flag parsing                           main.cpp:139-214
 *
 * The C++ drop-in API itself (poisson_solver, class Caustic_design) is declared in
 * poisson_caustic_design_b200/host/solver.h and host/caustic_design.h.
 */
#ifndef PCD_HOST_H
#define PCD_HOST_H

#ifdef __cplusplus
extern "C" {
#endif

/* Returns 0 on success; on failure returns nonzero and pcd_host_last_error() explains. */
const char *pcd_host_last_error(void);

/* main.cpp:29-109: decode a PNG and convert to gray = 0.299 r + 0.587 g + 0.114 b (r,g,b in [0,1], alpha
 * ignored).  *gray is malloc'ed [h*w], free with pcd_host_free. */
int pcd_host_load_png_gray(const char *path, int *width, int *height, double **gray);
void pcd_host_free(void *p);
/* main.cpp:13-27: nearest resampling, src = dst * old / new (integer division) */
void pcd_host_resize_nearest(const double *src, int old_w, int old_h, double *dst, int new_w, int new_h);

/* src/utils.cpp:198-260 (save_solid_obj) with find_perimeter_vertices :176-196; front = (fx,fy,fz), back plane
 * uses (bx,by); n = res_x*res_y points each */
int pcd_host_save_solid_obj(const double *fx, const double *fy, const double *fz, const double *bx, const double *by,
                            int res_x, int res_y, double width, double height, double thickness, const char *path);
/* main.cpp:111-135 */
int pcd_host_save_heightmap_json(const double *h, int res_x, int res_y, const char *path);
/* src/utils.cpp:262-307 (export_grid_to_svg) */
int pcd_host_export_grid_svg(const double *px, const double *py, int res_x, int res_y, double width, double height,
                             const char *path, double stroke_width);

typedef struct pcd_cli_options {   /* defaults: main.cpp:175-183 */
    char input_png[1024];
    char progress_out[1024];
    char output[1024];
    int has_progress_out;
    int res_w;
    double mesh_width, focal_l, thickness, conv_tres;   /* parsed as float, then widened (main.cpp:147-151) */
    int threads;
    int help;
    /* extensions (do not change any default behaviour) */
    int device;
    int solver_path;
    int quiet;
    int gpus;          /* --gpus N: Poisson solves as row slabs on devices device .. device+N-1 (default 1) */
} pcd_cli_options;
/* returns 0 ok, 1 parse error (message in pcd_host_last_error) */
int pcd_host_parse_cli(int argc, const char *const *argv, pcd_cli_options *out);

#ifdef __cplusplus
}
#endif
#endif
