"""TEST INFRASTRUCTURE ONLY: ctypes loaders for the two CPU checkers.

* ``RefLib``    -- oracle/_ref/libpcd_ref.so: the UNMODIFIED reference sources
  (/root/reference/src/*.cpp) behind oracle/ref_harness.cpp.
* ``OracleLib`` -- oracle/libpcd_oracle.so: the plain-C restatement (oracle/pcd_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.  The product package (poisson_caustic_design_b200) never does.

Also holds the host-side restatement of main.cpp's input preparation (grayscale, nearest resize,
derived sizes) used to feed both checkers and the CUDA path with identical inputs.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libpcd_ref.so")
PORT_SO = os.path.join(HERE, "libpcd_oracle.so")
REFERENCE_ROOT = "/root/reference"

# field ids: keep in sync with include/pcd.h (enum pcd_field) and oracle/ref_harness.cpp
FIELDS = {
    "phi": 0, "h": 1, "raster": 2, "pixels": 3, "divergence": 4, "norm_x": 5, "norm_y": 6,
    "gradient_x": 7, "gradient_y": 8, "errors": 9, "target_areas": 10,
    "vertex_gradient_x": 11, "vertex_gradient_y": 12, "normals_x": 13, "normals_y": 14,
    "target_x": 15, "target_y": 16, "target_z": 17, "source_x": 18, "source_y": 19, "source_z": 20,
}
GRID_FIELDS = {"phi", "h", "raster", "pixels", "divergence", "norm_x", "norm_y", "gradient_x", "gradient_y"}

_dp = C.POINTER(C.c_double)


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def build(ref: bool | None = None, quiet: bool = True) -> None:
    """Compile the checkers (make -C oracle port [ref]).  ``ref`` defaults to "if the reference is here"."""
    targets = ["port"]
    if ref is None:
        ref = os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))
    if ref:
        targets.append("ref")
    subprocess.run(["make", "-C", HERE] + targets, check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


# ------------------------------------------------------------------------------------------------
# main.cpp input preparation, restated (main.cpp:13-27, 95-98, 216-231)
# ------------------------------------------------------------------------------------------------
def rgba_to_gray(rgba: np.ndarray) -> np.ndarray:
    """main.cpp:95-98: gray = 0.299*(r/255) + 0.587*(g/255) + 0.114*(b/255), alpha ignored."""
    r = rgba[..., 0].astype(np.float64) / 255.0
    g = rgba[..., 1].astype(np.float64) / 255.0
    b = rgba[..., 2].astype(np.float64) / 255.0
    return (0.299 * r) + (0.587 * g) + (0.114 * b)


def resize_nearest(img: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    """main.cpp:13-27: src = dst * old / new with integer division."""
    old_h, old_w = img.shape
    sx = (np.arange(new_w, dtype=np.int64) * old_w) // new_w
    sy = (np.arange(new_h, dtype=np.int64) * old_h) // new_h
    return np.ascontiguousarray(img[sy][:, sx])


@dataclass
class Setup:
    """Derived sizes of main.cpp:216-231 for an image of img_w x img_h pixels."""
    mesh_nx: int
    mesh_ny: int
    res_x: int
    res_y: int
    width: float
    height: float
    focal_l: float
    thickness: float


def f32(x: float) -> float:
    """CLI numeric flags are args::ValueFlag<float> widened to double (main.cpp:147-151)."""
    return float(np.float32(x))


def derive_setup(img_w: int, img_h: int, res_w: int, mesh_width: float, focal_l: float,
                 thickness: float) -> Setup:
    aspect = float(img_w) / float(img_h)                      # main.cpp:219
    res_x = 4 * res_w
    res_y = int(4 * res_w / aspect)                           # main.cpp:222,227 (double -> int)
    mesh_ny = int(res_w / aspect)                             # main.cpp:226
    mesh_h = math.floor(res_w / aspect) * (mesh_width / res_w)  # main.cpp:229
    return Setup(res_w, mesh_ny, res_x, res_y, mesh_width, mesh_h, focal_l, thickness)


def prepare_image(gray: np.ndarray, res_w: int, mesh_width: float = 1.0, focal_l: float = 1.5,
                  thickness: float = 0.2):
    """Returns (Setup, resized image) exactly as main.cpp hands them to initialize_solvers."""
    img_h, img_w = gray.shape
    s = derive_setup(img_w, img_h, res_w, mesh_width, focal_l, thickness)
    return s, resize_nearest(gray, s.res_x, s.res_y)


def load_png_gray(path: str) -> np.ndarray:
    from PIL import Image
    return rgba_to_gray(np.asarray(Image.open(path).convert("RGBA")))


# ------------------------------------------------------------------------------------------------
# oracle/_ref: the reference itself
# ------------------------------------------------------------------------------------------------
class RefLib:
    def __init__(self, path: str = REF_SO):
        self.lib = L = C.CDLL(path)
        L.ref_poisson_solver.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
        L.ref_poisson_solver.restype = None
        L.ref_poisson_solver_timed.argtypes = L.ref_poisson_solver.argtypes
        L.ref_poisson_solver_timed.restype = C.c_double
        L.ref_subtract_average.argtypes = [_dp, C.c_int, C.c_int]
        L.ref_calculate_gradient.argtypes = [_dp, C.c_int, C.c_int, _dp, _dp]
        L.ref_calculate_divergence.argtypes = [_dp, _dp, C.c_int, C.c_int, _dp]
        L.ref_scale_matrix_proportional.argtypes = [_dp, C.c_int, C.c_int, C.c_double, C.c_double, _dp]
        L.ref_scale_matrix_proportional.restype = C.c_int
        L.ref_cd_create.argtypes = [C.c_int] * 4 + [C.c_double] * 4 + [C.c_int]
        L.ref_cd_create.restype = C.c_void_p
        L.ref_cd_destroy.argtypes = [C.c_void_p]
        L.ref_cd_initialize_solvers.argtypes = [C.c_void_p, _dp]
        L.ref_cd_perform_transport_iteration.argtypes = [C.c_void_p]
        L.ref_cd_perform_transport_iteration.restype = C.c_double
        L.ref_cd_perform_height_map_iteration.argtypes = [C.c_void_p, C.c_int]
        L.ref_cd_save_solid_obj_source.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_cd_export_parameterization_svg.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.ref_cd_export_inverted_svg.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.ref_cd_get_field.argtypes = [C.c_void_p, C.c_int, _dp]
        L.ref_cd_get_field.restype = C.c_long
        L.ref_cd_set_field.argtypes = [C.c_void_p, C.c_int, _dp]
        L.ref_cd_set_field.restype = C.c_int
        L.ref_cd_inverted_transport_map.argtypes = [C.c_void_p, _dp, _dp]
        L.ref_cd_inverted_transport_map.restype = C.c_long
        L.ref_set_quiet.argtypes = [C.c_int]

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def poisson_solver(self, D, phi, max_iterations, tol, threads=1, timed=False):
        D = np.ascontiguousarray(D, dtype=np.float64)
        phi = np.array(phi, dtype=np.float64, order="C")
        h, w = D.shape
        fn = self.lib.ref_poisson_solver_timed if timed else self.lib.ref_poisson_solver
        t = fn(_p(D), _p(phi), w, h, int(max_iterations), float(tol), int(threads))
        return (phi, t) if timed else phi

    def subtract_average(self, raster):
        r = np.array(raster, dtype=np.float64, order="C")
        self.lib.ref_subtract_average(_p(r), r.shape[1], r.shape[0])
        return r

    def gradient(self, grid):
        g = np.ascontiguousarray(grid, dtype=np.float64)
        gx, gy = np.empty_like(g), np.empty_like(g)
        self.lib.ref_calculate_gradient(_p(g), g.shape[1], g.shape[0], _p(gx), _p(gy))
        return gx, gy

    def divergence(self, nx, ny):
        nx = np.ascontiguousarray(nx, dtype=np.float64)
        ny = np.ascontiguousarray(ny, dtype=np.float64)
        out = np.empty_like(nx)
        self.lib.ref_calculate_divergence(_p(nx), _p(ny), nx.shape[1], nx.shape[0], _p(out))
        return out

    def design(self, setup: Setup, threads: int = 1) -> "RefDesign":
        return RefDesign(self, setup, threads)


class RefDesign:
    """One reference Caustic_design instance (src/caustic_design.h:7-66)."""

    def __init__(self, ref: RefLib, s: Setup, threads: int = 1):
        self.ref, self.s = ref, s
        self.h = ref.lib.ref_cd_create(s.mesh_nx, s.mesh_ny, s.res_x, s.res_y, s.width, s.height,
                                       s.focal_l, s.thickness, threads)

    def close(self):
        if self.h:
            self.ref.lib.ref_cd_destroy(self.h)
            self.h = None

    def initialize_solvers(self, image):
        img = np.ascontiguousarray(image, dtype=np.float64)
        assert img.shape == (self.s.res_y, self.s.res_x)
        self.ref.lib.ref_cd_initialize_solvers(self.h, _p(img))

    def transport_iteration(self) -> float:
        return self.ref.lib.ref_cd_perform_transport_iteration(self.h)

    def height_iteration(self, itr: int):
        self.ref.lib.ref_cd_perform_height_map_iteration(self.h, itr)

    def get(self, name: str) -> np.ndarray:
        fid = FIELDS[name]
        n = self.ref.lib.ref_cd_get_field(self.h, fid, None)
        out = np.empty(max(n, 0), dtype=np.float64)
        if n > 0:
            self.ref.lib.ref_cd_get_field(self.h, fid, _p(out))
        if name in GRID_FIELDS and n == self.s.res_x * self.s.res_y:
            out = out.reshape(self.s.res_y, self.s.res_x)
        return out

    def set(self, name: str, value):
        v = np.ascontiguousarray(value, dtype=np.float64).ravel()
        rc = self.ref.lib.ref_cd_set_field(self.h, FIELDS[name], _p(v))
        assert rc == 0, name

    def inverted_transport_map(self):
        V = self.s.mesh_nx * self.s.mesh_ny
        x, y = np.empty(V), np.empty(V)
        n = self.ref.lib.ref_cd_inverted_transport_map(self.h, _p(x), _p(y))
        return x[:n], y[:n]

    def save_obj(self, path: str):
        self.ref.lib.ref_cd_save_solid_obj_source(self.h, path.encode())


# ------------------------------------------------------------------------------------------------
# oracle/libpcd_oracle.so: the plain-C restatement
# ------------------------------------------------------------------------------------------------
class OracleLib:
    def __init__(self, path: str = PORT_SO):
        if not os.path.exists(path):
            build(ref=False)
        self.lib = L = C.CDLL(path)
        ip = C.POINTER(C.c_int)
        L.pcdo_poisson_lex.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double, _dp]
        L.pcdo_poisson_lex.restype = C.c_int
        L.pcdo_poisson_rb.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, ip, _dp]
        L.pcdo_poisson_rb.restype = C.c_int
        L.pcdo_subtract_average.argtypes = [_dp, C.c_long]
        L.pcdo_gradient.argtypes = [_dp, C.c_int, C.c_int, _dp, _dp]
        L.pcdo_divergence.argtypes = [_dp, _dp, C.c_int, C.c_int, _dp]
        L.pcdo_scale_matrix_proportional.argtypes = [_dp, C.c_long, C.c_double, C.c_double, _dp]
        L.pcdo_bilinear.argtypes = [_dp, C.c_int, C.c_int, C.c_double, C.c_double]
        L.pcdo_bilinear.restype = C.c_double
        L.pcdo_create.argtypes = [C.c_int] * 4 + [C.c_double] * 4 + [C.c_int]
        L.pcdo_create.restype = C.c_void_p
        L.pcdo_destroy.argtypes = [C.c_void_p]
        L.pcdo_initialize_solvers.argtypes = [C.c_void_p, _dp]
        L.pcdo_perform_transport_iteration.argtypes = [C.c_void_p, ip]
        L.pcdo_perform_transport_iteration.restype = C.c_double
        L.pcdo_perform_height_map_iteration.argtypes = [C.c_void_p, C.c_int]
        L.pcdo_perform_height_map_iteration.restype = C.c_int
        L.pcdo_last_sweeps.argtypes = [C.c_void_p]
        L.pcdo_last_sweeps.restype = C.c_int
        L.pcdo_get_field.argtypes = [C.c_void_p, C.c_int, _dp]
        L.pcdo_get_field.restype = C.c_long
        L.pcdo_set_field.argtypes = [C.c_void_p, C.c_int, _dp]
        L.pcdo_set_field.restype = C.c_int
        L.pcdo_inverted_transport_map.argtypes = [C.c_void_p, _dp, _dp]
        L.pcdo_inverted_transport_map.restype = C.c_long
        L.pcdo_stage_errors.argtypes = [C.c_void_p]
        L.pcdo_stage_raster.argtypes = [C.c_void_p]
        L.pcdo_stage_raster.restype = C.c_int
        L.pcdo_stage_step.argtypes = [C.c_void_p]
        L.pcdo_stage_step.restype = C.c_double

    def poisson_lex(self, D, phi, max_iterations, tol):
        """Reference ordering (src/solver.cpp, one tile).  Returns (phi, sweeps, last_max_update)."""
        D = np.ascontiguousarray(D, dtype=np.float64)
        phi = np.array(phi, dtype=np.float64, order="C")
        h, w = D.shape
        last = C.c_double(0.0)
        n = self.lib.pcdo_poisson_lex(_p(D), _p(phi), w, h, int(max_iterations), float(tol), C.byref(last))
        return phi, n, last.value

    def poisson_rb(self, D, phi, max_iterations, tol, extra_sweeps=0):
        """Red-black ordering.  Returns (phi, sweeps_executed, converged_at, last_max_update)."""
        D = np.ascontiguousarray(D, dtype=np.float64)
        phi = np.array(phi, dtype=np.float64, order="C")
        h, w = D.shape
        last, conv = C.c_double(0.0), C.c_int(0)
        n = self.lib.pcdo_poisson_rb(_p(D), _p(phi), w, h, int(max_iterations), float(tol), int(extra_sweeps),
                                     C.byref(conv), C.byref(last))
        return phi, n, conv.value, last.value

    def subtract_average(self, raster):
        r = np.array(raster, dtype=np.float64, order="C")
        self.lib.pcdo_subtract_average(_p(r), r.size)
        return r

    def gradient(self, grid):
        g = np.ascontiguousarray(grid, dtype=np.float64)
        gx, gy = np.empty_like(g), np.empty_like(g)
        self.lib.pcdo_gradient(_p(g), g.shape[1], g.shape[0], _p(gx), _p(gy))
        return gx, gy

    def divergence(self, nx, ny):
        nx = np.ascontiguousarray(nx, dtype=np.float64)
        ny = np.ascontiguousarray(ny, dtype=np.float64)
        out = np.empty_like(nx)
        self.lib.pcdo_divergence(_p(nx), _p(ny), nx.shape[1], nx.shape[0], _p(out))
        return out

    def design(self, setup: Setup, solver_mode: int = 0) -> "OracleDesign":
        return OracleDesign(self, setup, solver_mode)


class OracleDesign:
    """The restated Caustic_design; same surface as RefDesign."""

    def __init__(self, lib: OracleLib, s: Setup, solver_mode: int = 0):
        self.o, self.s = lib, s
        self.h = lib.lib.pcdo_create(s.mesh_nx, s.mesh_ny, s.res_x, s.res_y, s.width, s.height, s.focal_l,
                                     s.thickness, solver_mode)

    def close(self):
        if self.h:
            self.o.lib.pcdo_destroy(self.h)
            self.h = None

    def initialize_solvers(self, image):
        img = np.ascontiguousarray(image, dtype=np.float64)
        assert img.shape == (self.s.res_y, self.s.res_x)
        self.o.lib.pcdo_initialize_solvers(self.h, _p(img))

    def transport_iteration(self) -> float:
        miss = C.c_int(0)
        step = self.o.lib.pcdo_perform_transport_iteration(self.h, C.byref(miss))
        if miss.value:
            raise RuntimeError("interpolation miss")
        return step

    def height_iteration(self, itr: int):
        if self.o.lib.pcdo_perform_height_map_iteration(self.h, itr):
            raise RuntimeError("interpolation miss")

    @property
    def last_sweeps(self) -> int:
        return self.o.lib.pcdo_last_sweeps(self.h)

    def get(self, name: str) -> np.ndarray:
        fid = FIELDS[name]
        n = self.o.lib.pcdo_get_field(self.h, fid, None)
        out = np.empty(max(n, 0), dtype=np.float64)
        if n > 0:
            self.o.lib.pcdo_get_field(self.h, fid, _p(out))
        if name in GRID_FIELDS:
            out = out.reshape(self.s.res_y, self.s.res_x)
        return out

    def set(self, name: str, value):
        v = np.ascontiguousarray(value, dtype=np.float64).ravel()
        assert self.o.lib.pcdo_set_field(self.h, FIELDS[name], _p(v)) == 0, name

    def inverted_transport_map(self):
        V = self.s.mesh_nx * self.s.mesh_ny
        x, y = np.empty(V), np.empty(V)
        n = self.o.lib.pcdo_inverted_transport_map(self.h, _p(x), _p(y))
        return x[:n], y[:n]

    def stage_errors(self):
        self.o.lib.pcdo_stage_errors(self.h)

    def stage_raster(self) -> bool:
        return bool(self.o.lib.pcdo_stage_raster(self.h))

    def stage_step(self) -> float:
        return self.o.lib.pcdo_stage_step(self.h)
