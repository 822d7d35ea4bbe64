"""poisson_caustic_design_b200 -- B200-native (sm_100a) caustic-design hot path.

The product is the native code: ``libpcd_b200.so`` (hand-written CUDA kernels behind the C ABI of
``include/pcd.h``) and the C++17 host shim / CLI under ``host/`` that keeps the reference's
``poisson_solver`` / ``Caustic_design`` signatures.  This Python module is only the ctypes view of
that C ABI used by the tests and ``bench.py``; names follow the reference
(src/caustic_design.h:7-66, src/solver.h:8).

There is no CPU fallback: importing works without a GPU (so the ABI can be checked), but every
compute call fails with :class:`PcdError` unless the CUDA library is built and a device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PCD_LIB") or os.path.join(PKG_DIR, "libpcd_b200.so")  # PCD_LIB: diagnostic builds only

PCD_OK, PCD_ERR_INVALID, PCD_ERR_CUDA, PCD_ERR_NO_DEVICE, PCD_ERR_RASTER_MISS, PCD_ERR_STATE, PCD_ERR_UNSUPPORTED = range(7)
SOLVER_AUTO, SOLVER_STREAMING, SOLVER_RESIDENT, SOLVER_TILED, SOLVER_DCT = 0, 1, 2, 3, 4
SOLVER_PATH_NAMES = {SOLVER_AUTO: "auto", SOLVER_STREAMING: "streaming", SOLVER_RESIDENT: "resident", SOLVER_TILED: "tiled",
                     SOLVER_DCT: "dct"}

FIELDS = {
    "phi": 0, "h": 1, "raster": 2, "pixels": 3, "divergence": 4, "norm_x": 5, "norm_y": 6,
    "gradient_x": 7, "gradient_y": 8, "errors": 9, "target_areas": 10,
    "vertex_gradient_x": 11, "vertex_gradient_y": 12, "normals_x": 13, "normals_y": 14,
    "target_x": 15, "target_y": 16, "target_z": 17, "source_x": 18, "source_y": 19, "source_z": 20,
}
GRID_FIELDS = {"phi", "h", "raster", "pixels", "divergence", "norm_x", "norm_y", "gradient_x", "gradient_y"}


class PcdError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"pcd status {status}: {message}")
        self.status = status


class pcd_config(C.Structure):
    _fields_ = [("mesh_res_x", C.c_int), ("mesh_res_y", C.c_int), ("res_x", C.c_int), ("res_y", C.c_int),
                ("width", C.c_double), ("height", C.c_double), ("focal_l", C.c_double), ("thickness", C.c_double),
                ("device", C.c_int), ("solver_path", C.c_int)]


class pcd_solve_info(C.Structure):
    _fields_ = [("sweeps", C.c_int), ("converged_at", C.c_int), ("last_max_update", C.c_double),
                ("device_ms", C.c_double), ("kernel_ms", C.c_double), ("launches", C.c_int), ("path", C.c_int)]

    def as_dict(self):
        return {"sweeps": self.sweeps, "converged_at": self.converged_at, "last_max_update": self.last_max_update,
                "device_ms": self.device_ms, "kernel_ms": self.kernel_ms, "launches": self.launches, "path": SOLVER_PATH_NAMES.get(self.path, "?")}


_dp = C.POINTER(C.c_double)
_lib = None

# every symbol include/pcd.h declares: (name, restype, argtypes)
ABI = [
    ("pcd_abi_version", C.c_int, []),
    ("pcd_last_error", C.c_char_p, []),
    ("pcd_device_count", C.c_int, [C.POINTER(C.c_int)]),
    ("pcd_launch_count", C.c_longlong, []),
    ("pcd_create", C.c_int, [C.POINTER(pcd_config), C.POINTER(C.c_void_p)]),
    ("pcd_destroy", None, [C.c_void_p]),
    ("pcd_initialize_solvers", C.c_int, [C.c_void_p, _dp]),
    ("pcd_perform_transport_iteration", C.c_int, [C.c_void_p, _dp]),
    ("pcd_perform_height_map_iteration", C.c_int, [C.c_void_p, C.c_int, _dp]),
    ("pcd_run_transport", C.c_int, [C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_int), _dp]),
    ("pcd_field_size", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_long)]),
    ("pcd_get_field", C.c_int, [C.c_void_p, C.c_int, _dp]),
    ("pcd_set_field", C.c_int, [C.c_void_p, C.c_int, _dp]),
    ("pcd_inverted_transport_map", C.c_int, [C.c_void_p, _dp, _dp]),
    ("pcd_last_solve_info", C.c_int, [C.c_void_p, C.POINTER(pcd_solve_info)]),
    ("pcd_solve_totals", C.c_int, [C.c_void_p, C.POINTER(pcd_solve_info), C.c_int]),
    ("pcd_event_record", C.c_int, [C.c_void_p, C.c_int]),
    ("pcd_event_elapsed_ms", C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp]),
    ("pcd_flush_l2", C.c_int, [C.c_void_p]),
    ("pcd_stage_errors", C.c_int, [C.c_void_p]),
    ("pcd_stage_raster", C.c_int, [C.c_void_p]),
    ("pcd_stage_subtract_average", C.c_int, [C.c_void_p]),
    ("pcd_stage_solve_transport", C.c_int, [C.c_void_p]),
    ("pcd_stage_step", C.c_int, [C.c_void_p, _dp]),
    ("pcd_poisson_solver", C.c_int, [_dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(pcd_solve_info)]),
    ("pcd_solver_create", C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    ("pcd_solver_destroy", None, [C.c_void_p]),
    ("pcd_solver_upload", C.c_int, [C.c_void_p, _dp, _dp]),
    ("pcd_solver_download", C.c_int, [C.c_void_p, _dp]),
    ("pcd_solver_load_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("pcd_solver_store_device", C.c_int, [C.c_void_p, C.c_void_p]),
    ("pcd_solver_set_check_lag", C.c_int, [C.c_void_p, C.c_int]),
    ("pcd_solver_run", C.c_int, [C.c_void_p, C.c_int, C.c_double, C.POINTER(pcd_solve_info)]),
    ("pcd_solver_path_used", C.c_int, [C.c_void_p]),
    ("pcd_solver_resident_exchange", C.c_int, [C.c_void_p]),
    ("pcd_resident_exchange", C.c_int, [C.c_void_p]),
    ("pcd_solver_plan", C.c_int, [C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 5),
    ("pcd_slab_create", C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    ("pcd_slab_destroy", None, [C.c_void_p]),
    ("pcd_slab_device_ptrs", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    ("pcd_slab_upload", C.c_int, [C.c_void_p, _dp, _dp]),
    ("pcd_slab_download", C.c_int, [C.c_void_p, _dp]),
    ("pcd_slab_sweep_colour", C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    ("pcd_slab_pass", C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    ("pcd_slab_pass_part", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("pcd_slab_flip", C.c_int, [C.c_void_p]),
    ("pcd_slab_set_sm_reserve", C.c_int, [C.c_void_p, C.c_int]),
    ("pcd_slab_ghost_rows", C.c_int, []),
    ("pcd_slab_sweeps_per_pass", C.c_int, []),
    ("pcd_slab_current", C.c_int, [C.c_void_p]),
    ("pcd_slab_error_word", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    ("pcd_slab_has_nan", C.c_int, [C.c_void_p]),
    ("pcd_slab_clear_max", C.c_int, [C.c_void_p, C.c_int]),
    ("pcd_slab_clear_max_range", C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    ("pcd_slab_load_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("pcd_slab_store_device", C.c_int, [C.c_void_p, C.c_void_p]),
    ("pcd_set_solve_hook", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("pcd_set_tolerances", C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    ("pcd_slab_peer_handle_bytes", C.c_int, []),
    ("pcd_slab_peer_export", C.c_int, [C.c_void_p, C.c_void_p]),
    ("pcd_slab_peer_connect_ipc", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    ("pcd_slab_peer_connect_local", C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    ("pcd_slab_peer_run", C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    ("pcd_slab_peer_status", C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    ("pcd_slab_peer_error_to", C.c_int, [C.c_void_p, C.c_void_p]),
    ("pcd_multi_create", C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    ("pcd_multi_destroy", None, [C.c_void_p]),
    ("pcd_multi_device_count", C.c_int, [C.c_void_p]),
    ("pcd_multi_set_check_every", C.c_int, [C.c_void_p, C.c_int]),
    ("pcd_multi_solve", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.POINTER(pcd_solve_info)]),
    ("pcd_multi_attach", C.c_int, [C.c_void_p, C.c_void_p]),
]


def lib() -> C.CDLL:
    """Loads libpcd_b200.so (built in-tree by ``__graft_entry__.build()`` / ``build.py``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PcdError(PCD_ERR_NO_DEVICE, f"{LIB_PATH} is missing: build it with `python -m "
                           "poisson_caustic_design_b200.build` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, restype, argtypes in ABI:
            fn = getattr(L, name)
            fn.restype, fn.argtypes = restype, argtypes
        _lib = L
    return _lib


def _check(status: int):
    if status != PCD_OK:
        raise PcdError(status, lib().pcd_last_error().decode(errors="replace"))


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def device_count() -> int:
    n = C.c_int(0)
    lib().pcd_device_count(C.byref(n))
    return n.value


def launch_count() -> int:
    return int(lib().pcd_launch_count())


def poisson_solver(D, phi, width: int, height: int, max_iterations: int, convergence_threshold: float,
                   max_threads: int = 1, device: int = 0):
    """Drop-in for src/solver.h:8 (phi is updated in place when it is a C-contiguous float64 array).
    ``max_threads`` is accepted and ignored.  Returns the solve info dict."""
    D = np.ascontiguousarray(D, dtype=np.float64)
    out = phi if (isinstance(phi, np.ndarray) and phi.dtype == np.float64 and phi.flags["C_CONTIGUOUS"]) else None
    buf = out if out is not None else np.array(phi, dtype=np.float64, order="C")
    assert D.size == width * height and buf.size == width * height
    info = pcd_solve_info()
    _check(lib().pcd_poisson_solver(_p(D), _p(buf), width, height, int(max_iterations), float(convergence_threshold),
                                    device, C.byref(info)))
    if out is None:
        np.copyto(np.asarray(phi), buf.reshape(np.shape(phi)))
    return info.as_dict()


class Solver:
    """Device-resident poisson_solver state (D, phi stay in HBM between runs)."""

    def __init__(self, width: int, height: int, device: int = 0, path: int = SOLVER_AUTO):
        self.W, self.H = width, height
        self._h = C.c_void_p()
        _check(lib().pcd_solver_create(width, height, device, path, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().pcd_solver_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    @property
    def path(self) -> str:
        return SOLVER_PATH_NAMES[lib().pcd_solver_path_used(self._h)]

    def upload(self, D=None, phi=None):
        d = np.ascontiguousarray(D, dtype=np.float64) if D is not None else None
        p = np.ascontiguousarray(phi, dtype=np.float64) if phi is not None else None
        for name, a in (("D", d), ("phi", p)):
            if a is not None and a.size != self.W * self.H:   # the C ABI takes plain pointers: sizes are checked here
                raise PcdError(PCD_ERR_INVALID, f"{name} has {a.size} elements, the solver's grid {self.W}x{self.H}")
        _check(lib().pcd_solver_upload(self._h, _p(d) if d is not None else None, _p(p) if p is not None else None))

    def download(self) -> np.ndarray:
        out = np.empty((self.H, self.W), dtype=np.float64)
        _check(lib().pcd_solver_download(self._h, _p(out)))
        return out

    def load_device(self, D_dev: int | None = None, phi_dev: int | None = None):
        """D / phi from device arrays (raw pointers, e.g. ``tensor.data_ptr()``) on the solver's device."""
        _check(lib().pcd_solver_load_device(self._h, C.c_void_p(D_dev) if D_dev else None, C.c_void_p(phi_dev) if phi_dev else None))

    def store_device(self, phi_dev: int):
        _check(lib().pcd_solver_store_device(self._h, C.c_void_p(phi_dev)))

    def set_check_lag(self, lag: int):
        _check(lib().pcd_solver_set_check_lag(self._h, lag))

    def run(self, max_iterations: int, tol: float) -> dict:
        info = pcd_solve_info()
        _check(lib().pcd_solver_run(self._h, int(max_iterations), float(tol), C.byref(info)))
        return info.as_dict()

    @property
    def resident_exchange(self) -> int:
        """Resident kernel of the last run: 0 none, 1 one exchange per colour phase, 2 one per sweep (deep halos)."""
        return lib().pcd_solver_resident_exchange(self._h)


def solver_plan(width: int, height: int, sm_count: int = 148) -> dict:
    """What SOLVER_AUTO chooses for a width x height grid on a device with ``sm_count`` SMs (host logic only, no GPU)."""
    v = [C.c_int(0) for _ in range(5)]
    _check(lib().pcd_solver_plan(int(width), int(height), int(sm_count), *[C.byref(x) for x in v]))
    return {"path": SOLVER_PATH_NAMES.get(v[0].value, "?"), "rows_per_cta": v[1].value, "ctas": v[2].value,
            "transposed": bool(v[3].value), "deep_only": bool(v[4].value)}


class MultiGpuSolver:
    """One Poisson problem as row slabs on several GPUs of this process (include/pcd.h: pcd_multi_*)."""

    def __init__(self, width: int, height: int, devices):
        self.W, self.H = width, height
        devs = (C.c_int * len(devices))(*devices)
        self._h = C.c_void_p()
        _check(lib().pcd_multi_create(width, height, devs, len(devices), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().pcd_multi_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def set_check_every(self, sweeps: int):
        _check(lib().pcd_multi_set_check_every(self._h, sweeps))

    def solve(self, D_dev: int, phi_dev: int, max_iterations: int, tol: float) -> dict:
        """D_dev / phi_dev: raw device pointers (devices[0]) of W x H float64 arrays; phi in/out."""
        info = pcd_solve_info()
        _check(lib().pcd_multi_solve(self._h, C.c_void_p(D_dev), C.c_void_p(phi_dev), int(max_iterations), float(tol), C.byref(info)))
        return info.as_dict()

    def attach(self, design: "CausticDesign | None"):
        _check(lib().pcd_multi_attach(self._h, design._h if design is not None else None))


class CausticDesign:
    """ctypes mirror of ``class Caustic_design`` (src/caustic_design.h:7-66): same setters, same call
    order as main.cpp:224-262.  Fields are fetched on demand with :meth:`get`."""

    def __init__(self, device: int = 0, solver_path: int = SOLVER_AUTO):
        self._cfg = pcd_config(0, 0, 0, 0, 0.0, 0.0, 0.0, 0.0, device, solver_path)
        self._h = C.c_void_p()

    # setters, src/caustic_design.cpp:28-53
    def set_mesh_resolution(self, width: int, height: int):
        self._cfg.mesh_res_x, self._cfg.mesh_res_y = width, height

    def set_domain_resolution(self, width: int, height: int):
        self._cfg.res_x, self._cfg.res_y = width, height

    def set_mesh_size(self, width: float, height: float):
        self._cfg.width, self._cfg.height = width, height

    def set_lens_focal_length(self, focal_length: float):
        self._cfg.focal_l = focal_length

    def set_lens_thickness(self, thickness: float):
        self._cfg.thickness = thickness

    def set_solver_max_threads(self, n_threads: int):
        pass  # CPU threads have no meaning on the device path

    def close(self):
        if self._h:
            lib().pcd_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def initialize_solvers(self, image):
        img = np.ascontiguousarray(image, dtype=np.float64)
        if img.shape != (self._cfg.res_y, self._cfg.res_x):
            raise PcdError(PCD_ERR_INVALID, f"image shape {img.shape} != domain ({self._cfg.res_y}, {self._cfg.res_x})")
        self.close()
        _check(lib().pcd_create(C.byref(self._cfg), C.byref(self._h)))
        _check(lib().pcd_initialize_solvers(self._h, _p(img)))

    def perform_transport_iteration(self) -> float:
        step = C.c_double(0.0)
        _check(lib().pcd_perform_transport_iteration(self._h, C.byref(step)))
        return step.value

    def run_transport(self, max_iters: int = 50, conv_tres: float = 0.01):
        """main.cpp:243-256 in one call.  Returns the list of step sizes."""
        steps = np.zeros(max_iters, dtype=np.float64)
        n = C.c_int(0)
        _check(lib().pcd_run_transport(self._h, max_iters, float(conv_tres), C.byref(n), _p(steps)))
        return steps[: n.value].tolist()

    def perform_height_map_iteration(self, itr: int) -> float:
        upd = C.c_double(0.0)
        _check(lib().pcd_perform_height_map_iteration(self._h, itr, C.byref(upd)))
        return upd.value

    def set_tolerances(self, transport_tol: float = 0.0, height_tol: float = 0.0):
        """Stopping thresholds of the two Poisson solves (<= 0 keeps the current one); call after initialize_solvers."""
        _check(lib().pcd_set_tolerances(self._h, float(transport_tol), float(height_tol)))

    def last_solve_info(self) -> dict:
        info = pcd_solve_info()
        _check(lib().pcd_last_solve_info(self._h, C.byref(info)))
        return info.as_dict()

    @property
    def resident_exchange(self) -> int:
        """Resident kernel of the last built-in solve: 0 none, 1 one exchange per colour phase, 2 one per sweep."""
        return lib().pcd_resident_exchange(self._h)

    def solve_totals(self, reset: bool = False) -> dict:
        info = pcd_solve_info()
        _check(lib().pcd_solve_totals(self._h, C.byref(info), int(reset)))
        return info.as_dict()

    def event_record(self, slot: int):
        _check(lib().pcd_event_record(self._h, slot))

    def event_elapsed_ms(self, start: int, stop: int) -> float:
        ms = C.c_double(0.0)
        _check(lib().pcd_event_elapsed_ms(self._h, start, stop, C.byref(ms)))
        return ms.value

    def flush_l2(self):
        _check(lib().pcd_flush_l2(self._h))

    def get_into(self, name: str, out: np.ndarray):
        """Device -> caller-provided (e.g. pinned) host buffer."""
        _check(lib().pcd_get_field(self._h, FIELDS[name], _p(out)))

    def set_from(self, name: str, src: np.ndarray):
        """Caller-provided (e.g. pinned) host buffer -> device."""
        _check(lib().pcd_set_field(self._h, FIELDS[name], _p(src)))

    def get(self, name: str) -> np.ndarray:
        n = C.c_long(0)
        _check(lib().pcd_field_size(self._h, FIELDS[name], C.byref(n)))
        out = np.empty(n.value, dtype=np.float64)
        _check(lib().pcd_get_field(self._h, FIELDS[name], _p(out)))
        return out.reshape(self._cfg.res_y, self._cfg.res_x) if name in GRID_FIELDS else out

    def set(self, name: str, value):
        v = np.ascontiguousarray(value, dtype=np.float64).ravel()
        n = C.c_long(0)
        _check(lib().pcd_field_size(self._h, FIELDS[name], C.byref(n)))
        if v.size != n.value:
            raise PcdError(PCD_ERR_INVALID, f"{name}: {v.size} values, field has {n.value}")
        _check(lib().pcd_set_field(self._h, FIELDS[name], _p(v)))

    def inverted_transport_map(self):
        V = self._cfg.mesh_res_x * self._cfg.mesh_res_y
        x, y = np.empty(V), np.empty(V)
        _check(lib().pcd_inverted_transport_map(self._h, _p(x), _p(y)))
        return x, y

    # stage entry points (per-stage parity tests)
    def stage_errors(self):
        _check(lib().pcd_stage_errors(self._h))

    def stage_raster(self):
        _check(lib().pcd_stage_raster(self._h))

    def stage_subtract_average(self):
        _check(lib().pcd_stage_subtract_average(self._h))

    def stage_solve_transport(self) -> dict:
        _check(lib().pcd_stage_solve_transport(self._h))
        return self.last_solve_info()

    def stage_step(self) -> float:
        step = C.c_double(0.0)
        _check(lib().pcd_stage_step(self._h, C.byref(step)))
        return step.value


def from_setup(setup, device: int = 0, solver_path: int = SOLVER_AUTO) -> CausticDesign:
    """``setup`` has mesh_nx, mesh_ny, res_x, res_y, width, height, focal_l, thickness (main.cpp:224-235)."""
    cd = CausticDesign(device, solver_path)
    cd.set_mesh_resolution(setup.mesh_nx, setup.mesh_ny)
    cd.set_domain_resolution(setup.res_x, setup.res_y)
    cd.set_mesh_size(setup.width, setup.height)
    cd.set_lens_focal_length(setup.focal_l)
    cd.set_lens_thickness(setup.thickness)
    return cd
