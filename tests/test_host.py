"""CPU: host-side logic around the hot path (libpcd_host.so): PNG ingest + gray + nearest resize
(main.cpp:13-109), solid OBJ / heightmap JSON / SVG writers (src/utils.cpp:176-307, main.cpp:111-135),
CLI flag parsing (main.cpp:139-214)."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

from conftest import ROOT

HOST_SO = os.path.join(ROOT, "poisson_caustic_design_b200", "libpcd_host.so")
_dp = C.POINTER(C.c_double)


class CliOptions(C.Structure):
    _fields_ = [("input_png", C.c_char * 1024), ("progress_out", C.c_char * 1024), ("output", C.c_char * 1024),
                ("has_progress_out", C.c_int), ("res_w", C.c_int), ("mesh_width", C.c_double), ("focal_l", C.c_double),
                ("thickness", C.c_double), ("conv_tres", C.c_double), ("threads", C.c_int), ("help", C.c_int),
                ("device", C.c_int), ("solver_path", C.c_int), ("quiet", C.c_int), ("gpus", C.c_int)]


@pytest.fixture(scope="module")
def host(pcd):
    if not os.path.exists(HOST_SO):
        from poisson_caustic_design_b200 import build
        build.build_all()
    L = C.CDLL(HOST_SO)
    L.pcd_host_last_error.restype = C.c_char_p
    L.pcd_host_load_png_gray.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(_dp)]
    L.pcd_host_free.argtypes = [C.c_void_p]
    L.pcd_host_resize_nearest.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int]
    L.pcd_host_save_solid_obj.argtypes = [_dp] * 5 + [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_char_p]
    L.pcd_host_save_heightmap_json.argtypes = [_dp, C.c_int, C.c_int, C.c_char_p]
    L.pcd_host_export_grid_svg.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_char_p, C.c_double]
    L.pcd_host_parse_cli.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(CliOptions)]
    return L


def p(a):
    return a.ctypes.data_as(_dp)


def load_png(host, path):
    w, h, ptr = C.c_int(), C.c_int(), _dp()
    rc = host.pcd_host_load_png_gray(str(path).encode(), C.byref(w), C.byref(h), C.byref(ptr))
    if rc != 0:
        raise RuntimeError(host.pcd_host_last_error().decode())
    out = np.ctypeslib.as_array(ptr, shape=(h.value, w.value)).copy()
    host.pcd_host_free(ptr)
    return out


def test_png_decoder_matches_reference_conversion(host, oracle_mod, golden, tmp_path):
    from PIL import Image
    imgs = golden("images")
    rng = np.random.RandomState(0)
    cases = {
        "rgba_siggraph": Image.fromarray(np.dstack([imgs["siggraph"], np.full(imgs["siggraph"].shape[:2], 255, np.uint8)]), "RGBA"),
        "rgb_lena": Image.fromarray(imgs["lena"][::4, ::4].copy(), "RGB"),
        "rgba_hello_nonsquare": Image.fromarray(np.dstack([imgs["hello"], rng.randint(0, 255, imgs["hello"].shape[:2]).astype(np.uint8)]), "RGBA"),
        "gray8": Image.fromarray(rng.randint(0, 255, (37, 53)).astype(np.uint8), "L"),
        "palette": Image.fromarray(imgs["lena"][::8, ::8].copy(), "RGB").convert("P", palette=Image.ADAPTIVE, colors=64),
        "gray_alpha": Image.fromarray(rng.randint(0, 255, (20, 31, 2)).astype(np.uint8), "LA"),
    }
    for name, im in cases.items():
        path = tmp_path / f"{name}.png"
        im.save(path)
        want = oracle_mod.rgba_to_gray(np.asarray(Image.open(path).convert("RGBA")))   # main.cpp:95-98 restated
        got = load_png(host, path)
        assert got.shape == want.shape, name
        assert np.array_equal(got, want), name
    # 16-bit gray: libpng's strip_16 keeps the high byte
    arr16 = rng.randint(0, 65535, (16, 24)).astype(np.uint16)
    Image.fromarray(arr16).save(tmp_path / "g16.png")
    got = load_png(host, tmp_path / "g16.png")
    v = (arr16 >> 8).astype(np.float64) / 255.0
    assert np.array_equal(got, (0.299 * v) + (0.587 * v) + (0.114 * v))
    with pytest.raises(RuntimeError):
        load_png(host, tmp_path / "missing.png")
    (tmp_path / "bad.png").write_bytes(b"not a png")
    with pytest.raises(RuntimeError):
        load_png(host, tmp_path / "bad.png")


def test_resize_nearest(host, oracle_mod):
    rng = np.random.RandomState(1)
    src = rng.rand(50, 70)
    for (nw, nh) in ((400, 285), (35, 25), (70, 50), (64, 200)):
        dst = np.empty((nh, nw))
        host.pcd_host_resize_nearest(p(src), 70, 50, p(dst), nw, nh)
        assert np.array_equal(dst, oracle_mod.resize_nearest(src, nw, nh))


def test_solid_obj_is_byte_identical_to_the_reference(host, golden, oracle_mod, tmp_path):
    """C1's output.obj from the reference run: md5 e0aa88f3....  Same vertices in -> same bytes out."""
    g = golden("full_c1")
    nx, ny = int(g["params"][0]), int(g["params"][1])
    width, height, thickness = float(g["params"][4]), float(g["params"][5]), float(g["params"][7])
    j, i = np.meshgrid(np.arange(nx), np.arange(ny))
    sx = (j.astype(np.float64) * width / (nx - 1)).ravel()          # src/mesh.cpp:50-51
    sy = (i.astype(np.float64) * height / (ny - 1)).ravel()
    sz = np.ascontiguousarray(g["source_z"])
    out = tmp_path / "output.obj"
    assert host.pcd_host_save_solid_obj(p(sx), p(sy), p(sz), p(sx), p(sy), nx, ny, width, height, thickness, str(out).encode()) == 0
    data = out.read_bytes()
    assert hashlib.md5(data).hexdigest() == bytes(g["obj_md5"]).decode() == "e0aa88f3faab2d417ef7fc13fbee27c6"
    lines = data.decode().split("\n")
    assert len(lines) == int(g["obj_nlines"][0])
    assert "\n".join(lines[:12]) == bytes(g["obj_head"]).decode()
    assert sum(1 for l in lines if l.startswith("v ")) == 2 * nx * ny
    # the perimeter walk repeats the top-right and bottom-right corners (src/utils.cpp:176-196): 2nx + 2ny - 2 entries
    assert sum(1 for l in lines if l.startswith("f ")) == 4 * (nx - 1) * (ny - 1) + 2 * (2 * (nx + ny) - 2)


def test_heightmap_json_and_svg_format(host, tmp_path):
    import json
    h = np.array([[0.0, 1.5, -2.25e-7], [1e10, 3.0, 0.1]])
    out = tmp_path / "heightmap.json"
    assert host.pcd_host_save_heightmap_json(p(h), 3, 2, str(out).encode()) == 0
    text = out.read_text()
    assert text == "[\n  [0, 1.5, -2.25e-07],\n  [1e+10, 3, 0.1]\n]\n"          # default ostream formatting (main.cpp:118-133)
    assert np.allclose(np.array(json.loads(text)), h)
    nx, ny = 4, 3
    j, i = np.meshgrid(np.arange(nx), np.arange(ny))
    px, py = (j / (nx - 1) * 0.5).ravel().astype(np.float64), (i / (ny - 1) * 0.25).ravel().astype(np.float64)
    svg = tmp_path / "grid.svg"
    assert host.pcd_host_export_grid_svg(p(px), p(py), nx, ny, 0.5, 0.25, str(svg).encode(), 1.0) == 0
    s = svg.read_text()
    assert s.startswith('<?xml version="1.0" encoding="UTF-8" ?>\n<svg width="1000" height="500"')
    assert s.count("<path ") == nx + ny
    assert '<path d="M0.000000,0.000000L333.333333,0.000000L666.666667,0.000000L1000.000000,0.000000" fill="none"' in s
    # column paths of a non-square mesh keep the reference's `i < res_x - 1` test: a trailing L (src/utils.cpp:298)
    assert 'M0.000000,0.000000L0.000000,250.000000L0.000000,500.000000L"' in s


def test_svg_writer_matches_reference_bytes(host, ref, oracle_mod, tmp_path):
    """Build container only: the reference's own export_grid_to_svg on the same points."""
    s = oracle_mod.Setup(12, 7, 48, 28, 0.5, 0.5 * 7 / 12, 1.5, 0.1)
    d = ref.design(s)
    rng = np.random.RandomState(2)
    d.initialize_solvers(rng.rand(28, 48))
    d.transport_iteration()
    a = tmp_path / "ref.svg"
    ref.lib.ref_cd_export_parameterization_svg(d.h, str(a).encode(), 0.5)
    x, y = d.get("target_x"), d.get("target_y")
    b = tmp_path / "mine.svg"
    assert host.pcd_host_export_grid_svg(p(x), p(y), 12, 7, s.width, s.height, str(b).encode(), 0.5) == 0
    assert a.read_bytes() == b.read_bytes()
    # and the OBJ writer against the reference's on a non-square mesh
    for it in range(2):
        d.height_iteration(it)
    d.save_obj(str(tmp_path / "ref.obj"))
    sx, sy, sz = d.get("source_x"), d.get("source_y"), d.get("source_z")
    assert host.pcd_host_save_solid_obj(p(sx), p(sy), p(sz), p(sx), p(sy), 12, 7, s.width, s.height, s.thickness,
                                        str(tmp_path / "mine.obj").encode()) == 0
    assert (tmp_path / "ref.obj").read_bytes() == (tmp_path / "mine.obj").read_bytes()
    d.close()


def parse(host, args):
    argv = (C.c_char_p * (len(args) + 1))(b"caustic_design", *[a.encode() for a in args])
    o = CliOptions()
    rc = host.pcd_host_parse_cli(len(args) + 1, argv, C.byref(o))
    return rc, o


def test_cli_defaults_and_float_flags(host):
    rc, o = parse(host, [])
    assert rc == 0                                                               # main.cpp:175-183
    assert (o.res_w, o.mesh_width, o.focal_l, o.thickness, o.threads, o.conv_tres) == (100, 1.0, 1.5, 0.2, 1, 0.01)
    assert o.progress_out == b"./" and o.output == b"./" and not o.has_progress_out
    # README.md:121 command; numeric flags are ValueFlag<float> widened to double (main.cpp:147-151)
    rc, o = parse(host, ["--input_png=../img/siggraph.png", "--res_w=100", "--mesh_width=0.5", "--focal_l=1.5",
                         "--thickness=0.1", "--conv_tres=0.01", "--output=../output/"])
    assert rc == 0 and o.input_png == b"../img/siggraph.png" and o.output == b"../output/"
    assert o.thickness == float(np.float32(0.1)) != 0.1 and o.conv_tres == float(np.float32(0.01))
    assert o.mesh_width == 0.5 and o.res_w == 100
    rc, o = parse(host, ["--res_w", "64", "--progress_out", "/tmp/p/", "--threads", "8"])   # space-separated form
    assert rc == 0 and o.res_w == 64 and o.has_progress_out and o.progress_out == b"/tmp/p/" and o.threads == 8
    assert parse(host, ["--help"])[1].help == 1 and parse(host, ["-h"])[1].help == 1
    assert parse(host, ["--nonsense=1"])[0] == 1
    assert parse(host, ["--res_w=abc"])[0] == 1
    assert parse(host, ["--res_w"])[0] == 1
