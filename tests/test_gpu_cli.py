"""GPU: the C++17 host layer -- the `caustic_design` CLI with the reference's flags (main.cpp:137-272) and the
drop-in C++ API (poisson_solver, class Caustic_design) -- against the reference's own run of BASELINE.json
configs[0] (README.md:121 command) and against the oracle."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
PKG = os.path.join(ROOT, "poisson_caustic_design_b200")
CLI = os.path.join(PKG, "caustic_design")


def test_cli_readme_command_matches_reference_run(pcd, golden, tmp_path):
    from PIL import Image
    if not os.path.exists(CLI):
        from poisson_caustic_design_b200 import build
        build.build_all()
    g = golden("full_c1")
    Image.fromarray(golden("images")["siggraph"], "RGB").save(tmp_path / "siggraph.png")
    out_dir = str(tmp_path) + "/"
    cmd = [CLI, f"--input_png={tmp_path}/siggraph.png", "--res_w=100", "--mesh_width=0.5", "--focal_l=1.5", "--thickness=0.1",
           "--conv_tres=0.01", f"--output={out_dir}", f"--progress_out={out_dir}"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    steps = [float(m) for m in re.findall(r"Transport step size = ([0-9.]+)", r.stdout)]
    assert len(steps) == len(g["steps"]) == 13                                   # 13 iterations
    assert np.abs(np.array(steps) - g["steps"]).max() < 2e-6                     # printed with %f
    assert "starting iteration 12" in r.stdout and "Height solver done! Exporting as solidified obj" in r.stdout
    assert r.stdout.count("height max update") == 3
    lines = open(os.path.join(out_dir, "output.obj")).read().split("\n")
    assert len(lines) == int(g["obj_nlines"][0])
    v = np.array([[float(t) for t in l.split()[1:]] for l in lines if l.startswith("v ")])
    nv = 100 * 100
    zr = g["source_z"]
    rng = zr.max() - zr.min()
    assert np.abs(-v[:nv, 2] - zr).max() <= 5e-4 * rng + 1e-6                    # OBJ prints 6 significant digits
    assert np.abs(v[nv:, 2] - (-float(np.float32(0.1)))).max() < 1e-6            # back plane: -thickness - min(0, min z)
    assert lines[:2] == bytes(g["obj_head"]).decode().split("\n")[:2]
    import json
    h = np.array(json.load(open(os.path.join(out_dir, "heightmap.json"))))
    assert h.shape == (400, 400)
    hs, hr = h[::8, ::8], g["h_sub8"]
    assert np.abs((hs - hs.mean()) - (hr - hr.mean())).max() <= 1e-3 * (g["h_range"][1] - g["h_range"][0])
    assert os.path.exists(os.path.join(out_dir, "parameterization_0.svg")) and os.path.exists(os.path.join(out_dir, "parameterization_13.svg"))
    assert os.path.exists(os.path.join(out_dir, "inverted.svg"))


def test_cli_errors(pcd, tmp_path):
    if not os.path.exists(CLI):
        pytest.skip("CLI not built")
    r = subprocess.run([CLI, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--input_png" in r.stdout and "--conv_tres" in r.stdout
    r = subprocess.run([CLI, "--bogus=1"], capture_output=True, text=True)
    assert r.returncode == 1
    r = subprocess.run([CLI, f"--input_png={tmp_path}/nope.png"], capture_output=True, text=True)
    assert r.returncode != 0 and "Failed to open PNG file." in r.stderr        # main.cpp:32 (uncaught runtime_error)


def test_cpp_drop_in_api(pcd, port, oracle_mod, tmp_path):
    exe = tmp_path / "shim_test"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(PKG, "host"), "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "shim_test.cpp"), "-o", str(exe), "-L", PKG, "-lpcd_host", "-lpcd_b200",
                    f"-Wl,-rpath,{PKG}"], check=True)
    W, H, nx, ny = 64, 48, 16, 12
    rng = np.random.RandomState(4)
    D = rng.standard_normal((H, W))
    D -= D.mean()
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    img = 0.1 + np.exp(-((xx - 30) ** 2 + (yy - 20) ** 2) / 90.0)
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(struct.pack("4i", W, H, nx, ny))
        f.write(D.tobytes())
        f.write(img.tobytes())
    r = subprocess.run([str(exe), str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    buf = np.fromfile(tmp_path / "out.bin", dtype=np.uint8)
    off = 0

    def take(n, shape=None):
        nonlocal off
        a = np.frombuffer(buf[off:off + 8 * n].tobytes(), dtype=np.float64)
        off += 8 * n
        return a.reshape(shape) if shape else a

    phi = take(W * H, (H, W))
    lex = port.poisson_lex(D, np.zeros_like(D), 100000, 1e-7)[0]
    gl, gg = port.gradient(lex), port.gradient(phi)
    assert max(np.abs(gl[0] - gg[0]).max(), np.abs(gl[1] - gg[1]).max()) / np.abs(gl[0]).max() < 1e-5
    steps = take(2)
    V = nx * ny
    tp = take(3 * V, (V, 3))
    phi2 = take(W * H, (H, W))
    errors = take(V)
    again = take(1)[0]
    sp = take(3 * V, (V, 3))
    h = take(W * H, (H, W))
    caught = int(np.frombuffer(buf[off:off + 4].tobytes(), dtype=np.int32)[0])
    assert caught == 1
    s = oracle_mod.Setup(nx, ny, W, H, 0.5, 0.5 * ny / nx, 1.5, 0.1)
    od = port.design(s)
    od.initialize_solvers(img)
    want = [od.transport_iteration() for _ in range(2)]
    assert np.abs(steps - want).max() < 1e-6 * max(want)
    assert again == steps[0]          # edits of the public members (mesh put back, phi zeroed) are honoured: iteration 0 repeats
    disp = np.abs(od.get("target_x") - od.get("source_x")).max()
    assert np.abs(tp[:, 0] - od.get("target_x")).max() < 1e-6 * disp and np.abs(tp[:, 1] - od.get("target_y")).max() < 1e-6 * disp
    assert np.abs(errors - od.get("errors")).max() < 1e-6 * np.abs(od.get("errors")).max()
    assert phi2.shape == (H, W) and np.isfinite(phi2).all()
    od.set("target_x", od.get("source_x")); od.set("target_y", od.get("source_y")); od.set("phi", np.zeros((H, W)))
    assert abs(od.transport_iteration() - again) < 1e-6 * again
    od.height_iteration(0)
    zr = od.get("source_z")
    assert np.abs(sp[:, 2] - zr).max() <= 5e-4 * (zr.max() - zr.min())
    assert np.isfinite(h).all() and "height max update" in r.stdout and "built mesh" in r.stdout
