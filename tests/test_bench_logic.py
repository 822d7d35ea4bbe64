"""CPU: the measurement code itself (bench.py host logic, tools/wave_trace.py) -- the schedule of the timed steps
(consecutive iterations of complete designs, re-initialised at convergence), the reference arm's time budget, the
per-pass DRAM-traffic scaling and the trace parser.  No GPU, no oracle solve: fakes stand in for the designs."""
import importlib.util
import json
import os
import struct
import sys

import numpy as np

from conftest import ROOT


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


bench = _load("bench_mod", os.path.join(ROOT, "bench.py"))


class FakeCpu:
    """A design that converges after `n` iterations (step sizes 0.05, 0.04, ..., last one below conv_tres)."""

    def __init__(self, n, dt=0.0):
        self.n, self.dt, self.inits, self.it = n, dt, 0, 0

    def init(self):
        self.inits += 1
        self.it = 0
        return 0.0

    def iterate(self):
        self.it += 1
        return (0.005 if self.it == self.n else 0.05), self.dt


def test_reference_schedule_cycles_through_whole_designs():
    cpu = FakeCpu(3)
    times, sizes, _ = bench.run_cpu_iterations(cpu, 7, 1e9)
    assert len(times) == 7 and cpu.inits == 3          # designs of 3 iterations: 3 + 3 + 1, re-initialised twice
    assert [s < bench.CONV_TRES for s in sizes] == [False, False, True, False, False, True, False]
    cpu = FakeCpu(3)
    bench.run_cpu_iterations(cpu, 6, 1e9)
    assert cpu.inits == 2                              # no re-initialisation after the last timed step


def test_reference_budget_stops_before_the_next_step_would_not_fit():
    cpu = FakeCpu(100, dt=10.0)                        # reported 10 s per iteration (the wall clock barely moves here)
    times, _, _ = bench.run_cpu_iterations(cpu, 50, 5.0)
    assert len(times) == 1                             # the first step always runs; the next would exceed the budget


def test_design_runner_schedule():
    class FakeDesign:
        def __init__(self):
            self.inits, self.it = 0, 0

        def initialize_solvers(self, img):
            self.inits += 1
            self.it = 0

        def close(self):
            pass

    class FakeP:
        SOLVER_AUTO, SOLVER_DCT = 0, 4

        def from_setup(self, setup, device=0, solver_path=0):
            return FakeDesign()

    run = bench.DesignRunner(FakeP(), None, None, 0, "sor")
    run.init()
    for k in range(7):
        run.cd.it += 1
        run.after_step(0.005 if run.cd.it == 3 else 0.05, more=k + 1 < 7)
    assert run.designs_completed == 2 and run.design_lengths == [3, 3] and run.cd.inits == 3 and run.it_in_design == 1


def test_wave_traffic_scales_with_the_passes_of_a_launch():
    t = {"sor_wave_kernel": {"dram_bytes_per_cell_per_pass": 24.3}}
    cells = 8192 * 8192
    assert bench.wave_traffic(t, cells, 2) == int(24.3 * cells)
    assert bench.wave_traffic(t, cells, 64) == int(24.3 * cells * 32)
    assert bench.wave_traffic(t, 1024 * 1024, 64) is None      # L2-resident working set: the capture does not transfer
    committed = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    assert abs(committed["sor_wave_kernel"]["dram_bytes_per_cell_per_pass"] - 24.0) < 1.0


def test_wave_trace_parser(tmp_path):
    wt = _load("wave_trace_mod", os.path.join(ROOT, "tools", "wave_trace.py"))
    npass, maxc, ncta = 6, 1024, 34
    t = np.zeros((npass, maxc, 4), dtype=np.uint64)
    for i in range(npass):
        for c in range(ncta):
            start = 1_000_000 + i * 50_000 + c * 10
            t[i, c] = [start, start + 5_000, start + 45_000, start + 46_000]   # 5 us wait, 40 us compute, 1 us publish
    path = tmp_path / "x_row0.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("2i", npass, maxc))
        f.write(t.tobytes())
    s = wt.summarise(wt.load(str(path)), strips=17)
    assert s["ctas"] == ncta and s["passes"] == npass
    assert abs(s["wait_us_mean"] - 5.0) < 1e-9 and abs(s["compute_us_mean"] - 40.0) < 1e-9 and abs(s["publish_us_mean"] - 1.0) < 1e-9
    assert abs(s["pass_period_us_median"] - 50.0) < 1e-9 and len(s["compute_us_by_chunk_row"]) == 2
