#!/usr/bin/env python
"""bench.py -- transport iterations/s and Poisson-sweep GB/s at a 1024x1024 grid (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload c4|c1|c2k|c4k|c5|c5slab] [--distribute] [--backend sor|dct]

WHAT IS TIMED.  The workload is BASELINE.json configs[3]: a caustic lens designed from a synthetic 1024x1024
high-contrast density, mesh 256x256.  One "step" = one optimal-transport iteration (dual-cell areas -> density
mismatch -> rasterise -> mean removal -> Poisson solve (~10^4 red-black SOR sweeps, tol 1e-7, warm-started) ->
gradient step on the mesh vertices).  The timed steps are CONSECUTIVE ITERATIONS OF COMPLETE DESIGNS, each design
run from iteration 0 until its step size falls below conv_tres = 0.01 (main.cpp:243-256); when a design has
converged the context is re-initialised (untimed, like the reference's initialize_solvers) and the next design
starts.  `--steps 0` (the default) times exactly one complete design.  The warm-up is one complete design (at
least W iterations).  L2 is flushed (256 MiB memset, untimed) before every timed step.  `design` in the output is
the whole-design figure (iterations to convergence / device time of those iterations).

N > 1 (torchrun, one process per GPU): the headline stays the same metric on the same workload -- one lens design
per GPU, no data-path collective (weak scaling; a 16 MB design does not shard) -- and the SAME RUN measures the
row-slab Poisson solver that large grids use in the `slab` record:
  slab.c5      ONE 8192x8192 Poisson problem cut into N row slabs, ghost rows stored into the neighbour's HBM by
               the pass kernel itself (NVLink peer memory): sweeps/s (strong scaling), per-GPU algorithmic GB/s,
               and a BIT-IDENTITY check of the N-GPU field against a single-GPU solve of the same problem after
               the same number of sweeps (the bench FAILS on a mismatch);
  slab.c4grid  the 1024x1024 grid of the headline cut the same way (expected to scale poorly: 128 rows per GPU
               at N = 8 -- reported because the metric asks for it).
At N = 1 the `slab` record holds the single-GPU solver on the same two problems (the denominators).

Keys beyond the base contract:
  roofline      dominant kernel (the resident SOR kernel): algorithmic bytes = 24 B x W x H x sweeps
                (read phi, read D, write phi per cell per sweep) / CUDA-event time of those launches on the
                launching stream, against MEASURED_PEAKS.json's HBM copy bandwidth; `on_chip` adds the floors that
                actually bound a kernel whose fields never leave the SMs.
  cpu_baseline  the reference's own CPU code (oracle/_ref) on this box's host cores: iterations 0 and 1 of the
                same design, measured in full (no extrapolation).
  e2e           the same steps through the host-buffer path: per step the mesh vertices are uploaded from pinned
                host memory and step/vertices/errors/vertex-gradients are read back (wall clock, L2 flushed).
  --impl reference: the reference's CPU implementation only (none of this repo's kernels), REAL iterations of
                the same design from iteration 0, as many of the K steps as fit a 170 s budget.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "transport_iters_per_sec"
UNIT = "iters/s"
CONV_TRES = 0.01          # main.cpp:183 (default of --conv_tres)
MAX_TRANSPORT_ITERS = 50  # main.cpp:243
REFERENCE_BUDGET_S = 170.0

WORKLOADS = {
    # name: (res_w, domain W, domain H, seed, description)
    "c4": (256, 1024, 1024, 1024, "synthetic 1024x1024 high-contrast density, mesh 256x256 (BASELINE.json configs[3])"),
    "c1": (100, 400, 400, 400, "synthetic 400x400 density, mesh 100x100 (size of BASELINE.json configs[0])"),
    "c2k": (512, 2048, 2048, 2048, "synthetic 2048x2048 density, mesh 512x512 (wavefront K-SOR path)"),
    "c4k": (1024, 4096, 4096, 4096, "synthetic 4096x4096 density, mesh 1024x1024 (wavefront K-SOR path)"),
    "c5": (2048, 8192, 8192, 8192, "synthetic 8192x8192 density, mesh 2048x2048 (BASELINE.json configs[4])"),
}
BYTES_PER_CELL_SWEEP = 24.0
# Floors of the resident kernel at 1024^2 measured on this chip in round 1 (tools/ll_latency.cu):
# the per-phase halo exchange of 1024 16-byte messages through L2 with ZERO compute sustains 1528 cycles per phase
# (chain2) = 1.55 us per sweep; 9 fp64 instructions per cell update at 64 fp64 lanes per SM and clock = 0.5 us.
RESIDENT_FLOORS_1024 = {"exchange_floor_us_per_sweep": 1.55, "fp64_floor_us_per_sweep": 0.50,
                        "source": "tools/ll_latency.cu chain2 (zero-compute halo chain, 148 CTAs); 9 fp64 instr/cell at 64 lanes/SM/clk"}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def wave_traffic(traffic, cells, sweeps_per_launch=2.0):
    """DRAM bytes of one launch of the wavefront kernel over `cells` cells (a launch = sweeps_per_launch / 2 passes),
    scaled from the committed 8192^2 capture; None when the working set (3 arrays) is not far beyond the 126 MB L2,
    where the capture does not transfer."""
    w = (traffic or {}).get("sor_wave_kernel")
    if not w or 24.0 * cells < 4 * 126e6:
        return None
    return int(w["dram_bytes_per_cell_per_pass"] * cells * sweeps_per_launch / 2.0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(name: str, rank: int = 0):
    from poisson_caustic_design_b200 import synth
    res_w, W, H, seed, desc = WORKLOADS[name]
    img = synth.synth_density(W, H, seed + rank)
    setup = synth.Setup(res_w, W, H, mesh_width=1.0, focal_l=1.5, thickness=0.2)
    assert (setup.res_x, setup.res_y) == (W, H)
    return setup, img, desc


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (oracle/_ref), else the oracle port -- REAL iterations
# ---------------------------------------------------------------------------------------------------
class CpuDesign:
    """The same design on the host cores: oracle/_ref (the reference's own code, all the threads it can use) when
    it was built, else the plain-C port (one thread)."""

    def __init__(self, workload: str):
        from oracle import oracle as O
        self.setup, self.img, self.desc = make_workload(workload)
        s = self.setup
        self.osetup = O.Setup(s.mesh_nx, s.mesh_ny, s.res_x, s.res_y, s.width, s.height, s.focal_l, s.thickness)
        self.d = None
        if O.RefLib.available():
            self.kind, self.lib = "reference", O.RefLib()
            self.threads = self._pick_threads()
        else:
            self.kind, self.lib, self.threads = "port", O.OracleLib(), 1

    def _pick_threads(self):
        # the reference runs floor(sqrt(min(threads, cores)))^2 tiles (src/solver.cpp:73-81) and spawns + joins them
        # every sweep: probe a few square counts on this grid and keep the fastest
        cores = os.cpu_count() or 1
        W, H = self.setup.res_x, self.setup.res_y
        rng = np.random.RandomState(0)
        D = rng.standard_normal((H, W))
        D -= D.mean()
        best = None
        for t in sorted({s * s for s in (1, 2, 3, 4, 6, 8, int(cores ** 0.5)) if s * s <= cores}):
            _, dt = self.lib.poisson_solver(D, np.zeros_like(D), 8, 0.0, threads=t, timed=True)
            if best is None or dt < best[1]:
                best = (t, dt)
        return best[0]

    def warm(self):
        """Bounded warm-up step: pages the library in and spins the worker threads once (40 sweeps)."""
        W, H = self.setup.res_x, self.setup.res_y
        D = np.zeros((H, W))
        D[H // 2, W // 2] = 1.0
        D -= D.mean()
        if self.kind == "reference":
            self.lib.poisson_solver(D, np.zeros_like(D), 40, 0.0, threads=self.threads)
        else:
            self.lib.poisson_lex(D, np.zeros_like(D), 40, 0.0)

    def init(self):
        if self.d is not None:
            self.d.close()
        self.d = self.lib.design(self.osetup, threads=self.threads) if self.kind == "reference" else self.lib.design(self.osetup)
        t0 = time.perf_counter()
        self.d.initialize_solvers(self.img)
        return time.perf_counter() - t0

    def iterate(self):
        t0 = time.perf_counter()
        step = self.d.transport_iteration()
        return step, time.perf_counter() - t0

    def close(self):
        if self.d is not None:
            self.d.close()
            self.d = None


def run_cpu_iterations(cpu: CpuDesign, steps: int, budget_s: float):
    """Consecutive iterations of complete designs from iteration 0 (the GPU arm's schedule), real and in full, until
    `steps` are done or the next one would not fit the budget.  Returns (times, step sizes, init seconds)."""
    times, sizes = [], []
    t_init = cpu.init()
    it_in_design = 0
    t_begin = time.perf_counter()
    while len(times) < steps:
        if times and (time.perf_counter() - t_begin) + max(times) > budget_s:
            break
        step, dt = cpu.iterate()
        times.append(dt)
        sizes.append(step)
        it_in_design += 1
        if (step < CONV_TRES or it_in_design == MAX_TRANSPORT_ITERS) and len(times) < steps:
            cpu.init()
            it_in_design = 0
    return times, sizes, t_init


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    cpu = CpuDesign(args.workload)
    setup = cpu.setup
    for _ in range(args.warmup):
        cpu.warm()
    steps_req = args.steps if args.steps > 0 else 8
    times, sizes, t_init = run_cpu_iterations(cpu, steps_req, REFERENCE_BUDGET_S)
    cpu.close()
    t = float(np.sum(times))
    value = len(times) / t
    sample = (f"{cpu.kind} Caustic_design::perform_transport_iteration, iterations 0..{len(times) - 1} of the design, each run in "
              f"full on {cpu.threads} threads (no extrapolation); {steps_req} requested, {len(times)} fit the {REFERENCE_BUDGET_S:.0f} s budget")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "steps_requested": steps_req, "extrapolated": False,
        "warmup": args.warmup, "ms_per_step": t * 1e3 / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": cpu.desc, "mesh": [setup.mesh_nx, setup.mesh_ny], "domain": [setup.res_x, setup.res_y],
                   "conv_tres": CONV_TRES,
                   "schedule": "consecutive iterations of complete designs from iteration 0 (re-initialised, untimed, at convergence)",
                   "parallelism": f"cpu threads={cpu.threads} (reference tile grid)",
                   "l2": "n/a (host arm)", "solver_path": "reference poisson_solver (lexicographic SOR, thread tiles)", "backend": "sor"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.threads, "kind": cpu.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "step_sizes": sizes[:4], "s_per_step": [round(x, 3) for x in times], "init_s": t_init,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
class DesignRunner:
    """Consecutive iterations of complete designs on one GPU (see the module docstring)."""

    def __init__(self, P, setup, img, device: int, backend: str, hook_factory=None):
        self.P, self.setup, self.img, self.device = P, setup, img, device
        self.backend = backend
        self.hook_factory, self.hook = hook_factory, None
        self.cd = P.from_setup(setup, device=device, solver_path=P.SOLVER_DCT if backend == "dct" else P.SOLVER_AUTO)  # dct: the opt-in direct solve
        self.it_in_design = 0
        self.designs_completed = 0
        self.design_lengths = []

    def init(self):
        if self.hook is not None:
            self.hook.close()
            self.hook = None
        self.cd.initialize_solvers(self.img)
        if self.hook_factory is not None:
            self.hook = self.hook_factory(self.cd)
        self.it_in_design = 0

    def after_step(self, step: float, more: bool):
        self.it_in_design += 1
        if step < CONV_TRES or self.it_in_design == MAX_TRANSPORT_ITERS:
            self.designs_completed += 1
            self.design_lengths.append(self.it_in_design)
            if more:
                self.init()

    def timed_steps(self, n: int, sync):
        """n steps, each: L2 flush (untimed) -> device-timed iteration.  Returns per-step records."""
        P, recs = self.P, []
        for k in range(n):
            cd = self.cd
            cd.flush_l2()
            l0 = P.launch_count()
            cd.event_record(0)
            step = cd.perform_transport_iteration()
            cd.event_record(1)
            ms = cd.event_elapsed_ms(0, 1)
            info = cd.last_solve_info()
            recs.append({"ms": ms, "step": step, "sweeps": info["sweeps"], "kernel_ms": info["kernel_ms"],
                         "solve_launches": info["launches"], "launches": P.launch_count() - l0, "path": info["path"],
                         "exchange": cd.resident_exchange, "it": self.it_in_design})
            self.after_step(step, more=k + 1 < n)
        return recs

    def e2e_steps(self, n: int, sync, pin):
        """The same schedule through host buffers: per step H2D of the mesh vertices (public members in the reference)
        from pinned memory, the iteration, D2H of step / vertices / errors / vertex gradients.  Wall clock."""
        V = self.setup.mesh_nx * self.setup.mesh_ny
        h_tx, h_ty = pin(V), pin(V)
        o_tx, o_ty, o_err, o_vgx, o_vgy = pin(V), pin(V), pin(V), pin(V), pin(V)
        total = 0.0
        fresh = True
        for k in range(n):
            cd = self.cd
            if fresh:
                cd.get_into("target_x", h_tx)
                cd.get_into("target_y", h_ty)
                fresh = False
            cd.flush_l2()
            sync()
            t0 = time.perf_counter()
            cd.set_from("target_x", h_tx)
            cd.set_from("target_y", h_ty)
            step = cd.perform_transport_iteration()     # returns the step size (device -> host scalar)
            cd.get_into("target_x", o_tx)
            cd.get_into("target_y", o_ty)
            cd.get_into("errors", o_err)
            cd.get_into("vertex_gradient_x", o_vgx)
            cd.get_into("vertex_gradient_y", o_vgy)
            total += time.perf_counter() - t0
            h_tx[:] = o_tx
            h_ty[:] = o_ty
            before = self.designs_completed
            self.after_step(step, more=k + 1 < n)
            fresh = self.designs_completed != before
        return total, {"h2d_bytes_per_step": 2 * V * 8, "d2h_bytes_per_step": 5 * V * 8 + 8}

    def close(self):
        if self.hook is not None:
            self.hook.close()
        self.cd.close()


def slab_problem(torch, W: int, H: int, device):
    """Zero-mean right-hand side of the slab records, generated on the device (same formula as round 1)."""
    yy = ((torch.arange(H, dtype=torch.float64, device=device) + 0.5) / H)[:, None]
    xx = ((torch.arange(W, dtype=torch.float64, device=device) + 0.5) / W)[None, :]
    D = torch.cos(3 * np.pi * xx) * torch.cos(2 * np.pi * yy) + 0.3 * torch.cos(17 * np.pi * xx) * torch.cos(11 * np.pi * yy)
    return D.contiguous(), torch.zeros((H, W), dtype=torch.float64, device=device)


def slab_record(P, torch, dist, rank: int, local_rank: int, world: int, W: int, H: int, check_sweeps: int = 64,
                timed_sweeps: int = 256, reps: int = 3, check_every: int = 64):
    """ONE W x H Poisson problem on `world` GPUs as row slabs: sweeps/s + bit-identity against one GPU."""
    from poisson_caustic_design_b200 import slab
    dev = torch.device("cuda", local_rank)
    D, phi0 = slab_problem(torch, W, H, dev)
    peak, _ = load_peaks()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the single-GPU solver on the same problem: the denominator at N = 1, the checker at N > 1
    sv = P.Solver(W, H, local_rank, P.SOLVER_AUTO)
    sv.load_device(D.data_ptr(), phi0.data_ptr())
    sv.run(check_sweeps, 0.0)
    want = torch.empty_like(phi0)
    sv.store_device(want.data_ptr())
    rec = {"domain": [W, H], "n_gpus": world, "scaling": "strong", "checked_sweeps": check_sweeps}
    if world == 1:
        best = None
        for _ in range(reps):
            sv.load_device(None, phi0.data_ptr())
            info = sv.run(timed_sweeps, 0.0)
            best = info if best is None or info["kernel_ms"] < best["kernel_ms"] else best
        t = best["kernel_ms"] * 1e-3
        rec.update({"mode": "single GPU: " + best["path"], "bit_identical_to_1gpu": True, "launches": best["launches"]})
        sv.close()
    else:
        sv.close()
        row0, rows = slab.partition(H, world, rank)
        eng = slab.CudaSlabEngine(W, H, row0, rows, local_rank)
        eng.load_device(D.data_ptr(), phi0.data_ptr())
        info = slab.solve(eng, dist, rank, world, check_sweeps, 0.0, check_every)
        got = torch.zeros_like(phi0)
        eng.store_device(got.data_ptr())
        torch.cuda.synchronize()
        ok = torch.tensor([int(torch.equal(got[row0:row0 + rows], want[row0:row0 + rows]))], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) != 1:
            raise SystemExit(f"bench.py: the {world}-GPU slab solve of the {W}x{H} problem is NOT bit-identical to the "
                             f"single-GPU solve after {check_sweeps} sweeps (mode {info['mode']})")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best_ms = None
        for _ in range(reps):
            eng.load_device(D.data_ptr(), phi0.data_ptr())
            sync()
            e0.record()
            info = slab.solve(eng, dist, rank, world, timed_sweeps, 0.0, check_every)
            e1.record()
            sync()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            best_ms = float(ms.item()) if best_ms is None else min(best_ms, float(ms.item()))
        t = best_ms * 1e-3
        GH, TS = eng.GH, eng.TS
        rec.update({"mode": info["mode"], "bit_identical_to_1gpu": True, "check_every": check_every,
                    "halo_bytes_per_sweep_per_gpu": (2 if world > 2 else 1) * GH * W * 8 // TS,   # a slab with two neighbours
                    "rows_per_gpu": rows, "comm": "ghost rows (GH=%d per side, once per %d sweeps) stored into the neighbour's HBM by the "
                                                   "pass kernel over NVLink peer memory; one all-reduce(MAX) of the maxima per %d sweeps" % (GH, TS, check_every)})
        eng.close()
    gbs = BYTES_PER_CELL_SWEEP * W * H * timed_sweeps / t / 1e9
    rec.update({"timed_sweeps": timed_sweeps, "sweeps_per_s": timed_sweeps / t, "us_per_sweep": t * 1e6 / timed_sweeps,
                "algorithmic_gbs": gbs, "per_gpu_algorithmic_gbs": gbs / world, "per_gpu_frac_of_hbm_peak": gbs / world / peak})
    del D, phi0, want
    torch.cuda.empty_cache()
    return rec


def run_b200(args, rank: int, local_rank: int, world: int):
    import torch
    import poisson_caustic_design_b200 as P

    if P.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # --distribute: ONE design for the whole job; its Poisson solves are spread over the ranks as row slabs (strong
    # scaling), every other stage is replicated.  Default: one design per GPU (replicas, weak scaling).
    distributed = bool(args.distribute)
    setup, img, desc = make_workload(args.workload, 0 if distributed else rank)
    W, H, V = setup.res_x, setup.res_y, setup.mesh_nx * setup.mesh_ny
    hook_factory = None
    hooks = []
    if distributed:
        from poisson_caustic_design_b200 import slab

        def hook_factory(cd):
            h = slab.SlabSolveHook(cd, dist, rank, world, local_rank)
            hooks.append(h)
            return h
    run = DesignRunner(P, setup, img, local_rank, args.backend, hook_factory)
    designs = 1 if distributed else world

    # ---- warm-up: one complete design (>= W iterations) -------------------------------------------
    run.init()
    warm = run.timed_steps(1, barrier)
    while run.designs_completed == 0 or len(warm) < args.warmup:
        warm += run.timed_steps(1, barrier)
    n_it = run.design_lengths[0]
    steps = args.steps if args.steps > 0 else n_it

    # ---- device-timed steps ---------------------------------------------------------------------
    run.init()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    barrier()
    wall0 = time.perf_counter()
    recs = run.timed_steps(steps, barrier)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if sampler else None
    dev_ms = float(sum(r["ms"] for r in recs))
    t_max = dev_ms
    if dist is not None:
        t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_max = float(t.item())
    value = designs * steps / (t_max * 1e-3)
    for h in hooks:
        if h.error is not None:
            raise h.error

    # ---- end to end through host buffers: the same schedule, re-initialised ------------------------
    pin = lambda n: torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()  # noqa: E731
    run.init()
    barrier()
    e_wall, e_bytes = run.e2e_steps(steps, torch.cuda.synchronize, pin)
    barrier()
    if dist is not None:
        t = torch.tensor([e_wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_wall = float(t.item())
    e2e = {"value": designs * steps / e_wall, "unit": UNIT, **e_bytes,
           "note": "same steps (iterations from 0 of re-initialised designs), L2 flushed before each, wall clock"}
    mode_used = hooks[-1].solves[-1]["mode"] if hooks and hooks[-1].solves else None
    run.close()

    # ---- the row-slab solver in the same run ---------------------------------------------
    slab_rec = None
    if not args.no_slab and args.workload == "c4" and not distributed:
        slab_rec = {"c5": slab_record(P, torch, dist, rank, local_rank, world, 8192, 8192),
                    "c4grid": slab_record(P, torch, dist, rank, local_rank, world, 1024, 1024, timed_sweeps=1024)}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    sweeps = sum(r["sweeps"] for r in recs)
    kernel_ms = sum(r["kernel_ms"] for r in recs)
    solve_launches = sum(r["solve_launches"] for r in recs)
    launches = sum(r["launches"] for r in recs)
    path = recs[-1]["path"]
    alg_bytes = BYTES_PER_CELL_SWEEP * W * H * sweeps
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    if distributed:
        achieved /= world   # per GPU: every rank moves 1/world of the cells
    traffic = load_traffic()
    us_per_sweep = kernel_ms * 1e3 / sweeps if sweeps else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (traffic["dram_bytes_per_launch"] if path == "resident" else wave_traffic(traffic, W * H // (world if distributed else 1), sweeps / max(solve_launches, 1))) if traffic else None,
                "kernel": {"resident": "sor_resident_deep_kernel" if recs[-1].get("exchange") == 2 else "sor_resident_kernel", "tiled": "sor_wave_kernel", "dct": "dct_gemm_kernel"}.get(path, "sor_colour_kernel"),
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes / max(solve_launches, 1) / (world if distributed else 1),
                "sweeps_per_launch": sweeps / max(solve_launches, 1),
                "kernel_ms_per_launch": kernel_ms / max(solve_launches, 1),
                "kernel_share_of_step": kernel_ms / dev_ms if dev_ms > 0 else None,
                "us_per_sweep": us_per_sweep,
                "note": ("phi and D live in registers/shared memory for the whole solve: DRAM traffic per launch is one read "
                         "of both fields, far below the algorithmic bytes, so HBM is NOT this kernel's roof (see on_chip)" if path == "resident" else
                         "temporal blocking: 2 sweeps per HBM pass (12 B/cell/sweep of real traffic), one persistent launch per block of passes")}
    if path == "resident" and (W, H) == (1024, 1024) and us_per_sweep:
        # the deep-halo kernel exchanges once per sweep (same message volume, half the round trips): its chain floor is ONE
        # chain2 exchange per sweep, and its fp64 floor carries the redundantly updated colour-0 halo cells (8/7 at 7 rows/CTA)
        deep = recs[-1].get("exchange") == 2
        xf = RESIDENT_FLOORS_1024["exchange_floor_us_per_sweep"] * (0.5 if deep else 1.0)
        ff = RESIDENT_FLOORS_1024["fp64_floor_us_per_sweep"] * (8.0 / 7.0 if deep else 1.0)
        roofline["on_chip"] = {"exchange_floor_us_per_sweep": xf, "fp64_floor_us_per_sweep": ff,
                               "exchanges_per_sweep": 1 if deep else 2, "source": RESIDENT_FLOORS_1024["source"],
                               "us_per_sweep": us_per_sweep, "frac_of_exchange_floor": xf / us_per_sweep,
                               "frac_of_fp64_floor": ff / us_per_sweep}
    # the whole-design figure: a complete design inside the timed window if there is one, else the warm-up design
    src, src_name = (recs[:n_it], "timed window") if steps >= n_it else (warm[:n_it], "warm-up design (first launches included)")
    d_ms = sum(r["ms"] for r in src)
    design = {"iterations_to_convergence": n_it, "conv_tres": CONV_TRES, "device_ms": d_ms, "iters_per_s": n_it / (d_ms * 1e-3),
              "sweeps": sum(r["sweeps"] for r in src), "from": src_name}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": len(warm),
        "ms_per_step": t_max / steps, "higher_is_better": True, "scaling": "strong" if distributed else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "mesh": [setup.mesh_nx, setup.mesh_ny], "domain": [W, H], "conv_tres": CONV_TRES,
                   "schedule": "consecutive iterations of complete designs from iteration 0 (re-initialised, untimed, at convergence)",
                   "parallelism": ("1 GPU" if world == 1 else
                                   f"ONE design, Poisson solves on row slabs x{world} ({mode_used} exchange), other stages replicated"
                                   if distributed else f"replicas x{world} (one lens design per GPU, no collective; the row-slab solver is measured in `slab`)"),
                   "l2": "flushed before every step (256 MiB memset, untimed)", "solver_path": path, "backend": args.backend},
        "design": design,
        "poisson_sweeps_per_sec": designs * sweeps / (kernel_ms * 1e-3) if kernel_ms > 0 else None,
        "poisson_sweeps_per_step": sweeps / steps,
        "poisson_gbs": achieved,
        "roofline": roofline,
        "e2e": e2e,
        "gpu_launches": int(launches * world),
        "clocks": clocks,
        "step_sizes": [r["step"] for r in recs[:4]],
        "sweeps_by_step": [r["sweeps"] for r in recs],
        "ms_by_step": [round(r["ms"], 3) for r in recs],
        "wall_ms_per_step": wall * 1e3 / steps,
    }
    if slab_rec is not None:
        line["slab"] = slab_rec
    if world == 1 and not args.no_cpu:
        try:
            cpu = CpuDesign(args.workload)
            cpu.warm()
            times, sizes, _ = run_cpu_iterations(cpu, 2, 60.0)
            cpu.close()
            line["cpu_baseline"] = {"value": len(times) / float(np.sum(times)), "unit": UNIT, "cores": cpu.threads, "kind": cpu.kind,
                                    "sample": f"{cpu.kind} perform_transport_iteration, iterations 0..{len(times) - 1} of the same design run in full on "
                                              f"{cpu.threads} threads ({', '.join('%.1f s' % x for x in times)}); no extrapolation",
                                    "step_sizes": sizes}
        except Exception as e:  # the CPU leg must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_slab(args, rank: int, local_rank: int, world: int):
    """--workload c5slab: BASELINE.json configs[4] -- ONE 8192x8192 Poisson problem cut into row slabs across the
    N GPUs (strong scaling; ghost rows stored into the neighbour's HBM by the pass kernel).  One step = 64 sweeps."""
    import torch
    import torch.distributed as dist
    import poisson_caustic_design_b200 as P

    W = H = 8192
    sweeps_per_step = 64
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = P.launch_count()
    steps = args.steps if args.steps > 0 else 8
    rec = slab_record(P, torch, dist, rank, local_rank, world, W, H, timed_sweeps=steps * sweeps_per_step, reps=max(args.warmup, 1) + 1,
                      check_every=args.check_every)
    launches = P.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        peak, peak_src = load_peaks()
        t = rec["timed_sweeps"] / rec["sweeps_per_s"]
        line = {"metric": "poisson_sweeps_per_sec", "value": rec["sweeps_per_s"], "unit": "sweeps/s", "n_gpus": world, "steps": steps,
                "warmup": args.warmup, "ms_per_step": t * 1e3 / steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "synthetic 8192x8192 Poisson problem, row slabs (BASELINE.json configs[4])",
                           "domain": [W, H], "parallelism": f"row slabs x{world}, mode {rec['mode']}",
                           "sweeps_per_step": sweeps_per_step, "l2": "working set 1.6 GB >> L2"},
                "roofline": {"bound": "hbm", "achieved": rec["per_gpu_algorithmic_gbs"], "peak": peak, "unit": "GB/s",
                             "frac": rec["per_gpu_frac_of_hbm_peak"],
                             "traffic": wave_traffic(load_traffic(), W * H // world, 64), "kernel": "sor_wave_kernel", "peak_source": peak_src,
                             "note": "per-GPU algorithmic GB/s (24 B/cell/sweep); whole job = achieved x n_gpus"},
                "poisson_gbs": rec["algorithmic_gbs"], "gpu_launches": int(launches * world), "clocks": clocks, "slab": {"c5": rec},
                "e2e": {"value": rec["sweeps_per_s"], "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8 * sweeps_per_step,
                        "note": "device-resident solve; per step only the per-sweep maxima cross to the host"}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed transport iterations; 0 = exactly one complete design")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS) + ["c5slab"])
    ap.add_argument("--backend", default="sor", choices=["sor", "dct"],
                    help="Poisson backend of the design: the reference's SOR iteration (default, the parity path) or the opt-in direct DCT solve")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-slab", action="store_true", help="skip the row-slab record")
    ap.add_argument("--check-every", type=int, default=64, help="c5slab: sweeps between convergence all-reduces")
    ap.add_argument("--distribute", action="store_true",
                    help="one design for the whole job: Poisson solves spread over the GPUs as row slabs (strong scaling)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # not under torchrun: launch ourselves the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.workload == "c5slab":
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "c5slab is a GPU-only scaling workload; use --workload c5"}))
            return
        run_slab(args, rank, local_rank, world)
    elif args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
