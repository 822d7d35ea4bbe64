#!/usr/bin/env python
"""bench.py -- transport iterations/s and Poisson-sweep GB/s at a 1024x1024 grid (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload c4|c1|c2k|c4k|c5|c5slab] [--distribute]

One "step" = one optimal-transport iteration (dual-cell areas -> density mismatch -> rasterise -> mean
removal -> Poisson solve (~10^4 red-black SOR sweeps, tol 1e-7, warm-started) -> gradient step on the
mesh vertices) on the synthetic 1024x1024 high-contrast density of BASELINE.json configs[3] (mesh
256x256).  N>1: one process per GPU (torchrun), every rank designs its own lens (independent units,
no data-path collective) -> weak scaling.  Large grids: `--workload c5 --distribute` runs ONE
8192x8192 design whose Poisson solves are spread over the N GPUs as row slabs (ghost rows stored into the
neighbour's HBM by the pass kernel itself) -> strong scaling; `--workload c5slab` times the slab solver alone.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel (the resident SOR kernel): algorithmic bytes = 24 B x W x H x sweeps
                (read phi, read D, write phi per cell per sweep) / CUDA-event time of those launches
                on the launching stream, against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline  the reference's own CPU code (oracle/_ref) on this box's host cores, bounded sample.
  e2e           the same metric through the host-buffer path: per step the mesh vertices are uploaded
                from pinned host memory and step/vertices/errors/vertex-gradients are read back.
  --impl reference: the reference's CPU implementation only (none of this repo's kernels).
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

WORKLOADS = {
    # name: (res_w, domain W, domain H, seed, description)
    "c4": (256, 1024, 1024, 1024, "synthetic 1024x1024 high-contrast density, mesh 256x256 (BASELINE.json configs[3])"),
    "c1": (100, 400, 400, 400, "synthetic 400x400 density, mesh 100x100 (size of BASELINE.json configs[0])"),
    "c2k": (512, 2048, 2048, 2048, "synthetic 2048x2048 density, mesh 512x512 (wavefront K-SOR path)"),
    "c4k": (1024, 4096, 4096, 4096, "synthetic 4096x4096 density, mesh 1024x1024 (wavefront K-SOR path)"),
    "c5": (2048, 8192, 8192, 8192, "synthetic 8192x8192 density, mesh 2048x2048 (BASELINE.json configs[4])"),
}
# sweeps of the first transport solve (measured with the CUDA path; the reference's lexicographic
# ordering needs 3-10 % more) -- used only to extrapolate the CPU sample to a full iteration
EXPECTED_SWEEPS = {"c4": 11000, "c1": 4300, "c2k": 21500, "c4k": 48000, "c5": 85000}
BYTES_PER_CELL_SWEEP = 24.0


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


def wave_traffic(traffic, cells):
    """DRAM bytes of one wavefront pass over `cells` cells, scaled from the committed 8192^2 capture; None when the
    working set (3 arrays) is not far beyond the 126 MB L2, where the capture does not transfer."""
    w = (traffic or {}).get("sor_wave_kernel")
    if not w or 24.0 * cells < 4 * 126e6:
        return None
    return int(w["dram_bytes_per_cell_per_launch"] * cells)


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
# CPU arm: the reference's own implementation (oracle/_ref), else the oracle port
# ---------------------------------------------------------------------------------------------------
def cpu_sample(workload: str, sweeps_sample: int = 100, threads: int | None = None):
    """One bounded sample of a transport iteration on the host cores.  Returns a dict with the
    extrapolated seconds per full iteration."""
    from oracle import oracle as O
    setup, img, _ = make_workload(workload)
    osetup = O.Setup(setup.mesh_nx, setup.mesh_ny, setup.res_x, setup.res_y, setup.width, setup.height,
                     setup.focal_l, setup.thickness)
    n_full = EXPECTED_SWEEPS[workload]
    W, H = setup.res_x, setup.res_y
    if O.RefLib.available():
        ref = O.RefLib()
        cores = os.cpu_count() or 1
        if threads is None:
            # the reference runs floor(sqrt(min(threads, cores)))^2 tiles (src/solver.cpp:73-81) and
            # spawns + joins them every sweep: probe a few square counts and keep the fastest
            best = None
            rng = np.random.RandomState(0)
            D = rng.standard_normal((H, W))
            D -= D.mean()
            cands = sorted({s * s for s in (1, 2, 3, 4, 6, 8, int(cores ** 0.5)) if s * s <= cores})
            for t in cands:
                _, dt = ref.poisson_solver(D, np.zeros_like(D), 8, 0.0, threads=t, timed=True)
                if best is None or dt < best[1]:
                    best = (t, dt)
            threads = best[0]
        # non-Poisson stages: a reference design with nthreads=0 runs every stage but its solver does
        # nothing (0 tiles -> max_update 0 -> "converged", src/solver.cpp:73-83,142)
        d = ref.design(osetup, threads=0)
        t0 = time.perf_counter()
        d.initialize_solvers(img)
        t_init = time.perf_counter() - t0
        t0 = time.perf_counter()
        d.transport_iteration()
        t_other = time.perf_counter() - t0
        raster = d.get("raster")
        d.close()
        _, t_solve = ref.poisson_solver(raster, np.zeros_like(raster), sweeps_sample, 0.0, threads=threads, timed=True)
        kind = "reference"
    else:
        port = O.OracleLib()
        threads = 1
        d = port.design(osetup)
        t0 = time.perf_counter()
        d.initialize_solvers(img)
        t_init = time.perf_counter() - t0
        t0 = time.perf_counter()
        d.stage_errors()
        d.stage_raster()
        raster = port.subtract_average(d.get("raster"))
        d.stage_step()
        t_other = time.perf_counter() - t0
        d.close()
        t0 = time.perf_counter()
        port.poisson_lex(raster, np.zeros_like(raster), sweeps_sample, 0.0)
        t_solve = time.perf_counter() - t0
        kind = "port"
    t_sweep = t_solve / sweeps_sample
    t_iter = t_other + n_full * t_sweep
    return {"kind": kind, "cores": int(threads), "t_other_s": t_other, "t_sweep_s": t_sweep, "t_init_s": t_init,
            "t_iter_s": t_iter, "sweeps_sample": sweeps_sample, "sweeps_full": n_full,
            "sweeps_per_s": 1.0 / t_sweep, "gbs": BYTES_PER_CELL_SWEEP * W * H / t_sweep / 1e9,
            "sample": f"{kind} poisson_solver x{sweeps_sample} sweeps on {threads} threads ({t_sweep * 1e3:.2f} ms/sweep) + one "
                      f"full non-Poisson iteration ({t_other:.2f} s), extrapolated to {n_full} sweeps/iteration"}


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    setup, _, desc = make_workload(args.workload)
    threads = None
    times = []
    last = None
    for i in range(args.warmup + args.steps):
        s = cpu_sample(args.workload, sweeps_sample=40, threads=threads)
        threads = s["cores"]
        last = s
        if i >= args.warmup:
            times.append(s["t_iter_s"])
    t = float(np.mean(times))
    value = 1.0 / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "mesh": [setup.mesh_nx, setup.mesh_ny], "domain": [setup.res_x, setup.res_y],
                   "parallelism": f"cpu threads={last['cores']} (reference tile grid)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": last["kind"], "sample": last["sample"],
                         "poisson_sweeps_per_s": last["sweeps_per_s"], "poisson_gbs": last["gbs"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
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
    cd = P.from_setup(setup, device=local_rank)
    cd.initialize_solvers(img)
    hook = None
    if distributed:
        from poisson_caustic_design_b200 import slab
        hook = slab.SlabSolveHook(cd, dist, rank, world, local_rank)
    designs = 1 if distributed else world
    for _ in range(args.warmup):
        cd.perform_transport_iteration()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    cd.solve_totals(reset=True)
    launches0 = P.launch_count()
    dev_ms = 0.0
    steps_vals = []
    barrier()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        cd.flush_l2()                        # untimed: evict L2 between steps (256 MiB memset)
        cd.event_record(0)
        steps_vals.append(cd.perform_transport_iteration())
        cd.event_record(1)
        dev_ms += cd.event_elapsed_ms(0, 1)
    barrier()
    wall = time.perf_counter() - wall0
    launches = P.launch_count() - launches0
    totals = cd.solve_totals()
    clocks = sampler.stop() if sampler else None
    t_max = dev_ms
    if dist is not None:
        t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_max = float(t.item())
    value = designs * args.steps / (t_max * 1e-3)
    if hook is not None and hook.error is not None:
        raise hook.error

    # ---- end to end through host buffers --------------------------------------------------------
    pin = lambda n: torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()  # noqa: E731
    h_tx, h_ty = pin(V), pin(V)
    o_tx, o_ty, o_err, o_vgx, o_vgy = pin(V), pin(V), pin(V), pin(V), pin(V)
    cd.get_into("target_x", h_tx)
    cd.get_into("target_y", h_ty)
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        cd.set_from("target_x", h_tx)        # host-owned mesh (public member in the reference) -> device
        cd.set_from("target_y", h_ty)
        cd.perform_transport_iteration()     # returns the step size (device -> host scalar)
        cd.get_into("target_x", o_tx)
        cd.get_into("target_y", o_ty)
        cd.get_into("errors", o_err)
        cd.get_into("vertex_gradient_x", o_vgx)
        cd.get_into("vertex_gradient_y", o_vgy)
        h_tx[:] = o_tx
        h_ty[:] = o_ty
    barrier()
    e_wall = time.perf_counter() - e0
    if dist is not None:
        t = torch.tensor([e_wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_wall = float(t.item())
    e2e = {"value": designs * args.steps / e_wall, "unit": UNIT, "h2d_bytes_per_step": 2 * V * 8,
           "d2h_bytes_per_step": 5 * V * 8 + 8}

    mode_used = hook.solves[-1]["mode"] if hook is not None and hook.solves else None
    if hook is not None:
        hook.close()
    if rank != 0:
        cd.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    sweeps = totals["sweeps"]
    alg_bytes = BYTES_PER_CELL_SWEEP * W * H * sweeps
    achieved = alg_bytes / (totals["kernel_ms"] * 1e-3) / 1e9 if totals["kernel_ms"] > 0 else 0.0
    if distributed:
        achieved /= world   # per GPU: every rank moves 1/world of the cells
    traffic = load_traffic()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (traffic["dram_bytes_per_launch"] if totals["path"] == "resident" else wave_traffic(traffic, W * H // (world if distributed else 1))) if traffic else None,
                "kernel": {"resident": "sor_resident_kernel", "tiled": "sor_wave_kernel"}.get(totals["path"], "sor_colour_kernel"),
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes / max(totals["launches"], 1) / (world if distributed else 1),
                "sweeps_per_launch": sweeps / max(totals["launches"], 1),
                "kernel_ms_per_launch": totals["kernel_ms"] / max(totals["launches"], 1),
                "kernel_share_of_step": totals["kernel_ms"] / dev_ms if dev_ms > 0 else None,
                "note": ("phi and D live in registers/shared memory for the whole solve: DRAM traffic per launch is one read "
                         "of both fields, far below the algorithmic bytes" if totals["path"] == "resident" else
                         "temporal blocking: each launch applies 2 sweeps per HBM pass (12 B/cell/sweep of real traffic)")}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_max / args.steps, "higher_is_better": True, "scaling": "strong" if distributed else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "mesh": [setup.mesh_nx, setup.mesh_ny], "domain": [W, H],
                   "parallelism": ("1 GPU" if world == 1 else
                                   f"ONE design, Poisson solves on row slabs x{world} ({mode_used} exchange), other stages replicated"
                                   if distributed else f"replicas x{world} (one lens design per GPU, no collective)"),
                   "l2": "flushed between steps (256 MiB memset, untimed)", "solver_path": totals["path"]},
        "poisson_sweeps_per_sec": designs * sweeps / (totals["kernel_ms"] * 1e-3) if totals["kernel_ms"] > 0 else None,
        "poisson_sweeps_per_step": sweeps / args.steps,
        "poisson_gbs": achieved,
        "roofline": roofline,
        "e2e": e2e,
        "gpu_launches": int(launches * world),
        "clocks": clocks,
        "step_sizes": steps_vals[:4],
        "wall_ms_per_step": wall * 1e3 / args.steps,
    }
    if world == 1 and not args.no_cpu:
        try:
            s = cpu_sample(args.workload, sweeps_sample=100)
            line["cpu_baseline"] = {"value": 1.0 / s["t_iter_s"], "unit": UNIT, "cores": s["cores"], "kind": s["kind"],
                                    "sample": s["sample"], "poisson_sweeps_per_s": s["sweeps_per_s"], "poisson_gbs": s["gbs"]}
        except Exception as e:  # the CPU leg must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}
    print(json.dumps(line), flush=True)
    cd.close()
    if dist is not None:
        dist.destroy_process_group()


def run_slab(args, rank: int, local_rank: int, world: int):
    """--workload c5slab: BASELINE.json configs[4] -- ONE 8192x8192 Poisson problem cut into row slabs across the
    N GPUs (strong scaling; ghost rows exchanged over NVLink once per wavefront pass).  One step = 64 sweeps."""
    import torch
    import torch.distributed as dist
    from poisson_caustic_design_b200 import slab
    import poisson_caustic_design_b200 as P

    W = H = 8192
    sweeps_per_step = 64
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    row0, rows = slab.partition(H, world, rank)
    GH = P.lib().pcd_slab_ghost_rows()
    yy = (np.arange(row0 - GH, row0 + rows + GH, dtype=np.float64)[:, None] + 0.5) / H
    xx = (np.arange(W, dtype=np.float64)[None, :] + 0.5) / W
    D = np.cos(3 * np.pi * xx) * np.cos(2 * np.pi * yy) + 0.3 * np.cos(17 * np.pi * xx) * np.cos(11 * np.pi * yy)  # zero mean
    D[np.broadcast_to((yy < 0) | (yy > 1), D.shape)] = 0.0   # ghost rows outside the grid
    eng = slab.CudaSlabEngine(W, H, row0, rows, local_rank)
    eng.upload(D, np.zeros_like(D))

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        slab.solve(eng, dist, rank, world, sweeps_per_step, 0.0, sweeps_per_step)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    sync()
    if sampler:
        sampler.start()
    launches0 = P.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for _ in range(args.steps):
        info = slab.solve(eng, dist, rank, world, sweeps_per_step, 0.0, sweeps_per_step)
    e1.record()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = P.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        t = float(ms.item()) * 1e-3
        sweeps = args.steps * sweeps_per_step
        peak, peak_src = load_peaks()
        achieved = BYTES_PER_CELL_SWEEP * W * H * sweeps / t / 1e9
        line = {"metric": "poisson_sweeps_per_sec", "value": sweeps / t, "unit": "sweeps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t * 1e3 / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "synthetic 8192x8192 Poisson problem, row slabs (BASELINE.json configs[4])",
                           "domain": [W, H], "parallelism": f"row slabs x{world}, " + {"peer": "ghost rows stored into the neighbour's HBM by the pass kernel itself (NVLink peer memory, device-side flags)", "wavefront": "single slab, no exchange" if world == 1 else "ghost rows exchanged with NCCL send/recv once per pass", "colour": "one ghost row exchanged with NCCL per colour phase"}[info["mode"]],
                           "sweeps_per_step": sweeps_per_step, "l2": "working set 1.6 GB >> L2"},
                "roofline": {"bound": "hbm", "achieved": achieved / world, "peak": peak, "unit": "GB/s", "frac": achieved / world / peak,
                             "traffic": wave_traffic(load_traffic(), W * H // world), "kernel": "sor_wave_kernel", "peak_source": peak_src,
                             "note": "per-GPU algorithmic GB/s (24 B/cell/sweep); whole job = achieved x n_gpus"},
                "poisson_gbs": achieved, "gpu_launches": int(launches * world), "clocks": clocks,
                "e2e": {"value": sweeps / t, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8 * sweeps_per_step,
                        "note": "device-resident solve; per step only the per-sweep maxima cross to the host"}}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS) + ["c5slab"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
