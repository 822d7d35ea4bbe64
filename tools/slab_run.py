"""torchrun entry for the slab solver: solves a W x H Poisson problem split over WORLD_SIZE GPUs, prints timing
(device events, max over ranks) and optionally saves the gathered field.
    python -m torch.distributed.run --nproc-per-node N tools/slab_run.py --W 8192 --H 8192 --sweeps 200"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poisson_caustic_design_b200 import slab  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--W", type=int, default=8192)
    ap.add_argument("--H", type=int, default=8192)
    ap.add_argument("--sweeps", type=int, default=200)
    ap.add_argument("--tol", type=float, default=0.0)
    ap.add_argument("--check_every", type=int, default=64)
    ap.add_argument("--out", default="")
    ap.add_argument("--mode", default="auto", choices=["auto", "peer", "wavefront", "colour"])
    a = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rng = np.random.RandomState(0)
    D = rng.standard_normal((a.H, a.W))
    D -= D.mean()
    phi0 = np.zeros_like(D) if not a.out else rng.standard_normal((a.H, a.W))
    row0, rows = slab.partition(a.H, world, rank)
    eng = slab.CudaSlabEngine(a.W, a.H, row0, rows, local)
    eng.upload(slab.with_ghosts(D, row0, rows, eng.GH), slab.with_ghosts(phi0, row0, rows, eng.GH))
    slab.solve(eng, dist, rank, world, min(8, a.sweeps), 0.0, a.check_every, a.mode)           # warm-up
    eng.upload(slab.with_ghosts(D, row0, rows, eng.GH), slab.with_ghosts(phi0, row0, rows, eng.GH))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    info = slab.solve(eng, dist, rank, world, a.sweeps, a.tol, a.check_every, a.mode)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    own = eng.download()
    if a.out:
        parts = [None] * world
        if world > 1:
            dist.all_gather_object(parts, own)
        else:
            parts = [own]
        if rank == 0:
            np.savez(a.out, D=D, phi0=phi0, phi=np.concatenate(parts, axis=0))
    if rank == 0:
        us = float(ms.item()) * 1e3 / info["sweeps"]
        print(json.dumps({"W": a.W, "H": a.H, "gpus": world, "mode": info["mode"], "sweeps": info["sweeps"], "us_per_sweep": us,
                          "sweeps_per_s": 1e6 / us, "algorithmic_gbs": 24.0 * a.W * a.H / (us * 1e-6) / 1e9}), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
