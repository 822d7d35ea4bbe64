"""torchrun entry: a caustic design whose Poisson solves are spread over WORLD_SIZE GPUs as row slabs (slab.SlabSolveHook),
all other stages replicated.  Prints timing (device events on rank 0's torch stream are not enough here: the stages run
on the context's own stream, so the step is timed with synchronised wall clock, max over ranks) and optionally saves
rank 0's fields.
    python -m torch.distributed.run --nproc-per-node 2 tools/dist_design_run.py --res_w 320 --aspect 4 --iters 2"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poisson_caustic_design_b200 as P  # noqa: E402
from poisson_caustic_design_b200 import slab, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res_w", type=int, default=320)
    ap.add_argument("--aspect", type=float, default=4.0, help="image width / height")
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--height_iters", type=int, default=1)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--mode", default="auto")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = 4 * a.res_w
    H = int(W / a.aspect)
    image = synth.synth_density(W, H, a.seed)
    setup = synth.Setup(a.res_w, W, H)
    cd = P.from_setup(setup, local)
    cd.initialize_solvers(image)
    hook = slab.SlabSolveHook(cd, dist if world > 1 else None, rank, world, local, mode=a.mode)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    steps = []
    sync()
    t0 = time.perf_counter()
    for _ in range(a.iters):
        steps.append(cd.perform_transport_iteration())
    sync()
    t_transport = time.perf_counter() - t0
    if hook.error:
        raise hook.error
    for i in range(a.height_iters):
        cd.perform_height_map_iteration(i)
    sync()
    if hook.error:
        raise hook.error
    if rank == 0:
        sw = [s["sweeps"] for s in hook.solves]
        print(json.dumps({"W": W, "H": H, "gpus": world, "iters": a.iters, "steps": steps, "sweeps": sw, "mode": hook.solves[0]["mode"],
                          "s_per_transport_iter": t_transport / max(a.iters, 1),
                          "solve_ms": [round(s["ms"], 3) for s in hook.solves]}), flush=True)
        if a.out:
            np.savez(a.out, image=image, steps=np.array(steps), phi=cd.get("phi"), h=cd.get("h"), tx=cd.get("target_x"),
                     ty=cd.get("target_y"), sz=cd.get("source_z"), sweeps=np.array(sw))
    hook.close()
    cd.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
