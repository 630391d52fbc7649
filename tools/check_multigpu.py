#!/usr/bin/env python
"""Run under torchrun (one process per GPU): the image-row sharded result over N GPUs must equal
the single-GPU result — bit for bit for `image` in ASE mode (disjoint pixels, same per-pixel
order), to rounding for I_ang — and a seeded case must match to 1e-12.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/check_multigpu.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytrace_miniapp_b200 import dist as rdist, lib, problem_io, synth  # noqa: E402


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    small, _ = problem_io.load_npz(os.path.join(ROOT, "tests", "golden", "ase_small.npz"))
    seeded, _ = problem_io.load_npz(os.path.join(ROOT, "tests", "golden", "seed_small.npz"))
    seeded.N_parallel = 7  # every 7th ray of seed_small keeps the check short
    for name, p in (("ASE_small", small), ("ASE_medium-synth", synth.ase_medium_synth(small)),
                    ("seed_small/7", seeded)):
        e = p.euv_beam
        ctx = lib.Context(local)
        ctx.stage(p)
        image = torch.zeros(e.nx * e.ny * e.nv, dtype=torch.float64, device=dev)
        I_ang = torch.zeros(e.na * e.nb, dtype=torch.float64, device=dev)
        rdist.sharded_create_image(ctx, p, image, I_ang)  # row-cyclic shares + all_reduce
        ctx.sync()
        tile_i, tile_a = torch.zeros_like(image), torch.zeros_like(I_ang)
        rdist.sharded_create_image(ctx, p, tile_i, tile_a, cyclic=False)  # contiguous tiles + all_gather
        ctx.sync()
        whole_i, whole_a = torch.zeros_like(image), torch.zeros_like(I_ang)
        ctx.launch(0, ctx.staged_pixels, whole_i, whole_a, stream=torch.cuda.current_stream().cuda_stream)
        ctx.sync()
        if p.method == 1:
            same = torch.equal(image, whole_i) and torch.equal(tile_i, whole_i)
        else:
            same = float((image - whole_i).norm() / whole_i.norm()) < 1e-12
        ea = float((I_ang - whole_a).norm() / whole_a.norm())
        flag = torch.tensor([int(same and ea < 1e-12)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("%-18s world=%d image %s, I_ang relL2 %.2e -> %s" %
                  (name, world, "bit-identical" if p.method == 1 and same else ("match" if same else "MISMATCH"),
                   ea, "PASS" if int(flag) else "FAIL"))
        ok = ok and bool(int(flag))
        ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
