#!/usr/bin/env python
"""Every BASELINE.json configuration on the GPU, one JSON line each (profiles/r02_configs.json):
device time per image (CUDA events, inputs resident: the default overlapped launch, and the march /
integration split of a serialised launch), end-to-end time through rtb200_create_image,
ray-segments/s, the step-level SURVEY 8d fraction of the measured FP64 peak, and parity against
the CPU oracle - at full size where the oracle finishes in about a minute on the box's host
threads, else on a strided sample of the reference's own N_start / N_parallel decomposition
(stated in the line).

    python tools/run_configs.py [name ...] > profiles/r02_configs.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402  (the checker)
from raytrace_miniapp_b200 import abi, lib, problem_io, synth  # noqa: E402


def rel_l2(a, b):
    return float(np.linalg.norm(np.ravel(a) - np.ravel(b)) / max(np.linalg.norm(b), 1e-300))


def max_rel(a, b, floor=1e-6):
    a, b = np.ravel(a), np.ravel(b)
    m = np.abs(b) > floor * np.abs(b).max()
    return float(np.max(np.abs(a[m] - b[m]) / np.abs(b[m]))) if m.any() else 0.0


def configs():
    small, _ = problem_io.load_npz(os.path.join(ROOT, "tests", "golden", "ase_small.npz"))
    seed, _ = problem_io.load_npz(os.path.join(ROOT, "tests", "golden", "seed_small.npz"))
    yield "ASE_small", "config 1 (reference fixture)", small, 1
    yield "seed_small", "config 1 (reference fixture)", seed, 1
    yield "ASE_medium-synth", "config 2 (stand-in for the missing ASE_medium.dat)", synth.ase_medium_synth(small), 1
    yield "seed_medium-synth", "config 2 (stand-in for the missing seed_medium.dat)", synth.seed_medium_synth(seed), 1
    yield "S4 (gain 2x/axis, image 2x/axis)", "config 4", synth.s4(small, 2, 2), 1
    yield "S4b (gain 4x/axis, image 2x/axis)", "config 4", synth.s4(small, 4, 2), 1
    yield "S4x (gain 8x/axis: gv 35 MB/plane, image 2x/axis)", "config 4 (lineshape tables beyond L2)", synth.s4(small, 8, 2), 1
    for K, af in ((52, 1), (99, 1), (128, 1), (256, 1), (512, 1), (52, 2), (512, 2), (128, 4)):
        yield ("spectral K=%d, angles x%d" % (K, af), "config 5", synth.spectral(small, K, angle_factor=af),
               1)


def staged_times(ctx, p, flags, torch, n=5):
    """Best device time (CUDA events on the launching stream) of n launches of the staged problem,
    and the library's kernel timings of that launch."""
    e = p.euv_beam
    dev = torch.device("cuda", 0)
    n_pix = ctx.stage(p, flags=flags)
    img = torch.zeros(e.nx * e.ny * e.nv, dtype=torch.float64, device=dev)
    ang = torch.zeros(e.na * e.nb, dtype=torch.float64, device=dev)
    best, tm = 1e9, None
    st = torch.cuda.current_stream().cuda_stream
    for it in range(n + 1):
        img.zero_()
        ang.zero_()
        ctx.reset_timings()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.launch(0, n_pix, img, ang, stream=st)
        e1.record()
        ctx.sync(raise_on_failed=False)
        torch.cuda.synchronize()
        if it > 0 and e0.elapsed_time(e1) < best:
            best, tm = e0.elapsed_time(e1), ctx.timings()
    del img, ang
    return best, tm


def main():
    want = sys.argv[1:]
    ctx = lib.Context(0)
    peak = ctx.measure_fp64_peak()
    O = pyoracle.Oracle()
    threads = os.cpu_count() or 1
    import torch
    for name, cfg, p, stride in configs():
        if want and not any(w in name for w in want):
            continue
        e = p.euv_beam
        flags = abi.FLAG_NO_LIMITS
        t0 = time.perf_counter()
        img, ang = ctx.create_image(p, flags=flags)  # warm-up (allocations)
        best_e2e = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            img, ang = ctx.create_image(p, flags=flags)
            best_e2e = min(best_e2e, time.perf_counter() - t0)
        # device time with the inputs resident (the default, overlapped launch), then the per-kernel
        # split from a context that runs the two kernels strictly one after the other
        dev_ms, _ = staged_times(ctx, p, flags, torch)
        os.environ["RTB200_OVERLAP"] = "0"
        try:
            ctx_s = lib.Context(0)
        finally:
            os.environ.pop("RTB200_OVERLAP", None)
        serial_ms, tm = staged_times(ctx_s, p, flags, torch)
        ctx_s.close()
        # eikonal steps of the march (device counter; a counting instantiation of the kernel)
        os.environ["RTB200_COUNT_STEPS"] = "1"
        try:
            ctx_c = lib.Context(0)
        finally:
            os.environ.pop("RTB200_COUNT_STEPS", None)
        ctx_c.create_image(p, flags=flags)
        march_steps = int(ctx_c.timings()["march_steps"])
        ctx_c.close()
        W_seg = p.n_rays * (p.N - 1) * 3
        line = {"config": name, "baseline_config": cfg, "rays": p.n_rays, "N": p.N, "K": e.nv,
                "gain_grid": [p.gain[1].Nx, p.gain[1].Ny], "ray_segments": W_seg,
                "device_ms": dev_ms, "serialised_ms": serial_ms, "march_ms": tm["march_ms"],
                "integrate_ms": tm["integrate_ms"],
                "march_steps": march_steps, "march_steps_per_ray": march_steps / max(p.n_rays, 1),
                "march_steps_per_s": march_steps / (tm["march_ms"] * 1e-3),
                "e2e_ms": best_e2e * 1e3, "ray_segments_per_s_device": W_seg / (dev_ms * 1e-3),
                "ray_segments_per_s_e2e": W_seg / best_e2e}
        per_upd = 32 if p.seed is None else 25  # FP64 instr per frequency update / per (ray, bin)
        work = (W_seg * e.nv * 32) if p.seed is None else (p.n_rays * e.nv * 25)
        line["fp64_fraction_8d"] = work / (dev_ms * 1e-3) / peak
        line["fp64_convention"] = "%d FP64 instr per %s (SURVEY 8d), measured DFMA peak %.3e lane-instr/s" % (
            per_upd, "frequency update" if p.seed is None else "(ray, bin)", peak)
        # ---- parity -----------------------------------------------------------------------------
        old = p.N_start, p.N_parallel
        try:
            if stride > 1:
                p.N_start, p.N_parallel = 1, stride
                img_s, ang_s = ctx.create_image(p, flags=flags)
            else:
                img_s, ang_s = img, ang
            t0 = time.perf_counter()
            o = O.create_image(p, flags=flags, threads=threads)
            line["oracle_s"] = time.perf_counter() - t0
            line["parity"] = {"against": "CPU oracle (oracle/rt_oracle.c), %s" % (
                                  "full size" if stride == 1 else "every %d-th ray (N_start=1, N_parallel=%d: %d rays)"
                                  % (stride, stride, p.n_rays)),
                              "image_relL2": rel_l2(img_s, o["image"]), "I_ang_relL2": rel_l2(ang_s, o["I_ang"]),
                              "image_max_rel": max_rel(img_s, o["image"]), "I_ang_max_rel": max_rel(ang_s, o["I_ang"]),
                              "failure_code": ctx.failure_code}
        finally:
            p.N_start, p.N_parallel = old
        print(json.dumps(line), flush=True)
        del img, ang
        torch.cuda.empty_cache()
    ctx.close()


if __name__ == "__main__":
    main()
