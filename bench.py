#!/usr/bin/env python
"""bench.py — the headline benchmark of the image-formation path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl rtb200|reference]

Metric: ray-segments/s (one ray-segment = one (ray, length segment, sub-segment) triple,
SURVEY.md §8d).  A step is one create_image pass over the workload.
  N = 1  workload = "ASE_medium" of BASELINE.json configs[1].  The real ASE_medium.dat is not in
         the reference checkout (.MISSING_LARGE_BLOBS); the documented synthetic stand-in is
         built from ASE_small (raytrace_miniapp_b200.synth.ase_medium_synth).
  N > 1  one process per GPU (torchrun), image rows sharded across ranks, image tiles gathered
         and I_ang reduced over NCCL.  Weak scaling: ny is refined by N so rays/GPU is fixed.
`value` is timed with inputs resident in HBM (CUDA events on the launching stream, max over
ranks); `e2e` is the same metric through the reference-facing call with host buffers.
--impl reference times the reference's own CPU implementation (oracle/_ref, unmodified
sources, `threads` method on all host cores) on a bounded strided sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from raytrace_miniapp_b200 import problem_io, synth  # noqa: E402

FP64_INSTR_PER_UPDATE = 32  # SURVEY.md §8d convention (ASE mode)
TMP = os.path.join(ROOT, ".bench_tmp")


def workload(n_gpus, scaling):
    small, _ = problem_io.load_npz(os.path.join(ROOT, "tests", "golden", "ase_small.npz"))
    rows = n_gpus if scaling == "weak" else 1
    p = synth.ase_medium_synth(small, rows_factor=rows)
    e = p.euv_beam
    name = ("ASE_medium-synth (ASE_small refined as -scale=8: %dx%dx%dx%d rays, N=%d planes, "
            "nv=%d%s)" % (e.nx, e.ny, e.na, e.nb, p.N, e.nv,
                          ", ny x%d for weak scaling" % rows if rows > 1 else ""))
    return p, name


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def reference_cpu(problem, stride, steps, warmup, threads_method="threads"):
    """Times the reference's CPU path on a strided sample (every `stride`-th ray, the
    reference's own N_start/N_parallel decomposition).  Returns dict or None."""
    from oracle import pyoracle
    from raytrace_miniapp_b200 import write_dat
    seg_per_ray = (problem.N - 1) * 3
    old = problem.N_start, problem.N_parallel
    problem.N_start, problem.N_parallel = 0, stride
    n_rays = problem.n_rays
    try:
        if pyoracle.Reference.available():
            os.makedirs(TMP, exist_ok=True)
            path = os.path.join(TMP, "bench_sample_%d.dat" % os.getpid())
            write_dat(path, problem)
            R = pyoracle.Reference(path)
            cores = R.hardware_threads()
            times = []
            for i in range(warmup + steps):
                _, _, sec = R.create_image(threads_method)
                if i >= warmup:
                    times.append(sec)
            R.close()
            os.remove(path)
            kind = "reference"
        else:
            O = pyoracle.Oracle()
            cores = os.cpu_count() or 1
            times = []
            for i in range(warmup + steps):
                t0 = time.perf_counter()
                O.create_image(problem, threads=cores)
                if i >= warmup:
                    times.append(time.perf_counter() - t0)
            kind = "port"
    finally:
        problem.N_start, problem.N_parallel = old
    sec = sum(times) / len(times)
    return {"value": n_rays * seg_per_ray / sec, "unit": "ray-segments/s", "cores": cores,
            "kind": kind, "seconds_per_pass": sec, "n_rays": n_rays,
            "sample": "every %d-th ray of the workload (N_start=0, N_parallel=%d: %d rays), "
                      "reference method '%s' on %d host threads" % (stride, stride, n_rays,
                                                                     threads_method, cores)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    problem, name = workload(args.gpus, args.scaling)
    stride = args.cpu_stride or 16
    r = reference_cpu(problem, stride, args.steps, args.warmup)
    line = {"impl": "reference", "metric": "ray_segments_per_s", "value": r["value"],
            "unit": "ray-segments/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["seconds_per_pass"] * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "sample": r["sample"]},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "ray-segments/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from raytrace_miniapp_b200 import build as rbuild, dist as rdist, lib as rl
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        rbuild.build_library()  # no-op when librtb200.so is newer than its sources
    else:  # the other ranks wait for local rank 0's (normally instantaneous) build
        t_wait = time.time()
        while rbuild.needs_build() and time.time() - t_wait < 300:
            time.sleep(0.5)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the rtb200 path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    problem, name = workload(world, args.scaling)
    e = problem.euv_beam
    seg_per_ray = (problem.N - 1) * 3
    W_seg = problem.n_rays * seg_per_ray
    W_upd = W_seg * e.nv
    ctx = rl.Context(local)
    n_pix = ctx.stage(problem)  # inputs resident in HBM before the timed region
    image = torch.zeros(e.nx * e.ny * e.nv, dtype=torch.float64, device=dev)
    I_ang = torch.zeros(e.na * e.nb, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.Stream(device=dev)
    K, Wm = args.steps, args.warmup

    def step():
        if world > 1:
            rdist.sharded_create_image(ctx, problem, image, I_ang)
        else:
            image.zero_()
            I_ang.zero_()
            ctx.launch(0, n_pix, image, I_ang, stream=torch.cuda.current_stream().cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    with torch.cuda.stream(stream):
        for _ in range(Wm):
            flush.zero_()
            step()
        ctx.sync()
        barrier()
        ctx.reset_timings()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(K)]
        if rank == 0:
            sampler.start()
        t_wall0 = time.perf_counter()
        for k in range(K):
            flush.zero_()  # L2 flush between timed iterations (outside the per-step events)
            ev[k][0].record()
            step()
            ev[k][1].record()
        ctx.sync()
        barrier()
        t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    tm = ctx.timings()
    ms_total = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms_total, tm["march_ms"], tm["integrate_ms"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, march_ms, integ_ms = [float(v) for v in t.cpu()]
    ms_per_step = ms_total / K
    value = W_seg / (ms_per_step * 1e-3)
    launches_per_step = tm["kernel_launches"] // K
    image_norm = float(torch.linalg.vector_norm(image).cpu())

    # ---- end to end through the reference-facing call: host buffers, H2D + D2H inside -------------
    h_img = torch.empty(e.nx * e.ny * e.nv, dtype=torch.float64).pin_memory()
    h_ang = torch.empty(e.na * e.nb, dtype=torch.float64).pin_memory()
    h2d = d2h = 0
    e2e_times = []
    ctx2 = rl.Context(local)
    for i in range(Wm + K):
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            ctx2.create_image(problem, image=h_img.numpy(), I_ang=h_ang.numpy())
        else:
            with torch.cuda.stream(stream):
                ctx2.stage(problem)
                rdist.sharded_create_image(ctx2, problem, image, I_ang)
                h_img.copy_(image, non_blocking=True)
                h_ang.copy_(I_ang, non_blocking=True)
                ctx2.sync()
        barrier()
        if i >= Wm:
            e2e_times.append(time.perf_counter() - t0)
    gain_bytes = sum(g.x.nbytes + g.y.nbytes + g.n.size * 16 + g.gv.nbytes for g in problem.gain)
    h2d = gain_bytes + 8 * (e.nx + e.ny + e.na + e.nb + e.nv) + 16 * (e.nx + e.ny + e.na + e.nb)
    d2h = h_img.numel() * 8 + h_ang.numel() * 8 + 536
    te = torch.tensor([sum(e2e_times) / len(e2e_times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.cpu()[0])

    if rank == 0:
        fp64_peak = ctx.measure_fp64_peak()  # FP64 lane-instr/s (DFMA), measured on this box
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic = None
        executed_per_update = None
        try:  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["integrate_ase_owner_kernel"]
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
            executed_per_update = tj.get("fp64_instr_per_update_executed")
        except Exception:
            pass
        march_prof = {}
        try:
            march_prof = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["march_flat_kernel"]
        except Exception:
            pass
        launches_integ = max(1, launches_per_step // 2) * K
        integ_s = integ_ms * 1e-3
        upd_local = W_upd / world  # per rank (weak: identical tiles)
        fp64_instr = upd_local * FP64_INSTR_PER_UPDATE * K
        achieved_tflops = fp64_instr * 2 / integ_s / 1e12
        peak_tflops = fp64_peak * 2 / 1e12
        # compulsory HBM bytes of one pass: gain planes read once + image / I_ang written once
        alg_bytes = gain_bytes + image.numel() * 8 / world + I_ang.numel() * 8
        # Roofline of the dominant pipe of the integration kernel.  `achieved` counts the FP64
        # instructions the kernel really issues per frequency update (ncu, profiles/
        # r01_traffic.json); the SURVEY.md 8d convention (the reference formula with the library
        # exp and divide: 32 per update) is kept beside it - by that count the kernel is past 1.0
        # of the peak, which only says that its exp / reciprocal are cheaper than the library's.
        per_upd = executed_per_update if executed_per_update is not None else FP64_INSTR_PER_UPDATE
        roofline = {
            "bound": "fp64", "kernel": "integrate_ase_owner_kernel",
            "achieved": achieved_tflops * per_upd / FP64_INSTR_PER_UPDATE, "peak": peak_tflops,
            "unit": "TFLOP/s",
            "frac": achieved_tflops * per_upd / FP64_INSTR_PER_UPDATE / peak_tflops,
            "traffic": traffic,
            "traffic_note": "dram read+write bytes per launch (ncu, profiles/r01_traffic.json): the "
                            "march->integrate hand-off records, not re-reads of the inputs",
            "convention": "%.2f FP64 instr issued per frequency update (ncu count of this kernel, "
                          "profiles/r01_traffic.json) x 2 flop x updates / kernel time; peak = DFMA "
                          "micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP64 "
                          "entry).  The kernel is bound by instruction issue (72%% of the issue "
                          "slots busy, FP64 pipe 47%%)" % per_upd,
            "survey_convention": {
                "fp64_instr_per_update": FP64_INSTR_PER_UPDATE, "achieved": achieved_tflops,
                "frac": achieved_tflops / peak_tflops, "unit": "TFLOP/s",
                "note": "SURVEY.md 8d fixed convention: the reference formula with library exp and "
                        "divide; above 1.0 because this kernel's exp / reciprocal need fewer FP64 "
                        "instructions than that"},
            "avg_launch_ms": integ_ms / launches_integ,
            "hbm": {"achieved": alg_bytes * K / integ_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": alg_bytes * K / integ_s / 1e9 / hbm_peak,
                    "note": "algorithmic bytes only; the path is FP64/issue-bound, not HBM-bound"}}
        # The march (the larger half of the step) has no pipe roofline: scalar FP32 / mixed FP64
        # per ray with data-dependent trip counts.  Its bound is instruction issue x SIMT
        # efficiency; both factors from the committed ncu capture, the rate from this run.
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        if march_prof.get("warp_instr_per_launch") and sms and args.scaling == "weak":
            issue_peak = sms * 4 * (clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0) * 1e6
            rate = march_prof["warp_instr_per_launch"] * K / (march_ms * 1e-3)
            roofline["march_kernel"] = {
                "kernel": "march_flat_kernel", "bound": "issue", "unit": "warp-instr/s",
                "achieved": rate, "peak": issue_peak, "frac": rate / issue_peak,
                "simt_efficiency": march_prof.get("active_threads_per_instr", 0) / 32.0,
                "avg_launch_ms": march_ms / launches_integ,
                "note": "peak = SMs x 4 schedulers x SM clock; instructions per launch from ncu "
                        "(profiles/r01_traffic.json, N=1 workload; per GPU the same in weak scaling)"}
        line = {
            "metric": "ray_segments_per_s", "value": value, "unit": "ray-segments/s",
            "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "rays": problem.n_rays, "ray_segments": W_seg,
                       "frequency_updates": W_upd, "l2": "256 MiB buffer rewritten between timed steps",
                       "parallelism": "image rows sharded over %d GPU(s), NCCL all_gather(image) + "
                                      "all_reduce(I_ang)" % world if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": W_seg / e2e_s, "unit": "ray-segments/s", "image_time_ms": e2e_s * 1e3,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": launches_per_step * K,
            "image_time_ms_device": ms_per_step,
            "kernel_ms_per_step": {"march": march_ms / K, "integrate": integ_ms / K},
            "image_l2_norm": image_norm,
            "roofline": roofline,
            "wall_s_timed_region": t_wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = reference_cpu(problem, args.cpu_stride or 16, 1, 0)
                line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as ex:  # the baseline is informative; never lose the GPU line
                line["cpu_baseline"] = {"value": None, "unit": "ray-segments/s", "cores": 0,
                                        "kind": "unavailable", "sample": str(ex)}
        print(json.dumps(line))
    ctx.close()
    ctx2.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="rtb200", choices=["rtb200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-stride", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
