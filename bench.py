#!/usr/bin/env python
"""bench.py — the headline benchmark of the image-formation path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl rtb200|reference]

Metric: ray-segments/s (one ray-segment = one (ray, length segment, sub-segment) triple,
SURVEY.md §8d).  A step is one create_image pass over the workload.
  N = 1  workload = "ASE_medium" of BASELINE.json configs[1].  The real ASE_medium.dat is not in
         the reference checkout (.MISSING_LARGE_BLOBS); the documented synthetic stand-in is
         built from ASE_small (raytrace_miniapp_b200.synth.ase_medium_synth).
  N > 1  one process per GPU (torchrun), image rows sharded row-cyclically across ranks, each
         rank's compact rows all-gathered and I_ang reduced over NCCL.  `value`: weak scaling (ny
         is refined by N so rays/GPU is fixed).  `strong`: the FIXED ASE_medium image at N GPUs
         (BASELINE.json: "ASE_medium image time at 1/2/4/8 B200"), and `parity`: that sharded
         image against the single-GPU one of the same run.
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


WORKLOAD = "ase_medium"  # --workload: which BASELINE.json configuration the step traces


def workload(n_gpus, scaling):
    """The headline workload (BASELINE.json configs[1], and configs[2] when sharded), or with
    --workload the synthetic families of configs[3] / configs[4]: s4 / s4b / s4x = gain planes 2 /
    4 / 8 times finer per axis with 4x the image resolution, spectral<K> = K frequency bins with 4x
    the angles.  Weak scaling multiplies the image rows (the sharded axis) by the GPU count."""
    small, _ = problem_io.load_npz(os.path.join(ROOT, "tests", "golden", "ase_small.npz"))
    rows = n_gpus if scaling == "weak" else 1
    tail = ", ny x%d for weak scaling" % rows if rows > 1 else ""
    if WORKLOAD == "ase_medium":
        p = synth.ase_medium_synth(small, rows_factor=rows)
        e = p.euv_beam
        return p, ("ASE_medium-synth (ASE_small refined as -scale=8: %dx%dx%dx%d rays, N=%d planes, "
                   "nv=%d%s)" % (e.nx, e.ny, e.na, e.nb, p.N, e.nv, tail))
    if WORKLOAD in ("s4", "s4b", "s4x"):
        f = {"s4": 2, "s4b": 4, "s4x": 8}[WORKLOAD]
        p = synth.s4(small, f, 2)
        label = "S4 family: ASE_small with gain planes %dx finer per axis (%dx%d nodes), image 2x per axis" % (
            f, p.gain[1].Nx, p.gain[1].Ny)
    elif WORKLOAD.startswith("spectral"):
        K = int(WORKLOAD[len("spectral"):] or 512)
        p = synth.spectral(small, K, angle_factor=2)
        label = "spectral sweep: ASE_small with %d frequency bins, angles 2x per axis" % K
    else:
        raise SystemExit("unknown --workload %r" % WORKLOAD)
    if rows > 1:
        p.euv_beam = synth.refine_rows(p.euv_beam, rows)
    e = p.euv_beam
    return p, "%s: %dx%dx%dx%d rays, N=%d planes, nv=%d%s" % (label, e.nx, e.ny, e.na, e.nb, p.N, e.nv, tail)


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


def lib_hash():
    """Hash of the product library's sources + build flags (see build.source_hash)."""
    from raytrace_miniapp_b200 import build as rbuild
    return rbuild.source_hash()


def ncu_figures():
    """Static figures of the committed ncu captures (profiles/r02_traffic.json), or {}."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        return {}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from raytrace_miniapp_b200 import abi, build as rbuild, dist as rdist, lib as rl
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

    stream = torch.cuda.Stream(device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    K, Wm = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    class Job:
        """One problem resident on this rank's device, traced whole (world 1) or as this rank's
        rows of a row-cyclic decomposition followed by the exchange."""

        def __init__(self, problem, ctx, sharded):
            self.p, self.ctx, self.sharded = problem, ctx, sharded and world > 1
            e = problem.euv_beam
            self.n_pix = ctx.stage(problem)  # inputs resident in HBM before any timed region
            self.info = ctx.staged_info()
            self.image = torch.zeros(e.nx * e.ny * e.nv, dtype=torch.float64, device=dev)
            self.I_ang = torch.zeros(e.na * e.nb, dtype=torch.float64, device=dev)
            self.rows = rdist.RowGather(self.info, world, dev) if self.sharded and self.info["owner"] else None

        def step(self):
            if self.sharded:
                rdist.sharded_create_image(self.ctx, self.p, self.image, self.I_ang, rows=self.rows)
            else:
                self.image.zero_()
                self.I_ang.zero_()
                self.ctx.launch(0, self.n_pix, self.image, self.I_ang,
                                stream=torch.cuda.current_stream().cuda_stream)

        def exchange_name(self):
            if not self.sharded:
                return "single GPU"
            if self.rows is not None:
                return ("image rows sharded row-cyclically over %d GPUs; NCCL all_gather of each rank's "
                        "compact rows (1/%d of the image) + un-permute kernel, all_reduce(sum) of I_ang"
                        % (world, world))
            return ("image rows sharded row-cyclically over %d GPUs; NCCL all_reduce(sum) of the "
                    "full-size partial images and of I_ang" % world)

        def timed(self, steps, warmup):
            """Device time per step (CUDA events on the launching stream, L2 flushed between
            steps, barrier + synchronize on both sides), max over ranks; kernel times likewise."""
            with torch.cuda.stream(stream):
                for _ in range(warmup):
                    flush.zero_()
                    self.step()
                self.ctx.sync()
                barrier()
                self.ctx.reset_timings()
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                      for _ in range(steps)]
                t0 = time.perf_counter()
                for k in range(steps):
                    flush.zero_()  # L2 flush between timed iterations (outside the per-step events)
                    ev[k][0].record()
                    self.step()
                    ev[k][1].record()
                self.ctx.sync()
                barrier()
                wall = time.perf_counter() - t0
            tm = self.ctx.timings()
            ms = sum(a.elapsed_time(b) for a, b in ev)
            ms, march, integ = max_over_ranks([ms, tm["march_ms"], tm["integrate_ms"]])
            return {"ms_per_step": ms / steps, "march_ms": march / steps, "integrate_ms": integ / steps,
                    "launches": tm["kernel_launches"], "wall_s": wall}

        def e2e(self, ctx2, steps, warmup, h_img, h_ang):
            """The same step through the host-facing path: host arrays in (pack + H2D inside the
            timed region), host image / I_ang out (D2H inside), wall clock, max over ranks."""
            times = []
            # the rtb200_problem structure is built once, as an application that holds its
            # create_image_struct would: the timed call is the C entry point with host pointers
            mp = self.p.marshal()
            img_np, ang_np = h_img.numpy(), h_ang.numpy()
            for i in range(warmup + steps):
                barrier()
                t0 = time.perf_counter()
                if not self.sharded:
                    ctx2.create_image(mp, image=img_np, I_ang=ang_np)
                else:
                    with torch.cuda.stream(stream):
                        ctx2.stage(mp, flags=abi.FLAG_LAZY_TABLES)  # every rank stages its own copy
                        rdist.sharded_create_image(ctx2, self.p, self.image, self.I_ang, rows=self.rows)
                        if rank == 0:  # the caller's buffers live in one process: one download
                            h_img.copy_(self.image, non_blocking=True)
                            h_ang.copy_(self.I_ang, non_blocking=True)
                        ctx2.sync()
                barrier()
                if i >= warmup:
                    times.append(time.perf_counter() - t0)
            return max_over_ranks([sum(times) / len(times)])[0]

    ctx, ctx2 = rl.Context(local), rl.Context(local)
    problem, name = workload(world, args.scaling)
    e = problem.euv_beam
    seg_per_ray = (problem.N - 1) * 3
    W_seg = problem.n_rays * seg_per_ray
    W_upd = W_seg * e.nv
    job = Job(problem, ctx, sharded=True)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    main = job.timed(K, Wm)
    clocks = sampler.stop() if rank == 0 else None
    # The timed step launches the integration behind the march with programmatic stream
    # serialization (the march's tail overlaps the integration's start), so CUDA events cannot
    # split it.  The per-kernel times come from a few extra steps, OUTSIDE the timed region, of a
    # context that runs the two kernels strictly one after the other (RTB200_OVERLAP=0).
    saved = os.environ.get("RTB200_OVERLAP")
    os.environ["RTB200_OVERLAP"] = "0"
    try:
        split = Job(problem, rl.Context(local), sharded=True).timed(max(3, min(K, 5)), 2)
    finally:
        if saved is None:
            os.environ.pop("RTB200_OVERLAP", None)
        else:
            os.environ["RTB200_OVERLAP"] = saved
    overlapped = main["integrate_ms"] == 0.0
    main["march_ms"], main["integrate_ms"] = split["march_ms"], split["integrate_ms"]
    main["serialised_ms_per_step"] = split["ms_per_step"]
    ms_per_step = main["ms_per_step"]
    value = W_seg / (ms_per_step * 1e-3)
    image_norm = float(torch.linalg.vector_norm(job.image).cpu())

    h_img = torch.empty(e.nx * e.ny * e.nv, dtype=torch.float64).pin_memory()
    h_ang = torch.empty(e.na * e.nb, dtype=torch.float64).pin_memory()
    e2e_s = job.e2e(ctx2, K, Wm, h_img, h_ang)
    gain_bytes = sum(g.x.nbytes + g.y.nbytes + g.n.size * 16 + g.gv.nbytes for g in problem.gain)
    h2d = gain_bytes + 8 * (e.nx + e.ny + e.na + e.nb + e.nv) + 16 * (e.nx + e.ny + e.na + e.nb)
    d2h = h_img.numel() * 8 + h_ang.numel() * 8 + 536

    # ---- BASELINE.json's second metric: the FIXED ASE_medium image at `world` GPUs (strong
    # scaling), and the parity of the sharded result with the single-GPU one ----------------------
    strong = parity = None
    if world > 1:
        fixed, fixed_name = workload(1, "strong")
        ef = fixed.euv_beam
        ks, ws = max(3, min(K, 10)), max(3, min(Wm, 3))
        sj = Job(fixed, rl.Context(local), sharded=True)
        st = sj.timed(ks, ws)
        hf_img = torch.empty(ef.nx * ef.ny * ef.nv, dtype=torch.float64).pin_memory()
        hf_ang = torch.empty(ef.na * ef.nb, dtype=torch.float64).pin_memory()
        st_e2e = sj.e2e(rl.Context(local), ks, ws, hf_img, hf_ang)
        sharded_img, sharded_ang = sj.image.clone(), sj.I_ang.clone()
        single = {"ms_per_step": 0.0}
        same_bits, ang_err, e2e1 = True, 0.0, 0.0
        if rank == 0:  # the single-GPU reference of the same problem, on rank 0 alone
            old_world = world
            oj = Job(fixed, rl.Context(local), sharded=False)
            with torch.cuda.stream(stream):
                for _ in range(ws):
                    flush.zero_()
                    oj.step()
                oj.ctx.sync()
                torch.cuda.synchronize()
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ks)]
                for k in range(ks):
                    flush.zero_()
                    ev[k][0].record()
                    oj.step()
                    ev[k][1].record()
                oj.ctx.sync()
                torch.cuda.synchronize()
            single["ms_per_step"] = sum(a.elapsed_time(b) for a, b in ev) / ks
            c1 = rl.Context(local)
            t_e = []
            fm, hi_np, ha_np = fixed.marshal(), hf_img.numpy(), hf_ang.numpy()
            for i in range(ws + ks):
                t0 = time.perf_counter()
                c1.create_image(fm, image=hi_np, I_ang=ha_np)
                if i >= ws:
                    t_e.append(time.perf_counter() - t0)
            e2e1 = sum(t_e) / len(t_e)
            same_bits = bool(torch.equal(sharded_img, oj.image))
            ang_err = float((torch.linalg.vector_norm(sharded_ang - oj.I_ang) /
                             torch.linalg.vector_norm(oj.I_ang)).cpu())
            assert old_world == world
        barrier()
        strong = {"workload": fixed_name, "image_time_ms_device": st["ms_per_step"],
                  "image_time_ms_e2e": st_e2e * 1e3,
                  "single_gpu_image_time_ms_device": single["ms_per_step"],
                  "single_gpu_image_time_ms_e2e": e2e1 * 1e3,
                  "speedup_vs_1": single["ms_per_step"] / st["ms_per_step"] if rank == 0 else None,
                  "speedup_vs_1_e2e": e2e1 / st_e2e if rank == 0 else None,
                  "kernel_ms_per_step": ({"march_and_integrate_overlapped": st["march_ms"]}
                                         if st["integrate_ms"] == 0.0 else
                                         {"march": st["march_ms"], "integrate": st["integrate_ms"]}),
                  "steps": ks, "parallelism": sj.exchange_name()}
        parity = {"checked": "sharded image / I_ang of the fixed image (strong.workload) at %d GPUs against the "
                             "single-GPU result computed in this run on rank 0" % world,
                  "image_bit_identical_to_single_gpu": same_bits, "I_ang_relL2": ang_err}

    if rank == 0:
        fp64_peak = ctx.measure_fp64_peak()  # FP64 lane-instr/s (DFMA), measured on this box
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        prof = ncu_figures()
        prof_ok = bool(prof) and prof.get("src_sha16") == lib_hash()
        step_s = ms_per_step * 1e-3
        upd_local = W_upd / world  # per rank (weak: identical shares)
        peak_tflops = fp64_peak * 2 / 1e12
        # W_fp64 = frequency updates x 32 FP64 instructions (the reference formula with
        # the library exp and divide), against the measured DFMA peak, over the WHOLE step: march +
        # integration.  That is the fraction the north_star's 0.60 target is about.
        step_tflops = upd_local * FP64_INSTR_PER_UPDATE * 2 / step_s / 1e12
        integ_tflops = upd_local * FP64_INSTR_PER_UPDATE * 2 / (main["integrate_ms"] * 1e-3) / 1e12
        alg_bytes = gain_bytes + job.image.numel() * 8 / world + job.I_ang.numel() * 8
        dominant = "march_flat_kernel" if main["march_ms"] >= main["integrate_ms"] else "integrate_ase_owner_kernel"
        traffic = None
        if prof_ok:
            traffic = sum(prof[k]["dram_bytes_read"] + prof[k]["dram_bytes_write"]
                          for k in ("march_flat_kernel", "integrate_ase_owner_kernel") if k in prof)
        roofline = {
            "bound": "fp64", "kernel": dominant,
            "achieved": step_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
            "frac": step_tflops / peak_tflops,
            "traffic": traffic,
            "convention": "SURVEY.md 8d at step level: frequency updates x 32 FP64 instr x 2 flop / "
                          "device time of the whole step (march + integration); peak = DFMA "
                          "micro-benchmark of this run (%.3e FP64 lane-instr/s; MEASURED_PEAKS.json has "
                          "no FP64 entry)" % fp64_peak,
            "traffic_note": "dram read+write bytes per step of both kernels (ncu --set full of these very "
                            "sources, profiles/r02_traffic.json): the march->integrate hand-off "
                            "records, not re-reads of the inputs" if prof_ok else
                            "null: no ncu capture of these sources is committed (source hash mismatch)",
            "fp64_peak_lane_instr_per_s": fp64_peak,
            "per_kernel": {
                "march_flat_kernel": {"ms": main["march_ms"], "share_of_step": main["march_ms"] / main["serialised_ms_per_step"],
                                      "bound": "instruction issue x SIMT efficiency (FP32 / XU scalar "
                                               "work per ray, no pipe roofline)"},
                "integrate_ase_owner_kernel": {
                    "ms": main["integrate_ms"], "share_of_step": main["integrate_ms"] / main["serialised_ms_per_step"],
                    "survey_fraction": integ_tflops / peak_tflops,
                    "note": "8d convention for this kernel alone; its own exp / reciprocal need fewer "
                            "FP64 instructions than the library's 32, so this can pass 1.0"}},
            "hbm": {"achieved": alg_bytes / step_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": alg_bytes / step_s / 1e9 / hbm_peak,
                    "note": "algorithmic bytes (gain planes in, image / I_ang out) over the step; the "
                            "path is instruction-issue / FP64 bound, not HBM-bound"}}
        if prof_ok:  # pipe-busy figures of the same build (static: measured under ncu, not in this run)
            roofline["ncu_same_build"] = {k: prof[k] for k in prof if k not in ("src_sha16", "source")}
        launches_per_step = main["launches"] // K
        if job.sharded and job.rows is not None:
            launches_per_step += 1  # the un-permute kernel
        line = {
            "metric": "ray_segments_per_s", "value": value, "unit": "ray-segments/s",
            "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "rays": problem.n_rays, "ray_segments": W_seg,
                       "frequency_updates": W_upd, "l2": "256 MiB buffer rewritten between timed steps",
                       "parallelism": job.exchange_name()},
            "clocks": clocks,
            "e2e": {"value": W_seg / e2e_s, "unit": "ray-segments/s", "image_time_ms": e2e_s * 1e3,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": launches_per_step * K,
            "image_time_ms_device": ms_per_step,
            "kernel_ms_per_step": {"march": main["march_ms"], "integrate": main["integrate_ms"],
                                   "serialised_step": main["serialised_ms_per_step"],
                                   "overlapped_in_timed_step": overlapped,
                                   "note": "per-kernel times of extra steps with the two kernels strictly one "
                                           "after the other (RTB200_OVERLAP=0), outside the timed region; the "
                                           "timed step starts the integration while the march drains"},
            "image_l2_norm": image_norm,
            "src_sha16": lib_hash(),
            "roofline": roofline,
            "wall_s_timed_region": main["wall_s"],
        }
        if strong is not None:
            line["strong"] = strong
            line["parity"] = parity
        elif args.scaling == "weak":  # N = 1: the fixed image IS the workload
            line["strong"] = {"workload": name, "image_time_ms_device": ms_per_step,
                              "image_time_ms_e2e": e2e_s * 1e3, "speedup_vs_1": 1.0, "speedup_vs_1_e2e": 1.0}
        if world == 1:
            # The seeded half of BASELINE.json's config 2 (seed_small.dat: 7 803 000 rays, gain-only
            # integration binned by the exit ray), timed the same way as the headline; parity
            # against the unmodified reference's own CPU output (tests/golden, tools/make_golden.py).
            try:
                from raytrace_miniapp_b200 import problem_io
                sp, extra = problem_io.load_npz(os.path.join(ROOT, "tests", "golden", "seed_small.npz"))
                sjob = Job(sp, rl.Context(local), sharded=False)
                ks = max(3, min(K, 10))
                stm = sjob.timed(ks, 3)
                s_seg = sp.n_rays * (sp.N - 1) * 3
                ref_img = torch.from_numpy(extra["ref_cpu_image"]).to(dev).flatten()
                ref_ang = torch.from_numpy(extra["ref_cpu_I_ang"]).to(dev).flatten()
                rel = lambda a, b: float((torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b)).cpu())  # noqa: E731
                line["seeded"] = {
                    "workload": "seed_small.dat (seed beam %dx%dx%dx%d = %d rays, N=%d planes, nv=%d)" % (
                        sp.seed_beam.nx, sp.seed_beam.ny, sp.seed_beam.na, sp.seed_beam.nb, sp.n_rays, sp.N,
                        sp.euv_beam.nv),
                    "rays": sp.n_rays, "ray_segments": s_seg, "steps": ks,
                    "image_time_ms_device": stm["ms_per_step"],
                    "kernel_ms_per_step": {"march": stm["march_ms"], "integrate": stm["integrate_ms"]},
                    "value": s_seg / (stm["ms_per_step"] * 1e-3), "unit": "ray-segments/s",
                    "parity": {"checked": "image / I_ang against the reference's own CPU result (tests/golden)",
                               "image_relL2": rel(sjob.image, ref_img), "I_ang_relL2": rel(sjob.I_ang, ref_ang)}}
                sjob.ctx.close()
            except Exception as ex:  # informative; never lose the headline line
                line["seeded"] = {"error": str(ex)[-300:]}
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
    ap.add_argument("--workload", default="ase_medium",
                    help="ase_medium (default, the headline), s4 | s4b | s4x (config 4), spectral<K> (config 5)")
    ap.add_argument("--cpu-stride", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = args.workload
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
