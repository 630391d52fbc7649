#!/usr/bin/env python
"""A/B of library builds on the GPU box, device-resident inputs (the bench's `value` path):

    python tools/ab_variants.py [name ...]        name = "tree" or a file stem under .variants/
                                                  name@KEY=VAL,KEY=VAL = the same build with that environment

Each build runs in a process of its own (RTB200_LIB): ASE_medium-synth and ASE_small staged once,
then 3 warm-up + 9 timed launches each (CUDA events on the launching stream, L2 flushed between
launches); prints the median / best device time per image and the march / integration split,
and compares image and I_ang with the FIRST build of the list (bit-identical? relative L2).
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/tmp/ab_variants_ref_%s.npz"


def child(name):
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    from raytrace_miniapp_b200 import lib, problem_io, synth
    small, _ = problem_io.load_npz(os.path.join(ROOT, "tests", "golden", "ase_small.npz"))
    cases = [("ASE_medium-synth", synth.ase_medium_synth(small)), ("ASE_small", small)]
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ctx = lib.Context(0)
    for cname, p in cases:
        e = p.euv_beam
        n_pix = ctx.stage(p)
        img = torch.zeros(e.nx * e.ny * e.nv, dtype=torch.float64, device=dev)
        ang = torch.zeros(e.na * e.nb, dtype=torch.float64, device=dev)
        times, split = [], []
        with torch.cuda.stream(stream):
            for it in range(12):
                img.zero_()
                ang.zero_()
                flush.fill_(it & 1)
                ctx.reset_timings()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ctx.launch(0, n_pix, img, ang, stream=torch.cuda.current_stream().cuda_stream)
                e1.record()
                ctx.sync()
                torch.cuda.synchronize()
                if it >= 3:
                    t = ctx.timings()
                    times.append(e0.elapsed_time(e1))
                    split.append((t["march_ms"], t["integrate_ms"]))
        times_s = sorted(times)
        med = times_s[len(times_s) // 2]
        m = sorted(s[0] for s in split)[len(split) // 2]
        g = sorted(s[1] for s in split)[len(split) // 2]
        h_img, h_ang = img.cpu().numpy(), ang.cpu().numpy()
        ref = REF % cname
        if not os.path.exists(ref):
            np.savez(ref, image=h_img, I_ang=h_ang)
            par = "(reference of this run)"
        else:
            r = np.load(ref)
            par = "image %s relL2 %.2e  I_ang relL2 %.2e" % (
                "bit-identical" if np.array_equal(r["image"], h_img) else "differs",
                np.linalg.norm(h_img - r["image"]) / np.linalg.norm(r["image"]),
                np.linalg.norm(h_ang - r["I_ang"]) / np.linalg.norm(r["I_ang"]))
        print("%-22s %-17s median %.3f best %.3f ms (march %.3f + integrate %.3f) | %s"
              % (name, cname, med, times_s[0], m, g, par), flush=True)
    ctx.close()


def main():
    names = sys.argv[1:] or ["tree"]
    for cname in ("ASE_medium-synth", "ASE_small"):
        if os.path.exists(REF % cname):
            os.remove(REF % cname)
    for name in names:
        env = dict(os.environ)
        base = name
        if "@" in base:  # name@KEY=VAL,KEY=VAL: environment of that run
            base, kv = base.split("@", 1)
            for item in kv.split(","):
                k, v = item.split("=", 1)
                env[k] = v
        if base != "tree":
            env["RTB200_LIB"] = os.path.join(ROOT, ".variants", base + ".so")
        else:
            env.pop("RTB200_LIB", None)
        env["AB_CHILD"] = name
        try:
            subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, timeout=240, check=False)
        except subprocess.TimeoutExpired:
            print("%-22s TIMEOUT (240 s)" % name, flush=True)


if __name__ == "__main__":
    if os.environ.get("AB_CHILD"):
        child(os.environ["AB_CHILD"])
    else:
        main()
