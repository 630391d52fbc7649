"""Analysis tool: replay the march of a sample of ASE_medium-synth pixels under different
block-scheduling policies (tools/sim_march_policy.cpp) and print the warp instructions each
would issue.  CPU only."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytrace_miniapp_b200 import abi, problem_io, synth  # noqa: E402


def main():
    src = os.path.join(ROOT, "tools", "sim_march_policy.cpp")
    so = "/tmp/libsimmarch.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-o", so, src], check=True)
    L = C.CDLL(so)
    L.sim_trace.restype = C.c_longlong
    L.sim_trace.argtypes = [C.POINTER(abi.CProblem), C.c_longlong, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p]
    L.sim_policy.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                             C.c_int, C.c_void_p, C.c_void_p]
    small, _ = problem_io.load_npz(os.path.join(ROOT, "tests", "golden", "ase_small.npz"))
    which = sys.argv[1] if len(sys.argv) > 1 else "medium"
    p = synth.ase_medium_synth(small) if which == "medium" else small
    cp, keep = p.c_struct()
    e = p.euv_beam
    AB = e.na * e.nb
    npix = e.nx * e.ny
    step = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    pix = np.arange(0, npix, step)
    ops_all, off_all = [], [0]
    buf = np.zeros(AB * 4000, np.uint8)
    off = np.zeros(AB + 1, np.int64)
    for q in pix:
        n = L.sim_trace(C.byref(cp), int(q) * AB, AB, buf.ctypes.data, buf.size, off.ctypes.data)
        assert n >= 0
        ops_all.append(buf[:n].copy())
        base = off_all[-1]
        off_all.extend((off[1:] + base).tolist())
    ops = np.concatenate(ops_all)
    offs = np.array(off_all, np.int64)
    n_rays = offs.size - 1
    nS = int((ops == ord("S")).sum())
    print("rays %d ops %d: S %.1f I %.1f C %.1f per ray" % (n_rays, ops.size, nS / n_rays,
          (ops == ord("I")).sum() / n_rays, (ops == ord("C")).sum() / n_rays))
    cost = np.array([206.0, 98.0, 228.0, 10.0, 200.0])
    out = np.zeros(8)

    def run(policy, thrC=0, thrI=0, maxwait=1 << 30, per_warp=32 * 8):
        L.sim_policy(ops.ctypes.data, offs.ctypes.data, n_rays, per_warp, 6, policy, thrC, thrI, maxwait,
                     cost.ctypes.data, out.ctypes.data)
        t, eC, eI, eS, lC, lI, lS, ins = out
        return ins, "trips %.0f execC %.2f (%.1f lanes) execI %.2f (%.1f) execS %.2f (%.1f) instr/ray %.0f" % (
            t, eC / t, lC / max(eC, 1), eI / t, lI / max(eI, 1), eS / t, lS / max(eS, 1), ins / n_rays)

    base, s = run(0)
    print("baseline            ", s)
    for thr in (4, 8, 12, 16, 20, 24):
        for mw in (1, 2, 3, 1 << 30):
            ins, s = run(1, thr, 0, mw)
            print("C>=%2d maxwait %-6s %+.1f%% " % (thr, mw if mw < 100 else "inf", 100 * (ins / base - 1)), s)
    for thrC, thrI in ((12, 8), (12, 16), (16, 16), (8, 12)):
        ins, s = run(2, thrC, thrI, 2)
        print("C>=%2d I>=%2d mw 2    %+.1f%% " % (thrC, thrI, 100 * (ins / base - 1)), s)


if __name__ == "__main__":
    main()
