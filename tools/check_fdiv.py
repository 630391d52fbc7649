#!/usr/bin/env python
"""Exhaustive device-side proof for fdiv_refined: all 2^23 divisor significands x all 2^23
numerator significands (operands in [1, 2)), plus random slices at other operand scales.

    python tools/check_fdiv.py [out.txt]        (about a minute on a B200)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytrace_miniapp_b200 import lib  # noqa: E402

ctx = lib.Context(0)
lines = []
t0 = time.time()
bad_total = 0
STEP = 1 << 17
for b0 in range(0, 1 << 23, STEP):
    bad, a, b = ctx.check_fdiv(b0, STEP, 0, 0)
    bad_total += bad
    if bad:
        lines.append("MISMATCH divisor slice %d: %d pairs, e.g. a=%r b=%r" % (b0, bad, a, b))
lines.append("all 2^46 significand pairs, a, b in [1, 2): %d mismatches (%.1f s)" % (bad_total, time.time() - t0))
rng = np.random.default_rng(7)
t0 = time.time()
n = 0
for _ in range(64):
    ea, eb = int(rng.integers(-60, 60)), int(rng.integers(-60, 60))
    b0 = int(rng.integers(0, (1 << 23) - 256))
    bad, a, b = ctx.check_fdiv(b0, 256, ea, eb)
    bad_total += bad
    n += 256 << 23
    if bad:
        lines.append("MISMATCH scale 2^%d / 2^%d: %d pairs, e.g. a=%r b=%r" % (ea, eb, bad, a, b))
lines.append("%d pairs at 64 random operand scales 2^[-60, 60): %d mismatches in total (%.1f s)" % (n, bad_total, time.time() - t0))
text = "\n".join(lines)
print(text)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write("fdiv_refined (csrc/rtb200_math.cuh) vs the correctly rounded quotient, B200, sm_100a\n" + text + "\n")
sys.exit(1 if bad_total else 0)
