"""Device-side exactness of the march's branch-free FP32 division (csrc/rtb200_math.cuh:
frcp_refined / fdiv_refined) against the correctly rounded quotient, through rtb200_check_fdiv.
The full 2^46 significand pairs are swept by tools/check_fdiv.py (profiles/r01_fdiv_exhaustive.txt);
here: slices of the divisor range at several operand scales, all 2^23 numerators each."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_refined_division_is_correctly_rounded(ctx):
    rng = np.random.default_rng(2)
    total = 0
    # the ends of the significand range, and random slices; exponent pairs inside the guarded domain
    slices = [(0, 64), ((1 << 23) - 64, 64)] + [(int(b), 32) for b in rng.integers(0, (1 << 23) - 32, 10)]
    scales = [(0, 0), (-60, 59), (59, -60), (-37, -35), (12, -20)]
    for i, (b0, nb) in enumerate(slices):
        ea, eb = scales[i % len(scales)]
        bad, a, b = ctx.check_fdiv(b0, nb, ea, eb)
        assert bad == 0, (bad, a, b, ea, eb)
        total += nb << 23
    assert total > 3e9
    # the packed two-quotient form (FFMA2) the step runs: same slices, both lanes
    for i, (b0, nb) in enumerate(slices):
        ea, eb = scales[(i + 2) % len(scales)]
        bad, a, b = ctx.check_fdiv(b0, nb, ea, eb, variant=3)
        assert bad == 0, ("packed", bad, a, b, ea, eb)
    # the detector itself: without the correction step the quotient is only faithful
    bad, a, b = ctx.check_fdiv(12345, 8, 0, 0, variant=1)
    assert bad > 1000, bad


def test_branch_free_sqrt_and_reciprocal_exhaustive(ctx):
    """normalize_s: fsqrt_refined and the reciprocal after it, every significand, both exponent
    parities, at the centre and at both ends of the admitted range 2^-60 .. 2^60."""
    for eb in (0, -60, 58, -31, 30):
        bad, x, got = ctx.check_fdiv(0, 1 << 23, 0, eb, variant=2)
        assert bad == 0, (eb, x, got)
