"""The C-ABI library loads without a GPU, exports every symbol include/rtb200.h declares, and
refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from raytrace_miniapp_b200 import abi
from conftest import ROOT, has_gpu


def _declared():
    src = open(os.path.join(ROOT, "include", "rtb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rtb200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(rtlib):
    L = rtlib.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), "librtb200.so does not export %s" % n
    assert b"sm_100a" in L.rtb200_version()


def test_struct_layout_matches_header(rtlib):
    # sizes implied by the header on LP64: guards the ctypes mirror in abi.py
    assert C.sizeof(abi.Ray) == 16
    assert C.sizeof(abi.Beam) == 24 + 5 * 8 + 5 * 8
    assert C.sizeof(abi.GainPlane) == 16 + 6 * 8
    assert C.sizeof(abi.Seed) == 24 + 5 * 8 + 5 * 8 + 8
    assert C.sizeof(abi.CProblem) == 16 + 4 * 8
    assert C.sizeof(abi.Timings) == 5 * 4 + 2 * 4 + 4 + 2 * 8


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU behaviour")
def test_fails_loudly_without_a_device(rtlib):
    assert rtlib.device_count() == 0
    with pytest.raises(rtlib.RTB200Error) as e:
        rtlib.Context(0)
    assert e.value.code == abi.ERR_CUDA


def test_marshalled_problem_is_a_view_of_the_problem(ase_small):
    """Problem.marshal(): the rtb200_problem structure built once (what bench.py's end-to-end leg
    passes to every call); same pointers and scalars as a fresh c_struct(), attributes pass through."""
    p, _ = ase_small
    m = p.marshal()
    c0, _keep0 = p.c_struct()
    c1, _keep1 = m.c_struct()
    assert m.c_struct()[0] is c1  # built once
    assert (c1.N, c1.N_start, c1.N_parallel) == (c0.N, c0.N_start, c0.N_parallel) == (p.N, p.N_start, p.N_parallel)
    assert m.euv_beam is p.euv_beam and m.n_rays == p.n_rays and m.method == p.method
    for i in range(p.N):
        assert c1.gain[i].Nx == p.gain[i].Nx and c1.gain[i].Nv == p.euv_beam.nv
