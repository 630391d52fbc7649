"""Wire format: the .dat reader / writer (Python) and rtb200_parse_dat (C++), against the
reference's own files when present and against the committed fixtures otherwise."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

import raytrace_miniapp_b200 as rt
from raytrace_miniapp_b200 import abi, problem_io
from conftest import REF_ROOT

HAVE_REF = os.path.exists(os.path.join(REF_ROOT, "ASE_small.dat"))


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not present")
@pytest.mark.parametrize("name", ["ASE_small", "seed_small"])
def test_roundtrip_is_byte_identical(name):
    path = os.path.join(REF_ROOT, name + ".dat")
    raw = open(path, "rb").read()
    p, img, ang = rt.read_dat(path)
    payload = rt.pack_payload(p, img, ang)
    assert struct.pack("<Q", len(payload)) + payload == raw


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not present")
@pytest.mark.parametrize("name", ["ase_small", "seed_small"])
def test_fixture_matches_dat(name, request):
    p0, extra = request.getfixturevalue(name)
    p, img, ang = rt.read_dat(os.path.join(REF_ROOT, {"ase_small": "ASE_small", "seed_small": "seed_small"}[name] + ".dat"))
    a, b = problem_io.problem_arrays(p0), problem_io.problem_arrays(p)
    assert a.keys() == b.keys()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(extra["dat_golden_image"], img)
    assert np.array_equal(extra["dat_golden_I_ang"], ang)


def test_write_then_read(tmp_path, ase_small):
    p, extra = ase_small
    f = str(tmp_path / "x.dat")
    rt.write_dat(f, p, extra["dat_golden_image"], extra["dat_golden_I_ang"])
    q, img, ang = rt.read_dat(f)
    a, b = problem_io.problem_arrays(p), problem_io.problem_arrays(q)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(img, extra["dat_golden_image"])
    assert q.n_rays == 399000 and q.method == 1 and q.ray_segments == 2394000


def test_truncated_payload_is_rejected(ase_small):
    p, _ = ase_small
    payload = rt.pack_payload(p)
    with pytest.raises(Exception):
        rt.parse_payload(payload[:-5])


@pytest.mark.parametrize("name", ["ase_small", "seed_small"])
def test_c_parser_matches_python(name, request, rtlib):
    """rtb200_parse_dat (pure host code: runs without a GPU)."""
    p, extra = request.getfixturevalue(name)
    payload = rt.pack_payload(p, extra["dat_golden_image"], extra["dat_golden_I_ang"])
    L = rtlib.load()
    cp = C.POINTER(abi.CProblem)()
    gi, ga = abi.c_double_p(), abi.c_double_p()
    buf = C.create_string_buffer(payload, len(payload))
    assert L.rtb200_parse_dat(buf, len(payload), C.byref(cp), C.byref(gi), C.byref(ga)) == abi.OK
    try:
        q = cp.contents
        assert (q.N, q.N_start, q.N_parallel) == (p.N, p.N_start, p.N_parallel)
        e = q.euv_beam.contents
        assert (e.nx, e.ny, e.na, e.nb, e.nv) == (p.euv_beam.nx, p.euv_beam.ny, p.euv_beam.na, p.euv_beam.nb, p.euv_beam.nv)
        assert np.array_equal(np.ctypeslib.as_array(e.dv, (e.nv,)), p.euv_beam.dv)
        assert e.dz == p.euv_beam.dz
        for i in range(p.N):
            g = q.gain[i]
            n = g.Nx * g.Ny
            assert np.array_equal(np.ctypeslib.as_array(g.n, (n,)), p.gain[i].n.ravel())
            assert np.array_equal(np.ctypeslib.as_array(g.gv, (n * g.Nv,)), p.gain[i].gv.ravel())
            assert np.array_equal(np.ctypeslib.as_array(g.E0, (n,)), p.gain[i].E0.ravel())
        assert bool(q.seed) == (p.seed is not None)
        if p.seed is not None:
            s = q.seed.contents
            assert s.f0 == p.seed.f0
            for d in range(5):
                assert np.array_equal(np.ctypeslib.as_array(s.f[d], (s.dim[d],)), p.seed.f[d])
            sb = q.seed_beam.contents
            assert np.array_equal(np.ctypeslib.as_array(sb.a, (sb.na,)), p.seed_beam.a)
        n_img = e.nx * e.ny * e.nv
        assert np.array_equal(np.ctypeslib.as_array(gi, (n_img,)), extra["dat_golden_image"])
    finally:
        L.rtb200_free_problem(cp)
    # malformed input is refused, not crashed on
    bad = C.create_string_buffer(payload[:1000], 1000)
    assert L.rtb200_parse_dat(bad, 1000, C.byref(cp), None, None) == abi.ERR_FORMAT


def _plain(p):
    """The same problem without the off-path extras of the file it came from (the C++ writer
    emits neutral values for them)."""
    def grid(g, euv):
        if g is None:
            return None
        return abi.BeamGrid(g.x, g.y, g.a, g.b, g.dx, g.dy, g.da, g.db, dv=g.dv if euv else None, dz=g.dz)
    gain = [abi.Gain(g.x, g.y, g.n, g.g0, g.E0, g.gv) for g in p.gain]
    return abi.Problem(grid(p.euv_beam, True), gain, grid(p.seed_beam, False), p.seed, p.N_start, p.N_parallel)


@pytest.mark.parametrize("name", ["ase_small", "seed_small"])
def test_cpp_writer_matches_python_writer(name, request, rtlib):
    """rtb200_write_dat (create_image_struct::pack restated in C++) produces the bytes of the
    Python writer, and both parsers read them back."""
    p, extra = request.getfixturevalue(name)
    q = _plain(p)
    want = rt.pack_payload(q, extra["dat_golden_image"], extra["dat_golden_I_ang"])
    got = rtlib.write_dat_payload(q, extra["dat_golden_image"], extra["dat_golden_I_ang"])
    assert got == want
    r, img, ang = rt.parse_payload(got)
    a, b = problem_io.problem_arrays(p), problem_io.problem_arrays(r)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(img, extra["dat_golden_image"])
    # without golden arrays
    assert rtlib.write_dat_payload(q) == rt.pack_payload(q)


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                     "oracle", "_ref", "libref_oracle.so")),
                    reason="oracle/_ref (the reference built from its sources) not present")
def test_reference_loads_what_the_cpp_writer_wrote(tmp_path, ase_small, rtlib):
    """The unmodified reference's loader (create_image_struct::unpack) accepts the C++ writer's
    output and computes the fixture's image from it."""
    from oracle import pyoracle
    p, extra = ase_small
    q = _plain(p)
    q.N_start, q.N_parallel = 7, 61
    payload = rtlib.write_dat_payload(q)
    f = str(tmp_path / "w.dat")
    with open(f, "wb") as fh:
        fh.write(struct.pack("<Q", len(payload)) + payload)
    R = pyoracle.Reference(f)
    img, ang, _ = R.create_image("cpu")
    R.close()
    o = pyoracle.Oracle().create_image(q)
    assert np.array_equal(img, o["image"]) and np.array_equal(ang, o["I_ang"])
    assert np.linalg.norm(img) > 0
