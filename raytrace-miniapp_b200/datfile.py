"""Reader / writer for the reference's `.dat` input files (serialized create_image_struct).

File = uint64 payload length + payload (src/CreateImage.cpp:26-58).  Payload layout:
create_image_struct::pack/unpack (src/RayTraceStructures.cpp:2159-2292) nesting
EUV_beam_struct (:441-573), seed_beam_struct (:1028-1240), ray_gain_struct (:1987-2048),
ray_seed_struct (:1393-1431), each optionally preceded by the 16-byte byte_array_header
(src/RayTraceStructures.h:470-482).  Fields that are off the hot path (z, v, physics scalars,
the seed_beam's temporal-shape tail) are kept verbatim in `.extra` so that
write_dat(read_dat(f)) reproduces the file byte for byte; the writer is what emits the
synthetic ASE_medium / scaled inputs for the reference's own CreateImage driver.
"""
import struct

import numpy as np

from .abi import BeamGrid, Gain, Problem, SeedProfile

HEADER_ID = 237


class _Reader:
    def __init__(self, buf, pos=0):
        self.buf, self.pos = buf, pos

    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.buf, self.pos)
        self.pos += struct.calcsize("<" + fmt)
        return v if len(v) > 1 else v[0]

    def arr(self, dtype, n):
        a = np.frombuffer(self.buf, dtype=dtype, count=n, offset=self.pos).copy()
        self.pos += a.nbytes
        return a

    def raw(self, n):
        b = bytes(self.buf[self.pos:self.pos + n])
        self.pos += n
        return b


def _read_header(r):
    """load_byte_header (src/RayTraceStructures.cpp:118-138): returns dict or None (old data)."""
    if r.buf[r.pos] != HEADER_ID:
        return None
    b = r.raw(16)
    if b[1] != 4 or b[2] != 8:
        raise ValueError("byte_array_header: unexpected int/double size")
    return {"version": b[3], "type": b[4], "n_bytes": b[7] * (1 << 32) + struct.unpack("<I", b[8:12])[0],
            "flags": b[12:16]}


def _make_header(version, typ, n_bytes, flags=b"\0\0\0\0"):
    return (bytes([HEADER_ID, 4, 8, version, typ, 0, 0, n_bytes >> 32]) +
            struct.pack("<I", n_bytes & 0xFFFFFFFF) + bytes(flags))


def _parse_euv_beam(buf):
    r = _Reader(buf)
    head = _read_header(r)
    if head is not None and head["version"] > 0 and head["type"] != 2:
        raise ValueError("not euv_beam data")
    run = r.take("???")
    nx, ny, nz, na, nb, nv, nz_sub = r.take("7i")
    R_scale, G_scale, lam, Nc, dx, dy, dz, da, db, v0 = r.take("10d")
    x, y, z = r.arr("<f8", nx), r.arr("<f8", ny), r.arr("<f8", nz)
    a, b = r.arr("<f8", na), r.arr("<f8", nb)
    v, dv = r.arr("<f8", nv), r.arr("<f8", nv)
    if r.pos != len(buf):
        raise ValueError("euv_beam: trailing bytes")
    extra = dict(run=run, nz_sub=nz_sub, R_scale=R_scale, G_scale=G_scale, lam=lam, Nc=Nc, v0=v0,
                 z=z, v=v, header=head is not None)
    return BeamGrid(x, y, a, b, dx, dy, da, db, dv=dv, dz=dz, extra=extra)


def _pack_euv_beam(g):
    e = g.extra
    z = np.asarray(e.get("z", np.zeros(1)), "<f8")
    v = np.asarray(e.get("v", np.zeros(g.nv)), "<f8")
    body = struct.pack("<???", *e.get("run", (True, True, True)))
    body += struct.pack("<7i", g.nx, g.ny, z.size, g.na, g.nb, g.nv, e.get("nz_sub", 0))
    body += struct.pack("<10d", e.get("R_scale", 1.0), e.get("G_scale", 1.0), e.get("lam", 0.0),
                        e.get("Nc", 0.0), g.dx, g.dy, g.dz, g.da, g.db, e.get("v0", 0.0))
    for arr in (g.x, g.y, z, g.a, g.b, v, g.dv):
        body += np.asarray(arr, "<f8").tobytes()
    return _make_header(2, 2, 16 + len(body)) + body


def _parse_seed_beam(buf):
    r = _Reader(buf)
    head = _read_header(r)
    if head is not None and head["version"] > 0 and head["type"] != 3:
        raise ValueError("not seed_beam data")
    nx, ny, na, nb = r.take("4i")
    dx, dy, da, db = r.take("4d")
    phys = r.take("14d")  # Wx..chirp: off the path
    x, y, a, b = r.arr("<f8", nx), r.arr("<f8", ny), r.arr("<f8", na), r.arr("<f8", nb)
    tail = r.raw(len(buf) - r.pos)  # N, tau, use_transform, seed_shape blobs: off the path
    extra = dict(phys=phys, tail=tail, head=head)
    return BeamGrid(x, y, a, b, dx, dy, da, db, extra=extra)


def _pack_seed_beam(g):
    e = g.extra
    head = e.get("head") or {"version": 2, "flags": b"\0\0\0\0"}
    body = struct.pack("<4i", g.nx, g.ny, g.na, g.nb)
    body += struct.pack("<4d", g.dx, g.dy, g.da, g.db)
    body += struct.pack("<14d", *e.get("phys", (0.0,) * 14))
    for arr in (g.x, g.y, g.a, g.b):
        body += np.asarray(arr, "<f8").tobytes()
    body += e.get("tail", struct.pack("<i", 0))
    return _make_header(head["version"], 3, 16 + len(body), head["flags"]) + body


def _parse_gain(buf):
    r = _Reader(buf)
    Nx, Ny, Nv = r.take("3i")
    x, y = r.arr("<f8", Nx), r.arr("<f8", Ny)
    n = r.arr("<f8", Nx * Ny)
    g0, E0 = r.arr("<f4", Nx * Ny), r.arr("<f4", Nx * Ny)
    gv = r.arr("<f4", Nx * Ny * Nv)
    gv0 = r.arr("<f4", Nx * Ny)
    if r.pos != len(buf):
        raise ValueError("ray_gain: size mismatch")
    return Gain(x, y, n, g0, E0, gv.reshape(Ny, Nx, Nv), gv0)


def _pack_gain(g):
    E0 = g.E0 if g.E0 is not None else np.zeros_like(g.g0)
    return (struct.pack("<3i", g.Nx, g.Ny, g.Nv) + g.x.tobytes() + g.y.tobytes() + g.n.tobytes() +
            g.g0.tobytes() + E0.tobytes() + g.gv.tobytes() + g.gv0.tobytes())


def _parse_seed(buf):
    r = _Reader(buf)
    dim = r.take("5i")
    xs, fs = [], []
    for i in range(5):
        xs.append(r.arr("<f8", dim[i]))
        fs.append(r.arr("<f8", dim[i]))
    f0 = r.take("d")
    if r.pos != len(buf):
        raise ValueError("ray_seed: size mismatch")
    return SeedProfile(xs, fs, f0)


def _pack_seed(s):
    out = struct.pack("<5i", *[v.size for v in s.x])
    for i in range(5):
        out += s.x[i].tobytes() + s.f[i].tobytes()
    return out + struct.pack("<d", s.f0)


def parse_payload(buf):
    """create_image_struct::unpack.  Returns (Problem, golden_image|None, golden_I_ang|None)."""
    buf = memoryview(buf)
    r = _Reader(buf)
    N, N_start, N_parallel = r.take("3i")
    dz = r.take("d")
    nb = r.take("I")
    euv = _parse_euv_beam(buf[r.pos:r.pos + nb]) if nb else None
    r.pos += nb
    nb = r.take("I")
    seed_beam = _parse_seed_beam(buf[r.pos:r.pos + nb]) if nb else None
    r.pos += nb
    gain = []
    for _ in range(N):
        nb = r.take("I")
        gain.append(_parse_gain(buf[r.pos:r.pos + nb]))
        r.pos += nb
    nb = r.take("I")
    seed = _parse_seed(buf[r.pos:r.pos + nb]) if nb else None
    r.pos += nb
    image = I_ang = None
    if r.take("?"):
        image = r.arr("<f8", euv.nx * euv.ny * euv.nv)
    if r.take("?"):
        I_ang = r.arr("<f8", euv.na * euv.nb)
    if r.pos != len(buf):
        raise ValueError("create_image_struct: size mismatch")
    p = Problem(euv, gain, seed_beam, seed, N_start, N_parallel)
    p.header_dz = dz
    return p, image, I_ang


def pack_payload(p, image=None, I_ang=None):
    """create_image_struct::pack."""
    out = struct.pack("<3id", p.N, p.N_start, p.N_parallel, p.euv_beam.dz)
    blob = _pack_euv_beam(p.euv_beam)
    out += struct.pack("<I", len(blob)) + blob
    blob = _pack_seed_beam(p.seed_beam) if p.seed_beam is not None else b""
    out += struct.pack("<I", len(blob)) + blob
    for g in p.gain:
        blob = _pack_gain(g)
        out += struct.pack("<I", len(blob)) + blob
    blob = _pack_seed(p.seed) if p.seed is not None else b""
    out += struct.pack("<I", len(blob)) + blob
    out += struct.pack("<?", image is not None)
    if image is not None:
        out += np.asarray(image, "<f8").tobytes()
    out += struct.pack("<?", I_ang is not None)
    if I_ang is not None:
        out += np.asarray(I_ang, "<f8").tobytes()
    return out


def read_dat(path):
    with open(path, "rb") as f:
        data = f.read()
    (n,) = struct.unpack_from("<Q", data, 0)
    if n != len(data) - 8:
        raise ValueError("%s: length prefix %d != payload %d" % (path, n, len(data) - 8))
    return parse_payload(memoryview(data)[8:])


def write_dat(path, p, image=None, I_ang=None):
    payload = pack_payload(p, image, I_ang)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(payload)))
        f.write(payload)
