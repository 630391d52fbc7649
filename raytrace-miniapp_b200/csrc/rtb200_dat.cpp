// rtb200_dat.cpp — reader for the reference's serialized create_image_struct (the payload of a
// .dat file), straight into the POD problem description of include/rtb200.h.
//
// Wire format: create_image_struct::pack/unpack (src/RayTraceStructures.cpp:2159-2292) nesting
// EUV_beam_struct (:441-573), seed_beam_struct (:1028-1240), ray_gain_struct (:1987-2048) and
// ray_seed_struct (:1393-1431); sub-blobs may start with the 16-byte byte_array_header
// (src/RayTraceStructures.h:470-482, id 237).  Only the fields on the image-formation path are
// kept; arrays are copied into one owned arena (the stream is not aligned).
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/rtb200.h"

namespace {

struct Cursor {
    const unsigned char *p;
    size_t n, pos;
    bool ok;
    Cursor(const void *b, size_t len) : p((const unsigned char *) b), n(len), pos(0), ok(true) {}
    template <class T>
    T take()
    {
        T v;
        std::memset(&v, 0, sizeof(T));
        if (sizeof(T) > n - pos) { // (pos <= n always; no sum that could wrap)
            ok = false;
            return v;
        }
        std::memcpy(&v, p + pos, sizeof(T));
        pos += sizeof(T);
        return v;
    }
    const unsigned char *bytes(size_t len)
    {
        if (len > n - pos) {
            ok = false;
            return nullptr;
        }
        const unsigned char *r = p + pos;
        pos += len;
        return r;
    }
    // load_byte_header (src/RayTraceStructures.cpp:118-138): skip the header when present.
    int header_type()
    {
        if (pos < n && p[pos] == 237 && n - pos >= 16) {
            const int type = p[pos + 4];
            if (p[pos + 1] != 4 || p[pos + 2] != 8)
                ok = false;
            pos += 16;
            return type;
        }
        return -1;
    }
};

struct Owned {
    rtb200_problem p; // must stay the first member
    rtb200_beam euv, seed_beam;
    rtb200_seed seed;
    std::vector<rtb200_gain_plane> gain;
    std::vector<void *> blocks;
    double *golden_image = nullptr, *golden_I_ang = nullptr;
    ~Owned()
    {
        for (void *b : blocks)
            ::operator delete(b);
    }
    template <class T>
    T *copy(Cursor &c, size_t count)
    {
        if (count > (c.n - c.pos) / sizeof(T)) { // also keeps count * sizeof(T) from wrapping
            c.ok = false;
            return nullptr;
        }
        const unsigned char *src = c.bytes(count * sizeof(T));
        if (!src)
            return nullptr;
        T *dst = (T *) ::operator new(count * sizeof(T) + 1);
        blocks.push_back(dst);
        std::memcpy(dst, src, count * sizeof(T));
        return dst;
    }
};

bool parse_euv(Owned &o, const unsigned char *b, size_t n)
{
    Cursor c(b, n);
    const int type = c.header_type();
    if (type > 0 && type != 2)
        return false;
    c.bytes(3); // run_ASE, run_sat, run_refract
    rtb200_beam &e = o.euv;
    std::memset(&e, 0, sizeof(e));
    e.nx = c.take<int32_t>();
    e.ny = c.take<int32_t>();
    const int nz = c.take<int32_t>();
    e.na = c.take<int32_t>();
    e.nb = c.take<int32_t>();
    e.nv = c.take<int32_t>();
    c.take<int32_t>(); // nz_sub (unused)
    if (!c.ok || e.nx < 1 || e.ny < 1 || nz < 1 || e.na < 1 || e.nb < 1 || e.nv < 1)
        return false;
    c.take<double>(); // R_scale
    c.take<double>(); // G_scale
    c.take<double>(); // lambda
    c.take<double>(); // Nc
    e.dx = c.take<double>();
    e.dy = c.take<double>();
    e.dz = c.take<double>();
    e.da = c.take<double>();
    e.db = c.take<double>();
    c.take<double>(); // v0
    e.x = o.copy<double>(c, e.nx);
    e.y = o.copy<double>(c, e.ny);
    c.bytes(sizeof(double) * (size_t) nz); // z (nz < 2^31: no wrap)
    e.a = o.copy<double>(c, e.na);
    e.b = o.copy<double>(c, e.nb);
    c.bytes(sizeof(double) * (size_t) e.nv); // v
    e.dv = o.copy<double>(c, e.nv);
    return c.ok && c.pos == n;
}

bool parse_seed_beam(Owned &o, const unsigned char *b, size_t n)
{
    Cursor c(b, n);
    const int type = c.header_type();
    if (type > 0 && type != 3)
        return false;
    rtb200_beam &s = o.seed_beam;
    std::memset(&s, 0, sizeof(s));
    s.nx = c.take<int32_t>();
    s.ny = c.take<int32_t>();
    s.na = c.take<int32_t>();
    s.nb = c.take<int32_t>();
    s.dx = c.take<double>();
    s.dy = c.take<double>();
    s.da = c.take<double>();
    s.db = c.take<double>();
    if (!c.ok || s.nx < 1 || s.ny < 1 || s.na < 1 || s.nb < 1)
        return false;
    c.bytes(14 * sizeof(double)); // Wx .. chirp: off the path
    s.x = o.copy<double>(c, s.nx);
    s.y = o.copy<double>(c, s.ny);
    s.a = o.copy<double>(c, s.na);
    s.b = o.copy<double>(c, s.nb);
    return c.ok; // the temporal-shape tail (tau, use_transform, seed_shape) is off the path
}

bool parse_gain(Owned &o, rtb200_gain_plane &g, const unsigned char *b, size_t n)
{
    Cursor c(b, n);
    std::memset(&g, 0, sizeof(g));
    g.Nx = c.take<int32_t>();
    g.Ny = c.take<int32_t>();
    g.Nv = c.take<int32_t>();
    if (!c.ok || g.Nx < 1 || g.Ny < 1 || g.Nv < 1)
        return false;
    const size_t nn = (size_t) g.Nx * g.Ny; // < 2^62
    if (nn > n / sizeof(float) || (size_t) g.Nv > n / sizeof(float) / nn)
        return false; // nn * Nv floats cannot be in this stream (and the product could wrap)
    g.x = o.copy<double>(c, g.Nx);
    g.y = o.copy<double>(c, g.Ny);
    g.n = o.copy<double>(c, nn);
    g.g0 = o.copy<float>(c, nn);
    g.E0 = o.copy<float>(c, nn);
    g.gv = o.copy<float>(c, nn * (size_t) g.Nv);
    c.bytes(sizeof(float) * nn); // gv0: off the path (nn <= n / 4 was checked above)
    return c.ok && c.pos == n;
}

bool parse_seed(Owned &o, const unsigned char *b, size_t n)
{
    Cursor c(b, n);
    rtb200_seed &s = o.seed;
    std::memset(&s, 0, sizeof(s));
    for (int i = 0; i < 5; i++) {
        s.dim[i] = c.take<int32_t>();
        if (!c.ok || s.dim[i] < 1)
            return false;
    }
    for (int i = 0; i < 5; i++) {
        s.x[i] = o.copy<double>(c, s.dim[i]);
        s.f[i] = o.copy<double>(c, s.dim[i]);
    }
    s.f0 = c.take<double>();
    return c.ok && c.pos == n;
}

} // namespace

extern "C" {

int rtb200_parse_dat(const void *bytes, size_t n_bytes, rtb200_problem **problem,
                     const double **golden_image, const double **golden_I_ang)
{
    if (!bytes || !problem)
        return RTB200_ERR_ARG;
    *problem = nullptr;
    Owned *o = new (std::nothrow) Owned;
    if (!o)
        return RTB200_ERR_ARG;
    try { // nothing may propagate through the C boundary (std::bad_alloc from a hostile count)
    std::memset(&o->p, 0, sizeof(o->p));
    Cursor c(bytes, n_bytes);
    bool ok = true;
    o->p.N = c.take<int32_t>();
    o->p.N_start = c.take<int32_t>();
    o->p.N_parallel = c.take<int32_t>();
    c.take<double>(); // dz (duplicated inside euv_beam)
    ok = ok && c.ok && o->p.N >= 1 && (size_t) o->p.N <= n_bytes / 16; // a plane takes > 16 bytes
    if (ok) {
        const uint32_t nb = c.take<uint32_t>();
        const unsigned char *b = c.bytes(nb);
        ok = c.ok && nb > 0 && parse_euv(*o, b, nb);
        if (ok)
            o->p.euv_beam = &o->euv;
    }
    if (ok) {
        const uint32_t nb = c.take<uint32_t>();
        const unsigned char *b = c.bytes(nb);
        ok = c.ok;
        if (ok && nb > 0) {
            ok = parse_seed_beam(*o, b, nb);
            if (ok)
                o->p.seed_beam = &o->seed_beam;
        }
    }
    if (ok) {
        o->gain.resize((size_t) o->p.N);
        for (int i = 0; ok && i < o->p.N; i++) {
            const uint32_t nb = c.take<uint32_t>();
            const unsigned char *b = c.bytes(nb);
            ok = c.ok && parse_gain(*o, o->gain[i], b, nb) && o->gain[i].Nv == o->euv.nv;
        }
        o->p.gain = o->gain.data();
    }
    if (ok) {
        const uint32_t nb = c.take<uint32_t>();
        const unsigned char *b = c.bytes(nb);
        ok = c.ok;
        if (ok && nb > 0) {
            ok = parse_seed(*o, b, nb);
            if (ok)
                o->p.seed = &o->seed;
        }
    }
    if (ok && c.take<unsigned char>()) {
        const size_t nxy = (size_t) o->euv.nx * o->euv.ny; // < 2^62
        ok = nxy <= n_bytes / sizeof(double) && (size_t) o->euv.nv <= n_bytes / sizeof(double) / nxy;
        if (ok)
            o->golden_image = o->copy<double>(c, nxy * (size_t) o->euv.nv);
    }
    if (ok && c.ok && c.take<unsigned char>())
        o->golden_I_ang = o->copy<double>(c, (size_t) o->euv.na * o->euv.nb);
    ok = ok && c.ok && c.pos == n_bytes;
    if (!ok) {
        delete o;
        return RTB200_ERR_FORMAT;
    }
    if (golden_image)
        *golden_image = o->golden_image;
    if (golden_I_ang)
        *golden_I_ang = o->golden_I_ang;
    *problem = &o->p;
    return RTB200_OK;
    } catch (...) {
        delete o;
        return RTB200_ERR_FORMAT;
    }
}

void rtb200_free_problem(rtb200_problem *problem)
{
    if (problem)
        delete reinterpret_cast<Owned *>(problem);
}

} // extern "C"
