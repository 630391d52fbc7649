// rtb200_dat.cpp — reader for the reference's serialized create_image_struct (the payload of a
// .dat file), straight into the POD problem description of include/rtb200.h.
//
// Wire format: create_image_struct::pack/unpack (src/RayTraceStructures.cpp:2159-2292) nesting
// EUV_beam_struct (:441-573), seed_beam_struct (:1028-1240), ray_gain_struct (:1987-2048) and
// ray_seed_struct (:1393-1431); sub-blobs may start with the 16-byte byte_array_header
// (src/RayTraceStructures.h:470-482, id 237).  Only the fields on the image-formation path are
// kept.  rtb200_parse_dat copies every array into an owned arena (the stream is not aligned);
// the view form used by rtb200_create_image_from_dat copies only the small ones (grids, seed
// tables) and lets the large ones (n, g0, E0, gv of every plane) point INTO the byte stream:
// they are read exactly once, unaligned-safely, by the packer that writes the pinned staging
// blob (rtb200_pack.h), so a .dat goes file buffer -> pinned blob -> device with no copy in
// between.  Also the writer (create_image_struct::pack, :2159-2223) for synthetic inputs.
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
    bool view = false; // large arrays alias the byte stream instead of being copied
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
    // a large array: aliased (possibly misaligned; only ever read through memcpy) in view mode
    template <class T>
    const T *big(Cursor &c, size_t count)
    {
        if (!view)
            return copy<T>(c, count);
        if (count > (c.n - c.pos) / sizeof(T)) {
            c.ok = false;
            return nullptr;
        }
        return reinterpret_cast<const T *>(c.bytes(count * sizeof(T)));
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
    g.n = o.big<double>(c, nn);
    g.g0 = o.big<float>(c, nn);
    g.E0 = o.big<float>(c, nn);
    g.gv = o.big<float>(c, nn * (size_t) g.Nv);
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

static int parse_impl(const void *bytes, size_t n_bytes, bool view, rtb200_problem **problem,
                      const double **golden_image, const double **golden_I_ang)
{
    if (!bytes || !problem)
        return RTB200_ERR_ARG;
    *problem = nullptr;
    Owned *o = new (std::nothrow) Owned;
    if (!o)
        return RTB200_ERR_ARG;
    o->view = view;
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

int rtb200_parse_dat(const void *bytes, size_t n_bytes, rtb200_problem **problem,
                     const double **golden_image, const double **golden_I_ang)
{
    return parse_impl(bytes, n_bytes, false, problem, golden_image, golden_I_ang);
}

int rtb200_parse_dat_view(const void *bytes, size_t n_bytes, rtb200_problem **problem)
{
    return parse_impl(bytes, n_bytes, true, problem, nullptr, nullptr);
}

void rtb200_free_problem(rtb200_problem *problem)
{
    if (problem)
        delete reinterpret_cast<Owned *>(problem);
}

} // extern "C"

// ---- writer: create_image_struct::pack (src/RayTraceStructures.cpp:2159-2223) --------------------
namespace {

struct Sink {
    unsigned char *p;
    size_t cap, pos;
    template <class T>
    void put(const T &v)
    {
        if (p && pos + sizeof(T) <= cap)
            std::memcpy(p + pos, &v, sizeof(T));
        pos += sizeof(T);
    }
    void raw(const void *src, size_t n)
    {
        if (p && pos + n <= cap) {
            if (src)
                std::memcpy(p + pos, src, n);
            else
                std::memset(p + pos, 0, n);
        }
        pos += n;
    }
    // byte_array_header (src/RayTraceStructures.h:470-482; create_byte_header, .cpp:140-180)
    void header(unsigned char version, unsigned char type, uint64_t n_bytes)
    {
        const unsigned char h[8] = { 237, 4, 8, version, type, 0, 0, (unsigned char) (n_bytes >> 32) };
        raw(h, 8);
        put<uint32_t>((uint32_t) (n_bytes & 0xffffffffu));
        const unsigned char flags[4] = { 0, 0, 0, 0 };
        raw(flags, 4);
    }
};

size_t euv_bytes(const rtb200_beam &e) { return 16 + 3 + 7 * 4 + 10 * 8 + 8 * ((size_t) e.nx + e.ny + 1 + e.na + e.nb + 2 * (size_t) e.nv); }
size_t seed_beam_bytes(const rtb200_beam &s) { return 16 + 4 * 4 + 18 * 8 + 8 * ((size_t) s.nx + s.ny + s.na + s.nb) + 4; }
size_t gain_bytes(const rtb200_gain_plane &g)
{
    const size_t nn = (size_t) g.Nx * g.Ny;
    return 3 * 4 + 8 * ((size_t) g.Nx + g.Ny + nn) + 4 * (3 * nn + nn * (size_t) g.Nv);
}
size_t seed_bytes(const rtb200_seed &s)
{
    size_t n = 5 * 4 + 8;
    for (int d = 0; d < 5; d++)
        n += 16 * (size_t) s.dim[d];
    return n;
}

} // namespace

extern "C" {

int rtb200_write_dat(const rtb200_problem *p, const double *golden_image, const double *golden_I_ang,
                     void *out, size_t capacity, size_t *n_bytes)
{
    if (!p || !p->euv_beam || !p->gain || p->N < 1 || !n_bytes)
        return RTB200_ERR_ARG;
    const rtb200_beam &e = *p->euv_beam;
    Sink w{ (unsigned char *) out, out ? capacity : 0, 0 };
    w.put<int32_t>(p->N);
    w.put<int32_t>(p->N_start);
    w.put<int32_t>(p->N_parallel);
    w.put<double>(e.dz);
    // euv_beam (EUV_beam_struct::pack, :441-510).  Fields off the image-formation path get
    // neutral values: run_ASE / run_sat / run_refract = true, one z plane, R_scale = G_scale = 1.
    w.put<uint32_t>((uint32_t) euv_bytes(e));
    w.header(2, 2, euv_bytes(e));
    const unsigned char run[3] = { 1, 1, 1 };
    w.raw(run, 3);
    const int32_t dims[7] = { e.nx, e.ny, 1, e.na, e.nb, e.nv, 0 };
    w.raw(dims, sizeof(dims));
    const double sc[10] = { 1.0, 1.0, 0.0, 0.0, e.dx, e.dy, e.dz, e.da, e.db, 0.0 };
    w.raw(sc, sizeof(sc));
    w.raw(e.x, 8 * (size_t) e.nx);
    w.raw(e.y, 8 * (size_t) e.ny);
    w.raw(nullptr, 8); // z
    w.raw(e.a, 8 * (size_t) e.na);
    w.raw(e.b, 8 * (size_t) e.nb);
    w.raw(nullptr, 8 * (size_t) e.nv); // v
    w.raw(e.dv, 8 * (size_t) e.nv);
    if (p->seed_beam) { // seed_beam_struct::pack (:1028-1140), no temporal shapes
        const rtb200_beam &s = *p->seed_beam;
        w.put<uint32_t>((uint32_t) seed_beam_bytes(s));
        w.header(2, 3, seed_beam_bytes(s));
        const int32_t d4[4] = { s.nx, s.ny, s.na, s.nb };
        w.raw(d4, sizeof(d4));
        const double d[4] = { s.dx, s.dy, s.da, s.db };
        w.raw(d, sizeof(d));
        w.raw(nullptr, 14 * 8); // Wx .. chirp
        w.raw(s.x, 8 * (size_t) s.nx);
        w.raw(s.y, 8 * (size_t) s.ny);
        w.raw(s.a, 8 * (size_t) s.na);
        w.raw(s.b, 8 * (size_t) s.nb);
        w.put<int32_t>(0); // N seed shapes
    } else {
        w.put<uint32_t>(0);
    }
    for (int i = 0; i < p->N; i++) { // ray_gain_struct::pack (:1987-2016)
        const rtb200_gain_plane &g = p->gain[i];
        const size_t nn = (size_t) g.Nx * g.Ny;
        if (gain_bytes(g) > 0xffffffffull)
            return RTB200_ERR_LIMITS; // the format stores sub-blob sizes in 32 bits
        w.put<uint32_t>((uint32_t) gain_bytes(g));
        const int32_t d3[3] = { g.Nx, g.Ny, g.Nv };
        w.raw(d3, sizeof(d3));
        w.raw(g.x, 8 * (size_t) g.Nx);
        w.raw(g.y, 8 * (size_t) g.Ny);
        w.raw(g.n, 8 * nn);
        w.raw(g.g0, 4 * nn);
        w.raw(g.E0, 4 * nn); // NULL -> zeros
        w.raw(g.gv, 4 * nn * (size_t) g.Nv);
        w.raw(nullptr, 4 * nn); // gv0: off the path
    }
    if (p->seed) { // ray_seed_struct::pack (:1393-1411)
        const rtb200_seed &s = *p->seed;
        w.put<uint32_t>((uint32_t) seed_bytes(s));
        w.raw(s.dim, 5 * 4);
        for (int d = 0; d < 5; d++) {
            w.raw(s.x[d], 8 * (size_t) s.dim[d]);
            w.raw(s.f[d], 8 * (size_t) s.dim[d]);
        }
        w.put<double>(s.f0);
    } else {
        w.put<uint32_t>(0);
    }
    w.put<unsigned char>(golden_image ? 1 : 0);
    if (golden_image)
        w.raw(golden_image, 8 * (size_t) e.nx * e.ny * e.nv);
    w.put<unsigned char>(golden_I_ang ? 1 : 0);
    if (golden_I_ang)
        w.raw(golden_I_ang, 8 * (size_t) e.na * e.nb);
    *n_bytes = w.pos;
    if (out && w.pos > capacity)
        return RTB200_ERR_ARG;
    return RTB200_OK;
}

} // extern "C"
