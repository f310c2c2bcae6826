// fdes_b200 -- EMD (HDF5) writer and reader without libhdf5.
//
// What the reference does with libhdf5 (src/rwHdf5.cu): writeHdf5 lays out the groups /data
// (potential_slices, exit_wave, images), /microscope (+ aberrations), /sample, /imaging, /user,
// /comments with float32 / int32 / uint8 / fixed-string attributes and contiguous datasets
// (:27-1084, :1085-1945); readHdf5 reads the same names back (:1946-2571).  This file produces and
// parses the bytes libhdf5 1.8 emits for exactly those calls -- the subset of the HDF5 file format
// that the reference's shipped ExampleSpecimens/Au_cubeoctahedron_emd/Auparticle.emd uses:
//   superblock version 0 (8-byte offsets and lengths, group leaf K 4, internal K 16),
//   version-1 object headers, one symbol-table message per group (v1 B-tree node + local heap +
//   one symbol node), dataspace v1, datatype v1, fill-value v2, layout v3 (contiguous),
//   attribute v1.  Numeric attributes are 1-element simple dataspaces (writeAttributeSingle),
//   string attributes scalar with size = strlen (writeAttributeString, :2641-2655).
// tests/h5min.py (an independent pure-Python reader, checked against that real libhdf5 file) parses
// what this writes; read_emd below parses both.
#include "emd.h"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>

namespace fdes {
namespace {

constexpr uint64_t UNDEF = ~0ull;
enum DT { DT_F32, DT_I32, DT_U8, DT_STR };
using Bytes = std::vector<uint8_t>;

void put(Bytes& b, uint64_t v, int n) { for (int i = 0; i < n; i++) b.push_back((uint8_t)(v >> (8 * i))); }
void pad8(Bytes& b) { while (b.size() % 8) b.push_back(0); }
void append(Bytes& b, const Bytes& s) { b.insert(b.end(), s.begin(), s.end()); }

Bytes datatype_msg(DT dt, uint32_t strsize)
{
    Bytes b;
    switch (dt) {
        case DT_F32: b = {0x11, 0x20, 0x1f, 0x00, 4, 0, 0, 0, 0, 0, 0x20, 0, 0x17, 0x08, 0x00, 0x17, 0x7f, 0, 0, 0}; break;
        case DT_I32: b = {0x10, 0x08, 0x00, 0x00, 4, 0, 0, 0, 0, 0, 0x20, 0}; break;
        case DT_U8: b = {0x10, 0x00, 0x00, 0x00, 1, 0, 0, 0, 0, 0, 0x08, 0}; break;
        case DT_STR: b = {0x13, 0x00, 0x00, 0x00}; put(b, strsize, 4); break;
    }
    return b;
}

Bytes dataspace_msg(const std::vector<uint64_t>& dims, bool scalar)
{
    Bytes b;
    if (scalar) { b = {1, 0, 0, 0, 0, 0, 0, 0}; return b; }
    b = {1, (uint8_t)dims.size(), 1, 0, 0, 0, 0, 0};
    for (uint64_t d : dims) put(b, d, 8);
    for (uint64_t d : dims) put(b, d, 8);   // maximum dimensions = current
    return b;
}

struct Attr { std::string name; DT dt; bool scalar; uint32_t strsize; Bytes data; };

struct Obj {
    std::string name;
    bool group = true;
    std::vector<Attr> attrs;
    std::vector<std::unique_ptr<Obj>> kids;
    DT dt = DT_F32;
    uint32_t strsize = 0;
    std::vector<uint64_t> dims;
    const void* ext = nullptr;   // caller-owned data
    Bytes own;                   // data owned by the tree
    uint64_t nbytes = 0;
    uint64_t ohdr = 0, btree = 0, heap = 0;

    Obj* add_group(const std::string& n)
    {
        kids.emplace_back(new Obj);
        kids.back()->name = n;
        return kids.back().get();
    }
    Obj* add_dataset(const std::string& n, DT t, std::vector<uint64_t> d, Bytes data, uint32_t ssize = 0)
    {
        Obj* o = add_group(n);
        o->group = false; o->dt = t; o->dims = std::move(d); o->own = std::move(data); o->strsize = ssize;
        o->nbytes = o->own.size();
        return o;
    }
    Obj* add_f32(const std::string& n, std::vector<uint64_t> d, const std::vector<float>& v)
    {
        Bytes data(v.size() * 4);
        if (!v.empty()) memcpy(data.data(), v.data(), data.size());
        return add_dataset(n, DT_F32, std::move(d), std::move(data));
    }
    void attr_f32(const std::string& n, float v) { Bytes d(4); memcpy(d.data(), &v, 4); attrs.push_back({n, DT_F32, false, 0, d}); }
    void attr_i32(const std::string& n, int v) { Bytes d(4); memcpy(d.data(), &v, 4); attrs.push_back({n, DT_I32, false, 0, d}); }
    void attr_u8(const std::string& n, uint8_t v) { attrs.push_back({n, DT_U8, false, 0, Bytes{v}}); }
    void attr_str(const std::string& n, const std::string& s)
    {
        Bytes d(s.begin(), s.end());
        if (d.empty()) d.push_back(0);   // H5Tset_size(0) is not a valid type; keep one NUL
        attrs.push_back({n, DT_STR, true, (uint32_t)d.size(), d});
    }
};

Bytes message(uint16_t type, uint8_t flags, Bytes body)
{
    pad8(body);
    Bytes b;
    put(b, type, 2); put(b, body.size(), 2); b.push_back(flags); put(b, 0, 3);
    append(b, body);
    return b;
}

Bytes attr_message(const Attr& a)
{
    Bytes dt = datatype_msg(a.dt, a.strsize), ds = dataspace_msg({1}, a.scalar), body;
    body = {1, 0};
    put(body, a.name.size() + 1, 2); put(body, dt.size(), 2); put(body, ds.size(), 2);
    body.insert(body.end(), a.name.begin(), a.name.end()); body.push_back(0); pad8(body);
    append(body, dt); pad8(body);
    append(body, ds); pad8(body);
    append(body, a.data);
    if (body.size() + 8 > 0xfff8) throw std::runtime_error("EMD writer: attribute " + a.name + " too large");
    return message(0x000c, 0, body);
}

struct Writer {
    uint64_t end = 0;
    struct Block { uint64_t addr; Bytes bytes; const void* ext; uint64_t n; };
    std::vector<Block> blocks;
    uint64_t alloc(uint64_t n) { const uint64_t a = (end + 7) & ~7ull; end = a + n; return a; }
    void place(uint64_t addr, Bytes b) { const uint64_t n = b.size(); blocks.push_back({addr, std::move(b), nullptr, n}); }

    void object_header(uint64_t addr, const std::vector<Bytes>& msgs)
    {
        Bytes h = {1, 0};
        size_t total = 0;
        for (const Bytes& m : msgs) total += m.size();
        put(h, msgs.size(), 2); put(h, 1, 4); put(h, total, 4); put(h, 0, 4);
        for (const Bytes& m : msgs) append(h, m);
        place(addr, std::move(h));
    }
    static size_t header_size(const std::vector<Bytes>& msgs)
    {
        size_t total = 16;
        for (const Bytes& m : msgs) total += m.size();
        return total;
    }

    void emit(Obj& o)
    {
        for (auto& k : o.kids) emit(*k);   // children first: their addresses go into this group's symbol node
        std::vector<Bytes> msgs;
        if (!o.group) {
            const uint64_t daddr = o.nbytes ? alloc(o.nbytes) : UNDEF;
            if (o.nbytes) blocks.push_back({daddr, Bytes(), o.ext ? o.ext : o.own.data(), o.nbytes});
            msgs.push_back(message(0x0001, 0, dataspace_msg(o.dims, false)));
            msgs.push_back(message(0x0003, 1, datatype_msg(o.dt, o.strsize)));
            msgs.push_back(message(0x0005, 1, Bytes{2, 2, 2, 1, 0, 0, 0, 0}));
            Bytes lay = {3, 1};
            put(lay, daddr, 8); put(lay, o.nbytes, 8);
            msgs.push_back(message(0x0008, 0, lay));
            for (const Attr& a : o.attrs) msgs.push_back(attr_message(a));
            o.ohdr = alloc(header_size(msgs));
            object_header(o.ohdr, msgs);
            return;
        }
        if (o.kids.size() > 8) throw std::runtime_error("EMD writer: more than 8 links in group " + o.name);
        // local heap: "" at offset 0, the link names, one free block at the end
        Bytes seg(8, 0);
        std::vector<std::pair<std::string, size_t>> order;   // (name, index) sorted by name
        std::vector<uint64_t> name_off(o.kids.size());
        for (size_t i = 0; i < o.kids.size(); i++) {
            name_off[i] = seg.size();
            seg.insert(seg.end(), o.kids[i]->name.begin(), o.kids[i]->name.end());
            seg.push_back(0); pad8(seg);
            order.emplace_back(o.kids[i]->name, i);
        }
        std::sort(order.begin(), order.end());
        const uint64_t free_off = seg.size();
        put(seg, 1, 8); put(seg, 16, 8);
        o.btree = alloc(24 + 33 * 8 + 32 * 8);
        o.heap = alloc(32 + seg.size());
        const uint64_t snod = o.kids.empty() ? UNDEF : alloc(8 + 8 * 40);
        Bytes heap = {'H', 'E', 'A', 'P', 0, 0, 0, 0};
        put(heap, seg.size(), 8); put(heap, free_off, 8); put(heap, o.heap + 32, 8);
        append(heap, seg);
        place(o.heap, std::move(heap));
        Bytes tree = {'T', 'R', 'E', 'E', 0, 0};
        put(tree, o.kids.empty() ? 0 : 1, 2); put(tree, UNDEF, 8); put(tree, UNDEF, 8);
        put(tree, 0, 8);
        if (!o.kids.empty()) { put(tree, snod, 8); put(tree, name_off[order.back().second], 8); }
        tree.resize(24 + 33 * 8 + 32 * 8, 0);
        place(o.btree, std::move(tree));
        if (!o.kids.empty()) {
            Bytes sn = {'S', 'N', 'O', 'D', 1, 0};
            put(sn, o.kids.size(), 2);
            for (const auto& e : order) {
                const Obj& k = *o.kids[e.second];
                put(sn, name_off[e.second], 8); put(sn, k.ohdr, 8);
                put(sn, k.group ? 1 : 0, 4); put(sn, 0, 4);
                put(sn, k.group ? k.btree : 0, 8); put(sn, k.group ? k.heap : 0, 8);
            }
            sn.resize(8 + 8 * 40, 0);
            place(snod, std::move(sn));
        }
        Bytes st;
        put(st, o.btree, 8); put(st, o.heap, 8);
        msgs.push_back(message(0x0011, 0, st));
        for (const Attr& a : o.attrs) msgs.push_back(attr_message(a));
        o.ohdr = alloc(header_size(msgs));
        object_header(o.ohdr, msgs);
    }

    bool write(const char* file, Obj& root)
    {
        alloc(96);   // superblock
        emit(root);
        const uint64_t eof = (end + 7) & ~7ull;
        Bytes sb = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n', 0, 0, 0, 0, 0, 8, 8, 0};
        put(sb, 4, 2); put(sb, 16, 2); put(sb, 0, 4);
        put(sb, 0, 8); put(sb, UNDEF, 8); put(sb, eof, 8); put(sb, UNDEF, 8);
        put(sb, 0, 8); put(sb, root.ohdr, 8); put(sb, 1, 4); put(sb, 0, 4); put(sb, root.btree, 8); put(sb, root.heap, 8);
        place(0, std::move(sb));
        std::sort(blocks.begin(), blocks.end(), [](const Block& a, const Block& b) { return a.addr < b.addr; });
        FILE* f = fopen(file, "wb");
        if (!f) { fprintf(stderr, "  Cannot open %s for writing\n", file); return false; }
        uint64_t pos = 0;
        static const uint8_t zeros[8] = {0};
        bool ok = true;
        for (const Block& b : blocks) {
            while (pos < b.addr) { const size_t n = (size_t)std::min<uint64_t>(8, b.addr - pos); ok &= fwrite(zeros, 1, n, f) == n; pos += n; }
            const void* src = b.ext ? b.ext : (const void*)b.bytes.data();
            ok &= fwrite(src, 1, b.n, f) == b.n;
            pos += b.n;
        }
        while (pos < eof) { ok &= fwrite(zeros, 1, 1, f) == 1; pos++; }
        ok &= fclose(f) == 0;
        return ok;
    }
};

std::vector<float> centred_axis(int m)
{
    std::vector<float> v((size_t)m);
    for (int i = 0; i < m; i++) v[i] = (float)(i - (m - 1) / 2.0);   // src/rwHdf5.cu:104-107
    return v;
}
std::vector<float> index_axis(int n)
{
    std::vector<float> v((size_t)n);
    for (int i = 0; i < n; i++) v[i] = (float)i;
    return v;
}
void axis(Obj* grp, const char* dim, const char* name, std::vector<float> v)
{
    const uint64_t n = v.size();
    Obj* d = grp->add_f32(dim, {n}, v);
    d->attr_str("name", name);
    d->attr_str("units", "[m]");
}
// [k][j][i][c] (i fastest of the spatial indices) -> [i][j][k][c]: the reference's transposition
// (src/rwHdf5.cu:78-88, 248-256, 413-419)
Bytes transposed(const float* src, int n1, int n2, int n3, int comps)
{
    Bytes out((size_t)n1 * n2 * n3 * comps * 4);
    float* dst = reinterpret_cast<float*>(out.data());
    constexpr int TILE = 32;   // (i, j) tiles: both the reads along i and the writes along j stay in cache
    for (int k = 0; k < n3; k++)
        for (int j0 = 0; j0 < n2; j0 += TILE)
            for (int i0 = 0; i0 < n1; i0 += TILE) {
                const int j1 = std::min(n2, j0 + TILE), i1 = std::min(n1, i0 + TILE);
                for (int i = i0; i < i1; i++)
                    for (int j = j0; j < j1; j++) {
                        const float* s = src + (((size_t)k * n2 + j) * n1 + i) * comps;
                        float* d = dst + (((size_t)i * n2 + j) * n3 + k) * comps;
                        for (int c = 0; c < comps; c++) d[c] = s[c];
                    }
            }
    return out;
}
void complex_axis(Obj* grp)
{
    Bytes names = {'r', 'e', 'a', 'l', 'i', 'm', 'a', 'g'};   // S4 x 2 (src/rwHdf5.cu:191-204)
    Obj* d = grp->add_dataset("dim4", DT_STR, {2}, names, 4);
    d->attr_str("name", "complex");
    d->attr_str("units", "[]");
}

}  // namespace

bool write_emd(const char* file, const Params& p, const Atoms& atoms, const float* image, const float* potential,
               int pot_slices, const float* exitwave)
{
    Obj root;
    root.attr_f32("version", 0.1f);   // src/FDES.cu:52
    Obj* data = root.add_group("data");
    if (potential && pot_slices > 0) {
        Obj* g = data->add_group("potential_slices");
        g->attr_u8("emd_group_type", 1);
        g->add_dataset("data", DT_F32, {(uint64_t)p.m1, (uint64_t)p.m2, (uint64_t)pot_slices, 2},
                       transposed(potential, p.m1, p.m2, pot_slices, 2));
        axis(g, "dim1", "x", centred_axis(p.m1));
        axis(g, "dim2", "y", centred_axis(p.m2));
        axis(g, "dim3", "z", centred_axis(pot_slices));
        complex_axis(g);
    }
    if (exitwave) {
        Obj* g = data->add_group("exit_wave");
        g->attr_u8("emd_group_type", 1);
        g->add_dataset("data", DT_F32, {(uint64_t)p.m1, (uint64_t)p.m2, (uint64_t)p.n3, 2},
                       transposed(exitwave, p.m1, p.m2, p.n3, 2));
        axis(g, "dim1", "x", centred_axis(p.m1));
        axis(g, "dim2", "y", centred_axis(p.m2));
        axis(g, "dim3", "z", index_axis(p.n3));
        complex_axis(g);
    }
    if (image) {
        Obj* g = data->add_group("images");
        g->attr_u8("emd_group_type", 1);
        g->add_dataset("data", DT_F32, {(uint64_t)p.n1, (uint64_t)p.n2, (uint64_t)p.n3}, transposed(image, p.n1, p.n2, p.n3, 1));
        axis(g, "dim1", "x", centred_axis(p.n1));
        axis(g, "dim2", "y", centred_axis(p.n2));
        axis(g, "dim3", "z", index_axis(p.n3));
    }
    Obj* mic = root.add_group("microscope");
    mic->attr_f32("voltage", p.E0); mic->attr_str("voltage_units", "[v]");
    mic->attr_f32("gamma", p.gamma);
    mic->attr_f32("wavelength", p.lambda); mic->attr_str("wavelength_units", "[m]");
    mic->attr_f32("interaction_constant", p.sigma); mic->attr_str("interaction_constant_units", "[V^-1][m^-1]");
    mic->attr_f32("focus_spread", p.defocspread); mic->attr_str("focus_spread_units", "[m]");
    mic->attr_f32("illumination_angle", p.illangle); mic->attr_str("illumination_angle_units", "[rad]");
    mic->attr_f32("objective_aperture", p.ObjAp); mic->attr_str("objective_aperture_units", "[rad]");
    mic->attr_f32("mtf_a", p.mtfa); mic->attr_f32("mtf_b", p.mtfb); mic->attr_f32("mtf_c", p.mtfc); mic->attr_f32("mtf_d", p.mtfd);
    Obj* ab = mic->add_group("aberrations");
    ab->attr_str("amplitude_units", "[m]");
    ab->attr_str("angle_units", "[rad]");
    for (int a = 0; a < AB_COUNT; a++) {
        ab->attr_f32(std::string(kAberrationNames[a]) + "_amplitude", p.ab0[a]);
        if (a != AB_C1 && a != AB_C3 && a != AB_C5) ab->attr_f32(std::string(kAberrationNames[a]) + "_angle", p.ab1[a]);
    }
    Obj* sam = root.add_group("sample");
    sam->attr_str("name", p.sample_name);
    sam->attr_str("material", p.material);
    const uint64_t nAt = (uint64_t)atoms.size();
    {
        Bytes z(nAt * 4);
        if (nAt) memcpy(z.data(), atoms.Z.data(), z.size());
        sam->add_dataset("atomic_numbers", DT_I32, {nAt}, z);
        std::vector<float> c(nAt);
        const char* names[3] = {"x_coordinates", "y_coordinates", "z_coordinates"};
        for (int k = 0; k < 3; k++) {
            for (uint64_t i = 0; i < nAt; i++) c[i] = atoms.xyz[3 * i + k];
            sam->add_f32(names[k], {nAt}, c)->attr_str("units", "[m]");
        }
        sam->add_f32("debeye_waller_factors", {nAt}, atoms.dwf)->attr_str("units", "[m^2]");
        sam->add_f32("occupancy", {nAt}, atoms.occ);
    }
    sam->attr_f32("absorptive_potential_factor", p.imPot);
    Obj* im = root.add_group("imaging");
    im->attr_i32("mode", p.mode);
    im->attr_i32("sample_size_x", p.m1); im->attr_i32("sample_size_y", p.m2); im->attr_i32("sample_size_z", p.m3);
    im->attr_str("sample_size_units", "[pix]");
    im->attr_f32("pixel_size_x", p.d1); im->attr_f32("pixel_size_y", p.d2); im->attr_f32("pixel_size_z", p.d3);
    im->attr_str("pixel_size_units", "[m]");
    im->attr_i32("image_size_x", p.n1); im->attr_i32("image_size_y", p.n2); im->attr_i32("image_size_z", p.n3);
    im->attr_str("image_size_units", "[pix]");
    im->attr_i32("border_size_x", p.dn1); im->attr_i32("border_size_y", p.dn2);
    im->attr_str("border_size_units", "[pix]");
    im->attr_f32("specimen_tilt_offset_x", p.tilt_off[0]); im->attr_f32("specimen_tilt_offset_y", p.tilt_off[1]);
    im->attr_f32("specimen_tilt_offset_z", p.tilt_off[2]);
    im->attr_str("specimen_tilt_offset_units", "[rad]");
    im->attr_i32("frozen_phonons", p.frPh);
    im->attr_f32("pixel_dose", p.pD);
    im->attr_f32("subpixel_size_z", p.subSlTh);
    im->attr_str("subpixel_size_units", "[m]");
    {
        const uint64_t n3 = (uint64_t)p.n3;
        std::vector<float> a(n3), b(n3);
        for (uint64_t i = 0; i < n3; i++) { a[i] = p.tiltspec[2 * i]; b[i] = p.tiltspec[2 * i + 1]; }
        im->add_f32("specimen_tilt_x", {n3}, a)->attr_str("units", "[rad]");
        im->add_f32("specimen_tilt_y", {n3}, b)->attr_str("units", "[rad]");
        for (uint64_t i = 0; i < n3; i++) { a[i] = p.tiltbeam[2 * i]; b[i] = p.tiltbeam[2 * i + 1]; }
        im->add_f32("beam_tilt_x", {n3}, a)->attr_str("units", "[rad]");
        im->add_f32("beam_tilt_y", {n3}, b)->attr_str("units", "[rad]");
        im->add_f32("defoci", {n3}, p.defoci)->attr_str("units", "[rad]");   // "[rad]" as in the reference (:1015)
    }
    Obj* user = root.add_group("user");
    user->attr_str("name", p.user_name);
    user->attr_str("institution", p.institution);
    user->attr_str("department", p.department);
    user->attr_str("email", p.email);
    root.add_group("comments")->attr_str("comment", p.comments);
    Writer w;
    return w.write(file, root);
}

// ------------------------------------------------------------------------------------------------
// reader
// ------------------------------------------------------------------------------------------------
namespace {

struct H5In {
    Bytes b;
    uint64_t root = 0;

    uint64_t u(uint64_t pos, int n) const
    {
        // overflow-safe: an undefined address (0xFFFF...FFFF) stored in the file must not wrap the sum
        if (pos > b.size() || (uint64_t)n > b.size() - pos) throw std::runtime_error("EMD reader: read beyond end of file");
        uint64_t v = 0;
        for (int i = 0; i < n; i++) v |= (uint64_t)b[pos + i] << (8 * i);
        return v;
    }
    void open(const char* file)
    {
        FILE* f = fopen(file, "rb");
        if (!f) throw std::runtime_error(std::string("cannot open ") + file);
        fseek(f, 0, SEEK_END);
        const long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        b.resize((size_t)n);
        const size_t got = fread(b.data(), 1, b.size(), f);
        fclose(f);
        static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        if (got != b.size() || b.size() < 96 || memcmp(b.data(), sig, 8) != 0)
            throw std::runtime_error(std::string(file) + " is not an HDF5 file");
        if (b[8] != 0 || b[13] != 8 || b[14] != 8)
            throw std::runtime_error("EMD reader: only HDF5 superblock version 0 with 8-byte offsets is supported");
        root = u(64, 8);
    }
    struct Msg { uint16_t type; uint64_t pos; uint16_t size; };
    std::vector<Msg> messages(uint64_t ohdr) const
    {
        if (u(ohdr, 1) != 1) throw std::runtime_error("EMD reader: only version-1 object headers are supported");
        const size_t nmsg = (size_t)u(ohdr + 2, 2);
        std::vector<std::pair<uint64_t, uint64_t>> chunks = {{ohdr + 16, u(ohdr + 8, 4)}};
        std::vector<Msg> out;
        for (size_t c = 0; c < chunks.size() && out.size() < nmsg; c++) {
            uint64_t pos = chunks[c].first;
            const uint64_t end = pos + chunks[c].second;
            while (pos + 8 <= end && out.size() < nmsg) {
                const Msg m = {(uint16_t)u(pos, 2), pos + 8, (uint16_t)u(pos + 2, 2)};
                if (m.type == 0x10) chunks.emplace_back(u(m.pos, 8), u(m.pos + 8, 8));
                out.push_back(m);
                pos = m.pos + m.size;
            }
        }
        return out;
    }
    std::string heap_name(uint64_t heap, uint64_t off) const
    {
        if (memcmp(&b.at(heap), "HEAP", 4) != 0) throw std::runtime_error("EMD reader: bad local heap");
        uint64_t s = u(heap + 24, 8) + off;
        std::string r;
        while (s < b.size() && b[s]) r.push_back((char)b[s++]);
        return r;
    }
    void tree_links(uint64_t tree, uint64_t heap, std::vector<std::pair<std::string, uint64_t>>& out, int depth = 0) const
    {
        // a crafted file may contain cycles or absurdly deep trees: bound the recursion
        if (depth > 32) throw std::runtime_error("EMD reader: group B-tree nested too deeply");
        u(tree, 8);      // bounds check of the signature + header start
        if (memcmp(&b[tree], "TREE", 4) != 0 || u(tree + 4, 1) != 0) throw std::runtime_error("EMD reader: bad group B-tree");
        const int level = (int)u(tree + 5, 1), used = (int)u(tree + 6, 2);
        for (int i = 0; i < used; i++) {
            const uint64_t child = u(tree + 24 + 16 * i + 8, 8);
            if (level > 0) { tree_links(child, heap, out, depth + 1); continue; }
            u(child, 8);
            if (memcmp(&b[child], "SNOD", 4) != 0) throw std::runtime_error("EMD reader: bad symbol node");
            const int nsym = (int)u(child + 6, 2);
            for (int j = 0; j < nsym; j++)
                out.emplace_back(heap_name(heap, u(child + 8 + 40 * j, 8)), u(child + 8 + 40 * j + 8, 8));
        }
    }
    uint64_t child(uint64_t ohdr, const std::string& name) const
    {
        for (const Msg& m : messages(ohdr))
            if (m.type == 0x11) {
                std::vector<std::pair<std::string, uint64_t>> links;
                tree_links(u(m.pos, 8), u(m.pos + 8, 8), links);
                for (auto& l : links) if (l.first == name) return l.second;
            }
        return UNDEF;
    }
    uint64_t path(const std::string& p) const
    {
        uint64_t o = root;
        size_t s = 0;
        while (s < p.size() && o != UNDEF) {
            const size_t e = p.find('/', s);
            const std::string part = p.substr(s, e == std::string::npos ? std::string::npos : e - s);
            if (!part.empty()) o = child(o, part);
            if (e == std::string::npos) break;
            s = e + 1;
        }
        return o;
    }
    struct Type { int cls; uint32_t size; bool sign; uint64_t len; };
    Type datatype(uint64_t pos) const
    {
        Type t;
        t.cls = (int)(u(pos, 1) & 15);
        t.size = (uint32_t)u(pos + 4, 4);
        t.sign = (u(pos + 1, 1) & 8) != 0;
        if (u(pos + 1, 1) & 1) throw std::runtime_error("EMD reader: big-endian data is not supported");
        t.len = t.cls == 0 ? 12 : t.cls == 1 ? 20 : 8;
        if (t.cls != 0 && t.cls != 1 && t.cls != 3) throw std::runtime_error("EMD reader: unsupported datatype class");
        return t;
    }
    uint64_t space_count(uint64_t pos, uint64_t* dim0 = nullptr) const
    {
        const int ver = (int)u(pos, 1), rank = (int)u(pos + 1, 1);
        const uint64_t d = pos + (ver == 1 ? 8 : 4);
        uint64_t n = 1;
        for (int i = 0; i < rank; i++) n *= u(d + 8 * i, 8);
        if (dim0) *dim0 = rank ? u(d, 8) : 1;
        return n;
    }
    double number(const Type& t, uint64_t pos) const
    {
        if (t.cls == 1 && t.size == 4) { float v; memcpy(&v, &b.at(pos), 4); return v; }
        if (t.cls == 1 && t.size == 8) { double v; memcpy(&v, &b.at(pos), 8); return v; }
        if (t.cls == 0) {
            const uint64_t raw = u(pos, (int)t.size);
            if (t.sign && t.size == 8) return (double)(int64_t)raw;
            if (t.sign && t.size < 8 && (raw >> (8 * t.size - 1))) return (double)((int64_t)raw - ((int64_t)1 << (8 * t.size)));
            return (double)raw;
        }
        throw std::runtime_error("EMD reader: not a number");
    }
    // attribute message -> (type, element count, data position)
    bool attribute(uint64_t ohdr, const std::string& name, Type& t, uint64_t& count, uint64_t& data) const
    {
        for (const Msg& m : messages(ohdr)) {
            if (m.type != 0x0c) continue;
            const int ver = (int)u(m.pos, 1);
            if (ver < 1 || ver > 3) throw std::runtime_error("EMD reader: unsupported attribute message version");
            auto pad = [&](uint64_t n) { return ver == 1 ? (n + 7) & ~7ull : n; };
            const uint64_t nsz = u(m.pos + 2, 2), tsz = u(m.pos + 4, 2), ssz = u(m.pos + 6, 2);
            uint64_t p = m.pos + 8 + (ver == 3 ? 1 : 0);
            u(p, 1);                                           // p inside the file
            const uint64_t room = std::min<uint64_t>(nsz, b.size() - p);   // the name may not run past its field or the file
            if (std::string((const char*)&b[p], strnlen((const char*)&b[p], room)) != name) continue;
            p += pad(nsz);
            t = datatype(p);
            p += pad(tsz);
            count = space_count(p);
            data = p + pad(ssz);
            return true;
        }
        return false;
    }
    template <typename T> bool attr_num(uint64_t ohdr, const std::string& name, T& out) const
    {
        Type t; uint64_t n, d;
        if (ohdr == UNDEF || !attribute(ohdr, name, t, n, d) || n < 1 || t.cls == 3) return false;
        out = (T)number(t, d);
        return true;
    }
    bool attr_str(uint64_t ohdr, const std::string& name, std::string& out) const
    {
        Type t; uint64_t n, d;
        if (ohdr == UNDEF) return false;
        // names / comments are optional: a string this reader cannot decode (e.g. h5py's variable-length
        // strings) leaves the default in place instead of failing the whole file
        try { if (!attribute(ohdr, name, t, n, d) || t.cls != 3) return false; }
        catch (const std::runtime_error&) { return false; }
        out.assign((const char*)&b.at(d), strnlen((const char*)&b.at(d), t.size));
        return true;
    }
    template <typename T> bool dataset(uint64_t ohdr, std::vector<T>& out) const
    {
        if (ohdr == UNDEF) return false;
        Type t{}; uint64_t n = 0, addr = UNDEF, bytes = 0;
        bool have_t = false, have_s = false, have_l = false;
        for (const Msg& m : messages(ohdr)) {
            if (m.type == 3) { t = datatype(m.pos); have_t = true; }
            if (m.type == 1) { n = space_count(m.pos); have_s = true; }
            if (m.type == 8) {
                if (u(m.pos, 1) != 3 || u(m.pos + 1, 1) != 1) throw std::runtime_error("EMD reader: only contiguous datasets (layout v3) are supported");
                addr = u(m.pos + 2, 8); bytes = u(m.pos + 10, 8); have_l = true;
            }
        }
        if (!have_t || !have_s || !have_l || t.cls == 3) return false;
        if (t.size == 0 || n > b.size() / t.size || bytes != n * t.size)     // n * size cannot overflow; data must fit the file
            throw std::runtime_error("EMD reader: dataset size mismatch");
        if (addr != UNDEF) u(addr, 0), u(addr + bytes - (bytes ? 1 : 0), bytes ? 1 : 0);
        out.resize(n);
        for (uint64_t i = 0; i < n; i++) out[i] = addr == UNDEF ? T(0) : (T)number(t, addr + i * t.size);
        return true;
    }
};

}  // namespace

bool read_emd(const char* file, Params& p, Atoms* atoms, bool atoms_from_external)
{
    H5In h;
    h.open(file);
    p = Params();
    p.cst_pi = 3.141592654f;
    const uint64_t im = h.path("imaging");
    if (im == UNDEF) throw std::runtime_error("input emd file has no /imaging group");
    auto need = [&](const char* name, auto& v) {
        if (!h.attr_num(im, name, v)) throw std::runtime_error(std::string("input emd file missed the specification of ") + name);
    };
    need("image_size_z", p.n3);
    if (p.n3 < 1) throw std::runtime_error("input emd file: image_size_z < 1");
    need("image_size_x", p.n1); need("image_size_y", p.n2); need("mode", p.mode);
    need("sample_size_x", p.m1); need("sample_size_y", p.m2); need("sample_size_z", p.m3);
    need("pixel_size_x", p.d1); need("pixel_size_y", p.d2); need("pixel_size_z", p.d3);
    need("border_size_x", p.dn1); need("border_size_y", p.dn2);
    h.attr_num(im, "specimen_tilt_offset_x", p.tilt_off[0]);
    h.attr_num(im, "specimen_tilt_offset_y", p.tilt_off[1]);
    h.attr_num(im, "specimen_tilt_offset_z", p.tilt_off[2]);
    h.attr_num(im, "frozen_phonons", p.frPh);
    h.attr_num(im, "pixel_dose", p.pD);
    p.subSlTh = p.d3;
    h.attr_num(im, "subpixel_size_z", p.subSlTh);
    p.tiltspec.assign(2 * (size_t)p.n3, 0.f); p.tiltbeam.assign(2 * (size_t)p.n3, 0.f); p.defoci.assign((size_t)p.n3, 0.f);
    auto per_k = [&](const char* name, std::vector<float>& dst, int stride, int off) {
        std::vector<float> v;
        if (!h.dataset(h.child(im, name), v)) return;
        if ((int)v.size() != p.n3) throw std::runtime_error(std::string("input emd file: /imaging/") + name + " does not have image_size_z entries");
        for (int i = 0; i < p.n3; i++) dst[(size_t)stride * i + off] = v[i];
    };
    per_k("specimen_tilt_x", p.tiltspec, 2, 0); per_k("specimen_tilt_y", p.tiltspec, 2, 1);
    per_k("beam_tilt_x", p.tiltbeam, 2, 0); per_k("beam_tilt_y", p.tiltbeam, 2, 1);
    per_k("defoci", p.defoci, 1, 0);
    const uint64_t mic = h.path("microscope");
    h.attr_num(mic, "voltage", p.E0); h.attr_num(mic, "gamma", p.gamma); h.attr_num(mic, "wavelength", p.lambda);
    h.attr_num(mic, "interaction_constant", p.sigma); h.attr_num(mic, "focus_spread", p.defocspread);
    h.attr_num(mic, "illumination_angle", p.illangle); h.attr_num(mic, "objective_aperture", p.ObjAp);
    h.attr_num(mic, "mtf_a", p.mtfa); h.attr_num(mic, "mtf_b", p.mtfb); h.attr_num(mic, "mtf_c", p.mtfc); h.attr_num(mic, "mtf_d", p.mtfd);
    const uint64_t ab = h.path("microscope/aberrations");
    for (int a = 0; a < AB_COUNT; a++) {
        h.attr_num(ab, std::string(kAberrationNames[a]) + "_amplitude", p.ab0[a]);
        if (a != AB_C1 && a != AB_C3 && a != AB_C5) h.attr_num(ab, std::string(kAberrationNames[a]) + "_angle", p.ab1[a]);
    }
    const uint64_t user = h.path("user");
    h.attr_str(user, "name", p.user_name); h.attr_str(user, "institution", p.institution);
    h.attr_str(user, "department", p.department); h.attr_str(user, "email", p.email);
    h.attr_str(h.path("comments"), "comment", p.comments);
    const uint64_t sam = h.path("sample");
    h.attr_str(sam, "name", p.sample_name); h.attr_str(sam, "material", p.material);
    h.attr_num(sam, "absorptive_potential_factor", p.imPot);
    if (!atoms_from_external && atoms) {
        Atoms& at = *atoms;
        at = Atoms();
        if (sam == UNDEF || !h.dataset(h.child(sam, "atomic_numbers"), at.Z))
            throw std::runtime_error("input emd file has no /sample/atomic_numbers");
        const size_t n = at.Z.size();
        std::vector<float> c[3];
        const char* names[3] = {"x_coordinates", "y_coordinates", "z_coordinates"};
        at.xyz.assign(3 * n, 0.f);
        for (int k = 0; k < 3; k++) {
            if (!h.dataset(h.child(sam, names[k]), c[k]) || c[k].size() != n)
                throw std::runtime_error(std::string("input emd file: /sample/") + names[k] + " missing or of the wrong length");
            for (size_t i = 0; i < n; i++) at.xyz[3 * i + k] = c[k][i];
        }
        if (!h.dataset(h.child(sam, "debeye_waller_factors"), at.dwf) || at.dwf.size() != n)
            throw std::runtime_error("input emd file: /sample/debeye_waller_factors missing or of the wrong length");
        if (!h.dataset(h.child(sam, "occupancy"), at.occ) || at.occ.size() != n)
            throw std::runtime_error("input emd file: /sample/occupancy missing or of the wrong length");
        p.nAt = (int)n;
    }
    consistent_params(p);
    return true;
}

}  // namespace fdes
