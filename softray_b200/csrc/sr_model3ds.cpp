// sr_model3ds.cpp -- .3DS stream -> flattened, unit-cube-normalised softray_mesh, in native code.
//
// the on-disk format feeding the hot path.  Follows the reference's loader and
// model post-processing so that the arrays handed to softray_scene_create are the ones its C# host
// would pin:
//   3dsLoader/ThreeDSFile.cs:132-185 (LoadModel), :187-257 (ProcessChunk), :259-327 (material),
//   :420-460 (colour / percentage), :462-573 (object, faces, per-face material groups), :575-662
//   (SkipChunk, strings, vertices, triangles, chunk header);
//   Model.cs:522-653 (Load3ds: entity merge, NaN / inf / |v| > 1e6 scrub, extent),
//   Model.cs:750-790 (PostProcessGeometry: centre + scale into the unit cube),
//   Renderer.cs:1452-1469 + Surface.cs:131-138 (per-triangle PackColorAndAlpha(diffuse, 1.0)).
// Plain FP64 expressions in the reference's order; this file is compiled with -ffp-contract=off.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <vector>

#include "../../include/softray_cuda.h"

namespace {

struct FormatError {};   // FormatException / EndOfStreamException / IndexOutOfRange of the C# loader

class Reader {
public:
    Reader(const uint8_t* p, size_t n) : p_(p), n_(n) {}
    size_t pos() const { return pos_; }
    void seek(size_t to) { pos_ = to > n_ ? n_ : to; }
    bool seek_checked(int64_t to)
    {
        if (to < 0) to = 0;
        if ((uint64_t)to > n_) { pos_ = n_; return false; }
        pos_ = (size_t)to;
        return true;
    }
    uint8_t u8() { need(1); return p_[pos_++]; }
    uint16_t u16() { need(2); const uint16_t v = (uint16_t)(p_[pos_] | (p_[pos_ + 1] << 8)); pos_ += 2; return v; }
    uint32_t u32()
    {
        need(4);
        const uint32_t v = (uint32_t)p_[pos_] | ((uint32_t)p_[pos_ + 1] << 8) | ((uint32_t)p_[pos_ + 2] << 16) |
                           ((uint32_t)p_[pos_ + 3] << 24);
        pos_ += 4;
        return v;
    }
    float f32() { const uint32_t u = u32(); float f; std::memcpy(&f, &u, 4); return f; }

private:
    void need(size_t k) { if (pos_ + k > n_) throw FormatError(); }
    const uint8_t* p_;
    size_t n_, pos_ = 0;
};

// ThreeDSChunk (ThreeDSFile.cs:664-690): id, length, running count of consumed bytes
struct Chunk {
    uint16_t id;
    uint32_t length;
    size_t start;
    int64_t consumed;
    explicit Chunk(Reader& r) : start(r.pos())
    {
        id = r.u16();
        length = r.u32();
        consumed = 6;
        if (length < 6) throw FormatError();   // the reference would seek backwards forever (SkipChunk)
    }
    bool more() const { return consumed < (int64_t)length; }
    void jump_to_end(Reader& r) const { r.seek(start + length); }
    void skip_rest(Reader& r)                  // SkipChunk (ThreeDSFile.cs:575-588)
    {
        const int64_t left = (int64_t)length - consumed;
        if (!r.seek_checked((int64_t)r.pos() + left)) throw FormatError();
        consumed += left;
    }
};

struct Material { std::string name; float diffuse[3] = {0.0f, 0.0f, 0.0f}; };   // Material.cs:33
struct Face { int32_t v[3]; int32_t material = -1; };                            // -1: Triangle.defaultMaterial
struct Entity {
    std::vector<double> verts;   // x, y, z per vertex, already in the reference's axis convention
    std::vector<Face> faces;
    bool has_verts = false, has_faces = false;
};

class Loader {
public:
    explicit Loader(Reader& r) : r_(r) {}
    std::vector<Material> materials;
    std::vector<Entity> entities;

    void run()
    {
        Chunk top(r_);
        if (top.id != 0x4D4D) throw FormatError();   // "Not a proper 3DS file." (ThreeDSFile.cs:160-168)
        chunks(top);
    }

private:
    Reader& r_;

    std::string cstring(Chunk& c)              // ProcessString (ThreeDSFile.cs:590-606)
    {
        std::string s;
        int n = 0;
        for (uint8_t b = r_.u8(); b != 0; b = r_.u8()) { s.push_back((char)b); n++; }
        c.consumed += n + 1;
        return s;
    }

    void colour(Chunk& parent, float rgb[3])   // ProcessColorChunk (:420-450): first sub-chunk only
    {
        Chunk c(r_);
        rgb[0] = rgb[1] = rgb[2] = 1.0f;
        if (c.id == 0x0010) { rgb[0] = r_.f32(); rgb[1] = r_.f32(); rgb[2] = r_.f32(); }
        else if (c.id == 0x0011) {
            rgb[0] = (float)r_.u8() / 255.0f; rgb[1] = (float)r_.u8() / 255.0f; rgb[2] = (float)r_.u8() / 255.0f;
        }
        parent.consumed += (int64_t)c.length;
        c.jump_to_end(r_);
    }

    void percentage(Chunk& parent)             // ProcessPercentageChunk (:452-460)
    {
        Chunk c(r_);
        (void)r_.u16();
        c.consumed += 2;
        parent.consumed += c.consumed;
        c.jump_to_end(r_);
    }

    void texture_map(Chunk& parent)            // ProcessTexMapChunk (:329-418): unused by the raytracer
    {
        while (parent.more()) {
            Chunk c(r_);
            if (c.id == 0xA300) (void)cstring(c); else c.skip_rest(r_);
            parent.consumed += c.consumed;
            c.jump_to_end(r_);
        }
    }

    void material(Chunk& parent)               // ProcessMaterialChunk (:259-327)
    {
        Material m;
        float unused[3];
        while (parent.more()) {
            Chunk c(r_);
            switch (c.id) {
            case 0xA000: m.name = cstring(c); break;
            case 0xA010: colour(c, unused); break;
            case 0xA020: colour(c, m.diffuse); break;
            case 0xA030: colour(c, unused); break;
            case 0xA040: percentage(c); break;
            case 0xA200: percentage(c); texture_map(c); break;
            default: c.skip_rest(r_); break;
            }
            parent.consumed += c.consumed;
            c.jump_to_end(r_);
        }
        for (const Material& have : materials)
            if (have.name == m.name) return;   // duplicate names are ignored (:323-326)
        materials.push_back(m);
    }

    void face_materials(Chunk& parent, Entity& e)   // ProcessFaceChunk (:522-573)
    {
        while (parent.more()) {
            Chunk c(r_);
            if (c.id == 0x4130) {
                const std::string name = cstring(c);
                int32_t mat = -1;
                for (size_t i = 0; i < materials.size(); i++)
                    if (materials[i].name == name) { mat = (int32_t)i; break; }
                const int n = r_.u16();
                c.consumed += 2;
                for (int i = 0; i < n; i++) {
                    const size_t fi = r_.u16();
                    if (fi >= e.faces.size()) throw FormatError();
                    e.faces[fi].material = mat;
                    c.consumed += 2;
                }
            }
            c.skip_rest(r_);
            parent.consumed += c.consumed;
            c.jump_to_end(r_);
        }
    }

    void object(Chunk& parent, Entity& e)      // ProcessObjectChunk (:462-520)
    {
        while (parent.more()) {
            Chunk c(r_);
            switch (c.id) {
            case 0x4100: object(c, e); break;
            case 0x4110: {                     // ReadVertices (:608-632): (x, y, z)_file -> (x, z, -y)
                const int n = r_.u16();
                c.consumed += 2;
                e.verts.assign((size_t)3 * n, 0.0);
                e.has_verts = true;
                for (int i = 0; i < n; i++) {
                    const float a = r_.f32(), b = r_.f32(), cc = r_.f32();
                    e.verts[3 * (size_t)i] = a; e.verts[3 * (size_t)i + 1] = cc; e.verts[3 * (size_t)i + 2] = -b;
                }
                c.consumed += (int64_t)n * 12;
                break;
            }
            case 0x4120: {                     // ReadTriangles (:634-657); the face flags are dropped
                const int n = r_.u16();
                c.consumed += 2;
                e.faces.assign((size_t)n, Face());
                e.has_faces = true;
                for (int i = 0; i < n; i++) {
                    Face& f = e.faces[(size_t)i];
                    f.v[0] = r_.u16(); f.v[1] = r_.u16(); f.v[2] = r_.u16();
                    (void)r_.u16();
                }
                c.consumed += (int64_t)n * 8;
                if (c.more()) face_materials(c, e);
                break;
            }
            case 0x4140: {                     // texture coordinates: read and dropped
                const int n = r_.u16();
                c.consumed += 2;
                for (int i = 0; i < n; i++) { (void)r_.f32(); (void)r_.f32(); }
                c.consumed += (int64_t)n * 8;
                break;
            }
            default: c.skip_rest(r_); break;
            }
            parent.consumed += c.consumed;
            c.jump_to_end(r_);
        }
    }

    void chunks(Chunk& parent)                 // ProcessChunk (:187-257)
    {
        while (parent.more()) {
            Chunk c(r_);
            switch (c.id) {
            case 0x0002: (void)r_.u32(); c.consumed += 4; break;
            case 0x3D3D: {
                Chunk blind(r_);               // the first sub-chunk is skipped unseen (:208-216)
                blind.skip_rest(r_);
                c.consumed += blind.consumed;
                chunks(c);
                break;
            }
            case 0xAFFF: material(c); break;
            case 0x4000: {
                (void)cstring(c);
                Entity e;
                object(c, e);
                if (e.has_verts && e.has_faces) entities.push_back(std::move(e));
                break;
            }
            default: c.skip_rest(r_); break;
            }
            parent.consumed += c.consumed;
            if (c.id != 0x0002) c.jump_to_end(r_);
        }
    }
};

inline uint32_t channel(float v)               // (byte)(float -> double * 255.0): truncation (Surface.cs:131-138)
{
    return (uint32_t)(int32_t)((double)v * 255.0) & 0xffu;
}
inline uint32_t pack_color_and_alpha(const float rgb[3])
{
    return (0xffu << 24) + (channel(rgb[0]) << 16) + (channel(rgb[1]) << 8) + channel(rgb[2]);
}

}  // namespace

struct softray_model {
    std::vector<double> verts;
    std::vector<int32_t> tri_vidx;
    std::vector<uint32_t> tri_argb;
    double bbox_min[3], bbox_max[3];
};

extern "C" int softray_model_load_3ds(const uint8_t* bytes, uint64_t n_bytes, softray_model** out)
{
    if (!bytes || !out) return SOFTRAY_E_INVALID_ARG;
    *out = nullptr;
    try {
        Reader r(bytes, (size_t)n_bytes);
        Loader L(r);
        L.run();
        if (L.entities.empty()) return SOFTRAY_E_FORMAT;                        // "No entities in model"
        softray_model* m = new softray_model();
        double mn[3], mx[3];
        for (int k = 0; k < 3; k++) { mn[k] = std::numeric_limits<double>::max(); mx[k] = -std::numeric_limits<double>::max(); }
        const float none[3] = {0.0f, 0.0f, 0.0f};
        for (const Entity& e : L.entities) {                                     // Model.cs:560-653: merge entities
            if (e.verts.size() < 9 || e.faces.empty()) { delete m; return SOFTRAY_E_FORMAT; }
            const int32_t base = (int32_t)(m->verts.size() / 3);
            for (size_t i = 0; i < e.verts.size(); i++) {
                double x = e.verts[i];
                if (std::isnan(x) || std::isinf(x) || std::fabs(x) > 1e6) x = 0.0;   // maxCoordinateSize (Model.cs:585-609)
                m->verts.push_back(x);
                mn[i % 3] = std::fmin(mn[i % 3], x); mx[i % 3] = std::fmax(mx[i % 3], x);
            }
            for (const Face& f : e.faces) {
                for (int k = 0; k < 3; k++) m->tri_vidx.push_back(base + f.v[k]);
                m->tri_argb.push_back(pack_color_and_alpha(f.material >= 0 ? L.materials[(size_t)f.material].diffuse : none));
            }
        }
        // the C# host would fault on a vertex index beyond the merged list when it builds the triangles
        const int32_t n_verts = (int32_t)(m->verts.size() / 3);
        for (int32_t vi : m->tri_vidx)
            if (vi < 0 || vi >= n_verts) { delete m; return SOFTRAY_E_FORMAT; }
        // Model.PostProcessGeometry (Model.cs:750-790): centre on the origin, longest axis -> [-0.5, 0.5]
        const double centre[3] = {(mn[0] + mx[0]) / 2, (mn[1] + mx[1]) / 2, (mn[2] + mx[2]) / 2};
        const double ext[3] = {mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]};
        const double scale = 1.0 / std::fmax(std::fmax(ext[0], ext[1]), ext[2]);
        for (size_t i = 0; i < m->verts.size(); i++) m->verts[i] = (m->verts[i] - centre[i % 3]) * scale;
        for (int k = 0; k < 3; k++) { m->bbox_min[k] = (mn[k] - centre[k]) * scale; m->bbox_max[k] = (mx[k] - centre[k]) * scale; }
        *out = m;
        return SOFTRAY_OK;
    } catch (const FormatError&) {
        return SOFTRAY_E_FORMAT;
    } catch (const std::bad_alloc&) {
        return SOFTRAY_E_OOM;
    }
}

extern "C" int softray_model_get_mesh(const softray_model* model, softray_mesh* out)
{
    if (!model || !out) return SOFTRAY_E_INVALID_ARG;
    out->verts_xyz = model->verts.data();
    out->tri_vidx = model->tri_vidx.data();
    out->tri_argb = model->tri_argb.data();
    out->n_verts = (int32_t)(model->verts.size() / 3);
    out->n_tris = (int32_t)model->tri_argb.size();
    for (int k = 0; k < 3; k++) { out->bbox_min[k] = model->bbox_min[k]; out->bbox_max[k] = model->bbox_max[k]; }
    return SOFTRAY_OK;
}

extern "C" void softray_model_destroy(softray_model* model) { delete model; }
