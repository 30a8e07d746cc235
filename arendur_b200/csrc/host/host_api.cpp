// Host layer C entry points (include/arn_host.h) plus the host-only arn.h functions
// arn_film_finalize.  Product code: scene ingest, flattening, camera set-up, image output.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include "../../../include/arn_host.h"
#include "flat_scene.hpp"
#include "image_io.hpp"
#include "json.hpp"
#include "obj_loader.hpp"

using namespace arnhost;

// tex_cache: image textures already loaded, keyed by file + ImageInfo + UVMapping (the reference shares MipMaps the same way,
// texturing/textures/image.rs:170-199)
struct arn_hscene { FlatScene fs; std::map<std::string, int> tex_cache; };

namespace {

// ImageTexture::new_as_arc (image.rs:151-199) for a PNG file: builds the pyramid (host/image_io.hpp), appends it to the scene's
// texture table.  Returns the texture id (>= 1), or 0 when the picture cannot be opened / decoded (`why` says so).
int load_image_texture(arn_hscene& hs, const std::string& path, uint32_t channels, bool trilinear, float max_aniso, uint32_t wrapping,
                       bool gamma, float scale, const float* scaling2, const float* shifting2, float* mean_out, std::string* why) {
    char keybuf[160];
    std::snprintf(keybuf, sizeof keybuf, "|%u|%d|%.9g|%u|%d|%.9g|%.9g,%.9g|%.9g,%.9g", channels, (int)trilinear, max_aniso, wrapping, (int)gamma, scale,
                  scaling2[0], scaling2[1], shifting2[0], shifting2[1]);
    const std::string key = path + keybuf;
    arn_texture t; std::memset(&t, 0, sizeof t);
    std::vector<float> texels; float mean[3] = {0.f, 0.f, 0.f};
    auto it = hs.tex_cache.find(key);
    if (it != hs.tex_cache.end() && !mean_out) return it->second;
    try {
        if (!build_pyramid(path, channels, gamma, scale, &t, &texels, mean, why)) return 0;
    } catch (const std::exception& e) { if (why) *why = path + ": " + e.what(); return 0; }      // e.g. bad_alloc on a huge picture: never across the C boundary
    if (mean_out) for (uint32_t c = 0; c < channels; c++) mean_out[c] = mean[c];
    if (it != hs.tex_cache.end()) return it->second;
    t.trilinear = trilinear ? 1u : 0u; t.wrapping = wrapping; t.max_aniso = max_aniso;
    t.scale_u = scaling2[0]; t.scale_v = scaling2[1]; t.shift_u = shifting2[0]; t.shift_v = shifting2[1];
    int id = hs.fs.add_texture(t, texels.data(), (uint64_t)texels.size());
    if (id <= 0) { if (why) *why = hs.fs.err; return 0; }
    hs.tex_cache[key] = id;
    return id;
}
std::string path_join(const std::string& dir, const std::string& name) {
    if (dir.empty() || (!name.empty() && name[0] == '/')) return name;
    return dir + (dir.back() == '/' ? "" : "/") + name;
}

// component::load_obj's material choice (src/component/mod.rs:70-173).  map_Kd / map_Ks are opened next to the OBJ file, map_bump
// as written (:76, :97, :125): EWA look-ups (trilinear false), max_aniso 16, Repeat, no gamma, scale 1, identity UV mapping;
// a picture that cannot be opened leaves the MTL constant in place (:88-95,:111-118) — here that includes every non-PNG file.
int obj_material_to_arn(const ObjMaterial& mtl, arn_hscene& hs, const std::string& parent_dir, std::string* warn) {
    FlatScene& fs = hs.fs;
    arn_material m; std::memset(&m, 0, sizeof m);
    for (int k = 0; k < 3; k++) { m.kd[k] = mtl.diffuse[k]; m.ks[k] = mtl.specular[k]; }
    const float one2[2] = {1.f, 1.f}, zero2[2] = {0.f, 0.f};
    float spec_mean[3] = {mtl.specular[0], mtl.specular[1], mtl.specular[2]};           // specular.mean(): the constant, or MipMap::mean
    if (!mtl.diffuse_texture.empty()) {
        std::string why;
        int id = load_image_texture(hs, path_join(parent_dir, mtl.diffuse_texture), 3, false, 16.f, ARN_WRAP_REPEAT, false, 1.f, one2, zero2, nullptr, &why);
        if (id > 0) m.kd_tex = (uint32_t)id; else if (warn) *warn += "diffuse texture " + mtl.diffuse_texture + " unfound (" + why + "); constant used\n";
    }
    if (!mtl.specular_texture.empty()) {
        std::string why; float mean[3];
        int id = load_image_texture(hs, path_join(parent_dir, mtl.specular_texture), 3, false, 16.f, ARN_WRAP_REPEAT, false, 1.f, one2, zero2, mean, &why);
        if (id > 0) { m.ks_tex = (uint32_t)id; for (int k = 0; k < 3; k++) spec_mean[k] = mean[k]; }
        else if (warn) *warn += "specular texture " + mtl.specular_texture + " unfound (" + why + "); constant used\n";
    }
    {
        auto bt = mtl.unknown_param.find("map_bump");
        if (bt != mtl.unknown_param.end() && !bt->second.empty()) {
            std::string why;
            int id = load_image_texture(hs, bt->second, 1, false, 16.f, ARN_WRAP_REPEAT, false, 1.f, one2, zero2, nullptr, &why);
            if (id > 0) m.bump_tex = (uint32_t)id; else if (warn) *warn += "bump map " + bt->second + " unfound (" + why + "); none used\n";
        }
    }
    float rough = (1000.f - mtl.shininess) / 1000.f;                         // :120-122
    rough = rough < 1.f ? rough : 1.f; rough = rough > 0.f ? rough : 0.f;     // .min(1.).max(0.)
    m.roughness = rough;
    std::string illum = "2";
    auto it = mtl.unknown_param.find("illum"); if (it != mtl.unknown_param.end()) illum = it->second;
    float dissolve = mtl.dissolve > 0.f ? mtl.dissolve : 0.f; dissolve = dissolve < 1.f ? dissolve : 1.f;  // .max(0.).min(1.)
    bool spec_black = spec_mean[0] == 0.f && spec_mean[1] == 0.f && spec_mean[2] == 0.f;
    bool spec_valid = spec_mean[0] >= 0.f && spec_mean[1] >= 0.f && spec_mean[2] >= 0.f && std::isfinite(spec_mean[0]) && std::isfinite(spec_mean[1]) && std::isfinite(spec_mean[2]);
    // relative_eq!(dissolve, 1.0): |d - 1| <= eps or <= eps * max(|d|, 1)
    bool dissolve_is_one = dissolve == 1.f || std::fabs(dissolve - 1.f) <= 1.1920929e-7f;
    if (illum.find('4') != std::string::npos) { m.type = ARN_MAT_GLASS; m.eta = mtl.optical_density; }
    else if (!dissolve_is_one) { m.type = ARN_MAT_TRANSLUCENT; m.dissolve = dissolve; }
    else if (spec_black || !spec_valid) { m.type = ARN_MAT_MATTE; m.sigma = 0.f; m.ks[0] = m.ks[1] = m.ks[2] = 0.f; m.ks_tex = 0; }
    else m.type = ARN_MAT_PLASTIC;
    return fs.add_material(m);
}

int load_obj_into(arn_hscene& hs, const std::string& path, const float* transform16, std::string* err) {
    FlatScene& fs = hs.fs;
    std::vector<ObjModel> models; std::vector<ObjMaterial> mtls;
    if (!load_obj_file(path, models, mtls, err)) return ARN_E_IO;
    const size_t slash = path.find_last_of('/');
    const std::string parent_dir = slash == std::string::npos ? std::string() : path.substr(0, slash);      // path.parent()
    std::vector<int> mat_ids; std::string warn;
    for (auto& mtl : mtls) { int id = obj_material_to_arn(mtl, hs, parent_dir, &warn); if (id < 0) { *err = fs.err; return id; } mat_ids.push_back(id); }
    if (!warn.empty() && std::getenv("ARN_VERBOSE")) std::fprintf(stderr, "%s", warn.c_str());
    // fallback material appended after the MTL ones (:165-171)
    arn_material dflt; std::memset(&dflt, 0, sizeof dflt);
    dflt.type = ARN_MAT_MATTE; dflt.kd[0] = 0.5f; dflt.kd[1] = 0.6f; dflt.kd[2] = 0.7f;
    int dflt_id = fs.add_material(dflt);
    int ntri = 0;
    for (auto& mdl : models) {
        int mid = mdl.material_id >= 0 ? mat_ids[(size_t)mdl.material_id] : dflt_id;
        size_t nv = mdl.positions.size() / 3;
        const float* nrm = mdl.normals.empty() ? nullptr : mdl.normals.data();
        const float* uv = mdl.texcoords.empty() ? nullptr : mdl.texcoords.data();
        if (nrm && mdl.normals.size() != nv * 3) { *err = "model '" + mdl.name + "': some vertices lack normals (the reference would index out of bounds)"; return ARN_E_UNSUPPORTED; }
        if (uv && mdl.texcoords.size() != nv * 2) { *err = "model '" + mdl.name + "': some vertices lack texture coordinates"; return ARN_E_UNSUPPORTED; }
        int rc = fs.add_mesh(mdl.positions.data(), (uint32_t)nv, mdl.indices.data(), (uint32_t)mdl.indices.size(), nrm, uv, transform16, (uint32_t)mid);
        if (rc < 0) { *err = fs.err; return rc; }
        ntri += (int)(mdl.indices.size() / 3);
    }
    return ntri;
}

// cgmath Matrix4 via serde: 4 columns, as array-of-arrays or as a map {x, y, z, w}
bool json_matrix(const Json& j, float* m16) {
    const Json* cols[4] = {nullptr, nullptr, nullptr, nullptr};
    if (j.kind == Json::Arr && j.arr.size() == 4) for (int i = 0; i < 4; i++) cols[i] = &j.arr[(size_t)i];
    else if (j.kind == Json::Obj) { cols[0] = j.get("x"); cols[1] = j.get("y"); cols[2] = j.get("z"); cols[3] = j.get("w"); }
    for (int c = 0; c < 4; c++) {
        if (!cols[c]) return false;
        const Json& col = *cols[c];
        if (col.kind == Json::Arr && col.arr.size() == 4) { for (int r = 0; r < 4; r++) { if (col.arr[(size_t)r].kind != Json::Num) return false; m16[c * 4 + r] = (float)col.arr[(size_t)r].num; } }
        else if (col.kind == Json::Obj) { const char* k[4] = {"x", "y", "z", "w"}; for (int r = 0; r < 4; r++) { const Json* e = col.get(k[r]); if (!e || e->kind != Json::Num) return false; m16[c * 4 + r] = (float)e->num; } }
        else return false;
    }
    return true;
}
bool json_num(const Json* j, float* out) { if (!j || j->kind != Json::Num) return false; *out = (float)j->num; return true; }
bool json_vec(const Json* j, const char* const* keys, int n, float* out) {
    if (!j) return false;
    if (j->kind == Json::Arr && (int)j->arr.size() == n) { for (int i = 0; i < n; i++) if (!json_num(&j->arr[(size_t)i], &out[i])) return false; return true; }
    if (j->kind == Json::Obj) { for (int i = 0; i < n; i++) if (!json_num(j->get(keys[i]), &out[i])) return false; return true; }
    return false;
}
const char* const XY[2] = {"x", "y"};
const char* const XYZ[3] = {"x", "y", "z"};
// {RGB,Gray}TextureDesc::Image{info: ImageInfo{name, trilinear, max_aniso, wrapping, gamma, scale}, mapping: UVMapping{scaling, shifting}}
// (examples/arencli.rs:377-380,433-436; image.rs:576-583; mappings.rs:14-19).  `name` is opened as written (image::open), then next to
// the scene file.  A picture that cannot be opened makes the reference drop the material's primitive silently (to_arc -> None,
// arencli.rs:307-316); here it is an error.
int json_image_texture(const Json* img, arn_hscene* hs, const std::string& base_dir, uint32_t channels, uint32_t* tex_out, std::string* err) {
    const Json* info = img->get("info"); const Json* map = img->get("mapping");
    if (!hs) { *err = "image textures are not available here (emission textures are constants)"; return ARN_E_UNSUPPORTED; }
    if (!info || !map || !info->get("name") || info->get("name")->kind != Json::Str) { *err = "malformed Image texture"; return ARN_E_INVALID; }
    float max_aniso = 8.f, scale = 1.f, scaling[2] = {1.f, 1.f}, shifting[2] = {0.f, 0.f};
    if (!json_num(info->get("max_aniso"), &max_aniso) || !json_num(info->get("scale"), &scale) || !json_vec(map->get("scaling"), XY, 2, scaling) ||
        !json_vec(map->get("shifting"), XY, 2, shifting)) { *err = "malformed Image texture parameters"; return ARN_E_INVALID; }
    auto flag = [](const Json* j, bool* out) { if (!j || j->kind != Json::Bool) return false; *out = j->b; return true; };
    bool trilinear = false, gamma = false;
    if (!flag(info->get("trilinear"), &trilinear) || !flag(info->get("gamma"), &gamma)) { *err = "malformed Image texture flags"; return ARN_E_INVALID; }
    const Json* wr = info->get("wrapping"); uint32_t wrapping;
    if (!wr || wr->kind != Json::Str) { *err = "malformed Image wrapping mode"; return ARN_E_INVALID; }
    if (wr->str == "Repeat") wrapping = ARN_WRAP_REPEAT; else if (wr->str == "Black") wrapping = ARN_WRAP_BLACK; else if (wr->str == "Clamp") wrapping = ARN_WRAP_CLAMP;
    else { *err = "unknown Image wrapping mode " + wr->str; return ARN_E_INVALID; }
    std::string why;
    int id = load_image_texture(*hs, info->get("name")->str, channels, trilinear, max_aniso, wrapping, gamma, scale, scaling, shifting, nullptr, &why);
    if (id <= 0) id = load_image_texture(*hs, path_join(base_dir, info->get("name")->str), channels, trilinear, max_aniso, wrapping, gamma, scale, scaling, shifting, nullptr, &why);
    if (id <= 0) { *err = "image texture: " + why; return ARN_E_IO; }
    *tex_out = (uint32_t)id;
    return ARN_OK;
}
// RGBTextureDesc::Constant{value: RGBSpectrumf{inner: Vector3}} or ::Image (hs != nullptr: `*tex_out` receives the texture id)
int json_rgb_texture(const Json* named, float* rgb, std::string* err, arn_hscene* hs = nullptr, const std::string& base_dir = std::string(), uint32_t* tex_out = nullptr) {
    if (!named) { *err = "missing RGB texture"; return ARN_E_INVALID; }
    const Json* v = named->get("value");
    if (!v || v->is_null()) { *err = "texture referenced by name only (no value); texture reuse by name never resolves in the reference either"; return ARN_E_UNSUPPORTED; }
    if (const Json* img = v->get("Image")) return json_image_texture(img, tex_out ? hs : nullptr, base_dir, 3, tex_out, err);
    const Json* c = v->get("Constant");
    if (!c) { *err = "Product textures never resolve in arencli (examples/arencli.rs:413-427): only Constant and Image textures are taken"; return ARN_E_UNSUPPORTED; }
    const Json* val = c->get("value"); const Json* inner = val ? val->get("inner") : nullptr;
    if (!json_vec(inner, XYZ, 3, rgb)) { *err = "malformed constant RGB texture"; return ARN_E_INVALID; }
    return ARN_OK;
}
int json_gray_texture(const Json* named, float* g, std::string* err, arn_hscene* hs = nullptr, const std::string& base_dir = std::string(), uint32_t* tex_out = nullptr) {
    if (!named) { *err = "missing gray texture"; return ARN_E_INVALID; }
    const Json* v = named->get("value");
    if (!v || v->is_null()) { *err = "texture referenced by name only (no value)"; return ARN_E_UNSUPPORTED; }
    if (const Json* img = v->get("Image")) return json_image_texture(img, tex_out ? hs : nullptr, base_dir, 1, tex_out, err);
    const Json* c = v->get("Constant");
    if (!c) { *err = "Product textures never resolve in arencli: only Constant and Image textures are taken"; return ARN_E_UNSUPPORTED; }
    if (!json_num(c->get("value"), g)) { *err = "malformed constant gray texture"; return ARN_E_INVALID; }
    return ARN_OK;
}
// MaterialDesc::to_arc (examples/arencli.rs:290-379)
int json_material(const Json& desc, arn_hscene& hs, const std::string& base_dir, std::string* err) {
    FlatScene& fs = hs.fs;
    arn_material m; std::memset(&m, 0, sizeof m);
    const Json* body; int rc;
    if ((body = desc.get("Matte"))) {
        m.type = ARN_MAT_MATTE;
        if ((rc = json_rgb_texture(body->get("kd"), m.kd, err, &hs, base_dir, &m.kd_tex)) != ARN_OK) return rc;
        if ((rc = json_gray_texture(body->get("sigma"), &m.sigma, err, &hs, base_dir, &m.aux_tex)) != ARN_OK) return rc;
    } else if ((body = desc.get("Glass")) || (body = desc.get("Plastic")) || (body = desc.get("Translucent"))) {
        m.type = desc.get("Glass") ? ARN_MAT_GLASS : (desc.get("Plastic") ? ARN_MAT_PLASTIC : ARN_MAT_TRANSLUCENT);
        if ((rc = json_rgb_texture(body->get("diffuse"), m.kd, err, &hs, base_dir, &m.kd_tex)) != ARN_OK) return rc;
        if ((rc = json_rgb_texture(body->get("specular"), m.ks, err, &hs, base_dir, &m.ks_tex)) != ARN_OK) return rc;
        if ((rc = json_gray_texture(body->get("roughness"), &m.roughness, err, &hs, base_dir, &m.aux_tex)) != ARN_OK) return rc;
        if (m.type == ARN_MAT_GLASS && !json_num(body->get("eta"), &m.eta)) { *err = "Glass needs eta"; return ARN_E_INVALID; }
        if (m.type == ARN_MAT_TRANSLUCENT && !json_num(body->get("dissolve"), &m.dissolve)) { *err = "Translucent needs dissolve"; return ARN_E_INVALID; }
    } else { *err = "unknown material description"; return ARN_E_INVALID; }
    const Json* bump = body->get("bump");
    if (bump && !bump->is_null()) {         // Option<Named<GrayTextureDesc>>: an Image displacement map (material/mod.rs:42-86)
        float unused = 0.f;
        if ((rc = json_gray_texture(bump, &unused, err, &hs, base_dir, &m.bump_tex)) != ARN_OK) return rc;
        if (!m.bump_tex) { *err = "constant bump textures are not taken (only Image displacement maps)"; return ARN_E_UNSUPPORTED; }
    }
    int id = fs.add_material(m);
    if (id < 0) *err = fs.err;
    return id;
}

std::string g_host_error;

}  // namespace

extern "C" {

int arn_hscene_create(arn_hscene** out) { if (!out) return ARN_E_INVALID; *out = new arn_hscene; return ARN_OK; }
void arn_hscene_destroy(arn_hscene* h) { delete h; }
const char* arn_hscene_last_error(const arn_hscene* h) { return h ? h->fs.err.c_str() : g_host_error.c_str(); }

int arn_hscene_add_material(arn_hscene* h, const arn_material* m) { if (!h || !m) return ARN_E_INVALID; return h->fs.add_material(*m); }
int arn_hscene_add_mesh(arn_hscene* h, const float* positions, uint32_t n_vertices, const uint32_t* indices, uint32_t n_indices,
                        const float* normals, const float* uvs, const float* transform16, uint32_t material) {
    if (!h) return ARN_E_INVALID;
    return h->fs.add_mesh(positions, n_vertices, indices, n_indices, normals, uvs, transform16, material);
}
int arn_hscene_add_sphere(arn_hscene* h, float radius, float zmin, float zmax, float phimax, uint32_t material,
                          const float* emission3, const float* transform16) {
    if (!h) return ARN_E_INVALID;
    return h->fs.add_sphere(radius, zmin, zmax, phimax, material, emission3, transform16);
}
int arn_hscene_add_texture(arn_hscene* h, const arn_texture* tex, const float* texels, uint64_t n_floats) { if (!h || !tex) return ARN_E_INVALID; return h->fs.add_texture(*tex, texels, n_floats); }
int arn_hscene_add_texture_file(arn_hscene* h, const char* path, const arn_texture* params, int gamma, float scale, float* mean3_out) {
    if (!h || !path || !params) return ARN_E_INVALID;
    if ((params->channels != 1 && params->channels != 3) || params->wrapping > ARN_WRAP_CLAMP) return h->fs.fail(ARN_E_INVALID, "texture file: channels must be 1 or 3, a valid wrap mode");
    const float sc[2] = {params->scale_u, params->scale_v}, sh[2] = {params->shift_u, params->shift_v};
    float mean[3] = {0.f, 0.f, 0.f}; std::string why;
    int id = load_image_texture(*h, path, params->channels, params->trilinear != 0, params->max_aniso, params->wrapping, gamma != 0, scale, sc, sh, mean, &why);
    if (id <= 0) return h->fs.fail(ARN_E_IO, "texture file: " + why);
    if (mean3_out) for (uint32_t c = 0; c < params->channels; c++) mean3_out[c] = mean[c];
    return id;
}
int arn_hscene_add_light(arn_hscene* h, const arn_analytic_light* light) { if (!h || !light) return ARN_E_INVALID; return h->fs.add_light(*light); }

int arn_point_light_make(const float* pos3, const float* intensity3, arn_analytic_light* out) {
    if (!pos3 || !intensity3 || !out) return ARN_E_INVALID;
    std::memset(out, 0, sizeof *out);
    out->type = ARN_LIGHT_POINT;
    for (int k = 0; k < 3; k++) { out->pos[k] = pos3[k]; out->intensity[k] = intensity3[k]; }
    return ARN_OK;
}
int arn_distant_light_make(const float* intensity3, const float* dir3, float world_radius, arn_analytic_light* out) {
    if (!intensity3 || !dir3 || !out) return ARN_E_INVALID;
    std::memset(out, 0, sizeof *out);
    out->type = ARN_LIGHT_DISTANT;
    Vec3 d = normalize(Vec3{dir3[0], dir3[1], dir3[2]});                    // DistantLight::new (distantlight.rs:27)
    out->dir[0] = d.x; out->dir[1] = d.y; out->dir[2] = d.z;
    for (int k = 0; k < 3; k++) out->intensity[k] = intensity3[k];
    out->world_radius = world_radius;
    return ARN_OK;
}
int arn_spot_light_make(const float* pos3, const float* towards3, const float* intensity3, float total_angle,
                        float start_falloff_angle, arn_analytic_light* out) {
    if (!pos3 || !towards3 || !intensity3 || !out) return ARN_E_INVALID;
    const float pi = 3.14159265358979323846f;
    // assert!s of SpotLight::new (pointlights.rs:105-107)
    if (!(total_angle > start_falloff_angle) || !(start_falloff_angle > 0.f) || !(total_angle < pi * 2.0f)) { g_host_error = "spot light: need 0 < start_falloff_angle < total_angle < 2 pi"; return ARN_E_INVALID; }
    std::memset(out, 0, sizeof *out);
    out->type = ARN_LIGHT_SPOT;
    Vec3 src = normalize(Vec3{towards3[0], towards3[1], towards3[2]}), dst{0.f, 0.f, 1.f};
    // cgmath 0.14 Quaternion::from_arc(src, dst, None)
    auto ulps_eq = [](float a, float b) {
        if (std::fabs(a - b) <= 1.1920929e-7f) return true;
        if ((a < 0.f) != (b < 0.f)) return false;
        int32_t ia, ib; std::memcpy(&ia, &a, 4); std::memcpy(&ib, &b, 4);
        int64_t diff = (int64_t)ia - (int64_t)ib; if (diff < 0) diff = -diff;
        return diff <= 4;
    };
    float mag_avg = std::sqrt((src.x * src.x + src.y * src.y + src.z * src.z) * (dst.x * dst.x + dst.y * dst.y + dst.z * dst.z));
    float dt = src.x * dst.x + src.y * dst.y + src.z * dst.z;
    float qs, qx, qy, qz;
    if (ulps_eq(dt, mag_avg)) { qs = 1.f; qx = qy = qz = 0.f; }
    else if (ulps_eq(dt, -mag_avg)) {
        // fallback axis: unit_x x src, or unit_y x src when that vanishes; rotation by half a turn
        Vec3 v{0.f * src.z - 0.f * src.y, 0.f * src.x - 1.f * src.z, 1.f * src.y - 0.f * src.x};
        if (ulps_eq(v.x, 0.f) && ulps_eq(v.y, 0.f) && ulps_eq(v.z, 0.f)) v = Vec3{1.f * src.z - 0.f * src.y, 0.f * src.x - 0.f * src.z, 0.f * src.y - 1.f * src.x};
        v = normalize(v);
        float half = pi * 0.5f, sn = (float)std::sin((double)half), cs = (float)std::cos((double)half);
        qs = cs; qx = v.x * sn; qy = v.y * sn; qz = v.z * sn;
    } else {
        qs = mag_avg + dt;
        qx = src.y * dst.z - src.z * dst.y; qy = src.z * dst.x - src.x * dst.z; qz = src.x * dst.y - src.y * dst.x;
        float inv = 1.f / std::sqrt(qs * qs + qx * qx + qy * qy + qz * qz);
        qs *= inv; qx *= inv; qy *= inv; qz *= inv;
    }
    // Matrix4::from(Quaternion)
    float x2 = qx + qx, y2 = qy + qy, z2 = qz + qz;
    float xx2 = x2 * qx, xy2 = x2 * qy, xz2 = x2 * qz, yy2 = y2 * qy, yz2 = y2 * qz, zz2 = z2 * qz;
    float sy2 = y2 * qs, sz2 = z2 * qs, sx2 = x2 * qs;
    Mat4 rot = Mat4::identity();
    rot.c[0][0] = 1.f - yy2 - zz2; rot.c[0][1] = xy2 + sz2; rot.c[0][2] = xz2 - sy2;
    rot.c[1][0] = xy2 - sz2; rot.c[1][1] = 1.f - xx2 - zz2; rot.c[1][2] = yz2 + sx2;
    rot.c[2][0] = xz2 + sy2; rot.c[2][1] = yz2 - sx2; rot.c[2][2] = 1.f - xx2 - yy2;
    Mat4 parent_local = rot * Mat4::translation(pos3[0] - 0.f, pos3[1] - 0.f, pos3[2] - 0.f), local_parent;
    if (!invert(parent_local, &local_parent)) { g_host_error = "spot light: invalid inversion"; return ARN_E_INVALID; }   // .expect("invalid inversion")
    parent_local.to_array(out->parent_local);
    for (int k = 0; k < 3; k++) { out->pos[k] = pos3[k]; out->intensity[k] = intensity3[k]; }
    out->cost = (float)std::cos((double)total_angle); out->cosf = (float)std::cos((double)start_falloff_angle);
    return ARN_OK;
}

int arn_hscene_load_obj(arn_hscene* h, const char* path, const float* transform16) {
    if (!h || !path) return ARN_E_INVALID;
    std::string err; int rc = load_obj_into(*h, path, transform16, &err);
    if (rc < 0) h->fs.err = err;
    return rc;
}
int arn_hscene_build(arn_hscene* h, int strategy) { if (!h) return ARN_E_INVALID; return h->fs.build(strategy); }
int arn_hscene_build_gpu(arn_hscene* h, arn_ctx* ctx, float* build_ms_out) {
    if (!h || !ctx) return ARN_E_INVALID;
    int rc = h->fs.build(ARN_BVH_SAH, ctx);
    if (rc == ARN_OK && build_ms_out) *build_ms_out = h->fs.last_build_ms;
    return rc;
}
const arn_scene_desc* arn_hscene_desc(const arn_hscene* h) { return (h && h->fs.built) ? &h->fs.desc : nullptr; }

int arn_camera_make(const float* parent_view16, const float* screen4, float znear, float zfar, float fov, int has_lens,
                    float lens_radius, float focal_distance, float res_x, float res_y, arn_camera* out) {
    if (!parent_view16 || !screen4 || !out) return ARN_E_INVALID;
    if (!(znear < zfar) || !(fov < 3.14159265358979323846f)) { g_host_error = "camera: need znear < zfar and fov < pi"; return ARN_E_INVALID; }  // assert! perspective.rs:94-95
    Mat4 parent_view = Mat4::from_array(parent_view16), view_parent;
    if (!invert(parent_view, &view_parent)) { g_host_error = "camera transform: matrix inversion failure"; return ARN_E_INVALID; }
    // PerspecCam::perspective_transform (filming/perspective.rs:93-107)
    Mat4 persp; std::memset(&persp, 0, sizeof persp);
    persp.c[0][0] = 1.f; persp.c[1][1] = 1.f; persp.c[2][2] = zfar / (zfar - znear); persp.c[2][3] = 1.f;
    persp.c[3][2] = -zfar * znear / (zfar - znear); persp.c[3][3] = 0.f;
    float inv_tan = 1.f / (float)std::tan((double)(fov * 0.5f));
    Mat4 view_screen = Mat4::scale(inv_tan, inv_tan, 1.f) * persp;
    // ProjCameraInfo::new (filming/projective.rs:24-45)
    Mat4 raster_screen = Mat4::translation(screen4[0], screen4[3], 0.f)
                       * Mat4::scale((screen4[2] - screen4[0]) / res_x, (screen4[1] - screen4[3]) / res_y, 1.f);
    Mat4 tmp, inv_vs;
    if (!invert(raster_screen, &tmp)) { g_host_error = "camera: raster_screen is singular"; return ARN_E_INVALID; }
    if (!invert(view_screen, &inv_vs)) { g_host_error = "camera: matrix inversion failure"; return ARN_E_INVALID; }
    Mat4 raster_view = inv_vs * raster_screen;
    raster_view.to_array(out->raster_view); view_parent.to_array(out->view_parent);
    out->has_lens = has_lens ? 1u : 0u; out->lens_radius = lens_radius; out->focal_distance = focal_distance; out->ortho = 0u;
    return ARN_OK;
}

int arn_ortho_camera_make(const float* view_parent16, const float* screen4, float znear, float zfar, int has_lens,
                          float lens_radius, float focal_distance, float res_x, float res_y, arn_camera* out) {
    if (!view_parent16 || !screen4 || !out) return ARN_E_INVALID;
    Mat4 view_parent = Mat4::from_array(view_parent16), parent_view;
    if (!invert(view_parent, &parent_view)) { g_host_error = "camera transform: matrix inversion failure"; return ARN_E_INVALID; }   // ortho.rs:39
    // OrthoCam::ortho_transform (ortho.rs:57-66): scale(1, 1, 1 / (zfar - znear)) * translation(0, 0, -znear)
    Mat4 view_screen = Mat4::scale(1.f, 1.f, 1.f / (zfar - znear)) * Mat4::translation(0.f, 0.f, -znear);
    // ProjCameraInfo::new (filming/projective.rs:24-45)
    Mat4 raster_screen = Mat4::translation(screen4[0], screen4[3], 0.f)
                       * Mat4::scale((screen4[2] - screen4[0]) / res_x, (screen4[1] - screen4[3]) / res_y, 1.f);
    Mat4 tmp, inv_vs;
    if (!invert(raster_screen, &tmp)) { g_host_error = "camera: raster_screen is singular"; return ARN_E_INVALID; }
    if (!invert(view_screen, &inv_vs)) { g_host_error = "camera: matrix inversion failure"; return ARN_E_INVALID; }
    Mat4 raster_view = inv_vs * raster_screen;
    raster_view.to_array(out->raster_view); view_parent.to_array(out->view_parent);
    out->has_lens = has_lens ? 1u : 0u; out->lens_radius = lens_radius; out->focal_distance = focal_distance; out->ortho = 1u;
    return ARN_OK;
}

// examples/arencli.rs parse_input (:70-204)
int arn_hscene_load_json(arn_hscene* h, const char* json_path, const char* base_dir, arn_camera* cam, arn_film* film,
                         arn_sampler* sampler, arn_pt_params* params, char* outname, size_t outname_cap) {
    if (!h || !json_path || !cam || !film || !sampler || !params) return ARN_E_INVALID;
    FlatScene& fs = h->fs;
    std::ifstream f(json_path);
    if (!f) return fs.fail(ARN_E_IO, std::string("cannot open ") + json_path);
    std::stringstream ss; ss << f.rdbuf(); std::string text = ss.str();
    Json root; std::string perr;
    if (!JsonParser(text).parse(&root, &perr)) return fs.fail(ARN_E_IO, "JSON decode error: " + perr);
    // `for light in scenedesc.lights.iter() { lights.push(light.to_arc()) }` (arencli.rs:95-98): serde's
    // externally tagged LightDesc::{Point, Spot, Distant} carrying the light structs field by field
    const Json* lights = root.get("lights");
    if (lights && lights->kind == Json::Arr) for (const Json& l : lights->arr) {
        arn_analytic_light al; std::memset(&al, 0, sizeof al);
        const Json* b;
        auto rgb = [&](const Json* j, float* o) { return j && json_vec(j->get("inner"), XYZ, 3, o); };
        if ((b = l.get("Point"))) {
            al.type = ARN_LIGHT_POINT;
            if (!json_vec(b->get("posw"), XYZ, 3, al.pos) || !rgb(b->get("intensity"), al.intensity)) return fs.fail(ARN_E_INVALID, "malformed Point light");
        } else if ((b = l.get("Spot"))) {
            al.type = ARN_LIGHT_SPOT;
            const Json* pl = b->get("parent_local"); const Json* lp = b->get("local_parent"); float unused[16];
            if (!json_vec(b->get("posw"), XYZ, 3, al.pos) || !rgb(b->get("intensity"), al.intensity) || !json_num(b->get("cost"), &al.cost)
                || !json_num(b->get("cosf"), &al.cosf) || !pl || !json_matrix(*pl, al.parent_local) || !lp || !json_matrix(*lp, unused))
                return fs.fail(ARN_E_INVALID, "malformed Spot light");
        } else if ((b = l.get("Distant"))) {
            al.type = ARN_LIGHT_DISTANT;
            float centre[3];
            if (!rgb(b->get("intensity"), al.intensity) || !json_vec(b->get("dir"), XYZ, 3, al.dir) || !json_vec(b->get("world_center"), XYZ, 3, centre)
                || !json_num(b->get("world_radius"), &al.world_radius)) return fs.fail(ARN_E_INVALID, "malformed Distant light");
        } else return fs.fail(ARN_E_INVALID, "unknown light description");
        int rc = fs.add_light(al);
        if (rc < 0) return rc;
    }
    const Json* comps = root.get("components");
    if (!comps || comps->kind != Json::Arr) return fs.fail(ARN_E_INVALID, "scene has no components array");
    std::string base = base_dir ? std::string(base_dir) : std::string();
    if (!base.empty() && base.back() != '/') base.push_back('/');
    std::map<std::string, int> materials;
    struct ShapedRecord { float radius, zmin, zmax, phimax; uint32_t material; bool emissive; float emission[3]; bool transformed; };
    std::map<std::string, ShapedRecord> shaped_by_name;          // `primitives` (arencli.rs:90): what a Transformed component may refer to
    // pass 1: meshes in file order; pass 2: shaped primitives in file order (fixed component order)
    for (int pass = 0; pass < 2; pass++) for (const Json& c : comps->arr) {
        const Json* value = c.get("value");
        if (!value || value->is_null()) continue;                         // "ignoring empty component"
        const Json* mesh = value->get("Mesh"); const Json* shaped = value->get("Shaped");
        const Json* transformed = value->get("Transformed");
        if (pass == 1 && transformed) {
            // ComponentDesc::Transformed{transform, original} (arencli.rs:162-181): TransformedComposable over a primitive
            // already in `primitives`.  One level — an instance of a BARE sphere — is exactly a transformed sphere; an
            // instance of an already transformed primitive would round-trip rays through two matrices in turn, which the
            // flattened sphere record does not model.  The instance is not pushed to `lights` even when it is emissive.
            const Json* orig = transformed->get("original"); const Json* tr = transformed->get("transform");
            float t16[16];
            if (!orig || orig->kind != Json::Str || !tr || !json_matrix(*tr, t16)) return fs.fail(ARN_E_INVALID, "malformed Transformed component");
            { Mat4 m = Mat4::from_array(t16), inv; if (!invert(m, &inv)) continue; }          // "load transformed failed, invalid matrix invert"
            auto it = shaped_by_name.find(orig->str);
            if (it == shaped_by_name.end()) continue;                                      // "original doesn't exists": skipped
            const ShapedRecord& r = it->second;
            if (r.transformed) return fs.fail(ARN_E_UNSUPPORTED, "Transformed component over an already transformed primitive (nested TransformedComposable) is not flattened");
            int rc = fs.add_sphere(r.radius, r.zmin, r.zmax, r.phimax, r.material, r.emissive ? r.emission : nullptr, t16, false);
            if (rc < 0) return rc;
            const Json* nm = c.get("name");
            if (nm && nm->kind == Json::Str) { ShapedRecord inst = r; inst.transformed = true; shaped_by_name[nm->str] = inst; }
        }
        if (pass == 0 && mesh) {
            const Json* fn = mesh->get("filename");
            if (!fn || fn->kind != Json::Str) return fs.fail(ARN_E_INVALID, "Mesh without filename");
            float t16[16]; const float* tp = nullptr;
            const Json* tr = mesh->get("transform");
            if (tr && !tr->is_null()) { if (!json_matrix(*tr, t16)) return fs.fail(ARN_E_INVALID, "malformed mesh transform"); tp = t16; }
            // `transform.unwrap_or(identity)` then from_model_transformed: identity is applied as a matrix too
            float ident[16] = {1,0,0,0, 0,1,0,0, 0,0,1,0, 0,0,0,1};
            std::string path = (fn->str.size() && fn->str[0] == '/') ? fn->str : base + fn->str;
            std::string err; int rc = load_obj_into(*h, path, tp ? tp : ident, &err);
            if (rc < 0) return fs.fail(rc, err);                           // the reference prints "load mesh failed" and goes on
        }
        if (pass == 1 && shaped) {
            const Json* shape = shaped->get("shape"); const Json* sph = shape ? shape->get("Sphere") : nullptr;
            if (!sph) return fs.fail(ARN_E_INVALID, "Shaped component without a Sphere shape");
            float radius, zmin, zmax, phimax;
            if (!json_num(sph->get("radius"), &radius) || !json_num(sph->get("zmin"), &zmin) || !json_num(sph->get("zmax"), &zmax) || !json_num(sph->get("phimax"), &phimax))
                return fs.fail(ARN_E_INVALID, "malformed Sphere");
            // Named<MaterialDesc>::find_or_insert_with (:241-255)
            const Json* mat = shaped->get("material");
            const Json* mname = mat ? mat->get("name") : nullptr; const Json* mval = mat ? mat->get("value") : nullptr;
            if (!mname || mname->kind != Json::Str) return fs.fail(ARN_E_INVALID, "material without a name");
            if (mval && !mval->is_null()) { std::string err; int id = json_material(*mval, *h, base, &err); if (id < 0) return fs.fail(id, err); materials[mname->str] = id; }
            auto it = materials.find(mname->str);
            if (it == materials.end()) return fs.fail(ARN_E_INVALID, "load shape failed: unknown material " + mname->str);
            float emission[3]; const float* ep = nullptr;
            const Json* light = shaped->get("light");
            if (light && !light->is_null()) { std::string err; int rc = json_rgb_texture(light, emission, &err); if (rc != ARN_OK) return fs.fail(rc, err); ep = emission; }
            float t16[16]; const float* tp = nullptr;
            const Json* tr = shaped->get("transform");
            if (tr && !tr->is_null()) { if (!json_matrix(*tr, t16)) return fs.fail(ARN_E_INVALID, "malformed shape transform"); tp = t16; }
            bool kept = false;
            int rc = fs.add_sphere(radius, zmin, zmax, phimax, (uint32_t)it->second, ep, tp, true, &kept);
            if (rc < 0) return rc;
            const Json* nm = c.get("name");
            if (nm && nm->kind == Json::Str) {
                ShapedRecord r; r.radius = radius; r.zmin = zmin; r.zmax = zmax; r.phimax = phimax; r.material = (uint32_t)it->second;
                r.emissive = ep != nullptr; if (ep) { r.emission[0] = ep[0]; r.emission[1] = ep[1]; r.emission[2] = ep[2]; }
                r.transformed = kept;
                shaped_by_name[nm->str] = r;
            }
        }
    }
    // sampler: StrataSampler {sampledx, sampledy, ndim} (sample/strata.rs:93-164)
    const Json* js = root.get("sampler");
    float sx, sy, nd;
    if (!js || !json_num(js->get("sampledx"), &sx) || !json_num(js->get("sampledy"), &sy) || !json_num(js->get("ndim"), &nd)) return fs.fail(ARN_E_INVALID, "malformed sampler");
    sampler->sampledx = (uint32_t)sx; sampler->sampledy = (uint32_t)sy; sampler->ndim = (uint32_t)nd; sampler->seed = 0; sampler->mode = ARN_SAMPLER_PARITY;
    // camera: PerspecCam (filming/perspective.rs:139-260) with its Film (filming/film.rs:38-45)
    const Json* jc = root.get("camera");
    if (!jc) return fs.fail(ARN_E_INVALID, "scene has no camera");
    float pv[16]; const Json* jt = jc->get("transform");
    if (!jt || !json_matrix(*jt, pv)) return fs.fail(ARN_E_INVALID, "malformed camera transform");
    const Json* scr = jc->get("screen"); float smin[2], smax[2];
    if (!scr || !json_vec(scr->get("pmin"), XY, 2, smin) || !json_vec(scr->get("pmax"), XY, 2, smax)) return fs.fail(ARN_E_INVALID, "malformed camera screen");
    float znear, zfar, fov;
    if (!json_num(jc->get("znear"), &znear) || !json_num(jc->get("zfar"), &zfar) || !json_num(jc->get("fov"), &fov)) return fs.fail(ARN_E_INVALID, "malformed camera");
    int has_lens = 0; float lens[2] = {0.f, 0.f};
    const Json* jl = jc->get("lens");
    if (jl && !jl->is_null()) { if (jl->kind != Json::Arr || jl->arr.size() != 2 || !json_num(&jl->arr[0], &lens[0]) || !json_num(&jl->arr[1], &lens[1])) return fs.fail(ARN_E_INVALID, "malformed lens"); has_lens = 1; }
    const Json* jf = jc->get("film");
    float res[2], cmin[2], cmax[2], fr[2];
    const Json* crop = jf ? jf->get("crop_window") : nullptr;
    if (!jf || !json_vec(jf->get("resolution"), XY, 2, res) || !crop || !json_vec(crop->get("pmin"), XY, 2, cmin) || !json_vec(crop->get("pmax"), XY, 2, cmax)
        || !json_vec(jf->get("filter_radius"), XY, 2, fr)) return fs.fail(ARN_E_INVALID, "malformed film");
    film->res_x = (uint32_t)res[0]; film->res_y = (uint32_t)res[1];
    film->crop_min_x = (int32_t)cmin[0]; film->crop_min_y = (int32_t)cmin[1]; film->crop_max_x = (int32_t)cmax[0]; film->crop_max_y = (int32_t)cmax[1];
    film->filter_radius_x = fr[0]; film->filter_radius_y = fr[1];
    film->filter_kind = ARN_FILTER_LANCZOS; film->filter_a = 0.f; film->filter_b = 0.f;   // `filter` is skip_deserializing (film.rs:42,47-51)
    float screen4[4] = {smin[0], smin[1], smax[0], smax[1]};
    int rc = arn_camera_make(pv, screen4, znear, zfar, fov, has_lens, lens[0], lens[1], (float)film->res_x, (float)film->res_y, cam);
    if (rc != ARN_OK) return fs.fail(rc, g_host_error);
    float md;
    if (!json_num(root.get("max_depth"), &md)) return fs.fail(ARN_E_INVALID, "missing max_depth");
    std::memset(params, 0, sizeof *params);
    params->max_depth = (uint32_t)md; params->min_depth = params->max_depth / 2; params->rr_threshold = 0.05f;   // renderer/pt.rs:47-48
    params->tiles_x = 16; params->tiles_y = 16; params->rank = 0; params->world_size = 1;                          // pt.rs:131
    params->spp_begin = 0; params->spp_end = sampler->sampledx * sampler->sampledy; params->partition_subdiv = 0;
    if (outname && outname_cap) { const Json* o = root.get("outputfilename"); std::string s = (o && o->kind == Json::Str) ? o->str : ""; std::snprintf(outname, outname_cap, "%s", s.c_str()); }
    return ARN_OK;
}

// TilePixel::finalize (filming/film.rs:338-344) + ToNorm::from_norm for u8 (spectrum/macros.rs:164-180)
int arn_film_finalize(const float* film, size_t n_pixels, float* rgb_out, uint8_t* rgb8_out) {
    if (!film) return ARN_E_INVALID;
    for (size_t i = 0; i < n_pixels; i++) {
        const float* p = film + 4 * i; float c[3];
        if (p[3] == 0.f) c[0] = c[1] = c[2] = 0.f;
        else { c[0] = p[0] / p[3]; c[1] = p[1] / p[3]; c[2] = p[2] / p[3]; }
        for (int k = 0; k < 3; k++) {
            if (rgb_out) rgb_out[3 * i + k] = c[k];
            if (rgb8_out) { float v = c[k]; v = v < 0.f ? 0.f : (v < 1.f ? v : 1.f); rgb8_out[3 * i + k] = (uint8_t)(v * 255.f); }
        }
    }
    return ARN_OK;
}

// Image::save (filming/film.rs:380-391): RGB8 PNG, rows top to bottom, no gamma.
// Stored-deflate PNG writer (no zlib dependency).
int arn_save_png(const char* path, const float* film, uint32_t width, uint32_t height) {
    if (!path || !film || !width || !height) return ARN_E_INVALID;
    std::vector<uint8_t> rgb((size_t)width * height * 3);
    arn_film_finalize(film, (size_t)width * height, nullptr, rgb.data());
    static uint32_t crc_table[256]; static bool crc_init = false;
    if (!crc_init) { for (uint32_t n = 0; n < 256; n++) { uint32_t c = n; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; crc_table[n] = c; } crc_init = true; }
    auto crc = [&](const uint8_t* d, size_t n, uint32_t c) { for (size_t i = 0; i < n; i++) c = crc_table[(c ^ d[i]) & 0xFF] ^ (c >> 8); return c; };
    std::vector<uint8_t> raw; raw.reserve(((size_t)width * 3 + 1) * height);
    for (uint32_t y = 0; y < height; y++) { raw.push_back(0); raw.insert(raw.end(), rgb.begin() + (size_t)y * width * 3, rgb.begin() + (size_t)(y + 1) * width * 3); }
    std::vector<uint8_t> z; z.push_back(0x78); z.push_back(0x01);
    size_t pos = 0; uint32_t a = 1, b = 0;
    for (uint8_t v : raw) { a = (a + v) % 65521u; b = (b + a) % 65521u; }
    while (pos < raw.size()) {
        size_t n = raw.size() - pos; if (n > 65535) n = 65535;
        z.push_back(pos + n == raw.size() ? 1 : 0);
        z.push_back((uint8_t)(n & 0xFF)); z.push_back((uint8_t)(n >> 8)); z.push_back((uint8_t)(~n & 0xFF)); z.push_back((uint8_t)((~n >> 8) & 0xFF));
        z.insert(z.end(), raw.begin() + (long)pos, raw.begin() + (long)(pos + n)); pos += n;
    }
    uint32_t adler = (b << 16) | a;
    z.push_back((uint8_t)(adler >> 24)); z.push_back((uint8_t)(adler >> 16)); z.push_back((uint8_t)(adler >> 8)); z.push_back((uint8_t)adler);
    FILE* fp = std::fopen(path, "wb");
    if (!fp) return ARN_E_IO;
    auto be32 = [](uint32_t v, uint8_t* o) { o[0] = (uint8_t)(v >> 24); o[1] = (uint8_t)(v >> 16); o[2] = (uint8_t)(v >> 8); o[3] = (uint8_t)v; };
    auto chunk = [&](const char* type, const uint8_t* data, size_t n) {
        uint8_t len[4]; be32((uint32_t)n, len); std::fwrite(len, 1, 4, fp);
        std::fwrite(type, 1, 4, fp); if (n) std::fwrite(data, 1, n, fp);
        uint32_t c = crc((const uint8_t*)type, 4, 0xFFFFFFFFu); if (n) c = crc(data, n, c); c ^= 0xFFFFFFFFu;
        uint8_t cb[4]; be32(c, cb); std::fwrite(cb, 1, 4, fp);
    };
    const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::fwrite(sig, 1, 8, fp);
    uint8_t ihdr[13]; be32(width, ihdr); be32(height, ihdr + 4); ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
    chunk("IHDR", ihdr, 13); chunk("IDAT", z.data(), z.size()); chunk("IEND", nullptr, 0);
    std::fclose(fp);
    return ARN_OK;
}

}  // extern "C"
