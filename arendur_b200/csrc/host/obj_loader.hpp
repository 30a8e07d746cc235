// OBJ / MTL ingest with the semantics of the `tobj` 0.1 crate that arendur's `load_obj`
// relies on (src/component/mod.rs:65-185): one model per `o`/`g` group in file order,
// vertices re-indexed per unique (v, vt, vn) triple in first-use order, polygons fanned
// into triangles, one material id per model, unknown MTL keys kept verbatim
// (`illum`, `map_bump`, ...).  tobj's source is not part of the reference tree; this
// follows its documented behaviour.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

namespace arnhost {

struct ObjMaterial {
    std::string name;
    float ambient[3] = {0, 0, 0}, diffuse[3] = {0, 0, 0}, specular[3] = {0, 0, 0};
    float shininess = 0.f, dissolve = 1.f, optical_density = 1.f;
    std::string ambient_texture, diffuse_texture, specular_texture, normal_texture, dissolve_texture;
    std::map<std::string, std::string> unknown_param;
};
struct ObjModel {
    std::string name;
    std::vector<float> positions, normals, texcoords;
    std::vector<uint32_t> indices;
    int material_id = -1;
};

namespace objdetail {
inline std::string trim(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r\n"); if (a == std::string::npos) return "";
    size_t b = s.find_last_not_of(" \t\r\n"); return s.substr(a, b - a + 1);
}
inline std::vector<std::string> words(const std::string& s) {
    std::vector<std::string> w; std::istringstream is(s); std::string t; while (is >> t) w.push_back(t); return w;
}
inline bool floats(const std::vector<std::string>& w, size_t first, int n, float* out) {
    if (w.size() < first + (size_t)n) return false;
    for (int i = 0; i < n; i++) { char* e = nullptr; out[i] = std::strtof(w[first + i].c_str(), &e); if (e == w[first + i].c_str()) return false; }
    return true;
}
struct VertexIndices { long v, vt, vn; bool operator<(const VertexIndices& o) const { return std::tie(v, vt, vn) < std::tie(o.v, o.vt, o.vn); } };
const long MISSING = -1;
inline bool parse_index(const std::string& tok, size_t count, long* out) {
    char* e = nullptr; long i = std::strtol(tok.c_str(), &e, 10);
    if (e == tok.c_str()) return false;
    if (i < 0) *out = (long)count + i; else *out = i - 1;      // relative / 1-based
    return true;
}
inline bool parse_face_vertex(const std::string& tok, size_t np, size_t nt, size_t nn, VertexIndices* out) {
    out->v = out->vt = out->vn = MISSING;
    size_t s1 = tok.find('/');
    if (s1 == std::string::npos) return parse_index(tok, np, &out->v);
    if (!parse_index(tok.substr(0, s1), np, &out->v)) return false;
    size_t s2 = tok.find('/', s1 + 1);
    std::string t = s2 == std::string::npos ? tok.substr(s1 + 1) : tok.substr(s1 + 1, s2 - s1 - 1);
    if (!t.empty() && !parse_index(t, nt, &out->vt)) return false;
    if (s2 != std::string::npos) { std::string n = tok.substr(s2 + 1); if (!n.empty() && !parse_index(n, nn, &out->vn)) return false; }
    return true;
}
}  // namespace objdetail

inline bool load_mtl(const std::string& path, std::vector<ObjMaterial>& mats, std::map<std::string, int>& mat_map, std::string* err) {
    using namespace objdetail;
    std::ifstream f(path);
    if (!f) { if (err) *err = "cannot open MTL file " + path; return false; }
    ObjMaterial cur; bool have = false;
    std::string line;
    auto flush = [&]() { if (have) { mat_map[cur.name] = (int)mats.size(); mats.push_back(cur); } };
    while (std::getline(f, line)) {
        std::string l = trim(line);
        if (l.empty() || l[0] == '#') continue;
        std::vector<std::string> w = words(l);
        const std::string& key = w[0];
        std::string rest = trim(l.substr(key.size()));
        if (key == "newmtl") { flush(); cur = ObjMaterial(); cur.name = rest; have = true; }
        else if (key == "Ka") floats(w, 1, 3, cur.ambient);
        else if (key == "Kd") floats(w, 1, 3, cur.diffuse);
        else if (key == "Ks") floats(w, 1, 3, cur.specular);
        else if (key == "Ns") floats(w, 1, 1, &cur.shininess);
        else if (key == "Ni") floats(w, 1, 1, &cur.optical_density);
        else if (key == "d") floats(w, 1, 1, &cur.dissolve);
        else if (key == "map_Ka") cur.ambient_texture = rest;
        else if (key == "map_Kd") cur.diffuse_texture = rest;
        else if (key == "map_Ks") cur.specular_texture = rest;
        else if (key == "map_Ns") cur.normal_texture = rest;
        else if (key == "map_d") cur.dissolve_texture = rest;
        else cur.unknown_param[key] = rest;
    }
    flush();
    return true;
}

inline bool load_obj_file(const std::string& path, std::vector<ObjModel>& models, std::vector<ObjMaterial>& mats, std::string* err) {
    using namespace objdetail;
    std::ifstream f(path);
    if (!f) { if (err) *err = "cannot open OBJ file " + path; return false; }
    std::string dir; { size_t p = path.find_last_of('/'); if (p != std::string::npos) dir = path.substr(0, p + 1); }
    std::vector<float> pos, tex, nrm;
    std::vector<std::vector<VertexIndices>> faces;
    std::string name = "unnamed_object";
    std::map<std::string, int> mat_map; int mat_id = -1;
    auto export_faces = [&]() {
        if (faces.empty()) return;
        ObjModel m; m.name = name; m.material_id = mat_id;
        std::map<VertexIndices, uint32_t> index_map;
        auto add_vertex = [&](const VertexIndices& vi) {
            auto it = index_map.find(vi);
            if (it != index_map.end()) { m.indices.push_back(it->second); return; }
            uint32_t ni = (uint32_t)(m.positions.size() / 3);
            m.positions.push_back(pos[3 * vi.v]); m.positions.push_back(pos[3 * vi.v + 1]); m.positions.push_back(pos[3 * vi.v + 2]);
            if (!tex.empty() && vi.vt != MISSING) { m.texcoords.push_back(tex[2 * vi.vt]); m.texcoords.push_back(tex[2 * vi.vt + 1]); }
            if (!nrm.empty() && vi.vn != MISSING) { m.normals.push_back(nrm[3 * vi.vn]); m.normals.push_back(nrm[3 * vi.vn + 1]); m.normals.push_back(nrm[3 * vi.vn + 2]); }
            index_map[vi] = ni; m.indices.push_back(ni);
        };
        for (auto& face : faces) {
            if (face.size() < 3) continue;                       // points / lines are not renderable components
            for (size_t k = 1; k + 1 < face.size(); k++) { add_vertex(face[0]); add_vertex(face[k]); add_vertex(face[k + 1]); }
        }
        models.push_back(std::move(m));
        faces.clear();
    };
    std::string line;
    while (std::getline(f, line)) {
        std::string l = trim(line);
        if (l.empty() || l[0] == '#') continue;
        std::vector<std::string> w = words(l);
        const std::string& key = w[0];
        if (key == "v") { float v[3]; if (!floats(w, 1, 3, v)) { if (err) *err = "bad position: " + l; return false; } pos.insert(pos.end(), v, v + 3); }
        else if (key == "vt") { float v[2]; if (!floats(w, 1, 2, v)) { if (err) *err = "bad texcoord: " + l; return false; } tex.insert(tex.end(), v, v + 2); }
        else if (key == "vn") { float v[3]; if (!floats(w, 1, 3, v)) { if (err) *err = "bad normal: " + l; return false; } nrm.insert(nrm.end(), v, v + 3); }
        else if (key == "f" || key == "l") {
            std::vector<VertexIndices> face;
            for (size_t k = 1; k < w.size(); k++) {
                VertexIndices vi;
                if (!parse_face_vertex(w[k], pos.size() / 3, tex.size() / 2, nrm.size() / 3, &vi) || vi.v < 0 || (size_t)vi.v >= pos.size() / 3
                    || (vi.vt != MISSING && (vi.vt < 0 || (size_t)vi.vt >= tex.size() / 2)) || (vi.vn != MISSING && (vi.vn < 0 || (size_t)vi.vn >= nrm.size() / 3))) {
                    if (err) *err = "bad face: " + l; return false; }
                face.push_back(vi);
            }
            faces.push_back(std::move(face));
        }
        else if (key == "o" || key == "g") { export_faces(); name = trim(l.substr(1)); if (name.empty()) name = "unnamed_object"; }
        else if (key == "mtllib") {
            std::string mtl = trim(l.substr(6));
            std::string e2;
            if (!load_mtl(dir + mtl, mats, mat_map, &e2)) { if (err) *err = e2; return false; }
        }
        else if (key == "usemtl") {
            std::string mname = trim(l.substr(6));
            auto it = mat_map.find(mname);
            int new_mat = it == mat_map.end() ? -1 : it->second;
            if (new_mat != mat_id && !faces.empty()) export_faces();   // a material change inside a group emits a model
            mat_id = new_mat;
        }
        // everything else (s, comments handled above) is ignored
    }
    export_faces();
    return true;
}

}  // namespace arnhost
