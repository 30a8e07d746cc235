"""Generates tests/golden/cornell_scene.npz from the reference's Cornell fixture
(/root/reference/examples/cornellbox/{cb.json,CornellBox-Glossy.obj,.mtl}).

Run in the build container (the reference tree does not exist on the GPU box):
    python tests/golden/make_cornell_fixture.py

The OBJ/MTL are parsed here by an independent Python restatement of tobj 0.1's semantics
(one model per g/o group, vertices re-indexed per unique (v,vt,vn) in first-use order) and of
component::load_obj's material choice (src/component/mod.rs:65-185); tests/test_scene_ingest.py
checks the C++ host loader against this fixture.  Stored: raw (untransformed) model arrays,
the material table, the mesh transform, the two spheres, camera and integrator parameters.
"""
import json
import os
import sys

import numpy as np

REF = "/root/reference/examples/cornellbox"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cornell_scene.npz")


def parse_mtl(path):
    mats, cur = [], None
    for line in open(path):
        l = line.strip()
        if not l or l.startswith("#"):
            continue
        w = l.split()
        k, rest = w[0], l[len(w[0]):].strip()
        if k == "newmtl":
            cur = dict(name=rest, Kd=[0, 0, 0], Ks=[0, 0, 0], Ns=0.0, Ni=1.0, d=1.0, unknown={})
            mats.append(cur)
        elif k in ("Kd", "Ks"):
            cur[k] = [float(x) for x in w[1:4]]
        elif k in ("Ns", "Ni", "d"):
            cur[k] = float(w[1])
        elif k in ("Ka", "map_Ka", "map_Kd", "map_Ks", "map_Ns", "map_d"):
            pass
        else:
            cur["unknown"][k] = rest
    return mats


def parse_obj(path):
    pos, tex, nrm, models = [], [], [], []
    faces, name, mat = [], "unnamed_object", None
    mats, mat_map = [], {}

    def export():
        nonlocal faces
        if not faces:
            return
        imap, P, T, N, I = {}, [], [], [], []
        for f in faces:
            for k in range(1, len(f) - 1):
                for vi in (f[0], f[k], f[k + 1]):
                    if vi not in imap:
                        imap[vi] = len(P)
                        P.append(pos[vi[0]])
                        if vi[1] is not None:
                            T.append(tex[vi[1]])
                        if vi[2] is not None:
                            N.append(nrm[vi[2]])
                    I.append(imap[vi])
        models.append(dict(name=name, positions=np.array(P, np.float32), texcoords=np.array(T, np.float32),
                           normals=np.array(N, np.float32), indices=np.array(I, np.uint32), material=mat))
        faces = []

    def idx(tok, n):
        i = int(tok)
        return n + i if i < 0 else i - 1

    for line in open(path):
        l = line.strip()
        if not l or l.startswith("#"):
            continue
        w = l.split()
        if w[0] == "v":
            pos.append([float(x) for x in w[1:4]])
        elif w[0] == "vt":
            tex.append([float(x) for x in w[1:3]])
        elif w[0] == "vn":
            nrm.append([float(x) for x in w[1:4]])
        elif w[0] == "f":
            f = []
            for tok in w[1:]:
                p = tok.split("/")
                v = idx(p[0], len(pos))
                vt = idx(p[1], len(tex)) if len(p) > 1 and p[1] else None
                vn = idx(p[2], len(nrm)) if len(p) > 2 and p[2] else None
                f.append((v, vt, vn))
            faces.append(f)
        elif w[0] in ("o", "g"):
            export()
            name = l[1:].strip()
        elif w[0] == "mtllib":
            for m in parse_mtl(os.path.join(os.path.dirname(path), l[6:].strip())):
                mat_map[m["name"]] = len(mats)
                mats.append(m)
        elif w[0] == "usemtl":
            new = mat_map.get(l[6:].strip())
            if new != mat and faces:
                export()
            mat = new
    export()
    return models, mats


def material_row(m):
    """component::load_obj's choice (mod.rs:118-164) -> (type, kd3, ks3, sigma, roughness, eta, dissolve)"""
    f32 = np.float32
    rough = f32(min(max((f32(1000.0) - f32(m["Ns"])) / f32(1000.0), 0.0), 1.0)) if True else 0
    rough = f32(max(min(float((f32(1000.0) - f32(m["Ns"])) / f32(1000.0)), 1.0), 0.0))
    illum = m["unknown"].get("illum", "2")
    dissolve = f32(min(max(m["d"], 0.0), 1.0))
    kd, ks = m["Kd"], m["Ks"]
    if "4" in illum:
        t, eta, dis = 2, m["Ni"], 1.0
    elif abs(float(dissolve) - 1.0) > 1.1920929e-7:
        t, eta, dis = 3, 1.0, float(dissolve)
    elif ks == [0, 0, 0]:
        t, eta, dis = 0, 1.0, 1.0
    else:
        t, eta, dis = 1, 1.0, 1.0
    return [t] + kd + ks + [0.0, float(rough), eta, dis]


def matrix(j):
    cols = [j[k] for k in "xyzw"] if isinstance(j, dict) else j
    return np.array(cols, np.float32).reshape(16)


def main():
    scene = json.load(open(os.path.join(REF, "cb.json")))
    out = {}
    mesh = [c for c in scene["components"] if "Mesh" in c["value"]][0]["value"]["Mesh"]
    models, mats = parse_obj(os.path.join(REF, os.path.basename(mesh["filename"])))
    rows = [material_row(m) for m in mats]
    rows.append([0, 0.5, 0.6, 0.7, 0, 0, 0, 0.0, 0.0, 1.0, 1.0])        # fallback Matte(0.5,0.6,0.7), mod.rs:165-171
    out["n_models"] = np.int32(len(models))
    out["model_material"] = np.array([m["material"] if m["material"] is not None else len(rows) - 1 for m in models], np.int32)
    for i, m in enumerate(models):
        out[f"m{i}_positions"] = m["positions"]
        out[f"m{i}_indices"] = m["indices"]
        if len(m["normals"]):
            assert len(m["normals"]) == len(m["positions"])
            out[f"m{i}_normals"] = m["normals"]
        if len(m["texcoords"]):
            assert len(m["texcoords"]) == len(m["positions"])
            out[f"m{i}_texcoords"] = m["texcoords"]
    out["model_names"] = np.array([m["name"] for m in models])
    out["mesh_transform"] = matrix(mesh["transform"])
    # shaped primitives, file order; materials by name (Named<T>::find_or_insert_with)
    sph_rows, named = [], {}
    for c in scene["components"]:
        s = c["value"].get("Shaped")
        if not s:
            continue
        md = s["material"]
        if md.get("value"):
            mat = md["value"]["Matte"]
            named[md["name"]] = len(rows)
            rows.append([0] + mat["kd"]["value"]["Constant"]["value"]["inner"] + [0, 0, 0] + [mat["sigma"]["value"]["Constant"]["value"], 0.0, 1.0, 1.0])
        sp = s["shape"]["Sphere"]
        em = s["light"]["value"]["Constant"]["value"]["inner"]
        sph_rows.append([sp["radius"], sp["zmin"], sp["zmax"], sp["phimax"], named[md["name"]]] + em + list(matrix(s["transform"])))
    out["materials"] = np.array(rows, np.float32)
    out["spheres"] = np.array(sph_rows, np.float32)
    cam = scene["camera"]
    out["camera"] = np.array(list(matrix(cam["transform"])) + [cam["screen"]["pmin"]["x"], cam["screen"]["pmin"]["y"], cam["screen"]["pmax"]["x"],
                             cam["screen"]["pmax"]["y"], cam["znear"], cam["zfar"], cam["fov"]], np.float32)
    out["film"] = np.array([cam["film"]["resolution"]["x"], cam["film"]["resolution"]["y"], cam["film"]["filter_radius"]["x"], cam["film"]["filter_radius"]["y"]], np.float32)
    out["sampler"] = np.array([scene["sampler"]["sampledx"], scene["sampler"]["sampledy"], scene["sampler"]["ndim"]], np.int32)
    out["max_depth"] = np.int32(scene["max_depth"])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "models:", [(m["name"], len(m["indices"]) // 3) for m in models], "materials:", len(rows), file=sys.stderr)


if __name__ == "__main__":
    main()
