// scenekit_io.h — mesh ingestion for the scene kit.
//
//   load_3d   the reference's only mesh format, read by plyInfo (shape/plyRead.h:19-48): whitespace-separated text,
//             "vertex <nV> face <nT>" (either order), nV x "x y z", nT x "<token> i0 i1 i2" (the token, usually the
//             vertex count 3, is read and ignored).  plyInfo multiplies every vertex by 20 while reading (:38); this
//             reader returns the file's values and the caller applies the x20 (add_mesh's pre-scale), so that the float
//             operation is the same one.  Unlike plyInfo it checks what it reads: a short file, a header other than
//             vertex / face or an index out of range is an error, not undefined behaviour.
//   load_obj  Wavefront OBJ (v / vt / vn / f with v, v/vt, v//vn, v/vt/vn corners, negative indices, polygons fanned
//             into triangles).  The reference has no OBJ loader (config 3 names an asset it cannot open): corners are
//             unified into the single-index vertices TriangleMesh takes (shape/Triangle.cpp:14-60).
//   load_mtl  the MTL library an OBJ names (newmtl / Kd / Ks / Ns / Ni / d / Tr / illum) and `material_recipe`, the mapping
//             of an MTL entry onto the reference's material classes (materials/*.cpp): mirror for illum 3 / 5 with a black
//             Kd, glass for illum 4 / 6 / 7 / 9 or d < 1, plastic when Ks is not black, matte otherwise.  The scene kit turns
//             a recipe into a gnx_material, the oracle harness into the reference's own Material object — one recipe, two
//             consumers, so both sides render the same thing.
//   save_3d   writes a Mesh in the .3d layout (tests, tools).
//
// Header-only, plain C++; no dependency on the reference or on CUDA.
#ifndef GNX_SCENEKIT_IO_H
#define GNX_SCENEKIT_IO_H

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <tuple>

#include "scenekit_mesh.h"

namespace gnxsk {

inline bool read_file(const std::string &path, std::string *out, std::string *err) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { *err = "cannot open " + path; return false; }
    std::ostringstream ss;
    ss << f.rdbuf();
    *out = ss.str();
    return true;
}

// Whitespace tokenizer over a buffer (operator>> of the reference's std::fstream splits the same way).
struct Tokens {
    const char *p, *end;
    explicit Tokens(const std::string &s) : p(s.data()), end(s.data() + s.size()) {}
    bool next(const char **b, const char **e) {
        while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r' || *p == '\f' || *p == '\v')) ++p;
        if (p >= end) return false;
        *b = p;
        while (p < end && !(*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r' || *p == '\f' || *p == '\v')) ++p;
        *e = p;
        return true;
    }
    bool word(std::string *w) { const char *b, *e; if (!next(&b, &e)) return false; w->assign(b, e); return true; }
    bool integer(long *v) {
        const char *b, *e; if (!next(&b, &e)) return false;
        char *q; *v = strtol(b, &q, 10); return q == e;
    }
    bool real(float *v) {  // istream >> float and strtof both round the decimal string correctly
        const char *b, *e; if (!next(&b, &e)) return false;
        char *q; *v = strtof(b, &q); return q == e;
    }
};

inline bool load_3d(const std::string &path, Mesh *m, std::string *err) {
    std::string buf;
    if (!read_file(path, &buf, err)) return false;
    Tokens t(buf);
    long nV = -1, nT = -1;
    for (int i = 0; i < 2; ++i) {
        std::string w;
        if (!t.word(&w)) { *err = path + ": truncated header"; return false; }
        if (w == "vertex") { if (!t.integer(&nV)) { *err = path + ": bad vertex count"; return false; } }
        else if (w == "face") { if (!t.integer(&nT)) { *err = path + ": bad face count"; return false; } }
        else { *err = path + ": header must be 'vertex <n> face <m>', found '" + w + "'"; return false; }
    }
    if (nV < 0 || nT < 0 || nV > (1l << 30) || nT > (1l << 30)) { *err = path + ": vertex / face counts missing or out of range"; return false; }
    m->P.resize((size_t)nV * 3);
    m->N.clear(); m->UV.clear();
    m->idx.resize((size_t)nT * 3);
    for (size_t i = 0; i < (size_t)nV * 3; ++i)
        if (!t.real(&m->P[i])) { *err = path + ": truncated or malformed vertex list"; return false; }
    for (long f = 0; f < nT; ++f) {
        std::string w;
        if (!t.word(&w)) { *err = path + ": truncated face list"; return false; }
        for (int c = 0; c < 3; ++c) {
            long v;
            if (!t.integer(&v)) { *err = path + ": truncated or malformed face list"; return false; }
            if (v < 0 || v >= nV) { *err = path + ": face " + std::to_string(f) + " indexes vertex " + std::to_string(v) + " of " + std::to_string(nV); return false; }
            m->idx[(size_t)f * 3 + c] = (int)v;
        }
    }
    return true;
}

inline bool save_3d(const std::string &path, const Mesh &m, std::string *err) {
    FILE *f = fopen(path.c_str(), "w");
    if (!f) { *err = "cannot write " + path; return false; }
    fprintf(f, "vertex %d\nface %d\n", m.nVerts(), m.nTris());
    for (int i = 0; i < m.nVerts(); ++i) fprintf(f, "%.9g %.9g %.9g\n", m.P[3 * i], m.P[3 * i + 1], m.P[3 * i + 2]);
    for (int i = 0; i < m.nTris(); ++i) fprintf(f, "3 %d %d %d\n", m.idx[3 * i], m.idx[3 * i + 1], m.idx[3 * i + 2]);
    fclose(f);
    return true;
}

struct ObjMaterial {
    std::string name;
    float Kd[3] = {0.8f, 0.8f, 0.8f}, Ks[3] = {0, 0, 0};
    float Ns = 0, Ni = 1.5f, d = 1;
    int illum = 2;
};
// What a material of the reference is built from (same constants on the kit side and on the reference side).
struct MaterialRecipe {
    enum Kind { Matte, Plastic, Mirror, Glass } kind = Matte;
    float kd[3] = {0, 0, 0}, ks[3] = {0, 0, 0};
    float roughness = 0, index = 1.5f;
};
inline MaterialRecipe material_recipe(const ObjMaterial &m) {
    MaterialRecipe r;
    const bool kdBlack = m.Kd[0] == 0 && m.Kd[1] == 0 && m.Kd[2] == 0, ksBlack = m.Ks[0] == 0 && m.Ks[1] == 0 && m.Ks[2] == 0;
    for (int c = 0; c < 3; ++c) { r.kd[c] = m.Kd[c]; r.ks[c] = m.Ks[c]; }
    r.index = m.Ni;
    if (m.illum == 4 || m.illum == 6 || m.illum == 7 || m.illum == 9 || m.d < 1) {
        r.kind = MaterialRecipe::Glass;                       // GlassMaterial(Kr = Ks or 1, Kt = 1 - (1 - d) Kd ... kept simple: Kt = Kr)
        for (int c = 0; c < 3; ++c) r.ks[c] = ksBlack ? 1.f : m.Ks[c];
    } else if ((m.illum == 3 || m.illum == 5) && kdBlack && !ksBlack) r.kind = MaterialRecipe::Mirror;   // MirrorMaterial(Kr = Ks)
    else if (!ksBlack) {
        r.kind = MaterialRecipe::Plastic;                     // PlasticMaterial(Kd, Ks, roughness), remapRoughness = false
        r.roughness = std::sqrt(2.f / (m.Ns + 2.f));          // Blinn-Phong exponent -> microfacet alpha
    } else r.kind = MaterialRecipe::Matte;                    // MatteMaterial(Kd, sigma = 0)
    return r;
}
inline bool load_mtl(const std::string &path, std::vector<ObjMaterial> *out, std::string *err) {
    std::string buf;
    if (!read_file(path, &buf, err)) return false;
    std::istringstream in(buf);
    std::string line;
    int lineNo = 0;
    while (std::getline(in, line)) {
        ++lineNo;
        size_t h = line.find('#');
        if (h != std::string::npos) line.resize(h);
        Tokens t(line);
        std::string w;
        if (!t.word(&w)) continue;
        if (w == "newmtl") {
            out->emplace_back();
            if (!t.word(&out->back().name)) { *err = path + ":" + std::to_string(lineNo) + ": newmtl without a name"; return false; }
            continue;
        }
        if (out->empty()) continue;
        ObjMaterial &m = out->back();
        auto three = [&](float *v) {
            if (!t.real(&v[0])) return false;
            if (!t.real(&v[1])) { v[1] = v[2] = v[0]; return true; }  // "Kd r" means grey
            return t.real(&v[2]);
        };
        bool ok = true;
        if (w == "Kd") ok = three(m.Kd);
        else if (w == "Ks") ok = three(m.Ks);
        else if (w == "Ns") ok = t.real(&m.Ns);
        else if (w == "Ni") ok = t.real(&m.Ni);
        else if (w == "d") ok = t.real(&m.d);
        else if (w == "Tr") { float tr; ok = t.real(&tr); m.d = 1 - tr; }
        else if (w == "illum") { float f; ok = t.real(&f); m.illum = (int)f; }
        // Ka / Ke / Tf / map_*: not part of the reference's material models
        if (!ok) { *err = path + ":" + std::to_string(lineNo) + ": malformed '" + w + "'"; return false; }
    }
    return true;
}

// mats (optional): receives the materials of the OBJ's mtllib files; m->tri_material then indexes it per triangle.
inline bool load_obj(const std::string &path, Mesh *m, std::string *err, std::vector<ObjMaterial> *mats = nullptr) {
    std::string buf;
    if (!read_file(path, &buf, err)) return false;
    std::vector<int> faceMat;
    int curMat = -1;
    const std::string dir = path.find_last_of('/') == std::string::npos ? std::string() : path.substr(0, path.find_last_of('/') + 1);
    std::vector<float> P, T, N;
    std::map<std::tuple<int, int, int>, int> corner;  // (v, vt, vn) -> unified vertex
    *m = Mesh();
    bool anyT = false, anyN = false, allT = true, allN = true;
    std::istringstream in(buf);
    std::string line;
    int lineNo = 0;
    auto fail = [&](const std::string &what) { *err = path + ":" + std::to_string(lineNo) + ": " + what; return false; };
    std::vector<std::tuple<int, int, int>> faces;  // corner triples, 3 per triangle
    while (std::getline(in, line)) {
        ++lineNo;
        size_t h = line.find('#');
        if (h != std::string::npos) line.resize(h);
        Tokens t(line);
        std::string w;
        if (!t.word(&w)) continue;
        if (w == "v" || w == "vn") {
            float x[3];
            for (int c = 0; c < 3; ++c) if (!t.real(&x[c])) return fail("malformed '" + w + "'");
            (w == "v" ? P : N).insert((w == "v" ? P : N).end(), x, x + 3);
        } else if (w == "vt") {
            float x[2] = {0, 0};
            if (!t.real(&x[0])) return fail("malformed 'vt'");
            t.real(&x[1]);
            T.insert(T.end(), x, x + 2);
        } else if (w == "f") {
            std::vector<std::tuple<int, int, int>> poly;
            std::string c;
            while (t.word(&c)) {
                int id[3] = {0, 0, 0};  // 1-based, 0 = absent
                const char *q = c.c_str();
                for (int k = 0; k < 3 && *q; ++k) {
                    if (*q != '/') { char *e; id[k] = (int)strtol(q, &e, 10); if (e == q) return fail("malformed face corner '" + c + "'"); q = e; }
                    if (*q == '/') ++q; else break;
                }
                const int cnt[3] = {(int)(P.size() / 3), (int)(T.size() / 2), (int)(N.size() / 3)};
                for (int k = 0; k < 3; ++k) {
                    if (id[k] < 0) id[k] = cnt[k] + id[k] + 1;  // relative to the elements read so far
                    if (id[k] < 0 || id[k] > cnt[k]) return fail("face corner '" + c + "' is out of range");
                }
                if (id[0] == 0) return fail("face corner without a vertex index");
                poly.emplace_back(id[0], id[1], id[2]);
            }
            if (poly.size() < 3) return fail("face with fewer than 3 corners");
            for (size_t k = 1; k + 1 < poly.size(); ++k) {
                faces.push_back(poly[0]); faces.push_back(poly[k]); faces.push_back(poly[k + 1]);
                faceMat.push_back(curMat);
            }
        } else if (w == "mtllib" && mats) {
            std::string f;
            // (a library that is not there is not an error: OBJ files routinely outlive their MTL; the faces then keep the
            // scene's default material.  A library that IS there must parse.)
            while (t.word(&f)) {
                FILE *probe = fopen((dir + f).c_str(), "rb");
                if (!probe) continue;
                fclose(probe);
                if (!load_mtl(dir + f, mats, err)) return false;
            }
        } else if (w == "usemtl" && mats) {
            std::string nm;
            t.word(&nm);
            curMat = -1;
            for (size_t k = 0; k < mats->size(); ++k) if ((*mats)[k].name == nm) curMat = (int)k;
            if (curMat < 0 && !mats->empty()) return fail("usemtl names an unknown material '" + nm + "'");
        }  // o / g / s: ignored
    }
    if (faces.empty()) return fail("no faces");
    for (auto &c : faces) {
        anyT |= std::get<1>(c) != 0; anyN |= std::get<2>(c) != 0;
        allT &= std::get<1>(c) != 0; allN &= std::get<2>(c) != 0;
    }
    // TriangleMesh takes uv / n for every vertex or for none
    const bool useT = anyT && allT, useN = anyN && allN;
    for (auto c : faces) {
        if (!useT) std::get<1>(c) = 0;
        if (!useN) std::get<2>(c) = 0;
        auto it = corner.find(c);
        if (it == corner.end()) {
            int v = m->nVerts();
            it = corner.emplace(c, v).first;
            const int iv = std::get<0>(c) - 1, it2 = std::get<1>(c) - 1, in = std::get<2>(c) - 1;
            m->P.insert(m->P.end(), {P[3 * iv], P[3 * iv + 1], P[3 * iv + 2]});
            if (useT) m->UV.insert(m->UV.end(), {T[2 * it2], T[2 * it2 + 1]});
            if (useN) m->N.insert(m->N.end(), {N[3 * in], N[3 * in + 1], N[3 * in + 2]});
        }
        m->idx.push_back(it->second);
    }
    if (mats && !mats->empty()) m->tri_material = faceMat;
    return true;
}

// Uniform scale + translation that fits a mesh into the sphere of radius `radius` around `centre` (OBJ assets come in
// arbitrary units; the .3d path keeps the reference's fixed x20 and (0, -2.9, 0) instead).
inline void fit_to_sphere(Mesh *m, float radius, const float centre[3]) {
    if (m->P.empty()) return;
    float lo[3] = {m->P[0], m->P[1], m->P[2]}, hi[3] = {m->P[0], m->P[1], m->P[2]};
    for (int i = 0; i < m->nVerts(); ++i)
        for (int c = 0; c < 3; ++c) { lo[c] = std::min(lo[c], m->P[3 * i + c]); hi[c] = std::max(hi[c], m->P[3 * i + c]); }
    float mid[3], r2 = 0;
    for (int c = 0; c < 3; ++c) mid[c] = 0.5f * (lo[c] + hi[c]);
    for (int i = 0; i < m->nVerts(); ++i) {
        float d2 = 0;
        for (int c = 0; c < 3; ++c) { float d = m->P[3 * i + c] - mid[c]; d2 += d * d; }
        r2 = std::max(r2, d2);
    }
    const float s = r2 > 0 ? radius / std::sqrt(r2) : 1.f;
    for (int i = 0; i < m->nVerts(); ++i)
        for (int c = 0; c < 3; ++c) m->P[3 * i + c] = (m->P[3 * i + c] - mid[c]) * s + centre[c];
}

}  // namespace gnxsk
#endif
