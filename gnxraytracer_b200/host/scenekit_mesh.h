// scenekit_mesh.h — deterministic procedural meshes for the BASELINE.json configs.
//
// The reference snapshot ships neither dragon.3d nor nanosuit (/.MISSING_LARGE_BLOBS) and its
// Sphere is a stub (shape/Sphere.h:28-56), so every config is realised with the stand-ins that
// SURVEY.md §8(d) specifies.  The same generators feed (a) the product's own scene builder
// (scenekit.cpp) and (b) the oracle harness, which pushes the very same vertex arrays through the
// reference's TriangleMesh class, so both sides render identical geometry.
//
// Header-only, no dependency on the reference or on CUDA.
#ifndef GNX_SCENEKIT_MESH_H
#define GNX_SCENEKIT_MESH_H

#include <array>
#include <cmath>
#include <cstdint>
#include <map>
#include <utility>
#include <vector>

namespace gnxsk {

struct Mesh {
    std::vector<float> P;    // 3 * nVerts
    std::vector<float> N;    // 3 * nVerts or empty
    std::vector<float> UV;   // 2 * nVerts or empty
    std::vector<int> idx;    // 3 * nTris
    std::vector<int> tri_material;  // nTris indices into the file's material list (OBJ usemtl), -1 = none; empty = one material
    int nVerts() const { return (int)(P.size() / 3); }
    int nTris() const { return (int)(idx.size() / 3); }
};

// The five walls of the UI's Cornell box (ui/ModelList.cpp:71-125): a cube of edge L with its
// min corner at the origin, 2 triangles per wall, unshared vertices, in the order
// floor, ceiling, back, right(x=0), left(x=L).  The caller translates by -L/2 (ModelList.cpp:109).
inline Mesh cornell_walls(float L = 5.0f) {
    const float o = 0.f;
    const float v[30][3] = {
        {o, o, L}, {L, o, L}, {o, o, o},  {L, o, L}, {L, o, o}, {o, o, o},   // floor
        {o, L, L}, {o, L, o}, {L, L, L},  {L, L, L}, {o, L, o}, {L, L, o},   // ceiling
        {o, o, o}, {L, o, o}, {L, L, o},  {o, o, o}, {L, L, o}, {o, L, o},   // back wall
        {o, o, o}, {o, L, L}, {o, o, L},  {o, o, o}, {o, L, o}, {o, L, L},   // x = 0 wall
        {L, o, o}, {L, L, L}, {L, o, L},  {L, o, o}, {L, L, o}, {L, L, L}};  // x = L wall
    Mesh m;
    for (int i = 0; i < 30; ++i) {
        m.P.insert(m.P.end(), {v[i][0], v[i][1], v[i][2]});
        m.idx.push_back(i);
    }
    return m;
}

// The 2.8 x 2.8 ceiling light quad (ui/ModelList.cpp:127-146), y = 0 before the caller's translate.
inline Mesh area_light_quad(float h = 1.4f) {
    const float v[6][3] = {{-h, 0, h}, {-h, 0, -h}, {h, 0, h}, {h, 0, h}, {-h, 0, -h}, {h, 0, -h}};
    Mesh m;
    for (int i = 0; i < 6; ++i) {
        m.P.insert(m.P.end(), {v[i][0], v[i][1], v[i][2]});
        m.idx.push_back(i);
    }
    return m;
}

// Axis-aligned quad in the plane y = y0 spanning [-h, h]^2 (ground plane for config 4).
inline Mesh ground_quad(float h, float y0) {
    const float v[4][3] = {{-h, y0, h}, {h, y0, h}, {h, y0, -h}, {-h, y0, -h}};
    Mesh m;
    for (int i = 0; i < 4; ++i) m.P.insert(m.P.end(), {v[i][0], v[i][1], v[i][2]});
    m.idx = {0, 1, 2, 0, 2, 3};
    return m;
}

// Closed axis-aligned box, outward-facing triangles (12), used as a medium boundary.
inline Mesh box(const float lo[3], const float hi[3]) {
    Mesh m;
    for (int i = 0; i < 8; ++i)
        m.P.insert(m.P.end(), {(i & 1) ? hi[0] : lo[0], (i & 2) ? hi[1] : lo[1], (i & 4) ? hi[2] : lo[2]});
    const int f[12][3] = {{0, 2, 1}, {1, 2, 3}, {4, 5, 6}, {5, 7, 6}, {0, 1, 4}, {1, 5, 4},
                          {2, 6, 3}, {3, 6, 7}, {0, 4, 2}, {2, 4, 6}, {1, 3, 5}, {3, 7, 5}};
    for (auto &t : f) m.idx.insert(m.idx.end(), {t[0], t[1], t[2]});
    return m;
}

// Icosphere: `subdiv` 4-way subdivisions of an icosahedron (subdiv 3 -> 1280 triangles),
// outward-facing, shared vertices, optional per-vertex normals.
inline Mesh icosphere(int subdiv, float radius, float cx, float cy, float cz, bool withNormals = false) {
    const double t = (1.0 + std::sqrt(5.0)) / 2.0;
    std::vector<std::array<double, 3>> V = {{-1, t, 0}, {1, t, 0}, {-1, -t, 0}, {1, -t, 0},
                                            {0, -1, t}, {0, 1, t}, {0, -1, -t}, {0, 1, -t},
                                            {t, 0, -1}, {t, 0, 1}, {-t, 0, -1}, {-t, 0, 1}};
    std::vector<std::array<int, 3>> F = {{0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11},
                                         {1, 5, 9}, {5, 11, 4}, {11, 10, 2}, {10, 7, 6}, {7, 1, 8},
                                         {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8}, {3, 8, 9},
                                         {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};
    auto norm = [](std::array<double, 3> &p) {
        double l = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
        p[0] /= l; p[1] /= l; p[2] /= l;
    };
    for (auto &p : V) norm(p);
    for (int s = 0; s < subdiv; ++s) {
        std::map<std::pair<int, int>, int> mid;
        auto midpoint = [&](int a, int b) {
            std::pair<int, int> key(std::min(a, b), std::max(a, b));
            auto it = mid.find(key);
            if (it != mid.end()) return it->second;
            std::array<double, 3> p = {(V[a][0] + V[b][0]) / 2, (V[a][1] + V[b][1]) / 2, (V[a][2] + V[b][2]) / 2};
            norm(p);
            V.push_back(p);
            return mid[key] = (int)V.size() - 1;
        };
        std::vector<std::array<int, 3>> F2;
        F2.reserve(F.size() * 4);
        for (auto &f : F) {
            int a = midpoint(f[0], f[1]), b = midpoint(f[1], f[2]), c = midpoint(f[2], f[0]);
            F2.push_back({f[0], a, c});
            F2.push_back({f[1], b, a});
            F2.push_back({f[2], c, b});
            F2.push_back({a, b, c});
        }
        F.swap(F2);
    }
    Mesh m;
    for (auto &p : V) {
        m.P.insert(m.P.end(), {(float)(cx + radius * p[0]), (float)(cy + radius * p[1]), (float)(cz + radius * p[2])});
        if (withNormals) m.N.insert(m.N.end(), {(float)p[0], (float)p[1], (float)p[2]});
    }
    for (auto &f : F) m.idx.insert(m.idx.end(), {f[0], f[1], f[2]});
    return m;
}

// "Dragon-class" stand-in for the missing Resources/dragon.3d (Stanford dragon: 871 414 tris):
// a closed (2,3) torus-knot tube, nu x nv quads -> 2*nu*nv triangles (2048 x 213 -> 872 448),
// tube radius modulated by a fixed 3-term sine displacement (no RNG).  Coordinates are in the
// units of the .3d file, i.e. BEFORE plyInfo's x20 scale (shape/plyRead.h:38) and AddModel's
// translate (0,-2.9,0) (ui/ModelList.cpp:56); `scale` folds those in when the caller wants
// world units directly.
inline Mesh torus_knot(int nu = 2048, int nv = 213, float scale = 1.0f, bool withNormals = false,
                       bool withUV = false) {
    const double PI = 3.14159265358979323846;
    const int p = 2, q = 3;
    const double R = 0.062, r = 0.026, tube = 0.0135;
    auto center = [&](double u, double c[3]) {
        double phi = u * 2 * PI;
        double rr = R + r * std::cos(q * phi);
        c[0] = rr * std::cos(p * phi);
        c[1] = rr * std::sin(p * phi) + 0.145;  // 0.145 * 20 = 2.9: centred after the UI's translate
        c[2] = r * std::sin(q * phi) * 1.35;    // knot lies in the xy plane, facing the UI camera
    };
    Mesh m;
    m.P.resize((size_t)nu * nv * 3);
    if (withNormals) m.N.resize((size_t)nu * nv * 3);
    if (withUV) m.UV.resize((size_t)nu * nv * 2);
    for (int i = 0; i < nu; ++i) {
        double u = (double)i / nu;
        double c0[3], c1[3];
        center(u, c0);
        center(u + 1e-4, c1);
        double T[3] = {c1[0] - c0[0], c1[1] - c0[1], c1[2] - c0[2]};
        double tl = std::sqrt(T[0] * T[0] + T[1] * T[1] + T[2] * T[2]);
        for (double &x : T) x /= tl;
        // frame: B = T x up', Nn = B x T, with a reference vector pointing away from the knot axis
        double ref[3] = {c0[0], c0[1] - 0.145, 0.0};
        double rl = std::sqrt(ref[0] * ref[0] + ref[1] * ref[1]);
        ref[0] /= rl; ref[1] /= rl;
        double B[3] = {T[1] * ref[2] - T[2] * ref[1], T[2] * ref[0] - T[0] * ref[2], T[0] * ref[1] - T[1] * ref[0]};
        double bl = std::sqrt(B[0] * B[0] + B[1] * B[1] + B[2] * B[2]);
        for (double &x : B) x /= bl;
        double Nn[3] = {B[1] * T[2] - B[2] * T[1], B[2] * T[0] - B[0] * T[2], B[0] * T[1] - B[1] * T[0]};
        for (int j = 0; j < nv; ++j) {
            double v = (double)j / nv, th = v * 2 * PI;
            double disp = 1.0 + 0.11 * std::sin(17 * u * 2 * PI + 3 * th) + 0.07 * std::sin(41 * u * 2 * PI - 5 * th) +
                          0.04 * std::sin(7 * th + 97 * u * 2 * PI);
            double rad = tube * disp;
            double d[3] = {std::cos(th) * Nn[0] + std::sin(th) * B[0], std::cos(th) * Nn[1] + std::sin(th) * B[1],
                           std::cos(th) * Nn[2] + std::sin(th) * B[2]};
            size_t k = (size_t)i * nv + j;
            for (int a = 0; a < 3; ++a) m.P[3 * k + a] = (float)(scale * (c0[a] + rad * d[a]));
            if (withNormals) for (int a = 0; a < 3; ++a) m.N[3 * k + a] = (float)d[a];
            if (withUV) { m.UV[2 * k] = (float)(u * 16.0); m.UV[2 * k + 1] = (float)v; }
        }
    }
    m.idx.reserve((size_t)nu * nv * 6);
    for (int i = 0; i < nu; ++i) {
        int i1 = (i + 1) % nu;
        for (int j = 0; j < nv; ++j) {
            int j1 = (j + 1) % nv;
            int a = i * nv + j, b = i1 * nv + j, c = i1 * nv + j1, d = i * nv + j1;
            m.idx.insert(m.idx.end(), {a, b, c, a, c, d});
        }
    }
    return m;
}

// UV sphere with per-vertex normals and UVs (config 3's textured stand-in building block).
inline Mesh uv_sphere(int nu, int nv, float radius, float cx, float cy, float cz) {
    const double PI = 3.14159265358979323846;
    Mesh m;
    for (int j = 0; j <= nv; ++j) {
        double v = (double)j / nv, th = v * PI;
        for (int i = 0; i <= nu; ++i) {
            double u = (double)i / nu, ph = u * 2 * PI;
            double d[3] = {std::sin(th) * std::cos(ph), std::cos(th), std::sin(th) * std::sin(ph)};
            m.P.insert(m.P.end(), {(float)(cx + radius * d[0]), (float)(cy + radius * d[1]), (float)(cz + radius * d[2])});
            m.N.insert(m.N.end(), {(float)d[0], (float)d[1], (float)d[2]});
            m.UV.insert(m.UV.end(), {(float)(u * 4.0), (float)(v * 2.0)});
        }
    }
    for (int j = 0; j < nv; ++j)
        for (int i = 0; i < nu; ++i) {
            int a = j * (nu + 1) + i, b = a + 1, c = a + nu + 1, d = c + 1;
            if (j != 0) m.idx.insert(m.idx.end(), {a, b, c});
            if (j != nv - 1) m.idx.insert(m.idx.end(), {b, d, c});
        }
    return m;
}

inline void translate(Mesh &m, float x, float y, float z) {
    for (size_t i = 0; i < m.P.size(); i += 3) { m.P[i] += x; m.P[i + 1] += y; m.P[i + 2] += z; }
}

}  // namespace gnxsk
#endif
