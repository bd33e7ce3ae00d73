// scenekit.cpp — host-side scene kit (include/gnx_scenekit.h): builds the BASELINE.json configs
// directly as gnx_scene_desc buffers.  Plain C++ (g++), no CUDA, no dependency on the reference or on
// the oracle.  tests/test_scenekit.py and tests/test_gpu_parity.py check that these scenes render like the ones the
// bridge flattens out of the reference's own objects.
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "gnx_scenekit.h"
#include "scenekit_mesh.h"
#include "scenekit_io.h"

namespace {

constexpr float kPiF = 3.14159265358979323846f;

struct Vec3d { double x, y, z; };
static Vec3d sub(Vec3d a, Vec3d b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static Vec3d crs(Vec3d a, Vec3d b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static Vec3d nrm(Vec3d a) { double l = std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); return {a.x / l, a.y / l, a.z / l}; }

struct Mat4 {
    double m[4][4];
    static Mat4 identity() { Mat4 r{}; for (int i = 0; i < 4; ++i) r.m[i][i] = 1; return r; }
};
static Mat4 mul(const Mat4 &a, const Mat4 &b) {
    Mat4 r{};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            for (int k = 0; k < 4; ++k) r.m[i][j] += a.m[i][k] * b.m[k][j];
    return r;
}
static Mat4 inverse(const Mat4 &a) {  // Gauss-Jordan with partial pivoting, in double
    double aug[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) { aug[i][j] = a.m[i][j]; aug[i][4 + j] = i == j; }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        for (int r = c + 1; r < 4; ++r) if (std::fabs(aug[r][c]) > std::fabs(aug[piv][c])) piv = r;
        for (int j = 0; j < 8; ++j) std::swap(aug[c][j], aug[piv][j]);
        double d = aug[c][c];
        for (int j = 0; j < 8; ++j) aug[c][j] /= d;
        for (int r = 0; r < 4; ++r) {
            if (r == c) continue;
            double f = aug[r][c];
            for (int j = 0; j < 8; ++j) aug[r][j] -= f * aug[c][j];
        }
    }
    Mat4 r{};
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = aug[i][4 + j];
    return r;
}
static Mat4 scale(double x, double y, double z) { Mat4 r = Mat4::identity(); r.m[0][0] = x; r.m[1][1] = y; r.m[2][2] = z; return r; }
static Mat4 translate(double x, double y, double z) { Mat4 r = Mat4::identity(); r.m[0][3] = x; r.m[1][3] = y; r.m[2][3] = z; return r; }
static Mat4 rotate_axis(int axis, double deg) {
    double t = deg * 3.14159265358979323846 / 180.0, s = std::sin(t), c = std::cos(t);
    Mat4 r = Mat4::identity();
    int a = (axis + 1) % 3, b = (axis + 2) % 3;
    r.m[a][a] = c; r.m[a][b] = -s; r.m[b][a] = s; r.m[b][b] = c;
    return r;
}
static void to_float16(const Mat4 &m, float out[16]) { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out[4 * i + j] = (float)m.m[i][j]; }

// ---------------------------------------------------------------------------------------------------
// Input soup: one record per triangle in the caller's order.
struct Soup {
    std::vector<float> p;          // 9 per triangle
    std::vector<float> uv, n;      // 6 / 9 per triangle (may stay empty)
    std::vector<uint8_t> has_n;
    std::vector<int32_t> material, light, med_in, med_out;
    std::vector<uint8_t> transition;
    bool anyUV = false, anyN = false, anyMedia = false;
    int count() const { return (int)material.size(); }
};

// Appends a mesh after the float operations the reference applies to it: an optional uniform
// pre-scale (plyInfo's x20, shape/plyRead.h:38) and a translation (TriangleMesh's ObjectToWorld,
// shape/Triangle.cpp:24-29: with a pure translation the 4x4 product reduces to x + tx).
static void add_mesh(Soup &s, const gnxsk::Mesh &m, float pre, const float t[3], int material, int medIn = -1, int medOut = -1) {
    for (int f = 0; f < m.nTris(); ++f) {
        for (int v = 0; v < 3; ++v) {
            int vi = m.idx[3 * f + v];
            for (int c = 0; c < 3; ++c) {
                float x = m.P[3 * vi + c];
                if (pre != 1.0f) x *= pre;
                s.p.push_back(x + t[c]);
            }
        }
        // (a mesh with per-triangle materials — OBJ usemtl — indexes them behind `material`; slot `material` itself is the default)
        s.material.push_back(m.tri_material.empty() ? material : material + 1 + m.tri_material[f]);
        s.light.push_back(-1);
        s.med_in.push_back(medIn); s.med_out.push_back(medOut); s.transition.push_back(medIn != medOut);
        if (medIn >= 0 || medOut >= 0) s.anyMedia = true;
        // per-triangle copies of the optional vertex attributes (a pure translation leaves normals unchanged)
        for (int v = 0; v < 3; ++v) {
            int vi = m.idx[3 * f + v];
            if (!m.UV.empty()) { s.uv.push_back(m.UV[2 * vi]); s.uv.push_back(m.UV[2 * vi + 1]); s.anyUV = true; }
            else { const float d[3][2] = {{0, 0}, {1, 0}, {1, 1}}; s.uv.push_back(d[v][0]); s.uv.push_back(d[v][1]); }
            for (int c = 0; c < 3; ++c) s.n.push_back(m.N.empty() ? 0.f : m.N[3 * vi + c]);
        }
        s.has_n.push_back(m.N.empty() ? 0 : 1);
        if (!m.N.empty()) s.anyN = true;
    }
}

// ---------------------------------------------------------------------------------------------------
// Binned-SAH BVH over triangle bounds, emitted directly in the 32-byte depth-first node layout the
// kernels traverse (first child = next node, second child by index).  One triangle per leaf unless
// centroids coincide.  Large subtrees are built as OpenMP tasks into private node arrays and spliced.
struct Box { float lo[3], hi[3]; };
static Box empty_box() { return {{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}}; }
static void grow(Box &b, const float *p) { for (int c = 0; c < 3; ++c) { b.lo[c] = std::min(b.lo[c], p[c]); b.hi[c] = std::max(b.hi[c], p[c]); } }
// (component-wise: an EMPTY box {+inf, -inf} must leave b unchanged — growing by its corners as points would not)
static void grow(Box &b, const Box &o) { for (int c = 0; c < 3; ++c) { b.lo[c] = std::min(b.lo[c], o.lo[c]); b.hi[c] = std::max(b.hi[c], o.hi[c]); } }
static float area(const Box &b) {
    float d[3] = {b.hi[0] - b.lo[0], b.hi[1] - b.lo[1], b.hi[2] - b.lo[2]};
    return 2 * (d[0] * d[1] + d[0] * d[2] + d[1] * d[2]);
}

struct BvhBuilder {
    const std::vector<Box> &pb;
    const std::vector<float> &cen;  // 3 per prim
    std::vector<int> order;
    static constexpr int kBins = 16;
    static constexpr int kTaskCutoff = 4096;
    // maxPrimsInNode of the SAH termination rule (accelerator/BVHAccel.cpp:297-318).  The reference's UI passes 1
    // (ui/RenderThread.cpp:155); GNXSK_MAX_LEAF overrides it for experiments (profiles/README.md).
    int maxLeaf = 1;

    BvhBuilder(const std::vector<Box> &b, const std::vector<float> &c) : pb(b), cen(c), order(b.size()) {
        for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
        if (const char *e = getenv("GNXSK_MAX_LEAF")) maxLeaf = std::max(1, std::min(16, atoi(e)));
    }

    void leaf(std::vector<gnx_bvh_node> &out, int self, const Box &b, int lo, int hi) {
        gnx_bvh_node &n = out[self];
        memcpy(n.bmin, b.lo, 12); memcpy(n.bmax, b.hi, 12);
        n.offset = lo; n.n_prims = (uint16_t)(hi - lo); n.axis = 0; n.pad = 0;
    }

    // Builds the subtree over order[lo, hi) appending to `out`; node indices are relative to out.
    void build(std::vector<gnx_bvh_node> &out, int lo, int hi) {
        const int self = (int)out.size();
        out.emplace_back();
        Box b = empty_box(), cb = empty_box();
        for (int i = lo; i < hi; ++i) { grow(b, pb[order[i]]); grow(cb, &cen[3 * order[i]]); }
        const int n = hi - lo;
        if (n == 1) { leaf(out, self, b, lo, hi); return; }
        int axis = 0;
        float ext[3] = {cb.hi[0] - cb.lo[0], cb.hi[1] - cb.lo[1], cb.hi[2] - cb.lo[2]};
        if (ext[1] > ext[axis]) axis = 1;
        if (ext[2] > ext[axis]) axis = 2;
        if (ext[axis] == 0 && n <= 65535) { leaf(out, self, b, lo, hi); return; }
        int mid = (lo + hi) / 2;
        auto byCentroid = [&](int a, int c) { return cen[3 * a + axis] < cen[3 * c + axis]; };
        if (n <= 2 || ext[axis] == 0) {
            std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi, byCentroid);
        } else {
            int cnt[kBins] = {0};
            Box bb[kBins];
            for (auto &x : bb) x = empty_box();
            const float scale = kBins / ext[axis];
            auto binOf = [&](int prim) { int k = (int)((cen[3 * prim + axis] - cb.lo[axis]) * scale); return k >= kBins ? kBins - 1 : k; };
            for (int i = lo; i < hi; ++i) { int k = binOf(order[i]); ++cnt[k]; grow(bb[k], pb[order[i]]); }
            float rightArea[kBins];
            int rightCnt[kBins];
            Box acc = empty_box();
            int c = 0;
            for (int k = kBins - 1; k > 0; --k) { grow(acc, bb[k]); c += cnt[k]; rightArea[k] = cnt[k] || c ? area(acc) : 0; rightCnt[k] = c; }
            acc = empty_box(); c = 0;
            float best = INFINITY;
            int bestK = -1;
            const float invA = 1.f / area(b);
            for (int k = 0; k < kBins - 1; ++k) {
                grow(acc, bb[k]); c += cnt[k];
                if (c == 0 || rightCnt[k + 1] == 0) continue;
                float cost = 0.125f + (c * area(acc) + rightCnt[k + 1] * rightArea[k + 1]) * invA;
                if (cost < best) { best = cost; bestK = k; }
            }
            // leaf instead of a split when it is no more expensive (cost of a leaf: one unit per primitive)
            if (n <= maxLeaf && !(best < (float)n)) { leaf(out, self, b, lo, hi); return; }
            if (bestK >= 0) {
                auto it = std::partition(order.begin() + lo, order.begin() + hi, [&](int prim) { return binOf(prim) <= bestK; });
                mid = (int)(it - order.begin());
            }
            if (bestK < 0 || mid == lo || mid == hi) {
                mid = (lo + hi) / 2;
                std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi, byCentroid);
            }
        }
        if (n >= 2 * kTaskCutoff) {
            std::vector<gnx_bvh_node> left, right;
#pragma omp task shared(left) firstprivate(lo, mid)
            build(left, lo, mid);
#pragma omp task shared(right) firstprivate(mid, hi)
            build(right, mid, hi);
#pragma omp taskwait
            const int lbase = (int)out.size();
            for (auto nd : left) { if (nd.n_prims == 0) nd.offset += lbase; out.push_back(nd); }
            const int rbase = (int)out.size();
            for (auto nd : right) { if (nd.n_prims == 0) nd.offset += rbase; out.push_back(nd); }
            gnx_bvh_node &nn = out[self];
            memcpy(nn.bmin, b.lo, 12); memcpy(nn.bmax, b.hi, 12);
            nn.offset = rbase; nn.n_prims = 0; nn.axis = (uint8_t)axis; nn.pad = 0;
        } else {
            build(out, lo, mid);
            const int second = (int)out.size();
            build(out, mid, hi);
            gnx_bvh_node &nn = out[self];
            memcpy(nn.bmin, b.lo, 12); memcpy(nn.bmax, b.hi, 12);
            nn.offset = second; nn.n_prims = 0; nn.axis = (uint8_t)axis; nn.pad = 0;
        }
    }
};

// ---------------------------------------------------------------------------------------------------
// Radiance .hdr (RGBE, new-style RLE) -> float RGB, value = mantissa * 2^(e - 136) like stb_image's
// stbi_loadf (3rd/stb_image.h:7130-7155), which is what InfiniteAreaLight feeds on.
static bool load_hdr(const std::string &path, int *w, int *h, std::vector<float> &rgb, std::string *err) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { *err = "cannot open " + path; return false; }
    std::string line;
    bool fmt = false;
    while (std::getline(f, line)) {
        if (line.empty() || line == "\r") break;
        if (line.find("FORMAT=32-bit_rle_rgbe") != std::string::npos) fmt = true;
    }
    if (!fmt) { *err = "not an RGBE .hdr: " + path; return false; }
    std::getline(f, line);
    if (sscanf(line.c_str(), "-Y %d +X %d", h, w) != 2) { *err = "unsupported .hdr orientation"; return false; }
    const int W = *w, H = *h;
    rgb.resize((size_t)W * H * 3);
    std::vector<unsigned char> scan((size_t)W * 4);
    auto convert = [&](const unsigned char *px, float *out) {
        if (px[3] != 0) {
            float f1 = (float)std::ldexp(1.0f, (int)px[3] - (128 + 8));
            out[0] = px[0] * f1; out[1] = px[1] * f1; out[2] = px[2] * f1;
        } else out[0] = out[1] = out[2] = 0;
    };
    for (int y = 0; y < H; ++y) {
        unsigned char hd[4];
        f.read((char *)hd, 4);
        if (!f) { *err = "truncated .hdr"; return false; }
        if (W < 8 || W >= 32768 || hd[0] != 2 || hd[1] != 2 || (hd[2] & 0x80)) {
            // flat scanline
            memcpy(scan.data(), hd, 4);
            f.read((char *)scan.data() + 4, (size_t)(W - 1) * 4);
        } else {
            if (((int)hd[2] << 8 | hd[3]) != W) { *err = "bad .hdr scanline width"; return false; }
            for (int c = 0; c < 4; ++c) {
                int x = 0;
                while (x < W) {
                    int cnt = f.get();
                    if (cnt > 128) { int val = f.get(); cnt -= 128; while (cnt-- && x < W) scan[4 * (x++) + c] = (unsigned char)val; }
                    else while (cnt-- && x < W) scan[4 * (x++) + c] = (unsigned char)f.get();
                }
            }
        }
        if (!f) { *err = "truncated .hdr"; return false; }
        for (int x = 0; x < W; ++x) convert(&scan[4 * x], &rgb[((size_t)y * W + x) * 3]);
    }
    return true;
}

static float lanczos(float x, float tau = 2) {  // core/Texture.cpp:150-160
    x = std::fabs(x);
    if (x < 1e-5f) return 1;
    if (x > 1.f) return 0;
    x *= kPiF;
    float s = std::sin(x * tau) / (x * tau);
    float l = std::sin(x) / x;
    return s * l;
}
static int round_up_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

struct EnvTables {
    int w = 0, h = 0, dw = 0, dh = 0;
    std::vector<float> texels, cond_func, cond_cdf, cond_int, marg_func, marg_cdf;
    float marg_int = 0;
};

// Distribution1D (core/Sampling.h:22-35)
static float dist1d(const float *f, int n, float *cdf) {
    cdf[0] = 0;
    for (int i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + f[i - 1] / n;
    float funcInt = cdf[n];
    if (funcInt == 0) for (int i = 1; i < n + 1; ++i) cdf[i] = float(i) / float(n);
    else for (int i = 1; i < n + 1; ++i) cdf[i] /= funcInt;
    return funcInt;
}

// InfiniteAreaLight's constructor (lights/InfiniteAreaLight.cpp:12-82): texel = (L * rgb)^1.5, MIPMap
// resample to a power of two with a 4-tap Lanczos filter and Repeat wrap (core/MIPMap.h:86-150), then
// the sin-weighted luminance image at twice the resolution and its Distribution2D.
static void build_env(const std::vector<float> &rgb, int w0, int h0, float L, EnvTables &e) {
    std::vector<float> tex((size_t)w0 * h0 * 3);
    for (size_t i = 0; i < tex.size(); ++i) { float r = L * rgb[i]; tex[i] = r * std::sqrt(r); }
    int W = w0, H = h0;
    if ((w0 & (w0 - 1)) || (h0 & (h0 - 1))) {
        W = round_up_pow2(w0); H = round_up_pow2(h0);
        auto weights = [](int oldRes, int newRes, std::vector<int> &first, std::vector<float> &wt) {
            first.resize(newRes); wt.resize((size_t)newRes * 4);
            const float filterwidth = 2.f;
            for (int i = 0; i < newRes; ++i) {
                float center = (i + .5f) * oldRes / newRes;
                first[i] = (int)std::floor((center - filterwidth) + 0.5f);
                for (int j = 0; j < 4; ++j) { float pos = first[i] + j + .5f; wt[4 * i + j] = lanczos((pos - center) / filterwidth); }
                float inv = 1 / (wt[4 * i] + wt[4 * i + 1] + wt[4 * i + 2] + wt[4 * i + 3]);
                for (int j = 0; j < 4; ++j) wt[4 * i + j] *= inv;
            }
        };
        auto modp = [](int a, int b) { int r = a % b; return r < 0 ? r + b : r; };
        std::vector<int> sf, tf;
        std::vector<float> sw, tw;
        weights(w0, W, sf, sw);
        weights(h0, H, tf, tw);
        std::vector<float> tmp((size_t)W * H * 3, 0.f);
        for (int t = 0; t < h0; ++t)
            for (int s = 0; s < W; ++s)
                for (int c = 0; c < 3; ++c) {
                    float v = 0.f;
                    for (int j = 0; j < 4; ++j) { int os = modp(sf[s] + j, w0); v += sw[4 * s + j] * tex[((size_t)t * w0 + os) * 3 + c]; }
                    tmp[((size_t)t * W + s) * 3 + c] = v;
                }
        std::vector<float> col((size_t)H * 3);
        for (int s = 0; s < W; ++s) {
            for (int t = 0; t < H; ++t)
                for (int c = 0; c < 3; ++c) {
                    float v = 0.f;
                    for (int j = 0; j < 4; ++j) { int ot = modp(tf[t] + j, h0); v += tw[4 * t + j] * tmp[((size_t)ot * W + s) * 3 + c]; }
                    col[3 * t + c] = v;
                }
            for (int t = 0; t < H; ++t)
                for (int c = 0; c < 3; ++c) tmp[((size_t)t * W + s) * 3 + c] = std::max(0.f, col[3 * t + c]);
        }
        tex.swap(tmp);
    }
    e.w = W; e.h = H; e.texels = tex;
    e.dw = 2 * W; e.dh = 2 * H;
    auto texel = [&](int s, int t, int c) {
        s %= W; if (s < 0) s += W;
        t %= H; if (t < 0) t += H;
        return e.texels[((size_t)t * W + s) * 3 + c];
    };
    std::vector<float> img((size_t)e.dw * e.dh);
#pragma omp parallel for
    for (int v = 0; v < e.dh; ++v) {
        float vp = (v + .5f) / (float)e.dh;
        float sinTheta = std::sin(kPiF * (v + .5f) / e.dh);
        for (int u = 0; u < e.dw; ++u) {
            float up = (u + .5f) / (float)e.dw;
            float s = up * W - 0.5f, t = vp * H - 0.5f;
            int s0 = (int)std::floor(s), t0 = (int)std::floor(t);
            float ds = s - s0, dt = t - t0;
            float rgbv[3];
            for (int c = 0; c < 3; ++c)
                rgbv[c] = (1 - ds) * (1 - dt) * texel(s0, t0, c) + (1 - ds) * dt * texel(s0, t0 + 1, c) +
                          ds * (1 - dt) * texel(s0 + 1, t0, c) + ds * dt * texel(s0 + 1, t0 + 1, c);
            float y = 0.212671f * rgbv[0] + 0.715160f * rgbv[1] + 0.072169f * rgbv[2];
            img[(size_t)v * e.dw + u] = y * sinTheta;
        }
    }
    e.cond_func = img;
    e.cond_cdf.resize((size_t)(e.dw + 1) * e.dh);
    e.cond_int.resize(e.dh);
    for (int v = 0; v < e.dh; ++v) e.cond_int[v] = dist1d(&img[(size_t)v * e.dw], e.dw, &e.cond_cdf[(size_t)v * (e.dw + 1)]);
    e.marg_func = e.cond_int;
    e.marg_cdf.resize(e.dh + 1);
    e.marg_int = dist1d(e.marg_func.data(), e.dh, e.marg_cdf.data());
}

static gnx_material make_material(int type, unsigned flags) {
    gnx_material m{};
    m.type = type; m.flags = flags;
    for (int &t : m.rgb_tex) t = -1;
    for (int &t : m.f_tex) t = -1;
    return m;
}
static void set_rgb(gnx_material &m, int slot, float r, float g, float b) { m.rgb[slot][0] = r; m.rgb[slot][1] = g; m.rgb[slot][2] = b; }

}  // namespace

struct gnxsk_scene {
    std::string error;
    gnx_scene_desc desc{};
    std::vector<gnx_bvh_node> nodes;
    std::vector<float> prim_p, prim_uv, prim_n;
    std::vector<uint8_t> prim_has_n, prim_flags;
    std::vector<int32_t> prim_material, prim_light, prim_id, prim_med_in, prim_med_out;
    std::vector<uint8_t> prim_transition;
    std::vector<gnx_medium> media;
    std::vector<float> density;
    bool pcgSampler = false;
    std::vector<gnx_material> materials;
    std::vector<gnx_texture> textures;
    std::vector<float> texels;
    std::vector<gnx_light> lights;
    std::vector<int32_t> light_n_samples;  // Light::nSamples: 5 for the UI's area lights (ui/ModelList.cpp:140-146), else 1
    EnvTables env;
    double build_seconds = 0;

    // camera: LookAt(eye -> look, up +y); the UI's default is (0,0,5) -> origin (ui/RenderThread.cpp:60-68)
    double cam_eye[3] = {0, 0, 5}, cam_look[3] = {0, 0, 0};
    float world_radius = 0;  // Bounds3::BoundingSphere radius of the scene, for DistantLight::Preprocess

    static Mat4 look_at_c2w(const double eye_[3], const double look_[3]) {
        Vec3d eye{eye_[0], eye_[1], eye_[2]}, look{look_[0], look_[1], look_[2]}, up{0, 1, 0};
        Vec3d dir = nrm(sub(look, eye)), right = nrm(crs(nrm(up), dir)), newUp = crs(dir, right);
        Mat4 c2w = Mat4::identity();
        c2w.m[0][0] = right.x; c2w.m[1][0] = right.y; c2w.m[2][0] = right.z;
        c2w.m[0][1] = newUp.x; c2w.m[1][1] = newUp.y; c2w.m[2][1] = newUp.z;
        c2w.m[0][2] = dir.x; c2w.m[1][2] = dir.y; c2w.m[2][2] = dir.z;
        c2w.m[0][3] = eye.x; c2w.m[1][3] = eye.y; c2w.m[2][3] = eye.z;
        return c2w;
    }

    void finalize(Soup &soup, int width, int height, int spp, bool withEnv, const Mat4 &envL2W) {
        const int n = soup.count();
        // ---- BVH
        std::vector<Box> pb(n);
        std::vector<float> cen((size_t)n * 3);
        Box wb = empty_box();
        for (int i = 0; i < n; ++i) {
            Box b = empty_box();
            for (int v = 0; v < 3; ++v) grow(b, &soup.p[(size_t)i * 9 + 3 * v]);
            pb[i] = b;
            for (int c = 0; c < 3; ++c) cen[3 * (size_t)i + c] = .5f * b.lo[c] + .5f * b.hi[c];
            grow(wb, b);
        }
        auto t0 = std::chrono::steady_clock::now();
        BvhBuilder bb(pb, cen);
        if (n > 0) {
#pragma omp parallel
#pragma omp single
            bb.build(nodes, 0, n);
        }
        build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        // ---- reorder primitives into BVH order
        prim_p.resize((size_t)n * 9);
        prim_material.resize(n); prim_light.assign(n, -1); prim_id.resize(n); prim_flags.assign(n, 0);
        std::vector<int> newIndex(n);
        for (int k = 0; k < n; ++k) {
            int src = bb.order[k];
            newIndex[src] = k;
            memcpy(&prim_p[(size_t)k * 9], &soup.p[(size_t)src * 9], 36);
            prim_material[k] = soup.material[src];
            prim_id[k] = src;
        }
        if (soup.anyMedia) {
            prim_med_in.resize(n); prim_med_out.resize(n); prim_transition.resize(n);
            for (int k = 0; k < n; ++k) { int src = bb.order[k]; prim_med_in[k] = soup.med_in[src]; prim_med_out[k] = soup.med_out[src]; prim_transition[k] = soup.transition[src]; }
        }
        if (soup.anyUV) {
            prim_uv.resize((size_t)n * 6);
            for (int k = 0; k < n; ++k) memcpy(&prim_uv[(size_t)k * 6], &soup.uv[(size_t)bb.order[k] * 6], 24);
        }
        if (soup.anyN) {
            prim_n.resize((size_t)n * 9);
            prim_has_n.resize(n);
            for (int k = 0; k < n; ++k) { memcpy(&prim_n[(size_t)k * 9], &soup.n[(size_t)bb.order[k] * 9], 36); prim_has_n[k] = soup.has_n[bb.order[k]]; }
        }
        for (size_t i = 0; i < lights.size(); ++i)
            if (lights[i].type == GNX_LIGHT_AREA_TRI) {
                lights[i].prim = newIndex[lights[i].prim];
                prim_light[lights[i].prim] = (int32_t)i;
            }
        gnx_geometry &g = desc.geom;
        g.n_nodes = (int32_t)nodes.size(); g.nodes = nodes.data();
        g.n_prims = n; g.prim_p = prim_p.data();
        g.prim_material = prim_material.data(); g.prim_light = prim_light.data();
        g.prim_flags = prim_flags.data(); g.prim_id = prim_id.data();
        g.prim_medium_in = soup.anyMedia ? prim_med_in.data() : nullptr;
        g.prim_medium_out = soup.anyMedia ? prim_med_out.data() : nullptr;
        g.prim_is_transition = soup.anyMedia ? prim_transition.data() : nullptr;
        desc.n_media = (int32_t)media.size(); desc.media = media.data();
        g.prim_uv = soup.anyUV ? prim_uv.data() : nullptr;
        g.prim_n = soup.anyN ? prim_n.data() : nullptr;
        g.prim_has_n = soup.anyN ? prim_has_n.data() : nullptr;
        desc.n_textures = (int32_t)textures.size(); desc.textures = textures.data();
        memcpy(g.world_bound, wb.lo, 12); memcpy(g.world_bound + 3, wb.hi, 12);
        desc.abi_version = GNX_ABI_VERSION;
        desc.n_materials = (int32_t)materials.size(); desc.materials = materials.data();
        // ---- environment light
        if (withEnv) {
            gnx_light l{};
            l.type = GNX_LIGHT_INFINITE; l.prim = -1; l.medium = -1;
            lights.push_back(l);
            gnx_envmap &e = desc.env;
            e.present = 1; e.light_index = (int32_t)lights.size() - 1;
            e.width = env.w; e.height = env.h; e.texels = env.texels.data();
            e.dist_w = env.dw; e.dist_h = env.dh;
            e.cond_func = env.cond_func.data(); e.cond_cdf = env.cond_cdf.data(); e.cond_int = env.cond_int.data();
            e.marg_func = env.marg_func.data(); e.marg_cdf = env.marg_cdf.data(); e.marg_int = env.marg_int;
            to_float16(envL2W, e.light_to_world);
            to_float16(inverse(envL2W), e.world_to_light);
            // Bounds3::BoundingSphere (core/Geometry.h:770-773)
            float c[3], r2 = 0;
            for (int a = 0; a < 3; ++a) { c[a] = (wb.lo[a] + wb.hi[a]) / 2; e.world_center[a] = c[a]; float d = c[a] - wb.hi[a]; r2 += d * d; }
            e.world_radius = std::sqrt(r2);
        }
        {
            // Bounds3::BoundingSphere (core/Geometry.h:770-773): what DistantLight::Preprocess caches
            float r2 = 0;
            for (int a = 0; a < 3; ++a) { float c = (wb.lo[a] + wb.hi[a]) / 2, d = c - wb.hi[a]; r2 += d * d; }
            world_radius = std::sqrt(r2);
            for (gnx_light &l : lights) if (l.type == GNX_LIGHT_DISTANT) l.area = world_radius;
        }
        desc.n_lights = (int32_t)lights.size(); desc.lights = lights.data();
        light_n_samples.clear();
        for (const gnx_light &l : lights) light_n_samples.push_back(l.type == GNX_LIGHT_AREA_TRI ? 5 : 1);
        desc.light_n_samples = light_n_samples.data();
        // ---- camera: LookAt((0,0,5) -> origin, up +y), fov 90, near 1e-2, far 1000, screen window by
        // aspect (ui/RenderThread.cpp:60-68, camera/Perspective.cpp:114-135, core/Camera.h:54-75)
        Mat4 c2w = look_at_c2w(cam_eye, cam_look);
        double frame = (double)width / height, sx0, sx1, sy0, sy1;
        if (frame > 1) { sx0 = -frame; sx1 = frame; sy0 = -1; sy1 = 1; }
        else { sx0 = -1; sx1 = 1; sy0 = -1 / frame; sy1 = 1 / frame; }
        const double nr = 1e-2f, fr = 1000.0;
        Mat4 persp{};
        persp.m[0][0] = 1; persp.m[1][1] = 1; persp.m[2][2] = fr / (fr - nr); persp.m[2][3] = -fr * nr / (fr - nr); persp.m[3][2] = 1;
        double invTan = 1 / std::tan((90.0 * 3.14159265358979323846 / 180.0) / 2);
        Mat4 c2s = mul(scale(invTan, invTan, 1), persp);
        Mat4 s2r = mul(mul(scale(width, height, 1), scale(1 / (sx1 - sx0), 1 / (sy0 - sy1), 1)), translate(-sx0, -sy1, 0));
        Mat4 r2c = mul(inverse(c2s), inverse(s2r));
        to_float16(r2c, desc.camera.raster_to_camera);
        to_float16(c2w, desc.camera.camera_to_world);
        desc.camera.lens_radius = 0; desc.camera.focal_distance = 3.0f;
        desc.camera.shutter_open = 0; desc.camera.shutter_close = 0;
        desc.camera.medium = -1;
        // ---- Halton parameters (samplers/HaltonSampler.cpp:33-61): scales >= min(res, 128)
        gnx_sampler &s = desc.sampler;
        s.type = pcgSampler ? GNX_SAMPLER_PCG32 : GNX_SAMPLER_HALTON; s.samples_per_pixel = spp;
        const int res[2] = {width, height};
        for (int i = 0; i < 2; ++i) {
            int base = i == 0 ? 2 : 3, sc = 1, ex = 0;
            while (sc < std::min(res[i], 128)) { sc *= base; ++ex; }
            s.base_scales[i] = sc; s.base_exponents[i] = ex;
        }
        s.sample_stride = s.base_scales[0] * s.base_scales[1];
        auto mulinv = [](long long a, long long n) {  // modular inverse by brute force (n <= 243)
            for (long long x = 0; x < n; ++x) if ((a * x) % n == 1 % n) return (int)x;
            return 0;
        };
        s.mult_inverse[0] = mulinv(s.base_scales[1], s.base_scales[0]);
        s.mult_inverse[1] = mulinv(s.base_scales[0], s.base_scales[1]);
        s.sample_at_pixel_center = 0;
        s.perms = nullptr;  // the library derives the permutation table itself
        s.n_perm_entries = 0;
    }
};

extern "C" {

gnxsk_scene *gnxsk_create(const char *name, int width, int height, int spp, int p0, int p1, int p2, const char *resources) {
    auto *sc = new gnxsk_scene;
    if (!name || width <= 0 || height <= 0 || spp <= 0) { sc->error = "bad arguments"; return sc; }
    const std::string nm = name;
    Soup soup;
    const float zero[3] = {0, 0, 0};
    if (nm == "cornell") {
        // same recipe as oracle/ref_harness.cpp::BuildCornell (ui/ModelList.cpp:71-146, ui/RenderThread.cpp:79-99)
        const float sigma = p0 == 1 ? 60.f : 0.f;
        auto matte = [&](float r, float g, float b) { gnx_material m = make_material(GNX_MAT_MATTE, GNX_MATF_BUMP_IDENTITY); set_rgb(m, 0, r, g, b); m.f[0] = sigma; return m; };
        sc->materials.push_back(matte(0.91f, 0.91f, 0.91f));  // 0 white
        sc->materials.push_back(matte(0.9f, 0.1f, 0.17f));    // 1 red
        sc->materials.push_back(matte(0.14f, 0.21f, 0.87f));  // 2 blue
        gnx_material mirror = make_material(GNX_MAT_MIRROR, GNX_MATF_BUMP_IDENTITY); set_rgb(mirror, 0, 0.9f, 0.9f, 0.9f);
        gnx_material glass = make_material(GNX_MAT_GLASS, GNX_MATF_BUMP_IDENTITY);
        set_rgb(glass, 0, 0.98f, 0.98f, 0.98f); set_rgb(glass, 1, 0.98f, 0.98f, 0.98f); glass.f[0] = 0; glass.f[1] = 0; glass.f[2] = 1.5f;
        sc->materials.push_back(mirror);  // 3
        sc->materials.push_back(glass);   // 4
        if (p1 >= 0) {
            add_mesh(soup, gnxsk::icosphere(p1, 0.8f, -1.0f, -1.7f, -0.5f), 1.f, zero, 3);
            add_mesh(soup, gnxsk::icosphere(p1, 0.8f, 1.0f, -1.7f, 0.8f), 1.f, zero, 4);
        }
        gnxsk::Mesh walls = gnxsk::cornell_walls(5.0f);
        const float boxT[3] = {-2.5f, -2.5f, -2.5f};
        for (int i = 0; i < 10; ++i) {
            gnxsk::Mesh one;
            for (int v = 0; v < 3; ++v) { one.P.insert(one.P.end(), {walls.P[9 * i + 3 * v], walls.P[9 * i + 3 * v + 1], walls.P[9 * i + 3 * v + 2]}); one.idx.push_back(v); }
            add_mesh(soup, one, 1.f, boxT, (i == 6 || i == 7) ? 1 : (i == 8 || i == 9) ? 2 : 0);
        }
        const float lightT[3] = {0.0f, 2.45f, 0.0f};
        int firstLight = soup.count();
        add_mesh(soup, gnxsk::area_light_quad(1.4f), 1.f, lightT, 0);
        for (int k = firstLight; k < soup.count(); ++k) {
            const float *p = &soup.p[(size_t)k * 9];
            Vec3d a{p[3] - p[0], p[4] - p[1], p[5] - p[2]}, b{p[6] - p[0], p[7] - p[1], p[8] - p[2]};
            Vec3d c = crs(a, b);
            float cx = (float)c.x, cy = (float)c.y, cz = (float)c.z;  // Cross() is evaluated in double, stored as float
            gnx_light l{};
            l.type = GNX_LIGHT_AREA_TRI; l.prim = k; l.two_sided = 0; l.medium = -1;
            l.L[0] = l.L[1] = l.L[2] = 5.0f;
            l.area = 0.5f * std::sqrt(cx * cx + cy * cy + cz * cz);  // Triangle::Area, shape/Triangle.cpp:455-462
            sc->lights.push_back(l);
        }
        sc->finalize(soup, width, height, spp, false, Mat4::identity());
    } else if (nm == "lights") {
        // SURVEY §8f rank 1, same recipe as oracle/ref_harness.cpp::BuildLightsRoom: the Cornell room with a Mirror, a
        // Glass and a Plastic sphere, lit by area + Point + Spot + Distant + SkyBox lights (p0 = light mask, bits as in
        // the harness; the kit has no JPEG decoder, so the SkyBox shows its procedural colours).  For the Whitted and
        // DirectLighting integrators.
        const int mask = p0 > 0 ? p0 : 31, subdiv = p1 > 0 ? p1 : 2;
        auto matte = [&](float r, float g, float b, float sigma) { gnx_material m = make_material(GNX_MAT_MATTE, GNX_MATF_BUMP_IDENTITY); set_rgb(m, 0, r, g, b); m.f[0] = sigma; return m; };
        sc->materials.push_back(matte(0.91f, 0.91f, 0.91f, 0.f));   // 0 white
        sc->materials.push_back(matte(0.9f, 0.1f, 0.17f, 0.f));     // 1 red
        sc->materials.push_back(matte(0.14f, 0.21f, 0.87f, 30.f));  // 2 blue (Oren-Nayar)
        gnx_material mirror = make_material(GNX_MAT_MIRROR, GNX_MATF_BUMP_IDENTITY); set_rgb(mirror, 0, 0.9f, 0.9f, 0.9f);
        gnx_material glass = make_material(GNX_MAT_GLASS, GNX_MATF_BUMP_IDENTITY);
        set_rgb(glass, 0, 0.98f, 0.98f, 0.98f); set_rgb(glass, 1, 0.98f, 0.98f, 0.98f); glass.f[0] = 0; glass.f[1] = 0; glass.f[2] = 1.5f;
        gnx_material plastic = make_material(GNX_MAT_PLASTIC, GNX_MATF_BUMP_IDENTITY | GNX_MATF_REMAP_ROUGHNESS);
        set_rgb(plastic, 0, 0.35f, 0.12f, 0.48f); set_rgb(plastic, 1, 0.65f, 0.88f, 0.52f); plastic.f[0] = 0.1f;
        sc->materials.push_back(mirror);   // 3
        sc->materials.push_back(glass);    // 4
        sc->materials.push_back(plastic);  // 5
        add_mesh(soup, gnxsk::icosphere(subdiv, 0.8f, -1.0f, -1.7f, -0.5f), 1.f, zero, 3);
        add_mesh(soup, gnxsk::icosphere(subdiv, 0.8f, 1.0f, -1.7f, 0.8f), 1.f, zero, 4);
        add_mesh(soup, gnxsk::icosphere(subdiv > 1 ? subdiv - 1 : subdiv, 0.5f, 0.0f, -2.0f, 1.6f), 1.f, zero, 5);
        gnxsk::Mesh walls = gnxsk::cornell_walls(5.0f);
        const float boxT[3] = {-2.5f, -2.5f, -2.5f};
        for (int i = 0; i < 10; ++i) {
            gnxsk::Mesh one;
            for (int v = 0; v < 3; ++v) { one.P.insert(one.P.end(), {walls.P[9 * i + 3 * v], walls.P[9 * i + 3 * v + 1], walls.P[9 * i + 3 * v + 2]}); one.idx.push_back(v); }
            add_mesh(soup, one, 1.f, boxT, (i == 6 || i == 7) ? 1 : (i == 8 || i == 9) ? 2 : 0);
        }
        if (mask & 1) {
            const float lightT[3] = {0.0f, 2.45f, 0.0f};
            int firstLight = soup.count();
            add_mesh(soup, gnxsk::area_light_quad(1.4f), 1.f, lightT, 0);
            for (int k = firstLight; k < soup.count(); ++k) {
                const float *p = &soup.p[(size_t)k * 9];
                Vec3d a{p[3] - p[0], p[4] - p[1], p[5] - p[2]}, b{p[6] - p[0], p[7] - p[1], p[8] - p[2]};
                Vec3d c = crs(a, b);
                float cx = (float)c.x, cy = (float)c.y, cz = (float)c.z;
                gnx_light l{};
                l.type = GNX_LIGHT_AREA_TRI; l.prim = k; l.two_sided = 0; l.medium = -1;
                l.L[0] = l.L[1] = l.L[2] = 3.0f;
                l.area = 0.5f * std::sqrt(cx * cx + cy * cy + cz * cz);
                sc->lights.push_back(l);
            }
        }
        if (mask & 2) {
            gnx_light l{};
            l.type = GNX_LIGHT_POINT; l.prim = -1; l.medium = -1;
            l.L[0] = 6.f; l.L[1] = 5.f; l.L[2] = 4.f;
            l.p[0] = 1.5f; l.p[1] = 1.8f; l.p[2] = 1.0f;
            sc->lights.push_back(l);
        }
        if (mask & 4) {
            gnx_light l{};
            l.type = GNX_LIGHT_SPOT; l.prim = -1; l.medium = -1;
            l.L[0] = 9.f; l.L[1] = 14.f; l.L[2] = 18.f;
            const double eye[3] = {-1.8f, 2.0f, 1.8f}, look[3] = {0.2f, -2.0f, 0.0f};
            l.p[0] = (float)eye[0]; l.p[1] = (float)eye[1]; l.p[2] = (float)eye[2];
            l.cos_total = std::cos((3.14159265358979323846f / 180.f) * 32.f);    // Radians(totalWidth), float like the reference
            l.cos_falloff = std::cos((3.14159265358979323846f / 180.f) * 22.f);
            to_float16(inverse(gnxsk_scene::look_at_c2w(eye, look)), l.world_to_light);
            sc->lights.push_back(l);
        }
        if (mask & 8) {
            gnx_light l{};
            l.type = GNX_LIGHT_DISTANT; l.prim = -1; l.medium = -1;
            l.L[0] = 0.5f; l.L[1] = 0.45f; l.L[2] = 0.35f;
            // Normalize(LightToWorld(w)), LightToWorld = RotateY(15)
            Mat4 r = rotate_axis(1, 15);
            double w[3] = {0.3f, 0.4f, 1.0f}, o[3];
            for (int i = 0; i < 3; ++i) o[i] = r.m[i][0] * w[0] + r.m[i][1] * w[1] + r.m[i][2] * w[2];
            float fo[3] = {(float)o[0], (float)o[1], (float)o[2]};
            float len = std::sqrt(fo[0] * fo[0] + fo[1] * fo[1] + fo[2] * fo[2]), inv = 1.f / len;
            for (int i = 0; i < 3; ++i) l.p[i] = fo[i] * inv;
            sc->lights.push_back(l);
        }
        if (mask & (16 | 32)) {
            gnx_light l{};
            l.type = GNX_LIGHT_SKYBOX; l.prim = -1; l.medium = -1;
            to_float16(rotate_axis(0, 10), l.world_to_light);  // SKYBOX: LightToWorld
            sc->lights.push_back(l);
            gnx_skybox &k = sc->desc.skybox;
            k.present = 1; k.light_index = (int32_t)sc->lights.size() - 1;
            k.width = k.height = k.channels = 0; k.data = nullptr;
            k.center[0] = k.center[1] = k.center[2] = 0.f;
            k.radius = 50.f;
        }
        sc->cam_eye[2] = 6.5; sc->cam_look[1] = -0.4f;
        sc->finalize(soup, width, height, spp, false, Mat4::identity());
    } else if (nm == "ui" || nm.rfind("ui3d:", 0) == 0) {
        // The reference UI's live scene (ui/RenderThread.cpp:60-164), same recipe as oracle/ref_harness.cpp::BuildUI: the
        // mesh (Matte sigma 60) inside the five-wall Cornell box (Oren-Nayar sigma 60), the two-triangle area light with
        // the mesh's material, a SkyBoxLight of radius 10 without an image.  Every camera ray hits a surface.
        // p1 x p2 knot quads (0 = 2048 x 213); "ui3d:<path>" reads the mesh from a .3d file instead.
        auto matte = [&](float r, float g, float b) { gnx_material m = make_material(GNX_MAT_MATTE, GNX_MATF_BUMP_IDENTITY); set_rgb(m, 0, r, g, b); m.f[0] = 60.f; return m; };
        sc->materials.push_back(matte(0.2f, 0.8f, 0.2f));     // 0 mesh + light quad
        sc->materials.push_back(matte(0.91f, 0.91f, 0.91f));  // 1 white
        sc->materials.push_back(matte(0.9f, 0.1f, 0.17f));    // 2 red
        sc->materials.push_back(matte(0.14f, 0.21f, 0.87f));  // 3 blue
        const float T[3] = {0.f, -2.9f, 0.f};
        if (nm == "ui") add_mesh(soup, gnxsk::torus_knot(p1 > 0 ? p1 : 2048, p2 > 0 ? p2 : 213), 20.f, T, 0);
        else {
            gnxsk::Mesh file;
            if (!gnxsk::load_3d(nm.substr(5), &file, &sc->error)) return sc;
            if (file.nTris() == 0) { sc->error = "mesh file without triangles"; return sc; }
            add_mesh(soup, file, 20.f, T, 0);
        }
        gnxsk::Mesh walls = gnxsk::cornell_walls(5.0f);
        const float boxT[3] = {-2.5f, -2.5f, -2.5f};
        for (int i = 0; i < 10; ++i) {
            gnxsk::Mesh one;
            for (int v = 0; v < 3; ++v) { one.P.insert(one.P.end(), {walls.P[9 * i + 3 * v], walls.P[9 * i + 3 * v + 1], walls.P[9 * i + 3 * v + 2]}); one.idx.push_back(v); }
            add_mesh(soup, one, 1.f, boxT, (i == 6 || i == 7) ? 2 : (i == 8 || i == 9) ? 3 : 1);
        }
        const float lightT[3] = {0.0f, 2.45f, 0.0f};
        int firstLight = soup.count();
        add_mesh(soup, gnxsk::area_light_quad(1.4f), 1.f, lightT, 0);
        for (int k = firstLight; k < soup.count(); ++k) {
            const float *p = &soup.p[(size_t)k * 9];
            Vec3d a{p[3] - p[0], p[4] - p[1], p[5] - p[2]}, b{p[6] - p[0], p[7] - p[1], p[8] - p[2]};
            Vec3d c = crs(a, b);
            float cx = (float)c.x, cy = (float)c.y, cz = (float)c.z;
            gnx_light l{};
            l.type = GNX_LIGHT_AREA_TRI; l.prim = k; l.two_sided = 0; l.medium = -1;
            l.L[0] = l.L[1] = l.L[2] = 5.0f;
            l.area = 0.5f * std::sqrt(cx * cx + cy * cy + cz * cz);
            sc->lights.push_back(l);
        }
        {
            gnx_light l{};
            l.type = GNX_LIGHT_SKYBOX; l.prim = -1; l.medium = -1;
            to_float16(Mat4::identity(), l.world_to_light);  // SKYBOX: LightToWorld
            sc->lights.push_back(l);
            gnx_skybox &k = sc->desc.skybox;
            k.present = 1; k.light_index = (int32_t)sc->lights.size() - 1;
            k.width = k.height = k.channels = 0; k.data = nullptr;
            k.center[0] = k.center[1] = k.center[2] = 0.f;
            k.radius = 10.f;
        }
        sc->finalize(soup, width, height, spp, false, Mat4::identity());
    } else if (nm == "dragon" || nm.rfind("dragon3d:", 0) == 0 || nm.rfind("obj:", 0) == 0) {
        // "dragon3d:<path>": the mesh comes from a .3d file, placed like ui/ModelList.cpp:49-69 (plyInfo's x20, then
        // Translate(0, -2.9, 0)); "obj:<path>": a Wavefront OBJ fitted into a sphere of radius 2.5 around (0, -0.4, 0)
        // oracle/ref_harness.cpp::BuildDragon (ui/MaterialList.cpp:48-69, ui/ModelList.cpp:49-69,172-178)
        if (p0 == 1) {
            gnx_material m = make_material(GNX_MAT_METAL, GNX_MATF_BUMP_IDENTITY);
            set_rgb(m, 0, 0.2f, 0.2f, 0.8f); set_rgb(m, 1, 0.11f, 0.11f, 0.11f); m.f[0] = 0.15f; m.f[1] = 0.15f;
            sc->materials.push_back(m);
        } else {
            gnx_material m = make_material(GNX_MAT_PLASTIC, GNX_MATF_BUMP_IDENTITY | GNX_MATF_REMAP_ROUGHNESS);
            set_rgb(m, 0, 0.35f, 0.12f, 0.48f); set_rgb(m, 1, 1.f - 0.35f, 1.f - 0.12f, 1.f - 0.48f); m.f[0] = 0.1f;
            sc->materials.push_back(m);
        }
        const float T[3] = {0.f, -2.9f, 0.f};
        if (nm == "dragon") add_mesh(soup, gnxsk::torus_knot(p1 > 0 ? p1 : 2048, p2 > 0 ? p2 : 213), 20.f, T, 0);
        else {
            gnxsk::Mesh file;
            const bool obj = nm[0] == 'o';
            std::vector<gnxsk::ObjMaterial> objMats;
            if (!(obj ? gnxsk::load_obj(nm.substr(4), &file, &sc->error, &objMats) : gnxsk::load_3d(nm.substr(9), &file, &sc->error))) return sc;
            if (file.nTris() == 0) { sc->error = "mesh file without triangles"; return sc; }
            // OBJ materials (MTL): one gnx_material per entry behind the default one, through the same recipe the oracle
            // harness feeds to the reference's material classes
            for (const gnxsk::ObjMaterial &om : objMats) {
                const gnxsk::MaterialRecipe r = gnxsk::material_recipe(om);
                gnx_material gm = make_material(GNX_MAT_MATTE, GNX_MATF_BUMP_IDENTITY);
                if (r.kind == gnxsk::MaterialRecipe::Matte) { set_rgb(gm, 0, r.kd[0], r.kd[1], r.kd[2]); gm.f[0] = 0.f; }
                else if (r.kind == gnxsk::MaterialRecipe::Plastic) {
                    gm.type = GNX_MAT_PLASTIC;
                    set_rgb(gm, 0, r.kd[0], r.kd[1], r.kd[2]); set_rgb(gm, 1, r.ks[0], r.ks[1], r.ks[2]); gm.f[0] = r.roughness;
                } else if (r.kind == gnxsk::MaterialRecipe::Mirror) { gm.type = GNX_MAT_MIRROR; set_rgb(gm, 0, r.ks[0], r.ks[1], r.ks[2]); }
                else {
                    gm.type = GNX_MAT_GLASS;
                    set_rgb(gm, 0, r.ks[0], r.ks[1], r.ks[2]); set_rgb(gm, 1, r.ks[0], r.ks[1], r.ks[2]);
                    gm.f[0] = gm.f[1] = 0.f; gm.f[2] = r.index;
                }
                sc->materials.push_back(gm);
            }
            if (obj) {
                const float centre[3] = {0.f, -0.4f, 0.f}, zero[3] = {0.f, 0.f, 0.f};
                gnxsk::fit_to_sphere(&file, 2.5f, centre);
                add_mesh(soup, file, 1.f, zero, 0);
            } else add_mesh(soup, file, 20.f, T, 0);
        }
        int w0, h0;
        std::vector<float> rgb;
        std::string path = std::string(resources ? resources : ".") + "/MonValley1000.hdr";
        if (!load_hdr(path, &w0, &h0, rgb, &sc->error)) return sc;
        build_env(rgb, w0, h0, 1.0f, sc->env);
        Mat4 l2w = mul(mul(rotate_axis(0, 20), rotate_axis(1, -90)), rotate_axis(0, -90));
        sc->finalize(soup, width, height, spp, true, l2w);
    } else if (nm == "nano") {
        // config 3 stand-in, oracle/ref_harness.cpp::BuildNano: Disney material on a UV-mapped, smooth-shaded
        // mesh under TropicalRuins1000.hdr.  The oracle side textures it with awesomeface.jpg through the
        // reference's stb loader; the kit has no JPEG decoder, so it paints a procedural 512 x 512 colour map
        // (same code path: ImageTexture level-0 bilinear, Repeat wrap).
        const bool thin = p0 == 1;
        gnx_material m = make_material(GNX_MAT_DISNEY, thin ? GNX_MATF_THIN : 0u);
        m.rgb_tex[0] = 0;
        const float fv[12] = {0.2f, 1.5f, 0.4f, 0.f, thin ? 0.3f : 0.f, 0.5f, 0.5f, 0.5f, 0.8f, thin ? 0.4f : 0.f, thin ? 0.3f : 0.f, thin ? 0.5f : 0.f};
        memcpy(m.f, fv, sizeof(fv));
        sc->materials.push_back(m);
        const int TW = 512;
        sc->texels.resize((size_t)TW * TW * 3);
        for (int y = 0; y < TW; ++y)
            for (int x = 0; x < TW; ++x) {
                float u = (x + .5f) / TW, v = (y + .5f) / TW;
                bool chk = ((x / 64) + (y / 64)) & 1;
                float *q = &sc->texels[((size_t)y * TW + x) * 3];
                q[0] = chk ? 0.9f : 0.15f + 0.5f * u;
                q[1] = chk ? 0.75f * v + 0.1f : 0.6f;
                q[2] = chk ? 0.1f : 0.8f * (1 - v);
            }
        gnx_texture t{};
        t.width = t.height = TW; t.n_channels = 3; t.n_levels = 1; t.wrap = GNX_WRAP_REPEAT; t.max_aniso = 8.f;
        t.su = t.sv = 1.f; t.texels = sc->texels.data();
        sc->textures.push_back(t);
        const int nu = p1 > 0 ? p1 : 320, nv = p2 > 0 ? p2 : 64;
        const float T[3] = {0.f, -2.9f, 0.f};
        add_mesh(soup, gnxsk::torus_knot(nu, nv, 1.0f, true, true), 20.f, T, 0);
        add_mesh(soup, gnxsk::uv_sphere(std::max(8, nu / 4), std::max(6, nv), 0.9f, -2.6f, -1.2f, 0.6f), 1.f, zero, 0);
        add_mesh(soup, gnxsk::uv_sphere(std::max(8, nu / 4), std::max(6, nv), 0.7f, 2.7f, 1.4f, -0.4f), 1.f, zero, 0);
        int w0, h0;
        std::vector<float> rgb;
        std::string path = std::string(resources ? resources : ".") + "/TropicalRuins1000.hdr";
        if (!load_hdr(path, &w0, &h0, rgb, &sc->error)) return sc;
        build_env(rgb, w0, h0, 1.0f, sc->env);
        Mat4 l2w = mul(mul(rotate_axis(0, 20), rotate_axis(1, -90)), rotate_axis(0, -90));
        sc->finalize(soup, width, height, spp, true, l2w);
    } else if (nm == "smoke") {
        // config 4, oracle/ref_harness.cpp::BuildSmoke: GridDensityMedium (density_render.70.volume) inside a
        // HomogeneousMedium fog box, material-less boundary triangles, Matte ground, MonValley environment.
        // p0: 0 = grid + fog with the PCG32 stream sampler, 1 = fog only with Halton.
        const std::string res = resources ? resources : ".";
        gnx_medium fog{};
        fog.type = GNX_MEDIUM_HOMOGENEOUS;
        for (int c = 0; c < 3; ++c) { fog.sigma_a[c] = 0.02f; fog.sigma_s[c] = 0.08f; }
        fog.g = 0.5f;
        sc->media.push_back(fog);
        if (p0 == 0) {
            std::ifstream f(res + "/density_render.70.volume");
            if (!f) { sc->error = "cannot open " + res + "/density_render.70.volume"; return sc; }
            std::string tok;
            int nx = 0, ny = 0, nz = 0;
            float q0[3], q1[3], sa[3], ss[3];
            f >> tok >> nx >> tok >> ny >> tok >> nz;
            f >> tok >> q0[0] >> q0[1] >> q0[2] >> tok >> q1[0] >> q1[1] >> q1[2];
            f >> tok >> sa[0] >> sa[1] >> sa[2] >> tok >> ss[0] >> ss[1] >> ss[2];
            sc->density.resize((size_t)nx * ny * nz);
            float mx = 0;
            for (float &d : sc->density) { f >> d; mx = std::max(mx, d); }
            if (!f || nx <= 0) { sc->error = "cannot parse the .volume file"; return sc; }
            gnx_medium grid{};
            grid.type = GNX_MEDIUM_GRID;
            for (int c = 0; c < 3; ++c) { grid.sigma_a[c] = sa[0]; grid.sigma_s[c] = ss[0]; }
            grid.g = 0.f; grid.nx = nx; grid.ny = ny; grid.nz = nz;
            grid.density = sc->density.data();
            grid.inv_max_density = 1 / mx;
            Mat4 m2w = mul(mul(translate(-1, -1, -0.4f), translate(q0[0], q0[1], q0[2])), scale(q1[0] - q0[0], q1[1] - q0[1], q1[2] - q0[2]));
            to_float16(inverse(m2w), grid.world_to_medium);
            sc->media.push_back(grid);
            const float lo[3] = {-1.f + q0[0] - 0.01f, -1.f + q0[1] - 0.01f, -0.4f + q0[2] - 0.01f};
            const float hi[3] = {-1.f + q1[0] + 0.01f, -1.f + q1[1] + 0.01f, -0.4f + q1[2] + 0.01f};
            add_mesh(soup, gnxsk::box(lo, hi), 1.f, zero, -1, 1, 0);
            sc->pcgSampler = true;
        }
        const float flo[3] = {-2.4f, -2.4f, -2.4f}, fhi[3] = {2.4f, 2.4f, 2.4f};
        add_mesh(soup, gnxsk::box(flo, fhi), 1.f, zero, -1, 0, -1);
        gnx_material m = make_material(GNX_MAT_MATTE, GNX_MATF_BUMP_IDENTITY);
        set_rgb(m, 0, 0.5f, 0.5f, 0.5f);
        sc->materials.push_back(m);
        add_mesh(soup, gnxsk::ground_quad(2.2f, -1.5f), 1.f, zero, 0, 0, 0);
        int w0, h0;
        std::vector<float> rgb;
        if (!load_hdr(res + "/MonValley1000.hdr", &w0, &h0, rgb, &sc->error)) return sc;
        build_env(rgb, w0, h0, 1.0f, sc->env);
        Mat4 l2w = mul(mul(rotate_axis(0, 20), rotate_axis(1, -90)), rotate_axis(0, -90));
        sc->finalize(soup, width, height, spp, true, l2w);
    } else {
        sc->error = "unknown scene '" + nm + "'";
    }
    return sc;
}

void gnxsk_destroy(gnxsk_scene *s) { delete s; }
const char *gnxsk_error(const gnxsk_scene *s) { return s ? s->error.c_str() : "null scene"; }
const gnx_scene_desc *gnxsk_desc(const gnxsk_scene *s) { return (s && s->error.empty()) ? &s->desc : nullptr; }
int gnxsk_num_prims(const gnxsk_scene *s) { return s ? s->desc.geom.n_prims : 0; }
double gnxsk_build_seconds(const gnxsk_scene *s) { return s ? s->build_seconds : 0; }
// Mesh files without a scene: 0 on success, else -1 with the reason in err (when err_len > 0).
static int io_result(bool ok, const std::string &e, char *err, int err_len) {
    if (!ok && err && err_len > 0) { strncpy(err, e.c_str(), (size_t)err_len - 1); err[err_len - 1] = 0; }
    return ok ? 0 : -1;
}
int gnxsk_write_knot_3d(const char *path, int nu, int nv, char *err, int err_len) {
    std::string e;
    return io_result(gnxsk::save_3d(path, gnxsk::torus_knot(nu, nv), &e), e, err, err_len);
}
int gnxsk_mesh_info(const char *path, int *n_vertices, int *n_triangles, int *has_uv, int *has_normals, char *err, int err_len) {
    std::string e, p(path);
    gnxsk::Mesh m;
    const bool obj = p.size() > 4 && p.compare(p.size() - 4, 4, ".obj") == 0;
    bool ok = obj ? gnxsk::load_obj(p, &m, &e) : gnxsk::load_3d(p, &m, &e);
    if (ok) {
        if (n_vertices) *n_vertices = m.nVerts();
        if (n_triangles) *n_triangles = m.nTris();
        if (has_uv) *has_uv = !m.UV.empty();
        if (has_normals) *has_normals = !m.N.empty();
    }
    return io_result(ok, e, err, err_len);
}
// Damages the description in one specific way (tests of the library's upload validation): 1 = a BVH leaf whose
// primitive range leaves the array, 2 = prim_light past the light list, 3 = a texture index below -1, 4 = materials NULL
// with a non-zero count, 5 = prim_light pointing at a light that is not an area light.  Returns 0 when applied.
int gnxsk_corrupt(gnxsk_scene *s, int kind) {
    if (!s || !s->error.empty()) return -1;
    gnx_scene_desc &d = s->desc;
    if (kind == 1) {
        for (gnx_bvh_node &n : s->nodes) if (n.n_prims > 0) { n.offset = d.geom.n_prims - n.n_prims + 1; return 0; }
        return -1;
    }
    if (kind == 2 && !s->prim_light.empty()) { s->prim_light[0] = d.n_lights; return 0; }
    if (kind == 3 && !s->materials.empty()) { s->materials[0].rgb_tex[0] = -5; return 0; }
    if (kind == 4) { d.materials = nullptr; return 0; }
    if (kind == 5 && !s->prim_light.empty()) {
        for (int i = 0; i < d.n_lights; ++i) if (s->lights[i].type != GNX_LIGHT_AREA_TRI) { s->prim_light[0] = i; return 0; }
        return -1;
    }
    return -1;
}

void gnxsk_strip_bvh(gnxsk_scene *s) {
    if (!s) return;
    s->desc.geom.n_nodes = 0;
    s->desc.geom.nodes = nullptr;
}

}  // extern "C"
