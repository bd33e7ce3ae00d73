// gnx_pack.h — host-side packing of a gnx_scene_desc into the layouts the kernels read.  Shared by
// gnx_upload_scene (gnx_render.cu) and by the CPU emulation used in tests (tests/emul), so both see
// bit-identical tables.  Host code only.
#pragma once
#include <algorithm>
#include <cmath>
#include <functional>
#include <cstring>
#include <string>
#include <vector>

#include "gnx_scene.cuh"
#include "gnx_sampler.cuh"
#include "gnx_bvh.cuh"
#include "gnx_bvh8.cuh"

namespace gnx {

// ---- Halton tables, generated exactly as the reference does ----------------------------------------------
// Primes / PrimeSums are the first 1000 primes and their running sums (samplers/LowDiscrepancy.cpp:9-355);
// the permutations are ComputeRadicalInversePermutations with a default-constructed RNG
// (samplers/LowDiscrepancy.cpp:2459-2473, samplers/HaltonSampler.cpp:36-39, core/Sampling.h:129-137).
inline const int kPrimeTableSize = 1000;
inline void make_primes(std::vector<int> &primes, std::vector<int> &sums) {
    primes.clear();
    std::vector<char> sieve(8200, 1);
    for (int i = 2; i < (int)sieve.size() && (int)primes.size() < kPrimeTableSize; ++i) {
        if (!sieve[i]) continue;
        primes.push_back(i);
        for (int j = i * 2; j < (int)sieve.size(); j += i) sieve[j] = 0;
    }
    sums.assign(primes.size(), 0);
    int s = 0;
    for (size_t i = 0; i < primes.size(); ++i) { sums[i] = s; s += primes[i]; }
}
// Per-dimension record of the scrambled radical inverse: the prime, the offset of its permutation in the
// concatenated table (PrimeSums, samplers/LowDiscrepancy.cpp:86-) and the exact-division constant of
// gnx_sampler.cuh.
inline void make_dim_table(const std::vector<int> &primes, const std::vector<int> &sums, std::vector<uint4> &dims) {
    dims.resize(primes.size());
    for (size_t i = 0; i < primes.size(); ++i) {
        uint64_t d = (uint64_t)primes[i];
        uint64_t m = ((1ull << 38) + d - 1) / d;
        dims[i] = make_uint4((uint32_t)primes[i], (uint32_t)sums[i], (uint32_t)(m & 0xffffffffu), (uint32_t)(m >> 32));
    }
}
// Environment-map texels [h][w][3] -> one 16-byte record per texel (a bilinear lookup gathers four
// unrelated texels: 4 x LDG.128 instead of 12 scalar loads).
inline void pack_env_texels(const float *rgb, size_t n, std::vector<float4> &out) {
    out.resize(n);
    for (size_t i = 0; i < n; ++i) out[i] = make_float4(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], 0.f);
}
// Guide table of find_interval_guided for one cdf of `size` entries (size - 1 <= 65535).
inline void make_cdf_guide(const float *cdf, int size, int G, uint16_t *guide) {
    for (int b = 0; b <= G; ++b) guide[b] = (uint16_t)find_interval_cdf(cdf, size, (float)b / (float)G);
}
inline int guide_buckets(int n) {  // power of two, about n / 8 buckets
    int g = 1;
    while (g * 8 < n) g <<= 1;
    return g;
}
inline uint32_t pcg_bounded(Pcg32 &rng, uint32_t b) {
    uint32_t threshold = (~b + 1u) % b;
    while (true) {
        uint32_t r = rng.next_u32();
        if (r >= threshold) return r % b;
    }
}
inline void make_permutations(const std::vector<int> &primes, std::vector<uint16_t> &perms) {
    size_t total = 0;
    for (int p : primes) total += p;
    perms.resize(total);
    Pcg32 rng;
    rng.state = 0x853c49e6748fea9bULL;  // PCG32_DEFAULT_STATE / STREAM, core/RNG.h:26-27,95
    rng.inc = 0xda3e39cb94b95bdbULL;
    uint16_t *p = perms.data();
    for (int prime : primes) {
        for (int j = 0; j < prime; ++j) p[j] = (uint16_t)j;
        for (int i = 0; i < prime; ++i) {
            int other = i + (int)pcg_bounded(rng, (uint32_t)(prime - i));
            std::swap(p[i], p[other]);
        }
        p += prime;
    }
}


// 48-byte triangle records: a = (p0.xyz, p1.x), b = (p1.yz, p2.xy), c = (p2.z, word, light, id) with
// word = (material + 1) | shade type << 20 | prim flags << 24.
inline bool pack_triangles(const gnx_scene_desc &d, std::vector<float4> &tris, unsigned *typeMask, std::string *err) {
    const gnx_geometry &g = d.geom;
    tris.resize((size_t)g.n_prims * 3);
    *typeMask = 0;
    for (int k = 0; k < g.n_prims; ++k) {
        const float *p = g.prim_p + (size_t)k * 9;
        int mat = g.prim_material[k];
        if (mat >= d.n_materials) { *err = "prim_material out of range"; return false; }
        int type = mat < 0 ? kNumShadeTypes - 1 : d.materials[mat].type;
        *typeMask |= 1u << type;
        unsigned flags = g.prim_flags ? g.prim_flags[k] : 0u;
        unsigned word = (unsigned)(mat + 1) | ((unsigned)type << 20) | (flags << 24);
        int light = g.prim_light ? g.prim_light[k] : -1;
        int id = g.prim_id ? g.prim_id[k] : k;
        float fw, fl, fi;
        memcpy(&fw, &word, 4); memcpy(&fl, &light, 4); memcpy(&fi, &id, 4);
        tris[3 * (size_t)k + 0] = make_float4(p[0], p[1], p[2], p[3]);
        tris[3 * (size_t)k + 1] = make_float4(p[4], p[5], p[6], p[7]);
        tris[3 * (size_t)k + 2] = make_float4(p[8], fw, fl, fi);
    }
    return true;
}

// LinearBVHNode array (32 B nodes, depth-first, first child = index + 1) -> Node2 array (gnx_bvh.cuh): one
// 64-byte record per INTERIOR node holding both children's bounds.  Leaves with more than
// kLeafMaxPrims primitives (coincident centroids) become a chain of "c0 first" nodes.
inline bool build_node2(const gnx_bvh_node *nodes, int n, std::vector<float4> &out, std::string *err) {
    out.clear();
    if (n <= 0) return true;
    std::vector<int> id((size_t)n, -1);
    int count = 0;
    for (int i = 0; i < n; ++i) if (nodes[i].n_prims == 0) id[i] = count++;
    struct N2 { float lo0[3], hi0[3], lo1[3], hi1[3]; int ref0, ref1, axis; };
    std::vector<N2> v((size_t)count);
    auto empty = [](float *lo, float *hi) { for (int c = 0; c < 3; ++c) { lo[c] = 0; hi[c] = 0; } };
    // leaf reference, chaining oversized leaves through extra nodes appended at the end
    auto leafRef = [&](const gnx_bvh_node &lf) -> int {
        int offset = lf.offset, cnt = lf.n_prims;
        if (cnt <= kLeafMaxPrims) return leaf_ref(offset, cnt);
        int first = -1, prev = -1;
        while (cnt > 0) {
            int take = std::min(cnt, kLeafMaxPrims);
            N2 c{};
            memcpy(c.lo0, lf.bmin, 12); memcpy(c.hi0, lf.bmax, 12);
            memcpy(c.lo1, lf.bmin, 12); memcpy(c.hi1, lf.bmax, 12);
            c.ref0 = leaf_ref(offset, take); c.ref1 = kRefNone; c.axis = 3;
            v.push_back(c);
            int me = (int)v.size() - 1;
            if (first < 0) first = me; else v[prev].ref1 = me;
            prev = me; offset += take; cnt -= take;
        }
        return first;
    };
    if (nodes[0].n_prims > 0) {  // single-leaf tree
        N2 r{};
        memcpy(r.lo0, nodes[0].bmin, 12); memcpy(r.hi0, nodes[0].bmax, 12);
        empty(r.lo1, r.hi1);
        r.ref1 = kRefNone; r.axis = 3;
        v.push_back(r);
        int ref = leafRef(nodes[0]);
        v[0].ref0 = ref;
    }
    for (int i = 0; i < n; ++i) {
        if (nodes[i].n_prims != 0) continue;
        const int c0 = i + 1, c1 = nodes[i].offset;
        if (c0 >= n || c1 <= i || c1 >= n) { *err = "malformed BVH node array"; return false; }
        N2 r{};
        memcpy(r.lo0, nodes[c0].bmin, 12); memcpy(r.hi0, nodes[c0].bmax, 12);
        memcpy(r.lo1, nodes[c1].bmin, 12); memcpy(r.hi1, nodes[c1].bmax, 12);
        r.axis = nodes[i].axis > 2 ? 0 : nodes[i].axis;
        r.ref0 = nodes[c0].n_prims ? leafRef(nodes[c0]) : id[c0];
        r.ref1 = nodes[c1].n_prims ? leafRef(nodes[c1]) : id[c1];
        v[id[i]] = r;
    }
    out.resize(v.size() * 4);
    for (size_t k = 0; k < v.size(); ++k) {
        const N2 &r = v[k];
        float f0, f1, fa;
        memcpy(&f0, &r.ref0, 4); memcpy(&f1, &r.ref1, 4); memcpy(&fa, &r.axis, 4);
        out[4 * k + 0] = make_float4(r.lo0[0], r.lo0[1], r.lo0[2], r.hi0[0]);
        out[4 * k + 1] = make_float4(r.hi0[1], r.hi0[2], r.lo1[0], r.lo1[1]);
        out[4 * k + 2] = make_float4(r.lo1[2], r.hi1[0], r.hi1[1], r.hi1[2]);
        out[4 * k + 3] = make_float4(f0, f1, fa, 0.f);
    }
    return true;
}

// Same input, 4-wide: every Node4 (gnx_bvh.cuh) covers an interior LinearBVHNode X and its two children, i.e.
// holds the bounds of X's (up to) four grandchildren, slots 0-1 under X's first child and slots 2-3 under its
// second child; a child that is itself a leaf takes one slot of its pair.  Halves the dependent fetches of
// the two-child layout without changing which leaves are visited or in which order.
inline bool build_node4(const gnx_bvh_node *nodes, int n, std::vector<float4> &out, std::string *err) {
    out.clear();
    if (n <= 0) return true;
    struct N4 { float lo[3][4], hi[3][4]; int ref[4]; int axTop, ax0, ax1; };
    std::vector<N4> v;
    auto blank = [] { N4 r{}; for (int k = 0; k < 4; ++k) r.ref[k] = kRefNone; r.axTop = r.ax0 = r.ax1 = 3; return r; };
    auto setBox = [](N4 &r, int slot, const gnx_bvh_node &b) { for (int c = 0; c < 3; ++c) { r.lo[c][slot] = b.bmin[c]; r.hi[c][slot] = b.bmax[c]; } };
    // chain of "slot order" nodes for a leaf with more than kLeafMaxPrims primitives; returns a child reference
    std::function<int(const gnx_bvh_node &, int, int)> leafRef = [&](const gnx_bvh_node &lf, int offset, int cnt) -> int {
        if (cnt <= kLeafMaxPrims) return leaf_ref(offset, cnt);
        int me = (int)v.size();
        v.push_back(blank());
        setBox(v[me], 0, lf); setBox(v[me], 2, lf);
        v[me].ref[0] = leaf_ref(offset, kLeafMaxPrims);
        int rest = leafRef(lf, offset + kLeafMaxPrims, cnt - kLeafMaxPrims);
        v[me].ref[2] = rest;
        return me;
    };
    std::vector<std::pair<int, int>> todo;  // (source interior node, Node4 index)
    v.push_back(blank());
    if (nodes[0].n_prims > 0) {
        setBox(v[0], 0, nodes[0]);
        int r = leafRef(nodes[0], nodes[0].offset, nodes[0].n_prims);
        v[0].ref[0] = r;
    } else todo.push_back({0, 0});
    while (!todo.empty()) {
        auto [x, me] = todo.back();
        todo.pop_back();
        const int c[2] = {x + 1, nodes[x].offset};
        if (c[0] >= n || c[1] <= x || c[1] >= n) { *err = "malformed BVH node array"; return false; }
        v[me].axTop = nodes[x].axis > 2 ? 0 : nodes[x].axis;
        for (int k = 0; k < 2; ++k) {
            const gnx_bvh_node &C = nodes[c[k]];
            int &ax = k == 0 ? v[me].ax0 : v[me].ax1;
            if (C.n_prims > 0) {
                ax = 3;
                setBox(v[me], 2 * k, C);
                int r = leafRef(C, C.offset, C.n_prims);
                v[me].ref[2 * k] = r;
            } else {
                ax = C.axis > 2 ? 0 : C.axis;
                const int g[2] = {c[k] + 1, C.offset};
                if (g[0] >= n || g[1] <= c[k] || g[1] >= n) { *err = "malformed BVH node array"; return false; }
                for (int j = 0; j < 2; ++j) {
                    const gnx_bvh_node &G = nodes[g[j]];
                    setBox(v[me], 2 * k + j, G);
                    if (G.n_prims > 0) { int r = leafRef(G, G.offset, G.n_prims); v[me].ref[2 * k + j] = r; }
                    else {
                        int idx = (int)v.size();
                        v.push_back(blank());
                        v[me].ref[2 * k + j] = idx;
                        todo.push_back({g[j], idx});
                    }
                }
            }
        }
    }
    out.resize(v.size() * 8);
    for (size_t i = 0; i < v.size(); ++i) {
        const N4 &r = v[i];
        for (int c = 0; c < 3; ++c) {
            out[8 * i + c] = make_float4(r.lo[c][0], r.lo[c][1], r.lo[c][2], r.lo[c][3]);
            out[8 * i + 3 + c] = make_float4(r.hi[c][0], r.hi[c][1], r.hi[c][2], r.hi[c][3]);
        }
        float f[4], a[3];
        for (int k = 0; k < 4; ++k) memcpy(&f[k], &r.ref[k], 4);
        memcpy(&a[0], &r.axTop, 4); memcpy(&a[1], &r.ax0, 4); memcpy(&a[2], &r.ax1, 4);
        out[8 * i + 6] = make_float4(f[0], f[1], f[2], f[3]);
        out[8 * i + 7] = make_float4(a[0], a[1], a[2], 0.f);
    }
    return true;
}

// MIPMap pyramid from a power-of-two level 0 (core/MIPMap.h:161-186): level i is the 2 x 2 box filter of level i - 1
// through MIPMap::Texel's wrap rule.  `texels` receives all levels back to back, `offsets` each level's start in texels.
// The caller's own levels (n_levels > 1: the reference's pyramid, read by the bridge) are copied as they are.
inline void build_mip_pyramid(const gnx_texture &t, std::vector<float> &texels, std::vector<int> &offsets, int *nLevelsOut) {
    const int nch = t.n_channels;
    auto levelW = [&](int l) { return std::max(1, t.width >> l); };
    auto levelH = [&](int l) { return std::max(1, t.height >> l); };
    const bool pow2 = (t.width & (t.width - 1)) == 0 && (t.height & (t.height - 1)) == 0;
    int nLevels = t.n_levels;
    if (nLevels <= 1) {
        nLevels = 1;
        if (pow2) { int m = std::max(t.width, t.height); while ((1 << nLevels) <= m) ++nLevels; }  // 1 + Log2Int(max)
    }
    offsets.assign(nLevels, 0);
    size_t total = 0;
    for (int l = 0; l < nLevels; ++l) { offsets[l] = (int)total; total += (size_t)levelW(l) * levelH(l); }
    texels.resize(total * nch);
    if (t.n_levels > 1) { memcpy(texels.data(), t.texels, total * nch * sizeof(float)); *nLevelsOut = nLevels; return; }
    memcpy(texels.data(), t.texels, (size_t)t.width * t.height * nch * sizeof(float));
    for (int l = 1; l < nLevels; ++l) {
        const int pw = levelW(l - 1), ph = levelH(l - 1), w = levelW(l), h = levelH(l);
        const float *src = texels.data() + (size_t)offsets[l - 1] * nch;
        float *dst = texels.data() + (size_t)offsets[l] * nch;
        auto texel = [&](int s, int tt, int c) -> float {
            if (t.wrap == GNX_WRAP_REPEAT) { s %= pw; if (s < 0) s += pw; tt %= ph; if (tt < 0) tt += ph; }
            else if (t.wrap == GNX_WRAP_CLAMP) { s = std::min(std::max(s, 0), pw - 1); tt = std::min(std::max(tt, 0), ph - 1); }
            else if (s < 0 || s >= pw || tt < 0 || tt >= ph) return 0.f;
            return src[((size_t)tt * pw + s) * nch + c];
        };
        for (int tt = 0; tt < h; ++tt)
            for (int ss = 0; ss < w; ++ss)
                for (int c = 0; c < nch; ++c)
                    dst[((size_t)tt * w + ss) * nch + c] = .25f * (texel(2 * ss, 2 * tt, c) + texel(2 * ss + 1, 2 * tt, c) +
                                                                   texel(2 * ss, 2 * tt + 1, c) + texel(2 * ss + 1, 2 * tt + 1, c));
    }
    *nLevelsOut = nLevels;
}
// MIPMap::weightLut (core/MIPMap.h:189-196)
inline void make_ewa_lut(float lut[128]) {
    for (int i = 0; i < 128; ++i) {
        float alpha = 2;
        float r2 = float(i) / float(128 - 1);
        lut[i] = std::exp(-alpha * r2) - std::exp(-alpha);
    }
}
inline void fill_dev_texture(const gnx_texture &t, const float *texels, const std::vector<int> &offsets, int nLevels, DevTexture &o) {
    o.w = t.width; o.h = t.height; o.nch = t.n_channels; o.wrap = t.wrap;
    o.su = t.su; o.sv = t.sv; o.du = t.du; o.dv = t.dv;
    o.texels = texels;
    o.n_levels = nLevels; o.do_trilinear = t.do_trilinear; o.max_aniso = t.max_aniso;
    for (int l = 0; l < kMaxMipLevels; ++l) o.level_off[l] = l < nLevels ? offsets[l] : 0;
}

// Index hygiene at the ABI boundary: every index the kernels will dereference is checked here, once, so that a bad
// description is GNX_ERR_INVALID at upload instead of an out-of-range device read at render time.
inline bool validate_scene_desc(const gnx_scene_desc &d, std::string *err) {
    const gnx_geometry &g = d.geom;
    if (d.n_materials < 0 || d.n_textures < 0 || d.n_lights < 0 || d.n_media < 0) { *err = "negative array count"; return false; }
    if (d.n_materials > 0 && !d.materials) { *err = "materials is NULL with n_materials > 0"; return false; }
    if (d.n_textures > 0 && !d.textures) { *err = "textures is NULL with n_textures > 0"; return false; }
    if (d.n_lights > 0 && !d.lights) { *err = "lights is NULL with n_lights > 0"; return false; }
    if (d.n_media > 0 && !d.media) { *err = "media is NULL with n_media > 0"; return false; }
    for (int i = 0; i < g.n_nodes; ++i) {
        const gnx_bvh_node &nd = g.nodes[i];
        if (nd.n_prims == 0) continue;
        if (nd.offset < 0 || (long long)nd.offset + nd.n_prims > (long long)g.n_prims) { *err = "BVH leaf primitive range out of bounds"; return false; }
    }
    for (int k = 0; k < g.n_prims; ++k) {
        if (g.prim_material[k] < -1 || g.prim_material[k] >= d.n_materials) { *err = "prim_material out of range"; return false; }
        if (g.prim_light) {
            const int l = g.prim_light[k];
            if (l < -1 || l >= d.n_lights) { *err = "prim_light out of range"; return false; }
            if (l >= 0 && d.lights[l].type != GNX_LIGHT_AREA_TRI) { *err = "prim_light points at a light that is not an area light"; return false; }
        }
        if (g.prim_medium_in && g.prim_medium_out &&
            (g.prim_medium_in[k] < -1 || g.prim_medium_in[k] >= d.n_media || g.prim_medium_out[k] < -1 || g.prim_medium_out[k] >= d.n_media)) {
            *err = "prim_medium index out of range"; return false;
        }
    }
    for (int i = 0; i < d.n_materials; ++i) {
        for (int s = 0; s < GNX_MAT_MAX_RGB; ++s)
            if (d.materials[i].rgb_tex[s] < -1 || d.materials[i].rgb_tex[s] >= d.n_textures) { *err = "texture index out of range"; return false; }
        for (int s = 0; s < GNX_MAT_MAX_F; ++s)
            if (d.materials[i].f_tex[s] < -1 || d.materials[i].f_tex[s] >= d.n_textures) { *err = "texture index out of range"; return false; }
    }
    for (int i = 0; i < d.n_lights; ++i)
        if (d.lights[i].medium < -1 || d.lights[i].medium >= d.n_media) { *err = "light medium index out of range"; return false; }
    if (d.camera.medium < -1 || d.camera.medium >= d.n_media) { *err = "camera medium index out of range"; return false; }
    return true;
}

// Packs the node array in the layout the library was compiled for (GNX_BVH_WIDTH).
inline bool build_nodes(const gnx_bvh_node *nodes, int n, std::vector<float4> &out, int *count, std::string *err) {
#if GNX_BVH_WIDTH == 4
    if (!build_node4(nodes, n, out, err)) return false;
    *count = (int)(out.size() / 8);
#else
    if (!build_node2(nodes, n, out, err)) return false;
    *count = (int)(out.size() / 4);
#endif
    return true;
}

// Node2 array (64-byte two-child records, gnx_bvh.cuh) -> compressed 8-wide tree (gnx_bvh8.cuh) over the same ordered
// primitives, for the any-hit queries.  A Node8 starts from the two children of its source node and keeps replacing the
// internal child of largest surface area by that child's own two children until it holds eight (or only leaves are
// left); leaves keep their reference; the children then take octant slots (front-to-back visiting order by XOR).  Child boxes are rounded OUTWARD onto the node's grid (origin = corner of the union,
// step = the smallest power of two that spans it in 255 steps), checked in double precision.  Returns false (and leaves
// `out` empty) for bounds that are not finite: the caller then keeps the two-child tree for every query.
inline bool build_node8(const float4 *n2, int count, std::vector<uint4> &out) {
    out.clear();
    if (count <= 0) return true;
    struct Child { float lo[3], hi[3]; int ref; };
    auto children2 = [&](int i, Child *c) {
        const float4 a = n2[4 * (size_t)i], b = n2[4 * (size_t)i + 1], d = n2[4 * (size_t)i + 2], r = n2[4 * (size_t)i + 3];
        int ref0, ref1, n = 0;
        memcpy(&ref0, &r.x, 4); memcpy(&ref1, &r.y, 4);
        if (ref0 != kRefNone) { c[n] = Child{{a.x, a.y, a.z}, {a.w, b.x, b.y}, ref0}; ++n; }
        if (ref1 != kRefNone) { c[n] = Child{{b.z, b.w, d.x}, {d.y, d.z, d.w}, ref1}; ++n; }
        return n;
    };
    auto area = [](const Child &c) {
        const double dx = (double)c.hi[0] - c.lo[0], dy = (double)c.hi[1] - c.lo[1], dz = (double)c.hi[2] - c.lo[2];
        return dx * dy + dy * dz + dz * dx;
    };
    std::vector<std::pair<int, int>> todo;  // (source Node2, Node8 index)
    out.resize(kNode8Words);
    todo.push_back({0, 0});
    while (!todo.empty()) {
        const auto [src, me] = todo.back();
        todo.pop_back();
        if (src < 0 || src >= count) { out.clear(); return false; }
        Child c[8];
        int n = children2(src, c);
        if (n == 0) { out.clear(); return false; }
        while (n < 8) {
            int pick = -1;
            double best = -1;
            for (int k = 0; k < n; ++k) if (c[k].ref >= 0) { const double a = area(c[k]); if (a > best || pick < 0) { best = a; pick = k; } }
            if (pick < 0) break;
            if (c[pick].ref >= count) { out.clear(); return false; }
            Child g[2];
            const int m = children2(c[pick].ref, g);
            if (m == 0) { c[pick] = c[--n]; continue; }
            c[pick] = g[0];
            if (m > 1) c[n++] = g[1];
        }
        // octant slots (gnx_bvh8.cuh): the child lying furthest towards (+-x, +-y, +-z) of the node's centre takes that
        // corner's slot, greedily over the 8 x n (child, slot) pairs
        {
            double ctr[3], cc[8][3];
            for (int ax = 0; ax < 3; ++ax) {
                double lo = c[0].lo[ax], hi = c[0].hi[ax];
                for (int k = 1; k < n; ++k) { lo = std::min(lo, (double)c[k].lo[ax]); hi = std::max(hi, (double)c[k].hi[ax]); }
                ctr[ax] = 0.5 * (lo + hi);
                for (int k = 0; k < n; ++k) cc[k][ax] = 0.5 * ((double)c[k].lo[ax] + (double)c[k].hi[ax]) - ctr[ax];
            }
            Child placed[8];
            bool slotUsed[8] = {}, childUsed[8] = {};
            int slotOf[8];
            for (int round = 0; round < n; ++round) {
                int bc = -1, bs = -1;
                double best = 0;
                for (int k = 0; k < n; ++k) {
                    if (childUsed[k]) continue;
                    for (int sl = 0; sl < 8; ++sl) {
                        if (slotUsed[sl]) continue;
                        const double v = ((sl & 1) ? cc[k][0] : -cc[k][0]) + ((sl & 2) ? cc[k][1] : -cc[k][1]) + ((sl & 4) ? cc[k][2] : -cc[k][2]);
                        if (bc < 0 || v > best) { best = v; bc = k; bs = sl; }
                    }
                }
                childUsed[bc] = true; slotUsed[bs] = true; slotOf[bc] = bs;
            }
            for (int k = 0; k < n; ++k) placed[slotOf[k]] = c[k];
            for (int sl = 0; sl < 8; ++sl) if (!slotUsed[sl]) { placed[sl] = c[0]; placed[sl].ref = kRefNone; }  // empty slot
            for (int sl = 0; sl < 8; ++sl) c[sl] = placed[sl];
        }
        float org[3];
        int ebyte[3];
        uint8_t qlo[3][8], qhi[3][8];
        bool occ[8];
        for (int k = 0; k < 8; ++k) occ[k] = c[k].ref != kRefNone;
        for (int ax = 0; ax < 3; ++ax) {
            float lo = 0, hi = 0;
            bool first = true;
            for (int k = 0; k < 8; ++k) {
                if (!occ[k]) continue;
                if (!std::isfinite(c[k].lo[ax]) || !std::isfinite(c[k].hi[ax]) || c[k].hi[ax] < c[k].lo[ax]) { out.clear(); return false; }
                lo = first ? c[k].lo[ax] : std::min(lo, c[k].lo[ax]);
                hi = first ? c[k].hi[ax] : std::max(hi, c[k].hi[ax]);
                first = false;
            }
            org[ax] = lo;
            const double ext = (double)hi - (double)lo;
            int e = -126;
            if (ext > 0) { int fe; std::frexp(ext / 255.0, &fe); e = std::max(-126, fe - 1); }  // 2^(fe-1) <= ext/255 < 2^fe
            for (;; ++e) {
                if (e > 127) { out.clear(); return false; }
                const double s = std::ldexp(1.0, e);
                bool ok = true;
                for (int k = 0; k < 8 && ok; ++k) {
                    if (!occ[k]) { qlo[ax][k] = 255; qhi[ax][k] = 0; continue; }
                    double ql = std::floor(((double)c[k].lo[ax] - (double)lo) / s), qh = std::ceil(((double)c[k].hi[ax] - (double)lo) / s);
                    while (ql > 0 && (double)lo + ql * s > (double)c[k].lo[ax]) ql -= 1;
                    while ((double)lo + qh * s < (double)c[k].hi[ax]) qh += 1;
                    if (ql < 0) ql = 0;
                    if (qh > 255) { ok = false; break; }
                    qlo[ax][k] = (uint8_t)ql; qhi[ax][k] = (uint8_t)qh;
                }
                if (ok) break;
            }
            ebyte[ax] = e + 127;
        }
        int refs[8];
        unsigned valid = 0;
        for (int k = 0; k < 8; ++k) {
            refs[k] = kRefNone;
            if (!occ[k]) continue;
            valid |= 1u << k;
            if (c[k].ref < 0) refs[k] = c[k].ref;
            else {
                refs[k] = (int)(out.size() / kNode8Words);
                out.resize(out.size() + kNode8Words);
                todo.push_back({c[k].ref, refs[k]});
            }
        }
        auto pack4 = [](const uint8_t *b) { return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24); };
        uint32_t ob[3];
        memcpy(ob, org, 12);
        uint4 *w = out.data() + (size_t)kNode8Words * me;
        w[0] = make_uint4(ob[0], ob[1], ob[2], (uint32_t)ebyte[0] | ((uint32_t)ebyte[1] << 8) | ((uint32_t)ebyte[2] << 16) | (valid << 24));
        w[1] = make_uint4((uint32_t)refs[0], (uint32_t)refs[1], (uint32_t)refs[2], (uint32_t)refs[3]);
        w[2] = make_uint4((uint32_t)refs[4], (uint32_t)refs[5], (uint32_t)refs[6], (uint32_t)refs[7]);
        w[3] = make_uint4(pack4(qlo[0]), pack4(qlo[0] + 4), pack4(qlo[1]), pack4(qlo[1] + 4));
        w[4] = make_uint4(pack4(qlo[2]), pack4(qlo[2] + 4), pack4(qhi[0]), pack4(qhi[0] + 4));
        w[5] = make_uint4(pack4(qhi[1]), pack4(qhi[1] + 4), pack4(qhi[2]), pack4(qhi[2] + 4));
    }
    return true;
}

// gnx_medium -> DevMedium (density pointer supplied by the caller: host for the emulation, device for upload)
inline void fill_dev_medium(const gnx_medium &m, const float *density, DevMedium &dm) {
    memset(&dm, 0, sizeof(dm));
    dm.type = m.type;
    for (int c = 0; c < 3; ++c) {
        dm.sigma_a[c] = m.sigma_a[c]; dm.sigma_s[c] = m.sigma_s[c];
        dm.sigma_t[c] = m.sigma_s[c] + m.sigma_a[c];  // HomogeneousMedium: sigma_s + sigma_a (media/HomogeneousMedium.h)
    }
    dm.g = m.g;
    if (m.type == GNX_MEDIUM_GRID) {
        dm.nx = m.nx; dm.ny = m.ny; dm.nz = m.nz;
        dm.density = density;
        memcpy(dm.w2m.m, m.world_to_medium, 64);
        dm.inv_max_density = m.inv_max_density;
        dm.sigma_t_scalar = m.sigma_a[0] + m.sigma_s[0];  // (sigma_a + sigma_s)[0], media/GridDensityMedium.h:32
    }
}

// UniformLightDistribution: Distribution1D over n ones (core/LightDistribution.cpp:35-38, core/Sampling.h:22-35)
inline float uniform_light_distribution(int n, std::vector<float> &func, std::vector<float> &cdf) {
    func.assign((size_t)n, 1.f);
    cdf.assign((size_t)n + 1, 0.f);
    for (int i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + func[i - 1] / n;
    float funcInt = cdf[n];
    for (int i = 1; i < n + 1; ++i) cdf[i] /= funcInt;
    return funcInt;
}

// PowerLightDistribution: Distribution1D over Light::Power().y() (core/Integrator.cpp:212-220,
// core/Sampling.h:22-35; an all-zero table falls back to a linear cdf like the reference's).
inline float power_light_distribution(int n, const float *power, std::vector<float> &func, std::vector<float> &cdf) {
    func.assign(power, power + n);
    cdf.assign((size_t)n + 1, 0.f);
    for (int i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + func[i - 1] / n;
    float funcInt = cdf[n];
    if (funcInt == 0) for (int i = 1; i < n + 1; ++i) cdf[i] = float(i) / float(n);
    else for (int i = 1; i < n + 1; ++i) cdf[i] /= funcInt;
    return funcInt;
}

// Light::Power().y() from the flattened records, for callers that pass no light_power table:
// DiffuseAreaLight (lights/DiffuseAreaLight.cpp:32-35), PointLight (lights/PointLight.cpp:24),
// SpotLight (lights/SpotLight.cpp:42-45), DistantLight (lights/DistantLight.cpp:28-31).
// Returns false for light types whose power depends on private state the record does not hold.
inline bool derive_light_power(const gnx_light &l, float *out) {
    const float kPiF = 3.14159265358979323846f;
    // Spectrum::y(), core/Spectrum.h (RGBSpectrum): YWeight = {0.212671, 0.715160, 0.072169}
    auto lum = [](float r, float g, float b) { return 0.212671f * r + 0.715160f * g + 0.072169f * b; };
    // each product in the reference's association order (left to right)
    float c[3];
    switch (l.type) {
    case GNX_LIGHT_AREA_TRI: for (int i = 0; i < 3; ++i) c[i] = ((l.two_sided ? 2.f : 1.f) * l.L[i]) * l.area * kPiF; break;
    case GNX_LIGHT_POINT: for (int i = 0; i < 3; ++i) c[i] = (4 * kPiF) * l.L[i]; break;
    case GNX_LIGHT_SPOT: for (int i = 0; i < 3; ++i) c[i] = l.L[i] * 2 * kPiF * (1 - .5f * (l.cos_falloff + l.cos_total)); break;
    case GNX_LIGHT_DISTANT: for (int i = 0; i < 3; ++i) c[i] = l.L[i] * kPiF * l.area * l.area; break;
    case GNX_LIGHT_SKYBOX: *out = 0.f; return true;  // SkyBoxLight::Power() is Spectrum(0) (lights/SkyBoxLight.h)
    default: return false;
    }
    *out = lum(c[0], c[1], c[2]);
    return true;
}

// SpatialLightDistribution ctor (core/LightDistribution.cpp:70-87): 64 voxels on the widest axis.
inline size_t spatial_voxel_resolution(const float wb[6], int nvox[3]) {
    float diag[3] = {wb[3] - wb[0], wb[4] - wb[1], wb[5] - wb[2]};
    int mx = (diag[0] > diag[1] && diag[0] > diag[2]) ? 0 : (diag[1] > diag[2] ? 1 : 2);
    float bmax = diag[mx];
    size_t nv = 1;
    for (int i = 0; i < 3; ++i) {
        nvox[i] = std::max(1, (int)std::round(diag[i] / bmax * 64));
        nv *= (size_t)nvox[i];
    }
    return nv;
}

}  // namespace gnx
