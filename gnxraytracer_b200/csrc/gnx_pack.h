// gnx_pack.h — host-side packing of a gnx_scene_desc into the layouts the kernels read.  Shared by
// gnx_upload_scene (gnx_render.cu) and by the CPU emulation used in tests (tests/emul), so both see
// bit-identical tables.  Host code only.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "gnx_scene.cuh"
#include "gnx_sampler.cuh"

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

// UniformLightDistribution: Distribution1D over n ones (core/LightDistribution.cpp:35-38, core/Sampling.h:22-35)
inline float uniform_light_distribution(int n, std::vector<float> &func, std::vector<float> &cdf) {
    func.assign((size_t)n, 1.f);
    cdf.assign((size_t)n + 1, 0.f);
    for (int i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + func[i - 1] / n;
    float funcInt = cdf[n];
    for (int i = 1; i < n + 1; ++i) cdf[i] /= funcInt;
    return funcInt;
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
