// gnx_sampler.cuh — HaltonSampler / GlobalSampler on the device.
//
// Follows samplers/HaltonSampler.cpp:63-94 (GetIndexForSample, SampleDimension),
// samplers/LowDiscrepancy.cpp:358-405 (RadicalInverseSpecialized, ScrambledRadicalInverseSpecialized,
// base-2 bit reversal in double) and samplers/LowDiscrepancy.h:47-56 (InverseRadicalInverse).
// Integer work is exact; the float products are taken in the same order as the reference so the
// sample values are bit-identical (tests/test_sampler_parity.py).
#pragma once
#include "gnx_scene.cuh"

namespace gnx {

constexpr int kMaxResolution = 128;  // samplers/HaltonSampler.cpp:11

GNX_HD uint64_t inverse_radical_inverse(uint64_t inverse, int base, int nDigits) {
    uint64_t index = 0;
    for (int i = 0; i < nDigits; ++i) {
        uint64_t digit = inverse % base;
        inverse /= base;
        index = index * base + digit;
    }
    return index;
}

// HaltonSampler::GetIndexForSample(0) for pixel (px, py): the offset of the pixel's first sample.
// With stride * 256 < 2^31 (always: stride = 128 * 243 at most) every intermediate fits 32 bits, the
// base-2 digit reversal is one BREV and the base-3 one divides by a compile-time constant.
GNX_HD uint64_t halton_pixel_offset(const DevSampler &s, int px, int py) {
    if (s.stride <= 1) return 0;
    int pmx = px % kMaxResolution, pmy = py % kMaxResolution;  // Mod() of non-negative ints
    if ((uint64_t)s.stride * 256u < (1ull << 31)) {
        // InverseRadicalInverse<2>: the low base_exp0 bits of pmx, reversed (pmx < 128 = 2^7 >= scale)
        uint32_t d0 = s.base_exp0 > 0 ? brev32((uint32_t)pmx) >> (32 - s.base_exp0) : 0u;
        uint32_t d1 = 0, y = (uint32_t)pmy;
        for (int i = 0; i < s.base_exp1; ++i) { uint32_t q = y / 3u; d1 = d1 * 3u + (y - q * 3u); y = q; }
        // d < scale, so d * (stride / scale) < stride; times multInverse (< scale <= 243) stays below 2^31
        uint32_t a = d0 * (uint32_t)s.stride_over_scale0 * (uint32_t)s.mult_inv0;
        uint32_t b = d1 * (uint32_t)s.stride_over_scale1 * (uint32_t)s.mult_inv1;
        return (uint64_t)((a + b) % (uint32_t)s.stride);
    }
    uint64_t offset = 0;
    uint64_t d0 = inverse_radical_inverse((uint64_t)pmx, 2, s.base_exp0);
    uint64_t d1 = inverse_radical_inverse((uint64_t)pmy, 3, s.base_exp1);
    offset += d0 * (uint64_t)(s.stride / s.base_scale0) * (uint64_t)s.mult_inv0;
    offset += d1 * (uint64_t)(s.stride / s.base_scale1) * (uint64_t)s.mult_inv1;
    offset %= (uint64_t)s.stride;
    return offset;
}

// Exact division of a < 2^25 by a prime d < 2^13 through the precomputed m = ceil(2^38 / d):
// floor(a * m / 2^38) == floor(a / d) because a * (m * d - 2^38) < 2^25 * 2^13 (Granlund-Montgomery).
// One 64-bit multiply instead of the ~25-instruction 32-bit division sequence per digit.
constexpr int kMagicShift = 38;
constexpr uint32_t kMagicLimit = 1u << 25;

// The digit loops run on 32-bit operands whenever the index fits (always, at the configs: the
// largest index is 31 104 * 1024 + 31 103 < 2^25); the reversed-digit accumulator stays 64-bit.
GNX_D float radical_inverse_base(uint64_t a64, uint32_t base) {
    const float invBase = 1.0f / (float)base;
    uint64_t reversed = 0;
    float invBaseN = 1;
    if (a64 <= 0xffffffffull) {
        uint32_t a = (uint32_t)a64;
        while (a) {
            uint32_t next = a / base;
            uint32_t digit = a - next * base;
            reversed = reversed * base + digit;
            invBaseN *= invBase;
            a = next;
        }
    } else {
        while (a64) {
            uint64_t next = a64 / base;
            uint64_t digit = a64 - next * base;
            reversed = reversed * base + digit;
            invBaseN *= invBase;
            a64 = next;
        }
    }
    return fminf((float)reversed * invBaseN, kOneMinusEpsilon);
}

// RadicalInverseSpecialized<3>: the divisions are by a compile-time constant (multiply-high)
GNX_D float radical_inverse_base3(uint64_t a64) {
    if (a64 > 0xffffffffull) return radical_inverse_base(a64, 3u);
    const float invBase = 1.0f / 3.0f;
    uint64_t reversed = 0;
    float invBaseN = 1;
    uint32_t a = (uint32_t)a64;
    while (a) {
        uint32_t next = a / 3u;
        uint32_t digit = a - next * 3u;
        reversed = reversed * 3u + digit;
        invBaseN *= invBase;
        a = next;
    }
    return fminf((float)reversed * invBaseN, kOneMinusEpsilon);
}

GNX_D float scrambled_radical_inverse_base(uint64_t a64, uint32_t base, const uint16_t *perm, uint64_t magic) {
    const float invBase = 1.0f / (float)base;
    uint64_t reversed = 0;
    float invBaseN = 1;
    if (a64 < kMagicLimit) {
        uint32_t a = (uint32_t)a64;
        while (a) {
            uint32_t next = (uint32_t)(((uint64_t)a * magic) >> kMagicShift);
            uint32_t digit = a - next * base;
            reversed = reversed * base + ldg(perm + digit);
            invBaseN *= invBase;
            a = next;
        }
    } else {
        while (a64) {
            uint64_t next = a64 / base;
            uint64_t digit = a64 - next * base;
            reversed = reversed * base + ldg(perm + digit);
            invBaseN *= invBase;
            a64 = next;
        }
    }
    float p0 = (float)ldg(perm);
    return fminf(invBaseN * ((float)reversed + invBase * p0 / (1 - invBase)), kOneMinusEpsilon);
}

// RadicalInverse(0, a): ReverseBits64(a) * 2^-64 in double, narrowed (LowDiscrepancy.cpp:396-405)
GNX_D float radical_inverse_base2(uint64_t a) {
    // a < 2^32: the reversed bits fill the top word only, so the double holds them exactly and the one rounding
    // that matters is the narrowing to float — the same rounding as uint32 -> float (no fp64 on the fast path)
    if (a <= 0xffffffffull) return (float)brev32((uint32_t)a) * 2.3283064365386963e-10f;
    uint64_t r = brev64(a);
    return (float)((double)r * 5.4210108624275222e-20);
}

// Unscrambled RadicalInverse(baseIndex, a) for the first few bases (used by the spatial light
// distribution, core/LightDistribution.cpp:230-236).
GNX_D float radical_inverse(const DevSampler &s, int baseIndex, uint64_t a) {
    if (baseIndex == 0) return radical_inverse_base2(a);
    return radical_inverse_base(a, (uint32_t)ldg(s.primes + baseIndex));
}

// ---- Sobol' (samplers/LowDiscrepancy.h:194-252) ----------------------------------------------------------------
constexpr int kSobolMatrixSize = 52;   // samplers/SobolMatrices.h:14
constexpr int kSobolConstDims = 192;   // dimensions whose matrices live in __constant__ memory (39 936 B); a path of
                                       // maxDepth 5 uses 5 + 8 * 6 = 53 of them.  All lanes of a warp walk the bits of
                                       // their indices in step, so every matrix read is one broadcast.
#if defined(__CUDACC__)
static __constant__ uint32_t c_sobol32[kSobolConstDims * kSobolMatrixSize];
}  // namespace gnx
void gnx_register_sobol_uploader(void (*fn)(const uint32_t *));  // gnx_render.cu
namespace gnx {
namespace {
struct SobolConstRegistrar {
    SobolConstRegistrar() {
        gnx_register_sobol_uploader([](const uint32_t *head) { cudaMemcpyToSymbol(c_sobol32, head, sizeof(c_sobol32)); });
    }
};
static SobolConstRegistrar g_sobol_const_registrar;  // one per translation unit: each has its own copy of the symbol
}  // namespace
#endif
// SobolSampleFloat(a, dimension, scramble = 0)
GNX_D float sobol_sample_float(const DevSampler &s, uint64_t a, int dimension) {
    if (dimension >= s.sobol_dims) return 0.f;  // (the reference reads past its tables here)
    uint32_t v = 0;
#if defined(__CUDA_ARCH__)
    if (dimension < kSobolConstDims) {
        const uint32_t *M = c_sobol32 + dimension * kSobolMatrixSize;
        for (int i = 0; a != 0; a >>= 1, i++) if (a & 1) v ^= M[i];
    } else
#endif
    {
        const uint32_t *M = s.sobol32 + (size_t)dimension * kSobolMatrixSize;
        for (int i = 0; a != 0; a >>= 1, i++) if (a & 1) v ^= ldg(M + i);
    }
    return fminf(v * 2.3283064365386963e-10f, kOneMinusEpsilon);
}
// SobolIntervalToIndex(m, frame, p)
GNX_D uint64_t sobol_interval_to_index(const DevSampler &s, uint64_t frame, int px, int py) {
    const uint32_t m = (uint32_t)s.sobol_log2res;
    if (m == 0) return 0;
    const uint32_t m2 = m << 1;
    uint64_t index = frame << m2;
    uint64_t delta = 0;
    for (int c = 0; frame; frame >>= 1, ++c)
        if (frame & 1) delta ^= ldg(s.sobol_vdc + c);
    uint64_t b = ((((uint64_t)((uint32_t)px)) << m) | ((uint32_t)py)) ^ delta;
    for (int c = 0; b; b >>= 1, ++c)
        if (b & 1) index ^= ldg(s.sobol_vdc_inv + c);
    return index;
}

// GlobalSampler::GetIndexForSample(sampleNum) for pixel (px, py): HaltonSampler (samplers/HaltonSampler.cpp:63-82) or the
// Sobol' sampler (SobolIntervalToIndex over the pixel grid of resolution 2^m).
GNX_D uint64_t sampler_index(const DevSampler &s, int px, int py, uint64_t sampleNum);

// HaltonSampler::SampleDimension, samplers/HaltonSampler.cpp:85-94 — or, for GNX_SAMPLER_SOBOL, SobolSample(index, dim)
// WITHOUT the pixel remap of the first two dimensions (see sampler_film_dimensions)
GNX_D float halton_sample_dimension(const DevSampler &s, uint64_t index, int dim) {
    if (s.type == GNX_SAMPLER_SOBOL) return sobol_sample_float(s, index, dim);
    if (s.at_center && (dim == 0 || dim == 1)) return 0.5f;
    if (dim == 0) return radical_inverse_base2(index >> s.base_exp0);
    if (dim == 1)
        return radical_inverse_base3(index <= 0xffffffffull ? (uint64_t)((uint32_t)index / (uint32_t)s.base_scale1)
                                                           : index / (uint64_t)s.base_scale1);
    // ScrambledRadicalInverse returns 0 for base indices it has no case for (>= 1024); the
    // reference's PrimeSums read at dim >= 1000 is out of range, we return 0.
    if (dim >= s.n_primes) return 0.f;
    uint4 rec = ldg(s.dims + dim);  // {prime, offset of its permutation, magic lo, magic hi}
    return scrambled_radical_inverse_base(index, rec.x, s.perms + rec.y, ((uint64_t)rec.w << 32) | rec.z);
}

GNX_D uint64_t sampler_index(const DevSampler &s, int px, int py, uint64_t sampleNum) {
    if (s.type == GNX_SAMPLER_SOBOL) return sobol_interval_to_index(s, sampleNum, px, py);
    return halton_pixel_offset(s, px, py) + sampleNum * (uint64_t)s.stride;
}
// Dimensions 0 and 1 of a camera sample (the film offset inside the pixel).  The Sobol' sampler stretches them over the
// whole image and takes the part inside the current pixel: s * resolution - pixel, clamped to [0, 1).
GNX_D void sampler_film_dimensions(const DevSampler &s, uint64_t index, int px, int py, float *u0, float *u1) {
    *u0 = halton_sample_dimension(s, index, 0);
    *u1 = halton_sample_dimension(s, index, 1);
    if (s.type == GNX_SAMPLER_SOBOL) {
        float a = *u0 * s.sobol_res + 0.f, b = *u1 * s.sobol_res + 0.f;  // + sampleBounds.pMin (the image starts at 0)
        a -= (float)px; b -= (float)py;
        *u0 = a < 0.f ? 0.f : (a > kOneMinusEpsilon ? kOneMinusEpsilon : a);
        *u1 = b < 0.f ? 0.f : (b > kOneMinusEpsilon ? kOneMinusEpsilon : b);
    }
}

// PCG32 (core/RNG.h:30-110), used as the per-pixel stream when the sampler is not Halton.
struct Pcg32 {
    uint64_t state, inc;
    GNX_HD void set_sequence(uint64_t seq) {
        state = 0u;
        inc = (seq << 1u) | 1u;
        next_u32();
        state += 0x853c49e6748fea9bULL;
        next_u32();
    }
    GNX_HD uint32_t next_u32() {
        uint64_t old = state;
        state = old * 0x5851f42d4c957f2dULL + inc;
        uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
    }
    GNX_HD float uniform_float() { return fminf(kOneMinusEpsilon, (float)(next_u32() * 2.3283064365386963e-10f)); }
};

// Sampler view of one path.
//   Halton: GlobalSampler::Get1D/Get2D (core/Sampler.cpp:162-179).  arrayStartDim is 5 and no sample arrays
//   are ever requested by Path/VolPath, so the array-skip branch is inert.
//   PCG32:  one stream per camera sample, sequence id = (W * py + px) << 20 | sampleNumber — the
//   convention of the harness's Sampler subclass (oracle/ref_harness.cpp, PcgStreamSampler), which the
//   unbounded delta-tracking loops of GridDensityMedium need (the reference's Halton tables stop at
//   dimension 1000 and read out of range beyond).
struct PathSampler {
    const DevSampler &s;
    uint64_t index;
    int dim;
    bool pcg;
    Pcg32 rng;
    GNX_D PathSampler(const DevSampler &smp, uint64_t idx, int d) : s(smp), index(idx), dim(d), pcg(false) {}
    GNX_D static PathSampler stream(const DevSampler &smp, uint64_t sequence) {
        PathSampler p(smp, 0, 0);
        p.pcg = true;
        p.rng.set_sequence(sequence);
        return p;
    }
    // takes over another view's position in the stream (the table reference stays: both views are of the same sampler)
    GNX_D void take(const PathSampler &o) { index = o.index; dim = o.dim; pcg = o.pcg; rng = o.rng; }
    GNX_D float get1d() {
        if (pcg) { ++dim; return rng.uniform_float(); }
        return halton_sample_dimension(s, index, dim++);
    }
    // Sampler::GetCameraSample's film sample: dimensions 0 and 1 (remapped into the pixel by the Sobol' sampler)
    GNX_D void get_film(int px, int py, float *a, float *b) {
        if (pcg) { get2d(a, b); return; }
        sampler_film_dimensions(s, index, px, py, a, b);
        dim += 2;
    }
    GNX_D void get2d(float *a, float *b) {
        if (pcg) { *a = rng.uniform_float(); *b = rng.uniform_float(); dim += 2; return; }
        *a = halton_sample_dimension(s, index, dim);
        *b = halton_sample_dimension(s, index, dim + 1);
        dim += 2;
    }
};

}  // namespace gnx
