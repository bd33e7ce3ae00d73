// SobolSampler — the Sobol' GlobalSampler the reference has the ingredients for but no class of.
//
// The reference ships the generator matrices (samplers/SobolMatrices.h:12-17) and the helpers SobolIntervalToIndex /
// SobolSample (samplers/LowDiscrepancy.h:194-252) without a Sampler built on them; BASELINE.json's north_star names
// "Halton/Sobol" ray generation.  This header is that class, written against the reference's own GlobalSampler
// interface (core/Sampler.h:75-94) and helpers only, in the shape of the renderer the reference descends from
// (pbrt-v3's SobolSampler): samples per pixel rounded up to a power of two, one global sequence over the image's
// bounding power-of-two square, the first two dimensions remapped into the current pixel.  It runs on the CPU inside
// the reference's integrators (that is the oracle for it) and gnx::FlattenScene recognises it and hands its state —
// and the reference's matrices, by pointer — to the device sampler (GNX_SAMPLER_SOBOL, csrc/gnx_sampler.cuh).
#ifndef GNX_SOBOL_SAMPLER_H
#define GNX_SOBOL_SAMPLER_H

#include <algorithm>
#include <memory>

#include "core/Sampler.h"
#include "samplers/LowDiscrepancy.h"
#include "samplers/SobolMatrices.h"

namespace gnx {

class SobolSampler : public pbr::GlobalSampler {
  public:
    SobolSampler(int64_t samplesPerPixel, const pbr::Bounds2i &sampleBounds)
        : pbr::GlobalSampler(pbr::RoundUpPow2((int32_t)samplesPerPixel)), sampleBounds(sampleBounds) {
        const pbr::Vector2i diag = sampleBounds.Diagonal();
        resolution = pbr::RoundUpPow2(std::max(diag.x, diag.y));
        log2Resolution = pbr::Log2Int(resolution);
    }
    int64_t GetIndexForSample(int64_t sampleNum) const override {
        return (int64_t)pbr::SobolIntervalToIndex((uint32_t)log2Resolution, (uint64_t)sampleNum,
                                                  pbr::Point2i(currentPixel - sampleBounds.pMin));
    }
    pbr::Float SampleDimension(int64_t index, int dim) const override {
        pbr::Float s = pbr::SobolSample(index, dim);
        if (dim == 0 || dim == 1) {  // the film dimensions: stretch over the image, keep the part inside the pixel
            s = s * resolution + sampleBounds.pMin[dim];
            s = pbr::Clamp(s - currentPixel[dim], (pbr::Float)0, pbr::OneMinusEpsilon);
        }
        return s;
    }
    std::unique_ptr<pbr::Sampler> Clone(int seed) override { return std::unique_ptr<pbr::Sampler>(new SobolSampler(*this)); }

    const pbr::Bounds2i sampleBounds;
    int resolution, log2Resolution;
};

}  // namespace gnx
#endif
