// gnx_film.cuh — optional Gaussian reconstruction film (north_star stage 5).
//
// The reference owns a GaussianFilter class (filters/GaussianFilter.h:12-33, filters/GaussianFilter.cpp:9-12) but its
// Render() averages a pixel's samples with a box (core/Integrator.cpp:274-293); there is no Film that uses the filter.
// This film gives the filter the semantics of the renderer the reference descends from (pbrt-v3 Film::AddSample /
// WriteImage): a sample at pFilm contributes L * f(x - pd.x, y - pd.y) to every pixel (x, y) with
// ceil(pd - radius) <= (x, y) <= floor(pd + radius), pd = pFilm - (0.5, 0.5), and a pixel is
// max(0, sum(L f) / sum(f)).  The filter is evaluated exactly (no 16 x 16 lookup table) with the reference's own
// formula, max(0, exp(-alpha d^2) - exp(-alpha r^2)) per axis.
//
// On the GPU the splat becomes a GATHER: one thread per output pixel walks the samples of the pixels within reach in a
// fixed order, so the image is deterministic (no float atomics) like the box film.  All functions are
// __host__ __device__; tests/emul runs the same code on the CPU.
#pragma once
#include "gnx_path.cuh"

namespace gnx {

struct FilmFilter {
    float radius, alpha, expv;  // expv = exp(-alpha * radius^2)  (GaussianFilter ctor, filters/GaussianFilter.h:15-19)
    int reach;                  // source pixels up to `reach` away can touch a pixel: floor(radius + 0.5)
};

GNX_D float gaussian_1d(const FilmFilter &f, float d) {  // GaussianFilter::Gaussian, filters/GaussianFilter.h:28-31
    return fmaxf(0.f, expf(-f.alpha * d * d) - f.expv);
}

// Film offsets (CameraSample::pFilm - pixel) of camera sample `sample` of pixel (px, py): the first two values of the
// sample vector (Sampler::GetCameraSample, core/Sampler.cpp:14-20), for either sampler.
GNX_D void film_sample_offset(const DeviceScene &sc, int width, int px, int py, int sample, float *u0, float *u1) {
    if (sc.smp.type == GNX_SAMPLER_PCG32) {
        PathSampler smp = PathSampler::stream(sc.smp, ((uint64_t)(width * py + px) << 20) | (uint64_t)sample);
        smp.get2d(u0, u1);
    } else {
        uint64_t hidx = sampler_index(sc.smp, px, py, (uint64_t)sample);
        sampler_film_dimensions(sc.smp, hidx, px, py, u0, u1);
    }
}

// (sum of L * weight, sum of weight) over the samples of one batch that touch pixel (x, y).  L[slot].xyz is the
// sample's radiance, off[slot].xy its film offset; slot = pixel * batch_spp + sample (slot_to_sample).
GNX_D float4 gaussian_gather(const float4 *L, const float4 *off, const RenderConsts &rc, const FilmFilter &f, int x, int y) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    const int y0 = y - f.reach < 0 ? 0 : y - f.reach, y1 = y + f.reach > rc.height - 1 ? rc.height - 1 : y + f.reach;
    const int x0 = x - f.reach < 0 ? 0 : x - f.reach, x1 = x + f.reach > rc.width - 1 ? rc.width - 1 : x + f.reach;
    const float fx = (float)x, fy = (float)y;
    for (int sy = y0; sy <= y1; ++sy)
        for (int sx = x0; sx <= x1; ++sx) {
            const size_t base = ((size_t)sy * rc.width + sx) * rc.batch_spp;
            for (int s = 0; s < rc.batch_spp; ++s) {
                const float4 o = off[base + s];
                const float pdx = ((float)sx + o.x) - 0.5f, pdy = ((float)sy + o.y) - 0.5f;
                if (fx < ceilf(pdx - f.radius) || fx > floorf(pdx + f.radius) || fy < ceilf(pdy - f.radius) ||
                    fy > floorf(pdy + f.radius))
                    continue;
                const float w = gaussian_1d(f, fx - pdx) * gaussian_1d(f, fy - pdy);
                const float4 l = L[base + s];
                a.x += l.x * w; a.y += l.y * w; a.z += l.z * w; a.w += w;
            }
        }
    return a;
}

// Film::WriteImage's normalisation: rgb / filterWeightSum, clamped at 0.
GNX_D float4 gaussian_resolve(float4 a) {
    if (a.w == 0.f) return make_float4(0.f, 0.f, 0.f, 1.f);
    const float inv = 1.f / a.w;
    return make_float4(fmaxf(0.f, a.x * inv), fmaxf(0.f, a.y * inv), fmaxf(0.f, a.z * inv), 1.f);
}

}  // namespace gnx
