// gnx_kernels.cuh — the wavefront kernels (sm_100a).  Per-path logic is in gnx_path.cuh; a kernel
// walks its input queue with a block-uniform grid-stride loop so that every lane of a warp reaches
// the warp-aggregated queue pushes (__ballot_sync + one atomicAdd per warp and queue).
//
//   k_raygen  -> k_extend -> k_shade<type> ... -> k_shadow x2, k_probe -> (next bounce) ... -> k_accumulate -> k_film
//
// Grid sizes are multiples of the SM count (148 on B200); the traversal kernels keep their stacks
// in shared memory (12 KB per 128-thread block).
#pragma once
#include "gnx_path.cuh"

namespace gnx {

constexpr int kBlock = 128;
constexpr unsigned kFull = 0xffffffffu;

// Lanes with `pred` receive consecutive indices of *counter (one atomic per warp); others get -1.
__device__ __forceinline__ int warp_push(int *counter, bool pred) {
    unsigned m = __ballot_sync(kFull, pred);
    if (m == 0) return -1;
    int lane = threadIdx.x & 31;
    int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(kFull, base, leader);
    return pred ? base + __popc(m & ((1u << lane) - 1)) : -1;
}

__device__ __forceinline__ void flush_stats(DevStats *st, int kind, unsigned nodes, unsigned tris, unsigned rays) {
    for (int o = 16; o > 0; o >>= 1) {
        nodes += __shfl_down_sync(kFull, nodes, o);
        tris += __shfl_down_sync(kFull, tris, o);
        rays += __shfl_down_sync(kFull, rays, o);
    }
    if ((threadIdx.x & 31) == 0 && rays) {
        atomicAdd(&st->nodes[kind], (unsigned long long)nodes);
        atomicAdd(&st->tris[kind], (unsigned long long)tris);
        atomicAdd(&st->rays[kind], (unsigned long long)rays);
    }
}

__global__ void __launch_bounds__(256) k_raygen(const DeviceScene sc, PathState ps, Queues q, RenderConsts rc, DevStats *st) {
    const int n = rc.npix * rc.batch_spp;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += gridDim.x * blockDim.x) {
        raygen_slot(sc, ps, rc, slot);
        q.extend_q[0][slot] = slot;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        q.counts[kCntExtend0] = n;
        atomicAdd(&st->paths, (unsigned long long)n);
    }
}

// Zeroes every queue counter except the extend queue that is about to be consumed.
__global__ void k_reset_counts(int *counts, int outExtend) {
    int i = threadIdx.x;
    if (i < kNumCounters && i != (1 - outExtend)) counts[i] = 0;
}

__global__ void __launch_bounds__(kBlock) k_extend(const DeviceScene sc, PathState ps, Queues q, RenderConsts rc,
                                                    int inQ, DevStats *st) {
    __shared__ int s_stack[kSmemStack * kBlock];
    int *stack = s_stack + threadIdx.x;
    const int n = q.counts[inQ];
    const int *inList = q.extend_q[inQ];
    TraversalCounters cnt{0, 0};
    unsigned rays = 0;
    const int stride = gridDim.x * blockDim.x;
    for (int base = blockIdx.x * blockDim.x; base < n; base += stride) {
        const int i = base + threadIdx.x;
        int type = -1, slot = 0;
        if (i < n) {
            slot = inList[i];
            ++rays;
            type = extend_slot(sc, ps, rc, slot, stack, kBlock, cnt);
        }
#pragma unroll
        for (int t = 0; t < kNumShadeTypes; ++t) {
            int idx = warp_push(&q.counts[kCntShade0 + t], type == t);
            if (idx >= 0) q.shade_q[(size_t)t * q.capacity + idx] = slot;
        }
    }
    flush_stats(st, 0, cnt.nodes, cnt.tris, rays);
}

__global__ void __launch_bounds__(kBlock) k_shade_null(const DeviceScene sc, PathState ps, Queues q, RenderConsts rc,
                                                        int outQ) {
    const int n = q.counts[kCntShade0 + (kNumShadeTypes - 1)];
    const int *list = q.shade_q + (size_t)(kNumShadeTypes - 1) * q.capacity;
    const int stride = gridDim.x * blockDim.x;
    for (int base = blockIdx.x * blockDim.x; base < n; base += stride) {
        const int i = base + threadIdx.x;
        int slot = 0;
        bool alive = false;
        if (i < n) {
            slot = list[i];
            alive = shade_null_slot(sc, ps, rc, slot);
        }
        int idx = warp_push(&q.counts[outQ], alive);
        if (idx >= 0) q.extend_q[outQ][idx] = slot;
    }
}

template <int MAXL>
__global__ void __launch_bounds__(kBlock) k_shade(const DeviceScene sc, PathState ps, Queues q, RenderConsts rc,
                                                   int type, int outQ) {
    const int n = q.counts[kCntShade0 + type];
    const int *list = q.shade_q + (size_t)type * q.capacity;
    const int stride = gridDim.x * blockDim.x;
    for (int base = blockIdx.x * blockDim.x; base < n; base += stride) {
        const int i = base + threadIdx.x;
        int slot = 0;
        ShadeOut out;
        out.alive = out.haveShadowA = out.haveShadowB = out.haveProbe = false;
        if (i < n) {
            slot = list[i];
            shade_slot<MAXL>(sc, ps, rc, slot, out);
        }
        int idx = warp_push(&q.counts[kCntShadow], out.haveShadowA);
        if (idx >= 0) q.shadow_q[idx] = out.shA;
        idx = warp_push(&q.counts[kCntShadow + 1], out.haveShadowB);
        if (idx >= 0) q.shadow_q[(size_t)q.capacity + idx] = out.shB;
        idx = warp_push(&q.counts[kCntProbe], out.haveProbe);
        if (idx >= 0) q.probe_q[idx] = out.pr;
        idx = warp_push(&q.counts[outQ], out.alive);
        if (idx >= 0) q.extend_q[outQ][idx] = slot;
    }
}

// which == 0: VisibilityTester rays; which == 1: the BSDF-sampled MIS ray toward the environment.
// Two launches (not one mixed queue) keep the two additions into a path's L in a fixed order.
__global__ void __launch_bounds__(kBlock) k_shadow(const DeviceScene sc, PathState ps, Queues q, int which, DevStats *st) {
    __shared__ int s_stack[kSmemStack * kBlock];
    int *stack = s_stack + threadIdx.x;
    const int n = q.counts[kCntShadow + which];
    const ShadowItem *items = q.shadow_q + (size_t)which * q.capacity;
    TraversalCounters cnt{0, 0};
    unsigned rays = 0;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        ++rays;
        shadow_item(sc, ps, items + i, stack, kBlock, cnt);
    }
    flush_stats(st, which == 0 ? 1 : 2, cnt.nodes, cnt.tris, rays);
}

__global__ void __launch_bounds__(kBlock) k_probe(const DeviceScene sc, PathState ps, Queues q, DevStats *st) {
    __shared__ int s_stack[kSmemStack * kBlock];
    int *stack = s_stack + threadIdx.x;
    const int n = q.counts[kCntProbe];
    TraversalCounters cnt{0, 0};
    unsigned rays = 0;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        ++rays;
        probe_item(sc, ps, q.probe_q + i, stack, kBlock, cnt);
    }
    flush_stats(st, 2, cnt.nodes, cnt.tris, rays);
}

// colObj += Li(...) over the samples of the pixel, in sample order (core/Integrator.cpp:274-291)
__global__ void k_accumulate(PathState ps, float4 *accum, RenderConsts rc) {
    for (int pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < rc.npix; pixel += gridDim.x * blockDim.x) {
        float4 a = accum[pixel];
        for (int s = 0; s < rc.batch_spp; ++s) {
            const float4 L = ps.L[(size_t)s * rc.npix + pixel];
            a.x += L.x; a.y += L.y; a.z += L.z;
        }
        accum[pixel] = a;
    }
}

// colObj / samplesPerPixel, alpha 1 (core/Integrator.cpp:293,307-310)
__global__ void k_film(const float4 *accum, float4 *rgba, int npix, float spp) {
    for (int pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < npix; pixel += gridDim.x * blockDim.x) {
        const float4 a = accum[pixel];
        rgba[pixel] = make_float4(a.x / spp, a.y / spp, a.z / spp, 1.f);
    }
}

__global__ void __launch_bounds__(kBlock) k_primary_hits(const DeviceScene sc, int width, int height, int sample, int *out) {
    __shared__ int s_stack[kSmemStack * kBlock];
    int *stack = s_stack + threadIdx.x;
    const int npix = width * height;
    for (int pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < npix; pixel += gridDim.x * blockDim.x)
        out[pixel] = primary_hit_id(sc, pixel % width, pixel / width, sample, stack, kBlock);
}

__global__ void k_sample_dims(const DeviceScene sc, int n, const long long *index, const int *dim, float *out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = halton_sample_dimension(sc.smp, (uint64_t)index[i], dim[i]);
}

__global__ void k_build_spatial(const DeviceScene sc, float *func, float *cdf, float *fint) {
    const int nv = sc.ld.nvox[0] * sc.ld.nvox[1] * sc.ld.nvox[2];
    for (int vox = blockIdx.x * blockDim.x + threadIdx.x; vox < nv; vox += gridDim.x * blockDim.x)
        build_spatial_voxel(sc, vox, func, cdf, fint);
}

// FrameBuffer::update_f_u_c's tonemap on the first pass (ui/FrameBuffer.h:141-147)
__global__ void k_tonemap(const float4 *rgba, uchar4 *out, int npix) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        const float4 c = rgba[i];
        const float exposure = 0.75f;
        float r = 1.0f - expf(-c.x * 1.0f / (1 - exposure));
        float g = 1.0f - expf(-c.y * 1.0f / (1 - exposure));
        float b = 1.0f - expf(-c.z * 1.0f / (1 - exposure));
        out[i] = make_uchar4((unsigned char)(r * 255), (unsigned char)(g * 255), (unsigned char)(b * 255), 255);
    }
}

}  // namespace gnx
