// gnx_kernels.cuh — the wavefront kernels (sm_100a).  Per-path logic is in gnx_path.cuh; a kernel
// walks its input queue with a block-uniform grid-stride loop so that every lane of a warp reaches
// the warp-aggregated queue pushes (__ballot_sync + one atomicAdd per warp and queue).
//
//   PathIntegrator:  k_trace<3> (ray-gen + first extension) -> [k_qs_* : dense shade queues back into slot order] -> k_shade<type> ...
//                    -> k_trace<2> (area-light MIS probes) -> { k_trace<0> (extension rays, reference-order two-child tree)  ||
//                    k_anyhit8<1> (the previous bounce's shadow / environment-MIS rays, compressed 8-wide tree, second stream) }
//                    -> k_shade ... -> (next bounce) ... -> k_anyhit8<0> (last bounce's any-hit rays) -> k_accumulate -> k_film
//   VolPath:         k_vp_logic<extend> -> rounds of { k_vp_track, k_vp_logic<vertex | shadow | MIS | extend> } -> k_accumulate -> k_film
//   Whitted:         k_trace<3> -> k_whitted_vertex -> k_anyhit8<0> -> k_whitted_sum -> k_recursive<0> over the listed samples -> ...
//   DirectLighting:  k_recursive<1 | 2> (one camera sample per lane)
//
// Grids are sized from the occupancy query: SM count (148 on B200) x resident blocks per SM, so a
// persistent kernel never has a partial second wave.  The traversal kernel keeps its stacks in shared
// memory (24 KB per 128-thread block).
#pragma once
#include "gnx_whitted.cuh"
#include "gnx_volwave.cuh"
#include "gnx_film.cuh"

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
    if ((threadIdx.x & 31) == 0 && (rays | nodes | tris)) {
        atomicAdd(&st->nodes[kind], (unsigned long long)nodes);
        atomicAdd(&st->tris[kind], (unsigned long long)tris);
        atomicAdd(&st->rays[kind], (unsigned long long)rays);
    }
}

// Zeroes every queue counter except the extend queue that is about to be consumed.
#ifndef GNX_KERNELS_TEMPLATES_ONLY
__global__ void k_reset_counts(int *counts, int outExtend) {
    int i = threadIdx.x;
    if (i < kNumCounters && i != (1 - outExtend)) counts[i] = 0;
}
#endif

// After the reference-order launch over the rays a wide closest-hit launch flagged: their list (extend queue qi) and the
// cursor that launch fetched through are free again.
#ifndef GNX_KERNELS_TEMPLATES_ONLY
__global__ void k_retrace_reset(int *counts, int qi) {
    if (threadIdx.x == 0) { counts[qi] = 0; counts[kCntFetch + 3] = 0; }
}
#endif

// ---- a queue back into slot order ---------------------------------------------------------------------------
// The traversal kernel hands out rays dynamically, so the slots it pushes to the shade queues arrive in completion order:
// neighbouring lanes of the shade kernel then work on unrelated paths (every path-state access its own 32-byte sector),
// and the shadow / extension rays they emit start at unrelated points.  A queue holds distinct slots of [0, nSlots), so
// sorting it is a bitmap: mark (k_qs_mark), count the bits per block of 256 words (k_qs_count), scan the block counts
// (k_qs_scan, one block), write the set bits back in order and clear them (k_qs_emit).  4 MB of bitmap for 32 M slots.
#ifndef GNX_KERNELS_TEMPLATES_ONLY
constexpr int kQsBlock = 256;  // threads = bitmap words per block
// camera rays that hit a surface (the shade queues after the first traversal launch), for the host's decision below
__global__ void k_qs_probe(const int *counts, int *out) {
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < kNumShadeTypes; ++k) t += counts[kCntShade0 + k];
        *out = t;
    }
}
// (a sparse queue is left alone — every kernel returns at once: walking the bitmap of ALL slots then costs more than the
// order gains; measured on C2 / C3, where a tenth of the camera rays hit anything: -1.3 % / -2.7 % when sorted regardless,
// against +5 % on the UI's closed scene)
constexpr int kQsMinDensity = 4;  // sort when count * kQsMinDensity >= slots
__global__ void k_qs_mark(const int *list, const int *count, unsigned *bits, int nSlots) {
    const int n = *count;
    if ((long long)n * kQsMinDensity < nSlots) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int e = list[i];
        atomicOr(&bits[e >> 5], 1u << (e & 31));
    }
}
__global__ void k_qs_count(const unsigned *bits, int nWords, int *blockCounts, const int *count, int nSlots) {
    __shared__ int s_sum[kQsBlock / 32];
    if ((long long)*count * kQsMinDensity < nSlots) return;
    const int w = blockIdx.x * kQsBlock + threadIdx.x;
    int c = w < nWords ? __popc(bits[w]) : 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(kFull, c, o);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < kQsBlock / 32; ++k) t += s_sum[k];
        blockCounts[blockIdx.x] = t;
    }
}
// exclusive scan of the block counts in place (one block of 1024 threads, any number of entries)
__global__ void k_qs_scan(int *blockCounts, int nBlocks, const int *count, int nSlots) {
    __shared__ int s_part[1024];
    if ((long long)*count * kQsMinDensity < nSlots) return;
    const int per = (nBlocks + 1023) / 1024, lo = threadIdx.x * per, hi = min(lo + per, nBlocks);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += blockCounts[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over the 1024 partial sums
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = s_part[threadIdx.x] - sum;
    for (int i = lo; i < hi; ++i) { const int c = blockCounts[i]; blockCounts[i] = run; run += c; }
}
__global__ void k_qs_emit(unsigned *bits, int nWords, const int *blockOffsets, int *list, const int *count, int nSlots) {
    __shared__ int s_warp[kQsBlock / 32];
    if ((long long)*count * kQsMinDensity < nSlots) return;
    const int w = blockIdx.x * kQsBlock + threadIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned word = w < nWords ? bits[w] : 0u;
    const int c = __popc(word);
    int incl = c;
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    int base = blockOffsets[blockIdx.x];
    for (int k = 0; k < wid; ++k) base += s_warp[k];
    int pos = base + incl - c;
    if (word) {
        bits[w] = 0u;
        while (word) {
            const int b = __ffs(word) - 1;
            word &= word - 1;
            list[pos++] = (w << 5) + b;
        }
    }
}
#endif

// Zeroes the shadow / probe counters and the fetch cursors (between a mixed trace launch and the shade stage).
#ifndef GNX_KERNELS_TEMPLATES_ONLY
__global__ void k_reset_ray_counts(int *counts) {
    int i = threadIdx.x;
    if (i >= kCntShadow && i < kNumCounters) counts[i] = 0;
}
#endif
// Zeroes the counters of the queues the coming stages will fill, keeping the extend queue `in` AND the shadow queues
// (both are consumed by the mixed trace launch that follows); fetch cursors are zeroed.
#ifndef GNX_KERNELS_TEMPLATES_ONLY
__global__ void k_reset_counts_keep_rays(int *counts, int outExtend) {
    int i = threadIdx.x;
    if (i < kNumCounters && i != (1 - outExtend) && !(i >= kCntShadow && i < kCntShadow + 2)) counts[i] = 0;
}
#endif

// ---- the traversal kernel --------------------------------------------------------------------------------
// Persistent threads with dynamic ray fetch (Aila & Laine's while-while scheme): the grid is exactly the
// number of resident blocks; every lane owns at most one ray; lanes whose ray has finished pull the next
// unprocessed queue entry through a warp-aggregated atomic on a global cursor as soon as fewer than
// kRefetchBelow lanes of the warp are still traversing.  Inside, lanes first all descend interior nodes
// (two slab tests per 64-byte fetch), then all intersect their pending leaf, so a warp executes one
// kind of work at a time.  This replaced a static ray-per-thread mapping whose SIMD efficiency profiled
// at 5-11 active lanes of 32 (profiles/r01_extend_static.txt).
//   KIND 0: extension rays of the path slots in extend_q[arg]   -> hit record, env radiance, shade queues
//   KIND 1: shadow_q half `arg` (any-hit)                        -> L += contrib when unoccluded
//   KIND 2: probe_q (closest hit must be the light's triangle)   -> L += contrib
//   KIND 4: KIND 0 and KIND 1 (both halves) in one launch: the extension rays of bounce d+1 and the any-hit rays of
//           bounce d are independent work, and every launch of this persistent kernel ends in a tail of a few long rays
//   KIND 3: camera rays of the batch, generated in registers (ray-gen fused with the first extension;
//           path state is written only for rays that hit)      -> L = Le or hit record + shade queues
#ifndef GNX_REFETCH
#define GNX_REFETCH 24
#endif
constexpr int kRefetchBelow = GNX_REFETCH;
#ifndef GNX_REFETCH_PRIMARY
#define GNX_REFETCH_PRIMARY 1
#endif
// Camera rays refill only when the whole warp has drained: generating a camera ray (Halton digits, lens,
// ray transform) costs about as much as tracing it, so ray-gen must run with all 32 lanes (measured on C2:
// threshold 24 -> 39.4 ms, 12 -> 36.2, 4 -> 35.1, 1 -> 34.5).
constexpr int kRefetchBelowPrimary = GNX_REFETCH_PRIMARY;
#ifndef GNX_LEAF_BATCH
#define GNX_LEAF_BATCH 4
#endif
constexpr int kLeafBatch = GNX_LEAF_BATCH;  // parked lanes that trigger a joint triangle-test round (1..8 measure the same)

#ifndef GNX_TRACE_BLOCKS
#define GNX_TRACE_BLOCKS 8
#endif
//   WIDE (KIND 0 and 3): the closest hit through the compressed 8-wide tree (gnx_bvh8.cuh), front-to-back by octant slots.
//           A ray whose best hit has a rival within the tie band is not finished here: its slot (KIND 0) or sample index
//           (KIND 3) goes to the free extend queue `1 - arg` (KIND 3: queue 1), and the caller launches the reference-order
//           kernel over that list (KIND 0 with arg = the list; KIND 3 with arg = the list instead of -1 = "all samples").
constexpr int kPendRetrace = -3;
#ifndef GNX_TRACE8_BLOCKS
#define GNX_TRACE8_BLOCKS 8
#endif
template <int KIND, bool WIDE = false>
__global__ void __launch_bounds__(kBlock, WIDE ? GNX_TRACE8_BLOCKS : GNX_TRACE_BLOCKS) k_trace(const DeviceScene sc, PathState ps, Queues q, RenderConsts rc, int arg,
                                                      DevStats *st) {
    static_assert(!WIDE || KIND == 0 || KIND == 3, "the wide closest-hit traversal serves extension and camera rays");
    __shared__ int2 s_stack[kSmemStack * kBlock];
    int2 *stack = s_stack + threadIdx.x;
    const uint32_t sb = stack_shared_base(stack);
    const int lane = threadIdx.x & 31;
    const bool kExtend = KIND == 0 || KIND == 3 || KIND == 4;
    constexpr bool kMixed = KIND == 4;
    constexpr bool kPopRound = KIND == 1;
    // KIND 1: arg 0 = shadow A items, 1 = shadow B items, 2 = both halves in one launch (A first)
    // KIND 4: arg = extend queue; shadow items (A then B) follow the nE extension rays in the index space
    const int nShA = (KIND == 1 || kMixed) ? q.counts[kCntShadow] : 0, nShB = (KIND == 1 || kMixed) ? q.counts[kCntShadow + 1] : 0;
    const int qa = KIND == 0 ? (arg & 1) : arg;  // KIND 0: bit 1 of arg marks the retrace launch of flagged rays
    const int nE = (KIND == 0 || kMixed) ? q.counts[qa] : 0;
    const int *primList = (KIND == 3 && !WIDE && arg >= 0) ? q.extend_q[arg] : nullptr;  // retrace of flagged camera rays
    const int retraceQ = KIND == 3 ? 1 : 1 - (arg & 1);                                         // WIDE: where flagged rays go
    const int n = KIND == 3 ? (primList ? q.counts[arg] : rc.npix * rc.batch_spp)
                            : (KIND == 0 ? nE : (kMixed ? nE + nShA + nShB
                                                        : (KIND == 1 ? (arg == 0 ? nShA : (arg == 1 ? nShB : nShA + nShB)) : q.counts[kCntProbe])));
    // (the retrace launches of flagged rays fetch through the probe cursor: the extension cursor has been used by the wide launch)
    const bool retraceLaunch = !WIDE && ((KIND == 3 && arg >= 0) || (KIND == 0 && (arg & 2)));
    int *cursor = &q.counts[kCntFetch + (kExtend ? (retraceLaunch ? 3 : 0) : (KIND == 1 ? (arg == 1 ? 2 : 1) : 3))];
    const int *inList = (KIND == 0 || kMixed) ? q.extend_q[qa] : nullptr;
    const int firstB = kMixed ? nShA : (KIND == 1 ? (arg == 0 ? 0x7fffffff : (arg == 1 ? 0 : nShA)) : 0);  // shadow items from here on are B items
    bool shLane = false;  // KIND 4: the lane's ray is an any-hit ray
    unsigned raysSh = 0;
    auto shadowItem = [&](int i) { return i >= firstB ? q.shadow_q + (size_t)q.capacity + (i - firstB) : q.shadow_q + i; };
    unsigned raysB = 0;
    TraversalCounters cnt{0, 0};
    unsigned rays = 0;
    typename std::conditional<WIDE, Trav8, Trav>::type t;
    int2 spill[kSpillStack];
    t.spill = spill;
    t.cur = kRefNone;
    bool active = false, exhausted = false;
    int item = 0;            // path slot (KIND 0, 3) or queue index (KIND 1, 2) of the lane's ray
    int pendType = -1, pendSlot = 0;
    uint32_t hidx = 0;       // KIND 3 only
    V3 camD;                 // KIND 3 only
    bool needFinish = false;  // KIND 0 / 3: traversal over, result not consumed yet
    while (true) {
        if (kExtend) {
            // The rays that finished since the last visit are consumed here, together: the environment lookup
            // of escaped rays (atan2f / acosf / bilinear fetch) and the path-state writes of hits run with as
            // many lanes as were idle instead of one lane at a time inside the traversal loop.
            if (needFinish) {
                bool flagged = false;
                if constexpr (WIDE) flagged = t.tie;
                if (flagged) pendType = kPendRetrace;
                else if (KIND == 3) pendType = primary_finish(sc, ps, rc, item, hidx, camD, t);
                else pendType = extend_finish(sc, ps, rc, item, t);
                pendSlot = item;
                needFinish = false;
            }
            // shade-queue pushes, one atomic per warp and queue
            if (__any_sync(kFull, pendType != -1)) {
#pragma unroll
                for (int ty = 0; ty < kNumShadeTypes; ++ty) {
                    int idx = warp_push(&q.counts[kCntShade0 + ty], pendType == ty);
                    if (idx >= 0) q.shade_q[(size_t)ty * q.capacity + idx] = pendSlot;
                }
                if (q.miss_q) {  // scenes with a SkyBoxLight: escaped rays wait for k_escape
                    int idx = warp_push(&q.counts[kCntMiss], pendType == kPendEscape);
                    if (idx >= 0) q.miss_q[idx] = pendSlot;
                }
                if constexpr (WIDE) {      // flagged rays wait for the reference-order launch
                    int idx = warp_push(&q.counts[retraceQ], pendType == kPendRetrace);
                    if (idx >= 0) q.extend_q[retraceQ][idx] = pendSlot;
                }
                pendType = -1;
            }
        }
        if (!exhausted) {
            const unsigned idle = __ballot_sync(kFull, !active);
            if (idle) {
                const int leader = __ffs(idle) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(cursor, __popc(idle));
                base = __shfl_sync(kFull, base, leader);
                if (!active) {
                    const int i = base + __popc(idle & ((1u << lane) - 1));
                    if (i < n) {
                        bool valid = true;
                        if (KIND == 3) {
                            item = primList ? primList[i] : i;
                            int pixel, sample, px, py;
                            slot_to_sample(rc, item, &pixel, &sample);
                            valid = pixel_xy(rc, pixel, &px, &py);
                            if (valid) primary_begin(sc, px, py, sample, &hidx, &camD, t);
                            else ps.L[item] = make_float4(0.f, 0.f, 0.f, 0.f);  // pixel of an edge tile outside the image
                        } else if (kMixed) {
                            shLane = i >= nE;
                            if (shLane) { item = i - nE; shadow_begin(sc, shadowItem(item), t); ++raysSh; if (item >= firstB) ++raysB; }
                            else { item = inList[i]; extend_begin(sc, ps, item, t); }
                        } else if (KIND == 0) { item = inList[i]; extend_begin(sc, ps, item, t); }
                        else if (KIND == 1) { item = i; shadow_begin(sc, shadowItem(i), t); if (i >= firstB) ++raysB; }
                        else { item = i; probe_begin(sc, q.probe_q + i, t); }
                        if (valid) {
                            if constexpr (WIDE) trav8_init(sc, t);
                            active = true;
                            if (!retraceLaunch) ++rays;  // a retraced ray has been counted by the wide launch
                        }
                    }
                }
                exhausted = base + __popc(idle) >= n;
            }
        }
        if (!__any_sync(kFull, active)) break;
        // Warp-synchronous traversal rounds: all 32 lanes run the loop control together (full-mask votes
        // force reconvergence every round), lanes without work are predicated off.
        while (true) {
            // A lane that reaches a leaf parks until kLeafBatch lanes of the warp hold one (or nobody can
            // advance any more); then the warp runs the triangle test for all of them at once.  No speculation:
            // node and triangle visits are exactly those of the reference order.  (Measured alternatives, all
            // slower: deferring the leaf while the lane goes on with its next stack entry, +4..30 % with the batch
            // size; larger batches; see profiles/README.md.)
            // Pops requested by the previous round's interior steps and leaves.  Any-hit rays (entries never go stale):
            // predicated at the top of the round, no divergent block (shadow stage -3 %); closest-hit rays measure
            // 1-2 % better with the divergent pop below, whose re-validation loop skips culled entries at once.
            if constexpr (WIDE) { if (active && trav_needs_pop(t)) trav8_next(sc, t, stack, kBlock, sb); }
            else if (kPopRound) trav_pop_round<KIND == 1>(t, stack, kBlock, sb, active);
            bool parked = active && trav_is_leaf(t);
            if (active && !parked && trav_is_interior(t)) {
                if constexpr (WIDE) trav8_interior<true>(sc, t, stack, kBlock, cnt, sb);
                else trav_interior(sc, t, stack, kBlock, cnt, sb);
                parked = trav_is_leaf(t);
            }
            const unsigned parkedMask = __ballot_sync(kFull, parked);
            const unsigned advMask = __ballot_sync(kFull, active && !parked && !trav_done(t));
            if (parked && (__popc(parkedMask) >= kLeafBatch || advMask == 0)) {
                if constexpr (WIDE) { trav8_leaf_closest(sc, t, t.cur, cnt); t.cur = kRefPop; }
                else trav_leaf<KIND == 1>(sc, t, stack, kBlock, cnt);
                if (kMixed && shLane && t.hit) t.cur = kRefNone;  // any-hit ray in a mixed launch
            }
            if constexpr (!WIDE) { if (!kPopRound && active && trav_needs_pop(t)) trav_pop<KIND == 1>(t, stack, kBlock, sb); }
            if (active && trav_done(t)) {
                if (kExtend && !(kMixed && shLane)) needFinish = true;
                else if (KIND == 1 || kMixed) shadow_finish(ps, shadowItem(item), t, item >= firstB);
                else probe_finish(ps, q.probe_q + item, t);
                active = false;
            }
            const unsigned actMask = __ballot_sync(kFull, active);
            if (actMask == 0 || (!exhausted && __popc(actMask) < (KIND == 3 ? kRefetchBelowPrimary : kRefetchBelow))) break;
        }
    }
    if (KIND == 1) {
        // any-hit rays: shadow rays (A) are booked as kind 1, environment-MIS rays (B) as kind 2
        flush_stats(st, 1, cnt.nodes, cnt.tris, rays - raysB);
        flush_stats(st, 2, 0, 0, raysB);
    } else if (kMixed) {
        // nodes / triangles of both ray kinds are booked under the extension rays; the ray counts stay per kind
        flush_stats(st, 0, cnt.nodes, cnt.tris, rays - raysSh);
        flush_stats(st, 1, 0, 0, raysSh - raysB);
        flush_stats(st, 2, 0, 0, raysB);
        for (int o = 16; o > 0; o >>= 1) raysSh += __shfl_down_sync(kFull, raysSh, o);
        if ((threadIdx.x & 31) == 0 && raysSh) atomicAdd(&st->shadow_rays_in_extend_launches, (unsigned long long)raysSh);
    } else flush_stats(st, kExtend ? 0 : 2, cnt.nodes, cnt.tris, rays);
    if (KIND == 3 && !retraceLaunch && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&st->paths, (unsigned long long)n);
}

// ---- the any-hit kernel over the compressed 8-wide tree (gnx_bvh8.cuh) ---------------------------------------
// Shadow rays (queue half A) and environment-MIS rays (half B) of one bounce: same persistent scheme as k_trace (one ray
// per lane, refill through a warp-aggregated cursor, warp-synchronous rounds), but a round is one 8-wide node per lane:
// six 16-byte loads feed eight slab tests, the stack holds one entry per node (node, children still to visit), and there
// is no visiting order to keep.  Leaves wait until kLeafBatch lanes hold one, like in k_trace.
//   VARIANT 0: both queue halves (A first); VARIANT 1: the same, launched next to an extension launch (booked to that stage)
#ifndef GNX_ANY8_BLOCKS
#define GNX_ANY8_BLOCKS 8
#endif
#ifndef GNX_ANY8_REFETCH
#define GNX_ANY8_REFETCH 24
#endif
template <int VARIANT>
__global__ void __launch_bounds__(kBlock, GNX_ANY8_BLOCKS) k_anyhit8(const DeviceScene sc, PathState ps, Queues q, RenderConsts rc, DevStats *st) {
    __shared__ int2 s_stack[kSmemStack * kBlock];
    int2 *stack = s_stack + threadIdx.x;
    const uint32_t sb = stack_shared_base(stack);
    const int lane = threadIdx.x & 31;
    const int nShA = q.counts[kCntShadow], nShB = q.counts[kCntShadow + 1];
    const int n = nShA + nShB;
    int *cursor = &q.counts[kCntFetch + 1];
    auto shadowItem = [&](int i) { return i >= nShA ? q.shadow_q + (size_t)q.capacity + (i - nShA) : q.shadow_q + i; };
    TraversalCounters cnt{0, 0};
    unsigned rays = 0, raysB = 0;
    Trav8 t;
    int2 spill[kSpillStack];
    t.spill = spill;
    t.cur = kRefNone;
    t.gnode = 0; t.gmask = 0;
    bool active = false, exhausted = false;
    int item = 0;
    while (true) {
        if (!exhausted) {
            const unsigned idle = __ballot_sync(kFull, !active);
            if (idle) {
                const int leader = __ffs(idle) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(cursor, __popc(idle));
                base = __shfl_sync(kFull, base, leader);
                if (!active) {
                    const int i = base + __popc(idle & ((1u << lane) - 1));
                    if (i < n) {
                        item = i;
                        shadow_begin(sc, shadowItem(i), t);
                        trav8_init(sc, t);
                        if (i >= nShA) ++raysB;
                        active = true;
                        ++rays;
                    }
                }
                exhausted = base + __popc(idle) >= n;
            }
        }
        if (!__any_sync(kFull, active)) break;
        while (true) {
            if (active && trav_needs_pop(t)) trav8_next(sc, t, stack, kBlock, sb);
            bool parked = active && trav_is_leaf(t);
            if (active && !parked && trav_is_interior(t)) { trav8_interior<false>(sc, t, stack, kBlock, cnt, sb); parked = trav_is_leaf(t); }
            const unsigned parkedMask = __ballot_sync(kFull, parked);
            const unsigned advMask = __ballot_sync(kFull, active && !parked && !trav_done(t));
            if (parked && (__popc(parkedMask) >= kLeafBatch || advMask == 0))
                t.cur = trav_leaf_ref<true>(sc, t, t.cur, cnt) ? kRefNone : kRefPop;
            if (active && trav_done(t)) {
                shadow_finish(ps, shadowItem(item), t, item >= nShA);
                active = false;
            }
            const unsigned actMask = __ballot_sync(kFull, active);
            if (actMask == 0 || (!exhausted && __popc(actMask) < GNX_ANY8_REFETCH)) break;
        }
    }
    flush_stats(st, 1, cnt.nodes, cnt.tris, rays - raysB);
    flush_stats(st, 2, 0, 0, raysB);
    if (VARIANT == 1) {  // launched inside the extend stage: its work belongs to that stage's algorithmic bytes
        unsigned nn = cnt.nodes, nt = cnt.tris;
        for (int o = 16; o > 0; o >>= 1) {
            rays += __shfl_down_sync(kFull, rays, o);
            nn += __shfl_down_sync(kFull, nn, o);
            nt += __shfl_down_sync(kFull, nt, o);
        }
        if (lane == 0 && rays) {
            atomicAdd(&st->shadow_rays_in_extend_launches, (unsigned long long)rays);
            atomicAdd(&st->any_nodes_in_extend, (unsigned long long)nn);
            atomicAdd(&st->any_tris_in_extend, (unsigned long long)nt);
        }
    }
}

// Escaped rays of a scene with a SkyBoxLight (queued by the traversal kernel): L += beta * Le(ray).
#ifndef GNX_KERNELS_TEMPLATES_ONLY
__global__ void k_escape(const DeviceScene sc, PathState ps, Queues q) {
    const int n = q.counts[kCntMiss];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) escape_slot(sc, ps, q.miss_q[i]);
}
#endif

#ifndef GNX_KERNELS_TEMPLATES_ONLY
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
#endif

#ifndef GNX_SHADE_BLOCK
#define GNX_SHADE_BLOCK 128
#endif
#ifndef GNX_SHADE_SYNC
#define GNX_SHADE_SYNC 0  // measured: block barriers between stages help C1 (-10 %), cost C2 / C3 (+5 %)
#endif
constexpr int kShadeBlock = GNX_SHADE_BLOCK;
// Resident blocks per SM asked of the compiler (it caps the registers accordingly).  The kernel is bound by the
// latency of its dependent table loads, so more warps pay although the code then spills: measured on C2 / C3 shade
// stage, 4 blocks (128 regs) 4.85 / 27.9 ms, 6: 4.59 / 26.2, 7: 4.52 / 24.0, 8 (64 regs): 4.59 / 22.9 (Disney).
#ifndef GNX_SHADE_MINBLOCKS
#define GNX_SHADE_MINBLOCKS(MAXL) ((MAXL) > 2 ? 8 : 7)
#endif
template <int MAXL, bool NEXT = false>
__global__ void __launch_bounds__(kShadeBlock, GNX_SHADE_MINBLOCKS(MAXL)) k_shade(const DeviceScene sc, PathState ps, Queues q, RenderConsts rc,
                                                        int type, int outQ) {
    const int n = q.counts[kCntShade0 + type];
    const int *list = q.shade_q + (size_t)type * q.capacity;
    const int stride = gridDim.x * blockDim.x;
    for (int base = blockIdx.x * blockDim.x; base < n; base += stride) {
        const int i = base + threadIdx.x;
        int slot = 0;
        ShadeOut out;
        if (i < n) slot = list[i];
        // every thread of the block walks through the stages (barriers inside), with or without an item
        shade_slot<MAXL, GNX_SHADE_SYNC != 0, NEXT>(sc, ps, rc, slot, out, i < n);
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

// VolPathIntegrator: one path per lane from camera to termination (gnx_volpath.cuh explains why this
// integrator is not cut into wavefront stages).  Lanes pull camera samples through a warp-aggregated cursor.
// Resident blocks per SM for the two per-lane kernels (latency-bound: more warps pay despite the spills).  Measured:
// k_volpath C4 4 blocks 2.1 s, 6: 1.97 s, 8: 1.85 s; k_recursive W1 / D1 4: 25.5 / 34.0 ms, 6: 23.9 / 32.0, 8: 25.0 / 30.2.
template <bool TEX>
__global__ void __launch_bounds__(kBlock, 8) k_volpath(const DeviceScene sc, PathState ps, Queues q, RenderConsts rc, DevStats *st) {
    __shared__ int2 s_stack[kSmemStack * kBlock];
    int2 *stack = s_stack + threadIdx.x;
    const int n = rc.npix * rc.batch_spp;
    const int lane = threadIdx.x & 31;
    int *cursor = &q.counts[kCntFetch];
    TraversalCounters cnt{0, 0};
    VolCounters vc{0, 0, 0};
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(cursor, 32);
        base = __shfl_sync(kFull, base, 0);
        if (base >= n) break;
        const int slot = base + lane;
        if (slot < n) {
            int pixel, sample, px, py;
            slot_to_sample(rc, slot, &pixel, &sample);
            V3 L(0.f);
            if (pixel_xy(rc, pixel, &px, &py)) L = volpath_li<TEX>(sc, rc, px, py, sample, stack, kBlock, cnt, vc);
            ps.L[slot] = make_float4(L.x, L.y, L.z, 0.f);
        }
        __syncwarp();
    }
    // all traversal work is booked under the ray kind that issued it
    unsigned nodes = cnt.nodes, tris = cnt.tris;
    flush_stats(st, 0, nodes, tris, vc.extend);
    flush_stats(st, 1, 0, 0, vc.shadow);
    flush_stats(st, 2, 0, 0, vc.mis);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&st->paths, (unsigned long long)n);
}

// ---- VolPathIntegrator as a staged wavefront (gnx_volwave.cuh) -------------------------------------------------------
// Queues (the PathIntegrator's arrays, re-used): walks waiting for k_vp_track in extend_q[0] as (slot << 1 | mode), count
// in counts[0]; paths waiting for a logic kernel in shade_q[VolQueue], counts in counts[kCntShade0 + queue].  One fetch
// cursor (counts[kCntFetch]) serves every launch: k_vp_reset zeroes it, and the count of the queue the previous launch
// has consumed, between two launches.
#ifndef GNX_KERNELS_TEMPLATES_ONLY
__global__ void k_vp_reset(int *counts, int consumed) {
    const int i = threadIdx.x;
    if (consumed == -1) { if (i < kNumCounters) counts[i] = 0; }  // before the start launch: everything
    else if (i == consumed || i == kCntFetch) counts[i] = 0;      // (-2: the cursor only)
}
#endif

// One logic kernel (VolKernel) over one of its queues; queue < 0: the camera samples of the batch (VP_START).
#ifndef GNX_VOL_MAXL
#define GNX_VOL_MAXL 8
#endif
#ifndef GNX_VP_LOGIC_BLOCKS
#define GNX_VP_LOGIC_BLOCKS 4  // measured on C4: 4 blocks (128 registers) 148 ms, 6: 158 ms, 8: 177 ms — the spills cost more than the warps hide
#endif
// MAXL: lobe capacity of the BSDFs the kernel builds — 8 with a DisneyMaterial in the scene, 2 otherwise (like k_shade<2 | 8>).
// The per-thread frame (Surface, Bsdf, walk state) lives in local memory; across all resident threads it is larger than
// L2, so its size is DRAM traffic: the vertex kernel of C4 takes 54.8 ms with MAXL 8 and 49.8 ms with MAXL 2.
template <int KERNEL, bool TEX = false, int MAXL = GNX_VOL_MAXL>
__global__ void __launch_bounds__(kBlock, GNX_VP_LOGIC_BLOCKS) k_vp_logic(const DeviceScene sc, PathState ps, VolWave vw, Queues q, RenderConsts rc, int queue,
                                                        DevStats *st) {
    __shared__ int2 s_stack[kSmemStack * kBlock];
    int2 *stack = s_stack + threadIdx.x;
    const int n = queue < 0 ? rc.npix * rc.batch_spp : q.counts[kCntShade0 + queue];
    const int *list = queue < 0 ? nullptr : q.shade_q + (size_t)queue * q.capacity;
    const int entry = queue < 0 ? VP_START : vol_queue_phase(queue);
    int *cursor = &q.counts[kCntFetch];
    const int lane = threadIdx.x & 31;
    TraversalCounters cnt{0, 0};
    VolCounters vc{0, 0, 0};
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(cursor, 32);
        base = __shfl_sync(kFull, base, 0);
        if (base >= n) break;
        const int i = base + lane;
        int slot = 0, y = VY_DONE;
        if (i < n) {
            slot = list ? list[i] : i;
            y = vol_advance<MAXL, TEX>(sc, rc, ps, vw, slot, entry, KERNEL, stack, kBlock, cnt, vc);
        }
        __syncwarp();
        int idx = warp_push(&q.counts[kCntExtend0], y == VY_TRACK_MAIN || y == VY_TRACK_SUB);
        if (idx >= 0) q.extend_q[0][idx] = (slot << 1) | (y == VY_TRACK_SUB ? 1 : 0);
#pragma unroll
        for (int k = 0; k < VQ_COUNT; ++k) {
            if (vol_queue_kernel(k) == KERNEL) continue;  // a kernel never queues for itself
            idx = warp_push(&q.counts[kCntShade0 + k], y == VY_QUEUE0 + k);
            if (idx >= 0) q.shade_q[(size_t)k * q.capacity + idx] = slot;
        }
    }
    flush_stats(st, 0, cnt.nodes, cnt.tris, vc.extend);
    flush_stats(st, 1, 0, 0, vc.shadow);
    flush_stats(st, 2, 0, 0, vc.mis);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (queue < 0) atomicAdd(&st->paths, (unsigned long long)n);
        atomicAdd(&st->vp_items[KERNEL], (unsigned long long)n);
    }
}

// The tracking walks of GridDensityMedium (delta tracking for the medium sample of a path segment, ratio tracking for a
// shadow / MIS walk segment): persistent lanes, each holding one walk; every iteration of the inner loop is ONE tracking
// step for all lanes that hold a walk, and lanes whose walk has ended take the next queue entry as soon as kTrackRefill
// of them are idle — the lanes stay full whatever the lengths of the individual walks.
#ifndef GNX_TRACK_REFILL
#define GNX_TRACK_REFILL 8
#endif
#ifndef GNX_TRACK_BURST
#define GNX_TRACK_BURST 8
#endif
#ifndef GNX_KERNELS_TEMPLATES_ONLY
__global__ void __launch_bounds__(kBlock, 8) k_vp_track(const DeviceScene sc, PathState ps, VolWave vw, Queues q, RenderConsts rc, DevStats *st) {
    const int n = q.counts[kCntExtend0];
    const int *list = q.extend_q[0];
    int *cursor = &q.counts[kCntFetch];
    const int lane = threadIdx.x & 31;
    TrackLane tl;
    tl.slot = 0; tl.mode = 0; tl.medium = 0;
    PathSampler smp(sc.smp, 0, 0);
    bool active = false, exhausted = false;
    int pendQueue = -1, pendSlot = 0;   // finished walk: resume queue of its path
    unsigned long long steps = 0;
    while (true) {
        // paths whose walk ended since the last visit go to the queue of their resume phase, one atomic per warp and queue
        if (__any_sync(kFull, pendQueue >= 0)) {
#pragma unroll
            for (int k = 0; k < VQ_COUNT; ++k) {
                if (k != VQ_VERTEX && k != VQ_SHADOW_RESUME && k != VQ_MIS_RESUME) continue;
                const int idx = warp_push(&q.counts[kCntShade0 + k], pendQueue == k);
                if (idx >= 0) q.shade_q[(size_t)k * q.capacity + idx] = pendSlot;
            }
            pendQueue = -1;
        }
        const unsigned idle = __ballot_sync(kFull, !active);
        if (!exhausted && __popc(idle) >= GNX_TRACK_REFILL) {
            const int leader = __ffs(idle) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(cursor, __popc(idle));
            base = __shfl_sync(kFull, base, leader);
            if (!active) {
                const int i = base + __popc(idle & ((1u << lane) - 1));
                if (i < n) {
                    const int item = list[i];
                    const int slot = item >> 1, mode = item & 1;
                    int pixel, sample, px, py;
                    slot_to_sample(rc, slot, &pixel, &sample);
                    pixel_xy(rc, pixel, &px, &py);
                    smp.take(vol_load_sampler(sc, rc, ps, vw, slot, px, py, sample));
                    if (vol_track_begin(sc, ps, vw, slot, mode, tl)) active = true;
                    else {  // the ray misses the medium's box: no steps, no draws
                        vol_track_finish(sc, ps, vw, tl, smp);
                        pendSlot = slot;
                        pendQueue = vol_resume_queue(mode, (int)(ps.meta[slot] & kVolPhaseMask));
                    }
                }
            }
            exhausted = base + __popc(idle) >= n;
        }
        if (!__any_sync(kFull, active || pendQueue >= 0)) { if (exhausted) break; else continue; }
#pragma unroll 1
        for (int k = 0; k < GNX_TRACK_BURST; ++k) {
            if (active) {
                ++steps;
                if (track_step(sc.media[tl.medium], tl.ts, smp)) {
                    vol_track_finish(sc, ps, vw, tl, smp);
                    pendSlot = tl.slot;
                    pendQueue = vol_resume_queue(tl.mode, (int)(ps.meta[tl.slot] & kVolPhaseMask));
                    active = false;
                }
            }
            // stop the burst early once enough lanes are idle to be worth a refill (or nobody is left)
            const unsigned act = __ballot_sync(kFull, active);
            if (act == 0 || (!exhausted && 32 - __popc(act) >= GNX_TRACK_REFILL)) break;
        }
    }
    for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(kFull, steps, o);
    if (lane == 0 && steps) atomicAdd(&st->track_steps, steps);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&st->vp_items[4], (unsigned long long)n);
}
#endif

// WhittedIntegrator / DirectLightingIntegrator: one camera sample per lane, the recursion as a depth-first frame
// stack (gnx_whitted.cuh).  Same dynamic fetch as k_volpath.
#ifndef GNX_REC_MAXL
#define GNX_REC_MAXL 8
#endif
template <int DIRECT, bool TEX>  // DIRECT: 0 Whitted, 1 DirectLighting UniformSampleOne, 2 UniformSampleAll; TEX: carry ray differentials
__global__ void __launch_bounds__(kBlock, 6) k_recursive(const DeviceScene sc, PathState ps, Queues q, RenderConsts rc, DevStats *st) {
    __shared__ int2 s_stack[kSmemStack * kBlock];
    int2 *stack = s_stack + threadIdx.x;
    // rc.rec_list: only the samples the staged first vertex turned down (k_whitted_vertex), counted by the camera-ray launch
    const int *list = rc.rec_list ? q.extend_q[rc.rec_list - 1] : nullptr;
    const int n = list ? q.counts[rc.rec_list - 1] : rc.npix * rc.batch_spp;
    const int lane = threadIdx.x & 31;
    int *cursor = &q.counts[kCntFetch];
    TraversalCounters cnt{0, 0};
    RecCounters rcnt{0, 0, 0};
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(cursor, 32);
        base = __shfl_sync(kFull, base, 0);
        if (base >= n) break;
        if (base + lane < n) {
            const int slot = list ? list[base + lane] : base + lane;
            int pixel, sample, px, py;
            slot_to_sample(rc, slot, &pixel, &sample);
            V3 L(0.f);
            if (pixel_xy(rc, pixel, &px, &py)) L = recursive_li<GNX_REC_MAXL, DIRECT, TEX>(sc, rc, px, py, sample, stack, kBlock, cnt, rcnt);
            ps.L[slot] = make_float4(L.x, L.y, L.z, 0.f);
        }
        __syncwarp();
    }
    unsigned nodes = cnt.nodes, tris = cnt.tris;
    flush_stats(st, 0, nodes, tris, rcnt.extend);
    flush_stats(st, 1, 0, 0, rcnt.shadow);
    flush_stats(st, 2, 0, 0, rcnt.mis);
    if (!list && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&st->paths, (unsigned long long)n);
}

// WhittedIntegrator's first vertex as a stage (gnx_whitted.cuh): over the camera rays' hits of one shade queue.  A vertex
// without specular lobes gets its emitted light into ps.L and one shadow item per light that can contribute (plane index
// light * capacity + slot in d_path.w); the others are listed in extend queue 0 for k_recursive.
template <bool TEX>
__global__ void __launch_bounds__(kBlock, 4) k_whitted_vertex(const DeviceScene sc, PathState ps, Queues q, RenderConsts rc, int type) {
    const int n = q.counts[kCntShade0 + type];
    const int *list = q.shade_q + (size_t)type * q.capacity;
    const int stride = gridDim.x * blockDim.x;
    for (int base = blockIdx.x * blockDim.x; base < n; base += stride) {
        const int i = base + threadIdx.x;
        int slot = 0;
        bool staged = false, recurse = false;
        WhittedVertex<8> v;
        if (i < n) {
            slot = list[i];
            int pixel, sample, px, py;
            slot_to_sample(rc, slot, &pixel, &sample);
            pixel_xy(rc, pixel, &px, &py);
            const float4 hit = ps.hit[slot];
            V3 Le;
            staged = whitted_vertex_begin<8, TEX>(sc, px, py, sample, f2i(hit.w), hit.x, hit.y, hit.z, v, &Le);
            recurse = !staged;
            if (staged) ps.L[slot] = make_float4(Le.x, Le.y, Le.z, 0.f);
        }
        int idx = warp_push(&q.counts[kCntExtend0], recurse);
        if (idx >= 0) q.extend_q[0][idx] = slot;
        for (int j = 0; j < sc.n_lights; ++j) {
            ShadowItem it;
            const bool have = staged && whitted_vertex_light<8>(sc, v, j, &it);
            idx = warp_push(&q.counts[kCntShadow], have);
            if (idx >= 0) {
                it.d_path.w = i2f(j * q.capacity + slot);
                q.shadow_q[idx] = it;
            }
        }
    }
}

// L = Le + lightL, lightL = the visible lights' contributions in light order (WhittedIntegrator.cpp:42-60).  Slots the
// stage did not handle have zero planes: escaped rays keep their radiance, recursion slots are overwritten afterwards.
#ifndef GNX_KERNELS_TEMPLATES_ONLY
__global__ void k_whitted_sum(float4 *L, const float4 *planes, int nLights, int capacity, int n) {
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += gridDim.x * blockDim.x) {
        float lx = 0.f, ly = 0.f, lz = 0.f;
        for (int j = 0; j < nLights; ++j) {
            const float4 c = planes[(size_t)j * capacity + slot];
            lx += c.x; ly += c.y; lz += c.z;
        }
        float4 v = L[slot];
        v.x += lx; v.y += ly; v.z += lz;
        v.w = 0.f;
        L[slot] = v;
    }
}
__global__ void k_zero_counter(int *counts, int i) {
    if (threadIdx.x == 0) counts[i] = 0;
}
#endif

// colObj += Li(...) over the samples of the pixel, in sample order (core/Integrator.cpp:274-291)
#ifndef GNX_KERNELS_TEMPLATES_ONLY
__global__ void k_accumulate(PathState ps, float4 *accum, RenderConsts rc) {
    for (int pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < rc.npix; pixel += gridDim.x * blockDim.x) {
        float4 a = accum[pixel];
        for (int s = 0; s < rc.batch_spp; ++s) {
            float4 L = ps.L[(size_t)pixel * rc.batch_spp + s];
            if (L.w != 0.f) {  // the path has separate accumulators for its shadow / environment-MIS contributions
                if (ps.La) { const float4 La = ps.La[(size_t)pixel * rc.batch_spp + s]; L.x += La.x; L.y += La.y; L.z += La.z; }
                if (ps.Lb) { const float4 Lb = ps.Lb[(size_t)pixel * rc.batch_spp + s]; L.x += Lb.x; L.y += Lb.y; L.z += Lb.z; }
            }
            a.x += L.x; a.y += L.y; a.z += L.z;
        }
        accum[pixel] = a;
    }
}
#endif

// Gaussian film, step 1 (per path slot): the sample's radiance summed into L.xyz and its film offset written into
// ps.film_off (dense in slot order).
#ifndef GNX_KERNELS_TEMPLATES_ONLY
__global__ void k_film_prepare(const DeviceScene sc, PathState ps, RenderConsts rc) {
    const int n = rc.npix * rc.batch_spp;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += gridDim.x * blockDim.x) {
        int pixel, sample;
        slot_to_sample(rc, slot, &pixel, &sample);
        float u0, u1;
        film_sample_offset(sc, rc.width, pixel % rc.width, pixel / rc.width, sample, &u0, &u1);  // (no tile partition with the Gaussian film)
        float4 L = ps.L[slot];
        if (L.w != 0.f) {
            if (ps.La) { const float4 La = ps.La[slot]; L.x += La.x; L.y += La.y; L.z += La.z; }
            if (ps.Lb) { const float4 Lb = ps.Lb[slot]; L.x += Lb.x; L.y += Lb.y; L.z += Lb.z; }
            ps.L[slot] = make_float4(L.x, L.y, L.z, 0.f);
        }
        ps.film_off[slot] = make_float4(u0, u1, 0.f, 0.f);  // (shares Lb's storage: written after Lb has been folded in)
    }
}
#endif

// Gaussian film, step 2 (per pixel): gather the batch's samples within the filter's reach; accum = (sum L f, sum f).
#ifndef GNX_KERNELS_TEMPLATES_ONLY
__global__ void k_accumulate_gauss(PathState ps, float4 *accum, RenderConsts rc, FilmFilter f) {
    for (int pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < rc.npix; pixel += gridDim.x * blockDim.x) {
        const float4 g = gaussian_gather(ps.L, ps.film_off, rc, f, pixel % rc.width, pixel / rc.width);
        float4 a = accum[pixel];
        a.x += g.x; a.y += g.y; a.z += g.z; a.w += g.w;
        accum[pixel] = a;
    }
}
#endif

// Gaussian film, step 2 as a shared-memory tiled gather: a block owns a 32 x 8 tile of output pixels and stages, for
// `chunk` samples per pixel at a time, the radiance of the (32 + 2 reach) x (8 + 2 reach) source pixels around it
// together with each sample's 1-D filter weights toward the 2 reach + 1 output columns and rows it can touch
// (2 (2 reach + 1) exponentials per staged sample instead of two per (output pixel, sample) pair, and coalesced loads
// instead of one sample stream per lane).  The weights vanish outside the radius (max(0, exp(-a d^2) - exp(-a r^2))), so
// Film::AddSample's pixel bounds need no test of their own.  A warp is one row of 32 pixels and the per-pixel sample
// stride is odd: the shared-memory reads are conflict-free.  Threads sum in a fixed order: deterministic.
// RT / CT > 0: reach and chunk known at compile time (the index divisions become multiplications).
constexpr int kFilmTW = 32, kFilmTH = 8;
template <int RT, int CT>
__global__ void __launch_bounds__(kFilmTW * kFilmTH) k_accumulate_gauss_tiled(PathState ps, float4 *accum, RenderConsts rc, FilmFilter f,
                                                                              int chunkArg, int tilesX, int tilesY) {
    extern __shared__ float film_sm[];
    const int r = RT > 0 ? RT : f.reach, chunk = CT > 0 ? CT : chunkArg;
    const int nw = 2 * r + 1, SW = kFilmTW + 2 * r, SH = kFilmTH + 2 * r, npx = SW * SH, stride = chunk | 1;
    const int plane = npx * stride;
    float *sLx = film_sm, *sLy = film_sm + plane, *sLz = film_sm + 2 * plane, *sWX = film_sm + 3 * plane, *sWY = sWX + nw * plane;
    const int lx = threadIdx.x % kFilmTW, ly = threadIdx.x / kFilmTW;
    for (int tile = blockIdx.x; tile < tilesX * tilesY; tile += gridDim.x) {
        const int tx0 = (tile % tilesX) * kFilmTW, ty0 = (tile / tilesX) * kFilmTH;
        float ax = 0.f, ay = 0.f, az = 0.f, aw = 0.f;
        for (int s0 = 0; s0 < rc.batch_spp; s0 += chunk) {
            const int n = rc.batch_spp - s0 < chunk ? rc.batch_spp - s0 : chunk;
            __syncthreads();
            // (compile-time trip count for the <RT, CT> variants: unrolled, so that the global loads of a thread's five or six
            // items are in flight together instead of one dependent load -> exp -> store chain per item)
#pragma unroll
            for (int i = threadIdx.x; i < npx * chunk; i += kFilmTW * kFilmTH) {
                const int spx = i / chunk, s = i - spx * chunk;
                const int row = spx / SW;
                const int sx = tx0 - r + (spx - row * SW), sy = ty0 - r + row;
                const int idx = spx * stride + s;
                const bool in = s < n && sx >= 0 && sx < rc.width && sy >= 0 && sy < rc.height;
                float4 l = make_float4(0.f, 0.f, 0.f, 0.f), o = l;
                if (in) {
                    const size_t slot = ((size_t)sy * rc.width + sx) * rc.batch_spp + s0 + s;
                    l = ps.L[slot];
                    o = ps.film_off[slot];
                }
                sLx[idx] = l.x; sLy[idx] = l.y; sLz[idx] = l.z;
                const float pdx = ((float)sx + o.x) - 0.5f, pdy = ((float)sy + o.y) - 0.5f;
#pragma unroll
                for (int k = 0; k < nw; ++k) {
                    sWX[k * plane + idx] = in ? gaussian_1d(f, (float)(sx + k - r) - pdx) : 0.f;
                    sWY[k * plane + idx] = in ? gaussian_1d(f, (float)(sy + k - r) - pdy) : 0.f;
                }
            }
            __syncthreads();
            // source pixel (x + dx - r, y + dy - r) reaches output column x through its weight number 2r - dx
#pragma unroll
            for (int dy = 0; dy < nw; ++dy)
#pragma unroll
                for (int dx = 0; dx < nw; ++dx) {
                    const int base = ((ly + dy) * SW + lx + dx) * stride;
                    const float *wx = sWX + (2 * r - dx) * plane + base, *wy = sWY + (2 * r - dy) * plane + base;
#pragma unroll
                    for (int s = 0; s < chunk; ++s) {
                        const float w = wx[s] * wy[s];
                        ax = __fmaf_rn(sLx[base + s], w, ax); ay = __fmaf_rn(sLy[base + s], w, ay); az = __fmaf_rn(sLz[base + s], w, az);
                        aw += w;
                    }
                }
        }
        const int x = tx0 + lx, y = ty0 + ly;
        if (x < rc.width && y < rc.height) {
            float4 a = accum[(size_t)y * rc.width + x];
            a.x += ax; a.y += ay; a.z += az; a.w += aw;
            accum[(size_t)y * rc.width + x] = a;
        }
    }
}

// resolve = 1: sum(L f) / sum(f); 0: the raw sums (N-GPU jobs reduce them across ranks before dividing)
#ifndef GNX_KERNELS_TEMPLATES_ONLY
__global__ void k_film_gauss(const float4 *accum, float4 *rgba, int npix, int resolve) {
    for (int pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < npix; pixel += gridDim.x * blockDim.x)
        rgba[pixel] = resolve ? gaussian_resolve(accum[pixel]) : accum[pixel];
}
#endif

// colObj / samplesPerPixel, alpha 1 (core/Integrator.cpp:293,307-310).  alpha: 1, or 0 on the non-root devices of an
// N-device job whose partial framebuffers are summed afterwards.
#ifndef GNX_KERNELS_TEMPLATES_ONLY
__global__ void k_film(const float4 *accum, float4 *rgba, int npix, float spp, float alpha) {
    for (int pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < npix; pixel += gridDim.x * blockDim.x) {
        const float4 a = accum[pixel];
        rgba[pixel] = make_float4(a.x / spp, a.y / spp, a.z / spp, alpha);
    }
}
#endif
// Tile partition: this device's tiles scattered into the (zeroed) full frame; the sum over the devices' frames is the
// image, every pixel being x + 0 + ... + 0: bit-equal to the single-device render.
#ifndef GNX_KERNELS_TEMPLATES_ONLY
__global__ void k_film_tiles(const float4 *accum, float4 *rgba, RenderConsts rc, float spp) {
    for (int pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < rc.npix; pixel += gridDim.x * blockDim.x) {
        int px, py;
        if (!pixel_xy(rc, pixel, &px, &py)) continue;
        const float4 a = accum[pixel];
        rgba[(size_t)py * rc.width + px] = make_float4(a.x / spp, a.y / spp, a.z / spp, 1.f);
    }
}
#endif
// The N-device reduce as ONE kernel on the root: partial frames of the peers are read through peer-to-peer loads over
// NVLink and summed in device order (deterministic), in place into the root's frame.
constexpr int kMaxDevices = 16;
struct PeerFrames { const float4 *part[kMaxDevices]; int n; };
#ifndef GNX_KERNELS_TEMPLATES_ONLY
__global__ void k_reduce_peers(float4 *out, PeerFrames pf, int npix) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        float4 a = out[i];
        for (int g = 0; g < pf.n; ++g) {
            const float4 b = pf.part[g][i];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        out[i] = a;
    }
}
#endif

#ifndef GNX_KERNELS_TEMPLATES_ONLY
__global__ void __launch_bounds__(kBlock) k_primary_hits(const DeviceScene sc, int width, int height, int sample, int *out) {
    __shared__ int2 s_stack[kSmemStack * kBlock];
    int2 *stack = s_stack + threadIdx.x;
    const int npix = width * height;
    for (int pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < npix; pixel += gridDim.x * blockDim.x)
        out[pixel] = primary_hit_id(sc, width, pixel % width, pixel / width, sample, stack, kBlock);
}
#endif

#ifndef GNX_KERNELS_TEMPLATES_ONLY
__global__ void k_sample_dims(const DeviceScene sc, int n, const long long *index, const int *dim, float *out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = halton_sample_dimension(sc.smp, (uint64_t)index[i], dim[i]);
}
#endif

#ifndef GNX_KERNELS_TEMPLATES_ONLY
__global__ void k_build_spatial(const DeviceScene sc, float *func, float *cdf, float *fint) {
    const int nv = sc.ld.nvox[0] * sc.ld.nvox[1] * sc.ld.nvox[2];
    for (int vox = blockIdx.x * blockDim.x + threadIdx.x; vox < nv; vox += gridDim.x * blockDim.x)
        build_spatial_voxel(sc, vox, func, cdf, fint);
}
#endif

// FrameBuffer::update_f_u_c for a whole pass (ui/FrameBuffer.h:127-149): running mean over Render() calls, then the
// exposure tonemap of the updated value; set_uc(.., 3, 255) for the alpha byte.  `weight` = 1 / curRenderCount.
#ifndef GNX_KERNELS_TEMPLATES_ONLY
__global__ void k_framebuffer_update(const float4 *frame, float4 *state, uchar4 *u8, int npix, float weight) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        const float4 v = frame[i];
        float4 f = state[i];
        f.x = weight * v.x + (1.0f - weight) * f.x;
        f.y = weight * v.y + (1.0f - weight) * f.y;
        f.z = weight * v.z + (1.0f - weight) * f.z;
        state[i] = f;
        const float exposure = 0.75f;
        const float r = 1.0f - expf(-f.x * 1.0f / (1 - exposure));
        const float g = 1.0f - expf(-f.y * 1.0f / (1 - exposure));
        const float b = 1.0f - expf(-f.z * 1.0f / (1 - exposure));
        u8[i] = make_uchar4((unsigned char)(r * 255), (unsigned char)(g * 255), (unsigned char)(b * 255), 255);
    }
}
#endif

// FrameBuffer::update_f_u_c's tonemap on the first pass (ui/FrameBuffer.h:141-147)
#ifndef GNX_KERNELS_TEMPLATES_ONLY
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
#endif

// ---- translation units -------------------------------------------------------------------------------------------
// The heavy kernel templates are instantiated in their own translation units (csrc/tu_*.cu, compiled in parallel): each
// defines GNX_KERNELS_TEMPLATES_ONLY and ONE GNX_TU_<name>; gnx_render.cu (GNX_TU_MAIN) only declares them.
#if defined(GNX_TU_MAIN)
extern template __global__ void k_trace<0>(const DeviceScene, PathState, Queues, RenderConsts, int, DevStats *);
extern template __global__ void k_trace<1>(const DeviceScene, PathState, Queues, RenderConsts, int, DevStats *);
extern template __global__ void k_trace<2>(const DeviceScene, PathState, Queues, RenderConsts, int, DevStats *);
extern template __global__ void k_trace<3>(const DeviceScene, PathState, Queues, RenderConsts, int, DevStats *);
extern template __global__ void k_trace<4>(const DeviceScene, PathState, Queues, RenderConsts, int, DevStats *);
extern template __global__ void k_trace<0, true>(const DeviceScene, PathState, Queues, RenderConsts, int, DevStats *);
extern template __global__ void k_trace<3, true>(const DeviceScene, PathState, Queues, RenderConsts, int, DevStats *);
extern template __global__ void k_shade<2, false>(const DeviceScene, PathState, Queues, RenderConsts, int, int);
extern template __global__ void k_shade<2, true>(const DeviceScene, PathState, Queues, RenderConsts, int, int);
extern template __global__ void k_shade<8, false>(const DeviceScene, PathState, Queues, RenderConsts, int, int);
extern template __global__ void k_shade<8, true>(const DeviceScene, PathState, Queues, RenderConsts, int, int);
extern template __global__ void k_recursive<0, false>(const DeviceScene, PathState, Queues, RenderConsts, DevStats *);
extern template __global__ void k_recursive<0, true>(const DeviceScene, PathState, Queues, RenderConsts, DevStats *);
extern template __global__ void k_whitted_vertex<false>(const DeviceScene, PathState, Queues, RenderConsts, int);
extern template __global__ void k_whitted_vertex<true>(const DeviceScene, PathState, Queues, RenderConsts, int);
extern template __global__ void k_recursive<1, false>(const DeviceScene, PathState, Queues, RenderConsts, DevStats *);
extern template __global__ void k_recursive<1, true>(const DeviceScene, PathState, Queues, RenderConsts, DevStats *);
extern template __global__ void k_recursive<2, false>(const DeviceScene, PathState, Queues, RenderConsts, DevStats *);
extern template __global__ void k_recursive<2, true>(const DeviceScene, PathState, Queues, RenderConsts, DevStats *);
extern template __global__ void k_anyhit8<0>(const DeviceScene, PathState, Queues, RenderConsts, DevStats *);
extern template __global__ void k_anyhit8<1>(const DeviceScene, PathState, Queues, RenderConsts, DevStats *);
extern template __global__ void k_vp_logic<VK_VERTEX, false>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
extern template __global__ void k_vp_logic<VK_VERTEX, true>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
extern template __global__ void k_vp_logic<VK_MIS, false>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
extern template __global__ void k_vp_logic<VK_MIS, true>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
extern template __global__ void k_vp_logic<VK_VERTEX, false, 2>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
extern template __global__ void k_vp_logic<VK_VERTEX, true, 2>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
extern template __global__ void k_vp_logic<VK_MIS, false, 2>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
extern template __global__ void k_vp_logic<VK_MIS, true, 2>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
extern template __global__ void k_vp_logic<VK_EXTEND, false>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
extern template __global__ void k_vp_logic<VK_EXTEND, true>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
extern template __global__ void k_vp_logic<VK_SHADOW, false>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
extern template __global__ void k_volpath<false>(const DeviceScene, PathState, Queues, RenderConsts, DevStats *);
extern template __global__ void k_volpath<true>(const DeviceScene, PathState, Queues, RenderConsts, DevStats *);
#endif
#if defined(GNX_TU_TRACE)
template __global__ void k_anyhit8<0>(const DeviceScene, PathState, Queues, RenderConsts, DevStats *);
template __global__ void k_anyhit8<1>(const DeviceScene, PathState, Queues, RenderConsts, DevStats *);
template __global__ void k_trace<0>(const DeviceScene, PathState, Queues, RenderConsts, int, DevStats *);
template __global__ void k_trace<1>(const DeviceScene, PathState, Queues, RenderConsts, int, DevStats *);
template __global__ void k_trace<2>(const DeviceScene, PathState, Queues, RenderConsts, int, DevStats *);
template __global__ void k_trace<3>(const DeviceScene, PathState, Queues, RenderConsts, int, DevStats *);
template __global__ void k_trace<4>(const DeviceScene, PathState, Queues, RenderConsts, int, DevStats *);
template __global__ void k_trace<0, true>(const DeviceScene, PathState, Queues, RenderConsts, int, DevStats *);
template __global__ void k_trace<3, true>(const DeviceScene, PathState, Queues, RenderConsts, int, DevStats *);
#endif
#if defined(GNX_TU_SHADE2)
template __global__ void k_shade<2, false>(const DeviceScene, PathState, Queues, RenderConsts, int, int);
template __global__ void k_shade<2, true>(const DeviceScene, PathState, Queues, RenderConsts, int, int);
#endif
#if defined(GNX_TU_SHADE8)
template __global__ void k_shade<8, false>(const DeviceScene, PathState, Queues, RenderConsts, int, int);
template __global__ void k_shade<8, true>(const DeviceScene, PathState, Queues, RenderConsts, int, int);
#endif
#if defined(GNX_TU_REC0)
template __global__ void k_recursive<0, false>(const DeviceScene, PathState, Queues, RenderConsts, DevStats *);
template __global__ void k_recursive<0, true>(const DeviceScene, PathState, Queues, RenderConsts, DevStats *);
template __global__ void k_whitted_vertex<false>(const DeviceScene, PathState, Queues, RenderConsts, int);
template __global__ void k_whitted_vertex<true>(const DeviceScene, PathState, Queues, RenderConsts, int);
#endif
#if defined(GNX_TU_REC1)
template __global__ void k_recursive<1, false>(const DeviceScene, PathState, Queues, RenderConsts, DevStats *);
template __global__ void k_recursive<1, true>(const DeviceScene, PathState, Queues, RenderConsts, DevStats *);
#endif
#if defined(GNX_TU_REC2)
template __global__ void k_recursive<2, false>(const DeviceScene, PathState, Queues, RenderConsts, DevStats *);
template __global__ void k_recursive<2, true>(const DeviceScene, PathState, Queues, RenderConsts, DevStats *);
#endif
#if defined(GNX_TU_VOL1)
template __global__ void k_vp_logic<VK_VERTEX, false>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
template __global__ void k_vp_logic<VK_VERTEX, true>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
#endif
#if defined(GNX_TU_VOL4)
template __global__ void k_vp_logic<VK_VERTEX, false, 2>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
template __global__ void k_vp_logic<VK_VERTEX, true, 2>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
#endif
#if defined(GNX_TU_VOL5)
template __global__ void k_vp_logic<VK_MIS, false, 2>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
template __global__ void k_vp_logic<VK_MIS, true, 2>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
#endif
#if defined(GNX_TU_VOL2)
template __global__ void k_vp_logic<VK_MIS, false>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
template __global__ void k_vp_logic<VK_MIS, true>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
#endif
#if defined(GNX_TU_VOL3)
template __global__ void k_vp_logic<VK_EXTEND, false>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
template __global__ void k_vp_logic<VK_EXTEND, true>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
template __global__ void k_vp_logic<VK_SHADOW, false>(const DeviceScene, PathState, VolWave, Queues, RenderConsts, int, DevStats *);
template __global__ void k_volpath<false>(const DeviceScene, PathState, Queues, RenderConsts, DevStats *);
template __global__ void k_volpath<true>(const DeviceScene, PathState, Queues, RenderConsts, DevStats *);
#endif

}  // namespace gnx
