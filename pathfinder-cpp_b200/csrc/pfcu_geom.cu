// dice + bin: the geometry half of the path. Compiled with -fmad=false: every float operation is a separately
// rounded IEEE single-precision op in the same order as the reference's x86/SSE hybrid tiler, which is what
// makes the fill lists and backdrops bit-exact (BASELINE.json north_star; reference build has no FMA,
// CMakeLists.txt:58-62).
//
//   dice  : pathfinder/shaders/d3d11/dice.comp:127-223 stage, with the arithmetic of the hybrid tiler's
//           recursive flattening (core/d3d9/tiler.cpp:284-315, core/data/segment.cpp:12-135), its contour walk
//           (core/data/contour.cpp:110-173) and its view-box clip (tiler.cpp:46-125,138-153), so the lines are the
//           ones the CPU tiler bins.
//   bin   : pathfinder/shaders/d3d11/bin.comp:140-248 stage, with the arithmetic of
//           core/d3d9/tiler.cpp:155-279 (tile walk) and core/d3d9/object_builder.cpp:19-114
//           (fill conversion: clamp [0,4095], round-to-nearest-even; backdrop bookkeeping).
//
// B200 mapping. No linked lists, no overflow-and-retry read-backs, and the tile walk runs ONCE.
//   dice: the reference flattens a curve by recursion (one thread per segment would serialise up to ~100 leaves).
//         Here a CTA owns a chunk of segments and walks all their subdivision trees breadth-first through a
//         shared-memory node queue: every thread tests / splits one node per level, so a level of any curve is
//         processed in parallel; each node's control points come from the same sequence of de Casteljau halvings as
//         the recursion, so the leaves are bit-identical. Lines are collected in shared memory and flushed with one
//         global atomic per level and CTA.
//   bin : one thread per line for short walks; a line that crosses many tiles is walked by the whole warp: all lanes
//         replay the cheap part of the DDA (the t_max chain, which must stay a chain of float additions to be
//         bit-exact) and each lane does the expensive part (fill conversion, atomics, stores) of every 32nd step.
//         Every line owns an upper bound of staging slots (2 fills per visited tile, one warp-aggregated atomic);
//         fills are counted per tile with fire-and-forget reductions; a device-wide scan then gives every tile a
//         contiguous range and a fully parallel scatter (pfcu_tiles.cu) moves the staged fills into it (CSR), so the
//         fill stage reads each tile's fills with coalesced loads.
#include "pfcu_device.h"

namespace pfcu {

// ------------------------------------------------------------------------------------------------ shared helpers

__device__ __forceinline__ float lerp_clamped(float a, float bq, float t) {  // common/math/basic.h:50-53
    t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
    return a + (bq - a) * t;
}

__device__ __forceinline__ unsigned outcode(float x, float y, float left, float right, float bottom) {
    unsigned c = 0;  // compute_outcode, tiler.cpp:46-62; the top bound is -inf (tiler.cpp:143-144) so TOP never sets
    if (x < left) c |= 1u; else if (x > right) c |= 2u;
    if (y > bottom) c |= 8u;
    return c;
}

// clip_line_segment_to_rect (Cohen-Sutherland), tiler.cpp:65-125, against the view box with an open top.
__device__ __forceinline__ bool clip_to_view_box(float &l0, float &l1, float &l2, float &l3, float left, float right,
                                                 float bottom) {
    unsigned of = outcode(l0, l1, left, right, bottom), ot = outcode(l2, l3, left, right, bottom);
    for (int guard = 0; guard < 16; guard++) {
        if (of == 0 && ot == 0) return true;
        if ((of & ot) != 0) return false;
        const bool clip_from = of > ot;
        const unsigned oc = clip_from ? of : ot;
        float px = clip_from ? l0 : l2, py = clip_from ? l1 : l3;
        if (oc & 1u) {
            py = lerp_clamped(l1, l3, (left - l0) / (l2 - l0));
            px = left;
        } else if (oc & 2u) {
            py = lerp_clamped(l1, l3, (right - l0) / (l2 - l0));
            px = right;
        } else if (oc & 8u) {
            px = lerp_clamped(l0, l2, (bottom - l1) / (l3 - l1));
            py = bottom;
        }
        if (clip_from) { l0 = px; l1 = py; of = outcode(px, py, left, right, bottom); }
        else { l2 = px; l3 = py; ot = outcode(px, py, left, right, bottom); }
    }
    return false;
}

constexpr uint32_t BIN_LONG_STEPS = 12;  // walks longer than this go to the long-line kernel, one warp per line

__device__ __forceinline__ uint32_t walk_steps(const float4 &ln) {
    // the walk visits |dx| + |dy| + 1 tiles (tiler.cpp:191-278) and adds at most two fills per tile
    const long long tx0 = (int)floorf(ln.x * 0.0625f), ty0 = (int)floorf(ln.y * 0.0625f);
    const long long tx1 = (int)floorf(ln.z * 0.0625f), ty1 = (int)floorf(ln.w * 0.0625f);
    long long d = llabs(tx1 - tx0) + llabs(ty1 - ty0) + 1;
    if (d > MAX_DDA_STEPS) d = MAX_DDA_STEPS;
    return (uint32_t)d;
}

// Where a line's fills will be staged, decided as soon as the line exists (dice): 2 slots per tile of its walk. Lines
// with long walks are queued for k_bin_long at the same time, so that the two bin kernels need nothing from each other
// and run side by side. One line per thread, per-thread atomics (the rare paths of dice).
template <class B>
__device__ __forceinline__ uint32_t reserve_line_slots(const B &b, const float4 &ln, uint32_t g) {
    const uint32_t steps = walk_steps(ln), slots = 2u * steps;
    const uint32_t slot0 = atomicAdd(&b.counters->n_staging, slots);
    if (slot0 + slots > b.staging_capacity) {
        atomicOr(&b.counters->overflow, (uint32_t)OVF_STAGING);
        return 0xffffffffu;
    }
    if (steps > BIN_LONG_STEPS) b.long_lines[atomicAdd(&b.counters->n_long, 1u)] = g;  // capacity == line capacity
    return slot0;
}

// ------------------------------------------------------------------------------------------------ dice

struct Cubic {
    float2 p0, p1, p2, p3;
};

// Segment::is_flat_cubic, core/data/segment.cpp:12-20
__device__ __forceinline__ bool is_flat_cubic(const Cubic &c) {
    float u0 = 3.0f * c.p1.x - c.p0.x - c.p0.x - c.p3.x;
    float u1 = 3.0f * c.p1.y - c.p0.y - c.p0.y - c.p3.y;
    float u2 = 3.0f * c.p2.x - c.p3.x - c.p3.x - c.p0.x;
    float u3 = 3.0f * c.p2.y - c.p3.y - c.p3.y - c.p0.y;
    u0 = u0 * u0;
    u1 = u1 * u1;
    u2 = u2 * u2;
    u3 = u3 * u3;
    float m0 = u0 > u2 ? u0 : u2;  // _mm_max_ps(a, b) = a > b ? a : b
    float m1 = u1 > u3 ? u1 : u3;
    return m0 + m1 <= FLATTENING_TOLERANCE;
}

// Segment::is_flat_quadratic, core/data/segment.cpp:22-37 (p2 of the Cubic struct is unused)
__device__ __forceinline__ bool is_flat_quadratic(const Cubic &c) {
    float mx = (c.p0.x + c.p3.x) * 0.5f, my = (c.p0.y + c.p3.y) * 0.5f;
    float dx = c.p1.x - mx, dy = c.p1.y - my;
    return dx * dx + dy * dy <= FLATTENING_TOLERANCE * 0.25f;
}

__device__ __forceinline__ float2 lerp_half(float2 a, float2 b) {  // a + 0.5 * (b - a), segment.cpp:71-81
    return make_float2(a.x + 0.5f * (b.x - a.x), a.y + 0.5f * (b.y - a.y));
}

__device__ __forceinline__ bool finite2(float2 p) { return isfinite(p.x) && isfinite(p.y); }

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, unsigned lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += t;
    }
    return v;
}

constexpr int DICE_THREADS = 256;
constexpr int DICE_CHUNK = 32;    // fewest segments a CTA takes at a time (launch_dice picks 32 .. DICE_THREADS)
// Two shared-memory configurations (template parameters Q = nodes per level held in shared memory -- wider levels fall
// back to a private stack --, O = lines collected between flushes): few segments with deep trees (an SVG at 4096^2)
// take <1024, 1024>, 98 KB, 2 CTAs per SM; many segments with shallow trees (text density) take <512, 512>, 50 KB,
// 4 CTAs per SM, so that twice as many chunks overlap their barriers and dependent loads.
#ifndef DICE_INTERLEAVE
#define DICE_INTERLEAVE 1
#endif
constexpr int DICE_PATHS = 512;   // dice metadata entries staged per chunk (more: the search stays in global memory)

template <int Q, int O>
struct DiceSharedT {
    float4 qa[2][Q];    // p0, p1
    float4 qb[2][Q];    // p2, p3
    uint32_t qm[2][Q];  // path | depth << 24 | cubic << 31
    float4 out_line[O];
    uint32_t out_path[O];
    uint32_t q_count[2];
    uint32_t out_count, out_base;
    uint32_t path_lo, path_hi;         // paths that own the chunk's first / last segment
    uint2 path_seg[DICE_PATHS];        // their (first_batch_segment_index, first_global_segment_index)
};

// One flattened line: view-box clip, then into the CTA's output buffer (or straight to global memory when full).
template <int Q, int O, class B>
__device__ __forceinline__ void dice_emit(const B &b, DiceSharedT<Q, O> &sh, float2 from, float2 to, uint32_t path) {
    float l0 = from.x, l1 = from.y, l2 = to.x, l3 = to.y;
    if (!clip_to_view_box(l0, l1, l2, l3, b.view_box[0], b.view_box[2], b.view_box[3])) return;
    const uint32_t at = atomicAdd(&sh.out_count, 1u);
    if (at < O) {
        sh.out_line[at] = make_float4(l0, l1, l2, l3);
        sh.out_path[at] = path;
    } else {
        const uint32_t g = atomicAdd(&b.counters->n_lines, 1u);
        if (g < b.line_capacity) {
            const float4 ln = make_float4(l0, l1, l2, l3);
            b.lines[g] = ln;
            b.line_meta[g] = make_uint2(path, reserve_line_slots(b, ln, g));
        } else {
            atomicOr(&b.counters->overflow, (uint32_t)OVF_LINES);
        }
    }
}

__device__ __forceinline__ void split_node(const Cubic &cur, bool cubic, Cubic &left, Cubic &right) {
    if (cubic) {  // Segment::split_cubic(0.5), segment.cpp:43-108
        const float2 p01 = lerp_half(cur.p0, cur.p1), p12 = lerp_half(cur.p1, cur.p2), p23 = lerp_half(cur.p2, cur.p3);
        const float2 p012 = lerp_half(p01, p12), p123 = lerp_half(p12, p23);
        const float2 p0123 = lerp_half(p012, p123);
        left = {cur.p0, p01, p012, p0123};
        right = {p0123, p123, p23, cur.p3};
    } else {  // Segment::split_quadratic(0.5), segment.cpp:110-135: a = p0 + (p1 - p0) * t ...
        const float2 a = make_float2(cur.p0.x + (cur.p1.x - cur.p0.x) * 0.5f, cur.p0.y + (cur.p1.y - cur.p0.y) * 0.5f);
        const float2 bq = make_float2(cur.p1.x + (cur.p3.x - cur.p1.x) * 0.5f, cur.p1.y + (cur.p3.y - cur.p1.y) * 0.5f);
        const float2 c = make_float2(a.x + (bq.x - a.x) * 0.5f, a.y + (bq.y - a.y) * 0.5f);
        left = {cur.p0, a, a, c};
        right = {c, bq, bq, cur.p3};
    }
}

__device__ __forceinline__ bool node_is_flat(const Cubic &c, bool cubic, int depth) {
    return depth >= MAX_FLATTEN_DEPTH || (cubic ? is_flat_cubic(c) : is_flat_quadratic(c));
}

// Depth-first walk of one subtree with a private stack: only used when a level does not fit the shared queue.
template <int Q, int O, class B>
__device__ __noinline__ void dice_subtree_serial(const B &b, DiceSharedT<Q, O> &sh, Cubic cur, bool cubic, int depth,
                                                 uint32_t path) {
    Cubic stack[MAX_FLATTEN_DEPTH];
    unsigned char stack_depth[MAX_FLATTEN_DEPTH];
    int sp = 0;
    while (true) {
        if (node_is_flat(cur, cubic, depth)) {
            dice_emit(b, sh, cur.p0, cur.p3, path);
            if (sp == 0) break;
            sp--;
            cur = stack[sp];
            depth = stack_depth[sp];
        } else {
            Cubic left, right;
            split_node(cur, cubic, left, right);
            depth++;
            stack[sp] = right;
            stack_depth[sp] = (unsigned char)depth;
            sp++;
            cur = left;
        }
    }
}

template <int Q, int O, class B>
__device__ __forceinline__ void dice_push(const B &b, DiceSharedT<Q, O> &sh, int buf, const Cubic &c, bool cubic,
                                          int depth, uint32_t path) {
    const uint32_t at = atomicAdd(&sh.q_count[buf], 1u);
    if (at < Q) {
        sh.qa[buf][at] = make_float4(c.p0.x, c.p0.y, c.p1.x, c.p1.y);
        sh.qb[buf][at] = make_float4(c.p2.x, c.p2.y, c.p3.x, c.p3.y);
        sh.qm[buf][at] = path | ((uint32_t)depth << 24) | (cubic ? 0x80000000u : 0u);
    } else {
        dice_subtree_serial(b, sh, c, cubic, depth, path);
    }
}

// Warp-aggregated versions of dice_emit / dice_push (one shared-memory atomic per warp instead of one per lane: 256
// lanes bumping the same counter serialise). Must be called by all 32 lanes of a warp.
template <int Q, int O, class B>
__device__ __forceinline__ void dice_emit_warp(const B &b, DiceSharedT<Q, O> &sh, bool want, float2 from, float2 to,
                                               uint32_t path, unsigned lane) {
    float l0 = from.x, l1 = from.y, l2 = to.x, l3 = to.y;
    const bool em = want && clip_to_view_box(l0, l1, l2, l3, b.view_box[0], b.view_box[2], b.view_box[3]);
    const unsigned mask = __ballot_sync(0xffffffffu, em);
    if (!mask) return;
    uint32_t base = 0;
    const int leader = __ffs(mask) - 1;
    if ((int)lane == leader) base = atomicAdd(&sh.out_count, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (!em) return;
    const uint32_t at = base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
    if (at < O) {
        sh.out_line[at] = make_float4(l0, l1, l2, l3);
        sh.out_path[at] = path;
    } else {
        const uint32_t g = atomicAdd(&b.counters->n_lines, 1u);
        if (g < b.line_capacity) {
            const float4 ln = make_float4(l0, l1, l2, l3);
            b.lines[g] = ln;
            b.line_meta[g] = make_uint2(path, reserve_line_slots(b, ln, g));
        } else {
            atomicOr(&b.counters->overflow, (uint32_t)OVF_LINES);
        }
    }
}

// Queues `n_nodes` (1 or 2) nodes per wanting lane; nodes that do not fit are flattened on the spot.
template <int Q, int O, class B>
__device__ __forceinline__ void dice_push_warp(const B &b, DiceSharedT<Q, O> &sh, int buf, bool want, int n_nodes,
                                               const Cubic &c0, const Cubic &c1, bool cubic, int depth, uint32_t path,
                                               unsigned lane) {
    const unsigned mask = __ballot_sync(0xffffffffu, want);
    if (!mask) return;
    uint32_t base = 0;
    const int leader = __ffs(mask) - 1;
    if ((int)lane == leader) base = atomicAdd(&sh.q_count[buf], (uint32_t)(n_nodes * __popc(mask)));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (!want) return;
    const uint32_t at = base + (uint32_t)(n_nodes * __popc(mask & ((1u << lane) - 1u)));
    const uint32_t meta = path | ((uint32_t)depth << 24) | (cubic ? 0x80000000u : 0u);
    for (int k = 0; k < n_nodes; k++) {
        const Cubic &c = k ? c1 : c0;
        if (at + k < Q) {
            sh.qa[buf][at + k] = make_float4(c.p0.x, c.p0.y, c.p1.x, c.p1.y);
            sh.qb[buf][at + k] = make_float4(c.p2.x, c.p2.y, c.p3.x, c.p3.y);
            sh.qm[buf][at + k] = meta;
        } else {
            dice_subtree_serial(b, sh, c, cubic, depth, path);
        }
    }
}

// Last path p in [0, n) with first_batch_segment_index[p] <= key, found by one warp with a 32-ary search: 2 dependent
// loads for 300 paths, 4 for 200 000 (dice.comp:134-149 does a binary search per thread: 9 and 18).
__device__ __forceinline__ uint32_t warp_find_path(const pfcu_dice_metadata *dice, uint32_t n, uint32_t key, unsigned lane) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const uint32_t step = (hi - lo + 31) / 32;
        const uint32_t idx = lo + lane * step;
        const bool ok = idx < hi && (idx == lo || __ldg(&dice[idx].first_batch_segment_index) <= key);
        const unsigned mask = __ballot_sync(0xffffffffu, ok);  // monotone: a prefix of the lanes (lane 0 always)
        const uint32_t j = 31u - (uint32_t)__clz((int)mask);
        const uint32_t nlo = lo + j * step;
        hi = min(nlo + step, hi);
        lo = nlo;
    }
    return lo;
}

// Batch segment number of the i-th segment to dice (identity unless the frame dices a subset: BatchView::dice_ranges)
template <class B>
__device__ __forceinline__ uint32_t dice_segment(const B &b, uint32_t i) {
    if (!b.dice_ranges) return i;
    uint32_t lo = 0, hi = b.n_dice_ranges;  // last range whose prefix count is <= i
    while (lo + 1 < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&b.dice_ranges[mid].y) <= i) lo = mid; else hi = mid;
    }
    const uint2 r = __ldg(&b.dice_ranges[lo]);
    return r.x + (i - r.y);
}

// Copies the CTA's collected lines to global memory: one atomic per flush.
template <int Q, int O, class B>
__device__ __forceinline__ void dice_flush(const B &b, DiceSharedT<Q, O> &sh) {
    // (called by all threads, between two __syncthreads() of the caller's making: out_count is stable)
    const uint32_t n_out = min(sh.out_count, (uint32_t)O);
    if (threadIdx.x == 0) sh.out_base = n_out ? atomicAdd(&b.counters->n_lines, n_out) : 0u;
    __syncthreads();
    const uint32_t base = sh.out_base;
    if (base + n_out > b.line_capacity) {
        if (threadIdx.x == 0 && n_out) atomicOr(&b.counters->overflow, (uint32_t)OVF_LINES);
    } else {
        // every line also gets its staging slots here (2 per tile of its walk; one atomic per warp) and long walks are
        // queued for k_bin_long, so the two bin kernels are independent of each other
        const unsigned lane = threadIdx.x & 31;
        for (uint32_t i0 = 0; i0 < n_out; i0 += DICE_THREADS) {  // whole warps: the scan below needs all lanes
            const uint32_t i = i0 + threadIdx.x;
            bool active = i < n_out;
            float4 ln = make_float4(0.f, 0.f, 0.f, 0.f);
            uint32_t steps = 0;
            if (active) {
                ln = sh.out_line[i];
                steps = walk_steps(ln);
            }
            const uint32_t slots = 2u * steps;
            uint32_t incl = slots;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= (unsigned)d) incl += t;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            uint32_t sbase = 0;
            if (lane == 31 && total) sbase = atomicAdd(&b.counters->n_staging, total);
            sbase = __shfl_sync(0xffffffffu, sbase, 31);
            uint32_t slot0 = sbase + incl - slots;
            if (active && slot0 + slots > b.staging_capacity) {
                atomicOr(&b.counters->overflow, (uint32_t)OVF_STAGING);
                slot0 = 0xffffffffu;
            }
            const bool is_long = active && slot0 != 0xffffffffu && steps > BIN_LONG_STEPS;
            const unsigned long_mask = __ballot_sync(0xffffffffu, is_long);
            if (long_mask) {
                uint32_t lbase = 0;
                const int leader = __ffs(long_mask) - 1;
                if ((int)lane == leader) lbase = atomicAdd(&b.counters->n_long, (uint32_t)__popc(long_mask));
                lbase = __shfl_sync(0xffffffffu, lbase, leader);
                if (is_long) b.long_lines[lbase + (uint32_t)__popc(long_mask & ((1u << lane) - 1u))] = base + i;
            }
            if (active) {
                b.lines[base + i] = ln;
                b.line_meta[base + i] = make_uint2(sh.out_path[i], slot0);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) sh.out_count = 0;
}

template <int Q, int O>
__global__ void __launch_bounds__(DICE_THREADS) k_dice(DiceArgs b, uint32_t chunk_size) {
    extern __shared__ __align__(16) unsigned char dice_smem[];
    DiceSharedT<Q, O> &sh = *reinterpret_cast<DiceSharedT<Q, O> *>(dice_smem);
    PFCU_KERNEL_BEGIN(b, PFCU_STAGE_DICE);
    // segments to dice: all of the batch, or -- incremental frames -- the ranges of the paths that changed
    const uint32_t n_dice = b.dice_ranges ? b.n_dice_segments : b.segment_count;
    const uint32_t n_chunks = (n_dice + chunk_size - 1) / chunk_size;
    if (threadIdx.x == 0) {
        sh.q_count[0] = sh.q_count[1] = 0;
        sh.out_count = 0;
    }
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    for (uint32_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        // ---- which paths own this chunk's segments: warps 0 and 1 search, everybody stages that slice of the metadata
        // A batch whose whole path table fits the staging buffer deals its segments to the chunks round-robin (segment
        // s -> chunk s mod n_chunks): the segments of one path -- similar sizes, so similar subdivision depths -- spread
        // over all CTAs instead of making one CTA walk 32 deep trees (tiger 4096^2: SM-active time 22 k cycles on average,
        // 35 k on the slowest SM with contiguous chunks). Larger batches (text-density scenes: uniform small paths) keep
        // contiguous chunks and stage only the paths of their range.
        const bool interleave = DICE_INTERLEAVE && b.path_count <= (uint32_t)DICE_PATHS && n_chunks > 1;
        const uint32_t seg0 = dice_segment(b, chunk * chunk_size), seg1 = dice_segment(b, min((chunk + 1) * chunk_size, n_dice) - 1);
        if (interleave) {
            if (threadIdx.x == 0) {
                sh.path_lo = 0;
                sh.path_hi = b.path_count - 1;
            }
        } else if (threadIdx.x < 64) {
            const uint32_t p = warp_find_path(b.dice, b.path_count, threadIdx.x < 32 ? seg0 : seg1, lane);
            if (lane == 0) (threadIdx.x < 32 ? sh.path_lo : sh.path_hi) = p;
        }
        __syncthreads();
        const uint32_t path_lo = sh.path_lo, n_staged = sh.path_hi - sh.path_lo + 1;
        const bool paths_staged = n_staged <= DICE_PATHS;
        if (paths_staged)
            for (uint32_t i = threadIdx.x; i < n_staged; i += DICE_THREADS) {
                const uint4 d = __ldg(reinterpret_cast<const uint4 *>(&b.dice[path_lo + i]));
                sh.path_seg[i] = make_uint2(d.z, d.y);
            }
        __syncthreads();
        // ---- roots: one thread per segment of the chunk (chunk_size is a multiple of 32: whole warps)
        if (threadIdx.x < chunk_size) {
            const uint32_t s_lin = interleave ? threadIdx.x * n_chunks + chunk : chunk * chunk_size + threadIdx.x;
            const uint32_t s = s_lin < n_dice ? dice_segment(b, s_lin) : 0xffffffffu;
            bool root_line = false, root_curve = false, root_cubic = false;
            Cubic root = {};
            uint32_t root_path = 0;
            if (s < b.segment_count) {
                // Owner path: last p with first_batch_segment_index <= s (dice.comp:134-149 binary search).
                uint32_t lo = 0, hi = n_staged, path, g;
                if (paths_staged) {
                    while (lo + 1 < hi) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (sh.path_seg[mid].x <= s) lo = mid; else hi = mid;
                    }
                    path = path_lo + lo;
                    g = sh.path_seg[lo].y + (s - sh.path_seg[lo].x);
                } else {
                    lo = path_lo;
                    hi = path_lo + n_staged;
                    while (lo + 1 < hi) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (__ldg(&b.dice[mid].first_batch_segment_index) <= s) lo = mid; else hi = mid;
                    }
                    path = lo;
                    g = __ldg(&b.dice[path].first_global_segment_index) +
                        (s - __ldg(&b.dice[path].first_batch_segment_index));
                }
                if (g < b.n_segments_total) {
                    const uint2 ix = __ldg(&b.indices[g]);
                    const uint32_t fp = ix.x, flag = ix.y;
                    const uint32_t next_fp = g + 1 < b.n_segments_total ? __ldg(&b.indices[g + 1]).x : b.n_points;
                    const uint32_t npt = (flag & CURVE_IS_CUBIC) ? 4u : (flag & CURVE_IS_QUADRATIC) ? 3u : 2u;
                    if (fp + npt <= b.n_points) {
                        float2 q[4];
                        bool ok = true;
#pragma unroll
                        for (uint32_t k = 0; k < 4; k++) {
                            q[k] = make_float2(0.f, 0.f);
                            if (k < npt) {
                                float2 p = __ldg(&b.points[fp + k]);
                                if (!b.identity_transform) {  // Transform2::operator*(Vec2F), common/math/transform2.h:89-91
                                    const float tx = b.transform[0] * p.x + b.transform[2] * p.y + b.transform[4];
                                    const float ty = b.transform[1] * p.x + b.transform[3] * p.y + b.transform[5];
                                    p = make_float2(tx, ty);
                                }
                                q[k] = p;
                                ok = ok && finite2(p);  // line_segment.cpp:97-104
                            }
                        }
                        // SegmentsD3D11::add_path appends points[0] after each contour (gpu_data.cpp:109): a line whose
                        // successor starts two points later is the closing line, which the hybrid tiler only emits when
                        // the contour is not already closed (contour.cpp:157-166).
                        if (npt == 2 && next_fp == fp + 2) {
                            const float dx = q[1].x - q[0].x, dy = q[1].y - q[0].y;
                            if (sqrtf(dx * dx + dy * dy) <= FLOAT_EPSILON) ok = false;
                        }
                        if (ok) {
                            root_path = path;
                            root_line = npt == 2;
                            root_curve = npt != 2;
                            root_cubic = npt == 4;
                            root = npt == 2 ? Cubic{q[0], q[1], q[1], q[1]}
                                            : npt == 3 ? Cubic{q[0], q[1], q[1], q[2]} : Cubic{q[0], q[1], q[2], q[3]};
                        }
                    }
                }
            }
            dice_emit_warp(b, sh, root_line, root.p0, root.p3, root_path, lane);
            // (a quadratic and a cubic root may share a warp: two passes keep the kind warp-uniform per call)
            dice_push_warp(b, sh, 0, root_curve && root_cubic, 1, root, root, true, 0, root_path, lane);
            dice_push_warp(b, sh, 0, root_curve && !root_cubic, 1, root, root, false, 0, root_path, lane);
        }
        __syncthreads();
        // ---- breadth-first over the subdivision trees (tiler.cpp:284-315): one node per thread per level
        int cur = 0;
        while (true) {
            const uint32_t count = min(sh.q_count[cur], (uint32_t)Q);
            // a level emits at most one line per node: flush first if they might not fit
            if (sh.out_count + count > O) {
                __syncthreads();
                dice_flush(b, sh);
                __syncthreads();
            }
            if (count == 0) break;
            for (uint32_t base = 0; base < count; base += DICE_THREADS) {
                const uint32_t i = base + threadIdx.x;
                const bool active = i < count;
                Cubic c = {};
                uint32_t m = 0;
                if (active) {
                    const float4 a = sh.qa[cur][i], bq = sh.qb[cur][i];
                    m = sh.qm[cur][i];
                    c = {make_float2(a.x, a.y), make_float2(a.z, a.w), make_float2(bq.x, bq.y), make_float2(bq.z, bq.w)};
                }
                const bool cubic = (m & 0x80000000u) != 0;
                const int depth = (int)((m >> 24) & 0x7fu);
                const uint32_t path = m & 0x00ffffffu;
                const bool flat = active && node_is_flat(c, cubic, depth);
                dice_emit_warp(b, sh, flat, c.p0, c.p3, path, lane);
                Cubic left = {}, right = {};
                if (active && !flat) split_node(c, cubic, left, right);
                dice_push_warp(b, sh, cur ^ 1, active && !flat && cubic, 2, left, right, true, depth + 1, path, lane);
                dice_push_warp(b, sh, cur ^ 1, active && !flat && !cubic, 2, left, right, false, depth + 1, path, lane);
            }
            __syncthreads();
            if (threadIdx.x == 0) sh.q_count[cur] = 0;
            cur ^= 1;
            __syncthreads();
        }
        // (both queues are empty here; the lines stay in the buffer until it fills up or the CTA is done)
    }
    __syncthreads();
    dice_flush(b, sh);
}

#ifndef DICE_WIDE_SEGMENTS
#define DICE_WIDE_SEGMENTS 4096  // batches with more segments than this take the 4-CTAs-per-SM configuration (49 KB of shared
                                 // memory per CTA instead of 98: tiger 4096^2 dices as fast -- 20.4 us -- and the other frames in
                                 // flight find more room beside it, 58.9 instead of 59.4 us per frame streamed; 65536 before)
#endif
#ifndef DICE_WIDE_CHUNK
#define DICE_WIDE_CHUNK 128
#endif

template <int Q, int O>
static cudaError_t launch_dice_cfg(const BatchView &b, cudaStream_t s, uint32_t ctas_per_sm, uint32_t max_chunk) {
    static bool configured[MAX_DEVICES] = {};  // (one per instantiation and device: function attributes are per device)
    const int dev = current_device();
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_dice<Q, O>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(DiceSharedT<Q, O>));
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    // Segments per CTA and pass: few segments -> small chunks (every SM gets one, deep trees fit the queue);
    // many segments -> large chunks (fewer block-wide barriers per segment).
    const uint32_t ctas = (uint32_t)sm_count() * ctas_per_sm;
    const uint32_t n_dice = b.dice_ranges ? b.n_dice_segments : b.segment_count;
    uint32_t chunk = (n_dice / (ctas * 4u) + 31u) & ~31u;
    chunk = chunk < DICE_CHUNK ? DICE_CHUNK : (chunk > max_chunk ? max_chunk : chunk);
    const uint32_t n_chunks = (n_dice + chunk - 1) / chunk;
    const uint32_t grid = min(n_chunks, ctas);
    return launch_pdl(k_dice<Q, O>, grid, DICE_THREADS, sizeof(DiceSharedT<Q, O>), s, DiceArgs(b), chunk);
}

cudaError_t launch_dice(const BatchView &b, cudaStream_t s) {
    if (!b.segment_count || (b.dice_ranges && !b.n_dice_segments)) return cudaSuccess;
#ifndef DICE_WIDE_Q
#define DICE_WIDE_Q 512
#define DICE_WIDE_CTAS 4
#endif
    if (b.segment_count > DICE_WIDE_SEGMENTS) return launch_dice_cfg<DICE_WIDE_Q, DICE_WIDE_Q>(b, s, DICE_WIDE_CTAS, DICE_WIDE_CHUNK);
    return launch_dice_cfg<1024, 1024>(b, s, 2u, DICE_THREADS);
}

// ------------------------------------------------------------------------------------------------ bin

struct PathTiles {
    int min_x, min_y, max_x, max_y;
    uint32_t tile_offset, backdrop_offset;
};

// ObjectBuilder::add_fill, core/d3d9/object_builder.cpp:19-64. Returns true when a fill was staged.
template <class B>
__device__ __forceinline__ bool add_fill(const B &b, const PathTiles &pt, float fx, float fy, float tx,
                                         float ty, int tcx, int tcy, uint32_t slot) {
    if (!(pt.min_x <= tcx && tcx <= pt.max_x - 1 && pt.min_y <= tcy && tcy <= pt.max_y - 1)) return false;
    const float ulx = (float)tcx * 16.0f, uly = (float)tcy * 16.0f;
    float s0 = (fx - ulx) * 256.0f, s1 = (fy - uly) * 256.0f, s2 = (tx - ulx) * 256.0f, s3 = (ty - uly) * 256.0f;
    // clamp(0, 4095) then round to nearest even (F32x4::clamp / round, common/f32x4.h:56-66)
    s0 = rintf(fminf(s0 > 0.0f ? s0 : 0.0f, 4095.0f));
    s1 = rintf(fminf(s1 > 0.0f ? s1 : 0.0f, 4095.0f));
    s2 = rintf(fminf(s2 > 0.0f ? s2 : 0.0f, 4095.0f));
    s3 = rintf(fminf(s3 > 0.0f ? s3 : 0.0f, 4095.0f));
    const uint32_t u0 = (uint32_t)s0, u1 = (uint32_t)s1, u2 = (uint32_t)s2, u3 = (uint32_t)s3;
    if (u0 == u2) return false;  // degenerate (vertical after quantisation)
    const uint32_t ti = pt.tile_offset + (uint32_t)(tcx - pt.min_x) +
                        (uint32_t)(pt.max_x - pt.min_x) * (uint32_t)(tcy - pt.min_y);
    *reinterpret_cast<uint4 *>(&b.staging[slot]) = make_uint4(ti, u0 | (u1 << 16), u2 | (u3 << 16), 0u);
    atomicAdd(&b.tile_word[ti], 1u);  // no return value: compiles to a fire-and-forget RED
    return true;
}

// ObjectBuilder::adjust_alpha_tile_backdrop, core/d3d9/object_builder.cpp:95-114
template <class B>
__device__ __forceinline__ void adjust_backdrop(const B &b, const PathTiles &pt, int tcx, int tcy, int delta) {
    const int ox = tcx - pt.min_x, oy = tcy - pt.min_y;
    const int w = pt.max_x - pt.min_x, h = pt.max_y - pt.min_y;
    if (ox < 0 || ox >= w || oy >= h) return;
    if (oy < 0) {
        atomicAdd(&b.col_backdrop[pt.backdrop_offset + (uint32_t)ox], delta);
        return;
    }
    // int8 delta lives in the top byte of the tile word (wraps mod 256 like the reference's int8_t)
    atomicAdd(&b.tile_word[pt.tile_offset + (uint32_t)ox + (uint32_t)w * (uint32_t)oy], (uint32_t)delta << 24);
}

template <class B>
__device__ __forceinline__ PathTiles load_path_tiles(const B &b, uint32_t path) {
    PathTiles pt;
    const int4 r = __ldg(reinterpret_cast<const int4 *>(&b.meta[path].tile_rect[0]));
    pt.min_x = r.x; pt.min_y = r.y; pt.max_x = r.z; pt.max_y = r.w;
    pt.tile_offset = __ldg(&b.meta[path].tile_offset);
    pt.backdrop_offset = __ldg(&b.meta[path].backdrop_offset);
    return pt;
}

// The set-up of the tile walk (process_line_segment, tiler.cpp:155-190).
struct Walk {
    float l0, l1, vx, vy;
    float t_max_x, t_max_y, t_delta_x, t_delta_y;
    int tcx, tcy, to_tx, to_ty, step_x, step_y;
    __device__ __forceinline__ void init(float a0, float a1, float a2, float a3) {
        const float ts = 16.0f;
        l0 = a0; l1 = a1;
        tcx = (int)floorf(a0 * 0.0625f); tcy = (int)floorf(a1 * 0.0625f);
        to_tx = (int)floorf(a2 * 0.0625f); to_ty = (int)floorf(a3 * 0.0625f);
        vx = a2 - a0; vy = a3 - a1;
        step_x = vx < 0 ? -1 : 1; step_y = vy < 0 ? -1 : 1;
        const float fcx = ((float)tcx + (vx >= 0 ? 1.0f : 0.0f)) * ts;
        const float fcy = ((float)tcy + (vy >= 0 ? 1.0f : 0.0f)) * ts;
        t_max_x = (fcx - a0) / vx; t_max_y = (fcy - a1) / vy;
        t_delta_x = fabsf(ts / vx); t_delta_y = fabsf(ts / vy);
    }
    // Decision of one step (tiler.cpp:191-215): direction of the next tile crossing and its clamped t.
    __device__ __forceinline__ void decide(int &next_dir, float &next_t) const {
        if (t_max_x < t_max_y) next_dir = 1;
        else if (t_max_x > t_max_y) next_dir = 2;
        else next_dir = step_x > 0 ? 1 : 2;
        next_t = next_dir == 1 ? t_max_x : t_max_y;
        next_t = next_t < 1.0f ? next_t : 1.0f;
        if (tcx == to_tx && tcy == to_ty) next_dir = 0;
    }
    __device__ __forceinline__ void advance(int next_dir) {
        if (next_dir == 1) { t_max_x += t_delta_x; tcx += step_x; }
        else if (next_dir == 2) { t_max_y += t_delta_y; tcy += step_y; }
    }
};

// The expensive part of one step (tiler.cpp:216-270): up to two fills into this step's two staging slots, backdrop
// bookkeeping. cur = where the line enters the tile, (nx, ny) = where it leaves it.
template <class B>
__device__ __forceinline__ uint32_t walk_step_emit(const B &b, const PathTiles &pt, const Walk &w, float cur_x,
                                                   float cur_y, float nx, float ny, int tcx, int tcy, int last_dir,
                                                   int next_dir, uint32_t slot) {
    const float ts = 16.0f;
    uint32_t used = 0;
    if (add_fill(b, pt, cur_x, cur_y, nx, ny, tcx, tcy, slot + used)) used++;
    if (w.step_y < 0 && next_dir == 2) {
        if (add_fill(b, pt, nx, ny, (float)tcx * ts, (float)tcy * ts, tcx, tcy, slot + used)) used++;
    } else if (w.step_y > 0 && last_dir == 2) {
        if (add_fill(b, pt, (float)tcx * ts, (float)tcy * ts, cur_x, cur_y, tcx, tcy, slot + used)) used++;
    }
    if (w.step_x < 0 && last_dir == 1) adjust_backdrop(b, pt, tcx, tcy, 1);
    else if (w.step_x > 0 && next_dir == 1) adjust_backdrop(b, pt, tcx, tcy, -1);
    return used;
}

#ifndef BIN_PERMUTE
#define BIN_PERMUTE 1
#endif
constexpr int BIN_CHAIN = 1032;          // crossings per axis the long-line kernel keeps in shared memory (16 K pixels)
constexpr int BIN_LONG_WARPS = 4;

// Incremental frames: line i is a retained line (diced in an earlier frame) of a path that has changed since
template <class B>
__device__ __forceinline__ bool stale_line(const B &b, uint32_t i, uint32_t path) {
    return i < b.n_static_lines && b.dirty_paths && ((__ldg(&b.dirty_paths[path >> 5]) >> (path & 31u)) & 1u);
}

__global__ void __launch_bounds__(128) k_bin(BinArgs b) {
    PFCU_KERNEL_BEGIN(b, PFCU_STAGE_BIN);
    const uint32_t n_lines = min(b.counters->n_lines, b.line_capacity);
    const unsigned lane = threadIdx.x & 31;
    // consecutive 32-line chunks go to consecutive CTAs: the grid is one wave sized for the largest batches, and with
    // CTA-major numbering a small batch would keep the first few hundred CTAs (two or three per SM on some SMs, none on
    // others) busy and leave the rest idle
#if BIN_PERMUTE
    const uint32_t warp0 = (threadIdx.x >> 5) * gridDim.x + blockIdx.x, n_warps = (gridDim.x * blockDim.x) >> 5;
#else
    const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
#endif
    for (uint32_t base = warp0 * 32; base < n_lines; base += n_warps * 32) {
        const uint32_t i = base + lane;
        bool active = i < n_lines;
        float4 ln = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t path = 0, steps = 0;
        uint32_t slot0 = 0xffffffffu;
        if (active) {
            ln = b.lines[i];
            const uint2 lm = b.line_meta[i];
            path = lm.x;
            slot0 = lm.y;
            steps = walk_steps(ln);
        }
        // the staging slots were reserved by dice (line_meta.y; ~0: no room, the frame is replayed with more); long
        // walks are k_bin_long's
        const uint32_t slots = 2u * steps;
        const bool is_long = steps > BIN_LONG_STEPS;
        if (slot0 == 0xffffffffu) active = false;
        if (!active || is_long) continue;
        if (stale_line(b, i, path)) {  // a retained line of a path that changed: its slots stay empty
            for (uint32_t slot = slot0; slot < slot0 + slots; slot++) b.staging[slot].tile = 0xffffffffu;
            continue;
        }
        const PathTiles pt = load_path_tiles(b, path);
        Walk w;
        w.init(ln.x, ln.y, ln.z, ln.w);
        float cur_x = ln.x, cur_y = ln.y;
        int last_dir = 0;  // 0 none, 1 X, 2 Y
        uint32_t slot = slot0;
        const uint32_t slot_end = slot0 + slots;
        for (uint32_t s = 0;; s++) {
            int next_dir;
            float next_t;
            w.decide(next_dir, next_t);
            if (s >= steps) {  // the walk left the |dx|+|dy|+1 envelope (only with non-finite arithmetic)
                atomicOr(&b.counters->overflow, (uint32_t)OVF_DDA);
                break;
            }
            const float nx = w.l0 + w.vx * next_t, ny = w.l1 + w.vy * next_t;  // LineSegmentF::sample
            slot += walk_step_emit(b, pt, w, cur_x, cur_y, nx, ny, w.tcx, w.tcy, last_dir, next_dir, slot);
            if (next_dir == 0) break;
            w.advance(next_dir);
            cur_x = nx;
            cur_y = ny;
            last_dir = next_dir;
        }
        for (; slot < slot_end; slot++) b.staging[slot].tile = 0xffffffffu;  // unused slots
    }
}

// Exact replay of the serial walk by a whole warp: all lanes run the cheap part (the t_max chain and the direction
// decisions) and lane k does the expensive part of every 32nd step. Used when the crossings of a line do not fit the
// shared-memory chains or when the merge below cannot be proven to equal the serial walk.
template <class B>
__device__ __noinline__ void walk_replay(const B &b, const PathTiles &pt, float a0, float a1, float a2, float a3,
                                         uint32_t lslot0, uint32_t lsteps, unsigned lane) {
    Walk w;
    w.init(a0, a1, a2, a3);
    int last_dir = 0;
    float prev_t = 0.0f;
    uint32_t s = 0;
    bool done = false;
    while (!done) {
        int my_dir = -1, my_last = 0, my_tcx = 0, my_tcy = 0;
        float my_t = 0.0f, my_prev_t = 0.0f;
        const uint32_t s0 = s;
#pragma unroll 4
        for (int k = 0; k < 32; k++) {
            int next_dir;
            float next_t;
            w.decide(next_dir, next_t);
            if (s >= lsteps) {
                if (lane == 0) atomicOr(&b.counters->overflow, (uint32_t)OVF_DDA);
                done = true;
                break;
            }
            if ((int)lane == k) {
                my_dir = next_dir; my_last = last_dir; my_tcx = w.tcx; my_tcy = w.tcy;
                my_t = next_t; my_prev_t = prev_t;
            }
            s++;
            if (next_dir == 0) {
                done = true;
                break;
            }
            w.advance(next_dir);
            prev_t = next_t;
            last_dir = next_dir;
        }
        if (my_dir >= 0) {
            const uint32_t my_s = s0 + lane;
            const float nx = w.l0 + w.vx * my_t, ny = w.l1 + w.vy * my_t;
            const float cx = my_s == 0 ? a0 : w.l0 + w.vx * my_prev_t, cy = my_s == 0 ? a1 : w.l1 + w.vy * my_prev_t;
            const uint32_t slot = lslot0 + 2u * my_s;
            const uint32_t used = walk_step_emit(b, pt, w, cx, cy, nx, ny, my_tcx, my_tcy, my_last, my_dir, slot);
            for (uint32_t u = used; u < 2; u++) b.staging[slot + u].tile = 0xffffffffu;
        }
    }
    for (uint32_t u = lslot0 + 2u * s + lane; u < lslot0 + 2u * lsteps; u += 32) b.staging[u].tile = 0xffffffffu;
}

// The order in which the serial walk consumes tile crossings: an X crossing at time a goes before a Y crossing at
// time b_ exactly when the walk, comparing the two, steps in X (tiler.cpp:193-205).
__device__ __forceinline__ bool x_first(float a, float b_, bool tie_x) { return a < b_ || (!(a > b_) && tie_x); }

// One warp per long line. The serial walk is a merge of two monotone sequences -- the times at which the line crosses
// vertical and horizontal tile boundaries, each built by REPEATED float addition (which is why it cannot be jumped
// into) -- so two lanes build the sequences once (one dependent add per crossing), and then every step of the walk is
// independent: a lane finds how many X crossings precede its step with a merge-path binary search and does the step's
// fill conversion, atomics and stores. 32 steps of the walk per pass instead of one.
__global__ void __launch_bounds__(BIN_LONG_WARPS * 32) k_bin_long(BinArgs b) {
    __shared__ float chain[BIN_LONG_WARPS][2][BIN_CHAIN];
    PFCU_KERNEL_BEGIN(b, PFCU_STAGE_BIN | 0x100);
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t n_long = min(b.counters->n_long, b.line_capacity);
#if BIN_PERMUTE
    const uint32_t warp0 = wib * gridDim.x + blockIdx.x, n_warps = (gridDim.x * blockDim.x) >> 5;  // (as in k_bin)
#else
    const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
#endif
    float *sx = chain[wib][0], *sy = chain[wib][1];
    for (uint32_t at = warp0; at < n_long; at += n_warps) {
        const uint32_t i = b.long_lines[at];
        const float4 ln = b.lines[i];
        const uint2 lm = b.line_meta[i];
        const uint32_t steps = walk_steps(ln), slot0 = lm.y;
        if (stale_line(b, i, lm.x)) {  // (warp-uniform)
            for (uint32_t slot = slot0 + lane; slot < slot0 + 2u * steps; slot += 32) b.staging[slot].tile = 0xffffffffu;
            continue;
        }
        const PathTiles pt = load_path_tiles(b, lm.x);
        Walk w;
        w.init(ln.x, ln.y, ln.z, ln.w);
        const uint32_t nx = (uint32_t)abs(w.to_tx - w.tcx), ny = (uint32_t)abs(w.to_ty - w.tcy);
        if (nx + 1 > BIN_CHAIN || ny + 1 > BIN_CHAIN || nx + ny + 1 != steps) {
            walk_replay(b, pt, ln.x, ln.y, ln.z, ln.w, slot0, steps, lane);
            continue;
        }
        __syncwarp();
        if (lane < 2) {  // sx[k] / sy[k]: t_max_x / t_max_y after k steps in that axis (tiler.cpp:271-277)
            float t = lane ? w.t_max_y : w.t_max_x;
            const float dt = lane ? w.t_delta_y : w.t_delta_x;
            const uint32_t n = lane ? ny : nx;
            float *dst = lane ? sy : sx;
            for (uint32_t k = 0; k <= n; k++) {
                dst[k] = t;
                t += dt;
            }
        }
        __syncwarp();
        const bool tie_x = w.step_x > 0;
        // The merge equals the serial walk iff the walk never steps past the last tile column / row before reaching
        // the last tile: once in the last column it must keep stepping in Y (and vice versa).
        const bool ok = (ny == 0 || !x_first(sx[nx], sy[ny - 1], tie_x)) && (nx == 0 || x_first(sx[nx - 1], sy[ny], tie_x));
        if (!ok) {
            walk_replay(b, pt, ln.x, ln.y, ln.z, ln.w, slot0, steps, lane);
            continue;
        }
        const uint32_t last = nx + ny;
        for (uint32_t s0 = 0; s0 <= last; s0 += 32) {
            const uint32_t s = s0 + lane;
            if (s > last) break;
            // state before step s: ix X crossings and s - ix Y crossings consumed
            uint32_t lo = s > ny ? s - ny : 0u, hi = min(s, nx);
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (x_first(sx[mid], sy[s - 1 - mid], tie_x)) lo = mid + 1; else hi = mid;
            }
            const uint32_t ix = lo, iy = s - lo;
            const float a = sx[ix], c = sy[iy];
            int next_dir = x_first(a, c, tie_x) ? 1 : 2;
            float next_t = next_dir == 1 ? a : c;
            next_t = next_t < 1.0f ? next_t : 1.0f;
            if (s == last) next_dir = 0;
            int last_dir = 0;
            float cur_x = ln.x, cur_y = ln.y;
            if (s > 0) {  // the crossing consumed last: the later of sx[ix - 1], sy[iy - 1]
                float prev_t;
                if (ix == 0) { last_dir = 2; prev_t = sy[iy - 1]; }
                else if (iy == 0) { last_dir = 1; prev_t = sx[ix - 1]; }
                else {
                    const float pa = sx[ix - 1], pc = sy[iy - 1];
                    if (x_first(pa, pc, tie_x)) { last_dir = 2; prev_t = pc; } else { last_dir = 1; prev_t = pa; }
                }
                prev_t = prev_t < 1.0f ? prev_t : 1.0f;
                cur_x = w.l0 + w.vx * prev_t;
                cur_y = w.l1 + w.vy * prev_t;
            }
            const float px = w.l0 + w.vx * next_t, py = w.l1 + w.vy * next_t;  // LineSegmentF::sample
            const int tcx = w.tcx + (int)ix * w.step_x, tcy = w.tcy + (int)iy * w.step_y;
            const uint32_t slot = slot0 + 2u * s;
            const uint32_t used = walk_step_emit(b, pt, w, cur_x, cur_y, px, py, tcx, tcy, last_dir, next_dir, slot);
            for (uint32_t u = used; u < 2; u++) b.staging[slot + u].tile = 0xffffffffu;
        }
    }
}

cudaError_t launch_bin(const BatchView &b, cudaStream_t s) {
    if (!b.segment_count) return cudaSuccess;
    return launch_pdl(k_bin, sm_count() * 8, 128, 0, s, BinArgs(b));
}

cudaError_t launch_bin_long(const BatchView &b, cudaStream_t s) {
    if (!b.segment_count) return cudaSuccess;
    return launch_pdl(k_bin_long, sm_count() * 2, BIN_LONG_WARPS * 32, 0, s, BinArgs(b));
}

}  // namespace pfcu
