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
// B200 mapping. No linked lists, no overflow-and-retry read-backs, and the tile walk runs ONCE: dice reserves, per
// clipped line, an upper bound of staging slots (2 fills per visited tile) with one atomic per warp; bin writes its
// fills there and counts them per tile with fire-and-forget reductions; a device-wide scan then gives every tile a
// contiguous range and a fully parallel scatter (pfcu_tiles.cu) moves the staged fills into it (CSR), so the fill
// kernel reads each tile's fills with coalesced loads.
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

// Upper bound of the fills one clipped line can emit: the tile walk (tiler.cpp:191-278) visits
// |dx| + |dy| + 1 tiles and adds at most two fills per tile.
__device__ __forceinline__ uint32_t fill_bound(float l0, float l1, float l2, float l3) {
    const long long tx0 = (int)floorf(l0 * 0.0625f), ty0 = (int)floorf(l1 * 0.0625f);
    const long long tx1 = (int)floorf(l2 * 0.0625f), ty1 = (int)floorf(l3 * 0.0625f);
    long long d = llabs(tx1 - tx0) + llabs(ty1 - ty0) + 1;
    if (d > MAX_DDA_STEPS) d = MAX_DDA_STEPS;
    return (uint32_t)(2 * d);
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

// Walks the subdivision tree of one curve depth-first with an explicit stack of right siblings
// (tiler.cpp:284-315 recursion). `emit(from, to)` is called once per leaf, left to right.
template <bool CUBIC, typename Emit>
__device__ __forceinline__ void flatten(Cubic cur, Emit &&emit) {
    Cubic stack[MAX_FLATTEN_DEPTH];
    unsigned char stack_depth[MAX_FLATTEN_DEPTH];
    int sp = 0, depth = 0;
    while (true) {
        bool flat = depth >= MAX_FLATTEN_DEPTH || (CUBIC ? is_flat_cubic(cur) : is_flat_quadratic(cur));
        if (flat) {
            emit(cur.p0, cur.p3);
            if (sp == 0) break;
            sp--;
            cur = stack[sp];
            depth = stack_depth[sp];
        } else {
            Cubic left, right;
            if (CUBIC) {  // Segment::split_cubic(0.5), segment.cpp:43-108
                float2 p01 = lerp_half(cur.p0, cur.p1), p12 = lerp_half(cur.p1, cur.p2), p23 = lerp_half(cur.p2, cur.p3);
                float2 p012 = lerp_half(p01, p12), p123 = lerp_half(p12, p23);
                float2 p0123 = lerp_half(p012, p123);
                left = {cur.p0, p01, p012, p0123};
                right = {p0123, p123, p23, cur.p3};
            } else {  // Segment::split_quadratic(0.5), segment.cpp:110-135: a = p0 + (p1 - p0) * t ...
                float2 a = make_float2(cur.p0.x + (cur.p1.x - cur.p0.x) * 0.5f, cur.p0.y + (cur.p1.y - cur.p0.y) * 0.5f);
                float2 b = make_float2(cur.p1.x + (cur.p3.x - cur.p1.x) * 0.5f, cur.p1.y + (cur.p3.y - cur.p1.y) * 0.5f);
                float2 c = make_float2(a.x + (b.x - a.x) * 0.5f, a.y + (b.y - a.y) * 0.5f);
                left = {cur.p0, a, a, c};
                right = {c, b, b, cur.p3};
            }
            depth++;
            stack[sp] = right;
            stack_depth[sp] = (unsigned char)depth;
            sp++;
            cur = left;
        }
    }
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

__global__ void __launch_bounds__(128) k_dice(BatchView b) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    const float left = b.view_box[0], right = b.view_box[2], bottom = b.view_box[3];
    uint32_t path = 0, npt = 0;
    Cubic c;
    bool active = false;
    if (s < b.segment_count) {
        // Owner path: last p with first_batch_segment_index <= s (dice.comp:134-149 binary search).
        uint32_t lo = 0, hi = b.path_count;
        while (lo + 1 < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if (__ldg(&b.dice[mid].first_batch_segment_index) <= s) lo = mid; else hi = mid;
        }
        path = lo;
        const uint32_t g = __ldg(&b.dice[path].first_global_segment_index) +
                           (s - __ldg(&b.dice[path].first_batch_segment_index));
        if (g < b.n_segments_total) {
            const uint2 ix = __ldg(&b.indices[g]);
            const uint32_t fp = ix.x, flag = ix.y;
            const uint32_t next_fp = g + 1 < b.n_segments_total ? __ldg(&b.indices[g + 1]).x : b.n_points;
            npt = (flag & CURVE_IS_CUBIC) ? 4u : (flag & CURVE_IS_QUADRATIC) ? 3u : 2u;
            if (fp + npt <= b.n_points) {
                float2 q[4];
#pragma unroll
                for (uint32_t k = 0; k < 4; k++) {
                    if (k < npt) {
                        float2 p = __ldg(&b.points[fp + k]);
                        if (!b.identity_transform) {  // Transform2::operator*(Vec2F), common/math/transform2.h:89-91
                            float tx = b.transform[0] * p.x + b.transform[2] * p.y + b.transform[4];
                            float ty = b.transform[1] * p.x + b.transform[3] * p.y + b.transform[5];
                            p = make_float2(tx, ty);
                        }
                        q[k] = p;
                    } else {
                        q[k] = make_float2(0.f, 0.f);
                    }
                }
                active = true;
                for (uint32_t k = 0; k < npt; k++) active = active && finite2(q[k]);  // line_segment.cpp:97-104
                // SegmentsD3D11::add_path appends points[0] after each contour (gpu_data.cpp:109): a line whose
                // successor starts two points later is the closing line, which the hybrid tiler only emits when the
                // contour is not already closed (contour.cpp:157-166).
                if (npt == 2 && next_fp == fp + 2) {
                    float dx = q[1].x - q[0].x, dy = q[1].y - q[0].y;
                    if (sqrtf(dx * dx + dy * dy) <= FLOAT_EPSILON) active = false;
                }
                if (npt == 2) c = {q[0], q[1], q[1], q[1]};
                else if (npt == 3) c = {q[0], q[1], q[1], q[2]};
                else c = {q[0], q[1], q[2], q[3]};
            }
        }
    }

    // Pass 1: count this segment's surviving lines and the staging slots they may need.
    uint32_t n = 0, slots = 0;
    auto count = [&](float2 from, float2 to) {
        float l0 = from.x, l1 = from.y, l2 = to.x, l3 = to.y;
        if (clip_to_view_box(l0, l1, l2, l3, left, right, bottom)) {
            n++;
            slots += fill_bound(l0, l1, l2, l3);
        }
    };
    if (active) {
        if (npt == 2) count(c.p0, c.p3);
        else if (npt == 3) flatten<false>(c, count);
        else flatten<true>(c, count);
    }
    // Warp-aggregated reservation: one atomic per warp and per counter.
    const uint32_t incl_n = warp_incl_scan(n, lane), incl_s = warp_incl_scan(slots, lane);
    const uint32_t total_n = __shfl_sync(0xffffffffu, incl_n, 31), total_s = __shfl_sync(0xffffffffu, incl_s, 31);
    uint32_t base_n = 0, base_s = 0;
    if (lane == 31 && total_n) {
        base_n = atomicAdd(&b.counters->n_lines, total_n);
        base_s = atomicAdd(&b.counters->n_staging, total_s);
    }
    base_n = __shfl_sync(0xffffffffu, base_n, 31);
    base_s = __shfl_sync(0xffffffffu, base_s, 31);
    if (!n) return;
    uint32_t at = base_n + incl_n - n, slot = base_s + incl_s - slots;
    if (at + n > b.line_capacity) {
        atomicOr(&b.counters->overflow, (uint32_t)OVF_LINES);
        return;
    }
    if (slot + slots > b.staging_capacity) {
        atomicOr(&b.counters->overflow, (uint32_t)OVF_STAGING);
        return;
    }
    // Pass 2: emit the clipped lines.
    auto emit = [&](float2 from, float2 to) {
        float l0 = from.x, l1 = from.y, l2 = to.x, l3 = to.y;
        if (clip_to_view_box(l0, l1, l2, l3, left, right, bottom)) {
            b.lines[at] = make_float4(l0, l1, l2, l3);
            b.line_meta[at] = make_uint2(path, slot);
            at++;
            slot += fill_bound(l0, l1, l2, l3);
        }
    };
    if (npt == 2) emit(c.p0, c.p3);
    else if (npt == 3) flatten<false>(c, emit);
    else flatten<true>(c, emit);
}

cudaError_t launch_dice(const BatchView &b, cudaStream_t s) {
    if (!b.segment_count) return cudaSuccess;
    k_dice<<<(b.segment_count + 127) / 128, 128, 0, s>>>(b);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ bin

struct PathTiles {
    int min_x, min_y, max_x, max_y;
    uint32_t tile_offset, backdrop_offset;
};

// ObjectBuilder::add_fill, core/d3d9/object_builder.cpp:19-64. Returns true when a fill was staged.
__device__ __forceinline__ bool add_fill(const BatchView &b, const PathTiles &pt, float fx, float fy, float tx,
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
__device__ __forceinline__ void adjust_backdrop(const BatchView &b, const PathTiles &pt, int tcx, int tcy, int delta) {
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

__global__ void __launch_bounds__(128) k_bin(BatchView b) {
    const uint32_t n_lines = min(b.counters->n_lines, b.line_capacity);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_lines; i += gridDim.x * blockDim.x) {
        const float4 ln = b.lines[i];
        const uint2 lm = b.line_meta[i];
        const uint32_t path = lm.x;
        const float l0 = ln.x, l1 = ln.y, l2 = ln.z, l3 = ln.w;
        uint32_t slot = lm.y;
        const uint32_t slot_end = slot + fill_bound(l0, l1, l2, l3);
        if (slot_end > b.staging_capacity) continue;  // dice flagged the overflow; the frame will be replayed
        PathTiles pt;
        {
            const int4 r = __ldg(reinterpret_cast<const int4 *>(&b.meta[path].tile_rect[0]));
            pt.min_x = r.x; pt.min_y = r.y; pt.max_x = r.z; pt.max_y = r.w;
            pt.tile_offset = __ldg(&b.meta[path].tile_offset);
            pt.backdrop_offset = __ldg(&b.meta[path].backdrop_offset);
        }
        // process_line_segment, tiler.cpp:155-278
        const float ts = 16.0f;
        int tcx = (int)floorf(l0 * 0.0625f), tcy = (int)floorf(l1 * 0.0625f);
        const int to_tx = (int)floorf(l2 * 0.0625f), to_ty = (int)floorf(l3 * 0.0625f);
        const float vx = l2 - l0, vy = l3 - l1;
        const int step_x = vx < 0 ? -1 : 1, step_y = vy < 0 ? -1 : 1;
        const float fcx = ((float)tcx + (vx >= 0 ? 1.0f : 0.0f)) * ts;
        const float fcy = ((float)tcy + (vy >= 0 ? 1.0f : 0.0f)) * ts;
        float t_max_x = (fcx - l0) / vx, t_max_y = (fcy - l1) / vy;
        const float t_delta_x = fabsf(ts / vx), t_delta_y = fabsf(ts / vy);
        float cur_x = l0, cur_y = l1;
        int last_dir = 0;  // 0 none, 1 X, 2 Y
        for (int iter = 0; iter < MAX_DDA_STEPS; iter++) {
            int next_dir;
            if (t_max_x < t_max_y) next_dir = 1;
            else if (t_max_x > t_max_y) next_dir = 2;
            else next_dir = step_x > 0 ? 1 : 2;
            float next_t = next_dir == 1 ? t_max_x : t_max_y;
            next_t = next_t < 1.0f ? next_t : 1.0f;
            if (tcx == to_tx && tcy == to_ty) next_dir = 0;
            const float nx = l0 + vx * next_t, ny = l1 + vy * next_t;  // LineSegmentF::sample
            if (slot + 2 > slot_end) {  // the walk left the |dx|+|dy|+1 envelope (only with non-finite arithmetic)
                atomicOr(&b.counters->overflow, (uint32_t)OVF_DDA);
                break;
            }
            if (add_fill(b, pt, cur_x, cur_y, nx, ny, tcx, tcy, slot)) slot++;
            if (step_y < 0 && next_dir == 2) {
                if (add_fill(b, pt, nx, ny, (float)tcx * ts, (float)tcy * ts, tcx, tcy, slot)) slot++;
            } else if (step_y > 0 && last_dir == 2) {
                if (add_fill(b, pt, (float)tcx * ts, (float)tcy * ts, cur_x, cur_y, tcx, tcy, slot)) slot++;
            }
            if (step_x < 0 && last_dir == 1) adjust_backdrop(b, pt, tcx, tcy, 1);
            else if (step_x > 0 && next_dir == 1) adjust_backdrop(b, pt, tcx, tcy, -1);
            if (next_dir == 1) { t_max_x += t_delta_x; tcx += step_x; }
            else if (next_dir == 2) { t_max_y += t_delta_y; tcy += step_y; }
            else break;
            cur_x = nx;
            cur_y = ny;
            last_dir = next_dir;
        }
        for (; slot < slot_end; slot++) b.staging[slot].tile = 0xffffffffu;  // unused slots
    }
}

cudaError_t launch_bin(const BatchView &b, cudaStream_t s) {
    if (!b.segment_count) return cudaSuccess;
    k_bin<<<sm_count() * 8, 128, 0, s>>>(b);
    return cudaGetLastError();
}

}  // namespace pfcu
