// Stroke-to-fill on the GPU: OutlineStrokeToFill::offset of the reference front end
// (pathfinder/core/stroke.cpp:124-167) for a whole batch of outlines in two kernel passes.
//
// With a 0.06 - 0.1 ms GPU frame the CPU front end (SVG parse + stroke: 4 ms for tiger.svg) is what an animated scene
// waits for; the stroker is the arithmetic half of it. The output must be the reference's, bit for bit (the tile
// geometry downstream is bit-exact against the reference's tiler, and a stroke that differs in the last bit moves fills):
// this file is compiled with -fmad=false and every function below performs the reference's float operations in the
// reference's order (its x86 build has no FMA; SSE lanes round like scalars).
//
// Mapping: one thread per INPUT CONTOUR. A contour's stroke is a sequential construction -- every join looks at the
// last two points pushed so far (Contour::add_join, stroke.cpp:273-324), every offset curve is accepted or split
// recursively (Segment::offset, stroke.cpp:499-540) -- so the parallelism is across contours (a glyph-density scene has
// hundreds of thousands; tiger.svg a few hundred). Pass 1 runs the construction and only counts points, a scan turns the
// counts into offsets, pass 2 runs it again and writes. Both passes are the same code (template parameter), so the counts
// cannot disagree.
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "pfcu_device.h"

namespace pfcu {
namespace stroke {

constexpr float STROKE_TOL = 0.1f;         // stroke.cpp:15
constexpr int SAMPLE_COUNT = 8;            // stroke.cpp:21
constexpr float GEOMETRIC_EPSILON = 0.001f;  // common/math/basic.h:14
constexpr int MAX_RECURSION = 16;          // stroke.cpp:514
enum Kind : int { K_NONE = 0, K_LINE = 1, K_QUAD = 2, K_CUBIC = 3 };
enum Flag : uint8_t { ON_CURVE = 0, CTRL0 = 1, CTRL1 = 2 };
enum Join : int { JOIN_MITER = 0, JOIN_BEVEL = 1, JOIN_ROUND = 2 };  // LineJoin, core/data/data.h
enum Cap : int { CAP_BUTT = 0, CAP_SQUARE = 1, CAP_ROUND = 2 };      // LineCap

struct V2 {
    float x, y;
};
__device__ __forceinline__ V2 v2(float x, float y) { return V2{x, y}; }
__device__ __forceinline__ V2 operator+(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ V2 operator-(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ V2 operator*(V2 a, V2 b) { return v2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ V2 operator*(V2 a, float s) { return v2(a.x * s, a.y * s); }
__device__ __forceinline__ V2 operator/(V2 a, float s) { return v2(a.x / s, a.y / s); }
__device__ __forceinline__ V2 neg(V2 a) { return v2(-a.x, -a.y); }
__device__ __forceinline__ float sq_len(V2 a) { return a.x * a.x + a.y * a.y; }  // Vec2::square_length
__device__ __forceinline__ float len(V2 a) { return sqrtf(sq_len(a)); }
__device__ __forceinline__ V2 normalize(V2 a) { return a / len(a); }            // Vec2::normalize, vec2.h:62-69
__device__ __forceinline__ V2 yx(V2 a) { return v2(a.y, a.x); }
__device__ __forceinline__ V2 lerp(V2 a, V2 b, float t) { return v2(a.x + t * (b.x - a.x), a.y + t * (b.y - a.y)); }  // vec2.h:178
__device__ __forceinline__ bool finite2(V2 a) { return isfinite(a.x) && isfinite(a.y); }

// Mat2 (column major: m11 m21 m12 m22) and Transform2, common/math/mat2.h, transform2.h
struct M2 {
    float v[4];
};
struct T2 {
    M2 m;
    V2 t;
};
__device__ __forceinline__ M2 mat_mul(const M2 &a, const M2 &b) {  // mat2.h:68-73
    return M2{{a.v[0] * b.v[0] + a.v[2] * b.v[1], a.v[1] * b.v[0] + a.v[3] * b.v[1], a.v[0] * b.v[2] + a.v[2] * b.v[3],
               a.v[1] * b.v[2] + a.v[3] * b.v[3]}};
}
__device__ __forceinline__ V2 mat_apply(const M2 &a, V2 p) { return v2(a.v[0] * p.x + a.v[2] * p.y, a.v[1] * p.x + a.v[3] * p.y); }
__device__ __forceinline__ V2 xf_apply(const T2 &a, V2 p) { return mat_apply(a.m, p) + a.t; }             // transform2.h:88-90
__device__ __forceinline__ T2 xf_mul(const T2 &a, const T2 &b) { return T2{mat_mul(a.m, b.m), xf_apply(a, b.t)}; }  // :104-106
__device__ __forceinline__ T2 xf_scale(V2 s) { return T2{M2{{s.x, 0.0f, 0.0f, s.y}}, v2(0.0f, 0.0f)}; }
__device__ __forceinline__ T2 xf_translation(V2 t) { return T2{M2{{1.0f, 0.0f, 0.0f, 1.0f}}, t}; }
__device__ __forceinline__ T2 xf_rotation_vector(V2 u) {  // mat2.h:32-34
    return T2{M2{{1.0f * u.x, 1.0f * u.y, -1.0f * u.y, 1.0f * u.x}}, v2(0.0f, 0.0f)};
}

// UnitVector, common/math/unit_vector.cpp
__device__ __forceinline__ V2 rotate_by(V2 a, V2 o) { return v2(a.x * o.x - a.y * o.y, a.y * o.x + a.x * o.y); }
__device__ __forceinline__ V2 rev_rotate_by(V2 a, V2 o) { return v2(a.x * o.x + a.y * o.y, a.y * o.x - a.x * o.y); }
__device__ __forceinline__ V2 halve_angle(V2 a) {
    const V2 term = v2(a.x, -a.x);
    const V2 h = v2(0.5f, 0.5f) * (v2(1.0f, 1.0f) + term);
    return v2(sqrtf(fmaxf(h.x, 0.0f)), sqrtf(fmaxf(h.y, 0.0f)));
}

// LineSegmentF, core/data/line_segment.cpp
struct L2 {
    V2 from, to;
};
__device__ __forceinline__ V2 vec(const L2 &l) { return l.to - l.from; }
__device__ __forceinline__ V2 sample(const L2 &l, float t) { return l.from + vec(l) * t; }
__device__ __forceinline__ L2 offset(const L2 &l, float distance) {  // line_segment.cpp:95-101
    const V2 v = vec(l);
    if (v.x == 0.0f && v.y == 0.0f) return l;
    const V2 d = normalize(yx(v)) * v2(-distance, distance);
    return L2{l.from + d, l.to + d};
}
__device__ __forceinline__ bool intersection_t(const L2 &a, const L2 &b, float &out) {  // line_segment.cpp:84-93
    const V2 p0p1 = vec(a), bv = vec(b);
    const float m0 = bv.x, m1 = bv.y, m2 = -p0p1.x, m3 = -p0p1.y;
    const float det = m0 * m3 - m1 * m2;
    if (fabsf(det) < FLOAT_EPSILON) return false;
    const float inv = 1.0f / det;
    // adjugate (v3 * 1, v1 * -1, v2 * -1, v0 * 1) * (1 / det); only the second row is needed
    const float i1 = (m1 * -1.0f) * inv, i3 = (m0 * 1.0f) * inv;
    const V2 d = a.from - b.from;
    out = i1 * d.x + i3 * d.y;
    return true;
}

// Segment, core/data/segment.cpp. Quadratics keep their control point in c0 (c1 is whatever the reference leaves there:
// it is never read for a quadratic).
struct Seg {
    V2 p0, c0, c1, p3;
    int kind;
};
__device__ __forceinline__ bool seg_valid(const Seg &s) {  // segment.cpp:189-201
    if (s.kind == K_LINE) return finite2(s.p0) && finite2(s.p3);
    if (s.kind == K_QUAD || s.kind == K_CUBIC) return finite2(s.p0) && finite2(s.p3) && finite2(s.c0) && finite2(s.c1);
    return false;
}
__device__ void seg_split(const Seg &s, float t, Seg &a, Seg &b) {  // segment.cpp:43-139
    a.kind = b.kind = s.kind;
    if (s.kind == K_LINE) {  // LineSegmentF::split
        const V2 mid = s.p0 + (s.p3 - s.p0) * t;
        a.p0 = s.p0; a.p3 = mid; b.p0 = mid; b.p3 = s.p3;
        a.c0 = a.c1 = b.c0 = b.c1 = v2(0.0f, 0.0f);
    } else if (s.kind == K_QUAD) {
        const V2 qa = s.p0 + (s.c0 - s.p0) * t, qb = s.c0 + (s.p3 - s.c0) * t;
        const V2 c = qa + (qb - qa) * t;
        a.p0 = s.p0; a.c0 = a.c1 = qa; a.p3 = c;
        b.p0 = c; b.c0 = b.c1 = qb; b.p3 = s.p3;
    } else if (t <= 0.0f) {
        a.p0 = a.c0 = a.c1 = a.p3 = s.p0;
        b = s;
    } else if (t >= 1.0f) {
        a = s;
        b.p0 = b.c0 = b.c1 = b.p3 = s.p3;
        b.kind = K_CUBIC;
    } else {
        const V2 p01 = s.p0 + (s.c0 - s.p0) * t, p12 = s.c0 + (s.c1 - s.c0) * t, p23 = s.c1 + (s.p3 - s.c1) * t;
        const V2 p012 = p01 + (p12 - p01) * t, p123 = p12 + (p23 - p12) * t;
        const V2 p0123 = p012 + (p123 - p012) * t;
        a.p0 = s.p0; a.c0 = p01; a.c1 = p012; a.p3 = p0123;
        b.p0 = p0123; b.c0 = p123; b.c1 = p23; b.p3 = s.p3;
    }
}
__device__ __forceinline__ V2 seg_sample(const Seg &s, float t) {  // segment.cpp:163-168
    Seg a, b;
    seg_split(s, t, a, b);
    return a.p3;
}
__device__ __forceinline__ Seg seg_transform(const Seg &s, const T2 &x) {  // segment.cpp:156-161
    return Seg{xf_apply(x, s.p0), xf_apply(x, s.c0), xf_apply(x, s.c1), xf_apply(x, s.p3), s.kind};
}
__device__ __forceinline__ V2 join_ctrl(const L2 &s0, const L2 &s1) {
    float t;
    return intersection_t(s0, s1, t) ? sample(s0, t) : lerp(s0.to, s1.from, 0.5f);
}
__device__ Seg offset_once(const Seg &s, float distance) {  // stroke.cpp:370-441
    if (s.kind == K_LINE) {
        const L2 l = offset(L2{s.p0, s.p3}, distance);
        return Seg{l.from, v2(0.0f, 0.0f), v2(0.0f, 0.0f), l.to, K_LINE};
    }
    if (s.kind == K_QUAD) {
        const L2 s0 = offset(L2{s.p0, s.c0}, distance), s1 = offset(L2{s.c0, s.p3}, distance);
        const V2 c = join_ctrl(s0, s1);
        return Seg{s0.from, c, c, s1.to, K_QUAD};
    }
    if (s.p0.x == s.c0.x && s.p0.y == s.c0.y) {
        const L2 s0 = offset(L2{s.p0, s.c1}, distance), s1 = offset(L2{s.c1, s.p3}, distance);
        const V2 c = join_ctrl(s0, s1);
        return Seg{s0.from, s0.from, c, s1.to, K_CUBIC};
    }
    if (s.c1.x == s.p3.x && s.c1.y == s.p3.y) {
        const L2 s0 = offset(L2{s.p0, s.c0}, distance), s1 = offset(L2{s.c0, s.p3}, distance);
        const V2 c = join_ctrl(s0, s1);
        return Seg{s0.from, c, s1.to, s1.to, K_CUBIC};
    }
    const L2 s0 = offset(L2{s.p0, s.c0}, distance), s1 = offset(L2{s.c0, s.c1}, distance), s2 = offset(L2{s.c1, s.p3}, distance);
    V2 c0, c1;
    float t0, t1;
    if (intersection_t(s0, s1, t0) && intersection_t(s1, s2, t1)) {
        c0 = sample(s0, t0);
        c1 = sample(s1, t1);
    } else {
        c0 = lerp(s0.to, s1.from, 0.5f);
        c1 = lerp(s1.to, s2.from, 0.5f);
    }
    return Seg{s0.from, c0, c1, s2.to, K_CUBIC};
}
__device__ bool error_within_tolerance(const Seg &s, const Seg &other, float distance) {  // stroke.cpp:454-477
    float lo = fabsf(distance) - STROKE_TOL, hi = fabsf(distance) + STROKE_TOL;
    lo = lo <= 0.0f ? 0.0f : lo * lo;
    hi = hi <= 0.0f ? 0.0f : hi * hi;
    for (int k = 0; k < SAMPLE_COUNT + 1; k++) {
        const float t = (float)k / (float)SAMPLE_COUNT;
        const float d2 = sq_len(seg_sample(s, t) - seg_sample(other, t));
        if (d2 < lo || d2 > hi) return false;
    }
    return true;
}
__device__ __forceinline__ Seg quarter_circle_arc() {  // segment.cpp:203-212
    const float r2 = sqrtf(2.0f);
    const V2 p0 = v2(r2 * 0.5f, r2 * 0.5f);
    const V2 p1 = v2(-r2 / 6.0f + 4.0f / 3.0f, 7.0f * r2 / 6.0f - 4.0f / 3.0f);
    const V2 flip = v2(1.0f, -1.0f);
    const V2 p2 = p1 * flip, p3 = p0 * flip;
    return Seg{p3, p2, p1, p0, K_CUBIC};
}
__device__ __forceinline__ Seg arc_from_cos(float c) {  // segment.cpp:214-240
    if (c >= 1.0 - GEOMETRIC_EPSILON) return Seg{v2(1.0f, 0.0f), v2(0.0f, 0.0f), v2(0.0f, 0.0f), v2(1.0f, 0.0f), K_LINE};
    const V2 term = v2(c, -c);
    const V2 h = (term + v2(1.0f, 1.0f)) * v2(0.5f, 0.5f);
    const V2 r = v2(sqrtf(h.x), sqrtf(h.y));
    const V2 p3 = r * v2(1.0f, -1.0f), p0 = r * v2(1.0f, 1.0f);
    const float p1x = 4.0f - p0.x, p1y = (1.0f - p0.x) * (3.0f - p0.x) / p0.y;
    const V2 p2 = v2(p1x, -p1y) * (1.0f / 3.0f), p1 = v2(p1x, p1y) * (1.0f / 3.0f);
    return Seg{p3, p2, p1, p0, K_CUBIC};
}

// The output contour under construction (Contour::push_point and the accessors the stroker uses). WRITE = false counts.
// The last 8 points are kept in a ring: add_join reads the last two, add_cap walks back from the end over coincident
// points (never more than a couple; both passes walk the same ring, so they agree by construction).
template <bool WRITE>
struct Out {
    float2 *pts;
    uint8_t *flags;
    uint32_t cap;  // points this contour may write (pass 2: its count from pass 1)
    uint32_t n;
    V2 ring[8];
    V2 first[2];
    T2 xform;      // Canvas::push_path's Outline::transform on the finished outline (canvas.cpp:190): applied to what is
    bool has_xform;  // WRITTEN only -- the construction itself runs in the path's own coordinates
    __device__ __forceinline__ void reset(float2 *p, uint8_t *f, uint32_t c) {
        pts = p;
        flags = f;
        cap = c;
        n = 0;
    }
    __device__ __forceinline__ void push(V2 p, uint8_t flag) {
        if (WRITE && n < cap) {
            const V2 q = has_xform ? xf_apply(xform, p) : p;  // Contour::transform, contour.cpp:100-107
            pts[n] = make_float2(q.x, q.y);
            flags[n] = flag;
        }
        ring[n & 7u] = p;
        if (n < 2) first[n] = p;
        n++;
    }
    __device__ __forceinline__ V2 last(uint32_t index) const { return ring[(n - index) & 7u]; }  // Contour::position_of_last
    __device__ __forceinline__ void push_segment(const Seg &s) {  // contour.cpp:11-41
        if (s.kind == K_NONE || !seg_valid(s)) return;
        push(s.p0, ON_CURVE);
        if (s.kind != K_LINE) {
            push(s.c0, CTRL0);
            if (s.kind != K_QUAD) push(s.c1, CTRL1);
        }
        push(s.p3, ON_CURVE);
    }
    __device__ __forceinline__ bool might_need_join(int join) const { return n >= 2 && (join == JOIN_MITER || join == JOIN_ROUND); }
};

template <bool WRITE>
__device__ void push_arc_from_unit_chord(Out<WRITE> &out, const T2 &transform, V2 chord_from, V2 chord_to) {  // stroke.cpp:326-368 (CW)
    const T2 direction = xf_scale(v2(1.0f, 1.0f));  // Transform2()
    V2 vector = chord_from;
    const V2 end_vector = chord_to;
    for (int i = 0; i < 4; i++) {
        V2 sweep = rev_rotate_by(end_vector, vector);
        const bool last = sweep.x >= -FLOAT_EPSILON && sweep.y >= -FLOAT_EPSILON;
        Seg seg;
        if (!last) {
            sweep = v2(0.0f, 1.0f);
            seg = quarter_circle_arc();
        } else {
            seg = arc_from_cos(sweep.x);
        }
        const V2 half = halve_angle(sweep);
        const T2 rotation = xf_rotation_vector(rotate_by(half, vector));
        seg = seg_transform(seg, xf_mul(xf_mul(transform, direction), rotation));
        out.push_segment(seg);
        if (last) break;
        vector = rotate_by(vector, sweep);
    }
}

template <bool WRITE>
__device__ void add_join(Out<WRITE> &out, float distance, int join, V2 join_point, const L2 &next_tangent, float miter_limit) {  // stroke.cpp:273-324
    const V2 p0 = out.last(2), p1 = out.last(1);
    const L2 prev_tangent{p0, p1};
    if (sq_len(vec(prev_tangent)) < FLOAT_EPSILON || sq_len(vec(next_tangent)) < FLOAT_EPSILON) return;
    if (join == JOIN_MITER) {
        float t;
        if (intersection_t(prev_tangent, next_tangent, t)) {
            if (t < -FLOAT_EPSILON) return;
            const V2 miter_endpoint = sample(prev_tangent, t);
            const float threshold = miter_limit * distance;
            if (sq_len(miter_endpoint - join_point) > threshold * threshold) return;
            out.push(miter_endpoint, ON_CURVE);
        }
    } else if (join == JOIN_ROUND) {
        if (sq_len(prev_tangent.to - join_point) == 0.0f || sq_len(next_tangent.to - join_point) == 0.0f) return;
        const float scale = fabsf(distance);
        const T2 transform = xf_mul(xf_translation(join_point), xf_scale(v2(scale, scale)));  // from_scale(s).translate(p)
        push_arc_from_unit_chord(out, transform, normalize(prev_tangent.to - join_point), normalize(next_tangent.to - join_point));
    }
}

template <bool WRITE>
__device__ __forceinline__ void add_to_contour(Out<WRITE> &out, const Seg &s, float distance, int join, V2 join_point, float miter_limit) {  // stroke.cpp:479-497
    if (out.might_need_join(join)) {
        const V2 p3 = s.p0, p4 = s.kind == K_LINE ? s.p3 : s.c0;
        add_join(out, distance, join, join_point, L2{p4, p3}, miter_limit);
    }
    out.push_segment(s);
}

// Segment::offset (stroke.cpp:499-540), the recursion unrolled onto an explicit stack (depth <= 16, one pending sibling per
// level). Every accepted piece -- the offset curve, or the segment itself when it is too short or too deep -- goes to
// `sink(piece, join_point)` in curve order: what Segment::add_to_contour is called with upstream.
template <typename Sink>
__device__ void for_each_leaf(const Seg &root, float distance, Sink &&sink) {
    Seg stack[MAX_RECURSION + 1];
    unsigned char depth_of[MAX_RECURSION + 1];
    int sp = 0;
    Seg cur = root;
    int depth = 0;
    while (true) {
        bool leaf = true;
        if (seg_valid(cur)) {
            const V2 join_point = cur.p0;
            if (sq_len(cur.p3 - cur.p0) < STROKE_TOL * STROKE_TOL || depth >= MAX_RECURSION) {
                sink(cur, join_point);
            } else {
                const Seg candidate = offset_once(cur, distance);
                if (cur.kind == K_LINE || error_within_tolerance(cur, candidate, distance)) {
                    sink(candidate, join_point);
                } else {
                    Seg before, after;
                    seg_split(cur, 0.5f, before, after);
                    depth++;
                    stack[sp] = after;
                    depth_of[sp] = (unsigned char)depth;
                    sp++;
                    cur = before;
                    leaf = false;
                }
            }
        }
        if (leaf) {
            if (sp == 0) break;
            sp--;
            cur = stack[sp];
            depth = depth_of[sp];
        }
    }
}

struct ContourIn {
    const float2 *pts;
    const uint8_t *flags;
    int n;
    bool closed;
    __device__ __forceinline__ V2 p(int i) const { return v2(pts[i].x, pts[i].y); }
};

__device__ __forceinline__ bool approx_eq(V2 a, V2 b, float eps) { return len(a - b) <= eps; }  // vec2.h:87-89

// SegmentsIter (contour.cpp:108-177)
struct SegIter {
    int head = 0;
    bool has_next = true;
    __device__ Seg next(const ContourIn &c) {
        Seg s;
        s.kind = K_NONE;
        s.p0 = s.c0 = s.c1 = s.p3 = v2(0.0f, 0.0f);
        if (head < c.n && c.flags[head] == ON_CURVE) {
            s.p0 = c.p(head);
            if (head + 1 < c.n) {
                if (c.flags[head + 1] == ON_CURVE) {
                    s.p3 = c.p(head + 1);
                    s.kind = K_LINE;
                    head += 1;
                } else if (head + 2 < c.n) {
                    if (c.flags[head + 1] == CTRL0 && c.flags[head + 2] == ON_CURVE) {
                        s.c0 = c.p(head + 1);
                        s.p3 = c.p(head + 2);
                        s.kind = K_QUAD;
                        head += 2;
                    } else if (head + 3 < c.n) {
                        if (c.flags[head + 1] == CTRL0 && c.flags[head + 2] == CTRL1 && c.flags[head + 3] == ON_CURVE) {
                            s.c0 = c.p(head + 1);
                            s.c1 = c.p(head + 2);
                            s.p3 = c.p(head + 3);
                            s.kind = K_CUBIC;
                            head += 3;
                        }
                    }
                }
            } else {
                if (c.closed && c.n > 1 && !approx_eq(c.p(0), c.p(c.n - 1), FLOAT_EPSILON)) {
                    s.p0 = c.p(c.n - 1);
                    s.p3 = c.p(0);
                    s.kind = K_LINE;
                }
                has_next = false;
            }
        }
        return s;
    }
};

// The segments ContourStrokeToFill::offset_forward (stroke.cpp:27-51) offsets, in order: sink(segment)
template <typename Sink>
__device__ void forward_segments(const ContourIn &c, Sink &&sink) {
    SegIter it;
    while (it.has_next) {
        const Seg s = it.next(c);
        if (s.kind == K_NONE) break;
        sink(s);
    }
}

// ... and ContourStrokeToFill::offset_backward (stroke.cpp:53-122): the closing line first, then the contour back to front
template <typename Sink>
__device__ void backward_segments(const ContourIn &c, Sink &&sink) {
    int tail = c.n - 1;
    if (c.closed && c.n > 1 && !approx_eq(c.p(0), c.p(c.n - 1), FLOAT_EPSILON))
        sink(Seg{c.p(0), v2(0.0f, 0.0f), v2(0.0f, 0.0f), c.p(c.n - 1), K_LINE});
    while (tail >= 0) {
        if (c.flags[tail] != ON_CURVE) break;
        Seg s;
        if (tail >= 3 && c.flags[tail - 1] == CTRL1 && c.flags[tail - 2] == CTRL0 && c.flags[tail - 3] == ON_CURVE) {
            s = Seg{c.p(tail), c.p(tail - 1), c.p(tail - 2), c.p(tail - 3), K_CUBIC};
            tail -= 3;
        } else if (tail >= 2 && c.flags[tail - 1] == CTRL0 && c.flags[tail - 2] == ON_CURVE) {
            s = Seg{c.p(tail), c.p(tail - 1), v2(0.0f, 0.0f), c.p(tail - 2), K_QUAD};
            tail -= 2;
        } else if (tail >= 1 && c.flags[tail - 1] == ON_CURVE) {
            s = Seg{c.p(tail), v2(0.0f, 0.0f), v2(0.0f, 0.0f), c.p(tail - 1), K_LINE};
            tail -= 1;
        } else {
            break;
        }
        sink(s);
    }
}

template <bool WRITE>
__device__ void add_cap(Out<WRITE> &out, float width, int cap) {  // stroke.cpp:197-255
    if (cap == CAP_BUTT || out.n < 2) return;
    const V2 p1 = out.last(1);
    V2 p0;
    uint32_t back = 2;  // p0_index = size - back
    while (true) {
        p0 = out.last(back);
        if (sq_len(p1 - p0) > FLOAT_EPSILON) break;
        if (back == out.n || back == 8) return;  // p0_index == 0 (or the ring's reach: 8 coincident points)
        back++;
    }
    const V2 gradient = normalize(p1 - p0);
    if (cap == CAP_SQUARE) {
        const V2 off = gradient * (width * 0.5f);
        const V2 p2 = p1 + off;
        const V2 p3 = p2 + yx(gradient) * v2(-width, width);
        const V2 p4 = p3 - off;
        out.push(p2, ON_CURVE);
        out.push(p3, ON_CURVE);
        out.push(p4, ON_CURVE);
    } else {
        const float scale = width * 0.5f;
        const V2 off = yx(gradient) * v2(-1.0f, 1.0f);
        const V2 translation = p1 + off * (width * 0.5f);
        const T2 transform = xf_mul(xf_translation(translation), xf_scale(v2(scale, scale)));
        push_arc_from_unit_chord(out, transform, neg(off), off);
    }
}

template <bool WRITE>
__device__ void close_stroked(Out<WRITE> &out, const ContourIn &c, float width, int join, float miter_limit, bool closed) {  // stroke.cpp:178-195
    if (closed && out.might_need_join(join)) {
        const L2 final_segment{out.first[1], out.first[0]};
        add_join(out, width * 0.5f, join, c.p(0), final_segment, miter_limit);
    }
}

// ---- dashing (OutlineDash / ContourDash, pathfinder/core/dash.cpp:1-126): one thread per OUTLINE -- the dash state runs
// on from one contour of an outline into the next

__device__ __forceinline__ float arc_length(const Seg &s) {  // segment.cpp:170-187 (to_cubic: :141-154)
    if (s.kind == K_LINE) return len(s.p3 - s.p0);
    if (s.kind != K_QUAD && s.kind != K_CUBIC) return 0.0f;
    V2 c0 = s.c0, c1 = s.c1;
    if (s.kind == K_QUAD) {
        const V2 p1_2 = s.c0 + s.c0;
        c0 = (s.p0 + p1_2) / 3.0f;
        c1 = (p1_2 + s.p3) / 3.0f;
    }
    const float chord = len(s.p3 - s.p0);
    const float cont_net = len(s.p0 - c0) + len(c1 - c0) + len(s.p3 - c1);
    return (cont_net + chord) * 0.5f;
}

// The dashed outline under construction: the current contour (DashState::output) and the contours already pushed.
template <bool WRITE>
struct DashOut {
    float2 *pts;          // this outline's points
    uint8_t *flags;
    uint32_t *contour_end;  // this outline's contours: end offset (relative to the outline's first point) of each
    uint32_t n_points = 0, n_contours = 0, contour_start = 0, cap_points = 0, cap_contours = 0;
    __device__ __forceinline__ void push(V2 p, uint8_t flag) {
        if (WRITE && n_points < cap_points) {
            pts[n_points] = make_float2(p.x, p.y);
            flags[n_points] = flag;
        }
        n_points++;
    }
    __device__ __forceinline__ void push_segment(const Seg &s) {  // contour.cpp:11-41
        if (s.kind == K_NONE || !seg_valid(s)) return;
        push(s.p0, ON_CURVE);
        if (s.kind != K_LINE) {
            push(s.c0, CTRL0);
            if (s.kind != K_QUAD) push(s.c1, CTRL1);
        }
        push(s.p3, ON_CURVE);
    }
    __device__ __forceinline__ void push_contour() {  // Outline::push_contour (path.cpp:24-27 drops an empty contour)
        if (n_points == contour_start) return;
        if (WRITE && n_contours < cap_contours) contour_end[n_contours] = n_points;
        n_contours++;
        contour_start = n_points;
    }
};

template <bool WRITE>
__device__ void dash_outline(DashOut<WRITE> &out, const float2 *pts, const uint8_t *flags, const uint32_t *contour_first,
                             const uint8_t *closed, uint32_t c0, uint32_t c1, const float *dashes, uint32_t n_dashes, float offset) {
    if (!n_dashes) return;
    // DashState::DashState, dash.cpp:9-29
    float total = 0.0f;
    for (uint32_t k = 0; k < n_dashes; k++) total += dashes[k];
    offset = fmodf(offset, total);
    uint32_t index = 0;
    while (index < n_dashes) {
        const float d = dashes[index];
        if (offset < d) break;
        offset -= d;
        index += 1;
    }
    float distance_left = offset;
    // (an offset that runs off the end of the pattern leaves index == n_dashes: the reference then reads past its vector;
    // a Canvas always passes offset 0, core/canvas.cpp:288)
    for (uint32_t ci = c0; ci < c1; ci++) {  // ContourDash::dash, dash.cpp:70-124
        const uint32_t first = contour_first[ci];
        const ContourIn c{pts + first, flags + first, (int)(contour_first[ci + 1] - first), closed[ci] != 0};
        SegIter it;
        Seg queued;
        bool queued_none = true;
        while (true) {
            if (queued_none) {
                if (!it.has_next) break;
                queued = it.next(c);
                if (queued.kind == K_NONE) break;
                queued_none = false;
            }
            Seg current = queued;
            float distance = distance_left;
            const float t = distance / arc_length(current);
            if (t < 1.0) {
                Seg prev, next;
                seg_split(current, t, prev, next);
                current = prev;
                queued = next;
                queued_none = false;
            } else {
                distance = arc_length(current);
                queued_none = true;
            }
            const bool on = (index % 2u) == 0u;
            if (on) out.push_segment(current);
            distance_left -= distance;
            if (distance_left < FLOAT_EPSILON) {
                if (on) out.push_contour();
                index += 1;
                if (index == n_dashes) index = 0;
                distance_left = dashes[index];
            }
        }
    }
    if ((index % 2u) == 0u) out.push_contour();  // OutlineDash::into_outline, dash.cpp:59-65
}

}  // namespace stroke

using namespace stroke;

// One thread per outline. counts[2o] = points, counts[2o + 1] = contours of the dashed outline o (pass 1); pass 2 writes
// at the exclusive prefix sums of both (offsets[], interleaved the same way: one scan over 2 * n_outlines words would mix
// the two, so the points live at even and the contours at odd positions of two separate arrays -- see the launcher).
template <bool WRITE>
__global__ void __launch_bounds__(128) k_dash(const float2 *pts, const uint8_t *flags, const uint32_t *contour_first, const uint8_t *closed,
                                              const uint32_t *outline_first, const float *dashes, const uint32_t *dash_first,
                                              const float *dash_offset, uint32_t n_outlines, uint32_t *point_counts,
                                              uint32_t *contour_counts, const uint32_t *point_offsets, const uint32_t *contour_offsets,
                                              float2 *out_pts, uint8_t *out_flags, uint32_t *out_contour_first) {
    const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_outlines) return;
    DashOut<WRITE> out;
    if (WRITE) {
        out.pts = out_pts + point_offsets[o];
        out.flags = out_flags + point_offsets[o];
        out.contour_end = out_contour_first + contour_offsets[o] + 1;  // entry k + 1 = end of contour k
        out.cap_points = point_counts[o];
        out.cap_contours = contour_counts[o];
    }
    dash_outline(out, pts, flags, contour_first, closed, outline_first[o], outline_first[o + 1], dashes + dash_first[o],
                 dash_first[o + 1] - dash_first[o], dash_offset[o]);
    if (!WRITE) {
        point_counts[o] = out.n_points;
        contour_counts[o] = out.n_contours;
    } else {
        // contour ends were written relative to the outline's first point: make them absolute
        for (uint32_t k = 0; k < out.cap_contours; k++) out_contour_first[contour_offsets[o] + 1 + k] += point_offsets[o];
        if (o == 0) out_contour_first[0] = 0;
    }
}

// ---- stroke-to-fill in three levels of parallelism
//   segments : one thread per input contour lists the segments its two sides offset (forward_segments / backward_segments)
//   leaves   : one thread per SEGMENT runs Segment::offset's recursion -- offset_once + the 9-sample error check per node,
//              nine tenths of the stroker's arithmetic -- and records the accepted pieces ("leaves"); count pass, scan,
//              write pass
//   contours : one thread per input contour walks its leaves in order and does what is inherently sequential: joins (they
//              look at the last two points pushed so far), caps, the pushes themselves; count pass, scan, write pass
// (Round 2's first version ran everything in the per-contour thread: 1.4 ms for tiger.svg's 78 stroked contours, whose
// longest has 715 segments, against 0.28 ms for the reference on one CPU core.)

struct SegSlot {
    Seg s;           // kind K_NONE: unused slot (the used ones of a side come first)
    float distance;  // -radius of the contour's style
};
struct Leaf {
    float4 a, b, c;  // (p0, c0), (c1, p3), (join point, kind, -)
};

// seg_first[i] .. seg_first[i + 1]: the slots of contour i, first half its forward side, second half its backward side (the
// host sizes a side as on-curve points + 1: a segment starts at an on-curve point, plus the closing line)
__global__ void __launch_bounds__(128) k_stroke_segments(const float2 *pts, const uint8_t *flags, const uint32_t *contour_first,
                                                         const uint8_t *closed, const uint32_t *style_index,
                                                         const pfcu_stroke_style *styles, const uint32_t *seg_first,
                                                         uint32_t n_contours, SegSlot *slots) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_contours) return;
    const uint32_t first = contour_first[i];
    const ContourIn c{pts + first, flags + first, (int)(contour_first[i + 1] - first), closed[i] != 0};
    const float distance = -(styles[style_index[i]].line_width * 0.5f);
    const uint32_t cap = (seg_first[i + 1] - seg_first[i]) / 2;
    for (int side = 0; side < 2; side++) {
        SegSlot *out = slots + seg_first[i] + side * cap;
        uint32_t k = 0;
        auto sink = [&](const Seg &sg) {
            if (k < cap) out[k] = SegSlot{sg, distance};
            k++;
        };
        if (side == 0) forward_segments(c, sink);
        else backward_segments(c, sink);
        for (; k < cap; k++) out[k].s.kind = K_NONE;
    }
}

template <bool WRITE>
__global__ void __launch_bounds__(128) k_stroke_leaves(const SegSlot *slots, uint32_t n_slots, uint32_t *leaf_count,
                                                       const uint32_t *leaf_offset, Leaf *leaves) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    const SegSlot sl = slots[i];
    uint32_t n = 0;
    if (sl.s.kind != K_NONE) {
        Leaf *out = WRITE ? leaves + leaf_offset[i] : nullptr;
        const uint32_t cap = WRITE ? leaf_count[i] : 0u;
        for_each_leaf(sl.s, sl.distance, [&](const Seg &piece, V2 join_point) {
            if (WRITE && n < cap)
                out[n] = Leaf{make_float4(piece.p0.x, piece.p0.y, piece.c0.x, piece.c0.y),
                              make_float4(piece.c1.x, piece.c1.y, piece.p3.x, piece.p3.y),
                              make_float4(join_point.x, join_point.y, __int_as_float(piece.kind), 0.0f)};
            n++;
        });
    }
    if (!WRITE) leaf_count[i] = n;
}

// One side of a contour: Segment::add_to_contour for every leaf of every segment, in order (the first segment's leaves get
// a bevel join: stroke.cpp:46, :76, :118)
template <bool WRITE>
__device__ void add_side(Out<WRITE> &out, const SegSlot *slots, uint32_t base, uint32_t cap, const uint32_t *leaf_count,
                         const uint32_t *leaf_offset, const Leaf *leaves, int join, float miter_limit) {
    for (uint32_t k = 0; k < cap; k++) {
        if (slots[base + k].s.kind == K_NONE) break;
        const float distance = slots[base + k].distance;
        const int jn = k == 0 ? (int)JOIN_BEVEL : join;
        const uint32_t l0 = leaf_offset[base + k], l1 = l0 + leaf_count[base + k];
        for (uint32_t l = l0; l < l1; l++) {
            const Leaf lf = leaves[l];
            const Seg piece{v2(lf.a.x, lf.a.y), v2(lf.a.z, lf.a.w), v2(lf.b.x, lf.b.y), v2(lf.b.z, lf.b.w), __float_as_int(lf.c.z)};
            add_to_contour(out, piece, distance, jn, v2(lf.c.x, lf.c.y), miter_limit);
        }
    }
}

// One thread per input contour (OutlineStrokeToFill::offset's loop body, stroke.cpp:130-155). counts[2i], counts[2i+1]:
// points of the contour's first / second output contour (closed contours produce two, open ones one: the second stays 0).
// WRITE: offsets[] are the exclusive prefix sums of counts[].
template <bool WRITE>
__global__ void __launch_bounds__(128) k_stroke(const float2 *pts, const uint8_t *flags, const uint32_t *contour_first, const uint8_t *closed,
                                                const uint32_t *style_index, const pfcu_stroke_style *styles, const uint32_t *seg_first,
                                                const SegSlot *slots, const uint32_t *leaf_count, const uint32_t *leaf_offset,
                                                const Leaf *leaves, uint32_t n_contours, uint32_t *counts, const uint32_t *offsets,
                                                float2 *out_pts, uint8_t *out_flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_contours) return;
    const uint32_t first = contour_first[i];
    ContourIn c{pts + first, flags + first, (int)(contour_first[i + 1] - first), closed[i] != 0};
    const pfcu_stroke_style st = styles[style_index[i]];
    const uint32_t base = seg_first[i], cap = (seg_first[i + 1] - seg_first[i]) / 2;
    Out<WRITE> out;
    out.xform = T2{M2{{st.transform[0], st.transform[1], st.transform[2], st.transform[3]}}, v2(st.transform[4], st.transform[5])};
    // Outline::transform leaves an identity alone (path.cpp:8-10)
    out.has_xform = !(st.transform[0] == 1.0f && st.transform[1] == 0.0f && st.transform[2] == 0.0f && st.transform[3] == 1.0f &&
                      st.transform[4] == 0.0f && st.transform[5] == 0.0f);
    out.reset(WRITE ? out_pts + offsets[2 * i] : nullptr, WRITE ? out_flags + offsets[2 * i] : nullptr, WRITE ? counts[2 * i] : 0u);
    add_side(out, slots, base, cap, leaf_count, leaf_offset, leaves, st.line_join, st.miter_limit);
    uint32_t n_first = 0;
    if (c.closed) {
        close_stroked(out, c, st.line_width, st.line_join, st.miter_limit, true);
        n_first = out.n;
        out.reset(WRITE ? out_pts + offsets[2 * i + 1] : nullptr, WRITE ? out_flags + offsets[2 * i + 1] : nullptr, WRITE ? counts[2 * i + 1] : 0u);
    } else {
        add_cap(out, st.line_width, st.line_cap);
    }
    add_side(out, slots, base + cap, cap, leaf_count, leaf_offset, leaves, st.line_join, st.miter_limit);
    if (!c.closed) add_cap(out, st.line_width, st.line_cap);
    close_stroked(out, c, st.line_width, st.line_join, st.miter_limit, c.closed);
    if (!WRITE) {
        counts[2 * i] = c.closed ? n_first : out.n;
        counts[2 * i + 1] = c.closed ? out.n : 0u;
    }
}

// ---- exclusive scan of n words in three kernels (CTA-local scan of 4096 words + CTA totals; scan of the totals by one
// CTA; add): the counts of a glyph-density scene are millions of words. total[0] = sum.
constexpr int SCAN_BLOCK = 1024, SCAN_PER_THREAD = 4, SCAN_CHUNK = SCAN_BLOCK * SCAN_PER_THREAD;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_sums, uint32_t &block_total) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (unsigned)d) incl += t;
    }
    __syncthreads();  // (warp_sums may still be read from an earlier call)
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = warp_sums[lane];
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= (unsigned)d) wi += t;
        }
        warp_sums[lane] = wi - w;
        if (lane == 31) warp_sums[32] = wi;
    }
    __syncthreads();
    block_total = warp_sums[32];
    return warp_sums[warp] + incl - v;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_chunks(const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *chunk_totals) {
    __shared__ uint32_t warp_sums[33];
    const uint32_t base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_PER_THREAD;
    uint32_t v[SCAN_PER_THREAD], sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        v[k] = base + k < n ? in[base + k] : 0u;
        sum += v[k];
    }
    uint32_t total;
    uint32_t off = block_exclusive_scan(sum, warp_sums, total);
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        if (base + k < n) out[base + k] = off;
        off += v[k];
    }
    if (threadIdx.x == 0) chunk_totals[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_totals(uint32_t *chunk_totals, uint32_t n_chunks, uint32_t *total) {
    __shared__ uint32_t warp_sums[33];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_chunks; base += SCAN_BLOCK) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_chunks ? chunk_totals[i] : 0u;
        uint32_t block_total;
        const uint32_t off = block_exclusive_scan(v, warp_sums, block_total);
        if (i < n_chunks) chunk_totals[i] = carry + off;
        __syncthreads();
        if (threadIdx.x == 0) carry += block_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_add(uint32_t *out, uint32_t n, const uint32_t *chunk_totals) {
    const uint32_t add = chunk_totals[blockIdx.x];
    const uint32_t base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_PER_THREAD;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++)
        if (base + k < n) out[base + k] += add;
}

// scratch: (n + SCAN_CHUNK - 1) / SCAN_CHUNK words
cudaError_t launch_scan_u32(const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *total, uint32_t *scratch, cudaStream_t s) {
    const uint32_t n_chunks = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    if (!n_chunks) return cudaMemsetAsync(total, 0, 4, s);
    k_scan_chunks<<<n_chunks, SCAN_BLOCK, 0, s>>>(in, out, n, scratch);
    k_scan_totals<<<1, SCAN_BLOCK, 0, s>>>(scratch, n_chunks, total);
    if (n_chunks > 1) k_scan_add<<<n_chunks, SCAN_BLOCK, 0, s>>>(out, n, scratch);
    return cudaGetLastError();
}

cudaError_t launch_stroke_segments(const StrokeArgs &a, cudaStream_t s) {
    if (!a.n_contours) return cudaSuccess;
    k_stroke_segments<<<(a.n_contours + 127) / 128, 128, 0, s>>>(a.pts, a.flags, a.contour_first, a.closed, a.style_index, a.styles,
                                                                 a.seg_first, a.n_contours, static_cast<SegSlot *>(a.slots));
    if (a.n_slots) {
        k_stroke_leaves<false><<<(a.n_slots + 127) / 128, 128, 0, s>>>(static_cast<const SegSlot *>(a.slots), a.n_slots, a.leaf_count,
                                                                      nullptr, nullptr);
        cudaError_t e = launch_scan_u32(a.leaf_count, a.leaf_offset, a.n_slots, a.total, a.scratch, s);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

cudaError_t launch_stroke_count(const StrokeArgs &a, cudaStream_t s) {
    if (!a.n_contours) return cudaSuccess;
    if (a.n_slots)
        k_stroke_leaves<true><<<(a.n_slots + 127) / 128, 128, 0, s>>>(static_cast<const SegSlot *>(a.slots), a.n_slots, a.leaf_count,
                                                                     a.leaf_offset, static_cast<Leaf *>(a.leaves));
    k_stroke<false><<<(a.n_contours + 127) / 128, 128, 0, s>>>(a.pts, a.flags, a.contour_first, a.closed, a.style_index, a.styles,
                                                             a.seg_first, static_cast<const SegSlot *>(a.slots), a.leaf_count,
                                                             a.leaf_offset, static_cast<const Leaf *>(a.leaves), a.n_contours, a.counts,
                                                             nullptr, nullptr, nullptr);
    return launch_scan_u32(a.counts, a.offsets, 2 * a.n_contours, a.total, a.scratch, s);
}

cudaError_t launch_stroke_write(const StrokeArgs &a, float2 *out_pts, uint8_t *out_flags, cudaStream_t s) {
    if (!a.n_contours) return cudaSuccess;
    k_stroke<true><<<(a.n_contours + 127) / 128, 128, 0, s>>>(a.pts, a.flags, a.contour_first, a.closed, a.style_index, a.styles,
                                                            a.seg_first, static_cast<const SegSlot *>(a.slots), a.leaf_count,
                                                            a.leaf_offset, static_cast<const Leaf *>(a.leaves), a.n_contours, a.counts,
                                                            a.offsets, out_pts, out_flags);
    return cudaGetLastError();
}

size_t stroke_slot_bytes() { return sizeof(SegSlot); }
size_t stroke_leaf_bytes() { return sizeof(Leaf); }

cudaError_t launch_dash_count(const float2 *pts, const uint8_t *flags, const uint32_t *contour_first, const uint8_t *closed,
                              const uint32_t *outline_first, const float *dashes, const uint32_t *dash_first, const float *dash_offset,
                              uint32_t n_outlines, uint32_t *point_counts, uint32_t *contour_counts, uint32_t *point_offsets,
                              uint32_t *contour_offsets, uint32_t *totals, uint32_t *scratch, cudaStream_t s) {
    if (!n_outlines) return cudaSuccess;
    k_dash<false><<<(n_outlines + 127) / 128, 128, 0, s>>>(pts, flags, contour_first, closed, outline_first, dashes, dash_first, dash_offset,
                                                           n_outlines, point_counts, contour_counts, nullptr, nullptr, nullptr, nullptr, nullptr);
    cudaError_t e = launch_scan_u32(point_counts, point_offsets, n_outlines, totals, scratch, s);
    if (e != cudaSuccess) return e;
    return launch_scan_u32(contour_counts, contour_offsets, n_outlines, totals + 1, scratch, s);
}

cudaError_t launch_dash_write(const float2 *pts, const uint8_t *flags, const uint32_t *contour_first, const uint8_t *closed,
                              const uint32_t *outline_first, const float *dashes, const uint32_t *dash_first, const float *dash_offset,
                              uint32_t n_outlines, uint32_t *point_counts, uint32_t *contour_counts, const uint32_t *point_offsets,
                              const uint32_t *contour_offsets, float2 *out_pts, uint8_t *out_flags, uint32_t *out_contour_first,
                              cudaStream_t s) {
    if (!n_outlines) return cudaSuccess;
    k_dash<true><<<(n_outlines + 127) / 128, 128, 0, s>>>(pts, flags, contour_first, closed, outline_first, dashes, dash_first, dash_offset,
                                                          n_outlines, point_counts, contour_counts, point_offsets, contour_offsets, out_pts,
                                                          out_flags, out_contour_first);
    return cudaGetLastError();
}

}  // namespace pfcu
